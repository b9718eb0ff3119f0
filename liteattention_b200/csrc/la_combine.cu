// la_combine.cu -- merge partial attention results by their log-sum-exp (HBM-bound, elementwise).
//
// The reference tells callers to split text/video attention into several LiteAttention calls with
// return_softmax_lse=True and "combine the partial results using their LSE values" (README.md:222-250) but
// ships no combiner in the default build (flash_fwd_combine_kernel.h is compiled out, hopper/setup.py:48).
//   lse = log(sum_i exp(lse_i));   out = sum_i exp(lse_i - lse) * o_i
// One thread owns 8 consecutive head-dim elements (one 16-byte vector) of one (b, s, h) row.
#include "la_kernels.h"

namespace la {

__global__ void __launch_bounds__(256) la_combine_kernel(const CombineKernelArgs args) {
  const int vec_per_row = args.d >> 3;
  const int64_t total = (int64_t)args.b * args.s * args.h * vec_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int vec = (int)(idx % vec_per_row);
    const int64_t row = idx / vec_per_row;  // (b, s, h) flattened, h fastest
    const int hh = (int)(row % args.h);
    const int64_t bs = row / args.h;
    const int ss = (int)(bs % args.s);
    const int bb = (int)(bs / args.s);
    const int64_t lse_idx = ((int64_t)bb * args.h + hh) * args.s + ss;

    float lmax = -INFINITY;
    float li[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      li[p] = (p < args.n_parts) ? args.lse_parts[p][lse_idx] : -INFINITY;
      lmax = fmaxf(lmax, li[p]);
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float wsum = 0.f;
    if (lmax != -INFINITY) {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        if (p < args.n_parts) {
          const float w = __expf(li[p] - lmax);
          wsum += w;
          const uint4 raw = *reinterpret_cast<const uint4*>(args.o_parts[p] + row * args.d + vec * 8);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            acc[2 * j] += w * f.x;
            acc[2 * j + 1] += w * f.y;
          }
        }
      }
    }
    const float inv = (wsum > 0.f) ? 1.0f / wsum : 0.f;
    uint4 outv;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(acc[2 * j] * inv, acc[2 * j + 1] * inv);
    *reinterpret_cast<uint4*>(args.out + row * args.d + vec * 8) = outv;
    if (vec == 0 && args.lse != nullptr) args.lse[lse_idx] = (wsum > 0.f) ? lmax + logf(wsum) : -INFINITY;
  }
}

}  // namespace la
