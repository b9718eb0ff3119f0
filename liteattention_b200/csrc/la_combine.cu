// la_combine.cu -- merge partial attention results by their log-sum-exp (HBM-bound, elementwise).
//
// The reference tells callers to split text/video attention into several LiteAttention calls with
// return_softmax_lse=True and "combine the partial results using their LSE values" (README.md:222-250) but
// ships no combiner in the default build (flash_fwd_combine_kernel.h is compiled out, hopper/setup.py:48; its op
// `lite_attention::fwd_combine`, flash_api.cpp:1620-1720, takes fp32 partials).
//   lse = log(sum_i exp(lse_i));   out = sum_i exp(lse_i - lse) * o_i
// One thread owns 8 consecutive head-dim elements of one (b, s, h) row; partials may be bf16 (what LiteAttention
// returns) or fp32 (the reference op's contract), the result bf16 or fp32.
#include "la_kernels.h"

namespace la {

template <typename In, typename Out>
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
          if constexpr (sizeof(In) == 2) {
            const uint4 raw = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(args.o_parts[p]) + row * args.d + vec * 8);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h2[j]);
              acc[2 * j] += w * f.x;
              acc[2 * j + 1] += w * f.y;
            }
          } else {
            const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(args.o_parts[p]) + row * args.d + vec * 8);
            const float4 a = src[0], b4 = src[1];
            acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
            acc[4] += w * b4.x; acc[5] += w * b4.y; acc[6] += w * b4.z; acc[7] += w * b4.w;
          }
        }
      }
    }
    const float inv = (wsum > 0.f) ? 1.0f / wsum : 0.f;
    if constexpr (sizeof(Out) == 2) {
      uint4 outv;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
      for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(acc[2 * j] * inv, acc[2 * j + 1] * inv);
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) + row * args.d + vec * 8) = outv;
    } else {
      float4* dst = reinterpret_cast<float4*>(static_cast<float*>(args.out) + row * args.d + vec * 8);
      dst[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
      dst[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
    }
    if (vec == 0 && args.lse != nullptr) args.lse[lse_idx] = (wsum > 0.f) ? lmax + logf(wsum) : -INFINITY;
  }
}

template __global__ void la_combine_kernel<__nv_bfloat16, __nv_bfloat16>(const CombineKernelArgs);
template __global__ void la_combine_kernel<__nv_bfloat16, float>(const CombineKernelArgs);
template __global__ void la_combine_kernel<float, __nv_bfloat16>(const CombineKernelArgs);
template __global__ void la_combine_kernel<float, float>(const CombineKernelArgs);

}  // namespace la
