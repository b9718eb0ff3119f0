// la_rope_cast.cu -- the caller-side step in front of the attention call, fused:
// 3-D rotary position embedding of Q or K + cast to bf16, one HBM pass instead of the ~10 elementwise passes the
// reference's Wan integration runs (README.md:301-315 of the reference: `rope_apply(q, grid_sizes, freqs)` then
// `.bfloat16()`; rope_apply itself lives in Wan2.1's wan/modules/model.py, not in the reference tree -- its
// published algorithm is restated in oracle/rope.py).
//
//   token t of sample b sits at (f, y, x) = (t / (gh*gw), (t / gw) % gh, t % gw) of that sample's (gf, gh, gw) grid;
//   complex pair c of the head dim takes its angle from axis 0 (frames) for c < c0, axis 1 (height) for
//   c < c0 + c1, axis 2 (width) otherwise, with c0 = D/2 - 2*(D/2/3), c1 = c2 = D/2/3 (22/21/21 at D = 128);
//   (out[2c], out[2c+1]) = (x[2c] + i x[2c+1]) * (cos + i sin);  tokens >= gf*gh*gw are only cast.
//
// HBM-bound: 4 (or 2) bytes read + 2 bytes written per element, cos/sin (D/2 values per token, shared by all heads)
// come from a 512 KB table that lives in L2.  One thread = 8 consecutive elements (4 complex pairs) of a row:
// 32-byte (fp32) or 16-byte (bf16) loads, 16-byte stores, fully coalesced.
#include "la_kernels.h"

namespace la {

// One CTA per token (b, t): the token's D/2 (cos, sin) pairs are shared by all heads, so each thread fetches the 4 it
// needs once (L1/L2 hits) and then streams its 8-element vector of every 20th head.  No per-element index division:
// the (frame, y, x) position is block-uniform scalar arithmetic.
constexpr int kRopeThreads = 320;   // 20 heads x 16 vectors of 8 elements in flight per trip (d = 128)

template <typename In>
__global__ void __launch_bounds__(kRopeThreads) la_rope_cast_kernel(const RopeKernelArgs args) {
  const int vec_per_row = args.d >> 3;
  const int half = args.d >> 1;
  const int c1 = half / 3, c0 = half - 2 * c1;
  const int heads_per_trip = kRopeThreads / vec_per_row;
  const int vec = threadIdx.x % vec_per_row;
  const int h0 = threadIdx.x / vec_per_row;
  if (h0 >= heads_per_trip) return;
  for (int64_t token = blockIdx.x; token < (int64_t)args.b * args.s; token += gridDim.x) {
    const int t = (int)(token % args.s);
    const int bb = (int)(token / args.s);
    const int gf = args.grid[bb * 3 + 0], gh = args.grid[bb * 3 + 1], gw = args.grid[bb * 3 + 2];
    const bool rotate = t < gf * gh * gw;
    float2 cs[4];
    if (rotate) {
      const int pf = t / (gh * gw), py = (t / gw) % gh, px = t % gw;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = vec * 4 + j;                      // complex pair index in [0, D/2)
        const int pos = (c < c0) ? pf : ((c < c0 + c1) ? py : px);
        cs[j] = __ldg(args.cos_sin + (int64_t)pos * half + c);   // table [max_pos, D/2] of (cos, sin)
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) cs[j] = make_float2(1.f, 0.f);
    }
    const In* src_tok = reinterpret_cast<const In*>(args.x) + (int64_t)bb * args.x_batch_stride +
                        (int64_t)t * args.x_row_stride + vec * 8;
    __nv_bfloat16* dst_tok = args.out + token * (int64_t)args.h * args.d + vec * 8;
    // Two heads per trip: both rows' loads are issued before either is rotated (bytes in flight, not math, is what
    // an HBM-bound pass is short of).
    for (int hh = h0; hh < args.h; hh += 2 * heads_per_trip) {
      const int hb = hh + heads_per_trip;
      const bool two = hb < args.h;
      float v[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const In* src = src_tok + (int64_t)(u == 0 ? hh : hb) * args.x_head_stride;
        if constexpr (sizeof(In) == 4) {
          const float4 a = __ldcs(reinterpret_cast<const float4*>(src));        // streamed once: evict-first
          const float4 b4 = __ldcs(reinterpret_cast<const float4*>(src + 4));
          v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
          v[u][4] = b4.x; v[u][5] = b4.y; v[u][6] = b4.z; v[u][7] = b4.w;
        } else {
          const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(src));
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            v[u][2 * j] = f.x;
            v[u][2 * j + 1] = f.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        uint4 outv;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float re = v[u][2 * j], im = v[u][2 * j + 1];
          o2[j] = __floats2bfloat162_rn(re * cs[j].x - im * cs[j].y, re * cs[j].y + im * cs[j].x);
        }
        *reinterpret_cast<uint4*>(dst_tok + (int64_t)(u == 0 ? hh : hb) * args.d) = outv;
      }
    }
  }
}

template __global__ void la_rope_cast_kernel<float>(const RopeKernelArgs);
template __global__ void la_rope_cast_kernel<__nv_bfloat16>(const RopeKernelArgs);

}  // namespace la
