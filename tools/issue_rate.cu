// tools/issue_rate.cu -- warp-instruction issue rates of the softmax's instruction classes on one SMSP (1 / 2 warps),
// independent instruction streams: FFMA, FFMA2, FADD2, FMNMX3, F2FP, MUFU.EX2, IMAD and the mixes the softmax uses.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "la_ptx.cuh"
using namespace la;
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, int warps_per_smsp, long long* res, float* sink) {
  const int warp = threadIdx.x >> 5;
  float a[16]; uint64_t b[8];
  for (int j = 0; j < 16; ++j) a[j] = threadIdx.x * 0.001f + j;
  for (int j = 0; j < 8; ++j) b[j] = pack2(a[2 * j], a[2 * j + 1]);
  const uint64_t c2 = pack2(1.0001f, 0.9999f), d2 = pack2(1e-6f, -1e-6f);
  long long t0 = 0, t1 = 0;
  if ((warp >> 2) < warps_per_smsp) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) { a[j] = fmaf(a[j], 1.0001f, a[j + 8]); a[j + 8] = fmaf(a[j + 8], 0.9999f, a[j]); }                 // 2 FFMA (3-reg-ish)
          if (MODE == 1) { b[j] = ffma2(b[j], c2, d2); b[(j + 4) & 7] = ffma2(b[(j + 4) & 7], c2, b[j]); }                   // 2 FFMA2
          if (MODE == 2) { b[j] = fadd2(b[j], d2); b[(j + 4) & 7] = fadd2(b[(j + 4) & 7], b[j]); }                            // 2 FADD2
          if (MODE == 3) { a[j] = fmax3(a[j], a[j + 8], a[(j + 1) & 15]); a[j + 8] = fmax3(a[j + 8], a[j], a[(j + 3) & 15]); } // 2 FMNMX3
          if (MODE == 4) { a[j] = __uint_as_float(pack_bf16(a[j], a[j + 8])); a[j + 8] = __uint_as_float(pack_bf16(a[j + 8], a[(j + 1) & 15])); }  // 2 F2FP
          if (MODE == 5) { a[j] = ex2_approx(a[j]); a[j + 8] = ex2_approx(a[j + 8]); }                                       // 2 MUFU
          if (MODE == 6) { a[j] = __int_as_float(__float_as_int(a[j]) * 8388608 + __float_as_int(a[j + 8])); a[j + 8] = __int_as_float(__float_as_int(a[j + 8]) * 8388608 + __float_as_int(a[j])); }  // 2 IMAD
          if (MODE == 7) { b[j] = ffma2(b[j], c2, d2); a[j] = fmax3(a[j], a[j + 8], a[(j + 1) & 15]); }                       // FFMA2 + FMNMX3
          if (MODE == 8) { b[j] = ffma2(b[j], c2, d2); a[j] = ex2_approx(a[j]); }                                              // FFMA2 + MUFU
          if (MODE == 9) { a[j] = fmax3(a[j], a[j + 8], a[(j + 1) & 15]); a[j + 8] = ex2_approx(a[j + 8]); }                   // FMNMX3 + MUFU
          if (MODE == 10) {  // the softmax mix per 4 elements: 2 FMNMX3, 2 FFMA2, 4 MUFU, 2 FADD2, 2 F2FP
            a[j] = fmax3(a[j], a[j + 8], a[(j + 1) & 15]);
            b[j] = ffma2(b[j], c2, d2);
            float x0, x1; unpack2(b[j], x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            b[(j + 4) & 7] = fadd2(b[(j + 4) & 7], pack2(p0, p1));
            a[j + 8] = __uint_as_float(pack_bf16(p0, p1));
          }
          if (MODE == 12) {  // 2 MUFU.EX2.F16 (scalar half)
            unsigned short h0 = __half_as_ushort(__float2half_rn(a[j])), h1 = __half_as_ushort(__float2half_rn(a[j + 8])), r0, r1;
            asm volatile("ex2.approx.f16 %0, %1;" : "=h"(r0) : "h"(h0));
            asm volatile("ex2.approx.f16 %0, %1;" : "=h"(r1) : "h"(h1));
            a[j] = __half2float(__ushort_as_half(r0)); a[j + 8] = __half2float(__ushort_as_half(r1));
          }
          if (MODE == 13) {  // 1 ex2.approx.f16x2 (= 2 MUFU.EX2.F16 in SASS?)
            unsigned int hh = (unsigned int)__float_as_uint(a[j]), rr;
            asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(rr) : "r"(hh));
            a[j] = __uint_as_float(rr);
          }
          if (MODE == 14) {  // 1 ex2.approx.ftz.bf16x2
            unsigned int hh = (unsigned int)__float_as_uint(a[j]), rr;
            asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(rr) : "r"(hh));
            a[j] = __uint_as_float(rr);
          }
          if (MODE == 11) { a[j] = a[j] * 1.0001f + 0.5f; a[j + 8] = a[j + 8] * 0.9999f + 0.25f; }                           // 2 FFMA imm-form
        }
      }
    }
    t1 = clock64();
  }
  float acc = 0; for (int j = 0; j < 16; ++j) acc += a[j]; for (int j = 0; j < 8; ++j) { float x, y; unpack2(b[j], x, y); acc += x + y; }
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0 && warp == 0) res[0] = t1 - t0;
}
template <int MODE> void run(const char* name, int instr_per_inner, long long* d_res, float* d_sink) {
  const int iters = 2000;
  for (int w = 1; w <= 2; ++w) {
    k<MODE><<<148, 256>>>(iters, w, d_res, d_sink);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d_res, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 4 * 8 * instr_per_inner;   // warp-instructions per warp
    printf("%-34s warps/SMSP=%d  %.2f clk per warp-instr per warp  => %.2f instr/clk/SMSP\n", name, w, h / n, n * w / h);
  }
}
int main() {
  long long* d_res; float* d_sink; cudaMalloc(&d_res, 8); cudaMalloc(&d_sink, 148 * 256 * 4);
  run<0>("FFMA (reg)", 2, d_res, d_sink);
  run<11>("FFMA (imm)", 2, d_res, d_sink);
  run<1>("FFMA2", 2, d_res, d_sink);
  run<2>("FADD2", 2, d_res, d_sink);
  run<3>("FMNMX3", 2, d_res, d_sink);
  run<4>("F2FP.BF16.PACK_AB", 2, d_res, d_sink);
  run<5>("MUFU.EX2", 2, d_res, d_sink);
  run<6>("IMAD", 2, d_res, d_sink);
  run<7>("FFMA2 + FMNMX3", 2, d_res, d_sink);
  run<8>("FFMA2 + MUFU", 2, d_res, d_sink);
  run<9>("FMNMX3 + MUFU", 2, d_res, d_sink);
  run<10>("softmax mix (6 instr / 2 elem)", 6, d_res, d_sink);
  run<12>("MUFU.EX2.F16 x2 (+4 cvt)", 2, d_res, d_sink);
  run<13>("ex2.approx.f16x2 (per PTX instr)", 1, d_res, d_sink);
  run<14>("ex2.approx.ftz.bf16x2 (per PTX instr)", 1, d_res, d_sink);
  return 0;
}
