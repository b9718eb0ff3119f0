// tools/softmax_rate.cu -- throughput of the softmax building blocks on one SM (x148): tcgen05.ld/st, MUFU.EX2,
// and the kernel's per-tile loop body, for 4 / 8 warps.  Prints clocks per "tile" (128 rows x 176 columns).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I liteattention_b200/csrc -o tools/_build/softmax_rate tools/softmax_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "la_ptx.cuh"
#include "la_tmem_ptx.cuh"
using namespace la;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void exp2_poly_pair(float t0, float t1, float& p0, float& p1) {
  const float kMagic = 12582912.f;
  t0 = fmaxf(t0, -126.f); t1 = fmaxf(t1, -126.f);
  const uint64_t t = pack2(t0, t1);
  const uint64_t xf = fadd2(t, pack2(kMagic, kMagic));
  const uint64_t n = fadd2(xf, pack2(-kMagic, -kMagic));
  const uint64_t r = ffma2(n, pack2(-1.f, -1.f), t);
  uint64_t p = ffma2(pack2(0.05517167f, 0.05517167f), r, pack2(0.24261113f, 0.24261113f));
  p = ffma2(p, r, pack2(0.69326097f, 0.69326097f));
  p = ffma2(p, r, pack2(0.99992806f, 0.99992806f));
  float x0, x1, q0, q1; unpack2(xf, x0, x1); unpack2(p, q0, q1);
  p0 = __int_as_float(__float_as_int(x0) * (1 << 23) + __float_as_int(q0));
  p1 = __int_as_float(__float_as_int(x1) * (1 << 23) + __float_as_int(q1));
}
// mode 0: LDTM only (88 cols/warp as x32,x32,x16,x8 + wait)   1: LDTM x32 back-to-back, one wait per 4
// 2: STTM only (44 cols/warp)     3: 88 MUFU.EX2 per warp (independent)    4: LDTM + max only
// 5: full fast-path body (ld, max, ffma2, ex2, fadd2, pack, st)            6: body without LDTM/STTM (regs only)
// 7: LDTM 16x256b... (not implemented)
template <uint32_t MASK>
__global__ void __launch_bounds__(288, 1) k(int mode, int iters, int nwarps, unsigned long long* res, float* sink, int with_mma) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t s_addr = tm + lane_field + (warp >> 2) * 88;
  const uint32_t p_addr = tm + lane_field + 352 + (warp >> 2) * 44;
  float acc = 0.f;
  if (warp == 8) {
    if (with_mma && (threadIdx.x & 31) == 0) {
      uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
      const uint32_t sb = smem_u32(smem);
      const uint64_t qd = make_smem_desc_sw128(sb, 16, 1024), kd = make_smem_desc_sw128(sb + 32768, 16, 1024);
      const uint64_t vd = make_smem_desc_sw128(sb + 32768 + 45056, 176 * 128, 1024);
      constexpr uint32_t idQK = make_idesc_bf16(128, 176, 0), idPV = make_idesc_bf16(128, 128, 1);
      long long n = 0;
      while (!stop_flag) {
        for (int j = 0; j < 8; ++j) umma_ss(tm + (n & 1) * 176, qd + ((((j >> 2) * 16384 + (j & 3) * 32)) >> 4), kd + ((((j >> 2) * 22528 + (j & 3) * 32)) >> 4), idQK, j > 0);
        for (int j = 0; j < 11; ++j) umma_ts(tm + 352, tm + ((n + 1) & 1) * 176 + j * 8, vd + ((j * 16 * 128) >> 4), idPV, 1);
        ++n;
      }
      res[blockIdx.x * 8 + 7 + 148 * 8] = (unsigned long long)n;
    }
  } else if (warp < nwarps) {
    // init TMEM region with small values
    { uint32_t z[32]; for (int j = 0; j < 32; ++j) z[j] = __float_as_uint(-1.0f - 0.01f * j);
      tmem_st_x32(s_addr, z); tmem_st_x32(s_addr + 32, z); tmem_st_x32(s_addr + 56, z); tmem_wait_st(); }
    __syncwarp();
    const long long t0 = clock64();
    const float c = 0.1275f, nm = 0.3f;
    const uint64_t c2 = pack2(c, c), nm2 = pack2(nm, nm);
    for (int it = 0; it < iters; ++it) {
      float s[88]; uint32_t* sr = reinterpret_cast<uint32_t*>(s);
      if (mode == 0 || mode == 4 || mode == 5) {
        tmem_ld_x32(s_addr, sr); tmem_ld_x32(s_addr + 32, sr + 32); tmem_ld_x16(s_addr + 64, sr + 64); tmem_ld_x8(s_addr + 80, sr + 80);
        tmem_wait_ld();
        if (mode == 0) { acc += s[0] + s[40] + s[70] + s[85]; }
      }
      if (mode == 1) {
        tmem_ld_x32(s_addr, sr); tmem_ld_x32(s_addr + 32, sr + 32); tmem_ld_x32(s_addr + 56, sr + 56);
        tmem_wait_ld(); acc += s[0] + s[33] + s[60];
      }
      if (mode == 2) {
        uint32_t pr[44]; for (int j = 0; j < 44; ++j) pr[j] = it + j;
        tmem_st_x32(p_addr, pr); tmem_st_x8(p_addr + 32, pr + 32); tmem_st_x4(p_addr + 40, pr + 40); tmem_wait_st();
      }
      if (mode == 3) {
#pragma unroll
        for (int j = 0; j < 88; ++j) acc += ex2_approx(-(float)(it + j) * 1e-3f - acc * 1e-30f);
      }
      if (mode == 6) {
#pragma unroll
        for (int j = 0; j < 88; ++j) s[j] = -(float)(it + j) * 1e-3f - acc * 1e-30f;
      }
      if (mode == 4) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 88; j += 4) { mx0 = fmax3(mx0, s[j], s[j + 1]); mx1 = fmax3(mx1, s[j + 2], s[j + 3]); }
        acc += fmaxf(mx0, mx1);
      }
      if (mode == 5 || mode == 6) {
        uint32_t pr[44];
        uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 88; j += 4) {
          mx0 = fmax3(mx0, s[j], s[j + 1]); mx1 = fmax3(mx1, s[j + 2], s[j + 3]);
          float t0_, t1_, t2_, t3_;
          unpack2(ffma2(pack2(s[j], s[j + 1]), c2, nm2), t0_, t1_);
          unpack2(ffma2(pack2(s[j + 2], s[j + 3]), c2, nm2), t2_, t3_);
          float p0, p1, p2, p3;
          if ((MASK >> ((j >> 1) & 7)) & 1u) exp2_poly_pair(t0_, t1_, p0, p1); else { p0 = ex2_approx(t0_); p1 = ex2_approx(t1_); }
          if ((MASK >> (((j >> 1) + 1) & 7)) & 1u) exp2_poly_pair(t2_, t3_, p2, p3); else { p2 = ex2_approx(t2_); p3 = ex2_approx(t3_); }
          acc0 = fadd2(acc0, pack2(p0, p1)); acc1 = fadd2(acc1, pack2(p2, p3));
          pr[j / 2] = pack_bf16(p0, p1); pr[j / 2 + 1] = pack_bf16(p2, p3);
        }
        float a0, a1, a2, a3; unpack2(acc0, a0, a1); unpack2(acc1, a2, a3);
        acc += a0 + a1 + a2 + a3 + fmaxf(mx0, mx1);
        if (mode == 5) { tmem_st_x32(p_addr, pr); tmem_st_x8(p_addr + 32, pr + 32); tmem_st_x4(p_addr + 40, pr + 40); tmem_wait_st(); }
        else { uint32_t x = 0; for (int j = 0; j < 44; ++j) x ^= pr[j]; acc += __uint_as_float(x & 0xff); }
      }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) res[blockIdx.x * 8 + warp] = (unsigned long long)(t1 - t0);
    __syncwarp();
    if (threadIdx.x == 0) stop_flag = 1;
    sink[blockIdx.x * 256 + threadIdx.x] = acc;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  unsigned long long* d_res; float* d_sink;
  cudaMalloc(&d_res, 148 * 8 * 8 * 2 + 64); cudaMalloc(&d_sink, 148 * 256 * 4);
  const char* names[] = {"LDTM 88 cols/warp (x32,x32,x16,x8) + wait", "LDTM 3 x x32 + wait (96 cols)", "STTM 44 cols/warp + wait",
                         "88 MUFU.EX2 per warp", "LDTM + row max", "full fast-path body (ld..st)", "body, registers only",
                         "body, poly 1/8", "body, poly 2/8", "body, poly 3/8", "body, poly 4/8", "body, poly 8/8"};
  const int SM = 200 * 1024;
  cudaFuncSetAttribute(k<0x00u>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncSetAttribute(k<0x10u>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncSetAttribute(k<0x44u>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncSetAttribute(k<0x92u>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncSetAttribute(k<0x55u>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncSetAttribute(k<0xFFu>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  for (int with_mma = 0; with_mma < 2; ++with_mma)
  for (int nw : {8}) {
    printf("---- concurrent tcgen05.mma stream: %s\n", with_mma ? "YES" : "no");
    for (int mode = 0; mode < 12; ++mode) {
      cudaMemset(d_res, 0, 148 * 8 * 8);
      if (mode < 7) k<0x00u><<<148, 288, SM>>>(mode, iters, nw, d_res, d_sink, with_mma);
      else if (mode == 7) k<0x10u><<<148, 288, SM>>>(5, iters, nw, d_res, d_sink, with_mma);
      else if (mode == 8) k<0x44u><<<148, 288, SM>>>(5, iters, nw, d_res, d_sink, with_mma);
      else if (mode == 9) k<0x92u><<<148, 288, SM>>>(5, iters, nw, d_res, d_sink, with_mma);
      else if (mode == 10) k<0x55u><<<148, 288, SM>>>(5, iters, nw, d_res, d_sink, with_mma);
      else k<0xFFu><<<148, 288, SM>>>(5, iters, nw, d_res, d_sink, with_mma);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<unsigned long long> h(148 * 8);
      cudaMemcpy(h.data(), d_res, 148 * 8 * 8, cudaMemcpyDeviceToHost);
      double mx = 0; for (auto v : h) mx = v > mx ? (double)v : mx;
      printf("warps=%d  %-44s %8.1f clk per iteration (slowest warp)\n", nw, names[mode], mx / iters);
    }
  }
  return 0;
}
