// tools/mma_rate.cu -- tcgen05.mma issue-rate microbenchmark (no TMA, no softmax): what fraction of the
// 8192 FLOP/clk/SM tensor floor does a given instruction stream reach?  Used to separate "the MMA stream itself"
// from "the softmax/loads around it" when reading la_fwd_kernel's tensor-pipe utilisation.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I liteattention_b200/csrc -o tools/_build/mma_rate tools/mma_rate.cu
//   tools/_build/mma_rate [iters]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>

#include "la_ptx.cuh"
#include "la_tmem_ptx.cuh"

using namespace la;

constexpr int kM = 128, kN = 176, kD = 128;
constexpr uint32_t kQBlock = kM * 128, kKVBlock = kN * 128;
constexpr uint32_t kSmem = 200 * 1024;

struct Res { unsigned long long cycles; };

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// mode: see main()
// bg: what warps 0-7 do while warp 9 issues the MMA stream
//   0 nothing  1 LDTM x32 back to back (S0/S1 region)  2 LDTM paced like the kernel (88 cols / warp / ~1400 clk)
//   3 STTM x32 back to back into the (unused) columns 480-511   4 MUFU.EX2 loop   5 FFMA loop
//   6 kernel-like mix: 88 cols LDTM + 88 ex2 + 44 cols STTM per ~tile
__global__ void __launch_bounds__(320, 1) mma_rate_kernel(int mode, int iters, int fill, int bg, Res* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_mem[2];
  // operands: Q0 @0 (32 KB), Q1 @32 KB, K @64 KB (44 KB, padded to 48 KB for N=256 tests: 64 KB), V @128 KB (44 KB -> 48 KB)
  const uint32_t offQ0 = 0, offQ1 = 32768, offK = 65536, offV = 65536 + 65536;
  uint16_t* s16 = reinterpret_cast<uint16_t*>(smem);
  for (uint32_t i = threadIdx.x; i < (offV + 49152) / 2; i += blockDim.x) {
    uint16_t v = 0;
    if (fill) {
      const uint32_t h = hash32(i * 2654435761u + blockIdx.x);
      const float f = ((h & 0xFFFF) / 32768.0f - 1.0f);
      __nv_bfloat16 b = __float2bfloat16(f);
      v = *reinterpret_cast<uint16_t*>(&b);
    }
    s16[i] = v;
  }
  fence_proxy_async_smem();
  const uint32_t bar = smem_u32(&bar_mem[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;

  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  __syncthreads();
  if (threadIdx.x == 9 * 32) {
    const uint64_t q0 = make_smem_desc_sw128(sb + offQ0, 16, 1024);
    const uint64_t q1 = make_smem_desc_sw128(sb + offQ1, 16, 1024);
    const uint64_t kd = make_smem_desc_sw128(sb + offK, 16, 1024);
    const uint64_t vd_mn = make_smem_desc_sw128(sb + offV, kKVBlock, 1024);  // MN-major V [176][64] x 2
    const uint64_t vd_k = make_smem_desc_sw128(sb + offV, 16, 1024);         // K-major "V^T" [128][64] x 3
    constexpr uint32_t idQK176 = make_idesc_bf16(128, 176, 0);
    constexpr uint32_t idQK128 = make_idesc_bf16(128, 128, 0);
    constexpr uint32_t idQK256 = make_idesc_bf16(128, 256, 0);
    constexpr uint32_t idQK88 = make_idesc_bf16(128, 88, 0);
    constexpr uint32_t idQK96 = make_idesc_bf16(128, 96, 0);
    constexpr uint32_t idQK80 = make_idesc_bf16(128, 80, 0);
    constexpr uint32_t idQK192 = make_idesc_bf16(128, 192, 0);
    constexpr uint32_t idPVmn = make_idesc_bf16(128, 128, 1);
    constexpr uint32_t idPVk = make_idesc_bf16(128, 128, 0);

    auto qk = [&](uint64_t qd, uint32_t d, uint32_t idesc, uint32_t kblock, uint32_t kbyteoff) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t a_off = ((j >> 2) * kQBlock + (j & 3) * 32) >> 4;
        const uint32_t b_off = ((j >> 2) * kblock + (j & 3) * 32 + kbyteoff) >> 4;
        umma_ss(d, qd + a_off, kd + b_off, idesc, j > 0);
      }
    };
    auto pv_ts_mn = [&](uint32_t d, uint32_t p, int ksteps) {
      for (int j = 0; j < ksteps; ++j) umma_ts(d, p + j * 8, vd_mn + ((j * 16 * 128) >> 4), idPVmn, 1);
    };
    auto pv_ss_mn = [&](uint32_t d, int ksteps) {  // A = "P" from smem (K-major, Q0/Q1 buffers reused, 11 k-steps span 3 blocks -> wrap)
      for (int j = 0; j < ksteps; ++j) {
        const uint32_t a_off = (((j >> 2) & 1) * kQBlock + (j & 3) * 32) >> 4;
        umma_ss(d, q0 + a_off, vd_mn + ((j * 16 * 128) >> 4), idPVmn, 1);
      }
    };
    auto pv_ts_k = [&](uint32_t d, uint32_t p, int ksteps) {
      for (int j = 0; j < ksteps; ++j) {
        const uint32_t b_off = ((j >> 2) * (128 * 128) + (j & 3) * 32) >> 4;
        umma_ts(d, p + j * 8, vd_k + b_off, idPVk, 1);
      }
    };

    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sbuf = tm + (it & 1) * kN;
      switch (mode) {
        case 0: qk(q0, sbuf, idQK176, kKVBlock, 0); break;                           // QK only, N=176
        case 1: pv_ts_mn(tm + 352, sbuf, 11); break;                                 // PV only (TS, V MN-major)
        case 2: qk(q0, sbuf, idQK176, kKVBlock, 0); pv_ts_mn(tm + 352, tm + ((it + 1) & 1) * kN, 11); break;  // tile = QK + PV
        case 3: qk(q0, tm + (it & 1) * 128, idQK128, 128 * 128, 0); break;          // QK only, N=128
        case 4: qk(q0, tm + (it & 1) * 256, idQK256, 256 * 128, 0); break;          // QK only, N=256
        case 5: pv_ss_mn(tm + 352, 11); break;                                       // PV only, SS
        case 6: pv_ts_k(tm + 352, sbuf, 11); break;                                  // PV only, TS, K-major B
        case 7:                                                                      // tile with commits like the kernel
          qk(q0, sbuf, idQK176, kKVBlock, 0);
          tc_commit(bar + 8);
          tc_commit(bar + 8);
          pv_ts_mn(tm + 352, tm + ((it + 1) & 1) * kN, 11);
          tc_commit(bar + 8);
          tc_commit(bar + 8);
          break;
        case 8:                                                                      // two Q tiles share K/V: 2 x (QK, PV), S 88-col halves
          qk(q0, tm + 0, idQK88, kKVBlock, 0);
          qk(q1, tm + 88, idQK88, kKVBlock, 0);
          qk(q0, tm + 176, idQK88, kKVBlock, 88 * 128);
          qk(q1, tm + 264, idQK88, kKVBlock, 88 * 128);
          break;
        case 9: qk(q0, tm + (it & 1) * 96, idQK96, kKVBlock, 0); qk(q0, tm + 192 + (it & 1) * 80, idQK80, kKVBlock, 96 * 128); break;  // 96 + 80 split
        case 10: qk(q0, sbuf, idQK176, kKVBlock, 0); pv_ts_k(tm + 352, tm + ((it + 1) & 1) * kN, 11); break;  // tile, K-major V
        case 11: qk(q0, tm + (it & 1) * 192, idQK192, 192 * 128, 0); break;         // QK only, N=192
        case 12:                                                                     // tile pair for 2 Q tiles: QK_A QK_B PV_A PV_B (O_A@256.., no room: reuse)
          qk(q0, tm + 0, idQK176, kKVBlock, 0);
          qk(q1, tm + 176, idQK176, kKVBlock, 0);
          pv_ts_mn(tm + 352, tm + 0, 11);
          pv_ts_mn(tm + 352, tm + 176, 11);
          break;
        case 13: pv_ts_mn(tm + 352, sbuf, 8); break;                                 // PV, 8 k-steps (128-wide kv tile)
        case 14: qk(q0, tm + (it & 1) * 128, idQK128, 128 * 128, 0); pv_ts_mn(tm + 352, tm + ((it + 1) & 1) * 128, 8); break;  // FA4-like 128x128 tile
        default: break;
      }
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    res[blockIdx.x].cycles = (unsigned long long)(t1 - t0);
    stop_flag = 1;
  } else if (threadIdx.x < 256 && bg != 0) {
    const int w = threadIdx.x >> 5;
    const uint32_t lane_field = (uint32_t)((w & 3) * 32) << 16;
    const uint32_t base = tm + lane_field + (w >> 2) * 88;
    float acc = 0.f;
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) r[j] = threadIdx.x + j;
    long long next = clock64();
    while (!stop_flag) {
      if (bg == 1) {
        tmem_ld_x32(base, r); tmem_ld_x32(base + 32, r); tmem_wait_ld();
        acc += __uint_as_float(r[3]);
      } else if (bg == 2 || bg == 6) {
        tmem_ld_x32(base, r); tmem_wait_ld(); acc += __uint_as_float(r[1]);
        if (bg == 6) { for (int j = 0; j < 32; ++j) acc += ex2_approx(__uint_as_float(r[j]) * 1e-30f); }
        tmem_ld_x32(base + 32, r); tmem_wait_ld(); acc += __uint_as_float(r[2]);
        if (bg == 6) { for (int j = 0; j < 32; ++j) acc += ex2_approx(__uint_as_float(r[j]) * 1e-30f); }
        tmem_ld_x16(base + 64, r); tmem_ld_x8(base + 80, r + 16); tmem_wait_ld(); acc += __uint_as_float(r[3]);
        if (bg == 6) {
          for (int j = 0; j < 24; ++j) acc += ex2_approx(__uint_as_float(r[j]) * 1e-30f);
          tmem_st_x32(tm + lane_field + 480, r); tmem_wait_st();
        }
        next += 1400;
        while (clock64() < next && !stop_flag) {}
      } else if (bg == 3) {
        tmem_st_x32(tm + lane_field + 480, r); tmem_wait_st();
      } else if (bg == 4) {
        for (int j = 0; j < 32; ++j) acc += ex2_approx(acc * 1e-30f + j);
      } else if (bg == 5) {
        for (int j = 0; j < 32; ++j) acc = fmaf(acc, 1.0001f, 1e-9f * j);
      }
    }
    if (acc == 123.456f) res[0].cycles = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  const int nsm = 148;
  Res* d_res;
  cudaMalloc(&d_res, nsm * sizeof(Res));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  struct Mode { int id; const char* name; double floor_clk; };
  // floor = sum over MMAs of M*N*16 / 4096 MAC/clk
  const Mode modes[] = {
      {0, "QK SS N=176 x8", 8 * 88.0},
      {3, "QK SS N=128 x8", 8 * 64.0},
      {11, "QK SS N=192 x8", 8 * 96.0},
      {4, "QK SS N=256 x8", 8 * 128.0},
      {8, "QK SS N=88 x8 x4 (2 Q tiles x 2 halves)", 4 * 8 * 44.0},
      {9, "QK SS N=96 x8 + N=80 x8", 8 * 48.0 + 8 * 40.0},
      {1, "PV TS N=128 x11 (V MN-major)", 11 * 64.0},
      {13, "PV TS N=128 x8 (V MN-major)", 8 * 64.0},
      {6, "PV TS N=128 x11 (V K-major)", 11 * 64.0},
      {5, "PV SS N=128 x11 (V MN-major)", 11 * 64.0},
      {2, "tile: QK176 x8 + PV x11", 8 * 88.0 + 11 * 64.0},
      {7, "tile + 4 commits", 8 * 88.0 + 11 * 64.0},
      {10, "tile, V K-major", 8 * 88.0 + 11 * 64.0},
      {12, "2 tiles: QK QK PV PV", 2 * (8 * 88.0 + 11 * 64.0)},
      {14, "FA4-like tile: QK128 x8 + PV x8", 16 * 64.0},
  };
  const int only_mode = argc > 2 ? atoi(argv[2]) : -1;
  const int bg_max = argc > 3 ? atoi(argv[3]) : 0;
  for (int bg = 0; bg <= bg_max; ++bg)
  for (int fill = (only_mode >= 0 ? 1 : 0); fill < 2; ++fill) {
    printf("---- operands %s, background work %d\n", fill ? "random bf16 in [-1,1)" : "zero", bg);
    for (const Mode& m : modes) {
      if (only_mode >= 0 && m.id != only_mode) continue;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      mma_rate_kernel<<<nsm, 320, kSmem>>>(m.id, 50, fill, bg, d_res);  // warm
      cudaEventRecord(e0);
      mma_rate_kernel<<<nsm, 320, kSmem>>>(m.id, iters, fill, bg, d_res);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("mode %d: %s\n", m.id, cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<Res> h(nsm);
      cudaMemcpy(h.data(), d_res, nsm * sizeof(Res), cudaMemcpyDeviceToHost);
      double sum = 0, mn = 1e30, mx = 0;
      for (auto& r : h) { double c = (double)r.cycles / iters; sum += c; mn = c < mn ? c : mn; mx = c > mx ? c : mx; }
      const double avg = sum / nsm;
      const double flops = m.floor_clk * 8192.0 * iters * nsm;
      printf("%-44s floor %7.1f clk  measured avg %7.1f (min %7.1f max %7.1f)  util %5.1f%%  %7.1f TFLOP/s  (%.3f ms, ~%.0f MHz)\n",
             m.name, m.floor_clk, avg, mn, mx, 100.0 * m.floor_clk / avg, flops / (ms * 1e-3) / 1e12, ms,
             avg * iters / (ms * 1e-3) / 1e6);
    }
  }
  return 0;
}
