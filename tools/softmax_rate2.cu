// tools/softmax_rate2.cu -- round 2: how should the 128x176 softmax tile be spread over warps?
// Compares, per SM (x148 CTAs), the clocks per tile of the kernel's speculative softmax body
//   (tcgen05.ld -> row max / ffma2 / exp2 (MUFU or FMA-pipe polynomial) / row sum / bf16 pack -> tcgen05.st,
//    half-row-max exchange through smem + named barrier, warp vote)
// for 8 warps x 88 columns (the round-1 organisation) and 16 warps x 44 columns (four warps per SM sub-partition),
// with and without the row-sum adds (a ones-column in the PV MMA could produce the row sum instead), several
// polynomial shares, and with a concurrent tcgen05.mma stream (QK^T N=176 SS + PV N=128 TS, the kernel's own).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I liteattention_b200/csrc \
//        -o tools/_build/softmax_rate2 tools/softmax_rate2.cu
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

template <int COLS>
__device__ __forceinline__ void ld_cols(uint32_t a, uint32_t* r) {
  if constexpr (COLS == 88) { tmem_ld_x32(a, r); tmem_ld_x32(a + 32, r + 32); tmem_ld_x16(a + 64, r + 64); tmem_ld_x8(a + 80, r + 80); }
  else { tmem_ld_x32(a, r); tmem_ld_x8(a + 32, r + 32); tmem_ld_x4(a + 40, r + 40); }
}
template <int COLS>
__device__ __forceinline__ void st_cols(uint32_t a, const uint32_t* r) {   // COLS/2 packed columns
  if constexpr (COLS == 88) { tmem_st_x32(a, r); tmem_st_x8(a + 32, r + 32); tmem_st_x4(a + 40, r + 40); }
  else { tmem_st_x16(a, r); tmem_st_x4(a + 16, r + 16); tmem_st_x2(a + 20, r + 20); }
}

// One tile's worth of work for one thread: COLS columns of one row.  SPLIT: the row max of ALL columns is taken during
// the first half of the loop and handed to `post` at the midpoint, so that an exchange can run under the second half.
template <int COLS, uint32_t MASK, int ROWSUM, int ORDERED, int SPLIT, typename Post>
__device__ __forceinline__ void body(const float* s, uint32_t* pr, uint64_t c2, uint64_t nm2, float& m_half, float& sum, Post post) {
  uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
  float mx0 = -INFINITY, mx1 = -INFINITY;
  constexpr int Q = COLS / 4, QH = (Q + 1) / 2;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = 4 * q;
    if (SPLIT) {
      if (q < QH) {
        mx0 = fmax3(mx0, s[j], s[j + 1]); mx1 = fmax3(mx1, s[j + 2], s[j + 3]);
        if (q + QH < Q) { mx0 = fmax3(mx0, s[j + 4 * QH], s[j + 4 * QH + 1]); mx1 = fmax3(mx1, s[j + 4 * QH + 2], s[j + 4 * QH + 3]); }
      }
      if (q == QH) { m_half = fmaxf(mx0, mx1); post(m_half); }
    } else {
      mx0 = fmax3(mx0, s[j], s[j + 1]); mx1 = fmax3(mx1, s[j + 2], s[j + 3]);
    }
    float t0, t1, t2, t3, p0, p1, p2, p3;
    unpack2(ffma2(pack2(s[j], s[j + 1]), c2, nm2), t0, t1);
    unpack2(ffma2(pack2(s[j + 2], s[j + 3]), c2, nm2), t2, t3);
    if ((MASK >> ((j >> 1) & 7)) & 1u) exp2_poly_pair(t0, t1, p0, p1);
    else if (ORDERED) { p0 = ex2_approx_ordered(t0); p1 = ex2_approx_ordered(t1); }
    else { p0 = ex2_approx(t0); p1 = ex2_approx(t1); }
    if ((MASK >> (((j >> 1) + 1) & 7)) & 1u) exp2_poly_pair(t2, t3, p2, p3);
    else if (ORDERED) { p2 = ex2_approx_ordered(t2); p3 = ex2_approx_ordered(t3); }
    else { p2 = ex2_approx(t2); p3 = ex2_approx(t3); }
    if (ROWSUM) { acc0 = fadd2(acc0, pack2(p0, p1)); acc1 = fadd2(acc1, pack2(p2, p3)); }
    pr[j / 2] = pack_bf16(p0, p1); pr[j / 2 + 1] = pack_bf16(p2, p3);
  }
  float a0, a1, a2, a3; unpack2(acc0, a0, a1); unpack2(acc1, a2, a3);
  sum += (a0 + a1) + (a2 + a3);
  if (!SPLIT) m_half = fmaxf(mx0, mx1);
}

__host__ __device__ constexpr uint32_t rotl8(uint32_t m, int r) { return ((m << (r & 7)) | (m >> ((8 - r) & 7))) & 0xFFu; }

// NW softmax warps (8: 88 columns each, 16: 44 columns each) + 1 MMA warp.
template <int NW, uint32_t MASK, int ROWSUM, int ORDERED, int XCHG>
__global__ void __launch_bounds__(NW * 32 + 128, 1) k(int iters, unsigned long long* res, float* sink, int with_mma) {
  constexpr int COLS = 176 / (NW / 4);
  constexpr int kRegsSoftmax = NW == 8 ? 216 : 112, kRegsOther = NW == 8 ? 72 : 32;   // setmaxnreg only redistributes the launch allocation (640 x 96 = 16 x 32 x 112 + 4 x 32 x 32)
  constexpr int G = NW / 4;   // warps per row group (= per SM sub-partition)
  extern __shared__ uint8_t smem_raw[];
  __shared__ int done_warps;
  __shared__ uint32_t tmem_slot;
  __shared__ float xchg[2][G][128];
  __shared__ __align__(8) uint64_t xbar[4][2];          // XCHG == 2: one mbarrier per row group and tile parity, G arrivals
  if (threadIdx.x == 0) {
    done_warps = 0;
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&xbar[i >> 1][i & 1]), G);
    fence_mbar_init();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  float acc = 0.f;
  if (warp >= NW) {
    setmaxnreg_dec<kRegsOther>();
    if (warp == NW && with_mma && lane == 0) {
      uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
      const uint32_t sb = smem_u32(smem);
      const uint64_t qd = make_smem_desc_sw128(sb, 16, 1024), kd = make_smem_desc_sw128(sb + 32768, 16, 1024);
      const uint64_t vd = make_smem_desc_sw128(sb + 32768 + 45056, 176 * 128, 1024);
      constexpr uint32_t idQK = make_idesc_bf16(128, 176, 0), idPV = make_idesc_bf16(128, 128, 1);
      long long n = 0;
      // D goes to columns 352.. (O) and to a scratch S at 176 so that the softmax warps' region [0,176) keeps its values
      while (*reinterpret_cast<volatile int*>(&done_warps) < NW) {
        for (int j = 0; j < 8; ++j) umma_ss(tm + 176, qd + ((((j >> 2) * 16384 + (j & 3) * 32)) >> 4), kd + ((((j >> 2) * 22528 + (j & 3) * 32)) >> 4), idQK, j > 0);
        for (int j = 0; j < 11; ++j) umma_ts(tm + 352, tm + 176 + j * 8, vd + ((j * 16 * 128) >> 4), idPV, 1);
        ++n;
      }
      res[148 * 32 + blockIdx.x] = (unsigned long long)n;
    }
  } else {
    setmaxnreg_inc<kRegsSoftmax>();
    const int r = warp & 3, g = warp >> 2;
    const int row = r * 32 + lane;
    const uint32_t lane_field = (uint32_t)(r * 32) << 16;
    const uint32_t s_addr = tm + lane_field + g * COLS;
    const uint32_t p_addr = tm + lane_field + 480 + 0;   // 32 spare columns: every warp group stores to the same place (timing only)
    { uint32_t z[32]; for (int j = 0; j < 32; ++j) z[j] = __float_as_uint(-1.0f - 0.01f * j);
      for (int c0 = 0; c0 + 32 <= COLS; c0 += 32) tmem_st_x32(s_addr + c0, z);
      tmem_st_x32(s_addr + COLS - 32, z); tmem_wait_st(); }
    __syncwarp();
    const float c = 0.1275f;
    float m_ref = 0.3f, l_run = 0.f, m_true = -1e30f;
    const uint64_t c2 = pack2(c, c);
    float s[COLS]; uint32_t* sr = reinterpret_cast<uint32_t*>(s);
    ld_cols<COLS>(s_addr, sr);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      tmem_wait_ld();
      const float neg = -m_ref * c;
      const uint64_t nm2 = pack2(neg, neg);
      uint32_t pr[COLS / 2];
      float m_half, sum = 0.f;
      // four staggered masks, one per warp of the sub-partition
      const uint32_t xb = smem_u32(&xbar[r][it & 1]);
      auto post = [&](float mh) {
        if (XCHG == 2) {
          xchg[it & 1][g][row] = mh;
          __syncwarp();
          if (lane == 0) mbar_arrive(xb);
        }
      };
      if (g == 0) body<COLS, rotl8(MASK, 0), ROWSUM, ORDERED, XCHG == 2>(s, pr, c2, nm2, m_half, sum, post);
      else if (g == 1) body<COLS, rotl8(MASK, G == 2 ? 2 : 1), ROWSUM, ORDERED, XCHG == 2>(s, pr, c2, nm2, m_half, sum, post);
      else if (g == 2) body<COLS, rotl8(MASK, 2), ROWSUM, ORDERED, XCHG == 2>(s, pr, c2, nm2, m_half, sum, post);
      else body<COLS, rotl8(MASK, 3), ROWSUM, ORDERED, XCHG == 2>(s, pr, c2, nm2, m_half, sum, post);
      float m_loc = m_half;
      if (XCHG == 1) {
        xchg[it & 1][g][row] = m_half;
        named_bar_sync(1 + r, G * 32);
#pragma unroll
        for (int o = 1; o < G; ++o) m_loc = fmaxf(m_loc, xchg[it & 1][(g + o) % G][row]);
      }
      if (XCHG == 2) {
        mbar_wait(xb, (it >> 1) & 1);
#pragma unroll
        for (int o = 1; o < G; ++o) m_loc = fmaxf(m_loc, xchg[it & 1][(g + o) % G][row]);
      }
      const bool exact = __any_sync(0xffffffffu, !((m_loc - m_ref) * c <= 8.0f));
      if (exact) m_ref = m_loc;   // never taken with these values
      l_run += sum;
      m_true = fmaxf(m_true, m_loc);
      // publish P, then prefetch S for the next iteration (as the kernel does)
      if constexpr (COLS == 88) { tmem_st_x32(p_addr, pr); tmem_st_x8(p_addr, pr + 32); tmem_st_x4(p_addr, pr + 40); }
      else { tmem_st_x16(p_addr, pr); tmem_st_x4(p_addr + 16, pr + 16); tmem_st_x2(p_addr + 20, pr + 20); }
      ld_cols<COLS>(s_addr, sr);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
    }
    tmem_wait_ld();
    const long long t1 = clock64();
    acc = l_run + m_true + s[0];
    if (lane == 0) { res[blockIdx.x * 32 + warp] = (unsigned long long)(t1 - t0); atomicAdd(&done_warps, 1); }
    __syncwarp();
    sink[blockIdx.x * 512 + threadIdx.x] = acc;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static unsigned long long* d_res; static float* d_sink;
template <int NW, uint32_t MASK, int ROWSUM, int ORDERED, int XCHG>
void run(const char* name, int iters) {
  const int SM = 200 * 1024;
  auto fn = k<NW, MASK, ROWSUM, ORDERED, XCHG>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, fn);
  double out[2] = {0, 0}; double mma_frac = 0;
  for (int with_mma = 0; with_mma < 2; ++with_mma) {
    cudaMemset(d_res, 0, (148 * 32 + 148) * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fn<<<148, NW * 32 + 128, SM>>>(iters, d_res, d_sink, with_mma);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: err %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<unsigned long long> h(148 * 32 + 148);
    cudaMemcpy(h.data(), d_res, h.size() * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < 148 * 32; ++i) mx = h[i] > mx ? (double)h[i] : mx;
    out[with_mma] = mx / iters;
    if (with_mma) { double tiles = 0; for (int i = 0; i < 148; ++i) tiles += (double)h[148 * 32 + i]; mma_frac = tiles / 148 * 1408.0 / mx; }
  }
  printf("%-64s regs %3d  %7.1f clk/tile alone   %7.1f clk/tile with MMA stream (tensor pipe %.0f%% busy)\n", name, fa.numRegs, out[0], out[1], mma_frac * 100);
  fflush(stdout);
}


// Ping-pong organisation probe: ONE warp per SM sub-partition (4 warps), a thread owns a whole 176-column row and walks
// it as two chunks of 88 columns (no exchange of partial row maxima at all).  How long is one tile with the
// sub-partition to itself?  (TEAMS = 2: a second team of 4 warps does the same on another buffer, half a tile out of phase.)
template <uint32_t MASK, int ROWSUM, int TEAMS>
__global__ void __launch_bounds__(TEAMS * 128 + 128, 1) kpp(int iters, unsigned long long* res, float* sink, int with_mma) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int done_warps;
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x == 0) done_warps = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  constexpr int NWS = TEAMS * 4;
  if (warp >= NWS) {
    setmaxnreg_dec<72>();
    if (warp == NWS && with_mma && lane == 0) {
      uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
      const uint32_t sb = smem_u32(smem);
      const uint64_t qd = make_smem_desc_sw128(sb, 16, 1024), kd = make_smem_desc_sw128(sb + 32768, 16, 1024);
      const uint64_t vd = make_smem_desc_sw128(sb + 32768 + 45056, 176 * 128, 1024);
      constexpr uint32_t idQK = make_idesc_bf16(128, 160, 0), idPV = make_idesc_bf16(128, 128, 1);   // N = 160: D stays inside columns 352..511
      long long n = 0;
      while (*reinterpret_cast<volatile int*>(&done_warps) < NWS) {
        for (int j = 0; j < 8; ++j) umma_ss(tm + 352, qd + ((((j >> 2) * 16384 + (j & 3) * 32)) >> 4), kd + ((((j >> 2) * 22528 + (j & 3) * 32)) >> 4), idQK, j > 0);
        for (int j = 0; j < 11; ++j) umma_ts(tm + 352, tm + 352 + j * 8, vd + ((j * 16 * 128) >> 4), idPV, 1);
        ++n;
      }
      res[148 * 32 + blockIdx.x] = (unsigned long long)n;
    }
  } else {
    setmaxnreg_inc<TEAMS == 1 ? 216 : 216>();
    const int r = warp & 3, team = warp >> 2;
    const uint32_t lane_field = (uint32_t)(r * 32) << 16;
    const uint32_t s_addr = tm + lane_field + team * 176;
    const uint32_t p_addr = s_addr;      // P over the first 88 columns of the team's own S buffer (re-initialised below)
    uint32_t z[32]; for (int j = 0; j < 32; ++j) z[j] = __float_as_uint(-1.0f - 0.01f * j);
    auto reinit = [&]() { for (int c0 = 0; c0 < 96; c0 += 32) tmem_st_x32(s_addr + c0, z); };
    for (int c0 = 0; c0 + 32 <= 176; c0 += 32) tmem_st_x32(s_addr + c0, z);
    tmem_st_x16(s_addr + 160, z); tmem_wait_st();
    __syncwarp();
    const float c = 0.1275f;
    float m_ref = 0.3f, l_run = 0.f, m_true = -1e30f;
    const uint64_t c2 = pack2(c, c);
    float s[88]; uint32_t* sr = reinterpret_cast<uint32_t*>(s);
    if (team == 1) { for (int w = 0; w < 40; ++w) l_run += ex2_approx(l_run * 1e-20f - w); }   // some phase offset
    ld_cols<88>(s_addr, sr);
    const long long t0 = clock64();
    auto nopost = [](float) {};
    for (int it = 0; it < iters; ++it) {
      const float neg = -m_ref * c;
      const uint64_t nm2 = pack2(neg, neg);
      uint32_t pr[88];
      float mh0, mh1, sum = 0.f;
      tmem_wait_ld();
      body<88, MASK, ROWSUM, 1, 0>(s, pr, c2, nm2, mh0, sum, nopost);
      ld_cols<88>(s_addr + 88, sr);
      tmem_wait_ld();
      body<88, rotl8(MASK, 2), ROWSUM, 1, 0>(s, pr + 44, c2, nm2, mh1, sum, nopost);
      const float m_loc = fmaxf(mh0, mh1);
      const bool exact = __any_sync(0xffffffffu, !((m_loc - m_ref) * c <= 8.0f));
      if (exact) m_ref = m_loc;
      l_run += sum;
      m_true = fmaxf(m_true, m_loc);
      tmem_st_x32(p_addr, pr); tmem_st_x32(p_addr + 32, pr + 32); tmem_st_x16(p_addr + 64, pr + 64); tmem_st_x8(p_addr + 80, pr + 80);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      reinit();                      // stands in for the next QK^T writing S (keeps the values finite)
      tmem_wait_st();
      ld_cols<88>(s_addr, sr);
    }
    tmem_wait_ld();
    const long long t1 = clock64();
    if (lane == 0) { res[blockIdx.x * 32 + warp] = (unsigned long long)(t1 - t0); atomicAdd(&done_warps, 1); }
    __syncwarp();
    sink[blockIdx.x * 512 + threadIdx.x] = l_run + m_true + s[0];
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <uint32_t MASK, int ROWSUM, int TEAMS>
void runpp(const char* name, int iters) {
  const int SM = 200 * 1024;
  auto fn = kpp<MASK, ROWSUM, TEAMS>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  double out[2] = {0, 0};
  for (int with_mma = 0; with_mma < 2; ++with_mma) {
    cudaMemset(d_res, 0, (148 * 32 + 148) * 8);
    fn<<<148, TEAMS * 128 + 128, SM>>>(iters, d_res, d_sink, with_mma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: err %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<unsigned long long> h(148 * 32 + 148);
    cudaMemcpy(h.data(), d_res, h.size() * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < 148 * 32; ++i) mx = h[i] > mx ? (double)h[i] : mx;
    out[with_mma] = mx / iters;
  }
  printf("%-64s           %7.1f clk per tile per team alone   %7.1f with MMA stream\n", name, out[0], out[1]);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 3000;
  cudaMalloc(&d_res, (148 * 32 + 148) * 8); cudaMalloc(&d_sink, 148 * 512 * 4);
  runpp<0x11u, 1, 1>("PP: 4 warps (1 per SMSP) x 176 cols, poly 2/8, rowsum (one team alone; includes a 96-col TMEM re-init)", iters);
  runpp<0x11u, 1, 2>("PP: 2 teams x 4 warps x 176 cols, poly 2/8, rowsum (per team; includes a 96-col TMEM re-init)", iters);
  runpp<0x49u, 1, 1>("PP: 4 warps x 176 cols, poly 3/8, rowsum", iters);
  runpp<0x49u, 1, 2>("PP: 2 teams, poly 3/8, rowsum", iters);
  runpp<0x55u, 1, 1>("PP: 4 warps x 176 cols, poly 4/8, rowsum", iters);
  runpp<0x00u, 1, 1>("PP: 4 warps x 176 cols, poly 0/8, rowsum", iters);
  //                 NW  MASK  ROWSUM ORDERED XCHG(0 none, 1 bar.sync at the end, 2 post at the midpoint + mbarrier wait at the end)
  run<16, 0x11u, 1, 1, 1>("16 warps x 44, poly 2/8, rowsum, bar.sync xchg", iters);
  run<16, 0x11u, 1, 1, 2>("16 warps x 44, poly 2/8, rowsum, split mbarrier xchg", iters);
  run<16, 0x11u, 1, 1, 0>("16 warps x 44, poly 2/8, rowsum, no xchg", iters);
  run<16, 0x11u, 0, 1, 2>("16 warps x 44, poly 2/8, no rowsum, split mbarrier xchg", iters);
  run<16, 0x49u, 1, 1, 2>("16 warps x 44, poly 3/8, rowsum, split mbarrier xchg", iters);
  run<16, 0x49u, 0, 1, 2>("16 warps x 44, poly 3/8, no rowsum, split mbarrier xchg", iters);
  run<16, 0x55u, 1, 1, 2>("16 warps x 44, poly 4/8, rowsum, split mbarrier xchg", iters);
  run<16, 0x00u, 1, 1, 2>("16 warps x 44, poly 0/8, rowsum, split mbarrier xchg", iters);
  run<16, 0x11u, 1, 0, 2>("16 warps x 44, poly 2/8, rowsum, unordered, split mbarrier xchg", iters);
  run<8, 0x11u, 1, 1, 1>(" 8 warps x 88, poly 2/8, rowsum, bar.sync xchg   (round-1 kernel)", iters);
  run<8, 0x11u, 1, 1, 2>(" 8 warps x 88, poly 2/8, rowsum, split mbarrier xchg", iters);
  run<8, 0x49u, 1, 1, 2>(" 8 warps x 88, poly 3/8, rowsum, split mbarrier xchg", iters);
  run<8, 0x11u, 0, 1, 2>(" 8 warps x 88, poly 2/8, no rowsum, split mbarrier xchg", iters);
  run<8, 0x49u, 0, 1, 2>(" 8 warps x 88, poly 3/8, no rowsum, split mbarrier xchg", iters);
  return 0;
}
