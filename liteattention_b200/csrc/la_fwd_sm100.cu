// la_fwd_sm100.cu -- skip-list-gated attention forward for B200 (sm_100a).
//
// What it computes (restating the reference's behaviour, not its code):
//   for one (batch b, head h, 128-row Q tile m) walk the K tiles named by the read list
//   (descending inclusive ranges, SkipListReader, hopper/_internal/cpp/mainloop_fwd_sm90_tma_gmma_ws.hpp:47-115)
//   and run S = Q K^T -> online softmax -> O += P V on exactly those tiles
//   (mainloop :1612-1663 first tile, :1667-1755 steady state, softmax.h:139-222, :263-296).
//   Per visited tile it also emits the left-hand side of the QK-skip predicate (softmax.h:194)
//   reduced over the tile's 128 rows, which la_skip_update.cu turns into the next list.
//
// How (B200-first, nothing shared with the Hopper kernel):
//   * 10 warps: 2 softmax warpgroups, 1 TMA producer warp, 1 tcgen05 issuer warp.
//   * Q (128x128), K and V tiles (176x128, 2 stages each) are TMA-loaded into 128B-swizzled smem.
//   * S = Q K^T : tcgen05.mma kind::f16, M=128 N=176 K=16 x8, SS operands, fp32 accumulator in TMEM.
//   * P (bf16) is written back over S in TMEM and fed to O += P V as the TMEM A operand
//     (M=128 N=128 K=16 x11, V is the MN-major smem B operand); O lives in TMEM for the whole CTA.
//   * TMEM map (512 columns allocated): S0 @0, S1 @176, O @352..479.
//   * Tile ping-pong: warpgroup A owns the even visited tiles (S buffer 0), warpgroup B the odd ones
//     (S buffer 1); one thread = one query row (TMEM lane) over all 176 columns, in two passes over TMEM
//     (pass 1 row max, pass 2 exp2 / row sum / bf16 P).  While A's tile is in the tensor pipe (PV, next QK^T)
//     B's tile is in the MUFU/FMA pipes and vice versa; the running row max travels A -> B -> A through
//     shared memory + 64-thread named barriers.  (ncu on the first version -- both warpgroups on the same
//     tile -- showed MUFU idle during every max/exchange phase and the tensor pipe at 53 %.)
//   * exp2: MUFU.EX2 is exactly as expensive as the two MMAs at d=128 (16 ex2/clk/SM vs 8192 MAC/clk/SM), so
//     kPolyPairs of every 8 column pairs are evaluated on the FMA pipe instead (Cody-Waite + degree-3
//     minimax polynomial, packed f32x2 math, relative error 7.5e-5 -- below bf16 rounding of P).
//   * O is rescaled in TMEM only when some row max of the warp actually moved (exact, not lazy).
#include <cuda_bf16.h>

#include "la_kernels.h"
#include "la_ptx.cuh"
#include "la_tmem_ptx.cuh"

namespace la {

#ifdef LA_PROFILE_CLOCKS
__device__ unsigned long long g_la_prof[16];
#define LA_CLK(var) const long long var = clock64()
#define LA_ACC(idx, a, b) prof_acc[idx] += (b) - (a)
#else
#define LA_CLK(var)
#define LA_ACC(idx, a, b)
#endif

namespace {

constexpr int kM = 128;        // query rows per CTA
constexpr int kN = 176;        // key rows per tile (skip-list granularity, tile_size.h:35-40)
constexpr int kD = 128;        // head dim
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;

constexpr uint32_t kQBlockBytes = kM * 128;   // one 64-column (128 B) swizzled block of Q
constexpr uint32_t kKVBlockBytes = kN * 128;  // one 64-column block of a K or V tile (22528 = 22 * 1024)
constexpr uint32_t kQBytes = 2 * kQBlockBytes;
constexpr uint32_t kKVBytes = 2 * kKVBlockBytes;

constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffK = kOffQ + kQBytes;
constexpr uint32_t kOffV = kOffK + 2 * kKVBytes;
constexpr uint32_t kOffBar = kOffV + 2 * kKVBytes;
constexpr uint32_t kOffMbox = kOffBar + 256;          // float[2][128]  running max mailbox (A->B, B->A)
constexpr uint32_t kOffFin = kOffMbox + 2 * kM * 4;   // float[2][2][128] final (l, m_ref) per warpgroup
constexpr uint32_t kOffStat = kOffFin + 4 * kM * 4;
constexpr uint32_t kOffSeq = kOffStat + kFwdMaxTiles * 4;
constexpr uint32_t kSmemUsed = kOffSeq + kFwdMaxTiles * 2;
static_assert(kSmemUsed + 1024 <= 232448, "shared memory budget (227 KB) exceeded");
static_assert(kFwdSmemBytes == kSmemUsed + 1024, "keep la_kernels.h in sync");

enum Bar : uint32_t {
  kBarQFull = 0,
  kBarKFull = 1,   // +stage
  kBarKEmpty = 3,  // +stage
  kBarVFull = 5,
  kBarVEmpty = 7,
  kBarSFull = 9,   // +buf
  kBarPFull = 11,  // +buf
  kBarPvDone = 13, // one completion per visited tile
  kBarOFinal = 14, // completes once, after the last PV
  kNumBars = 15
};

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemS = 0;        // S buffers at 0 and 176
constexpr uint32_t kTmemO = 2 * kN;   // 352

constexpr uint32_t kIdescQK = make_idesc_bf16(kM, kN, /*b_mn_major=*/0);
constexpr uint32_t kIdescPV = make_idesc_bf16(kM, kD, /*b_mn_major=*/1);

// exp2 on the FMA pipe for these column pairs of every 8 (bit p set => pair p of the group uses the polynomial)
#ifndef LA_POLY_MASK
#define LA_POLY_MASK 0x94u   // pairs 2, 4, 7 -> 3/8 of the elements
#endif
constexpr uint32_t kPolyMask = LA_POLY_MASK;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x (low 16 bits) = lo
  return *reinterpret_cast<uint32_t*>(&v);
}

// 2^t for a pair, t <= 0, on the FMA pipe: n = round(t) via the 1.5*2^23 trick, r = t - n in [-0.5, 0.5],
// 2^r by a degree-3 minimax polynomial (max rel. error 7.5e-5), exponent patched in with one IMAD per element.
__device__ __forceinline__ void exp2_poly_pair(float t0, float t1, float& p0, float& p1) {
  const float kMagic = 12582912.f;  // 1.5 * 2^23
  t0 = fmaxf(t0, -126.f);
  t1 = fmaxf(t1, -126.f);
  const uint64_t t = pack2(t0, t1);
  const uint64_t xf = fadd2(t, pack2(kMagic, kMagic));
  const uint64_t n = fadd2(xf, pack2(-kMagic, -kMagic));
  const uint64_t r = ffma2(n, pack2(-1.f, -1.f), t);
  uint64_t p = ffma2(pack2(0.05517167f, 0.05517167f), r, pack2(0.24261113f, 0.24261113f));
  p = ffma2(p, r, pack2(0.69326097f, 0.69326097f));
  p = ffma2(p, r, pack2(0.99992806f, 0.99992806f));
  float x0, x1, q0, q1;
  unpack2(xf, x0, x1);
  unpack2(p, q0, q1);
  p0 = __int_as_float(__float_as_int(x0) * (1 << 23) + __float_as_int(q0));
  p1 = __int_as_float(__float_as_int(x1) * (1 << 23) + __float_as_int(q1));
}

// One chunk of pass 2: NC S columns (already in registers) -> P (bf16 pairs), returns the chunk's row-sum part.
template <int NC>
__device__ __forceinline__ float softmax_chunk(const float* s, uint32_t* pr, float c, float neg_mc) {
  const uint64_t c2 = pack2(c, c);
  const uint64_t nm2 = pack2(neg_mc, neg_mc);
  uint64_t acc = pack2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < NC; j += 2) {
    float t0, t1, p0, p1;
    unpack2(ffma2(pack2(s[j], s[j + 1]), c2, nm2), t0, t1);
    if ((kPolyMask >> ((j >> 1) & 7)) & 1u) {
      exp2_poly_pair(t0, t1, p0, p1);
    } else {
      p0 = ex2_approx(t0);
      p1 = ex2_approx(t1);
    }
    acc = fadd2(acc, pack2(p0, p1));  // row sum uses fp32 P, before bf16 rounding (softmax.h:263-273)
    pr[j >> 1] = pack_bf16(p0, p1);
  }
  float a0, a1;
  unpack2(acc, a0, a1);
  return a0 + a1;
}

// Cold path: write -inf over the S columns >= lim of this thread's TMEM lane (ragged first tile).
__device__ __noinline__ void mask_ragged_tile(uint32_t s_addr, int lim) {
#pragma unroll 1
  for (int c0 = lim & ~7; c0 < kN; c0 += 8) {
    float v[8];
    if (c0 < lim) {
      tmem_ld_x8(s_addr + c0, reinterpret_cast<uint32_t*>(v));
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j >= lim) v[j] = -INFINITY;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = -INFINITY;
    }
    tmem_st_x8(s_addr + c0, reinterpret_cast<uint32_t*>(v));
  }
  tmem_wait_st();
}

template <int NC>
__device__ __forceinline__ float chunk_max(const float* s, float m) {
  float m0 = m, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j + 3 < NC; j += 4) {
    m0 = fmax3(m0, s[j], s[j + 1]);
    m1 = fmax3(m1, s[j + 2], s[j + 3]);
  }
  return fmaxf(m0, m1);
}

}  // namespace

__global__ void __launch_bounds__(kFwdThreads, 1)
la_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const FwdKernelArgs args) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzle atoms need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_block = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int head_kv = head / args.h_per_kv;

  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kOffBar + kNumBars * 8);
  volatile int* num_tiles_smem = reinterpret_cast<volatile int*>(smem + kOffBar + kNumBars * 8 + 4);
  volatile float* mbox = reinterpret_cast<volatile float*>(smem + kOffMbox);
  volatile float* fin = reinterpret_cast<volatile float*>(smem + kOffFin);
  int* stat_s = reinterpret_cast<int*>(smem + kOffStat);
  uint16_t* seq = reinterpret_cast<uint16_t*>(smem + kOffSeq);
  auto bar = [&](uint32_t idx) { return smem_base + kOffBar + idx * 8; };

  // ------------------------------------------------------------------ one-time setup
  if (threadIdx.x == 0) {
    mbar_init(bar(kBarQFull), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarKFull + s), 1);
      mbar_init(bar(kBarKEmpty + s), 1);
      mbar_init(bar(kBarVFull + s), 1);
      mbar_init(bar(kBarVEmpty + s), 1);
      mbar_init(bar(kBarSFull + s), 1);
      mbar_init(bar(kBarPFull + s), 4);  // one arrival per warp of the owning warpgroup
    }
    mbar_init(bar(kBarPvDone), 1);
    mbar_init(bar(kBarOFinal), 1);
    fence_mbar_init();
  }
  if (warp == kProducerWarp) {
    // Decode the read-list row into the flat sequence of K tiles this CTA visits.
    if (lane == 0) {
      prefetch_tmap(&tmap_q);
      prefetch_tmap(&tmap_k);
      prefetch_tmap(&tmap_v);
    }
    int count = 0;
    if (args.read_list != nullptr) {
      const int32_t* row = args.read_list +
                           ((int64_t)(batch * args.h + head) * args.qtiles + m_block) * (int64_t)(args.ktiles + 1);
      int len = row[0];
      len = min(max(len, 0), args.ktiles) & ~1;
      if (args.ktiles == 1) {
        // A one-tile row is [len, 0]: the range end would live in the next row's slot.  The reference
        // still visits tile 0 (the first listed tile is processed unconditionally, mainloop :1612-1663).
        len = 0;
        if (row[0] > 0) {
          if (lane == 0) seq[0] = 0;
          count = 1;
        }
      }
      for (int r = 0; r < len; r += 2) {
        int s = row[1 + r], e = row[2 + r];
        // The reference does not validate list contents (a malformed list is an OOB tile index there);
        // here out-of-range ranges are clamped so that the kernel can never read outside K/V.
        s = min(s, args.ktiles - 1);
        e = max(e, 0);
        const int nt = s - e + 1;
        if (nt <= 0) continue;
        const int room = kFwdMaxTiles - count;
        const int take = min(nt, room);
        for (int j = lane; j < take; j += 32) seq[count + j] = (uint16_t)(s - j);
        count += take;
      }
    } else {
      count = min(args.ktiles, kFwdMaxTiles);
      for (int j = lane; j < count; j += 32) seq[j] = (uint16_t)(args.ktiles - 1 - j);
    }
    const int neg_inf_ord = float_to_ordered(-INFINITY);
    for (int j = lane; j < count; j += 32) stat_s[j] = (j == 0) ? float_to_ordered(INFINITY) : neg_inf_ord;
    if (lane == 0) *num_tiles_smem = count;
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t tmem_base = *tmem_ptr_smem;
  const int T = *num_tiles_smem;

  if (warp == kProducerWarp) {
    // ================================================================ TMA producer (one lane)
    if (lane == 0 && T > 0) {
      mbar_arrive_expect_tx(bar(kBarQFull), kQBytes);
      tma_load_4d(smem_base + kOffQ, &tmap_q, bar(kBarQFull), 0, m_block * kM, head, batch);
      tma_load_4d(smem_base + kOffQ + kQBlockBytes, &tmap_q, bar(kBarQFull), 64, m_block * kM, head, batch);

      auto load_kv = [&](const CUtensorMap* tm, uint32_t off, uint32_t full0, uint32_t empty0, int i) {
        const int s = i & 1;
        const int n = seq[i];
        mbar_wait(bar(empty0 + s), ((i >> 1) & 1) ^ 1, 1, i);
#ifdef LA_EXPERIMENT_NOLOAD   // timing experiment only: reuse the first two tiles' smem, no further TMA traffic
        if (i >= 2) {
          mbar_arrive(bar(full0 + s));
          return;
        }
#endif
        mbar_arrive_expect_tx(bar(full0 + s), kKVBytes);
        const uint32_t dst = smem_base + off + s * kKVBytes;
        tma_load_4d(dst, tm, bar(full0 + s), 0, n * kN, head_kv, batch);
        tma_load_4d(dst + kKVBlockBytes, tm, bar(full0 + s), 64, n * kN, head_kv, batch);
      };
      load_kv(&tmap_k, kOffK, kBarKFull, kBarKEmpty, 0);
      for (int i = 0; i < T; ++i) {
        if (i + 1 < T) load_kv(&tmap_k, kOffK, kBarKFull, kBarKEmpty, i + 1);
        load_kv(&tmap_v, kOffV, kBarVFull, kBarVEmpty, i);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================================================ tcgen05 issuer (one lane)
    if (lane == 0 && T > 0) {
      const uint64_t q_desc = make_smem_desc_sw128(smem_base + kOffQ, 16, 1024);
      auto issue_qk = [&](int i) {
        const int s = i & 1;
        mbar_wait(bar(kBarKFull + s), (i >> 1) & 1, 2, i);
        tc_fence_after();
        const uint64_t k_desc = make_smem_desc_sw128(smem_base + kOffK + s * kKVBytes, 16, 1024);
        const uint32_t d_tmem = tmem_base + kTmemS + s * kN;
#pragma unroll
        for (int j = 0; j < kD / 16; ++j) {
          // K-major SW128: 4 k-steps of 32 B inside a 128 B swizzle row, then the next 64-column block.
          const uint32_t a_off = ((j >> 2) * kQBlockBytes + (j & 3) * 32) >> 4;
          const uint32_t b_off = ((j >> 2) * kKVBlockBytes + (j & 3) * 32) >> 4;
          umma_ss(d_tmem, q_desc + a_off, k_desc + b_off, kIdescQK, j > 0);
        }
        tc_commit(bar(kBarKEmpty + s));  // K stage reusable once these MMAs retire
        tc_commit(bar(kBarSFull + s));   // S(i) ready for its softmax warpgroup
      };
#ifdef LA_PROFILE_CLOCKS
      long long prof_acc[4] = {0, 0, 0, 0};
#endif
      mbar_wait(bar(kBarQFull), 0, 3, 0);
      issue_qk(0);
      for (int i = 0; i < T; ++i) {
        LA_CLK(m0);
        if (i + 1 < T) issue_qk(i + 1);
        const int s = i & 1;
        LA_CLK(m1);
        LA_ACC(0, m0, m1);
        mbar_wait(bar(kBarVFull + s), (i >> 1) & 1, 4, i);
        LA_CLK(m2);
        LA_ACC(1, m1, m2);
        mbar_wait(bar(kBarPFull + s), (i >> 1) & 1, 5, i);
        tc_fence_after();
        LA_CLK(m3);
        LA_ACC(2, m2, m3);
        // V tile is [176 kv rows][64 d] x 2 blocks, i.e. the MN-major B operand:
        //   LBO = distance between the two 64-wide d blocks, SBO = 8 kv rows (1024 B); one k-step = 16 rows.
        const uint64_t v_desc = make_smem_desc_sw128(smem_base + kOffV + s * kKVBytes, kKVBlockBytes, 1024);
        const uint32_t p_tmem = tmem_base + kTmemS + s * kN;
#pragma unroll
        for (int j = 0; j < kN / 16; ++j) {
          umma_ts(tmem_base + kTmemO, p_tmem + j * 8, v_desc + ((j * 16 * 128) >> 4), kIdescPV, (i > 0) || (j > 0));
        }
        tc_commit(bar(kBarVEmpty + s));
        tc_commit(bar(kBarPvDone));
        LA_CLK(m4);
        LA_ACC(3, m3, m4);
      }
      tc_commit(bar(kBarOFinal));
#ifdef LA_PROFILE_CLOCKS
      for (int j = 0; j < 4; ++j) atomicAdd(&g_la_prof[8 + j], (unsigned long long)prof_acc[j]);
#endif
    }
    __syncwarp();
  } else {
    // ================================================================ softmax warpgroups (2 x 128 threads)
    const int wg = warp >> 2;                 // 0 = A (even visited tiles, S buffer 0), 1 = B (odd, buffer 1)
    const int wq = warp & 3;
    const int row = wq * 32 + lane;           // TMEM lane == query row inside the tile
    const uint32_t lane_field = (uint32_t)(wq * 32) << 16;
    const float c = args.scale_log2;
    const int q_row = m_block * kM + row;
    const uint32_t s_addr = tmem_base + kTmemS + wg * kN + lane_field;
    const uint32_t o_addr = tmem_base + kTmemO + lane_field;
    const uint32_t bar_recv = 1 + (1 - wg) * 4 + wq;   // the other warpgroup arrives here after publishing its max
    const uint32_t bar_send = 1 + wg * 4 + wq;

    float m_ref = -INFINITY;  // running max this warpgroup's l is expressed against (raw S units)
    float l_run = 0.f;        // row sum over this warpgroup's tiles

#ifdef LA_PROFILE_CLOCKS
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int i = wg; i < T; i += 2) {
      const int n = seq[i];
      LA_CLK(t0);
      mbar_wait(bar(kBarSFull + wg), (i >> 1) & 1, 6, i);
      tc_fence_after();
      LA_CLK(t1);
      LA_ACC(0, t0, t1);
#ifdef LA_EXPERIMENT_MMAONLY   // timing experiment only: no softmax work at all, P = whatever is in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarPFull + wg));
      continue;
#endif

      // The FIRST visited tile is the only one that is seqlen-masked (mask.h:66-76, mainloop :1626): when it is
      // ragged, -inf is written over its out-of-range S columns in TMEM once (cold, compact loop), so the hot
      // passes below carry no mask code at all (the fully unrolled masked variants cost more in instruction
      // fetch stalls than in arithmetic -- ncu: 20-45 % "no instruction" in the softmax passes).
      if (i == 0) {
        const int lim = args.seqlen_k - n * kN;
        if (lim < kN) mask_ragged_tile(s_addr, lim);
      }

      // ---------------- pass 1: row max of the tile
      float m_loc;
      {
        float s[64];
        uint32_t* sr = reinterpret_cast<uint32_t*>(s);
        tmem_ld_x64(s_addr, sr);
        tmem_wait_ld();
        m_loc = chunk_max<64>(s, -INFINITY);
        tmem_ld_x64(s_addr + 64, sr);
        tmem_wait_ld();
        m_loc = chunk_max<64>(s, m_loc);
        tmem_ld_x32(s_addr + 128, sr);
        tmem_ld_x16(s_addr + 160, sr + 32);
        tmem_wait_ld();
        m_loc = chunk_max<48>(s, m_loc);
      }

      // ---------------- running max: receive m(i-1) from the other warpgroup, publish m(i)
      LA_CLK(t2);
      LA_ACC(1, t1, t2);
      float m_prev = -INFINITY;
      if (i > 0) {
        named_bar_sync(bar_recv, 64);
        m_prev = mbox[(1 - wg) * kM + row];
      }
      LA_CLK(t3);
      LA_ACC(2, t2, t3);
      const float m_new = fmaxf(m_prev, m_loc);
      if (i + 1 < T) {
        mbox[wg * kM + row] = m_new;
        named_bar_arrive(bar_send, 64);
      }
      if (i > 0) {
        // QK-skip statistic: (m_local - m_prev) * scale_log2, reduced with max over the tile's rows.
        const float d = __fmul_rn(__fsub_rn(m_loc, m_prev), c);
        int od = (d != d) ? float_to_ordered(-INFINITY) : float_to_ordered(d);  // NaN compares false upstream
        od = __reduce_max_sync(0xffffffffu, od);
        if (lane == 0) atomicMax(&stat_s[i], od);
      }
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha_o = ex2_approx((m_prev - m_safe) * c);   // O is expressed against m(i-1)
      const float alpha_l = ex2_approx((m_ref - m_safe) * c);    // this warpgroup's l against its previous tile
      m_ref = m_new;
      const float neg_mc = -m_safe * c;

      // ---------------- pass 2: P = exp2(S*c - m*c), row sum, bf16 P written over the S buffer
      // Five chunks of 32 S columns + one of 16; P chunk k (16 columns of bf16 pairs) lands on columns
      // [16k, 16k+16) -- always behind the S columns [32k, 32k+32) still to be read, and only this thread
      // touches this TMEM lane.  Two register buffers: the next chunk's tcgen05.ld is in flight while the
      // current one is exponentiated.  The loop is kept rolled (two chunk bodies + tail) to stay inside the
      // instruction cache; 32-wide chunks keep the per-chunk MUFU-latency drain to six per tile.
      float lsum = 0.f;
      {
        float sa[32], sb[32];
        uint32_t pr[16];
        tmem_ld_x32(s_addr, reinterpret_cast<uint32_t*>(sa));
        tmem_wait_ld();
#pragma unroll 1
        for (int k = 0; k < 4; k += 2) {
          tmem_ld_x32(s_addr + 32 * (k + 1), reinterpret_cast<uint32_t*>(sb));
          lsum += softmax_chunk<32>(sa, pr, c, neg_mc);
          tmem_st_x16(s_addr + 16 * k, pr);
          tmem_wait_ld();
          tmem_ld_x32(s_addr + 32 * (k + 2), reinterpret_cast<uint32_t*>(sa));
          lsum += softmax_chunk<32>(sb, pr, c, neg_mc);
          tmem_st_x16(s_addr + 16 * (k + 1), pr);
          tmem_wait_ld();
        }
        // sa = chunk 4 (columns 128..159); the 16-column tail goes through sb
        tmem_ld_x16(s_addr + 160, reinterpret_cast<uint32_t*>(sb));
        lsum += softmax_chunk<32>(sa, pr, c, neg_mc);
        tmem_st_x16(s_addr + 64, pr);
        tmem_wait_ld();
        lsum += softmax_chunk<16>(sb, pr, c, neg_mc);
        tmem_st_x8(s_addr + 80, pr);
      }
      l_run = l_run * alpha_l + lsum;
      LA_CLK(t4);
      LA_ACC(3, t3, t4);

      if (i > 0) {
        // O may only be touched between PV(i-1) retiring and PV(i) being issued.
        mbar_wait(bar(kBarPvDone), (i - 1) & 1, 7, i);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha_o != 1.0f)) {
#pragma unroll 1
          for (int h = 0; h < 4; ++h) {
            float o[32];
            tmem_ld_x32(o_addr + h * 32, reinterpret_cast<uint32_t*>(o));
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] *= alpha_o;
            tmem_st_x32(o_addr + h * 32, reinterpret_cast<uint32_t*>(o));
          }
        }
      }
      LA_CLK(t5);
      LA_ACC(4, t4, t5);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarPFull + wg));
      LA_CLK(t6);
      LA_ACC(5, t5, t6);
    }
#ifdef LA_PROFILE_CLOCKS
    if (lane == 0 && wq == 0) {
      for (int j = 0; j < 6; ++j) atomicAdd(&g_la_prof[j], (unsigned long long)prof_acc[j]);
      if (wg == 0) atomicAdd(&g_la_prof[15], (unsigned long long)T);
    }
#endif

    // ---------------- epilogue: merge the two warpgroups' (l, m_ref), normalise, store
    fin[(wg * 2 + 0) * kM + row] = l_run;
    fin[(wg * 2 + 1) * kM + row] = m_ref;
    named_bar_sync(9 + wq, 64);
    float inv = 0.f, lse = -INFINITY;
    if (T > 0) {
      const float l_o = fin[((1 - wg) * 2 + 0) * kM + row];
      const float m_o = fin[((1 - wg) * 2 + 1) * kM + row];
      const float m_fin = fmaxf(m_ref, m_o);                      // the running max is monotone
      const float m_fs = (m_fin == -INFINITY) ? 0.f : m_fin;
      const float l_tot = l_run * ex2_approx((m_ref - m_fs) * c) + l_o * ex2_approx((m_o - m_fs) * c);
      const bool bad = (l_tot == 0.f) || (l_tot != l_tot);
      inv = bad ? 0.f : 1.0f / l_tot;                             // softmax.h:283-293
      lse = bad ? -INFINITY : m_fin * args.softmax_scale + logf(l_tot);
      mbar_wait(bar(kBarOFinal), 0, 8, T);
      tc_fence_after();
    }
    // warpgroup A stores O[:, 0:64), B stores O[:, 64:128)
    float o[64];
    if (T > 0) {
      tmem_ld_x64(o_addr + wg * 64, reinterpret_cast<uint32_t*>(o));
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) o[j] = 0.f;
    }
    if (q_row < args.seqlen_q) {
      __nv_bfloat16* optr = args.out + (int64_t)batch * args.o_batch_stride + (int64_t)q_row * args.o_row_stride +
                            (int64_t)head * args.o_head_stride + wg * 64;
      uint4* dst = reinterpret_cast<uint4*>(optr);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 v;
        v.x = pack_bf16(o[8 * j + 0] * inv, o[8 * j + 1] * inv);
        v.y = pack_bf16(o[8 * j + 2] * inv, o[8 * j + 3] * inv);
        v.z = pack_bf16(o[8 * j + 4] * inv, o[8 * j + 5] * inv);
        v.w = pack_bf16(o[8 * j + 6] * inv, o[8 * j + 7] * inv);
        dst[j] = v;
      }
      if (wg == 0 && args.lse != nullptr)
        args.lse[((int64_t)batch * args.h + head) * args.seqlen_q + q_row] = lse;
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (args.tile_stat != nullptr) {
    float* stat_row = args.tile_stat + ((int64_t)(batch * args.h + head) * args.qtiles + m_block) * (int64_t)args.ktiles;
    for (int j = threadIdx.x; j < T; j += kFwdThreads) stat_row[seq[j]] = ordered_to_float(stat_s[j]);
  }
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace la
