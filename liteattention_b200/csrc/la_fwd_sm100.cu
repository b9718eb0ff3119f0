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
//   * 12 warps: 8 softmax warps (two warpgroups that split the 176 S columns 88/88; one TMEM lane = one query row
//     per thread), 1 TMA producer warp, 1 tcgen05 issuer warp (+2 idle warps that complete its warpgroup, so that
//     setmaxnreg can move registers: 216 per softmax thread, 72 per producer/issuer thread).
//   * Q (128x128), K and V tiles (176x128, 2 stages each) are TMA-loaded into 128B-swizzled smem.
//   * S = Q K^T : tcgen05.mma kind::f16, M=128 N=176 K=16 x8, SS operands, fp32 accumulator in TMEM;
//     two S buffers so QK^T of tile i+1 runs under the softmax of tile i.
//   * P (bf16) is written back over S in TMEM and fed to O += P V as the TMEM A operand
//     (M=128 N=128 K=16 x11, V is the MN-major smem B operand); O lives in TMEM for the whole CTA.
//   * TMEM map (512 columns allocated): S0 @0, S1 @176, O @352..479.
//   * The issuer warp walks its loop converged and issues through elect.sync, so descriptors sit in uniform
//     registers (19 UTCHMMA per tile back to back instead of an ELECT + R2UR chain per instruction).
//   * Softmax, per tile (tools/prof_clocks.py, tools/softmax_rate.cu, tools/issue_rate.cu give the budget: MUFU.EX2
//     is 16/clk/SM = 1408 clk for a 128x176 tile, exactly the tile's MMA time; FFMA2/FADD2/FMNMX3/F2FP dispatch at
//     one warp-instruction per 2 clk per SM sub-partition, 3 of them per element = ~1060 clk):
//       - lazy scaling reference: P = 2^((S - m_ref) c) with m_ref trailing the TRUE running max by <= 8 (log2
//         units); the exponentials never wait for the tile's own max and O is almost never rescaled.  The QK-skip
//         statistic is always computed from the true running max.
//       - the verdict (does any row max run ahead of m_ref by more than 8?) needs the full-row max: the two warps
//         that own a row exchange half-row maxima through smem and one named barrier.  If it fails,
//         the tile is redone exactly, out of line (softmax_slow_tile): m_ref := true max, l and O rescaled.
//       - 2 of every 8 column pairs (staggered between the two column halves) take their exp2 on the FMA pipe (Cody-Waite + cubic), which balances the MUFU
//         pipe against instruction dispatch; S(i+1) is pulled from TMEM while P(i) is being published; the
//         statistic of tile i is reduced inside tile i+1's exponential loop.
//
// Measured ceiling of this organisation (one 128-row Q tile per CTA, 512 TMEM columns = 2 x S(176) + O(128)):
// the tensor pipe alone runs the tile's 19 MMAs at the 1408-clk floor (tools/mma_rate.cu), L2->smem delivers
// 75-80 B/clk/SM against the 64 needed (tools/l2_rate.cu); the softmax chain S(i) -> P(i) is ~2000 clk, and with
// only two S buffers PV(i) -> QK(i+2) cannot be decoupled from it.  See DESIGN.md section 3.1.
#include <cuda_bf16.h>

#include <type_traits>

#include "la_kernels.h"
#include "la_ptx.cuh"
#include "la_tmem_ptx.cuh"

namespace la {

// Developer-only cycle accounting (tools/prof_clocks.py builds a variant with -DLA_PROFILE_CLOCKS): clock64 deltas
// of the role leaders summed per phase.  Compiled out of the product build.
#ifdef LA_PROFILE_CLOCKS
__device__ unsigned long long g_la_prof[32];
#define LA_CLK(var) const long long var = clock64()
#define LA_ACC(idx, a, b) prof_acc[idx] += (b) - (a)
#define LA_PROF_DECL(n) long long prof_acc[n] = {}
#define LA_PROF_FLUSH(base, n) for (int j_ = 0; j_ < (n); ++j_) atomicAdd(&g_la_prof[(base) + j_], (unsigned long long)prof_acc[j_])
#else
#define LA_CLK(var)
#define LA_ACC(idx, a, b)
#define LA_PROF_DECL(n)
#define LA_PROF_FLUSH(base, n)
#endif

namespace {

constexpr int kM = 128;        // query rows per CTA
constexpr int kN = 176;        // key rows per tile (skip-list granularity, tile_size.h:35-40)
constexpr int kD = 128;        // head dim
constexpr int kHalfN = kN / 2; // S columns per softmax warpgroup
constexpr int kSoftmaxThreads = 256;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kRegsSoftmax = 216;   // 8 warps x 32 x 216 + 4 warps x 32 x 72 = 64512 = the 12 x 32 x 168 the CTA launches with
constexpr int kRegsOther = 72;

constexpr uint32_t kQBlockBytes = kM * 128;   // one 64-column (128 B) swizzled block of Q
constexpr uint32_t kKVBlockBytes = kN * 128;  // one 64-column block of a K or V tile (22528 = 22 * 1024)
constexpr uint32_t kQBytes = 2 * kQBlockBytes;
constexpr uint32_t kKVBytes = 2 * kKVBlockBytes;

constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffK = kOffQ + kQBytes;
constexpr uint32_t kOffV = kOffK + 2 * kKVBytes;
constexpr uint32_t kOffBar = kOffV + 2 * kKVBytes;
constexpr uint32_t kOffXchg = kOffBar + 256;
constexpr uint32_t kOffLx = kOffXchg + 2 * kSoftmaxThreads * 8;   // {tile tag, half-row max} per thread and S buffer
constexpr uint32_t kOffStat = kOffLx + kSoftmaxThreads * 4;
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
  kBarOFull = 13,   // the last PV retired: O is final
  kNumBars = 14
};

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemS = 0;        // S buffers at 0 and 176
constexpr uint32_t kTmemO = 2 * kN;   // 352

constexpr uint32_t kIdescQK = make_idesc_bf16(kM, kN, /*b_mn_major=*/0);
constexpr uint32_t kIdescPV = make_idesc_bf16(kM, kD, /*b_mn_major=*/1);

// Cold path: write -inf over the S columns >= lim of this thread's TMEM lane (ragged first tile), in TMEM,
// before the hot path loads S -- so the hot path carries no mask code and no register array escapes.
__device__ __noinline__ void mask_ragged_tile(uint32_t s_addr, int lim, int ncols) {
#pragma unroll 1
  for (int c0 = lim & ~7; c0 < ncols; c0 += 8) {
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

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x (low 16 bits) = lo
  return *reinterpret_cast<uint32_t*>(&v);
}

// exp2 on the FMA pipe for these column pairs of every 8 (bit p set => pair p of the group uses the polynomial).
// MUFU.EX2 runs at 16/clk/SM: 128x176 exponentials take exactly as long as the tile's two MMAs (1408 clk), so a
// share of them has to come off the MUFU pipe for the softmax to fit under the tensor pipe at all.
#ifndef LA_POLY_MASK
#define LA_POLY_MASK 0x11u   // pairs 0 and 4 of every 8 for the first column half ...
#endif
constexpr uint32_t kPolyMask = LA_POLY_MASK;
#ifndef LA_POLY_MASK_B
#define LA_POLY_MASK_B 0x44u  // ... pairs 2 and 6 for the second: the two warps of a sub-partition run in near lock-step,
                              // staggering their FMA-pipe pairs is worth 2 % (same-box A/B, tools/build_variants.py)
#endif
constexpr uint32_t kPolyMaskB = LA_POLY_MASK_B;

// The scaling reference m_ref of a row trails its true running max by at most kLazyTau (log2 units): P <= 2^tau.
// While the true max stays within tau of m_ref, P(i) does not depend on tile i's own max, so the exponentials
// start as soon as S arrives and O never needs rescaling.  (The QK-skip statistic always uses the TRUE max.)
#ifndef LA_LAZY_TAU
#define LA_LAZY_TAU 8.0f    // larger is ~2 % faster on video-like data (fewer exact tiles) but costs accuracy: with the true
                            // max the dominant P of a row is exactly 1.0, with a stale reference it is 2^x rounded to bf16
                            // (2^-9 relative); tau = 32 fails tests/test_fwd_gpu.py::test_edge_layouts' tolerance by 12 %
#endif
constexpr float kLazyTau = LA_LAZY_TAU;

// 2^t for a pair on the FMA pipe: n = round(t) via the 1.5*2^23 trick, r = t - n in [-0.5, 0.5], 2^r by a
// degree-3 minimax polynomial (max rel. error 7.5e-5, below bf16 rounding of P), exponent patched in with one
// integer multiply-add per element.
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

struct SlowTileArgs {
  uint32_t s_addr;      // this thread's half of S (TMEM, lane field included)
  uint32_t p_addr;      // where this thread's 44 bf16x2 P columns go
  uint32_t o_addr;      // this thread's 64 O columns
  uint32_t xchg_mine, xchg_other;  // smem byte addresses of the half-row max exchange slots
  uint32_t bar_id;      // named barrier of the warp pair owning these 32 rows
  uint32_t bar_pv_done; // mbarrier that flips when PV(i-1) retires (the V-empty barrier of its stage) ...
  uint32_t pv_parity;   // ... and the parity of that phase
  int i;                // visit index of the tile
  int mask_lim;         // < kHalfN: columns >= mask_lim of this half are out of range (first tile only)
  float c;
  float m_loc;          // in: full-row max of this tile if known (have_mloc); out: always
  int have_mloc;
  float m_true;         // in: true running max before this tile
  float m_ref;          // in/out: scaling reference
  float l_run;          // in/out: partial row sum (this thread's columns), relative to m_ref
};

// The exact (non-speculative) tile: first visited tile of every CTA, and any tile whose row max runs more than
// kLazyTau ahead of the reference.  m_ref := true running max, l and O are rescaled, P is recomputed from S
// (still intact in TMEM) entirely on MUFU.  Out of line: it runs for ~0.5 % of the tiles.
__device__ __noinline__ void softmax_slow_tile(SlowTileArgs* a) {
  const float c = a->c;
  if (a->mask_lim < kHalfN) mask_ragged_tile(a->s_addr, max(a->mask_lim, 0), kHalfN);
  float s[kHalfN];
  uint32_t* sr = reinterpret_cast<uint32_t*>(s);
  tmem_ld_x32(a->s_addr, sr);
  tmem_ld_x32(a->s_addr + 32, sr + 32);
  tmem_ld_x16(a->s_addr + 64, sr + 64);
  tmem_ld_x8(a->s_addr + 80, sr + 80);
  tmem_wait_ld();
  float m_loc = a->m_loc;
  if (!a->have_mloc) {
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < kHalfN; j += 4) {
      mx0 = fmax3(mx0, s[j], s[j + 1]);
      mx1 = fmax3(mx1, s[j + 2], s[j + 3]);
    }
    const float m_half = fmaxf(mx0, mx1);
    sts_f32(a->xchg_mine, m_half);
    named_bar_sync(a->bar_id, 64);
    m_loc = fmaxf(m_half, lds_f32(a->xchg_other));
    a->m_loc = m_loc;
  } else {
    // The partner warp's P lands on S columns this warp has just re-read: order its store after our load.
    named_bar_sync(a->bar_id, 64);
  }
  const float m_new = fmaxf(a->m_true, m_loc);
  const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
  const float alpha = ex2_approx((a->m_ref - m_safe) * c);   // m_ref = -inf before the first tile -> 0
  a->m_ref = m_safe;
  const float neg_mc = -m_safe * c;
  uint32_t pr[kHalfN / 2];
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int j = 0; j < kHalfN; j += 2) {
    const float p0 = ex2_approx(fmaf(s[j], c, neg_mc)), p1 = ex2_approx(fmaf(s[j + 1], c, neg_mc));
    sum0 += p0;
    sum1 += p1;
    pr[j / 2] = pack_bf16(p0, p1);
  }
  a->l_run = a->l_run * alpha + (sum0 + sum1);
  if (a->i > 0) {
    // O may only be touched between PV(i-1) retiring and PV(i) being issued.
    mbar_wait(a->bar_pv_done, a->pv_parity, 7, a->i);
    tc_fence_after();
    if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        float o[16];
        tmem_ld_x16(a->o_addr + h * 16, reinterpret_cast<uint32_t*>(o));
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] *= alpha;
        tmem_st_x16(a->o_addr + h * 16, reinterpret_cast<uint32_t*>(o));
      }
    }
  }
  tmem_st_x32(a->p_addr, pr);
  tmem_st_x8(a->p_addr + 32, pr + 32);
  tmem_st_x4(a->p_addr + 40, pr + 40);
  tmem_wait_st();
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
  float* lx = reinterpret_cast<float*>(smem + kOffLx);
  int* stat_s = reinterpret_cast<int*>(smem + kOffStat);
  uint16_t* seq = reinterpret_cast<uint16_t*>(smem + kOffSeq);
  auto bar = [&](uint32_t idx) { return smem_base + kOffBar + idx * 8; };

#ifdef LA_PROFILE_CLOCKS
  const long long prof_t_entry = clock64();
#endif
  // ------------------------------------------------------------------ one-time setup
  if (threadIdx.x == 0) {
    mbar_init(bar(kBarQFull), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarKFull + s), 1);
      mbar_init(bar(kBarKEmpty + s), 1);
      mbar_init(bar(kBarVFull + s), 1);
      mbar_init(bar(kBarVEmpty + s), 1);
      mbar_init(bar(kBarSFull + s), 1);
      mbar_init(bar(kBarPFull + s), kSoftmaxThreads / 32);  // one arrival per softmax warp
    }
    mbar_init(bar(kBarOFull), 1);
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

  // 12 warps x 168 registers at launch; the softmax threads need ~200 (88 S values prefetched for the next tile
  // while 44 packed P registers of the current one are still live), the producer / issuer warps almost none.
  // setmaxnreg is a warpgroup-wide instruction: all four warps of a group must run the SAME instance of it.
  if (warp >= 8) {
  setmaxnreg_dec<kRegsOther>();   // warps 10-11 only complete this warpgroup
  if (warp == kProducerWarp) {
    // ================================================================ TMA producer (one lane)
    if (lane == 0 && T > 0) {
      // Optional L2 eviction hints like the reference's (mainloop :1002, :1020 EVICT_LAST on K/V, :1092 EVICT_FIRST on Q):
      // K/V of a head are re-read by its 591 Q tiles out of L2, Q is read once.
#ifndef LA_L2_HINTS
#define LA_L2_HINTS 0     // bit 0: K/V evict_last, bit 1: Q evict_first, bit 2: O stores evict_first
#endif
#if LA_L2_HINTS & 3
      const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
#endif
#define LA_TMA_LOAD_PLAIN(dst, tm, bar_, c0, c1, c2, c3, pol) tma_load_4d(dst, tm, bar_, c0, c1, c2, c3)
#define LA_TMA_LOAD_HINT(dst, tm, bar_, c0, c1, c2, c3, pol) tma_load_4d_hint(dst, tm, bar_, c0, c1, c2, c3, pol)
#if LA_L2_HINTS & 1
#define LA_TMA_LOAD_KV LA_TMA_LOAD_HINT
#else
#define LA_TMA_LOAD_KV LA_TMA_LOAD_PLAIN
#endif
#if LA_L2_HINTS & 2
#define LA_TMA_LOAD_Q LA_TMA_LOAD_HINT
#else
#define LA_TMA_LOAD_Q LA_TMA_LOAD_PLAIN
#endif
      mbar_arrive_expect_tx(bar(kBarQFull), kQBytes);
      LA_TMA_LOAD_Q(smem_base + kOffQ, &tmap_q, bar(kBarQFull), 0, m_block * kM, head, batch, pol_stream);
      LA_TMA_LOAD_Q(smem_base + kOffQ + kQBlockBytes, &tmap_q, bar(kBarQFull), 64, m_block * kM, head, batch, pol_stream);

      LA_PROF_DECL(2);
      auto load_kv = [&](const CUtensorMap* tm, uint32_t off, uint32_t full0, uint32_t empty0, int i) {
        const int s = i & 1;
        const int n = seq[i];
        LA_CLK(p0);
        mbar_wait(bar(empty0 + s), ((i >> 1) & 1) ^ 1, 1, i);
        LA_CLK(p1);
        LA_ACC(empty0 == kBarKEmpty ? 0 : 1, p0, p1);
#ifdef LA_EXPERIMENT_NOLOAD   // timing experiment only (wrong results): reuse the first two tiles' smem, no further TMA traffic
        if (i >= 2) {
          mbar_arrive(bar(full0 + s));
          return;
        }
#endif
        mbar_arrive_expect_tx(bar(full0 + s), kKVBytes);
        const uint32_t dst = smem_base + off + s * kKVBytes;
        LA_TMA_LOAD_KV(dst, tm, bar(full0 + s), 0, n * kN, head_kv, batch, pol_keep);
        LA_TMA_LOAD_KV(dst + kKVBlockBytes, tm, bar(full0 + s), 64, n * kN, head_kv, batch, pol_keep);
      };
      load_kv(&tmap_k, kOffK, kBarKFull, kBarKEmpty, 0);
      for (int i = 0; i < T; ++i) {
        if (i + 1 < T) load_kv(&tmap_k, kOffK, kBarKFull, kBarKEmpty, i + 1);
        load_kv(&tmap_v, kOffV, kBarVFull, kBarVEmpty, i);
      }
      LA_PROF_FLUSH(16, 2);
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================================================ tcgen05 issuer
    // The whole warp walks the loop converged and one elected lane issues: with warp-uniform control flow the
    // descriptors live in uniform registers.  (Issuing from inside `if (lane == 0)` made ptxas rebuild every
    // operand with ELECT + 4-6 dependent R2UR per MMA, ~80 clk of issue latency per instruction -- more than the
    // MMA itself takes to execute -- and the tensor pipe starved: tools/prof_clocks.py.)
    const int Tu = __reduce_max_sync(0xffffffffu, T);   // same value in every lane; tells nvcc it is uniform
    if (Tu > 0) {
      const uint32_t q_lo = ((smem_base + kOffQ) >> 4) & 0x3FFFu;
      const uint32_t k_lo = ((smem_base + kOffK) >> 4) & 0x3FFFu;
      const uint32_t v_lo = ((smem_base + kOffV) >> 4) & 0x3FFFu;
      constexpr uint64_t kDescKMajorHi = make_smem_desc_sw128(0, 16, 1024);            // Q and K: K-major
      constexpr uint64_t kDescVHi = make_smem_desc_sw128(0, kKVBlockBytes, 1024);      // V: MN-major, see below
      LA_PROF_DECL(5);
      auto issue_qk = [&](int i, int s) {
        LA_CLK(k0);
        mbar_wait(bar(kBarKFull + s), (i >> 1) & 1, 2, i);
        LA_CLK(k1);
        LA_ACC(0, k0, k1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d_tmem = tmem_base + kTmemS + s * kN;
#pragma unroll
          for (int j = 0; j < kD / 16; ++j) {
            // K-major SW128: 4 k-steps of 32 B inside a 128 B swizzle row, then the next 64-column block.
            const uint32_t a_off = ((j >> 2) * kQBlockBytes + (j & 3) * 32) >> 4;
            const uint32_t b_off = (s * kKVBytes + (j >> 2) * kKVBlockBytes + (j & 3) * 32) >> 4;
            umma_ss(d_tmem, kDescKMajorHi | (uint64_t)(q_lo + a_off), kDescKMajorHi | (uint64_t)(k_lo + b_off),
                    kIdescQK, j > 0);
          }
          tc_commit(bar(kBarKEmpty + s));  // K stage reusable once these MMAs retire
          tc_commit(bar(kBarSFull + s));   // S(i) ready for the softmax warps
        }
        __syncwarp();
        LA_CLK(k2);
        LA_ACC(1, k1, k2);
      };
      auto issue_pv = [&](int i, int s) {
        LA_CLK(v0);
        mbar_wait(bar(kBarVFull + s), (i >> 1) & 1, 4, i);
        LA_CLK(v1);
        mbar_wait(bar(kBarPFull + s), (i >> 1) & 1, 5, i);
        LA_CLK(v2);
        LA_ACC(2, v0, v1);
        LA_ACC(3, v1, v2);
        tc_fence_after();
        if (elect_one_sync()) {
          // V tile is [176 kv rows][64 d] x 2 blocks, i.e. the MN-major B operand:
          //   LBO = distance between the two 64-wide d blocks, SBO = 8 kv rows (1024 B); one k-step = 16 rows.
          const uint32_t p_tmem = tmem_base + kTmemS + s * kN;
          umma_ts(tmem_base + kTmemO, p_tmem, kDescVHi | (uint64_t)(v_lo + ((s * kKVBytes) >> 4)), kIdescPV, i > 0);
#pragma unroll
          for (int j = 1; j < kN / 16; ++j) {
            umma_ts(tmem_base + kTmemO, p_tmem + j * 8,
                    kDescVHi | (uint64_t)(v_lo + ((s * kKVBytes + j * 16 * 128) >> 4)), kIdescPV, 1);
          }
          tc_commit(bar(kBarVEmpty + s));   // V stage reusable; also "PV(i) retired" for a tile that must rescale O
          // O-final gets its own barrier: a parity wait is only meaningful one phase behind, and a softmax warp may
          // not have waited on V-empty for many tiles by the time it reaches the epilogue.
          if (i == Tu - 1) tc_commit(bar(kBarOFull));
        }
        __syncwarp();
        LA_CLK(v3);
        LA_ACC(4, v2, v3);
      };
      mbar_wait(bar(kBarQFull), 0, 3, 0);
      issue_qk(0, 0);
      for (int i = 0; i < Tu; i += 2) {   // two tiles per trip so that the stage index is a compile-time constant
        if (i + 1 < Tu) issue_qk(i + 1, 1);
        issue_pv(i, 0);
        if (i + 1 < Tu) {
          if (i + 2 < Tu) issue_qk(i + 2, 0);
          issue_pv(i + 1, 1);
        }
      }
      if (lane == 0) { LA_PROF_FLUSH(8, 5); }
    }
    __syncwarp();
  }
  } else {
    // ================================================================ softmax warps (256 threads)
    setmaxnreg_inc<kRegsSoftmax>();
    const int wg = warp >> 2;                 // column half: 0 -> S[:, 0:88), 1 -> S[:, 88:176)
    const int row = (warp & 3) * 32 + lane;   // TMEM lane == query row inside the tile
    const int tid = threadIdx.x;              // 0..255
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    const float c = args.scale_log2;
    const int q_row = m_block * kM + row;

    float m_true = -INFINITY;  // true running row max (raw S units): the QK-skip statistic is defined on it
    float m_ref = -INFINITY;   // scaling reference: P = 2^((S - m_ref) c); m_true - m_ref <= kLazyTau / c
    float l_run = 0.f;         // this thread's partial row sum (its 88 columns), relative to m_ref

    const uint32_t xchg_base = smem_base + kOffXchg;
    const uint32_t stat_base = smem_base + kOffStat;
    const uint64_t c2 = pack2(c, c);

    // QK-skip statistic of a tile: (m_local - m_prev) * scale_log2 (softmax.h:194) reduced with max over the tile's
    // 128 rows (one warp-level redux + one shared-memory reduction per warp).  It is emitted one tile late, inside
    // the next tile's exponential phase, where its latency is free; the two column halves alternate tiles.
    float stat_d = 0.f;    // this row's (m_local - m_prev) * c of the previous tile
    int stat_i = 0;        // its visit index (0 = nothing pending: the first visited tile has no statistic)
    auto emit_stat = [&]() {
      if (stat_i > 0 && (stat_i & 1) == wg) {
        int od = (stat_d != stat_d) ? float_to_ordered(-INFINITY) : float_to_ordered(stat_d);  // NaN compares false upstream
        od = __reduce_max_sync(0xffffffffu, od);
        if (lane == 0) red_smax_s32(stat_base + stat_i * 4, od);
      }
    };

    // Raw S of the tile being processed (this thread's 88 columns).  It is (re)loaded at the END of the previous
    // tile's iteration -- after that tile's P stores were issued and, on the exact path, after the out-of-line
    // call -- so the TMEM read latency of tile i+1 hides under the tail of tile i and the array is never live
    // across the call.
    float s[kHalfN];
    uint32_t* sr = reinterpret_cast<uint32_t*>(s);
    auto s_tmem = [&](int i) { return tmem_base + kTmemS + (i & 1) * kN + wg * kHalfN + lane_field; };
    auto load_s = [&](int i) {
      const uint32_t a = s_tmem(i);
      tc_fence_after();
      tmem_ld_x32(a, sr);
      tmem_ld_x32(a + 32, sr + 32);
      tmem_ld_x16(a + 64, sr + 64);
      tmem_ld_x8(a + 80, sr + 80);
    };
    // Named barrier of the warp pair that owns the same 32 rows (ids 1..4; 0 is __syncthreads).  Always a full
    // rendezvous (bar.sync by both warps), never arrive + sync: with a split barrier a warp that runs ahead can
    // contribute twice to one phase -- its arrive for tile i and its next use of the same id -- while its partner
    // sits between two instructions (cold instruction cache is enough); the pair then stays one phase apart and
    // the last sync never completes.  (Seen as a 1-in-15 hang on fresh processes; tools/flaky2.py.)
    const uint32_t pair_bar = 1 + (warp & 3);
    auto exact_tile = [&](int i, float m_loc_known) -> float {
      const int buf = i & 1;
      SlowTileArgs a;
      a.s_addr = s_tmem(i);
      a.p_addr = tmem_base + kTmemS + buf * kN + wg * (kHalfN / 2) + lane_field;
      a.o_addr = tmem_base + kTmemO + wg * 64 + lane_field;
      a.xchg_mine = xchg_base + (buf * kSoftmaxThreads + tid) * 8 + 4;
      a.xchg_other = xchg_base + (buf * kSoftmaxThreads + (tid ^ 128)) * 8 + 4;
      a.bar_id = pair_bar;
      // PV(i-1) retiring is what frees V stage (i-1)&1: phase ((i-1)>>1) of that stage's empty barrier.  When tile i
      // is being processed PV(i-2) has retired (S(i) was committed after it) and PV(i+1) cannot have been issued, so
      // the barrier is at most one phase away from the one waited for and the parity test is unambiguous.
      a.bar_pv_done = bar(kBarVEmpty + ((i - 1) & 1));
      a.pv_parity = (uint32_t)(((i - 1) >> 1) & 1);
      a.i = i;
      // Key columns >= seqlen_k are masked in the FIRST processed tile only (mask.h:66-76, mainloop :1626).
      a.mask_lim = (i == 0) ? args.seqlen_k - (seq[0] * kN + wg * kHalfN) : kHalfN;
      a.c = c;
      a.m_loc = m_loc_known;
      a.have_mloc = (i != 0);
      a.m_true = m_true;
      a.m_ref = m_ref;
      a.l_run = l_run;
      softmax_slow_tile(&a);
      m_ref = a.m_ref;
      l_run = a.l_run;
      return a.m_loc;
    };
    auto publish_p = [&](int i) {   // P(i) is in TMEM (and O rescaled if it had to be): let the issuer run PV(i)
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarPFull + (i & 1)));
    };

    // Exchange slots are {tile tag, value}; no hot-loop tile has tag 0xffffffff.
    sts_u32(xchg_base + tid * 8, 0xffffffffu);
    sts_u32(xchg_base + (kSoftmaxThreads + tid) * 8, 0xffffffffu);
    LA_PROF_DECL(7);
#ifdef LA_PROFILE_CLOCKS
    long long prof_t_s0 = 0, prof_t_first = 0, prof_t_loop_end = 0;
#endif
    if (T > 0) {
      // ---------------- first visited tile: always exact (there is no reference yet)
      mbar_wait(bar(kBarSFull + 0), 0, 6, 0);
      tc_fence_after();
#ifdef LA_PROFILE_CLOCKS
      prof_t_s0 = clock64();
#endif
      m_true = exact_tile(0, 0.f);
      publish_p(0);
#ifdef LA_PROFILE_CLOCKS
      prof_t_first = clock64();
#endif
      if (T > 1) {
        mbar_wait(bar(kBarSFull + 1), 0, 6, 1);
        load_s(1);
      }
    }
    for (int i = 1; i < T; ++i) {
      const int buf = i & 1;
      // P (bf16 pairs) goes over the first 88 columns of this S buffer: wg0 -> [0,44), wg1 -> [44,88).
      const uint32_t p_addr = tmem_base + kTmemS + buf * kN + wg * (kHalfN / 2) + lane_field;
      const uint32_t xchg_mine = xchg_base + (buf * kSoftmaxThreads + tid) * 8;
      const uint32_t xchg_other = xchg_base + (buf * kSoftmaxThreads + (tid ^ 128)) * 8;
      const uint32_t next_bar = bar(kBarSFull + (buf ^ 1));
      const uint32_t next_par = ((i + 1) >> 1) & 1;
      const bool more = i + 1 < T;
#ifdef LA_EXPERIMENT_MMAONLY   // timing experiment only (wrong results): no softmax work at all, P = whatever is in TMEM
      tmem_wait_ld();
      publish_p(i);
      if (more) { mbar_wait(next_bar, next_par, 6, i + 1); load_s(i + 1); }
      continue;
#endif
      // ---------------- speculative tile: exponentials against the current reference, row max on the side
      LA_CLK(t0);
      tmem_wait_ld();
      LA_CLK(t1);
      LA_ACC(0, t0, t1);
      const float neg_mc = -m_ref * c;
      const uint64_t nm2 = pack2(neg_mc, neg_mc);
      uint32_t pr[kHalfN / 2];
      uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
      float mx0 = -INFINITY, mx1 = -INFINITY;
      constexpr int kQuads = kHalfN / 4;   // 22 groups of 4 columns
#ifndef LA_XCHG_FLAG
#define LA_XCHG_FLAG 0                     // 1: half-row max exchange through tagged smem slots (round-2 experiment), 0: named barrier
#endif
#ifndef LA_POST_Q
#define LA_POST_Q (LA_XCHG_FLAG ? 18 : 22) // quad index at which the half-row max is posted (22 = after the loop)
#endif
      constexpr int kPostQ = LA_POST_Q;
#ifndef LA_STAT_Q
#define LA_STAT_Q 11                       // quad at which the previous tile's statistic is reduced
#endif
#ifndef LA_EX2_ORDERED
#define LA_EX2_ORDERED 1                   // 1: MUFU statements pinned in program order (volatile asm)
#endif
      float m_half = 0.f, m_loc = 0.f;
      bool ready = false;
      // Half-row max exchange between the two warps that own the same 32 rows: a full-rendezvous named barrier after the
      // loop (default), or -- LA_XCHG_FLAG, a round-2 experiment kept as a knob -- tagged smem slots: the max of the
      // remaining quads is taken ahead of their exponentials and posted at quad LA_POST_Q as {tile index, value}, the
      // partner's slot is polled after the loop; no barrier, and unlike round 1's split arrive/sync barrier it cannot
      // slip a phase (a tag matches exactly one tile).  It is 1.2-1.5 % faster in short bursts (tools/ab.py) and
      // 0.4 % slower inside bench.py's power-capped steady state, where the busier pipes only lower the clock
      // (profiles/ab_r2.txt, calls 28-31); the barrier stays.  Where the statistic is reduced and whether the MUFU
      // statements are pinned in program order measured within +-1 % of each other.
      auto post_and_fetch = [&](int q_from) {
#pragma unroll
        for (int qq = q_from; qq < kQuads; ++qq) {
          mx0 = fmax3(mx0, s[4 * qq], s[4 * qq + 1]);
          mx1 = fmax3(mx1, s[4 * qq + 2], s[4 * qq + 3]);
        }
        m_half = fmaxf(mx0, mx1);
#if LA_XCHG_FLAG
        // Tagged post: {tile index, half-row max} in one 8-byte store.  The partner polls for the tag after its own loop
        // (consume_max below), so nothing blocks here and the exchange latency runs under the remaining exponentials.
        sts_v2_u32(xchg_mine, (uint32_t)i, __float_as_uint(m_half));
        ready = more && mbar_test_wait(next_bar, next_par);
#else
        sts_f32(xchg_mine + 4, m_half);
        // Is S(i+1) there yet?  (Asked here so that the answer's latency runs under the verdict.)
        ready = more && mbar_try_wait(next_bar, next_par);
        // Exchange the half-row maxima between the two warps that own the same 32 rows.  Passing the barrier also
        // means the partner has all of its S columns in registers, the condition for our P to land on them.
        named_bar_sync(pair_bar, 64);
        m_loc = fmaxf(m_half, lds_f32(xchg_other + 4));
#endif
      };
      auto consume_max = [&]() {
#if LA_XCHG_FLAG
        // Seeing the partner's tag for this tile also means it has all of its S columns in registers (it posts after
        // its wait::ld), the condition for our P to land on them.  Two slots (one per S buffer) are enough: the
        // partner cannot post tile i+2 before it has seen our post of tile i+1, which follows this read.
        uint32_t tag, val, spins = 0;
        do {
          lds_v2_u32_volatile(xchg_other, tag, val);
          if (++spins > LA_WATCHDOG_SPINS) __trap();      // a protocol bug must not hang the box (see la_ptx.cuh)
        } while (tag != (uint32_t)i);
        m_loc = fmaxf(m_half, __uint_as_float(val));
        // The poll loop exits lane by lane: reconverge before anything .sync.aligned depends on a per-lane answer
        // (`ready` gates the warp-wide tcgen05.ld of S(i+1)).
        __syncwarp();
        if (more && !ready) ready = mbar_test_wait(next_bar, next_par);
        ready = __all_sync(0xffffffffu, ready);
#endif
      };
      // The two warps that share a sub-partition run this loop in near lock-step; each column half has its own mask so
      // that their FMA-pipe pairs do not coincide.
      auto exp_loop = [&](auto mask_tag) {
        constexpr uint32_t kMask = decltype(mask_tag)::value;
#pragma unroll
      for (int q = 0; q < kQuads; ++q) {
        const int j = 4 * q;
        if (q == kPostQ) post_and_fetch(q);
        if (q < kPostQ) {
          mx0 = fmax3(mx0, s[j], s[j + 1]);
          mx1 = fmax3(mx1, s[j + 2], s[j + 3]);
        }
        if (q == LA_STAT_Q) emit_stat();   // of tile i-1
        float t0, t1, t2, t3, p0, p1, p2, p3;
        unpack2(ffma2(pack2(s[j], s[j + 1]), c2, nm2), t0, t1);
        unpack2(ffma2(pack2(s[j + 2], s[j + 3]), c2, nm2), t2, t3);
        if ((kMask >> ((j >> 1) & 7)) & 1u) {
          exp2_poly_pair(t0, t1, p0, p1);
        } else if (q < kPostQ && LA_EX2_ORDERED) {
          p0 = ex2_approx_ordered(t0);
          p1 = ex2_approx_ordered(t1);
        } else {
          p0 = ex2_approx(t0);
          p1 = ex2_approx(t1);
        }
        if ((kMask >> (((j >> 1) + 1) & 7)) & 1u) {
          exp2_poly_pair(t2, t3, p2, p3);
        } else if (q < kPostQ && LA_EX2_ORDERED) {
          p2 = ex2_approx_ordered(t2);
          p3 = ex2_approx_ordered(t3);
        } else {
          p2 = ex2_approx(t2);
          p3 = ex2_approx(t3);
        }
        acc0 = fadd2(acc0, pack2(p0, p1));   // row sum uses fp32 P, before bf16 rounding (softmax.h:263-273)
        acc1 = fadd2(acc1, pack2(p2, p3));
        pr[j / 2] = pack_bf16(p0, p1);
        pr[j / 2 + 1] = pack_bf16(p2, p3);
      }
      };
      if (kPolyMaskB == kPolyMask || wg == 0) exp_loop(std::integral_constant<uint32_t, kPolyMask>{});
      else exp_loop(std::integral_constant<uint32_t, kPolyMaskB>{});
      if (kPostQ >= kQuads) post_and_fetch(kQuads);
      consume_max();
      LA_CLK(t2);
      LA_ACC(1, t1, t2);
      // Both warps of the pair see the same m_loc and m_ref for the same rows => the same verdict.
      const bool exact = __any_sync(0xffffffffu, !((m_loc - m_ref) * c <= kLazyTau));
      LA_CLK(t3);
      LA_ACC(2, t2, t3);
      if (!exact) {
        float a0, a1, a2, a3;
        unpack2(acc0, a0, a1);
        unpack2(acc1, a2, a3);
        l_run += (a0 + a1) + (a2 + a3);
        tmem_st_x32(p_addr, pr);
        tmem_st_x8(p_addr + 32, pr + 32);
        tmem_st_x4(p_addr + 40, pr + 40);
        if (ready) load_s(i + 1);   // S(i+1)'s read latency runs under the publication of P(i)
        LA_CLK(t4);
        LA_ACC(3, t3, t4);
        publish_p(i);
        if (more && !ready) {
          mbar_wait(next_bar, next_par, 6, i + 1);
          load_s(i + 1);
        }
        LA_CLK(t8);
        LA_ACC(4, t4, t8);
      } else {
        LA_CLK(t5);
        m_loc = exact_tile(i, m_loc);
        publish_p(i);
        if (more) {
          mbar_wait(next_bar, next_par, 6, i + 1);
          load_s(i + 1);
        }
        LA_CLK(t6);
        LA_ACC(5, t5, t6);
      }
      stat_d = __fmul_rn(__fsub_rn(m_loc, m_true), c);
      stat_i = i;
      m_true = fmaxf(m_true, m_loc);
#ifdef LA_PROFILE_CLOCKS
      prof_acc[6] += 1;
#endif
    }
    emit_stat();   // of the last tile
#ifdef LA_PROFILE_CLOCKS
    prof_t_loop_end = clock64();
    if (lane == 0 && (warp == 0 || warp == 4)) { LA_PROF_FLUSH(warp == 0 ? 0 : 20, 7); }
#endif

    // ---------------------------------------------------------------- epilogue
    float inv = 0.f, lse = -INFINITY;
    if (T > 0) {
      lx[tid] = l_run;
      named_bar_sync(1 + (warp & 3), 64);
      const float l_tot = l_run + lx[tid ^ 128];
      const bool bad = (l_tot == 0.f) || (l_tot != l_tot);
      inv = bad ? 0.f : 1.0f / l_tot;                                              // softmax.h:283-293
      lse = bad ? -INFINITY : m_ref * args.softmax_scale + logf(l_tot);
      mbar_wait(bar(kBarOFull), 0, 8, T);
      tc_fence_after();
    }
    float o[64];
    if (T > 0) {
      const uint32_t o_addr = tmem_base + kTmemO + wg * 64 + lane_field;
      tmem_ld_x32(o_addr, reinterpret_cast<uint32_t*>(o));
      tmem_ld_x32(o_addr + 32, reinterpret_cast<uint32_t*>(o) + 32);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) o[j] = 0.f;
    }
    if (q_row < args.seqlen_q) {
      int64_t o_off = (int64_t)batch * args.o_batch_stride + (int64_t)q_row * args.o_row_stride +
                      (int64_t)head * args.o_head_stride + wg * 64;
      __nv_bfloat16* obase = args.out;
      if (args.rows_per_peer > 0) {
        // Sequence-parallel return path fused into the epilogue: this row belongs to the rank that owns its token.
        const int dest = q_row / args.rows_per_peer;
        o_off = (int64_t)batch * args.o_batch_stride + (int64_t)(q_row - dest * args.rows_per_peer) * args.o_row_stride +
                (int64_t)head * args.o_head_stride + wg * 64;
        obase = args.out_peer[dest];
      }
#ifndef LA_NO_V8
#define LA_V8_OK args.o_align32
#else
#define LA_V8_OK false
#endif
      if (args.out_f32 == nullptr || args.rows_per_peer > 0) {
        // 128 contiguous bytes per thread.  With 32-byte aligned rows every store instruction fills whole 32-byte
        // sectors (st.global.v8, STG.E.256: no half-written sectors for L2 to merge); otherwise 16-byte vectors.
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = pack_bf16(o[2 * j] * inv, o[2 * j + 1] * inv);
        if (LA_V8_OK) {
#if LA_L2_HINTS & 4
          const uint64_t pol_o = policy_evict_first();     // O is written once and not read by this kernel
#pragma unroll
          for (int j = 0; j < 4; ++j) stg_v8_hint(obase + o_off + 16 * j, w + 8 * j, pol_o);
#else
#pragma unroll
          for (int j = 0; j < 4; ++j) stg_v8(obase + o_off + 16 * j, w + 8 * j);
#endif
        } else {
          uint4* dst = reinterpret_cast<uint4*>(obase + o_off);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
      } else {
        // The consumer's dtype directly: what the reference's caller gets from `x = x.float()`
        // after the call (README.md:312-313) -- i.e. the bf16-rounded value widened, bit for bit -- without the
        // extra elementwise pass.
        float* dstf = args.out_f32 + o_off;
#pragma unroll
        for (int j = 0; j < 64; ++j) o[j] = __bfloat162float(__float2bfloat16_rn(o[j] * inv));
        if (LA_V8_OK) {
#pragma unroll
          for (int j = 0; j < 8; ++j) stg_v8(dstf + 8 * j, reinterpret_cast<const uint32_t*>(o) + 8 * j);
        } else {
          float4* dst = reinterpret_cast<float4*>(dstf);
#pragma unroll
          for (int j = 0; j < 16; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
      }
      if (wg == 0 && args.lse != nullptr)
        args.lse[((int64_t)batch * args.h + head) * args.seqlen_q + q_row] = lse;
    }
#ifdef LA_PROFILE_CLOCKS
    if (warp == 0 && lane == 0 && T > 0) {
      const long long t_end = clock64();
      atomicAdd(&g_la_prof[27], (unsigned long long)(prof_t_s0 - prof_t_entry));      // entry -> S(0) ready
      atomicAdd(&g_la_prof[28], (unsigned long long)(prof_t_first - prof_t_s0));      // first (exact) tile
      atomicAdd(&g_la_prof[29], (unsigned long long)(t_end - prof_t_loop_end));       // epilogue (O final wait + stores)
      atomicAdd(&g_la_prof[30], (unsigned long long)(t_end - prof_t_entry));          // whole CTA
      atomicAdd(&g_la_prof[31], 1ull);
    }
#endif
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
