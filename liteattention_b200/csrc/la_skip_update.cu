// la_skip_update.cu -- QK-Skip mask update for B200 (sm_100a): turns the per-tile statistic emitted by
// la_fwd_kernel into the next run-length skip list.
//
// Restates the behaviour of SkipListWriter + the range loop of the reference's fused kernel
// (hopper/_internal/cpp/mainloop_fwd_sm90_tma_gmma_ws.hpp:121-192 writer, :1804-1827 loop,
//  softmax.h:194/207 predicate `any_row((m_loc - m_prev) * scale_log2 > thr)` => "do"):
//
//   w = 1; skipping = true; first visited tile: vote = do, never tested, no must-do lookup.
//   every later visited tile n: vote_skip = !(stat[n] > thr); if vote_skip and n lies inside the current
//   must-do range (n <= start && n > end, reader advanced by a single `if`) the vote becomes "do";
//   a change of state writes n; at the end of each READ range the state is forced back to "skipping" and the
//   inclusive end is written iff the last RAW vote (before the must-do override) was "do"; row[0] = w - 1.
//
// This is integer work gated by one fp32 compare per tile: one warp per (b, h, q-tile) row.
// Three paths, all producing the same bits:
//   * bitmap path (list sorted descending and disjoint, no must-do list -- every list this kernel itself writes; round 2):
//     the row becomes three bitmaps (range starts, range ends, skip votes), one 32-tile word per lane; the writer's state
//     machine is then a handful of word operations (see the comment at the path) and a warp prefix sum places the
//     entries.  The statistic is read with coalesced 128-byte loads.  Replaces round 1's range-parallel and
//     tile-parallel paths for these rows (42 % Wan list: 42 -> 34 us, dense: 49 -> 37 us).
//   * tile-parallel with a must-do list (sorted read row, <= 32 sorted zero-padded must-do ranges): lanes map to K tiles,
//     the state machine collapses to neighbour compares + ballot prefix sums over smem bitmaps.  The must-do reader is a
//     "lagging follower": on every skip-voted tile it moves at most one range towards the range that contains the
//     tile, m_k = min(m_{k-1} + 1, R_k)  =>  m_k = k + min_{j <= k}(R_j - j): a warp prefix-min over the skip-voted
//     tiles in visit order (round 2; it used to send the whole row to one lane).
//   * general (a hand-made unsorted list, or a must-do row with more than 32 or unsorted ranges): lane 0 walks the row.
//
// Bounds: the reference writer has no capacity check and can overflow a row into its neighbour
// (SURVEY.md section 8 a12-iii).  Here a row that would need more than `ktiles` entries is replaced by a
// copy of the read row (always valid, merely not sparser) and counted in *overflow_count.
#include "la_kernels.h"
#include "la_ptx.cuh"

namespace la {

namespace {
constexpr int kMaskWords = kFwdMaxTiles / 32;

__device__ __forceinline__ int clamp_len(int len, int ktiles) { return min(max(len, 0), ktiles) & ~1; }
}  // namespace

size_t la_skip_update_smem_bytes(int) { return (size_t)kUpdWarpsPerBlock * 3 * kMaskWords * sizeof(uint32_t); }

__global__ void __launch_bounds__(kUpdWarpsPerBlock * 32) la_skip_update_kernel(const UpdateKernelArgs args) {
  extern __shared__ uint32_t upd_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row_id = blockIdx.x * kUpdWarpsPerBlock + warp;
  if (row_id >= args.rows) return;

  const int ktiles = args.ktiles;
  const int64_t stride = (int64_t)ktiles + 1;
  const int32_t* rd = args.read_list + row_id * stride;
  const int32_t* md = args.must_do_list ? args.must_do_list + row_id * stride : nullptr;
  int32_t* wr = args.write_list + row_id * stride;
  const float* stat = args.tile_stat + (int64_t)row_id * ktiles;
  const float thr = args.thr;

  if (ktiles == 1) {
    // One-tile rows are [len, 0] (no room for a range end): the only tile is always kept.
    if (lane == 0) {
      if (rd[0] > 0) {
        wr[0] = 2;
        wr[1] = 0;
      } else {
        wr[0] = 0;
      }
    }
    return;
  }
  // The row is walked through three dependent HBM round trips (length -> ranges -> statistic).  Fold the first two:
  // the first 32 ranges are requested together with the length (a row always has ktiles + 1 >= 65 words here).
  int pre_s = 0, pre_e = 0, pre2_s = 0, pre2_e = 0;
  const bool can_pre = ktiles >= 64, can_pre2 = ktiles >= 128;
  if (can_pre) {
    pre_s = __ldg(rd + 1 + 2 * lane);
    pre_e = __ldg(rd + 2 + 2 * lane);
  }
  if (can_pre2) {                      // ranges 32..63 as well: a 42 %-sparse Wan row has ~62 of them
    pre2_s = __ldg(rd + 65 + 2 * lane);
    pre2_e = __ldg(rd + 66 + 2 * lane);
  }
  const int len = clamp_len(__ldg(rd), ktiles);
  const int nranges = len >> 1;
  if (nranges == 0) {
    if (lane == 0) wr[0] = 0;
    return;
  }

  // ---- classify: is the must-do row trivial, is the read row sorted/disjoint?
  bool general = false;
  bool with_md = false;            // a non-trivial must-do row that the tile-parallel path can handle
  int md_s = 0, md_e = 0, md_n = 0;   // lane j holds must-do range j (start, end); ranges >= md_n are the (0, 0) padding
  if (md != nullptr) {
    const int mdlen = md[0];
    // `[2, 0, 0]` (lite_attention.py:229-231 default) protects nothing: n <= 0 && n > 0 is never true.
    if (!(mdlen <= 0 || (mdlen == 2 && md[1] == 0 && md[2] == 0))) {
      md_n = min(mdlen, ktiles) >> 1;
      // What the serial reader sees past the listed ranges is whatever follows in the row: the expanded lists are
      // zero-padded (lite_attention.py:236-238); the parallel path relies on that and on descending, disjoint ranges.
      bool ok = (mdlen > 0) && (mdlen % 2 == 0) && md_n <= 32 && (2 * md_n + 2 <= ktiles);
      if (ok) {
        if (lane < md_n) {
          md_s = md[1 + 2 * lane];
          md_e = md[2 + 2 * lane];
        }
        const int nxt_s = __shfl_down_sync(0xffffffffu, md_s, 1);
        bool bad = false;
        if (lane < md_n) {
          if (md_s < md_e || md_e < 0) bad = true;
          if (lane + 1 < md_n && !(md_e > nxt_s)) bad = true;        // strictly descending, disjoint
        }
        if (lane == 0 && (md[1 + 2 * md_n] != 0 || md[2 + 2 * md_n] != 0)) bad = true;   // padding must be (0, 0)
        ok = !__any_sync(0xffffffffu, bad);
      }
      if (ok) with_md = true;
      else general = true;
    }
  }
  int first_n = min(rd[1], ktiles - 1);
  int w = 1;  // next write slot
  bool overflow = false;

  // ---- one pass over the ranges: sorted / disjoint check, and the row as bitmaps in smem (bit n <-> K tile n):
  // smask = a range starts at n, emask = a range ends at n (and, for the must-do walk only, vis = n is visited).
  uint32_t* vis = upd_smem + warp * 3 * kMaskWords;
  uint32_t* smask = vis + kMaskWords;
  uint32_t* emask = smask + kMaskWords;
  const int words = (ktiles + 31) >> 5;
  for (int j = lane; j < words; j += 32) vis[j] = smask[j] = emask[j] = 0u;
  __syncwarp();
  int prev_e_carry = 0x7fffffff;   // end of the last range of the previous trip
  for (int r0 = 0; r0 < nranges && !general; r0 += 32) {
    const int r = r0 + lane;
    int s_ = 0, e_ = 0;
    bool bad = false;
    if (r < nranges) {
      s_ = min((can_pre && r0 == 0) ? pre_s : (can_pre2 && r0 == 32) ? pre2_s : rd[1 + 2 * r], ktiles - 1);
      e_ = max((can_pre && r0 == 0) ? pre_e : (can_pre2 && r0 == 32) ? pre2_e : rd[2 + 2 * r], 0);
      if (s_ < e_) bad = true;                                  // empty after clamping: leave to the general path
    }
    int prev_e = __shfl_up_sync(0xffffffffu, e_, 1);
    if (lane == 0) prev_e = prev_e_carry;
    if (r < nranges && !(prev_e > s_)) bad = true;              // previous end must be strictly above this start
    if (__any_sync(0xffffffffu, bad)) {
      general = true;
      break;
    }
    if (r < nranges) {
      atomicOr(&smask[s_ >> 5], 1u << (s_ & 31));
      atomicOr(&emask[e_ >> 5], 1u << (e_ & 31));
      if (with_md) {
        for (int n = e_; n <= s_;) {                             // set bits [e, s]
          const int wi = n >> 5, lo = n & 31;
          const int hi = min(31, s_ - (wi << 5));
          const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
          atomicOr(&vis[wi], m);
          n = (wi + 1) << 5;
        }
      }
    }
    prev_e_carry = __shfl_sync(0xffffffffu, e_, 31);
  }
  __syncwarp();

  if (!general && !with_md) {
    // ---------------------------------------------------------------- bitmap path (round 2): lane = one 32-tile word
    // Descending visit order = descending bit order.  With V = skip votes of the visited tiles (the first visited tile
    // votes "do"), the writer's state before tile n is 1 ("skipping") at a range start and V[n+1] elsewhere, so
    //   TRANS = VIS & (V ^ (ST | (~ST & V>>1)))   -> n is written (state change)
    //   ENDDO = EN & ~V                           -> the range end is written again (last raw vote was "do")
    // and VIS itself is the running XOR (from the top bit down) of the toggles ST ^ (EN >> 1).  Entries are placed by a
    // warp prefix sum over the words; the statistic is read with coalesced 128-byte loads, one word per instruction.
    uint32_t carry_par = 0u, carry_en = 0u, carry_v = 0u;
    for (int hi = words - 1; hi >= 0; hi -= 32) {
      const int wi = hi - lane;                       // this lane's word (lane 0 = the highest tiles of the trip)
      const bool act = wi >= 0;
      const int nw = min(32, hi + 1);
      const uint32_t st = act ? smask[wi] : 0u, en = act ? emask[wi] : 0u;
      uint32_t en_up = __shfl_up_sync(0xffffffffu, en, 1);
      if (lane == 0) en_up = carry_en;
      const uint32_t tg = st ^ ((en >> 1) | (en_up << 31));
      uint32_t y = tg;
      y ^= y >> 1;
      y ^= y >> 2;
      y ^= y >> 4;
      y ^= y >> 8;
      y ^= y >> 16;
      const uint32_t pb = __ballot_sync(0xffffffffu, (__popc(tg) & 1) != 0);
      const uint32_t par = ((uint32_t)__popc(pb & ((1u << lane) - 1u)) & 1u) ^ carry_par;
      const uint32_t visw = par ? ~y : y;
      uint32_t vw = 0u;
      for (int u0 = 0; u0 < nw; u0 += 4) {
        float sv[4];
        bool ld[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int u = u0 + k;
          const uint32_t vc = __shfl_sync(0xffffffffu, visw, u & 31);
          const int n = ((hi - u) << 5) + lane;
          ld[k] = (u < nw) && ((vc >> lane) & 1u) && (n != first_n);
          sv[k] = ld[k] ? __ldg(stat + n) : INFINITY;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t bal = __ballot_sync(0xffffffffu, ld[k] && !(sv[k] > thr));
          if (lane == u0 + k) vw = bal;
        }
      }
      uint32_t v_up = __shfl_up_sync(0xffffffffu, vw, 1);
      if (lane == 0) v_up = carry_v;
      const uint32_t prev = st | (~st & ((vw >> 1) | (v_up << 31)));
      const uint32_t trans = visw & (vw ^ prev);
      const uint32_t enddo = en & ~vw;
      const int cnt = __popc(trans) + __popc(enddo);
      int incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      int pos = w + incl - cnt;
      uint32_t m = trans | enddo;
      while (m) {
        const int b = 31 - __clz(m);
        const int n = (wi << 5) + b;
        if ((trans >> b) & 1u) {
          if (pos <= ktiles) wr[pos] = n;
          ++pos;
        }
        if ((enddo >> b) & 1u) {
          if (pos <= ktiles) wr[pos] = n;
          ++pos;
        }
        m &= ~(1u << b);
      }
      w += __shfl_sync(0xffffffffu, incl, 31);
      carry_par ^= (uint32_t)__popc(pb) & 1u;
      carry_en = __shfl_sync(0xffffffffu, en, 31);
      carry_v = __shfl_sync(0xffffffffu, vw, 31);
    }
    overflow = (w - 1) > ktiles;
    if (overflow) {
      for (int j = lane; j <= len; j += 32) wr[j] = (j == 0) ? len : rd[j];
      if (lane == 0 && args.overflow_count != nullptr) atomicAdd(args.overflow_count, 1);
    } else if (lane == 0) {
      wr[0] = w - 1;
    }
    return;
  }

  if (!general) {
    // ---------------------------------------------------------------- tile-parallel path
    bool carry_ev = true;  // effective vote of the previous (higher) tile; irrelevant at range starts
    int md_k = 0;          // skip-voted tiles seen so far (visit order = descending tile index)
    int md_min = 0;        // min(m_0, min_j (R_j - j)) over them, m_0 = 0: the follower's position is k + md_min
    // Four 32-tile chunks per trip: their statistic loads are issued together (the walk itself is a serial
    // ballot/prefix chain, so without this every chunk would expose one full HBM latency).
    constexpr int kUnroll = 4;
    for (int base = ktiles - 1; base >= 0; base -= 32 * kUnroll) {
      float sv[kUnroll];
      bool vv[kUnroll], stt[kUnroll], enn[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int n = base - 32 * u - lane;
        bool v = false, st = false, en = false;
        if (n >= 0) {
          const uint32_t bit = 1u << (n & 31);
          v = vis[n >> 5] & bit;
          st = smask[n >> 5] & bit;
          en = emask[n >> 5] & bit;
        }
        vv[u] = v;
        stt[u] = st;
        enn[u] = en;
        sv[u] = (v && n != first_n) ? __ldg(stat + n) : INFINITY;   // +inf > thr: "do", like the untested first tile
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int n = base - 32 * u - lane;
        const bool v = vv[u], st = stt[u], en = enn[u];
        bool rv = false;  // raw vote: true = skip
        if (v && n != first_n) rv = !(sv[u] > thr);
        bool ev = rv;
        if (with_md) {
          // R = first must-do range whose end is <= n (ranges are descending; index md_n = the (0, 0) padding).
          // The serial reader tests `end > n` on every skip-voted tile and then moves ONE range (writer :156-159).
          int R = 0;
          for (int j = 0; j < md_n; ++j) R += (__shfl_sync(0xffffffffu, md_e, j) > n) ? 1 : 0;
          const uint32_t mv = __ballot_sync(0xffffffffu, rv);
          const int k = md_k + __popc(mv & ((1u << lane) - 1u)) + 1;     // 1-based rank of this tile among skip votes
          int cand = rv ? (R - k) : 0x3fffffff;                           // R_k - k, only skip-voted tiles take part
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {                              // inclusive prefix min in lane (= visit) order
            const int o = __shfl_up_sync(0xffffffffu, cand, d);
            if (lane >= d) cand = min(cand, o);
          }
          const int run_min = min(md_min, cand);
          const int m = rv ? min(k + run_min, md_n) : 0;                  // range the reader points at AFTER its single step
          const int ms = __shfl_sync(0xffffffffu, md_s, m & 31), me = __shfl_sync(0xffffffffu, md_e, m & 31);
          if (rv && m < md_n && n <= ms && n > me) ev = false;            // protected: the vote becomes "do"
          md_min = min(md_min, __shfl_sync(0xffffffffu, cand, 31));
          md_k += __popc(mv);
        }
        bool prev_ev = __shfl_up_sync(0xffffffffu, ev, 1);
        if (lane == 0) prev_ev = carry_ev;
        const bool ps = st ? true : prev_ev;
        const bool a = v && (ev != ps);
        const bool b = v && en && !rv;
        const uint32_t ma = __ballot_sync(0xffffffffu, a);
        const uint32_t mb = __ballot_sync(0xffffffffu, b);
        const uint32_t lt = (1u << lane) - 1u;
        const int pos = w + __popc(ma & lt) + __popc(mb & lt);
        if (a && pos <= ktiles) wr[pos] = n;
        if (b && pos + (a ? 1 : 0) <= ktiles) wr[pos + (a ? 1 : 0)] = n;
        w += __popc(ma) + __popc(mb);
        carry_ev = __shfl_sync(0xffffffffu, ev, 31);
      }
    }
    overflow = (w - 1) > ktiles;
  } else {
    // ---------------------------------------------------------------- general path (lane 0)
    if (lane == 0) {
      bool skipping = true, raw = false, first = true;
      int mdlen = 2, mi = 1, ms = 0, me = 0;
      auto md_at = [&](int idx) { return (md != nullptr && idx <= ktiles) ? md[idx] : 0; };
      if (md != nullptr) {
        mdlen = md[0];
        ms = md_at(1);
        me = md_at(2);
      }
      auto put = [&](int val) {
        if (w <= ktiles) wr[w] = val;
        else overflow = true;
        ++w;
      };
      for (int r = 0; r < nranges; ++r) {
        int s = min(rd[1 + 2 * r], ktiles - 1);
        const int e = max(rd[2 + 2 * r], 0);
        if (s < e) continue;
        for (int n = s; n >= e; --n) {
          bool vote;
          if (first) {
            vote = false;
            raw = false;
            first = false;
          } else {
            raw = !(stat[n] > thr);
            vote = raw;
            if (vote) {
              if (me > n && mi <= mdlen) {  // single `if`, not `while` (writer :156-159)
                mi += 2;
                ms = md_at(mi);
                me = md_at(mi + 1);
              }
              if (n <= ms && n > me) vote = false;
            }
          }
          if (vote != skipping) {
            put(n);
            skipping = vote;
          }
        }
        skipping = true;            // record_range_end gets the RAW vote (Appendix A quirk)
        if (!raw) put(e);
      }
    }
    w = __shfl_sync(0xffffffffu, w, 0);
    overflow = __shfl_sync(0xffffffffu, (int)overflow, 0);
  }

  if (overflow) {
    for (int j = lane; j <= len; j += 32) wr[j] = (j == 0) ? len : rd[j];
    if (lane == 0 && args.overflow_count != nullptr) atomicAdd(args.overflow_count, 1);
  } else if (lane == 0) {
    wr[0] = w - 1;
  }
}

}  // namespace la
