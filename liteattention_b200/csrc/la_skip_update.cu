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
// This is integer work gated by one fp32 compare per tile, HBM-bound: one warp per (b, h, q-tile) row.
// Three paths, all producing the same bits:
//   * range-parallel (list sorted descending, no must-do ranges, every range <= 64 tiles -- what a sparse list looks
//     like): the writer's state is reset at every range start, so ranges are independent: one LANE per range
//     loads that range's statistics, packs the votes into a 64-bit mask, derives the transitions with two bit
//     operations, and a warp prefix sum places each lane's entries.  ~5x fewer instructions than the tile walk.
//   * tile-parallel (sorted; some long range -- dense or early lists -- or a must-do list): lanes map to K tiles, the
//     state machine collapses to neighbour compares + ballot prefix sums over smem bitmaps.  The must-do reader is a
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
  int pre_s = 0, pre_e = 0;
  const bool can_pre = ktiles >= 64;
  if (can_pre) {
    pre_s = __ldg(rd + 1 + 2 * lane);
    pre_e = __ldg(rd + 2 + 2 * lane);
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

  // ---- sorted / disjoint / short-range check (no smem): decides between the range-parallel and the other paths
  bool sorted_ok = !general, short_ok = true;
  for (int r0 = 0; r0 < nranges && sorted_ok; r0 += 32) {
    const int r = r0 + lane;
    bool bad = false, lng = false;
    if (r < nranges) {
      int s_ = (can_pre && r0 == 0) ? pre_s : rd[1 + 2 * r];
      int e_ = (can_pre && r0 == 0) ? pre_e : rd[2 + 2 * r];
      s_ = min(s_, ktiles - 1);
      e_ = max(e_, 0);
      if (s_ < e_) bad = true;                                  // empty after clamping: leave to the general path
      if (r > 0 && !(max(rd[2 * r], 0) > s_)) bad = true;       // previous end must be strictly above this start
      lng = (s_ - e_ + 1) > 64;
    }
    if (__any_sync(0xffffffffu, bad)) sorted_ok = false;
    if (__any_sync(0xffffffffu, lng)) short_ok = false;
  }
  if (!sorted_ok) general = true;

  if (!general && short_ok && !with_md) {
    // ---------------------------------------------------------------- range-parallel path
    for (int r0 = 0; r0 < nranges; r0 += 32) {
      const int r = r0 + lane;
      int s_ = 0, e_ = 0, nt = 0;
      if (r < nranges) {
        s_ = min((can_pre && r0 == 0) ? pre_s : rd[1 + 2 * r], ktiles - 1);
        e_ = max((can_pre && r0 == 0) ? pre_e : rd[2 + 2 * r], 0);
        nt = s_ - e_ + 1;
      }
      // votes of tiles s_, s_-1, ..., e_ (bit j = tile s_ - j): 1 = skip.  Loads of a lane's tiles are independent.
      unsigned long long votes = 0ull;
      const int nt_max = __reduce_max_sync(0xffffffffu, nt);
      for (int j0 = 0; j0 < nt_max; j0 += 8) {
        float sv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u;
          sv[u] = (j < nt && (s_ - j) != first_n) ? __ldg(stat + (s_ - j)) : INFINITY;   // +inf > thr: "do"
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u;
          if (j < nt && !(sv[u] > thr) && (s_ - j) != first_n) votes |= 1ull << j;
        }
      }
      // state before tile j: skipping (1) at the range start, else the vote of tile j-1  =>  transitions:
      const unsigned long long live = (nt >= 64) ? ~0ull : ((1ull << nt) - 1ull);
      unsigned long long trans = (votes ^ ((votes << 1) | 1ull)) & live;
      const bool end_do = nt > 0 && !((votes >> (nt - 1)) & 1ull);    // last raw vote "do": the range end is written
      const int cnt = __popcll(trans) + (end_do ? 1 : 0);
      int incl = cnt;                                                 // warp inclusive prefix sum of the counts
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      int pos = w + incl - cnt;
      while (trans) {
        const int j = __ffsll((long long)trans) - 1;
        if (pos <= ktiles) wr[pos] = s_ - j;
        ++pos;
        trans &= trans - 1ull;
      }
      if (end_do) {
        if (pos <= ktiles) wr[pos] = e_;
      }
      w += __shfl_sync(0xffffffffu, incl, 31);
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

  uint32_t* vis = upd_smem + warp * 3 * kMaskWords;
  uint32_t* smask = vis + kMaskWords;
  uint32_t* emask = smask + kMaskWords;
  const int words = (ktiles + 31) >> 5;
  for (int j = lane; j < words; j += 32) vis[j] = smask[j] = emask[j] = 0u;
  __syncwarp();
  for (int r0 = 0; r0 < nranges && !general; r0 += 32) {
    const int r = r0 + lane;
    if (r < nranges) {
      int s = (can_pre && r0 == 0) ? pre_s : rd[1 + 2 * r];
      int e = (can_pre && r0 == 0) ? pre_e : rd[2 + 2 * r];
      s = min(s, ktiles - 1);
      e = max(e, 0);
      for (int n = e; n <= s;) {                             // set bits [e, s]
        const int wi = n >> 5, lo = n & 31;
        const int hi = min(31, s - (wi << 5));
        const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        atomicOr(&vis[wi], m);
        n = (wi + 1) << 5;
      }
      atomicOr(&smask[s >> 5], 1u << (s & 31));
      atomicOr(&emask[e >> 5], 1u << (e & 31));
    }
  }
  __syncwarp();

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
