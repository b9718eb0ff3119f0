// la_list_codec.cu -- compact resident form of the skip state.
//
// The reference keeps, per layer object, an int32 double buffer [2, max_batch, H, qtiles, ktiles + 1] of run-length rows
// (hopper/lite_attention.py:113-153): 326 MB at the Wan2.1-14B shape with the default max_batch_size = 4, times 40
// layers.  A row is a set of visited K tiles cut into descending ranges, so two bits per tile hold it exactly:
//     vis[n]   tile n is listed            start[n]   a range starts at tile n (its highest tile)
// (range boundaries matter -- the writer's state is reset at every range start -- so the tile set alone is not enough).
// 2 x ceil(ktiles / 32) words per row = 2.6 MB per layer at that shape.  The int32 rows the kernels consume are expanded
// into a scratch list shared by all layers right before the forward and packed again right after the update; both
// directions are one warp per row, HBM-bound, a few microseconds.
// Lossless for rows whose ranges are descending and disjoint (everything the update kernel writes); other rows are
// reported through *bad_rows and must stay in list form.
#include "la_kernels.h"
#include "la_ptx.cuh"

namespace la {

// list row [len, s0, e0, ...] -> bits[row][0][w] = vis, bits[row][1][w] = start
__global__ void __launch_bounds__(128) la_list_pack_kernel(const int32_t* __restrict__ list, uint32_t* __restrict__ bits,
                                                          int rows, int ktiles, int32_t* bad_rows) {
  __shared__ uint32_t sm[4][2][kFwdMaxTiles / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const int words = (ktiles + 31) >> 5;
  const int32_t* rd = list + (int64_t)row * (ktiles + 1);
  uint32_t* vis = sm[warp][0];
  uint32_t* st = sm[warp][1];
  for (int j = lane; j < words; j += 32) vis[j] = st[j] = 0u;
  __syncwarp();
  const int len = min(max(__ldg(rd), 0), ktiles) & ~1;
  bool bad = false;
  for (int r0 = 0; r0 < (len >> 1); r0 += 32) {
    const int r = r0 + lane;
    if (r < (len >> 1)) {
      const int s = __ldg(rd + 1 + 2 * r), e = __ldg(rd + 2 + 2 * r);
      if (s >= ktiles || e < 0 || s < e) bad = true;
      else if (r > 0 && !(__ldg(rd + 2 * r) > s)) bad = true;      // previous end must be strictly above this start
      else {
        for (int n = e; n <= s;) {                                  // set bits [e, s]
          const int wi = n >> 5, lo = n & 31;
          const int hi = min(31, s - (wi << 5));
          const uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
          atomicOr(&vis[wi], m);
          n = (wi + 1) << 5;
        }
        atomicOr(&st[s >> 5], 1u << (s & 31));
      }
    }
  }
  if (ktiles == 1 && __ldg(rd) > 0) {       // one-tile rows are [len, 0]: tile 0 listed
    if (lane == 0) { vis[0] = 1u; st[0] = 1u; }
    bad = false;
  }
  bad = __any_sync(0xffffffffu, bad);
  __syncwarp();
  uint32_t* out = bits + (int64_t)row * 2 * words;
  for (int j = lane; j < words; j += 32) {
    out[j] = bad ? 0u : vis[j];
    out[words + j] = bad ? 0u : st[j];
  }
  if (bad && lane == 0 && bad_rows != nullptr) atomicAdd(bad_rows, 1);
}

// bits -> list row; entries beyond the length are left untouched (stale by design, SkipListWriter :184-191)
__global__ void __launch_bounds__(128) la_list_unpack_kernel(const uint32_t* __restrict__ bits, int32_t* __restrict__ list,
                                                            int rows, int ktiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= rows) return;
  const int words = (ktiles + 31) >> 5;
  const uint32_t* vis = bits + (int64_t)row * 2 * words;
  const uint32_t* st = vis + words;
  int32_t* wr = list + (int64_t)row * (ktiles + 1);
  if (ktiles == 1) {
    if (lane == 0) {
      const bool on = __ldg(vis) & 1u;
      wr[0] = on ? 2 : 0;
      if (on) wr[1] = 0;
    }
    return;
  }
  int w = 1;
  for (int base = ktiles - 1; base >= 0; base -= 32) {        // lanes walk the tiles downwards: lane 0 = highest tile
    const int n = base - lane;
    bool v = false, s = false, below_v = false, below_s = false;
    if (n >= 0) {
      v = (__ldg(vis + (n >> 5)) >> (n & 31)) & 1u;
      s = (__ldg(st + (n >> 5)) >> (n & 31)) & 1u;
      if (n > 0) {
        below_v = (__ldg(vis + ((n - 1) >> 5)) >> ((n - 1) & 31)) & 1u;
        below_s = (__ldg(st + ((n - 1) >> 5)) >> ((n - 1) & 31)) & 1u;
      }
    }
    const bool is_start = v && s;
    const bool is_end = v && (n == 0 || !below_v || below_s);
    const uint32_t ma = __ballot_sync(0xffffffffu, is_start), mb = __ballot_sync(0xffffffffu, is_end);
    const uint32_t lt = (1u << lane) - 1u;
    const int pos = w + __popc(ma & lt) + __popc(mb & lt);
    if (is_start && pos <= ktiles) wr[pos] = n;
    if (is_end && pos + (is_start ? 1 : 0) <= ktiles) wr[pos + (is_start ? 1 : 0)] = n;
    w += __popc(ma) + __popc(mb);
  }
  if (lane == 0) wr[0] = min(w - 1, ktiles) & ~1;
}

}  // namespace la
