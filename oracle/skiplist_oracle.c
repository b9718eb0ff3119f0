/* skiplist_oracle.c -- plain-C restatement of the reference's skip-list codec.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates SkipListReader (hopper/_internal/cpp/mainloop_fwd_sm90_tma_gmma_ws.hpp:47-115), SkipListWriter
 * (:121-192) and the skippable range loop (:1804-1827), with the vote taken from a per-tile statistic:
 *     vote_skip(n) = !(stat[n] > thr)            (softmax.h:194,207; NaN compares false => skip)
 * Same semantics as oracle/skiplist.py:skip_list_step (used for big lists where Python is too slow).
 * Rows: int32 [ktiles+1] = [len, s0, e0, ...], inclusive descending ranges.
 *
 * on_overflow: 0 = unbounded (caller gives out rows of out_stride >= 2*(ktiles+1) ints),
 *              1 = the CUDA kernel's policy: a row needing more than ktiles entries becomes a copy of the read row.
 * Returns the number of rows that overflowed.
 */
#include <stdint.h>

static int md_at(const int32_t* md, int idx, int ktiles) { return (md && idx >= 0 && idx <= ktiles) ? md[idx] : 0; }

int skiplist_oracle_step(const int32_t* read, const int32_t* mustdo, int32_t* write, const float* stat, int rows,
                         int ktiles, float thr, int out_stride, int on_overflow) {
  int overflowed = 0;
  for (int r = 0; r < rows; ++r) {
    const int32_t* rd = read + (int64_t)r * (ktiles + 1);
    const int32_t* md = mustdo ? mustdo + (int64_t)r * (ktiles + 1) : 0;
    int32_t* wr = write + (int64_t)r * out_stride;
    const float* st = stat + (int64_t)r * ktiles;
    if (ktiles == 1) { /* one-tile rows are [len, 0] */
      if (rd[0] > 0) { wr[0] = 2; wr[1] = 0; } else { wr[0] = 0; }
      continue;
    }
    int len = rd[0];
    if (len < 0) len = 0;
    if (len > ktiles) len = ktiles;
    len &= ~1;
    int mdlen = md ? md[0] : 2, mi = 1, ms = md_at(md, 1, ktiles), me = md_at(md, 2, ktiles);
    int w = 1, skipping = 1, raw = 0, first = 1;
    const int cap = on_overflow ? ktiles : out_stride - 1;
    for (int i = 0; i < len; i += 2) {
      int s = rd[1 + i], e = rd[2 + i];
      if (s > ktiles - 1) s = ktiles - 1;
      if (e < 0) e = 0;
      if (s < e) continue;
      for (int n = s; n >= e; --n) {
        int vote;
        if (first) {
          vote = raw = 0;
          first = 0;
        } else {
          raw = !(st[n] > thr);
          vote = raw;
          if (vote) {
            if (me > n && mi <= mdlen) { /* single `if`, writer :156-159 */
              mi += 2;
              ms = md_at(md, mi, ktiles);
              me = md_at(md, mi + 1, ktiles);
            }
            if (n <= ms && n > me) vote = 0;
          }
        }
        if (vote != skipping) {
          if (w <= cap) wr[w] = n;
          ++w;
          skipping = vote;
        }
      }
      skipping = 1; /* record_range_end: raw vote, :172-181 + Appendix A quirk */
      if (!raw) {
        if (w <= cap) wr[w] = e;
        ++w;
      }
    }
    if (w - 1 > cap) {
      ++overflowed;
      if (on_overflow) {
        for (int j = 1; j <= len; ++j) wr[j] = rd[j];
        wr[0] = len;
      } else {
        wr[0] = w - 1;
      }
    } else {
      wr[0] = w - 1;
    }
  }
  return overflowed;
}
