/*
 * liteattn_b200.h -- C ABI of the B200 (sm_100a) QK-Skip attention forward.
 *
 * This is the drop-in boundary for the ONE hot path of moonmath-ai/LiteAttention:
 * the skip-list-gated attention forward + skip-list update that the reference reaches through
 *     torch.ops.lite_attention.fwd            (hopper/_internal/cpp/flash_api.cpp:1722-1763 schema,
 *                                              :667-1249 mha_fwd, :915-963 skip-arg plumbing)
 * Every entry point takes plain pointers/sizes (device pointers unless stated otherwise) and a
 * cudaStream_t passed as void*.  No torch types.  All functions return 0 on success, a negative
 * LA_ERR_* code otherwise; la_last_error() returns a thread-local human-readable message.
 * Nothing here synchronises the host with the device (the reference does not either,
 * flash_api.cpp:1219-1220).
 */
#ifndef LITEATTN_B200_H_
#define LITEATTN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LA_OK 0
#define LA_ERR_INVALID (-1)     /* bad argument (shape / stride / alignment / null pointer)          */
#define LA_ERR_UNSUPPORTED (-2) /* valid for the reference but not built here (e.g. head_dim != 128) */
#define LA_ERR_CUDA (-3)        /* CUDA driver / runtime error, see la_last_error()                  */

#define LA_ABI_VERSION 3   /* 2: la_fwd_params.out_is_f32, la_rope_cast_sm100; 3: la_combine_params dtypes, la_list_pack/unpack_sm100 */

/* Tile geometry of the skip list.  API-visible: mirrors tile_size_fwd_sm90
 * (hopper/_internal/cpp/tile_size.h:10-62) == LiteAttention.get_MN (hopper/lite_attention.py:87-111).
 * bf16, d=128, non-causal -> (128, 176). */
#define LA_BLOCK_M 128
#define LA_BLOCK_N 176
#define LA_HEAD_DIM 128

/* Replaces the used fields of Flash_fwd_params / Qkv_params (hopper/_internal/cpp/flash.h:22-44,
 * :48-83) and QKSkipMaskArgs (flash.h:12-18, :181-184).  Strides are in ELEMENTS, as in the
 * reference (set_params_fprop, flash_api.cpp:81-103).  q/k/v/out are bf16 with last-dim stride 1. */
typedef struct la_fwd_params {
  const void* q;   /* (b, seqlen_q, h,   d) bf16 */
  const void* k;   /* (b, seqlen_k, h_k, d) bf16 */
  const void* v;   /* (b, seqlen_k, h_k, d) bf16 */
  void* out;       /* (b, seqlen_q, h,   d) bf16, written */
  float* lse;      /* (b, h, seqlen_q) fp32 contiguous, written; may be NULL */
  int64_t q_batch_stride, q_row_stride, q_head_stride;
  int64_t k_batch_stride, k_row_stride, k_head_stride;
  int64_t v_batch_stride, v_row_stride, v_head_stride;
  int64_t o_batch_stride, o_row_stride, o_head_stride;
  int32_t b, h, h_k, seqlen_q, seqlen_k, d;
  float softmax_scale;
  /* Skip list that gates whole (Q-tile, K-tile) iterations.  int32, contiguous,
   * [>=b, h, qtiles, ktiles+1], row = [len, s0, e0, s1, e1, ...] with inclusive descending ranges
   * (SkipListReader, mainloop_fwd_sm90_tma_gmma_ws.hpp:47-115).  NULL => dense (every tile). */
  const int32_t* read_list;
  /* Per-tile QK-skip statistic, fp32 contiguous [b, h, qtiles, ktiles]; for every VISITED tile n
   * (except the first one of a row, which gets +inf) the kernel writes
   *     max over the 128 rows of ((m_local - m_prev) * softmax_scale * log2(e))
   * i.e. the left-hand side of the skip predicate (softmax.h:194).  May be NULL. Unvisited entries
   * are left untouched. */
  float* tile_stat;
  /* 0: out is bf16 (the reference's op).  1: out is fp32 with the same element strides and holds the bf16-rounded
   * result widened -- exactly what the reference's caller computes with `x = x.float()` after the call
   * (README.md:312-313), without the extra pass. */
  int32_t out_is_f32;
  /* Sequence-parallel scatter of O (out_rows_per_peer > 0, bf16 only): query row r is written to
   * out_peer[r / out_rows_per_peer] at row (r % out_rows_per_peer), with the o_*_stride element strides -- the peers'
   * buffers mapped into this process over NVLink (CUDA IPC / symmetric memory); `out` is ignored.  n_out_peers <= 8
   * and n_out_peers * out_rows_per_peer >= seqlen_q.  (the return all-to-all of a head-parallel
   * single-prompt run happens inside the forward's epilogue.) */
  int32_t n_out_peers;
  int32_t out_rows_per_peer;
  int32_t reserved_;
  void* out_peer[8];
} la_fwd_params;

/* Replaces SkipListWriter + the record/loop logic that the reference fuses into the forward
 * (mainloop_fwd_sm90_tma_gmma_ws.hpp:121-192, :1804-1827).  Here it is a separate HBM-bound kernel. */
typedef struct la_update_params {
  const int32_t* read_list;    /* [>=b, h, qtiles, ktiles+1] the list the forward just used  */
  const int32_t* must_do_list; /* same shape (expanded, lite_attention.py:214-242); may be NULL */
  int32_t* write_list;         /* same shape; rows are rewritten: [len, entries...]            */
  const float* tile_stat;      /* [b, h, qtiles, ktiles] from la_fwd_sm100                      */
  int32_t b, h, qtiles, ktiles;
  float thr;                   /* vote skip iff !(stat > thr)  (softmax.h:194,207)              */
  int32_t* overflow_count;     /* optional device counter: rows whose new list would not fit   */
} la_update_params;

/* O_i/LSE_i partial-attention merge.  Replaces the reference op lite_attention::fwd_combine = mha_combine
 * (hopper/_internal/cpp/flash_api.cpp:1620-1720, kernel flash_fwd_combine_kernel.h -- compiled out of the shipped
 * build, hopper/setup.py:48; README.md:222-250 asks callers to merge partial results by LSE themselves). */
typedef struct la_combine_params {
  const void* const* o_parts;    /* HOST array of n_parts device pointers, each (b, s, h, d) contiguous, bf16 or fp32 */
  const float* const* lse_parts; /* HOST array of n_parts device pointers, each (b, h, s) fp32 contiguous    */
  int32_t n_parts;
  void* out;                     /* (b, s, h, d) contiguous, bf16 or fp32 */
  float* lse;                    /* (b, h, s) fp32, may be NULL  */
  int32_t b, h, s, d;
  int32_t parts_are_f32;         /* 0: bf16 partials (what la_fwd_sm100 writes), 1: fp32 (the reference op's contract) */
  int32_t out_is_f32;            /* 0: bf16 result, 1: fp32 */
} la_combine_params;

/* Caller-side step in front of the attention call, fused: 3-D rotary embedding of Q or K + cast to bf16.
 * Replaces `rope_apply(x, grid_sizes, freqs)` followed by `.bfloat16()` in the reference's Wan integration
 * (README.md:301-315; rope_apply is Wan2.1's wan/modules/model.py, restated in oracle/rope.py).   */
typedef struct la_rope_params {
  const void* x;           /* (b, s, h, d) fp32 (x_is_bf16 = 0) or bf16 (1), last-dim stride 1, strides in elements */
  void* out;               /* (b, s, h, d) bf16 contiguous, written */
  const float* cos_sin;    /* [max_pos, d/2, 2] fp32 (cos, sin): angle of position p for complex pair c; the three
                            * axes' tables concatenated along d/2 in the order frames | height | width with widths
                            * d/2 - 2*(d/2/3), d/2/3, d/2/3 (how Wan builds `freqs`) */
  const int32_t* grid;     /* [b, 3] device int32: (frames, height, width) of each sample; tokens >= f*h*w are cast only */
  int64_t x_batch_stride, x_row_stride, x_head_stride;
  int32_t b, s, h, d, max_pos;
  int32_t x_is_bf16;
} la_rope_params;

int la_abi_version(void);
const char* la_last_error(void);

/* (kBlockM, kBlockN) for a head dim / element size.  Same table as get_MN. Returns LA_ERR_UNSUPPORTED
 * for geometries whose kernel is not built. */
int la_get_tile_mn(int head_dim, int element_size, int v_colmajor, int* block_m, int* block_n);

/* Skip-list-gated attention forward (tcgen05/TMEM/TMA kernel). */
int la_fwd_sm100(const la_fwd_params* p, void* stream);

/* Skip-list update from the per-tile statistic. */
int la_skip_update_sm100(const la_update_params* p, void* stream);

/* Convenience: la_fwd_sm100 followed by la_skip_update_sm100 on the same stream -- the exact
 * behaviour of one lite_attention::fwd call with (attn_read_list, attn_must_do_list,
 * attn_write_list, thr).  upd->read_list / tile_stat are taken from fwd when NULL. */
int la_fwd_skip_sm100(const la_fwd_params* fwd, const la_update_params* upd, void* stream);

/* Merge n partial attention results by their LSE. */
int la_combine_sm100(const la_combine_params* p, void* stream);

/* Fused 3-D RoPE + bf16 cast (HBM-bound elementwise kernel). */
int la_rope_cast_sm100(const la_rope_params* p, void* stream);

/* Compact resident skip state (the reference keeps int32 [2, max_batch, H, qtiles, ktiles+1] per layer
 * object, hopper/lite_attention.py:113-153): two bits per (row, K tile) -- "listed" and "a range starts here" --
 * uint32 bits[rows][2][ceil(ktiles/32)].  pack: int32 list rows -> bits (rows whose ranges are not descending and
 * disjoint cannot be represented: they are zeroed and counted in *bad_rows, which may be NULL); unpack: bits -> rows
 * [len, s0, e0, ...] (entries past len untouched).  unpack(pack(row)) == row on [0, len] for every row the update
 * kernel writes. */
int la_list_pack_sm100(const int32_t* list, uint32_t* bits, int64_t rows, int ktiles, int32_t* bad_rows, void* stream);
int la_list_unpack_sm100(const uint32_t* bits, int32_t* list, int64_t rows, int ktiles, void* stream);

/* Host-resident activations (offloaded pipelines; what the reference's users do with `.cuda()` / `.cpu()` around the
 * call, hopper README "Quick start"): an asynchronous pitched copy of `rows` rows of `width_bytes` between pinned host
 * memory and device memory (direction inferred from the pointers), on `stream`.  It lets the host-side mirror move one
 * HEAD GROUP of a (batch, seq, heads, dim) tensor at a time, so that the forward of group g runs while group g+1 is
 * still on the wire (liteattention_b200/lite_attention.py: LiteAttention.__call__ on pinned CPU tensors). */
int la_copy2d_async(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes,
                    size_t rows, void* stream);

/* Number of kernels launched through this library by the calling process (for bench accounting). */
uint64_t la_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LITEATTN_B200_H_ */
