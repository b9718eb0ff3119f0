// la_api.cu -- host side of the C ABI declared in include/liteattn_b200.h.
// Plays the role of mha_fwd / run_mha_fwd / run_flash_fwd in the reference
// (hopper/_internal/cpp/flash_api.cpp:667-1249, :250-380; flash_fwd_launch_template.h:52-363):
// validate, build TMA descriptors, launch on the caller's stream, never synchronise.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/liteattn_b200.h"
#include "la_kernels.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define LA_CHECK_ARG(cond, ...) \
  do {                          \
    if (!(cond)) return fail(LA_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define LA_CUDA(expr)                                                                            \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) return fail(LA_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// (b, s, h, d) bf16 tensor with element strides -> 4-D tensor map {d, s, h, b}, box {64, box_rows, 1, 1},
// 128-byte swizzle, out-of-bounds rows filled with zeros (what the reference's TMA loads do for the ragged
// last Q / K tile).
int make_tmap(CUtensorMap* tm, const void* ptr, int b, int s, int h, int d, int64_t bs, int64_t rs, int64_t hs,
              int box_rows, const char* name) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(LA_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)s, (cuuint64_t)h, (cuuint64_t)b};
  cuuint64_t strides[3] = {(cuuint64_t)rs * 2, (cuuint64_t)hs * 2, (cuuint64_t)bs * 2};
  // Size-1 dimensions may carry arbitrary strides in torch; give TMA something legal.
  if (h == 1) strides[1] = (cuuint64_t)d * 2;
  if (b == 1) strides[2] = (cuuint64_t)d * 2;
  if (s == 1) strides[0] = (cuuint64_t)d * 2;
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LA_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", name, (int)r);
  return LA_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

int la_abi_version(void) { return LA_ABI_VERSION; }
const char* la_last_error(void) { return g_err; }
uint64_t la_launch_count(void) { return g_launches.load(); }

int la_get_tile_mn(int head_dim, int element_size, int v_colmajor, int* block_m, int* block_n) {
  // Same table as LiteAttention.get_MN (hopper/lite_attention.py:87-111) for non-causal, non-local attention.
  int m, n;
  if (element_size == 2) {
    if (head_dim <= 64) { m = 192; n = 192; }
    else if (head_dim <= 96) { m = 192; n = 144; }
    else if (head_dim <= 128) { m = 128; n = 176; }
    else if (head_dim <= 192) { m = 128; n = 112; }
    else { m = 128; n = 80; }
  } else {
    if (head_dim <= 64) { m = 192; n = 160; }
    else if (head_dim <= 96) { m = 192; n = 128; }
    else if (head_dim <= 128) { m = 128; n = v_colmajor ? 192 : 224; }
    else if (head_dim <= 192) { m = 128; n = 160; }
    else { m = 128; n = 128; }
  }
  if (block_m) *block_m = m;
  if (block_n) *block_n = n;
  return (element_size == 2 && head_dim == LA_HEAD_DIM) ? LA_OK : LA_ERR_UNSUPPORTED;
}

int la_rope_cast_sm100(const la_rope_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LA_CHECK_ARG(p != nullptr && p->x && p->out && p->cos_sin && p->grid, "la_rope_cast_sm100: NULL argument");
  LA_CHECK_ARG(p->b > 0 && p->s > 0 && p->h > 0 && p->d > 0 && p->d % 8 == 0 && p->max_pos > 0,
               "la_rope_cast_sm100: bad sizes (d must be a multiple of 8)");
  const int elt = p->x_is_bf16 ? 2 : 4;
  LA_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) % 16) == 0 && aligned16(p->out) &&
                   (p->x_batch_stride * elt) % 16 == 0 && (p->x_row_stride * elt) % 16 == 0 &&
                   (p->x_head_stride * elt) % 16 == 0,
               "la_rope_cast_sm100: x / out must be 16-byte aligned with 16-byte aligned strides");
  la::RopeKernelArgs a;
  a.x = p->x;
  a.out = static_cast<__nv_bfloat16*>(p->out);
  a.cos_sin = reinterpret_cast<const float2*>(p->cos_sin);
  a.grid = p->grid;
  a.x_batch_stride = p->x_batch_stride;
  a.x_row_stride = p->x_row_stride;
  a.x_head_stride = p->x_head_stride;
  a.b = p->b; a.s = p->s; a.h = p->h; a.d = p->d; a.max_pos = p->max_pos;
  LA_CHECK_ARG(p->d / 8 <= la::kRopeThreads, "la_rope_cast_sm100: head_dim too large");
  const int64_t tokens = (int64_t)p->b * p->s;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = (int)std::min<int64_t>(tokens, (int64_t)sms * 64);   // a multiple of the SM count, token-stride
  if (p->x_is_bf16) la::la_rope_cast_kernel<__nv_bfloat16><<<blocks, la::kRopeThreads, 0, stream>>>(a);
  else la::la_rope_cast_kernel<float><<<blocks, la::kRopeThreads, 0, stream>>>(a);
  LA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return LA_OK;
}

#ifdef LA_PROFILE_CLOCKS
int la_prof_read(unsigned long long out[32], int reset) {
  LA_CUDA(cudaMemcpyFromSymbol(out, la::g_la_prof, 32 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[32] = {0};
    LA_CUDA(cudaMemcpyToSymbol(la::g_la_prof, z, sizeof(z)));
  }
  return LA_OK;
}
#endif

int la_watchdog_read(unsigned int out[4]) {
  LA_CUDA(cudaMemcpyFromSymbol(out, la::g_la_watchdog, 4 * sizeof(unsigned int)));
  return LA_OK;
}

int la_fwd_sm100(const la_fwd_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LA_CHECK_ARG(p != nullptr, "la_fwd_sm100: params is NULL");
  LA_CHECK_ARG(p->q && p->k && p->v && p->out, "la_fwd_sm100: q/k/v/out must be non-NULL");
  if (p->d != LA_HEAD_DIM)
    return fail(LA_ERR_UNSUPPORTED, "la_fwd_sm100: head_dim %d not built (only %d)", p->d, LA_HEAD_DIM);
  LA_CHECK_ARG(p->b > 0 && p->h > 0 && p->h_k > 0 && p->seqlen_q > 0 && p->seqlen_k > 0,
               "la_fwd_sm100: sizes must be positive (b=%d h=%d h_k=%d sq=%d sk=%d)", p->b, p->h, p->h_k,
               p->seqlen_q, p->seqlen_k);
  LA_CHECK_ARG(p->h % p->h_k == 0, "la_fwd_sm100: h (%d) must be a multiple of h_k (%d)", p->h, p->h_k);
  LA_CHECK_ARG(p->h <= 65535 && p->b <= 65535, "la_fwd_sm100: h and b must be <= 65535");
  LA_CHECK_ARG(aligned16(p->q) && aligned16(p->k) && aligned16(p->v) && aligned16(p->out),
               "la_fwd_sm100: q/k/v/out must be 16-byte aligned");
  const int64_t strides[] = {p->q_batch_stride, p->q_row_stride, p->q_head_stride, p->k_batch_stride,
                             p->k_row_stride,   p->k_head_stride, p->v_batch_stride, p->v_row_stride,
                             p->v_head_stride,  p->o_batch_stride, p->o_row_stride, p->o_head_stride};
  for (int64_t s : strides)
    LA_CHECK_ARG(s % 8 == 0 && s >= 0, "la_fwd_sm100: strides must be non-negative multiples of 8 elements (got %lld)",
                 (long long)s);
  const int qtiles = (p->seqlen_q + LA_BLOCK_M - 1) / LA_BLOCK_M;
  const int ktiles = (p->seqlen_k + LA_BLOCK_N - 1) / LA_BLOCK_N;
  if (ktiles > la::kFwdMaxTiles)
    return fail(LA_ERR_UNSUPPORTED, "la_fwd_sm100: seqlen_k %d needs %d K tiles (max %d)", p->seqlen_k, ktiles,
                la::kFwdMaxTiles);

  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap(&tq, p->q, p->b, p->seqlen_q, p->h, p->d, p->q_batch_stride, p->q_row_stride, p->q_head_stride,
                      LA_BLOCK_M, "q")))
    return rc;
  if ((rc = make_tmap(&tk, p->k, p->b, p->seqlen_k, p->h_k, p->d, p->k_batch_stride, p->k_row_stride,
                      p->k_head_stride, LA_BLOCK_N, "k")))
    return rc;
  if ((rc = make_tmap(&tv, p->v, p->b, p->seqlen_k, p->h_k, p->d, p->v_batch_stride, p->v_row_stride,
                      p->v_head_stride, LA_BLOCK_N, "v")))
    return rc;

  la::FwdKernelArgs a;
  memset(&a, 0, sizeof(a));
  a.out = p->out_is_f32 ? nullptr : static_cast<__nv_bfloat16*>(p->out);
  a.out_f32 = p->out_is_f32 ? static_cast<float*>(p->out) : nullptr;
  a.rows_per_peer = 0;
  if (p->out_rows_per_peer > 0) {
    LA_CHECK_ARG(!p->out_is_f32 && p->n_out_peers >= 1 && p->n_out_peers <= 8 &&
                     (int64_t)p->n_out_peers * p->out_rows_per_peer >= p->seqlen_q,
                 "la_fwd_sm100: bad peer scatter (bf16 only, 1..8 peers covering seqlen_q)");
    for (int i = 0; i < p->n_out_peers; ++i) {
      LA_CHECK_ARG(p->out_peer[i] != nullptr && aligned16(p->out_peer[i]), "la_fwd_sm100: out_peer[%d] is NULL or unaligned", i);
      a.out_peer[i] = static_cast<__nv_bfloat16*>(p->out_peer[i]);
    }
    a.rows_per_peer = p->out_rows_per_peer;
  }
  {
    // 256-bit stores need 32-byte aligned destinations: base pointers and row / head / batch strides
    const int elt = p->out_is_f32 ? 4 : 2;
    bool ok = ((p->o_batch_stride * elt) % 32 == 0) && ((p->o_row_stride * elt) % 32 == 0) && ((p->o_head_stride * elt) % 32 == 0);
    if (a.rows_per_peer > 0) {
      for (int i = 0; i < p->n_out_peers; ++i) ok = ok && (reinterpret_cast<uintptr_t>(p->out_peer[i]) % 32 == 0);
    } else {
      ok = ok && (reinterpret_cast<uintptr_t>(p->out) % 32 == 0);
    }
    a.o_align32 = ok ? 1 : 0;
  }
  a.lse = p->lse;
  a.read_list = p->read_list;
  a.tile_stat = p->tile_stat;
  a.o_batch_stride = p->o_batch_stride;
  a.o_row_stride = p->o_row_stride;
  a.o_head_stride = p->o_head_stride;
  a.h = p->h;
  a.h_per_kv = p->h / p->h_k;
  a.seqlen_q = p->seqlen_q;
  a.seqlen_k = p->seqlen_k;
  a.qtiles = qtiles;
  a.ktiles = ktiles;
  a.softmax_scale = p->softmax_scale;
  a.scale_log2 = p->softmax_scale * (float)M_LOG2E;  // mainloop :760

  // The attribute is per function per DEVICE: remember it per device ordinal (a process may drive several GPUs).
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  LA_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    LA_CUDA(cudaFuncSetAttribute(la::la_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, la::kFwdSmemBytes));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  dim3 grid(qtiles, p->h, p->b);
  la::la_fwd_kernel<<<grid, la::kFwdThreads, la::kFwdSmemBytes, stream>>>(tq, tk, tv, a);
  LA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return LA_OK;
}

int la_skip_update_sm100(const la_update_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LA_CHECK_ARG(p != nullptr, "la_skip_update_sm100: params is NULL");
  LA_CHECK_ARG(p->read_list && p->write_list && p->tile_stat,
               "la_skip_update_sm100: read_list/write_list/tile_stat must be non-NULL");
  LA_CHECK_ARG(p->read_list != p->write_list, "la_skip_update_sm100: read and write lists must not alias");
  LA_CHECK_ARG(p->b > 0 && p->h > 0 && p->qtiles > 0 && p->ktiles > 0, "la_skip_update_sm100: bad sizes");
  if (p->ktiles > la::kFwdMaxTiles)
    return fail(LA_ERR_UNSUPPORTED, "la_skip_update_sm100: ktiles %d > %d", p->ktiles, la::kFwdMaxTiles);
  const int64_t rows64 = (int64_t)p->b * p->h * p->qtiles;
  LA_CHECK_ARG(rows64 < (1ll << 31), "la_skip_update_sm100: too many rows");
  la::UpdateKernelArgs a;
  a.read_list = p->read_list;
  a.must_do_list = p->must_do_list;
  a.write_list = p->write_list;
  a.tile_stat = p->tile_stat;
  a.overflow_count = p->overflow_count;
  a.rows = (int)rows64;
  a.ktiles = p->ktiles;
  a.thr = p->thr;
  const int blocks = (a.rows + la::kUpdWarpsPerBlock - 1) / la::kUpdWarpsPerBlock;
  la::la_skip_update_kernel<<<blocks, la::kUpdWarpsPerBlock * 32, la::la_skip_update_smem_bytes(p->ktiles), stream>>>(a);
  LA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return LA_OK;
}

int la_fwd_skip_sm100(const la_fwd_params* fwd, const la_update_params* upd, void* stream) {
  LA_CHECK_ARG(fwd != nullptr && upd != nullptr, "la_fwd_skip_sm100: params are NULL");
  LA_CHECK_ARG(fwd->read_list != nullptr && fwd->tile_stat != nullptr,
               "la_fwd_skip_sm100: the forward needs read_list and tile_stat to feed the update");
  // The update walks the rows the forward just produced: its geometry must be the forward's (a mismatch would read
  // and write the lists and the statistic out of bounds).
  {
    const int qt = (fwd->seqlen_q + 127) / 128, kt = (fwd->seqlen_k + 175) / 176;
    LA_CHECK_ARG(upd->b == fwd->b && upd->h == fwd->h && upd->qtiles == qt && upd->ktiles == kt,
                 "la_fwd_skip_sm100: update geometry (b %d, h %d, qtiles %d, ktiles %d) does not match the forward's "
                 "(b %d, h %d, qtiles %d, ktiles %d)", upd->b, upd->h, upd->qtiles, upd->ktiles, fwd->b, fwd->h, qt, kt);
  }
  int rc = la_fwd_sm100(fwd, stream);
  if (rc) return rc;
  la_update_params u = *upd;
  if (!u.read_list) u.read_list = fwd->read_list;
  if (!u.tile_stat) u.tile_stat = fwd->tile_stat;
  return la_skip_update_sm100(&u, stream);
}

int la_list_pack_sm100(const int32_t* list, uint32_t* bits, int64_t rows, int ktiles, int32_t* bad_rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LA_CHECK_ARG(list && bits && rows > 0 && rows < (1ll << 31) && ktiles > 0, "la_list_pack_sm100: bad arguments");
  if (ktiles > la::kFwdMaxTiles) return fail(LA_ERR_UNSUPPORTED, "la_list_pack_sm100: ktiles %d > %d", ktiles, la::kFwdMaxTiles);
  la::la_list_pack_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(list, bits, (int)rows, ktiles, bad_rows);
  LA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return LA_OK;
}

int la_list_unpack_sm100(const uint32_t* bits, int32_t* list, int64_t rows, int ktiles, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LA_CHECK_ARG(list && bits && rows > 0 && rows < (1ll << 31) && ktiles > 0, "la_list_unpack_sm100: bad arguments");
  if (ktiles > la::kFwdMaxTiles) return fail(LA_ERR_UNSUPPORTED, "la_list_unpack_sm100: ktiles %d > %d", ktiles, la::kFwdMaxTiles);
  la::la_list_unpack_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(bits, list, (int)rows, ktiles);
  LA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return LA_OK;
}

int la_copy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width, size_t rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LA_CHECK_ARG(dst && src && width > 0 && rows > 0 && dst_pitch >= width && src_pitch >= width, "la_copy2d_async: bad arguments");
  LA_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, rows, cudaMemcpyDefault, stream));
  return LA_OK;
}

int la_combine_sm100(const la_combine_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LA_CHECK_ARG(p != nullptr && p->o_parts && p->lse_parts && p->out, "la_combine_sm100: NULL argument");
  LA_CHECK_ARG(p->n_parts >= 1 && p->n_parts <= 8, "la_combine_sm100: n_parts must be in [1, 8] (got %d)", p->n_parts);
  LA_CHECK_ARG(p->b > 0 && p->h > 0 && p->s > 0 && p->d > 0 && p->d % 8 == 0, "la_combine_sm100: bad sizes");
  la::CombineKernelArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < p->n_parts; ++i) {
    LA_CHECK_ARG(p->o_parts[i] && p->lse_parts[i] && aligned16(p->o_parts[i]), "la_combine_sm100: bad part %d", i);
    a.o_parts[i] = p->o_parts[i];
    a.lse_parts[i] = p->lse_parts[i];
  }
  LA_CHECK_ARG(aligned16(p->out), "la_combine_sm100: out must be 16-byte aligned");
  a.out = p->out;
  a.lse = p->lse;
  a.n_parts = p->n_parts;
  a.b = p->b;
  a.h = p->h;
  a.s = p->s;
  a.d = p->d;
  const int64_t total = (int64_t)p->b * p->s * p->h * (p->d / 8);
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  if (p->parts_are_f32) {
    if (p->out_is_f32) la::la_combine_kernel<float, float><<<blocks, 256, 0, stream>>>(a);
    else la::la_combine_kernel<float, __nv_bfloat16><<<blocks, 256, 0, stream>>>(a);
  } else {
    if (p->out_is_f32) la::la_combine_kernel<__nv_bfloat16, float><<<blocks, 256, 0, stream>>>(a);
    else la::la_combine_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, stream>>>(a);
  }
  LA_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return LA_OK;
}

}  // extern "C"
