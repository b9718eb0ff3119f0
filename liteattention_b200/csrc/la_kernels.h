// la_kernels.h -- internal interface between the host glue (la_api.cu) and the kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace la {

constexpr int kFwdThreads = 384;        // 8 softmax warps + one warpgroup holding the TMA warp and the MMA warp
constexpr int kFwdMaxTiles = 2048;      // K tiles one CTA can visit (seqlen_k <= 360448)
constexpr int kFwdSmemBytes = 231680;   // dynamic shared memory of la_fwd_kernel (incl. 1 KB alignment slack)

struct FwdKernelArgs {
  __nv_bfloat16* out;
  float* out_f32;       // when non-NULL the epilogue writes fp32 here (same element strides) instead of bf16 to out
  // Sequence-parallel scatter (rows_per_peer > 0): query row r goes to out_peer[r / rows_per_peer] at row
  // r % rows_per_peer -- peer GPUs' buffers mapped over NVLink; `out` is ignored.
  __nv_bfloat16* out_peer[8];
  int32_t rows_per_peer;
  int32_t o_align32;    // every O row segment the epilogue writes starts on a 32-byte boundary: 256-bit stores
  float* lse;
  const int32_t* read_list;
  float* tile_stat;
  int64_t o_batch_stride, o_row_stride, o_head_stride;
  int32_t h, h_per_kv, seqlen_q, seqlen_k, qtiles, ktiles;
  float softmax_scale;  // multiplies raw S
  float scale_log2;     // softmax_scale * log2(e)
};

__global__ void la_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                              const __grid_constant__ CUtensorMap tmap_v, const FwdKernelArgs args);

struct UpdateKernelArgs {
  const int32_t* read_list;
  const int32_t* must_do_list;
  int32_t* write_list;
  const float* tile_stat;
  int32_t* overflow_count;
  int32_t rows;    // b * h * qtiles
  int32_t ktiles;
  float thr;
};

constexpr int kUpdWarpsPerBlock = 4;
__global__ void la_skip_update_kernel(const UpdateKernelArgs args);
size_t la_skip_update_smem_bytes(int ktiles);

__global__ void la_list_pack_kernel(const int32_t* list, uint32_t* bits, int rows, int ktiles, int32_t* bad_rows);
__global__ void la_list_unpack_kernel(const uint32_t* bits, int32_t* list, int rows, int ktiles);

struct CombineKernelArgs {
  const void* o_parts[8];      // (b, s, h, d) contiguous, bf16 or fp32 (template parameter In)
  const float* lse_parts[8];   // (b, h, s) fp32 contiguous
  void* out;                   // (b, s, h, d) contiguous, bf16 or fp32 (template parameter Out)
  float* lse;
  int32_t n_parts, b, h, s, d;
};
template <typename In, typename Out>
__global__ void la_combine_kernel(const CombineKernelArgs args);

struct RopeKernelArgs {
  const void* x;            // (b, s, h, d) fp32 or bf16, last dim contiguous, element strides below
  __nv_bfloat16* out;       // (b, s, h, d) bf16 contiguous
  const float2* cos_sin;    // [max_pos, d/2] (cos, sin) of the three axes' angles, concatenated along d/2
  const int32_t* grid;      // [b, 3] (frames, height, width) per sample, device memory
  int64_t x_batch_stride, x_row_stride, x_head_stride;
  int32_t b, s, h, d, max_pos;
};
template <typename In>
__global__ void la_rope_cast_kernel(const RopeKernelArgs args);

}  // namespace la
