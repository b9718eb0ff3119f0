// ref_softmax_harness.cu -- drives the REFERENCE's own online-softmax / QK-skip arithmetic on the GPU.
// TEST INFRASTRUCTURE ONLY.
//
// The reference forward kernel is sm_90a-only (wgmma), but the code that decides m, l, alpha, P and the skip vote
// is plain CuTe register-tensor code with no wgmma in it.  This harness includes it, unmodified, from where it lies:
//   /root/reference/hopper/_internal/cpp/softmax.h   flash::Softmax<2,0>::max_get_scale_detect_qk_skip (:139-222),
//                                                     online_softmax (:263-273) -> scale_apply_exp2 (:81-121) +
//                                                     reduce_sum (:71-79), finalize (:275-296)
//   /root/reference/hopper/_internal/cpp/mask.h      flash::Mask<128,176,false,TiledMmaQK>::apply<Seqlenk_mask> (:46-76)
//   /root/reference/hopper/_internal/cpp/utils.h     convert_layout_acc_rowcol (:124-146), convert_type_out (:211-225)
// and instantiates them exactly as the kernel does for hdim 128 / bf16 (flash_fwd_kernel_sm90.h:514,
// mainloop_fwd_sm90_tma_gmma_ws.hpp:283-290, :1520-1522): 256 MMA threads = two warpgroups of 64 rows each, kNRows = 2,
// the S fragment laid out by CuTe's own SM90 64x176x16 GMMA atom (only its LAYOUT is used -- partition_C of an
// identity tensor tells every thread which (row, col) each fragment register holds; no wgmma is executed).
// The call sequence per visited tile is the mainloop's (:1626-1643 first tile, :1702-1740 further tiles, :1843 finalize).
// The two warpgroups share one skip_tests[4] array in the reference (assign by warpgroup 0, &= by warpgroup 1,
// :1374, softmax.h:208-218, with a latent race); the harness orders them with a block barrier, i.e. it implements the
// race-free reading: tile vote = AND over all 8 consumer warps.
//
// Built by oracle/Makefile (target `ref`) into oracle/_ref/softmax_ref (git-ignored, travels to the GPU box).
// tests/test_ref_softmax_gpu.py feeds it S tiles and compares m / alpha / vote / bf16 P / LSE / 1/l with
// oracle/attention.py (softmax_tile_sequence) and the vote with la_fwd_kernel's statistic.
//
// File format (little endian):
//   in : int32 T, n_block_first, seqlen_q, seqlen_k, m_block;  float scale_log2, thr;  then T x 128 x 176 floats
//        (raw S = Q K^T of the visited tiles in visit order; the first one still unmasked)
//   out: for each tile: 128 row_max, 128 scores_scale, 1 vote (0/1 as float), 128 x 176 P (bf16 widened to float);
//        then 128 LSE (row_sum after finalize, natural log domain) and 128 final scales (1 / row_sum)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cute/tensor.hpp>
#include <cutlass/numeric_types.h>
#include <cutlass/numeric_conversion.h>
#include <cutlass/fast_math.h>
#include <cute/arch/mma_sm90_gmma.hpp>
#include <cute/atom/mma_traits_sm90_gmma.hpp>
#include <cute/atom/mma_atom.hpp>

#include "utils.h"
#include "softmax.h"
#include "mask.h"

using namespace cute;

constexpr int kBlockM = 128, kBlockN = 176, kHeadDim = 128;
using Element = cutlass::bfloat16_t;
using TileShape_MNK = Shape<Int<kBlockM>, Int<kBlockN>, Int<kHeadDim>>;
using AtomLayoutQK = Layout<Shape<Int<kBlockM / 64>, _1, _1>>;                       // mainloop :283
using TiledMmaQK = decltype(cute::make_tiled_mma(
    decltype(cute::GMMA::ss_op_selector<Element, Element, float, TileShape_MNK>()){}, AtomLayoutQK{}));   // :284-290
static_assert(size(TiledMmaQK{}) == 256, "two consumer warpgroups");

__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ S, float* __restrict__ out, int T,
                                                      int n_block_first, int seqlen_q, int seqlen_k, int m_block,
                                                      float scale_log2, float thr) {
  __shared__ int skip_tests[4];
  const int thread_idx = threadIdx.x;
  const int wg = thread_idx / 128;
  TiledMmaQK tiled_mma_qk;
  auto thr_mma = tiled_mma_qk.get_thread_slice(thread_idx);
  Tensor tSrS = partition_fragment_C(tiled_mma_qk, select<0, 1>(TileShape_MNK{}));      // mainloop :1614
  Tensor cS = cute::make_identity_tensor(Shape<Int<kBlockM>, Int<kBlockN>>{});
  Tensor tScS = thr_mma.partition_C(cS);                                               // (row, col) of every register

  flash::Softmax<2 * (2 * kBlockM / 256), 0> softmax(scale_log2);                       // flash_fwd_kernel_sm90.h:514
  flash::Mask<kBlockM, kBlockN, false, TiledMmaQK> mask(thread_idx, seqlen_q, seqlen_k, -1, -1, 0,
                                                        cutlass::FastDivmod(1), cutlass::FastDivmod(1));   // :1520-1522
  const size_t per_tile = 128 + 128 + 1 + (size_t)kBlockM * kBlockN;
  if (thread_idx < 4) skip_tests[thread_idx] = 0;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const float* St = S + (size_t)t * kBlockM * kBlockN;
#pragma unroll
    for (int i = 0; i < size(tSrS); ++i) tSrS(i) = St[get<0>(tScS(i)) * kBlockN + get<1>(tScS(i))];
    float* o = out + (size_t)t * per_tile;
    auto scores_scale = make_fragment_like(softmax.row_max);
    if (t == 0) {
      mask.template apply<true /*Seqlenk_mask*/, false, false>(tSrS, m_block, n_block_first);           // :1626
      if (wg == 0) cute::copy(softmax.template max_get_scale_detect_qk_skip<true, true, false>(tSrS, thr, skip_tests), scores_scale);
      __syncthreads();
      if (wg == 1) cute::copy(softmax.template max_get_scale_detect_qk_skip<true, true, true>(tSrS, thr, skip_tests), scores_scale);
      __syncthreads();
      softmax.template online_softmax<true, true>(tSrS);                                                 // :1641
    } else {
      if (wg == 0) cute::copy(softmax.template max_get_scale_detect_qk_skip<false, false, false>(tSrS, thr, skip_tests), scores_scale);
      __syncthreads();
      if (wg == 1) cute::copy(softmax.template max_get_scale_detect_qk_skip<false, false, true>(tSrS, thr, skip_tests), scores_scale);
      __syncthreads();
      softmax.template online_softmax<false, false>(tSrS);                                               // :1716
    }
    const int vote = skip_tests[0] & skip_tests[1] & skip_tests[2] & skip_tests[3];                      // :1721-1725
    // P -> bf16 exactly as the kernel does before the PV GEMM (:1735, utils.h:211-225)
    Tensor tOrP = make_tensor_like<Element>(tSrS);
    flash::convert_type_out(tSrS, tOrP);
    Tensor rows = make_tensor(tScS.data(), flash::convert_layout_acc_rowcol(tScS.layout()));            // (nrow, ncol) coords
#pragma unroll
    for (int mi = 0; mi < size(softmax.row_max); ++mi) {
      const int row = get<0>(rows(mi, _0{}));
      if ((thread_idx & 3) == 0) {          // the four threads of a quad hold the same row statistics
        o[row] = softmax.row_max(mi);
        o[128 + row] = scores_scale(mi);
      }
    }
    if (thread_idx == 0) o[256] = (float)vote;
#pragma unroll
    for (int i = 0; i < size(tSrS); ++i)
      o[257 + get<0>(tScS(i)) * kBlockN + get<1>(tScS(i))] = static_cast<float>(tOrP(i));
    __syncthreads();
  }
  auto fin = softmax.finalize(1.f);                                                                      // :1843
  Tensor rows = make_tensor(tScS.data(), flash::convert_layout_acc_rowcol(tScS.layout()));
  float* o = out + (size_t)T * per_tile;
#pragma unroll
  for (int mi = 0; mi < size(softmax.row_sum); ++mi) {
    const int row = get<0>(rows(mi, _0{}));
    if ((thread_idx & 3) == 0) {
      o[row] = softmax.row_sum(mi);        // finalize() leaves the LSE here (softmax.h:293)
      o[128 + row] = fin(mi);
    }
  }
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open in"); return 2; }
  int hdr[5]; float fh[2];
  if (fread(hdr, 4, 5, f) != 5 || fread(fh, 4, 2, f) != 2) return 2;
  const int T = hdr[0];
  const size_t ns = (size_t)T * kBlockM * kBlockN;
  const size_t per_tile = 128 + 128 + 1 + (size_t)kBlockM * kBlockN;
  const size_t no = (size_t)T * per_tile + 256;
  std::vector<float> S(ns), O(no, 0.f);
  if (fread(S.data(), 4, ns, f) != ns) return 2;
  fclose(f);
  float *d_s, *d_o;
  cudaMalloc(&d_s, ns * 4); cudaMalloc(&d_o, no * 4);
  cudaMemcpy(d_s, S.data(), ns * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_o, 0, no * 4);
  softmax_kernel<<<1, 256>>>(d_s, d_o, T, hdr[1], hdr[2], hdr[3], hdr[4], fh[0], fh[1]);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(O.data(), d_o, no * 4, cudaMemcpyDeviceToHost);
  f = fopen(argv[2], "wb");
  fwrite(O.data(), 4, no, f);
  fclose(f);
  return 0;
}
