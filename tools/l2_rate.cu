// tools/l2_rate.cu -- how many bytes per SM clock can TMA pull from L2 into shared memory when every SM streams
// K/V tiles the way la_fwd_kernel does (176 rows x 256 B, rows strided by H*D*2 bytes, 2 boxes of 64 columns)?
// At 128 query rows per CTA the forward kernel needs 64 B/clk/SM at the tensor floor (90 KB per 1408 clk), so this
// number bounds the kernel no matter how good the softmax schedule is.
//   mode 0: unicast, every CTA its own tile stream      mode 1: unicast, CTA pairs stream the SAME tiles
//   mode 2: 2-CTA clusters, each CTA loads half a tile and multicasts it to both
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I liteattention_b200/csrc -o tools/_build/l2_rate tools/l2_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>

#include "la_ptx.cuh"

using namespace la;

constexpr int kN = 176;
constexpr uint32_t kBlockBytes = kN * 128, kTileBytes = 2 * kBlockBytes;
constexpr int kStages = 4;
constexpr uint32_t kSmem = kStages * kTileBytes + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                               int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster_acq(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}

struct Res { unsigned long long cycles; };

template <int CL>
__global__ void __launch_bounds__(64, 1) l2_rate_kernel(const __grid_constant__ CUtensorMap tm, int mode, int iters,
                                                        int heads, int ktiles, Res* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kStages * kTileBytes;
  auto full = [&](int s) { return bar0 + s * 8; };
  auto empty = [&](int s) { return bar0 + (kStages + s) * 8; };
  const uint32_t rank = (CL > 1) ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), CL);
    }
    fence_mbar_init();
  }
  if (CL > 1) cluster_sync_all(); else __syncthreads();

  // which tile stream does this CTA follow?
  const int stream = (mode == 0) ? blockIdx.x : blockIdx.x / 2;
  const int head = stream % heads;
  int tile = (stream * 37) % ktiles;

  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      const uint32_t par = (it / kStages) & 1;
      if (CL > 1) { while (!mbar_try_wait_cluster_acq(empty(s), par ^ 1)) {} }
      else mbar_wait(empty(s), par ^ 1);
      mbar_arrive_expect_tx(full(s), kTileBytes);
      const uint32_t dst = sb + s * kTileBytes;
      if (CL == 1) {
        tma_load_4d(dst, &tm, full(s), 0, tile * kN, head, 0);
        tma_load_4d(dst + kBlockBytes, &tm, full(s), 64, tile * kN, head, 0);
      } else {
        tma_load_4d_mc(dst + rank * kBlockBytes, &tm, full(s), rank * 64, tile * kN, head, 0, (uint16_t)0x3);
      }
      tile = (tile == 0) ? ktiles - 1 : tile - 1;
    }
    // drain
    for (int it = iters; it < iters + kStages; ++it) {
      const int s = it % kStages;
      const uint32_t par = (it / kStages) & 1;
      if (CL > 1) { while (!mbar_try_wait_cluster_acq(empty(s), par ^ 1)) {} }
      else mbar_wait(empty(s), par ^ 1);
    }
    res[blockIdx.x].cycles = (unsigned long long)(clock64() - t0);
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages;
      mbar_wait(full(s), (it / kStages) & 1);
      if (CL == 1) mbar_arrive(empty(s));
      else {
        mbar_arrive_cluster(mapa(empty(s), 0));
        mbar_arrive_cluster(mapa(empty(s), 1));
      }
    }
  }
  if (CL > 1) cluster_sync_all(); else __syncthreads();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  const int S = 75600, H = 4, D = 128, ktiles = (S + kN - 1) / kN;
  void* buf;
  const size_t bytes = (size_t)S * H * D * 2;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0x11, bytes);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)H, 1};
  cuuint64_t strides[3] = {(cuuint64_t)H * D * 2, (cuuint64_t)D * 2, (cuuint64_t)S * H * D * 2};
  cuuint32_t box[4] = {64, kN, 1, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  Res* d_res;
  const int nsm = 148;
  cudaMalloc(&d_res, nsm * sizeof(Res));
  cudaFuncSetAttribute(l2_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  cudaFuncSetAttribute(l2_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  const char* names[3] = {"unicast, own stream per CTA", "unicast, CTA pairs share a stream", "cluster-2 multicast (half tile each)"};
  for (int heads = 1; heads <= 4; heads *= 2) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (mode < 2) {
          l2_rate_kernel<1><<<nsm, 64, kSmem>>>(tm, mode, iters, heads, ktiles, d_res);
        } else {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(nsm); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = kSmem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, l2_rate_kernel<2>, tm, mode, iters, heads, ktiles, d_res);
        }
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(err)); return 1; }
        if (rep == 0) continue;
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<Res> h(nsm);
        cudaMemcpy(h.data(), d_res, nsm * sizeof(Res), cudaMemcpyDeviceToHost);
        double sum = 0, mx = 0;
        for (auto& x : h) { sum += (double)x.cycles; mx = x.cycles > mx ? (double)x.cycles : mx; }
        const double avg = sum / nsm;
        const double bpc = (double)iters * kTileBytes / avg;
        printf("heads=%d (%5.1f MB)  %-40s %6.1f B/clk/SM into smem (%7.0f B/clk chip)  %6.2f TB/s  %.3f ms  ~%.0f MHz  => tensor-floor share %.0f%%\n",
               heads, heads * (double)S * D * 2 / 1e6, names[mode], bpc, bpc * nsm, (double)iters * kTileBytes * nsm / (ms * 1e-3) / 1e12,
               ms, avg / (ms * 1e-3) / 1e6, 100.0 * bpc / 64.0);
      }
    }
  }
  return 0;
}
