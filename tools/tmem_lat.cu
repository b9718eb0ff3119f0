// tools/tmem_lat.cu -- latency of tcgen05.wait::st / wait::ld as a function of how long ago the st/ld was issued.
#include <cstdio>
#include <cuda_bf16.h>
#include "la_ptx.cuh"
#include "la_tmem_ptx.cuh"
using namespace la;
__global__ void __launch_bounds__(256, 1) k(int delay_iters, int nw, long long* res, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  float acc = threadIdx.x;
  long long t_issue = 0, t_wait = 0, t_ld = 0, t_fence = 0;
  if (warp < nw) {
    for (int it = 0; it < 200; ++it) {
      uint32_t pr[44];
      for (int j = 0; j < 44; ++j) pr[j] = it + j;
      long long a = clock64();
      tmem_st_x32(tm, pr); tmem_st_x8(tm + 32, pr + 32); tmem_st_x4(tm + 40, pr + 40);
      long long b = clock64();
      for (int j = 0; j < delay_iters; ++j) acc = fmaf(acc, 1.0001f, 0.5f);   // dependent chain: 4 clk each
      long long c = clock64();
      tmem_wait_st();
      long long d = clock64();
      tc_fence_before();
      long long e = clock64();
      uint32_t r[32];
      tmem_ld_x32(tm, r); tmem_wait_ld();
      long long f = clock64();
      acc += __uint_as_float(r[5]) * 1e-30f;
      t_issue += b - a; t_wait += d - c; t_fence += e - d; t_ld += f - e;
    }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { res[warp * 4] = t_issue / 200; res[warp * 4 + 1] = t_wait / 200; res[warp * 4 + 2] = t_fence / 200; res[warp * 4 + 3] = t_ld / 200; }
  }
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}
int main() {
  long long* d_res; float* d_sink; cudaMalloc(&d_res, 8 * 4 * 8); cudaMalloc(&d_sink, 148 * 256 * 4);
  for (int nw : {1, 8}) for (int delay : {0, 25, 50, 100, 200}) {
    k<<<148, 256>>>(delay, nw, d_res, d_sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
    long long h[32]; cudaMemcpy(h, d_res, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps=%d delay=%4d clk: st issue %4lld  wait::st %4lld  fence %3lld  ld x32+wait %4lld   (warp 0);  warp %d: wait::st %4lld\n",
           nw, delay * 4, h[0], h[1], h[2], h[3], nw - 1, h[(nw - 1) * 4 + 1]);
  }
  return 0;
}
