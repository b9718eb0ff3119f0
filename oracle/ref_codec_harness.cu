// ref_codec_harness.cu -- drives the REFERENCE's own skip-list codec structs on the GPU.  TEST INFRASTRUCTURE ONLY.
//
// flash::SkipListReader / flash::SkipListWriter are taken, unmodified, from where they lie:
//   /root/reference/hopper/_internal/cpp/mainloop_fwd_sm90_tma_gmma_ws.hpp:47-192
// (they are plain CUDA device structs; only their template `init` depends on the CuTe kernel params and is
// not used here).  The driver loop below follows the reference's range loop (same file, :1804-1827) with the
// per-tile vote supplied from a file instead of from the softmax.  Built by oracle/Makefile into
// oracle/_ref/skiplist_ref (git-ignored, travels to the GPU box); the gpu tests feed it random lists/votes
// and compare its output with oracle/skiplist.py and with la_skip_update_kernel.
//
// File format (little endian int32): in : rows, ktiles, use_md, then rows*(ktiles+1) read list,
//                                         rows*(ktiles+1) must-do list, rows*ktiles votes (1 = skip)
//                                    out: rows * 2*(ktiles+1) ints (row stride doubled: the reference writer
//                                         is unbounded, the slack catches its overflow instead of corrupting).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "flash.h"
#include "utils.h"  // block.h (pulled in by the mainloop header) needs flash::round_up from here
#include "mainloop_fwd_sm90_tma_gmma_ws.hpp"

__global__ void codec_kernel(const int* read, const int* mustdo, int* write, const int* votes, int rows, int ktiles,
                             int use_md) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int stride = ktiles + 1;
  flash::SkipListReader skip_reader, must_do_reader;
  skip_reader.list_ptr = read + (size_t)row * stride;
  skip_reader.skip_list_len = skip_reader.list_ptr[0];
  skip_reader.load_range();
  must_do_reader.list_ptr = mustdo + (size_t)row * stride;
  must_do_reader.skip_list_len = must_do_reader.list_ptr[0];
  must_do_reader.load_range();
  flash::SkipListWriter skip_writer;
  skip_writer.list_ptr = write + (size_t)row * 2 * stride;
  skip_writer.is_saving_thread = true;
  const int* v = votes + (size_t)row * ktiles;

  int n_block = skip_reader.start_idx;
  bool skip = false;
  skip_writer.record_transition(skip, n_block);
  --n_block;
  do {
    for (; n_block >= skip_reader.end_idx; n_block--) {
      skip = v[n_block] != 0;
      skip_writer.record_transition(skip, n_block, use_md ? &must_do_reader : nullptr);
    }
    skip_writer.record_range_end(skip, skip_reader.end_idx);
    skip_reader.advance();
    if (!skip_reader.has_more()) break;
    skip_reader.load_range();
    n_block = skip_reader.start_idx;
  } while (true);
  skip_writer.finalize();
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open in"); return 2; }
  int hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 2;
  const int rows = hdr[0], ktiles = hdr[1], use_md = hdr[2];
  const size_t nl = (size_t)rows * (ktiles + 1), nv = (size_t)rows * ktiles;
  std::vector<int> rd(nl), md(nl), vt(nv), wr(2 * nl, 0);
  if (fread(rd.data(), 4, nl, f) != nl || fread(md.data(), 4, nl, f) != nl || fread(vt.data(), 4, nv, f) != nv) return 2;
  fclose(f);
  int *d_rd, *d_md, *d_wr, *d_vt;
  cudaMalloc(&d_rd, nl * 4); cudaMalloc(&d_md, nl * 4); cudaMalloc(&d_wr, 2 * nl * 4); cudaMalloc(&d_vt, nv * 4);
  cudaMemcpy(d_rd, rd.data(), nl * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_md, md.data(), nl * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_vt, vt.data(), nv * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_wr, 0, 2 * nl * 4);
  codec_kernel<<<(rows + 127) / 128, 128>>>(d_rd, d_md, d_wr, d_vt, rows, ktiles, use_md);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(wr.data(), d_wr, 2 * nl * 4, cudaMemcpyDeviceToHost);
  f = fopen(argv[2], "wb");
  fwrite(wr.data(), 4, 2 * nl, f);
  fclose(f);
  return 0;
}
