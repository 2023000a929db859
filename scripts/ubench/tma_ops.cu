// Micro-benchmark (profiling aid): cost structure of TMA tile loads on one SM -- is a load paid per byte, per 128-byte
// row request, or per instruction?  Each CTA streams "stages" made of several boxes through a ring and frees them at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_ops tma_ops.cu ../../autoformer_b200/csrc/avc_host.o
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../autoformer_b200/csrc/avc_host.h"
#include "../../autoformer_b200/csrc/avc_ptx.cuh"

using namespace avc;

constexpr int kMaxOps = 4;
constexpr int kRing = 192 * 1024;

struct Params {
  CUtensorMap tmap[kMaxOps];   // 2-D maps over [rows][pitch], box {64 bf16, rows_op}
  int rows_op[kMaxOps];
  int n_ops;
  int stage_bytes, stages;
  int iters;                   // stages streamed per CTA
  int region_rows;             // rows each CTA cycles through (L2-resident when small)
  long long cta_stride_rows;   // distance between the regions of consecutive CTAs
  long long* cycles;
};

__global__ void __launch_bounds__(64, 1) ops_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + kRing);
  uint64_t* empty = full + 64;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const long long row_base = blockIdx.x * p.cta_stride_rows;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    uint32_t stage = 0, phase = 0;
    int row = 0;
    for (int i = 0; i < p.iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1u);
      mbar_arrive_expect_tx(&full[stage], p.stage_bytes);
      uint8_t* dst = base + stage * p.stage_bytes;
      for (int o = 0; o < p.n_ops; ++o) {
        tma_load_2d(dst, &p.tmap[o], &full[stage], 0, (int)(row_base + row));
        dst += p.rows_op[o] * 128;
        row += p.rows_op[o];
        if (row + 256 > p.region_rows) row = 0;
      }
      if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
    }
    // drain
    for (int i = 0; i < p.stages; ++i) {
      mbar_wait(&empty[stage], phase ^ 1u);
      if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
    }
    p.cycles[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 32) {
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < p.iters; ++i) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
    }
  }
}

static void run(const char* label, void* buf, long long buf_rows, int pitch_bytes, std::vector<int> rows_op,
                int region_rows, bool dram, int grid = 148, bool shared_region = false, int max_stages = 8) {
  Params p;
  memset(&p, 0, sizeof(p));
  p.n_ops = (int)rows_op.size();
  p.stage_bytes = 0;
  for (int o = 0; o < p.n_ops; ++o) {
    p.rows_op[o] = rows_op[o];
    p.stage_bytes += rows_op[o] * 128;
    if (!encode_tmap_2d(&p.tmap[o], 2, buf, pitch_bytes / 2, (uint64_t)buf_rows, pitch_bytes, 64, rows_op[o])) {
      printf("tmap failed\n");
      exit(1);
    }
  }
  p.stages = kRing / p.stage_bytes;
  if (p.stages > max_stages) p.stages = max_stages;
  p.iters = 3000;
  p.region_rows = region_rows;
  p.cta_stride_rows = dram ? buf_rows / grid : region_rows;
  if (dram) p.region_rows = (int)(buf_rows / grid);
  if (shared_region) p.cta_stride_rows = 0;      // every CTA walks the same rows
  cudaMalloc(&p.cycles, grid * sizeof(long long));
  const int smem = kRing + 2 * 64 * 8 + 1024;
  cudaFuncSetAttribute(ops_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    ops_kernel<<<grid, 64, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: failed %s\n", label, cudaGetErrorString(e));
      exit(1);
    }
  }
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), p.cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (long long c : cyc) mean += (double)c;
  mean /= grid;
  const double per_stage = mean / p.iters;
  printf("%-46s ops %d stage %3d KB x %d stages: %7.0f cycles/stage  %6.0f cycles/op  %5.1f B/clk/SM\n", label, p.n_ops,
         p.stage_bytes / 1024, p.stages, per_stage, per_stage / p.n_ops, p.stage_bytes / per_stage);
  cudaFree(p.cycles);
}

int main() {
  // 4 GB buffer viewed with two pitches
  const size_t bytes = (size_t)4 << 30;
  void* buf;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 1, bytes);
  const long long rows128 = bytes / 128, rows2k = bytes / 2048;
  const int L2rows = 1024;    // per-CTA region that stays in L2
  run("1 x 16KB, pitch 2KB, L2", buf, rows2k, 2048, {128}, L2rows, false);
  run("1 x 16KB, pitch 128B (contiguous), L2", buf, rows128, 128, {128}, L2rows, false);
  run("1 x 32KB (256 rows), pitch 128B, L2", buf, rows128, 128, {256}, L2rows, false);
  run("1 x 8KB, pitch 128B, L2", buf, rows128, 128, {64}, L2rows, false);
  run("1 x 4KB, pitch 128B, L2", buf, rows128, 128, {32}, L2rows, false);
  run("2 ops 16KB + 4KB (narrow conv k-block), L2", buf, rows128, 128, {128, 32}, L2rows, false);
  run("4 ops 16+16+8+8KB (LSTM split stage), L2", buf, rows2k, 2048, {128, 128, 64, 64}, L2rows, false);
  run("2 ops 16+16KB (conv BN=256 pair k-block), L2", buf, rows2k, 2048, {128, 128}, L2rows, false);
  run("1 x 16KB, pitch 128B, DRAM stream", buf, rows128, 128, {128}, 0, true);
  run("2 ops 16KB + 4KB, DRAM stream", buf, rows128, 128, {128, 32}, 0, true);
  run("1 x 16KB, pitch 128B, L2, 74 CTAs", buf, rows128, 128, {128}, L2rows, false, 74);
  run("1 x 16KB, pitch 128B, DRAM stream, 74 CTAs", buf, rows128, 128, {128}, 0, true, 74);
  // rows of one box far apart in memory, as the recurrent-state tiles of the LSTM are (one row per utterance,
  // utterances T * H * 4 bytes apart): does the number of distinct pages a box touches matter?
  const long long rows512k = bytes / (512 << 10);
  run("1 x 16KB (128 rows), pitch 512KB, L2", buf, rows512k, 512 << 10, {128}, 1024, false, 148, true);
  run("1 x 4KB (32 rows), pitch 512KB, L2", buf, rows512k, 512 << 10, {32}, 1024, false, 148, true);
  run("2 x 4KB (32 rows), pitch 512KB, L2", buf, rows512k, 512 << 10, {32, 32}, 1024, false, 148, true);
  run("1 x 16KB (128 rows), pitch 2KB, shared region", buf, rows2k, 2048, {128}, 1024, false, 148, true);
  run("2 x 4KB (32 rows), pitch 2KB, shared region", buf, rows2k, 2048, {32, 32}, 1024, false, 148, true);
  // latency or throughput?  the same single 4 KB op with 2 .. 32 stages in flight
  for (int st : {2, 4, 8, 16, 32}) {
    char label[64];
    snprintf(label, sizeof(label), "1 x 4KB, pitch 128B, L2, ring of %d", st);
    run(label, buf, rows128, 128, {32}, L2rows, false, 148, false, st);
  }
  for (int st : {2, 4, 8}) {
    char label[64];
    snprintf(label, sizeof(label), "4 x 4KB, pitch 128B, L2, ring of %d", st);
    run(label, buf, rows128, 128, {32, 32, 32, 32}, L2rows, false, 148, false, st);
  }
  return 0;
}
