// Micro-benchmark (profiling aid, not part of the library): how fast can the SMs of a B200 pull 16 KB operand tiles
// out of L2 with TMA when (a) every CTA reads its own data, (b) groups of CTAs read the SAME tiles at the same time
// (the h tiles of the LSTM recurrence), (c) the tiles are TMA-multicast across a cluster.  Decides whether the
// recurrence mainloop (measured ~47 B/clk/SM) is bound by L2 slice bandwidth (multicast helps) or by SM ingress.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_stream tma_stream.cu ../../autoformer_b200/csrc/avc_host.o
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../autoformer_b200/csrc/avc_host.h"
#include "../../autoformer_b200/csrc/avc_ptx.cuh"

using namespace avc;

constexpr int kTile = 16384;     // 128 rows x 128 B
constexpr int kMaxStages = 12;

struct Params {
  CUtensorMap tmap;     // [rows][1024] bf16, box {64, 128 / CS}
  int iters;            // passes over the 16 k-chunks of a block ("frames")
  int group;            // CTAs (CS = 1) or clusters (CS > 1) that read the same block
  int stages;
  long long* cycles;    // per CTA
};

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int CS>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + kMaxStages * kTile);
  uint64_t* empty = full + kMaxStages;
  const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CS;
  const int blk = unit / p.group;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CS);
    }
    fence_mbar_init();
  }
  if (CS > 1) cluster_sync_all(); else __syncthreads();
  const int total = p.iters * 16;
  long long t0 = 0;
  if (threadIdx.x == 0) {          // producer
    t0 = clock64();
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(&empty[stage], phase ^ 1u);
      mbar_arrive_expect_tx(&full[stage], kTile);
      const int kc = (i & 15) * 64;
      if (CS == 1) {
        tma_load_2d(base + stage * kTile, &p.tmap, &full[stage], kc, blk * 128);
      } else {
        constexpr int rows = 128 / CS;
        tma_load_2d_mc(base + stage * kTile + rank * rows * 128, &p.tmap, &full[stage], kc, blk * 128 + rank * rows,
                       (uint16_t)((1u << CS) - 1u));
      }
      if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (threadIdx.x == 32) {  // consumer: frees the stage (in every CTA of the cluster) as soon as it has landed
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(&full[stage], phase);
      if (CS == 1) {
        mbar_arrive(&empty[stage]);
      } else {
        const uint32_t a = smem_u32(&empty[stage]);
#pragma unroll
        for (int r = 0; r < CS; ++r) mbar_arrive_remote(mapa(a, r));
      }
      if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
    }
    p.cycles[blockIdx.x] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
  if (CS > 1) cluster_sync_all();
}

template <int CS>
static void run(const char* label, void* buf, int grid, int group, int iters, int stages) {
  Params p;
  const int rows_total = 128 * 256;
  if (!encode_tmap_2d(&p.tmap, 2, buf, 1024, rows_total, 2048, 64, 128 / CS)) {
    printf("tmap failed\n");
    exit(1);
  }
  p.iters = iters;
  p.group = group;
  p.stages = stages;
  cudaMalloc(&p.cycles, grid * sizeof(long long));
  auto kern = stream_kernel<CS>;
  const int smem = kMaxStages * kTile + 2 * kMaxStages * 8 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (CS > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = -1;
  cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) {
      printf("%s: launch failed %s / %s\n", label, cudaGetErrorString(e), cudaGetErrorString(e2));
      exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), p.cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  long long mx = 0;
  for (long long c : cyc) { mean += (double)c; if (c > mx) mx = c; }
  mean /= grid;
  const double bytes_cta = (double)iters * 16 * kTile;
  printf("%-34s grid %3d cs %2d group %3d stages %2d max_clusters %3d: %7.3f ms  %6.2f TB/s delivered  %5.1f B/clk/SM (mean) %5.1f (slowest)\n",
         label, grid, CS, group, stages, max_clusters, best, bytes_cta * grid / (best * 1e-3) / 1e12, bytes_cta / mean,
         bytes_cta / (double)mx);
  cudaFree(p.cycles);
}

int main() {
  void* buf;
  const size_t bytes = (size_t)128 * 256 * 2048;   // 256 blocks of 128 rows x 2 KB = 64 MB
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 1, bytes);
  const int it = 400;
  for (int st : {4, 8, 12}) {
    run<1>("distinct", buf, 8, 1, it, st);
    run<1>("distinct", buf, 64, 1, it, st);
    run<1>("distinct", buf, 128, 1, it, st);
    run<1>("distinct", buf, 148, 1, it, st);
  }
  run<1>("shared by 2", buf, 128, 2, it, 8);
  run<1>("shared by 4", buf, 128, 4, it, 8);
  run<1>("shared by 8", buf, 128, 8, it, 8);
  run<1>("shared by 32 (h-like)", buf, 128, 32, it, 8);
  run<1>("shared by 128", buf, 128, 128, it, 8);
  run<1>("shared by 148", buf, 148, 148, it, 8);
  for (int st : {4, 8, 12}) {
    run<2>("multicast 2, distinct clusters", buf, 128, 1, it, st);
    run<4>("multicast 4, distinct clusters", buf, 128, 1, it, st);
    run<8>("multicast 8, distinct clusters", buf, 128, 1, it, st);
  }
  run<2>("multicast 2, 16 clusters share", buf, 128, 16, it, 8);
  run<4>("multicast 4, 8 clusters share", buf, 128, 8, it, 8);
  run<8>("multicast 8, 4 clusters share", buf, 128, 4, it, 8);
  run<16>("multicast 16, distinct", buf, 128, 1, it, 8);
  return 0;
}
