// A whole stack of small-batch LSTM layers as ONE wavefront (see include/avc_b200.h: avc_lstm_stack_ws).
//
// avc_lstm_seq_ws runs one layer: a frame is a chain of L2 round trips (grid barrier, h slice, reduction) of about
// 4.3 us whatever the batch is, and the layers of a stack run one after another -- LstmDV (factory/LstmDV.py:12,20:
// 3 x LSTM(768), 1000 frames) spends 3 x 1000 such frames.  Layer l at frame t needs only layer l-1 at frame t and its own
// frame t-1, so the layers can run ONE TICK APART: at tick k layer l works on frame k - l, all layers at once, on
// disjoint SMs, and the stack costs T + L - 1 ticks instead of L x T frames.
//
// For that the input projection of the layers above the first has to move into the recurrence (a dense projection in
// front needs the whole sequence of the layer below), and every weight has to stay on chip for the whole sequence:
//   layer 0:   z = xproj_0[t] (dense GEMM in front, as before)     + W_hh0 h0_{t-1}
//   layer l>0: z = bias_l + W_ih_l h^{l-1}_t                       + W_hh_l h^l_{t-1}
// The grid is L layers x R = 4H/128 row blocks x 2 CTAs, each pair a thread-block cluster.  In layer 0 the two CTAs split
// the K range of W_hh0 (two fp16 terms each, as in "fp16x2"); above, CTA 0 holds the row block of W_ih and CTA 1 that of
// W_hh as ONE fp16 term each -- that is what fits: 128 rows x 768 channels x 2 bytes = 384 of the 512 tensor-memory
// columns per CTA beside the accumulator, 144 CTAs for LstmDV.  (Two terms for every matrix would need 47 MB on chip.
// The single-term layers cost precision: embedding rel-L2 2.6e-4 instead of 1.4e-4 on the test weights, 5.5e-4 instead of
// 2.3e-4 on the x3-gain stress weights, gate 1e-3 -- scripts/lstm_stack_precision.py; the caller chooses.)
// The weights are the A operand of tcgen05.mma read from tensor memory; per tick a CTA TMA-loads its operand rows
// (h^{l-1}_t or h^l_{t-1}: B rows x K channels, in three groups so the MMAs start on the first), issues K/16 MMAs of
// width N = B', and the pair reduces its two partial sums through distributed shared memory exactly like
// avc_lstm_seq_ws (st.async crediting the owner's mbarrier); each CTA finalises 16 hidden units: cell update with c in
// registers, h_t written as fp16 into the stack's scratch sequence hs[l][t+1] (frame 0 is the zero initial state).
// One release/acquire grid barrier per tick orders the h stores of tick k before the TMA reads of tick k + 1.
//
// Warp roles: 0 = grid barrier + operand producer, 1 = MMA issuer, 2..5 = reduction + cell (warps 2, 3 store h_t).
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

#include "../../include/avc_b200.h"
#include "avc_host.h"
#include "avc_pipe.cuh"

namespace avc {

constexpr int kStThreads = 192;
constexpr int kStCellThreads = 128;
constexpr int kStMaxLayers = AVC_STACK_MAX_LAYERS;
constexpr int kStRowsOwn = kBlockM / 2;        // gate rows each CTA of the pair finalises
constexpr int kStUnitsOwn = kStRowsOwn / 4;    // = 16 hidden units
constexpr int kStGroups = 3;                   // operand-load groups per tick (one mbarrier each)

struct alignas(64) StackParams {
  // hs[l] as (64 channels, B rows, H/64 chunks, T + 1 frames), box {64, AR, chunks of one load group, 1}: one TMA
  // operation brings a whole group of 64-channel operand tiles (rows >= B and chunks >= H/64 are zero-filled)
  CUtensorMap tmap_h[kStMaxLayers];   // groups of the layers above the first (ceil(H/64 / 3) chunks)
  CUtensorMap tmap_h0;                // hs[0] in the groups of layer 0's own CTAs (ceil(H/128 / 3) chunks)
  const __half* w_hh0;                // [4H][2H] = [hi | lo]
  const __half* w_ih[kStMaxLayers];   // l >= 1: [4H][H]
  const __half* w_hh[kStMaxLayers];   // l >= 1: [4H][H]
  const float* bias[kStMaxLayers];    // l >= 1: [4H]
  const float* xproj0;
  __half* hs;
  float* h_last;
  unsigned int* grid_barrier;
  long long* debug_clk;               // optional: 16 clock64 stamps per (tick, CTA)
  int B, T, H, L;
};

__device__ __forceinline__ uint32_t st_map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f4(uint32_t addr, uint32_t mbar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(mbar)
               : "memory");
}

template <int AR>
struct StackCfg {
  static constexpr int NQ = AR / 4;                                   // groups of 4 utterances
  static constexpr int kItems = kStUnitsOwn * NQ;                     // (unit, utterance group) pairs an owner finalises
  static constexpr int IT = (kItems + kStCellThreads - 1) / kStCellThreads;
  static constexpr int kHTile = AR * kRowBytes;                       // one 64-channel chunk of the operand rows
  static constexpr int kRedFrameBytes = 2 * AR * kStRowsOwn * 4;      // what one tick pushes into an owner
  static constexpr int kRedBytes = (kRedFrameBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = (AR * kStUnitsOwn * 4 + 1023) / 1024 * 1024;
  static constexpr uint32_t kAccCols = AR < 32 ? 32 : AR;
  // operand tiles: whole load groups (the last group of a K range that does not divide by 3 is zero-filled)
  __host__ __device__ static int h_tiles(int H) { return kStGroups * ((H / 64 + kStGroups - 1) / kStGroups); }
  static int smem_bytes(int H) { return h_tiles(H) * kHTile + kRedBytes + kStageBytes + 128 + 1024 /* alignment slack */; }
};

template <int AR>
__global__ void __launch_bounds__(kStThreads, 1) lstm_stack_kernel(const __grid_constant__ StackParams p) {
  using Cfg = StackCfg<AR>;
  constexpr int NQ = Cfg::NQ, IT = Cfg::IT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_h = base;
  float* s_red = reinterpret_cast<float*>(s_h + Cfg::h_tiles(p.H) * Cfg::kHTile);
  float* s_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_red) + Cfg::kRedBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_stage) + Cfg::kStageBytes);
  uint64_t* w_full = bars + 0;
  uint64_t* d_full = bars + 1;        // the frame's MMAs have completed (accumulator ready)
  uint64_t* red_full = bars + 2;      // both partial sums of the rows this CTA owns have landed
  uint64_t* epi_done = bars + 3;      // this CTA's h_t is stored (and the accumulator drained)
  uint64_t* h_full = bars + 4;        // [kStGroups]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int R = 4 * p.H / kBlockM;
  const int cid = blockIdx.x >> 1;
  const int layer = cid / R;
  const int r = cid % R;
  const unsigned int n_ctas = gridDim.x;
  long long* const clk = p.debug_clk;
#define AVC_ST_STAMP(k_, i_) \
  if (clk) clk[((long long)(k_) * n_ctas + blockIdx.x) * 16 + (i_)] = clock64()

  // what this CTA multiplies: `chunks` 64-channel chunks starting at channel k0 of sequence `src`, frame t + frame_off
  int chunks, terms, k0, src, frame_off;
  const __half* w;
  long long w_ld;
  if (layer == 0) {
    chunks = p.H / 128; terms = 2; k0 = (int)rank * (p.H / 2); src = 0; frame_off = 0; w = p.w_hh0; w_ld = 2LL * p.H;
  } else if (rank == 0) {
    chunks = p.H / 64; terms = 1; k0 = 0; src = layer - 1; frame_off = 1; w = p.w_ih[layer]; w_ld = p.H;
  } else {
    chunks = p.H / 64; terms = 1; k0 = 0; src = layer; frame_off = 0; w = p.w_hh[layer]; w_ld = p.H;
  }
  const int cpg = (chunks + kStGroups - 1) / kStGroups;          // chunks per load group
  const int ngroups = (chunks + cpg - 1) / cpg;
  const CUtensorMap* const tmap = layer == 0 ? &p.tmap_h0 : &p.tmap_h[src];

  if (threadIdx.x == 0) {
    mbar_init(w_full, kStCellThreads / 32);
    mbar_init(d_full, 1);
    mbar_init(red_full, 1);
    mbar_init(epi_done, AR > 32 ? 2 : 1);     // one arrival per warp that stores h_t
    for (int g = 0; g < kStGroups; ++g) mbar_init(h_full + g, 1);
    fence_mbar_init();
    prefetch_tmap(tmap);
  }
  constexpr uint32_t tmem_cols = 512;
  if (warp == 1) {
    tmem_alloc(tmem_ptr, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();                 // the peer credits bytes to this CTA's red_full: its init must be visible cluster-wide
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);
  const uint32_t tmem_w = tmem_base + Cfg::kAccCols;        // term 0 at columns [0, 32 chunks), term 1 behind it
  const int nticks = p.T + p.L - 1;

  if (warp == 0) {
    if (lane == 0) {
      for (int k = 0; k < nticks; ++k) {
        const int t = k - layer;
        if (k > 0) {
          // this CTA's part of the previous tick's h is stored, then every CTA's
          if (t - 1 >= 0 && t - 1 < p.T) mbar_wait(epi_done, (t - 1) & 1);
          grid_arrive_wait(p.grid_barrier, (unsigned)k * n_ctas,
                           clk ? clk + ((long long)k * n_ctas + blockIdx.x) * 16 : nullptr);  // stamp 0: arrival issued
        }
        AVC_ST_STAMP(k, 1);
        if (t >= 0 && t < p.T) {
          fence_proxy_async_global();   // the operand rows were written with generic stores
          for (int g = 0; g < ngroups; ++g) {
            mbar_arrive_expect_tx(h_full + g, cpg * Cfg::kHTile);          // the whole box, zero-filled part included
            tma_load_4d(s_h + g * cpg * Cfg::kHTile, tmap, h_full + g, 0, 0, k0 / 64 + g * cpg, t + frame_off);
          }
          if (clk)                       // profiling only: when each group really landed (the MMA thread sees a group
            for (int g = 0; g < ngroups; ++g) {   // only after it has issued the previous group's MMAs)
              mbar_wait(h_full + g, t & 1);
              AVC_ST_STAMP(k, 11 + g);
            }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBlockM, AR, false) ^ kIdescF16Xor;
      mbar_wait(w_full, 0);
      tc_fence_after();
      for (int t = 0; t < p.T; ++t) {
        // (the accumulator is free: the operands of frame t exist only after this CTA's cell warps drained frame t - 1)
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait(h_full + g, t & 1);
          AVC_ST_STAMP(t + layer, g == 0 ? 2 : 7 + g);              // 2, 8, 9: operand group g landed
          tc_fence_after();             // also orders the cell warps' accumulator reads of frame t - 1 before these MMAs
          const int c0 = g * cpg, c1 = min(chunks, c0 + cpg);
          for (int c = c0; c < c1; ++c) {
            const uint32_t h = smem_u32(s_h + c * Cfg::kHTile);
            for (int term = 0; term < terms; ++term) {
              const uint32_t wt = tmem_w + (term * chunks + c) * 32;       // 8 columns per K = 16
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16_ts(tmem_base, wt + kk * 8, umma_desc_sw128(h + kk * 32), idesc,
                             (c == 0 && term == 0 && kk == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(d_full);
        AVC_ST_STAMP(t + layer, 10);                                  // all MMAs of the frame issued
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                           // TMEM lane quarter
    const int row = q * 32 + lane;                    // accumulator row = packed gate row within the row block
    const int ctid = threadIdx.x - 64;                // 0..127
    // reduction buffer of an owner: [source rank][group of 4 utterances][gate][owned unit][4] fp32 -- a warp's 16-byte
    // stores of one group fill one contiguous 512-byte run, and the cell threads of adjacent units read adjacent float4
    const uint32_t owner = row / kStRowsOwn;
    const int slot = (row & 3) * kStUnitsOwn + (row % kStRowsOwn) / 4;
    const uint32_t push_addr = st_map_to_cta(smem_u32(s_red + ((int)rank * NQ * kStRowsOwn + slot) * 4), owner);
    const uint32_t owner_bar = st_map_to_cta(smem_u32(red_full), owner);
    {
      // this thread's W row -> its TMEM lane, once: 32 fp16 (16 columns) per store
      const uint32_t lane_base = tmem_w + (static_cast<uint32_t>(q * 32) << 16);
      for (int term = 0; term < terms; ++term) {
        const uint4* srcw =
            reinterpret_cast<const uint4*>(w + ((long long)r * kBlockM + row) * w_ld + (long long)term * p.H + k0);
        for (int g = 0; g < 2 * chunks; ++g) {
          uint32_t v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = __ldg(srcw + g * 4 + i);
            v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
          }
          tmem_st_32x16(lane_base + term * chunks * 32 + g * 16, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_full);
    }
    int u_own[IT], n4[IT];
    bool active[IT];
    float4 bias4[IT];
    float c_state[IT][4];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int e = it * kStCellThreads + ctid;       // adjacent threads: adjacent units (contiguous reads)
      u_own[it] = e % kStUnitsOwn;
      n4[it] = e / kStUnitsOwn;
      active[it] = n4[it] < NQ;
      bias4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (layer > 0)
        bias4[it] = __ldg(reinterpret_cast<const float4*>(p.bias[layer] + r * kBlockM +
                                                          4 * ((int)rank * kStUnitsOwn + u_own[it])));
#pragma unroll
      for (int j = 0; j < 4; ++j) c_state[it][j] = 0.f;
    }
    const long long H4 = 4LL * p.H;
    __half* const hs_l = p.hs + (long long)layer * (p.T + 1) * p.B * p.H;
    const bool top = layer == p.L - 1;
    for (int t = 0; t < p.T; ++t) {
      // the additive term of this thread's cells: issued first, consumed after the reduction
      float z[IT][4][4];                              // [item][utterance j][gate]
#pragma unroll
      for (int it = 0; it < IT; ++it)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 xp = bias4[it];
          const int n = n4[it] * 4 + j;
          if (layer == 0 && active[it] && n < p.B)
            xp = __ldg(reinterpret_cast<const float4*>(p.xproj0 + ((long long)n * p.T + t) * H4 + r * kBlockM +
                                                       4 * ((int)rank * kStUnitsOwn + u_own[it])));
          z[it][j][0] = xp.x; z[it][j][1] = xp.y; z[it][j][2] = xp.z; z[it][j][3] = xp.w;
        }
      if (threadIdx.x == 64) mbar_arrive_expect_tx(red_full, Cfg::kRedFrameBytes);
      mbar_wait(d_full, t & 1);
      if (threadIdx.x == 64) AVC_ST_STAMP(t + layer, 3);
      tc_fence_after();
      // this thread's accumulator row, pushed to the CTA that owns it
      {
        uint32_t a[AR / 16][16];
#pragma unroll
        for (int j = 0; j < AR / 16; ++j) tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 16, a[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < AR / 16; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_async_f4(push_addr + (j * 4 + i) * kStRowsOwn * 16, owner_bar, __uint_as_float(a[j][i * 4]),
                        __uint_as_float(a[j][i * 4 + 1]), __uint_as_float(a[j][i * 4 + 2]), __uint_as_float(a[j][i * 4 + 3]));
      }
      tc_fence_before();                // accumulator reads before the next frame's MMAs (via epi_done -> h_full)
      if (threadIdx.x == 64) AVC_ST_STAMP(t + layer, 4);
      mbar_wait_cluster(red_full, t & 1);
      if (threadIdx.x == 64) AVC_ST_STAMP(t + layer, 5);
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        if (!active[it]) continue;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 v = *reinterpret_cast<const float4*>(
                s_red + ((s * NQ + n4[it]) * kStRowsOwn + g * kStUnitsOwn + u_own[it]) * 4);   // gate g, 4 utterances
            z[it][0][g] += v.x; z[it][1][g] += v.y; z[it][2][g] += v.z; z[it][3][g] += v.w;
          }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float cn, hn;
          lstm_cell(z[it][j][0], z[it][j][1], z[it][j][2], z[it][j][3], c_state[it][j], cn, hn);
          c_state[it][j] = cn;
          s_stage[(n4[it] * 4 + j) * kStUnitsOwn + u_own[it]] = hn;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStCellThreads) : "memory");
      if (threadIdx.x == 64) AVC_ST_STAMP(t + layer, 6);
      if (warp == 2 || (AR > 32 && warp == 3)) {
        // utterance n: the 16 units this CTA finalised = 32 bytes of fp16 in frame t + 1 of this layer's sequence
        const int n = (warp - 2) * 32 + lane;
        if (n < AR && n < p.B) {
          const int ug = r * 32 + (int)rank * kStUnitsOwn;
          const float4* sv = reinterpret_cast<const float4*>(s_stage + n * kStUnitsOwn);
          const float4 h0 = sv[0], h1 = sv[1], h2 = sv[2], h3 = sv[3];
          uint4* o = reinterpret_cast<uint4*>(hs_l + ((long long)(t + 1) * p.B + n) * p.H + ug);
          o[0] = make_uint4(pack_f16(h0.x, h0.y), pack_f16(h0.z, h0.w), pack_f16(h1.x, h1.y), pack_f16(h1.z, h1.w));
          o[1] = make_uint4(pack_f16(h2.x, h2.y), pack_f16(h2.z, h2.w), pack_f16(h3.x, h3.y), pack_f16(h3.z, h3.w));
          if (top && t == p.T - 1 && p.h_last) {
            float4* hl = reinterpret_cast<float4*>(p.h_last + (long long)n * p.H + ug);
            hl[0] = h0; hl[1] = h1; hl[2] = h2; hl[3] = h3;
          }
          fence_proxy_async_global();   // order the h stores before later async-proxy (TMA) reads
        }
        __syncwarp();
        if (lane == 0) {
          AVC_ST_STAMP(t + layer, 7);
          mbar_arrive(epi_done);        // after bar.sync 1: every cell warp has drained the accumulator too
        }
      }
    }
  }
#undef AVC_ST_STAMP
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                   // no CTA may exit while its peer can still push into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

template <int AR>
static int launch_stack(const StackParams& p, const avc_lstm_stack_desc* d, cudaStream_t stream) {
  using Cfg = StackCfg<AR>;
  auto kern = lstm_stack_kernel<AR>;
  const int smem = Cfg::smem_bytes(d->H);
  AVC_REQUIRE(smem <= 227 * 1024, "avc_lstm_stack_ws: H=%d B=%d needs %d bytes of shared memory", d->H, d->B, smem);
  static int configured[PerDeviceOnce::kMaxDevices] = {};   // largest size set so far, per device
  int dev = 0;
  AVC_CHECK_CUDA(cudaGetDevice(&dev));
  const bool known = dev >= 0 && dev < PerDeviceOnce::kMaxDevices;
  if (!known || configured[dev] < smem) {
    AVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (known) configured[dev] = smem;
  }
  const int grid = d->L * (4 * d->H / kBlockM) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kStThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  static const bool no_coop = getenv("AVC_LSTM_NO_COOP") != nullptr;   // profiling aid, see avc_lstm.cu
  if (!no_coop) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  int clusters = 0;
  AVC_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg));
  if (clusters * 2 < grid) {
    set_error("avc_lstm_stack_ws: grid %d does not fit (%d CTAs resident)", grid, clusters * 2);
    return AVC_ERR_NOT_RESIDENT;
  }
  AVC_CHECK_CUDA(cudaMemsetAsync(d->grid_barrier, 0, sizeof(unsigned int), stream));
  // frame 0 of every layer's sequence: the zero initial state
  const size_t frame_bytes = (size_t)d->B * d->H * 2, layer_bytes = frame_bytes * (size_t)(d->T + 1);
  for (int l = 0; l < d->L; ++l)
    AVC_CHECK_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(d->hs) + l * layer_bytes, 0, frame_bytes, stream));
  AVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  return 0;
}

}  // namespace avc

extern "C" int avc_lstm_stack_ws(const avc_lstm_stack_desc* d, void* stream_v) {
  using namespace avc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  AVC_REQUIRE(d != nullptr, "avc_lstm_stack_ws: null descriptor");
  AVC_REQUIRE(d->B > 0 && d->B <= 64 && d->T > 0, "avc_lstm_stack_ws: B=%d T=%d (1 <= B <= 64)", d->B, d->T);
  AVC_REQUIRE(d->L >= 1 && d->L <= kStMaxLayers, "avc_lstm_stack_ws: L=%d (1..%d layers)", d->L, kStMaxLayers);
  AVC_REQUIRE(d->H > 0 && d->H % 128 == 0 && d->H / 2 + 64 <= 512,
              "avc_lstm_stack_ws: unsupported H=%d (multiple of 128, at most 896: the weights of a CTA fill H/2 "
              "tensor-memory columns beside the accumulator)", d->H);
  AVC_REQUIRE(d->xproj0 && d->w_hh0 && d->hs && d->grid_barrier, "avc_lstm_stack_ws: missing buffer");
  for (int l = 1; l < d->L; ++l)
    AVC_REQUIRE(d->w_ih[l] && d->w_hh[l] && d->bias[l], "avc_lstm_stack_ws: missing weights of layer %d", l);
  if (d->L * (4 * d->H / kBlockM) * 2 > num_sms()) {
    set_error("avc_lstm_stack_ws: %d layers x %d row blocks x 2 CTAs exceed the %d SMs", d->L, 4 * d->H / kBlockM, num_sms());
    return AVC_ERR_NOT_RESIDENT;
  }
  const uint64_t H = (uint64_t)d->H;
  const int ar = d->B <= 16 ? 16 : (d->B <= 32 ? 32 : 64);
  StackParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t frame_bytes = (uint64_t)d->B * H * 2, layer_bytes = frame_bytes * (uint64_t)(d->T + 1);
  const uint32_t cpg_upper = (uint32_t)((d->H / 64 + kStGroups - 1) / kStGroups);
  const uint32_t cpg_first = (uint32_t)((d->H / 128 + kStGroups - 1) / kStGroups);
  const uint64_t dh[4] = {64, (uint64_t)d->B, H / 64, (uint64_t)d->T + 1};
  const uint64_t sh[3] = {H * 2, 128, frame_bytes};
  const uint32_t box_first[4] = {64, (uint32_t)ar, cpg_first, 1};
  if (!encode_tmap_4d(&p.tmap_h0, 2, d->hs, dh, sh, box_first)) return -3;
  for (int l = 0; l < d->L; ++l) {
    const uint32_t box[4] = {64, (uint32_t)ar, cpg_upper, 1};
    if (!encode_tmap_4d(&p.tmap_h[l], 2, static_cast<uint8_t*>(d->hs) + l * layer_bytes, dh, sh, box)) return -3;
    p.w_ih[l] = static_cast<const __half*>(d->w_ih[l]);
    p.w_hh[l] = static_cast<const __half*>(d->w_hh[l]);
    p.bias[l] = d->bias[l];
  }
  p.w_hh0 = static_cast<const __half*>(d->w_hh0);
  p.xproj0 = d->xproj0;
  p.hs = static_cast<__half*>(d->hs);
  p.h_last = d->h_last;
  p.grid_barrier = d->grid_barrier;
  p.debug_clk = d->debug_clk;
  p.B = d->B;
  p.T = d->T;
  p.H = d->H;
  p.L = d->L;
  return ar == 16 ? launch_stack<16>(p, d, stream) : ar == 32 ? launch_stack<32>(p, d, stream) : launch_stack<64>(p, d, stream);
}
