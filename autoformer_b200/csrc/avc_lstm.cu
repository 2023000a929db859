// Large-H LSTM recurrence on tcgen05 (see include/avc_b200.h: avc_lstm_seq).
//
// One time step is the GEMM  Z[B x 4H] = h_{t-1}[B x H] . W_hh^T  with the LSTM cell as its epilogue.
// W_hh rows (and the xproj columns) are gate-interleaved in groups of G hidden units, so a 128 x 4G
// accumulator tile holds all four gates of G units for 128 utterances: the epilogue thread that owns
// accumulator row b updates c[b, u] / h[b, u] for those G units without any cross-thread traffic.
// h_{t-1} is read by TMA straight out of the output sequence [B][T][H] (a 3-D box {128 B, 1 frame,
// 128 utterances} at frame t-1), so there is no separate recurrent-state buffer.
//
// Two launch modes share the kernel:
//   per-step    t_end = t_begin + 1, one launch per frame (stream order is the time dependency)
//   persistent  one cooperative launch for all frames; a grid-wide barrier separates frames
#include <cuda_bf16.h>

#include "../../include/avc_b200.h"
#include "avc_host.h"
#include "avc_pipe.cuh"

namespace avc {

struct alignas(64) LstmParams {
  CUtensorMap tmap_h;   // hseq as (H, T, B), box {kc, 1, 128}
  CUtensorMap tmap_w;   // w_hh as (H, 4H), box {kc, BN}
  const float* xproj;
  void* hseq;
  float* hseq_f32;
  float* h_last;
  float* c_state;
  unsigned int* grid_barrier;
  int B, T, H;
  int num_kb, kc_elems;
  int a_wrap;           // A channel coordinate = (kb * kc) % a_wrap  (split bf16: [h_hi|h_lo] then h_hi again)
  int n_tiles;
  int t_begin, t_end;
};

// All threads of all CTAs call this; `target` = number of arrivals expected so far (monotonic counter).
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (seen < target && clock64() - t0 > 4000000000LL) {
        printf("avc: grid barrier timeout block %d seen %u target %u\n", (int)blockIdx.x, seen, target);
        __trap();
      }
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// MODE: 0 = tf32, 1 = bf16, 2 = split bf16 (three bf16 products per fp32 product)
template <int BN, int MODE>
__global__ void __launch_bounds__(kNumThreads, 1) lstm_step_kernel(const __grid_constant__ LstmParams p) {
  using C = PipeCfg<BN>;
  constexpr int G = BN / 4;
  constexpr bool BF16 = MODE != 0;
  extern __shared__ uint8_t smem_raw[];
  const PipeSmem s = carve_smem<BN>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int b0 = m_tile * kBlockM;
  const int n0 = n_tile * BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmap_h);
    prefetch_tmap(&p.tmap_w);
  }
  const uint32_t tmem_base = pipe_setup<BN>(s);

  RingState rs;              // producer and MMA issuer each keep their own copy (same sequence)
  uint32_t acc_phase = 0;    // epilogue: parity of tmem_full
  unsigned int sync_count = 0;

  for (int t = p.t_begin; t < p.t_end; ++t) {
    const bool has_mma = t > 0;   // h_{-1} = 0: the first frame has no recurrent term
    if (warp == 0) {
      if (lane == 0 && has_mma) {
        fence_proxy_async_all();   // h_{t-1} was written with generic stores (other CTAs / previous launch)
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&s.empty[rs.stage], rs.phase ^ 1u);
          uint8_t* a_dst = s.base + rs.stage * C::kStageBytes;
          mbar_arrive_expect_tx(&s.full[rs.stage], C::kStageBytes);
          tma_load_3d(a_dst, &p.tmap_h, &s.full[rs.stage], (kb * p.kc_elems) % p.a_wrap, t - 1, b0);
          tma_load_2d(a_dst + kATileBytes, &p.tmap_w, &s.full[rs.stage], kb * p.kc_elems, n0);
          rs.advance<C::kStages>();
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0 && has_mma) {
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&s.full[rs.stage], rs.phase);
          tc_fence_after();
          issue_kblock<BN, BF16>(s, rs.stage, tmem_base, kb == 0);
          umma_commit(&s.empty[rs.stage]);
          rs.advance<C::kStages>();
        }
        umma_commit(s.tmem_full);
      }
      __syncwarp();
    } else {
      // ---------------- cell epilogue: thread owns utterance b; the two epilogue warps of a TMEM lane quarter
      // (warp_id % 4) take alternate groups of 8 hidden units of this tile
      const int q = warp & 3;
      const int half = (warp - 2) >> 2;
      const int b = b0 + q * 32 + lane;
      const bool valid = b < p.B;
      const int u0 = n_tile * G;
      const long long row = (long long)b * p.T + t;
      const float* xp = p.xproj + row * (4LL * p.H) + n0;
      float* cp = p.c_state + (long long)b * p.H + u0;
      if (has_mma) {
        mbar_wait(s.tmem_full, acc_phase);
        acc_phase ^= 1u;
        tc_fence_after();
      }
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int j = half; j < G / 8; j += kEpiWarps / 4) {
        uint32_t acc[4][8];
        if (has_mma) {
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld_32x8(lane_addr + g * G + j * 8, acc[g]);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[g][e] = 0u;
        }
        if (valid) {
          float z[4][8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(xp + g * G + j * 8));
            const float4 x1 = __ldg(reinterpret_cast<const float4*>(xp + g * G + j * 8 + 4));
            z[g][0] = __uint_as_float(acc[g][0]) + x0.x;
            z[g][1] = __uint_as_float(acc[g][1]) + x0.y;
            z[g][2] = __uint_as_float(acc[g][2]) + x0.z;
            z[g][3] = __uint_as_float(acc[g][3]) + x0.w;
            z[g][4] = __uint_as_float(acc[g][4]) + x1.x;
            z[g][5] = __uint_as_float(acc[g][5]) + x1.y;
            z[g][6] = __uint_as_float(acc[g][6]) + x1.z;
            z[g][7] = __uint_as_float(acc[g][7]) + x1.w;
          }
          float cprev[8];
          if (has_mma) {
            const float4 c0 = *reinterpret_cast<const float4*>(cp + j * 8);
            const float4 c1 = *reinterpret_cast<const float4*>(cp + j * 8 + 4);
            cprev[0] = c0.x; cprev[1] = c0.y; cprev[2] = c0.z; cprev[3] = c0.w;
            cprev[4] = c1.x; cprev[5] = c1.y; cprev[6] = c1.z; cprev[7] = c1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) cprev[e] = 0.0f;
          }
          float cn[8], hn[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float ig = sigmoid_fast(z[0][e]);
            const float fg = sigmoid_fast(z[1][e]);
            const float gg = tanh_fast(z[2][e]);
            const float og = sigmoid_fast(z[3][e]);
            cn[e] = fg * cprev[e] + ig * gg;
            hn[e] = og * tanh_fast(cn[e]);
          }
          *reinterpret_cast<float4*>(cp + j * 8) = make_float4(cn[0], cn[1], cn[2], cn[3]);
          *reinterpret_cast<float4*>(cp + j * 8 + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
          const long long hoff = row * p.H + u0 + j * 8;
          if (MODE == 2) {
            float lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) lo[e] = hn[e] - __bfloat162float(__float2bfloat16_rn(hn[e]));
            __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(p.hseq) + row * (2LL * p.H) + u0 + j * 8;
            *reinterpret_cast<uint4*>(hp) = make_uint4(pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]),
                                                       pack_bf16(hn[4], hn[5]), pack_bf16(hn[6], hn[7]));
            *reinterpret_cast<uint4*>(hp + p.H) = make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]),
                                                             pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
          } else if (MODE == 1) {
            uint4 pk = make_uint4(pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]), pack_bf16(hn[4], hn[5]),
                                  pack_bf16(hn[6], hn[7]));
            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.hseq) + hoff) = pk;
          } else {
            float* hp = static_cast<float*>(p.hseq) + hoff;
            *reinterpret_cast<float4*>(hp) =
                make_float4(round_tf32(hn[0]), round_tf32(hn[1]), round_tf32(hn[2]), round_tf32(hn[3]));
            *reinterpret_cast<float4*>(hp + 4) =
                make_float4(round_tf32(hn[4]), round_tf32(hn[5]), round_tf32(hn[6]), round_tf32(hn[7]));
          }
          if (p.hseq_f32) {
            float* hp = p.hseq_f32 + hoff;
            *reinterpret_cast<float4*>(hp) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            *reinterpret_cast<float4*>(hp + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
          }
          if (p.h_last && t == p.T - 1) {
            float* hp = p.h_last + (long long)b * p.H + u0 + j * 8;
            *reinterpret_cast<float4*>(hp) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            *reinterpret_cast<float4*>(hp + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
          }
        }
      }
      fence_proxy_async_all();   // order the h stores before later async-proxy (TMA) reads
    }
    if (t + 1 < p.t_end) {
      // the next frame's MMAs overwrite the accumulator and read h_t from every CTA: grid-wide barrier
      tc_fence_before();
      ++sync_count;
      grid_sync(p.grid_barrier, sync_count * gridDim.x);
      tc_fence_after();
    }
  }
  pipe_teardown<BN>(tmem_base);
}

template <int BN, int MODE>
static int run(LstmParams p, const avc_lstm_desc* d, cudaStream_t stream) {
  auto kern = lstm_step_kernel<BN, MODE>;
  static bool configured = false;
  if (!configured) {
    AVC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PipeCfg<BN>::kSmemBytes));
    configured = true;
  }
  const int m_tiles = (d->B + kBlockM - 1) / kBlockM;
  const int grid = m_tiles * p.n_tiles;
  if (d->persistent) {
    AVC_REQUIRE(d->grid_barrier != nullptr, "avc_lstm_seq: persistent mode needs grid_barrier scratch");
    int per_sm = 0;
    AVC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kNumThreads, PipeCfg<BN>::kSmemBytes));
    AVC_REQUIRE(per_sm * num_sms() >= grid, "avc_lstm_seq: persistent grid %d does not fit (%d x %d resident)", grid,
                per_sm, num_sms());
    AVC_CHECK_CUDA(cudaMemsetAsync(d->grid_barrier, 0, sizeof(unsigned int), stream));
    p.t_begin = 0;
    p.t_end = d->T;
    void* args[] = {&p};
    AVC_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kNumThreads), args,
                                               PipeCfg<BN>::kSmemBytes, stream));
    count_launch();
  } else {
    for (int t = 0; t < d->T; ++t) {
      p.t_begin = t;
      p.t_end = t + 1;
      kern<<<grid, kNumThreads, PipeCfg<BN>::kSmemBytes, stream>>>(p);
    }
    AVC_CHECK_CUDA(cudaGetLastError());
    count_launch(d->T);
  }
  return 0;
}

}  // namespace avc

extern "C" int avc_lstm_seq(const avc_lstm_desc* d, void* stream_v) {
  using namespace avc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  AVC_REQUIRE(d != nullptr, "avc_lstm_seq: null descriptor");
  AVC_REQUIRE(d->dtype >= AVC_DTYPE_TF32 && d->dtype <= AVC_DTYPE_BF16X3, "avc_lstm_seq: bad dtype %d", d->dtype);
  AVC_REQUIRE(d->gate_group == 16 || d->gate_group == 32 || d->gate_group == 64, "avc_lstm_seq: gate_group %d",
              d->gate_group);
  AVC_REQUIRE(d->B > 0 && d->T > 0 && d->H > 0 && d->H % d->gate_group == 0, "avc_lstm_seq: bad shape B=%d T=%d H=%d",
              d->B, d->T, d->H);
  AVC_REQUIRE(d->xproj && d->w_hh && d->hseq && d->c_state, "avc_lstm_seq: missing buffer");
  const int es = d->dtype == AVC_DTYPE_TF32 ? 4 : 2;
  const int kc = kRowBytes / es;
  AVC_REQUIRE(d->H % kc == 0, "avc_lstm_seq: H=%d must be a multiple of %d", d->H, kc);
  const int bn = 4 * d->gate_group;
  const bool split = d->dtype == AVC_DTYPE_BF16X3;
  const uint64_t hc = split ? 2ull * d->H : (uint64_t)d->H;   // channels of the h sequence buffer
  const uint64_t wk = split ? 3ull * d->H : (uint64_t)d->H;   // K extent of the packed recurrent weights

  LstmParams p;
  memset(&p, 0, sizeof(p));
  if (!encode_tmap_3d(&p.tmap_h, es, d->hseq, hc, (uint64_t)d->T, (uint64_t)d->B, hc * es, (uint64_t)d->T * hc * es, kc,
                      1, kBlockM))
    return -3;
  if (!encode_tmap_2d(&p.tmap_w, es, d->w_hh, wk, (uint64_t)4 * d->H, wk * es, kc, bn)) return -3;
  p.a_wrap = (int)hc;
  p.xproj = d->xproj;
  p.hseq = d->hseq;
  p.hseq_f32 = d->hseq_f32;
  p.h_last = d->h_last;
  p.c_state = d->c_state;
  p.grid_barrier = d->grid_barrier;
  p.B = d->B;
  p.T = d->T;
  p.H = d->H;
  p.kc_elems = kc;
  p.num_kb = (int)(wk / kc);
  p.n_tiles = 4 * d->H / bn;
#define AVC_LSTM_DISPATCH(BN_)                                  \
  switch (d->dtype) {                                           \
    case AVC_DTYPE_TF32: return run<BN_, 0>(p, d, stream);      \
    case AVC_DTYPE_BF16: return run<BN_, 1>(p, d, stream);      \
    default: return run<BN_, 2>(p, d, stream);                  \
  }
  switch (bn) {
    case 64: AVC_LSTM_DISPATCH(64)
    case 128: AVC_LSTM_DISPATCH(128)
    default: AVC_LSTM_DISPATCH(256)
  }
#undef AVC_LSTM_DISPATCH
}
