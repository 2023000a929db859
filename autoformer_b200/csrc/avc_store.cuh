// Element / 4-element stores into the operand formats: 0 fp32 [TF32-rounded], 1 bf16, 2 split bf16 [hi | lo], 3 fp16,
// 4 split fp16 [hi | lo] (two fp16 terms, ~2^-22: the MelGAN residual stream of the "fp16s" precision).
#pragma once
#include <cuda_bf16.h>

#include "avc_pipe.cuh"

namespace avc {

// One logical channel c of a row with C logical channels; ld = elements per row of the buffer.
__device__ __forceinline__ void store_op1(void* out, int mode, int round, long long row, long long ld, int c, int C,
                                          float v) {
  if (mode == 2) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16* p = static_cast<__nv_bfloat16*>(out) + row * ld + c;
    p[0] = hi;
    p[C] = __float2bfloat16_rn(v - __bfloat162float(hi));
  } else if (mode == 4) {
    const __half hi = __float2half_rn(sat_f16(v));
    __half* p = static_cast<__half*>(out) + row * ld + c;
    p[0] = hi;
    p[C] = __float2half_rn(sat_f16(v) - __half2float(hi));
  } else if (mode == 1) {
    static_cast<__nv_bfloat16*>(out)[row * ld + c] = __float2bfloat16_rn(v);
  } else if (mode == 3) {
    static_cast<__half*>(out)[row * ld + c] = __float2half_rn(sat_f16(v));
  } else {
    static_cast<float*>(out)[row * ld + c] = round ? round_tf32(v) : v;
  }
}

// Four consecutive logical channels [c, c+4), c % 4 == 0.
__device__ __forceinline__ void store_op4(void* out, int mode, int round, long long row, long long ld, int c, int C,
                                          float4 v) {
  if (mode == 2) {
    __nv_bfloat16* p = static_cast<__nv_bfloat16*>(out) + row * ld + c;
    const float lx = v.x - __bfloat162float(__float2bfloat16_rn(v.x));
    const float ly = v.y - __bfloat162float(__float2bfloat16_rn(v.y));
    const float lz = v.z - __bfloat162float(__float2bfloat16_rn(v.z));
    const float lw = v.w - __bfloat162float(__float2bfloat16_rn(v.w));
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    *reinterpret_cast<uint2*>(p + C) = make_uint2(pack_bf16(lx, ly), pack_bf16(lz, lw));
  } else if (mode == 4) {
    __half* p = static_cast<__half*>(out) + row * ld + c;
    uint2 hi, lo;
    split_f16_pair(v.x, v.y, hi.x, lo.x);
    split_f16_pair(v.z, v.w, hi.y, lo.y);
    *reinterpret_cast<uint2*>(p) = hi;
    *reinterpret_cast<uint2*>(p + C) = lo;
  } else if (mode == 1) {
    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + row * ld + c) =
        make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  } else if (mode == 3) {
    *reinterpret_cast<uint2*>(static_cast<__half*>(out) + row * ld + c) =
        make_uint2(pack_f16(v.x, v.y), pack_f16(v.z, v.w));
  } else {
    if (round) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    *reinterpret_cast<float4*>(static_cast<float*>(out) + row * ld + c) = v;
  }
}

__host__ __device__ inline long long op_ld(int C, int mode) { return (mode == 2 || mode == 4) ? 2LL * C : C; }

}  // namespace avc
