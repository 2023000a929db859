// Experiment (design check for the fused ResnetBlock kernel): may the start address of a K-major SWIZZLE_128B
// shared-memory operand descriptor be advanced by an ARBITRARY number of 128-byte rows?  If the tensor core applies
// the swizzle XOR to the absolute shared-memory address (as the usual +32-byte K advance suggests), one window of
// 128 + 2d rows serves all three taps of a dilated k3 convolution: tap j reads rows [j*d, j*d + 128).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o desc_shift desc_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../autoformer_b200/csrc/avc_ptx.cuh"

using namespace avc;

constexpr int kRows = 160;

__global__ void __launch_bounds__(128, 1) shift_kernel(int d, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* a = base;                       // kRows x 128 B, swizzled
  uint8_t* b = base + kRows * 128;         // 32 x 128 B, swizzled (kRows * 128 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(b + 32 * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < kRows * 64; i += blockDim.x) {
    const int R = i / 64, k = i % 64;
    const float v = (float)((R * 7 + k * 3) % 61);
    *reinterpret_cast<__nv_bfloat16*>(a + R * 128 + (((k >> 3) ^ (R & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
    const int n = i / 64, k = i % 64;
    const float v = (k % 32 == n) ? (k < 32 ? 1.f : 2.f) : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(b + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(v);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) {
    tmem_alloc(tptr, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tptr);
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc(128, 32, false);
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem, umma_desc_sw128(smem_u32(a) + d * 128 + k * 32), umma_desc_sw128(smem_u32(b) + k * 32), idesc,
                k > 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[32];
  const int warp = threadIdx.x >> 5;
  tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int n = 0; n < 32; ++n) out[threadIdx.x * 32 + n] = __uint_as_float(v[n]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 32 * 4);
  std::vector<float> h(128 * 32);
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int bad_total = 0;
  for (int d : {0, 1, 2, 3, 5, 8, 9, 18, 27}) {
    shift_kernel<<<1, 128, 64 * 1024>>>(d, out);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("launch failed for d=%d: %s\n", d, cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 32; ++n) {
        const int R = r + d;
        const float want = (float)((R * 7 + n * 3) % 61) + 2.f * (float)((R * 7 + (n + 32) * 3) % 61);
        if (h[r * 32 + n] != want) ++bad;
      }
    printf("row shift d=%2d: %d mismatches of %d\n", d, bad, 128 * 32);
    bad_total += bad;
  }
  printf(bad_total ? "descriptor row shift: NOT usable\n" : "descriptor row shift: exact for every d\n");
  return 0;
}
