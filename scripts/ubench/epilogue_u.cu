// Micro-benchmark of the fused-MLP epilogue iteration (tcgen05.ld x32 + bias + bf16 pack + swizzled st.shared.v4)
// in isolation, with and without other warps spinning on an mbarrier, with bias from constant / global / none.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ float c_bias[4096];

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
  uint32_t r; asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}

// MODE 0: no bias, 1: constant memory, 2: global (__ldg)
template <int MODE>
__global__ void __launch_bounds__(576, 1) k(int iters, int work_warps, int spin_warps, const float* gbias, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t never;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&never)) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 200 * 1024);
  if (threadIdx.x == 0) *stop = 0;
  __syncthreads();
  if (warp < work_warps) {
    const int q = warp & 3, hcol = (warp >> 2) & 1;
    const int row = q * 32 + lane;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t trow = slot + ((uint32_t)(q << 5) << 16);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = hcol * 128 + cc * 32;
        uint32_t v[32];
        tmem_ld32(trow + col0, v);
        float4 b4[8];
        if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) b4[j] = reinterpret_cast<const float4*>(c_bias + (it & 7) * 256 + col0)[j];
        } else if (MODE == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(gbias + (it & 7) * 256 + col0) + j);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) b4[j] = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float h[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          h[4 * j] = __uint_as_float(v[4 * j]) + b4[j].x; h[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4[j].y;
          h[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4[j].z; h[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4[j].w;
        }
        const uint32_t cb = base + (col0 >> 6) * 16384;
        const int u0 = (col0 & 63) >> 3;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t addr = cb + row * 128 + ((((u0 + u) ^ (row & 7)) & 7) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(pack_relu(h[8 * u], h[8 * u + 1])),
                       "r"(pack_relu(h[8 * u + 2], h[8 * u + 3])), "r"(pack_relu(h[8 * u + 4], h[8 * u + 5])),
                       "r"(pack_relu(h[8 * u + 6], h[8 * u + 7])) : "memory");
        }
      }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; }
    __syncwarp();
    if (warp == 0 && lane == 0) *stop = 1;
  } else if (warp < work_warps + spin_warps) {
    // spin on a barrier that never completes, like the idle roles of the fused kernel
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&never);
    while (!*stop) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(slot) : "memory");
}

template <int MODE>
void run(const char* name, int work, int spin, const float* gb) {
  long long* out; cudaMalloc(&out, 148 * 8);
  const int iters = 500;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 576, 225 * 1024>>>(iters, work, spin, gb, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%-10s work_warps %2d spin_warps %2d: %7.1f cycles per 32-column iteration (%s)\n", name, work, spin,
         (double)h / iters / 4, cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  float* gb; cudaMalloc(&gb, 4096 * 4); cudaMemset(gb, 0, 4096 * 4);
  for (int work : {4, 8})
    for (int spin : {0, 10}) {
      run<0>("nobias", work, spin, gb);
      run<1>("constant", work, spin, gb);
      run<2>("global", work, spin, gb);
    }
  return 0;
}
