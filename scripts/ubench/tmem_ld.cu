// Micro-benchmark: tcgen05.ld throughput / latency per SM as a function of warps and loads in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

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

template <int INFLIGHT>
__global__ void k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) << 5) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[INFLIGHT][32];
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) tmem_ld32(base + ((i * INFLIGHT + j) * 32) % 512, v[j]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j)
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= v[j][q];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(slot) : "memory");
}

template <int INFLIGHT>
void run(int warps, int blocks) {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, blocks * sizeof(long long)); cudaMalloc(&sink, blocks * warps * 32 * 4);
  const int iters = 2000;
  k<INFLIGHT><<<blocks, warps * 32>>>(iters, out, sink);
  k<INFLIGHT><<<blocks, warps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * INFLIGHT * warps * 32 * 32 * 4;
  printf("warps %2d inflight %d blocks %3d: %8lld cycles, %6.1f cycles/iter, %7.1f B/clk/SM  (%s)\n", warps, INFLIGHT, blocks,
         h, (double)h / iters, bytes / h, cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int blocks : {1, 148})
    for (int warps : {1, 4, 8, 16}) {
      run<1>(warps, blocks); run<2>(warps, blocks); run<4>(warps, blocks);
    }
  return 0;
}
