// Micro-benchmark: how fast can ONE SM pull global memory into shared memory, as a function of how many SMs pull at
// the same time, the copy mechanism and the bytes in flight?  (Question behind it: the grouped weight-gradient kernel
// sustains ~47 GB/s per SM at every grid size from 68 to 148 CTAs — is that a per-SM limit of cp.async.bulk?)
//   mode 0: cp.async.bulk (UBLKCP) pieces of `piece` bytes, `stages` x `stage_bytes` ring, mbarrier completion
//   mode 1: LDG.128 by all threads into registers (sum), `unroll` independent loads per thread in flight
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ingest ingest.cu && ./ingest
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// every CTA streams `bytes_per_cta` starting at its own offset (wrapping inside `span` bytes: span <= L2 -> L2 hits)
__global__ void __launch_bounds__(128, 1) bulk_kernel(const char* src, size_t span, size_t bytes_per_cta, int stage_bytes,
                                                      int stages, int piece, unsigned long long* t_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bars = base + stages * stage_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const size_t nst = bytes_per_cta / stage_bytes;
    size_t off = ((size_t)blockIdx.x * bytes_per_cta) % span;
    // prologue: fill the ring; steady state: wait stage i, immediately re-issue it
    for (size_t i = 0; i < nst + stages; ++i) {
      const int s = (int)(i % stages);
      if (i >= (size_t)stages) mbar_wait(bars + 8 * s, (uint32_t)(((i / stages) - 1) & 1));
      if (i < nst) {
        mbar_expect(bars + 8 * s, stage_bytes);
        for (int p = 0; p < stage_bytes; p += piece) {
          bulk_g2s(base + s * stage_bytes + p, src + off, piece, bars + 8 * s);
          off += piece;
          if (off + piece > span) off = 0;
        }
      }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    t_out[2 * blockIdx.x] = t0;
    t_out[2 * blockIdx.x + 1] = t1;
  }
}

template <int UNROLL>
__global__ void __launch_bounds__(512, 1) ldg_kernel(const uint4* src, size_t span16, size_t n16_per_cta, unsigned long long* t_out,
                                                     unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];       // only to force one CTA per SM
  unsigned long long t0, t1;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  size_t off = ((size_t)blockIdx.x * n16_per_cta) % span16;
  unsigned acc = 0;
  for (size_t i = 0; i < n16_per_cta; i += (size_t)blockDim.x * UNROLL) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      size_t idx = off + i + (size_t)u * blockDim.x + threadIdx.x;
      if (idx >= span16) idx -= span16;
      v[u] = __ldg(src + idx);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    t_out[2 * blockIdx.x] = t0;
    t_out[2 * blockIdx.x + 1] = t1;
  }
}

static void report(const char* what, int grid, size_t bytes_per_cta, unsigned long long* d_t) {
  static unsigned long long h[2 * 160];
  cudaMemcpy(h, d_t, sizeof(unsigned long long) * 2 * grid, cudaMemcpyDeviceToHost);
  unsigned long long lo = ~0ull, hi = 0;
  double per = 0;
  for (int b = 0; b < grid; ++b) {
    if (h[2 * b] < lo) lo = h[2 * b];
    if (h[2 * b + 1] > hi) hi = h[2 * b + 1];
    per += (double)bytes_per_cta / (double)(h[2 * b + 1] - h[2 * b]);
  }
  printf("%-44s grid %3d : %7.1f GB/s per SM (mean of CTAs), %8.1f GB/s aggregate\n", what, grid, per / grid,
         (double)bytes_per_cta * grid / (double)(hi - lo));
}

int main() {
  const size_t BIG = (size_t)4 << 30, SMALL = (size_t)48 << 20;
  char* buf;
  cudaMalloc(&buf, BIG);
  cudaMemset(buf, 1, BIG);
  unsigned long long* d_t;
  cudaMalloc(&d_t, sizeof(unsigned long long) * 2 * 160);
  unsigned* sink;
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(ldg_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int grids[] = {1, 16, 37, 74, 111, 148};
  for (int l2 = 0; l2 < 2; ++l2) {
    const size_t span = l2 ? SMALL : BIG;
    const size_t per_cta = l2 ? ((size_t)96 << 20) : ((size_t)24 << 20);
    printf("==== source: %s ====\n", l2 ? "48 MB span (L2 resident after the first pass)" : "4 GB span (HBM)");
    struct Cfg { int stage_bytes, stages, piece; } cfgs[] = {
        {65536, 3, 8192}, {65536, 3, 65536}, {32768, 6, 8192}, {24576, 8, 8192}, {16384, 12, 16384}, {8192, 24, 8192}, {65536, 3, 2048}};
    for (const Cfg& c : cfgs) {
      char what[96];
      snprintf(what, sizeof what, "bulk %2d x %3d KB stages, %2d KB pieces", c.stages, c.stage_bytes / 1024, c.piece / 1024);
      for (int g : grids) {
        if (l2) bulk_kernel<<<g, 128, 200 * 1024>>>(buf, span, span / 2, c.stage_bytes, c.stages, c.piece, d_t);   // warm L2
        bulk_kernel<<<g, 128, 200 * 1024>>>(buf, span, per_cta, c.stage_bytes, c.stages, c.piece, d_t);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        report(what, g, per_cta, d_t);
      }
    }
    for (int g : grids) {
      ldg_kernel<8><<<g, 512, 200 * 1024>>>((const uint4*)buf, span / 16, per_cta / 16, d_t, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      report("LDG.128 x 8 in flight per thread, 512 threads", g, per_cta, d_t);
    }
  }
  return 0;
}
