// Shared helpers for libnerfail_b200 (sm_100a).  Host-side error plumbing + small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/nerfail_b200.h"

namespace nfb {

// ---- host: error text + launch accounting --------------------------------------------------------
char* err_buf();                       // thread-local, 512 bytes
int   fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(NFB_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return NFB_OK;
}

#define NFB_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) return nfb::fail(NFB_E_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

#define NFB_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return nfb::fail(NFB_E_ARG, __VA_ARGS__);      \
  } while (0)

int sm_count();   // cached multiprocessor count of the current device (148 on B200)

// One product dW = dY^T X of a grouped tensor-core weight-gradient launch (wgrad.cu).  dy / x point at chunk 0 of tile 0
// of a bf16 tile image; *_pitch = bytes between tiles; ndy / nx = 64-column chunks (dY: 2 or 4, X: 1, 2 or 4); dY chunks
// [ndy_real, ndy) are taken from a block of zeros.  Output window: rows [row_begin, row_end) x cols_valid columns are
// ACCUMULATED into out_w[(row - row_begin) * ld + col0 + c], the column sums of dY into out_b[row - row_begin] (or null).
struct WgradJob {
  const void* dy;  int64_t dy_pitch;
  const void* x;   int64_t x_pitch;
  float* out_w;
  float* out_b;
  int ndy, ndy_real, nx;
  int ld, col0, cols_valid, row_begin, row_end;
  // Optional side product on CUDA cores, riding on the stages this job loads anyway (the two small heads of the network):
  //   out_side_w[r * side_ld + c] += sum_rows g[row][side_gcol + r] * S[row][c],  out_side_b[r] += sum_rows g[row][side_gcol + r]
  // for r < side_rows (<= 3), c < side_cols (<= 256); g = the fp32 upstream gradient g_raw [M,4] (launch argument), S = this
  // job's X chunks (side_x == null) or side_nx (<= 2) extra chunks at side_x loaded next to them.
  int side_rows = 0, side_gcol = 0, side_cols = 0, side_ld = 0, side_nx = 0;
  const void* side_x = nullptr;
  float* out_side_w = nullptr;
  float* out_side_b = nullptr;
};
// ready / job_need / consumer_ctas: consumer mode (see wgrad.cu): dY is produced concurrently by mlp_train_kernel<BWD>, which
// publishes ready[tile] = number of its store groups that have landed; job j may load a tile once ready[tile] >= job_need[j].
// g_raw / M: the fp32 upstream gradient [M,4] the side products read (required when a job has side_rows > 0).
int launch_wgrad_grouped(const WgradJob* jobs, int njobs, int64_t ntiles, const void* zero16k, int* status, void* stream,
                         const char* what, const int* ready = nullptr, const signed char* job_need = nullptr, int consumer_ctas = 0,
                         const float* g_raw = nullptr, int64_t M = 0);

// ---- device ------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// streaming 128-bit load that does not allocate in L1 (read-once inputs)
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace nfb
