// Exchange steps of the two data-parallel workloads as ONE kernel per GPU over NVLink peer memory (no NCCL on the data path).
//
// Reference: the reference is single-process; what has to be exchanged follows from its loops:
//   * NeRFail-S attack iteration (attack_NeRFail_S.py:331-392): grad_spatial_rgb [P,H,W,4] summed over the views of all
//     ranks, then the I-FGSM update  rgb <- clamp(rgb -/+ a * sign(grad), init - eps, init + eps)  on the rows with A > 0
//     (:357-392), identical on every rank.
//   * NeRF retraining step (run_nerf.py:791-800): parameter gradients averaged over ranks, then Adam.
// An all-reduce followed by a replicated element-wise update moves 2 (G-1)/G x N bytes per GPU and then touches all N
// elements on every GPU.  Here every GPU owns 1/G of the elements: it READS that slice of every peer's gradient buffer
// through peer pointers (reduce-scatter by P2P loads over NVSwitch), applies the update to its slice, and WRITES the
// updated slice into every peer's copy (all-gather by P2P stores) — reduce, update and broadcast in one launch, the
// transfer overlapping the arithmetic row by row.  For the attack only rows with A > 0 are touched at all.
//
// Buffers live in "peer memory": plain cudaMalloc allocations exported with cudaIpcGetMemHandle and opened by the other
// ranks of the node (one process per GPU), so a rank holds G pointers per symmetric buffer.  Cross-GPU ordering uses flag
// words in the same memory: st.release.sys / ld.acquire.sys on monotonically increasing epochs (never reset, graph-safe:
// the epoch is a device-resident counter the kernel itself advances).
//   phase 0  "my gradient is complete" -> every peer; wait until every peer said so       (all CTAs poll local flags)
//   phase 1  slice reduce + update + broadcast
//   phase 2  the last CTA of a GPU tells every peer "I have read your gradient and written your table", then waits for the
//            same message from every peer, so kernel completion on a GPU implies its own copies are final.
// Every wait is bounded; a time-out raises a status word in mapped pinned memory that the host polls (nfb_peer_status).
#include "common.cuh"

namespace nfb {

constexpr int MAX_PEERS = 8;
constexpr int FLAG_READY = 0, FLAG_DONE = MAX_PEERS;     // flag words per rank: ready[8], done[8]
constexpr int FLAG_WORDS = 2 * MAX_PEERS;

struct PeerPtrs {
  float* grad[MAX_PEERS];        // every rank's gradient buffer (index = rank); [rank] is local
  float* value[MAX_PEERS];       // every rank's copy of the updated quantity (perturbation table / flat parameters)
  uint32_t* flags[MAX_PEERS];    // every rank's flag words
  int rank, G;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float4* p) {        // peer data: bypass L1 (written by another GPU)
  float4 r;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_peer4(float4* p, float4 v) {
  asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// bounded spin on a local flag word until it reaches `epoch`
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch, volatile int* status) {
  for (uint32_t n = 0;; ++n) {
    if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return true;
    if ((n & 1023u) == 1023u) {
      if (*status) return false;
      if (n > (1u << 26)) { *status = 1; return false; }          // ~ seconds: a peer never arrived
    }
    __nanosleep(32);
  }
}

// phase 0: announce + wait.  Returns the epoch of this call.
__device__ __forceinline__ uint32_t peer_begin(const PeerPtrs& p, const uint32_t* epoch_ctr, volatile int* status) {
  const uint32_t epoch = *epoch_ctr + 1u;
  if (blockIdx.x == 0 && threadIdx.x < p.G) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + FLAG_READY + p.rank, epoch);
  }
  if (threadIdx.x < p.G) wait_flag(p.flags[p.rank] + FLAG_READY + threadIdx.x, epoch, status);
  __syncthreads();
  return epoch;
}

// phase 2: the last CTA of this GPU tells the peers and waits for them; it also advances the device-resident epoch.
__device__ __forceinline__ void peer_end(const PeerPtrs& p, uint32_t epoch, uint32_t* epoch_ctr, unsigned int* cta_counter,
                                         volatile int* status) {
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(cta_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x < p.G) {
    st_release_sys(p.flags[threadIdx.x] + FLAG_DONE + p.rank, epoch);
    wait_flag(p.flags[p.rank] + FLAG_DONE + threadIdx.x, epoch, status);
  }
  __syncthreads();
  if (threadIdx.x == 0) { *cta_counter = 0u; *epoch_ctr = epoch; }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }   // torch.sign

// attack_NeRFail_S.py:348-392 across G GPUs: rows [row0, row1) of the [T,4] table are this rank's slice.
template <int U>          // rows per thread and iteration: 2 with peers (NVLink latency), 1 without (plain bandwidth)
__global__ void __launch_bounds__(256, U == 2 ? 2 : 3)
attack_exchange_kernel(PeerPtrs p, const float4* __restrict__ init, int64_t T, float step, float eps,
                       uint32_t* epoch_ctr, unsigned int* cta_counter, int* status) {
  const uint32_t epoch = peer_begin(p, epoch_ctr, status);
  const int64_t per = (T + p.G - 1) / p.G;
  const int64_t row0 = per * p.rank, row1 = min(T, row0 + per);
  float4* mine = reinterpret_cast<float4*>(p.value[p.rank]);
  // U rows per thread and iteration, every load of all of them in flight before the first use: with peers the kernel is
  // bound by the latency of its dependent loads (table row -> peers' gradients), not by bandwidth.
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = row0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < row1; r += U * stride) {
    int64_t rr[U];
    float4 t[U];
    bool act[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      rr[u] = r + u * stride;
      act[u] = rr[u] < row1;
      t[u] = act[u] ? mine[rr[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
      act[u] = act[u] && (t[u].w > 0.f);              // inactive point: the update multiplies by (A > 0), nothing moves
    }
    float4 gq[U][MAX_PEERS], t0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      t0[u] = __ldg(init + rr[u]);
#pragma unroll
      for (int q = 0; q < MAX_PEERS; ++q)
        if (q < p.G) gq[u][q] = ld_peer4(reinterpret_cast<const float4*>(p.grad[q]) + rr[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      float gx = 0.f, gy = 0.f, gz = 0.f;             // rank order; every row has exactly one owner -> deterministic
#pragma unroll
      for (int q = 0; q < MAX_PEERS; ++q)
        if (q < p.G) { gx += gq[u][q].x; gy += gq[u][q].y; gz += gq[u][q].z; }
      float4 o = t[u];
      o.x = fmaxf(fminf(o.x - step * sgn(gx), t0[u].x + eps), t0[u].x - eps);
      o.y = fmaxf(fminf(o.y - step * sgn(gy), t0[u].y + eps), t0[u].y - eps);
      o.z = fmaxf(fminf(o.z - step * sgn(gz), t0[u].z + eps), t0[u].z - eps);
#pragma unroll
      for (int q = 0; q < MAX_PEERS; ++q)
        if (q < p.G) st_peer4(reinterpret_cast<float4*>(p.value[q]) + rr[u], o);
    }
  }
  peer_end(p, epoch, epoch_ctr, cta_counter, status);
}

// run_nerf.py:791-800 across G GPUs: flat parameter index space [0, n), this rank owns [i0, i1): mean gradient over ranks,
// Adam (torch.optim.Adam, no amsgrad / weight decay: the arithmetic of optim.cu's adam_kernel), new parameters to every rank.
// scalars[0] = lr / (1 - beta1^step), scalars[1] = 1 / sqrt(1 - beta2^step) (device memory: graph replays see the current step).
__global__ void __launch_bounds__(256, 3)
adam_exchange_kernel(PeerPtrs p, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, int64_t n,
                     const float* __restrict__ scalars, float one_minus_beta1, float beta2, float one_minus_beta2, float eps,
                     float grad_scale,
                     uint32_t* epoch_ctr, unsigned int* cta_counter, int* status) {
  const uint32_t epoch = peer_begin(p, epoch_ctr, status);
  const int64_t n4 = (n + 3) / 4;                                   // buffers are padded to a multiple of 4 floats
  const int64_t per = (n4 + p.G - 1) / p.G;
  const int64_t q0 = per * p.rank, q1 = min(n4, q0 + per);
  const float step_size = scalars[0], inv_bc2_sqrt = scalars[1];
  float4* mine = reinterpret_cast<float4*>(p.value[p.rank]);
  float4* m4 = reinterpret_cast<float4*>(exp_avg);
  float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
  for (int64_t i = q0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q1; i += (int64_t)gridDim.x * blockDim.x) {
    float4 gq[MAX_PEERS];
#pragma unroll
    for (int q = 0; q < MAX_PEERS; ++q)
      if (q < p.G) gq[q] = ld_peer4(reinterpret_cast<const float4*>(p.grad[q]) + i);
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < MAX_PEERS; ++q)
      if (q < p.G) { g[0] += gq[q].x; g[1] += gq[q].y; g[2] += gq[q].z; g[3] += gq[q].w; }
    float4 w = mine[i], m = m4[i], v = v4[i];
    float* wp = &w.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = g[k] * grad_scale;                          // the arithmetic of optim.cu's adam_kernel, bit for bit
      mp[k] = fmaf(gk - mp[k], one_minus_beta1, mp[k]);            // exp_avg.lerp_(grad, 1 - beta1)
      vp[k] = fmaf(vp[k], beta2, one_minus_beta2 * gk * gk);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vp[k]) * inv_bc2_sqrt + eps;
      wp[k] = wp[k] - step_size * (mp[k] / denom);
    }
    m4[i] = m; v4[i] = v;
#pragma unroll
    for (int q = 0; q < MAX_PEERS; ++q)
      if (q < p.G) st_peer4(reinterpret_cast<float4*>(p.value[q]) + i, w);
  }
  peer_end(p, epoch, epoch_ctr, cta_counter, status);
}

}  // namespace nfb

struct nfb_peer {
  nfb::PeerPtrs ptrs;
  uint32_t* epoch_ctr;          // device
  unsigned int* cta_counter;    // device
  int* status_dev;              // device alias of status_host
  volatile int* status_host;    // mapped pinned
};

extern "C" {

int nfb_peer_alloc(size_t bytes, void** out) {
  NFB_REQUIRE(out && bytes > 0, "peer_alloc: bad argument");
  void* p = nullptr;
  NFB_CUDA(cudaMalloc(&p, bytes));
  NFB_CUDA(cudaMemset(p, 0, bytes));
  NFB_CUDA(cudaDeviceSynchronize());
  *out = p;
  return NFB_OK;
}

int nfb_peer_free(void* p) {
  if (p) NFB_CUDA(cudaFree(p));
  return NFB_OK;
}

int nfb_peer_export(void* p, void* handle64) {
  NFB_REQUIRE(p && handle64, "peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  NFB_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  return NFB_OK;
}

int nfb_peer_import(const void* handle64, void** out) {
  NFB_REQUIRE(handle64 && out, "peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  NFB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return NFB_OK;
}

int nfb_peer_close(void* p) {
  if (p) NFB_CUDA(cudaIpcCloseMemHandle(p));
  return NFB_OK;
}

int nfb_peer_flag_bytes(void) { return nfb::FLAG_WORDS * (int)sizeof(uint32_t); }

int nfb_peer_create(nfb_peer_t** out, int rank, int G, void* const* grad_ptrs, void* const* value_ptrs, void* const* flag_ptrs) {
  NFB_REQUIRE(out && grad_ptrs && value_ptrs && flag_ptrs, "peer_create: null pointer");
  NFB_REQUIRE(G >= 1 && G <= nfb::MAX_PEERS && rank >= 0 && rank < G, "peer_create: rank %d of %d (at most %d ranks)", rank, G, nfb::MAX_PEERS);
  nfb_peer* h = new nfb_peer();
  memset(&h->ptrs, 0, sizeof(h->ptrs));
  h->ptrs.rank = rank; h->ptrs.G = G;
  for (int q = 0; q < G; ++q) {
    NFB_REQUIRE(grad_ptrs[q] && value_ptrs[q] && flag_ptrs[q], "peer_create: missing pointer of rank %d", q);
    NFB_REQUIRE(((reinterpret_cast<uintptr_t>(grad_ptrs[q]) | reinterpret_cast<uintptr_t>(value_ptrs[q])) & 15) == 0, "peer_create: buffers must be 16-byte aligned");
    h->ptrs.grad[q] = static_cast<float*>(grad_ptrs[q]);
    h->ptrs.value[q] = static_cast<float*>(value_ptrs[q]);
    h->ptrs.flags[q] = static_cast<uint32_t*>(flag_ptrs[q]);
  }
  h->epoch_ctr = nullptr; h->cta_counter = nullptr; h->status_host = nullptr; h->status_dev = nullptr;
  cudaError_t e = cudaMalloc(&h->epoch_ctr, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->epoch_ctr, 0, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->cta_counter, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->cta_counter, 0, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->status_host, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *h->status_host = 0; e = cudaHostGetDevicePointer((void**)&h->status_dev, (void*)h->status_host, 0); }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(h->epoch_ctr); cudaFree(h->cta_counter);
    if (h->status_host) cudaFreeHost((void*)h->status_host);
    delete h;
    return nfb::fail(NFB_E_CUDA, "peer_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return NFB_OK;
}

int nfb_peer_destroy(nfb_peer_t* h) {
  if (!h) return NFB_OK;
  cudaFree(h->epoch_ctr); cudaFree(h->cta_counter);
  if (h->status_host) cudaFreeHost((void*)h->status_host);
  delete h;
  return NFB_OK;
}

int nfb_peer_status(const nfb_peer_t* h) {
  NFB_REQUIRE(h, "peer_status: null handle");
  if (*h->status_host) return nfb::fail(NFB_E_CUDA, "peer exchange: a rank of the node never arrived (flag wait timed out); results are invalid");
  return NFB_OK;
}

static int exchange_grid(int64_t items, int ctas_per_sm = 4) {
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)nfb::sm_count() * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

int nfb_attack_exchange_step(nfb_peer_t* h, const float* init, int64_t T, float signed_step, float eps, void* stream) {
  NFB_REQUIRE(h && init && T > 0 && eps >= 0.f, "attack_exchange_step: bad argument");
  NFB_REQUIRE((reinterpret_cast<uintptr_t>(init) & 15) == 0, "attack_exchange_step: init must be 16-byte aligned");
  int rc = nfb_peer_status(h);
  if (rc != NFB_OK) return rc;
  const int64_t per = (T + h->ptrs.G - 1) / h->ptrs.G;
  if (h->ptrs.G > 1)      // two rows per thread, 2 CTAs per SM resident
    nfb::attack_exchange_kernel<2><<<exchange_grid((per + 1) / 2, 2), 256, 0, (cudaStream_t)stream>>>(
        h->ptrs, reinterpret_cast<const float4*>(init), T, signed_step, eps, h->epoch_ctr, h->cta_counter, h->status_dev);
  else
    nfb::attack_exchange_kernel<1><<<exchange_grid(per, 3), 256, 0, (cudaStream_t)stream>>>(
        h->ptrs, reinterpret_cast<const float4*>(init), T, signed_step, eps, h->epoch_ctr, h->cta_counter, h->status_dev);
  return nfb::check_launch("attack_exchange_step");
}

int nfb_adam_exchange_step(nfb_peer_t* h, float* exp_avg, float* exp_avg_sq, int64_t n, const float* step_scalars,
                           double beta1, double beta2, double eps, double grad_scale, void* stream) {
  NFB_REQUIRE(h && exp_avg && exp_avg_sq && step_scalars && n > 0, "adam_exchange_step: bad argument");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adam_exchange_step: state must be 16-byte aligned");
  int rc = nfb_peer_status(h);
  if (rc != NFB_OK) return rc;
  const int64_t per = ((n + 3) / 4 + h->ptrs.G - 1) / h->ptrs.G;
  nfb::adam_exchange_kernel<<<exchange_grid(per, 3), 256, 0, (cudaStream_t)stream>>>(      // 3 CTAs per SM are resident: one wave
      h->ptrs, exp_avg, exp_avg_sq, n, step_scalars, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
      (float)grad_scale,
      h->epoch_ctr, h->cta_counter, h->status_dev);
  return nfb::check_launch("adam_exchange_step");
}

}  // extern "C"
