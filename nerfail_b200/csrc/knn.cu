// Exact 8-nearest-neighbour search of every pixel's 3-D point against the P*H*W base-view points.
//
// Reference: Create_spatial_point_set/create_index_and_dist.py:126-145 — torch.cdist + torch.sort over
// 1600 candidate chunks with a running top-8 merge (it materialises and sorts ~3 GB per chunk).  The
// reference's cdist runs in matmul mode, whose fp32 cancellation makes its own ordering noisy
// (SURVEY.md §0.4); this kernel computes the direct-difference squared distance
//     d2 = ((dx*dx + dy*dy) + dz*dz)          (fp32, round-to-nearest, no FMA contraction)
// orders by (d2, candidate index) and returns sqrt(d2), which is what oracle/knn_oracle.py pins bit-exactly.
//
// Layout: one thread owns QPT queries and keeps their top-8 (d2, idx) sorted in registers; candidates stream
// through shared memory in tiles (every lane reads the same candidate -> LDS.128 broadcast).  Compute-bound
// on the FP32 pipe: ~9 instructions per (query, candidate) pair.
#include "common.cuh"

namespace nfb {

constexpr int KNN_TILE = 2048;     // candidates per shared-memory tile (32 KB as float4)
constexpr int KNN_THREADS = 256;

struct Top8 {
  float d[8];
  int id[8];
};

__device__ __forceinline__ void top8_insert(Top8& t, float d2, int c) {
  if (d2 < t.d[7]) {                 // strict: an equal later candidate never displaces an earlier one
    t.d[7] = d2; t.id[7] = c;
#pragma unroll
    for (int k = 7; k > 0; --k) {
      if (t.d[k] < t.d[k - 1]) {     // strict: stays behind equal, lower-index entries
        const float td = t.d[k]; t.d[k] = t.d[k - 1]; t.d[k - 1] = td;
        const int ti = t.id[k]; t.id[k] = t.id[k - 1]; t.id[k - 1] = ti;
      }
    }
  }
}

__global__ void __launch_bounds__(KNN_THREADS)
knn8_kernel(const float* __restrict__ query, int64_t Q, const float* __restrict__ cand, int64_t C,
            float* __restrict__ out_dist, float* __restrict__ out_idx, int32_t* __restrict__ out_idx_i32) {
  __shared__ float4 tile[KNN_TILE];
  const int64_t q = blockIdx.x * (int64_t)KNN_THREADS + threadIdx.x;
  const bool live = q < Q;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) { qx = __ldg(query + q * 3); qy = __ldg(query + q * 3 + 1); qz = __ldg(query + q * 3 + 2); }
  Top8 best;
#pragma unroll
  for (int k = 0; k < 8; ++k) { best.d[k] = INFINITY; best.id[k] = -1; }

  for (int64_t c0 = 0; c0 < C; c0 += KNN_TILE) {
    const int m = (int)min((int64_t)KNN_TILE, C - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += KNN_THREADS) {
      const float* s = cand + (c0 + i) * 3;
      tile[i] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.f);
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int i = 0; i < m; ++i) {
        const float4 p = tile[i];
        const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        top8_insert(best, d2, (int)(c0 + i));
      }
    }
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (out_dist) out_dist[q * 8 + k] = __fsqrt_rn(best.d[k]);
      if (out_idx) out_idx[q * 8 + k] = (float)best.id[k];        // the reference stores indices as float32 (:148-151)
      if (out_idx_i32) out_idx_i32[q * 8 + k] = best.id[k];
    }
  }
}

}  // namespace nfb

extern "C" {

int nfb_knn8(const float* query, int64_t Q, const float* cand, int64_t C,
             float* out_dist, float* out_idx, int32_t* out_idx_i32, void* stream) {
  NFB_REQUIRE(query && cand && (out_dist || out_idx || out_idx_i32), "knn8: null pointer");
  NFB_REQUIRE(Q >= 0 && C >= 8, "knn8: need at least 8 candidates (Q=%lld C=%lld)", (long long)Q, (long long)C);
  if (C > (1 << 24)) return nfb::fail(NFB_E_UNSUPPORTED, "knn8: %lld candidates do not fit a float32 index", (long long)C);
  if (Q == 0) return NFB_OK;
  const int64_t blocks = (Q + nfb::KNN_THREADS - 1) / nfb::KNN_THREADS;
  nfb::knn8_kernel<<<(unsigned)blocks, nfb::KNN_THREADS, 0, (cudaStream_t)stream>>>(
      query, Q, cand, C, out_dist, out_idx, out_idx_i32);
  return nfb::check_launch("knn8");
}

}  // extern "C"
