// Exact 8-nearest-neighbour search of every pixel's 3-D point against the P*H*W base-view points.
//
// Reference: Create_spatial_point_set/create_index_and_dist.py:126-145 — torch.cdist + torch.sort over
// 1600 candidate chunks with a running top-8 merge (it materialises and sorts ~3 GB per chunk).  The
// reference's cdist runs in matmul mode, whose fp32 cancellation makes its own ordering noisy
// (SURVEY.md §0.4); this kernel computes the direct-difference squared distance
//     d2 = ((dx*dx + dy*dy) + dz*dz)          (fp32, round-to-nearest, no FMA contraction)
// orders by (d2, candidate index) and returns sqrt(d2), which is what oracle/gauss_oracle.py:knn8_exact pins bit-exactly.
//
// Layout: one thread owns QPT queries and keeps their top-8 (d2, idx) sorted in registers; candidates stream
// through shared memory in tiles (every lane reads the same candidate -> LDS.128 broadcast).  Compute-bound
// on the FP32 pipe: ~9 instructions per (query, candidate) pair.
#include "common.cuh"
#include <stdlib.h>

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


// ---------------------------------------------------------------------------------------------------------------
// Exact 8-NN through a uniform grid (same results as knn8_kernel, bit for bit, ~1000x fewer distance evaluations)
// ---------------------------------------------------------------------------------------------------------------
// The candidates (the P*H*W base-view surface points, fixed for a whole data set) are bucketed once into a 256^3 grid
// over their bounding box, cells in Morton order, by a counting sort (histogram -> scan -> scatter; the order inside a
// cell is arbitrary because results are ordered by (d2, candidate index), never by visiting order).  Because Morton
// codes nest, the same table serves coarser levels: the level-s cell (edge h * 2^s) with code c owns the fine cells
// [c << 3s, (c + 1) << 3s).  A query visits Chebyshev shells r = 0, 1, 2 around its cell at level 0, 2, 4, 6, 8 in turn
// and stops as soon as its 8th best squared distance is below the (margin-reduced) squared gap to the unvisited
// region; level 8 is the whole grid, so background pixels far from every candidate degrade to the brute-force scan
// instead of failing.  Distances use the same fp32 expression as the brute-force kernel.
constexpr int KG_BITS = 8;
constexpr int KG_N = 1 << KG_BITS;                 // cells per axis
constexpr int KG_CELLS = 1 << (3 * KG_BITS);       // 16 777 216
constexpr int KG_RMAX = 2;
constexpr int KG_SCAN_BLOCK = 4096;                // elements per scan block (256 threads x 16)

struct KnnGridParams { float ox, oy, oz, h, inv_h, margin_abs; };

__host__ __device__ __forceinline__ uint32_t kg_spread(uint32_t v) {      // 8 bits -> every third bit
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__host__ __device__ __forceinline__ uint32_t kg_morton(int x, int y, int z) {
  return kg_spread((uint32_t)x) | (kg_spread((uint32_t)y) << 1) | (kg_spread((uint32_t)z) << 2);
}
__device__ __forceinline__ int kg_cell(float v, float o, float inv_h) {
  const float t = __fmul_rn(__fsub_rn(v, o), inv_h);
  int c = (int)floorf(t);
  c = c < 0 ? 0 : c;
  return c > KG_N - 1 ? KG_N - 1 : c;
}

__global__ void kg_hist_kernel(const float* __restrict__ cand, int64_t C, KnnGridParams g, int32_t* __restrict__ count) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= C) return;
  const uint32_t code = kg_morton(kg_cell(cand[i * 3], g.ox, g.inv_h), kg_cell(cand[i * 3 + 1], g.oy, g.inv_h), kg_cell(cand[i * 3 + 2], g.oz, g.inv_h));
  atomicAdd(count + code, 1);
}

// exclusive scan of KG_CELLS counts in three passes (block totals, scan of the 4096 totals, re-scan with offsets)
__global__ void __launch_bounds__(256) kg_scan_totals_kernel(const int32_t* __restrict__ count, int32_t* __restrict__ totals) {
  __shared__ int32_t red[8];
  const int4* src = reinterpret_cast<const int4*>(count + (int64_t)blockIdx.x * KG_SCAN_BLOCK) + threadIdx.x * 4;
  int32_t s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int4 v = src[k]; s += v.x + v.y + v.z + v.w; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { int32_t t = 0; for (int w = 0; w < 8; ++w) t += red[w]; totals[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) kg_scan_offsets_kernel(int32_t* __restrict__ totals, int nblocks) {   // one block
  __shared__ int32_t warp_sum_s[32];
  int32_t v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int i = threadIdx.x * 4 + k; v[k] = i < nblocks ? totals[i] : 0; s += v[k]; }
  int32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(FULL, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
  if ((threadIdx.x & 31) == 31) warp_sum_s[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    int32_t w = warp_sum_s[threadIdx.x], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(FULL, wi, o); if (threadIdx.x >= o) wi += t; }
    warp_sum_s[threadIdx.x] = wi - w;
  }
  __syncthreads();
  int32_t run = warp_sum_s[threadIdx.x >> 5] + incl - s;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int i = threadIdx.x * 4 + k; if (i < nblocks) totals[i] = run; run += v[k]; }
}
__global__ void __launch_bounds__(256) kg_scan_apply_kernel(const int32_t* __restrict__ count, const int32_t* __restrict__ totals,
                                                             int32_t* __restrict__ cell_start, int32_t C) {
  __shared__ int32_t warp_sum_s[8];
  const int64_t base = (int64_t)blockIdx.x * KG_SCAN_BLOCK + threadIdx.x * 16;
  int32_t v[16], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int4 q = reinterpret_cast<const int4*>(count + base)[k];
    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    s += q.x + q.y + q.z + q.w;
  }
  int32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(FULL, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
  if ((threadIdx.x & 31) == 31) warp_sum_s[threadIdx.x >> 5] = incl;
  __syncthreads();
  int32_t run = totals[blockIdx.x] + incl - s;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) run += warp_sum_s[w];
#pragma unroll
  for (int k = 0; k < 16; ++k) { cell_start[base + k] = run; run += v[k]; }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) cell_start[KG_CELLS] = C;
}
__global__ void kg_scatter_kernel(const float* __restrict__ cand, int64_t C, KnnGridParams g, const int32_t* __restrict__ cell_start,
                                  int32_t* __restrict__ cursor, float4* __restrict__ sorted) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= C) return;
  const float x = cand[i * 3], y = cand[i * 3 + 1], z = cand[i * 3 + 2];
  const uint32_t code = kg_morton(kg_cell(x, g.ox, g.inv_h), kg_cell(y, g.oy, g.inv_h), kg_cell(z, g.oz, g.inv_h));
  const int32_t pos = cell_start[code] + atomicAdd(cursor + code, 1);
  sorted[pos] = make_float4(x, y, z, __int_as_float((int)i));
}

// total order (d2, index): independent of the order candidates are visited in
__device__ __forceinline__ void top8_insert_ordered(Top8& t, float d2, int c, bool check_dup) {
  if (d2 < t.d[7] || (d2 == t.d[7] && c < t.id[7])) {
    if (check_dup) {
      bool dup = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) dup |= (t.id[k] == c);
      if (dup) return;
    }
    t.d[7] = d2; t.id[7] = c;
#pragma unroll
    for (int k = 7; k > 0; --k) {
      if (t.d[k] < t.d[k - 1] || (t.d[k] == t.d[k - 1] && t.id[k] < t.id[k - 1])) {
        const float td = t.d[k]; t.d[k] = t.d[k - 1]; t.d[k - 1] = td;
        const int ti = t.id[k]; t.id[k] = t.id[k - 1]; t.id[k - 1] = ti;
      }
    }
  }
}

__global__ void __launch_bounds__(KNN_THREADS)
knn8_grid_kernel(const float* __restrict__ query, int64_t Q, const float4* __restrict__ sorted, const int32_t* __restrict__ cell_start,
                 KnnGridParams g, float* __restrict__ out_dist, float* __restrict__ out_idx, int32_t* __restrict__ out_idx_i32,
                 unsigned long long* __restrict__ stats) {
  const int64_t q = blockIdx.x * (int64_t)KNN_THREADS + threadIdx.x;
  if (q >= Q) return;
  const float qx = __ldg(query + q * 3), qy = __ldg(query + q * 3 + 1), qz = __ldg(query + q * 3 + 2);
  const int cx = kg_cell(qx, g.ox, g.inv_h), cy = kg_cell(qy, g.oy, g.inv_h), cz = kg_cell(qz, g.oz, g.inv_h);
  Top8 best;
#pragma unroll
  for (int k = 0; k < 8; ++k) { best.d[k] = INFINITY; best.id[k] = -1; }
  unsigned long long evals = 0;
  bool done = false;
  for (int s = 0; s <= KG_BITS && !done; s += 2) {
    const int n = KG_N >> s;
    const int lx = cx >> s, ly = cy >> s, lz = cz >> s;
    const float hs = g.h * (float)(1 << s);
    const float margin = fmaf(1e-3f, hs, g.margin_abs);
    for (int r = 0; r <= KG_RMAX && !done; ++r) {
      const int z0 = max(lz - r, 0), z1 = min(lz + r, n - 1), y0 = max(ly - r, 0), y1 = min(ly + r, n - 1);
      const int x0 = max(lx - r, 0), x1 = min(lx + r, n - 1);
      for (int z = z0; z <= z1; ++z) {
        for (int y = y0; y <= y1; ++y) {
          const bool face = (z - lz == r) || (lz - z == r) || (y - ly == r) || (ly - y == r);
          // on a z / y face of the shell the whole x run belongs to it, otherwise only its two end cells
          const int step = face ? 1 : max(2 * r, 1);
          for (int x = face ? x0 : lx - r; x <= (face ? x1 : lx + r); x += step) {
            if (x < 0 || x >= n) continue;
            const uint32_t code = kg_morton(x, y, z);
            const int b = __ldg(cell_start + ((int64_t)code << (3 * s)));
            const int e = __ldg(cell_start + ((int64_t)(code + 1) << (3 * s)));
            for (int i = b; i < e; ++i) {
              const float4 p = __ldg(sorted + i);
              const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
              const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
              top8_insert_ordered(best, d2, __float_as_int(p.w), s > 0);
            }
            evals += (unsigned)(e - b);
          }
        }
      }
      // distance from the query to the nearest face of the visited block that still has cells behind it
      float gap = INFINITY;
      if (lx - r > 0) gap = fminf(gap, qx - fmaf((float)(lx - r), hs, g.ox));
      if (lx + r < n - 1) gap = fminf(gap, fmaf((float)(lx + r + 1), hs, g.ox) - qx);
      if (ly - r > 0) gap = fminf(gap, qy - fmaf((float)(ly - r), hs, g.oy));
      if (ly + r < n - 1) gap = fminf(gap, fmaf((float)(ly + r + 1), hs, g.oy) - qy);
      if (lz - r > 0) gap = fminf(gap, qz - fmaf((float)(lz - r), hs, g.oz));
      if (lz + r < n - 1) gap = fminf(gap, fmaf((float)(lz + r + 1), hs, g.oz) - qz);
      if (gap == INFINITY) { done = true; break; }                 // the block covers the whole grid
      const float gm = gap - margin;
      if (best.id[7] >= 0 && gm > 0.f && best.d[7] < gm * gm * 0.999999f) done = true;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (out_dist) out_dist[q * 8 + k] = __fsqrt_rn(best.d[k]);
    if (out_idx) out_idx[q * 8 + k] = (float)best.id[k];
    if (out_idx_i32) out_idx_i32[q * 8 + k] = best.id[k];
  }
  if (stats) {      // optional: total distance evaluations (pruning factor = Q * C / evals)
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(__activemask(), evals, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(stats, evals);
  }
}


// Warp-cooperative variant (default): one warp answers one query at a time.  The lanes fetch the cell ranges of a shell
// in parallel, then the warp scans each non-empty range with coalesced 16-byte loads, 32 candidates per step; a ballot
// finds the (rare) candidates that beat the current 8th best and they are inserted into the top-8, which every lane
// keeps identically in registers.  The per-thread walk above does 181x fewer distance evaluations than brute force but
// is only 3x faster (dependent, uncoalesced loads); this one turns the pruning into time.
__global__ void __launch_bounds__(KNN_THREADS)
knn8_grid_warp_kernel(const float* __restrict__ query, int64_t Q, const float4* __restrict__ sorted, const int32_t* __restrict__ cell_start,
                      KnnGridParams g, float* __restrict__ out_dist, float* __restrict__ out_idx, int32_t* __restrict__ out_idx_i32,
                      unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = (blockIdx.x * (int64_t)KNN_THREADS + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * (KNN_THREADS / 32);
  unsigned long long evals = 0;
  for (int64_t q = warp_id; q < Q; q += nwarps) {
    const float qx = __ldg(query + q * 3), qy = __ldg(query + q * 3 + 1), qz = __ldg(query + q * 3 + 2);
    const int cx = kg_cell(qx, g.ox, g.inv_h), cy = kg_cell(qy, g.oy, g.inv_h), cz = kg_cell(qz, g.oz, g.inv_h);
    Top8 best;
#pragma unroll
    for (int k = 0; k < 8; ++k) { best.d[k] = INFINITY; best.id[k] = -1; }
    bool done = false;
    for (int s = 0; s <= KG_BITS && !done; s += 2) {
      const int n = KG_N >> s;
      const int lx = cx >> s, ly = cy >> s, lz = cz >> s;
      const float hs = g.h * (float)(1 << s);
      const float margin = fmaf(1e-3f, hs, g.margin_abs);
      for (int r = 0; r <= KG_RMAX && !done; ++r) {
        const int w = 2 * r + 1, ncell = w * w * w;
        for (int t0 = 0; t0 < ncell; t0 += 32) {
          const int t = t0 + lane;
          int b = 0, e = 0;
          if (t < ncell) {
            const int dz = t / (w * w) - r, dy = (t / w) % w - r, dx = t % w - r;
            const int cheb = max(max(abs(dx), abs(dy)), abs(dz));
            const int x = lx + dx, y = ly + dy, z = lz + dz;
            if (cheb == r && x >= 0 && x < n && y >= 0 && y < n && z >= 0 && z < n) {
              const uint32_t code = kg_morton(x, y, z);
              b = __ldg(cell_start + ((int64_t)code << (3 * s)));
              e = __ldg(cell_start + ((int64_t)(code + 1) << (3 * s)));
            }
          }
          unsigned m = __ballot_sync(FULL, e > b);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int bb = __shfl_sync(FULL, b, src), ee = __shfl_sync(FULL, e, src);
            for (int i0 = bb; i0 < ee; i0 += 32) {
              const int i = i0 + lane;
              float d2 = INFINITY;
              int id = 0x7fffffff;
              if (i < ee) {
                const float4 p = __ldg(sorted + i);
                const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
                d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                id = __float_as_int(p.w);
              }
              unsigned cm = __ballot_sync(FULL, d2 < best.d[7] || (d2 == best.d[7] && id < best.id[7]));
              while (cm) {
                const int l = __ffs(cm) - 1;
                cm &= cm - 1;
                top8_insert_ordered(best, __shfl_sync(FULL, d2, l), __shfl_sync(FULL, id, l), s > 0);
              }
            }
            evals += (unsigned)(ee - bb);
          }
        }
        float gap = INFINITY;
        if (lx - r > 0) gap = fminf(gap, qx - fmaf((float)(lx - r), hs, g.ox));
        if (lx + r < n - 1) gap = fminf(gap, fmaf((float)(lx + r + 1), hs, g.ox) - qx);
        if (ly - r > 0) gap = fminf(gap, qy - fmaf((float)(ly - r), hs, g.oy));
        if (ly + r < n - 1) gap = fminf(gap, fmaf((float)(ly + r + 1), hs, g.oy) - qy);
        if (lz - r > 0) gap = fminf(gap, qz - fmaf((float)(lz - r), hs, g.oz));
        if (lz + r < n - 1) gap = fminf(gap, fmaf((float)(lz + r + 1), hs, g.oz) - qz);
        if (gap == INFINITY) { done = true; break; }
        const float gm = gap - margin;
        if (best.id[7] >= 0 && gm > 0.f && best.d[7] < gm * gm * 0.999999f) done = true;
      }
    }
    if (lane == 0 && out_dist) {
      reinterpret_cast<float4*>(out_dist + q * 8)[0] = make_float4(__fsqrt_rn(best.d[0]), __fsqrt_rn(best.d[1]), __fsqrt_rn(best.d[2]), __fsqrt_rn(best.d[3]));
      reinterpret_cast<float4*>(out_dist + q * 8)[1] = make_float4(__fsqrt_rn(best.d[4]), __fsqrt_rn(best.d[5]), __fsqrt_rn(best.d[6]), __fsqrt_rn(best.d[7]));
    }
    if (lane == 1 && out_idx) {
      reinterpret_cast<float4*>(out_idx + q * 8)[0] = make_float4((float)best.id[0], (float)best.id[1], (float)best.id[2], (float)best.id[3]);
      reinterpret_cast<float4*>(out_idx + q * 8)[1] = make_float4((float)best.id[4], (float)best.id[5], (float)best.id[6], (float)best.id[7]);
    }
    if (lane == 2 && out_idx_i32) {
      reinterpret_cast<int4*>(out_idx_i32 + q * 8)[0] = make_int4(best.id[0], best.id[1], best.id[2], best.id[3]);
      reinterpret_cast<int4*>(out_idx_i32 + q * 8)[1] = make_int4(best.id[4], best.id[5], best.id[6], best.id[7]);
    }
  }
  if (stats && lane == 0 && evals) atomicAdd(stats, evals);
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


int64_t nfb_knn_grid_cells(void) { return nfb::KG_CELLS; }

static int kg_params(const float* bbox_min_host, float h, float margin_abs, nfb::KnnGridParams* g) {
  NFB_REQUIRE(bbox_min_host && h > 0.f && margin_abs >= 0.f, "knn_grid: bad grid (h=%g)", (double)h);
  *g = nfb::KnnGridParams{bbox_min_host[0], bbox_min_host[1], bbox_min_host[2], h, 1.0f / h, margin_abs};
  return NFB_OK;
}

// Buckets cand [C,3] into the 256^3 Morton grid with origin bbox_min (host, 3 floats) and cell edge h: fills
// sorted [C,4] (x, y, z, original index as int bits) and cell_start [cells + 1]; workspace: cells + 4096 int32.
int nfb_knn_grid_build(const float* cand, int64_t C, const float* bbox_min_host, float h, float* sorted, int32_t* cell_start,
                       int32_t* workspace, void* stream) {
  NFB_REQUIRE(cand && sorted && cell_start && workspace, "knn_grid_build: null pointer");
  NFB_REQUIRE(C >= 8 && C <= (1 << 24), "knn_grid_build: C=%lld (need 8 <= C <= 2^24 for float32 indices)", (long long)C);
  nfb::KnnGridParams g;
  int rc = kg_params(bbox_min_host, h, 0.f, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* count = workspace;
  int32_t* totals = workspace + nfb::KG_CELLS;
  const int nblk = nfb::KG_CELLS / nfb::KG_SCAN_BLOCK;
  NFB_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * nfb::KG_CELLS, st));
  nfb::kg_hist_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(cand, C, g, count);
  if ((rc = nfb::check_launch("knn_grid_build.hist"))) return rc;
  nfb::kg_scan_totals_kernel<<<nblk, 256, 0, st>>>(count, totals);
  if ((rc = nfb::check_launch("knn_grid_build.totals"))) return rc;
  nfb::kg_scan_offsets_kernel<<<1, 1024, 0, st>>>(totals, nblk);
  if ((rc = nfb::check_launch("knn_grid_build.offsets"))) return rc;
  nfb::kg_scan_apply_kernel<<<nblk, 256, 0, st>>>(count, totals, cell_start, (int32_t)C);
  if ((rc = nfb::check_launch("knn_grid_build.scan"))) return rc;
  NFB_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * nfb::KG_CELLS, st));
  nfb::kg_scatter_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(cand, C, g, cell_start, count, reinterpret_cast<float4*>(sorted));
  return nfb::check_launch("knn_grid_build.scatter");
}

// Exact 8-NN of query [Q,3] against a grid built by nfb_knn_grid_build (same bbox_min / h); outputs as nfb_knn8.
// margin_abs: absolute slack on the pruning bound (a few ulps of the largest coordinate).  stats (device uint64, may be
// NULL) accumulates the number of distance evaluations.
int nfb_knn8_grid(const float* query, int64_t Q, const float* sorted, const int32_t* cell_start, const float* bbox_min_host, float h,
                  float margin_abs, float* out_dist, float* out_idx, int32_t* out_idx_i32, unsigned long long* stats, void* stream) {
  NFB_REQUIRE(query && sorted && cell_start && (out_dist || out_idx || out_idx_i32), "knn8_grid: null pointer");
  NFB_REQUIRE(Q >= 0, "knn8_grid: Q=%lld", (long long)Q);
  NFB_REQUIRE((reinterpret_cast<uintptr_t>(sorted) & 15) == 0, "knn8_grid: sorted must be 16-byte aligned");
  nfb::KnnGridParams g;
  int rc = kg_params(bbox_min_host, h, margin_abs, &g);
  if (rc) return rc;
  if (Q == 0) return NFB_OK;
  // NERFAIL_B200_KNN_GRID=thread selects the per-thread walk (kept as a cross-check); default: warp per query
  static const bool per_thread = []() { const char* e = getenv("NERFAIL_B200_KNN_GRID"); return e && e[0] == 't'; }();
  const bool aligned = ((reinterpret_cast<uintptr_t>(out_dist) | reinterpret_cast<uintptr_t>(out_idx) | reinterpret_cast<uintptr_t>(out_idx_i32)) & 15) == 0;
  if (per_thread || !aligned) {
    const int64_t blocks = (Q + nfb::KNN_THREADS - 1) / nfb::KNN_THREADS;
    nfb::knn8_grid_kernel<<<(unsigned)blocks, nfb::KNN_THREADS, 0, (cudaStream_t)stream>>>(
        query, Q, reinterpret_cast<const float4*>(sorted), cell_start, g, out_dist, out_idx, out_idx_i32, stats);
  } else {
    const int64_t warps_needed = Q, per_block = nfb::KNN_THREADS / 32;
    int64_t blocks = (warps_needed + per_block - 1) / per_block;
    const int64_t cap = (int64_t)nfb::sm_count() * 64;           // each warp then walks its queries with a grid stride
    if (blocks > cap) blocks = cap;
    nfb::knn8_grid_warp_kernel<<<(unsigned)blocks, nfb::KNN_THREADS, 0, (cudaStream_t)stream>>>(
        query, Q, reinterpret_cast<const float4*>(sorted), cell_start, g, out_dist, out_idx, out_idx_i32, stats);
  }
  return nfb::check_launch("knn8_grid");
}

}  // extern "C"
