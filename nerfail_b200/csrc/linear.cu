// fp32 layer primitives for the exact-parity / training path, plus the stand-alone positional encoding.
//
// Reference arithmetic: Create_spatial_point_set/nerf_pytorch/run_nerf_helpers.py:100-123 (the addmm / relu
// chain of NeRF.forward and its autograd) and :36-50 (Embedder).  These run on the FP32 FFMA pipe with fp32
// accumulation so that gradients can be checked against the fp32 oracle at 1e-3 relative; the render hot
// path uses the fused bf16 tcgen05 kernel in mlp_fused.cu instead.
//
// One tiled kernel serves the three contractions.  With r the reduction index:
//     C[i,j] = sum_r A(i,r) * B(j,r)
//   forward      i=row m, j=out n, r=k   A=X (r contiguous)   B=W (r contiguous)
//   bwd_data     i=row m, j=in  k, r=n   A=dY (r contiguous)  B=W (j contiguous)
//   bwd_weight   i=out n, j=in  k, r=m   A=dY (i contiguous)  B=X (j contiguous)   split over r
#include "common.cuh"

namespace nfb {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, GEMM_THREADS = 256;

struct GemmArgs {
  const float* A; int64_t lda;       // pitch of A's slow dimension
  const float* Amask; int64_t ldm;   // optional: A element is zeroed where Amask <= 0 (relu backward), same indexing
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  const float* bias;                 // per-j bias (forward) or null
  int64_t I, J, Rn;                  // extents
  int relu, accumulate;
  int64_t r_per_split;               // split-r: block z handles r in [z*r_per_split, ...), C advances by I*ldc per split
};

// One group = 4 consecutive elements of an operand tile along its CONTIGUOUS dimension, zero-filled outside the matrix
// (and where the relu mask says so).  RC (reduction index contiguous): group q of 512 = row q / 4, reduction offsets
// (q % 4) * 4 ..; otherwise: reduction row q / 32, output offsets (q % 32) * 4 ...  `vec` (pointer and pitch 16-byte
// aligned) takes the 128-bit load when the whole group lies inside the matrix.
template <bool RC, bool MASK>
__device__ __forceinline__ float4 load_group(const float* __restrict__ P, int64_t ld, const float* __restrict__ Mk, int64_t ldm,
                                             int64_t o0, int64_t extent, int64_t r0, int64_t r_end, int q, bool vec) {
  int64_t go, gr;                       // first element: output index, reduction index
  if (RC) { go = o0 + (q >> 2); gr = r0 + ((q & 3) << 2); } else { gr = r0 + (q >> 5); go = o0 + ((q & 31) << 2); }
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t off = RC ? go * ld + gr : gr * ld + go;
  const bool inside = RC ? (go < extent && gr + 3 < r_end) : (gr < r_end && go + 3 < extent);
  if (vec && inside) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(P + off));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    if (MASK) {
      const int64_t moff = RC ? go * ldm + gr : gr * ldm + go;
#pragma unroll
      for (int c = 0; c < 4; ++c) if (!(__ldg(Mk + moff + c) > 0.f)) v[c] = 0.f;
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool ok = RC ? (go < extent && gr + c < r_end) : (gr < r_end && go + c < extent);
      if (ok) {
        v[c] = __ldg(P + off + c);
        if (MASK) {
          const int64_t moff = RC ? go * ldm + gr : gr * ldm + go;
          if (!(__ldg(Mk + moff + c) > 0.f)) v[c] = 0.f;
        }
      }
    }
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

template <bool RC>
__device__ __forceinline__ void store_group(float (*S)[BM + PAD], int q, const float4 v) {
  if (RC) {
    const int oo = q >> 2, rr = (q & 3) << 2;
    S[rr][oo] = v.x; S[rr + 1][oo] = v.y; S[rr + 2][oo] = v.z; S[rr + 3][oo] = v.w;
  } else {
    *reinterpret_cast<float4*>(&S[q >> 5][(q & 31) << 2]) = v;
  }
}

// A_RC / B_RC: reduction index is the contiguous one for that operand.  The next K-slab is fetched into registers
// (128-bit loads where alignment allows) while the current one is multiplied out of shared memory; the products are
// accumulated in ascending reduction order whatever the tiling, so results do not depend on it.
template <bool A_RC, bool B_RC, bool MASK>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_kernel(GemmArgs g) {
  static_assert(BM == BN, "operand tiles share the loader");
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.x * BM, j0 = (int64_t)blockIdx.y * BN;
  const int64_t r_begin = (int64_t)blockIdx.z * g.r_per_split;
  const int64_t r_end = min(g.Rn, r_begin + g.r_per_split);
  const int ty = tid / 16, tx = tid % 16;     // 16 x 16 threads, 8 x 8 outputs each
  const bool vecA = ((reinterpret_cast<uintptr_t>(g.A) | (uintptr_t)(g.lda * 4)) & 15) == 0;
  const bool vecB = ((reinterpret_cast<uintptr_t>(g.B) | (uintptr_t)(g.ldb * 4)) & 15) == 0;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  float4 pa[2], pb[2];
  auto fetch = [&](int64_t r0) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      pa[e] = load_group<A_RC, MASK>(g.A, g.lda, g.Amask, g.ldm, i0, g.I, r0, r_end, tid + e * GEMM_THREADS, vecA);
      pb[e] = load_group<B_RC, false>(g.B, g.ldb, nullptr, 0, j0, g.J, r0, r_end, tid + e * GEMM_THREADS, vecB);
    }
  };
  if (r_begin < r_end) fetch(r_begin);
  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      store_group<A_RC>(As, tid + e * GEMM_THREADS, pa[e]);
      store_group<B_RC>(Bs, tid + e * GEMM_THREADS, pb[e]);
    }
    __syncthreads();
    if (r0 + BK < r_end) fetch(r0 + BK);
#pragma unroll
    for (int rr = 0; rr < BK; ++rr) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[rr][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[rr][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[rr][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[rr][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }

  float* C = g.C + (int64_t)blockIdx.z * g.I * g.ldc;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t gi = i0 + ty * 8 + a;
    if (gi >= g.I) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int64_t gj = j0 + tx * 8 + b;
      if (gj >= g.J) continue;
      float v = acc[a][b];
      if (g.bias) v += __ldg(g.bias + gj);
      if (g.relu) v = fmaxf(v, 0.f);
      float* dst = C + gi * g.ldc + gj;
      if (g.accumulate) v += *dst;
      *dst = v;
    }
  }
}

// dst[e] = sum_s part[s][e]   (deterministic split reduction)
__global__ void split_reduce_kernel(const float* __restrict__ part, int64_t n, int splits, float* __restrict__ dst) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(int64_t)k * n + e];
    dst[e] = s;
  }
}

// part[s][n] = sum over rows of split s of (dY o mask)[m][n]
__global__ void bias_partial_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ Y, int64_t ldy,
                                    int relu, int64_t M, int N, int64_t rows_per_split, float* __restrict__ part) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_split, m1 = min(M, m0 + rows_per_split);
  float s = 0.f;
  for (int64_t m = m0; m < m1; ++m) {
    float v = __ldg(dY + m * lddy + n);
    if (relu && !(__ldg(Y + m * ldy + n) > 0.f)) v = 0.f;
    s += v;
  }
  part[(int64_t)blockIdx.y * N + n] = s;
}

__global__ void embed_kernel(const float* __restrict__ x, int64_t M, int L, float* __restrict__ out, int ld, int col0,
                             int64_t row_repeat) {
  const int width = 3 + 6 * L;
  const int64_t n = M * 3;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = p / 3;
    const int c = (int)(p % 3);
    const float v = __ldg(x + (m / row_repeat) * 3 + c);
    float* o = out + m * ld + col0;
    o[c] = v;
    float f = 1.f;
    for (int l = 0; l < L; ++l) {          // freq = 2^l exactly (run_nerf_helpers.py:32), so x*freq is exact
      const float a = v * f;
      o[3 + 6 * l + c] = sinf(a);
      o[3 + 6 * l + 3 + c] = cosf(a);
      f *= 2.f;
    }
    (void)width;
  }
}

static int pick_splits(int64_t Rn, int64_t tiles) {
  int64_t want = (2LL * sm_count() + tiles - 1) / tiles;      // ~2 waves of blocks
  int64_t max_by_len = (Rn + 4 * BK - 1) / (4 * BK);
  if (want > max_by_len) want = max_by_len;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  return (int)want;
}

}  // namespace nfb

extern "C" {

int nfb_linear_fwd(const float* X, int ldx, const float* Wt, int ldw, const float* b, int64_t M, int N, int K,
                   int relu, float* Y, int ldy, void* stream) {
  NFB_REQUIRE(X && Wt && Y && M >= 0 && N > 0 && K > 0 && ldx >= K && ldw >= K && ldy >= N, "linear_fwd: bad argument");
  if (M == 0) return NFB_OK;
  nfb::GemmArgs g{X, ldx, nullptr, 0, Wt, ldw, Y, ldy, b, M, N, K, relu, 0, K};
  dim3 grid((unsigned)((M + nfb::BM - 1) / nfb::BM), (unsigned)((N + nfb::BN - 1) / nfb::BN), 1);
  nfb::gemm_kernel<true, true, false><<<grid, nfb::GEMM_THREADS, 0, (cudaStream_t)stream>>>(g);
  return nfb::check_launch("linear_fwd");
}

int nfb_linear_bwd_data(const float* dY, int lddy, const float* Y, int ldy, int relu, const float* Wt, int ldw,
                        int64_t M, int N, int K, float* dX, int lddx, int accumulate, void* stream) {
  NFB_REQUIRE(dY && Wt && dX && (!relu || Y) && M >= 0 && N > 0 && K > 0 && lddy >= N && ldw >= K && lddx >= K,
              "linear_bwd_data: bad argument");
  if (M == 0) return NFB_OK;
  // C[m,k] = sum_n dY[m,n] * W[n,k]: A = dY (reduction contiguous), B = W (output index contiguous)
  nfb::GemmArgs g{dY, lddy, relu ? Y : nullptr, ldy, Wt, ldw, dX, lddx, nullptr, M, K, N, 0, accumulate, N};
  dim3 grid((unsigned)((M + nfb::BM - 1) / nfb::BM), (unsigned)((K + nfb::BN - 1) / nfb::BN), 1);
  if (relu) nfb::gemm_kernel<true, false, true><<<grid, nfb::GEMM_THREADS, 0, (cudaStream_t)stream>>>(g);
  else      nfb::gemm_kernel<true, false, false><<<grid, nfb::GEMM_THREADS, 0, (cudaStream_t)stream>>>(g);
  return nfb::check_launch("linear_bwd_data");
}

static int bwd_weight_splits(int64_t M, int N, int K) {
  const int64_t tiles = (int64_t)((N + nfb::BM - 1) / nfb::BM) * ((K + nfb::BN - 1) / nfb::BN);
  return nfb::pick_splits(M, tiles);
}

int64_t nfb_linear_bwd_weight_workspace(int64_t M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int splits = bwd_weight_splits(M, N, K);
  return (int64_t)splits * ((int64_t)N * K + N) * (int64_t)sizeof(float);
}

int nfb_linear_bwd_weight(const float* dY, int lddy, const float* Y, int ldy, int relu, const float* X, int ldx,
                          int64_t M, int N, int K, float* dW, int lddw, float* db,
                          float* workspace, int64_t workspace_bytes, void* stream) {
  NFB_REQUIRE(dY && X && dW && (!relu || Y) && M > 0 && N > 0 && K > 0 && lddy >= N && ldx >= K && lddw >= K,
              "linear_bwd_weight: bad argument");
  NFB_REQUIRE(workspace && workspace_bytes >= nfb_linear_bwd_weight_workspace(M, N, K),
              "linear_bwd_weight: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
              (long long)nfb_linear_bwd_weight_workspace(M, N, K));
  NFB_REQUIRE(lddw == K, "linear_bwd_weight: dW must be dense (lddw == K)");
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = bwd_weight_splits(M, N, K);
  int64_t per = (M + splits - 1) / splits;
  per = (per + nfb::BK - 1) / nfb::BK * nfb::BK;
  // C[n,k] = sum_m dY[m,n] * X[m,k]: both operands have the output index contiguous; partials per split
  nfb::GemmArgs g{dY, lddy, relu ? Y : nullptr, ldy, X, ldx, workspace, K, nullptr, N, K, M, 0, 0, per};
  dim3 grid((unsigned)((N + nfb::BM - 1) / nfb::BM), (unsigned)((K + nfb::BN - 1) / nfb::BN), (unsigned)splits);
  if (relu) nfb::gemm_kernel<false, false, true><<<grid, nfb::GEMM_THREADS, 0, st>>>(g);
  else      nfb::gemm_kernel<false, false, false><<<grid, nfb::GEMM_THREADS, 0, st>>>(g);
  int rc = nfb::check_launch("linear_bwd_weight");
  if (rc) return rc;
  const int64_t nk = (int64_t)N * K;
  nfb::split_reduce_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(workspace, nk, splits, dW);
  rc = nfb::check_launch("linear_bwd_weight.reduce");
  if (rc) return rc;
  if (db) {
    float* bpart = workspace + (int64_t)splits * nk;
    dim3 bgrid((unsigned)((N + 127) / 128), (unsigned)splits, 1);
    nfb::bias_partial_kernel<<<bgrid, 128, 0, st>>>(dY, lddy, Y, ldy, relu, M, N, per, bpart);
    rc = nfb::check_launch("linear_bwd_weight.bias");
    if (rc) return rc;
    nfb::split_reduce_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(bpart, N, splits, db);
    rc = nfb::check_launch("linear_bwd_weight.bias_reduce");
  }
  return rc;
}

int nfb_embed(const float* x, int64_t M, int L, float* out, int ld, int col0, int64_t row_repeat, void* stream) {
  NFB_REQUIRE(x && out && M >= 0 && L >= 0 && L <= 16 && col0 >= 0 && ld >= col0 + 3 + 6 * L && row_repeat >= 1,
              "embed: bad argument");
  if (M == 0) return NFB_OK;
  int64_t blocks = (M * 3 + 255) / 256;
  const int64_t cap = (int64_t)nfb::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  nfb::embed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, M, L, out, ld, col0, row_repeat);
  return nfb::check_launch("embed");
}

}  // extern "C"
