// Ray generation, coarse depths, inverse-CDF sampling and the fused hierarchical step.
//
// Reference arithmetic:
//   get_rays        Create_spatial_point_set/nerf_pytorch/run_nerf_helpers.py:157-166, run_nerf.py:102-123
//   coarse depths   run_nerf.py:357-379
//   sample_pdf      run_nerf_helpers.py:200-243
//   hierarchical    run_nerf.py:392-396, :412
// All of these are HBM-bound and tiny next to the MLP (1 524 B/ray); the point of the
// kernels is that the CDF, the binary search and the 192-way sort live in shared memory per ray instead
// of the reference's expanded [R,128,63] gathers and a global sort.
#include "common.cuh"
#include <stdlib.h>

namespace nfb {

// torch.linspace(0, 1, n)[i] in fp32 exactly as ATen's CPU kernel evaluates it: step = 1/(n-1); forward from 0 below
// the midpoint (step*i), backward from 1 above it with a FUSED multiply-subtract (1 - step*(n-1-i), one rounding).
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n <= 1) return 0.f;
  const float step = __fdiv_rn(1.f, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(n - 1 - i), 1.f);
}

// In-kernel uniform numbers for the stochastic render path (run_nerf.py:365-379 t_rand, run_nerf_helpers.py:213 u): the
// reference draws them with torch.rand into [R,64] + [R,128] HBM tensors; here element p of stream `sid` is word p & 3 of
// Philox4x32-10 (Salmon et al. 2011 — the counter-based generator torch.cuda itself uses) at counter (p >> 2, sid, offset)
// under key `seed`, mapped to [0, 1) with 24 bits like torch.rand.  Stateless: any thread can produce any element.
struct Rng { unsigned long long seed, offset; int on; };
__device__ __forceinline__ float philox_uniform(const Rng& g, uint32_t sid, uint64_t p) {
  uint32_t c0 = (uint32_t)(p >> 2), c1 = (uint32_t)(p >> 34), c2 = sid ^ (uint32_t)(g.offset >> 32), c3 = (uint32_t)g.offset;
  uint32_t k0 = (uint32_t)g.seed, k1 = (uint32_t)(g.seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint32_t w = (p & 3) == 0 ? c0 : (p & 3) == 1 ? c1 : (p & 3) == 2 ? c2 : c3;
  return (float)(w >> 8) * 5.9604644775390625e-08f;      // 2^-24
}

__global__ void philox_uniform_kernel(Rng rng, uint32_t sid, int64_t n, float* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    out[p] = philox_uniform(rng, sid, (uint64_t)p);
}

__global__ void get_rays_kernel(int H, int W, float fx, float fy, float cx, float cy,
                                float r00, float r01, float r02, float r10, float r11, float r12,
                                float r20, float r21, float r22, float tx, float ty, float tz,
                                float near_, float far_, float* __restrict__ rays) {
  const int64_t n = (int64_t)H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(p % W), j = (int)(p / W);
    const float a = __fdiv_rn(__fsub_rn((float)i, cx), fx);
    const float b = -__fdiv_rn(__fsub_rn((float)j, cy), fy);
    const float c = -1.f;
    // sum(dirs * c2w[:3,:3], -1): products then left-to-right adds, no contraction
    const float d0 = __fadd_rn(__fadd_rn(__fmul_rn(a, r00), __fmul_rn(b, r01)), __fmul_rn(c, r02));
    const float d1 = __fadd_rn(__fadd_rn(__fmul_rn(a, r10), __fmul_rn(b, r11)), __fmul_rn(c, r12));
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(a, r20), __fmul_rn(b, r21)), __fmul_rn(c, r22));
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    float* o = rays + p * 11;
    o[0] = tx; o[1] = ty; o[2] = tz;
    o[3] = d0; o[4] = d1; o[5] = d2;
    o[6] = near_; o[7] = far_;
    o[8] = __fdiv_rn(d0, nrm); o[9] = __fdiv_rn(d1, nrm); o[10] = __fdiv_rn(d2, nrm);
  }
}

// render(rays=(rays_o, rays_d)) of run_nerf.py:95-123 for use_viewdirs = True, ndc = False: the [N,11] ray batch
// (o, d, near, far, d / |d|) from a [2,N,3] batch in one launch (the reference: norm, div, two ones_like, two cats).
__global__ void rays_from_batch_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t n, float near_,
                                       float far_, float* __restrict__ rays) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const float d0 = __ldg(d + 3 * p), d1 = __ldg(d + 3 * p + 1), d2 = __ldg(d + 3 * p + 2);
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    float* r = rays + p * 11;
    r[0] = __ldg(o + 3 * p); r[1] = __ldg(o + 3 * p + 1); r[2] = __ldg(o + 3 * p + 2);
    r[3] = d0; r[4] = d1; r[5] = d2;
    r[6] = near_; r[7] = far_;
    r[8] = __fdiv_rn(d0, nrm); r[9] = __fdiv_rn(d1, nrm); r[10] = __fdiv_rn(d2, nrm);
  }
}

// loss = img2mse(rgb, target) + img2mse(rgb0, target) (run_nerf.py:781-789, run_nerf_helpers.py:9) with its gradient in the
// same pass: out[0] = loss, out[1] = mse(rgb), out[2] = mse(rgb0), out[3] / out[4] = their PSNR; g = 2 (rgb - target) / n, g0 likewise.  One CTA, fixed
// reduction order (deterministic); n = 3 R is a few thousand to a few hundred thousand elements.
__global__ void __launch_bounds__(1024)
mse_loss2_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb0, const float* __restrict__ target, int64_t n,
                 float* __restrict__ out, float* __restrict__ g, float* __restrict__ g0) {
  __shared__ float s0[32], s1[32];
  float a = 0.f, b = 0.f;
  const float scale = 2.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float t = __ldg(target + i);
    const float e = rgb[i] - t;
    a = fmaf(e, e, a);
    g[i] = scale * e;
    if (rgb0) {
      const float e0 = rgb0[i] - t;
      b = fmaf(e0, e0, b);
      g0[i] = scale * e0;
    }
  }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? s0[threadIdx.x] : 0.f;
    b = threadIdx.x < (blockDim.x >> 5) ? s1[threadIdx.x] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
    if (threadIdx.x == 0) {
      const float m = a / (float)n, m0 = b / (float)n;
      out[0] = m + m0; out[1] = m; out[2] = m0;
      out[3] = -10.f * log10f(m);                       // mse2psnr (run_nerf_helpers.py:10), the two statistics of :783 / :788
      out[4] = rgb0 ? -10.f * log10f(m0) : 0.f;
    }
  }
}

__device__ __forceinline__ float coarse_depth(float near_, float far_, int i, int S, int lindisp) {
  const float t = linspace01(i, S);
  const float omt = __fsub_rn(1.f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near_, omt), __fmul_rn(far_, t));
  return __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, near_), omt), __fmul_rn(__fdiv_rn(1.f, far_), t)));
}

__global__ void coarse_z_kernel(const float* __restrict__ rays, int R, int S, int lindisp,
                                const float* __restrict__ t_rand, const Rng rng, float* __restrict__ z_vals) {
  const int64_t n = (int64_t)R * S;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(p / S), i = (int)(p % S);
    const float near_ = __ldg(rays + (int64_t)r * 11 + 6), far_ = __ldg(rays + (int64_t)r * 11 + 7);
    float zc = coarse_depth(near_, far_, i, S, lindisp);
    if (t_rand || rng.on) {   // run_nerf.py:365-379
      const float zl = (i > 0) ? coarse_depth(near_, far_, i - 1, S, lindisp) : zc;
      const float zu = (i + 1 < S) ? coarse_depth(near_, far_, i + 1, S, lindisp) : zc;
      const float lower = (i > 0) ? __fmul_rn(0.5f, __fadd_rn(zc, zl)) : zc;
      const float upper = (i + 1 < S) ? __fmul_rn(0.5f, __fadd_rn(zu, zc)) : zc;
      const float t = t_rand ? __ldg(t_rand + p) : philox_uniform(rng, 0u, (uint64_t)p);
      zc = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
    }
    z_vals[p] = zc;
  }
}

// Builds the CDF of run_nerf_helpers.py:202-205 for one ray into shared memory.
// w = weights + 1e-5; pdf = w / sum(w); cdf = [0, cumsum(pdf)] — the cumulative sum runs sequentially in
// double and is rounded to fp32 per element, which is what ATen's CPU cumsum does for float tensors.
// sum(w + 1e-5) over nw contiguous floats in the exact order of ATen's CPU reduction (vectorized_inner_sum in
// aten/src/ATen/native/cpu/SumKernel.cpp with 8-float vectors, which is what produced tests/golden): four
// interleaved vector accumulators over groups of four 8-wide vectors, left-over vectors into accumulator 0,
// accumulators folded 0 += 1, 2, 3, then the scalar tail summed first and the 8 lanes added to it in order.
// Valid while every accumulator sees < 16 vectors (nw < 512); the cascade above that is not reproduced.
// Why bother: searchsorted(cdf, u) sits on a knife edge wherever u equals a CDF knot (always for u = 1), so one ulp
// of the normaliser moves a sample across a bin; matching the summation order makes the sample indices bit-exact.
__device__ __forceinline__ float aten_cpu_sum_w(const float* __restrict__ w, int nw, int lane) {
  const int nv = nw >> 3, groups = nv >> 2;
  const int k = lane >> 3, l = lane & 7;          // accumulator k, vector lane l
  float acc = 0.f;
  for (int i = 0; i < groups; ++i) acc = __fadd_rn(acc, __fadd_rn(__ldg(w + ((i * 4 + k) << 3) + l), 1e-5f));
  if (k == 0)
    for (int v = groups * 4; v < nv; ++v) acc = __fadd_rn(acc, __fadd_rn(__ldg(w + (v << 3) + l), 1e-5f));
  const float a1 = __shfl_sync(FULL, acc, 8 + l), a2 = __shfl_sync(FULL, acc, 16 + l), a3 = __shfl_sync(FULL, acc, 24 + l);
  acc = __fadd_rn(__fadd_rn(__fadd_rn(acc, a1), a2), a3);      // meaningful in lanes 0..7
  float fin = 0.f;
  for (int i = nv << 3; i < nw; ++i) fin = __fadd_rn(fin, __fadd_rn(__ldg(w + i), 1e-5f));
#pragma unroll
  for (int j = 0; j < 8; ++j) fin = __fadd_rn(fin, __shfl_sync(FULL, acc, j));
  return fin;
}

__device__ __forceinline__ void build_cdf(const float* __restrict__ w, int nw, float* cdf, int lane) {
  float total;
  if (nw < 512) {
    total = aten_cpu_sum_w(w, nw, lane);
  } else {                                         // correctly rounded sum (double accumulation)
    double part = 0.0;
    for (int i = lane; i < nw; i += 32) part += (double)__fadd_rn(__ldg(w + i), 1e-5f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
    total = (float)part;
  }
  for (int i = lane; i < nw; i += 32) cdf[i + 1] = __fdiv_rn(__fadd_rn(__ldg(w + i), 1e-5f), total);
  __syncwarp();
  // ATen's CPU cumsum: sequential, double accumulator, every prefix rounded to fp32.  When each pdf value is >= 2^-29 all
  // partial sums (< 2) are multiples of 2^-52, i.e. EXACT in double, so the order of the additions cannot matter and a
  // warp scan of per-lane blocks yields the same bits as the sequential loop; otherwise (weights spanning > 2^29) lane 0
  // runs the sequential loop.  (The sequential loop on one lane was 60 % of this kernel's time.)
  const int per = (nw + 31) >> 5;
  const int b = lane * per, e = min(b + per, nw);
  double part = 0.0;
  float mn = INFINITY;
  for (int i = b; i < e; ++i) { const float p = cdf[i + 1]; part += (double)p; mn = fminf(mn, p); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
  if (mn >= 1.862645149230957e-09f && total == total) {         // 2^-29
    double incl = part;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += t;
    }
    double run = incl - part;
    for (int i = b; i < e; ++i) { run += (double)cdf[i + 1]; cdf[i + 1] = (float)run; }
    if (lane == 0) cdf[0] = 0.f;
  } else if (lane == 0) {
    double run = 0.0;
    cdf[0] = 0.f;
    for (int i = 1; i <= nw; ++i) { run += (double)cdf[i]; cdf[i] = (float)run; }
  }
  __syncwarp();
}

// 32-bit shared-window addressing for the search loops: with generic pointers into the dynamic shared-memory window ptxas
// re-derives the window base (S2R SR_CgaCtaId + LEA) inside every loop iteration, 13 instructions per search step.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }

// torch.searchsorted(cdf, u, right=True) then the gather / lerp of :227-241.
__device__ __forceinline__ float invert_cdf(uint32_t cdf, uint32_t bins, int nb, float u, int* ind_out) {
  // first index with cdf[idx] > u = number of leading knots that are not > u; branch-free descent with a trip count that
  // depends on nb only (the while-loop form cost ~16 instructions per step in divergent BSSY/BSYNC regions)
  int lo = 0;
  for (int step = 1 << (31 - __clz(nb)); step > 0; step >>= 1) {
    const int np = lo + step;
    if (np <= nb && !(lds_f32(cdf + 4 * (np - 1)) > u)) lo = np;
  }
  if (ind_out) *ind_out = lo;
  const int below = max(0, lo - 1), above = min(nb - 1, lo);
  const float c0 = lds_f32(cdf + 4 * below), c1 = lds_f32(cdf + 4 * above);
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.f;
  const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  const float b0 = lds_f32(bins + 4 * below), b1 = lds_f32(bins + 4 * above);
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

// dynamic smem: per warp 2*nb floats
__global__ void __launch_bounds__(128)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, int w_pitch,
                  const float* __restrict__ u, int R, int nb, int N, float* __restrict__ samples,
                  int32_t* __restrict__ inds) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* cdf = smem + (size_t)wib * 2 * nb;
  float* sb = cdf + nb;
  for (int r = blockIdx.x * wpb + wib; r < R; r += gridDim.x * wpb) {
    for (int i = lane; i < nb; i += 32) sb[i] = __ldg(bins + (int64_t)r * nb + i);
    build_cdf(weights + (int64_t)r * w_pitch, nb - 1, cdf, lane);
    for (int k = lane; k < N; k += 32) {
      const float uk = u ? __ldg(u + (int64_t)r * N + k) : linspace01(k, N);
      int ind;
      const float s = invert_cdf(smem_addr(cdf), smem_addr(sb), nb, uk, &ind);
      samples[(int64_t)r * N + k] = s;
      if (inds) inds[(int64_t)r * N + k] = ind;
    }
    __syncwarp();
  }
}

// dynamic smem per warp: 2*(Sc-1) floats (cdf, bins) + Sc (coarse) + P (new samples, padded to a power of two) +
// Sc+N (merged) floats.
// torch.sort(cat([z_vals, z_samples])) returns values only, so any procedure that emits the same multiset in
// ascending order is equivalent.  The coarse depths are already sorted; the new samples are sorted too in the
// deterministic path (u ascending, inverse CDF monotone) and need a 128-wide bitonic network otherwise.  The two
// sorted runs are then merged by rank: every element binary-searches the other run for its output position
// (coarse elements first among equals), ~40x fewer instructions than sorting all 192 values from scratch.
__global__ void __launch_bounds__(128)
hierarchical_kernel(const float* __restrict__ z_coarse, const float* __restrict__ weights,
                    const float* __restrict__ u, const Rng rng, int R, int Sc, int N, int P,
                    float* __restrict__ z_fine, float* __restrict__ z_samples, float* __restrict__ z_std) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nb = Sc - 1;
  const int Sf = Sc + N;
  const int per_warp = 2 * nb + Sc + P + Sf;
  float* cdf = smem + (size_t)wib * per_warp;
  float* sb = cdf + nb;
  float* zc = sb + nb;
  float* zs = zc + Sc;
  float* out = zs + P;
  const uint32_t cdf_a = smem_addr(cdf), sb_a = smem_addr(sb), zc_a = smem_addr(zc), zs_a = smem_addr(zs), out_a = smem_addr(out);
  for (int r = blockIdx.x * wpb + wib; r < R; r += gridDim.x * wpb) {
    for (int i = lane; i < Sc; i += 32) zc[i] = __ldg(z_coarse + (int64_t)r * Sc + i);
    for (int i = N + lane; i < P; i += 32) zs[i] = INFINITY;
    __syncwarp();
    for (int i = lane; i < nb; i += 32) sb[i] = __fmul_rn(0.5f, __fadd_rn(zc[i + 1], zc[i]));   // :392
    build_cdf(weights + (int64_t)r * Sc + 1, nb - 1, cdf, lane);                                // weights[...,1:-1]
    float sum = 0.f;
    for (int k = lane; k < N; k += 32) {
      const float uk = u ? __ldg(u + (int64_t)r * N + k)
                         : (rng.on ? philox_uniform(rng, 1u, (uint64_t)r * N + k) : linspace01(k, N));
      const float s = invert_cdf(cdf_a, sb_a, nb, uk, nullptr);
      zs[k] = s;
      sum += s;
      if (z_samples) z_samples[(int64_t)r * N + k] = s;
    }
    __syncwarp();
    if (z_std) {   // torch.std(z_samples, -1, unbiased=False)
      const float mean = warp_sum(sum) / (float)N;
      float sq = 0.f;
      for (int k = lane; k < N; k += 32) { const float dlt = zs[k] - mean; sq += dlt * dlt; }
      sq = warp_sum(sq);
      if (lane == 0) z_std[r] = sqrtf(sq / (float)N);
    }
    // random u: the new samples come out unordered.  Deterministic u: ordered up to rounding at bin boundaries —
    // verify (one pass) and only sort if an inversion exists, since the rank merge below needs sorted runs.
    bool unsorted = false;
    for (int k = lane; k + 1 < N; k += 32) unsorted |= zs[k] > zs[k + 1];
    unsorted = __any_sync(FULL, unsorted);
    if (unsorted && !u && !rng.on) {
      // deterministic u: inversions are isolated adjacent pairs at bin boundaries; a few odd-even transposition rounds
      // repair them for ~60 instructions each instead of the 28-pass bitonic network (which was 3/4 of this kernel)
      for (int round = 0; round < 4 && unsorted; ++round) {
        bool swapped = false;
        for (int par = 0; par < 2; ++par) {
          for (int k = 2 * lane + par; k + 1 < N; k += 64) {
            const float a = zs[k], b = zs[k + 1];
            if (a > b) { zs[k] = b; zs[k + 1] = a; swapped = true; }
          }
          __syncwarp();
        }
        unsorted = __any_sync(FULL, swapped);      // a round without swaps means the run is sorted
      }
    }
    if (unsorted) {
      for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < P; i += 32) {
            const int l = i ^ j;
            if (l > i) {
              const float a = zs[i], b = zs[l];
              const bool asc = (i & k) == 0;
              if ((a > b) == asc) { zs[i] = b; zs[l] = a; }
            }
          }
          __syncwarp();
        }
      }
    }
    // merge by rank
    const int step_n = 1 << (31 - __clz(N)), step_c = 1 << (31 - __clz(Sc));
    for (int i = lane; i < Sc; i += 32) {
      const float v = lds_f32(zc_a + 4 * i);
      int lo = 0;                            // number of new samples strictly below v
      for (int step = step_n; step > 0; step >>= 1) {
        const int np = lo + step;
        if (np <= N && lds_f32(zs_a + 4 * (np - 1)) < v) lo = np;
      }
      sts_f32(out_a + 4 * (i + lo), v);
    }
    for (int k = lane; k < N; k += 32) {
      const float v = lds_f32(zs_a + 4 * k);
      int lo = 0;                            // number of coarse depths <= v
      for (int step = step_c; step > 0; step >>= 1) {
        const int np = lo + step;
        if (np <= Sc && lds_f32(zc_a + 4 * (np - 1)) <= v) lo = np;
      }
      sts_f32(out_a + 4 * (k + lo), v);
    }
    __syncwarp();
    for (int i = lane; i < Sf; i += 32) z_fine[(int64_t)r * Sf + i] = out[i];
    __syncwarp();
  }
}

// ---- the lego shape (64 coarse depths, 128 new samples) with everything sized at compile time ------------------------
// Same arithmetic, same results as hierarchical_kernel (tests/test_gpu_kernels.py compares them bit for bit); what changes
// is the instruction count per ray (1920 -> ~1100 warp instructions, the kernel is issue-bound):
//   * searches are unrolled descents without bound checks: 63 knots = 32+16+8+4+2+1, so lo + step never leaves the
//     array; a 64- or 128-long run is one comparison with its last element plus a 6- / 7-step descent;
//   * a new sample needs NO search to find its place among the coarse depths: it was interpolated between the bin edges
//     mid[below] and mid[above], and mid[i] lies between z[i] and z[i+1], so "coarse depths <= sample" is below + 1 plus at
//     most a couple of comparisons (kept as a loop, so it stays correct for any rounding);
//   * loop trip counts (2 or 4 per lane) are constants, the shared-memory windows are fixed offsets.
template <int SC, int N>
__device__ __forceinline__ int count_less_unrolled(uint32_t arr, float v) {       // number of arr[0..N) strictly below v
  static_assert((N & (N - 1)) == 0, "power of two");
  if (lds_f32(arr + 4 * (N - 1)) < v) return N;
  int lo = 0;
#pragma unroll
  for (int step = N >> 1; step > 0; step >>= 1)
    if (lds_f32(arr + 4 * (lo + step - 1)) < v) lo += step;
  return lo;
}

template <int SC, int N>
__global__ void __launch_bounds__(128)
hierarchical_fixed_kernel(const float* __restrict__ z_coarse, const float* __restrict__ weights,
                          const float* __restrict__ u, const Rng rng, int R,
                          float* __restrict__ z_fine, float* __restrict__ z_samples, float* __restrict__ z_std) {
  constexpr int NB = SC - 1, SF = SC + N, PER_WARP = 2 * NB + SC + N + SF;
  static_assert(NB == 63 && (N & (N - 1)) == 0, "built for 64 coarse depths and a power-of-two sample count");
  __shared__ float smem[4 * PER_WARP];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* cdf = smem + wib * PER_WARP;
  float* sb = cdf + NB;
  float* zc = sb + NB;
  float* zs = zc + SC;
  float* out = zs + N;
  const uint32_t cdf_a = smem_addr(cdf), sb_a = smem_addr(sb), zc_a = smem_addr(zc), zs_a = smem_addr(zs), out_a = smem_addr(out);
  for (int r = blockIdx.x * 4 + wib; r < R; r += gridDim.x * 4) {
#pragma unroll
    for (int i = lane; i < SC; i += 32) zc[i] = __ldg(z_coarse + (int64_t)r * SC + i);
    __syncwarp();
#pragma unroll
    for (int i = lane; i < NB; i += 32) sb[i] = __fmul_rn(0.5f, __fadd_rn(zc[i + 1], zc[i]));   // :392
    build_cdf(weights + (int64_t)r * SC + 1, NB - 1, cdf, lane);                                // weights[...,1:-1]
    float sum = 0.f;
    int rank_c[N / 32];                        // coarse depths <= sample, per sample of this lane
    float val[N / 32];
#pragma unroll
    for (int j = 0; j < N / 32; ++j) {
      const int k = lane + 32 * j;
      const float uk = u ? __ldg(u + (int64_t)r * N + k) : (rng.on ? philox_uniform(rng, 1u, (uint64_t)r * N + k) : linspace01(k, N));
      int lo = 0;                              // torch.searchsorted(cdf, u, right=True): knots that are not > u
#pragma unroll
      for (int step = 32; step > 0; step >>= 1)
        if (!(lds_f32(cdf_a + 4 * (lo + step - 1)) > uk)) lo += step;
      const int below = max(0, lo - 1), above = min(NB - 1, lo);
      const float c0 = lds_f32(cdf_a + 4 * below), c1 = lds_f32(cdf_a + 4 * above);
      float denom = __fsub_rn(c1, c0);
      if (denom < 1e-5f) denom = 1.f;
      const float t = __fdiv_rn(__fsub_rn(uk, c0), denom);
      const float b0 = lds_f32(sb_a + 4 * below), b1 = lds_f32(sb_a + 4 * above);
      const float s = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
      zs[k] = s;
      val[j] = s;
      sum += s;
      if (z_samples) z_samples[(int64_t)r * N + k] = s;
      int c = below + 1;                       // z[below] <= mid[below] <= s; settle the few neighbours by comparison
      while (c < SC && lds_f32(zc_a + 4 * c) <= s) ++c;
      while (c > 0 && lds_f32(zc_a + 4 * (c - 1)) > s) --c;
      rank_c[j] = c;
    }
    __syncwarp();
    if (z_std) {   // torch.std(z_samples, -1, unbiased=False)
      const float mean = warp_sum(sum) / (float)N;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < N / 32; ++j) { const float dlt = val[j] - mean; sq += dlt * dlt; }
      sq = warp_sum(sq);
      if (lane == 0) z_std[r] = sqrtf(sq / (float)N);
    }
    bool unsorted = false;
#pragma unroll
    for (int j = 0; j < N / 32; ++j) { const int k = lane + 32 * j; if (k + 1 < N) unsorted |= zs[k] > zs[k + 1]; }
    unsorted = __any_sync(FULL, unsorted);
    if (unsorted) {
      // rare for deterministic u (isolated inversions at bin boundaries), the rule for random u: sort the new samples
      // (bitonic network), then re-derive each sample's coarse rank; the rank merge below needs sorted runs
      for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < N; i += 32) {
            const int l = i ^ j;
            if (l > i) {
              const float a = zs[i], b = zs[l];
              const bool asc = (i & k) == 0;
              if ((a > b) == asc) { zs[i] = b; zs[l] = a; }
            }
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int j = 0; j < N / 32; ++j) {
        const float v = zs[lane + 32 * j];
        val[j] = v;
        int c = 0;                             // number of coarse depths <= v
        if (lds_f32(zc_a + 4 * (SC - 1)) <= v) c = SC;
        else {
#pragma unroll
          for (int step = SC >> 1; step > 0; step >>= 1)
            if (lds_f32(zc_a + 4 * (c + step - 1)) <= v) c += step;
        }
        rank_c[j] = c;
      }
    }
    // merge by rank: coarse element i goes to i + (new samples strictly below it), sample k to k + (coarse depths <= it)
#pragma unroll
    for (int i = lane; i < SC; i += 32) {
      const float v = lds_f32(zc_a + 4 * i);
      sts_f32(out_a + 4 * (i + count_less_unrolled<SC, N>(zs_a, v)), v);
    }
#pragma unroll
    for (int j = 0; j < N / 32; ++j) sts_f32(out_a + 4 * (lane + 32 * j + rank_c[j]), val[j]);
    __syncwarp();
#pragma unroll
    for (int i = lane; i < SF; i += 32) z_fine[(int64_t)r * SF + i] = out[i];
    __syncwarp();
  }
}

static int grid_for(int64_t items, int per_block, int blocks_per_sm) {
  int64_t blocks = (items + per_block - 1) / per_block;
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace nfb

extern "C" {

int nfb_get_rays(int H, int W, const double* K, const float* c, float near_, float far_, float* rays, void* stream) {
  NFB_REQUIRE(H > 0 && W > 0 && K && c && rays, "get_rays: H=%d W=%d and non-null K, c2w, rays required", H, W);
  nfb::get_rays_kernel<<<nfb::grid_for((int64_t)H * W, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      H, W, (float)K[0], (float)K[4], (float)K[2], (float)K[5],
      c[0], c[1], c[2], c[4], c[5], c[6], c[8], c[9], c[10], c[3], c[7], c[11], near_, far_, rays);
  return nfb::check_launch("get_rays");
}

static int coarse_z_launch(const float* rays, int R, int S, int lindisp, const float* t_rand, nfb::Rng rng, float* z_vals, void* stream) {
  NFB_REQUIRE(rays && z_vals && R >= 0 && S > 0, "coarse_z: R=%d S=%d", R, S);
  if (R == 0) return NFB_OK;
  nfb::coarse_z_kernel<<<nfb::grid_for((int64_t)R * S, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      rays, R, S, lindisp, t_rand, rng, z_vals);
  return nfb::check_launch("coarse_z");
}

int nfb_coarse_z(const float* rays, int R, int S, int lindisp, const float* t_rand, float* z_vals, void* stream) {
  return coarse_z_launch(rays, R, S, lindisp, t_rand, nfb::Rng{0ull, 0ull, 0}, z_vals, stream);
}

int nfb_rays_from_batch(const float* rays_o, const float* rays_d, int64_t N, float near_, float far_, float* rays, void* stream) {
  NFB_REQUIRE(rays_o && rays_d && rays && N >= 0, "rays_from_batch: bad argument");
  if (N == 0) return NFB_OK;
  nfb::rays_from_batch_kernel<<<nfb::grid_for(N, 256, 8), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, N, near_, far_, rays);
  return nfb::check_launch("rays_from_batch");
}

int nfb_mse_loss2(const float* rgb, const float* rgb0, const float* target, int64_t n, float* out5, float* g_rgb, float* g_rgb0,
                  void* stream) {
  NFB_REQUIRE(rgb && target && out5 && g_rgb && n > 0 && (!rgb0 || g_rgb0), "mse_loss2: bad argument");
  nfb::mse_loss2_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rgb, rgb0, target, n, out5, g_rgb, g_rgb0);
  return nfb::check_launch("mse_loss2");
}

int nfb_philox_uniform(uint64_t seed, uint64_t offset, uint32_t stream_id, int64_t n, float* out, void* stream) {
  NFB_REQUIRE(out && n >= 0, "philox_uniform: bad argument");
  if (n == 0) return NFB_OK;
  nfb::philox_uniform_kernel<<<nfb::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(nfb::Rng{seed, offset, 1}, stream_id, n, out);
  return nfb::check_launch("philox_uniform");
}

int nfb_coarse_z_rng(const float* rays, int R, int S, int lindisp, uint64_t seed, uint64_t offset, float* z_vals, void* stream) {
  return coarse_z_launch(rays, R, S, lindisp, nullptr, nfb::Rng{seed, offset, 1}, z_vals, stream);
}

int nfb_sample_pdf(const float* bins, const float* weights, int w_pitch, const float* u,
                   int R, int nb, int N, float* samples, int32_t* inds, void* stream) {
  NFB_REQUIRE(bins && weights && samples, "sample_pdf: null pointer");
  NFB_REQUIRE(R >= 0 && nb >= 2 && N > 0 && w_pitch >= nb - 1, "sample_pdf: R=%d nb=%d N=%d w_pitch=%d", R, nb, N, w_pitch);
  if (nb > 2048) return nfb::fail(NFB_E_UNSUPPORTED, "sample_pdf: %d bins > 2048", nb);
  if (R == 0) return NFB_OK;
  // a block keeps (cdf, bins) of one ray per warp: 4 warps while that fits the default 48 KB window, fewer above (nb <= 2048
  // always fits with one warp: 16 KB)
  int warps = 4;
  while (warps > 1 && (size_t)warps * 2 * nb * sizeof(float) > 48 * 1024) warps >>= 1;
  const size_t smem = (size_t)warps * 2 * nb * sizeof(float);
  nfb::sample_pdf_kernel<<<nfb::grid_for(R, warps, 16), 32 * warps, smem, (cudaStream_t)stream>>>(
      bins, weights, w_pitch, u, R, nb, N, samples, inds);
  return nfb::check_launch("sample_pdf");
}

static int hierarchical_launch(const float* z_coarse, const float* weights, const float* u, nfb::Rng rng, int R, int Sc, int N,
                               float* z_fine, float* z_samples, float* z_std, void* stream) {
  NFB_REQUIRE(z_coarse && weights && z_fine, "hierarchical: null pointer");
  NFB_REQUIRE(R >= 0 && N > 0, "hierarchical: R=%d N=%d", R, N);
  if (Sc < 3 || Sc > 128 || Sc + N > 512)
    return nfb::fail(NFB_E_UNSUPPORTED, "hierarchical: need 3 <= Sc <= 128 and Sc+N <= 512 (Sc=%d N=%d)", Sc, N);
  if (R == 0) return NFB_OK;
  // the lego shape has its own compile-time-sized kernel; NERFAIL_B200_HIER=generic keeps the general one (cross-check)
  const char* hier_env = getenv("NERFAIL_B200_HIER");
  if (Sc == 64 && N == 128 && !(hier_env && hier_env[0] == 'g')) {
    nfb::hierarchical_fixed_kernel<64, 128><<<nfb::grid_for(R, 4, 16), 128, 0, (cudaStream_t)stream>>>(
        z_coarse, weights, u, rng, R, z_fine, z_samples, z_std);
    return nfb::check_launch("hierarchical");
  }
  int P = 1;
  while (P < N) P <<= 1;
  const size_t smem = (size_t)4 * (2 * (Sc - 1) + Sc + P + Sc + N) * sizeof(float);
  nfb::hierarchical_kernel<<<nfb::grid_for(R, 4, 16), 128, smem, (cudaStream_t)stream>>>(
      z_coarse, weights, u, rng, R, Sc, N, P, z_fine, z_samples, z_std);
  return nfb::check_launch("hierarchical");
}

int nfb_hierarchical(const float* z_coarse, const float* weights, const float* u, int R, int Sc, int N,
                     float* z_fine, float* z_samples, float* z_std, void* stream) {
  return hierarchical_launch(z_coarse, weights, u, nfb::Rng{0ull, 0ull, 0}, R, Sc, N, z_fine, z_samples, z_std, stream);
}

int nfb_hierarchical_rng(const float* z_coarse, const float* weights, uint64_t seed, uint64_t offset, int R, int Sc, int N,
                         float* z_fine, float* z_samples, float* z_std, void* stream) {
  return hierarchical_launch(z_coarse, weights, nullptr, nfb::Rng{seed, offset, 1}, R, Sc, N, z_fine, z_samples, z_std, stream);
}

}  // extern "C"
