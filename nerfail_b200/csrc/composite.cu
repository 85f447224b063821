// Alpha compositing (raw2outputs) forward / backward and the pts_max epilogue.
//
// Reference arithmetic: Create_spatial_point_set/nerf_pytorch/run_nerf.py:262-305 (raw2outputs) and
// Create_spatial_point_set/nerf_to_coord.py:418-421 (pts_max).  HBM-bound: 24 B/sample forward,
// 40 B/sample backward.  One warp owns one ray; samples are visited in chunks of 32
// (lane = sample within the chunk, so every global access is a fully coalesced 128/512-byte request)
// and the exclusive transmittance product is a warp shuffle scan carried from chunk to chunk.
#include "common.cuh"
#include <stdlib.h>

namespace nfb {

struct SampleTerms {
  float alpha, t, z, dist, sigma_in;   // sigma_in = raw sigma + noise (before relu)
  float4 raw;
};

// Loads sample i of ray r and evaluates the per-sample quantities of run_nerf.py:277-293.
__device__ __forceinline__ SampleTerms load_sample(const float* __restrict__ raw, const float* __restrict__ z,
                                                   const float* __restrict__ noise, int64_t base, int i, int S,
                                                   float dnorm, bool valid) {
  SampleTerms s;
  s.raw = make_float4(0.f, 0.f, 0.f, 0.f);
  s.z = 0.f; s.dist = 0.f; s.alpha = 0.f; s.t = 1.f; s.sigma_in = 0.f;
  if (valid) {
    s.raw = ld_stream4(reinterpret_cast<const float4*>(raw) + base + i);
    s.z = __ldg(z + base + i);
    float gap = (i + 1 < S) ? __fsub_rn(__ldg(z + base + i + 1), s.z) : 1e10f;   // :277-278
    s.dist = __fmul_rn(gap, dnorm);                                              // :280
    s.sigma_in = s.raw.w + (noise ? __ldg(noise + base + i) : 0.f);
    float sig = fmaxf(s.sigma_in, 0.f);
    s.alpha = 1.f - expf(-__fmul_rn(sig, s.dist));                              // :275
    s.t = (1.f - s.alpha) + 1e-10f;                                             // :295
  }
  return s;
}

// inclusive multiplicative scan over the 32 lanes
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
// inclusive additive suffix scan (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_down_sync(FULL, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}

// 1 / y and __frcp_rn(y) are both the correctly rounded reciprocal: identical bits, no division slow path
__device__ __forceinline__ float sigmoidf(float x) { return __frcp_rn(1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     int ray_pitch, const float* __restrict__ noise, int R, int S, int white,
                     float* __restrict__ rgb_map, float* __restrict__ disp, float* __restrict__ acc_out,
                     float* __restrict__ weights, float* __restrict__ depth_out, float* __restrict__ pts_max) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nchunk = (S + 31) >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < R; r += gridDim.x * warps_per_block) {
    const float* d = rays_d + (int64_t)r * ray_pitch;
    const float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const int64_t base = (int64_t)r * S;
    float carry = 1.f, acc = 0.f, dep = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    float best_w = -INFINITY, best_z = 0.f;
    int best_i = 0x7fffffff;
    for (int j = 0; j < nchunk; ++j) {
      const int i = (j << 5) + lane;
      const bool valid = i < S;
      SampleTerms s = load_sample(raw, z, noise, base, i, S, dnorm, valid);
      float incl = warp_scan_mul(s.t, lane);
      float excl = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) excl = 1.f;
      const float T = carry * excl;
      carry *= __shfl_sync(FULL, incl, 31);
      const float w = s.alpha * T;
      if (valid) {
        weights[base + i] = w;
        acc += w;
        dep += w * s.z;
        cr += w * sigmoidf(s.raw.x);
        cg += w * sigmoidf(s.raw.y);
        cb += w * sigmoidf(s.raw.z);
        if (w > best_w) { best_w = w; best_i = i; best_z = s.z; }   // strict > keeps the first maximum in-lane
      }
    }
    acc = warp_sum(acc); dep = warp_sum(dep);
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    if (pts_max) {   // torch.argmax: first index attaining the maximum
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ow = __shfl_xor_sync(FULL, best_w, o);
        int oi = __shfl_xor_sync(FULL, best_i, o);
        float oz = __shfl_xor_sync(FULL, best_z, o);
        if (ow > best_w || (ow == best_w && oi < best_i)) { best_w = ow; best_i = oi; best_z = oz; }
      }
    }
    if (lane == 0) {
      if (white) { const float bg = 1.f - acc; cr += bg; cg += bg; cb += bg; }   // :302-303
      if (rgb_map) { rgb_map[(int64_t)r * 3] = cr; rgb_map[(int64_t)r * 3 + 1] = cg; rgb_map[(int64_t)r * 3 + 2] = cb; }
      if (acc_out) acc_out[r] = acc;
      if (depth_out) depth_out[r] = dep;
      if (disp) {
        const float q = dep / acc;                 // 0/0 -> NaN exactly as the reference (:299)
        disp[r] = 1.f / fmaxf(1e-10f, q);          // fmaxf(1e-10, NaN) = 1e-10: torch.max propagates NaN instead
        if (q != q) disp[r] = q;                   // keep the reference's NaN
      }
      if (pts_max) {                               // nerf_to_coord.py:397 + :421 : o + d * z
        const float* o = d - 3;
        pts_max[(int64_t)r * 3]     = __fadd_rn(__ldg(o),     __fmul_rn(dx, best_z));
        pts_max[(int64_t)r * 3 + 1] = __fadd_rn(__ldg(o + 1), __fmul_rn(dy, best_z));
        pts_max[(int64_t)r * 3 + 2] = __fadd_rn(__ldg(o + 2), __fmul_rn(dz, best_z));
      }
    }
  }
}

// Same arithmetic with the chunk loop unrolled (NCH = ceil(S / 32) known at compile time): all of a ray's loads are in
// flight at once, the gap to the next depth comes from a shuffle instead of a second load, and the NCH chunk-local
// scans are independent; only the product of the chunk totals is a serial chain.  Bit-identical to the loop form
// (same per-chunk scan order, same carry products) and ~2x faster: the loop form exposed one load latency plus one
// 5-step shuffle scan per chunk on the critical path.
template <int NCH>
__global__ void __launch_bounds__(256)
composite_fwd_unrolled_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                              int ray_pitch, const float* __restrict__ noise, int R, int S, int white,
                              float* __restrict__ rgb_map, float* __restrict__ disp, float* __restrict__ acc_out,
                              float* __restrict__ weights, float* __restrict__ depth_out, float* __restrict__ pts_max) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < R; r += gridDim.x * warps_per_block) {
    const float* d = rays_d + (int64_t)r * ray_pitch;
    const float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const int64_t base = (int64_t)r * S;
    float4 rw[NCH];
    float zz[NCH], nz[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int i = (j << 5) + lane;
      const bool valid = i < S;
      rw[j] = valid ? ld_stream4(reinterpret_cast<const float4*>(raw) + base + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      zz[j] = valid ? __ldg(z + base + i) : 0.f;
      nz[j] = (valid && noise) ? __ldg(noise + base + i) : 0.f;
    }
    float alpha[NCH], incl[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int i = (j << 5) + lane;
      const bool valid = i < S;
      float znext = __shfl_down_sync(FULL, zz[j], 1);
      const float zfirst_next = __shfl_sync(FULL, zz[j + 1 < NCH ? j + 1 : j], 0);
      if (lane == 31) znext = zfirst_next;
      const float gap = (i + 1 < S) ? __fsub_rn(znext, zz[j]) : 1e10f;                 // :277-278
      const float dist = __fmul_rn(gap, dnorm);                                        // :280
      const float sig = fmaxf(rw[j].w + nz[j], 0.f);
      alpha[j] = valid ? 1.f - expf(-__fmul_rn(sig, dist)) : 0.f;                      // :275
      const float t = valid ? (1.f - alpha[j]) + 1e-10f : 1.f;                         // :295
      incl[j] = warp_scan_mul(t, lane);
    }
    float carry = 1.f, acc = 0.f, dep = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    float best_w = -INFINITY, best_z = 0.f;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int i = (j << 5) + lane;
      float excl = __shfl_up_sync(FULL, incl[j], 1);
      if (lane == 0) excl = 1.f;
      const float T = carry * excl;
      carry *= __shfl_sync(FULL, incl[j], 31);
      const float w = alpha[j] * T;
      if (i < S) {
        weights[base + i] = w;
        acc += w;
        dep += w * zz[j];
        cr += w * sigmoidf(rw[j].x);
        cg += w * sigmoidf(rw[j].y);
        cb += w * sigmoidf(rw[j].z);
        if (w > best_w) { best_w = w; best_i = i; best_z = zz[j]; }
      }
    }
    acc = warp_sum(acc); dep = warp_sum(dep);
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    if (pts_max) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ow = __shfl_xor_sync(FULL, best_w, o);
        int oi = __shfl_xor_sync(FULL, best_i, o);
        float oz = __shfl_xor_sync(FULL, best_z, o);
        if (ow > best_w || (ow == best_w && oi < best_i)) { best_w = ow; best_i = oi; best_z = oz; }
      }
    }
    if (lane == 0) {
      if (white) { const float bg = 1.f - acc; cr += bg; cg += bg; cb += bg; }
      if (rgb_map) { rgb_map[(int64_t)r * 3] = cr; rgb_map[(int64_t)r * 3 + 1] = cg; rgb_map[(int64_t)r * 3 + 2] = cb; }
      if (acc_out) acc_out[r] = acc;
      if (depth_out) depth_out[r] = dep;
      if (disp) {
        const float q = dep / acc;
        disp[r] = 1.f / fmaxf(1e-10f, q);
        if (q != q) disp[r] = q;
      }
      if (pts_max) {
        const float* o = d - 3;
        pts_max[(int64_t)r * 3]     = __fadd_rn(__ldg(o),     __fmul_rn(dx, best_z));
        pts_max[(int64_t)r * 3 + 1] = __fadd_rn(__ldg(o + 1), __fmul_rn(dy, best_z));
        pts_max[(int64_t)r * 3 + 2] = __fadd_rn(__ldg(o + 2), __fmul_rn(dz, best_z));
      }
    }
  }
}

// Backward.  With G_i = dL/dw_i (all consumers of the weights folded in),
//   dL/dalpha_i = G_i T_i - (sum_{j>i} G_j w_j) / t_i          (cumprod backward, suffix-sum form)
//   dL/dsigma_i = dL/dalpha_i * dist_i * exp(-sigma_i dist_i)   masked by relu
//   dL/draw_rgb = w_i * g_rgb * c (1 - c)
template <int MAXC>
__global__ void __launch_bounds__(256)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     int ray_pitch, const float* __restrict__ noise, int R, int S, int white,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_disp,
                     const float* __restrict__ g_acc, const float* __restrict__ g_w,
                     const float* __restrict__ g_depth, float* __restrict__ g_raw) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nchunk = (S + 31) >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < R; r += gridDim.x * warps_per_block) {
    const float* d = rays_d + (int64_t)r * ray_pitch;
    const float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const int64_t base = (int64_t)r * S;
    float Tj[MAXC], aj[MAXC];
    float carry = 1.f, acc = 0.f, dep = 0.f;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      Tj[j] = 0.f; aj[j] = 0.f;
      if (j < nchunk) {
        const int i = (j << 5) + lane;
        const bool valid = i < S;
        SampleTerms s = load_sample(raw, z, noise, base, i, S, dnorm, valid);
        float incl = warp_scan_mul(s.t, lane);
        float excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = 1.f;
        const float T = carry * excl;
        carry *= __shfl_sync(FULL, incl, 31);
        Tj[j] = T; aj[j] = s.alpha;
        if (valid) { acc += s.alpha * T; dep += s.alpha * T * s.z; }
      }
    }
    acc = warp_sum(acc); dep = warp_sum(dep);

    float gr = 0.f, gg = 0.f, gb = 0.f;
    if (g_rgb) { gr = __ldg(g_rgb + (int64_t)r * 3); gg = __ldg(g_rgb + (int64_t)r * 3 + 1); gb = __ldg(g_rgb + (int64_t)r * 3 + 2); }
    float k_acc = g_acc ? __ldg(g_acc + r) : 0.f;
    float k_dep = g_depth ? __ldg(g_depth + r) : 0.f;
    if (white) k_acc -= (gr + gg + gb);
    if (g_disp) {
      const float gd = __ldg(g_disp + r);
      const float q = dep / acc;
      if (q > 1e-10f) {                       // max(1e-10, q) routes the gradient to q
        const float dq = -gd / (q * q);
        k_dep += dq / acc;
        k_acc += -dq * dep / (acc * acc);
      }
    }

    float suffix_carry = 0.f;
#pragma unroll
    for (int j = MAXC - 1; j >= 0; --j) {
      if (j < nchunk) {
        const int i = (j << 5) + lane;
        const bool valid = i < S;
        SampleTerms s = load_sample(raw, z, noise, base, i, S, dnorm, valid);   // L2/L1 hit: read in pass 1
        const float cr = sigmoidf(s.raw.x), cg = sigmoidf(s.raw.y), cb = sigmoidf(s.raw.z);
        const float T = Tj[j], a = aj[j];
        const float w = a * T;
        float G = gr * cr + gg * cg + gb * cb + k_dep * s.z + k_acc;
        if (g_w && valid) G += __ldg(g_w + base + i);
        float v = valid ? G * w : 0.f;
        float incl = warp_rscan_add(v, lane);
        float excl = __shfl_down_sync(FULL, incl, 1);
        if (lane == 31) excl = 0.f;
        const float suffix = suffix_carry + excl;
        suffix_carry += __shfl_sync(FULL, incl, 0);
        if (valid) {
          const float t = (1.f - a) + 1e-10f;
          const float dalpha = G * T - suffix / t;
          const float sig = fmaxf(s.sigma_in, 0.f);
          const float e = expf(-__fmul_rn(sig, s.dist));
          const float dsigma = (s.sigma_in > 0.f) ? dalpha * s.dist * e : 0.f;
          float4 o;
          o.x = w * gr * cr * (1.f - cr);
          o.y = w * gg * cg * (1.f - cg);
          o.z = w * gb * cb * (1.f - cb);
          o.w = dsigma;
          st_stream4(reinterpret_cast<float4*>(g_raw) + base + i, o);
        }
      }
    }
  }
}

static int composite_grid(int R) {
  const int warps_per_block = 8;
  int64_t blocks = ((int64_t)R + warps_per_block - 1) / warps_per_block;
  int64_t cap = (int64_t)sm_count() * 8;     // 8 resident 256-thread blocks / SM = 64 warps / SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace nfb

extern "C" {

int nfb_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, int ray_pitch,
                      const float* noise, int R, int S, int white_bkgd,
                      float* rgb_map, float* disp, float* acc, float* weights, float* depth,
                      float* pts_max, void* stream) {
  NFB_REQUIRE(R >= 0 && S > 0 && ray_pitch >= 3, "composite_fwd: R=%d S=%d ray_pitch=%d", R, S, ray_pitch);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(raw && z_vals && rays_d && weights, "composite_fwd: raw, z_vals, rays_d and weights are required");
  NFB_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "composite_fwd: raw must be 16-byte aligned");
  if (R == 0) return NFB_OK;
  const int grid = nfb::composite_grid(R);
  cudaStream_t st = (cudaStream_t)stream;
#define NFB_LAUNCH_FWD(NCH)                                                                                  \
  nfb::composite_fwd_unrolled_kernel<NCH><<<grid, 256, 0, st>>>(raw, z_vals, rays_d, ray_pitch, noise, R, S, \
      white_bkgd, rgb_map, disp, acc, weights, depth, pts_max)
  static const bool loop_form = []() { const char* e = getenv("NERFAIL_B200_COMPOSITE_LOOP"); return e && e[0] == '1'; }();
  if (loop_form || S > 256)
    nfb::composite_fwd_kernel<<<grid, 256, 0, st>>>(raw, z_vals, rays_d, ray_pitch, noise, R, S, white_bkgd, rgb_map, disp, acc,
                                                    weights, depth, pts_max);
  else if (S <= 64) NFB_LAUNCH_FWD(2);
  else if (S <= 128) NFB_LAUNCH_FWD(4);
  else if (S <= 192) NFB_LAUNCH_FWD(6);
  else NFB_LAUNCH_FWD(8);
#undef NFB_LAUNCH_FWD
  return nfb::check_launch("composite_fwd");
}

int nfb_composite_bwd(const float* raw, const float* z_vals, const float* rays_d, int ray_pitch,
                      const float* noise, int R, int S, int white_bkgd,
                      const float* g_rgb, const float* g_disp, const float* g_acc,
                      const float* g_weights, const float* g_depth, float* g_raw, void* stream) {
  NFB_REQUIRE(R >= 0 && S > 0 && ray_pitch >= 3, "composite_bwd: R=%d S=%d ray_pitch=%d", R, S, ray_pitch);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(raw && z_vals && rays_d && g_raw, "composite_bwd: raw, z_vals, rays_d and g_raw are required");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(g_raw)) & 15) == 0,
              "composite_bwd: raw and g_raw must be 16-byte aligned");
  if (S > 1024) return nfb::fail(NFB_E_UNSUPPORTED, "composite_bwd: S=%d > 1024 samples per ray", S);
  if (R == 0) return NFB_OK;
  const int grid = nfb::composite_grid(R);
  cudaStream_t st = (cudaStream_t)stream;
#define NFB_LAUNCH_BWD(MAXC)                                                                          \
  nfb::composite_bwd_kernel<MAXC><<<grid, 256, 0, st>>>(raw, z_vals, rays_d, ray_pitch, noise, R, S,  \
      white_bkgd, g_rgb, g_disp, g_acc, g_weights, g_depth, g_raw)
  if (S <= 64) NFB_LAUNCH_BWD(2);
  else if (S <= 256) NFB_LAUNCH_BWD(8);
  else NFB_LAUNCH_BWD(32);
#undef NFB_LAUNCH_BWD
  return nfb::check_launch("composite_bwd");
}

}  // extern "C"
