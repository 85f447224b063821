// GaussNet: Gaussian weights, 8-neighbour weighted gather (forward) and L2-reduction scatter (backward; per-lane
// red.global.add.v4.f32 by default, warp-aggregated variant opt-in because it measured slower — see the launcher).
//
// Reference arithmetic: model/GaussNet.py:169-186 (create_gauss_w.forward), :53-119 (gauss_net.forward up
// to x_rgba).  HBM/L2-bound: 228 B/pixel each way; the 30.7 MB [P*H*W,4] table is L2
// resident on B200, so the gathers and the red.global.add.v4.f32 scatters are L2 traffic and the
// streaming operands (weights, indices, original image, outputs) are read/written exactly once with
// 128-bit accesses.
#include "common.cuh"
#include <stdlib.h>

namespace nfb {

__global__ void __launch_bounds__(256)
gauss_weights_kernel(const float* __restrict__ di, int64_t B, int64_t HW, float c, float* __restrict__ out) {
  const int64_t n = B * HW;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, q = p % HW;
    const float4* dsrc = reinterpret_cast<const float4*>(di + ((b * 2 + 0) * HW + q) * 8);
    const float4* isrc = reinterpret_cast<const float4*>(di + ((b * 2 + 1) * HW + q) * 8);
    float4* wdst = reinterpret_cast<float4*>(out + ((b * 2 + 0) * HW + q) * 8);
    float4* idst = reinterpret_cast<float4*>(out + ((b * 2 + 1) * HW + q) * 8);
    const float4 d0 = ld_stream4(dsrc), d1 = ld_stream4(dsrc + 1);
    float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = __fdiv_rn(d[k], c);
      d[k] = expf(-__fdiv_rn(__fmul_rn(t, t), 2.f));        // exp(-(dist/c)^2 / 2)   :174-175
      s = __fadd_rn(s, d[k]);
    }
    const float dq = __fadd_rn(s, 0.001f);                   // :178
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = (s > 0.f) ? __fdiv_rn(d[k], dq) : 0.f;   // :181
    st_stream4(wdst, make_float4(d[0], d[1], d[2], d[3]));
    st_stream4(wdst + 1, make_float4(d[4], d[5], d[6], d[7]));
    st_stream4(idst, ld_stream4(isrc));
    st_stream4(idst + 1, ld_stream4(isrc + 1));
  }
}

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {   // valid for any sign mix
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct PixelIn {
  float w[8];
  int idx[8];
  float ori[4];
};

__device__ __forceinline__ PixelIn load_pixel(const float* __restrict__ w_idx, const uint8_t* __restrict__ ori,
                                              int64_t b, int64_t q, int64_t HW) {
  PixelIn px;
  const float4* ws = reinterpret_cast<const float4*>(w_idx + ((b * 2 + 0) * HW + q) * 8);
  const float4* is = reinterpret_cast<const float4*>(w_idx + ((b * 2 + 1) * HW + q) * 8);
  const float4 w0 = ld_stream4(ws), w1 = ld_stream4(ws + 1), i0 = ld_stream4(is), i1 = ld_stream4(is + 1);
  px.w[0] = w0.x; px.w[1] = w0.y; px.w[2] = w0.z; px.w[3] = w0.w;
  px.w[4] = w1.x; px.w[5] = w1.y; px.w[6] = w1.z; px.w[7] = w1.w;
  px.idx[0] = (int)i0.x; px.idx[1] = (int)i0.y; px.idx[2] = (int)i0.z; px.idx[3] = (int)i0.w;   // .type(torch.long) :62
  px.idx[4] = (int)i1.x; px.idx[5] = (int)i1.y; px.idx[6] = (int)i1.z; px.idx[7] = (int)i1.w;
  const uchar4 o = *reinterpret_cast<const uchar4*>(ori + (b * HW + q) * 4);
  px.ori[0] = (float)o.x; px.ori[1] = (float)o.y; px.ori[2] = (float)o.z; px.ori[3] = (float)o.w;
  return px;
}

__global__ void __launch_bounds__(256)
gauss_gather_fwd_kernel(const float4* __restrict__ table, int64_t T, const float* __restrict__ w_idx,
                        const uint8_t* __restrict__ ori, int64_t B, int64_t HW, float eps,
                        float4* __restrict__ x_out, float4* __restrict__ xrgba_out, float* __restrict__ minmax) {
  const int64_t n = B * HW;
  float vmin = 0.f, vmax = 0.f;
  bool any = false;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, q = p % HW;
    const PixelIn px = load_pixel(w_idx, ori, b, q, HW);
    float4 rows[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {      // issue all 8 gathers before the first use
      int id = px.idx[k];
      id = id < 0 ? 0 : (id >= T ? (int)(T - 1) : id);
      rows[k] = __ldg(table + id);
    }
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {      // (x * w).sum(-2), k ascending   :81-83
      x.x = __fadd_rn(x.x, __fmul_rn(rows[k].x, px.w[k]));
      x.y = __fadd_rn(x.y, __fmul_rn(rows[k].y, px.w[k]));
      x.z = __fadd_rn(x.z, __fmul_rn(rows[k].z, px.w[k]));
      x.w = __fadd_rn(x.w, __fmul_rn(rows[k].w, px.w[k]));
    }
    const float alpha = __fdiv_rn(x.w, 255.f);                                   // :85
    float pr[3] = {__fmul_rn(x.x, alpha), __fmul_rn(x.y, alpha), __fmul_rn(x.z, alpha)};
    if (minmax) {                                                                // :89-103 without the host sync
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float v = (alpha > 0.f) ? pr[ch] : 0.f;
        if (!any) { vmin = v; vmax = v; any = true; }
        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
      }
    }
    float4 out;
    float* o = &out.x;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float v = pr[ch];
      if (eps >= 0.f) v = fminf(fmaxf(v, -eps), eps);                            // :109
      v = __fadd_rn(px.ori[ch], v);                                              // :107 / :110
      if (!(px.ori[3] > 0.f)) v = 0.f;                                           // :112-113
      o[ch] = fminf(fmaxf(v, 0.f), 255.f);                                       // :119
    }
    out.w = fminf(fmaxf(px.ori[3], 0.f), 255.f);
    st_stream4(x_out + p, x);
    st_stream4(xrgba_out + p, out);
  }
  if (minmax) {
    unsigned have = __ballot_sync(FULL, any);
    if (have) {
      const int src = __ffs(have) - 1;
      const float fill_min = __shfl_sync(FULL, vmin, src), fill_max = __shfl_sync(FULL, vmax, src);
      if (!any) { vmin = fill_min; vmax = fill_max; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(FULL, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
      }
      if ((threadIdx.x & 31) == 0) { atomic_min_f(minmax, vmin); atomic_max_f(minmax + 1, vmax); }
    }
  }
}

__device__ __forceinline__ void red_add_v4(float4* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One thread per pixel computes G = dL/dx (4 channels); the 8 contributions w_k * G go to table rows idx_k as
// 128-bit L2 reductions.  AGGREGATE = true first merges lanes of the warp that hit the same row at the same k
// (match.any) so that only the lowest lane issues the reduction; kept selectable, off by default (see the launcher).
template <bool AGGREGATE>
__global__ void __launch_bounds__(256)
gauss_scatter_bwd_kernel(const float4* __restrict__ g_x, const float4* __restrict__ g_xrgba,
                         const float4* __restrict__ x_saved, const float* __restrict__ w_idx,
                         const uint8_t* __restrict__ ori, int64_t B, int64_t HW, float eps, int64_t T,
                         float4* __restrict__ g_table) {
  const int64_t n = B * HW;
  const int lane = threadIdx.x & 31;
  const int64_t n_round = (n + 31) & ~(int64_t)31;     // whole warps stay convergent for the shuffles
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_round; p += (int64_t)gridDim.x * blockDim.x) {
    const bool live = p < n;
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
    PixelIn px;
#pragma unroll
    for (int k = 0; k < 8; ++k) { px.w[k] = 0.f; px.idx[k] = -1 - lane; }
    if (live) {
      const int64_t b = p / HW, q = p % HW;
      px = load_pixel(w_idx, ori, b, q, HW);
      if (g_x) G = ld_stream4(g_x + p);
      if (g_xrgba) {
        const float4 go = ld_stream4(g_xrgba + p);
        const float4 x = ld_stream4(x_saved + p);
        const float alpha = __fdiv_rn(x.w, 255.f);
        const float xs[3] = {x.x, x.y, x.z};
        const float gs[3] = {go.x, go.y, go.z};
        float* Gp = &G.x;
        float g_alpha = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float pr = __fmul_rn(xs[ch], alpha);
          float v = pr;
          bool pass = true;
          if (eps >= 0.f) { pass = (pr >= -eps) && (pr <= eps); v = fminf(fmaxf(pr, -eps), eps); }
          v = __fadd_rn(px.ori[ch], v);
          pass = pass && (px.ori[3] > 0.f) && (v >= 0.f) && (v <= 255.f);   // torch.clip passes grad on the closed interval
          if (pass) { Gp[ch] += gs[ch] * alpha; g_alpha += gs[ch] * xs[ch]; }
        }
        G.w += g_alpha / 255.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int id = px.idx[k];
      if (live) id = id < 0 ? 0 : (id >= T ? (int)(T - 1) : id);
      float4 v = make_float4(G.x * px.w[k], G.y * px.w[k], G.z * px.w[k], G.w * px.w[k]);
      if (AGGREGATE) {
        const unsigned peers = __match_any_sync(FULL, id);
        const int leader = __ffs(peers) - 1;
        unsigned rem = peers & ~(1u << leader);
        while (__any_sync(FULL, rem != 0)) {
          const int src = rem ? (__ffs(rem) - 1) : lane;
          const float ox = __shfl_sync(FULL, v.x, src), oy = __shfl_sync(FULL, v.y, src);
          const float oz = __shfl_sync(FULL, v.z, src), ow = __shfl_sync(FULL, v.w, src);
          if (lane == leader && rem) { v.x += ox; v.y += oy; v.z += oz; v.w += ow; }
          rem &= rem - 1;
        }
        if (live && lane == leader) red_add_v4(g_table + id, v);
      } else {
        if (live) red_add_v4(g_table + id, v);
      }
    }
  }
}

// The same backward for NC cotangents of x_rgba at once (DeepFool's per-class gradients, deepfool.py:72-86: 14
// torch.autograd.grad calls per iteration through the same forward): weights, indices, original pixel and the saved x are
// read ONCE per pixel, the clip / alpha masks are evaluated once, and the NC gradient tables [NC][T][4] receive their
// reductions from one launch.  g_xrgba [NC][B*HW][4]; g_x (the gradient w.r.t. the un-composited x) is not an input: the
// classifier sees x_rgba only.
__global__ void __launch_bounds__(256)
gauss_scatter_bwd_batched_kernel(const float4* __restrict__ g_xrgba, int NC, const float4* __restrict__ x_saved,
                                 const float* __restrict__ w_idx, const uint8_t* __restrict__ ori, int64_t B, int64_t HW,
                                 float eps, int64_t T, float4* __restrict__ g_table) {
  const int64_t n = B * HW;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, q = p % HW;
    PixelIn px = load_pixel(w_idx, ori, b, q, HW);
    const float4 x = ld_stream4(x_saved + p);
    const float alpha = __fdiv_rn(x.w, 255.f);
    const float xs[3] = {x.x, x.y, x.z};
    bool pass[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float pr = __fmul_rn(xs[ch], alpha);
      float v = pr;
      bool ok = true;
      if (eps >= 0.f) { ok = (pr >= -eps) && (pr <= eps); v = fminf(fmaxf(pr, -eps), eps); }
      v = __fadd_rn(px.ori[ch], v);
      pass[ch] = ok && (px.ori[3] > 0.f) && (v >= 0.f) && (v <= 255.f);
    }
    if (!(pass[0] || pass[1] || pass[2])) continue;           // this pixel passes no gradient for any cotangent
#pragma unroll
    for (int k = 0; k < 8; ++k) px.idx[k] = px.idx[k] < 0 ? 0 : (px.idx[k] >= T ? (int)(T - 1) : px.idx[k]);
    for (int c = 0; c < NC; ++c) {
      const float4 go = ld_stream4(g_xrgba + (int64_t)c * n + p);
      const float gs[3] = {go.x, go.y, go.z};
      float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
      float* Gp = &G.x;
      float g_alpha = 0.f;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        if (pass[ch]) { Gp[ch] = gs[ch] * alpha; g_alpha += gs[ch] * xs[ch]; }
      G.w = g_alpha / 255.f;
      float4* dst = g_table + (int64_t)c * T;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        red_add_v4(dst + px.idx[k], make_float4(G.x * px.w[k], G.y * px.w[k], G.z * px.w[k], G.w * px.w[k]));
    }
  }
}

static int stream_grid(int64_t items) {
  int64_t blocks = (items + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace nfb

namespace nfb {

// RGBA [B,HW,4] -> planar RGB [B,3,HW] on a white (fill) background: model/GaussNet.py:121-145
//   cla = where(alpha > 0, rgb, 255) after the NHWC -> NCHW transpose (alpha = channel 3 of the same image).
// One thread per pixel: a float4 (or uchar4) read, three coalesced plane writes.
template <typename Pix>
__global__ void __launch_bounds__(256)
rgba_to_chw_kernel(const Pix* __restrict__ img, const float4* __restrict__ alpha_src, int64_t B, int64_t HW, float fill,
                   float* __restrict__ out) {
  const int64_t n = B * HW;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, q = p % HW;
    const Pix v = img[p];
    const float a = alpha_src ? alpha_src[p].w : (float)v.w;
    const bool on = a > 0.f;
    float* o = out + b * 3 * HW + q;
    o[0] = on ? (float)v.x : fill;
    o[HW] = on ? (float)v.y : fill;
    o[2 * HW] = on ? (float)v.z : fill;
  }
}

// adjoint of the above with fill = 0: g_img[b,q,c] = alpha > 0 ? g_out[b,c,q] : 0 (c < 3), g_img[b,q,3] = 0
__global__ void __launch_bounds__(256)
chw_to_rgba_kernel(const float* __restrict__ g_out, const float4* __restrict__ alpha_src, int64_t B, int64_t HW,
                   float4* __restrict__ g_img) {
  const int64_t n = B * HW;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / HW, q = p % HW;
    const bool on = alpha_src[p].w > 0.f;
    const float* g = g_out + b * 3 * HW + q;
    g_img[p] = on ? make_float4(g[0], g[HW], g[2 * HW], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace nfb

namespace nfb {

// The RGB columns of the active rows of grad_spatial_rgb, packed: out[k] = grad[idx[k]].xyz  (what the sign step consumes)
__global__ void __launch_bounds__(256)
attack_pack_rgb_kernel(const float4* __restrict__ grad, const int64_t* __restrict__ idx, int64_t n, float* __restrict__ out) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = __ldg(grad + __ldg(idx + k));
    out[3 * k] = g.x; out[3 * k + 1] = g.y; out[3 * k + 2] = g.z;
  }
}

// attack_NeRFail_S.py:357-392 on the active rows: rgb <- clamp(rgb -/+ step * sign(g), init - eps, init + eps)
__global__ void __launch_bounds__(256)
attack_sign_step_kernel(float4* __restrict__ table, const float4* __restrict__ init, const int64_t* __restrict__ idx,
                        const float* __restrict__ g, int64_t n, float step, float eps) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = __ldg(idx + k);
    float4 t = table[r];
    const float4 t0 = __ldg(init + r);
    const float gx = g[3 * k], gy = g[3 * k + 1], gz = g[3 * k + 2];
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };     // torch.sign
    t.x = fmaxf(fminf(t.x - step * sgn(gx), t0.x + eps), t0.x - eps);
    t.y = fmaxf(fminf(t.y - step * sgn(gy), t0.y + eps), t0.y - eps);
    t.z = fmaxf(fminf(t.z - step * sgn(gz), t0.z + eps), t0.z - eps);
    table[r] = t;
  }
}

}  // namespace nfb

extern "C" {

int nfb_gauss_weights(const float* dist_idx, int64_t B, int64_t HW, float c, float* i_w, void* stream) {
  NFB_REQUIRE(dist_idx && i_w && B >= 0 && HW >= 0 && c != 0.f, "gauss_weights: bad argument");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(dist_idx) | reinterpret_cast<uintptr_t>(i_w)) & 15) == 0,
              "gauss_weights: buffers must be 16-byte aligned");
  if (B * HW == 0) return NFB_OK;
  nfb::gauss_weights_kernel<<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(dist_idx, B, HW, c, i_w);
  return nfb::check_launch("gauss_weights");
}

int nfb_gauss_gather_fwd(const float* table, int64_t T, const float* w_idx, const uint8_t* ori,
                         int64_t B, int64_t HW, float eps, float* x, float* x_rgba, float* minmax, void* stream) {
  NFB_REQUIRE(table && w_idx && ori && x && x_rgba, "gauss_gather_fwd: null pointer");
  NFB_REQUIRE(T > 0 && T < (1 << 24) + 1 && B >= 0 && HW >= 0, "gauss_gather_fwd: T=%lld B=%lld HW=%lld", (long long)T, (long long)B, (long long)HW);
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(w_idx) | reinterpret_cast<uintptr_t>(x) |
                reinterpret_cast<uintptr_t>(x_rgba)) & 15) == 0 && (reinterpret_cast<uintptr_t>(ori) & 3) == 0,
              "gauss_gather_fwd: buffers must be 16-byte aligned (ori: 4)");
  if (B * HW == 0) return NFB_OK;
  nfb::gauss_gather_fwd_kernel<<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(table), T, w_idx, ori, B, HW, eps,
      reinterpret_cast<float4*>(x), reinterpret_cast<float4*>(x_rgba), minmax);
  return nfb::check_launch("gauss_gather_fwd");
}

int nfb_gauss_scatter_bwd(const float* g_x, const float* g_xrgba, const float* x, const float* w_idx,
                          const uint8_t* ori, int64_t B, int64_t HW, float eps, int64_t T,
                          float* g_table, void* stream) {
  NFB_REQUIRE(w_idx && ori && g_table && (g_x || g_xrgba), "gauss_scatter_bwd: null pointer");
  NFB_REQUIRE(!g_xrgba || x, "gauss_scatter_bwd: x (saved forward output) is required with g_xrgba");
  NFB_REQUIRE(T > 0 && B >= 0 && HW >= 0, "gauss_scatter_bwd: bad size");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(g_table) | reinterpret_cast<uintptr_t>(w_idx) | reinterpret_cast<uintptr_t>(g_x) |
                reinterpret_cast<uintptr_t>(g_xrgba) | reinterpret_cast<uintptr_t>(x)) & 15) == 0,
              "gauss_scatter_bwd: buffers must be 16-byte aligned");
  if (B * HW == 0) return NFB_OK;
  // Default: every lane issues its own red.global.add.v4.f32.  NERFAIL_B200_SCATTER_AGG=1 enables the warp-level
  // merge of equal rows (match.any + shuffle tree).  Measured on B200 at 800x800, P=3 (scripts/profile_gauss.py):
  // aggregated 64-73 us per view vs 42-47 us plain for clustered, 3x3-shared and uniformly random indices alike -
  // the L2 atomic units absorb the duplicates faster than 8 match.any rounds per pixel can remove them.
  static const bool agg = []() { const char* e = getenv("NERFAIL_B200_SCATTER_AGG"); return e && e[0] == '1'; }();
  if (agg)
    nfb::gauss_scatter_bwd_kernel<true><<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(g_x), reinterpret_cast<const float4*>(g_xrgba),
        reinterpret_cast<const float4*>(x), w_idx, ori, B, HW, eps, T, reinterpret_cast<float4*>(g_table));
  else
    nfb::gauss_scatter_bwd_kernel<false><<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(g_x), reinterpret_cast<const float4*>(g_xrgba),
        reinterpret_cast<const float4*>(x), w_idx, ori, B, HW, eps, T, reinterpret_cast<float4*>(g_table));
  return nfb::check_launch("gauss_scatter_bwd");
}

int nfb_gauss_scatter_bwd_batched(const float* g_xrgba, int NC, const float* x, const float* w_idx, const uint8_t* ori,
                                  int64_t B, int64_t HW, float eps, int64_t T, float* g_table, void* stream) {
  NFB_REQUIRE(g_xrgba && x && w_idx && ori && g_table, "gauss_scatter_bwd_batched: null pointer");
  NFB_REQUIRE(NC >= 0 && T > 0 && B >= 0 && HW >= 0, "gauss_scatter_bwd_batched: bad size");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(g_table) | reinterpret_cast<uintptr_t>(w_idx) | reinterpret_cast<uintptr_t>(g_xrgba) |
                reinterpret_cast<uintptr_t>(x)) & 15) == 0, "gauss_scatter_bwd_batched: buffers must be 16-byte aligned");
  if (B * HW == 0 || NC == 0) return NFB_OK;
  nfb::gauss_scatter_bwd_batched_kernel<<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(g_xrgba), NC, reinterpret_cast<const float4*>(x), w_idx, ori, B, HW, eps, T,
      reinterpret_cast<float4*>(g_table));
  return nfb::check_launch("gauss_scatter_bwd_batched");
}

int nfb_rgba_to_chw(const float* img_f32, const uint8_t* img_u8, const float* alpha_src, int64_t B, int64_t HW, float fill,
                    float* out, void* stream) {
  NFB_REQUIRE((img_f32 != nullptr) != (img_u8 != nullptr) && out, "rgba_to_chw: exactly one of img_f32 / img_u8, and out");
  NFB_REQUIRE(B >= 0 && HW >= 0, "rgba_to_chw: B=%lld HW=%lld", (long long)B, (long long)HW);
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(img_f32) | reinterpret_cast<uintptr_t>(alpha_src)) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(img_u8) & 3) == 0, "rgba_to_chw: images must be 16-byte aligned (uint8: 4)");
  if (B * HW == 0) return NFB_OK;
  if (img_f32)
    nfb::rgba_to_chw_kernel<float4><<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(img_f32), reinterpret_cast<const float4*>(alpha_src), B, HW, fill, out);
  else
    nfb::rgba_to_chw_kernel<uchar4><<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uchar4*>(img_u8), reinterpret_cast<const float4*>(alpha_src), B, HW, fill, out);
  return nfb::check_launch("rgba_to_chw");
}

int nfb_chw_to_rgba(const float* g_out, const float* alpha_src, int64_t B, int64_t HW, float* g_img, void* stream) {
  NFB_REQUIRE(g_out && alpha_src && g_img, "chw_to_rgba: null pointer");
  NFB_REQUIRE(B >= 0 && HW >= 0, "chw_to_rgba: B=%lld HW=%lld", (long long)B, (long long)HW);
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(alpha_src) | reinterpret_cast<uintptr_t>(g_img)) & 15) == 0,
              "chw_to_rgba: images must be 16-byte aligned");
  if (B * HW == 0) return NFB_OK;
  nfb::chw_to_rgba_kernel<<<nfb::stream_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
      g_out, reinterpret_cast<const float4*>(alpha_src), B, HW, reinterpret_cast<float4*>(g_img));
  return nfb::check_launch("chw_to_rgba");
}

int nfb_attack_pack_rgb(const float* grad, const int64_t* active_idx, int64_t n, float* packed, void* stream) {
  NFB_REQUIRE(grad && active_idx && packed && n >= 0, "attack_pack_rgb: bad argument");
  NFB_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "attack_pack_rgb: grad must be 16-byte aligned");
  if (n == 0) return NFB_OK;
  nfb::attack_pack_rgb_kernel<<<nfb::stream_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(grad), active_idx, n, packed);
  return nfb::check_launch("attack_pack_rgb");
}

int nfb_attack_sign_step(float* table, const float* init, const int64_t* active_idx, const float* packed_grad, int64_t n,
                         float signed_step, float eps, void* stream) {
  NFB_REQUIRE(table && init && active_idx && packed_grad && n >= 0 && eps >= 0.f, "attack_sign_step: bad argument");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(init)) & 15) == 0, "attack_sign_step: tables must be 16-byte aligned");
  if (n == 0) return NFB_OK;
  nfb::attack_sign_step_kernel<<<nfb::stream_grid(n), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(table), reinterpret_cast<const float4*>(init), active_idx, packed_grad, n, signed_step, eps);
  return nfb::check_launch("attack_sign_step");
}

}  // extern "C"
