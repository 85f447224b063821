// Classifier-input stage of GaussNet fused into one kernel: RGBA -> RGB on white -> NCHW -> bilinear Resize.
//
// Reference arithmetic: model/GaussNet.py:121-145 (where(alpha > 0, rgb, 255) after the NHWC -> NCHW transpose) followed by
// :147-154, torchvision.transforms.Resize([299, 299]) (or 224 for vit_b_16) on the float NCHW tensor, i.e. ATen's
// upsample_bilinear2d (antialias = False: two taps per axis, align_corners = False) or _upsample_bilinear2d_aa
// (antialias = True — the default of the torchvision the container ships; a triangle filter whose support grows with the
// down-scaling factor, ~7 taps per axis for 800 -> 299).  Both are separable linear maps: per axis and output index a start,
// a tap count and normalised weights (nfb_resize_weights computes them on the host exactly as ATen's
// area_pixel_compute_source_index / _compute_indices_min_size_weights_aa do, in fp32).
// The adjoint (gradient w.r.t. the RGBA image) uses the transposed tables (per INPUT index the contiguous range of outputs
// it feeds) so that it is a gather as well: no atomics, deterministic.  Forward and adjoint are each other's derivative
// (both linear), which keeps gauss_net differentiable twice for deepfool.py:76-77.
// HBM/L2-bound and small next to a classifier: 10.2 MB in, 1.07 MB out per 800x800 -> 299x299 image.
#include "common.cuh"
#include <math.h>
#include <vector>

namespace nfb {

struct AxisTable {          // device pointers
  const int* start;         // [n]
  const int* count;         // [n]
  const float* w;           // [n][maxk]
  int maxk;
};

template <typename Pix>
__global__ void __launch_bounds__(256)
resize_rgba_to_chw_kernel(const Pix* __restrict__ img, const float4* __restrict__ alpha_src, int64_t alpha_batch,
                          int64_t B, int H, int W, int OH, int OW, float fill, AxisTable ty, AxisTable tx,
                          float* __restrict__ out) {
  const int64_t n = B * OH * OW;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(p % OW), oy = (int)((p / OW) % OH);
    const int64_t b = p / ((int64_t)OW * OH);
    const int y0 = __ldg(ty.start + oy), ny = __ldg(ty.count + oy), x0 = __ldg(tx.start + ox), nx = __ldg(tx.count + ox);
    const Pix* src = img + b * (int64_t)H * W;
    const float4* asrc = alpha_src ? alpha_src + (b % alpha_batch) * (int64_t)H * W : nullptr;
    float r = 0.f, g = 0.f, bl = 0.f;
    for (int j = 0; j < ny; ++j) {
      const float wy = __ldg(ty.w + (int64_t)oy * ty.maxk + j);
      float rr = 0.f, gg = 0.f, bb = 0.f;                    // horizontal pass of this row first, like ATen's separable form
      const int64_t rowoff = (int64_t)(y0 + j) * W + x0;
      for (int i = 0; i < nx; ++i) {
        const float wx = __ldg(tx.w + (int64_t)ox * tx.maxk + i);
        const Pix v = src[rowoff + i];
        const float a = asrc ? asrc[rowoff + i].w : (float)v.w;
        const bool on = a > 0.f;
        rr = fmaf(wx, on ? (float)v.x : fill, rr);
        gg = fmaf(wx, on ? (float)v.y : fill, gg);
        bb = fmaf(wx, on ? (float)v.z : fill, bb);
      }
      r = fmaf(wy, rr, r); g = fmaf(wy, gg, g); bl = fmaf(wy, bb, bl);
    }
    const int64_t plane = (int64_t)OH * OW;
    float* o = out + b * 3 * plane + (int64_t)oy * OW + ox;
    o[0] = r; o[plane] = g; o[2 * plane] = bl;
  }
}

// g_img[b, y, x, c] = alpha > 0 ? sum_{oy in Ty(y)} sum_{ox in Tx(x)} wy * wx * g_out[b, c, oy, ox] : 0 ; channel 3 = 0
__global__ void __launch_bounds__(256)
resize_chw_to_rgba_adjoint_kernel(const float* __restrict__ g_out, const float4* __restrict__ alpha_src, int64_t alpha_batch,
                                  int64_t B, int H, int W, int OH, int OW, AxisTable ty, AxisTable tx,
                                  float4* __restrict__ g_img) {
  const int64_t n = B * H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(p % W), y = (int)((p / W) % H);
    const int64_t b = p / ((int64_t)W * H);
    const float a = alpha_src[(b % alpha_batch) * (int64_t)H * W + (int64_t)y * W + x].w;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a > 0.f) {
      const int oy0 = __ldg(ty.start + y), ny = __ldg(ty.count + y), ox0 = __ldg(tx.start + x), nx = __ldg(tx.count + x);
      const int64_t plane = (int64_t)OH * OW;
      const float* g = g_out + b * 3 * plane;
      for (int j = 0; j < ny; ++j) {
        const float wy = __ldg(ty.w + (int64_t)y * ty.maxk + j);
        const float* row = g + (int64_t)(oy0 + j) * OW + ox0;
        float rr = 0.f, gg = 0.f, bb = 0.f;
        for (int i = 0; i < nx; ++i) {
          const float wx = __ldg(tx.w + (int64_t)x * tx.maxk + i);
          rr = fmaf(wx, __ldg(row + i), rr);
          gg = fmaf(wx, __ldg(row + plane + i), gg);
          bb = fmaf(wx, __ldg(row + 2 * plane + i), bb);
        }
        acc.x = fmaf(wy, rr, acc.x); acc.y = fmaf(wy, gg, acc.y); acc.z = fmaf(wy, bb, acc.z);
      }
    }
    g_img[p] = acc;
  }
}

static int resize_grid(int64_t items) {
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// One axis of ATen's resampling, fp32 like the float kernels.  Returns the largest tap count.
static int axis_weights(int in, int out, bool antialias, std::vector<int>& start, std::vector<int>& count,
                        std::vector<float>& w, int maxk) {
  const float scale = (float)in / (float)out;              // area_pixel_compute_scale, align_corners = False, no scale_factor
  int worst = 0;
  for (int i = 0; i < out; ++i) {
    float* wi = w.data() + (size_t)i * maxk;
    if (!antialias) {
      float src = scale * ((float)i + 0.5f) - 0.5f;        // area_pixel_compute_source_index (bilinear: clamped at 0)
      if (src < 0.f) src = 0.f;
      int i0 = (int)src;
      if (i0 > in - 1) i0 = in - 1;
      const int i1 = i0 + ((i0 < in - 1) ? 1 : 0);
      const float l1 = src - (float)i0, l0 = 1.f - l1;
      start[i] = i0;
      if (i1 == i0) { count[i] = 1; wi[0] = l0 + l1; }
      else { count[i] = 2; wi[0] = l0; wi[1] = l1; }
    } else {                                               // _compute_indices_min_size_weights_aa, triangle filter
      const float support = (scale >= 1.f) ? scale : 1.f;  // (interp_size / 2) * scale with interp_size = 2
      const float center = scale * ((float)i + 0.5f);
      const float invscale = (scale >= 1.f) ? 1.f / scale : 1.f;
      int xmin = (int)(center - support + 0.5f);
      if (xmin < 0) xmin = 0;
      int xmax = (int)(center + support + 0.5f);
      if (xmax > in) xmax = in;
      int xsize = xmax - xmin;
      if (xsize > maxk) xsize = maxk;
      float total = 0.f;
      for (int j = 0; j < xsize; ++j) {
        float t = ((float)(j + xmin) - center + 0.5f) * invscale;
        t = fabsf(t);
        const float v = t < 1.f ? 1.f - t : 0.f;
        wi[j] = v; total += v;
      }
      for (int j = 0; j < xsize; ++j) wi[j] = total != 0.f ? wi[j] / total : 0.f;
      start[i] = xmin; count[i] = xsize;
    }
    if (count[i] > worst) worst = count[i];
  }
  return worst;
}

}  // namespace nfb

extern "C" {

int nfb_resize_max_taps(int in_size, int out_size, int antialias, int transposed) {
  if (in_size <= 0 || out_size <= 0) return 0;
  const double scale = (double)in_size / (double)out_size;
  const int fwd = antialias ? (int)ceil(2.0 * (scale >= 1.0 ? scale : 1.0)) + 2 : 2;
  if (!transposed) return fwd;
  // outputs fed by one input index: the taps of neighbouring outputs overlap it for about fwd / scale outputs
  return (int)ceil((double)fwd / (scale < 1e-9 ? 1e-9 : scale)) + 3;
}

int nfb_resize_weights(int in_size, int out_size, int antialias, int transposed, int maxk,
                       int* start_host, int* count_host, float* weights_host) {
  NFB_REQUIRE(in_size > 0 && out_size > 0 && start_host && count_host && weights_host, "resize_weights: bad argument");
  NFB_REQUIRE(maxk >= nfb_resize_max_taps(in_size, out_size, antialias, transposed), "resize_weights: maxk %d too small", maxk);
  const int kf = nfb_resize_max_taps(in_size, out_size, antialias, 0);
  std::vector<int> st(out_size), ct(out_size);
  std::vector<float> w((size_t)out_size * kf, 0.f);
  nfb::axis_weights(in_size, out_size, antialias != 0, st, ct, w, kf);
  if (!transposed) {
    for (int i = 0; i < out_size; ++i) {
      start_host[i] = st[i]; count_host[i] = ct[i];
      for (int j = 0; j < maxk; ++j) weights_host[(size_t)i * maxk + j] = j < ct[i] ? w[(size_t)i * kf + j] : 0.f;
    }
    return NFB_OK;
  }
  // transposed: for input index x the outputs o with st[o] <= x < st[o] + ct[o] (a contiguous range: st and st + ct are
  // non-decreasing in o), weight w[o][x - st[o]]
  for (int x = 0; x < in_size; ++x) { start_host[x] = 0; count_host[x] = 0; }
  for (size_t k = 0; k < (size_t)in_size * maxk; ++k) weights_host[k] = 0.f;
  for (int o = 0; o < out_size; ++o)
    for (int j = 0; j < ct[o]; ++j) {
      const int x = st[o] + j;
      if (count_host[x] == 0) start_host[x] = o;
      const int slot = o - start_host[x];
      if (slot >= maxk) return nfb::fail(NFB_E_ARG, "resize_weights: transposed tap count exceeds maxk %d", maxk);
      weights_host[(size_t)x * maxk + slot] = w[(size_t)o * kf + j];
      if (slot + 1 > count_host[x]) count_host[x] = slot + 1;
    }
  return NFB_OK;
}

int nfb_rgba_to_chw_resized(const float* img_f32, const uint8_t* img_u8, const float* alpha_src, int64_t alpha_batch,
                            int64_t B, int H, int W, int OH, int OW, float fill,
                            const int* y_start, const int* y_count, const float* y_w, int y_maxk,
                            const int* x_start, const int* x_count, const float* x_w, int x_maxk,
                            float* out, void* stream) {
  NFB_REQUIRE((img_f32 != nullptr) != (img_u8 != nullptr) && out, "rgba_to_chw_resized: exactly one of img_f32 / img_u8, and out");
  NFB_REQUIRE(B >= 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "rgba_to_chw_resized: bad size");
  NFB_REQUIRE(y_start && y_count && y_w && x_start && x_count && x_w && y_maxk > 0 && x_maxk > 0, "rgba_to_chw_resized: tables missing");
  NFB_REQUIRE(!alpha_src || alpha_batch > 0, "rgba_to_chw_resized: alpha_batch must be positive with alpha_src");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(img_f32) | reinterpret_cast<uintptr_t>(alpha_src)) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(img_u8) & 3) == 0, "rgba_to_chw_resized: images must be 16-byte aligned (uint8: 4)");
  if (B == 0) return NFB_OK;
  const nfb::AxisTable ty{y_start, y_count, y_w, y_maxk}, tx{x_start, x_count, x_w, x_maxk};
  const int grid = nfb::resize_grid(B * OH * OW);
  if (img_f32)
    nfb::resize_rgba_to_chw_kernel<float4><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(img_f32), reinterpret_cast<const float4*>(alpha_src), alpha_batch > 0 ? alpha_batch : 1,
        B, H, W, OH, OW, fill, ty, tx, out);
  else
    nfb::resize_rgba_to_chw_kernel<uchar4><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uchar4*>(img_u8), reinterpret_cast<const float4*>(alpha_src), alpha_batch > 0 ? alpha_batch : 1,
        B, H, W, OH, OW, fill, ty, tx, out);
  return nfb::check_launch("rgba_to_chw_resized");
}

int nfb_chw_resized_to_rgba(const float* g_out, const float* alpha_src, int64_t alpha_batch, int64_t B, int H, int W,
                            int OH, int OW,
                            const int* yt_start, const int* yt_count, const float* yt_w, int yt_maxk,
                            const int* xt_start, const int* xt_count, const float* xt_w, int xt_maxk,
                            float* g_img, void* stream) {
  NFB_REQUIRE(g_out && alpha_src && g_img && alpha_batch > 0, "chw_resized_to_rgba: null pointer");
  NFB_REQUIRE(B >= 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "chw_resized_to_rgba: bad size");
  NFB_REQUIRE(yt_start && yt_count && yt_w && xt_start && xt_count && xt_w && yt_maxk > 0 && xt_maxk > 0, "chw_resized_to_rgba: tables missing");
  NFB_REQUIRE(((reinterpret_cast<uintptr_t>(alpha_src) | reinterpret_cast<uintptr_t>(g_img)) & 15) == 0,
              "chw_resized_to_rgba: images must be 16-byte aligned");
  if (B == 0) return NFB_OK;
  const nfb::AxisTable ty{yt_start, yt_count, yt_w, yt_maxk}, tx{xt_start, xt_count, xt_w, xt_maxk};
  nfb::resize_chw_to_rgba_adjoint_kernel<<<nfb::resize_grid(B * H * W), 256, 0, (cudaStream_t)stream>>>(
      g_out, reinterpret_cast<const float4*>(alpha_src), alpha_batch, B, H, W, OH, OW, ty, tx, reinterpret_cast<float4*>(g_img));
  return nfb::check_launch("chw_resized_to_rgba");
}

}  // extern "C"
