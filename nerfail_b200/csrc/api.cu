// Error plumbing, launch accounting and device queries behind the C ABI.
#include "common.cuh"

namespace nfb {

std::atomic<uint64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace nfb

extern "C" {

int nfb_abi_version(void) { return NFB_ABI_VERSION; }

const char* nfb_last_error(void) { return nfb::err_buf(); }

uint64_t nfb_launch_count(void) { return nfb::g_launches.load(std::memory_order_relaxed); }

int nfb_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  NFB_CUDA(cudaGetDevice(&dev));
  NFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  NFB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

// ---- one ray batch, coarse + fine, without autograd: the kernel sequence of render_rays in one call ----
static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

size_t nfb_render_rays_workspace_bytes(int R, int N_samples, int N_importance) {
  if (R <= 0 || N_samples <= 0 || N_importance < 0) return 0;
  const size_t Sc = (size_t)N_samples, Sf = Sc + (size_t)N_importance, r = (size_t)R;
  // z_coarse, raw (sized for the fine pass, reused), weights (fine size, reused), z_fine
  return align256(r * Sc * 4) + align256(r * Sf * 16) + align256(r * Sf * 4) + align256(r * Sf * 4);
}

int nfb_render_rays_fwd(const nfb_mlp_t* coarse, const nfb_mlp_t* fine, const float* rays, int R, int N_samples,
                        int N_importance, int lindisp, int white_bkgd, const float* t_rand, const float* u,
                        int perturb, uint64_t rng_seed, uint64_t rng_offset, float* rgb, float* disp, float* acc, float* rgb0, float* disp0, float* acc0, float* z_std,
                        float* pts_max, void* workspace, size_t workspace_bytes, void* stream) {
  NFB_REQUIRE(coarse && rays && rgb && disp && acc, "render_rays_fwd: null pointer");
  NFB_REQUIRE(N_importance == 0 || (rgb0 && disp0 && acc0), "render_rays_fwd: coarse outputs are required when N_importance > 0");
  NFB_REQUIRE(R >= 0 && N_samples >= 1 && N_importance >= 0, "render_rays_fwd: R=%d N_samples=%d N_importance=%d", R, N_samples, N_importance);
  if (R == 0) return NFB_OK;
  NFB_REQUIRE(workspace && workspace_bytes >= nfb_render_rays_workspace_bytes(R, N_samples, N_importance) &&
              (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "render_rays_fwd: workspace too small or not 256-byte aligned");
  const size_t Sc = (size_t)N_samples, Sf = Sc + (size_t)N_importance, r = (size_t)R;
  char* w = static_cast<char*>(workspace);
  float* z_c = reinterpret_cast<float*>(w);  w += align256(r * Sc * 4);
  float* raw = reinterpret_cast<float*>(w);  w += align256(r * Sf * 16);
  float* wts = reinterpret_cast<float*>(w);  w += align256(r * Sf * 4);
  float* z_f = reinterpret_cast<float*>(w);
  // stratified draws: the caller's numbers when given, else (perturb != 0) Philox in the kernels, else deterministic
  const bool rng_t = perturb && !t_rand, rng_u = perturb && !u;
  int rc = rng_t ? nfb_coarse_z_rng(rays, R, N_samples, lindisp, rng_seed, rng_offset, z_c, stream)
                 : nfb_coarse_z(rays, R, N_samples, lindisp, t_rand, z_c, stream);                     // run_nerf.py:357-379
  if (rc == NFB_OK) rc = nfb_mlp_fwd(coarse, 1, nullptr, nullptr, rays, z_c, R, N_samples, raw, stream);   // :381-385
  const bool two = N_importance > 0;
  if (rc == NFB_OK)                                                                                     // :386
    rc = nfb_composite_fwd(raw, z_c, rays + 3, 11, nullptr, R, N_samples, white_bkgd, two ? rgb0 : rgb, two ? disp0 : disp,
                           two ? acc0 : acc, wts, nullptr, two ? nullptr : pts_max, stream);
  if (rc != NFB_OK || !two) return rc;
  rc = rng_u ? nfb_hierarchical_rng(z_c, wts, rng_seed, rng_offset, R, N_samples, N_importance, z_f, nullptr, z_std, stream)
             : nfb_hierarchical(z_c, wts, u, R, N_samples, N_importance, z_f, nullptr, z_std, stream);  // :392-396, :412
  if (rc == NFB_OK) rc = nfb_mlp_fwd(fine ? fine : coarse, 1, nullptr, nullptr, rays, z_f, R, (int)Sf, raw, stream);   // :397-401
  if (rc == NFB_OK)                                                                                     // :403, nerf_to_coord.py:418-421
    rc = nfb_composite_fwd(raw, z_f, rays + 3, 11, nullptr, R, (int)Sf, white_bkgd, rgb, disp, acc, wts, nullptr, pts_max, stream);
  return rc;
}

// ---- one ray batch, coarse + fine, UNDER AUTOGRAD: forward that keeps what the backward needs, and the backward ----
// Workspace of the pair (caller-owned, handed from the forward to the backward untouched):
//   saved by the forward : z_c [R,Sc], z_f [R,Sf], raw_c [R,Sc,4], raw_f [R,Sf,4], act / mask tile images of both networks
//   scratch              : weights [R,Sf] (forward), g_raw [R,Sf,4] + dY tile image sized for the fine pass (backward)
namespace {
struct TrainWs {
  size_t z_c, z_f, raw_c, raw_f, wts, act_c, mask_c, act_f, mask_f, g_raw, dy, total;
};
TrainWs train_ws(int R, int N_samples, int N_importance) {
  const size_t Sc = (size_t)N_samples, Sf = Sc + (size_t)N_importance, r = (size_t)R;
  const size_t tc = (size_t)nfb_mlp_train_tiles((int64_t)(r * Sc)), tf = (size_t)nfb_mlp_train_tiles((int64_t)(r * Sf));
  TrainWs w{};
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += align256(bytes); return at; };
  w.z_c = take(r * Sc * 4);    w.z_f = take(r * Sf * 4);
  w.raw_c = take(r * Sc * 16); w.raw_f = take(r * Sf * 16);
  w.wts = take(r * Sf * 4);
  w.act_c = take(tc * 40 * 16384); w.mask_c = take(tc * 9 * 8 * 128 * 4);
  w.act_f = take(tf * 40 * 16384); w.mask_f = take(tf * 9 * 8 * 128 * 4);
  w.g_raw = take(r * Sf * 16);
  w.dy = take(tf * 39 * 16384);
  w.total = o;
  return w;
}
}  // namespace

size_t nfb_render_rays_train_workspace_bytes(int R, int N_samples, int N_importance) {
  if (R <= 0 || N_samples <= 0 || N_importance <= 0) return 0;
  return train_ws(R, N_samples, N_importance).total;
}

size_t nfb_render_rays_train_raw_offset(int R, int N_samples, int N_importance) {
  if (R <= 0 || N_samples <= 0 || N_importance <= 0) return 0;
  return train_ws(R, N_samples, N_importance).raw_f;
}

int nfb_render_rays_train_fwd(const nfb_mlp_t* coarse, const nfb_mlp_t* fine, const float* rays, int R, int N_samples,
                              int N_importance, int lindisp, int white_bkgd, const float* t_rand, const float* u,
                              float* rgb, float* disp, float* acc, float* rgb0, float* disp0, float* acc0, float* z_std,
                              void* workspace, size_t workspace_bytes, void* stream) {
  NFB_REQUIRE(coarse && fine && rays && rgb && disp && acc && rgb0 && disp0 && acc0, "render_rays_train_fwd: null pointer");
  NFB_REQUIRE(R >= 0 && N_samples >= 1 && N_importance >= 1, "render_rays_train_fwd: R=%d N_samples=%d N_importance=%d (both passes required)", R, N_samples, N_importance);
  if (R == 0) return NFB_OK;
  const TrainWs w = train_ws(R, N_samples, N_importance);
  NFB_REQUIRE(workspace && workspace_bytes >= w.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "render_rays_train_fwd: workspace too small or not 256-byte aligned");
  char* b = static_cast<char*>(workspace);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(b + off); };
  const int Sf = N_samples + N_importance;
  int rc = nfb_coarse_z(rays, R, N_samples, lindisp, t_rand, F(w.z_c), stream);                                   // run_nerf.py:357-379
  if (rc == NFB_OK) rc = nfb_mlp_fwd_train(coarse, rays, F(w.z_c), R, N_samples, F(w.raw_c), b + w.act_c,
                                           reinterpret_cast<uint32_t*>(b + w.mask_c), stream);                     // :381-385
  if (rc == NFB_OK) rc = nfb_composite_fwd(F(w.raw_c), F(w.z_c), rays + 3, 11, nullptr, R, N_samples, white_bkgd,
                                           rgb0, disp0, acc0, F(w.wts), nullptr, nullptr, stream);                 // :386
  if (rc == NFB_OK) rc = nfb_hierarchical(F(w.z_c), F(w.wts), u, R, N_samples, N_importance, F(w.z_f), nullptr, z_std, stream);   // :392-396 (detached)
  if (rc == NFB_OK) rc = nfb_mlp_fwd_train(fine, rays, F(w.z_f), R, Sf, F(w.raw_f), b + w.act_f,
                                           reinterpret_cast<uint32_t*>(b + w.mask_f), stream);                     // :397-401
  if (rc == NFB_OK) rc = nfb_composite_fwd(F(w.raw_f), F(w.z_f), rays + 3, 11, nullptr, R, Sf, white_bkgd,
                                           rgb, disp, acc, F(w.wts), nullptr, nullptr, stream);                    // :403
  return rc;
}

int nfb_render_rays_bwd(const nfb_mlp_t* coarse, const nfb_mlp_t* fine, const float* rays, int R, int N_samples,
                        int N_importance, int white_bkgd, const float* g_rgb, const float* g_disp, const float* g_acc,
                        const float* g_rgb0, const float* g_disp0, const float* g_acc0, float* grad_coarse, float* grad_fine,
                        void* workspace, size_t workspace_bytes, void* stream) {
  NFB_REQUIRE(coarse && fine && rays && grad_coarse && grad_fine, "render_rays_bwd: null pointer");
  NFB_REQUIRE(R >= 0 && N_samples >= 1 && N_importance >= 1, "render_rays_bwd: R=%d N_samples=%d N_importance=%d", R, N_samples, N_importance);
  if (R == 0) return NFB_OK;
  const TrainWs w = train_ws(R, N_samples, N_importance);
  NFB_REQUIRE(workspace && workspace_bytes >= w.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "render_rays_bwd: workspace too small or not 256-byte aligned");
  char* b = static_cast<char*>(workspace);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(b + off); };
  const int Sf = N_samples + N_importance;
  int rc = NFB_OK;
  if (g_rgb || g_disp || g_acc) {                    // fine pass: compositing -> data gradient chain -> all weight gradients
    const int64_t M = (int64_t)R * Sf;
    rc = nfb_composite_bwd(F(w.raw_f), F(w.z_f), rays + 3, 11, nullptr, R, Sf, white_bkgd, g_rgb, g_disp, g_acc, nullptr, nullptr,
                           F(w.g_raw), stream);
    if (rc == NFB_OK) rc = nfb_mlp_bwd_data(fine, F(w.g_raw), M, reinterpret_cast<const uint32_t*>(b + w.mask_f), b + w.dy, stream);
    if (rc == NFB_OK) rc = nfb_mlp_bwd_weights(fine, b + w.act_f, b + w.dy, F(w.g_raw), M, grad_fine, stream);
  }
  if (rc == NFB_OK && (g_rgb0 || g_disp0 || g_acc0)) {   // coarse pass (the resampled depths are detached, run_nerf.py:394)
    const int64_t M = (int64_t)R * N_samples;
    rc = nfb_composite_bwd(F(w.raw_c), F(w.z_c), rays + 3, 11, nullptr, R, N_samples, white_bkgd, g_rgb0, g_disp0, g_acc0, nullptr,
                           nullptr, F(w.g_raw), stream);
    if (rc == NFB_OK) rc = nfb_mlp_bwd_data(coarse, F(w.g_raw), M, reinterpret_cast<const uint32_t*>(b + w.mask_c), b + w.dy, stream);
    if (rc == NFB_OK) rc = nfb_mlp_bwd_weights(coarse, b + w.act_c, b + w.dy, F(w.g_raw), M, grad_coarse, stream);
  }
  return rc;
}

}  // extern "C"
