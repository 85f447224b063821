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

}  // extern "C"
