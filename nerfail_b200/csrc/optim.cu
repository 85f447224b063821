// Fused multi-tensor Adam step.
//
// Reference: torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999)) created at
// Create_spatial_point_set/nerf_pytorch/run_nerf.py:213 and stepped at :792, followed by the exponential learning-rate
// decay of :796-800 (the new rate is passed in by the host for the NEXT step, as the reference sets it after step()).
// Arithmetic follows torch/optim/adam.py::_single_tensor_adam (no amsgrad, no weight decay, maximize = False):
//     m   = m + (g - m) * (1 - beta1)                 (lerp)
//     v   = v * beta2 + (1 - beta2) * g * g
//     p  -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// One launch updates every parameter tensor of both networks (48 tensors, 1.19 M floats): HBM-bound, 28 B per element
// (read p, g, m, v; write p, m, v).  grad_scale folds the 1/world_size of the data-parallel mean into the same pass.
#include "common.cuh"
#include <math.h>

namespace nfb {

constexpr int ADAM_MAX_TENSORS = 64;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = 4096;          // elements per CTA work item

struct AdamTable {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  int chunk0[ADAM_MAX_TENSORS + 1];       // first work item of tensor t (prefix sum of ceil(n / ADAM_CHUNK))
  int64_t n[ADAM_MAX_TENSORS];
  int count;
};

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const __grid_constant__ AdamTable tb, float step_size, float one_minus_beta1, float beta2, float one_minus_beta2,
            float inv_sqrt_bc2, float eps, float grad_scale, const float* __restrict__ dyn) {
  // dyn (device, 2 floats) overrides the two step-dependent constants: a launch captured in a CUDA graph is replayed
  // with the step size and bias correction of the CURRENT step, written by the host before each replay
  if (dyn) { step_size = __ldg(dyn); inv_sqrt_bc2 = __ldg(dyn + 1); }
  // locate the tensor of this work item (binary search over <= 64 prefix sums held in the parameter bank)
  int lo = 0, hi = tb.count;
  const int item = blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tb.chunk0[mid] <= item) lo = mid; else hi = mid;
  }
  const int t = lo;
  const int64_t begin = (int64_t)(item - tb.chunk0[t]) * ADAM_CHUNK;
  const int64_t end = min(begin + ADAM_CHUNK, tb.n[t]);
  float* __restrict__ p = tb.p[t];
  const float* __restrict__ g = tb.g[t];
  float* __restrict__ m = tb.m[t];
  float* __restrict__ v = tb.v[t];
  for (int64_t i = begin + threadIdx.x; i < end; i += ADAM_THREADS) {
    const float gi = g[i] * grad_scale;
    const float mi = fmaf(gi - m[i], one_minus_beta1, m[i]);
    const float vi = fmaf(v[i], beta2, one_minus_beta2 * gi * gi);
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

}  // namespace nfb

extern "C" {

// the two step-dependent constants of the update, derived in double like torch's Python scalars and rounded once
int nfb_adam_step_scalars(int64_t step, double lr, double beta1, double beta2, float* out2) {
  NFB_REQUIRE(out2 && step >= 1, "adam_step_scalars: step=%lld", (long long)step);
  NFB_REQUIRE(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1., "adam_step_scalars: bad betas");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  out2[0] = (float)(lr / bc1);
  out2[1] = (float)(1.0 / sqrt(bc2));
  return NFB_OK;
}

static int adam_launch(const nfb_adam_tensor* tensors, int count, float step_size, float inv_sqrt_bc2, const float* dyn,
                       double beta1, double beta2, double eps, double grad_scale, void* stream) {
  for (int t0 = 0; t0 < count; t0 += nfb::ADAM_MAX_TENSORS) {
    nfb::AdamTable tb{};
    const int n = count - t0 < nfb::ADAM_MAX_TENSORS ? count - t0 : nfb::ADAM_MAX_TENSORS;
    int items = 0;
    for (int i = 0; i < n; ++i) {
      const nfb_adam_tensor& s = tensors[t0 + i];
      NFB_REQUIRE(s.numel >= 0 && (s.numel == 0 || (s.param && s.grad && s.exp_avg && s.exp_avg_sq)), "adam_step: tensor %d: null pointer", t0 + i);
      tb.p[i] = s.param; tb.g[i] = s.grad; tb.m[i] = s.exp_avg; tb.v[i] = s.exp_avg_sq; tb.n[i] = s.numel;
      tb.chunk0[i] = items;
      items += (int)((s.numel + nfb::ADAM_CHUNK - 1) / nfb::ADAM_CHUNK);
    }
    tb.chunk0[n] = items;
    tb.count = n;
    if (items == 0) continue;
    nfb::adam_kernel<<<items, nfb::ADAM_THREADS, 0, (cudaStream_t)stream>>>(tb, step_size, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                                                  inv_sqrt_bc2, (float)eps, (float)grad_scale, dyn);
    int rc = nfb::check_launch("adam_step");
    if (rc) return rc;
  }
  return NFB_OK;
}

int nfb_adam_step(const nfb_adam_tensor* tensors, int count, int64_t step, double lr, double beta1, double beta2, double eps,
                  double grad_scale, void* stream) {
  NFB_REQUIRE(tensors || count == 0, "adam_step: null tensor table");
  NFB_REQUIRE(count >= 0 && step >= 1, "adam_step: count=%d step=%lld", count, (long long)step);
  NFB_REQUIRE(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., "adam_step: bad hyper-parameters");
  float sc[2];
  int rc = nfb_adam_step_scalars(step, lr, beta1, beta2, sc);
  if (rc) return rc;
  return adam_launch(tensors, count, sc[0], sc[1], nullptr, beta1, beta2, eps, grad_scale, stream);
}

// Same update with the step-dependent constants read from device memory (step_scalars = what nfb_adam_step_scalars
// returns, copied to the device by the caller before the launch runs): the form that can be captured in a CUDA graph.
int nfb_adam_step_dev(const nfb_adam_tensor* tensors, int count, const float* step_scalars, double beta1, double beta2,
                      double eps, double grad_scale, void* stream) {
  NFB_REQUIRE(tensors || count == 0, "adam_step_dev: null tensor table");
  NFB_REQUIRE(count >= 0 && step_scalars, "adam_step_dev: count=%d", count);
  NFB_REQUIRE(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., "adam_step_dev: bad hyper-parameters");
  return adam_launch(tensors, count, 0.f, 0.f, step_scalars, beta1, beta2, eps, grad_scale, stream);
}

}  // extern "C"
