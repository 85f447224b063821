// Fused positional encoding + 8x256 skip MLP (+ alpha / feature / view / rgb heads) on tcgen05 tensor cores.
//
// Reference arithmetic: Create_spatial_point_set/nerf_pytorch/run_nerf.py:37-51 (run_network),
// run_nerf_helpers.py:36-50 (Embedder), :100-123 (NeRF.forward), run_nerf.py:381 (pts = o + d*z).
//
// Data flow for one 128-sample tile (one CTA works on two tiles, "slots", in ping-pong):
//   input stage : each epilogue thread owns one sample (= one TMEM lane): builds the 63 (+1 pad) positional
//                 features in registers and writes them as bf16 into the slot's PE chunk (128x64, K-major,
//                 128-byte swizzle) -> A operand of layer 0 and, again, of the skip layer 5.
//   layer l     : the MMA thread issues tcgen05.mma (M=128, N=128 per weight half, K=16) with
//                 A = the slot's activation chunks in shared memory, B = a 16 KB weight block streamed by
//                 the bulk-copy engine (cp.async.bulk + mbarrier tx-count) from a pre-swizzled bf16 image,
//                 D = 128x256 fp32 accumulator in TMEM (columns slot*256 ..).
//   epilogue l  : 4 warps read the accumulator with tcgen05.ld (32x32b.x32), add the fp32 bias, relu,
//                 round to bf16 and overwrite the slot's activation chunks in place (the MMAs that read
//                 them have completed).  Layer 7's epilogue also forms sigma = w_alpha . h + b in fp32;
//                 the view layer's epilogue forms the three rgb logits in fp32 and stores the float4 raw.
//   While slot 0 is in its epilogue the tensor core runs slot 1's layer, and vice versa.
//
// Activations never leave the SM; HBM traffic is 16 B read + 16 B written per sample.
#include "common.cuh"
#include <stdlib.h>
#include <mutex>

namespace nfb {

// ---------------------------------------------------------------------------------------------------
// Architecture constants (the reference's one shipped NeRF: configs/lego.txt through run_nerf.py:181-198)
// ---------------------------------------------------------------------------------------------------
constexpr int W_ = 256;            // netwidth
constexpr int L_PTS = 10;          // multires      -> 63 features
constexpr int L_DIR = 4;           // multires_views-> 27 features
constexpr int CH_PTS = 63, CH_DIR = 27;
constexpr int TILE_M = 128;
constexpr int KCH = 64;            // K elements per smem chunk (= one 128-byte swizzle row of bf16)
constexpr int CHUNK_BYTES = TILE_M * KCH * 2;     // 16 KB: activation chunk and weight block alike
constexpr int NSTEP = 10;          // MMA steps per tile: L0..L7, feature, views
constexpr int NSTAGE = 4;          // weight ring depth
constexpr int NUM_THREADS = 576;   // warp 0 producer, warp 1 MMA + TMEM owner, warps 2-9 slot 0, warps 10-17 slot 1

// weight blocks (16 KB each: 128 output rows x 64 K) per step, in consumption order (chunk-major, half-minor)
__host__ __device__ constexpr int step_kchunks(int s) { return s == 0 ? 1 : (s == 5 || s == 9) ? 5 : 4; }
__host__ __device__ constexpr int step_halves(int s) { return s == 9 ? 1 : 2; }
__host__ __device__ constexpr int step_block0(int s) {
  int b = 0;
  for (int i = 0; i < s; ++i) b += step_kchunks(i) * step_halves(i);
  return b;
}
constexpr int TOTAL_BLOCKS = step_block0(NSTEP);   // 2 + 32 + 10 + 16 + 8 + 5 = 73
static_assert(TOTAL_BLOCKS == 73, "weight block count");

// shared memory map (offsets from a 1024-byte aligned base)
constexpr int SM_ACT = 0;                                   // [2 slots][4 chunks][16 KB]
constexpr int SM_PE = SM_ACT + 2 * 4 * CHUNK_BYTES;         // [2 slots][16 KB]
constexpr int SM_W = SM_PE + 2 * CHUNK_BYTES;               // [NSTAGE][16 KB]
constexpr int SM_BAR = SM_W + NSTAGE * CHUNK_BYTES;         // barriers + tmem pointer
constexpr int SM_BIAS = SM_BAR + 256;                        // [2 slots][256 floats]: bias row of the step being drained
constexpr int SM_TOTAL = SM_BIAS + 2 * 256 * 4;
constexpr int SMEM_BYTES = SM_TOTAL;                        // the dynamic window starts 1024-byte aligned (checked in-kernel)
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory per CTA");

// fp32 side parameters (biases and the two small heads evaluated on CUDA cores), device resident
struct MlpSide {
  float bias[8][W_];       // pts_linears.0..7
  float bias_feat[W_];
  float bias_views[128];
  float w_alpha[W_];
  float w_rgb[3][128];
  float b_alpha;
  float b_rgb[3];
};

// The epilogues read biases / head weights warp-uniformly.  With ~225 KB of shared memory per CTA the L1 carve-out is
// a few KB, so global loads of these 12 KB per network miss L1 and cost an L2 round trip per 32 columns (measured: 900
// cycles per 32-column epilogue iteration).  They live in constant memory instead: MAX_NETS entries per device, shared
// LRU among however many networks are alive (ensure_cslot below).
constexpr int MAX_NETS = 4;
__constant__ MlpSide c_side[MAX_NETS];

}  // namespace nfb

struct nfb_mlp {
  __nv_bfloat16* image;      // TOTAL_BLOCKS x 16 KB pre-swizzled weight blocks
  __nv_bfloat16* image_t;    // transposed blocks for the data-gradient chain (68 x 16 KB)
  nfb::MlpSide* side;        // staging copy in global memory (written by the pack kernel)
  mutable int cslot;         // index into the __constant__ c_side table of this device, or -1 while evicted (see ensure_cslot)
  int* abort_flag;           // set by the kernel if a barrier wait timed out: DEVICE alias of abort_host (mapped pinned memory)
  volatile int* abort_host;  // the same word seen from the host: polled without synchronising (nfb_mlp_poll, every launch)
  void* zero16k;             // 16 KB of zeros: the padding dY chunk of the head weight-gradient products
  cudaStream_t side_stream;  // second stream + fork/join events: the weight-gradient kernel of nfb_mlp_bwd runs next to
  cudaEvent_t ev_fork, ev_join;   // the data-gradient kernel on its own SMs
  int device;
  int64_t n_params;
};

namespace nfb {

// state_dict offsets in the flat fp32 parameter buffer (nn.Module registration order, run_nerf_helpers.py:82-96)
struct ParamLayout {
  int64_t w_pts[8], b_pts[8], w_views, b_views, w_feat, b_feat, w_alpha, b_alpha, w_rgb, b_rgb, total;
};
__host__ __device__ inline ParamLayout param_layout() {
  ParamLayout p{};
  int64_t o = 0;
  for (int l = 0; l < 8; ++l) {
    const int in = (l == 0) ? CH_PTS : (l == 5 ? W_ + CH_PTS : W_);
    p.w_pts[l] = o; o += (int64_t)W_ * in;
    p.b_pts[l] = o; o += W_;
  }
  p.w_views = o; o += 128 * (W_ + CH_DIR);
  p.b_views = o; o += 128;
  p.w_feat = o; o += W_ * W_;
  p.b_feat = o; o += W_;
  p.w_alpha = o; o += W_;
  p.b_alpha = o; o += 1;
  p.w_rgb = o; o += 3 * 128;
  p.b_rgb = o; o += 3;
  p.total = o;
  return p;
}

// byte offset of element (row, k) inside a 128x64 bf16 K-major chunk with the 128-byte swizzle
// (16-byte unit index XOR row%8; chunk base must be 1024-byte aligned) — the layout both the UMMA smem
// descriptors below and the epilogue's manual stores use.
__host__ __device__ __forceinline__ int swz_off(int row, int k) {
  return row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + ((k & 7) << 1);
}

// ---------------------------------------------------------------------------------------------------
// weight packing: fp32 state_dict -> bf16 swizzled blocks + fp32 side parameters
// ---------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ image,
                                    MlpSide* __restrict__ side) {
  const ParamLayout pl = param_layout();
  // one thread per bf16 element of the image
  const int64_t total = (int64_t)TOTAL_BLOCKS * TILE_M * KCH;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int blk = (int)(e / (TILE_M * KCH));
    const int within = (int)(e % (TILE_M * KCH));
    const int nrow = within / KCH, kk = within % KCH;
    int s = 0;
    while (s + 1 < NSTEP && step_block0(s + 1) <= blk) ++s;
    const int local = blk - step_block0(s);
    const int halves = step_halves(s);
    const int chunk = local / halves, half = local % halves;
    const int n = half * 128 + nrow;           // output feature
    float v = 0.f;
    if (s <= 7) {
      const int in = (s == 0) ? CH_PTS : (s == 5 ? W_ + CH_PTS : W_);
      int col = -1;                            // column of the reference weight [256, in]
      if (s == 0) col = (kk < CH_PTS) ? kk : -1;
      else if (s == 5) {                       // cat([input_pts, h]) : chunks 0..3 = h part, chunk 4 = PE part
        if (chunk == 4) col = (kk < CH_PTS) ? kk : -1;
        else col = CH_PTS + chunk * KCH + kk;
      } else col = chunk * KCH + kk;
      if (col >= 0) v = params[pl.w_pts[s] + (int64_t)n * in + col];
    } else if (s == 8) {
      v = params[pl.w_feat + (int64_t)n * W_ + chunk * KCH + kk];
    } else {                                   // views: cat([feature, input_views]) : chunks 0..3 feature, chunk 4 dirs
      const int in = W_ + CH_DIR;
      int col = -1;
      if (chunk < 4) col = chunk * KCH + kk;
      else col = (kk < CH_DIR) ? W_ + kk : -1;
      if (col >= 0) v = params[pl.w_views + (int64_t)n * in + col];
    }
    char* dst = reinterpret_cast<char*>(image) + (int64_t)blk * CHUNK_BYTES + swz_off(nrow, kk);
    *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(v);
  }
  // side parameters
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nt = gridDim.x * blockDim.x;
  for (int i = t; i < 8 * W_; i += nt) side->bias[i / W_][i % W_] = params[pl.b_pts[i / W_] + i % W_];
  for (int i = t; i < W_; i += nt) { side->bias_feat[i] = params[pl.b_feat + i]; side->w_alpha[i] = params[pl.w_alpha + i]; }
  for (int i = t; i < 128; i += nt) side->bias_views[i] = params[pl.b_views + i];
  for (int i = t; i < 3 * 128; i += nt) side->w_rgb[i / 128][i % 128] = params[pl.w_rgb + i];
  if (t == 0) {
    side->b_alpha = params[pl.b_alpha];
    side->b_rgb[0] = params[pl.b_rgb]; side->b_rgb[1] = params[pl.b_rgb + 1]; side->b_rgb[2] = params[pl.b_rgb + 2];
  }
}

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  On timeout the abort flag is raised, every later wait
// in the grid falls through, the kernel finishes with garbage and the host reports NFB_E_CUDA.
// The spin loop lives inside ONE asm statement so that the compiler sees straight-line code: control flow around
// the waits stays warp-uniform and the MMA warp can keep descriptors in uniform registers (no per-instruction
// ELECT + R2UR when issuing tcgen05.mma).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .u32 n, f;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra NFB_DONE;\n\t"
      "mov.u32 n, 0;\n\t"
      "NFB_SPIN:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra NFB_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "and.b32 f, n, 1023;\n\t"
      "setp.ne.u32 q, f, 0;\n\t"
      "@q bra NFB_SPIN;\n\t"
      "ld.volatile.global.u32 f, [%2];\n\t"
      "setp.ne.u32 q, f, 0;\n\t"
      "@q bra NFB_DONE;\n\t"
      "setp.lt.u32 q, n, 4194304;\n\t"
      "@q bra NFB_SPIN;\n\t"
      "mov.u32 f, 1;\n\t"
      "st.volatile.global.u32 [%2], f;\n\t"
      "NFB_DONE:\n\t"
      "}"
      :: "r"(bar), "r"(parity), "l"(abort_flag) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset  bits [16,30) (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                            // version = 1 (sm_100) bits [46,48)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B         bits [61,64)
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// two fp32 additions in one instruction (add.rn.f32x2 -> FADD2 on sm_100): the bias add of the accumulator drain
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  unsigned long long a, b, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// positional encoding of one 3-vector into `feat` (bf16 pairs), reference order [x, sin(2^0 x), cos(2^0 x), ...]
// sin/cos of 2^l x by exact angle doubling from an accurate sincosf(x): the error doubles per octave and stays
// below 3e-5 at l = 9, two orders under the bf16 rounding that follows.
// ---------------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ void encode3(float x, float y, float z, float (&f)[64]) {
  f[0] = x; f[1] = y; f[2] = z;
  float s[3], c[3];
  sincosf(x, &s[0], &c[0]); sincosf(y, &s[1], &c[1]); sincosf(z, &s[2], &c[2]);
#pragma unroll
  for (int l = 0; l < L; ++l) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      f[3 + 6 * l + a] = s[a];
      f[3 + 6 * l + 3 + a] = c[a];
      const float s2 = 2.f * s[a] * c[a];
      const float c2 = 1.f - 2.f * s[a] * s[a];
      s[a] = s2; c[a] = c2;
    }
  }
#pragma unroll
  for (int i = 3 + 6 * L; i < 64; ++i) f[i] = 0.f;
}

__device__ __forceinline__ void store_row_chunk(uint32_t chunk_base, int row, const float (&f)[64]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint32_t addr = chunk_base + row * 128 + (((u ^ (row & 7)) & 7) << 4);
    st_shared_v4(addr, pack_bf16(f[8 * u], f[8 * u + 1]), pack_bf16(f[8 * u + 2], f[8 * u + 3]),
                 pack_bf16(f[8 * u + 4], f[8 * u + 5]), pack_bf16(f[8 * u + 6], f[8 * u + 7]));
  }
}

struct FwdArgs {
  const __nv_bfloat16* image;
  const MlpSide* side;      // global copy (bias rows are staged from here into shared memory)
  int cslot;                // index into c_side (constant memory): head weights read through the uniform datapath
  int* abort_flag;
  int mode;                 // 0: explicit pts + dirs, 1: rays + z_vals
  const float* pts;         // [M,3]
  const float* dirs;        // [R,3]
  const float* rays;        // [R,11]
  const float* z_vals;      // [M]
  int64_t M;                // R*S samples
  int S;
  float* raw;               // [M,4]
  int nsteps;               // 10 normally; < 10 = debug: stop after that many MMA steps and dump activations
  float* dbg;               // [M,256] fp32 post-activation values of the last executed step (debug only)
  unsigned long long* trace;   // optional timeline of CTA 0: [3 roles][TRACE_CAP][4] = tag, t_begin, t_end, aux (profiling aid)
};
constexpr int TRACE_CAP = 2048;

// ---- cluster / cta_group::2 helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      :: "r"(bar), "r"(cta) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_issue(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// tcgen05.commit: the mbarrier (same offset in every CTA of the group) is arrived on when all MMAs issued so far retire
template <int CG>
__device__ __forceinline__ void umma_commit_group(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
  }
}
__device__ __forceinline__ uint32_t umma_idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// A-operand chunk of MMA step s, K-chunk c (consumption order: the 256-wide activation first, then the encoding chunk)
__device__ __forceinline__ uint32_t a_chunk_addr(int s, int c, uint32_t act, uint32_t pe) {
  if (s == 0) return pe;
  if ((s == 5 || s == 9) && c == 4) return pe;
  return act + c * CHUNK_BYTES;
}

// ---------------------------------------------------------------------------------------------------
// the kernel.  CG = 1: one CTA per SM, 128-row tiles, every CTA streams whole weight blocks per tile slot.
//              CG = 2: CTA pairs (cluster of 2) run tcgen05.mma.cta_group::2 on 256-row tiles: each CTA holds its 128
//                      rows of A and HALF of each weight block (N split), so L2->SMEM weight traffic and SMEM operand
//                      reads per CTA are halved, one MMA covers N = 256, and both tile slots reuse the resident weights.
// ---------------------------------------------------------------------------------------------------
// AUX = true: the instantiation behind nfb_mlp_fwd_debug / nfb_mlp_fwd_trace (partial runs, activation dumps, timeline);
// the production instantiation has those branches compiled out of the issue loop and the epilogue.
template <int CG, bool AUX = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_fused_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);          // swizzled operands need a 1024-byte aligned window
  const uint32_t bar0 = base + SM_BAR;
  auto W_FULL = [&](int s) { return bar0 + 8 * s; };
  auto W_EMPTY = [&](int s) { return bar0 + 8 * (NSTAGE + s); };
  auto A_READY = [&](int g) { return bar0 + 8 * (2 * NSTAGE + g); };
  auto ACC_FULL = [&](int g) { return bar0 + 8 * (2 * NSTAGE + 2 + g); };
  const uint32_t tmem_slot = bar0 + 8 * (2 * NSTAGE + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + SM_BAR + 8 * (2 * NSTAGE + 4));

  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);   // broadcast => warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  volatile int* abort_flag = a.abort_flag;
  const int nsteps = AUX ? a.nsteps : NSTEP;
  float* const dbg_out = AUX ? a.dbg : nullptr;
  if ((base & 1023u) != 0u && threadIdx.x == 0) *abort_flag = 1;   // misaligned window: results would be garbage
  const bool tracing = AUX && a.trace != nullptr && blockIdx.x == 0;
  int trace_n = 0;
  auto trace_evt = [&](int role, unsigned long long tag, unsigned long long t0, unsigned long long t1, unsigned long long aux) {
    if (tracing && lane == 0 && trace_n < TRACE_CAP) {
      unsigned long long* e = a.trace + ((size_t)role * TRACE_CAP + trace_n) * 4;
      e[0] = tag; e[1] = t0; e[2] = t1; e[3] = aux;
      ++trace_n;
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(W_FULL(s), (CG == 2 && leader) ? 2 : 1);      // CG=2 leader: own producer + the peer's relay
      mbar_init(W_EMPTY(s), 1);
    }
    for (int g = 0; g < 2; ++g) { mbar_init(A_READY(g), 8 * CG); mbar_init(ACC_FULL(g), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {   // TMEM: all 512 columns (two 128x256 fp32 accumulators per CTA)
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();       // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  constexpr int ROWS_PER_UNIT = 2 * TILE_M * CG;    // two tile slots
  const int64_t nunits = (a.M + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT;
  const int64_t group = blockIdx.x / CG, ngroups = gridDim.x / CG;

  if (warp == 0) {
    // ================= weight producer (whole warp waits, one elected lane issues the bulk copies) =================
    uint32_t pos = 0;
    for (int64_t unit = group; unit < nunits; unit += ngroups) {
      for (int s = 0; s < nsteps; ++s) {
        const int kch = step_kchunks(s), halves = step_halves(s);
        const char* src = reinterpret_cast<const char*>(a.image) + (int64_t)step_block0(s) * CHUNK_BYTES;
        if constexpr (CG == 1) {
          for (int g = 0; g < 2; ++g)
            for (int b = 0; b < kch * halves; ++b, ++pos) {
              const int stage = pos & (NSTAGE - 1);
              mbar_wait(W_EMPTY(stage), ((pos >> 2) & 1) ^ 1, abort_flag);
              if (elect_one()) {
                mbar_arrive_expect_tx(W_FULL(stage), CHUNK_BYTES);
                bulk_g2s(base + SM_W + stage * CHUNK_BYTES, src + (int64_t)b * CHUNK_BYTES, CHUNK_BYTES, W_FULL(stage));
              }
            }
        } else {
          // this CTA's half of every block, once per step (both slots reuse it): N rows [rank*N/2, (rank+1)*N/2)
          const uint32_t bytes = (halves == 2) ? CHUNK_BYTES : CHUNK_BYTES / 2;
          for (int c = 0; c < kch; ++c, ++pos) {
            const int stage = pos & (NSTAGE - 1);
            const char* blk = (halves == 2) ? src + (int64_t)(c * 2 + rank) * CHUNK_BYTES
                                            : src + (int64_t)c * CHUNK_BYTES + rank * (CHUNK_BYTES / 2);
            const unsigned long long t0 = tracing ? clock64() : 0;
            mbar_wait(W_EMPTY(stage), ((pos >> 2) & 1) ^ 1, abort_flag);
            if (elect_one()) {
              mbar_arrive_expect_tx(W_FULL(stage), bytes);
              bulk_g2s(base + SM_W + stage * CHUNK_BYTES, blk, bytes, W_FULL(stage));
            }
            if (tracing) trace_evt(0, pos, t0, clock64(), (unsigned long long)s << 8 | c);
          }
        }
      }
    }
  } else if (warp == 1) {
    // Warp-uniform control flow: every lane walks the loops and waits on the barriers; one elected lane issues the
    // tcgen05.mma / tcgen05.commit instructions, whose descriptors then live in uniform registers.
    const uint32_t dlo_or = 1u << 16;
    const uint32_t dhi = (uint32_t)(umma_desc(0) >> 32);
    auto desc_of = [&](uint32_t saddr) { return ((uint64_t)dhi << 32) | (uint64_t)(((saddr & 0x3FFFF) >> 4) | dlo_or); };
    if (leader) {
      // ================= MMA issuer =================
      uint32_t pos = 0, ready_phase0 = 0, ready_phase1 = 0;
      for (int64_t unit = group; unit < nunits; unit += ngroups) {
        for (int s = 0; s < nsteps; ++s) {
          const int kch = step_kchunks(s), halves = step_halves(s);
          if constexpr (CG == 1) {
            const uint32_t idesc = umma_idesc_mn(128, 128);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              mbar_wait(A_READY(g), g == 0 ? ready_phase0 : ready_phase1, abort_flag);
              if (g == 0) ready_phase0 ^= 1; else ready_phase1 ^= 1;
              tc_fence_after();
              const uint32_t act = base + SM_ACT + g * 4 * CHUNK_BYTES, pe = base + SM_PE + g * CHUNK_BYTES;
              for (int c = 0; c < kch; ++c) {
                const uint64_t ad = desc_of(a_chunk_addr(s, c, act, pe));
                for (int h = 0; h < halves; ++h, ++pos) {
                  const int stage = pos & (NSTAGE - 1);
                  mbar_wait(W_FULL(stage), (pos >> 2) & 1, abort_flag);
                  tc_fence_after();
                  const uint64_t bd = desc_of(base + SM_W + stage * CHUNK_BYTES);
                  const uint32_t d = tmem_base + g * 256 + h * 128;
                  if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KCH / 16; ++k)
                      umma_issue<1>(d, ad + 2 * k, bd + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    umma_commit_group<1>(W_EMPTY(stage));
                    if (c == kch - 1 && h == halves - 1) umma_commit_group<1>(ACC_FULL(g));
                  }
                }
              }
            }
          } else {
            const uint32_t idesc = umma_idesc_mn(256, halves == 2 ? 256 : 128);
            const int main_ch = kch < NSTAGE ? kch : NSTAGE;     // chunks resident together; a 5th one goes last
            const uint32_t p0 = pos;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const unsigned long long ta = tracing ? clock64() : 0;
              mbar_wait(A_READY(g), g == 0 ? ready_phase0 : ready_phase1, abort_flag);
              if (tracing) trace_evt(1, 0x1000 | (s << 4) | g, ta, clock64(), 0);
              if (g == 0) ready_phase0 ^= 1; else ready_phase1 ^= 1;
              tc_fence_after();
              const uint32_t act = base + SM_ACT + g * 4 * CHUNK_BYTES, pe = base + SM_PE + g * CHUNK_BYTES;
              const uint32_t d = tmem_base + g * 256;
              for (int c = 0; c < main_ch; ++c) {
                const uint32_t p = p0 + c;
                const int stage = p & (NSTAGE - 1);
                if (g == 0) {
                  const unsigned long long tw = tracing ? clock64() : 0;
                  mbar_wait(W_FULL(stage), (p >> 2) & 1, abort_flag);
                  tc_fence_after();
                  if (tracing) trace_evt(1, 0x2000 | (s << 4) | c, tw, clock64(), p);
                }
                const uint64_t ad = desc_of(a_chunk_addr(s, c, act, pe));
                const uint64_t bd = desc_of(base + SM_W + stage * CHUNK_BYTES);
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < KCH / 16; ++k)
                    umma_issue<2>(d, ad + 2 * k, bd + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
                  if (g == 1) umma_commit_group<2>(W_EMPTY(stage));   // both slots have consumed the block
                  if (kch == main_ch && c == main_ch - 1) umma_commit_group<2>(ACC_FULL(g));
                }
              }
            }
            if (kch > main_ch) {                                  // the encoding chunk of steps 5 and 9
              const uint32_t p = p0 + main_ch;
              const int stage = p & (NSTAGE - 1);
              mbar_wait(W_FULL(stage), (p >> 2) & 1, abort_flag);
              tc_fence_after();
              const uint64_t bd = desc_of(base + SM_W + stage * CHUNK_BYTES);
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const uint32_t act = base + SM_ACT + g * 4 * CHUNK_BYTES, pe = base + SM_PE + g * CHUNK_BYTES;
                const uint64_t ad = desc_of(a_chunk_addr(s, main_ch, act, pe));
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < KCH / 16; ++k)
                    umma_issue<2>(tmem_base + g * 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
                  if (g == 1) umma_commit_group<2>(W_EMPTY(stage));
                  umma_commit_group<2>(ACC_FULL(g));
                }
              }
            }
            pos += kch;
          }
        }
      }
    } else if (CG == 2) {
      // ================= peer relay: tells the leader when this CTA's half of a block has landed =================
      uint32_t pos = 0;
      for (int64_t unit = group; unit < nunits; unit += ngroups)
        for (int s = 0; s < nsteps; ++s)
          for (int c = 0; c < step_kchunks(s); ++c, ++pos) {
            const int stage = pos & (NSTAGE - 1);
            mbar_wait(W_FULL(stage), (pos >> 2) & 1, abort_flag);
            if (elect_one()) mbar_arrive_remote(W_FULL(stage), 0);
          }
    }
  } else {
    // ================= input stage + epilogues =================
    // 8 warps per tile slot: warp id % 4 fixes the TMEM lane quarter (32 samples), (warp - 2) / 4 % 2 picks the column
    // half, so one sample's accumulator row is drained by two threads (128 columns each).  tcgen05.ld retires one
    // 32-column load per ~94 cycles per warp (measured, scripts/ubench/tmem_ld.cu); two warps per quarter halve the
    // drain latency that the other slot's MMAs have to cover.
    const int e = warp - 2;
    const int g = e >> 3;                             // slot
    const int hcol = (e >> 2) & 1;                    // column half of this thread
    const int q = warp & 3;                           // TMEM lane quarter
    const int row = (q << 5) + lane;
    const uint32_t act = base + SM_ACT + g * 4 * CHUNK_BYTES;
    const uint32_t pe = base + SM_PE + g * CHUNK_BYTES;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q << 5) << 16) + g * 256;
    const uint32_t pair_bar = 1 + g * 4 + q;          // named barrier shared by the two warps of a (slot, quarter)
    const MlpSide* __restrict__ sd = &c_side[a.cslot];
    uint32_t full_phase = 0;
    auto signal_a_ready = [&]() {                     // this warp's part of the A operand is in shared memory
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || leader) mbar_arrive(A_READY(g)); else mbar_arrive_remote(A_READY(g), 0);
      }
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" :: "r"(pair_bar) : "memory"); };
    // shared-memory scratch (generic proxy) used to combine the two column halves of a row
    float* scratch_pe = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + SM_PE + g * CHUNK_BYTES);
    float* scratch_act = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + SM_ACT + g * 4 * CHUNK_BYTES);
    // Bias staging: per-thread broadcast loads of the bias (global or constant memory) cost 2-4x the rest of a
    // 32-column epilogue iteration (scripts/ubench/epilogue.cu), so the 256 threads of a slot keep the bias row of
    // the step being drained in shared memory (read back as broadcast LDS.128) and refill it once per step with
    // the next step's row, prefetched into a register while the accumulator is drained.
    float* bias_s = reinterpret_cast<float*>(smem_raw + SM_BIAS) + g * 256;
    const int tslot = ((e & 7) << 5) + lane;          // 0..255 within the slot
    const uint32_t slot_bar = 9 + g;                  // named barrier of the slot's 8 epilogue warps
    auto slot_sync = [&]() { asm volatile("bar.sync %0, 256;" :: "r"(slot_bar) : "memory"); };
    const MlpSide* __restrict__ sg = a.side;
    auto bias_elem = [&](int step) -> float {         // element `tslot` of the bias row used by MMA step `step`
      if (step < 8) return __ldg(&sg->bias[step][tslot]);
      if (step == 8) return __ldg(&sg->bias_feat[tslot]);
      return tslot < 128 ? __ldg(&sg->bias_views[tslot]) : 0.f;
    };
    bias_s[tslot] = bias_elem(0);
    slot_sync();

    for (int64_t unit = group; unit < nunits; unit += ngroups) {
      const int64_t m = unit * ROWS_PER_UNIT + (int64_t)g * (TILE_M * CG) + rank * TILE_M + row;
      const bool live = m < a.M;
      // ---- input stage (the column-half-0 thread of each row encodes the point) ----
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (hcol == 0) {
        float px = 0.f, py = 0.f, pz = 0.f;
        if (live) {
          const int64_t r = m / a.S;
          if (a.mode == 0) {
            px = __ldg(a.pts + m * 3); py = __ldg(a.pts + m * 3 + 1); pz = __ldg(a.pts + m * 3 + 2);
            vx = __ldg(a.dirs + r * 3); vy = __ldg(a.dirs + r * 3 + 1); vz = __ldg(a.dirs + r * 3 + 2);
          } else {
            const float* ray = a.rays + r * 11;
            const float z = __ldg(a.z_vals + m);
            px = __fadd_rn(__ldg(ray), __fmul_rn(__ldg(ray + 3), z));        // run_nerf.py:381
            py = __fadd_rn(__ldg(ray + 1), __fmul_rn(__ldg(ray + 4), z));
            pz = __fadd_rn(__ldg(ray + 2), __fmul_rn(__ldg(ray + 5), z));
            vx = __ldg(ray + 8); vy = __ldg(ray + 9); vz = __ldg(ray + 10);
          }
        }
        float f[64];
        encode3<L_PTS>(px, py, pz, f);
        store_row_chunk(pe, row, f);
      }
      signal_a_ready();

      float sigma = 0.f;
      for (int s = 0; s < nsteps; ++s) {
        const unsigned long long te = tracing ? clock64() : 0;
        mbar_wait(ACC_FULL(g), full_phase, abort_flag);
        if (tracing && e == 0) trace_evt(2, 0x3000 | (s << 4) | g, te, clock64(), 0);
        full_phase ^= 1;
        tc_fence_after();
        const bool last = (s == nsteps - 1);
        const float next_bias = bias_elem(last ? 0 : s + 1);      // consumed after this step's drain (latency hidden)
        if (s < 9) {
          const bool relu = (s < 8);
          float sig_acc = 0.f, sig_b = 0.f, sig_c = 0.f, sig_d = 0.f;     // four partial sums: a 32-deep FFMA chain per drain iteration sat on step 7's critical path
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {             // 4 x 32 accumulator columns of this thread's half
            const int col0 = hcol * 128 + cc * 32;
            uint32_t v[32];
            tmem_ld32(tmem_row + col0, v);
            float4 b4[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) b4[j] = reinterpret_cast<const float4*>(bias_s + col0)[j];
            tmem_ld_wait();
            float h[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              h[4 * j] = __uint_as_float(v[4 * j]); h[4 * j + 1] = __uint_as_float(v[4 * j + 1]);
              h[4 * j + 2] = __uint_as_float(v[4 * j + 2]); h[4 * j + 3] = __uint_as_float(v[4 * j + 3]);
              add2(h[4 * j], h[4 * j + 1], b4[j].x, b4[j].y);
              add2(h[4 * j + 2], h[4 * j + 3], b4[j].z, b4[j].w);
            }
            if (s == 7) {                              // alpha_linear on the fp32 activations (run_nerf_helpers.py:110)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 w4 = reinterpret_cast<const float4*>(sd->w_alpha + col0)[j];
                sig_acc = fmaf(fmaxf(h[4 * j], 0.f), w4.x, sig_acc);
                sig_b = fmaf(fmaxf(h[4 * j + 1], 0.f), w4.y, sig_b);
                sig_c = fmaf(fmaxf(h[4 * j + 2], 0.f), w4.z, sig_c);
                sig_d = fmaf(fmaxf(h[4 * j + 3], 0.f), w4.w, sig_d);
              }
            }
            if (AUX && last && dbg_out) {
              if (live) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float4 o = make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
                  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                  reinterpret_cast<float4*>(dbg_out + m * 256 + col0)[j] = o;
                }
              }
            } else {
              const uint32_t cb = act + (col0 >> 6) * CHUNK_BYTES;     // activation K-chunk holding these columns
              const int u0 = (col0 & 63) >> 3;                         // first 16-byte unit inside the 128-byte row
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t addr = cb + row * 128 + ((((u0 + u) ^ (row & 7)) & 7) << 4);
                if (relu)
                  st_shared_v4(addr, pack_bf16_relu(h[8 * u], h[8 * u + 1]), pack_bf16_relu(h[8 * u + 2], h[8 * u + 3]),
                               pack_bf16_relu(h[8 * u + 4], h[8 * u + 5]), pack_bf16_relu(h[8 * u + 6], h[8 * u + 7]));
                else
                  st_shared_v4(addr, pack_bf16(h[8 * u], h[8 * u + 1]), pack_bf16(h[8 * u + 2], h[8 * u + 3]),
                               pack_bf16(h[8 * u + 4], h[8 * u + 5]), pack_bf16(h[8 * u + 6], h[8 * u + 7]));
              }
            }
          }
          // This warp's share of the next A operand is in shared memory: hand it over NOW.  What follows (the sigma
          // combine, the view-direction encoding, the refill of the bias row) touches neither the activation chunks nor,
          // before step 9, the encoding chunk the next MMA reads, so it runs beside that MMA instead of delaying it.
          sig_acc = (sig_acc + sig_b) + (sig_c + sig_d);
          if (tracing && e == 0) trace_evt(2, 0x6000 | (s << 4) | g, clock64(), clock64(), 2);
          if (!last) signal_a_ready();
          if (tracing && e == 0) trace_evt(2, 0x4000 | (s << 4) | g, clock64(), clock64(), 0);
          if (s == 7) {
            // The encoding chunk is free from here (its last reader was the MMA of step 5) until step 9.  First the two
            // column halves of sigma are combined through the row's own last padding columns (logical columns 62, 63),
            // then the column-half-0 thread writes the view-direction features of its row over the whole row (the
            // padding becomes zero again); the arrive at the end of step 8 orders these stores before step 9's MMA.
            float* slot = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + SM_PE + g * CHUNK_BYTES +
                                                   row * 128 + (((7 ^ (row & 7)) & 7) << 4) + 12);
            if (hcol == 1) *slot = sig_acc;
            pair_sync();
            if (hcol == 0) {
              sigma = sig_acc + *slot + sd->b_alpha;
              float f[64];
              encode3<L_DIR>(vx, vy, vz, f);
              store_row_chunk(pe, row, f);
            }
          }
          slot_sync();                                 // every warp of the slot has read this step's bias row
          bias_s[tslot] = next_bias;
          slot_sync();
          if (tracing && e == 0) trace_evt(2, 0x6800 | (s << 4) | g, clock64(), clock64(), 2);
        } else {
          // ---- views layer epilogue: relu(acc + b) . w_rgb -> raw; this thread covers 64 of the 128 columns ----
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int col0 = hcol * 64 + cc * 32;
            uint32_t v[32];
            tmem_ld32(tmem_row + col0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = reinterpret_cast<const float4*>(bias_s + col0)[j];
              const float4 w0 = reinterpret_cast<const float4*>(sd->w_rgb[0] + col0)[j];
              const float4 w1 = reinterpret_cast<const float4*>(sd->w_rgb[1] + col0)[j];
              const float4 w2 = reinterpret_cast<const float4*>(sd->w_rgb[2] + col0)[j];
              const float h0 = fmaxf(__uint_as_float(v[4 * j]) + b4.x, 0.f);
              const float h1 = fmaxf(__uint_as_float(v[4 * j + 1]) + b4.y, 0.f);
              const float h2 = fmaxf(__uint_as_float(v[4 * j + 2]) + b4.z, 0.f);
              const float h3 = fmaxf(__uint_as_float(v[4 * j + 3]) + b4.w, 0.f);
              r0 = fmaf(h0, w0.x, r0); r0 = fmaf(h1, w0.y, r0); r0 = fmaf(h2, w0.z, r0); r0 = fmaf(h3, w0.w, r0);
              r1 = fmaf(h0, w1.x, r1); r1 = fmaf(h1, w1.y, r1); r1 = fmaf(h2, w1.z, r1); r1 = fmaf(h3, w1.w, r1);
              r2 = fmaf(h0, w2.x, r2); r2 = fmaf(h1, w2.y, r2); r2 = fmaf(h2, w2.z, r2); r2 = fmaf(h3, w2.w, r2);
              if (AUX && dbg_out && live) {
                reinterpret_cast<float4*>(dbg_out + m * 256 + col0)[j] = make_float4(h0, h1, h2, h3);
              }
            }
          }
          // combine the halves through the slot's activation buffer (free: every MMA of this tile has retired)
          if (hcol == 1) { scratch_act[row * 4] = r0; scratch_act[row * 4 + 1] = r1; scratch_act[row * 4 + 2] = r2; }
          pair_sync();
          if (hcol == 0 && live) {
            const float4 o = make_float4(r0 + scratch_act[row * 4] + sd->b_rgb[0],
                                         r1 + scratch_act[row * 4 + 1] + sd->b_rgb[1],
                                         r2 + scratch_act[row * 4 + 2] + sd->b_rgb[2], sigma);
            st_stream4(reinterpret_cast<float4*>(a.raw) + m, o);
          }
          pair_sync();                                   // scratch is re-used as the A operand of the next tile
          slot_sync();
          bias_s[tslot] = next_bias;                     // bias row of step 0 for the next unit
          slot_sync();
        }
      }
      // the accumulator has been drained (tcgen05.wait::ld above); signal_a_ready() of the next unit orders it
    }
  }

  tc_fence_before();
  __syncthreads();
  __syncwarp();
  if constexpr (CG == 2) cluster_sync_all();       // the peer may still be reading TMEM the leader's MMAs wrote
  if (warp == 1) {
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace nfb

#include "mlp_train.inl"

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
namespace {
// The head weights / biases of a network are read through constant memory (c_side, MAX_NETS entries per device).  More
// networks than entries may be alive: a network without an entry takes the least recently used one at its next launch
// (the evicted network re-acquires one the same way).  An eviction first waits, on the host, for the last kernel that read
// the victim's entry; entries referenced by a CUDA graph capture are never evicted.
struct SideSlot {
  const nfb_mlp* owner = nullptr;
  cudaEvent_t last_use = nullptr;
  bool event_valid = false;     // last_use was recorded after the owner's latest launch
  bool pinned = false;          // used inside a stream capture: the graph replays with this entry, keep it
  uint64_t tick = 0;
};
std::mutex g_slot_mutex;
SideSlot g_slots[64][nfb::MAX_NETS];
int g_alive[64] = {0};
uint64_t g_tick = 0;

int try_free_cslot(int device, const nfb_mlp* h) {
  for (int i = 0; i < nfb::MAX_NETS; ++i)
    if (!g_slots[device][i].owner) {
      g_slots[device][i].owner = h; g_slots[device][i].event_valid = false; g_slots[device][i].pinned = false;
      g_slots[device][i].tick = ++g_tick;
      return i;
    }
  return -1;
}
void release_cslot(int device, int slot) {
  if (slot >= 0) { g_slots[device][slot].owner = nullptr; g_slots[device][slot].pinned = false; g_slots[device][slot].event_valid = false; }
}
bool is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st != cudaStreamCaptureStatusNone;
}
// Makes sure h owns a constant-memory entry holding its side parameters before a kernel of h is launched on `stream`.
int ensure_cslot(const nfb_mlp* h, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_slot_mutex);
  const int dev = h->device;
  if (h->cslot >= 0) { g_slots[dev][h->cslot].tick = ++g_tick; return NFB_OK; }
  int slot = try_free_cslot(dev, h);
  if (slot < 0) {
    int victim = -1;
    for (int i = 0; i < nfb::MAX_NETS; ++i)
      if (!g_slots[dev][i].pinned && (victim < 0 || g_slots[dev][i].tick < g_slots[dev][victim].tick)) victim = i;
    if (victim < 0)
      return nfb::fail(NFB_E_UNSUPPORTED, "mlp: all %d constant-memory entries are referenced by captured CUDA graphs; destroy a network first", nfb::MAX_NETS);
    if (is_capturing(stream))
      return nfb::fail(NFB_E_UNSUPPORTED, "mlp: more than %d fused networks alive: launch this network once outside the stream capture first", nfb::MAX_NETS);
    SideSlot& v = g_slots[dev][victim];
    cudaError_t e = v.event_valid ? cudaEventSynchronize(v.last_use) : cudaDeviceSynchronize();
    if (e != cudaSuccess) return nfb::fail(NFB_E_CUDA, "mlp: waiting for the evicted network: %s", cudaGetErrorString(e));
    v.owner->cslot = -1;
    v.owner = h; v.event_valid = false; v.tick = ++g_tick;
    slot = victim;
  }
  h->cslot = slot;
  NFB_CUDA(cudaMemcpyToSymbolAsync(nfb::c_side, h->side, sizeof(nfb::MlpSide), (size_t)slot * sizeof(nfb::MlpSide),
                                   cudaMemcpyDeviceToDevice, stream));
  return NFB_OK;
}
// After a launch of h on `stream`: remember when its entry was last read (only needed once entries are contended).
void note_cslot_use(const nfb_mlp* h, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_slot_mutex);
  if (h->cslot < 0) return;
  SideSlot& s = g_slots[h->device][h->cslot];
  if (is_capturing(stream)) { s.pinned = true; return; }
  if (g_alive[h->device] <= nfb::MAX_NETS) { s.event_valid = false; return; }
  if (!s.last_use && cudaEventCreateWithFlags(&s.last_use, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); s.last_use = nullptr; return; }
  s.event_valid = cudaEventRecord(s.last_use, stream) == cudaSuccess;
}
// A barrier time-out of an earlier launch of this network leaves its flag raised (sticky until nfb_mlp_status clears it):
// refuse to run on top of garbage.
int refuse_if_aborted(const nfb_mlp* h, const char* what) {
  if (h->abort_host && *h->abort_host)
    return nfb::fail(NFB_E_CUDA, "%s: an earlier launch of this network hit a pipeline-barrier time-out (its outputs are invalid); "
                                 "nfb_mlp_status() reports and clears it", what);
  return NFB_OK;
}
}  // namespace

extern "C" {

int nfb_mlp_create(nfb_mlp_t** out, int D, int W, int input_ch, int input_ch_views, int skip) {
  NFB_REQUIRE(out, "mlp_create: null out");
  if (D != 8 || W != nfb::W_ || input_ch != nfb::CH_PTS || input_ch_views != nfb::CH_DIR || skip != 4)
    return nfb::fail(NFB_E_UNSUPPORTED,
                     "mlp_create: fused kernel is built for D=8 W=256 input_ch=63 input_ch_views=27 skips=[4] "
                     "(got D=%d W=%d input_ch=%d input_ch_views=%d skip=%d)", D, W, input_ch, input_ch_views, skip);
  nfb_mlp* h = new nfb_mlp();
  h->image = nullptr; h->image_t = nullptr; h->side = nullptr; h->abort_flag = nullptr; h->abort_host = nullptr; h->zero16k = nullptr; h->cslot = -1;
  h->side_stream = nullptr; h->ev_fork = nullptr; h->ev_join = nullptr;
  h->n_params = nfb::param_layout().total;
  cudaError_t e = cudaGetDevice(&h->device);
  if (e == cudaSuccess && (h->device < 0 || h->device >= 64)) { delete h; return nfb::fail(NFB_E_UNSUPPORTED, "mlp_create: device index out of range"); }
  if (e == cudaSuccess) {     // no free constant-memory entry is not an error: the first launch evicts the LRU one
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    h->cslot = try_free_cslot(h->device, h);
    ++g_alive[h->device];
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->image, (size_t)nfb::TOTAL_BLOCKS * nfb::CHUNK_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&h->image_t, (size_t)nfb::tr::TOTAL_BLOCKS_T * nfb::CHUNK_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&h->side, sizeof(nfb::MlpSide));
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->abort_host, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *h->abort_host = 0; e = cudaHostGetDevicePointer((void**)&h->abort_flag, (void*)h->abort_host, 0); }
  if (e == cudaSuccess) e = cudaMalloc(&h->zero16k, nfb::CHUNK_BYTES);
  if (e == cudaSuccess) e = cudaMemset(h->zero16k, 0, nfb::CHUNK_BYTES);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::mlp_fused_fwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::mlp_fused_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::mlp_fused_fwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::mlp_fused_fwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::tr::mlp_train_kernel<nfb::tr::MODE_FWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::tr::mlp_train_kernel<nfb::tr::MODE_FWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::tr::mlp_train_kernel<nfb::tr::MODE_BWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(nfb::tr::mlp_train_kernel<nfb::tr::MODE_BWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::SMEM_BYTES);
  if (e != cudaSuccess) {
    cudaFree(h->image); cudaFree(h->image_t); cudaFree(h->side); cudaFree(h->zero16k);
    if (h->abort_host) cudaFreeHost((void*)h->abort_host);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->device >= 0 && h->device < 64) {
      std::lock_guard<std::mutex> lock(g_slot_mutex);
      release_cslot(h->device, h->cslot);
      --g_alive[h->device];
    }
    delete h;
    return nfb::fail(NFB_E_CUDA, "mlp_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return NFB_OK;
}

int64_t nfb_mlp_param_count(const nfb_mlp_t* h) { return h ? h->n_params : nfb::param_layout().total; }

int nfb_mlp_update(nfb_mlp_t* h, const float* params, int64_t n_params, void* stream) {
  NFB_REQUIRE(h && params, "mlp_update: null pointer");
  NFB_REQUIRE(n_params == h->n_params, "mlp_update: expected %lld parameters, got %lld", (long long)h->n_params, (long long)n_params);
  // NERFAIL_B200_PACK=split: the two one-element-per-thread kernels (the readable specification of the layouts) instead of
  // the fused 16-bytes-per-thread kernel; the images are bit-identical (tests/test_gpu_mlp.py)
  const char* pack_env = getenv("NERFAIL_B200_PACK");
  const bool split = pack_env && pack_env[0] == 's';
  int rc;
  if (split) {
    nfb::pack_weights_kernel<<<nfb::sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(params, h->image, h->side);
    rc = nfb::check_launch("mlp_update");
    if (rc) return rc;
    nfb::tr::pack_weights_T_kernel<<<nfb::sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(params, h->image_t);
    rc = nfb::check_launch("mlp_update.transposed");
  } else {
    const int units = (nfb::TOTAL_BLOCKS + nfb::tr::TOTAL_BLOCKS_T) * nfb::TILE_M * 8;
    nfb::tr::pack_both_kernel<<<(units + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, h->image, h->image_t, h->side);
    rc = nfb::check_launch("mlp_update");
  }
  if (rc) return rc;
  // stream-ordered device-to-device copy of the fp32 side parameters into this network's constant-memory entry (a network
  // that holds none right now copies them when it acquires one, ensure_cslot)
  std::lock_guard<std::mutex> lock(g_slot_mutex);
  if (h->cslot >= 0)
    NFB_CUDA(cudaMemcpyToSymbolAsync(nfb::c_side, h->side, sizeof(nfb::MlpSide), (size_t)h->cslot * sizeof(nfb::MlpSide),
                                     cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return NFB_OK;
}

int nfb_mlp_destroy(nfb_mlp_t* h) {
  if (!h) return NFB_OK;
  cudaFree(h->image); cudaFree(h->image_t); cudaFree(h->side); cudaFree(h->zero16k);
  if (h->abort_host) cudaFreeHost((void*)h->abort_host);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  {
    std::lock_guard<std::mutex> lock(g_slot_mutex);
    release_cslot(h->device, h->cslot);
    --g_alive[h->device];
  }
  delete h;
  return NFB_OK;
}

// 0 = healthy; 1 = a barrier wait inside the fused kernel timed out (results invalid).  Synchronises the device.
int nfb_mlp_status(nfb_mlp_t* h) {
  NFB_REQUIRE(h, "mlp_status: null handle");
  NFB_CUDA(cudaDeviceSynchronize());
  if (*h->abort_host) {
    *h->abort_host = 0;
    return nfb::fail(NFB_E_CUDA, "mlp_fwd: pipeline barrier timed out inside the fused kernel");
  }
  return NFB_OK;
}

// The same flag WITHOUT synchronising or clearing: what the kernels that have finished so far reported.  Costs one read of
// pinned host memory, so product paths poll it at their natural hand-over points (a finished view, optimizer.step).
int nfb_mlp_poll(const nfb_mlp_t* h) {
  NFB_REQUIRE(h, "mlp_poll: null handle");
  return refuse_if_aborted(h, "mlp_poll");
}

// Test hook: raise the time-out flag from the host (what a kernel does when a barrier wait expires), so the reporting
// contract — launches refuse, nfb_mlp_poll reports, nfb_mlp_status reports and clears — can be exercised without a hang.
int nfb_mlp_debug_raise_abort(nfb_mlp_t* h) {
  NFB_REQUIRE(h && h->abort_host, "mlp_debug_raise_abort: null handle");
  *h->abort_host = 1;
  return NFB_OK;
}

static int mlp_launch(const nfb_mlp_t* h, int mode, const float* pts, const float* dirs, const float* rays,
                      const float* z_vals, int R, int S, float* raw, int nsteps, float* dbg, void* stream,
                      unsigned long long* trace = nullptr) {
  NFB_REQUIRE(h && raw, "mlp_fwd: null handle or output");
  NFB_REQUIRE(R >= 0 && S > 0, "mlp_fwd: R=%d S=%d", R, S);
  NFB_REQUIRE(mode == 0 ? (pts && dirs) : (mode == 1 && rays && z_vals), "mlp_fwd: inputs missing for mode %d", mode);
  NFB_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "mlp_fwd: raw must be 16-byte aligned");
  NFB_REQUIRE(nsteps >= 1 && nsteps <= nfb::NSTEP && (nsteps == nfb::NSTEP || dbg), "mlp_fwd: bad nsteps");
  if (R == 0) return NFB_OK;
  int rc0 = refuse_if_aborted(h, "mlp_fwd");
  if (rc0 == NFB_OK) rc0 = ensure_cslot(h, (cudaStream_t)stream);
  if (rc0 != NFB_OK) return rc0;
  nfb::FwdArgs a;
  a.image = h->image; a.side = h->side; a.cslot = h->cslot; a.abort_flag = h->abort_flag;
  a.mode = mode; a.pts = pts; a.dirs = dirs; a.rays = rays; a.z_vals = z_vals;
  a.M = (int64_t)R * S; a.S = S; a.raw = raw; a.nsteps = nsteps; a.dbg = dbg; a.trace = trace;
  // NERFAIL_B200_CG=1 selects the single-CTA variant (cta_group::1); default is the CTA-pair kernel (cta_group::2).
  static const int cg = []() { const char* e = getenv("NERFAIL_B200_CG"); return (e && e[0] == '1') ? 1 : 2; }();
  const int64_t rows_per_unit = 2 * nfb::TILE_M * cg;
  const int64_t nunits = (a.M + rows_per_unit - 1) / rows_per_unit;
  int groups = nfb::sm_count() / cg;
  if (nunits < groups) groups = (int)nunits;
  const bool aux = dbg != nullptr || trace != nullptr || nsteps != nfb::NSTEP;
  if (cg == 1) {
    if (aux) nfb::mlp_fused_fwd_kernel<1, true><<<groups, nfb::NUM_THREADS, nfb::SMEM_BYTES, (cudaStream_t)stream>>>(a);
    else nfb::mlp_fused_fwd_kernel<1, false><<<groups, nfb::NUM_THREADS, nfb::SMEM_BYTES, (cudaStream_t)stream>>>(a);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(groups * 2));
    cfg.blockDim = dim3(nfb::NUM_THREADS);
    cfg.dynamicSmemBytes = nfb::SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = aux ? cudaLaunchKernelEx(&cfg, nfb::mlp_fused_fwd_kernel<2, true>, a)
                        : cudaLaunchKernelEx(&cfg, nfb::mlp_fused_fwd_kernel<2, false>, a);
    if (e != cudaSuccess) return nfb::fail(NFB_E_CUDA, "mlp_fwd: cluster launch: %s", cudaGetErrorString(e));
  }
  note_cslot_use(h, (cudaStream_t)stream);
  return nfb::check_launch("mlp_fwd");
}

int nfb_mlp_fwd(const nfb_mlp_t* h, int mode, const float* pts, const float* dirs,
                const float* rays, const float* z_vals, int R, int S, float* raw, void* stream) {
  return mlp_launch(h, mode, pts, dirs, rays, z_vals, R, S, raw, nfb::NSTEP, nullptr, stream);
}

// Debug/validation entry: run only the first `nsteps` MMA steps and dump the fp32 post-activation values of
// the last one to dbg [R*S,256] (first 128 columns for the view layer).  raw is written only when nsteps == 10.
int nfb_mlp_fwd_debug(const nfb_mlp_t* h, int mode, const float* pts, const float* dirs,
                      const float* rays, const float* z_vals, int R, int S, float* raw, int nsteps, float* dbg,
                      void* stream) {
  return mlp_launch(h, mode, pts, dirs, rays, z_vals, R, S, raw, nsteps, dbg, stream);
}

// ---- training (bf16 tensor-core forward that saves activations, and the data-gradient chain) ----
static int train_launch(int mode, const nfb_mlp_t* h, nfb::tr::TrainArgs& a, void* stream, int max_groups = 0) {
  static const int skip = []() { const char* e = getenv("NERFAIL_B200_TRAIN_SKIP"); return e ? atoi(e) : 0; }();
  a.skip = skip;
  int rc0 = refuse_if_aborted(h, mode == nfb::tr::MODE_FWD ? "mlp_fwd_train" : "mlp_bwd_data");
  if (rc0 == NFB_OK) rc0 = ensure_cslot(h, (cudaStream_t)stream);
  if (rc0 != NFB_OK) return rc0;
  a.cslot = h->cslot;
  const int64_t rows_per_unit = 2 * nfb::TILE_M * 2;
  const int64_t nunits = (a.M + rows_per_unit - 1) / rows_per_unit;
  int groups = nfb::sm_count() / 2;
  if (max_groups > 0 && max_groups < groups) groups = max_groups;
  if (nunits < groups) groups = (int)nunits;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(groups * 2));
  cfg.blockDim = dim3(nfb::NUM_THREADS);
  cfg.dynamicSmemBytes = nfb::SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static unsigned long long* wait_dbg = []() -> unsigned long long* {       // NERFAIL_B200_TRAIN_WAITS=1: wait-cycle counters
    const char* e = getenv("NERFAIL_B200_TRAIN_WAITS");
    if (!(e && e[0] == '1')) return nullptr;
    unsigned long long* p = nullptr;
    cudaMalloc(&p, 4 * sizeof(unsigned long long));
    return p;
  }();
  a.wait_cycles = wait_dbg;
  if (wait_dbg) cudaMemsetAsync(wait_dbg, 0, 4 * sizeof(unsigned long long), (cudaStream_t)stream);
  const bool aux = a.skip != 0 || a.ready != nullptr || wait_dbg != nullptr;
  cudaError_t e = (mode == nfb::tr::MODE_FWD)
      ? (aux ? cudaLaunchKernelEx(&cfg, nfb::tr::mlp_train_kernel<nfb::tr::MODE_FWD, true>, a)
             : cudaLaunchKernelEx(&cfg, nfb::tr::mlp_train_kernel<nfb::tr::MODE_FWD, false>, a))
      : (aux ? cudaLaunchKernelEx(&cfg, nfb::tr::mlp_train_kernel<nfb::tr::MODE_BWD, true>, a)
             : cudaLaunchKernelEx(&cfg, nfb::tr::mlp_train_kernel<nfb::tr::MODE_BWD, false>, a));
  if (e != cudaSuccess) return nfb::fail(NFB_E_CUDA, "mlp_train: cluster launch: %s", cudaGetErrorString(e));
  note_cslot_use(h, (cudaStream_t)stream);
  if (wait_dbg) {
    unsigned long long h[4];
    cudaMemcpy(h, wait_dbg, sizeof h, cudaMemcpyDeviceToHost);
    const double n = groups, T = (double)h[3];
    fprintf(stderr, "%s waits (share of CTA 0's %.0f cycles): MMA warp A_READY %.1f %%, W_FULL %.1f %%; epilogue warp ACC_FULL %.1f %%\n",
            mode == nfb::tr::MODE_FWD ? "train fwd" : "bwd data", T, 100.0 * h[0] / n / T, 100.0 * h[1] / n / T, 100.0 * h[2] / (2.0 * n) / T);
  }
  return nfb::check_launch(mode == nfb::tr::MODE_FWD ? "mlp_fwd_train" : "mlp_bwd_data");
}

// number of 128-row tiles the training images must hold for M = R*S samples (whole units of 512 rows)
int64_t nfb_mlp_train_tiles(int64_t M) { return M <= 0 ? 0 : ((M + 511) / 512) * 4; }

int nfb_mlp_fwd_train(const nfb_mlp_t* h, const float* rays, const float* z_vals, int R, int S, float* raw,
                      void* act_img, uint32_t* mask, void* stream) {
  NFB_REQUIRE(h && rays && z_vals && raw && act_img && mask, "mlp_fwd_train: null pointer");
  NFB_REQUIRE(R >= 0 && S > 0, "mlp_fwd_train: R=%d S=%d", R, S);
  if (R == 0) return NFB_OK;
  nfb::tr::TrainArgs a{};
  a.image = h->image; a.side = h->side; a.cslot = h->cslot; a.abort_flag = h->abort_flag;
  a.rays = rays; a.z_vals = z_vals; a.M = (int64_t)R * S; a.S = S; a.raw = raw;
  a.act_img = (char*)act_img; a.mask = mask;
  return train_launch(nfb::tr::MODE_FWD, h, a, stream);
}

int nfb_mlp_bwd_data(const nfb_mlp_t* h, const float* g_raw, int64_t M, const uint32_t* mask, void* dy_img, void* stream) {
  NFB_REQUIRE(h && g_raw && mask && dy_img, "mlp_bwd_data: null pointer");
  NFB_REQUIRE(M >= 0, "mlp_bwd_data: M=%lld", (long long)M);
  if (M == 0) return NFB_OK;
  nfb::tr::TrainArgs a{};
  a.image = h->image_t; a.side = h->side; a.cslot = h->cslot; a.abort_flag = h->abort_flag;
  a.M = M; a.S = 1; a.g_raw = g_raw; a.mask = const_cast<uint32_t*>(mask); a.dy_img = (char*)dy_img;
  return train_launch(nfb::tr::MODE_BWD, h, a, stream);
}

// The products dW = dY^T X (+ bias sums) of one network over the saved images, gradient in state_dict order.
// need[j] = value of the data-gradient kernel's per-tile ready counter from which job j's dY chunks of that tile are
// complete (mlp_train.inl: 1 = input stage (view-layer dY), b + 2 = output of backward step b).
static int wgrad_jobs(const void* act_img, const void* dy_img, float* grad, nfb::WgradJob* jobs, signed char* need) {
  using namespace nfb;
  const ParamLayout pl = param_layout();
  const int64_t ap = (int64_t)tr::FWD_CHUNKS * CHUNK_BYTES, dp = (int64_t)tr::BWD_CHUNKS * CHUNK_BYTES;
  auto A = [&](int chunk) { return (const void*)((const char*)act_img + (int64_t)chunk * CHUNK_BYTES); };
  auto D = [&](int chunk) { return (const void*)((const char*)dy_img + (int64_t)chunk * CHUNK_BYTES); };
  auto dYl = [&](int l) { return D(6 + 4 * (7 - l)); };       // output of backward step 8 - l
  auto job = [&](const void* dy, const void* x, float* gw, float* gb, int ndy, int nx, int ld, int col0, int cols, int rows) {
    WgradJob j{};
    j.dy = dy; j.dy_pitch = dp; j.x = x; j.x_pitch = ap; j.out_w = gw; j.out_b = gb;
    j.ndy = ndy; j.ndy_real = ndy; j.nx = nx; j.ld = ld; j.col0 = col0; j.cols_valid = cols; j.row_begin = 0; j.row_end = rows;
    return j;
  };
  int n = 0;
  // pts_linears.l, heaviest first so the equal-cost cut starts on full-width products
  for (int l = 1; l < 8; ++l) {
    const int ld = (l == 5) ? W_ + CH_PTS : W_;
    need[n] = (signed char)(10 - l);
    jobs[n++] = job(dYl(l), A(4 * (l - 1)), grad + pl.w_pts[l], grad + pl.b_pts[l], 4, 4, ld, l == 5 ? CH_PTS : 0, W_, W_);
  }
  // feature_linear (X = h_7); alpha_linear rides on it as a side product: dW_alpha = sum_rows g_sigma * h_7
  need[n] = 2; jobs[n] = job(D(2), A(28), grad + pl.w_feat, grad + pl.b_feat, 4, 4, W_, 0, W_, W_);
  jobs[n].side_rows = 1; jobs[n].side_gcol = 3; jobs[n].side_cols = W_; jobs[n].side_ld = W_;
  jobs[n].out_side_w = grad + pl.w_alpha; jobs[n].out_side_b = grad + pl.b_alpha;
  ++n;
  need[n] = 1; jobs[n++] = job(D(0), A(tr::IMG_FEAT), grad + pl.w_views, grad + pl.b_views, 2, 4, W_ + CH_DIR, 0, W_, 128);             // views_linears.0 [:, :256]
  need[n] = 10; jobs[n++] = job(dYl(0), A(tr::IMG_PE), grad + pl.w_pts[0], grad + pl.b_pts[0], 4, 1, CH_PTS, 0, CH_PTS, W_);            // pts_linears.0
  need[n] = 5; jobs[n++] = job(dYl(5), A(tr::IMG_PE), grad + pl.w_pts[5], nullptr, 4, 1, W_ + CH_PTS, 0, CH_PTS, W_);                   // pts_linears.5 [:, :63] (skip input)
  // views_linears.0 [:, 256:] (X = encoded direction); rgb_linear rides on it: its operand hv is loaded as two extra
  // slices, dW_rgb = sum_rows g_rgb * hv
  need[n] = 1; jobs[n] = job(D(0), A(tr::IMG_DIR), grad + pl.w_views, nullptr, 2, 1, W_ + CH_DIR, W_, CH_DIR, 128);
  jobs[n].side_rows = 3; jobs[n].side_gcol = 0; jobs[n].side_cols = 128; jobs[n].side_ld = 128; jobs[n].side_nx = 2;
  jobs[n].side_x = A(tr::IMG_HV); jobs[n].out_side_w = grad + pl.w_rgb; jobs[n].out_side_b = grad + pl.b_rgb;
  ++n;
  return n;
}

// Weight gradients of one network from the saved images: 14 tensor-core products dW = dY^T X (+ bias sums), the two head
// products as CUDA-core side sums of g_raw [M,4] against operands two of them load anyway, in ONE grouped launch, accumulated into grad [n_params] in state_dict order (the caller zeroes it once per step).
int nfb_mlp_bwd_weights(const nfb_mlp_t* h, const void* act_img, const void* dy_img, const float* g_raw, int64_t M,
                        float* grad, void* stream) {
  NFB_REQUIRE(h && act_img && dy_img && g_raw && grad, "mlp_bwd_weights: null pointer");
  NFB_REQUIRE(M >= 0, "mlp_bwd_weights: M=%lld", (long long)M);
  if (M == 0) return NFB_OK;
  const int64_t ntiles = nfb_mlp_train_tiles(M);
  nfb::WgradJob jobs[16];
  signed char need[16];
  const int n = wgrad_jobs(act_img, dy_img, grad, jobs, need);
  const int rc0 = refuse_if_aborted(h, "mlp_bwd_weights");
  if (rc0 != NFB_OK) return rc0;
  return nfb::launch_wgrad_grouped(jobs, n, ntiles, h->zero16k, h->abort_flag, stream, "mlp_bwd_weights", nullptr, nullptr, 0, g_raw, M);
}

// Whole backward of one network: the data-gradient chain and the grouped weight-gradient kernel run CONCURRENTLY on
// disjoint SMs (producer clusters on `stream`, consumer CTAs on the handle's side stream, fork / join by events).  The
// producer publishes per tile how many of its dY store groups have landed (ready [tiles] int32, zeroed here); a consumer
// CTA loads a tile's dY as soon as its product's chunks are complete, i.e. out of L2 while the lines are still resident,
// so the 5 KB / sample of dY is written to HBM once and never read back from it.  Results are identical to
// nfb_mlp_bwd_data followed by nfb_mlp_bwd_weights up to the order of the fp32 reductions.
// NERFAIL_B200_BWD_PRODUCERS = number of producer CTA pairs (default 40 of 74); batches smaller than that run serially.
int nfb_mlp_bwd(const nfb_mlp_t* h, const float* g_raw, int64_t M, const uint32_t* mask, const void* act_img,
                void* dy_img, float* grad, int* ready, void* stream) {
  NFB_REQUIRE(h && g_raw && mask && act_img && dy_img && grad && ready, "mlp_bwd: null pointer");
  NFB_REQUIRE(M >= 0, "mlp_bwd: M=%lld", (long long)M);
  if (M == 0) return NFB_OK;
  static const int env_groups = []() { const char* e = getenv("NERFAIL_B200_BWD_PRODUCERS"); return e ? atoi(e) : 40; }();
  const int64_t ntiles = nfb_mlp_train_tiles(M);
  const int pairs = nfb::sm_count() / 2;
  int groups = env_groups;
  const int consumers = 2 * (pairs - groups);
  if (groups <= 0 || consumers < 16 || ntiles / 4 < groups || ntiles >= 32768) {
    int rc = nfb_mlp_bwd_data(h, g_raw, M, mask, dy_img, stream);
    if (rc != NFB_OK) return rc;
    return nfb_mlp_bwd_weights(h, act_img, dy_img, g_raw, M, grad, stream);
  }
  cudaStream_t main_s = (cudaStream_t)stream;
  // profiling only (NERFAIL_B200_BWD_MODE): 1 = producer alone on its CTA pairs, 2 = consumer alone on its CTAs (ready left
  // at 10 by an earlier call), 3 = both, one after the other on `stream`
  static const int dbg_mode = []() { const char* e = getenv("NERFAIL_B200_BWD_MODE"); return e ? atoi(e) : 0; }();
  cudaStream_t cons_s = (dbg_mode == 0) ? h->side_stream : main_s;
  if (dbg_mode != 2) NFB_CUDA(cudaMemsetAsync(ready, 0, sizeof(int) * ntiles, main_s));
  if (dbg_mode == 0) {
    NFB_CUDA(cudaEventRecord(h->ev_fork, main_s));
    NFB_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
  }
  nfb::tr::TrainArgs a{};
  a.image = h->image_t; a.side = h->side; a.cslot = h->cslot; a.abort_flag = h->abort_flag;
  a.M = M; a.S = 1; a.g_raw = g_raw; a.mask = const_cast<uint32_t*>(mask); a.dy_img = (char*)dy_img; a.ready = ready;
  int rc = (dbg_mode == 2) ? NFB_OK : train_launch(nfb::tr::MODE_BWD, h, a, stream, groups);
  nfb::WgradJob jobs[16];
  signed char need[16];
  const int n = wgrad_jobs(act_img, dy_img, grad, jobs, need);
  if (rc == NFB_OK && dbg_mode != 1)
    rc = nfb::launch_wgrad_grouped(jobs, n, ntiles, h->zero16k, h->abort_flag, cons_s, "mlp_bwd", ready, need, consumers, g_raw, M);
  if (dbg_mode != 0) return rc;
  // join even after a failed launch so the caller's stream never runs ahead of the side stream
  cudaEventRecord(h->ev_join, h->side_stream);
  cudaStreamWaitEvent(main_s, h->ev_join, 0);
  return rc;
}

// Profiling aid: full forward with a timeline of CTA 0 written to trace [3][2048][4] uint64 (see FwdArgs::trace).
int nfb_mlp_fwd_trace(const nfb_mlp_t* h, const float* rays, const float* z_vals, int R, int S, float* raw,
                      unsigned long long* trace, void* stream) {
  return mlp_launch(h, 1, nullptr, nullptr, rays, z_vals, R, S, raw, nfb::NSTEP, nullptr, stream, trace);
}

}  // extern "C"
