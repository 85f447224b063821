// Training kernels of the fused MLP (included at the end of mlp_fused.cu: same translation unit, same helpers).
//
//   MODE_FWD : the inference pipeline of mlp_fused_fwd_kernel<2> that additionally leaves in HBM, per 128-row tile,
//              the bf16 activation of every layer as a "tile image" (bulk-stored straight out of the shared-memory
//              A operand: cp.async.bulk.global.shared::cta) and the relu masks as bit words.
//   MODE_BWD : the data-gradient chain as the same pipeline run backwards: A = dY (bf16, shared memory), B = the
//              TRANSPOSED weight blocks, epilogue = relu mask (bit words) instead of bias + relu; every dY is also
//              bulk-stored as a tile image for the weight-gradient GEMMs (wgrad.cu).
// Reference: autograd of run_nerf_helpers.py:100-123 under loss.backward() (run_nerf.py:791).  No gradient flows to the
// encoded inputs (points and directions are constants of the step), exactly as SURVEY.md §8d counts the dgrad FLOPs.
//
// Tile-image chunk indices (16 KB chunks, 128 rows x 64 columns, 128-byte swizzle):
//   forward  : h_s (s = 0..7) at 4 s .. 4 s + 3, feature at 32..35, view-layer output at 36..37, encoded point 38,
//              encoded direction 39                                            -> FWD_CHUNKS = 40 per tile
//   backward : dY_views 0..1, dY_feature 2..5, dY_l (l = 7..0) at 6 + 4 (7 - l); chunk 38 is reserved and not written (it
//              held dY of the two heads before their weight gradients became side sums of g_raw) -> BWD_CHUNKS = 39 per tile
//   masks    : [tile][9][8 words][128 rows]: words of h_0..h_7 (index 0..7) and of the view layer (index 8, 4 words);
//              word w of a row covers columns 32 w .. 32 w + 31 (a warp stores / loads 128 contiguous bytes).
//              h_0..h_7: PAIR layout — column 32 w + 2 p + 1 = bit 31 - p, column 32 w + 2 p = bit 15 - p, so that
//              (word << p) carries the two bits of the bf16 pair p in the sign positions of its two halves and ONE
//              byte-permute with sign replication (PRMT 0xBB99) turns them into the AND mask of the packed pair
//              (1.5 instructions per column in the data-gradient drain instead of 3).  View layer (index 8): column
//              32 w + j = bit 31 - j (its consumer works on fp32 values, one column at a time).

namespace nfb {
namespace tr {

constexpr int MODE_FWD = 1, MODE_BWD = 2;
constexpr int FWD_CHUNKS = 40, BWD_CHUNKS = 39;
constexpr int IMG_DY_HEAD = 38;
constexpr int IMG_FEAT = 32, IMG_HV = 36, IMG_PE = 38, IMG_DIR = 39;
constexpr int MASK_WORDS_PER_TILE = 9 * 128 * 8;
constexpr int BWD_STEPS = 9;
constexpr int TOTAL_BLOCKS_T = 4 + 8 * 8;     // transposed image: views^T (2 K-chunks) + feature^T + W_7^T .. W_1^T

__host__ __device__ constexpr int dy_chunk0(int b) { return b == 0 ? 2 : 6 + 4 * (b - 1); }   // output image chunk of bwd step b

struct TrainArgs {
  const __nv_bfloat16* image;     // forward blocks (MODE_FWD) or transposed blocks (MODE_BWD)
  const MlpSide* side;
  int cslot;
  int* abort_flag;
  unsigned long long* wait_cycles; // profiling only (AUX instantiation): [0] += cycles the MMA warps waited for A_READY, [1] for W_FULL,
                                  // [2] += cycles the epilogue warps of slot 0 / quarter 0 waited for ACC_FULL, [3] += kernel cycles of CTA 0
  const float* rays;              // [R,11]           (MODE_FWD)
  const float* z_vals;            // [M]              (MODE_FWD)
  int64_t M;
  int S;
  float* raw;                     // [M,4] out        (MODE_FWD)
  const float* g_raw;             // [M,4] in         (MODE_BWD)
  char* act_img;                  // [ntiles][40][16 KB]  written by MODE_FWD
  uint32_t* mask;                 // [ntiles][9][8][128]  written by MODE_FWD, read by MODE_BWD
  int* ready;                     // [ntiles] or null (MODE_BWD): published count of landed store groups per tile, for a
                                  // concurrently running wgrad_kernel in consumer mode: 1 = dY_views,
                                  // 1 + b = output of backward step b - 1 as well (10 = everything)
  int skip;                       // profiling only (NERFAIL_B200_TRAIN_SKIP): bit 0 = no mask stores, bit 1 = no image stores,
                                  // bit 2 = image stores are not waited for before their source is overwritten, bit 3 = every CTA
                                  // rewrites the same two tile images, so the stores never leave L2 (both WRONG results: timing only)
  char* dy_img;                   // [ntiles][39][16 KB]  written by MODE_BWD
};

// transposed weight image for the data-gradient chain
__global__ void pack_weights_T_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ image) {
  const ParamLayout pl = param_layout();
  const int64_t total = (int64_t)TOTAL_BLOCKS_T * TILE_M * KCH;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int blk = (int)(e / (TILE_M * KCH));
    const int within = (int)(e % (TILE_M * KCH));
    const int nrow = within / KCH, kk = within % KCH;
    int b, local;
    if (blk < 4) { b = 0; local = blk; } else { b = 1 + (blk - 4) / 8; local = (blk - 4) % 8; }
    const int chunk = local / 2, half = local % 2;
    const int n = half * 128 + nrow;          // input-feature index of the forward layer (row of the transposed block)
    const int k = chunk * KCH + kk;           // output-feature index of the forward layer
    float v;
    if (b == 0) v = params[pl.w_views + (int64_t)k * (W_ + CH_DIR) + n];
    else if (b == 1) v = params[pl.w_feat + (int64_t)k * W_ + n];
    else {
      const int l = 9 - b;                    // 7 .. 1
      const int in = (l == 5) ? W_ + CH_PTS : W_;
      v = params[pl.w_pts[l] + (int64_t)k * in + n + (l == 5 ? CH_PTS : 0)];
    }
    char* dst = reinterpret_cast<char*>(image) + (int64_t)blk * CHUNK_BYTES + swz_off(nrow, kk);
    *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(v);
  }
}

// Both weight images and the fp32 side parameters of one network in ONE launch (nfb_mlp_update): a thread produces one
// 16-byte unit (8 bf16) of an image.  Forward image: unit-fastest mapping (8 lanes read 64 consecutive floats of a weight
// row); transposed image: row-fastest mapping (32 lanes read 32 consecutive input features of one output row), so both
// read the fp32 parameters coalesced.  Same bits as pack_weights_kernel / pack_weights_T_kernel (kept as the readable
// specification and used by the tests' cross-check); 4-5x faster, which matters once a step is a few hundred microseconds.
__device__ __forceinline__ float fwd_image_value(const float* __restrict__ params, const ParamLayout& pl, int s, int chunk, int n, int kk) {
  if (s <= 7) {
    const int in = (s == 0) ? CH_PTS : (s == 5 ? W_ + CH_PTS : W_);
    int col;
    if (s == 0) col = (kk < CH_PTS) ? kk : -1;
    else if (s == 5) col = (chunk == 4) ? ((kk < CH_PTS) ? kk : -1) : CH_PTS + chunk * KCH + kk;
    else col = chunk * KCH + kk;
    return col >= 0 ? __ldg(params + pl.w_pts[s] + (int64_t)n * in + col) : 0.f;
  }
  if (s == 8) return __ldg(params + pl.w_feat + (int64_t)n * W_ + chunk * KCH + kk);
  const int col = chunk < 4 ? chunk * KCH + kk : ((kk < CH_DIR) ? W_ + kk : -1);
  return col >= 0 ? __ldg(params + pl.w_views + (int64_t)n * (W_ + CH_DIR) + col) : 0.f;
}

__global__ void __launch_bounds__(256)
pack_both_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ image, __nv_bfloat16* __restrict__ image_t,
                 MlpSide* __restrict__ side) {
  const ParamLayout pl = param_layout();
  constexpr int UNITS_PER_BLOCK = TILE_M * 8;
  const int total = (TOTAL_BLOCKS + TOTAL_BLOCKS_T) * UNITS_PER_BLOCK;
  for (int uid = blockIdx.x * blockDim.x + threadIdx.x; uid < total; uid += gridDim.x * blockDim.x) {
    int blk = uid / UNITS_PER_BLOCK;
    const int within = uid % UNITS_PER_BLOCK;
    float v[8];
    char* dst;
    int nrow, unit;
    if (blk < TOTAL_BLOCKS) {
      nrow = within >> 3; unit = within & 7;
      int s = 0;
#pragma unroll
      for (int t = 1; t < NSTEP; ++t) s += (blk >= step_block0(t)) ? 1 : 0;
      const int local = blk - step_block0(s);
      const int halves = step_halves(s);
      const int chunk = local / halves, half = local % halves;
      const int n = half * 128 + nrow;
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = fwd_image_value(params, pl, s, chunk, n, unit * 8 + q);
      dst = reinterpret_cast<char*>(image) + (int64_t)blk * CHUNK_BYTES;
    } else {
      blk -= TOTAL_BLOCKS;
      nrow = within & 127; unit = within >> 7;
      int b, local;
      if (blk < 4) { b = 0; local = blk; } else { b = 1 + (blk - 4) / 8; local = (blk - 4) % 8; }
      const int chunk = local / 2, half = local % 2;
      const int n = half * 128 + nrow;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = chunk * KCH + unit * 8 + q;
        if (b == 0) v[q] = __ldg(params + pl.w_views + (int64_t)k * (W_ + CH_DIR) + n);
        else if (b == 1) v[q] = __ldg(params + pl.w_feat + (int64_t)k * W_ + n);
        else {
          const int l = 9 - b;
          const int in = (l == 5) ? W_ + CH_PTS : W_;
          v[q] = __ldg(params + pl.w_pts[l] + (int64_t)k * in + n + (l == 5 ? CH_PTS : 0));
        }
      }
      dst = reinterpret_cast<char*>(image_t) + (int64_t)blk * CHUNK_BYTES;
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + nrow * 128 + (((unit ^ (nrow & 7)) & 7) << 4)) = o;
  }
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

__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_but_last() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all_but_two() { asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); }
// the bulk stores counted so far have landed (wait_group without .read): make them visible device-wide, then publish
__device__ __forceinline__ void publish_ready(int* p, int count) {
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __threadfence();
  asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(count) : "memory");
}

// AUX = true: the instantiation with the profiling switches (TrainArgs::skip) and the ready counters of the overlapped
// backward (TrainArgs::ready); the production instantiation has both compiled out of the epilogue.
template <int MODE, bool AUX = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_train_kernel(const TrainArgs a) {
  const int skip = AUX ? a.skip : 0;
  int* const ready = AUX ? a.ready : nullptr;
  constexpr int CG = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t bar0 = base + SM_BAR;
  auto W_FULL = [&](int s) { return bar0 + 8 * s; };
  auto W_EMPTY = [&](int s) { return bar0 + 8 * (NSTAGE + s); };
  auto A_READY = [&](int g) { return bar0 + 8 * (2 * NSTAGE + g); };
  auto ACC_FULL = [&](int g) { return bar0 + 8 * (2 * NSTAGE + 2 + g); };
  const uint32_t tmem_slot = bar0 + 8 * (2 * NSTAGE + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + SM_BAR + 8 * (2 * NSTAGE + 4));

  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  volatile int* abort_flag = a.abort_flag;
  if ((base & 1023u) != 0u && threadIdx.x == 0) *abort_flag = 1;
  constexpr int nsteps = (MODE == MODE_FWD) ? NSTEP : BWD_STEPS;
  auto kchunks = [](int s) { return MODE == MODE_FWD ? step_kchunks(s) : (s == 0 ? 2 : 4); };
  auto halves_of = [](int s) { return MODE == MODE_FWD ? step_halves(s) : 2; };
  auto block0 = [](int s) { return MODE == MODE_FWD ? step_block0(s) : (s == 0 ? 0 : 4 + (s - 1) * 8); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(W_FULL(s), leader ? 2 : 1); mbar_init(W_EMPTY(s), 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(A_READY(g), 8 * CG); mbar_init(ACC_FULL(g), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  constexpr int ROWS_PER_UNIT = 2 * TILE_M * CG;
  const int64_t nunits = (a.M + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT;
  const int64_t group = blockIdx.x / CG, ngroups = gridDim.x / CG;

  if (warp == 0) {
    // ================= weight producer =================
    uint32_t pos = 0;
    for (int64_t unit = group; unit < nunits; unit += ngroups) {
      for (int s = 0; s < nsteps; ++s) {
        const int kch = kchunks(s), halves = halves_of(s);
        const char* src = reinterpret_cast<const char*>(a.image) + (int64_t)block0(s) * CHUNK_BYTES;
        const uint32_t bytes = (halves == 2) ? CHUNK_BYTES : CHUNK_BYTES / 2;
        for (int c = 0; c < kch; ++c, ++pos) {
          const int stage = pos & (NSTAGE - 1);
          const char* blk = (halves == 2) ? src + (int64_t)(c * 2 + rank) * CHUNK_BYTES
                                          : src + (int64_t)c * CHUNK_BYTES + rank * (CHUNK_BYTES / 2);
          mbar_wait(W_EMPTY(stage), ((pos >> 2) & 1) ^ 1, abort_flag);
          if (elect_one()) {
            mbar_arrive_expect_tx(W_FULL(stage), bytes);
            bulk_g2s(base + SM_W + stage * CHUNK_BYTES, blk, bytes, W_FULL(stage));
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t dlo_or = 1u << 16;
    const uint32_t dhi = (uint32_t)(umma_desc(0) >> 32);
    auto desc_of = [&](uint32_t saddr) { return ((uint64_t)dhi << 32) | (uint64_t)(((saddr & 0x3FFFF) >> 4) | dlo_or); };
    if (leader) {
      // ================= MMA issuer =================
      uint32_t pos = 0, ready_phase0 = 0, ready_phase1 = 0;
      long long w_a = 0, w_w = 0;
      const long long t_begin = AUX ? clock64() : 0;
      for (int64_t unit = group; unit < nunits; unit += ngroups) {
        for (int s = 0; s < nsteps; ++s) {
          const int kch = kchunks(s), halves = halves_of(s);
          const uint32_t idesc = umma_idesc_mn(256, halves == 2 ? 256 : 128);
          const int main_ch = kch < NSTAGE ? kch : NSTAGE;
          const uint32_t p0 = pos;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            { const long long t0 = AUX ? clock64() : 0; mbar_wait(A_READY(g), g == 0 ? ready_phase0 : ready_phase1, abort_flag); if (AUX) w_a += clock64() - t0; }
            if (g == 0) ready_phase0 ^= 1; else ready_phase1 ^= 1;
            tc_fence_after();
            const uint32_t act = base + SM_ACT + g * 4 * CHUNK_BYTES, pe = base + SM_PE + g * CHUNK_BYTES;
            const uint32_t d = tmem_base + g * 256;
            for (int c = 0; c < main_ch; ++c) {
              const uint32_t p = p0 + c;
              const int stage = p & (NSTAGE - 1);
              if (g == 0) { const long long t0 = AUX ? clock64() : 0; mbar_wait(W_FULL(stage), (p >> 2) & 1, abort_flag); if (AUX) w_w += clock64() - t0; tc_fence_after(); }
              const uint64_t ad = desc_of(MODE == MODE_FWD ? a_chunk_addr(s, c, act, pe) : act + c * CHUNK_BYTES);
              const uint64_t bd = desc_of(base + SM_W + stage * CHUNK_BYTES);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < KCH / 16; ++k)
                  umma_issue<2>(d, ad + 2 * k, bd + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
                if (g == 1) umma_commit_group<2>(W_EMPTY(stage));
                if (kch == main_ch && c == main_ch - 1) umma_commit_group<2>(ACC_FULL(g));
              }
            }
          }
          if (kch > main_ch) {
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
      if (AUX && a.wait_cycles && lane == 0) {
        atomicAdd(a.wait_cycles, (unsigned long long)w_a);
        atomicAdd(a.wait_cycles + 1, (unsigned long long)w_w);
        if (blockIdx.x == 0) atomicAdd(a.wait_cycles + 3, (unsigned long long)(clock64() - t_begin));
      }
    } else {
      uint32_t pos = 0;
      for (int64_t unit = group; unit < nunits; unit += ngroups)
        for (int s = 0; s < nsteps; ++s)
          for (int c = 0; c < kchunks(s); ++c, ++pos) {
            const int stage = pos & (NSTAGE - 1);
            mbar_wait(W_FULL(stage), (pos >> 2) & 1, abort_flag);
            if (elect_one()) mbar_arrive_remote(W_FULL(stage), 0);
          }
    }
  } else {
    // ================= input stage + epilogues (8 warps per slot: lane quarter x column half) =================
    const int e = warp - 2;
    const int g = e >> 3;
    const int hcol = (e >> 2) & 1;
    const int q = warp & 3;
    const int row = (q << 5) + lane;
    const uint32_t act = base + SM_ACT + g * 4 * CHUNK_BYTES;
    const uint32_t pe = base + SM_PE + g * CHUNK_BYTES;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q << 5) << 16) + g * 256;
    const uint32_t pair_bar = 1 + g * 4 + q;
    const uint32_t slot_bar = 9 + g;
    const MlpSide* __restrict__ sd = &c_side[a.cslot];
    const MlpSide* __restrict__ sg = a.side;
    uint32_t full_phase = 0;
    long long w_acc = 0;
    const int tslot = ((e & 7) << 5) + lane;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" :: "r"(pair_bar) : "memory"); };
    auto slot_sync = [&]() { asm volatile("bar.sync %0, 256;" :: "r"(slot_bar) : "memory"); };
    auto signal_a_ready = [&]() {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(A_READY(g)); else mbar_arrive_remote(A_READY(g), 0); }
    };
    float* scratch_pe = reinterpret_cast<float*>(smem_raw + SM_PE + g * CHUNK_BYTES);
    float* scratch_act = reinterpret_cast<float*>(smem_raw + SM_ACT + g * 4 * CHUNK_BYTES);
    float* bias_s = reinterpret_cast<float*>(smem_raw + SM_BIAS) + g * 256;
    auto bias_elem = [&](int step) -> float {
      if (MODE == MODE_BWD) return __ldg(&sg->w_alpha[tslot]);            // the row stays resident: w_alpha for the sigma term
      if (step < 8) return __ldg(&sg->bias[step][tslot]);
      if (step == 8) return __ldg(&sg->bias_feat[tslot]);
      return tslot < 128 ? __ldg(&sg->bias_views[tslot]) : 0.f;
    };
    bias_s[tslot] = bias_elem(0);
    slot_sync();

    int64_t prev_tile = -1;
    for (int64_t unit = group; unit < nunits; unit += ngroups) {
      const int64_t m = unit * ROWS_PER_UNIT + (int64_t)g * (TILE_M * CG) + rank * TILE_M + row;
      const int64_t tile = (unit * 2 + g) * CG + rank;          // = m / 128
      const bool live = m < a.M;
      const int64_t img_tile = (skip & 8) ? (int64_t)(blockIdx.x * 2 + g) : tile;     // bit 3: every CTA rewrites its own two tiles (stores stay in L2)
      char* act_tile = (MODE == MODE_FWD) ? a.act_img + img_tile * (int64_t)FWD_CHUNKS * CHUNK_BYTES : nullptr;
      char* dy_tile = (MODE == MODE_BWD) ? a.dy_img + img_tile * (int64_t)BWD_CHUNKS * CHUNK_BYTES : nullptr;
      uint32_t* mask_tile = a.mask + tile * (int64_t)MASK_WORDS_PER_TILE;
      // every bulk store of the previous unit must have finished READING shared memory before it is rewritten
      if (tslot == 0 && !(skip & 4)) bulk_wait_read();
      slot_sync();

      float vx = 0.f, vy = 0.f, vz = 0.f, gsig = 0.f;
      if (MODE == MODE_FWD) {
        // ---- input stage: encode the point (column-half-0 thread of each row) ----
        if (hcol == 0) {
          float px = 0.f, py = 0.f, pz = 0.f;
          if (live) {
            const int64_t r = m / a.S;
            const float* ray = a.rays + r * 11;
            const float z = __ldg(a.z_vals + m);
            px = __fadd_rn(__ldg(ray), __fmul_rn(__ldg(ray + 3), z));
            py = __fadd_rn(__ldg(ray + 1), __fmul_rn(__ldg(ray + 4), z));
            pz = __fadd_rn(__ldg(ray + 2), __fmul_rn(__ldg(ray + 5), z));
            vx = __ldg(ray + 8); vy = __ldg(ray + 9); vz = __ldg(ray + 10);
          }
          float f[64];
          encode3<L_PTS>(px, py, pz, f);
          store_row_chunk(pe, row, f);
        }
        fence_proxy_async();
        slot_sync();
        if (tslot == 0) { bulk_s2g(act_tile + (int64_t)IMG_PE * CHUNK_BYTES, pe, CHUNK_BYTES); bulk_commit(); }
      } else {
        // ---- input stage of the backward chain: dY of the view layer = (g_rgb . W_rgb) o relu'(hv), 64 columns per thread ----
        float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) gr = __ldg(reinterpret_cast<const float4*>(a.g_raw) + m);
        gsig = gr.w;
        const uint2 mw = make_uint2(mask_tile[(64 + hcol * 2) * 128 + row], mask_tile[(64 + hcol * 2 + 1) * 128 + row]);
        const uint32_t cb = act + hcol * CHUNK_BYTES;                    // chunk hcol holds view-layer columns hcol*64 ..
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = hcol * 64 + u * 8 + j;
            const float t = fmaf(gr.x, sd->w_rgb[0][c], fmaf(gr.y, sd->w_rgb[1][c], gr.z * sd->w_rgb[2][c]));
            const uint32_t word = (u < 4) ? mw.x : mw.y;
            v[j] = ((word >> (31 - ((u & 3) * 8 + j))) & 1u) ? t : 0.f;
          }
          const uint32_t addr = cb + row * 128 + (((u ^ (row & 7)) & 7) << 4);
          st_shared_v4(addr, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
        // (the two heads' weight gradients are side sums of g_raw itself in wgrad.cu: no dY chunk is written for them)
        fence_proxy_async();
        slot_sync();
        if (tslot == 0) {
          bulk_s2g(dy_tile, act, 2 * CHUNK_BYTES);
          bulk_commit();
        }
      }
      signal_a_ready();

      float sigma = 0.f;
      for (int s = 0; s < nsteps; ++s) {
        { const long long t0 = AUX ? clock64() : 0; mbar_wait(ACC_FULL(g), full_phase, abort_flag); if (AUX) w_acc += clock64() - t0; }
        full_phase ^= 1;
        tc_fence_after();
        const bool last = (s == nsteps - 1);
        // Image stores go out in two bulk groups per step: chunks {0, 2} as soon as the first half of the drain has
        // written them, chunks {1, 3} at the end.  Before overwriting chunks {0, 2} only the older group of the previous
        // step must be done reading shared memory; the younger one is waited for half a drain later (cc == 2).  The
        // stores are HBM-bound (64 KB per slot and step), so this doubles the time they have before they stall the drain.
        const bool wait_all = (MODE == MODE_BWD && s == 0) || (MODE == MODE_FWD && s == 9);   // previous group read chunks 0/1
        if (tslot == 0 && !(skip & 4)) { if (wait_all) bulk_wait_read(); else bulk_wait_read_but_last(); }
        if (AUX && MODE == MODE_BWD && ready && tslot == 0 && s >= 1) {
          // committed so far: the input-stage group and two groups per finished step; all but the last two have landed
          bulk_wait_all_but_two();
          publish_ready(ready + tile, s);
          if (s == 1 && prev_tile >= 0) publish_ready(ready + prev_tile, 1 + BWD_STEPS);
        }
        slot_sync();

        if (MODE == MODE_FWD && s == 9) {
          // ---- view layer: relu(acc + b) -> rgb head (fp32) and the bf16 image of hv (for the rgb / view weight gradients) ----
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int col0 = hcol * 64 + cc * 32;
            uint32_t v[32];
            tmem_ld32(tmem_row + col0, v);
            tmem_ld_wait();
            float h[32];
            uint32_t mword = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = reinterpret_cast<const float4*>(bias_s + col0)[j];
              const float4 w0 = reinterpret_cast<const float4*>(sd->w_rgb[0] + col0)[j];
              const float4 w1 = reinterpret_cast<const float4*>(sd->w_rgb[1] + col0)[j];
              const float4 w2 = reinterpret_cast<const float4*>(sd->w_rgb[2] + col0)[j];
              h[4 * j] = __uint_as_float(v[4 * j]) + b4.x;
              h[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
              h[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
              h[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
#pragma unroll
              for (int t = 0; t < 4; ++t) {                      // mask bit first (sign of the pre-activation), then relu
                mword = __funnelshift_l(__float_as_uint(h[4 * j + t]), mword, 1);
                h[4 * j + t] = fmaxf(h[4 * j + t], 0.f);
              }
              r0 = fmaf(h[4 * j], w0.x, r0); r0 = fmaf(h[4 * j + 1], w0.y, r0); r0 = fmaf(h[4 * j + 2], w0.z, r0); r0 = fmaf(h[4 * j + 3], w0.w, r0);
              r1 = fmaf(h[4 * j], w1.x, r1); r1 = fmaf(h[4 * j + 1], w1.y, r1); r1 = fmaf(h[4 * j + 2], w1.z, r1); r1 = fmaf(h[4 * j + 3], w1.w, r1);
              r2 = fmaf(h[4 * j], w2.x, r2); r2 = fmaf(h[4 * j + 1], w2.y, r2); r2 = fmaf(h[4 * j + 2], w2.z, r2); r2 = fmaf(h[4 * j + 3], w2.w, r2);
            }
            if (!(skip & 1)) mask_tile[(64 + hcol * 2 + cc) * 128 + row] = ~mword;
            const uint32_t cb = act + hcol * CHUNK_BYTES;                // view-layer columns hcol*64 .. -> chunk hcol
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t addr = cb + row * 128 + ((((cc * 4 + u) ^ (row & 7)) & 7) << 4);
              st_shared_v4(addr, pack_bf16(h[8 * u], h[8 * u + 1]), pack_bf16(h[8 * u + 2], h[8 * u + 3]),
                           pack_bf16(h[8 * u + 4], h[8 * u + 5]), pack_bf16(h[8 * u + 6], h[8 * u + 7]));
            }
          }
          // combine the two halves of the rgb logits through chunk 3 of the slot's activation buffer (unused here)
          float* scr = scratch_act + 3 * (CHUNK_BYTES / 4);
          if (hcol == 1) { scr[row * 4] = r0; scr[row * 4 + 1] = r1; scr[row * 4 + 2] = r2; }
          pair_sync();
          if (hcol == 0 && live) {
            const float4 o = make_float4(r0 + scr[row * 4] + sd->b_rgb[0], r1 + scr[row * 4 + 1] + sd->b_rgb[1],
                                         r2 + scr[row * 4 + 2] + sd->b_rgb[2], sigma);
            st_stream4(reinterpret_cast<float4*>(a.raw) + m, o);
          }
          fence_proxy_async();
          slot_sync();
          if (tslot == 0) { bulk_s2g(act_tile + (int64_t)IMG_HV * CHUNK_BYTES, act, 2 * CHUNK_BYTES); bulk_commit(); }
          bias_s[tslot] = bias_elem(0);
          slot_sync();
          continue;
        }

        const float next_bias = bias_elem(last ? 0 : s + 1);
        const bool relu = (MODE == MODE_FWD) ? (s < 8) : (s >= 1);       // bwd: step 0 outputs dY_feature (no activation)
        const int mask_layer = (MODE == MODE_FWD) ? s : 8 - s;           // bwd step b masks with relu'(h_{8-b})
        uint4 mw4 = make_uint4(0u, 0u, 0u, 0u);
        if (MODE == MODE_BWD && relu) {
          const uint32_t* mp = mask_tile + (mask_layer * 8 + hcol * 4) * 128 + row;
          mw4 = make_uint4(mp[0], mp[128], mp[256], mp[384]);
        }
        float sig_acc = 0.f, sig_b = 0.f, sig_c = 0.f, sig_d = 0.f;       // same four partial sums as the inference kernel (identical raw)
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          if (cc == 2) {
            fence_proxy_async();
            if (tslot == 0 && !(skip & 4)) bulk_wait_read();
            slot_sync();                               // chunks 0 and 2 are final; chunks 1 and 3 may be overwritten
            if (tslot == 0 && !(skip & 2)) {
              char* dst = (MODE == MODE_FWD) ? act_tile + (int64_t)(s * 4) * CHUNK_BYTES : dy_tile + (int64_t)dy_chunk0(s) * CHUNK_BYTES;
              bulk_s2g(dst, act, CHUNK_BYTES);
              bulk_s2g(dst + 2 * CHUNK_BYTES, act + 2 * CHUNK_BYTES, CHUNK_BYTES);
              bulk_commit();
            }
          }
          const int col0 = hcol * 128 + cc * 32;
          uint32_t v[32];
          tmem_ld32(tmem_row + col0, v);
          float4 b4[8];
          if (MODE == MODE_BWD && s == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) b4[j] = reinterpret_cast<const float4*>(bias_s + col0)[j];
          }
          tmem_ld_wait();
          float h[32];
          if (MODE == MODE_FWD) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = reinterpret_cast<const float4*>(bias_s + col0)[j];     // loaded at use: 32 fewer live registers
              h[4 * j] = __uint_as_float(v[4 * j]); h[4 * j + 1] = __uint_as_float(v[4 * j + 1]);
              h[4 * j + 2] = __uint_as_float(v[4 * j + 2]); h[4 * j + 3] = __uint_as_float(v[4 * j + 3]);
              add2(h[4 * j], h[4 * j + 1], b.x, b.y);
              add2(h[4 * j + 2], h[4 * j + 3], b.z, b.w);
            }
            if (s == 7) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 w4 = reinterpret_cast<const float4*>(sd->w_alpha + col0)[j];
                sig_acc = fmaf(fmaxf(h[4 * j], 0.f), w4.x, sig_acc); sig_b = fmaf(fmaxf(h[4 * j + 1], 0.f), w4.y, sig_b);
                sig_c = fmaf(fmaxf(h[4 * j + 2], 0.f), w4.z, sig_c); sig_d = fmaf(fmaxf(h[4 * j + 3], 0.f), w4.w, sig_d);
              }
            }
            if (relu) {
              // relu mask = complement of the gathered sign bits, one funnel shift per column: column j of the word is
              // bit 31 - j (an exactly zero pre-activation counts as active: measure-zero deviation from relu'(0) = 0)
              uint32_t mword = 0;                      // pair layout: odd columns first (bits 31..16), then even (15..0)
#pragma unroll
              for (int j = 0; j < 16; ++j) mword = __funnelshift_l(__float_as_uint(h[2 * j + 1]), mword, 1);
#pragma unroll
              for (int j = 0; j < 16; ++j) mword = __funnelshift_l(__float_as_uint(h[2 * j]), mword, 1);
              if (!(skip & 1)) mask_tile[(mask_layer * 8 + hcol * 4 + cc) * 128 + row] = ~mword;
            }
          } else {
            // backward: dX (+ g_sigma * w_alpha for the layer-7 activation), then relu'
            if (s == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                h[4 * j] = fmaf(gsig, b4[j].x, __uint_as_float(v[4 * j])); h[4 * j + 1] = fmaf(gsig, b4[j].y, __uint_as_float(v[4 * j + 1]));
                h[4 * j + 2] = fmaf(gsig, b4[j].z, __uint_as_float(v[4 * j + 2])); h[4 * j + 3] = fmaf(gsig, b4[j].w, __uint_as_float(v[4 * j + 3]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) h[j] = __uint_as_float(v[j]);
            }
          }
          // backward with relu': the mask is applied to the PACKED bf16 pairs (see the pair layout above)
          const bool bwd_mask = (MODE == MODE_BWD) && relu;
          const uint32_t bw = cc == 0 ? mw4.x : cc == 1 ? mw4.y : cc == 2 ? mw4.z : mw4.w;
          auto pack_masked = [&](float lo, float hi, int p) -> uint32_t {
            uint32_t v2 = pack_bf16(lo, hi), m;
            asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(bw << p));     // bytes 1 and 3, sign-replicated
            return v2 & m;
          };
          if (bwd_mask) {
            const uint32_t cb = act + (col0 >> 6) * CHUNK_BYTES;
            const int u0 = (col0 & 63) >> 3;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t addr = cb + row * 128 + ((((u0 + u) ^ (row & 7)) & 7) << 4);
              st_shared_v4(addr, pack_masked(h[8 * u], h[8 * u + 1], 4 * u), pack_masked(h[8 * u + 2], h[8 * u + 3], 4 * u + 1),
                           pack_masked(h[8 * u + 4], h[8 * u + 5], 4 * u + 2), pack_masked(h[8 * u + 6], h[8 * u + 7], 4 * u + 3));
            }
            continue;
          }
          const uint32_t cb = act + (col0 >> 6) * CHUNK_BYTES;
          const int u0 = (col0 & 63) >> 3;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t addr = cb + row * 128 + ((((u0 + u) ^ (row & 7)) & 7) << 4);
            if (MODE == MODE_FWD && relu)
              st_shared_v4(addr, pack_bf16_relu(h[8 * u], h[8 * u + 1]), pack_bf16_relu(h[8 * u + 2], h[8 * u + 3]),
                           pack_bf16_relu(h[8 * u + 4], h[8 * u + 5]), pack_bf16_relu(h[8 * u + 6], h[8 * u + 7]));
            else
              st_shared_v4(addr, pack_bf16(h[8 * u], h[8 * u + 1]), pack_bf16(h[8 * u + 2], h[8 * u + 3]),
                           pack_bf16(h[8 * u + 4], h[8 * u + 5]), pack_bf16(h[8 * u + 6], h[8 * u + 7]));
          }
        }
        sig_acc = (sig_acc + sig_b) + (sig_c + sig_d);
        if (MODE == MODE_FWD && s == 8 && hcol == 0) {
          float f[64];
          encode3<L_DIR>(vx, vy, vz, f);
          store_row_chunk(pe, row, f);
        }
        // this warp's share of the next A operand is complete: hand it to the MMA warp before the barriers that only
        // serve the sigma combine, the image stores and the bias row (they then run beside the next MMA)
        if (!last) signal_a_ready();
        if (MODE == MODE_FWD && s == 7) {
          if (hcol == 1) scratch_pe[row] = sig_acc;
          pair_sync();
          if (hcol == 0) sigma = sig_acc + scratch_pe[row] + sd->b_alpha;
          pair_sync();
        }
        fence_proxy_async();
        slot_sync();                                   // activation chunks (and the bias row) of this step are final
        if (tslot == 0) {
          if (!(skip & 2)) {
            char* dst = (MODE == MODE_FWD) ? act_tile + (int64_t)(s * 4) * CHUNK_BYTES : dy_tile + (int64_t)dy_chunk0(s) * CHUNK_BYTES;
            bulk_s2g(dst + CHUNK_BYTES, act + CHUNK_BYTES, CHUNK_BYTES);
            bulk_s2g(dst + 3 * CHUNK_BYTES, act + 3 * CHUNK_BYTES, CHUNK_BYTES);
            if (MODE == MODE_FWD && s == 8) bulk_s2g(act_tile + (int64_t)IMG_DIR * CHUNK_BYTES, pe, CHUNK_BYTES);
          }
          bulk_commit();
        }
        if (MODE == MODE_FWD) bias_s[tslot] = next_bias;
        slot_sync();
      }
      prev_tile = tile;
    }
    if (AUX && a.wait_cycles && e == 0 && lane == 0) atomicAdd(a.wait_cycles + 2, (unsigned long long)w_acc);
    if (tslot == 0) {
      bulk_wait_all();                                 // the images must be complete in HBM when the kernel ends
      if (AUX && MODE == MODE_BWD && ready && prev_tile >= 0) publish_ready(ready + prev_tile, 1 + BWD_STEPS);
    }
  }

  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512) : "memory");
}

}  // namespace tr
}  // namespace nfb
