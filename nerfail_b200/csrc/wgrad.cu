// Weight-gradient GEMM on tcgen05:  dW[n_out, k_in] = sum_rows dY[row, n_out] * X[row, k_in]   (+ db = column sums of dY)
//
// Reference: the autograd of the addmm chain in Create_spatial_point_set/nerf_pytorch/run_nerf_helpers.py:100-123 as
// driven by loss.backward() at run_nerf.py:791 (the wgrad third of the 3 489 024 FLOP/sample of SURVEY.md §8d).
//
// Operands are the bf16 "tile images" the fused training kernels leave in HBM: per 128-row tile, 64-column chunks of
// 128 rows x 128 bytes with the 128-byte swizzle (exactly the shared-memory image of an A operand).  Read as a matrix
// with K = rows and MN = the 64 columns, such a chunk IS the canonical MN-major SWIZZLE_128B operand layout, so both
// dY^T (A, M = n_out) and X (B, N = k_in) are fed to tcgen05.mma straight from bulk copies, no transposition.
//   stage    = 64 rows of every chunk of one tile (8 KB slices, chunk pitch 8 KB inside the stage); the 192 KB ring
//              holds 3 (64 KB stages) to 8 (24 KB stages) of them, so narrow products keep as many bytes in flight
//              as wide ones (with a fixed 3-deep ring their CTAs were latency-bound and finished 1.5x late)
//   MMA      = M 128 (two dY chunks) x N 64*nx x K 16 rows, fp32 accumulators for both M halves stay in TMEM (<= 512
//              columns) across all tiles of the CTA
//   bias     = 4 CUDA-core warps sum the dY slices of each stage out of shared memory
//   epilogue = a CTA adds its partial [256, N] (+ [256]) into the gradient tensor with coalesced red.global.add.f32
//   grouped  = one launch computes every product of a network: a job table, (job, tile) pairs cut into equal-cost
//              contiguous ranges, one range per CTA (see job_segment)
// HBM-bound by design: 1 KB per row and layer against 2*256*256 FLOP.
#include "common.cuh"
#include <stdlib.h>
#include <vector>

namespace nfb {
namespace wg {

constexpr int ROWS_PER_STAGE = 64;
constexpr int SLICE_BYTES = ROWS_PER_STAGE * 128;       // 8 KB: 64 rows of one chunk
constexpr int MAX_CHUNKS = 8;                           // 4 dY + 4 X
constexpr int STAGE_BYTES = MAX_CHUNKS * SLICE_BYTES;   // 64 KB
constexpr int RING_BYTES = 3 * STAGE_BYTES;             // 192 KB ring, cut into as many stages as the job's stage size allows
constexpr int NSTAGE = 8;                               // upper bound (stage of 24 KB: a 2-chunk dY against a 1-chunk X)
constexpr int SM_BAR = RING_BYTES;
constexpr int SM_SCRATCH = SM_BAR + 256;                  // 4 epilogue warps x [32][33] floats (transpose for coalesced reductions)
constexpr int SM_SIDE = SM_SCRATCH + 4 * 32 * 33 * 4;     // [NSTAGE][64 rows][4] fp32: the stage's rows of g_raw (side products)
constexpr int SIDE_BYTES = ROWS_PER_STAGE * 16;
constexpr int SMEM_BYTES = SM_SIDE + NSTAGE * SIDE_BYTES;
static_assert(SM_SIDE % 16 == 0 && SMEM_BYTES <= 232448, "shared memory map");
constexpr int THREADS = 32 * 10;                        // warp 0 producer, 1 MMA/TMEM, 2-5 epilogue, 6-9 bias sums

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
// bounded spin inside one asm statement (see mlp_fused.cu); on timeout raises *flag and falls through
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* flag) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 n, f;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WG_DONE;\n\tmov.u32 n, 0;\n\t"
      "WG_SPIN:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WG_DONE;\n\t"
      "add.u32 n, n, 1;\n\tand.b32 f, n, 1023;\n\tsetp.ne.u32 q, f, 0;\n\t@q bra WG_SPIN;\n\t"
      "ld.volatile.global.u32 f, [%2];\n\tsetp.ne.u32 q, f, 0;\n\t@q bra WG_DONE;\n\t"
      "setp.lt.u32 q, n, 4194304;\n\t@q bra WG_SPIN;\n\tmov.u32 f, 1;\n\tst.volatile.global.u32 [%2], f;\n\t"
      "WG_DONE:\n\t}"
      :: "r"(bar), "r"(parity), "l"(flag) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// MN-major SWIZZLE_128B descriptor: 64-element MN groups `lbo` bytes apart, 8-row K groups 1024 bytes apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// bf16 x bf16 -> fp32, A and B both MN-major (bits 15, 16), M = 128, N = n
__device__ __forceinline__ uint32_t idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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

constexpr int MAX_JOBS = 16;
constexpr int MAX_CTAS = 160;

// One product dW = dY^T X of a grouped launch.  dY chunks [ndy_real, ndy) are read from a 16 KB block of zeros (pitch 0,
// L2 resident) so that a product with fewer than 128 output rows still fills an M = 128 MMA.
struct Job {
  const char* dy;  int64_t dy_pitch;    // chunk 0 of tile 0; bytes between tiles
  const char* x;   int64_t x_pitch;
  float* out_w;                         // accumulated into out_w[(row - row_begin) * ld + col0 + c], c < cols_valid
  float* out_b;                         // accumulated into out_b[row - row_begin] (column sums of dY), or null
  int ndy, ndy_real, nx;                // 64-column chunks: dY (2 or 4), of which read from HBM, X (1, 2 or 4)
  int ld, col0, cols_valid, row_begin, row_end;
  // side product (see WgradJob): side_nx extra chunks at side_x (loaded behind the X slices) or the job's own X
  int side_rows, side_gcol, side_cols, side_ld, side_nx;
  const char* side_x;
  float* out_side_w;
  float* out_side_b;
  int cost;                             // relative time of one tile of this job (work split): 4 per chunk moved + side sums
};

struct Args {
  Job job[MAX_JOBS];
  int njobs;
  int64_t ntiles;
  const char* zero;                     // 16 KB of zeros (only read when some job has ndy_real < ndy)
  int* flag;
  const float* g_raw;                   // [M,4] fp32 upstream gradient (side products), M rows are valid
  int64_t M;
  // consumer mode (ready != null): the dY image is being written by a concurrently running mlp_train_kernel<BWD>, which
  // publishes per tile how many of its store groups have landed.  Every CTA then owns ONE job and the tiles
  // cta_first, cta_first + cta_stride, ... in increasing order (the order the producer finishes them in), and waits for
  // ready[tile] >= job_need[job] before loading a tile, so dY is consumed out of L2 while it is still resident.
  const int* ready;
  short cta_job[MAX_CTAS], cta_first[MAX_CTAS], cta_stride[MAX_CTAS];
  signed char job_need[MAX_JOBS];
  unsigned long long* dbg;              // profiling only (NERFAIL_B200_WGRAD_DBG=1): per CTA globaltimer at start / end
};

// Work split: the (job, tile) pairs, jobs in table order, are cut into gridDim.x contiguous ranges of equal HBM cost
// (cost of a tile of job j = ndy + nx chunks moved into shared memory, zero padding included).  A CTA therefore works on one to three consecutive jobs, keeps a
// job's whole dW in TMEM while it walks that job's tiles and flushes it once per job.
struct Segment { int64_t lo, hi, stride; };
__device__ __forceinline__ int64_t seg_tiles(const Segment& s) { return s.hi > s.lo ? (s.hi - s.lo + s.stride - 1) / s.stride : 0; }
__device__ __forceinline__ int ring_stages(int stage_bytes) {
  const int n = RING_BYTES / stage_bytes;
  return n < NSTAGE ? n : NSTAGE;
}
__device__ __forceinline__ Segment job_segment(const Args& a, int j, int64_t cost_before, int64_t cost_total) {
  if (a.ready) {
    if (a.cta_job[blockIdx.x] != j) return Segment{0, 0, 1};
    return Segment{a.cta_first[blockIdx.x], a.ntiles, a.cta_stride[blockIdx.x]};
  }
  const int64_t W = a.ntiles * cost_total;
  const int64_t w0 = W * blockIdx.x / gridDim.x, w1 = W * (blockIdx.x + 1) / gridDim.x;
  const int64_t base = cost_before * a.ntiles;
  const int64_t c = a.job[j].cost;
  auto cut = [&](int64_t w) -> int64_t {
    if (w <= base) return 0;
    const int64_t t = (w - base + c - 1) / c;
    return t < a.ntiles ? t : a.ntiles;
  };
  return Segment{cut(w0), cut(w1), 1};
}

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" :: "l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void red_add_v4_f32(float* addr, float a0, float a1, float a2, float a3) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(addr), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar0 = base + SM_BAR;
  auto FULL_B = [&](int s) { return bar0 + 8 * s; };
  auto EMPTY_B = [&](int s) { return bar0 + 8 * (NSTAGE + s); };
  const uint32_t DONE_B = bar0 + 8 * (2 * NSTAGE);
  const uint32_t FREE_B = bar0 + 8 * (2 * NSTAGE + 1);
  const uint32_t tmem_slot = bar0 + 8 * (2 * NSTAGE + 2);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * (2 * NSTAGE + 2));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  volatile int* flag = a.flag;

  if (threadIdx.x == 0) {
    if (base & 1023u) *flag = 1;
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(FULL_B(s), 1); mbar_init(EMPTY_B(s), 1 + 4); }
    mbar_init(DONE_B, 1);
    mbar_init(FREE_B, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  int64_t cost_total = 0;
  for (int j = 0; j < a.njobs; ++j) cost_total += a.job[j].cost;
  if (a.dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[2 * blockIdx.x] = t;
  }

  if (warp == 0) {
    // ================= producer: 64-row slices of every chunk of a tile, one stage per half tile =================
    int64_t cost_before = 0;
    uint32_t pmask = 0;                                  // bit s = parity of the next use of stage index s
    bool first = true;
    for (int j = 0; j < a.njobs; ++j) {
      const Job& jb = a.job[j];
      const Segment sg = job_segment(a, j, cost_before, cost_total);
      cost_before += jb.cost;
      if (sg.hi <= sg.lo) continue;
      const int stage_bytes = (jb.ndy + jb.nx + jb.side_nx) * SLICE_BYTES;
      const int nst = ring_stages(stage_bytes);
      if (!first)                                        // the ring is re-cut: every stage of the previous job must be free
        for (int s2 = 0; s2 < NSTAGE; ++s2) mbar_wait(EMPTY_B(s2), ((pmask >> s2) & 1u) ^ 1u, flag);
      first = false;
      int st = 0;
      for (int64_t tile = sg.lo; tile < sg.hi; tile += sg.stride) {
        if (a.ready) {                                   // wait until the producer kernel has published this tile's dY
          if (lane == 0) {
            const int need = a.job_need[j];
            int v, spins = 0;
            while (true) {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a.ready + tile) : "memory");
              if (v >= need) break;
              __nanosleep(200);
              if ((++spins & 1023) == 0 && (*flag != 0 || spins > (1 << 22))) { *flag = 1; break; }
            }
          }
          __syncwarp();
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        for (int half = 0; half < 2; ++half) {
          const uint32_t ph = (pmask >> st) & 1u;
          mbar_wait(EMPTY_B(st), ph ^ 1, flag);
          const uint32_t dst = base + st * stage_bytes;
          const uint32_t full = FULL_B(st);
          const int st_cur = st;
          pmask ^= 1u << st;
          st = (st + 1 == nst) ? 0 : st + 1;
          // the stage's 64 rows of g_raw (16 B each) for the side product: only the rows below M exist
          int64_t side_valid = 0;
          if (jb.side_rows > 0) {
            side_valid = a.M - (tile * 128 + half * ROWS_PER_STAGE);
            side_valid = side_valid < 0 ? 0 : (side_valid > ROWS_PER_STAGE ? ROWS_PER_STAGE : side_valid);
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(full, stage_bytes + (uint32_t)side_valid * 16);
            if (side_valid > 0)
              bulk_g2s(base + SM_SIDE + (st_cur) * SIDE_BYTES, a.g_raw + (tile * 128 + half * ROWS_PER_STAGE) * 4, (uint32_t)side_valid * 16, full);
            for (int c = 0; c < jb.side_nx; ++c)
              bulk_g2s(dst + (jb.ndy + jb.nx + c) * SLICE_BYTES, jb.side_x + tile * jb.x_pitch + (int64_t)c * 16384 + half * SLICE_BYTES,
                       SLICE_BYTES, full);
            for (int c = 0; c < jb.ndy; ++c) {
              const char* src = c < jb.ndy_real ? jb.dy + tile * jb.dy_pitch + (int64_t)c * 16384 + half * SLICE_BYTES : a.zero;
              bulk_g2s(dst + c * SLICE_BYTES, src, SLICE_BYTES, full);
            }
            for (int c = 0; c < jb.nx; ++c)
              bulk_g2s(dst + (jb.ndy + c) * SLICE_BYTES, jb.x + tile * jb.x_pitch + (int64_t)c * 16384 + half * SLICE_BYTES,
                       SLICE_BYTES, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: dW (both 128-row halves) accumulates in TMEM over the job's tiles =================
    int64_t cost_before = 0;
    uint32_t seg = 0, pmask = 0;
    for (int j = 0; j < a.njobs; ++j) {
      const Job& jb = a.job[j];
      const Segment sg = job_segment(a, j, cost_before, cost_total);
      cost_before += jb.cost;
      if (sg.hi <= sg.lo) continue;
      const int stage_bytes = (jb.ndy + jb.nx + jb.side_nx) * SLICE_BYTES;
      const int nst_ring = ring_stages(stage_bytes);
      int st = 0;
      const uint32_t idesc = idesc_mn(64 * jb.nx);
      const int mhalves = jb.ndy / 2;
      if (seg > 0) {                                   // the previous job's accumulators must have been drained
        mbar_wait(FREE_B, (seg - 1) & 1, flag);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const int64_t nst = seg_tiles(sg) * 2;
      for (int64_t i = 0; i < nst; ++i) {
        mbar_wait(FULL_B(st), (pmask >> st) & 1u, flag);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sbase = base + st * stage_bytes;
        const uint32_t empty = EMPTY_B(st);
        pmask ^= 1u << st;
        st = (st + 1 == nst_ring) ? 0 : st + 1;
        if (elect_one()) {
          for (int mh = 0; mh < mhalves; ++mh) {
#pragma unroll
            for (int k = 0; k < ROWS_PER_STAGE / 16; ++k) {
              const uint64_t ad = desc_mn(sbase + (2 * mh) * SLICE_BYTES + k * 2048, SLICE_BYTES);
              const uint64_t bd = desc_mn(sbase + jb.ndy * SLICE_BYTES + k * 2048, SLICE_BYTES);
              const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                  :: "r"(tmem_base + mh * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(empty) : "memory");
          if (i == nst - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(DONE_B) : "memory");
        }
      }
      ++seg;
    }
  } else if (warp >= 6) {
    // ================= bias sums: warp (6 + c) owns dY chunk c; lane = (row group of 4) x (16-byte unit = 8 columns) ====
    const int c = warp - 6;
    const int unit = lane & 7, rg = lane >> 3;
    int64_t cost_before = 0;
    uint32_t pmask = 0;
    for (int j = 0; j < a.njobs; ++j) {
      const Job& jb = a.job[j];
      const Segment sg = job_segment(a, j, cost_before, cost_total);
      cost_before += jb.cost;
      if (sg.hi <= sg.lo) continue;
      const int stage_bytes = (jb.ndy + jb.nx + jb.side_nx) * SLICE_BYTES;
      const int nst_ring = ring_stages(stage_bytes);
      int st = 0;
      const bool mine = jb.out_b && c < jb.ndy_real;
      float acc8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc8[q] = 0.f;
      // side product: a thread owns one 16-byte unit (8 columns) of S and every RG-th row of the stage; the 128 threads
      // cover side_cols / 8 units x RG row groups (256 columns: 32 x 4, 128 columns: 16 x 8)
      const bool side = jb.side_rows > 0;
      const bool three = jb.side_rows == 3;
      const int side_slice0 = jb.side_nx ? jb.ndy + jb.nx : jb.ndy;       // first slice of S inside the stage
      const int nunits = side ? jb.side_cols >> 3 : 32;
      const int uidx = lane & (nunits - 1);
      const int RG = 4 * (32 / nunits);
      const int my_rg = c * (32 / nunits) + lane / nunits;
      float sa[8], sy[8], sz[8], sb0 = 0.f, sb1 = 0.f, sb2 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) { sa[q] = 0.f; sy[q] = 0.f; sz[q] = 0.f; }
      const int64_t nst = seg_tiles(sg) * 2;
      for (int64_t i = 0; i < nst; ++i) {
        mbar_wait(FULL_B(st), (pmask >> st) & 1u, flag);
        const uint32_t empty = EMPTY_B(st);
        if (mine) {
          const uint8_t* sl = smem + st * stage_bytes + c * SLICE_BYTES;
          for (int r = rg; r < ROWS_PER_STAGE; r += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(sl + r * 128 + (((unit ^ (r & 7)) & 7) << 4));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc8[2 * q] += __uint_as_float(w[q] << 16);
              acc8[2 * q + 1] += __uint_as_float(w[q] & 0xFFFF0000u);
            }
          }
        }
        if (side) {
          const int64_t tile = sg.lo + (i >> 1) * sg.stride;
          int64_t valid = a.M - (tile * 128 + (i & 1) * ROWS_PER_STAGE);
          valid = valid < 0 ? 0 : (valid > ROWS_PER_STAGE ? ROWS_PER_STAGE : valid);
          const float4* gs = reinterpret_cast<const float4*>(smem + SM_SIDE + st * SIDE_BYTES);
          const uint8_t* sx = smem + st * stage_bytes + (side_slice0 + (uidx >> 3)) * SLICE_BYTES;
          const int unit_s = uidx & 7;
          const int gc = jb.side_gcol;
          const int nvalid = (int)valid;
#pragma unroll 2
          for (int r = my_rg; r < nvalid; r += RG) {
            const float4 g = gs[r];                                       // same address across the row group: broadcast
            const uint4 xv = *reinterpret_cast<const uint4*>(sx + r * 128 + (((unit_s ^ (r & 7)) & 7) << 4));
            // rows = 1: the column side_gcol of g (sigma: 3); rows = 3: columns 0..2 (the rgb logits)
            const float g0 = three ? g.x : (gc == 3 ? g.w : gc == 2 ? g.z : gc == 1 ? g.y : g.x);
            const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w};
            sb0 += g0;
            if (three) { sb1 += g.y; sb2 += g.z; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float xl = __uint_as_float(w[q] << 16), xh = __uint_as_float(w[q] & 0xFFFF0000u);
              sa[2 * q] = fmaf(g0, xl, sa[2 * q]);
              sa[2 * q + 1] = fmaf(g0, xh, sa[2 * q + 1]);
              if (three) {
                sy[2 * q] = fmaf(g.y, xl, sy[2 * q]); sy[2 * q + 1] = fmaf(g.y, xh, sy[2 * q + 1]);
                sz[2 * q] = fmaf(g.z, xl, sz[2 * q]); sz[2 * q + 1] = fmaf(g.z, xh, sz[2 * q + 1]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty);
        pmask ^= 1u << st;
        st = (st + 1 == nst_ring) ? 0 : st + 1;
      }
      if (mine) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          acc8[q] += __shfl_xor_sync(0xffffffffu, acc8[q], 8);
          acc8[q] += __shfl_xor_sync(0xffffffffu, acc8[q], 16);
        }
        if (rg == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int r = c * 64 + unit * 8 + q;
            if (r >= jb.row_begin && r < jb.row_end) red_add_f32(jb.out_b + (r - jb.row_begin), acc8[q]);
          }
        }
      }
      if (side) {                                   // every row group adds its partial sums (once per job and CTA)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = uidx * 8 + q;
          red_add_f32(jb.out_side_w + col, sa[q]);
          if (three) { red_add_f32(jb.out_side_w + jb.side_ld + col, sy[q]); red_add_f32(jb.out_side_w + 2 * jb.side_ld + col, sz[q]); }
        }
        if (uidx == 0 && jb.out_side_b) {
          red_add_f32(jb.out_side_b, sb0);
          if (three) { red_add_f32(jb.out_side_b + 1, sb1); red_add_f32(jb.out_side_b + 2, sb2); }
        }
      }
    }
  } else {
    // ================= epilogue: TMEM -> shared-memory transpose -> coalesced L2 reductions into the gradient =========
    const int q = warp & 3;
    float* scr = reinterpret_cast<float*>(smem + SM_SCRATCH) + (warp - 2) * (32 * 33);
    int64_t cost_before = 0;
    uint32_t seg = 0;
    for (int j = 0; j < a.njobs; ++j) {
      const Job& jb = a.job[j];
      const Segment sg = job_segment(a, j, cost_before, cost_total);
      cost_before += jb.cost;
      if (sg.hi <= sg.lo) continue;
      mbar_wait(DONE_B, seg & 1, flag);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int N = 64 * jb.nx, mhalves = jb.ndy / 2;
      // Flush.  A lane holds 32 consecutive columns of ONE output row (TMEM lane = row): where the destination rows are
      // 16-byte aligned (ld, col0, cols_valid multiples of 4: the seven 256 x 256 products and feature_linear) they go out
      // as eight 128-bit L2 reductions per lane, straight from the registers; the odd-pitch products (63-, 319- and
      // 283-column weights) keep the shared-memory transpose + scalar reductions.  The flush is a fixed cost per (CTA,
      // product) that the pipeline cannot hide (the accumulators fill TMEM), which is what a small batch pays for.
      const bool vec = ((jb.ld | jb.col0 | jb.cols_valid) & 3) == 0 && (reinterpret_cast<uintptr_t>(jb.out_w) & 15) == 0;
      for (int mh = 0; mh < mhalves; ++mh) {
        const int row0 = mh * 128 + q * 32;                      // this warp's 32 output rows (TMEM lanes q*32 ..)
        if (row0 >= jb.row_end || row0 + 32 <= jb.row_begin) continue;
        for (int c0 = 0; c0 < N && c0 < jb.cols_valid; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q << 5) << 16) + mh * 256 + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (vec) {
            const int row = row0 + lane;
            if (row >= jb.row_begin && row < jb.row_end) {
              float* dst = jb.out_w + (int64_t)(row - jb.row_begin) * jb.ld + jb.col0 + c0;
#pragma unroll
              for (int t = 0; t < 32; t += 4)
                if (c0 + t < jb.cols_valid)
                  red_add_v4_f32(dst + t, __uint_as_float(v[t]), __uint_as_float(v[t + 1]), __uint_as_float(v[t + 2]), __uint_as_float(v[t + 3]));
            }
            continue;
          }
#pragma unroll
          for (int t = 0; t < 32; ++t) scr[lane * 33 + t] = __uint_as_float(v[t]);
          __syncwarp();
          const bool col_ok = c0 + lane < jb.cols_valid;
          for (int r = 0; r < 32; ++r) {
            const int row = row0 + r;
            if (row >= jb.row_begin && row < jb.row_end && col_ok)
              red_add_f32(jb.out_w + (int64_t)(row - jb.row_begin) * jb.ld + jb.col0 + c0 + lane, scr[r * 33 + lane]);
          }
          __syncwarp();
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(FREE_B);
      ++seg;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (a.dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[2 * blockIdx.x + 1] = t;
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512) : "memory");
}

}  // namespace wg

// Launches the grouped weight-gradient kernel (declared in common.cuh; the NeRF job table is built in mlp_fused.cu).
int launch_wgrad_grouped(const WgradJob* jobs, int njobs, int64_t ntiles, const void* zero16k, int* status, void* stream,
                         const char* what, const int* ready, const signed char* job_need, int consumer_ctas,
                         const float* g_raw, int64_t M) {
  NFB_REQUIRE(jobs && njobs > 0 && njobs <= wg::MAX_JOBS && status, "%s: bad job table", what);
  if (ntiles <= 0) return NFB_OK;
  wg::Args a{};
  int64_t cost = 0;
  for (int j = 0; j < njobs; ++j) {
    const WgradJob& s = jobs[j];
    NFB_REQUIRE(s.dy && s.x && s.out_w, "%s: job %d: null pointer", what, j);
    NFB_REQUIRE((s.ndy == 2 || s.ndy == 4) && s.ndy_real >= 1 && s.ndy_real <= s.ndy && (s.nx == 1 || s.nx == 2 || s.nx == 4),
                "%s: job %d: ndy=%d (%d real) nx=%d", what, j, s.ndy, s.ndy_real, s.nx);
    NFB_REQUIRE(s.ndy_real == s.ndy || zero16k, "%s: job %d needs the zero block", what, j);
    NFB_REQUIRE(s.dy_pitch % 16 == 0 && s.x_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(s.dy) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(s.x) & 15) == 0, "%s: job %d: images must be 16-byte aligned", what, j);
    NFB_REQUIRE(s.cols_valid > 0 && s.cols_valid <= 64 * s.nx && s.row_begin >= 0 && s.row_end <= 64 * s.ndy &&
                s.row_begin < s.row_end && s.ld >= s.col0 + s.cols_valid, "%s: job %d: bad output window", what, j);
    if (s.side_rows > 0) {
      NFB_REQUIRE(g_raw && M > 0 && (reinterpret_cast<uintptr_t>(g_raw) & 15) == 0, "%s: job %d: side product needs g_raw [M,4]", what, j);
      NFB_REQUIRE((s.side_rows == 1 || (s.side_rows == 3 && s.side_gcol == 0)) && s.side_gcol >= 0 && s.side_gcol < 4 &&
                  (s.side_cols == 256 || s.side_cols == 128) && s.side_ld >= s.side_cols && s.out_side_w,
                  "%s: job %d: bad side product", what, j);
      NFB_REQUIRE(s.side_x ? (s.side_nx >= 1 && s.side_nx <= 2 && s.side_cols <= 64 * s.side_nx &&
                              (reinterpret_cast<uintptr_t>(s.side_x) & 15) == 0)
                           : (s.side_nx == 0 && s.side_cols <= 64 * s.nx), "%s: job %d: bad side operand", what, j);
    } else {
      NFB_REQUIRE(s.side_nx == 0 && !s.side_x, "%s: job %d: side operand without side rows", what, j);
    }
    a.job[j] = wg::Job{(const char*)s.dy, s.dy_pitch, (const char*)s.x, s.x_pitch, s.out_w, s.out_b,
                       s.ndy, s.ndy_real, s.nx, s.ld, s.col0, s.cols_valid, s.row_begin, s.row_end,
                       s.side_rows, s.side_gcol, s.side_cols, s.side_ld, s.side_nx, (const char*)s.side_x, s.out_side_w, s.out_side_b, 0};
    // measured (B200, 786 k samples): a stage with the 256-column side sum takes 1.25x, with the 3 x 128 one 1.13x
    a.job[j].cost = 4 * (s.ndy + s.nx + s.side_nx) + (s.side_rows == 1 ? 8 : s.side_rows == 3 ? 3 : 0);
    cost += a.job[j].cost;
  }
  a.njobs = njobs; a.ntiles = ntiles; a.zero = (const char*)zero16k; a.flag = status;
  a.g_raw = g_raw; a.M = M;
  a.ready = ready;
  NFB_CUDA(cudaFuncSetAttribute(wg::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES));
  int64_t grid = sm_count();
  if ((int64_t)njobs * ntiles < grid) grid = (int64_t)njobs * ntiles;
  if (ready) {
    // consumer mode: CTAs per job proportional to the job's cost (largest-remainder rounding, at least one each)
    NFB_REQUIRE(job_need && consumer_ctas >= njobs && consumer_ctas <= wg::MAX_CTAS && consumer_ctas % 2 == 0 && ntiles < 32768,
                "%s: bad consumer split", what);
    grid = consumer_ctas;
    int n[wg::MAX_JOBS], given = 0;
    double frac[wg::MAX_JOBS];
    for (int j = 0; j < njobs; ++j) {
      const double share = (double)grid * a.job[j].cost / (double)cost;
      n[j] = (int)share < 1 ? 1 : (int)share;
      frac[j] = share - n[j];
      given += n[j];
      a.job_need[j] = job_need[j];
    }
    while (given < grid) { int b = 0; for (int j = 1; j < njobs; ++j) if (frac[j] > frac[b]) b = j; ++n[b]; frac[b] -= 1.0; ++given; }
    while (given > grid) { int b = -1; for (int j = 0; j < njobs; ++j) if (n[j] > 1 && (b < 0 || frac[j] < frac[b])) b = j; --n[b]; frac[b] += 1.0; --given; }
    int c = 0;
    for (int j = 0; j < njobs; ++j)
      for (int k = 0; k < n[j]; ++k, ++c) { a.cta_job[c] = (short)j; a.cta_first[c] = (short)k; a.cta_stride[c] = (short)n[j]; }
  }
  static const bool dbg = []() { const char* e = getenv("NERFAIL_B200_WGRAD_DBG"); return e && e[0] == '1'; }();
  if (dbg) NFB_CUDA(cudaMalloc(&a.dbg, sizeof(unsigned long long) * 2 * grid));
  if (ready) {
    // CTA pairs (clusters of 2, no cluster communication) so that the consumers take whole TPCs: whichever of the two
    // concurrent kernels is placed first, the producer's 2-CTA clusters still find free SM pairs
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(wg::THREADS);
    cfg.dynamicSmemBytes = wg::SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, wg::wgrad_kernel, a);
    if (e != cudaSuccess) return fail(NFB_E_CUDA, "%s: cluster launch: %s", what, cudaGetErrorString(e));
  } else {
    wg::wgrad_kernel<<<(unsigned)grid, wg::THREADS, wg::SMEM_BYTES, (cudaStream_t)stream>>>(a);
  }
  if (dbg) {     // per-CTA duration against its first job: shows whether the equal-cost split is equal-time
    std::vector<unsigned long long> t(2 * grid);
    NFB_CUDA(cudaMemcpy(t.data(), a.dbg, sizeof(unsigned long long) * 2 * grid, cudaMemcpyDeviceToHost));
    cudaFree(a.dbg);
    unsigned long long t0 = ~0ull;
    for (int64_t b = 0; b < grid; ++b) t0 = t[2 * b] < t0 ? t[2 * b] : t0;
    const int64_t W = ntiles * cost;
    for (int64_t b = 0; b < grid; ++b) {
      const int64_t w0 = W * b / grid;
      int64_t cb = 0; int j0 = 0;
      for (int j = 0; j < njobs; ++j) { const int64_t c = a.job[j].cost; if (w0 < (cb + c) * ntiles) { j0 = j; break; } cb += c; }
      fprintf(stderr, "wgrad cta %3lld first job %2d (ndy %d/%d nx %d)  start %8.1f us  end %8.1f us\n", (long long)b, j0, jobs[j0].ndy_real,
              jobs[j0].ndy, jobs[j0].nx, (t[2 * b] - t0) / 1e3, (t[2 * b + 1] - t0) / 1e3);
    }
  }
  return check_launch(what);
}

}  // namespace nfb

extern "C" {

// dy / x: tile images (see header of this file); ndy in {2,4}, nx in {1,2,4}.
// Accumulates (L2 float reductions) dW rows [row_begin, row_end) x columns [0, cols_valid) into
// out_w[(row - row_begin) * ld + col0 + c] and the column sums of dY into out_b[row - row_begin] (or NULL): the caller
// zeroes the gradient buffers once per step.  status: device int raised if a pipeline barrier timed out.
int nfb_wgrad_bf16(const void* dy, int64_t dy_tile_pitch, int ndy, const void* x, int64_t x_tile_pitch, int nx,
                   int64_t ntiles, float* out_w, int ld, int col0, int cols_valid, int row_begin, int row_end,
                   float* out_b, int* status, void* stream) {
  nfb::WgradJob j{};
  j.dy = dy; j.dy_pitch = dy_tile_pitch; j.x = x; j.x_pitch = x_tile_pitch; j.out_w = out_w; j.out_b = out_b;
  j.ndy = ndy; j.ndy_real = ndy; j.nx = nx; j.ld = ld; j.col0 = col0; j.cols_valid = cols_valid; j.row_begin = row_begin; j.row_end = row_end;
  return nfb::launch_wgrad_grouped(&j, 1, ntiles, nullptr, status, stream, "wgrad_bf16");
}

}  // extern "C"
