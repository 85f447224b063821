// Weight-gradient GEMM on tcgen05:  dW[n_out, k_in] = sum_rows dY[row, n_out] * X[row, k_in]   (+ db = column sums of dY)
//
// Reference: the autograd of the addmm chain in Create_spatial_point_set/nerf_pytorch/run_nerf_helpers.py:100-123 as
// driven by loss.backward() at run_nerf.py:791 (the wgrad third of the 3 489 024 FLOP/sample of SURVEY.md §8d).
//
// Operands are the bf16 "tile images" the fused training kernels leave in HBM: per 128-row tile, 64-column chunks of
// 128 rows x 128 bytes with the 128-byte swizzle (exactly the shared-memory image of an A operand).  Read as a matrix
// with K = rows and MN = the 64 columns, such a chunk IS the canonical MN-major SWIZZLE_128B operand layout, so both
// dY^T (A, M = n_out) and X (B, N = k_in) are fed to tcgen05.mma straight from bulk copies, no transposition.
//   stage    = 64 rows of every chunk of one tile (8 KB slices, chunk pitch 8 KB inside the stage), 3-deep ring
//   MMA      = M 128 (two dY chunks) x N 64*nx x K 16 rows, fp32 accumulators for both M halves stay in TMEM (<= 512
//              columns) across all tiles of the CTA
//   bias     = 4 CUDA-core warps sum the dY slices of each stage out of shared memory
//   epilogue = per-CTA partial [256, N] (+ [256]) to a workspace; nfb_wgrad_reduce sums partials in fixed order.
// HBM-bound by design: 1 KB per row and layer against 2*256*256 FLOP.
#include "common.cuh"

namespace nfb {
namespace wg {

constexpr int ROWS_PER_STAGE = 64;
constexpr int SLICE_BYTES = ROWS_PER_STAGE * 128;       // 8 KB: 64 rows of one chunk
constexpr int MAX_CHUNKS = 8;                           // 4 dY + 4 X
constexpr int STAGE_BYTES = MAX_CHUNKS * SLICE_BYTES;   // 64 KB
constexpr int NSTAGE = 3;
constexpr int SM_BAR = NSTAGE * STAGE_BYTES;
constexpr int SMEM_BYTES = SM_BAR + 256;
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

struct Args {
  const char* dy;  int64_t dy_tile_pitch;  int ndy;     // bytes between tiles; number of 64-column chunks (2 or 4)
  const char* x;   int64_t x_tile_pitch;   int nx;      // 1, 2 or 4 chunks (N = 64*nx)
  int64_t ntiles;
  float* part_w;     // [grid][64*ndy][64*nx]
  float* part_b;     // [grid][64*ndy] or null
  int* flag;
};

__global__ void __launch_bounds__(THREADS, 1) wgrad_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar0 = base + SM_BAR;
  auto FULL_B = [&](int s) { return bar0 + 8 * s; };
  auto EMPTY_B = [&](int s) { return bar0 + 8 * (NSTAGE + s); };
  const uint32_t DONE_B = bar0 + 8 * (2 * NSTAGE);
  const uint32_t tmem_slot = bar0 + 8 * (2 * NSTAGE + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * (2 * NSTAGE + 1));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  volatile int* flag = a.flag;
  const int nchunks = a.ndy + a.nx;
  const int N = 64 * a.nx;
  const int mhalves = a.ndy / 2;

  if (threadIdx.x == 0) {
    if (base & 1023u) *flag = 1;
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(FULL_B(s), 1); mbar_init(EMPTY_B(s), 1 + (a.part_b ? a.ndy : 0)); }
    mbar_init(DONE_B, 1);
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

  // this CTA's tiles: blockIdx, blockIdx + grid, ...; each tile = 2 stages of 64 rows
  const int64_t my_tiles = a.ntiles > blockIdx.x ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t nstages_total = my_tiles * 2;

  if (warp == 0) {
    for (int64_t it = 0; it < nstages_total; ++it) {
      const int st = (int)(it % NSTAGE);
      const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
      mbar_wait(EMPTY_B(st), ph ^ 1, flag);
      if (elect_one()) {
        const int64_t tile = blockIdx.x + (it >> 1) * gridDim.x;
        const int half = (int)(it & 1);
        mbar_arrive_expect_tx(FULL_B(st), nchunks * SLICE_BYTES);
        const uint32_t dst = base + st * STAGE_BYTES;
        for (int c = 0; c < a.ndy; ++c)
          bulk_g2s(dst + c * SLICE_BYTES, a.dy + tile * a.dy_tile_pitch + (int64_t)c * 16384 + half * SLICE_BYTES, SLICE_BYTES, FULL_B(st));
        for (int c = 0; c < a.nx; ++c)
          bulk_g2s(dst + (a.ndy + c) * SLICE_BYTES, a.x + tile * a.x_tile_pitch + (int64_t)c * 16384 + half * SLICE_BYTES, SLICE_BYTES, FULL_B(st));
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_mn(N);
    for (int64_t it = 0; it < nstages_total; ++it) {
      const int st = (int)(it % NSTAGE);
      const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
      mbar_wait(FULL_B(st), ph, flag);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sbase = base + st * STAGE_BYTES;
      if (elect_one()) {
        for (int mh = 0; mh < mhalves; ++mh) {
#pragma unroll
          for (int k = 0; k < ROWS_PER_STAGE / 16; ++k) {
            const uint64_t ad = desc_mn(sbase + (2 * mh) * SLICE_BYTES + k * 2048, SLICE_BYTES);
            const uint64_t bd = desc_mn(sbase + a.ndy * SLICE_BYTES + k * 2048, SLICE_BYTES);
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                :: "r"(tmem_base + mh * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(EMPTY_B(st)) : "memory");
        if (it == nstages_total - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(DONE_B) : "memory");
      }
    }
  } else if (warp >= 6) {
    // ---- bias sums: warp (6 + c) owns dY chunk c; lane = (row group of 4) x (16-byte unit = 8 columns) ----
    const int c = warp - 6;
    if (a.part_b && c < a.ndy) {
      float acc8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc8[j] = 0.f;
      const int unit = lane & 7, rg = lane >> 3;
      for (int64_t it = 0; it < nstages_total; ++it) {
        const int st = (int)(it % NSTAGE);
        const uint32_t ph = (uint32_t)((it / NSTAGE) & 1);
        mbar_wait(FULL_B(st), ph, flag);
        const uint8_t* sl = smem + st * STAGE_BYTES + c * SLICE_BYTES;
        for (int r = rg; r < ROWS_PER_STAGE; r += 4) {
          const uint4 v = *reinterpret_cast<const uint4*>(sl + r * 128 + (((unit ^ (r & 7)) & 7) << 4));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc8[2 * j] += __uint_as_float(w[j] << 16);
            acc8[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(EMPTY_B(st));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc8[j] += __shfl_xor_sync(0xffffffffu, acc8[j], 8);
        acc8[j] += __shfl_xor_sync(0xffffffffu, acc8[j], 16);
      }
      if (rg == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a.part_b[(int64_t)blockIdx.x * 64 * a.ndy + c * 64 + unit * 8 + j] = acc8[j];
      }
    }
  } else {
    // ---- final epilogue: TMEM -> per-CTA partial ----
    if (nstages_total > 0) {
      mbar_wait(DONE_B, 0, flag);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int q = warp & 3;
    for (int mh = 0; mh < mhalves; ++mh) {
      const int row = mh * 128 + q * 32 + lane;
      float* out = a.part_w + ((int64_t)blockIdx.x * 64 * a.ndy + row) * N;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        if (nstages_total > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q << 5) << 16) + mh * 256 + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(out + c0)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
    if (a.part_b && nstages_total == 0 && warp == 2)
      for (int j = lane; j < 64 * a.ndy; j += 32) a.part_b[(int64_t)blockIdx.x * 64 * a.ndy + j] = 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512) : "memory");
}

// out[r * ld + col0 + c] (+)= sum_p part[p][r][c]  for r < rows, c < cols_valid (cols = pitch of the partial)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int rows, int cols, int cols_valid,
                                    float* __restrict__ out, int ld, int col0, int accumulate) {
  const int64_t n = (int64_t)rows * cols_valid;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols_valid), c = (int)(e % cols_valid);
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[((int64_t)p * rows + r) * cols + c];
    float* dst = out + (int64_t)r * ld + col0 + c;
    *dst = accumulate ? *dst + s : s;
  }
}

}  // namespace wg
}  // namespace nfb

extern "C" {

// Number of CTAs nfb_wgrad_bf16 launches (= leading dimension of the partial buffers).
int nfb_wgrad_parts(void) { return nfb::sm_count(); }

// dy / x: tile images (see header of this file); ndy in {2,4}, nx in {1,2,4}.
// part_w: [parts][64*ndy][64*nx] fp32, part_b: [parts][64*ndy] fp32 or NULL, status: device int (0 = ok).
int nfb_wgrad_bf16(const void* dy, int64_t dy_tile_pitch, int ndy, const void* x, int64_t x_tile_pitch, int nx,
                   int64_t ntiles, float* part_w, float* part_b, int* status, void* stream) {
  NFB_REQUIRE(dy && x && part_w && status, "wgrad_bf16: null pointer");
  NFB_REQUIRE((ndy == 2 || ndy == 4) && (nx == 1 || nx == 2 || nx == 4), "wgrad_bf16: ndy=%d nx=%d", ndy, nx);
  NFB_REQUIRE(ntiles >= 0 && dy_tile_pitch % 16 == 0 && x_tile_pitch % 16 == 0, "wgrad_bf16: bad pitch");
  static bool attr_set = false;
  if (!attr_set) {
    NFB_CUDA(cudaFuncSetAttribute(nfb::wg::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nfb::wg::SMEM_BYTES));
    attr_set = true;
  }
  nfb::wg::Args a{(const char*)dy, dy_tile_pitch, ndy, (const char*)x, x_tile_pitch, nx, ntiles, part_w, part_b, status};
  nfb::wg::wgrad_kernel<<<nfb::sm_count(), nfb::wg::THREADS, nfb::wg::SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return nfb::check_launch("wgrad_bf16");
}

int nfb_wgrad_reduce(const float* part, int nparts, int rows, int cols, int cols_valid, float* out, int ld, int col0,
                     int accumulate, void* stream) {
  NFB_REQUIRE(part && out && nparts > 0 && rows > 0 && cols > 0 && cols_valid > 0 && cols_valid <= cols, "wgrad_reduce: bad argument");
  const int64_t n = (int64_t)rows * cols_valid;
  nfb::wg::wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part, nparts, rows, cols, cols_valid,
                                                                                         out, ld, col0, accumulate);
  return nfb::check_launch("wgrad_reduce");
}

}  // extern "C"
