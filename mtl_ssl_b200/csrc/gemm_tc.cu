// tcgen05 / TMEM / TMA implicit-GEMM convolution engine for sm_100a.
//
// One persistent, warp-specialised kernel serves the three GEMMs of a
// convolution layer on NHWC bf16 activations (fp32 accumulate in TMEM):
//   FPROP  y[NPQ,K]   = im2col(x)[NPQ,RSC] * w[K,RSC]^T       (A K-major, B K-major)
//   DGRAD  dx[NHW,C]  = im2col^T(dy)[NHW,RSK] * w[K,(RS)C]    (A K-major, B MN-major)
//   WGRAD  dw[K,RSC] += dy[NPQ,K]^T * im2col(x)[NPQ,RSC]      (A MN-major, B MN-major, split-K)
// The reference delegates these to cuDNN through slim.conv2d
// (/root/reference/slim/nets/resnet_v1.py:107-126, resnet_utils.py:77-122); there is no
// reference kernel to mirror, so the design below is new.
//
// Data path: weight / plain-matrix operands arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) issued by
// one producer thread from a division-free loop; stride-1 R x S filters use TMA im2col loads, strided / odd
// geometries are staged into the same swizzled layout by four gather warps (16-byte cp.async + zero fill).
// One thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM accumulator.  Eight epilogue
// warps (two groups taking alternate 32-column chunks; four in the gather variants) drain it with
// tcgen05.ld and apply bias / residual / ReLU / mask; bf16 outputs, residuals and masks move as 32 x 32
// tiles through swizzled shared memory by TMA (stores, and a ring of prefetched loads), WGRAD's fp32 tile
// leaves by TMA reduce-add.  Shared memory (operand stages vs. epilogue ring) is carved at launch time.
// Optional schedules, off by default (see DESIGN.md 2 / profiles/r1_final.md): fp32-workspace split-K for
// FPROP / DGRAD with a last-arriver epilogue; CTA pairs (clusters of 2) multicasting the weight tile.
#include "common.cuh"
#include <cuda.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

namespace tc {

constexpr int BM = 128;            // rows per tile (UMMA M)
constexpr int BK = 64;             // bf16 per K chunk (= one 128 B swizzle row)
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int NUM_GATHER_THREADS = 128;

enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct Params {
  int M, N;                  // GEMM output rows / cols
  int k_iters;               // total K iterations (taps*cpt, or pixel blocks for WGRAD)
  int taps, cpt;             // filter taps and 64-wide channel chunks per tap
  int tiles_m, tiles_n, splits;
  // gather geometry (operand staged by the gather warps)
  const bf16* gsrc;          // NHWC tensor being gathered
  int gH, gW, gC;            // its spatial dims and channel count
  int gld;                   // its pixel pitch in elements (>= gC: the tensor may be a channel slice)
  int oH, oW;                // row-space spatial dims: row -> (n, oh, ow)
  int rows;                  // number of valid rows in row space
  int S, stride, pad_h, pad_w, dil, transposed;
  int ntot;                  // DGRAD: Cin (columns per tap of w); WGRAD: Cin
  unsigned long long div_ohw, div_ow;   // multiply-shift division constants for oH*oW and oW
  int im2col;                // gathered operand comes by TMA im2col (stride-1 filters), no gather warps
  int im_low_h, im_low_w;    // base-pixel offset of filter offset (0,0)
  int R;                     // filter rows (mirrored offsets in DGRAD)
  // epilogue
  void* out; long long ldo; int out_fp32;
  const float* bias;         // per output column (FPROP/DGRAD) or nullptr
  float bias_scale;          // bias is added as bias[n] * bias_scale (scaled residual branches)
  const float* rowscale;     // WGRAD: per output row scale or nullptr
  const void* res; long long ldr; int res_fp32;
  const bf16* mask; long long ldm;
  int relu;
  float alpha;
  float mask_hi;             // > 0: the mask is a ReLU6 output, gradient also dies at mask >= mask_hi
  int epi_tma;               // FPROP/DGRAD bf16 output (and bf16 residual / mask) move through smem + TMA
  int stages;                // operand pipeline depth (shared memory is carved at launch time)
  int res_slots;             // per epilogue warp: 32x32 residual(+mask) tiles kept in flight by TMA
  int epi_warp_bytes;        // per epilogue warp: 2 output staging tiles + res_slots * slot bytes
  int cluster;               // 2: CTA pairs with multicast weight tiles (host-side dispatch only)
  int max_ctas;              // host-side: grid cap (0 = number of SMs)
  float* ws;                 // FPROP/DGRAD split-K: fp32 partial-sum tiles [tiles_m*tiles_n][BN/4][BM][4], all zero between launches
  int* ws_cnt;               // FPROP/DGRAD split-K: arrival counter per output tile, zero between launches
  int nprod;                 // TMA producer threads (1..3), K steps round-robin
  int epi_fast;              // FPROP/DGRAD: the streamlined drain (64-column stores, compile-time flags) applies
  float* pool;               // FPROP (streamlined drain): per 32-row group partial row sums instead of the output tensor
  int pool_hw;               //   rows per pooling window (>= 32): pool[(2g + s) * N + n], s = first / second window of group g
};

// One problem of a grouped launch (see mtl_conv_tc_group_*): its tensor maps, parameters and the first CTA-wide tile
// index it owns.  The array lives in global memory; TMA takes the maps by their (64-byte aligned) global address.
struct alignas(128) GroupEntry {
  CUtensorMap tmA, tmB, tmO, tmR, tmM;
  Params p;
  int tile_begin;            // first tile id of this problem in the concatenated tile space of the group
  int tiles;                 // its tile count (tiles_m * tiles_n * splits)
};

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// x / d for 0 <= x < 2^31 with a host-made constant: low 32 bits = multiplier ceil(2^s / d), high = s
__device__ __forceinline__ int fast_div(int x, unsigned long long c) {
  const uint32_t mul = (uint32_t)c, sh = (uint32_t)(c >> 32);
  return (int)(((unsigned long long)(uint32_t)x * mul) >> sh);
}
__device__ __forceinline__ bool mask_dead(float m, float hi) { return !(m > 0.0f) || (hi > 0.0f && !(m < hi)); }
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must kill the kernel (trap) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fffu) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("gemm_tc: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n",
               blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// TMA im2col load (rank-4 NHWC tensor map made by cuTensorMapEncodeIm2col): `pixels` consecutive
// output positions starting at base pixel (w, h, n) -- wrapping over rows / images inside the
// bounding box in hardware -- for filter offset (offw, offh), 64 channels from c; zero fill outside.
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c,
                                                int w, int h, int n, int offw, int offh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"((unsigned short)offw), "h"((unsigned short)offh)
      : "memory");
}
// The same load delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask` of the cluster.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// fp32 tile += shared-memory tile, element-wise in L2 (the tensor map carries the element type)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// commit that arrives on the barrier at the same offset in both CTAs of a pair
__device__ __forceinline__ void tcgen05_commit_mc2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp bit layout):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=2 (SW128).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BN> struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int DEF_STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;   // double-buffered accumulator (power of two >= 32)
};
constexpr int MAX_STAGES = 8;
constexpr int MAX_RES_SLOTS = 8;
// barrier block: full[8] empty[8] tfull[2] tempty[2] | tmem slot, split flag | res_bar[8 warps][8 slots]
constexpr int MAX_EPI_WARPS = 8;
constexpr int BAR_BYTES = 8 * (2 * MAX_STAGES + 4) + 8 + 8 * MAX_EPI_WARPS * MAX_RES_SLOTS;
constexpr int SMEM_LIMIT = 232448;           // 227 KB: most dynamic shared memory a CTA may opt in to
// shared memory of one launch: alignment slack + operand stages + epilogue staging + barriers
static inline int smem_bytes(int stage_bytes, int stages, int epi_warps, int epi_warp_bytes) {
  return 1024 + stages * stage_bytes + epi_warps * epi_warp_bytes + BAR_BYTES;
}

// CL = 2 (FPROP / DGRAD, TMA operands, no split-K): CTA pairs (thread-block clusters of two) work on
// vertically adjacent output tiles of the same column block; each CTA fetches half of the shared weight
// tile and multicasts it into both, which cuts the L2 -> SM operand traffic of a 128 x 256 tile by a third
// (the big second-stage GEMMs sit on the L2 fabric limit, ~6.1 kB/clk chip-wide, not on the tensor pipe).
template <int MODE, int BN, bool GATHER, int CL>
__global__ void __launch_bounds__(384, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmM, const __grid_constant__ Params p0,
               const GroupEntry* __restrict__ grp, const int ngrp) {
  // grp == nullptr: one problem, described by the kernel parameters.  Otherwise a grouped launch: `ngrp` independent
  // problems of the same kernel instance and shared-memory carve share one persistent grid; the CTA-wide tile index
  // runs over the concatenation of their tile spaces (tile t belongs to CTA t % gridDim.x), p0 = the first problem.
  using C = Cfg<BN>;
  const int STAGES = p0.stages;
  constexpr bool A_MN = (MODE == WGRAD);
  constexpr bool B_MN = (MODE != FPROP);
  // which operand the gather warps stage (the other always comes by TMA)
  constexpr bool GATHER_A = GATHER && (MODE != WGRAD);
  constexpr bool GATHER_B = GATHER && (MODE == WGRAD);
  // warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 epilogue; warps 8-11 stage the im2col operand
  // in the GATHER variants and are a second epilogue group otherwise (the drain of a 128 x BN tile is
  // issue-latency bound with one warp per scheduler: two groups take alternate 32-column chunks)
  constexpr int EPI_GROUPS = GATHER ? 1 : 2;
  constexpr int EPI_THREADS = 128 * EPI_GROUPS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = epi_base + (uint32_t)(4 * EPI_GROUPS) * (uint32_t)p0.epi_warp_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);
  const uint32_t split_flag = tmem_slot + 4u;
  auto res_bar = [&](int q, int b) { return bar_base + 8u * (2 * MAX_STAGES + 4) + 8u + 8u * (q * MAX_RES_SLOTS + b); };
  auto stage_a = [&](int s) { return smem_base + s * C::STAGE_BYTES; };
  auto stage_b = [&](int s) { return smem_base + s * C::STAGE_BYTES + A_STAGE_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 32) {      // descriptors are kernel parameters: fetch them while the previous kernel drains
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (p0.epi_tma && !grp) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
      if (p0.res) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmR)) : "memory");
      if (p0.mask) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmM)) : "memory");
    }
  }
  {
    // one barrier per thread (up to 84 of them): a serial init loop would cost ~1 us of prologue
    constexpr int NBAR = 2 * MAX_STAGES + 4;
    const int t = threadIdx.x - 64;          // warps 2-4 are idle here
    if (t >= 0 && t < NBAR) {
      uint32_t count = 1;
      if (t < MAX_STAGES) count = 1 + (GATHER ? NUM_GATHER_THREADS : 0);          // full[]
      else if (t < 2 * MAX_STAGES) count = CL;                                     // empty[]: every CTA of the pair
      else if (t >= 2 * MAX_STAGES + 2) count = EPI_THREADS;                       // tempty[]
      mbar_init(bar_base + 8u * t, count);
    } else if (t >= NBAR && t < NBAR + 4 * EPI_GROUPS * MAX_RES_SLOTS) {
      const int r = t - NBAR;
      mbar_init(res_bar(r / MAX_RES_SLOTS, r % MAX_RES_SLOTS), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  if (CL == 2) cluster_sync_all();      // the peer's barriers must exist before anything is multicast into it
  else __syncthreads();
  tcgen05_fence_after();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlaps the
  // tail of the previous kernel in the stream; global memory is only touched after this wait.
  // The next kernel may start its own prologue right away (this grid is fully resident).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  uint32_t cta_rank = 0;
  if (CL == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int ngj = grp ? ngrp : 1;
  // Per-problem view of a role's tile loop.  `p` is a by-value copy (only the fields a role touches are loaded, once
  // per problem); tile id -> (split, output tile): WGRAD walks all output tiles of one split first; FPROP/DGRAD keep
  // the splits of one output tile adjacent so that their partial sums meet in L2 at about the same time.
  // CL == 2: clusters walk (pair of m-tiles, n-tile); a pair's second m-tile may lie past the matrix (odd tile
  // count): that CTA still takes part in the multicast protocol but loads a clamped tile and stores nothing.
#define TC_PROBLEM_SETUP(gj)                                                                                          \
  const Params* const pp_ = grp ? &grp[gj].p : &p0;                                                                   \
  const Params p = *pp_;                                                                                              \
  const CUtensorMap* const tmA_p = grp ? &grp[gj].tmA : &tmA;                                                         \
  const CUtensorMap* const tmB_p = grp ? &grp[gj].tmB : &tmB;                                                         \
  const CUtensorMap* const tmO_p = grp ? &grp[gj].tmO : &tmO;                                                         \
  const CUtensorMap* const tmR_p = grp ? &grp[gj].tmR : &tmR;                                                         \
  const CUtensorMap* const tmM_p = grp ? &grp[gj].tmM : &tmM;                                                         \
  (void)tmA_p; (void)tmB_p; (void)tmO_p; (void)tmR_p; (void)tmM_p;                                                    \
  const int tiles_mn = p.tiles_m * p.tiles_n;                                                                         \
  const int total_tiles = tiles_mn * p.splits;                                                                        \
  const int ips = (p.k_iters + p.splits - 1) / p.splits;   /* K iterations per split */                               \
  const int gsz_ = (int)gridDim.x;                                                                                    \
  const int t_begin = CL == 2 ? (int)(blockIdx.x >> 1)                                                                \
                              : (grp ? (((int)blockIdx.x - grp[gj].tile_begin) % gsz_ + gsz_) % gsz_ : (int)blockIdx.x); \
  const int t_step = CL == 2 ? (int)(gridDim.x >> 1) : gsz_;                                                          \
  const int t_end = CL == 2 ? ((p.tiles_m + 1) >> 1) * p.tiles_n : total_tiles;                                       \
  (void)ips; (void)t_begin; (void)t_step; (void)t_end;                                                                \
  auto decode_tile = [&](int tile, int& split, int& rem) {                                                            \
    if (MODE == WGRAD) { split = tile / tiles_mn; rem = tile - split * tiles_mn; }                                    \
    else if (p.splits == 1) { split = 0; rem = tile; }                                                                \
    else { rem = tile / p.splits; split = tile - rem * p.splits; }                                                    \
  };                                                                                                                  \
  auto tile_coords = [&](int tile, int& split, int& m_tile, int& n_tile) {                                            \
    if (CL == 2) {                                                                                                    \
      split = 0;                                                                                                      \
      const int mp = tile / p.tiles_n;                                                                                \
      n_tile = tile - mp * p.tiles_n;                                                                                 \
      m_tile = 2 * mp + (int)cta_rank;                                                                                \
    } else {                                                                                                          \
      int rem;                                                                                                        \
      decode_tile(tile, split, rem);                                                                                  \
      m_tile = rem / p.tiles_n;                                                                                       \
      n_tile = rem - m_tile * p.tiles_n;                                                                              \
    }                                                                                                                 \
  };

  if (warp == 0 || warp == 2 || warp == 3) {
    // ============================ TMA producers (one thread in each of up to three warps) ====
    // One K step of DGRAD / WGRAD issues five or six TMA loads (the MN-major operand arrives in 64-column boxes);
    // with a single issuing thread those steps took longer than the 512 cycles their four MMAs need (measured on
    // B200, profiles/r2_kloop.md: dgrad 3x3 on 256 ROI tiles 79 -> 61 us, wgrad 86 -> 63 us with two producers;
    // FPROP, two loads per step, gains 5-15 %; a third producer adds nothing).  The otherwise idle warps 2 and 3
    // join warp 0 and take the K steps round-robin: step g of this CTA belongs to producer g % NPROD, lives in stage
    // g % STAGES and is signalled on that stage's barrier, so no ordering between the producers is needed.
    const int NPROD = p0.nprod;
    const int prod = warp == 0 ? 0 : warp - 1;
    if (lane == 0 && prod < NPROD) {
      int g = prod;                 // CTA-wide index of the next K step this producer issues
      int G0 = 0;                   // CTA-wide index of the current tile's first K step
      int s = prod;                 // g % STAGES (STAGES >= 3 = NPROD)
      uint32_t ph = 0;              // (g / STAGES) & 1
      constexpr uint32_t tx_bytes =
          (GATHER_A ? 0u : (uint32_t)A_STAGE_BYTES) + (GATHER_B ? 0u : (uint32_t)C::B_STAGE_BYTES);
      for (int gj = 0; gj < ngj; ++gj) {
      TC_PROBLEM_SETUP(gj)
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        int split, m_tile, n_tile;
        tile_coords(tile, split, m_tile, n_tile);
        if (CL == 2 && m_tile >= p.tiles_m) m_tile = p.tiles_m - 1;     // padding tile of an odd pair
        const int m0 = m_tile * BM;
        const int kb = split * ips, ke = min(kb + ips, p.k_iters);
        int k = kb + (g - G0);      // this producer's first K step inside the tile
        G0 += ke - kb;
        if (k >= ke) continue;
        int tn0 = 0, tp0 = 0, tq0 = 0;      // (n, p, q) of the tile's first row (im2col TMA base pixel)
        if (MODE != WGRAD && p.im2col) {
          tn0 = fast_div(m0, p.div_ohw);
          const int r2 = m0 - tn0 * (p.oH * p.oW);
          tp0 = fast_div(r2, p.div_ow);
          tq0 = r2 - tp0 * p.oW;
        }
        // Coordinates are carried incrementally (no division inside the K loop), NPROD steps at a time.
        if (MODE == FPROP || MODE == DGRAD) {
          int tap = 0, chunk = k;
          if (k >= p.cpt) { tap = k / p.cpt; chunk = k - tap * p.cpt; }
          int fr = 0, fs = tap;
          if (tap >= p.S) { fr = tap / p.S; fs = tap - fr * p.S; }
          const int ntap = (MODE == FPROP) ? p.gC : p.ntot;    // B-matrix columns per filter tap
          int acol = chunk * BK;
          int bbase = tap * ntap + (MODE == FPROP ? 0 : n_tile * BN);
          const int bw = tq0 + p.im_low_w, bh = tp0 + p.im_low_h, n0 = n_tile * BN;
          for (; k < ke; k += NPROD, g += NPROD) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t fb = full_bar(s), sa = stage_a(s);
            mbar_arrive_expect_tx(fb, tx_bytes);
            if (!GATHER_A) {
              if (p.im2col) {
                const int offh = (MODE == FPROP ? fr : p.R - 1 - fr) * p.dil;
                const int offw = (MODE == FPROP ? fs : p.S - 1 - fs) * p.dil;
                tma_load_im2col(sa, tmA_p, fb, acol, bw, bh, tn0, offw, offh);
              } else {
                tma_load_2d(sa, tmA_p, fb, acol, m0);
              }
            }
            if (CL == 2) {
              // this CTA's half of the weight tile, delivered to both CTAs of the pair
              if (MODE == FPROP) {
                tma_load_2d_mc(sa + A_STAGE_BYTES + cta_rank * (BN / 2) * 128, tmB_p, fb, bbase + acol,
                               n0 + (int)cta_rank * (BN / 2), (uint16_t)3);
              } else {
#pragma unroll
                for (int jj = 0; jj < BN / 128; ++jj) {
                  const int j = (int)cta_rank * (BN / 128) + jj;
                  tma_load_2d_mc(sa + A_STAGE_BYTES + j * 8192, tmB_p, fb, bbase + j * 64, acol, (uint16_t)3);
                }
              }
            } else if (MODE == FPROP) {
              tma_load_2d(sa + A_STAGE_BYTES, tmB_p, fb, bbase + acol, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sa + A_STAGE_BYTES + j * 8192, tmB_p, fb, bbase + j * 64, acol);
            }
            for (int j = 0; j < NPROD; ++j) {
              acol += BK;
              if (++chunk == p.cpt) {
                chunk = 0; acol = 0; bbase += ntap;
                if (++fs == p.S) { fs = 0; ++fr; }
              }
            }
            s += NPROD;
            if (s >= STAGES) { s -= STAGES; ph ^= 1u; }
          }
        } else {
          const int nper = (p.ntot + BN - 1) / BN;   // n-tiles per tap
          const int tap = n_tile / nper;
          const int ci0 = (n_tile - tap * nper) * BN;
          const int fr = tap / p.S, fs = tap - fr * p.S;
          for (; k < ke; k += NPROD, g += NPROD) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t fb = full_bar(s), sa = stage_a(s);
            mbar_arrive_expect_tx(fb, tx_bytes);
            const int p0 = k * BK;
            tma_load_2d(sa, tmA_p, fb, m0, p0);
            tma_load_2d(sa + 8192, tmA_p, fb, m0 + 64, p0);
            if (!GATHER_B) {
              if (p.im2col) {
                const int pn = fast_div(p0, p.div_ohw);
                const int r2 = p0 - pn * (p.oH * p.oW);
                const int pp = fast_div(r2, p.div_ow);
                const int pq = r2 - pp * p.oW;
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_im2col(sa + A_STAGE_BYTES + j * 8192, tmB_p, fb, ci0 + j * 64, pq + p.im_low_w,
                                  pp + p.im_low_h, pn, fs * p.dil, fr * p.dil);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_2d(sa + A_STAGE_BYTES + j * 8192, tmB_p, fb, ci0 + j * 64, p0);
              }
            }
            s += NPROD;
            if (s >= STAGES) { s -= STAGES; ph ^= 1u; }
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (one thread) ==============================
    if (lane == 0) {
      // instruction descriptor: fp32 accum, bf16 A/B, majors, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      uint32_t tcount = 0;
      int s = 0;
      uint32_t ph = 0;
      // The issue loop of this one thread bounds every K step (measured: 0.28 us per step whatever the tile width and
      // with the operand loads switched off, profiles/r2_kloop.md), so it carries no address arithmetic: the low word
      // of a shared-memory descriptor is (address >> 4) | LBO << 16 and all stage / K-slice addresses are below
      // 256 KB, i.e. advancing a descriptor is a 32-bit add on its low word; the high word (SBO, version, 128-byte
      // swizzle) never changes.
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t KS_A = (A_MN ? 2048u : 32u) >> 4, KS_B = (B_MN ? 2048u : 32u) >> 4;
      const uint32_t sstep = (uint32_t)C::STAGE_BYTES >> 4;
      const uint32_t lo_a0 = ((stage_a(0) >> 4) & 0x3fffu) | ((A_MN ? (8192u >> 4) : 1u) << 16);
      const uint32_t lo_b0 = ((stage_b(0) >> 4) & 0x3fffu) | ((B_MN ? (8192u >> 4) : 1u) << 16);
      uint32_t lo_a = lo_a0, lo_b = lo_b0;
      auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; };
      for (int gj = 0; gj < ngj; ++gj) {
      TC_PROBLEM_SETUP(gj)
      for (int tile = t_begin; tile < t_end; tile += t_step, ++tcount) {
        int split, m_tile_, n_tile_;
        tile_coords(tile, split, m_tile_, n_tile_);
        const int kb = split * ips, ke = min(kb + ips, p.k_iters);
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        uint32_t accum = 0u;
        for (int k = kb; k < ke; ++k) {
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          tcgen05_mma_bf16(tmem_d, desc(lo_a), desc(lo_b), idesc, accum);
          accum = 1u;
#pragma unroll
          for (uint32_t ks = 1; ks < BK / 16; ++ks)
            tcgen05_mma_bf16(tmem_d, desc(lo_a + ks * KS_A), desc(lo_b + ks * KS_B), idesc, 1u);
          // frees the smem stage when these MMAs retire (in both CTAs of a pair: the peer multicasts into it)
          if (CL == 2) tcgen05_commit_mc2(empty_bar(s));
          else tcgen05_commit(empty_bar(s));
          lo_a += sstep; lo_b += sstep;
          if (++s == STAGES) { s = 0; ph ^= 1u; lo_a = lo_a0; lo_b = lo_b0; }
        }
        tcgen05_commit(tfull_bar(acc));     // accumulator ready for the epilogue
      }
      }
    }
  } else if (warp >= 4 && (warp < 8 || !GATHER)) {
    // ============================ epilogue warps =======================================
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int egrp = (EPI_GROUPS == 2 && warp >= 8) ? 1 : 0;   // group g takes chunks g, g + EPI_GROUPS, ...
    const int ew = quad + 4 * egrp;         // epilogue warp index (staging memory, barriers)
    const int row = quad * 32 + lane;
    constexpr int NCH = BN / 32;            // 32-column chunks per tile
    constexpr int CPW = NCH / EPI_GROUPS;   // chunks per warp per tile
    uint32_t tcount = 0;
    uint32_t gchunk = 0;                    // stored-chunk counter (output staging buffer parity)
    // Residual / mask tiles (32 rows x 32 columns bf16) are fetched by TMA into a ring of `res_slots`
    // slots per warp, `res_slots` chunks ahead of their use and across tile boundaries, so the ~1 us
    // HBM round trip of a 500 MB residual stream never sits on the drain path.  Chunks are numbered
    // in processing order; chunk g lives in slot g % res_slots.  iq_* is the issue cursor, cq_slot /
    // slot_ph the consume cursor and the expected phase bit of every slot's barrier.
    // (A grouped launch re-primes the ring at every problem: all its problems share the carve, i.e. the slot count
    // and which of residual / mask exist, and the issue and consume cursors meet again at each problem's end.)
    const uint32_t ebase = epi_base + (uint32_t)ew * (uint32_t)p0.epi_warp_bytes;
    int iq_slot = 0, cq_slot = 0;
    uint32_t slot_ph = 0;
    for (int gj = 0; gj < ngj; ++gj) {
    TC_PROBLEM_SETUP(gj)
    if (MODE != WGRAD && p.epi_fast) {
      // ---------------- streamlined drain (host: TMA epilogue, no split-K, N a multiple of 64) ----------------
      // The generic drain below spends ~490 warp instructions per 32 x 32 chunk (ncu source view, second-stage
      // conv3: 190 of them useful), mostly run-time flag tests, rematerialised addresses and register moves, and with
      // two epilogue warps per scheduler it is issue-latency bound: a 128 x 256 tile takes 4.3 us to drain against
      // 2.2 us of MMAs in a K loop of 8 steps.  Here the flags are compile-time, a warp owns 64-column PAIRS of
      // chunks (two TMEM loads in flight, one 32 x 64 staging tile, one TMA store of full 128-byte rows per pair)
      // and the residual / mask ring keeps its 32-column slots and its order (pair q, half h -> chunk 2q + h).
      // 64-wide tiles (the trunk at batch 1): one 32-column chunk per warp, same code with H = 1.
      constexpr int EG = EPI_GROUPS;
      constexpr int CW = BN >= 128 ? 64 : 32;      // columns a warp handles per step (a pair of chunks, or one)
      constexpr int H = CW / 32;                   // 32-column halves per step
      constexpr int PPW = BN / (CW * EG);          // steps per warp per tile
      constexpr int CPT = H * PPW;                 // ring chunks per warp per tile
      const int R = p.res_slots;
      const int Nn = p.N;
      const int m_lim = p.tiles_m * BM;
      auto drain = [&](auto BIAS_, auto RES_, auto MASK_, auto RELU_, auto POOL_) {
        constexpr bool BIAS = decltype(BIAS_)::value, RES = decltype(RES_)::value, MASK = decltype(MASK_)::value;
        constexpr bool POOL = decltype(POOL_)::value;
        constexpr int RELU = decltype(RELU_)::value;
        constexpr bool RING = RES || MASK;
        constexpr uint32_t SLOT = 2048u * ((RES ? 1u : 0u) + (MASK ? 1u : 0u));
        const float bscale = p.bias_scale;
        const float mhi = p.mask_hi;
        const float* const bias = p.bias;
        const uint32_t ring = ebase + 4096u;
        const uint32_t rbar0 = res_bar(ew, 0);
        const uint32_t swz = ((uint32_t)(lane >> 1) & 3u);
        const uint32_t srow = ebase + (uint32_t)lane * 128u;
        const uint32_t sx = (uint32_t)lane & 7u;
        int iq_tile = t_begin, iq_c = 0, iq_m0 = 0, iq_n0 = 0;
        auto issue_next = [&]() {
          if (!RING) return;
          if (iq_tile < t_end) {
            if (iq_c == 0) {
              int sp, mt, nt;
              tile_coords(iq_tile, sp, mt, nt);
              iq_m0 = mt * BM + quad * 32;       // a padding tile starts past the last row: nothing to fetch
              iq_n0 = nt * BN + egrp * CW;
            }
            const int n0 = iq_n0 + (iq_c / H) * (CW * EG) + (iq_c % H) * 32;
            if (n0 < Nn && iq_m0 < m_lim && lane == 0) {
              const uint32_t rb = rbar0 + 8u * (uint32_t)iq_slot;
              const uint32_t dst = ring + (uint32_t)iq_slot * SLOT;
              mbar_arrive_expect_tx(rb, SLOT);
              if (RES) tma_load_2d(dst, tmR_p, rb, n0, iq_m0);
              if (MASK) tma_load_2d(dst + (RES ? 2048u : 0u), tmM_p, rb, n0, iq_m0);
            }
          }
          if (++iq_c == CPT) { iq_c = 0; iq_tile += t_step; }
          if (++iq_slot == R) iq_slot = 0;
        };
        if (RING)
          for (int i = 0; i < R; ++i) issue_next();
        for (int tile = t_begin; tile < t_end; tile += t_step, ++tcount) {
          int split, m_tile, n_tile;
          tile_coords(tile, split, m_tile, n_tile);
          const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
          mbar_wait(tfull_bar(acc), aph);
          tcgen05_fence_after();
          if (CL == 2 && m_tile >= p.tiles_m) {
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(acc));
            for (int i = 0; i < CPT; ++i) {
              if (RING && ++cq_slot == R) cq_slot = 0;
              issue_next();
            }
            continue;
          }
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + egrp * CW;
          const int m0w = m_tile * BM + quad * 32;
          const int ncol0 = n_tile * BN + egrp * CW;
          uint32_t va[32], vb[32];
          tmem_ld32(taddr, va);
          if (H == 2) tmem_ld32(taddr + 32, vb);
#pragma unroll 1
          for (int i = 0; i < PPW; ++i) {
            const int n0 = ncol0 + i * (CW * EG);
            tmem_ld_wait();
            if (i == PPW - 1) {
              // both halves of the last pair are in registers: the MMA warp may reuse the accumulator
              tcgen05_fence_before();
              mbar_arrive(tempty_bar(acc));
            }
            if (n0 >= Nn) {        // tile columns past the matrix (N is a multiple of 64: whole pairs)
              for (int h = 0; h < H; ++h) {
                if (RING && ++cq_slot == R) cq_slot = 0;
                issue_next();
              }
              if (i + 1 < PPW) {
                tmem_ld32(taddr + (i + 1) * (CW * EG), va);
                if (H == 2) tmem_ld32(taddr + (i + 1) * (CW * EG) + 32, vb);
              }
              continue;
            }
            uint32_t o[16 * H];    // the step's 64 (32) outputs, packed bf16
            auto half = [&](uint32_t (&v)[32], auto H_) {
              constexpr int h = decltype(H_)::value;
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              if (BIAS) {
                const float4* bq = reinterpret_cast<const float4*>(bias + n0 + h * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = __ldg(bq + j);
                  f[4 * j] += b.x * bscale; f[4 * j + 1] += b.y * bscale;
                  f[4 * j + 2] += b.z * bscale; f[4 * j + 3] += b.w * bscale;
                }
              }
              uint32_t rrow = 0;
              if (RING) {
                const int slot = cq_slot;
                if (++cq_slot == R) cq_slot = 0;
                mbar_wait(rbar0 + 8u * (uint32_t)slot, (slot_ph >> slot) & 1u);
                slot_ph ^= 1u << slot;
                rrow = ring + (uint32_t)slot * SLOT + (uint32_t)lane * 64u;
              }
              if (RES) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 t = lds128(rrow + ((j ^ swz) << 4));
                  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    f[8 * j + 2 * q] += __uint_as_float(w[q] << 16);
                    f[8 * j + 2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
                  }
                }
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                if (RELU) hh = __hmax2(hh, __float2bfloat162_rn(0.0f));
                if (RELU == 2) hh = __hmin2(hh, __float2bfloat162_rn(6.0f));
                o[16 * h + j] = *reinterpret_cast<const uint32_t*>(&hh);
              }
              if (MASK) {
                const uint32_t mrow = rrow + (RES ? 2048u : 0u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 t = lds128(mrow + ((j ^ swz) << 4));
                  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    uint32_t keep = 0xffffffffu;
                    if (mask_dead(__uint_as_float(w[q] << 16), mhi)) keep &= 0xffff0000u;
                    if (mask_dead(__uint_as_float(w[q] & 0xffff0000u), mhi)) keep &= 0x0000ffffu;
                    o[16 * h + 4 * j + q] &= keep;
                  }
                }
              }
              if (RING) {
                __syncwarp();        // every lane has read the slot: refill it with the chunk `R` ahead
                issue_next();
              }
            };
            half(va, std::integral_constant<int, 0>{});
            if (i + 1 < PPW) tmem_ld32(taddr + (i + 1) * (CW * EG), va);
            if (H == 2) {
              half(vb, std::integral_constant<int, H - 1>{});
              if (i + 1 < PPW) tmem_ld32(taddr + (i + 1) * (CW * EG) + 32, vb);
            }
            if (!POOL && lane == 0) bulk_wait_read<0>();     // the previous pair's store has read the staging tile
            __syncwarp();
            if (H == 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts128(srow + (((uint32_t)j ^ sx) << 4), make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
            } else {       // 32 columns: 64-byte rows, 64 B swizzle (the generic drain's staging layout)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(ebase + (uint32_t)lane * 64u + (((uint32_t)j ^ swz) << 4),
                       make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
            }
            if (POOL && H == 2) {
              // Fused spatial mean (forward-only tails): the 32 x 64 tile is not stored; lane l sums columns 2l, 2l+1
              // down the rows of the group's first / second pooling window (a window has >= 32 rows: at most two per
              // group) and writes the two partial sums.  Fixed summation order: deterministic.
              __syncwarp();
              const int rows_ok = min(32, p.M - m0w);
              if (rows_ok > 0) {
                const int hw = p.pool_hw;
                const int b = min((m0w / hw + 1) * hw - m0w, rows_ok);
                const uint32_t unit = (uint32_t)lane >> 2, wsel = ((uint32_t)lane & 3u) * 4u;
                float s00 = 0.0f, s01 = 0.0f, s10 = 0.0f, s11 = 0.0f;
                // fully unrolled (all 32 loads in flight; b and rows_ok are warp-uniform: predicated adds, same
                // ascending order as a loop)
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                  const uint32_t w = lds32(ebase + (uint32_t)r * 128u + ((unit ^ ((uint32_t)r & 7u)) << 4) + wsel);
                  const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
                  if (r < b) { s00 += lo; s01 += hi; }
                  else if (r < rows_ok) { s10 += lo; s11 += hi; }
                }
                float* dst = p.pool + (long long)(m0w >> 5) * 2 * Nn + n0 + 2 * lane;
                *reinterpret_cast<float2*>(dst) = make_float2(s00, s01);
                *reinterpret_cast<float2*>(dst + Nn) = make_float2(s10, s11);
              }
              __syncwarp();        // every lane has read the tile: the next pair may overwrite it
              continue;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(tmO_p, ebase, n0, m0w);
              bulk_commit();
            }
          }
        }
        if (lane == 0) bulk_wait_read<0>();       // the next problem (or the generic drain) may reuse the staging tile
        __syncwarp();
      };
      using T_ = std::true_type; using F_ = std::false_type;
      using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
      const bool hr = p.res != nullptr, hm = p.mask != nullptr;
      if (MODE == FPROP) {          // host guarantees: bias, no mask (pooled output: residual + ReLU, the bottleneck's conv3)
        if (p.pool) drain(T_{}, T_{}, F_{}, I1{}, T_{});
        else if (hr) { if (p.relu == 0) drain(T_{}, T_{}, F_{}, I0{}, F_{}); else if (p.relu == 1) drain(T_{}, T_{}, F_{}, I1{}, F_{}); else drain(T_{}, T_{}, F_{}, I2{}, F_{}); }
        else    { if (p.relu == 0) drain(T_{}, F_{}, F_{}, I0{}, F_{}); else if (p.relu == 1) drain(T_{}, F_{}, F_{}, I1{}, F_{}); else drain(T_{}, F_{}, F_{}, I2{}, F_{}); }
      } else {                      // DGRAD: no bias, no activation
        if (hr) { if (hm) drain(F_{}, T_{}, T_{}, I0{}, F_{}); else drain(F_{}, T_{}, F_{}, I0{}, F_{}); }
        else    { if (hm) drain(F_{}, F_{}, T_{}, I0{}, F_{}); else drain(F_{}, F_{}, F_{}, I0{}, F_{}); }
      }
      continue;
    }
    const int R = (MODE != WGRAD && p.epi_tma) ? p.res_slots : 0;
    const bool has_res = p.res != nullptr;
    const bool has_mask = p.mask != nullptr;
    const uint32_t slot_bytes = 2048u * ((has_res ? 1u : 0u) + (has_mask ? 1u : 0u));
    int iq_tile = t_begin, iq_c = 0, iq_m0 = 0, iq_n0 = 0;
    auto issue_next = [&](int cur_tile) {
      if (R == 0) return;
      if (p.splits > 1 && iq_tile != cur_tile) return;   // split-K: only the finishing CTA reads them
      if (iq_tile < t_end) {
        if (iq_c == 0) {
          int sp, mt, nt;
          tile_coords(iq_tile, sp, mt, nt);
          iq_m0 = mt * BM + quad * 32;       // a padding tile starts past the last row: nothing to fetch
          iq_n0 = nt * BN;
        }
        const int n0 = iq_n0 + (egrp + iq_c * EPI_GROUPS) * 32;
        if (n0 < p.N && iq_m0 < p.tiles_m * BM && lane == 0) {
          const uint32_t rb = res_bar(ew, iq_slot);
          const uint32_t dst = ebase + 4096u + (uint32_t)iq_slot * slot_bytes;
          mbar_arrive_expect_tx(rb, slot_bytes);
          if (has_res) tma_load_2d(dst, tmR_p, rb, n0, iq_m0);
          if (has_mask) tma_load_2d(dst + (has_res ? 2048u : 0u), tmM_p, rb, n0, iq_m0);
        }
      }
      if (++iq_c == CPW) { iq_c = 0; iq_tile += t_step; }
      if (++iq_slot == R) iq_slot = 0;
    };
    if (p.splits == 1)
      for (int i = 0; i < R; ++i) issue_next(-1);
    for (int tile = t_begin; tile < t_end; tile += t_step, ++tcount) {
      int split, m_tile, n_tile;
      tile_coords(tile, split, m_tile, n_tile);
      const int rem = m_tile * p.tiles_n + n_tile;
      const int m = m_tile * BM + row;
      const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
      mbar_wait(tfull_bar(acc), aph);
      tcgen05_fence_after();
      if (CL == 2 && m_tile >= p.tiles_m) {
        // padding tile of an odd pair: release the accumulator, keep the residual ring cursors in step
        tcgen05_fence_before();
        mbar_arrive(tempty_bar(acc));
        for (int i = 0; i < CPW; ++i) {
          if (R && ++cq_slot == R) cq_slot = 0;
          issue_next(tile);
        }
        continue;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
      long long ncol0;     // first output column of this tile
      int wg_col_in_tap = 0;   // WGRAD: channel offset of this tile inside its filter tap
      if (MODE == WGRAD) {
        const int nper = (p.ntot + BN - 1) / BN;
        ncol0 = (long long)(n_tile / nper) * p.ntot + (long long)(n_tile % nper) * BN;
        wg_col_in_tap = (n_tile % nper) * BN;
      } else {
        ncol0 = (long long)n_tile * BN;
      }
      const bool row_ok = m < p.M;
      // Software-pipelined drain: the TMEM load of chunk c+1 is in flight while chunk c is combined
      // and stored, so the tcgen05.ld -> store dependency chain is paid once per tile, not per chunk.
      uint32_t v[32];
      tmem_ld32(taddr + egrp * 32, v);
      if (MODE == WGRAD) {
        const float sc = p.alpha * ((p.rowscale && row_ok) ? __ldg(p.rowscale + m) : 1.0f);
        if (p.epi_tma) {
          // 32 x 32 fp32 chunks are staged in swizzled shared memory and added to dw by TMA reduce-add:
          // full 128-byte rows reach L2 instead of 32 scattered 16-byte atomics per warp instruction.
          // Padding columns (channels past C) hold exact zeros and rows past K are clipped by hardware.
          const uint32_t sw = (uint32_t)lane & 7u;
#pragma unroll 1
          for (int c = egrp; c < NCH; c += EPI_GROUPS) {
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * sc;
            if (c + EPI_GROUPS < NCH) tmem_ld32(taddr + (c + EPI_GROUPS) * 32, v);
            if (wg_col_in_tap + c * 32 >= p.ntot) continue;
            if (lane == 0) bulk_wait_read<0>();       // the previous reduce has read the staging tile
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(ebase + lane * 128u + ((j ^ sw) << 4),
                     make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3])));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(tmO_p, ebase, (int)(ncol0 + c * 32), m_tile * BM + quad * 32);
              bulk_commit();
            }
          }
        } else {
#pragma unroll 1
        for (int c = egrp; c < NCH; c += EPI_GROUPS) {
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * sc;
          if (c + EPI_GROUPS < NCH) tmem_ld32(taddr + (c + EPI_GROUPS) * 32, v);
          if (row_ok) {      // columns past C inside a tap are padding (C need not be a multiple of BN)
            float* o = reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + ncol0 + c * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (wg_col_in_tap + c * 32 + 4 * j < p.ntot)
                atomicAdd(reinterpret_cast<float4*>(o) + j, make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
          }
        }
        }
      } else if (p.epi_tma) {
        // Output, residual and mask tiles travel through swizzled shared memory and TMA: no per-thread
        // 64-byte strided global accesses, rows/columns outside the tensor are clipped by hardware.
        const int m0w = m_tile * BM + quad * 32;
        const uint32_t swz = ((uint32_t)(lane >> 1) & 3u);
        bool from_ws = false;
        float4* wst = nullptr;          // this thread's column of the workspace tile [BN/4 float4 columns][128 rows]
        if (p.splits > 1) {
          // Split-K: add this CTA's partial tile into the fp32 workspace (L2 reduces); the CTA that
          // arrives last reads the sum back, runs the fused epilogue on it and re-zeroes the workspace.
          wst = reinterpret_cast<float4*>(p.ws) + (long long)rem * (BM * BN / 4) + row;
#pragma unroll 1
          for (int c = egrp; c < NCH; c += EPI_GROUPS) {
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (c + EPI_GROUPS < NCH) tmem_ld32(taddr + (c + EPI_GROUPS) * 32, v);
            if (row_ok && ncol0 + c * 32 < p.N) {
              float4* q = wst + c * (8 * BM);      // a warp's 32 lanes hit 512 contiguous bytes
#pragma unroll
              for (int j = 0; j < 8; ++j)
                atomicAdd(q + j * BM, make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
            }
          }
          tcgen05_fence_before();
          mbar_arrive(tempty_bar(acc));
          __threadfence();
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
          if (threadIdx.x == 128) {
            const int old = atomicAdd(p.ws_cnt + rem, 1);
            const uint32_t fin = (old == p.splits - 1) ? 1u : 0u;
            if (fin) p.ws_cnt[rem] = 0;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(split_flag), "r"(fin) : "memory");
          }
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
          uint32_t fin;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(fin) : "r"(split_flag) : "memory");
          if (!fin) continue;
          __threadfence();
          from_ws = true;
          iq_tile = tile; iq_c = 0; iq_slot = cq_slot;
          for (int i = 0; i < R && i < CPW; ++i) issue_next(tile);
        }
        auto load_acc = [&](int c) {
          if (from_ws) {
            if (row_ok && ncol0 + c * 32 < p.N) {
              float4* q = wst + c * (8 * BM);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = __ldcg(q + j * BM);
                v[4 * j] = __float_as_uint(t.x); v[4 * j + 1] = __float_as_uint(t.y);
                v[4 * j + 2] = __float_as_uint(t.z); v[4 * j + 3] = __float_as_uint(t.w);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) __stcg(q + j * BM, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            }
          } else {
            tmem_ld32(taddr + c * 32, v);
          }
        };
        if (from_ws) load_acc(egrp);
#pragma unroll 1
        for (int c = egrp; c < NCH; c += EPI_GROUPS) {
          const long long n0 = ncol0 + c * 32;
          const int nvalid = (int)min((long long)32, (long long)p.N - n0);
          const bool vec = (nvalid == 32);
          // the bias vector does not depend on the accumulator: fetch it while the TMEM load is in flight
          float4 bq[8];
          if (p.bias && vec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
          }
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (c + EPI_GROUPS < NCH) {
            load_acc(c + EPI_GROUPS);
          } else if (!from_ws) {
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(acc));       // accumulator fully drained: the MMA warp may reuse it
          }
          const int slot = cq_slot;
          if (R && ++cq_slot == R) cq_slot = 0;
          if (nvalid <= 0) { issue_next(tile); continue; }
          const uint32_t buf = gchunk & 1u;
          if (p.bias) {
            if (vec) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[4 * j] += bq[j].x * p.bias_scale; f[4 * j + 1] += bq[j].y * p.bias_scale;
                f[4 * j + 2] += bq[j].z * p.bias_scale; f[4 * j + 3] += bq[j].w * p.bias_scale;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) f[j] += __ldg(p.bias + n0 + j) * p.bias_scale;
            }
          }
          const uint32_t rrow = ebase + 4096u + (uint32_t)slot * slot_bytes + lane * 64u;
          if (R) {
            mbar_wait(res_bar(ew, slot), (slot_ph >> slot) & 1u);
            slot_ph ^= 1u << slot;
          }
          if (has_res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = lds128(rrow + ((j ^ swz) << 4));
              const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                f[8 * j + 2 * q] += __uint_as_float(w[q] << 16);
                f[8 * j + 2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
              }
            }
          }
          // round to bf16 first: ReLU / ReLU6 / the gradient mask commute with the rounding and cost half as
          // many instructions on packed pairs
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            if (p.relu) {
              h = __hmax2(h, __float2bfloat162_rn(0.0f));
              if (p.relu == 2) h = __hmin2(h, __float2bfloat162_rn(6.0f));
            }
            o[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          if (has_mask) {      // rows outside the tensor arrive as zeros: dead
            const uint32_t mrow = rrow + (has_res ? 2048u : 0u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = lds128(mrow + ((j ^ swz) << 4));
              const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t keep = 0xffffffffu;
                if (mask_dead(__uint_as_float(w[q] << 16), p.mask_hi)) keep &= 0xffff0000u;
                if (mask_dead(__uint_as_float(w[q] & 0xffff0000u), p.mask_hi)) keep &= 0x0000ffffu;
                o[4 * j + q] &= keep;
              }
            }
          }
          if (R) {
            __syncwarp();        // every lane has read the slot: refill it with the chunk `R` ahead right away
            issue_next(tile);
          }
          // staging buffer `buf` is free: lane 0 waited (end of the previous chunk) for the store issued from
          // it two chunks ago, and the __syncwarp that followed ordered that wait before these writes
          const uint32_t orow = ebase + buf * 2048u + lane * 64u;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(orow + ((j ^ swz) << 4), make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
          fence_proxy_async();
          __syncwarp();        // staging tile complete
          if (lane == 0) {
            tma_store_2d(tmO_p, ebase + buf * 2048u, (int)n0, m0w);
            bulk_commit();
          }
          if (lane == 0) bulk_wait_read<1>();     // the other staging buffer has been read by its store
          __syncwarp();
          ++gchunk;            // counts stored chunks only: consecutive stores always alternate buffers
        }
        continue;     // tempty already signalled
      } else {
        // Direct epilogue (fp32 outputs / residuals, unaligned pitches: the small FC heads): per-thread
        // row accesses straight to global memory.
        const bool res_bf16 = p.res && !p.res_fp32 && (p.ldr & 7) == 0;
        const bool mask_vec = p.mask && (p.ldm & 7) == 0;
#pragma unroll 1
        for (int c = egrp; c < NCH; c += EPI_GROUPS) {
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (c + EPI_GROUPS < NCH) tmem_ld32(taddr + (c + EPI_GROUPS) * 32, v);
          const long long n0 = ncol0 + c * 32;
          if (!row_ok) continue;
          const int nvalid = (int)min((long long)32, (long long)p.N - n0);
          if (nvalid <= 0) continue;
          const bool vec = (nvalid == 32);
          if (p.bias) {
            if (vec) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
                f[4 * j] += t.x * p.bias_scale; f[4 * j + 1] += t.y * p.bias_scale;
                f[4 * j + 2] += t.z * p.bias_scale; f[4 * j + 3] += t.w * p.bias_scale;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) f[j] += __ldg(p.bias + n0 + j) * p.bias_scale;
            }
          }
          if (p.res) {
            if (p.res_fp32) {
              const float* r = reinterpret_cast<const float*>(p.res) + (long long)m * p.ldr + n0;
              if (vec && (p.ldr & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 t = __ldg(reinterpret_cast<const float4*>(r) + j);
                  f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nvalid) f[j] += __ldg(r + j);
              }
            } else if (vec && res_bf16) {
              const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.res) + (long long)m * p.ldr + n0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 t = __ldg(r + j);
                const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  f[8 * j + 2 * q] += __uint_as_float(w[q] << 16);
                  f[8 * j + 2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
                }
              }
            } else {
              const bf16* r = reinterpret_cast<const bf16*>(p.res) + (long long)m * p.ldr + n0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) f[j] += __bfloat162float(r[j]);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
            if (p.relu == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fminf(f[j], 6.0f);
            }
          }
          if (p.mask) {
            if (vec && mask_vec) {
              const uint4* r = reinterpret_cast<const uint4*>(p.mask + (long long)m * p.ldm + n0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 t = __ldg(r + j);
                const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (mask_dead(__uint_as_float(w[q] << 16), p.mask_hi)) f[8 * j + 2 * q] = 0.0f;
                  if (mask_dead(__uint_as_float(w[q] & 0xffff0000u), p.mask_hi)) f[8 * j + 2 * q + 1] = 0.0f;
                }
              }
            } else {
              const bf16* r = p.mask + (long long)m * p.ldm + n0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid && mask_dead(__bfloat162float(r[j]), p.mask_hi)) f[j] = 0.0f;
            }
          }
          if (p.out_fp32) {
            float* o = reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + n0;
            if (vec && (p.ldo & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(o)[j] =
                    make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) o[j] = f[j];
            }
          } else {
            bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)m * p.ldo + n0;
            if (vec && (p.ldo & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 t;
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * j + 2 * q], f[8 * j + 2 * q + 1]);
                  w[q] = *reinterpret_cast<const uint32_t*>(&h);
                }
                t.x = w[0]; t.y = w[1]; t.z = w[2]; t.w = w[3];
                reinterpret_cast<uint4*>(o)[j] = t;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) o[j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
    }
    // the staging tiles must have been read before the CTA's shared memory is released; the global writes
    // themselves complete before the grid does (and before a dependent grid's griddepcontrol.wait returns)
    if (p0.epi_tma && lane == 0) bulk_wait_read<0>();
  } else if (GATHER && warp >= 8) {
    // ============================ im2col gather warps ==================================
    // Each of the 128 threads owns one 128-byte row (FPROP/DGRAD: one output pixel of the A
    // tile; WGRAD: one pixel of the B tile for half of the 64-channel column blocks).
    constexpr int DEPTH = (BN <= 128) ? 3 : 2;   // cp.async groups kept in flight besides the current one (< stages)
    const int g = threadIdx.x - 256;
    uint32_t it = 0;
    uint32_t pending_first = 0;    // oldest iteration whose full-barrier arrive is outstanding
    int s = 0, pf_s = 0;           // stage of iteration `it` / of iteration `pending_first`
    uint32_t ph = 0;
    // The gather is on the critical path of every 3x3 / strided layer: all per-iteration index math
    // is strength-reduced (no divisions inside the K loop; row decode uses multiply-shift division).
    for (int gj = 0; gj < ngj; ++gj) {
    TC_PROBLEM_SETUP(gj)
    const int ohw = p.oH * p.oW;
    for (int tile = t_begin; tile < t_end; tile += t_step) {
      int split, m_tile, n_tile;
      tile_coords(tile, split, m_tile, n_tile);
      const int kb = split * ips, ke = min(kb + ips, p.k_iters);
      if (GATHER_A) {
        const int m = m_tile * BM + g;
        const bool rvalid = m < p.rows;
        int rn = 0, roh = 0, row_ = 0;
        if (rvalid) {
          rn = fast_div(m, p.div_ohw);
          const int r2 = m - rn * ohw;
          roh = fast_div(r2, p.div_ow);
          row_ = r2 - roh * p.oW;
        }
        const bf16* img = p.gsrc + (long long)rn * p.gH * p.gW * p.gld;
        // fprop: input coordinate of tap (0,0); dgrad: output-gradient coordinate before the stride division
        const int h0 = p.transposed ? roh + p.pad_h : roh * p.stride - p.pad_h;
        const int w0 = p.transposed ? row_ + p.pad_w : row_ * p.stride - p.pad_w;
        int tap = kb / p.cpt;
        int chunk = kb - tap * p.cpt;
        int fr = tap / p.S, fs = tap - fr * p.S;
        const uint32_t dst_row = g * 128;
        const int sw = g & 7;
        for (int k = kb; k < ke; ++k, ++it) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          int hi, wi;
          bool ok = rvalid;
          if (!p.transposed) {
            hi = h0 + fr * p.dil;
            wi = w0 + fs * p.dil;
          } else {
            hi = h0 - fr * p.dil;
            wi = w0 - fs * p.dil;
            ok = ok && hi >= 0 && wi >= 0;
            if (p.stride != 1) {
              ok = ok && (hi % p.stride == 0) && (wi % p.stride == 0);
              hi /= p.stride;
              wi /= p.stride;
            }
          }
          ok = ok && hi >= 0 && hi < p.gH && wi >= 0 && wi < p.gW;
          const int c0 = chunk * BK;
          const bf16* src = ok ? img + ((hi * p.gW + wi) * p.gld + c0) : p.gsrc;
          const uint32_t dst = stage_a(s) + dst_row;
          if (c0 + BK <= p.gC) {
            const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) cp_async16(dst + ((q ^ sw) << 4), src + (ok ? q * 8 : 0), nb);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const bool pv = ok && (c0 + q * 8 < p.gC);
              cp_async16(dst + ((q ^ sw) << 4), pv ? (const void*)(src + q * 8) : (const void*)p.gsrc, pv ? 16u : 0u);
            }
          }
          if (++chunk == p.cpt) {
            chunk = 0;
            if (++fs == p.S) { fs = 0; ++fr; }
          }
          cp_async_commit();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (it - pending_first >= (uint32_t)DEPTH) {
            cp_async_wait<DEPTH>();
            fence_proxy_async();
            mbar_arrive(full_bar(pf_s));
            if (++pf_s == STAGES) pf_s = 0;
            ++pending_first;
          }
        }
      } else {
        // WGRAD: rows are pixels of this K block, columns are ci of tap (n_tile / nper)
        const int nper = (p.ntot + BN - 1) / BN;
        const int tap = n_tile / nper, ci0 = (n_tile - tap * nper) * BN;
        const int fr = tap / p.S, fs = tap - fr * p.S;
        const int dh = fr * p.dil - p.pad_h, dw = fs * p.dil - p.pad_w;
        const int pr = g & 63;
        const int sw = pr & 7;
        for (int k = kb; k < ke; ++k, ++it) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const int pix = k * BK + pr;
          bool ok = pix < p.rows;
          const int n = fast_div(pix, p.div_ohw);
          const int r2 = pix - n * ohw;
          const int oh = fast_div(r2, p.div_ow);
          const int ow = r2 - oh * p.oW;
          const int hi = oh * p.stride + dh;
          const int wi = ow * p.stride + dw;
          ok = ok && hi >= 0 && hi < p.gH && wi >= 0 && wi < p.gW;
          const bf16* src = ok ? p.gsrc + ((long long)(n * p.gH + hi) * p.gW + wi) * p.gld + ci0 : p.gsrc;
#pragma unroll
          for (int jj = 0; jj < BN / 128; ++jj) {
            const int j = (g >> 6) + 2 * jj;
            const uint32_t dst = stage_b(s) + j * 8192 + pr * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const bool pv = ok && (ci0 + j * 64 + q * 8 < p.gC);      // channels past C are padding
              cp_async16(dst + ((q ^ sw) << 4), src + (pv ? j * 64 + q * 8 : 0), pv ? 16u : 0u);
            }
          }
          if (BN == 64) {   // single column block: threads 0..63 stage it, 64..127 only arrive
            if ((g >> 6) == 0) {
              const uint32_t dst = stage_b(s) + pr * 128;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const bool pv = ok && (ci0 + q * 8 < p.gC);
                cp_async16(dst + ((q ^ sw) << 4), src + (pv ? q * 8 : 0), pv ? 16u : 0u);
              }
            }
          }
          cp_async_commit();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (it - pending_first >= (uint32_t)DEPTH) {
            cp_async_wait<DEPTH>();
            fence_proxy_async();
            mbar_arrive(full_bar(pf_s));
            if (++pf_s == STAGES) pf_s = 0;
            ++pending_first;
          }
        }
      }
    }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (; pending_first < it; ++pending_first) {
      mbar_arrive(full_bar(pf_s));
      if (++pf_s == STAGES) pf_s = 0;
    }
  }

#undef TC_PROBLEM_SETUP
  tcgen05_fence_before();
  if (CL == 2) cluster_sync_all();      // no CTA leaves while its peer may still signal or write into it
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 row-major matrix [rows, cols] with row pitch ld (elements); box = 64 cols x box_rows.
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                    int box_rows, int box_cols = 64, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
                    bool fp32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { mtl_set_error("gemm_tc: cuTensorMapEncodeTiled unavailable"); return MTL_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & (fp32 ? 3 : 7))) {
    mtl_set_error("gemm_tc: TMA operand must be 16B aligned with a 16B-multiple row pitch (ld=%lld)", ld);
    return MTL_ERR_ARG;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * (fp32 ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mtl_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return MTL_ERR_CUDA; }
  return MTL_OK;
}

static unsigned long long make_fast_div(int d) {
  // Granlund-Montgomery round-up method, exact for 0 <= x < 2^31 (d >= 1)
  if (d <= 1) return 1ull;                      // mul = 1, shift = 0
  int l = 0;
  while ((1ll << l) < d) ++l;
  const int s = 31 + l;
  const unsigned long long mul = ((1ull << s) + (unsigned long long)d - 1) / (unsigned long long)d;
  return (mul & 0xffffffffull) | ((unsigned long long)s << 32);
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC bf16 activation [N,H,W,C] as an im2col tensor map: 64 channels x `pixels` positions per load;
// the base pixel ranges over [low, dim + up) in each spatial dimension (stride-1 traversal).
static int make_im2col_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int low_h, int low_w,
                           int up_h, int up_w, int pixels, long long ld = 0) {
  if (ld == 0) ld = C;
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      mtl_set_error("gemm_tc: cuTensorMapEncodeIm2col unavailable");
      return MTL_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeIm2colFn>(ptr);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (C & 7) || (ld & 7)) {
    mtl_set_error("gemm_tc: im2col operand must be 16B aligned with C %% 8 == 0");
    return MTL_ERR_ARG;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  int lower[2] = {low_w, low_h};
  int upper[2] = {up_w, up_h};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, lower, upper, 64u,
                  (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mtl_set_error("gemm_tc: cuTensorMapEncodeIm2col failed (%d)", (int)r); return MTL_ERR_CUDA; }
  return MTL_OK;
}

template <int MODE, int BN, bool GATHER, int CL>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
                  const CUtensorMap& tmM, const Params& p, cudaStream_t stream, const GroupEntry* grp = nullptr,
                  int ngrp = 0, int group_tiles = 0) {
  using C = Cfg<BN>;
  auto kern = tc_gemm_kernel<MODE, BN, GATHER, CL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) { mtl_set_error("gemm_tc: smem attribute: %s", cudaGetErrorString(e)); return MTL_ERR_CUDA; }
    attr_set = true;
  }
  const int smem = smem_bytes(C::STAGE_BYTES, p.stages, GATHER ? 4 : 8, p.epi_warp_bytes);
  if (smem > SMEM_LIMIT || p.stages < 3 || p.stages > MAX_STAGES || (GATHER && BN <= 128 && p.stages < 4)) {
    mtl_set_error("gemm_tc: bad shared-memory carve (stages %d, epilogue %d B/warp, %d B)", p.stages,
                  p.epi_warp_bytes, smem);
    return MTL_ERR_ARG;
  }
  int grid;
  if (CL == 2) {
    // persistent clusters: never more than can be co-resident (a GPC with an odd SM count strands one SM)
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(mtl_num_sms() / 2 * 2);
      q.blockDim = dim3(384);
      q.dynamicSmemBytes = SMEM_LIMIT;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = mtl_num_sms() / 2 - 2;
      }
      max_clusters = n < mtl_num_sms() / 2 ? n : mtl_num_sms() / 2;
    }
    const int pairs = ((p.tiles_m + 1) / 2) * p.tiles_n;
    const int clusters = pairs < max_clusters ? pairs : max_clusters;
    grid = 2 * clusters;
  } else {
    const int total = grp ? group_tiles : p.tiles_m * p.tiles_n * p.splits;
    const int cap = (p.max_ctas > 0 && p.max_ctas < mtl_num_sms()) ? p.max_ctas : mtl_num_sms();
    grid = total < cap ? total : cap;
  }
  static const bool no_pdl = getenv("MTL_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CL == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (CL == 2 && grp) { mtl_set_error("gemm_tc: grouped launches do not pair CTAs"); return MTL_ERR_UNSUPPORTED; }
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, tmR, tmM, p, grp, ngrp);
  if (le != cudaSuccess) {
    mtl_set_error("tc_gemm_kernel: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return MTL_ERR_CUDA;
  }
  MTL_CUDA_LAUNCH_CHECK("tc_gemm_kernel");
  return MTL_OK;
}

struct Maps { CUtensorMap a, b, o, r, m; };

template <int MODE, bool GATHER>
static int dispatch_bn(int bn, const Maps& t, const Params& p, cudaStream_t st, const GroupEntry* grp = nullptr,
                       int ngrp = 0, int group_tiles = 0) {
  if (p.cluster == 2) {       // FPROP / DGRAD, TMA operands, 128- or 256-wide tiles (chosen by the caller)
    if (MODE != WGRAD && !GATHER) {
      constexpr int M2 = (MODE == WGRAD || GATHER) ? FPROP : MODE;     // keeps the dead branch instantiable
      if (bn == 256) return launch<M2, 256, false, 2>(t.a, t.b, t.o, t.r, t.m, p, st);
      if (bn == 128) return launch<M2, 128, false, 2>(t.a, t.b, t.o, t.r, t.m, p, st);
    }
    mtl_set_error("gemm_tc: CTA pairs need FPROP/DGRAD with TMA operands and BN >= 128");
    return MTL_ERR_UNSUPPORTED;
  }
  switch (bn) {
    case 256: return launch<MODE, 256, GATHER, 1>(t.a, t.b, t.o, t.r, t.m, p, st, grp, ngrp, group_tiles);
    case 128: return launch<MODE, 128, GATHER, 1>(t.a, t.b, t.o, t.r, t.m, p, st, grp, ngrp, group_tiles);
    case 64: return launch<MODE, 64, GATHER, 1>(t.a, t.b, t.o, t.r, t.m, p, st, grp, ngrp, group_tiles);
  }
  mtl_set_error("gemm_tc: unsupported BN %d", bn);
  return MTL_ERR_UNSUPPORTED;
}

// Tile width: the widest tile that still yields enough CTAs (measured on B200: for M = 2394 rows
// 64-wide tiles win by 10-20 %, for >= 100 tiles of width 128 narrower tiles only add operand re-reads).
static int pick_bn(int M, int N, int k_iters) {
  const int sms = mtl_num_sms();
  const int tm = ceil_div(M, BM);
  // short K loops over few rows (block3 conv3 and its mirror dgrad at batch 1: 19 row tiles, 4 K steps) are
  // pure drain: two waves of narrow tiles beat half a wave of wide ones (7.8 vs 9.2 us)
  if (tm <= 32 && k_iters <= 8 && N >= 512 && tm * ceil_div(N, 256) < sms) return 64;
  int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  if (bn == 256 && tm * ceil_div(N, 256) < sms / 2) bn = 128;
  if (bn == 128 && tm * ceil_div(N, 128) < (sms * 2) / 5) bn = 64;
  return bn;
}

// FPROP / DGRAD tile width and split-K factor.  Split-K (fp32 workspace, see the epilogue) pays off only
// when a long K loop meets too few output tiles to fill the machine: the 3x3 trunk layers at batch 1
// (2394 pixels: 13.2 vs 14.3 us) and the RPN 3x3 conv (31 vs 47 us), measured with tools/sweep_conv.py.
static void plan_tile(int M, int N, int k_iters, bool can_split, int force_bn, int force_splits, int* bn_out,
                      int* splits_out) {
  int bn = force_bn ? force_bn : pick_bn(M, N, k_iters);
  int splits = 1;
  if (can_split) {
    if (force_splits > 0) {
      splits = force_splits;
    } else if (!force_bn) {
      const int sms = mtl_num_sms(), tm = ceil_div(M, BM);
      if (k_iters >= 96 && N >= 256 && tm * ceil_div(N, 256) * 3 <= sms) { bn = 256; splits = 3; }
      else if (k_iters >= 32 && N >= 128 && tm * ceil_div(N, 128) * 3 <= sms) { bn = 128; splits = 3; }
    }
    if (splits > k_iters) splits = k_iters;
    if (splits < 1) splits = 1;
    splits = ceil_div(k_iters, ceil_div(k_iters, splits));     // every split owns at least one K iteration
  }
  *bn_out = bn; *splits_out = splits;
}

// CTA pairs with multicast weight tiles: worth it once there is at least a full wave of wide tiles
static int pick_cluster(int M, int N, int bn, int splits, bool gather, int force) {
  static const bool off = getenv("MTL_NO_CLUSTER") != nullptr;
  if (gather || splits > 1 || bn < 128 || off || force == 1) return 1;
  if (force == 2) return 2;
  // Measured on B200 (tools/sweep_conv.py pairs): within +-2 % of the unpaired schedule on every second-stage
  // GEMM -- L2 already merges the near-simultaneous requests of neighbouring CTAs for one weight tile -- so
  // pairs stay opt-in (MTL_CLUSTER=1 / force_cluster=2).
  static const bool on = getenv("MTL_CLUSTER") != nullptr;
  (void)M; (void)N;
  return on && ceil_div(M, BM) * ceil_div(N, bn) >= mtl_num_sms() ? 2 : 1;
}

// the output (and residual / mask) of an FPROP / DGRAD call can travel through the TMA epilogue
static bool epi_tma_ok(const void* out, int out_fp32, long long ldo, const void* res, int res_fp32, long long ldr,
                       const void* mask, long long ldm) {
  static const bool no_epi_tma = getenv("MTL_NO_TMA_EPILOGUE") != nullptr;
  return !no_epi_tma && !out_fp32 && (ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
         (!res || (!res_fp32 && (ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(res) & 15) == 0)) &&
         (!mask || ((ldm & 7) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0));
}

}  // namespace tc

// Public C ABI --------------------------------------------------------------------------
struct mtl_conv_args {
  int mode;                 // 0 fprop, 1 dgrad, 2 wgrad
  int N, H, W, C;           // input-side activation geometry x[N,H,W,C]
  int K;                    // output channels
  int R, S, stride, pad_h, pad_w, dil;
  int P, Q;                 // output-side spatial dims y[N,P,Q,K]
  const void* x;            // fprop/wgrad: x (bf16). dgrad: unused
  const void* w;            // fprop/dgrad: w[K,R,S,C] bf16. wgrad: unused
  const void* dy;           // dgrad/wgrad: dy[N,P,Q,K] bf16
  void* out;                // fprop: y; dgrad: dx; wgrad: dw fp32 [K,R,S,C] (accumulated)
  int out_fp32;             // fprop/dgrad: output element type
  const float* bias;        // fprop: per-K bias (or null)
  const float* rowscale;    // wgrad: per-K scale (or null)
  const void* res; int res_fp32;   // tensor added before relu/mask, same shape as out
  const void* mask;         // bf16 tensor, same shape as out: out = mask > 0 ? out : 0
  int relu;
  float alpha;              // wgrad scale
  int force_bn;             // 0 = auto
  int force_splits;         // 0 = auto
  float mask_hi;            // dgrad: > 0 -> mask is a ReLU6 output (gradient only where 0 < mask < mask_hi)
  // channel-slice support (concatenated branches): pixel pitches in elements, 0 = dense
  long long dy_ld, out_ld, res_ld, mask_ld;
  float bias_scale;         // 0 = 1.0
  int force_stages;         // 0 = auto operand pipeline depth
  void* ws;                 // fprop/dgrad split-K workspace (all zero between launches) or null
  long long ws_bytes;
  int force_cluster;        // 0 auto, 1 never pair CTAs, 2 pair CTAs (fprop/dgrad with TMA operands)
  int max_ctas;             // 0 = whole GPU; otherwise the persistent grid (and wgrad's split-K) is sized for this many
                            // SMs: work that overlaps a latency-bound chain on another stream leaves it room
  float* pool_out;          // fprop (+ residual + ReLU): partial row sums per 32-row group instead of y (see mtlssl.h)
  int pool_hw;
};

// Everything mtl_conv_tc decides on the host for one problem: kernel parameters, tensor maps, tile width, operand path.
struct mtl_conv_plan {
  tc::Params p;
  tc::Maps t;
  int bn;
  bool gather;
  int mode;
};

static int plan_conv(const mtl_conv_args* a, mtl_conv_plan* out) {
  using namespace tc;
  MTL_CHECK_ARG(a != nullptr, "mtl_conv_tc: null args");
  MTL_CHECK_ARG(a->C % 8 == 0 && a->K % 8 == 0, "mtl_conv_tc: C and K must be multiples of 8 (C=%d K=%d)", a->C, a->K);
  MTL_CHECK_ARG(a->N > 0 && a->H > 0 && a->W > 0 && a->P > 0 && a->Q > 0, "mtl_conv_tc: empty geometry");
  const bool plain = (a->R == 1 && a->S == 1 && a->stride == 1 && a->pad_h == 0 && a->pad_w == 0);
  if (plain) MTL_CHECK_ARG(a->P == a->H && a->Q == a->W, "mtl_conv_tc: 1x1 geometry mismatch");
  const long long npq = (long long)a->N * a->P * a->Q, nhw = (long long)a->N * a->H * a->W;
  MTL_CHECK_ARG(npq < (1ll << 31) && nhw < (1ll << 31), "mtl_conv_tc: too many pixels");

  Params p;
  memset(&p, 0, sizeof(p));
  p.S = a->S; p.stride = a->stride; p.pad_h = a->pad_h; p.pad_w = a->pad_w; p.dil = a->dil > 0 ? a->dil : 1;
  p.out = a->out; p.out_fp32 = a->out_fp32; p.bias = a->bias; p.rowscale = a->rowscale;
  p.res = a->res; p.res_fp32 = a->res_fp32; p.mask = reinterpret_cast<const bf16*>(a->mask);
  p.relu = a->relu; p.mask_hi = a->mask_hi; p.alpha = a->alpha;
  static const int nprod_env = getenv("MTL_NPROD") ? atoi(getenv("MTL_NPROD")) : 2;
  p.nprod = nprod_env < 1 ? 1 : (nprod_env > 3 ? 3 : nprod_env);
  p.bias_scale = a->bias_scale != 0.0f ? a->bias_scale : 1.0f; p.splits = 1; p.cluster = 1; p.max_ctas = a->max_ctas; p.taps = a->R * a->S;
  CUtensorMap tmA, tmB;
  memset(&tmA, 0, sizeof(tmA)); memset(&tmB, 0, sizeof(tmB));
  int bn, rc;
  bool epi_ok = false;
  // stride-1 filters: the gathered operand is fetched by TMA im2col loads (no gather warps)
  static const bool no_im2col = getenv("MTL_NO_TMA_IM2COL") != nullptr;
  const bool im2col = !plain && !no_im2col && a->stride == 1 && a->R * p.dil < 120 && a->S * p.dil < 120 &&
                      a->pad_h < 120 && a->pad_w < 120;
  const bool gather = !plain && !im2col;
  p.im2col = im2col ? 1 : 0;
  p.R = a->R;
  if (a->mode == FPROP) {
    MTL_CHECK_ARG(a->x && a->w && a->out, "mtl_conv_tc fprop: null tensor");
    p.M = (int)npq; p.N = a->K; p.cpt = ceil_div(a->C, BK); p.k_iters = p.taps * p.cpt;
    p.gsrc = reinterpret_cast<const bf16*>(a->x); p.gH = a->H; p.gW = a->W; p.gC = a->C; p.gld = a->C;
    p.oH = a->P; p.oW = a->Q; p.rows = p.M; p.transposed = 0;
    p.ldo = a->out_ld ? a->out_ld : a->K; p.ldr = a->res_ld ? a->res_ld : a->K;
    p.ldm = a->mask_ld ? a->mask_ld : a->K;
    epi_ok = epi_tma_ok(a->out, a->out_fp32, p.ldo, a->res, a->res_fp32, p.ldr, a->mask, p.ldm);
    plan_tile(p.M, p.N, p.k_iters, epi_ok && a->ws, a->force_bn, a->force_splits, &bn, &p.splits);
    if (a->pool_out && !a->force_bn && bn < 128) { bn = 128; p.splits = 1; }      // pooled output: the streamlined drain only
    p.cluster = pick_cluster(p.M, p.N, bn, p.splits, gather, a->force_cluster);
    if (im2col) {
      p.im_low_h = -a->pad_h; p.im_low_w = -a->pad_w;
      if ((rc = make_im2col_map(&tmA, a->x, a->N, a->H, a->W, a->C, p.im_low_h, p.im_low_w,
                                a->P + p.im_low_h - a->H, a->Q + p.im_low_w - a->W, BM))) return rc;
    } else if (!gather && (rc = make_map(&tmA, a->x, npq, a->C, a->C, BM))) return rc;
    // CTA pairs: each CTA fetches (and multicasts) half of the weight tile
    if ((rc = make_map(&tmB, a->w, a->K, (long long)p.taps * a->C, (long long)p.taps * a->C, bn / p.cluster))) return rc;
    if (gather) tmA = tmB;
  } else if (a->mode == DGRAD) {
    MTL_CHECK_ARG(a->dy && a->w && a->out, "mtl_conv_tc dgrad: null tensor");
    p.M = (int)nhw; p.N = a->C; p.cpt = ceil_div(a->K, BK); p.k_iters = p.taps * p.cpt;
    const long long dyld = a->dy_ld ? a->dy_ld : a->K;
    p.gsrc = reinterpret_cast<const bf16*>(a->dy); p.gH = a->P; p.gW = a->Q; p.gC = a->K; p.gld = (int)dyld;
    p.oH = a->H; p.oW = a->W; p.rows = p.M; p.transposed = 1; p.ntot = a->C;
    p.ldo = a->out_ld ? a->out_ld : a->C; p.ldr = a->res_ld ? a->res_ld : a->C;
    p.ldm = a->mask_ld ? a->mask_ld : a->C;
    epi_ok = epi_tma_ok(a->out, a->out_fp32, p.ldo, a->res, a->res_fp32, p.ldr, a->mask, p.ldm);
    plan_tile(p.M, p.N, p.k_iters, epi_ok && a->ws, a->force_bn, a->force_splits, &bn, &p.splits);
    p.cluster = pick_cluster(p.M, p.N, bn, p.splits, gather, a->force_cluster);
    if (im2col) {
      // dx[h] = sum_r dy[h + pad - r*dil]: a stride-1 correlation over dy with mirrored filter offsets
      p.im_low_h = a->pad_h - (a->R - 1) * p.dil; p.im_low_w = a->pad_w - (a->S - 1) * p.dil;
      if ((rc = make_im2col_map(&tmA, a->dy, a->N, a->P, a->Q, a->K, p.im_low_h, p.im_low_w,
                                a->H + p.im_low_h - a->P, a->W + p.im_low_w - a->Q, BM, dyld))) return rc;
    } else if (!gather && (rc = make_map(&tmA, a->dy, npq, a->K, dyld, BM))) return rc;
    if ((rc = make_map(&tmB, a->w, a->K, (long long)p.taps * a->C, (long long)p.taps * a->C, 64))) return rc;
    if (gather) tmA = tmB;
  } else if (a->mode == WGRAD) {
    MTL_CHECK_ARG(a->dy && a->x && a->out, "mtl_conv_tc wgrad: null tensor");
    const long long dyld = a->dy_ld ? a->dy_ld : a->K;
    p.M = a->K; p.N = p.taps * a->C; p.cpt = 1; p.k_iters = (int)ceil_div_ll(npq, BK);
    p.gsrc = reinterpret_cast<const bf16*>(a->x); p.gH = a->H; p.gW = a->W; p.gC = a->C; p.gld = a->C;
    p.oH = a->P; p.oW = a->Q; p.rows = (int)npq; p.transposed = 0; p.ntot = a->C;
    p.ldo = (long long)p.taps * a->C;
    // channels per tap need not fill the tile: columns past C are zero-filled operands and masked stores
    bn = a->force_bn ? a->force_bn : (a->C >= 256 ? 256 : (a->C > 64 ? 128 : 64));
    if (!a->force_bn && bn == 256 && a->C % 256 != 0 && a->C % 256 <= 128) bn = 128;
    if ((rc = make_map(&tmA, a->dy, npq, a->K, dyld, 64))) return rc;
    if (im2col) {
      p.im_low_h = -a->pad_h; p.im_low_w = -a->pad_w;
      if ((rc = make_im2col_map(&tmB, a->x, a->N, a->H, a->W, a->C, p.im_low_h, p.im_low_w,
                                a->P + p.im_low_h - a->H, a->Q + p.im_low_w - a->W, 64))) return rc;
    } else if (!gather && (rc = make_map(&tmB, a->x, nhw, a->C, a->C, 64))) return rc;
    if (gather) tmB = tmA;
    // split K (pixels) so that about one wave of CTAs is launched
    const int tiles = ceil_div(p.M, BM) * p.taps * ceil_div(a->C, bn);
    int splits = a->force_splits;
    if (splits <= 0) {
      // minimise waves x (K iterations per CTA + drain cost): never spill a few tiles into a second wave
      // reductions of up to 256 K steps (the trunk at batch 1, the 64- and 256-ROI tails) run beside the dgrad chain
      // that feeds them: leave
      // that chain its SMs instead of grabbing the whole machine for a few microseconds (MTL_WGRAD_SMALL_CTAS)
      // measured on the full step: 7.51 -> 7.37 ms with a cap anywhere in 32..72; 0 switches it off
      static const int small_cap = getenv("MTL_WGRAD_SMALL_CTAS") ? atoi(getenv("MTL_WGRAD_SMALL_CTAS")) : 64;
      static const int small_k = getenv("MTL_WGRAD_SMALL_K") ? atoi(getenv("MTL_WGRAD_SMALL_K")) : 256;
      int sms = (small_cap > 0 && p.k_iters <= small_k) ? (small_cap < mtl_num_sms() ? small_cap : mtl_num_sms())
                                                   : mtl_num_sms();
      if (a->max_ctas > 0 && a->max_ctas < sms) sms = a->max_ctas;
      const int drain = bn / 48 + 1;
      long long best = -1;
      splits = 1;
      for (int sp = 1; sp <= p.k_iters && sp <= 64 && (sp == 1 || tiles * sp <= 2 * sms); ++sp) {
        const long long cost = (long long)ceil_div(tiles * sp, sms) * (ceil_div(p.k_iters, sp) + drain);
        if (best < 0 || cost < best) { best = cost; splits = sp; }
      }
    }
    if (splits > p.k_iters) splits = p.k_iters;
    if (splits < 1) splits = 1;
    const int ips = ceil_div(p.k_iters, splits);
    p.splits = ceil_div(p.k_iters, ips);   // every split owns at least one K iteration
  } else {
    mtl_set_error("mtl_conv_tc: bad mode %d", a->mode);
    return MTL_ERR_ARG;
  }
  p.div_ohw = make_fast_div(p.oH * p.oW);
  p.div_ow = make_fast_div(p.oW);
  p.tiles_m = ceil_div(p.M, BM);
  p.tiles_n = (a->mode == WGRAD) ? p.taps * ceil_div(a->C, bn) : ceil_div(p.N, bn);
  // epilogue through shared memory + TMA whenever the output (and residual / mask) are plain bf16 matrices
  Maps t;
  memset(&t, 0, sizeof(t));
  t.a = tmA; t.b = tmB;
  static const bool no_epi_tma = getenv("MTL_NO_TMA_EPILOGUE") != nullptr;
  p.epi_tma = 0;
  int nmaps = 0;
  if (a->mode != WGRAD && epi_ok) {
    // streamlined drain (see the kernel): whole 64-column pairs, no split-K, the flag combinations of the conv layers
    static const bool no_fast = getenv("MTL_NO_FAST_EPI") != nullptr;
    p.epi_fast = !no_fast && p.splits == 1 && (p.N % 64) == 0 &&
                 (a->mode == FPROP ? (a->bias != nullptr && a->mask == nullptr)
                                   : (a->bias == nullptr && a->relu == 0));
    if (a->pool_out) {
      if (!p.epi_fast || bn < 128 || a->mode != FPROP || !a->res || a->relu != 1 || a->pool_hw < 32) {
        mtl_set_error("gemm_tc: pooled output needs the streamlined drain (bf16, N %% 64 == 0, tile width >= 128), a "
                      "residual, ReLU and windows of >= 32 rows");
        return MTL_ERR_UNSUPPORTED;
      }
      p.pool = a->pool_out; p.pool_hw = a->pool_hw;
    }
    if (p.epi_fast && bn >= 128) { if ((rc = make_map(&t.o, a->out, p.M, p.N, p.ldo, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc; }
    else
    if ((rc = make_map(&t.o, a->out, p.M, p.N, p.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (a->res && (rc = make_map(&t.r, a->res, p.M, p.N, p.ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (a->mask && (rc = make_map(&t.m, a->mask, p.M, p.N, p.ldm, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (!a->res) t.r = t.o;
    if (!a->mask) t.m = t.o;
    p.epi_tma = 1;
    nmaps = (a->res ? 1 : 0) + (a->mask ? 1 : 0);
  } else if (a->mode == WGRAD && !no_epi_tma && (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0) {
    if ((rc = make_map(&t.o, a->out, p.M, p.N, p.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B, true))) return rc;
    t.r = t.o; t.m = t.o;
    p.epi_tma = 1;
  } else {
    t.o = tmB; t.r = tmB; t.m = tmB;
  }
  // Shared-memory carve.  Layers with a residual / mask stream give up one or two operand stages for a
  // deep ring of residual tiles in the epilogue (their K loops are short, the 2nd-stage residual
  // stream is not L2 resident).
  const int stage_bytes = A_STAGE_BYTES + bn * 128;
  const int def_stages = bn == 256 ? 4 : (bn == 128 ? 6 : 8);
  p.stages = def_stages;
  if (nmaps) p.stages = bn == 256 ? 3 : (bn == 128 ? 5 : 6);
  if (a->force_stages) p.stages = a->force_stages;
  p.res_slots = 0;
  p.epi_warp_bytes = 4096;
  if (p.epi_tma) {
    const int epi_warps = gather ? 4 : 8;
    int blocks = (SMEM_LIMIT - 1024 - BAR_BYTES - p.stages * stage_bytes) / (epi_warps * 2048);   // 2 KB tiles per warp
    if (blocks < 2 + nmaps) {
      mtl_set_error("gemm_tc: %d stages leave no room for the epilogue", p.stages);
      return MTL_ERR_ARG;
    }
    if (nmaps) {
      p.res_slots = (blocks - 2) / nmaps;
      if (p.res_slots > MAX_RES_SLOTS) p.res_slots = MAX_RES_SLOTS;
    }
    p.epi_warp_bytes = 2048 * (2 + p.res_slots * nmaps);     // WGRAD: one 32 x 32 fp32 staging tile
  }
  // FPROP / DGRAD split-K (plan_tile): needs the TMA epilogue and a zeroed workspace from the caller
  if (a->mode != WGRAD && p.splits > 1) {
    const int tiles = p.tiles_m * p.tiles_n;
    const long long need = (long long)tiles * BM * bn * 4 + (long long)tiles * 4;
    if (need > a->ws_bytes) {
      mtl_set_error("gemm_tc: split-K workspace too small (%lld < %lld bytes)", a->ws_bytes, need);
      return MTL_ERR_ARG;
    }
    p.ws = reinterpret_cast<float*>(a->ws);
    p.ws_cnt = reinterpret_cast<int*>(p.ws + (long long)tiles * BM * bn);
  }
  out->p = p; out->t = t; out->bn = bn; out->gather = gather; out->mode = a->mode;
  return MTL_OK;
}

static int run_plan(const mtl_conv_plan& pl, cudaStream_t stream, const tc::GroupEntry* grp = nullptr, int ngrp = 0,
                    int group_tiles = 0) {
  using namespace tc;
  const int bn = pl.bn;
  const bool gather = pl.gather;
  if (pl.mode == FPROP) return gather ? dispatch_bn<FPROP, true>(bn, pl.t, pl.p, stream, grp, ngrp, group_tiles)
                                      : dispatch_bn<FPROP, false>(bn, pl.t, pl.p, stream, grp, ngrp, group_tiles);
  if (pl.mode == DGRAD) return gather ? dispatch_bn<DGRAD, true>(bn, pl.t, pl.p, stream, grp, ngrp, group_tiles)
                                      : dispatch_bn<DGRAD, false>(bn, pl.t, pl.p, stream, grp, ngrp, group_tiles);
  return gather ? dispatch_bn<WGRAD, true>(bn, pl.t, pl.p, stream, grp, ngrp, group_tiles)
                : dispatch_bn<WGRAD, false>(bn, pl.t, pl.p, stream, grp, ngrp, group_tiles);
}

extern "C" int mtl_conv_tc(const mtl_conv_args* a, cudaStream_t stream) {
  mtl_conv_plan pl;
  int rc = plan_conv(a, &pl);
  if (rc) return rc;
  return run_plan(pl, stream);
}

// ---------------------------------------------------------------------------------------- grouped launches
// Several INDEPENDENT problems of one kernel instance (same mode, tile width, operand path and shared-memory carve)
// in one persistent launch.  The trunk's weight-gradient GEMMs at batch 1 are the case: 81 launches of 8-54 CTAs, each
// paying ~4 us of launch / fill / drain for 3-10 us of MMA work next to a latency-bound dgrad chain, become one grid
// whose CTAs walk the concatenated tile space (K never split: every tile of dw is written by exactly one CTA, which
// also makes the sums reproducible).
//   mtl_conv_tc_group_entry_bytes()            size of one table entry (128-byte aligned)
//   mtl_conv_tc_group_build(args, n, target_k_iters, host_table, info)
//        plans the n problems and writes the table into HOST memory (n entries); info[0..3] = total tiles, tile
//        width, mode, operand path.  K is split into pieces of about `target_k_iters` 64-deep steps (0: never).
//        The caller copies the table to device memory (any 128-byte aligned allocation) and keeps both alive.
//   mtl_conv_tc_group_launch(host_table, dev_table, n, info, max_ctas, stream)
extern "C" long long mtl_conv_tc_group_entry_bytes() { return (long long)sizeof(tc::GroupEntry); }

// Problems with equal keys can share a grouped launch (kernel instance + shared-memory carve); < 0: planning failed.
extern "C" long long mtl_conv_tc_group_key(const mtl_conv_args* a) {
  mtl_conv_args b = *a;
  b.force_splits = 1; b.force_cluster = 1;
  if (b.mode != tc::WGRAD) b.ws = nullptr;
  mtl_conv_plan pl;
  if (plan_conv(&b, &pl)) return -1;
  const tc::Params& p = pl.p;
  return (long long)pl.mode | ((long long)pl.bn << 4) | ((long long)(pl.gather ? 1 : 0) << 16) |
         ((long long)p.stages << 20) | ((long long)p.res_slots << 24) | ((long long)p.epi_tma << 28) |
         ((long long)(p.res != nullptr) << 29) | ((long long)(p.mask != nullptr) << 30) |
         ((long long)(p.epi_warp_bytes >> 10) << 32);
}

extern "C" int mtl_conv_tc_group_build(const mtl_conv_args* args, int n, int target_k_iters, void* host_table,
                                       int* info) {
  using namespace tc;
  MTL_CHECK_ARG(args && n > 0 && host_table && info, "mtl_conv_tc_group_build: bad args");
  GroupEntry* tab = reinterpret_cast<GroupEntry*>(host_table);
  int tile = 0;
  mtl_conv_plan first;
  for (int i = 0; i < n; ++i) {
    mtl_conv_args a = args[i];
    if (a.mode == WGRAD) {
      const long long npq = (long long)a.N * a.P * a.Q;
      const int k_iters = (int)ceil_div_ll(npq, BK);
      a.force_splits = target_k_iters > 0 ? (k_iters + target_k_iters / 2) / target_k_iters : 1;
      if (a.force_splits < 1) a.force_splits = 1;
    } else {
      a.force_splits = 1;       // FPROP / DGRAD split-K needs a workspace and a last-arriver epilogue: not in groups
      a.ws = nullptr;
    }
    a.force_cluster = 1;
    mtl_conv_plan pl;
    int rc = plan_conv(&a, &pl);
    if (rc) return rc;
    if (i == 0) first = pl;
    const Params &p = pl.p, &q = first.p;
    if (pl.mode != first.mode || pl.bn != first.bn || pl.gather != first.gather || p.stages != q.stages ||
        p.epi_warp_bytes != q.epi_warp_bytes || p.res_slots != q.res_slots || p.epi_tma != q.epi_tma ||
        (p.res != nullptr) != (q.res != nullptr) || (p.mask != nullptr) != (q.mask != nullptr) || p.cluster != 1 ||
        (pl.mode != WGRAD && p.splits != 1)) {
      mtl_set_error("mtl_conv_tc_group_build: problem %d does not share problem 0's kernel instance / carve "
                    "(mode %d/%d bn %d/%d gather %d/%d stages %d/%d)", i, pl.mode, first.mode, pl.bn, first.bn,
                    (int)pl.gather, (int)first.gather, p.stages, q.stages);
      return MTL_ERR_UNSUPPORTED;
    }
    GroupEntry& e = tab[i];
    memset(&e, 0, sizeof(e));
    e.tmA = pl.t.a; e.tmB = pl.t.b; e.tmO = pl.t.o; e.tmR = pl.t.r; e.tmM = pl.t.m;
    e.p = p;
    e.tile_begin = tile;
    e.tiles = p.tiles_m * p.tiles_n * p.splits;
    tile += e.tiles;
  }
  info[0] = tile; info[1] = first.bn; info[2] = first.mode; info[3] = first.gather ? 1 : 0;
  return MTL_OK;
}

extern "C" int mtl_conv_tc_group_launch(const void* host_table, const void* dev_table, int n, const int* info,
                                        int max_ctas, cudaStream_t stream) {
  using namespace tc;
  MTL_CHECK_ARG(host_table && dev_table && n > 0 && info, "mtl_conv_tc_group_launch: bad args");
  MTL_CHECK_ARG((reinterpret_cast<uintptr_t>(dev_table) & 127) == 0, "mtl_conv_tc_group_launch: table must be 128-byte aligned");
  const GroupEntry* tab = reinterpret_cast<const GroupEntry*>(host_table);
  mtl_conv_plan pl;
  pl.p = tab[0].p;
  pl.p.max_ctas = max_ctas;
  pl.t.a = tab[0].tmA; pl.t.b = tab[0].tmB; pl.t.o = tab[0].tmO; pl.t.r = tab[0].tmR; pl.t.m = tab[0].tmM;
  pl.bn = info[1]; pl.mode = info[2]; pl.gather = info[3] != 0;
  return run_plan(pl, stream, reinterpret_cast<const GroupEntry*>(dev_table), n, info[0]);
}

// Bytes of zeroed split-K workspace mtl_conv_tc would use for this FPROP / DGRAD problem (0: it would not
// split).  The caller allocates it once (zero filled; the kernel leaves it zeroed) and passes it as args->ws.
extern "C" long long mtl_conv_tc_ws_bytes(const mtl_conv_args* a) {
  using namespace tc;
  if (!a || a->mode == WGRAD || a->out_fp32) return 0;
  const long long M = a->mode == FPROP ? (long long)a->N * a->P * a->Q : (long long)a->N * a->H * a->W;
  const int N = a->mode == FPROP ? a->K : a->C;
  const int red = a->mode == FPROP ? a->C : a->K;
  if (M <= 0 || M >= (1ll << 31) || N <= 0) return 0;
  const int k_iters = a->R * a->S * ceil_div(red, BK);
  const long long ldo = a->out_ld ? a->out_ld : N, ldr = a->res_ld ? a->res_ld : N, ldm = a->mask_ld ? a->mask_ld : N;
  if (!epi_tma_ok(a->out, a->out_fp32, ldo, a->res, a->res_fp32, ldr, a->mask, ldm)) return 0;
  int bn, splits;
  plan_tile((int)M, N, k_iters, true, a->force_bn, a->force_splits, &bn, &splits);
  if (splits <= 1) return 0;
  const int tiles = ceil_div((int)M, BM) * ceil_div(N, bn);
  return (long long)tiles * BM * bn * 4 + (long long)tiles * 4;
}
