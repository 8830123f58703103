// conv_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a, fp32-grade via the 3xTF32 split:
//     x*w ~= hi(x)*hi(w) + lo(x)*hi(w) + hi(x)*lo(w),   hi = cvt.rna.tf32(.), lo = . - hi  (exact in fp32)
// with fp32 accumulation in TMEM (a separate correction accumulator holds the two small terms).
//
// GEMM view: M = N*H*W output pixels (tile of 128 = one TMA box of the NHWC activation), N = output channels, K = taps x input
// channels, one pipeline stage = (one tap, 32 channels) = a 128-byte swizzle row.  The activation box is im2col-free: the tap shift is
// a coordinate offset, image borders are TMA out-of-bounds zero fill, stride 2 is a TMA element stride.
//
// Kernels (launch_tc picks by output width):
//   conv_tc_ps_kernel     fprop / dgrad for > 64 output channels: persistent (1 CTA per SM loops over tiles), A hi/lo in shared memory,
//                         3 x 64 KB stages, two TMEM accumulator sets (the epilogue of tile i overlaps the main loop of tile i+1), and
//                         a_hi x [b_hi | b_lo] issued as ONE N=256 instruction into [main | correction]
//   conv_tc_ts_kernel<64> <= 64 output channels: one tile per CTA, the split A operand goes through TENSOR MEMORY (tcgen05.st)
//   wgrad_tc_kernel       weight gradient: dY^T through TMEM, X split in shared memory, both MN-major, split-K over pixels
//   pack_tc / split_tf32 / transpose_batched helpers; dp_gemm_nt_tc runs the attention GEMMs on the persistent kernel.
// The experimental variants measured in round 1 (SS one-tile, cluster multicast, decoupled rings, two-issuer, persistent TS, 16-float
// stages, the clock64() stage tracer; profiles/r01_experiments.md) live on the git tag `lab-kernels-r01`, the cta_group::2 CTA-pair
// kernel of round 2 (correct, 1.4x SLOWER: the kernel is shared-memory-bandwidth bound, profiles/r02_experiments.md) on
// `lab-pair-kernel-r02` — neither is in the product library.
#include <cuda.h>
#include <mutex>
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 32, NTHREADS = 192;
constexpr int A_BYTES = BM * BK * 4;  // 16 KB

struct TcParams {
  int Nimg, H, W;
  int Nout;            // GEMM N (valid output channels)
  int kchunks;         // ceil(Kg / 32)
  int bw, bh, bn, tiles_w, tiles_h;
  float* y; long long ldy;
  const float* bias;
  const float* rowadd; long long ld_rowadd;
  const float* residual; long long ld_res;
  int accumulate;
  int vec4;            // all epilogue pointers / strides are 16-byte aligned
  // generalised tap table (stride-2 dgrad runs as 4 parity classes with 1/2/2/4 taps each) and output pixel mapping
  int ntaps;
  signed char dh[9], dw[9], wt[9];
  int os, oa, ob, Ho, Wo;   // output pixel = (p*os + oa, q*os + ob) on an [Ho][Wo] grid
  int in_stride;            // strided fprop: input pixel = in_stride * output pixel + tap offset (the A map traverses W, H with that stride)
  float alpha;              // epilogue scale of the accumulator (attention logits); 1 for convolutions
  int b_from_img;           // batched GEMM: the B tile index is the tile's image (bn == 1) instead of a filter tap
  // split-K (persistent kernel, launches with fewer tiles than half the SMs: the 4x4 / 8x8 / 16x16 levels): work item = (tile, K split);
  // a split walks `it_per_split` pipeline stages of the tile and writes its raw accumulator to ws[split][row][channel]
  // (row = tile_m * 128 + TMEM lane, pitch ws_ld); splitk_epilogue_kernel sums the splits in fixed order and applies the epilogue
  int ksplit, it_per_split;
  float* ws; long long ws_split_stride; int ws_ld;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major, 128B-swizzled operand tile descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 | SBO(1024B>>4)<<32
// | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// warp-converged single-lane election (elect.sync): lets ptxas keep descriptors / barrier addresses in UNIFORM registers and emit
// straight-line UTCHMMA / UTMALDG; a plain `if (lane == 0)` makes it wrap every such instruction in an ELECT / BRA.U.ANY loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
// A whole (converged) warp waits on a barrier (every lane polls: hardware-suspended try_wait; lane-0-only polling measured 14 % slower).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}


// ------------------------------------------------------------------------------------------------ TS kernel (<= 64 output channels)
// The split A operand goes registers -> TENSOR MEMORY (tcgen05.st) and the MMA reads A from TMEM (tcgen05.mma [d], [a_tmem], b_desc):
// the shared-memory pipe only carries the TMA writes, one read of the raw A tile and the B operand reads.
// TMEM map (512 columns): [0,BN) main accumulator | [BN,2BN) correction accumulator | 2BN + 64*s: A_hi(32) A_lo(32) of stage s.
constexpr int STAGES_TS = 4;
constexpr int PF_DIST = 8;   // L2 prefetch distance (stages) for the activation boxes
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}
template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_ts_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                  const __grid_constant__ CUtensorMap mapBl, const TcParams p) {
  constexpr int B_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  constexpr uint32_t A_COL0 = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to;
  const uint32_t sbase = raw + pad_to;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES_TS * STAGE_BYTES);
  const uint32_t bar0 = sbase + STAGES_TS * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto conv_bar = [&](int s) { return bar0 + 8u * (STAGES_TS + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES_TS + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (3 * STAGES_TS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES_TS + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES_TS; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int tile_m = blockIdx.x, nblk = blockIdx.y;
  const int tw = tile_m % p.tiles_w;
  const int th = (tile_m / p.tiles_w) % p.tiles_h;
  const int tn = tile_m / (p.tiles_w * p.tiles_h);
  const int q0 = tw * p.bw, p0 = th * p.bh, n0 = tn * p.bn;
  const int num_iters = p.ntaps * p.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
      for (int it = 0; it < num_iters; ++it) {
        const int s = it % STAGES_TS;
        const uint32_t ph = (uint32_t)(it / STAGES_TS) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
        const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
        const uint32_t st = sbase + s * STAGE_BYTES;
        tma_load_4d(st, &mapA, full_bar(s), kc * BK, q0 + p.dw[tap], p0 + p.dh[tap], n0);
        if (it + PF_DIST < num_iters) {   // pull the A box of a later stage into L2 while this one is in flight
          const int it2 = it + PF_DIST, tap2 = it2 / p.kchunks, kc2 = it2 - tap2 * p.kchunks;
          asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                       ::"l"(reinterpret_cast<uint64_t>(&mapA)), "r"(kc2 * BK), "r"(q0 + p.dw[tap2]), "r"(p0 + p.dh[tap2]), "r"(n0) : "memory");
        }
        const int tapb = p.wt[tap];
        tma_load_3d(st + A_BYTES, &mapBh, full_bar(s), kc * BK, nblk * BN, tapb);
        tma_load_3d(st + A_BYTES + B_BYTES, &mapBl, full_bar(s), kc * BK, nblk * BN, tapb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // UMMA N = valid channels of this N tile rounded up to 16 (pruned widths 96 / 192 / 179 do not pay for 128)
      const int n_valid = min(BN, p.Nout - nblk * BN);
      const uint32_t n_instr = (uint32_t)((n_valid + 15) & ~15);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_instr >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < num_iters; ++it) {
        const int s = it % STAGES_TS;
        const uint32_t ph = (uint32_t)(it / STAGES_TS) & 1u;
        mbar_wait(conv_bar(s), ph);     // A hi/lo of this stage are in TMEM (and, transitively, B has landed in smem)
        mbar_wait(full_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = sbase + s * STAGE_BYTES;
        const uint32_t a_t = tmem_base + A_COL0 + 64u * s;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t b_hi = umma_desc(st + A_BYTES + k * 32), b_lo = umma_desc(st + A_BYTES + B_BYTES + k * 32);
          const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
          umma_tf32_ts(tmem_base + BN, a_t + 32 + k * 8, b_hi, idesc, first);   // lo * hi
          umma_tf32_ts(tmem_base + BN, a_t + k * 8, b_lo, idesc, 1u);           // hi * lo
          umma_tf32_ts(tmem_base, a_t + k * 8, b_hi, idesc, first);             // hi * hi
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---- splitter: thread <-> tile row (TMEM lane).  Row r of the 128B-swizzled tile: 16-byte chunk j sits at (j ^ (r & 7)).
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < num_iters; ++it) {
      const int s = it % STAGES_TS;
      const uint32_t ph = (uint32_t)(it / STAGES_TS) & 1u;
      mbar_wait(full_bar(s), ph);
      const uint8_t* arow = smem + s * STAGE_BYTES + row * 128;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(arow + ((j ^ (row & 7)) << 4));
        const float h0 = tf32_rna(v.x), h1 = tf32_rna(v.y), h2 = tf32_rna(v.z), h3 = tf32_rna(v.w);
        hi[4 * j + 0] = __float_as_uint(h0); hi[4 * j + 1] = __float_as_uint(h1);
        hi[4 * j + 2] = __float_as_uint(h2); hi[4 * j + 3] = __float_as_uint(h3);
        lo[4 * j + 0] = __float_as_uint(v.x - h0); lo[4 * j + 1] = __float_as_uint(v.y - h1);
        lo[4 * j + 2] = __float_as_uint(v.z - h2); lo[4 * j + 3] = __float_as_uint(v.w - h3);
      }
      const uint32_t a_t = tmem_base + lane_addr + A_COL0 + 64u * s;
      tmem_st32(a_t, hi);
      tmem_st32(a_t + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(conv_bar(s));
    }
    // ---- epilogue
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int w_l = row % p.bw, h_l = (row / p.bw) % p.bh, n_l = row / (p.bw * p.bh);
    const int img = n0 + n_l;
    const bool row_ok = img < p.Nimg;
    const long long m = ((long long)img * p.Ho + ((p0 + h_l) * p.os + p.oa)) * p.Wo + ((q0 + w_l) * p.os + p.ob);
    float* yrow = p.y + m * p.ldy;
    const float* rrow = p.residual ? p.residual + m * p.ld_res : nullptr;
    const float* arow2 = p.rowadd ? p.rowadd + (long long)img * p.ld_rowadd : nullptr;
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      uint32_t v[32], u[32];
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(j * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
            "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
            "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
            "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
          : "r"(taddr + (uint32_t)BN));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row_ok) {
        const int c0 = nblk * BN + j * 32;
        if (p.vec4 && c0 + 32 <= p.Nout) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o = make_float4(__uint_as_float(v[i]) + __uint_as_float(u[i]), __uint_as_float(v[i + 1]) + __uint_as_float(u[i + 1]),
                                   __uint_as_float(v[i + 2]) + __uint_as_float(u[i + 2]), __uint_as_float(v[i + 3]) + __uint_as_float(u[i + 3]));
            if (p.bias) { float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            if (arow2) { float4 t = __ldg(reinterpret_cast<const float4*>(arow2 + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            if (rrow) { float4 t = __ldg(reinterpret_cast<const float4*>(rrow + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            float4* dst = reinterpret_cast<float4*>(yrow + c0 + i);
            if (p.accumulate) { float4 t = *dst; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            *dst = o;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = c0 + i;
            if (c < p.Nout) {
              float o = __uint_as_float(v[i]) + __uint_as_float(u[i]);
              if (p.bias) o += __ldg(p.bias + c);
              if (arow2) o += __ldg(arow2 + c);
              if (rrow) o += __ldg(rrow + c);
              if (p.accumulate) o += yrow[c];
              yrow[c] = o;
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------ persistent variant
// One CTA per SM loops over output tiles (static stride), 10 warps: TMA producer | MMA issuer | 4 splitter warps |
// 4 epilogue warps.  Two accumulator sets in TMEM (2 x [main 128 | correction 128] = 512 columns) let the epilogue of tile
// i (TMEM -> registers -> global, ~20-50 % of a short-K tile) overlap the main loop of tile i+1; barrier init, TMEM
// allocation and descriptor prefetch are paid once per SM instead of once per tile.  A operand hi/lo in shared memory
// (SS mode; the TS variant needs the TMEM columns the second accumulator set occupies).
constexpr int PS_THREADS = 320, PS_STAGES = 3;

__global__ void __launch_bounds__(PS_THREADS, 1)
conv_tc_ps_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                  const __grid_constant__ CUtensorMap mapBl, const TcParams p, const int tiles_m, const int total_tiles) {
  const int total_work = total_tiles * p.ksplit;     // work item wi = split * total_tiles + tile (the splits of one tile run on different SMs)
  constexpr int BN = 128;
  constexpr int B_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to;
  const uint32_t sbase = raw + pad_to;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PS_STAGES * STAGE_BYTES);
  const uint32_t bar0 = sbase + PS_STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto conv_bar = [&](int s) { return bar0 + 8u * (PS_STAGES + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * PS_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (3 * PS_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (3 * PS_STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * PS_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PS_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int iters_per_tile = p.ntaps * p.kchunks;

  auto tile_coords = [&](int tile, int& q0, int& p0, int& n0, int& nblk) {
    nblk = tile / tiles_m;
    const int tile_m = tile - nblk * tiles_m;
    const int tw = tile_m % p.tiles_w;
    const int th = (tile_m / p.tiles_w) % p.tiles_h;
    const int tn = tile_m / (p.tiles_w * p.tiles_h);
    q0 = tw * p.bw; p0 = th * p.bh; n0 = tn * p.bn;
  };

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
      uint32_t g = 0;
      for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x) {
        const int tile = wi % total_tiles, it0 = (wi / total_tiles) * p.it_per_split, it1 = min(iters_per_tile, it0 + p.it_per_split);
        int q0, p0, n0, nblk;
        tile_coords(tile, q0, p0, n0, nblk);
        for (int it = it0; it < it1; ++it, ++g) {
          const int s = g % PS_STAGES;
          const uint32_t ph = (g / PS_STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
          const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
          const uint32_t st = sbase + s * STAGE_BYTES;
          tma_load_4d(st, &mapA, full_bar(s), kc * BK, q0 * p.in_stride + p.dw[tap], p0 * p.in_stride + p.dh[tap], n0);
          const int tapb = p.b_from_img ? n0 : p.wt[tap];
          tma_load_3d(st + 2 * A_BYTES, &mapBh, full_bar(s), kc * BK, nblk * BN, tapb);
          tma_load_3d(st + 2 * A_BYTES + B_BYTES, &mapBl, full_bar(s), kc * BK, nblk * BN, tapb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t g = 0, tl = 0;
      for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x, ++tl) {
        const int tile = wi % total_tiles, it0 = (wi / total_tiles) * p.it_per_split, it1 = min(iters_per_tile, it0 + p.it_per_split);
        const int nblk = tile / tiles_m;
        const int n_valid = min(BN, p.Nout - nblk * BN);
        const uint32_t n_instr = (uint32_t)((n_valid + 15) & ~15);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_instr >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t b = tl & 1u, use = tl >> 1;
        mbar_wait(tempty_bar(b), (use & 1u) ^ 1u);          // epilogue has drained this accumulator set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + b * 256u;
        for (int it = it0; it < it1; ++it, ++g) {
          const int s = g % PS_STAGES;
          const uint32_t ph = (g / PS_STAGES) & 1u;
          mbar_wait(conv_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = sbase + s * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t a_hi = umma_desc(st + k * 32), a_lo = umma_desc(st + A_BYTES + k * 32);
            const uint64_t b_hi = umma_desc(st + 2 * A_BYTES + k * 32);
            const uint32_t first = (it > it0 || k > 0) ? 1u : 0u;
            // a_hi x [b_hi | b_lo] -> [main | correction] as ONE N=256 instruction (the two B tiles are adjacent in shared memory):
            // 8 instead of 12 instructions per stage and 5/6 of the operand reads
            umma_tf32(acc, a_hi, b_hi, idesc256, first);
            umma_tf32(acc + 128, a_lo, b_hi, idesc, 1u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(b));
      }
    }
  } else if (warp < 6) {
    // ---- splitter warps 2..5
    const int ct = threadIdx.x - 64;
    uint32_t g = 0;
    for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x) {
      const int it0 = (wi / total_tiles) * p.it_per_split, it1 = min(iters_per_tile, it0 + p.it_per_split);
      for (int it = it0; it < it1; ++it, ++g) {
        const int s = g % PS_STAGES;
        const uint32_t ph = (g / PS_STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        float4* A = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
        float4* Al = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = ct + 128 * i;
          float4 v = A[idx], h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          A[idx] = h;
          Al[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(conv_bar(s));
      }
    }
  } else {
    // ---- epilogue warps 6..9 (TMEM lane quarter = warp & 3)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int w_l = row % p.bw, h_l = (row / p.bw) % p.bh, n_l = row / (p.bw * p.bh);
    uint32_t tl = 0;
    for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x, ++tl) {
      const int tile = wi % total_tiles;
      int q0, p0, n0, nblk;
      tile_coords(tile, q0, p0, n0, nblk);
      const uint32_t b = tl & 1u, use = tl >> 1;
      mbar_wait(tfull_bar(b), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int img = n0 + n_l;
      const bool row_ok = img < p.Nimg;
      const long long m = ((long long)img * p.Ho + ((p0 + h_l) * p.os + p.oa)) * p.Wo + ((q0 + w_l) * p.os + p.ob);
      float* yrow = p.y + m * p.ldy;
      const float* rrow = p.residual ? p.residual + m * p.ld_res : nullptr;
      const float* arow2 = p.rowadd ? p.rowadd + (long long)img * p.ld_rowadd : nullptr;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t v[32], u[32];
        const uint32_t taddr = tmem_base + lane_addr + b * 256u + (uint32_t)(j * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
              "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
              "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
              "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
            : "r"(taddr + 128u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (j == BN / 32 - 1) {   // accumulators are in registers: hand the TMEM set back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(tempty_bar(b));
        }
        if (p.ksplit > 1) {   // split-K: raw partial sums into this split's slab of the (padded) workspace; splitk_epilogue_kernel finishes
          float* wrow = p.ws + (long long)(wi / total_tiles) * p.ws_split_stride + ((long long)(tile - nblk * tiles_m) * BM + row) * p.ws_ld + nblk * BN + j * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(wrow + i) = make_float4(__uint_as_float(v[i]) + __uint_as_float(u[i]), __uint_as_float(v[i + 1]) + __uint_as_float(u[i + 1]),
                                                               __uint_as_float(v[i + 2]) + __uint_as_float(u[i + 2]), __uint_as_float(v[i + 3]) + __uint_as_float(u[i + 3]));
        } else if (row_ok) {
          const int c0 = nblk * BN + j * 32;
          if (p.vec4 && c0 + 32 <= p.Nout) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 o = make_float4(p.alpha * (__uint_as_float(v[i]) + __uint_as_float(u[i])), p.alpha * (__uint_as_float(v[i + 1]) + __uint_as_float(u[i + 1])),
                                     p.alpha * (__uint_as_float(v[i + 2]) + __uint_as_float(u[i + 2])), p.alpha * (__uint_as_float(v[i + 3]) + __uint_as_float(u[i + 3])));
              if (p.bias) { float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
              if (arow2) { float4 t = __ldg(reinterpret_cast<const float4*>(arow2 + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
              if (rrow) { float4 t = __ldg(reinterpret_cast<const float4*>(rrow + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
              float4* dst = reinterpret_cast<float4*>(yrow + c0 + i);
              if (p.accumulate) { float4 t = *dst; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
              *dst = o;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = c0 + i;
              if (c < p.Nout) {
                float o = p.alpha * (__uint_as_float(v[i]) + __uint_as_float(u[i]));
                if (p.bias) o += __ldg(p.bias + c);
                if (arow2) o += __ldg(arow2 + c);
                if (rrow) o += __ldg(rrow + c);
                if (p.accumulate) o += yrow[c];
                yrow[c] = o;
              }
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Sums the K splits of conv_tc_ps_kernel in fixed order (deterministic) and applies its epilogue: alpha, bias, per-image row, residual,
// accumulate, the (strided) output pixel mapping.  One thread per (GEMM row, 4 channels).
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(const TcParams p, const int tiles_m) {
  const int c4 = (p.Nout + 3) >> 2;
  const long long total = (long long)tiles_m * BM * c4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long grow = i / c4;
    const int c = (int)(i - grow * c4) << 2;
    const int tile_m = (int)(grow / BM), row = (int)(grow - (long long)tile_m * BM);
    const int tw = tile_m % p.tiles_w, th = (tile_m / p.tiles_w) % p.tiles_h, tn = tile_m / (p.tiles_w * p.tiles_h);
    const int w_l = row % p.bw, h_l = (row / p.bw) % p.bh, n_l = row / (p.bw * p.bh);
    const int img = tn * p.bn + n_l;
    if (img >= p.Nimg) continue;
    const float* src = p.ws + grow * p.ws_ld + c;
    float4 acc = *reinterpret_cast<const float4*>(src);
    for (int ks = 1; ks < p.ksplit; ++ks) {
      const float4 t = *reinterpret_cast<const float4*>(src + ks * p.ws_split_stride);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float o[4] = {p.alpha * acc.x, p.alpha * acc.y, p.alpha * acc.z, p.alpha * acc.w};
    const long long m = ((long long)img * p.Ho + ((th * p.bh + h_l) * p.os + p.oa)) * p.Wo + ((tw * p.bw + w_l) * p.os + p.ob);
    float* yrow = p.y + m * p.ldy;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int cc = c + e;
      if (cc < p.Nout) {
        float v = o[e];
        if (p.bias) v += __ldg(p.bias + cc);
        if (p.rowadd) v += __ldg(p.rowadd + (long long)img * p.ld_rowadd + cc);
        if (p.residual) v += __ldg(p.residual + m * p.ld_res + cc);
        if (p.accumulate) v += yrow[cc];
        yrow[cc] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// dW[k][tap][c] = sum_pix dy[pix][k] * x[pix @ tap][c]  as GEMM  M = k (128), N = c (128), K = pixels.
// Both operands are activations stored pixel-major / channel-contiguous, i.e. "MN-major" for the tensor core:
// a stage = 32 pixels (one TMA box of the pixel grid) x 128 channels = 4 swizzle blocks [32 px][128 B] per operand
// (MN-block stride LBO = 4 KB, 8-pixel K-group stride SBO = 1 KB).  Both tiles are split hi/lo in shared memory.
// grid = (k tiles * c tiles * taps, splits): split z covers pixel chunks [z*cps, (z+1)*cps) and writes its partial
// tile to workspace[z][k][tap*C + c]; dp_conv2d_wgrad_reduce sums splits in fixed order (deterministic).
struct WgParams {
  int Nimg, H, W, C, K;
  int R, S, pad;
  int bw, bh, bn, tiles_w, tiles_h;   // 32-pixel box
  int total_chunks, chunks_per_split;
  int c_tiles;
  float* ws;
  int in_stride;             // x pixel = in_stride * dy pixel + tap offset
};
constexpr int WG_KPIX = 32, WG_T = 128 * WG_KPIX * 4;   // one operand tile = 16 KB

// MN-major TF32 operands must use the SWIZZLE_128B_BASE32B layout (cute: Layout_MN_SW128_32B_Atom, "the only available
// smem layout for mn-major tf32"): atoms of 4 K-rows x 128 B with the four 32-byte chunks of a row XOR-ed by (row & 3);
// TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = 4096 B between 32-channel blocks, SBO = 512 B between
// 4-pixel K-groups, layout_type = 1.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (256ull << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}

constexpr int WG_THREADS = 224;   // warps: 0 TMA | 1, 6 MMA issuers (alternate stages, see conv_tc_ps_kernel) | 2-5 splitters + epilogue
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapDy, const __grid_constant__ CUtensorMap mapX, const WgParams p) {
  // stage smem: dy raw (16 KB, read once by the splitter) | x (hi in place, 16 KB) | x_lo (16 KB)
  // The A operand (dY^T: lane = out-channel, column = pixel) is built in TENSOR MEMORY: thread <-> out-channel reads its
  // channel across the 32 pixel rows of the (32B-atom swizzled) tile — one conflict-free 128 B row per warp instruction —
  // splits hi/lo and writes 2 x 32 columns with tcgen05.st; the MMA then runs in TS mode (A from TMEM, B = x MN-major smem).
  // TMEM: [0,128) main acc | [128,256) correction acc | 256 + 64*s: A_hi(32) A_lo(32) of stage s (4 stages -> 512 columns).
  constexpr int WSTAGES = 4;
  constexpr int STAGE_BYTES = 3 * WG_T;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to;
  const uint32_t sbase = raw + pad_to;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WSTAGES * STAGE_BYTES);
  const uint32_t bar0 = sbase + WSTAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto conv_bar = [&](int s) { return bar0 + 8u * (WSTAGES + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * WSTAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (3 * WSTAGES);
  auto iss_bar = [&](int s) { return bar0 + 8u * (3 * WSTAGES + 1 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * WSTAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); mbar_init(iss_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int T = p.R * p.S;
  int tile = blockIdx.x;
  const int tap = tile % T; tile /= T;
  const int ct = tile % p.c_tiles;
  const int kt = tile / p.c_tiles;
  const int r = tap / p.S, sx = tap - r * p.S;
  const int chunk0 = blockIdx.y * p.chunks_per_split;
  const int chunk1 = min(p.total_chunks, chunk0 + p.chunks_per_split);
  const int num_iters = max(0, chunk1 - chunk0);   // 0 for a trailing empty split: its workspace tile is zero-filled

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapDy)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapX)) : "memory");
      for (int it = 0; it < num_iters; ++it) {
        const int s = it % WSTAGES;
        const uint32_t ph = (uint32_t)(it / WSTAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), 2 * WG_T);
        const int chunk = chunk0 + it;
        const int tw = chunk % p.tiles_w;
        const int th = (chunk / p.tiles_w) % p.tiles_h;
        const int tn = chunk / (p.tiles_w * p.tiles_h);
        const int q0 = tw * p.bw, p0 = th * p.bh, n0 = tn * p.bn;
        const uint32_t st = sbase + s * STAGE_BYTES;
#pragma unroll
        for (int b = 0; b < 4; ++b) {   // 4 blocks of 32 channels = 128 channels per operand
          tma_load_4d(st + b * 4096, &mapDy, full_bar(s), kt * 128 + b * 32, q0, p0, n0);
          tma_load_4d(st + WG_T + b * 4096, &mapX, full_bar(s), ct * 128 + b * 32, q0 * p.in_stride + sx - p.pad, p0 * p.in_stride + r - p.pad, n0);
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // two issuer warps on alternate stages, warp-converged with one elected lane
    const uint32_t mw = (warp == 1) ? 0u : 1u;
    const uint32_t two = 1u;   // two issuer warps on alternate stages: a lone issuer cannot run ahead of the tensor queue (profiles/r01_experiments.md)
    // B MN-major (bit 16); A comes from TMEM
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (num_iters == 0 && mw == 0) {   // empty split: release the epilogue (it writes zeros)
      if (elect_one()) umma_commit(tmem_full_bar);
      __syncwarp();
    }
    if (two || mw == 0) {
      for (int it = 0; it < num_iters; ++it) {
        if (two && ((uint32_t)it & 1u) != mw) continue;
        const int s = it % WSTAGES;
        const uint32_t ph = (uint32_t)(it / WSTAGES) & 1u;
        mbar_wait_warp(conv_bar(s), ph);
        mbar_wait_warp(full_bar(s), ph);
        if (two && it > 0) mbar_wait_warp(iss_bar((it - 1) % WSTAGES), (uint32_t)((it - 1) / WSTAGES) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = sbase + s * STAGE_BYTES;
        const uint32_t a_t = tmem_base + 256u + 64u * s;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < WG_KPIX / 8; ++k) {
            const uint64_t b_hi = umma_desc_mn(st + WG_T + k * 1024);
            const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
            // x_hi | x_lo are adjacent 4 x 32-channel block groups: a_hi x [x_hi | x_lo] -> [main | correction] in one N=256 instruction
            umma_tf32_ts(tmem_base, a_t + k * 8, b_hi, idesc256, first);
            umma_tf32_ts(tmem_base + 128, a_t + 32 + k * 8, b_hi, idesc, 1u);
          }
          umma_commit(empty_bar(s));
          if (it == num_iters - 1) umma_commit(tmem_full_bar);
          if (two) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(iss_bar(s));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    const int tid = threadIdx.x - 64;
    const int q = warp & 3;                   // TMEM lane quarter == 32-channel block of dy
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < num_iters; ++it) {
      const int s = it % WSTAGES;
      const uint32_t ph = (uint32_t)(it / WSTAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      // (1) dy^T -> TMEM.  Block q holds channels 32q..32q+31 as rows [pix][32 ch]; SWIZZLE_128B_ATOM_32B: the 32-byte
      //     chunk j of row `pix` sits at chunk position j ^ (pix & 3).
      {
        const uint8_t* blk = smem + s * STAGE_BYTES + q * 4096;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int pix = 0; pix < 32; ++pix) {
          const int chunk = (lane >> 3) ^ (pix & 3);
          const float v = *reinterpret_cast<const float*>(blk + pix * 128 + chunk * 32 + (lane & 7) * 4);
          const float h = tf32_rna(v);
          hi[pix] = __float_as_uint(h);
          lo[pix] = __float_as_uint(v - h);
        }
        const uint32_t a_t = tmem_base + lane_addr + 256u + 64u * s;
        tmem_st32(a_t, hi);
        tmem_st32(a_t + 32, lo);
      }
      // (2) x tile: hi in place, lo to the side (elementwise, layout agnostic)
      {
        float4* A = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + WG_T);
        float4* Al = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + 2 * WG_T);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = tid + 128 * i;
          float4 v = A[idx], h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          A[idx] = h;
          Al[idx] = l;
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(conv_bar(s));
    }
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = q * 32 + lane;            // k_out within the tile
    const int kout = kt * 128 + row;
    const long long TC_ = (long long)T * p.C;
    float* wrow = p.ws + ((long long)blockIdx.y * p.K + kout) * TC_ + (long long)tap * p.C;
    if (num_iters == 0) {                     // nothing was accumulated (TMEM holds garbage): this split contributes zeros
      if (kout < p.K)
        for (int c = ct * 128; c < min(p.C, ct * 128 + 128); ++c) wrow[c] = 0.f;
    } else
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      uint32_t v[32], u[32];
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(j * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
            "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
            "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
            "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
          : "r"(taddr + 128u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (kout < p.K) {
        const int c0 = ct * 128 + j * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.C) wrow[c0 + i] = __uint_as_float(v[i]) + __uint_as_float(u[i]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_tc_state = -1;  // -1 unknown, 0 unavailable, 1 ok
int g_num_sms = 148;
std::mutex g_tc_mutex;
constexpr int PS_SMEM = PS_STAGES * (2 * A_BYTES + 2 * 128 * BK * 4) + 2048;
constexpr int TS64_SMEM = STAGES_TS * (A_BYTES + 2 * 64 * BK * 4) + 2048;
constexpr int WG_SMEM = 4 * 3 * WG_T + 2048;

// Row length of the packed TF32 weight tiles (dp_pack_conv_weight_tc): rows longer than 32 floats are zero-padded to a multiple of 32
// floats (128 B) so that every 32-float TMA box row is exactly one aligned 128-byte line; short rows to a multiple of 4 (the TMA
// 16-byte stride rule).  With 16-byte padding only, pruned widths (90 / 179 input channels) ran 20-25 % slower than the next
// multiple of 32 (scripts/time_conv_shapes.py: 90 -> 90 3x3 @32x32 167 us vs 134 us).
static int wrow(int c) { return c > 32 ? ((c + 31) & ~31) : ((c + 3) & ~3); }

int tc_init() {
  std::lock_guard<std::mutex> lk(g_tc_mutex);
  if (g_tc_state >= 0) return g_tc_state;
  g_tc_state = 0;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) { (void)cudaGetLastError(); return 0; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) { (void)cudaGetLastError(); return 0; }
  g_encode = (EncodeTiledFn)fn;
  bool ok = cudaFuncSetAttribute(conv_tc_ts_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS64_SMEM) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM) == cudaSuccess;
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (!ok) { (void)cudaGetLastError(); return 0; }
  (void)cudaGetLastError();
  g_tc_state = 1;
  return 1;
}

// pix_stride > 1 (strided convolution): dims 1 and 2 (W, H) are traversed with that element stride; the caller passes the box
// extents in traversed elements (box = loaded pixels x pix_stride)
bool make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int pix_stride = 1) {
  cuuint32_t estr[5] = {1, (cuuint32_t)pix_stride, (cuuint32_t)pix_stride, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 128-pixel box of an [N][H][W] grid
bool pick_box(int N, int H, int W, int& bw, int& bh, int& bn) {
  if (W >= BM) {
    if (W % BM) return false;
    bw = BM; bh = 1; bn = 1; return true;
  }
  if (BM % W) return false;
  bw = W;
  int rem = BM / W;
  if (H >= rem) {
    if (H % rem) return false;
    bh = rem; bn = 1; return true;
  }
  if (rem % H) return false;
  bh = H; bn = rem / H;
  return true;
}

struct TapTable { int n; signed char dh[9], dw[9], wt[9]; };
// K splits of a persistent-kernel launch with `tiles` output tiles of `iters` pipeline stages each: enough work items to fill the SMs,
// at least 4 stages per split, none when the tiles already cover half the machine
static int pick_ksplit(int tiles, int iters, int& it_per_split) {
  it_per_split = iters;
  if (tiles * 2 > g_num_sms || iters < 8) return 1;
  int ks = g_num_sms / tiles;
  if (ks > iters / 4) ks = iters / 4;
  if (ks > 16) ks = 16;
  if (ks < 2) return 1;
  it_per_split = (iters + ks - 1) / ks;
  return (iters + it_per_split - 1) / it_per_split;      // no empty split
}

// Shared launcher.  act: [Nimg][H][W][Kg] view (ld_act) = A operand on whose pixel grid the M tiles live; w_hi/w_lo: [T][Nout][Kg];
// out: [Nimg][Ho][Wo][Nout] view, output pixel = (p*os+oa, q*os+ob).
// ws: optional split-K workspace (dp_conv_splitk_workspace_floats floats); *ws_need != nullptr: only report the floats a split launch needs
int launch_tc(const float* act, long long ld_act, int Nimg, int H, int W, int Kg, const float* w_hi, const float* w_lo, int Nout,
              int T, const TapTable& taps, int os, int oa, int ob, int Ho, int Wo, float* out, long long ld_out, const float* bias,
              const float* rowadd, long long ld_rowadd, const float* residual, long long ld_res, int accumulate, cudaStream_t st,
              float alpha = 1.0f, int b_from_img = 0, int in_stride = 1, int ldb = -1, float* ws = nullptr, long long* ws_need = nullptr) {
  if (ldb < 0) ldb = wrow(Kg);   // packed conv weights; batched GEMM callers pass their own row pitch
  if (!tc_init()) return DP_ERR_UNSUPPORTED;
  if (!ws_need && (!w_hi || !w_lo)) return DP_ERR_UNSUPPORTED;
  if (!ws_need && (ld_act % 4 || ((uintptr_t)act & 15) || ((uintptr_t)w_hi & 15) || ((uintptr_t)w_lo & 15))) return DP_ERR_UNSUPPORTED;
  int bw, bh, bn;
  if (!pick_box(Nimg, H, W, bw, bh, bn)) return DP_ERR_UNSUPPORTED;
  const int BN = (Nout <= 64) ? 64 : 128;
  if (ws_need) {     // geometry-only query
    *ws_need = 0;
    if (BN != 128 || b_from_img) return DP_OK;
    const int tiles_m = (W / bw) * (H / bh) * ((Nimg + bn - 1) / bn), n_tiles = (Nout + 127) / 128;
    int ips;
    const int ks = pick_ksplit(tiles_m * n_tiles, taps.n * ((Kg + BK - 1) / BK), ips);
    if (ks > 1) *ws_need = (long long)ks * tiles_m * BM * n_tiles * 128;
    return DP_OK;
  }
  if ((in_stride != 1 || alpha != 1.0f || b_from_img) && BN != 128) return DP_ERR_UNSUPPORTED;   // only the persistent kernel scales the tile origin / applies alpha / image-indexed B
  if (b_from_img && bn != 1) return DP_ERR_UNSUPPORTED;
  CUtensorMap mA, mBh, mBl;
  {
    // strided fprop: the M tiles live on the OUTPUT grid [H][W]; the activation is [H*in_stride][W*in_stride] and the box picks every
    // in_stride-th pixel (TMA element strides), so a tile is still one 128-pixel box
    const cuuint64_t Hin = (cuuint64_t)H * in_stride, Win = (cuuint64_t)W * in_stride;
    cuuint64_t dims[4] = {(cuuint64_t)Kg, Win, Hin, (cuuint64_t)Nimg};
    cuuint64_t str[3] = {(cuuint64_t)ld_act * 4, Win * ld_act * 4, Hin * Win * ld_act * 4};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(bw * in_stride), (cuuint32_t)(bh * in_stride), (cuuint32_t)bn};
    if (box[1] > 256 || box[2] > 256) return DP_ERR_UNSUPPORTED;
    if (!make_map(&mA, act, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, in_stride)) return DP_ERR_UNSUPPORTED;
  }
  {
    const cuuint64_t Kg4 = (cuuint64_t)ldb;
    cuuint64_t dims[3] = {Kg4, (cuuint64_t)Nout, (cuuint64_t)T};
    cuuint64_t str[2] = {Kg4 * 4, (cuuint64_t)Nout * Kg4 * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
    if (!make_map(&mBh, w_hi, 3, dims, str, box) || !make_map(&mBl, w_lo, 3, dims, str, box)) return DP_ERR_UNSUPPORTED;
  }
  TcParams p{};
  p.Nimg = Nimg; p.H = H; p.W = W; p.Nout = Nout;
  p.ntaps = taps.n;
  for (int i = 0; i < 9; ++i) { p.dh[i] = taps.dh[i]; p.dw[i] = taps.dw[i]; p.wt[i] = taps.wt[i]; }
  p.os = os; p.oa = oa; p.ob = ob; p.Ho = Ho; p.Wo = Wo;
  p.alpha = alpha; p.b_from_img = b_from_img; p.in_stride = in_stride;
  p.kchunks = (Kg + BK - 1) / BK;
  p.bw = bw; p.bh = bh; p.bn = bn; p.tiles_w = W / bw; p.tiles_h = H / bh;
  p.y = out; p.ldy = ld_out; p.bias = bias; p.rowadd = rowadd; p.ld_rowadd = ld_rowadd; p.residual = residual; p.ld_res = ld_res;
  p.accumulate = accumulate;
  auto al16 = [](const void* q, long long ld) { return q == nullptr || ((((uintptr_t)q) & 15) == 0 && (ld % 4) == 0); };
  p.vec4 = (al16(out, ld_out) && al16(bias, 0) && al16(rowadd, ld_rowadd) && al16(residual, ld_res)) ? 1 : 0;
  const int tiles_n = (Nimg + bn - 1) / bn;
  dim3 grid((unsigned)(p.tiles_w * p.tiles_h * tiles_n), (unsigned)((Nout + BN - 1) / BN));
  p.ksplit = 1; p.it_per_split = p.ntaps * p.kchunks;
  if (BN == 128) {
    const int tiles_m = (int)grid.x, total = (int)(grid.x * grid.y);
    if (ws && !b_from_img) {
      p.ksplit = pick_ksplit(total, p.ntaps * p.kchunks, p.it_per_split);
      p.ws = ws; p.ws_ld = (int)grid.y * 128; p.ws_split_stride = (long long)tiles_m * BM * p.ws_ld;
    }
    const int work = total * p.ksplit;
    const int ctas = work < g_num_sms ? work : g_num_sms;
    conv_tc_ps_kernel<<<ctas, PS_THREADS, PS_SMEM, st>>>(mA, mBh, mBl, p, tiles_m, total);
    if (p.ksplit > 1) {
      int rc = dp_check_launch();
      if (rc) return rc;
      const long long items = (long long)tiles_m * BM * ((Nout + 3) / 4);
      long long blocks = (items + 255) / 256;
      if (blocks > g_num_sms * 8) blocks = g_num_sms * 8;
      splitk_epilogue_kernel<<<(int)blocks, 256, 0, st>>>(p, tiles_m);
    }
  } else {
    conv_tc_ts_kernel<64><<<grid, NTHREADS, TS64_SMEM, st>>>(mA, mBh, mBl, p);
  }
  return dp_check_launch();
}

__global__ void pack_tc_kernel(const float* __restrict__ w, int K, int C, int RS, int C4, int K4, float* __restrict__ kc_hi,
                               float* __restrict__ kc_lo, float* __restrict__ ck_hi, float* __restrict__ ck_lo) {
  // rows are zero-padded to C4 = dp_tc_weight_row(C), K4 = dp_tc_weight_row(K): kc [RS][K][C4], ck [RS][C][K4]
  const long long na = (long long)RS * K * C4, nb = (long long)RS * C * K4;
  const long long total = na > nb ? na : nb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < na && (kc_hi || kc_lo)) {
      int c = (int)(i % C4); long long t = i / C4; int k = (int)(t % K), tap = (int)(t / K);
      float v = c < C ? w[((long long)k * C + c) * RS + tap] : 0.f, h = tf32_rna(v);
      if (kc_hi) kc_hi[i] = h;
      if (kc_lo) kc_lo[i] = v - h;
    }
    if (i < nb && (ck_hi || ck_lo)) {
      int k = (int)(i % K4); long long t = i / K4; int c = (int)(t % C), tap = (int)(t / C);
      float v = k < K ? w[((long long)k * C + c) * RS + tap] : 0.f, h = tf32_rna(v);
      if (ck_hi) ck_hi[i] = h;
      if (ck_lo) ck_lo[i] = v - h;
    }
  }
}
__global__ void split_tf32_kernel(const float* __restrict__ x, long long ld, long long bs, int rows, int cols, int transpose,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  // 32x32 tile through shared memory so both the read (along cols) and the transposed write (along rows) are coalesced
  __shared__ float t[32][33];
  const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* xb = x + (long long)b * bs;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? xb[(long long)r * ld + c] : 0.f;
  }
  __syncthreads();
  if (!transpose) {
    const int cols4 = (cols + 3) & ~3;
    for (int i = threadIdx.y; i < 32; i += 8) {
      int r = r0 + i, c = c0 + threadIdx.x;
      if (r < rows && c < cols4) {
        float v = t[i][threadIdx.x], h = tf32_rna(v);
        long long o = ((long long)b * rows + r) * cols4 + c;
        hi[o] = h; lo[o] = v - h;
      }
    }
  } else {
    const int rows4 = (rows + 3) & ~3;
    for (int i = threadIdx.y; i < 32; i += 8) {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (c < cols && r < rows4) {
        float v = t[threadIdx.x][i], h = tf32_rna(v);
        long long o = ((long long)b * cols + c) * rows4 + r;
        hi[o] = h; lo[o] = v - h;
      }
    }
  }
}
__global__ void transpose_batched_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* ib = in + (long long)b * rows * cols;
  float* ob = out + (long long)b * rows * cols;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) t[i][threadIdx.x] = ib[(long long)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) ob[(long long)c * rows + r] = t[threadIdx.x][i];
  }
}
}  // namespace

extern "C" int dp_split_tf32(const float* x, int64_t ld, int64_t bs, int32_t batch, int32_t rows, int32_t cols, int32_t transpose,
                             float* hi, float* lo, dp_stream_t stream) {
  DP_REQUIRE(x && hi && lo, DP_ERR_NULL);
  DP_REQUIRE(batch > 0 && rows > 0 && cols > 0 && ld >= cols && batch <= 65535, DP_ERR_SHAPE);
  // the padded tail of a row (cols4 / rows4) must be covered by the grid: round the covered extent up
  const int ccov = transpose ? cols : ((cols + 3) & ~3), rcov = transpose ? ((rows + 3) & ~3) : rows;
  dim3 grid((ccov + 31) / 32, (rcov + 31) / 32, batch);
  split_tf32_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, ld, bs, rows, cols, transpose, hi, lo);
  return dp_check_launch();
}
extern "C" int dp_transpose_batched(const float* in, float* out, int32_t batch, int32_t rows, int32_t cols, dp_stream_t stream) {
  DP_REQUIRE(in && out, DP_ERR_NULL);
  DP_REQUIRE(batch > 0 && rows > 0 && cols > 0 && batch <= 65535, DP_ERR_SHAPE);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  transpose_batched_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, rows, cols);
  return dp_check_launch();
}
extern "C" int dp_gemm_nt_tc(const dp_gemm_nt_args* a, dp_stream_t stream) {
  DP_REQUIRE(a && a->A && a->b_hi && a->b_lo && a->C, DP_ERR_NULL);
  DP_REQUIRE(a->batch > 0 && a->H > 0 && a->W > 0 && a->Kg > 0 && a->N > 0 && a->ld_a >= a->Kg && a->ldc >= a->N, DP_ERR_SHAPE);
  if (a->batch > 127) { /* the B "tap" index travels in a signed char table only for real taps; images use n0 directly */ }
  TapTable t{};
  t.n = 1;
  return launch_tc(a->A, a->ld_a, a->batch, a->H, a->W, a->Kg, a->b_hi, a->b_lo, a->N, a->batch, t, 1, 0, 0, a->H, a->W, a->C, a->ldc,
                   nullptr, nullptr, 0, nullptr, 0, 0, (cudaStream_t)stream, a->alpha, 1, 1, (a->Kg + 3) & ~3);
}

int dp_tc_runtime_ok() { return tc_init(); }

static TapTable dense_taps(int R, int S, int pad, bool flip) {
  TapTable t{};
  t.n = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      int i = r * S + s;
      t.dh[i] = (signed char)(r - pad); t.dw[i] = (signed char)(s - pad);
      t.wt[i] = (signed char)(flip ? (R * S - 1 - i) : i);
    }
  return t;
}

int dp_conv2d_fprop_tc(const dp_conv_args* a, dp_stream_t stream) {
  if (!a || !a->x || !a->y) return DP_ERR_UNSUPPORTED;   // let the SIMT entry produce the precise error
  if (a->R != a->S || (a->R != 1 && a->R != 3) || a->pad_l != a->pad_t) return DP_ERR_UNSUPPORTED;
  // stride 1: 'same' padding.  stride 2: 3x3 with pad 1, or pad 0 + the (0,1,0,1) zero border of Downsample2D (resnet.py:213-218) which
  // TMA out-of-bounds zero fill provides for free
  if (!((a->stride == 1 && a->pad_t == (a->R - 1) / 2) || (a->stride == 2 && a->R == 3 && (a->pad_t == 0 || a->pad_t == 1)))) return DP_ERR_UNSUPPORTED;
  if (a->P * a->stride != a->H || a->Q * a->stride != a->W) return DP_ERR_UNSUPPORTED;   // stride 2: even extents, out = in / 2 (Downsample2D, pad 1)
  if (a->N <= 0 || a->H <= 0 || a->W <= 0 || a->C <= 0 || a->K <= 0 || a->ldx < a->C || a->ldy < a->K) return DP_ERR_UNSUPPORTED;
  return launch_tc((const float*)a->x, a->ldx, a->N, a->P, a->Q, a->C, a->w_tc_hi, a->w_tc_lo, a->K, a->R * a->S,
                   dense_taps(a->R, a->S, a->pad_t, false), 1, 0, 0, a->P, a->Q, (float*)a->y, a->ldy, a->bias, a->rowadd,
                   a->ld_rowadd, a->residual, a->ld_res, (a->flags & DP_CONV_ACCUMULATE) ? 1 : 0, (cudaStream_t)stream, 1.0f, 0, a->stride, -1,
                   a->workspace);
}

// stride-1 dgrad == fprop of dy with the taps flipped and the (K,C) roles swapped: dx[n,h,w,c] = sum dy[n,h+1-r,w+1-s,k] W[k,c,r,s].
// stride-2 dgrad: dx[2i+a, 2j+b] only sees taps with (a+pad-r), (b+pad-s) even -> 4 parity classes, each a dense GEMM over the
// dy grid with 1/2/2/4 taps and a strided output mapping (no MACs wasted on structural zeros).
int dp_conv2d_dgrad_tc(const dp_conv_args* a, dp_stream_t stream) {
  if (!a || !a->x || !a->y) return DP_ERR_UNSUPPORTED;
  if (a->N <= 0 || a->H <= 0 || a->W <= 0 || a->C <= 0 || a->K <= 0 || a->ldx < a->C || a->ldy < a->K) return DP_ERR_UNSUPPORTED;
  if (a->R != a->S || (a->R != 1 && a->R != 3)) return DP_ERR_UNSUPPORTED;
  const int acc = (a->flags & DP_CONV_ACCUMULATE) ? 1 : 0;
  if (a->stride == 1) {
    if (a->pad_t != (a->R - 1) / 2 || a->pad_l != a->pad_t || a->P != a->H || a->Q != a->W) return DP_ERR_UNSUPPORTED;
    return launch_tc((const float*)a->y, a->ldy, a->N, a->H, a->W, a->K, a->w_tc_hi, a->w_tc_lo, a->C, a->R * a->S,
                     dense_taps(a->R, a->S, a->pad_t, true), 1, 0, 0, a->H, a->W, (float*)a->x, a->ldx, nullptr, nullptr, 0, nullptr, 0,
                     acc, (cudaStream_t)stream, 1.0f, 0, 1, -1, a->workspace);
  }
  if (a->stride != 2 || a->R != 3 || a->H != 2 * a->P || a->W != 2 * a->Q) return DP_ERR_UNSUPPORTED;
  TapTable cls[4];
  for (int ca = 0; ca < 2; ++ca)
    for (int cb = 0; cb < 2; ++cb) {
      TapTable& t = cls[ca * 2 + cb];
      t = TapTable{};
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          int nh = ca + a->pad_t - r, nw = cb + a->pad_l - s;
          if ((nh & 1) || (nw & 1)) continue;
          t.dh[t.n] = (signed char)(nh / 2); t.dw[t.n] = (signed char)(nw / 2); t.wt[t.n] = (signed char)(r * 3 + s);
          ++t.n;
        }
      if (t.n == 0) return DP_ERR_UNSUPPORTED;
    }
  for (int ca = 0; ca < 2; ++ca)
    for (int cb = 0; cb < 2; ++cb) {
      int rc = launch_tc((const float*)a->y, a->ldy, a->N, a->P, a->Q, a->K, a->w_tc_hi, a->w_tc_lo, a->C, 9, cls[ca * 2 + cb], 2, ca, cb,
                         a->H, a->W, (float*)a->x, a->ldx, nullptr, nullptr, 0, nullptr, 0, acc, (cudaStream_t)stream, 1.0f, 0, 1, -1,
                         a->workspace);
      if (rc != DP_OK) return (ca == 0 && cb == 0) ? rc : (rc == DP_ERR_UNSUPPORTED ? DP_ERR_SHAPE : rc);
    }
  return DP_OK;
}

// Floats of split-K workspace dp_conv2d_fprop (op 0) / dp_conv2d_dgrad (op 1) can use for this geometry (0: the launch fills the SMs
// without splitting).  With a->workspace == NULL the launch simply does not split.
extern "C" long long dp_conv_splitk_workspace_floats(const dp_conv_args* a, int op) {
  if (!a || a->N <= 0 || a->H <= 0 || a->W <= 0 || a->C <= 0 || a->K <= 0 || a->R != a->S || (a->R != 1 && a->R != 3)) return 0;
  long long need = 0;
  TapTable t{};
  if (op == 0) {
    t.n = a->R * a->S;
    launch_tc(nullptr, 0, a->N, a->P, a->Q, a->C, nullptr, nullptr, a->K, t.n, t, 1, 0, 0, a->P, a->Q, nullptr, 0, nullptr, nullptr, 0, nullptr, 0,
              0, nullptr, 1.0f, 0, a->stride, -1, nullptr, &need);
  } else if (a->stride == 1) {
    t.n = a->R * a->S;
    launch_tc(nullptr, 0, a->N, a->H, a->W, a->K, nullptr, nullptr, a->C, t.n, t, 1, 0, 0, a->H, a->W, nullptr, 0, nullptr, nullptr, 0, nullptr, 0,
              0, nullptr, 1.0f, 0, 1, -1, nullptr, &need);
  } else {
    for (int taps = 1; taps <= 4; taps *= 2) {     // the parity classes of a stride-2 3x3 dgrad have 1 / 2 / 2 / 4 taps and run back to back
      long long n = 0;
      t.n = taps;
      launch_tc(nullptr, 0, a->N, a->P, a->Q, a->K, nullptr, nullptr, a->C, 9, t, 2, 0, 0, a->H, a->W, nullptr, 0, nullptr, nullptr, 0, nullptr, 0,
                0, nullptr, 1.0f, 0, 1, -1, nullptr, &n);
      if (n > need) need = n;
    }
  }
  return need;
}

// 32-pixel K-chunk box of an [N][H][W] grid
static bool pick_box32(int H, int W, int& bw, int& bh, int& bn) {
  if (W >= WG_KPIX) { if (W % WG_KPIX) return false; bw = WG_KPIX; bh = 1; bn = 1; return true; }
  if (WG_KPIX % W) return false;
  bw = W;
  int rem = WG_KPIX / W;
  if (H >= rem) { if (H % rem) return false; bh = rem; bn = 1; return true; }
  if (rem % H) return false;
  bh = H; bn = rem / H;
  return true;
}

int dp_conv2d_wgrad_tc(const dp_conv_args* a, dp_stream_t stream) {
  if (!a || !a->x || !a->y || !a->workspace) return DP_ERR_UNSUPPORTED;
  if (!tc_init()) return DP_ERR_UNSUPPORTED;
  if (a->R != a->S || (a->R != 1 && a->R != 3) || a->pad_l != a->pad_t) return DP_ERR_UNSUPPORTED;
  // stride 1: 'same' padding.  stride 2: 3x3 with pad 1, or pad 0 + the (0,1,0,1) zero border of Downsample2D (resnet.py:213-218) which
  // TMA out-of-bounds zero fill provides for free
  if (!((a->stride == 1 && a->pad_t == (a->R - 1) / 2) || (a->stride == 2 && a->R == 3 && (a->pad_t == 0 || a->pad_t == 1)))) return DP_ERR_UNSUPPORTED;
  if (a->P * a->stride != a->H || a->Q * a->stride != a->W || a->splits < 1) return DP_ERR_UNSUPPORTED;
  if (a->ldx % 4 || a->ldy % 4 || ((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15)) return DP_ERR_UNSUPPORTED;
  int bw, bh, bn;
  if (!pick_box32(a->P, a->Q, bw, bh, bn)) return DP_ERR_UNSUPPORTED;   // 32-pixel chunks of the dy (output) grid
  if (a->N % bn) return DP_ERR_UNSUPPORTED;   // a partial image box would be fine (OOB zero) but keep chunks exact
  CUtensorMap mDy, mX;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->K, (cuuint64_t)a->Q, (cuuint64_t)a->P, (cuuint64_t)a->N};
    cuuint64_t str[3] = {(cuuint64_t)a->ldy * 4, (cuuint64_t)a->Q * a->ldy * 4, (cuuint64_t)a->P * a->Q * a->ldy * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    if (!make_map(&mDy, a->y, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return DP_ERR_UNSUPPORTED;
  }
  {   // x is sampled at stride * (output pixel) + tap offset: TMA element strides on W, H
    cuuint64_t dims[4] = {(cuuint64_t)a->C, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->N};
    cuuint64_t str[3] = {(cuuint64_t)a->ldx * 4, (cuuint64_t)a->W * a->ldx * 4, (cuuint64_t)a->H * a->W * a->ldx * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(bw * a->stride), (cuuint32_t)(bh * a->stride), (cuuint32_t)bn};
    if (!make_map(&mX, a->x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, a->stride)) return DP_ERR_UNSUPPORTED;
  }
  WgParams p{};
  p.Nimg = a->N; p.H = a->P; p.W = a->Q; p.C = a->C; p.K = a->K; p.R = a->R; p.S = a->S; p.pad = a->pad_t; p.in_stride = a->stride;
  p.bw = bw; p.bh = bh; p.bn = bn; p.tiles_w = a->Q / bw; p.tiles_h = a->P / bh;
  p.total_chunks = p.tiles_w * p.tiles_h * (a->N / bn);
  p.chunks_per_split = (p.total_chunks + a->splits - 1) / a->splits;
  p.c_tiles = (a->C + 127) / 128;
  p.ws = a->workspace;
  const int k_tiles = (a->K + 127) / 128;
  dim3 grid((unsigned)(k_tiles * p.c_tiles * a->R * a->S), (unsigned)a->splits);
  wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM, (cudaStream_t)stream>>>(mDy, mX, p);
  return dp_check_launch();
}

extern "C" int dp_pack_conv_weight_tc(const float* w, int32_t K, int32_t C, int32_t R, int32_t S, float* kc_hi, float* kc_lo,
                                      float* ck_hi, float* ck_lo, dp_stream_t stream) {
  DP_REQUIRE(w, DP_ERR_NULL);
  DP_REQUIRE(K > 0 && C > 0 && R > 0 && S > 0, DP_ERR_SHAPE);
  const int C4 = wrow(C), K4 = wrow(K);
  long long total = (long long)R * S * ((long long)K * C4 > (long long)C * K4 ? (long long)K * C4 : (long long)C * K4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, K, C, R * S, C4, K4, kc_hi, kc_lo, ck_hi, ck_lo);
  return dp_check_launch();
}

extern "C" int dp_tc_weight_row(int channels) { return channels > 0 ? wrow(channels) : 0; }
