// conv_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (fprop, and stride-1 dgrad as a
// tap-flipped fprop), fp32-grade via the 3xTF32 split:
//     x*w ~= hi(x)*hi(w) + lo(x)*hi(w) + hi(x)*lo(w),   hi = cvt.rna.tf32(.), lo = . - hi  (exact in fp32)
// with fp32 accumulation in TMEM.
//
// GEMM view: M = N*H*W output pixels (tile of 128 = one TMA box of the NHWC activation), N = output channels
// (tile BN), K = taps x input channels, one pipeline stage = (one tap, 32 channels) = a 128-byte swizzle row.
//
// Warp roles (192 threads, 1 CTA / SM):
//   warp 0      TMA producer: raw fp32 A box [128 px][32 ch] (im2col-free: the tap shift is a coordinate offset,
//               image borders are TMA out-of-bounds zero fill) + pre-split B_hi / B_lo weight tiles
//   warps 2..5  splitter: A (in place) -> hi, A_lo <- lo, smem->reg->smem, then fence.proxy.async + mbarrier
//   warp 1      MMA issuer (one lane): 4 k-steps x 3 tcgen05.mma.kind::tf32 per stage, tcgen05.commit frees the stage
//   warps 2..5  epilogue: tcgen05.ld accumulator rows, + bias / + per-image temb row / + residual / accumulate, store
//
// Kernels in this file (launch_tc picks one; DPB200_TC_PERSISTENT / DPB200_TC_SS / DPB200_TC_CLUSTER select the non-defaults):
//   conv_tc_ps_kernel        DEFAULT fprop/dgrad for > 64 output channels: persistent (1 CTA/SM loops over tiles), A hi/lo in
//                            shared memory (SS), 3 x 64 KB stages, two accumulator sets, N=256 fused hi-product instruction
//   conv_tc_ts_kernel<BN,CL> one tile per CTA, A through TMEM (TS); default for <= 64 output channels; CL > 1 = weight multicast
//   conv_tc_kernel<BN>       first SS kernel (DPB200_TC_SS), one tile per CTA
//   conv_tc_ab_kernel        (=2) one tile per CTA, TS, decoupled A / B / TMEM rings
//   conv_tc_ps2_kernel<BK>   (=3) ps + elect.sync issue, two issuer warps, optional 16-float stages x 7, clock64() trace stamps
//   conv_tc_pt_kernel        (=4; =5 picks it per layer for tiles with >= 27 stages) persistent TS: decoupled rings, two issuers, register-drained epilogue (round-2 candidate)
//   wgrad_tc_kernel          weight gradient: dY^T through TMEM, X split in shared memory, split-K over pixels
//   pack_tc / split_tf32 / transpose_batched helpers, dp_gemm_nt_tc (attention GEMMs on the persistent kernel)
// Measurements of every variant: profiles/r01_experiments.md.
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 32, STAGES = 3, NTHREADS = 192;
constexpr int A_BYTES = BM * BK * 4;  // 16 KB

struct TcParams {
  int Nimg, H, W;
  int Nout;            // GEMM N (valid output channels)
  int R, S, pad, flip;
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
  int two_issuers;          // persistent kernel: number of MMA-issuing warps taking alternate pipeline stages (DPB200_TC_ISSUERS=1|2)
  int wide_n;               // persistent kernel: fuse a_hi x b_hi and a_hi x b_lo into one N=256 instruction (DPB200_TC_WIDE_N=0 switches it off)
  int dbg_skip;             // timing experiments only (DPB200_TC_DEBUG_SKIP bitmask: 1 no A load, 2 no B_hi, 4 no B_lo) — results are then wrong
  long long* trace;         // pipeline trace buffer (dp_conv_tc_set_trace; nullptr = off): CTA 0 stamps clock64() per stage / actor
  int b_sub;                // decoupled-ring kernel: weight tiles fetched as b_sub boxes of BN/b_sub rows (experiment knob DPB200_B_SUB)
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
// the N=256 fused hi-product instruction is compiled in unconditionally (-DDPB200_RUNTIME_WIDE_N restores the DPB200_TC_WIDE_N=0|1
// run-time switch; the extra code path costs ~0.7 % of the C1 pass)
#ifdef DPB200_RUNTIME_WIDE_N
#define DP_WIDE_N(p) ((p).wide_n != 0)
#else
#define DP_WIDE_N(p) true
#endif
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
// A whole (converged) warp waits on a barrier.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
#ifdef DPB200_POLL_LANE0   // measured 14 % slower on the C1 pass than letting every lane poll (hardware-suspended try_wait)
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
               const __grid_constant__ CUtensorMap mapBl, const TcParams p) {
  constexpr int B_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to;                 // 1024B aligned: required by SWIZZLE_128B (TMA and UMMA agree)
  const uint32_t sbase = raw + pad_to;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t bar0 = sbase + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto conv_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (3 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // 2*BN columns: [0,BN) main accumulator (hi*hi), [BN,2BN) correction accumulator (lo*hi + hi*lo).  The tensor core
    // adds into an fp32 accumulator with one (truncating) rounding per MMA; keeping the 2^-11-smaller correction terms
    // out of the main accumulator cuts its rounding count 3x and keeps the small terms from being absorbed.
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // ---- tile coordinates
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
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
        const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
        const uint32_t st = sbase + s * STAGE_BYTES;
        tma_load_4d(st, &mapA, full_bar(s), kc * BK, q0 + p.dw[tap], p0 + p.dh[tap], n0);
        const int tapb = p.wt[tap];
        tma_load_3d(st + 2 * A_BYTES, &mapBh, full_bar(s), kc * BK, nblk * BN, tapb);
        tma_load_3d(st + 2 * A_BYTES + B_BYTES, &mapBl, full_bar(s), kc * BK, nblk * BN, tapb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4) | A=TF32 (2<<7) | B=TF32 (2<<10) | K-major A,B |
      // N>>3 at bit 17 | M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < num_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(conv_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = sbase + s * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t a_hi = umma_desc(st + k * 32), a_lo = umma_desc(st + A_BYTES + k * 32);
          const uint64_t b_hi = umma_desc(st + 2 * A_BYTES + k * 32), b_lo = umma_desc(st + 2 * A_BYTES + B_BYTES + k * 32);
          const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
          umma_tf32(tmem_base + BN, a_lo, b_hi, idesc, first);
          umma_tf32(tmem_base + BN, a_hi, b_lo, idesc, 1u);
          umma_tf32(tmem_base, a_hi, b_hi, idesc, first);
        }
        umma_commit(empty_bar(s));   // stage free once these MMAs have consumed it
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---- splitter: raw fp32 A tile -> (hi in place, lo)
    const int ct = threadIdx.x - 64;
    for (int it = 0; it < num_iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      mbar_arrive(conv_bar(s));
    }
    // ---- epilogue
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int w_l = row % p.bw, h_l = (row / p.bw) % p.bh, n_l = row / (p.bw * p.bh);
    const int img = n0 + n_l;
    const bool row_ok = img < p.Nimg;
    const long long m = ((long long)img * p.Ho + ((p0 + h_l) * p.os + p.oa)) * p.Wo + ((q0 + w_l) * p.os + p.ob);
    float* yrow = p.y + m * p.ldy;
    const float* rrow = p.residual ? p.residual + m * p.ld_res : nullptr;
    const float* arow = p.rowadd ? p.rowadd + (long long)img * p.ld_rowadd : nullptr;
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      uint32_t v[32], u[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
            "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
            "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
            "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
          : "r"(taddr + (uint32_t)BN));
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row_ok) {
        const int c0 = nblk * BN + j * 32;
        if (p.vec4 && c0 + 32 <= p.Nout) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o = make_float4(__uint_as_float(v[i]) + __uint_as_float(u[i]), __uint_as_float(v[i + 1]) + __uint_as_float(u[i + 1]),
                                   __uint_as_float(v[i + 2]) + __uint_as_float(u[i + 2]), __uint_as_float(v[i + 3]) + __uint_as_float(u[i + 3]));
            if (p.bias) { float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            if (arow) { float4 t = __ldg(reinterpret_cast<const float4*>(arow + c0 + i)); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
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
              if (arow) o += __ldg(arow + c);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}



// ------------------------------------------------------------------------------------------------ TS variant
// Same GEMM, but the split A operand goes registers -> TENSOR MEMORY (tcgen05.st) and the MMA reads A from TMEM
// (tcgen05.mma [d], [a_tmem], b_desc): the shared-memory pipe only carries the TMA writes, one read of the raw A tile and
// the B operand reads (112 KB per 768-cycle stage instead of 192 KB), which is what bounds the SS kernel above.
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

__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CL = thread-block-cluster size along the M tiles (1, 2 or 4).  The CL CTAs of a cluster work on different pixel tiles but the
// SAME weight tile: each CTA fetches 1/CL of B_hi / B_lo and TMA-multicasts it into all of them, so the L2->SM traffic per MMA
// stage drops from 48 KB to 16 + 32/CL KB — the L2->SM path (~42 B/clk/SM), not the tensor pipe, is what bounds this kernel.
template <int BN, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_ts_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                  const __grid_constant__ CUtensorMap mapBl, const TcParams p) {
  constexpr int B_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  constexpr uint32_t A_COL0 = 2 * BN;
  constexpr int SLICE_ROWS = BN / CL, SLICE_BYTES = B_BYTES / CL;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
  uint32_t cta_rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
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
    for (int s = 0; s < STAGES_TS; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), CL); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast / remote commit can reach them
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
        if (CL == 1) {
          tma_load_3d(st + A_BYTES, &mapBh, full_bar(s), kc * BK, nblk * BN, tapb);
          tma_load_3d(st + A_BYTES + B_BYTES, &mapBl, full_bar(s), kc * BK, nblk * BN, tapb);
        } else {   // my 1/CL row-slice of both weight tiles, multicast to every CTA of the cluster (same smem offsets)
          const int row0 = nblk * BN + (int)cta_rank * SLICE_ROWS;
          tma_load_3d_mc(st + A_BYTES + cta_rank * SLICE_BYTES, &mapBh, full_bar(s), kc * BK, row0, tapb, MC_MASK);
          tma_load_3d_mc(st + A_BYTES + B_BYTES + cta_rank * SLICE_BYTES, &mapBl, full_bar(s), kc * BK, row0, tapb, MC_MASK);
        }
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
        if (CL == 1) umma_commit(empty_bar(s)); else umma_commit_mc(empty_bar(s), MC_MASK);   // release the stage in EVERY CTA that multicasts into it
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
    // ---- epilogue (identical to the SS kernel)
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
  if (CL > 1) cluster_sync_all();   // nobody leaves while a peer may still signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------ decoupled-ring variant
// The measured limiter of the kernels above is not a throughput resource but the per-stage dependency CHAIN (TMA arrival ->
// split -> fence -> barrier -> MMA -> commit -> barrier -> next TMA ~ 3000 cycles for 768 cycles of MMA) with one ring that
// holds A and B together.  Here the two operands get their own rings and their own producers:
//   A ring : 6 x 16 KB raw tiles in shared memory -> splitter warps -> 4 hi/lo slots in TENSOR MEMORY; a shared-memory slot is
//            released by the SPLITTER (as soon as the tile is in registers), so the A TMA runs several stages ahead
//   B ring : 3 x (hi 16 KB + lo 16 KB); released by tcgen05.commit; its chain has no splitter hop at all
//   MMA    : waits "A slot in TMEM" + "B stage landed", issues 12 MMAs (A from TMEM), commits to both rings
// warps: 0 TMA-A | 1 TMA-B | 2 MMA (+TMEM alloc) | 3-6 splitter then epilogue (TMEM lane quarter = warp & 3).
constexpr int AB_THREADS = 224, AB_SA = 6, AB_SB = 3, AB_TA = 4;

__global__ void __launch_bounds__(AB_THREADS, 1)
conv_tc_ab_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                  const __grid_constant__ CUtensorMap mapBl, const TcParams p) {
  constexpr int BN = 128;
  constexpr int B_BYTES = BN * BK * 4;
  constexpr uint32_t A_COL0 = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to;
  const uint32_t sbase = raw + pad_to;
  const uint32_t a_base = sbase, b_base = sbase + AB_SA * A_BYTES;
  constexpr int DATA_BYTES = AB_SA * A_BYTES + AB_SB * 2 * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DATA_BYTES);
  const uint32_t bar0 = sbase + DATA_BYTES;
  auto fullA = [&](int i) { return bar0 + 8u * i; };
  auto emptyA = [&](int i) { return bar0 + 8u * (AB_SA + i); };
  auto fullB = [&](int i) { return bar0 + 8u * (2 * AB_SA + i); };
  auto emptyB = [&](int i) { return bar0 + 8u * (2 * AB_SA + AB_SB + i); };
  auto convT = [&](int i) { return bar0 + 8u * (2 * AB_SA + 2 * AB_SB + i); };
  auto emptyT = [&](int i) { return bar0 + 8u * (2 * AB_SA + 2 * AB_SB + AB_TA + i); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * AB_SA + 2 * AB_SB + 2 * AB_TA);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * AB_SA + 2 * AB_SB + 2 * AB_TA + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < AB_SA; ++i) { mbar_init(fullA(i), 1); mbar_init(emptyA(i), 128); }
    for (int i = 0; i < AB_SB; ++i) { mbar_init(fullB(i), 1); mbar_init(emptyB(i), 1); }
    for (int i = 0; i < AB_TA; ++i) { mbar_init(convT(i), 128); mbar_init(emptyT(i), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
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
      for (int it = 0; it < num_iters; ++it) {
        const int s = it % AB_SA;
        const uint32_t ph = (uint32_t)(it / AB_SA) & 1u;
        mbar_wait(emptyA(s), ph ^ 1u);
        mbar_expect_tx(fullA(s), A_BYTES);
        const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
        tma_load_4d(a_base + s * A_BYTES, &mapA, fullA(s), kc * BK, q0 + p.dw[tap], p0 + p.dh[tap], n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
      for (int it = 0; it < num_iters; ++it) {
        const int s = it % AB_SB;
        const uint32_t ph = (uint32_t)(it / AB_SB) & 1u;
        mbar_wait(emptyB(s), ph ^ 1u);
        mbar_expect_tx(fullB(s), 2 * B_BYTES);
        const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
        const int tapb = p.b_from_img ? n0 : p.wt[tap];
        const uint32_t st = b_base + s * 2 * B_BYTES;
        if (p.b_sub <= 1) {
          tma_load_3d(st, &mapBh, fullB(s), kc * BK, nblk * BN, tapb);
          tma_load_3d(st + B_BYTES, &mapBl, fullB(s), kc * BK, nblk * BN, tapb);
        } else {   // the same bytes as p.b_sub smaller boxes per operand: more TMA operations in flight, shorter per-load latency
          const int rows = BN / p.b_sub;
          for (int j = 0; j < p.b_sub; ++j) {
            tma_load_3d(st + j * rows * 128, &mapBh, fullB(s), kc * BK, nblk * BN + j * rows, tapb);
            tma_load_3d(st + B_BYTES + j * rows * 128, &mapBl, fullB(s), kc * BK, nblk * BN + j * rows, tapb);
          }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      const int n_valid = min(BN, p.Nout - nblk * BN);
      const uint32_t n_instr = (uint32_t)((n_valid + 15) & ~15);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_instr >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < num_iters; ++it) {
        const int sb = it % AB_SB, ta = it % AB_TA;
        mbar_wait(convT(ta), (uint32_t)(it / AB_TA) & 1u);     // A hi/lo of this step sit in TMEM slot ta
        mbar_wait(fullB(sb), (uint32_t)(it / AB_SB) & 1u);     // B hi/lo landed in shared memory
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = b_base + sb * 2 * B_BYTES;
        const uint32_t a_t = tmem_base + A_COL0 + 64u * ta;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t b_hi = umma_desc(st + k * 32), b_lo = umma_desc(st + B_BYTES + k * 32);
          const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
          umma_tf32_ts(tmem_base + BN, a_t + 32 + k * 8, b_hi, idesc, first);   // lo * hi
          umma_tf32_ts(tmem_base + BN, a_t + k * 8, b_lo, idesc, 1u);           // hi * lo
          umma_tf32_ts(tmem_base, a_t + k * 8, b_hi, idesc, first);             // hi * hi
        }
        umma_commit(emptyB(sb));
        umma_commit(emptyT(ta));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < num_iters; ++it) {
      const int sa = it % AB_SA, ta = it % AB_TA;
      mbar_wait(fullA(sa), (uint32_t)(it / AB_SA) & 1u);
      const uint8_t* arow = smem + sa * A_BYTES + row * 128;
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
      mbar_arrive(emptyA(sa));                                  // tile is in registers: the A TMA may refill this slot
      mbar_wait(emptyT(ta), ((uint32_t)(it / AB_TA) & 1u) ^ 1u);   // the MMAs that read TMEM slot ta (4 steps ago) are done
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_t = tmem_base + lane_addr + A_COL0 + 64u * ta;
      tmem_st32(a_t, hi);
      tmem_st32(a_t + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(convT(ta));
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int q0, p0, n0, nblk;
        tile_coords(tile, q0, p0, n0, nblk);
        for (int it = 0; it < iters_per_tile; ++it, ++g) {
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const int nblk = tile / tiles_m;
        const int n_valid = min(BN, p.Nout - nblk * BN);
        const uint32_t n_instr = (uint32_t)((n_valid + 15) & ~15);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_instr >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t b = tl & 1u, use = tl >> 1;
        mbar_wait(tempty_bar(b), (use & 1u) ^ 1u);          // epilogue has drained this accumulator set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + b * 256u;
        for (int it = 0; it < iters_per_tile; ++it, ++g) {
          const int s = g % PS_STAGES;
          const uint32_t ph = (g / PS_STAGES) & 1u;
          mbar_wait(conv_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = sbase + s * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t a_hi = umma_desc(st + k * 32), a_lo = umma_desc(st + A_BYTES + k * 32);
            const uint64_t b_hi = umma_desc(st + 2 * A_BYTES + k * 32), b_lo = umma_desc(st + 2 * A_BYTES + B_BYTES + k * 32);
            const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
            if (DP_WIDE_N(p)) {   // a_hi x [b_hi | b_lo] -> [main | correction] as ONE N=256 instruction (the two B tiles are adjacent)
              umma_tf32(acc, a_hi, b_hi, idesc256, first);
              umma_tf32(acc + 128, a_lo, b_hi, idesc, 1u);
            } else {
              umma_tf32(acc + 128, a_lo, b_hi, idesc, first);
              umma_tf32(acc + 128, a_hi, b_lo, idesc, 1u);
              umma_tf32(acc, a_hi, b_hi, idesc, first);
            }
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
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int it = 0; it < iters_per_tile; ++it, ++g) {
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
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
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
        if (row_ok) {
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

// ---- experimental successor of conv_tc_ps_kernel (DPB200_TC_PERSISTENT=3): warp-converged elect.sync issue, two MMA-issuer
// warps on alternate stages, optional 16-float stages x 7, optional N=256 fused instruction, per-stage clock64() trace
// (dp_conv_tc_set_trace).  Measured in profiles/r01_experiments.md; not the default because on the whole C1 pass it is
// ~8 % slower than the kernel above (21.7 vs 20.0 ms of fprop+dgrad) although its steady-state stage period is shorter.
constexpr int PS2_THREADS = 352;   // warps: 0 TMA producer | 1, 10 MMA issuers (alternate stages) | 2-5 splitters | 6-9 epilogue
// K-chunk per pipeline stage and ring depth: 16 floats (64-byte swizzle atoms, 32 KB stages) x 7, or 32 floats (128-byte atoms, 64 KB) x 3.
// The per-stage round trip (slot freed -> TMA -> split -> MMA -> commit) is ~3000 clk; 3 x 768 clk of MMA work in flight cannot
// cover it, 7 x 384 clk with a shorter split/MMA leg can (profiles/r01_experiments.md).
template <int BKT> struct PsCfg { static constexpr int STAGES = (BKT == 16) ? 7 : 3; };
template <int BKT> __device__ __forceinline__ uint64_t umma_desc_ps(uint32_t saddr) {
  if (BKT == 32) return umma_desc(saddr);
  // K-major SWIZZLE_64B: 8-row groups of 64-byte rows, SBO = 512 B, layout_type 4
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}   // warps: 0 TMA producer | 1 MMA issuer (even stages) | 2-5 splitters | 6-9 epilogue | 10 MMA issuer (odd stages)

constexpr int TRACE_STAGES = 1024;
#define DP_TRACE_TILE(slot, tl) do { if (p.trace && blockIdx.x == 0 && (tl) < 64u) p.trace[TRACE_STAGES * 16 + (tl) * 4 + (slot)] = clock64(); } while (0)
#define DP_TRACE(slot, g) do { if (p.trace && blockIdx.x == 0 && (g) < (uint32_t)TRACE_STAGES) p.trace[(g) * 16 + (slot)] = clock64(); } while (0)

template <int BKT>
__global__ void __launch_bounds__(PS2_THREADS, 1)
conv_tc_ps2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                  const __grid_constant__ CUtensorMap mapBl, const TcParams p, const int tiles_m, const int total_tiles) {
  constexpr int BN = 128;
  constexpr int PS_STAGES = PsCfg<BKT>::STAGES;
  constexpr int BK = BKT;                       // shadows the file-scope K chunk
  constexpr int A_BYTES = BM * BKT * 4;
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
  auto iss_bar = [&](int s) { return bar0 + 8u * (3 * PS_STAGES + 4 + s); };   // "MMAs of the stage in slot s are in the tensor queue"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * PS_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PS_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); mbar_init(iss_bar(s), 1); }
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
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
    }
    __syncwarp();
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int q0, p0, n0, nblk;
      tile_coords(tile, q0, p0, n0, nblk);
      int tap = 0, kc = 0;
      for (int it = 0; it < iters_per_tile; ++it, ++g) {
        const int s = g % PS_STAGES;
        const uint32_t ph = (g / PS_STAGES) & 1u;
        const int c_k = kc * BK, c_w = q0 + p.dw[tap], c_h = p0 + p.dh[tap];
        const int tapb = p.b_from_img ? n0 : p.wt[tap];
        const uint32_t st = sbase + s * STAGE_BYTES;
        if (++kc == p.kchunks) { kc = 0; ++tap; }
        mbar_wait_warp(empty_bar(s), ph ^ 1u);
        if (elect_one()) {
          DP_TRACE(0, g);
          if (p.dbg_skip == 0) {
            mbar_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
            tma_load_4d(st, &mapA, full_bar(s), c_k, c_w, c_h, n0);
            tma_load_3d(st + 2 * A_BYTES, &mapBh, full_bar(s), c_k, nblk * BN, tapb);
            tma_load_3d(st + 2 * A_BYTES + B_BYTES, &mapBl, full_bar(s), c_k, nblk * BN, tapb);
          } else {   // timing experiments: drop some of the loads
            const int sk = p.dbg_skip;
            mbar_expect_tx(full_bar(s), ((sk & 1) ? 0 : A_BYTES) + ((sk & 2) ? 0 : B_BYTES) + ((sk & 4) ? 0 : B_BYTES));
            if (!(sk & 1)) tma_load_4d(st, &mapA, full_bar(s), c_k, c_w, c_h, n0);
            if (!(sk & 2)) tma_load_3d(st + 2 * A_BYTES, &mapBh, full_bar(s), c_k, nblk * BN, tapb);
            if (!(sk & 4)) tma_load_3d(st + 2 * A_BYTES + B_BYTES, &mapBl, full_bar(s), c_k, nblk * BN, tapb);
          }
          DP_TRACE(1, g);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || warp == 10) {
    // Two issuer warps take alternate pipeline stages.  Every tcgen05.mma / tcgen05.commit holds its uniform-register
    // operands until the tensor queue has consumed it, so ptxas makes the issuing warp wait on that scoreboard before it
    // may set up the next stage: a lone issuer therefore stalls until ITS stage has drained and the tensor pipe idles for
    // the ~450 clk it then needs to poll the barrier and rebuild descriptors (profiles/r01_experiments.md, pipeline trace).
    // With two warps one is always ahead, queueing stage g+1 behind stage g.  Queue order across the two warps is kept by
    // the iss_bar hand-off (arrive after the stage's MMAs are issued; the other warp waits on it before issuing).
    const uint32_t mw = (warp == 1) ? 0u : 1u;
    const uint32_t ni = (uint32_t)p.two_issuers;   // number of issuer warps in the rotation (1 or 2)
    const uint32_t two = ni > 1u ? 1u : 0u;
    if (mw < ni) {
      uint32_t g = 0, tl = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const int nblk = tile / tiles_m;
        const int n_valid = min(BN, p.Nout - nblk * BN);
        const uint32_t n_instr = (uint32_t)((n_valid + 15) & ~15);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_instr >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t b = tl & 1u, use = tl >> 1;
        mbar_wait_warp(tempty_bar(b), (use & 1u) ^ 1u);          // epilogue has drained this accumulator set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mw == 0 && lane == 0) DP_TRACE_TILE(3, tl);
        const uint32_t acc = tmem_base + b * 256u;
        for (int it = 0; it < iters_per_tile; ++it, ++g) {
          if (g % ni != mw) continue;
          const int s = g % PS_STAGES;
          const uint32_t ph = (g / PS_STAGES) & 1u;
          if (lane == 0) DP_TRACE(8, g);
          mbar_wait_warp(conv_bar(s), ph);
          if (lane == 0) DP_TRACE(9, g);
          if (two && g > 0) mbar_wait_warp(iss_bar((g - 1) % PS_STAGES), ((g - 1) / PS_STAGES) & 1u);
          if (lane == 0) DP_TRACE(10, g);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = sbase + s * STAGE_BYTES;
          if (elect_one()) {
            DP_TRACE(4, g);
            const uint64_t a_hi0 = umma_desc_ps<BKT>(st), a_lo0 = umma_desc_ps<BKT>(st + A_BYTES);
            const uint64_t b_hi0 = umma_desc_ps<BKT>(st + 2 * A_BYTES), b_lo0 = umma_desc_ps<BKT>(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {   // +32 B per K step = +2 in the descriptor's 16-byte address field
              const uint64_t a_hi = a_hi0 + 2 * k, a_lo = a_lo0 + 2 * k, b_hi = b_hi0 + 2 * k, b_lo = b_lo0 + 2 * k;
              const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
              if (DP_WIDE_N(p)) {
                // B_hi and B_lo are adjacent 128-row tiles: ONE N=256 instruction computes a_hi x [b_hi | b_lo] into
                // [main | correction] (acc .. acc+255); a_lo x b_hi then adds into the correction half.  Same tensor time as
                // three N=128 instructions, but 2/3 of the instructions and 5/6 of the operand reads from shared memory.
                umma_tf32(acc, a_hi, b_hi, idesc256, first);
                umma_tf32(acc + 128, a_lo, b_hi, idesc, 1u);
              } else {
                umma_tf32(acc + 128, a_lo, b_hi, idesc, first);
                umma_tf32(acc + 128, a_hi, b_lo, idesc, 1u);
                umma_tf32(acc, a_hi, b_hi, idesc, first);
              }
            }
            umma_commit(empty_bar(s));
            if (it == iters_per_tile - 1) umma_commit(tfull_bar(b));
            if (two) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              mbar_arrive(iss_bar(s));
            }
            DP_TRACE(5, g);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 6) {
    // ---- splitter warps 2..5: A tile -> tf32 hi (in place) + lo (side buffer), elementwise so the swizzle is irrelevant
    const int ct = (int)threadIdx.x - 64;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int it = 0; it < iters_per_tile; ++it, ++g) {
        const int s = g % PS_STAGES;
        const uint32_t ph = (g / PS_STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        if (ct == 0) DP_TRACE(2, g);
        float4* A = reinterpret_cast<float4*>(smem + s * STAGE_BYTES);
        float4* Al = reinterpret_cast<float4*>(smem + s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
          const int idx = ct + 128 * i;
          float4 v = A[idx], h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          A[idx] = h;
          Al[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (ct == 0) DP_TRACE(6, g);
        mbar_arrive(conv_bar(s));
        if ((ct & 31) == 0) DP_TRACE(12 + (ct >> 5), g);   // per-warp arrival
      }
    }
  } else if (warp < 10) {
    // ---- epilogue warps 6..9 (TMEM lane quarter = warp & 3)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int w_l = row % p.bw, h_l = (row / p.bw) % p.bh, n_l = row / (p.bw * p.bh);
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      int q0, p0, n0, nblk;
      tile_coords(tile, q0, p0, n0, nblk);
      const uint32_t b = tl & 1u, use = tl >> 1;
      mbar_wait(tfull_bar(b), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (threadIdx.x == 6 * 32) DP_TRACE_TILE(0, tl);
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
          if (threadIdx.x == 6 * 32) DP_TRACE_TILE(1, tl);
        }
        if (row_ok) {
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
      if (threadIdx.x == 6 * 32) DP_TRACE_TILE(2, tl);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ persistent TS variant
// Round-2 candidate built from what the pipeline timeline showed (profiles/r01_experiments.md), DPB200_TC_PERSISTENT=4:
//  * A operand through TENSOR MEMORY (TS mode): the raw fp32 A box lands in a 6-deep ring of 16 KB stages, a splitter thread
//    per pixel row converts it in registers and writes hi/lo with tcgen05.st into a 4-deep ring of TMEM slots — the raw
//    stage is free as soon as it is in registers, no generic->async proxy fence, and the MMA reads only B from shared memory;
//  * B (pre-split weights, hi | lo adjacent) in its own 3-deep ring of 32 KB stages: 192 KB in total, but the A, B and TMEM
//    rings are released independently, so the ~3000-clk TMA -> split -> MMA -> commit round trip is covered;
//  * per 8-float K step two instructions: a_hi x [b_hi | b_lo] (N=256) -> [main | correction], a_lo x b_hi (N=128) -> correction;
//  * two issuer warps on alternate stages (a lone issuer cannot run ahead of the tensor queue);
//  * persistent tiles with ONE accumulator set (the A slots use the other 256 TMEM columns): the epilogue warps first drain
//    main + correction into 128 registers per thread (a few hundred clk), hand the accumulator back, and only then do the
//    bias / residual / store work, which overlaps the next tile's main loop.
constexpr int PT_THREADS = 384;   // warps: 0 A producer | 1 B producer | 2, 11 MMA issuers | 3-6 splitters | 7-10 epilogue
constexpr int PT_SA = 6, PT_SB = 3, PT_TA = 4;

__global__ void __launch_bounds__(PT_THREADS, 1)
conv_tc_pt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                  const __grid_constant__ CUtensorMap mapBl, const TcParams p, const int tiles_m, const int total_tiles) {
  constexpr int BN = 128;
  constexpr int B_BYTES = BN * BK * 4;
  constexpr uint32_t A_COL0 = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad_to = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad_to;
  const uint32_t sbase = raw + pad_to;
  const uint32_t a_base = sbase, b_base = sbase + PT_SA * A_BYTES;
  constexpr int DATA_BYTES = PT_SA * A_BYTES + PT_SB * 2 * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DATA_BYTES);
  const uint32_t bar0 = sbase + DATA_BYTES;
  auto fullA = [&](int i) { return bar0 + 8u * i; };
  auto emptyA = [&](int i) { return bar0 + 8u * (PT_SA + i); };
  auto fullB = [&](int i) { return bar0 + 8u * (2 * PT_SA + i); };
  auto emptyB = [&](int i) { return bar0 + 8u * (2 * PT_SA + PT_SB + i); };
  auto convT = [&](int i) { return bar0 + 8u * (2 * PT_SA + 2 * PT_SB + i); };
  auto emptyT = [&](int i) { return bar0 + 8u * (2 * PT_SA + 2 * PT_SB + PT_TA + i); };
  auto issT = [&](int i) { return bar0 + 8u * (2 * PT_SA + 2 * PT_SB + 2 * PT_TA + i); };   // "stage in TMEM slot i is queued"
  constexpr int NB0 = 2 * PT_SA + 2 * PT_SB + 3 * PT_TA;
  const uint32_t tfull_bar = bar0 + 8u * NB0, tempty_bar = bar0 + 8u * (NB0 + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NB0 + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < PT_SA; ++i) { mbar_init(fullA(i), 1); mbar_init(emptyA(i), 128); }
    for (int i = 0; i < PT_SB; ++i) { mbar_init(fullB(i), 1); mbar_init(emptyB(i), 1); }
    for (int i = 0; i < PT_TA; ++i) { mbar_init(convT(i), 128); mbar_init(emptyT(i), 1); mbar_init(issT(i), 1); }
    mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
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
    // ---- A producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int q0, p0, n0, nblk;
        tile_coords(tile, q0, p0, n0, nblk);
        int tap = 0, kc = 0;
        for (int it = 0; it < iters_per_tile; ++it, ++g) {
          const int s = g % PT_SA;
          const int c_k = kc * BK, c_w = q0 * p.in_stride + p.dw[tap], c_h = p0 * p.in_stride + p.dh[tap];
          if (++kc == p.kchunks) { kc = 0; ++tap; }
          mbar_wait(emptyA(s), ((g / PT_SA) & 1u) ^ 1u);
          DP_TRACE(0, g);
          mbar_expect_tx(fullA(s), A_BYTES);
          tma_load_4d(a_base + s * A_BYTES, &mapA, fullA(s), c_k, c_w, c_h, n0);
          DP_TRACE(1, g);
        }
      }
    }
  } else if (warp == 1) {
    // ---- B producer (hi | lo adjacent: one N=256 descriptor covers both)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int q0, p0, n0, nblk;
        tile_coords(tile, q0, p0, n0, nblk);
        int tap = 0, kc = 0;
        for (int it = 0; it < iters_per_tile; ++it, ++g) {
          const int s = g % PT_SB;
          const int c_k = kc * BK;
          const int tapb = p.b_from_img ? n0 : p.wt[tap];
          if (++kc == p.kchunks) { kc = 0; ++tap; }
          mbar_wait(emptyB(s), ((g / PT_SB) & 1u) ^ 1u);
          const uint32_t st = b_base + s * 2 * B_BYTES;
          if (p.dbg_skip == 0) {
            mbar_expect_tx(fullB(s), 2 * B_BYTES);
            tma_load_3d(st, &mapBh, fullB(s), c_k, nblk * BN, tapb);
            tma_load_3d(st + B_BYTES, &mapBl, fullB(s), c_k, nblk * BN, tapb);
          } else {   // timing experiments (DPB200_TC_DEBUG_SKIP: 2 no B_hi, 4 no B_lo): results are then wrong
            mbar_expect_tx(fullB(s), ((p.dbg_skip & 2) ? 0 : B_BYTES) + ((p.dbg_skip & 4) ? 0 : B_BYTES));
            if (!(p.dbg_skip & 2)) tma_load_3d(st, &mapBh, fullB(s), c_k, nblk * BN, tapb);
            if (!(p.dbg_skip & 4)) tma_load_3d(st + B_BYTES, &mapBl, fullB(s), c_k, nblk * BN, tapb);
          }
        }
      }
    }
  } else if (warp == 2 || warp == 11) {
    // ---- two MMA issuer warps on alternate stages (warp-converged, one elected lane; queue order kept by the issT hand-off)
    const uint32_t mw = (warp == 2) ? 0u : 1u;
    const uint32_t ni = (uint32_t)p.two_issuers;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    if (mw < ni) {
      uint32_t g = 0, tl = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        mbar_wait(tempty_bar, (tl & 1u) ^ 1u);            // the epilogue has drained the accumulator of the previous tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mw == 0 && lane == 0) DP_TRACE_TILE(3, tl);
        for (int it = 0; it < iters_per_tile; ++it, ++g) {
          if (g % ni != mw) continue;
          const int sb = g % PT_SB, ta = g % PT_TA;
          if (lane == 0) DP_TRACE(8, g);
          mbar_wait(convT(ta), (g / PT_TA) & 1u);         // A hi/lo of this stage sit in TMEM slot ta
          if (lane == 0) DP_TRACE(9, g);
          mbar_wait(fullB(sb), (g / PT_SB) & 1u);         // B hi/lo landed in shared memory
          if (lane == 0) DP_TRACE(7, g);
          if (ni > 1u && g > 0) mbar_wait(issT((g - 1) % PT_TA), ((g - 1) / PT_TA) & 1u);
          if (lane == 0) DP_TRACE(10, g);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = b_base + sb * 2 * B_BYTES;
          const uint32_t a_t = tmem_base + A_COL0 + 64u * ta;
          if (elect_one()) {
            DP_TRACE(4, g);
            const uint64_t b_hi0 = umma_desc(st);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t b_hi = b_hi0 + 2 * k;
              const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
              umma_tf32_ts(tmem_base, a_t + k * 8, b_hi, idesc256, first);        // a_hi x [b_hi | b_lo] -> [main | correction]
              umma_tf32_ts(tmem_base + BN, a_t + 32 + k * 8, b_hi, idesc, 1u);    // a_lo x b_hi -> correction
            }
            umma_commit(emptyB(sb));
            umma_commit(emptyT(ta));
            if (it == iters_per_tile - 1) umma_commit(tfull_bar);
            if (ni > 1u) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              mbar_arrive(issT(ta));
            }
            DP_TRACE(5, g);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 7) {
    // ---- splitter warps 3..6: thread <-> pixel row (TMEM lane quarter = warp & 3)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int it = 0; it < iters_per_tile; ++it, ++g) {
        const int sa = g % PT_SA, ta = g % PT_TA;
        mbar_wait(fullA(sa), (g / PT_SA) & 1u);
        if (row == 0) DP_TRACE(2, g);
        const uint8_t* arow = smem + sa * A_BYTES + row * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // SWIZZLE_128B: 16-byte chunk j of row r sits at chunk position j ^ (r & 7)
          const float4 v = *reinterpret_cast<const float4*>(arow + ((j ^ (row & 7)) << 4));
          const float h0 = tf32_rna(v.x), h1 = tf32_rna(v.y), h2 = tf32_rna(v.z), h3 = tf32_rna(v.w);
          hi[4 * j + 0] = __float_as_uint(h0); hi[4 * j + 1] = __float_as_uint(h1);
          hi[4 * j + 2] = __float_as_uint(h2); hi[4 * j + 3] = __float_as_uint(h3);
          lo[4 * j + 0] = __float_as_uint(v.x - h0); lo[4 * j + 1] = __float_as_uint(v.y - h1);
          lo[4 * j + 2] = __float_as_uint(v.z - h2); lo[4 * j + 3] = __float_as_uint(v.w - h3);
        }
        mbar_arrive(emptyA(sa));                                     // tile is in registers: the A TMA may refill this slot
        if (row == 0) DP_TRACE(6, g);
        mbar_wait(emptyT(ta), ((g / PT_TA) & 1u) ^ 1u);              // the MMAs that read TMEM slot ta (4 stages ago) are done
        if (row == 0) DP_TRACE(11, g);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_t = tmem_base + lane_addr + A_COL0 + 64u * ta;
        tmem_st32(a_t, hi);
        tmem_st32(a_t + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(convT(ta));
        if (lane == 0) DP_TRACE(12 + q, g);   // per-warp arrival (slot 12 + TMEM lane quarter); row 0 is quarter 0
        if (row == 0) DP_TRACE(3, g);
      }
    }
  } else if (warp < 11) {
    // ---- epilogue warps 7..10 (TMEM lane quarter = warp & 3): drain main + correction into registers, release, then store
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int w_l = row % p.bw, h_l = (row / p.bw) % p.bh, n_l = row / (p.bw * p.bh);
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      int q0, p0, n0, nblk;
      tile_coords(tile, q0, p0, n0, nblk);
      mbar_wait(tfull_bar, tl & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (row == 0) DP_TRACE_TILE(0, tl);
      float acc[BN];
#pragma unroll
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
            : "r"(taddr + 128u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[j * 32 + i] = p.alpha * (__uint_as_float(v[i]) + __uint_as_float(u[i]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty_bar);                                        // the issuers may start the next tile
      if (row == 0) DP_TRACE_TILE(1, tl);
      const int img = n0 + n_l;
      if (img < p.Nimg) {
        const long long m = ((long long)img * p.Ho + ((p0 + h_l) * p.os + p.oa)) * p.Wo + ((q0 + w_l) * p.os + p.ob);
        float* yrow = p.y + m * p.ldy;
        const float* rrow = p.residual ? p.residual + m * p.ld_res : nullptr;
        const float* arow2 = p.rowadd ? p.rowadd + (long long)img * p.ld_rowadd : nullptr;
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
          const int c0 = nblk * BN + j * 32;
          if (p.vec4 && c0 + 32 <= p.Nout) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 o = make_float4(acc[j * 32 + i], acc[j * 32 + i + 1], acc[j * 32 + i + 2], acc[j * 32 + i + 3]);
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
                float o = acc[j * 32 + i];
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
      if (row == 0) DP_TRACE_TILE(2, tl);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
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
  int two_issuers, wide_n;   // DPB200_TC_ISSUERS / DPB200_TC_WIDE_N (see TcParams)
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
    // two issuer warps on alternate stages, warp-converged with one elected lane (see conv_tc_ps_kernel)
    const uint32_t mw = (warp == 1) ? 0u : 1u;
    const uint32_t two = p.two_issuers ? 1u : 0u;
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
            const uint64_t b_hi = umma_desc_mn(st + WG_T + k * 1024), b_lo = umma_desc_mn(st + 2 * WG_T + k * 1024);
            const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
            if (DP_WIDE_N(p)) {   // x_hi | x_lo are adjacent 4 x 32-channel block groups: a_hi x [x_hi | x_lo] -> [main | correction]
              umma_tf32_ts(tmem_base, a_t + k * 8, b_hi, idesc256, first);
              umma_tf32_ts(tmem_base + 128, a_t + 32 + k * 8, b_hi, idesc, 1u);
            } else {
              umma_tf32_ts(tmem_base + 128, a_t + 32 + k * 8, b_hi, idesc, first);
              umma_tf32_ts(tmem_base + 128, a_t + k * 8, b_lo, idesc, 1u);
              umma_tf32_ts(tmem_base, a_t + k * 8, b_hi, idesc, first);
            }
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
bool g_use_ss = false;
int g_persistent = 1;  // DPB200_TC_PERSISTENT: 1 = persistent SS kernel, 2 = decoupled A/B-ring TS kernel, 3 = two-issuer experimental kernel, 4 = persistent TS kernel (round-2 candidate), 5 = TS kernel for long-K tiles / default kernel otherwise, 0 = one-ring TS kernel
static int g_pt_min_stages = 27;        // DPB200_TC_PT_MIN_STAGES: with DPB200_TC_PERSISTENT=5, tiles with at least this many pipeline stages use conv_tc_pt_kernel
static int g_ps_bk = 32;               // DPB200_TC_PS_BK: K chunk per stage of the persistent kernel (16 -> 7 stages, 32 -> 3 stages)
static constexpr int ps_smem_bytes(int bk) { return (bk == 16 ? 7 : 3) * 4 * 128 * bk * 4 + 2048; }
// Row length of the packed TF32 weight tiles (dp_pack_conv_weight_tc): rows longer than 32 floats are zero-padded to a multiple of 32
// floats (128 B) so that every 32-float TMA box row is exactly one aligned 128-byte line; short rows to a multiple of 4 (the TMA
// 16-byte stride rule).  With 16-byte padding only, pruned widths (90 / 179 input channels) ran 20-25 % slower than the next
// multiple of 32 (scripts/time_conv_shapes.py: 90 -> 90 3x3 @32x32 167 us vs 134 us).  DPB200_WROW_PAD=4 restores the old layout.
static int wrow(int c) {
  static const int pad = getenv("DPB200_WROW_PAD") ? atoi(getenv("DPB200_WROW_PAD")) : 32;
  return (pad == 32 && c > 32) ? ((c + 31) & ~31) : ((c + 3) & ~3);
}
static long long* g_trace = nullptr;   // see dp_conv_tc_set_trace
int g_num_sms = 148;
int g_cluster = 1;     // DPB200_TC_CLUSTER=2|4: CTAs per cluster sharing (TMA-multicasting) one weight tile.  Measured on B200
                       // (profiles/r01_experiments.md): 46.7 / 47.5 / 48.3 ms per pass for 1 / 2 / 4 -> off by default. // DPB200_TC_SS=1: keep the A operand in shared memory (SS-mode kernel) instead of TMEM (TS-mode)
std::mutex g_tc_mutex;

int tc_init() {
  std::lock_guard<std::mutex> lk(g_tc_mutex);
  if (g_tc_state >= 0) return g_tc_state;
  g_tc_state = 0;
  if (getenv("DPB200_FORCE_SIMT")) return 0;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) { (void)cudaGetLastError(); return 0; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) { (void)cudaGetLastError(); return 0; }
  g_encode = (EncodeTiledFn)fn;
  bool ok = cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 STAGES * (2 * A_BYTES + 2 * 128 * BK * 4) + 2048) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  STAGES * (2 * A_BYTES + 2 * 64 * BK * 4) + 2048) == cudaSuccess;
  const int smem128 = STAGES_TS * (A_BYTES + 2 * 128 * BK * 4) + 2048, smem64 = STAGES_TS * (A_BYTES + 2 * 64 * BK * 4) + 2048;
  ok = ok && cudaFuncSetAttribute(conv_tc_ts_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_ts_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_ts_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_ts_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64) == cudaSuccess;
  if (const char* e = getenv("DPB200_TC_CLUSTER")) g_cluster = atoi(e);
  ok = ok && cudaFuncSetAttribute(conv_tc_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_STAGES * (2 * A_BYTES + 2 * 128 * BK * 4) + 2048) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_pt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SA * A_BYTES + PT_SB * 2 * 128 * BK * 4 + 2048) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_ps2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ps_smem_bytes(32)) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(conv_tc_ps2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ps_smem_bytes(16)) == cudaSuccess;
  if (const char* e = getenv("DPB200_TC_PS_BK")) g_ps_bk = atoi(e) == 32 ? 32 : 16;
  if (const char* e = getenv("DPB200_TC_PT_MIN_STAGES")) g_pt_min_stages = atoi(e);
  if (const char* e = getenv("DPB200_TC_PERSISTENT")) g_persistent = atoi(e);
  ok = ok && cudaFuncSetAttribute(conv_tc_ab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SA * A_BYTES + AB_SB * 2 * 128 * BK * 4 + 2048) == cudaSuccess;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); }
  g_use_ss = getenv("DPB200_TC_SS") != nullptr;
  ok = ok && cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * WG_T + 2048) == cudaSuccess;
  if (!ok) { (void)cudaGetLastError(); return 0; }
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
// Shared launcher.  act: [Nimg][H][W][Kg] view (ld_act) = A operand on whose pixel grid the M tiles live; w_hi/w_lo: [T][Nout][Kg];
// out: [Nimg][Ho][Wo][Nout] view, output pixel = (p*os+oa, q*os+ob).
int launch_tc(const float* act, long long ld_act, int Nimg, int H, int W, int Kg, const float* w_hi, const float* w_lo, int Nout,
              int T, const TapTable& taps, int os, int oa, int ob, int Ho, int Wo, float* out, long long ld_out, const float* bias,
              const float* rowadd, long long ld_rowadd, const float* residual, long long ld_res, int accumulate, cudaStream_t st,
              float alpha = 1.0f, int b_from_img = 0, int in_stride = 1, int ldb = -1) {
  if (ldb < 0) ldb = wrow(Kg);   // packed conv weights; batched GEMM callers pass their own row pitch
  if (!tc_init()) return DP_ERR_UNSUPPORTED;
  if (!w_hi || !w_lo) return DP_ERR_UNSUPPORTED;
  if (ld_act % 4 || ((uintptr_t)act & 15) || ((uintptr_t)w_hi & 15) || ((uintptr_t)w_lo & 15)) return DP_ERR_UNSUPPORTED;
  int bw, bh, bn;
  if (!pick_box(Nimg, H, W, bw, bh, bn)) return DP_ERR_UNSUPPORTED;
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
  const int BN = (Nout <= 64) ? 64 : 128;
  if (in_stride != 1 && !(BN == 128 && (g_persistent == 1 || g_persistent == 4 || g_persistent == 5) && !g_use_ss)) return DP_ERR_UNSUPPORTED;   // only these kernels scale the tile origin
  {
    const cuuint64_t Kg4 = (cuuint64_t)ldb;   // dp_pack_conv_weight_tc pads rows to 16 B
    cuuint64_t dims[3] = {Kg4, (cuuint64_t)Nout, (cuuint64_t)T};
    cuuint64_t str[2] = {Kg4 * 4, (cuuint64_t)Nout * Kg4 * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
    if (!make_map(&mBh, w_hi, 3, dims, str, box) || !make_map(&mBl, w_lo, 3, dims, str, box)) return DP_ERR_UNSUPPORTED;
  }
  TcParams p{};
  p.Nimg = Nimg; p.H = H; p.W = W; p.Nout = Nout; p.R = 0; p.S = 0; p.pad = 0; p.flip = 0;
  p.ntaps = taps.n;
  for (int i = 0; i < 9; ++i) { p.dh[i] = taps.dh[i]; p.dw[i] = taps.dw[i]; p.wt[i] = taps.wt[i]; }
  p.os = os; p.oa = oa; p.ob = ob; p.Ho = Ho; p.Wo = Wo;
  p.alpha = alpha; p.b_from_img = b_from_img; p.trace = g_trace; p.in_stride = in_stride;
  { static const int skip = getenv("DPB200_TC_DEBUG_SKIP") ? atoi(getenv("DPB200_TC_DEBUG_SKIP")) : 0; p.dbg_skip = skip; }
  { static const int issuers = getenv("DPB200_TC_ISSUERS") ? atoi(getenv("DPB200_TC_ISSUERS")) : 2; p.two_issuers = issuers >= 2 ? 2 : 1; }
  { static const int wn = getenv("DPB200_TC_WIDE_N") ? atoi(getenv("DPB200_TC_WIDE_N")) : 1; p.wide_n = wn; }
  if ((alpha != 1.0f || b_from_img) && !(g_persistent && !g_use_ss && Nout > 64 && bn == 1)) return DP_ERR_UNSUPPORTED;   // ps and ab kernels apply alpha / image-indexed B
  p.kchunks = (Kg + BK - 1) / BK;
  p.bw = bw; p.bh = bh; p.bn = bn; p.tiles_w = W / bw; p.tiles_h = H / bh;
  p.y = out; p.ldy = ld_out; p.bias = bias; p.rowadd = rowadd; p.ld_rowadd = ld_rowadd; p.residual = residual; p.ld_res = ld_res;
  p.accumulate = accumulate;
  auto al16 = [](const void* q, long long ld) { return q == nullptr || ((((uintptr_t)q) & 15) == 0 && (ld % 4) == 0); };
  p.vec4 = (al16(out, ld_out) && al16(bias, 0) && al16(rowadd, ld_rowadd) && al16(residual, ld_res)) ? 1 : 0;
  const int tiles_n = (Nimg + bn - 1) / bn;
  dim3 grid((unsigned)(p.tiles_w * p.tiles_h * tiles_n), (unsigned)((Nout + BN - 1) / BN));
  if (g_use_ss) {
    if (BN == 64) conv_tc_kernel<64><<<grid, NTHREADS, STAGES * (2 * A_BYTES + 2 * 64 * BK * 4) + 2048, st>>>(mA, mBh, mBl, p);
    else conv_tc_kernel<128><<<grid, NTHREADS, STAGES * (2 * A_BYTES + 2 * 128 * BK * 4) + 2048, st>>>(mA, mBh, mBl, p);
  } else if (BN == 128 && g_persistent == 2) {
    static const int bsub = getenv("DPB200_B_SUB") ? atoi(getenv("DPB200_B_SUB")) : 1;
    if (bsub > 1) {
      const cuuint64_t Kg4 = (cuuint64_t)ldb;
      cuuint64_t dims[3] = {Kg4, (cuuint64_t)Nout, (cuuint64_t)T};
      cuuint64_t str[2] = {Kg4 * 4, (cuuint64_t)Nout * Kg4 * 4};
      cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(128 / bsub), 1};
      if (!make_map(&mBh, w_hi, 3, dims, str, box) || !make_map(&mBl, w_lo, 3, dims, str, box)) return DP_ERR_UNSUPPORTED;
      p.b_sub = bsub;
    }
    conv_tc_ab_kernel<<<grid, AB_THREADS, AB_SA * A_BYTES + AB_SB * 2 * 128 * BK * 4 + 2048, st>>>(mA, mBh, mBl, p);
  } else if (BN == 128 && (g_persistent == 4 || (g_persistent == 5 && p.ntaps * p.kchunks >= g_pt_min_stages))) {   // 5 = per layer: long-K tiles on the TS kernel
    const int tiles_m = (int)grid.x, total = (int)(grid.x * grid.y);
    const int ctas = total < g_num_sms ? total : g_num_sms;
    conv_tc_pt_kernel<<<ctas, PT_THREADS, PT_SA * A_BYTES + PT_SB * 2 * 128 * BK * 4 + 2048, st>>>(mA, mBh, mBl, p, tiles_m, total);
  } else if (BN == 128 && g_persistent == 3) {
    const int tiles_m = (int)grid.x, total = (int)(grid.x * grid.y);
    const int ctas = total < g_num_sms ? total : g_num_sms;
    if (g_ps_bk == 16) {   // 16-float K chunks: 64-byte-swizzled boxes
      {
        cuuint64_t dims[4] = {(cuuint64_t)Kg, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
        cuuint64_t str[3] = {(cuuint64_t)ld_act * 4, (cuuint64_t)W * ld_act * 4, (cuuint64_t)H * W * ld_act * 4};
        cuuint32_t box[4] = {16, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
        if (!make_map(&mA, act, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return DP_ERR_UNSUPPORTED;
        const cuuint64_t Kg4 = (cuuint64_t)ldb;
        cuuint64_t bdims[3] = {Kg4, (cuuint64_t)Nout, (cuuint64_t)T};
        cuuint64_t bstr[2] = {Kg4 * 4, (cuuint64_t)Nout * Kg4 * 4};
        cuuint32_t bbox[3] = {16, 128, 1};
        if (!make_map(&mBh, w_hi, 3, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !make_map(&mBl, w_lo, 3, bdims, bstr, bbox, CU_TENSOR_MAP_SWIZZLE_64B)) return DP_ERR_UNSUPPORTED;
      }
      p.kchunks = (Kg + 15) / 16;
      conv_tc_ps2_kernel<16><<<ctas, PS2_THREADS, ps_smem_bytes(16), st>>>(mA, mBh, mBl, p, tiles_m, total);
    } else {
      conv_tc_ps2_kernel<32><<<ctas, PS2_THREADS, ps_smem_bytes(32), st>>>(mA, mBh, mBl, p, tiles_m, total);
    }
  } else if (BN == 128 && g_persistent) {
    const int tiles_m = (int)grid.x, total = (int)(grid.x * grid.y);
    const int ctas = total < g_num_sms ? total : g_num_sms;
    conv_tc_ps_kernel<<<ctas, PS_THREADS, PS_STAGES * (2 * A_BYTES + 2 * 128 * BK * 4) + 2048, st>>>(mA, mBh, mBl, p, tiles_m, total);
  } else if (BN == 64) {
    conv_tc_ts_kernel<64, 1><<<grid, NTHREADS, STAGES_TS * (A_BYTES + 2 * 64 * BK * 4) + 2048, st>>>(mA, mBh, mBl, p);
  } else {
    int cl = g_cluster;
    while (cl > 1 && (grid.x % cl)) cl >>= 1;
    const size_t smem = STAGES_TS * (A_BYTES + 2 * 128 * BK * 4) + 2048;
    if (cl <= 1) {
      conv_tc_ts_kernel<128, 1><<<grid, NTHREADS, smem, st>>>(mA, mBh, mBl, p);
    } else {
      // weight-tile slices of 128/cl rows per CTA need their own (smaller-box) tensor maps
      CUtensorMap sBh, sBl;
      const cuuint64_t Kg4 = (cuuint64_t)ldb;
      cuuint64_t dims[3] = {Kg4, (cuuint64_t)Nout, (cuuint64_t)T};
      cuuint64_t str[2] = {Kg4 * 4, (cuuint64_t)Nout * Kg4 * 4};
      cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(128 / cl), 1};
      if (!make_map(&sBh, w_hi, 3, dims, str, box) || !make_map(&sBl, w_lo, 3, dims, str, box)) return DP_ERR_UNSUPPORTED;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = grid; cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaError_t e = (cl == 2) ? cudaLaunchKernelEx(&cfg, conv_tc_ts_kernel<128, 2>, mA, sBh, sBl, p)
                                : cudaLaunchKernelEx(&cfg, conv_tc_ts_kernel<128, 4>, mA, sBh, sBl, p);
      if (e != cudaSuccess) { g_dp_last_cuda_error = (int)e; (void)cudaGetLastError(); return DP_ERR_CUDA; }
    }
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

namespace {
}  // namespace

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
                   a->ld_rowadd, a->residual, a->ld_res, (a->flags & DP_CONV_ACCUMULATE) ? 1 : 0, (cudaStream_t)stream, 1.0f, 0, a->stride);
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
                     acc, (cudaStream_t)stream);
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
                         a->H, a->W, (float*)a->x, a->ldx, nullptr, nullptr, 0, nullptr, 0, acc, (cudaStream_t)stream);
      if (rc != DP_OK) return (ca == 0 && cb == 0) ? rc : (rc == DP_ERR_UNSUPPORTED ? DP_ERR_SHAPE : rc);
    }
  return DP_OK;
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
  { static const int issuers = getenv("DPB200_TC_ISSUERS") ? atoi(getenv("DPB200_TC_ISSUERS")) : 2; p.two_issuers = issuers >= 2; }
  { static const int wn = getenv("DPB200_TC_WIDE_N") ? atoi(getenv("DPB200_TC_WIDE_N")) : 1; p.wide_n = wn; }
  wgrad_tc_kernel<<<grid, WG_THREADS, 4 * 3 * WG_T + 2048, (cudaStream_t)stream>>>(mDy, mX, p);
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

// Debug hook: CTA 0 of every following persistent conv launch stamps clock64() into buf[stage*8 + slot] for its first 1024 pipeline
// stages (conv_tc_ps2_kernel / conv_tc_pt_kernel; slots: 0 producer woke on "empty", 1 TMA issued, 2 splitter woke on "full", 6 split done, 3 splitter arrived,
// 4 MMA lane woke on "converted", 5 MMAs + commit issued).  nullptr switches it off.  buf must hold 16640 int64 (16 slots per stage; 8..10: issuer at loop top / after the "converted" wait / after the hand-off wait) (the last 256: per tile, epilogue woke / released TMEM / done, issuer got the accumulator).
extern "C" int dp_conv_tc_set_trace(long long* buf) { g_trace = buf; return 0; }

extern "C" int dp_tc_weight_row(int channels) { return channels > 0 ? wrow(channels) : 0; }
