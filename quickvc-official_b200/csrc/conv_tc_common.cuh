// conv_tc_common.cuh -- PTX wrappers and epilogue building blocks shared by the tcgen05 series-convolution
// kernels (conv_tc.cu: one CTA per tile, cta_group::1; conv_tc2.cu: CTA pairs, cta_group::2).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace qvc {

namespace tc {

constexpr int CHUNK_M = 128;              // output channels per CTA and MMA (TMEM lanes)
constexpr int ROW_BYTES = 128;            // one swizzle-128B row = one K block of one frame / filter row
constexpr int CHUNK_BYTES = CHUNK_M * ROW_BYTES;
constexpr int MAX_SMEM = 232448;          // 227 KB
// 8 epilogue warps (two per TMEM lane quarter).  Tried and dropped (profiles/r02_summary.md): 16 epilogue warps on 32-frame
// superblocks (576 threads, 96 registers each, 50-370 bytes of spills): +5 % step time in tf32, +1.5 % in fp16, +7 % on
// a fused WN layer.
constexpr int NTHREADS = 320;
constexpr int N_EPI_WARPS = 8;
constexpr int ACC_COLS = 256;             // TMEM columns per accumulator set

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  // bounded spin: a protocol bug must trap, not hang the GPU
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// One lane of a converged warp (the pattern ptxas recognises: no per-instruction election loop around
// the warp-level tcgen05 instructions, unlike `if (lane == 0)`).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// operand format field of the instruction descriptor (a_format / b_format): kind::f16 takes 0 = F16, 1 = BF16;
// kind::tf32 takes 2 = TF32
constexpr uint32_t mma_format(int opf) { return opf == QVC_OPF_F16 ? 0u : (opf == QVC_OPF_BF16 ? 1u : 2u); }

template <int OPF>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (opf_is16(OPF)) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// Same MMA with the A operand read from TMEM (M lanes x K 32-bit columns) instead of shared memory.
template <int OPF>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (opf_is16(OPF)) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// K-major, 128-byte-swizzled operand: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
// Measured on B200 (profiles/r01_tc_desc_mode.log): the 128B swizzle is a function of the absolute
// shared-memory address, so a descriptor may start on any 128-byte row of a TMA-written slab with
// base_offset 0 -- which is what lets one slab serve all k taps.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

// ---- thread-block clusters / CTA pairs (cta_group::2): shared by conv_tc2.cu and conv_wn.cu ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Remote arrive WITHOUT release semantics: for barriers that only say "my tcgen05.ld reads of this accumulator are done"
// (completed by tcgen05.wait::ld, ordered by tcgen05.fence::before_thread_sync) and publish no memory.  The release form
// compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. it waits until every global store of the tile just finished has drained:
// 15 % of all warp-stall samples of an epilogue-bound layer sat there (profiles/r02_summary.md).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// TMA loads of a CTA pair: data lands in the issuing CTA, the bytes are counted on the barrier at `bar`
// (a shared::cluster address -- the leader's)
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the barrier at the same offset in both CTAs once all prior MMAs of this thread have retired
__device__ __forceinline__ void tc2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
template <int OPF>
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (opf_is16(OPF)) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// 32 consecutive TMEM columns of this thread's lane; completion via tmem_wait().
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// exp / sigmoid / tanh on the SFU (ex2.approx, rcp.approx): absolute error of a few 1e-7 on the gate
// output, two orders below the TF32 rounding of the operand it becomes.  The exact-fp32 FMA back end
// keeps expf / tanhf.
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.f + fast_exp(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - 2.f * fast_rcp(1.f + fast_exp(2.f * x)); }

// ----------------------------------------------------------------------------------------------
// epilogues: one thread = one output channel, 32 consecutive frames t .. t+31 (nv of them live)
// ----------------------------------------------------------------------------------------------
// LINEAR epilogue on a "superblock" of up to 64 frames: the global loads of the whole superblock are
// issued first (and, for the first superblock of a tile, before the accumulator is even complete), then the
// two 32-frame halves are read out of TMEM, finished and stored.  Measured (scripts/micro/membench.cu): with
// 8 warps per SM a 32-line-deep load burst per warp sustains 3.0 TB/s, a 64-line-deep one 5.3 TB/s.
struct LinCtx {
  const EpiSeg* sg;
  int b, c;                 // utterance, channel within the segment
  bool ok;                  // this lane owns a live channel
  bool all_ok;              // ... and so does every lane of the warp
  float bias;
  int lim;                  // live rows of utterance b (ragged batches): operand rows >= lim are written as zero
};

// loads only: the residual if the segment has one (fp32, or the operand-format copy -- kept as raw bits here, decoded
// in lin_finish so that nothing waits for the loads), else the accumulate-into tensor (else nothing)
template <int OPF>
__device__ __forceinline__ void lin_load(const LinCtx& k, int t, int nv, float* r) {
  const EpiSeg& sg = *k.sg;
  if (sg.res_op.present()) {
    using OT = typename OpType<OPF>::type;
    const OT* rp = sg.res_op.at<OT>(k.b, t, k.c);
    const int ld = sg.res_op.ld;
    if constexpr (opf_is16(OPF)) {
      const uint16_t* rp16 = reinterpret_cast<const uint16_t*>(rp);
      if (k.all_ok && nv == 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) r[i] = __uint_as_float((uint32_t)rp16[i * ld]);
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) r[i] = (k.ok && i < nv) ? __uint_as_float((uint32_t)rp16[i * ld]) : 0.f;
      }
    } else {
      const float* rpf = reinterpret_cast<const float*>(rp);
      if (k.all_ok && nv == 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) r[i] = rpf[i * ld];
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) r[i] = (k.ok && i < nv) ? rpf[i * ld] : 0.f;
      }
    }
    return;
  }
  const TRef& src = sg.res.present() ? sg.res : sg.accin;
  if (!src.present()) return;
  const float* rp = src.at<float>(k.b, t, k.c);
  const int ld = src.ld;
  if (k.all_ok && nv == 64) {
    // common case, no per-element predicate: one IMAD.WIDE + LDG per element (the predicated form costs
    // ~9 instructions per load and made this epilogue issue-bound, profiles/r01_v2_summary.md)
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = rp[i * ld];
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = (k.ok && i < nv) ? rp[i * ld] : 0.f;
  }
}

template <int OPF>
__device__ __forceinline__ void lin_finish(const LinCtx& k, int t, int nv, float* r, uint32_t taddr) {
  using OT = typename OpType<OPF>::type;
  const EpiSeg& sg = *k.sg;
  const bool res_from_op = sg.res_op.present();
  const bool has_res = sg.res.present() || res_from_op, has_acc = sg.accin.present();
  const float alpha = sg.alpha, beta = sg.beta, slope = sg.slope;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int th = t + 32 * h, nvh = nv - 32 * h;
    if (nvh <= 0) break;
    float v[32];
    tmem_ld32(taddr + 32 * h, v);
    tmem_wait();
    float* rr = r + 32 * h;
    if (res_from_op) {                               // decode the operand copy in place: undo the leaky-relu
      const float inv = sg.res_inv_slope;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = rr[i];
        if constexpr (opf_is16(OPF)) x = op16_to_float<OPF>(__float_as_uint(x));
        rr[i] = x > 0.f ? x : x * inv;
      }
    }
    if (has_res && has_acc) {
      // both streams (last convolution of MRF blocks 2 and 3): the accumulate-into tensor is loaded late, 8 at a time
      const float* ap = sg.accin.at<float>(k.b, th, k.c);
      const int64_t ld = sg.accin.ld;
#pragma unroll
      for (int g = 0; g < 32; g += 8) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (k.ok && g + i < nvh) ? ap[(g + i) * ld] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[g + i] = fmaf(beta, fmaf(alpha, v[g + i] + k.bias, rr[g + i]), a[i]);
      }
    } else if (has_res) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = beta * fmaf(alpha, v[i] + k.bias, rr[i]);
    } else if (has_acc) {
      const float ab = alpha * beta;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaf(ab, v[i] + k.bias, rr[i]);
    } else {
      const float ab = alpha * beta;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ab * (v[i] + k.bias);
    }
    const int nlive = k.lim - th;                    // rows of this block inside the utterance's own length
    const bool full = k.all_ok && nvh >= 32;
    if (sg.raw.present()) {
      float* wp = sg.raw.at<float>(k.b, th, k.c);
      const int ld = sg.raw.ld;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) wp[i * ld] = v[i];
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k.ok && i < nvh) wp[i * ld] = v[i];
      }
    }
    if (sg.op.present()) {
      OT* op = sg.op.at<OT>(k.b, th, k.c);
      const int ld = sg.op.ld;
      if (full && nlive >= 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) op[i * ld] = to_operand<OPF>(fmaxf(v[i], v[i] * slope));   // leaky-relu, slope <= 1
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k.ok && i < nvh) op[i * ld] = to_operand<OPF>(i < nlive ? leaky(v[i], slope) : 0.f);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Sum of convolutions (qvc_conv1d_sum): one segment, up to three residual sources, blocks of 32 frames.
//   v = beta * (alpha * (acc + bias) + res_0 + res_1 + res_2)        (bias = sum of the sources' biases, in LinCtx)
// sum_load only issues the loads of a block (raw bits, decoded in sum_finish), so the first block of a tile is in
// flight before the accumulator is complete.
// ----------------------------------------------------------------------------------------------
template <int OPF>
__device__ __forceinline__ void sum_load(const EpiParams& ep, const LinCtx& k, int t, int nv, float* r) {
  using OT = typename OpType<OPF>::type;
  const EpiSeg& sg = *k.sg;
  const bool full = k.all_ok && nv >= 32;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s >= ep.nsum) break;
    const TRef& ro = s == 0 ? sg.res_op : ep.xres_op[s - 1];
    const TRef& rf = s == 0 ? sg.res : ep.xres[s - 1];
    float* rr = r + 32 * s;
    if (ro.present()) {
      const OT* rp = ro.at<OT>(k.b, t, k.c);
      const int ld = ro.ld;
      if constexpr (opf_is16(OPF)) {
        const uint16_t* rp16 = reinterpret_cast<const uint16_t*>(rp);
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = __uint_as_float((uint32_t)rp16[i * ld]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = (k.ok && i < nv) ? __uint_as_float((uint32_t)rp16[i * ld]) : 0.f;
        }
      } else {
        const float* rpf = reinterpret_cast<const float*>(rp);
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = rpf[i * ld];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = (k.ok && i < nv) ? rpf[i * ld] : 0.f;
        }
      }
    } else if (rf.present()) {
      const float* rp = rf.at<float>(k.b, t, k.c);
      const int ld = rf.ld;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) rr[i] = rp[i * ld];
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) rr[i] = (k.ok && i < nv) ? rp[i * ld] : 0.f;
      }
    }
  }
}

template <int OPF>
__device__ __forceinline__ void sum_finish(const EpiParams& ep, const LinCtx& k, int t, int nv, float* r, uint32_t taddr) {
  using OT = typename OpType<OPF>::type;
  const EpiSeg& sg = *k.sg;
  const bool from_op = sg.res_op.present();
  const bool has_res = from_op || sg.res.present();
  float v[32];
  tmem_ld32(taddr, v);
  tmem_wait();
  if (from_op) {                                   // operand copies hold leaky_relu(x): undo it
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (s >= ep.nsum) break;
      const float inv = s == 0 ? sg.res_inv_slope : ep.xres_inv_slope[s - 1];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = r[32 * s + i];
        if constexpr (opf_is16(OPF)) x = op16_to_float<OPF>(__float_as_uint(x));
        r[32 * s + i] = x > 0.f ? x : x * inv;
      }
    }
  }
  const float alpha = sg.alpha, beta = sg.beta, slope = sg.slope;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float a = has_res ? fmaf(alpha, v[i] + k.bias, r[i]) : alpha * (v[i] + k.bias);
    if (has_res && ep.nsum > 1) a += r[32 + i];
    if (has_res && ep.nsum > 2) a += r[64 + i];
    v[i] = beta * a;
  }
  const int nlive = k.lim - t;
  const bool full = k.all_ok && nv >= 32;
  if (sg.raw.present()) {
    float* wp = sg.raw.at<float>(k.b, t, k.c);
    const int ld = sg.raw.ld;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; ++i) wp[i * ld] = v[i];
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (k.ok && i < nv) wp[i * ld] = v[i];
    }
  }
  if (sg.op.present()) {
    OT* op = sg.op.at<OT>(k.b, t, k.c);
    const int ld = sg.op.ld;
    if (full && nlive >= 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) op[i * ld] = to_operand<OPF>(fmaxf(v[i], v[i] * slope));
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (k.ok && i < nv) op[i * ld] = to_operand<OPF>(i < nlive ? leaky(v[i], slope) : 0.f);
    }
  }
}

// bias of a sum of convolutions: the sources' bias vectors added in source order
__device__ __forceinline__ float sum_bias(const EpiParams& ep, int b, int n, bool ok) {
  float bias = (ep.bias && ok) ? ep.bias[(int64_t)b * ep.bias_bs + n] : 0.f;
  if (ep.nsum > 1 && ep.xbias[0] && ok) bias += ep.xbias[0][n];
  if (ep.nsum > 2 && ep.xbias[1] && ok) bias += ep.xbias[1][n];
  return bias;
}

// tanh(x) * sigmoid(y) with two ex2 and ONE rcp:  (e^{2x} - 1) / ((e^{2x} + 1) (1 + e^{-y})).
// x is clamped to 15 (tanh(15) = 1 - 2e-13) so that e^{2x} stays finite; a huge e^{-y} makes the
// denominator inf and the quotient 0, the correct limit.  Absolute error a few 1e-7.
__device__ __forceinline__ float fast_gate(float x, float y) {
  float e2x, emy;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2x) : "f"(fminf(x, 15.f) * 2.8853900817779268f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(emy) : "f"(y * -1.4426950408889634f));
  return (e2x - 1.f) * fast_rcp((e2x + 1.f) * (1.f + emy));
}

template <int OPF>
__device__ __forceinline__ void epi_gate_cols(const EpiParams& ep, int b, int t, int nv, int nlive, int n, bool ok, bool all_ok,
                                              float bias_lo, float bias_hi, uint32_t taddr_lo, uint32_t taddr_hi) {
  using OT = typename OpType<OPF>::type;
  float lo[32], hi[32];
  tmem_ld32(taddr_lo, lo);
  tmem_ld32(taddr_hi, hi);
  tmem_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) lo[i] = fast_gate(lo[i] + bias_lo, hi[i] + bias_hi);
  const EpiSeg& sg = ep.seg[0];
  const bool full = all_ok && nv >= 32;
  if (sg.raw.present()) {
    float* wp = sg.raw.at<float>(b, t, n);
    const int ld = sg.raw.ld;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (ok && i < nv) wp[i * ld] = lo[i];
  }
  if (sg.op.present()) {
    OT* op = sg.op.at<OT>(b, t, n);
    const int ld = sg.op.ld;
    if (full && nlive >= 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) op[i * ld] = to_operand<OPF>(lo[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (ok && i < nv) op[i * ld] = to_operand<OPF>(i < nlive ? lo[i] : 0.f);
    }
  }
}

template <int OPF>
__device__ __forceinline__ void epi_sample_cols(const EpiParams& ep, int b, int t, int nv, int nlive, int n, bool ok,
                                                float bias_lo, float bias_hi, uint32_t taddr_lo, uint32_t taddr_hi) {
  using OT = typename OpType<OPF>::type;
  float m[32], lg[32], nz[32];
  {
    const float* np = ep.noise.at<float>(b, t, n);
    const int64_t ld = ep.noise.ld;
#pragma unroll
    for (int i = 0; i < 32; ++i) nz[i] = (ok && i < nv) ? np[i * ld] : 0.f;
  }
  tmem_ld32(taddr_lo, m);
  tmem_ld32(taddr_hi, lg);
  tmem_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    m[i] += bias_lo;
    lg[i] += bias_hi;
    nz[i] = fmaf(nz[i], fast_exp(lg[i]), m[i]);        // z = m + noise * exp(logs)   (models.py:94)
  }
  if (ep.aux0.present()) {
    float* wp = ep.aux0.at<float>(b, t, n);
    const int64_t ld = ep.aux0.ld;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (ok && i < nv) wp[i * ld] = m[i];
  }
  if (ep.aux1.present()) {
    float* wp = ep.aux1.at<float>(b, t, n);
    const int64_t ld = ep.aux1.ld;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (ok && i < nv) wp[i * ld] = lg[i];
  }
  const EpiSeg& sg = ep.seg[0];
  if (sg.raw.present()) {
    float* wp = sg.raw.at<float>(b, t, n);
    const int64_t ld = sg.raw.ld;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (ok && i < nv) wp[i * ld] = nz[i];
  }
  if (sg.op.present()) {
    OT* op = sg.op.at<OT>(b, t, n);
    const int64_t ld = sg.op.ld;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (ok && i < nv) op[i * ld] = to_operand<OPF>(i < nlive ? nz[i] : 0.f);
  }
}


}  // namespace tc

// host helpers defined in conv_tc.cu
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_get_encode();
int tc_sm_count();                  // SMs available to a persistent convolution grid (minus tc_reserve_sms)
void tc_reserve_sms(int n);         // per host thread; see conv_tc.cu
int tc_env_int(const char* name, int dflt);
bool tc_prof_next(cudaEvent_t* e0, cudaEvent_t* e1);
int launch_conv_tc2(const qvc_conv_args& a, cudaStream_t stream);     // conv_tc2.cu; QVC_ERR_UNSUPPORTED = not applicable
int launch_conv_tc2_sum(const qvc_conv_args* const* srcs, int nsrc, bool sum, cudaStream_t stream);
int launch_conv_tcr(const qvc_conv_args& a, cudaStream_t stream);   // conv_tcr.cu (frames on the accumulator rows); QVC_ERR_UNSUPPORTED = not applicable

}  // namespace qvc
