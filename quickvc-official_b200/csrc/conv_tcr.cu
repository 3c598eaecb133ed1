// conv_tcr.cu -- series convolution on CTA pairs with FRAMES on the accumulator rows (TMEM lanes) and output
// channels on its columns: the transpose of conv_tc2.cu.  Built for the 192-channel WN stacks of the prior encoder
// and the flow (modules.py:88-112), where the channel-major kernels were bound by everything but the tensor pipe:
//
//   * A = activation slab (M = 2 x 128 frames, one half per CTA, taps = descriptors shifted by j * dil rows),
//     B = filter rows (N = one "piece" of up to 256 output columns, half of them loaded by each CTA).  Per MMA and CTA
//     that is 4 KB of A + 16 N bytes of B for N / 2 cycles: N = 192 reads 7 KB per 96 cycles, where the channel-major
//     pair kernel at 128 frames per tile read 6 KB per 64 cycles (capped near 2/3 of the tensor rate, DESIGN.md
//     section 4) and needed all 512 TMEM columns for one tile of a gate layer.
//   * gate layers: piece p = [tanh rows of channels 96 p .. 96 p + 95 | sigmoid rows of the same channels]: both
//     members of a gate pair sit in the same TMEM lane (same frame) 96 columns apart, two pieces double-buffer in
//     2 x 192 columns, so the gate epilogue of one piece runs under the MMAs of the next.
//   * an epilogue thread owns one FRAME and walks channels: every global access is a 32-byte sector of one row
//     (256-bit LDG / STG), 8 instructions per 32 fp32 values instead of 32.
#include <cstdlib>

#include "conv_tc_common.cuh"

namespace qvc {

namespace {

using namespace tc;

constexpr int MAXPIECES = 8;
constexpr int TM = 128;                   // frames per CTA = TMEM lanes; a pair tile spans 2 TM frames
constexpr int BUF_COLS = 256;             // TMEM columns between the two accumulator sets

struct alignas(64) TcrParams {
  CUtensorMap mx;                          // x as (channel, frame, utterance)
  CUtensorMap mw;                          // w as (tap*cin + channel, output channel); box = np / 2 rows
  int32_t cin, k, dil, pad_left;
  int32_t np;                              // accumulator columns per piece (the N of the MMA)
  int32_t npieces;
  int32_t wrow[MAXPIECES][2];              // first filter row CTA r loads for piece p
  int32_t n0[MAXPIECES];                   // LINEAR: output column of accumulator column 0; GATE: gate channel of column 0
  int32_t ntb, ntiles;
  int32_t slab_box_rows, slab_stages, w_stages;
  uint32_t slab_stage_bytes, w_stage_bytes;
  int32_t batch;
  EpiParams ep;
};

// 256-bit global accesses (sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void ldg256(const void* p, uint32_t* v) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 32 consecutive fp32 values of one row
__device__ __forceinline__ void load_row32(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) ldg256(p + 8 * i, reinterpret_cast<uint32_t*>(v) + 8 * i);
}
__device__ __forceinline__ void store_row32(float* p, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) stg256(p + 8 * i, reinterpret_cast<const uint32_t*>(v) + 8 * i);
}
// 32 consecutive operand values of one row
template <int OPF>
__device__ __forceinline__ void store_op32(typename OpType<OPF>::type* p, const float* v) {
  if constexpr (opf_is16(OPF)) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = op16_pack2<OPF>(v[2 * i], v[2 * i + 1]);
    stg256(p, w);
    stg256(p + 16, w + 8);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(to_operand<OPF>(v[8 * i + j]));
      stg256(p + 8 * i, w);
    }
  }
}

// ---- LINEAR epilogue of one 32-column block of one row (same arithmetic as tc::lin_finish) ----
struct RowSeg {
  const EpiSeg* sg;
  int c;                    // channel of the block's first column within the segment
  bool in;                  // the block lies inside the segment
};
__device__ __forceinline__ RowSeg row_seg(const EpiParams& ep, int n) {
  RowSeg s;
  s.sg = &ep.seg[(ep.nseg > 1 && n >= ep.seg[1].col0) ? 1 : 0];
  s.c = n - s.sg->col0;
  s.in = s.c >= 0 && s.c + 32 <= s.sg->ncols;
  return s;
}
// the fp32 stream a block adds first: the residual if there is one, else the accumulate-into tensor
__device__ __forceinline__ void row_lin_load(const RowSeg& s, int b, int t, bool ok, float* r) {
  const TRef& src = s.sg->res.present() ? s.sg->res : s.sg->accin;
  if (ok && s.in && src.present()) load_row32(src.at<float>(b, t, s.c), r);
}
// 16 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// Finishes one 32-column block of one row in two halves of 16 columns (the accumulator is read 16 columns at a time so
// that three blocks of prefetched residual fit the register file beside it).  All lanes execute the TMEM loads.
template <int OPF>
__device__ __forceinline__ void row_lin_finish(const EpiParams& ep, const RowSeg& s, int b, int t, int n, bool ok, bool live,
                                               uint32_t taddr, const float* r) {
  using OT = typename OpType<OPF>::type;
  const EpiSeg& sg = *s.sg;
  const bool act = ok && s.in;
  const float alpha = sg.alpha, beta = sg.beta, slope = sg.slope;
  const bool has_res = sg.res.present(), has_acc = sg.accin.present();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
    tmem_ld16(taddr + 16 * h, v);
    tmem_wait();
    if (!act) continue;
    const float* rr = r + 16 * h;
    if (ep.bias) {
      const float4* bp = reinterpret_cast<const float4*>(ep.bias + (int64_t)b * ep.bias_bs + n + 16 * h);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = __ldg(bp + i);
        v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
      }
    }
    if (has_res && has_acc) {
      const float* ap = sg.accin.at<float>(b, t, s.c + 16 * h);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float a[8];
        ldg256(ap + 8 * g, reinterpret_cast<uint32_t*>(a));
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * g + i] = fmaf(beta, fmaf(alpha, v[8 * g + i], rr[8 * g + i]), a[i]);
      }
    } else if (has_res) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = beta * fmaf(alpha, v[i], rr[i]);
    } else if (has_acc) {
      const float ab = alpha * beta;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(ab, v[i], rr[i]);
    } else {
      const float ab = alpha * beta;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ab * v[i];
    }
    if (sg.raw.present()) {
      float* wp = sg.raw.at<float>(b, t, s.c + 16 * h);
      stg256(wp, reinterpret_cast<const uint32_t*>(v));
      stg256(wp + 8, reinterpret_cast<const uint32_t*>(v) + 8);
    }
    if (sg.op.present()) {
      if (live) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * slope);      // leaky-relu, slope <= 1
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      OT* op = sg.op.at<OT>(b, t, s.c + 16 * h);
      if constexpr (opf_is16(OPF)) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = op16_pack2<OPF>(v[2 * i], v[2 * i + 1]);
        stg256(op, w);
      } else {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(to_operand<OPF>(v[i]));
        stg256(op, w);
        stg256(op + 8, w + 8);
      }
    }
  }
}

template <int OPF, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) conv_tcr_kernel(const __grid_constant__ TcrParams p) {
  constexpr int ESIZE = opf_is16(OPF) ? 2 : 4;
  constexpr int KC = ROW_BYTES / ESIZE;
  constexpr uint32_t FMT = mma_format(OPF);
  using OT = typename OpType<OPF>::type;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slab0 = smem_base;
  const uint32_t w0 = slab0 + p.slab_stages * p.slab_stage_bytes;
  const uint32_t bar0 = w0 + p.w_stages * p.w_stage_bytes;
  // barriers (same offsets in both CTAs): full_slab[SS] empty_slab[SS] full_w[WS] empty_w[WS] tmem_full[2]
  // tmem_empty[2], then the TMEM base word.  full_* and tmem_empty are only used in the leader.
  const uint32_t full_slab = bar0, empty_slab = full_slab + 8 * p.slab_stages;
  const uint32_t full_w = empty_slab + 8 * p.slab_stages, empty_w = full_w + 8 * p.w_stages;
  const uint32_t tmem_full = empty_w + 8 * p.w_stages, tmem_empty = tmem_full + 16;
  const uint32_t tmem_slot = tmem_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  int32_t* live_s = reinterpret_cast<int32_t*>(tmem_slot_ptr + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * p.slab_stages + 2 * p.w_stages + 2; ++i) mbar_init(bar0 + 8 * i, 1);
    // an accumulator set is drained by ONE group of four epilogue warps per CTA (group h <-> set h)
    mbar_init(tmem_empty, 2 * (N_EPI_WARPS / 2));
    mbar_init(tmem_empty + 8, 2 * (N_EPI_WARPS / 2));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BUF_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.ep.live != nullptr) {
    load_live_cache(p.ep, p.batch, live_s);
    __syncthreads();
  }
  const int n_cchunks = p.cin / KC;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mx) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mw) : "memory");
      const uint32_t lead_full_slab = map_to_cta(full_slab, 0), lead_full_w = map_to_cta(full_w, 0);
      const uint32_t slab_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
      const int cin = p.cin, k = p.k, pad_left = p.pad_left;
      uint32_t s = 0, ph = 0, ws = 0, wph = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs) {
        const int piece = tile % p.npieces;
        const int rest = tile / p.npieces;
        const int tb = rest % p.ntb, b = rest / p.ntb;
        if (tile_dead(p.ep, live_s, b, tb * 2 * TM)) continue;
        const int t0 = tb * 2 * TM + (int)rank * TM;
        const int wrow = p.wrow[piece][rank];
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(empty_slab + 8 * s, ph ^ 1u);
          if (leader) mbar_expect_tx(full_slab + 8 * s, 2 * slab_bytes);
          tma2_load_3d(slab0 + s * p.slab_stage_bytes, &p.mx, lead_full_slab + 8 * s, cc * KC, t0 - pad_left, b);
          if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
          for (int j = 0; j < k; ++j) {
            mbar_wait(empty_w + 8 * ws, wph ^ 1u);
            if (leader) mbar_expect_tx(full_w + 8 * ws, 2 * p.w_stage_bytes);
            tma2_load_2d(w0 + ws * p.w_stage_bytes, &p.mw, lead_full_w + 8 * ws, j * cin + cc * KC, wrow);
            if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.np >> 3) << 17) |
                             ((uint32_t)((2 * TM) >> 4) << 24);
      const uint64_t desc_hi = smem_desc(0);
      const int k = p.k, dil = p.dil;
      uint32_t s = 0, ph = 0, ws = 0, wph = 0, ait = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs) {
        {
          const int rest = tile / p.npieces;
          if (tile_dead(p.ep, live_s, rest / p.ntb, (rest % p.ntb) * 2 * TM)) continue;
        }
        const uint32_t buf = ait & 1u, bph = (ait >> 1) & 1u;
        mbar_wait(tmem_empty + 8 * buf, bph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * BUF_COLS;
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(full_slab + 8 * s, ph);
          const uint32_t slab = slab0 + s * p.slab_stage_bytes;
          for (int j = 0; j < k; ++j) {
            mbar_wait(full_w + 8 * ws, wph);
            tc_fence_after();
            const uint32_t first = (cc | j) == 0 ? 0u : 1u;
            const uint64_t adesc = desc_hi | (uint64_t)(((slab + (uint32_t)(j * dil) * ROW_BYTES) & 0x3FFFFu) >> 4);
            const uint64_t bdesc = desc_hi | (uint64_t)(((w0 + ws * p.w_stage_bytes) & 0x3FFFFu) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma2<OPF>(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
              tc2_commit(empty_w + 8 * ws);
              if (j == k - 1) tc2_commit(empty_slab + 8 * s);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
          }
          if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) tc2_commit(tmem_full + 8 * buf);
        __syncwarp();
        ++ait;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs): group h drains accumulator set h =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const uint32_t h = (uint32_t)(warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lead_tmem_empty = map_to_cta(tmem_empty, 0);
    uint32_t ait = 0;
    for (int tile = pair; tile < p.ntiles; tile += npairs) {
      const int piece = tile % p.npieces;
      const int rest = tile / p.npieces;
      const int tb = rest % p.ntb, b = rest / p.ntb;
      if (tile_dead(p.ep, live_s, b, tb * 2 * TM)) continue;
      const uint32_t buf = ait & 1u, bph = (ait >> 1) & 1u;
      ++ait;
      if (buf != h) continue;
      const int t = tb * 2 * TM + (int)rank * TM + row;
      const bool ok = t < p.ep.out_rows;
      const bool live = t < live_rows(p.ep, b);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BUF_COLS;
      if constexpr (EPI == QVC_EPI_LINEAR) {
        // The fp32 stream of a block (residual / accumulate-into) is requested TWO blocks ahead of its use, the first two
        // before the accumulator is even complete: measured (profiles/r02_summary.md), one 128-byte row segment in flight
        // per thread (32 KB per SM) sustains only ~1.9 TB/s of a memory-bound layer, the latency under load being ~2.5 us.
        const int nblk = p.np >> 5;
        const int nbase = p.n0[piece];
        float r0[32], r1[32], r2[32];
        RowSeg s0 = row_seg(p.ep, nbase), s1 = row_seg(p.ep, nbase + 32), s2 = s0;
        row_lin_load(s0, b, t, ok, r0);
        if (nblk > 1) row_lin_load(s1, b, t, ok, r1);
        mbar_wait(tmem_full + 8 * buf, bph);
        tc_fence_after();
        auto step = [&](int blk, RowSeg& sc, float* rc, RowSeg& sn, float* rn) {   // finish block blk (stream rc), request blk + 2 into rn
          if (blk + 2 < nblk) {
            sn = row_seg(p.ep, nbase + 32 * (blk + 2));
            row_lin_load(sn, b, t, ok, rn);
          }
          row_lin_finish<OPF>(p.ep, sc, b, t, nbase + 32 * blk, ok, live, taddr + 32 * blk, rc);
        };
        for (int blk = 0; blk < nblk; blk += 3) {
          step(blk, s0, r0, s2, r2);
          if (blk + 1 < nblk) step(blk + 1, s1, r1, s0, r0);
          if (blk + 2 < nblk) step(blk + 2, s2, r2, s1, r1);
        }
      } else {
        const int hp = p.np >> 1;                   // gate channels of this piece: columns [0, hp) tanh, [hp, np) sigmoid
        const int ch0 = p.n0[piece];
        const float* gb = p.ep.bias + (int64_t)b * p.ep.bias_bs;
        const EpiSeg& sgm = p.ep.seg[0];
        mbar_wait(tmem_full + 8 * buf, bph);
        tc_fence_after();
        for (int blk = 0; blk < (hp >> 5); ++blk) {
          float lo[32], hi[32];
          tmem_ld32(taddr + 32 * blk, lo);
          tmem_ld32(taddr + hp + 32 * blk, hi);
          const int n = ch0 + 32 * blk;
          float bl[32], bh[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(gb + n) + i);
            const float4 c = __ldg(reinterpret_cast<const float4*>(gb + p.ep.half + n) + i);
            bl[4 * i] = a.x; bl[4 * i + 1] = a.y; bl[4 * i + 2] = a.z; bl[4 * i + 3] = a.w;
            bh[4 * i] = c.x; bh[4 * i + 1] = c.y; bh[4 * i + 2] = c.z; bh[4 * i + 3] = c.w;
          }
          tmem_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) lo[i] = fast_gate(lo[i] + bl[i], hi[i] + bh[i]);
          if (ok) {
            if (sgm.raw.present()) store_row32(sgm.raw.at<float>(b, t, n), lo);
            if (sgm.op.present()) {
              if (!live) {
#pragma unroll
                for (int i = 0; i < 32; ++i) lo[i] = 0.f;
              }
              store_op32<OPF>(sgm.op.at<OT>(b, t, n), lo);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_tmem_empty + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  // execution barrier only (the peer may still multicast into / arrive on this CTA's barriers): the release form would
  // first drain every global store of the CTA (MEMBAR.ALL.GPU: 10 % of the stall samples of a memory-bound layer)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BUF_COLS) : "memory");
  }
}

template <int OPF, int EPI>
int launch_r(const TcrParams& p, int grid, size_t smem, cudaStream_t stream) {
  static std::atomic<bool> attr_done[MAX_DEVICES];
  if (first_use_on_device(attr_done))
    QVC_CHECK_CUDA(cudaFuncSetAttribute(conv_tcr_kernel<OPF, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc_env_int("QVC_TC_PDL", 1) ? 1 : 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = tc_prof_next(&e0, &e1);
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e0, stream));
  QVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tcr_kernel<OPF, EPI>, p));
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e1, stream));
  return post_launch("conv_tcr_kernel");
}

bool aligned32(const qvc_tensor& t, size_t esize) {
  return t.ptr == nullptr || (((uintptr_t)t.ptr & 31) == 0 && ((size_t)t.ld * esize) % 32 == 0 && ((size_t)t.bstride * esize) % 32 == 0);
}

}  // namespace

// Frames-on-rows pair kernel.  Returns QVC_ERR_UNSUPPORTED (error string untouched) when the layer is not one of its
// cases.  QVC_TC_ROWS (bit mask): 1 = the gate layers of the WN stacks (384 columns), 2 = their 192 / 384-column LINEAR
// layers, 4 = every eligible layer; 0 = never.
int launch_conv_tcr(const qvc_conv_args& a, cudaStream_t stream) {
  const int mode = tc_env_int("QVC_TC_ROWS", 3);
  if (mode <= 0) return QVC_ERR_UNSUPPORTED;
  if (a.epilogue != QVC_EPI_LINEAR && a.epilogue != QVC_EPI_GATE) return QVC_ERR_UNSUPPORTED;
  const bool gate = a.epilogue == QVC_EPI_GATE;
  if (a.tap_split > 0) return QVC_ERR_UNSUPPORTED;
  const int esize = (int)opformat_bytes(a.opformat);
  const int kc = ROW_BYTES / esize;
  if (a.cin % kc) return QVC_ERR_UNSUPPORTED;
  const int halo = (a.k - 1) * a.dil;
  if (TM + halo > 256) return QVC_ERR_UNSUPPORTED;
  if (a.out_rows <= TM) return QVC_ERR_UNSUPPORTED;
  // piece width: the widest of 256 .. 64 columns that tiles the output (gate: half of it per gate half)
  int np = 0;
  if (gate) {
    const int H = a.cout / 2;
    for (int hp = 128; hp >= 32; hp -= 32)
      if (H % hp == 0) { np = 2 * hp; break; }
  } else {
    for (int c = 256; c >= 64; c -= 32)
      if (a.cout % c == 0) { np = c; break; }
  }
  if (np == 0) return QVC_ERR_UNSUPPORTED;
  if (!(mode & 4)) {
    const bool wn_family = np == 192 && (a.cout == 192 || a.cout == 384);
    if (!wn_family || !(mode & (gate ? 1 : 2))) return QVC_ERR_UNSUPPORTED;
  }
  const int npieces = a.cout / np;
  if (npieces > MAXPIECES) return QVC_ERR_UNSUPPORTED;
  const int ntb = (a.out_rows + 2 * TM - 1) / (2 * TM);
  const int ntiles = a.batch * ntb * npieces;
  if (ntiles < tc_sm_count() / 4 && !tc_env_int("QVC_TC_2CTA_FORCE", 0)) return QVC_ERR_UNSUPPORTED;
  // the row epilogue moves 32-byte sectors
  if (!aligned32(a.noise, 4)) return QVC_ERR_UNSUPPORTED;
  for (int s = 0; s < (gate ? 1 : a.nseg); ++s) {
    const qvc_epi_segment& g = a.seg[s];
    if (g.res_op.ptr) return QVC_ERR_UNSUPPORTED;
    if (!aligned32(g.res, 4) || !aligned32(g.accin, 4) || !aligned32(g.raw, 4) || !aligned32(g.op, esize)) return QVC_ERR_UNSUPPORTED;
    if (!gate && (g.col0 % 32 || g.ncols % 32)) return QVC_ERR_UNSUPPORTED;
  }
  if (a.bias && (((uintptr_t)a.bias & 15) || a.bias_bstride % 4)) return QVC_ERR_UNSUPPORTED;
  EncodeTiledFn encode = tc_get_encode();
  if (!encode) return QVC_ERR_UNSUPPORTED;
  QVC_REQUIRE((a.x.ld * esize) % 16 == 0 && ((uintptr_t)a.x.ptr & 15) == 0 && ((uintptr_t)a.w & 15) == 0,
              "conv1d(tcgen05): x / w must be 16-byte aligned with 16-byte row pitch");
  QVC_REQUIRE(a.batch == 1 || (a.x.bstride * esize) % 16 == 0, "conv1d(tcgen05): utterance pitch not 16-byte aligned");

  TcrParams p{};
  QVC_PROPAGATE(build_epi_params(a, &p.ep));
  p.cin = a.cin; p.k = a.k; p.dil = a.dil; p.pad_left = a.pad_left;
  p.np = np; p.npieces = npieces; p.ntb = ntb; p.ntiles = ntiles; p.batch = a.batch;
  for (int i = 0; i < npieces; ++i) {
    if (gate) {
      p.n0[i] = i * (np / 2);
      p.wrow[i][0] = p.n0[i];
      p.wrow[i][1] = a.cout / 2 + p.n0[i];
    } else {
      p.n0[i] = i * np;
      p.wrow[i][0] = p.n0[i];
      p.wrow[i][1] = p.n0[i] + np / 2;
    }
  }
  p.slab_box_rows = (TM + halo + 7) & ~7;
  p.slab_stage_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
  p.w_stage_bytes = (uint32_t)(np / 2) * ROW_BYTES;
  if (p.w_stage_bytes % 1024) return QVC_ERR_UNSUPPORTED;          // swizzle atoms: stages must stay 1024-byte aligned
  static const int stage_options[][2] = {{4, 8}, {3, 8}, {3, 6}, {2, 6}, {2, 4}, {2, 3}, {2, 2}};
  size_t smem = 0;
  bool fits = false;
  for (const auto& opt : stage_options) {
    smem = (size_t)opt[0] * p.slab_stage_bytes + (size_t)opt[1] * p.w_stage_bytes + 1024 + 256 + live_cache_bytes(a);
    if (smem <= (size_t)MAX_SMEM) { p.slab_stages = opt[0]; p.w_stages = opt[1]; fits = true; break; }
  }
  if (!fits) return QVC_ERR_UNSUPPORTED;

  const CUtensorMapDataType dt = a.opformat == QVC_OPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : (a.opformat == QVC_OPF_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.cin, (cuuint64_t)a.x_rows, (cuuint64_t)a.batch};
    cuuint64_t strides[2] = {(cuuint64_t)a.x.ld * esize,
                             (cuuint64_t)(a.batch > 1 ? a.x.bstride : (int64_t)a.x_rows * a.x.ld) * esize};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)p.slab_box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.mx, dt, 3, a.x.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05, rows): cuTensorMapEncodeTiled(x) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.k * a.cin, (cuuint64_t)a.cout};
    cuuint64_t strides[1] = {(cuuint64_t)a.k * a.cin * esize};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(np / 2)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&p.mw, dt, 2, const_cast<void*>(a.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05, rows): cuTensorMapEncodeTiled(w) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  int pairs = tc_sm_count() / 2;
  if (ntiles < pairs) pairs = ntiles;
#define QVC_TCR_DISPATCH(OPF)                                                            \
  return gate ? launch_r<OPF, QVC_EPI_GATE>(p, 2 * pairs, smem, stream)                  \
              : launch_r<OPF, QVC_EPI_LINEAR>(p, 2 * pairs, smem, stream);
  if (a.opformat == QVC_OPF_BF16) { QVC_TCR_DISPATCH(QVC_OPF_BF16) }
  if (a.opformat == QVC_OPF_F16) { QVC_TCR_DISPATCH(QVC_OPF_F16) }
  QVC_TCR_DISPATCH(QVC_OPF_TF32)
#undef QVC_TCR_DISPATCH
}

}  // namespace qvc
