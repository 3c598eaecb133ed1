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
  CUtensorMap mw;                          // w as (channel, output channel, tap); box = one channel chunk x np / 2 rows x tg taps
  int32_t cin, k, dil, pad_left;
  int32_t tg;                              // taps per filter stage (one TMA operation brings tg taps of one channel chunk)
  int32_t np;                              // accumulator columns per piece (the N of the MMA)
  int32_t npieces;
  int32_t wrow[MAXPIECES][2];              // first filter row CTA r loads for piece p
  int32_t n0[MAXPIECES];                   // LINEAR: output column of accumulator column 0; GATE: gate channel of column 0
  int32_t ntb, ntiles;
  int32_t slab_box_rows, slab_stages, w_stages;
  uint32_t slab_stage_bytes, w_stage_bytes;
  int32_t batch;
  int32_t live_cache_words;                // shared-memory words of the ragged-batch length cache
  int32_t bias_words;                      // LINEAR, shared bias: cout words of it are staged in shared memory (0 = not)
  int32_t debug;                           // timing experiments only (QVC_TCR_DEBUG, results are garbage): 1 no epilogue body, 2 no MMAs
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

// ---- LINEAR epilogue of one row (same arithmetic as tc::lin_finish), 32-column blocks ----
// One thread's row pointers into the tensors of an epilogue segment, hoisted out of the block loop: indexing the
// segment array per block costs register-indexed constant loads on the critical path (10 % of the stall samples of a
// memory-bound layer, profiles/r02_summary.md).
template <typename OT>
struct RowPtrs {
  const float* first;       // the fp32 stream added first: the residual if there is one, else the accumulate-into tensor
  const float* second;      // the accumulate-into tensor when there is a residual too
  const OT* res_op;         // the residual as an operand-format tensor holding leaky_relu(r) (lean instance only)
  float* raw;
  OT* op;
  float alpha, beta, slope, res_inv_slope;
  int col0, ncols;
  bool has_res;
};
template <typename OT>
__device__ __forceinline__ void row_ptrs_from(const EpiSeg& sg, int b, int t, RowPtrs<OT>& r) {
  r.has_res = sg.res.present();
  const TRef& f = r.has_res ? sg.res : sg.accin;
  r.first = f.present() ? f.at<float>(b, t, 0) : nullptr;
  r.second = (r.has_res && sg.accin.present()) ? sg.accin.at<float>(b, t, 0) : nullptr;
  r.res_op = sg.res_op.present() ? sg.res_op.at<OT>(b, t, 0) : nullptr;
  r.res_inv_slope = sg.res_inv_slope;
  r.raw = sg.raw.present() ? sg.raw.at<float>(b, t, 0) : nullptr;
  r.op = sg.op.present() ? sg.op.at<OT>(b, t, 0) : nullptr;
  r.alpha = sg.alpha; r.beta = sg.beta; r.slope = sg.slope;
  r.col0 = sg.col0; r.ncols = sg.ncols;
}
template <typename OT>
__device__ __forceinline__ void row_ptrs(const EpiParams& ep, int si, int b, int t, RowPtrs<OT>& r) {
  if (si) row_ptrs_from<OT>(ep.seg[1], b, t, r);        // two branches: constant offsets into the parameter bank
  else    row_ptrs_from<OT>(ep.seg[0], b, t, r);
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

// Finishes the 32-column block starting at output column n (segment channel c) in two halves of 16 columns: the
// accumulator is read 16 columns at a time so that three blocks of prefetched residual fit the register file beside it.
// All lanes execute the TMEM loads; `act` = this lane's row is real and the block lies inside the segment.
template <int OPF>
__device__ __forceinline__ void row_lin_finish(const RowPtrs<typename OpType<OPF>::type>& rp, const float* bias, int n, int c,
                                               bool act, bool live, uint32_t taddr, const float* r) {
  const float alpha = rp.alpha, beta = rp.beta, slope = rp.slope;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
    tmem_ld16(taddr + 16 * h, v);
    float bv[16];
    if (bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(bias + n + 16 * h + 4 * i);
        bv[4 * i] = q.x; bv[4 * i + 1] = q.y; bv[4 * i + 2] = q.z; bv[4 * i + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) bv[i] = 0.f;
    }
    tmem_wait();
    if (!act) continue;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += bv[i];
    const float* rr = r + 16 * h;
    if (rp.second) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float a[8];
        ldg256(rp.second + c + 16 * h + 8 * g, reinterpret_cast<uint32_t*>(a));
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * g + i] = fmaf(beta, fmaf(alpha, v[8 * g + i], rr[8 * g + i]), a[i]);
      }
    } else if (rp.first && rp.has_res) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = beta * fmaf(alpha, v[i], rr[i]);
    } else if (rp.first) {
      const float ab = alpha * beta;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(ab, v[i], rr[i]);
    } else {
      const float ab = alpha * beta;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ab * v[i];
    }
    if (rp.raw) {
      float* wp = rp.raw + c + 16 * h;
      stg256(wp, reinterpret_cast<const uint32_t*>(v));
      stg256(wp + 8, reinterpret_cast<const uint32_t*>(v) + 8);
    }
    if (rp.op) {
      if (live) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * slope);      // leaky-relu, slope <= 1
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      auto* op = rp.op + c + 16 * h;
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

template <int OPF, int EPI, bool LEAN>
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
  // LINEAR layers with one bias vector for all utterances keep it in shared memory: the streaming epilogue leaves the L1
  // no room for it (5 % load hit rate), and a bias load that goes to L2 sits on the critical path of every block
  float* bias_s = reinterpret_cast<float*>(live_s + p.live_cache_words);
  const bool bias_in_smem = EPI == QVC_EPI_LINEAR && p.bias_words > 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * p.slab_stages + 2 * p.w_stages + 2; ++i) mbar_init(bar0 + 8 * i, 1);
    // an accumulator set is drained by all eight epilogue warps of both CTAs (each group of four takes half of the columns)
    mbar_init(tmem_empty, 2 * N_EPI_WARPS);
    mbar_init(tmem_empty + 8, 2 * N_EPI_WARPS);
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
  if (p.ep.live != nullptr) load_live_cache(p.ep, p.batch, live_s);
  if (bias_in_smem)
    for (int i = threadIdx.x; i < p.bias_words; i += blockDim.x) bias_s[i] = p.ep.bias[i];
  if (p.ep.live != nullptr || bias_in_smem) __syncthreads();
  const int n_cchunks = p.cin / KC;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mx) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mw) : "memory");
      const uint32_t lead_full_slab = map_to_cta(full_slab, 0), lead_full_w = map_to_cta(full_w, 0);
      const uint32_t slab_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
      const int k = p.k, pad_left = p.pad_left, tg = p.tg;
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
          // One operation per group of tg taps: the single producer thread spends ~400 cycles per TMA operation (wait,
          // expect_tx, issue), which bounded every layer whose stage feeds fewer than ~400 cycles of MMAs
          // (profiles/r02_summary.md: a k = 3, 128 -> 128 layer took 61 us with neither MMAs nor epilogue).
          for (int j = 0; j < k; j += tg) {
            mbar_wait(empty_w + 8 * ws, wph ^ 1u);
            if (leader) mbar_expect_tx(full_w + 8 * ws, 2 * p.w_stage_bytes);
            tma2_load_3d(w0 + ws * p.w_stage_bytes, &p.mw, lead_full_w + 8 * ws, cc * KC, wrow, j);
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
      const int k = p.k, dil = p.dil, tg = p.tg;
      const uint32_t w_tap_bytes = (uint32_t)(p.np >> 1) * ROW_BYTES;
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
          for (int j0 = 0; j0 < k; j0 += tg) {
            mbar_wait(full_w + 8 * ws, wph);
            tc_fence_after();
            const int jn = k - j0 < tg ? k - j0 : tg;
            const uint64_t adesc0 = desc_hi | (uint64_t)(((slab + (uint32_t)(j0 * dil) * ROW_BYTES) & 0x3FFFFu) >> 4);
            const uint64_t bdesc0 = desc_hi | (uint64_t)(((w0 + ws * p.w_stage_bytes) & 0x3FFFFu) >> 4);
            if (elect_one()) {
              if (!(p.debug & 2)) {
                for (int jj = 0; jj < jn; ++jj) {
                  const uint32_t first = (cc | j0 | jj) == 0 ? 0u : 1u;
                  const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(jj * dil) * (ROW_BYTES >> 4));
                  const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)jj * (w_tap_bytes >> 4));
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma2<OPF>(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
                }
              }
              tc2_commit(empty_w + 8 * ws);
              if (j0 + tg >= k) tc2_commit(empty_slab + 8 * s);
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
    // ===================== epilogue (warps 2..9 of both CTAs): group h takes half h of every tile's columns =====================
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
      const int t = tb * 2 * TM + (int)rank * TM + row;
      const bool ok = t < p.ep.out_rows;
      const bool live = t < live_rows(p.ep, b);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BUF_COLS;
      if constexpr (EPI == QVC_EPI_LINEAR) {
        // The fp32 stream of a block (residual / accumulate-into) is requested one block ahead of its use, the first one
        // before the accumulator is even complete.
        const int nblk = p.np >> 5;
        const int nbase = p.n0[piece];
        const float* bias = p.ep.bias == nullptr ? nullptr : (bias_in_smem ? bias_s : p.ep.bias + (int64_t)b * p.ep.bias_bs);
        RowPtrs<OT> rp;
        int seg_idx = -1;
        // segment of block blk (re-read only when it changes), channel of its first column, inside the segment?
        auto enter = [&](int blk, int& c, bool& act) {
          const int n = nbase + 32 * blk;
          const int si = (p.ep.nseg > 1 && n >= p.ep.seg[1].col0) ? 1 : 0;
          if (si != seg_idx) { row_ptrs<OT>(p.ep, si, b, t, rp); seg_idx = si; }
          c = n - rp.col0;
          act = ok && c >= 0 && c + 32 <= rp.ncols;
        };
        auto request = [&](int blk, float* r) {
          int c; bool act;
          enter(blk, c, act);
          if (act && rp.first) load_row32(rp.first + c, r);
        };
        const int blk0 = h ? (nblk + 1) >> 1 : 0, blk1 = h ? nblk : (nblk + 1) >> 1;      // this group's blocks
        // ---- lean instance (chosen by the host: every block of a group inside ONE segment, at most one fp32 input stream,
        // bias shared).  Straight-line code per 16 columns; the generic instance below spends ~15 instructions per element
        // on predicates, segment look-ups and address arithmetic and is issue-bound on memory-bound layers
        // (profiles/r02_summary.md); this one ~5.
        if constexpr (LEAN) {
          if (blk0 < blk1 && !(p.debug & 1)) {
            const int n_first = nbase + 32 * blk0;
            const int si = (p.ep.nseg > 1 && n_first >= p.ep.seg[1].col0) ? 1 : 0;
            row_ptrs<OT>(p.ep, si, b, t, rp);
            const float* first = rp.first ? rp.first + (n_first - rp.col0) : nullptr;
            const OT* rop = rp.res_op ? rp.res_op + (n_first - rp.col0) : nullptr;
            const float inv_slope = rp.res_inv_slope;
            float* raw = rp.raw ? rp.raw + (n_first - rp.col0) : nullptr;
            OT* op = rp.op ? rp.op + (n_first - rp.col0) : nullptr;
            const float* bs = p.ep.bias ? bias_s + n_first : nullptr;
            const float alpha = rp.alpha, beta = rp.beta, ab = rp.alpha * rp.beta, slope = rp.slope;
            const bool is_res = rp.has_res, ragged = p.ep.live != nullptr;
            const int nb = blk1 - blk0;
            float ra[32], rb[32];
            // operand-format residual: 32 values = 64 or 128 bytes, kept as raw bits until used
            auto load_rop = [&](const OT* src, float* dst) {
              ldg256(src, reinterpret_cast<uint32_t*>(dst));
              ldg256(reinterpret_cast<const char*>(src) + 32, reinterpret_cast<uint32_t*>(dst) + 8);
              if constexpr (!opf_is16(OPF)) {
                ldg256(reinterpret_cast<const char*>(src) + 64, reinterpret_cast<uint32_t*>(dst) + 16);
                ldg256(reinterpret_cast<const char*>(src) + 96, reinterpret_cast<uint32_t*>(dst) + 24);
              }
            };
            if (ok) {
              if (first) load_row32(first, ra);
              else if (rop) load_rop(rop, ra);
            }
            mbar_wait(tmem_full + 8 * buf, bph);
            tc_fence_after();
            auto lean_block = [&](int i, const float* rc, float* rn) {
              if (ok && i + 1 < nb) {
                if (first) load_row32(first + 32 * (i + 1), rn);
                else if (rop) load_rop(rop + 32 * (i + 1), rn);
              }
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {               // two halves of 16 columns: keeps the live registers low
                float v[16];
                tmem_ld16(taddr + 32 * (blk0 + i) + 16 * hh, v);
                float4 bq[4];
                if (bs) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) bq[j] = *reinterpret_cast<const float4*>(bs + 32 * i + 16 * hh + 4 * j);
                }
                tmem_wait();
                if (!ok) continue;
                if (bs) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    v[4 * j] += bq[j].x; v[4 * j + 1] += bq[j].y; v[4 * j + 2] += bq[j].z; v[4 * j + 3] += bq[j].w;
                  }
                }
                const float* rr = rc + 16 * hh;
                if (first) {
                  if (is_res) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = beta * fmaf(alpha, v[j], rr[j]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaf(ab, v[j], rr[j]);
                  }
                } else if (rop) {                             // decode the operand copy: undo its leaky-relu
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    float x;
                    if constexpr (opf_is16(OPF)) {
                      const uint32_t w = __float_as_uint(rc[8 * hh + (j >> 1)]);
                      x = op16_to_float<OPF>((j & 1) ? (w >> 16) : (w & 0xffffu));
                    } else {
                      x = rr[j];
                    }
                    x = x > 0.f ? x : x * inv_slope;
                    v[j] = beta * fmaf(alpha, v[j], x);
                  }
                } else if (ab != 1.f) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] *= ab;
                }
                if (raw) {
                  stg256(raw + 32 * i + 16 * hh, reinterpret_cast<const uint32_t*>(v));
                  stg256(raw + 32 * i + 16 * hh + 8, reinterpret_cast<const uint32_t*>(v) + 8);
                }
                if (op) {
                  if (slope != 1.f) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * slope);
                  }
                  if (ragged && !live) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                  }
                  OT* o = op + 32 * i + 16 * hh;
                  if constexpr (opf_is16(OPF)) {
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = op16_pack2<OPF>(v[2 * j], v[2 * j + 1]);
                    stg256(o, w);
                  } else {
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(to_operand<OPF>(v[j]));
                    stg256(o, w);
                    stg256(o + 8, w + 8);
                  }
                }
              }
            };
            for (int i = 0; i < nb; i += 2) {
              lean_block(i, ra, rb);
              if (i + 1 < nb) lean_block(i + 1, rb, ra);
            }
          } else {
            mbar_wait(tmem_full + 8 * buf, bph);
            tc_fence_after();
          }
        } else {
        float r0[32], r1[32];
        if (blk0 < blk1) request(blk0, r0);
        mbar_wait(tmem_full + 8 * buf, bph);
        tc_fence_after();
        auto step = [&](int blk, float* rc, float* rn) {     // request block blk + 1 into rn, finish block blk (stream rc)
          if (blk + 1 < blk1) request(blk + 1, rn);
          int c; bool act;
          enter(blk, c, act);
          row_lin_finish<OPF>(rp, bias, nbase + 32 * blk, c, act, live, taddr + 32 * blk, rc);
        };
        for (int blk = blk0; blk < blk1; blk += 2) {
          step(blk, r0, r1);
          if (blk + 1 < blk1) step(blk + 1, r1, r0);
        }
        }
      } else {
        const int hp = p.np >> 1;                   // gate channels of this piece: columns [0, hp) tanh, [hp, np) sigmoid
        const int ch0 = p.n0[piece];
        const float* gb = p.ep.bias + (int64_t)b * p.ep.bias_bs;
        const EpiSeg& sgm = p.ep.seg[0];
        mbar_wait(tmem_full + 8 * buf, bph);
        tc_fence_after();
        const int nun = hp >> 4;                    // units of 16 gate channels; this group takes half of them
        const int u0 = h ? (nun + 1) >> 1 : 0, u1 = (p.debug & 1) ? 0 : (h ? nun : (nun + 1) >> 1);
        for (int u = u0; u < u1; ++u) {
          float lo[16], hi[16];
          tmem_ld16(taddr + 16 * u, lo);
          tmem_ld16(taddr + hp + 16 * u, hi);
          const int n = ch0 + 16 * u;
          float bl[16], bh[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(gb + n) + i);
            const float4 c = __ldg(reinterpret_cast<const float4*>(gb + p.ep.half + n) + i);
            bl[4 * i] = a.x; bl[4 * i + 1] = a.y; bl[4 * i + 2] = a.z; bl[4 * i + 3] = a.w;
            bh[4 * i] = c.x; bh[4 * i + 1] = c.y; bh[4 * i + 2] = c.z; bh[4 * i + 3] = c.w;
          }
          tmem_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) lo[i] = fast_gate(lo[i] + bl[i], hi[i] + bh[i]);
          if (ok) {
            if (sgm.raw.present()) {
              float* wp = sgm.raw.at<float>(b, t, n);
              stg256(wp, reinterpret_cast<const uint32_t*>(lo));
              stg256(wp + 8, reinterpret_cast<const uint32_t*>(lo) + 8);
            }
            if (sgm.op.present()) {
              if (!live) {
#pragma unroll
                for (int i = 0; i < 16; ++i) lo[i] = 0.f;
              }
              OT* op = sgm.op.at<OT>(b, t, n);
              if constexpr (opf_is16(OPF)) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = op16_pack2<OPF>(lo[2 * i], lo[2 * i + 1]);
                stg256(op, w);
              } else {
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(to_operand<OPF>(lo[i]));
                stg256(op, w);
                stg256(op + 8, w + 8);
              }
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
  // Release / acquire, not the relaxed form: the peer's epilogue warps arrive REMOTELY (and relaxed) on this CTA's
  // tmem_empty barriers, and this CTA's shared memory is handed to the next grid's CTA the moment it exits -- the release
  // of the cluster barrier is what guarantees that every such remote arrive has landed before that.  (A relaxed execution
  // barrier was tried: it saved the MEMBAR at the end, ~1 us per launch, and left a remote arrive free to land in the
  // NEXT kernel's freshly initialised barrier.)
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BUF_COLS) : "memory");
  }
}

template <int OPF, int EPI, bool LEAN>
int launch_r(const TcrParams& p, int grid, size_t smem, cudaStream_t stream) {
  static std::atomic<bool> attr_done[MAX_DEVICES];
  if (first_use_on_device(attr_done))
    QVC_CHECK_CUDA(cudaFuncSetAttribute(conv_tcr_kernel<OPF, EPI, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
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
  QVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tcr_kernel<OPF, EPI, LEAN>, p));
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e1, stream));
  return post_launch("conv_tcr_kernel");
}

bool aligned32(const qvc_tensor& t, size_t esize) {
  return t.ptr == nullptr || (((uintptr_t)t.ptr & 31) == 0 && ((size_t)t.ld * esize) % 32 == 0 && ((size_t)t.bstride * esize) % 32 == 0);
}

}  // namespace

// Frames-on-rows pair kernel.  Returns QVC_ERR_UNSUPPORTED (error string untouched) when the layer is not one of its
// cases.  QVC_TC_ROWS (bit mask): 1 = the gate layers of the WN stacks (384 columns), 2 = their 192 / 384-column LINEAR
// layers, 8 = the 128 -> 128 LINEAR layers (MRF-2), 4 = every eligible layer; 0 = never.
int launch_conv_tcr(const qvc_conv_args& a, cudaStream_t stream) {
  const int mode = tc_env_int("QVC_TC_ROWS", 11);
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
    const bool mrf2_family = !gate && a.cin == 128 && a.cout == 128;
    if (!((wn_family && (mode & (gate ? 1 : 2))) || (mrf2_family && (mode & 8)))) return QVC_ERR_UNSUPPORTED;
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
    if (!aligned32(g.res_op, esize)) return QVC_ERR_UNSUPPORTED;
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
  p.debug = tc_env_int("QVC_TCR_DEBUG", 0);
  p.live_cache_words = (int32_t)((live_cache_bytes(a) + 15) / 16 * 4);
  p.bias_words = (!gate && a.bias && a.bias_bstride == 0 && a.cout <= 2048) ? a.cout : 0;
  p.slab_box_rows = (TM + halo + 7) & ~7;
  p.slab_stage_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
  const uint32_t w_tap_bytes = (uint32_t)(np / 2) * ROW_BYTES;
  if (w_tap_bytes % 1024) return QVC_ERR_UNSUPPORTED;              // swizzle atoms: every tap's tile must stay 1024-byte aligned
  // taps per filter stage: as many as keep a stage at or under 32 KB, spread evenly over the groups
  {
    int tg = (int)(32768u / w_tap_bytes);
    tg = tg < 1 ? 1 : (tg > a.k ? a.k : tg);
    const int tg_env = tc_env_int("QVC_TCR_TG", 0);
    if (tg_env >= 1 && tg_env <= a.k) tg = tg_env;
    const int groups = (a.k + tg - 1) / tg;
    p.tg = (a.k + groups - 1) / groups;
  }
  p.w_stage_bytes = (uint32_t)p.tg * w_tap_bytes;
  static const int stage_options[][2] = {{4, 4}, {4, 3}, {3, 3}, {3, 2}, {2, 2}};
  size_t smem = 0;
  bool fits = false;
  {
    const int ss = tc_env_int("QVC_TCR_SS", 0), ws = tc_env_int("QVC_TCR_WS", 0);       // tuning override
    if (ss >= 1 && ss <= 16 && ws >= 2 && ws <= 16) {
      smem = (size_t)ss * p.slab_stage_bytes + (size_t)ws * p.w_stage_bytes + 1024 + 256 + 4 * (size_t)p.live_cache_words +
             4 * (size_t)p.bias_words;
      if (smem <= (size_t)MAX_SMEM) { p.slab_stages = ss; p.w_stages = ws; fits = true; }
    }
  }
  for (const auto& opt : stage_options) {
    if (fits) break;
    smem = (size_t)opt[0] * p.slab_stage_bytes + (size_t)opt[1] * p.w_stage_bytes + 1024 + 256 + 4 * (size_t)p.live_cache_words +
           4 * (size_t)p.bias_words;
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
    cuuint64_t dims[3] = {(cuuint64_t)a.cin, (cuuint64_t)a.cout, (cuuint64_t)a.k};
    cuuint64_t strides[2] = {(cuuint64_t)a.k * a.cin * esize, (cuuint64_t)a.cin * esize};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)(np / 2), (cuuint32_t)p.tg};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.mw, dt, 3, const_cast<void*>(a.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05, rows): cuTensorMapEncodeTiled(w) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  int pairs = tc_sm_count() / 2;
  if (ntiles < pairs) pairs = ntiles;
  // lean LINEAR instance: every group's column range ([0, ceil(nblk / 2)) and the rest of each piece) lies inside one
  // segment, no segment has both a residual and an accumulate-into tensor, and the bias (if any) sits in shared memory
  bool lean = !gate && (a.bias == nullptr || p.bias_words > 0);
  if (lean) {
    const int nblk = np / 32, split = (nblk + 1) / 2;
    for (int i = 0; i < npieces && lean; ++i)
      for (int hgrp = 0; hgrp < 2 && lean; ++hgrp) {
        const int c0 = p.n0[i] + 32 * (hgrp ? split : 0), c1 = p.n0[i] + 32 * (hgrp ? nblk : split);
        if (c0 >= c1) continue;
        const int si = (a.nseg > 1 && c0 >= a.seg[1].col0) ? 1 : 0;
        const qvc_epi_segment& g = a.seg[si];
        if (c0 < g.col0 || c1 > g.col0 + g.ncols || ((g.res.ptr || g.res_op.ptr) && g.accin.ptr)) lean = false;
        if (si == 0 && a.nseg > 1 && c1 > a.seg[1].col0) lean = false;
      }
  }
  if (tc_env_int("QVC_TCR_LEAN", 1) == 0) lean = false;
  if (!lean)                                        // the generic instance has no operand-format residual
    for (int s = 0; s < (gate ? 0 : a.nseg); ++s)
      if (a.seg[s].res_op.ptr) return QVC_ERR_UNSUPPORTED;
#define QVC_TCR_DISPATCH(OPF)                                                                  \
  return gate ? launch_r<OPF, QVC_EPI_GATE, false>(p, 2 * pairs, smem, stream)                 \
              : (lean ? launch_r<OPF, QVC_EPI_LINEAR, true>(p, 2 * pairs, smem, stream)        \
                      : launch_r<OPF, QVC_EPI_LINEAR, false>(p, 2 * pairs, smem, stream));
  if (a.opformat == QVC_OPF_BF16) { QVC_TCR_DISPATCH(QVC_OPF_BF16) }
  if (a.opformat == QVC_OPF_F16) { QVC_TCR_DISPATCH(QVC_OPF_F16) }
  QVC_TCR_DISPATCH(QVC_OPF_TF32)
#undef QVC_TCR_DISPATCH
}

}  // namespace qvc
