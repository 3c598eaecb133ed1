// conv_tc2.cu -- the series convolution of conv_tc.cu on CTA PAIRS (tcgen05 cta_group::2), for layers with
// an even number of 128-channel output chunks (MRF-1 256 -> 256, the polyphase upsamplers, conv_pre).
//
// Why: measured on B200 (profiles/r01_v2_summary.md) a cta_group::1 MMA with M = 128, N = 256 and both
// operands in shared memory runs at ~2/3 of its ideal rate -- 12 KB of operand reads per MMA at ~64 B/cycle.
// A CTA pair issues ONE M = 256, N = 256 MMA: CTA r supplies the 128 filter rows of chunk r (A) and 128 of
// the 256 frames (its half of B), i.e. 8 KB per CTA for the same 128 cycles, and receives the accumulator of
// its own 128 channels x all 256 frames in its own TMEM.
//
//   pair tile  = (utterance, 256-frame block, group of two 128-channel chunks)
//   CTA r      : TMA-loads frames [t0 + 128 r - pad, + 128 + halo) of every channel chunk (its slab) and the
//                filter rows of chunk r; epilogue over its 128 channels x 256 frames (same code as conv_tc)
//   leader (r=0): issues the MMAs; its "full" barriers collect the TMA bytes of BOTH CTAs (the peer's loads
//                signal the leader's barrier), commits are multicast to both CTAs' "empty" / "accumulator
//                ready" barriers; both CTAs' epilogue warps arrive on the leader's "accumulator drained" barrier.
#include <cstdlib>

#include "conv_tc_common.cuh"

namespace qvc {

namespace {

using namespace tc;

constexpr int MAXGROUPS2 = 8;
constexpr int MAXSRC = QVC_MAX_SUM_SOURCES;

// one (input series, filter) pair; an ordinary convolution has one, qvc_conv1d_sum up to MAXSRC
struct alignas(64) Src2 {
  CUtensorMap mx;                          // x as (channel, frame, utterance)
  CUtensorMap mw;                          // w as (tap*cin + channel, output channel)
  int32_t cin, k, dil, pad_left;
  // structured zeros of the filter (qvc_conv_args.tap_split): channel chunks from split_chunk on use taps
  // [jlo[1], jhi[1]], the ones before it [jlo[0], jhi[0]]; without a hint both ranges are [0, k-1]
  int32_t split_chunk;
  int32_t jlo[2], jhi[2];
  int32_t slab_box_rows;                   // one TMA box per slab (half tile + halo <= 256 rows)
};

struct alignas(64) Tc2Params {
  Src2 src[MAXSRC];
  int32_t nsrc;
  int32_t pair_n;                          // frames per pair tile: 256 (LINEAR), 128 (GATE: two accumulators)
  int32_t nacc;                            // accumulators per CTA: 1, or 2 = (lo, hi) halves of a gate pair
  int32_t nbuf;                            // accumulator sets in TMEM: 2 (epilogue overlaps the next tile), or 1 when nacc * pair_n = 512
  int32_t ntb;                             // frame blocks per utterance
  int32_t ngroups, ntiles;                 // groups of two chunks; pair tiles
  int32_t row0[MAXGROUPS2][2];             // filter row on lane 0 of the chunk of CTA r (lo half for GATE)
  int32_t valid[MAXGROUPS2][2];
  int32_t slab_stages, w_stages;
  uint32_t slab_stage_bytes;
  int32_t batch;
  EpiParams ep;
};

// SUM: the K loop runs over p.nsrc sources into one accumulator and the epilogue is the multi-residual one of
// qvc_conv1d_sum; every other instance has exactly one source (the loops over sources collapse at compile time).
template <int OPF, int EPI, bool SUM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) conv_tc2_kernel(const __grid_constant__ Tc2Params p) {
  constexpr int ESIZE = opf_is16(OPF) ? 2 : 4;
  constexpr int KC = ROW_BYTES / ESIZE;
  constexpr uint32_t FMT = mma_format(OPF);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slab0 = smem_base;
  const uint32_t w0 = slab0 + p.slab_stages * p.slab_stage_bytes;
  const uint32_t w_stage_bytes = (uint32_t)p.nacc * CHUNK_BYTES;
  const uint32_t bar0 = w0 + p.w_stages * w_stage_bytes;
  const int PN = p.pair_n, HN = p.pair_n >> 1;
  // barrier layout (same offsets in both CTAs): full_slab[SS] empty_slab[SS] full_w[WS] empty_w[WS]
  // tmem_full[2] tmem_empty[2], then the TMEM base word.  full_* and tmem_empty are only used in the leader.
  const uint32_t full_slab = bar0, empty_slab = full_slab + 8 * p.slab_stages;
  const uint32_t full_w = empty_slab + 8 * p.slab_stages, empty_w = full_w + 8 * p.w_stages;
  const uint32_t tmem_full = empty_w + 8 * p.w_stages, tmem_empty = tmem_full + 16;
  const uint32_t tmem_slot = tmem_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  int32_t* live_s = reinterpret_cast<int32_t*>(tmem_slot_ptr + 4);       // ragged batches: copy of ep.live

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int nsrc = SUM ? p.nsrc : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * p.slab_stages + 2 * p.w_stages + 2; ++i) mbar_init(bar0 + 8 * i, 1);
    mbar_init(tmem_empty, 2 * N_EPI_WARPS);
    mbar_init(tmem_empty + 8, 2 * N_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * ACC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.ep.live != nullptr) {
    load_live_cache(p.ep, p.batch, live_s);
    __syncthreads();
  }

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      for (int si = 0; si < nsrc; ++si) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.src[si].mx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.src[si].mw) : "memory");
      }
      const uint32_t lead_full_slab = map_to_cta(full_slab, 0), lead_full_w = map_to_cta(full_w, 0);
      uint32_t s = 0, ph = 0, ws = 0, wph = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs) {
        const int gi = tile % p.ngroups;
        const int rest = tile / p.ngroups;
        const int tb = rest % p.ntb, b = rest / p.ntb;
        if (tile_dead(p.ep, live_s, b, tb * PN)) continue;
        const int t0 = tb * PN + (int)rank * HN;
        const int wrow = p.row0[gi][rank];
        for (int si = 0; si < nsrc; ++si) {
          const Src2& S = p.src[si];
          // this source's scalars in registers: a constant-bank load per channel chunk on this single thread is on the
          // critical path
          const int n_cchunks = S.cin / KC, cin = S.cin, pad_left = S.pad_left;
          const uint32_t slab_bytes = (uint32_t)S.slab_box_rows * ROW_BYTES;
          const int jlo0 = S.jlo[0], jhi0 = S.jhi[0], jlo1 = S.jlo[1], jhi1 = S.jhi[1], split = S.split_chunk;
          for (int cc = 0; cc < n_cchunks; ++cc) {
            mbar_wait(empty_slab + 8 * s, ph ^ 1u);
            if (leader) mbar_expect_tx(full_slab + 8 * s, 2 * slab_bytes);       // bytes of both CTAs
            tma2_load_3d(slab0 + s * p.slab_stage_bytes, &S.mx, lead_full_slab + 8 * s, cc * KC, t0 - pad_left, b);
            if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
            const int jbeg = cc >= split ? jlo1 : jlo0, jend = cc >= split ? jhi1 : jhi0;
            for (int j = jbeg; j <= jend; ++j) {
              mbar_wait(empty_w + 8 * ws, wph ^ 1u);
              if (leader) mbar_expect_tx(full_w + 8 * ws, 2 * w_stage_bytes);
              tma2_load_2d(w0 + ws * w_stage_bytes, &S.mw, lead_full_w + 8 * ws, j * cin + cc * KC, wrow);
              if (EPI != QVC_EPI_LINEAR)       // hi half of the gate pair: the same lanes, H rows further down
                tma2_load_2d(w0 + ws * w_stage_bytes + CHUNK_BYTES, &S.mw, lead_full_w + 8 * ws, j * cin + cc * KC,
                             p.ep.half + wrow);
              if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(PN >> 3) << 17) |
                             ((uint32_t)((2 * CHUNK_M) >> 4) << 24);
      const uint64_t desc_hi = smem_desc(0);
      uint32_t s = 0, ph = 0, ws = 0, wph = 0, ait = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs) {
        {
          const int rest = tile / p.ngroups;
          if (tile_dead(p.ep, live_s, rest / p.ntb, (rest % p.ntb) * PN)) continue;
        }
        const uint32_t buf = p.nbuf == 2 ? (ait & 1u) : 0u;
        const uint32_t bph = p.nbuf == 2 ? ((ait >> 1) & 1u) : (ait & 1u);
        mbar_wait(tmem_empty + 8 * buf, bph ^ 1u);                 // both CTAs' epilogues drained this set
        tc_fence_after();
        const uint32_t d = tmem_base + buf * ACC_COLS;
        for (int si = 0; si < nsrc; ++si) {
          const Src2& S = p.src[si];
          const int n_cchunks = S.cin / KC, dil = S.dil;
          const int jlo0 = S.jlo[0], jhi0 = S.jhi[0], jlo1 = S.jlo[1], jhi1 = S.jhi[1], split = S.split_chunk;
          for (int cc = 0; cc < n_cchunks; ++cc) {
            mbar_wait(full_slab + 8 * s, ph);
            const uint32_t slab = slab0 + s * p.slab_stage_bytes;
            const int jbeg = cc >= split ? jlo1 : jlo0, jlast = cc >= split ? jhi1 : jhi0;
            for (int j = jbeg; j <= jlast; ++j) {
              mbar_wait(full_w + 8 * ws, wph);
              tc_fence_after();
              const uint32_t first = (si == 0 && cc == 0 && j == jlo0) ? 0u : 1u;
              const uint64_t bdesc = desc_hi | (uint64_t)(((slab + (uint32_t)(j * dil) * ROW_BYTES) & 0x3FFFFu) >> 4);
              const uint64_t adesc = desc_hi | (uint64_t)(((w0 + ws * w_stage_bytes) & 0x3FFFFu) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma2<OPF>(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
                if (EPI != QVC_EPI_LINEAR) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma2<OPF>(d + (uint32_t)PN, adesc + (CHUNK_BYTES >> 4) + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
                }
                tc2_commit(empty_w + 8 * ws);
                if (j == jlast) tc2_commit(empty_slab + 8 * s);
              }
              __syncwarp();
              if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
            }
            if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
          }
        }
        if (elect_one()) tc2_commit(tmem_full + 8 * buf);
        __syncwarp();
        ++ait;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs): own 128 channels x 256 frames =====================
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int col_begin = h * HN, col_end = col_begin + HN;
    const int lic = q * 32 + lane;
    const uint32_t lead_tmem_empty = map_to_cta(tmem_empty, 0);
    uint32_t ait = 0;
    for (int tile = pair; tile < p.ntiles; tile += npairs) {
      const int gi = tile % p.ngroups;
      const int rest = tile / p.ngroups;
      const int tb = rest % p.ntb, b = rest / p.ntb;
      const int t0 = tb * PN;
      if (tile_dead(p.ep, live_s, b, t0)) continue;
      const uint32_t buf = p.nbuf == 2 ? (ait & 1u) : 0u;
      const uint32_t bph = p.nbuf == 2 ? ((ait >> 1) & 1u) : (ait & 1u);
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS;
      const int nvalid = p.valid[gi][rank];
      const bool warp_live = q * 32 < nvalid;
      if constexpr (EPI == QVC_EPI_LINEAR) {
        LinCtx k;
        {
          const int n_w = p.row0[gi][rank] + q * 32;
          const int n = n_w + lane;
          k.sg = &p.ep.seg[(p.ep.nseg > 1 && n_w >= p.ep.seg[1].col0) ? 1 : 0];
          const int c = n - k.sg->col0;
          k.ok = warp_live && lic < nvalid && c >= 0 && c < k.sg->ncols;
          k.all_ok = __all_sync(0xffffffffu, k.ok);
          k.c = k.ok ? c : 0;
          k.b = b;
          k.lim = live_rows(p.ep, b);
          if constexpr (SUM) k.bias = sum_bias(p.ep, b, n, k.ok);
          else               k.bias = (p.ep.bias && k.ok) ? p.ep.bias[(int64_t)b * p.ep.bias_bs + n] : 0.f;
        }
        if constexpr (SUM) {
          auto frames_at = [&](int col) -> int {
            const int left = min(col_end - col, p.ep.out_rows - (t0 + col));
            return left < 32 ? left : 32;
          };
          float r[96];
          bool primed = false;
          if (warp_live) {
            const int nv = frames_at(col_begin);
            if (nv > 0) sum_load<OPF>(p.ep, k, t0 + col_begin, nv, r);
            primed = true;
          }
          mbar_wait(tmem_full + 8 * buf, bph);
          tc_fence_after();
          if (warp_live) {
            for (int col = col_begin; col < col_end; col += 32) {
              const int nv = frames_at(col);
              if (nv <= 0) break;
              if (!(primed && col == col_begin)) sum_load<OPF>(p.ep, k, t0 + col, nv, r);
              sum_finish<OPF>(p.ep, k, t0 + col, nv, r, tbase + (uint32_t)col);
            }
          }
        } else {
          auto frames_at = [&](int col) -> int {
            const int left = min(col_end - col, p.ep.out_rows - (t0 + col));
            return left < 64 ? left : 64;
          };
          // Two experiments on the residual loads of the 16-bit modes, both dropped (profiles/r02_summary.md): (1) both
          // 64-frame superblocks of a warp in ONE 128-deep burst, two values per register: c2 layers 25-35 % SLOWER (with
          // 227 KB of the SM carved out as shared memory almost no L1 is left to hold that many loads in flight);
          // (2) the residual staged through shared memory by cp.async one superblock ahead (no exposed load latency at
          // all): no change.
          float r[64];
          bool primed = false;
          if (warp_live) {
            const int nv = frames_at(col_begin);
            if (nv > 0) lin_load<OPF>(k, t0 + col_begin, nv, r);
            primed = true;
          }
          mbar_wait(tmem_full + 8 * buf, bph);
          tc_fence_after();
          if (warp_live) {
            for (int col = col_begin; col < col_end; col += 64) {
              const int nv = frames_at(col);
              if (nv <= 0) break;
              if (!(primed && col == col_begin)) lin_load<OPF>(k, t0 + col, nv, r);
              lin_finish<OPF>(k, t0 + col, nv, r, tbase + (uint32_t)col);
            }
          }
        }
      } else {
        mbar_wait(tmem_full + 8 * buf, bph);
        tc_fence_after();
        if (warp_live) {
          const int n = p.row0[gi][rank] + lic;
          const bool ok = lic < nvalid;
          const bool all_ok = __all_sync(0xffffffffu, ok);
          const float* bias = p.ep.bias + (int64_t)b * p.ep.bias_bs;
          const float bias_lo = ok ? bias[n] : 0.f, bias_hi = ok ? bias[p.ep.half + n] : 0.f;
          const int lim = live_rows(p.ep, b);
          for (int col = col_begin; col < col_end; col += 32) {
            const int t = t0 + col;
            const int nv = min(32, p.ep.out_rows - t);
            if (nv <= 0) break;
            epi_gate_cols<OPF>(p.ep, b, t, nv, lim - t, ok ? n : 0, ok, all_ok, bias_lo, bias_hi, tbase + (uint32_t)col,
                               tbase + (uint32_t)(PN + col));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_tmem_empty + 8 * buf);
      ++ait;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer may still multicast into / arrive on this CTA's barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ACC_COLS) : "memory");
  }
}

template <int OPF, int EPI, bool SUM>
int launch2(const Tc2Params& p, int grid, size_t smem, cudaStream_t stream) {
  static std::atomic<bool> attr_done[MAX_DEVICES];
  if (first_use_on_device(attr_done))
    QVC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<OPF, EPI, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
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
  QVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<OPF, EPI, SUM>, p));
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e1, stream));
  return post_launch("conv_tc2_kernel");
}

// fills p->src[si] (tensor maps, taps) for source `a`; QVC_ERR_UNSUPPORTED = not a CTA-pair case
int fill_source(const qvc_conv_args& a, int half_n, Src2* S, int* halo_out) {
  EncodeTiledFn encode = tc_get_encode();
  if (!encode) return QVC_ERR_UNSUPPORTED;
  const int esize = (int)opformat_bytes(a.opformat);
  const int kc = ROW_BYTES / esize;
  if (a.cin % kc) return QVC_ERR_UNSUPPORTED;
  const int halo = (a.k - 1) * a.dil;
  if (half_n + halo > 256) return QVC_ERR_UNSUPPORTED;
  QVC_REQUIRE((a.x.ld * esize) % 16 == 0 && ((uintptr_t)a.x.ptr & 15) == 0 && ((uintptr_t)a.w & 15) == 0,
              "conv1d(tcgen05): x / w must be 16-byte aligned with 16-byte row pitch");
  QVC_REQUIRE(a.batch == 1 || (a.x.bstride * esize) % 16 == 0, "conv1d(tcgen05): utterance pitch not 16-byte aligned");
  S->cin = a.cin; S->k = a.k; S->dil = a.dil; S->pad_left = a.pad_left;
  // structured-zero hint: a pair tile holds both output halves, so a (tap, input half) block is skipped only when it
  // is zero for both of them
  S->split_chunk = 1 << 30;
  S->jlo[0] = S->jlo[1] = 0;
  S->jhi[0] = S->jhi[1] = a.k - 1;
  if (a.tap_split > 0 && a.tap_split % kc == 0 && a.tap_split < a.cin && a.epilogue == QVC_EPI_LINEAR) {
    S->split_chunk = a.tap_split / kc;
    for (int q = 0; q < 2; ++q) {
      S->jlo[q] = a.tap_lo[0][q] < a.tap_lo[1][q] ? a.tap_lo[0][q] : a.tap_lo[1][q];
      S->jhi[q] = a.tap_hi[0][q] > a.tap_hi[1][q] ? a.tap_hi[0][q] : a.tap_hi[1][q];
      QVC_REQUIRE(S->jlo[q] >= 0 && S->jhi[q] < a.k && S->jlo[q] <= S->jhi[q], "conv1d: bad tap range [%d, %d] for k = %d", S->jlo[q], S->jhi[q], a.k);
    }
  }
  S->slab_box_rows = (half_n + halo + 7) & ~7;
  *halo_out = halo;
  const CUtensorMapDataType dt = a.opformat == QVC_OPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : (a.opformat == QVC_OPF_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.cin, (cuuint64_t)a.x_rows, (cuuint64_t)a.batch};
    cuuint64_t strides[2] = {(cuuint64_t)a.x.ld * esize,
                             (cuuint64_t)(a.batch > 1 ? a.x.bstride : (int64_t)a.x_rows * a.x.ld) * esize};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)S->slab_box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&S->mx, dt, 3, a.x.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05, pairs): cuTensorMapEncodeTiled(x) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.k * a.cin, (cuuint64_t)a.cout};
    cuuint64_t strides[1] = {(cuuint64_t)a.k * a.cin * esize};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)CHUNK_M};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&S->mw, dt, 2, const_cast<void*>(a.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05, pairs): cuTensorMapEncodeTiled(w) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  return QVC_OK;
}

}  // namespace

// Sum of nsrc >= 1 convolutions on CTA pairs (nsrc == 1, sum == false: an ordinary convolution).
// Returns QVC_ERR_UNSUPPORTED (without touching the error string) when the layer is not a CTA-pair case.
int launch_conv_tc2_sum(const qvc_conv_args* const* srcs, int nsrc, bool sum, cudaStream_t stream) {
  const qvc_conv_args& a = *srcs[0];
  if (!tc_env_int("QVC_TC_2CTA", 1)) return QVC_ERR_UNSUPPORTED;
  if (a.epilogue != QVC_EPI_LINEAR && a.epilogue != QVC_EPI_GATE) return QVC_ERR_UNSUPPORTED;
  const bool gate = a.epilogue == QVC_EPI_GATE;
  // GATE: CTA r owns channels [128 r, 128 r + 128) of the H = cout / 2 gate channels and holds two accumulators
  // (tanh half, sigmoid half) of 128 frames; LINEAR: CTA r owns chunk 2 g + r and one accumulator of 256 frames.
  const int span = gate ? a.cout / 2 : a.cout;
  const int chunks = (span + CHUNK_M - 1) / CHUNK_M;
  if (gate ? chunks != 2 : (chunks < 2 || (chunks & 1) || chunks / 2 > MAXGROUPS2)) return QVC_ERR_UNSUPPORTED;
  const int pair_n = gate ? (tc_env_int("QVC_TC_GATE_N", 128) == 256 ? 256 : 128) : 256, half_n = pair_n / 2;
  if (a.out_rows <= half_n) return QVC_ERR_UNSUPPORTED;                       // short series: conv_tc packs chunks instead
  if (!gate && a.nseg == 2 && (a.seg[1].col0 % 32 || a.seg[0].col0 % 32 || a.seg[0].ncols % 32)) return QVC_ERR_UNSUPPORTED;

  Tc2Params p{};
  QVC_PROPAGATE(build_epi_params(a, &p.ep));
  if (sum) QVC_PROPAGATE(add_sum_sources(srcs, nsrc, &p.ep));
  p.nsrc = nsrc;
  p.batch = a.batch;
  int max_box = 0;
  for (int si = 0; si < nsrc; ++si) {
    int halo = 0;
    const int st = fill_source(*srcs[si], half_n, &p.src[si], &halo);
    if (st != QVC_OK) return st;
    if (p.src[si].slab_box_rows > max_box) max_box = p.src[si].slab_box_rows;
  }
  p.pair_n = pair_n;
  p.nacc = gate ? 2 : 1;
  p.nbuf = p.nacc * pair_n > ACC_COLS ? 1 : 2;
  p.ngroups = chunks / 2;
  for (int gi = 0; gi < p.ngroups; ++gi)
    for (int r = 0; r < 2; ++r) {
      const int r0 = (2 * gi + r) * CHUNK_M;
      p.row0[gi][r] = r0;
      p.valid[gi][r] = span - r0 < CHUNK_M ? span - r0 : CHUNK_M;
    }
  p.ntb = (a.out_rows + pair_n - 1) / pair_n;
  p.ntiles = a.batch * p.ntb * p.ngroups;
  // too few pair tiles to occupy the machine: conv_tc with narrower tiles has the shorter critical path
  if (p.ntiles < tc_sm_count() / 4 && !tc_env_int("QVC_TC_2CTA_FORCE", 0)) return QVC_ERR_UNSUPPORTED;
  p.slab_stage_bytes = (uint32_t)max_box * ROW_BYTES;
  static const int stage_options[][2] = {{3, 8}, {3, 6}, {2, 6}, {2, 4}, {2, 3}, {2, 2}};
  size_t smem = 0;
  bool fits = false;
  for (const auto& opt : stage_options) {
    smem = (size_t)opt[0] * p.slab_stage_bytes + (size_t)opt[1] * p.nacc * CHUNK_BYTES + 1024 + 256 + live_cache_bytes(a);
    if (smem <= (size_t)MAX_SMEM) { p.slab_stages = opt[0]; p.w_stages = opt[1]; fits = true; break; }
  }
  if (!fits) return QVC_ERR_UNSUPPORTED;

  int pairs = tc_sm_count() / 2;
  if (p.ntiles < pairs) pairs = p.ntiles;
  const int grid_env = tc_env_int("QVC_TC_GRID", 0);
  if (grid_env >= 2 && grid_env / 2 < pairs) pairs = grid_env / 2;
  if (gate) {
    if (a.opformat == QVC_OPF_BF16) return launch2<QVC_OPF_BF16, QVC_EPI_GATE, false>(p, 2 * pairs, smem, stream);
    if (a.opformat == QVC_OPF_F16) return launch2<QVC_OPF_F16, QVC_EPI_GATE, false>(p, 2 * pairs, smem, stream);
    return launch2<QVC_OPF_TF32, QVC_EPI_GATE, false>(p, 2 * pairs, smem, stream);
  }
  if (sum) {
    if (a.opformat == QVC_OPF_BF16) return launch2<QVC_OPF_BF16, QVC_EPI_LINEAR, true>(p, 2 * pairs, smem, stream);
    if (a.opformat == QVC_OPF_F16) return launch2<QVC_OPF_F16, QVC_EPI_LINEAR, true>(p, 2 * pairs, smem, stream);
    return launch2<QVC_OPF_TF32, QVC_EPI_LINEAR, true>(p, 2 * pairs, smem, stream);
  }
  if (a.opformat == QVC_OPF_BF16) return launch2<QVC_OPF_BF16, QVC_EPI_LINEAR, false>(p, 2 * pairs, smem, stream);
  if (a.opformat == QVC_OPF_F16) return launch2<QVC_OPF_F16, QVC_EPI_LINEAR, false>(p, 2 * pairs, smem, stream);
  return launch2<QVC_OPF_TF32, QVC_EPI_LINEAR, false>(p, 2 * pairs, smem, stream);
}

int launch_conv_tc2(const qvc_conv_args& a, cudaStream_t stream) {
  const qvc_conv_args* one[1] = {&a};
  return launch_conv_tc2_sum(one, 1, false, stream);
}

}  // namespace qvc
