// conv_tc.cu -- series convolution as an implicit GEMM on the 5th-generation tensor cores
// (QVC_BACKEND_TCGEN05): TMA -> shared memory -> tcgen05.mma -> TMEM -> fused epilogue.
//
//   D[n][t] = sum_{j<k} sum_{c<cin}  W[n][j][c] * X[t + j*dil - pad][c]
//   A (M = 128 output channels per MMA) = folded filter rows [n][j][c], K-major, 128-byte rows
//   B (N = up to 256 frames)            = activation rows, K-major (channels contiguous)
//   one K block = one filter tap j x one 128-byte channel chunk (32 TF32 / 64 bf16 channels)
//
// Output channels sit on the TMEM lanes and frames on the TMEM columns.  An epilogue thread therefore
// owns one output channel and walks over frames, so that a warp touches 32 consecutive channels of
// one frame per instruction: every residual / skip / noise load and every raw / operand store of the
// series-major [utterance][frame][channel] tensors is a fully used 128-byte line.  (The v1 kernel had
// frames on the lanes; its epilogue touched 32 different lines per instruction and ran at 10-20 % of
// the tensor pipe, profiles/r01_v1_summary.md.)
//
// What makes it a convolution rather than im2col + GEMM: the activation chunk is brought in ONCE per
// channel chunk as a "slab" of (N + (k-1)*dil) consecutive frames (TMA, 128B swizzle, rows outside
// the utterance zero-filled by the tensor map => the reference's zero padding, no cross-utterance
// leakage); the k taps are k shared-memory descriptors into the same slab, each shifted by j*dil rows.
//
// Persistent CTAs (one per SM) walk over tiles (utterance, frame block, output-channel group); TMEM
// holds two accumulator sets of 256 columns so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 =
// epilogue (two warps per TMEM lane quarter, each taking half of the tile's frames).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

#include "conv_tc_common.cuh"

namespace qvc {

namespace {

using namespace tc;

constexpr int MAXG = 4;                   // output-channel chunks per tile
constexpr int MAXGROUPS = 10;             // output-channel groups per layer

constexpr int MAXSRC = QVC_MAX_SUM_SOURCES;

// one (input series, filter) pair; an ordinary convolution has one, qvc_conv1d_sum up to MAXSRC
struct alignas(64) Src1 {
  CUtensorMap mx;                          // x as (channel, frame, utterance)
  CUtensorMap mw;                          // w as (tap*cin + channel, output channel)
  int32_t cin, k, dil, pad_left;
  int32_t slab_boxes, slab_box_rows;       // slab = slab_boxes TMA boxes of slab_box_rows frames
  // Structured zeros of the filter (qvc_conv_args.tap_split; only read by the TAPS = true instances): channel chunks
  // from split_chunk on are input half q = 1; taps[h] packs, for output half h (2 = both), the tap range of input half q
  // as nibbles: lo at bits 8q, hi at bits 8q + 4.
  int32_t split_chunk;
  uint32_t taps[3];
};

struct alignas(64) TcParams {
  Src1 src[MAXSRC];
  int32_t nsrc;
  int32_t ntime;                           // N: frames per tile (multiple of 32, <= 256)
  int32_t ntb;                             // frame blocks per utterance
  int32_t ngroups, ntiles;
  int32_t gsize[MAXGROUPS];                // chunks of each group
  int32_t row0[MAXGROUPS][MAXG];           // filter row (GEMM column n) on lane 0 of each chunk
  int32_t valid[MAXGROUPS][MAXG];          // live lanes of each chunk
  int32_t slab_stages, w_stages;
  uint32_t slab_stage_bytes, w_stage_bytes;
  int32_t debug;                           // diagnostics only (QVC_TC_DEBUG): 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue I/O
  int32_t cout_half;                       // TAPS: first filter row of output half 1
  int32_t batch;
  EpiParams ep;
};

// ----------------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------------
// SUM (qvc_conv1d_sum): the K loop runs over p.nsrc sources into one accumulator, multi-residual epilogue; every other
// instance has exactly one source (the loops over sources collapse at compile time).
template <int OPF, int EPI, bool TAPS, bool SUM>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int ESIZE = opf_is16(OPF) ? 2 : 4;
  constexpr int KC = ROW_BYTES / ESIZE;            // channels per K block
  constexpr uint32_t FMT = mma_format(OPF);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slab0 = smem_base;
  const uint32_t w0 = slab0 + p.slab_stages * p.slab_stage_bytes;
  const uint32_t bar0 = w0 + p.w_stages * p.w_stage_bytes;
  // barrier layout: full_slab[SS] empty_slab[SS] full_w[WS] empty_w[WS] tmem_full[2] tmem_empty[2], TMEM base word
  const uint32_t full_slab = bar0, empty_slab = full_slab + 8 * p.slab_stages;
  const uint32_t full_w = empty_slab + 8 * p.slab_stages, empty_w = full_w + 8 * p.w_stages;
  const uint32_t tmem_full = empty_w + 8 * p.w_stages, tmem_empty = tmem_full + 16;
  const uint32_t tmem_slot = tmem_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  int32_t* live_s = reinterpret_cast<int32_t*>(tmem_slot_ptr + 4);       // ragged batches: copy of ep.live

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.ntime;
  const int nsrc = SUM ? p.nsrc : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * p.slab_stages + 2 * p.w_stages + 2; ++i) mbar_init(bar0 + 8 * i, 1);
    mbar_init(tmem_empty, N_EPI_WARPS);
    mbar_init(tmem_empty + 8, N_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * ACC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as every CTA
  // of this grid got here (its barrier init / TMEM allocation / tensor-map prefetch then overlap our main
  // loop's tail); nothing above touched global memory, everything below waits for the previous grid.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.ep.live != nullptr) {
    load_live_cache(p.ep, p.batch, live_s);
    __syncthreads();
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int si = 0; si < nsrc; ++si) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.src[si].mx) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.src[si].mw) : "memory");
      }
      // stage indices / phase parities are carried incrementally: a runtime integer division per K block
      // on this single thread (I2F / MUFU.RCP chains) cost more than the MMAs it feeds
      uint32_t s = 0, ph = 0, ws = 0, wph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int gi = tile % p.ngroups;
        const int rest = tile / p.ngroups;
        const int tb = rest % p.ntb, b = rest / p.ntb;
        const int t0 = tb * N;
        if (tile_dead(p.ep, live_s, b, t0)) continue;
        const int gs = p.gsize[gi];
        for (int si = 0; si < nsrc; ++si) {
          const Src1& S = p.src[si];
          const int n_cchunks = S.cin / KC;
          const uint32_t slab_bytes = (uint32_t)S.slab_boxes * S.slab_box_rows * ROW_BYTES;
          uint32_t tp = 0;                           // TAPS: tap ranges of this tile's output half (union if it spans both)
          if constexpr (TAPS) tp = S.taps[gs > 1 ? 2 : (p.row0[gi][0] >= p.cout_half ? 1 : 0)];
          for (int cc = 0; cc < n_cchunks; ++cc) {
            mbar_wait(empty_slab + 8 * s, ph ^ 1u);
            if (p.debug & 1) mbar_arrive(full_slab + 8 * s);
            else             mbar_expect_tx(full_slab + 8 * s, slab_bytes);
            for (int i = 0; i < S.slab_boxes && !(p.debug & 1); ++i)
              tma_load_3d(slab0 + s * p.slab_stage_bytes + i * S.slab_box_rows * ROW_BYTES, &S.mx, full_slab + 8 * s,
                          cc * KC, t0 - S.pad_left + i * S.slab_box_rows, b);
            if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
            int jbeg = 0, jend = S.k - 1;
            if constexpr (TAPS) {
              const uint32_t r = cc >= S.split_chunk ? tp >> 8 : tp;
              jbeg = (int)(r & 15u); jend = (int)((r >> 4) & 15u);
            }
            for (int j = jbeg; j <= jend; ++j) {
              mbar_wait(empty_w + 8 * ws, wph ^ 1u);
              if (p.debug & 1) mbar_arrive(full_w + 8 * ws);
              else             mbar_expect_tx(full_w + 8 * ws, (uint32_t)gs * CHUNK_BYTES);
              for (int ci = 0; ci < gs && !(p.debug & 1); ++ci)
                tma_load_2d(w0 + ws * p.w_stage_bytes + ci * CHUNK_BYTES, &S.mw, full_w + 8 * ws,
                            j * S.cin + cc * KC, p.row0[gi][ci]);
              if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the pipeline (waits are warp-wide); one elected lane issues the MMAs and commits.
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(CHUNK_M >> 4) << 24);
    const uint64_t desc_hi = smem_desc(0);         // everything but the start address
    // Stage indices / phase parities are carried incrementally (no runtime division on this path).
    // Measured on B200 (scripts/conv_bench.py with QVC_TC_DEBUG): this loop is not the limiter -- an
    // M = 128, N = 256 MMA with both operands in shared memory takes ~190 cycles instead of 128, i.e. the
    // 12 KB of operand reads per MMA run at ~64 B/cycle; batching several K blocks per trip changes nothing.
    uint32_t s = 0, ph = 0, ws = 0, wph = 0, ait = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int gi = tile % p.ngroups;
      {
        const int rest = tile / p.ngroups;
        if (tile_dead(p.ep, live_s, rest / p.ntb, (rest % p.ntb) * N)) continue;
      }
      const int gs = (p.debug & 2) ? 0 : p.gsize[gi];
      const uint32_t buf = ait & 1u;
      mbar_wait(tmem_empty + 8 * buf, ((ait >> 1) & 1u) ^ 1u);     // epilogue drained this accumulator set
      tc_fence_after();
      const uint32_t dbase = tmem_base + buf * ACC_COLS;
      for (int si = 0; si < nsrc; ++si) {
        const Src1& S = p.src[si];
        const int n_cchunks = S.cin / KC;
        uint32_t tp = 0;
        if constexpr (TAPS) tp = S.taps[p.gsize[gi] > 1 ? 2 : (p.row0[gi][0] >= p.cout_half ? 1 : 0)];
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(full_slab + 8 * s, ph);
          const uint32_t slab = slab0 + s * p.slab_stage_bytes;
          int jbeg = 0, jlast = S.k - 1, jfirst = 0;
          if constexpr (TAPS) {
            const uint32_t r = cc >= S.split_chunk ? tp >> 8 : tp;
            jbeg = (int)(r & 15u); jlast = (int)((r >> 4) & 15u); jfirst = (int)(tp & 15u);
          }
          for (int j = jbeg; j <= jlast; ++j) {
            mbar_wait(full_w + 8 * ws, wph);
            tc_fence_after();
            const uint32_t wst = w0 + ws * p.w_stage_bytes;
            const uint32_t first = (si == 0 && cc == 0 && j == jfirst) ? 0u : 1u;
            const uint64_t bdesc = desc_hi | (uint64_t)(((slab + (uint32_t)(j * S.dil) * ROW_BYTES) & 0x3FFFFu) >> 4);
            if (elect_one()) {
              for (int ci = 0; ci < gs; ++ci) {
                const uint64_t adesc = desc_hi | (uint64_t)(((wst + (uint32_t)ci * CHUNK_BYTES) & 0x3FFFFu) >> 4);
                const uint32_t d = dbase + (uint32_t)(ci * N);
                if (p.debug & 16) {
                  // timing experiment only (garbage A): A operand from TMEM -- the other accumulator set
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_ts<OPF>(d, tmem_base + (buf ^ 1u) * ACC_COLS + 8 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
                } else {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)      // +32 bytes per K step = +2 in the (addr >> 4) field
                    umma<OPF>(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
                }
              }
              tc_commit(empty_w + 8 * ws);            // filter stage free once these MMAs retire
              if (j == jlast) tc_commit(empty_slab + 8 * s);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
          }
          if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
        }
      }
      if (elect_one()) tc_commit(tmem_full + 8 * buf);
      __syncwarp();
      ++ait;
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int h = (warp - 2) >> 2;                 // which half of the tile's frames
    const int nblk = N >> 5;
    const int col_begin = h == 0 ? 0 : ((nblk + 1) >> 1) << 5;
    const int col_end = h == 0 ? ((nblk + 1) >> 1) << 5 : N;
    const int lic = q * 32 + lane;                 // lane within the 128-channel chunk
    uint32_t ait = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int gi = tile % p.ngroups;
      const int rest = tile / p.ngroups;
      const int tb = rest % p.ntb, b = rest / p.ntb;
      const int t0 = tb * N;
      if (tile_dead(p.ep, live_s, b, t0)) continue;
      const int gs = p.gsize[gi];
      const uint32_t buf = ait & 1u;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS;
      const int lim = live_rows(p.ep, b);
      if constexpr (EPI == QVC_EPI_LINEAR) {
        auto context = [&](int ci, LinCtx& k) -> bool {            // false: no live channel in this lane quarter
          const int nvalid = p.valid[gi][ci];
          if (q * 32 >= nvalid) return false;
          const int n_w = p.row0[gi][ci] + q * 32;
          const int n = n_w + lane;
          k.sg = &p.ep.seg[(p.ep.nseg > 1 && n_w >= p.ep.seg[1].col0) ? 1 : 0];
          const int c = n - k.sg->col0;
          k.ok = lic < nvalid && c >= 0 && c < k.sg->ncols;
          k.all_ok = __all_sync(0xffffffffu, k.ok);
          k.c = k.ok ? c : 0;
          k.b = b;
          k.lim = lim;
          if constexpr (SUM) k.bias = sum_bias(p.ep, b, n, k.ok);
          else               k.bias = (p.ep.bias && k.ok) ? p.ep.bias[(int64_t)b * p.ep.bias_bs + n] : 0.f;
          return true;
        };
        constexpr int SB = SUM ? 32 : 64;                          // frames per superblock (SUM: three residual streams)
        auto frames_at = [&](int col) -> int {                     // live frames of the superblock starting at col
          if (p.debug & 4) return 0;
          const int left = min(col_end - col, p.ep.out_rows - (t0 + col));
          return left < SB ? left : SB;
        };
        float r[SUM ? 96 : 64];
        LinCtx k;
        // first superblock of the tile: loads in flight before the accumulator is complete
        bool primed = false;
        if (context(0, k)) {
          const int nv = frames_at(col_begin);
          if (nv > 0) {
            if constexpr (SUM) sum_load<OPF>(p.ep, k, t0 + col_begin, nv, r);
            else               lin_load<OPF>(k, t0 + col_begin, nv, r);
          }
          primed = true;
        }
        mbar_wait(tmem_full + 8 * buf, (ait >> 1) & 1u);
        tc_fence_after();
        for (int ci = 0; ci < gs; ++ci) {
          if (!context(ci, k)) continue;
          for (int col = col_begin; col < col_end; col += SB) {
            const int nv = frames_at(col);
            if (nv <= 0) break;
            if constexpr (SUM) {
              if (!(primed && ci == 0 && col == col_begin)) sum_load<OPF>(p.ep, k, t0 + col, nv, r);
              sum_finish<OPF>(p.ep, k, t0 + col, nv, r, tbase + (uint32_t)(ci * N + col));
            } else {
              if (!(primed && ci == 0 && col == col_begin)) lin_load<OPF>(k, t0 + col, nv, r);
              lin_finish<OPF>(k, t0 + col, nv, r, tbase + (uint32_t)(ci * N + col));
            }
          }
        }
      } else {
        mbar_wait(tmem_full + 8 * buf, (ait >> 1) & 1u);
        tc_fence_after();
        const int nlo = gs >> 1;
        for (int ci = 0; ci < nlo; ++ci) {
          const int nvalid = p.valid[gi][ci];
          if (q * 32 >= nvalid) continue;
          const int n = p.row0[gi][ci] + lic;
          const bool ok = lic < nvalid;
          const bool all_ok = __all_sync(0xffffffffu, ok);
          const float* bias = p.ep.bias + (int64_t)b * p.ep.bias_bs;
          const float bias_lo = ok ? bias[n] : 0.f, bias_hi = ok ? bias[p.ep.half + n] : 0.f;
          for (int col = col_begin; col < col_end; col += 32) {
            const int t = t0 + col;
            const int nv = min(32, p.ep.out_rows - t);
            if (nv <= 0) break;
            const uint32_t ta_lo = tbase + (uint32_t)(ci * N + col), ta_hi = tbase + (uint32_t)((ci + nlo) * N + col);
            if constexpr (EPI == QVC_EPI_GATE) epi_gate_cols<OPF>(p.ep, b, t, nv, lim - t, ok ? n : 0, ok, all_ok, bias_lo, bias_hi, ta_lo, ta_hi);
            else                               epi_sample_cols<OPF>(p.ep, b, t, nv, lim - t, ok ? n : 0, ok, bias_lo, bias_hi, ta_lo, ta_hi);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(tmem_empty + 8 * buf);
      ++ait;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ACC_COLS) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int sm_count() {                                  // of the current device, cached per device
  static std::atomic<int> cache[MAX_DEVICES];
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 148;
  v = cache[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// ---- optional per-launch timing (qvc_profile / qvc_profile_read): CUDA events on the launching stream ----
struct ProfState {
  std::mutex mu;
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t used = 0;
};
ProfState& prof() {
  static ProfState st;
  return st;
}
// returns the event pair to record around the next launch, or false when profiling is off
bool prof_next(cudaEvent_t* e0, cudaEvent_t* e1) {
  ProfState& st = prof();
  std::lock_guard<std::mutex> lk(st.mu);
  if (!st.on) return false;
  if (st.used == st.ev.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return false;
    st.ev.emplace_back(a, b);
  }
  *e0 = st.ev[st.used].first;
  *e1 = st.ev[st.used].second;
  ++st.used;
  return true;
}

template <int OPF, int EPI, bool TAPS, bool SUM = false>
int launch_variant(const TcParams& p, int grid, size_t smem, cudaStream_t stream) {
  static std::atomic<bool> attr_done[MAX_DEVICES];
  if (first_use_on_device(attr_done))
    QVC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<OPF, EPI, TAPS, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_int("QVC_TC_PDL", 1) ? 1 : 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = prof_next(&e0, &e1);
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e0, stream));
  QVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<OPF, EPI, TAPS, SUM>, p));
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e1, stream));
  return post_launch("conv_tc_kernel");
}

}  // namespace

EncodeTiledFn tc_get_encode() { return get_encode(); }
// SMs the persistent convolution grids leave free (per host thread): while the speaker encoder's cluster runs on its
// side stream, a grid of one CTA per SM would leave 8 CTAs waiting for it -- with static tile assignment their tiles
// would finish ~0.5 ms late (measured: +0.45 ms per step tf32, +0.65 ms bf16, scripts/spk_overlap.py).
thread_local int g_reserved_sms = 0;
void tc_reserve_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }
int tc_sm_count() {
  const int n = sm_count() - g_reserved_sms;
  return n < 2 ? 2 : n;
}
int tc_env_int(const char* name, int dflt) { return env_int(name, dflt); }
bool tc_prof_next(cudaEvent_t* e0, cudaEvent_t* e1) { return prof_next(e0, e1); }

// Sum of nsrc >= 1 convolutions into one accumulator (nsrc == 1, sum == false: an ordinary convolution; srcs[0]
// carries the epilogue).  Layers with an even number of 128-channel chunks and long series go to the CTA-pair kernel.
int launch_conv_tc_any(const qvc_conv_args* const* srcs, int nsrc, bool sum, cudaStream_t stream) {
  const qvc_conv_args& a = *srcs[0];
  if (!sum && nsrc == 1) {
    const int st = launch_conv_tcr(a, stream);
    if (st != QVC_ERR_UNSUPPORTED) return st;
  }
  {
    const int st = launch_conv_tc2_sum(srcs, nsrc, sum, stream);
    if (st != QVC_ERR_UNSUPPORTED) return st;
  }
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("conv1d(tcgen05): cuTensorMapEncodeTiled not available from the driver");
    return QVC_ERR_CUDA;
  }
  const int esize = (int)opformat_bytes(a.opformat);
  const int kc = ROW_BYTES / esize;
  for (int si = 0; si < nsrc; ++si) {
    const qvc_conv_args& x = *srcs[si];
    QVC_REQUIRE(x.cin % kc == 0, "conv1d(tcgen05): cin %d not a multiple of %d", x.cin, kc);
    QVC_REQUIRE((x.x.ld * esize) % 16 == 0 && ((uintptr_t)x.x.ptr & 15) == 0 && ((uintptr_t)x.w & 15) == 0,
                "conv1d(tcgen05): x / w must be 16-byte aligned with 16-byte row pitch");
    QVC_REQUIRE(x.batch == 1 || (x.x.bstride * esize) % 16 == 0, "conv1d(tcgen05): utterance pitch not 16-byte aligned");
  }

  TcParams p{};
  QVC_PROPAGATE(build_epi_params(a, &p.ep));
  if (sum) QVC_PROPAGATE(add_sum_sources(srcs, nsrc, &p.ep));
  p.nsrc = nsrc;
  p.batch = a.batch;
  p.debug = env_int("QVC_TC_DEBUG", 0);

  // ---- output-channel chunks (TMEM lanes) and groups ----
  const bool paired = a.epilogue != QVC_EPI_LINEAR;
  const int rows32 = ((a.out_rows + 31) / 32) * 32;
  int g;                                           // chunks per tile
  if (paired) {
    // group i = { lo chunk i, hi chunk i }: both members of a gate pair on the same TMEM lane, two
    // chunks per tile so that a tile spans 128 frames (weights re-streamed per 128 frames, not 64)
    const int half = a.cout / 2;
    const int nlo = (half + CHUNK_M - 1) / CHUNK_M;
    QVC_REQUIRE(nlo <= MAXGROUPS, "conv1d(tcgen05): paired epilogue supports cout <= %d (got %d)", 2 * MAXGROUPS * CHUNK_M, a.cout);
    g = 2;
    p.ngroups = nlo;
    for (int i = 0; i < nlo; ++i) {
      p.gsize[i] = 2;
      p.row0[i][0] = i * CHUNK_M;        p.valid[i][0] = half - i * CHUNK_M < CHUNK_M ? half - i * CHUNK_M : CHUNK_M;
      p.row0[i][1] = half + i * CHUNK_M; p.valid[i][1] = p.valid[i][0];
    }
  } else {
    if (a.nseg == 2)
      QVC_REQUIRE(a.seg[1].col0 % 32 == 0 && a.seg[0].col0 % 32 == 0 && a.seg[0].ncols % 32 == 0,
                  "conv1d(tcgen05): segment boundaries must be multiples of 32 columns");
    const int chunks = (a.cout + CHUNK_M - 1) / CHUNK_M;
    // One chunk x 256 frames per tile: measured on B200 (scripts/conv_bench.py) an N = 256 MMA costs about
    // the same issue slot as an N = 128 one, so wide-N tiles win even though the slab is re-read once per
    // chunk.  Short series put more chunks on a tile instead of frames.
    g = 1;
    {
      // ... but only when there are tiles to spare: with few tiles (single clips, streaming chunks) one chunk per
      // tile keeps the serial MMA chain of a CTA short and spreads the chunks over more SMs
      const long tiles_g1 = (long)a.batch * ((a.out_rows + 255) / 256) * chunks;
      if (tiles_g1 >= 2L * sm_count()) {
        if (rows32 <= 128 && chunks >= 2) g = 2;
        if (rows32 <= 64 && chunks >= 4) g = 4;
      }
    }
    const int g_env = env_int("QVC_TC_G", 0);
    if (g_env >= 1 && g_env <= MAXG && g_env <= chunks) g = g_env;
    p.ngroups = (chunks + g - 1) / g;
    QVC_REQUIRE(p.ngroups <= MAXGROUPS, "conv1d(tcgen05): cout %d needs more than %d channel groups", a.cout, MAXGROUPS);
    for (int gi = 0; gi < p.ngroups; ++gi) {
      const int first = gi * g;
      p.gsize[gi] = chunks - first < g ? chunks - first : g;
      for (int ci = 0; ci < p.gsize[gi]; ++ci) {
        const int r0 = (first + ci) * CHUNK_M;
        p.row0[gi][ci] = r0;
        p.valid[gi][ci] = a.cout - r0 < CHUNK_M ? a.cout - r0 : CHUNK_M;
      }
    }
  }
  int ntime = (ACC_COLS / g) & ~31;
  if (rows32 < ntime) ntime = rows32;
  // Small problems (single clips, streaming chunks): a tile's K loop is a serial chain of MMAs whose cost
  // barely shrinks with N (N = 64: ~94 cycles, N = 256: ~190), so narrower tiles on more SMs cut latency.
  {
    const int groups = paired ? p.ngroups : (( (a.cout + CHUNK_M - 1) / CHUNK_M) + g - 1) / g;
    while (ntime > 64 && (long)a.batch * ((a.out_rows + ntime - 1) / ntime) * groups < sm_count()) ntime = (ntime / 2 + 31) & ~31;
  }
  const int n_env = env_int("QVC_TC_N", 0);
  if (n_env >= 32 && n_env % 32 == 0 && n_env <= ntime) ntime = n_env;
  p.ntime = ntime;
  p.ntb = (a.out_rows + ntime - 1) / ntime;
  p.ntiles = a.batch * p.ntb * p.ngroups;

  // ---- shared-memory pipeline: one slab ring serves every source, sized for the tallest slab ----
  p.slab_stage_bytes = 0;
  for (int si = 0; si < nsrc; ++si) {
    const qvc_conv_args& x = *srcs[si];
    Src1& S = p.src[si];
    S.cin = x.cin; S.k = x.k; S.dil = x.dil; S.pad_left = x.pad_left;
    const int rows = ntime + (x.k - 1) * x.dil;
    S.slab_boxes = (rows + 255) / 256;
    S.slab_box_rows = (((rows + S.slab_boxes - 1) / S.slab_boxes) + 7) & ~7;
    QVC_REQUIRE(S.slab_box_rows <= 256, "conv1d(tcgen05): slab box too tall");
    const uint32_t bytes = (uint32_t)S.slab_boxes * S.slab_box_rows * ROW_BYTES;
    if (bytes > p.slab_stage_bytes) p.slab_stage_bytes = bytes;
  }
  p.w_stage_bytes = (uint32_t)g * CHUNK_BYTES;
  static const int stage_options[][2] = {{3, 6}, {3, 4}, {2, 4}, {2, 3}, {2, 2}, {1, 2}};
  size_t smem = 0;
  bool fits = false;
  const int ss_env = env_int("QVC_TC_SS", 0), ws_env = env_int("QVC_TC_WS", 0);
  if (ss_env >= 1 && ws_env >= 1 && ss_env <= 8 && ws_env <= 12 && ws_env >= 2) {
    smem = (size_t)ss_env * p.slab_stage_bytes + (size_t)ws_env * p.w_stage_bytes + 1024 + 256 + live_cache_bytes(a);
    if (smem <= (size_t)MAX_SMEM) { p.slab_stages = ss_env; p.w_stages = ws_env; fits = true; }
  }
  for (const auto& opt : stage_options) {
    if (fits) break;
    smem = (size_t)opt[0] * p.slab_stage_bytes + (size_t)opt[1] * p.w_stage_bytes + 1024 + 256 + live_cache_bytes(a);
    if (smem <= (size_t)MAX_SMEM) { p.slab_stages = opt[0]; p.w_stages = opt[1]; fits = true; }
  }
  if (!fits) {
    set_error("conv1d(tcgen05): tile does not fit shared memory (k=%d dil=%d cout=%d)", a.k, a.dil, a.cout);
    return QVC_ERR_UNSUPPORTED;
  }

  // ---- tensor maps ----
  const CUtensorMapDataType dt = a.opformat == QVC_OPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : (a.opformat == QVC_OPF_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  for (int si = 0; si < nsrc; ++si) {
    const qvc_conv_args& x = *srcs[si];
    Src1& S = p.src[si];
    {
      cuuint64_t dims[3] = {(cuuint64_t)x.cin, (cuuint64_t)x.x_rows, (cuuint64_t)x.batch};
      cuuint64_t strides[2] = {(cuuint64_t)x.x.ld * esize,
                               (cuuint64_t)(x.batch > 1 ? x.x.bstride : (int64_t)x.x_rows * x.x.ld) * esize};
      cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)S.slab_box_rows, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = encode(&S.mx, dt, 3, x.x.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05): cuTensorMapEncodeTiled(x) failed: %d", (int)r); return QVC_ERR_CUDA; }
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)x.k * x.cin, (cuuint64_t)x.cout};
      cuuint64_t strides[1] = {(cuuint64_t)x.k * x.cin * esize};
      cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)CHUNK_M};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&S.mw, dt, 2, const_cast<void*>(x.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05): cuTensorMapEncodeTiled(w) failed: %d", (int)r); return QVC_ERR_CUDA; }
    }
  }

  // structured-zero hint (LINEAR layers whose two output halves are whole chunks): packed tap ranges, TAPS instance.
  // A sum uses the TAPS instance whenever ANY source has a hint; sources without one get the full tap range.
  bool taps = false;
  p.cout_half = a.cout / 2;
  for (int si = 0; si < nsrc; ++si) {
    const qvc_conv_args& x = *srcs[si];
    if (x.tap_split > 0 && x.tap_split % kc == 0 && x.tap_split < x.cin && !paired && (x.cout / 2) % CHUNK_M == 0 && x.k <= 16) taps = true;
  }
  for (int si = 0; si < nsrc && taps; ++si) {
    const qvc_conv_args& x = *srcs[si];
    Src1& S = p.src[si];
    QVC_REQUIRE(x.k <= 16, "conv1d(tcgen05): tap hints need k <= 16 in every source");
    const bool has = x.tap_split > 0 && x.tap_split % kc == 0 && x.tap_split < x.cin;
    S.split_chunk = has ? x.tap_split / kc : (1 << 30);
    for (int h = 0; h < 3; ++h) {
      uint32_t w = 0;
      for (int q = 0; q < 2; ++q) {
        int lo = 0, hi = x.k - 1;
        if (has) {
          if (h < 2) { lo = x.tap_lo[h][q]; hi = x.tap_hi[h][q]; }
          else {
            lo = x.tap_lo[0][q] < x.tap_lo[1][q] ? x.tap_lo[0][q] : x.tap_lo[1][q];
            hi = x.tap_hi[0][q] > x.tap_hi[1][q] ? x.tap_hi[0][q] : x.tap_hi[1][q];
          }
        }
        QVC_REQUIRE(lo >= 0 && hi < x.k && lo <= hi, "conv1d: bad tap range [%d, %d] for k = %d", lo, hi, x.k);
        w |= ((uint32_t)lo | (uint32_t)hi << 4) << (8 * q);
      }
      S.taps[h] = w;
    }
  }
  int grid = p.ntiles < tc_sm_count() ? p.ntiles : tc_sm_count();
  const int grid_env = env_int("QVC_TC_GRID", 0);
  if (grid_env >= 1 && grid_env < grid) grid = grid_env;
#define QVC_TC_DISPATCH(OPF)                                                                         \
  if (sum) return taps ? launch_variant<OPF, QVC_EPI_LINEAR, true, true>(p, grid, smem, stream)       \
                       : launch_variant<OPF, QVC_EPI_LINEAR, false, true>(p, grid, smem, stream);     \
  switch (a.epilogue) {                                                                              \
    case QVC_EPI_LINEAR: return taps ? launch_variant<OPF, QVC_EPI_LINEAR, true>(p, grid, smem, stream)    \
                                     : launch_variant<OPF, QVC_EPI_LINEAR, false>(p, grid, smem, stream);  \
    case QVC_EPI_GATE:   return launch_variant<OPF, QVC_EPI_GATE, false>(p, grid, smem, stream);     \
    default:             return launch_variant<OPF, QVC_EPI_SAMPLE, false>(p, grid, smem, stream);   \
  }
  if (a.opformat == QVC_OPF_BF16) { QVC_TC_DISPATCH(QVC_OPF_BF16) }
  if (a.opformat == QVC_OPF_F16) { QVC_TC_DISPATCH(QVC_OPF_F16) }
  QVC_TC_DISPATCH(QVC_OPF_TF32)
#undef QVC_TC_DISPATCH
}

int launch_conv_tc(const qvc_conv_args& a, cudaStream_t stream) {
  const qvc_conv_args* one[1] = {&a};
  return launch_conv_tc_any(one, 1, false, stream);
}

int launch_conv_tc_sum(const qvc_conv_args* const* srcs, int nsrc, cudaStream_t stream) {
  return launch_conv_tc_any(srcs, nsrc, true, stream);
}

}  // namespace qvc

extern "C" int qvc_profile(int enable) {
  qvc::ProfState& st = qvc::prof();
  std::lock_guard<std::mutex> lk(st.mu);
  st.on = enable != 0;
  if (st.on) st.used = 0;
  return QVC_OK;
}

extern "C" int qvc_profile_read(double* ms_total, uint64_t* launches) {
  qvc::ProfState& st = qvc::prof();
  std::lock_guard<std::mutex> lk(st.mu);
  double tot = 0.0;
  for (size_t i = 0; i < st.used; ++i) {
    QVC_CHECK_CUDA(cudaEventSynchronize(st.ev[i].second));
    float ms = 0.f;
    QVC_CHECK_CUDA(cudaEventElapsedTime(&ms, st.ev[i].first, st.ev[i].second));
    tot += ms;
  }
  if (ms_total) *ms_total = tot;
  if (launches) *launches = st.used;
  return QVC_OK;
}

