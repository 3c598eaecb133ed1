// conv_wn.cu -- one whole WN layer (modules.py:88-112) in one kernel on CTA pairs:
//
//     acts = tanh(conv_k5(x)[:H] + b) * sigmoid(conv_k5(x)[H:] + b)        (in_layer + fused gate, modules.py:91-101)
//     rs   = W_rs . acts + b_rs                                             (res_skip 1x1, modules.py:104)
//     x   += rs[:H] ; skip += rs[H:]      (last layer: skip += rs)           (modules.py:106-112)
//
// The gated activations never leave the SM pair: the gate epilogue writes them, already in operand format,
// into shared memory in the K-major 128-byte-swizzled layout the second GEMM reads as its B operand.
// Unfused (conv_tc2 GATE + conv_tc LINEAR) a layer costs 50 + 40 us at B = 64 x 10 s; the 1x1 GEMM is tiny, so its
// launch is all epilogue traffic plus the activation round trip.
//
// Pair tile = (utterance, 128 frames); CTA r of the pair owns gate channels [128 r, 128 r + 128) in GEMM 1 and
// output rows [128 r, +128) and [256 + 128 r, +128) of the res_skip filter in GEMM 2, and holds 64 of the 128
// frames of every B operand (slab for GEMM 1, activations for GEMM 2).  Since GEMM 2 wants, per CTA, all
// channels of ITS 64 frames while the gate epilogue produced all frames of ITS channels, every epilogue warp writes
// one half of its results into the peer's shared memory (DSMEM).
//
// TMEM (512 columns): [0,128) tanh half, [128,256) sigmoid half of GEMM 1; [256,384) and [384,512) the two
// accumulators of GEMM 2.  The MMA issuer orders its work  G1(0), { G2(i), G1(i+1) } ...  so the memory-bound
// res_skip epilogue of tile i overlaps the long GEMM 1 of tile i+1.
//
// The caller must ping-pong the operand copy of x between layers: tile A's epilogue writes the new x operand while
// a neighbouring tile may still need the old one as convolution halo.
#include <cstdlib>

#include "conv_tc_common.cuh"

namespace qvc {

namespace {

using namespace tc;

constexpr int PN = 128;                   // frames per pair tile
constexpr int HN = PN / 2;                // frames whose B rows one CTA holds
constexpr uint32_t W_STAGE = 2 * CHUNK_BYTES;
constexpr uint32_t ACTS_CHUNK = HN * ROW_BYTES;     // one 128-byte channel chunk of the activations: 64 rows

struct alignas(64) WnParams {
  CUtensorMap mx;                          // x as (channel, frame, utterance)
  CUtensorMap mw_in;                       // in_layer filter as (tap*cin + channel, 2H rows)
  CUtensorMap mw_rs;                       // res_skip filter as (channel, rs rows)
  int32_t hid, k, pad_left;                // H (= cin of both GEMMs), taps of GEMM 1
  int32_t rs_cout;                         // 2H, or H on the last layer of a stack
  int32_t ntb, ntiles;
  int32_t slab_box_rows, slab_stages, w_stages;
  uint32_t slab_stage_bytes;
  int32_t out_rows, batch;
  int32_t debug;                           // timing experiments only (QVC_WN_DEBUG, results are garbage): 1 no gate epilogue body,
                                           // 2 no res_skip epilogue loads, 4 no res_skip epilogue body, 8 no GEMM 1, 16 no GEMM 2, 32 relaxed acts_ready arrive
  const float* gate_bias;                  // [2H] (+ per-utterance stride)
  int64_t gate_bias_bs;
  EpiParams ep;                            // res_skip epilogue (LINEAR, 1 or 2 segments)
};

// wait that also acquires at cluster scope: the data guarded by the barrier was written by the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void st_cluster_b32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_b16(uint32_t cluster_addr, uint16_t v) {
  asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(cluster_addr), "h"(v) : "memory");
}

template <int OPF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) conv_wn_kernel(const __grid_constant__ WnParams p) {
  constexpr int ESIZE = opf_is16(OPF) ? 2 : 4;
  constexpr int KC = ROW_BYTES / ESIZE;
  constexpr uint32_t FMT = mma_format(OPF);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int n_cchunks = p.hid / KC;
  const uint32_t slab0 = smem_base;
  const uint32_t w0 = slab0 + p.slab_stages * p.slab_stage_bytes;
  const uint32_t acts0 = w0 + p.w_stages * W_STAGE;
  const uint32_t bar0 = acts0 + (uint32_t)n_cchunks * ACTS_CHUNK;
  // barriers (same offsets in both CTAs): full_slab[SS] empty_slab[SS] full_w[WS] empty_w[WS] acc1_full acc2_full
  // | acts_ready acc2_empty (leader only, 16 arrivals each), then the TMEM base word
  const uint32_t full_slab = bar0, empty_slab = full_slab + 8 * p.slab_stages;
  const uint32_t full_w = empty_slab + 8 * p.slab_stages, empty_w = full_w + 8 * p.w_stages;
  const uint32_t acc1_full = empty_w + 8 * p.w_stages, acc2_full = acc1_full + 8;
  const uint32_t acts_ready = acc2_full + 8, acc2_empty = acts_ready + 8;
  const uint32_t tmem_slot = acc2_empty + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  int32_t* live_s = reinterpret_cast<int32_t*>(tmem_slot_ptr + 4);       // ragged batches: copy of ep.live

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * p.slab_stages + 2 * p.w_stages + 2; ++i) mbar_init(bar0 + 8 * i, 1);
    mbar_init(acts_ready, 2 * N_EPI_WARPS);
    mbar_init(acc2_empty, 2 * N_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // ragged batches: every warp role walks the same list of live tiles (tiles wholly past their utterance's end are skipped)
  if (p.ep.live != nullptr) {
    load_live_cache(p.ep, p.batch, live_s);
    __syncthreads();
  }
  auto next_live = [&](int tile) -> int {
    while (tile < p.ntiles && tile_dead(p.ep, live_s, tile / p.ntb, (tile % p.ntb) * PN)) tile += npairs;
    return tile;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // order of filter stages = order of the MMA issuer: G1(0), then per tile { G2(i), G1(i+1) }
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mx) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mw_in) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mw_rs) : "memory");
      const uint32_t slab_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
      const uint32_t lead_full_slab = map_to_cta(full_slab, 0), lead_full_w = map_to_cta(full_w, 0);
      uint32_t s = 0, ph = 0, ws = 0, wph = 0;
      auto load_g1 = [&](int tile) {
        const int tb = tile % p.ntb, b = tile / p.ntb;
        const int t0 = tb * PN + (int)rank * HN;
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(empty_slab + 8 * s, ph ^ 1u);
          if (leader) mbar_expect_tx(full_slab + 8 * s, 2 * slab_bytes);
          tma2_load_3d(slab0 + s * p.slab_stage_bytes, &p.mx, lead_full_slab + 8 * s, cc * KC, t0 - p.pad_left, b);
          if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
          for (int j = 0; j < p.k; ++j) {
            mbar_wait(empty_w + 8 * ws, wph ^ 1u);
            if (leader) mbar_expect_tx(full_w + 8 * ws, 2 * W_STAGE);
            tma2_load_2d(w0 + ws * W_STAGE, &p.mw_in, lead_full_w + 8 * ws, j * p.hid + cc * KC, (int)rank * CHUNK_M);
            tma2_load_2d(w0 + ws * W_STAGE + CHUNK_BYTES, &p.mw_in, lead_full_w + 8 * ws, j * p.hid + cc * KC,
                         p.hid + (int)rank * CHUNK_M);
            if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
          }
        }
      };
      auto load_g2 = [&]() {
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(empty_w + 8 * ws, wph ^ 1u);
          if (leader) mbar_expect_tx(full_w + 8 * ws, 2 * W_STAGE);
          tma2_load_2d(w0 + ws * W_STAGE, &p.mw_rs, lead_full_w + 8 * ws, cc * KC, (int)rank * CHUNK_M);
          tma2_load_2d(w0 + ws * W_STAGE + CHUNK_BYTES, &p.mw_rs, lead_full_w + 8 * ws, cc * KC, 2 * CHUNK_M + (int)rank * CHUNK_M);
          if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
        }
      };
      int tile = next_live(pair);
      if (tile < p.ntiles) load_g1(tile);
      while (tile < p.ntiles) {
        load_g2();
        tile = next_live(tile + npairs);
        if (tile < p.ntiles) load_g1(tile);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(PN >> 3) << 17) |
                             ((uint32_t)((2 * CHUNK_M) >> 4) << 24);
      const uint64_t desc_hi = smem_desc(0);
      uint32_t s = 0, ph = 0, ws = 0, wph = 0;
      auto gemm1 = [&]() {                          // in_layer: acc1 = [tanh half | sigmoid half]
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(full_slab + 8 * s, ph);
          const uint32_t slab = slab0 + s * p.slab_stage_bytes;
          for (int j = 0; j < p.k; ++j) {
            mbar_wait(full_w + 8 * ws, wph);
            tc_fence_after();
            const uint32_t first = (cc | j) == 0 ? 0u : 1u;
            const uint64_t bdesc = desc_hi | (uint64_t)(((slab + (uint32_t)j * ROW_BYTES) & 0x3FFFFu) >> 4);
            const uint64_t adesc = desc_hi | (uint64_t)(((w0 + ws * W_STAGE) & 0x3FFFFu) >> 4);
            if (elect_one()) {
              if (!(p.debug & 8)) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2<OPF>(tmem_base, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma2<OPF>(tmem_base + PN, adesc + (CHUNK_BYTES >> 4) + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
              }
              tc2_commit(empty_w + 8 * ws);
              if (j == p.k - 1) tc2_commit(empty_slab + 8 * s);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
          }
          if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) tc2_commit(acc1_full);
        __syncwarp();
      };
      auto gemm2 = [&]() {                          // res_skip: acc2 = W_rs . acts
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(full_w + 8 * ws, wph);
          tc_fence_after();
          const uint32_t first = cc == 0 ? 0u : 1u;
          const uint64_t bdesc = desc_hi | (uint64_t)(((acts0 + (uint32_t)cc * ACTS_CHUNK) & 0x3FFFFu) >> 4);
          const uint64_t adesc = desc_hi | (uint64_t)(((w0 + ws * W_STAGE) & 0x3FFFFu) >> 4);
          if (elect_one()) {
            if (!(p.debug & 16)) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma2<OPF>(tmem_base + 2 * PN, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma2<OPF>(tmem_base + 3 * PN, adesc + (CHUNK_BYTES >> 4) + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
            }
            tc2_commit(empty_w + 8 * ws);
          }
          __syncwarp();
          if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
        }
        if (elect_one()) tc2_commit(acc2_full);
        __syncwarp();
      };
      uint32_t it = 0;
      int tile = next_live(pair);
      if (tile < p.ntiles) gemm1();
      for (; tile < p.ntiles; ++it) {
        // activations of this tile written by all 16 epilogue warps (=> acc1 drained too), acc2 drained by the
        // previous tile's res_skip epilogue
        mbar_wait_cluster(acts_ready, it & 1u);
        mbar_wait(acc2_empty, (it & 1u) ^ 1u);
        tc_fence_after();
        gemm2();
        tile = next_live(tile + npairs);
        if (tile < p.ntiles) gemm1();
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;                 // which 64 of the tile's 128 frames
    const int lic = q * 32 + lane;                 // lane within this CTA's 128 channels / filter rows
    const uint32_t lead_acts_ready = map_to_cta(acts_ready, 0), lead_acc2_empty = map_to_cta(acc2_empty, 0);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    // gate channel of this lane and where its activations go: chunk cc of the B operand of GEMM 2, in the CTA that
    // owns frames [64 h, 64 h + 64) of the tile
    const int gch = (int)rank * CHUNK_M + lic;
    const bool gate_ok = gch < p.hid;
    const bool gate_warp_live = (int)rank * CHUNK_M + q * 32 < p.hid;
    const uint32_t acts_dst = map_to_cta(acts0, (uint32_t)h) + (uint32_t)(gch / KC) * ACTS_CHUNK;
    const uint32_t col_byte = (uint32_t)(gch % KC) * ESIZE;           // byte offset within the 128-byte row
    uint32_t it = 0;
    for (int tile = next_live(pair); tile < p.ntiles; tile = next_live(tile + npairs), ++it) {
      const int tb = tile % p.ntb, b = tile / p.ntb;
      const int t0 = tb * PN;
      const uint32_t par = it & 1u;
      // ---- gate epilogue: acc1 -> activations in shared memory (own or peer CTA) ----
      mbar_wait(acc1_full, par);
      tc_fence_after();
      if (gate_warp_live && !(p.debug & 1)) {
        const float* gb = p.gate_bias + (int64_t)b * p.gate_bias_bs;
        const float b_lo = gate_ok ? gb[gch] : 0.f, b_hi = gate_ok ? gb[p.hid + gch] : 0.f;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const int col = h * HN + blk * 32;
          float lo[32], hi[32];
          tmem_ld32(tlane + (uint32_t)col, lo);
          tmem_ld32(tlane + (uint32_t)(PN + col), hi);
          tmem_wait();
          if (gate_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float a = fast_gate(lo[i] + b_lo, hi[i] + b_hi);
              const uint32_t row = (uint32_t)(blk * 32 + i);                        // row within the owner's 64 frames
              const uint32_t off = row * ROW_BYTES + ((((col_byte >> 4) ^ (row & 7u)) << 4) | (col_byte & 15u));
              if constexpr (opf_is16(OPF)) {
                const auto v = to_operand<OPF>(a);
                st_cluster_b16(acts_dst + off, *reinterpret_cast<const uint16_t*>(&v));
              } else {
                st_cluster_b32(acts_dst + off, __float_as_uint(round_tf32(a)));
              }
            }
          }
        }
      }
      // generic-proxy writes (local and remote) -> visible to the tensor core's async-proxy reads; TMEM reads done
      asm volatile("fence.proxy.async;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (p.debug & 32) mbar_arrive_cluster_relaxed(lead_acts_ready);      // timing experiment only: NOT a correct publication
        else              mbar_arrive_cluster(lead_acts_ready);
      }

      // ---- res_skip epilogue: acc2 -> x += res, skip += skip (LINEAR epilogue of conv_tc on two accumulators) ----
      LinCtx kx[2];
      bool live[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int n_w = a * 2 * CHUNK_M + (int)rank * CHUNK_M + q * 32;
        live[a] = n_w < p.rs_cout;
        const int n = n_w + lane;
        kx[a].sg = &p.ep.seg[(p.ep.nseg > 1 && n_w >= p.ep.seg[1].col0) ? 1 : 0];
        const int c = n - kx[a].sg->col0;
        kx[a].ok = live[a] && n < p.rs_cout && c >= 0 && c < kx[a].sg->ncols;
        kx[a].all_ok = __all_sync(0xffffffffu, kx[a].ok);
        kx[a].c = kx[a].ok ? c : 0;
        kx[a].b = b;
        kx[a].lim = live_rows(p.ep, b);
        kx[a].bias = (p.ep.bias && kx[a].ok) ? p.ep.bias[(int64_t)b * p.ep.bias_bs + n] : 0.f;
      }
      const int col = h * HN;
      int nv = p.out_rows - (t0 + col);
      nv = nv < 0 ? 0 : (nv > HN ? HN : nv);
      float r[64];
      if (p.debug & 4) nv = 0;
      if (live[0] && nv > 0 && !(p.debug & 2)) lin_load<OPF>(kx[0], t0 + col, nv, r);           // in flight before the accumulator is complete
      mbar_wait(acc2_full, par);
      tc_fence_after();
      if (nv > 0) {
        if (live[0]) lin_finish<OPF>(kx[0], t0 + col, nv, r, tlane + (uint32_t)(2 * PN + col));
        if (live[1]) {
          if (!(p.debug & 2)) lin_load<OPF>(kx[1], t0 + col, nv, r);
          lin_finish<OPF>(kx[1], t0 + col, nv, r, tlane + (uint32_t)(3 * PN + col));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_acc2_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int OPF>
int launch_wn(const WnParams& p, int grid, size_t smem, cudaStream_t stream) {
  static std::atomic<bool> attr_done[MAX_DEVICES];
  if (first_use_on_device(attr_done))
    QVC_CHECK_CUDA(cudaFuncSetAttribute(conv_wn_kernel<OPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
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
  QVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_wn_kernel<OPF>, p));
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e1, stream));
  return post_launch("conv_wn_kernel");
}

}  // namespace

}  // namespace qvc

using namespace qvc;

// See include/qvc_b200.h.  Returns QVC_ERR_UNSUPPORTED (error string untouched) when the pair of layers is not a
// case this kernel handles; the caller then issues the two qvc_conv1d calls.
extern "C" int qvc_wn_layer(const qvc_conv_args* gi, const qvc_conv_args* rs, qvc_stream_t stream_) {
  using namespace qvc::tc;
  QVC_REQUIRE(gi && rs, "qvc_wn_layer: null args");
  cudaStream_t stream = (cudaStream_t)stream_;
  // Default: off whenever the frames-on-rows pair kernel is enabled (conv_tcr.cu, QVC_TC_ROWS): the layer then runs as its
  // two convolutions there, which is faster at every batch size measured (profiles/r02_summary.md).
  if (!tc_env_int("QVC_WN_FUSED", (tc_env_int("QVC_TC_ROWS", 11) & 1) ? 0 : 1)) return QVC_ERR_UNSUPPORTED;
  if (gi->backend != QVC_BACKEND_TCGEN05 || rs->backend != QVC_BACKEND_TCGEN05) return QVC_ERR_UNSUPPORTED;
  if (gi->opformat != rs->opformat || gi->opformat == QVC_OPF_F32) return QVC_ERR_UNSUPPORTED;
  if (gi->epilogue != QVC_EPI_GATE || rs->epilogue != QVC_EPI_LINEAR) return QVC_ERR_UNSUPPORTED;
  const int H = gi->cout / 2;
  const int esize = (int)opformat_bytes(gi->opformat);
  const int kc = ROW_BYTES / esize;
  if (gi->cin != H || rs->cin != H || H % kc || H <= CHUNK_M || H > 2 * CHUNK_M) return QVC_ERR_UNSUPPORTED;
  if (rs->k != 1 || gi->dil != 1 || gi->k < 1 || gi->k > 16) return QVC_ERR_UNSUPPORTED;
  if (rs->cout > 4 * CHUNK_M || gi->batch != rs->batch || gi->out_rows != rs->out_rows || gi->x_rows != gi->out_rows)
    return QVC_ERR_UNSUPPORTED;
  if (rs->nseg == 2 && (rs->seg[1].col0 % 32 || rs->seg[0].col0 % 32 || rs->seg[0].ncols % 32)) return QVC_ERR_UNSUPPORTED;
  if (gi->batch == 0 || gi->out_rows == 0) return QVC_OK;
  const int ntb = (gi->out_rows + PN - 1) / PN;
  const int ntiles = gi->batch * ntb;
  {
    // latency shapes (few tiles): QVC_WN_MIN_TILES overrides the threshold below which the layer runs as two launches
    // (measured, profiles/r02_summary.md: the fused layer on a single 5 s clip saves 32 launches but only ~50 us of 1.9 ms:
    // the dependent MMA chains, not the launch count, set the latency of a clip)
    const int min_tiles = tc_env_int("QVC_WN_MIN_TILES", tc_sm_count() / 4);
    if (ntiles < min_tiles && !tc_env_int("QVC_TC_2CTA_FORCE", 0)) return QVC_ERR_UNSUPPORTED;
  }
  QVC_REQUIRE(gi->bias != nullptr, "qvc_wn_layer: the gate needs a bias vector");
  EncodeTiledFn encode = tc_get_encode();
  if (!encode) return QVC_ERR_UNSUPPORTED;
  QVC_REQUIRE((gi->x.ld * esize) % 16 == 0 && ((uintptr_t)gi->x.ptr & 15) == 0 && ((uintptr_t)gi->w & 15) == 0 &&
                  ((uintptr_t)rs->w & 15) == 0,
              "qvc_wn_layer: x / w must be 16-byte aligned with 16-byte row pitch");
  // the fused kernel writes the new operand copy of x while other tiles still read the old one as halo
  for (int s = 0; s < rs->nseg; ++s)
    QVC_REQUIRE(rs->seg[s].op.ptr == nullptr || rs->seg[s].op.ptr != gi->x.ptr,
                "qvc_wn_layer: the operand output of res_skip must not alias the layer input (ping-pong it)");

  WnParams p{};
  QVC_PROPAGATE(build_epi_params(*rs, &p.ep));
  p.hid = H; p.k = gi->k; p.pad_left = gi->pad_left; p.rs_cout = rs->cout;
  p.ntb = ntb; p.ntiles = ntiles; p.out_rows = gi->out_rows; p.batch = gi->batch;
  p.gate_bias = gi->bias; p.gate_bias_bs = gi->bias_bstride;
  p.debug = tc_env_int("QVC_WN_DEBUG", 0);
  const int halo = gi->k - 1;
  p.slab_box_rows = (HN + halo + 7) & ~7;
  p.slab_stage_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
  const int n_cchunks = H / kc;
  static const int stage_options[][2] = {{3, 5}, {3, 4}, {2, 4}, {2, 3}};
  size_t smem = 0;
  bool fits = false;
  for (const auto& opt : stage_options) {
    smem = (size_t)opt[0] * p.slab_stage_bytes + (size_t)opt[1] * W_STAGE + (size_t)n_cchunks * ACTS_CHUNK + 1024 + 256 + live_cache_bytes(*rs);
    if (smem <= (size_t)MAX_SMEM) { p.slab_stages = opt[0]; p.w_stages = opt[1]; fits = true; break; }
  }
  if (!fits) return QVC_ERR_UNSUPPORTED;

  const CUtensorMapDataType dt = gi->opformat == QVC_OPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : (gi->opformat == QVC_OPF_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  {
    cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)gi->x_rows, (cuuint64_t)gi->batch};
    cuuint64_t strides[2] = {(cuuint64_t)gi->x.ld * esize,
                             (cuuint64_t)(gi->batch > 1 ? gi->x.bstride : (int64_t)gi->x_rows * gi->x.ld) * esize};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)p.slab_box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.mx, dt, 3, gi->x.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("qvc_wn_layer: cuTensorMapEncodeTiled(x) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)gi->k * H, (cuuint64_t)gi->cout};
    cuuint64_t strides[1] = {(cuuint64_t)gi->k * H * esize};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)CHUNK_M};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&p.mw_in, dt, 2, const_cast<void*>(gi->w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("qvc_wn_layer: cuTensorMapEncodeTiled(w_in) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)rs->cout};
    cuuint64_t strides[1] = {(cuuint64_t)H * esize};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)CHUNK_M};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&p.mw_rs, dt, 2, const_cast<void*>(rs->w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("qvc_wn_layer: cuTensorMapEncodeTiled(w_rs) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  int pairs = tc_sm_count() / 2;
  if (ntiles < pairs) pairs = ntiles;
  if (gi->opformat == QVC_OPF_BF16) return launch_wn<QVC_OPF_BF16>(p, 2 * pairs, smem, stream);
  if (gi->opformat == QVC_OPF_F16) return launch_wn<QVC_OPF_F16>(p, 2 * pairs, smem, stream);
  return launch_wn<QVC_OPF_TF32>(p, 2 * pairs, smem, stream);
}
