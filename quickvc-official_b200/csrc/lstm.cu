// lstm.cu -- SpeakerEncoder (models.py:507-546) as a persistent-RNN.
//
// Per layer: (1) the input projection W_ih x_t + b for every step at once (one exact-fp32 series
// GEMM), (2) a persistent recurrent kernel: a thread-block cluster of 8 CTAs keeps the layer's
// 1 MB W_hh resident on chip (32 hidden units x 4 gates per CTA, in registers) for all steps; each
// step every CTA computes its 32 units for up to 8 sequences and pushes the new h slice into all 8
// CTAs' shared memory with st.async, whose bytes complete a transaction mbarrier in the receiving CTA:
// a step is closed by data arrival, not by a cluster barrier (whose release fence also waited for the
// step's global stores and prefetches: ~1 us of every 2.7 us step, profiles/r01_final_summary.md).
// (3) a small finishing kernel: Linear, ReLU, L2-normalise, mean over the windows.
// All arithmetic is fp32 FMA: the recurrence amplifies operand rounding, and the whole encoder is
// 2.4 % of the path's FLOPs.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace qvc {

namespace {

constexpr int HID = 256;
constexpr int GATES = 4 * HID;
constexpr int CLUSTER = 8;
constexpr int UNITS_PER_CTA = HID / CLUSTER;   // 32
constexpr int MAXSEQ = 8;                      // sequences per cluster
constexpr int REC_THREADS = 256;
constexpr size_t REC_SMEM = (size_t)UNITS_PER_CTA * 4 * HID * sizeof(float)      // W_hh slice
                            + (size_t)2 * MAXSEQ * HID * sizeof(float)           // h double buffer
                            + 16;                                                // one mbarrier per h buffer

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 4 bytes into a CTA of the cluster; the write completes 4 bytes of the transaction count of `mbar` (same CTA)
__device__ __forceinline__ void st_async_f32(uint32_t cluster_addr, float v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(cluster_addr), "r"(__float_as_uint(v)), "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {          // bounded: a protocol bug must trap, not hang the GPU
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

struct RecParams {
  const float* w_hh;      // (1024, 256)
  const float* gx;        // [rows][1024] input projection (bias included)
  int32_t nseq, steps;
  int32_t spc;            // sequences per cluster (<= MAXSEQ): fewer = more clusters, shorter steps
  int32_t row_stride;     // first gx row of sequence s is s*row_stride ...
  int32_t last_row;       // ... except, when >= 0, the last sequence starts here
  float* hseq;            // [nseq][steps][256] or NULL
  float* hlast;           // [nseq][256] or NULL
};

// NS = sequences a cluster carries (compile time: the per-step loop has no branches, so the h loads of all k slices
// are scheduled ahead of the FMAs that use them -- with a runtime count every (k slice, sequence) block was a
// branch + LDS + dependent FMAs and the step was one long latency chain, ~4100 cycles for ~650 instructions per warp).
template <int NS>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(REC_THREADS, 1)
lstm_recurrent_kernel(const RecParams p) {
  extern __shared__ __align__(16) float smem[];
  float* Wsm = smem;                                        // [32 units][4 gates][256]
  float* hbuf = smem + UNITS_PER_CTA * 4 * HID;             // [2][MAXSEQ][256]
  const uint32_t mbar = smem_addr(hbuf + 2 * MAXSEQ * HID);  // [2]: h buffer b is complete when mbar[b]'s phase is

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int group = blockIdx.x / CLUSTER;                   // which block of <= NS sequences
  const int s_base = group * NS;
  const int ns = min(NS, p.nseq - s_base);
  const int tid = threadIdx.x;
  const int ks = tid & 7;                                   // k-slice of the matrix-vector product
  const int u = tid >> 3;                                   // local hidden unit
  const int U = rank * UNITS_PER_CTA + u;                   // global hidden unit

  // resident weights: rows (gate*256 + U) of W_hh
  for (int i = tid; i < UNITS_PER_CTA * 4 * (HID / 4); i += REC_THREADS) {
    const int k4 = i % (HID / 4);
    const int g = (i / (HID / 4)) & 3;
    const int uu = i / (HID);                               // (HID/4)*4 entries per unit
    const float4 v = *reinterpret_cast<const float4*>(
        p.w_hh + ((int64_t)(g * HID + rank * UNITS_PER_CTA + uu)) * HID + 4 * k4);
    *reinterpret_cast<float4*>(Wsm + ((uu * 4 + g) * HID) + 4 * k4) = v;
  }
  for (int i = tid; i < 2 * MAXSEQ * HID; i += REC_THREADS) hbuf[i] = 0.f;
  // every step, each of the 8 CTAs sends this CTA 32 units x ns sequences x 4 bytes
  const uint32_t step_bytes = (uint32_t)(CLUSTER * UNITS_PER_CTA * ns * sizeof(float));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // The 8 k-slice lanes of a unit reduce-scatter their partial sums (below): with NS sequences the lane whose low
  // log2(8 / NS) bits are zero ends up owning sequence ks / (8 / NS) and performs its cell update.
  constexpr int LANES_PER_SEQ = 8 / NS;
  const int s_own = ks / LANES_PER_SEQ;
  const bool own = (ks % LANES_PER_SEQ) == 0 && s_own < ns;
  int64_t gx_row0 = 0;
  if (own) {
    const int sg = s_base + s_own;
    gx_row0 = (p.last_row >= 0 && sg == p.nseq - 1) ? p.last_row : (int64_t)sg * p.row_stride;
  }
  float c_state = 0.f;
  float gxv[4] = {0.f, 0.f, 0.f, 0.f}, gxn[4] = {0.f, 0.f, 0.f, 0.f};     // this step's / the next step's input projection
  if (own) {
#pragma unroll
    for (int g = 0; g < 4; ++g) gxv[g] = __ldg(p.gx + gx_row0 * GATES + g * HID + U);
    if (p.steps > 1) {
#pragma unroll
      for (int g = 0; g < 4; ++g) gxn[g] = __ldg(p.gx + (gx_row0 + 1) * GATES + g * HID + U);
    }
  }
  cluster.sync();

  // this thread's slice of W_hh (unit u, 4 gates, 32 of the 256 k) stays in registers for all steps: the
  // per-step loop then reads only h from shared memory (the W reads were 4 of every 11 LDS.128)
  float4 wreg[4][HID / 32];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int kk = 0; kk < HID / 32; ++kk)
      wreg[g][kk] = *reinterpret_cast<const float4*>(Wsm + (u * 4 + g) * HID + kk * 32 + ks * 4);

  uint32_t remote_h[CLUSTER], remote_bar[CLUSTER];
#pragma unroll
  for (int r = 0; r < CLUSTER; ++r) {
    remote_h[r] = map_rank(smem_addr(hbuf), (uint32_t)r);
    remote_bar[r] = map_rank(mbar, (uint32_t)r);
  }

  for (int t = 0; t < p.steps; ++t) {
    // h_t is in buffer t & 1: zeros for t = 0, else complete once all 8 CTAs' slices of step t-1 have landed
    // (phase (t-1) >> 1 of that buffer's barrier).  Then arm the other buffer's barrier for h_{t+1}: nobody can
    // send h_{t+1} before this CTA has sent its part of it, below.
    if (t > 0) mbar_wait_parity(mbar + 8 * (t & 1), (uint32_t)(((t - 1) >> 1) & 1));
    if (tid == 0 && t + 1 < p.steps) mbar_expect(mbar + 8 * ((t + 1) & 1), step_bytes);
    const float* hc = hbuf + (t & 1) * MAXSEQ * HID;
    float acc[4][NS];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[g][s] = 0.f;

    // rows of sequences past ns are never written: they stay zero
#pragma unroll
    for (int kk = 0; kk < HID / 32; ++kk) {
      const int k = kk * 32 + ks * 4;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + s * HID + k);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          acc[g][s] = fmaf(wreg[g][kk].x, hv.x, acc[g][s]);
          acc[g][s] = fmaf(wreg[g][kk].y, hv.y, acc[g][s]);
          acc[g][s] = fmaf(wreg[g][kk].z, hv.z, acc[g][s]);
          acc[g][s] = fmaf(wreg[g][kk].w, hv.w, acc[g][s]);
        }
      }
    }
    // reduce(-scatter) over the 8 k-slice lanes, bit 2 then 1 then 0 of ks: while more than one sequence is left a
    // stage halves the sequences a lane keeps (the kept half chosen by that bit), afterwards it is a plain sum
    float gate[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) v[i] = acc[g][i];
      int n = NS;
#pragma unroll
      for (int bit = 4; bit >= 1; bit >>= 1) {
        if (n > 1) {
          const int half = n >> 1;
#pragma unroll
          for (int i = 0; i < NS / 2; ++i) {
            if (i < half) {
              const float keep = (ks & bit) ? v[half + i] : v[i];
              const float send = (ks & bit) ? v[i] : v[half + i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
          }
          n = half;
        } else {
          v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
        }
      }
      gate[g] = v[0];
    }
    if (own) {
      const float ig = sigmoid_acc(gate[0] + gxv[0]);
      const float fg = sigmoid_acc(gate[1] + gxv[1]);
      const float gg = tanhf(gate[2] + gxv[2]);
      const float og = sigmoid_acc(gate[3] + gxv[3]);
      c_state = fg * c_state + ig * gg;
      const float h_new = og * tanhf(c_state);
      if (t + 1 < p.steps) {
        const uint32_t off = (uint32_t)((((t + 1) & 1) * MAXSEQ * HID + s_own * HID + U) * sizeof(float));
        const uint32_t boff = 8u * (uint32_t)((t + 1) & 1);
#pragma unroll
        for (int r = 0; r < CLUSTER; ++r) st_async_f32(remote_h[r] + off, h_new, remote_bar[r] + boff);
      }
      if (p.hseq) p.hseq[((int64_t)(s_base + s_own) * p.steps + t) * HID + U] = h_new;
      if (p.hlast && t == p.steps - 1) p.hlast[(int64_t)(s_base + s_own) * HID + U] = h_new;
      // input projection two steps ahead: an L2 round trip is longer than one step
#pragma unroll
      for (int g = 0; g < 4; ++g) gxv[g] = gxn[g];
      if (t + 2 < p.steps) {
#pragma unroll
        for (int g = 0; g < 4; ++g) gxn[g] = __ldg(p.gx + (gx_row0 + t + 2) * GATES + g * HID + U);
      }
    }
  }
  cluster.sync();          // no CTA may exit while a peer could still be writing into its shared memory
}

// emb[w][n] = normalise(relu(lin_w h[w] + lin_b))   (models.py:517-518): one CTA per window
__global__ void __launch_bounds__(HID) spk_project_kernel(const float* hlast, const float* lin_w, const float* lin_b,
                                                          float* emb) {
  __shared__ float hs[HID];
  __shared__ float red[HID / 32];
  const int w = blockIdx.x, n = threadIdx.x;
  hs[n] = hlast[(int64_t)w * HID + n];
  __syncthreads();
  float a = lin_b[n];
  const float4* wr = reinterpret_cast<const float4*>(lin_w + (int64_t)n * HID);
#pragma unroll 4
  for (int k4 = 0; k4 < HID / 4; ++k4) {
    const float4 wv = __ldg(wr + k4);
    a = fmaf(wv.x, hs[4 * k4], a);
    a = fmaf(wv.y, hs[4 * k4 + 1], a);
    a = fmaf(wv.z, hs[4 * k4 + 2], a);
    a = fmaf(wv.w, hs[4 * k4 + 3], a);
  }
  a = fmaxf(a, 0.f);
  float sq = a * a;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((n & 31) == 0) red[n >> 5] = sq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < HID / 32; ++i) tot += red[i];
  emb[(int64_t)w * HID + n] = a / sqrtf(tot);
}

// g[e][n] = mean over the windows of embedding e, summed in window order (models.py:540): deterministic
__global__ void __launch_bounds__(HID) spk_mean_kernel(const float* emb, int seq_per_embed, float* g_out) {
  const int e = blockIdx.x, n = threadIdx.x;
  float mean = 0.f;
  for (int w = 0; w < seq_per_embed; ++w) mean += emb[((int64_t)e * seq_per_embed + w) * HID + n];
  g_out[(int64_t)e * HID + n] = mean / (float)seq_per_embed;
}

struct SpkPlan {
  int nseq, steps, n_embed, seq_per_embed, row_stride, last_row;
  size_t off_mel, off_gx, off_hseq, off_hlast, off_emb, total;
};

SpkPlan make_plan(int bm, int tm) {
  SpkPlan pl;
  if (tm > 128) {
    pl.nseq = (tm - 128 + 63) / 64 + 1;       // len(range(0, tm-128, 64)) + the last window
    pl.steps = 128;
    pl.n_embed = 1;
    pl.seq_per_embed = pl.nseq;
    pl.row_stride = 64;
    pl.last_row = tm - 128;
  } else {
    pl.nseq = bm;
    pl.steps = tm;
    pl.n_embed = bm;
    pl.seq_per_embed = 1;
    pl.row_stride = tm;
    pl.last_row = -1;
  }
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t o = 0;
  pl.off_mel = o;   o += up((size_t)bm * tm * 80 * 4);
  const size_t gx_rows = (size_t)pl.nseq * pl.steps > (size_t)bm * tm ? (size_t)pl.nseq * pl.steps : (size_t)bm * tm;
  pl.off_gx = o;    o += up(gx_rows * GATES * 4);
  pl.off_hseq = o;  o += up((size_t)pl.nseq * pl.steps * HID * 4);
  pl.off_hlast = o; o += up((size_t)pl.nseq * HID * 4);
  pl.off_emb = o;   o += up((size_t)pl.nseq * HID * 4);
  pl.total = o;
  return pl;
}

}  // namespace

int spk_sequences_per_cluster(int nseq) {
  int spc = nseq <= 4 ? 1 : 2;
  if (const char* e = getenv("QVC_SPK_SPC")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) spc = v; }
  return spc;
}

}  // namespace qvc

extern "C" size_t qvc_spk_workspace_bytes(int bm, int tm) {
  if (bm <= 0 || tm <= 0) return 0;
  return qvc::make_plan(bm, tm).total;
}

extern "C" int qvc_spk_embed(const qvc_spk_weights* w, const float* mel, int bm, int tm, float* g_out,
                             void* workspace, size_t workspace_bytes, qvc_stream_t stream_) {
  using namespace qvc;
  cudaStream_t stream = (cudaStream_t)stream_;
  QVC_REQUIRE(w && mel && g_out && workspace, "qvc_spk_embed: null pointer");
  QVC_REQUIRE(bm >= 1 && tm >= 1, "qvc_spk_embed: bad shape (%d, 80, %d)", bm, tm);
  // the reference stacks (W, Bm, 128, 80).squeeze(1): Bm > 1 is a 4-D LSTM input (models.py:536)
  QVC_REQUIRE(tm <= 128 || bm == 1, "qvc_spk_embed: mel longer than 128 frames must have batch 1 (got %d)", bm);
  const SpkPlan pl = make_plan(bm, tm);
  if (workspace_bytes < pl.total) {
    set_error("qvc_spk_embed: workspace %zu < %zu", workspace_bytes, pl.total);
    return QVC_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  float* mel_sm = reinterpret_cast<float*>(ws + pl.off_mel);
  float* gx = reinterpret_cast<float*>(ws + pl.off_gx);
  float* hseq = reinterpret_cast<float*>(ws + pl.off_hseq);
  float* hlast = reinterpret_cast<float*>(ws + pl.off_hlast);

  QVC_PROPAGATE(qvc_to_series_major(mel, mel_sm, bm, 80, tm, QVC_OPF_F32, stream_));

  static std::atomic<bool> attr_set[MAX_DEVICES];
  if (first_use_on_device(attr_set)) {
    QVC_CHECK_CUDA(cudaFuncSetAttribute(lstm_recurrent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REC_SMEM));
    QVC_CHECK_CUDA(cudaFuncSetAttribute(lstm_recurrent_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REC_SMEM));
    QVC_CHECK_CUDA(cudaFuncSetAttribute(lstm_recurrent_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REC_SMEM));
    QVC_CHECK_CUDA(cudaFuncSetAttribute(lstm_recurrent_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REC_SMEM));
  }

  for (int layer = 0; layer < 3; ++layer) {
    // input projection for every step: gx = W_ih x + (b_ih + b_hh)
    qvc_conv_args a{};
    const int rows = layer == 0 ? bm * tm : pl.nseq * pl.steps;
    a.x = qvc_tensor{layer == 0 ? (void*)mel_sm : (void*)hseq, 0, layer == 0 ? 80 : HID, 0};
    a.batch = 1; a.x_rows = rows; a.out_rows = rows;
    a.cin = layer == 0 ? 80 : HID;
    a.w = w->w_ih[layer]; a.bias = w->bias[layer]; a.bias_bstride = 0;
    a.cout = GATES; a.k = 1; a.dil = 1; a.pad_left = 0;
    a.epilogue = QVC_EPI_LINEAR; a.nseg = 1;
    a.seg[0].col0 = 0; a.seg[0].ncols = GATES; a.seg[0].alpha = 1.f; a.seg[0].beta = 1.f; a.seg[0].slope = 1.f;
    a.seg[0].raw = qvc_tensor{gx, 0, GATES, 0};
    a.opformat = QVC_OPF_F32; a.backend = QVC_BACKEND_FMA;
    QVC_PROPAGATE(launch_conv_fma(a, stream));

    RecParams rp;
    rp.w_hh = w->w_hh[layer];
    rp.gx = gx;
    rp.nseq = pl.nseq; rp.steps = pl.steps;
    rp.row_stride = layer == 0 ? pl.row_stride : pl.steps;
    rp.last_row = layer == 0 ? pl.last_row : -1;
    rp.hseq = layer < 2 ? hseq : nullptr;
    rp.hlast = layer == 2 ? hlast : nullptr;
    // Sequences (windows) are independent: spreading them over more clusters shortens every one of the 384
    // dependent steps (per-step FMA work is proportional to the sequences of a cluster).  spk_sequences_per_cluster():
    // one window per cluster up to 4 windows (a 5 s target mel = 3 windows = 24 SMs), two beyond (10 s = 7 windows =
    // 32 SMs), which keeps the encoder hidden behind the prior encoder it runs beside, for single clips too.
    const int spc = spk_sequences_per_cluster(pl.nseq);
    rp.spc = spc;
    const int groups = (pl.nseq + spc - 1) / spc;
    switch (spc) {
      case 1:  lstm_recurrent_kernel<1><<<groups * CLUSTER, REC_THREADS, REC_SMEM, stream>>>(rp); break;
      case 2:  lstm_recurrent_kernel<2><<<groups * CLUSTER, REC_THREADS, REC_SMEM, stream>>>(rp); break;
      case 4:  lstm_recurrent_kernel<4><<<groups * CLUSTER, REC_THREADS, REC_SMEM, stream>>>(rp); break;
      default: lstm_recurrent_kernel<8><<<groups * CLUSTER, REC_THREADS, REC_SMEM, stream>>>(rp); break;
    }
    QVC_PROPAGATE(post_launch("lstm_recurrent_kernel"));
  }
  float* emb = reinterpret_cast<float*>(ws + pl.off_emb);
  spk_project_kernel<<<pl.nseq, HID, 0, stream>>>(hlast, w->lin_w, w->lin_b, emb);
  QVC_PROPAGATE(post_launch("spk_project_kernel"));
  spk_mean_kernel<<<pl.n_embed, HID, 0, stream>>>(emb, pl.seq_per_embed, g_out);
  return post_launch("spk_mean_kernel");
}
