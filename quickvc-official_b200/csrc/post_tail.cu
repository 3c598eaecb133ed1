// post_tail.cu -- subband_conv_post and everything after it in ONE kernel (models.py:388-406):
//
//   conv_post (128 -> 72, k = 7) as a tcgen05 GEMM with frames on the accumulator rows (the mainloop of conv_tcr.cu),
//   whose epilogue IS the tail of tail.cu:  exp / pi*sin -> 16-point inverse real DFT in registers -> window ->
//   hop-4 overlap-add / envelope / trim (torch.istft semantics) -> x4 zero-stuffing folded into a 4-phase, 17-tap-per-band
//   synthesis FIR -> float4 stores of the waveform.
//
// The 72-channel post-net tensor (184 MB fp32 at B = 64 x 10 s, written by one launch and read back by the next) never
// exists in memory; the path reads the 128-channel operand series once and writes the waveform.
//
// Geometry.  A CTA owns TMR = 128 consecutive post-net frames = the 128 TMEM lanes of its accumulator, and produces the
// TQ = 4 * 121 sub-band samples (1936 waveform samples) whose hops start in frames [F0, F0 + 121): the other 7 frames are
// the halo (3 frames of overlap-add before, 2 + 2 frames of FIR reach), recomputed by the neighbouring CTA -- 5.8 % more
// MMA work instead of an exchange through memory.  A CTA pair (cta_group::2, M = 256) carries two consecutive such
// tiles; the 80 filter rows (72 live) are the N of the MMA, half of them loaded by each CTA.
//
// Epilogue (8 warps per CTA): each thread owns one frame and reads two of its four bands out of TMEM (the accumulator is
// released right there, so the MMAs of the next tile run under all of the following), does the polar / DFT / window work
// of tail.cu for them, and the three stages meet in shared memory exactly as in tail.cu.
#include <cstring>
#include <type_traits>

#include "conv_tc_common.cuh"

namespace qvc {

namespace {

using namespace tc;

constexpr int TMR = 128;                  // frames per CTA = TMEM lanes
constexpr int TR = TMR - 7;               // frames whose hop starts inside the CTA's tile
constexpr int TQ = 4 * TR;                // sub-band samples per CTA
constexpr int YW = TQ + 16;               // sub-band samples held per band (8 of FIR halo on each side)
constexpr int NCH = 72, NCOL = 80;        // live / padded post-net channels
constexpr int PT_BUF_COLS = 128;          // TMEM columns between the two accumulator sets
constexpr int N_EPI_THREADS = 32 * N_EPI_WARPS;

struct alignas(64) PostTailParams {
  CUtensorMap mx;                          // x as (channel, frame, utterance)
  CUtensorMap mw;                          // w as (channel, output channel, tap); box = one channel chunk x 40 rows x tg taps
  int32_t cin, k, pad_left, tg;
  int32_t nct;                             // CTA tiles per utterance
  int32_t ntb;                             // pair tiles per utterance
  int32_t ntiles;
  int32_t slab_box_rows, slab_stages, w_stages;
  uint32_t slab_stage_bytes, w_stage_bytes;
  int32_t batch, frames;                   // post-net frames the buffers are laid out for
  const float* bias;                       // [80]
  float* post_raw;                         // optional copy of the post-net output [b][frame][post_ld] (debug tap)
  int64_t post_bs;
  int32_t post_ld;
  float* wave;
  float* y_mb;
  const int32_t* live_units;
  int32_t frames_per_unit;
  float Ec[4 * 4 * 17];                    // synthesis filter and window as kernel parameters: FFMA operands from the constant bank
  float Wc[16];
};

__constant__ float pt_cos16[16] = {
    1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
    0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
    -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
    0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};

__device__ __forceinline__ void pt_tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void epi_bar_sync() {       // the 256 epilogue threads of this CTA (named barrier 1)
  asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_THREADS) : "memory");
}

template <int OPF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) post_tail_kernel(const __grid_constant__ PostTailParams p) {
  constexpr int ESIZE = opf_is16(OPF) ? 2 : 4;
  constexpr int KC = ROW_BYTES / ESIZE;
  constexpr uint32_t FMT = mma_format(OPF);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slab0 = smem_base;
  const uint32_t w0 = slab0 + p.slab_stages * p.slab_stage_bytes;
  const uint32_t bar0 = w0 + p.w_stages * p.w_stage_bytes;
  const uint32_t full_slab = bar0, empty_slab = full_slab + 8 * p.slab_stages;
  const uint32_t full_w = empty_slab + 8 * p.slab_stages, empty_w = full_w + 8 * p.w_stages;
  const uint32_t tmem_full = empty_w + 8 * p.w_stages, tmem_empty = tmem_full + 16;
  const uint32_t tmem_slot = tmem_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // epilogue work area: bias, squared window, windowed frame samples, trimmed sub-band signal
  float* bias_s = reinterpret_cast<float*>(tmem_slot_ptr + 4);
  float* W2 = bias_s + NCOL;
  float* inv_env = W2 + 16;                 // 1 / (sum of the four squared window values that overlap at position n mod 4)
  float (*Xf)[TMR][17] = reinterpret_cast<float (*)[TMR][17]>(inv_env + 4);
  float (*Y)[YW] = reinterpret_cast<float (*)[YW]>(&Xf[4][0][0]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * p.slab_stages + 2 * p.w_stages + 2; ++i) mbar_init(bar0 + 8 * i, 1);
    mbar_init(tmem_empty, 2 * N_EPI_WARPS);
    mbar_init(tmem_empty + 8, 2 * N_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * PT_BUF_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < NCOL; i += blockDim.x) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  if (threadIdx.x < 16) W2[threadIdx.x] = p.Wc[threadIdx.x] * p.Wc[threadIdx.x];
  if (threadIdx.x < 4) {
    const int i = threadIdx.x;
    inv_env[i] = 1.f / (((p.Wc[i] * p.Wc[i] + p.Wc[i + 4] * p.Wc[i + 4]) + p.Wc[i + 8] * p.Wc[i + 8]) + p.Wc[i + 12] * p.Wc[i + 12]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int n_cchunks = p.cin / KC;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mx) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mw) : "memory");
      const uint32_t lead_full_slab = map_to_cta(full_slab, 0), lead_full_w = map_to_cta(full_w, 0);
      const uint32_t slab_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
      const int k = p.k, tg = p.tg;
      uint32_t s = 0, ph = 0, ws = 0, wph = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs) {
        const int tb = tile % p.ntb, b = tile / p.ntb;
        const int f_first = TR * (2 * tb + (int)rank) - 3;           // frame on TMEM lane 0 of this CTA
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(empty_slab + 8 * s, ph ^ 1u);
          if (leader) mbar_expect_tx(full_slab + 8 * s, 2 * slab_bytes);
          tma2_load_3d(slab0 + s * p.slab_stage_bytes, &p.mx, lead_full_slab + 8 * s, cc * KC, f_first - p.pad_left, b);
          if (++s == (uint32_t)p.slab_stages) { s = 0; ph ^= 1u; }
          for (int j = 0; j < k; j += tg) {
            mbar_wait(empty_w + 8 * ws, wph ^ 1u);
            if (leader) mbar_expect_tx(full_w + 8 * ws, 2 * p.w_stage_bytes);
            tma2_load_3d(w0 + ws * p.w_stage_bytes, &p.mw, lead_full_w + 8 * ws, cc * KC, (int)rank * (NCOL / 2), j);
            if (++ws == (uint32_t)p.w_stages) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(NCOL >> 3) << 17) |
                             ((uint32_t)((2 * TMR) >> 4) << 24);
      const uint64_t desc_hi = smem_desc(0);
      const int k = p.k, tg = p.tg;
      const uint32_t w_tap_bytes = (uint32_t)(NCOL / 2) * ROW_BYTES;
      uint32_t s = 0, ph = 0, ws = 0, wph = 0, ait = 0;
      for (int tile = pair; tile < p.ntiles; tile += npairs) {
        const uint32_t buf = ait & 1u, bph = (ait >> 1) & 1u;
        mbar_wait(tmem_empty + 8 * buf, bph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * PT_BUF_COLS;
        for (int cc = 0; cc < n_cchunks; ++cc) {
          mbar_wait(full_slab + 8 * s, ph);
          const uint32_t slab = slab0 + s * p.slab_stage_bytes;
          for (int j0 = 0; j0 < k; j0 += tg) {
            mbar_wait(full_w + 8 * ws, wph);
            tc_fence_after();
            const int jn = k - j0 < tg ? k - j0 : tg;
            const uint64_t adesc0 = desc_hi | (uint64_t)(((slab + (uint32_t)j0 * ROW_BYTES) & 0x3FFFFu) >> 4);
            const uint64_t bdesc0 = desc_hi | (uint64_t)(((w0 + ws * p.w_stage_bytes) & 0x3FFFFu) >> 4);
            if (elect_one()) {
              for (int jj = 0; jj < jn; ++jj) {
                const uint32_t first = (cc | j0 | jj) == 0 ? 0u : 1u;
                const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)jj * (ROW_BYTES >> 4));
                const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)jj * (w_tap_bytes >> 4));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2<OPF>(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks == 0 ? first : 1u);
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
    // ===================== epilogue (warps 2..9 of both CTAs) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int h = (warp - 2) >> 2;                 // bands 2h, 2h + 1
    const int row = q * 32 + lane;                 // frame slot = TMEM lane
    const int te = (int)threadIdx.x - 64;          // 0 .. 255
    const uint32_t lead_tmem_empty = map_to_cta(tmem_empty, 0);
    const int FS = p.frames, nys = 4 * (FS - 1);
    uint32_t ait = 0;
    for (int tile = pair; tile < p.ntiles; tile += npairs, ++ait) {
      const int tb = tile % p.ntb, b = tile / p.ntb;
      const int ct = 2 * tb + (int)rank;           // this CTA's tile of the utterance (may lie past the last one)
      const int F0 = TR * ct, f_lo = F0 - 3, q0 = 4 * F0;
      const int F = p.live_units ? min(FS, p.live_units[b] * p.frames_per_unit + 1) : FS;   // this utterance's own frames
      const int ny = 4 * (F - 1);
      const uint32_t buf = ait & 1u, bph = (ait >> 1) & 1u;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * PT_BUF_COLS + (h ? 32u : 0u);
      // 1. this thread's two bands of its frame: 36 channels inside 48 aligned accumulator columns
      float acc[48];
      mbar_wait(tmem_full + 8 * buf, bph);
      tc_fence_after();
      pt_tmem_ld16(taddr, acc);
      pt_tmem_ld16(taddr + 16, acc + 16);
      pt_tmem_ld16(taddr + 32, acc + 32);
      tmem_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_tmem_empty + 8 * buf);      // the accumulator is free again
      const int f = f_lo + row;
      const bool frame_ok = ct < p.nct && f >= 0 && f < F;
      // (the column offset of a band inside `acc` must be a compile-time constant: a runtime one puts the array in local memory)
      auto bands = [&](auto off_c) {
        constexpr int off = decltype(off_c)::value;  // band 2 starts at channel 36 = column 32 + 4
#pragma unroll
        for (int sb = 0; sb < 2; ++sb) {
          const int s = 2 * h + sb;
          float in18[18];
#pragma unroll
          for (int i = 0; i < 18; ++i) in18[i] = acc[off + 18 * sb + i] + bias_s[18 * s + i];
          if (p.post_raw && row >= 3 && (row < 3 + TR || ct == p.nct - 1)) {       // debug tap: each frame written once (the last tile also owns the frames after its hops)
            float* dst = p.post_raw + (int64_t)b * p.post_bs + (int64_t)f * p.post_ld + 18 * s;
#pragma unroll
            for (int i = 0; i < 18; ++i) dst[i] = in18[i];
          }
          float re[9], im[9];
#pragma unroll
          for (int kk = 0; kk < 9; ++kk) {
            const float mag = __expf(in18[kk]);
            const float xp = in18[9 + kk];
            const float xr = fmaf(-6.28318530717958647692f, rintf(xp * 0.15915494309189533577f), xp);
            const float phs = 3.14159265358979323846f * __sinf(xr);               // in [-pi, pi]
            float sn, cs;
            __sincosf(phs, &sn, &cs);
            re[kk] = mag * cs;
            im[kk] = mag * sn;
          }
          // x[n] = A[n] - Bo[n], x[16-n] = A[n] + Bo[n]; Im of DC / Nyquist is ignored (irfft).  Even / odd bins:
          //   A[n] = Ae[n] + Ao[n], A[8-n] = Ae[n] - Ao[n];   Bo[n] = Be[n] + Bd[n], Bo[8-n] = Bd[n] - Be[n]
          // (cos(2 pi k (8-n) / 16) = (-1)^k cos(2 pi k n / 16), sin likewise with the opposite sign): n = 0..4 only.
          float xs[17];
#pragma unroll
          for (int n = 0; n <= 4; ++n) {
            float ae = re[0] + re[8], ao = 0.f, be = 0.f, bd = 0.f;
#pragma unroll
            for (int kk = 1; kk < 8; ++kk) {
              const int m = (kk * n) & 15;
              if (kk & 1) {
                ao = fmaf(2.f * re[kk], pt_cos16[m], ao);
                bd = fmaf(2.f * im[kk], pt_cos16[(m + 12) & 15], bd);
              } else {
                ae = fmaf(2.f * re[kk], pt_cos16[m], ae);
                be = fmaf(2.f * im[kk], pt_cos16[(m + 12) & 15], be);
              }
            }
            if (n & 1) ae -= 2.f * re[8];            // the Nyquist bin enters as (-1)^n re[8]
            const float a1 = 0.0625f * (ae + ao), b1 = 0.0625f * (be + bd);       // A[n], Bo[n]
            const float a2 = 0.0625f * (ae - ao), b2 = 0.0625f * (bd - be);       // A[8-n], Bo[8-n]
            xs[n] = a1 - b1;
            xs[16 - n] = a1 + b1;
            xs[8 - n] = a2 - b2;
            xs[8 + n] = a2 + b2;
          }
#pragma unroll
          for (int n = 0; n < 16; ++n) Xf[s][row][n] = xs[n] * p.Wc[n];
        }
      };
      if (frame_ok) {
        if (h == 0) bands(std::integral_constant<int, 0>{});
        else        bands(std::integral_constant<int, 4>{});
      }
      epi_bar_sync();

      // 2. overlap-add, envelope, trim: Y[s][i] = y[s][q0 - 8 + i].  Away from the utterance's edges four frames overlap at
      // every position and the envelope only depends on the position modulo the hop: a multiplication by one of four
      // reciprocals; the first and last three frames take the general path.
      for (int s = 0; s < 4; ++s) {
        for (int ii = te; ii < YW; ii += N_EPI_THREADS) {
          const int qq = q0 - 8 + ii;
          float v = 0.f;
          if (qq >= 0 && qq < ny) {
            const int pos = qq + 8;
            const int fhi = pos >> 2, n0 = pos & 3;
            if (fhi >= 3 && fhi < F) {
              const float* xr = &Xf[s][fhi - f_lo][n0];
              const float a = (xr[0] + xr[4 - 17]) + (xr[8 - 34] + xr[12 - 51]);
              v = a * inv_env[n0];
            } else {
              float a = 0.f, env = 0.f;
#pragma unroll
              for (int dd = 0; dd < 4; ++dd) {
                const int ff = fhi - dd;
                const int n = pos - 4 * ff;
                if (ff >= 0 && ff < F) {
                  a += Xf[s][ff - f_lo][n];
                  env += W2[n];
                }
              }
              v = a / env;
            }
          }
          Y[s][ii] = v;
        }
      }
      epi_bar_sync();

      // 3. polyphase synthesis: wave[4q + r] = sum_s sum_e E[s][r][e] * y[s][q + 8 - e]
      if (ct < p.nct) {
        for (int tt = te; tt < TQ; tt += N_EPI_THREADS) {
          const int qq = q0 + tt;
          if (qq >= nys) break;
          float o[4] = {0.f, 0.f, 0.f, 0.f};
          if (qq < ny) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
#pragma unroll
              for (int e = 0; e < 17; ++e) {
                const float yv = Y[s][tt + 16 - e];
#pragma unroll
                for (int r = 0; r < 4; ++r) o[r] = fmaf(p.Ec[(s * 4 + r) * 17 + e], yv, o[r]);
              }
            }
          }
          *reinterpret_cast<float4*>(p.wave + (int64_t)b * 4 * nys + 4 * (int64_t)qq) = make_float4(o[0], o[1], o[2], o[3]);
          if (p.y_mb) {
#pragma unroll
            for (int s = 0; s < 4; ++s) p.y_mb[((int64_t)b * 4 + s) * nys + qq] = Y[s][tt + 8];
          }
        }
      }
      // the next tile's Xf writes come after its own accumulator wait; its Y writes after the first barrier above
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // release / acquire: the peer's remote tmem_empty arrives have landed before this CTA's shared memory is reused
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * PT_BUF_COLS) : "memory");
  }
}

template <int OPF>
int launch_pt(const PostTailParams& p, int grid, size_t smem, cudaStream_t stream) {
  static std::atomic<bool> attr_done[MAX_DEVICES];
  if (first_use_on_device(attr_done))
    QVC_CHECK_CUDA(cudaFuncSetAttribute(post_tail_kernel<OPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
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
  QVC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, post_tail_kernel<OPF>, p));
  if (timed) QVC_CHECK_CUDA(cudaEventRecord(e1, stream));
  return post_launch("post_tail_kernel");
}

}  // namespace

}  // namespace qvc

using namespace qvc;

// See include/qvc_b200.h.
extern "C" int qvc_post_tail(const qvc_conv_args* post, const qvc_tail_weights* w, const int32_t* live_units,
                             int frames_per_unit, float* wave, float* y_mb, qvc_stream_t stream_) {
  using namespace qvc::tc;
  QVC_REQUIRE(post && w && wave, "qvc_post_tail: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const qvc_conv_args& a = *post;
  if (!tc_env_int("QVC_POST_TAIL", 1)) return QVC_ERR_UNSUPPORTED;
  if (a.backend != QVC_BACKEND_TCGEN05 || a.opformat == QVC_OPF_F32) return QVC_ERR_UNSUPPORTED;
  if (!w->synth_host || !w->window_host) return QVC_ERR_UNSUPPORTED;     // the coefficients travel as kernel parameters
  if (a.epilogue != QVC_EPI_LINEAR || a.cout != NCOL || a.dil != 1 || a.x_rows != a.out_rows || a.tap_split > 0) return QVC_ERR_UNSUPPORTED;
  const int esize = (int)opformat_bytes(a.opformat);
  const int kc = ROW_BYTES / esize;
  if (a.cin % kc || a.k < 1 || TMR + a.k - 1 > 256) return QVC_ERR_UNSUPPORTED;
  QVC_REQUIRE(a.x.ptr && a.w, "qvc_post_tail: null x / w");
  QVC_REQUIRE(!live_units || frames_per_unit >= 1, "qvc_post_tail: live_units needs frames_per_unit >= 1");
  QVC_REQUIRE(a.batch >= 0 && a.batch <= 65535 && a.out_rows >= 1, "qvc_post_tail: bad shape");
  const int frames = a.out_rows, nys = 4 * (frames - 1);
  if (a.batch == 0 || nys == 0) return QVC_OK;
  EncodeTiledFn encode = tc_get_encode();
  if (!encode) return QVC_ERR_UNSUPPORTED;
  QVC_REQUIRE((a.x.ld * esize) % 16 == 0 && ((uintptr_t)a.x.ptr & 15) == 0 && ((uintptr_t)a.w & 15) == 0 &&
                  ((uintptr_t)wave & 15) == 0,
              "qvc_post_tail: x / w / wave must be 16-byte aligned with 16-byte row pitch");
  QVC_REQUIRE(a.batch == 1 || (a.x.bstride * esize) % 16 == 0, "qvc_post_tail: utterance pitch not 16-byte aligned");

  PostTailParams p{};
  p.cin = a.cin; p.k = a.k; p.pad_left = a.pad_left;
  p.nct = (nys + TQ - 1) / TQ;
  p.ntb = (p.nct + 1) / 2;
  p.ntiles = a.batch * p.ntb;
  p.batch = a.batch; p.frames = frames;
  p.bias = a.bias;
  p.post_raw = reinterpret_cast<float*>(a.seg[0].raw.ptr); p.post_bs = a.seg[0].raw.bstride; p.post_ld = a.seg[0].raw.ld;
  p.wave = wave; p.y_mb = y_mb; p.live_units = live_units; p.frames_per_unit = frames_per_unit;
  memcpy(p.Ec, w->synth_host, sizeof(p.Ec));
  memcpy(p.Wc, w->window_host, sizeof(p.Wc));
  const uint32_t w_tap_bytes = (uint32_t)(NCOL / 2) * ROW_BYTES;            // 5 KB: 40 rows, five 8-row swizzle atoms
  {
    int tg = (int)(32768u / w_tap_bytes);
    tg = tg < 1 ? 1 : (tg > a.k ? a.k : tg);
    const int groups = (a.k + tg - 1) / tg;
    p.tg = (a.k + groups - 1) / groups;
  }
  p.w_stage_bytes = (uint32_t)p.tg * w_tap_bytes;
  p.slab_box_rows = (TMR + a.k - 1 + 7) & ~7;
  p.slab_stage_bytes = (uint32_t)p.slab_box_rows * ROW_BYTES;
  p.slab_stages = 4; p.w_stages = 3;
  const size_t epi_bytes = 4 * (size_t)(NCOL + 16 + 4 + 4 * TMR * 17 + 4 * YW);
  const size_t smem = (size_t)p.slab_stages * p.slab_stage_bytes + (size_t)p.w_stages * p.w_stage_bytes + 1024 + 256 + epi_bytes;
  if (smem > (size_t)MAX_SMEM) return QVC_ERR_UNSUPPORTED;

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
    if (r != CUDA_SUCCESS) { set_error("qvc_post_tail: cuTensorMapEncodeTiled(x) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.cin, (cuuint64_t)a.cout, (cuuint64_t)a.k};
    cuuint64_t strides[2] = {(cuuint64_t)a.k * a.cin * esize, (cuuint64_t)a.cin * esize};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)(NCOL / 2), (cuuint32_t)p.tg};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.mw, dt, 3, const_cast<void*>(a.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("qvc_post_tail: cuTensorMapEncodeTiled(w) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  int pairs = tc_sm_count() / 2;
  if (p.ntiles < pairs) pairs = p.ntiles;
  if (a.opformat == QVC_OPF_BF16) return launch_pt<QVC_OPF_BF16>(p, 2 * pairs, smem, stream);
  if (a.opformat == QVC_OPF_F16) return launch_pt<QVC_OPF_F16>(p, 2 * pairs, smem, stream);
  return launch_pt<QVC_OPF_TF32>(p, 2 * pairs, smem, stream);
}
