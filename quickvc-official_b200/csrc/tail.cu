// tail.cu -- the bandwidth-bound end of the decoder, one kernel:
//   subband_conv_post output  ->  exp / pi*sin  ->  16-point inverse real DFT in registers
//   ->  window  ->  hop-4 overlap-add / envelope / trim  (torch.istft semantics, models.py:399-401)
//   ->  x4 zero-stuffing folded into a 4-phase, 17-tap-per-band synthesis FIR (models.py:404-406)
//   ->  coalesced float4 stores of the waveform.
//
// A CTA owns one utterance and TQ = 256 sub-band samples (1024 waveform samples).  It stages the
// FT+7 = 71 post-net frames it needs through shared memory (3 frames of overlap-add halo and the
// +-8 sub-band samples of FIR halo on each side), so every HBM byte is read once per CTA and the
// only re-read is the 7/64 halo.
//
// Algorithmic bytes per post-net frame: 72*4 read + 16*4 written = 352 B.
#include <cstring>

#include "common.cuh"

namespace qvc {

namespace {

constexpr int TQ = 256;            // sub-band samples per CTA
constexpr int FT = TQ / 4;         // post-net frames whose hop starts inside the tile
constexpr int NFR = FT + 7;        // frames staged (3 before, 4 after incl. FIR halo)
constexpr int NTHREADS = 288;      // 72 frame slots x 4 bands
constexpr int YW = TQ + 16;        // sub-band samples held per band (8 halo each side)
constexpr int NCH = 72;

struct TailParams {
  const float* post;
  int32_t ld, frames;
  const float* window;
  const float* synth;   // [4][4][17]
  float* wave;
  float* y_mb;
  const int32_t* live_units;   // ragged batches: utterance b has live_units[b] * frames_per_unit + 1 frames (nullptr: all)
  int32_t frames_per_unit;
  // host copies, present when the caller supplied them: as kernel parameters they live in the constant bank,
  // so every coefficient is an FFMA operand instead of a shared-memory load
  float Ec[4 * 4 * 17];
  float Wc[16];
};

__constant__ float c_cos16[16] = {
    1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
    0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
    -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
    0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};

template <bool CONST_COEF>
__global__ void __launch_bounds__(NTHREADS) tail_kernel(const __grid_constant__ TailParams p) {
  __shared__ float Xf[4][NFR][17];                  // windowed frame samples (17: bank spread)
  __shared__ float Y[4][YW];                        // trimmed, normalised sub-band signal
  __shared__ float E[4 * 4 * 17];                   // synthesis filter
  __shared__ float W[16], W2[16];

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * TQ;                   // first sub-band sample of the tile
  const int fq = q0 >> 2;
  const int f_lo = fq - 3;                          // frame held in slot 0
  const int FS = p.frames;                          // frames / sub-band samples the buffers are laid out for
  const int nys = 4 * (FS - 1);
  const int F = p.live_units ? min(FS, p.live_units[b] * p.frames_per_unit + 1) : FS;   // this utterance's own
  const int ny = 4 * (F - 1);                       // sub-band samples per band

  if (tid < 16) {
    const float w = CONST_COEF ? p.Wc[tid] : p.window[tid];
    W[tid] = w;
    W2[tid] = w * w;
  }
  if (!CONST_COEF)
    for (int i = tid; i < 4 * 4 * 17; i += NTHREADS) E[i] = p.synth[i];

  const float* pb = p.post + (int64_t)b * FS * p.ld;
  __syncthreads();                                  // W / W2 / E visible

  // 2. one (frame, band) per thread: polar -> inverse real DFT -> window
  {
    const int slot = tid >> 2, s = tid & 3;
    const int f = f_lo + slot;
    if (slot < NFR && f >= 0 && f < F) {
      // the 18 inputs of this (frame, band) straight from global memory: 72 contiguous bytes, 8-byte aligned;
      // the four bands of a frame and consecutive frames are contiguous, so a warp reads one 2304-byte run.
      // (Staging whole frames through shared memory first cost 20 KB per CTA and a load phase nothing overlapped:
      // the kernel sat at 4 CTAs per SM, 43 % of its stall samples in that phase.)
      float in18[18];
      {
        const float2* src = reinterpret_cast<const float2*>(pb + (int64_t)f * p.ld + 18 * s);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float2 t2 = __ldg(src + k);
          in18[2 * k] = t2.x;
          in18[2 * k + 1] = t2.y;
        }
      }
      float re[9], im[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        // SFU paths (ex2 / sin / cos approx, abs error ~5e-7 after the 2*pi range reduction): the libm
        // versions made this kernel instruction-bound at 18 % of the HBM roofline
        const float mag = __expf(in18[k]);
        const float xp = in18[9 + k];
        const float xr = fmaf(-6.28318530717958647692f, rintf(xp * 0.15915494309189533577f), xp);
        const float ph = 3.14159265358979323846f * __sinf(xr);            // in [-pi, pi]
        float sn, cs;
        __sincosf(ph, &sn, &cs);
        re[k] = mag * cs;
        im[k] = mag * sn;
      }
      // x[n] = A[n] - Bo[n], x[16-n] = A[n] + Bo[n]; Im of DC / Nyquist is ignored (irfft)
#pragma unroll
      for (int n = 0; n <= 8; ++n) {
        float a = re[0] + ((n & 1) ? -re[8] : re[8]);
        float bo = 0.f;
#pragma unroll
        for (int k = 1; k < 8; ++k) {
          const int m = (k * n) & 15;
          a = fmaf(2.f * re[k], c_cos16[m], a);
          bo = fmaf(2.f * im[k], c_cos16[(m + 12) & 15], bo);   // sin(x) = cos(x - pi/2)
        }
        a *= 0.0625f;
        bo *= 0.0625f;
        Xf[s][slot][n] = (a - bo) * W[n];
        if (n >= 1 && n <= 7) Xf[s][slot][16 - n] = (a + bo) * W[16 - n];
      }
    }
  }
  __syncthreads();

  // 3. overlap-add, envelope, trim: Y[s][i] = y[s][q0 - 8 + i]
  for (int i = tid; i < 4 * YW; i += NTHREADS) {
    const int s = i / YW, ii = i % YW;
    const int q = q0 - 8 + ii;
    float v = 0.f;
    if (q >= 0 && q < ny) {
      const int pos = q + 8;                       // position in the untrimmed overlap-add
      const int fhi = pos >> 2;
      float acc = 0.f, env = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int f = fhi - d;
        const int n = pos - 4 * f;                 // 0..15
        if (f >= 0 && f < F) {
          acc += Xf[s][f - f_lo][n];
          env += W2[n];
        }
      }
      v = acc / env;
    }
    Y[s][ii] = v;
  }
  __syncthreads();

  // 4. polyphase synthesis: wave[4q + r] = sum_s sum_e E[s][r][e] * y[s][q + 8 - e]
  if (tid < TQ) {
    const int q = q0 + tid;
    if (q < nys) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (q < ny) {                                 // past the utterance's own end: zeros
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
          for (int e = 0; e < 17; ++e) {
            const float yv = Y[s][tid + 16 - e];
#pragma unroll
            for (int r = 0; r < 4; ++r) o[r] = fmaf(CONST_COEF ? p.Ec[(s * 4 + r) * 17 + e] : E[(s * 4 + r) * 17 + e], yv, o[r]);
          }
        }
      }
      *reinterpret_cast<float4*>(p.wave + (int64_t)b * 4 * nys + 4 * (int64_t)q) =
          make_float4(o[0], o[1], o[2], o[3]);
      if (p.y_mb) {
#pragma unroll
        for (int s = 0; s < 4; ++s) p.y_mb[((int64_t)b * 4 + s) * nys + q] = Y[s][tid + 8];
      }
    }
  }
}

}  // namespace

}  // namespace qvc

extern "C" int qvc_tail(const qvc_tail_weights* w, const float* post, int ld, int batch, int frames,
                        const int32_t* live_units, int frames_per_unit, float* wave, float* y_mb, qvc_stream_t stream) {
  using namespace qvc;
  QVC_REQUIRE(w && w->window && w->synth && post && wave, "qvc_tail: null pointer");
  QVC_REQUIRE(ld >= NCH && ld % 4 == 0 && ((uintptr_t)post & 15) == 0, "qvc_tail: post must be 16-byte aligned with ld %% 4 == 0");
  QVC_REQUIRE(batch >= 0 && frames >= 1, "qvc_tail: bad shape");
  QVC_REQUIRE(!live_units || frames_per_unit >= 1, "qvc_tail: live_units needs frames_per_unit >= 1");
  const int ny = 4 * (frames - 1);
  if (batch == 0 || ny == 0) return QVC_OK;
  QVC_REQUIRE(batch <= 65535, "qvc_tail: batch too large for one launch");
  TailParams p;                           // ~1.2 KB on the stack: launches from several host threads do not serialise
  p.post = post; p.ld = ld; p.frames = frames; p.window = w->window; p.synth = w->synth; p.wave = wave; p.y_mb = y_mb;
  p.live_units = live_units; p.frames_per_unit = frames_per_unit;
  dim3 grid((ny + TQ - 1) / TQ, batch);
  if (w->synth_host && w->window_host) {
    memcpy(p.Ec, w->synth_host, sizeof(p.Ec));
    memcpy(p.Wc, w->window_host, sizeof(p.Wc));
    tail_kernel<true><<<grid, NTHREADS, 0, (cudaStream_t)stream>>>(p);
  } else {
    tail_kernel<false><<<grid, NTHREADS, 0, (cudaStream_t)stream>>>(p);
  }
  return post_launch("tail_kernel");
}
