// mel.cu -- the target-speaker mel front end, wave_to_mel (mel_processing.py:79-98): the step that produces the
// `mel` argument of SynthesizerTrn.infer (convert.py:75-77; SURVEY.md section 8f, "next" #1).
//
//   reflect-pad (n_fft - hop)/2  ->  STFT(n_fft, hop, Hann(n_fft), center=False)  ->  sqrt(re^2 + im^2 + 1e-6)
//   ->  mel filterbank (Slaney scale and normalisation, librosa.filters.mel)  ->  log(max(., 1e-5))
//
// n_fft = 1280 is not a power of two and one call handles one target utterance (a few hundred frames), so the
// STFT is a plain GEMM: frames x windowed DFT basis.  The frames are never materialised: the padded waveform is
// handed to the exact-fp32 series convolution as a [frame][n_fft] tensor whose row pitch is the hop (overlapping
// rows).  fp32 throughout -- the result feeds the recurrent speaker encoder.
#include "common.cuh"

namespace qvc {

namespace {

// padded[b][i] = y[b][reflect(i - pad)]  (torch.nn.functional.pad(mode='reflect'), mel_processing.py:46)
__global__ void __launch_bounds__(256) reflect_pad_kernel(const float* __restrict__ y, float* __restrict__ out,
                                                          int samples, int pad, int padded, int64_t out_bs) {
  const int b = blockIdx.y;
  const float* src = y + (int64_t)b * samples;
  float* dst = out + (int64_t)b * out_bs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += gridDim.x * blockDim.x) {
    int j = i - pad;
    if (j < 0) j = -j;
    if (j >= samples) j = 2 * (samples - 1) - j;
    dst[i] = src[j];
  }
}

// one CTA per (frame, utterance): magnitudes into shared memory, then one thread per mel band
__global__ void __launch_bounds__(128) mel_log_kernel(const float* __restrict__ spec, int spec_ld, int64_t spec_bs,
                                                      const float* __restrict__ fbank_t /* [bins][mels] */,
                                                      int bins, int mels, int frames, float* __restrict__ mel) {
  extern __shared__ float mag[];
  const int f = blockIdx.x, b = blockIdx.y;
  const float* row = spec + (int64_t)b * spec_bs + (int64_t)f * spec_ld;
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    const float re = row[k], im = row[bins + k];
    mag[k] = sqrtf(re * re + im * im + 1e-6f);                 // mel_processing.py:54
  }
  __syncthreads();
  for (int m = threadIdx.x; m < mels; m += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < bins; ++k) acc = fmaf(fbank_t[(int64_t)k * mels + m], mag[k], acc);
    mel[((int64_t)b * mels + m) * frames + f] = logf(fmaxf(acc, 1e-5f));   // mel_processing.py:8,73-74
  }
}

struct MelPlan { int pad, padded, padded_ld, frames, rows16; size_t off_pad, off_spec, total; };

MelPlan mel_plan(const qvc_mel_weights* w, int batch, int samples) {
  MelPlan p{};
  p.pad = (w->n_fft - w->hop) / 2;
  p.padded = samples + 2 * p.pad;
  p.padded_ld = (p.padded + 3) & ~3;                            // 16-byte aligned utterance pitch
  p.frames = p.padded >= w->n_fft ? 1 + (p.padded - w->n_fft) / w->hop : 0;
  p.rows16 = (2 * (w->n_fft / 2 + 1) + 15) & ~15;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  p.off_pad = 0;
  p.off_spec = up((size_t)batch * p.padded_ld * 4 + 16);
  p.total = p.off_spec + up((size_t)batch * (p.frames > 0 ? p.frames : 1) * p.rows16 * 4);
  return p;
}

}  // namespace

}  // namespace qvc

using namespace qvc;

extern "C" int qvc_mel_frames(const qvc_mel_weights* w, int samples) {
  if (!w || samples <= 0 || w->hop <= 0 || w->n_fft <= 0) return 0;
  return mel_plan(w, 1, samples).frames;
}

extern "C" size_t qvc_mel_workspace_bytes(const qvc_mel_weights* w, int batch, int samples) {
  if (!w || batch <= 0 || samples <= 0) return 0;
  return mel_plan(w, batch, samples).total + 256;
}

extern "C" int qvc_wave_to_mel(const qvc_mel_weights* w, const float* wave, int batch, int samples, float* mel,
                               void* workspace, size_t workspace_bytes, qvc_stream_t stream_) {
  QVC_REQUIRE(w && wave && mel && workspace, "qvc_wave_to_mel: null pointer");
  QVC_REQUIRE(w->basis && w->fbank_t, "qvc_wave_to_mel: weights not populated");
  QVC_REQUIRE(w->n_fft > 0 && w->n_fft % 16 == 0 && w->hop > 0 && w->hop % 4 == 0 && w->hop <= w->n_fft && (w->n_fft - w->hop) % 2 == 0,
              "qvc_wave_to_mel: unsupported n_fft %d / hop %d", w->n_fft, w->hop);
  QVC_REQUIRE(w->n_mels >= 1 && w->n_mels <= 1024, "qvc_wave_to_mel: bad n_mels %d", w->n_mels);
  QVC_REQUIRE(batch >= 1 && batch <= 65535, "qvc_wave_to_mel: bad batch %d", batch);
  const MelPlan p = mel_plan(w, batch, samples);
  // reflect padding needs pad < samples (torch raises otherwise, mel_processing.py:46)
  QVC_REQUIRE(samples > p.pad, "qvc_wave_to_mel: %d samples are not more than the reflect padding %d", samples, p.pad);
  QVC_REQUIRE(p.frames >= 1, "qvc_wave_to_mel: waveform too short for one frame");
  const uintptr_t mis = (uintptr_t)workspace & 255;
  char* ws = reinterpret_cast<char*>(workspace) + (mis ? 256 - mis : 0);
  if (workspace_bytes < p.total + (mis ? 256 - mis : 0)) {
    set_error("qvc_wave_to_mel: workspace %zu < %zu", workspace_bytes, p.total + 256);
    return QVC_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  float* padded = reinterpret_cast<float*>(ws + p.off_pad);
  float* spec = reinterpret_cast<float*>(ws + p.off_spec);
  const int bins = w->n_fft / 2 + 1;

  {
    dim3 grid((p.padded + 255) / 256 < 1024 ? (p.padded + 255) / 256 : 1024, batch);
    reflect_pad_kernel<<<grid, 256, 0, stream>>>(wave, padded, samples, p.pad, p.padded, p.padded_ld);
    QVC_PROPAGATE(post_launch("reflect_pad_kernel"));
  }
  {
    // STFT = [frames][n_fft] (row pitch = hop: overlapping rows of the padded waveform) x basis^T, exact fp32
    qvc_conv_args a{};
    a.x = qvc_tensor{padded, (int64_t)p.padded_ld, w->hop, 0};
    a.batch = batch; a.x_rows = p.frames; a.out_rows = p.frames; a.cin = w->n_fft;
    a.w = w->basis; a.bias = nullptr; a.cout = p.rows16; a.k = 1; a.dil = 1; a.pad_left = 0;
    a.epilogue = QVC_EPI_LINEAR; a.nseg = 1;
    a.seg[0].col0 = 0; a.seg[0].ncols = p.rows16; a.seg[0].alpha = 1.f; a.seg[0].beta = 1.f; a.seg[0].slope = 1.f;
    a.seg[0].raw = qvc_tensor{spec, (int64_t)p.frames * p.rows16, p.rows16, 0};
    a.opformat = QVC_OPF_F32; a.backend = QVC_BACKEND_FMA;
    QVC_PROPAGATE(qvc_conv1d(&a, stream_));
  }
  {
    dim3 grid(p.frames, batch);
    mel_log_kernel<<<grid, 128, (size_t)bins * sizeof(float), stream>>>(spec, p.rows16, (int64_t)p.frames * p.rows16, w->fbank_t,
                                                                       bins, w->n_mels, p.frames, mel);
    QVC_PROPAGATE(post_launch("mel_log_kernel"));
  }
  return QVC_OK;
}
