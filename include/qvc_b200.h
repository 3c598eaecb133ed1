/*
 * qvc_b200.h -- C ABI of libqvc_b200.so: the B200 (sm_100a) implementation of QuickVC's
 * conversion forward pass, `SynthesizerTrn.infer` (/root/reference/models.py:625-642).
 *
 * The reference has no native boundary at all (it is pure PyTorch); the only interface for this
 * path is the Python method `SynthesizerTrn.infer(unit, mel)`.  This header is the boundary a
 * binding for that method sits on: plain pointers, sizes and a cudaStream_t, no torch types.
 * The Python mirror of the reference class lives in quickvc-official_b200/models.py and calls
 * these entry points through ctypes; INTEGRATION.md shows the stub a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative qvc_status; qvc_last_error() returns a
 *     thread-local description of the last failure;
 *   - no function allocates device memory: the caller passes weights, workspace and outputs;
 *   - all work is enqueued on the caller's stream; functions are re-entrant across streams as long
 *     as each stream uses its own workspace;
 *   - activations inside the library are "series-major": [utterance][frame][channel], channel
 *     contiguous; the public tensors keep the reference layout (B, C, T).
 */
#ifndef QVC_B200_H_
#define QVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QVC_ABI_VERSION 6

typedef struct CUstream_st* qvc_stream_t;   /* == cudaStream_t */

typedef enum {
  QVC_OK = 0,
  QVC_ERR_ARG = -1,        /* bad argument (shape, alignment, null pointer)                  */
  QVC_ERR_CUDA = -2,       /* a CUDA runtime / driver call failed                            */
  QVC_ERR_NO_DEVICE = -3,  /* no sm_100 device / driver                                       */
  QVC_ERR_WORKSPACE = -4,  /* workspace too small                                             */
  QVC_ERR_UNSUPPORTED = -5 /* configuration outside what the kernels were built for           */
} qvc_status;

/* Operand format of the convolution GEMM operands (activations as stored between layers, and the
 * folded weights).  Accumulation is always fp32. */
typedef enum {
  QVC_OPF_F32 = 0,   /* unrounded fp32 operands; exact-fp32 FMA kernels only                  */
  QVC_OPF_TF32 = 1,  /* fp32 storage, values rounded-to-nearest to TF32 by the producer       */
  QVC_OPF_BF16 = 2,  /* bf16 storage                                                          */
  QVC_OPF_F16 = 3    /* IEEE half storage: the 10-bit mantissa of TF32 in 2 bytes, range +-65504 (the tensor
                        cores run it at the bf16 rate; activations and folded filters of this model stay far
                        inside the range, values below 6e-5 lose relative -- not absolute -- precision)    */
} qvc_opformat;

typedef enum {
  QVC_BACKEND_FMA = 0,     /* CUDA-core fp32 FMA kernels (exact fp32; validation / strict mode) */
  QVC_BACKEND_TCGEN05 = 1  /* tcgen05 tensor-core implicit-GEMM kernels fed by TMA              */
} qvc_backend;

/* ---------------------------------------------------------------------------------------------
 * Generic series convolution -- the building block behind every Conv1d / ConvTranspose1d on the
 * path (modules.py:64,67,134-143,193,195; models.py:71,73,327-335,346).
 *
 *   acc[b][t][n] = sum_{j<k} sum_{c<cin} w[n][j][c] * x[b][t + j*dil - pad_left][c]   (rows outside
 *                  [0, x_rows) read as zero -- the reference's zero "same" padding)
 *
 * followed by one of the fused epilogues below.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  void*   ptr;      /* base of [utterance][row][channel]                                      */
  int64_t bstride;  /* elements between utterances (0 = shared by all utterances)             */
  int32_t ld;       /* elements between rows                                                  */
  int32_t _pad;
} qvc_tensor;

typedef enum {
  QVC_EPI_LINEAR = 0, /* per column segment: v = alpha*(acc+bias) [+ res | res_op]; w = [accin +] beta*v;
                         raw <- w ; op <- round(leaky_relu(w, slope))                          */
  QVC_EPI_GATE = 1,   /* WN gate (modules.py:14-34): cout = 2H; op[n] <- round(tanh(a[n]) *
                         sigmoid(a[n+H])), a = acc + bias                                      */
  QVC_EPI_SAMPLE = 2  /* prior sample (models.py:93-94): cout = 2H; z = a[n] + noise*exp(a[n+H]);
                         raw <- z ; op <- round(z); optional m / logs copies                   */
} qvc_epilogue;

typedef struct {
  int32_t    col0, ncols;   /* output columns [col0, col0+ncols) of the GEMM feed this segment;
                               column n lands in channel n-col0 of the tensors below           */
  float      alpha, beta;   /* see QVC_EPI_LINEAR                                              */
  float      slope;         /* leaky-relu slope applied to the operand copy (1 = identity)     */
  int32_t    _pad;
  qvc_tensor res;           /* fp32, optional (ptr NULL = absent)                              */
  qvc_tensor res_op;        /* alternative to `res`: the residual given as an operand-format tensor that
                               holds leaky_relu(r, s); it is undone with r = v > 0 ? v : v * res_inv_slope
                               (res_inv_slope = 1/s).  Lets a residual stream live only as the operand
                               copy the next convolution reads anyway.                          */
  float      res_inv_slope;
  int32_t    _pad2;
  qvc_tensor accin;         /* fp32, optional                                                  */
  qvc_tensor raw;           /* fp32 out, optional                                              */
  qvc_tensor op;            /* operand-format out, optional                                    */
} qvc_epi_segment;

typedef struct {
  /* input series */
  qvc_tensor x;             /* operand format                                                  */
  int32_t    batch;         /* utterances                                                      */
  int32_t    x_rows;        /* rows (frames) per utterance in x                                */
  int32_t    out_rows;      /* rows produced per utterance                                     */
  int32_t    cin;           /* multiple of 16                                                  */
  /* filter */
  const void*  w;           /* [cout][k][cin], operand format                                  */
  const float* bias;        /* [cout] fp32, or per-utterance [batch][bias_bstride]; may be NULL */
  int64_t    bias_bstride;  /* 0 = shared                                                      */
  int32_t    cout;          /* GEMM columns (multiple of 16)                                   */
  int32_t    k, dil, pad_left;
  /* epilogue */
  int32_t    epilogue;      /* qvc_epilogue                                                    */
  int32_t    nseg;          /* QVC_EPI_LINEAR: 1 or 2 segments                                 */
  qvc_epi_segment seg[2];   /* GATE / SAMPLE use seg[0].raw / seg[0].op with H = cout/2 channels */
  qvc_tensor noise;         /* SAMPLE: fp32 [b][t][H]                                          */
  qvc_tensor aux0, aux1;    /* SAMPLE: optional fp32 copies of m and logs                      */
  int32_t    opformat;      /* qvc_opformat of x, w and every `op` output                      */
  int32_t    backend;       /* qvc_backend                                                     */
  /* Ragged batches (utterances of different lengths padded to out_rows): when live_units is non-NULL,
   * utterance b has live_units[b] * live_mul live output rows, and every `op` output row at or past that
   * count is written as ZERO -- exactly the zero "same" padding the next convolution would have seen at
   * the end of that utterance alone.  fp32 outputs (raw, aux) past the count are unspecified. */
  const int32_t* live_units; /* device, [batch], or NULL                                       */
  int32_t    live_mul;
  /* Structured zeros of the filter (a hint: results never depend on it).  When tap_split > 0, the block of w
   * with output columns in half p (p = n >= cout/2), input channels in half q (q = c >= tap_split) and tap j
   * is all zero unless tap_lo[p][q] <= j <= tap_hi[p][q]; kernels may skip such blocks.  Every range must be
   * non-empty.  Used by the frame-paired form of the 128-channel MRF layers (see qvc_model.paired). */
  int32_t    tap_split;
  int32_t    tap_lo[2][2], tap_hi[2][2];
} qvc_conv_args;

int qvc_conv1d(const qvc_conv_args* args, qvc_stream_t stream);

/* Sum of series convolutions with ONE fused epilogue (tcgen05 back end): the accumulator collects
 *     acc[b][t][n] = sum_{i < nsrc} conv(srcs[i]->x, srcs[i]->w)[b][t][n]        (each source with its own k, dil, pad_left, taps)
 * in source order inside tensor memory, then the QVC_EPI_LINEAR epilogue of srcs[0] (one segment, no accin) runs with
 *     bias = sum_i srcs[i]->bias,   residual = sum_i srcs[i]->seg[0].res (or .res_op, undone with that source's res_inv_slope):
 *     v = beta * (alpha * (acc + bias) + res_0 + res_1 + ...);  raw <- v;  op <- round(leaky_relu(v, slope)).
 * This is the mean of the three ResBlocks of an MRF stage (models.py:378-384) taken where their last convolutions
 * (modules.py:153-154) accumulate: xs / 3 = (1/3) sum_r (x_r + c2_r(t_r)) -- the fp32 running sum never exists in memory.
 * All sources share batch, x_rows == out_rows, cin, cout, opformat and backend; 1 <= nsrc <= QVC_MAX_SUM_SOURCES.
 * Returns QVC_ERR_UNSUPPORTED (error string untouched) for the FMA back end: the caller then chains qvc_conv1d calls
 * through `accin`. */
#define QVC_MAX_SUM_SOURCES 3
int qvc_conv1d_sum(const qvc_conv_args* const* srcs, int nsrc, qvc_stream_t stream);

/* One whole WN layer (modules.py:88-112) in one launch: `in_layer` (QVC_EPI_GATE) followed by `res_skip`
 * (QVC_EPI_LINEAR, k = 1) whose input is the gate output.  Equivalent to qvc_conv1d(in_layer) then
 * qvc_conv1d(res_skip) with res_skip->x == in_layer->seg[0].op, except that the gated activations stay in shared
 * memory (in_layer->seg[0] and res_skip->x are ignored) and that res_skip's operand output must not alias
 * in_layer->x (other tiles still read it as halo: ping-pong the operand copy of x between layers).
 * Returns QVC_ERR_UNSUPPORTED, leaving the error string alone, when the shapes are not ones the fused kernel
 * handles (it needs the tcgen05 back end, 128 < H <= 256 gate channels, enough tiles to fill the machine);
 * the caller then issues the two calls. */
int qvc_wn_layer(const qvc_conv_args* in_layer, const qvc_conv_args* res_skip, qvc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Layout / format helpers
 * ------------------------------------------------------------------------------------------- */
/* (B, C, T) fp32 -> [B][T][C] in `opformat` (rounded).  Replaces the implicit NCT layout of
 * `unit` / noise at models.py:625,94. */
int qvc_to_series_major(const float* src, void* dst, int batch, int channels, int frames,
                        int opformat, qvc_stream_t stream);
/* [B][T][ld>=C] fp32 -> (B, C, T) fp32 (debug taps). */
int qvc_from_series_major(const float* src, int ld, float* dst, int batch, int channels, int frames,
                          qvc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Speaker encoder: SpeakerEncoder.embed_utterance / forward (models.py:507-546).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* w_ih[3];   /* (1024, 80|256|256) row-major, gate rows i,f,g,o                  */
  const float* w_hh[3];   /* (1024, 256)                                                       */
  const float* bias[3];   /* (1024) = bias_ih + bias_hh, summed once at fold time              */
  const float* lin_w;     /* (256, 256)                                                        */
  const float* lin_b;     /* (256)                                                             */
} qvc_spk_weights;

/* mel (Bm, 80, Tm) fp32, reference layout.  Tm > 128 requires Bm == 1 and yields one embedding
 * (mean over the 128-frame windows, hop 64, plus the last window); Tm <= 128 yields Bm embeddings.
 * g_out :: [n_embed][256] fp32.  workspace: qvc_spk_workspace_bytes(). */
size_t qvc_spk_workspace_bytes(int bm, int tm);
int qvc_spk_embed(const qvc_spk_weights* w, const float* mel, int bm, int tm, float* g_out,
                  void* workspace, size_t workspace_bytes, qvc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Target-mel front end: wave_to_mel (mel_processing.py:15-98, called at convert.py:75-77) -- the step right
 * before the path; SURVEY.md section 8f "next" #1.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* basis;     /* [rows16][n_fft] fp32: row r < bins = hann[n] cos(2 pi r n / n_fft), row bins + r =
                             -hann[n] sin(...), zero rows up to rows16 = round_up(2 bins, 16); bins = n_fft/2+1 */
  const float* fbank_t;   /* [bins][n_mels] fp32: the mel filterbank, transposed                              */
  int32_t n_fft, hop, n_mels, _pad;
} qvc_mel_weights;

/* frames produced for `samples` input samples: 1 + (samples + (n_fft - hop) - n_fft) / hop */
int qvc_mel_frames(const qvc_mel_weights* w, int samples);
size_t qvc_mel_workspace_bytes(const qvc_mel_weights* w, int batch, int samples);
/* wave (B, samples) fp32 -> mel (B, n_mels, frames) fp32 = log(max(fbank . sqrt(|STFT|^2 + 1e-6), 1e-5)), with the
 * reference's reflect padding of (n_fft - hop)/2 samples on both sides and center=False. */
int qvc_wave_to_mel(const qvc_mel_weights* w, const float* wave, int batch, int samples, float* mel,
                    void* workspace, size_t workspace_bytes, qvc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Tail: magnitude exp / phase pi*sin, 16-point inverse real DFT, Hann window, hop-4 overlap-add,
 * envelope division and trim (torch.istft as called at models.py:350,399-401), then the x4
 * zero-stuffing and the learnable 4->1 synthesis filter (models.py:404-406), fused.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* window;    /* (16) dec.stft.window                                              */
  const float* synth;     /* [4 bands][4 phases][17 taps] folded from dec.updown_filter and
                             dec.multistream_conv_post (see fold.py)                           */
  const float* window_host; /* optional HOST copies of the two arrays above: when both are given */
  const float* synth_host;  /* the coefficients travel as kernel parameters (constant bank)      */
} qvc_tail_weights;

/* post :: [B][frames][ld] fp32 with 72 live channels (band*18 + {0..8 log-mag, 9..17 phase});
 * wave :: (B, 1, 16*(frames-1)) fp32; y_mb (optional) :: (B, 4, 4*(frames-1)) fp32.
 * live_units (optional, device [B]): utterance b only has live_units[b] * frames_per_unit + 1 post-net frames;
 * its waveform is what torch.istft and the synthesis filter give for that many frames, followed by zeros. */
int qvc_tail(const qvc_tail_weights* w, const float* post, int ld, int batch, int frames,
             const int32_t* live_units, int frames_per_unit, float* wave, float* y_mb, qvc_stream_t stream);

/* subband_conv_post AND the tail in one launch (tcgen05 back end): `post` describes the post-net convolution as for
 * qvc_conv1d (x = the reflection-padded 128-channel operand series, cout = 80 with 72 live rows, k = 7, QVC_EPI_LINEAR;
 * seg[0].raw, optional, receives a copy of the 72-channel post-net output), and the kernel's epilogue is qvc_tail: the
 * 72-channel tensor is never written and read back.  wave / y_mb / live_units as for qvc_tail with frames = post->out_rows.
 * Returns QVC_ERR_UNSUPPORTED (error string untouched) when the fused kernel does not apply (FMA back end, tail weights
 * without host copies, other shapes): the caller then issues qvc_conv1d + qvc_tail. */
int qvc_post_tail(const qvc_conv_args* post, const qvc_tail_weights* w, const int32_t* live_units, int frames_per_unit,
                  float* wave, float* y_mb, qvc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Whole path
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void*  w;          /* [cout][k][cin] operand format                                    */
  const float* bias;       /* [cout] fp32 or NULL                                              */
  int32_t cin, cout, k, dil, pad_left, _pad;
} qvc_layer;

/* Canonical layer order of qvc_model.layers (see qvc_layer_index_* below):
 *   0            enc_p.pre                      (models.py:71)
 *   1..16        enc_p.enc.in_layers.i          (modules.py:64)
 *   17..32       enc_p.enc.res_skip_layers.i    (modules.py:67)
 *   33           enc_p.proj                     (models.py:73)
 *   34+10c ..    coupling c = 0..3 in EXECUTION order (flows 6,4,2,0; modules.py:193-195):
 *                  +0 pre, +1..+4 in_layers, +5..+8 res_skip_layers, +9 post
 *   74           dec.conv_pre                   (models.py:327)
 *   75, 76       dec.ups.0, dec.ups.1 as polyphase series convolutions (models.py:333-335)
 *   77+6r ..     dec.resblocks.r, r = 0..5: +0..+2 convs1.j, +3..+5 convs2.j (modules.py:133-144)
 *   113          dec.subband_conv_post          (models.py:346)
 */
#define QVC_NUM_LAYERS 114

typedef struct {
  int32_t abi_version;      /* QVC_ABI_VERSION                                                  */
  int32_t opformat;         /* qvc_opformat of every layer's w                                  */
  int32_t backend;          /* qvc_backend                                                      */
  int32_t chunk_utts;       /* decoder sub-batch (utterances) kept L2-resident; 0 = auto        */
  qvc_layer layers[QVC_NUM_LAYERS];
  /* Frame-paired form of a layer (w == NULL: none), for dilation-1 layers with 128 input and output channels
   * (MRF-2, modules.py:133-144): two consecutive frames are one row of a series with 2 x 128 channels,
   *   out'[n][p*128 + c] = out[2n + p][c],   x'[m][q*128 + ci] = x[2m + q][ci]
   * -- the same memory, row pitch doubled -- which turns the k-tap 128 -> 128 convolution into a
   * (k+1)/2+1-tap 256 -> 256 one whose filter w'[p*128+c][a][q*128+ci] = w[c][2a + q - p + pad][ci] is zero outside
   * k + 1 of its (tap, input half) blocks.  256 output rows let a CTA pair issue M = 256 MMAs at the full tensor rate
   * where the 128-row form is capped at 2/3 by its shared-memory operand reads. */
  qvc_layer paired[QVC_NUM_LAYERS];
  /* Deferred skip sum of a WN stack (modules.py:106-112).  The stack's output  sum_i (W_skip_i . acts_i + b_skip_i)  is ONE
   * 1x1 convolution over the gated activations of its L layers laid side by side along the channel axis:
   *   wn_skip[s].w :: [192][1][L*192],  w[n][0][i*192 + c] = res_skip_i.w[(i < L-1 ? 192 : 0) + n][c],  bias[n] = sum_i b_skip_i[n]
   * (s = 0: enc_p.enc, L = 16;  s = 1 + c: coupling c in execution order, L = 4).  The engine then runs only the residual
   * half of every res_skip layer (rows [0, 192) of layers[...].w, none for the last layer) and the fp32 running skip sum
   * the reference reads and rewrites in every layer never exists in memory. */
  qvc_layer wn_skip[5];
  /* speaker conditioning folded to per-utterance bias vectors:
   * cond_w :: [cond_rows][256] fp32, cond_b :: [cond_rows] fp32 where the rows are
   * 4 couplings x (4 layers x 384) gate biases (cond_layer + in_layer bias, modules.py:83-96)
   * followed by 512 rows of dec.cond + dec.conv_pre bias (models.py:372). */
  const float* cond_w;
  const float* cond_b;
  int32_t cond_rows;        /* 4*1536 + 512 = 6656                                              */
  int32_t _pad;
  qvc_spk_weights spk;
  qvc_tail_weights tail;
} qvc_model;

/* ---------------------------------------------------------------------------------------------
 * Weight preparation: reference-layout state_dict -> qvc_model, once per load (replaces what the reference recomputes
 * on every forward: weight_norm at every weight_norm(...) site, Flip, cond_layer(g), ConvTranspose1d, the zero-stuffing
 * + synthesis filter -- see csrc/fold.cu).  Host arithmetic in double precision; the device receives one copy.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const char*  name;     /* key of the reference's state_dict (models.py:551-591; utils.py:161-176), e.g.
                            "dec.resblocks.3.convs1.0.weight_v"; enc_q.* entries may be present and are ignored */
  const float* data;     /* HOST pointer, fp32, contiguous, in the reference's (PyTorch) element order */
  int64_t      numel;
} qvc_state_entry;

/* bytes of the block qvc_prepare_weights / qvc_fold_host fill for `opformat` (0 = bad format) */
size_t qvc_prepared_bytes(int opformat);
/* Folds the entries into `device_block` (device memory of >= qvc_prepared_bytes(opformat), owned by the caller and kept
 * alive as long as `model` is used) and fills *model with pointers into it.  `tail_host` (optional): 16 + 272 host
 * floats, owned by the caller for the same lifetime, that receive the host copies of the tail coefficients
 * (qvc_tail_weights.window_host / synth_host).  The copy is enqueued on `stream`, which is synchronised before
 * returning.  A missing key or a wrong element count fails with QVC_ERR_ARG naming the key. */
int qvc_prepare_weights(const qvc_state_entry* entries, int n_entries, int opformat, int backend,
                        void* device_block, size_t device_bytes, float* tail_host, qvc_model* model,
                        qvc_stream_t stream);
/* The same fold into HOST memory (pointers in *model then refer to `block`): what the tests compare with the Python
 * statement of the fold (quickvc-official_b200/fold.py); no device needed. */
int qvc_fold_host(const qvc_state_entry* entries, int n_entries, int opformat, int backend, void* block,
                  size_t block_bytes, float* tail_host, qvc_model* model);

/* Optional per-stage copies in the reference layout (B, C, T) fp32; NULL = skip.
 * Names follow SURVEY.md section 8a. */
typedef struct {
  float* g;          /* (n_embed, 256)         */
  float* m_p;        /* (B, 192, T)            */
  float* logs_p;     /* (B, 192, T)            */
  float* z_p;        /* (B, 192, T)            */
  float* flow[4];    /* after couplings 6,4,2,0 (B, 192, T), reference channel order */
  float* conv_pre;   /* (B, 512, T)            */
  float* ups0;       /* (B, 256, 5T)           */
  float* mrf0;       /* (B, 256, 5T)           */
  float* ups1;       /* (B, 128, 20T)          */
  float* mrf1;       /* (B, 128, 20T)          */
  float* conv_post;  /* (B, 72, 20T+1)         */
  float* y_mb;       /* (B, 4, 80T)            */
} qvc_taps;

size_t qvc_infer_workspace_bytes(const qvc_model* model, int batch, int frames, int mel_batch,
                                 int mel_frames);

/* unit (B,256,T), mel (Bm,80,Tm), noise (B,192,T) [the torch.randn_like draw of models.py:94],
 * wave (B,1,320T); all fp32 device pointers in the reference layout.
 * g_in (optional, [n_embed][256]): a cached speaker embedding; when non-NULL the speaker encoder
 * is skipped and mel may be NULL.
 * lengths (optional, device int32 [B], 1 <= lengths[b] <= T): a ragged batch padded to T frames.  Utterance b
 * then gets, in wave[b][0 .. 320*lengths[b]), exactly the samples a call with that utterance alone
 * (batch 1, frames lengths[b]) produces, followed by zeros; the padding frames of unit / noise are ignored.
 * The reference has no such argument: it converts one utterance per call (convert.py:59-86). */
int qvc_infer(const qvc_model* model, const float* unit, const float* mel, const float* noise,
              const float* g_in, const int32_t* lengths, int batch, int frames, int mel_batch,
              int mel_frames, float* wave, const qvc_taps* taps, void* workspace,
              size_t workspace_bytes, qvc_stream_t stream);

/* Decoder only (BASELINE.json config 3): z (B,192,T) fp32, g [1|B][256] fp32; lengths as for qvc_infer. */
int qvc_decode(const qvc_model* model, const float* z, const float* g, int g_batch,
               const int32_t* lengths, int batch, int frames, float* wave, const qvc_taps* taps,
               void* workspace, size_t workspace_bytes, qvc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Misc
 * ------------------------------------------------------------------------------------------- */
const char* qvc_last_error(void);
int qvc_abi_version(void);
/* number of kernel launches enqueued by this process through the library (all streams). */
uint64_t qvc_launch_count(void);
/* name of the kernel the calling thread launched last through the library ("" before the first launch): lets a test
 * assert WHICH kernel served a call (e.g. "conv_tcr_kernel" for the frames-on-rows pair kernel). */
const char* qvc_last_kernel(void);
/* Page-locks / releases a host range the caller owns (cudaHostRegister, portable): the multi-GPU host gather registers one
 * buffer shared by the ranks of a box in each of them (quickvc-official_b200/shard.py).  A failure is reported here and
 * leaves no pending error behind in the caller's own CUDA runtime. */
int qvc_host_register(void* ptr, size_t bytes);
int qvc_host_unregister(void* ptr);
/* 0 when device `dev` is an sm_100 part this library has code for. */
int qvc_check_device(int dev);
/* Measurement aid (bench.py's roofline): while enabled, every tcgen05 series-convolution launch is
 * bracketed by CUDA events on its own stream.  qvc_profile(1) resets and starts, qvc_profile(0) stops;
 * qvc_profile_read synchronises the recorded events and returns their summed duration and count. */
int qvc_profile(int enable);
int qvc_profile_read(double* ms_total, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* QVC_B200_H_ */
