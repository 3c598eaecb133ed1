// engine.cu -- enqueues SynthesizerTrn.infer (models.py:625-642) as a fixed sequence of series
// convolutions with fused epilogues, the persistent-RNN speaker encoder and the fused tail.
//
// Data layout in HBM (all "series-major" [utterance][frame][channel]):
//   *R buffers: fp32 residual / skip streams (never rounded)
//   *O buffers: GEMM operands in the model's operand format (TF32-rounded fp32, or bf16), with the
//               next layer's leaky-relu already applied by the producing epilogue
// Whole-batch buffers serve the prior encoder and the flow (192 channels at the unit frame rate);
// the decoder runs per sub-batch of `chunk_utts` utterances (default: the whole batch, bounded only by
// the workspace size).
#include <cstdlib>

#include "common.cuh"
#include "engine.h"

namespace qvc {

namespace {

constexpr int HID = 192;        // inter / hidden channels (configs/quickvc.json:41-42)
constexpr int UNIT_CH = 256;    // models.py:579
constexpr int GIN = 256;
constexpr int C_PRE = 512, C_UP0 = 256, C_UP1 = 128, C_POST = 72;
constexpr int UP0 = 5, UP1 = 4;

// canonical layer order (qvc_b200.h)
constexpr int L_ENC_PRE = 0, L_ENC_IN = 1, L_ENC_RS = 17, L_ENC_PROJ = 33, L_FLOW = 34, L_DEC_PRE = 74,
              L_UPS = 75, L_RES = 77, L_POST = 113;

struct Bump {
  char* base; size_t off = 0, cap;
  Bump(void* b, size_t c) : base(reinterpret_cast<char*>(b)), cap(c) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  }
};

struct Shapes { int B, T, n_embed; int cb; };

struct Buffers {
  // whole batch
  void *unitO, *xO, *xO2, *actsO, *skipO, *zO;      // actsO: gated activations of every layer of a WN stack side by side ([frame][L * 192])
  float *noiseT, *xR, *skipR, *zR, *tapA, *tapB, *condvec, *g;
  void* spk_ws; size_t spk_ws_bytes;
  // per decoder sub-batch; [3]: one per ResBlock of an MRF stage (the three blocks' last convolutions run as ONE sum
  // of convolutions, so their inputs and residual streams are alive together)
  void *aO, *x1O, *t1O[3], *xaO[3], *uO, *y1O, *t2O[3], *yaO[3], *pO;
  float *x1R, *xaR[3], *sum1R, *y1R, *yaR[3], *sum2R, *cpR;
};

// Utterances per chunk of the WN stacks (prior encoder, flow).  Default: the whole batch.  QVC_WN_CHUNK=n runs them n
// utterances at a time so that the fp32 residual / skip streams of a chunk stay L2-resident between layers -- measured
// on B200 (profiles/r02_summary.md): at B = 64 x 10 s chunks of 32 / 16 / 8 utterances cost +3 % / +9 % / +45 % of the
// step in tf32 (+2.5 / +12 / +54 % in fp16): a fused WN layer is bound by its serial per-tile chain and by streaming its
// filters from L2, not by HBM, and short launches lose the overlap between consecutive tiles.
int wn_chunk_utts(int B, int T) {
  static const int env = [] {
    const char* e = getenv("QVC_WN_CHUNK");
    return e ? atoi(e) : 0;
  }();
  (void)T;
  return (env <= 0 || env > B) ? B : env;
}

// Deferred skip sum (qvc_model.wn_skip, tcgen05 back end): every layer of a WN stack writes its gated activations into its
// own 192-channel column block of one [frame][L * 192] tensor, runs only the residual half of res_skip, and ONE 1x1
// convolution over all L * 192 channels produces the stack's output.  The reference's running fp32 skip sum (read and
// rewritten by every layer, modules.py:108-112) is never stored: -40 % of the traffic of the memory-bound res_skip layers.
// QVC_WN_DEFER=0 keeps the layer-by-layer form (the FMA back end always does).
constexpr int WN_MAX_LAYERS = 16;
bool wn_defer_skip(const qvc_model* m) {
  static const int env = [] {
    const char* e = getenv("QVC_WN_DEFER");
    return e ? atoi(e) : 1;
  }();
  return env != 0 && m->backend == QVC_BACKEND_TCGEN05 && m->wn_skip[0].w != nullptr;
}

int pick_chunk(const qvc_model* m, int B, int T) {
  int cb = m->chunk_utts;
  if (cb <= 0) {
    // Whole batch per launch unless the decoder working set would pass ~24 GB (measured on B200,
    // profiles/r01_v1_summary.md: sub-batching for L2 residency starves the grid and is 2.2x slower).
    // Per utterance: ~24 live series of 20 T rows x 128 channels (or 5 T x 256), 4 bytes each.
    const double per_utt = 24.0 * 20.0 * T * C_UP1 * 4.0;
    cb = (int)(24.0 * 1024 * 1024 * 1024 / (per_utt > 1 ? per_utt : 1));
    if (cb < 1) cb = 1;
  }
  return cb > B ? B : cb;
}

// The 16-bit modes of the tensor-core back end keep the residual streams of the MRF only as the operand copy the next
// convolution reads anyway (qvc_epi_segment.res_op): the c2 layers are memory-bound there and this halves their
// traffic.  In the fp32 (TF32) mode the residual streams stay unrounded fp32.
int residual_from_operand_mode() {                // 0: never, 1 (default): 16-bit modes, 2: experiment -- TF32 mode too
  static const int mode = [] {
    const char* e = getenv("QVC_RES_FROM_OP");
    return e ? atoi(e) : 1;
  }();
  return mode;
}
bool residual_from_operand(const qvc_model* m) {
  const int mode = residual_from_operand_mode();
  if (m->backend != QVC_BACKEND_TCGEN05) return false;
  return (mode >= 1 && opf_is16(m->opformat)) || (mode >= 2 && m->opformat == QVC_OPF_TF32);
}

size_t carve(const qvc_model* m, const Shapes& s, int mel_batch, int mel_frames, bool with_spk,
             void* ws, size_t cap, Buffers* b) {
  Bump a(ws, cap);
  const size_t E = opformat_bytes(m->opformat);
  const size_t n1 = (size_t)s.B * s.T;
  b->unitO = a.take(n1 * UNIT_CH * E);
  b->noiseT = (float*)a.take(n1 * HID * 4);
  b->xR = (float*)a.take(n1 * HID * 4);      b->xO = a.take(n1 * HID * E);
  b->xO2 = a.take(n1 * HID * E);             // ping-pong partner of xO (fused WN layers)
  b->actsO = a.take(n1 * (wn_defer_skip(m) ? WN_MAX_LAYERS : 1) * HID * E);
  b->skipR = (float*)a.take(n1 * HID * 4);   b->skipO = a.take(n1 * HID * E);
  b->zR = (float*)a.take(n1 * HID * 4);      b->zO = a.take(n1 * HID * E);
  b->tapA = (float*)a.take(n1 * HID * 4);    b->tapB = (float*)a.take(n1 * HID * 4);
  b->condvec = (float*)a.take((size_t)s.n_embed * m->cond_rows * 4);
  b->g = (float*)a.take((size_t)s.n_embed * GIN * 4);
  b->spk_ws_bytes = with_spk ? qvc_spk_workspace_bytes(mel_batch, mel_frames) : 0;
  b->spk_ws = a.take(b->spk_ws_bytes);
  const size_t c = (size_t)s.cb, T = (size_t)s.T;
  const size_t r0 = c * UP0 * T, r1 = c * UP0 * UP1 * T, rp = c * (UP0 * UP1 * T + 1);
  b->aO = a.take(c * T * C_PRE * E);
  // the fp32 residual streams exist only where the mode keeps them (the 16-bit tensor-core modes read the residual back
  // from the operand copy, residual_from_operand); the fp32 running sum only on the FMA back end
  const bool tc = m->backend == QVC_BACKEND_TCGEN05;
  const bool need_r = !residual_from_operand(m);
  b->x1R = (float*)a.take(r0 * C_UP0 * 4);   b->x1O = a.take(r0 * C_UP0 * E);
  for (int r = 0; r < 3; ++r) {
    b->t1O[r] = a.take(r0 * C_UP0 * E);
    b->xaR[r] = (float*)a.take(need_r || r == 0 ? r0 * C_UP0 * 4 : 0);
    b->xaO[r] = a.take(r0 * C_UP0 * E);
  }
  b->sum1R = (float*)a.take(tc ? 0 : r0 * C_UP0 * 4); b->uO = a.take(r0 * C_UP0 * E);
  b->y1R = (float*)a.take(r1 * C_UP1 * 4);   b->y1O = a.take(r1 * C_UP1 * E);
  for (int r = 0; r < 3; ++r) {
    b->t2O[r] = a.take(r1 * C_UP1 * E);
    b->yaR[r] = (float*)a.take(need_r || r == 0 ? r1 * C_UP1 * 4 : 0);
    b->yaO[r] = a.take(r1 * C_UP1 * E);
  }
  b->sum2R = (float*)a.take(tc ? 0 : r1 * C_UP1 * 4);
  b->pO = a.take(rp * C_UP1 * E);
  b->cpR = (float*)a.take(rp * C_POST * 4);
  return a.off;
}

struct Ctx {
  const qvc_model* m;
  cudaStream_t st;
  size_t E;
  int T;                         // unit frames the batch is padded to
  const int32_t* lengths;        // device [batch] live unit frames per utterance, or nullptr (ragged batches)
};

inline qvc_tensor tens(const void* p, int64_t bs, int ld) { return qvc_tensor{const_cast<void*>(p), bs, ld, 0}; }
inline qvc_tensor none() { return qvc_tensor{nullptr, 0, 0, 0}; }

inline qvc_epi_segment seg(int col0, int ncols) {
  qvc_epi_segment s{};
  s.col0 = col0; s.ncols = ncols; s.alpha = 1.f; s.beta = 1.f; s.slope = 1.f;
  return s;
}

qvc_conv_args layer_args(const Ctx& c, const qvc_layer& L, qvc_tensor x, int batch, int x_rows, int out_rows);
qvc_conv_args layer_args(const Ctx& c, int li, qvc_tensor x, int batch, int x_rows, int out_rows) {
  return layer_args(c, c.m->layers[li], x, batch, x_rows, out_rows);
}
qvc_conv_args layer_args(const Ctx& c, const qvc_layer& L, qvc_tensor x, int batch, int x_rows, int out_rows) {
  qvc_conv_args a{};
  a.x = x; a.batch = batch; a.x_rows = x_rows; a.out_rows = out_rows;
  a.cin = L.cin; a.w = L.w; a.bias = L.bias; a.bias_bstride = 0;
  a.cout = L.cout; a.k = L.k; a.dil = L.dil; a.pad_left = L.pad_left;
  a.epilogue = QVC_EPI_LINEAR; a.nseg = 1;
  a.opformat = c.m->opformat; a.backend = c.m->backend;
  // ragged batches: operand rows past an utterance's own length are written as zero, so that every later
  // convolution sees the zero padding it would see at the end of that utterance alone
  if (c.lengths && out_rows % c.T == 0) { a.live_units = c.lengths; a.live_mul = out_rows / c.T; }
  return a;
}

int run(const Ctx& c, const qvc_conv_args& a) { return qvc_conv1d(&a, (qvc_stream_t)c.st); }

// Frame-paired form of a dilation-1, 128-channel layer (qvc_model.paired): used whenever it exists, whatever the
// batch, so that an utterance's samples do not depend on what it is batched with (the summation order inside a
// dot product differs between the two forms).
bool use_frame_pairs(const Ctx& c, int li, int rows) {
  const char* e = getenv("QVC_FRAME_PAIR");          // read per call: tests switch it
  const bool enabled = !(e && e[0] == '0');
  return enabled && c.m->backend == QVC_BACKEND_TCGEN05 && c.m->paired[li].w != nullptr && rows % 2 == 0;
}

bool rows_kernel_serves_mrf2() {
  const char* e = getenv("QVC_TC_ROWS");               // read per call: tests switch it (default mask 11, conv_tcr.cu)
  return ((e ? atoi(e) : 11) & (8 | 4)) != 0;
}

// arguments of layer li in frame-paired form: x holds `rows` frames of `ch` channels with row pitch ch
qvc_conv_args paired_args(const Ctx& c, int li, const void* x, int64_t bs, int batch, int rows, int ch) {
  const qvc_layer& L = c.m->paired[li];
  const qvc_layer& L0 = c.m->layers[li];
  qvc_conv_args a{};
  a.x = tens(x, bs, 2 * ch); a.batch = batch; a.x_rows = rows / 2; a.out_rows = rows / 2;
  a.cin = L.cin; a.w = L.w; a.bias = L.bias; a.bias_bstride = 0;
  a.cout = L.cout; a.k = L.k; a.dil = 1; a.pad_left = L.pad_left;
  a.epilogue = QVC_EPI_LINEAR; a.nseg = 1;
  a.opformat = c.m->opformat; a.backend = c.m->backend;
  if (c.lengths && (rows / 2) % c.T == 0) { a.live_units = c.lengths; a.live_mul = (rows / 2) / c.T; }
  // w'[p C + c][a][q cin + ci] = w[c][2 (a - pad') + q - p + pad][ci]: non-zero for 0 <= 2 (a - pad') + q - p + pad < k
  a.tap_split = ch;
  for (int p = 0; p < 2; ++p)
    for (int q = 0; q < 2; ++q) {
      const int off = p - q - L0.pad_left;                       // need 0 <= 2 a0 - off < k with a0 = a - pad'
      const int lo = off >= 0 ? (off + 1) / 2 : -((-off) / 2);    // ceil(off / 2)
      const int hi_num = L0.k - 1 + off;                         // floor((k - 1 + off) / 2)
      const int hi = hi_num >= 0 ? hi_num / 2 : -((-hi_num + 1) / 2);
      a.tap_lo[p][q] = lo + L.pad_left;
      a.tap_hi[p][q] = hi + L.pad_left;
    }
  return a;
}

// One WN stack (modules.py:69-114) over the whole batch.  x: operand/raw pair holding the stack
// input; on return skipO holds the operand copy of the summed skip output.
int run_wn_deferred(const Ctx& c, const Buffers& bf, int B, int T, int stack, int l_in, int l_rs, int n_layers,
                    const float* gate_bias, int64_t gate_bias_bs, int gate_bias_layer_stride, int reserved_layers,
                    int reserved_sms) {
  const int64_t bs = (int64_t)T * HID;
  const int ald = n_layers * HID;                       // row pitch of the side-by-side activations
  const int64_t abs_ = (int64_t)T * ald;
  const void* x_in = bf.xO;
  void* x_out = bf.xO2;
  for (int i = 0; i < n_layers; ++i) {
    tc_reserve_sms(i < reserved_layers ? reserved_sms : 0);
    char* acts_i = reinterpret_cast<char*>(bf.actsO) + (size_t)i * HID * c.E;
    qvc_conv_args a = layer_args(c, l_in + i, tens(x_in, bs, HID), B, T, T);
    a.epilogue = QVC_EPI_GATE;
    if (gate_bias) { a.bias = gate_bias + (int64_t)i * gate_bias_layer_stride; a.bias_bstride = gate_bias_bs; }
    a.seg[0] = seg(0, HID);
    a.seg[0].op = tens(acts_i, abs_, ald);
    QVC_PROPAGATE(run(c, a));
    if (i == n_layers - 1) break;                       // the last layer has no residual half (modules.py:110-112)
    qvc_conv_args r = layer_args(c, l_rs + i, tens(acts_i, abs_, ald), B, T, T);
    r.cout = HID;                                       // rows [0, HID) of the filter: the residual half
    r.seg[0] = seg(0, HID);
    r.seg[0].res = tens(bf.xR, bs, HID);
    r.seg[0].raw = tens(bf.xR, bs, HID);
    r.seg[0].op = tens(x_out, bs, HID);
    QVC_PROPAGATE(run(c, r));
    const void* tmp = x_in;
    x_in = x_out;
    x_out = const_cast<void*>(tmp);
  }
  tc_reserve_sms(n_layers <= reserved_layers ? reserved_sms : 0);
  qvc_conv_args k = layer_args(c, c.m->wn_skip[stack], tens(bf.actsO, abs_, ald), B, T, T);
  k.seg[0] = seg(0, HID);
  k.seg[0].op = tens(bf.skipO, bs, HID);
  return run(c, k);
}

int run_wn(const Ctx& c, const Buffers& bf, int B, int T, int stack, int l_in, int l_rs, int n_layers,
           const float* gate_bias, int64_t gate_bias_bs, int gate_bias_layer_stride, int reserved_layers = 0,
           int reserved_sms = 0) {
  if (wn_defer_skip(c.m) && n_layers <= WN_MAX_LAYERS && c.m->wn_skip[stack].cin == n_layers * HID)
    return run_wn_deferred(c, bf, B, T, stack, l_in, l_rs, n_layers, gate_bias, gate_bias_bs, gate_bias_layer_stride,
                           reserved_layers, reserved_sms);
  const int64_t bs = (int64_t)T * HID;
  // the operand copy of x ping-pongs between two buffers: a fused layer (qvc_wn_layer) writes the new x while other
  // tiles of the same launch still read the old one as convolution halo
  const void* x_in = bf.xO;
  void* x_out = bf.xO2;
  for (int i = 0; i < n_layers; ++i) {
    // the first `reserved_layers` layers run beside the speaker encoder and leave its SMs out of their grids
    tc_reserve_sms(i < reserved_layers ? reserved_sms : 0);
    qvc_conv_args a = layer_args(c, l_in + i, tens(x_in, bs, HID), B, T, T);
    a.epilogue = QVC_EPI_GATE;
    if (gate_bias) { a.bias = gate_bias + (int64_t)i * gate_bias_layer_stride; a.bias_bstride = gate_bias_bs; }
    a.seg[0] = seg(0, HID);
    a.seg[0].op = tens(bf.actsO, bs, HID);

    qvc_conv_args r = layer_args(c, l_rs + i, tens(bf.actsO, bs, HID), B, T, T);
    if (i < n_layers - 1) {
      r.nseg = 2;
      r.seg[0] = seg(0, HID);                       // residual half: x += ...
      r.seg[0].res = tens(bf.xR, bs, HID);
      r.seg[0].raw = tens(bf.xR, bs, HID);
      r.seg[0].op = tens(x_out, bs, HID);
      r.seg[1] = seg(HID, HID);                     // skip half: out += ...
      if (i > 0) r.seg[1].accin = tens(bf.skipR, bs, HID);
      r.seg[1].raw = tens(bf.skipR, bs, HID);
    } else {
      r.nseg = 1;                                   // last layer: skip only (modules.py:110-112)
      r.seg[0] = seg(0, HID);
      if (i > 0) r.seg[0].accin = tens(bf.skipR, bs, HID);
      r.seg[0].op = tens(bf.skipO, bs, HID);
    }
    int st = QVC_ERR_UNSUPPORTED;
    if (c.m->backend == QVC_BACKEND_TCGEN05) st = qvc_wn_layer(&a, &r, (qvc_stream_t)c.st);
    if (st == QVC_ERR_UNSUPPORTED) {
      QVC_PROPAGATE(run(c, a));
      QVC_PROPAGATE(run(c, r));
    } else {
      QVC_PROPAGATE(st);
    }
    const void* tmp = x_in;
    x_in = x_out;
    x_out = const_cast<void*>(tmp);
  }
  return QVC_OK;
}

// MRF = mean of three ResBlock1 (models.py:378-384, modules.py:147-154) on a sub-batch.
// tcgen05 back end: the last convolution of the three blocks (convs2[2], modules.py:153-154) runs as ONE sum of
// convolutions (qvc_conv1d_sum) -- xs / 3 = (1/3) sum_r (x_r + c2_r(t_r)) accumulates in tensor memory, the fp32 running
// sum of the reference (models.py:380-383) never exists in HBM.  FMA back end: the blocks chain through `accin`.
int run_mrf(const Ctx& c, int cb, int rows, int ch, int l_res0, float* x1R, void* x1O, void* const* tO,
            float* const* xaR, void* const* xaO, float* sumR, qvc_tensor final_op, float final_slope, float* final_raw) {
  const int64_t bs = (int64_t)rows * ch;
  const bool rfo = residual_from_operand(c.m);
  const bool fused_sum = c.m->backend == QVC_BACKEND_TCGEN05;
  qvc_conv_args last[3];
  for (int r = 0; r < 3; ++r) {
    const int lb = l_res0 + 6 * r;
    const float* srcR = x1R;
    const void* srcO = x1O;
    for (int j = 0; j < 3; ++j) {
      // frame-paired layers see every tensor as [rows / 2][2 ch]: the same memory with the row pitch doubled
      // 128-channel layers run in their PLAIN form wherever the frames-on-rows pair kernel serves the family (conv_tcr.cu,
      // QVC_TC_ROWS bit 8: measured faster than the frame-paired channel-major form on every MRF-2 layer,
      // profiles/r02_summary.md) -- at every batch size, so that an utterance's samples do not depend on what it is
      // batched with: small shapes then run the same form on conv_tc, whose accumulation order is the same.  The blocks'
      // last convolutions (j == 2) keep the form the sum kernel shares.
      const bool plain_mrf2 = c.m->backend == QVC_BACKEND_TCGEN05 && rows_kernel_serves_mrf2();
      const bool p1 = !plain_mrf2 && use_frame_pairs(c, lb + j, rows);
      const bool p2 = !(plain_mrf2 && (j < 2 || !fused_sum)) && use_frame_pairs(c, lb + 3 + j, rows);
      const int w1 = p1 ? 2 * ch : ch, w2 = p2 ? 2 * ch : ch;
      qvc_conv_args a = p1 ? paired_args(c, lb + j, srcO, bs, cb, rows, ch)
                           : layer_args(c, lb + j, tens(srcO, bs, ch), cb, rows, rows);
      a.seg[0] = seg(0, w1);
      a.seg[0].slope = 0.1f;
      a.seg[0].op = tens(tO[r], bs, w1);
      QVC_PROPAGATE(run(c, a));

      qvc_conv_args d = p2 ? paired_args(c, lb + 3 + j, tO[r], bs, cb, rows, ch)
                           : layer_args(c, lb + 3 + j, tens(tO[r], bs, ch), cb, rows, rows);
      d.seg[0] = seg(0, w2);
      if (rfo) {
        d.seg[0].res_op = tens(srcO, bs, w2);       // leaky_relu(x, 0.1) in operand format
        d.seg[0].res_inv_slope = 10.f;
      } else {
        d.seg[0].res = tens(srcR, bs, w2);
      }
      if (j < 2) {
        if (!rfo) d.seg[0].raw = tens(xaR[r], bs, w2);
        d.seg[0].op = tens(xaO[r], bs, w2);
        d.seg[0].slope = 0.1f;
        srcR = xaR[r]; srcO = xaO[r];
        QVC_PROPAGATE(run(c, d));
      } else if (fused_sum) {
        last[r] = d;                                // issued below, all three blocks at once
      } else {
        d.seg[0].beta = 1.f / 3.f;
        if (r > 0) d.seg[0].accin = tens(sumR, bs, w2);
        if (r < 2) {
          d.seg[0].raw = tens(sumR, bs, w2);
        } else {
          d.seg[0].op = final_op;
          d.seg[0].op.ld = final_op.ld * (w2 / ch);
          d.seg[0].slope = final_slope;
          if (final_raw) d.seg[0].raw = tens(final_raw, bs, w2);
        }
        QVC_PROPAGATE(run(c, d));
      }
    }
  }
  if (fused_sum) {
    // the three sources share one form (plain or frame-paired: cout decides the view of the output tensors)
    const int w2 = last[0].cout;
    QVC_REQUIRE(last[1].cout == w2 && last[2].cout == w2, "MRF: the ResBlocks' last convolutions differ in form");
    last[0].seg[0].beta = 1.f / 3.f;
    last[0].seg[0].op = final_op;
    last[0].seg[0].op.ld = final_op.ld * (w2 / ch);
    last[0].seg[0].slope = final_slope;
    if (final_raw) last[0].seg[0].raw = tens(final_raw, bs, w2);
    const qvc_conv_args* srcs[3] = {&last[0], &last[1], &last[2]};
    QVC_PROPAGATE(qvc_conv1d_sum(srcs, 3, (qvc_stream_t)c.st));
  }
  return QVC_OK;
}

// Multistream_iSTFT_Generator.forward (models.py:360-408) on the whole batch, sub-batch by sub-batch.
int run_decoder(const Ctx& whole, const Buffers& bf, const Shapes& s, const void* zO, const float* condvec,
                float* wave, const qvc_taps* taps) {
  const int T = s.T, R0 = UP0 * T, R1 = UP0 * UP1 * T, RP = R1 + 1;
  const size_t E = whole.E;
  for (int b0 = 0; b0 < s.B; b0 += s.cb) {
    const int cb = (s.B - b0 < s.cb) ? s.B - b0 : s.cb;
    Ctx c = whole;
    if (c.lengths) c.lengths += b0;
    const char* zO_b = reinterpret_cast<const char*>(zO) + (size_t)b0 * T * HID * E;
    // conv_pre + cond(g): the conditioning is a per-utterance bias (models.py:372)
    {
      qvc_conv_args a = layer_args(c, L_DEC_PRE, tens(zO_b, (int64_t)T * HID, HID), cb, T, T);
      a.bias = condvec + (c.m->cond_rows - C_PRE) + (s.n_embed > 1 ? (int64_t)b0 * c.m->cond_rows : 0);
      a.bias_bstride = s.n_embed > 1 ? c.m->cond_rows : 0;
      a.seg[0] = seg(0, C_PRE);
      a.seg[0].slope = 0.1f;
      a.seg[0].op = tens(bf.aO, (int64_t)T * C_PRE, C_PRE);
      if (taps && taps->conv_pre) a.seg[0].raw = tens(bf.x1R, (int64_t)T * C_PRE, C_PRE);
      QVC_PROPAGATE(run(c, a));
      if (taps && taps->conv_pre)
        QVC_PROPAGATE(from_series_major(bf.x1R, C_PRE, (int64_t)T * C_PRE, taps->conv_pre + (size_t)b0 * C_PRE * T,
                                        cb, C_PRE, T, false, c.st));
    }
    // ups.0 as a 4-tap series convolution with 5*256 phase-major output columns: [T][1280] == [5T][256]
    {
      qvc_conv_args a = layer_args(c, L_UPS + 0, tens(bf.aO, (int64_t)T * C_PRE, C_PRE), cb, T, T);
      a.seg[0] = seg(0, UP0 * C_UP0);
      a.seg[0].slope = 0.1f;
      if (!residual_from_operand(c.m) || (taps && taps->ups0)) a.seg[0].raw = tens(bf.x1R, (int64_t)T * UP0 * C_UP0, UP0 * C_UP0);
      a.seg[0].op = tens(bf.x1O, (int64_t)T * UP0 * C_UP0, UP0 * C_UP0);
      QVC_PROPAGATE(run(c, a));
      if (taps && taps->ups0)
        QVC_PROPAGATE(from_series_major(bf.x1R, C_UP0, (int64_t)R0 * C_UP0, taps->ups0 + (size_t)b0 * C_UP0 * R0,
                                        cb, C_UP0, R0, false, c.st));
    }
    {
      const bool tap = taps && taps->mrf0;
      QVC_PROPAGATE(run_mrf(c, cb, R0, C_UP0, L_RES, bf.x1R, bf.x1O, bf.t1O, bf.xaR, bf.xaO, bf.sum1R,
                            tens(bf.uO, (int64_t)R0 * C_UP0, C_UP0), 0.1f, tap ? bf.xaR[0] : nullptr));
      if (tap)
        QVC_PROPAGATE(from_series_major(bf.xaR[0], C_UP0, (int64_t)R0 * C_UP0, taps->mrf0 + (size_t)b0 * C_UP0 * R0,
                                        cb, C_UP0, R0, false, c.st));
    }
    // ups.1: 5 taps, 4*128 phase-major columns: [5T][512] == [20T][128]
    {
      qvc_conv_args a = layer_args(c, L_UPS + 1, tens(bf.uO, (int64_t)R0 * C_UP0, C_UP0), cb, R0, R0);
      a.seg[0] = seg(0, UP1 * C_UP1);
      a.seg[0].slope = 0.1f;
      if (!residual_from_operand(c.m) || (taps && taps->ups1)) a.seg[0].raw = tens(bf.y1R, (int64_t)R0 * UP1 * C_UP1, UP1 * C_UP1);
      a.seg[0].op = tens(bf.y1O, (int64_t)R0 * UP1 * C_UP1, UP1 * C_UP1);
      QVC_PROPAGATE(run(c, a));
      if (taps && taps->ups1)
        QVC_PROPAGATE(from_series_major(bf.y1R, C_UP1, (int64_t)R1 * C_UP1, taps->ups1 + (size_t)b0 * C_UP1 * R1,
                                        cb, C_UP1, R1, false, c.st));
    }
    {
      // final operand: leaky_relu with the DEFAULT slope 0.01 (models.py:385), written one row
      // down into the reflection-padded series (models.py:388)
      const bool tap = taps && taps->mrf1;
      qvc_tensor fo = tens(reinterpret_cast<char*>(bf.pO) + (size_t)C_UP1 * E, (int64_t)RP * C_UP1, C_UP1);
      QVC_PROPAGATE(run_mrf(c, cb, R1, C_UP1, L_RES + 18, bf.y1R, bf.y1O, bf.t2O, bf.yaR, bf.yaO, bf.sum2R,
                            fo, 0.01f, tap ? bf.yaR[0] : nullptr));
      if (tap)
        QVC_PROPAGATE(from_series_major(bf.yaR[0], C_UP1, (int64_t)R1 * C_UP1, taps->mrf1 + (size_t)b0 * C_UP1 * R1,
                                        cb, C_UP1, R1, false, c.st));
      QVC_PROPAGATE(reflect_row(bf.pO, (int64_t)RP * C_UP1 * E, (int)(C_UP1 * E), cb, c.st));
    }
    {
      qvc_conv_args a = layer_args(c, L_POST, tens(bf.pO, (int64_t)RP * C_UP1, C_UP1), cb, RP, RP);
      a.seg[0] = seg(0, C_POST);
      float* y_mb = (taps && taps->y_mb) ? taps->y_mb + (size_t)b0 * 4 * 4 * R1 : nullptr;
      // post-net convolution and tail in one launch (post_tail.cu): the 72-channel tensor only exists when a tap asks for it
      const bool tap = taps && taps->conv_post;
      if (tap) a.seg[0].raw = tens(bf.cpR, (int64_t)RP * C_POST, C_POST);
      int st = QVC_ERR_UNSUPPORTED;
      if (c.m->backend == QVC_BACKEND_TCGEN05)
        st = qvc_post_tail(&a, &c.m->tail, c.lengths, UP0 * UP1, wave + (size_t)b0 * 16 * R1, y_mb, (qvc_stream_t)c.st);
      if (st == QVC_ERR_UNSUPPORTED) {
        a.seg[0].raw = tens(bf.cpR, (int64_t)RP * C_POST, C_POST);
        QVC_PROPAGATE(run(c, a));
        QVC_PROPAGATE(qvc_tail(&c.m->tail, bf.cpR, C_POST, cb, RP, c.lengths, UP0 * UP1, wave + (size_t)b0 * 16 * R1, y_mb,
                               (qvc_stream_t)c.st));
      } else {
        QVC_PROPAGATE(st);
      }
      if (tap)
        QVC_PROPAGATE(from_series_major(bf.cpR, C_POST, (int64_t)RP * C_POST, taps->conv_post + (size_t)b0 * C_POST * RP,
                                        cb, C_POST, RP, false, c.st));
    }
  }
  return QVC_OK;
}

// Fork / join helper: the speaker encoder (384 strictly sequential LSTM cell steps on one 8-CTA
// cluster, ~1.7 ms) depends only on `mel`, so it runs on a side stream beside the prior encoder and
// joins before the first layer that needs the conditioning vectors (the flow).  One side stream and
// event pair per host thread and device; works under stream capture (fork and join are event edges).
struct SideStream {
  int dev = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};

int side_stream(SideStream** out) {
  static thread_local SideStream pool[16];
  int dev = 0;
  QVC_CHECK_CUDA(cudaGetDevice(&dev));
  QVC_REQUIRE(dev >= 0 && dev < 16, "device index %d out of range", dev);
  SideStream& s = pool[dev];
  if (s.dev != dev) {
    QVC_CHECK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    QVC_CHECK_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    QVC_CHECK_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    s.dev = dev;
  }
  *out = &s;
  return QVC_OK;
}

int check_model(const qvc_model* m) {
  QVC_REQUIRE(m != nullptr, "null model");
  QVC_REQUIRE(m->abi_version == QVC_ABI_VERSION, "model built for ABI %d, library is %d", m->abi_version, QVC_ABI_VERSION);
  QVC_REQUIRE(m->opformat >= QVC_OPF_F32 && m->opformat <= QVC_OPF_F16, "bad opformat %d", m->opformat);
  QVC_REQUIRE(m->backend == QVC_BACKEND_FMA || m->backend == QVC_BACKEND_TCGEN05, "bad backend %d", m->backend);
  QVC_REQUIRE(!(m->backend == QVC_BACKEND_TCGEN05 && m->opformat == QVC_OPF_F32),
              "the tcgen05 back end needs TF32 or BF16 operands");
  QVC_REQUIRE(m->cond_rows == 4 * 4 * 2 * HID + C_PRE, "cond_rows %d != %d", m->cond_rows, 4 * 4 * 2 * HID + C_PRE);
  for (int i = 0; i < QVC_NUM_LAYERS; ++i)
    QVC_REQUIRE(m->layers[i].w != nullptr && m->layers[i].cin > 0 && m->layers[i].cout > 0, "layer %d not populated", i);
  QVC_REQUIRE(m->cond_w && m->cond_b && m->tail.window && m->tail.synth, "model misses cond / tail weights");
  return QVC_OK;
}

}  // namespace

}  // namespace qvc

using namespace qvc;

extern "C" size_t qvc_infer_workspace_bytes(const qvc_model* m, int batch, int frames, int mel_batch, int mel_frames) {
  if (!m || batch <= 0 || frames <= 0) return 0;
  Shapes s{batch, frames, batch, pick_chunk(m, batch, frames)};   // n_embed upper bound = batch
  Buffers b;
  return carve(m, s, mel_batch, mel_frames, mel_batch > 0 && mel_frames > 0, nullptr, 0, &b) + 256;
}

extern "C" int qvc_infer(const qvc_model* m, const float* unit, const float* mel, const float* noise,
                         const float* g_in, const int32_t* lengths, int batch, int frames, int mel_batch, int mel_frames,
                         float* wave, const qvc_taps* taps, void* workspace, size_t workspace_bytes,
                         qvc_stream_t stream) {
  QVC_PROPAGATE(check_model(m));
  QVC_REQUIRE(unit && noise && wave && workspace, "qvc_infer: null pointer");
  QVC_REQUIRE(batch >= 1 && batch <= 65535 && frames >= 1, "qvc_infer: bad shape B=%d T=%d", batch, frames);
  QVC_REQUIRE(g_in || mel, "qvc_infer: need mel or a cached embedding");
  QVC_REQUIRE(mel_batch >= 1, "qvc_infer: mel_batch must be >= 1");
  const bool with_spk = g_in == nullptr;
  if (with_spk) {
    QVC_REQUIRE(mel_frames >= 1, "qvc_infer: empty mel");
    QVC_REQUIRE(mel_frames <= 128 || mel_batch == 1,
                "qvc_infer: mel longer than 128 frames must have batch 1 (got %d), as in the reference (models.py:536)", mel_batch);
  }
  const int n_embed = with_spk ? (mel_frames > 128 ? 1 : mel_batch) : mel_batch;
  QVC_REQUIRE(n_embed == 1 || n_embed == batch, "qvc_infer: %d embeddings cannot broadcast over %d utterances", n_embed, batch);

  Shapes s{batch, frames, n_embed, pick_chunk(m, batch, frames)};
  Buffers bf;
  const uintptr_t mis = (uintptr_t)workspace & 255;
  char* ws = reinterpret_cast<char*>(workspace) + (mis ? 256 - mis : 0);
  const size_t cap = workspace_bytes - (mis ? 256 - mis : 0);
  const size_t need = carve(m, s, mel_batch, mel_frames, with_spk, ws, cap, &bf);
  if (need > cap) {
    set_error("qvc_infer: workspace %zu < %zu", workspace_bytes, need + 256);
    return QVC_ERR_WORKSPACE;
  }
  Ctx c{m, (cudaStream_t)stream, opformat_bytes(m->opformat), frames, lengths};
  const int B = batch, T = frames;
  const int64_t bs = (int64_t)T * HID;

  // inputs -> series-major
  QVC_PROPAGATE(to_series_major(unit, bf.unitO, B, UNIT_CH, T, m->opformat, lengths, c.st));
  QVC_PROPAGATE(to_series_major(noise, bf.noiseT, B, HID, T, QVC_OPF_F32, lengths, c.st));

  // speaker embedding and the per-utterance bias vectors derived from it (models.py:635), on the
  // side stream when the encoder has to run
  const float* g = g_in;
  SideStream* side = nullptr;
  cudaStream_t gst = c.st;
  if (with_spk) {
    QVC_PROPAGATE(side_stream(&side));
    QVC_CHECK_CUDA(cudaEventRecord(side->fork, c.st));
    QVC_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    gst = side->stream;
    QVC_PROPAGATE(qvc_spk_embed(&m->spk, mel, mel_batch, mel_frames, bf.g, bf.spk_ws, bf.spk_ws_bytes, (qvc_stream_t)gst));
    g = bf.g;
  }
  if (taps && taps->g)
    QVC_CHECK_CUDA(cudaMemcpyAsync(taps->g, g, (size_t)n_embed * GIN * 4, cudaMemcpyDeviceToDevice, gst));
  QVC_PROPAGATE(cond_vectors(m->cond_w, m->cond_b, g, n_embed, m->cond_rows, bf.condvec, gst));
  if (side) QVC_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
  const int64_t cond_bs = n_embed > 1 ? m->cond_rows : 0;
  // the prior encoder runs beside the speaker encoder: leave its cluster(s) of 8 SMs out of the convolution grids
  struct Reserve {
    explicit Reserve(int n) { tc_reserve_sms(n); }
    ~Reserve() { tc_reserve_sms(0); }
  };
  int spk_sms = 0;
  if (with_spk) {
    const int nseq = mel_frames > 128 ? (mel_frames - 128 + 63) / 64 + 1 : mel_batch;
    const int spc = spk_sequences_per_cluster(nseq);
    spk_sms = 8 * ((nseq + spc - 1) / spc);         // one 8-CTA cluster per spc windows (lstm.cu)
    if (spk_sms > 64) spk_sms = 64;
  }

  // The prior encoder and the flow run per CHUNK of utterances (wn_chunk_utts): the fp32 residual / skip streams and the
  // operand copies of a chunk (~0.8 MB per 10 s utterance and tensor) then stay in the 126 MB L2 from one WN layer to
  // the next instead of round-tripping HBM 32 times per step.  Utterances are independent, so the arithmetic of an
  // utterance does not depend on the chunking.
  const int wcb = wn_chunk_utts(B, T);
  for (int b0 = 0; b0 < B; b0 += wcb) {
    const int Bc = B - b0 < wcb ? B - b0 : wcb;
    const size_t o1 = (size_t)b0 * T;                       // frames before this chunk
    Ctx cc = c;
    if (cc.lengths) cc.lengths += b0;
    Buffers cf = bf;
    const size_t E = c.E;
    cf.unitO = (char*)bf.unitO + o1 * UNIT_CH * E;
    cf.xO = (char*)bf.xO + o1 * HID * E;     cf.xO2 = (char*)bf.xO2 + o1 * HID * E;
    cf.actsO = (char*)bf.actsO + o1 * (wn_defer_skip(m) ? WN_MAX_LAYERS : 1) * HID * E;
    cf.skipO = (char*)bf.skipO + o1 * HID * E; cf.zO = (char*)bf.zO + o1 * HID * E;
    cf.noiseT = bf.noiseT + o1 * HID;        cf.xR = bf.xR + o1 * HID;
    cf.skipR = bf.skipR + o1 * HID;          cf.zR = bf.zR + o1 * HID;
    cf.tapA = bf.tapA + o1 * HID;            cf.tapB = bf.tapB + o1 * HID;
    const bool first = b0 == 0;

    // prior encoder enc_p (models.py:75-95)
    {
      Reserve reserve(first ? spk_sms : 0);
      qvc_conv_args a = layer_args(cc, L_ENC_PRE, tens(cf.unitO, (int64_t)T * UNIT_CH, UNIT_CH), Bc, T, T);
      a.seg[0] = seg(0, HID);
      a.seg[0].raw = tens(cf.xR, bs, HID);
      a.seg[0].op = tens(cf.xO, bs, HID);
      QVC_PROPAGATE(run(cc, a));
      // The encoder needs ~0.5 ms (three layers of 128 steps at ~1 us); a WN layer takes ~0.1 ms per 32000 frames of
      // batch.  Only the layers that can overlap it give up its SMs; the rest of the stack uses the whole machine
      // (a grid launched while the encoder still holds SMs just has its last CTAs start late).
      const int64_t frames = (int64_t)Bc * T;
      int64_t res_layers = (first && spk_sms) ? (160000 + frames - 1) / frames : 0;
      if (res_layers > 16) res_layers = 16;
      QVC_PROPAGATE(run_wn(cc, cf, Bc, T, 0, L_ENC_IN, L_ENC_RS, 16, nullptr, 0, 0, (int)res_layers, spk_sms));
      tc_reserve_sms(res_layers >= 16 ? spk_sms : 0);
      qvc_conv_args p = layer_args(cc, L_ENC_PROJ, tens(cf.skipO, bs, HID), Bc, T, T);
      p.epilogue = QVC_EPI_SAMPLE;
      p.noise = tens(cf.noiseT, bs, HID);
      p.seg[0] = seg(0, HID);
      p.seg[0].raw = tens(cf.zR, bs, HID);
      p.seg[0].op = tens(cf.zO, bs, HID);
      if (taps && taps->m_p) p.aux0 = tens(cf.tapA, bs, HID);
      if (taps && taps->logs_p) p.aux1 = tens(cf.tapB, bs, HID);
      QVC_PROPAGATE(run(cc, p));
      const size_t ot = (size_t)b0 * HID * T;
      if (taps && taps->m_p) QVC_PROPAGATE(from_series_major(cf.tapA, HID, bs, taps->m_p + ot, Bc, HID, T, false, c.st));
      if (taps && taps->logs_p) QVC_PROPAGATE(from_series_major(cf.tapB, HID, bs, taps->logs_p + ot, Bc, HID, T, false, c.st));
      if (taps && taps->z_p) QVC_PROPAGATE(from_series_major(cf.zR, HID, bs, taps->z_p + ot, Bc, HID, T, false, c.st));
    }

    // flow, reverse direction, Flips folded into the weights (models.py:39-51, modules.py:165-224)
    if (side && first) QVC_CHECK_CUDA(cudaStreamWaitEvent(c.st, side->join, 0));      // conditioning vectors ready
    const float* cond_c = bf.condvec + (n_embed > 1 ? (int64_t)b0 * m->cond_rows : 0);
    for (int cpl = 0; cpl < 4; ++cpl) {
      const int lb = L_FLOW + 10 * cpl;
      qvc_conv_args a = layer_args(cc, lb, tens(cf.zO, bs, HID), Bc, T, T);
      a.seg[0] = seg(0, HID);
      a.seg[0].raw = tens(cf.xR, bs, HID);
      a.seg[0].op = tens(cf.xO, bs, HID);
      QVC_PROPAGATE(run(cc, a));
      QVC_PROPAGATE(run_wn(cc, cf, Bc, T, 1 + cpl, lb + 1, lb + 5, 4, cond_c + cpl * 4 * 2 * HID, cond_bs, 2 * HID));
      qvc_conv_args p = layer_args(cc, lb + 9, tens(cf.skipO, bs, HID), Bc, T, T);
      p.seg[0] = seg(0, HID);
      p.seg[0].alpha = -1.f;                          // x1 - m (modules.py:217)
      p.seg[0].res = tens(cf.zR, bs, HID);
      p.seg[0].raw = tens(cf.zR, bs, HID);
      p.seg[0].op = tens(cf.zO, bs, HID);
      QVC_PROPAGATE(run(cc, p));
      if (taps && taps->flow[cpl])
        QVC_PROPAGATE(from_series_major(cf.zR, HID, bs, taps->flow[cpl] + (size_t)b0 * HID * T, Bc, HID, T, (cpl & 1) == 0, c.st));
    }
  }

  return run_decoder(c, bf, s, bf.zO, bf.condvec, wave, taps);
}

extern "C" int qvc_decode(const qvc_model* m, const float* z, const float* g, int g_batch,
                          const int32_t* lengths, int batch, int frames, float* wave, const qvc_taps* taps, void* workspace,
                          size_t workspace_bytes, qvc_stream_t stream) {
  QVC_PROPAGATE(check_model(m));
  QVC_REQUIRE(z && g && wave && workspace, "qvc_decode: null pointer");
  QVC_REQUIRE(batch >= 1 && batch <= 65535 && frames >= 1, "qvc_decode: bad shape B=%d T=%d", batch, frames);
  QVC_REQUIRE(g_batch == 1 || g_batch == batch, "qvc_decode: %d embeddings cannot broadcast over %d utterances", g_batch, batch);
  Shapes s{batch, frames, g_batch, pick_chunk(m, batch, frames)};
  Buffers bf;
  const uintptr_t mis = (uintptr_t)workspace & 255;
  char* ws = reinterpret_cast<char*>(workspace) + (mis ? 256 - mis : 0);
  const size_t cap = workspace_bytes - (mis ? 256 - mis : 0);
  const size_t need = carve(m, s, 0, 0, false, ws, cap, &bf);
  if (need > cap) {
    set_error("qvc_decode: workspace %zu < %zu", workspace_bytes, need + 256);
    return QVC_ERR_WORKSPACE;
  }
  Ctx c{m, (cudaStream_t)stream, opformat_bytes(m->opformat), frames, lengths};
  QVC_PROPAGATE(to_series_major(z, bf.zO, batch, HID, frames, m->opformat, lengths, c.st));
  QVC_PROPAGATE(cond_vectors(m->cond_w, m->cond_b, g, g_batch, m->cond_rows, bf.condvec, c.st));
  return run_decoder(c, bf, s, bf.zO, bf.condvec, wave, taps);
}
