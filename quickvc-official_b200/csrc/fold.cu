// fold.cu -- qvc_prepare_weights: reference-layout state_dict -> kernel-ready qvc_model, in the library (host code).
//
// Everything the reference recomputes on every forward but that depends only on the weights is done here once per load,
// in double precision on the host (the device only ever runs this library's kernels):
//   * old-style weight norm  w = g * v / ||v||  (every weight_norm(...) site: modules.py:64,67,134-143,193,195;
//     models.py:71,73,327-335,346,356);
//   * Flip (modules.py:165-170) folded into channel permutations of each coupling's pre / post, the 96-channel coupling
//     halves (modules.py:209-222) zero-embedded into 192-channel filters;
//   * cond_layer(g) / dec.cond(g) (modules.py:83-96, models.py:372) as one matrix whose product with the speaker
//     embedding gives all per-utterance bias vectors;
//   * ConvTranspose1d (models.py:333-335) as a stride-1 polyphase series convolution ([T][s*Cout] == [s*T][Cout]);
//   * updown_filter zero-stuffing + the 63-tap synthesis Conv1d (models.py:353-357,405-406) as a 4-phase, 17-tap filter;
//   * the frame-paired form of the dilation-1, 128-channel MRF layers (qvc_model.paired);
//   * operands rounded to the GEMM operand format (TF32 round-to-nearest-away, IEEE half / bf16 round-to-nearest-even).
// quickvc-official_b200/fold.py states the same fold in Python; tests/test_fold_native.py holds the two against each other.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace qvc {

namespace {

constexpr int HID = 192, GIN = 256, N_WN_ENC = 16, N_WN_FLOW = 4, N_CPL = 4;
constexpr int COND_ROWS = N_CPL * N_WN_FLOW * 2 * HID + 512;

struct Store {
  std::unordered_map<std::string, std::pair<const float*, int64_t>> m;
  bool ok = true;
  const float* get(const std::string& key, int64_t numel) {
    auto it = m.find(key);
    if (it == m.end()) {
      if (ok) set_error("qvc_prepare_weights: state_dict entry '%s' is missing", key.c_str());
      ok = false;
      return nullptr;
    }
    if (it->second.second != numel || it->second.first == nullptr) {
      if (ok) set_error("qvc_prepare_weights: '%s' has %lld elements, expected %lld", key.c_str(),
                        (long long)it->second.second, (long long)numel);
      ok = false;
      return nullptr;
    }
    return it->second.first;
  }
  bool has(const std::string& key) const { return m.count(key) != 0; }
};

using Vec = std::vector<double>;

// weight of `prefix` with dim0 x inner elements: weight-norm resolved (norm over every dim but 0), or the plain weight
Vec weight(Store& st, const std::string& prefix, int64_t dim0, int64_t inner) {
  Vec w((size_t)(dim0 * inner), 0.0);
  if (st.has(prefix + ".weight_v")) {
    const float* v = st.get(prefix + ".weight_v", dim0 * inner);
    const float* g = st.get(prefix + ".weight_g", dim0);
    if (!v || !g) return w;
    for (int64_t r = 0; r < dim0; ++r) {
      double ss = 0.0;
      for (int64_t i = 0; i < inner; ++i) ss += (double)v[r * inner + i] * (double)v[r * inner + i];
      const double scale = (double)g[r] / std::sqrt(ss);
      for (int64_t i = 0; i < inner; ++i) w[(size_t)(r * inner + i)] = (double)v[r * inner + i] * scale;
    }
  } else {
    const float* p = st.get(prefix + ".weight", dim0 * inner);
    if (!p) return w;
    for (int64_t i = 0; i < dim0 * inner; ++i) w[(size_t)i] = (double)p[i];
  }
  return w;
}

Vec bias(Store& st, const std::string& prefix, int64_t n) {
  Vec b((size_t)n, 0.0);
  if (st.has(prefix + ".bias")) {
    const float* p = st.get(prefix + ".bias", n);
    if (p)
      for (int64_t i = 0; i < n; ++i) b[(size_t)i] = (double)p[i];
  }
  return b;
}

// ---- output placement: one caller-provided block, 256-byte aligned pieces ----
struct Arena {
  char* host;        // staging (or the final host block for qvc_fold_host)
  char* target;      // address the pointers in qvc_model refer to (device block, or == host)
  size_t off = 0, cap;
  bool dry;
  void* take(size_t bytes, void** host_ptr) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    if (dry || off > cap) { *host_ptr = nullptr; return nullptr; }
    *host_ptr = host + o;
    return target + o;
  }
};

inline uint16_t to_bf16(float f) {           // round to nearest even, as torch's .to(bfloat16)
  uint32_t b;
  memcpy(&b, &f, 4);
  if ((b & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((b >> 16) | 0x40);
  b += 0x7fffu + ((b >> 16) & 1u);
  return (uint16_t)(b >> 16);
}

inline uint16_t to_f16(float f) {            // IEEE half, round to nearest even, overflow to infinity (torch's .to(float16))
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

inline float to_tf32(float f) {              // cvt.rna.tf32.f32: nearest, ties away from zero
  uint32_t b;
  memcpy(&b, &f, 4);
  b = (b + 0x1000u) & ~0x1FFFu;
  memcpy(&f, &b, 4);
  return f;
}

// writes `n` values (double -> fp32 -> operand format) and returns the pointer the model should hold
void* put_operand(Arena& A, const Vec& w, int opf) {
  const size_t n = w.size();
  void* hp = nullptr;
  void* tp = A.take(n * opformat_bytes(opf), &hp);
  if (!hp) return tp;
  if (opf_is16(opf)) {
    uint16_t* d = reinterpret_cast<uint16_t*>(hp);
    for (size_t i = 0; i < n; ++i) d[i] = opf == QVC_OPF_BF16 ? to_bf16((float)w[i]) : to_f16((float)w[i]);
  } else {
    float* d = reinterpret_cast<float*>(hp);
    for (size_t i = 0; i < n; ++i) d[i] = opf == QVC_OPF_TF32 ? to_tf32((float)w[i]) : (float)w[i];
  }
  return tp;
}

float* put_f32(Arena& A, const Vec& w) {
  void* hp = nullptr;
  void* tp = A.take(w.size() * 4, &hp);
  if (hp) {
    float* d = reinterpret_cast<float*>(hp);
    for (size_t i = 0; i < w.size(); ++i) d[i] = (float)w[i];
  }
  return reinterpret_cast<float*>(tp);
}

float* put_raw(Arena& A, const float* src, int64_t n) {
  void* hp = nullptr;
  void* tp = A.take((size_t)n * 4, &hp);
  if (hp && src) memcpy(hp, src, (size_t)n * 4);
  return reinterpret_cast<float*>(tp);
}

struct Builder {
  Store& st;
  Arena& A;
  qvc_model* m;
  int opf;
  int next = 0;

  // filter given as [cout][k][cin]; pads cout to a multiple of 16 with zero rows
  void add_layer(Vec w, Vec b, bool with_bias, int cout, int k, int cin, int dil, int pad_left) {
    const int pad_rows = (16 - cout % 16) % 16;
    if (pad_rows) {
      w.resize((size_t)(cout + pad_rows) * k * cin, 0.0);
      b.resize((size_t)(cout + pad_rows), 0.0);
    }
    const int li = next++;
    if (li >= QVC_NUM_LAYERS) return;
    // dilation-1 layers with 128 channels in and out (MRF-2): also the frame-paired form (qvc_model.paired)
    //   out[2n + p][c] = sum_j sum_ci w[c][j][ci] x[2n + p + j - pad][ci],  2n + p + j - pad = 2 (n + a) + q
    //   =>  w'[p C + c][a - a_min][q cin + ci] = w[c][2a + q - p + pad][ci]
    if (dil == 1 && cin == 128 && cout == 128 && (k & 1) && k > 1 && with_bias) {
      auto fdiv2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };
      const int a_min = fdiv2(0 - pad_left), a_max = fdiv2(1 + (k - 1) - pad_left);
      const int kp = a_max - a_min + 1, C = cout;
      Vec wp((size_t)2 * C * kp * 2 * cin, 0.0);
      for (int p = 0; p < 2; ++p)
        for (int j = 0; j < k; ++j) {
          const int v = p + j - pad_left;
          const int a = fdiv2(v), q = v - 2 * a;
          for (int c = 0; c < C; ++c)
            for (int ci = 0; ci < cin; ++ci)
              wp[(((size_t)(p * C + c) * kp) + (a - a_min)) * 2 * cin + q * cin + ci] = w[((size_t)c * k + j) * cin + ci];
        }
      Vec bp((size_t)2 * C);
      for (int i = 0; i < 2 * C; ++i) bp[(size_t)i] = (double)(float)b[(size_t)(i % C)];
      qvc_layer& P = m->paired[li];
      P.w = put_operand(A, wp, opf);
      P.bias = put_f32(A, bp);
      P.cin = 2 * cin; P.cout = 2 * cout; P.k = kp; P.dil = 1; P.pad_left = -a_min;
    }
    qvc_layer& L = m->layers[li];
    L.w = put_operand(A, w, opf);
    L.bias = with_bias ? put_f32(A, b) : nullptr;
    L.cin = cin; L.cout = cout + pad_rows; L.k = k; L.dil = dil; L.pad_left = pad_left;
  }

  // qvc_model.wn_skip[s]: the skip halves of a WN stack's L res_skip layers (1x1, HID inputs each) side by side along
  // the input-channel axis, biases summed (modules.py:106-112: every layer but the last has 2 HID outputs, the skip
  // half second; the last layer's HID outputs are all skip)
  void skip_sum(int s, const std::string& prefix, int L) {
    Vec f((size_t)HID * L * HID, 0.0), b((size_t)HID, 0.0);
    for (int i = 0; i < L; ++i) {
      const int cout = i < L - 1 ? 2 * HID : HID, r0 = cout - HID;
      const Vec w = weight(st, prefix + std::to_string(i), cout, HID);
      const Vec bi = bias(st, prefix + std::to_string(i), cout);
      for (int n = 0; n < HID; ++n) {
        for (int c = 0; c < HID; ++c) f[(size_t)n * L * HID + (size_t)i * HID + c] = w[(size_t)(r0 + n) * HID + c];
        b[(size_t)n] += (double)(float)bi[(size_t)(r0 + n)];
      }
    }
    qvc_layer& S = m->wn_skip[s];
    S.w = put_operand(A, f, opf);
    S.bias = put_f32(A, b);
    S.cin = L * HID; S.cout = HID; S.k = 1; S.dil = 1; S.pad_left = 0;
  }

  // Conv1d weight (cout, cin, k) -> [cout][k][cin], 'same' padding
  void plain(const std::string& prefix, int cout, int cin, int k, int dil = 1, bool with_bias = true) {
    const Vec w = weight(st, prefix, cout, (int64_t)cin * k);
    Vec f((size_t)cout * k * cin);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int j = 0; j < k; ++j) f[((size_t)o * k + j) * cin + c] = w[((size_t)o * cin + c) * k + j];
    add_layer(std::move(f), with_bias ? bias(st, prefix, cout) : Vec((size_t)cout, 0.0), with_bias, cout, k, cin, dil,
              (k - 1) * dil / 2);
  }
};

// the whole fold; `dry` only measures.  Pointers written into *m refer to A.target.
int fold_all(Store& st, Arena& A, qvc_model* m, int opf, int backend, float* window_host, float* synth_host) {
  memset(m, 0, sizeof(*m));
  m->abi_version = QVC_ABI_VERSION;
  m->opformat = opf;
  m->backend = backend;
  m->chunk_utts = 0;
  Builder B{st, A, m, opf};

  // ---- prior encoder (models.py:71-73, modules.py:64-67) ----
  B.plain("enc_p.pre", HID, 256, 1);
  for (int i = 0; i < N_WN_ENC; ++i) B.plain("enc_p.enc.in_layers." + std::to_string(i), 2 * HID, HID, 5);
  for (int i = 0; i < N_WN_ENC; ++i)
    B.plain("enc_p.enc.res_skip_layers." + std::to_string(i), i < N_WN_ENC - 1 ? 2 * HID : HID, HID, 1);
  B.plain("enc_p.proj", 2 * HID, HID, 1);
  B.skip_sum(0, "enc_p.enc.res_skip_layers.", N_WN_ENC);

  // ---- flow in the execution order of reverse=True: flows 6, 4, 2, 0 (models.py:48).  Couplings 6 and 2 see the
  // channel-reversed state (an odd number of Flips before them); the state is kept in its original orientation and the
  // channel indexing of pre (inputs) and post (outputs) is reversed instead. ----
  Vec cond_w((size_t)COND_ROWS * GIN, 0.0), cond_b((size_t)COND_ROWS, 0.0);
  const int half = HID / 2, flows[4] = {6, 4, 2, 0};
  for (int c = 0; c < N_CPL; ++c) {
    const bool flipped = c % 2 == 0;
    const std::string p = "flow.flows." + std::to_string(flows[c]);
    const Vec w_pre = weight(st, p + ".pre", HID, half);            // (192, 96, 1)
    const Vec w_post = weight(st, p + ".post", half, HID);          // (96, 192, 1)
    const Vec b_post = bias(st, p + ".post", half);
    Vec wp((size_t)HID * HID, 0.0), wq((size_t)HID * HID, 0.0), bq((size_t)HID, 0.0);
    for (int o = 0; o < HID; ++o)
      for (int i = 0; i < half; ++i) {
        if (flipped) wp[(size_t)o * HID + (HID - half) + i] = w_pre[(size_t)o * half + (half - 1 - i)];
        else         wp[(size_t)o * HID + i] = w_pre[(size_t)o * half + i];
      }
    for (int o = 0; o < half; ++o) {
      const int src = flipped ? half - 1 - o : o, dst = flipped ? o : half + o;
      for (int i = 0; i < HID; ++i) wq[(size_t)dst * HID + i] = w_post[(size_t)src * HID + i];
      bq[(size_t)dst] = b_post[(size_t)src];
    }
    B.add_layer(std::move(wp), bias(st, p + ".pre", HID), true, HID, 1, HID, 1, 0);
    for (int i = 0; i < N_WN_FLOW; ++i) B.plain(p + ".enc.in_layers." + std::to_string(i), 2 * HID, HID, 5, 1, false);
    for (int i = 0; i < N_WN_FLOW; ++i)
      B.plain(p + ".enc.res_skip_layers." + std::to_string(i), i < N_WN_FLOW - 1 ? 2 * HID : HID, HID, 1);
    B.add_layer(std::move(wq), std::move(bq), true, HID, 1, HID, 1, 0);
    B.skip_sum(1 + c, p + ".enc.res_skip_layers.", N_WN_FLOW);
    const int rows = N_WN_FLOW * 2 * HID, r0 = c * rows;
    const Vec cw = weight(st, p + ".enc.cond_layer", rows, GIN);
    const Vec cb = bias(st, p + ".enc.cond_layer", rows);
    for (int r = 0; r < rows; ++r) {
      for (int g = 0; g < GIN; ++g) cond_w[(size_t)(r0 + r) * GIN + g] = cw[(size_t)r * GIN + g];
      cond_b[(size_t)(r0 + r)] = cb[(size_t)r];
    }
    for (int i = 0; i < N_WN_FLOW; ++i) {
      const Vec ib = bias(st, p + ".enc.in_layers." + std::to_string(i), 2 * HID);
      for (int r = 0; r < 2 * HID; ++r) cond_b[(size_t)(r0 + i * 2 * HID + r)] += ib[(size_t)r];
    }
  }

  // ---- decoder (models.py:327-346) ----
  B.plain("dec.conv_pre", 512, HID, 7, 1, false);
  {
    const Vec cw = weight(st, "dec.cond", 512, GIN), cb = bias(st, "dec.cond", 512), pb = bias(st, "dec.conv_pre", 512);
    for (int r = 0; r < 512; ++r) {
      for (int g = 0; g < GIN; ++g) cond_w[(size_t)(COND_ROWS - 512 + r) * GIN + g] = cw[(size_t)r * GIN + g];
      cond_b[(size_t)(COND_ROWS - 512 + r)] = cb[(size_t)r] + pb[(size_t)r];
    }
  }
  {
    // ConvTranspose1d weight (Cin, Cout, k) -> [stride*Cout][taps][Cin]:
    //   y[co, s q + r] = sum_{ci, j : (r + p - j) % s == 0} w[ci, co, j] x[ci, q + (r + p - j) / s]
    const int ups[2][4] = {{512, 256, 5, 6}, {256, 128, 4, 6}};       // Cin, Cout, stride, padding; k = 16 (models.py:335)
    for (int u = 0; u < 2; ++u) {
      const int cin = ups[u][0], cout = ups[u][1], s = ups[u][2], pad = ups[u][3], k = 16;
      const std::string p = "dec.ups." + std::to_string(u);
      const Vec w = weight(st, p, cin, (int64_t)cout * k);             // norm per input channel (dim 0)
      auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
      int o_min = 1 << 30, o_max = -(1 << 30);
      for (int r = 0; r < s; ++r)
        for (int j = 0; j < k; ++j)
          if (((r + pad - j) % s + s) % s == 0) {
            const int o = fdiv(r + pad - j, s);
            o_min = o < o_min ? o : o_min;
            o_max = o > o_max ? o : o_max;
          }
      const int taps = o_max - o_min + 1;
      Vec f((size_t)s * cout * taps * cin, 0.0);
      for (int r = 0; r < s; ++r)
        for (int j = 0; j < k; ++j) {
          if (((r + pad - j) % s + s) % s) continue;
          const int o = fdiv(r + pad - j, s);
          for (int co = 0; co < cout; ++co)
            for (int ci = 0; ci < cin; ++ci)
              f[(((size_t)(r * cout + co) * taps) + (o - o_min)) * cin + ci] = w[((size_t)ci * cout + co) * k + j];
        }
      const Vec b1 = bias(st, p, cout);
      Vec bb((size_t)s * cout);
      for (int i = 0; i < s * cout; ++i) bb[(size_t)i] = b1[(size_t)(i % cout)];
      B.add_layer(std::move(f), std::move(bb), true, s * cout, taps, cin, 1, -o_min);
    }
  }
  {
    const int ks[3] = {3, 7, 11}, dils[3] = {1, 3, 5};
    for (int r = 0; r < 6; ++r) {
      const int ch = r < 3 ? 256 : 128, k = ks[r % 3];
      const std::string p = "dec.resblocks." + std::to_string(r);
      for (int j = 0; j < 3; ++j) B.plain(p + ".convs1." + std::to_string(j), ch, ch, k, dils[j]);
      for (int j = 0; j < 3; ++j) B.plain(p + ".convs2." + std::to_string(j), ch, ch, k, 1);
    }
  }
  B.plain("dec.subband_conv_post", 72, 128, 7);
  if (B.next != QVC_NUM_LAYERS) {
    set_error("qvc_prepare_weights: built %d layers, expected %d", B.next, QVC_NUM_LAYERS);
    return QVC_ERR_ARG;
  }
  m->cond_w = put_f32(A, cond_w);
  m->cond_b = put_f32(A, cond_b);
  m->cond_rows = COND_ROWS;

  // ---- speaker encoder (models.py:510-511) ----
  for (int l = 0; l < 3; ++l) {
    const std::string s = std::to_string(l);
    const int in = l == 0 ? 80 : 256;
    m->spk.w_ih[l] = put_raw(A, st.get("enc_spk.lstm.weight_ih_l" + s, 1024 * in), 1024 * in);
    m->spk.w_hh[l] = put_raw(A, st.get("enc_spk.lstm.weight_hh_l" + s, 1024 * 256), 1024 * 256);
    const float* bi = st.get("enc_spk.lstm.bias_ih_l" + s, 1024);
    const float* bh = st.get("enc_spk.lstm.bias_hh_l" + s, 1024);
    Vec b(1024, 0.0);
    if (bi && bh)
      for (int i = 0; i < 1024; ++i) b[(size_t)i] = (double)bi[i] + (double)bh[i];
    m->spk.bias[l] = put_f32(A, b);
  }
  m->spk.lin_w = put_raw(A, st.get("enc_spk.linear.weight", 256 * 256), 256 * 256);
  m->spk.lin_b = put_raw(A, st.get("enc_spk.linear.bias", 256), 256);

  // ---- tail (models.py:350-357): E[s][r][e] with wave[4q + r] = sum_s sum_e E[s][r][e] y[s][q + 8 - e];
  //   up[s'][4m + tau] = 4 sum_s F[s][s'][tau] y[s][m],  wave[n] = sum_{s', j} w_syn[s'][j] up[s'][n + j - 31]
  //   => tap index j = 63 - 4e + tau - r ----
  const float* win = st.get("dec.stft.window", 16);
  const float* ud = st.get("dec.updown_filter", 64);
  const Vec wsyn = weight(st, "dec.multistream_conv_post", 1, 4 * 63);
  Vec E(4 * 4 * 17, 0.0);
  if (ud)
    for (int e = 0; e < 17; ++e)
      for (int r = 0; r < 4; ++r)
        for (int tau = 0; tau < 4; ++tau) {
          const int j = 63 - 4 * e + tau - r;
          if (j < 0 || j >= 63) continue;
          for (int s = 0; s < 4; ++s) {
            double acc = 0.0;
            for (int s2 = 0; s2 < 4; ++s2) acc += (double)ud[(s * 4 + s2) * 4 + tau] * wsyn[(size_t)s2 * 63 + j];
            E[(size_t)(s * 4 + r) * 17 + e] += 4.0 * acc;
          }
        }
  m->tail.window = put_raw(A, win, 16);
  m->tail.synth = put_f32(A, E);
  if (window_host && synth_host && win) {
    memcpy(window_host, win, 16 * 4);
    for (int i = 0; i < 4 * 4 * 17; ++i) synth_host[i] = (float)E[(size_t)i];
    m->tail.window_host = window_host;
    m->tail.synth_host = synth_host;
  }
  return st.ok ? QVC_OK : QVC_ERR_ARG;
}

int make_store(const qvc_state_entry* entries, int n, Store* st) {
  QVC_REQUIRE(entries != nullptr && n > 0, "qvc_prepare_weights: empty state_dict");
  for (int i = 0; i < n; ++i) {
    QVC_REQUIRE(entries[i].name != nullptr, "qvc_prepare_weights: entry %d has no name", i);
    st->m[entries[i].name] = {entries[i].data, entries[i].numel};
  }
  return QVC_OK;
}

int check_format(int opformat, int backend) {
  QVC_REQUIRE(opformat >= QVC_OPF_F32 && opformat <= QVC_OPF_F16, "qvc_prepare_weights: bad opformat %d", opformat);
  QVC_REQUIRE(backend == QVC_BACKEND_FMA || backend == QVC_BACKEND_TCGEN05, "qvc_prepare_weights: bad backend %d", backend);
  QVC_REQUIRE(!(backend == QVC_BACKEND_TCGEN05 && opformat == QVC_OPF_F32),
              "qvc_prepare_weights: the tcgen05 back end needs TF32, FP16 or BF16 operands");
  return QVC_OK;
}

}  // namespace

}  // namespace qvc

using namespace qvc;

extern "C" size_t qvc_prepared_bytes(int opformat) {
  if (opformat < QVC_OPF_F32 || opformat > QVC_OPF_F16) return 0;
  // a dry run of the same placement sequence over an empty store (sizes do not depend on the values)
  Store st;
  st.ok = false;                                   // suppress "missing entry" messages
  Arena A{nullptr, nullptr, 0, 0, true};
  qvc_model m;
  fold_all(st, A, &m, opformat, QVC_BACKEND_FMA, nullptr, nullptr);
  return A.off + 256;
}

extern "C" int qvc_fold_host(const qvc_state_entry* entries, int n_entries, int opformat, int backend, void* block,
                             size_t block_bytes, float* tail_host, qvc_model* model) {
  QVC_REQUIRE(block && model, "qvc_fold_host: null pointer");
  QVC_PROPAGATE(check_format(opformat, backend));
  Store st;
  QVC_PROPAGATE(make_store(entries, n_entries, &st));
  const uintptr_t mis = (uintptr_t)block & 255;
  char* base = reinterpret_cast<char*>(block) + (mis ? 256 - mis : 0);
  Arena A{base, base, 0, block_bytes - (mis ? 256 - mis : 0), false};
  const int s = fold_all(st, A, model, opformat, backend, tail_host, tail_host ? tail_host + 16 : nullptr);
  if (A.off > A.cap) {
    set_error("qvc_fold_host: block of %zu bytes < %zu", block_bytes, A.off + 256);
    return QVC_ERR_WORKSPACE;
  }
  return s;
}

extern "C" int qvc_prepare_weights(const qvc_state_entry* entries, int n_entries, int opformat, int backend,
                                   void* device_block, size_t device_bytes, float* tail_host, qvc_model* model,
                                   qvc_stream_t stream) {
  QVC_REQUIRE(device_block && model, "qvc_prepare_weights: null pointer");
  QVC_PROPAGATE(check_format(opformat, backend));
  Store st;
  QVC_PROPAGATE(make_store(entries, n_entries, &st));
  const uintptr_t mis = (uintptr_t)device_block & 255;
  char* dbase = reinterpret_cast<char*>(device_block) + (mis ? 256 - mis : 0);
  const size_t cap = device_bytes - (mis ? 256 - mis : 0);
  std::vector<char> staging(cap);
  Arena A{staging.data(), dbase, 0, cap, false};
  const int s = fold_all(st, A, model, opformat, backend, tail_host, tail_host ? tail_host + 16 : nullptr);
  if (A.off > A.cap) {
    set_error("qvc_prepare_weights: device block of %zu bytes < %zu (qvc_prepared_bytes)", device_bytes, A.off + 256);
    return QVC_ERR_WORKSPACE;
  }
  QVC_PROPAGATE(s);
  // the staging buffer is pageable: the copy returns once the data has left it, and the stream is synchronised so that
  // any other stream may use the model afterwards
  QVC_CHECK_CUDA(cudaMemcpyAsync(dbase, staging.data(), A.off, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  QVC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return QVC_OK;
}
