// conv_fma.cu -- exact-fp32 CUDA-core series convolution (QVC_BACKEND_FMA).
//
// Role: the strict-fp32 mode of the library and the on-GPU cross-check of the tcgen05 kernels
// (same operand tensors, same epilogues, different accumulator producer).  It is also what the
// speaker-encoder input projections run on, where fp32 matters (recurrent error growth).
//
// Tiling: a CTA of 256 threads owns 64 output rows x 64 GEMM columns; each thread a 4x4 micro-tile.
// For the paired epilogues (GATE / SAMPLE) the 64 columns are 32 columns of the first half and the
// matching 32 of the second half, so both members of a pair sit in one thread.
#include "common.cuh"

namespace qvc {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct FmaConvParams {
  const void* x;
  int64_t x_bs;
  int32_t x_ld, x_rows;
  const void* w;
  int32_t cin, cout, k, dil, pad_left;
  EpiParams ep;
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float* v);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* v) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float* v) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<__half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<__half2*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <int OPF, bool PAIRED>
__global__ void __launch_bounds__(256) conv_fma_kernel(const FmaConvParams p) {
  using T = typename OpType<OPF>::type;
  __shared__ float Xs[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];

  const int tid = threadIdx.x;
  const int cn = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * BM;
  const int half = p.cout >> 1;
  // local column l in [0,64) -> GEMM column
  const int nbase = PAIRED ? blockIdx.y * 32 : blockIdx.y * 64;
  auto col_of = [&](int l) -> int {
    if (PAIRED) return l < 32 ? nbase + l : half + nbase + (l - 32);
    return nbase + l;
  };
  const int col_limit_lo = PAIRED ? half : p.cout;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const T* xb = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.x_bs;
  const T* wg = reinterpret_cast<const T*>(p.w);
  const int lrow = tid >> 2, lc4 = (tid & 3) * 4;
  const int wcol = col_of(lrow);
  const bool wvalid = PAIRED ? ((lrow < 32 ? nbase + lrow : nbase + lrow - 32) < half) : (wcol < p.cout);
  (void)col_limit_lo;

  for (int j = 0; j < p.k; ++j) {
    const int xr = t0 + lrow + j * p.dil - p.pad_left;
    const bool xvalid = xr >= 0 && xr < p.x_rows;
    for (int c0 = 0; c0 < p.cin; c0 += BK) {
      float xv[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
      if (xvalid) load4<T>(xb + (int64_t)xr * p.x_ld + c0 + lc4, xv);
      if (wvalid) load4<T>(wg + ((int64_t)wcol * p.k + j) * p.cin + c0 + lc4, wv);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        Xs[lc4 + q][lrow] = xv[q];
        Ws[lc4 + q][lrow] = wv[q];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 xa = *reinterpret_cast<const float4*>(&Xs[kk][4 * ty]);
        const float2 w0 = *reinterpret_cast<const float2*>(&Ws[kk][2 * cn]);
        const float2 w1 = *reinterpret_cast<const float2*>(&Ws[kk][32 + 2 * cn]);
        const float xr4[4] = {xa.x, xa.y, xa.z, xa.w};
        const float wr4[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] = fmaf(xr4[i], wr4[q], acc[i][q]);
      }
    }
  }

  // epilogue: thread owns rows t0+4ty+i, local columns {2cn,2cn+1} and {32+2cn,32+2cn+1}
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + 4 * ty + i;
    if (t >= p.ep.out_rows) continue;
    if (PAIRED) {
      const int n = nbase + 2 * cn;
      if (n >= half) continue;
      const float lo[2] = {acc[i][0], acc[i][1]};
      const float hi[2] = {acc[i][2], acc[i][3]};
      if (p.ep.mode == QVC_EPI_GATE) epi_gate<OPF, 2>(p.ep, b, t, n, lo, hi);
      else                           epi_sample<OPF, 2>(p.ep, b, t, n, lo, hi);
    } else {
      const int n0 = nbase + 2 * cn, n1 = nbase + 32 + 2 * cn;
      const float lo[2] = {acc[i][0], acc[i][1]};
      const float hi[2] = {acc[i][2], acc[i][3]};
      if (n0 < p.cout) epi_linear<OPF, 2>(p.ep, b, t, n0, lo);
      if (n1 < p.cout) epi_linear<OPF, 2>(p.ep, b, t, n1, hi);
    }
  }
}

}  // namespace

int build_epi_params(const qvc_conv_args& a, EpiParams* ep) {
  ep->mode = a.epilogue;
  ep->nseg = a.nseg;
  ep->half = a.cout / 2;
  ep->out_rows = a.out_rows;
  ep->bias = a.bias;
  ep->bias_bs = a.bias_bstride;
  ep->live = a.live_units;
  ep->live_mul = a.live_mul;
  ep->nsum = 1;
  for (int i = 0; i < 2; ++i) {
    ep->xres[i] = TRef{nullptr, 0, 0};
    ep->xres_op[i] = TRef{nullptr, 0, 0};
    ep->xres_inv_slope[i] = 1.f;
    ep->xbias[i] = nullptr;
  }
  QVC_REQUIRE(!a.live_units || a.live_mul >= 1, "conv1d: live_units needs live_mul >= 1 (got %d)", a.live_mul);
  QVC_REQUIRE(a.epilogue >= QVC_EPI_LINEAR && a.epilogue <= QVC_EPI_SAMPLE, "conv1d: bad epilogue %d", a.epilogue);
  if (a.epilogue == QVC_EPI_LINEAR) {
    QVC_REQUIRE(a.nseg == 1 || a.nseg == 2, "conv1d: nseg must be 1 or 2 (got %d)", a.nseg);
  } else {
    QVC_REQUIRE(a.bias != nullptr, "conv1d: GATE/SAMPLE epilogues need a bias vector");
    QVC_REQUIRE(a.cout % 16 == 0, "conv1d: paired epilogue needs cout %% 16 == 0");
    ep->nseg = 1;
  }
  for (int s = 0; s < 2; ++s) {
    const qvc_epi_segment& g = a.seg[s];
    EpiSeg& d = ep->seg[s];
    d.col0 = g.col0; d.ncols = g.ncols; d.alpha = g.alpha; d.beta = g.beta; d.slope = g.slope;
    d.res = make_tref(g.res); d.accin = make_tref(g.accin); d.raw = make_tref(g.raw); d.op = make_tref(g.op);
    d.res_op = make_tref(g.res_op); d.res_inv_slope = g.res_inv_slope;
    if (s < ep->nseg && a.epilogue == QVC_EPI_LINEAR)
      QVC_REQUIRE(!(g.res.ptr && g.res_op.ptr), "conv1d: segment %d has both res and res_op", s);
    if (s < ep->nseg && a.epilogue == QVC_EPI_LINEAR) {
      QVC_REQUIRE(g.col0 % 8 == 0 && g.ncols % 8 == 0, "conv1d: segment %d not 8-column aligned", s);
      const qvc_tensor* ts[5] = {&g.res, &g.accin, &g.raw, &g.op, &g.res_op};
      for (const qvc_tensor* t : ts)
        if (t->ptr) QVC_REQUIRE(t->ld % 8 == 0 && t->bstride % 8 == 0 && ((uintptr_t)t->ptr & 15) == 0,
                                "conv1d: epilogue tensor of segment %d not 8-element / 16-byte aligned", s);
    }
  }
  ep->noise = make_tref(a.noise);
  ep->aux0 = make_tref(a.aux0);
  ep->aux1 = make_tref(a.aux1);
  if (a.epilogue == QVC_EPI_SAMPLE) QVC_REQUIRE(a.noise.ptr != nullptr, "conv1d: SAMPLE epilogue needs noise");
  return QVC_OK;
}

// qvc_conv1d_sum: the residuals and biases of sources 1.. join the epilogue description of source 0
int add_sum_sources(const qvc_conv_args* const* srcs, int nsrc, EpiParams* ep) {
  const qvc_conv_args& a0 = *srcs[0];
  QVC_REQUIRE(a0.epilogue == QVC_EPI_LINEAR && a0.nseg == 1 && !a0.seg[0].accin.ptr,
              "conv1d_sum: source 0 must carry a one-segment LINEAR epilogue without accin");
  QVC_REQUIRE(a0.seg[0].col0 == 0 && a0.seg[0].ncols == a0.cout, "conv1d_sum: the segment must span all %d columns", a0.cout);
  const bool from_op = a0.seg[0].res_op.ptr != nullptr;
  ep->nsum = nsrc;
  for (int i = 1; i < nsrc; ++i) {
    const qvc_conv_args& a = *srcs[i];
    QVC_REQUIRE(a.batch == a0.batch && a.out_rows == a0.out_rows && a.x_rows == a0.x_rows && a.cin == a0.cin &&
                    a.cout == a0.cout && a.opformat == a0.opformat && a.backend == a0.backend && a.bias_bstride == 0 &&
                    a0.bias_bstride == 0,
                "conv1d_sum: source %d does not share the geometry of source 0", i);
    const qvc_epi_segment& g = a.seg[0];
    QVC_REQUIRE((g.res_op.ptr != nullptr) == from_op && (g.res.ptr != nullptr) == (a0.seg[0].res.ptr != nullptr),
                "conv1d_sum: every source must give its residual the same way (res, res_op or none)");
    const qvc_tensor* ts[2] = {&g.res, &g.res_op};
    for (const qvc_tensor* t : ts)
      if (t->ptr) QVC_REQUIRE(t->ld % 8 == 0 && t->bstride % 8 == 0 && ((uintptr_t)t->ptr & 15) == 0,
                              "conv1d_sum: residual of source %d not 8-element / 16-byte aligned", i);
    ep->xres[i - 1] = make_tref(g.res);
    ep->xres_op[i - 1] = make_tref(g.res_op);
    ep->xres_inv_slope[i - 1] = g.res_inv_slope;
    ep->xbias[i - 1] = a.bias;
  }
  return QVC_OK;
}

int launch_conv_fma(const qvc_conv_args& a, cudaStream_t stream) {
  FmaConvParams p;
  p.x = a.x.ptr; p.x_bs = a.x.bstride; p.x_ld = a.x.ld; p.x_rows = a.x_rows;
  p.w = a.w; p.cin = a.cin; p.cout = a.cout; p.k = a.k; p.dil = a.dil; p.pad_left = a.pad_left;
  QVC_PROPAGATE(build_epi_params(a, &p.ep));
  QVC_REQUIRE(a.cin % 16 == 0, "conv1d(fma): cin %d not a multiple of 16", a.cin);
  QVC_REQUIRE(a.x.ld % 4 == 0, "conv1d(fma): x.ld %d not a multiple of 4", a.x.ld);
  const bool paired = a.epilogue != QVC_EPI_LINEAR;
  dim3 grid((a.out_rows + BM - 1) / BM, paired ? (a.cout / 2 + 31) / 32 : (a.cout + BN - 1) / BN, a.batch);
  dim3 block(256);
  if (grid.x == 0 || grid.z == 0) return QVC_OK;
#define QVC_LAUNCH_FMA(OPF)                                                         \
  do {                                                                              \
    if (paired) conv_fma_kernel<OPF, true><<<grid, block, 0, stream>>>(p);          \
    else        conv_fma_kernel<OPF, false><<<grid, block, 0, stream>>>(p);         \
  } while (0)
  switch (a.opformat) {
    case QVC_OPF_F32:  QVC_LAUNCH_FMA(QVC_OPF_F32); break;
    case QVC_OPF_TF32: QVC_LAUNCH_FMA(QVC_OPF_TF32); break;
    case QVC_OPF_BF16: QVC_LAUNCH_FMA(QVC_OPF_BF16); break;
    case QVC_OPF_F16:  QVC_LAUNCH_FMA(QVC_OPF_F16); break;
    default: set_error("conv1d: bad opformat %d", a.opformat); return QVC_ERR_ARG;
  }
#undef QVC_LAUNCH_FMA
  return post_launch("conv_fma_kernel");
}

}  // namespace qvc
