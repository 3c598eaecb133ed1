// common.cuh -- shared device/host helpers for libqvc_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qvc_b200.h"

namespace qvc {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void note_kernel(const char* name);        // static string: qvc_last_kernel()

#define QVC_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::qvc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return QVC_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define QVC_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::qvc::set_error(__VA_ARGS__);           \
      return QVC_ERR_ARG;                      \
    }                                          \
  } while (0)

#define QVC_PROPAGATE(expr)      \
  do {                           \
    int _s = (expr);             \
    if (_s != QVC_OK) return _s; \
  } while (0)

// Once per (kernel instance, device) set-up such as cudaFuncSetAttribute: function attributes are per device, and a
// process may drive several (one host thread per GPU).  `flags` is the instance's own static array.
constexpr int MAX_DEVICES = 64;
inline bool first_use_on_device(std::atomic<bool>* flags) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return true;
  return !flags[dev].exchange(true, std::memory_order_acq_rel);
}

inline int post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return QVC_ERR_CUDA;
  }
  count_launch();
  note_kernel(what);
  return QVC_OK;
}

// ---------------------------------------------------------------------------------------------
// operand formats
// ---------------------------------------------------------------------------------------------
template <int OPF> struct OpType;
template <> struct OpType<QVC_OPF_F32>  { using type = float; };
template <> struct OpType<QVC_OPF_TF32> { using type = float; };
template <> struct OpType<QVC_OPF_BF16> { using type = __nv_bfloat16; };
template <> struct OpType<QVC_OPF_F16>  { using type = __half; };

constexpr bool opf_is16(int opf) { return opf == QVC_OPF_BF16 || opf == QVC_OPF_F16; }
inline size_t opformat_bytes(int opf) { return opf_is16(opf) ? 2 : 4; }

__device__ __forceinline__ float round_tf32(float v) {
  // round-to-nearest (ties away) to the 10-bit-mantissa TF32 grid; stays an fp32 bit pattern
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

template <int OPF> __device__ __forceinline__ typename OpType<OPF>::type to_operand(float v);
template <> __device__ __forceinline__ float to_operand<QVC_OPF_F32>(float v) { return v; }
template <> __device__ __forceinline__ float to_operand<QVC_OPF_TF32>(float v) { return round_tf32(v); }
template <> __device__ __forceinline__ __nv_bfloat16 to_operand<QVC_OPF_BF16>(float v) {
  return __float2bfloat16_rn(v);
}

// half saturates instead of overflowing to infinity: one out-of-range activation must not poison a whole utterance.
// cvt.rn.satfinite is ONE instruction (F2FP.SATFINITE); the two-FMNMX clamp it replaces made every fp16 epilogue 2
// instructions per element heavier than the bf16 one (a memory-bound MRF-2 layer: 95 us against 83, profiles/r02_summary.md).
template <> __device__ __forceinline__ __half to_operand<QVC_OPF_F16>(float v) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return __ushort_as_half(h);
}

__device__ __forceinline__ float op_to_float(float v) { return v; }
__device__ __forceinline__ float op_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float op_to_float(__half v) { return __half2float(v); }

// the 16 raw bits of a 2-byte operand (zero-extended) -> float
template <int OPF>
__device__ __forceinline__ float op16_to_float(uint32_t bits) {
  if constexpr (OPF == QVC_OPF_BF16) return __uint_as_float(bits << 16);
  else return __half2float(__ushort_as_half((unsigned short)bits));
}
// two floats -> two 2-byte operands packed low / high
template <int OPF>
__device__ __forceinline__ uint32_t op16_pack2(float lo, float hi) {
  if constexpr (OPF == QVC_OPF_BF16) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
  } else {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
}

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + expf(-v)); }

// ---------------------------------------------------------------------------------------------
// device view of qvc_tensor / epilogue description (POD, passed by value to kernels)
// ---------------------------------------------------------------------------------------------
struct TRef {
  void*   ptr;
  int64_t bs;
  int32_t ld;
  __device__ __forceinline__ bool present() const { return ptr != nullptr; }
  template <typename T>
  __device__ __forceinline__ T* at(int b, int row, int col) const {
    return reinterpret_cast<T*>(ptr) + (int64_t)b * bs + (int64_t)row * ld + col;
  }
};

struct EpiSeg {
  int32_t col0, ncols;
  float alpha, beta, slope;
  float res_inv_slope;
  TRef res, res_op, accin, raw, op;
};

struct EpiParams {
  int32_t mode;       // qvc_epilogue
  int32_t nseg;
  int32_t half;       // H for GATE / SAMPLE
  int32_t out_rows;
  const float* bias;
  int64_t bias_bs;
  EpiSeg seg[2];
  TRef noise, aux0, aux1;
  const int32_t* live;  // ragged batches: utterance b has live[b] * live_mul live rows (nullptr: all of them)
  int32_t live_mul;
  // sum of convolutions (qvc_conv1d_sum): residuals / biases of sources 1 and 2 (source 0 is seg[0] / bias)
  int32_t nsum;         // number of sources (1 = an ordinary convolution)
  TRef xres[2], xres_op[2];
  float xres_inv_slope[2];
  const float* xbias[2];
};

// rows of utterance b whose operand outputs are real; the rest are written as zero (qvc_conv_args.live_units)
__device__ __forceinline__ int live_rows(const EpiParams& ep, int b) {
  return ep.live ? ep.live[b] * ep.live_mul : 0x7fffffff;
}

// Ragged batches: a tile whose first row lies DEAD_MARGIN rows or more past its utterance's own end is skipped by all
// warp roles of the tensor-core kernels.  The margin keeps the zero rows a later layer reads as halo past the end
// (at most (k - 1) * dil / 2 = 25 rows on this path) written; rows beyond it are never read by a live tile's live columns.
// `live_s` is the kernel's shared-memory copy of ep.live[0 .. LIVE_CACHE) (load_live_cache): the single-thread TMA / MMA
// roles would otherwise pay a dependent global load per tile, dead or not.
constexpr int DEAD_MARGIN = 32;
constexpr int LIVE_CACHE = 2048;
inline size_t live_cache_bytes(const qvc_conv_args& a) {
  return a.live_units ? 4u * (size_t)(a.batch < LIVE_CACHE ? a.batch : LIVE_CACHE) : 0u;
}
__device__ __forceinline__ void load_live_cache(const EpiParams& ep, int batch, int32_t* live_s) {   // all threads of the CTA
  if (ep.live == nullptr) return;
  const int n = batch < LIVE_CACHE ? batch : LIVE_CACHE;
  for (int i = threadIdx.x; i < n; i += blockDim.x) live_s[i] = ep.live[i];
}
__device__ __forceinline__ bool tile_dead(const EpiParams& ep, const int32_t* live_s, int b, int t0) {
  if (ep.live == nullptr) return false;
  const int lv = b < LIVE_CACHE ? live_s[b] : ep.live[b];
  return t0 >= lv * ep.live_mul + DEAD_MARGIN;
}

inline TRef make_tref(const qvc_tensor& t) { return TRef{t.ptr, t.bstride, t.ld}; }

// Store VEC consecutive operand values.
template <int OPF, int VEC>
__device__ __forceinline__ void store_operand(typename OpType<OPF>::type* dst, const float* v) {
  if constexpr (opf_is16(OPF)) {
    if constexpr (VEC == 2) {
      *reinterpret_cast<uint32_t*>(dst) = op16_pack2<OPF>(v[0], v[1]);
    } else if constexpr (VEC == 4) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(op16_pack2<OPF>(v[0], v[1]), op16_pack2<OPF>(v[2], v[3]));
    } else {
      static_assert(VEC == 8, "VEC");
      *reinterpret_cast<uint4*>(dst) = make_uint4(op16_pack2<OPF>(v[0], v[1]), op16_pack2<OPF>(v[2], v[3]),
                                                  op16_pack2<OPF>(v[4], v[5]), op16_pack2<OPF>(v[6], v[7]));
    }
  } else {
    float r[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) r[i] = to_operand<OPF>(v[i]);
    if constexpr (VEC == 2) {
      *reinterpret_cast<float2*>(dst) = make_float2(r[0], r[1]);
    } else if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
      static_assert(VEC == 8, "VEC");
      reinterpret_cast<float4*>(dst)[0] = make_float4(r[0], r[1], r[2], r[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(r[4], r[5], r[6], r[7]);
    }
  }
}

template <int VEC>
__device__ __forceinline__ void load_f32(const float* src, float* v) {
  if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(src);
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(src + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  }
}

template <int VEC>
__device__ __forceinline__ void store_f32(float* dst, const float* v) {
  if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 4)
      *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Shared epilogues.  `acc` holds VEC consecutive GEMM columns starting at column n (n % VEC == 0)
// of output row (b, t).  Used by both the FMA and the tcgen05 convolution kernels so that the two
// back ends differ only in how the accumulator was produced.
// ---------------------------------------------------------------------------------------------
template <int OPF, int VEC>
__device__ __forceinline__ void epi_linear(const EpiParams& ep, int b, int t, int n, const float* acc) {
  using OT = typename OpType<OPF>::type;
  const bool live = t < live_rows(ep, b);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (s >= ep.nseg) break;
    const EpiSeg& sg = ep.seg[s];
    if (n < sg.col0 || n >= sg.col0 + sg.ncols) continue;
    const int c = n - sg.col0;
    float v[VEC];
    const float* bias = ep.bias ? ep.bias + (int64_t)b * ep.bias_bs + n : nullptr;
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = sg.alpha * (acc[i] + (bias ? bias[i] : 0.f));
    if (sg.res.present()) {
      float r[VEC];
      load_f32<VEC>(sg.res.at<float>(b, t, c), r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] += r[i];
    } else if (sg.res_op.present()) {
      const OT* rp = sg.res_op.at<OT>(b, t, c);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float r = op_to_float(rp[i]);
        v[i] += r > 0.f ? r : r * sg.res_inv_slope;
      }
    }
    if (sg.accin.present()) {
      float r[VEC];
      load_f32<VEC>(sg.accin.at<float>(b, t, c), r);
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = r[i] + sg.beta * v[i];
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = sg.beta * v[i];
    }
    if (sg.raw.present()) store_f32<VEC>(sg.raw.at<float>(b, t, c), v);
    if (sg.op.present()) {
      float a[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) a[i] = live ? leaky(v[i], sg.slope) : 0.f;
      store_operand<OPF, VEC>(sg.op.at<OT>(b, t, c), a);
    }
  }
}

// lo = columns [n, n+VEC) of the first half, hi = the matching columns of the second half.
template <int OPF, int VEC>
__device__ __forceinline__ void epi_gate(const EpiParams& ep, int b, int t, int n, const float* lo,
                                         const float* hi) {
  using OT = typename OpType<OPF>::type;
  const float* bias = ep.bias + (int64_t)b * ep.bias_bs;
  float a[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float ta = lo[i] + bias[n + i];
    float sa = hi[i] + bias[ep.half + n + i];
    a[i] = tanhf(ta) * sigmoid_acc(sa);
  }
  const EpiSeg& sg = ep.seg[0];
  if (sg.raw.present()) store_f32<VEC>(sg.raw.at<float>(b, t, n), a);
  if (sg.op.present()) {
    if (t >= live_rows(ep, b)) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) a[i] = 0.f;
    }
    store_operand<OPF, VEC>(sg.op.at<OT>(b, t, n), a);
  }
}

template <int OPF, int VEC>
__device__ __forceinline__ void epi_sample(const EpiParams& ep, int b, int t, int n, const float* lo,
                                           const float* hi) {
  using OT = typename OpType<OPF>::type;
  const float* bias = ep.bias + (int64_t)b * ep.bias_bs;
  float m[VEC], lg[VEC], z[VEC], nz[VEC];
  load_f32<VEC>(ep.noise.at<float>(b, t, n), nz);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    m[i] = lo[i] + bias[n + i];
    lg[i] = hi[i] + bias[ep.half + n + i];
    z[i] = m[i] + nz[i] * expf(lg[i]);
  }
  const EpiSeg& sg = ep.seg[0];
  if (ep.aux0.present()) store_f32<VEC>(ep.aux0.at<float>(b, t, n), m);
  if (ep.aux1.present()) store_f32<VEC>(ep.aux1.at<float>(b, t, n), lg);
  if (sg.raw.present()) store_f32<VEC>(sg.raw.at<float>(b, t, n), z);
  if (sg.op.present()) {
    if (t >= live_rows(ep, b)) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) z[i] = 0.f;
    }
    store_operand<OPF, VEC>(sg.op.at<OT>(b, t, n), z);
  }
}

// host-side translation of the public argument struct
int build_epi_params(const qvc_conv_args& a, EpiParams* ep);
int add_sum_sources(const qvc_conv_args* const* srcs, int nsrc, EpiParams* ep);

// kernels' host launchers (defined in the respective .cu files)
int launch_conv_fma(const qvc_conv_args& a, cudaStream_t stream);
int launch_conv_tc(const qvc_conv_args& a, cudaStream_t stream);
// sum of nsrc convolutions (qvc_conv1d_sum); srcs[0] carries the epilogue
int launch_conv_tc_sum(const qvc_conv_args* const* srcs, int nsrc, cudaStream_t stream);

}  // namespace qvc
