// misc.cu -- layout conversion, speaker-conditioning bias vectors, error plumbing, generic entry.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"
#include "engine.h"

namespace qvc {

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
thread_local const char* g_last_kernel = "";
void note_kernel(const char* name) { g_last_kernel = name; }

namespace {

// (B, C, T) -> [B][T][C], 32x32 tiles through shared memory, both sides coalesced
template <int OPF>
__global__ void __launch_bounds__(256) to_series_kernel(const float* __restrict__ src,
                                                        typename OpType<OPF>::type* __restrict__ dst,
                                                        int C, int T, const int32_t* __restrict__ live) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + (int64_t)b * C * T;
  const int Tb = live ? min(T, live[b]) : T;          // ragged batches: padding frames become zero rows
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, t = t0 + tx;
    tile[ty + 8 * i][tx] = (c < C && t < Tb) ? s[(int64_t)c * T + t] : 0.f;
  }
  __syncthreads();
  auto* d = dst + (int64_t)b * T * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty + 8 * i, c = c0 + tx;
    if (t < T && c < C) d[(int64_t)t * C + c] = to_operand<OPF>(tile[tx][ty + 8 * i]);
  }
}

// [B][T][ld] -> (B, C, T); `reverse` writes channel c to C-1-c (undoes a folded Flip for taps)
__global__ void __launch_bounds__(256) from_series_kernel(const float* __restrict__ src, int ld,
                                                          int64_t src_bs, float* __restrict__ dst,
                                                          int C, int T, int reverse) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + (int64_t)b * src_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (c < C && t < T) ? s[(int64_t)t * ld + c] : 0.f;
  }
  __syncthreads();
  float* d = dst + (int64_t)b * C * T;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, t = t0 + tx;
    if (c < C && t < T) d[(int64_t)(reverse ? C - 1 - c : c) * T + t] = tile[tx][ty + 8 * i];
  }
}

// out[e][r] = cond_b[r] + sum_c cond_w[r][c] * g[e][c]; one warp per row  (modules.py:83-84,
// models.py:372: k=1 convolutions on a length-1 series are per-utterance bias vectors)
__global__ void __launch_bounds__(256) cond_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                   const float* __restrict__ g, int rows, float* __restrict__ out) {
  const int e = blockIdx.y;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float4* wr = reinterpret_cast<const float4*>(w + (int64_t)r * 256);
  const float4* gv = reinterpret_cast<const float4*>(g + (int64_t)e * 256);
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float4 x = __ldg(wr + lane + 32 * i), y = __ldg(gv + lane + 32 * i);
    a = fmaf(x.x, y.x, a); a = fmaf(x.y, y.y, a); a = fmaf(x.z, y.z, a); a = fmaf(x.w, y.w, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[(int64_t)e * rows + r] = a + bias[r];
}

// ReflectionPad1d((1,0)) (models.py:345,388): row 0 of the padded series = row 2 (= x[1])
__global__ void reflect_row_kernel(char* base, int64_t bstride_bytes, int row_bytes) {
  char* p = base + (int64_t)blockIdx.x * bstride_bytes;
  for (int i = threadIdx.x * 16; i < row_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(p + i) = *reinterpret_cast<const uint4*>(p + 2 * (int64_t)row_bytes + i);
}

}  // namespace

int from_series_major(const float* src, int ld, int64_t src_bs, float* dst, int batch, int channels,
                      int frames, bool reverse, cudaStream_t stream) {
  if (batch == 0 || frames == 0) return QVC_OK;
  dim3 grid((frames + 31) / 32, (channels + 31) / 32, batch);
  from_series_kernel<<<grid, 256, 0, stream>>>(src, ld, src_bs, dst, channels, frames, reverse ? 1 : 0);
  return post_launch("from_series_kernel");
}

int cond_vectors(const float* w, const float* bias, const float* g, int n_embed, int rows, float* out,
                 cudaStream_t stream) {
  dim3 grid((rows + 7) / 8, n_embed);
  cond_kernel<<<grid, 256, 0, stream>>>(w, bias, g, rows, out);
  return post_launch("cond_kernel");
}

int reflect_row(void* base, int64_t bstride_bytes, int row_bytes, int batch, cudaStream_t stream) {
  reflect_row_kernel<<<batch, 64, 0, stream>>>(reinterpret_cast<char*>(base), bstride_bytes, row_bytes);
  return post_launch("reflect_row_kernel");
}

}  // namespace qvc

using namespace qvc;

namespace qvc {
int to_series_major(const float* src, void* dst, int batch, int channels, int frames, int opformat,
                    const int32_t* live, cudaStream_t st) {
  QVC_REQUIRE(src && dst, "qvc_to_series_major: null pointer");
  QVC_REQUIRE(batch >= 0 && batch <= 65535 && channels > 0 && frames >= 0, "qvc_to_series_major: bad shape");
  if (batch == 0 || frames == 0) return QVC_OK;
  dim3 grid((frames + 31) / 32, (channels + 31) / 32, batch);
  switch (opformat) {
    case QVC_OPF_F32:  to_series_kernel<QVC_OPF_F32><<<grid, 256, 0, st>>>(src, (float*)dst, channels, frames, live); break;
    case QVC_OPF_TF32: to_series_kernel<QVC_OPF_TF32><<<grid, 256, 0, st>>>(src, (float*)dst, channels, frames, live); break;
    case QVC_OPF_BF16: to_series_kernel<QVC_OPF_BF16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16*)dst, channels, frames, live); break;
    case QVC_OPF_F16:  to_series_kernel<QVC_OPF_F16><<<grid, 256, 0, st>>>(src, (__half*)dst, channels, frames, live); break;
    default: set_error("qvc_to_series_major: bad opformat %d", opformat); return QVC_ERR_ARG;
  }
  return post_launch("to_series_kernel");
}
}  // namespace qvc

extern "C" int qvc_to_series_major(const float* src, void* dst, int batch, int channels, int frames,
                                   int opformat, qvc_stream_t stream) {
  return to_series_major(src, dst, batch, channels, frames, opformat, nullptr, (cudaStream_t)stream);
}

extern "C" int qvc_from_series_major(const float* src, int ld, float* dst, int batch, int channels,
                                     int frames, qvc_stream_t stream) {
  QVC_REQUIRE(src && dst, "qvc_from_series_major: null pointer");
  QVC_REQUIRE(batch >= 0 && batch <= 65535 && channels > 0 && ld >= channels, "qvc_from_series_major: bad shape");
  return from_series_major(src, ld, (int64_t)frames * ld, dst, batch, channels, frames, false, (cudaStream_t)stream);
}

extern "C" int qvc_conv1d(const qvc_conv_args* a, qvc_stream_t stream) {
  QVC_REQUIRE(a != nullptr, "qvc_conv1d: null args");
  QVC_REQUIRE(a->x.ptr && a->w, "qvc_conv1d: null x / w");
  QVC_REQUIRE(a->batch >= 0 && a->batch <= 65535, "qvc_conv1d: batch %d out of range", a->batch);
  QVC_REQUIRE(a->k >= 1 && a->dil >= 1 && a->cin > 0 && a->cout > 0 && a->x_rows >= 0 && a->out_rows >= 0,
              "qvc_conv1d: bad geometry");
  QVC_REQUIRE(a->cout % 16 == 0, "qvc_conv1d: cout %d not a multiple of 16", a->cout);
  if (a->batch == 0 || a->out_rows == 0) return QVC_OK;
  if (a->backend == QVC_BACKEND_FMA) return launch_conv_fma(*a, (cudaStream_t)stream);
  if (a->backend == QVC_BACKEND_TCGEN05) {
    QVC_REQUIRE(a->opformat != QVC_OPF_F32, "qvc_conv1d: the tcgen05 back end needs TF32 or BF16 operands");
    return launch_conv_tc(*a, (cudaStream_t)stream);
  }
  set_error("qvc_conv1d: bad backend %d", a->backend);
  return QVC_ERR_ARG;
}

extern "C" int qvc_conv1d_sum(const qvc_conv_args* const* srcs, int nsrc, qvc_stream_t stream) {
  QVC_REQUIRE(srcs != nullptr && nsrc >= 1 && nsrc <= QVC_MAX_SUM_SOURCES, "qvc_conv1d_sum: 1 <= nsrc <= %d (got %d)",
              QVC_MAX_SUM_SOURCES, nsrc);
  for (int i = 0; i < nsrc; ++i) {
    const qvc_conv_args* a = srcs[i];
    QVC_REQUIRE(a != nullptr && a->x.ptr && a->w, "qvc_conv1d_sum: source %d has null x / w", i);
    QVC_REQUIRE(a->batch >= 0 && a->batch <= 65535, "qvc_conv1d_sum: batch %d out of range", a->batch);
    QVC_REQUIRE(a->k >= 1 && a->dil >= 1 && a->cin > 0 && a->cout > 0 && a->x_rows >= 0 && a->out_rows >= 0 &&
                    a->x_rows == a->out_rows && a->cout % 16 == 0,
                "qvc_conv1d_sum: bad geometry in source %d", i);
  }
  if (srcs[0]->backend != QVC_BACKEND_TCGEN05) return QVC_ERR_UNSUPPORTED;
  QVC_REQUIRE(srcs[0]->opformat != QVC_OPF_F32, "qvc_conv1d_sum: the tcgen05 back end needs TF32, FP16 or BF16 operands");
  if (srcs[0]->batch == 0 || srcs[0]->out_rows == 0) return QVC_OK;
  return launch_conv_tc_sum(srcs, nsrc, (cudaStream_t)stream);
}

extern "C" const char* qvc_last_error(void) { return g_err; }
extern "C" int qvc_abi_version(void) { return QVC_ABI_VERSION; }
extern "C" uint64_t qvc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* qvc_last_kernel(void) { return g_last_kernel; }

extern "C" int qvc_host_register(void* ptr, size_t bytes) {
  QVC_REQUIRE(ptr != nullptr && bytes > 0, "qvc_host_register: empty range");
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();                      // do not leave the error pending for the next launch check
    set_error("qvc_host_register: cudaHostRegister(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
    return QVC_ERR_CUDA;
  }
  return QVC_OK;
}

extern "C" int qvc_host_unregister(void* ptr) {
  QVC_REQUIRE(ptr != nullptr, "qvc_host_unregister: null pointer");
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("qvc_host_unregister: cudaHostUnregister -> %s", cudaGetErrorString(e));
    return QVC_ERR_CUDA;
  }
  return QVC_OK;
}

extern "C" int qvc_check_device(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device / driver: %s", cudaGetErrorString(e));
    return QVC_ERR_NO_DEVICE;
  }
  QVC_REQUIRE(dev >= 0 && dev < n, "device %d out of range (%d devices)", dev, n);
  cudaDeviceProp pr;
  QVC_CHECK_CUDA(cudaGetDeviceProperties(&pr, dev));
  // the library holds sm_100a SASS only and no PTX: arch-specific ("a") code does not run on any other part, sm_103 included
  if (pr.major != 10 || pr.minor != 0) {
    set_error("device %d is sm_%d%d; libqvc_b200 only carries sm_100a code (B200)", dev, pr.major, pr.minor);
    return QVC_ERR_NO_DEVICE;
  }
  return QVC_OK;
}
