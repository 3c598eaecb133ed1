// membench.cu -- which global-memory access shapes reach the HBM copy peak on B200?
// Models the conv_tc epilogue (a warp owns 32 channels = one 128-byte line per frame, 32 frames per block,
// rows `pitch` bytes apart) against a plain vectorised copy.  nvcc -arch=sm_100a -O3 membench.cu -o membench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

// P0: plain float4 grid-stride copy
__global__ void copy_vec(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// P1: strip copy.  Tensor [rows][C] fp32.  A block of W warps handles tiles of FR frames; warp w takes channel
// strip (w % (C/32)) and frame sub-range; each thread: 32 loads (stride C), then 32 stores.  MODE: 0 = load+store
// same tensor shape (copy), 1 = read-only (sum to a sink), 2 = write-only.
template <int MODE, int DEPTH>
__global__ void strip_copy(const float* __restrict__ a, float* __restrict__ b, int rows, int C, int frames_per_tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int strips = C / 32;
  const int ntiles = rows / frames_per_tile;
  const int blocks_per_tile = frames_per_tile / 32;          // 32-frame blocks
  float sink = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // work items of the tile: strips x blocks, dealt to warps so that the `strips` warps of one 32-frame block run together
    for (int item = warp; item < strips * blocks_per_tile; item += nw * DEPTH) {
      float v[DEPTH][32];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int it = item + d * nw;
        if (it < strips * blocks_per_tile) {
          const int s = it % strips, blk = it / strips;
          const float* p = a + ((size_t)tile * frames_per_tile + blk * 32) * C + s * 32 + lane;
          if (MODE != 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[d][i] = p[(size_t)i * C];
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[d][i] = (float)(i + lane);
          }
        }
      }
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int it = item + d * nw;
        if (it < strips * blocks_per_tile) {
          const int s = it % strips, blk = it / strips;
          float* q = b + ((size_t)tile * frames_per_tile + blk * 32) * C + s * 32 + lane;
          if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) q[(size_t)i * C] = v[d][i] * 1.0001f;
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) sink += v[d][i];
          }
        }
      }
    }
  }
  if (MODE == 1 && sink == 12345.678f) b[0] = sink;
}

template <typename F>
float time_it(F f, int iters = 10) {
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  f(); f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(s);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(e);
  CK(cudaEventSynchronize(e));
  float ms; cudaEventElapsedTime(&ms, s, e);
  return ms / iters;
}

int main() {
  const int C = 128, rows = 640000;                    // MRF-2 series at B = 64 x 10 s
  const size_t n = (size_t)rows * C;
  float *a, *b;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  const double gb = n * 4 / 1e9;
  {
    float ms = time_it([&] { copy_vec<<<148 * 8, 512>>>((const float4*)a, (float4*)b, n / 4); });
    printf("P0 float4 copy                         : %7.1f us  %6.0f GB/s (r+w)\n", ms * 1e3, 2 * gb / (ms * 1e-3));
  }
  for (int warps : {4, 8, 16, 32}) {
    for (int grid : {148, 296}) {
      float ms = time_it([&] { strip_copy<0, 1><<<grid, warps * 32>>>(a, b, rows, C, 256); });
      printf("P1 strip copy  warps=%2d grid=%3d depth=1 : %7.1f us  %6.0f GB/s (r+w)\n", warps, grid, ms * 1e3, 2 * gb / (ms * 1e-3));
      ms = time_it([&] { strip_copy<0, 2><<<grid, warps * 32>>>(a, b, rows, C, 256); });
      printf("P1 strip copy  warps=%2d grid=%3d depth=2 : %7.1f us  %6.0f GB/s (r+w)\n", warps, grid, ms * 1e3, 2 * gb / (ms * 1e-3));
    }
  }
  for (int warps : {8, 16}) {
    float ms = time_it([&] { strip_copy<1, 1><<<148, warps * 32>>>(a, b, rows, C, 256); });
    printf("P1 strip READ  warps=%2d grid=148 depth=1 : %7.1f us  %6.0f GB/s (r)\n", warps, ms * 1e3, gb / (ms * 1e-3));
    ms = time_it([&] { strip_copy<1, 2><<<148, warps * 32>>>(a, b, rows, C, 256); });
    printf("P1 strip READ  warps=%2d grid=148 depth=2 : %7.1f us  %6.0f GB/s (r)\n", warps, ms * 1e3, gb / (ms * 1e-3));
    ms = time_it([&] { strip_copy<2, 1><<<148, warps * 32>>>(a, b, rows, C, 256); });
    printf("P1 strip WRITE warps=%2d grid=148 depth=1 : %7.1f us  %6.0f GB/s (w)\n", warps, ms * 1e3, gb / (ms * 1e-3));
  }
  return 0;
}
