// conv_tc.cu -- series convolution as an implicit GEMM on the 5th-generation tensor cores
// (QVC_BACKEND_TCGEN05): TMA -> shared memory -> tcgen05.mma -> TMEM -> fused epilogue.
//
//   D[t][n] = sum_{j<k} sum_{c<cin}  X[t + j*dil - pad][c] * W[n][j][c]
//   A (M = 128 frames per MMA)  = activation rows, K-major (channels contiguous), 128-byte rows
//   B (N = up to 256 columns)   = folded filter rows [n][j][c], K-major
//   one K block = one filter tap j x one 128-byte channel chunk (32 TF32 / 64 bf16 channels)
//
// What makes it a convolution rather than im2col + GEMM: the activation chunk is brought in ONCE per
// channel chunk as a "slab" of (MT*128 + (k-1)*dil) consecutive frames (TMA, 128B swizzle, rows outside
// the utterance zero-filled by the tensor map => the reference's zero padding, no cross-utterance
// leakage); the k taps are k shared-memory descriptors into the same slab, each shifted by j*dil rows.
// Only the filter streams from L2, and one filter stage feeds MT accumulator tiles.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 =
// epilogue (TMEM -> registers -> fused epilogue of common.cuh -> global).
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace qvc {

namespace {

constexpr int TILE_M = 128;
constexpr int ROW_BYTES = 128;            // one swizzle-128B row = one K block of one frame
constexpr int MAX_SMEM = 232448;          // 227 KB
constexpr int NTHREADS = 192;

struct alignas(64) TcParams {
  CUtensorMap mx;                          // x as (channel, frame, utterance)
  CUtensorMap mw;                          // w as (tap*cin + channel, column)
  int32_t cin, k, dil, pad_left;
  int32_t nc;                              // GEMM columns owned by one CTA
  int32_t nsplit, nchunk;                  // nc = nsplit * nchunk, nchunk = N of one MMA
  int32_t mt;                              // 128-frame accumulator tiles per CTA
  int32_t slab_boxes, slab_box_rows;       // slab = slab_boxes TMA boxes of slab_box_rows frames
  int32_t w_boxes, w_box_rows;
  int32_t slab_stages, w_stages;
  uint32_t slab_stage_bytes, w_stage_bytes;
  uint32_t tmem_cols;
  int32_t desc_mode;                       // 0 (default): base_offset 0 -- measured on B200: the 128B swizzle is a
                                           // function of the absolute shared-memory address, so a descriptor may start
                                           // on any 128-byte row of a TMA-written slab; 1 = (addr >> 7) & 7 (wrong, kept
                                           // only for the experiment recorded in profiles/r01_tc_desc_mode.log)
  EpiParams ep;
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  // bounded spin: a protocol bug must trap, not hang the GPU
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int OPF>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (OPF == QVC_OPF_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// K-major, 128-byte-swizzled operand: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, int mode) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  if (mode == 1) d |= (uint64_t)((addr >> 7) & 7u) << 49;   // base offset: phase of the start row in the swizzle pattern
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ----------------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------------
template <int OPF, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int ESIZE = OPF == QVC_OPF_BF16 ? 2 : 4;
  constexpr int KC = ROW_BYTES / ESIZE;            // channels per K block
  constexpr uint32_t FMT = OPF == QVC_OPF_BF16 ? 1u : 2u;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slab0 = smem_base;
  const uint32_t w0 = slab0 + p.slab_stages * p.slab_stage_bytes;
  const uint32_t bar0 = w0 + p.w_stages * p.w_stage_bytes;
  // barrier layout: full_slab[SS] empty_slab[SS] full_w[WS] empty_w[WS] tmem_full, then the TMEM base word
  const uint32_t full_slab = bar0, empty_slab = full_slab + 8 * p.slab_stages;
  const uint32_t full_w = empty_slab + 8 * p.slab_stages, empty_w = full_w + 8 * p.w_stages;
  const uint32_t tmem_full = empty_w + 8 * p.w_stages;
  const uint32_t tmem_slot = tmem_full + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * p.mt * TILE_M;
  const int n0 = blockIdx.y * p.nc;
  const int n_cchunks = p.cin / KC;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * p.slab_stages + 2 * p.w_stages + 1; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mx) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mw) : "memory");
      const uint32_t slab_bytes = (uint32_t)p.slab_boxes * p.slab_box_rows * ROW_BYTES;
      const uint32_t w_bytes = (uint32_t)p.nc * ROW_BYTES;
      int wit = 0;
      for (int cc = 0; cc < n_cchunks; ++cc) {
        const int s = cc % p.slab_stages;
        const uint32_t ph = (uint32_t)(cc / p.slab_stages) & 1u;
        mbar_wait(empty_slab + 8 * s, ph ^ 1u);
        mbar_expect_tx(full_slab + 8 * s, slab_bytes);
        for (int i = 0; i < p.slab_boxes; ++i)
          tma_load_3d(slab0 + s * p.slab_stage_bytes + i * p.slab_box_rows * ROW_BYTES, &p.mx, full_slab + 8 * s,
                      cc * KC, t0 - p.pad_left + i * p.slab_box_rows, b);
        for (int j = 0; j < p.k; ++j, ++wit) {
          const int ws = wit % p.w_stages;
          const uint32_t wph = (uint32_t)(wit / p.w_stages) & 1u;
          mbar_wait(empty_w + 8 * ws, wph ^ 1u);
          mbar_expect_tx(full_w + 8 * ws, w_bytes);
          for (int i = 0; i < p.w_boxes; ++i)
            tma_load_2d(w0 + ws * p.w_stage_bytes + i * p.w_box_rows * ROW_BYTES, &p.mw, full_w + 8 * ws,
                        j * p.cin + cc * KC, n0 + i * p.w_box_rows);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.nchunk >> 3) << 17) |
                             ((uint32_t)(TILE_M >> 4) << 24);
      int wit = 0;
      for (int cc = 0; cc < n_cchunks; ++cc) {
        const int s = cc % p.slab_stages;
        const uint32_t ph = (uint32_t)(cc / p.slab_stages) & 1u;
        mbar_wait(full_slab + 8 * s, ph);
        tc_fence_after();
        const uint32_t slab = slab0 + s * p.slab_stage_bytes;
        for (int j = 0; j < p.k; ++j, ++wit) {
          const int ws = wit % p.w_stages;
          const uint32_t wph = (uint32_t)(wit / p.w_stages) & 1u;
          mbar_wait(full_w + 8 * ws, wph);
          tc_fence_after();
          const uint32_t wst = w0 + ws * p.w_stage_bytes;
          const uint32_t first = (cc | j) == 0 ? 0u : 1u;
          for (int m = 0; m < p.mt; ++m) {
            const uint32_t a_row = slab + (uint32_t)(m * TILE_M + j * p.dil) * ROW_BYTES;
            for (int ns = 0; ns < p.nsplit; ++ns) {
              const uint32_t b_row = wst + (uint32_t)(ns * p.nchunk) * ROW_BYTES;
              const uint32_t d = tmem_base + (uint32_t)(m * p.nc + ns * p.nchunk);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma<OPF>(d, smem_desc(a_row + ks * 32, p.desc_mode), smem_desc(b_row + ks * 32, p.desc_mode), idesc,
                          ks == 0 ? first : 1u);
            }
          }
          tc_commit(empty_w + 8 * ws);            // filter stage free once these MMAs retire
        }
        tc_commit(empty_slab + 8 * s);
      }
      tc_commit(tmem_full);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    for (int m = 0; m < p.mt; ++m) {
      const int t = t0 + m * TILE_M + q * 32 + lane;
      const bool live = t < p.ep.out_rows;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * p.nc);
      if constexpr (EPI == QVC_EPI_LINEAR) {
        for (int c = 0; c < p.nc; c += 16) {
          float v[16];
          tmem_ld16(trow + c, v);
          if (live) {
            epi_linear<OPF, 8>(p.ep, b, t, n0 + c, v);
            epi_linear<OPF, 8>(p.ep, b, t, n0 + c + 8, v + 8);
          }
        }
      } else {
        const int half = p.nc >> 1;
        for (int c = 0; c < half; c += 8) {
          float lo[8], hi[8];
          tmem_ld8(trow + c, lo);
          tmem_ld8(trow + half + c, hi);
          if (live) {
            if constexpr (EPI == QVC_EPI_GATE) epi_gate<OPF, 8>(p.ep, b, t, c, lo, hi);
            else                               epi_sample<OPF, 8>(p.ep, b, t, c, lo, hi);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int desc_mode_from_env() {
  static int mode = [] {
    const char* e = getenv("QVC_TC_DESC_MODE");
    return e ? atoi(e) : 0;
  }();
  return mode;
}

template <int OPF, int EPI>
int launch_variant(const TcParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    QVC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<OPF, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_done = true;
  }
  conv_tc_kernel<OPF, EPI><<<grid, NTHREADS, smem, stream>>>(p);
  return post_launch("conv_tc_kernel");
}

}  // namespace

int launch_conv_tc(const qvc_conv_args& a, cudaStream_t stream) {
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("conv1d(tcgen05): cuTensorMapEncodeTiled not available from the driver");
    return QVC_ERR_CUDA;
  }
  const int esize = (int)opformat_bytes(a.opformat);
  const int kc = ROW_BYTES / esize;
  QVC_REQUIRE(a.cin % kc == 0, "conv1d(tcgen05): cin %d not a multiple of %d", a.cin, kc);
  QVC_REQUIRE((a.x.ld * esize) % 16 == 0 && ((uintptr_t)a.x.ptr & 15) == 0 && ((uintptr_t)a.w & 15) == 0,
              "conv1d(tcgen05): x / w must be 16-byte aligned with 16-byte row pitch");
  QVC_REQUIRE(a.batch == 1 || (a.x.bstride * esize) % 16 == 0, "conv1d(tcgen05): utterance pitch not 16-byte aligned");

  TcParams p{};
  QVC_PROPAGATE(build_epi_params(a, &p.ep));
  p.cin = a.cin; p.k = a.k; p.dil = a.dil; p.pad_left = a.pad_left;
  p.desc_mode = desc_mode_from_env();

  // ---- tile shape ----
  const bool paired = a.epilogue != QVC_EPI_LINEAR;
  int nc;
  if (a.cout <= 512) nc = a.cout;
  else if (a.cout % 256 == 0) nc = 256;
  else if (a.cout % 128 == 0) nc = 128;
  else { set_error("conv1d(tcgen05): cout %d unsupported", a.cout); return QVC_ERR_UNSUPPORTED; }
  QVC_REQUIRE(!paired || nc == a.cout, "conv1d(tcgen05): paired epilogue needs cout <= 512");
  p.nc = nc;
  p.nsplit = nc <= 256 ? 1 : 2;
  p.nchunk = nc / p.nsplit;
  QVC_REQUIRE(p.nchunk % 16 == 0 && p.nchunk <= 256, "conv1d(tcgen05): MMA N %d invalid", p.nchunk);
  const int halo = (a.k - 1) * a.dil;
  const int grid_y = a.cout / nc;

  // filter stage
  p.w_boxes = nc <= 256 ? 1 : 2;
  p.w_box_rows = nc / p.w_boxes;
  p.w_stage_bytes = (uint32_t)nc * ROW_BYTES;

  // accumulator tiles per CTA: as many as TMEM / shared memory allow, but keep the grid >= ~2 waves
  int mt = 512 / nc;
  if (mt > 4) mt = 4;
  if (mt < 1) mt = 1;
  const char* mt_env = getenv("QVC_TC_MT");
  if (mt_env && atoi(mt_env) >= 1 && atoi(mt_env) <= mt) mt = atoi(mt_env);
  while (mt > 1) {
    const long tiles = (long)((a.out_rows + mt * TILE_M - 1) / (mt * TILE_M)) * grid_y * a.batch;
    if (tiles >= 2 * 148 || a.out_rows > (mt / 2) * TILE_M && tiles >= 148) break;
    mt >>= 1;
  }
  int slab_stages = 2, w_stages = 4;
  for (;;) {
    const int rows = mt * TILE_M + halo;
    p.slab_boxes = (rows + 255) / 256;
    p.slab_box_rows = (((rows + p.slab_boxes - 1) / p.slab_boxes) + 7) & ~7;
    p.slab_stage_bytes = (uint32_t)p.slab_boxes * p.slab_box_rows * ROW_BYTES;
    const size_t need = (size_t)slab_stages * p.slab_stage_bytes + (size_t)w_stages * p.w_stage_bytes + 1024 + 256;
    if (need <= (size_t)MAX_SMEM) break;
    if (w_stages > 2) { --w_stages; continue; }
    if (mt > 1) { mt >>= 1; w_stages = 4; continue; }
    set_error("conv1d(tcgen05): tile does not fit shared memory (k=%d dil=%d nc=%d)", a.k, a.dil, nc);
    return QVC_ERR_UNSUPPORTED;
  }
  QVC_REQUIRE(p.slab_box_rows <= 256, "conv1d(tcgen05): slab box too tall");
  p.mt = mt;
  p.slab_stages = slab_stages;
  p.w_stages = w_stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(mt * nc)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)slab_stages * p.slab_stage_bytes + (size_t)w_stages * p.w_stage_bytes + 1024 + 256;

  // ---- tensor maps ----
  const CUtensorMapDataType dt = a.opformat == QVC_OPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.cin, (cuuint64_t)a.x_rows, (cuuint64_t)a.batch};
    cuuint64_t strides[2] = {(cuuint64_t)a.x.ld * esize,
                             (cuuint64_t)(a.batch > 1 ? a.x.bstride : (int64_t)a.x_rows * a.x.ld) * esize};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)p.slab_box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&p.mx, dt, 3, a.x.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05): cuTensorMapEncodeTiled(x) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.k * a.cin, (cuuint64_t)a.cout};
    cuuint64_t strides[1] = {(cuuint64_t)a.k * a.cin * esize};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)p.w_box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&p.mw, dt, 2, const_cast<void*>(a.w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv1d(tcgen05): cuTensorMapEncodeTiled(w) failed: %d", (int)r); return QVC_ERR_CUDA; }
  }

  dim3 grid((a.out_rows + mt * TILE_M - 1) / (mt * TILE_M), grid_y, a.batch);
#define QVC_TC_DISPATCH(OPF)                                                                         \
  switch (a.epilogue) {                                                                              \
    case QVC_EPI_LINEAR: return launch_variant<OPF, QVC_EPI_LINEAR>(p, grid, smem, stream);          \
    case QVC_EPI_GATE:   return launch_variant<OPF, QVC_EPI_GATE>(p, grid, smem, stream);            \
    default:             return launch_variant<OPF, QVC_EPI_SAMPLE>(p, grid, smem, stream);          \
  }
  if (a.opformat == QVC_OPF_BF16) { QVC_TC_DISPATCH(QVC_OPF_BF16) }
  QVC_TC_DISPATCH(QVC_OPF_TF32)
#undef QVC_TC_DISPATCH
}

}  // namespace qvc
