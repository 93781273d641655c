// gemm.cu -- dense-layer engine on tcgen05 tensor cores (sm_100a).
//
//   nunerf_linear     C[M,N] = epi(A[M,K] * B[N,K]^T)      (forward layers and the dX backward GEMMs)
//   nunerf_linear_dw  dW[N,K] += dZ[M,N]^T * X[M,K]        (weight gradients, reduction over points)
//
// Both are persistent, warp-specialised kernels: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer
// (+ TMEM allocation), warps 2..5 = epilogue (TMEM -> registers -> global).  Operands are bf16 "planes":
// one plane (fast mode) or hi/lo planes (split mode, 3 MMAs per product: hi*hi + hi*lo + lo*hi, fp32
// accumulation in TMEM) -- the split mode is what meets the 1e-4 rgb parity gate against the fp32 reference.
//
// nunerf_linear keeps the whole weight tile resident in shared memory (it is re-used by every 128-row tile
// of points) and streams only the activations through a TMA ring; the accumulator is double-buffered in
// TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#include <cuda.h>
#include <mutex>

namespace nunerf {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int env_int(const char* name, int dflt);

// ------------------------------------------------------------------------------------------- TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// bf16 row-major [rows, ld]; box = box_cols x box_rows, 128B swizzle (box_cols must be 64)
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("%s", "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  static const int promo_env = env_int("NUNERF_L2_PROMO", 128);
  const CUtensorMapL2promotion promo = promo_env == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : promo_env == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo_env == 0  ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                                         : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d) rows=%lld ld=%lld box=%dx%d base=%p", (int)r,
             rows, ld, box_cols, box_rows, base);
    return -3;
  }
  return 0;
}

int env_int(const char* name, int dflt);
static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// ------------------------------------------------------------------------------------------- dW kernel
constexpr int DW_THREADS = 192;
constexpr int DW_BOX_BYTES = 64 * 128;  // 64 points x 64 bf16

struct DwK {
  int M, N, Kc;            // Kc = columns of X handled by this launch (<= 256)
  int n0_stride;           // 128
  int z_planes, z_lo, x_planes, x_lo, x_col0;
  int nseg;                // 1 or 3
  int chunk_pts;           // points per CTA (multiple of 64)
  int stages, stage_bytes, z_bytes_plane, x_bytes_plane, tmem_cols;
  float* dW; int lddw; int dw_col0;
  float* db;               // optional bias gradient: db[n] += sum_m dZ[m,n] (fused column sum, k0 == 0 launch only)
  int n_halves;            // 128-row halves of dZ^T handled by one CTA (2 = whole 256-wide layer per CTA)
  uint32_t lbo, sbo;
};

__global__ void __launch_bounds__(DW_THREADS, 1)
dw_tc_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapX, const DwK p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* t_full = bars + 2 * p.stages;
  uint32_t* tmem_ptr = (uint32_t*)(t_full + 1);
  float* s_colsum = (float*)(bars + 32);   // 256 floats, 256 bytes past the barrier block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128 * p.n_halves;
  const long long m_begin = (long long)blockIdx.y * p.chunk_pts;
  long long m_end = m_begin + p.chunk_pts;
  if (m_end > p.M) m_end = p.M;
  const int nsteps = m_begin < m_end ? (int)((m_end - m_begin + 63) / 64) : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapZ);
    ptx::prefetch_tmap(&mapX);
    // a stage is released by the MMA commit and, when the bias gradient is fused, by the 4 column-sum warps
    for (int i = 0; i < p.stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], p.db ? 5 : 1); }
    ptx::mbar_init(t_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_colsum[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int xboxes = p.Kc / 64;
  const int zboxes = 2 * p.n_halves;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int st = 0; st < nsteps; ++st) {
        const int m = (int)(m_begin + (long long)st * 64);
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full[stage], (uint32_t)p.stage_bytes);
        uint8_t* base = smem + (size_t)stage * p.stage_bytes;
        for (int pl = 0; pl < p.z_planes; ++pl)
          for (int j = 0; j < zboxes; ++j)
            ptx::tma_load_2d(base + pl * p.z_bytes_plane + j * DW_BOX_BYTES, &mapZ, &full[stage],
                             pl * p.z_lo + n0 + j * 64, m);
        uint8_t* xb = base + p.z_planes * p.z_bytes_plane;
        for (int pl = 0; pl < p.x_planes; ++pl)
          for (int j = 0; j < xboxes; ++j)
            ptx::tma_load_2d(xb + pl * p.x_bytes_plane + j * DW_BOX_BYTES, &mapX, &full[stage],
                             pl * p.x_lo + p.x_col0 + j * 64, m);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nsteps > 0) {
      const uint32_t idesc = ptx::idesc_bf16(128, p.Kc, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int st = 0; st < nsteps; ++st) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t zb = ptx::smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t xb = zb + p.z_planes * p.z_bytes_plane;
        for (int h = 0; h < p.n_halves; ++h)
          for (int s = 0; s < p.nseg; ++s) {
            // segments: (z_hi,x_hi), (z_hi,x_lo), (z_lo,x_hi)
            const uint32_t za = zb + (s == 2 ? p.z_bytes_plane : 0) + h * 2 * DW_BOX_BYTES;
            const uint32_t xa = xb + (s == 1 ? p.x_bytes_plane : 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint64_t ad = ptx::smem_desc(za + k * 2048, p.lbo, p.sbo);
              uint64_t bd = ptx::smem_desc(xa + k * 2048, p.lbo, p.sbo);
              ptx::umma_bf16(tmem_base + (uint32_t)(h * p.Kc), ad, bd, idesc, (accum || s || k) ? 1u : 0u);
            }
          }
        accum = 1;
        ptx::tc_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      ptx::tc_commit(t_full);
    }
  } else if (nsteps > 0) {
    const int q = warp & 3;
    if (p.db) {
      // ---- fused bias gradient: column sums of the dZ tiles as they pass through shared memory.
      // thread t owns the 8 columns of group g = t & 15 (per 128-column half) for the points p = psub + 8 i;
      // (p & 7) == psub, so the 128B-swizzled 16-byte chunk of those columns sits at ((g & 7) ^ psub).
      const int t = threadIdx.x - 64;
      const int g = t & 15, psub = t >> 4;
      const uint32_t off = (uint32_t)((g >> 3) * DW_BOX_BYTES + psub * 128 + (((g & 7) ^ psub) << 4));
      float acc[2][8];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[h][j] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int st = 0; st < nsteps; ++st) {
        ptx::mbar_wait(&full[stage], phase);
        const uint8_t* zb = smem + (size_t)stage * p.stage_bytes;
        for (int pl = 0; pl < p.z_planes; ++pl)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < p.n_halves) {
              const uint8_t* src = zb + pl * p.z_bytes_plane + h * 2 * DW_BOX_BYTES + off;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(src + i * 1024);
                acc[h][0] += bf16lo_to_f(v.x); acc[h][1] += bf16hi_to_f(v.x);
                acc[h][2] += bf16lo_to_f(v.y); acc[h][3] += bf16hi_to_f(v.y);
                acc[h][4] += bf16lo_to_f(v.z); acc[h][5] += bf16hi_to_f(v.z);
                acc[h][6] += bf16lo_to_f(v.w); acc[h][7] += bf16hi_to_f(v.w);
              }
            }
          }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (h < p.n_halves)
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(&s_colsum[h * 128 + g * 8 + j], acc[h][j]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = t; c < 128 * p.n_halves; c += 128)
        if (n0 + c < p.N) atomicAdd(p.db + n0 + c, s_colsum[c]);
    }
    ptx::mbar_wait(t_full, 0);
    ptx::tc_fence_after();
    for (int h = 0; h < p.n_halves; ++h) {
      const int n = n0 + h * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * p.Kc);
      for (int c0 = 0; c0 < p.Kc; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(taddr + c0, v);
        ptx::tmem_ld_wait();
        if (n < p.N) {
          // 16-byte vector reductions (REDG.ADD.F32x4): a quarter of the L2 atomic operations of scalar atomicAdd
          float* dst = p.dW + (long long)n * p.lddw + p.dw_col0 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                         : "memory");
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

__global__ void dw_simt_kernel(const __nv_bfloat16* Z, int ldz, int z_lo, const __nv_bfloat16* X, int ldx, int x_lo,
                               int M, int N, int K, float* dW, int lddw, int rows_per_block) {
  // grid: (N*K/256, chunks); each thread one (n,k) entry, partial sum over a chunk of points
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * K) return;
  int n = e / K, k = e % K;
  long long m0 = (long long)blockIdx.y * rows_per_block, m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float acc = 0.f;
  for (long long m = m0; m < m1; ++m)
    acc = fmaf(load_planes(Z, m * ldz + n, z_lo), load_planes(X, m * ldx + k, x_lo), acc);
  atomicAdd(dW + (long long)n * lddw + k, acc);
}

// ------------------------------------------------------------------------------------------- helpers
__global__ void colsum_kernel(const __nv_bfloat16* Z, int ldz, int z_lo, int M, int N, float* out, int rows_per_block) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  long long m0 = (long long)blockIdx.y * rows_per_block, m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float acc = 0.f;
  for (long long m = m0; m < m1; ++m) acc += load_planes(Z, m * ldz + n, z_lo);
  atomicAdd(out + n, acc);
}

__global__ void to_planes_kernel(const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                                 __nv_bfloat16* dst, int dst_rows, int ld, int lo_off, int col_off, int row_off,
                                 int width) {
  // writes dst[row_off + r, col_off + c] for r < dst_rows, c < width (zero where outside the source)
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)dst_rows * width) return;
  int r = (int)(idx / width), c = (int)(idx % width);
  int sr = transpose ? c : r, sc = transpose ? r : c;
  float v = 0.f;
  if (sr < rows && sc < cols) v = src[(long long)sr * ld_src + sc] * scale;
  store_planes(dst, (long long)(row_off + r) * ld + col_off + c, lo_off, v);
}

__global__ void from_planes_kernel(const __nv_bfloat16* src, int rows, int cols, int ld, int lo_off, float* dst,
                                   int ld_dst) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  int r = (int)(idx / cols), c = (int)(idx % cols);
  dst[(long long)r * ld_dst + c] = load_planes(src, (long long)r * ld + c, lo_off);
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

}  // namespace nunerf

using namespace nunerf;

extern "C" const char* nunerf_last_error(void) { return g_err; }
extern "C" int nunerf_version(void) { return 100; }
extern "C" long long nunerf_launch_count(void) { return g_launches.load(); }

extern "C" int nunerf_linear_dw(const nunerf_dw_t* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(a && a->dZ && a->X && a->dW, "dw: null operand");
  NUNERF_REQUIRE(a->M > 0 && a->N >= 1 && a->N <= 256, "dw: bad M/N");
  NUNERF_REQUIRE(a->K >= 64 && a->K % 64 == 0 && a->K <= 1024, "dw: K must be a multiple of 64");
  NUNERF_REQUIRE(a->ldz % 8 == 0 && a->ldx % 8 == 0, "dw: ld must be multiples of 8");
  NUNERF_REQUIRE(a->lddw % 4 == 0 && ((uintptr_t)a->dW & 15) == 0, "dw: dW must be 16-byte aligned with a pitch % 4 == 0");
  if (a->impl == 1) {
    int chunks = 64;
    int rpb = cdiv(a->M, chunks);
    dim3 grid(cdiv((long long)a->N * a->K, 256), cdiv(a->M, rpb));
    dw_simt_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)a->dZ, a->ldz, a->z_lo_off,
                                             (const __nv_bfloat16*)a->X, a->ldx, a->x_lo_off, a->M, a->N, a->K, a->dW,
                                             a->lddw, rpb);
    NUNERF_CHECK_LAUNCH("dw_simt_kernel");
    return 0;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail("dw: cudaFuncSetAttribute: %s", cudaGetErrorString(e), -2);
    configured = true;
  }
  const bool split = a->z_lo_off != 0 && a->x_lo_off != 0;
  CUtensorMap mapZ, mapX;
  if (int r = make_map(&mapZ, a->dZ, a->M, a->ldz, a->ldz, 64, 64)) return r;
  if (int r = make_map(&mapX, a->X, a->M, a->ldx, a->ldx, 64, 64)) return r;
  // one 128-row half of dW per CTA by default (grid = halves x point chunks; the second half re-reads X from L2).
  // NUNERF_DW_HALVES=2 lets one CTA own the whole <= 256-wide layer (both halves in TMEM): measured SLOWER on B200
  // (166 us vs 108 us for 384k x 256 x 256, profiles/r1b_micro_dw.txt), kept for experiments only.
  const int n_halves = (a->N > 128 && !split && env_int("NUNERF_DW_HALVES", 1) == 2) ? 2 : 1;
  const int n_tiles = cdiv(a->N, 128 * n_halves);
  for (int k0 = 0; k0 < a->K; k0 += 256) {
    DwK k;
    memset(&k, 0, sizeof(k));
    k.M = a->M; k.N = a->N; k.Kc = (a->K - k0) < 256 ? (a->K - k0) : 256;
    k.z_planes = split ? 2 : 1; k.x_planes = split ? 2 : 1;
    k.z_lo = a->z_lo_off; k.x_lo = a->x_lo_off; k.x_col0 = k0;
    k.nseg = split ? 3 : 1;
    k.n_halves = n_halves;
    k.db = (k0 == 0) ? a->db : nullptr;
    k.z_bytes_plane = n_halves * 2 * DW_BOX_BYTES;
    k.x_bytes_plane = (k.Kc / 64) * DW_BOX_BYTES;
    k.stage_bytes = k.z_planes * k.z_bytes_plane + k.x_planes * k.x_bytes_plane;
    int stages = (int)((size_t)(223 * 1024) / k.stage_bytes);
    if (stages > 6) stages = 6;
    NUNERF_REQUIRE(stages >= 2, "dw: stage does not fit");
    k.stages = stages;
    int tc = 32;
    while (tc < n_halves * k.Kc) tc <<= 1;
    k.tmem_cols = tc;
    k.dW = a->dW; k.lddw = a->lddw; k.dw_col0 = k0;
    k.lbo = (uint32_t)env_int("NUNERF_DW_LBO", 8192);
    k.sbo = (uint32_t)env_int("NUNERF_DW_SBO", 1024);
    int chunks = num_sms() / n_tiles;
    if (chunks < 1) chunks = 1;
    int steps_total = cdiv(a->M, 64);
    if (chunks > steps_total) chunks = steps_total;
    k.chunk_pts = cdiv(steps_total, chunks) * 64;
    chunks = cdiv(a->M, k.chunk_pts);
    const size_t smem = 1024 + (size_t)k.stages * k.stage_bytes + 256 + 1024;
    dw_tc_kernel<<<dim3(n_tiles, chunks), DW_THREADS, smem, stream>>>(mapZ, mapX, k);
    NUNERF_CHECK_LAUNCH("dw_tc_kernel");
  }
  return 0;
}

extern "C" int nunerf_colsum(const void* Z, int ldz, int z_lo_off, int M, int N, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(Z && out && M > 0 && N > 0, "colsum: bad arguments");
  int chunks = 4 * num_sms();
  int rpb = cdiv(M, chunks);
  if (rpb < 64) rpb = 64;
  dim3 grid(cdiv(N, 128), cdiv(M, rpb));
  colsum_kernel<<<grid, 128, 0, stream>>>((const __nv_bfloat16*)Z, ldz, z_lo_off, M, N, out, rpb);
  NUNERF_CHECK_LAUNCH("colsum_kernel");
  return 0;
}

extern "C" int nunerf_to_planes(const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                                void* dst, int dst_rows, int dst_cols, int ld, int lo_off, int col_off, int row_off,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(src && dst && dst_rows > 0 && dst_cols > 0, "to_planes: bad arguments");
  long long total = (long long)dst_rows * dst_cols;
  to_planes_kernel<<<cdiv(total, 256), 256, 0, stream>>>(src, rows, cols, ld_src, transpose, scale,
                                                        (__nv_bfloat16*)dst, dst_rows, ld, lo_off, col_off, row_off,
                                                        dst_cols);
  NUNERF_CHECK_LAUNCH("to_planes_kernel");
  return 0;
}

extern "C" int nunerf_from_planes(const void* src, int rows, int cols, int ld, int lo_off, float* dst, int ld_dst,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(src && dst && rows > 0 && cols > 0, "from_planes: bad arguments");
  long long total = (long long)rows * cols;
  from_planes_kernel<<<cdiv(total, 256), 256, 0, stream>>>((const __nv_bfloat16*)src, rows, cols, ld, lo_off, dst,
                                                          ld_dst);
  NUNERF_CHECK_LAUNCH("from_planes_kernel");
  return 0;
}
