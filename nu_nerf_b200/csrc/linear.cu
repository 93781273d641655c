// linear.cu -- nunerf_linear: C[M,N] = epi(A[M,K] * B[N,K]^T), the dense-layer kernel (forward layers and the
// dX backward GEMMs) on tcgen05 tensor cores.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0       TMA producer: loads the whole weight tile B once (resident in shared memory, reused by every
//                128-row tile of points), then streams A tiles (128 x 64 bf16, 128B swizzle) through a ring;
//   warp 1       tcgen05.mma issuer (one elected lane), accumulator double-buffered in TMEM so the epilogue of
//                tile i overlaps the MMAs of tile i+1; also owns the TMEM allocation;
//   warps 2..17  epilogue: four warps per TMEM lane quarter, one 64-column chunk each.  TMEM -> registers ->
//                bias / activation / mask / addend -> bf16 -> swizzled shared-memory staging -> TMA store.
//                ReLU layers also emit a 1-bit mask per element (32 B per 256-wide row) so the backward never
//                re-reads the activations.
// The epilogue is compiled per feature set (template FEAT) so the hot variants carry no dead code or registers.
//
// Operands are bf16 "planes": one plane (fast mode) or hi/lo planes (split mode: hi*hi + hi*lo + lo*hi, fp32
// accumulation in TMEM).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace nunerf {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int EPI_WARPS = 16;
constexpr int LIN_THREADS = 32 * (2 + EPI_WARPS);
constexpr int STAGE_BYTES = 32 * 128;       // one epilogue staging buffer: 32 rows x 64 bf16

enum : int {
  F_BIAS = 1, F_RELU = 2, F_SOFTPLUS = 4, F_MASKOUT = 8, F_MASKIN = 16, F_AUX = 32, F_ADD = 64, F_LO = 128,
  F_F32 = 256, F_SCALE = 512, F_DIRECT = 1024, F_ALL = 2047
};

struct LinearK {
  int M, N, nkb, nseg;
  int a_seg_col[3];    // column offset of the A plane used by segment s
  int b_seg_plane[3];  // which resident B plane segment s multiplies with
  int b_planes, b_lo_off;
  int stages, num_tiles, acc_stride, tmem_cols;
  const float* bias;
  int act, aux_mode;
  const __nv_bfloat16* aux; int ldaux, aux_lo;
  const __nv_bfloat16* add; int ldadd, add_lo;
  const uint8_t* mask_in; int ldmask_in;
  uint8_t* mask_out; int ldmask_out;
  float out_scale;
  __nv_bfloat16* out; int ldo, out_lo;
  int tma_out;
  float* out32; int ldo32;
  int n_store;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float* f) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) { f[2 * j] = bf16lo_to_f(w[j]); f[2 * j + 1] = bf16hi_to_f(w[j]); }
}

// One 16-column block of one row: x[16] (accumulator + everything the feature set asks for) -> outputs.
template <int FEAT>
__device__ __forceinline__ void epilogue16(const LinearK& p, float* x, const float* s_bias, int c0, long long row,
                                           bool row_ok, uint32_t mbits16, bool to_stage, uint8_t* stage_row, int j16,
                                           int lane, uint32_t* obits16) {
  if constexpr (FEAT & F_BIAS) {
    const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 b = b4[j];
      x[4 * j] += b.x; x[4 * j + 1] += b.y; x[4 * j + 2] += b.z; x[4 * j + 3] += b.w;
    }
  }
  if constexpr (FEAT & F_RELU) {
    if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.0f);
    }
  }
  if constexpr (FEAT & F_SOFTPLUS) {
    if (p.act == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = softplus100(x[j]);
    }
  }
  if constexpr (FEAT & F_MASKIN) {
    if (p.aux_mode == 3) {
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = ((mbits16 >> j) & 1u) ? x[j] : 0.0f;
    }
  }
  if constexpr (FEAT & F_AUX) {
    if (p.aux_mode == 1 || p.aux_mode == 2) {
      const long long rowc = row_ok ? row : 0;
      const uint4* ap = reinterpret_cast<const uint4*>(p.aux + rowc * p.ldaux + c0);
      float a[16];
      unpack16(__ldg(ap), __ldg(ap + 1), a);
      if (p.aux_lo) {
        const uint4* lp = reinterpret_cast<const uint4*>(p.aux + rowc * p.ldaux + p.aux_lo + c0);
        float l[16];
        unpack16(__ldg(lp), __ldg(lp + 1), l);
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] += l[j];
      }
      if (p.aux_mode == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = a[j] > 0.0f ? x[j] : 0.0f;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] *= (1.0f - __expf(-100.0f * a[j]));
      }
    }
  }
  if constexpr (FEAT & F_SCALE) {
    if (p.out_scale != 1.0f) {
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] *= p.out_scale;
    }
  }
  if constexpr (FEAT & F_ADD) {
    if (p.add) {
      const long long rowc = row_ok ? row : 0;
      const uint4* ap = reinterpret_cast<const uint4*>(p.add + rowc * p.ldadd + c0);
      float a[16];
      unpack16(ap[0], ap[1], a);   // plain loads: the addend may alias the output (in-place accumulation)
      if (p.add_lo) {
        const uint4* lp = reinterpret_cast<const uint4*>(p.add + rowc * p.ldadd + p.add_lo + c0);
        float l[16];
        unpack16(lp[0], lp[1], l);
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] += l[j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] += a[j];
    }
  }
  if constexpr (FEAT & F_MASKOUT) {
    uint32_t ob = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) ob |= (x[j] > 0.0f ? 1u : 0u) << j;
    *obits16 = ob;
  }
  uint32_t hw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hw[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
  const bool full16 = (c0 + 16 <= p.n_store);
  if (to_stage) {
    // swizzle-128B staging: 16-byte chunk j16 of row r lives at r*128 + ((j16 ^ (r & 7)) << 4)
    *reinterpret_cast<uint4*>(stage_row + (((j16) ^ (lane & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(stage_row + (((j16 + 1) ^ (lane & 7)) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
  } else {
    if constexpr (FEAT & F_DIRECT) {
      if (p.out && row_ok) {
        __nv_bfloat16* op = p.out + row * p.ldo + c0;
        if (full16) {
          uint4* o4 = reinterpret_cast<uint4*>(op);
          o4[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          o4[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        } else {
          for (int j = 0; j < 16 && c0 + j < p.n_store; ++j) op[j] = __float2bfloat16_rn(x[j]);
        }
      }
    }
  }
  if constexpr (FEAT & F_LO) {
    if (p.out && p.out_lo && row_ok) {
      __nv_bfloat16* op = p.out + row * p.ldo + p.out_lo + c0;
      if (full16) {
        uint32_t lw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          lw[j] = pack_bf16x2(x[2 * j] - bf16lo_to_f(hw[j]), x[2 * j + 1] - bf16hi_to_f(hw[j]));
        uint4* l4 = reinterpret_cast<uint4*>(op);
        l4[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        l4[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
      } else {
        for (int j = 0; j < 16 && c0 + j < p.n_store; ++j) {
          float hi = __bfloat162float(__float2bfloat16_rn(x[j]));
          op[j] = __float2bfloat16_rn(x[j] - hi);
        }
      }
    }
  }
  if constexpr (FEAT & F_F32) {
    if (p.out32 && row_ok) {
      float* o32 = p.out32 + row * p.ldo32 + c0;
      if (full16 && (p.ldo32 & 3) == 0) {
        float4* o4 = reinterpret_cast<float4*>(o32);
#pragma unroll
        for (int j = 0; j < 4; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
      } else {
        for (int j = 0; j < 16 && c0 + j < p.n_store; ++j) o32[j] = x[j];
      }
    }
  }
}

template <int FEAT>
__global__ void __launch_bounds__(LIN_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapO, const LinearK p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_block_bytes = p.N * 128;
  uint8_t* sB = smem;
  uint8_t* sA = sB + (size_t)p.b_planes * p.nkb * b_block_bytes;
  uint8_t* sStage = sA + (size_t)p.stages * A_STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(sStage + (p.tma_out ? EPI_WARPS * STAGE_BYTES : 0));
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* b_full = bars + 2 * p.stages;
  uint64_t* t_full = b_full + 1;   // [2]
  uint64_t* t_empty = t_full + 2;  // [2]
  uint32_t* tmem_ptr = (uint32_t*)(t_empty + 2);
  float* s_bias = (float*)(tmem_ptr + 2);   // 16-byte aligned: (2*stages + 5) * 8 + 8 bytes past a 1 KB boundary

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
    if (p.tma_out) ptx::prefetch_tmap(&mapO);
    for (int i = 0; i < p.stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.0f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      ptx::mbar_expect_tx(b_full, (uint32_t)(p.b_planes * p.nkb * b_block_bytes));
      for (int pl = 0; pl < p.b_planes; ++pl)
        for (int kb = 0; kb < p.nkb; ++kb)
          ptx::tma_load_2d(sB + (size_t)(pl * p.nkb + kb) * b_block_bytes, &mapB, b_full, pl * p.b_lo_off + kb * BK, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int s = 0; s < p.nseg; ++s)
          for (int kb = 0; kb < p.nkb; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full[stage], A_STAGE_BYTES);
            ptx::tma_load_2d(sA + (size_t)stage * A_STAGE_BYTES, &mapA, &full[stage], p.a_seg_col[s] + kb * BK,
                             tile * BM);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = ptx::idesc_bf16(BM, p.N, 0, 0);
      ptx::mbar_wait(b_full, 0);
      ptx::tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
        ptx::mbar_wait(&t_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        uint32_t accum = 0;
        for (int s = 0; s < p.nseg; ++s)
          for (int kb = 0; kb < p.nkb; ++kb) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
            const uint32_t b_addr = ptx::smem_u32(sB + (size_t)(p.b_seg_plane[s] * p.nkb + kb) * b_block_bytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              uint64_t ad = ptx::smem_desc(a_addr + k * 32, 16, 1024);
              uint64_t bd = ptx::smem_desc(b_addr + k * 32, 16, 1024);
              ptx::umma_bf16(d_tmem, ad, bd, idesc, accum);
              accum = 1;
            }
            ptx::tc_commit(&empty[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        ptx::tc_commit(&t_full[acc]);
      }
    }
  } else {
    // ================= epilogue warps (2..17) =================
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int c = ew >> 2;    // the 64-column chunk it owns
    uint8_t* stage_buf = sStage + (size_t)ew * STAGE_BYTES;
    const int cbase = c * 64;
    const bool chunk_active = cbase < p.N && cbase < p.n_store;
    const int ncols = chunk_active ? ((p.N - cbase) < 64 ? (p.N - cbase) : 64) : 0;
    // TMA store only for chunks that are entirely stored (the column clip of a partially valid box is not
    // element exact: whole 16-byte groups get written); other chunks take the guarded per-thread path
    const bool chunk_tma = p.tma_out && chunk_active && (cbase + 64 <= p.n_store);
    bool store_pending = false;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
      const long long row = (long long)tile * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      if (lane == 0) ptx::mbar_wait(&t_full[acc], acc_phase);
      __syncwarp();
      ptx::tc_fence_after();
      if (chunk_active) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride) + cbase;
        unsigned long long mbits = 0ull;
        if constexpr (FEAT & F_MASKIN) {
          if (p.aux_mode == 3)
            mbits = *reinterpret_cast<const unsigned long long*>(p.mask_in + (row_ok ? row : 0) * p.ldmask_in + c * 8);
        }
        if (chunk_tma && store_pending) {
          if (lane == 0) bulk_wait_read0();   // the previous TMA store of this warp has finished reading stage_buf
          __syncwarp();
          store_pending = false;
        }
        unsigned long long obits = 0ull;
        uint8_t* stage_row = stage_buf + lane * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half * 32 < ncols) {
            uint32_t v[32];
            ptx::tmem_ld16(taddr + half * 32, v);
            if (half * 32 + 16 < ncols) ptx::tmem_ld16(taddr + half * 32 + 16, v + 16);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const int sblk = half * 2 + s2;
              if (sblk * 16 < ncols) {
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[s2 * 16 + j]);
                uint32_t ob16 = 0;
                epilogue16<FEAT>(p, x, s_bias, cbase + sblk * 16, row, row_ok, (uint32_t)(mbits >> (sblk * 16)) & 0xffffu,
                                 chunk_tma, stage_row, sblk * 2, lane, &ob16);
                if constexpr (FEAT & F_MASKOUT) obits |= (unsigned long long)ob16 << (sblk * 16);
              }
            }
          }
        }
        if constexpr (FEAT & F_MASKOUT) {
          if (p.mask_out && row_ok)
            *reinterpret_cast<unsigned long long*>(p.mask_out + row * p.ldmask_out + c * 8) = obits;
        }
        if (chunk_tma) {
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapO, stage_buf, cbase, tile * BM + q * 32);   // rows >= M are clipped
            bulk_commit();
          }
          store_pending = true;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&t_empty[acc]);
    }
    if (p.tma_out && lane == 0) bulk_wait_all();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// SIMT debug kernel (same contract, reconstructs hi+lo in fp32).  Selected with impl=1; not a fallback:
// the Python layer only uses it when NUNERF_GEMM_IMPL=simt is set for debugging.
__global__ void linear_simt_kernel(const __nv_bfloat16* A, int lda, int a_lo, const __nv_bfloat16* B, int ldb, int b_lo,
                                   int K, const LinearK p) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)p.M * p.n_store;
  if (idx >= total) return;
  long long row = idx / p.n_store;
  int col = (int)(idx % p.n_store);
  const __nv_bfloat16* a = A + row * lda;
  const __nv_bfloat16* b = B + (long long)col * ldb;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float av = __bfloat162float(a[k]) + (a_lo ? __bfloat162float(a[k + a_lo]) : 0.f);
    float bv = __bfloat162float(b[k]) + (b_lo ? __bfloat162float(b[k + b_lo]) : 0.f);
    acc = fmaf(av, bv, acc);
  }
  float x = acc + (p.bias ? p.bias[col] : 0.f);
  if (p.act == 1) x = fmaxf(x, 0.f);
  else if (p.act == 2) x = softplus100(x);
  if (p.aux_mode == 3) {
    x = ((p.mask_in[row * p.ldmask_in + (col >> 3)] >> (col & 7)) & 1) ? x : 0.f;
  } else if (p.aux_mode) {
    float av = load_planes(p.aux, row * p.ldaux + col, p.aux_lo);
    if (p.aux_mode == 1) x = av > 0.f ? x : 0.f;
    else x *= (1.0f - __expf(-100.0f * av));
  }
  x *= p.out_scale;
  if (p.add) x += load_planes(p.add, row * p.ldadd + col, p.add_lo);
  if (p.out) store_planes(p.out, row * p.ldo + col, p.out_lo, x);
  if (p.out32) p.out32[row * p.ldo32 + col] = x;
  if (p.mask_out) {
    // one thread per element: set the bit atomically (debug path only; the caller zero-fills the mask)
    if (x > 0.f) atomicOr(reinterpret_cast<unsigned int*>(p.mask_out + row * p.ldmask_out) + (col >> 5), 1u << (col & 31));
  }
}

template <int FEAT>
static int launch_tc(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO, const LinearK& k,
                     int grid, size_t smem, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail("linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e), -2);
    configured = true;
  }
  linear_tc_kernel<FEAT><<<grid, LIN_THREADS, smem, stream>>>(mapA, mapB, mapO, k);
  NUNERF_CHECK_LAUNCH("linear_tc_kernel");
  return 0;
}

}  // namespace nunerf

using namespace nunerf;

extern "C" int nunerf_linear(const nunerf_linear_t* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(a && a->A && a->B, "linear: null operand");
  NUNERF_REQUIRE(a->M > 0, "linear: M must be positive");
  NUNERF_REQUIRE(a->N >= 16 && a->N <= 256 && a->N % 16 == 0, "linear: N must be a multiple of 16 in [16,256]");
  NUNERF_REQUIRE(a->K >= 64 && a->K % 64 == 0, "linear: K must be a positive multiple of 64");
  NUNERF_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "linear: lda/ldb must be multiples of 8");
  NUNERF_REQUIRE(a->out || a->out_f32, "linear: no output");
  NUNERF_REQUIRE(!a->out || (a->ldo % 8 == 0 && a->out_lo_off % 8 == 0), "linear: ldo/out_lo_off must be multiples of 8");
  NUNERF_REQUIRE(a->aux_mode >= 0 && a->aux_mode <= 3, "linear: aux_mode must be 0..3");
  NUNERF_REQUIRE(!(a->aux_mode == 1 || a->aux_mode == 2) || (a->aux && a->ldaux % 8 == 0 && a->aux_lo_off % 8 == 0),
                 "linear: bad aux");
  NUNERF_REQUIRE(a->aux_mode != 3 || (a->mask_in && a->ldmask_in % 8 == 0 && ((uintptr_t)a->mask_in & 7) == 0),
                 "linear: bad mask_in (8-byte aligned rows required)");
  NUNERF_REQUIRE(!a->mask_out || (a->ldmask_out % 8 == 0 && ((uintptr_t)a->mask_out & 7) == 0),
                 "linear: bad mask_out (8-byte aligned rows required)");
  NUNERF_REQUIRE(!a->add || (a->ldadd % 8 == 0 && a->add_lo_off % 8 == 0), "linear: bad addend");
  NUNERF_REQUIRE(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->B & 15) == 0, "linear: operands must be 16B aligned");
  LinearK k;
  memset(&k, 0, sizeof(k));
  k.M = a->M; k.N = a->N; k.nkb = a->K / BK;
  const bool a2 = a->a_lo_off != 0, b2 = a->b_lo_off != 0;
  // split mode: hi*hi + hi*lo + lo*hi ; mixed cases degrade gracefully
  k.nseg = 0;
  k.a_seg_col[k.nseg] = 0; k.b_seg_plane[k.nseg] = 0; k.nseg++;
  if (b2) { k.a_seg_col[k.nseg] = 0; k.b_seg_plane[k.nseg] = 1; k.nseg++; }
  if (a2) { k.a_seg_col[k.nseg] = a->a_lo_off; k.b_seg_plane[k.nseg] = 0; k.nseg++; }
  k.b_planes = b2 ? 2 : 1; k.b_lo_off = a->b_lo_off;
  k.num_tiles = cdiv(a->M, BM);
  k.acc_stride = ((a->N + 31) / 32) * 32;
  int tc = 32;
  while (tc < 2 * k.acc_stride) tc <<= 1;
  k.tmem_cols = tc;
  k.bias = a->bias; k.act = a->act; k.aux_mode = a->aux_mode;
  k.aux = (const __nv_bfloat16*)a->aux; k.ldaux = a->ldaux; k.aux_lo = a->aux_lo_off;
  k.add = (const __nv_bfloat16*)a->add; k.ldadd = a->ldadd; k.add_lo = a->add_lo_off;
  k.mask_in = (const uint8_t*)a->mask_in; k.ldmask_in = a->ldmask_in;
  k.mask_out = (uint8_t*)a->mask_out; k.ldmask_out = a->ldmask_out;
  k.out_scale = a->out_scale == 0.f ? 1.f : a->out_scale;
  k.out = (__nv_bfloat16*)a->out; k.ldo = a->ldo; k.out_lo = a->out_lo_off;
  k.out32 = a->out_f32; k.ldo32 = a->ldo32;
  k.n_store = (a->n_store > 0 && a->n_store < a->N) ? a->n_store : a->N;

  if (a->impl == 1) {
    long long total = (long long)k.M * k.n_store;
    linear_simt_kernel<<<cdiv(total, 256), 256, 0, stream>>>((const __nv_bfloat16*)a->A, a->lda, a->a_lo_off,
                                                            (const __nv_bfloat16*)a->B, a->ldb, a->b_lo_off, a->K, k);
    NUNERF_CHECK_LAUNCH("linear_simt_kernel");
    return 0;
  }
  const size_t b_bytes = (size_t)k.b_planes * k.nkb * a->N * 128;
  const size_t fixed0 = 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/;
  const size_t staging = (size_t)EPI_WARPS * STAGE_BYTES;
  NUNERF_REQUIRE(b_bytes + 2 * A_STAGE_BYTES <= 227 * 1024 - fixed0,
                 "linear: weight tile does not fit in shared memory (split N)");
  // bf16 output path: TMA store through swizzled staging when at least 2 activation stages still fit next to the
  // resident weights, else guarded per-thread stores (NUNERF_STORE_MODE = 1 / 0 forces one of them)
  const int stages_tma = (int)(((long long)227 * 1024 - (long long)(fixed0 + staging + b_bytes)) / A_STAGE_BYTES);
  int mode = env_int("NUNERF_STORE_MODE", -1);
  if (mode < 0) mode = stages_tma >= 2 ? 1 : 0;
  if (mode == 1 && stages_tma < 2) mode = 0;
  k.tma_out = (a->out != nullptr && ((uintptr_t)a->out & 15) == 0 && k.n_store >= 64) ? mode : 0;
  const size_t fixed = fixed0 + (k.tma_out ? staging : 0);
  int stages = (int)((227 * 1024 - fixed - b_bytes) / A_STAGE_BYTES);
  if (stages > 8) stages = 8;
  k.stages = stages;
  const size_t smem = fixed + b_bytes + (size_t)stages * A_STAGE_BYTES;
  CUtensorMap mapA, mapB, mapO;
  if (int r = make_map(&mapA, a->A, a->M, a->lda, a->lda, BK, BM)) return r;
  if (int r = make_map(&mapB, a->B, a->N, a->ldb, a->ldb, BK, a->N)) return r;
  if (k.tma_out) {
    if (int r = make_map(&mapO, a->out, a->M, k.n_store, a->ldo, 64, 32)) return r;
  } else {
    mapO = mapA;
  }
  // feature set of this call -> smallest compiled epilogue that covers it
  int need = 0;
  if (a->bias) need |= F_BIAS;
  if (a->act == 1) need |= F_RELU;
  if (a->act == 2) need |= F_SOFTPLUS;
  if (a->mask_out) need |= F_MASKOUT;
  if (a->aux_mode == 3) need |= F_MASKIN;
  if (a->aux_mode == 1 || a->aux_mode == 2) need |= F_AUX;
  if (a->add) need |= F_ADD;
  if (a->out && a->out_lo_off) need |= F_LO;
  if (a->out_f32) need |= F_F32;
  if (k.out_scale != 1.0f) need |= F_SCALE;
  if (a->out && (!k.tma_out || (k.n_store % 64) != 0)) need |= F_DIRECT;
  const int grid = k.num_tiles < num_sms() ? k.num_tiles : num_sms();
  constexpr int V_RELU = F_BIAS | F_RELU | F_MASKOUT | F_DIRECT, V_MASKIN = F_MASKIN | F_DIRECT,
                V_SOFTPLUS = F_BIAS | F_SOFTPLUS | F_DIRECT, V_AUX = F_AUX | F_ADD | F_DIRECT, V_PLAIN = F_DIRECT;
  if ((need & ~V_PLAIN) == 0) return launch_tc<V_PLAIN>(mapA, mapB, mapO, k, grid, smem, stream);
  if ((need & ~V_RELU) == 0) return launch_tc<V_RELU>(mapA, mapB, mapO, k, grid, smem, stream);
  if ((need & ~V_MASKIN) == 0) return launch_tc<V_MASKIN>(mapA, mapB, mapO, k, grid, smem, stream);
  if ((need & ~V_SOFTPLUS) == 0) return launch_tc<V_SOFTPLUS>(mapA, mapB, mapO, k, grid, smem, stream);
  if ((need & ~V_AUX) == 0) return launch_tc<V_AUX>(mapA, mapB, mapO, k, grid, smem, stream);
  return launch_tc<F_ALL>(mapA, mapB, mapO, k, grid, smem, stream);
}
