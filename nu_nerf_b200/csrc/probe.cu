// probe.cu -- micro-benchmark of the tcgen05.mma issue rate on this part (development aid, tools/mma_probe.py):
// one thread per CTA issues `iters` groups of four K = 16 MMAs (M = 128, N given) back to back on fixed operands and
// commits once; the CTA reports the cycles from the first issue to the arrival of the commit.
//   mode 0: A and B from shared memory (SS)      mode 1: A from tensor memory, B from shared memory (TS)
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace nunerf {

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int mode, int N, int iters, int b_stages, int dmode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                         // 128 x 64 bf16
  uint8_t* sB = smem + 16384;                 // b_stages x (256 x 64 bf16)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < (16384 + b_stages * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_ptr, 512u); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint64_t desc_hi = ptx::smem_desc(0, 16, 1024);
    const uint64_t ad = desc_hi | (uint64_t)((ptx::smem_u32(sA) >> 4) & 0x3fff);
    const uint32_t idesc = ptx::idesc_bf16(128, N, 0, 0);
    // every operand of the timed loop is precomputed: the loop body is eight MMAs on fixed registers
    uint64_t bdk[4], adk[4];
    uint32_t atk[4], dk[3];
    const uint64_t bd = desc_hi | (uint64_t)((ptx::smem_u32(sB) >> 4) & 0x3fff);
    for (int k = 0; k < 4; ++k) { bdk[k] = bd + 2 * k; adk[k] = ad + 2 * k; atk[k] = tb + (uint32_t)(k * 8); }
    for (int j = 0; j < 3; ++j) dk[j] = tb + 128u + (uint32_t)(j * 128);
    const uint32_t d1 = dmode >= 1 ? dk[1] : dk[0], d2 = dmode == 2 ? dk[2] : dk[0];
    const uint32_t dA = dk[0], dB = d1, dC = d2, dD = dmode == 2 ? dk[0] : d1;   // rotation of consecutive instructions
    const long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < iters; i += 2) {
        ptx::umma_bf16(dA, adk[0], bdk[0], idesc, 1u); ptx::umma_bf16(dB, adk[1], bdk[1], idesc, 1u);
        ptx::umma_bf16(dC, adk[2], bdk[2], idesc, 1u); ptx::umma_bf16(dD, adk[3], bdk[3], idesc, 1u);
        ptx::umma_bf16(dA, adk[0], bdk[0], idesc, 1u); ptx::umma_bf16(dB, adk[1], bdk[1], idesc, 1u);
        ptx::umma_bf16(dC, adk[2], bdk[2], idesc, 1u); ptx::umma_bf16(dD, adk[3], bdk[3], idesc, 1u);
      }
    } else {
      for (int i = 0; i < iters; i += 2) {
        ptx::umma_bf16_ts(dA, atk[0], bdk[0], idesc, 1u); ptx::umma_bf16_ts(dB, atk[1], bdk[1], idesc, 1u);
        ptx::umma_bf16_ts(dC, atk[2], bdk[2], idesc, 1u); ptx::umma_bf16_ts(dD, atk[3], bdk[3], idesc, 1u);
        ptx::umma_bf16_ts(dA, atk[0], bdk[0], idesc, 1u); ptx::umma_bf16_ts(dB, atk[1], bdk[1], idesc, 1u);
        ptx::umma_bf16_ts(dC, atk[2], bdk[2], idesc, 1u); ptx::umma_bf16_ts(dD, atk[3], bdk[3], idesc, 1u);
      }
    }
    const long long t1 = clock64();
    ptx::tc_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tb, 512u);
}

}  // namespace nunerf

using namespace nunerf;

// out: 2 int64 per CTA (cycles until all MMAs were issued, cycles until they completed)
extern "C" int nunerf_mma_probe(int mode, int N, int iters, int grid, int b_stages, int dmode, long long* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(out && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && grid > 0 && b_stages >= 1 && b_stages <= 6 && (dmode == 0 || N <= 128), "mma_probe: bad arguments");
  const size_t smem = 1024 + 16384 + (size_t)b_stages * 32768;
  cudaError_t e = cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail("mma_probe: cudaFuncSetAttribute: %s", cudaGetErrorString(e), -2);
  mma_probe_kernel<<<grid, 128, smem, stream>>>(mode, N, iters, b_stages, dmode, out);
  NUNERF_CHECK_LAUNCH("mma_probe_kernel");
  return 0;
}
