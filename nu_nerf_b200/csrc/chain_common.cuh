// chain_common.cuh -- parameter block shared by the two fused-chain kernels (chain.cu: activations in shared memory,
// SS-mode MMAs, two tiles per CTA; chain_ts.cu: activations in tensor memory, TS-mode MMAs, one tile per CTA).
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

// 1 (default): packed fp32 pairs (FFMA2 ...) in the activation epilogues; 0: one lane per instruction (A/B measurements)
#ifndef NUNERF_PACKED_EPI
#define NUNERF_PACKED_EPI 1
#endif

namespace nunerf {

constexpr int CH_BLOCK_BYTES = 128 * 64 * 2;     // one activation K-block: 128 rows x 64 bf16
constexpr int CH_WROWS = 256;                    // output rows per weight block (128 = half layers: measured slower)
constexpr int CH_WSTAGE_BYTES = CH_WROWS * 64 * 2;   // one weight block: <= CH_WROWS output rows x 64 bf16 of K
constexpr int CH_EPI_WARPS = 16;
constexpr int CH_THREADS = 32 * (2 + CH_EPI_WARPS);
constexpr int CH_MAXL = NUNERF_CHAIN_MAX_LAYERS;

struct ChainLayer {
  int N;            // MMA N (multiple of 16, 16..256)
  int n_real;       // produced columns >= n_real are replaced (zeros, or the PE side block when cat_pe)
  int kb0, nkb;     // input K-blocks [kb0, kb0 + nkb) of the tile's four activation blocks
  int act;          // 0 none, 1 relu, 2 softplus(beta = 100)
  int cat_pe;       // columns [n_real, 256) <- PE-6 columns [0, 256 - n_real) of the point (SDF skip concat)
  int to_x;         // write the activation back to X blocks 0..3 (a next layer or a TMA store consumes it)
  int store_chunks; // > 0: TMA-store that many 64-column chunks of the activation through out_map
  int w_box_bytes;  // bytes of one weight TMA box: 128 B x min(128, N) rows
  int hot;          // plain 256-wide hidden layer with a specialised epilogue (ch_hot16 KIND 1..3), 0 = generic path
  const float* bias;
  // aux epilogues of the SDF network's reverse passes (hot kinds 4..6, plain 256-wide layers without bias):
  //   4: y = acc . s                      s = 1 - exp(-100 aux1)   (softplus'(z) from the stored activation)
  //   5: y = acc . s ;  e_out = acc . aux2 . 100 (1 - s)           (reverse-over-reverse glue, field.cu sdf_bwd2_ew)
  //   6: y = acc . s + aux2
  const __nv_bfloat16* aux1; int ld_aux1;
  const __nv_bfloat16* aux2; int ld_aux2;
  __nv_bfloat16* e_out; int ld_e;
  int mask_perm;    // hot layers only: mask words in THREAD order -- byte j*8 + c*2 holds the 16 bits of columns
                    // c*64 + j*16 .. +15 (one 8-byte access per thread and tile instead of four 2-byte ones)
  uint8_t* mask_out; int ldmask_out;       // optional 1-bit (x > 0) mask, 32 bytes per row
  const uint8_t* mask_in; int ldmask_in;   // optional 1-bit multiplicative mask
  float* out32; int ldo32; int n32;        // optional fp32 copy of the first n32 columns
  // chain_ts.cu only: the kept activation goes to global memory straight from the epilogue registers (no TMA store)
  __nv_bfloat16* store; int ld_store;      // optional bf16 copy of columns [0, store_cols)
  int store_cols;
  int keep;                                // the activation becomes the next layer's A operand (written to TMEM)
};

struct ChainParams {
  CUtensorMap in_map;
  CUtensorMap w_map[CH_MAXL];
  CUtensorMap out_map[CH_MAXL];
  ChainLayer layer[CH_MAXL];
  int n_layers, M, num_tiles;
  int in_mode;           // 0: X0 = rows of a bf16 matrix (TMA, into the activation blocks), 1: X0 = PE-6 of pts
  int in_blocks;         // K-blocks of the TMA input (1..4)
  int in_release_layer;  // (unused)
  int w_stages;
  const float* pts;
  int dbg_flags;         // timing experiments only (NUNERF_CHAIN_DEBUG): 1 = skip the TMEM load, 2 = skip the smem store,
                         // 8 = nanosleep in the epilogue, 16 = skip the MMAs, 32 = skip the weight loads, 64 = no bias loads
  long long* dbg;        // optional timeline buffer (NUNERF_CHAIN_TIMELINE): [2][256] clock64 stamps of CTA 0
  int role_hi;           // chain.cu: 1 = producer / MMA issuer are the two HIGHEST warps of the CTA (the scheduler arbitrates
                         // highest warp id first, B300_MICROARCH.md), 0 = warps 0 / 1
  int x_blocks;          // chain.cu: activation K-blocks in shared memory (8; 10 when the input is 320 columns wide)
  int epi_wait;          // chain.cu (NUNERF_CHAIN_EPIWAIT, default 1): bit 0 = every epilogue warp waits for the accumulator
                         // barrier on its own (0: the 16 warps start each tile together, one waiter + bar.sync -- measured
                         // 6 % slower on the fused SDF query); bit 1 = the issuer parks on x_done instead of spinning (no
                         // effect); bit 3 (default on) = ReLU-backward layers issue their four accumulator loads together
                         // and wait once (their ~40 instructions per chunk cannot hide a TMEM round trip; -2.5 % on the
                         // plain chains).  The same burst for ReLU-forward layers costs 3 registers + spills and slowed
                         // every path by 3-4 %: not kept; bit 5 = DISABLE the early hand-back of a tile's input blocks in
                         // chains that end in a narrow head (on by default: predictor forward chains +5 %)
};


__device__ __forceinline__ void ch_unpack16(const uint4& a, const uint4& b, float* f) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[2 * i] = bf16lo_to_f(w[i]); f[2 * i + 1] = bf16hi_to_f(w[i]); }
}


// PE-6 value of column pc (0..38) of [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] (field.py:14-61), rounded to bf16
// exactly as the tile-input writer does (the SDF skip concat re-creates these columns instead of keeping a copy in
// shared memory: the 32 KB go to a third weight stage, which the L2 latency of the weight stream needs).
__device__ __forceinline__ float ch_pe_col(const float* x, int pc) {
  if (pc < 3) return __bfloat162float(__float2bfloat16_rn(x[pc]));
  const int t = (pc - 3) / 3, c = (pc - 3) % 3;
  float sn, co;
  sincosf(x[c] * (float)(1 << (t >> 1)), &sn, &co);
  return __bfloat162float(__float2bfloat16_rn((t & 1) ? co : sn));
}


// chain_ts.cu
int chain_ts_launch(ChainParams& P, cudaStream_t stream);

}  // namespace nunerf
