// chain.cu -- fused multi-layer perceptron chains on tcgen05 tensor cores (sm_100a).
//
// Two 128-row tiles of points per CTA run through a whole chain of dense layers without their activations ever
// leaving the SM: the bf16 activation of layer l sits in shared memory (four 128x64 K-blocks per tile, 128B swizzle)
// as the A operand of layer l+1, each tile owns one 256-column fp32 accumulator in TMEM, and only the weights stream
// in (from L2, by TMA, through a ring of 32 KB K-blocks).  The two tiles run in PING-PONG: while the tensor core
// computes layer l of one tile, the 16 epilogue warps activate layer l (or l-1) of the other one, so neither the
// tensor pipe nor the epilogue warps wait for each other in steady state (measured timelines of the earlier
// single-tile version, tools/chain_timeline.py: the MMA warp trails the epilogue by ~3000 cycles per layer).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0       TMA producer (weight K-blocks; the tile's input rows when they come from global memory)
//   warp 1       tcgen05.mma issuer (one lane) + TMEM allocation + TMA stores of activations that must be kept
//   warps 2..17  epilogue: warp (q, j) owns TMEM lane quarter q and the 16 columns j of every 64-column chunk
//
// Programs built on it (host side, below):
//   nunerf_sdf_infer   SDFNetwork.sdf (field.py:133-152): PE-6 in-kernel, 8 Softplus(beta=100) layers with the skip
//                      concat at layer 4, 1-row sdf head -- the no-grad SDF queries of sample_ray (ZT:598, :563),
//                      extract_fields (field.py:1286-1307) and the occlusion probes (field.py:524-554)
//   nunerf_mlp_chain   generic ReLU / Softplus chains with optional per-layer stores (predictors, NeRF++)
#include "chain_common.cuh"

// timeline stamps inside the MMA / producer loops cost issue slots on a starved warp: compile them in only on demand
#ifndef NUNERF_CHAIN_TIMELINE_DETAIL
#define NUNERF_CHAIN_TIMELINE_DETAIL 0
#endif

namespace nunerf {

// Timing experiments (NUNERF_CHAIN_DEBUG bits) and the per-layer timeline stamps are compiled in only with
// -DNUNERF_CHAIN_DEBUG_BUILD=1 (`make EXTRA=-DNUNERF_CHAIN_DEBUG_BUILD=1`, tools/chain_timeline.py / tools/gpu_d.sh): in the
// product build the flag tests and the keep-alive compares they need cost ~4 % of the epilogue instructions.
#ifndef NUNERF_CHAIN_DEBUG_BUILD
#define NUNERF_CHAIN_DEBUG_BUILD 0
#endif
#define CH_DBGF(flags) (NUNERF_CHAIN_DEBUG_BUILD ? (flags) : 0)
#define CH_TL(ptr) (NUNERF_CHAIN_DEBUG_BUILD ? (ptr) : (long long*)nullptr)

// Hot epilogue of a plain 256-wide hidden layer -> bf16 -> shared memory: 16 columns of one row.
//   KIND 1: bias + Softplus(beta = 100)              (SDF network)
//   KIND 2: bias + ReLU, emits the 16 (x > 0) bits    (predictor / NeRF++ forward)
//   KIND 3: multiply by the 16 mask bits `mbits`      (ReLU backward: dZ_l = (dZ_{l+1} W_{l+1}) . [z_l > 0])
template <int KIND>
__device__ __forceinline__ void ch_hot16(const uint32_t* v, const float4* b, uint8_t* dst, int j, uint32_t sw,
                                         uint32_t* obits, uint32_t mbits, int dbg_flags) {
  float x[16];
  if (KIND == 3) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = ((mbits >> i) & 1u) ? __uint_as_float(v[i]) : 0.0f;
  } else {
#if NUNERF_PACKED_EPI
    if (KIND == 1) {
      // bias + Softplus two lanes per instruction (FADD2 / FMUL2 / FFMA2): the fma pipe bounds this epilogue
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        x[4 * i] = __uint_as_float(v[4 * i]); x[4 * i + 1] = __uint_as_float(v[4 * i + 1]);
        x[4 * i + 2] = __uint_as_float(v[4 * i + 2]); x[4 * i + 3] = __uint_as_float(v[4 * i + 3]);
        softplus100_x2(x[4 * i], x[4 * i + 1], b[i].x, b[i].y);
        softplus100_x2(x[4 * i + 2], x[4 * i + 3], b[i].z, b[i].w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), f2_pack(b[i].x, b[i].y)), x[4 * i], x[4 * i + 1]);
        f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), f2_pack(b[i].z, b[i].w)), x[4 * i + 2], x[4 * i + 3]);
      }
    }
#else
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[4 * i] = __uint_as_float(v[4 * i]) + b[i].x;
      x[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b[i].y;
      x[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b[i].z;
      x[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b[i].w;
    }
#endif
    if (KIND == 1) {
#if !NUNERF_PACKED_EPI
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = softplus100(x[i]);
#endif
    } else {
      uint32_t ob = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        ob |= (x[i] > 0.0f ? 1u : 0u) << i;
        x[i] = fmaxf(x[i], 0.0f);
      }
      *obits = ob;
    }
  }
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
  if (CH_DBGF(dbg_flags) & 2) {
    if (h[0] == 0x12345678u && h[5] == 0x9abcdef0u) *obits = h[3];   // keep the math alive without the store
    return;
  }
  ptx::st_shared_v4(dst + (((uint32_t)(2 * j) ^ sw) << 4), make_uint4(h[0], h[1], h[2], h[3]));
  ptx::st_shared_v4(dst + (((uint32_t)(2 * j + 1) ^ sw) << 4), make_uint4(h[4], h[5], h[6], h[7]));
}

// Aux epilogues (kinds 4..6, see ChainLayer): 16 columns of one row; a0/a1 = the 16 bf16 of aux1, b0/b1 of aux2.
template <int KIND>
__device__ __forceinline__ void ch_aux16(uint32_t taddr, const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1,
                                         uint8_t* dst, int j, uint32_t sw, __nv_bfloat16* e_dst, bool row_ok) {
  uint32_t v[16];
  ptx::tmem_ld16(taddr, v);
  float av[16], s[16];
  ch_unpack16(a0, a1, av);
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = 1.0f - __expf(-100.0f * av[i]);
  float bv[16];
  if (KIND >= 5) ch_unpack16(b0, b1, bv);
  ptx::tmem_ld_wait();
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float acc = __uint_as_float(v[i]);
    x[i] = acc * s[i];
    if (KIND == 6) x[i] += bv[i];
    if (KIND == 5) bv[i] = acc * bv[i] * 100.0f * (1.0f - s[i]);
    if (!row_ok) x[i] = 0.0f;
  }
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
  ptx::st_shared_v4(dst + (((uint32_t)(2 * j) ^ sw) << 4), make_uint4(h[0], h[1], h[2], h[3]));
  ptx::st_shared_v4(dst + (((uint32_t)(2 * j + 1) ^ sw) << 4), make_uint4(h[4], h[5], h[6], h[7]));
  if (KIND == 5 && row_ok) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = pack_bf16x2(bv[2 * i], bv[2 * i + 1]);
    ptx::st_global_v8(e_dst, e);
  }
}

__device__ __forceinline__ void ch_tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

// Block ids inside sX: tile t in {0,1} owns activation blocks t*4 + (0..3).  The tile input (TMA rows or the in-kernel
// PE) is written into those same blocks: they are free between the last layer of one pair and the first of the next.
// A 320-column input (the [feature | p] rows of the material predictors) has a fifth K-block per tile: blocks 8 + t, which
// exist only in that configuration (ChainParams::x_blocks = 10, one weight stage less).
__device__ __forceinline__ int ch_block(int t, int id) { return id < 4 ? t * 4 + id : 8 + t; }

// PAIR = 1: one CTA per pair of tiles, cta_group::1 MMAs (M = 128).
// PAIR = 2: clusters of two CTAs, cta_group::2 MMAs (M = 256 = one tile of each CTA): every CTA keeps its own tiles,
//           accumulators, epilogue warps and activation stores, but holds only HALF of each weight block (N/2 rows) --
//           the pair's tensor cores share the two halves, so per SM the operand reads drop from 12 KB to 8 KB per
//           instruction and the weight bytes written into shared memory halve (the bound of PAIR = 1).
//           Only the leader CTA (cluster rank 0) issues MMAs and commits; the peer's warp 1 replays the same schedule as a
//           proxy: it waits for ITS local conditions (epilogue done, input / weight half landed, stores drained) and
//           arrives on the leader's peer_go[stage] barrier, which the leader waits for before each K-block.
// KM = the epilogue kinds this instantiation carries (bit 0: Softplus layers, bit 1: ReLU-forward, bit 2: ReLU-backward,
//      bit 3: the SDF reverse-pass kinds 4..6; the generic path is always present).  A launch picks the instantiation that
//      matches the kinds of its layers: smaller code (fewer instruction-cache misses and branches) and a register
//      allocation that is not the maximum over every epilogue of the file.
template <int PAIR, int KM>
__global__ void __launch_bounds__(CH_THREADS, 1) mlp_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int n_xblocks = p.x_blocks;
  uint8_t* sX = smem;                                         // activation K-blocks of both tiles
  uint8_t* sW = sX + (size_t)n_xblocks * CH_BLOCK_BYTES;      // weight ring
  const int wstage_bytes = CH_WSTAGE_BYTES / PAIR;             // PAIR = 2: a stage holds this CTA's half of a weight block
  uint64_t* bars = (uint64_t*)(sW + (size_t)p.w_stages * wstage_bytes);
  uint64_t* w_full = bars;                  // [w_stages]
  uint64_t* w_empty = bars + 8;             // [w_stages]
  uint64_t* x_done = bars + 16;             // [2]  epilogue of tile t finished (X written, accumulator drained)
  uint64_t* t_full = bars + 18;             // [2]  accumulator of tile t complete
  uint64_t* in_full = bars + 20;            // [2]  tile input is in shared memory
  uint64_t* in_empty = bars + 22;           // [2]  X blocks of tile t may be refilled (TMA input mode)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 24);
  uint64_t* peer_go = bars + 32;            // [w_stages]  PAIR = 2, leader only: the peer CTA is ready for this K-block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // service warps: the two lowest warp ids, or (role_hi) the two highest -- the issue arbiter serves the highest warp id of
  // a scheduler first, so the single MMA-issuing thread is no longer starved by the epilogue warps it shares it with
  const int w_prod = p.role_hi ? CH_EPI_WARPS : 0, w_mma = p.role_hi ? CH_EPI_WARPS + 1 : 1;
  const int num_pairs = (p.num_tiles + 1) >> 1;
  // PAIR = 2: both CTAs of a cluster run the same number of tile pairs (a surplus pair is all-masked rows)
  const uint32_t cl_rank = PAIR == 2 ? ptx::cluster_ctarank() : 0u;
  const int num_iters = (num_pairs + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == w_prod && lane == 0) {
    for (int l = 0; l < p.n_layers; ++l) {
      ptx::prefetch_tmap(&p.w_map[l]);
      if (p.layer[l].store_chunks > 0) ptx::prefetch_tmap(&p.out_map[l]);
    }
    if (p.in_mode == 0) ptx::prefetch_tmap(&p.in_map);
    for (int i = 0; i < p.w_stages; ++i) {
      ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); ptx::mbar_init(&peer_go[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&x_done[i], CH_EPI_WARPS);
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&in_full[i], p.in_mode == 0 ? 1 : CH_EPI_WARPS);
      ptx::mbar_init(&in_empty[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == w_mma) {
    if (PAIR == 2) { ptx::tmem_alloc2(tmem_ptr, 512u); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(tmem_ptr, 512u); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR == 2) ptx::cluster_sync();      // the peer must not signal barriers that are not initialised yet
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == w_prod) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_iters; ++it) {
        const int pair = (int)blockIdx.x + it * (int)gridDim.x;
        if (p.in_mode == 0) {
          for (int t = 0; t < 2; ++t) {
            ptx::mbar_wait_parked(&in_empty[t], (uint32_t)(it & 1) ^ 1);
            ptx::mbar_expect_tx(&in_full[t], (uint32_t)(p.in_blocks * CH_BLOCK_BYTES));
            for (int b = 0; b < p.in_blocks; ++b)      // rows beyond M (or a missing second tile) are zero filled
              ptx::tma_load_2d(sX + (size_t)ch_block(t, b) * CH_BLOCK_BYTES, &p.in_map, &in_full[t], b * 64,
                               (pair * 2 + t) * 128);
          }
        }
        for (int l = 0; l < p.n_layers; ++l) {
          const int nkb = p.layer[l].nkb;
          const int row0 = (int)cl_rank * (p.layer[l].N / PAIR);     // PAIR = 2: this CTA's half of the weight rows
          for (int t = 0; t < 2; ++t)
            for (int kb = 0; kb < nkb; ++kb) {
              // the stage is free once the MMAs that read it have completed (tcgen05.commit, multicast to the pair)
              ptx::mbar_wait_parked(&w_empty[stage], phase ^ 1);
              if (CH_TL(p.dbg) && blockIdx.x == 0 && it == 1 && l == 5) p.dbg[400 + (t * 4 + kb)] = clock64();
              // box = 64 K-columns x N / PAIR rows
              if (CH_DBGF(p.dbg_flags) & 32) {            // timing experiment: no weight traffic at all (stale weights)
                ptx::mbar_arrive(&w_full[stage]);
              } else {
                ptx::mbar_expect_tx(&w_full[stage], (uint32_t)p.layer[l].w_box_bytes);
                ptx::tma_load_2d(sW + (size_t)stage * wstage_bytes, &p.w_map[l], &w_full[stage], kb * 64, row0);
              }
              if (++stage == p.w_stages) { stage = 0; phase ^= 1; }
            }
        }
      }
    }
  } else if (warp == w_mma) {
    // ================================================================ MMA issuer
    // The WHOLE warp runs this control flow (every value is warp-uniform: descriptors and counters live in uniform
    // registers, a handful of instructions per MMA); only MMA / commit / TMA-store are issued by lane 0.
    // Ping-pong: (layer l, tile 0), (layer l, tile 1), (layer l+1, tile 0), ... -- while the tensor core works on one
    // tile the 16 epilogue warps activate the other one.
    // the issuing lane of this warp, chosen by elect.sync: ptxas then keeps the MMA operands in uniform registers; from a
    // `lane == 0` branch it wraps every tcgen05.mma in an elect / broadcast loop (200+ cycles per instruction)
    const bool leader = ptx::elect_one();
    const bool cta_leader = PAIR == 1 || cl_rank == 0;      // the CTA that issues the pair's MMAs
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    long long g = 0;             // layer counter per tile (both tiles advance together): x_done / t_full phase g
    const uint64_t desc_hi = ptx::smem_desc(0, 16, 1024);
    const uint32_t sx_base = ptx::smem_u32(sX), sw_base = ptx::smem_u32(sW);
    const bool early_in = PAIR == 1 && p.in_mode == 0 && p.layer[p.n_layers - 1].store_chunks == 0 &&
                          !p.layer[p.n_layers - 1].to_x && !(p.epi_wait & 32);
    for (it = 0; it < num_iters; ++it) {
      const int pair = (int)blockIdx.x + it * (int)gridDim.x;
      for (int l = 0; l < p.n_layers; ++l, ++g) {
        const ChainLayer& L = p.layer[l];
        const int prev_store = l > 0 ? p.layer[l - 1].store_chunks : 0;
        for (int t = 0; t < 2; ++t) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(t * 256);
          // the previous layer's epilogue of this tile has drained the accumulator and written the activation
#if NUNERF_CHAIN_TIMELINE_DETAIL
          if (CH_TL(p.dbg) && leader && blockIdx.x == 0 && it == 1 && l == 5) p.dbg[450 + t] = clock64();
#endif
          if (g > 0) {
            if (p.epi_wait & 2) ptx::mbar_wait_parked(&x_done[t], (uint32_t)((g - 1) & 1));
            else ptx::mbar_wait(&x_done[t], (uint32_t)((g - 1) & 1));
          }
          if (l == 0) ptx::mbar_wait(&in_full[t], (uint32_t)(it & 1));
          ptx::tc_fence_after();
#if NUNERF_CHAIN_TIMELINE_DETAIL
          if (CH_TL(p.dbg) && leader && blockIdx.x == 0 && it == 1 && l == 5) p.dbg[420 + t] = clock64();
#endif
          bool stores_pending = false;
          if (prev_store > 0) {
            if (leader) {
              for (int c = 0; c < prev_store; ++c)
                ch_tma_store_2d(&p.out_map[l - 1], sX + (size_t)(t * 4 + c) * CH_BLOCK_BYTES, c * 64, (pair * 2 + t) * 128);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            stores_pending = true;
          }
          const uint32_t idesc = ptx::idesc_bf16(128 * PAIR, L.N, 0, 0);
          for (int kb = 0; kb < L.nkb; ++kb) {
            const int blk = ch_block(t, L.kb0 + kb);
            ptx::mbar_wait(&w_full[stage], phase);
            ptx::tc_fence_after();
#if NUNERF_CHAIN_TIMELINE_DETAIL
            if (CH_TL(p.dbg) && leader && blockIdx.x == 0 && it == 1 && l == 5) p.dbg[430 + (t * 4 + kb) * 2] = clock64();
#endif
            if (PAIR == 2 && !cta_leader) {
              // proxy of the peer CTA: everything the pair's MMA on this K-block needs from THIS CTA is in place (the
              // activation tile, the drained accumulator, this half of the weight block; before the last K-block also
              // the TMA stores that still read the activation blocks) -> tell the leader
              if (leader) {
                if (kb == L.nkb - 1 && stores_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                ptx::mbar_arrive_remote(&peer_go[stage], 0u);
              }
            } else {
              if (PAIR == 2) { ptx::mbar_wait_cluster(&peer_go[stage], phase); ptx::tc_fence_after(); }
              const uint64_t ad0 = desc_hi | (uint64_t)(((sx_base + (uint32_t)blk * CH_BLOCK_BYTES) >> 4) & 0x3fff);
              const uint64_t bd0 = desc_hi | (uint64_t)(((sw_base + (uint32_t)stage * wstage_bytes) >> 4) & 0x3fff);
              if (leader) {
                if (!(CH_DBGF(p.dbg_flags) & 16)) {         // (16: timing experiment without the MMAs themselves)
                  if (PAIR == 2) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)    // +32 bytes per K = 16 step: +2 in the (address >> 4) field
                      ptx::umma_bf16_2cta(d_tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, (uint32_t)(kb | k));
                  } else {
                    // the four K steps as one statement on 32-bit descriptor words (ptx.cuh)
                    ptx::umma_bf16_ss_x4(d_tmem, (uint32_t)ad0, (uint32_t)bd0, (uint32_t)(desc_hi >> 32), idesc, (uint32_t)kb);
                  }
                }
                if (PAIR == 2) ptx::tc_commit2_mc(&w_empty[stage], (uint16_t)3);
                else ptx::tc_commit(&w_empty[stage]);
#if NUNERF_CHAIN_TIMELINE_DETAIL
                if (CH_TL(p.dbg) && blockIdx.x == 0 && it == 1 && l == 5) p.dbg[430 + (t * 4 + kb) * 2 + 1] = clock64();
#endif
              }
            }
            if (++stage == p.w_stages) { stage = 0; phase ^= 1; }
          }
          if (leader && cta_leader) {
            // this layer's epilogue overwrites the X blocks: outstanding TMA stores must have read them
            if (stores_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (PAIR == 2) ptx::tc_commit2_mc(&t_full[t], (uint16_t)3);
            else ptx::tc_commit(&t_full[t]);
            // a chain that ends in a narrow head (nothing written back to X, no store of it): the tile's X blocks are free
            // as soon as these MMAs have read them -- the producer fetches the next pair's input while the head's epilogue
            // still runs, instead of after it (short chains lost ~10 % to that bubble)
            if (early_in && l == p.n_layers - 1) ptx::tc_commit(&in_empty[t]);
            if (CH_TL(p.dbg) && blockIdx.x == 0 && it == 1 && l < 12) p.dbg[(l * 2 + t) * 2] = clock64();
          }
          __syncwarp();
        }
      }
      // end of the pair: store a kept last layer, then hand the X blocks back to the producer (TMA input mode)
      const int last_store = p.layer[p.n_layers - 1].store_chunks;
      if (!early_in && (last_store > 0 || p.in_mode == 0)) {
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(&x_done[t], (uint32_t)((g - 1) & 1));
          if (leader) {
            if (last_store > 0) {
              for (int c = 0; c < last_store; ++c)
                ch_tma_store_2d(&p.out_map[p.n_layers - 1], sX + (size_t)(t * 4 + c) * CH_BLOCK_BYTES, c * 64,
                                (pair * 2 + t) * 128);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (p.in_mode == 0) ptx::mbar_arrive(&in_empty[t]);
          }
          __syncwarp();
        }
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ================================================================ epilogue warps
    const int ew = p.role_hi ? warp : warp - 2;
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int j = ew >> 2;       // its 16 columns inside every 64-column chunk
    const int r = q * 32 + lane; // row inside the tile
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    long long g = 0;
    for (int it = 0; it < num_iters; ++it) {
      const int pair = (int)blockIdx.x + it * (int)gridDim.x;
      if (p.in_mode == 1) {
        // ---- tile inputs: PE-6 of the point, 39 columns + zero padding to 64, into the tile's input block
        for (int t = 0; t < 2; ++t) {
          if (j == 0) {
            const long long row = ((long long)pair * 2 + t) * 128 + r;
            float x[3] = {0.f, 0.f, 0.f};
            if (row < p.M) { x[0] = p.pts[3 * row]; x[1] = p.pts[3 * row + 1]; x[2] = p.pts[3 * row + 2]; }
            float pe[64];
#pragma unroll
            for (int c = 0; c < 3; ++c) pe[c] = x[c];
#pragma unroll
            for (int k = 0; k < 6; ++k)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float sn, co;
                sincosf(x[c] * (float)(1 << k), &sn, &co);
                pe[3 + 6 * k + c] = sn;
                pe[6 + 6 * k + c] = co;
              }
#pragma unroll
            for (int c = 39; c < 64; ++c) pe[c] = 0.f;
            uint8_t* dst = sX + (size_t)(t * 4) * CH_BLOCK_BYTES + row_off;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              uint4 v = make_uint4(pack_bf16x2(pe[8 * ch], pe[8 * ch + 1]), pack_bf16x2(pe[8 * ch + 2], pe[8 * ch + 3]),
                                   pack_bf16x2(pe[8 * ch + 4], pe[8 * ch + 5]), pack_bf16x2(pe[8 * ch + 6], pe[8 * ch + 7]));
              ptx::st_shared_v4(dst + (((uint32_t)ch ^ sw) << 4), v);
            }
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&in_full[t]);
        }
      }
      for (int l = 0; l < p.n_layers; ++l, ++g) {
        const ChainLayer& L = p.layer[l];
        for (int t = 0; t < 2; ++t) {
          const long long row = ((long long)pair * 2 + t) * 128 + r;
          const bool row_ok = row < p.M;
          // ReLU-backward layers: this thread's 4 x 16 mask bits, fetched while the tensor core is still busy
          unsigned long long hot_mask = 0ull;
          if ((KM & 4) && L.hot == 3) {
            if (L.mask_perm) {
              hot_mask = __ldg(reinterpret_cast<const unsigned long long*>(L.mask_in + (row_ok ? row : 0) * L.ldmask_in + 8 * j));
            } else {
              const uint8_t* mrow = L.mask_in + (row_ok ? row : 0) * L.ldmask_in + 2 * j;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                hot_mask |= (unsigned long long)__ldg(reinterpret_cast<const uint16_t*>(mrow + 8 * c)) << (16 * c);
            }
            if (!row_ok) hot_mask = 0ull;
          }
          unsigned long long out_mask = 0ull;
          uint4 aux_a[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)}, aux_b[2] = {aux_a[0], aux_a[0]};
          const long long aux_row = row_ok ? row : 0;
          if ((KM & 8) && L.hot >= 4) {
            // (32-byte loads: each lane reads one whole sector of its row with one request)
            ptx::ld_global_nc_v8(L.aux1 + aux_row * L.ld_aux1 + j * 16, aux_a[0], aux_a[1]);
            if (L.hot >= 5) ptx::ld_global_nc_v8(L.aux2 + aux_row * L.ld_aux2 + j * 16, aux_b[0], aux_b[1]);
          }
          if ((p.epi_wait & 1) == 0) {
            // one spinning waiter per CTA; the other 15 epilogue warps sleep on a hardware barrier (all 16 warps start a
            // tile's epilogue together)
            if (ew == 0 && lane == 0) ptx::mbar_wait(&t_full[t], (uint32_t)(g & 1));
            asm volatile("bar.sync 1, 512;" ::: "memory");
          } else {
            // every warp waits for the accumulator on its own (parked try_wait): no rendezvous between the warps, so a
            // warp that is done with one tile moves on to the other while its neighbours still work
            if (lane == 0) ptx::mbar_wait_parked(&t_full[t], (uint32_t)(g & 1));
            __syncwarp();
          }
          ptx::tc_fence_after();
          if (CH_TL(p.dbg) && blockIdx.x == 0 && it == 1 && l < 12 && lane == 0 && ew == 0) p.dbg[256 + (l * 2 + t) * 2] = clock64();
          uint8_t* xt = sX + (size_t)(t * 4) * CH_BLOCK_BYTES;
          if ((KM & 8) && L.hot >= 4) {
            // aux operands stream from global memory one 64-column chunk ahead of their use (32 bytes per thread, row
            // and chunk: whole sectors); the first chunk was requested before the accumulator wait
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int c0 = c * 64 + j * 16;
              const uint4 a0 = aux_a[0], a1 = aux_a[1], b0 = aux_b[0], b1 = aux_b[1];
              if (c < 3) {
                ptx::ld_global_nc_v8(L.aux1 + aux_row * L.ld_aux1 + c0 + 64, aux_a[0], aux_a[1]);
                if (L.hot >= 5) ptx::ld_global_nc_v8(L.aux2 + aux_row * L.ld_aux2 + c0 + 64, aux_b[0], aux_b[1]);
              }
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + c0);
              uint8_t* dst = xt + (size_t)c * CH_BLOCK_BYTES + row_off;
              if (L.hot == 4) ch_aux16<4>(taddr, a0, a1, b0, b1, dst, j, sw, nullptr, row_ok);
              else if (L.hot == 5) ch_aux16<5>(taddr, a0, a1, b0, b1, dst, j, sw, L.e_out + aux_row * L.ld_e + c0, row_ok);
              else ch_aux16<6>(taddr, a0, a1, b0, b1, dst, j, sw, nullptr, row_ok);
            }
          } else if ((KM & 4) && L.hot == 3 && (p.epi_wait & 8)) {
            // ---- ReLU-backward layers do ~40 instructions per 16-column chunk: too little to hide a TMEM round trip
            // behind, so all four chunk loads of the row go out together and are awaited once
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + j * 16);
            uint32_t v0[16], v1[16], v2[16], v3[16];
            ptx::tmem_ld16(taddr0, v0);
            ptx::tmem_ld16(taddr0 + 64u, v1);
            ptx::tmem_ld16(taddr0 + 128u, v2);
            ptx::tmem_ld16(taddr0 + 192u, v3);
            ptx::tmem_ld_wait();
            uint32_t ob = 0;
            const float4 b[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f),
                                 make_float4(0.f, 0.f, 0.f, 0.f)};
            ch_hot16<3>(v0, b, xt + row_off, j, sw, &ob, (uint32_t)hot_mask & 0xffffu, p.dbg_flags);
            ch_hot16<3>(v1, b, xt + (size_t)CH_BLOCK_BYTES + row_off, j, sw, &ob, (uint32_t)(hot_mask >> 16) & 0xffffu, p.dbg_flags);
            ch_hot16<3>(v2, b, xt + (size_t)2 * CH_BLOCK_BYTES + row_off, j, sw, &ob, (uint32_t)(hot_mask >> 32) & 0xffffu, p.dbg_flags);
            ch_hot16<3>(v3, b, xt + (size_t)3 * CH_BLOCK_BYTES + row_off, j, sw, &ob, (uint32_t)(hot_mask >> 48) & 0xffffu, p.dbg_flags);
          } else if ((KM & 7) && L.hot >= 1 && L.hot <= 3) {
            // ---- plain 256-wide hidden layers: the accumulator read of chunk c + 1 is in flight while chunk c is activated
            // (one exposed TMEM round trip per tile instead of four)
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + j * 16);
            uint32_t va[16], vb[16];
            ptx::tmem_ld16(taddr0, va);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int c0 = c * 64 + j * 16;     // first of this thread's 16 columns
              uint32_t* cur = (c & 1) ? vb : va;
              // this chunk's 16 biases are requested BEFORE the accumulator wait: their (L1) latency hides behind it
              // instead of being exposed at the first add (ncu: 7 % of all stall samples sat there)
              float4 b[4];
              if (L.hot != 3 && !(CH_DBGF(p.dbg_flags) & 64)) {
                {
                  uint4 q0, q1, q2, q3;        // 2 x 32-byte loads (c0 is a multiple of 16 floats)
                  ptx::ld_global_nc_v8(L.bias + c0, q0, q1);
                  ptx::ld_global_nc_v8(L.bias + c0 + 8, q2, q3);
                  b[0] = *reinterpret_cast<float4*>(&q0); b[1] = *reinterpret_cast<float4*>(&q1);
                  b[2] = *reinterpret_cast<float4*>(&q2); b[3] = *reinterpret_cast<float4*>(&q3);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              ptx::tmem_ld_wait();
              if (c < 3) ptx::tmem_ld16(taddr0 + (uint32_t)((c + 1) * 64), (c & 1) ? va : vb);
              uint32_t ob = 0;
              uint8_t* dst = xt + (size_t)c * CH_BLOCK_BYTES + row_off;
              if ((KM & 1) && L.hot == 1) ch_hot16<1>(cur, b, dst, j, sw, &ob, 0u, p.dbg_flags);
              else if ((KM & 2) && L.hot == 2) {
                ch_hot16<2>(cur, b, dst, j, sw, &ob, 0u, p.dbg_flags);
                if (L.mask_perm) out_mask |= (unsigned long long)ob << (16 * c);
                else if (L.mask_out && row_ok)
                  *reinterpret_cast<uint16_t*>(L.mask_out + row * L.ldmask_out + (c0 >> 3)) = (uint16_t)ob;
              } else if (KM & 4) {
                ch_hot16<3>(cur, b, dst, j, sw, &ob, (uint32_t)(hot_mask >> (16 * c)) & 0xffffu, p.dbg_flags);
              }
            }
          } else
          for (int c = 0; c < 4; ++c) {
            const int c0 = c * 64 + j * 16;     // first of this thread's 16 columns
            const bool in_acc = c0 < L.N;       // columns the MMA produced
            // columns that must be (re)written in shared memory: the produced ones, the concatenated PE columns,
            // and the zero tail of a partially produced K-block (the next layer reads whole 64-column blocks)
            const bool in_x = L.to_x && c0 < (L.cat_pe ? 256 : ((L.N + 63) & ~63));
            if (in_acc || in_x) {
              float x[16];
              if (in_acc) {
                uint32_t v[16];
                ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + c0), v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
                if (L.bias) {
                  const float4* b4 = reinterpret_cast<const float4*>(L.bias + c0);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float4 b = __ldg(b4 + i);
                    x[4 * i] += b.x; x[4 * i + 1] += b.y; x[4 * i + 2] += b.z; x[4 * i + 3] += b.w;
                  }
                }
                if (L.act == 1) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], 0.0f);
                } else if (L.act == 2) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) x[i] = softplus100(x[i]);
                }
                if (L.mask_in) {
                  const uint32_t mb =
                      *reinterpret_cast<const uint16_t*>(L.mask_in + (row_ok ? row : 0) * L.ldmask_in + (c0 >> 3));
#pragma unroll
                  for (int i = 0; i < 16; ++i) x[i] = ((mb >> i) & 1u) ? x[i] : 0.0f;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = 0.0f;
              }
              if (c0 + 16 > L.n_real) {
                if (L.cat_pe) {
                  float px[3] = {0.f, 0.f, 0.f};
                  if (row_ok) { px[0] = p.pts[3 * row]; px[1] = p.pts[3 * row + 1]; px[2] = p.pts[3 * row + 2]; }
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const int col = c0 + i;
                    if (col >= L.n_real) x[i] = ch_pe_col(px, col - L.n_real);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (c0 + i >= L.n_real) x[i] = 0.0f;
                }
              }
              if (in_acc && L.mask_out) {
                uint32_t ob = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) ob |= (x[i] > 0.0f ? 1u : 0u) << i;
                if (row_ok) *reinterpret_cast<uint16_t*>(L.mask_out + row * L.ldmask_out + (c0 >> 3)) = (uint16_t)ob;
              }
              if (in_x) {
                uint8_t* dst = xt + (size_t)c * CH_BLOCK_BYTES + row_off;
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
                ptx::st_shared_v4(dst + (((uint32_t)(2 * j) ^ sw) << 4), make_uint4(h[0], h[1], h[2], h[3]));
                ptx::st_shared_v4(dst + (((uint32_t)(2 * j + 1) ^ sw) << 4), make_uint4(h[4], h[5], h[6], h[7]));
              }
              if (in_acc && L.out32 && row_ok) {
                float* o = L.out32 + row * L.ldo32 + c0;
                if (c0 + 16 <= L.n32 && (L.ldo32 & 7) == 0 && (((uintptr_t)L.out32) & 31) == 0) {
                  // two 32-byte stores: whole sectors of the thread's own row
                  uint32_t w[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(x[i]);
                  ptx::st_global_v8(o, w);
                  ptx::st_global_v8(o + 8, w + 8);
                } else if (c0 + 16 <= L.n32 && (L.ldo32 & 3) == 0) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    reinterpret_cast<float4*>(o)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (c0 + i < L.n32) o[i] = x[i];
                }
              }
            }
          }
          if ((KM & 2) && L.hot == 2 && L.mask_perm && L.mask_out && row_ok)
            *reinterpret_cast<unsigned long long*>(L.mask_out + row * L.ldmask_out + 8 * j) = out_mask;
          // publish: generic-proxy writes of this warp -> visible to the tensor core / TMA (async proxy)
          ptx::fence_proxy_async();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&x_done[t]);
          if (CH_TL(p.dbg) && blockIdx.x == 0 && it == 1 && l < 12 && lane == 0 && ew == 0) p.dbg[256 + (l * 2 + t) * 2 + 1] = clock64();
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR == 2) ptx::cluster_sync();      // no CTA leaves while the pair's MMAs / commits may still touch it
  if (warp == w_mma) {
    if (PAIR == 2) ptx::tmem_dealloc2(tmem_base, 512u);
    else ptx::tmem_dealloc(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------- host side
// 1 (default): cta_group::1 kernel.  2 (NUNERF_CHAIN_PAIR=2): CTA pairs sharing the weight operand (cta_group::2) --
// functionally identical (tests/test_engine_gpu.py runs both) but measured SLOWER on B200 for these 256-wide layers
// (fused SDF query 496 vs 570 TFLOP/s, gpurun_out/s22/s24: the pair's instruction takes ~300 cycles against ~210 for
// one CTA's, and every tile-layer waits for the slower of two epilogues), so it stays an opt-in experiment.
static int chain_pair() { return env_int("NUNERF_CHAIN_PAIR", 1) == 2 ? 2 : 1; }

template <int PAIR, int KM>
static int chain_launch_t(ChainParams& P, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_chain_kernel<PAIR, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail("chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e), -2);
    if (PAIR == 2) {
      e = cudaFuncSetAttribute(mlp_chain_kernel<PAIR, KM>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
      (void)e; (void)cudaGetLastError();
    }
    configured = true;
  }
  NUNERF_REQUIRE(P.in_mode == 1 || (P.in_blocks >= 1 && P.in_blocks <= (PAIR == 1 ? 5 : 4)),
                 "chain: TMA input must be 1..4 K-blocks (5 with the single-CTA kernel)");
  P.x_blocks = (P.in_mode == 0 && P.in_blocks == 5) ? 10 : 8;
  const size_t xbytes = (size_t)P.x_blocks * CH_BLOCK_BYTES;
  const size_t fixed = 1024 + 512;
  const int stage_bytes = CH_WSTAGE_BYTES / PAIR;
  int stages = (int)((227 * 1024 - fixed - xbytes) / stage_bytes);
  if (stages > 8) stages = 8;
  NUNERF_REQUIRE(stages >= 2, "chain: input too wide for shared memory");
  { const int s_env = env_int("NUNERF_CHAIN_STAGES", 0); if (s_env >= 2 && s_env < stages) stages = s_env; }   // experiments
  for (int l = 0; l < P.n_layers; ++l) {
    P.layer[l].w_box_bytes = 128 * (P.layer[l].N / PAIR);
  }
  P.w_stages = stages;
  P.num_tiles = cdiv(P.M, 128);
  const size_t smem = fixed + xbytes + (size_t)stages * stage_bytes;
  const int num_pairs = (P.num_tiles + 1) / 2;
  const int cluster = PAIR;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // persistent grid = the number of clusters that are co-resident (GPC sizes need not be multiples of the cluster)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cfg.gridDim = dim3(num_sms() / cluster * cluster);
    int n = 0;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&n, mlp_chain_kernel<PAIR, KM>, &cfg);
    max_clusters = (qe == cudaSuccess && n > 0) ? n : num_sms() / cluster;
    (void)cudaGetLastError();
  }
  int grid = max_clusters * cluster;
  if (grid > num_sms()) grid = num_sms() / cluster * cluster;
  const int need = (num_pairs + cluster - 1) / cluster * cluster;
  if (grid > need) grid = need;
  { const int g_env = env_int("NUNERF_CHAIN_GRID", 0); if (g_env >= cluster && g_env < grid) grid = g_env / cluster * cluster; }
  cfg.gridDim = dim3(grid);
  cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_chain_kernel<PAIR, KM>, P);
  if (e != cudaSuccess) return fail("chain: cudaLaunchKernelEx: %s", cudaGetErrorString(e), -2);
  NUNERF_CHECK_LAUNCH("mlp_chain_kernel");
  return 0;
}

// NUNERF_CHAIN_IMPL: "ss" (default) = this file's kernel (two tiles per CTA in ping-pong, activations in shared memory);
// "ts" = activations in tensor memory, TS-mode MMAs, one tile per CTA (chain_ts.cu).  Both pass the same parity tests; on
// B200 the SS kernel is the faster one for every chain of the step (DESIGN.md 4 has the measurements and the reasons).
static bool chain_use_ts() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NUNERF_CHAIN_IMPL"); v = (e && e[0] == 't') ? 1 : 0; }
  return v == 1 && chain_pair() == 1;
}

static int chain_launch(ChainParams& P, cudaStream_t stream) {
  if (chain_use_ts()) return chain_ts_launch(P, stream);
  P.role_hi = env_int("NUNERF_CHAIN_HIPRIO", 1);
  P.epi_wait = env_int("NUNERF_CHAIN_EPIWAIT", 9);
  if (chain_pair() == 2) return chain_launch_t<2, 15>(P, stream);
  // the kinds of this chain's layers pick the instantiation (NUNERF_CHAIN_SPECIALISE=0: always the full kernel)
  int km = 0;
  for (int l = 0; l < P.n_layers; ++l) {
    const int h = P.layer[l].hot;
    km |= h == 1 ? 1 : h == 2 ? 2 : h == 3 ? 4 : h >= 4 ? 8 : 0;
  }
  if (!env_int("NUNERF_CHAIN_SPECIALISE", 1)) km = 15;
  switch (km) {
    case 0: case 1: return chain_launch_t<1, 1>(P, stream);
    case 2: return chain_launch_t<1, 2>(P, stream);
    case 4: return chain_launch_t<1, 4>(P, stream);
    case 8: return chain_launch_t<1, 8>(P, stream);
    default: return chain_launch_t<1, 15>(P, stream);
  }
}

}  // namespace nunerf

using namespace nunerf;

extern "C" int nunerf_sdf_infer(const nunerf_sdf_infer_t* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(a && a->pts && a->sdf && a->M > 0, "sdf_infer: bad arguments");
  static const int Ns[9] = {256, 256, 256, 224, 256, 256, 256, 256, 16};
  static const int Ks[9] = {64, 256, 256, 256, 256, 256, 256, 256, 256};
  ChainParams P;
  memset(&P, 0, sizeof(P));
  P.n_layers = 9; P.M = a->M; P.in_mode = 1; P.in_blocks = 1; P.in_release_layer = 0; P.pts = a->pts;
  for (int l = 0; l < 9; ++l) {
    NUNERF_REQUIRE(a->w[l] && a->ldw[l] >= Ks[l] && a->ldw[l] % 8 == 0, "sdf_infer: bad weight operand");
    if (int r = make_map(&P.w_map[l], a->w[l], Ns[l], Ks[l], a->ldw[l], 64, Ns[l] / chain_pair())) return r;
    ChainLayer& L = P.layer[l];
    L.N = Ns[l]; L.n_real = Ns[l];
    L.kb0 = 0; L.nkb = Ks[l] / 64;
    L.act = l < 8 ? 2 : 0; L.to_x = l < 8 ? 1 : 0; L.keep = L.to_x;
    L.bias = a->bias[l];
    L.hot = (l < 8 && l != 3) ? 1 : 0;             // KIND 1: bias + softplus
  }
  if (env_int("NUNERF_CHAIN_DEBUG_RELU", 0))           // timing experiment only: ReLU epilogue instead of Softplus
    for (int l = 0; l < 8; ++l) { P.layer[l].act = 1; if (P.layer[l].hot) P.layer[l].hot = 2; }
  P.layer[3].n_real = 217; P.layer[3].cat_pe = 1;      // x <- cat([x, PE]) / sqrt(2) (the scale lives in lin4's weights)
  P.layer[8].n_real = 16; P.layer[8].out32 = a->sdf; P.layer[8].ldo32 = a->ld_sdf; P.layer[8].n32 = 1;
  P.dbg = (long long*)a->timeline;
  P.dbg_flags = env_int("NUNERF_CHAIN_DEBUG", 0);
  return chain_launch(P, stream);
}

extern "C" int nunerf_mlp_chain(const nunerf_mlp_chain_t* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(a && (a->x || a->pts) && a->M > 0, "mlp_chain: bad arguments");
  NUNERF_REQUIRE(a->n_layers >= 1 && a->n_layers <= CH_MAXL, "mlp_chain: 1..NUNERF_CHAIN_MAX_LAYERS layers");
  ChainParams P;
  memset(&P, 0, sizeof(P));
  P.n_layers = a->n_layers; P.M = a->M;
  int width;                         // valid (written) columns of the activation blocks
  if (a->x) {
    const int k0_max = (chain_use_ts() || chain_pair() == 2) ? 256 : 320;
    NUNERF_REQUIRE(a->K0 >= 64 && a->K0 % 64 == 0 && a->K0 <= k0_max && a->ldx % 8 == 0 && a->ldx >= a->K0,
                   "mlp_chain: input must be 64..256 columns (320 with the default kernel; multiple of 64), pitch % 8 == 0");
    P.in_mode = 0; P.in_blocks = a->K0 / 64;
    if (int r = make_map(&P.in_map, a->x, a->M, a->K0, a->ldx, 64, 128)) return r;
    width = a->K0;
  } else {
    P.in_mode = 1; P.in_blocks = 1; width = 64;      // x_0 = PE-6 of pts (39 columns, zero padded to 64), in-kernel
  }
  P.pts = a->pts;
  for (int l = 0; l < a->n_layers; ++l) {
    const nunerf_chain_layer_t& s = a->layer[l];
    ChainLayer& L = P.layer[l];
    NUNERF_REQUIRE(s.w && s.N >= 16 && s.N % 16 == 0 && s.N <= 256, "mlp_chain: N must be a multiple of 16, <= 256");
    NUNERF_REQUIRE(s.K >= 64 && s.K % 64 == 0 && s.K <= width, "mlp_chain: K exceeds the activation produced so far");
    NUNERF_REQUIRE(s.ldw % 8 == 0 && s.ldw >= s.K, "mlp_chain: bad weight pitch");
    NUNERF_REQUIRE(s.act >= 0 && s.act <= 2, "mlp_chain: act must be 0 (none), 1 (relu) or 2 (softplus 100)");
    if (int r = make_map(&P.w_map[l], s.w, s.N, s.K, s.ldw, 64, s.N / chain_pair())) return r;
    L.N = s.N; L.n_real = s.n_real > 0 ? s.n_real : s.N;
    L.kb0 = 0; L.nkb = s.K / 64;
    L.act = s.act; L.bias = s.bias;
    L.cat_pe = s.cat_pe;
    if (s.cat_pe) NUNERF_REQUIRE(a->pts && L.n_real < 256 && 256 - L.n_real <= 39, "mlp_chain: cat_pe needs pts and n_real >= 217");
    L.aux1 = (const __nv_bfloat16*)s.aux1; L.ld_aux1 = s.ld_aux1;
    L.aux2 = (const __nv_bfloat16*)s.aux2; L.ld_aux2 = s.ld_aux2;
    L.e_out = (__nv_bfloat16*)s.e_out; L.ld_e = s.ld_e;
    L.mask_out = s.mask_out; L.ldmask_out = s.ldmask_out;
    L.mask_in = s.mask_in; L.ldmask_in = s.ldmask_in;
    L.out32 = s.out32; L.ldo32 = s.ldo32; L.n32 = s.n32;
    const bool last = l + 1 == a->n_layers;
    L.to_x = (s.keep || s.store) ? 1 : 0;
    L.keep = (L.to_x && !last) ? 1 : 0;
    if (s.store) {
      NUNERF_REQUIRE(s.ld_store % 8 == 0 && s.ld_store >= s.N, "mlp_chain: bad store pitch");
      const int wide = s.cat_pe ? 256 : (s.N + 63) / 64 * 64;
      const int cols = wide <= s.ld_store ? wide : s.N;
      L.store_chunks = wide / 64;
      L.store = (__nv_bfloat16*)s.store; L.ld_store = s.ld_store; L.store_cols = cols;
      if (!chain_use_ts())
        if (int r = make_map(&P.out_map[l], s.store, a->M, cols, s.ld_store, 64, 128)) return r;
    }
    if (L.to_x) width = s.cat_pe ? 256 : (s.N + 63) / 64 * 64;   // a layer that keeps nothing leaves the previous activation
    // specialised epilogues
    const bool plain = s.N == 256 && L.n_real == 256 && L.to_x && !s.out32;
    if (plain && s.bias && s.act == 2 && !s.mask_in && !s.mask_out) L.hot = 1;
    else if (plain && s.bias && s.act == 1 && !s.mask_in) L.hot = 2;
    else if (plain && !s.bias && s.act == 0 && s.mask_in && !s.mask_out) L.hot = 3;
    if (s.aux_mode) {
      NUNERF_REQUIRE(s.aux_mode >= 4 && s.aux_mode <= 6, "mlp_chain: aux_mode must be 0 or 4..6");
      NUNERF_REQUIRE(plain && !s.bias && s.act == 0 && !s.mask_in && !s.mask_out, "mlp_chain: aux layers are plain 256-wide");
      NUNERF_REQUIRE(s.aux1 && s.ld_aux1 % 16 == 0 && ((uintptr_t)s.aux1 & 31) == 0, "mlp_chain: aux1 must be 32-byte aligned (pitch % 16)");
      if (s.aux_mode >= 5)
        NUNERF_REQUIRE(s.aux2 && s.ld_aux2 % 16 == 0 && ((uintptr_t)s.aux2 & 31) == 0, "mlp_chain: aux2 must be 32-byte aligned (pitch % 16)");
      if (s.aux_mode == 5)
        NUNERF_REQUIRE(s.e_out && s.ld_e % 16 == 0 && ((uintptr_t)s.e_out & 31) == 0, "mlp_chain: e_out must be 32-byte aligned (pitch % 16)");
      L.hot = s.aux_mode;
    }
    L.mask_perm = s.mask_perm;
    if (s.mask_perm)
      NUNERF_REQUIRE((L.hot == 2 || L.hot == 3) && ((s.mask_in ? s.ldmask_in : s.ldmask_out) % 8 == 0) &&
                         (((uintptr_t)(s.mask_in ? (const void*)s.mask_in : (const void*)s.mask_out)) & 7) == 0,
                     "mlp_chain: mask_perm needs a plain 256-wide ReLU layer and 8-byte aligned mask rows");
    if (s.mask_in) NUNERF_REQUIRE(s.ldmask_in >= s.N / 8, "mlp_chain: mask_in pitch");
    if (s.mask_out) NUNERF_REQUIRE(s.ldmask_out >= s.N / 8, "mlp_chain: mask_out pitch");
  }
  // timing experiments only (results are wrong with these set): drop the activation stores / the mask writes
  if (env_int("NUNERF_CHAIN_NOSTORE", 0)) for (int l = 0; l < a->n_layers; ++l) P.layer[l].store_chunks = 0;
  if (env_int("NUNERF_CHAIN_NOMASK", 0)) for (int l = 0; l < a->n_layers; ++l) P.layer[l].mask_out = nullptr;
  P.dbg = (long long*)a->timeline;
  P.dbg_flags = 0;
  return chain_launch(P, stream);
}
