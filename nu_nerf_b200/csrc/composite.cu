// composite.cu -- render_core geometry, inner/outer compaction and the fused compositing kernels
// (ZT:725-785).  One warp per ray; sample s lives in lane (s & 31), block (s >> 5): all [R,S] accesses are
// coalesced.  Backward recomputes the transmittance with the same scan instead of reading a stored copy.
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace nunerf {

constexpr int WPB = 4;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float scan_mul32(float x, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float t = __shfl_up_sync(FULL, x, off);
    if (lane >= off) x *= t;
  }
  return x;
}
__device__ __forceinline__ float scan_add32_rev(float x, int lane) {
  // inclusive suffix sum within the warp
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float t = __shfl_down_sync(FULL, x, off);
    if (lane + off < 32) x += t;
  }
  return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
  return x;
}

// ---- geometry pass 1: dists, mid points, per-ray inner count (ZT:730-736)
__global__ void geometry_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z,
                                int R, int S, float* dists, float* pts, int32_t* ray_inner) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (r >= R) return;
  const float ox = o[3 * r], oy = o[3 * r + 1], oz = o[3 * r + 2];
  const float dx = d[3 * r], dy = d[3 * r + 1], dz = d[3 * r + 2];
  const float* zr = z + (long long)r * S;
  int cnt = 0;
  for (int s = lane; s < S + (32 - (S & 31)) % 32; s += 32) {
    bool ok = s < S;
    float z0 = ok ? zr[s] : 0.f;
    float dist;
    if (ok) {
      if (s + 1 < S) dist = __fsub_rn(zr[s + 1], z0);
      else dist = __fsub_rn(z0, zr[s - 1]);
      float zm = __fadd_rn(z0, __fmul_rn(dist, 0.5f));
      float px = __fadd_rn(ox, __fmul_rn(dx, zm)), py = __fadd_rn(oy, __fmul_rn(dy, zm)),
            pz = __fadd_rn(oz, __fmul_rn(dz, zm));
      long long i = (long long)r * S + s;
      if (dists) dists[i] = dist;
      if (pts) { pts[3 * i] = px; pts[3 * i + 1] = py; pts[3 * i + 2] = pz; }
      float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
      ok = nrm <= 1.0f;
    }
    cnt += __popc(__ballot_sync(FULL, ok));
  }
  if (lane == 0) ray_inner[r] = cnt;
}

// ---- geometry pass 2: exclusive scan of the per-ray inner counts, one block per 1024 rays (coalesced loads, shuffle
//      scans).  ray_off[r] = offset inside the ray's block; the block total overwrites the block's own first count
//      (ray_inner[b * 1024], already consumed) and pass 3 adds the totals of the preceding blocks.
constexpr int SCAN_BLOCK = 1024;
__global__ void __launch_bounds__(SCAN_BLOCK) ray_scan_kernel(int32_t* __restrict__ ray_inner, int R, int32_t* ray_off) {
  __shared__ int s_warp[32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int r = blockIdx.x * SCAN_BLOCK + t;
  const int v = r < R ? ray_inner[r] : 0;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int u = __shfl_up_sync(FULL, incl, off);
    if (lane >= off) incl += u;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = s_warp[lane], wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int u = __shfl_up_sync(FULL, wi, off);
      if (lane >= off) wi += u;
    }
    s_warp[lane] = wi - w;                       // exclusive prefix of the warp totals
    if (lane == 31) ray_inner[blockIdx.x * SCAN_BLOCK] = wi;     // block total (every count of the block was read above)
  }
  __syncthreads();
  if (r < R) ray_off[r] = s_warp[wid] + incl - v;
}

// inner samples of all rays before block b of the scan (<= 64 blocks for R <= 65536: one warp-wide sum per warp)
__device__ __forceinline__ int scan_block_base(const int32_t* __restrict__ ray_inner, int b, int lane) {
  int acc = 0;
  for (int i = lane; i < b; i += 32) acc += ray_inner[i * SCAN_BLOCK];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
  return acc;
}

// The compaction map in its compact form, RAY_MAP ints per ray: [0] index of the ray's first inner sample in the inner
// list, [1] of its first outer sample in the outer list, [2 + k] bit mask of the inner samples among samples 32k..32k+31.
// (Boolean-mask indexing keeps the row-major order, so a ray's inner / outer samples are two contiguous runs.)  The
// compositing kernels rebuild a sample's list index from it with two popcounts instead of reading a 4-byte slot per
// sample.
constexpr int RAY_MAP = 10;

// F.normalize(dirs) (ZT:740): d / max(||d||, 1e-12) with explicitly rounded operations (both compaction kernels share it)
__device__ __forceinline__ void normalized_dir(const float* __restrict__ d, int r, float& dx, float& dy, float& dz) {
  dx = d[3 * r]; dy = d[3 * r + 1]; dz = d[3 * r + 2];
  const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);
  dx = __fdiv_rn(dx, nrm); dy = __fdiv_rn(dy, nrm); dz = __fdiv_rn(dz, nrm);
}

// ---- geometry pass 3: compact gather of both sets (row-major mask order, as points[inner_mask]); the per-sample
//      geometry is recomputed from z with the arithmetic of pass 1 (bit-identical) instead of being re-read
__global__ void compact_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z,
                               const int32_t* __restrict__ ray_inner, const int32_t* __restrict__ ray_off,
                               int32_t* __restrict__ counts, int R, int S, int32_t* slot, int32_t* ray_map,
                               float* pts_in, float* dists_in, float* dirs_in, int32_t* id_in, float* pts_out,
                               float* dists_out, float* dirs_out, int32_t* id_out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (r >= R) return;
  const float ox = o[3 * r], oy = o[3 * r + 1], oz = o[3 * r + 2];
  const float rdx = d[3 * r], rdy = d[3 * r + 1], rdz = d[3 * r + 2];
  float dx, dy, dz;
  normalized_dir(d, r, dx, dy, dz);
  int in_run = ray_off[r] + scan_block_base(ray_inner, r / SCAN_BLOCK, lane);
  int out_run = r * S - in_run;
  if (r == 0) {                                 // totals for the host (the one sync of the step reads them)
    const int n_in = scan_block_base(ray_inner, (R + SCAN_BLOCK - 1) / SCAN_BLOCK, lane);
    if (lane == 0) { counts[0] = n_in; counts[1] = R * S - n_in; }
  }
  if (ray_map && lane == 0) { ray_map[(long long)r * RAY_MAP] = in_run; ray_map[(long long)r * RAY_MAP + 1] = out_run; }
  const float* zr = z + (long long)r * S;
  for (int s0 = 0; s0 < S; s0 += 32) {
    int s = s0 + lane;
    bool ok = s < S;
    long long i = (long long)r * S + (ok ? s : 0);
    float px = 0.f, py = 0.f, pz = 0.f, dist = 0.f;
    if (ok) {
      const float z0 = zr[s];
      if (s + 1 < S) dist = __fsub_rn(zr[s + 1], z0);
      else dist = __fsub_rn(z0, zr[s - 1]);
      const float zm = __fadd_rn(z0, __fmul_rn(dist, 0.5f));
      px = __fadd_rn(ox, __fmul_rn(rdx, zm)); py = __fadd_rn(oy, __fmul_rn(rdy, zm)); pz = __fadd_rn(oz, __fmul_rn(rdz, zm));
    }
    float pn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
    bool inner = ok && pn <= 1.0f;
    bool outer = ok && !inner;
    unsigned bi = __ballot_sync(FULL, inner), bo = __ballot_sync(FULL, outer);
    unsigned lt = (1u << lane) - 1u;
    if (ray_map && lane == 0) ray_map[(long long)r * RAY_MAP + 2 + (s0 >> 5)] = (int32_t)bi;
    if (inner) {
      int k = in_run + __popc(bi & lt);
      if (slot) slot[i] = k;
      pts_in[3 * k] = px; pts_in[3 * k + 1] = py; pts_in[3 * k + 2] = pz;
      dists_in[k] = dist;
      dirs_in[3 * k] = dx; dirs_in[3 * k + 1] = dy; dirs_in[3 * k + 2] = dz;
      if (id_in) id_in[k] = (int32_t)i;
    } else if (outer) {
      int k = out_run + __popc(bo & lt);
      if (slot) slot[i] = -1 - k;
      pts_out[3 * k] = px; pts_out[3 * k + 1] = py; pts_out[3 * k + 2] = pz;
      dists_out[k] = dist;
      dirs_out[3 * k] = dx; dirs_out[3 * k + 1] = dy; dirs_out[3 * k + 2] = dz;
      if (id_out) id_out[k] = (int32_t)i;
    }
    in_run += __popc(bi);
    out_run += __popc(bo);
  }
}

// list index of sample (block k, lane) from the ray map: >= 0 inner, < 0 -> -1 - outer index (the slot convention)
__device__ __forceinline__ int slot_from_map(int in_off, int out_off, unsigned mask, int inner_before, int k, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  const int in_lt = inner_before + __popc(mask & lt);
  if ((mask >> lane) & 1u) return in_off + in_lt;
  return -1 - (out_off + (k * 32 + lane) - in_lt);
}

// ---- compositing
struct Sample { float a, ab, c0, c1, c2; };

__device__ __forceinline__ Sample fetch(const float* __restrict__ a_in, const float* __restrict__ c_in,
                                        const float* __restrict__ a_out, const float* __restrict__ c_out, int sl) {
  Sample s;
  if (sl >= 0) {
    s.a = a_in[sl]; s.ab = 0.f;
    s.c0 = c_in[3 * sl]; s.c1 = c_in[3 * sl + 1]; s.c2 = c_in[3 * sl + 2];
  } else {
    int k = -1 - sl;
    s.a = a_out[k]; s.ab = s.a;
    s.c0 = c_out[3 * k]; s.c1 = c_out[3 * k + 1]; s.c2 = c_out[3 * k + 2];
  }
  return s;
}

// NBLK = ceil(S / 32) blocks of 32 samples (5 for the product's S = 160, 8 for the general S <= 256).  All loads of a
// ray are issued before the first scan: the prefix products are a dependent chain (shuffles + the carry between blocks),
// so without this the kernel alternates between one block's memory latency and one block's scan latency.
template <int NBLK>
__global__ void __launch_bounds__(32 * WPB) composite_fwd_kernel(const float* __restrict__ a_in, const float* __restrict__ c_in,
                                     const float* __restrict__ a_out, const float* __restrict__ c_out,
                                     const int32_t* __restrict__ slot, const int32_t* __restrict__ ray_map, int R, int S,
                                     int is_nerf, float* rgb, float* rgb_raw, float* acc, float* rgb_b, float* weights) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (r >= R) return;
  int sl[NBLK];
  if (ray_map) {
    const int32_t* rm = ray_map + (long long)r * RAY_MAP;
    const int in_off = __ldg(rm), out_off = __ldg(rm + 1);
    int before = 0;
#pragma unroll
    for (int k = 0; k < NBLK; ++k) {
      const unsigned m = (unsigned)__ldg(rm + 2 + k);
      sl[k] = slot_from_map(in_off, out_off, m, before, k, lane);
      before += __popc(m);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NBLK; ++k) {
      const int s = k * 32 + lane;
      sl[k] = s < S ? __ldg(slot + (long long)r * S + s) : 0;
    }
  }
  Sample sm[NBLK];
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    const int s = k * 32 + lane;
    sm[k] = Sample{0.f, 0.f, 0.f, 0.f, 0.f};
    if (s < S) sm[k] = fetch(a_in, c_in, a_out, c_out, sl[k]);
  }
  float carry = 1.f, carry_b = 1.f;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, sa = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    const int s = k * 32 + lane;
    const bool ok = s < S;
    float v = ok ? (1.0f - sm[k].a + 1e-7f) : 1.0f;
    float vb = ok ? (1.0f - sm[k].ab + 1e-7f) : 1.0f;
    float incl = scan_mul32(v, lane), incl_b = scan_mul32(vb, lane);
    float ex = __shfl_up_sync(FULL, incl, 1), exb = __shfl_up_sync(FULL, incl_b, 1);
    if (lane == 0) { ex = 1.f; exb = 1.f; }
    float w = sm[k].a * (carry * ex), wb = sm[k].ab * (carry_b * exb);
    carry *= __shfl_sync(FULL, incl, 31);
    carry_b *= __shfl_sync(FULL, incl_b, 31);
    if (ok && weights) weights[(long long)r * S + s] = w;
    s0 += w * sm[k].c0; s1 += w * sm[k].c1; s2 += w * sm[k].c2; sa += w;
    b0 += wb * sm[k].c0; b1 += wb * sm[k].c1; b2 += wb * sm[k].c2;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); sa = warp_sum(sa);
  b0 = warp_sum(b0); b1 = warp_sum(b1); b2 = warp_sum(b2);
  if (lane == 0) {
    if (is_nerf) { float bg = 1.0f - sa; s0 += bg; s1 += bg; s2 += bg; }
    rgb_raw[3 * r] = s0; rgb_raw[3 * r + 1] = s1; rgb_raw[3 * r + 2] = s2;
    rgb[3 * r] = fminf(fmaxf(s0, 0.f), 1.f); rgb[3 * r + 1] = fminf(fmaxf(s1, 0.f), 1.f);
    rgb[3 * r + 2] = fminf(fmaxf(s2, 0.f), 1.f);
    acc[r] = sa;
    rgb_b[3 * r] = b0; rgb_b[3 * r + 1] = b1; rgb_b[3 * r + 2] = b2;
  }
}

// d_alpha_i = g_i T_i - (sum_{k>i} g_k w_k) / (1 - a_i + eps),  g_k = d_rgb . c_k + d_acc - [is_nerf] sum(d_rgb)
// NBLK = ceil(S / 32) blocks of 32 samples are kept in registers between the forward and the reverse pass: the kernel is
// instantiated for the product's S = 160 (5 blocks) and for the general S <= 256 (8 blocks) so that the common case does
// not pay the register footprint (and the occupancy) of the largest one.
template <int NBLK>
__global__ void __launch_bounds__(32 * WPB) composite_bwd_kernel(const float* __restrict__ a_in, const float* __restrict__ c_in,
                                     const float* __restrict__ a_out, const float* __restrict__ c_out,
                                     const int32_t* __restrict__ slot, const int32_t* __restrict__ ray_map, int R, int S,
                                     int is_nerf, const float* __restrict__ rgb_raw, const float* __restrict__ d_rgb,
                                     const float* __restrict__ d_acc, const float* __restrict__ d_rgb_b,
                                     float* d_a_in, float* d_c_in, float* d_a_out, float* d_c_out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (r >= R) return;
  float g0 = d_rgb ? d_rgb[3 * r] : 0.f, g1 = d_rgb ? d_rgb[3 * r + 1] : 0.f, g2 = d_rgb ? d_rgb[3 * r + 2] : 0.f;
  // clamp(color, 0, 1) passes the gradient on the closed interval
  float q0 = rgb_raw[3 * r], q1 = rgb_raw[3 * r + 1], q2 = rgb_raw[3 * r + 2];
  if (!(q0 >= 0.f && q0 <= 1.f)) g0 = 0.f;
  if (!(q1 >= 0.f && q1 <= 1.f)) g1 = 0.f;
  if (!(q2 >= 0.f && q2 <= 1.f)) g2 = 0.f;
  float ga = (d_acc ? d_acc[r] : 0.f) - (is_nerf ? (g0 + g1 + g2) : 0.f);
  float h0 = d_rgb_b ? d_rgb_b[3 * r] : 0.f, h1 = d_rgb_b ? d_rgb_b[3 * r + 1] : 0.f,
        h2 = d_rgb_b ? d_rgb_b[3 * r + 2] : 0.f;
  const int nblk = (S + 31) >> 5;
  // pass 1 (forward): T per sample, kept in registers (S <= 256)
  float T[NBLK], Tb[NBLK], gw[NBLK], gwb[NBLK], va[NBLK], vab[NBLK], gk[NBLK], gkb[NBLK];
  int sl[NBLK];
  float carry = 1.f, carry_b = 1.f;
  // all loads of the ray first (see composite_fwd_kernel)
  if (ray_map) {
    const int32_t* rm = ray_map + (long long)r * RAY_MAP;
    const int in_off = __ldg(rm), out_off = __ldg(rm + 1);
    int before = 0;
#pragma unroll
    for (int k = 0; k < NBLK; ++k) {
      const unsigned m = (unsigned)__ldg(rm + 2 + k);
      sl[k] = slot_from_map(in_off, out_off, m, before, k, lane);
      before += __popc(m);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NBLK; ++k) {
      const int s = k * 32 + lane;
      sl[k] = (k < nblk && s < S) ? __ldg(slot + (long long)r * S + s) : 0;
    }
  }
  Sample smp[NBLK];
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    const int s = k * 32 + lane;
    smp[k] = Sample{0.f, 0.f, 0.f, 0.f, 0.f};
    if (k < nblk && s < S) smp[k] = fetch(a_in, c_in, a_out, c_out, sl[k]);
  }
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    if (k < nblk) {
      int s = k * 32 + lane;
      bool ok = s < S;
      const Sample sm = smp[k];
      float v = ok ? (1.0f - sm.a + 1e-7f) : 1.0f, vb = ok ? (1.0f - sm.ab + 1e-7f) : 1.0f;
      float incl = scan_mul32(v, lane), incl_b = scan_mul32(vb, lane);
      float ex = __shfl_up_sync(FULL, incl, 1), exb = __shfl_up_sync(FULL, incl_b, 1);
      if (lane == 0) { ex = 1.f; exb = 1.f; }
      T[k] = carry * ex; Tb[k] = carry_b * exb;
      carry *= __shfl_sync(FULL, incl, 31);
      carry_b *= __shfl_sync(FULL, incl_b, 31);
      va[k] = v; vab[k] = vb;
      float w = sm.a * T[k], wb = sm.ab * Tb[k];
      gk[k] = ok ? (g0 * sm.c0 + g1 * sm.c1 + g2 * sm.c2 + ga) : 0.f;
      gkb[k] = ok ? (h0 * sm.c0 + h1 * sm.c1 + h2 * sm.c2) : 0.f;
      gw[k] = gk[k] * w; gwb[k] = gkb[k] * wb;
      if (ok) {
        // colour gradient: both composites share the colour of outer samples
        float dc0 = w * g0 + wb * h0, dc1 = w * g1 + wb * h1, dc2 = w * g2 + wb * h2;
        if (sl[k] >= 0) { float* p = d_c_in + 3 * (long long)sl[k]; p[0] = dc0; p[1] = dc1; p[2] = dc2; }
        else { float* p = d_c_out + 3 * (long long)(-1 - sl[k]); p[0] = dc0; p[1] = dc1; p[2] = dc2; }
      }
    }
  }
  // pass 2 (reverse): exclusive suffix sums of g*w
  float suf = 0.f, suf_b = 0.f;
#pragma unroll
  for (int k = NBLK - 1; k >= 0; --k) {
    if (k < nblk) {
      int s = k * 32 + lane;
      bool ok = s < S;
      float incl = scan_add32_rev(gw[k], lane), incl_b = scan_add32_rev(gwb[k], lane);
      float ex = __shfl_down_sync(FULL, incl, 1), exb = __shfl_down_sync(FULL, incl_b, 1);
      if (lane == 31) { ex = 0.f; exb = 0.f; }
      float after = suf + ex, after_b = suf_b + exb;
      suf += __shfl_sync(FULL, incl, 0);
      suf_b += __shfl_sync(FULL, incl_b, 0);
      if (ok) {
        float da = gk[k] * T[k] - after / va[k];
        if (sl[k] >= 0) d_a_in[sl[k]] = da;
        else d_a_out[-1 - sl[k]] = da + (gkb[k] * Tb[k] - after_b / vab[k]);
      }
    }
  }
}

// =====================================================================================================================
// Staged compositing (the product path: S <= 160, per-ray compaction map).  One warp per ray.  The ray's four contiguous
// runs (inner alpha / colour, outer alpha / colour) are brought into shared memory as whole 16-byte chunks with
// cp.async (6 per lane; no per-sample address arithmetic), and every lane then owns SPL = 5 CONSECUTIVE samples: the
// transmittance is a 5-term product in registers plus ONE warp scan of the lane totals per composite (the lane-per-sample
// form needs one scan per block of 32 samples, five per ray), and the seven per-ray sums go through one transposing
// butterfly.  Stride-5 / stride-15 shared-memory reads are bank-conflict free.  The backward kernel overwrites the staged
// alpha / colour in place with their gradients and writes the chunks back (16-byte stores inside the run, scalar stores
// for the <= 3 elements of a boundary chunk, so nothing outside the ray's own run is written).  Loads may touch the
// <= 3 neighbouring elements that share a 16-byte chunk with the run (always inside the same 16-byte aligned buffer).
constexpr int CW = 8;            // warps (= rays in flight) per block
constexpr int SPL = 5;           // samples per lane
constexpr int A_ST = 176;        // staged floats per ray, alpha:  <= 43 chunks
constexpr int C_ST = 496;        // staged floats per ray, colour: <= 123 chunks

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// sums of 8 per-lane values over the warp with 9 shuffles: after the call lane l holds the total of value (l >> 2) & 7
// (identical in the four lanes that share it)
__device__ __forceinline__ float warp_sum8(float (&v)[8], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h4 ? v[i] : v[i + 4], keep = h4 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(FULL, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h3 ? v[i] : v[i + 2], keep = h3 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(FULL, send, 8);
  }
  {
    const float send = h2 ? v[0] : v[1], keep = h2 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(FULL, send, 4);
  }
  v[0] += __shfl_xor_sync(FULL, v[0], 2);
  v[0] += __shfl_xor_sync(FULL, v[0], 1);
  return v[0];
}

// the per-ray compaction map in registers (loaded one ray ahead)
struct RayMapRegs {
  int in_off, out_off;
  unsigned m[5];
};
__device__ __forceinline__ RayMapRegs load_ray_map(const int32_t* __restrict__ ray_map, int r) {
  RayMapRegs mp;
  const int32_t* rm = ray_map + (long long)r * RAY_MAP;
  mp.in_off = __ldg(rm); mp.out_off = __ldg(rm + 1);
#pragma unroll
  for (int k = 0; k < 5; ++k) mp.m[k] = (unsigned)__ldg(rm + 2 + k);
  return mp;
}

struct RayStage {
  int pa[SPL], pc[SPL];      // staged positions of the lane's samples (alpha / first colour component)
  unsigned inner;            // bit j: sample j of the lane is an inner sample
  unsigned valid;            // bit j: sample exists (s < S)
  int in_off, n_in, out_off, n_out;
  int nch_ai, nch_ao, nch_ci, nch_co;     // 16-byte chunks of the four runs
};

// run lengths / chunk counts of a ray.  Staged layout: [inner chunks | outer chunks]; a run keeps its global alignment
// modulo 16 B
__device__ __forceinline__ void ray_runs(const RayMapRegs& mp, int S, RayStage& st) {
  const int n_in = __popc(mp.m[0]) + __popc(mp.m[1]) + __popc(mp.m[2]) + __popc(mp.m[3]) + __popc(mp.m[4]);
  const int n_out = S - n_in;
  st.in_off = mp.in_off; st.out_off = mp.out_off; st.n_in = n_in; st.n_out = n_out;
  st.nch_ai = n_in ? ((mp.in_off & 3) + n_in + 3) >> 2 : 0;
  st.nch_ao = n_out ? ((mp.out_off & 3) + n_out + 3) >> 2 : 0;
  st.nch_ci = n_in ? (((3 * mp.in_off) & 3) + 3 * n_in + 3) >> 2 : 0;
  st.nch_co = n_out ? (((3 * mp.out_off) & 3) + 3 * n_out + 3) >> 2 : 0;
}

// start the chunk loads of a ray into its staging buffers (one commit group); returns the ray's run lengths / chunk
// counts packed into one word (n_in | nch_ai << 8 | nch_ci << 16 -- the outer ones follow from S) for decode_ray
__device__ __forceinline__ unsigned issue_ray_loads(const RayMapRegs& mp, int S, int lane, const float* __restrict__ a_in,
                                                const float* __restrict__ c_in, const float* __restrict__ a_out,
                                                const float* __restrict__ c_out, float* sa, float* sc) {
  RayStage st;
  ray_runs(mp, S, st);
  {
    const float* gi = a_in + (st.in_off & ~3);
    const float* go = a_out + (st.out_off & ~3) - 4 * st.nch_ai;
#pragma unroll
    for (int t = lane; t < 64; t += 32)
      if (t < st.nch_ai + st.nch_ao) cp_async16(sa + 4 * t, (t < st.nch_ai ? gi : go) + 4 * t);
  }
  {
    const float* gi = c_in + ((3ll * st.in_off) & ~3ll);
    const float* go = c_out + ((3ll * st.out_off) & ~3ll) - 4 * st.nch_ci;
#pragma unroll
    for (int t = lane; t < 128; t += 32)
      if (t < st.nch_ci + st.nch_co) cp_async16(sc + 4 * t, (t < st.nch_ci ? gi : go) + 4 * t);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  return (unsigned)st.n_in | ((unsigned)st.nch_ai << 8) | ((unsigned)st.nch_ci << 16);
}

// staged positions of the lane's samples (FULLS: S == 160, every sample of every lane exists)
template <bool FULLS>
__device__ __forceinline__ RayStage decode_ray(const RayMapRegs& mp, unsigned runs, int S, int lane) {
  RayStage st;
  st.in_off = mp.in_off; st.out_off = mp.out_off;
  st.n_in = (int)(runs & 255u); st.n_out = S - st.n_in;
  st.nch_ai = (int)((runs >> 8) & 255u); st.nch_ci = (int)(runs >> 16);
  st.nch_ao = st.n_out ? ((mp.out_off & 3) + st.n_out + 3) >> 2 : 0;
  st.nch_co = st.n_out ? (((3 * mp.out_off) & 3) + 3 * st.n_out + 3) >> 2 : 0;
  const int hi_a = mp.in_off & 3, ho_a = mp.out_off & 3, hi_c = (3 * mp.in_off) & 3, ho_c = (3 * mp.out_off) & 3;
  // the lane's 5 mask bits and the number of inner samples in front of them
  const int b0 = SPL * lane, w = b0 >> 5, sh = b0 & 31;
  unsigned lo = mp.m[0], hi = mp.m[1];
  int before = 0, run = 0;
#pragma unroll
  for (int k = 1; k < 5; ++k) {
    run += __popc(mp.m[k - 1]);
    if (w == k) { lo = mp.m[k]; hi = k < 4 ? mp.m[k < 4 ? k + 1 : 4] : 0u; before = run; }
  }
  st.inner = __funnelshift_r(lo, hi, sh) & 31u;
  before += __popc(lo & ((1u << sh) - 1u));
  st.valid = 0;
  const int qa = 4 * st.nch_ai + ho_a, qc = 4 * st.nch_ci + ho_c;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = b0 + j;
    const int rin = before + __popc(st.inner & ((1u << j) - 1u));
    const bool in = (st.inner >> j) & 1u;
    st.pa[j] = in ? hi_a + rin : qa + (s - rin);
    st.pc[j] = in ? hi_c + 3 * rin : qc + 3 * (s - rin);
    if (FULLS || s < S) st.valid |= 1u << j;
  }
  if (FULLS) st.valid = 31u;
  return st;
}

// write chunk t of a staged region back: elements [lo, hi) of the destination (global element indices relative to the
// 16-byte aligned base g) belong to the ray
__device__ __forceinline__ void unstage_chunk(float* __restrict__ g, int lo, int hi, const float* s, int t) {
  const float4 v = *reinterpret_cast<const float4*>(s + 4 * t);
  const int e = 4 * t;
  if (e >= lo && e + 4 <= hi) {
    *reinterpret_cast<float4*>(g + e) = v;
  } else {
    if (e >= lo && e < hi) g[e] = v.x;
    if (e + 1 >= lo && e + 1 < hi) g[e + 1] = v.y;
    if (e + 2 >= lo && e + 2 < hi) g[e + 2] = v.z;
    if (e + 3 >= lo && e + 3 < hi) g[e + 3] = v.w;
  }
}

// Both kernels are persistent: a warp walks over rays r = w, w + W, ... with two staging buffers -- the map of ray i+2
// is requested and the chunk loads of ray i+1 are in flight while ray i is computed, so neither the map's nor the data's
// memory latency is exposed.
template <bool FULLS, bool WEIGHTS>
__global__ void __launch_bounds__(32 * CW, 4) composite_fwd_staged_kernel(
    const float* __restrict__ a_in, const float* __restrict__ c_in, const float* __restrict__ a_out,
    const float* __restrict__ c_out, const int32_t* __restrict__ ray_map, int R, int S, int is_nerf, float* rgb,
    float* rgb_raw, float* acc, float* rgb_b, float* weights) {
  __shared__ __align__(16) float s_a[2][CW][A_ST];
  __shared__ __align__(16) float s_c[2][CW][C_ST];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int stride = gridDim.x * CW;
  int r = blockIdx.x * CW + wid;
  if (r >= R) return;
  RayMapRegs cur = load_ray_map(ray_map, r);
  unsigned cur_runs = issue_ray_loads(cur, S, lane, a_in, c_in, a_out, c_out, s_a[0][wid], s_c[0][wid]);
  RayMapRegs nxt = load_ray_map(ray_map, min(r + stride, R - 1));
  for (int buf = 0; r < R; r += stride, buf ^= 1) {
    float* sa = s_a[buf][wid];
    float* sc = s_c[buf][wid];
    cp_async_wait_all();
    __syncwarp();
    const RayStage st = decode_ray<FULLS>(cur, cur_runs, S, lane);
    float a[SPL], ab[SPL], c0[SPL], c1[SPL], c2[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const bool ok = (st.valid >> j) & 1u;
      a[j] = ok ? sa[st.pa[j]] : 0.f;
      c0[j] = ok ? sc[st.pc[j]] : 0.f;
      c1[j] = ok ? sc[st.pc[j] + 1] : 0.f;
      c2[j] = ok ? sc[st.pc[j] + 2] : 0.f;
      ab[j] = ((st.inner >> j) & 1u) ? 0.f : a[j];
    }
    // next ray: its map arrived during the previous iteration; start its loads, request the map after it
    cur = nxt;
    if (r + stride < R) {
      cur_runs = issue_ray_loads(cur, S, lane, a_in, c_in, a_out, c_out, s_a[buf ^ 1][wid], s_c[buf ^ 1][wid]);
      nxt = load_ray_map(ray_map, min(r + 2 * stride, R - 1));
    }
    // transmittance: local exclusive products, one warp scan of the lane totals per composite
    float p[SPL], pb[SPL], tot = 1.f, totb = 1.f;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const bool ok = (st.valid >> j) & 1u;
      p[j] = tot; pb[j] = totb;
      tot *= ok ? (1.0f - a[j] + 1e-7f) : 1.0f;
      totb *= ok ? (1.0f - ab[j] + 1e-7f) : 1.0f;
    }
    const float incl = scan_mul32(tot, lane), inclb = scan_mul32(totb, lane);
    float ex = __shfl_up_sync(FULL, incl, 1), exb = __shfl_up_sync(FULL, inclb, 1);
    if (lane == 0) { ex = 1.f; exb = 1.f; }
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const float w = a[j] * (ex * p[j]), wb = ab[j] * (exb * pb[j]);
      if (WEIGHTS && ((st.valid >> j) & 1u)) weights[(long long)r * S + SPL * lane + j] = w;
      v[0] += w * c0[j]; v[1] += w * c1[j]; v[2] += w * c2[j]; v[3] += w;
      v[4] += wb * c0[j]; v[5] += wb * c1[j]; v[6] += wb * c2[j];
    }
    const float tsum = warp_sum8(v, lane);
    const float sacc = __shfl_sync(FULL, tsum, 12);
    // value q = lane >> 2: 0..2 rgb, 3 acc, 4..6 bkgr; lane 4q writes it (lanes 4q+1 of the rgb values write the clamped copy)
    {
      const int q = lane >> 2, sub = lane & 3;
      const float x = tsum + ((is_nerf && q < 3) ? 1.0f - sacc : 0.f);
      float* dst = q < 3 ? (sub == 0 ? rgb_raw : rgb) + 3 * r + q : (q == 3 ? acc + r : rgb_b + 3 * r + (q - 4));
      const bool wr = q < 3 ? sub < 2 : (q < 7 && sub == 0);
      if (wr) *dst = (q < 3 && sub == 1) ? fminf(fmaxf(x, 0.f), 1.f) : x;
    }
  }
}

template <bool FULLS>
__global__ void __launch_bounds__(32 * CW, 2) composite_bwd_staged_kernel(
    const float* __restrict__ a_in, const float* __restrict__ c_in, const float* __restrict__ a_out,
    const float* __restrict__ c_out, const int32_t* __restrict__ ray_map, int R, int S, int is_nerf,
    const float* __restrict__ rgb_raw, const float* __restrict__ d_rgb, const float* __restrict__ d_acc,
    const float* __restrict__ d_rgb_b, float* d_a_in, float* d_c_in, float* d_a_out, float* d_c_out) {
  __shared__ __align__(16) float s_a[2][CW][A_ST];
  __shared__ __align__(16) float s_c[2][CW][C_ST];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int stride = gridDim.x * CW;
  int r = blockIdx.x * CW + wid;
  if (r >= R) return;
  RayMapRegs cur = load_ray_map(ray_map, r);
  unsigned cur_runs = issue_ray_loads(cur, S, lane, a_in, c_in, a_out, c_out, s_a[0][wid], s_c[0][wid]);
  RayMapRegs nxt = load_ray_map(ray_map, min(r + stride, R - 1));
  for (int buf = 0; r < R; r += stride, buf ^= 1) {
    float* sa = s_a[buf][wid];
    float* sc = s_c[buf][wid];
    // per-ray upstream gradients
    float g0 = d_rgb ? d_rgb[3 * r] : 0.f, g1 = d_rgb ? d_rgb[3 * r + 1] : 0.f, g2 = d_rgb ? d_rgb[3 * r + 2] : 0.f;
    const float q0 = rgb_raw[3 * r], q1 = rgb_raw[3 * r + 1], q2 = rgb_raw[3 * r + 2];
    if (!(q0 >= 0.f && q0 <= 1.f)) g0 = 0.f;       // clamp(color, 0, 1) passes the gradient on the closed interval
    if (!(q1 >= 0.f && q1 <= 1.f)) g1 = 0.f;
    if (!(q2 >= 0.f && q2 <= 1.f)) g2 = 0.f;
    const float ga = (d_acc ? d_acc[r] : 0.f) - (is_nerf ? (g0 + g1 + g2) : 0.f);
    const float h0 = d_rgb_b ? d_rgb_b[3 * r] : 0.f, h1 = d_rgb_b ? d_rgb_b[3 * r + 1] : 0.f,
                h2 = d_rgb_b ? d_rgb_b[3 * r + 2] : 0.f;
    cp_async_wait_all();
    __syncwarp();
    const RayStage st = decode_ray<FULLS>(cur, cur_runs, S, lane);
    float a[SPL], gk[SPL], gkb[SPL], vv[SPL];
    float p[SPL], pb[SPL], tot = 1.f, totb = 1.f;
    const float v_in = 1.0f - 0.f + 1e-7f;        // 1 - alpha * outer_mask + 1e-7 of an inner sample in the bkgr composite
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const bool ok = (st.valid >> j) & 1u;
      a[j] = ok ? sa[st.pa[j]] : 0.f;
      const float c0 = ok ? sc[st.pc[j]] : 0.f, c1 = ok ? sc[st.pc[j] + 1] : 0.f, c2 = ok ? sc[st.pc[j] + 2] : 0.f;
      gk[j] = ok ? (g0 * c0 + g1 * c1 + g2 * c2 + ga) : 0.f;
      gkb[j] = ok ? (h0 * c0 + h1 * c1 + h2 * c2) : 0.f;
      vv[j] = ok ? (1.0f - a[j] + 1e-7f) : 1.0f;
      p[j] = tot; pb[j] = totb;
      tot *= vv[j]; totb *= ok ? (((st.inner >> j) & 1u) ? v_in : vv[j]) : 1.0f;
    }
    // next ray: start its loads into the other buffer, request the map after it
    cur = nxt;
    if (r + stride < R) {
      cur_runs = issue_ray_loads(cur, S, lane, a_in, c_in, a_out, c_out, s_a[buf ^ 1][wid], s_c[buf ^ 1][wid]);
      nxt = load_ray_map(ray_map, min(r + 2 * stride, R - 1));
    }
    const float incl = scan_mul32(tot, lane), inclb = scan_mul32(totb, lane);
    float ex = __shfl_up_sync(FULL, incl, 1), exb = __shfl_up_sync(FULL, inclb, 1);
    if (lane == 0) { ex = 1.f; exb = 1.f; }
    // forward quantities, colour gradient in place, local exclusive suffix sums of g * w
    float q[SPL], qb[SPL], suf = 0.f, sufb = 0.f;
#pragma unroll
    for (int j = SPL - 1; j >= 0; --j) {
      p[j] *= ex; pb[j] *= exb;                       // T, T_bkgr
      const float w = a[j] * p[j], wb = ((st.inner >> j) & 1u) ? 0.f : a[j] * pb[j];
      if ((st.valid >> j) & 1u) {
        sc[st.pc[j]] = w * g0 + wb * h0; sc[st.pc[j] + 1] = w * g1 + wb * h1; sc[st.pc[j] + 2] = w * g2 + wb * h2;
      }
      q[j] = suf; qb[j] = sufb;
      suf += gk[j] * w; sufb += gkb[j] * wb;
    }
    const float rin = scan_add32_rev(suf, lane), rinb = scan_add32_rev(sufb, lane);
    float rex = __shfl_down_sync(FULL, rin, 1), rexb = __shfl_down_sync(FULL, rinb, 1);
    if (lane == 31) { rex = 0.f; rexb = 0.f; }
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      if ((st.valid >> j) & 1u) {
        float da = gk[j] * p[j] - __fdividef(rex + q[j], vv[j]);
        if (!((st.inner >> j) & 1u)) da += gkb[j] * pb[j] - __fdividef(rexb + qb[j], vv[j]);
        sa[st.pa[j]] = da;
      }
    }
    __syncwarp();
    // send the four runs back, chunk by chunk
    {
      const int hi = st.in_off & 3, ho = st.out_off & 3;
      float* gi = d_a_in + (st.in_off & ~3);
      float* go = d_a_out + (st.out_off & ~3);
#pragma unroll
      for (int t = lane; t < 64; t += 32) {
        const bool in = t < st.nch_ai;
        if (t < st.nch_ai + st.nch_ao)
          unstage_chunk(in ? gi : go, in ? hi : ho, in ? hi + st.n_in : ho + st.n_out, sa + (in ? 0 : 4 * st.nch_ai),
                        in ? t : t - st.nch_ai);
      }
    }
    {
      const int hi = (3 * st.in_off) & 3, ho = (3 * st.out_off) & 3;
      float* gi = d_c_in + ((3ll * st.in_off) & ~3ll);
      float* go = d_c_out + ((3ll * st.out_off) & ~3ll);
#pragma unroll
      for (int t = lane; t < 128; t += 32) {
        const bool in = t < st.nch_ci;
        if (t < st.nch_ci + st.nch_co)
          unstage_chunk(in ? gi : go, in ? hi : ho, in ? hi + 3 * st.n_in : ho + 3 * st.n_out,
                        sc + (in ? 0 : 4 * st.nch_ci), in ? t : t - st.nch_ci);
      }
    }
    __syncwarp();
  }
}

// =====================================================================================================================
// Lean count pass of the geometry (the product path: only the compact lists and the per-ray map are wanted, S <= 160):
// instantiated for 5 sample blocks, no optional outputs, and the inner test without the square root (192 instead of
// ~1000 warp instructions per ray).  Two further variants were built and measured at 32 768 rays and are NOT kept:
// a single-pass kernel with a decoupled look-back over tiles of 8 rays (96 us against 69 us for the three-kernel form:
// a block's 8 warps idle through the look-back), and a compaction pass that stages the lists through shared memory into
// 16-byte stores (53 us against 48 us for the scalar-store kernel: 926 warp instructions per ray, issue-bound).
constexpr int GT = 8;                                      // rays (= warps) per block

// the samples of a ray, lane-per-sample (ZT:730-736): mid points, dists and the inner bit masks; returns the inner count
template <int NBLK>
__device__ __forceinline__ int ray_samples(const float* __restrict__ o, const float* __restrict__ d,
                                           const float* __restrict__ z, int r, int S, int lane, float (&px)[NBLK],
                                           float (&py)[NBLK], float (&pz)[NBLK], float (&ds)[NBLK], unsigned (&m)[NBLK]) {
  const float ox = o[3 * r], oy = o[3 * r + 1], oz = o[3 * r + 2];
  const float rdx = d[3 * r], rdy = d[3 * r + 1], rdz = d[3 * r + 2];
  const float* zr = z + (long long)r * S;
  int n_in = 0;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    const int s = 32 * k + lane;
    const bool ok = s < S;
    px[k] = py[k] = pz[k] = ds[k] = 0.f;
    if (ok) {
      const float z0 = zr[s];
      const float dist = (s + 1 < S) ? __fsub_rn(zr[s + 1], z0) : __fsub_rn(z0, zr[s - 1]);
      const float zm = __fadd_rn(z0, __fmul_rn(dist, 0.5f));
      px[k] = __fadd_rn(ox, __fmul_rn(rdx, zm)); py[k] = __fadd_rn(oy, __fmul_rn(rdy, zm));
      pz[k] = __fadd_rn(oz, __fmul_rn(rdz, zm));
      ds[k] = dist;
    }
    // ||p|| <= 1 with ||p|| = sqrt_rn(q)  <=>  q <= 1 + 2^-23 (sqrt_rn maps (1, 1 + 2^-23] onto 1): no square root needed
    const float q = __fadd_rn(__fadd_rn(__fmul_rn(px[k], px[k]), __fmul_rn(py[k], py[k])), __fmul_rn(pz[k], pz[k]));
    m[k] = __ballot_sync(FULL, ok && q <= 1.00000011920928955078125f);
    n_in += __popc(m[k]);
  }
  return n_in;
}

template <int NBLK>
__global__ void __launch_bounds__(32 * GT) geometry_count_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                                 const float* __restrict__ z, int R, int S,
                                                                 int32_t* __restrict__ ray_inner) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * GT + (threadIdx.x >> 5);
  if (r >= R) return;
  float px[NBLK], py[NBLK], pz[NBLK], ds[NBLK];
  unsigned m[NBLK];
  const int n_in = ray_samples<NBLK>(o, d, z, r, S, lane, px, py, pz, ds, m);
  if (lane == 0) ray_inner[r] = n_in;
}


// ---------------------------------------------------------------------------------------------------------------------
// Stage-2 per-segment compositing in LINEAR colour (ZT:1942-1951): dense rows alpha[N,S], sRGB colour[N,S,3], S <= 256.
//   w_j = alpha_j prod_{i<j}(1 - alpha_i + 1e-7);   rgb_lin = sum_j w_j srgb_to_linear(c_j);   T_end = prod_j(1 - alpha_j + 1e-7)
// One warp per ray, sample j in lane j & 31 of block j >> 5 (coalesced rows); the backward recomputes the transmittance
// with the same scans instead of reading a stored copy.
__device__ __forceinline__ float srgb_to_linear_f(float x) {        // utils/raw_utils.py:21-27
  const float eps = 1.1920928955078125e-07f;
  return x <= 0.04045f ? (25.0f / 323.0f) * x : powf(fmaxf((200.0f * x + 11.0f) / 211.0f, eps), 12.0f / 5.0f);
}
__device__ __forceinline__ float srgb_to_linear_df(float x) {
  const float eps = 1.1920928955078125e-07f;
  if (x <= 0.04045f) return 25.0f / 323.0f;
  const float b = (200.0f * x + 11.0f) / 211.0f;
  return b < eps ? 0.0f : (12.0f / 5.0f) * powf(b, 7.0f / 5.0f) * (200.0f / 211.0f);
}

constexpr int SEG_NBLK = 8;

__global__ void __launch_bounds__(32 * WPB) seg_composite_fwd_kernel(const float* __restrict__ alpha,
                                                                    const float* __restrict__ color, int N, int S,
                                                                    float* __restrict__ rgb_lin, float* __restrict__ t_end) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (r >= N) return;
  const float* a_row = alpha + (long long)r * S;
  const float* c_row = color + (long long)r * S * 3;
  float carry = 1.0f, acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < SEG_NBLK; ++k) {
    if (k * 32 >= S) break;
    const int j = lane + 32 * k;
    const bool ok = j < S;
    const float a = ok ? a_row[j] : 0.0f;
    const float incl = scan_mul32(ok ? 1.0f - a + 1e-7f : 1.0f, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float w = a * carry * excl;
    carry *= __shfl_sync(FULL, incl, 31);
    if (ok) {
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] = fmaf(w, srgb_to_linear_f(c_row[3 * j + c]), acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) acc[c] = warp_sum(acc[c]);
  if (lane == 0) {
    rgb_lin[3 * r] = acc[0]; rgb_lin[3 * r + 1] = acc[1]; rgb_lin[3 * r + 2] = acc[2];
    t_end[r] = carry;
  }
}

__global__ void __launch_bounds__(32 * WPB) seg_composite_bwd_kernel(const float* __restrict__ alpha,
                                                                    const float* __restrict__ color, int N, int S,
                                                                    const float* __restrict__ g_rgb,
                                                                    const float* __restrict__ g_t,
                                                                    float* __restrict__ d_alpha, float* __restrict__ d_color) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (r >= N) return;
  const float* a_row = alpha + (long long)r * S;
  const float* c_row = color + (long long)r * S * 3;
  const float g[3] = {g_rgb[3 * r], g_rgb[3 * r + 1], g_rgb[3 * r + 2]};
  float av[SEG_NBLK], Tv[SEG_NBLK], sv[SEG_NBLK], qin[SEG_NBLK];
  float carry = 1.0f;
#pragma unroll
  for (int k = 0; k < SEG_NBLK; ++k) {
    av[k] = 0.f; Tv[k] = 0.f; sv[k] = 0.f; qin[k] = 0.f;
    if (k * 32 >= S) continue;
    const int j = lane + 32 * k;
    const bool ok = j < S;
    const float a = ok ? a_row[j] : 0.0f;
    const float incl = scan_mul32(ok ? 1.0f - a + 1e-7f : 1.0f, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = carry * excl;
    carry *= __shfl_sync(FULL, incl, 31);
    float s = 0.0f;                                   // d L / d w_j = <srgb_to_linear(c_j), g>
    if (ok) {
      const float w = a * T;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float cs = c_row[3 * j + c];
        s = fmaf(srgb_to_linear_f(cs), g[c], s);
        d_color[((long long)r * S + j) * 3 + c] = w * g[c] * srgb_to_linear_df(cs);
      }
    }
    av[k] = a; Tv[k] = T; sv[k] = s;
    qin[k] = ok ? a * T * s : 0.0f;                    // q_j = w_j s_j
  }
  // exclusive SUFFIX sums of q, accumulated from the far end (a prefix-sum difference would cancel catastrophically
  // behind an opaque sample, where it is divided by 1 - alpha ~ 1e-4)
  const float gt_end = g_t[r] * carry;
  float carry_rev = 0.0f;
#pragma unroll
  for (int k = SEG_NBLK - 1; k >= 0; --k) {
    if (k * 32 >= S) continue;
    const int j = lane + 32 * k;
    const float incl = scan_add32_rev(qin[k], lane);
    const float suffix = carry_rev + (incl - qin[k]);
    carry_rev += __shfl_sync(FULL, incl, 0);
    if (j < S)       // d/d alpha_j: T_j s_j - (sum_{i>j} w_i s_i + g_T T_end) / (1 - alpha_j + 1e-7)
      d_alpha[(long long)r * S + j] = Tv[k] * sv[k] - (suffix + gt_end) / (1.0f - av[k] + 1e-7f);
  }
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, long long M, int C, const int32_t* __restrict__ id,
                                    float* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  long long m = i / C;
  int c = (int)(i % C);
  dst[(long long)id[m] * C + c] = src[i];
}

}  // namespace nunerf

using namespace nunerf;

// NUNERF_COMPOSITE_LEGACY=1 selects the lane-per-sample kernels also where the staged ones apply (A/B measurements)
static const bool g_composite_legacy = [] { const char* e = getenv("NUNERF_COMPOSITE_LEGACY"); return e && e[0] == '1'; }();

// persistent grid of the staged compositing kernels: `per_sm` resident blocks on every SM (or fewer when R is small)
static int staged_grid(int R, int per_sm) {
  static int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
  return min(cdiv(R, CW), sms * per_sm);
}

extern "C" int nunerf_render_geometry(const float* o, const float* d, const float* z, int R, int S, float* dists,
                                      float* pts, int32_t* slot, int32_t* counts, int32_t* ray_scratch, float* pts_in,
                                      float* dists_in, float* dirs_in, int32_t* id_in, float* pts_out, float* dists_out,
                                      float* dirs_out, int32_t* id_out, int32_t* ray_map, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(o && d && z && counts && ray_scratch && (slot || ray_map), "render_geometry: null argument");
  NUNERF_REQUIRE(pts_in && dists_in && dirs_in && pts_out && dists_out && dirs_out, "render_geometry: null compact buffer");
  NUNERF_REQUIRE(R > 0 && S >= 2 && S <= 256, "render_geometry: need 2 <= S <= 256");
  if (!dists && !pts && S <= 160 && !g_composite_legacy) {
    geometry_count_kernel<5><<<cdiv(R, GT), 32 * GT, 0, stream>>>(o, d, z, R, S, ray_scratch);
    NUNERF_CHECK_LAUNCH("geometry_count_kernel");
  } else {
    geometry_kernel<<<cdiv(R, WPB), 32 * WPB, 0, stream>>>(o, d, z, R, S, dists, pts, ray_scratch);
    NUNERF_CHECK_LAUNCH("geometry_kernel");
  }
  ray_scan_kernel<<<cdiv(R, SCAN_BLOCK), SCAN_BLOCK, 0, stream>>>(ray_scratch, R, ray_scratch + R);
  NUNERF_CHECK_LAUNCH("ray_scan_kernel");
  compact_kernel<<<cdiv(R, WPB), 32 * WPB, 0, stream>>>(o, d, z, ray_scratch, ray_scratch + R, counts, R, S, slot, ray_map,
                                                       pts_in, dists_in, dirs_in, id_in, pts_out, dists_out, dirs_out,
                                                       id_out);
  NUNERF_CHECK_LAUNCH("compact_kernel");
  return 0;
}

// per-ray number of samples inside the unit sphere (the same geometry code as nunerf_render_geometry, no compaction): lets
// a chunked / sharded trainer know the global eikonal denominator before any chunk is differentiated
extern "C" int nunerf_inner_counts(const float* o, const float* d, const float* z, int R, int S, int32_t* ray_inner,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NUNERF_REQUIRE(o && d && z && ray_inner && R > 0 && S >= 2 && S <= 160, "inner_counts: bad arguments");
  geometry_count_kernel<5><<<cdiv(R, GT), 32 * GT, 0, stream>>>(o, d, z, R, S, ray_inner);
  NUNERF_CHECK_LAUNCH("geometry_count_kernel");
  return 0;
}

// stage-2 segment compositing (ZT:1942-1951): rgb_lin[N,3] = sum_j w_j srgb_to_linear(colour_j), t_end[N] = prod_j(1 - alpha_j + 1e-7)
extern "C" int nunerf_seg_composite_fwd(const float* alpha, const float* color, int N, int S, float* rgb_lin, float* t_end,
                                        void* stream) {
  NUNERF_REQUIRE(alpha && color && rgb_lin && t_end && N > 0 && S > 0 && S <= 32 * SEG_NBLK, "seg_composite_fwd: bad arguments");
  seg_composite_fwd_kernel<<<cdiv(N, WPB), 32 * WPB, 0, (cudaStream_t)stream>>>(alpha, color, N, S, rgb_lin, t_end);
  NUNERF_CHECK_LAUNCH("seg_composite_fwd_kernel");
  return 0;
}
// its backward: g_rgb[N,3] = dL/d rgb_lin, g_t[N] = dL/d t_end -> d_alpha[N,S], d_color[N,S,3] (transmittance recomputed)
extern "C" int nunerf_seg_composite_bwd(const float* alpha, const float* color, int N, int S, const float* g_rgb,
                                        const float* g_t, float* d_alpha, float* d_color, void* stream) {
  NUNERF_REQUIRE(alpha && color && g_rgb && g_t && d_alpha && d_color && N > 0 && S > 0 && S <= 32 * SEG_NBLK,
                 "seg_composite_bwd: bad arguments");
  seg_composite_bwd_kernel<<<cdiv(N, WPB), 32 * WPB, 0, (cudaStream_t)stream>>>(alpha, color, N, S, g_rgb, g_t, d_alpha, d_color);
  NUNERF_CHECK_LAUNCH("seg_composite_bwd_kernel");
  return 0;
}

extern "C" int nunerf_composite_fwd(const float* alpha_in, const float* color_in, const float* alpha_out,
                                    const float* color_out, const int32_t* slot, int R, int S, int is_nerf, float* rgb,
                                    float* rgb_raw, float* acc, float* rgb_bkgr, float* weights,
                                    const int32_t* ray_map, void* stream) {
  NUNERF_REQUIRE((slot || ray_map) && rgb && rgb_raw && acc && rgb_bkgr && R > 0 && S > 0 && S <= 256,
                 "composite_fwd: bad arguments");
  const bool aligned = (((uintptr_t)alpha_in | (uintptr_t)color_in | (uintptr_t)alpha_out | (uintptr_t)color_out) & 15) == 0;
  if (ray_map && S <= 32 * SPL && aligned && !g_composite_legacy)
    (S == 32 * SPL ? (weights ? composite_fwd_staged_kernel<true, true> : composite_fwd_staged_kernel<true, false>)
                   : (weights ? composite_fwd_staged_kernel<false, true> : composite_fwd_staged_kernel<false, false>))
        <<<staged_grid(R, 4), 32 * CW, 0, (cudaStream_t)stream>>>(
        alpha_in, color_in, alpha_out, color_out, ray_map, R, S, is_nerf, rgb, rgb_raw, acc, rgb_bkgr, weights);
  else if (S <= 160)
    composite_fwd_kernel<5><<<cdiv(R, WPB), 32 * WPB, 0, (cudaStream_t)stream>>>(
        alpha_in, color_in, alpha_out, color_out, slot, ray_map, R, S, is_nerf, rgb, rgb_raw, acc, rgb_bkgr, weights);
  else
    composite_fwd_kernel<8><<<cdiv(R, WPB), 32 * WPB, 0, (cudaStream_t)stream>>>(
        alpha_in, color_in, alpha_out, color_out, slot, ray_map, R, S, is_nerf, rgb, rgb_raw, acc, rgb_bkgr, weights);
  NUNERF_CHECK_LAUNCH("composite_fwd_kernel");
  return 0;
}

extern "C" int nunerf_composite_bwd(const float* alpha_in, const float* color_in, const float* alpha_out,
                                    const float* color_out, const int32_t* slot, int R, int S, int is_nerf,
                                    const float* rgb_raw, const float* d_rgb, const float* d_acc,
                                    const float* d_rgb_bkgr, float* d_alpha_in, float* d_color_in, float* d_alpha_out,
                                    float* d_color_out, const int32_t* ray_map, void* stream) {
  NUNERF_REQUIRE((slot || ray_map) && rgb_raw && d_alpha_in && d_color_in && d_alpha_out && d_color_out && R > 0 && S > 0 &&
                     S <= 256,
                 "composite_bwd: bad arguments");
  const bool aligned = (((uintptr_t)alpha_in | (uintptr_t)color_in | (uintptr_t)alpha_out | (uintptr_t)color_out |
                         (uintptr_t)d_alpha_in | (uintptr_t)d_color_in | (uintptr_t)d_alpha_out | (uintptr_t)d_color_out) & 15) == 0;
  if (ray_map && S <= 32 * SPL && aligned && !g_composite_legacy)
    (S == 32 * SPL ? composite_bwd_staged_kernel<true> : composite_bwd_staged_kernel<false>)<<<staged_grid(R, 2), 32 * CW, 0,
                                                                                             (cudaStream_t)stream>>>(
        alpha_in, color_in, alpha_out, color_out, ray_map, R, S, is_nerf, rgb_raw, d_rgb, d_acc, d_rgb_bkgr, d_alpha_in,
        d_color_in, d_alpha_out, d_color_out);
  else if (S <= 160)
    composite_bwd_kernel<5><<<cdiv(R, WPB), 32 * WPB, 0, (cudaStream_t)stream>>>(
        alpha_in, color_in, alpha_out, color_out, slot, ray_map, R, S, is_nerf, rgb_raw, d_rgb, d_acc, d_rgb_bkgr,
        d_alpha_in, d_color_in, d_alpha_out, d_color_out);
  else
    composite_bwd_kernel<8><<<cdiv(R, WPB), 32 * WPB, 0, (cudaStream_t)stream>>>(
        alpha_in, color_in, alpha_out, color_out, slot, ray_map, R, S, is_nerf, rgb_raw, d_rgb, d_acc, d_rgb_bkgr,
        d_alpha_in, d_color_in, d_alpha_out, d_color_out);
  NUNERF_CHECK_LAUNCH("composite_bwd_kernel");
  return 0;
}

extern "C" int nunerf_scatter_rows(const float* src, int M, int C, const int32_t* sample_id, float* dst, void* stream) {
  NUNERF_REQUIRE(src && sample_id && dst && M > 0 && C > 0, "scatter_rows: bad arguments");
  long long total = (long long)M * C;
  scatter_rows_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, M, C, sample_id, dst);
  NUNERF_CHECK_LAUNCH("scatter_rows_kernel");
  return 0;
}
