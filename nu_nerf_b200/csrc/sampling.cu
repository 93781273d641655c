// sampling.cu -- per-ray warp-cooperative sampling kernels (stage-1 sample_ray, ZT:572-612).
//
// One warp owns one ray.  Sample j of a ray lives in lane (j & 31), block (j >> 5), so every global access of
// a [R, n] row is a coalesced 128-byte request and no shared-memory staging is needed for the scans: each
// block of 32 samples is scanned with 5 shuffle steps (Hillis-Steele) and chained through a scalar carry.
//
// Arithmetic contract (bit-exact against oracle/sampling_oracle.c): every fp32 operation is an explicitly
// rounded __f{add,sub,mul,div}_rn (no FMA contraction), exp() is nunerf::det_exp, and the scan order is
//   incl = HS32(v);  out_i = carry (op) shfl_up(incl, 1);  carry <- carry (op) incl_31      (exclusive product)
//   cdf_i = carry + incl_i;  carry <- cdf_31                                                  (inclusive sum)
//   total = butterfly xor-reduce over lanes of the per-lane sequential block sums.
#include "common.cuh"

namespace nunerf {

constexpr int WARPS_PER_BLOCK = 4;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float hs_scan_mul(float x, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float t = __shfl_up_sync(FULL, x, off);
    if (lane >= off) x = det_mul(t, x);
  }
  return x;
}
__device__ __forceinline__ float hs_scan_add(float x, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float t = __shfl_up_sync(FULL, x, off);
    if (lane >= off) x = det_add(t, x);
  }
  return x;
}
__device__ __forceinline__ float xor_reduce_add(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x = det_add(x, __shfl_xor_sync(FULL, x, off));
  return x;
}

// tables: [0,64) linspace(0,1,64) | [64,96) bg lower | [96,128) bg (upper-lower) | [128,160) bg unperturbed
__global__ void ray_setup_kernel(const float* __restrict__ o, const float* __restrict__ d, float* near, float* far,
                                 const float* __restrict__ U0, const float* __restrict__ U1,
                                 const float* __restrict__ tables, int R, int sphere, int perturb, float* z,
                                 float* z_bg) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= R) return;
  float nr, fr;
  if (sphere) {
    // ZT:320-327
    float ox = o[3 * r], oy = o[3 * r + 1], oz = o[3 * r + 2];
    float dx = d[3 * r], dy = d[3 * r + 1], dz = d[3 * r + 2];
    float a = det_add(det_add(det_mul(dx, dx), det_mul(dy, dy)), det_mul(dz, dz));
    float b = det_mul(2.0f, det_add(det_add(det_mul(ox, dx), det_mul(oy, dy)), det_mul(oz, dz)));
    float mid = det_div(det_mul(0.5f, -b), a);
    nr = fmaxf(det_sub(mid, 1.0f), 1e-3f);
    fr = det_add(mid, 1.0f);
    if (lane == 0) { near[r] = nr; far[r] = fr; }
  } else {
    nr = near[r];
    fr = far[r];
  }
  const float span = det_sub(fr, nr);
  float shift = 0.f;
  if (perturb) shift = det_div(det_mul(det_sub(U0[r], 0.5f), 2.0f), 64.0f);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    int j = lane + 32 * k;
    float v = det_add(nr, det_mul(span, tables[j]));
    if (perturb) v = det_add(v, shift);
    z[(long long)r * 64 + j] = v;
  }
  // background: far / flip(b) + 1/32  (ZT:582-594)
  int jj = 31 - lane;
  float b = perturb ? det_add(tables[64 + jj], det_mul(tables[96 + jj], U1[(long long)r * 32 + jj])) : tables[128 + jj];
  z_bg[(long long)r * 32 + lane] = det_add(det_div(fr, b), 0.03125f);
}

__global__ void points_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z,
                              long long total, int n, float* pts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long r = i / n;
  float zz = z[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[3 * i + c] = det_add(o[3 * r + c], det_mul(d[3 * r + c], zz));
}

// One importance round.  n <= 128 samples in, n_new <= 32 out.  NBLK = ceil(n / 32) blocks of 32 samples per lane: the
// kernel is issue-bound (fixed-order scans, IEEE divisions and expf of the logistic CDF), so it is instantiated for 2, 3
// and 4 blocks instead of always walking four (n = 64 is half of that); skipped blocks only ever contributed
// identities (weight 0, factor 1), so the results are bit-identical.
template <int NBLK>
__global__ void upsample_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z,
                                const float* __restrict__ sdf, int R, int n, int n_new,
                                const float* __restrict__ inv_s_dev, float inv_s_cap, const float* __restrict__ u_tab,
                                float* z_new, int32_t* inds, float* z_merged, int32_t* perm) {
  __shared__ float s_z[WARPS_PER_BLOCK][128];
  __shared__ float s_cdf[WARPS_PER_BLOCK][132];
  __shared__ float s_new[WARPS_PER_BLOCK][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS_PER_BLOCK + w;
  if (r >= R) return;
  const float inv_s = fminf(inv_s_dev[0], inv_s_cap);
  const float ox = o[3 * r], oy = o[3 * r + 1], oz = o[3 * r + 2];
  const float dx = d[3 * r], dy = d[3 * r + 1], dz = d[3 * r + 2];

  // ---- load + radius
  float zv[NBLK], sv[NBLK], rad[NBLK];
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    int j = lane + 32 * k;
    bool ok = j < n;
    zv[k] = ok ? z[(long long)r * n + j] : 0.f;
    sv[k] = ok ? sdf[(long long)r * n + j] : 0.f;
    float px = det_add(ox, det_mul(dx, zv[k])), py = det_add(oy, det_mul(dy, zv[k])), pz = det_add(oz, det_mul(dz, zv[k]));
    rad[k] = det_sqrt(det_add(det_add(det_mul(px, px), det_mul(py, py)), det_mul(pz, pz)));
    if (ok) s_z[w][j] = zv[k];
  }
  __syncwarp();
  // ---- per-section quantities; section j uses samples j, j+1 and the raw cosine of section j-1
  float alpha[NBLK], wgt[NBLK];
  float carry_prev_cos = 0.f;  // raw cosine of the last section of the previous block
  float carryT = 1.0f;
  float lane_sum = 0.f;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    int j = lane + 32 * k;
    // neighbour j+1: lane+1 of this block, or lane 0 of the next block
    float zn = __shfl_down_sync(FULL, zv[k], 1), sn = __shfl_down_sync(FULL, sv[k], 1),
          rn = __shfl_down_sync(FULL, rad[k], 1);
    constexpr int LAST = NBLK - 1;
    const int kn = k + 1 > LAST ? LAST : k + 1;
    float zn2 = __shfl_sync(FULL, k < LAST ? zv[kn] : 0.f, 0);
    float sn2 = __shfl_sync(FULL, k < LAST ? sv[kn] : 0.f, 0);
    float rn2 = __shfl_sync(FULL, k < LAST ? rad[kn] : 0.f, 0);
    if (lane == 31) { zn = zn2; sn = sn2; rn = rn2; }
    const bool sec_ok = (j < n - 1);
    float dist = det_sub(zn, zv[k]);
    float cosv = det_div(det_sub(sn, sv[k]), det_add(dist, 1e-5f));
    float prev = __shfl_up_sync(FULL, cosv, 1);
    if (lane == 0) prev = carry_prev_cos;
    carry_prev_cos = __shfl_sync(FULL, cosv, 31);
    float c = fminf(prev, cosv);
    c = fminf(fmaxf(c, -1e3f), 0.0f);
    const bool inside = (rad[k] < 1.0f) || (rn < 1.0f);
    c = inside ? c : det_mul(c, 0.0f);
    float mid = det_mul(det_add(sv[k], sn), 0.5f);
    float half = det_mul(det_mul(c, dist), 0.5f);
    float pe = det_sub(mid, half), ne = det_add(mid, half);
    float pc = det_sigmoid(det_mul(pe, inv_s)), nc = det_sigmoid(det_mul(ne, inv_s));
    float a = det_div(det_add(det_sub(pc, nc), 1e-5f), det_add(pc, 1e-5f));
    alpha[k] = sec_ok ? a : 0.f;
    float v = sec_ok ? det_add(det_sub(1.0f, a), 1e-7f) : 1.0f;
    float incl = hs_scan_mul(v, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    float T = det_mul(carryT, excl);
    carryT = det_mul(carryT, __shfl_sync(FULL, incl, 31));
    wgt[k] = sec_ok ? det_add(det_mul(alpha[k], T), 1e-5f) : 0.f;  // sample_pdf's +1e-5 (field.py:471)
    lane_sum = det_add(lane_sum, wgt[k]);
  }
  const float total = xor_reduce_add(lane_sum);
  // ---- cdf = [0, cumsum(pdf)]
  float carryC = 0.f;
  if (lane == 0) s_cdf[w][0] = 0.f;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    int j = lane + 32 * k;
    float pdf = (j < n - 1) ? det_div(wgt[k], total) : 0.f;
    float incl = hs_scan_add(pdf, lane);
    float c = det_add(carryC, incl);
    carryC = __shfl_sync(FULL, c, 31);
    if (j < n - 1) s_cdf[w][j + 1] = c;
  }
  __syncwarp();
  // ---- invert the cdf at n_new deterministic u's (field.py:476-496)
  float zs = 0.f;
  int ind = 0;
  if (lane < n_new) {
    const float u = u_tab[lane];
    int lo = 0, hi = n;  // first index with cdf > u  (searchsorted right=True)
    while (lo < hi) {
      int m = (lo + hi) >> 1;
      if (s_cdf[w][m] <= u) lo = m + 1; else hi = m;
    }
    ind = lo;
    int below = ind - 1 < 0 ? 0 : ind - 1;
    int above = ind > n - 1 ? n - 1 : ind;
    float c0 = s_cdf[w][below], c1 = s_cdf[w][above];
    float b0 = s_z[w][below], b1 = s_z[w][above];
    float den = det_sub(c1, c0);
    if (den < 1e-5f) den = 1.0f;
    float t = det_div(det_sub(u, c0), den);
    zs = det_add(b0, det_mul(t, det_sub(b1, b0)));
    s_new[w][lane] = zs;
    z_new[(long long)r * n_new + lane] = zs;
    inds[(long long)r * n_new + lane] = ind;
  }
  __syncwarp();
  // ---- stable merge (old samples first on ties): equals a stable sort of cat(z, z_new) (ZT:560-561)
  const int nm = n + n_new;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    int j = lane + 32 * k;
    if (j < n) {
      int cnt = 0;
      for (int t = 0; t < n_new; ++t) cnt += (s_new[w][t] < zv[k]) ? 1 : 0;
      z_merged[(long long)r * nm + j + cnt] = zv[k];
      perm[(long long)r * nm + j + cnt] = j;
    }
  }
  if (lane < n_new) {
    int lo = 0, hi = n;  // number of old samples <= zs
    while (lo < hi) {
      int m = (lo + hi) >> 1;
      if (s_z[w][m] <= zs) lo = m + 1; else hi = m;
    }
    z_merged[(long long)r * nm + lane + lo] = zs;
    perm[(long long)r * nm + lane + lo] = n + lane;
  }
}


// NeRF-guided importance sampling of the stage-2 miss rays (upsample_nerf + cat_z_vals_nerf, ZT:1367-1397, called at
// ZT:1762-1799): weights_j = alpha_j prod_{i<j}(1 - alpha_i + 1e-7) over the n samples, sample_pdf(z, weights[:-1], n_new,
// det) (field.py:468-498) and the sorted merge of z with the new samples.  One warp per ray, n <= 32 NBLK, n_new <= 64.
// Same fixed scan order / explicitly rounded fp32 arithmetic as upsample_kernel.
template <int NBLK>
__global__ void alpha_importance_kernel(const float* __restrict__ z, const float* __restrict__ alpha, int R, int n, int n_new,
                                        const float* __restrict__ u_tab, float* __restrict__ z_merged) {
  __shared__ float s_z[WARPS_PER_BLOCK][32 * NBLK];
  __shared__ float s_cdf[WARPS_PER_BLOCK][32 * NBLK + 4];
  __shared__ float s_new[WARPS_PER_BLOCK][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS_PER_BLOCK + w;
  if (r >= R) return;
  float zv[NBLK], wgt[NBLK];
  float carryT = 1.0f, lane_sum = 0.f;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    const int j = lane + 32 * k;
    const bool ok = j < n;
    zv[k] = ok ? z[(long long)r * n + j] : 0.f;
    const float a = ok ? alpha[(long long)r * n + j] : 0.f;
    if (ok) s_z[w][j] = zv[k];
    const float incl = hs_scan_mul(ok ? det_add(det_sub(1.0f, a), 1e-7f) : 1.0f, lane);
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = det_mul(carryT, excl);
    carryT = det_mul(carryT, __shfl_sync(FULL, incl, 31));
    wgt[k] = (j < n - 1) ? det_add(det_mul(a, T), 1e-5f) : 0.f;       // weights[:, :-1] + 1e-5 (field.py:471)
    lane_sum = det_add(lane_sum, wgt[k]);
  }
  const float total = xor_reduce_add(lane_sum);
  float carryC = 0.f;
  if (lane == 0) s_cdf[w][0] = 0.f;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    const int j = lane + 32 * k;
    const float pdf = (j < n - 1) ? det_div(wgt[k], total) : 0.f;
    const float c = det_add(carryC, hs_scan_add(pdf, lane));
    carryC = __shfl_sync(FULL, c, 31);
    if (j < n - 1) s_cdf[w][j + 1] = c;
  }
  __syncwarp();
  for (int t = lane; t < n_new; t += 32) {
    const float u = u_tab[t];
    int lo = 0, hi = n;                         // first index with cdf > u (searchsorted right=True) over the n cdf entries
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (s_cdf[w][m] <= u) lo = m + 1; else hi = m;
    }
    const int below = lo - 1 < 0 ? 0 : lo - 1, above = lo > n - 1 ? n - 1 : lo;
    const float c0 = s_cdf[w][below], c1 = s_cdf[w][above];
    const float b0 = s_z[w][below], b1 = s_z[w][above];
    float den = det_sub(c1, c0);
    if (den < 1e-5f) den = 1.0f;
    s_new[w][t] = det_add(b0, det_mul(det_div(det_sub(u, c0), den), det_sub(b1, b0)));
  }
  __syncwarp();
  // sorted merge (values only): old sample j moves up by the number of new samples below it, new sample t by the number of
  // old samples <= it (both sequences are non-decreasing)
  const int nm = n + n_new;
#pragma unroll
  for (int k = 0; k < NBLK; ++k) {
    const int j = lane + 32 * k;
    if (j < n) {
      int cnt = 0;
      for (int t = 0; t < n_new; ++t) cnt += (s_new[w][t] < zv[k]) ? 1 : 0;
      z_merged[(long long)r * nm + j + cnt] = zv[k];
    }
  }
  for (int t = lane; t < n_new; t += 32) {
    const float zs = s_new[w][t];
    int lo = 0, hi = n;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (s_z[w][m] <= zs) lo = m + 1; else hi = m;
    }
    z_merged[(long long)r * nm + t + lo] = zs;
  }
}

// Occlusion probe (get_weights + sample_pdf of get_intersection, field.py:501-554): one warp per probe ray with n <= 64
// samples z / sdf.  weights_j = alpha_j * prod_{k<j}(1 - alpha_k + 1e-7) with alpha from the logistic CDF at the section
// ends, zeroed where the SDF does not decrease (surface_mask).  Either inverts the CDF of (weights + 1e-5) at n_new
// deterministic u's (first pass -> z_new) or returns sum_j weights_j (second pass -> the hit probability).  Float
// arithmetic only feeds a regression target, so no fixed operation order is imposed here.
__global__ void probe_weights_kernel(const float* __restrict__ z, const float* __restrict__ sdf, int P, int n,
                                     const float* __restrict__ inv_s_dev, int n_new, const float* __restrict__ u_tab,
                                     float* z_new, float* wsum) {
  __shared__ float s_z[WARPS_PER_BLOCK][64];
  __shared__ float s_cdf[WARPS_PER_BLOCK][68];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * WARPS_PER_BLOCK + w;
  if (r >= P) return;
  const float inv_s = inv_s_dev[0];
  float zv[2], sv[2], wgt[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int j = lane + 32 * k;
    const bool ok = j < n;
    zv[k] = ok ? z[(long long)r * n + j] : 0.f;
    sv[k] = ok ? sdf[(long long)r * n + j] : 0.f;
    if (ok) s_z[w][j] = zv[k];
  }
  float carryT = 1.0f, lane_sum = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int j = lane + 32 * k;
    float zn = __shfl_down_sync(FULL, zv[k], 1), sn = __shfl_down_sync(FULL, sv[k], 1);
    const float zn2 = __shfl_sync(FULL, k == 0 ? zv[1] : 0.f, 0), sn2 = __shfl_sync(FULL, k == 0 ? sv[1] : 0.f, 0);
    if (lane == 31) { zn = zn2; sn = sn2; }
    const bool sec_ok = j < n - 1;
    const float dist = zn - zv[k];
    float cosv = (sn - sv[k]) / (dist + 1e-5f);
    const bool surf = cosv < 0.f;
    cosv = fminf(cosv, 0.f);
    const float mid = (sv[k] + sn) * 0.5f, half = cosv * dist * 0.5f;
    const float pc = 1.0f / (1.0f + __expf(-(mid - half) * inv_s)), nc = 1.0f / (1.0f + __expf(-(mid + half) * inv_s));
    float a = (pc - nc + 1e-5f) / (pc + 1e-5f);
    a = (sec_ok && surf) ? a : 0.f;
    float v = sec_ok ? (1.0f - a + 1e-7f) : 1.0f;
    float incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float t = __shfl_up_sync(FULL, incl, off);
      if (lane >= off) incl *= t;
    }
    float excl = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) excl = 1.0f;
    wgt[k] = sec_ok ? a * (carryT * excl) : 0.f;
    carryT *= __shfl_sync(FULL, incl, 31);
    lane_sum += wgt[k];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) lane_sum += __shfl_xor_sync(FULL, lane_sum, off);
  if (wsum && lane == 0) wsum[r] = lane_sum;
  if (!z_new) return;
  // cdf of (weights + 1e-5) over the n - 1 sections (sample_pdf, field.py:468-498)
  const float total = lane_sum + 1e-5f * (float)(n - 1);
  float carryC = 0.f;
  if (lane == 0) s_cdf[w][0] = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int j = lane + 32 * k;
    float pdf = (j < n - 1) ? (wgt[k] + 1e-5f) / total : 0.f;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float t = __shfl_up_sync(FULL, pdf, off);
      if (lane >= off) pdf += t;
    }
    const float c = carryC + pdf;
    carryC = __shfl_sync(FULL, c, 31);
    if (j < n - 1) s_cdf[w][j + 1] = c;
  }
  __syncwarp();
  if (lane < n_new) {
    const float u = u_tab[lane];
    int lo = 0, hi = n;                      // first index with cdf > u (searchsorted right=True)
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (s_cdf[w][m] <= u) lo = m + 1; else hi = m;
    }
    const int below = lo - 1 < 0 ? 0 : lo - 1, above = lo > n - 1 ? n - 1 : lo;
    const float c0 = s_cdf[w][below], c1 = s_cdf[w][above];
    const float b0 = s_z[w][below], b1 = s_z[w][above];
    float den = c1 - c0;
    if (den < 1e-5f) den = 1.0f;
    z_new[(long long)r * n_new + lane] = b0 + (u - c0) / den * (b1 - b0);
  }
}

__global__ void merge_sdf_kernel(const float* __restrict__ sdf, const float* __restrict__ sdf_new,
                                 const int32_t* __restrict__ perm, long long total, int n, int n_new, float* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int nm = n + n_new;
  long long r = i / nm;
  int src = perm[i];
  out[i] = src < n ? sdf[r * n + src] : sdf_new[r * n_new + (src - n)];
}

}  // namespace nunerf

using namespace nunerf;

extern "C" int nunerf_ray_setup(const float* o, const float* d, float* near, float* far, const float* U0,
                                const float* U1, const float* tables, int R, int sphere, int perturb, float* z,
                                float* z_bg, void* stream) {
  NUNERF_REQUIRE(o && d && near && far && tables && z && z_bg && R > 0, "ray_setup: bad arguments");
  NUNERF_REQUIRE(!perturb || (U0 && U1), "ray_setup: perturb needs uniforms");
  ray_setup_kernel<<<cdiv(R, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      o, d, near, far, U0, U1, tables, R, sphere, perturb, z, z_bg);
  NUNERF_CHECK_LAUNCH("ray_setup_kernel");
  return 0;
}

// reverse of points_kernel: pts[r, j] = o[r] + d[r] z[r, j]  ->  g_o[r] = sum_j g[r, j],  g_d[r] = sum_j z[r, j] g[r, j]
// (one warp per ray, lanes stride over the samples, shuffle reduction)
__global__ void points_bwd_kernel(const float* __restrict__ g_pts, const float* __restrict__ z, int R, int n, float* g_o,
                                  float* g_d) {
  const int ray = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (ray >= R) return;
  float a[3] = {0.f, 0.f, 0.f}, b[3] = {0.f, 0.f, 0.f};
  for (int j = lane; j < n; j += 32) {
    const float zz = z[(long long)ray * n + j];
    const float* g = g_pts + ((long long)ray * n + j) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) { a[c] += g[c]; b[c] += zz * g[c]; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a[c] += __shfl_xor_sync(0xffffffffu, a[c], off);
      b[c] += __shfl_xor_sync(0xffffffffu, b[c], off);
    }
  if (lane == 0)
    for (int c = 0; c < 3; ++c) { g_o[3 * ray + c] = a[c]; g_d[3 * ray + c] = b[c]; }
}

extern "C" int nunerf_points_bwd(const float* g_pts, const float* z, int R, int n, float* g_o, float* g_d, void* stream) {
  NUNERF_REQUIRE(g_pts && z && g_o && g_d && R > 0 && n > 0, "points_bwd: bad arguments");
  points_bwd_kernel<<<cdiv((long long)R * 32, 256), 256, 0, (cudaStream_t)stream>>>(g_pts, z, R, n, g_o, g_d);
  NUNERF_CHECK_LAUNCH("points_bwd_kernel");
  return 0;
}

extern "C" int nunerf_points(const float* o, const float* d, const float* z, int R, int n, float* pts, void* stream) {
  NUNERF_REQUIRE(o && d && z && pts && R > 0 && n > 0, "points: bad arguments");
  long long total = (long long)R * n;
  points_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(o, d, z, total, n, pts);
  NUNERF_CHECK_LAUNCH("points_kernel");
  return 0;
}

extern "C" int nunerf_upsample(const float* o, const float* d, const float* z, const float* sdf, int R, int n,
                               int n_new, const float* inv_s_dev, float inv_s_cap, const float* u_tab, float* z_new,
                               int32_t* inds, float* z_merged, int32_t* perm, void* stream) {
  NUNERF_REQUIRE(o && d && z && sdf && inv_s_dev && u_tab && z_new && inds && z_merged && perm, "upsample: null argument");
  NUNERF_REQUIRE(R > 0 && n >= 2 && n <= 128 && n_new >= 1 && n_new <= 32, "upsample: need 2<=n<=128, 1<=n_new<=32");
  const dim3 grid(cdiv(R, WARPS_PER_BLOCK)), block(32 * WARPS_PER_BLOCK);
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 64)
    upsample_kernel<2><<<grid, block, 0, st>>>(o, d, z, sdf, R, n, n_new, inv_s_dev, inv_s_cap, u_tab, z_new, inds, z_merged, perm);
  else if (n <= 96)
    upsample_kernel<3><<<grid, block, 0, st>>>(o, d, z, sdf, R, n, n_new, inv_s_dev, inv_s_cap, u_tab, z_new, inds, z_merged, perm);
  else
    upsample_kernel<4><<<grid, block, 0, st>>>(o, d, z, sdf, R, n, n_new, inv_s_dev, inv_s_cap, u_tab, z_new, inds, z_merged, perm);
  NUNERF_CHECK_LAUNCH("upsample_kernel");
  return 0;
}

// z_merged[R, n + n_new] = sort(cat(z, sample_pdf(z, (alpha T)[:-1], n_new)))  (ZT:1367-1397); u_tab[n_new] = the det. u's
extern "C" int nunerf_alpha_importance(const float* z, const float* alpha, int R, int n, int n_new, const float* u_tab,
                                       float* z_merged, void* stream) {
  NUNERF_REQUIRE(z && alpha && u_tab && z_merged && R > 0 && n >= 2 && n <= 256 && n_new >= 1 && n_new <= 64,
                 "alpha_importance: need 2 <= n <= 256, 1 <= n_new <= 64");
  const dim3 grid(cdiv(R, WARPS_PER_BLOCK)), block(32 * WARPS_PER_BLOCK);
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 128) alpha_importance_kernel<4><<<grid, block, 0, st>>>(z, alpha, R, n, n_new, u_tab, z_merged);
  else if (n <= 192) alpha_importance_kernel<6><<<grid, block, 0, st>>>(z, alpha, R, n, n_new, u_tab, z_merged);
  else alpha_importance_kernel<8><<<grid, block, 0, st>>>(z, alpha, R, n, n_new, u_tab, z_merged);
  NUNERF_CHECK_LAUNCH("alpha_importance_kernel");
  return 0;
}

extern "C" int nunerf_probe_weights(const float* z, const float* sdf, int P, int n, const float* inv_s_dev, int n_new,
                                    const float* u_tab, float* z_new, float* wsum, void* stream) {
  NUNERF_REQUIRE(z && sdf && inv_s_dev && (z_new || wsum) && P > 0, "probe_weights: bad arguments");
  NUNERF_REQUIRE(n >= 2 && n <= 64 && (!z_new || (u_tab && n_new >= 1 && n_new <= 32)), "probe_weights: need 2<=n<=64, n_new<=32");
  probe_weights_kernel<<<cdiv(P, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      z, sdf, P, n, inv_s_dev, n_new, u_tab, z_new, wsum);
  NUNERF_CHECK_LAUNCH("probe_weights_kernel");
  return 0;
}

extern "C" int nunerf_merge_sdf(const float* sdf, const float* sdf_new, const int32_t* perm, int R, int n, int n_new,
                                float* sdf_merged, void* stream) {
  NUNERF_REQUIRE(sdf && sdf_new && perm && sdf_merged && R > 0, "merge_sdf: bad arguments");
  long long total = (long long)R * (n + n_new);
  merge_sdf_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(sdf, sdf_new, perm, total, n, n_new, sdf_merged);
  NUNERF_CHECK_LAUNCH("merge_sdf_kernel");
  return 0;
}
