// field.cu -- encodings and pointwise field kernels (everything on the path that is not a dense layer,
// a scan or a traversal).  Per-point math lives in pointwise.cuh; the kernels here only move data:
// fp32 compact per-sample arrays in, bf16 "planes" (tensor-core operands) or fp32 out.
#include "common.cuh"
#include "pointwise.cuh"

namespace nunerf {

__constant__ pw::IdeTable c_ide;
static pw::IdeTable h_ide;
static bool g_ide_ready = false;

static int ensure_ide() {
  if (g_ide_ready) return 0;
  pw::build_ide_table(&h_ide);
  cudaError_t e = cudaMemcpyToSymbol(c_ide, &h_ide, sizeof(h_ide));
  if (e != cudaSuccess) return fail("ide table upload: %s", cudaGetErrorString(e), -2);
  g_ide_ready = true;
  return 0;
}

// ------------------------------------------------------------------------------------------- positional encoding
// Row helpers: a thread builds one whole operand row in registers (every index is a compile-time constant once the
// loops are unrolled) and writes it as 16-byte packed bf16 chunks, hi plane and (split mode) lo plane.
__device__ __forceinline__ void store8(__nv_bfloat16* dst, long long idx, int lo, const float* v) {
  uint32_t h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(dst + idx) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) {
    uint32_t l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) l[j] = pack_bf16x2(v[2 * j] - bf16lo_to_f(h[j]), v[2 * j + 1] - bf16hi_to_f(h[j]));
    *reinterpret_cast<uint4*>(dst + idx + lo) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// 16 columns as ONE 32-byte store per plane (STG.256: a whole sector per lane and half the store requests of two store8)
// when the address allows it
__device__ __forceinline__ void store16(__nv_bfloat16* dst, long long idx, int lo, const float* v) {
  if ((((uintptr_t)(dst + idx)) & 31) != 0 || (lo & 15) != 0) {
    store8(dst, idx, lo, v);
    store8(dst, idx + 8, lo, v + 8);
    return;
  }
  uint32_t h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + idx), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]),
               "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7])
               : "memory");
  if (lo) {
    uint32_t l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = pack_bf16x2(v[2 * j] - bf16lo_to_f(h[j]), v[2 * j + 1] - bf16hi_to_f(h[j]));
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + idx + lo), "r"(l[0]), "r"(l[1]), "r"(l[2]),
                 "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7])
                 : "memory");
  }
}

// [x (D), sin(2^0 x) (D), cos(2^0 x) (D), sin(2^1 x) (D), ...] (field.py:14-61) -> row[0 .. D(1+2F))
template <int D, int F>
__device__ __forceinline__ void fill_pe(float* row, const float* x) {
#pragma unroll
  for (int c = 0; c < D; ++c) row[c] = x[c];
#pragma unroll
  for (int k = 0; k < F; ++k)
#pragma unroll
    for (int c = 0; c < D; ++c) {
      float sn, co;
      sincosf(x[c] * (float)(1 << k), &sn, &co);
      row[D + (2 * k) * D + c] = sn;
      row[D + (2 * k + 1) * D + c] = co;
    }
}

// dst[row_off + m, col_off + j] for j < width; j >= D(1+2F) is zero fill.  One thread per row.
template <int D, int F>
__global__ void __launch_bounds__(128)
encode_pe_kernel(const float* __restrict__ x, long long M, __nv_bfloat16* dst, int ld, int lo_off, int col_off,
                 long long row_off, int width) {
  constexpr int REAL = D * (1 + 2 * F);
  constexpr int PADW = (REAL + 7) / 8 * 8;
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float xin[D];
#pragma unroll
  for (int c = 0; c < D; ++c) xin[c] = x[m * D + c];
  float row[PADW];
  fill_pe<D, F>(row, xin);
#pragma unroll
  for (int j = REAL; j < PADW; ++j) row[j] = 0.f;
  const long long base = (row_off + m) * ld + col_off;
  if (((col_off | ld | lo_off) & 7) == 0 && (width & 7) == 0) {
    // 16 columns per store where a pair of 8-column groups is available (32-byte stores), 8 for an odd tail group
    const float zero[16] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    constexpr int G = PADW / 8;                 // 8-column groups that hold real values
#pragma unroll
    for (int c = 0; c + 1 < G; c += 2)
      if (c * 8 + 16 <= width) store16(dst, base + c * 8, lo_off, row + c * 8);
      else if (c * 8 < width) store8(dst, base + c * 8, lo_off, row + c * 8);
    int c_next = G & ~1;
    if (G & 1) {
      if ((G - 1) * 8 < width) store8(dst, base + (G - 1) * 8, lo_off, row + (G - 1) * 8);
      c_next = G;
    }
    for (int c = c_next; c * 8 < width;) {
      if ((c & 1) == 0 && c * 8 + 16 <= width) { store16(dst, base + c * 8, lo_off, zero); c += 2; }
      else { store8(dst, base + c * 8, lo_off, zero); c += 1; }
    }
  } else {
#pragma unroll
    for (int j = 0; j < PADW; ++j)
      if (j < width) store_planes(dst, base + j, lo_off, row[j]);
    for (int j = PADW; j < width; ++j) store_planes(dst, base + j, lo_off, 0.f);
  }
}

// grad_x = J_pe^T (ga + gb), PE-6 of a 3-vector (39 columns)
__global__ void sdf_grad_pe_kernel(const float* __restrict__ x, const float* __restrict__ ga, int lda,
                                   const float* __restrict__ gb, int ldb, long long M, float* grad) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* a = ga + m * lda;
  const float* b = gb ? gb + m * ldb : nullptr;
  for (int c = 0; c < 3; ++c) {
    float xc = x[3 * m + c];
    float g = a[c] + (b ? b[c] : 0.f);
    for (int k = 0; k < 6; ++k) {
      float f = (float)(1 << k);
      float s, co;
      sincosf(xc * f, &s, &co);
      int is = 3 + 6 * k + c, ic = 6 + 6 * k + c;
      g += (a[is] + (b ? b[is] : 0.f)) * f * co - (a[ic] + (b ? b[ic] : 0.f)) * f * s;
    }
    grad[3 * m + c] = g;
  }
}

// u~ = J_pe d_grad written to up to two plane destinations (39 real columns, `width` written, zero padded)
__global__ void sdf_grad_pe_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dgrad, long long M,
                                       __nv_bfloat16* d1, int ld1, int lo1, int col1, int width1, __nv_bfloat16* d2,
                                       int ld2, int lo2, int col2, int width2) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int wmax = width1 > width2 ? width1 : width2;
  if (idx >= M * wmax) return;
  long long m = idx / wmax;
  int j = (int)(idx % wmax);
  float v = 0.f;
  if (j < 3) v = dgrad[3 * m + j];
  else if (j < 39) {
    int t = (j - 3) / 3, c = (j - 3) % 3, k = t >> 1;
    float f = (float)(1 << k);
    float s, co;
    sincosf(x[3 * m + c] * f, &s, &co);
    v = dgrad[3 * m + c] * ((t & 1) ? -f * s : f * co);
  }
  if (d1 && j < width1) store_planes(d1, m * ld1 + col1 + j, lo1, v);
  if (d2 && j < width2) store_planes(d2, m * ld2 + col2 + j, lo2, v);
}

// dx[m, c] (+)= [J_pe(x)^T (ga + gb)]_c for a d-dimensional input with F frequencies (d <= 4): the input gradient of any
// positional encoding (PE-6 of points / directions, PE-10 of the NeRF++ 4-vector, PE-4 of its view direction)
__global__ void pe_bwd_kernel(const float* __restrict__ x, int d, int F, const float* __restrict__ ga, int lda,
                              const float* __restrict__ gb, int ldb, long long M, float* dx, int accumulate) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * d) return;
  const long long m = idx / d;
  const int c = (int)(idx % d);
  const int W = d * (1 + 2 * F);
  const float* a = ga + m * lda;
  const float* b = gb ? gb + m * ldb : nullptr;
  const float xc = x[m * d + c];
  float acc = a[c] + (b ? b[c] : 0.f);
  for (int k = 0; k < F; ++k) {
    const float f = (float)(1 << k);
    float sn, co;
    sincosf(xc * f, &sn, &co);
    const int is = pw::pe_index(d, k, 0, c), ic = pw::pe_index(d, k, 1, c);
    acc += f * (co * (a[is] + (b ? b[is] : 0.f)) - sn * (a[ic] + (b ? b[ic] : 0.f)));
  }
  (void)W;
  dx[idx] = accumulate ? dx[idx] + acc : acc;
}

// dx[m, c] += d_grad[m, c] * d/dx_c [J_pe(x)^T (ga + gb)]_c  (PE-6 of a 3-vector): the explicit x-dependence of
// SDFNetwork.gradient through the Jacobian of the encoding (field.py:158-170 differentiated once more)
__global__ void sdf_pe_hess_kernel(const float* __restrict__ x, const float* __restrict__ ga, int lda,
                                   const float* __restrict__ gb, int ldb, const float* __restrict__ d_grad, long long M,
                                   float* dx) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 3) return;
  const long long m = idx / 3;
  const int c = (int)(idx % 3);
  const float* a = ga + m * lda;
  const float* b = gb ? gb + m * ldb : nullptr;
  const float xc = x[idx];
  float acc = 0.f;
  for (int k = 0; k < 6; ++k) {
    const float f = (float)(1 << k);
    float sn, co;
    sincosf(xc * f, &sn, &co);
    const int is = 3 + 6 * k + c, ic = 6 + 6 * k + c;
    acc -= f * f * (sn * (a[is] + (b ? b[is] : 0.f)) + co * (a[ic] + (b ? b[ic] : 0.f)));
  }
  dx[idx] += d_grad[idx] * acc;
}

// backward of nerf_prep_kernel: pts4 = (p / |p|, 1 / |p|), views = -dirs
__global__ void nerf_prep_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ d_pts4,
                                     const float* __restrict__ d_views, long long M, float* d_pts, float* d_dirs) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float p[3] = {pts[3 * m], pts[3 * m + 1], pts[3 * m + 2]};
  const float n = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  const float q[3] = {p[0] / n, p[1] / n, p[2] / n};
  const float dq[3] = {d_pts4[4 * m], d_pts4[4 * m + 1], d_pts4[4 * m + 2]};
  const float dw = d_pts4[4 * m + 3];
  const float qdq = q[0] * dq[0] + q[1] * dq[1] + q[2] * dq[2];
  for (int c = 0; c < 3; ++c) {
    d_pts[3 * m + c] = (dq[c] - q[c] * qdq) / n - dw * q[c] / (n * n);
    d_dirs[3 * m + c] = -d_views[3 * m + c];
  }
}

// ------------------------------------------------------------------------------------------- sdf -> alpha
__global__ void sdf_alpha_fwd_kernel(nunerf_sdf_alpha_t p) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  float inv_s = fminf(fmaxf(p.inv_s_dev[0], 1e-6f), 1e6f);
  float g[3] = {p.grad[3 * m], p.grad[3 * m + 1], p.grad[3 * m + 2]};
  float dr[3] = {p.dirs[3 * m], p.dirs[3 * m + 1], p.dirs[3 * m + 2]};
  pw::SdfAlphaOut o = pw::sdf_alpha_fwd(p.sdf[m * p.ld_sdf], g, p.dists[m], dr, inv_s, p.cos_anneal);
  p.alpha[m] = o.alpha;
  p.grad_err[m] = o.gerr;
}

__global__ void sdf_alpha_bwd_kernel(nunerf_sdf_alpha_t p) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float dinv = 0.f;
  if (m < p.M) {
    float raw = p.inv_s_dev[0];
    float inv_s = fminf(fmaxf(raw, 1e-6f), 1e6f);
    float g[3] = {p.grad[3 * m], p.grad[3 * m + 1], p.grad[3 * m + 2]};
    float dr[3] = {p.dirs[3 * m], p.dirs[3 * m + 1], p.dirs[3 * m + 2]};
    float dsdf, dg[3], ddist, ddir[3];
    pw::sdf_alpha_bwd(p.sdf[m * p.ld_sdf], g, p.dists[m], dr, inv_s, p.cos_anneal, p.d_alpha[m],
                      p.d_grad_err ? p.d_grad_err[m] : 0.f, &dsdf, dg, &dinv, &ddist, ddir);
    if (p.d_dists) p.d_dists[m] = ddist;
    if (p.d_dirs) { p.d_dirs[3 * m] = ddir[0]; p.d_dirs[3 * m + 1] = ddir[1]; p.d_dirs[3 * m + 2] = ddir[2]; }
    if (!(raw >= 1e-6f && raw <= 1e6f)) dinv = 0.f;
    p.d_sdf[m] = dsdf;
    p.d_grad[3 * m] = dg[0]; p.d_grad[3 * m + 1] = dg[1]; p.d_grad[3 * m + 2] = dg[2];
  }
  if (p.d_inv_s) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dinv += __shfl_xor_sync(0xffffffffu, dinv, off);
    if ((threadIdx.x & 31) == 0 && dinv != 0.f) atomicAdd(p.d_inv_s, dinv);
  }
}

// ------------------------------------------------------------------------------------------- NeRF++ glue
__global__ void nerf_prep_kernel(const float* __restrict__ pts, const float* __restrict__ dirs, long long M,
                                 float* pts4, float* views) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float x = pts[3 * m], y = pts[3 * m + 1], z = pts[3 * m + 2];
  float n = sqrtf(x * x + y * y + z * z);  // ZT:688-689
  pts4[4 * m] = x / n; pts4[4 * m + 1] = y / n; pts4[4 * m + 2] = z / n; pts4[4 * m + 3] = 1.0f / n;
  views[3 * m] = -dirs[3 * m]; views[3 * m + 1] = -dirs[3 * m + 1]; views[3 * m + 2] = -dirs[3 * m + 2];
}

__global__ void nerf_out_fwd_kernel(const float* __restrict__ sigma, int ld_s, const float* __restrict__ rgb, int ld_c,
                                    const float* __restrict__ dists, long long M, float* alpha, float* color) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float c[3] = {rgb[m * ld_c], rgb[m * ld_c + 1], rgb[m * ld_c + 2]};
  float a, col[3];
  pw::nerf_out_fwd(sigma[m * ld_s], c, dists[m], &a, col);
  alpha[m] = a;
  color[3 * m] = col[0]; color[3 * m + 1] = col[1]; color[3 * m + 2] = col[2];
}

__global__ void nerf_out_bwd_kernel(const float* __restrict__ sigma, int ld_s, const float* __restrict__ rgb, int ld_c,
                                    const float* __restrict__ dists, long long M, const float* __restrict__ d_alpha,
                                    const float* __restrict__ d_color, __nv_bfloat16* d_sig, int ld_ds, int lo_ds,
                                    int col_ds, __nv_bfloat16* d_rgb, int ld_dr, int lo_dr, int col_dr, float* d_dists) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float c[3] = {rgb[m * ld_c], rgb[m * ld_c + 1], rgb[m * ld_c + 2]};
  float dc[3] = {d_color[3 * m], d_color[3 * m + 1], d_color[3 * m + 2]};
  float ds, drgb[3], dd;
  pw::nerf_out_bwd(sigma[m * ld_s], c, dists[m], d_alpha[m], dc, &ds, drgb, &dd);
  if (d_dists) d_dists[m] = dd;
  store_planes(d_sig, m * ld_ds + col_ds, lo_ds, ds);
  for (int k = 0; k < 3; ++k) store_planes(d_rgb, m * ld_dr + col_dr + k, lo_dr, drgb[k]);
}

// ------------------------------------------------------------------------------------------- shading encode
// One thread per (point, job): job 0: IDE(n,1) -> x_outer; 1: IDE(r,rough) -> x_outer + [PE6(p) | IDE] -> x_inner;
// 2: same at roughness 0; 3: [PE6(p) | PE6(r)] -> x_weight, [PE6(p) | PE6(v)] -> x_refrac, NoV.
// Every row is written whole (128 columns, zero padded), so the operand buffers need no prior zero fill.
__global__ void __launch_bounds__(128) shade_encode_fwd_kernel(nunerf_shade_encode_t p) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = p.M;
  if (idx >= 4 * M) return;
  int job = (int)(idx / M);
  long long m = idx % M;
  float g[3] = {p.grad[3 * m], p.grad[3 * m + 1], p.grad[3 * m + 2]};
  float rd[3] = {p.dirs[3 * m], p.dirs[3 * m + 1], p.dirs[3 * m + 2]};
  float pt[3] = {p.pts[3 * m], p.pts[3 * m + 1], p.pts[3 * m + 2]};
  pw::ShadeDirs s = pw::shade_dirs(g, rd);
  float row[128];
  if (job < 3) {
    __nv_bfloat16* xo = (__nv_bfloat16*)p.x_outer;
    __nv_bfloat16* xi = (__nv_bfloat16*)p.x_inner;
    float* out = row + 40;     // IDE block, 16-byte aligned inside the row buffer
    if (job == 0) pw::ide_fwd(c_ide, s.n[0], s.n[1], s.n[2], 1.0f, out);
    else pw::ide_fwd(c_ide, s.r[0], s.r[1], s.r[2], job == 1 ? pw::sigmoidf_(p.rough_raw[m * p.ld_rough]) : 0.0f, out);
#pragma unroll
    for (int j = 112; j < 128; ++j) row[j] = 0.f;
    const long long ro = ((long long)job * M + m) * p.ld_outer;
#pragma unroll
    for (int c = 0; c < 4; ++c) store16(xo, ro + c * 16, p.lo_outer, out + c * 16);
    store8(xo, ro + 64, p.lo_outer, out + 64);
    store8(xo, ro + 72, p.lo_outer, row + 112);
#pragma unroll
    for (int c = 5; c < 8; ++c) store16(xo, ro + c * 16, p.lo_outer, row + 112);
    if (job >= 1) {
      // inner input row = [PE6(p) (39) | IDE (72) | 0]: shift the IDE block down by one column
#pragma unroll
      for (int j = 39; j < 111; ++j) row[j] = row[j + 1];
      row[111] = 0.f;
      fill_pe<3, 6>(row, pt);
      const long long ri = ((long long)(job - 1) * M + m) * p.ld_inner;
#pragma unroll
      for (int c = 0; c < 8; ++c) store16(xi, ri + c * 16, p.lo_inner, row + c * 16);
    }
  } else {
    __nv_bfloat16* xw = (__nv_bfloat16*)p.x_weight;
    __nv_bfloat16* xr = (__nv_bfloat16*)p.x_refrac;
    fill_pe<3, 6>(row, pt);
    fill_pe<3, 6>(row + 39, s.r);
#pragma unroll
    for (int j = 78; j < 128; ++j) row[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) store16(xw, m * p.ld_weight + c * 16, p.lo_weight, row + c * 16);
    fill_pe<3, 6>(row + 39, s.v);
#pragma unroll
    for (int c = 0; c < 8; ++c) store16(xr, m * p.ld_refrac + c * 16, p.lo_refrac, row + c * 16);
    p.nov[m] = s.nov;
    if (p.refl) { p.refl[3 * m] = s.r[0]; p.refl[3 * m + 1] = s.r[1]; p.refl[3 * m + 2] = s.r[2]; }
  }
}

__global__ void shade_encode_bwd_kernel(nunerf_shade_encode_t p) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = p.M;
  if (m >= M) return;
  float g[3] = {p.grad[3 * m], p.grad[3 * m + 1], p.grad[3 * m + 2]};
  float rd[3] = {p.dirs[3 * m], p.dirs[3 * m + 1], p.dirs[3 * m + 2]};
  pw::ShadeDirs s = pw::shade_dirs(g, rd);
  float rough = pw::sigmoidf_(p.rough_raw[m * p.ld_rough]);
  float dn[3] = {0.f, 0.f, 0.f}, dr[3] = {0.f, 0.f, 0.f}, drough = 0.f;
  float dout[72], gx, gy, gz, gk;
  // every thread walks its own 512-byte rows: 16-byte loads (4 x fewer L1 requests than scalar ones).  The IDE block of an
  // inner-light row starts at column 39: aligned vectors over columns [36, 112), shifted by 3 in registers
  const bool vec = ((p.ld_dxo | p.ld_dxi) & 3) == 0 && (((uintptr_t)p.d_x_outer | (uintptr_t)p.d_x_inner) & 15) == 0;
  auto load_outer = [&](long long row) {
    const float* src = p.d_x_outer + row * p.ld_dxo;
    if (vec) {
#pragma unroll
      for (int k = 0; k < 18; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + k);
        dout[4 * k] = v.x; dout[4 * k + 1] = v.y; dout[4 * k + 2] = v.z; dout[4 * k + 3] = v.w;
      }
    } else {
      for (int j = 0; j < 72; ++j) dout[j] = src[j];
    }
  };
  auto add_inner = [&](long long row) {
    const float* src2 = p.d_x_inner + row * p.ld_dxi;
    if (vec) {
#pragma unroll
      for (int k = 9; k < 28; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src2) + k);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * k + i - 39;
          if (j >= 0 && j < 72) dout[j] += e[i];
        }
      }
    } else {
      for (int j = 0; j < 72; ++j) dout[j] += src2[39 + j];
    }
  };
  // IDE(n, 1)
  load_outer(m);
  pw::ide_bwd(c_ide, s.n[0], s.n[1], s.n[2], 1.0f, dout, &gx, &gy, &gz, &gk);
  dn[0] += gx; dn[1] += gy; dn[2] += gz;
  // IDE(r, rough): outer block 1 + inner block 0
  load_outer(M + m);
  add_inner(m);
  pw::ide_bwd(c_ide, s.r[0], s.r[1], s.r[2], rough, dout, &gx, &gy, &gz, &gk);
  dr[0] += gx; dr[1] += gy; dr[2] += gz; drough += gk;
  // IDE(r, 0): outer block 2 + inner block 1
  load_outer(2 * M + m);
  add_inner(M + m);
  pw::ide_bwd(c_ide, s.r[0], s.r[1], s.r[2], 0.0f, dout, &gx, &gy, &gz, &gk);
  dr[0] += gx; dr[1] += gy; dr[2] += gz;
  float dg[3];
  if (p.d_pts) {
    // position-gradient outputs (stage 2): d p through PE6(p) of both inner-light inputs and of the refraction-light input;
    // d ray direction through the reflected direction, NoV and PE6(v) of the refraction-light input (the occlusion weight
    // network sees detached encodings, field.py:657)
    const float* gi0 = p.d_x_inner + m * p.ld_dxi;
    const float* gi1 = p.d_x_inner + (M + m) * p.ld_dxi;
    const float* gr = p.d_x_refrac ? p.d_x_refrac + m * p.ld_dxr : nullptr;
    float gpe[39];
    for (int j = 0; j < 39; ++j) gpe[j] = gi0[j] + gi1[j] + (gr ? gr[j] : 0.f);
    float dvd[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < 3; ++c) {
      p.d_pts[3 * m + c] += pw::pe_bwd_coord(gpe, 3, 6, c, p.pts[3 * m + c]);
      if (gr) dvd[c] = pw::pe_bwd_coord(gr + 39, 3, 6, c, s.v[c]);
    }
    const float vn = fmaxf(sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]), 1e-12f);
    float drd[3];
    pw::shade_dirs_bwd(s, dr, dn, p.d_nov[m], dg, dvd, vn, drd);
    p.d_dirs[3 * m] += drd[0]; p.d_dirs[3 * m + 1] += drd[1]; p.d_dirs[3 * m + 2] += drd[2];
  } else {
    pw::shade_dirs_bwd(s, dr, dn, p.d_nov[m], dg);
  }
  p.d_grad[3 * m] += dg[0]; p.d_grad[3 * m + 1] += dg[1]; p.d_grad[3 * m + 2] += dg[2];
  p.d_rough_raw[m * p.ld_drough] += drough * rough * (1.0f - rough);
}

// ---- the same two kernels for the other shading-network variants.  Kept apart from the (6, 6) kernels above, which are on
// the measured stage-1 path.
//   PF / RF: light_pos_freq / refrac_freq (AppShadingNetwork_SpecInner, field.py:1320-1330: 8 / 2).  Row layouts:
//     x_inner = [PE_PF(p) | IDE | 0], x_weight = [PE_PF(p) | PE6(r) | 0], x_refrac = [PE_RF(p) | PE_RF(v) | 0], 128 columns.
//   SPH: the `sphere_direction` variant (field.py:594-597, :641-651, :675-680): the outer-light input is 144 wide,
//     x_outer = [IDE(u, k) | IDE(q(p, u), k') | 0] in 192-column rows, q = where the ray (p, u) leaves the unit sphere
//     (pw::sphere_dir_fwd); (u, k, k') = (n, 1, 1) | (r, rough, rough) | (r, 0, rough) for the three row blocks.
template <int PF, int RF, bool SPH>
__global__ void __launch_bounds__(128) shade_encode_fwd_var_kernel(nunerf_shade_encode_t p) {
  constexpr int PD = 3 + 6 * PF, RD = 3 + 6 * RF, OW = SPH ? 192 : 128;
  static_assert(PD + 72 <= 128 && PD + 39 <= 128 && 2 * RD <= 128, "row layout");
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = p.M;
  if (idx >= 4 * M) return;
  int job = (int)(idx / M);
  long long m = idx % M;
  float g[3] = {p.grad[3 * m], p.grad[3 * m + 1], p.grad[3 * m + 2]};
  float rd[3] = {p.dirs[3 * m], p.dirs[3 * m + 1], p.dirs[3 * m + 2]};
  float pt[3] = {p.pts[3 * m], p.pts[3 * m + 1], p.pts[3 * m + 2]};
  pw::ShadeDirs s = pw::shade_dirs(g, rd);
  float row[128];
  if (job < 3) {
    __nv_bfloat16* xo = (__nv_bfloat16*)p.x_outer;
    __nv_bfloat16* xi = (__nv_bfloat16*)p.x_inner;
    float orow[OW];
    const float* u = job == 0 ? s.n : s.r;
    const float rough = job == 0 ? 1.0f : pw::sigmoidf_(p.rough_raw[m * p.ld_rough]);
    pw::ide_fwd(c_ide, u[0], u[1], u[2], job == 2 ? 0.0f : rough, orow);
    if (SPH) {
      float q[3];
      pw::sphere_dir_fwd(pt, u, q);
      pw::ide_fwd(c_ide, q[0], q[1], q[2], rough, orow + 72);
    }
#pragma unroll
    for (int j = SPH ? 144 : 72; j < OW; ++j) orow[j] = 0.f;
    const long long ro = ((long long)job * M + m) * p.ld_outer;
#pragma unroll
    for (int c = 0; c < OW / 16; ++c) store16(xo, ro + c * 16, p.lo_outer, orow + c * 16);
    if (job >= 1) {
#pragma unroll
      for (int j = 0; j < 72; ++j) row[PD + j] = orow[j];
#pragma unroll
      for (int j = PD + 72; j < 128; ++j) row[j] = 0.f;
      fill_pe<3, PF>(row, pt);
      const long long ri = ((long long)(job - 1) * M + m) * p.ld_inner;
#pragma unroll
      for (int c = 0; c < 8; ++c) store16(xi, ri + c * 16, p.lo_inner, row + c * 16);
    }
  } else {
    __nv_bfloat16* xw = (__nv_bfloat16*)p.x_weight;
    __nv_bfloat16* xr = (__nv_bfloat16*)p.x_refrac;
    fill_pe<3, PF>(row, pt);
    fill_pe<3, 6>(row + PD, s.r);
#pragma unroll
    for (int j = PD + 39; j < 128; ++j) row[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) store16(xw, m * p.ld_weight + c * 16, p.lo_weight, row + c * 16);
    fill_pe<3, RF>(row, pt);
    fill_pe<3, RF>(row + RD, s.v);
#pragma unroll
    for (int j = 2 * RD; j < 128; ++j) row[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) store16(xr, m * p.ld_refrac + c * 16, p.lo_refrac, row + c * 16);
    p.nov[m] = s.nov;
    if (p.refl) { p.refl[3 * m] = s.r[0]; p.refl[3 * m + 1] = s.r[1]; p.refl[3 * m + 2] = s.r[2]; }
  }
}

// dout[j] (+)= row[COL0 + j], j < 72: every thread walks its own row, so 16-byte loads (4 x fewer L1 requests than scalar
// ones; COL0 need not be a multiple of 4: aligned vectors over [COL0 & ~3, COL0 + 72) with the shift done in registers)
template <int COL0, bool ADD>
__device__ __forceinline__ void load72(const float* __restrict__ row, float* dout, bool vec) {
  if (vec) {
#pragma unroll
    for (int k = COL0 / 4; k <= (COL0 + 71) / 4; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row) + k);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = 4 * k + i - COL0;
        if (j >= 0 && j < 72) dout[j] = ADD ? dout[j] + e[i] : e[i];
      }
    }
  } else {
    for (int j = 0; j < 72; ++j) dout[j] = ADD ? dout[j] + row[COL0 + j] : row[COL0 + j];
  }
}

template <int PF, int RF, bool SPH>
__global__ void shade_encode_bwd_var_kernel(nunerf_shade_encode_t p) {
  constexpr int PD = 3 + 6 * PF, RD = 3 + 6 * RF;
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = p.M;
  if (m >= M) return;
  float g[3] = {p.grad[3 * m], p.grad[3 * m + 1], p.grad[3 * m + 2]};
  float rd[3] = {p.dirs[3 * m], p.dirs[3 * m + 1], p.dirs[3 * m + 2]};
  float pt[3] = {p.pts[3 * m], p.pts[3 * m + 1], p.pts[3 * m + 2]};
  pw::ShadeDirs s = pw::shade_dirs(g, rd);
  float rough = pw::sigmoidf_(p.rough_raw[m * p.ld_rough]);
  float dn[3] = {0.f, 0.f, 0.f}, dr[3] = {0.f, 0.f, 0.f}, dpt[3] = {0.f, 0.f, 0.f}, drough = 0.f;
  float dout[72], gx, gy, gz, gk;
  // one row block: IDE(u, k) gradient = outer-light columns 0..71 (+ inner-light columns PD..PD+71), and with SPH the
  // second IDE block (columns 72..143) through the sphere direction q(p, u)
  const bool vec = ((p.ld_dxo | p.ld_dxi) & 3) == 0 && (((uintptr_t)p.d_x_outer | (uintptr_t)p.d_x_inner) & 15) == 0;
  auto block = [&](long long orow, long long irow, const float* u, float k, bool k_is_rough, float* du) {
    const float* src = p.d_x_outer + orow * p.ld_dxo;
    load72<0, false>(src, dout, vec);
    if (irow >= 0) load72<PD, true>(p.d_x_inner + irow * p.ld_dxi, dout, vec);
    pw::ide_bwd(c_ide, u[0], u[1], u[2], k, dout, &gx, &gy, &gz, &gk);
    du[0] += gx; du[1] += gy; du[2] += gz;
    if (k_is_rough) drough += gk;
    if (SPH) {
      load72<72, false>(src, dout, vec);
      float q[3], dq[3];
      pw::sphere_dir_fwd(pt, u, q);
      const bool rough_k = irow >= 0;               // blocks 1 and 2 encode q at the predicted roughness, block 0 at 1
      pw::ide_bwd(c_ide, q[0], q[1], q[2], rough_k ? rough : 1.0f, dout, &dq[0], &dq[1], &dq[2], &gk);
      if (rough_k) drough += gk;
      pw::sphere_dir_bwd(pt, u, dq, du, p.d_pts ? dpt : nullptr);
    }
  };
  block(m, -1, s.n, 1.0f, false, dn);
  block(M + m, m, s.r, rough, true, dr);
  block(2 * M + m, M + m, s.r, 0.0f, false, dr);
  float dg[3];
  if (p.d_pts) {
    const float* gi0 = p.d_x_inner + m * p.ld_dxi;
    const float* gi1 = p.d_x_inner + (M + m) * p.ld_dxi;
    const float* gr = p.d_x_refrac ? p.d_x_refrac + m * p.ld_dxr : nullptr;
    float gpe[PD];
    for (int j = 0; j < PD; ++j) gpe[j] = gi0[j] + gi1[j];
    float dvd[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < 3; ++c) {
      float acc = dpt[c] + pw::pe_bwd_coord(gpe, 3, PF, c, pt[c]);
      if (gr) {
        acc += pw::pe_bwd_coord(gr, 3, RF, c, pt[c]);
        dvd[c] = pw::pe_bwd_coord(gr + RD, 3, RF, c, s.v[c]);
      }
      p.d_pts[3 * m + c] += acc;
    }
    const float vn = fmaxf(sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]), 1e-12f);
    float drd[3];
    pw::shade_dirs_bwd(s, dr, dn, p.d_nov[m], dg, dvd, vn, drd);
    p.d_dirs[3 * m] += drd[0]; p.d_dirs[3 * m + 1] += drd[1]; p.d_dirs[3 * m + 2] += drd[2];
  } else {
    pw::shade_dirs_bwd(s, dr, dn, p.d_nov[m], dg);
  }
  p.d_grad[3 * m] += dg[0]; p.d_grad[3 * m + 1] += dg[1]; p.d_grad[3 * m + 2] += dg[2];
  p.d_rough_raw[m * p.ld_drough] += drough * rough * (1.0f - rough);
}

// IDE of M unit directions at a constant roughness (the per-ray specular probe, ZT:780); writes 128 columns
// (72 real + zero padding) starting at the 8-aligned column `col`.
__global__ void __launch_bounds__(128)
ide_encode_kernel(const float* __restrict__ x, long long M, float kinv, __nv_bfloat16* dst, int ld, int lo, int col) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float out[72];
  pw::ide_fwd(c_ide, x[3 * m], x[3 * m + 1], x[3 * m + 2], kinv, out);
  const float zero[16] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 4; ++c) store16(dst, m * ld + col + c * 16, lo, out + c * 16);
  store8(dst, m * ld + col + 64, lo, out + 64);
  store8(dst, m * ld + col + 72, lo, zero);
#pragma unroll
  for (int c = 5; c < 8; ++c) store16(dst, m * ld + col + c * 16, lo, zero);
}

// ------------------------------------------------------------------------------------------- shading mix
__device__ __forceinline__ void load_mix(const nunerf_shade_mix_t& p, long long m, pw::ShadeMixIn* in) {
  long long M = p.M;
  in->metallic = p.metallic[m * p.ld_mat];
  in->rough = p.rough[m * p.ld_mat];
  in->trans = p.trans[m * p.ld_mat];
  in->occ = p.weight[m * p.ld_weight];
  in->nov = p.nov[m];
  for (int c = 0; c < 3; ++c) {
    in->albedo[c] = p.albedo[m * p.ld_mat + c];
    in->diffuse_l[c] = p.outer[m * p.ld_outer + c];
    in->direct[c] = p.outer[(M + m) * p.ld_outer + c];
    in->direct0[c] = p.outer[(2 * M + m) * p.ld_outer + c];
    in->indirect[c] = p.inner[m * p.ld_inner + c];
    in->indirect0[c] = p.inner[(M + m) * p.ld_inner + c];
    in->refrac[c] = p.refrac[m * p.ld_refrac + c];
  }
}

__global__ void shade_mix_fwd_kernel(nunerf_shade_mix_t p) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  pw::ShadeMixIn in;
  load_mix(p, m, &in);
  pw::ShadeMixOut o = pw::shade_mix_fwd(in, p.lut, p.exp_max, p.use_exp_max_refrac ? p.exp_max_refrac : p.exp_max);
  p.color[3 * m] = o.color[0]; p.color[3 * m + 1] = o.color[1]; p.color[3 * m + 2] = o.color[2];
  p.trans_out[m] = o.trans; p.metallic_out[m] = o.metallic; p.occ_prob[m] = o.occ_prob;
}

// gradients w.r.t. the raw heads are written as bf16 planes (the dZ operands of the head layers, 64-wide,
// zero padded) plus fp32 d_nov / d_rough_raw for the encode backward.
__global__ void shade_mix_bwd_kernel(nunerf_shade_mix_t p) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long M = p.M;
  if (m >= M) return;
  pw::ShadeMixIn in, d;
  load_mix(p, m, &in);
  float dc[3] = {p.d_color[3 * m], p.d_color[3 * m + 1], p.d_color[3 * m + 2]};
  pw::shade_mix_bwd(in, p.lut, p.exp_max, p.use_exp_max_refrac ? p.exp_max_refrac : p.exp_max, dc,
                    p.d_trans_out ? p.d_trans_out[m] : 0.f,
                    p.d_metallic_out ? p.d_metallic_out[m] : 0.f, &d);
  if (p.d_occ_prob) d.occ += 0.5f * p.d_occ_prob[m];   // occ_prob = raw * 0.5 + 0.5 (field.py:659), unclipped output
  const int ld = p.ld_dz, lo = p.lo_dz;
  // every head row is written WHOLE (64 columns: up to 3 values + zero padding) with 32-byte stores, so the dZ operand
  // buffers need no prior zero fill and no 2-byte partial-sector writes
  auto row64 = [&](void* dst, long long row, float a, float b, float c) {
    float v[16] = {a, b, c, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float z[16] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    __nv_bfloat16* d_ = (__nv_bfloat16*)dst;
    store16(d_, row * ld, lo, v);
    if (ld >= 64) { store16(d_, row * ld + 16, lo, z); store16(d_, row * ld + 32, lo, z); store16(d_, row * ld + 48, lo, z); }
  };
  row64(p.dz_metallic, m, d.metallic, 0.f, 0.f);
  row64(p.dz_trans, m, d.trans, 0.f, 0.f);
  row64(p.dz_weight, m, d.occ, 0.f, 0.f);
  row64(p.dz_albedo, m, d.albedo[0], d.albedo[1], d.albedo[2]);
  row64(p.dz_outer, m, d.diffuse_l[0], d.diffuse_l[1], d.diffuse_l[2]);
  row64(p.dz_outer, M + m, d.direct[0], d.direct[1], d.direct[2]);
  row64(p.dz_outer, 2 * M + m, d.direct0[0], d.direct0[1], d.direct0[2]);
  row64(p.dz_inner, m, d.indirect[0], d.indirect[1], d.indirect[2]);
  row64(p.dz_inner, M + m, d.indirect0[0], d.indirect0[1], d.indirect0[2]);
  row64(p.dz_refrac, m, d.refrac[0], d.refrac[1], d.refrac[2]);
  p.d_rough_raw[m] = d.rough;  // completed (and turned into planes) after shade_encode_bwd adds its share
  p.d_nov[m] = d.nov;
}

// ------------------------------------------------------------------------------------------- SDF-net glue
// out[m,n] = w[n] * (1 - exp(-100 a[m,n]))          (gs_7 = w_sdf . s_7)
// 8 consecutive columns of bf16 planes <-> fp32 (16-byte accesses; callers guarantee 8-column alignment)
__device__ __forceinline__ void load8_planes(const __nv_bfloat16* base, long long idx, int lo, float* v) {
  const uint4 h = *reinterpret_cast<const uint4*>(base + idx);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[2 * j] = bf16lo_to_f(hw[j]); v[2 * j + 1] = bf16hi_to_f(hw[j]); }
  if (lo) {
    const uint4 l = *reinterpret_cast<const uint4*>(base + idx + lo);
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[2 * j] += bf16lo_to_f(lw[j]); v[2 * j + 1] += bf16hi_to_f(lw[j]); }
  }
}

// one thread per 8 columns
__global__ void rowvec_mask_kernel(const float* __restrict__ w, const __nv_bfloat16* __restrict__ a, int lda, int a_lo,
                                   long long M, int N, __nv_bfloat16* out, int ldo, int o_lo) {
  const int groups = N >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * groups) return;
  const long long m = idx / groups;
  const int n = (int)(idx % groups) << 3;
  float av[8], r[8];
  load8_planes(a, m * lda + n, a_lo, av);
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = w[n + j] * (1.0f - __expf(-100.0f * av[j]));
  store8(out, m * ldo + n, o_lo, r);
}
// u4 [M,256] fp32 = d sdf / d (input of lin4): columns < 217 -> gs_3 = u . s_3 (planes), columns 217.. -> g_skip (the PE part)
__global__ void sdf_skip_split_kernel(const float* __restrict__ u4, const __nv_bfloat16* __restrict__ a3, int lda,
                                      int a_lo, long long M, __nv_bfloat16* gs3, int ldg, int g_lo, float* g_skip) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 32) return;
  const long long m = idx >> 5;
  const int n = (int)(idx & 31) << 3;
  const float4 u0 = *reinterpret_cast<const float4*>(u4 + m * 256 + n);
  const float4 u1 = *reinterpret_cast<const float4*>(u4 + m * 256 + n + 4);
  const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
  float av[8], r[8];
  load8_planes(a3, m * lda + n, a_lo, av);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = n + j;
    if (col < 217) r[j] = u[j] * (1.0f - __expf(-100.0f * av[j]));
    else { r[j] = 0.f; g_skip[m * 39 + (col - 217)] = u[j]; }
  }
  store8(gs3, m * ldg + n, g_lo, r);
}

// reverse-over-reverse glue of one softplus layer:
//   u_next = gts . s ;  e = gts . gs . 100 (1 - s)      with s = 1 - exp(-100 a)
__global__ void sdf_bwd2_ew_kernel(const __nv_bfloat16* __restrict__ gts, int ldt, int t_lo,
                                   const __nv_bfloat16* __restrict__ a, int lda, int a_lo,
                                   const __nv_bfloat16* __restrict__ gs, int ldg, int g_lo, long long M, int N,
                                   int n_real, __nv_bfloat16* u_next, int ldu, int u_lo, __nv_bfloat16* e, int lde,
                                   int e_lo) {
  // 8 consecutive columns per thread, 16-byte loads / stores
  const int n8 = N >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * n8) return;
  long long m = idx / n8;
  int n0 = (int)(idx % n8) << 3;
  auto ld8 = [](const __nv_bfloat16* base, long long off, int lo, float* f) {
    uint4 h = *reinterpret_cast<const uint4*>(base + off);
    const uint32_t w[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = bf16lo_to_f(w[j]); f[2 * j + 1] = bf16hi_to_f(w[j]); }
    if (lo) {
      uint4 l = *reinterpret_cast<const uint4*>(base + off + lo);
      const uint32_t v[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { f[2 * j] += bf16lo_to_f(v[j]); f[2 * j + 1] += bf16hi_to_f(v[j]); }
    }
  };
  auto st8 = [](__nv_bfloat16* base, long long off, int lo, const float* f) {
    uint32_t hw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) hw[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (lo) {
      uint32_t lw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        lw[j] = pack_bf16x2(f[2 * j] - bf16lo_to_f(hw[j]), f[2 * j + 1] - bf16hi_to_f(hw[j]));
      *reinterpret_cast<uint4*>(base + off + lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  };
  float t[8], av[8], g[8], un[8], ev[8];
  ld8(gts, m * ldt + n0, t_lo, t);
  ld8(a, m * lda + n0, a_lo, av);
  ld8(gs, m * ldg + n0, g_lo, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float s = 1.0f - __expf(-100.0f * av[j]);
    bool real = (n0 + j) < n_real;
    un[j] = real ? t[j] * s : 0.f;
    ev[j] = real ? t[j] * g[j] * 100.0f * (1.0f - s) : 0.f;
  }
  if (u_next) {
    if (n0 + 8 <= n_real) st8(u_next, m * ldu + n0, u_lo, un);
    else
      for (int j = 0; j < 8 && n0 + j < n_real; ++j) store_planes(u_next, m * ldu + n0 + j, u_lo, un[j]);
  }
  st8(e, m * lde + n0, e_lo, ev);
}

// fp32 [M, C] (+ optional second addend) -> planes at a column offset
__global__ void f32_to_planes_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                     long long M, int C, int width, __nv_bfloat16* dst, int ld, int lo, int col) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * width) return;
  long long m = idx / width;
  int c = (int)(idx % width);
  float v = 0.f;
  if (c < C) v = a[m * lda + c] + (b ? b[m * ldb + c] : 0.f);
  store_planes(dst, m * ld + col + c, lo, v);
}

__global__ void f32_to_planes8_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, long long M,
                                      int C, int groups, __nv_bfloat16* dst, int ld, int lo, int col) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * groups) return;
  const long long m = idx / groups;
  const int c0 = (int)(idx % groups) << 3;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    v[j] = c < C ? a[m * lda + c] + (b ? b[m * ldb + c] : 0.f) : 0.f;
  }
  store8(dst, m * ld + col + c0, lo, v);
}

__global__ void adam_kernel(float* p, const float* __restrict__ g, float* m, float* v, long long n, float lr, float b1,
                            float b2, float eps, float bc1, float bc2) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  float mi = b1 * m[i] + (1.0f - b1) * gi;
  float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  // torch.optim.Adam: p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
  p[i] -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
}

// ------------------------------------------------------------------------------------------- grid sweep
__global__ void grid_points_kernel(int res, long long start, long long count, const float* __restrict__ lin, float* pts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  long long g = start + i;
  int zi = (int)(g % res), yi = (int)((g / res) % res), xi = (int)(g / ((long long)res * res));
  pts[3 * i] = lin[xi]; pts[3 * i + 1] = lin[res + yi]; pts[3 * i + 2] = lin[2 * res + zi];   // per-axis tables
}
__global__ void grid_mask_kernel(const float* __restrict__ pts, const float* __restrict__ sdf, int ld, long long count,
                                 float outside, float* u) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  float n = sqrtf(x * x + y * y + z * z);
  u[i] = n >= 1.0f ? outside : sdf[i * ld];
}

}  // namespace nunerf

using namespace nunerf;
#define ST(s) ((cudaStream_t)(s))
#define G1(n) cdiv((n), 256), 256

extern "C" int nunerf_encode_pe(const float* x, int M, int d, int nfreq, void* dst, int ld, int lo_off, int col_off,
                                int row_off, int width, void* stream) {
  NUNERF_REQUIRE(x && dst && M > 0 && d > 0 && nfreq >= 0 && width >= d * (1 + 2 * nfreq), "encode_pe: bad arguments");
  const int grid = cdiv(M, 128);
  __nv_bfloat16* o = (__nv_bfloat16*)dst;
  if (d == 3 && nfreq == 6) encode_pe_kernel<3, 6><<<grid, 128, 0, ST(stream)>>>(x, M, o, ld, lo_off, col_off, row_off, width);
  else if (d == 4 && nfreq == 10) encode_pe_kernel<4, 10><<<grid, 128, 0, ST(stream)>>>(x, M, o, ld, lo_off, col_off, row_off, width);
  else if (d == 3 && nfreq == 4) encode_pe_kernel<3, 4><<<grid, 128, 0, ST(stream)>>>(x, M, o, ld, lo_off, col_off, row_off, width);
  else return fail("%s", "encode_pe: supported (d, nfreq) are (3,6), (4,10), (3,4)");
  NUNERF_CHECK_LAUNCH("encode_pe_kernel");
  return 0;
}

extern "C" int nunerf_sdf_grad_pe(const float* x, const float* ga, int lda, const float* gb, int ldb, int M,
                                  float* grad, void* stream) {
  NUNERF_REQUIRE(x && ga && grad && M > 0, "sdf_grad_pe: bad arguments");
  sdf_grad_pe_kernel<<<G1(M), 0, ST(stream)>>>(x, ga, lda, gb, ldb, M, grad);
  NUNERF_CHECK_LAUNCH("sdf_grad_pe_kernel");
  return 0;
}

// one thread per row: the 39 values once, written as 32-byte stores (the 64-wide destination whole; the 39-wide one -- the
// skip columns 217..255 of a 256-wide row -- as 7 single columns up to the next 16-column boundary + two 32-byte stores)
__global__ void sdf_grad_pe_bwd_row_kernel(const float* __restrict__ x, const float* __restrict__ dgrad, long long M,
                                           __nv_bfloat16* d1, int ld1, int lo1, __nv_bfloat16* d2, int ld2, int lo2) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float v[64];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float g = dgrad[3 * m + c], xc = x[3 * m + c];
    v[c] = g;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float f = (float)(1 << k);
      float s, co;
      sincosf(xc * f, &s, &co);
      v[3 + 6 * k + c] = g * f * co;
      v[6 + 6 * k + c] = -g * f * s;
    }
  }
#pragma unroll
  for (int j = 39; j < 64; ++j) v[j] = 0.f;
  if (d1) {
#pragma unroll
    for (int c = 0; c < 4; ++c) store16(d1, m * ld1 + c * 16, lo1, v + c * 16);
  }
  if (d2) {
#pragma unroll
    for (int j = 0; j < 7; ++j) store_planes(d2, m * ld2 + 217 + j, lo2, v[j]);
    store16(d2, m * ld2 + 224, lo2, v + 7);
    store16(d2, m * ld2 + 240, lo2, v + 23);
  }
}

extern "C" int nunerf_sdf_grad_pe_bwd(const float* x, const float* dgrad, int M, void* d1, int ld1, int lo1, int col1,
                                      int width1, void* d2, int ld2, int lo2, int col2, int width2, void* stream) {
  NUNERF_REQUIRE(x && dgrad && M > 0 && (d1 || d2), "sdf_grad_pe_bwd: bad arguments");
  if ((!d1 || (col1 == 0 && width1 == 64 && (ld1 & 15) == 0 && (lo1 & 15) == 0 && (((uintptr_t)d1) & 31) == 0)) &&
      (!d2 || (col2 == 217 && width2 == 39 && (ld2 & 15) == 0 && (lo2 & 15) == 0 && (((uintptr_t)d2) & 31) == 0))) {
    sdf_grad_pe_bwd_row_kernel<<<G1(M), 0, ST(stream)>>>(x, dgrad, M, (__nv_bfloat16*)d1, ld1, lo1, (__nv_bfloat16*)d2, ld2,
                                                        lo2);
    NUNERF_CHECK_LAUNCH("sdf_grad_pe_bwd_row_kernel");
    return 0;
  }
  int wmax = width1 > width2 ? width1 : width2;
  long long total = (long long)M * wmax;
  sdf_grad_pe_bwd_kernel<<<G1(total), 0, ST(stream)>>>(x, dgrad, M, (__nv_bfloat16*)d1, ld1, lo1, col1, width1,
                                                      (__nv_bfloat16*)d2, ld2, lo2, col2, width2);
  NUNERF_CHECK_LAUNCH("sdf_grad_pe_bwd_kernel");
  return 0;
}

extern "C" int nunerf_pe_bwd(const float* x, int d, int nfreq, const float* ga, int lda, const float* gb, int ldb, int M,
                             float* dx, int accumulate, void* stream) {
  NUNERF_REQUIRE(x && ga && dx && M > 0 && d >= 1 && d <= 4 && nfreq >= 0 && nfreq <= 12 && lda >= d * (1 + 2 * nfreq) &&
                     (!gb || ldb >= d * (1 + 2 * nfreq)), "pe_bwd: bad arguments");
  pe_bwd_kernel<<<G1((long long)M * d), 0, ST(stream)>>>(x, d, nfreq, ga, lda, gb, ldb, M, dx, accumulate);
  NUNERF_CHECK_LAUNCH("pe_bwd_kernel");
  return 0;
}
extern "C" int nunerf_sdf_pe_hess(const float* x, const float* ga, int lda, const float* gb, int ldb, const float* d_grad,
                                  int M, float* dx, void* stream) {
  NUNERF_REQUIRE(x && ga && d_grad && dx && M > 0 && lda >= 39 && (!gb || ldb >= 39), "sdf_pe_hess: bad arguments");
  sdf_pe_hess_kernel<<<G1((long long)M * 3), 0, ST(stream)>>>(x, ga, lda, gb, ldb, d_grad, M, dx);
  NUNERF_CHECK_LAUNCH("sdf_pe_hess_kernel");
  return 0;
}
extern "C" int nunerf_nerf_prep_bwd(const float* pts, const float* d_pts4, const float* d_views, int M, float* d_pts,
                                    float* d_dirs, void* stream) {
  NUNERF_REQUIRE(pts && d_pts4 && d_views && d_pts && d_dirs && M > 0, "nerf_prep_bwd: bad arguments");
  nerf_prep_bwd_kernel<<<G1(M), 0, ST(stream)>>>(pts, d_pts4, d_views, M, d_pts, d_dirs);
  NUNERF_CHECK_LAUNCH("nerf_prep_bwd_kernel");
  return 0;
}

extern "C" int nunerf_sdf_alpha_fwd(const nunerf_sdf_alpha_t* p, void* stream) {
  NUNERF_REQUIRE(p && p->M > 0 && p->sdf && p->grad && p->dists && p->dirs && p->alpha && p->grad_err && p->inv_s_dev,
                 "sdf_alpha_fwd: bad arguments");
  sdf_alpha_fwd_kernel<<<G1(p->M), 0, ST(stream)>>>(*p);
  NUNERF_CHECK_LAUNCH("sdf_alpha_fwd_kernel");
  return 0;
}
extern "C" int nunerf_sdf_alpha_bwd(const nunerf_sdf_alpha_t* p, void* stream) {
  NUNERF_REQUIRE(p && p->M > 0 && p->sdf && p->grad && p->dists && p->dirs && p->d_alpha && p->d_sdf && p->d_grad &&
                     p->inv_s_dev,
                 "sdf_alpha_bwd: bad arguments");
  sdf_alpha_bwd_kernel<<<G1(p->M), 0, ST(stream)>>>(*p);
  NUNERF_CHECK_LAUNCH("sdf_alpha_bwd_kernel");
  return 0;
}

extern "C" int nunerf_nerf_prep(const float* pts, const float* dirs, int M, float* pts4, float* views, void* stream) {
  NUNERF_REQUIRE(pts && dirs && pts4 && views && M > 0, "nerf_prep: bad arguments");
  nerf_prep_kernel<<<G1(M), 0, ST(stream)>>>(pts, dirs, M, pts4, views);
  NUNERF_CHECK_LAUNCH("nerf_prep_kernel");
  return 0;
}
extern "C" int nunerf_nerf_out_fwd(const float* sigma, int ld_s, const float* rgb, int ld_c, const float* dists, int M,
                                   float* alpha, float* color, void* stream) {
  NUNERF_REQUIRE(sigma && rgb && dists && alpha && color && M > 0, "nerf_out_fwd: bad arguments");
  nerf_out_fwd_kernel<<<G1(M), 0, ST(stream)>>>(sigma, ld_s, rgb, ld_c, dists, M, alpha, color);
  NUNERF_CHECK_LAUNCH("nerf_out_fwd_kernel");
  return 0;
}
extern "C" int nunerf_nerf_out_bwd(const float* sigma, int ld_s, const float* rgb, int ld_c, const float* dists, int M,
                                   const float* d_alpha, const float* d_color, void* d_sig, int ld_ds, int lo_ds,
                                   int col_ds, void* d_rgb, int ld_dr, int lo_dr, int col_dr, void* stream) {
  NUNERF_REQUIRE(sigma && rgb && dists && d_alpha && d_color && d_sig && d_rgb && M > 0, "nerf_out_bwd: bad arguments");
  nerf_out_bwd_kernel<<<G1(M), 0, ST(stream)>>>(sigma, ld_s, rgb, ld_c, dists, M, d_alpha, d_color,
                                               (__nv_bfloat16*)d_sig, ld_ds, lo_ds, col_ds, (__nv_bfloat16*)d_rgb, ld_dr,
                                               lo_dr, col_dr, nullptr);
  NUNERF_CHECK_LAUNCH("nerf_out_bwd_kernel");
  return 0;
}
// same, additionally d alpha / d dist -> d_dists[M] (stage 2: the sample spacing depends on the refracted path)
extern "C" int nunerf_nerf_out_bwd_geo(const float* sigma, int ld_s, const float* rgb, int ld_c, const float* dists, int M,
                                       const float* d_alpha, const float* d_color, void* d_sig, int ld_ds, int lo_ds,
                                       int col_ds, void* d_rgb, int ld_dr, int lo_dr, int col_dr, float* d_dists,
                                       void* stream) {
  NUNERF_REQUIRE(sigma && rgb && dists && d_alpha && d_color && d_sig && d_rgb && d_dists && M > 0, "nerf_out_bwd_geo: bad arguments");
  nerf_out_bwd_kernel<<<G1(M), 0, ST(stream)>>>(sigma, ld_s, rgb, ld_c, dists, M, d_alpha, d_color,
                                               (__nv_bfloat16*)d_sig, ld_ds, lo_ds, col_ds, (__nv_bfloat16*)d_rgb, ld_dr,
                                               lo_dr, col_dr, d_dists);
  NUNERF_CHECK_LAUNCH("nerf_out_bwd_kernel");
  return 0;
}

extern "C" int nunerf_shade_encode_fwd(const nunerf_shade_encode_t* p, void* stream) {
  NUNERF_REQUIRE(p && p->M > 0 && p->pts && p->grad && p->dirs && p->rough_raw && p->x_outer && p->x_inner &&
                     p->x_weight && p->x_refrac && p->nov,
                 "shade_encode_fwd: bad arguments");
  if (int r = ensure_ide()) return r;
  NUNERF_REQUIRE(((p->ld_outer | p->lo_outer | p->ld_inner | p->lo_inner | p->ld_weight | p->lo_weight | p->ld_refrac |
                   p->lo_refrac) & 7) == 0 && p->ld_outer >= 128 && p->ld_inner >= 128 && p->ld_weight >= 128 &&
                     p->ld_refrac >= 128,
                 "shade_encode_fwd: operand rows must be >= 128 columns with 8-column aligned pitches");
  const int pf = p->pos_freq ? p->pos_freq : 6, rf = p->refrac_freq ? p->refrac_freq : 6, sph = p->sphere_direction;
  NUNERF_REQUIRE(p->ld_outer >= (sph ? 192 : 128), "shade_encode_fwd: sphere_direction needs 192-column x_outer rows");
  const int nb = (int)cdiv(4LL * p->M, 128);
  if (pf == 6 && rf == 6 && !sph) shade_encode_fwd_kernel<<<nb, 128, 0, ST(stream)>>>(*p);
  else if (pf == 6 && rf == 6) shade_encode_fwd_var_kernel<6, 6, true><<<nb, 128, 0, ST(stream)>>>(*p);
  else if (pf == 8 && rf == 2 && !sph) shade_encode_fwd_var_kernel<8, 2, false><<<nb, 128, 0, ST(stream)>>>(*p);
  else if (pf == 8 && rf == 2) shade_encode_fwd_var_kernel<8, 2, true><<<nb, 128, 0, ST(stream)>>>(*p);
  else return fail("%s", "shade_encode_fwd: supported (pos_freq, refrac_freq) are (6, 6) and (8, 2)");
  NUNERF_CHECK_LAUNCH("shade_encode_fwd_kernel");
  return 0;
}
extern "C" int nunerf_shade_encode_bwd(const nunerf_shade_encode_t* p, void* stream) {
  NUNERF_REQUIRE(p && p->M > 0 && p->grad && p->dirs && p->rough_raw && p->d_x_outer && p->d_x_inner && p->d_nov &&
                     p->d_grad && p->d_rough_raw,
                 "shade_encode_bwd: bad arguments");
  NUNERF_REQUIRE(!p->d_pts || (p->pts && p->d_dirs), "shade_encode_bwd: d_pts needs pts and d_dirs");
  if (int r = ensure_ide()) return r;
  const int pf = p->pos_freq ? p->pos_freq : 6, rf = p->refrac_freq ? p->refrac_freq : 6, sph = p->sphere_direction;
  NUNERF_REQUIRE(!sph || (p->ld_dxo >= 144 && p->pts), "shade_encode_bwd: sphere_direction needs 144 gradient columns and pts");
  const int nb = (int)cdiv(p->M, 128);
  if (pf == 6 && rf == 6 && !sph) shade_encode_bwd_kernel<<<nb, 128, 0, ST(stream)>>>(*p);
  else if (pf == 6 && rf == 6) shade_encode_bwd_var_kernel<6, 6, true><<<nb, 128, 0, ST(stream)>>>(*p);
  else if (pf == 8 && rf == 2 && !sph) shade_encode_bwd_var_kernel<8, 2, false><<<nb, 128, 0, ST(stream)>>>(*p);
  else if (pf == 8 && rf == 2) shade_encode_bwd_var_kernel<8, 2, true><<<nb, 128, 0, ST(stream)>>>(*p);
  else return fail("%s", "shade_encode_bwd: supported (pos_freq, refrac_freq) are (6, 6) and (8, 2)");
  NUNERF_CHECK_LAUNCH("shade_encode_bwd_kernel");
  return 0;
}
extern "C" int nunerf_ide_encode(const float* x, int M, float kinv, void* dst, int ld, int lo, int col, void* stream) {
  NUNERF_REQUIRE(x && dst && M > 0 && ((ld | lo | col) & 7) == 0 && ld >= col + 128, "ide_encode: bad arguments");
  if (int r = ensure_ide()) return r;
  ide_encode_kernel<<<cdiv(M, 128), 128, 0, ST(stream)>>>(x, M, kinv, (__nv_bfloat16*)dst, ld, lo, col);
  NUNERF_CHECK_LAUNCH("ide_encode_kernel");
  return 0;
}
extern "C" int nunerf_shade_mix_fwd(const nunerf_shade_mix_t* p, void* stream) {
  NUNERF_REQUIRE(p && p->M > 0 && p->metallic && p->rough && p->albedo && p->trans && p->outer && p->inner &&
                     p->weight && p->refrac && p->nov && p->lut && p->color && p->trans_out && p->metallic_out &&
                     p->occ_prob,
                 "shade_mix_fwd: bad arguments");
  shade_mix_fwd_kernel<<<G1(p->M), 0, ST(stream)>>>(*p);
  NUNERF_CHECK_LAUNCH("shade_mix_fwd_kernel");
  return 0;
}
extern "C" int nunerf_shade_mix_bwd(const nunerf_shade_mix_t* p, void* stream) {
  NUNERF_REQUIRE(p && p->M > 0 && p->d_color && p->dz_metallic && p->dz_albedo && p->dz_trans && p->dz_outer &&
                     p->dz_inner && p->dz_weight && p->dz_refrac && p->d_rough_raw && p->d_nov,
                 "shade_mix_bwd: bad arguments");
  shade_mix_bwd_kernel<<<cdiv(p->M, 128), 128, 0, ST(stream)>>>(*p);
  NUNERF_CHECK_LAUNCH("shade_mix_bwd_kernel");
  return 0;
}

extern "C" int nunerf_rowvec_mask(const float* w, const void* a, int lda, int a_lo, int M, int N, void* out, int ldo,
                                  int o_lo, void* stream) {
  NUNERF_REQUIRE(w && a && out && M > 0 && N > 0, "rowvec_mask: bad arguments");
  NUNERF_REQUIRE(((N | lda | a_lo | ldo | o_lo) & 7) == 0, "rowvec_mask: columns / pitches must be multiples of 8");
  rowvec_mask_kernel<<<G1((long long)M * (N / 8)), 0, ST(stream)>>>(w, (const __nv_bfloat16*)a, lda, a_lo, M, N,
                                                             (__nv_bfloat16*)out, ldo, o_lo);
  NUNERF_CHECK_LAUNCH("rowvec_mask_kernel");
  return 0;
}
extern "C" int nunerf_sdf_skip_split(const float* u4, const void* a3, int lda, int a_lo, int M, void* gs3, int ldg,
                                     int g_lo, float* g_skip, void* stream) {
  NUNERF_REQUIRE(u4 && a3 && gs3 && g_skip && M > 0, "sdf_skip_split: bad arguments");
  NUNERF_REQUIRE(((lda | a_lo | ldg | g_lo) & 7) == 0, "sdf_skip_split: pitches must be multiples of 8");
  sdf_skip_split_kernel<<<G1((long long)M * 32), 0, ST(stream)>>>(u4, (const __nv_bfloat16*)a3, lda, a_lo, M,
                                                                  (__nv_bfloat16*)gs3, ldg, g_lo, g_skip);
  NUNERF_CHECK_LAUNCH("sdf_skip_split_kernel");
  return 0;
}
extern "C" int nunerf_sdf_bwd2_ew(const void* gts, int ldt, int t_lo, const void* a, int lda, int a_lo, const void* gs,
                                  int ldg, int g_lo, int M, int N, int n_real, void* u_next, int ldu, int u_lo, void* e,
                                  int lde, int e_lo, void* stream) {
  NUNERF_REQUIRE(gts && a && gs && e && M > 0 && N > 0 && N % 8 == 0, "sdf_bwd2_ew: bad arguments");
  NUNERF_REQUIRE(ldt % 8 == 0 && lda % 8 == 0 && ldg % 8 == 0 && lde % 8 == 0 && (!u_next || ldu % 8 == 0),
                 "sdf_bwd2_ew: leading dimensions must be multiples of 8");
  sdf_bwd2_ew_kernel<<<G1((long long)M * (N / 8)), 0, ST(stream)>>>(
      (const __nv_bfloat16*)gts, ldt, t_lo, (const __nv_bfloat16*)a, lda, a_lo, (const __nv_bfloat16*)gs, ldg, g_lo, M, N,
      n_real, (__nv_bfloat16*)u_next, ldu, u_lo, (__nv_bfloat16*)e, lde, e_lo);
  NUNERF_CHECK_LAUNCH("sdf_bwd2_ew_kernel");
  return 0;
}
extern "C" int nunerf_f32_to_planes(const float* a, int lda, const float* b, int ldb, int M, int C, int width, void* dst,
                                    int ld, int lo, int col, void* stream) {
  NUNERF_REQUIRE(a && dst && M > 0 && C > 0 && width >= C, "f32_to_planes: bad arguments");
  if (((width | col | ld | lo) & 7) == 0 && (((uintptr_t)dst) & 15) == 0) {
    // 8 columns per thread, one 16-byte store per plane (the usual case: a few real columns + zero padding to 64)
    f32_to_planes8_kernel<<<G1((long long)M * (width >> 3)), 0, ST(stream)>>>(a, lda, b, ldb, M, C, width >> 3,
                                                                             (__nv_bfloat16*)dst, ld, lo, col);
    NUNERF_CHECK_LAUNCH("f32_to_planes8_kernel");
    return 0;
  }
  f32_to_planes_kernel<<<G1((long long)M * width), 0, ST(stream)>>>(a, lda, b, ldb, M, C, width, (__nv_bfloat16*)dst,
                                                                   ld, lo, col);
  NUNERF_CHECK_LAUNCH("f32_to_planes_kernel");
  return 0;
}
extern "C" int nunerf_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                           float eps, int step, void* stream) {
  NUNERF_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adam: bad arguments");
  float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  adam_kernel<<<G1(n), 0, ST(stream)>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, bc2);
  NUNERF_CHECK_LAUNCH("adam_kernel");
  return 0;
}
extern "C" int nunerf_grid_points(int res, long long start, int count, const float* lin, float* pts, void* stream) {
  NUNERF_REQUIRE(res > 1 && count > 0 && lin && pts, "grid_points: bad arguments");
  grid_points_kernel<<<G1(count), 0, ST(stream)>>>(res, start, count, lin, pts);
  NUNERF_CHECK_LAUNCH("grid_points_kernel");
  return 0;
}
extern "C" int nunerf_grid_mask(const float* pts, const float* sdf, int ld_sdf, int count, float outside_val, float* u,
                                void* stream) {
  NUNERF_REQUIRE(pts && sdf && u && count > 0, "grid_mask: bad arguments");
  grid_mask_kernel<<<G1(count), 0, ST(stream)>>>(pts, sdf, ld_sdf, count, outside_val, u);
  NUNERF_CHECK_LAUNCH("grid_mask_kernel");
  return 0;
}
