// pointwise.cuh -- per-point math of the fields (forward and hand-derived backward), written as
// __host__ __device__ functions so the same code is exercised by the CPU derivative checks
// (tests/hostsim, test infrastructure only) and by the CUDA kernels in field.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PW_HD __host__ __device__ __forceinline__
#define PW_UNROLL _Pragma("unroll")
#else
#define PW_HD inline
#define PW_UNROLL
#endif

namespace nunerf {
namespace pw {

// ----------------------------------------------------------------------------- IDE tables
// (m, l) pairs with l = 2^i, m = 0..l (utils/ref_utils.py:39-50) and the z-polynomial coefficients
// mat[k][i] = sph_harm_coeff(l, m, k) (utils/ref_utils.py:71-81).
constexpr int IDE_TERMS = 36;
constexpr int IDE_DEG = 17;  // z^0 .. z^16
struct IdeTable {
  float mat[IDE_DEG][IDE_TERMS];
  int m[IDE_TERMS];
  int l[IDE_TERMS];
  float sigma[IDE_TERMS];  // 0.5 * l * (l + 1)
};

inline void build_ide_table(IdeTable* t) {
  auto fact = [](int n) { double r = 1; for (int i = 2; i <= n; ++i) r *= i; return r; };
  int i = 0;
  for (int e = 0; e < 5; ++e) {
    int l = 1 << e;
    for (int m = 0; m <= l; ++m, ++i) {
      t->m[i] = m; t->l[i] = l; t->sigma[i] = 0.5f * l * (l + 1);
      for (int k = 0; k < IDE_DEG; ++k) t->mat[k][i] = 0.f;
      for (int k = 0; k <= l - m; ++k) {
        // generalized binomial coefficient C(0.5*(l+k+m-1), l)
        double a = 0.5 * (l + k + m - 1.0), gb = 1.0;
        for (int j = 0; j < l; ++j) gb *= (a - j);
        gb /= fact(l);
        double leg = ((m & 1) ? -1.0 : 1.0) * pow(2.0, l) * fact(l) / fact(k) / fact(l - k - m) * gb;
        double sh = sqrt((2.0 * l + 1.0) * fact(l - m) / (4.0 * M_PI * fact(l + m))) * leg;
        t->mat[k][i] = (float)sh;
      }
    }
  }
}

// (m, l) of term i as compile-time functions: once the term loop is fully unrolled every table index below is a
// constant, so re/im/out live in registers and the coefficients are constant-bank operands (no local memory).
PW_HD constexpr int ide_l(int i) { return i < 2 ? 1 : i < 5 ? 2 : i < 10 ? 4 : i < 19 ? 8 : 16; }
PW_HD constexpr int ide_m(int i) { return i < 2 ? i : i < 5 ? i - 2 : i < 10 ? i - 5 : i < 19 ? i - 10 : i - 19; }

// out[0..35] = Re, out[36..71] = Im of (x+iy)^m P_i(z) exp(-sigma_i kinv)
PW_HD void ide_fwd(const IdeTable& tb, float x, float y, float z, float kinv, float* out) {
  float re[17], im[17];
  re[0] = 1.f; im[0] = 0.f;
PW_UNROLL
  for (int m = 1; m <= 16; ++m) { re[m] = re[m - 1] * x - im[m - 1] * y; im[m] = re[m - 1] * y + im[m - 1] * x; }
  // attenuation per degree l (5 distinct values): exp(-0.5 l (l+1) kinv)
  float att[5];
PW_UNROLL
  for (int e = 0; e < 5; ++e) { const int l = 1 << e; att[e] = expf(-0.5f * (float)(l * (l + 1)) * kinv); }
PW_UNROLL
  for (int i = 0; i < IDE_TERMS; ++i) {
    const int l = ide_l(i), m = ide_m(i), deg = l - m;
    const int e = l == 1 ? 0 : l == 2 ? 1 : l == 4 ? 2 : l == 8 ? 3 : 4;
    float p = tb.mat[deg][i];
PW_UNROLL
    for (int k = deg - 1; k >= 0; --k) p = p * z + tb.mat[k][i];
    out[i] = re[m] * p * att[e];
    out[IDE_TERMS + i] = im[m] * p * att[e];
  }
}

PW_HD void ide_bwd(const IdeTable& tb, float x, float y, float z, float kinv, const float* dout, float* dx, float* dy,
                   float* dz, float* dkinv) {
  float re[17], im[17];
  re[0] = 1.f; im[0] = 0.f;
PW_UNROLL
  for (int m = 1; m <= 16; ++m) { re[m] = re[m - 1] * x - im[m - 1] * y; im[m] = re[m - 1] * y + im[m - 1] * x; }
  float att[5];
PW_UNROLL
  for (int e = 0; e < 5; ++e) { const int l = 1 << e; att[e] = expf(-0.5f * (float)(l * (l + 1)) * kinv); }
  float gx = 0.f, gy = 0.f, gz = 0.f, gk = 0.f;
PW_UNROLL
  for (int i = 0; i < IDE_TERMS; ++i) {
    const int l = ide_l(i), m = ide_m(i), deg = l - m;
    const int e = l == 1 ? 0 : l == 2 ? 1 : l == 4 ? 2 : l == 8 ? 3 : 4;
    float p = tb.mat[deg][i], dp = 0.f;
PW_UNROLL
    for (int k = deg - 1; k >= 0; --k) { dp = dp * z + p; p = p * z + tb.mat[k][i]; }
    const float a = att[e];
    float dr = dout[i], di = dout[IDE_TERMS + i];
    float s = dr * re[m] + di * im[m];
    gz += s * a * dp;
    gk += s * p * a * (-0.5f * (float)(l * (l + 1)));
    if (m > 0) {
      float dre = dr * p * a, dim = di * p * a, fm = (float)m;
      gx += fm * (dre * re[m - 1] + dim * im[m - 1]);
      gy += fm * (-dre * im[m - 1] + dim * re[m - 1]);
    }
  }
  *dx = gx; *dy = gy; *dz = gz; *dkinv = gk;
}

// ----------------------------------------------------------------------------- small helpers
PW_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
PW_HD float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

PW_HD float srgb_fwd(float x) {  // utils/raw_utils.py:5-12
  const float eps = 1.1920928955078125e-07f;
  return x <= 0.0031308f ? (323.0f / 25.0f) * x : (211.0f * powf(fmaxf(x, eps), 5.0f / 12.0f) - 11.0f) / 200.0f;
}
PW_HD float srgb_bwd(float x) {
  const float eps = 1.1920928955078125e-07f;
  if (x <= 0.0031308f) return 323.0f / 25.0f;
  if (x < eps) return 0.f;
  return (211.0f / 200.0f) * (5.0f / 12.0f) * powf(x, -7.0f / 12.0f);
}

// positional encoding of one scalar coordinate layout: out index for (freq k, fn, dim c) with d dims:
//   [x (d)] [sin(2^0 x) (d)] [cos(2^0 x) (d)] [sin(2^1 x) (d)] ...
PW_HD int pe_index(int d, int k, int is_cos, int c) { return d + (2 * k + is_cos) * d + c; }

// J_pe(x)^T g for one coordinate c of a d-dimensional input with F frequencies (layout of pe_index): g points at the PE row
PW_HD float pe_bwd_coord(const float* g, int d, int F, int c, float xc) {
  float acc = g[c];
  for (int k = 0; k < F; ++k) {
    float f = (float)(1 << k), sn = sinf(xc * f), co = cosf(xc * f);
    acc += f * (co * g[pe_index(d, k, 0, c)] - sn * g[pe_index(d, k, 1, c)]);
  }
  return acc;
}
// d/dx_c of [J_pe(x)^T a]_c = sum_k f^2 (-sin(f x_c) a_sin - cos(f x_c) a_cos)   (second derivative of the encoding)
PW_HD float pe_hess_coord(const float* a, int d, int F, int c, float xc) {
  float acc = 0.f;
  for (int k = 0; k < F; ++k) {
    float f = (float)(1 << k), sn = sinf(xc * f), co = cosf(xc * f);
    acc -= f * f * (sn * a[pe_index(d, k, 0, c)] + co * a[pe_index(d, k, 1, c)]);
  }
  return acc;
}

// ----------------------------------------------------------------------------- sdf -> alpha (ZT:657-685, :769)
struct SdfAlphaOut { float alpha, gerr; };
PW_HD SdfAlphaOut sdf_alpha_fwd(float sdf, const float* g, float dist, const float* dir, float inv_s, float anneal) {
  float tc = dir[0] * g[0] + dir[1] * g[1] + dir[2] * g[2];
  float r1 = fmaxf(-tc * 0.5f + 0.5f, 0.f), r2 = fmaxf(-tc, 0.f);
  float ic = -(r1 * (1.0f - anneal) + r2 * anneal);
  float en = sdf + ic * dist * 0.5f, ep = sdf - ic * dist * 0.5f;
  float pc = sigmoidf_(ep * inv_s), nc = sigmoidf_(en * inv_s);
  float a = (pc - nc + 1e-5f) / (pc + 1e-5f);
  float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  SdfAlphaOut o;
  o.alpha = clamp01(a);
  o.gerr = (gn - 1.0f) * (gn - 1.0f);
  return o;
}
PW_HD void sdf_alpha_bwd(float sdf, const float* g, float dist, const float* dir, float inv_s, float anneal,
                         float d_alpha, float d_gerr, float* d_sdf, float* d_g, float* d_inv_s, float* d_dist = nullptr,
                         float* d_dir = nullptr) {
  float tc = dir[0] * g[0] + dir[1] * g[1] + dir[2] * g[2];
  float q1 = -tc * 0.5f + 0.5f, q2 = -tc;
  float r1 = fmaxf(q1, 0.f), r2 = fmaxf(q2, 0.f);
  float ic = -(r1 * (1.0f - anneal) + r2 * anneal);
  float en = sdf + ic * dist * 0.5f, ep = sdf - ic * dist * 0.5f;
  float pc = sigmoidf_(ep * inv_s), nc = sigmoidf_(en * inv_s);
  float num = pc - nc + 1e-5f, den = pc + 1e-5f;
  float a = num / den;
  float da = (a >= 0.f && a <= 1.f) ? d_alpha : 0.f;
  float dnum = da / den, dden = -da * num / (den * den);
  float dpc = dnum + dden, dnc = -dnum;
  float dpa = dpc * pc * (1.0f - pc), dna = dnc * nc * (1.0f - nc);
  float dep = dpa * inv_s, den_ = dna * inv_s;
  *d_inv_s = dpa * ep + dna * en;
  *d_sdf = dep + den_;
  float dic = (den_ - dep) * dist * 0.5f;
  float dr1 = -dic * (1.0f - anneal), dr2 = -dic * anneal;
  float dtc = (q1 > 0.f ? dr1 * -0.5f : 0.f) + (q2 > 0.f ? -dr2 : 0.f);
  float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  float ge = gn > 0.f ? d_gerr * 2.0f * (gn - 1.0f) / gn : 0.f;
  for (int c = 0; c < 3; ++c) d_g[c] = dtc * dir[c] + ge * g[c];
  // position-gradient outputs (stage 2: the sample positions depend on the refracted path): distance and ray direction
  if (d_dist) *d_dist = (den_ - dep) * ic * 0.5f;
  if (d_dir)
    for (int c = 0; c < 3; ++c) d_dir[c] = dtc * g[c];
}

// ----------------------------------------------------------------------------- NeRF++ output (ZT:515-516, :691-692)
PW_HD float softplus1(float x) { return x > 20.f ? x : log1pf(expf(x)); }
PW_HD void nerf_out_fwd(float sigma, const float* rgb, float dist, float* alpha, float* color) {
  *alpha = 1.0f - expf(-softplus1(sigma) * dist);
  for (int c = 0; c < 3; ++c) color[c] = srgb_fwd(expf(fminf(rgb[c], 5.0f)));
}
PW_HD void nerf_out_bwd(float sigma, const float* rgb, float dist, float d_alpha, const float* d_color, float* d_sigma,
                        float* d_rgb, float* d_dist = nullptr) {
  float sp = softplus1(sigma);
  float dsp = d_alpha * expf(-sp * dist) * dist;
  if (d_dist) *d_dist = d_alpha * expf(-sp * dist) * sp;
  *d_sigma = dsp * (sigma > 20.f ? 1.0f : sigmoidf_(sigma));
  for (int c = 0; c < 3; ++c) {
    float e = expf(fminf(rgb[c], 5.0f));
    d_rgb[c] = rgb[c] <= 5.0f ? d_color[c] * srgb_bwd(e) * e : 0.f;
  }
}

// ----------------------------------------------------------------------------- shading directions (field.py:686-689)
struct ShadeDirs { float n[3], v[3], r[3], nov, gn; };
PW_HD ShadeDirs shade_dirs(const float* g, const float* raydir) {
  ShadeDirs s;
  s.gn = fmaxf(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-12f);
  float vn = fmaxf(sqrtf(raydir[0] * raydir[0] + raydir[1] * raydir[1] + raydir[2] * raydir[2]), 1e-12f);
  for (int c = 0; c < 3; ++c) { s.n[c] = g[c] / s.gn; s.v[c] = -raydir[c] / vn; }
  s.nov = s.n[0] * s.v[0] + s.n[1] * s.v[1] + s.n[2] * s.v[2];
  for (int c = 0; c < 3; ++c) s.r[c] = s.nov * s.n[c] * 2.0f - s.v[c];
  return s;
}
// given d_r, d_n (direct), d_nov -> d_g; optionally also d_raydir (v = -raydir / |raydir|; d_v_direct = gradient that
// reaches v directly, e.g. through PE(v)), |raydir| = vn
PW_HD void shade_dirs_bwd(const ShadeDirs& s, const float* d_r, const float* d_n_direct, float d_nov, float* d_g,
                          const float* d_v_direct = nullptr, float vn = 1.0f, float* d_raydir = nullptr) {
  float ndr = s.n[0] * d_r[0] + s.n[1] * d_r[1] + s.n[2] * d_r[2];
  if (d_raydir) {
    // r = 2 (n.v) n - v  ->  d_v = 2 n (n.d_r) - d_r ; nov = n.v -> d_v += d_nov n
    float dv[3];
    for (int c = 0; c < 3; ++c) dv[c] = 2.0f * s.n[c] * ndr - d_r[c] + d_nov * s.n[c] + (d_v_direct ? d_v_direct[c] : 0.f);
    float vdv = s.v[0] * dv[0] + s.v[1] * dv[1] + s.v[2] * dv[2];
    for (int c = 0; c < 3; ++c) d_raydir[c] = -(dv[c] - s.v[c] * vdv) / vn;
  }
  float dn[3];
  for (int c = 0; c < 3; ++c) dn[c] = d_n_direct[c] + 2.0f * s.v[c] * ndr + 2.0f * s.nov * d_r[c] + d_nov * s.v[c];
  float ndn = s.n[0] * dn[0] + s.n[1] * dn[1] + s.n[2] * dn[2];
  for (int c = 0; c < 3; ++c) d_g[c] = (dn[c] - s.n[c] * ndn) / s.gn;
}

// ----------------------------------------------------------------------------- sphere direction (field.py:447-465, :641-644)
// The `sphere_direction` shading variant looks the direct light up at the point where the ray (p, u) leaves the unit sphere:
// q = normalize(ps + u t), ps = p pulled inside radius 0.999 (offset_points_to_sphere), t = -ps.u + sqrt((ps.u)^2 - ps.ps + 1
// + 1e-6) (get_sphere_intersection).
PW_HD void sphere_dir_fwd(const float* p, const float* u, float* q) {
  float pn = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  float ps[3] = {p[0], p[1], p[2]};
  if (pn > 0.999f)
    for (int c = 0; c < 3; ++c) ps[c] = p[c] / pn * 0.999f;
  float dtx = ps[0] * u[0] + ps[1] * u[1] + ps[2] * u[2];
  float xtx = ps[0] * ps[0] + ps[1] * ps[1] + ps[2] * ps[2];
  float t = -dtx + sqrtf(dtx * dtx - xtx + 1.0f + 1e-6f);
  float y[3] = {ps[0] + u[0] * t, ps[1] + u[1] * t, ps[2] + u[2] * t};
  float yn = fmaxf(sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]), 1e-12f);
  for (int c = 0; c < 3; ++c) q[c] = y[c] / yn;
}
// d_q -> d_u (added) and d_p (added when given)
PW_HD void sphere_dir_bwd(const float* p, const float* u, const float* d_q, float* d_u, float* d_p) {
  float pn = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  bool scaled = pn > 0.999f;
  float ps[3] = {p[0], p[1], p[2]};
  if (scaled)
    for (int c = 0; c < 3; ++c) ps[c] = p[c] / pn * 0.999f;
  float dtx = ps[0] * u[0] + ps[1] * u[1] + ps[2] * u[2];
  float xtx = ps[0] * ps[0] + ps[1] * ps[1] + ps[2] * ps[2];
  float sq = sqrtf(dtx * dtx - xtx + 1.0f + 1e-6f);
  float t = -dtx + sq;
  float y[3] = {ps[0] + u[0] * t, ps[1] + u[1] * t, ps[2] + u[2] * t};
  float yn = fmaxf(sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]), 1e-12f);
  float q[3] = {y[0] / yn, y[1] / yn, y[2] / yn};
  float qdq = q[0] * d_q[0] + q[1] * d_q[1] + q[2] * d_q[2];
  float dy[3];
  for (int c = 0; c < 3; ++c) dy[c] = (d_q[c] - q[c] * qdq) / yn;
  float dt = dy[0] * u[0] + dy[1] * u[1] + dy[2] * u[2];
  float ddtx = dt * (-1.0f + dtx / sq), dxtx = -dt * 0.5f / sq;
  float dps[3];
  for (int c = 0; c < 3; ++c) {
    d_u[c] += dy[c] * t + ddtx * ps[c];
    dps[c] = dy[c] + ddtx * u[c] + 2.0f * dxtx * ps[c];
  }
  if (d_p) {
    if (scaled) {
      float ph[3] = {p[0] / pn, p[1] / pn, p[2] / pn};
      float pd = ph[0] * dps[0] + ph[1] * dps[1] + ph[2] * dps[2];
      for (int c = 0; c < 3; ++c) d_p[c] += 0.999f * (dps[c] - ph[c] * pd) / pn;
    } else {
      for (int c = 0; c < 3; ++c) d_p[c] += dps[c];
    }
  }
}

// ----------------------------------------------------------------------------- non-zero-thickness bounce
// network/renderer.py:1690-2009 per hit ray (see nu_nerf_b200/shell.py for the derivation and the reference line map): the
// outer mesh is one face of a glass shell of thickness tau = 0.01 thick_sig, locally two concentric spheres of radius
// r = 1 / sqrt(|K|) and r -+ tau.  Entering: refract (ratio ior), cross the shell along the chord, refract again (ratio ioo).
// Leaving: first pull the hit back onto the inner face.  s = -1 / +1 folds the two curvature-sign cases (exact sign factor).
struct ShellIn { float x[3], n[3], d[3], gk, ior_sig, th_sig; };
struct ShellOut { int ok, tir; float x_mod[3], start[3], dir[3], ratio; };
#define PW_SHELL_IOR_INNER (1.0f / 1.0001f)

PW_HD float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// u = v / (|v| + 1e-4)
PW_HD float unit4(const float* v, float* u) {
  float nv = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  float den = nv + 0.0001f;
  for (int c = 0; c < 3; ++c) u[c] = v[c] / den;
  return nv;
}
// adjoint of unit4: d_v += (d_u - u (u.d_u) |v| / (|v| + 1e-4)) / (|v| + 1e-4)   [d|v|/dv = v/|v|]
PW_HD void unit4_bwd(const float* v, float nv, const float* u, const float* d_u, float* d_v) {
  float den = nv + 0.0001f;
  float udu = dot3(u, d_u);
  for (int c = 0; c < 3; ++c) d_v[c] += (d_u[c] - (nv > 0.f ? v[c] / nv : 0.f) * udu) / den;
}

PW_HD void shell_bounce_fwd(const ShellIn& in, int inside, ShellOut* o) {
  const float* x = in.x; const float* n = in.n; const float* d = in.d;
  float cos_i = -dot3(n, d);
  float sin2_i = 1.0f - cos_i * cos_i;
  float a = 1.0f / (in.ior_sig * 1.0f + 0.6f);
  float b = PW_SHELL_IOR_INNER / a;
  float th = in.th_sig * 0.01f;
  float ior = inside ? 1.0f / b : a, ioo = inside ? 1.0f / a : b;
  o->ok = !(ior * ior * sin2_i > 0.999f);
  o->tir = o->ok;
  o->ratio = ior;
  for (int c = 0; c < 3; ++c) { o->x_mod[c] = x[c]; o->start[c] = 0.f; o->dir[c] = 0.f; }
  if (!o->ok) return;
  float sin2_t = sin2_i * ior * ior;
  float r = 1.0f / sqrtf(fmaxf(fabsf(in.gk), 0.000001f));
  if (r != r) r = 0.1f;
  float cos_t = sqrtf(fmaxf(1.0f - sin2_t, 0.0001f));
  float s, x_in[3], n_in[3], d_in[3], v1[3];
  if (!inside) {
    s = in.gk >= 0.f ? -1.0f : 1.0f;
    float k1 = ior * cos_i - sqrtf(fmaxf(1.0f - sin2_t, 0.0001f));
    for (int c = 0; c < 3; ++c) { v1[c] = ior * d[c] + k1 * n[c]; x_in[c] = x[c]; n_in[c] = n[c]; }
    unit4(v1, d_in);
  } else {
    s = in.gk <= 0.f ? -1.0f : 1.0f;
    float c_r = r * cos_i;
    float delta = sqrtf(fmaxf(c_r * c_r + s * (2.0f * r * th) + th * th, 0.0001f));
    float L1 = fabsf(c_r - delta);
    float nm[3];
    for (int c = 0; c < 3; ++c) {
      float center = x[c] + s * (n[c] * r);
      x_in[c] = x[c] - L1 * d[c];
      nm[c] = -s * (x_in[c] - center);
      o->x_mod[c] = x_in[c];
    }
    unit4(nm, n_in);
    float cos_m = -dot3(n_in, d);
    float e2 = (1.0f - cos_m * cos_m) * ior * ior;
    if (e2 > 0.999f) o->tir = 0;
    float k1 = ior * cos_m - sqrtf(fmaxf(1.0f - fminf(e2, 0.999f), 0.0001f));
    for (int c = 0; c < 3; ++c) v1[c] = ior * d[c] + k1 * n_in[c];
    unit4(v1, d_in);
  }
  float ctr = r * cos_t;
  float delta2 = sqrtf(fmaxf(ctr * ctr + s * (2.0f * r * th) + th * th, 0.0001f));
  float L2 = fabsf(ctr - delta2);
  float na[3], n_af[3], v3[3];
  for (int c = 0; c < 3; ++c) {
    float center = x_in[c] + s * (n_in[c] * r);
    o->start[c] = x_in[c] + d_in[c] * (L2 + 0.001f);
    na[c] = -s * (o->start[c] - center);
  }
  unit4(na, n_af);
  float cos_2 = -dot3(n_af, d_in);
  float e3 = (1.0f - cos_2 * cos_2) * ioo * ioo;
  if (e3 > 0.999f) o->tir = 0;
  float k3 = ioo * cos_2 - sqrtf(fmaxf(1.0f - fminf(e3, 0.999f), 0.0001f));
  for (int c = 0; c < 3; ++c) v3[c] = ioo * d_in[c] + k3 * n_af[c];
  unit4(v3, o->dir);
}

// Adjoint of shell_bounce_fwd for a ray with ok = 1: gradients of (start, dir, ratio, x_mod) -> d_in (x, n, d, gk, ior_sig,
// th_sig); every clamp / abs / min passes the gradient on its active side only, like autograd.
PW_HD void shell_bounce_bwd(const ShellIn& in, int inside, const float* g_start, const float* g_dir, float g_ratio,
                            const float* g_xmod, ShellIn* di) {
  const float* x = in.x; const float* n = in.n; const float* d = in.d;
  for (int c = 0; c < 3; ++c) { di->x[c] = 0.f; di->n[c] = 0.f; di->d[c] = 0.f; }
  di->gk = 0.f; di->ior_sig = 0.f; di->th_sig = 0.f;
  // ---- forward again, keeping the intermediates
  float cos_i = -dot3(n, d);
  float sin2_i = 1.0f - cos_i * cos_i;
  float a = 1.0f / (in.ior_sig * 1.0f + 0.6f);
  float b = PW_SHELL_IOR_INNER / a;
  float th = in.th_sig * 0.01f;
  float ior = inside ? 1.0f / b : a, ioo = inside ? 1.0f / a : b;
  float sin2_t = sin2_i * ior * ior;
  float ak = fabsf(in.gk);
  float r = 1.0f / sqrtf(fmaxf(ak, 0.000001f));
  float ct_arg = 1.0f - sin2_t;
  float cos_t = sqrtf(fmaxf(ct_arg, 0.0001f));
  float s, x_in[3], n_in[3], d_in[3], v1[3], nv1, nm[3] = {0.f, 0.f, 0.f}, nnm = 0.f;
  float c_r = 0.f, q1 = 0.f, delta1 = 0.f, L1 = 0.f, cos_m = 0.f, e2 = 0.f, k1;
  if (!inside) {
    s = in.gk >= 0.f ? -1.0f : 1.0f;
    k1 = ior * cos_i - cos_t;
    for (int c = 0; c < 3; ++c) { v1[c] = ior * d[c] + k1 * n[c]; x_in[c] = x[c]; n_in[c] = n[c]; }
  } else {
    s = in.gk <= 0.f ? -1.0f : 1.0f;
    c_r = r * cos_i;
    q1 = c_r * c_r + s * (2.0f * r * th) + th * th;
    delta1 = sqrtf(fmaxf(q1, 0.0001f));
    L1 = fabsf(c_r - delta1);
    for (int c = 0; c < 3; ++c) {
      float center = x[c] + s * (n[c] * r);
      x_in[c] = x[c] - L1 * d[c];
      nm[c] = -s * (x_in[c] - center);
    }
    nnm = unit4(nm, n_in);
    cos_m = -dot3(n_in, d);
    e2 = (1.0f - cos_m * cos_m) * ior * ior;
    k1 = ior * cos_m - sqrtf(fmaxf(1.0f - fminf(e2, 0.999f), 0.0001f));
    for (int c = 0; c < 3; ++c) v1[c] = ior * d[c] + k1 * n_in[c];
  }
  nv1 = unit4(v1, d_in);
  float ctr = r * cos_t;
  float q2 = ctr * ctr + s * (2.0f * r * th) + th * th;
  float delta2 = sqrtf(fmaxf(q2, 0.0001f));
  float L2 = fabsf(ctr - delta2);
  float start[3], na[3], n_af[3], v3[3], dirn[3];
  for (int c = 0; c < 3; ++c) {
    float center = x_in[c] + s * (n_in[c] * r);
    start[c] = x_in[c] + d_in[c] * (L2 + 0.001f);
    na[c] = -s * (start[c] - center);
  }
  float nna = unit4(na, n_af);
  float cos_2 = -dot3(n_af, d_in);
  float e3 = (1.0f - cos_2 * cos_2) * ioo * ioo;
  float a3 = 1.0f - fminf(e3, 0.999f);
  float sq3 = sqrtf(fmaxf(a3, 0.0001f));
  float k3 = ioo * cos_2 - sq3;
  for (int c = 0; c < 3; ++c) v3[c] = ioo * d_in[c] + k3 * n_af[c];
  float nv3 = unit4(v3, dirn);
  // ---- reverse
  float d_ior = g_ratio, d_ioo = 0.f, d_r = 0.f, d_th = 0.f, d_cos_t = 0.f, d_cos_i = 0.f;
  float d_xin[3] = {0.f, 0.f, 0.f}, d_nin[3] = {0.f, 0.f, 0.f}, d_din[3] = {0.f, 0.f, 0.f};
  float d_v3[3] = {0.f, 0.f, 0.f}, d_naf[3] = {0.f, 0.f, 0.f}, d_start[3];
  unit4_bwd(v3, nv3, dirn, g_dir, d_v3);
  // v3 = ioo d_in + k3 n_af
  float d_k3 = dot3(d_v3, n_af);
  d_ioo += dot3(d_v3, d_in);
  for (int c = 0; c < 3; ++c) { d_din[c] += ioo * d_v3[c]; d_naf[c] += k3 * d_v3[c]; }
  // k3 = ioo cos_2 - sqrt(max(1 - min(e3, .999), 1e-4))
  d_ioo += d_k3 * cos_2;
  float d_cos_2 = d_k3 * ioo;
  float d_e3 = (a3 > 0.0001f && e3 < 0.999f) ? d_k3 * 0.5f / sq3 : 0.f;       // d(-sqrt(1 - e3)) / d e3 = +1 / (2 sqrt)
  // e3 = (1 - cos_2^2) ioo^2
  d_cos_2 += d_e3 * (-2.0f * cos_2) * ioo * ioo;
  d_ioo += d_e3 * (1.0f - cos_2 * cos_2) * 2.0f * ioo;
  // cos_2 = -(n_af . d_in)
  for (int c = 0; c < 3; ++c) { d_naf[c] -= d_cos_2 * d_in[c]; d_din[c] -= d_cos_2 * n_af[c]; }
  // n_af = unit4(na); na = -s (start - center2); center2 = x_in + s n_in r
  float d_na[3] = {0.f, 0.f, 0.f};
  unit4_bwd(na, nna, n_af, d_naf, d_na);
  for (int c = 0; c < 3; ++c) {
    d_start[c] = g_start[c] - s * d_na[c];
    float d_center = s * d_na[c];
    d_xin[c] += d_center;
    d_nin[c] += d_center * s * r;
    d_r += d_center * s * n_in[c];
  }
  // start = x_in + d_in (L2 + 0.001)
  float d_L2 = 0.f;
  for (int c = 0; c < 3; ++c) { d_xin[c] += d_start[c]; d_din[c] += d_start[c] * (L2 + 0.001f); d_L2 += d_start[c] * d_in[c]; }
  // L2 = |ctr - delta2|; delta2 = sqrt(max(q2, 1e-4)); q2 = ctr^2 + s 2 r th + th^2; ctr = r cos_t
  float sg2 = (ctr - delta2) > 0.f ? 1.0f : ((ctr - delta2) < 0.f ? -1.0f : 0.f);
  float d_ctr = d_L2 * sg2, d_delta2 = -d_L2 * sg2;
  float d_q2 = q2 > 0.0001f ? d_delta2 * 0.5f / delta2 : 0.f;
  d_ctr += d_q2 * 2.0f * ctr;
  d_r += d_q2 * s * 2.0f * th;
  d_th += d_q2 * (s * 2.0f * r + 2.0f * th);
  d_r += d_ctr * cos_t;
  d_cos_t += d_ctr * r;
  // d_in = unit4(v1)
  float d_v1[3] = {0.f, 0.f, 0.f};
  unit4_bwd(v1, nv1, d_in, d_din, d_v1);
  float d_d[3] = {0.f, 0.f, 0.f}, d_n[3] = {0.f, 0.f, 0.f}, d_x[3] = {0.f, 0.f, 0.f};
  if (!inside) {
    // v1 = ior d + k1 n, k1 = ior cos_i - cos_t; x_in = x, n_in = n
    float d_k1 = dot3(d_v1, n);
    d_ior += dot3(d_v1, d);
    for (int c = 0; c < 3; ++c) { d_d[c] += ior * d_v1[c]; d_n[c] += k1 * d_v1[c] + d_nin[c]; d_x[c] += d_xin[c] + g_xmod[c]; }
    d_ior += d_k1 * cos_i;
    d_cos_i += d_k1 * ior;
    d_cos_t -= d_k1;
  } else {
    // v1 = ior d + k1 n_in, k1 = ior cos_m - sqrt(max(1 - min(e2, .999), 1e-4))
    float d_k1 = dot3(d_v1, n_in);
    d_ior += dot3(d_v1, d);
    for (int c = 0; c < 3; ++c) { d_d[c] += ior * d_v1[c]; d_nin[c] += k1 * d_v1[c]; }
    d_ior += d_k1 * cos_m;
    float d_cos_m = d_k1 * ior;
    float a2 = 1.0f - fminf(e2, 0.999f);
    float d_e2 = (a2 > 0.0001f && e2 < 0.999f) ? d_k1 * 0.5f / sqrtf(a2) : 0.f;
    d_cos_m += d_e2 * (-2.0f * cos_m) * ior * ior;
    d_ior += d_e2 * (1.0f - cos_m * cos_m) * 2.0f * ior;
    // cos_m = -(n_in . d)
    for (int c = 0; c < 3; ++c) { d_nin[c] -= d_cos_m * d[c]; d_d[c] -= d_cos_m * n_in[c]; }
    // n_in = unit4(nm); nm = -s (x_in - center1); center1 = x + s n r; x_in = x - L1 d  (x_mod = x_in)
    float d_nm[3] = {0.f, 0.f, 0.f};
    unit4_bwd(nm, nnm, n_in, d_nin, d_nm);
    float d_L1 = 0.f;
    for (int c = 0; c < 3; ++c) {
      float d_xi = d_xin[c] + g_xmod[c] - s * d_nm[c];
      float d_center = s * d_nm[c];
      d_x[c] += d_xi + d_center;
      d_d[c] -= d_xi * L1;
      d_L1 -= d_xi * d[c];
      d_n[c] += d_center * s * r;
      d_r += d_center * s * n[c];
    }
    // L1 = |c_r - delta1|; delta1 = sqrt(max(q1, 1e-4)); q1 = c_r^2 + s 2 r th + th^2; c_r = r cos_i
    float sg1 = (c_r - delta1) > 0.f ? 1.0f : ((c_r - delta1) < 0.f ? -1.0f : 0.f);
    float d_cr = d_L1 * sg1, d_delta1 = -d_L1 * sg1;
    float d_q1 = q1 > 0.0001f ? d_delta1 * 0.5f / delta1 : 0.f;
    d_cr += d_q1 * 2.0f * c_r;
    d_r += d_q1 * s * 2.0f * th;
    d_th += d_q1 * (s * 2.0f * r + 2.0f * th);
    d_r += d_cr * cos_i;
    d_cos_i += d_cr * r;
  }
  // cos_t = sqrt(max(1 - sin2_t, 1e-4)); sin2_t = sin2_i ior^2; sin2_i = 1 - cos_i^2; cos_i = -(n . d)
  float d_sin2_t = ct_arg > 0.0001f ? -d_cos_t * 0.5f / cos_t : 0.f;
  float d_sin2_i = d_sin2_t * ior * ior;
  d_ior += d_sin2_t * sin2_i * 2.0f * ior;
  d_cos_i += d_sin2_i * (-2.0f * cos_i);
  for (int c = 0; c < 3; ++c) { d_n[c] -= d_cos_i * d[c]; d_d[c] -= d_cos_i * n[c]; }
  // r = 1 / sqrt(max(|gk|, 1e-6))
  if (ak > 0.000001f) di->gk = d_r * (-0.5f) * r / ak * (in.gk > 0.f ? 1.0f : -1.0f);
  // ior / ioo -> a, b -> ior_sig;  a = 1 / (sig + 0.6), b = IOR_INNER / a
  float d_a, d_b;
  if (inside) { d_b = -d_ior / (b * b); d_a = -d_ioo / (a * a); }
  else { d_a = d_ior; d_b = d_ioo; }
  d_a += d_b * (-PW_SHELL_IOR_INNER / (a * a));
  di->ior_sig = d_a * (-a * a);
  di->th_sig = d_th * 0.01f;
  for (int c = 0; c < 3; ++c) { di->x[c] = d_x[c]; di->n[c] = d_n[c]; di->d[c] = d_d[c]; }
}

// ----------------------------------------------------------------------------- FG LUT (dr.texture linear/clamp)
PW_HD void fg_lookup(const float* lut, float u, float v, float* fg, float* dfg_du, float* dfg_dv) {
  const int W = 256, H = 256;
  float fx = u * W - 0.5f, fy = v * H - 0.5f;
  bool cx = fx > 0.f && fx < (float)(W - 1), cy = fy > 0.f && fy < (float)(H - 1);
  fx = fminf(fmaxf(fx, 0.f), (float)(W - 1));
  fy = fminf(fmaxf(fy, 0.f), (float)(H - 1));
  int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
  float tx = fx - x0, ty = fy - y0;
  int x1 = x0 + 1 > W - 1 ? W - 1 : x0 + 1, y1 = y0 + 1 > H - 1 ? H - 1 : y0 + 1;
  for (int c = 0; c < 2; ++c) {
    float c00 = lut[(y0 * W + x0) * 2 + c], c01 = lut[(y0 * W + x1) * 2 + c];
    float c10 = lut[(y1 * W + x0) * 2 + c], c11 = lut[(y1 * W + x1) * 2 + c];
    fg[c] = (c00 * (1 - tx) + c01 * tx) * (1 - ty) + (c10 * (1 - tx) + c11 * tx) * ty;
    if (dfg_du) {
      dfg_du[c] = cx ? W * ((c01 - c00) * (1 - ty) + (c11 - c10) * ty) : 0.f;
      dfg_dv[c] = cy ? H * ((c10 * (1 - tx) + c11 * tx) - (c00 * (1 - tx) + c01 * tx)) : 0.f;
    }
  }
}

// ----------------------------------------------------------------------------- shading mix (field.py:691-741)
struct ShadeMixIn {
  float metallic, rough, albedo[3], trans;               // raw heads (pre-sigmoid)
  float diffuse_l[3], direct[3], direct0[3];             // raw heads (pre-exp)
  float indirect[3], indirect0[3], occ, refrac[3];
  float nov;
};
struct ShadeMixOut { float color[3], trans, metallic, occ_prob; };

PW_HD float exp_act(float x, float mx) { return expf(fminf(x, mx)); }

// exp_max_r: clamp of the refraction-light head (AppShadingNetwork: = exp_max, field.py:602; AppShadingNetwork_SpecInner: -0.2,
// field.py:1373)
PW_HD ShadeMixOut shade_mix_fwd(const ShadeMixIn& in, const float* lut, float exp_max, float exp_max_r) {
  ShadeMixOut o;
  float met = sigmoidf_(in.metallic), rough = sigmoidf_(in.rough), T = sigmoidf_(in.trans);
  float occ = in.occ * 0.5f + 0.5f, oc = clamp01(occ);
  float t = clamp01(1.0f - in.nov);
  float rw = clamp01(0.04f + 0.96f * t * t * t * t * t);
  float fg[2];
  fg_lookup(lut, clamp01(in.nov), clamp01(rough), fg, nullptr, nullptr);
  for (int c = 0; c < 3; ++c) {
    float alb = sigmoidf_(in.albedo[c]);
    float dl = exp_act(in.diffuse_l[c], exp_max), dr = exp_act(in.direct[c], exp_max),
          d0 = exp_act(in.direct0[c], exp_max), il = exp_act(in.indirect[c], exp_max),
          i0 = exp_act(in.indirect0[c], exp_max), rf = exp_act(in.refrac[c], exp_max_r);
    float diffuse = (1.0f - met) * alb * dl;
    float sa = 0.04f * (1.0f - met) + met * alb;
    float light = il * oc + dr * (1.0f - oc), light0 = i0 * oc + d0 * (1.0f - oc);
    float spec = (sa * fg[0] + fg[1]) * light;
    float lin = (diffuse + spec) * (1.0f - T) + (rw * light0 + (1.0f - rw) * rf) * T;
    o.color[c] = srgb_fwd(lin);
  }
  o.trans = T; o.metallic = met; o.occ_prob = occ;
  return o;
}

// d_in receives gradients w.r.t. every raw head and nov
PW_HD void shade_mix_bwd(const ShadeMixIn& in, const float* lut, float exp_max, float exp_max_r, const float* d_color,
                         float d_trans_out, float d_met_out, ShadeMixIn* d_in) {
  float met = sigmoidf_(in.metallic), rough = sigmoidf_(in.rough), T = sigmoidf_(in.trans);
  float occ = in.occ * 0.5f + 0.5f, oc = clamp01(occ);
  float t1 = 1.0f - in.nov, t = clamp01(t1);
  float sch = 0.04f + 0.96f * t * t * t * t * t, rw = clamp01(sch);
  float fg[2], dfu[2], dfv[2];
  float un = clamp01(in.nov), vr = clamp01(rough);
  fg_lookup(lut, un, vr, fg, dfu, dfv);
  float d_met = 0.f, d_T = 0.f, d_oc = 0.f, d_rw = 0.f, d_fg0 = 0.f, d_fg1 = 0.f;
  for (int c = 0; c < 3; ++c) {
    float alb = sigmoidf_(in.albedo[c]);
    float dl = exp_act(in.diffuse_l[c], exp_max), dr = exp_act(in.direct[c], exp_max),
          d0 = exp_act(in.direct0[c], exp_max), il = exp_act(in.indirect[c], exp_max),
          i0 = exp_act(in.indirect0[c], exp_max), rf = exp_act(in.refrac[c], exp_max_r);
    float diffuse = (1.0f - met) * alb * dl;
    float sa = 0.04f * (1.0f - met) + met * alb;
    float light = il * oc + dr * (1.0f - oc), light0 = i0 * oc + d0 * (1.0f - oc);
    float sref = sa * fg[0] + fg[1];
    float spec = sref * light;
    float mixB = rw * light0 + (1.0f - rw) * rf;
    float lin = (diffuse + spec) * (1.0f - T) + mixB * T;
    float dlin = d_color[c] * srgb_bwd(lin);
    d_T += dlin * (mixB - (diffuse + spec));
    float d_diffuse = dlin * (1.0f - T), d_spec = d_diffuse, d_mixB = dlin * T;
    // diffuse = (1-met) alb dl
    d_met += d_diffuse * (-alb * dl);
    float d_alb = d_diffuse * (1.0f - met) * dl;
    float d_dl = d_diffuse * (1.0f - met) * alb;
    // spec = sref * light
    float d_sref = d_spec * light, d_light = d_spec * sref;
    float d_sa = d_sref * fg[0];
    d_fg0 += d_sref * sa; d_fg1 += d_sref;
    d_met += d_sa * (-0.04f + alb);
    d_alb += d_sa * met;
    // mixB
    d_rw += d_mixB * (light0 - rf);
    float d_light0 = d_mixB * rw, d_rf = d_mixB * (1.0f - rw);
    // lights
    float d_il = d_light * oc, d_dr = d_light * (1.0f - oc);
    float d_i0 = d_light0 * oc, d_d0 = d_light0 * (1.0f - oc);
    d_oc += d_light * (il - dr) + d_light0 * (i0 - d0);
    // raw heads
    d_in->albedo[c] = d_alb * alb * (1.0f - alb);
    d_in->diffuse_l[c] = in.diffuse_l[c] <= exp_max ? d_dl * dl : 0.f;
    d_in->direct[c] = in.direct[c] <= exp_max ? d_dr * dr : 0.f;
    d_in->direct0[c] = in.direct0[c] <= exp_max ? d_d0 * d0 : 0.f;
    d_in->indirect[c] = in.indirect[c] <= exp_max ? d_il * il : 0.f;
    d_in->indirect0[c] = in.indirect0[c] <= exp_max ? d_i0 * i0 : 0.f;
    d_in->refrac[c] = in.refrac[c] <= exp_max_r ? d_rf * rf : 0.f;
  }
  d_met += d_met_out;
  d_T += d_trans_out;
  d_in->metallic = d_met * met * (1.0f - met);
  d_in->trans = d_T * T * (1.0f - T);
  d_in->occ = (occ >= 0.f && occ <= 1.f) ? d_oc * 0.5f : 0.f;
  // FG lookup: u = clamp(nov), v = clamp(rough)
  float d_u = d_fg0 * dfu[0] + d_fg1 * dfu[1], d_v = d_fg0 * dfv[0] + d_fg1 * dfv[1];
  float d_rough = (rough >= 0.f && rough <= 1.f) ? d_v : 0.f;
  d_in->rough = d_rough * rough * (1.0f - rough);
  float d_nov = (in.nov >= 0.f && in.nov <= 1.f) ? d_u : 0.f;
  // schlick
  float d_sch = (sch >= 0.f && sch <= 1.f) ? d_rw : 0.f;
  float d_t = d_sch * 0.96f * 5.0f * t * t * t * t;
  if (t1 >= 0.f && t1 <= 1.f) d_nov -= d_t;
  d_in->nov = d_nov;
}

}  // namespace pw
}  // namespace nunerf
