/* TEST INFRASTRUCTURE ONLY -- plain-C oracle for the bit-exact parts of the NU-NeRF hot path.
 *
 * Restates, on the CPU and independently of the CUDA sources, the reference algorithms whose results are
 * integer / index valued (sample indices, merge permutation, triangle hit ids) in the ONE arithmetic order
 * the build fixes for them (SURVEY.md section 7 hard part 3): strict fp32, no FMA contraction (compile with
 * -ffp-contract=off), own exp() polynomial, and the scan orders written out below.  Reference code followed:
 *   ray set-up      network/renderer_zerothick.py:320-327, 580-594
 *   up-sample round network/renderer_zerothick.py:525-554 ; sample_pdf network/field.py:468-498
 *   merge           network/renderer_zerothick.py:556-561 (stable: old samples first on ties)
 *   closest hit     network/DiffRender.py:61-92 (Moeller-Trumbore op order), cuda/triangle.cu:48-99 (miss id)
 * Pinned by tests/test_oracle_golden.py against fixtures produced by the unmodified reference
 * (index flips vs stock torch are counted and must be ulp-level ties).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float f_exp(float x) {
  /* Cody-Waite reduction + degree-7 Horner, every op a rounded fp32 op */
  if (x < -87.0f) return 0.0f;
  if (x > 88.0f) x = 88.0f;
  float t = x * 1.44269504088896341f;
  float n = (t >= 0.0f) ? (float)(int)(t + 0.5f) : (float)(int)(t - 0.5f);
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float p = 1.0f / 5040.0f;
  p = p * r + 1.0f / 720.0f;
  p = p * r + 1.0f / 120.0f;
  p = p * r + 1.0f / 24.0f;
  p = p * r + 1.0f / 6.0f;
  p = p * r + 0.5f;
  p = p * r + 1.0f;
  p = p * r + 1.0f;
  int e = (int)n + 127;
  if (e <= 0) return 0.0f;
  union { uint32_t u; float f; } s;
  s.u = (uint32_t)e << 23;
  return p * s.f;
}

static float f_sigmoid(float x) {
  if (x >= 0.0f) return 1.0f / (1.0f + f_exp(-x));
  float e = f_exp(x);
  return e / (1.0f + e);
}

/* Hillis-Steele inclusive scan over 32 lanes (simultaneous update per step) */
static void hs32_mul(float* x) {
  float y[32];
  for (int off = 1; off < 32; off <<= 1) {
    for (int i = 0; i < 32; ++i) y[i] = i >= off ? x[i - off] * x[i] : x[i];
    memcpy(x, y, sizeof(y));
  }
}
static void hs32_add(float* x) {
  float y[32];
  for (int off = 1; off < 32; off <<= 1) {
    for (int i = 0; i < 32; ++i) y[i] = i >= off ? x[i - off] + x[i] : x[i];
    memcpy(x, y, sizeof(y));
  }
}
static float xor_reduce(float* x) {
  float y[32];
  for (int off = 16; off > 0; off >>= 1) {
    for (int i = 0; i < 32; ++i) y[i] = x[i] + x[i ^ off];
    memcpy(x, y, sizeof(y));
  }
  return x[0];
}

/* tables: [0,64) linspace(0,1,64) | [64,96) bg lower | [96,128) bg upper-lower | [128,160) bg unperturbed */
void oracle_ray_setup(const float* o, const float* d, float* near, float* far, const float* U0, const float* U1,
                      const float* tables, int R, int sphere, int perturb, float* z, float* z_bg) {
  for (int r = 0; r < R; ++r) {
    float nr, fr;
    if (sphere) {
      const float* oo = o + 3 * r; const float* dd = d + 3 * r;
      float a = (dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2];
      float b = 2.0f * ((oo[0] * dd[0] + oo[1] * dd[1]) + oo[2] * dd[2]);
      float mid = (0.5f * -b) / a;
      nr = fmaxf(mid - 1.0f, 1e-3f);
      fr = mid + 1.0f;
      near[r] = nr; far[r] = fr;
    } else { nr = near[r]; fr = far[r]; }
    float span = fr - nr;
    float shift = perturb ? ((U0[r] - 0.5f) * 2.0f) / 64.0f : 0.0f;
    for (int j = 0; j < 64; ++j) {
      float v = nr + span * tables[j];
      if (perturb) v = v + shift;
      z[r * 64 + j] = v;
    }
    for (int j = 0; j < 32; ++j) {
      int jj = 31 - j;
      float b = perturb ? tables[64 + jj] + tables[96 + jj] * U1[r * 32 + jj] : tables[128 + jj];
      z_bg[r * 32 + j] = fr / b + 0.03125f;
    }
  }
}

/* one importance round; n <= 128, n_new <= 32 */
void oracle_upsample(const float* o, const float* d, const float* z, const float* sdf, int R, int n, int n_new,
                     float inv_s_full, float inv_s_cap, const float* u_tab, float* z_new, int32_t* inds,
                     float* z_merged, int32_t* perm) {
  const float inv_s = fminf(inv_s_full, inv_s_cap);
  for (int r = 0; r < R; ++r) {
    const float* zr = z + (long)r * n; const float* sr = sdf + (long)r * n;
    const float* oo = o + 3 * r; const float* dd = d + 3 * r;
    float zv[128], sv[128], rad[128], wgt[128], cdf[129];
    for (int j = 0; j < 128; ++j) {
      zv[j] = j < n ? zr[j] : 0.0f; sv[j] = j < n ? sr[j] : 0.0f;
      float px = oo[0] + dd[0] * zv[j], py = oo[1] + dd[1] * zv[j], pz = oo[2] + dd[2] * zv[j];
      rad[j] = sqrtf((px * px + py * py) + pz * pz);
    }
    float carryT = 1.0f, lane_sum[32];
    for (int i = 0; i < 32; ++i) lane_sum[i] = 0.0f;
    float prev_cos_raw = 0.0f;
    for (int k = 0; k < 4; ++k) {
      float v[32], alpha[32];
      int ok[32];
      for (int l = 0; l < 32; ++l) {
        int j = 32 * k + l;
        ok[l] = j < n - 1;
        float zn = j + 1 < 128 ? zv[j + 1] : 0.0f, sn = j + 1 < 128 ? sv[j + 1] : 0.0f, rn = j + 1 < 128 ? rad[j + 1] : 0.0f;
        float dist = zn - zv[j];
        float cosv = (sn - sv[j]) / (dist + 1e-5f);
        float c = fminf(prev_cos_raw, cosv);
        prev_cos_raw = cosv;
        c = fminf(fmaxf(c, -1e3f), 0.0f);
        int inside = (rad[j] < 1.0f) || (rn < 1.0f);
        if (!inside) c = c * 0.0f;
        float mid = (sv[j] + sn) * 0.5f;
        float half = (c * dist) * 0.5f;
        float pe = mid - half, ne = mid + half;
        float pc = f_sigmoid(pe * inv_s), nc = f_sigmoid(ne * inv_s);
        float a = ((pc - nc) + 1e-5f) / (pc + 1e-5f);
        alpha[l] = ok[l] ? a : 0.0f;
        v[l] = ok[l] ? (1.0f - a) + 1e-7f : 1.0f;
      }
      float incl[32];
      memcpy(incl, v, sizeof(v));
      hs32_mul(incl);
      for (int l = 0; l < 32; ++l) {
        float excl = l == 0 ? 1.0f : incl[l - 1];
        float T = carryT * excl;
        int j = 32 * k + l;
        wgt[j] = ok[l] ? alpha[l] * T + 1e-5f : 0.0f;
        lane_sum[l] = lane_sum[l] + wgt[j];
      }
      carryT = carryT * incl[31];
    }
    float total = xor_reduce(lane_sum);
    float carryC = 0.0f;
    cdf[0] = 0.0f;
    for (int k = 0; k < 4; ++k) {
      float pdf[32];
      for (int l = 0; l < 32; ++l) { int j = 32 * k + l; pdf[l] = j < n - 1 ? wgt[j] / total : 0.0f; }
      hs32_add(pdf);
      float last = 0.0f;
      for (int l = 0; l < 32; ++l) {
        int j = 32 * k + l;
        float c = carryC + pdf[l];
        if (j < n - 1) cdf[j + 1] = c;
        if (l == 31) last = c;
      }
      carryC = last;
    }
    float zs[32];
    for (int t = 0; t < n_new; ++t) {
      float u = u_tab[t];
      int lo = 0, hi = n;
      while (lo < hi) { int m = (lo + hi) >> 1; if (cdf[m] <= u) lo = m + 1; else hi = m; }
      int ind = lo;
      int below = ind - 1 < 0 ? 0 : ind - 1, above = ind > n - 1 ? n - 1 : ind;
      float c0 = cdf[below], c1 = cdf[above], b0 = zv[below], b1 = zv[above];
      float den = c1 - c0;
      if (den < 1e-5f) den = 1.0f;
      float tt = (u - c0) / den;
      zs[t] = b0 + tt * (b1 - b0);
      z_new[(long)r * n_new + t] = zs[t];
      inds[(long)r * n_new + t] = ind;
    }
    int nm = n + n_new;
    for (int j = 0; j < n; ++j) {
      int cnt = 0;
      for (int t = 0; t < n_new; ++t) cnt += zs[t] < zv[j];
      z_merged[(long)r * nm + j + cnt] = zv[j];
      perm[(long)r * nm + j + cnt] = j;
    }
    for (int t = 0; t < n_new; ++t) {
      int cnt = 0;
      for (int j = 0; j < n; ++j) cnt += zv[j] <= zs[t];
      z_merged[(long)r * nm + t + cnt] = zs[t];
      perm[(long)r * nm + t + cnt] = n + t;
    }
  }
}

/* compositing forward (tolerance-checked, sequential order): ZT:773-788 */
void oracle_composite(const float* alpha, const float* color, const uint8_t* inner, int R, int S, int is_nerf,
                      float* rgb, float* acc, float* rgb_b, float* weights) {
  for (int r = 0; r < R; ++r) {
    double T = 1.0, Tb = 1.0, c[3] = {0, 0, 0}, cb[3] = {0, 0, 0}, a_sum = 0;
    for (int s = 0; s < S; ++s) {
      long i = (long)r * S + s;
      double a = alpha[i], ab = inner[i] ? 0.0 : a;
      double w = a * T, wb = ab * Tb;
      if (weights) weights[i] = (float)w;
      for (int k = 0; k < 3; ++k) { c[k] += w * color[3 * i + k]; cb[k] += wb * color[3 * i + k]; }
      a_sum += w;
      T *= (1.0 - a + 1e-7); Tb *= (1.0 - ab + 1e-7);
    }
    for (int k = 0; k < 3; ++k) {
      double v = c[k] + (is_nerf ? 1.0 - a_sum : 0.0);
      rgb[3 * r + k] = (float)(v < 0 ? 0 : (v > 1 ? 1 : v));
      rgb_b[3 * r + k] = (float)cb[k];
    }
    acc[r] = (float)a_sum;
  }
}

/* brute-force closest hit: Moeller-Trumbore in the op order of DiffRender.JIT_Dintersect */
static float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static void cross3(const float* a, const float* b, float* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
void oracle_closest_hit(const float* tri_verts, int F, const float* rays_o, const float* rays_d, int N, float tmax,
                        float* hit, int32_t* tri, float* tout, float* uv) {
  for (int i = 0; i < N; ++i) {
    const float* o = rays_o + 3 * i; const float* d = rays_d + 3 * i;
    float best = tmax, bu = 0, bv = 0;
    int id = 10000000;
    for (int f = 0; f < F; ++f) {
      const float* t9 = tri_verts + 9 * (long)f;
      float e1[3], e2[3], p[3], q[3], s[3];
      for (int c = 0; c < 3; ++c) { e1[c] = t9[3 + c] - t9[c]; e2[c] = t9[6 + c] - t9[c]; }
      cross3(d, e2, p);
      float det = dot3(e1, p);
      if (det == 0.0f) continue;
      float inv = 1.0f / det;
      for (int c = 0; c < 3; ++c) s[c] = o[c] - t9[c];
      float u = dot3(s, p) * inv;
      cross3(s, e1, q);
      float v = dot3(d, q) * inv;
      float t = dot3(e2, q) * inv;
      if (!(u >= 0.0f && v >= 0.0f && u + v <= 1.0f)) continue;
      if (!(t > 0.0f && t < tmax)) continue;
      if (t < best || (t == best && f < id)) { best = t; id = f; bu = u; bv = v; }
    }
    hit[i] = id != 10000000 ? 1.0f : 0.0f;
    tri[i] = id;
    if (tout) tout[i] = best;
    if (uv) { uv[2 * i] = bu; uv[2 * i + 1] = bv; }
  }
}
