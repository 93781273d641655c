// TEST INFRASTRUCTURE ONLY -- host build of pw::shell_bounce_fwd / _bwd (nu_nerf_b200/csrc/pointwise.cuh, the source the CUDA
// kernels of csrc/shell.cu compile) behind a C ABI for ctypes: tests/test_kernel_source_cpu.py checks it against the torch
// restatement nu_nerf_b200/shell.py (pinned to the unmodified reference) and against its autograd.
#include <cmath>
#include <cstdint>
#include "pointwise.cuh"
using namespace nunerf;
extern "C" void shell_fwd(const float* in, int M, int inside, float* out) {      // in [M,12], out [M,12]
  for (int m = 0; m < M; ++m) {
    pw::ShellIn a; pw::ShellOut o;
    const float* p = in + 12 * m;
    for (int c = 0; c < 3; ++c) { a.x[c] = p[c]; a.n[c] = p[3 + c]; a.d[c] = p[6 + c]; }
    a.gk = p[9]; a.ior_sig = p[10]; a.th_sig = p[11];
    pw::shell_bounce_fwd(a, inside, &o);
    float* q = out + 12 * m;
    q[0] = (float)o.ok; q[1] = (float)o.tir;
    for (int c = 0; c < 3; ++c) { q[2 + c] = o.x_mod[c]; q[5 + c] = o.start[c]; q[8 + c] = o.dir[c]; }
    q[11] = o.ratio;
  }
}
extern "C" void shell_bwd(const float* in, int M, int inside, const float* g, float* din) {   // g [M,10]: start, dir, ratio, xmod
  for (int m = 0; m < M; ++m) {
    pw::ShellIn a, d;
    const float* p = in + 12 * m;
    for (int c = 0; c < 3; ++c) { a.x[c] = p[c]; a.n[c] = p[3 + c]; a.d[c] = p[6 + c]; }
    a.gk = p[9]; a.ior_sig = p[10]; a.th_sig = p[11];
    const float* gg = g + 10 * m;
    pw::shell_bounce_bwd(a, inside, gg, gg + 3, gg[6], gg + 7, &d);
    float* q = din + 12 * m;
    for (int c = 0; c < 3; ++c) { q[c] = d.x[c]; q[3 + c] = d.n[c]; q[6 + c] = d.d[c]; }
    q[9] = d.gk; q[10] = d.ior_sig; q[11] = d.th_sig;
  }
}

// pw::sphere_dir_fwd / _bwd (the `sphere_direction` shader variant, field.py:447-465, :641-644) on arrays
extern "C" void sphere_dir(const float* p, const float* u, int M, float* q) {
  for (int m = 0; m < M; ++m) pw::sphere_dir_fwd(p + 3 * m, u + 3 * m, q + 3 * m);
}
extern "C" void sphere_dir_bwd(const float* p, const float* u, const float* dq, int M, float* du, float* dp) {
  for (int m = 0; m < M; ++m) {
    for (int c = 0; c < 3; ++c) { du[3 * m + c] = 0.f; dp[3 * m + c] = 0.f; }
    pw::sphere_dir_bwd(p + 3 * m, u + 3 * m, dq + 3 * m, du + 3 * m, dp + 3 * m);
  }
}

// pw::shade_mix_fwd / _bwd (AppShadingNetwork.forward mixing, field.py:684-741; exp_max_r = the refraction-light clamp of
// AppShadingNetwork_SpecInner, field.py:1373) on one point: in / d_in are the 26 floats of ShadeMixIn in declaration order
static void mix_in(const float* v, pw::ShadeMixIn* a) {
  int i = 0;
  a->metallic = v[i++]; a->rough = v[i++];
  for (int c = 0; c < 3; ++c) a->albedo[c] = v[i++];
  a->trans = v[i++];
  for (int c = 0; c < 3; ++c) a->diffuse_l[c] = v[i++];
  for (int c = 0; c < 3; ++c) a->direct[c] = v[i++];
  for (int c = 0; c < 3; ++c) a->direct0[c] = v[i++];
  for (int c = 0; c < 3; ++c) a->indirect[c] = v[i++];
  for (int c = 0; c < 3; ++c) a->indirect0[c] = v[i++];
  a->occ = v[i++];
  for (int c = 0; c < 3; ++c) a->refrac[c] = v[i++];
  a->nov = v[i++];
}
extern "C" void mix_fwd(const float* in, const float* lut, float exp_max, float exp_max_r, float* out) {   // out: rgb, T, met, occ
  pw::ShadeMixIn a;
  mix_in(in, &a);
  pw::ShadeMixOut o = pw::shade_mix_fwd(a, lut, exp_max, exp_max_r);
  out[0] = o.color[0]; out[1] = o.color[1]; out[2] = o.color[2]; out[3] = o.trans; out[4] = o.metallic; out[5] = o.occ_prob;
}
extern "C" void mix_bwd(const float* in, const float* lut, float exp_max, float exp_max_r, const float* d_color,
                        float d_trans, float d_met, float* d_in) {
  pw::ShadeMixIn a, d;
  mix_in(in, &a);
  pw::shade_mix_bwd(a, lut, exp_max, exp_max_r, d_color, d_trans, d_met, &d);
  int i = 0;
  d_in[i++] = d.metallic; d_in[i++] = d.rough;
  for (int c = 0; c < 3; ++c) d_in[i++] = d.albedo[c];
  d_in[i++] = d.trans;
  for (int c = 0; c < 3; ++c) d_in[i++] = d.diffuse_l[c];
  for (int c = 0; c < 3; ++c) d_in[i++] = d.direct[c];
  for (int c = 0; c < 3; ++c) d_in[i++] = d.direct0[c];
  for (int c = 0; c < 3; ++c) d_in[i++] = d.indirect[c];
  for (int c = 0; c < 3; ++c) d_in[i++] = d.indirect0[c];
  d_in[i++] = d.occ;
  for (int c = 0; c < 3; ++c) d_in[i++] = d.refrac[c];
  d_in[i++] = d.nov;
}

// pw::ide_fwd / ide_bwd (integrated directional encoding, utils/ref_utils.py:85-114) with the table build_ide_table() makes
// for the device's constant memory
extern "C" void ide_host(const float* xyz, const float* kinv, int M, float* out) {
  static pw::IdeTable tb;
  static bool init = false;
  if (!init) { pw::build_ide_table(&tb); init = true; }
  for (int m = 0; m < M; ++m) pw::ide_fwd(tb, xyz[3 * m], xyz[3 * m + 1], xyz[3 * m + 2], kinv[m], out + 72 * m);
}
extern "C" void ide_bwd_host(const float* xyz, const float* kinv, const float* dout, int M, float* dxyz, float* dk) {
  static pw::IdeTable tb;
  static bool init = false;
  if (!init) { pw::build_ide_table(&tb); init = true; }
  for (int m = 0; m < M; ++m)
    pw::ide_bwd(tb, xyz[3 * m], xyz[3 * m + 1], xyz[3 * m + 2], kinv[m], dout + 72 * m, dxyz + 3 * m, dxyz + 3 * m + 1,
                dxyz + 3 * m + 2, dk + m);
}

// pw::sdf_alpha_fwd / _bwd (compute_sdf_alpha ZT:657-685 + the eikonal term ZT:769) on arrays: in [M,9] = sdf, g[3], dist,
// dir[3], (unused); out [M,2] = alpha, gerr; bwd: cot [M,2] -> d [M,9] = d_sdf, d_g[3], d_dist, d_dir[3], d_inv_s
extern "C" void sdf_alpha_host(const float* in, int M, float inv_s, float anneal, float* out) {
  for (int m = 0; m < M; ++m) {
    const float* p = in + 9 * m;
    pw::SdfAlphaOut o = pw::sdf_alpha_fwd(p[0], p + 1, p[4], p + 5, inv_s, anneal);
    out[2 * m] = o.alpha; out[2 * m + 1] = o.gerr;
  }
}
extern "C" void sdf_alpha_bwd_host(const float* in, int M, float inv_s, float anneal, const float* cot, float* d) {
  for (int m = 0; m < M; ++m) {
    const float* p = in + 9 * m;
    float* q = d + 9 * m;
    pw::sdf_alpha_bwd(p[0], p + 1, p[4], p + 5, inv_s, anneal, cot[2 * m], cot[2 * m + 1], q, q + 1, q + 8, q + 4, q + 5);
  }
}

// pw::nerf_out_fwd / _bwd (compute_density_alpha ZT:687-693, density_activation ZT:515-516, linear_to_srgb) and
// pw::shade_dirs / shade_dirs_bwd (normal, view, reflected direction, NoV: field.py:686-689) on arrays
extern "C" void nerf_out_host(const float* in, int M, float* out) {          // in [M,5] = sigma, rgb[3], dist; out [M,4]
  for (int m = 0; m < M; ++m) pw::nerf_out_fwd(in[5 * m], in + 5 * m + 1, in[5 * m + 4], out + 4 * m, out + 4 * m + 1);
}
extern "C" void nerf_out_bwd_host(const float* in, int M, const float* cot, float* d) {     // cot [M,4] -> d [M,5]
  for (int m = 0; m < M; ++m)
    pw::nerf_out_bwd(in[5 * m], in + 5 * m + 1, in[5 * m + 4], cot[4 * m], cot + 4 * m + 1, d + 5 * m, d + 5 * m + 1,
                     d + 5 * m + 4);
}
extern "C" void shade_dirs_host(const float* g, const float* raydir, int M, float* out) {   // out [M,10] = n, v, r, nov
  for (int m = 0; m < M; ++m) {
    pw::ShadeDirs s = pw::shade_dirs(g + 3 * m, raydir + 3 * m);
    float* q = out + 10 * m;
    for (int c = 0; c < 3; ++c) { q[c] = s.n[c]; q[3 + c] = s.v[c]; q[6 + c] = s.r[c]; }
    q[9] = s.nov;
  }
}
// cot [M,10] = d_n (direct), d_v (direct), d_r, d_nov -> d_g [M,3], d_raydir [M,3]
extern "C" void shade_dirs_bwd_host(const float* g, const float* raydir, const float* cot, int M, float* d_g, float* d_rd) {
  for (int m = 0; m < M; ++m) {
    pw::ShadeDirs s = pw::shade_dirs(g + 3 * m, raydir + 3 * m);
    const float* c = cot + 10 * m;
    const float* rd = raydir + 3 * m;
    float vn = fmaxf(sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]), 1e-12f);
    pw::shade_dirs_bwd(s, c + 6, c, c[9], d_g + 3 * m, c + 3, vn, d_rd + 3 * m);
  }
}
