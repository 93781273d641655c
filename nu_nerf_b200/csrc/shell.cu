// shell.cu -- the non-zero-thickness bounce of network/renderer.py:1690-2009, one thread per hit ray,
// forward and hand-derived reverse (pw::shell_bounce_fwd / _bwd in pointwise.cuh: the same source is compiled for the host by
// tests/hostsim/shell_host.cpp and checked there against the torch restatement nu_nerf_b200/shell.py and its autograd, which
// in turn is pinned bounce by bounce to the unmodified reference).  A few thousand rays per launch: latency-, not
// bandwidth-bound; the point is ONE launch per bounce instead of ~120 element-wise ones.
#include "common.cuh"
#include "pointwise.cuh"

namespace nunerf {

__device__ __forceinline__ void load_shell(const float* x, const float* n, const float* d, const float* gk,
                                           const float* ior_sig, const float* th_sig, int m, pw::ShellIn* in) {
#pragma unroll
  for (int c = 0; c < 3; ++c) { in->x[c] = x[3 * m + c]; in->n[c] = n[3 * m + c]; in->d[c] = d[3 * m + c]; }
  in->gk = gk[m]; in->ior_sig = ior_sig[m]; in->th_sig = th_sig[m];
}

__global__ void shell_bounce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ n, const float* __restrict__ d,
                                        const float* __restrict__ gk, const float* __restrict__ ior_sig,
                                        const float* __restrict__ th_sig, int M, int inside, uint8_t* ok, uint8_t* tir,
                                        float* x_mod, float* start, float* dir, float* ratio) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  pw::ShellIn in;
  pw::ShellOut o;
  load_shell(x, n, d, gk, ior_sig, th_sig, m, &in);
  pw::shell_bounce_fwd(in, inside, &o);
  ok[m] = (uint8_t)o.ok; tir[m] = (uint8_t)o.tir; ratio[m] = o.ratio;
#pragma unroll
  for (int c = 0; c < 3; ++c) { x_mod[3 * m + c] = o.x_mod[c]; start[3 * m + c] = o.start[c]; dir[3 * m + c] = o.dir[c]; }
}

__global__ void shell_bounce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ n, const float* __restrict__ d,
                                        const float* __restrict__ gk, const float* __restrict__ ior_sig,
                                        const float* __restrict__ th_sig, const uint8_t* __restrict__ ok, int M, int inside,
                                        const float* __restrict__ g_start, const float* __restrict__ g_dir,
                                        const float* __restrict__ g_ratio, const float* __restrict__ g_xmod, float* d_x,
                                        float* d_n, float* d_d, float* d_gk, float* d_ior, float* d_th) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  pw::ShellIn in, di;
  if (!ok[m]) {
    // a ray stopped by total internal reflection: only x_mod = x is a function of the inputs
#pragma unroll
    for (int c = 0; c < 3; ++c) { d_x[3 * m + c] = g_xmod[3 * m + c]; d_n[3 * m + c] = 0.f; d_d[3 * m + c] = 0.f; }
    d_gk[m] = 0.f; d_ior[m] = 0.f; d_th[m] = 0.f;
    return;
  }
  load_shell(x, n, d, gk, ior_sig, th_sig, m, &in);
  pw::shell_bounce_bwd(in, inside, g_start + 3 * m, g_dir + 3 * m, g_ratio[m], g_xmod + 3 * m, &di);
#pragma unroll
  for (int c = 0; c < 3; ++c) { d_x[3 * m + c] = di.x[c]; d_n[3 * m + c] = di.n[c]; d_d[3 * m + c] = di.d[c]; }
  d_gk[m] = di.gk; d_ior[m] = di.ior_sig; d_th[m] = di.th_sig;
}

}  // namespace nunerf

using namespace nunerf;

extern "C" int nunerf_shell_bounce(const float* x_hit, const float* n_signed, const float* rays_d, const float* g_k,
                                   const float* ior_sig, const float* thick_sig, int M, int inside, uint8_t* pass,
                                   uint8_t* tir, float* x_mod, float* o_next, float* d_next, float* ratio, void* stream) {
  NUNERF_REQUIRE(x_hit && n_signed && rays_d && g_k && ior_sig && thick_sig && pass && tir && x_mod && o_next && d_next &&
                     ratio && M > 0, "shell_bounce: bad arguments");
  shell_bounce_fwd_kernel<<<cdiv(M, 128), 128, 0, (cudaStream_t)stream>>>(x_hit, n_signed, rays_d, g_k, ior_sig, thick_sig, M,
                                                                         inside, pass, tir, x_mod, o_next, d_next, ratio);
  NUNERF_CHECK_LAUNCH("shell_bounce_fwd_kernel");
  return 0;
}

extern "C" int nunerf_shell_bounce_bwd(const float* x_hit, const float* n_signed, const float* rays_d, const float* g_k,
                                       const float* ior_sig, const float* thick_sig, const uint8_t* pass, int M, int inside,
                                       const float* g_onext, const float* g_dnext, const float* g_ratio, const float* g_xmod,
                                       float* d_x, float* d_n, float* d_d, float* d_gk, float* d_ior, float* d_thick,
                                       void* stream) {
  NUNERF_REQUIRE(x_hit && n_signed && rays_d && g_k && ior_sig && thick_sig && pass && g_onext && g_dnext && g_ratio &&
                     g_xmod && d_x && d_n && d_d && d_gk && d_ior && d_thick && M > 0, "shell_bounce_bwd: bad arguments");
  shell_bounce_bwd_kernel<<<cdiv(M, 128), 128, 0, (cudaStream_t)stream>>>(x_hit, n_signed, rays_d, g_k, ior_sig, thick_sig,
                                                                         pass, M, inside, g_onext, g_dnext, g_ratio, g_xmod,
                                                                         d_x, d_n, d_d, d_gk, d_ior, d_thick);
  NUNERF_CHECK_LAUNCH("shell_bounce_bwd_kernel");
  return 0;
}
