// bvh.cu -- ray / triangle-mesh closest hit for the stage-2 bounce (replaces the OptiX pipeline of
// network/tracing_optix.py + cuda/triangle.cu and the vendored raytracing/ BVH4 extension).
//
//   * host build: 4-wide BVH, recursive median split on the axis of largest centroid variance (the scheme of
//     raytracing/src/bvh.cu:526-602), <= 4 triangles per leaf, one 128-byte node = 4 child boxes + links;
//   * device traversal: one ray per thread, per-thread stack in shared memory laid out [depth][thread]
//     (bank-conflict free), children visited near-to-far, subtree pruned when its entry distance exceeds the
//     best hit; SoA-friendly ray I/O;
//   * closest-hit contract (the oracle's definition): Moeller-Trumbore in the operation order of
//     DiffRender.JIT_Dintersect (DiffRender.py:61-92), every op explicitly rounded (bit-exact against
//     oracle/sampling_oracle.c), double sided, u,v >= 0, u+v <= 1, 0 < t < tmax; ties: min t then min face id.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace nunerf {

constexpr int MISS_ID = 10000000;  // cuda/triangle.cu:85-89

struct Hit { float t; int id; };

__host__ __device__ __forceinline__ float dot3(const float* a, const float* b) {
  return det_add(det_add(det_mul(a[0], b[0]), det_mul(a[1], b[1])), det_mul(a[2], b[2]));
}
__host__ __device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
  c[0] = det_sub(det_mul(a[1], b[2]), det_mul(a[2], b[1]));
  c[1] = det_sub(det_mul(a[2], b[0]), det_mul(a[0], b[2]));
  c[2] = det_sub(det_mul(a[0], b[1]), det_mul(a[1], b[0]));
}
// returns true on a valid hit; u, v, t as JIT_Dintersect
__host__ __device__ __forceinline__ bool moller_trumbore(const float* o, const float* d, const float* tv9, float* u,
                                                         float* v, float* t) {
  float e1[3], e2[3], p[3], q[3], s[3];
  for (int c = 0; c < 3; ++c) { e1[c] = det_sub(tv9[3 + c], tv9[c]); e2[c] = det_sub(tv9[6 + c], tv9[c]); }
  cross3(d, e2, p);
  float det = dot3(e1, p);
  if (det == 0.0f) return false;
  float inv = det_div(1.0f, det);
  for (int c = 0; c < 3; ++c) s[c] = det_sub(o[c], tv9[c]);
  *u = det_mul(dot3(s, p), inv);
  cross3(s, e1, q);
  *v = det_mul(dot3(d, q), inv);
  *t = det_mul(dot3(e2, q), inv);
  return (*u >= 0.0f) && (*v >= 0.0f) && (det_add(*u, *v) <= 1.0f);
}

__device__ __forceinline__ void consider(const float* o, const float* d, const float* tv9, int face, float tmax,
                                         Hit* best) {
  float u, v, t;
  if (!moller_trumbore(o, d, tv9, &u, &v, &t)) return;
  if (!(t > 0.0f && t < tmax)) return;
  if (t < best->t || (t == best->t && face < best->id)) { best->t = t; best->id = face; }
}

// ------------------------------------------------------------------------------------------- traversal
constexpr int TRACE_THREADS = 128;
constexpr int STACK_DEPTH = 48;
// A pop pushes at most the four children of the node, so a tree of depth D needs 3 (D - 1) + 4 entries.  The builder
// splits at the median (depth <= 15 for any int32 triangle count) and REFUSES trees that would not fit; should a foreign
// node array overflow anyway, the push is counted here (nunerf_bvh_overflow_count) instead of vanishing silently.
__device__ unsigned int g_bvh_overflow = 0;

__global__ void __launch_bounds__(TRACE_THREADS)
bvh_trace_kernel(const nunerf_bvh_node_t* __restrict__ nodes, const float* __restrict__ tri_verts,
                 const int32_t* __restrict__ tri_order, const float* __restrict__ rays_o,
                 const float* __restrict__ rays_d, int N, float tmax, float* hit, int32_t* tri, float* tout) {
  __shared__ int s_stack[STACK_DEPTH][TRACE_THREADS];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * TRACE_THREADS + tid;
  if (i >= N) return;
  const float o[3] = {rays_o[3 * i], rays_o[3 * i + 1], rays_o[3 * i + 2]};
  const float d[3] = {rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]};
  float inv[3];
  for (int c = 0; c < 3; ++c) inv[c] = 1.0f / d[c];  // +-inf on axis-parallel rays is what the slab test wants
  Hit best = {tmax, MISS_ID};
  int sp = 0;
  s_stack[sp++][tid] = 0;
  while (sp > 0) {
    const int ni = s_stack[--sp][tid];
    const nunerf_bvh_node_t* nd = nodes + ni;
    float tn[4];
    int ord[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ord[k] = k;
      tn[k] = INFINITY;
      if (nd->count[k] < 0) continue;
      float t0 = 0.0f, t1 = best.t;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = (nd->lo[k][c] - o[c]) * inv[c], b = (nd->hi[k][c] - o[c]) * inv[c];
        // NaN (0 * inf) must not cull: fmin/fmax drop NaNs
        t0 = fmaxf(t0, fminf(a, b));
        t1 = fminf(t1, fmaxf(a, b));
      }
      if (t0 <= t1) tn[k] = t0;
    }
    // sorting network on 4 keys (near first)
#define CSWAP(a, b) if (tn[ord[a]] > tn[ord[b]]) { int t_ = ord[a]; ord[a] = ord[b]; ord[b] = t_; }
    CSWAP(0, 1) CSWAP(2, 3) CSWAP(0, 2) CSWAP(1, 3) CSWAP(1, 2)
#undef CSWAP
    // leaves first (they can only shrink best.t), then push inner children far-to-near
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      int k = ord[kk];
      if (tn[k] == INFINITY || nd->count[k] <= 0) continue;
      if (tn[k] > best.t) continue;
      int first = nd->child[k];
      for (int j = 0; j < nd->count[k]; ++j) consider(o, d, tri_verts + 9 * (long long)(first + j), tri_order[first + j], tmax, &best);
    }
#pragma unroll
    for (int kk = 3; kk >= 0; --kk) {
      int k = ord[kk];
      if (tn[k] == INFINITY || nd->count[k] != 0) continue;
      if (tn[k] > best.t) continue;
      if (sp < STACK_DEPTH) s_stack[sp++][tid] = nd->child[k];
      else atomicAdd(&g_bvh_overflow, 1u);
    }
  }
  hit[i] = best.id != MISS_ID ? 1.0f : 0.0f;
  tri[i] = best.id;
  if (tout) tout[i] = best.t;
}

__global__ void trace_brute_kernel(const float* __restrict__ tri_verts, int F, const float* __restrict__ rays_o,
                                   const float* __restrict__ rays_d, int N, float tmax, float* hit, int32_t* tri,
                                   float* tout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float o[3] = {rays_o[3 * i], rays_o[3 * i + 1], rays_o[3 * i + 2]};
  const float d[3] = {rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]};
  Hit best = {tmax, MISS_ID};
  for (int f = 0; f < F; ++f) consider(o, d, tri_verts + 9 * (long long)f, f, tmax, &best);
  hit[i] = best.id != MISS_ID ? 1.0f : 0.0f;
  tri[i] = best.id;
  if (tout) tout[i] = best.t;
}

// re-intersection with the hit triangle + interpolated unit normal (DiffRender.py:61-111, Intersection :302-311)
__global__ void hit_interp_kernel(const float* __restrict__ tri_verts, const float* __restrict__ tri_normals,
                                  const int32_t* __restrict__ tri, const float* __restrict__ rays_o,
                                  const float* __restrict__ rays_d, int N, float* uvt, float* x_hit, float* n_hit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int f = tri[i];
  if (f < 0 || f >= MISS_ID) {
    for (int c = 0; c < 3; ++c) { uvt[3 * i + c] = 0.f; x_hit[3 * i + c] = 0.f; n_hit[3 * i + c] = 0.f; }
    return;
  }
  const float o[3] = {rays_o[3 * i], rays_o[3 * i + 1], rays_o[3 * i + 2]};
  const float d[3] = {rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]};
  float u, v, t;
  moller_trumbore(o, d, tri_verts + 9 * (long long)f, &u, &v, &t);
  const float* nn = tri_normals + 9 * (long long)f;
  float w = 1.0f - u - v, n[3];
  for (int c = 0; c < 3; ++c) n[c] = w * nn[c] + u * nn[3 + c] + v * nn[6 + c];
  float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  uvt[3 * i] = u; uvt[3 * i + 1] = v; uvt[3 * i + 2] = t;
  for (int c = 0; c < 3; ++c) { n_hit[3 * i + c] = n[c] / len; x_hit[3 * i + c] = o[c] + t * d[c]; }
}

// zero-thickness bounce (ZT:1633-1684): Snell refraction with TIR test, next ray
__global__ void refract_bounce_kernel(const float* __restrict__ x_hit, const float* __restrict__ n_hit,
                                      const float* __restrict__ rays_d, const float* __restrict__ eta_in,
                                      const int32_t* __restrict__ tri, int N, int inside, float* d_out, float* o_out,
                                      uint8_t* pass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int f = tri[i];
  if (f < 0 || f >= MISS_ID) { pass[i] = 0; return; }
  float n[3], d[3] = {rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]};
  float len = fmaxf(sqrtf(n_hit[3 * i] * n_hit[3 * i] + n_hit[3 * i + 1] * n_hit[3 * i + 1] + n_hit[3 * i + 2] * n_hit[3 * i + 2]), 1e-12f);
  for (int c = 0; c < 3; ++c) n[c] = (inside ? -1.0f : 1.0f) * n_hit[3 * i + c] / len;
  float cosi = -(n[0] * d[0] + n[1] * d[1] + n[2] * d[2]);
  float sin2 = 1.0f - cosi * cosi;
  float eta = eta_in[i];            // 1 / (IoR_net(x) + 1)   (ZT:1642-1643)
  if (inside) eta = 1.0f / eta;     // ZT:1653-1654
  bool ok = !(eta * eta * sin2 > 0.999f);
  pass[i] = ok ? 1 : 0;
  if (!ok) return;
  float sin2t = sin2 * eta * eta;
  float k = eta * cosi - sqrtf(1.0f - sin2t);
  float nd[3];
  for (int c = 0; c < 3; ++c) nd[c] = eta * d[c] + k * n[c];
  float nl = sqrtf(nd[0] * nd[0] + nd[1] * nd[1] + nd[2] * nd[2]) + 0.0001f;
  for (int c = 0; c < 3; ++c) {
    o_out[3 * i + c] = x_hit[3 * i + c] + nd[c] * 1e-5f;
    d_out[3 * i + c] = nd[c] / nl;
  }
}

// ---- reverse of hit_interp_kernel (the differentiable ray-triangle intersection of DiffRender.py:61-124, which the
// reference keeps inside autograd so that the IoR network is trained through the refracted path geometry).
// Inputs: gradient arriving at x = o + t d and at the signed unit normal sign * normalize(interp(u, v)); outputs:
// gradient with respect to the ray origin and direction (the triangle is a constant).
__global__ void hit_interp_bwd_kernel(const float* __restrict__ tri_verts, const float* __restrict__ tri_normals,
                                      const int32_t* __restrict__ tri, const float* __restrict__ rays_o,
                                      const float* __restrict__ rays_d, int N, float sign, const float* __restrict__ g_x,
                                      const float* __restrict__ g_n, float* g_o, float* g_d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int f = tri[i];
  if (f < 0 || f >= MISS_ID) {
    for (int c = 0; c < 3; ++c) { g_o[3 * i + c] = 0.f; g_d[3 * i + c] = 0.f; }
    return;
  }
  const float* tv = tri_verts + 9 * (long long)f;
  const float* nn = tri_normals + 9 * (long long)f;
  const float o[3] = {rays_o[3 * i], rays_o[3 * i + 1], rays_o[3 * i + 2]};
  const float d[3] = {rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]};
  float e1[3], e2[3], p[3], q[3], s[3];
  for (int c = 0; c < 3; ++c) { e1[c] = tv[3 + c] - tv[c]; e2[c] = tv[6 + c] - tv[c]; s[c] = o[c] - tv[c]; }
  cross3(d, e2, p);
  const float inv = 1.0f / dot3(e1, p);
  cross3(s, e1, q);
  const float sp = dot3(s, p), dq = dot3(d, q), eq = dot3(e2, q);
  const float u = sp * inv, v = dq * inv, t = eq * inv;
  // ---- interpolated normal n = ni / |ni| (the second F.normalize of the reference is the identity)
  float ni[3], n[3];
  for (int c = 0; c < 3; ++c) ni[c] = (1.0f - u - v) * nn[c] + u * nn[3 + c] + v * nn[6 + c];
  const float len = sqrtf(dot3(ni, ni));
  for (int c = 0; c < 3; ++c) n[c] = ni[c] / len;
  const float gx[3] = {g_x[3 * i], g_x[3 * i + 1], g_x[3 * i + 2]};
  const float gn[3] = {sign * g_n[3 * i], sign * g_n[3 * i + 1], sign * g_n[3 * i + 2]};
  const float ngn = dot3(n, gn);
  float gni[3];
  for (int c = 0; c < 3; ++c) gni[c] = (gn[c] - n[c] * ngn) / len;
  float gu = 0.f, gv = 0.f;
  for (int c = 0; c < 3; ++c) { gu += gni[c] * (nn[3 + c] - nn[c]); gv += gni[c] * (nn[6 + c] - nn[c]); }
  // ---- x = o + t d
  const float gt = dot3(gx, d);
  float go[3] = {gx[0], gx[1], gx[2]}, gd[3] = {t * gx[0], t * gx[1], t * gx[2]};
  // ---- Moeller-Trumbore: u = (s . p) inv, v = (d . q) inv, t = (e2 . q) inv, p = d x e2, q = s x e1, inv = 1 / (e1 . p)
  const float g_inv = gu * sp + gv * dq + gt * eq;
  const float g_det = -g_inv * inv * inv;
  float gp[3], gq[3], tmp[3];
  for (int c = 0; c < 3; ++c) {
    gp[c] = gu * inv * s[c] + g_det * e1[c];
    gq[c] = gv * inv * d[c] + gt * inv * e2[c];
  }
  cross3(e1, gq, tmp);                       // q = s x e1  ->  g_s += e1 x g_q
  for (int c = 0; c < 3; ++c) go[c] += gu * inv * p[c] + tmp[c];
  cross3(e2, gp, tmp);                       // p = d x e2  ->  g_d += e2 x g_p
  for (int c = 0; c < 3; ++c) gd[c] += gv * inv * q[c] + tmp[c];
  for (int c = 0; c < 3; ++c) { g_o[3 * i + c] = go[c]; g_d[3 * i + c] = gd[c]; }
}

// ---- reverse of the Snell step of refract_bounce_kernel (ZT:1633-1684) for rays that pass the TIR test:
//   d' = eta d + (eta cos_i - sqrt(1 - (1 - cos_i^2) eta^2)) n,  o_next = x + 1e-5 d',  d_next = d' / (|d'| + 1e-4)
// with n the signed unit normal, cos_i = -n . d and eta the effective ratio (already inverted when inside).
// Outputs: gradient with respect to x, n, d and eta.
__global__ void refract_bounce_bwd_kernel(const float* __restrict__ n_in, const float* __restrict__ rays_d,
                                          const float* __restrict__ eta_in, int N, const float* __restrict__ g_onext,
                                          const float* __restrict__ g_dnext, float* g_x, float* g_n, float* g_d,
                                          float* g_eta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float n[3] = {n_in[3 * i], n_in[3 * i + 1], n_in[3 * i + 2]};
  const float d[3] = {rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]};
  const float eta = eta_in[i];
  const float cosi = -dot3(n, d);
  const float sin2 = 1.0f - cosi * cosi;
  const float r = sqrtf(1.0f - sin2 * eta * eta);
  const float k = eta * cosi - r;
  float nd[3];
  for (int c = 0; c < 3; ++c) nd[c] = eta * d[c] + k * n[c];
  const float L = sqrtf(dot3(nd, nd)), Le = L + 0.0001f;
  const float gs[3] = {g_onext[3 * i], g_onext[3 * i + 1], g_onext[3 * i + 2]};
  const float gdn[3] = {g_dnext[3 * i], g_dnext[3 * i + 1], g_dnext[3 * i + 2]};
  const float proj = dot3(gdn, nd) / (Le * Le * L);
  float gnd[3];
  for (int c = 0; c < 3; ++c) gnd[c] = gdn[c] / Le - nd[c] * proj + 1e-5f * gs[c];
  const float gk = dot3(gnd, n);
  const float ge = dot3(gnd, d) + gk * (cosi + sin2 * eta / r);
  const float gcos = gk * (eta - cosi * eta * eta / r);
  for (int c = 0; c < 3; ++c) {
    g_x[3 * i + c] = gs[c];
    g_n[3 * i + c] = k * gnd[c] - gcos * d[c];
    g_d[3 * i + c] = eta * gnd[c] - gcos * n[c];
  }
  g_eta[i] = ge;
}

// ------------------------------------------------------------------------------------------- host build
struct BuildCtx {
  const float* verts; const int32_t* faces;
  std::vector<float> cent;  // [F,3]
  std::vector<int> order;
  nunerf_bvh_node_t* nodes; int max_nodes; int n_nodes;
  int depth;                // deepest node level reached (root = 1)
};

static void tri_bounds(const BuildCtx& c, int b, int e, float* lo, float* hi) {
  for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
  for (int i = b; i < e; ++i) {
    const int32_t* f = c.faces + 3 * c.order[i];
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) {
        float v = c.verts[3 * f[j] + k];
        lo[k] = std::min(lo[k], v); hi[k] = std::max(hi[k], v);
      }
  }
  for (int k = 0; k < 3; ++k) {  // conservative padding: the slab test must never lose a boundary triangle
    float pad = 1e-5f + 1e-5f * std::max(std::fabs(lo[k]), std::fabs(hi[k]));
    lo[k] -= pad; hi[k] += pad;
  }
}

static int split_median(BuildCtx& c, int b, int e) {
  double mean[3] = {0, 0, 0}, var[3] = {0, 0, 0};
  for (int i = b; i < e; ++i) for (int k = 0; k < 3; ++k) mean[k] += c.cent[3 * c.order[i] + k];
  for (int k = 0; k < 3; ++k) mean[k] /= (e - b);
  for (int i = b; i < e; ++i) for (int k = 0; k < 3; ++k) { double d = c.cent[3 * c.order[i] + k] - mean[k]; var[k] += d * d; }
  int ax = var[0] >= var[1] ? (var[0] >= var[2] ? 0 : 2) : (var[1] >= var[2] ? 1 : 2);
  int mid = (b + e) / 2;
  std::nth_element(c.order.begin() + b, c.order.begin() + mid, c.order.begin() + e, [&](int x, int y) {
    float cx = c.cent[3 * x + ax], cy = c.cent[3 * y + ax];
    return cx < cy || (cx == cy && x < y);
  });
  return mid;
}

static int build_node(BuildCtx& c, int b, int e, int level = 1) {
  if (c.n_nodes >= c.max_nodes) return -1;
  int me = c.n_nodes++;
  c.depth = std::max(c.depth, level);
  int rb[4], re[4], nr = 0;
  if (e - b <= 4) { rb[0] = b; re[0] = e; nr = 1; }
  else {
    int m = split_median(c, b, e);
    int halves[2][2] = {{b, m}, {m, e}};
    for (auto& h : halves) {
      if (h[1] - h[0] <= 4) { rb[nr] = h[0]; re[nr] = h[1]; nr++; }
      else { int mm = split_median(c, h[0], h[1]); rb[nr] = h[0]; re[nr] = mm; nr++; rb[nr] = mm; re[nr] = h[1]; nr++; }
    }
  }
  for (int k = 0; k < 4; ++k) {
    nunerf_bvh_node_t& nd = c.nodes[me];
    if (k >= nr) { nd.count[k] = -1; nd.child[k] = -1; for (int j = 0; j < 3; ++j) { nd.lo[k][j] = 0; nd.hi[k][j] = 0; } continue; }
    float lo[3], hi[3];
    tri_bounds(c, rb[k], re[k], lo, hi);
    for (int j = 0; j < 3; ++j) { c.nodes[me].lo[k][j] = lo[j]; c.nodes[me].hi[k][j] = hi[j]; }
    if (re[k] - rb[k] <= 4) { c.nodes[me].count[k] = re[k] - rb[k]; c.nodes[me].child[k] = rb[k]; }
    else {
      int ch = build_node(c, rb[k], re[k], level + 1);
      if (ch < 0) return -1;
      c.nodes[me].count[k] = 0; c.nodes[me].child[k] = ch;
    }
  }
  return me;
}

}  // namespace nunerf

using namespace nunerf;

extern "C" int nunerf_bvh_build_host(const float* verts, int V, const int32_t* faces, int F, nunerf_bvh_node_t* nodes,
                                     int max_nodes, int32_t* tri_order) {
  NUNERF_REQUIRE(verts && faces && nodes && tri_order && V > 0 && F > 0 && max_nodes > 0, "bvh_build: bad arguments");
  for (int i = 0; i < 3 * F; ++i) NUNERF_REQUIRE(faces[i] >= 0 && faces[i] < V, "bvh_build: face index out of range");
  BuildCtx c;
  c.verts = verts; c.faces = faces; c.nodes = nodes; c.max_nodes = max_nodes; c.n_nodes = 0; c.depth = 0;
  c.cent.resize(3 * (size_t)F);
  c.order.resize(F);
  for (int f = 0; f < F; ++f) {
    c.order[f] = f;
    for (int k = 0; k < 3; ++k)
      c.cent[3 * f + k] = (verts[3 * faces[3 * f] + k] + verts[3 * faces[3 * f + 1] + k] + verts[3 * faces[3 * f + 2] + k]) / 3.0f;
  }
  int root = build_node(c, 0, F);
  NUNERF_REQUIRE(root == 0, "bvh_build: node capacity exceeded");
  NUNERF_REQUIRE(3 * (c.depth - 1) + 4 <= STACK_DEPTH, "bvh_build: tree too deep for the traversal stack");
  for (int f = 0; f < F; ++f) tri_order[f] = c.order[f];
  return c.n_nodes;
}

extern "C" int nunerf_bvh_trace(const nunerf_bvh_node_t* nodes, const float* tri_verts, const int32_t* tri_order,
                                const float* rays_o, const float* rays_d, int N, float tmax, float* hit, int32_t* tri,
                                float* t, void* stream) {
  NUNERF_REQUIRE(nodes && tri_verts && tri_order && rays_o && rays_d && hit && tri && N > 0, "bvh_trace: bad arguments");
  bvh_trace_kernel<<<cdiv(N, TRACE_THREADS), TRACE_THREADS, 0, (cudaStream_t)stream>>>(nodes, tri_verts, tri_order,
                                                                                     rays_o, rays_d, N, tmax, hit, tri, t);
  NUNERF_CHECK_LAUNCH("bvh_trace_kernel");
  return 0;
}

// number of traversal-stack pushes dropped since the library was loaded (0 for every tree nunerf_bvh_build_host accepts);
// synchronises the device
extern "C" int nunerf_bvh_overflow_count(unsigned int* out) {
  NUNERF_REQUIRE(out, "bvh_overflow_count: bad arguments");
  cudaError_t e = cudaMemcpyFromSymbol(out, g_bvh_overflow, sizeof(unsigned int));
  if (e != cudaSuccess) return fail("bvh_overflow_count: %s", cudaGetErrorString(e), -2);
  return 0;
}

extern "C" int nunerf_trace_brute(const float* tri_verts, int F, const float* rays_o, const float* rays_d, int N,
                                  float tmax, float* hit, int32_t* tri, float* t, void* stream) {
  NUNERF_REQUIRE(tri_verts && rays_o && rays_d && hit && tri && N > 0 && F > 0, "trace_brute: bad arguments");
  trace_brute_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(tri_verts, F, rays_o, rays_d, N, tmax, hit, tri, t);
  NUNERF_CHECK_LAUNCH("trace_brute_kernel");
  return 0;
}

extern "C" int nunerf_hit_interp(const float* tri_verts, const float* tri_normals, const int32_t* tri,
                                 const float* rays_o, const float* rays_d, int N, float* uvt, float* x_hit, float* n_hit,
                                 void* stream) {
  NUNERF_REQUIRE(tri_verts && tri_normals && tri && rays_o && rays_d && uvt && x_hit && n_hit && N > 0,
                 "hit_interp: bad arguments");
  hit_interp_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(tri_verts, tri_normals, tri, rays_o, rays_d, N, uvt,
                                                                   x_hit, n_hit);
  NUNERF_CHECK_LAUNCH("hit_interp_kernel");
  return 0;
}

extern "C" int nunerf_hit_interp_bwd(const float* tri_verts, const float* tri_normals, const int32_t* tri, const float* rays_o,
                                     const float* rays_d, int N, int inside, const float* g_x, const float* g_n, float* g_o,
                                     float* g_d, void* stream) {
  NUNERF_REQUIRE(tri_verts && tri_normals && tri && rays_o && rays_d && g_x && g_n && g_o && g_d && N > 0,
                 "hit_interp_bwd: bad arguments");
  hit_interp_bwd_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(tri_verts, tri_normals, tri, rays_o, rays_d, N,
                                                                       inside ? -1.0f : 1.0f, g_x, g_n, g_o, g_d);
  NUNERF_CHECK_LAUNCH("hit_interp_bwd_kernel");
  return 0;
}

extern "C" int nunerf_refract_bounce_bwd(const float* n_signed, const float* rays_d, const float* eta_eff, int N,
                                         const float* g_onext, const float* g_dnext, float* g_x, float* g_n, float* g_d,
                                         float* g_eta, void* stream) {
  NUNERF_REQUIRE(n_signed && rays_d && eta_eff && g_onext && g_dnext && g_x && g_n && g_d && g_eta && N > 0,
                 "refract_bounce_bwd: bad arguments");
  refract_bounce_bwd_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(n_signed, rays_d, eta_eff, N, g_onext, g_dnext,
                                                                           g_x, g_n, g_d, g_eta);
  NUNERF_CHECK_LAUNCH("refract_bounce_bwd_kernel");
  return 0;
}

extern "C" int nunerf_refract_bounce(const float* x_hit, const float* n_hit, const float* rays_d, const float* eta,
                                     const int32_t* tri, int N, int inside, float* d_out, float* o_out, uint8_t* pass,
                                     void* stream) {
  NUNERF_REQUIRE(x_hit && n_hit && rays_d && eta && tri && d_out && o_out && pass && N > 0, "refract_bounce: bad arguments");
  refract_bounce_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(x_hit, n_hit, rays_d, eta, tri, N, inside, d_out,
                                                                       o_out, pass);
  NUNERF_CHECK_LAUNCH("refract_bounce_kernel");
  return 0;
}
