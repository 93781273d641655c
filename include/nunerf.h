/* nunerf.h -- C-ABI of libnunerf_b200.so, the sm_100a engine behind the NU-NeRF renderer boundary.
 *
 * The reference (jjjkkyz/NU-NeRF) has no FFI of its own: its hot path is PyTorch eager code plus two
 * third-party native tracers.  Each entry point below names the reference code it replaces
 * (paths relative to the reference root; "ZT" = network/renderer_zerothick.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all buffers;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, negative = error; nunerf_last_error() gives the message (thread local);
 *   - row-major, fp32 unless stated.  "bf16 planes": a [rows, ld] bf16 matrix whose columns
 *     [c, c+K) hold hi = bf16(x) and, in the split (fp32-accurate) mode, columns [lo_off + c, ...)
 *     hold lo = bf16(x - hi); lo_off = 0 means single plane.
 */
#ifndef NUNERF_H
#define NUNERF_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* nunerf_last_error(void);
int nunerf_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long nunerf_launch_count(void);

/* ------------------------------------------------------------------ dense layers (tcgen05)
 * Replace every nn.Linear / F.linear on the path (cuBLAS in the reference): field.py:145 (SDFNetwork),
 * :273-287 (NeRFNetwork), :386-394 (make_predictor), and their autograd backward / double backward.
 */
typedef struct {
  /* C[M,N] = epi( A[M,K] * B[N,K]^T ), A and B bf16 planes, K-major */
  const void* A; int lda; int a_lo_off;
  const void* B; int ldb; int b_lo_off;
  int M, N, K;                               /* N <= 256, N % 16 == 0, K % 64 == 0 */
  const float* bias;                         /* [N] or NULL */
  int act;                                   /* 0 none, 1 relu, 2 softplus(beta=100) */
  const void* aux; int ldaux; int aux_lo_off;/* bf16 planes [M, >=N] or NULL */
  int aux_mode;                              /* 0 none, 1 *= (aux>0), 2 *= 1-exp(-100*aux), 3 *= bit of mask_in */
  const void* add; int ldadd; int add_lo_off;/* bf16 planes added after the mask, or NULL */
  const void* mask_in; int ldmask_in;        /* aux_mode 3: 1 bit / element, row pitch in bytes (multiple of 8) */
  void* mask_out; int ldmask_out;            /* optional: bit = (output > 0), same layout (ReLU layers) */
  float out_scale;                           /* applied to the accumulator product (0 or 1.0 = none) */
  void* out; int ldo; int out_lo_off;        /* bf16 planes or NULL */
  float* out_f32; int ldo32;                 /* optional fp32 copy or NULL */
  int n_store;                               /* store columns [0, n_store) only (0 = all) */
  int impl;                                  /* 0 tcgen05, 1 SIMT debug kernel */
} nunerf_linear_t;
int nunerf_linear(const nunerf_linear_t* p, void* stream);

typedef struct {
  /* dW[N,K] += dZ[M,N]^T * X[M,K]  (fp32 atomics; caller zeroes dW), both operands MN-major bf16 planes */
  const void* dZ; int ldz; int z_lo_off;
  const void* X; int ldx; int x_lo_off;
  int M, N, K;                               /* N <= 256, K % 64 == 0 */
  float* dW; int lddw;
  int impl;
  float* db;                                 /* optional: db[n] += sum_m dZ[m,n] (the bias gradient), fused */
} nunerf_dw_t;
int nunerf_linear_dw(const nunerf_dw_t* p, void* stream);

/* Fused SDF query (SDFNetwork.sdf, field.py:133-152): PE-6 computed in-kernel, 8 Softplus(beta=100) layers with the
 * skip concat, 1-row sdf head; the activations stay in shared memory / TMEM between layers (csrc/chain.cu).  Used for
 * the no-grad SDF passes of sample_ray (ZT:598, :563), extract_fields (field.py:1286-1307) and the occlusion probes
 * (field.py:524-554).  w[l]: K-major bf16 weight of lin_l zero padded to [256,64], [256,256] x2, [224,256], [256,256] x4
 * and (w[8]) the sdf row of lin8 padded to [16,256]; lin4 pre-scaled by 1/sqrt(2); bias[l] fp32 padded alike. */
#define NUNERF_CHAIN_MAX_LAYERS 10
typedef struct {
  const float* pts; int M;                   /* [M,3] */
  const void* w[9]; int ldw[9];
  const float* bias[9];
  float* sdf; int ld_sdf;                    /* sdf[m * ld_sdf] */
  void* timeline;                            /* NULL, or 512 int64 (device): clock64 stamps of CTA 0's second tile
                                                (development aid, tools/chain_timeline.py) */
} nunerf_sdf_infer_t;
int nunerf_sdf_infer(const nunerf_sdf_infer_t* p, void* stream);

/* Generic fused chain of dense layers on ONE 128-row tile pair per SM (csrc/chain.cu): layer l computes
 *   y = act(x_l W_l^T + bias) [. mask_in]   with x_0 = the bf16 input rows and x_{l+1} = y when `keep` (or `store`) is set;
 * a layer with keep = store = 0 (a narrow head) leaves the activation in place, so the next layer still reads x_l.
 * The activations never leave shared memory / TMEM unless `store` asks for a bf16 copy (needed by the weight-gradient
 * GEMMs of the training step).  Replaces the per-layer evaluation of make_predictor MLPs (field.py:371-408: forward
 * with ReLU masks kept, and the dX chain of their backward). bf16 single-plane operands only. */
typedef struct {
  const void* w; int ldw;                    /* K-major bf16 [N (padded to 16), K] */
  int N, K;                                  /* N % 16 == 0, <= 256; K % 64 == 0, <= 256 */
  int n_real;                                /* 0 = N; produced columns >= n_real are zeroed */
  const float* bias;                         /* NULL or fp32 [N] */
  int act;                                   /* 0 none, 1 relu, 2 softplus(beta = 100) */
  unsigned char* mask_out; int ldmask_out;   /* optional 1-bit (y > 0) mask, N/8 bytes per row */
  const unsigned char* mask_in; int ldmask_in; /* optional 1-bit multiplicative mask */
  void* store; int ld_store;                 /* optional bf16 [M, pad64(N)] copy of y (TMA store) */
  float* out32; int ldo32; int n32;          /* optional fp32 copy of the first n32 columns */
  int keep;                                  /* y becomes the next layer's input */
  int cat_pe;                                /* columns [n_real, 256) <- PE-6 columns of the point (SDF skip concat) */
  /* epilogues of the SDF network's reverse passes (plain 256-wide layers, no bias / act / masks), s = 1 - exp(-100 aux1):
   *   aux_mode 4: y = acc . s      5: y = acc . s, e_out = acc . aux2 . 100 (1 - s)      6: y = acc . s + aux2 */
  int aux_mode;
  const void* aux1; int ld_aux1;             /* bf16 [M, 256] */
  const void* aux2; int ld_aux2;
  void* e_out; int ld_e;
  int mask_perm;                             /* plain 256-wide ReLU layers: the mask (in or out) is in the kernel's thread
                                                order -- 2-byte word j*4 + c holds columns c*64 + j*16 .. +15 -- so that
                                                each thread moves its 64 bits with one 8-byte access */
} nunerf_chain_layer_t;
typedef struct {
  const void* x; int ldx; int K0;            /* bf16 [M, K0], K0 in {64,128,192,256}; or NULL: */
  const float* pts;                          /* [M,3]: x_0 = PE-6(pts) computed in-kernel (also feeds cat_pe layers) */
  int M, n_layers;
  nunerf_chain_layer_t layer[NUNERF_CHAIN_MAX_LAYERS];
  void* timeline;                            /* NULL, or 512 int64 (device) clock64 stamps (development aid) */
} nunerf_mlp_chain_t;
int nunerf_mlp_chain(const nunerf_mlp_chain_t* p, void* stream);

/* Development aid (tools/mma_probe.py): issue rate of tcgen05.mma on this part.  One thread per CTA issues `iters` groups
 * of four M128 x N x K16 instructions (mode 0: both operands from shared memory, 1: A from tensor memory) on `b_stages`
 * rotating weight blocks (dmode: how consecutive instructions rotate over accumulator regions); out[2*cta] = cycles until issued, out[2*cta+1] = cycles until complete. */
int nunerf_mma_probe(int mode, int N, int iters, int grid, int b_stages, int dmode, long long* out, void* stream);

/* out[n] += sum_m Z[m,n] (hi + lo) -- the bias gradient */
int nunerf_colsum(const void* Z, int ldz, int z_lo_off, int M, int N, float* out, void* stream);
/* fp32 [rows, cols] (ld_src) -> the block dst[row_off:+dst_rows, col_off:+dst_cols] of a bf16 plane matrix, zero
 * filled outside the source; optional transpose and scale (weights: W -> W_kmajor and W^T_kmajor). */
int nunerf_to_planes(const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                     void* dst, int dst_rows, int dst_cols, int ld, int lo_off, int col_off, int row_off, void* stream);
int nunerf_from_planes(const void* src, int rows, int cols, int ld, int lo_off, float* dst, int ld_dst, void* stream);
/* dst[:, col:col+width] = a (+ b), fp32 [M,C] -> planes, zero padded to `width` columns */
int nunerf_f32_to_planes(const float* a, int lda, const float* b, int ldb, int M, int C, int width, void* dst, int ld,
                         int lo, int col, void* stream);

/* ------------------------------------------------------------------ sampling (ZT:572-612, field.py:468-498)
 * tables (host-computed with torch.linspace, 160 floats): [0,64) linspace(0,1,64) | [64,96) bg lower |
 *   [96,128) bg upper-lower | [128,160) bg unperturbed.  u_tab: linspace(.5/n_new, 1-.5/n_new, n_new).
 * nunerf_ray_setup : near/far (ZT:320-327, when sphere != 0), 64 coarse z + 32 inverse-depth bg z (ZT:580-594).
 * nunerf_upsample  : one importance round (ZT:525-554 + sample_pdf(det) + sorted merge ZT:556-561):
 *                    z[R,n], sdf[R,n] -> z_new[R,n_new], inds[R,n_new], z_merged[R,n+n_new], perm[R,n+n_new];
 *                    inv_s = min(*inv_s_dev, inv_s_cap) (ZT:604-605).
 * nunerf_merge_sdf : sdf_merged = cat(sdf, sdf_new)[perm]  (ZT:564-568).
 * nunerf_points    : p = o + d * z (ZT:598, :559).
 */
int nunerf_ray_setup(const float* o, const float* d, float* near, float* far, const float* U0, const float* U1,
                     const float* tables, int R, int sphere, int perturb, float* z, float* z_bg, void* stream);
int nunerf_points(const float* o, const float* d, const float* z, int R, int n, float* pts, void* stream);
/* reverse of nunerf_points (stage-2 path points as a function of the refracted segment start / direction):
 * g_o[R,3] = sum_j g_pts[r,j], g_d[R,3] = sum_j z[r,j] g_pts[r,j] */
int nunerf_points_bwd(const float* g_pts, const float* z, int R, int n, float* g_o, float* g_d, void* stream);
int nunerf_upsample(const float* o, const float* d, const float* z, const float* sdf, int R, int n, int n_new,
                    const float* inv_s_dev, float inv_s_cap, const float* u_tab, float* z_new, int32_t* inds,
                    float* z_merged, int32_t* perm, void* stream);
int nunerf_merge_sdf(const float* sdf, const float* sdf_new, const int32_t* perm, int R, int n, int n_new,
                     float* sdf_merged, void* stream);
/* Occlusion probe along P rays (get_weights + sample_pdf of get_intersection, field.py:501-554; caller ZT:695-723):
 * weights from n <= 64 samples z / sdf [P,n] with the probe's alpha (logistic CDF at the section ends, zero where the SDF
 * does not decrease); z_new != NULL: CDF inversion of (weights + 1e-5) at the n_new <= 32 deterministic u's -> z_new
 * [P,n_new]; wsum != NULL: sum of the weights [P] (the hit probability of the second pass). */
int nunerf_probe_weights(const float* z, const float* sdf, int P, int n, const float* inv_s_dev, int n_new,
                         const float* u_tab, float* z_new, float* wsum, void* stream);

/* ray_inner[r] = number of samples of ray r whose mid-point lies inside the unit sphere (ZT:730-741: the inner mask of
 * render_core), computed by the geometry code of nunerf_render_geometry without the compaction.  S <= 160. */
int nunerf_inner_counts(const float* o, const float* d, const float* z, int R, int S, int32_t* ray_inner, void* stream);

/* Stage-2 per-segment compositing in linear colour (ZT:1942-1951) on dense rows alpha[N,S], sRGB colour[N,S,3], S <= 256:
 *   rgb_lin[N,3] = sum_j alpha_j prod_{i<j}(1 - alpha_i + 1e-7) srgb_to_linear(colour_j),  t_end[N] = prod_j(1 - alpha_j + 1e-7)
 * and the backward w.r.t. alpha / colour given dL/d rgb_lin and dL/d t_end (the transmittance is recomputed). */
int nunerf_seg_composite_fwd(const float* alpha, const float* color, int N, int S, float* rgb_lin, float* t_end, void* stream);
int nunerf_seg_composite_bwd(const float* alpha, const float* color, int N, int S, const float* g_rgb, const float* g_t,
                             float* d_alpha, float* d_color, void* stream);
/* NeRF-guided importance sampling of the rays that leave the scene (upsample_nerf + cat_z_vals_nerf, ZT:1367-1397):
 * z_merged[R, n + n_new] = sorted merge of z[R,n] with sample_pdf(z, (alpha T)[:, :-1], n_new, det); n <= 256, n_new <= 64 */
int nunerf_alpha_importance(const float* z, const float* alpha, int R, int n, int n_new, const float* u_tab,
                            float* z_merged, void* stream);

/* ------------------------------------------------------------------ render_core geometry + compositing
 * nunerf_render_geometry (ZT:730-741): dists, mid points, inner mask and the row-major compaction of the
 *   inner / outer sample sets (the order of the reference's boolean-mask indexing): slot[R*S] >= 0 -> index in
 *   the inner list, < 0 -> -1 - index in the outer list; counts[2] = {N_in, N_out}; ray_scratch: 2R ints.
 *   The compact arrays (capacity R*S each) receive points, dists, normalised ray dirs and the flat sample id.
 *   ray_map (optional, 10 ints per ray): the same map in compact form -- [0] / [1] index of the ray's first inner / outer
 *   sample in its list, [2 + k] bit mask of the inner samples among samples 32k .. 32k+31; the compositing kernels take
 *   either form.  dists, pts, slot, id_in, id_out may be NULL when not needed.
 * nunerf_composite_fwd/bwd (ZT:773-788): w = a * excl_cumprod(1-a+1e-7); rgb = clamp(sum w c (+1-acc)); acc;
 *   background-only composite.  Backward recomputes the transmittance instead of storing it.
 *   With the per-ray map, S <= 160 and all list pointers 16-byte aligned the persistent shared-memory-staged kernels
 *   run (a ray's list runs are fetched as whole 16-byte chunks: reads may touch the <= 3 elements that share a chunk
 *   with the run, writes never leave the run.  BUFFER CONTRACT: alpha_in / color_in / alpha_out / color_out (and the
 *   gradient lists of the backward) must be ALLOCATED with a size that is a multiple of 16 bytes -- the last chunk of the
 *   last ray is read whole, i.e. up to 12 bytes past the logical end of a list whose length is not a multiple of 4
 *   floats; the entry points cannot check allocation sizes.  Callers with exact-size buffers set
 *   NUNERF_COMPOSITE_LEGACY=1 or pad); any other case (slot form,
 *   S > 160, unaligned views) takes the lane-per-sample kernels.  `weights` (dense [R,S], optional) is only needed by
 *   the validation outputs; pass NULL in training.  NUNERF_COMPOSITE_LEGACY=1 forces the lane-per-sample kernels.
 */
int nunerf_render_geometry(const float* o, const float* d, const float* z, int R, int S, float* dists, float* pts,
                           int32_t* slot, int32_t* counts, int32_t* ray_scratch, float* pts_in, float* dists_in,
                           float* dirs_in, int32_t* id_in, float* pts_out, float* dists_out, float* dirs_out,
                           int32_t* id_out, int32_t* ray_map, void* stream);
int nunerf_composite_fwd(const float* alpha_in, const float* color_in, const float* alpha_out,
                         const float* color_out, const int32_t* slot, int R, int S, int is_nerf, float* rgb,
                         float* rgb_raw, float* acc, float* rgb_bkgr, float* weights, const int32_t* ray_map,
                         void* stream);
int nunerf_composite_bwd(const float* alpha_in, const float* color_in, const float* alpha_out,
                         const float* color_out, const int32_t* slot, int R, int S, int is_nerf, const float* rgb_raw,
                         const float* d_rgb, const float* d_acc, const float* d_rgb_bkgr, float* d_alpha_in,
                         float* d_color_in, float* d_alpha_out, float* d_color_out, const int32_t* ray_map,
                         void* stream);
int nunerf_scatter_rows(const float* src, int M, int C, const int32_t* sample_id, float* dst, void* stream);

/* ------------------------------------------------------------------ encodings + pointwise field math */
/* positional encoding (field.py:14-61) of x[M,d] into planes; columns >= d(1+2F) up to `width` are zeroed */
int nunerf_encode_pe(const float* x, int M, int d, int nfreq, void* dst, int ld, int lo_off, int col_off,
                     int row_off, int width, void* stream);
/* grad_x = J_pe(x)^T (ga + gb): last step of SDFNetwork.gradient (field.py:158-170); _bwd: J_pe(x) d_grad */
int nunerf_sdf_grad_pe(const float* x, const float* ga, int lda, const float* gb, int ldb, int M, float* grad,
                       void* stream);
int nunerf_sdf_grad_pe_bwd(const float* x, const float* dgrad, int M, void* d1, int ld1, int lo1, int col1, int width1,
                           void* d2, int ld2, int lo2, int col2, int width2, void* stream);

/* Input gradients of the encodings (the reference gets them from autograd; stage 2 needs them because the sample
 * positions depend on IORs_pred through the refracted path, ZT:1633-1684):
 *   nunerf_pe_bwd      dx[M,d] (+)= J_pe(x)^T (ga + gb) for a d-dimensional input (d <= 4) with nfreq frequencies
 *   nunerf_sdf_pe_hess dx[M,3] += d_grad . d/dx [J_pe6(x)^T (ga + gb)]: the x-dependence of SDFNetwork.gradient through
 *                      the Jacobian of the encoding itself (field.py:158-170, create_graph=True)
 *   nunerf_nerf_prep_bwd backward of (p / |p|, 1 / |p|) and views = -dirs (ZT:688-689) */
int nunerf_pe_bwd(const float* x, int d, int nfreq, const float* ga, int lda, const float* gb, int ldb, int M, float* dx,
                  int accumulate, void* stream);
int nunerf_sdf_pe_hess(const float* x, const float* ga, int lda, const float* gb, int ldb, const float* d_grad, int M,
                       float* dx, void* stream);
int nunerf_nerf_prep_bwd(const float* pts, const float* d_pts4, const float* d_views, int M, float* d_pts, float* d_dirs,
                         void* stream);

/* compute_sdf_alpha ZT:657-685 + eikonal term ZT:769 */
typedef struct {
  int M; float cos_anneal; const float* inv_s_dev;       /* exp(10*variance), clipped to [1e-6,1e6] in-kernel */
  const float* sdf; int ld_sdf; const float* grad; const float* dists; const float* dirs;
  float* alpha; float* grad_err;
  const float* d_alpha; const float* d_grad_err; float* d_sdf; float* d_grad; float* d_inv_s; /* d_inv_s NULL: frozen */
  float* d_dists; float* d_dirs;                         /* optional backward outputs [M], [M,3]: d alpha / d (dist, dir) --
                                                            stage 2, where the sample positions depend on the IoR network */
} nunerf_sdf_alpha_t;
int nunerf_sdf_alpha_fwd(const nunerf_sdf_alpha_t* p, void* stream);
int nunerf_sdf_alpha_bwd(const nunerf_sdf_alpha_t* p, void* stream);

/* NeRF++ glue: ZT:687-693 */
int nunerf_nerf_prep(const float* pts, const float* dirs, int M, float* pts4, float* views, void* stream);
int nunerf_nerf_out_fwd(const float* sigma, int ld_s, const float* rgb, int ld_c, const float* dists, int M,
                        float* alpha, float* color, void* stream);
int nunerf_nerf_out_bwd(const float* sigma, int ld_s, const float* rgb, int ld_c, const float* dists, int M,
                        const float* d_alpha, const float* d_color, void* d_sig, int ld_ds, int lo_ds, int col_ds,
                        void* d_rgb, int ld_dr, int lo_dr, int col_dr, void* stream);
/* same + d alpha / d dist -> d_dists[M] */
int nunerf_nerf_out_bwd_geo(const float* sigma, int ld_s, const float* rgb, int ld_c, const float* dists, int M,
                            const float* d_alpha, const float* d_color, void* d_sig, int ld_ds, int lo_ds, int col_ds,
                            void* d_rgb, int ld_dr, int lo_dr, int col_dr, float* d_dists, void* stream);

/* AppShadingNetwork.forward field.py:684-741: directions + encodings (IDE utils/ref_utils.py:85-114) */
typedef struct {
  int M;
  const float* pts; const float* grad; const float* dirs;   /* [M,3]; dirs = normalised ray direction (view = -dirs) */
  const float* rough_raw; int ld_rough;                      /* pre-sigmoid roughness head */
  void* x_outer; int ld_outer; int lo_outer;                 /* [3M, ld]: IDE(n,1) | IDE(r,rough) | IDE(r,0) */
  void* x_inner; int ld_inner; int lo_inner;                 /* [2M, ld]: PE6(p)++IDE(r,rough) | PE6(p)++IDE(r,0) */
  void* x_weight; int ld_weight; int lo_weight;              /* [M, ld]: PE6(p) ++ PE6(r) */
  void* x_refrac; int ld_refrac; int lo_refrac;              /* [M, ld]: PE6(p) ++ PE6(v) */
  float* nov;                                                /* [M] */
  /* backward */
  const float* d_x_outer; int ld_dxo;                        /* fp32 [3M, ld] (72 cols used) */
  const float* d_x_inner; int ld_dxi;                        /* fp32 [2M, ld] (cols 39..110 used) */
  const float* d_nov;
  float* d_grad;                                             /* [M,3] accumulated */
  float* d_rough_raw; int ld_drough;                         /* accumulated */
  float* refl;                                               /* optional forward output [M,3]: reflected direction
                                                                (occ_info['reflective'], field.py:688, :744) */
  /* optional position-gradient part of the backward (stage 2): with d_pts set, d p through PE6(p) of the inner- and
   * refraction-light inputs is ADDED to d_pts[M,3] and d ray-direction (reflected direction, NoV, PE6(v)) to d_dirs[M,3] */
  const float* d_x_refrac; int ld_dxr;                       /* fp32 [M, ld] (78 cols used) or NULL */
  float* d_pts; float* d_dirs;
  /* encoding frequencies of the network evaluated: 0 = the AppShadingNetwork defaults (6 / 6, field.py:562-567).  The
   * inner field of the non-zero-thickness stage 2, AppShadingNetwork_SpecInner (field.py:1320-1330), has
   * light_pos_freq 8 (the PE(p) block of x_inner / x_weight is 51 columns, the IDE block starts at column 51) and
   * refrac_freq 2 (x_refrac = PE2(p) ++ PE2(v), 30 columns).  Supported pairs: (6, 6) and (8, 2). */
  int pos_freq; int refrac_freq;
  /* the `sphere_direction` shader variant (field.py:594-597, :641-651, :675-680): x_outer rows are 192 columns,
   * [IDE(u, k) | IDE(q(p, u), k') | 0] with q the exit point of the ray (p, u) on the unit sphere; d_x_outer carries 144
   * gradient columns and the backward needs pts */
  int sphere_direction;
} nunerf_shade_encode_t;
int nunerf_shade_encode_fwd(const nunerf_shade_encode_t* p, void* stream);
/* IDE(x, kinv) of M unit directions at one constant roughness -> planes (per-ray specular probe, ZT:780) */
int nunerf_ide_encode(const float* x, int M, float kinv, void* dst, int ld, int lo, int col, void* stream);
int nunerf_shade_encode_bwd(const nunerf_shade_encode_t* p, void* stream);

/* material / light mixing, FG-LUT taps (dr.texture), sRGB (utils/raw_utils.py:5-12) */
typedef struct {
  int M; float exp_max;
  const float* metallic; const float* rough; const float* albedo; const float* trans; int ld_mat; /* raw heads */
  const float* outer; int ld_outer;       /* [3M, ld]: diffuse | direct | direct0 (3 cols) */
  const float* inner; int ld_inner;       /* [2M, ld]: indirect | indirect0 */
  const float* weight; int ld_weight;     /* [M, ld] */
  const float* refrac; int ld_refrac;     /* [M, ld] */
  const float* nov; const float* lut;     /* [M]; FG LUT [256,256,2] */
  float* color; float* trans_out; float* metallic_out; float* occ_prob;
  /* backward: dZ operands of the head layers as bf16 planes [*, ld_dz >= 64]: the kernel writes 64 columns of every row
   * (the values + zero padding), the buffers need no prior fill */
  const float* d_color; const float* d_trans_out; const float* d_metallic_out;
  void* dz_metallic; void* dz_albedo; void* dz_trans; void* dz_outer; void* dz_inner; void* dz_weight; void* dz_refrac;
  int ld_dz; int lo_dz;
  float* d_rough_raw; float* d_nov;       /* fp32 [M] */
  const float* d_occ_prob;                /* optional [M]: gradient w.r.t. the occ_prob output (occlusion loss) */
  /* clamp of the refraction-light head when it differs from exp_max (AppShadingNetwork_SpecInner: -0.2, field.py:1373) */
  float exp_max_refrac; int use_exp_max_refrac;
} nunerf_shade_mix_t;
int nunerf_shade_mix_fwd(const nunerf_shade_mix_t* p, void* stream);
int nunerf_shade_mix_bwd(const nunerf_shade_mix_t* p, void* stream);

/* SDF-network glue for the gradient pass and its reverse-over-reverse (field.py:158-170 with create_graph) */
int nunerf_rowvec_mask(const float* w, const void* a, int lda, int a_lo, int M, int N, void* out, int ldo, int o_lo,
                       void* stream);
int nunerf_sdf_skip_split(const float* u4, const void* a3, int lda, int a_lo, int M, void* gs3, int ldg, int g_lo,
                          float* g_skip, void* stream);
int nunerf_sdf_bwd2_ew(const void* gts, int ldt, int t_lo, const void* a, int lda, int a_lo, const void* gs, int ldg,
                       int g_lo, int M, int N, int n_real, void* u_next, int ldu, int u_lo, void* e, int lde, int e_lo,
                       void* stream);
/* Per-step weight pipeline (one launch each way instead of hundreds of framework kernels): weight-norm
 * W = g v / |v| (nn.utils.weight_norm, field.py:121-122, :386-394), optional row/column rotation + scale, conversion to
 * the bf16 plane operands (K-major W and W^T), padded bias copy; and the adjoint that turns effective-weight gradients
 * (prepared layout) into gradients of weight_v / weight_g / bias (or the plain weight), accumulated into .grad storage.
 * One thread block per emitted row: blk_desc[b] / blk_row[b] give the descriptor and the row of block b. */
typedef struct {
  const float* v; const float* g;            /* source [N,K] pitch ld; g NULL = plain weight */
  int N, K, ld; float scale;
  int row_rot, col_rot, src_row0, n_rows;    /* dst row r <- src row (src_row0 + r + row_rot) % N ; dst col c <- (c + col_rot) % K */
  void* wk; int wk_ld, wk_lo, wk_row_off;    /* K-major destination or NULL */
  void* wtk; int wtk_ld, wtk_lo, wtk_col_off;/* transposed destination or NULL */
  float* inv_norm; float* row_f32;           /* [N] 1/|v| (by source row) ; optional fp32 copy [n_rows,K] */
  const float* bias_src; float* bias_dst;    /* bias_dst[wk_row_off + r] = bias_src[src row] */
  const float* dW; int lddw, dw_row_off;     /* backward: gradient of the prepared matrix (NULL = none) */
  float* dv; float* dg;                      /* accumulated: d weight_v (or d weight) pitch ld ; d weight_g [N] */
  const float* db; float* dbias;             /* accumulated: dbias[src row] += db[dw_row_off + r] */
} nunerf_wdesc_t;
int nunerf_weights_prepare(const nunerf_wdesc_t* descs, const int32_t* blk_desc, const int32_t* blk_row, int n_blocks,
                           void* stream);
int nunerf_weights_backward(const nunerf_wdesc_t* descs, const int32_t* blk_desc, const int32_t* blk_row, int n_blocks,
                            void* stream);
/* torch.optim.Adam update (train/lr_common_manager.py:11-15) over a flat parameter vector */
int nunerf_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                int step, void* stream);

/* Traversal-stack pushes dropped since the library was loaded: always 0 for trees built by nunerf_bvh_build_host (which
 * refuses trees deeper than the stack allows); synchronises the device.  Replaces the silent clamp of the reference's
 * fixed-size stack (raytracing/src/bvh.cu:526-602). */
int nunerf_bvh_overflow_count(unsigned int* out);

/* ------------------------------------------------------------------ tracing (DiffRender.py:410-416, :61-125;
 * cuda/triangle.cu:48-99; raytracing/src/bvh.cu:259-301,526-602,694-713)
 * nunerf_bvh_build_host: host build of a 4-wide BVH (recursive median split on the axis of largest centroid
 *   variance, <= 4 triangles per leaf) into caller-provided host arrays; returns the node count (or negative).
 * nunerf_bvh_trace    : closest hit with t in (0, tmax), double sided; hit[N] in {0,1}, tri[N] = ORIGINAL face
 *   index (miss: 10000000), t[N] (miss: tmax).  Tie rule: min t, then min face index.
 * nunerf_trace_brute  : same contract, exhaustive (on-device cross-check).
 * nunerf_hit_interp   : DiffRender.py:61-125 re-intersection with the hit face: (u,v,t), x = o + t d, interpolated
 *   unit vertex normal (tri_normals [F,9] = per-corner angle-weighted vertex normals, DiffRender.py:342-359).
 * nunerf_refract_bounce: zero-thickness bounce ZT:1633-1684: eta = 1/(IoR(x)+1) (inverted when inside), TIR test
 *   eta^2 sin^2 > 0.999, Snell direction, next origin x + 1e-5 d', d' / (|d'| + 1e-4).
 * nunerf_hit_interp_bwd / nunerf_refract_bounce_bwd: the reverse of those two steps (the reference keeps them inside
 *   autograd, DiffRender.py:61-124 / ZT:1633-1684, so that IORs_pred is trained through the path geometry).
 *   hit_interp_bwd: rows with a valid `tri`; g_x = d loss / d (o + t d), g_n = d loss / d (signed unit normal, negated
 *   when `inside`) -> g_o, g_d.  refract_bounce_bwd: rows that PASSED the TIR test, compact; n_signed = the signed unit
 *   normal, eta_eff = the effective ratio (inverted when inside); g_onext / g_dnext = d loss / d (next origin,
 *   next direction) -> g_x, g_n, g_d [N,3] and g_eta [N].
 */
typedef struct { float lo[4][3]; float hi[4][3]; int32_t child[4]; int32_t count[4]; } nunerf_bvh_node_t;
int nunerf_bvh_build_host(const float* verts_host, int V, const int32_t* faces_host, int F,
                          nunerf_bvh_node_t* nodes_host, int max_nodes, int32_t* tri_order_host);
int nunerf_bvh_trace(const nunerf_bvh_node_t* nodes, const float* tri_verts, const int32_t* tri_order,
                     const float* rays_o, const float* rays_d, int N, float tmax, float* hit, int32_t* tri, float* t,
                     void* stream);
int nunerf_trace_brute(const float* tri_verts, int F, const float* rays_o, const float* rays_d, int N, float tmax,
                       float* hit, int32_t* tri, float* t, void* stream);
int nunerf_hit_interp(const float* tri_verts, const float* tri_normals, const int32_t* tri, const float* rays_o,
                      const float* rays_d, int N, float* uvt, float* x_hit, float* n_hit, void* stream);
int nunerf_hit_interp_bwd(const float* tri_verts, const float* tri_normals, const int32_t* tri, const float* rays_o,
                          const float* rays_d, int N, int inside, const float* g_x, const float* g_n, float* g_o, float* g_d,
                          void* stream);
int nunerf_refract_bounce_bwd(const float* n_signed, const float* rays_d, const float* eta_eff, int N, const float* g_onext,
                              const float* g_dnext, float* g_x, float* g_n, float* g_d, float* g_eta, void* stream);
int nunerf_refract_bounce(const float* x_hit, const float* n_hit, const float* rays_d, const float* eta,
                          const int32_t* tri, int N, int inside, float* d_out, float* o_out, uint8_t* pass,
                          void* stream);

/* ------------------------------------------------------------------ non-zero-thickness bounce (network/renderer.py:1690-2009)
 * One launch per bounce on the M rays that hit the mesh.  The mesh is one face of a glass shell of thickness
 * 0.01 * thick_sig; at a hit the shell is modelled as two concentric spheres of radius r = 1 / sqrt(max(|g_k|, 1e-6)) and
 * r -+ thickness (g_k = interpolated vertex Gaussian curvature, DiffRender.py:113-116): refraction into the glass with
 * ior = 1 / (ior_sig + 0.6) (NZ:1733-1734), chord through the shell, second refraction with (1 / 1.0001) / ior (NZ:1739-1744);
 * `inside` (the ray leaves the object): the ratios are inverted and swapped (NZ:1753-1756) and the hit is first pulled back
 * onto the inner face (NZ:1884-1933).  n_signed = unit vertex normal, negated when inside.
 *   pass[M]   the first refraction is no total internal reflection (ior^2 sin^2 <= 0.999; `converged_out`, NZ:1768)
 *   tir[M]    pass and neither later refraction was clamped (`tir`, NZ:1773, :1860, :1939, :1993)
 *   x_mod     the hit point (pulled back when inside); o_next / d_next / ratio: next segment's origin, direction and the
 *             effective ratio of the first refraction (`ior_ratios`) -- defined on rows with pass = 1
 * nunerf_shell_bounce_bwd: the reverse (the reference keeps the bounce inside autograd so that IORs_pred AND thickness_pred
 *   are trained through the path geometry): g_* = d loss / d (o_next, d_next, ratio, x_mod) over all M rows (ignored where
 *   pass = 0, except g_xmod) -> d loss / d (x_hit, n_signed, rays_d, g_k, ior_sig, thick_sig). */
int nunerf_shell_bounce(const float* x_hit, const float* n_signed, const float* rays_d, const float* g_k,
                        const float* ior_sig, const float* thick_sig, int M, int inside, uint8_t* pass, uint8_t* tir,
                        float* x_mod, float* o_next, float* d_next, float* ratio, void* stream);
int nunerf_shell_bounce_bwd(const float* x_hit, const float* n_signed, const float* rays_d, const float* g_k,
                            const float* ior_sig, const float* thick_sig, const uint8_t* pass, int M, int inside,
                            const float* g_onext, const float* g_dnext, const float* g_ratio, const float* g_xmod,
                            float* d_x, float* d_n, float* d_d, float* d_gk, float* d_ior, float* d_thick, void* stream);

/* ------------------------------------------------------------------ grid sweep (field.py:1286-1307)
 * grid_points: points start..start+count of the res^3 grid in x-major order (the order of torch.meshgrid 'ij' +
 *   reshape, field.py:1297-1300); lin = the three per-axis torch.linspace tables, [3, res].
 * grid_mask  : u = outside_val where |p| >= 1 else sdf (field.py:1303-1304). */
int nunerf_grid_points(int res, long long start, int count, const float* lin, float* pts, void* stream);
int nunerf_grid_mask(const float* pts, const float* sdf, int ld_sdf, int count, float outside_val, float* u,
                     void* stream);

/* ------------------------------------------------------------------ marching cubes (field.py:1312, replaces the
 * host-side mcubes.marching_cubes(u, threshold) of extract_geometry)
 * Cells are visited in lexicographic (x, y, z) order; a block owns 8 consecutive (x, y) cell rows, blocks numbered
 * row-major: nunerf_mc_blocks(res) = (res-1) * ceil((res-1) / 8) blocks.  res <= 2048.
 * mc_count : block_counts[b] = number of triangles of block b (int32, nunerf_mc_blocks(res) entries).
 * mc_emit  : block_offsets = exclusive prefix sum of block_counts (int64, computed by the caller); writes the triangle
 *   soup verts [3 * T, 3] (grid-index coordinates, cell order) and keys [3 * T] = global id of the grid edge carrying
 *   the vertex (((x0 * res + y0) * res + z0) * 3 + axis), so that equal keys are the same vertex.
 * Tables (nu_nerf_b200/mc_tables.py): tri_table int8 [256, 3 * max_tris], n_tris int8 [256], edges int8 [12, 2],
 *   edge_axis int8 [12]; bit k of the case index is set when corner k (offset (k&1, k>>1&1, k>>2&1)) is < iso. */
long long nunerf_mc_blocks(int res);
int nunerf_mc_count(const float* u, int res, float iso, const int8_t* n_tris, int32_t* block_counts, void* stream);
int nunerf_mc_emit(const float* u, int res, float iso, const int8_t* tri_table, int max_tris, const int8_t* n_tris,
                   const int8_t* edges, const int8_t* edge_axis, const long long* block_offsets, float* verts,
                   long long* keys, void* stream);

#ifdef __cplusplus
}
#endif
#endif
