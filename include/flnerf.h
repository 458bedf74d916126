/* flnerf.h -- C ABI of libflnerf.so: the B200 (sm_100a) hot path of Fast-Learning-NeRF.
 *
 * The reference (wen-yuan-zhang/Fast-Learning-NeRF, nerf-ours/) has no FFI of its own: the path
 * is Python functions wired by name (run_nerf.py:21-33).  Each entry point below replaces the
 * arithmetic of the cited reference function; the Python modules in fast-learning-nerf_b200/
 * (render.py, run_nerf_helpers.py, model.py, tree.py) keep the reference signatures and bind
 * these symbols with ctypes (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the parameter name starts with h_ (host);
 *  - all work is enqueued on the caller's cudaStream_t (passed as void*), nothing synchronises;
 *  - the library never allocates or frees caller memory; scratch comes from caller workspaces
 *    sized by the *_bytes() queries;
 *  - return value: 0 = OK, non-zero = error, text via flnerf_last_error() (thread-local);
 *  - fp32 unless stated; "rays11" is [B,11] = o(3) d(3) near far viewdir(3) (render.py:74-80).
 */
#ifndef FLNERF_H
#define FLNERF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flnerf_ctx flnerf_ctx;

/* MLP arithmetic modes (model.py:38-63 is fp32; see DESIGN.md "precision modes") */
#define FLNERF_MODE_FP32 0 /* CUDA-core fp32 GEMMs: the parity path (<=1e-4 rel of the reference)      */
#define FLNERF_MODE_BF16 1 /* tcgen05 bf16 x bf16 -> fp32 TMEM accumulators: the throughput path        */
#define FLNERF_MODE_BF16X3 2 /* tcgen05, every operand split into bf16 hi + lo, products hi*hi + lo*hi + hi*lo into the
                              * same fp32 accumulator: the tensor-core parity path (<=1e-4 rel of the reference)  */

#define FLNERF_MLP_PARAMS 595844 /* parameters of one NeRF(D=8,W=256,skips=[4],use_viewdirs) (model.py:20-34) */
#define FLNERF_TILE_ROWS 128
#define FLNERF_PAIR_ROWS 256     /* the tensor-core kernels process rows in pairs of 128-row tiles */

int flnerf_version(void);
const char *flnerf_last_error(void);
flnerf_ctx *flnerf_create(int device);
void flnerf_destroy(flnerf_ctx *ctx);
int flnerf_sm_count(flnerf_ctx *ctx);

/* ---- a1: get_rays (run_nerf_helpers.py:68-78). h_K = 3x3 row-major doubles, h_c2w = 3x4 floats. */
int flnerf_raygen(flnerf_ctx *, int H, int W, const double *h_K, const float *h_c2w, float *rays_o, float *rays_d,
                  void *stream);

/* ---- a2+a3: ndc_rays + ray packing (run_nerf_helpers.py:91-108, render.py:59-80). */
int flnerf_pack_rays(flnerf_ctx *, int64_t B, const float *rays_o, const float *rays_d, float near_, float far_,
                     int ndc, int H, int W, double focal, float *rays11, void *stream);

/* ---- a4: stratified depths (render.py:244-266). t_vals = linspace(0,1,Nc) [Nc]; t_rand [B,Nc] or NULL
 * (NULL with perturb!=0 draws Philox uniforms from (seed, offset)). */
int flnerf_coarse_depths(flnerf_ctx *, int64_t B, int Nc, const float *rays11, const float *t_vals,
                         const float *t_rand, int perturb, int lindisp, uint64_t seed, uint64_t offset, float *z,
                         void *stream);

/* ---- a5: positional encoding (run_nerf_helpers.py:15-63, run_nerf.py:50-64). */
/* generic embedder: x[n,3] -> out[n,3+6L] */
int flnerf_posenc(flnerf_ctx *, int64_t n, int L, const float *x, float *out, void *stream);
/* fused sample-point + PE, reference layout: x90[B*S,90] = [PE10(o+d*z), PE4(viewdir)] */
int flnerf_encode_f32(flnerf_ctx *, int64_t B, int S, const float *rays11, const float *z, float *x90, void *stream);
/* fused sample-point + PE for the tensor-core MLP: pe_tiles = ceil(B*S/256)*2 tiles of [128 rows x 64] bf16 in
 * the SWIZZLE_128B K-major shared-memory image (16 KB each, column 63 = 0); dirpe[B,32] fp32 (27 used). */
int flnerf_encode_tc(flnerf_ctx *, int64_t B, int S, const float *rays11, const float *z, void *pe_tiles,
                     float *dirpe, void *stream);
/* the eval path (render.py:94-146 -> render(c2w=...)): ONE launch from the camera pose to the coarse network's input -- for the
 * B pixels [pixel0, pixel0 + B) of an H x W frame: get_rays + (ndc) + ray packing -> rays11[B,11]; un-jittered coarse depths
 * z[B,S] (render.py:244-249; t_vals = linspace(0,1,S)); sample points + PE tiles (x3 != 0: hi and lo tile sets, as
 * flnerf_encode_tc_x3) and dirpe[B,32].  Replaces flnerf_raygen + flnerf_pack_rays + flnerf_coarse_depths + flnerf_encode_tc. */
int flnerf_encode_frame_tc(flnerf_ctx *, int x3, int H, int W, const double *h_K, const float *h_c2w, float near_, float far_,
                           int ndc, int lindisp, int64_t pixel0, int64_t B, int S, const float *t_vals, float *rays11, float *z,
                           void *pe_tiles, float *dirpe, void *stream);
/* already-embedded rows x90[n,90] (NeRF.forward API) -> pe_tiles + dirpe[n,32] (one "ray" per row, S = 1) */
int flnerf_pack_x90(flnerf_ctx *, int64_t n, const float *x90, void *pe_tiles, float *dirpe, void *stream);
/* the same two for FLNERF_MODE_BF16X3: pe_tiles holds TWO tile sets back to back, hi = bf16(PE) then lo = bf16(PE - hi)
 * (2 x 16 KB per 128 rows), and every octave takes an exact sincosf */
int flnerf_encode_tc_x3(flnerf_ctx *, int64_t B, int S, const float *rays11, const float *z, void *pe_tiles,
                        float *dirpe, void *stream);
int flnerf_pack_x90_x3(flnerf_ctx *, int64_t n, const float *x90, void *pe_tiles, float *dirpe, void *stream);
int64_t flnerf_padded_rows(int64_t n);                  /* n rounded up to FLNERF_PAIR_ROWS */

/* ---- a6: NeRF MLP (model.py:38-63).  params/grads: flat fp32[FLNERF_MLP_PARAMS] in parameters() order:
 * pts_linears.0..7 {weight,bias}, views_linears.0, feature_linear, alpha_linear, rgb_linear. */
/* scratch + activations kept for backward.  S = samples per ray (n = B*S rows, ray-major); training=0 sizes the
 * buffer for inference (FP32 mode always keeps the per-layer activations). */
size_t flnerf_mlp_stash_bytes(int mode, int64_t n, int S, int training);
size_t flnerf_mlp_bwd_workspace_bytes(int mode, int64_t n);
size_t flnerf_mlp_packed_bytes(void);                   /* tensor-core weight image of one net (hi part, then lo part) */
/* fp32 master weights -> pre-swizzled bf16 chunks (forward and transposed for dgrad), hi = bf16(w) and lo = bf16(w - hi) */
int flnerf_mlp_pack_weights(flnerf_ctx *, const float *params, void *packed, void *stream);
/* mode FP32: x = x90 fp32 [n,90];  modes BF16 / BF16X3: x = pe_tiles (of flnerf_encode_tc / flnerf_encode_tc_x3),
 * dirpe[B,32], S = samples per ray (row -> ray = row/S).
 * raw_out [n,4] = (r,g,b,sigma).  stash: flnerf_mlp_stash_bytes(mode, n, S, training) bytes, 1 KB aligned. */
int flnerf_mlp_forward(flnerf_ctx *, int mode, const float *params, const void *packed, int64_t n, int S,
                       const void *x, const float *dirpe, float *raw_out, void *stash, int training, void *stream);
/* grads += d(loss)/d(params) given draw[n,4]; no input gradient is produced (z, pts carry no grad: render.py:281) */
int flnerf_mlp_backward(flnerf_ctx *, int mode, const float *params, const void *packed, int64_t n, int S,
                        const void *x, const float *dirpe, const void *stash, const float *draw, float *grads,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ---- the same tensor-core MLP for a network with in_pts position channels: 63 (model.NeRF, the nerf++ foreground MLPNet) or
 * 84 (the nerf++ background MLPNet, PE of (x,y,z,1/r): nerf++-ours/nerf_network.py:70-142 -- two 64-wide input slabs; layer 5
 * is 340 wide).  params / grads: flat fp32[flnerf_mlp_param_count_g(in_pts, 27)] in the order above; packed image of
 * flnerf_mlp_packed_bytes_g(in_pts) bytes; x = pe_tiles with ceil(in_pts/64) slabs per 128-row tile as written by
 * flnerf_pack_xrows (mode BF16X3: hi set then lo set); stash / workspace sizes are those of flnerf_mlp_stash_bytes(mode, ..). */
size_t flnerf_mlp_packed_bytes_g(int in_pts);
int flnerf_mlp_pack_weights_g(flnerf_ctx *, int in_pts, const float *params, void *packed, void *stream);
int flnerf_mlp_forward_g(flnerf_ctx *, int mode, int in_pts, const float *params, const void *packed, int64_t n, int S,
                         const void *x, const float *dirpe, float *raw_out, void *stash, int training, void *stream);
int flnerf_mlp_backward_g(flnerf_ctx *, int mode, int in_pts, const float *params, const void *packed, int64_t n, int S,
                          const void *x, const float *dirpe, const void *stash, const float *draw, float *grads, void *workspace,
                          size_t workspace_bytes, void *stream);
/* already-embedded fp32 rows x[n, in_pts + 27] -> pe_tiles (ceil(in_pts/64) slabs per tile; x3 != 0: + the lo set) and, from the
 * first row of every S-row ray, dirpe[n/S, 32] */
int flnerf_pack_xrows(flnerf_ctx *, int x3, int in_pts, int in_views, int64_t n, int S, const float *x, void *pe_tiles,
                      float *dirpe, void *stream);

/* ---- the same MLP with a caller-chosen number of position / view channels, FP32 (CUDA-core) path only: in_pts = 63,
 * in_views = 27 is the nerf-ours network; in_pts = 84 is the nerf++ background network (nerf++-ours/nerf_network.py:70-142,
 * PE of (x,y,z,1/r)).  params / grads: flat fp32[flnerf_mlp_param_count_g()] in the order above; x fp32 [n, in_pts+in_views];
 * stash / workspace sizes are those of FLNERF_MODE_FP32. */
int64_t flnerf_mlp_param_count_g(int in_pts, int in_views);
int flnerf_mlp_fp32_forward_g(flnerf_ctx *, int in_pts, int in_views, const float *params, int64_t n, const float *x,
                              float *raw_out, void *stash, void *stream);
int flnerf_mlp_fp32_backward_g(flnerf_ctx *, int in_pts, int in_views, const float *params, int64_t n, const float *x,
                               const void *stash, const float *draw, float *grads, void *workspace, size_t workspace_bytes,
                               void *stream);

/* same, running only the selected backward kernels (bit 0: data-gradient chain, bit 1: weight gradients, bit 2:
 * rgb/view-direction heads) -- used by bench.py to time each kernel with CUDA events; FP32 mode ignores it */
int flnerf_mlp_backward_stages(flnerf_ctx *, int mode, const float *params, const void *packed, int64_t n, int S,
                               const void *x, const float *dirpe, const void *stash, const float *draw, float *grads,
                               void *workspace, size_t workspace_bytes, int stages, void *stream);

/* ---- a7: raw2outputs (render.py:149-192) and its backward. One warp per ray. */
int flnerf_composite_forward(flnerf_ctx *, int64_t B, int S, const float *raw, const float *z, const float *rays_d,
                             int64_t rays_d_stride, const float *noise, int white_bkgd, float *rgb, float *disp,
                             float *acc, float *depth, float *weights, void *stream);
/* g_* may be NULL (treated as zero). draw[B,S,4] is overwritten. */
int flnerf_composite_backward(flnerf_ctx *, int64_t B, int S, const float *raw, const float *z, const float *rays_d,
                              int64_t rays_d_stride, const float *noise, int white_bkgd, const float *g_rgb,
                              const float *g_disp, const float *g_acc, const float *g_depth, float *draw,
                              void *stream);

/* ---- a8: sample_pdf + sort-merge (run_nerf_helpers.py:112-155, render.py:279-284, 299).
 * u [B,Nf] or NULL; det!=0 -> u = linspace(0,1,Nf); NULL & det==0 -> Philox(seed, offset).
 * Outputs: z_merged[B,Nc+Nf] ascending, z_samples[B,Nf] (may be NULL), z_std[B] (may be NULL). */
int flnerf_sample_pdf_merge(flnerf_ctx *, int64_t B, int Nc, int Nf, const float *z, const float *weights,
                            const float *u, int det, uint64_t seed, uint64_t offset, float *z_merged,
                            float *z_samples, float *z_std, void *stream);

/* the plain sample_pdf(bins[B,n_bins], weights[B,n_bins-1], N_samples) API (run_nerf_helpers.py:112-155) */
int flnerf_sample_pdf(flnerf_ctx *, int64_t B, int n_bins, int Nf, const float *bins, const float *weights,
                      const float *u, int det, uint64_t seed, uint64_t offset, float *z_samples, void *stream);

/* ---- a9 + a13 accumulation: img2mse for fine and coarse (run_nerf_helpers.py:9, run_nerf.py:482-490), their
 * gradients, and the per-leaf max |gt - pred| table adjust_tree consumes (tree.py:538,642).
 * loss_out[2] = {mse(rgb), mse(rgb0)} (written, not accumulated); denom = number of rays the mean runs over
 * (global batch under data parallelism); rgb0/d_rgb0 may be NULL; leaf_gid (int32 [B], <0 = skip) and
 * leaf_max (fp32 table, atomicMax) may be NULL. */
int flnerf_mse_leafmax(flnerf_ctx *, int64_t B, const float *rgb, const float *rgb0, const float *target,
                       int64_t denom, const int32_t *leaf_gid, float *loss_out, float *d_rgb, float *d_rgb0,
                       float *leaf_max, void *stream);

/* ---- the MEAN refinement statistic of the nerf++ / plenoxels copies of the quadtree (nerf++-ours/tree.py:613-622): per leaf
 * leaf_sum[gid] += sum_c |target - pred| (double), leaf_cnt[gid] += 1 for every ray with leaf_gid >= 0; flnerf_leaf_mean turns
 * the two tables into leaf_stat[gid] = sum / (3 cnt) (-1 for a leaf without rays), which flnerf_qt_refine takes in place of the
 * per-leaf max. */
int flnerf_leaf_sum(flnerf_ctx *, int64_t B, const float *pred, const float *target, const int32_t *leaf_gid,
                    double *leaf_sum, int32_t *leaf_cnt, void *stream);
int flnerf_leaf_mean(flnerf_ctx *, int64_t n_slots, const double *leaf_sum, const int32_t *leaf_cnt, float *leaf_stat,
                     void *stream);

/* ---- torch.optim.Adam step (run_nerf.py:99,494) over flat buffers, t = 1-based step number; the bias
 * corrections are computed on the host in double like torch does. */
int flnerf_adam_step(flnerf_ctx *, int64_t n, float *param, float *m, float *v, const float *grad, double lr,
                     double b1, double b2, double eps, int64_t t, void *stream);

/* ---- a10-a13: GPU-resident quadtrees. Leaves of image i live at boxes[(i*cap + j)*4 .. +3] =
 * (x0,y0,x1,y1) doubles in DFS order (tree.py:61-72,679-686), count[i] leaves, min_area[i]. */
/* uniform tree of depth max_depth (tree.py:84-94 with mseThres=0) */
int flnerf_qt_init(flnerf_ctx *, int n_images, int cap, int H, int W, int max_depth, double *boxes, int32_t *count,
                   double *min_area, void *stream);
/* adjust_tree (tree.py:533-557,629-652): split leaves with leaf_max > thres and area == min_area; shrink min_area
 * of trees that split.  boxes_out must not alias boxes_in.  leaf_max is [n_images*cap]. */
int flnerf_qt_refine(flnerf_ctx *, int n_images, int cap, const double *boxes_in, const int32_t *count_in,
                     double *min_area, const float *leaf_max, float thres, double *boxes_out, int32_t *count_out,
                     void *stream);
/* gen_rays_v3_1 (tree.py:569-626) per-leaf ray counts: 10 if area > min_area+0.01 else int(area*rays_per_pixel);
 * ray_offset[n_images*cap+1] = exclusive scan over (image, leaf); total rays is ray_offset[n_images*cap]. */
int flnerf_qt_count(flnerf_ctx *, int n_images, int cap, const double *boxes, const int32_t *count,
                    const double *min_area, double rays_per_pixel, int64_t *ray_offset, void *stream);
/* emits the epoch's shuffled ray index buffer: for ray j of leaf (i,l): row ~ U[ceil(x0),ceil(x1)),
 * col ~ U[ceil(y0),ceil(y1-0.01)) (Philox(seed)), written at position perm(j) where perm is a keyed bijection of
 * [0,N) (replaces torch.randperm, tree.py:416).  ray_pix[N] = row*W+col, ray_gid[N] = i*cap+l. */
int flnerf_qt_emit(flnerf_ctx *, int n_images, int cap, int W, const double *boxes, const int32_t *count,
                   const int64_t *ray_offset, int64_t n_rays, uint64_t seed, int32_t *ray_pix, int32_t *ray_gid,
                   void *stream);
/* ---- probability-guided pixel sampling (prob=True; image_process.py:9-96, tree.py:583-595) ------------------------
 * sharp[n,H,W] = get_sharp_img(images[n,H,W,3]): 3x3 box-blur variance per channel -> sqrt|.| -> luminance
 * (image_process.py:26-39; cv2.blur border = BORDER_REFLECT_101, sums in double). */
int flnerf_sharp_map(flnerf_ctx *, int n_images, int H, int W, const float *images, float *sharp, void *stream);
/* row_offset[n_images*cap+1] = exclusive scan of the block heights int(x1)-int(x0) of every leaf (the slice
 * sharp[int(x0):int(x1), int(y0):int(y1)] of tree.py:587); total rows = row_offset[n_images*cap]. */
int flnerf_qt_prob_rows(flnerf_ctx *, int n_images, int cap, const double *boxes, const int32_t *count,
                        int64_t *row_offset, void *stream);
/* to_prob_v2 (image_process.py:59-74) per leaf: leaf_thr = 0.01*mean(gray+1e-6) and row_cdf[row_offset[s]+r] = inclusive
 * prefix over rows of sum_c max(gray+1e-6, leaf_thr) (float64).  row_cdf holds row_offset[n_images*cap] doubles. */
int flnerf_qt_prob_prepare(flnerf_ctx *, int n_images, int cap, int H, int W, const double *boxes,
                           const int32_t *count, const float *sharp, const int64_t *row_offset, double *row_cdf,
                           double *leaf_thr, void *stream);
/* gen_rays_v3_1 with prob=True: of a leaf's rays the first int(ray_num*(1-rand_frac)) follow np.random.choice(p) over
 * its block (inverse CDF, row-major, searchsorted side='right'), the rest the uniform rule of flnerf_qt_emit.
 * u [n_rays][2] fp32 (optional) replaces the Philox draws; shuffle=0 keeps emission order (parity tests). */
int flnerf_qt_emit_prob(flnerf_ctx *, int n_images, int cap, int H, int W, const double *boxes, const int32_t *count,
                        const int64_t *ray_offset, int64_t n_rays, uint64_t seed, double rand_frac, const float *sharp,
                        const int64_t *row_offset, const double *row_cdf, const double *leaf_thr, const float *u,
                        int shuffle, int32_t *ray_pix, int32_t *ray_gid, void *stream);
/* gathers a batch: for k in [0,B): ray = first + k*stride; target rgb from images (fp32 [n,H,W,3]) and the ray
 * (o,d) regenerated from pose/intrinsics (== get_rays at that pixel).  poses fp32 [n_images,12]. */
int flnerf_gather_batch(flnerf_ctx *, int64_t B, int64_t first, int64_t stride, const int32_t *ray_pix,
                        const int32_t *ray_gid, int cap, int H, int W, const double *h_K, const float *poses,
                        const float *images, float *rays_o, float *rays_d, float *target, int32_t *leaf_gid,
                        void *stream);

/* the same with the training images held as uint8 [n,H,W,3] (what the loaders read from disk: load_blender.py:37 divides by
 * 255 on the host) and decoded through lut256[u] = float32(u / 255.): 192 MB instead of 768 MB at lego size, the same targets */
int flnerf_gather_batch_u8(flnerf_ctx *, int64_t B, int64_t first, int64_t stride, const int32_t *ray_pix,
                           const int32_t *ray_gid, int cap, int H, int W, const double *h_K, const float *poses,
                           const uint8_t *images, const float *lut256, float *rays_o, float *rays_d, float *target,
                           int32_t *leaf_gid, void *stream);

/* ---- gen_rays_v3 (tree.py:231-307), the sub-pixel variant of the emitter: positions on a 1/1000 grid inside every leaf,
 * ray_xy[N,2] fp32 = (x = row coordinate, y = column coordinate) at the shuffled position; then colour / direction / origin by
 * F.grid_sample(bilinear, zeros, align_corners=False) semantics INCLUDING the reference's transposed grid (it passes (row, col)
 * where grid_sample expects (width, height)).  images: fp32 [n,H,W,3], or uint8 with lut256. */
int flnerf_qt_emit_sub(flnerf_ctx *, int n_images, int cap, const double *boxes, const int64_t *ray_offset, int64_t n_rays,
                       uint64_t seed, float *ray_xy, int32_t *ray_gid, void *stream);
int flnerf_gather_sub(flnerf_ctx *, int64_t B, const float *ray_xy, const int32_t *ray_gid, int cap, int H, int W,
                      const double *h_K, const float *poses, const void *images, const float *lut256_or_null, float *rays_o,
                      float *rays_d, float *target, void *stream);

/* ---- next row: nerf++-ours dual-MLP path -- parity-tested building blocks, no complete path yet.
 * Level-0 sample placement (ddp_train_nerf.py:54-81,352-366): fg_far[B] = unit-sphere exit depth, fg_z[B,N] linear in
 * [1e-4, fg_far], bg_z[B,N] = linspace(0,1) inverse depths; perturb != 0 jitters both inside their mid-point intervals with
 * the caller's uniforms t_fg/t_bg [B,N] (both or neither) or Philox(seed, offset). */
int flnerf_pp_depths0(flnerf_ctx *, int64_t B, int N, const float *rays_o, const float *rays_d, const float *t_fg,
                      const float *t_bg, int perturb, uint64_t seed, uint64_t offset, float *fg_far, float *fg_z, float *bg_z,
                      void *stream);
/* depth2pts_outside (ddp_model.py:16-45) + Embedder(4 -> 84) + Embedder(viewdir -> 27) in the FLIPPED sample order
 * NerfNet.forward feeds its background network (ddp_model.py:113-117): x111[B,N,111], bg_z_flip[B,N]; pts4[B,N,4]
 * (optional) = the un-flipped (unit direction, 1/r) points. */
int flnerf_pp_bg_encode(flnerf_ctx *, int64_t B, int N, const float *rays_o, const float *rays_d, const float *bg_z,
                        float *x111, float *bg_z_flip, float *pts4, void *stream);
/* NerfNet.forward compositing (ddp_model.py:93-133) from the two networks' raw outputs [.,4] = (rgb before the sigmoid,
 * sigma before |.|): rgb[B,3] = fg + bg_lambda*bg; fg_weights[B,Sf], bg_weights[B,Sb] (optional); aux9[B,9] (optional) =
 * fg_rgb(3), fg_depth, bg_rgb(3), bg_depth, bg_lambda.  bg_z_flip runs 1 -> 0 as produced by flnerf_pp_bg_encode. */
int flnerf_pp_composite_forward(flnerf_ctx *, int64_t B, int Sf, int Sb, const float *raw_fg, const float *fg_z,
                                const float *fg_far, const float *raw_bg, const float *bg_z_flip, const float *rays_d,
                                float *rgb, float *fg_weights, float *bg_weights, float *aux9, void *stream);
/* its backward for a loss on rgb: draw_fg[B,Sf,4], draw_bg[B,Sb,4] */
int flnerf_pp_composite_backward(flnerf_ctx *, int64_t B, int Sf, int Sb, const float *raw_fg, const float *fg_z,
                                 const float *fg_far, const float *raw_bg, const float *bg_z_flip, const float *rays_d,
                                 const float *g_rgb, float *draw_fg, float *draw_bg, void *stream);
/* level-1 sample placement (ddp_train_nerf.py:84-133,369-382): Nf depths resampled from weights[...,1:-1] over the
 * mid-point bins with the nerf++ sample_pdf (1e-6 floors, count-based inverse CDF), sort-merged with z[B,Nc]. */
int flnerf_pp_sample_pdf_merge(flnerf_ctx *, int64_t B, int Nc, int Nf, const float *z, const float *weights,
                               const float *u, int det, uint64_t seed, uint64_t offset, float *z_merged,
                               float *z_samples, void *stream);

/* ---- eval metrics (render.py:120-126): sums[0] = sum of the per-pixel, per-channel SSIM map of compute_ssim
 * (run_nerf_helpers.py:158-228: 11-tap Gaussian, sigma 1.5, zero padding, k1 0.01, k2 0.03), sums[1] = sum (img0 - img1)^2;
 * images [H,W,3] fp32; mean = sum / (3 H W), PSNR = -10 log10(sums[1] / (3 H W)). */
int flnerf_ssim_psnr(flnerf_ctx *, int H, int W, const float *img0, const float *img1, double max_val, double *sums,
                     void *stream);

/* ---- one CUDA graph per training step (run_nerf.py:470-516 is a fixed kernel sequence; only four scalars change from
 * one iteration to the next).  While a device-side step record is attached to the context, flnerf_gather_batch ADDS
 * rec->first to its `first`, flnerf_coarse_depths / flnerf_sample_pdf_merge (and the nerf++ pair flnerf_pp_depths0 /
 * flnerf_pp_sample_pdf_merge) ADD rec->rng_offset to their `offset`, and
 * flnerf_adam_step takes its step size and bias correction from the record instead of (lr, t): the captured launches carry
 * no per-step host value, and the host updates the record with ONE tiny kernel before each replay. */
typedef struct flnerf_step_record {
  int64_t first;        /* row of the epoch's ray index buffer where this step's batch starts */
  uint64_t rng_offset;  /* Philox counter offset of this step's uniforms */
  float adam_step_size; /* lr / (1 - b1^t) */
  float adam_bc2_sqrt;  /* sqrt(1 - b2^t) */
} flnerf_step_record;
int flnerf_set_step_record(flnerf_ctx *, const flnerf_step_record *rec /* device pointer, or NULL to detach */);
int flnerf_step_record_write(flnerf_ctx *, flnerf_step_record *rec, int64_t first, uint64_t rng_offset, double lr, double b1,
                             double b2, int64_t t, void *stream);

/* ---- diagnostics: number of kernels this library launched since the counter was last reset; _add accounts for
 * launches replayed from a captured graph (the capture itself counts once) */
int64_t flnerf_launch_count(int reset);
void flnerf_launch_count_add(int64_t n);

#ifdef __cplusplus
}
#endif
#endif
