// api.cu -- context, error reporting and the mode dispatch of the MLP entry points of libflnerf.so.
#include "common.cuh"
#include <string.h>

long long g_flnerf_launches = 0;
static thread_local char g_err[512] = "";

void flnerf_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// mlp_simt.cu
size_t mlp_simt_stash_bytes(int64_t n);
size_t mlp_simt_bwd_workspace_bytes(int64_t n);
int mlp_simt_forward(const float *P, int64_t n, const float *x90, float *raw, float *stash, cudaStream_t st);
int mlp_simt_backward(const float *P, int64_t n, const float *x90, const float *stash, const float *draw, float *G,
                      float *ws, cudaStream_t st);
int64_t mlp_simt_param_count(int in_pts, int in_views);
int mlp_simt_forward_g(int in_pts, int in_views, const float *P, int64_t n, const float *x, float *raw, float *stash,
                       cudaStream_t st);
int mlp_simt_backward_g(int in_pts, int in_views, const float *P, int64_t n, const float *x, const float *stash,
                        const float *draw, float *G, float *ws, cudaStream_t st);
int mlp_simt_layout_selfcheck();
// mlp_tc.cu
size_t mlp_tc_packed_bytes(int kind);
int64_t mlp_tc_param_count(int kind);
size_t mlp_tc_stash_bytes(int64_t n, int S, int training, bool x3);
size_t mlp_tc_bwd_workspace_bytes(int64_t n, bool x3);
int mlp_tc_pack_weights(flnerf_ctx *ctx, int kind, const float *params, void *packed, cudaStream_t st);
int mlp_tc_forward(flnerf_ctx *ctx, bool x3, int kind, const float *params, const void *packed, int64_t n, int S, const void *pe_tiles,
                   const float *dirpe, float *raw, void *stash, int training, cudaStream_t st);
int mlp_tc_backward(flnerf_ctx *ctx, bool x3, int kind, const float *params, const void *packed, int64_t n, int S, const void *pe_tiles,
                    const float *dirpe, const void *stash, const float *draw, float *grads, void *ws, int stages,
                    cudaStream_t st);

extern "C" {

int flnerf_version(void) { return 100; }
const char *flnerf_last_error(void) { return g_err; }

flnerf_ctx *flnerf_create(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    flnerf_set_error("flnerf_create: no CUDA device %d (count %d): %s", device, count,
                     cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    flnerf_set_error("flnerf_create: cannot query device %d", device);
    return nullptr;
  }
  if (prop.major != 10) {
    flnerf_set_error("flnerf_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                     prop.major, prop.minor);
    return nullptr;
  }
  flnerf_ctx *c = new flnerf_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  return c;
}

void flnerf_destroy(flnerf_ctx *ctx) { delete ctx; }
int flnerf_sm_count(flnerf_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int64_t flnerf_launch_count(int reset) {
  long long v = g_flnerf_launches;
  if (reset) g_flnerf_launches = 0;
  return v;
}

void flnerf_launch_count_add(int64_t n) { g_flnerf_launches += n; }

size_t flnerf_mlp_packed_bytes(void) { return mlp_tc_packed_bytes(0); }

static int net_kind(int in_pts) { return in_pts == 63 ? 0 : (in_pts == 84 ? 1 : -1); }
size_t flnerf_mlp_packed_bytes_g(int in_pts) { return net_kind(in_pts) < 0 ? 0 : mlp_tc_packed_bytes(net_kind(in_pts)); }

size_t flnerf_mlp_stash_bytes(int mode, int64_t n, int S, int training) {
  if (mode == FLNERF_MODE_FP32) return mlp_simt_stash_bytes(n);
  return mlp_tc_stash_bytes(n, S, training, mode == FLNERF_MODE_BF16X3);
}

size_t flnerf_mlp_bwd_workspace_bytes(int mode, int64_t n) {
  if (mode == FLNERF_MODE_FP32) return mlp_simt_bwd_workspace_bytes(n);
  return mlp_tc_bwd_workspace_bytes(n, mode == FLNERF_MODE_BF16X3);
}

int flnerf_mlp_pack_weights(flnerf_ctx *ctx, const float *params, void *packed, void *stream) {
  FL_REQUIRE(ctx && params && packed, "flnerf_mlp_pack_weights: bad arguments");
  FL_CHECK_CUDA(cudaSetDevice(ctx->device));
  return mlp_tc_pack_weights(ctx, 0, params, packed, (cudaStream_t)stream);
}

int flnerf_mlp_pack_weights_g(flnerf_ctx *ctx, int in_pts, const float *params, void *packed, void *stream) {
  FL_REQUIRE(ctx && params && packed && net_kind(in_pts) >= 0, "flnerf_mlp_pack_weights_g: bad arguments (in_pts must be 63 or 84)");
  FL_CHECK_CUDA(cudaSetDevice(ctx->device));
  return mlp_tc_pack_weights(ctx, net_kind(in_pts), params, packed, (cudaStream_t)stream);
}

int flnerf_mlp_forward(flnerf_ctx *ctx, int mode, const float *params, const void *packed, int64_t n, int S,
                       const void *x, const float *dirpe, float *raw_out, void *stash, int training, void *stream) {
  FL_REQUIRE(ctx && params && x && raw_out && stash && n > 0 && S > 0, "flnerf_mlp_forward: bad arguments");
  FL_REQUIRE(((uintptr_t)raw_out & 15) == 0, "flnerf_mlp_forward: raw_out must be 16-byte aligned");
  if (mode == FLNERF_MODE_FP32) return mlp_simt_forward(params, n, (const float *)x, raw_out, (float *)stash, (cudaStream_t)stream);
  FL_REQUIRE(mode == FLNERF_MODE_BF16 || mode == FLNERF_MODE_BF16X3, "flnerf_mlp_forward: unknown mode %d", mode);
  FL_REQUIRE(packed && dirpe, "flnerf_mlp_forward: the tensor-core modes need packed weights and dirpe");
  FL_REQUIRE((((uintptr_t)packed | (uintptr_t)x | (uintptr_t)stash) & 1023) == 0,
             "flnerf_mlp_forward: packed / pe_tiles / stash must be 1024-byte aligned");
  return mlp_tc_forward(ctx, mode == FLNERF_MODE_BF16X3, 0, params, packed, n, S, x, dirpe, raw_out, stash, training,
                        (cudaStream_t)stream);
}

int flnerf_mlp_forward_g(flnerf_ctx *ctx, int mode, int in_pts, const float *params, const void *packed, int64_t n, int S,
                         const void *x, const float *dirpe, float *raw_out, void *stash, int training, void *stream) {
  FL_REQUIRE(ctx && params && x && raw_out && stash && packed && dirpe && n > 0 && S > 0 && net_kind(in_pts) >= 0,
             "flnerf_mlp_forward_g: bad arguments (in_pts must be 63 or 84)");
  FL_REQUIRE(mode == FLNERF_MODE_BF16 || mode == FLNERF_MODE_BF16X3, "flnerf_mlp_forward_g: tensor-core modes only (mode %d)", mode);
  FL_REQUIRE(((uintptr_t)raw_out & 15) == 0 && (((uintptr_t)packed | (uintptr_t)x | (uintptr_t)stash) & 1023) == 0,
             "flnerf_mlp_forward_g: raw_out must be 16-byte, packed / pe_tiles / stash 1024-byte aligned");
  return mlp_tc_forward(ctx, mode == FLNERF_MODE_BF16X3, net_kind(in_pts), params, packed, n, S, x, dirpe, raw_out, stash,
                        training, (cudaStream_t)stream);
}

int flnerf_mlp_backward_g(flnerf_ctx *ctx, int mode, int in_pts, const float *params, const void *packed, int64_t n, int S,
                          const void *x, const float *dirpe, const void *stash, const float *draw, float *grads, void *workspace,
                          size_t workspace_bytes, void *stream) {
  FL_REQUIRE(ctx && params && x && stash && draw && grads && workspace && packed && dirpe && n > 0 && S > 0 &&
                 net_kind(in_pts) >= 0,
             "flnerf_mlp_backward_g: bad arguments (in_pts must be 63 or 84)");
  FL_REQUIRE(mode == FLNERF_MODE_BF16 || mode == FLNERF_MODE_BF16X3, "flnerf_mlp_backward_g: tensor-core modes only (mode %d)", mode);
  FL_REQUIRE(workspace_bytes >= flnerf_mlp_bwd_workspace_bytes(mode, n), "flnerf_mlp_backward_g: workspace too small");
  FL_REQUIRE(((uintptr_t)draw & 15) == 0 && (((uintptr_t)packed | (uintptr_t)x | (uintptr_t)stash | (uintptr_t)workspace) & 1023) == 0,
             "flnerf_mlp_backward_g: draw must be 16-byte, packed / pe_tiles / stash / workspace 1024-byte aligned");
  return mlp_tc_backward(ctx, mode == FLNERF_MODE_BF16X3, net_kind(in_pts), params, packed, n, S, x, dirpe, stash, draw, grads,
                         workspace, 7, (cudaStream_t)stream);
}

int flnerf_mlp_backward_stages(flnerf_ctx *ctx, int mode, const float *params, const void *packed, int64_t n, int S,
                               const void *x, const float *dirpe, const void *stash, const float *draw, float *grads,
                               void *workspace, size_t workspace_bytes, int stages, void *stream);

int flnerf_mlp_backward(flnerf_ctx *ctx, int mode, const float *params, const void *packed, int64_t n, int S,
                        const void *x, const float *dirpe, const void *stash, const float *draw, float *grads,
                        void *workspace, size_t workspace_bytes, void *stream) {
  return flnerf_mlp_backward_stages(ctx, mode, params, packed, n, S, x, dirpe, stash, draw, grads, workspace,
                                    workspace_bytes, 7, stream);
}

int flnerf_mlp_backward_stages(flnerf_ctx *ctx, int mode, const float *params, const void *packed, int64_t n, int S,
                               const void *x, const float *dirpe, const void *stash, const float *draw, float *grads,
                               void *workspace, size_t workspace_bytes, int stages, void *stream) {
  FL_REQUIRE(ctx && params && x && stash && draw && grads && workspace && n > 0 && S > 0,
             "flnerf_mlp_backward: bad arguments");
  FL_REQUIRE(workspace_bytes >= flnerf_mlp_bwd_workspace_bytes(mode, n), "flnerf_mlp_backward: workspace too small");
  FL_REQUIRE(((uintptr_t)draw & 15) == 0, "flnerf_mlp_backward: draw must be 16-byte aligned");
  if (mode == FLNERF_MODE_FP32)
    return mlp_simt_backward(params, n, (const float *)x, (const float *)stash, draw, grads, (float *)workspace,
                             (cudaStream_t)stream);
  FL_REQUIRE(mode == FLNERF_MODE_BF16 || mode == FLNERF_MODE_BF16X3, "flnerf_mlp_backward: unknown mode %d", mode);
  FL_REQUIRE(packed && dirpe, "flnerf_mlp_backward: the tensor-core modes need packed weights and dirpe");
  FL_REQUIRE((((uintptr_t)packed | (uintptr_t)x | (uintptr_t)stash | (uintptr_t)workspace) & 1023) == 0,
             "flnerf_mlp_backward: packed / pe_tiles / stash / workspace must be 1024-byte aligned");
  return mlp_tc_backward(ctx, mode == FLNERF_MODE_BF16X3, 0, params, packed, n, S, x, dirpe, stash, draw, grads, workspace,
                         stages, (cudaStream_t)stream);
}

int64_t flnerf_mlp_param_count_g(int in_pts, int in_views) {
  if (in_pts < 1 || in_views < 1) return -1;
  return mlp_simt_param_count(in_pts, in_views);
}

int flnerf_mlp_fp32_forward_g(flnerf_ctx *ctx, int in_pts, int in_views, const float *params, int64_t n, const float *x,
                              float *raw_out, void *stash, void *stream) {
  FL_REQUIRE(ctx && params && x && raw_out && stash && n > 0 && in_pts > 0 && in_views > 0, "flnerf_mlp_fp32_forward_g: bad arguments");
  FL_REQUIRE(((uintptr_t)raw_out & 15) == 0, "flnerf_mlp_fp32_forward_g: raw_out must be 16-byte aligned");
  FL_REQUIRE(mlp_simt_layout_selfcheck() == 0, "flnerf: generic parameter layout disagrees with mlp_layout.h");
  return mlp_simt_forward_g(in_pts, in_views, params, n, x, raw_out, (float *)stash, (cudaStream_t)stream);
}

int flnerf_mlp_fp32_backward_g(flnerf_ctx *ctx, int in_pts, int in_views, const float *params, int64_t n, const float *x,
                               const void *stash, const float *draw, float *grads, void *workspace, size_t workspace_bytes,
                               void *stream) {
  FL_REQUIRE(ctx && params && x && stash && draw && grads && workspace && n > 0 && in_pts > 0 && in_views > 0,
             "flnerf_mlp_fp32_backward_g: bad arguments");
  FL_REQUIRE(workspace_bytes >= mlp_simt_bwd_workspace_bytes(n), "flnerf_mlp_fp32_backward_g: workspace too small");
  return mlp_simt_backward_g(in_pts, in_views, params, n, x, (const float *)stash, draw, grads, (float *)workspace,
                             (cudaStream_t)stream);
}

}  // extern "C"
