// mlp_layout.h -- offsets (in floats) of the NeRF parameters inside the flat fp32 buffer, in the
// reference's parameters() order (model.py:20-34): pts_linears.0..7 {weight,bias}, views_linears.0,
// feature_linear, alpha_linear, rgb_linear.  D=8, W=256, input_ch=63, input_ch_views=27, skips=[4].
#pragma once

namespace mlp_layout {
// pts_linears.l: weight [256, K_l] with K_0 = 63, K_5 = 319 (skip after layer index 4), else 256
constexpr int K_PTS[8] = {63, 256, 256, 256, 256, 319, 256, 256};
constexpr int W_PTS[8] = {0, 16384, 82176, 147968, 213760, 279552, 361472, 427264};
constexpr int B_PTS[8] = {16128, 81920, 147712, 213504, 279296, 361216, 427008, 492800};
// device code cannot index a namespace-scope constexpr array with a runtime subscript
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int b_pts(int l) {
  return l == 0 ? 16128 : l == 1 ? 81920 : l == 2 ? 147712 : l == 3 ? 213504 : l == 4 ? 279296 : l == 5 ? 361216
         : l == 6 ? 427008 : 492800;
}
static_assert(b_pts(0) == B_PTS[0] && b_pts(3) == B_PTS[3] && b_pts(5) == B_PTS[5] && b_pts(7) == B_PTS[7], "layout");
constexpr int W_VIEWS = 493056;  // [128, 283] : columns 0..255 feature, 256..282 view-direction PE
constexpr int B_VIEWS = 529280;  // [128]
constexpr int W_FEAT = 529408;   // [256, 256]
constexpr int B_FEAT = 594944;   // [256]
constexpr int W_ALPHA = 595200;  // [1, 256]
constexpr int B_ALPHA = 595456;  // [1]
constexpr int W_RGB = 595457;    // [3, 128]
constexpr int B_RGB = 595841;    // [3]
constexpr int TOTAL = 595844;
static_assert(W_PTS[1] == W_PTS[0] + 256 * 63 + 256, "layout");
static_assert(W_PTS[6] == W_PTS[5] + 256 * 319 + 256, "layout");
static_assert(W_VIEWS == W_PTS[7] + 256 * 256 + 256, "layout");
static_assert(B_VIEWS == W_VIEWS + 128 * 283, "layout");
static_assert(W_FEAT == B_VIEWS + 128, "layout");
static_assert(W_ALPHA == B_FEAT + 256, "layout");
static_assert(W_RGB == B_ALPHA + 1, "layout");
static_assert(TOTAL == B_RGB + 3, "layout");
}  // namespace mlp_layout
