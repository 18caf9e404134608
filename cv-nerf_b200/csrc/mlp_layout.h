// Packed-parameter blob layout of one field network (shared by pack, vterm and MLP kernels).
//
// The blob holds the BF16 weights as a sequence of 16 KB "stages".  One stage is the shared-memory
// image of a [128 output rows][64 input columns] K-major tile in the 128-byte-swizzled UMMA
// layout, so the producer warp can move it with one cp.async.bulk (TMA) and the MMA warp can point
// a tcgen05 shared-memory descriptor at it with no further shuffling.  Stages are stored in the
// exact order the MMA warp consumes them:
//
//   layer  reference tensor (model.py:57-71)   K chunks (64 wide)            N halves  stages
//   L1     l1.weight [256,63]                  1  (PE, col 63 zero)          2         2
//   L2-L5  l2..l5.weight [256,256]             4                             2         8 each
//   L6     l6.weight [256,319]                 5  (chunk 0 = cols 0..62 = PE, 2         10
//                                                  chunks 1..4 = cols 63..318 = h5)
//   L7,L8  l7,l8.weight                        4                             2         8 each
//   L10'   l10.weight[:, :256] . l9.weight     4                             1         4
//          [128,256] (folded, see below)
//                                                                             total     64
// Within a layer the order is chunk-major, half-minor.  The fp32 tail holds everything the
// epilogues consume on CUDA cores: biases, the sigma head (l_alpha), l11 and the view-direction
// columns of l10 (used by nerf_viewdir_term).
//
// l9 is folded into l10.  model.py:100-104 computes feat = l9(h8) WITHOUT an activation and feeds it
// (with the view encoding) straight into l10, and nothing else reads feat (sigma comes from h8), so
//     l10(cat[l9(h8), pe]) = (W10a W9) h8 + (W10a b9 + b10) + W10v pe,     W10a = l10.weight[:, :256]
// The pack kernels form W' = W10a W9 in FP32 and round it to BF16 once; the bias W10a b9 + b10 joins
// the per-ray view term.  One 256x256 contraction per sample (and its activation record, its dZ
// record and its dW job in training) disappears: 9 tensor-core layers instead of 10.  The gradients of
// l9 and of l10's first 256 columns follow from G = dZ10^T h8 by the chain rule (nerf_mlp_bwd_unfold).
// Against the two-step form this drops one BF16 rounding of an activation (feat) per sample.
#pragma once
#include <stdint.h>

namespace nerf {

constexpr int kStageRows = 128;
constexpr int kStageCols = 64;
constexpr int kStageBytes = kStageRows * kStageCols * 2;  // 16384
constexpr int kNumMmaLayers = 9;                          // l1..l8, l10' (l9 folded into l10)
constexpr int kNumStages = 64;
constexpr int kHidden = 256;
constexpr int kPeDim = 63;
constexpr int kViewPeDim = 27;
constexpr int kL10Out = 128;

// per MMA layer: number of 64-wide K chunks and number of 128-row N halves
__host__ __device__ constexpr int layer_chunks(int l) { return l == 0 ? 1 : (l == 5 ? 5 : 4); }
__host__ __device__ constexpr int layer_halves(int l) { return l == 8 ? 1 : 2; }
__host__ __device__ constexpr int layer_first_stage(int l) {
    int s = 0;
    for (int i = 0; i < l; ++i) s += layer_chunks(i) * layer_halves(i);
    return s;
}
static_assert(layer_first_stage(kNumMmaLayers) == kNumStages, "stage count");

// fp32 tail (offsets in floats from the start of the tail)
constexpr int kTailBias = 0;                                   // [9][256]  b1..b8 (row 8 unused: b9 is folded into kTailB10)
constexpr int kTailWAlpha = kTailBias + 9 * kHidden;           // [256]     l_alpha.weight
constexpr int kTailBAlpha = kTailWAlpha + kHidden;             // [4]       l_alpha.bias (+pad)
constexpr int kTailW11 = kTailBAlpha + 4;                      // [3][128]  l11.weight
constexpr int kTailB11 = kTailW11 + 3 * kL10Out;               // [4]       l11.bias (+pad)
constexpr int kTailW10View = kTailB11 + 4;                     // [128][28] l10.weight[:, 256:283] (+pad)
constexpr int kTailB10 = kTailW10View + kL10Out * 28;          // [128]     l10.bias + l10.weight[:, :256] . l9.bias
constexpr int kTailFloats = kTailB10 + kL10Out;

constexpr size_t kWeightBytes = (size_t)kNumStages * kStageBytes;
constexpr size_t kPackedBytes = kWeightBytes + (size_t)kTailFloats * 4;

// byte offset of element (row r, col k) inside one swizzled stage image
__host__ __device__ constexpr uint32_t swz128_offset(int r, int k) {
    return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1));
}

}  // namespace nerf
