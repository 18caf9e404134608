// Layouts shared by the training kernels of the field network (forward with saved activations,
// backward dZ chain, dW contraction, gradient unpack / optimizer).
//
// Saved tensors are kept as *tile images*: for every 128-row tile, each [128 rows][64 columns]
// BF16 block is stored exactly as it sits in shared memory for tcgen05 (128-byte rows, 16-byte
// chunks XOR-swizzled by row & 7; see swz128_offset in mlp_layout.h).  The forward kernel dumps
// its activation tiles with one bulk copy per layer, the backward kernels load them back with
// bulk copies and point UMMA descriptors at them: K-major when the 64 columns are the contraction
// (dX = dZ . W), MN-major when the 128 rows are (dW = dZ^T . X).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "mlp_layout.h"

namespace nerf {

constexpr int kTileRows = 128;
constexpr size_t kBlockBytes = (size_t)kTileRows * 64 * 2;   // one [128][64] block = 16 KB

// ---- activation record of one tile (written by nerf_mlp_fwd with act_save) -------------------
constexpr size_t kActPE = 0;                                  // PE10(point), 1 block (column 63 = 0)
__host__ __device__ constexpr size_t act_hidden(int i) {      // i = 1..8: h_i (post-ReLU)
    return kBlockBytes + (size_t)(i - 1) * 4 * kBlockBytes;
}
constexpr size_t kActH10 = kBlockBytes + 8 * 4 * kBlockBytes; // h10 (post-ReLU), 2 blocks
// ReLU masks of h1..h8 and h10 as bits, for the dZ chain (which then never re-reads the BF16
// activations): [9 layers][128 rows][2 column halves] x 16 bytes.  The thread that owns (row, half)
// in the forward epilogue writes one uint4 per layer: word p covers its 16-column iterations 2p and
// 2p+1; the BF16 pair j (columns 2j, 2j+1 of iteration with parity b) puts its low element at bit
// 2j+b and its high element at bit 16+2j+b (relu_mask_bits below).  h10 (64 columns per thread)
// uses words 0 and 1 of its slot.
constexpr size_t kActMasks = kActH10 + 2 * kBlockBytes;
constexpr int kMaskLayers = 9;                                // index 0..7: h1..h8, 8: h10
constexpr size_t kActMaskBytes = (size_t)kMaskLayers * kTileRows * 2 * 16;   // 36864
constexpr size_t kActTileBytes = kActMasks + kActMaskBytes;   // 610304
__host__ __device__ constexpr size_t act_mask_slot(int layer, int row, int half) {
    return kActMasks + ((size_t)(layer * kTileRows + row) * 2 + half) * 16;
}
__host__ __device__ constexpr uint32_t relu_mask_bits(int parity, int j) {
    return (1u << (2 * j + parity)) | (1u << (16 + 2 * j + parity));
}

// ---- dZ record of one tile (written by nerf_mlp_bwd_dz) ---------------------------------------
// dZ_i = dL/d(pre-activation of layer i): i = 1..8 trunk, 10 = l10 (l9 is folded into l10: no dZ9).
__host__ __device__ constexpr size_t dz_hidden(int i) { return (size_t)(i - 1) * 4 * kBlockBytes; }
constexpr size_t kDz10 = 8 * 4 * kBlockBytes;                 // 2 blocks
constexpr size_t kDzTileBytes = kDz10 + 2 * kBlockBytes;      // 557056

// ---- transposed weights for the dZ chain (nerf_pack_model_bwd) --------------------------------
// Stage = [128 rows n = input feature][64 columns k = output feature] of W^T, K-major swizzled.
// Consumption order: (l10[:, :256] . l9)^T (the folded W', 2 K chunks), then l8, l7, l6[:, 63:],
// l5, l4, l3, l2 (4 K chunks each); every chunk has two 128-row halves, chunk-major / half-minor.
constexpr int kBwdLayers = 8;
__host__ __device__ constexpr int bwd_chunks(int j) { return j == 0 ? 2 : 4; }
__host__ __device__ constexpr int bwd_first_stage(int j) { return j == 0 ? 0 : 4 + (j - 1) * 8; }
constexpr int kBwdStages = 4 + 7 * 8;                         // 60
constexpr size_t kBwdWeightBytes = (size_t)kBwdStages * kStageBytes;
// fp32 tail: l_alpha.weight [256], l11.weight [3][128]
constexpr int kBwdTailWAlpha = 0;
constexpr int kBwdTailW11 = 256;
constexpr int kBwdTailFloats = 256 + 3 * 128;
constexpr size_t kBwdPackedBytes = kBwdWeightBytes + (size_t)kBwdTailFloats * 4;

// ---- gradient blob of one Model (fp32, padded so that every row is 16-byte aligned) -----------
// weight gradients are [out][pitch]; l6 keeps the PE columns (0..62, 63 = pad) in front of the
// h5 columns (64..319); l10 keeps the view-direction columns at 256..282.
constexpr int kG_W1 = 0;                                      // [256][64]
__host__ __device__ constexpr int grad_w_square(int l) {      // l = 2..5, 7..9 : [256][256]
    return 256 * 64 + (l <= 5 ? (l - 2) : (l - 3)) * 65536 + (l >= 7 ? 256 * 320 : 0);
}
constexpr int kG_W6 = 256 * 64 + 4 * 65536;                   // [256][320]
constexpr int kG_W10 = 256 * 64 + 7 * 65536 + 256 * 320;      // [128][288]
constexpr int kG_WAlpha = kG_W10 + 128 * 288;                 // [256]
constexpr int kG_W11 = kG_WAlpha + 256;                       // [3][128]
constexpr int kG_B = kG_W11 + 3 * 128;                        // [9][256] b1..b9
constexpr int kG_BAlpha = kG_B + 9 * 256;                     // [4]
constexpr int kG_B10 = kG_BAlpha + 4;                         // [128]
constexpr int kG_B11 = kG_B10 + 128;                          // [4]
// scratch behind the parameter gradients: G = dZ10^T . h8 [128][256], from which nerf_mlp_bwd_unfold
// forms the gradients of l9 and of l10's first 256 columns (the optimizer never reads this region)
constexpr int kG_Fold = kG_B11 + 4;                           // [128][256]
constexpr int kGradFloats = kG_Fold + 128 * 256;
static_assert(grad_w_square(2) == 256 * 64 && grad_w_square(5) + 65536 == kG_W6, "grad layout");
static_assert(grad_w_square(7) == kG_W6 + 256 * 320 && grad_w_square(9) + 65536 == kG_W10, "grad layout");

}  // namespace nerf
