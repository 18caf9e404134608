// Backward of the field network, part 2: weight and bias gradients.
//
//   dW_i[out][in] = sum over samples  dZ_i[s][out] * X_i[s][in]        (X_i = input of layer i)
//   db_i[out]     = sum over samples  dZ_i[s][out]
// for the tensor-core layers l1..l10 (what autograd accumulates into .grad for
// /root/reference/model.py:57-71 during loss.backward(), main.py:385).  The contraction runs over
// the SAMPLE axis, so both operands are read "sideways" from the tile images the forward pass
// (activations) and the dZ chain (gradients) left in HBM: UMMA MN-major descriptors over the very
// same 128-byte-swizzled blocks, no transposes anywhere.
//
// One CTA = one job = (layer, contiguous range of 128-sample tiles).  Jobs are sized so that all
// CTAs stream about the same number of bytes (the kernel is HBM-bound: 128 KB per tile and layer
// against 4.2 MFLOP/KB) and the grid never exceeds the SM count (one CTA per SM, one wave).
//   warp 0      producer: 32-sample K slices of the dZ and X blocks -> 6 x 32 KB ring (bulk copies)
//   warp 1      MMA issuer: tcgen05.mma M=128 (x2 for 256 outputs), N = 64 or 256, K=16,
//               FP32 accumulators for the whole [out][in] tile stay in tensor memory
//   warps 2-5   bias sums: column sums of the dZ slices straight from the ring
//   warps 2-9   at the end: accumulators -> red.global.add.v4.f32 into the gradient blob
#include <cuda_bf16.h>

#include "common.cuh"
#include "mlp_bwd_layout.h"
#include "umma.cuh"

namespace {

using namespace nerf;

constexpr int kRing = 6;
constexpr int kSliceRows = 32;                           // samples per ring slot
constexpr uint32_t kPieceBytes = kSliceRows * 128;       // one block's share of a slot: 4 KB
constexpr uint32_t kSlotBytes = 8 * kPieceBytes;         // up to 4 dZ + 4 X blocks: 32 KB
constexpr int kSlicesPerTile = kTileRows / kSliceRows;   // 4
constexpr int kBiasWarps = 4;
constexpr int kDrainWarps = 8;
constexpr int kThreads = 64 + kDrainWarps * 32;          // 320
constexpr uint32_t kOffBar = kRing * kSlotBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256;
constexpr int kNumJobsLayers = 10;     // l1, l2..l5, l6 (PE), l6 (h5), l7, l8, l10' (l9 folded: G = dZ10^T h8)

struct LayerJob {
    uint32_t a_off;      // byte offset of the dZ image inside a dZ tile record
    uint32_t b_off;      // byte offset of the X image inside an activation tile record
    int a_blocks;        // 64-column blocks of dZ: 4 (256 outputs) or 2 (128)
    int b_blocks;        // 64-column blocks of X: 4 (256 inputs) or 1 (the positional encoding)
    int out_off;         // float offset of dW[0][0] in the gradient blob
    int out_pitch;       // floats per dW row
    int bias_off;        // float offset of db in the gradient blob, or -1
    int first_cta;       // CTAs [first_cta, first_cta + parts) work on this layer
    int parts;
};

struct DwParams {
    LayerJob jobs[kNumJobsLayers];
    const uint8_t* act;
    const uint8_t* dz;
    float* grad;
    long n_tiles;
    float* partial;      // deterministic mode: per-CTA partial sums [grid][kPartFloats] instead of atomics into grad
};

// one CTA's partial result in deterministic mode: dW rows [256][256] (job-local row-major, pitch 256) + db [256]
constexpr int kPartFloats = 256 * 256 + 256;

__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_dw_kernel(const __grid_constant__ DwParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    if ((sbase & 1023u) != 0) __trap();
    const uint32_t bar_full = sbase + kOffBar;               // [kRing]
    const uint32_t bar_empty = bar_full + 8 * kRing;         // [kRing]
    const uint32_t bar_done = bar_empty + 8 * kRing;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (2 * kRing + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // which job?
    int ji = 0;
#pragma unroll 1
    for (int k = 0; k < kNumJobsLayers; ++k)
        if ((int)blockIdx.x >= P.jobs[k].first_cta) ji = k;
    const LayerJob& J = P.jobs[ji];
    const int part = (int)blockIdx.x - J.first_cta;
    const long t_begin = P.n_tiles * part / J.parts;
    const long t_end = P.n_tiles * (part + 1) / J.parts;
    const bool with_bias = J.bias_off >= 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kRing; ++s) {
            umma::mbar_init(bar_full + 8 * s, 1);
            umma::mbar_init(bar_empty + 8 * s, 1 + (with_bias ? kBiasWarps : 0));
        }
        umma::mbar_init(bar_done, 1);
        umma::fence_barrier_init();
    }
    if (warp == 1) {
        umma::tmem_alloc(umma::smem_u32(tmem_slot), 512);
        umma::tmem_relinquish();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_cols = J.b_blocks * 64;

    if (warp == 0) {
        // ===================== producer =====================
        // The up to eight 4 KB pieces of a slot are issued by eight LANES in one warp instruction: bulk
        // copies issued one after the other by a single thread do not overlap (16 KB copies: 30 B/cycle
        // from one thread, 60 from two, tools/probes/sw64_mma_probe.cu), and eight serial 4 KB copies per
        // 32 KB slot capped this HBM-bound kernel at the issue rate of one thread.
        {
            uint32_t it = 0;
            const uint32_t bytes = (uint32_t)(J.a_blocks + J.b_blocks) * kPieceBytes;
            const bool mine_a = lane < J.a_blocks;
            const bool mine_b = lane >= 4 && lane - 4 < J.b_blocks;
            for (long t = t_begin; t < t_end; ++t) {
                const uint8_t* a_src = P.dz + (size_t)t * kDzTileBytes + J.a_off;
                const uint8_t* b_src = P.act + (size_t)t * kActTileBytes + J.b_off;
                for (int ks = 0; ks < kSlicesPerTile; ++ks, ++it) {
                    const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
                    umma::mbar_wait_warp(bar_empty + 8 * slot, ph ^ 1);
                    if (lane == 0) umma::mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
                    __syncwarp();
                    const uint32_t dst = sbase + slot * kSlotBytes;
                    if (mine_a)
                        umma::bulk_g2s(dst + lane * kPieceBytes, a_src + (size_t)lane * kBlockBytes + ks * kPieceBytes,
                                       kPieceBytes, bar_full + 8 * slot);
                    else if (mine_b)
                        umma::bulk_g2s(dst + lane * kPieceBytes, b_src + (size_t)(lane - 4) * kBlockBytes + ks * kPieceBytes,
                                       kPieceBytes, bar_full + 8 * slot);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma::instr_desc_bf16_mn(128, n_cols);
            const int m_blocks = J.a_blocks / 2;
            uint32_t it = 0;
            for (long t = t_begin; t < t_end; ++t) {
                for (int ks = 0; ks < kSlicesPerTile; ++ks, ++it) {
                    const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
                    umma::mbar_wait(bar_full + 8 * slot, ph);
                    umma::tc_fence_after();
                    const uint32_t a_base = sbase + slot * kSlotBytes;
                    const uint32_t b_base = a_base + 4 * kPieceBytes;
#pragma unroll
                    for (int kk = 0; kk < kSliceRows / 16; ++kk) {
                        const uint64_t b_desc = umma::smem_desc_sw128_mn(b_base + kk * 2048, kPieceBytes);
                        for (int mb = 0; mb < m_blocks; ++mb) {
                            umma::mma_bf16_ss(tmem_base + mb * 256,
                                              umma::smem_desc_sw128_mn(a_base + mb * 2 * kPieceBytes + kk * 2048, kPieceBytes),
                                              b_desc, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                        }
                    }
                    umma::mma_commit(bar_empty + 8 * slot);
                }
            }
            umma::mma_commit(bar_done);
        }
    } else {
        const int dw = warp - 2;                         // 0..7
        // ===================== bias sums (warps 2-5) =====================
        float s0 = 0.f, s1 = 0.f;
        if (with_bias && dw < kBiasWarps) {
            const int f = (dw * 32 + lane) * 2;          // features f, f+1 of the 256 outputs
            const uint32_t col = (uint32_t)(f >> 6) * kPieceBytes + (uint32_t)(f & 7) * 2;
            const uint32_t c16 = (uint32_t)(f & 63) >> 3;
            uint32_t it = 0;
            for (long t = t_begin; t < t_end; ++t) {
                for (int ks = 0; ks < kSlicesPerTile; ++ks, ++it) {
                    const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
                    umma::mbar_wait(bar_full + 8 * slot, ph);
                    const uint32_t base = sbase + slot * kSlotBytes + col;
#pragma unroll 8
                    for (int r = 0; r < kSliceRows; ++r) {
                        const uint32_t w = umma::ld_shared_u32(base + r * 128 + ((c16 ^ (uint32_t)(r & 7)) << 4));
                        s0 += __uint_as_float(w << 16);
                        s1 += __uint_as_float(w & 0xffff0000u);
                    }
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(bar_empty + 8 * slot);
                }
            }
            if (t_end > t_begin) {
                if (P.partial) {
                    float* pb = P.partial + (size_t)blockIdx.x * kPartFloats + 256 * 256;
                    pb[f] = s0;
                    pb[f + 1] = s1;
                } else {
                    atomicAdd(P.grad + J.bias_off + f, s0);
                    atomicAdd(P.grad + J.bias_off + f + 1, s1);
                }
            }
        }
        // ===================== drain: accumulators -> gradient blob =====================
        if (t_end > t_begin) {
            umma::mbar_wait(bar_done, 0);
            umma::tc_fence_after();
            const int quad = warp & 3;
            const int m_blocks = J.a_blocks / 2;
            // 8 warps: two per TMEM lane quadrant; with two M blocks each takes one, otherwise they
            // split the columns
            const int second = dw >> 2;
            const int mb = m_blocks == 2 ? second : 0;
            const int c_begin = m_blocks == 2 ? 0 : second * (n_cols / 2);
            const int c_end = m_blocks == 2 ? n_cols : c_begin + n_cols / 2;
            const int out_row = mb * 128 + quad * 32 + lane;
            float* dst = P.grad + J.out_off + (size_t)out_row * J.out_pitch;
            float* dst_det = P.partial ? P.partial + (size_t)blockIdx.x * kPartFloats + (size_t)out_row * 256 : nullptr;
            const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + mb * 256;
            for (int c = c_begin; c < c_end; c += 16) {
                uint32_t v[16];
                umma::tmem_ld16(tacc + c, v);
                umma::tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (dst_det)
                        *reinterpret_cast<float4*>(dst_det + c + q * 4) =
                            make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                        __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                    else
                        umma::red_add_v4(dst + c + q * 4, __uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                         __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                }
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, 512);
    }
}

// Deterministic mode, second launch: every blob element adds its CTAs' partial sums in CTA order.  Four
// consecutive columns per thread (16-byte loads; every partial of an element is an independent load, only the
// additions are ordered): the launch is latency-bound, 56 us with one scalar element per thread.
__global__ void __launch_bounds__(256) dw_reduce_kernel(const __grid_constant__ DwParams P) {
    const LayerJob& J = P.jobs[blockIdx.y];
    const int rows = J.a_blocks * 64, cols = J.b_blocks * 64;
    const int n_w4 = rows * cols / 4, n_items = n_w4 + (J.bias_off >= 0 ? rows : 0);
    for (int item = blockIdx.x * 256 + threadIdx.x; item < n_items; item += gridDim.x * 256) {
        const bool is_w = item < n_w4;
        const int r = is_w ? (item * 4) / cols : item - n_w4, c = is_w ? (item * 4) % cols : 0;
        const size_t local = is_w ? (size_t)r * 256 + c : (size_t)256 * 256 + r;
        const float* src = P.partial + (size_t)J.first_cta * kPartFloats + local;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int part = 0; part < J.parts; ++part) {
            if (P.n_tiles * (part + 1) / J.parts <= P.n_tiles * part / J.parts) continue;     // this CTA had no tiles
            if (is_w) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)part * kPartFloats));
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            } else {
                sum.x += __ldg(src + (size_t)part * kPartFloats);
            }
        }
        if (is_w) {
            float4* dst = reinterpret_cast<float4*>(P.grad + J.out_off + (size_t)r * J.out_pitch + c);
            float4 d = *dst;
            d.x += sum.x; d.y += sum.y; d.z += sum.z; d.w += sum.w;
            *dst = d;
        } else {
            P.grad[J.bias_off + r] += sum.x;
        }
    }
}

// bytes one tile of layer job k moves (balances the split of the grid)
int job_weight(const LayerJob& j) { return (j.a_blocks + j.b_blocks) * 16; }

}  // namespace

extern "C" size_t nerf_grad_blob_bytes(void) { return (size_t)nerf::kGradFloats * 4; }

// dW / db of l1..l10 (tensor-core layers) accumulated into grad_blob (+=).
static int launch_dw(const void* act_save, const void* dz, long M, float* grad_blob, float* partial, void* stream);

extern "C" int nerf_mlp_bwd_dw(const void* act_save, const void* dz, long M, float* grad_blob, void* stream) {
    return launch_dw(act_save, dz, M, grad_blob, nullptr, stream);
}

// Deterministic accumulation: per-CTA partial sums into `scratch`, then an ordered reduction (second launch).
extern "C" int nerf_mlp_bwd_dw_det(const void* act_save, const void* dz, long M, float* grad_blob, void* scratch,
                                   void* stream) {
    if (!scratch) return nerf::arg_error("nerf_mlp_bwd_dw_det: scratch");
    return launch_dw(act_save, dz, M, grad_blob, (float*)scratch, stream);
}

extern "C" size_t nerf_mlp_bwd_dw_det_scratch_bytes(void) {
    int sms = nerf_b200_sm_count();
    if (sms <= 0) sms = 148;
    return (size_t)sms * kPartFloats * 4;
}

static int launch_dw(const void* act_save, const void* dz, long M, float* grad_blob, float* partial, void* stream) {
    nerf::DeviceGuard device_guard(grad_blob);
    if (M < 0 || (M > 0 && (!act_save || !dz || !grad_blob))) return nerf::arg_error("nerf_mlp_bwd_dw");
    if (M == 0) return 0;
    static int sm_count = 0;
    static bool configured = false;
    if (sm_count == 0) {
        sm_count = nerf_b200_sm_count();
        if (sm_count <= 0) {
            sm_count = 0;
            nerf::set_last_error("nerf_mlp_bwd_dw: no CUDA device");
            return (int)cudaErrorNoDevice;
        }
    }
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            nerf::set_last_error("nerf_mlp_bwd_dw setup: %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    DwParams P;
    P.act = (const uint8_t*)act_save; P.dz = (const uint8_t*)dz; P.grad = grad_blob; P.partial = partial;
    P.n_tiles = (M + kTileRows - 1) / kTileRows;
    // layer table: l1, l2..l5, l6 (PE columns), l6 (h5 columns), l7, l8, and the folded l10': G = dZ10^T . h8
    // into the scratch region of the blob (nerf_mlp_bwd_unfold turns it into the gradients of l9 and l10)
    int k = 0;
    auto add = [&](size_t a_off, int a_blocks, size_t b_off, int b_blocks, int out_off, int pitch, int bias_off) {
        LayerJob& j = P.jobs[k++];
        j.a_off = (uint32_t)a_off; j.a_blocks = a_blocks; j.b_off = (uint32_t)b_off; j.b_blocks = b_blocks;
        j.out_off = out_off; j.out_pitch = pitch; j.bias_off = bias_off; j.first_cta = 0; j.parts = 1;
    };
    add(dz_hidden(1), 4, kActPE, 1, kG_W1, 64, kG_B + 0 * 256);
    for (int l = 2; l <= 5; ++l) add(dz_hidden(l), 4, act_hidden(l - 1), 4, grad_w_square(l), 256, kG_B + (l - 1) * 256);
    add(dz_hidden(6), 4, kActPE, 1, kG_W6, 320, -1);
    add(dz_hidden(6), 4, act_hidden(5), 4, kG_W6 + 64, 320, kG_B + 5 * 256);
    for (int l = 7; l <= 8; ++l) add(dz_hidden(l), 4, act_hidden(l - 1), 4, grad_w_square(l), 256, kG_B + (l - 1) * 256);
    add(kDz10, 2, act_hidden(8), 4, kG_Fold, 256, -1);
    // split the SMs over the layers in proportion to their traffic, at most one CTA per tile
    const long cap = P.n_tiles;
    int total = kNumJobsLayers;
    while (total < sm_count) {
        int best = -1;
        double best_load = 0.;
        for (int i = 0; i < kNumJobsLayers; ++i) {
            if (P.jobs[i].parts >= cap) continue;
            double load = (double)job_weight(P.jobs[i]) / P.jobs[i].parts;
            if (load > best_load) { best_load = load; best = i; }
        }
        if (best < 0) break;
        ++P.jobs[best].parts;
        ++total;
    }
    int first = 0;
    for (int i = 0; i < kNumJobsLayers; ++i) {
        P.jobs[i].first_cta = first;
        first += P.jobs[i].parts;
    }
    mlp_bwd_dw_kernel<<<(unsigned)first, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    if (partial) dw_reduce_kernel<<<dim3(64, kNumJobsLayers), 256, 0, (cudaStream_t)stream>>>(P);
    return nerf::check_launch("nerf_mlp_bwd_dw");
}
