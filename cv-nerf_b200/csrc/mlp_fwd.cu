// K2: fused field-network forward for sm_100a.
//
// Replaces net_forward + FreqEmbedding.embed + Model.forward
// (/root/reference/model.py:9-31, 77-131) and the point construction of render_rays
// (/root/reference/main.py:238,252): sample position -> positional encoding -> 8x256 trunk with
// the skip at l6 -> sigma head, l9, l10 (+ hoisted view term), l11 -> raw[rgb(3), sigma].
// Per-sample activations never leave the SM: BF16 activations live in shared memory, FP32
// accumulators in tensor memory.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      producer: streams 16 KB weight stages L2 -> smem with cp.async.bulk (TMA engine)
//               through a 4-deep mbarrier ring, in the order mlp_layout.h stores them;
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=128, K=16, BF16 -> FP32 in
//               TMEM) and tcgen05.commit; owns the 512-column TMEM allocation;
//   warps 2-5   epilogue group X, warps 6-9 epilogue group Y: each group owns one 128-sample
//               sub-tile (thread = sample row = TMEM lane).  It computes the positional encoding,
//               and after every layer reads the accumulator (tcgen05.ld), adds bias, applies ReLU,
//               rounds to BF16 and writes the next layer's A operand back to shared memory in the
//               UMMA K-major SWIZZLE_128B layout.  sigma (l_alpha) and rgb (l11) are evaluated in
//               FP32 on CUDA cores from the FP32 accumulators of l8 / l10.
// The two sub-tiles ping-pong: while the tensor core runs layer l of Y, group X runs the epilogue
// of layer l of X, so MMA and epilogue overlap.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mlp_layout.h"
#include "umma.cuh"

namespace {

using namespace nerf;

constexpr int kTileM = 128;
constexpr int kRing = 4;
constexpr int kThreads = 320;
constexpr uint32_t kOffA = 0;                         // 2 x [4][128][64] bf16
constexpr uint32_t kOffPE = 2 * 65536;                // 2 x [128][64] bf16
constexpr uint32_t kOffW = kOffPE + 2 * 16384;        // kRing x 16 KB
constexpr uint32_t kOffBar = kOffW + kRing * kStageBytes;
constexpr uint32_t kSmemBytes = kOffBar + 128 + 1024;  // + barriers + alignment slack
constexpr uint32_t kIdesc = umma::instr_desc_bf16(128, 128);

struct FwdParams {
    const uint8_t* blob;     // packed model
    int in_mode;
    const float* in0;
    const float* in1;
    int in_stride;
    long M;
    int S;
    const float* vterm;
    int vterm_div;
    float* raw_out;
    float* probe_out;        // debug: [M][256] post-activation of layer probe_layer (or NULL)
    int probe_layer;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------- input stage
// Positional encoding of one point, model.py:15-31: [x, sin(2^k x), cos(2^k x)]_{k<10}, 63 wide,
// written as one 128-byte swizzled row (column 63 = 0).  sin/cos(2^k x) are produced from two
// accurate sincosf anchors (k = 0 and k = 5) by angle doubling; the accumulated error (<1e-5) is
// far below the BF16 rounding applied next.
__device__ __forceinline__ void encode_point(float px, float py, float pz, float (&f)[64]) {
    const float p[3] = {px, py, pz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        f[a] = p[a];
        float s, c;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            if (k == 0) sincosf(p[a], &s, &c);
            else if (k == 5) sincosf(p[a] * 32.f, &s, &c);
            else {
                float s2 = 2.f * s * c;
                float c2 = fmaf(-2.f * s, s, 1.f);
                s = s2; c = c2;
            }
            f[3 + 6 * k + a] = s;
            f[3 + 6 * k + 3 + a] = c;
        }
    }
    f[63] = 0.f;
}

__device__ __forceinline__ void store_row_bf16(uint8_t* tile, int row, const float (&f)[64]) {
#pragma unroll
    for (int c16 = 0; c16 < 8; ++c16) {
        uint4 q;
        q.x = pack_bf16x2(f[c16 * 8 + 0], f[c16 * 8 + 1]);
        q.y = pack_bf16x2(f[c16 * 8 + 2], f[c16 * 8 + 3]);
        q.z = pack_bf16x2(f[c16 * 8 + 4], f[c16 * 8 + 5]);
        q.w = pack_bf16x2(f[c16 * 8 + 6], f[c16 * 8 + 7]);
        *reinterpret_cast<uint4*>(tile + row * 128 + ((c16 ^ (row & 7)) << 4)) = q;
    }
}

__device__ __forceinline__ void input_stage(const FwdParams& P, long grow, uint8_t* pe_tile, int row) {
    float f[64];
    if (P.in_mode == NERF_IN_EMBEDDED) {
        const float* x = P.in0 + grow * P.in_stride;
#pragma unroll
        for (int c = 0; c < 63; ++c) f[c] = __ldg(x + c);
        f[63] = 0.f;
    } else {
        float px, py, pz;
        if (P.in_mode == NERF_IN_RAYS) {
            const float* r = P.in0 + (grow / P.S) * NERF_RAY_STRIDE;
            float z = __ldg(P.in1 + grow);
            // main.py:238: o + d * z, multiply and add rounded separately
            px = __fadd_rn(__ldg(r + 0), __fmul_rn(__ldg(r + 3), z));
            py = __fadd_rn(__ldg(r + 1), __fmul_rn(__ldg(r + 4), z));
            pz = __fadd_rn(__ldg(r + 2), __fmul_rn(__ldg(r + 5), z));
        } else {
            const float* x = P.in0 + grow * 3;
            px = __ldg(x); py = __ldg(x + 1); pz = __ldg(x + 2);
        }
        encode_point(px, py, pz, f);
    }
    store_row_bf16(pe_tile, row, f);
}

// ---------------------------------------------------------------------------- epilogues
// Hidden layer: h = act(acc + bias) -> BF16 -> A tile (in place).  MODE 0: ReLU; 1: ReLU and
// accumulate the FP32 sigma head (l_alpha) ; 2: no activation (l9).
template <int MODE, bool PROBE>
__device__ __forceinline__ void epilogue_hidden(uint32_t tacc, uint8_t* a_tile, int row,
                                                const float* __restrict__ bias,
                                                const float* __restrict__ walpha, float& sigma,
                                                float* probe_row) {
    uint32_t v[2][32];
    umma::tmem_ld32(tacc, v[0]);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        umma::tmem_wait_ld();
        if (it + 1 < 8) umma::tmem_ld32(tacc + (it + 1) * 32, v[(it + 1) & 1]);
        const uint32_t(&cur)[32] = v[it & 1];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = it * 32 + q * 8;
            float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
            float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
            float h[8];
            h[0] = __uint_as_float(cur[q * 8 + 0]) + b0.x;
            h[1] = __uint_as_float(cur[q * 8 + 1]) + b0.y;
            h[2] = __uint_as_float(cur[q * 8 + 2]) + b0.z;
            h[3] = __uint_as_float(cur[q * 8 + 3]) + b0.w;
            h[4] = __uint_as_float(cur[q * 8 + 4]) + b1.x;
            h[5] = __uint_as_float(cur[q * 8 + 5]) + b1.y;
            h[6] = __uint_as_float(cur[q * 8 + 6]) + b1.z;
            h[7] = __uint_as_float(cur[q * 8 + 7]) + b1.w;
            if (MODE != 2) {
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = fmaxf(h[e], 0.f);
            }
            if (MODE == 1) {
                float4 w0 = __ldg(reinterpret_cast<const float4*>(walpha + c));
                float4 w1 = __ldg(reinterpret_cast<const float4*>(walpha + c + 4));
                sigma = fmaf(w0.x, h[0], sigma); sigma = fmaf(w0.y, h[1], sigma);
                sigma = fmaf(w0.z, h[2], sigma); sigma = fmaf(w0.w, h[3], sigma);
                sigma = fmaf(w1.x, h[4], sigma); sigma = fmaf(w1.y, h[5], sigma);
                sigma = fmaf(w1.z, h[6], sigma); sigma = fmaf(w1.w, h[7], sigma);
            }
            if (PROBE && probe_row) {
#pragma unroll
                for (int e = 0; e < 8; ++e) probe_row[c + e] = h[e];
            }
            uint4 o;
            o.x = pack_bf16x2(h[0], h[1]);
            o.y = pack_bf16x2(h[2], h[3]);
            o.z = pack_bf16x2(h[4], h[5]);
            o.w = pack_bf16x2(h[6], h[7]);
            const int blk = it >> 1, c16 = (it & 1) * 4 + q;
            *reinterpret_cast<uint4*>(a_tile + blk * 16384 + row * 128 + ((c16 ^ (row & 7)) << 4)) = o;
        }
    }
}

// l10 (+ hoisted view term, ReLU) and l11 in FP32: returns rgb_raw.
template <bool PROBE>
__device__ __forceinline__ void epilogue_rgb(uint32_t tacc, const float* __restrict__ vt,
                                             const float* __restrict__ w11, float (&rgb)[3],
                                             float* probe_row) {
    uint32_t v[2][32];
    umma::tmem_ld32(tacc, v[0]);
    rgb[0] = rgb[1] = rgb[2] = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        umma::tmem_wait_ld();
        if (it + 1 < 4) umma::tmem_ld32(tacc + (it + 1) * 32, v[(it + 1) & 1]);
        const uint32_t(&cur)[32] = v[it & 1];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = it * 32 + q * 4;
            float4 t = __ldg(reinterpret_cast<const float4*>(vt + c));
            float h0 = fmaxf(__uint_as_float(cur[q * 4 + 0]) + t.x, 0.f);
            float h1 = fmaxf(__uint_as_float(cur[q * 4 + 1]) + t.y, 0.f);
            float h2 = fmaxf(__uint_as_float(cur[q * 4 + 2]) + t.z, 0.f);
            float h3 = fmaxf(__uint_as_float(cur[q * 4 + 3]) + t.w, 0.f);
            if (PROBE && probe_row) {
                probe_row[c] = h0; probe_row[c + 1] = h1; probe_row[c + 2] = h2; probe_row[c + 3] = h3;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 w = __ldg(reinterpret_cast<const float4*>(w11 + k * kL10Out + c));
                rgb[k] = fmaf(w.x, h0, rgb[k]); rgb[k] = fmaf(w.y, h1, rgb[k]);
                rgb[k] = fmaf(w.z, h2, rgb[k]); rgb[k] = fmaf(w.w, h3, rgb[k]);
            }
        }
    }
}

// ---------------------------------------------------------------------------- kernel
template <bool PROBE>
__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_kernel(const FwdParams P) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B atoms need 1024-byte aligned tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = umma::smem_u32(smem);
    const uint32_t bar_w_full = sbase + kOffBar;              // [kRing]
    const uint32_t bar_w_empty = bar_w_full + 8 * kRing;      // [kRing]
    const uint32_t bar_a_ready = bar_w_empty + 8 * kRing;     // [2]
    const uint32_t bar_acc_full = bar_a_ready + 16;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (2 * kRing + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long n_tiles = (P.M + kTileM - 1) / kTileM;
    const long n_pairs = (n_tiles + 1) / 2;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kRing; ++s) {
            umma::mbar_init(bar_w_full + 8 * s, 1);
            umma::mbar_init(bar_w_empty + 8 * s, 1);
        }
        for (int g = 0; g < 2; ++g) {
            umma::mbar_init(bar_a_ready + 8 * g, kTileM);
            umma::mbar_init(bar_acc_full + 8 * g, 1);
        }
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

    if (warp == 0) {
        // ===================== producer: weight stages, L2 -> smem =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int l = 0; l < kNumMmaLayers; ++l) {
                    const int first = layer_first_stage(l), cnt = layer_chunks(l) * layer_halves(l);
                    for (int g = 0; g < 2; ++g) {
                        for (int s = 0; s < cnt; ++s, ++it) {
                            const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
                            umma::mbar_wait(bar_w_empty + 8 * slot, ph ^ 1);
                            umma::mbar_arrive_expect_tx(bar_w_full + 8 * slot, kStageBytes);
                            umma::bulk_g2s(sbase + kOffW + slot * kStageBytes,
                                           P.blob + (size_t)(first + s) * kStageBytes, kStageBytes,
                                           bar_w_full + 8 * slot);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, n_ready[2] = {0, 0};
            for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int l = 0; l < kNumMmaLayers; ++l) {
                    const int chunks = layer_chunks(l), halves = layer_halves(l);
                    for (int g = 0; g < 2; ++g) {
                        umma::mbar_wait(bar_a_ready + 8 * g, n_ready[g] & 1);
                        ++n_ready[g];
                        umma::tc_fence_after();
                        const uint32_t d_base = tmem_base + g * 256;
                        const uint32_t a_tile = sbase + kOffA + g * 65536;
                        const uint32_t pe_tile = sbase + kOffPE + g * 16384;
                        for (int j = 0; j < chunks; ++j) {
                            uint32_t a_addr;
                            if (l == 0) a_addr = pe_tile;
                            else if (l == 5) a_addr = (j == 0) ? pe_tile : a_tile + (j - 1) * 16384;
                            else a_addr = a_tile + j * 16384;
                            for (int h = 0; h < halves; ++h, ++it) {
                                const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
                                umma::mbar_wait(bar_w_full + 8 * slot, ph);
                                umma::tc_fence_after();
                                const uint32_t b_addr = sbase + kOffW + slot * kStageBytes;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    umma::mma_bf16_ss(d_base + h * 128,
                                                      umma::smem_desc_sw128(a_addr + kk * 32),
                                                      umma::smem_desc_sw128(b_addr + kk * 32), kIdesc,
                                                      (j > 0 || kk > 0) ? 1u : 0u);
                                }
                                umma::mma_commit(bar_w_empty + 8 * slot);
                            }
                        }
                        umma::mma_commit(bar_acc_full + 8 * g);
                    }
                }
            }
        }
    } else {
        // ===================== epilogue groups =====================
        const int g = (warp - 2) >> 2;
        const int quad = warp & 3;               // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        uint8_t* a_tile = smem + kOffA + g * 65536;
        uint8_t* pe_tile = smem + kOffPE + g * 16384;
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + g * 256;
        const float* tail = reinterpret_cast<const float*>(P.blob + kWeightBytes);
        uint32_t n_full = 0;
        for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const long grow_raw = (pair * 2 + g) * kTileM + row;
            const bool valid = grow_raw < P.M;
            const long grow = valid ? grow_raw : P.M - 1;
            input_stage(P, grow, pe_tile, row);
            umma::fence_proxy_async_smem();
            umma::mbar_arrive(bar_a_ready + 8 * g);
            float sigma = 0.f;
            float* probe_row = nullptr;
#pragma unroll 1
            for (int l = 0; l < kNumMmaLayers; ++l) {
                umma::mbar_wait(bar_acc_full + 8 * g, n_full & 1);
                ++n_full;
                umma::tc_fence_after();
                if (PROBE) probe_row = (P.probe_out && P.probe_layer == l && valid) ? P.probe_out + grow * 256 : nullptr;
                if (l < 9) {
                    const float* bias = tail + kTailBias + l * kHidden;
                    if (l == 7) {
                        sigma = __ldg(tail + kTailBAlpha);
                        epilogue_hidden<1, PROBE>(tacc, a_tile, row, bias, tail + kTailWAlpha, sigma, probe_row);
                    } else if (l == 8) {
                        epilogue_hidden<2, PROBE>(tacc, a_tile, row, bias, nullptr, sigma, probe_row);
                    } else {
                        epilogue_hidden<0, PROBE>(tacc, a_tile, row, bias, nullptr, sigma, probe_row);
                    }
                    umma::fence_proxy_async_smem();
                    umma::tc_fence_before();
                    umma::mbar_arrive(bar_a_ready + 8 * g);
                } else {
                    float rgb[3];
                    const float* vt = P.vterm + (grow / P.vterm_div) * kL10Out;
                    epilogue_rgb<PROBE>(tacc, vt, tail + kTailW11, rgb, probe_row);
                    umma::tc_fence_before();
                    if (valid) {
                        float4 o;
                        o.x = rgb[0] + __ldg(tail + kTailB11 + 0);
                        o.y = rgb[1] + __ldg(tail + kTailB11 + 1);
                        o.z = rgb[2] + __ldg(tail + kTailB11 + 2);
                        o.w = sigma;
                        reinterpret_cast<float4*>(P.raw_out)[grow] = o;
                    }
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

int launch_fwd(const FwdParams& P, bool probe, void* stream) {
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            sm_count = 0;
            nerf::set_last_error("nerf_mlp_fwd setup: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    const long n_tiles = (P.M + kTileM - 1) / kTileM;
    const long n_pairs = (n_tiles + 1) / 2;
    const unsigned grid = (unsigned)(n_pairs < sm_count ? n_pairs : sm_count);
    if (probe) mlp_fwd_kernel<true><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    else mlp_fwd_kernel<false><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    return nerf::check_launch("nerf_mlp_fwd");
}

int fill_params(FwdParams& P, const void* packed, int in_mode, const float* in0, const float* in1,
                int in_stride, long M, int S, const float* vterm, int vterm_div, float* raw_out) {
    if (!packed || !in0 || !vterm || !raw_out || M < 0 || vterm_div < 1) return nerf::arg_error("nerf_mlp_fwd");
    if (in_mode == NERF_IN_RAYS) {
        if (!in1 || S < 1) return nerf::arg_error("nerf_mlp_fwd: NERF_IN_RAYS needs z and S");
    } else if (in_mode == NERF_IN_EMBEDDED) {
        if (in_stride < 63) return nerf::arg_error("nerf_mlp_fwd: in_stride < 63");
    } else if (in_mode != NERF_IN_POINTS) {
        return nerf::arg_error("nerf_mlp_fwd: in_mode");
    }
    P.blob = (const uint8_t*)packed;
    P.in_mode = in_mode; P.in0 = in0; P.in1 = in1; P.in_stride = in_stride;
    P.M = M; P.S = S < 1 ? 1 : S; P.vterm = vterm; P.vterm_div = vterm_div; P.raw_out = raw_out;
    P.probe_out = nullptr; P.probe_layer = -1;
    return 0;
}

}  // namespace

extern "C" int nerf_mlp_fwd(const void* packed, int in_mode, const float* in0, const float* in1,
                            int in_stride, long M, int S, const float* vterm, int vterm_div,
                            float* raw_out, void* act_save, void* stream) {
    if (act_save) { nerf::set_last_error("nerf_mlp_fwd: act_save not supported by this entry"); return NERF_ERR_UNSUPPORTED; }
    FwdParams P;
    int rc = fill_params(P, packed, in_mode, in0, in1, in_stride, M, S, vterm, vterm_div, raw_out);
    if (rc) return rc;
    if (M == 0) return 0;
    return launch_fwd(P, false, stream);
}

// Debug entry (tests only): additionally dumps the FP32 post-activation output of MMA layer
// `probe_layer` (0 = l1 ... 8 = l9, 9 = l10; 256 floats per row, l10 uses the first 128).
extern "C" int nerf_mlp_fwd_probe(const void* packed, int in_mode, const float* in0, const float* in1,
                                  int in_stride, long M, int S, const float* vterm, int vterm_div,
                                  float* raw_out, int probe_layer, float* probe_out, void* stream) {
    FwdParams P;
    int rc = fill_params(P, packed, in_mode, in0, in1, in_stride, M, S, vterm, vterm_div, raw_out);
    if (rc) return rc;
    if (M == 0) return 0;
    P.probe_out = probe_out; P.probe_layer = probe_layer;
    return launch_fwd(P, true, stream);
}

extern "C" size_t nerf_mlp_act_bytes(long M) { (void)M; return 0; }
