// K2: fused field-network forward for sm_100a.
//
// Replaces net_forward + FreqEmbedding.embed + Model.forward
// (/root/reference/model.py:9-31, 77-131) and the point construction of render_rays
// (/root/reference/main.py:238,252): sample position -> positional encoding -> 8x256 trunk with
// the skip at l6 -> sigma head, l9 folded into l10 (mlp_layout.h; + hoisted view term), l11 ->
// raw[rgb(3), sigma].
// Per-sample activations never leave the SM: BF16 activations live in shared memory, FP32
// accumulators in tensor memory.
//
// Structure (one persistent CTA per SM, 576 threads):
//   warp 0       producer: streams 32 KB weight slots (one 64-wide K chunk, all 256 output rows)
//                L2 -> smem with cp.async.bulk (TMA engine) through an mbarrier ring, in the order
//                mlp_layout.h stores them;
//   warp 1       MMA issuer: one thread issues tcgen05.mma (M=128, N=256, K=16, BF16 -> FP32 in
//                TMEM) and tcgen05.commit; owns the 512-column TMEM allocation;
//   warps 2-9    epilogue group X, warps 10-17 epilogue group Y.  A group owns one 128-sample
//                sub-tile; two threads share a sample row (= TMEM lane), each handling half of
//                the columns.  The group computes the positional encoding, and after every layer
//                reads the accumulator (tcgen05.ld), adds bias, applies ReLU, rounds to BF16 and
//                writes the next layer's A operand back to shared memory in the UMMA K-major
//                SWIZZLE_128B layout.  sigma (l_alpha) and rgb (l11) are evaluated in FP32 on
//                CUDA cores from the FP32 accumulators of l8 / l10.
// The two sub-tiles ping-pong: while the tensor core runs layer l of Y, group X runs the epilogue
// of layer l of X, so MMA and epilogue overlap.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "mlp_bwd_layout.h"
#include "mlp_layout.h"
#include "umma.cuh"

namespace {

using namespace nerf;

constexpr int kTileM = 128;
constexpr int kRing = 2;                              // 32 KB weight slots in flight
constexpr int kMaxRing = 8;
constexpr int kSlotBytes = 2 * kStageBytes;           // one K chunk, both N halves: [256][64] bf16
constexpr int kEpiWarpsPerGroup = 8;
constexpr int kThreads = 64 + 2 * kEpiWarpsPerGroup * 32;   // 576
// mlp_fwd_kernel can run a SECOND producer warp (warp 18, Cfg::two_producers): every 32 KB slot is then
// fetched as two 16 KB bulk copies issued by two threads (copies issued by one thread do not overlap:
// tools/probes/sw64_mma_probe.cu), which lands a slot ~150-200 cycles earlier.
constexpr int kThreadsFwd2 = kThreads + 32;           // 608
constexpr uint32_t kOffA = 0;                         // 2 x [4][128][64] bf16
constexpr uint32_t kOffPE = 2 * 65536;                // 2 x [128][64] bf16
constexpr uint32_t kOffW = kOffPE + 2 * 16384;        // ring x 32 KB
constexpr uint32_t kOffBar = kOffW + kRing * kSlotBytes;
constexpr uint32_t kOffBias = kOffBar + 256;          // 2 x [256] fp32: the current layer's bias, per group
constexpr uint32_t kOffPlan = kOffBias + 2 * 1024;     // the weight-ring plan of this launch (RingPlan, 412 bytes)
constexpr uint32_t kSmemBytes = kOffPlan + 416;
static_assert(kSmemBytes <= 232448, "shared memory budget");
constexpr uint32_t kIdescN256 = umma::instr_desc_bf16(128, 256);
constexpr uint32_t kIdescN128 = umma::instr_desc_bf16(128, 128);

// Kernel variants: kRing is the production ring depth.  The experiment variants (debug entry
// only) change the ring depth; ALIAS additionally places the PE tiles on top of the A tiles to
// free 32 KB for a deeper ring -- numerically wrong, used only to time the pipeline.
// EXP (timing experiments, wrong numerics): bit0 skip the A-tile stores, bit1 skip the bias loads,
// bit2 skip the TMEM loads, bit3 no weight streaming at all (the MMAs read whatever the ring holds:
// the speed the kernel would have if weight slots were always ready).
// PEA ("PE in the A tile", inference): the positional-encoding tile does not get shared memory of its
// own.  It is written into block 0 of the sub-tile's A tile for l1, and ENCODED AGAIN into the same
// block for l6 once the MMAs that read h5's block 0 have completed (l6 accumulates its four h5 chunks
// first and the PE chunk last; the group's re-encoding runs under the other sub-tile's MMAs).  The
// 32 KB this frees hold a THIRD weight slot: measured with the timing-only ALIAS layout, 3 x 32 KB
// instead of 2 x 32 KB takes 7.4 % off the kernel's cycles (the refill of a slot takes ~1000-1400
// cycles after its MMAs complete; a slot is consumed in 640).
template <int RING, bool ALIAS, int EXP = 0, bool PEA = false>
struct Cfg {
    static constexpr int ring = RING;
    static constexpr int exp = EXP;
    static constexpr bool pea = PEA;
    static constexpr uint32_t off_pe = ALIAS ? 0u : kOffPE;
    static constexpr uint32_t off_w = (ALIAS || PEA) ? kOffPE : kOffW;
    static_assert(!(ALIAS && PEA), "ALIAS is the timing-only precursor of PEA");
    static_assert(off_w + RING * kSlotBytes <= kOffBar, "ring does not fit");
};

// Host copy of the head of the fp32 tail, passed BY VALUE inside the kernel parameters: the
// epilogues then read biases / l_alpha / l11 through the constant bank with uniform loads
// (LDCU + FADD2 R, R, UR), which costs nothing on the L1 data pipe the tensor core needs for its
// operands (a warp-uniform LDS.128 costs 2 wavefronts there, an LDG.128 four).
struct alignas(16) ConstTail {
    float bias[9 * kHidden];
    float walpha[kHidden];
    float balpha[4];
    float w11[3 * kL10Out];
    float b11[4];
};
static_assert(sizeof(ConstTail) == (size_t)kTailW10View * 4, "ConstTail mirrors the head of the device tail");

struct FwdParams {
    ConstTail ct;            // valid for the CT kernel variants only
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
    long long* stats_out;    // debug: [grid][8] cycle counters (or NULL)
    long long* trace_out;    // debug: [4 roles][1024] time-stamped events of CTA 0's 4th and 5th tile pair (or NULL)
    int tile0;               // parity of the global index of this launch's first 128-row tile (chunk order, see chunk_at)
    uint8_t* act_save;       // training: activation records, kActTileBytes per 128-row tile (or NULL)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// max(x, 0) folded into the conversion (F2FP.RELU)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// ---------------------------------------------------------------------------- input stage
// Positional encoding of one point, model.py:15-31: [x, sin(2^k x), cos(2^k x)]_{k<10}, 63 wide,
// stored as one 128-byte swizzled row (column 63 = 0).  The two threads of a row split it:
// HALF 0 produces columns 0..31 (x and octaves 0..3 plus most of octave 4) from a sincosf anchor
// at octave 0, HALF 1 columns 32..63 (cos(16 z), octaves 5..9, zero pad) from an anchor at
// octave 5; the other octaves follow by angle doubling.  The accumulated error (<1e-5) is far
// below the BF16 rounding applied next.
template <int HALF>
__device__ __forceinline__ void encode_half(float px, float py, float pz, float (&f)[32]) {
    const float p[3] = {px, py, pz};
    if (HALF == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            f[a] = p[a];
            float s, c;
            sincosf(p[a], &s, &c);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                if (k > 0) {
                    float s2 = 2.f * s * c;
                    float c2 = fmaf(-2.f * s, s, 1.f);
                    s = s2; c = c2;
                }
                if (3 + 6 * k + a < 32) f[3 + 6 * k + a] = s;
                if (3 + 6 * k + 3 + a < 32) f[3 + 6 * k + 3 + a] = c;
            }
        }
    } else {
        f[0] = cosf(pz * 16.f);                       // column 32 = cos(2^4 z)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float s, c;
            sincosf(p[a] * 32.f, &s, &c);
#pragma unroll
            for (int k = 5; k < 10; ++k) {
                if (k > 5) {
                    float s2 = 2.f * s * c;
                    float c2 = fmaf(-2.f * s, s, 1.f);
                    s = s2; c = c2;
                }
                f[3 + 6 * k + a - 32] = s;
                f[3 + 6 * k + 3 + a - 32] = c;
            }
        }
        f[31] = 0.f;
    }
}

// the 32 PE columns of this thread's half row as four packed 16-byte chunks
template <int HALF>
__device__ __forceinline__ void input_encode(const FwdParams& P, long grow, uint4 (&o)[4]) {
    float f[32];
    if (P.in_mode == NERF_IN_EMBEDDED) {
        const float* x = P.in0 + grow * P.in_stride + HALF * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) f[c] = (HALF * 32 + c < kPeDim) ? __ldg(x + c) : 0.f;
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
        encode_half<HALF>(px, py, pz, f);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        o[q].x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
        o[q].y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
        o[q].z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
        o[q].w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
    }
}

template <int HALF>
__device__ __forceinline__ void input_store(const uint4 (&o)[4], uint8_t* pe_tile, int row) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c16 = HALF * 4 + q;
        *reinterpret_cast<uint4*>(pe_tile + row * 128 + ((c16 ^ (row & 7)) << 4)) = o[q];
    }
}

template <int HALF>
__device__ __forceinline__ void input_stage(const FwdParams& P, long grow, uint8_t* pe_tile, int row) {
    uint4 o[4];
    input_encode<HALF>(P, grow, o);
    input_store<HALF>(o, pe_tile, row);
}

// ---------------------------------------------------------------------------- epilogues
__device__ __forceinline__ void load_bias16(const float* __restrict__ b, float4 (&dst)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = __ldg(reinterpret_cast<const float4*>(b) + i);
}
// 16 floats of the staged bias: warp-uniform LDS.128 (a broadcast costs one wavefront, where a
// warp-uniform LDG.128 costs four on the same L1 data pipe the tensor core reads its operands through)
__device__ __forceinline__ void lds_bias16(uint32_t addr, float4 (&dst)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = umma::ld_shared_v4f(addr + i * 16);
}

// Hidden layer, this thread's 128 columns [c0, c0+128): h = act(acc + bias) -> BF16 -> A tile (in
// place).  MODE 0: ReLU; 1: ReLU and accumulate the FP32 sigma head (l_alpha); 2: no activation
// (l9).  16 columns per step; the next step's TMEM load and bias loads are in flight while the
// current step is processed.  The bias add runs on packed FP32 pairs (FADD2) and the ReLU is
// folded into the BF16 conversion (F2FP.RELU), so a column costs about one issue slot.
// row_addr: shared-space address of this row's 128-byte line in block 0 of the A tile;
// swz = (row & 7) << 4.
// CT: bias and l_alpha come from the kernel parameters (ct, layer l) instead of shared / global memory.
// L >= 0: the layer index is a compile-time constant, so with CT every bias is an immediate
// constant-bank operand of its FADD2 (no load instruction at all); L = -1: runtime index l.
// [x > 0] of a post-ReLU BF16 pair as an all-ones/all-zeros mask per 16-bit half (HSET2.BM)
__device__ __forceinline__ uint32_t bf16x2_positive_mask(uint32_t w) {
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
    return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w), zero);
}

// MASKS (training): additionally returns the ReLU mask bits of this thread's columns in mw[0..3]
// (layout in mlp_bwd_layout.h).
// C0 >= 0: this thread's first column is a compile-time constant too (the callers branch on the column
// half), which turns every bias / l_alpha address into a constant-bank immediate.
template <int MODE, bool PROBE, int EXP, bool CT, int NIT = 8, int L = -1, bool MASKS = false, int C0 = -1>
__device__ __forceinline__ void epilogue_hidden(uint32_t tacc, int c0_rt, uint32_t row_addr, uint32_t swz,
                                                uint32_t bias_addr,
                                                const float* __restrict__ walpha, float& sigma,
                                                float* probe_row, const ConstTail& ct, int l,
                                                uint32_t* mw = nullptr) {
    const int c0 = C0 >= 0 ? C0 : c0_rt;
    // NB accumulator buffers: tcgen05.wait::ld waits for every load issued so far, so with NB = 3 the
    // youngest load outstanding at a wait is one whole step old (two steps of cover per load).
    // EXP bit6 selects three (A/B): +2.5..7 % in short bursts, -1.6 % over 60 interleaved rounds of
    // sustained load (the kernel spills at its 96-register cap), so two stay the default.
    constexpr int NB = (CT && C0 >= 0 && (EXP & 64)) ? 3 : 2;
    uint32_t v[NB][16] = {};
    float4 b[2][4] = {};
    float2 sig2 = make_float2(0.f, 0.f);
    if (!(EXP & 4)) umma::tmem_ld16(tacc + c0, v[0]);
    if (NB == 3 && NIT > 1 && !(EXP & 4)) umma::tmem_ld16(tacc + c0 + 16, v[1]);
    if (!CT && !(EXP & 2)) lds_bias16(bias_addr + c0 * 4, b[0]);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int c = c0 + it * 16;
        if (!CT && it + 1 < NIT && !(EXP & 2)) lds_bias16(bias_addr + (c + 16) * 4, b[(it + 1) & 1]);
        if (!(EXP & 4)) umma::tmem_wait_ld();
        if (it + NB - 1 < NIT && !(EXP & 4)) umma::tmem_ld16(tacc + c + 16 * (NB - 1), v[(it + NB - 1) % NB]);
        const uint32_t(&cur)[16] = v[it % NB];
        const float4(&bc)[4] = b[it & 1];
        float2 h[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 b0 = make_float2(bc[q].x, bc[q].y), b1 = make_float2(bc[q].z, bc[q].w);
            if (CT && !(EXP & 2)) {
                const float4 bl = *reinterpret_cast<const float4*>(ct.bias + (L >= 0 ? L : l) * kHidden + c + q * 4);
                b0 = make_float2(bl.x, bl.y);
                b1 = make_float2(bl.z, bl.w);
            }
            h[2 * q] = __fadd2_rn(make_float2(__uint_as_float(cur[q * 4 + 0]), __uint_as_float(cur[q * 4 + 1])), b0);
            h[2 * q + 1] = __fadd2_rn(make_float2(__uint_as_float(cur[q * 4 + 2]), __uint_as_float(cur[q * 4 + 3])), b1);
        }
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 w;
                if (CT) w = *reinterpret_cast<const float4*>(ct.walpha + c + q * 4);
                else w = __ldg(reinterpret_cast<const float4*>(walpha + c) + q);
                h[2 * q].x = fmaxf(h[2 * q].x, 0.f); h[2 * q].y = fmaxf(h[2 * q].y, 0.f);
                h[2 * q + 1].x = fmaxf(h[2 * q + 1].x, 0.f); h[2 * q + 1].y = fmaxf(h[2 * q + 1].y, 0.f);
                sig2 = __ffma2_rn(make_float2(w.x, w.y), h[2 * q], sig2);
                sig2 = __ffma2_rn(make_float2(w.z, w.w), h[2 * q + 1], sig2);
            }
        }
        if (PROBE && probe_row) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                probe_row[c + 2 * e] = MODE == 0 ? fmaxf(h[e].x, 0.f) : h[e].x;
                probe_row[c + 2 * e + 1] = MODE == 0 ? fmaxf(h[e].y, 0.f) : h[e].y;
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint32_t o0, o1, o2, o3;
            if (MODE == 0) {
                o0 = pack_relu_bf16x2(h[q * 4 + 0].x, h[q * 4 + 0].y); o1 = pack_relu_bf16x2(h[q * 4 + 1].x, h[q * 4 + 1].y);
                o2 = pack_relu_bf16x2(h[q * 4 + 2].x, h[q * 4 + 2].y); o3 = pack_relu_bf16x2(h[q * 4 + 3].x, h[q * 4 + 3].y);
            } else {
                o0 = pack_bf16x2(h[q * 4 + 0].x, h[q * 4 + 0].y); o1 = pack_bf16x2(h[q * 4 + 1].x, h[q * 4 + 1].y);
                o2 = pack_bf16x2(h[q * 4 + 2].x, h[q * 4 + 2].y); o3 = pack_bf16x2(h[q * 4 + 3].x, h[q * 4 + 3].y);
            }
            if (MASKS && MODE != 2) {
                uint32_t& m = mw[it >> 1];
                if ((it & 1) == 0 && q == 0) m = 0;
                m |= bf16x2_positive_mask(o0) & relu_mask_bits(it & 1, q * 4 + 0);
                m |= bf16x2_positive_mask(o1) & relu_mask_bits(it & 1, q * 4 + 1);
                m |= bf16x2_positive_mask(o2) & relu_mask_bits(it & 1, q * 4 + 2);
                m |= bf16x2_positive_mask(o3) & relu_mask_bits(it & 1, q * 4 + 3);
            }
            const int cc = c + q * 8;
            const int blk = cc >> 6, c16 = (cc & 63) >> 3;
            if (!(EXP & 1) || o0 == 0x12345678u)
                umma::st_shared_v4(row_addr + blk * 16384 + ((uint32_t)(c16 << 4) ^ swz), o0, o1, o2, o3);
        }
    }
    if (MODE == 1) sigma += sig2.x + sig2.y;
}

// The same arithmetic with 32-column TMEM loads (inference, biases in the kernel parameters): four
// tcgen05.ld.x32 per thread and layer instead of eight .x16, each in flight while the previous 32 columns
// are processed.  tcgen05.wait::ld waits for every outstanding load of the thread, so the depth of the
// pipeline is one load whatever its width; wider loads halve the number of exposed load latencies, which
// is what the sub-tile's accumulator -> A-operand chain consists of (the instructions themselves need
// ~400 issue cycles of the ~1950 the epilogue takes).  Bit-identical results.
template <int MODE, int C0>
__device__ __forceinline__ void epilogue_hidden_w32(uint32_t tacc, uint32_t row_addr, uint32_t swz, float& sigma,
                                                    const ConstTail& ct, int l) {
    uint32_t v[2][32];
    float2 sig2 = make_float2(0.f, 0.f);
    umma::tmem_ld32(tacc + C0, v[0]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        umma::tmem_wait_ld();
        if (g + 1 < 4) umma::tmem_ld32(tacc + C0 + 32 * (g + 1), v[(g + 1) & 1]);
        const uint32_t(&cur)[32] = v[g & 1];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int c = C0 + g * 32 + hh * 16;
            float2 h[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 bl = *reinterpret_cast<const float4*>(ct.bias + l * kHidden + c + q * 4);
                h[2 * q] = __fadd2_rn(make_float2(__uint_as_float(cur[hh * 16 + q * 4 + 0]), __uint_as_float(cur[hh * 16 + q * 4 + 1])),
                                      make_float2(bl.x, bl.y));
                h[2 * q + 1] = __fadd2_rn(make_float2(__uint_as_float(cur[hh * 16 + q * 4 + 2]), __uint_as_float(cur[hh * 16 + q * 4 + 3])),
                                          make_float2(bl.z, bl.w));
            }
            if (MODE == 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w = *reinterpret_cast<const float4*>(ct.walpha + c + q * 4);
                    h[2 * q].x = fmaxf(h[2 * q].x, 0.f); h[2 * q].y = fmaxf(h[2 * q].y, 0.f);
                    h[2 * q + 1].x = fmaxf(h[2 * q + 1].x, 0.f); h[2 * q + 1].y = fmaxf(h[2 * q + 1].y, 0.f);
                    sig2 = __ffma2_rn(make_float2(w.x, w.y), h[2 * q], sig2);
                    sig2 = __ffma2_rn(make_float2(w.z, w.w), h[2 * q + 1], sig2);
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t o0, o1, o2, o3;
                if (MODE == 0) {
                    o0 = pack_relu_bf16x2(h[q * 4 + 0].x, h[q * 4 + 0].y); o1 = pack_relu_bf16x2(h[q * 4 + 1].x, h[q * 4 + 1].y);
                    o2 = pack_relu_bf16x2(h[q * 4 + 2].x, h[q * 4 + 2].y); o3 = pack_relu_bf16x2(h[q * 4 + 3].x, h[q * 4 + 3].y);
                } else {
                    o0 = pack_bf16x2(h[q * 4 + 0].x, h[q * 4 + 0].y); o1 = pack_bf16x2(h[q * 4 + 1].x, h[q * 4 + 1].y);
                    o2 = pack_bf16x2(h[q * 4 + 2].x, h[q * 4 + 2].y); o3 = pack_bf16x2(h[q * 4 + 3].x, h[q * 4 + 3].y);
                }
                const int cc = c + q * 8;
                const int blk = cc >> 6, c16 = (cc & 63) >> 3;
                umma::st_shared_v4(row_addr + blk * 16384 + ((uint32_t)(c16 << 4) ^ swz), o0, o1, o2, o3);
            }
        }
    }
    if (MODE == 1) sigma += sig2.x + sig2.y;
}

// CT epilogue of hidden layer l (0..8) with the layer index turned into a template constant.
template <int NIT, int EXP = 0, int C0 = -1>
__device__ __forceinline__ void epilogue_hidden_ct(int l, uint32_t tacc, int c0, uint32_t row_addr, uint32_t swz,
                                                   float& sigma, const ConstTail& ct) {
    // One copy of the code per activation MODE, not per layer: the layer index stays a (warp-uniform)
    // run-time value, so a bias address is "uniform register + immediate" and the loads stay on the uniform
    // datapath, while the kernel loses six of its nine unrolled epilogue bodies per column half -- the
    // instruction cache was missing 17 % of its requests (ncu sm__icc_request_hit_rate 82.7 %).
    if (EXP & 8192) {         // EXP bit13 (A/B): one body per layer (every bias address an immediate)
        switch (l) {
#define NERF_CT_LAYER(LL, MODE) \
            case LL: epilogue_hidden<MODE, false, EXP, true, NIT, LL, false, C0>(tacc, c0, row_addr, swz, 0, nullptr, sigma, nullptr, ct, LL); break;
            NERF_CT_LAYER(0, 0) NERF_CT_LAYER(1, 0) NERF_CT_LAYER(2, 0) NERF_CT_LAYER(3, 0) NERF_CT_LAYER(4, 0)
            NERF_CT_LAYER(5, 0) NERF_CT_LAYER(6, 0)
            default: epilogue_hidden<1, false, EXP, true, NIT, 7, false, C0>(tacc, c0, row_addr, swz, 0, nullptr, sigma, nullptr, ct, 7); break;
#undef NERF_CT_LAYER
        }
        return;
    }
    if ((EXP & 16384) && NIT == 8 && C0 >= 0) {     // EXP bit14: 32-column TMEM loads
        constexpr int C = C0 >= 0 ? C0 : 0;
        if (l < 7) epilogue_hidden_w32<0, C>(tacc, row_addr, swz, sigma, ct, (int)__reduce_max_sync(0xffffffffu, (unsigned)l));
        else epilogue_hidden_w32<1, C>(tacc, row_addr, swz, sigma, ct, 7);
        return;
    }
    // (the warp-wide reduction hands the compiler a value it knows to be uniform: REDUX writes a uniform register)
    if (l < 7) epilogue_hidden<0, false, EXP, true, NIT, -1, false, C0>(tacc, c0, row_addr, swz, 0, nullptr, sigma, nullptr, ct,
                                                                          (int)__reduce_max_sync(0xffffffffu, (unsigned)l));
    else epilogue_hidden<1, false, EXP, true, NIT, 7, false, C0>(tacc, c0, row_addr, swz, 0, nullptr, sigma, nullptr, ct, 7);
}

// l10 (+ hoisted view term, ReLU) and l11 in FP32 over this thread's 64 columns [c0, c0+64):
// partial rgb_raw.
// SAVE: h10 (post-ReLU, BF16) is written to blocks 0..1 of the A tile for the activation record.
// C0 >= 0: the first column as a compile-time constant, so that the host-tail build addresses l11's weights
// with immediate offsets into the parameter bank (uniform loads feeding FFMA2 directly; a runtime c0 turns
// them into 96 register-indexed constant loads per thread, which cost the tile boundary ~1500 cycles).
template <bool PROBE, bool SAVE, bool CT, int NIT = 4, int C0 = -1>
__device__ __forceinline__ void epilogue_rgb(uint32_t tacc, int c0_rt, const float* __restrict__ vt,
                                             const float* __restrict__ w11, float (&rgb)[3],
                                             float* probe_row, uint32_t row_addr, uint32_t swz, const ConstTail& ct,
                                             uint32_t* mw = nullptr) {
    const int c0 = C0 >= 0 ? C0 : c0_rt;
    uint32_t v[2][16];
    float4 t[2][4];
    umma::tmem_ld16(tacc + c0, v[0]);
    load_bias16(vt + c0, t[0]);
    float2 acc[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int c = c0 + it * 16;
        if (it + 1 < NIT) load_bias16(vt + c + 16, t[(it + 1) & 1]);
        umma::tmem_wait_ld();
        if (it + 1 < NIT) umma::tmem_ld16(tacc + c + 16, v[(it + 1) & 1]);
        const uint32_t(&cur)[16] = v[it & 1];
        const float4(&tc)[4] = t[it & 1];
        uint32_t pk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 ha = __fadd2_rn(make_float2(__uint_as_float(cur[q * 4 + 0]), __uint_as_float(cur[q * 4 + 1])),
                                   make_float2(tc[q].x, tc[q].y));
            float2 hb = __fadd2_rn(make_float2(__uint_as_float(cur[q * 4 + 2]), __uint_as_float(cur[q * 4 + 3])),
                                   make_float2(tc[q].z, tc[q].w));
            ha.x = fmaxf(ha.x, 0.f); ha.y = fmaxf(ha.y, 0.f);
            hb.x = fmaxf(hb.x, 0.f); hb.y = fmaxf(hb.y, 0.f);
            if (PROBE && probe_row) {
                float* pr = probe_row + c + q * 4;
                pr[0] = ha.x; pr[1] = ha.y; pr[2] = hb.x; pr[3] = hb.y;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 w;
                if (CT) {
                    w = *reinterpret_cast<const float4*>(ct.w11 + k * kL10Out + c + q * 4);
                } else {
                    w = __ldg(reinterpret_cast<const float4*>(w11 + k * kL10Out + c) + q);
                }
                acc[k] = __ffma2_rn(make_float2(w.x, w.y), ha, acc[k]);
                acc[k] = __ffma2_rn(make_float2(w.z, w.w), hb, acc[k]);
            }
            if (SAVE) {
                pk[q * 2 + 0] = pack_bf16x2(ha.x, ha.y);
                pk[q * 2 + 1] = pack_bf16x2(hb.x, hb.y);
            }
        }
        if (SAVE) {
            uint32_t& m = mw[it >> 1];
            if ((it & 1) == 0) m = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) m |= bf16x2_positive_mask(pk[j]) & relu_mask_bits(it & 1, j);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int cc = c + q * 8;
                const int blk = cc >> 6, c16 = (cc & 63) >> 3;
                umma::st_shared_v4(row_addr + blk * 16384 + ((uint32_t)(c16 << 4) ^ swz), pk[q * 4 + 0], pk[q * 4 + 1],
                                   pk[q * 4 + 2], pk[q * 4 + 3]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[k] = acc[k].x + acc[k].y;
}

// ---------------------------------------------------------------------------- weight ring plan
// Order in which a sub-tile consumes the K chunks of layer l.  Tiles of even GLOBAL index (FwdParams::tile0
// + local tile) run forward (0, 1, .., n-1), odd ones backward, so that the sub-tile issued second starts with
// the chunk the first one finished with: the three slots of the ring still hold the first sub-tile's last
// three chunks, and of the 8 chunk uses of a layer only 5 need a copy from L2 (43 instead of 68 per tile pair:
// the kernel was bound by the refill chain of its three 32 KB slots, not by the tensor pipe).  l6 keeps its
// PE chunk (0) last and the chunk that reads A block 0 (1) ahead of a later h5 chunk in both orders (PEA).
// The summation order of a row therefore depends on the parity of its global 128-row tile; callers that shard
// a batch at multiples of 128 rows and pass the shard's first row get bit-identical results.
__host__ __device__ constexpr int chunk_at(int l, bool rev, int idx) {
    if (l == 5) {
        if (!rev) return idx == 4 ? 0 : idx + 1;                                   // 1 2 3 4 0
        return idx == 0 ? 4 : idx == 1 ? 3 : idx == 2 ? 1 : idx == 3 ? 2 : 0;      // 4 3 1 2 0
    }
    return rev ? layer_chunks(l) - 1 - idx : idx;
}

// The first sub-tile keeps (does not release) its last RING chunks, provided the second sub-tile starts with one
// of them (otherwise nothing is kept and the second sub-tile fetches everything: the plan can then never run
// out of free slots).
// Slot bookkeeping.  Fills take the free slots in the order they were released (a FIFO).  The plan is a pure
// function of (ring depth, tile parity) and periodic over one to three tile pairs (the slots come back
// permuted after a pair), so it is worked out at COMPILE time into a table of 16-bit entries in issue order,
// one per (layer, sub-tile, chunk) = 68 per tile pair, copied to shared memory at kernel start.  The one
// thread that issues tcgen05.mma then runs a flat loop over the entries, fetching the next one under the
// current chunk's MMAs: whatever it executes or waits for between two chunks is tensor idle time (the same
// bookkeeping done in that thread made the kernel 1.5x slower; nested loops with two wait sites, +20 %).
//   bits 0-2  chunk (source of the copy)             bit 9   LAST: last chunk of the sub-tile's layer
//   bits 3-5  slot                                            (commit the accumulator)
//   bit 6     first use of a fill (wait / copy)       bit 10  PE chunk of l6 (PEA: wait for the restored block)
//   bit 7     release the slot after use              bit 11  chunk that frees A block 0 in l6 (PEA: signal it)
//   bit 8     sub-tile (issue order)                  bits 12-13 A block read by the chunk
//                                                     bit 15  A operand is the PE tile (layouts with PE tiles)
constexpr int kPlanMaxPairs = 3;
constexpr int kPlanPerPair = 68;
template <int MAXP>
struct RingPlanT {
    unsigned short e[MAXP * kPlanPerPair];
    unsigned short period, pad;
};
using RingPlan = RingPlanT<kPlanMaxPairs>;
struct RingPlans {
    RingPlan p[3][2];      // [ring depth - 1][tile parity]
};
enum : unsigned {
    kPlanSlotShift = 3, kPlanSlotMask = 7,
    kPlanFirstUse = 1u << 6, kPlanRelease = 1u << 7, kPlanG = 1u << 8, kPlanLast = 1u << 9,
    kPlanPeChunk = 1u << 10, kPlanFreesBlock0 = 1u << 11, kPlanPeSrc = 1u << 15
};

template <int MAXP = kPlanMaxPairs>
constexpr RingPlanT<MAXP> make_ring_plan(int ring, bool rev0, bool reuse = true, bool both_forward = false) {
    RingPlanT<MAXP> t{};
    int fifo[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    int head = 0, count = ring;
    int where[5] = {-1, -1, -1, -1, -1};
    t.period = 0;
    for (int pair = 0; pair < 2 * MAXP; ++pair) {
        int q = 0;
        for (int l = 0; l < kNumMmaLayers; ++l) {
            const int n = layer_chunks(l);
            int keep_from = n > ring ? n - ring : 0;
            {
                const int want = chunk_at(l, !rev0, 0);
                bool found = false;
                for (int idx = keep_from; idx < n; ++idx) found = found || chunk_at(l, rev0, idx) == want;
                if (!found || !reuse) keep_from = n;
            }
            for (int g = 0; g < 2; ++g) {
                for (int idx = 0; idx < n; ++idx, ++q) {
                    const int c = chunk_at(l, both_forward ? false : (g == 0 ? rev0 : !rev0), idx);
                    const bool first_use = where[c] < 0;
                    if (first_use) {
                        where[c] = fifo[head];
                        head = (head + 1) & 7;
                        --count;
                    }
                    const bool release = g == 1 || idx < keep_from;
                    const int a_block = l == 0 ? 0 : (l == 5 ? (c == 0 ? 0 : c - 1) : c);
                    unsigned e = (unsigned)c | ((unsigned)where[c] << kPlanSlotShift) | (first_use ? kPlanFirstUse : 0u) |
                                 (release ? kPlanRelease : 0u) | (g ? kPlanG : 0u) |
                                 (idx == n - 1 ? kPlanLast : 0u) | ((l == 5 && c == 0) ? kPlanPeChunk : 0u) |
                                 ((l == 5 && c == 1) ? kPlanFreesBlock0 : 0u) | ((unsigned)a_block << 12) |
                                 ((l == 0 || (l == 5 && c == 0)) ? kPlanPeSrc : 0u);
                    if (pair < MAXP) t.e[pair * kPlanPerPair + q] = (unsigned short)e;
                    if (release) {
                        fifo[(head + count) & 7] = where[c];
                        ++count;
                        where[c] = -1;
                    }
                }
            }
        }
        // the plan repeats once the free slots are back in their initial order
        bool initial = true;
        for (int i = 0; i < ring; ++i) initial = initial && fifo[(head + i) & 7] == i;
        if (initial && t.period == 0) t.period = (unsigned short)(pair + 1);
    }
    return t;
}
static_assert(2 * (1 + 7 * 4 + 5) == kPlanPerPair, "entries per tile pair");

constexpr RingPlans make_ring_plans() {
    RingPlans r{};
    for (int ring = 1; ring <= 3; ++ring)
        for (int par = 0; par < 2; ++par) r.p[ring - 1][par] = make_ring_plan(ring, par != 0);
    return r;
}
constexpr bool ring_plans_ok(const RingPlans& r) {
    for (int ring = 0; ring < 3; ++ring)
        for (int par = 0; par < 2; ++par)
            if (r.p[ring][par].period < 1 || r.p[ring][par].period > kPlanMaxPairs) return false;
    return true;
}
static_assert(ring_plans_ok(make_ring_plans()), "the ring plan must repeat within kPlanMaxPairs tile pairs");

__constant__ RingPlans c_ring_plans = make_ring_plans();

// CTA-pair kernel: five 16 KB slots (each CTA stages its half of a chunk).  From five slots on the second
// sub-tile finds EVERY chunk of the layer resident: 34 fills per tile pair, each chunk of the model once.  (A
// sixth slot -- two chunks of the next layer fetched ahead instead of one -- measured 1.4 % SLOWER: 3.292e7
// against 3.246e7 cycles per CTA.)
constexpr int kPairRing = 5;
constexpr int kPairPlanMaxPairs = 3;
using PairPlan = RingPlanT<kPairPlanMaxPairs>;
struct PairPlans {
    PairPlan p[2];         // [tile parity]
};
constexpr PairPlans make_pair_plans() {
    PairPlans r{};
    for (int par = 0; par < 2; ++par) r.p[par] = make_ring_plan<kPairPlanMaxPairs>(kPairRing, par != 0);
    return r;
}
static_assert(make_pair_plans().p[0].period >= 1 && make_pair_plans().p[0].period <= kPairPlanMaxPairs &&
              make_pair_plans().p[1].period >= 1 && make_pair_plans().p[1].period <= kPairPlanMaxPairs, "pair plan period");
__constant__ PairPlans c_pair_plans = make_pair_plans();
#ifdef NERF_B200_EXPERIMENTS
constexpr RingPlans make_ring_plans_noreuse() {      // A/B: the same chunk orders, every use copies its chunk again
    RingPlans r{};
    for (int ring = 1; ring <= 3; ++ring)
        for (int par = 0; par < 2; ++par) r.p[ring - 1][par] = make_ring_plan(ring, par != 0, false);
    return r;
}
__constant__ RingPlans c_ring_plans_noreuse = make_ring_plans_noreuse();
constexpr RingPlans make_ring_plans_round1() {       // A/B: both sub-tiles forward, no re-use (the round-1 schedule)
    RingPlans r{};
    for (int ring = 1; ring <= 3; ++ring)
        for (int par = 0; par < 2; ++par) r.p[ring - 1][par] = make_ring_plan(ring, false, false, true);
    return r;
}
__constant__ RingPlans c_ring_plans_round1 = make_ring_plans_round1();
static_assert(ring_plans_ok(make_ring_plans_noreuse()) && ring_plans_ok(make_ring_plans_round1()), "plan period");
#endif

// ---------------------------------------------------------------------------- kernel
// WIDE (inference, CT only): all 16 epilogue warps work on ONE sub-tile at a time (four threads per
// row, 64 columns each) and alternate between the two sub-tiles, instead of 8 warps per sub-tile.
// The MMA -> epilogue -> MMA chain of a sub-tile is latency-bound on the epilogue; with twice the
// warps on it the next layer's operand is ready in about half the time and the tensor core idles less.
template <bool PROBE, class CFG, bool SAVE = false, bool CT = false, bool WIDE = false>
__device__ __forceinline__ void mlp_fwd_body(const FwdParams& P) {
    static_assert(!WIDE || (CT && !SAVE && !PROBE), "WIDE is an inference-only variant");
    constexpr bool kPEA = CFG::pea;
    static_assert(!kPEA || (!WIDE && !PROBE), "PEA is implemented for the 8-warp epilogue groups");
    constexpr bool kStageBias = !CT;          // per-layer bias staged in shared memory
    constexpr bool kGroupSync = !CT || SAVE;  // group barriers around the staging / the record copies
    constexpr int kRing = CFG::ring;
    constexpr uint32_t kOffPE = CFG::off_pe, kOffW = CFG::off_w;
    // SWIZZLE_128B atoms need 1024-byte aligned tiles (the kernel has no static shared memory, so
    // the dynamic segment starts at the aligned base of the CTA's window; checked below)
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    if ((sbase & 1023u) != 0) __trap();
    const uint32_t bar_w_full = sbase + kOffBar;              // [kMaxRing]
    const uint32_t bar_w_empty = bar_w_full + 8 * kMaxRing;   // [kMaxRing]
    const uint32_t bar_a_ready = bar_w_empty + 8 * kMaxRing;  // [2]
    const uint32_t bar_acc_full = bar_a_ready + 16;           // [2]
    const uint32_t bar_pe_free = bar_acc_full + 16;           // [2]  PEA: l6's MMAs on A block 0 (h5) completed
    const uint32_t bar_unused = bar_w_full + 8 * (kMaxRing - 1);   // takes the commits of chunks that keep their slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (2 * kMaxRing + 6));
    static_assert(8 * (2 * kMaxRing + 6) + 4 <= 256, "barrier region");
    long long t_wait0 = 0, t_wait1 = 0, t_begin = 0;
    if (PROBE || P.stats_out) t_begin = clock64();
    // event trace (PROBE only): role 0 producer, 1 MMA issuer, 2 / 3 first warp of epilogue group X / Y;
    // entry = tag << 56 | layer << 48 | group << 44 | chunk << 40 | cycles since kernel start
    int n_ev = 0;
    auto rec = [&](int role, long pair_no, int tag, int l, int g, int j) {
        if (!PROBE || !P.trace_out || blockIdx.x != 0 || pair_no < 3 || pair_no > 4 || n_ev >= 1024) return;
        P.trace_out[role * 1024 + n_ev++] = ((long long)tag << 56) | ((long long)l << 48) | ((long long)g << 44) |
                                            ((long long)j << 40) | ((clock64() - t_begin) & 0xFFFFFFFFFFll);
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long n_tiles = (P.M + kTileM - 1) / kTileM;
    const long n_pairs = (n_tiles + 1) / 2;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kRing; ++s) {
            umma::mbar_init(bar_w_full + 8 * s, 1);
            umma::mbar_init(bar_w_empty + 8 * s, 1);
        }
        static_assert(kRing < kMaxRing, "the last full-barrier slot is the unused one");
        umma::mbar_init(bar_unused, 1);
        for (int g = 0; g < 2; ++g) {
            umma::mbar_init(bar_a_ready + 8 * g, (WIDE ? 2 : 1) * kEpiWarpsPerGroup * 32);
            umma::mbar_init(bar_acc_full + 8 * g, 1);
            umma::mbar_init(bar_pe_free + 8 * g, 1);
        }
        umma::fence_barrier_init();
    }
    if (warp == 1) {
        umma::tmem_alloc(umma::smem_u32(tmem_slot), 512);
        umma::tmem_relinquish();
    }
    // The ring plan of this launch goes to shared memory: the constant bank is streamed through by the
    // epilogues (11.8 KB of biases per tile pair), so an indexed constant load in the MMA thread misses its
    // cache almost every time (measured: +390 cycles per chunk, the kernel 1.5x slower).
    static_assert(kRing >= 1 && kRing <= 3, "ring plans exist for one to three slots");
    static_assert(sizeof(RingPlan) <= 416, "plan region");
    {
#ifdef NERF_B200_EXPERIMENTS
        const RingPlan& src = (CFG::exp & 65536)   ? c_ring_plans_round1.p[kRing - 1][P.tile0 & 1]
                              : (CFG::exp & 32768) ? c_ring_plans_noreuse.p[kRing - 1][P.tile0 & 1]
                                                   : c_ring_plans.p[kRing - 1][P.tile0 & 1];
#else
        const RingPlan& src = c_ring_plans.p[kRing - 1][P.tile0 & 1];
#endif
        for (int i = threadIdx.x; i < (int)(sizeof(RingPlan) / 2); i += blockDim.x)
            reinterpret_cast<unsigned short*>(smem + kOffPlan)[i] = reinterpret_cast<const unsigned short*>(&src)[i];
    }
    const uint32_t plan_addr = sbase + kOffPlan;                  // 16-bit entries, kPlanPerPair per tile pair
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int plan_period = (int)reinterpret_cast<const RingPlan*>(smem + kOffPlan)->period;

    if (warp == 0) {
        // ===================== producer: weight slots, L2 -> smem =====================
        // Follows the ring plan (c_ring_plans): per layer it copies the chunks of the sub-tile that is issued
        // first, and for the second sub-tile only the chunks that are no longer resident.
        if (lane == 0 && !(CFG::exp & 8)) {
            uint32_t parity = 0;         // bit s: how often slot s was filled so far (mod 2)
            long pair_no = 0;
            int pp = 0;                  // pair_no mod the plan's period
            for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pair_no) {
                uint32_t q_addr = plan_addr + (uint32_t)pp * (kPlanPerPair * 2);
                for (int l = 0; l < kNumMmaLayers; ++l) {
                    // one slot = one K chunk with all its N halves (adjacent stages in the blob)
                    const int first = layer_first_stage(l), chunks = layer_chunks(l);
                    // EXP bit5 (timing): the same copies and hand-offs, but only 1 KB per slot
                    const uint32_t bytes = (CFG::exp & 32) ? 1024u : layer_halves(l) * kStageBytes;
                    for (int i = 0; i < 2 * chunks; ++i, q_addr += 2) {
                        const uint32_t e = umma::ld_shared_u16(q_addr);
                        if (!(e & kPlanFirstUse)) continue;          // the other sub-tile's copy is still resident
                        const int j = (int)(e & 7u), g = (e & kPlanG) ? 1 : 0;
                        const uint32_t slot = (e >> kPlanSlotShift) & kPlanSlotMask, ph = (parity >> slot) & 1u;
                        parity ^= 1u << slot;
                        long long t0 = PROBE ? clock64() : 0;
                        rec(0, pair_no, 1, l, g, j);                 // waits for a free slot
                        umma::mbar_wait(bar_w_empty + 8 * slot, ph ^ 1);
                        rec(0, pair_no, 2, l, g, j);                 // slot free: copy issued
                        if (PROBE) t_wait0 += clock64() - t0;
                        umma::mbar_arrive_expect_tx(bar_w_full + 8 * slot, bytes);
                        umma::bulk_g2s(sbase + kOffW + slot * kSlotBytes,
                                       P.blob + (size_t)first * kStageBytes + (size_t)j * bytes, bytes, bar_w_full + 8 * slot);
                    }
                }
                if (++pp == plan_period) pp = 0;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // Everything this one thread executes between tcgen05.mma instructions is tensor-pipe idle time
            // once the pipe's short queue has drained (four clock64 per chunk cost the probe build 20 %), so
            // the ring state is a few packed register words and the descriptors are built from precomputed words.
            // (The warp-uniform form of this loop that the CTA-pair kernel uses -- whole warp, elected issuer,
            // operands on the uniform datapath -- was tried here too: 4.62e7 cycles per CTA against 3.67e7.  At
            // 156 cycles per MMA this kernel is not bound by the issue instructions, and the extra divergence /
            // reconvergence per chunk lengthens every wait -> first-MMA hand-off.)
            uint32_t parity = 0;         // bit s: how often slot s was filled so far (mod 2)
            int pp = 0;                  // pair_no mod the plan's period
            uint32_t n_ready = 0;        // bit g: parity of sub-tile g's next a_ready phase
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << (46 - 32)) | (2u << (61 - 32));   // smem_desc_sw128, high word
            auto desc = [&](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
            auto desc_lo = [&](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
            const uint32_t w_lo = desc_lo(sbase + kOffW);
            const uint32_t a_lo0 = desc_lo(sbase + kOffA), pe_lo0 = desc_lo(sbase + kOffPE);
            long pair_no = 0;
            for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pair_no) {
                // EXP bit11: wait profile of the production layout, sampled on one tile pair in eight (two
                // clock64 around every wait would cost ~10 % if taken always)
                const bool sampled = (CFG::exp & 2048) && P.trace_out && (pair_no & 7) == 3;
                uint32_t q_addr = plan_addr + (uint32_t)pp * (kPlanPerPair * 2);
                uint32_t e = umma::ld_shared_u16(q_addr);
                // 18 segments per tile pair: (layer, sub-tile in issue order).  What is constant over a segment
                // (accumulator, instruction descriptor, A tile) stays out of the chunk loop: every operand that
                // changes between two tcgen05.mma costs the lone issuing thread a register -> uniform-register move.
#pragma unroll 1
                for (int seg = 0; seg < 2 * kNumMmaLayers; ++seg) {
                    const int l = seg >> 1;
                    const uint32_t g = (uint32_t)seg & 1u;
                    const uint32_t d_base = tmem_base + g * 256;
                    const uint32_t idesc = layer_halves(l) == 2 ? kIdescN256 : kIdescN128;
                    const uint32_t a_seg = a_lo0 + g * (65536 >> 4), pe_seg = pe_lo0 + g * (16384 >> 4);
                    {
                        // waits until sub-tile g's A operand is written
                        long long t0 = (PROBE || sampled) ? clock64() : 0;
                        rec(1, pair_no, 1, l, (int)g, 0);
                        umma::mbar_wait(bar_a_ready + 8 * g, (n_ready >> g) & 1u);
                        if (sampled) atomicAdd((unsigned long long*)P.trace_out + 64 + l * 2 + g, (unsigned long long)(clock64() - t0));
                        rec(1, pair_no, 2, l, (int)g, 0);
                        if (PROBE) t_wait0 += clock64() - t0;
                        n_ready ^= 1u << g;
                        umma::tc_fence_after();
                    }
                    uint32_t accumulate = 0u;
                    bool last;
#pragma unroll 1
                    do {
                        q_addr += 2;
                        // the next entry is fetched under this chunk's MMAs (the last fetch reads the pad word)
                        const uint32_t e_next = umma::ld_shared_u16(q_addr);
                        // everything the MMAs need is worked out BEFORE the waits: after a wait resolves, the
                        // tensor pipe is idle until the first tcgen05.mma is issued
                        const uint32_t slot = (e >> kPlanSlotShift) & kPlanSlotMask;
                        // B = [N rows][64] K-major; the two 128-row halves are contiguous
                        const uint32_t b_lo = w_lo + slot * (kSlotBytes >> 4);
                        const uint32_t a_lo = (!kPEA && (e & kPlanPeSrc)) ? pe_seg : a_seg + ((e >> 12) & 3u) * (16384 >> 4);
                        // one commit per chunk without a branch: the slot's release, or a barrier nobody waits on
                        const uint32_t release_bar = (e & kPlanRelease) ? bar_w_empty + 8 * slot : bar_unused;
                        const uint32_t ph_full = (parity >> slot) & 1u;
                        last = (e & kPlanLast) != 0;
                        if (kPEA && (e & kPlanPeChunk)) {
                            // PEA, l6: the PE block is restored into block 0 of the A tile once the MMAs that read
                            // h5's block 0 have completed; the PE chunk closes the layer
                            umma::mbar_wait(bar_a_ready + 8 * g, (n_ready >> g) & 1u);
                            n_ready ^= 1u << g;
                            umma::tc_fence_after();
                        }
                        if (e & kPlanFirstUse) {
                            parity ^= 1u << slot;
                            if (!(CFG::exp & 8)) {
                                // first use of a fill: wait for the copy (no tcgen05 fence here: the slot was written by
                                // the TMA engine, whose complete_tx on this barrier orders it before the MMAs that follow)
                                long long t1 = (PROBE || sampled) ? clock64() : 0;
                                rec(1, pair_no, 3, l, (int)g, (int)(e & 7u));
                                umma::mbar_wait(bar_w_full + 8 * slot, ph_full);
                                if (sampled) {
                                    atomicAdd((unsigned long long*)P.trace_out + l * 5 + (e & 7u), (unsigned long long)(clock64() - t1));
                                    if (seg == 0) atomicAdd((unsigned long long*)P.trace_out + 100, 1ull);
                                }
                                rec(1, pair_no, 4, l, (int)g, (int)(e & 7u));
                                if (PROBE) t_wait1 += clock64() - t1;
                            }
                        }
                        umma::mma_bf16_ss(d_base, desc(a_lo), desc(b_lo), idesc, accumulate);
#pragma unroll
                        for (int kk = 1; kk < 4; ++kk) umma::mma_bf16_ss(d_base, desc(a_lo + kk * 2), desc(b_lo + kk * 2), idesc, 1u);
                        accumulate = 1u;
                        umma::mma_commit(release_bar);
                        if (kPEA && (e & kPlanFreesBlock0)) umma::mma_commit(bar_pe_free + 8 * g);
                        rec(1, pair_no, 5, l, (int)g, (int)(e & 7u));
                        e = e_next;
                    } while (!last);
                    umma::mma_commit(bar_acc_full + 8 * g);
                }
                if (++pp == plan_period) pp = 0;
            }
        }
    } else if (WIDE) {
        // ===================== one 16-warp epilogue crew, alternating between the sub-tiles =====================
        const int cg = (warp - 2) >> 2;          // column group: columns [64 cg, 64 cg + 64) of a hidden layer
        const int quad = warp & 3;               // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        const uint32_t quad_bar = 1 + quad;      // named barrier of the four warps sharing rows
        const uint32_t swz = (uint32_t)(row & 7) << 4;
        float sigma[2] = {0.f, 0.f};
        uint32_t n_full[2] = {0, 0};
        // PE of sub-tile g of tile pair `pair` (two of the four threads of a row encode, all arrive)
        auto in_stage = [&](int g, long pair) {
            const long grow_raw = (pair * 2 + g) * kTileM + row;
            const long grow = grow_raw < P.M ? grow_raw : P.M - 1;
            uint8_t* pe_tile = smem + kOffPE + g * 16384;
            if (cg == 0) input_stage<0>(P, grow, pe_tile, row);
            else if (cg == 1) input_stage<1>(P, grow, pe_tile, row);
            umma::fence_proxy_async_smem();
            umma::mbar_arrive(bar_a_ready + 8 * g);
        };
        if ((long)blockIdx.x < n_pairs) { in_stage(0, blockIdx.x); in_stage(1, blockIdx.x); }
        for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll 1
            for (int l = 0; l < kNumMmaLayers; ++l) {
#pragma unroll 1
                for (int g = 0; g < 2; ++g) {
                    const uint32_t a_row_addr = sbase + kOffA + g * 65536 + row * 128;
                    const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + g * 256;
                    umma::mbar_wait_warp(bar_acc_full + 8 * g, n_full[g] & 1);
                    ++n_full[g];
                    umma::tc_fence_after();
                    if (l < kNumMmaLayers - 1) {
                        float sg = 0.f;
                        switch (cg) {
                            case 0: epilogue_hidden_ct<4, 0, 0>(l, tacc, 0, a_row_addr, swz, sg, P.ct); break;
                            case 1: epilogue_hidden_ct<4, 0, 64>(l, tacc, 64, a_row_addr, swz, sg, P.ct); break;
                            case 2: epilogue_hidden_ct<4, 0, 128>(l, tacc, 128, a_row_addr, swz, sg, P.ct); break;
                            default: epilogue_hidden_ct<4, 0, 192>(l, tacc, 192, a_row_addr, swz, sg, P.ct); break;
                        }
                        if (l == 7) sigma[g] = sg;
                        umma::fence_proxy_async_smem();
                        umma::tc_fence_before();
                        umma::mbar_arrive(bar_a_ready + 8 * g);
                    } else {
                        const long grow_raw = (pair * 2 + g) * kTileM + row;
                        const bool valid = grow_raw < P.M;
                        const long grow = valid ? grow_raw : P.M - 1;
                        float rgb[3];
                        const float* vt = P.vterm + (grow / P.vterm_div) * kL10Out;
                        epilogue_rgb<false, false, true, 2>(tacc, cg * 32, vt, nullptr, rgb, nullptr, a_row_addr, swz, P.ct);
                        umma::tc_fence_before();
                        // FP32 hand-over between the four threads of a row, inside the row's own PE
                        // line (free between l6's MMA and the next tile's encoding)
                        float4* xchg = reinterpret_cast<float4*>(smem + kOffPE + g * 16384 + row * 128);
                        if (cg > 0) xchg[cg - 1] = make_float4(rgb[0], rgb[1], rgb[2], sigma[g]);
                        umma::named_bar_sync(quad_bar, 128);
                        if (cg == 0 && valid) {
                            const float4 p1 = xchg[0], p2 = xchg[1], p3 = xchg[2];
                            float4 o;
                            o.x = rgb[0] + p1.x + p2.x + p3.x + P.ct.b11[0];
                            o.y = rgb[1] + p1.y + p2.y + p3.y + P.ct.b11[1];
                            o.z = rgb[2] + p1.z + p2.z + p3.z + P.ct.b11[2];
                            o.w = sigma[g] + p1.w + p2.w + p3.w + P.ct.balpha[0];
                            reinterpret_cast<float4*>(P.raw_out)[grow] = o;
                        }
                        // the next tile's encoding overwrites the hand-over slots: wait for the reads
                        umma::named_bar_sync(quad_bar, 128);
                        if (pair + gridDim.x < n_pairs) in_stage(g, pair + gridDim.x);
                    }
                }
            }
        }
    } else {
        // ===================== epilogue groups =====================
        const int ew = warp - 2;
        const int g = ew >> 3;                   // sub-tile / group
        const int half = (ew >> 2) & 1;          // which half of the columns this thread handles
        const int quad = warp & 3;               // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        const uint32_t pair_bar = 1 + g * 4 + quad;   // named barrier of the two warps sharing rows
        const uint32_t group_bar = 9 + g;              // named barrier of the group's 8 warps
        const int gtid = (ew & 7) * 32 + lane;         // thread index within the group
        const uint32_t bias_addr = sbase + kOffBias + g * 1024;
        const uint32_t a_row_addr = sbase + kOffA + g * 65536 + row * 128;
        const uint32_t swz = (uint32_t)(row & 7) << 4;
        // PEA: the PE tile is block 0 of the group's A tile
        uint8_t* pe_tile = kPEA ? smem + kOffA + g * 65536 : smem + kOffPE + g * 16384;
        // FP32 hand-over slot between the two threads of a row, inside the row's own PE line (free between
        // l6's MMA and the next tile's encoding); PEA: the row's line in block 3 of the A tile (h9, dead after
        // l10's MMAs; the training build parks h10 in blocks 0..1 at that point)
        float4* xchg = reinterpret_cast<float4*>((kPEA ? smem + kOffA + g * 65536 + 3 * 16384 : pe_tile) + row * 128);
        uint32_t n_pe_free = 0;
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + g * 256;
        const float* tail = reinterpret_cast<const float*>(P.blob + kWeightBytes);
        uint32_t n_full = 0;
        if (kStageBias) {
            umma::st_shared_f32(bias_addr + gtid * 4, __ldg(tail + kTailBias + gtid));
            umma::named_bar_sync(group_bar, kEpiWarpsPerGroup * 32);
        }
        long pair_no = 0;
        const bool tracer = (ew & 7) == 0 && lane == 0;        // first warp of the group
        uint4 pe_regs[4];
        bool pe_ahead = false, pe_stored = false;
        const float* vt_row = nullptr;
        long long t_rgb = 0;      // variant 18: start of the sampled tile's rgb epilogue
        for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pair_no) {
            const long grow_raw = (pair * 2 + g) * kTileM + row;
            const bool valid = grow_raw < P.M;
            const long grow = valid ? grow_raw : P.M - 1;
            if (SAVE && kPEA) {
                // the previous tile's h10 copy reads blocks 0..1 of the A tile: it must be done before the
                // PE is encoded into block 0
                if (gtid == 0) umma::bulk_wait_read0();
                umma::named_bar_sync(group_bar, kEpiWarpsPerGroup * 32);
            }
            // the tile's PE: encoded ahead of time (into registers, while the group waited for the previous
            // tile's l10 accumulator) whenever there was a previous tile -- at a tile boundary only the stores
            // are left on the sub-tile's MMA -> epilogue -> MMA chain
            if (!pe_ahead) {
                if (half == 0) input_encode<0>(P, grow, pe_regs);
                else input_encode<1>(P, grow, pe_regs);
            }
            if (!pe_stored) {      // (stored at the end of the previous tile otherwise)
                if (half == 0) input_store<0>(pe_regs, pe_tile, row);
                else input_store<1>(pe_regs, pe_tile, row);
                umma::fence_proxy_async_smem();
                umma::mbar_arrive(bar_a_ready + 8 * g);
            }
            pe_ahead = false;
            pe_stored = false;
            if ((CFG::exp & 2048) && t_rgb) atomicAdd((unsigned long long*)P.trace_out + 101 + g, (unsigned long long)(clock64() - t_rgb));
            t_rgb = 0;
            const bool esampled = (CFG::exp & 2048) && P.trace_out && tracer && (pair_no & 7) == 3;
            uint8_t* act_tile = nullptr;
            uint8_t* mask_tile = nullptr;          // every thread stores its own mask words
            if (SAVE) {
                const long tile = pair * 2 + g;
                const bool saver = gtid == 0 && tile < n_tiles && !(CFG::exp & 16);   // EXP bit4: no record copies (timing)
                act_tile = saver ? P.act_save + (size_t)tile * kActTileBytes : nullptr;
                mask_tile = tile < n_tiles ? P.act_save + (size_t)tile * kActTileBytes : nullptr;
                // the previous tile's h10 copy must have left the A tile before l1's epilogue
                // rewrites it; every thread passes the barrier after thread 0 has seen that
                if (gtid == 0) umma::bulk_wait_read0();
                umma::named_bar_sync(group_bar, kEpiWarpsPerGroup * 32);
                if (act_tile) {
                    umma::bulk_s2g(act_tile + kActPE, kPEA ? sbase + kOffA + g * 65536 : sbase + kOffPE + g * 16384, 16384);
                    umma::bulk_commit();
                }
            }
            float sigma = 0.f;
            float* probe_row = nullptr;
            uint32_t mw[4] = {0, 0, 0, 0};
#pragma unroll 1
            for (int l = 0; l < kNumMmaLayers; ++l) {
                long long t0 = PROBE ? clock64() : 0;
                if (!SAVE && l == kNumMmaLayers - 1 && pair + gridDim.x < n_pairs) {
                    const long nxt = ((pair + gridDim.x) * 2 + g) * kTileM + row;
                    if (half == 0) input_encode<0>(P, nxt < P.M ? nxt : P.M - 1, pe_regs);
                    else input_encode<1>(P, nxt < P.M ? nxt : P.M - 1, pe_regs);
                    pe_ahead = true;
                }
                if (l == kNumMmaLayers - 1) {
                    // the row's view term (l10's hoisted view-direction columns), 64 floats per thread: brought
                    // into L1 while l10's MMAs run, its L2 latency would otherwise sit on the tile boundary's
                    // MMA -> epilogue -> MMA chain four times
                    vt_row = P.vterm + (grow / P.vterm_div) * kL10Out;
                    umma::prefetch_l1(vt_row + half * 64);
                    umma::prefetch_l1(vt_row + half * 64 + 32);
                }
                if (tracer) rec(2 + g, pair_no, 1, l, g, 0);           // waits for the accumulator
                umma::mbar_wait_warp(bar_acc_full + 8 * g, n_full & 1);
                if (tracer) rec(2 + g, pair_no, 2, l, g, 0);           // accumulator complete
                if (PROBE) t_wait0 += clock64() - t0;
                ++n_full;
                umma::tc_fence_after();
                const long long t_epi = ((CFG::exp & 2048) && esampled) ? clock64() : 0;
                if ((CFG::exp & 2048) && esampled && l == kNumMmaLayers - 1) t_rgb = t_epi;
                if (PROBE) probe_row = (P.probe_out && P.probe_layer == l && valid) ? P.probe_out + grow * 256 : nullptr;
                if (SAVE && kPEA && l == 0) {
                    // the PE record copy reads block 0, which l1's epilogue is about to overwrite
                    if (gtid == 0) umma::bulk_wait_read0();
                    umma::named_bar_sync(group_bar, kEpiWarpsPerGroup * 32);
                }
                if (l < kNumMmaLayers - 1) {
                    const int c0 = half * 128;
                    if (CT && !PROBE) {
                        if (half == 0) epilogue_hidden_ct<8, CFG::exp & (7 | 64 | 8192 | 16384), 0>(l, tacc, 0, a_row_addr, swz, sigma, P.ct);
                        else epilogue_hidden_ct<8, CFG::exp & (7 | 64 | 8192 | 16384), 128>(l, tacc, 128, a_row_addr, swz, sigma, P.ct);
                    } else if (l == 7) {       // l8: its FP32 accumulator also feeds the sigma head (l_alpha)
                        epilogue_hidden<1, PROBE, CFG::exp, CT, 8, -1, SAVE>(tacc, c0, a_row_addr, swz, bias_addr, tail + kTailWAlpha, sigma, probe_row, P.ct, l, mw);
                    } else {
                        epilogue_hidden<0, PROBE, CFG::exp, CT, 8, -1, SAVE>(tacc, c0, a_row_addr, swz, bias_addr, nullptr, sigma, probe_row, P.ct, l, mw);
                    }
                    if (SAVE && mask_tile)
                        *reinterpret_cast<uint4*>(mask_tile + act_mask_slot(l, row, half)) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
                    umma::fence_proxy_async_smem();
                    umma::tc_fence_before();
                    umma::mbar_arrive(bar_a_ready + 8 * g);
                    if ((CFG::exp & 2048) && esampled) atomicAdd((unsigned long long*)P.trace_out + 104 + l * 2 + g, (unsigned long long)(clock64() - t_epi));
                    if (tracer) rec(2 + g, pair_no, 3, l, g, 0);       // this warp's part of the A operand written
                    // stage the next hidden layer's bias (after l9 comes l1 of the next tile); both
                    // barriers fall into the time the group would wait for the tensor core anyway
                    float bnext = 0.f;
                    if (kStageBias) bnext = __ldg(tail + kTailBias + (l == 7 ? 0 : l + 1) * kHidden + gtid);
                    if (kGroupSync) umma::named_bar_sync(group_bar, kEpiWarpsPerGroup * 32);
                    if (SAVE && gtid == 0) {
                        // activation record: h_{l+1} straight from the A tile;
                        // the copy must have read the tile before the next epilogue overwrites it
                        if (act_tile) {
                            umma::bulk_s2g(act_tile + act_hidden(l + 1), sbase + kOffA + g * 65536, 65536);
                            umma::bulk_commit();
                        }
                        umma::bulk_wait_read0();
                    }
                    if (kStageBias) umma::st_shared_f32(bias_addr + gtid * 4, bnext);
                    if (kGroupSync) umma::named_bar_sync(group_bar, kEpiWarpsPerGroup * 32);
                    if (kPEA && l == 4) {
                        // l6's MMAs on block 0 (h5) have completed: put the tile's PE (still in pe_regs -- the
                        // compiler parks them in local memory across l2..l5) back into it
                        umma::mbar_wait_warp(bar_pe_free + 8 * g, n_pe_free & 1);
                        ++n_pe_free;
                        if (half == 0) input_store<0>(pe_regs, pe_tile, row);
                        else input_store<1>(pe_regs, pe_tile, row);
                        umma::fence_proxy_async_smem();
                        umma::mbar_arrive(bar_a_ready + 8 * g);
                    }
                } else {
                    float rgb[3];
                    if (CT && !PROBE && !SAVE) {
                        if (half == 0) epilogue_rgb<false, false, true, 4, 0>(tacc, 0, vt_row, nullptr, rgb, nullptr, a_row_addr, swz, P.ct, mw);
                        else epilogue_rgb<false, false, true, 4, 64>(tacc, 64, vt_row, nullptr, rgb, nullptr, a_row_addr, swz, P.ct, mw);
                    } else {
                        epilogue_rgb<PROBE, SAVE, CT>(tacc, half * 64, vt_row, tail + kTailW11, rgb, probe_row, a_row_addr, swz, P.ct, mw);
                    }
                    umma::tc_fence_before();
                    if (kPEA && !SAVE && pe_ahead) {
                        // the accumulator is read and l10's MMAs have left the A tile: hand the next tile's PE
                        // (block 0) to the tensor core before this tile's output is combined and written
                        if (half == 0) input_store<0>(pe_regs, pe_tile, row);
                        else input_store<1>(pe_regs, pe_tile, row);
                        umma::fence_proxy_async_smem();
                        umma::mbar_arrive(bar_a_ready + 8 * g);
                        pe_stored = true;
                        if ((CFG::exp & 2048) && t_rgb) atomicAdd((unsigned long long*)P.trace_out + 101 + g, (unsigned long long)(clock64() - t_rgb));
                        t_rgb = 0;
                    }
                    if (SAVE && mask_tile)
                        *reinterpret_cast<uint2*>(mask_tile + act_mask_slot(8, row, half)) = make_uint2(mw[0], mw[1]);
                    if (SAVE) {
                        umma::fence_proxy_async_smem();
                        umma::named_bar_sync(group_bar, kEpiWarpsPerGroup * 32);
                        if (act_tile) {
                            umma::bulk_s2g(act_tile + kActH10, sbase + kOffA + g * 65536, 32768);
                            umma::bulk_commit();
                        }
                    }
                    // combine the two column halves of the row: half 1 hands its partial sums over
                    if (half == 1) *xchg = make_float4(rgb[0], rgb[1], rgb[2], sigma);
                    umma::named_bar_sync(pair_bar, 64);
                    if (half == 0) {
                        const float4 o2 = *xchg;
                        if (valid) {
                            float4 o;
                            o.x = rgb[0] + o2.x + (CT ? P.ct.b11[0] : __ldg(tail + kTailB11 + 0));
                            o.y = rgb[1] + o2.y + (CT ? P.ct.b11[1] : __ldg(tail + kTailB11 + 1));
                            o.z = rgb[2] + o2.z + (CT ? P.ct.b11[2] : __ldg(tail + kTailB11 + 2));
                            o.w = sigma + o2.w + (CT ? P.ct.balpha[0] : __ldg(tail + kTailBAlpha));
                            reinterpret_cast<float4*>(P.raw_out)[grow] = o;
                        }
                    }
                    // the hand-over slot is rewritten by the next tile's encoding (PEA: by half 1's l1 epilogue, block
                    // 3): wait for the read.  With the early PE hand-over above this barrier is off the tile
                    // boundary's MMA -> epilogue -> MMA chain.
                    umma::named_bar_sync(pair_bar, 64);
                }
            }
        }
        if (SAVE && gtid == 0) umma::bulk_wait_all();
    }
    if (PROBE && P.stats_out && lane == 0) {
        // [0] producer wait-empty, [1] mma wait a_ready, [2] mma wait w_full, [3] epi X wait acc,
        // [4] epi Y wait acc, [5] total cycles
        long long* o = P.stats_out + (long)blockIdx.x * 8;
        const long long total = clock64() - t_begin;
        if (warp == 0) { o[0] = t_wait0; o[5] = total; }
        if (warp == 1) { o[1] = t_wait0; o[2] = t_wait1; }
        if (warp == 2) o[3] = t_wait0;
        if (warp == 2 + kEpiWarpsPerGroup) o[4] = t_wait0;
    }
    umma::tc_fence_before();
    __syncthreads();
    // every variant reports its cycle count when the debug entry passes a counter block: under the power
    // cap a variant can be faster in milliseconds (higher clock) without being faster in cycles
    if (!PROBE && P.stats_out && threadIdx.x == 0) P.stats_out[(long)blockIdx.x * 8 + 5] = clock64() - t_begin;
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, 512);
    }
}

template <bool PROBE, class CFG, bool SAVE = false, bool CT = false, bool WIDE = false>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fwd_kernel(const __grid_constant__ FwdParams P) {
    mlp_fwd_body<PROBE, CFG, SAVE, CT, WIDE>(P);
}

// The same body with a 112-register budget (576 threads x 112 = 64 512 of the SM's 65 536 registers;
// with __launch_bounds__(576, 1) ptxas stops at 96): room for the 2 x 32-register accumulator buffers
// of the 32-column epilogue.
template <bool PROBE, class CFG, bool SAVE = false, bool CT = false, bool WIDE = false>
__global__ void __maxnreg__(112) mlp_fwd_kernel_r112(const __grid_constant__ FwdParams P) {
    mlp_fwd_body<PROBE, CFG, SAVE, CT, WIDE>(P);
}

// ---------------------------------------------------------------------------- CTA-pair kernel (inference)
// The same network on tcgen05 cta_group::2: the two CTAs of a cluster (one TPC) run ONE M=256 x N=256 MMA per
// K step over their two 128-row sub-tiles, and each CTA stages only ITS half of every weight chunk (128 of the
// 256 output rows; 64 of l10's 128).  The SS-form MMA of a single CTA fetches 12 KB of operands from shared
// memory per step and takes 156 cycles for it; the pair MMA fetches 8 KB per SM and runs at the tensor core's
// 128 cycles (tools/probes/pair_mma_rate_probe.cu), and the bytes the copy engine writes into shared memory
// halve as well.  Round 1 had this design and lost 13 % with it: a sub-tile's epilogue took ~3500 cycles then
// and, with the hand-offs crossing the pair, did not fit under the other sub-tile's 2048 MMA cycles.  The
// epilogue now takes ~1000.
//   tiles      a cluster takes four 128-row tiles at a time: CTA r, group g -> tile 4q + 2r + g, so g is the
//              parity of the global tile and the chunk orders (chunk_at) -- hence every row's summation order
//              and the output bits -- are those of mlp_fwd_kernel
//   warp 0     (both CTAs) producer of the CTA's own 16 KB half-chunks: ring of five, plan c_pair_plans
//   warp 1     leader: MMA issuer (tcgen05.mma.cta_group::2; commits multicast to both CTAs);
//              peer: relays "my half-chunk has landed" to the leader's slot barrier (count 2 there: own copy +
//              relay, so the issuing warp waits once per fill)
//   warps 2-17 the two epilogue groups of mlp_fwd_kernel (PE in the A tile, host tail); a warp signals "my part
//              of the A tile is written" with one arrive on the LEADER's barrier (remote from the peer)
constexpr uint32_t kPairOffW = kOffPE;                 // no PE tiles: 96 KB for the ring
static_assert(kPairOffW + kPairRing * kStageBytes <= kOffBar, "pair ring does not fit");
constexpr uint32_t kIdescPairN256 = umma::instr_desc_bf16(256, 256);
constexpr uint32_t kIdescPairN128 = umma::instr_desc_bf16(256, 128);

template <int EXP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_fwd_pair_kernel(const __grid_constant__ FwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    if ((sbase & 1023u) != 0) __trap();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = umma::cluster_ctarank();
    constexpr int R = kPairRing;
    const uint32_t bar_w_full = sbase + kOffBar;              // [R] this CTA's half-chunk has landed
    const uint32_t bar_w_empty = bar_w_full + 8 * R;          // [R] the pair's MMAs are done with the slot (multicast commit)
    const uint32_t bar_w_peer = bar_w_empty + 8 * R;          // [R] leader only: the peer's half-chunk has landed
    const uint32_t bar_a_ready = bar_w_peer + 8 * R;          // [2] leader only: both CTAs' A tiles written
    const uint32_t bar_acc_full = bar_a_ready + 16;           // [2] accumulators complete (multicast commit)
    const uint32_t bar_pe_free = bar_acc_full + 16;           // [2] l6's MMAs on A block 0 (h5) completed (multicast commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (3 * R + 6));
    static_assert(8 * (3 * R + 6) + 4 <= 256, "barrier region");

    const long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const long n_tiles = (P.M + kTileM - 1) / kTileM;
    const long n_quads = (n_tiles + 3) / 4;
    const long long t_begin = P.stats_out ? clock64() : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < R; ++s) {
            // leader: the slot is full when its own copy has landed (producer's arrive + bytes) AND the peer's
            // relay has arrived: one barrier, one wait per fill in the issuing warp
            umma::mbar_init(bar_w_full + 8 * s, (rank == 0 && !(EXP & (1 << 18))) ? 2 : 1);
            umma::mbar_init(bar_w_empty + 8 * s, 1);
            umma::mbar_init(bar_w_peer + 8 * s, 1);
        }
        for (int g = 0; g < 2; ++g) {
            umma::mbar_init(bar_a_ready + 8 * g, 2 * kEpiWarpsPerGroup);   // one arrive per warp and CTA
            umma::mbar_init(bar_acc_full + 8 * g, 1);
            umma::mbar_init(bar_pe_free + 8 * g, 1);
        }
        umma::fence_barrier_init();
    }
    if (warp == 1) {
        umma::tmem_alloc_pair(umma::smem_u32(tmem_slot), 512);
        umma::tmem_relinquish_pair();
    }
    // the plan goes where the device-tail kernels stage their biases (unused here: the tail is in the parameters)
    constexpr uint32_t kOffPairPlan = kOffBias;
    static_assert(sizeof(PairPlan) <= 2 * 1024, "pair plan region");
    {
        const PairPlan& src = c_pair_plans.p[P.tile0 & 1];
        for (int i = threadIdx.x; i < (int)(sizeof(PairPlan) / 2); i += blockDim.x)
            reinterpret_cast<unsigned short*>(smem + kOffPairPlan)[i] = reinterpret_cast<const unsigned short*>(&src)[i];
    }
    const uint32_t plan_addr = sbase + kOffPairPlan;
    umma::tc_fence_before();
    umma::cluster_sync_all();          // barriers of BOTH CTAs initialised before any remote arrive
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int plan_period = (int)reinterpret_cast<const PairPlan*>(smem + kOffPairPlan)->period;

    if (warp == 0) {
        // ===================== producer: this CTA's half of every weight chunk =====================
        if (lane == 0 && !(EXP & 8)) {
            uint32_t parity = 0;         // bit s: how often slot s was filled so far (mod 2)
            int pp = 0;
            for (long quad = cluster_id; quad < n_quads; quad += n_clusters) {
                uint32_t q_addr = plan_addr + (uint32_t)pp * (kPlanPerPair * 2);
                for (int l = 0; l < kNumMmaLayers; ++l) {
                    const int first = layer_first_stage(l), chunks = layer_chunks(l);
                    // N = 256: stage (chunk, half = rank); N = 128 (l10): rows [64 rank, 64 rank + 64) of the chunk's stage
                    const bool n256 = layer_halves(l) == 2;
                    const uint32_t bytes = n256 ? kStageBytes : kStageBytes / 2;
                    const uint8_t* src = P.blob + (size_t)first * kStageBytes + (size_t)rank * bytes;
                    const size_t stride = n256 ? 2 * kStageBytes : kStageBytes;
                    for (int i = 0; i < 2 * chunks; ++i, q_addr += 2) {
                        const uint32_t e = umma::ld_shared_u16(q_addr);
                        if (!(e & kPlanFirstUse)) continue;          // the other sub-tile's copy is still resident
                        const uint32_t slot = (e >> kPlanSlotShift) & kPlanSlotMask, ph = (parity >> slot) & 1u;
                        parity ^= 1u << slot;
                        umma::mbar_wait(bar_w_empty + 8 * slot, ph ^ 1);
                        umma::mbar_arrive_expect_tx(bar_w_full + 8 * slot, bytes);
                        umma::bulk_g2s(sbase + kPairOffW + slot * kStageBytes, src + (size_t)(e & 7u) * stride, bytes,
                                       bar_w_full + 8 * slot);
                    }
                }
                if (++pp == plan_period) pp = 0;
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 1 && (EXP & 8)) {
            // timing variant without weight streaming: nothing to relay
        } else if (lane == 0 && rank == 1) {
            // ===================== peer: relay slot arrivals to the leader =====================
            uint32_t parity = 0;
            int pp = 0;
            const uint32_t leader_w_peer = umma::map_to_cta((EXP & (1 << 18)) ? bar_w_peer : bar_w_full, 0);
            for (long quad = cluster_id; quad < n_quads; quad += n_clusters) {
                uint32_t q_addr = plan_addr + (uint32_t)pp * (kPlanPerPair * 2);
                for (int i = 0; i < kPlanPerPair; ++i, q_addr += 2) {
                    const uint32_t e = umma::ld_shared_u16(q_addr);
                    if (!(e & kPlanFirstUse)) continue;
                    const uint32_t slot = (e >> kPlanSlotShift) & kPlanSlotMask, ph = (parity >> slot) & 1u;
                    parity ^= 1u << slot;
                    umma::mbar_wait(bar_w_full + 8 * slot, ph);
                    umma::mbar_arrive_remote_cta(leader_w_peer + 8 * slot);
                }
                if (++pp == plan_period) pp = 0;
            }
        } else if (rank == 0) {
            // ===================== leader: MMA issuer for the pair =====================
            // The WHOLE warp runs this loop with warp-uniform values, and one elected lane issues the tensor-core
            // instructions.  With a lone thread (if (lane == 0) around the loop) the compiler has to move every
            // operand of every tcgen05.mma from vector registers into uniform registers inside an
            // elect / R2UR.BROADCAST x7 / branch loop (~17 instructions per MMA): more than the 128 cycles a pair
            // MMA leaves, so the issue thread, not the tensor pipe, set the pace (~200 cycles per MMA).  Values
            // the compiler can prove uniform (the plan entry goes through a warp reduction: REDUX writes a
            // uniform register) stay on the uniform datapath: UIADD + UTCHMMA.
            const bool issuer = umma::elect_one();
            uint32_t parity = 0, n_ready = 0;
            int pp = 0;
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << (46 - 32)) | (2u << (61 - 32));   // smem_desc_sw128, high word
            auto desc_lo = [&](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
            auto uniform = [&](uint32_t v) { return __reduce_max_sync(0xffffffffu, v); };
            const uint32_t w_lo = uniform(desc_lo(sbase + kPairOffW));
            const uint32_t a_lo0 = uniform(desc_lo(sbase + kOffA));
            const uint32_t tmem_u = uniform(tmem_base);
            const uint32_t bars_u = uniform(bar_w_full);
            // all lanes poll together (a uniform loop: no divergence to reconverge from behind a wait; the
            // elected-lane form "if (issuer) wait; __syncwarp()" costs ~140 cycles per chunk)
            auto wait = [&](uint32_t bar, uint32_t ph) { umma::mbar_wait_warp(bar, ph); };
            long quad_no = 0;
            for (long quad = cluster_id; quad < n_quads; quad += n_clusters, ++quad_no) {
                const bool sampled = (EXP & 2048) && P.trace_out && (quad_no & 7) == 3;
                uint32_t q_addr = plan_addr + (uint32_t)pp * (kPlanPerPair * 2);
                uint32_t e = uniform(umma::ld_shared_u16(q_addr));
#pragma unroll 1
                for (int seg = 0; seg < 2 * kNumMmaLayers; ++seg) {
                    const int l = seg >> 1;
                    const uint32_t g = (uint32_t)seg & 1u;
                    const uint32_t d_base = tmem_u + g * 256;
                    const uint32_t a_seg = a_lo0 + g * (65536 >> 4);
                    const uint32_t u_a_ready = bars_u + 8 * (3 * R) + 8 * g, u_acc_full = u_a_ready + 16, u_pe_free = u_a_ready + 32;
                    const uint32_t idesc = layer_halves(l) == 2 ? kIdescPairN256 : kIdescPairN128;
                    // waits until sub-tile g's A operand is written (and its accumulator read)
                    auto wait_a = [&]() {
                        const long long t0 = sampled ? clock64() : 0;
                        wait(u_a_ready, (n_ready >> g) & 1u);
                        if (sampled && issuer) atomicAdd((unsigned long long*)P.trace_out + 64 + l * 2 + g, (unsigned long long)(clock64() - t0));
                        n_ready ^= 1u << g;
                        umma::tc_fence_after();
                    };
                    wait_a();
                    uint32_t accumulate = 0u;
                    bool last;
#pragma unroll 1
                    do {
                        q_addr += 2;
                        // the next entry is fetched under this chunk's MMAs (the last fetch reads the pad word)
                        const uint32_t e_next = uniform(umma::ld_shared_u16(q_addr));
                        const uint32_t slot = (e >> kPlanSlotShift) & kPlanSlotMask;
                        const uint32_t b_lo = w_lo + slot * (kStageBytes >> 4);
                        const uint32_t a_lo = a_seg + ((e >> 12) & 3u) * (16384 >> 4);
                        const uint32_t release_bar = ((e & kPlanRelease) && !(EXP & 16)) ? bars_u + 8 * R + 8 * slot : 0u;   // bar_w_empty[slot]
                        last = (e & kPlanLast) != 0;
                        if (e & kPlanPeChunk) wait_a();      // l6: both CTAs have restored the PE block into block 0 of their A tiles
                        if (e & kPlanFirstUse) {
                            const uint32_t ph_full = (parity >> slot) & 1u;
                            parity ^= 1u << slot;
                            if (!(EXP & 8)) {
                                const long long t1 = sampled ? clock64() : 0;
                                wait(bars_u + 8 * slot, ph_full);                  // bar_w_full[slot]: own half landed + the peer's relay
                                if (EXP & (1 << 18)) wait(bars_u + 8 * (2 * R) + 8 * slot, ph_full);    // (A/B: the relay on a barrier of its own)
                                if (sampled && issuer) {
                                    atomicAdd((unsigned long long*)P.trace_out + l * 5 + (e & 7u), (unsigned long long)(clock64() - t1));
                                    if (seg == 0) atomicAdd((unsigned long long*)P.trace_out + 100, 1ull);
                                }
                            }
                        }
                        // four MMAs, the slot's release and, after the chunk that frees A block 0 in l6, bar_pe_free[g]
                        umma::mma_chunk_pair_elect(d_base, a_lo, b_lo, kDescHi, idesc, accumulate, release_bar,
                                                   (e & kPlanFreesBlock0) ? u_pe_free : 0u);
                        accumulate = 1u;
                        e = e_next;
                    } while (!last);
                    umma::mma_commit_pair_elect(u_acc_full);
                }
                if (++pp == plan_period) pp = 0;
            }
        }
    } else {
        // ===================== epilogue groups (as in mlp_fwd_kernel: PE in the A tile, host tail) =====================
        const int ew = warp - 2;
        const int g = ew >> 3;
        const int half = (ew >> 2) & 1;
        const int quadrant = warp & 3;
        const int row = quadrant * 32 + lane;
        const uint32_t pair_bar = 1 + g * 4 + quadrant;
        const uint32_t a_row_addr = sbase + kOffA + g * 65536 + row * 128;
        const uint32_t swz = (uint32_t)(row & 7) << 4;
        uint8_t* pe_tile = smem + kOffA + g * 65536;
        float4* xchg = reinterpret_cast<float4*>(smem + kOffA + g * 65536 + 3 * 16384 + row * 128);
        const uint32_t tacc = tmem_base + ((uint32_t)(quadrant * 32) << 16) + g * 256;
        const uint32_t leader_a_ready = umma::map_to_cta(bar_a_ready + 8 * g, 0);
        // "this warp's part of the group's A tile is written": every lane has fenced its stores towards the
        // async proxy and its TMEM loads towards the tensor core; one lane tells the leader's MMA warp.  The
        // arrive has CTA-scope release semantics, as in CUTLASS's ClusterBarrier::arrive(cta_id): the
        // .release.cluster form is a cluster-wide fence for the thread and cost ~2000 cycles per hand-over here.
        auto signal_a_ready = [&]() {
            umma::fence_proxy_async_smem();
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive_remote_cta(leader_a_ready);
        };
        const bool tracer = (EXP & 2048) && (ew & 7) == 0 && lane == 0;
        uint32_t n_full = 0, n_pe_free = 0;
        uint4 pe_regs[4];
        bool pe_ahead = false, pe_stored = false;
        const float* vt_row = nullptr;
        long long t_rgb = 0;
        long quad_no = 0;
        for (long quad = cluster_id; quad < n_quads; quad += n_clusters, ++quad_no) {
            const long grow_raw = (quad * 4 + rank * 2 + g) * kTileM + row;
            const bool valid = grow_raw < P.M;
            const long grow = valid ? grow_raw : P.M - 1;
            if (!pe_ahead) {
                if (half == 0) input_encode<0>(P, grow, pe_regs);
                else input_encode<1>(P, grow, pe_regs);
            }
            if (!pe_stored) {      // (stored at the end of the previous tile otherwise)
                if (half == 0) input_store<0>(pe_regs, pe_tile, row);
                else input_store<1>(pe_regs, pe_tile, row);
                signal_a_ready();
            }
            pe_ahead = false;
            pe_stored = false;
            const bool esampled = tracer && P.trace_out && (quad_no & 7) == 3;
            float sigma = 0.f;
#pragma unroll 1
            for (int l = 0; l < kNumMmaLayers; ++l) {
                if (l == kNumMmaLayers - 1) {
                    if (quad + n_clusters < n_quads) {
                        const long nxt = ((quad + n_clusters) * 4 + rank * 2 + g) * kTileM + row;
                        if (half == 0) input_encode<0>(P, nxt < P.M ? nxt : P.M - 1, pe_regs);
                        else input_encode<1>(P, nxt < P.M ? nxt : P.M - 1, pe_regs);
                        pe_ahead = true;
                    }
                    vt_row = P.vterm + (grow / P.vterm_div) * kL10Out;
                    umma::prefetch_l1(vt_row + half * 64);
                    umma::prefetch_l1(vt_row + half * 64 + 32);
                }
                umma::mbar_wait_warp(bar_acc_full + 8 * g, n_full & 1);
                ++n_full;
                umma::tc_fence_after();
                const long long t_epi = esampled ? clock64() : 0;
                if (esampled && l == kNumMmaLayers - 1) t_rgb = t_epi;
                if (l < kNumMmaLayers - 1) {
                    if (half == 0) epilogue_hidden_ct<8, EXP & (7 | 64 | 16384), 0>(l, tacc, 0, a_row_addr, swz, sigma, P.ct);
                    else epilogue_hidden_ct<8, EXP & (7 | 64 | 16384), 128>(l, tacc, 128, a_row_addr, swz, sigma, P.ct);
                    if ((EXP & 2048) && esampled) {
                        // split of the hand-over: epilogue body / proxy + tcgen05 fences / warp sync + remote arrive
                        const long long t1 = clock64();
                        umma::fence_proxy_async_smem();
                        umma::tc_fence_before();
                        const long long t2 = clock64();
                        __syncwarp();
                        if (lane == 0) umma::mbar_arrive_remote_cta(leader_a_ready);
                        const long long t3 = clock64();
                        atomicAdd((unsigned long long*)P.trace_out + 130, (unsigned long long)(t1 - t_epi));
                        atomicAdd((unsigned long long*)P.trace_out + 131, (unsigned long long)(t2 - t1));
                        atomicAdd((unsigned long long*)P.trace_out + 132, (unsigned long long)(t3 - t2));
                    } else {
                        signal_a_ready();
                    }
                    if (esampled) atomicAdd((unsigned long long*)P.trace_out + 104 + l * 2 + g, (unsigned long long)(clock64() - t_epi));
                    if (l == 4) {
                        // l6's MMAs on block 0 (h5) have completed in both CTAs: put the tile's PE back into it
                        umma::mbar_wait_warp(bar_pe_free + 8 * g, n_pe_free & 1);
                        ++n_pe_free;
                        if (half == 0) input_store<0>(pe_regs, pe_tile, row);
                        else input_store<1>(pe_regs, pe_tile, row);
                        signal_a_ready();
                    }
                } else {
                    float rgb[3];
                    if (half == 0) epilogue_rgb<false, false, true, 4, 0>(tacc, 0, vt_row, nullptr, rgb, nullptr, a_row_addr, swz, P.ct);
                    else epilogue_rgb<false, false, true, 4, 64>(tacc, 64, vt_row, nullptr, rgb, nullptr, a_row_addr, swz, P.ct);
                    if (pe_ahead) {
                        // the accumulator is read and l10's MMAs have left the A tile: hand the next tile's PE
                        // (block 0) to the tensor core before this tile's output is combined and written
                        if (half == 0) input_store<0>(pe_regs, pe_tile, row);
                        else input_store<1>(pe_regs, pe_tile, row);
                        signal_a_ready();
                        pe_stored = true;
                        if (t_rgb) atomicAdd((unsigned long long*)P.trace_out + 101 + g, (unsigned long long)(clock64() - t_rgb));
                        t_rgb = 0;
                    } else {
                        umma::tc_fence_before();
                    }
                    // combine the two column halves of the row: half 1 hands its partial sums over (slot in block 3 of
                    // the A tile, dead after l10's MMAs)
                    if (half == 1) *xchg = make_float4(rgb[0], rgb[1], rgb[2], sigma);
                    umma::named_bar_sync(pair_bar, 64);
                    if (half == 0) {
                        const float4 o2 = *xchg;
                        if (valid) {
                            float4 o;
                            o.x = rgb[0] + o2.x + P.ct.b11[0];
                            o.y = rgb[1] + o2.y + P.ct.b11[1];
                            o.z = rgb[2] + o2.z + P.ct.b11[2];
                            o.w = sigma + o2.w + P.ct.balpha[0];
                            reinterpret_cast<float4*>(P.raw_out)[grow] = o;
                        }
                    }
                    umma::named_bar_sync(pair_bar, 64);     // half 1's next l1 epilogue rewrites the slot
                }
            }
        }
    }
    umma::tc_fence_before();
    umma::cluster_sync_all();          // no CTA frees tensor memory while its partner still uses it
    if (P.stats_out && threadIdx.x == 0) P.stats_out[(long)blockIdx.x * 8 + 5] = clock64() - t_begin;
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc_pair(tmem_base, 512);
    }
}

using FwdKernel = void (*)(const FwdParams);   // (__grid_constant__ does not change the type)

// variant 0 = production; 1 = probe (production config + debug outputs);
// 2, 3 = pipeline-timing experiments (probe kernels with other ring depths)
FwdKernel fwd_variant(int v) {
    switch (v) {
        case 0: return mlp_fwd_kernel<false, Cfg<3, false, 0, true>>;                    // PE in the A tile, three weight slots
        case 1: return mlp_fwd_kernel<true, Cfg<kRing, false>>;
        case 8: return mlp_fwd_kernel<false, Cfg<3, false, 0, true>, true>;   // training: saves activations (PE in the A tile, three slots)
        case 9: return mlp_fwd_kernel<false, Cfg<3, false, 0, true>, false, true>;      // inference with the host tail (production)
#ifdef NERF_B200_EXPERIMENTS
        case 2: return mlp_fwd_kernel<true, Cfg<1, false>>;
        case 3: return mlp_fwd_kernel<true, Cfg<3, true>>;
        case 4: return mlp_fwd_kernel<true, Cfg<kRing, false, 1>>;
        case 5: return mlp_fwd_kernel<true, Cfg<kRing, false, 2>>;
        case 6: return mlp_fwd_kernel<true, Cfg<kRing, false, 4>>;
        case 7: return mlp_fwd_kernel<true, Cfg<kRing, false, 7>>;
        case 10: return mlp_fwd_kernel<false, Cfg<kRing, false>, false, true, true>;   // + 16-warp epilogue crew
        case 11: return mlp_fwd_kernel<true, Cfg<kRing, false, 8>>;
        case 12: return mlp_fwd_kernel<false, Cfg<kRing, false>, true>;       // training forward, round-1 layout: PE tiles + two slots (A/B)
        case 13: return mlp_fwd_kernel<false, Cfg<kRing, false, 8>, false, true, true>;   // no weight streaming + 16-warp crew (timing)
        case 14: return mlp_fwd_kernel<false, Cfg<kRing, false, 8>, false, true>;         // no weight streaming, host tail (timing)
        case 15: return mlp_fwd_kernel<false, Cfg<3, false, 8192, true>, false, true>;    // production with one epilogue body per layer (A/B)
        case 16: return mlp_fwd_kernel<false, Cfg<kRing, false>, false, true>;            // host tail, round-1 layout: PE tiles + two weight slots (A/B)
        case 17: return mlp_fwd_kernel<false, Cfg<kRing, false, 128>, false, true>;       // host tail, whole-warp MMA issuer with elect.sync (A/B)
        case 18: return mlp_fwd_kernel<false, Cfg<3, false, 2048, true>, false, true>;    // production + sampled wait profile (trace_out[0..50) slot waits, [64..84) A waits, [100] pairs)
        case 19: return mlp_fwd_kernel<false, Cfg<3, false, 16384, true>, false, true>;   // production with 32-column TMEM loads in the hidden epilogue (A/B)
        case 20: return mlp_fwd_kernel_r112<false, Cfg<3, false, 16384, true>, false, true>;   // ... and a 112-register budget
        case 21: return mlp_fwd_kernel_r112<false, Cfg<3, false, 0, true>, false, true>;       // production with a 112-register budget (A/B)
        case 22: return mlp_fwd_kernel<false, Cfg<3, false, 32768, true>, false, true>;        // production without weight-slot re-use (A/B)
        case 23: return mlp_fwd_kernel<false, Cfg<3, false, 65536, true>, false, true>;        // ... and both sub-tiles forward: the round-1 schedule (A/B)
#endif
        default: return nullptr;
    }
}

// CTA-pair kernel: clusters of two CTAs, one cluster per TPC.  kPairsUnavailable: the device cannot keep a
// single two-CTA cluster of this kernel resident (the occupancy query says so once per device); the caller then
// runs mlp_fwd_kernel, which computes the same bits.
constexpr int kPairsUnavailable = -1000;
int launch_fwd_pair(const FwdParams& P, int exp, void* stream) {
    constexpr int kMaxDevices = 64;
    static int max_clusters[kMaxDevices] = {};      // 0 not asked yet, -1 unavailable
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        nerf::set_last_error("nerf_mlp_fwd (CTA pairs) setup: %s", e0 != cudaSuccess ? cudaGetErrorString(e0) : "device index");
        return e0 != cudaSuccess ? (int)e0 : NERF_ERR_UNSUPPORTED;
    }
    FwdKernel k = mlp_fwd_pair_kernel<0>;
#ifdef NERF_B200_EXPERIMENTS
    if (exp == 2048) k = mlp_fwd_pair_kernel<2048>;      // + sampled wait profile
    // timing-only variants (wrong numerics): bit0 no A-tile stores, bit2 no TMEM loads, bit3 no weight streaming
    if (exp == 1) k = mlp_fwd_pair_kernel<1>;
    if (exp == 4) k = mlp_fwd_pair_kernel<4>;
    if (exp == 5) k = mlp_fwd_pair_kernel<5>;
    if (exp == 8) k = mlp_fwd_pair_kernel<8>;
    if (exp == 24) k = mlp_fwd_pair_kernel<24>;
    if (exp == 40) k = mlp_fwd_pair_kernel<40>;
    if (exp == 64) k = mlp_fwd_pair_kernel<64>;
    if (exp == 16384) k = mlp_fwd_pair_kernel<16384>;
    if (exp == (1 << 18)) k = mlp_fwd_pair_kernel<(1 << 18)>;
    if (exp != 0) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
    }
#else
    if (exp != 0) return nerf::arg_error("nerf_mlp_fwd (CTA pairs): variant");
#endif
    if (max_clusters[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        int sm_count = 0, n = 0;
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) {
            // how many 2-CTA clusters the device keeps resident at once (a GPC with an odd SM count loses one)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(sm_count & ~1));
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = kSmemBytes;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            e = cudaOccupancyMaxActiveClusters(&n, mlp_fwd_pair_kernel<0>, &cfg);
            if (e != cudaSuccess || n < 1) {       // no room for a cluster (or no cluster launch at all): not an error
                (void)cudaGetLastError();
                e = cudaSuccess;
                n = -1;
            }
        }
        if (e != cudaSuccess) {
            nerf::set_last_error("nerf_mlp_fwd (CTA pairs) setup: %s", cudaGetErrorString(e));
            return (int)e;
        }
        max_clusters[dev] = n;
    }
    if (max_clusters[dev] < 0) return kPairsUnavailable;
    const long n_tiles = (P.M + kTileM - 1) / kTileM;
    const long n_quads = (n_tiles + 3) / 4;
    const long clusters = n_quads < max_clusters[dev] ? n_quads : max_clusters[dev];
    k<<<(unsigned)(2 * clusters), kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    return nerf::check_launch("nerf_mlp_fwd (CTA pairs)");
}

int launch_fwd(const FwdParams& P, int variant, void* stream) {
    // per device (a process may hold tensors on several GPUs; the entry points switch to the buffers'
    // device): SM count, and the opt-in shared-memory size, which is a per-device function attribute
    constexpr int kMaxDevices = 64;
    static int sm_counts[kMaxDevices] = {};
    static bool configured_dev[kMaxDevices][24] = {};
    FwdKernel k = fwd_variant(variant);
    if (!k) return nerf::arg_error("nerf_mlp_fwd: variant");
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        nerf::set_last_error("nerf_mlp_fwd setup: %s", e0 != cudaSuccess ? cudaGetErrorString(e0) : "device index");
        return e0 != cudaSuccess ? (int)e0 : NERF_ERR_UNSUPPORTED;
    }
    int& sm_count = sm_counts[dev];
    bool* configured = configured_dev[dev];
    if (sm_count == 0) {
        cudaError_t e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) {
            sm_count = 0;
            nerf::set_last_error("nerf_mlp_fwd setup: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    if (!configured[variant]) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            nerf::set_last_error("nerf_mlp_fwd setup: %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured[variant] = true;
    }
    const long n_tiles = (P.M + kTileM - 1) / kTileM;
    const long n_pairs = (n_tiles + 1) / 2;
    const unsigned grid = (unsigned)(n_pairs < sm_count ? n_pairs : sm_count);
    k<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    return nerf::check_launch("nerf_mlp_fwd");
}

int fill_params(FwdParams& P, const void* packed, int in_mode, const float* in0, const float* in1,
                int in_stride, long M, int S, const float* vterm, int vterm_div, float* raw_out, long row0 = 0) {
    if (M < 0 || vterm_div < 1) return nerf::arg_error("nerf_mlp_fwd");
    if (M > 0 && (!packed || !in0 || !vterm || !raw_out)) return nerf::arg_error("nerf_mlp_fwd: null pointer");
    if (in_mode == NERF_IN_RAYS) {
        if ((M > 0 && !in1) || S < 1) return nerf::arg_error("nerf_mlp_fwd: NERF_IN_RAYS needs z and S");
    } else if (in_mode == NERF_IN_EMBEDDED) {
        if (in_stride < 63) return nerf::arg_error("nerf_mlp_fwd: in_stride < 63");
    } else if (in_mode != NERF_IN_POINTS) {
        return nerf::arg_error("nerf_mlp_fwd: in_mode");
    }
    P.blob = (const uint8_t*)packed;
    P.in_mode = in_mode; P.in0 = in0; P.in1 = in1; P.in_stride = in_stride;
    P.M = M; P.S = S < 1 ? 1 : S; P.vterm = vterm; P.vterm_div = vterm_div; P.raw_out = raw_out;
    P.probe_out = nullptr; P.probe_layer = -1; P.stats_out = nullptr; P.trace_out = nullptr; P.act_save = nullptr;
    // parity of the global index of the first tile: the chunk order of a row's contraction follows the parity of
    // its global 128-row tile (chunk_at); a shard that does not start on a tile boundary is numbered from 0
    P.tile0 = (row0 > 0 && (row0 % kTileM) == 0) ? (int)((row0 / kTileM) & 1) : 0;
    return 0;
}

}  // namespace

extern "C" int nerf_mlp_fwd(const void* packed, int in_mode, const float* in0, const float* in1,
                            int in_stride, long M, int S, const float* vterm, int vterm_div,
                            float* raw_out, void* act_save, long row0, void* stream) {
    nerf::DeviceGuard device_guard(raw_out);
    FwdParams P;
    int rc = fill_params(P, packed, in_mode, in0, in1, in_stride, M, S, vterm, vterm_div, raw_out, row0);
    if (rc) return rc;
    if (M == 0) return 0;
    if (act_save && ((uintptr_t)act_save & 15)) return nerf::arg_error("nerf_mlp_fwd: act_save must be 16-byte aligned");
    P.act_save = (uint8_t*)act_save;
#ifdef NERF_B200_EXPERIMENTS
    static const bool old_layout = getenv("NERF_B200_EXP_SAVE_OLD_LAYOUT") != nullptr;   // A/B timing only
    if (act_save && old_layout) return launch_fwd(P, 12, stream);
#endif
    return launch_fwd(P, act_save ? 8 : 0, stream);
}

extern "C" size_t nerf_model_host_tail_bytes(void) { return sizeof(ConstTail); }

// Copies the part of the packed blob's fp32 tail that the epilogues consume into HOST memory
// (asynchronously: synchronise the stream before passing it to nerf_mlp_fwd_host_tail).
extern "C" int nerf_model_host_tail(const void* packed, void* host_tail_out, void* stream) {
    nerf::DeviceGuard device_guard(packed);
    if (!packed || !host_tail_out) return nerf::arg_error("nerf_model_host_tail");
    cudaError_t e = cudaMemcpyAsync(host_tail_out, (const uint8_t*)packed + kWeightBytes, sizeof(ConstTail),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        nerf::set_last_error("nerf_model_host_tail: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// nerf_mlp_fwd with biases, l_alpha and l11 travelling in the kernel parameters (inference path).
extern "C" int nerf_mlp_fwd_host_tail(const void* packed, const void* host_tail, int in_mode, const float* in0,
                                      const float* in1, int in_stride, long M, int S, const float* vterm,
                                      int vterm_div, float* raw_out, long row0, void* stream) {
    nerf::DeviceGuard device_guard(raw_out);
    if (!host_tail) return nerf::arg_error("nerf_mlp_fwd_host_tail: host_tail");
    FwdParams P;
    int rc = fill_params(P, packed, in_mode, in0, in1, in_stride, M, S, vterm, vterm_div, raw_out, row0);
    if (rc) return rc;
    if (M == 0) return 0;
    memcpy(&P.ct, host_tail, sizeof(ConstTail));
    // The CTA-pair kernel (tcgen05 cta_group::2) is the production inference kernel; mlp_fwd_kernel produces the
    // same bits, stays selectable for A/B runs (NERF_B200_FWD_PAIRS=0) and takes over by itself on a device
    // that cannot co-schedule two-CTA clusters of this size.
    static const bool use_pairs = [] {
        const char* v = getenv("NERF_B200_FWD_PAIRS");
        return !(v && v[0] == '0');
    }();
    if (use_pairs) {
        rc = launch_fwd_pair(P, 0, stream);
        if (rc != kPairsUnavailable) return rc;
    }
    return launch_fwd(P, 9, stream);
}

// Test-support entry: additionally dumps the FP32 post-activation output of MMA layer
// `probe_layer` (0 = l1 ... 7 = l8, 8 = l10 with l9 folded in; 256 floats per row, l10 uses the first 128).
extern "C" int nerf_mlp_fwd_probe(const void* packed, int in_mode, const float* in0, const float* in1,
                                  int in_stride, long M, int S, const float* vterm, int vterm_div,
                                  float* raw_out, int probe_layer, float* probe_out, void* stream) {
    nerf::DeviceGuard device_guard(raw_out);
    FwdParams P;
    int rc = fill_params(P, packed, in_mode, in0, in1, in_stride, M, S, vterm, vterm_div, raw_out);
    if (rc) return rc;
    if (M == 0) return 0;
    P.probe_out = probe_out; P.probe_layer = probe_layer;
    return launch_fwd(P, 1, stream);
}

#ifdef NERF_B200_EXPERIMENTS
// Debug entry (tools/gpu_diag.py): run a pipeline variant and collect per-CTA cycle counters
// (stats_out [148][8] int64).  Variants other than 1 exist to time the pipeline only.
extern "C" int nerf_mlp_fwd_stats(const void* packed, const float* rays, const float* z, long M, int S,
                                  const float* vterm, float* raw_out, int variant, long long* stats_out,
                                  void* stream) {
    nerf::DeviceGuard device_guard(raw_out);
    FwdParams P;
    int rc = fill_params(P, packed, NERF_IN_RAYS, rays, z, 0, M, S, vterm, S, raw_out);
    if (rc) return rc;
    if (M == 0) return 0;
    P.stats_out = stats_out;
    if (variant >= 1000) {     // event trace behind the 148 x 8 counters (probe kernel)
        P.trace_out = stats_out + 148 * 8;
        variant -= 1000;
    }
    if (variant == 9 || variant == 10 || (variant >= 13 && variant <= 23) || (variant >= 30 && variant <= 45)) {   // host-tail kernels
        cudaError_t e = cudaMemcpy(&P.ct, (const uint8_t*)packed + kWeightBytes, sizeof(ConstTail), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return (int)e;
    }
    if (variant == 30) return launch_fwd_pair(P, 0, stream);          // CTA-pair kernel
    if (variant == 31) return launch_fwd_pair(P, 2048, stream);       // ... with the sampled wait profile
    if (variant == 32) return launch_fwd_pair(P, 4, stream);          // timing: no TMEM loads in the hidden epilogues
    if (variant == 33) return launch_fwd_pair(P, 1, stream);          // timing: no A-tile stores
    if (variant == 34) return launch_fwd_pair(P, 8, stream);          // timing: no weight streaming
    if (variant == 35) return launch_fwd_pair(P, 5, stream);          // timing: neither loads nor stores
    if (variant == 36) return launch_fwd_pair(P, 24, stream);         // timing: no weight streaming, no per-chunk commits
    if (variant == 37) return launch_fwd_pair(P, 40, stream);         // timing: no weight streaming, per-chunk commits without multicast
    if (variant == 38) return launch_fwd_pair(P, 16384, stream);      // A/B: 32-column TMEM loads in the hidden epilogue
    if (variant == 39) return launch_fwd_pair(P, 64, stream);         // A/B: three accumulator buffers in the hidden epilogue
    if (variant == 45) return launch_fwd_pair(P, 1 << 18, stream);            // A/B: the peer's relay on a barrier of its own (two waits per fill)
    return launch_fwd(P, variant < 1 ? 1 : variant, stream);
}

#endif  // NERF_B200_EXPERIMENTS

extern "C" size_t nerf_mlp_act_bytes(long M) {
    return M <= 0 ? 0 : (size_t)((M + kTileM - 1) / kTileM) * nerf::kActTileBytes;
}
