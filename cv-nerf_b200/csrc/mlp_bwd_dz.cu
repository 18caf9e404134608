// Backward of the field network, part 1: the dZ chain (gradients of every layer's pre-activation).
//
// Replaces what autograd does for Model.forward (/root/reference/model.py:77-107) when the train
// loop calls loss.backward() (/root/reference/main.py:385), restricted to the activations: for
// each 128-sample tile
//     dZ10 = (grad_rgb . W11) * [h10 > 0]                        CUDA cores, FP32
//     dZ8  = (dZ10 . W' + grad_sigma * w_alpha) * [h8 > 0]       W' = W10[:, :256] . W9: l9 has no
//                                                                 activation and is folded into l10
//                                                                 (mlp_layout.h), so there is no dZ9
//     dZi  = (dZ(i+1) . W(i+1)) * [hi > 0]            i = 7..1    (l6 contributes its h5 columns)
// as BF16 tensor-core contractions against the transposed weights (mlp_bwd_layout.h) with FP32
// accumulation in tensor memory.  Every dZ tile is dumped to HBM as a tile image (one bulk copy
// per layer) for the dW contraction (mlp_bwd_dw.cu); the ReLU masks come from the activation
// record the forward kernel saved.
//
// Same warp roles and ping-pong as mlp_fwd.cu: warp 0 streams 32 KB weight slots, warp 1 issues
// tcgen05.mma (M=128, N=256, K=16), warps 2-9 / 10-17 own one 128-row sub-tile each.
#include <cuda_bf16.h>

#include "common.cuh"
#include "mlp_bwd_layout.h"
#include "umma.cuh"

namespace {

using namespace nerf;

constexpr int kTileM = 128;
constexpr int kRing = 3;                                 // 32 KB weight slots (three fit beside the two A tiles)
constexpr int kSlotBytes = 2 * kStageBytes;
constexpr int kEpiWarpsPerGroup = 8;
constexpr int kGroupThreads = kEpiWarpsPerGroup * 32;
constexpr int kThreads = 64 + 2 * kGroupThreads;         // 576
constexpr uint32_t kOffA = 0;                            // 2 x [4][128][64] bf16
constexpr uint32_t kOffW = 2 * 65536;                    // ring x 32 KB
constexpr uint32_t kOffBar = kOffW + kRing * kSlotBytes;
constexpr uint32_t kOffTail = kOffBar + 256;             // l_alpha.weight [256], l11.weight [3][128] fp32
constexpr uint32_t kSmemBytes = kOffTail + kBwdTailFloats * 4;
static_assert(kSmemBytes <= 232448, "shared memory budget");
constexpr uint32_t kIdescN256 = umma::instr_desc_bf16(128, 256);

struct DzParams {
    const uint8_t* blob;      // transposed weights + tail
    const float* grad_raw;    // [M][4]
    const uint8_t* act;       // activation records
    uint8_t* dz;              // dZ records
    long M;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// x * [h > 0] on packed BF16 pairs
__device__ __forceinline__ uint32_t relu_mask_mul(uint32_t x, uint32_t h) {
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
    __nv_bfloat162 hv = *reinterpret_cast<__nv_bfloat162*>(&h);
    __nv_bfloat162 xv = *reinterpret_cast<__nv_bfloat162*>(&x);
    __nv_bfloat162 r = __hmul2(xv, __hgt2(hv, zero));
    return *reinterpret_cast<uint32_t*>(&r);
}

// one 32-byte sector = a swizzled pair of 16-byte chunks (LDG.256)
__device__ __forceinline__ void ldg_nc_v8(const void* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Head stage: dZ10 for this thread's 64 columns [half*64, half*64+64) of its row -> A tile block
// `half` (K-major operand of the first contraction).
// expand the two mask bits of BF16 pair j (iteration parity b) to all-ones / all-zeros halves
__device__ __forceinline__ uint32_t pair_keep_mask(uint32_t m, int b, int j) {
    return ((m >> (2 * j + b)) & 0x00010001u) * 0xffffu;
}

__device__ __forceinline__ void head_stage(float4 g, uint2 mask, uint32_t row_addr, uint32_t swz,
                                           int half, uint32_t w11_addr) {
#pragma unroll
    for (int c16 = 0; c16 < 8; ++c16) {
        const uint32_t mword = (c16 >> 2) ? mask.y : mask.x;      // 16-column iterations 2p, 2p+1 -> word p
        const int b = (c16 >> 1) & 1, j0 = (c16 & 1) * 4;
        const int c = half * 64 + c16 * 8;
        float dh[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 w0 = umma::ld_shared_v4f(w11_addr + (0 * kL10Out + c + q * 4) * 4);
            const float4 w1 = umma::ld_shared_v4f(w11_addr + (1 * kL10Out + c + q * 4) * 4);
            const float4 w2 = umma::ld_shared_v4f(w11_addr + (2 * kL10Out + c + q * 4) * 4);
            dh[q * 4 + 0] = fmaf(g.z, w2.x, fmaf(g.y, w1.x, g.x * w0.x));
            dh[q * 4 + 1] = fmaf(g.z, w2.y, fmaf(g.y, w1.y, g.x * w0.y));
            dh[q * 4 + 2] = fmaf(g.z, w2.z, fmaf(g.y, w1.z, g.x * w0.z));
            dh[q * 4 + 3] = fmaf(g.z, w2.w, fmaf(g.y, w1.w, g.x * w0.w));
        }
        const uint32_t o0 = pack_bf16x2(dh[0], dh[1]) & pair_keep_mask(mword, b, j0 + 0);
        const uint32_t o1 = pack_bf16x2(dh[2], dh[3]) & pair_keep_mask(mword, b, j0 + 1);
        const uint32_t o2 = pack_bf16x2(dh[4], dh[5]) & pair_keep_mask(mword, b, j0 + 2);
        const uint32_t o3 = pack_bf16x2(dh[6], dh[7]) & pair_keep_mask(mword, b, j0 + 3);
        umma::st_shared_v4(row_addr + half * 16384 + (((uint32_t)c16 << 4) ^ swz), o0, o1, o2, o3);
    }
}

// One layer's epilogue for this thread's 128 columns [c0, c0+128): accumulator (+ grad_sigma *
// w_alpha when ADD_SIGMA) -> BF16 -> * [h > 0] (when MASK) -> A tile in place.
// h_row: this row's line in block 0 of the saved activation (global), or NULL for zero rows.
template <bool MASK, bool ADD_SIGMA>
__device__ __forceinline__ void epilogue_dz(uint32_t tacc, int c0, uint32_t row_addr, uint32_t swz,
                                            uint4 mask, float dsig, uint32_t walpha_addr) {
    uint32_t v[2][16];
    umma::tmem_ld16(tacc + c0, v[0]);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int c = c0 + it * 16;
        umma::tmem_wait_ld();
        if (it + 1 < 8) umma::tmem_ld16(tacc + c + 16, v[(it + 1) & 1]);
        const uint32_t(&cur)[16] = v[it & 1];
        float2 d[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = make_float2(__uint_as_float(cur[2 * e]), __uint_as_float(cur[2 * e + 1]));
        if (ADD_SIGMA) {
            const float2 s2 = make_float2(dsig, dsig);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = umma::ld_shared_v4f(walpha_addr + (c + q * 4) * 4);
                d[2 * q] = __ffma2_rn(s2, make_float2(w.x, w.y), d[2 * q]);
                d[2 * q + 1] = __ffma2_rn(s2, make_float2(w.z, w.w), d[2 * q + 1]);
            }
        }
        const uint32_t mword = (it >> 1) == 0 ? mask.x : (it >> 1) == 1 ? mask.y : (it >> 1) == 2 ? mask.z : mask.w;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint32_t o0 = pack_bf16x2(d[q * 4 + 0].x, d[q * 4 + 0].y), o1 = pack_bf16x2(d[q * 4 + 1].x, d[q * 4 + 1].y);
            uint32_t o2 = pack_bf16x2(d[q * 4 + 2].x, d[q * 4 + 2].y), o3 = pack_bf16x2(d[q * 4 + 3].x, d[q * 4 + 3].y);
            if (MASK) {
                o0 &= pair_keep_mask(mword, it & 1, q * 4 + 0); o1 &= pair_keep_mask(mword, it & 1, q * 4 + 1);
                o2 &= pair_keep_mask(mword, it & 1, q * 4 + 2); o3 &= pair_keep_mask(mword, it & 1, q * 4 + 3);
            }
            const int cc = c + q * 8;
            const int blk = cc >> 6, c16 = (cc & 63) >> 3;
            umma::st_shared_v4(row_addr + blk * 16384 + (((uint32_t)c16 << 4) ^ swz), o0, o1, o2, o3);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_dz_kernel(const DzParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = umma::smem_u32(smem);
    if ((sbase & 1023u) != 0) __trap();
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
            umma::mbar_init(bar_a_ready + 8 * g, kGroupThreads);
            umma::mbar_init(bar_acc_full + 8 * g, 1);
        }
        umma::fence_barrier_init();
    }
    if (warp == 1) {
        umma::tmem_alloc(umma::smem_u32(tmem_slot), 512);
        umma::tmem_relinquish();
    }
    // fp32 tail -> shared memory (read with broadcast LDS by the epilogues)
    {
        const float* tail = reinterpret_cast<const float*>(P.blob + kBwdWeightBytes);
        for (int i = threadIdx.x; i < kBwdTailFloats; i += kThreads)
            reinterpret_cast<float*>(smem + kOffTail)[i] = __ldg(tail + i);
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: transposed weight slots, L2 -> smem =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int j = 0; j < kBwdLayers; ++j) {
                    const int first = bwd_first_stage(j), chunks = bwd_chunks(j);
                    for (int g = 0; g < 2; ++g) {
                        for (int c = 0; c < chunks; ++c, ++it) {
                            const uint32_t slot = it % kRing, ph = (it / kRing) & 1;
                            umma::mbar_wait(bar_w_empty + 8 * slot, ph ^ 1);
                            umma::mbar_arrive_expect_tx(bar_w_full + 8 * slot, kSlotBytes);
                            umma::bulk_g2s(sbase + kOffW + slot * kSlotBytes,
                                           P.blob + (size_t)first * kStageBytes + (size_t)c * kSlotBytes, kSlotBytes,
                                           bar_w_full + 8 * slot);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // lean issue loop as in mlp_fwd.cu: ring position kept incrementally, descriptors from
            // precomputed words, no tcgen05 fence after the TMA-signalled slot barrier
            uint32_t slot = 0, ph = 0, n_ready[2] = {0, 0};
            constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << (46 - 32)) | (2u << (61 - 32));
            auto desc = [&](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
            auto desc_lo = [&](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
            const uint32_t w_lo = desc_lo(sbase + kOffW);
            for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int j = 0; j < kBwdLayers; ++j) {
                    const int chunks = bwd_chunks(j);
                    for (int g = 0; g < 2; ++g) {
                        umma::mbar_wait(bar_a_ready + 8 * g, n_ready[g] & 1);
                        ++n_ready[g];
                        umma::tc_fence_after();
                        const uint32_t d_base = tmem_base + g * 256;
                        const uint32_t a_tile_lo = desc_lo(sbase + kOffA + g * 65536);
                        for (int c = 0; c < chunks; ++c) {
                            umma::mbar_wait(bar_w_full + 8 * slot, ph);
                            const uint32_t a_lo = a_tile_lo + c * (16384 >> 4);
                            const uint32_t b_lo = w_lo + slot * (kSlotBytes >> 4);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                umma::mma_bf16_ss(d_base, desc(a_lo + kk * 2), desc(b_lo + kk * 2), kIdescN256,
                                                  (c > 0 || kk > 0) ? 1u : 0u);
                            }
                            umma::mma_commit(bar_w_empty + 8 * slot);
                            if (++slot == (uint32_t)kRing) { slot = 0; ph ^= 1; }
                        }
                        umma::mma_commit(bar_acc_full + 8 * g);
                    }
                }
            }
        }
    } else {
        // ===================== epilogue groups =====================
        const int ew = warp - 2;
        const int g = ew >> 3;
        const int half = (ew >> 2) & 1;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t group_bar = 1 + g;
        const int gtid = (ew & 7) * 32 + lane;
        const uint32_t a_tile_addr = sbase + kOffA + g * 65536;
        const uint32_t a_row_addr = a_tile_addr + row * 128;
        const uint32_t swz = (uint32_t)(row & 7) << 4;
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + g * 256;
        const uint32_t walpha_addr = sbase + kOffTail + kBwdTailWAlpha * 4;
        const uint32_t w11_addr = sbase + kOffTail + kBwdTailW11 * 4;
        uint32_t n_full = 0;
        for (long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            const long tile = pair * 2 + g;
            const bool tile_ok = tile < n_tiles;
            const long grow = tile * kTileM + row;
            const bool valid = grow < P.M;
            const uint8_t* act_tile = P.act + (size_t)tile * kActTileBytes;
            uint8_t* dz_tile = P.dz + (size_t)tile * kDzTileBytes;
            const bool saver = gtid == 0 && tile_ok;
            // rows past M carry a zero gradient, so their (duplicated) activations never reach dW
            const float4 graw = valid ? __ldg(reinterpret_cast<const float4*>(P.grad_raw) + grow)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            // ReLU masks are bits written by the forward epilogue thread that owned the same (row, half)
            const uint2 m10 = tile_ok ? __ldg(reinterpret_cast<const uint2*>(act_tile + act_mask_slot(8, row, half)))
                                      : make_uint2(0, 0);
            head_stage(graw, m10, a_row_addr, swz, half, w11_addr);
            umma::fence_proxy_async_smem();
            umma::mbar_arrive(bar_a_ready + 8 * g);
            umma::named_bar_sync(group_bar, kGroupThreads);
            if (gtid == 0) {
                if (saver) {
                    umma::bulk_s2g(dz_tile + kDz10, a_tile_addr, 32768);
                    umma::bulk_commit();
                }
                umma::bulk_wait_read0();
            }
            umma::named_bar_sync(group_bar, kGroupThreads);
#pragma unroll 1
            for (int j = 0; j < kBwdLayers; ++j) {
                const int i = 8 - j;               // this epilogue produces dZ_i
                const int c0 = half * 128;
                // one 16-byte load per thread and layer, issued before the accumulator wait
                uint4 mk = make_uint4(0, 0, 0, 0);
                if (tile_ok) mk = __ldg(reinterpret_cast<const uint4*>(act_tile + act_mask_slot(i - 1, row, half)));
                umma::mbar_wait(bar_acc_full + 8 * g, n_full & 1);
                ++n_full;
                umma::tc_fence_after();
                if (j == 0) {          // dZ8: the sigma head's gradient joins here
                    epilogue_dz<true, true>(tacc, c0, a_row_addr, swz, mk, graw.w, walpha_addr);
                } else {
                    epilogue_dz<true, false>(tacc, c0, a_row_addr, swz, mk, 0.f, walpha_addr);
                }
                umma::fence_proxy_async_smem();
                umma::tc_fence_before();
                if (j + 1 < kBwdLayers) umma::mbar_arrive(bar_a_ready + 8 * g);
                umma::named_bar_sync(group_bar, kGroupThreads);
                if (gtid == 0) {
                    if (saver) {
                        umma::bulk_s2g(dz_tile + dz_hidden(i), a_tile_addr, 65536);
                        umma::bulk_commit();
                    }
                    // the copy must have read the tile before the next epilogue / head stage rewrites it
                    umma::bulk_wait_read0();
                }
                umma::named_bar_sync(group_bar, kGroupThreads);
            }
        }
        if (gtid == 0) umma::bulk_wait_all();
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

extern "C" size_t nerf_mlp_dz_bytes(long M) {
    return M <= 0 ? 0 : (size_t)((M + kTileM - 1) / kTileM) * nerf::kDzTileBytes;
}

extern "C" int nerf_mlp_bwd_dz(const void* packed_bwd, const float* grad_raw, const void* act_save, long M,
                               void* dz_out, void* stream) {
    nerf::DeviceGuard device_guard(dz_out);
    if (M < 0 || (M > 0 && (!packed_bwd || !grad_raw || !act_save || !dz_out))) return nerf::arg_error("nerf_mlp_bwd_dz");
    if (M == 0) return 0;
    if (((uintptr_t)act_save | (uintptr_t)dz_out | (uintptr_t)grad_raw) & 15)
        return nerf::arg_error("nerf_mlp_bwd_dz: buffers must be 16-byte aligned");
    static int sm_count = 0;
    static bool configured = false;
    if (sm_count == 0) {
        sm_count = nerf_b200_sm_count();
        if (sm_count <= 0) {
            sm_count = 0;
            nerf::set_last_error("nerf_mlp_bwd_dz: no CUDA device");
            return (int)cudaErrorNoDevice;
        }
    }
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_bwd_dz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            nerf::set_last_error("nerf_mlp_bwd_dz setup: %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    DzParams P;
    P.blob = (const uint8_t*)packed_bwd; P.grad_raw = grad_raw; P.act = (const uint8_t*)act_save;
    P.dz = (uint8_t*)dz_out; P.M = M;
    const long n_pairs = ((M + kTileM - 1) / kTileM + 1) / 2;
    const unsigned grid = (unsigned)(n_pairs < sm_count ? n_pairs : sm_count);
    mlp_bwd_dz_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    return nerf::check_launch("nerf_mlp_bwd_dz");
}
