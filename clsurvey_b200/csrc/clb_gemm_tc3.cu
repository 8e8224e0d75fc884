// tcgen05 implicit-GEMM, third generation: operands fed by TMA (cp.async.bulk.tensor) from pre-split planes.
//
// Measured on gen-1/gen-2 (profiles/): TF32x1 and TF32x3 take the same time -- the gather warps (LDG -> mask/subtract ->
// STS/STTM, ~0.2 IPC under one CTA per SM) bound the kernel, not the tensor pipe.  Two facts make TMA possible here:
//   * the tensor core ignores the low 13 mantissa bits of an fp32 operand by itself (feeding raw fp32 as the "hi" term
//     gives bit-identical results to masking first: profiles/r1_rawhi_truncation.log), so the hi plane IS the tensor;
//   * the lo plane (x - trunc_tf32(x)) is one cheap streaming kernel per operand (8 B/element).
// What TMA can and cannot feed from the reference's NCHW fp32 layout (profiles/r1_tma_probe.log):
//   + weights re-ordered to [K][R*S*C] (2-D map), dY viewed as [img][kout][pq] (3-D map): K-major, no shifts -> legal;
//   - filter-tap windows of X: a +-1 pixel shift is a 4-byte offset in the innermost dimension; the TMA unit requires the
//     innermost start to be 16-byte aligned and raises "illegal instruction" otherwise (out-of-bounds in OUTER dimensions
//     zero-fills fine) -> tap rows stay on gather warps.  (An NHWC activation layout would move the shift to an outer
//     dimension; wgrad would then need MN-major operands, i.e. the SWIZZLE_128B_BASE32B descriptor for tf32.)
// Forward / dgrad keep the gen-2 A path (pixel gather -> TMEM) but take B (re-ordered weights, hi and lo planes written
// by the permute kernel) by TMA, which frees the B-loader warps: four A-loader groups fit in the register budget.
#include <stdlib.h>

#include "clb_tc_loaders.cuh"
#include "clb_tma.cuh"

namespace clb {
namespace tc3 {
using namespace clb::tc;
using clb::tcl::EpiNCHW;
using clb::tcl::PixelRows;
using clb::tcl::PixelRowsSmallC;
using clb::tcl::tmem_st32;
using clb::tcl::tmem_wait_st;
using clb::tcl::umma_tf32_ts;

// ---------------------------------------------------------------------------------------------- small helpers
__global__ void split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t n) {
    const int64_t n4 = n >> 2, gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gs) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        float4 o;
        o.x = v.x - __uint_as_float(__float_as_uint(v.x) & kHiMask);
        o.y = v.y - __uint_as_float(__float_as_uint(v.y) & kHiMask);
        o.z = v.z - __uint_as_float(__float_as_uint(v.z) & kHiMask);
        o.w = v.w - __uint_as_float(__float_as_uint(v.w) & kHiMask);
        reinterpret_cast<float4*>(lo)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t e = (n4 << 2) + threadIdx.x;
        lo[e] = x[e] - __uint_as_float(__float_as_uint(x[e]) & kHiMask);
    }
}

using clb::tcl::EpiSplitK;

constexpr int kTile = BM * 128;                     // 16 KB: 128 rows x 128 B

// ---------------------------------------------------------------------------------------------- fwd / dgrad: A gather -> TMEM, B by TMA
// 4 A-loader groups (16 warps, one GEMM row per thread, registers -> TMEM), warp 16 = TMA producer for the weight tiles,
// warp 17 = TMEM allocator + MMA issuer; warps 0-7 run the epilogue.  18 warps -> 20 warp slots -> 96 registers.
constexpr int kFwGroups = 4;
constexpr int kFwWarpTma = 4 * kFwGroups, kFwWarpMma = kFwWarpTma + 1;
constexpr int kFwThreads = (kFwWarpMma + 1) * 32;     // 576
#ifndef CLB_FW_STAGES_B
#define CLB_FW_STAGES_B 4
#endif
constexpr int kFwStagesA = 4, kFwStagesB = CLB_FW_STAGES_B;

template <int BN, bool WITH_LO> struct FwLayout {
    static constexpr int kBTile = BN * 128;
    static constexpr int kStageB = kBTile * (WITH_LO ? 2 : 1);
    static constexpr int kBarOff = kStageB * kFwStagesB;
    static constexpr int kTotal = kBarOff + 256 + 1024;
    static constexpr int kAccCols = WITH_LO ? 2 * BN : BN;
    static constexpr int kAStageCols = WITH_LO ? 64 : 32;
    static constexpr int kColsNeeded = kAccCols + kFwStagesA * kAStageCols;
    static constexpr int kTmemCols = kColsNeeded <= 128 ? 128 : (kColsNeeded <= 256 ? 256 : 512);
    static_assert(kColsNeeded <= 512, "TMEM budget");
    static_assert(kFwGroups == kFwStagesA, "group g must own TMEM stage g (parity waits stay within one phase)");
};

template <int BN, bool WITH_LO, class ALoad, class Epi>
__global__ void __launch_bounds__(kFwThreads, 1)
fwd_tma_kernel(ALoad A, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w_lo, Epi epi,
               int nkb) {
    using L = FwLayout<BN, WITH_LO>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + L::kBarOff;
    const uint32_t a_full = bar, a_empty = bar + 8 * kFwStagesA;
    const uint32_t b_full = bar + 16 * kFwStagesA, b_empty = b_full + 8 * kFwStagesB;
    const uint32_t bar_tmem = b_empty + 8 * kFwStagesB, slot = bar_tmem + 8;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kFwStagesA; ++s) { mbar_init(a_full + 8 * s, 4); mbar_init(a_empty + 8 * s, 1); }
        for (int s = 0; s < kFwStagesB; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
        tma::prefetch_desc(&map_w);
        if (WITH_LO) tma::prefetch_desc(&map_w_lo);
    }
    if (warp == kFwWarpMma) tmem_alloc(slot, L::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const uint32_t tmem_a0 = tmem + L::kAccCols;

    if (warp < kFwWarpTma) {
        const int group = warp >> 2, tg = threadIdx.x & 127;
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const typename ALoad::Ctx actx = A.prep(m0 + tg);
        for (int i = group; i < nkb; i += kFwGroups) {
            float v[BK];
#ifdef CLB_DIAG_NO_A_LOAD                       // timing experiments only (results are garbage)
#pragma unroll
            for (int j = 0; j < BK; ++j) v[j] = (float)(i + j + tg);
#else
            A.row(i, actx, v);
#endif
            const int s = i % kFwStagesA;
            mbar_wait(a_empty + 8 * s, (((uint32_t)(i / kFwStagesA)) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t col = tmem_a0 + (uint32_t)s * L::kAStageCols;
            {
                uint32_t raw[BK];                               // the tensor core ignores the low 13 mantissa bits itself
#pragma unroll
                for (int j = 0; j < BK; ++j) raw[j] = __float_as_uint(v[j]);
                tmem_st32(lane_field + col, raw);
            }
            if (WITH_LO) {
                uint32_t lo[BK];
#pragma unroll
                for (int j = 0; j < BK; ++j) lo[j] = __float_as_uint(v[j] - __uint_as_float(__float_as_uint(v[j]) & kHiMask));
                tmem_st32(lane_field + col + 32, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * s);
        }
    } else if (warp == kFwWarpTma) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kFwStagesB;
                mbar_wait(b_empty + 8 * s, (((uint32_t)(i / kFwStagesB)) & 1u) ^ 1u);
#ifdef CLB_DIAG_NO_B_TMA
                mbar_arrive(b_full + 8 * s);
#else
                tma::mbar_arrive_expect_tx(b_full + 8 * s, (uint32_t)L::kStageB);
                const uint32_t st = base + (uint32_t)s * L::kStageB;
                tma::load_2d(st, &map_w, b_full + 8 * s, i * BK, n0);
                if (WITH_LO) tma::load_2d(st + L::kBTile, &map_w_lo, b_full + 8 * s, i * BK, n0);
#endif
            }
        }
    } else if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(BN);
        for (int i = 0; i < nkb; ++i) {
            const int sa = i % kFwStagesA, sb = i % kFwStagesB;
            mbar_wait(a_full + 8 * sa, ((uint32_t)(i / kFwStagesA)) & 1u);
            mbar_wait(b_full + 8 * sb, ((uint32_t)(i / kFwStagesB)) & 1u);
            tc_fence_after();
            const uint32_t st = base + (uint32_t)sb * L::kStageB;
            const uint64_t b_hi = make_desc(st), b_lo = make_desc(st + L::kBTile);
            const uint32_t a_hi = tmem_a0 + (uint32_t)sa * L::kAStageCols, a_lo = a_hi + 32;
#ifndef CLB_DIAG_NO_MMA
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
                if (WITH_LO) {
                    umma_tf32_ts(tmem + BN, a_lo + 8 * k, b_hi + 2 * k, idesc, (i | k) != 0);
                    umma_tf32_ts(tmem + BN, a_hi + 8 * k, b_lo + 2 * k, idesc, 1);
                }
                umma_tf32_ts(tmem, a_hi + 8 * k, b_hi + 2 * k, idesc, (i | k) != 0);
            }
#endif
            umma_commit(a_empty + 8 * sa);
            umma_commit(b_empty + 8 * sb);
        }
        umma_commit(bar_tmem);
    }

    if (warp < 8) {
        if (nkb > 0) {
            mbar_wait(bar_tmem, 0);
            tc_fence_after();
        }
        const int lane_grp = warp & 3, col_half = warp >> 2;
        const int m = m0 + lane_grp * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 16) {
            const int col = col_half * (BN / 2) + c;
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)col, r);
            if (WITH_LO) {
                uint32_t r2[16];
                tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(BN + col), r2);
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            }
            epi.store16(m, n0 + col, r, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kFwWarpMma) tmem_dealloc(tmem, L::kTmemCols);
}

template <int BN, bool WITH_LO, class ALoad>
static int launch_fwd(const ALoad& A, const float* w2, const float* w2_lo, int n_rows, int ld, const EpiNCHW& e, int M, int nkb,
                      cudaStream_t s) {
    using L = FwLayout<BN, WITH_LO>;
    CUtensorMap mw, mwl;
    const uint64_t dims[2] = {(uint64_t)ld, (uint64_t)n_rows};
    const uint64_t str[1] = {(uint64_t)ld * 4};
    const uint32_t box[2] = {32, (uint32_t)BN};
    int rc = tma::encode_f32(&mw, w2, 2, dims, str, box, true);
    if (rc) return rc;
    rc = tma::encode_f32(&mwl, WITH_LO ? w2_lo : w2, 2, dims, str, box, true);
    if (rc) return rc;
    auto kern = fwd_tma_kernel<BN, WITH_LO, ALoad, EpiNCHW>;
    static bool configured = false;
    if (!configured) { CLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal)); configured = true; }
    dim3 grid((M + BM - 1) / BM, (n_rows + BN - 1) / BN, 1);
    kern<<<grid, kFwThreads, L::kTotal, s>>>(A, mw, mwl, e, nkb); clb::count_launch();
    return CLB_OK;
}

// ---------------------------------------------------------------------------------------------- wgrad, mixed feed
// A = dY rows [kout][32 pixels] by TMA (raw + lo plane; K-major in NCHW, no shifts -> TMA-legal), B = filter-tap rows
// gathered by 4 loader groups (the +-1 pixel shifts are 4-byte misaligned in the innermost dimension, which the TMA
// unit rejects: profiles/r1_tma_probe.log), both operands in swizzled smem, SS MMAs, split accumulators.
// Loader groups == stages so that group g always refills stage g: a parity wait on an mbarrier is only safe when the
// waiter is at most one phase behind; 4 groups over 3 stages let a group run two phases ahead of a stage it had not
// touched for a while (observed as a hang at > 8 K blocks per CTA).
constexpr int kMxGroups = 3;
constexpr int kMxWarpTma = 4 * kMxGroups, kMxWarpMma = kMxWarpTma + 1;
constexpr int kMxThreads = (kMxWarpMma + 1) * 32;        // 448
constexpr int kMxStages = 3;
constexpr int kMxStage = 4 * kTile;                       // A_hi, A_lo, B_hi, B_lo (16 KB each)
constexpr int kMxSmem = kMxStages * kMxStage + 256 + 1024;
static_assert(kMxGroups == kMxStages, "group g must own stage g");

template <bool WITH_LO>
__global__ void __launch_bounds__(kMxThreads, 1)
wgrad_mixed_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_dy_lo,
                   clb::tcl::TapRows<128> B, EpiSplitK epi, int rows_per_img, int num_kb_total, int kb_per_split) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + kMxStages * kMxStage;
    const uint32_t a_full = bar, b_full = bar + 8 * kMxStages, empty = bar + 16 * kMxStages;
    const uint32_t bar_tmem = empty + 8 * kMxStages, slot = bar_tmem + 8;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * 128, z = blockIdx.z;
    const int kb_begin = z * kb_per_split;
    const int nkb = max(min(num_kb_total, kb_begin + kb_per_split) - kb_begin, 0);
    constexpr uint32_t kCols = WITH_LO ? 256 : 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMxStages; ++s) {
            mbar_init(a_full + 8 * s, 1);
            mbar_init(b_full + 8 * s, 4);
            mbar_init(empty + 8 * s, 1);
        }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
        tma::prefetch_desc(&map_dy);
        if (WITH_LO) tma::prefetch_desc(&map_dy_lo);
    }
    if (warp == kMxWarpMma) tmem_alloc(slot, kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;

    if (warp < kMxWarpTma) {
        const int group = warp >> 2, tg = threadIdx.x & 127;
        const typename clb::tcl::TapRows<128>::Ctx bctx = B.prep(tg, n0);
        for (int i = group; i < nkb; i += kMxGroups) {
            clb::tcl::BRegs<128> cur;
            B.load(kb_begin + i, tg, bctx, cur);
            const int s = i % kMxStages;
            mbar_wait(empty + 8 * s, (((uint32_t)(i / kMxStages)) & 1u) ^ 1u);
            const uint32_t st = base + (uint32_t)s * kMxStage;
            clb::tcl::store_b<128, WITH_LO>(cur, tg, st + 2 * kTile, st + 3 * kTile);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_full + 8 * s);
        }
    } else if (warp == kMxWarpTma) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kMxStages;
                mbar_wait(empty + 8 * s, (((uint32_t)(i / kMxStages)) & 1u) ^ 1u);
                const int kb = kb_begin + i;
                const int img = kb / rows_per_img, pq0 = (kb - img * rows_per_img) * 32;
                tma::mbar_arrive_expect_tx(a_full + 8 * s, (uint32_t)((WITH_LO ? 2 : 1) * kTile));
                const uint32_t st = base + (uint32_t)s * kMxStage;
                tma::load_3d(st, &map_dy, a_full + 8 * s, pq0, m0, img);
                if (WITH_LO) tma::load_3d(st + kTile, &map_dy_lo, a_full + 8 * s, pq0, m0, img);
            }
        }
    } else if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(128);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % kMxStages;
            const uint32_t ph = ((uint32_t)(i / kMxStages)) & 1u;
            mbar_wait(a_full + 8 * s, ph);
            mbar_wait(b_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t st = base + (uint32_t)s * kMxStage;
            const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + kTile);
            const uint64_t b_hi = make_desc(st + 2 * kTile), b_lo = make_desc(st + 3 * kTile);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
                if (WITH_LO) {
                    umma_tf32(tmem + 128, a_lo + 2 * k, b_hi + 2 * k, idesc, (i | k) != 0);
                    umma_tf32(tmem + 128, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
                }
                umma_tf32(tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (i | k) != 0);
            }
            umma_commit(empty + 8 * s);
        }
        umma_commit(bar_tmem);
    }

    if (warp < 8) {
        if (nkb > 0) {
            mbar_wait(bar_tmem, 0);
            tc_fence_after();
        }
        const int lane_grp = warp & 3, col_half = warp >> 2;
        const int m = m0 + lane_grp * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 64; c += 16) {
            const int col = col_half * 64 + c;
            uint32_t r[16];
            if (nkb > 0) {
                tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)col, r);
                if (WITH_LO) {
                    uint32_t r2[16];
                    tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(128 + col), r2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            epi.store16(m, n0 + col, r, z);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMxWarpMma) tmem_dealloc(tmem, kCols);
}

// ---------------------------------------------------------------------------------------------- plain GEMM, all-TMA (Linear layers)
// C[M][N] = act(A[M][K] * B[N][K]^T + bias): both operands are K-major row-major matrices, i.e. exactly what a 2-D tensor
// map describes -- no gather warps at all.  warp 0 = TMA producer (A, A_lo, B, B_lo boxes of 128 x 32), warp 1 = MMA
// issuer (3-pass split, two TMEM accumulators), warps 2-5 = epilogue.  nn.Linear fwd / dgrad / wgrad map onto it through
// (small) transposed copies; see clb_linear_* in clb_gemm_simt.cu.
constexpr int kGmStages = 3;
constexpr int kGmThreads = 192;
constexpr int kGmStage = 4 * kTile;
constexpr int kGmSmem = kGmStages * kGmStage + 256 + 1024;

struct EpiRowMajor {
    float* c; int64_t ldc; const float* bias; int relu, M, N;
    __device__ __forceinline__ void store16(int m, int n0, const uint32_t (&r)[16], int) const {
        if (m >= M) return;
        float* dst = c + (int64_t)m * ldc + n0;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            v[j] = __uint_as_float(r[j]) + ((bias && n0 + j < N) ? __ldg(bias + n0 + j) : 0.f);
            if (relu) v[j] = fmaxf(v[j], 0.f);
        }
        if (n0 + 15 < N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (n0 + j < N) dst[j] = v[j];
        }
    }
};

template <bool WITH_LO>
__global__ void __launch_bounds__(kGmThreads, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_b_lo, EpiRowMajor epi, int nkb) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + kGmStages * kGmStage;
    const uint32_t full = bar, empty = bar + 8 * kGmStages, bar_tmem = empty + 8 * kGmStages, slot = bar_tmem + 8;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * 128;
    constexpr uint32_t kCols = WITH_LO ? 256 : 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kGmStages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
        tma::prefetch_desc(&map_a); tma::prefetch_desc(&map_b);
    }
    if (warp == 1) tmem_alloc(slot, kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < nkb; ++i) {
            const int s = i % kGmStages;
            mbar_wait(empty + 8 * s, (((uint32_t)(i / kGmStages)) & 1u) ^ 1u);
            tma::mbar_arrive_expect_tx(full + 8 * s, (uint32_t)((WITH_LO ? 4 : 2) * kTile));
            const uint32_t st = base + (uint32_t)s * kGmStage;
            tma::load_2d(st, &map_a, full + 8 * s, i * BK, m0);
            tma::load_2d(st + 2 * kTile, &map_b, full + 8 * s, i * BK, n0);
            if (WITH_LO) {
                tma::load_2d(st + kTile, &map_a_lo, full + 8 * s, i * BK, m0);
                tma::load_2d(st + 3 * kTile, &map_b_lo, full + 8 * s, i * BK, n0);
            }
        }
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc(128);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % kGmStages;
            mbar_wait(full + 8 * s, ((uint32_t)(i / kGmStages)) & 1u);
            tc_fence_after();
            const uint32_t st = base + (uint32_t)s * kGmStage;
            const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + kTile);
            const uint64_t b_hi = make_desc(st + 2 * kTile), b_lo = make_desc(st + 3 * kTile);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
                if (WITH_LO) {
                    umma_tf32(tmem + 128, a_lo + 2 * k, b_hi + 2 * k, idesc, (i | k) != 0);
                    umma_tf32(tmem + 128, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
                }
                umma_tf32(tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (i | k) != 0);
            }
            umma_commit(empty + 8 * s);
        }
        umma_commit(bar_tmem);
    } else if (warp >= 2) {
        mbar_wait(bar_tmem, 0);
        tc_fence_after();
        const int lane_grp = warp & 3;
        const int m = m0 + lane_grp * 32 + lane;
#pragma unroll 1
        for (int col = 0; col < 128; col += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)col, r);
            if (WITH_LO) {
                uint32_t r2[16];
                tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(128 + col), r2);
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            }
            epi.store16(m, n0 + col, r, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kCols);
}

// out[j][i] (row pitch ldo) = in[i][j] (rows x cols, row pitch ldi); optional lo plane of the transposed matrix
__global__ void transpose_split_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out_lo,
                                       int rows, int cols, int ldi, int ldo) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(int64_t)r * ldi + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;          // output row = input column
        if (c < cols && r < ldo) {
            const float v = r < rows ? tile[threadIdx.x][i] : 0.f;
            out[(int64_t)c * ldo + r] = v;
            if (out_lo) out_lo[(int64_t)c * ldo + r] = v - __uint_as_float(__float_as_uint(v) & kHiMask);
        }
    }
}

static inline int ew_blocks(int64_t n) {
    int64_t b = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace tc3

// forward conv (also dgrad, as a forward conv of dY): w2 / w2_lo are the re-ordered weight planes [K][ld]
int tc3_conv_fwd(const float* x, const float* w2, const float* w2_lo, const float* bias, float* y, int N, int C, int H, int W,
                 int K, int R, int S, int pad, int relu, bool with_lo, cudaStream_t s) {
    using namespace tc3;
    const int P = H, Q = W, M = N * P * Q;
    EpiNCHW e{y, bias, relu, M, K, P * Q, FastDiv32(P * Q)};
    const bool wide = (K % 128 == 0 || K > 64);
    if (C % 32 != 0) {
        PixelRowsSmallC A{x, C, H, W, R, S, pad, P, Q, M, R * S * C, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
        if (wide) return with_lo ? launch_fwd<128, true>(A, w2, w2_lo, K, 32, e, M, 1, s) : launch_fwd<128, false>(A, w2, w2_lo, K, 32, e, M, 1, s);
        return with_lo ? launch_fwd<64, true>(A, w2, w2_lo, K, 32, e, M, 1, s) : launch_fwd<64, false>(A, w2, w2_lo, K, 32, e, M, 1, s);
    }
    PixelRows A{x, C, H, W, R, S, pad, P, Q, M, FastDiv32(P * Q), FastDiv32(Q), FastDiv32(C), FastDiv32(S)};
    const int ld = R * S * C, nkb = ld / BK;
    if (wide) return with_lo ? launch_fwd<128, true>(A, w2, w2_lo, K, ld, e, M, nkb, s) : launch_fwd<128, false>(A, w2, w2_lo, K, ld, e, M, nkb, s);
    return with_lo ? launch_fwd<64, true>(A, w2, w2_lo, K, ld, e, M, nkb, s) : launch_fwd<64, false>(A, w2, w2_lo, K, ld, e, M, nkb, s);
}

// mixed-feed wgrad: dY rows by TMA need the 32-pixel K blocks to stay inside one image (H*W % 32 == 0)
bool tc3_wgrad_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("CLB_TC_WGRAD_TMA");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    if (!enabled) return false;
    if (!(stride == 1 && R == S && 2 * pad == R - 1)) return false;
    return (H * W) % 32 == 0 && (W % 4) == 0;
}

// extra floats of workspace: the lo plane of dy
size_t tc3_wgrad_extra_floats(int N, int C, int H, int W, int K) { return (size_t)N * K * H * W + 8; }

// ws layout: [split-K partials] [dy_lo]
int tc3_conv_wgrad(const float* x, const float* dy, float* ws_partials, float* bias_part, float* dy_lo, int N, int C,
                   int H, int W, int K, int R, int S, int pad, bool with_lo, int splits, int kb_per_split, cudaStream_t s) {
    using namespace tc3;
    const int PQ = H * W, n_rows = R * S * C, npix = N * PQ;
    if (with_lo && bias_part) {
        clb::conv_bias_partials_and_lo(dy, dy_lo, bias_part, N, K, PQ, s);      // one read of dY for both
    } else if (with_lo) {
        split_lo_kernel<<<ew_blocks(((int64_t)N * K * PQ) >> 2), 256, 0, s>>>(dy, dy_lo, (int64_t)N * K * PQ); clb::count_launch();
    }
    CUtensorMap m_dy, m_dy_lo;
    const uint64_t dims[3] = {(uint64_t)PQ, (uint64_t)K, (uint64_t)N};
    const uint64_t str[2] = {(uint64_t)PQ * 4, (uint64_t)K * PQ * 4};
    const uint32_t box[3] = {32, 128, 1};
    int rc = tma::encode_f32(&m_dy, dy, 3, dims, str, box, true);
    if (rc) return rc;
    rc = tma::encode_f32(&m_dy_lo, with_lo ? dy_lo : dy, 3, dims, str, box, true);
    if (rc) return rc;
    clb::tcl::TapRows<128> B{x, C, H, W, R, S, pad, H, W, n_rows, npix, FastDiv32(PQ), FastDiv32(W), FastDiv32(C), FastDiv32(S)};
    EpiSplitK e{ws_partials, K, n_rows, (int64_t)K * n_rows};
    dim3 grid((K + BM - 1) / BM, (n_rows + 127) / 128, splits);
    static bool configured[2] = {false, false};
    if (with_lo) {
        if (!configured[1]) { CLB_CUDA(cudaFuncSetAttribute(wgrad_mixed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMxSmem)); configured[1] = true; }
        wgrad_mixed_kernel<true><<<grid, kMxThreads, kMxSmem, s>>>(m_dy, m_dy_lo, B, e, PQ / 32, npix / 32, kb_per_split); clb::count_launch();
    } else {
        if (!configured[0]) { CLB_CUDA(cudaFuncSetAttribute(wgrad_mixed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMxSmem)); configured[0] = true; }
        wgrad_mixed_kernel<false><<<grid, kMxThreads, kMxSmem, s>>>(m_dy, m_dy_lo, B, e, PQ / 32, npix / 32, kb_per_split); clb::count_launch();
    }
    return CLB_OK;
}

// ---- Linear layers on the TMA GEMM ------------------------------------------------------------------------------
static int gemm_tma(const float* A, const float* A_lo, int64_t lda, const float* B, const float* B_lo, int64_t ldb, float* C,
                    int64_t ldc, const float* bias, int relu, int M, int N, int K, bool with_lo, cudaStream_t s) {
    using namespace tc3;
    CUtensorMap ma, mal, mb, mbl;
    const uint32_t box[2] = {32, 128};
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        const uint64_t str[1] = {(uint64_t)lda * 4};
        int rc = tma::encode_f32(&ma, A, 2, dims, str, box, true);
        if (rc) return rc;
        rc = tma::encode_f32(&mal, with_lo ? A_lo : A, 2, dims, str, box, true);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        const uint64_t str[1] = {(uint64_t)ldb * 4};
        int rc = tma::encode_f32(&mb, B, 2, dims, str, box, true);
        if (rc) return rc;
        rc = tma::encode_f32(&mbl, with_lo ? B_lo : B, 2, dims, str, box, true);
        if (rc) return rc;
    }
    EpiRowMajor e{C, ldc, bias, relu, M, N};
    dim3 grid((M + BM - 1) / BM, (N + 127) / 128, 1);
    const int nkb = (K + BK - 1) / BK;
    static bool configured[2] = {false, false};
    if (with_lo) {
        if (!configured[1]) { CLB_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGmSmem)); configured[1] = true; }
        gemm_tma_kernel<true><<<grid, kGmThreads, kGmSmem, s>>>(ma, mal, mb, mbl, e, nkb); clb::count_launch();
    } else {
        if (!configured[0]) { CLB_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGmSmem)); configured[0] = true; }
        gemm_tma_kernel<false><<<grid, kGmThreads, kGmSmem, s>>>(ma, mal, mb, mbl, e, nkb); clb::count_launch();
    }
    return CLB_OK;
}

static inline size_t r4(size_t n) { return (n + 3) & ~(size_t)3; }
static void split_lo(const float* x, float* lo, int64_t n, cudaStream_t s) {
    tc3::split_lo_kernel<<<tc3::ew_blocks(n >> 2), 256, 0, s>>>(x, lo, n); clb::count_launch();
}
static void transpose_split(const float* in, float* out, float* out_lo, int rows, int cols, int ldi, int ldo, cudaStream_t s) {
    dim3 grid((cols + 31) / 32, (ldo + 31) / 32), block(32, 8);
    tc3::transpose_split_kernel<<<grid, block, 0, s>>>(in, out, out_lo, rows, cols, ldi, ldo); clb::count_launch();
}

bool tc3_linear_supported(int M, int in, int out) { return (in % 4) == 0 && (out % 4) == 0 && in >= 32; }
// workspace floats for the three Linear passes (the largest is used for all)
size_t tc3_linear_ws_floats(int M, int in, int out) {
    const size_t ldm = r4((size_t)M), ldo = r4((size_t)out);
    const size_t fwd = r4((size_t)M * in) + r4((size_t)out * in);
    const size_t dgr = r4((size_t)M * out) + 2 * r4((size_t)in * ldo);
    const size_t wgr = 2 * r4((size_t)out * ldm) + 2 * r4((size_t)in * ldm);
    size_t m = fwd > dgr ? fwd : dgr;
    return (m > wgr ? m : wgr) + 16;
}
int tc3_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* ws, int M, int in, int out, int relu,
                   bool with_lo, cudaStream_t s) {
    float* x_lo = ws;
    float* w_lo = ws + r4((size_t)M * in);
    if (with_lo) { split_lo(x, x_lo, (int64_t)M * in, s); split_lo(w, w_lo, (int64_t)out * in, s); }
    return gemm_tma(x, x_lo, in, w, w_lo, in, y, out, bias, relu, M, out, in, with_lo, s);
}
int tc3_linear_dgrad(const float* dy, const float* w, float* dx, float* ws, int M, int in, int out, bool with_lo, cudaStream_t s) {
    const int ldo = (int)r4((size_t)out);
    float* dy_lo = ws;
    float* wt = ws + r4((size_t)M * out);
    float* wt_lo = wt + r4((size_t)in * ldo);
    if (with_lo) split_lo(dy, dy_lo, (int64_t)M * out, s);
    transpose_split(w, wt, with_lo ? wt_lo : nullptr, out, in, in, ldo, s);          // W[out][in] -> Wt[in][ldo]
    return gemm_tma(dy, dy_lo, out, wt, wt_lo, ldo, dx, in, nullptr, 0, M, in, out, with_lo, s);
}
int tc3_linear_wgrad(const float* x, const float* dy, float* dw, float* ws, int M, int in, int out, bool with_lo, cudaStream_t s) {
    const int ldm = (int)r4((size_t)M);
    float* dyt = ws;
    float* dyt_lo = dyt + r4((size_t)out * ldm);
    float* xt = dyt_lo + r4((size_t)out * ldm);
    float* xt_lo = xt + r4((size_t)in * ldm);
    transpose_split(dy, dyt, with_lo ? dyt_lo : nullptr, M, out, out, ldm, s);       // dY[M][out] -> dYt[out][ldm]
    transpose_split(x, xt, with_lo ? xt_lo : nullptr, M, in, in, ldm, s);            // X[M][in]   -> Xt[in][ldm]
    return gemm_tma(dyt, dyt_lo, ldm, xt, xt_lo, ldm, dw, in, nullptr, 0, out, in, M, with_lo, s);
}

}  // namespace clb
