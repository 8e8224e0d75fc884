// "Planes" convolution kernels: 3x3 / stride 1 / pad 1 convolutions over activations kept as NHWC bf16 hi/lo planes.
//
// Replaces nn.Conv2d forward / backward of the reference's VGG stacks (src/models/VGGSlim.py:27-40 called from
// src/methods/EWC/train_EWC.py:181-187 and twins).  Round-1 kernels (clb_gemm_tc4.cu) gathered NCHW fp32 activations with
// 4-byte loads and converted them to bf16 hi/lo in the main loop; tensor pipe 25-50 % (profiles/README.md).  Here every
// activation lives as two bf16 planes x = hi + lo (hi = bf16_rn(x), lo = bf16_rn(x - hi)) in NHWC, written once by the
// producing epilogue, so that BOTH operands of every conv GEMM are plain TMA box loads:
//   * fwd / dgrad : A = 128 output pixels x 64 channels of ONE filter tap = a 4-D box (c, w, h, n) of the input planes
//                   shifted by the tap offset (TMA zero-fills the halo), K-major, 128B swizzle;
//                   B = BN filters x 64 channels of the re-ordered weight planes [K][tap][C], K-major.
//   * wgrad       : reduction over pixels, so both operands are MN-major: A = dY box [64 pixels][128 kout],
//                   B = X box shifted by the tap [64 pixels][2 x 64 channels].
// One persistent CTA per SM: warp 0 = TMA producer, warp 1 = MMA issuer (tcgen05.mma kind::f16, SS), warps 2-9 = epilogue.
// fp32 parity through the 3-pass split hi*hi + (hi*lo + lo*hi), main and cross terms in separate TMEM accumulators
// (DESIGN.md 4.1); two accumulator sets so the epilogue of tile i runs under the main loop of tile i+1.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "clb_planes.cuh"
#include "clb_tma.cuh"

namespace clb {
namespace pl {
using namespace clb::tc;

constexpr int kStages = 3;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;          // 320
constexpr int BKP = 64;                                  // K elements (bf16) per K block = one 128-byte swizzled row
#ifndef CLB_PLANES_CHUNK
#define CLB_PLANES_CHUNK 16
#endif
constexpr int kChunk = CLB_PLANES_CHUNK;                 // K blocks chained into one TMEM accumulator before it is drained to registers
constexpr int kTileBytes = 128 * 128;                    // 128 rows x 64 bf16
constexpr int kStageBytes = 4 * kTileBytes;              // A_hi, A_lo, B_hi, B_lo (BN <= 128)
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;

__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ((++spins & 0xFFFu) == 0)) {                       // watchdog: a lost arrive must not hang the GPU box
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000LL) __trap();
        }
    } while (!ok);
}

__device__ __forceinline__ void umma_bf16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, N >> 3 at bit 17, M = 128 at bit 24, majors at bits 15 / 16
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// MN-major, SWIZZLE_128B descriptor: rows of 64 bf16 (128 B) along M/N, 8 K-rows per 1024-byte atom (SBO), 64-wide atoms
// along M/N `lbo` bytes apart  (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    const uint64_t lo = (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16);
    const uint64_t hi = (uint64_t)(sbo >> 4) | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
    return lo | (hi << 32);
}
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);                     // .x = v0 -> low half
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// tile of `rows` consecutive NHWC pixels: whole image rows (bw == W), bh rows, bn images
struct TileGeom {
    int bh, bn, tiles_h, log_rows_img;                   // rows of one image inside a tile = bh * W = 1 << log_rows_img
};
static TileGeom make_geom(int rows, int H, int W) {
    TileGeom g;
    g.bh = rows / W < H ? rows / W : H;
    g.bn = rows / (W * g.bh);
    g.tiles_h = H / g.bh;
    int l = 0;
    while ((1 << l) < g.bh * W) ++l;
    g.log_rows_img = l;
    return g;
}

struct ConvArgs {
    int N, H, W, Cred, Cout;                             // Cred = reduction channels, Cout = output channels of the GEMM
    int taps, pad;                                       // 9 / 1: 3x3 conv;  1 / 0: nn.Linear as a 1x1 'conv' on a 1x1 map (rows = samples)
    int n_items, n_tiles_n;                              // work items (persistent loop) and N tiles per M tile
    TileGeom tg;                                         // M tile (fwd / dgrad) or K block (wgrad) geometry
    // fwd / dgrad epilogue
    const float* bias; int relu;
    const uint16_t* mask_hi;                             // dgrad: zero where the (post-ReLU) input activation is <= 0
    uint16_t *y_hi, *y_lo;
    // wgrad
    float* ws; int splits, kb_per_split, n_kb_total, n_tiles_m, ncb;     // ncb = 64-wide (tap, c) column blocks
    uint32_t mn_lbo, mn_sbo;                             // MN-major descriptor strides (bytes)
};

// ------------------------------------------------------------------------------------------------ the kernel
// KIND 0: fwd / dgrad (A, B K-major).  KIND 1: wgrad (A, B MN-major).
template <int KIND, int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_planes_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const ConvArgs p) {
    pdl_trigger();                                       // the next kernel of the chain may start its own prologue
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + kStages * kStageBytes;
    const uint32_t full = bar, empty = bar + 8 * kStages, acc_full = empty + 8 * kStages, acc_empty = acc_full + 16;
    const uint32_t slot = acc_empty + 16;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kAccCols = 2 * BN;                     // main + cross-term accumulator
    constexpr int kTmemCols = 2 * kAccCols <= 256 ? 256 : 512;
    constexpr uint32_t kABytes = kTileBytes, kBBytes = (uint32_t)BN * 128u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + 8 * s, 1); mbar_init(acc_empty + 8 * s, kEpiWarps); }
        fence_barrier_init();
        tma::prefetch_desc(&map_a_hi); tma::prefetch_desc(&map_a_lo);
        tma::prefetch_desc(&map_b_hi); tma::prefetch_desc(&map_b_lo);
    }
    if (warp == 1) tmem_alloc(slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const int cpb = p.Cred >> 6;                         // 64-channel blocks per tap
    pdl_wait();                                          // barriers, TMEM and descriptors are set up: now wait for the producer of
    //                                                      the operands (the predecessor kernel) to complete

    if (warp == 0) {
        // ================================================================================ TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                if (KIND == 0) {
                    const int nt = item % p.n_tiles_n, mt = item / p.n_tiles_n;
                    const int n0 = (mt / p.tg.tiles_h) * p.tg.bn, h0 = (mt % p.tg.tiles_h) * p.tg.bh;
                    const int nkb = p.taps * cpb;
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % kStages;
                        mbar_wait_wd(empty + 8 * s, ((it / kStages) & 1u) ^ 1u);
                        const int tap = kb / cpb, cb = kb - tap * cpb;
                        const int r = tap / 3, sx = tap - r * 3;
                        const uint32_t st = base + (uint32_t)s * kStageBytes;
                        tma::mbar_arrive_expect_tx(full + 8 * s, 2 * kABytes + 2 * kBBytes);
                        tma::load_4d(st, &map_a_hi, full + 8 * s, cb * 64, sx - p.pad, h0 + r - p.pad, n0);
                        tma::load_4d(st + kTileBytes, &map_a_lo, full + 8 * s, cb * 64, sx - p.pad, h0 + r - p.pad, n0);
                        tma::load_2d(st + 2 * kTileBytes, &map_b_hi, full + 8 * s, kb * 64, nt * BN);
                        tma::load_2d(st + 2 * kTileBytes + kBBytes, &map_b_lo, full + 8 * s, kb * 64, nt * BN);
                    }
                } else {
                    const int mt = item % p.n_tiles_m, nt = (item / p.n_tiles_m) % p.n_tiles_n;
                    const int z = item / (p.n_tiles_m * p.n_tiles_n);
                    const int kb0 = z * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.n_kb_total);
                    const int j0 = 2 * nt, nblk = (j0 + 1 < p.ncb) ? 2 : 1;
                    int tap[2], cb[2];
                    for (int j = 0; j < 2; ++j) { tap[j] = (j0 + j) / cpb; cb[j] = (j0 + j) - tap[j] * cpb; }
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % kStages;
                        mbar_wait_wd(empty + 8 * s, ((it / kStages) & 1u) ^ 1u);
                        const int n0 = (kb / p.tg.tiles_h) * p.tg.bn, h0 = (kb % p.tg.tiles_h) * p.tg.bh;
                        const uint32_t st = base + (uint32_t)s * kStageBytes;
                        tma::mbar_arrive_expect_tx(full + 8 * s, (uint32_t)(2 * kTileBytes + nblk * 2 * 8192));
                        for (int j = 0; j < 2; ++j) {                 // dY: two 64-kout sub-tiles [64 pixels][64]
                            tma::load_4d(st + j * 8192, &map_a_hi, full + 8 * s, mt * 128 + j * 64, 0, h0, n0);
                            tma::load_4d(st + kTileBytes + j * 8192, &map_a_lo, full + 8 * s, mt * 128 + j * 64, 0, h0, n0);
                        }
                        for (int j = 0; j < nblk; ++j) {              // X shifted by the tap of column block j
                            const int r = tap[j] / 3, sx = tap[j] - r * 3;
                            tma::load_4d(st + 2 * kTileBytes + j * 8192, &map_b_hi, full + 8 * s, cb[j] * 64, sx - p.pad, h0 + r - p.pad, n0);
                            tma::load_4d(st + 3 * kTileBytes + j * 8192, &map_b_lo, full + 8 * s, cb[j] * 64, sx - p.pad, h0 + r - p.pad, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================================ MMA issuer
        if (lane == 0) {
            uint32_t it = 0, ccount = 0;                  // ccount: accumulator chunks issued so far (chunk c uses set c & 1)
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int nkb;
                uint32_t idesc, idesc_wide = 0;
                if (KIND == 0) {
                    nkb = p.taps * cpb;
                    idesc = idesc_bf16(BN, false);
                    idesc_wide = idesc_bf16(2 * BN, false);
                } else {
                    const int nt = (item / p.n_tiles_m) % p.n_tiles_n, z = item / (p.n_tiles_m * p.n_tiles_n);
                    const int kb0 = z * p.kb_per_split;
                    nkb = min(kb0 + p.kb_per_split, p.n_kb_total) - kb0;
                    idesc = idesc_bf16((2 * nt + 1 < p.ncb) ? 128 : 64, true);
                }
                uint32_t d_main = 0, d_cross = 0;
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int ic = i % kChunk;                        // position inside the accumulator chunk
                    if (ic == 0) {
                        const uint32_t as = ccount & 1u;
                        mbar_wait_wd(acc_empty + 8 * as, ((ccount >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        d_main = tmem + as * kAccCols;
                        d_cross = d_main + BN;
                    }
                    const int s = it % kStages;
                    mbar_wait_wd(full + 8 * s, (it / kStages) & 1u);
                    tc_fence_after();
                    const uint32_t st = base + (uint32_t)s * kStageBytes;
                    if (KIND == 0) {
                        // B_lo lies right behind B_hi, and the cross accumulator right behind the main one: ONE N = 2 BN
                        // instruction computes A_hi.[B_hi | B_lo] -> [main | cross] and reads A_hi from shared memory once
                        // (20 KB instead of 24 KB of operand reads per K step -- the kernel is shared-memory bound)
                        const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + kTileBytes);
                        const uint64_t b_hi = make_desc(st + 2 * kTileBytes);
#pragma unroll
                        for (int k = 0; k < BKP / 16; ++k) {
                            umma_bf16_ss(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc_wide, (ic | k) != 0);
                            umma_bf16_ss(d_cross, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < BKP / 16; ++k) {
                            const uint32_t o = (uint32_t)k * 2048u;   // 16 pixel rows = two 1024-byte K groups
                            const uint64_t a_hi = make_desc_mn(st + o, p.mn_lbo, p.mn_sbo), a_lo = make_desc_mn(st + kTileBytes + o, p.mn_lbo, p.mn_sbo);
                            const uint64_t b_hi = make_desc_mn(st + 2 * kTileBytes + o, p.mn_lbo, p.mn_sbo);
                            const uint64_t b_lo = make_desc_mn(st + 3 * kTileBytes + o, p.mn_lbo, p.mn_sbo);
                            // (the N = 256 form over [X_hi | X_lo] measured 7 % SLOWER with MN-major operands: three MMAs here)
                            umma_bf16_ss(d_cross, a_lo, b_hi, idesc, (ic | k) != 0);
                            umma_bf16_ss(d_cross, a_hi, b_lo, idesc, 1);
                            umma_bf16_ss(d_main, a_hi, b_hi, idesc, (ic | k) != 0);
                        }
                    }
                    umma_commit(empty + 8 * s);
                    if (ic == kChunk - 1 || i == nkb - 1) {           // chunk complete: hand it to the epilogue warps
                        umma_commit(acc_full + 8 * (ccount & 1u));
                        ++ccount;
                    }
                }
            }
        }
    } else {
        // ================================================================================ epilogue (8 warps)
        const int e = warp - 2, lane_grp = warp & 3, half = e >> 2;
        const int row = lane_grp * 32 + lane;
        const uint32_t lane_field = (uint32_t)(lane_grp * 32) << 16;
        uint32_t ccount = 0;
        constexpr int kCols = BN / 2;                     // columns of this thread's half
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int nkb;
            if (KIND == 0) nkb = p.taps * cpb;
            else {
                const int kb0 = (item / (p.n_tiles_m * p.n_tiles_n)) * p.kb_per_split;
                nkb = min(kb0 + p.kb_per_split, p.n_kb_total) - kb0;
            }
            // Drain every chunk of kChunk K blocks into registers with round-to-nearest fp32 adds: the TMEM accumulate
            // truncates, and its bias grows with the number of MMAs chained into one accumulator
            float v[kCols];
#pragma unroll
            for (int j = 0; j < kCols; ++j) v[j] = 0.f;
            const int nchunks = (nkb + kChunk - 1) / kChunk;
            for (int ch = 0; ch < nchunks; ++ch, ++ccount) {
                const uint32_t as = ccount & 1u;
                mbar_wait_wd(acc_full + 8 * as, (ccount >> 1) & 1u);
                tc_fence_after();
                const uint32_t t_main = tmem + as * kAccCols + lane_field + (uint32_t)(half * kCols), t_cross = t_main + BN;
#pragma unroll
                for (int c = 0; c < kCols; c += 16) {
                    uint32_t r[16], r2[16];
                    tmem_ld16(t_main + (uint32_t)c, r);
                    tmem_ld16(t_cross + (uint32_t)c, r2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c + j] += __uint_as_float(r[j]) + __uint_as_float(r2[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * as);
            }
            if (KIND == 0) {
                const int nt = item % p.n_tiles_n, mt = item / p.n_tiles_n;
                const int n0 = (mt / p.tg.tiles_h) * p.tg.bn, h0 = (mt % p.tg.tiles_h) * p.tg.bh;
                const int img = n0 + (row >> p.tg.log_rows_img);
                const int pix_in = row & ((1 << p.tg.log_rows_img) - 1);
                const bool valid = img < p.N;
                const size_t off = (((size_t)img * p.H + h0) * p.W + pix_in) * p.Cout + (size_t)nt * BN + half * kCols;
#pragma unroll
                for (int c = 0; c < kCols; c += 16) {
                    if (p.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + nt * BN + half * kCols + c);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b = __ldg(bp + j);
                            v[c + 4 * j] += b.x; v[c + 4 * j + 1] += b.y; v[c + 4 * j + 2] += b.z; v[c + 4 * j + 3] += b.w;
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[c + j] = fmaxf(v[c + j], 0.f);
                    }
                    if (valid) {
                        if (p.mask_hi) {
                            const uint4* mp = reinterpret_cast<const uint4*>(p.mask_hi + off + c);
                            const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
                            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {             // bf16 > 0  <=>  sign clear and magnitude non-zero
                                const uint32_t lo16 = mw[j] & 0xFFFFu, hi16 = mw[j] >> 16;
                                if (!(lo16 != 0 && lo16 < 0x8000u)) v[c + 2 * j] = 0.f;
                                if (!(hi16 != 0 && hi16 < 0x8000u)) v[c + 2 * j + 1] = 0.f;
                            }
                        }
                        uint32_t h[8], l[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) split_pair(v[c + 2 * j], v[c + 2 * j + 1], h[j], l[j]);
                        st_global_v4(p.y_hi + off + c, h[0], h[1], h[2], h[3]);
                        st_global_v4(p.y_hi + off + c + 8, h[4], h[5], h[6], h[7]);
                        st_global_v4(p.y_lo + off + c, l[0], l[1], l[2], l[3]);
                        st_global_v4(p.y_lo + off + c + 8, l[4], l[5], l[6], l[7]);
                    }
                }
            } else {
                const int mt = item % p.n_tiles_m, nt = (item / p.n_tiles_m) % p.n_tiles_n;
                const int z = item / (p.n_tiles_m * p.n_tiles_n);
                const int ld = p.ncb * 64;                            // taps * Cred
                const int kout = mt * 128 + row;
                float* dst = p.ws + ((size_t)z * p.Cout + kout) * ld + (size_t)nt * 128 + half * 64;
                const bool second = 2 * nt + 1 < p.ncb;
                if ((half == 0 || second) && kout < p.Cout) {
#pragma unroll
                    for (int c = 0; c < kCols; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ host side
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// nn.Linear(in, out) on the same kernel: a 1x1 "conv" over a 1x1 map whose 128-row tiles are 128 samples
bool linear_supported(int in, int out) { return in >= 64 && (in % 64) == 0 && (out % 64) == 0; }

bool conv_supported(int C, int H, int W, int K, int R, int S, int stride, int pad) {
    return R == 3 && S == 3 && stride == 1 && pad == 1 && (C % 64) == 0 && (K % 64) == 0 && pow2(H) && pow2(W) && W >= 4 &&
           W <= 64 && H >= 2 && H * W >= 16;
}

// 4-D map over NHWC planes [N][H][W][C]: dims (C, W, H, N), box (64, W, bh, bn)
static int encode_act(CUtensorMap* m, const uint16_t* base, int N, int H, int W, int C, int bh, int bn) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)C * W * 2, (uint64_t)C * W * H * 2};
    const uint32_t box[4] = {64u, (uint32_t)W, (uint32_t)bh, (uint32_t)bn};
    return tma::encode_bf16(m, base, 4, dims, str, box, true);
}

template <int KIND, int BN>
static int launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                  const ConvArgs& p, cudaStream_t s) {
    auto kern = conv_planes_kernel<KIND, BN>;
    static bool configured = false;
    if (!configured) { CLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); configured = true; }
    const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
    launch_pdl(kern, dim3(grid), dim3(kThreads), kSmemBytes, s, a_hi, a_lo, b_hi, b_lo, p); clb::count_launch();
    return CLB_OK;
}

// y = conv3x3(x, w) [+ bias] [ReLU] [masked by mask_hi > 0], all planes NHWC.  w planes: [Cout][9][Cred] (K-major rows).
int conv_fwd(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, int relu,
             const uint16_t* mask_hi, uint16_t* y_hi, uint16_t* y_lo, int N, int H, int W, int Cred, int Cout, int taps, cudaStream_t s) {
    ConvArgs p{};
    p.N = N; p.H = H; p.W = W; p.Cred = Cred; p.Cout = Cout;
    p.taps = taps; p.pad = taps == 9 ? 1 : 0;
    p.tg = make_geom(128, H, W);
    const int bn_tile = (Cout % 128 == 0) ? 128 : 64;
    p.n_tiles_n = Cout / bn_tile;
    const int m_tiles = ((N + p.tg.bn - 1) / p.tg.bn) * p.tg.tiles_h;
    p.n_items = m_tiles * p.n_tiles_n;
    p.bias = bias; p.relu = relu; p.mask_hi = mask_hi; p.y_hi = y_hi; p.y_lo = y_lo;
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    int rc;
    if ((rc = encode_act(&a_hi, x_hi, N, H, W, Cred, p.tg.bh, p.tg.bn))) return rc;
    if ((rc = encode_act(&a_lo, x_lo, N, H, W, Cred, p.tg.bh, p.tg.bn))) return rc;
    const uint64_t dims[2] = {(uint64_t)taps * Cred, (uint64_t)Cout};
    const uint64_t str[1] = {(uint64_t)taps * Cred * 2};
    const uint32_t box[2] = {64u, (uint32_t)bn_tile};
    if ((rc = tma::encode_bf16(&b_hi, w_hi, 2, dims, str, box, true))) return rc;
    if ((rc = tma::encode_bf16(&b_lo, w_lo, 2, dims, str, box, true))) return rc;
    return bn_tile == 128 ? launch<0, 128>(a_hi, a_lo, b_hi, b_lo, p, s) : launch<0, 64>(a_hi, a_lo, b_hi, b_lo, p, s);
}

// split-K plan of the wgrad GEMM [Cout x 9 Cred] over N*H*W pixels: as many splits as fill ONE wave of the 148 SMs (every
// persistent CTA then does exactly one equally long work item -- no tail round -- and the partial-sum workspace is as small as
// the parallelism allows: 1 split, i.e. no reduction traffic at all, for the 144-tile layers).  Fixed 148 => the summation
// order, and with it every bit of dW, does not depend on the device it runs on.
void wgrad_plan(int N, int H, int W, int Cred, int Cout, int taps, int* splits, int* kb_per_split, int* n_kb) {
    const TileGeom g = make_geom(64, H, W);
    const int nkb = ((N + g.bn - 1) / g.bn) * g.tiles_h;
    const int tiles = ((Cout + 127) / 128) * ((taps * (Cred / 64) + 1) / 2);
    int want = 148 / tiles;
    if (want > nkb) want = nkb;
    if (want < 1) want = 1;
    const int per = (nkb + want - 1) / want;
    *kb_per_split = per;
    *splits = (nkb + per - 1) / per;
    *n_kb = nkb;
}
size_t wgrad_ws_floats(int N, int H, int W, int Cred, int Cout, int taps) {
    int splits, per, nkb;
    wgrad_plan(N, H, W, Cred, Cout, taps, &splits, &per, &nkb);
    return (size_t)splits * Cout * taps * Cred;
}

// ws[z][Cout][9][Cred] = partial sums of dY^T X_tap over the pixel range of split z
int conv_wgrad_partials(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* dy_hi, const uint16_t* dy_lo, float* ws,
                        int* splits_out, int N, int H, int W, int Cred, int Cout, int taps, cudaStream_t s) {
    ConvArgs p{};
    p.N = N; p.H = H; p.W = W; p.Cred = Cred; p.Cout = Cout;
    p.taps = taps; p.pad = taps == 9 ? 1 : 0;
    p.tg = make_geom(64, H, W);
    wgrad_plan(N, H, W, Cred, Cout, taps, &p.splits, &p.kb_per_split, &p.n_kb_total);
    p.ncb = taps * (Cred / 64);
    p.n_tiles_m = (Cout + 127) / 128;
    p.n_tiles_n = (p.ncb + 1) / 2;
    p.n_items = p.n_tiles_m * p.n_tiles_n * p.splits;
    p.ws = ws;
    p.mn_lbo = 8192u; p.mn_sbo = 1024u;              // MN-major, 128B swizzle: 64-wide MN atoms 8 KB apart, 8-row K groups 1 KB apart
    *splits_out = p.splits;
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    int rc;
    if ((rc = encode_act(&a_hi, dy_hi, N, H, W, Cout, p.tg.bh, p.tg.bn))) return rc;
    if ((rc = encode_act(&a_lo, dy_lo, N, H, W, Cout, p.tg.bh, p.tg.bn))) return rc;
    if ((rc = encode_act(&b_hi, x_hi, N, H, W, Cred, p.tg.bh, p.tg.bn))) return rc;
    if ((rc = encode_act(&b_lo, x_lo, N, H, W, Cred, p.tg.bh, p.tg.bn))) return rc;
    return launch<1, 128>(a_hi, a_lo, b_hi, b_lo, p, s);
}

}  // namespace pl
}  // namespace clb
