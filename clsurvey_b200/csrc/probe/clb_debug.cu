// Diagnostic kernels used while bringing the tcgen05 paths up (not on the hot path).
// clb_debug_umma_mn: D[128x128] = A[128x32] * B[128x32]^T with A staged in shared memory in an MN-major (M contiguous)
// 128B-swizzled layout -- probes which (atom order, LBO/SBO) convention the MN-major matrix descriptor expects, so that
// NCHW activations (pixels contiguous) can be fed to the tensor core by TMA without a register transpose.
#include "../clb_tc_ptx.cuh"
#include "../clb_tc_loaders.cuh"

namespace clb {
namespace dbg {
using namespace clb::tc;

__global__ void __launch_bounds__(128, 1) umma_mn_kernel(const float* __restrict__ At /*[32][128]*/,
                                                         const float* __restrict__ B /*[128][32]*/,
                                                         float* __restrict__ D /*[128][128]*/, int variant) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_tile = base, b_tile = base + 16384, bar = base + 32768, slot = bar + 8;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool k_group_major = variant & 1;     // atom order
    const bool swap_lbo_sbo = variant & 2;      // which stride goes to which descriptor field
    const uint32_t stride_ma = k_group_major ? 1024u : 4096u, stride_kg = k_group_major ? 4096u : 1024u;
    // A: element (m, k) -> atom (m/32, k/8), row k%8, 16B chunk ((m%32)/4) ^ (k%8)
    if (variant >= 4) {                                     // harness self-check: A staged K-major (the known-good layout)
        for (int idx = tid; idx < 32 * 128; idx += 128) {
            const int k = idx >> 7, m = idx & 127;
            const uint32_t off = (uint32_t)m * 128u + (uint32_t)((((k >> 2) ^ (m & 7)) << 4) + (k & 3) * 4);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_tile + off), "r"(__float_as_uint(At[k * 128 + m])) : "memory");
        }
    } else
    for (int idx = tid; idx < 32 * 32; idx += 128) {       // (k, m4)
        const int k = idx >> 5, m4 = idx & 31, m = m4 * 4;
        const float4 v = *reinterpret_cast<const float4*>(At + k * 128 + m);
        const uint32_t ma = m >> 5, kg = k >> 3, kr = k & 7, ch = (m & 31) >> 2;
        const uint32_t off = ma * stride_ma + kg * stride_kg + kr * 128u + ((ch ^ kr) << 4);
        st_shared_v4(a_tile + off, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
    }
    // B: K-major rows (known-good layout)
    for (int idx = tid; idx < 128 * 8; idx += 128) {
        const int r = idx >> 3, c = idx & 7;
        const float4 v = *reinterpret_cast<const float4*>(B + r * 32 + c * 4);
        st_shared_v4(b_tile + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4), __float_as_uint(v.x), __float_as_uint(v.y),
                     __float_as_uint(v.z), __float_as_uint(v.w));
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    if (tid == 0) {
        const uint32_t lbo = swap_lbo_sbo ? stride_kg : stride_ma, sbo = swap_lbo_sbo ? stride_ma : stride_kg;
        // idesc: as make_idesc(128) plus a_major = MN (bit 15)
        const uint32_t idesc = make_idesc(128) | (1u << 15);
        const uint64_t bdesc = make_desc(b_tile);
        if (variant >= 4) {
            const uint64_t adesc = make_desc(a_tile);
            for (int k = 0; k < 4; ++k) umma_tf32(tmem, adesc + 2 * k, bdesc + 2 * k, make_idesc(128), k != 0);
        } else
        for (int k = 0; k < 4; ++k) {
            const uint32_t a_addr = a_tile + (uint32_t)k * stride_kg;         // next 8 K rows
            const uint64_t lo = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16);
            const uint64_t hi = (uint64_t)(sbo >> 4) | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
            umma_tf32(tmem, lo | (hi << 32), bdesc + 2 * k, idesc, k != 0);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < 128; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
        for (int j = 0; j < 16; ++j) D[(warp * 32 + (tid & 31)) * 128 + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace dbg
}  // namespace clb

extern "C" int clb_debug_umma_mn(const float* At, const float* B, float* D, int variant, void* stream) {
    using namespace clb;
    CLB_CHECK_ARG(At && B && D && variant >= 0 && variant < 5);
    static bool configured = false;
    if (!configured) {
        CLB_CUDA(cudaFuncSetAttribute(dbg::umma_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
        configured = true;
    }
    dbg::umma_mn_kernel<<<1, 128, 40 * 1024, as_stream(stream)>>>(At, B, D, variant); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

// clb_debug_umma_bf16: D[128x128] = bf16(A[128x64]) * bf16(B[128x64])^T with kind::f16 (K = 16 per MMA, 4 MMAs).
// variant 0: A and B in shared memory (K-major, 64 bf16 = one 128-byte swizzled row); variants 1/2: A through TMEM,
// written with tcgen05.st.32x32b as 32-bit cells holding two consecutive K elements (1: even k in the low half,
// 2: even k in the high half), 8 cells per MMA.  Probes the layout a bf16 hi/lo split (bf16x3) would have to produce.
namespace clb {
namespace dbg {
__device__ __forceinline__ uint32_t bf16_rn_bits(float x) {
    uint32_t u = __float_as_uint(x);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return u >> 16;
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) umma_bf16_kernel(const float* __restrict__ A /*[128][64]*/, const float* __restrict__ B /*[128][64]*/,
                                                           float* __restrict__ D /*[128][128]*/, int variant) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_tile = base, b_tile = base + 16384, bar = base + 32768, slot = bar + 8;
    uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int tid = threadIdx.x, warp = tid >> 5;
    // smem tiles: row r = 64 bf16 = 8 chunks of 16 bytes, chunk c stored at (c ^ (r & 7))
    for (int idx = tid; idx < 128 * 32; idx += 128) {
        const int r = idx >> 5, kp = idx & 31;                        // kp = pair of K elements (2kp, 2kp+1)
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((((kp >> 2) ^ (r & 7)) << 4) + (kp & 3) * 4);
        const uint32_t pa = bf16_rn_bits(A[r * 64 + 2 * kp]) | (bf16_rn_bits(A[r * 64 + 2 * kp + 1]) << 16);
        const uint32_t pb = bf16_rn_bits(B[r * 64 + 2 * kp]) | (bf16_rn_bits(B[r * 64 + 2 * kp + 1]) << 16);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_tile + off), "r"(pa) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(b_tile + off), "r"(pb) : "memory");
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const uint32_t tmem_a = tmem + 128;
    if (variant != 0) {
        uint32_t cells[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t e = bf16_rn_bits(A[tid * 64 + 2 * j]), o = bf16_rn_bits(A[tid * 64 + 2 * j + 1]);
            cells[j] = variant == 1 ? (e | (o << 16)) : (o | (e << 16));
        }
        clb::tcl::tmem_st32(tmem_a + ((uint32_t)(warp * 32) << 16), cells);
        clb::tcl::tmem_wait_st();
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        // instruction descriptor: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t adesc = make_desc(a_tile), bdesc = make_desc(b_tile);
        for (int k = 0; k < 4; ++k) {
            if (variant == 0) umma_bf16_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
            else umma_bf16_ts(tmem, tmem_a + 8 * k, bdesc + 2 * k, idesc, k != 0);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < 128; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
        for (int j = 0; j < 16; ++j) D[(warp * 32 + (tid & 31)) * 128 + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}
}  // namespace dbg
}  // namespace clb

extern "C" int clb_debug_umma_bf16(const float* A, const float* B, float* D, int variant, void* stream) {
    using namespace clb;
    CLB_CHECK_ARG(A && B && D && variant >= 0 && variant < 3);
    static bool configured = false;
    if (!configured) {
        CLB_CUDA(cudaFuncSetAttribute(dbg::umma_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
        configured = true;
    }
    dbg::umma_bf16_kernel<<<1, 128, 40 * 1024, as_stream(stream)>>>(A, B, D, variant); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}

// clb_debug_tma3d: one 3-D TMA box load (optionally 128B-swizzled) at arbitrary (possibly out-of-bounds) coordinates,
// copied back verbatim from shared memory -- probes zero-fill / negative-coordinate behaviour and the swizzle pattern.
#include "../clb_tma.cuh"
namespace clb {
namespace dbg {
__global__ void tma3d_kernel(const __grid_constant__ CUtensorMap map, float* __restrict__ out, int nfloats, int c0, int c1, int c2) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 65536;
    float* tile = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_arrive_expect_tx(bar, (uint32_t)nfloats * 4u);
        tma::load_3d(base, &map, bar, c0, c1, c2);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = tile[i];
}
}  // namespace dbg
}  // namespace clb

extern "C" int clb_debug_tma3d(const float* x, int d0, int d1, int d2, int b0, int b1, int b2, int swizzle, int c0, int c1,
                               int c2, float* out, void* stream) {
    using namespace clb;
    CUtensorMap m;
    const uint64_t dims[3] = {(uint64_t)d0, (uint64_t)d1, (uint64_t)d2};
    const uint64_t str[2] = {(uint64_t)d0 * 4, (uint64_t)d0 * d1 * 4};
    const uint32_t box[3] = {(uint32_t)b0, (uint32_t)b1, (uint32_t)b2};
    int rc = tma::encode_f32(&m, x, 3, dims, str, box, swizzle != 0);
    if (rc) return rc;
    static bool configured = false;
    if (!configured) {
        CLB_CUDA(cudaFuncSetAttribute(dbg::tma3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024));
        configured = true;
    }
    dbg::tma3d_kernel<<<1, 128, 70 * 1024, as_stream(stream)>>>(m, out, b0 * b1 * b2, c0, c1, c2); clb::count_launch();
    CLB_CHECK_LAUNCH();
    return CLB_OK;
}
