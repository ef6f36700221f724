// tcgen05 fused attention for T = 256 keys/queries per (batch, head), head dim 64 (the UNet's 16x16
// attention level: 5 of its 6 attention blocks).
//
// One CTA = 128 query rows of one (batch, head).  160 threads:
//   warp 0      TMA loads of Q (128x64), K (256x64), V (256x64) with 128-byte swizzle; TMEM alloc;
//               single-thread tcgen05.mma:  S[128x256] = Q K^T  (both operands K-major), then after
//               the softmax warps have written P:  O[128x64] = P V  with P (bf16, K-major, written to
//               shared memory in the swizzled UMMA layout) and V as an MN-major B operand (the [key][d]
//               tile exactly as TMA delivers it)
//   warps 1..4  one thread per query row = TMEM lane: tcgen05.ld of the 256 logits (two passes: max,
//               then exp2 / sum / bf16 pack into smem), later the 64 output columns, scaled by 1/sum
// The whole 256-key row is resident in TMEM, so the softmax is exact (no online rescaling).
//
// Two CTAs per SM (round 2): a CTA is a serial chain (TMA round trip -> S -> two TMEM passes of softmax -> P V -> store) that
// keeps the tensor pipe busy for ~10 % of its life; with 145 KB of shared memory and a 512-column TMEM allocation only one
// fitted an SM.  P (64 KB, written after the last S MMA has read Q and K) now aliases the Q and K tiles, O (64 columns)
// aliases the first S columns (the P V MMAs are issued after every softmax thread has read S), so a CTA needs 96 KB and
// 256 TMEM columns and two of them overlap each other's latencies.
#include "common.cuh"
#include "ptx.cuh"

#include <mutex>
#include <stdio.h>

namespace attn_tc {

constexpr int T = 256, D = 64, BM = 128;
constexpr int SQ = BM * 128, SK = T * 128, SV = T * 128, SP = 4 * BM * 128;     // bytes
static_assert(SQ + SK <= SP, "P aliases the Q and K tiles");
constexpr int SMEM = SP + SV + 1024 + 128;
constexpr int THREADS = 160;

struct Params {
    bf16* o;
    long long o_bs, o_hs, o_rs;
    int H;
    int q_hs, k_hs, v_hs;      // head strides (elements) = TMA column coordinate step
    float scale_log2e;
};

__device__ __forceinline__ uint32_t idesc(int M, int N, bool b_mn_major) {
    return ptx::idesc_bf16_f32(M, N) | (b_mn_major ? (1u << 16) : 0u);
}

__global__ void __launch_bounds__(THREADS, 2)
attention_tc256_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const Params p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sP = smem;                                  // [4 k-blocks of 64 keys][128 rows][128 B]; aliases Q and K
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + SQ;
    uint8_t* sV = smem + SP;
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(sV + SV);
    uint64_t* bar_s = bar_load + 1;
    uint64_t* bar_p = bar_load + 2;
    uint64_t* bar_o = bar_load + 3;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_load + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slab = blockIdx.x & 1;
    const int bh = blockIdx.x >> 1;
    const int b = bh / p.H, h = bh % p.H;

    if (warp == 0) {
        if (lane == 0) {
            ptx::prefetch_tmap(&tmQ);
            ptx::prefetch_tmap(&tmK);
            ptx::prefetch_tmap(&tmV);
            ptx::mbar_init(bar_load, 1);
            ptx::mbar_init(bar_s, 1);
            ptx::mbar_init(bar_p, 128);
            ptx::mbar_init(bar_o, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_ptr, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t tmem_s = tmem, tmem_o = tmem;         // O re-uses the first 64 S columns (see the header)
    pdl_wait();                                          // prologue above overlaps the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            const int row0 = b * T;
            ptx::mbar_arrive_expect_tx(bar_load, SQ + SK + SV);
            ptx::tma_load_2d(sQ, &tmQ, bar_load, h * p.q_hs, row0 + slab * BM);
            ptx::tma_load_2d(sK, &tmK, bar_load, h * p.k_hs, row0);
            ptx::tma_load_2d(sK + SK / 2, &tmK, bar_load, h * p.k_hs, row0 + 128);
            ptx::tma_load_2d(sV, &tmV, bar_load, h * p.v_hs, row0);
            ptx::tma_load_2d(sV + SV / 2, &tmV, bar_load, h * p.v_hs, row0 + 128);
            ptx::mbar_wait(bar_load, 0);
            ptx::tc_fence_after();
            // S = Q K^T : M=128, N=256, K=64 (4 x UMMA_K)
            const uint64_t dq = ptx::smem_desc_sw128(ptx::smem_u32(sQ));
            const uint64_t dk = ptx::smem_desc_sw128(ptx::smem_u32(sK));
            const uint32_t id1 = idesc(BM, T, false);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, id1, k ? 1u : 0u);
            ptx::umma_commit(bar_s);
            // O = P V : M=128, N=64, K=256 (16 x UMMA_K); V is MN-major: 16 keys = 16 rows of 128 B
            ptx::mbar_wait(bar_p, 0);
            ptx::tc_fence_after();
            const uint32_t id2 = idesc(BM, D, true);
            const uint32_t p_addr = ptx::smem_u32(sP), v_addr = ptx::smem_u32(sV);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                const uint64_t dp = ptx::smem_desc_sw128(p_addr + (kk >> 2) * (BM * 128) + (kk & 3) * 32);
                const uint64_t dv = ptx::smem_desc_sw128(v_addr + kk * 16 * 128);
                ptx::umma_bf16(tmem_o, dp, dv, id2, kk ? 1u : 0u);
            }
            ptx::umma_commit(bar_o);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                          // TMEM lane quadrant of this warp
        const int r = q * 32 + lane;                     // query row inside the tile
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        ptx::mbar_wait(bar_s, 0);
        ptx::tc_fence_after();
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
        const float mb = mx * p.scale_log2e;
        float l = 0.f;
        uint8_t* prow = sP + r * 128;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {                    // 32 keys per chunk; 64 keys per 16 KB k-block
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
            ptx::tmem_ld_wait();
            uint8_t* blk = prow + (c >> 1) * (BM * 128);
#pragma unroll
            for (int g = 0; g < 4; ++g) {                // 8 keys = one 16-byte chunk
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    e[j] = exp2f(fmaf(__uint_as_float(v[g * 8 + j]), p.scale_log2e, -mb));
                    l += e[j];
                }
                const int chunk = (c & 1) * 4 + g;       // 16-byte chunk index inside the 128-byte row
                *reinterpret_cast<bf16x8*>(blk + ((chunk ^ (r & 7)) << 4)) = pack8(e);
            }
        }
        ptx::fence_proxy_async();                        // generic-proxy smem writes -> visible to the MMA
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_p);
        ptx::mbar_wait(bar_o, 0);
        ptx::tc_fence_after();
        const float inv = 1.0f / l;
        bf16* op = p.o + b * p.o_bs + h * p.o_hs + (long long)(slab * BM + r) * p.o_rs;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_o + lane_addr + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) e[j] = __uint_as_float(v[g * 8 + j]) * inv;
                *reinterpret_cast<bf16x8*>(op + c * 32 + g * 8) = pack8(e);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 256);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    return fn;
}

static int make_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long rs) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { xd_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled entry point not found"); return XD_ERR_TMAP; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)rs * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, str, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { xd_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed (attention)"); return XD_ERR_TMAP; }
    return XD_OK;
}

}  // namespace attn_tc

// Returns XD_OK when the tensor-core path took the call, -1 when the layout does not qualify (the
// caller then uses the CUDA-core kernel), or an error code.
int xd_attention_tc256_try(const void* q, long long q_bs, long long q_hs, long long q_rs, const void* k,
                           long long k_bs, long long k_hs, long long k_rs, const void* v, long long v_bs,
                           long long v_hs, long long v_rs, void* o, long long o_bs, long long o_hs, long long o_rs,
                           int B, int H, float scale, cudaStream_t st) {
    using namespace attn_tc;
    auto ok = [&](const void* ptr, long long bs, long long hs, long long rs) {
        return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && bs == (long long)T * rs && rs % 8 == 0 && hs % 8 == 0 &&
               (H - 1) * hs + D <= rs;
    };
    if (!ok(q, q_bs, q_hs, q_rs) || !ok(k, k_bs, k_hs, k_rs) || !ok(v, v_bs, v_hs, v_rs)) return -1;
    if (o_rs % 8 || o_hs % 8 || o_bs % 8 || (reinterpret_cast<uintptr_t>(o) & 15)) return -1;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_map(&tq, q, (long long)B * T, (H - 1) * q_hs + D, q_rs))) return rc;
    if ((rc = make_map(&tk, k, (long long)B * T, (H - 1) * k_hs + D, k_rs))) return rc;
    if ((rc = make_map(&tv, v, (long long)B * T, (H - 1) * v_hs + D, v_rs))) return rc;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(attention_tc256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute failed (attention_tc256)");
            return XD_ERR_CUDA;
        }
        configured = true;
    }
    Params p{(bf16*)o, o_bs, o_hs, o_rs, H, (int)q_hs, (int)k_hs, (int)v_hs, scale * 1.4426950408889634f};
    xd_launch(attention_tc256_kernel, (unsigned)(B * H * 2), THREADS, SMEM, st, tq, tk, tv, p);
    XD_CHECK_LAUNCH();
    return XD_OK;
}
