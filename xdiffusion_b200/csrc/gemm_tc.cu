// tcgen05 / TMEM / TMA GEMM and implicit-GEMM conv3x3 for sm_100a.
//
//   D[m, n] = sum_k A[m, k] * Wt[n, k]            (A bf16 row-major, Wt bf16 [N, K] row-major)
//
// Persistent kernel: one CTA per SM loops over 128 x BN output tiles (n fastest, so concurrently
// running CTAs share A tiles and all of Wt through L2).  Warp roles (320 threads):
//   warp 0      TMA producer: per 64-wide k-block one A box (128 rows x 128 B) and one B box
//               (BN rows x 128 B) into a STAGES-deep ring (128-byte swizzle, mbarrier complete_tx);
//               the ring runs ahead across tile boundaries
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 UMMA K=16 steps per k-block)
//               into one of TWO TMEM accumulators, so tile i+1 is computed while tile i drains
//   warps 2..9  epilogue, two warps per TMEM lane quadrant (each takes half of the columns):
//               tcgen05.ld 32x32b.x32 -> +bias, activation in registers -> per-warp padded smem
//               transpose -> row-contiguous 16-byte global accesses for gate / residual / output
// The epilogue never touches local memory and every global access is a full 128-byte (fp32) or
// 64-byte (bf16) row segment per 8 lanes.
//
// Implicit conv3x3 (NHWC, pad 1): the A operand of k-block (tap, 64-channel chunk) is a 4-D TMA
// box (64 ch, W, TH rows, TN images) of the activation tensor shifted by the tap offset; TMA's
// out-of-bounds zero fill *is* the padding.  An optional second A segment (1x1, no shift) lets a
// resblock's skip projection accumulate into the same TMEM tile (K = 9*C + Cskip).
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 32 * (2 + EPI_WARPS);
constexpr int STAGE_LD = 36;    // padded row (words) of the per-warp 32x32 transpose buffer

struct TcParams {
    int M, N;
    int nk0, nk1;       // k-blocks taken from A0 / A1
    int conv;           // 0: plain rows, 1: conv3x3 geometry
    int cpb;            // conv: 64-channel chunks per tap
    int H, W;           // conv: image height / width
    int vec_ok;         // unused by the kernel (all accesses are 16-byte; checked on the host)
    int direct_ok;      // bf16 out rows are 16-byte aligned: registers can be stored without the transpose
    int debug;          // experiments only (XDB200_DEBUG): 1 = skip global stores, 2 = skip the MMAs
    Epilogue epi;
};

template <int BN, int CG = 1> struct Cfg {
    static constexpr int B_ROWS = BN / CG;                       // B rows staged by one CTA
    static constexpr int STAGES = (B_ROWS <= 64) ? 6 : (B_ROWS <= 128 ? 5 : (B_ROWS <= 192 ? 4 : 3));
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int XPOSE_BYTES = EPI_WARPS * 32 * STAGE_LD * 4;
    static constexpr int SMEM = STAGES * STAGE_BYTES + XPOSE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);   // power of two
    static constexpr int CHUNKS = BN / 32;
    static constexpr int CHUNKS_PER_GROUP = CHUNKS / 2;
};

// fast activations for the hot epilogue (ex2.approx + rcp.approx; error << bf16 resolution)
template <int ACT> __device__ __forceinline__ float act_fast(float x) {
    if constexpr (ACT == XD_ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));
    if constexpr (ACT == XD_ACT_GELU_TANH) {
        // 0.5 x (1 + tanh(sqrt(2/pi)(x + 0.044715 x^3))) with ONE MUFU op (tanh.approx): the exp + rcp form
        // is MUFU-bound in the fc1 epilogue (2 MUFU x 25 M elements per launch)
        const float u = x * (0.7978845608028654f + 0.035677408136300125f * x * x);
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
        const float hx = 0.5f * x;
        return fmaf(hx, t, hx);
    }
    return x;
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per
// 256 x BN tile: each CTA stages its own 128 rows of A and HALF of the B tile, the leader issues one
// M = 256 MMA that reads both CTAs' shared memory and writes both CTAs' TMEM -- half the shared-memory
// and L2 operand traffic per FLOP, which is what bounds the 1-CTA kernel (see profiles/README.md).
template <int BN, int ACT, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    using C = Cfg<BN, CG>;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* xpose = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::XPOSE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full_bar = empty_bar + C::STAGES;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nk = p.nk0 + p.nk1;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int total_tiles = n_tiles * ((p.M + BM * CG - 1) / (BM * CG));
    const int rank = CG == 2 ? (int)ptx::cluster_ctarank() : 0;     // 0 = leader (issues the MMAs)
    const int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA0);
        ptx::prefetch_tmap(&tmB);
        if (p.nk1) ptx::prefetch_tmap(&tmA1);
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], CG);                // one arrive.expect_tx per CTA of the group
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&tmem_full_bar[b], 1);
            ptx::mbar_init(&tmem_empty_bar[b], EPI_WARPS * CG);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            ptx::tmem_alloc_2sm(tmem_ptr, C::TMEM_COLS);
            ptx::tmem_relinquish_2sm();
        } else {
            ptx::tmem_alloc(tmem_ptr, C::TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync();             // peer barriers initialised before any remote arrive
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped with the tail
    // of the previous kernel; global memory is only touched after this point
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;                                          // ring position, continuous across tiles
            uint32_t ph = 0;
            for (int tile = first_tile; tile < total_tiles && p.debug != 7; tile += tile_step) {
                const int m0 = (tile / n_tiles) * (BM * CG) + rank * BM;     // this CTA's 128 rows of A
                const int n0 = (tile % n_tiles) * BN + rank * C::B_ROWS;     // this CTA's share of the B tile
                int img = 0, h0 = 0;
                if (p.conv) {
                    const int hw = p.H * p.W;
                    img = m0 / hw;
                    h0 = (m0 - img * hw) / p.W;
                }
                for (int kb = 0; kb < nk; ++kb) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sA = smem + s * C::STAGE_BYTES;
                    uint8_t* sB = sA + C::A_BYTES;
                    const CUtensorMap* tm = &tmA0;
                    int c0, c1, c2, c3;
                    if (kb < p.nk0) {
                        if (!p.conv) {
                            c0 = kb * BK; c1 = m0; c2 = 0; c3 = 0;
                        } else {
                            const int tap = kb / p.cpb;
                            const int cc = kb - tap * p.cpb;
                            c0 = cc * BK; c1 = tap % 3 - 1; c2 = h0 + tap / 3 - 1; c3 = img;
                        }
                    } else {
                        tm = &tmA1;
                        const int k1 = kb - p.nk0;
                        if (!p.conv) { c0 = k1 * BK; c1 = m0; c2 = 0; c3 = 0; }
                        else { c0 = k1 * BK; c1 = 0; c2 = h0; c3 = img; }
                    }
                    if (p.debug >= 5) {                          // experiment: no operand traffic at all
                        if constexpr (CG == 2) ptx::mbar_arrive_leader(&full_bar[s]);
                        else ptx::mbar_arrive(&full_bar[s]);
                    } else if constexpr (CG == 2) {
                        ptx::mbar_arrive_expect_tx_leader(&full_bar[s], C::STAGE_BYTES);
                        ptx::tma_load_4d_2sm(sA, tm, &full_bar[s], c0, c1, c2, c3);
                        ptx::tma_load_2d_2sm(sB, &tmB, &full_bar[s], kb * BK, n0);
                    } else {
                        ptx::mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
                        ptx::tma_load_4d(sA, tm, &full_bar[s], c0, c1, c2, c3);
                        ptx::tma_load_2d(sB, &tmB, &full_bar[s], kb * BK, n0);
                    }
                    if (++s == C::STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp runs the (warp-uniform) loop so that descriptors, stage and phase live in
        // uniform registers; one elected lane issues the tcgen05 instructions.
        if (rank == 0) {
            constexpr uint32_t idesc = ptx::idesc_bf16_f32(BM * CG, BN);
            const uint64_t desc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem));
            uint32_t it = 0;
            int s = 0;
            uint32_t ph = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
                const uint32_t buf = it & 1;
                ptx::mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);  // epilogue drained this accumulator
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = 0; kb < nk; ++kb) {
                    if (p.debug != 7) {
                        ptx::mbar_wait(&full_bar[s], ph);
                        ptx::tc_fence_after();
                    }
                    // descriptors differ between stages only in the start-address field (16-byte units)
                    const uint64_t da = desc0 + (uint64_t)((s * C::STAGE_BYTES) >> 4);
                    const uint64_t db = da + (C::A_BYTES >> 4);
                    if (ptx::elect_one()) {
                        if (p.debug != 2 && p.debug != 4) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                // advance 32 bytes (16 bf16) inside the 128-byte swizzle row: +2 in 16-byte units
                                if constexpr (CG == 2) ptx::umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                                else ptx::umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                            }
                        }
                        if constexpr (CG == 2) {
                            if (p.debug != 7) ptx::umma_commit_2sm(&empty_bar[s]);
                            if (kb == nk - 1) ptx::umma_commit_2sm(&tmem_full_bar[buf]);
                        } else {
                            if (p.debug != 7) ptx::umma_commit(&empty_bar[s]);
                            if (kb == nk - 1) ptx::umma_commit(&tmem_full_bar[buf]);
                        }
                    }
                    __syncwarp();
                    if (++s == C::STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        const int e_warp = warp - 2;
        const int grp = e_warp >> 2;                    // which half of the tile's columns
        const int q = warp & 3;                         // TMEM lane quadrant this warp may access
        float* st = xpose + e_warp * 32 * STAGE_LD;     // private 32 x 32 (+pad) transpose buffer
        const Epilogue& e = p.epi;
        const int sub_row = lane >> 3;                  // coalesced pass: 4 rows x 8 lanes x 4 columns
        const int sub_col = (lane & 7) * 4;
        const bool res_f32 = e.res_dtype == XD_F32, out_f32 = e.out_dtype == XD_F32;
        const bool direct_bf16 = C::CHUNKS_PER_GROUP >= 2 && !out_f32 && !e.gate && !e.residual && (p.N % 8 == 0) && p.direct_ok;
        const int res_ld = (int)e.res_ld, out_ld = (int)e.out_ld;     // per-tile row offsets fit 32 bits
        uint32_t it = 0;
        for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
            const int m0 = (tile / n_tiles) * (BM * CG) + rank * BM + q * 32;   // first row of this warp's 32-row band
            const int n0 = (tile % n_tiles) * BN;
            const uint32_t buf = it & 1;
            // per-tile row bookkeeping, hoisted out of the chunk loop (all 32-bit)
            const int rows_left = p.M - m0;                             // rows rr < rows_left are valid
            unsigned gate_row[8];
            if (e.gate) {
#pragma unroll
                for (int i = 0; i < 8; ++i) gate_row[i] = (unsigned)(m0 + i * 4 + sub_row) / (unsigned)e.gate_rows;
            }
            const char* res_band = (const char*)e.residual + (long long)m0 * e.res_ld * (res_f32 ? 4 : 2);
            char* out_band = (char*)e.out + (long long)m0 * e.out_ld * (out_f32 ? 4 : 2);
            ptx::mbar_wait(&tmem_full_bar[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
            if (p.debug >= 3 && p.debug != 6) {         // experiment: mainloop only
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) ptx::mbar_arrive_leader(&tmem_empty_bar[buf]);
                    else ptx::mbar_arrive(&tmem_empty_bar[buf]);
                }
                continue;
            }
#pragma unroll 1
            for (int ci = 0; ci < (C::CHUNKS_PER_GROUP > 0 ? C::CHUNKS_PER_GROUP : 1); ++ci) {
                const int c = grp * C::CHUNKS_PER_GROUP + ci;
                const int nc = n0 + c * 32;             // first global column of this chunk
                // bias for this thread's 32 columns is requested before the TMEM load so that both
                // latencies overlap
                float4 bv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e.bias && nc + 4 * j < p.N) bv[j] = __ldg(reinterpret_cast<const float4*>(e.bias + nc) + j);
                }
                uint32_t r[32];
                ptx::tmem_ld_32x32(t_addr + c * 32, r);
                ptx::tmem_ld_wait();
                if (ci == C::CHUNKS_PER_GROUP - 1) {    // last TMEM read of this tile by this warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) ptx::mbar_arrive_leader(&tmem_empty_bar[buf]);
                        else ptx::mbar_arrive(&tmem_empty_bar[buf]);
                    }
                }
                if (direct_bf16) {
                    // No gate / residual, bf16 out.  Convert in registers, stage the warp's 32 rows x 64
                    // columns as bf16 (128-byte rows, XOR-swizzled 16-byte chunks: conflict-free both ways)
                    // and write full 128-byte lines: 8 lanes x 16 B per row, 4 rows per instruction.
                    uint8_t* sb = reinterpret_cast<uint8_t*>(st);
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        uint4 u;
                        u.x = f2_to_bf2(act_fast<ACT>(__uint_as_float(r[4 * j]) + bv[j].x),
                                        act_fast<ACT>(__uint_as_float(r[4 * j + 1]) + bv[j].y));
                        u.y = f2_to_bf2(act_fast<ACT>(__uint_as_float(r[4 * j + 2]) + bv[j].z),
                                        act_fast<ACT>(__uint_as_float(r[4 * j + 3]) + bv[j].w));
                        u.z = f2_to_bf2(act_fast<ACT>(__uint_as_float(r[4 * j + 4]) + bv[j + 1].x),
                                        act_fast<ACT>(__uint_as_float(r[4 * j + 5]) + bv[j + 1].y));
                        u.w = f2_to_bf2(act_fast<ACT>(__uint_as_float(r[4 * j + 6]) + bv[j + 1].z),
                                        act_fast<ACT>(__uint_as_float(r[4 * j + 7]) + bv[j + 1].w));
                        const int chunk16 = (ci & 1) * 4 + (j >> 1);           // 16-byte chunk inside the 128-byte row
                        *reinterpret_cast<uint4*>(sb + lane * 128 + ((chunk16 ^ (lane & 7)) << 4)) = u;
                    }
                    const bool last = ci == C::CHUNKS_PER_GROUP - 1;
                    if ((ci & 1) || last) {                                    // 64 (or a trailing 32) columns staged: flush
                        __syncwarp();
                        const int staged = (ci & 1) ? 64 : 32;
                        const int col = nc - (staged - 32) + (lane & 7) * 8;
                        const bool lane_ok = (lane & 7) * 8 < staged && col < p.N;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = i * 4 + sub_row;
                            const uint4 u = *reinterpret_cast<const uint4*>(sb + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
                            if (rr < rows_left && lane_ok && p.debug != 1)
                                *reinterpret_cast<uint4*>(out_band + (unsigned)(rr * out_ld + col) * 2u) = u;
                        }
                        __syncwarp();
                    }
                    continue;
                }
                // ---- column-wise part on the row-per-thread fragment: bias, activation
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v;
                    v.x = act_fast<ACT>(__uint_as_float(r[4 * j]) + bv[j].x);
                    v.y = act_fast<ACT>(__uint_as_float(r[4 * j + 1]) + bv[j].y);
                    v.z = act_fast<ACT>(__uint_as_float(r[4 * j + 2]) + bv[j].z);
                    v.w = act_fast<ACT>(__uint_as_float(r[4 * j + 3]) + bv[j].w);
                    *reinterpret_cast<float4*>(st + lane * STAGE_LD + 4 * j) = v;
                }
                __syncwarp();
                // ---- row-contiguous part: gate, residual, store (8 lanes cover one 32-column row segment)
                const int col = nc + sub_col;
                const bool col_ok = col < p.N;
                float4 v[8], g[8], rs[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + sub_row;
                    const bool ok = col_ok && rr < rows_left;
                    v[i] = *reinterpret_cast<const float4*>(st + rr * STAGE_LD + sub_col);
                    g[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                    rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok && e.gate) g[i] = __ldg(reinterpret_cast<const float4*>(e.gate + (long long)gate_row[i] * e.gate_ld + col));
                    if (ok && e.residual) {
                        if (res_f32) {
                            rs[i] = *reinterpret_cast<const float4*>(res_band + (unsigned)(rr * res_ld + col) * 4u);
                        } else {
                            const uint2 u = *reinterpret_cast<const uint2*>(res_band + (unsigned)(rr * res_ld + col) * 2u);
                            const float2 lo = bf2_to_f2(u.x), hi = bf2_to_f2(u.y);
                            rs[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + sub_row;
                    if (!(col_ok && rr < rows_left)) continue;
                    float4 o;
                    o.x = fmaf(v[i].x, g[i].x, rs[i].x); o.y = fmaf(v[i].y, g[i].y, rs[i].y);
                    o.z = fmaf(v[i].z, g[i].z, rs[i].z); o.w = fmaf(v[i].w, g[i].w, rs[i].w);
                    if (out_f32) {
                        *reinterpret_cast<float4*>(out_band + (unsigned)(rr * out_ld + col) * 4u) = o;
                    } else {
                        *reinterpret_cast<uint2*>(out_band + (unsigned)(rr * out_ld + col) * 2u) =
                            make_uint2(f2_to_bf2(o.x, o.y), f2_to_bf2(o.z, o.w));
                    }
                }
                __syncwarp();                           // transpose buffer is reused by the next chunk
            }
        }
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync();             // no remote arrive / peer smem read after a CTA exits
    else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

// bf16 tensor map, inner box = 64 elements (128 B), 128-byte swizzle, zero OOB fill.
int make_tmap(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { xd_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled entry point not found"); return XD_ERR_TMAP; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims,
                     strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static char msg[160];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        xd_set_error(__FILE__, __LINE__, msg);
        return XD_ERR_TMAP;
    }
    return XD_OK;
}

int tmap_rows(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    // rank-4 view (cols, rows, 1, 1) so that the plain GEMM shares the conv kernel's 4-D load.
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, 1, 1};
    cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rows, (cuuint64_t)ld * 2 * (cuuint64_t)rows};
    cuuint32_t box[4] = {BK, (cuuint32_t)box_rows, 1, 1};
    return make_tmap(tm, ptr, 4, dims, str, box);
}

int tmap_weights(CUtensorMap* tm, const void* ptr, long long n, long long k, long long ld, int bn) {
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)n};
    cuuint64_t str[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)bn};
    return make_tmap(tm, ptr, 2, dims, str, box);
}

int tmap_nhwc(CUtensorMap* tm, const void* ptr, int nimg, int H, int W, int C, long long ld) {
    // ld = elements between consecutive pixels (>= C: the tensor may be a channel slice of a wider buffer)
    const int hw = H * W;
    const int th = hw >= BM ? BM / W : H;
    const int tn = hw >= BM ? 1 : BM / hw;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
    cuuint32_t box[4] = {BK, (cuuint32_t)W, (cuuint32_t)th, (cuuint32_t)tn};
    return make_tmap(tm, ptr, 4, dims, str, box);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// The kernel's epilogue uses 16-byte (fp32) / 8-byte (bf16) accesses on 4-column groups.
int epilogue_vec_ok(const Epilogue& e, int N) {
    bool ok = aligned16(e.out) && aligned16(e.bias) && aligned16(e.gate) && aligned16(e.residual);
    ok = ok && (N % 4 == 0) && (e.out_ld % 4 == 0) && (e.res_ld % 4 == 0) && (e.gate_ld % 4 == 0);
    ok = ok && e.out_ld < (1 << 22) && e.res_ld < (1 << 22);      // 32-bit byte offsets inside a 128-row tile
    return ok ? 1 : 0;
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN, int ACT, int CG>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcParams& p, cudaStream_t st) {
    static bool configured = false;
    auto kernel = gemm_tc_kernel<BN, ACT, CG>;
    if (!configured) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, CG>::SMEM) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute(max dynamic smem) failed");
            return XD_ERR_CUDA;
        }
        configured = true;
    }
    const long long tiles = (long long)((p.N + BN - 1) / BN) * ((p.M + BM * CG - 1) / (BM * CG));
    // persistent: <= one CTA (pair) per SM (pair)
    const unsigned grid = (unsigned)std::min<long long>(tiles, sm_count() / CG) * CG;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg<BN, CG>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = xd_pdl_enabled() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, kernel, a0, a1, b, p) != cudaSuccess) {
        xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
        return XD_ERR_CUDA;
    }
    return XD_OK;
}

// Tile shape: (bn, cg).  force: 0 = auto; 64/128/256 = 1-CTA tiles; 1128/1256 = CTA-pair 256 x 128 / 256 x 256.
void pick_tile(int N, int M, int force, int* bn, int* cg) {
    if (force >= 1000) { *cg = 2; *bn = force - 1000; return; }
    if (force) { *cg = 1; *bn = force; return; }
    // tcgen05.mma 128 x N x 16 from shared memory runs at ~half rate for N = 128 (operand reads
    // saturate the shared-memory port) and at ~75% of peak for N >= 192 (measured, profiles/README.md),
    // so take the widest tile that divides N; a CTA pair for the N = 128 leftovers.
    static const int mode = getenv("XDB200_CG") ? atoi(getenv("XDB200_CG")) : 1;
    *cg = 1;
    if (N <= 64) { *bn = 64; return; }
    const long long m_tiles = (M + BM - 1) / BM;
    auto fills = [&](int w) { return m_tiles * ((N + w - 1) / w) >= sm_count(); };   // at least one full wave
    if (N % 192 == 0 && fills(192)) *bn = 192;          // measured: 192 beats 256 when both divide N
    else if (N % 256 == 0 && fills(256)) *bn = 256;
    else if (N > 1024 && fills(256)) *bn = 256;
    else {
        *bn = 128;
        if (mode == 2 && M > BM) *cg = 2;
    }
}

int dispatch(int bn, int cg, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcParams& p,
             cudaStream_t st) {
#define XD_TC_CASE(BN_, CG_)                                                                    \
    if (bn == BN_ && cg == CG_) {                                                               \
        switch (p.epi.act) {                                                                    \
            case XD_ACT_NONE: return launch<BN_, XD_ACT_NONE, CG_>(a0, a1, b, p, st);           \
            case XD_ACT_SILU: return launch<BN_, XD_ACT_SILU, CG_>(a0, a1, b, p, st);           \
            case XD_ACT_GELU_TANH: return launch<BN_, XD_ACT_GELU_TANH, CG_>(a0, a1, b, p, st); \
        }                                                                                       \
    }
    XD_TC_CASE(64, 1)
    XD_TC_CASE(128, 1)
    XD_TC_CASE(192, 1)
    XD_TC_CASE(256, 1)
    XD_TC_CASE(128, 2)
    XD_TC_CASE(256, 2)
#undef XD_TC_CASE
    xd_set_error(__FILE__, __LINE__, "unsupported tile shape");
    return XD_ERR_ARG;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI (declared in include/xdb200.h)
// ---------------------------------------------------------------------------------------------
extern "C" int xd_gemm_bf16_tc(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                               long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                               int gate_rows, long long gate_ld, const void* residual, int res_dtype,
                               long long res_ld, void* out, int out_dtype, long long out_ld, int force_bn,
                               void* stream) {
    XD_CHECK_ARG(A && Wt && out && M > 0 && N > 0 && K > 0);
    XD_CHECK_ARG(K % BK == 0 && K2 % BK == 0 && (A2 != nullptr) == (K2 > 0));
    XD_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda2 % 8 == 0 && aligned16(A) && aligned16(Wt) && aligned16(A2));
    XD_CHECK_ARG(!gate || gate_rows > 0);
    TcParams p{};
    p.M = M; p.N = N; p.nk0 = K / BK; p.nk1 = K2 / BK; p.conv = 0;
    p.epi = Epilogue{bias, gate, residual, out, gate_ld, res_ld, out_ld, act, gate_rows, res_dtype, out_dtype};
    p.vec_ok = epilogue_vec_ok(p.epi, N);
    XD_CHECK_ARG(p.vec_ok);
    p.direct_ok = (out_ld % 8 == 0) && getenv("XDB200_NO_DIRECT") == nullptr;
    p.debug = getenv("XDB200_DEBUG") ? atoi(getenv("XDB200_DEBUG")) : 0;
    int bn, cg;
    pick_tile(N, M, force_bn, &bn, &cg);
    CUtensorMap ta0, ta1, tb;
    int rc;
    if ((rc = tmap_rows(&ta0, A, M, K, lda, BM))) return rc;
    ta1 = ta0;
    if (A2 && (rc = tmap_rows(&ta1, A2, M, K2, lda2, BM))) return rc;
    if ((rc = tmap_weights(&tb, Wt, N, K + K2, ldw, bn / cg))) return rc;
    return dispatch(bn, cg, ta0, ta1, tb, p, (cudaStream_t)stream);
}

extern "C" int xd_conv3x3_bf16_tc(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                                  long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                                  const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                                  long long out_ld, int force_bn, void* stream) {
    XD_CHECK_ARG(X && Wp && out && nimg > 0 && H > 0 && W > 0 && Cout > 0);
    XD_CHECK_ARG(C % BK == 0 && Cs % BK == 0 && (Xs != nullptr) == (Cs > 0));
    XD_CHECK_ARG(ldx % 8 == 0 && lds % 8 == 0 && aligned16(X) && aligned16(Xs) && aligned16(Wp));
    const int hw = H * W;
    // an M tile is 128 consecutive NHWC pixels: whole rows of one image, or whole images
    XD_CHECK_ARG(W <= BM && BM % W == 0 && (hw % BM == 0 || BM % hw == 0));
    TcParams p{};
    p.M = nimg * hw; p.N = Cout; p.conv = 1; p.cpb = C / BK; p.nk0 = 9 * p.cpb; p.nk1 = Cs / BK; p.H = H; p.W = W;
    p.epi = Epilogue{bias, nullptr, residual, out, 0, res_ld, out_ld, act, 1, res_dtype, out_dtype};
    p.vec_ok = epilogue_vec_ok(p.epi, Cout);
    XD_CHECK_ARG(p.vec_ok);
    p.direct_ok = (out_ld % 8 == 0) && getenv("XDB200_NO_DIRECT") == nullptr;
    p.debug = getenv("XDB200_DEBUG") ? atoi(getenv("XDB200_DEBUG")) : 0;
    int bn, cg;
    pick_tile(Cout, p.M, force_bn, &bn, &cg);
    CUtensorMap ta0, ta1, tb;
    int rc;
    if ((rc = tmap_nhwc(&ta0, X, nimg, H, W, C, ldx))) return rc;
    ta1 = ta0;
    if (Xs && (rc = tmap_nhwc(&ta1, Xs, nimg, H, W, Cs, lds))) return rc;
    const long long ktot = 9LL * C + Cs;
    if ((rc = tmap_weights(&tb, Wp, Cout, ktot, ktot, bn / cg))) return rc;
    return dispatch(bn, cg, ta0, ta1, tb, p, (cudaStream_t)stream);
}
