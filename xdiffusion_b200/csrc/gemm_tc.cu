// tcgen05 / TMEM / TMA GEMM and implicit-GEMM conv3x3 for sm_100a.
//
//   D[m, n] = sum_k A[m, k] * Wt[n, k]            (A bf16 row-major, Wt bf16 [N, K] row-major)
//
// Persistent kernel gemm_tc_kernel<BN, ACT, CG, AS, EPI, LNA>: one CTA per SM (CG = 1) or one CTA pair per SM pair
// (CG = 2, tcgen05 cta_group::2, 256 x BN tiles) loops over output tiles, n fastest (concurrently running CTAs share A
// tiles and all of Wt through L2).  Warp roles (320 threads):
//   warp 0      TMA producer: per 64-wide k-block one A box (128 rows x 128 B) and one B box (BN / CG rows x 128 B)
//               into a STAGES-deep ring (128-byte swizzle, mbarrier complete_tx); the ring runs ahead across tiles
//   warp 1      TMEM allocator + MMA issuer (warp-uniform loop, one elected lane; 4 UMMA K = 16 steps per k-block)
//               into one of TWO TMEM accumulators, so tile i + 1 is computed while tile i drains
//   warps 2..9  epilogue, two warps per TMEM lane quadrant (each takes half of the columns), 32-column chunks with
//               double-buffered tcgen05.ld.  EPI_TMA_F32 / EPI_TMA_BF16 (default): work in the TMEM-native layout
//               (lane = row), residual box in by bulk tensor load, result out by bulk tensor store, bias slice staged in
//               shared memory, no predicates (TMA clips the edges).  EPI_LEGACY: register epilogue with a padded
//               shared-memory transpose and 16-byte global accesses (unaligned rows, mixed residual dtype).
// Options: AS = A-stationary work items (the 128 x K panel of A stays in shared memory across the n-tiles of an item),
// LNA = the epilogue warps build that panel as LayerNorm(x) * (1 + scale) + shift from fp32 rows, split-K (host side:
// fp32 partial tiles to a workspace + splitk_reduce_kernel), programmatic dependent launch.
//
// Implicit conv3x3 (NHWC, pad 1): the A operand of k-block (tap, 64-channel chunk) is a 4-D TMA
// box (64 ch, W, TH rows, TN images) of the activation tensor shifted by the tap offset; TMA's
// out-of-bounds zero fill *is* the padding.  An optional second A segment (1x1, no shift) lets a
// resblock's skip projection accumulate into the same TMEM tile (K = 9*C + Cskip).
//
// What bounds it (profiles/README.md): ~5 us fixed per launch; the K-loop is limited by operand delivery (TMA sustains
// ~64 B/clk/SM, tcgen05.mma operand reads share the 128 B/clk shared-memory port), not by the tensor pipe.
#include "common.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"

#include <algorithm>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

using namespace tcx;

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
    int tma_epi;        // 1: TMA epilogue (bulk tensor load of the residual, bulk tensor store of the output)
    // LNA kernels: the A operand is LayerNorm(x) * (1 + scale) + shift computed in the kernel from fp32 rows
    const float* ln_x;
    const float* ln_shift;
    const float* ln_scale;
    long long ln_ld, ln_mod_ld;
    int ln_rows_per_mod;
    float ln_eps;
    long long* prof;    // XDB200_PROF=1: per-CTA cycle counters (16 slots per CTA), nullptr otherwise
    // QS kernels: the epilogue also emits, per 32-row block and per group of 4 output columns ("quad"), the sum and the sum
    // of squares of the values it stores: qstats[(m / 32) * qstats_ld + (n / 4) * 2 + {0, 1}].  The GroupNorm that consumes
    // the output then needs no statistics pass of its own (norm.cu: gn_apply_quads_kernel).
    float* qstats;
    long long qstats_ld;
    int ksplit, kb_per_split, ws_rows;   // split-K: item = (tile, split); split s covers k-blocks [s * kb_per_split, ...) and
                                         // stores its fp32 partial tile at rows s * ws_rows + m of the workspace (out)
    int ng, tpg;        // work items: every m-tile is split into ng groups of tpg consecutive n-tiles
    Epilogue epi;
};

// AS ("A-stationary"): the 128 x K panel of A (K <= A_SLOTS * 64) stays in shared memory while the CTA walks the
// n-tiles of its work item; only B is streamed through the ring.  Operand traffic L2 -> SM per 128 x BN tile
// drops from (128 + BN) * K to BN * K (BN/2 * K for a CTA pair), which is what bounds the K = 384 DiT
// contractions (profiles/README.md).
template <int BN, int CG = 1, bool AS = false> struct Cfg {
    static constexpr int B_ROWS = BN / CG;                       // B rows staged by one CTA
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int A_SLOTS = AS ? 6 : 0;
    static constexpr int PANEL_BYTES = A_SLOTS * A_BYTES;
    static constexpr int STAGE_BYTES = AS ? B_BYTES : A_BYTES + B_BYTES;
    // per epilogue warp: 8 KB of staging (two 32 x 32 output boxes for the TMA epilogue; the legacy path uses the
    // first 4.5 KB as its padded transpose buffer) + 512 B for the warp's slice of the bias vector
    static constexpr int EPI_WARP_BYTES = 8192;
    static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
    static constexpr int BIAS_BYTES = EPI_WARPS * 512;
    static constexpr int FIXED_BYTES = EPI_BYTES + BIAS_BYTES + 1024 /*align*/ + 512 /*barriers*/;
    static constexpr int FIT = (232448 - FIXED_BYTES - PANEL_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT > (AS ? 8 : 6) ? (AS ? 8 : 6) : FIT;
    static constexpr int SMEM = PANEL_BYTES + STAGES * STAGE_BYTES + FIXED_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);   // power of two
    static constexpr int CHUNKS = BN / 32;
    static constexpr int CHUNKS_PER_GROUP = CHUNKS / 2;
    static_assert(STAGES >= 2 && STAGES <= 8, "ring depth");
};

// Experiments (XDB200_DEBUG knobs) and the in-kernel cycle accounting (XDB200_PROF) are compiled only with
// -DXDB200_INSTRUMENT (NVCC_EXTRA=-DXDB200_INSTRUMENT csrc/build.sh): they cost code size in the hot loops.
#ifdef XDB200_INSTRUMENT
constexpr bool kInst = true;
#else
constexpr bool kInst = false;
#endif
enum { EPI_LEGACY = 0, EPI_TMA_F32 = 1, EPI_TMA_BF16 = 2 };

// mbarrier wait that (optionally) accounts the waiting time; profiling builds of the wait are only taken when
// XDB200_PROF is set, the branch is warp-uniform.
__device__ __forceinline__ void wait_acc(uint64_t* bar, uint32_t parity, bool prof, long long& acc) {
    if (prof) {
        const long long t = clock64();
        ptx::mbar_wait(bar, parity);
        acc += clock64() - t;
    } else {
        ptx::mbar_wait(bar, parity);
    }
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per
// 256 x BN tile: each CTA stages its own 128 rows of A and HALF of the B tile, the leader issues one
// M = 256 MMA that reads both CTAs' shared memory and writes both CTAs' TMEM -- half the shared-memory
// and L2 operand traffic per FLOP, which is what bounds the 1-CTA kernel (see profiles/README.md).
template <int BN, int ACT, int CG, bool AS, int EPI, bool LNA = false, bool QS = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmRes,
               const __grid_constant__ CUtensorMap tmOut, const TcParams p) {
    using C = Cfg<BN, CG, AS>;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* panel = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem = panel + C::PANEL_BYTES;              // the ring
    uint8_t* epi = smem + C::STAGES * C::STAGE_BYTES;    // 1024-aligned (all stage sizes are multiples of 1024)
    uint8_t* bias_sm = epi + C::EPI_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_sm + C::BIAS_BYTES);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full_bar = empty_bar + 8;             // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
    uint64_t* a_full_bar = tmem_empty_bar + 2;           // [6]  (AS only)
    uint64_t* a_empty_bar = a_full_bar + 6;              // [6]
    uint64_t* res_bar = a_empty_bar + 6;                 // [EPI_WARPS][2]  (TMA epilogue: residual box landed)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 2 * EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool prof = kInst && p.prof != nullptr;
    const int dbg = kInst ? p.debug : 0;
    long long* pslot = prof ? p.prof + 16 * blockIdx.x : nullptr;
    if (prof && threadIdx.x == 0) pslot[8] = (long long)globaltimer_ns();
    const int nk_total = p.nk0 + p.nk1;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int total_items = p.ng * ((p.M + BM * CG - 1) / (BM * CG)) * p.ksplit;
    const int rank = CG == 2 ? (int)ptx::cluster_ctarank() : 0;     // 0 = leader (issues the MMAs)
    const int first_item = blockIdx.x / CG, item_step = gridDim.x / CG;
// work item w = (m-tile, group of tpg consecutive n-tiles); without AS every item is one tile (n fastest)
#define XD_ITEM_LOOP for (int w_ = first_item; w_ < total_items; w_ += item_step, ++item)
#define XD_ITEM_DECODE                                                              \
    const int sp = w_ % p.ksplit;                       /* split fastest: the splits of a tile run side by side */ \
    const int w2_ = w_ / p.ksplit;                                                  \
    const int mt = w2_ / p.ng;                                                      \
    const int nt0 = (w2_ - mt * p.ng) * p.tpg;                                      \
    const int nt1 = min(n_tiles, nt0 + p.tpg);                                      \
    const int kb0 = sp * p.kb_per_split;                                            \
    const int nk = min(p.kb_per_split, nk_total - kb0);
#define XD_TILE_LOOP for (int nt = nt0; nt < nt1; ++nt, ++it)

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA0);
        ptx::prefetch_tmap(&tmB);
        if (p.nk1) ptx::prefetch_tmap(&tmA1);
        if (EPI != EPI_LEGACY) {
            ptx::prefetch_tmap(&tmOut);
            if (p.epi.residual) ptx::prefetch_tmap(&tmRes);
        }
        for (int k = 0; k < 2 * EPI_WARPS; ++k) ptx::mbar_init(&res_bar[k], 1);
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], CG);                // one arrive.expect_tx per CTA of the group
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&tmem_full_bar[b], 1);
            ptx::mbar_init(&tmem_empty_bar[b], EPI_WARPS * CG);
        }
        for (int k = 0; k < C::A_SLOTS; ++k) {
            ptx::mbar_init(&a_full_bar[k], LNA ? EPI_WARPS * CG : CG);   // LNA: one arrive per transforming warp
            ptx::mbar_init(&a_empty_bar[k], 1);
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
            uint32_t ph = 0, it = 0, item = 0;
            long long w_empty = 0;
            const long long t_begin = prof ? clock64() : 0;
            if (dbg != 7 && dbg != 9) XD_ITEM_LOOP { XD_ITEM_DECODE XD_TILE_LOOP {
                const int m0 = mt * (BM * CG) + rank * BM;                   // this CTA's 128 rows of A
                const int n0 = nt * BN + rank * C::B_ROWS;                   // this CTA's share of the B tile
                int img = 0, h0 = 0;
                if (p.conv) {
                    const int hw = p.H * p.W;
                    img = m0 / hw;
                    h0 = (m0 - img * hw) / p.W;
                }
                for (int kb = 0; kb < nk; ++kb) {
                    const bool load_a = !LNA && (!AS || nt == nt0);
                    uint64_t* bar_a = AS ? &a_full_bar[kb] : &full_bar[s];
                    uint8_t* sA = AS ? panel + kb * C::A_BYTES : smem + s * C::STAGE_BYTES;
                    uint8_t* sB = AS ? smem + s * C::STAGE_BYTES : sA + C::A_BYTES;
                    if (AS && load_a) ptx::mbar_wait(&a_empty_bar[kb], (item & 1) ^ 1);
                    wait_acc(&empty_bar[s], ph ^ 1, prof, w_empty);
                    const CUtensorMap* tm = &tmA0;
                    int c0, c1, c2, c3;
                    const int kg = kb0 + kb;                     // k-block index in the whole contraction
                    if (kg < p.nk0) {
                        if (!p.conv) {
                            c0 = kg * BK; c1 = m0; c2 = 0; c3 = 0;
                        } else {
                            const int tap = kg / p.cpb;
                            const int cc = kg - tap * p.cpb;
                            c0 = cc * BK; c1 = tap % 3 - 1; c2 = h0 + tap / 3 - 1; c3 = img;
                        }
                    } else {
                        tm = &tmA1;
                        const int k1 = kg - p.nk0;
                        if (!p.conv) { c0 = k1 * BK; c1 = m0; c2 = 0; c3 = 0; }
                        else { c0 = k1 * BK; c1 = 0; c2 = h0; c3 = img; }
                    }
                    if (dbg >= 5) {                          // experiment: no operand traffic at all
                        if constexpr (CG == 2) {
                            if (AS && load_a) ptx::mbar_arrive_leader(bar_a);
                            ptx::mbar_arrive_leader(&full_bar[s]);
                        } else {
                            if (AS && load_a) ptx::mbar_arrive(bar_a);
                            ptx::mbar_arrive(&full_bar[s]);
                        }
                    } else if constexpr (CG == 2) {
                        if constexpr (AS) {
                            if (load_a) {
                                ptx::mbar_arrive_expect_tx_leader(bar_a, C::A_BYTES);
                                ptx::tma_load_4d_2sm(sA, tm, bar_a, c0, c1, c2, c3);
                            }
                            ptx::mbar_arrive_expect_tx_leader(&full_bar[s], C::B_BYTES);
                        } else {
                            ptx::mbar_arrive_expect_tx_leader(&full_bar[s], C::STAGE_BYTES);
                            ptx::tma_load_4d_2sm(sA, tm, bar_a, c0, c1, c2, c3);
                        }
                        ptx::tma_load_2d_2sm(sB, &tmB, &full_bar[s], kg * BK, n0);
                    } else {
                        if constexpr (AS) {
                            if (load_a) {
                                ptx::mbar_arrive_expect_tx(bar_a, C::A_BYTES);
                                ptx::tma_load_4d(sA, tm, bar_a, c0, c1, c2, c3);
                            }
                            ptx::mbar_arrive_expect_tx(&full_bar[s], C::B_BYTES);
                        } else {
                            ptx::mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
                            ptx::tma_load_4d(sA, tm, bar_a, c0, c1, c2, c3);
                        }
                        ptx::tma_load_2d(sB, &tmB, &full_bar[s], kg * BK, n0);
                    }
                    if (++s == C::STAGES) { s = 0; ph ^= 1; }
                }
            }}
            if (prof) { pslot[3] = clock64() - t_begin; pslot[4] = w_empty; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp runs the (warp-uniform) loop so that descriptors, stage and phase live in
        // uniform registers; one elected lane issues the tcgen05 instructions.
        if (rank == 0) {
            constexpr uint32_t idesc = ptx::idesc_bf16_f32(BM * CG, BN);
            const uint64_t desc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem));
            const uint64_t desc_panel = ptx::smem_desc_sw128(ptx::smem_u32(panel));
            uint32_t it = 0, item = 0;
            int s = 0;
            uint32_t ph = 0;
            long long w_tmem = 0, w_full = 0;
            const long long t_begin = prof ? clock64() : 0;
            XD_ITEM_LOOP { XD_ITEM_DECODE XD_TILE_LOOP {
                const uint32_t buf = it & 1;
                wait_acc(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1, prof, w_tmem);  // epilogue drained this accumulator
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = 0; kb < nk; ++kb) {
                    if (dbg != 7 && dbg != 9) {
                        if (AS && nt == nt0) ptx::mbar_wait(&a_full_bar[kb], item & 1);
                        wait_acc(&full_bar[s], ph, prof, w_full);
                        ptx::tc_fence_after();
                    }
                    // descriptors differ between stages only in the start-address field (16-byte units)
                    const uint64_t dring = desc0 + (uint64_t)((s * C::STAGE_BYTES) >> 4);
                    const uint64_t da = AS ? desc_panel + (uint64_t)((kb * C::A_BYTES) >> 4) : dring;
                    const uint64_t db = AS ? dring : dring + (C::A_BYTES >> 4);
                    if (ptx::elect_one()) {
                        if (dbg != 2 && dbg != 4) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                // advance 32 bytes (16 bf16) inside the 128-byte swizzle row: +2 in 16-byte units
                                if constexpr (CG == 2) ptx::umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                                else ptx::umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                            }
                        }
                        if constexpr (CG == 2) {
                            if (dbg != 7) ptx::umma_commit_2sm(&empty_bar[s]);
                            if (AS && nt == nt1 - 1) ptx::umma_commit_2sm(&a_empty_bar[kb]);   // panel slot free
                            if (kb == nk - 1) ptx::umma_commit_2sm(&tmem_full_bar[buf]);
                        } else {
                            if (dbg != 7) ptx::umma_commit(&empty_bar[s]);
                            if (AS && nt == nt1 - 1) ptx::umma_commit(&a_empty_bar[kb]);
                            if (kb == nk - 1) ptx::umma_commit(&tmem_full_bar[buf]);
                        }
                    }
                    __syncwarp();
                    if (++s == C::STAGES) { s = 0; ph ^= 1; }
                }
            }}
            if (prof && lane == 0) { pslot[0] = clock64() - t_begin; pslot[1] = w_tmem; pslot[2] = w_full; pslot[10] = it; }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        const int e_warp = warp - 2;
        const int grp = e_warp >> 2;                    // which half of the tile's columns
        const int q = warp & 3;                         // TMEM lane quadrant this warp may access
        float* st = reinterpret_cast<float*>(epi + e_warp * C::EPI_WARP_BYTES);   // private 32 x 32 (+pad) transpose buffer
        const Epilogue& e = p.epi;
        const int sub_row = lane >> 3;                  // coalesced pass: 4 rows x 8 lanes x 4 columns
        const int sub_col = (lane & 7) * 4;
        const bool res_f32 = e.res_dtype == XD_F32, out_f32 = e.out_dtype == XD_F32;
        const bool direct_bf16 = C::CHUNKS_PER_GROUP >= 2 && !out_f32 && !e.gate && !e.residual && (p.N % 8 == 0) && p.direct_ok;
        const int res_ld = (int)e.res_ld, out_ld = (int)e.out_ld;     // per-tile row offsets fit 32 bits
        uint32_t it = 0, item = 0;
        long long w_acc = 0, w_res = 0, seg[4] = {0, 0, 0, 0};
        const long long t_begin = prof ? clock64() : 0;
        uint32_t cc = 0;                                 // TMA epilogue: chunks processed by this warp (box / phase)
        const uint32_t wbuf_a = ptx::smem_u32(epi + e_warp * C::EPI_WARP_BYTES);
        const uint32_t bias_a = ptx::smem_u32(bias_sm + e_warp * 512);
        uint64_t* rbar = res_bar + 2 * e_warp;
        XD_ITEM_LOOP { XD_ITEM_DECODE
            if constexpr (LNA) {
                // ---- fused LayerNorm + modulate: build this CTA's 128 x K panel of A in shared memory.
                // MEASURED NOT PROFITABLE (profiles/README.md): the 8 epilogue warps need ~10 us for the 128 rows
                // (~3600 instructions per warp at ~5 clk each, two warps per scheduler) while the stand-alone
                // LayerNorm kernel takes 7.6 us on the whole machine; kept as an opt-in (XDB200_LN_FUSED=1) and tested.
                // Each epilogue warp owns 16 rows; one row per warp at a time, lane <-> columns (i * 32 + lane) * 4
                // (the decomposition and arithmetic of ln_modulate_kernel, norm.cu: bit-identical operand).
                // The panel is free once the MMAs of the previous item have completed (a_empty, multicast commit).
                const long long t_ln0 = prof ? clock64() : 0;
                for (int kb = 0; kb < nk; ++kb) ptx::mbar_wait(&a_empty_bar[kb], (item & 1) ^ 1);
                const int nv = nk >> 1;                                  // 128-column groups
                const float Df = (float)(nk * BK);
                const uint32_t panel_a = ptx::smem_u32(panel);
                const int row0 = e_warp * 16;
                // Eight rows per pass: all 24 row loads of the pass are issued before the first reduction, the eight
                // reductions run as independent shuffle chains (one row at a time is latency-bound: 2500 clk per row).
                // shift / scale are reloaded only when the modulation row changes (DiT: once per warp).
                // All index arithmetic is 32-bit and hoisted per pass: the eight epilogue warps are latency-bound
                // (two warps per scheduler), so the instruction count of this block is its cost.
                float4 sc[3], sh[3];
                int cur_mod = -1;
                const int ldx = (int)p.ln_ld;
                const unsigned rpm = (unsigned)p.ln_rows_per_mod;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int rbase = row0 + pass * 8;
                    const int mb = mt * (BM * CG) + rank * BM + rbase;            // first global row of the pass
                    const int valid = p.M - mb;                                   // rows rr < valid exist
                    const float* xb = p.ln_x + (long long)mb * p.ln_ld + lane * 4;
                    float4 v[8][3];
                    float red[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            v[rr][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (i < nv && rr < valid) v[rr][i] = *reinterpret_cast<const float4*>(xb + rr * ldx + i * 128);
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 3; ++i) acc += v[rr][i].x + v[rr][i].y + v[rr][i].z + v[rr][i].w;
                        red[rr] = acc;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) red[rr] += __shfl_xor_sync(0xffffffffu, red[rr], o);
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const float mean = red[rr] / Df;
                        float qacc = 0.f;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            if (i < nv) {
                                v[rr][i].x -= mean; v[rr][i].y -= mean; v[rr][i].z -= mean; v[rr][i].w -= mean;
                                qacc += v[rr][i].x * v[rr][i].x + v[rr][i].y * v[rr][i].y + v[rr][i].z * v[rr][i].z +
                                        v[rr][i].w * v[rr][i].w;
                            }
                        }
                        red[rr] = qacc;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) red[rr] += __shfl_xor_sync(0xffffffffu, red[rr], o);
                    }
                    const int m_last = p.M - 1;
                    const int mod_first = (int)((unsigned)min(mb, m_last) / rpm);
                    const bool mod_uniform = mod_first == (int)((unsigned)min(mb + 7, m_last) / rpm);
                    // this lane's byte offset inside a 128-byte swizzled row, without the row XOR
                    const uint32_t a_lane = panel_a + (lane >> 4) * C::A_BYTES + ((lane & 1) << 3);
                    const uint32_t chunk = (lane & 15) >> 1;
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int r = rbase + rr;                        // row inside the CTA's 128
                        const float rstd = rsqrtf(red[rr] / Df + p.ln_eps);
                        const int mod_row = mod_uniform ? mod_first : (int)((unsigned)min(mb + rr, m_last) / rpm);
                        if (mod_row != cur_mod) {                        // warp-uniform; DiT: once per warp and item
                            cur_mod = mod_row;
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                sc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                                sh[i] = sc[i];
                                if (i < nv) {
                                    const long long off = (long long)mod_row * p.ln_mod_ld + (i * 32 + lane) * 4;
                                    if (p.ln_scale) sc[i] = __ldg(reinterpret_cast<const float4*>(p.ln_scale + off));
                                    if (p.ln_shift) sh[i] = __ldg(reinterpret_cast<const float4*>(p.ln_shift + off));
                                }
                            }
                        }
                        const bool live = rr < valid;
                        const uint32_t a_row = a_lane + r * 128 + ((chunk ^ (r & 7)) << 4);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            if (i < nv) {
                                const float y0 = fmaf(v[rr][i].x * rstd, 1.0f + sc[i].x, sh[i].x), y1 = fmaf(v[rr][i].y * rstd, 1.0f + sc[i].y, sh[i].y);
                                const float y2 = fmaf(v[rr][i].z * rstd, 1.0f + sc[i].z, sh[i].z), y3 = fmaf(v[rr][i].w * rstd, 1.0f + sc[i].w, sh[i].w);
                                // k-block 2 i + (lane >> 4)
                                ptx::sts64(a_row + 2 * i * C::A_BYTES, live ? f2_to_bf2(y0, y1) : 0u, live ? f2_to_bf2(y2, y3) : 0u);
                            }
                        }
                    }
                }
                ptx::fence_proxy_async();                                // generic-proxy writes -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) {
                    for (int kb = 0; kb < nk; ++kb) {
                        if constexpr (CG == 2) ptx::mbar_arrive_leader(&a_full_bar[kb]);
                        else ptx::mbar_arrive(&a_full_bar[kb]);
                    }
                }
                if (prof) w_res += clock64() - t_ln0;                    // reported in the "epi wait residual" slot
            }
            XD_TILE_LOOP {
            const int m0 = mt * (BM * CG) + rank * BM + q * 32;   // first row of this warp's 32-row band
            const int n0 = nt * BN;
            const uint32_t buf = it & 1;
            if constexpr (EPI != EPI_LEGACY) {
                constexpr bool out_f32 = EPI == EPI_TMA_F32;
                // ---- TMA epilogue.  Each warp owns 32 rows x CPG 32-column chunks and works in the TMEM-native
                // layout (lane = row): the residual chunk arrives by a bulk tensor load into the warp's private
                // swizzled 32 x 32 box (prefetched one chunk ahead), is combined in place with the accumulator
                // (bias from the warp's smem slice, gate through L1) and leaves by a bulk tensor store.  No
                // transpose, no per-element predicates (TMA clips / zero-fills at the M and N edges).
                constexpr int CPG = C::CHUNKS_PER_GROUP > 0 ? C::CHUNKS_PER_GROUP : 1;
                const int ncol0 = n0 + grp * CPG * 32;
                const bool has_res = e.residual != nullptr && !(kInst && dbg >= 3 && dbg != 6);
                const uint32_t box_bytes = out_f32 ? 4096u : 2048u;
#pragma unroll
                for (int k = 0; k < CPG; ++k) {
                    const int col = ncol0 + k * 32 + lane;
                    ptx::sts32(bias_a + (k * 32 + lane) * 4, (e.bias && col < p.N) ? __float_as_uint(__ldg(e.bias + col)) : 0u);
                }
                if (has_res && lane == 0) {
                    ptx::bulk_wait_read<1>();                    // the store that last read this buffer (chunk cc - 2)
                    const uint32_t b = cc & 1;
                    ptx::mbar_arrive_expect_tx(&rbar[b], box_bytes);
                    ptx::tma_load_2d_u32(wbuf_a + b * 4096, &tmRes, ptx::smem_u32(&rbar[b]), ncol0, m0);
                }
                __syncwarp();
                const float* gp = nullptr;                       // this lane's (= row's) gate row
                if (e.gate) gp = e.gate + (long long)(min(m0 + lane, p.M - 1) / e.gate_rows) * e.gate_ld;
                wait_acc(&tmem_full_bar[buf], (it >> 1) & 1, prof, w_acc);
                ptx::tc_fence_after();
                const uint32_t t_addr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16) + grp * CPG * 32;
                if (kInst && dbg >= 3 && dbg != 6) {             // experiment: mainloop only
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) ptx::mbar_arrive_leader(&tmem_empty_bar[buf]);
                        else ptx::mbar_arrive(&tmem_empty_bar[buf]);
                    }
                    continue;
                }
                // TMEM reads are the scarce resource of the epilogue (~64 B/clk/SM, 96 KB per 128 x 192 tile): the load of
                // chunk ci + 1 is in flight while chunk ci is processed (two register buffers, loop fully unrolled).
                uint32_t rr[2][32];
                ptx::tmem_ld_32x32(t_addr, rr[0]);
#pragma unroll
                for (int ci = 0; ci < CPG; ++ci, ++cc) {
                    const int nc = ncol0 + ci * 32;
                    const uint32_t b = cc & 1;
                    const uint32_t wb = wbuf_a + b * 4096;
                    uint32_t* r = rr[ci & 1];
                    long long tA = prof ? clock64() : 0;
                    if (lane == 0) {
                        if (has_res) {
                            if (ci + 1 < CPG) {                  // prefetch the next residual chunk into the other box
                                ptx::bulk_wait_read<0>();
                                ptx::mbar_arrive_expect_tx(&rbar[b ^ 1], box_bytes);
                                ptx::tma_load_2d_u32(wbuf_a + (b ^ 1) * 4096, &tmRes, ptx::smem_u32(&rbar[b ^ 1]), nc + 32, m0);
                            }
                        } else {
                            ptx::bulk_wait_read<1>();            // box b was last read by the store of chunk cc - 2
                        }
                    }
                    if (prof) { const long long t = clock64(); seg[0] += t - tA; tA = t; }     // bulk_wait_read (+ prefetch issue)
                    ptx::tmem_ld_wait();
                    if (ci + 1 < CPG) ptx::tmem_ld_32x32(t_addr + (ci + 1) * 32, rr[(ci + 1) & 1]);
                    if (prof) { const long long t = clock64(); seg[1] += t - tA; tA = t; }     // TMEM load
                    if (ci == CPG - 1) ptx::tc_fence_before();
                    __syncwarp();
                    if (ci == CPG - 1 && lane == 0) {            // last TMEM read of this tile by this warp
                        if constexpr (CG == 2) ptx::mbar_arrive_leader(&tmem_empty_bar[buf]);
                        else ptx::mbar_arrive(&tmem_empty_bar[buf]);
                    }
                    if (has_res) wait_acc(&rbar[b], (cc >> 1) & 1, prof, w_res);
                    // The shared-memory accesses are volatile asm (program order), so loads are issued in batches
                    // ahead of the arithmetic: a load -> use -> store chain per 16 bytes costs one LDS latency each.
                    if constexpr (out_f32) {
                        const uint32_t rowa = wb + lane * 128;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {            // 16 columns per half
                            uint4 bq[4], rs[4];
                            float4 g[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = 4 * h + jj;
                                g[jj] = make_float4(1.f, 1.f, 1.f, 1.f);
                                if (gp && nc + 4 * j < p.N) g[jj] = __ldg(reinterpret_cast<const float4*>(gp + nc) + j);
                            }
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) bq[jj] = ptx::lds128(bias_a + (ci * 32 + 16 * h + 4 * jj) * 4);
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                rs[jj] = make_uint4(0u, 0u, 0u, 0u);
                                if (has_res) rs[jj] = ptx::lds128(rowa + (((4 * h + jj) ^ (lane & 7)) << 4));
                            }
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int c = 16 * h + 4 * jj;
                                rs[jj].x = __float_as_uint(fmaf(act_fast<ACT>(__uint_as_float(r[c]) + __uint_as_float(bq[jj].x)), g[jj].x, __uint_as_float(rs[jj].x)));
                                rs[jj].y = __float_as_uint(fmaf(act_fast<ACT>(__uint_as_float(r[c + 1]) + __uint_as_float(bq[jj].y)), g[jj].y, __uint_as_float(rs[jj].y)));
                                rs[jj].z = __float_as_uint(fmaf(act_fast<ACT>(__uint_as_float(r[c + 2]) + __uint_as_float(bq[jj].z)), g[jj].z, __uint_as_float(rs[jj].z)));
                                rs[jj].w = __float_as_uint(fmaf(act_fast<ACT>(__uint_as_float(r[c + 3]) + __uint_as_float(bq[jj].w)), g[jj].w, __uint_as_float(rs[jj].w)));
                            }
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) ptx::sts128(rowa + (((4 * h + jj) ^ (lane & 7)) << 4), rs[jj]);
                        }
                    } else {
                        const uint32_t rowa = wb + lane * 64;
                        uint4 bq[8], rs[4];
                        float qa[16];                            // QS: this row's 8 quad sums and 8 quad sums of squares
#pragma unroll
                        for (int j = 0; j < 8; ++j) bq[j] = ptx::lds128(bias_a + (ci * 32 + 4 * j) * 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            rs[j] = make_uint4(0u, 0u, 0u, 0u);
                            if (has_res) rs[j] = ptx::lds128(rowa + ((j ^ ((lane >> 1) & 3)) << 4));
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float v[8];
                            bias_act2<ACT>(r[8 * j], r[8 * j + 1], bq[2 * j].x, bq[2 * j].y, v[0], v[1]);
                            bias_act2<ACT>(r[8 * j + 2], r[8 * j + 3], bq[2 * j].z, bq[2 * j].w, v[2], v[3]);
                            bias_act2<ACT>(r[8 * j + 4], r[8 * j + 5], bq[2 * j + 1].x, bq[2 * j + 1].y, v[4], v[5]);
                            bias_act2<ACT>(r[8 * j + 6], r[8 * j + 7], bq[2 * j + 1].z, bq[2 * j + 1].w, v[6], v[7]);
                            if (gp) {
                                float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
                                if (nc + 8 * j < p.N) g0 = __ldg(reinterpret_cast<const float4*>(gp + nc) + 2 * j);
                                if (nc + 8 * j + 4 < p.N) g1 = __ldg(reinterpret_cast<const float4*>(gp + nc) + 2 * j + 1);
                                v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
                                v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
                            }
                            if (has_res) {
                                float2 t;
                                t = bf2_to_f2(rs[j].x); v[0] += t.x; v[1] += t.y;
                                t = bf2_to_f2(rs[j].y); v[2] += t.x; v[3] += t.y;
                                t = bf2_to_f2(rs[j].z); v[4] += t.x; v[5] += t.y;
                                t = bf2_to_f2(rs[j].w); v[6] += t.x; v[7] += t.y;
                            }
                            rs[j] = make_uint4(f2_to_bf2(v[0], v[1]), f2_to_bf2(v[2], v[3]), f2_to_bf2(v[4], v[5]), f2_to_bf2(v[6], v[7]));
                            if constexpr (QS) {
                                qa[2 * j] = (v[0] + v[1]) + (v[2] + v[3]);
                                qa[2 * j + 1] = (v[4] + v[5]) + (v[6] + v[7]);
                                qa[8 + 2 * j] = fmaf(v[0], v[0], v[1] * v[1]) + fmaf(v[2], v[2], v[3] * v[3]);
                                qa[9 + 2 * j] = fmaf(v[4], v[4], v[5] * v[5]) + fmaf(v[6], v[6], v[7] * v[7]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) ptx::sts128(rowa + ((j ^ ((lane >> 1) & 3)) << 4), rs[j]);
                        if constexpr (QS) {
                            // Sum the 16 per-row values over the warp's 32 rows with a transposing butterfly (16 + 8 + 4 + 2 + 1
                            // = 31 shuffles; fixed order, no atomics): afterwards lane L holds value (L >> 1) & 15, i.e.
                            // kind (sum / sum of squares) = L >> 4, quad = (L >> 1) & 7.
#pragma unroll
                            for (int w = 8; w >= 1; w >>= 1) {
                                const bool up = (lane & (2 * w)) != 0;
#pragma unroll
                                for (int i = 0; i < w; ++i) {
                                    const float send = up ? qa[i] : qa[i + w], keep = up ? qa[i + w] : qa[i];
                                    qa[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * w);
                                }
                            }
                            qa[0] += __shfl_xor_sync(0xffffffffu, qa[0], 1);
                            const int quad = (lane >> 1) & 7;
                            if (!(lane & 1) && m0 < p.M && nc + 4 * quad < p.N)
                                p.qstats[(long long)(m0 >> 5) * p.qstats_ld + ((nc >> 2) + quad) * 2 + (lane >> 4)] = qa[0];
                        }
                    }
                    if (prof) { const long long t = clock64(); seg[2] += t - tA; tA = t; }     // residual wait + math + STS
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (dbg != 1) ptx::tma_store_2d(&tmOut, wb, nc, m0 + sp * p.ws_rows);
                        ptx::bulk_commit();
                    }
                    if (prof) { const long long t = clock64(); seg[3] += t - tA; tA = t; }     // proxy fence + store issue
                }
            } else {
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
            if (dbg >= 3 && dbg != 6) {         // experiment: mainloop only (9 = 7 + per-stage commits)
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
                            if (rr < rows_left && lane_ok && dbg != 1)
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
            }   // EPI_LEGACY
        }}
        if (prof && warp == 2 && lane == 0) {
            pslot[5] = clock64() - t_begin; pslot[6] = w_acc; pslot[7] = w_res;
            pslot[12] = seg[0]; pslot[13] = seg[1]; pslot[14] = seg[2]; pslot[15] = seg[3];
        }
        if (EPI != EPI_LEGACY && lane == 0) ptx::bulk_wait<0>();    // all bulk stores of this warp have completed
        if (prof && warp == 2 && lane == 0) pslot[11] = clock64() - t_begin;
    }
#undef XD_ITEM_LOOP
#undef XD_ITEM_DECODE
#undef XD_TILE_LOOP
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync();             // no remote arrive / peer smem read after a CTA exits
    else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (prof && threadIdx.x == 0) pslot[9] = (long long)globaltimer_ns();
}

// Split-K second pass: out = epilogue( sum_s partial[s] ) over fp32 partial tiles, summed in split order (deterministic).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int S, long long split_stride, int M, int N, const Epilogue e) {
    pdl_prologue();
    const int n4 = N >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * n4) return;
    const int m = (int)(i / n4), c = (int)(i - (long long)m * n4) * 4;
    const float* src = ws + (long long)m * N + c;
    float4 v = *reinterpret_cast<const float4*>(src);
    for (int sidx = 1; sidx < S; ++sidx) {
        const float4 t = *reinterpret_cast<const float4*>(src + sidx * split_stride);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (e.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + c));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    v.x = apply_act(v.x, e.act); v.y = apply_act(v.y, e.act); v.z = apply_act(v.z, e.act); v.w = apply_act(v.w, e.act);
    if (e.gate) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(e.gate + (long long)(m / e.gate_rows) * e.gate_ld + c));
        v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    }
    if (e.residual) {
        if (e.res_dtype == XD_F32) {
            const float4 r = *reinterpret_cast<const float4*>((const float*)e.residual + (long long)m * e.res_ld + c);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        } else {
            const uint2 u = *reinterpret_cast<const uint2*>((const bf16*)e.residual + (long long)m * e.res_ld + c);
            const float2 lo = bf2_to_f2(u.x), hi = bf2_to_f2(u.y);
            v.x += lo.x; v.y += lo.y; v.z += hi.x; v.w += hi.y;
        }
    }
    if (e.out_dtype == XD_F32) *reinterpret_cast<float4*>((float*)e.out + (long long)m * e.out_ld + c) = v;
    else *reinterpret_cast<uint2*>((bf16*)e.out + (long long)m * e.out_ld + c) = make_uint2(f2_to_bf2(v.x, v.y), f2_to_bf2(v.z, v.w));
}

// scratch for split-K partial tiles, provided by the caller (xd_set_workspace); none -> no split-K
void* g_ws = nullptr;
size_t g_ws_bytes = 0;
int g_split_k = -1;      // -1: XDB200_SPLITK (default on), 0 / 1: set by xd_set_split_k

// ------------------------------------------------------------------ host side
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

// The kernel's epilogue uses 16-byte (fp32) / 8-byte (bf16) accesses on 4-column groups.
int epilogue_vec_ok(const Epilogue& e, int N) {
    bool ok = aligned16(e.out) && aligned16(e.bias) && aligned16(e.gate) && aligned16(e.residual);
    ok = ok && (N % 4 == 0) && (e.out_ld % 4 == 0) && (e.res_ld % 4 == 0) && (e.gate_ld % 4 == 0);
    ok = ok && e.out_ld < (1 << 22) && e.res_ld < (1 << 22);      // 32-bit byte offsets inside a 128-row tile
    return ok ? 1 : 0;
}

// The TMA epilogue needs 16-byte aligned rows (global strides are multiples of 16 bytes) and a residual of the
// output's dtype; everything else takes the legacy register / transpose epilogue.
int setup_epilogue(TcParams* p, CUtensorMap* tres, CUtensorMap* tout) {
    static const int mode = getenv("XDB200_EPI") ? atoi(getenv("XDB200_EPI")) : 1;
    const Epilogue& e = p->epi;
    const bool f32 = e.out_dtype == XD_F32;
    const int esz = f32 ? 4 : 2;
    bool ok = mode != 0 && (e.out_ld * esz) % 16 == 0;
    if (e.residual) ok = ok && e.res_dtype == e.out_dtype && (e.res_ld * esz) % 16 == 0;
    p->tma_epi = 0;
    memset(tres, 0, sizeof *tres);
    memset(tout, 0, sizeof *tout);
    if (!ok) return XD_OK;
    int rc;
    if ((rc = tmap_epi(tout, e.out, p->M, p->N, e.out_ld, f32))) return rc;
    *tres = *tout;
    if (e.residual && (rc = tmap_epi(tres, e.residual, p->M, p->N, e.res_ld, f32))) return rc;
    p->tma_epi = 1;
    return XD_OK;
}

template <int BN, int ACT, int CG, bool AS, int EPI, bool LNA = false, bool QS = false>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& tr,
           const CUtensorMap& to, const TcParams& p, cudaStream_t st) {
    static bool configured = false;
    auto kernel = gemm_tc_kernel<BN, ACT, CG, AS, EPI, LNA, QS>;
    if (!configured) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, CG, AS>::SMEM) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute(max dynamic smem) failed");
            return XD_ERR_CUDA;
        }
        configured = true;
    }
    const long long items = (long long)p.ng * ((p.M + BM * CG - 1) / (BM * CG)) * p.ksplit;
    // persistent: <= one CTA (pair) per SM (pair)
    const unsigned grid = (unsigned)std::min<long long>(items, sm_count() / CG) * CG;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg<BN, CG, AS>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = xd_pdl_enabled_gemm() ? 2 : 1;
    static const bool want_prof = kInst && getenv("XDB200_PROF") != nullptr;
    TcParams pp = p;
    static long long* prof_dev = nullptr;
    if (want_prof) {                                            // debugging aid: synchronous, never in a captured graph
        if (!prof_dev) cudaMalloc(&prof_dev, 16 * 256 * sizeof(long long));
        cudaMemsetAsync(prof_dev, 0, 16 * 256 * sizeof(long long), st);
        pp.prof = prof_dev;
    }
    if (cudaLaunchKernelEx(&cfg, kernel, a0, a1, b, tr, to, pp) != cudaSuccess) {
        xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
        return XD_ERR_CUDA;
    }
    if (want_prof) {
        static long long h[16 * 256];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, prof_dev, sizeof h, cudaMemcpyDeviceToHost);
        static const char* names[16] = {"mma total", "mma wait tmem_empty", "mma wait full", "prod total", "prod wait empty",
                                        "epi total", "epi wait tmem_full", "epi wait residual", "", "", "tiles", "epi total+drain",
                                        "epi seg wait_read", "epi seg tmem_ld", "epi seg math+sts", "epi seg fence+store"};
        long long t0 = h[8], t1 = h[9];
        for (unsigned c = 0; c < grid; ++c) { t0 = std::min(t0, h[16 * c + 8]); t1 = std::max(t1, h[16 * c + 9]); }
        fprintf(stderr, "[xdb200 prof] BN=%d CG=%d AS=%d M=%d N=%d nk=%d grid=%u span=%.2f us\n", BN, CG, (int)AS, p.M, p.N,
                p.nk0 + p.nk1, grid, (t1 - t0) * 1e-3);
        for (int k = 0; k < 16; ++k) {
            if (!names[k][0]) continue;
            double sum = 0; long long mx = 0;
            for (unsigned c = 0; c < grid; ++c) { sum += h[16 * c + k]; mx = std::max(mx, h[16 * c + k]); }
            fprintf(stderr, "[xdb200 prof]   %-22s mean %9.0f  max %9lld (cycles)\n", names[k], sum / grid, mx);
        }
        double first = 1e30, last = 0;
        for (unsigned c = 0; c < grid; ++c) { first = std::min(first, (double)(h[16 * c + 8] - t0)); last = std::max(last, (double)(h[16 * c + 8] - t0)); }
        fprintf(stderr, "[xdb200 prof]   CTA start skew %.2f us\n", last * 1e-3);
    }
    return XD_OK;
}

// Split-K second pass fused with the GroupNorm (+ scale / shift, + SiLU) that consumes the contraction's output (the first conv
// of a resblock at the low resolutions: 8 x 8 and 4 x 4 pixels per sample): one CTA per sample sums the fp32 partial tiles in
// split order, keeps its <= 64 values per thread in registers, reduces the 32 groups in a fixed order (two adjacent lanes x
// four row slots per group), normalises and stores bf16 -- the un-normalised activation never exists in memory and the
// stand-alone GroupNorm launch disappears.  C / 4 divides 256, so a thread always works on the same four channels.
constexpr int GN_MAX_V4 = 16;                            // float4 per thread: P * C <= 16384
struct GnFuse {
    int nsamples, P;
    const float *gamma, *beta, *ss;
    long long ss_ld;
    int ss_div, silu;
    float eps;
    void* out;
    long long out_ld;
    int done;                                            // set by dispatch(): the fused kernel ran
};
__global__ void __launch_bounds__(256)
splitk_reduce_gn_kernel(const float* __restrict__ ws, int S, long long split_stride, int P, int N, const float* __restrict__ bias,
                        const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ ss,
                        long long ss_ld, int ss_div, float eps, int silu, bf16* __restrict__ out, long long out_ld) {
    pdl_prologue();
    __shared__ float2 part[8][32];
    const int n4 = N >> 2, sample = blockIdx.x;
    const int col4 = threadIdx.x % n4, slot = threadIdx.x / n4, slots = 256 / n4;
    const int c = col4 * 4, cpg = N / 32;
    const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[GN_MAX_V4];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < GN_MAX_V4; ++i) {
        const int r = slot + i * slots;
        if (r < P) {
            const float* src = ws + ((long long)sample * P + r) * N + c;
            float4 t = *reinterpret_cast<const float4*>(src);
            for (int sidx = 1; sidx < S; ++sidx) {
                const float4 u = *reinterpret_cast<const float4*>(src + sidx * split_stride);
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w;
            v[i] = t;
            sum += (t.x + t.y) + (t.z + t.w);
            sq += fmaf(t.x, t.x, t.y * t.y) + fmaf(t.z, t.z, t.w * t.w);
        }
    }
    // group = cpg / 4 adjacent column quads (1, 2, 4 lanes) x `slots` row slots
    const int qpg = cpg >> 2;
    for (int o = 1; o < qpg; o <<= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    const int g = col4 / qpg;
    if (col4 % qpg == 0) part[slot][g] = make_float2(sum, sq);
    __syncthreads();
    float gs = 0.f, gq = 0.f;
    for (int k = 0; k < slots; ++k) { gs += part[k][g].x; gq += part[k][g].y; }
    const float inv_cnt = 1.0f / ((float)P * (float)cpg);
    const float mean = gs * inv_cnt;
    const float rstd = rsqrtf(fmaxf(gq * inv_cnt - mean * mean, 0.f) + eps);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
    float a[4] = {rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w};
    float bb[4] = {be.x - mean * a[0], be.y - mean * a[1], be.z - mean * a[2], be.w - mean * a[3]};
    if (ss) {
        const float* row = ss + (long long)(sample / ss_div) * ss_ld;
        const float4 sc = __ldg(reinterpret_cast<const float4*>(row + c)), sh = __ldg(reinterpret_cast<const float4*>(row + N + c));
        const float s4[4] = {1.0f + sc.x, 1.0f + sc.y, 1.0f + sc.z, 1.0f + sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] *= s4[k]; bb[k] = bb[k] * s4[k] + h4[k]; }
    }
#pragma unroll
    for (int i = 0; i < GN_MAX_V4; ++i) {
        const int r = slot + i * slots;
        if (r < P) {
            float y[4] = {fmaf(v[i].x, a[0], bb[0]), fmaf(v[i].y, a[1], bb[1]), fmaf(v[i].z, a[2], bb[2]), fmaf(v[i].w, a[3], bb[3])};
            if (silu) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float hh = 0.5f * y[k];
                    float t;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hh));
                    y[k] = fmaf(hh, t, hh);
                }
            }
            *reinterpret_cast<uint2*>(out + ((long long)sample * P + r) * out_ld + c) = make_uint2(f2_to_bf2(y[0], y[1]), f2_to_bf2(y[2], y[3]));
        }
    }
}

struct TileChoice {
    int bn, cg, as;
};

// force: 0 = auto; 64/128/192/256 = 1-CTA tiles; 1128/1256 = CTA-pair 256 x 128 / 256 x 256;
// 2192 = A-stationary 1-CTA 128 x 192; 3192 / 3256 = A-stationary CTA pair 256 x 192 / 256 x 256.
TileChoice pick_tile(int N, int M, int nk, int conv, int force) {
    if (force >= 3000) return {force - 3000, 2, 1};
    if (force >= 2000) return {force - 2000, 1, 1};
    if (force >= 1000) return {force - 1000, 2, 0};
    if (force) return {force, 1, 0};
    // tcgen05.mma 128 x N x 16 from shared memory runs at ~half rate for N = 128 (operand reads
    // saturate the shared-memory port) and at ~75% of peak for N >= 192 (measured, profiles/README.md),
    // so take the widest tile that divides N.
    static const int mode = getenv("XDB200_CG") ? atoi(getenv("XDB200_CG")) : 2;     // 1 = never pair CTAs
    static const int as_mode = getenv("XDB200_AS") ? atoi(getenv("XDB200_AS")) : 0;
    if (N <= 64) return {64, 1, 0};
    const long long m_tiles = (M + BM - 1) / BM;
    if (as_mode && !conv && nk <= Cfg<192, 1, true>::A_SLOTS && N % 192 == 0 && N > 192 && m_tiles * 4 >= 3 * sm_count()) {
        if (as_mode >= 2 && N % 256 == 0 && as_mode == 3) return {256, 2, 1};
        if (as_mode >= 2) return {192, 2, 1};
        return {192, 1, 1};
    }
    // CTA pairs (cta_group::2, 256 x BN tiles) halve the B bytes each SM stages per MMA cycle, so the same ring
    // covers ~1.4x more time: measured 15 % faster on the K = 1536 contraction, 10 % on K = 384 (profiles/README.md).
    const int cg = (mode != 1 && M >= 2 * BM) ? 2 : 1;
    const long long units = sm_count() / cg;
    const long long mt = (M + BM * cg - 1) / (BM * cg);
    // Widest tile is not always best and "fills one wave" is the wrong test (a 64 x 16 x 16 x 256 conv, M = 16384,
    // N = 256: one wave of 64 pair tiles of 256 x 256 takes 19 us, two waves of 256 x 128 tiles 31 us).  Estimated
    // cost = waves x tile width x rate penalty (128-wide MMAs run at ~2/3 of the 192/256 rate, 64-wide at ~1/2);
    // ragged N is charged through the tile count.
    int best = 128;
    long long best_cost = -1, best_pad = 0;
    const int widths[4] = {192, 256, 128, 64};
    for (int i = 0; i < 4; ++i) {
        const int w = widths[i];
        const long long tiles = mt * ((N + w - 1) / w);
        const long long waves = (tiles + units - 1) / units;
        const long long cost = waves * w * (w == 128 ? 3 : (w == 64 ? 4 : 2));       // x2 to keep integers
        const long long pad = (long long)((N + w - 1) / w) * w;                       // tie-break: fewer padded columns
        if (cg == 2 && w == 64) continue;                                             // no 64-wide pair kernel
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && pad < best_pad)) { best = w; best_cost = cost; best_pad = pad; }
    }
    return {best, cg, 0};
}

// Work items for the A-stationary kernels: groups of consecutive n-tiles, as few groups as keep >= 3/4 of the
// CTAs (pairs) busy.  Without AS every tile is its own item.
void set_items(TcParams* p, const TileChoice& t) {
    p->ksplit = 1; p->kb_per_split = p->nk0 + p->nk1; p->ws_rows = 0;
    const int n_tiles = (p->N + t.bn - 1) / t.bn;
    if (!t.as) { p->ng = n_tiles; p->tpg = 1; return; }
    const long long m_tiles = (p->M + BM * t.cg - 1) / (BM * t.cg);
    const long long units = sm_count() / t.cg;
    int ng = 1;
    while (m_tiles * ng * 4 < units * 3 && ng < n_tiles) ++ng;
    p->tpg = (n_tiles + ng - 1) / ng;
    p->ng = (n_tiles + p->tpg - 1) / p->tpg;
}

int dispatch(const TileChoice& t, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, TcParams& p,
             cudaStream_t st, int* qs_emitted = nullptr, GnFuse* gnf = nullptr) {
    set_items(&p, t);
    // Split-K: long contractions over few tiles (the 8x8 / 4x4 UNet convs: 16-32 tiles, K = 2304-9216) are bound by
    // the per-CTA TMA -> MMA round trip (~370 ns per k-block), not by the tensor pipe: S CTAs share a tile's K range,
    // write fp32 partial tiles to the workspace and a second small kernel applies the epilogue to their sum.
    static const int split_env = getenv("XDB200_SPLITK") ? atoi(getenv("XDB200_SPLITK")) : 1;
    const int split_mode = g_split_k >= 0 ? g_split_k : split_env;
    const Epilogue e_final = p.epi;
    int S = 1;
    const int nk_total = p.nk0 + p.nk1;
    const long long m_tiles_cg = (p.M + BM * t.cg - 1) / (BM * t.cg);
    if (split_mode && !t.as && g_ws) {
        const long long tiles = m_tiles_cg * ((p.N + t.bn - 1) / t.bn), units = sm_count() / t.cg;
        if (tiles * 2 <= units && nk_total >= 16) {
            long long want = std::min<long long>(std::min<long long>(units / tiles, nk_total / 8), 8);
            if (want >= 2) {
                const int per = (int)((nk_total + want - 1) / want);
                const int splits = (nk_total + per - 1) / per;
                const long long mpad = m_tiles_cg * BM * t.cg;
                if (splits >= 2 && (size_t)splits * mpad * p.N * sizeof(float) <= g_ws_bytes && mpad * splits < (1LL << 31)) {
                    S = splits;
                    p.ksplit = splits; p.kb_per_split = per; p.ws_rows = (int)mpad;
                    p.epi = Epilogue{nullptr, nullptr, nullptr, g_ws, 0, 0, p.N, XD_ACT_NONE, 1, XD_F32, XD_F32};
                }
            }
        }
    }
    CUtensorMap tr, to;
    if (int rc = setup_epilogue(&p, &tr, &to)) return rc;
    if (S > 1) {
        if (!p.tma_epi) { xd_set_error(__FILE__, __LINE__, "split-K needs the TMA epilogue"); return XD_ERR_ARG; }
        if (int rc = tmap_epi(&to, g_ws, (long long)S * p.ws_rows, p.N, p.N, true)) return rc;
        tr = to;
    }
    auto finish = [&](int rc) -> int {
        if (rc != XD_OK || S == 1) return rc;
        if (gnf && !e_final.gate && !e_final.residual && e_final.act == XD_ACT_NONE && p.N % 128 == 0 && 256 % (p.N / 4) == 0 &&
            (long long)gnf->P * p.N <= 1024LL * GN_MAX_V4 && (long long)gnf->nsamples * gnf->P == p.M && gnf->out_ld % 4 == 0) {
            if (xd_launch(splitk_reduce_gn_kernel, dim3((unsigned)gnf->nsamples), dim3(256), 0, st, (const float*)g_ws, S,
                          (long long)p.ws_rows * p.N, gnf->P, p.N, e_final.bias, gnf->gamma, gnf->beta, gnf->ss, gnf->ss_ld,
                          gnf->ss_div > 0 ? gnf->ss_div : 1, gnf->eps, gnf->silu, (bf16*)gnf->out, gnf->out_ld) != cudaSuccess) {
                xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
                return XD_ERR_CUDA;
            }
            gnf->done = 1;
            return XD_OK;
        }
        const long long n4 = (long long)p.M * (p.N / 4);
        if (xd_launch(splitk_reduce_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, (const float*)g_ws, S,
                      (long long)p.ws_rows * p.N, p.M, p.N, e_final) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
            return XD_ERR_CUDA;
        }
        return XD_OK;
    };
    if (t.as && p.nk0 + p.nk1 > Cfg<192, 1, true>::A_SLOTS) {
        xd_set_error(__FILE__, __LINE__, "A-stationary tile needs K <= 384");
        return XD_ERR_ARG;
    }
    const int epi = p.tma_epi ? (p.epi.out_dtype == XD_F32 ? EPI_TMA_F32 : EPI_TMA_BF16) : EPI_LEGACY;
    // Quad statistics for the consuming GroupNorm: only from the unsplit bf16 TMA epilogue without activation (every
    // conv3x3 / 1x1 projection that feeds a GroupNorm at the resolutions where the statistics pass matters); otherwise the
    // caller is told that nothing was written and runs the stand-alone GroupNorm.
    const bool qs = p.qstats && S == 1 && epi == EPI_TMA_BF16 && p.epi.act == XD_ACT_NONE && !t.as && t.bn >= 128 &&
                    p.M % 32 == 0 && p.N % 4 == 0;
    if (qs_emitted) *qs_emitted = qs ? 1 : 0;
    if (!qs) p.qstats = nullptr;
#define XD_TC_QS(BN_, CG_)                                                                                                 \
    if (qs && t.bn == BN_ && t.cg == CG_)                                                                                  \
        return launch<BN_, XD_ACT_NONE, CG_, false, EPI_TMA_BF16, false, true>(a0, a1, b, tr, to, p, st);
    XD_TC_QS(128, 1) XD_TC_QS(192, 1) XD_TC_QS(256, 1) XD_TC_QS(128, 2) XD_TC_QS(192, 2) XD_TC_QS(256, 2)
#undef XD_TC_QS
#define XD_TC_ACT(BN_, CG_, AS_, EPI_)                                                                     \
    switch (p.epi.act) {                                                                                   \
        case XD_ACT_NONE: return finish(launch<BN_, XD_ACT_NONE, CG_, AS_, EPI_>(a0, a1, b, tr, to, p, st));       \
        case XD_ACT_SILU: return finish(launch<BN_, XD_ACT_SILU, CG_, AS_, EPI_>(a0, a1, b, tr, to, p, st));       \
        case XD_ACT_GELU_TANH: return finish(launch<BN_, XD_ACT_GELU_TANH, CG_, AS_, EPI_>(a0, a1, b, tr, to, p, st)); \
    }
#define XD_TC_CASE(BN_, CG_, AS_)                                                  \
    if (t.bn == BN_ && t.cg == CG_ && (t.as != 0) == AS_) {                        \
        if (epi == EPI_TMA_F32) { XD_TC_ACT(BN_, CG_, AS_, EPI_TMA_F32) }          \
        else if (epi == EPI_TMA_BF16) { XD_TC_ACT(BN_, CG_, AS_, EPI_TMA_BF16) }   \
        else if (!AS_) { XD_TC_ACT(BN_, CG_, false, EPI_LEGACY) }                  \
    }
    XD_TC_CASE(64, 1, false)
    XD_TC_CASE(128, 1, false)
    XD_TC_CASE(192, 1, false)
    XD_TC_CASE(256, 1, false)
    XD_TC_CASE(128, 2, false)
    XD_TC_CASE(192, 2, false)
    XD_TC_CASE(256, 2, false)
    XD_TC_CASE(192, 1, true)
    XD_TC_CASE(192, 2, true)
    XD_TC_CASE(256, 2, true)
#undef XD_TC_ACT
#undef XD_TC_CASE
    xd_set_error(__FILE__, __LINE__, "unsupported tile shape (A-stationary tiles need the TMA epilogue)");
    return XD_ERR_ARG;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI (declared in include/xdb200.h)
// ---------------------------------------------------------------------------------------------
static int gemm_impl(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                     long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                     int gate_rows, long long gate_ld, const void* residual, int res_dtype,
                     long long res_ld, void* out, int out_dtype, long long out_ld, int force_bn,
                     float* qstats, long long qstats_ld, int* qs_emitted, void* stream) {
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
    const TileChoice t = pick_tile(N, M, p.nk0 + p.nk1, 0, force_bn);
    CUtensorMap ta0, ta1, tb;
    int rc;
    if ((rc = tmap_rows(&ta0, A, M, K, lda, BM))) return rc;
    ta1 = ta0;
    if (A2 && (rc = tmap_rows(&ta1, A2, M, K2, lda2, BM))) return rc;
    if ((rc = tmap_weights(&tb, Wt, N, K + K2, ldw, t.bn / t.cg))) return rc;
    p.qstats = qstats; p.qstats_ld = qstats_ld;
    return dispatch(t, ta0, ta1, tb, p, (cudaStream_t)stream, qs_emitted);
}

extern "C" int xd_gemm_bf16_tc(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                               long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                               int gate_rows, long long gate_ld, const void* residual, int res_dtype,
                               long long res_ld, void* out, int out_dtype, long long out_ld, int force_bn,
                               void* stream) {
    return gemm_impl(A, lda, A2, lda2, K2, Wt, ldw, M, N, K, bias, act, gate, gate_rows, gate_ld, residual, res_dtype, res_ld,
                     out, out_dtype, out_ld, force_bn, nullptr, 0, nullptr, stream);
}

extern "C" int xd_gemm_bf16_tc_qstats(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                                      long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                                      int gate_rows, long long gate_ld, const void* residual, int res_dtype,
                                      long long res_ld, void* out, int out_dtype, long long out_ld, int force_bn,
                                      float* qstats, long long qstats_ld, int* emitted, void* stream) {
    XD_CHECK_ARG(qstats && emitted && qstats_ld >= N / 4 * 2);
    return gemm_impl(A, lda, A2, lda2, K2, Wt, ldw, M, N, K, bias, act, gate, gate_rows, gate_ld, residual, res_dtype, res_ld,
                     out, out_dtype, out_ld, force_bn, qstats, qstats_ld, emitted, stream);
}

static int conv_impl(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                     long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                     const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                     long long out_ld, int force_bn, float* qstats, long long qstats_ld, int* qs_emitted, void* stream,
                     GnFuse* gnf = nullptr) {
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
    const TileChoice t = pick_tile(Cout, p.M, p.nk0 + p.nk1, 1, force_bn);
    CUtensorMap ta0, ta1, tb;
    int rc;
    if ((rc = tmap_nhwc(&ta0, X, nimg, H, W, C, ldx))) return rc;
    ta1 = ta0;
    if (Xs && (rc = tmap_nhwc(&ta1, Xs, nimg, H, W, Cs, lds))) return rc;
    const long long ktot = 9LL * C + Cs;
    if ((rc = tmap_weights(&tb, Wp, Cout, ktot, ktot, t.bn / t.cg))) return rc;
    p.qstats = qstats; p.qstats_ld = qstats_ld;
    return dispatch(t, ta0, ta1, tb, p, (cudaStream_t)stream, qs_emitted, gnf);
}

extern "C" int xd_groupnorm_apply_quads(const void* x, long long ld, int nsamples, int P, int C, int groups,
                                        const float* qstats, long long qstats_ld, const float* gamma, const float* beta,
                                        const float* scale_shift, long long ss_ld, int ss_div, float eps, int silu,
                                        void* out, long long ldo, void* stream);
extern "C" int xd_groupnorm_fused(const void* x, long long ld, int nsamples, int P, int C, int groups, const float* gamma,
                                  const float* beta, const float* scale_shift, long long ss_ld, int ss_div, float eps,
                                  int silu, void* out, long long ldo, void* stream);
extern "C" int xd_groupnorm_stats(const void* x, long long ld, int nsamples, int P, int C, int groups, int inner,
                                  float* stats, void* stream);
extern "C" int xd_groupnorm_apply(const void* x, long long ld, int nsamples, int P, int C, int groups,
                                  const float* stats, const float* gamma, const float* beta, const float* scale_shift,
                                  long long ss_ld, int ss_div, float eps, int silu, int inner, int split, void* out,
                                  long long ldo, void* stream);
extern "C" int xd_groupnorm_slabs(int nsamples, int P, int C);

// out = GroupNorm32(conv3x3(X) + bias) [* (1 + scale) + shift] [SiLU], the cheapest way the shape allows:
//   split-K contraction (few output tiles, long K)  -> the reduce pass normalises (splitk_reduce_gn_kernel), 2 launches;
//   unsplit bf16 TMA epilogue                        -> quad statistics from the epilogue + one streaming pass, 2 launches;
//   otherwise                                        -> conv, then the single-pass cluster GroupNorm (or statistics + apply).
// tmp: bf16 [nimg * H * W, Cout] scratch for the un-normalised activation (unused by the first path); scratch: fp32,
// max(M / 32 * Cout / 4 * 2, nsamples * 64 * xd_groupnorm_slabs(nsamples, P, Cout)) elements.
extern "C" int xd_conv3x3_groupnorm_bf16_tc(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Wp,
                                            int Cout, const float* bias, int nsamples, const float* gamma, const float* beta,
                                            const float* scale_shift, long long ss_ld, int ss_div, float eps, int silu,
                                            void* tmp, float* scratch, void* out, long long out_ld, void* stream) {
    XD_CHECK_ARG(tmp && scratch && out && gamma && beta && nsamples > 0 && ((long long)nimg * H * W) % nsamples == 0);
    const int M = nimg * H * W, P = M / nsamples;
    GnFuse gnf{nsamples, P, gamma, beta, scale_shift, ss_ld, ss_div, silu, eps, out, out_ld, 0};
    int emitted = 0;
    const bool want_q = M % 32 == 0 && P % 32 == 0 && Cout % 128 == 0;
    int rc = conv_impl(X, ldx, nimg, H, W, C, nullptr, 0, 0, Wp, Cout, bias, XD_ACT_NONE, nullptr, 0, 0, tmp, XD_BF16, Cout, 0,
                       want_q ? scratch : nullptr, want_q ? Cout / 4 * 2 : 0, &emitted, stream, &gnf);
    if (rc != XD_OK || gnf.done) return rc;
    if (emitted)
        return xd_groupnorm_apply_quads(tmp, Cout, nsamples, P, Cout, 32, scratch, Cout / 4 * 2, gamma, beta, scale_shift, ss_ld,
                                        ss_div, eps, silu, out, out_ld, stream);
    rc = xd_groupnorm_fused(tmp, Cout, nsamples, P, Cout, 32, gamma, beta, scale_shift, ss_ld, ss_div, eps, silu, out, out_ld, stream);
    if (rc != -1) return rc;
    if ((rc = xd_groupnorm_stats(tmp, Cout, nsamples, P, Cout, 32, 1, scratch, stream))) return rc;
    return xd_groupnorm_apply(tmp, Cout, nsamples, P, Cout, 32, scratch, gamma, beta, scale_shift, ss_ld, ss_div, eps, silu, 1, 0,
                              out, out_ld, stream);
}

extern "C" int xd_conv3x3_bf16_tc(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                                  long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                                  const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                                  long long out_ld, int force_bn, void* stream) {
    return conv_impl(X, ldx, nimg, H, W, C, Xs, lds, Cs, Wp, Cout, bias, act, residual, res_dtype, res_ld, out, out_dtype,
                     out_ld, force_bn, nullptr, 0, nullptr, stream);
}

extern "C" int xd_conv3x3_bf16_tc_qstats(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                                         long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                                         const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                                         long long out_ld, int force_bn, float* qstats, long long qstats_ld, int* emitted,
                                         void* stream) {
    XD_CHECK_ARG(qstats && emitted && qstats_ld >= Cout / 4 * 2);
    return conv_impl(X, ldx, nimg, H, W, C, Xs, lds, Cs, Wp, Cout, bias, act, residual, res_dtype, res_ld, out, out_dtype,
                     out_ld, force_bn, qstats, qstats_ld, emitted, stream);
}

// LayerNorm(no affine) + modulate fused into the A operand of the GEMM:
//   out[m, n] = act( sum_k LNmod(x)[m, k] * Wt[n, k] + bias[n] ),  LNmod(x)[m, :] = LN(x[m, :]) * (1 + scale[m / rows]) + shift[m / rows]
// x fp32 [M, K] (K = 128, 256 or 384: the whole row is one A panel), out bf16.  Replaces xd_layernorm_modulate followed by
// xd_gemm_bf16_tc (reference: `modulate(norm(x), shift, scale)` then `nn.Linear`, score_networks/dit.py:37-59).
extern "C" int xd_ln_gemm_bf16_tc(const float* X, long long ldx, const float* shift, const float* scale, long long mod_ld,
                                  int rows_per_mod, float eps, const void* Wt, long long ldw, int M, int N, int K,
                                  const float* bias, int act, void* out, long long out_ld, void* stream) {
    XD_CHECK_ARG(X && Wt && out && M > 0 && N > 0 && rows_per_mod > 0);
    constexpr int kMaxKb = Cfg<192, 2, true>::A_SLOTS;
    XD_CHECK_ARG(K % 128 == 0 && K / BK <= kMaxKb && N % 192 == 0);
    XD_CHECK_ARG(ldx % 4 == 0 && mod_ld % 4 == 0 && ldw % 8 == 0 && out_ld % 8 == 0);
    XD_CHECK_ARG(aligned16(X) && aligned16(shift) && aligned16(scale) && aligned16(Wt) && aligned16(out) && aligned16(bias));
    TcParams p{};
    p.M = M; p.N = N; p.nk0 = K / BK; p.nk1 = 0; p.conv = 0;
    p.epi = Epilogue{bias, nullptr, nullptr, out, 0, 0, out_ld, act, 1, XD_BF16, XD_BF16};
    p.vec_ok = 1; p.direct_ok = 1;
    p.ln_x = X; p.ln_ld = ldx; p.ln_shift = shift; p.ln_scale = scale; p.ln_mod_ld = mod_ld;
    p.ln_rows_per_mod = rows_per_mod; p.ln_eps = eps;
    const TileChoice t{192, 2, 1};
    set_items(&p, t);
    CUtensorMap tb, tr, to;
    int rc;
    if ((rc = tmap_weights(&tb, Wt, N, K, ldw, 96))) return rc;
    if ((rc = setup_epilogue(&p, &tr, &to))) return rc;
    XD_CHECK_ARG(p.tma_epi);
    cudaStream_t st = (cudaStream_t)stream;
    switch (act) {
        case XD_ACT_NONE: return launch<192, XD_ACT_NONE, 2, true, EPI_TMA_BF16, true>(tb, tb, tb, tr, to, p, st);
        case XD_ACT_GELU_TANH: return launch<192, XD_ACT_GELU_TANH, 2, true, EPI_TMA_BF16, true>(tb, tb, tb, tr, to, p, st);
    }
    XD_CHECK_ARG(false && "xd_ln_gemm_bf16_tc: activation must be none or gelu_tanh");
    return XD_ERR_ARG;
}

// Scratch memory for split-K partial tiles (device pointer, 16-byte aligned).  Launches that use it must be ordered on
// one stream.  Without a workspace (or with too small a one) the contraction runs unsplit.
extern "C" int xd_set_workspace(void* ptr, long long bytes) {
    XD_CHECK_ARG((ptr == nullptr) == (bytes == 0) && bytes >= 0 && aligned16(ptr));
    g_ws = ptr;
    g_ws_bytes = (size_t)bytes;
    return XD_OK;
}

// Split-K changes the summation order of a contraction with the number of output tiles, i.e. with the batch size: the
// low-order bits of a sample then depend on the batch it is computed in.  xd_set_split_k(0) restores bit-exact batch
// independence (sub-batches / ragged multi-GPU shards reproduce the same rows) at the cost of the split-K speed-up.
extern "C" int xd_set_split_k(int enabled) {
    g_split_k = enabled ? 1 : 0;
    return XD_OK;
}
// Work splits whose summation order depends on the row count (split-K here, the hidden-split clusters of dit_block.cu)
bool xd_split_enabled() {
    static const int split_env = getenv("XDB200_SPLITK") ? atoi(getenv("XDB200_SPLITK")) : 1;
    return (g_split_k >= 0 ? g_split_k : split_env) != 0;
}
