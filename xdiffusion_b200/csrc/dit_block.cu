// Fused DiT half-blocks on tcgen05 / TMEM / TMA (sm_100a), hidden size D = 384.
//
// A 128-row tile of the token stream (8 images x 16 tokens) never interacts with another tile inside a DiT block
// (LayerNorm is per token, attention per image, adaLN modulation per image), so a CTA pair can walk a 256-row tile
// through several operators without a grid-wide dependency and without the intermediates leaving the SM.
//
// dit_mlp_kernel  (xd_dit_proj_mlp_bf16_tc):   reference score_networks/dit.py:46-59, layers/mlp.py:30-45
//     h1 = h + gate1 * (O Wp^T + bp)                         attention output projection, gated residual
//     a  = bf16( LN(h1) * (1 + scale2) + shift2 )            LayerNorm (no affine) + adaLN modulate
//     u  = bf16( gelu_tanh(a W1^T + b1) )                    fc1, 128 hidden columns at a time: stays in shared memory
//     h  = h1 + gate2 * (u W2^T + b2)                        fc2 accumulates over the hidden chunks in TMEM
//     stats[m] = (mean, rstd) of the new h rows              consumed by the next block's LayerNorm (no second pass)
//   replaces 4 launches (proj GEMM, LayerNorm-modulate, fc1 GEMM, fc2 GEMM) and the 16384 x 1536 bf16 round trip.
//
// One cluster of two CTAs per 256-row tile, tcgen05 cta_group::2 (each CTA owns 128 rows of every A operand and stages
// HALF of every weight tile; the leader issues M = 256 MMAs).  TMEM (512 columns per CTA): [0, 384) = acc2 (proj result,
// then h1 parked for the LayerNorm second pass, then the fc2 accumulator), [384, 512) = acc1 (one fc1 chunk).
// Warp roles (320 threads): warp 0 TMA producer (weight tiles through a 5-slot ring, in the exact order the MMA warp
// consumes them), warp 1 MMA issuer (leader CTA), warps 2..9 epilogue (TMEM lane quadrant = warp & 3, column half =
// (warp - 2) / 4), all in the TMEM-native layout lane = row.
// Steady state of the MMA queue: fc1(c + 2), fc2(c), fc1(c + 3), fc2(c + 1), ... -- the GELU of chunk c + 1 runs while
// fc1(c + 2) and fc2(c) are on the tensor pipe, so the single acc1 buffer costs bubbles at start-up only.
#include "common.cuh"
#include "ptx.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace {

using namespace tcx;

constexpr int DM = 384;                     // model width
constexpr int KB = DM / 64;                 // k-blocks of the panel (6)
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 32 * (2 + EPI_WARPS);
constexpr int A_BYTES = 128 * 64 * 2;       // one 128-row x 64-column bf16 operand tile (128-byte swizzled rows)
constexpr int PANEL_BYTES = KB * A_BYTES;   // 96 KB: O (proj A operand), then a = LN2(h1) (fc1 A operand)
constexpr int HID_BYTES = 2 * A_BYTES;      // one fc1 chunk: 128 rows x 128 hidden columns
constexpr int SLOT_BYTES = 96 * 128;        // largest per-CTA weight tile: 96 rows x 64 columns
constexpr int RING = 5;
constexpr int MISC_BYTES = 4096;            // barriers, tmem pointer, per-row statistics exchange
constexpr int SMEM_BYTES = 1024 + PANEL_BYTES + 2 * HID_BYTES + RING * SLOT_BYTES + MISC_BYTES;
constexpr int ACC1_COL = 384;
// fc2(c) is issued FC2_LAG chunks behind fc1: with a lag of 1 the GELU of chunk c (TMEM load + bias + tanh + shared-memory
// store, ~2500 clk) sits between fc1(c + 1) (1536 clk) and fc2(c) on the tensor pipe's critical path (measured 45.5 us per
// launch); with a lag of 2 it has fc1(c + 1) + fc2(c - 1) = 3072 clk to finish and the pipe only waits at start-up.
constexpr int FC2_LAG = 2;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(RING < 2 * KB, "the producer waits for the preceding kernel after RING projection weight tiles");

struct MlpParams {
    int M, hidden, rows_per_mod;
    long long mod_ld;
    const float *bp, *b1, *b2;
    const float *gate1, *shift2, *scale2, *gate2;
    float2* stats_out;
    float eps;
    float* h_out;               // split kernels (G > 1): the final pass addresses h_out directly
    long long ldh;
    long long* prof;            // -DXDB200_INSTRUMENT + XDB200_DIT_PROF=<device pointer>: 64 clock64() stamps per CTA
    const uint8_t *wp, *w1, *w2;   // the three weight matrices (bytes), for the L2 prefetch at kernel entry
    const int* mod_rows;        // optional: modulation row of image i is mod_rows[i] (rows of a per-loop table) instead of i
};

// Every CTA pulls its share of a weight matrix into L2 (4 KB pieces, round-robin over the grid).  Issued before
// griddepcontrol.wait: weights do not depend on the preceding kernel, so the prefetch overlaps its tail.  Between two uses
// of a block's weights (one timestep, ~1 ms) some 0.5 GB of activations stream through the 126 MB L2, so the first CTA
// to touch a weight tile would otherwise pay DRAM latency on the 5-slot operand ring, and all CTAs run in lockstep.
__device__ __forceinline__ void prefetch_weights_l2(const uint8_t* w, uint32_t bytes) {
    for (uint32_t off = blockIdx.x * 4096u; off < bytes; off += gridDim.x * 4096u)
        ptx::prefetch_l2(w + off, min(4096u, bytes - off));
}

// phase stamps for tools/prof_dit_phases.py (compiled out of the product build)
#ifdef XDB200_INSTRUMENT
#define DIT_STAMP(slot) do { if (p.prof) p.prof[(long long)blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#else
#define DIT_STAMP(slot) do { } while (0)
#endif

// Warp-level pass over this warp's 32 rows x 192 columns of acc2 (six 32-column chunks):
//   v = res + gate * (acc + bias),  res = the h box (TMA load), v -> h (TMA store from the same box),
// optionally v -> TMEM (parked for a second pass), and the shifted sums  s = sum(v - c0), qq = sum((v - c0)^2).
// Every chunk has its own mbarrier and (except one) its own 4 KB box, so all residual loads are in flight at once: with
// two boxes the load of chunk i + 1 could only be issued after the store of chunk i - 1 had drained its box, a
// store-drain + load-latency chain of ~1.2 us per chunk (the pass then cost ~7 us of the kernel's 45).
//   FINAL = false (first pass): the caller issued the loads of chunks 0 and 1 before it waited for the accumulator; chunks
//           2..4 are issued here and chunk 5 re-uses the box of chunk 0 once that chunk's store has drained it.
//   FINAL = true: six distinct boxes; the caller issued chunks 0..2 (into the panel, idle after the last fc1 MMA) while the
//           last fc2 MMAs were still running, chunks 3..5 are issued here.
template <bool TO_TMEM, bool FINAL>
__device__ __forceinline__ void gated_residual_pass(const CUtensorMap* tmIn, const CUtensorMap* tmH, const uint32_t (&box)[6],
                                                    uint64_t* rbar, uint32_t parity, uint32_t t_addr, int m0w, int col_base,
                                                    const float* bias, const float* gp, int lane, float& c0, float& s,
                                                    float& qq) {
    if (lane == 0) {
#pragma unroll
        for (int ci = FINAL ? 3 : 2; ci < (FINAL ? 6 : 5); ++ci) {
            ptx::mbar_arrive_expect_tx(&rbar[ci], 4096);
            ptx::tma_load_2d_u32(box[ci], tmIn, ptx::smem_u32(&rbar[ci]), col_base + ci * 32, m0w);
        }
    }
    __syncwarp();
    uint32_t rr[2][32];
    ptx::tmem_ld_32x32(t_addr, rr[0]);
    s = 0.f; qq = 0.f; c0 = 0.f;
#pragma unroll
    for (int ci = 0; ci < 6; ++ci) {
        const int nc = col_base + ci * 32;
        const uint32_t wb = box[ci];
        uint32_t* r = rr[ci & 1];
        if (!FINAL && ci == 2 && lane == 0) {            // box[5] == box[0]: free once the store of chunk 0 has read it
            ptx::bulk_wait_read<1>();
            ptx::mbar_arrive_expect_tx(&rbar[5], 4096);
            ptx::tma_load_2d_u32(box[5], tmIn, ptx::smem_u32(&rbar[5]), col_base + 5 * 32, m0w);
        }
        ptx::tmem_ld_wait();
        if (ci + 1 < 6) ptx::tmem_ld_32x32(t_addr + (ci + 1) * 32, rr[(ci + 1) & 1]);
        ptx::mbar_wait(&rbar[ci], parity);
        const uint32_t rowa = wb + lane * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                    // 16 columns per half
            uint4 rs[4];
            float4 g[4], bq[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                g[jj] = __ldg(reinterpret_cast<const float4*>(gp + nc) + 4 * h + jj);
                bq[jj] = __ldg(reinterpret_cast<const float4*>(bias + nc) + 4 * h + jj);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) rs[jj] = ptx::lds128(rowa + (((4 * h + jj) ^ (lane & 7)) << 4));
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c = 16 * h + 4 * jj;
                float v0 = fmaf(__uint_as_float(r[c]) + bq[jj].x, g[jj].x, __uint_as_float(rs[jj].x));
                float v1 = fmaf(__uint_as_float(r[c + 1]) + bq[jj].y, g[jj].y, __uint_as_float(rs[jj].y));
                float v2 = fmaf(__uint_as_float(r[c + 2]) + bq[jj].z, g[jj].z, __uint_as_float(rs[jj].z));
                float v3 = fmaf(__uint_as_float(r[c + 3]) + bq[jj].w, g[jj].w, __uint_as_float(rs[jj].w));
                if (ci == 0 && h == 0 && jj == 0) c0 = v0;
                const float d0 = v0 - c0, d1 = v1 - c0, d2 = v2 - c0, d3 = v3 - c0;
                s += (d0 + d1) + (d2 + d3);
                qq = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, qq))));
                rs[jj] = make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
                if (TO_TMEM) { r[c] = rs[jj].x; r[c + 1] = rs[jj].y; r[c + 2] = rs[jj].z; r[c + 3] = rs[jj].w; }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) ptx::sts128(rowa + (((4 * h + jj) ^ (lane & 7)) << 4), rs[jj]);
        }
        if (TO_TMEM) ptx::tmem_st_32x32(t_addr + ci * 32, r);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            ptx::tma_store_2d(tmH, wb, nc, m0w);
            ptx::bulk_commit();
        }
    }
    if (TO_TMEM) ptx::tmem_st_wait();
}

// Combine the two column halves' shifted sums of a row (Chan's parallel formula; n = 192 each) -> mean, rstd.
__device__ __forceinline__ void combine_stats(float2* stat_sm, int grp, int row, int q, float c0, float s, float qq, float eps,
                                              float& mean, float& rstd) {
    const float n = 192.0f;
    const float mh = c0 + s / n;
    const float m2h = fmaxf(qq - s * s / n, 0.f);
    stat_sm[grp * 128 + row] = make_float2(mh, m2h);
    ptx::named_bar_sync(1 + q, 64);
    const float2 o = stat_sm[(grp ^ 1) * 128 + row];
    mean = 0.5f * (mh + o.x);
    const float dm = mh - o.x;
    const float m2 = m2h + o.y + dm * dm * (n * 0.5f);
    rstd = rsqrtf(m2 / (2.0f * n) + eps);
}

// G > 1 ("split" kernels, small M): a cluster of G CTA pairs shares one 256-row tile.  Every pair computes the projection,
// h1 and the LayerNorm panel redundantly (1/9 of the work), runs fc1 -> GELU -> fc2 over ITS 1/G of the hidden units, and the
// G partial fc2 tiles are reduced through distributed shared memory: pair g owns columns [g * 384/G, (g + 1) * 384/G) of
// the tile, every CTA pushes the matching slices of its TMEM partial into the owner's shared memory (st.shared::cluster),
// and the owner adds them in pair order (deterministic), applies bias / gate / residual and writes h.  h_in and h_out are
// distinct buffers here (another pair may still be reading h while this one writes h1).
template <int G>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dit_mlp_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWp,
               const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
               const __grid_constant__ CUtensorMap tmHin, const __grid_constant__ CUtensorMap tmH, const MlpParams p) {
    pdl_launch_dependents();
    if (threadIdx.x == 64) DIT_STAMP(63);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* panel = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* hid = panel + PANEL_BYTES;                 // [2][HID_BYTES]; aliased by the epilogue staging boxes
    uint8_t* ring = hid + 2 * HID_BYTES;
    uint8_t* misc = ring + RING * SLOT_BYTES;
    uint64_t* ring_full = reinterpret_cast<uint64_t*>(misc);     // [RING]
    uint64_t* ring_empty = ring_full + RING;                     // [RING]
    uint64_t* panel_full = ring_empty + RING;                    // [KB]
    uint64_t* acc2_full = panel_full + KB;                       // completes twice: proj done, fc2 done
    uint64_t* a_full = acc2_full + 1;                            // LN2 panel written (16 warp arrivals at the leader)
    uint64_t* acc1_full = a_full + 1;
    uint64_t* acc1_empty = acc1_full + 1;                        // 16 warp arrivals at the leader
    uint64_t* hid_full = acc1_empty + 1;                         // [2], 16 warp arrivals at the leader
    uint64_t* hid_empty = hid_full + 2;                          // [2], tcgen05.commit multicast
    uint64_t* res_bar = hid_empty + 2;                           // [EPI_WARPS][6]: one per residual box of a pass
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 6 * EPI_WARPS);
    float2* stat_sm = reinterpret_cast<float2*>(misc + 1024);    // [2][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int crank = (int)ptx::cluster_ctarank();               // pair g = crank / 2, CTA `rank` of the pair
    const int rank = crank & 1;
    const int grp_id = crank >> 1;
    const int tile = blockIdx.x / (2 * G);
    const int m0 = tile * 256 + rank * 128;                      // this CTA's first row
    const int nch = p.hidden / (128 * G);                        // hidden chunks of this pair
    const int ch0 = grp_id * nch;                                // ... starting at this chunk of the full hidden dimension
    const uint16_t pair_mask = (uint16_t)(3u << (crank & ~1));

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmO);
        ptx::prefetch_tmap(&tmWp);
        ptx::prefetch_tmap(&tmW1);
        ptx::prefetch_tmap(&tmW2);
        ptx::prefetch_tmap(&tmHin);
        ptx::prefetch_tmap(&tmH);
        for (int s = 0; s < RING; ++s) {
            ptx::mbar_init(&ring_full[s], 2);
            ptx::mbar_init(&ring_empty[s], 1);
        }
        for (int k = 0; k < KB; ++k) ptx::mbar_init(&panel_full[k], 2);
        ptx::mbar_init(acc2_full, 1);
        ptx::mbar_init(a_full, 2 * EPI_WARPS);
        ptx::mbar_init(acc1_full, 1);
        ptx::mbar_init(acc1_empty, 2 * EPI_WARPS);
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&hid_full[b], 2 * EPI_WARPS);
            ptx::mbar_init(&hid_empty[b], 1);
        }
        for (int k = 0; k < 6 * EPI_WARPS; ++k) ptx::mbar_init(&res_bar[k], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_2sm(tmem_ptr, 512);
        ptx::tmem_relinquish_2sm();
    }
    if (warp == 2 && lane == 0) {
        prefetch_weights_l2(p.wp, DM * DM * 2);
        prefetch_weights_l2(p.w1, (uint32_t)p.hidden * DM * 2);
        prefetch_weights_l2(p.w2, (uint32_t)p.hidden * DM * 2);
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (threadIdx.x == 64) DIT_STAMP(61);
    // griddepcontrol.wait is executed per role, by every thread that touches activations: the producer first fills the
    // operand ring with projection weights (constant data), the MMA warp only ever reads shared memory / TMEM.

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs: own rows / own tile halves)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            auto load_b = [&](const CUtensorMap* tm, uint32_t bytes, int ck, int cn) {
                ptx::mbar_wait(&ring_empty[s], ph ^ 1);
                ptx::mbar_arrive_expect_tx_leader(&ring_full[s], bytes);
                ptx::tma_load_2d_2sm(ring + s * SLOT_BYTES, tm, &ring_full[s], ck, cn);
                if (++s == RING) { s = 0; ph ^= 1; }
            };
            for (int i = 0; i < 2 * KB; ++i) {
                if (i == RING) {                         // the ring is full of weights: now wait for the producer of O
                    pdl_wait();
                    for (int kb = 0; kb < KB; ++kb) {
                        ptx::mbar_arrive_expect_tx_leader(&panel_full[kb], A_BYTES);
                        ptx::tma_load_2d_2sm(panel + kb * A_BYTES, &tmO, &panel_full[kb], kb * 64, m0);
                    }
                }
                load_b(&tmWp, 96 * 128, (i >> 1) * 64, (i & 1) * 192 + rank * 96);
            }
            for (int it = 0; it < nch + FC2_LAG; ++it) {
                if (it < nch)
                    for (int kb = 0; kb < KB; ++kb) load_b(&tmW1, 64 * 128, kb * 64, (ch0 + it) * 128 + rank * 64);
                if (it >= FC2_LAG)
                    for (int kb2 = 0; kb2 < 2; ++kb2)
                        for (int nh = 0; nh < 2; ++nh)
                            load_b(&tmW2, 96 * 128, (ch0 + it - FC2_LAG) * 128 + kb2 * 64, nh * 192 + rank * 96);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform loop)
        if (rank == 0) {
            constexpr uint32_t idesc192 = ptx::idesc_bf16_f32(256, 192);
            constexpr uint32_t idesc128 = ptx::idesc_bf16_f32(256, 128);
            const uint64_t d_panel = ptx::smem_desc_sw128(ptx::smem_u32(panel));
            const uint64_t d_hid = ptx::smem_desc_sw128(ptx::smem_u32(hid));
            const uint64_t d_ring = ptx::smem_desc_sw128(ptx::smem_u32(ring));
            int s = 0;
            uint32_t ph = 0;
            // four K = 16 steps on one (A tile, B slot); frees the slot when they have completed
            auto mma_block = [&](uint32_t tmem_d, uint64_t da, uint32_t idesc, bool fresh) {
                ptx::mbar_wait(&ring_full[s], ph);
                ptx::tc_fence_after();
                const uint64_t db = d_ring + (uint64_t)((s * SLOT_BYTES) >> 4);
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (fresh && k == 0) ? 0u : 1u);
                    ptx::umma_commit_2sm(&ring_empty[s], pair_mask);
                }
                __syncwarp();
                if (++s == RING) { s = 0; ph ^= 1; }
            };
            // ---- proj: acc2 = O Wp^T
            if (lane == 0) DIT_STAMP(48);
            for (int kb = 0; kb < KB; ++kb) {
                ptx::mbar_wait(&panel_full[kb], 0);
                if (kb == 0 && lane == 0) DIT_STAMP(49);
                for (int nh = 0; nh < 2; ++nh)
                    mma_block(tmem_base + nh * 192, d_panel + (uint64_t)((kb * A_BYTES) >> 4), idesc192, kb == 0);
            }
            if (ptx::elect_one()) ptx::umma_commit_2sm(acc2_full, pair_mask);
            __syncwarp();
            // ---- MLP: fc1(it), fc2(it - FC2_LAG)
            if (lane == 0) DIT_STAMP(50);
            ptx::mbar_wait(a_full, 0);
            ptx::tc_fence_after();
            if (lane == 0) DIT_STAMP(51);
            for (int it = 0; it < nch + FC2_LAG; ++it) {
                if (it < nch) {
                    if (it >= 1) {
                        ptx::mbar_wait(acc1_empty, (it - 1) & 1);
                        ptx::tc_fence_after();
                    }
                    for (int kb = 0; kb < KB; ++kb)
                        mma_block(tmem_base + ACC1_COL, d_panel + (uint64_t)((kb * A_BYTES) >> 4), idesc128, kb == 0);
                    if (ptx::elect_one()) ptx::umma_commit_2sm(acc1_full, pair_mask);
                    __syncwarp();
                }
                if (it >= FC2_LAG) {
                    const int c = it - FC2_LAG, b = c & 1;
                    ptx::mbar_wait(&hid_full[b], (c >> 1) & 1);
                    ptx::tc_fence_after();
                    for (int kb2 = 0; kb2 < 2; ++kb2)
                        for (int nh = 0; nh < 2; ++nh)
                            mma_block(tmem_base + nh * 192, d_hid + (uint64_t)((b * HID_BYTES + kb2 * A_BYTES) >> 4), idesc192,
                                      c == 0 && kb2 == 0);
                    if (ptx::elect_one()) ptx::umma_commit_2sm(&hid_empty[b], pair_mask);
                    __syncwarp();
                }
            }
            if (ptx::elect_one()) ptx::umma_commit_2sm(acc2_full, pair_mask);
            __syncwarp();
            if (lane == 0) DIT_STAMP(52);
        }
    } else {
        // ------------------------------------------------------------ epilogue warps
        const int e_warp = warp - 2;
        const int grp = e_warp >> 2;                    // column half
        const int q = warp & 3;                         // TMEM lane quadrant
        const int row = q * 32 + lane;                  // row inside this CTA's 128
        const int m0w = m0 + q * 32;
        const int gm = min(m0 + row, p.M - 1);
        const int img = gm / p.rows_per_mod;
        const long long mod_off = (long long)(p.mod_rows ? __ldg(p.mod_rows + img) : img) * p.mod_ld;
        // residual boxes of this warp (4 KB each, 1 KB aligned): two in the hidden buffers (idle until the first GELU), three
        // in the panel -- the very area this warp's LayerNorm pass fills afterwards (rows 32q.., k-blocks 3 grp..), idle between
        // the last proj MMA and that pass -- and, for the final pass (nothing else is live then), one in the weight ring
        const uint32_t hbox = ptx::smem_u32(hid + e_warp * 8192);
        const uint32_t pbox = ptx::smem_u32(panel) + 3 * grp * A_BYTES + q * 4096;
        const uint32_t box1[6] = {hbox, hbox + 4096, pbox, pbox + A_BYTES, pbox + 2 * A_BYTES, hbox};
        uint64_t* rbar = res_bar + 6 * e_warp;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const int col_base = grp * 192;
        float c0, s, qq, mean, rstd;

        // ---- h1 = h + gate1 * (acc2 + bp) -> h (global) and TMEM; LayerNorm statistics
        const bool st = e_warp == 0 && lane == 0;        // the stamping thread (instrumented builds)
        pdl_wait();
        if (st) DIT_STAMP(0);
        if (lane == 0) {                                 // the first two residual boxes travel while the projection runs
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                ptx::mbar_arrive_expect_tx(&rbar[ci], 4096);
                ptx::tma_load_2d_u32(box1[ci], &tmHin, ptx::smem_u32(&rbar[ci]), col_base + ci * 32, m0w);
            }
        }
        ptx::mbar_wait(acc2_full, 0);
        ptx::tc_fence_after();
        if (st) DIT_STAMP(1);
        gated_residual_pass<true, false>(&tmHin, &tmH, box1, rbar, 0, t_lane + col_base, m0w, col_base, p.bp, p.gate1 + mod_off, lane, c0, s, qq);
        if (st) DIT_STAMP(2);
        combine_stats(stat_sm, grp, row, q, c0, s, qq, p.eps, mean, rstd);
        if (lane == 0) ptx::bulk_wait_read<1>();         // the stores of chunks 2..4 have drained the panel boxes
        __syncwarp();
        // ---- second pass over the parked h1: a = LN(h1) * (1 + scale2) + shift2 -> bf16 panel (fc1's A operand)
        {
            const float* scp = p.scale2 + mod_off;
            const float* shp = p.shift2 + mod_off;
            const uint32_t panel_a = ptx::smem_u32(panel);
            uint32_t rr[2][32];
            ptx::tmem_ld_32x32(t_lane + col_base, rr[0]);
#pragma unroll
            for (int ci = 0; ci < 6; ++ci) {
                const int col = col_base + ci * 32;
                uint32_t* r = rr[ci & 1];
                ptx::tmem_ld_wait();
                if (ci + 1 < 6) ptx::tmem_ld_32x32(t_lane + col + 32, rr[(ci + 1) & 1]);
                const uint32_t rowa = panel_a + (col >> 6) * A_BYTES + row * 128;
                const int j0 = (col & 63) >> 3;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scp + col) + 2 * jj);
                    const float4 sc1 = __ldg(reinterpret_cast<const float4*>(scp + col) + 2 * jj + 1);
                    const float4 sh0 = __ldg(reinterpret_cast<const float4*>(shp + col) + 2 * jj);
                    const float4 sh1 = __ldg(reinterpret_cast<const float4*>(shp + col) + 2 * jj + 1);
                    const int c = 8 * jj;
                    const float y0 = fmaf((__uint_as_float(r[c]) - mean) * rstd, 1.0f + sc0.x, sh0.x);
                    const float y1 = fmaf((__uint_as_float(r[c + 1]) - mean) * rstd, 1.0f + sc0.y, sh0.y);
                    const float y2 = fmaf((__uint_as_float(r[c + 2]) - mean) * rstd, 1.0f + sc0.z, sh0.z);
                    const float y3 = fmaf((__uint_as_float(r[c + 3]) - mean) * rstd, 1.0f + sc0.w, sh0.w);
                    const float y4 = fmaf((__uint_as_float(r[c + 4]) - mean) * rstd, 1.0f + sc1.x, sh1.x);
                    const float y5 = fmaf((__uint_as_float(r[c + 5]) - mean) * rstd, 1.0f + sc1.y, sh1.y);
                    const float y6 = fmaf((__uint_as_float(r[c + 6]) - mean) * rstd, 1.0f + sc1.z, sh1.z);
                    const float y7 = fmaf((__uint_as_float(r[c + 7]) - mean) * rstd, 1.0f + sc1.w, sh1.w);
                    ptx::sts128(rowa + (((j0 + jj) ^ (row & 7)) << 4),
                                make_uint4(f2_to_bf2(y0, y1), f2_to_bf2(y2, y3), f2_to_bf2(y4, y5), f2_to_bf2(y6, y7)));
                }
            }
        }
        ptx::fence_proxy_async();                        // generic-proxy panel writes -> visible to tcgen05.mma
        ptx::tc_fence_before();
        if (lane == 0) ptx::bulk_wait_read<0>();         // staging boxes alias the hidden buffers: stores have read them
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_leader(a_full);
        if (st) DIT_STAMP(3);

        // ---- fc1 chunks: u = gelu(acc1 + b1) -> bf16 hidden buffer (fc2's A operand)
        const uint32_t hid_a = ptx::smem_u32(hid);
        // boxes of the final pass: the panel first (free as soon as the last fc1 chunk has been accumulated)
        const uint32_t box2[6] = {pbox, pbox + A_BYTES, pbox + 2 * A_BYTES, hbox, hbox + 4096, ptx::smem_u32(ring) + e_warp * 4096};
        for (int c = 0; c < nch; ++c) {
            const int b = c & 1;
            if (lane < 2)                                // this chunk's 64 bias values (2 lines) into L1 before they are needed
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p.b1 + (ch0 + c) * 128 + grp * 64 + lane * 32));
            ptx::mbar_wait(acc1_full, c & 1);
            ptx::tc_fence_after();
            if (st) DIT_STAMP(8 + 2 * c);
            if (G == 1 && c == nch - 1 && lane == 0) {   // no MMA reads the panel any more: h1 boxes 0..2 of the final pass
                ptx::bulk_wait<0>();                     // (this warp's h1 stores have landed before they are re-read)
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    ptx::mbar_arrive_expect_tx(&rbar[ci], 4096);
                    ptx::tma_load_2d_u32(box2[ci], &tmH, ptx::smem_u32(&rbar[ci]), col_base + ci * 32, m0w);
                }
            }
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32(t_lane + ACC1_COL + grp * 64, r0);
            ptx::tmem_ld_32x32(t_lane + ACC1_COL + grp * 64 + 32, r1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(acc1_empty);
            ptx::mbar_wait(&hid_empty[b], ((c >> 1) & 1) ^ 1);
            const uint32_t hb = hid_a + b * HID_BYTES + grp * A_BYTES + row * 128;
            const float* bptr = p.b1 + (ch0 + c) * 128 + grp * 64;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t* r = j < 4 ? r0 + 8 * j : r1 + 8 * (j - 4);
                const float4 ba = __ldg(reinterpret_cast<const float4*>(bptr) + 2 * j);
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bptr) + 2 * j + 1);
                float v[8];
                bias_act2<XD_ACT_GELU_TANH>(r[0], r[1], __float_as_uint(ba.x), __float_as_uint(ba.y), v[0], v[1]);
                bias_act2<XD_ACT_GELU_TANH>(r[2], r[3], __float_as_uint(ba.z), __float_as_uint(ba.w), v[2], v[3]);
                bias_act2<XD_ACT_GELU_TANH>(r[4], r[5], __float_as_uint(bb.x), __float_as_uint(bb.y), v[4], v[5]);
                bias_act2<XD_ACT_GELU_TANH>(r[6], r[7], __float_as_uint(bb.z), __float_as_uint(bb.w), v[6], v[7]);
                ptx::sts128(hb + ((j ^ (row & 7)) << 4),
                            make_uint4(f2_to_bf2(v[0], v[1]), f2_to_bf2(v[2], v[3]), f2_to_bf2(v[4], v[5]), f2_to_bf2(v[6], v[7])));
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(&hid_full[b]);
            if (st) DIT_STAMP(9 + 2 * c);
        }

        // ---- h = h1 + gate2 * (acc2 + b2); statistics of the new rows for the next block's LayerNorm
        if (lane == 0) ptx::bulk_wait<0>();              // this warp's h1 stores have landed before they are re-read
        __syncwarp();
        ptx::mbar_wait(acc2_full, 1);
        ptx::tc_fence_after();
        if (st) DIT_STAMP(40);
        if constexpr (G == 1) {
            gated_residual_pass<false, true>(&tmH, &tmH, box2, rbar, 1, t_lane + col_base, m0w, col_base, p.b2, p.gate2 + mod_off, lane, c0, s, qq);
            combine_stats(stat_sm, grp, row, q, c0, s, qq, p.eps, mean, rstd);
            if (p.stats_out && grp == 0 && m0 + row < p.M) p.stats_out[m0 + row] = make_float2(mean, rstd);
            if (lane == 0) ptx::bulk_wait_read<0>();     // the stores have read their boxes (grid completion publishes them)
            if (st) DIT_STAMP(41);
        }
    }
    if constexpr (G > 1) {
        // ------------------------------------------------------------ cross-pair reduction of the fc2 partial tiles (DSMEM)
        // Receive buffer (aliases panel / hidden buffers / weight ring, all idle once every MMA of the cluster has completed):
        // [G source pairs][128 rows][W columns] fp32, rows padded by 16 bytes (conflict-free 16-byte accesses, lane = row);
        // behind it [G][128] float2 (mean, M2) of the column slices, used by pair 0 only.
        constexpr int W = DM / G;
        constexpr int RS = W * 4 + 16;
        constexpr int STATX_OFF = G * 128 * RS;
        static_assert(W % 32 == 0 && STATX_OFF + G * 128 * 8 <= PANEL_BYTES + 2 * HID_BYTES + RING * SLOT_BYTES, "receive buffer");
        const uint32_t recv = ptx::smem_u32(panel);
        const int e_warp = warp - 2, grp = e_warp >> 2, q = warp & 3, row = q * 32 + lane;
        __syncwarp();
        ptx::tc_fence_before();
        ptx::cluster_sync();                             // #1: nobody in the cluster reads its operand buffers any more
        if (threadIdx.x == 64) DIT_STAMP(42);
        if (warp >= 2) {
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 192;
            uint32_t rr[2][32];
            ptx::tmem_ld_32x32(t_addr, rr[0]);
#pragma unroll
            for (int ci = 0; ci < 6; ++ci) {
                const int col = grp * 192 + ci * 32;
                const int owner = col / W, cin = col - owner * W;
                const uint32_t* r = rr[ci & 1];
                ptx::tmem_ld_wait();
                if (ci + 1 < 6) ptx::tmem_ld_32x32(t_addr + (ci + 1) * 32, rr[(ci + 1) & 1]);
                const uint32_t dst = ptx::mapa(recv + (grp_id * 128 + row) * RS + cin * 4, (uint32_t)(owner * 2 + rank));
#pragma unroll
                for (int j = 0; j < 8; ++j) ptx::sts128_cluster(dst + 16 * j, make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
            }
        }
        __syncwarp();
        if (threadIdx.x == 64) DIT_STAMP(43);
        ptx::cluster_sync();                             // #2: every partial slice has landed in its owner
        if (threadIdx.x == 64) DIT_STAMP(44);
        if (warp >= 2) {
            const int gm = m0 + row;
            const bool live = gm < p.M;
            const int img = min(gm, p.M - 1) / p.rows_per_mod;
            const long long mod_off = (long long)(p.mod_rows ? __ldg(p.mod_rows + img) : img) * p.mod_ld;
            float* hrow = p.h_out + (long long)min(gm, p.M - 1) * p.ldh + grp_id * W;
            const float* gp = p.gate2 + mod_off + grp_id * W;
            const float* bp2 = p.b2 + grp_id * W;
            float c0 = 0.f, s = 0.f, qq = 0.f;
            int ncols = 0;
#pragma unroll 1
            for (int ci = grp; ci < W / 32; ci += 2) {  // the two warps of a row quadrant alternate over the 32-column chunks
                float4 acc[8], h1[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) h1[j] = __ldcg(reinterpret_cast<const float4*>(hrow + ci * 32) + j);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = ptx::lds128(recv + row * RS + ci * 128 + 16 * j);
                    acc[j] = make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
                }
#pragma unroll
                for (int gg = 1; gg < G; ++gg) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 v = ptx::lds128(recv + (gg * 128 + row) * RS + ci * 128 + 16 * j);
                        acc[j].x += __uint_as_float(v.x); acc[j].y += __uint_as_float(v.y);
                        acc[j].z += __uint_as_float(v.z); acc[j].w += __uint_as_float(v.w);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 gt = __ldg(reinterpret_cast<const float4*>(gp + ci * 32) + j);
                    const float4 bq = __ldg(reinterpret_cast<const float4*>(bp2 + ci * 32) + j);
                    float4 v;
                    v.x = fmaf(acc[j].x + bq.x, gt.x, h1[j].x); v.y = fmaf(acc[j].y + bq.y, gt.y, h1[j].y);
                    v.z = fmaf(acc[j].z + bq.z, gt.z, h1[j].z); v.w = fmaf(acc[j].w + bq.w, gt.w, h1[j].w);
                    if (ncols == 0 && j == 0) c0 = v.x;
                    const float d0 = v.x - c0, d1 = v.y - c0, d2 = v.z - c0, d3 = v.w - c0;
                    s += (d0 + d1) + (d2 + d3);
                    qq = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, qq))));
                    if (live) *(reinterpret_cast<float4*>(hrow + ci * 32) + j) = v;
                }
                ncols += 32;
            }
            // (mean, M2) of this thread's columns -> of the pair's column slice -> pushed to pair 0 of this tile half
            const float n_a = (float)ncols;
            const float mean_a = c0 + s / n_a, m2_a = fmaxf(qq - s * s / n_a, 0.f);
            stat_sm[grp * 128 + row] = make_float2(mean_a, m2_a);
            ptx::named_bar_sync(1 + q, 64);
            if (grp == 0) {
                const float2 o = stat_sm[128 + row];
                const float n_b = (float)(W - ncols);     // columns taken by the partner warp (grp 1)
                const float dm = o.x - mean_a;
                const float mean_s = mean_a + dm * (n_b / (float)W);
                const float m2_s = m2_a + o.y + dm * dm * (n_a * n_b / (float)W);
                const uint32_t dst = ptx::mapa(recv + STATX_OFF + (grp_id * 128 + row) * 8, (uint32_t)rank);
                ptx::sts64_cluster(dst, __float_as_uint(mean_s), __float_as_uint(m2_s));
            }
        }
        __syncwarp();
        if (threadIdx.x == 64) DIT_STAMP(45);
        ptx::cluster_sync();                             // #3: slice statistics have landed in pair 0
        if (threadIdx.x == 64) DIT_STAMP(46);
        if (warp >= 2 && warp < 6 && grp_id == 0 && p.stats_out && m0 + row < p.M) {
            const float2* sx = reinterpret_cast<const float2*>(panel + STATX_OFF);
            float mean = sx[row].x, m2 = sx[row].y, n = (float)W;
#pragma unroll
            for (int gg = 1; gg < G; ++gg) {             // Chan's formula, slice after slice
                const float2 o = sx[gg * 128 + row];
                const float dm = o.x - mean, nn = n + (float)W;
                mean += dm * ((float)W / nn);
                m2 += o.y + dm * dm * (n * (float)W / nn);
                n = nn;
            }
            p.stats_out[m0 + row] = make_float2(mean, rsqrtf(m2 / (float)DM + p.eps));
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (threadIdx.x == 64) DIT_STAMP(60);
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm(tmem_base, 512);
    }
}

// =====================================================================================================================
// dit_attn_kernel  (xd_dit_ln_qkv_attn_bf16_tc):   reference score_networks/dit.py:46-51, layers/attention.py:350-375
//     a = bf16( LN(h) * (1 + scale1) + shift1 )                        LayerNorm (no affine) + adaLN modulate -> smem panel
//     [q_h | k_h | v_h] = a Wqkv_h^T + b_h   for each head h           one 256 x 192 tcgen05 tile per head (weights packed per head)
//     O[:, 64h : 64h + 64] = softmax(q_h k_h^T / sqrt(64)) v_h         per image (16 tokens): mma.sync m16n8k16 by the epilogue warps
//   replaces 3 launches (LayerNorm-modulate, qkv GEMM, attention) and the [M, 1152] qkv round trip.
// Work item = (256-row tile, group of heads): at small M the heads of a tile are spread over several CTA pairs.
// TMEM: two 192-column accumulators (head i + 1 is on the tensor pipe while head i is drained and attended).
constexpr int STAGE_BYTES = 4 * 3 * 4096;   // per TMEM lane quadrant: Q, K, V tiles of 32 rows x 64 columns bf16
constexpr int RING_A = 6;
constexpr int SMEM_A_BYTES = 1024 + PANEL_BYTES + STAGE_BYTES + RING_A * SLOT_BYTES + MISC_BYTES;
static_assert(SMEM_A_BYTES <= 232448, "shared memory budget");

struct AttnFParams {
    int M, rows_per_mod, heads, heads_per_item, groups;
    const uint8_t* w;           // packed qkv weights (bytes), for the L2 prefetch at kernel entry
    long long* prof;            // instrumented builds: clock64() stamps (see MlpParams)
    long long ldh, mod_ld, ldo;
    const float* h;
    const float2* stats;        // (mean, rstd) per row from the producer of h, or nullptr: computed here
    const float *shift, *scale;
    const float* bias;          // packed like the weights: [heads][q(64) | k(64) | v(64)]
    bf16* out;
    float eps, sm_scale;
    const int* mod_rows;        // optional indirection of the modulation rows (see MlpParams)
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// One warp, one image: Q, K, V tiles of 16 rows x 64 columns bf16 in shared memory (128-byte rows, 16-byte chunk c of row r
// at c ^ (r & 7)); O = softmax(scale * Q K^T) V is staged over the Q tile and stored as full 128-byte rows.  Same fragment
// algebra as attention16_mma_kernel (attention.cu): the S accumulators are re-used as the A operand of P V.
__device__ __forceinline__ void attend16(uint32_t aQ, uint32_t aK, uint32_t aV, float scale, bf16* op, long long ldo, int rows_ok,
                                         int lane) {
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
        {
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8, c = 2 * ks + (lane >> 4);
            ldsm_x4(aQ + row * 128 + ((c ^ (row & 7)) << 4), a0, a1, a2, a3);
        }
        {
            const int row = (lane & 7) + (lane >> 4) * 8, c = 2 * ks + ((lane >> 3) & 1);
            ldsm_x4(aK + row * 128 + ((c ^ (row & 7)) << 4), b0, b1, b2, b3);
        }
        mma_bf16_16816(s0, a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(s1, a0, a1, a2, a3, b2, b3);
    }
    float m_lo = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])), m_hi = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    const float c = scale * 1.4426950408889634f;
    s0[0] = exp2f((s0[0] - m_lo) * c); s0[1] = exp2f((s0[1] - m_lo) * c); s1[0] = exp2f((s1[0] - m_lo) * c); s1[1] = exp2f((s1[1] - m_lo) * c);
    s0[2] = exp2f((s0[2] - m_hi) * c); s0[3] = exp2f((s0[3] - m_hi) * c); s1[2] = exp2f((s1[2] - m_hi) * c); s1[3] = exp2f((s1[3] - m_hi) * c);
    float l_lo = s0[0] + s0[1] + s1[0] + s1[1], l_hi = s0[2] + s0[3] + s1[2] + s1[3];
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const uint32_t pa0 = f2_to_bf2(s0[0], s0[1]), pa1 = f2_to_bf2(s0[2], s0[3]);
    const uint32_t pa2 = f2_to_bf2(s1[0], s1[1]), pa3 = f2_to_bf2(s1[2], s1[3]);
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt += 2) {
        uint32_t b0, b1, b2, b3;
        const int row = (lane & 7) + ((lane >> 3) & 1) * 8, cc = nt + (lane >> 4);
        ldsm_x4_trans(aV + row * 128 + ((cc ^ (row & 7)) << 4), b0, b1, b2, b3);
#pragma unroll
        for (int j = 0; j < 4; ++j) { o[nt][j] = 0.f; o[nt + 1][j] = 0.f; }
        mma_bf16_16816(o[nt], pa0, pa1, pa2, pa3, b0, b1);
        mma_bf16_16816(o[nt + 1], pa0, pa1, pa2, pa3, b2, b3);
    }
    const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
    __syncwarp();                                         // every lane is done reading Q: stage O over it
    {
        const int r = lane >> 2, q4 = (lane & 3) * 4;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            ptx::sts32(aQ + r * 128 + ((nt ^ (r & 7)) << 4) + q4, f2_to_bf2(o[nt][0] * i_lo, o[nt][1] * i_lo));
            ptx::sts32(aQ + (r + 8) * 128 + ((nt ^ ((r + 8) & 7)) << 4) + q4, f2_to_bf2(o[nt][2] * i_hi, o[nt][3] * i_hi));
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i, row = idx >> 3, cch = idx & 7;
        const uint4 v = ptx::lds128(aQ + row * 128 + ((cch ^ (row & 7)) << 4));
        if (row < rows_ok) *reinterpret_cast<uint4*>(op + (long long)row * ldo + cch * 8) = v;
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
dit_attn_kernel(const __grid_constant__ CUtensorMap tmW, const AttnFParams p) {
    pdl_launch_dependents();
    if (threadIdx.x == 64) DIT_STAMP(63);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* panel = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage = panel + PANEL_BYTES;
    uint8_t* ring = stage + STAGE_BYTES;
    uint8_t* misc = ring + RING_A * SLOT_BYTES;
    uint64_t* ring_full = reinterpret_cast<uint64_t*>(misc);     // [RING_A]
    uint64_t* ring_empty = ring_full + RING_A;                   // [RING_A]
    uint64_t* a_full = ring_empty + RING_A;                      // LN panel written (16 warp arrivals at the leader)
    uint64_t* acc_full = a_full + 1;                             // [2]
    uint64_t* acc_empty = acc_full + 2;                          // [2], 16 warp arrivals at the leader
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = (int)ptx::cluster_ctarank();
    const int cid = blockIdx.x >> 1;
    const int tile = cid / p.groups, hg = cid - tile * p.groups;
    const int m0 = tile * 256 + rank * 128;
    const int hd0 = hg * p.heads_per_item;
    const int nh = min(p.heads_per_item, p.heads - hd0);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmW);
        for (int s = 0; s < RING_A; ++s) {
            ptx::mbar_init(&ring_full[s], 2);
            ptx::mbar_init(&ring_empty[s], 1);
        }
        ptx::mbar_init(a_full, 2 * EPI_WARPS);
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&acc_full[b], 1);
            ptx::mbar_init(&acc_empty[b], 2 * EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_2sm(tmem_ptr, 512);
        ptx::tmem_relinquish_2sm();
    }
    if (warp == 2 && lane == 0) prefetch_weights_l2(p.w, (uint32_t)p.heads * 192 * DM * 2);
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (threadIdx.x == 64) DIT_STAMP(61);
    // the producer streams weights only (no dependence on the preceding kernel): griddepcontrol.wait is executed by the
    // epilogue warps, the only ones that touch activations

    if (warp == 0) {
        if (lane == 0) {                                 // TMA producer: this CTA's 96 rows of every (head, k-block) weight tile
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nh; ++i)
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&ring_empty[s], ph ^ 1);
                    ptx::mbar_arrive_expect_tx_leader(&ring_full[s], SLOT_BYTES);
                    ptx::tma_load_2d_2sm(ring + s * SLOT_BYTES, &tmW, &ring_full[s], kb * 64, (hd0 + i) * 192 + rank * 96);
                    if (++s == RING_A) { s = 0; ph ^= 1; }
                }
        }
    } else if (warp == 1) {
        if (rank == 0) {                                 // MMA issuer
            constexpr uint32_t idesc192 = ptx::idesc_bf16_f32(256, 192);
            const uint64_t d_panel = ptx::smem_desc_sw128(ptx::smem_u32(panel));
            const uint64_t d_ring = ptx::smem_desc_sw128(ptx::smem_u32(ring));
            int s = 0;
            uint32_t ph = 0;
            if (lane == 0) DIT_STAMP(48);
            ptx::mbar_wait(a_full, 0);
            ptx::tc_fence_after();
            if (lane == 0) DIT_STAMP(51);
            for (int i = 0; i < nh; ++i) {
                const uint32_t buf = i & 1;
                ptx::mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&ring_full[s], ph);
                    ptx::tc_fence_after();
                    const uint64_t da = d_panel + (uint64_t)((kb * A_BYTES) >> 4);
                    const uint64_t db = d_ring + (uint64_t)((s * SLOT_BYTES) >> 4);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::umma_bf16_2sm(tmem_base + buf * 192, da + 2 * k, db + 2 * k, idesc192, (kb | k) ? 1u : 0u);
                        ptx::umma_commit_2sm(&ring_empty[s]);
                        if (kb == KB - 1) ptx::umma_commit_2sm(&acc_full[buf]);
                    }
                    __syncwarp();
                    if (++s == RING_A) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        const int e_warp = warp - 2;
        const int grp = e_warp >> 2;
        const int q = warp & 3;
        pdl_wait();
        const bool st = e_warp == 0 && lane == 0;
        if (st) DIT_STAMP(0);
        // ---- LayerNorm + modulate of this CTA's 128 rows -> bf16 panel.  Warp e_warp owns rows e_warp*16 .. +15 (one image
        // when rows_per_mod == 16); lane <-> columns (i * 32 + lane) * 4, eight rows per pass with all loads in flight.
        {
            const uint32_t panel_a = ptx::smem_u32(panel);
            const int row0 = e_warp * 16;
            float4 sc[3], sh[3];
            int cur_mod = -1;
            const int ldx = (int)p.ldh;
            const unsigned rpm = (unsigned)p.rows_per_mod;
            const int m_last = p.M - 1;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int rbase = row0 + pass * 8;
                const int mb = m0 + rbase;
                const int valid = p.M - mb;
                const float* xb = p.h + (long long)mb * p.ldh + lane * 4;
                float4 v[8][3];
                float mean[8], rstd[8];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        v[rr][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rr < valid) v[rr][i] = *reinterpret_cast<const float4*>(xb + rr * ldx + i * 128);
                    }
                }
                if (p.stats) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const float2 st = __ldg(p.stats + min(mb + rr, m_last));
                        mean[rr] = st.x; rstd[rr] = st.y;
                    }
                } else {
                    float red[8];
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
                        mean[rr] = red[rr] / (float)DM;
                        float qa = 0.f;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const float dx = v[rr][i].x - mean[rr], dy = v[rr][i].y - mean[rr], dz = v[rr][i].z - mean[rr],
                                        dw = v[rr][i].w - mean[rr];
                            qa += dx * dx + dy * dy + dz * dz + dw * dw;
                        }
                        red[rr] = qa;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) red[rr] += __shfl_xor_sync(0xffffffffu, red[rr], o);
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) rstd[rr] = rsqrtf(red[rr] / (float)DM + p.eps);
                }
                const uint32_t a_lane = panel_a + (lane >> 4) * A_BYTES + ((lane & 1) << 3);
                const uint32_t chunk = (lane & 15) >> 1;
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int r = rbase + rr;
                    const int mod_row = (int)((unsigned)min(mb + rr, m_last) / rpm);
                    if (mod_row != cur_mod) {            // warp-uniform; once per warp when rows_per_mod == 16
                        cur_mod = mod_row;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const long long off = (long long)(p.mod_rows ? __ldg(p.mod_rows + mod_row) : mod_row) * p.mod_ld + (i * 32 + lane) * 4;
                            sc[i] = __ldg(reinterpret_cast<const float4*>(p.scale + off));
                            sh[i] = __ldg(reinterpret_cast<const float4*>(p.shift + off));
                        }
                    }
                    const bool live = rr < valid;
                    const uint32_t a_row = a_lane + r * 128 + ((chunk ^ (r & 7)) << 4);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float y0 = fmaf((v[rr][i].x - mean[rr]) * rstd[rr], 1.0f + sc[i].x, sh[i].x);
                        const float y1 = fmaf((v[rr][i].y - mean[rr]) * rstd[rr], 1.0f + sc[i].y, sh[i].y);
                        const float y2 = fmaf((v[rr][i].z - mean[rr]) * rstd[rr], 1.0f + sc[i].z, sh[i].z);
                        const float y3 = fmaf((v[rr][i].w - mean[rr]) * rstd[rr], 1.0f + sc[i].w, sh[i].w);
                        ptx::sts64(a_row + 2 * i * A_BYTES, live ? f2_to_bf2(y0, y1) : 0u, live ? f2_to_bf2(y2, y3) : 0u);
                    }
                }
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(a_full);
            if (st) DIT_STAMP(1);
        }
        // ---- per head: drain [q | k | v] (+ bias) into the quadrant's bf16 tiles, then one image per warp
        const uint32_t st_a = ptx::smem_u32(stage) + q * (3 * 4096);
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const int img_row0 = m0 + q * 32 + grp * 16;
        for (int i = 0; i < nh; ++i) {
            const int hd = hd0 + i;
            const uint32_t buf = i & 1;
            const float* bptr = p.bias + hd * 192 + grp * 96;
            if (lane < 3) asm volatile("prefetch.global.L1 [%0];" ::"l"(bptr + lane * 32));
            ptx::mbar_wait(&acc_full[buf], (i >> 1) & 1);
            ptx::tc_fence_after();
            if (st) DIT_STAMP(8 + 3 * i);
            uint32_t r[3][32];
#pragma unroll
            for (int j = 0; j < 3; ++j) ptx::tmem_ld_32x32(t_lane + buf * 192 + grp * 96 + j * 32, r[j]);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_leader(&acc_empty[buf]);
            // this warp's 96 columns = 12 chunks of 8: column (grp * 96 + 8 * j) -> tile (col / 64), chunk (col % 64) / 8
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const int col = grp * 96 + 8 * j;
                const uint32_t* rv = r[j >> 2] + 8 * (j & 3);
                const float4 ba = __ldg(reinterpret_cast<const float4*>(bptr) + 2 * j);
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bptr) + 2 * j + 1);
                const uint4 pk = make_uint4(f2_to_bf2(__uint_as_float(rv[0]) + ba.x, __uint_as_float(rv[1]) + ba.y),
                                            f2_to_bf2(__uint_as_float(rv[2]) + ba.z, __uint_as_float(rv[3]) + ba.w),
                                            f2_to_bf2(__uint_as_float(rv[4]) + bb.x, __uint_as_float(rv[5]) + bb.y),
                                            f2_to_bf2(__uint_as_float(rv[6]) + bb.z, __uint_as_float(rv[7]) + bb.w));
                ptx::sts128(st_a + (col >> 6) * 4096 + lane * 128 + ((((col & 63) >> 3) ^ (lane & 7)) << 4), pk);
            }
            ptx::named_bar_sync(1 + q, 64);              // both column halves of the quadrant's 32 rows are staged
            if (st) DIT_STAMP(9 + 3 * i);
            const uint32_t io = grp * 2048;              // image `grp` of the quadrant: rows 16 * grp ..
            attend16(st_a + io, st_a + 4096 + io, st_a + 8192 + io, p.sm_scale,
                     p.out + (long long)img_row0 * p.ldo + hd * 64, p.ldo, p.M - img_row0, lane);
            ptx::named_bar_sync(1 + q, 64);              // the partner warp is done with the tiles before they are rewritten
            if (st) DIT_STAMP(10 + 3 * i);
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (threadIdx.x == 64) DIT_STAMP(60);
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm(tmem_base, 512);
    }
}

int tmap_2d_bf16(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    return make_tmap(tm, ptr, 2, dims, str, box);
}

}  // namespace

// h_out <- h1 + gate2 * (gelu(LNmod(h1) W1^T + b1) W2^T + b2),  h1 = h_in + gate1 * (O Wp^T + bp)   (fp32 rows of D = 384)
// LNmod(x)[m, :] = LN(x[m, :]) * (1 + scale2[m / rows_per_mod]) + shift2[m / rows_per_mod]; gate / shift / scale rows are
// mod_ld floats apart.  stats_out (optional) receives (mean, rstd) of every new row.  h_out may be h_in (in place) -- then one
// CTA pair owns each 256-row tile; with DISTINCT buffers and few tiles (small M: a shard of a strong-scaled batch) the hidden
// dimension of a tile is split over `split` CTA pairs of one cluster (2, 3 or 4; 0 = choose from M and the SM count).
namespace {
template <int G>
int launch_mlp(const CUtensorMap& tO, const CUtensorMap& tWp, const CUtensorMap& tW1, const CUtensorMap& tW2,
               const CUtensorMap& tHin, const CUtensorMap& tH, const MlpParams& p, int tiles, cudaStream_t stream, bool query,
               int* max_clusters) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dit_mlp_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute(max dynamic smem) failed");
            return XD_ERR_CUDA;
        }
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * G * tiles);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    if (query) {                                         // how many clusters of 2G CTAs can be resident at once
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(max_clusters, dit_mlp_kernel<G>, &cfg) != cudaSuccess) {
            cudaGetLastError();
            *max_clusters = 0;
        }
        return XD_OK;
    }
    cfg.numAttrs = xd_pdl_enabled_gemm() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, dit_mlp_kernel<G>, tO, tWp, tW1, tW2, tHin, tH, p) != cudaSuccess) {
        xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
        return XD_ERR_CUDA;
    }
    return XD_OK;
}
}  // namespace

static int dit_proj_mlp_impl(const void* O, long long ldo, const void* Wp, const float* bp, const void* W1,
                             const float* b1, const void* W2, const float* b2, int hidden, const float* h_in,
                             float* h_out, long long ldh, int M, int D, const float* gate1, const float* shift2,
                             const float* scale2, const float* gate2, long long mod_ld, int rows_per_mod,
                             float eps, float* stats_out, int split, const int* mod_rows, void* stream) {
    XD_CHECK_ARG(O && Wp && bp && W1 && b1 && W2 && b2 && h_in && h_out && gate1 && shift2 && scale2 && gate2 && M > 0);
    XD_CHECK_ARG(D == DM && hidden % 128 == 0 && hidden >= 128 && rows_per_mod > 0);
    XD_CHECK_ARG(ldo % 8 == 0 && ldh % 4 == 0 && mod_ld % 4 == 0);
    XD_CHECK_ARG(aligned16(O) && aligned16(Wp) && aligned16(W1) && aligned16(W2) && aligned16(h_in) && aligned16(h_out) &&
                 aligned16(bp) && aligned16(b1) && aligned16(b2) && aligned16(gate1) && aligned16(shift2) &&
                 aligned16(scale2) && aligned16(gate2) && (reinterpret_cast<uintptr_t>(stats_out) & 7) == 0);
    XD_CHECK_ARG(split >= 0 && split <= 4 && (split <= 1 || (h_in != h_out && hidden % (128 * split) == 0)));
    CUtensorMap tO, tWp, tW1, tW2, tHin, tH;
    int rc;
    if ((rc = tmap_2d_bf16(&tO, O, M, DM, ldo, 128))) return rc;
    if ((rc = tmap_2d_bf16(&tWp, Wp, DM, DM, DM, 96))) return rc;
    if ((rc = tmap_2d_bf16(&tW1, W1, hidden, DM, DM, 64))) return rc;
    if ((rc = tmap_2d_bf16(&tW2, W2, DM, hidden, hidden, 96))) return rc;
    if ((rc = tmap_epi(&tHin, h_in, M, DM, ldh, true))) return rc;
    if ((rc = tmap_epi(&tH, h_out, M, DM, ldh, true))) return rc;
    const int tiles = (M + 255) / 256;
    long long* prof = nullptr;
#ifdef XDB200_INSTRUMENT
    if (const char* e = getenv("XDB200_DIT_PROF")) prof = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
#endif
    MlpParams p{M, hidden, rows_per_mod, mod_ld, bp, b1, b2, gate1, shift2, scale2, gate2,
                reinterpret_cast<float2*>(stats_out), eps, h_out, ldh, prof,
                static_cast<const uint8_t*>(Wp), static_cast<const uint8_t*>(W1), static_cast<const uint8_t*>(W2), mod_rows};
    cudaStream_t st = (cudaStream_t)stream;
    int G = split;
    if (G == 0) {
        // widest split whose clusters are all resident at once (one wave) -- only worth it when the tiles alone leave most
        // SM pairs idle.  Cluster capacity is a property of the device: queried once per G.
        static int cap[5] = {0, 0, -1, -1, -1};
        G = 1;
        if (h_in != h_out && xd_split_enabled()) {       // (a split changes the fc2 summation order with the row count)
            for (int g = 2; g <= 4; ++g) {
                if (hidden % (128 * g)) continue;
                if (cap[g] < 0) {
                    int n = 0;
                    if (g == 2) launch_mlp<2>(tO, tWp, tW1, tW2, tHin, tH, p, tiles, st, true, &n);
                    if (g == 3) launch_mlp<3>(tO, tWp, tW1, tW2, tHin, tH, p, tiles, st, true, &n);
                    if (g == 4) launch_mlp<4>(tO, tWp, tW1, tW2, tHin, tH, p, tiles, st, true, &n);
                    cap[g] = n;
                }
                if (tiles <= cap[g]) G = g;
            }
        }
    }
    switch (G) {
        case 2: return launch_mlp<2>(tO, tWp, tW1, tW2, tHin, tH, p, tiles, st, false, nullptr);
        case 3: return launch_mlp<3>(tO, tWp, tW1, tW2, tHin, tH, p, tiles, st, false, nullptr);
        case 4: return launch_mlp<4>(tO, tWp, tW1, tW2, tHin, tH, p, tiles, st, false, nullptr);
        default: return launch_mlp<1>(tO, tWp, tW1, tW2, tHin, tH, p, tiles, st, false, nullptr);
    }
}

// O[m, 64h : 64h + 64] = softmax_per_image( q_h k_h^T / sqrt(64) ) v_h  with [q_h | k_h | v_h] = LNmod(h) Wqkv_h^T + b_h:
extern "C" int xd_dit_proj_mlp_bf16_tc(const void* O, long long ldo, const void* Wp, const float* bp, const void* W1,
                                       const float* b1, const void* W2, const float* b2, int hidden, const float* h_in,
                                       float* h_out, long long ldh, int M, int D, const float* gate1, const float* shift2,
                                       const float* scale2, const float* gate2, long long mod_ld, int rows_per_mod,
                                       float eps, float* stats_out, int split, void* stream) {
    return dit_proj_mlp_impl(O, ldo, Wp, bp, W1, b1, W2, b2, hidden, h_in, h_out, ldh, M, D, gate1, shift2, scale2, gate2, mod_ld,
                             rows_per_mod, eps, stats_out, split, nullptr, stream);
}
// ... with the modulation rows of image i taken from row mod_rows[i] of a table (device int32 [M / rows_per_mod])
extern "C" int xd_dit_proj_mlp_bf16_tc_rows(const void* O, long long ldo, const void* Wp, const float* bp, const void* W1,
                                            const float* b1, const void* W2, const float* b2, int hidden, const float* h_in,
                                            float* h_out, long long ldh, int M, int D, const float* gate1,
                                            const float* shift2, const float* scale2, const float* gate2, long long mod_ld,
                                            int rows_per_mod, float eps, float* stats_out, int split, const int* mod_rows,
                                            void* stream) {
    XD_CHECK_ARG(mod_rows != nullptr);
    return dit_proj_mlp_impl(O, ldo, Wp, bp, W1, b1, W2, b2, hidden, h_in, h_out, ldh, M, D, gate1, shift2, scale2, gate2, mod_ld,
                             rows_per_mod, eps, stats_out, split, mod_rows, stream);
}

// LayerNorm-modulate + QKV projection + attention of a DiT block in one kernel.  h fp32 [M, 384] (rows_per_mod = tokens per
// image = 16), Wh bf16 [heads * 192, 384] and bias fp32 [heads * 192] packed per head as [q | k | v] (64 rows each), stats
// (mean, rstd) per row or NULL (computed in the kernel), out bf16 [M, 384] with head h in columns [64h, 64h + 64).
static int dit_ln_qkv_attn_impl(const float* h, long long ldh, const float* stats, const float* shift,
                                const float* scale, long long mod_ld, int rows_per_mod, float eps, const void* Wh,
                                const float* bias, int heads, int M, int D, float sm_scale, void* out,
                                long long ldo, const int* mod_rows, void* stream) {
    XD_CHECK_ARG(h && shift && scale && Wh && bias && out && M > 0 && D == DM && heads * 64 == DM);
    XD_CHECK_ARG(rows_per_mod == 16 && M % 16 == 0);                  // one image = 16 token rows = one attention problem
    XD_CHECK_ARG(ldh % 4 == 0 && mod_ld % 4 == 0 && ldo % 8 == 0);
    XD_CHECK_ARG(aligned16(h) && aligned16(shift) && aligned16(scale) && aligned16(Wh) && aligned16(bias) && aligned16(out) &&
                 (reinterpret_cast<uintptr_t>(stats) & 7) == 0);
    CUtensorMap tW;
    int rc;
    if ((rc = tmap_2d_bf16(&tW, Wh, (long long)heads * 192, DM, DM, 96))) return rc;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dit_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_A_BYTES) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute(max dynamic smem) failed");
            return XD_ERR_CUDA;
        }
        configured = true;
    }
    // heads of a tile are spread over `groups` CTA pairs when there are fewer tiles than SM pairs
    const int tiles = (M + 255) / 256;
    const int pairs = sm_count() / 2;
    int groups = 1;
    for (int g : {2, 3, 6})
        if (heads % g == 0 && tiles * g <= pairs) groups = g;
    long long* prof = nullptr;
#ifdef XDB200_INSTRUMENT
    if (const char* e = getenv("XDB200_DIT_PROF")) prof = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
#endif
    AttnFParams p{M, rows_per_mod, heads, heads / groups, groups, static_cast<const uint8_t*>(Wh), prof, ldh, mod_ld, ldo, h, reinterpret_cast<const float2*>(stats),
                  shift, scale, bias, (bf16*)out, eps, sm_scale, mod_rows};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * tiles * groups);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_A_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = xd_pdl_enabled_gemm() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, dit_attn_kernel, tW, p) != cudaSuccess) {
        xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
        return XD_ERR_CUDA;
    }
    return XD_OK;
}

extern "C" int xd_dit_ln_qkv_attn_bf16_tc(const float* h, long long ldh, const float* stats, const float* shift,
                                          const float* scale, long long mod_ld, int rows_per_mod, float eps, const void* Wh,
                                          const float* bias, int heads, int M, int D, float sm_scale, void* out,
                                          long long ldo, void* stream) {
    return dit_ln_qkv_attn_impl(h, ldh, stats, shift, scale, mod_ld, rows_per_mod, eps, Wh, bias, heads, M, D, sm_scale, out, ldo,
                                nullptr, stream);
}
extern "C" int xd_dit_ln_qkv_attn_bf16_tc_rows(const float* h, long long ldh, const float* stats, const float* shift,
                                               const float* scale, long long mod_ld, int rows_per_mod, float eps,
                                               const void* Wh, const float* bias, int heads, int M, int D, float sm_scale,
                                               void* out, long long ldo, const int* mod_rows, void* stream) {
    XD_CHECK_ARG(mod_rows != nullptr);
    return dit_ln_qkv_attn_impl(h, ldh, stats, shift, scale, mod_ld, rows_per_mod, eps, Wh, bias, heads, M, D, sm_scale, out, ldo,
                                mod_rows, stream);
}
