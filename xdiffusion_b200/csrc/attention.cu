// Fused softmax attention, head dim 64, bf16 in / bf16 out, fp32 math, online softmax.
// One thread owns one query row (q and the output accumulator live in registers); K/V chunks of
// 16 keys are staged in shared memory as fp32 and broadcast to the rows that share them.  A CTA
// holds 128 query rows: one (batch, head) slab when Tq >= 128, else 128/Tq (batch, head) pairs.
//
// Element (b, h, row, d) of q/k/v/o lives at  ptr + b*bs + h*hs + row*rs + d  (elements), which
// covers every layout on the hot path:
//   UNet  QKVAttention            qkv[B*T, 3C], per-head interleaved  (layers/attention.py:161-188)
//   DiT   MultiHeadSelfAttention  qkv[B*T, 3D], [Q|K|V]-major         (layers/attention.py:350-380)
//   PixArt LastChannelCrossAttention  q[B*16, D], kv[B*77, 2D]        (layers/attention.py:209-228)
//   video TemporalSelfAttention: + relative-position logits q.E_k[h, j-i+L-1], no scale, and the
//       reference's raw (B,H,L,D)->(B,H*D,L) reinterpretation on store (layers/attention.py:551-676)
#include "common.cuh"

namespace {

constexpr int D = 64;
constexpr int KC = 16;          // keys per shared-memory chunk
constexpr int ROWS = 128;       // query rows (threads) per CTA

struct AttnParams {
    const bf16 *q, *k, *v;
    bf16* o;
    long long q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs;
    int B, H, Tq, Tk;
    float scale;
    const float* relk;          // optional [H][2*Tk-1][64] relative-position key table
    int scramble;               // temporal quirk: out viewed as (H*64, L) from a (H, L, 64) buffer
    long long o_cs;             // scramble: element stride between "channels" of the (C, L) view
    int in_f32;                 // q/k/v are fp32 (T = 16 relative-position path only)
    int hpg;                    // heads per group: index h = g * hpg + head; `head` selects the relk table / scramble row,
    long long o_gs;             // g (e.g. the pixel of a clip) adds g * o_gs to the scrambled store.  hpg == H: no groups
};

// q . k over 64 dims with four independent accumulators (a single 64-long FMA chain is latency-bound)
__device__ __forceinline__ float dot64(const float* q, const float4* k) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 t = k[d4];
        a0 = fmaf(q[4 * d4], t.x, a0); a1 = fmaf(q[4 * d4 + 1], t.y, a1);
        a2 = fmaf(q[4 * d4 + 2], t.z, a2); a3 = fmaf(q[4 * d4 + 3], t.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}
__device__ __forceinline__ float dot64_ldg(const float* q, const float4* k) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 t = __ldg(k + d4);
        a0 = fmaf(q[4 * d4], t.x, a0); a1 = fmaf(q[4 * d4 + 1], t.y, a1);
        a2 = fmaf(q[4 * d4 + 2], t.z, a2); a3 = fmaf(q[4 * d4 + 3], t.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

__device__ __forceinline__ void store_row(const AttnParams& p, int b, int h, int row, const float* o, float inv) {
    if (!p.scramble) {
        bf16* op = p.o + b * p.o_bs + h * p.o_hs + (long long)row * p.o_rs;
#pragma unroll
        for (int d = 0; d < D; d += 8) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = o[d + i] * inv;
            *reinterpret_cast<bf16x8*>(op + d) = pack8(f);
        }
    } else {
        // a[b, h, row, d] is element (h*L + row)*64 + d of a flat buffer that the reference then
        // views as (C = H*64, L): flat = c*L + l.
        const int L = p.Tq;
        const int g = h / p.hpg, hh = h - g * p.hpg;
        bf16* ob = p.o + b * p.o_bs + g * p.o_gs;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int flat = (hh * L + row) * D + d;
            const int c = flat / L, lpos = flat % L;
            ob[(long long)lpos * p.o_rs + (long long)c * p.o_cs] = __float2bfloat16_rn(o[d] * inv);
        }
    }
}

// Tq == Tk == 16 (DiT / PixArt tokens, video frames): 8 (batch, head) pairs per CTA, thread = query
// row; the same thread also fetches key row / value row `row` of its pair, so all 24 16-byte loads
// of a thread are in flight together, and the 16 logits of a row live in registers (exact softmax).
constexpr int T16 = 16;
constexpr int RS16 = D + 4;                  // padded key/value row: 8 consecutive rows cover all 32 banks
constexpr int PAIR_LD = 2 * T16 * RS16 + 4;  // +16 B: the two pairs of a warp hit different banks
__global__ void __launch_bounds__(ROWS)
attention16_kernel(const AttnParams p) {
    pdl_prologue();
    extern __shared__ float smem[];             // [8][2][16][64] (+pad) | [hpg][31][64] (+pad) relative-position tables
    const int tid = threadIdx.x;
    const int pl = tid >> 4, row = tid & 15;
    // The relative-position key tables (hpg x 31 rows of 64 floats) are staged once per CTA: read straight from global
    // memory every thread of a warp touched a different table row per load (32 cache lines per request) and the kernel
    // spent most of its time there (232 us at 8 clips x 32 x 32 pixels x 2 heads).  A CTA then walks groups of 8 problems.
    float* sR = smem + 8 * PAIR_LD;
    if (p.relk) {
        const int n4 = p.hpg * (2 * T16 - 1) * (D / 4);
        for (int i = tid; i < n4; i += ROWS) {
            const int r = i >> 4, c4 = i & 15;
            *reinterpret_cast<float4*>(sR + r * RS16 + c4 * 4) = __ldg(reinterpret_cast<const float4*>(p.relk) + i);
        }
    }
    const long long ngroups = ((long long)p.B * p.H + 7) / 8;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    if (grp != blockIdx.x) __syncthreads();     // the previous group's keys / values are no longer needed
    const long long bh = grp * 8 + pl;
    const bool active = bh < (long long)p.B * p.H;
    const int b = active ? (int)(bh / p.H) : 0, h = active ? (int)(bh % p.H) : 0;
    float* sK = smem + pl * PAIR_LD;
    float* sV = sK + T16 * RS16;
    float q[D];
    if (p.in_f32) {
        const float* qp = (const float*)p.q + b * p.q_bs + h * p.q_hs + (long long)row * p.q_rs;
        const float* kp = (const float*)p.k + b * p.k_bs + h * p.k_hs + (long long)row * p.k_rs;
        const float* vp = (const float*)p.v + b * p.v_bs + h * p.v_hs + (long long)row * p.v_rs;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(qp + i * 4);
            q[i * 4] = a.x * p.scale; q[i * 4 + 1] = a.y * p.scale; q[i * 4 + 2] = a.z * p.scale; q[i * 4 + 3] = a.w * p.scale;
            *reinterpret_cast<float4*>(sK + row * RS16 + i * 4) = *reinterpret_cast<const float4*>(kp + i * 4);
            *reinterpret_cast<float4*>(sV + row * RS16 + i * 4) = *reinterpret_cast<const float4*>(vp + i * 4);
        }
    } else {
        bf16x8 rq[8], rk[8], rv[8];
        const bf16* qp = p.q + b * p.q_bs + h * p.q_hs + (long long)row * p.q_rs;
        const bf16* kp = p.k + b * p.k_bs + h * p.k_hs + (long long)row * p.k_rs;
        const bf16* vp = p.v + b * p.v_bs + h * p.v_hs + (long long)row * p.v_rs;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            rq[i] = *reinterpret_cast<const bf16x8*>(qp + i * 8);
            rk[i] = *reinterpret_cast<const bf16x8*>(kp + i * 8);
            rv[i] = *reinterpret_cast<const bf16x8*>(vp + i * 8);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float f[8];
            unpack8(rq[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) q[i * 8 + j] = f[j] * p.scale;
            unpack8(rk[i], f);
            *reinterpret_cast<float4*>(sK + row * RS16 + i * 8) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(sK + row * RS16 + i * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
            unpack8(rv[i], f);
            *reinterpret_cast<float4*>(sV + row * RS16 + i * 8) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(sV + row * RS16 + i * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
    __syncthreads();
    float s[T16];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < T16; ++j) {
        float acc = dot64(q, reinterpret_cast<const float4*>(sK + j * RS16));
        if (p.relk)
            acc += dot64(q, reinterpret_cast<const float4*>(sR + ((h % p.hpg) * (2 * T16 - 1) + (j - row + T16 - 1)) * RS16));
        s[j] = acc;
        mx = fmaxf(mx, acc);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < T16; ++j) { s[j] = __expf(s[j] - mx); l += s[j]; }
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < T16; ++j) {
        const float4* vr = reinterpret_cast<const float4*>(sV + j * RS16);
#pragma unroll
        for (int d4 = 0; d4 < D / 4; ++d4) {
            const float4 t = vr[d4];
            o[4 * d4] = fmaf(s[j], t.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(s[j], t.y, o[4 * d4 + 1]);
            o[4 * d4 + 2] = fmaf(s[j], t.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(s[j], t.w, o[4 * d4 + 3]);
        }
    }
    if (active) store_row(p, b, h, row, o, 1.0f / l);
    }
}


// ---- Tq == Tk == 16 without relative positions: one WARP per (batch, head) on mma.sync m16n8k16 ----
// The problem is 16x16x64 twice: far too small for a 128-row tcgen05 tile, and on CUDA cores it costs
// ~3000 instructions per (batch, head).  Here a warp stages Q, K, V (2 KB each, XOR-swizzled 16-byte
// chunks) with coalesced 16-byte loads, runs 8 + 8 warp-level MMAs (S = Q K^T; O = P V with the S
// accumulator fragments re-used as the P operand), does the softmax on fragments with quad shuffles
// and writes O through shared memory as full 128-byte rows: ~150 instructions per (batch, head),
// which leaves the kernel bandwidth-bound.
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
constexpr int W16 = 4;                       // warps (= (batch, head) pairs) per CTA
__global__ void __launch_bounds__(32 * W16)
attention16_mma_kernel(const AttnParams p) {
    pdl_prologue();
    __shared__ __align__(128) uint8_t sm[W16][3][16 * 128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bh = (long long)blockIdx.x * W16 + warp;
    if (bh >= (long long)p.B * p.H) return;
    const int b = (int)(bh / p.H), h = (int)(bh % p.H);
    uint8_t* sQ = sm[warp][0];
    uint8_t* sK = sm[warp][1];
    uint8_t* sV = sm[warp][2];
    const bf16* qp = p.q + b * p.q_bs + h * p.q_hs;
    const bf16* kp = p.k + b * p.k_bs + h * p.k_hs;
    const bf16* vp = p.v + b * p.v_bs + h * p.v_hs;
    {   // 16 rows x 8 chunks of 16 B per tile: 4 chunks per lane per tile, all 12 loads in flight
        uint4 rq[4], rk[4], rv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            rq[i] = *reinterpret_cast<const uint4*>(qp + (long long)row * p.q_rs + c * 8);
            rk[i] = *reinterpret_cast<const uint4*>(kp + (long long)row * p.k_rs + c * 8);
            rv[i] = *reinterpret_cast<const uint4*>(vp + (long long)row * p.v_rs + c * 8);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            const int off = row * 128 + ((c ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(sQ + off) = rq[i];
            *reinterpret_cast<uint4*>(sK + off) = rk[i];
            *reinterpret_cast<uint4*>(sV + off) = rv[i];
        }
    }
    __syncwarp();
    const uint32_t aQ = (uint32_t)__cvta_generic_to_shared(sQ), aK = (uint32_t)__cvta_generic_to_shared(sK),
                   aV = (uint32_t)__cvta_generic_to_shared(sV);
    // S = Q K^T: two 8-key n-tiles, four 16-wide k-steps over d
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
        {   // A: Q rows (lane%8) + 8*((lane/8)%2), chunk 2*ks + lane/16
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8, c = 2 * ks + (lane >> 4);
            ldsm_x4(aQ + row * 128 + ((c ^ (row & 7)) << 4), a0, a1, a2, a3);
        }
        {   // B: K rows (keys) (lane%8) + 8*(lane/16), chunk 2*ks + (lane/8)%2
            const int row = (lane & 7) + (lane >> 4) * 8, c = 2 * ks + ((lane >> 3) & 1);
            ldsm_x4(aK + row * 128 + ((c ^ (row & 7)) << 4), b0, b1, b2, b3);
        }
        mma_bf16_16816(s0, a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(s1, a0, a1, a2, a3, b2, b3);
    }
    // softmax over the 16 keys of rows r = lane/4 (c0, c1) and r + 8 (c2, c3); a row lives in one quad
    float m_lo = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])), m_hi = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    const float c = p.scale * 1.4426950408889634f;
    s0[0] = exp2f((s0[0] - m_lo) * c); s0[1] = exp2f((s0[1] - m_lo) * c); s1[0] = exp2f((s1[0] - m_lo) * c); s1[1] = exp2f((s1[1] - m_lo) * c);
    s0[2] = exp2f((s0[2] - m_hi) * c); s0[3] = exp2f((s0[3] - m_hi) * c); s1[2] = exp2f((s1[2] - m_hi) * c); s1[3] = exp2f((s1[3] - m_hi) * c);
    float l_lo = s0[0] + s0[1] + s1[0] + s1[1], l_hi = s0[2] + s0[3] + s1[2] + s1[3];
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    // P as the A operand of the second MMA: the accumulator fragments are already in A layout
    const uint32_t pa0 = f2_to_bf2(s0[0], s0[1]), pa1 = f2_to_bf2(s0[2], s0[3]);
    const uint32_t pa2 = f2_to_bf2(s1[0], s1[1]), pa3 = f2_to_bf2(s1[2], s1[3]);
    // O = P V: eight 8-wide n-tiles over d, one k-step (16 keys); V^T fragments via ldmatrix.trans
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
    __syncwarp();                                         // everyone is done reading sQ: reuse it for O
    {
        const int r = lane >> 2, q4 = (lane & 3) * 4;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            *reinterpret_cast<uint32_t*>(sQ + r * 128 + ((nt ^ (r & 7)) << 4) + q4) = f2_to_bf2(o[nt][0] * i_lo, o[nt][1] * i_lo);
            *reinterpret_cast<uint32_t*>(sQ + (r + 8) * 128 + ((nt ^ ((r + 8) & 7)) << 4) + q4) =
                f2_to_bf2(o[nt][2] * i_hi, o[nt][3] * i_hi);
        }
    }
    __syncwarp();
    bf16* op = p.o + b * p.o_bs + h * p.o_hs;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i, row = idx >> 3, cch = idx & 7;
        *reinterpret_cast<uint4*>(op + (long long)row * p.o_rs + cch * 8) =
            *reinterpret_cast<const uint4*>(sQ + row * 128 + ((cch ^ (row & 7)) << 4));
    }
}

// ---- Tq == Tk == 16 with relative-position keys, fp32 q / k / v, scrambled store (the temporal attention of the video UNet):
// one WARP per (clip, pixel, head) on mma.sync m16n8k16.  The logits are NOT scaled in the reference, so bf16 operand
// rounding would be amplified by the softmax: every fp32 operand is split into bf16 hi + lo and each product is three MMAs
// (hi.hi + hi.lo + lo.hi, error ~2^-17), the same scheme as the split-precision qkv projection that feeds this kernel.
//   S = Q K^T (16 x 16),  R = Q E^T (16 x 31, E = this head's relative-position table),  logit[i][j] = S[i][j] + R[i][j - i + 15]
//   (R goes through a 2 KB shared-memory scratch for the diagonal gather),  exact softmax,  O = P V,
//   store a[b, h, row, d] as element (h * 16 + row) * 64 + d of the (C, L) view: frame d % 16, channel h * 64 + row * 4 + d / 16,
//   i.e. one contiguous 128-byte row of 64 channels per frame.
// Replaces the thread-per-row CUDA-core kernel (attention16_kernel), which was shared-memory bound.
constexpr int WRP = 4;                                   // warps (problems in flight) per CTA
constexpr int RP_WARP_BYTES = 6 * 2048 + 2176;           // Q, K, V as hi / lo tiles + the R scratch ([16][33] floats, padded to 128 B)
__device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* lo, int row, int c4, float4 x) {
    const bf16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y), h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
    const int off = row * 128 + (((c4 >> 1) ^ (row & 7)) << 4) + (c4 & 1) * 8;
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(f2_to_bf2(x.x, x.y), f2_to_bf2(x.z, x.w));
    *reinterpret_cast<uint2*>(lo + off) = make_uint2(
        f2_to_bf2(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1)), f2_to_bf2(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3)));
}
__global__ void __launch_bounds__(32 * WRP)
attention16_relpos_mma_kernel(const AttnParams p) {
    pdl_prologue();
    extern __shared__ __align__(128) uint8_t smr[];      // [WRP][RP_WARP_BYTES] | [hpg][hi, lo][32 rows x 128 B]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sE = smr + WRP * RP_WARP_BYTES;
    for (int i = threadIdx.x; i < p.hpg * 32 * 16; i += 32 * WRP) {
        const int t = i >> 9, r = (i >> 4) & 31, c4 = i & 15;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < 2 * T16 - 1) x = __ldg(reinterpret_cast<const float4*>(p.relk + ((long long)t * (2 * T16 - 1) + r) * D) + c4);
        split_store(sE + t * 8192, sE + t * 8192 + 4096, r, c4, x);
    }
    __syncthreads();
    uint8_t* sQh = smr + warp * RP_WARP_BYTES;
    uint8_t *sQl = sQh + 2048, *sKh = sQh + 4096, *sKl = sQh + 6144, *sVh = sQh + 8192, *sVl = sQh + 10240;
    float* sR = reinterpret_cast<float*>(sQh + 12288);
    const uint32_t aQh = (uint32_t)__cvta_generic_to_shared(sQh), aQl = aQh + 2048, aKh = aQh + 4096, aKl = aQh + 6144,
                   aVh = aQh + 8192, aVl = aQh + 10240, aE = (uint32_t)__cvta_generic_to_shared(sE);
    const long long total = (long long)p.B * p.H;
    for (long long bh = (long long)blockIdx.x * WRP + warp; bh < total; bh += (long long)gridDim.x * WRP) {
        const int b = (int)(bh / p.H), h = (int)(bh % p.H);
        const int g = h / p.hpg, hh = h - g * p.hpg;
        const float* qp = (const float*)p.q + b * p.q_bs + h * p.q_hs;
        const float* kp = (const float*)p.k + b * p.k_bs + h * p.k_hs;
        const float* vp = (const float*)p.v + b * p.v_bs + h * p.v_hs;
        {   // 16 rows x 16 float4 per tensor: 8 per lane, all 24 loads in flight
            float4 rq[8], rk[8], rv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = lane + 32 * i, row = idx >> 4, c4 = idx & 15;
                rq[i] = *reinterpret_cast<const float4*>(qp + (long long)row * p.q_rs + c4 * 4);
                rk[i] = *reinterpret_cast<const float4*>(kp + (long long)row * p.k_rs + c4 * 4);
                rv[i] = *reinterpret_cast<const float4*>(vp + (long long)row * p.v_rs + c4 * 4);
            }
            __syncwarp();                                 // the previous problem's tiles are no longer read
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = lane + 32 * i, row = idx >> 4, c4 = idx & 15;
                rq[i].x *= p.scale; rq[i].y *= p.scale; rq[i].z *= p.scale; rq[i].w *= p.scale;
                split_store(sQh, sQl, row, c4, rq[i]);
                split_store(sKh, sKl, row, c4, rk[i]);
                split_store(sVh, sVl, row, c4, rv[i]);
            }
        }
        __syncwarp();
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, r[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) r[nt][j] = 0.f;
        const uint32_t aEh = aE + hh * 8192, aEl = aEh + 4096;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t ah[4], al[4], bh4[4], bl4[4];
            {
                const int row = (lane & 7) + ((lane >> 3) & 1) * 8, c = 2 * ks + (lane >> 4);
                const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
                ldsm_x4(aQh + off, ah[0], ah[1], ah[2], ah[3]);
                ldsm_x4(aQl + off, al[0], al[1], al[2], al[3]);
            }
            const int brow = (lane & 7) + (lane >> 4) * 8, bc = 2 * ks + ((lane >> 3) & 1);
            {
                const uint32_t off = brow * 128 + ((bc ^ (brow & 7)) << 4);
                ldsm_x4(aKh + off, bh4[0], bh4[1], bh4[2], bh4[3]);
                ldsm_x4(aKl + off, bl4[0], bl4[1], bl4[2], bl4[3]);
            }
            mma_bf16_16816(s0, ah[0], ah[1], ah[2], ah[3], bh4[0], bh4[1]);
            mma_bf16_16816(s1, ah[0], ah[1], ah[2], ah[3], bh4[2], bh4[3]);
            mma_bf16_16816(s0, ah[0], ah[1], ah[2], ah[3], bl4[0], bl4[1]);
            mma_bf16_16816(s1, ah[0], ah[1], ah[2], ah[3], bl4[2], bl4[3]);
            mma_bf16_16816(s0, al[0], al[1], al[2], al[3], bh4[0], bh4[1]);
            mma_bf16_16816(s1, al[0], al[1], al[2], al[3], bh4[2], bh4[3]);
#pragma unroll
            for (int half = 0; half < 2; ++half) {       // table rows 16 * half .. 16 * half + 15
                const int er = brow + 16 * half;
                const uint32_t off = er * 128 + ((bc ^ (er & 7)) << 4);
                ldsm_x4(aEh + off, bh4[0], bh4[1], bh4[2], bh4[3]);
                ldsm_x4(aEl + off, bl4[0], bl4[1], bl4[2], bl4[3]);
                mma_bf16_16816(r[2 * half], ah[0], ah[1], ah[2], ah[3], bh4[0], bh4[1]);
                mma_bf16_16816(r[2 * half + 1], ah[0], ah[1], ah[2], ah[3], bh4[2], bh4[3]);
                mma_bf16_16816(r[2 * half], ah[0], ah[1], ah[2], ah[3], bl4[0], bl4[1]);
                mma_bf16_16816(r[2 * half + 1], ah[0], ah[1], ah[2], ah[3], bl4[2], bl4[3]);
                mma_bf16_16816(r[2 * half], al[0], al[1], al[2], al[3], bh4[0], bh4[1]);
                mma_bf16_16816(r[2 * half + 1], al[0], al[1], al[2], al[3], bh4[2], bh4[3]);
            }
        }
        // R -> scratch [16][33]; logit[i][j] += R[i][j - i + 15]
        const int i_lo = lane >> 2, i_hi = i_lo + 8, jc = (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            sR[i_lo * 33 + nt * 8 + jc] = r[nt][0]; sR[i_lo * 33 + nt * 8 + jc + 1] = r[nt][1];
            sR[i_hi * 33 + nt * 8 + jc] = r[nt][2]; sR[i_hi * 33 + nt * 8 + jc + 1] = r[nt][3];
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            s0[e] += sR[i_lo * 33 + (jc + e) - i_lo + 15];          s1[e] += sR[i_lo * 33 + (8 + jc + e) - i_lo + 15];
            s0[2 + e] += sR[i_hi * 33 + (jc + e) - i_hi + 15];      s1[2 + e] += sR[i_hi * 33 + (8 + jc + e) - i_hi + 15];
        }
        // exact softmax over the 16 keys of rows i_lo (elements 0, 1) and i_hi (2, 3); a row lives in one quad
        float m_lo = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])), m_hi = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
        s0[0] = __expf(s0[0] - m_lo); s0[1] = __expf(s0[1] - m_lo); s1[0] = __expf(s1[0] - m_lo); s1[1] = __expf(s1[1] - m_lo);
        s0[2] = __expf(s0[2] - m_hi); s0[3] = __expf(s0[3] - m_hi); s1[2] = __expf(s1[2] - m_hi); s1[3] = __expf(s1[3] - m_hi);
        float l_lo = s0[0] + s0[1] + s1[0] + s1[1], l_hi = s0[2] + s0[3] + s1[2] + s1[3];
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
        // P = hi + lo as the A operand of the second contraction (accumulator fragments are already in A layout)
        uint32_t pah[4], pal[4];
        {
            const float pv[8] = {s0[0], s0[1], s0[2], s0[3], s1[0], s1[1], s1[2], s1[3]};
            float lo8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) lo8[i] = pv[i] - __bfloat162float(__float2bfloat16_rn(pv[i]));
            pah[0] = f2_to_bf2(pv[0], pv[1]); pah[1] = f2_to_bf2(pv[2], pv[3]); pah[2] = f2_to_bf2(pv[4], pv[5]); pah[3] = f2_to_bf2(pv[6], pv[7]);
            pal[0] = f2_to_bf2(lo8[0], lo8[1]); pal[1] = f2_to_bf2(lo8[2], lo8[3]); pal[2] = f2_to_bf2(lo8[4], lo8[5]); pal[3] = f2_to_bf2(lo8[6], lo8[7]);
        }
        float o[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt += 2) {
            uint32_t vh[4], vl[4];
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8, cc = nt + (lane >> 4);
            const uint32_t off = row * 128 + ((cc ^ (row & 7)) << 4);
            ldsm_x4_trans(aVh + off, vh[0], vh[1], vh[2], vh[3]);
            ldsm_x4_trans(aVl + off, vl[0], vl[1], vl[2], vl[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) { o[nt][j] = 0.f; o[nt + 1][j] = 0.f; }
            mma_bf16_16816(o[nt], pah[0], pah[1], pah[2], pah[3], vh[0], vh[1]);
            mma_bf16_16816(o[nt + 1], pah[0], pah[1], pah[2], pah[3], vh[2], vh[3]);
            mma_bf16_16816(o[nt], pah[0], pah[1], pah[2], pah[3], vl[0], vl[1]);
            mma_bf16_16816(o[nt + 1], pah[0], pah[1], pah[2], pah[3], vl[2], vl[3]);
            mma_bf16_16816(o[nt], pal[0], pal[1], pal[2], pal[3], vh[0], vh[1]);
            mma_bf16_16816(o[nt + 1], pal[0], pal[1], pal[2], pal[3], vh[2], vh[3]);
        }
        const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
        __syncwarp();                                     // everyone is done with the Q tiles: the hi tile becomes the output tile
        // out_tile[frame = d % 16][channel = row * 4 + d / 16] (bf16, 128 bytes per frame)
        bf16* st = reinterpret_cast<bf16*>(sQh);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int f = (nt & 1) * 8 + jc + e, qd = nt >> 1;
                st[f * 64 + i_lo * 4 + qd] = __float2bfloat16_rn(o[nt][e] * inv_lo);
                st[f * 64 + i_hi * 4 + qd] = __float2bfloat16_rn(o[nt][2 + e] * inv_hi);
            }
        }
        __syncwarp();
        bf16* ob = p.o + b * p.o_bs + g * p.o_gs + hh * D;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, f = idx >> 3, ch = idx & 7;
            *reinterpret_cast<uint4*>(ob + (long long)f * p.o_rs + ch * 8) = *reinterpret_cast<const uint4*>(st + f * 64 + ch * 8);
        }
    }
}

__global__ void __launch_bounds__(ROWS)
attention_kernel(const AttnParams p) {
    pdl_prologue();
    extern __shared__ float smem[];             // [pairs][2][KC][D]
    const int pairs = p.Tq >= ROWS ? 1 : ROWS / p.Tq;
    const int slabs = p.Tq >= ROWS ? p.Tq / ROWS : 1;
    const int tid = threadIdx.x;
    int pair_local, row;
    long long bh;
    if (p.Tq >= ROWS) {
        bh = blockIdx.x / slabs;
        row = (blockIdx.x % slabs) * ROWS + tid;
        pair_local = 0;
    } else {
        pair_local = tid / p.Tq;
        row = tid % p.Tq;
        bh = (long long)blockIdx.x * pairs + pair_local;
    }
    const long long nbh = (long long)p.B * p.H;
    const bool active = bh < nbh;
    const int b = active ? (int)(bh / p.H) : 0, h = active ? (int)(bh % p.H) : 0;

    float q[D], o[D];
    if (active) {
        const bf16* qp = p.q + b * p.q_bs + h * p.q_hs + (long long)row * p.q_rs;
#pragma unroll
        for (int d = 0; d < D; d += 8) {
            float f[8];
            unpack8(*reinterpret_cast<const bf16x8*>(qp + d), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[d + i] = f[i] * p.scale;
        }
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = 0.f;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
    float mx = -INFINITY, l = 0.f;

    const long long bh0 = p.Tq >= ROWS ? bh : (long long)blockIdx.x * pairs;
    float* sK = smem + pair_local * 2 * KC * D;
    float* sV = sK + KC * D;
    for (int k0 = 0; k0 < p.Tk; k0 += KC) {
        const int kn = min(KC, p.Tk - k0);
        __syncthreads();
        // cooperative load: pairs * kn keys * 8 vectors, for K and V
        for (int i = tid; i < pairs * KC * 8; i += ROWS) {
            const int pr = i / (KC * 8), rem = i % (KC * 8), kk = rem / 8, vv = rem % 8;
            const long long bhl = bh0 + pr;
            if (kk < kn && bhl < nbh) {
                const int bb = (int)(bhl / p.H), hh = (int)(bhl % p.H);
                float f[8];
                unpack8(*reinterpret_cast<const bf16x8*>(p.k + bb * p.k_bs + hh * p.k_hs + (long long)(k0 + kk) * p.k_rs + vv * 8), f);
                float* dst = smem + (pr * 2 * KC + kk) * D + vv * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = f[j];
                unpack8(*reinterpret_cast<const bf16x8*>(p.v + bb * p.v_bs + hh * p.v_hs + (long long)(k0 + kk) * p.v_rs + vv * 8), f);
                dst += KC * D;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = f[j];
            }
        }
        __syncthreads();
        float s[KC];
        float cmax = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float acc = 0.f;
            if (kk < kn) {
                acc = dot64(q, reinterpret_cast<const float4*>(sK + kk * D));
                if (p.relk)
                    acc += dot64_ldg(q, reinterpret_cast<const float4*>(
                        p.relk + ((long long)(h % p.hpg) * (2 * p.Tk - 1) + (k0 + kk - row + p.Tk - 1)) * D));
            } else {
                acc = -INFINITY;
            }
            s[kk] = acc;
            cmax = fmaxf(cmax, acc);
        }
        const float nmx = fmaxf(mx, cmax);
        const float corr = __expf(mx - nmx);
        l *= corr;
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] *= corr;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            if (kk < kn) {
                const float pw = __expf(s[kk] - nmx);
                l += pw;
                const float4* vr = reinterpret_cast<const float4*>(sV + kk * D);
#pragma unroll
                for (int d4 = 0; d4 < D / 4; ++d4) {
                    const float4 t = vr[d4];
                    o[4 * d4] = fmaf(pw, t.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(pw, t.y, o[4 * d4 + 1]);
                    o[4 * d4 + 2] = fmaf(pw, t.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(pw, t.w, o[4 * d4 + 3]);
                }
            }
        }
        mx = nmx;
    }
    if (active) store_row(p, b, h, row, o, 1.0f / l);
}


// ---- Tq <= 16, Tk <= 16 * KT (cross-attention over a short context, e.g. PixArt: 16 latent tokens x 77 text tokens; the
// text-conditioned UNet: 16 or 4 pixels x (77 text + own) keys): one warp per (batch, head), same mma.sync structure as
// above with KT key blocks; keys >= Tk are masked to -inf, query rows >= Tq are zero and never stored.  Longer query
// sequences (the 8x8 level of the video UNet: 64 x 64) are cut into blocks of 16 rows, one warp each (the keys / values
// of a head are staged once per block: 16 KB from L2 per 2 x 16 x 64 x 64 MACs).
// Replaces the generic SIMT kernel, which took 131 us per launch on the PixArt workload (46 % of its step).
template <int KT>
__global__ void __launch_bounds__(KT > 5 ? 32 : 64) attention16xn_mma_kernel(const AttnParams p) {
    pdl_prologue();
    constexpr int W = KT > 5 ? 1 : 2;                      // warps per CTA (48 KB of static shared memory)
    __shared__ __align__(128) uint8_t sm[W][(1 + 2 * KT) * 2048];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqb = (p.Tq + 15) >> 4;                      // query blocks of 16 rows per (batch, head)
    const long long item = (long long)blockIdx.x * W + warp;
    if (item >= (long long)p.B * p.H * nqb) return;
    const long long bh = item / nqb;
    const int q0 = (int)(item - bh * nqb) * 16;
    const int tq = min(16, p.Tq - q0);                     // live query rows of this block
    const int b = (int)(bh / p.H), h = (int)(bh % p.H);
    uint8_t* sQ = sm[warp];
    uint8_t* sK = sQ + 2048;
    uint8_t* sV = sK + KT * 2048;
    const bf16* qp = p.q + b * p.q_bs + h * p.q_hs + (long long)q0 * p.q_rs;
    const bf16* kp = p.k + b * p.k_bs + h * p.k_hs;
    const bf16* vp = p.v + b * p.v_bs + h * p.v_hs;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    {   // Q and K: 4 + 4 KT chunks of 16 B per lane, all in flight; rows >= Tk are zero
        uint4 rq[4], rk[4 * KT];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            rq[i] = row < tq ? *reinterpret_cast<const uint4*>(qp + (long long)row * p.q_rs + c * 8) : zero;
        }
#pragma unroll
        for (int i = 0; i < 4 * KT; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            rk[i] = row < p.Tk ? *reinterpret_cast<const uint4*>(kp + (long long)row * p.k_rs + c * 8) : zero;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            *reinterpret_cast<uint4*>(sQ + row * 128 + ((c ^ (row & 7)) << 4)) = rq[i];
        }
#pragma unroll
        for (int i = 0; i < 4 * KT; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            *reinterpret_cast<uint4*>(sK + row * 128 + ((c ^ (row & 7)) << 4)) = rk[i];
        }
    }
    {
        uint4 rv[4 * KT];
#pragma unroll
        for (int i = 0; i < 4 * KT; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            rv[i] = row < p.Tk ? *reinterpret_cast<const uint4*>(vp + (long long)row * p.v_rs + c * 8) : zero;
        }
#pragma unroll
        for (int i = 0; i < 4 * KT; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, c = idx & 7;
            *reinterpret_cast<uint4*>(sV + row * 128 + ((c ^ (row & 7)) << 4)) = rv[i];
        }
    }
    __syncwarp();
    const uint32_t aQ = (uint32_t)__cvta_generic_to_shared(sQ), aK = (uint32_t)__cvta_generic_to_shared(sK),
                   aV = (uint32_t)__cvta_generic_to_shared(sV);
    // S = Q K^T: per key block two 8-key n-tiles, four 16-wide k-steps over d
    float s[KT][2][4];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[kt][t][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a0, a1, a2, a3;
        {
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8, c = 2 * ks + (lane >> 4);
            ldsm_x4(aQ + row * 128 + ((c ^ (row & 7)) << 4), a0, a1, a2, a3);
        }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            uint32_t b0, b1, b2, b3;
            const int row = kt * 16 + (lane & 7) + (lane >> 4) * 8, c = 2 * ks + ((lane >> 3) & 1);
            ldsm_x4(aK + row * 128 + ((c ^ (row & 7)) << 4), b0, b1, b2, b3);
            mma_bf16_16816(s[kt][0], a0, a1, a2, a3, b0, b1);
            mma_bf16_16816(s[kt][1], a0, a1, a2, a3, b2, b3);
        }
    }
    // mask keys >= Tk, softmax over rows r = lane/4 (elements 0, 1) and r + 8 (elements 2, 3); a row lives in one quad
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = kt * 16 + t * 8 + (lane & 3) * 2 + (j & 1);
                if (key >= p.Tk) s[kt][t][j] = -INFINITY;
                if (j < 2) m_lo = fmaxf(m_lo, s[kt][t][j]); else m_hi = fmaxf(m_hi, s[kt][t][j]);
            }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    const float c = p.scale * 1.4426950408889634f;
    float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            s[kt][t][0] = exp2f((s[kt][t][0] - m_lo) * c); s[kt][t][1] = exp2f((s[kt][t][1] - m_lo) * c);
            s[kt][t][2] = exp2f((s[kt][t][2] - m_hi) * c); s[kt][t][3] = exp2f((s[kt][t][3] - m_hi) * c);
            l_lo += s[kt][t][0] + s[kt][t][1];
            l_hi += s[kt][t][2] + s[kt][t][3];
        }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    // O = P V: KT k-steps of 16 keys, eight 8-wide n-tiles over d; V^T fragments via ldmatrix.trans
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[nt][j] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        const uint32_t pa0 = f2_to_bf2(s[kt][0][0], s[kt][0][1]), pa1 = f2_to_bf2(s[kt][0][2], s[kt][0][3]);
        const uint32_t pa2 = f2_to_bf2(s[kt][1][0], s[kt][1][1]), pa3 = f2_to_bf2(s[kt][1][2], s[kt][1][3]);
#pragma unroll
        for (int nt = 0; nt < 8; nt += 2) {
            uint32_t b0, b1, b2, b3;
            const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, cc = nt + (lane >> 4);
            ldsm_x4_trans(aV + row * 128 + ((cc ^ (row & 7)) << 4), b0, b1, b2, b3);
            mma_bf16_16816(o[nt], pa0, pa1, pa2, pa3, b0, b1);
            mma_bf16_16816(o[nt + 1], pa0, pa1, pa2, pa3, b2, b3);
        }
    }
    const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
    __syncwarp();                                         // everyone is done reading sQ: reuse it for O
    {
        const int r = lane >> 2, q4 = (lane & 3) * 4;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            *reinterpret_cast<uint32_t*>(sQ + r * 128 + ((nt ^ (r & 7)) << 4) + q4) = f2_to_bf2(o[nt][0] * i_lo, o[nt][1] * i_lo);
            *reinterpret_cast<uint32_t*>(sQ + (r + 8) * 128 + ((nt ^ ((r + 8) & 7)) << 4) + q4) =
                f2_to_bf2(o[nt][2] * i_hi, o[nt][3] * i_hi);
        }
    }
    __syncwarp();
    bf16* op = p.o + b * p.o_bs + h * p.o_hs + (long long)q0 * p.o_rs;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i, row = idx >> 3, cch = idx & 7;
        if (row < tq)
            *reinterpret_cast<uint4*>(op + (long long)row * p.o_rs + cch * 8) =
                *reinterpret_cast<const uint4*>(sQ + row * 128 + ((cch ^ (row & 7)) << 4));
    }
}

// ---- wide heads: head dim = 64 * DCH (the single-head attention of the EDM DDPM++ network: d = C = 256, T = 256 or 64).
// One CTA = 64 query rows of one (batch, head), four warps of 16 rows; keys / values stream through shared memory in blocks
// of 64 with the online-softmax recurrence; S = Q K^T and O += P V on mma.sync m16n8k16 (the S accumulators are re-used as
// the A operand of P V, as in the kernels above).  Shared-memory tiles: [DCH][64 rows][64 columns] bf16, 128-byte rows with
// 16-byte chunk c of row r at c ^ (r & 7) (conflict-free ldmatrix).
template <int DCH>
__global__ void __launch_bounds__(128) attention_wide_kernel(const AttnParams p) {
    pdl_prologue();
    extern __shared__ __align__(128) uint8_t smw[];
    constexpr int TILE = 64 * 128;                        // bytes of one [64 x 64] tile
    uint8_t* sQ = smw;
    uint8_t* sK = sQ + DCH * TILE;
    uint8_t* sV = sK + DCH * TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqb = (p.Tq + 63) >> 6;
    const long long bh = blockIdx.x / nqb;
    const int q0 = (int)(blockIdx.x - bh * nqb) * 64;
    const int b = (int)(bh / p.H), h = (int)(bh % p.H);
    const bf16* qp = p.q + b * p.q_bs + h * p.q_hs;
    const bf16* kp = p.k + b * p.k_bs + h * p.k_hs;
    const bf16* vp = p.v + b * p.v_bs + h * p.v_hs;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    // [64 rows x 64 * DCH columns] from global rows r0.. (rows >= limit are zero) into the tiled layout
    auto stage = [&](uint8_t* dst, const bf16* src, long long rs, int r0, int limit) {
        for (int i = threadIdx.x; i < 64 * 8 * DCH; i += 128) {
            const int row = i / (8 * DCH), cc = i % (8 * DCH), t = cc >> 3, c = cc & 7;
            const uint4 v = r0 + row < limit ? *reinterpret_cast<const uint4*>(src + (long long)(r0 + row) * rs + cc * 8) : zero;
            *reinterpret_cast<uint4*>(dst + t * TILE + row * 128 + ((c ^ (row & 7)) << 4)) = v;
        }
    };
    stage(sQ, qp, p.q_rs, q0, p.Tq);
    const uint32_t aQ = (uint32_t)__cvta_generic_to_shared(sQ) + warp * 16 * 128, aK = (uint32_t)__cvta_generic_to_shared(sK),
                   aV = (uint32_t)__cvta_generic_to_shared(sV);
    float o[8 * DCH][4];
#pragma unroll
    for (int nt = 0; nt < 8 * DCH; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[nt][j] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    const float c = p.scale * 1.4426950408889634f;
    for (int k0 = 0; k0 < p.Tk; k0 += 64) {
        __syncthreads();                                  // everyone is done with the previous key / value block
        stage(sK, kp, p.k_rs, k0, p.Tk);
        stage(sV, vp, p.v_rs, k0, p.Tk);
        __syncthreads();
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
        for (int t = 0; t < DCH; ++t) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t a0, a1, a2, a3;
                {
                    const int row = (lane & 7) + ((lane >> 3) & 1) * 8, cc = 2 * ks + (lane >> 4);
                    ldsm_x4(aQ + t * TILE + row * 128 + ((cc ^ (row & 7)) << 4), a0, a1, a2, a3);
                }
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) {          // 16 keys = two 8-key n-tiles per ldmatrix.x4
                    uint32_t b0, b1, b2, b3;
                    const int row = kt * 16 + (lane & 7) + (lane >> 4) * 8, cc = 2 * ks + ((lane >> 3) & 1);
                    ldsm_x4(aK + t * TILE + row * 128 + ((cc ^ (row & 7)) << 4), b0, b1, b2, b3);
                    mma_bf16_16816(s[2 * kt], a0, a1, a2, a3, b0, b1);
                    mma_bf16_16816(s[2 * kt + 1], a0, a1, a2, a3, b2, b3);
                }
            }
        }
        // mask keys >= Tk; block maxima of rows r = lane / 4 (elements 0, 1) and r + 8 (elements 2, 3)
        float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = k0 + nt * 8 + (lane & 3) * 2 + (j & 1);
                if (key >= p.Tk) s[nt][j] = -INFINITY;
                if (j < 2) bm_lo = fmaxf(bm_lo, s[nt][j]); else bm_hi = fmaxf(bm_hi, s[nt][j]);
            }
        bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 1)); bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 2));
        bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 1)); bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 2));
        const float n_lo = fmaxf(m_lo, bm_lo), n_hi = fmaxf(m_hi, bm_hi);
        const float f_lo = exp2f((m_lo - n_lo) * c), f_hi = exp2f((m_hi - n_hi) * c);      // (exp2(-inf) = 0 on the first block)
        m_lo = n_lo; m_hi = n_hi;
        float r_lo = 0.f, r_hi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f((s[nt][0] - n_lo) * c); s[nt][1] = exp2f((s[nt][1] - n_lo) * c);
            s[nt][2] = exp2f((s[nt][2] - n_hi) * c); s[nt][3] = exp2f((s[nt][3] - n_hi) * c);
            r_lo += s[nt][0] + s[nt][1];
            r_hi += s[nt][2] + s[nt][3];
        }
        l_lo = l_lo * f_lo + r_lo;                        // (per-lane partial sums; reduced over the quad at the end)
        l_hi = l_hi * f_hi + r_hi;
#pragma unroll
        for (int nt = 0; nt < 8 * DCH; ++nt) { o[nt][0] *= f_lo; o[nt][1] *= f_lo; o[nt][2] *= f_hi; o[nt][3] *= f_hi; }
        // O += P V: four k-steps of 16 keys; V^T fragments via ldmatrix.trans, two 8-wide n-tiles of d per instruction
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const uint32_t pa0 = f2_to_bf2(s[2 * kt][0], s[2 * kt][1]), pa1 = f2_to_bf2(s[2 * kt][2], s[2 * kt][3]);
            const uint32_t pa2 = f2_to_bf2(s[2 * kt + 1][0], s[2 * kt + 1][1]), pa3 = f2_to_bf2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
            for (int t = 0; t < DCH; ++t) {
#pragma unroll
                for (int nt = 0; nt < 8; nt += 2) {
                    uint32_t b0, b1, b2, b3;
                    const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, cc = nt + (lane >> 4);
                    ldsm_x4_trans(aV + t * TILE + row * 128 + ((cc ^ (row & 7)) << 4), b0, b1, b2, b3);
                    mma_bf16_16816(o[t * 8 + nt], pa0, pa1, pa2, pa3, b0, b1);
                    mma_bf16_16816(o[t * 8 + nt + 1], pa0, pa1, pa2, pa3, b2, b3);
                }
            }
        }
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
    const int r_lo_g = q0 + warp * 16 + (lane >> 2), r_hi_g = r_lo_g + 8;
    bf16* op = p.o + b * p.o_bs + h * p.o_hs;
#pragma unroll
    for (int nt = 0; nt < 8 * DCH; ++nt) {
        const int col = nt * 8 + (lane & 3) * 2;
        if (r_lo_g < p.Tq) *reinterpret_cast<uint32_t*>(op + (long long)r_lo_g * p.o_rs + col) = f2_to_bf2(o[nt][0] * i_lo, o[nt][1] * i_lo);
        if (r_hi_g < p.Tq) *reinterpret_cast<uint32_t*>(op + (long long)r_hi_g * p.o_rs + col) = f2_to_bf2(o[nt][2] * i_hi, o[nt][3] * i_hi);
    }
}

}  // namespace

int xd_attention_tc256_try(const void* q, long long q_bs, long long q_hs, long long q_rs, const void* k,
                           long long k_bs, long long k_hs, long long k_rs, const void* v, long long v_bs,
                           long long v_hs, long long v_rs, void* o, long long o_bs, long long o_hs, long long o_rs,
                           int B, int H, float scale, cudaStream_t st);

extern "C" int xd_attention_bf16(const void* q, long long q_bs, long long q_hs, long long q_rs, const void* k,
                                 long long k_bs, long long k_hs, long long k_rs, const void* v, long long v_bs,
                                 long long v_hs, long long v_rs, void* o, long long o_bs, long long o_hs,
                                 long long o_rs, int B, int H, int Tq, int Tk, int head_dim, float scale,
                                 const float* relk, int scramble, long long o_cs, int qkv_dtype, int heads_per_group,
                                 long long o_gs, void* stream) {
    XD_CHECK_ARG(q && k && v && o && (head_dim == D || head_dim == 4 * D) && B > 0 && H > 0 && Tq > 0 && Tk > 0);
    const int in_f32 = qkv_dtype == XD_F32;
    XD_CHECK_ARG(head_dim == D || (!in_f32 && !relk && !scramble));
    XD_CHECK_ARG(!in_f32 || (Tq == 16 && Tk == 16));            // fp32 q/k/v: SIMT T = 16 kernel only
    XD_CHECK_ARG(q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && q_hs % 8 == 0 && k_hs % 8 == 0 && v_hs % 8 == 0 &&
                 q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0);
    XD_CHECK_ARG(scramble || (o_rs % 8 == 0 && o_hs % 8 == 0 && o_bs % 8 == 0));
    XD_CHECK_ARG(!relk || Tq == Tk);
    if (heads_per_group <= 0) heads_per_group = H;
    XD_CHECK_ARG(H % heads_per_group == 0 && (heads_per_group == H || relk || scramble));
    AttnParams p{(const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, q_bs, q_hs, q_rs, k_bs, k_hs, k_rs,
                 v_bs, v_hs, v_rs, o_bs, o_hs, o_rs, B, H, Tq, Tk, scale, relk, scramble, o_cs, in_f32, heads_per_group, o_gs};
    if (head_dim == 4 * D) {                                      // wide single head (EDM DDPM++ network), mma.sync flash kernel
        XD_CHECK_ARG(o_rs % 2 == 0 && o_hs % 2 == 0 && o_bs % 2 == 0);
        const size_t smem = 3 * 4 * 64 * 128;
        static bool configured_w = false;
        if (!configured_w) {
            if (cudaFuncSetAttribute(attention_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute failed (attention_wide)");
                return XD_ERR_CUDA;
            }
            configured_w = true;
        }
        const long long blocks = (long long)B * H * ((Tq + 63) / 64);
        if (xd_launch(attention_wide_kernel<4>, (unsigned)blocks, 128, smem, (cudaStream_t)stream, p) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
            return XD_ERR_CUDA;
        }
        return XD_OK;
    }
    if (Tq == 256 && Tk == 256 && !relk && !scramble) {           // tcgen05 path (csrc/attention_tc.cu)
        const int rc = xd_attention_tc256_try(q, q_bs, q_hs, q_rs, k, k_bs, k_hs, k_rs, v, v_bs, v_hs, v_rs, o, o_bs,
                                              o_hs, o_rs, B, H, scale, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    if (Tq == T16 && Tk == T16 && !relk && !scramble && !in_f32) { // warp-level tensor-core path
        const long long nb = ((long long)B * H + W16 - 1) / W16;
        if (xd_launch(attention16_mma_kernel, (unsigned)nb, 32 * W16, 0, (cudaStream_t)stream, p) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
            return XD_ERR_CUDA;
        }
        return XD_OK;
    }
    if (Tq <= 128 && Tk <= 144 && !relk && !scramble && !in_f32) {  // short-context (cross-)attention (PixArt 16 x 77,
        const long long nbh = (long long)B * H * ((Tq + 15) / 16);  // video 64 x 64): one warp per 16 query rows
        cudaError_t e;
        if (Tk <= 32) e = xd_launch(attention16xn_mma_kernel<2>, (unsigned)((nbh + 1) / 2), 64, 0, (cudaStream_t)stream, p);
        else if (Tk <= 80) e = xd_launch(attention16xn_mma_kernel<5>, (unsigned)((nbh + 1) / 2), 64, 0, (cudaStream_t)stream, p);
        else if (Tk <= 128) e = xd_launch(attention16xn_mma_kernel<8>, (unsigned)nbh, 32, 0, (cudaStream_t)stream, p);
        else e = xd_launch(attention16xn_mma_kernel<9>, (unsigned)nbh, 32, 0, (cudaStream_t)stream, p);   // 64 + 77 keys (SR stage)
        if (e != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
            return XD_ERR_CUDA;
        }
        return XD_OK;
    }
    if (Tq == T16 && Tk == T16 && relk && scramble && in_f32 && o_cs == 1 && heads_per_group <= 8 && o_rs % 8 == 0 &&
        o_bs % 8 == 0 && o_gs % 8 == 0 && q_rs % 4 == 0 && q_hs % 4 == 0 && q_bs % 4 == 0 && getenv("XDB200_RELPOS_SIMT") == nullptr) {
        // temporal attention of the video UNet: split-precision mma.sync kernel, one warp per (clip, pixel, head)
        static bool configured_rp = false;
        const size_t smem = (size_t)WRP * RP_WARP_BYTES + (size_t)heads_per_group * 8192;
        if (!configured_rp) {
            if (cudaFuncSetAttribute(attention16_relpos_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)WRP * RP_WARP_BYTES + 8 * 8192)) != cudaSuccess) {
                xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute failed (attention16_relpos_mma)");
                return XD_ERR_CUDA;
            }
            configured_rp = true;
        }
        const long long warps = (long long)B * H;
        const long long nb = std::min<long long>((warps + WRP - 1) / WRP, 3 * 148);
        xd_launch(attention16_relpos_mma_kernel, (unsigned)nb, 32 * WRP, smem, (cudaStream_t)stream, p);
        XD_CHECK_LAUNCH();
        return XD_OK;
    }
    if (Tq == T16 && Tk == T16) {
        static bool configured16 = false;
        XD_CHECK_ARG(!relk || heads_per_group <= 8);
        const size_t smem16 = ((size_t)8 * PAIR_LD + (relk ? (size_t)heads_per_group * (2 * T16 - 1) * RS16 : 0)) * sizeof(float);
        if (!configured16) {
            cudaFuncSetAttribute(attention16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(((size_t)8 * PAIR_LD + (size_t)8 * (2 * T16 - 1) * RS16) * sizeof(float)));
            configured16 = true;
        }
        // with tables: a CTA stages them once and walks several groups of 8 problems (two CTAs per SM); without: one group
        const long long groups = ((long long)B * H + 7) / 8;
        const long long nb = relk ? std::min<long long>(groups, 2 * 148) : groups;
        xd_launch(attention16_kernel, (unsigned)nb, ROWS, smem16, (cudaStream_t)stream, p);
        XD_CHECK_LAUNCH();
        return XD_OK;
    }
    XD_CHECK_ARG((Tq >= ROWS && Tq % ROWS == 0) || (Tq < ROWS && ROWS % Tq == 0 && ROWS / Tq <= 8));   // generic kernel
    const int pairs = Tq >= ROWS ? 1 : ROWS / Tq;
    const long long nbh = (long long)B * H;
    const long long blocks = Tq >= ROWS ? nbh * (Tq / ROWS) : (nbh + pairs - 1) / pairs;
    const size_t smem = (size_t)pairs * 2 * KC * D * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * KC * D * 4);
        configured = true;
    }
    xd_launch(attention_kernel, (unsigned)blocks, ROWS, smem, (cudaStream_t)stream, p);
    XD_CHECK_LAUNCH();
    return XD_OK;
}
