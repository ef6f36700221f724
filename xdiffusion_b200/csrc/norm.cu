// Bandwidth-bound normalisation kernels.
//   GroupNorm(32 groups, eps) [+ (1+scale)*y+shift] [+ SiLU] over NHWC bf16   (layers/resnet.py:126-128,
//       151-153,193-197; layers/attention.py:64 in the reference)
//   LayerNorm(no affine, eps) * (1+scale) + shift over fp32 token rows -> bf16 (score_networks/dit.py:16-17,46-51)
// Statistics in fp32; 16-byte vector accesses; warp-shuffle / shared-memory reductions.
#include "common.cuh"
#include <stdio.h>
#include "ptx.cuh"

#include <algorithm>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

// ----------------------------------------------------------------------------- GroupNorm

// Deterministic block reduction of per-thread channel sums to per-group (sum, sumsq): fixed summation
// order (no floating-point atomics), so a sample's statistics do not depend on the batch it is in.
// scratch: [256][16] floats.  part[2*g], part[2*g+1] are written (not accumulated).
__device__ __forceinline__ void block_group_sums(const float (&s)[8], const float (&q)[8], bool active, int V, int ppb,
                                                 int C, int G, float* scratch, float* part) {
    float* mine = scratch + threadIdx.x * 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) { mine[k] = active ? s[k] : 0.f; mine[8 + k] = active ? q[k] : 0.f; }
    __syncthreads();
    // stage 1: one thread per channel folds the ppb pixel slots (in place, into slot 0: a thread only touches its
    // own channel's column); stage 2: one thread per group folds its cpg channels.  Serial depth ppb + cpg instead
    // of ppb * cpg (measured 4-11 us of a 20 us kernel with the single-stage loop), same fixed order for every batch.
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float* t = scratch + (c >> 3) * 16 + (c & 7);
        float cs = t[0], cq = t[8];
        for (int po = 1; po < ppb; ++po) {
            cs += t[po * V * 16];
            cq += t[po * V * 16 + 8];
        }
        t[0] = cs;
        t[8] = cq;
    }
    __syncthreads();
    const int cpg = C / G;
    if (threadIdx.x < G) {
        float gs = 0.f, gq = 0.f;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
            const float* t = scratch + (c >> 3) * 16 + (c & 7);
            gs += t[0];
            gq += t[8];
        }
        part[2 * threadIdx.x] = gs;
        part[2 * threadIdx.x + 1] = gq;
    }
    __syncthreads();
}
// Same reduction with a conflict-free shared-memory layout (the [thread][16] layout above makes the 16 scalar stores of a
// warp 16-way bank-conflicted: ncu counted 1.28 M conflicts per launch).  scratch: [16][NT + 1] floats, k-major with one
// pad column; chan: [2][C] floats of channel sums (the caller's coefficient buffer, not yet in use).
__device__ __forceinline__ void block_group_sums_t(const float (&s)[8], const float (&q)[8], bool active, int V, int ppb,
                                                   int C, int G, int NT, float* scratch, float* chan, float* part) {
    const int pitch = NT + 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        scratch[k * pitch + threadIdx.x] = active ? s[k] : 0.f;
        scratch[(8 + k) * pitch + threadIdx.x] = active ? q[k] : 0.f;
    }
    __syncthreads();
    // stage 1: thread u <-> (k = u / V, j = u % V): consecutive threads read consecutive words; fixed order over po
    for (int u = threadIdx.x; u < C; u += blockDim.x) {
        const int k = u / V, j = u - k * V;
        float cs = 0.f, cq = 0.f;
        for (int po = 0; po < ppb; ++po) {
            cs += scratch[k * pitch + po * V + j];
            cq += scratch[(8 + k) * pitch + po * V + j];
        }
        chan[j * 8 + k] = cs;
        chan[C + j * 8 + k] = cq;
    }
    __syncthreads();
    const int cpg = C / G;
    if (threadIdx.x < G) {
        float gs = 0.f, gq = 0.f;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) { gs += chan[c]; gq += chan[C + c]; }
        part[2 * threadIdx.x] = gs;
        part[2 * threadIdx.x + 1] = gq;
    }
    __syncthreads();
}
// SiLU with one MUFU op: y * sigmoid(y) = 0.5 y (1 + tanh(0.5 y)).  The exp + rcp form needs two and made the apply
// pass MUFU-bound (16 MUFU/clk/SM: 3.7 us for a 64 x 32 x 32 x 128 tensor).
__device__ __forceinline__ float silu_fast(float y) {
    const float h = 0.5f * y;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
// stats[sample][group] = (sum, sum of squares) accumulated with atomics from pixel slabs.
// "sample" = P consecutive pixels (rows of C channels, stride ld).
__global__ void __launch_bounds__(256)
gn_stats_kernel(const bf16* __restrict__ x, long long ld, int P, int C, int G, int inner,
                float* __restrict__ stats) {
    pdl_prologue();
    __shared__ float acc[128];
    __shared__ float scratch[256 * 16];
    const int V = C >> 3;                       // 16-byte vectors per pixel
    const int ppb = 256 / V;                    // pixels per block iteration
    const int sample = blockIdx.y;
    const int slabs = gridDim.x;
    const int per = (P + slabs - 1) / slabs;
    const int p0 = blockIdx.x * per, p1 = min(P, p0 + per);
    const int j = threadIdx.x % V, po = threadIdx.x / V;
    float s[8] = {}, q[8] = {};
    if (po < ppb) {
        // row(sample, p) = (sample / inner) * P * inner + sample % inner + p * inner
        const bf16* base = x + ((long long)(sample / inner) * P * inner + sample % inner) * ld + j * 8;
        for (int p = p0 + po; p < p1; p += ppb) {
            float f[8];
            unpack8(*reinterpret_cast<const bf16x8*>(base + (long long)p * inner * ld), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
        }
    }
    block_group_sums(s, q, po < ppb, V, ppb, C, G, scratch, acc);
    if (threadIdx.x < G) {                      // one partial per (sample, slab): summed in fixed order by gn_apply
        float* dst = &stats[(((long long)sample * slabs + blockIdx.x) * G + threadIdx.x) * 2];
        dst[0] = acc[2 * threadIdx.x];
        dst[1] = acc[2 * threadIdx.x + 1];
    }
}

// y = x * A[c] + B[c] with A = rstd*gamma*(1+scale), B = (beta - mean*rstd*gamma)*(1+scale) + shift
__global__ void __launch_bounds__(256)
gn_apply_kernel(const bf16* __restrict__ x, long long ld, int P, int C, int G, const float* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ ss,
                long long ss_ld, int ss_div, float eps, int silu, int inner, int stat_slabs, int split,
                bf16* __restrict__ out, long long ldo) {
    pdl_prologue();
    extern __shared__ float coef[];             // [2][C]
    const int sample = blockIdx.y;
    const int cpg = C / G;
    const float inv_cnt = 1.0f / ((float)P * (float)cpg);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float sum = 0.f, sq = 0.f;
        for (int sl = 0; sl < stat_slabs; ++sl) {       // fixed order: deterministic
            const float* st = &stats[(((long long)sample * stat_slabs + sl) * G + g) * 2];
            sum += st[0];
            sq += st[1];
        }
        const float mean = sum * inv_cnt;
        const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        float a = rstd * gamma[c], b = beta[c] - mean * rstd * gamma[c];
        if (ss) {
            const float* row = ss + (long long)(sample / ss_div) * ss_ld;
            const float sc = 1.0f + row[c], sh = row[C + c];
            a *= sc; b = b * sc + sh;
        }
        coef[c] = a; coef[C + c] = b;
    }
    __syncthreads();
    const int V = C >> 3;
    const int slabs = gridDim.x;
    const int per = (P + slabs - 1) / slabs;
    const int p0 = blockIdx.x * per, p1 = min(P, p0 + per);
    const long long n = (long long)(p1 - p0) * V;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = (int)(i % V);
        const long long p = (long long)(sample / inner) * P * inner + sample % inner + (p0 + i / V) * inner;
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + p * ld + j * 8), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float y = fmaf(f[k], coef[j * 8 + k], coef[C + j * 8 + k]);
            f[k] = silu ? silu_fast(y) : y;
        }
        const bf16x8 hi = pack8(f);
        *reinterpret_cast<bf16x8*>(out + p * ldo + j * 8) = hi;
        if (split) {                            // out row = [hi(C) | lo(C)], hi + lo ~ fp32 value (2^-16 relative)
            float h[8];
            unpack8(hi, h);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] -= h[k];
            *reinterpret_cast<bf16x8*>(out + p * ldo + C + j * 8) = pack8(f);
        }
    }
}


// ----------------------------------------------------------------------------- GroupNorm over (C / 32 x F) per pixel
// The temporal attention of the video UNet normalises every pixel of every clip over its F frames (samples (clip, pixel),
// rows = frames, row stride HW * ld) and feeds a split-precision GEMM (out row = [hi(C) | lo(C)]).  One WARP per sample,
// a lane owns C / 32 consecutive channels = exactly one group, so the statistics never leave the lane: one pass, one
// launch (the generic path is xd_groupnorm_stats + xd_groupnorm_apply: two launches and a second read).
template <int CPL>       // channels per lane: 4 (C = 128) or 8 (C = 256)
__global__ void __launch_bounds__(256)
gn_frames_kernel(const bf16* __restrict__ x, long long ld, int B, int F, int HW, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, bf16* __restrict__ out, long long ldo) {
    pdl_prologue();
    constexpr int C = 32 * CPL, MAXF = 16;
    const int lane = threadIdx.x & 31;
    const long long s = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= (long long)B * HW) return;
    const long long b = s / HW, pix = s - b * HW;
    const bf16* xp = x + (b * F * HW + pix) * ld + lane * CPL;
    bf16* op = out + (b * F * HW + pix) * ldo + lane * CPL;
    float v[MAXF][CPL];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int f = 0; f < MAXF; ++f) {
        if (f < F) {
            if constexpr (CPL == 8) {
                unpack8(*reinterpret_cast<const bf16x8*>(xp + (long long)f * HW * ld), v[f]);
            } else {
                const uint2 t = *reinterpret_cast<const uint2*>(xp + (long long)f * HW * ld);
                const float2 a = bf2_to_f2(t.x), c = bf2_to_f2(t.y);
                v[f][0] = a.x; v[f][1] = a.y; v[f][2] = c.x; v[f][3] = c.y;
            }
        }
    }
#pragma unroll
    for (int f = 0; f < MAXF; ++f) {
        if (f < F) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) { sum += v[f][k]; sq = fmaf(v[f][k], v[f][k], sq); }
        }
    }
    const float inv_cnt = 1.0f / ((float)F * (float)CPL);
    const float mean = sum * inv_cnt;
    const float rstd = rsqrtf(fmaxf(sq * inv_cnt - mean * mean, 0.f) + eps);
    float a[CPL], bb[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        a[k] = rstd * __ldg(gamma + lane * CPL + k);
        bb[k] = __ldg(beta + lane * CPL + k) - mean * a[k];
    }
#pragma unroll
    for (int f = 0; f < MAXF; ++f) {
        if (f < F) {
            float y[CPL], lo[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                y[k] = fmaf(v[f][k], a[k], bb[k]);
                lo[k] = y[k] - __bfloat162float(__float2bfloat16_rn(y[k]));
            }
            bf16* row = op + (long long)f * HW * ldo;
            if constexpr (CPL == 8) {
                *reinterpret_cast<bf16x8*>(row) = pack8(y);
                *reinterpret_cast<bf16x8*>(row + C) = pack8(lo);
            } else {
                *reinterpret_cast<uint2*>(row) = make_uint2(f2_to_bf2(y[0], y[1]), f2_to_bf2(y[2], y[3]));
                *reinterpret_cast<uint2*>(row + C) = make_uint2(f2_to_bf2(lo[0], lo[1]), f2_to_bf2(lo[2], lo[3]));
            }
        }
    }
}

// ----------------------------------------------------------------------------- fused GroupNorm (cluster)
// One thread-block CLUSTER per sample: each CTA keeps its slab of pixels in shared memory, the 32 group
// statistics are reduced across the cluster through distributed shared memory, and the slab is
// normalised straight from shared memory -- one HBM/L2 read and one write per element, one launch
// (the two-kernel path reads x twice and needs a memset + two launches).
template <int NT>
__global__ void __launch_bounds__(NT)
gn_fused_kernel(const bf16* __restrict__ x, long long ld, int P, int C, int G, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ ss, long long ss_ld, int ss_div, float eps,
                int silu, bf16* __restrict__ out, long long ldo, int slab_px) {
    pdl_prologue();
    extern __shared__ __align__(16) uint8_t smem_gn[];
#ifdef XDB200_INSTRUMENT
    long long tph[8]; int nph = 0;
#define GN_MARK() do { tph[nph++] = clock64(); } while (0)
#else
#define GN_MARK() do { } while (0)
#endif
    GN_MARK();
    // "every CTA of the cluster has started" barrier, split: arrive here, wait just before the first DSMEM store
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int sample = blockIdx.y;
    __shared__ float scratch[16 * (NT + 1)];
    float* all_part = reinterpret_cast<float*>(smem_gn);        // [8 ranks][64][2] partial (sum, sumsq) of every CTA
    float* part = all_part + rank * 128;                        // this CTA's row
    float* coef = all_part + 8 * 128;                           // [2][C]
    uint4* slab = reinterpret_cast<uint4*>(coef + 2 * C);       // [slab_px][C/8]
    const int V = C >> 3;
    const int p0 = rank * slab_px, p1 = min(P, p0 + slab_px);
    const int nvec = max(0, p1 - p0) * V;
    // pass 1: global -> smem with bulk async copies (one per pixel row, or 32 KB pieces when the rows are dense):
    // the whole slab is in flight at once, which a register-staged loop of 256 threads cannot do (measured 1.5 TB/s
    // on 32 x 32 x 128 samples, 0.9 TB/s on 32 x 32 x 384).  Then per-thread channel sums from shared memory; a
    // thread always sees the same 8 channels (blockDim is a multiple of V for every supported C, or po >= ppb idles).
    const bf16* base = x + ((long long)sample * P + p0) * ld;
    const int cpg = C / G;
    const int npx = max(0, p1 - p0);
    __shared__ __align__(8) uint64_t slab_bar;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&slab_bar, 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t bar = ptx::smem_u32(&slab_bar), dst = ptx::smem_u32(slab);
        if (threadIdx.x == 0) ptx::mbar_arrive_expect_tx(&slab_bar, (uint32_t)npx * C * 2);
        __syncwarp();
        if (ld == C) {
            const uint32_t bytes = (uint32_t)npx * C * 2, piece = 32768;
            for (uint32_t off = threadIdx.x * piece; off < bytes; off += 32 * piece)
                ptx::bulk_load(dst + off, reinterpret_cast<const char*>(base) + off, min(piece, bytes - off), bar);
        } else {
            for (int px = threadIdx.x; px < npx; px += 32)
                ptx::bulk_load(dst + px * C * 2, base + (long long)px * ld, C * 2, bar);
        }
    }
    ptx::mbar_wait(&slab_bar, 0);
    GN_MARK();
    float s[8] = {}, q[8] = {};
    const int ppb = NT / V;                                     // pixels per block iteration
    const int j = threadIdx.x % V, po = threadIdx.x / V;
    if (po < ppb) {
#pragma unroll 4
        for (int px = po; px < npx; px += ppb) {
            bf16x8 t; t.u = slab[px * V + j];
            float f[8];
            unpack8(t, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
        }
    }
    GN_MARK();
    block_group_sums_t(s, q, po < ppb, V, ppb, C, G, NT, scratch, coef, part);
    GN_MARK();
    // push this CTA's group partials into every peer's table (DSMEM stores), then ONE cluster barrier: afterwards all
    // statistics are local, nobody touches a peer's shared memory any more, and a CTA may exit as soon as it is done
    // (the pull version needed a second cluster barrier before exit).
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 2 * G) {
        const float vme = part[threadIdx.x];
        for (int r = 0; r < CS; ++r)
            if (r != rank) cluster.map_shared_rank(all_part, r)[rank * 128 + threadIdx.x] = vme;
    }
    cluster.sync();
    GN_MARK();
    // per-channel affine coefficients from the cluster-wide statistics (fixed-order sum over the ranks)
    const float inv_cnt = 1.0f / ((float)P * (float)cpg);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float sum = 0.f, sq = 0.f;
        for (int r = 0; r < CS; ++r) { sum += all_part[r * 128 + 2 * g]; sq += all_part[r * 128 + 2 * g + 1]; }
        const float mean = sum * inv_cnt;
        const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        float a = rstd * gamma[c], b = beta[c] - mean * rstd * gamma[c];
        if (ss) {
            const float* row = ss + (long long)(sample / ss_div) * ss_ld;
            const float sc = 1.0f + row[c], sh = row[C + c];
            a *= sc; b = b * sc + sh;
        }
        coef[c] = a; coef[C + c] = b;
    }
    GN_MARK();
    __syncthreads();                                            // coefs visible to the whole CTA
    GN_MARK();
    // pass 2: smem -> normalise -> global.  Same (pixel slot, vector) decomposition as pass 1: the thread's 16
    // coefficients live in registers and the loop has no integer division (the flat-index form spent ~800 cycles
    // per 16-byte vector: i / V, i % V and 16 scalar coefficient loads).
    bf16* obase = out + ((long long)sample * P + p0) * ldo;
    if (po < ppb) {
        float ca[8], cb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { ca[k] = coef[j * 8 + k]; cb[k] = coef[C + j * 8 + k]; }
#pragma unroll 2
        for (int px = po; px < npx; px += ppb) {
            bf16x8 t; t.u = slab[px * V + j];
            float f[8];
            unpack8(t, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float y = fmaf(f[k], ca[k], cb[k]);
                f[k] = silu ? silu_fast(y) : y;
            }
            *reinterpret_cast<bf16x8*>(obase + (long long)px * ldo + j * 8) = pack8(f);
        }
    }
#ifdef XDB200_INSTRUMENT
    GN_MARK();
    if (threadIdx.x == 0 && blockIdx.x == 0 && (blockIdx.y == 0 || blockIdx.y == gridDim.y - 1))
        printf("[gn prof] sample %d P=%d C=%d cs=%d: load %lld  sums %lld  blocksum %lld  csync %lld  coef %lld  csync %lld  apply %lld\n",
               sample, P, C, CS, tph[1] - tph[0], tph[2] - tph[1], tph[3] - tph[2], tph[4] - tph[3], tph[5] - tph[4],
               tph[6] - tph[5], tph[7] - tph[6]);
#endif
#undef GN_MARK
}

// ----------------------------------------------------------------------------- LayerNorm + modulate
// GroupNorm whose statistics come from the PRODUCER: the contraction that wrote x also emitted, per block of 32 rows and per
// quad of 4 channels, the sum and the sum of squares of what it stored (gemm_tc.cu, QS kernels).  What is left is one
// streaming pass: fold the sample's P / 32 partial rows in a fixed order (two stages: NPART contiguous ranges, then the
// ranges), group them, and apply y = x * A[c] + B[c] (+ SiLU) with the coefficients of the thread's 8 channels in registers.
// blockDim.x = the largest multiple of V = C / 8 that is <= 256, so a thread always works on the same vector column.
__global__ void __launch_bounds__(256)
gn_apply_quads_kernel(const bf16* __restrict__ x, long long ld, int P, int C, int G, const float* __restrict__ qstats,
                      long long qld, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ ss, long long ss_ld, int ss_div, float eps, int silu,
                      bf16* __restrict__ out, long long ldo) {
    pdl_prologue();
    extern __shared__ float sm[];               // [2][C] coefficients | [2][nq] quad sums | [NPARTS][2][nq] partial ranges
    const int nq = C >> 2, nb = P >> 5;
    float* coef = sm;
    float* qsum = sm + 2 * C;
    float* part = qsum + 2 * nq;
    const int sample = blockIdx.y;
    const int nparts = min(max((int)blockDim.x / nq, 1), nb);
    const int per = (nb + nparts - 1) / nparts;
    for (int u = threadIdx.x; u < nparts * nq; u += blockDim.x) {
        const int pt = u / nq, q = u - pt * nq;
        const int b0 = pt * per, b1 = min(nb, b0 + per);
        const float2* src = reinterpret_cast<const float2*>(qstats + ((long long)sample * nb + b0) * qld) + q;
        float s = 0.f, sq = 0.f;
        for (int b = b0; b < b1; ++b, src += qld / 2) {
            const float2 v = __ldg(src);
            s += v.x; sq += v.y;
        }
        part[(pt * 2) * nq + q] = s;
        part[(pt * 2 + 1) * nq + q] = sq;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        float s = 0.f, sq = 0.f;
        for (int pt = 0; pt < nparts; ++pt) { s += part[(pt * 2) * nq + q]; sq += part[(pt * 2 + 1) * nq + q]; }
        qsum[q] = s; qsum[nq + q] = sq;
    }
    __syncthreads();
    const int cpg = C / G, qpg = cpg >> 2;
    const float inv_cnt = 1.0f / ((float)P * (float)cpg);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float sum = 0.f, sq = 0.f;
        for (int q = g * qpg; q < (g + 1) * qpg; ++q) { sum += qsum[q]; sq += qsum[nq + q]; }
        const float mean = sum * inv_cnt;
        const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        float a = rstd * gamma[c], b = beta[c] - mean * rstd * gamma[c];
        if (ss) {
            const float* row = ss + (long long)(sample / ss_div) * ss_ld;
            const float sc = 1.0f + row[c], sh = row[C + c];
            a *= sc; b = b * sc + sh;
        }
        coef[c] = a; coef[C + c] = b;
    }
    __syncthreads();
    const int V = C >> 3;
    const int j = threadIdx.x % V, r0 = threadIdx.x / V, rstep = blockDim.x / V;
    float ca[8], cb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ca[k] = coef[j * 8 + k]; cb[k] = coef[C + j * 8 + k]; }
    const int slabs = gridDim.x;
    const int rows = (P + slabs - 1) / slabs;
    const int p0 = blockIdx.x * rows, p1 = min(P, p0 + rows);
    const bf16* xb = x + ((long long)sample * P) * ld + j * 8;
    bf16* ob = out + ((long long)sample * P) * ldo + j * 8;
    for (int r = p0 + r0; r < p1; r += 4 * rstep) {          // four rows in flight per thread
        bf16x8 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r + i * rstep < p1) v[i] = *reinterpret_cast<const bf16x8*>(xb + (long long)(r + i * rstep) * ld);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (r + i * rstep >= p1) break;
            float f[8];
            unpack8(v[i], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float y = fmaf(f[k], ca[k], cb[k]);
                f[k] = silu ? silu_fast(y) : y;
            }
            *reinterpret_cast<bf16x8*>(ob + (long long)(r + i * rstep) * ldo) = pack8(f);
        }
    }
}

template <int NV>   // D = NV * 128
__global__ void __launch_bounds__(256)
ln_modulate_kernel(const float* __restrict__ x, long long ld, int M, const float* __restrict__ shift,
                   const float* __restrict__ scale, long long mod_ld, int rows_per_mod, float eps,
                   bf16* __restrict__ out, long long ldo) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= M) return;
    const float* xr = x + m * ld;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float D = (float)(NV * 128);
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    const long long mr = (m / rows_per_mod) * mod_ld;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int d = (i * 32 + lane) * 4;
        float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
        if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + mr + d));
        if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + mr + d));
        const float y0 = fmaf(v[i].x * rstd, 1.0f + sc.x, sh.x), y1 = fmaf(v[i].y * rstd, 1.0f + sc.y, sh.y);
        const float y2 = fmaf(v[i].z * rstd, 1.0f + sc.z, sh.z), y3 = fmaf(v[i].w * rstd, 1.0f + sc.w, sh.w);
        __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1), hi = __floats2bfloat162_rn(y2, y3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out + m * ldo + d) = pk;
    }
}

}  // namespace

// Number of pixel slabs (= partial statistics per sample) the two-kernel path uses: a pure function of the
// problem shape, shared by xd_groupnorm_stats and xd_groupnorm_apply.
extern "C" int xd_groupnorm_slabs(int nsamples, int P, int C) {
    int slabs = (2 * 148 + nsamples - 1) / nsamples;
    const int ppb = 256 / (C / 8);
    return max(1, min(min(slabs, 32), (P + ppb * 4 - 1) / (ppb * 4)));
}

extern "C" int xd_groupnorm_stats(const void* x, long long ld, int nsamples, int P, int C, int groups, int inner,
                                  float* stats, void* stream) {
    XD_CHECK_ARG(x && stats && C % 8 == 0 && C <= 2048 && groups <= 64 && C % groups == 0 && ld % 8 == 0);
    XD_CHECK_ARG(inner >= 1 && nsamples % inner == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int slabs = xd_groupnorm_slabs(nsamples, P, C);
    xd_launch(gn_stats_kernel, dim3(slabs, nsamples), 256, 0, st, (const bf16*)x, ld, P, C, groups, inner, stats);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_groupnorm_apply(const void* x, long long ld, int nsamples, int P, int C, int groups,
                                  const float* stats, const float* gamma, const float* beta, const float* scale_shift,
                                  long long ss_ld, int ss_div, float eps, int silu, int inner, int split, void* out,
                                  long long ldo, void* stream) {
    XD_CHECK_ARG(x && stats && gamma && beta && out && C % 8 == 0 && C % groups == 0 && ld % 8 == 0 && ldo % 8 == 0);
    XD_CHECK_ARG(inner >= 1 && nsamples % inner == 0);
    int slabs = (4 * 148 + nsamples - 1) / nsamples;
    slabs = max(1, min(slabs, (P + 15) / 16));
    xd_launch(gn_apply_kernel, dim3(slabs, nsamples), 256, 2 * C * sizeof(float), (cudaStream_t)stream, 
        (const bf16*)x, ld, P, C, groups, stats, gamma, beta, scale_shift, ss_ld, ss_div > 0 ? ss_div : 1, eps, silu,
        inner, xd_groupnorm_slabs(nsamples, P, C), split, (bf16*)out, ldo);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

// GroupNorm from the quad statistics of the producing contraction (xd_conv3x3_bf16_tc_qstats / xd_gemm_bf16_tc_qstats):
// qstats[(row / 32) * qstats_ld + (c / 4) * 2 + {0, 1}] for the rows / channels of x; sample s = rows [s * P, (s + 1) * P).
extern "C" int xd_groupnorm_apply_quads(const void* x, long long ld, int nsamples, int P, int C, int groups,
                                        const float* qstats, long long qstats_ld, const float* gamma, const float* beta,
                                        const float* scale_shift, long long ss_ld, int ss_div, float eps, int silu,
                                        void* out, long long ldo, void* stream) {
    XD_CHECK_ARG(x && qstats && gamma && beta && out && nsamples > 0 && P > 0 && P % 32 == 0 && ld % 8 == 0 && ldo % 8 == 0);
    XD_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && groups > 0 && C % groups == 0 && (C / groups) % 4 == 0 && qstats_ld % 2 == 0);
    const int V = C / 8, nq = C / 4, nb = P / 32;
    const int threads = 256 / V * V;
    const int nparts = std::min(std::max(threads / nq, 1), nb);
    const size_t smem = (2 * (size_t)C + 2 * nq + 2 * (size_t)nparts * nq) * sizeof(float);
    XD_CHECK_ARG(smem <= 48 * 1024);
    int slabs = (4 * 148 + nsamples - 1) / nsamples;
    slabs = std::max(1, std::min(slabs, P / 32));
    xd_launch(gn_apply_quads_kernel, dim3(slabs, nsamples), threads, smem, (cudaStream_t)stream, (const bf16*)x, ld, P, C,
              groups, qstats, qstats_ld, gamma, beta, scale_shift, ss_ld, ss_div > 0 ? ss_div : 1, eps, silu, (bf16*)out, ldo);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

// GroupNorm(32 groups) over the F frames of every (clip, pixel) sample, split bf16 output [hi(C) | lo(C)]; rows ordered
// (clip, frame, pixel).  Returns -1 (and does nothing) for shapes the one-warp-per-sample kernel does not cover (C other than
// 128 / 256, F > 16): the caller then uses xd_groupnorm_stats + xd_groupnorm_apply with inner = HW, split = 1.
extern "C" int xd_groupnorm_frames_split(const void* x, long long ld, int B, int F, int HW, int C, const float* gamma,
                                         const float* beta, float eps, void* out, long long ldo, void* stream) {
    XD_CHECK_ARG(x && gamma && beta && out && B > 0 && F > 0 && HW > 0);
    if ((C != 128 && C != 256) || F > 16 || ld % 8 != 0 || ldo % 8 != 0) return -1;
    const unsigned grid = (unsigned)(((long long)B * HW + 7) / 8);
    if (C == 128) xd_launch(gn_frames_kernel<4>, grid, 256, 0, (cudaStream_t)stream, (const bf16*)x, ld, B, F, HW, gamma, beta, eps, (bf16*)out, ldo);
    else xd_launch(gn_frames_kernel<8>, grid, 256, 0, (cudaStream_t)stream, (const bf16*)x, ld, B, F, HW, gamma, beta, eps, (bf16*)out, ldo);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_layernorm_modulate(const float* x, long long ld, int M, int D, const float* shift,
                                     const float* scale, long long mod_ld, int rows_per_mod, float eps, void* out,
                                     long long ldo, void* stream) {
    XD_CHECK_ARG(x && out && M > 0 && D % 128 == 0 && D <= 1024 && ld % 4 == 0 && ldo % 4 == 0 && mod_ld % 4 == 0);
    XD_CHECK_ARG(rows_per_mod > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((M + 7) / 8);
#define XD_LN(NV) xd_launch(ln_modulate_kernel<NV>, grid, 256, 0, st, x, ld, M, shift, scale, mod_ld, rows_per_mod, eps, (bf16*)out, ldo)
    switch (D / 128) {
        case 1: XD_LN(1); break;
        case 2: XD_LN(2); break;
        case 3: XD_LN(3); break;
        case 4: XD_LN(4); break;
        case 6: XD_LN(6); break;
        case 8: XD_LN(8); break;
        default: XD_CHECK_ARG(false && "unsupported D");
    }
#undef XD_LN
    XD_CHECK_LAUNCH();
    return XD_OK;
}

// Returns XD_OK, an error code, or -1 when the sample does not fit the cluster's shared memory
// (the caller then uses xd_groupnorm_stats + xd_groupnorm_apply).
extern "C" int xd_groupnorm_fused(const void* x, long long ld, int nsamples, int P, int C, int groups,
                                  const float* gamma, const float* beta, const float* scale_shift, long long ss_ld,
                                  int ss_div, float eps, int silu, void* out, long long ldo, void* stream) {
    XD_CHECK_ARG(x && gamma && beta && out && C % 8 == 0 && C <= 2048 && groups <= 64 && C % groups == 0);
    XD_CHECK_ARG(ld % 8 == 0 && ldo % 8 == 0 && nsamples > 0 && P > 0);
    const size_t fixed = (8 * 128 + 2 * (size_t)C) * sizeof(float);
    const size_t budget = 110 * 1024;                           // (+ 16 KB static scratch) covers 32x32x384 with 8 CTAs
    // CTAs per sample: enough that a slab is ~16 KB (parallelism: the kernel is latency-bound, 13-17 us at one or two
    // CTAs per sample), at least what fits the budget.  A function of (P, C) only, never of the batch size, so a
    // sample's statistics are summed in the same order whatever batch it is part of.
    int cs = 1;
    while (cs < 8 && (size_t)P * C * 2 > (size_t)cs * 16 * 1024 && P / (2 * cs) >= 4) cs *= 2;
    while (cs <= 8 && (size_t)((P + cs - 1) / cs) * C * 2 + fixed > budget) cs *= 2;
    if (cs > 8) return -1;
    const int slab_px = (P + cs - 1) / cs;
    const size_t smem = fixed + (size_t)slab_px * C * 2;
    // big slabs (one CTA per SM) get 512 threads: twice the issue slots and loads in flight for the same smem
    const bool wide = (size_t)slab_px * C * 2 > 40 * 1024;       // the slab alone (one CTA per SM from ~48 KB)
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gn_fused_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess ||
            cudaFuncSetAttribute(gn_fused_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
            xd_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute failed (gn_fused)");
            return XD_ERR_CUDA;
        }
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs, nsamples);
    cfg.blockDim = dim3(wide ? 512 : 256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = xd_pdl_enabled() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, wide ? gn_fused_kernel<512> : gn_fused_kernel<256>, (const bf16*)x, ld, P, C, groups, gamma,
                           beta, scale_shift, ss_ld, ss_div > 0 ? ss_div : 1, eps, silu, (bf16*)out, ldo,
                           slab_px) != cudaSuccess) {
        xd_set_error(__FILE__, __LINE__, cudaGetErrorString(cudaGetLastError()));
        return XD_ERR_CUDA;
    }
    return XD_OK;
}
