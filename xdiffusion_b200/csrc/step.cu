// One fused kernel per reverse-process step (DDPM ancestral, DDIM, rectified-flow Euler).
//
// The host precomputes, with the same fp32 torch ops the reference uses, one row of eight
// coefficients per loop index i (scheduler tables gathered at i; see xdiffusion_b200/scheduler.py):
//     [0] a  [1] b  [2] c1  [3] c2  [4] sigma  [5] e1  [6] e2  [7] -
// and the kernel evaluates, with explicitly un-fused fp32 arithmetic (__fmul_rn/__fadd_rn: no FMA
// contraction, so results are bit-identical to the reference's elementwise chain):
//   x0   = a*x - b*o            (form 0: discrete eps/v, continuous v; scheduler.py:309-324,536-544)
//        | a*(x - o*b)          (form 1: continuous eps; scheduler.py:524-534)
//   x0   = clamp(x0,-1,1) | dynamic threshold (utils.py:379-396; per-sample quantile by in-smem sort)
//   ancestral: out = i==0 ? x0 : (c1*x0 + c2*x) + sigma*z           (ancestral.py:59-72,189-191)
//   ddim:      e = (pred v) e1*(x - x0_unclipped*e2) | (pred eps) o ;  out = i==0 ? x0 : c1*x0 + c2*e
//   euler:     out = x + o*a                                         (rectified_flow.py:76-84)
// The loop index is read from device memory when a pointer is given, so the whole step replays
// inside a CUDA graph; xd_schedule_advance decrements it and refreshes the per-step network inputs.
#include "common.cuh"

namespace {

enum { MODE_ANCESTRAL = 0, MODE_DDIM = 1, MODE_EULER = 2 };

struct StepParams {
    const float *x, *o, *z;
    float* out;
    const float* coefs;         // [N][8]
    const int* idx_dev;
    int idx_host;
    long long z_step_stride;    // elements between the noise of consecutive loop indices (0: one buffer)
    long long n_total;
    int n_per_sample;
    int mode, form, pred_v;
    int threshold;              // 0: clamp(-1,1), 1: dynamic threshold
    int thr_k;                  // floor(rank)
    float thr_w, thr_c;         // rank - floor(rank), hard cap c
    unsigned long long seed;    // in-kernel Philox noise when z == nullptr
    const unsigned long long* seed_dev;   // if non-null the seed is read from device memory (CUDA-graph replay)
    unsigned long long vec_offset;        // Philox counter of this shard's first float4: (first global row * n_per_sample) / 4,
                                          // so that a batch sharded over ranks draws the noise of the unsharded run
};

// ------------------------------------------------------------------ Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long idx, uint32_t step) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), step, 0x5851F42Du};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    const float u0 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)c[2] + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)c[3] + 0.5f) * 2.3283064365386963e-10f;
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__device__ __forceinline__ float x0_of(const StepParams& p, const float* cf, float x, float o) {
    if (p.form == 0) return __fsub_rn(__fmul_rn(cf[0], x), __fmul_rn(cf[1], o));
    return __fmul_rn(cf[0], __fsub_rn(x, __fmul_rn(o, cf[1])));
}
__device__ __forceinline__ float finish(const StepParams& p, const float* cf, int i, float x, float o, float x0u,
                                        float x0, float z) {
    if (i == 0) return x0;
    if (p.mode == MODE_ANCESTRAL)
        return __fadd_rn(__fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], x)), __fmul_rn(cf[4], z));
    const float e = p.pred_v ? __fmul_rn(cf[5], __fsub_rn(x, __fmul_rn(x0u, cf[6]))) : o;
    return __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], e));
}

// ------------------------------------------------------------------ elementwise variant (clamp / euler)
__global__ void __launch_bounds__(256) step_kernel(const StepParams p) {
    pdl_prologue();
    const int i = p.idx_dev ? *p.idx_dev : p.idx_host;
    const unsigned long long seed = p.seed_dev ? *p.seed_dev : p.seed;
    float cf[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cf[k] = __ldg(p.coefs + (long long)i * 8 + k);
    const float* z = p.z ? p.z + (long long)i * p.z_step_stride : nullptr;
    const long long n4 = p.n_total >> 2;
    const bool need_noise = p.mode == MODE_ANCESTRAL && i != 0;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (long long)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(p.x)[v];
        const float4 o = reinterpret_cast<const float4*>(p.o)[v];
        float4 r;
        if (p.mode == MODE_EULER) {
            r.x = __fadd_rn(x.x, __fmul_rn(o.x, cf[0])); r.y = __fadd_rn(x.y, __fmul_rn(o.y, cf[0]));
            r.z = __fadd_rn(x.z, __fmul_rn(o.z, cf[0])); r.w = __fadd_rn(x.w, __fmul_rn(o.w, cf[0]));
        } else {
            float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
            if (need_noise) zz = z ? reinterpret_cast<const float4*>(z)[v] : philox_normal4(seed, p.vec_offset + (unsigned long long)v, (uint32_t)i);
            const float xs[4] = {x.x, x.y, x.z, x.w}, os[4] = {o.x, o.y, o.z, o.w}, zs[4] = {zz.x, zz.y, zz.z, zz.w};
            float rs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float x0u = x0_of(p, cf, xs[k], os[k]);
                const float x0 = fminf(fmaxf(x0u, -1.0f), 1.0f);
                rs[k] = finish(p, cf, i, xs[k], os[k], x0u, x0, zs[k]);
            }
            r = make_float4(rs[0], rs[1], rs[2], rs[3]);
        }
        reinterpret_cast<float4*>(p.out)[v] = r;
    }
}

// ------------------------------------------------------------------ dynamic-threshold variant
// One CTA per sample.  torch.quantile's 'linear' interpolation needs two order statistics of |x0|: the k-th and
// (k+1)-th smallest.  They are found exactly by an MSB-first radix select on the float bit patterns (non-negative
// floats order like unsigned integers): four passes of a 256-bin shared-memory histogram (integer atomics:
// deterministic) + one pass for the successor -- ~5 block-wide steps instead of the 55 compare-exchange stages of a
// 1024-element bitonic sort (measured 39 us -> see profiles/README.md).
__global__ void __launch_bounds__(256) step_threshold_kernel(const StepParams p, int npow2) {
    pdl_prologue();
    extern __shared__ float sm[];
    float* s_x0 = sm;                // [n_per_sample]
    __shared__ unsigned hist[256];
    __shared__ unsigned sel[4];      // [0] bin, [1] rank inside the bin, [2] count(key <= v_k), [3] min(key > v_k)
    const int i = p.idx_dev ? *p.idx_dev : p.idx_host;
    const unsigned long long seed = p.seed_dev ? *p.seed_dev : p.seed;
    float cf[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cf[k] = __ldg(p.coefs + (long long)i * 8 + k);
    const long long base = (long long)blockIdx.x * p.n_per_sample;
    const int n = p.n_per_sample;
    for (int e = threadIdx.x; e < n; e += blockDim.x) s_x0[e] = x0_of(p, cf, p.x[base + e], p.o[base + e]);
    unsigned prefix = 0, rank = (unsigned)p.thr_k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[threadIdx.x] = 0;                                  // blockDim.x == 256
        __syncthreads();                                        // (also orders the s_x0 writes before the first read)
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const unsigned key = __float_as_uint(fabsf(s_x0[e]));
            if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {                                 // locate the bin that holds the wanted rank
            unsigned c[8], tot = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { c[k] = hist[threadIdx.x * 8 + k]; tot += c[k]; }
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)threadIdx.x >= o) incl += t;
            }
            unsigned before = incl - tot;
            if (rank >= before && rank < incl) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (rank >= before && rank < before + c[k]) { sel[0] = threadIdx.x * 8 + k; sel[1] = rank - before; }
                    before += c[k];
                }
            }
        }
        __syncthreads();
        prefix = (prefix << 8) | sel[0];
        rank = sel[1];
    }
    // prefix is the bit pattern of the k-th smallest; its successor is itself if it has duplicates beyond rank k,
    // else the smallest strictly larger key
    if (threadIdx.x == 0) { sel[2] = 0; sel[3] = 0xffffffffu; }
    __syncthreads();
    {
        unsigned le = 0, mn = 0xffffffffu;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const unsigned key = __float_as_uint(fabsf(s_x0[e]));
            if (key <= prefix) ++le; else mn = min(mn, key);
        }
        atomicAdd(&sel[2], le);
        atomicMin(&sel[3], mn);
    }
    __syncthreads();
    const float lo = __uint_as_float(prefix);
    const float hi = (p.thr_k + 1 >= n || sel[2] >= (unsigned)p.thr_k + 2) ? lo : __uint_as_float(sel[3]);

    // at::lerp: weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w); ATen's vectorised CPU kernel contracts
    // the multiply-add into one FMA (checked against torch.quantile: 100% bit match with fma,
    // 99.9% without), so the fused form is the reference behaviour here.
    const float diff = __fsub_rn(hi, lo);
    float s = p.thr_w < 0.5f ? __fmaf_rn(p.thr_w, diff, lo) : __fmaf_rn(-diff, __fsub_rn(1.0f, p.thr_w), hi);
    s = fminf(fmaxf(s, 1.0f), p.thr_c);
    const float* z = p.z ? p.z + (long long)i * p.z_step_stride : nullptr;
    const bool need_noise = p.mode == MODE_ANCESTRAL && i != 0;
    for (int e4 = threadIdx.x; e4 < n / 4; e4 += blockDim.x) {
        float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need_noise)
            zz = z ? reinterpret_cast<const float4*>(z + base)[e4]
                   : philox_normal4(seed, p.vec_offset + (unsigned long long)(base / 4 + e4), (uint32_t)i);
        const float zs[4] = {zz.x, zz.y, zz.z, zz.w};
        float rs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = e4 * 4 + k;
            const float x0u = s_x0[e];
            const float x0 = __fdiv_rn(fminf(fmaxf(x0u, -s), s), s);
            rs[k] = finish(p, cf, i, p.x[base + e], p.o[base + e], x0u, x0, zs[k]);
        }
        reinterpret_cast<float4*>(p.out + base)[e4] = make_float4(rs[0], rs[1], rs[2], rs[3]);
    }
}

// i <- i - 1 (or set), then refresh the per-step network inputs from host-built tables.
__global__ void advance_kernel(int* idx, int set_to, const long long* tab_i64, const float* tab_f32a,
                               const float* tab_f32b, long long* out_i64, float* out_f32a, float* out_f32b, int B) {
    pdl_prologue();
    __shared__ int s_i;
    if (threadIdx.x == 0) {
        int i = set_to >= 0 ? set_to : (set_to == -1 ? *idx - 1 : *idx);   // -2: refresh outputs only
        if (i < 0) i = 0;
        s_i = i;
    }
    __syncthreads();
    const int i = s_i;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        if (out_i64) out_i64[b] = tab_i64[i];
        if (out_f32a) out_f32a[b] = tab_f32a[i];
        if (out_f32b) out_f32b[b] = tab_f32b[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) *idx = i;
}

// samples = (clamp(x,-1,1)+1)*0.5   (utils.py:62-64)
__global__ void unnormalize_kernel(const float* x, float* out, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fmul_rn(__fadd_rn(fminf(fmaxf(x[i], -1.0f), 1.0f), 1.0f), 0.5f);
}

// ------------------------------------------------------------------ EDM (Karras et al.) Euler / Heun step, fp64 state
// samplers/edm.py:85-137 with the preconditioning of score_networks/edm.py:663-693 folded in.  Every operation is the
// un-fused IEEE operation the reference's tensor expression performs, in its order, so a step is bit-identical to it:
//   D     = c_skip * float(x) + c_out * F                       (fp32: EDMPrecond.forward, D_x)
//   stage 0 (Euler):  d = (x_hat - double(D)) / t_hat;          x_next = x_hat + h * d
//   stage 1 (Heun):   d' = (x_next - double(D')) / t_next;      x_next = x_hat + h * (0.5 * d + 0.5 * d')
//   stage 2:          D only (EDMPrecond.forward as a stand-alone call)
// and, fused behind it, the next network input  xin = c_in_next * float(x_next)  (fp32), so the fp64 state never makes an
// extra trip through HBM between two network evaluations.
struct EdmParams {
    int stage;
    const double *x_hat, *x_mid, *d_in;
    const float* F;
    double *d_out, *x_out;
    float *den_out, *xin_out;
    double t_div, h;
    float c_skip, c_out, c_in_next;
    long long n;
};

__global__ void edm_step_kernel(const EdmParams p) {
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
        const double xs = p.stage == 1 ? p.x_mid[i] : p.x_hat[i];       // the state the network was evaluated at
        const float den = __fadd_rn(__fmul_rn(p.c_skip, (float)xs), __fmul_rn(p.c_out, p.F[i]));
        if (p.stage == 2) { p.den_out[i] = den; continue; }
        const double d = __ddiv_rn(__dsub_rn(xs, (double)den), p.t_div);
        double xn;
        if (p.stage == 0) {
            xn = __dadd_rn(p.x_hat[i], __dmul_rn(p.h, d));
            p.d_out[i] = d;
        } else {
            const double mix = __dadd_rn(__dmul_rn(0.5, p.d_in[i]), __dmul_rn(0.5, d));
            xn = __dadd_rn(p.x_hat[i], __dmul_rn(p.h, mix));
        }
        p.x_out[i] = xn;
        if (p.xin_out) p.xin_out[i] = __fmul_rn(p.c_in_next, (float)xn);
    }
}

// Super-resolution stage input (layers/super_resolution.py:47-121): out[b] = [ x[b] | a * low[b] + c * z[b] ] on the channel
// axis; z = the Gaussian conditioning augmentation, RE-DRAWN at every network evaluation: row `idx` of an injected noise table
// (parity runs), or Philox normals keyed by (seed ^ kSrKey, element, loop index).  Un-fused fp32 multiply / add like q_sample.
struct SrParams {
    const float *x, *low, *z;
    float* out;
    const int* idx_dev;
    int idx_host;
    long long z_step_stride;
    int B, nx, nl;              // per sample: elements of x (Cx * HW) and of low (Cl * HW), multiples of 4
    float a, c;
    unsigned long long seed;
    const unsigned long long* seed_dev;
    unsigned long long vec_offset;
};
constexpr unsigned long long kSrKey = 0x9E3779B97F4A7C15ULL;

__global__ void sr_input_kernel(const SrParams p) {
    pdl_prologue();
    const int per = (p.nx + p.nl) / 4;
    const long long total = (long long)p.B * per;
    const int idx = p.idx_dev ? *p.idx_dev : p.idx_host;
    const unsigned long long seed = (p.seed_dev ? *p.seed_dev : p.seed) ^ kSrKey;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / per;
        const int e = (int)(i - b * per) * 4;
        float4 v;
        if (e < p.nx) {
            v = *reinterpret_cast<const float4*>(p.x + b * p.nx + e);
        } else {
            const long long li = b * p.nl + (e - p.nx);
            const float4 l = *reinterpret_cast<const float4*>(p.low + li);
            const float4 z = p.z ? *reinterpret_cast<const float4*>(p.z + (long long)idx * p.z_step_stride + li)
                                 : philox_normal4(seed, p.vec_offset + (unsigned long long)(li >> 2), (uint32_t)idx);
            v.x = __fadd_rn(__fmul_rn(p.a, l.x), __fmul_rn(p.c, z.x));
            v.y = __fadd_rn(__fmul_rn(p.a, l.y), __fmul_rn(p.c, z.y));
            v.z = __fadd_rn(__fmul_rn(p.a, l.z), __fmul_rn(p.c, z.z));
            v.w = __fadd_rn(__fmul_rn(p.a, l.w), __fmul_rn(p.c, z.w));
        }
        *reinterpret_cast<float4*>(p.out + b * (p.nx + p.nl) + e) = v;
    }
}

// xin = c_in * float(x)                                            (EDMPrecond.forward: (c_in * x).to(dtype))
// x_hat = x + c_noise * z, fp64                                     (samplers/edm.py:117-120, S_churn > 0)
__global__ void edm_in_kernel(const double* x, const double* z, double c_noise, double* x_hat, float c_in, float* xin,
                              long long n) {
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = x[i];
        if (z) {
            v = __dadd_rn(v, __dmul_rn(c_noise, z[i]));
            x_hat[i] = v;
        }
        if (xin) xin[i] = __fmul_rn(c_in, (float)v);
    }
}

}  // namespace

extern "C" int xd_edm_step(int stage, const double* x_hat, const double* x_mid, const double* d_in, const float* F,
                           double* d_out, double* x_out, float* den_out, float* xin_out, double t_div, double h,
                           float c_skip, float c_out, float c_in_next, long long n, void* stream) {
    XD_CHECK_ARG(stage >= 0 && stage <= 2 && x_hat && F && n > 0);
    XD_CHECK_ARG(stage != 0 || (d_out && x_out));
    XD_CHECK_ARG(stage != 1 || (x_mid && d_in && x_out));
    XD_CHECK_ARG(stage != 2 || den_out);
    XD_CHECK_ARG(stage == 2 || t_div != 0.0);
    EdmParams p{stage, x_hat, x_mid, d_in, F, d_out, x_out, den_out, xin_out, t_div, h, c_skip, c_out, c_in_next, n};
    const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 8);
    xd_launch(edm_step_kernel, grid, 256, 0, (cudaStream_t)stream, p);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_sr_input(const float* x, const float* low, const float* z, long long z_step_stride, float* out, int B,
                           int nx, int nl, float a, float c, const int* idx_dev, int idx_host, unsigned long long seed,
                           const unsigned long long* seed_dev, long long elem_offset, void* stream) {
    XD_CHECK_ARG(x && low && out && B > 0 && nx > 0 && nl > 0 && nx % 4 == 0 && nl % 4 == 0 && (idx_dev || idx_host >= 0));
    XD_CHECK_ARG(elem_offset >= 0 && elem_offset % 4 == 0);
    SrParams p{x, low, z, out, idx_dev, idx_host, z_step_stride, B, nx, nl, a, c, seed, seed_dev,
               (unsigned long long)(elem_offset / 4)};
    const long long total = (long long)B * ((nx + nl) / 4);
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 8);
    xd_launch(sr_input_kernel, grid, 256, 0, (cudaStream_t)stream, p);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_edm_prepare(const double* x, const double* z, double c_noise, double* x_hat, float c_in, float* xin,
                              long long n, void* stream) {
    XD_CHECK_ARG(x && n > 0 && (xin || z) && (!z || x_hat));
    const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 8);
    xd_launch(edm_in_kernel, grid, 256, 0, (cudaStream_t)stream, x, z, c_noise, x_hat, c_in, xin, n);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_sampler_step(int mode, int form, int pred_v, const float* x, const float* o, const float* z,
                               long long z_step_stride, float* out, const float* coefs, const int* idx_dev,
                               int idx_host, long long n_total, int n_per_sample, int threshold, int thr_k,
                               float thr_w, float thr_c, unsigned long long seed,
                               const unsigned long long* seed_dev, long long elem_offset, void* stream) {
    XD_CHECK_ARG(x && o && out && coefs && n_total > 0 && n_per_sample > 0 && n_total % n_per_sample == 0);
    XD_CHECK_ARG(elem_offset >= 0 && elem_offset % 4 == 0);
    XD_CHECK_ARG(n_per_sample % 4 == 0 && mode >= 0 && mode <= 2 && (idx_dev || idx_host >= 0));
    StepParams p{x, o, z, out, coefs, idx_dev, idx_host, z_step_stride, n_total, n_per_sample, mode, form, pred_v,
                 threshold, thr_k, thr_w, thr_c, seed, seed_dev, (unsigned long long)(elem_offset / 4)};
    cudaStream_t st = (cudaStream_t)stream;
    if (threshold && mode != MODE_EULER) {
        XD_CHECK_ARG(n_per_sample <= 8192 && thr_k >= 0 && thr_k < n_per_sample);
        int npow2 = 1;
        while (npow2 < n_per_sample) npow2 <<= 1;
        const size_t smem = (size_t)n_per_sample * sizeof(float);
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(step_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 4);
            configured = true;
        }
        xd_launch(step_threshold_kernel, (unsigned)(n_total / n_per_sample), 256, smem, st, p, npow2);
    } else {
        const long long n4 = n_total / 4;
        const unsigned grid = (unsigned)std::min<long long>((n4 + 255) / 256, 148LL * 8);
        xd_launch(step_kernel, grid, 256, 0, st, p);
    }
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_schedule_advance(int* idx_dev, int set_to, const long long* tab_i64, const float* tab_f32a,
                                   const float* tab_f32b, long long* out_i64, float* out_f32a, float* out_f32b,
                                   int B, void* stream) {
    XD_CHECK_ARG(idx_dev && B > 0 && (!out_i64 || tab_i64) && (!out_f32a || tab_f32a) && (!out_f32b || tab_f32b));
    xd_launch(advance_kernel, 1, 256, 0, (cudaStream_t)stream, idx_dev, set_to, tab_i64, tab_f32a, tab_f32b, out_i64,
                                                        out_f32a, out_f32b, B);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_unnormalize(const float* x, float* out, long long n, void* stream) {
    XD_CHECK_ARG(x && out && n > 0);
    xd_launch(unnormalize_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, x, out, n);
    XD_CHECK_LAUNCH();
    return XD_OK;
}
