// CUDA-core kernels: (a) bf16 GEMM / conv3x3 with the same epilogue contract as the tcgen05
// kernels (used to cross-check them on the device and for shapes the tensor-core path does not
// take: K % 64 != 0), (b) the UNet's first conv (C_in tiny, fp32 NCHW in) and last conv
// (C_out tiny, fp32 NCHW out), which are bandwidth-bound and not GEMM-shaped.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace {

// 64x64 output tile, 16-wide k step, 256 threads, 4x4 outputs per thread.
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const bf16* __restrict__ A, long long lda, const bf16* __restrict__ A2, long long lda2, int K2,
                 const bf16* __restrict__ Wt, long long ldw, int M, int N, int K, Epilogue e) {
    pdl_prologue();
    __shared__ float sA[16][65];
    __shared__ float sB[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long m0 = (long long)blockIdx.y * 64;
    const int n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    const int Kt = K + K2;
    for (int k0 = 0; k0 < Kt; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, c = i & 15;
            const int k = k0 + c;
            float a = 0.f, b = 0.f;
            if (m0 + r < M && k < Kt)
                a = __bfloat162float(k < K ? A[(m0 + r) * lda + k] : A2[(m0 + r) * lda2 + (k - K)]);
            if (n0 + r < N && k < Kt) b = __bfloat162float(Wt[(long long)(n0 + r) * ldw + k]);
            sA[c][r] = a;
            sB[c][r] = b;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sB[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long m = m0 + ty * 4 + i;
            const int n = n0 + tx * 4 + j;
            if (m < M && n < N) epi_store(e, epi_value(e, acc[i][j], m, n), m, n);
        }
}

// Direct conv3x3 (pad 1) over NHWC bf16 with packed weights [Cout][9*C + Cs]; one thread per output.
__global__ void conv3x3_simt_kernel(const bf16* __restrict__ X, long long ldx, int nimg, int H, int W, int C,
                                    const bf16* __restrict__ Xs, long long lds, int Cs,
                                    const bf16* __restrict__ Wp, int Cout, Epilogue e) {
    pdl_prologue();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nimg * H * W * Cout;
    if (idx >= total) return;
    const int n = (int)(idx % Cout);
    const long long m = idx / Cout;
    const int w = (int)(m % W), h = (int)((m / W) % H);
    const long long img = m / ((long long)W * H);
    const long long ktot = 9LL * C + Cs;
    const bf16* wrow = Wp + (long long)n * ktot;
    float acc = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
        const bf16* xp = X + ((img * H + hh) * W + ww) * ldx;
        const bf16* wp = wrow + (long long)tap * C;
        for (int c = 0; c < C; ++c) acc = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), acc);
    }
    if (Cs) {
        const bf16* xp = Xs + m * lds;
        const bf16* wp = wrow + 9LL * C;
        for (int c = 0; c < Cs; ++c) acc = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), acc);
    }
    epi_store(e, epi_value(e, acc, m, n), m, n);
}

// First conv: x fp32 NCHW (Cin <= 4), w fp32 [Cout][Cin][3][3] -> bf16 NHWC (ld = ldo).
// One thread per (pixel, 8 output channels); weights and bias staged once per block in shared memory as
// [Cin][tap][Cout] so that a thread reads its 8 channels with two 16-byte loads per tap.
__global__ void __launch_bounds__(256)
conv3x3_in_kernel(const float* __restrict__ x, int nimg, int Cin, int H, int W, const float* __restrict__ w,
                  const float* __restrict__ bias, int Cout, bf16* __restrict__ out, long long ldo) {
    pdl_prologue();
    extern __shared__ __align__(16) float ws_in[];          // [Cin * 9][Cout] + [Cout]
    float* bs = ws_in + Cin * 9 * Cout;
    for (int d = threadIdx.x; d < Cout * Cin * 9; d += blockDim.x) {       // d = r * Cout + co: conflict-free smem writes
        const int r = d / Cout, co = d - r * Cout;                          // w[co][c][tap], r = c * 9 + tap
        ws_in[d] = __ldg(w + co * (Cin * 9) + r);
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int cg = Cout / 8;
    const long long total = (long long)nimg * H * W * cg;
    // grid-stride: the weights are staged once per block, not once per 256 outputs
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    const long long m = idx / cg;
    const int ww = (int)(m % W), hh = (int)((m / W) % H);
    const long long img = m / ((long long)W * H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bs[g * 8 + j];
    for (int c = 0; c < Cin; ++c) {
        const float* xp = x + (img * Cin + c) * H * W;
        float xv[9];                                       // all nine taps in flight (clamped address, masked value)
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y = hh + tap / 3 - 1, xx = ww + tap % 3 - 1;
            const bool ok = y >= 0 && y < H && xx >= 0 && xx < W;
            const float t = __ldg(xp + min(max(y, 0), H - 1) * W + min(max(xx, 0), W - 1));
            xv[tap] = ok ? t : 0.f;
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float v = xv[tap];
            const float4 w0 = *reinterpret_cast<const float4*>(ws_in + (c * 9 + tap) * Cout + g * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(ws_in + (c * 9 + tap) * Cout + g * 8 + 4);
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
    }
    *reinterpret_cast<bf16x8*>(out + m * ldo + g * 8) = pack8(acc);
    }
}

// Last conv: bf16 NHWC (C % 8 == 0, C / 8 a power of two <= 32) -> fp32 NCHW, Cout tiny.  C / 8 lanes per output
// pixel (8 channels each, one 16-byte load per tap), weights staged in shared memory as [Cout][tap][C], segmented
// shuffle reduction.  (The first version used one warp per pixel with half of the lanes idle at C = 128 and
// stride-9 weight loads: 95 us for a 17 MB input.)
__global__ void __launch_bounds__(256)
conv3x3_out_kernel(const bf16* __restrict__ X, long long ldx, int nimg, int H, int W, int C,
                   const float* __restrict__ w /*[Cout][C][3][3]*/, const float* __restrict__ bias, int Cout,
                   float* __restrict__ out) {
    pdl_prologue();
    extern __shared__ __align__(16) float ws_out[];         // [Cout][9][C]
    for (int d = threadIdx.x; d < Cout * C * 9; d += blockDim.x) {          // d = (co * 9 + tap) * C + c: conflict-free writes
        const int c = d % C, t2 = d / C, tap = t2 % 9, co = t2 / 9;
        ws_out[d] = __ldg(w + ((long long)co * C + c) * 9 + tap);
    }
    __syncthreads();
    const int lpp = C >> 3;                                  // lanes per pixel
    const long long total = (long long)nimg * H * W;
    const long long nthreads = (total * lpp + 31) / 32 * 32;                 // whole warps stay in the loop together
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nthreads; idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx / lpp;
    const int g = (int)(idx - pix * lpp);
    const bool live = pix < total;
    const long long pc = live ? pix : total - 1;
    const int ww = (int)(pc % W), hh = (int)((pc / W) % H);
    const long long img = pc / ((long long)W * H);
    for (int co = 0; co < Cout; ++co) {
        float acc = 0.f;
        bf16x8 xt[9];                                      // all nine taps in flight (clamped address, masked value)
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y = min(max(hh + tap / 3 - 1, 0), H - 1), xx = min(max(ww + tap % 3 - 1, 0), W - 1);
            xt[tap] = *reinterpret_cast<const bf16x8*>(X + ((img * H + y) * W + xx) * ldx + g * 8);
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y = hh + tap / 3 - 1, xx = ww + tap % 3 - 1;
            if (y < 0 || y >= H || xx < 0 || xx >= W) continue;
            float f[8];
            unpack8(xt[tap], f);
            const float4 w0 = *reinterpret_cast<const float4*>(ws_out + (co * 9 + tap) * C + g * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(ws_out + (co * 9 + tap) * C + g * 8 + 4);
            acc = fmaf(f[0], w0.x, acc); acc = fmaf(f[1], w0.y, acc); acc = fmaf(f[2], w0.z, acc); acc = fmaf(f[3], w0.w, acc);
            acc = fmaf(f[4], w1.x, acc); acc = fmaf(f[5], w1.y, acc); acc = fmaf(f[6], w1.z, acc); acc = fmaf(f[7], w1.w, acc);
        }
        for (int o = lpp >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (g == 0 && live) out[((img * Cout + co) * H + hh) * W + ww] = acc + (bias ? bias[co] : 0.f);
    }
    }
}

// ---- the shapes the score networks actually use: one image channel in / out.  The thread's 8 output (input) channels are
// fixed for the whole kernel, so its 72 weights live in registers instead of 18 shared-memory loads per pixel; a CTA walks
// a band of image rows (32-bit index arithmetic; the 3 x 3 neighbourhood of a band stays in L1).
__global__ void __launch_bounds__(256)
conv3x3_in1_kernel(const float* __restrict__ x, int nimg, int H, int W, const float* __restrict__ w /*[Cout][1][3][3]*/,
                   const float* __restrict__ bias, int Cout, bf16* __restrict__ out, long long ldo, int band) {
    pdl_prologue();
    const int cg = Cout >> 3;                                // threads per pixel
    const int g = threadIdx.x % cg, slot = threadIdx.x / cg, slots = blockDim.x / cg;
    float wr[9][8], br[8];
    {
        float flat[72];                                      // w[(g * 8 + j) * 9 + tap]: 72 consecutive floats, 18 x 16 bytes
#pragma unroll
        for (int i = 0; i < 18; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(w + g * 72) + i);
            flat[4 * i] = t.x; flat[4 * i + 1] = t.y; flat[4 * i + 2] = t.z; flat[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            br[j] = bias ? __ldg(bias + g * 8 + j) : 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) wr[tap][j] = flat[j * 9 + tap];
        }
    }
    const int bands = (H + band - 1) / band;
    const int img = blockIdx.x / bands, y0 = (blockIdx.x - img * bands) * band, y1 = min(H, y0 + band);
    const float* xp = x + (long long)img * H * W;
    bf16* op = out + (long long)img * H * W * ldo + g * 8;
    for (int pix = y0 * W + slot; pix < y1 * W; pix += slots) {
        const int hh = pix / W, ww = pix - hh * W;
        float xv[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y = hh + tap / 3 - 1, xx = ww + tap % 3 - 1;
            const bool ok = y >= 0 && y < H && xx >= 0 && xx < W;
            const float t = __ldg(xp + min(max(y, 0), H - 1) * W + min(max(xx, 0), W - 1));
            xv[tap] = ok ? t : 0.f;
        }
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = br[j];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv[tap], wr[tap][j], acc[j]);
        *reinterpret_cast<bf16x8*>(op + (long long)pix * ldo) = pack8(acc);
    }
}

__global__ void __launch_bounds__(256)
conv3x3_out1_kernel(const bf16* __restrict__ X, long long ldx, int nimg, int H, int W, int C,
                    const float* __restrict__ w /*[1][C][3][3]*/, const float* __restrict__ bias, float* __restrict__ out,
                    int band) {
    pdl_prologue();
    const int lpp = C >> 3;                                  // lanes per pixel (a power of two <= 32)
    const int g = threadIdx.x % lpp, slot = threadIdx.x / lpp, slots = blockDim.x / lpp;
    float wr[9][8];
    {
        float flat[72];                                      // w[(g * 8 + j) * 9 + tap]: 72 consecutive floats, 18 x 16 bytes
#pragma unroll
        for (int i = 0; i < 18; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(w + g * 72) + i);
            flat[4 * i] = t.x; flat[4 * i + 1] = t.y; flat[4 * i + 2] = t.z; flat[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int j = 0; j < 8; ++j) wr[tap][j] = flat[j * 9 + tap];
    }
    const float b0 = bias ? __ldg(bias) : 0.f;
    const int bands = (H + band - 1) / band;
    const int img = blockIdx.x / bands, y0 = (blockIdx.x - img * bands) * band, y1 = min(H, y0 + band);
    const bf16* xp = X + (long long)img * H * W * ldx + g * 8;
    float* op = out + (long long)img * H * W;
    const int n = (y1 - y0) * W;
    for (int i = slot; i < (n + slots - 1) / slots * slots; i += slots) {      // whole warps stay in the loop together
        const bool live = i < n;
        const int pix = y0 * W + (live ? i : n - 1);
        const int hh = pix / W, ww = pix - hh * W;
        bf16x8 xt[9];                                        // all nine taps in flight (clamped address, masked value)
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y = min(max(hh + tap / 3 - 1, 0), H - 1), xx = min(max(ww + tap % 3 - 1, 0), W - 1);
            xt[tap] = *reinterpret_cast<const bf16x8*>(xp + (long long)(y * W + xx) * ldx);
        }
        float acc = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y = hh + tap / 3 - 1, xx = ww + tap % 3 - 1;
            if (y < 0 || y >= H || xx < 0 || xx >= W) continue;
            float f[8];
            unpack8(xt[tap], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(f[j], wr[tap][j], acc);
        }
        for (int o = lpp >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (g == 0 && live) op[pix] = acc + b0;
    }
}

// ---- sliding-window variants (W * C / 8 <= 512 threads): a thread owns one image column and walks down the rows of its
// band with a 3 x 3 window in registers, so an output costs 3 new loads instead of 9 clamped ones (the band kernels above
// spend ~430 / ~250 instructions per (pixel, 8 channels), most of them on addressing: profiles/README.md).
__global__ void __launch_bounds__(512)
conv3x3_in1_slide_kernel(const float* __restrict__ x, int nimg, int H, int W, const float* __restrict__ w,
                         const float* __restrict__ bias, int Cout, bf16* __restrict__ out, long long ldo, int band) {
    pdl_prologue();
    const int cg = Cout >> 3;
    const int g = threadIdx.x % cg, xx = threadIdx.x / cg;
    float wr[9][8], br[8];
    {
        float flat[72];
#pragma unroll
        for (int i = 0; i < 18; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(w + g * 72) + i);
            flat[4 * i] = t.x; flat[4 * i + 1] = t.y; flat[4 * i + 2] = t.z; flat[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            br[j] = bias ? __ldg(bias + g * 8 + j) : 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) wr[tap][j] = flat[j * 9 + tap];
        }
    }
    const int bands = (H + band - 1) / band;
    const int img = blockIdx.x / bands, y0 = (blockIdx.x - img * bands) * band, y1 = min(H, y0 + band);
    const float* xp = x + (long long)img * H * W;
    bf16* op = out + (long long)img * H * W * ldo + g * 8;
    const bool has_l = xx > 0, has_r = xx + 1 < W;
    auto load_row = [&](int y, float (&r)[3]) {
        const bool ok = y >= 0 && y < H;
        const float* row = xp + min(max(y, 0), H - 1) * W + xx;
        r[0] = ok && has_l ? __ldg(row - 1) : 0.f;
        r[1] = ok ? __ldg(row) : 0.f;
        r[2] = ok && has_r ? __ldg(row + 1) : 0.f;
    };
    float top[3], mid[3], bot[3];
    load_row(y0 - 1, top);
    load_row(y0, mid);
    for (int y = y0; y < y1; ++y) {
        load_row(y + 1, bot);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = br[j];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(top[k], wr[k][j], acc[j]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(mid[k], wr[3 + k][j], acc[j]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(bot[k], wr[6 + k][j], acc[j]);
        *reinterpret_cast<bf16x8*>(op + (long long)(y * W + xx) * ldo) = pack8(acc);
#pragma unroll
        for (int k = 0; k < 3; ++k) { top[k] = mid[k]; mid[k] = bot[k]; }
    }
}

__global__ void __launch_bounds__(512)
conv3x3_out1_slide_kernel(const bf16* __restrict__ X, long long ldx, int nimg, int H, int W, int C,
                          const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int band) {
    pdl_prologue();
    const int lpp = C >> 3;
    const int g = threadIdx.x % lpp, xx = threadIdx.x / lpp;
    float wr[9][8];
    {
        float flat[72];
#pragma unroll
        for (int i = 0; i < 18; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(w + g * 72) + i);
            flat[4 * i] = t.x; flat[4 * i + 1] = t.y; flat[4 * i + 2] = t.z; flat[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int j = 0; j < 8; ++j) wr[tap][j] = flat[j * 9 + tap];
    }
    const float b0 = bias ? __ldg(bias) : 0.f;
    const int bands = (H + band - 1) / band;
    const int img = blockIdx.x / bands, y0 = (blockIdx.x - img * bands) * band, y1 = min(H, y0 + band);
    const bf16* xp = X + (long long)img * H * W * ldx + g * 8;
    float* op = out + (long long)img * H * W;
    const bool has_l = xx > 0, has_r = xx + 1 < W;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    auto load_row = [&](int y, uint4 (&r)[3]) {
        const bool ok = y >= 0 && y < H;
        const bf16* row = xp + (long long)(min(max(y, 0), H - 1) * W + xx) * ldx;
        r[0] = ok && has_l ? *reinterpret_cast<const uint4*>(row - ldx) : zero;
        r[1] = ok ? *reinterpret_cast<const uint4*>(row) : zero;
        r[2] = ok && has_r ? *reinterpret_cast<const uint4*>(row + ldx) : zero;
    };
    auto dot8 = [](const uint4& v, const float (&ww)[8], float acc) {
        float2 t;
        t = bf2_to_f2(v.x); acc = fmaf(t.x, ww[0], acc); acc = fmaf(t.y, ww[1], acc);
        t = bf2_to_f2(v.y); acc = fmaf(t.x, ww[2], acc); acc = fmaf(t.y, ww[3], acc);
        t = bf2_to_f2(v.z); acc = fmaf(t.x, ww[4], acc); acc = fmaf(t.y, ww[5], acc);
        t = bf2_to_f2(v.w); acc = fmaf(t.x, ww[6], acc); acc = fmaf(t.y, ww[7], acc);
        return acc;
    };
    uint4 top[3], mid[3], bot[3];
    load_row(y0 - 1, top);
    load_row(y0, mid);
    for (int y = y0; y < y1; ++y) {
        load_row(y + 1, bot);
        float acc = 0.f;                          // same order as the band kernel: taps 0..8, 8 channels each
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = dot8(top[k], wr[k], acc);
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = dot8(mid[k], wr[3 + k], acc);
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = dot8(bot[k], wr[6 + k], acc);
        for (int o = lpp >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (g == 0) op[y * W + xx] = acc + b0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { top[k] = mid[k]; mid[k] = bot[k]; }
    }
}

}  // namespace

extern "C" int xd_gemm_bf16_simt(const void* A, long long lda, const void* A2, long long lda2, int K2, const void* Wt,
                                 long long ldw, int M, int N, int K, const float* bias, int act, const float* gate,
                                 int gate_rows, long long gate_ld, const void* residual, int res_dtype,
                                 long long res_ld, void* out, int out_dtype, long long out_ld, void* stream) {
    XD_CHECK_ARG(A && Wt && out && M > 0 && N > 0 && K > 0 && (A2 != nullptr) == (K2 > 0));
    XD_CHECK_ARG(!gate || gate_rows > 0);
    Epilogue e{bias, gate, residual, out, gate_ld, res_ld, out_ld, act, gate_rows > 0 ? gate_rows : 1, res_dtype, out_dtype};
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    xd_launch(gemm_simt_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)A, lda, (const bf16*)A2, lda2, K2,
                                                             (const bf16*)Wt, ldw, M, N, K, e);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_conv3x3_bf16_simt(const void* X, long long ldx, int nimg, int H, int W, int C, const void* Xs,
                                    long long lds, int Cs, const void* Wp, int Cout, const float* bias, int act,
                                    const void* residual, int res_dtype, long long res_ld, void* out, int out_dtype,
                                    long long out_ld, void* stream) {
    XD_CHECK_ARG(X && Wp && out && nimg > 0 && (Xs != nullptr) == (Cs > 0));
    Epilogue e{bias, nullptr, residual, out, 0, res_ld, out_ld, act, 1, res_dtype, out_dtype};
    const long long total = (long long)nimg * H * W * Cout;
    xd_launch(conv3x3_simt_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, 
        (const bf16*)X, ldx, nimg, H, W, C, (const bf16*)Xs, lds, Cs, (const bf16*)Wp, Cout, e);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_conv3x3_in_f32_nchw(const float* x, int nimg, int Cin, int H, int W, const float* w,
                                      const float* bias, int Cout, void* out, long long ldo, void* stream) {
    XD_CHECK_ARG(x && w && out && Cout % 8 == 0 && ldo % 8 == 0);
    static const bool slide = getenv("XDB200_CONV_SLIDE") == nullptr || atoi(getenv("XDB200_CONV_SLIDE")) != 0;
    if (slide && Cin == 1 && W * (Cout / 8) <= 512 && (W * (Cout / 8)) % 32 == 0 && (long long)H * W < (1LL << 30) &&
        (reinterpret_cast<uintptr_t>(w) & 15) == 0) {                                 // sliding window: thread = (column, 8 channels)
        int band = H;
        while (band > 4 && (long long)nimg * ((H + band - 1) / band) < 148) band = (band + 1) / 2;
        const unsigned grid = (unsigned)(nimg * ((H + band - 1) / band));
        xd_launch(conv3x3_in1_slide_kernel, grid, W * (Cout / 8), 0, (cudaStream_t)stream, x, nimg, H, W, w, bias, Cout, (bf16*)out, ldo, band);
        XD_CHECK_LAUNCH();
        return XD_OK;
    }
    if (Cin == 1 && 256 % (Cout / 8) == 0 && (long long)H * W < (1LL << 30) && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {       // register-weight kernel, one band of rows per CTA
        int band = H;
        while (band > 1 && (long long)nimg * ((H + band - 1) / band) < 2 * 148) band = (band + 1) / 2;
        const unsigned grid = (unsigned)(nimg * ((H + band - 1) / band));
        xd_launch(conv3x3_in1_kernel, grid, 256, 0, (cudaStream_t)stream, x, nimg, H, W, w, bias, Cout, (bf16*)out, ldo, band);
        XD_CHECK_LAUNCH();
        return XD_OK;
    }
    const size_t smem = ((size_t)Cin * 9 * Cout + Cout) * sizeof(float);
    XD_CHECK_ARG(smem <= 48 * 1024);
    const long long total = (long long)nimg * H * W * (Cout / 8);
    xd_launch(conv3x3_in_kernel, (unsigned)std::min<long long>((total + 255) / 256, 148 * 8), 256, smem, (cudaStream_t)stream, x, nimg, Cin, H, W, w, bias,
                                                                                        Cout, (bf16*)out, ldo);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_conv3x3_out_f32_nchw(const void* X, long long ldx, int nimg, int H, int W, int C, const float* w,
                                       const float* bias, int Cout, float* out, void* stream) {
    XD_CHECK_ARG(X && w && out && C % 8 == 0 && ldx % 8 == 0);
    const int lpp = C / 8;
    XD_CHECK_ARG(lpp >= 1 && lpp <= 32 && (lpp & (lpp - 1)) == 0);
    static const bool slide = getenv("XDB200_CONV_SLIDE") == nullptr || atoi(getenv("XDB200_CONV_SLIDE")) != 0;
    if (slide && Cout == 1 && W * lpp <= 512 && (W * lpp) % 32 == 0 && (long long)H * W * ldx < (1LL << 31) &&
        (reinterpret_cast<uintptr_t>(w) & 15) == 0) {                                 // sliding window: thread = (column, 8 channels)
        int band = H;
        while (band > 4 && (long long)nimg * ((H + band - 1) / band) < 148) band = (band + 1) / 2;
        const unsigned grid = (unsigned)(nimg * ((H + band - 1) / band));
        xd_launch(conv3x3_out1_slide_kernel, grid, W * lpp, 0, (cudaStream_t)stream, (const bf16*)X, ldx, nimg, H, W, C, w, bias, out, band);
        XD_CHECK_LAUNCH();
        return XD_OK;
    }
    if (Cout == 1 && (long long)H * W * ldx < (1LL << 31) && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {                          // register-weight kernel, one band of rows per CTA
        int band = H;
        while (band > 2 && (long long)nimg * ((H + band - 1) / band) < 2 * 148) band = (band + 1) / 2;
        const unsigned grid = (unsigned)(nimg * ((H + band - 1) / band));
        xd_launch(conv3x3_out1_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)X, ldx, nimg, H, W, C, w, bias, out, band);
        XD_CHECK_LAUNCH();
        return XD_OK;
    }
    const size_t smem = (size_t)Cout * 9 * C * sizeof(float);
    XD_CHECK_ARG(smem <= 48 * 1024);
    const long long threads = (long long)nimg * H * W * lpp;
    xd_launch(conv3x3_out_kernel, (unsigned)std::min<long long>((threads + 255) / 256, 148 * 8), 256, smem, (cudaStream_t)stream, 
        (const bf16*)X, ldx, nimg, H, W, C, w, bias, Cout, out);
    XD_CHECK_LAUNCH();
    return XD_OK;
}
