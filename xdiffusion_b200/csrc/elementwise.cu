// Small bandwidth-bound kernels around the score networks: sinusoidal timestep embeddings (both
// reference formulas), patchify / unpatchify, class-embedding combine, 2x2 average pool, nearest
// 2x upsample, activation casts, broadcast adds and the classifier-free-guidance combine.
#include "common.cuh"

namespace {

__device__ __forceinline__ float load_time(const void* t, int is_i64, long long i) {
    return is_i64 ? (float)((const long long*)t)[i] : ((const float*)t)[i];
}

// out[b, :] = [sin(a) | cos(a)] (order 0) or [cos(a) | sin(a)] (order 1), a = tx * freq[i], with
//   mode 0: tx = t                                  (layers/utils.py:102-117, DiT)
//   mode 1: tx = t * 1000 / max_time                (layers/embedding.py:66-76, UNet)
//   mode 2: tx = atan(exp(-0.5 clip(t,lo,hi))) / (pi/2) * 1000 / max_time   (embedding.py:131-133)
// freq is the host-built fp32 table (same torch ops as the reference), so only sinf/cosf differ.
__global__ void sinusoid_kernel(const void* t, int is_i64, int B, const float* __restrict__ freq, int half, int mode,
                                float max_time, float clip_lo, float clip_hi, int order, float* out_f32,
                                bf16* out_bf16) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, j = i % half;
    float tx = load_time(t, is_i64, b);
    if (mode == 2) {
        tx = fminf(fmaxf(tx, clip_lo), clip_hi);
        tx = atanf(expf(-0.5f * tx)) / (0.5f * 3.14159265358979323846f);
    }
    if (mode >= 1) tx = __fdiv_rn(__fmul_rn(tx, 1000.0f), max_time);
    const float a = __fmul_rn(tx, freq[j]);
    float s, c;
    sincosf(a, &s, &c);
    const float first = order == 0 ? s : c, second = order == 0 ? c : s;
    const long long o = (long long)b * 2 * half;
    if (out_f32) { out_f32[o + j] = first; out_f32[o + half + j] = second; }
    if (out_bf16) { out_bf16[o + j] = __float2bfloat16_rn(first); out_bf16[o + half + j] = __float2bfloat16_rn(second); }
}

// act + cast, n elements
__global__ void act_cast_kernel(const void* in, int in_dtype, void* out, int out_dtype, int act, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = in_dtype == XD_F32 ? ((const float*)in)[i] : __bfloat162float(((const bf16*)in)[i]);
    v = apply_act(v, act);
    if (out_dtype == XD_F32) ((float*)out)[i] = v;
    else ((bf16*)out)[i] = __float2bfloat16_rn(v);
}

// fp32 -> bf16, 8 elements per thread (two 16-byte loads, one 16-byte store); n % 8 == 0, 16-byte aligned pointers
__global__ void act_cast_f32_bf16_v8_kernel(const float4* __restrict__ in, uint4* __restrict__ out, int act, long long n8) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = in[2 * i], b = in[2 * i + 1];
    out[i] = make_uint4(f2_to_bf2(apply_act(a.x, act), apply_act(a.y, act)), f2_to_bf2(apply_act(a.z, act), apply_act(a.w, act)),
                        f2_to_bf2(apply_act(b.x, act), apply_act(b.y, act)), f2_to_bf2(apply_act(b.z, act), apply_act(b.w, act)));
}

// c = table[label] + temb ; optionally also silu(c) as bf16 (input of every adaLN GEMM)
__global__ void class_combine_kernel(const float* __restrict__ table, const long long* __restrict__ labels,
                                     const float* __restrict__ temb, int B, int Dm, float* c_out, bf16* silu_out) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Dm) return;
    const int b = i / Dm, d = i % Dm;
    float v = temb[i];
    if (table) v = table[labels[b] * Dm + d] + v;
    if (c_out) c_out[i] = v;
    if (silu_out) silu_out[i] = __float2bfloat16_rn(silu_f(v));
}

// The same with the timestep embedding taken from a per-loop table: temb = temb_table[*idx] for every row (the sampling
// loop evaluates the timestep MLP once for all N timesteps when it is built; the loop index lives on the device).
__global__ void class_combine_step_kernel(const float* __restrict__ table, const long long* __restrict__ labels,
                                          const float* __restrict__ temb_table, const int* __restrict__ idx, int B, int Dm,
                                          float* c_out, bf16* silu_out, int* rows_out, int n_steps) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Dm) return;
    const int b = i / Dm, d = i % Dm;
    if (rows_out && d == 0) rows_out[b] = (labels ? (int)labels[b] : 0) * n_steps + *idx;
    float v = temb_table[(long long)(*idx) * Dm + d];
    if (table) v = table[labels[b] * Dm + d] + v;
    if (c_out) c_out[i] = v;
    if (silu_out) silu_out[i] = __float2bfloat16_rn(silu_f(v));
}

// out[0 .. W) = table[*idx][0 .. W): the current timestep's row of a per-loop table (the loop index lives on the device)
__global__ void gather_row_kernel(const float4* __restrict__ table, long long ld4, const int* __restrict__ idx, int W4,
                                  float4* __restrict__ out) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W4) out[i] = table[(long long)(*idx) * ld4 + i];
}

// x fp32 NCHW -> bf16 [B*gh*gw, C*p*p], column order (c, py, px) = flattened Conv2d weight
__global__ void patchify_kernel(const float* __restrict__ x, int B, int C, int H, int W, int p, bf16* out) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int gh = H / p, gw = W / p, K = C * p * p;
    const long long total = (long long)B * gh * gw * K;
    if (i >= total) return;
    const int k = (int)(i % K);
    const long long tok = i / K;
    const int px = k % p, py = (k / p) % p, c = k / (p * p);
    const int tw = (int)(tok % gw), th = (int)((tok / gw) % gh);
    const long long b = tok / ((long long)gw * gh);
    out[i] = __float2bfloat16_rn(x[((b * C + c) * H + th * p + py) * W + tw * p + px]);
}

// y fp32 [B*gh*gw, p*p*c] -> fp32 NCHW : img[n, c, h*p+py, w*p+px] = y[n, h, w, py, px, c]  (dit.py:187-204)
__global__ void unpatchify_kernel(const float* __restrict__ y, long long ldy, int B, int C, int H, int W, int p,
                                  float* out) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * C * H * W;
    if (i >= total) return;
    const int xx = (int)(i % W), yy = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C);
    const long long b = i / ((long long)W * H * C);
    const int gw = W / p, gh = H / p;
    const long long tok = (b * gh + yy / p) * gw + xx / p;
    out[i] = y[tok * ldy + ((yy % p) * p + xx % p) * C + c];
}

// out[r, c] = a[r, c] + b[(r % period), c]   (fp32; token position embeddings: period = tokens per sample)
__global__ void add_rows_periodic_kernel(const float* a, const float* b, long long rows, int cols, int period,
                                         float* out) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = (int)(i % cols);
    out[i] = a[i] + b[(r % period) * cols + c];
}
// 16-byte version (cols % 4 == 0, 16-byte aligned pointers, < 2^32 vectors): 32-bit index arithmetic
__global__ void add_rows_periodic_v4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, unsigned nvec,
                                            unsigned cols4, unsigned period, float4* __restrict__ out) {
    pdl_prologue();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    const unsigned r = i / cols4, c = i - r * cols4;
    const float4 x = a[i], y = __ldg(&b[(r % period) * cols4 + c]);
    out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

// out[g, r, c] = a[r, c] + tab[g, c]    (PixArt adaLN-single: table(6*D) + t0, for all blocks at once)
__global__ void add_table_kernel(const float* a, const float* tab, int G, int R, int Ccols, float* out) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)G * R * Ccols;
    if (i >= total) return;
    const int c = (int)(i % Ccols);
    const int r = (int)((i / Ccols) % R);
    const int g = (int)(i / ((long long)Ccols * R));
    out[i] = a[(long long)r * Ccols + c] + tab[(long long)g * Ccols + c];
}

// 16-byte version of add_table (Ccols % 4 == 0, aligned pointers, < 2^32 vectors)
__global__ void add_table_v4_kernel(const float4* __restrict__ a, const float4* __restrict__ tab, unsigned G, unsigned R,
                                    unsigned C4, float4* __restrict__ out) {
    pdl_prologue();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * R * C4) return;
    const unsigned row = i / C4, c = i - row * C4;          // row = g * R + r
    const unsigned g = row / R, r = row - g * R;
    const float4 x = __ldg(&a[r * C4 + c]), t = __ldg(&tab[g * C4 + c]);
    out[i] = make_float4(x.x + t.x, x.y + t.y, x.z + t.z, x.w + t.w);
}

// 2x2 average pool / nearest 2x upsample over NHWC bf16 (8 channels per thread)
__global__ void avgpool2_kernel(const bf16* __restrict__ x, long long ldx, int nimg, int H, int W, int C, bf16* out,
                                long long ldo) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int V = C / 8, Ho = H / 2, Wo = W / 2;
    const long long total = (long long)nimg * Ho * Wo * V;
    if (i >= total) return;
    const int j = (int)(i % V);
    const long long m = i / V;
    const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho);
    const long long img = m / ((long long)Wo * Ho);
    float acc[8] = {};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            float f[8];
            unpack8(*reinterpret_cast<const bf16x8*>(x + ((img * H + 2 * ho + dy) * W + 2 * wo + dx) * ldx + j * 8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= 0.25f;
    *reinterpret_cast<bf16x8*>(out + m * ldo + j * 8) = pack8(acc);
}

__global__ void upsample2_kernel(const bf16* __restrict__ x, long long ldx, int nimg, int H, int W, int C, bf16* out,
                                 long long ldo) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int V = C / 8, Ho = H * 2, Wo = W * 2;
    const long long total = (long long)nimg * Ho * Wo * V;
    if (i >= total) return;
    const int j = (int)(i % V);
    const long long m = i / V;
    const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho);
    const long long img = m / ((long long)Wo * Ho);
    *reinterpret_cast<bf16x8*>(out + m * ldo + j * 8) =
        *reinterpret_cast<const bf16x8*>(x + ((img * H + ho / 2) * W + wo / 2) * ldx + j * 8);
}

// rows x C bf16 copy between strided buffers; row r = (batch r / rpb, row r % rpb) with separate batch strides (e.g. the
// self-attention [q | k | v] rows of every image appended behind that image's encoder rows: layers/attention.py:166-180)
__global__ void copy_rows_kernel(const bf16* __restrict__ x, long long ldx, long long x_bs, long long rows, long long rpb,
                                 int C, bf16* out, long long ldo, long long o_bs) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int V = C / 8;
    if (i >= rows * V) return;
    const long long r = i / V;
    const int j = (int)(i % V);
    const long long b = r / rpb, t = r - b * rpb;
    *reinterpret_cast<bf16x8*>(out + b * o_bs + t * ldo + j * 8) =
        *reinterpret_cast<const bf16x8*>(x + b * x_bs + t * ldx + j * 8);
}

// Stride-2 3x3 convolution (pad 1) as im2col + GEMM: out[(n, y, x), tap * C + c] = X[n, 2y + dy - 1, 2x + dx - 1, c] (0 outside),
// tap = 3 (dy) + dx, the K order of pack_conv3x3.  (DBlock._downsampling_convolution of the Imagen efficient UNet,
// layers/resnet.py:272-280; 8 channels = 16 bytes per thread.)
__global__ void im2col3x3_s2_kernel(const bf16* __restrict__ x, long long ldx, int H, int W, int C, bf16* __restrict__ out,
                                    long long total) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int V = C / 8, Ho = H / 2, Wo = W / 2;
    const int j = (int)(i % V);
    long long r = i / V;
    const int tap = (int)(r % 9);
    r /= 9;                                                 // output pixel (n, y, x)
    const int xo = (int)(r % Wo), yo = (int)((r / Wo) % Ho);
    const long long n = r / ((long long)Wo * Ho);
    const int yi = 2 * yo + tap / 3 - 1, xi = 2 * xo + tap % 3 - 1;
    bf16x8 v = {};
    if (yi >= 0 && yi < H && xi >= 0 && xi < W) v = *reinterpret_cast<const bf16x8*>(x + ((n * H + yi) * W + xi) * ldx + j * 8);
    *reinterpret_cast<bf16x8*>(out + (r * 9 + tap) * C + j * 8) = v;
}

// out[n, p, :] = x[n, p, :] + b[n, :]   (bf16 rows, fp32 per-sample channel bias: `h + Linear(SiLU(temb))[..., None, None]`,
// layers/resnet.py:303-310,395-402)
__global__ void add_channel_bias_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ b, long long ldb,
                                        long long P, int C, bf16* __restrict__ out, long long ldo, long long total) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int V = C / 8;
    const int j = (int)(i % V);
    const long long row = i / V, n = row / P;
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + row * ldx + j * 8), f);
    const float4 b0 = *reinterpret_cast<const float4*>(b + n * ldb + j * 8), b1 = *reinterpret_cast<const float4*>(b + n * ldb + j * 8 + 4);
    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    *reinterpret_cast<bf16x8*>(out + row * ldo + j * 8) = pack8(f);
}

// x[b, c, f, :] = mask[b, f] ? x[b, c, f, :] : x0[b, c, f, :]   (video-mask blend, diffusion/ddpm.py:963-982), float4
__global__ void blend_frames_kernel(float4* __restrict__ x, const float4* __restrict__ x0, const uint8_t* __restrict__ mask,
                                    int C, int F, int hw4, long long n4) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long plane = i / hw4;                       // (b, c, f)
    const int f = (int)(plane % F);
    const long long b = plane / ((long long)F * C);
    if (!mask[b * F + f]) x[i] = x0[i];
}

// eps = u + w (c - u)   (samplers/ancestral.py:229-231), float4 vectorised
__global__ void cfg_kernel(const float4* __restrict__ c, const float4* __restrict__ u, float w, float4* out,
                           long long n4) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 a = c[i], b = u[i];
    float4 r;
    r.x = __fadd_rn(b.x, __fmul_rn(w, __fsub_rn(a.x, b.x)));
    r.y = __fadd_rn(b.y, __fmul_rn(w, __fsub_rn(a.y, b.y)));
    r.z = __fadd_rn(b.z, __fmul_rn(w, __fsub_rn(a.z, b.z)));
    r.w = __fadd_rn(b.w, __fmul_rn(w, __fsub_rn(a.w, b.w)));
    out[i] = r;
}

inline unsigned blocks_for(long long n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

extern "C" int xd_timestep_embed(const void* t, int t_is_i64, int B, const float* freq, int half, int mode,
                                 float max_time, float clip_lo, float clip_hi, int order, float* out_f32,
                                 void* out_bf16, void* stream) {
    XD_CHECK_ARG(t && freq && B > 0 && half > 0 && (out_f32 || out_bf16) && mode >= 0 && mode <= 2);
    xd_launch(sinusoid_kernel, blocks_for((long long)B * half), 256, 0, (cudaStream_t)stream, 
        t, t_is_i64, B, freq, half, mode, max_time, clip_lo, clip_hi, order, out_f32, (bf16*)out_bf16);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_act_cast(const void* in, int in_dtype, void* out, int out_dtype, int act, long long n,
                           void* stream) {
    XD_CHECK_ARG(in && out && n > 0);
    if (in_dtype == XD_F32 && out_dtype == XD_BF16 && n % 8 == 0 &&
        ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
        xd_launch(act_cast_f32_bf16_v8_kernel, blocks_for(n / 8), 256, 0, (cudaStream_t)stream, (const float4*)in, (uint4*)out, act, n / 8);
    else
        xd_launch(act_cast_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, in, in_dtype, out, out_dtype, act, n);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_class_combine(const float* table, const long long* labels, const float* temb, int B, int Dm,
                                float* c_out, void* silu_out, void* stream) {
    XD_CHECK_ARG(temb && (c_out || silu_out) && (table == nullptr) == (labels == nullptr));
    xd_launch(class_combine_kernel, blocks_for((long long)B * Dm), 256, 0, (cudaStream_t)stream, table, labels, temb, B, Dm,
                                                                                        c_out, (bf16*)silu_out);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_class_combine_step(const float* table, const long long* labels, const float* temb_table, const int* idx_dev,
                                     int B, int Dm, float* c_out, void* silu_out, int* rows_out, int n_steps, void* stream) {
    XD_CHECK_ARG(temb_table && idx_dev && (c_out || silu_out) && (table == nullptr) == (labels == nullptr) && B > 0 && Dm > 0);
    xd_launch(class_combine_step_kernel, blocks_for((long long)B * Dm), 256, 0, (cudaStream_t)stream, table, labels, temb_table,
              idx_dev, B, Dm, c_out, (bf16*)silu_out, rows_out, n_steps);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_gather_row_f32(const float* table, long long ld, const int* idx_dev, int W, float* out, void* stream) {
    XD_CHECK_ARG(table && idx_dev && out && W > 0 && W % 4 == 0 && ld % 4 == 0 &&
                 ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    xd_launch(gather_row_kernel, blocks_for(W / 4), 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(table), ld / 4,
              idx_dev, W / 4, reinterpret_cast<float4*>(out));
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_patchify(const float* x, int B, int C, int H, int W, int p, void* out, void* stream) {
    XD_CHECK_ARG(x && out && p > 0 && H % p == 0 && W % p == 0);
    xd_launch(patchify_kernel, blocks_for((long long)B * C * H * W), 256, 0, (cudaStream_t)stream, x, B, C, H, W, p,
                                                                                           (bf16*)out);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_unpatchify(const float* y, long long ldy, int B, int C, int H, int W, int p, float* out,
                             void* stream) {
    XD_CHECK_ARG(y && out && p > 0 && H % p == 0 && W % p == 0);
    xd_launch(unpatchify_kernel, blocks_for((long long)B * C * H * W), 256, 0, (cudaStream_t)stream, y, ldy, B, C, H, W, p,
                                                                                             out);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_add_rows_periodic(const float* a, const float* b, long long rows, int cols, int period, float* out,
                                    void* stream) {
    XD_CHECK_ARG(a && b && out && period > 0);
    const long long nvec = rows * cols / 4;
    const bool v4 = cols % 4 == 0 && nvec < (1LL << 32) &&
                    ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (v4)
        xd_launch(add_rows_periodic_v4_kernel, blocks_for(nvec), 256, 0, (cudaStream_t)stream, (const float4*)a,
                  (const float4*)b, (unsigned)nvec, (unsigned)(cols / 4), (unsigned)period, (float4*)out);
    else
        xd_launch(add_rows_periodic_kernel, blocks_for(rows * cols), 256, 0, (cudaStream_t)stream, a, b, rows, cols, period, out);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_add_table(const float* a, const float* tab, int G, int R, int C, float* out, void* stream) {
    XD_CHECK_ARG(a && tab && out);
    const long long nvec = (long long)G * R * C / 4;
    if (C % 4 == 0 && nvec < (1LL << 32) &&
        ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(tab) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
        xd_launch(add_table_v4_kernel, blocks_for(nvec), 256, 0, (cudaStream_t)stream, (const float4*)a, (const float4*)tab,
                  (unsigned)G, (unsigned)R, (unsigned)(C / 4), (float4*)out);
    else
        xd_launch(add_table_kernel, blocks_for((long long)G * R * C), 256, 0, (cudaStream_t)stream, a, tab, G, R, C, out);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_avgpool2x2_nhwc(const void* x, long long ldx, int nimg, int H, int W, int C, void* out,
                                  long long ldo, void* stream) {
    XD_CHECK_ARG(x && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && ldx % 8 == 0 && ldo % 8 == 0);
    xd_launch(avgpool2_kernel, blocks_for((long long)nimg * (H / 2) * (W / 2) * (C / 8)), 256, 0, (cudaStream_t)stream, 
        (const bf16*)x, ldx, nimg, H, W, C, (bf16*)out, ldo);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_upsample2x_nhwc(const void* x, long long ldx, int nimg, int H, int W, int C, void* out,
                                  long long ldo, void* stream) {
    XD_CHECK_ARG(x && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0);
    xd_launch(upsample2_kernel, blocks_for((long long)nimg * H * 2 * W * 2 * (C / 8)), 256, 0, (cudaStream_t)stream, 
        (const bf16*)x, ldx, nimg, H, W, C, (bf16*)out, ldo);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_copy_rows_bf16(const void* x, long long ldx, long long x_bs, long long rows, long long rows_per_batch,
                                 int C, void* out, long long ldo, long long o_bs, void* stream) {
    XD_CHECK_ARG(x && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && x_bs % 8 == 0 && o_bs % 8 == 0 &&
                 rows_per_batch > 0 && rows > 0);
    xd_launch(copy_rows_kernel, blocks_for(rows * (C / 8)), 256, 0, (cudaStream_t)stream, (const bf16*)x, ldx, x_bs, rows,
              rows_per_batch, C, (bf16*)out, ldo, o_bs);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_im2col3x3_s2_nhwc(const void* x, long long ldx, int nimg, int H, int W, int C, void* out, void* stream) {
    XD_CHECK_ARG(x && out && nimg > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && ldx % 8 == 0);
    const long long total = (long long)nimg * (H / 2) * (W / 2) * 9 * (C / 8);
    xd_launch(im2col3x3_s2_kernel, blocks_for(total), 256, 0, (cudaStream_t)stream, (const bf16*)x, ldx, H, W, C, (bf16*)out,
              total);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_add_channel_bias_nhwc(const void* x, long long ldx, const float* b, long long ldb, int nsamples, long long P,
                                        int C, void* out, long long ldo, void* stream) {
    XD_CHECK_ARG(x && b && out && nsamples > 0 && P > 0 && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldb % 4 == 0);
    const long long total = (long long)nsamples * P * (C / 8);
    xd_launch(add_channel_bias_kernel, blocks_for(total), 256, 0, (cudaStream_t)stream, (const bf16*)x, ldx, b, ldb, P, C,
              (bf16*)out, ldo, total);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_blend_frames(float* x, const float* x0, const void* mask, int B, int C, int F, int HW, void* stream) {
    XD_CHECK_ARG(x && x0 && mask && B > 0 && C > 0 && F > 0 && HW > 0 && HW % 4 == 0);
    const long long n4 = (long long)B * C * F * (HW / 4);
    xd_launch(blend_frames_kernel, blocks_for(n4), 256, 0, (cudaStream_t)stream, (float4*)x, (const float4*)x0,
              (const uint8_t*)mask, C, F, HW / 4, n4);
    XD_CHECK_LAUNCH();
    return XD_OK;
}

extern "C" int xd_cfg_combine(const float* cond, const float* uncond, float w, float* out, long long n,
                              void* stream) {
    XD_CHECK_ARG(cond && uncond && out && n % 4 == 0);
    xd_launch(cfg_kernel, blocks_for(n / 4), 256, 0, (cudaStream_t)stream, (const float4*)cond, (const float4*)uncond, w,
                                                                    (float4*)out, n / 4);
    XD_CHECK_LAUNCH();
    return XD_OK;
}
