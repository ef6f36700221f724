// Shared device helpers for the xdb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define XD_OK 0
#define XD_ERR_ARG 1      // unsupported shape / alignment / null pointer
#define XD_ERR_CUDA 2     // launch or driver error
#define XD_ERR_TMAP 3     // cuTensorMapEncodeTiled failed

#define XD_CHECK_ARG(cond) do { if (!(cond)) { xd_set_error(__FILE__, __LINE__, #cond); return XD_ERR_ARG; } } while (0)
#define XD_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { xd_set_error(__FILE__, __LINE__, cudaGetErrorString(e__)); return XD_ERR_CUDA; } } while (0)

void xd_set_error(const char* file, int line, const char* msg);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// pdl_launch_dependents() + pdl_wait(): the NEXT kernel of the stream (or CUDA graph) may be scheduled
// as soon as all CTAs of this one have started, does its prologue (barrier init, TMEM allocation,
// tensor-map prefetch) and then blocks in griddepcontrol.wait until this kernel has completed and its
// memory is visible.  This removes the launch gap between the ~100-230 kernels of a sampling step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }

bool xd_pdl_enabled();       // XDB200_PDL=1: every kernel (abi.cu)
bool xd_split_enabled();     // xd_set_split_k / XDB200_SPLITK: batch-size-dependent work splits allowed (gemm_tc.cu)
bool xd_pdl_enabled_gemm();  // XDB200_PDL=1 or 2 (default): the tcgen05 GEMM / conv launches

template <typename... KArgs, typename... Args>
inline cudaError_t xd_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = xd_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

enum { XD_F32 = 0, XD_BF16 = 1 };
enum { XD_ACT_NONE = 0, XD_ACT_SILU = 1, XD_ACT_GELU_TANH = 2 };

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_tanh_f(float x) {
    // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  == x * sigmoid(2u)
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return x / (1.0f + __expf(-2.0f * u));
}
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == XD_ACT_SILU) return silu_f(v);
    if (act == XD_ACT_GELU_TANH) return gelu_tanh_f(v);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 8 x bf16 <-> 8 x float through ONE 16-byte access.  The payload is a uint4 on purpose: a struct of
// four __nv_bfloat162 is copied with four 32-bit LDG/STG (measured: 32 sectors per request), a uint4
// member compiles to LDG.128 / STG.128.
struct __align__(16) bf16x8 { uint4 u; };
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
    float2 t;
    t = bf2_to_f2(p.u.x); f[0] = t.x; f[1] = t.y;
    t = bf2_to_f2(p.u.y); f[2] = t.x; f[3] = t.y;
    t = bf2_to_f2(p.u.z); f[4] = t.x; f[5] = t.y;
    t = bf2_to_f2(p.u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
    bf16x8 p;
    p.u = make_uint4(f2_to_bf2(f[0], f[1]), f2_to_bf2(f[2], f[3]), f2_to_bf2(f[4], f[5]), f2_to_bf2(f[6], f[7]));
    return p;
}

// Generic epilogue description shared by the tcgen05 and the SIMT GEMM/conv kernels:
//   v = acc + bias[n];  v = act(v);  v *= gate[(m / gate_rows) * gate_ld + n];
//   v += residual[m * res_ld + n];  out[m * out_ld + n] = v   (fp32 or bf16)
struct Epilogue {
    const float* bias;
    const float* gate;
    const void* residual;
    void* out;
    long long gate_ld, res_ld, out_ld;
    int act, gate_rows, res_dtype, out_dtype;
};

__device__ __forceinline__ float epi_value(const Epilogue& e, float acc, long long m, int n) {
    float v = acc;
    if (e.bias) v += __ldg(e.bias + n);
    v = apply_act(v, e.act);
    if (e.gate) v *= __ldg(e.gate + (m / e.gate_rows) * e.gate_ld + n);
    if (e.residual) {
        if (e.res_dtype == XD_F32) v += ((const float*)e.residual)[m * e.res_ld + n];
        else v += __bfloat162float(((const bf16*)e.residual)[m * e.res_ld + n]);
    }
    return v;
}
__device__ __forceinline__ void epi_store(const Epilogue& e, float v, long long m, int n) {
    if (e.out_dtype == XD_F32) ((float*)e.out)[m * e.out_ld + n] = v;
    else ((bf16*)e.out)[m * e.out_ld + n] = __float2bfloat16_rn(v);
}
