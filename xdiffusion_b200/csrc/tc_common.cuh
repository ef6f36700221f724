// Helpers shared by the tcgen05 kernels (gemm_tc.cu, dit_block.cu): fast epilogue activations and the host-side
// tensor-map encoders.  sm_100a only.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

#include <mutex>
#include <stdio.h>

namespace tcx {

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

// (x + b) -> act, two columns at a time.  GELU uses the packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2): half the issue
// slots of the scalar form, same rounding (every step is the same rn operation), one MUFU.TANH per element.
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
template <int ACT> __device__ __forceinline__ void bias_act2(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, float& y0,
                                                             float& y1) {
    if constexpr (ACT == XD_ACT_GELU_TANH) {
        uint64_t x, x2, in, u, h, y;
        const uint64_t acc = pack_f32x2(__uint_as_float(a0), __uint_as_float(a1));
        const uint64_t bias = pack_f32x2(__uint_as_float(b0), __uint_as_float(b1));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(acc), "l"(bias));
        asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(x2) : "l"(x));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(in) : "l"(x2), "l"(pack_f32x2(0.035677408136300125f, 0.035677408136300125f)),
            "l"(pack_f32x2(0.7978845608028654f, 0.7978845608028654f)));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(x), "l"(in));
        float u0, u1, t0, t1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(u0), "=f"(u1) : "l"(u));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(h) : "l"(x), "l"(pack_f32x2(0.5f, 0.5f)));
        asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(y) : "l"(h), "l"(pack_f32x2(t0, t1)));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(y));
    } else {
        y0 = act_fast<ACT>(__uint_as_float(a0) + __uint_as_float(b0));
        y1 = act_fast<ACT>(__uint_as_float(a1) + __uint_as_float(b1));
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
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
inline int make_tmap(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
              CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { xd_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled entry point not found"); return XD_ERR_TMAP; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static char msg[160];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        xd_set_error(__FILE__, __LINE__, msg);
        return XD_ERR_TMAP;
    }
    return XD_OK;
}

// Epilogue boxes: 32 rows x 32 columns of the [M, N] output / residual (fp32: 128-byte rows, 128-byte swizzle;
// bf16: 64-byte rows, 64-byte swizzle).
inline int tmap_epi(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, bool f32) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {32, 32};
    return make_tmap(tm, ptr, 2, dims, str, box, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                     f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}


}  // namespace tcx
