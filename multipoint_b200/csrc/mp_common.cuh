// Shared helpers for the multipoint_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "multipoint_b200.h"

namespace mp {

// ---- error reporting across the C ABI: thread-local message, integer status ----
void set_error(const char *fmt, ...);
void count_launch(unsigned n = 1);
// profiling (mp_profile_begin / mp_profile_end): an event after each launch, one at each API entry;
// a kernel's duration is the time between its event and the previous one on the same stream
void prof_mark(const char *name, cudaStream_t stream);
inline void prof_entry(cudaStream_t stream) { prof_mark(nullptr, stream); }

#define MP_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            mp::set_error(__VA_ARGS__);         \
            return MP_ERR_INVALID;              \
        }                                       \
    } while (0)

#define MP_CUDA_OK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            mp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return MP_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// after every kernel launch: count it, (when profiling) drop a CUDA event behind it on the launching
// stream, and surface launch errors
#define MP_LAUNCH_OK_S(name, stream)                                                       \
    do {                                                                                   \
        mp::count_launch();                                                                \
        mp::prof_mark(name, (cudaStream_t)(stream));                                       \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            mp::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return MP_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Workspace {
    char *base;
    size_t size, off;
    Workspace(void *p, size_t n) : base((char *)p), size(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        off = align_up(off, 256);
        T *r = (T *)(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return base != nullptr && off <= size; }
};

// ---- device helpers ----
// a / b for many numerators and one denominator: r = RN(1 / b) once, then per value a multiply and one FMA correction
// step (Markstein: q' = q + (a - q b) r is the correctly rounded quotient when r is the correctly rounded reciprocal and
// nothing under- or overflows) -- 3-4 instructions instead of div.rn's ~12 with its range check and slow-path call.
// Valid for 1e-30 <= b <= 1e30 (callers rescale beyond that); an infinite numerator stays infinite.
__device__ __forceinline__ float rcp_rn(float x) {
#ifdef __CUDA_ARCH__
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
struct SharedDivisor {
    float b, r;
    __device__ __forceinline__ explicit SharedDivisor(float denom) : b(denom), r(rcp_rn(denom)) {}
    __device__ __forceinline__ float operator()(float a) const {
        const float q = a * r;
        const float q1 = fmaf(fmaf(-q, b, a), r, q);
        return fabsf(q) <= 3.4e38f ? q1 : q;   // +-inf (and NaN) pass through like a / b
    }
};
// v[i] = v[i] / denom for a register array; a denominator above 1e30 (no real descriptor) is scaled into range first
template <int N>
__device__ __forceinline__ void divide_all(float (&v)[N], float denom) {
    if (denom <= 1e30f) {
        const SharedDivisor div(denom);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = div(v[i]);
    } else {
        const SharedDivisor div(denom * 0x1p-64f);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = div(v[i] * 0x1p-64f);
    }
}

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace mp
