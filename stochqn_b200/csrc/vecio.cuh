// vecio.cuh - 16-byte vector load / store helpers shared by every streaming kernel.
#pragma once
#include <cuda_runtime.h>

namespace sqn {

// ---- 16-byte vector access -------------------------------------------------------------
template <typename T, int VEC> struct Pack;
template <> struct Pack<double, 2> { double2 v; __device__ __forceinline__ double get(int i) const { return i == 0 ? v.x : v.y; }
                                     __device__ __forceinline__ void set(int i, double a) { if (i == 0) v.x = a; else v.y = a; } };
template <> struct Pack<float, 4>  { float4 v;  __device__ __forceinline__ float get(int i) const { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
                                     __device__ __forceinline__ void set(int i, float a) { if (i == 0) v.x = a; else if (i == 1) v.y = a; else if (i == 2) v.z = a; else v.w = a; } };
template <typename T> struct Pack<T, 1> { T v; __device__ __forceinline__ T get(int) const { return v; }
                                          __device__ __forceinline__ void set(int, T a) { v = a; } };

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_stream(const T* p)
{
    Pack<T, VEC> r;
    if constexpr (VEC == 1) r.v = __ldg(p);
    else if constexpr (sizeof(T) == 8) r.v = __ldg(reinterpret_cast<const double2*>(p));
    else r.v = __ldg(reinterpret_cast<const float4*>(p));
    return r;
}
// streaming row load: read-only path, do not allocate in L1 (the rows are touched exactly once per kernel; this
// keeps L1 for the probe chunks that all row-groups of a CTA share).  Measured on B200 (tools/kbench.cu): +1.5 %.
template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_row(const T* p)
{
    Pack<T, VEC> r;
    if constexpr (VEC == 1) {
        if constexpr (sizeof(T) == 8) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r.v) : "l"(p));
        else asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v) : "l"(p));
    } else if constexpr (sizeof(T) == 8) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.v.x), "=d"(r.v.y) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.v.x), "=f"(r.v.y), "=f"(r.v.z), "=f"(r.v.w) : "l"(p));
    }
    return r;
}
// same, coherent path: for rows of a buffer the kernel also writes (each address is read before it is written)
template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_row_rw(const T* p)
{
    Pack<T, VEC> r;
    if constexpr (VEC == 1) {
        if constexpr (sizeof(T) == 8) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r.v) : "l"(p) : "memory");
        else asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v) : "l"(p) : "memory");
    } else if constexpr (sizeof(T) == 8) {
        asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.v.x), "=d"(r.v.y) : "l"(p) : "memory");
    } else {
        asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.v.x), "=f"(r.v.y), "=f"(r.v.z), "=f"(r.v.w) : "l"(p) : "memory");
    }
    return r;
}
// plain (coherent) load for buffers that the same kernel also writes
template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_rw(const T* p)
{
    Pack<T, VEC> r;
    if constexpr (VEC == 1) r.v = *p;
    else if constexpr (sizeof(T) == 8) r.v = *reinterpret_cast<const double2*>(p);
    else r.v = *reinterpret_cast<const float4*>(p);
    return r;
}
template <typename T, int VEC>
__device__ __forceinline__ void st_vec(T* p, const Pack<T, VEC>& r)
{
    if constexpr (VEC == 1) *p = r.v;
    else if constexpr (sizeof(T) == 8) *reinterpret_cast<double2*>(p) = r.v;
    else *reinterpret_cast<float4*>(p) = r.v;
}

}  // namespace sqn
