// Shared device helpers for the sslap_b200 CUDA path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SSLAPB_WARP 32
#define SSLAPB_FULL 0xffffffffu

// Order-preserving map double -> uint64 (total order of IEEE-754 non-NaN values, -inf < ... < +inf).
// 0 is below every real bid (it is the image of a negative NaN), so 0 is the "no bid yet" sentinel of bidkey[].
__device__ __forceinline__ unsigned long long sslapb_ord64(double x)
{
    long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}

__device__ __forceinline__ unsigned long long sslapb_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned sslapb_ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int sslapb_ld_volatile_s32(const int *p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// 16-byte streaming loads of immutable CSR data (read-only path, do not pollute L1 for the price gathers).
__device__ __forceinline__ int4 sslapb_ldg_stream_i4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 sslapb_ldg_stream_d2(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// One 32-byte object record in ONE request (LDG.E.256, sm_100a): a random gather costs the L1TEX pipe per request, not per
// byte, so a 256-bit load halves the cost of fetching {start, owner, deg, price} compared with 128 + 64 bits.
struct SslapbRec256 { unsigned long long start, owner_deg, price_bits, pad; };
__device__ __forceinline__ SslapbRec256 sslapb_ld_rec256(const void *p)
{
    SslapbRec256 r;
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.start), "=l"(r.owner_deg), "=l"(r.price_bits), "=l"(r.pad) : "l"(p));
    return r;
}

__device__ __forceinline__ void sslapb_st_rec256(void *p, unsigned long long start, unsigned long long owner_deg,
                                                 unsigned long long price_bits)
{
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(start), "l"(owner_deg), "l"(price_bits), "l"(0ull) : "memory");
}

#define SSLAPB_NEG_INF (__longlong_as_double((long long)0xfff0000000000000ull))
