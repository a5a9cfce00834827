// ptx_ops.cuh -- the carry-chain integer instructions the field arithmetic is built from.
//
// Device: one `asm volatile` per PTX instruction (mad.lo.cc / madc.hi.cc / add.cc ...).  ptxas
// fuses adjacent {mad.lo.cc, madc.hi.cc} pairs on the same operands into one IMAD.WIDE.U32 with
// carry-in/out predicates, which is what makes a Montgomery product cost ~2*N^2 IMAD.WIDE.
//
// Host (plain g++, used only by the CPU-side self tests in tests/host/): the same functions are
// emulated bit-exactly with an explicit carry flag, so the exact instruction sequences of
// field.cuh can be verified against the oracle without a GPU.  This is NOT a product code path:
// the shipped library is device code only and there is no CPU fallback.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __device__ __forceinline__
#define B200_HDM __device__ __forceinline__      /* static member functions */
#define B200_HOSTDEV __host__ __device__ __forceinline__   /* plain C++ helpers with no PTX inside */
#else
#define B200_HD static inline
#define B200_HDM inline
#define B200_HOSTDEV static inline
#endif

#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; };          // host test shim only
#endif

namespace ptx {

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
// carry-free 32 x 32 + 64 -> 64 (plain IMAD.WIDE.U32: full IMAD rate, unlike the .X carry variants)
__device__ __forceinline__ unsigned long long mad_wide(uint32_t a, uint32_t b, unsigned long long c) {
    unsigned long long r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ unsigned long long mul_wide(uint32_t a, uint32_t b) {
    unsigned long long r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}

// 256-bit global load / store (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256), p 32-byte aligned: one request per 32-byte
// sector instead of two 16-byte ones -- the gather-bound MSM kernels are limited by outstanding sector requests
__device__ __forceinline__ void ld_global_256(const void* p, uint32_t* r) {
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "l"(p)
                 : "memory");
}

#else  // ---------------------------------------------------------------- host emulation
static inline void ld_global_256(const void* p, uint32_t* r) { __builtin_memcpy(r, p, 32); }
static inline void st_global_256(void* p, const uint32_t* r) { __builtin_memcpy(p, r, 32); }

static thread_local uint32_t CF = 0;   // the PTX condition-code carry flag CC.CF

static inline uint32_t add_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a + b; CF = (uint32_t)(t >> 32); return (uint32_t)t;
}
static inline uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a + b + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t;
}
static inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF; }
static inline uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a - b; CF = (uint32_t)(t >> 63); return (uint32_t)t;
}
static inline uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a - b - CF; CF = (uint32_t)(t >> 63); return (uint32_t)t;
}
static inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF; }
static inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
static inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t t = (uint64_t)mul_lo(a, b) + c; CF = (uint32_t)(t >> 32); return (uint32_t)t;
}
static inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t t = (uint64_t)mul_lo(a, b) + c + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t;
}
static inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t t = (uint64_t)mul_hi(a, b) + c; CF = (uint32_t)(t >> 32); return (uint32_t)t;
}
static inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t t = (uint64_t)mul_hi(a, b) + c + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t;
}
static inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + CF; }
static inline unsigned long long mad_wide(uint32_t a, uint32_t b, unsigned long long c) { return (unsigned long long)a * b + c; }
static inline unsigned long long mul_wide(uint32_t a, uint32_t b) { return (unsigned long long)a * b; }

#endif

}  // namespace ptx
