// Shared device/host helpers for libadvmix_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include <cstdio>
#include <cstdarg>

#include "../../include/advmix_b200.h"

namespace advmix {

// ---- error plumbing ------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define ADVMIX_CUDA_OK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return ::advmix::fail(ADVMIX_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__,  \
                                  #expr, cudaGetErrorString(_e));                         \
    } while (0)

#define ADVMIX_LAUNCH_OK()                                                                \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess)                                                            \
            return ::advmix::fail(ADVMIX_ERR_CUDA, "%s:%d launch -> %s", __FILE__,        \
                                  __LINE__, cudaGetErrorString(_e));                      \
    } while (0)

#define ADVMIX_REQUIRE(cond, ...)                                                         \
    do {                                                                                  \
        if (!(cond)) return ::advmix::fail(ADVMIX_ERR_INVALID, __VA_ARGS__);              \
    } while (0)

int sm_count();  // SMs of the current device (cached)

// true exactly once per (current device, tag), under the library mutex: function attributes and __constant__
// symbols are per DEVICE, so one-time set-up must be repeated on every GPU a process touches.
bool first_use_on_device(const void* tag);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel)
template <class K>
static inline cudaError_t ensure_dyn_smem(K kernel, int bytes) {
    if (!first_use_on_device(reinterpret_cast<const void*>(kernel))) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// Device-resident constant tables, built once per device on first use.
// Returns a device pointer valid for the life of the process (nullptr on error).
const void* cached_table(const std::string& key, const void* host, size_t bytes);

void* cached_buffer(const std::string& key, size_t bytes);   // zero-initialised device scratch, once per (device, key)

static inline cudaStream_t as_stream(advmix_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- streaming loads / stores -----------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
    uint4 r = ld_stream_u4(p);
    return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z),
                       __uint_as_float(r.w));
}
__device__ __forceinline__ void st_stream_f4(void* p, const float4& v) {
    st_stream_u4(p, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z),
                               __float_as_uint(v.w)));
}

// ---- Philox4x32-10 counter RNG ------------------------------------------------------
// Stateless: draws for any (sample, element) are recomputable in-register, so the fused
// perf path never reads a random buffer from HBM.
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
        hi = __umulhi(a, b);
        lo = a * b;
#else
        uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
#endif
    }
    __host__ __device__ inline uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, c0, h0, l0);
            mulhilo(0xCD9E8D57u, c2, h1, l1);
            uint32_t n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// U[0,1) with 24 bits: exactly representable in float32 (and float64).
__host__ __device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// Two N(0,1) float32 from two u32 (Box-Muller).  Used by BOTH the fill kernel and the
// fused kernels, so injected and in-register draws are bit-identical.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
    float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);           // [0,1)
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// RNG stream tags (c3 of the Philox counter)
enum RandTag : uint32_t { TAG_FIELD0 = 0x1000, TAG_FIELD1 = 0x2000, TAG_PARAM = 0x3000, TAG_GLASS = 0x4000 };

}  // namespace advmix
