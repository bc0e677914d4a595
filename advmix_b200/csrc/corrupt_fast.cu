// Perf-mode kernels of the byte-bound corruptions (no injected draws): float32 arithmetic, random draws
// generated in-register, 16 bytes per thread per step (128-bit streaming loads / stores).  Tolerance
// against the float64 parity path / oracle on the dumped draws: <= 1 LSB after the final truncation
// (north_star's bar for floating-point ops); impulse and shot noise are integer decisions and stay exact.
#include "corrupt_common.cuh"

#include <atomic>
#include <cmath>
#include <vector>

namespace advmix {

constexpr int FT_THREADS = 256;

static inline dim3 fast_grid(int64_t work_per_image, int n) {
    int64_t bx = (work_per_image + FT_THREADS - 1) / FT_THREADS;
    int64_t cap = std::max<int64_t>(1, ((int64_t)sm_count() * 8 + n - 1) / n);
    return dim3((unsigned)std::min(bx, cap), (unsigned)n);
}

// float(byte k of w) without the conversion pipe
__device__ __forceinline__ float byte_f(uint32_t w, int k) { return byte_of_word_f(w, k); }
// clamp to [0,255] and truncate toward zero -> integer in the low byte (round-toward-zero add of 2^23)
__device__ __forceinline__ uint32_t trunc255(float v) {
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return __float_as_uint(__fadd_rz(v, 8388608.0f)) & 255u;
}
__device__ __forceinline__ uint32_t pack4u(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return a | (b << 8) | (c << 16) | (d << 24); }

// ---- gaussian_noise: x + 255*c*N(0,1) ----------------------------------------------------------------
__global__ void __launch_bounds__(FT_THREADS)
gaussian_noise_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                           uint64_t seed, int64_t sample_base, int64_t n16, float c255) {
    const int slot = slot_of(idx, blockIdx.y);
    const SampleRng rng(seed, sample_base + slot);
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n16;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)slot * n16;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n16; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 v = ld_stream_u4(src + q);
        float na[8], nb[8];
        noise_normal8(rng, TAG_FIELD0, 2 * q, na);
        noise_normal8(rng, TAG_FIELD0, 2 * q + 1, nb);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* n = j < 2 ? na + 4 * j : nb + 4 * (j - 2);
            o[j] = pack4u(trunc255(fmaf(n[0], c255, byte_f(w[j], 0))), trunc255(fmaf(n[1], c255, byte_f(w[j], 1))),
                          trunc255(fmaf(n[2], c255, byte_f(w[j], 2))), trunc255(fmaf(n[3], c255, byte_f(w[j], 3))));
        }
        st_stream_u4(dst + q, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// ---- impulse_noise: 15-bit flip draw + 1 salt bit per value --------------------------------------------------
__global__ void __launch_bounds__(FT_THREADS)
impulse_noise_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                          uint64_t seed, int64_t sample_base, int64_t n16, uint32_t thr15) {
    const int slot = slot_of(idx, blockIdx.y);
    const SampleRng rng(seed, sample_base + slot);
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n16;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)slot * n16;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n16; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 v = ld_stream_u4(src + q);
        uint32_t ka[8], kb[8];
        noise_bits8(rng, TAG_FIELD0, 2 * q, ka);
        noise_bits8(rng, TAG_FIELD0, 2 * q + 1, kb);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t* k = j < 2 ? ka + 4 * j : kb + 4 * (j - 2);
            uint32_t r = w[j];
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if ((k[b] & 0x7FFFu) < thr15) r = (r & ~(255u << (8 * b))) | ((k[b] & 0x8000u) ? (255u << (8 * b)) : 0u);
            o[j] = r;
        }
        st_stream_u4(dst + q, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// ---- shot_noise: inverse-CDF Poisson from integer tables in shared memory (exact, both modes) ------------
// Row v (pixel value) holds T[k] = ceil(cdf_v(k) * 2^24) for k in [kmin, kmin+len); a uniform u = m/2^24
// gives the count  k = kmin + #{T < = m}  (u < cdf  <=>  m < T on the 24-bit grid numpy and we draw from).
struct PoissonTables {
    std::vector<uint32_t> T;        // compact rows
    std::vector<uint32_t> meta;     // per v: offset | kmin << 16 | len << 24  (offset < 65536)
    std::vector<uint8_t> kout;      // output byte for count k
};

static PoissonTables build_poisson(double c) {
    PoissonTables P;
    P.meta.resize(256);
    P.kout.resize(128);
    for (int k = 0; k < 128; ++k) {
        double v = (double)k / c;
        v = std::min(std::max(v, 0.0), 1.0) * 255.0;
        P.kout[k] = (uint8_t)(int)v;
    }
    for (int v = 0; v < 256; ++v) {
        const double lam = ((double)v / 255.0) * c;
        double p = std::exp(-lam), acc = p;
        uint32_t row[128];
        for (int k = 0; k < 128; ++k) {
            if (k) { p = p * lam / k; acc = acc + p; }
            const double t = std::ceil(acc * 16777216.0);
            row[k] = (uint32_t)std::min(t, 16777216.0);
        }
        int kmin = 0;
        while (kmin < 127 && row[kmin] == 0) ++kmin;
        int kend = kmin;                                   // first k with T == 2^24 (always decides) or 127
        while (kend < 127 && row[kend] < 16777216u) ++kend;
        const int len = kend - kmin + 1;
        P.meta[v] = (uint32_t)P.T.size() | ((uint32_t)kmin << 16) | ((uint32_t)len << 24);
        for (int k = kmin; k <= kend; ++k) P.T.push_back(row[k]);
    }
    return P;
}

__global__ void __launch_bounds__(FT_THREADS)
shot_noise_table_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                        const float* __restrict__ field, size_t field_stride, uint64_t seed, int64_t sample_base,
                        int64_t n16, const uint32_t* __restrict__ gT, int nT, const uint32_t* __restrict__ gmeta,
                        const uint8_t* __restrict__ gkout) {
    extern __shared__ uint32_t s_T[];       // nT entries, then 256 meta, then 128 bytes kout
    uint32_t* s_meta = s_T + nT;
    uint8_t* s_kout = reinterpret_cast<uint8_t*>(s_meta + 256);
    for (int i = threadIdx.x; i < nT; i += FT_THREADS) s_T[i] = gT[i];
    for (int i = threadIdx.x; i < 256; i += FT_THREADS) s_meta[i] = gmeta[i];
    if (threadIdx.x < 128) s_kout[threadIdx.x] = gkout[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const SampleRng rng(seed, sample_base + slot);
    const float* inj = field ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(field) + (size_t)i * field_stride) : nullptr;
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n16;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)slot * n16;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n16; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 v = ld_stream_u4(src + q);
        uint32_t m[16];
        if (inj) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 u = *reinterpret_cast<const float4*>(inj + 16 * q + 4 * j);
                m[4 * j] = (uint32_t)(u.x * 16777216.0f); m[4 * j + 1] = (uint32_t)(u.y * 16777216.0f);
                m[4 * j + 2] = (uint32_t)(u.z * 16777216.0f); m[4 * j + 3] = (uint32_t)(u.w * 16777216.0f);
            }
        } else {
            uint32_t ka[8], kb[8];
            noise_bits8(rng, TAG_FIELD0, 2 * q, ka);
            noise_bits8(rng, TAG_FIELD0, 2 * q + 1, kb);
#pragma unroll
            for (int j = 0; j < 8; ++j) { m[j] = ka[j] << 8; m[8 + j] = kb[j] << 8; }
        }
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t r = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t meta = s_meta[(w[j] >> (8 * b)) & 255u];
                const uint32_t* row = s_T + (meta & 0xFFFFu);
                const int len = (int)(meta >> 24);
                const uint32_t mm = m[4 * j + b];
                int lo = 0, hi = len;                    // count of entries <= mm  (rows are non-decreasing)
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int mid = (lo + hi) >> 1;
                    if (lo < hi) { if (row[mid] <= mm) lo = mid + 1; else hi = mid; }
                }
                const int k = min((int)((meta >> 16) & 255u) + lo, 127);
                r |= (uint32_t)s_kout[k] << (8 * b);
            }
            o[j] = r;
        }
        st_stream_u4(dst + q, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// Perf mode (in-register draws): the uniform has 16 bits (m = k16 << 8), so the thresholds reduce to 16 bits exactly:
// T24 <= k16 * 256  <=>  ceil(T24 / 256) <= k16  <=>  ceil(T24 / 256) - 1 < k16.  Rows are padded to a power of two with a
// sentinel that is never below a draw, which makes the search branch-free with a fixed number of steps: 4 instructions per
// step instead of ~10 for the guarded bisection above (ncu: 115 instructions per byte, 70 % issue-active).  Same counts as
// shot_noise_table_kernel on the dumped draws (tests: perf == injected bit for bit).
template <int STEPS>
__global__ void __launch_bounds__(FT_THREADS)
shot_noise_perf16_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx, uint64_t seed,
                         int64_t sample_base, int64_t n16, const uint16_t* __restrict__ gT, const uint8_t* __restrict__ gkmin,
                         const uint8_t* __restrict__ gkout) {
    constexpr int L = 1 << STEPS, LP = L + 2;            // row pitch L + 2: with a pitch of L (a multiple of 64 words) element k of EVERY row sits in
                                                         // the same bank and the first steps (all lanes at k = L/2 - 1) are 32-way conflicts
    extern __shared__ __align__(16) uint16_t s_T16[];     // [256][LP], then kmin[256], kout[128]
    uint8_t* s_kmin = reinterpret_cast<uint8_t*>(s_T16 + 256 * LP);
    uint8_t* s_kout = s_kmin + 256;
    {
        const uint4* g4 = reinterpret_cast<const uint4*>(gT);
        uint4* s4 = reinterpret_cast<uint4*>(s_T16);
        for (int i = threadIdx.x; i < 256 * LP / 8; i += FT_THREADS) s4[i] = g4[i];
    }
    for (int i = threadIdx.x; i < 256; i += FT_THREADS) s_kmin[i] = gkmin[i];
    if (threadIdx.x < 128) s_kout[threadIdx.x] = gkout[threadIdx.x];
    __syncthreads();
    const int slot = slot_of(idx, blockIdx.y);
    const SampleRng rng(seed, sample_base + slot);
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n16;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)slot * n16;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n16; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 v = ld_stream_u4(src + q);
        uint32_t m[16];
        {
            uint32_t ka[8], kb[8];
            noise_bits8(rng, TAG_FIELD0, 2 * q, ka);
            noise_bits8(rng, TAG_FIELD0, 2 * q + 1, kb);
#pragma unroll
            for (int j = 0; j < 8; ++j) { m[j] = ka[j]; m[8 + j] = kb[j]; }
        }
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        // the 16 searches of a thread advance in lock step (step-major loops): a search is a chain of STEPS dependent
        // shared-memory reads, and one chain at a time left the kernel latency-bound (2x slower than the bisection it replaces)
        uint32_t pos[16], kmin[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const uint32_t px = __byte_perm(w[e >> 2], 0u, 0x4440u | (e & 3));
            pos[e] = px * LP;                           // element index of the row start; the count accumulates on top
            kmin[e] = s_kmin[px];
        }
#pragma unroll
        for (int st = STEPS - 1; st >= 0; --st) {
#pragma unroll
            for (int e = 0; e < 16; ++e) pos[e] += (s_T16[pos[e] + (1u << st) - 1u] < m[e]) ? (1u << st) : 0u;
        }
        uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const uint32_t px = __byte_perm(w[e >> 2], 0u, 0x4440u | (e & 3));
            const uint32_t lo = pos[e] - px * LP;
            o[e >> 2] |= (uint32_t)s_kout[min(kmin[e] + lo, 127u)] << (8 * (e & 3));
        }
        st_stream_u4(dst + q, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// ---- contrast: (x - mean)*c + mean ------------------------------------------------------------------------
// 48 bytes (16 pixels) per thread step, so the channel of every byte is a compile-time constant.
__device__ __forceinline__ uint32_t sum_bytes(uint32_t w, uint32_t mask) {   // sum of the bytes selected by mask bits 0..3
    uint32_t s = 0;
    if (mask & 1) s += w & 255u;
    if (mask & 2) s += (w >> 8) & 255u;
    if (mask & 4) s += (w >> 16) & 255u;
    if (mask & 8) s += w >> 24;
    return s;
}

__global__ void __launch_bounds__(FT_THREADS)
channel_sum48_kernel(const uint8_t* __restrict__ in, const int32_t* __restrict__ idx, int64_t n48,
                     unsigned long long* __restrict__ sums) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n48 * 3;
    uint32_t s0 = 0, s1 = 0, s2 = 0;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n48; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 a = ld_stream_u4(src + 3 * q), b = ld_stream_u4(src + 3 * q + 1), c = ld_stream_u4(src + 3 * q + 2);
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            // byte k of word j is element 4j+k, channel (4j+k) % 3
            uint32_t m0 = 0, m1 = 0, m2 = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ch = (4 * j + k) % 3;
                if (ch == 0) m0 |= 1u << k; else if (ch == 1) m1 |= 1u << k; else m2 |= 1u << k;
            }
            s0 += sum_bytes(w[j], m0); s1 += sum_bytes(w[j], m1); s2 += sum_bytes(w[j], m2);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[3 * i], (unsigned long long)s0);
        atomicAdd(&sums[3 * i + 1], (unsigned long long)s1);
        atomicAdd(&sums[3 * i + 2], (unsigned long long)s2);
    }
}

__global__ void __launch_bounds__(FT_THREADS)
contrast_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ idx,
                     int64_t n48, const unsigned long long* __restrict__ sums, float inv_npix, float c) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    // value domain 0..255: out = x*c + mean255*(1-c)
    const float add[3] = {(float)sums[3 * i] * inv_npix * (1.0f - c), (float)sums[3 * i + 1] * inv_npix * (1.0f - c),
                          (float)sums[3 * i + 2] * inv_npix * (1.0f - c)};
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n48 * 3;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)slot * n48 * 3;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n48; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 a = ld_stream_u4(src + 3 * q), b = ld_stream_u4(src + 3 * q + 1), cc = ld_stream_u4(src + 3 * q + 2);
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w};
        uint32_t o[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            uint32_t r = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) r |= trunc255(fmaf(byte_f(w[j], k), c, add[(4 * j + k) % 3])) << (8 * k);
            o[j] = r;
        }
        st_stream_u4(dst + 3 * q, make_uint4(o[0], o[1], o[2], o[3]));
        st_stream_u4(dst + 3 * q + 1, make_uint4(o[4], o[5], o[6], o[7]));
        st_stream_u4(dst + 3 * q + 2, make_uint4(o[8], o[9], o[10], o[11]));
    }
}

// ---- severity sweeps (advmix_corrupt_sweep_u8c3): the draws / the loaded words / the channel sums are shared by the five
// severities, each output is the expression of the single-severity kernel above -------------------------------------------
__global__ void __launch_bounds__(FT_THREADS)
gaussian_noise_sweep_fast_kernel(const uint8_t* __restrict__ in, Sweep5Out outs, const int32_t* __restrict__ idx, uint64_t seed,
                                 int64_t sample_base, int64_t n16, Sweep5F c255) {
    const int slot = slot_of(idx, blockIdx.y);
    const SampleRng rng(seed, sample_base + slot);
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n16;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n16; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 v = ld_stream_u4(src + q);
        float na[8], nb[8], nz[16];
        noise_normal8(rng, TAG_FIELD0, 2 * q, na);
        noise_normal8(rng, TAG_FIELD0, 2 * q + 1, nb);
#pragma unroll
        for (int j = 0; j < 8; ++j) { nz[j] = na[j]; nz[8 + j] = nb[j]; }
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = byte_f(w[j >> 2], j & 3);
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const float c = c255.v[s];
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o[j] = pack4u(trunc255(fmaf(nz[4 * j], c, x[4 * j])), trunc255(fmaf(nz[4 * j + 1], c, x[4 * j + 1])),
                              trunc255(fmaf(nz[4 * j + 2], c, x[4 * j + 2])), trunc255(fmaf(nz[4 * j + 3], c, x[4 * j + 3])));
            st_stream_u4(reinterpret_cast<uint4*>(outs.p[s]) + (int64_t)slot * n16 + q, make_uint4(o[0], o[1], o[2], o[3]));
        }
    }
}

__global__ void __launch_bounds__(FT_THREADS)
impulse_noise_sweep_fast_kernel(const uint8_t* __restrict__ in, Sweep5Out outs, const int32_t* __restrict__ idx, uint64_t seed,
                                int64_t sample_base, int64_t n16, Sweep5U thr15) {
    const int slot = slot_of(idx, blockIdx.y);
    const SampleRng rng(seed, sample_base + slot);
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n16;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n16; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 v = ld_stream_u4(src + q);
        uint32_t ka[8], kb[8], k[16];
        noise_bits8(rng, TAG_FIELD0, 2 * q, ka);
        noise_bits8(rng, TAG_FIELD0, 2 * q + 1, kb);
#pragma unroll
        for (int j = 0; j < 8; ++j) { k[j] = ka[j]; k[8 + j] = kb[j]; }
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const uint32_t thr = thr15.v[s];
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t r = w[j];
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if ((k[4 * j + b] & 0x7FFFu) < thr) r = (r & ~(255u << (8 * b))) | ((k[4 * j + b] & 0x8000u) ? (255u << (8 * b)) : 0u);
                o[j] = r;
            }
            st_stream_u4(reinterpret_cast<uint4*>(outs.p[s]) + (int64_t)slot * n16 + q, make_uint4(o[0], o[1], o[2], o[3]));
        }
    }
}

__global__ void __launch_bounds__(FT_THREADS)
contrast_sweep_fast_kernel(const uint8_t* __restrict__ in, Sweep5Out outs, const int32_t* __restrict__ idx, int64_t n48,
                           const unsigned long long* __restrict__ sums, float inv_npix, Sweep5F cs) {
    const int i = blockIdx.y, slot = slot_of(idx, i);
    const float m[3] = {(float)sums[3 * i] * inv_npix, (float)sums[3 * i + 1] * inv_npix, (float)sums[3 * i + 2] * inv_npix};
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)slot * n48 * 3;
    for (int64_t q = (int64_t)blockIdx.x * FT_THREADS + threadIdx.x; q < n48; q += (int64_t)gridDim.x * FT_THREADS) {
        const uint4 a = ld_stream_u4(src + 3 * q), b = ld_stream_u4(src + 3 * q + 1), cc = ld_stream_u4(src + 3 * q + 2);
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w};
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const float c = cs.v[s];
            const float add[3] = {m[0] * (1.0f - c), m[1] * (1.0f - c), m[2] * (1.0f - c)};
            uint32_t o[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                uint32_t r = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) r |= trunc255(fmaf(byte_f(w[j], k), c, add[(4 * j + k) % 3])) << (8 * k);
                o[j] = r;
            }
            uint4* dst = reinterpret_cast<uint4*>(outs.p[s]) + (int64_t)slot * n48 * 3;
            st_stream_u4(dst + 3 * q, make_uint4(o[0], o[1], o[2], o[3]));
            st_stream_u4(dst + 3 * q + 1, make_uint4(o[4], o[5], o[6], o[7]));
            st_stream_u4(dst + 3 * q + 2, make_uint4(o[8], o[9], o[10], o[11]));
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
bool fast48_ok(const CorruptArgs& a) {
    return ((int64_t)a.H * a.W) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
}

bool fast_ok(const CorruptArgs& a) {
    return ((int64_t)a.H * a.W * 3) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
}

int run_gaussian_noise_fast(const CorruptArgs& a) {
    const int64_t n16 = (int64_t)a.H * a.W * 3 / 16;
    gaussian_noise_fast_kernel<<<fast_grid(n16, a.n), FT_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, a.seed, a.sample_base, n16,
                                                                                (float)(sev_gaussian_noise(a.severity) * 255.0));
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_impulse_noise_fast(const CorruptArgs& a) {
    const int64_t n16 = (int64_t)a.H * a.W * 3 / 16;
    // u = k/32768 < c  <=>  k < ceil(c*32768)
    const uint32_t thr = (uint32_t)std::ceil(sev_impulse_noise(a.severity) * 32768.0);
    impulse_noise_fast_kernel<<<fast_grid(n16, a.n), FT_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, a.seed, a.sample_base, n16, thr);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

template <int STEPS>
static int launch_shot16(const CorruptArgs& a, const uint16_t* d_T, const uint8_t* d_kmin, const uint8_t* d_kout) {
    const size_t smem = (size_t)256 * ((1 << STEPS) + 2) * 2 + 256 + 128;
    ADVMIX_CUDA_OK(ensure_dyn_smem(shot_noise_perf16_kernel<STEPS>, (int)smem));
    const int64_t n16 = (int64_t)a.H * a.W * 3 / 16;
    shot_noise_perf16_kernel<STEPS><<<fast_grid(n16, a.n), FT_THREADS, smem, a.stream>>>(a.in, a.out, a.idx, a.seed, a.sample_base, n16, d_T, d_kmin, d_kout);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

struct Poisson16 { std::vector<uint16_t> T16; std::vector<uint8_t> kmin; int steps = 0; };

static Poisson16 build_poisson16(const PoissonTables& P) {
    Poisson16 R;
    // effective row length in 16 bits: entries whose threshold is already 65536 can never be below a draw
    int maxlen = 1;
    std::vector<std::vector<uint16_t>> rows(256);
    std::vector<uint8_t>& kmin = R.kmin;
    kmin.resize(256);
    for (int v = 0; v < 256; ++v) {
        const uint32_t meta = P.meta[v];
        const uint32_t* T = P.T.data() + (meta & 0xFFFFu);
        const int len = (int)(meta >> 24);
        kmin[v] = (uint8_t)((meta >> 16) & 255u);
        for (int k = 0; k < len; ++k) {
            const uint32_t t16 = (T[k] + 255u) >> 8;          // ceil(T24 / 256), >= 1 inside a row
            if (t16 >= 65536u) break;
            rows[v].push_back((uint16_t)(t16 - 1u));
        }
        maxlen = std::max(maxlen, (int)rows[v].size());
    }
    int steps = 1;
    while ((1 << steps) <= maxlen) ++steps;                  // L > maxlen: at least one sentinel per row
    steps = std::max(steps, 5);
    R.steps = steps;
    if (steps > 7) return R;
    const int LP = (1 << steps) + 2;                          // padded row pitch (bank spread), see the kernel
    R.T16.assign((size_t)256 * LP, (uint16_t)0xFFFF);
    for (int v = 0; v < 256; ++v) std::copy(rows[v].begin(), rows[v].end(), R.T16.begin() + (size_t)v * LP);
    return R;
}

// perf mode only (no injected field): 16-bit thresholds, branch-free search
static int run_shot_noise_perf16(const CorruptArgs& a, const PoissonTables& P) {
    static Poisson16 cache[5];
    static std::atomic<int> ready[5];
    static std::atomic<bool> guard{false};
    if (ready[a.severity - 1].load(std::memory_order_acquire) == 0) {
        while (guard.exchange(true)) {}
        if (ready[a.severity - 1].load() == 0) {
            cache[a.severity - 1] = build_poisson16(P);
            ready[a.severity - 1].store(1, std::memory_order_release);
        }
        guard.store(false);
    }
    const Poisson16& R = cache[a.severity - 1];
    const int steps = R.steps;
    if (steps > 7) return -1;
    const std::vector<uint16_t>& T16 = R.T16;
    const std::vector<uint8_t>& kmin = R.kmin;
    const std::string key = "poisson16_" + std::to_string(a.severity);
    const uint16_t* d_T = reinterpret_cast<const uint16_t*>(cached_table(key + "_T", T16.data(), T16.size() * 2));
    const uint8_t* d_kmin = reinterpret_cast<const uint8_t*>(cached_table(key + "_m", kmin.data(), kmin.size()));
    const uint8_t* d_kout = reinterpret_cast<const uint8_t*>(cached_table(key + "_k", P.kout.data(), P.kout.size()));
    if (!d_T || !d_kmin || !d_kout) return ADVMIX_ERR_CUDA;
    switch (steps) {
        case 5: return launch_shot16<5>(a, d_T, d_kmin, d_kout);
        case 6: return launch_shot16<6>(a, d_T, d_kmin, d_kout);
        default: return launch_shot16<7>(a, d_T, d_kmin, d_kout);
    }
}

// The tables are a function of the severity only: built once per process (building them - 256 rows x 128 exp / ceil - costs
// ~0.5 ms of host time, which at 512 images per call was MORE than the kernel: the op was host-bound in round 1).
static const PoissonTables& poisson_tables(int severity) {
    static PoissonTables cache[5];
    static std::atomic<int> ready[5];
    static std::atomic<bool> guard{false};
    if (ready[severity - 1].load(std::memory_order_acquire) == 0) {
        while (guard.exchange(true)) {}
        if (ready[severity - 1].load() == 0) {
            cache[severity - 1] = build_poisson(sev_shot_noise(severity));
            ready[severity - 1].store(1, std::memory_order_release);
        }
        guard.store(false);
    }
    return cache[severity - 1];
}

int run_shot_noise_table(const CorruptArgs& a) {
    const PoissonTables& P = poisson_tables(a.severity);
    if (!a.rand_field) {
        const int rc16 = run_shot_noise_perf16(a, P);
        if (rc16 != -1) return rc16;
    }
    const std::string key = "poisson_int_" + std::to_string(a.severity);
    const uint32_t* d_T = reinterpret_cast<const uint32_t*>(cached_table(key + "_T", P.T.data(), P.T.size() * 4));
    const uint32_t* d_meta = reinterpret_cast<const uint32_t*>(cached_table(key + "_m", P.meta.data(), P.meta.size() * 4));
    const uint8_t* d_kout = reinterpret_cast<const uint8_t*>(cached_table(key + "_k", P.kout.data(), P.kout.size()));
    if (!d_T || !d_meta || !d_kout) return ADVMIX_ERR_CUDA;
    ADVMIX_REQUIRE(P.T.size() < 65536, "shot_noise: table too large");
    const size_t smem = (P.T.size() + 256) * 4 + 128;
    ADVMIX_CUDA_OK(ensure_dyn_smem(shot_noise_table_kernel, 100 * 1024));
    const int64_t n16 = (int64_t)a.H * a.W * 3 / 16;
    shot_noise_table_kernel<<<fast_grid(n16, a.n), FT_THREADS, smem, a.stream>>>(
        a.in, a.out, a.idx, reinterpret_cast<const float*>(a.rand_field), a.field_bytes, a.seed, a.sample_base, n16, d_T,
        (int)P.T.size(), d_meta, d_kout);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_contrast_fast(const CorruptArgs& a, unsigned long long* sums) {
    const double c[5] = {0.4, 0.3, 0.2, 0.1, 0.05};
    const int64_t n48 = (int64_t)a.H * a.W * 3 / 48;
    ADVMIX_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)a.n * 3 * sizeof(unsigned long long), a.stream));
    channel_sum48_kernel<<<fast_grid(n48, a.n), FT_THREADS, 0, a.stream>>>(a.in, a.idx, n48, sums);
    ADVMIX_LAUNCH_OK();
    contrast_fast_kernel<<<fast_grid(n48, a.n), FT_THREADS, 0, a.stream>>>(a.in, a.out, a.idx, n48, sums,
                                                                          (float)(1.0 / ((double)a.H * a.W)), (float)c[a.severity - 1]);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

static bool sweep_outs_aligned(const SweepArgs& sw) {
    uintptr_t m = 0;
    for (int s = 0; s < 5; ++s) m |= reinterpret_cast<uintptr_t>(sw.outs[s]);
    return (m & 15) == 0;
}
static Sweep5Out sweep_outs(const SweepArgs& sw) {
    Sweep5Out o;
    for (int s = 0; s < 5; ++s) o.p[s] = sw.outs[s];
    return o;
}

int run_gaussian_noise_sweep(const SweepArgs& sw) {
    const CorruptArgs& a = sw.base;
    if (!a.fast || ((int64_t)a.H * a.W * 3) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || !sweep_outs_aligned(sw)) return -1;
    const int64_t n16 = (int64_t)a.H * a.W * 3 / 16;
    Sweep5F c;
    for (int s = 0; s < 5; ++s) c.v[s] = (float)(sev_gaussian_noise(s + 1) * 255.0);
    gaussian_noise_sweep_fast_kernel<<<fast_grid(n16, a.n), FT_THREADS, 0, a.stream>>>(a.in, sweep_outs(sw), a.idx, a.seed, a.sample_base, n16, c);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_impulse_noise_sweep(const SweepArgs& sw) {
    const CorruptArgs& a = sw.base;
    if (((int64_t)a.H * a.W * 3) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || !sweep_outs_aligned(sw)) return -1;
    const int64_t n16 = (int64_t)a.H * a.W * 3 / 16;
    Sweep5U t;
    for (int s = 0; s < 5; ++s) t.v[s] = (uint32_t)std::ceil(sev_impulse_noise(s + 1) * 32768.0);
    impulse_noise_sweep_fast_kernel<<<fast_grid(n16, a.n), FT_THREADS, 0, a.stream>>>(a.in, sweep_outs(sw), a.idx, a.seed, a.sample_base, n16, t);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int run_contrast_sweep(const SweepArgs& sw) {
    const CorruptArgs& a = sw.base;
    if (!a.fast || ((int64_t)a.H * a.W) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || !sweep_outs_aligned(sw)) return -1;
    const double c[5] = {0.4, 0.3, 0.2, 0.1, 0.05};
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(a.ws);
    const int64_t n48 = (int64_t)a.H * a.W * 3 / 48;
    ADVMIX_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)a.n * 3 * sizeof(unsigned long long), a.stream));
    channel_sum48_kernel<<<fast_grid(n48, a.n), FT_THREADS, 0, a.stream>>>(a.in, a.idx, n48, sums);
    ADVMIX_LAUNCH_OK();
    Sweep5F cf;
    for (int s = 0; s < 5; ++s) cf.v[s] = (float)c[s];
    contrast_sweep_fast_kernel<<<fast_grid(n48, a.n), FT_THREADS, 0, a.stream>>>(a.in, sweep_outs(sw), a.idx, n48, sums,
                                                                               (float)(1.0 / ((double)a.H * a.W)), cf);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // namespace advmix
