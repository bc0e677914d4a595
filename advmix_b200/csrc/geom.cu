// Row a1 / a7: affine crop (cv2.warpAffine fixed-point semantics), batched
// get_affine_transform, joint flip+affine, ToTensor+Normalize.
//
// Reference: lib/dataset/JointsDataset.py:167-199, :324-332; lib/utils/transforms.py:44-122.
#include "common.cuh"
#include "affine.cuh"

#include <atomic>
#include <climits>
#include <cstdlib>
#include <vector>

#include <cuda.h>            // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time, libcuda is not linked)

namespace advmix {

// ---- tensor-map TMA staging -----------------------------------------------------------------
// A source image is described to the TMA unit as a 2-D tensor of 8-byte elements {pitch / 8, H rows}; a box
// {2 * NCH elements = 16 * NCH bytes, 8 rows} lands in shared memory as 8 dense rows of 16 * NCH bytes - exactly the staged
// layout the blend loop reads (one row pitch per band), with out-of-image rows / columns zero-filled by the hardware.  (8-byte
// elements because a box dimension is limited to 256 elements: rows of up to 1536 bytes must be one box wide.  A 3-D view
// {16 B, chunks, rows} was measured first: correct, but the TMA unit then moves 16 bytes per request - 120 us per launch
// against 82 us for the LDGSTS path.)  Box dimensions are part of a
// tensor map, so every sample gets one map per width class (the band's row pitch is rounded up to the class); the maps
// are derived on the device from ONE host-encoded template with tensormap.replace (address, dims, stride, box width)
// by a small kernel in front of the crop kernel - the per-sample geometry lives in device arrays, the host never sees it.
// The lanes of one consumer warp then issue the band's ceil(rows / 8) cp.async.bulk.tensor copies, one box per lane
// (completion: complete_tx on the stage's mbarrier), instead of all 128 lanes of the group issuing 16-byte LDGSTS copies.
constexpr int TM_NCLS = 14;
constexpr int TM_ROWS = 16;
__constant__ int c_tm_cls[TM_NCLS] = {4, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64, 80, 96};

constexpr int AB_BITS = 10;
constexpr int INTER_BITS = 5;
constexpr int ROUND_DELTA = (1 << AB_BITS) / (1 << INTER_BITS) / 2;  // 16

// cv::warpAffine's inversion of the forward matrix, float64, same operation order.
__device__ __forceinline__ void invert_affine(const double* __restrict__ Min, double* M) {
    double m0 = Min[0], m1 = Min[1], m2 = Min[2], m3 = Min[3], m4 = Min[4], m5 = Min[5];
    double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
    D = (D != 0.0) ? __ddiv_rn(1.0, D) : 0.0;
    double A11 = __dmul_rn(m4, D), A22 = __dmul_rn(m0, D);
    m0 = A11;
    m1 = __dmul_rn(m1, -D);
    m3 = __dmul_rn(m3, -D);
    m4 = A22;
    double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
    double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
    M[0] = m0; M[1] = m1; M[2] = b1; M[3] = m3; M[4] = m4; M[5] = b2;
}

struct WarpArgs {
    const uint8_t* src_base;
    const int64_t* src_off;
    const int32_t* src_h;
    const int32_t* src_w;
    const int64_t* src_pitch;
    const uint8_t* flip;
    const double* M;
    uint8_t* dst_u8;
    void* dst_norm;
    const float* lut;
    int dw, dh, norm_dtype;
    // optional: evaluate get_affine_transform inside the kernel (M == nullptr)
    const float* center;
    const double* scale;
    const double* rot;
    int scale_f32;
    // dynamic tile scheduling (WS_CFG_DYNAMIC): global tile counter, zeroed before the launch
    unsigned int* tile_counter;
    const CUtensorMap* tmaps;       // [B][TM_NCLS]; nullptr: LDGSTS staging (the default; ADVMIX_WARP_TMA=1 selects the TMA path)
    int tma_flags;                  // experiment knobs (ADVMIX_TMA_FLAGS): 1 = skip the tensormap-proxy acquire fence
};

// ---- persistent, warp-specialised tile kernel: planner warps + a multi-stage cp.async pipeline ---------
// The destination is cut into 32x32 tiles.  A tile maps to a rotated rectangle in the source.
// PLANNER warps compute, several tiles ahead, each tile's integer source box (taps included) with
// OpenCV's fixed-point formulas, split it into row bands that fit a pipeline stage, and publish the
// descriptors through an mbarrier ring.  The 8 CONSUMER warps run a 4-stage cp.async pipeline over
// the bands: all 256 threads issue the 16-byte copies of the box of band n+3 (raw HWC bytes; the
// completion arrives on full[stage]), then blend the four taps of each pixel of band n from two
// 6-byte shared-memory windows and write the normalised / uint8 outputs.  DRAM latency is hidden by
// the pipeline depth, every footprint byte crosses HBM about once (overlap between neighbouring
// tiles hits L2).  Box parts outside the image are never copied: the consumers zero those taps,
// which IS cv2's BORDER_CONSTANT(0).  Boxes that do not fit even as 8-row bands, and sources whose
// rows are not 16-byte aligned, are sampled from global memory directly (same arithmetic).
// tile = 32 columns (one per lane) x WT_TH rows.  WT_TH = 64 (with the tiles of the last eighth of the samples 32 rows high, see the
// planners) halves the per-item hand-off cost per pixel and is 2 % faster when consecutive launches overlap on two streams (72.0
// instead of 73.5 us per bench step at the end of round 2), but a launch measured alone is slower (69.1 instead of 67.6 us: fewer,
// longer items per consumer group); the roofline number is the kernel alone, so 32 stays the default.
constexpr int WT_TW = 32, WT_TH = 32;
static_assert(WT_TH == 32 || WT_TH == 64, "the planners compute one or two rows per lane");
constexpr int WS_GROUPS = 2, WS_GROUP_WARPS = 4, WS_GSTAGES = 2;     // consumer groups, warps per group, pipeline stages per group
constexpr int WS_STAGES = WS_GROUPS * WS_GSTAGES, WS_DESC = 8, WS_CONSUMER_WARPS = WS_GROUPS * WS_GROUP_WARPS, WS_PLANNER_WARPS = 4;
constexpr int WS_THREADS = (WS_CONSUMER_WARPS + WS_PLANNER_WARPS) * 32;
constexpr int WS_STAGE_BYTES = 25 * 1024 + 512, WS_STAGE_ALLOC = WS_STAGE_BYTES + 128;   // 2 CTAs x (4 stages + 9.4 KB static + 1 KB reserved) just fit the 227 KB of an SM (26368 B per stage: one CTA per SM, 80.9 us)
constexpr int WS_DESC_LOG2 = 3, WS_GSTAGES_LOG2 = 1;          // the cursors are non-negative: masks and shifts instead of signed % and /
static_assert(WS_DESC == 1 << WS_DESC_LOG2 && WS_GSTAGES == 1 << WS_GSTAGES_LOG2 && (WS_PLANNER_WARPS & (WS_PLANNER_WARPS - 1)) == 0, "power-of-two ring sizes");
enum { WS_MODE_STAGED = 0, WS_MODE_BORDER = 1, WS_MODE_DIRECT = 2, WS_MODE_DONE = 3 };


__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory loads by 32-bit shared address; plain C++ loads, so the compiler interleaves the four
// pixels of a step, while the asm-volatile mbarrier waits (memory clobber) keep them behind the data.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    return *reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic((size_t)addr));
}
__device__ __forceinline__ float ldsf(uint32_t addr) {
    return *reinterpret_cast<const float*>(__cvta_shared_to_generic((size_t)addr));
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t tx) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(tx) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    // try_wait suspends the thread in hardware until the phase completes or the hint (ns) elapses, so a
    // waiting warp does not burn issue slots that the blending warps of the same SM partition need
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity), "r"(20000u) : "memory");
    } while (!ok);
}
// planner-side wait: back off between polls so that a planner waiting for a free descriptor slot (a long wait: the
// consumers set the pace) does not compete with the blending warps for issue slots
#ifndef WS_CFG_PLAN_SLEEP
#define WS_CFG_PLAN_SLEEP 0
#endif
#ifndef WS_CFG_PLANNER_LOW
#define WS_CFG_PLANNER_LOW 0
#endif
// Tiles cost between one and four pipeline items (rotated / down-scaled samples need more bands), so a static
// round-robin split leaves some CTAs with 10-20 % more work than the average and the kernel ends with the slowest one.
// With WS_CFG_DYNAMIC the planners claim tiles from a global counter instead.
#ifndef WS_CFG_DYNAMIC
#define WS_CFG_DYNAMIC 1
#endif
// timing experiment only (results are wrong): the consumers skip the copy plan and the copy loop and only arrive on the stage barrier
// the last 1 / WS_CFG_TAIL_DIV of the samples are cut into half-height tiles (dynamic scheduling: a group idles for at most one item at
// the end of the launch, so the items handed out last should be short)
#ifndef WS_CFG_TAIL_DIV
#define WS_CFG_TAIL_DIV 0      // measured with 32-row tiles + a 16-row tail (1/16 of the samples): 67.9 instead of 67.6 us - off
#endif
#ifndef WS_CFG_KNOCKOUT_COPY
#define WS_CFG_KNOCKOUT_COPY 0
#endif
__device__ __forceinline__ void mbar_wait_backoff(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    for (;;) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(parity), "r"(20000u) : "memory");
        if (ok) break;
        if (WS_CFG_PLAN_SLEEP > 0) __nanosleep(WS_CFG_PLAN_SLEEP);
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

__device__ __forceinline__ int floor16(int v) { return v & ~15; }
__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

// two horizontally adjacent source pixels (6 raw bytes at shared address `a`) -> R|B<<16 and G of each
__device__ __forceinline__ void load_pair(uint32_t a, uint32_t& rbL, uint32_t& gL, uint32_t& rbR, uint32_t& gR) {
    const uint32_t aw = a & ~3u;
    const uint32_t w0 = lds32(aw), w1 = lds32(aw + 4), w2 = lds32(aw + 8);
    const uint32_t sh = (a & 3u) << 3;
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    rbL = lo & 0x00FF00FFu;
    gL = __byte_perm(lo, 0, 0x4441);
    rbR = __byte_perm(lo, hi, 0x4543) & 0x00FF00FFu;
    gR = hi & 0xFFu;
}

// per-thread output cursor of one ring item (advance by WS_GROUP_WARPS rows per step).  The three colour planes of the
// normalised output sit `plane_b` bytes apart; for the configured output sizes (template DW x DH) both the row step and the
// plane distance are compile-time constants, so inside the 4-pixel unrolled loop every store address is base + immediate -
// with three separately advanced 64-bit pointers the loop spent ~8 of its 68 instructions per pixel on pointer arithmetic.
template <bool HAS_U8, bool HAS_NORM, bool BF16, int DW, int DH>
struct WarpOut {
    uint8_t* u8;
    char* n0;
    int64_t u8_step, n_step, plane_rt;
    __device__ __forceinline__ int64_t plane_b() const { return DW > 0 ? (int64_t)DW * DH * (BF16 ? 2 : 4) : plane_rt; }
    __device__ __forceinline__ void init(const WarpArgs& a, int64_t sample_off, int64_t pix, int64_t plane, int rows_step) {
        const int dw = DW > 0 ? DW : a.dw;
        if (HAS_U8) { u8 = a.dst_u8 + (sample_off + pix) * 3; u8_step = (int64_t)rows_step * dw * 3; }
        if (HAS_NORM) {
            const int es = BF16 ? 2 : 4;
            n0 = reinterpret_cast<char*>(a.dst_norm) + (3 * sample_off + pix) * es;
            plane_rt = plane * es;
            n_step = (int64_t)rows_step * dw * es;
        }
    }
    __device__ __forceinline__ void advance() {
        if (HAS_U8) u8 += u8_step;
        if (HAS_NORM) n0 += n_step;
    }
};

// blend + store of one destination pixel from the four (R|B<<16, G) tap pairs
template <bool HAS_U8, bool HAS_NORM, bool BF16, int DW, int DH>
__device__ __forceinline__ void warp_emit(const WarpOut<HAS_U8, HAS_NORM, BF16, DW, DH>& out, uint32_t rbA, uint32_t gA, uint32_t rbB,
                                          uint32_t gB, uint32_t rbC, uint32_t gC, uint32_t rbD, uint32_t gD, uint32_t wl,
                                          uint32_t wr, uint32_t fy, uint32_t lut32) {
    // sum_k p_k*w_k with w = a*b*32 == 32*[(32-fy)*(wl*pL + wr*pR)_top + fy*(...)_bottom]  (exact integer
    // identity) => out = (S + 512) >> 10.  R and B share a register (16-bit lanes).
    const uint32_t ify = 32u - fy;
    const uint32_t trb = wl * rbA + wr * rbB, brb = wl * rbC + wr * rbD;
    const uint32_t tg = wl * gA + wr * gB, bg = wl * gC + wr * gD;
    const uint32_t v0 = (ify * (trb & 0xFFFFu) + fy * (brb & 0xFFFFu) + 512u) >> 10;
    const uint32_t v1 = (ify * tg + fy * bg + 512u) >> 10;
    const uint32_t v2 = (ify * (trb >> 16) + fy * (brb >> 16) + 512u) >> 10;
    if (HAS_U8) { out.u8[0] = (uint8_t)v0; out.u8[1] = (uint8_t)v1; out.u8[2] = (uint8_t)v2; }
    if (HAS_NORM) {
        const float f0 = ldsf(lut32 + v0 * 4u), f1 = ldsf(lut32 + 1024u + v1 * 4u), f2 = ldsf(lut32 + 2048u + v2 * 4u);
        if (!BF16) {
            *reinterpret_cast<float*>(out.n0) = f0; *reinterpret_cast<float*>(out.n0 + out.plane_b()) = f1; *reinterpret_cast<float*>(out.n0 + 2 * out.plane_b()) = f2;
        } else {
            *reinterpret_cast<__nv_bfloat16*>(out.n0) = __float2bfloat16_rn(f0);
            *reinterpret_cast<__nv_bfloat16*>(out.n0 + out.plane_b()) = __float2bfloat16_rn(f1);
            *reinterpret_cast<__nv_bfloat16*>(out.n0 + 2 * out.plane_b()) = __float2bfloat16_rn(f2);
        }
    }
}

// Same blend with the TOP and BOTTOM tap of a channel packed in the 16-bit lanes of one register (xy = top | bottom << 16):
// the horizontal step is one IMUL + one IMAD per channel for both rows (lanes stay below 2^13), the vertical step
// ify * top + fy * bottom + 512 is one 2-way dot product (dp2a: two 16-bit lanes x two bytes) - 4 instructions per channel
// instead of ~8.  The integers are the same as in warp_emit.
template <bool HAS_U8, bool HAS_NORM, bool BF16, int DW, int DH>
__device__ __forceinline__ void warp_emit_tb(const WarpOut<HAS_U8, HAS_NORM, BF16, DW, DH>& out, uint32_t rL, uint32_t gL, uint32_t bL,
                                             uint32_t rR, uint32_t gR, uint32_t bR, uint32_t wl, uint32_t wr, uint32_t fyw, uint32_t lut32) {
    const uint32_t v0 = __dp2a_lo(wl * rL + wr * rR, fyw, 512u) >> 10;
    const uint32_t v1 = __dp2a_lo(wl * gL + wr * gR, fyw, 512u) >> 10;
    const uint32_t v2 = __dp2a_lo(wl * bL + wr * bR, fyw, 512u) >> 10;
    if (HAS_U8) { out.u8[0] = (uint8_t)v0; out.u8[1] = (uint8_t)v1; out.u8[2] = (uint8_t)v2; }
    if (HAS_NORM) {
        const float f0 = ldsf(lut32 + v0 * 4u), f1 = ldsf(lut32 + 1024u + v1 * 4u), f2 = ldsf(lut32 + 2048u + v2 * 4u);
        if (!BF16) {
            *reinterpret_cast<float*>(out.n0) = f0; *reinterpret_cast<float*>(out.n0 + out.plane_b()) = f1; *reinterpret_cast<float*>(out.n0 + 2 * out.plane_b()) = f2;
        } else {
            *reinterpret_cast<__nv_bfloat16*>(out.n0) = __float2bfloat16_rn(f0);
            *reinterpret_cast<__nv_bfloat16*>(out.n0 + out.plane_b()) = __float2bfloat16_rn(f1);
            *reinterpret_cast<__nv_bfloat16*>(out.n0 + 2 * out.plane_b()) = __float2bfloat16_rn(f2);
        }
    }
}

template <bool HAS_U8, bool HAS_NORM, bool BF16, int DW, int DH>
__device__ __forceinline__ void warp_store(const WarpOut<HAS_U8, HAS_NORM, BF16, DW, DH>& out, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t lut32) {
    if (HAS_U8) { out.u8[0] = (uint8_t)v0; out.u8[1] = (uint8_t)v1; out.u8[2] = (uint8_t)v2; }
    if (HAS_NORM) {
        const float f0 = ldsf(lut32 + v0 * 4u), f1 = ldsf(lut32 + 1024u + v1 * 4u), f2 = ldsf(lut32 + 2048u + v2 * 4u);
        if (!BF16) {
            *reinterpret_cast<float*>(out.n0) = f0; *reinterpret_cast<float*>(out.n0 + out.plane_b()) = f1; *reinterpret_cast<float*>(out.n0 + 2 * out.plane_b()) = f2;
        } else {
            *reinterpret_cast<__nv_bfloat16*>(out.n0) = __float2bfloat16_rn(f0);
            *reinterpret_cast<__nv_bfloat16*>(out.n0 + out.plane_b()) = __float2bfloat16_rn(f1);
            *reinterpret_cast<__nv_bfloat16*>(out.n0 + 2 * out.plane_b()) = __float2bfloat16_rn(f2);
        }
    }
}

// one destination pixel of a staged item.  K folds the stage base, the box origin and (when flipped)
// the mirror constant: staged byte address of the LEFT source pixel of the tap pair = K + sy*rowpitch + csgn*sx
template <bool HAS_U8, bool HAS_NORM, bool BF16, bool FLIP, bool BORDER, int DW, int DH>
__device__ __forceinline__ void warp_pixel_staged(const WarpOut<HAS_U8, HAS_NORM, BF16, DW, DH>& out, int X0r, int Y0r, int ad, int bd,
                                                  uint32_t K, int rowpitch, int H, int W, uint32_t lut32) {
    const int X = (X0r + ad) >> (AB_BITS - INTER_BITS), Y = (Y0r + bd) >> (AB_BITS - INTER_BITS);
    const uint32_t fx = X & 31, fy = Y & 31;
    const int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
    const uint32_t o = K + (uint32_t)(sy * rowpitch + (FLIP ? -3 : 3) * sx);
    const uint32_t aw = o & ~3u, sh = (o << 3) & 24u;
    const uint32_t t0 = lds32(aw), t1 = lds32(aw + 4), t2 = lds32(aw + 8);
    const uint32_t aw2 = aw + (uint32_t)rowpitch;                       // rowpitch is a multiple of 16
    const uint32_t u0 = lds32(aw2), u1 = lds32(aw2 + 4), u2 = lds32(aw2 + 8);
    const uint32_t tlo = __funnelshift_r(t0, t1, sh), thi = __funnelshift_r(t1, t2, sh);
    const uint32_t ulo = __funnelshift_r(u0, u1, sh), uhi = __funnelshift_r(u1, u2, sh);
    if (!BORDER) {
        // bytes: tlo = [R_A G_A B_A R_B], thi = [G_B B_B . .] (top row, A = left pixel, B = right), ulo / uhi the same for the
        // bottom row (C, D).  Per channel ONE register holds the four taps [top-left, top-right, bottom-left, bottom-right]
        // (5 PRMT for the three channels) and the blend is two 2-way dot products of those bytes with the 16-bit weight pairs
        // (wl*ify | wr*ify << 16) and (wl*fy | wr*fy << 16): S = sum of the four weight-tap products + 512, the same integers as
        // 32*[(32-fy)(wl*pL + wr*pR)_top + fy(...)_bottom]/32 of warp_emit.  18 instead of 25 instructions for unpack + blend.
        const uint32_t tR = __byte_perm(tlo, ulo, 0x7430);                       // [R_A R_B R_C R_D]
        const uint32_t tgb = __byte_perm(tlo, thi, 0x5241), ugb = __byte_perm(ulo, uhi, 0x5241);   // [G_A G_B B_A B_B], [G_C G_D B_C B_D]
        const uint32_t tG = __byte_perm(tgb, ugb, 0x5410), tB = __byte_perm(tgb, ugb, 0x7632);
        const uint32_t wlr = FLIP ? (fx | ((32u - fx) << 16)) : ((32u - fx) | (fx << 16));
        const uint32_t aT = wlr * (32u - fy), aB = wlr * fy;
        const uint32_t v0 = __dp2a_hi(aB, tR, __dp2a_lo(aT, tR, 512u)) >> 10;
        const uint32_t v1 = __dp2a_hi(aB, tG, __dp2a_lo(aT, tG, 512u)) >> 10;
        const uint32_t v2 = __dp2a_hi(aB, tB, __dp2a_lo(aT, tB, 512u)) >> 10;
        warp_store<HAS_U8, HAS_NORM, BF16, DW, DH>(out, v0, v1, v2, lut32);
        return;
    }
    uint32_t rbA = tlo & 0x00FF00FFu, gA = __byte_perm(tlo, 0, 0x4441);
    uint32_t rbB = __byte_perm(tlo, thi, 0x4543) & 0x00FF00FFu, gB = thi & 0xFFu;
    uint32_t rbC = ulo & 0x00FF00FFu, gC = __byte_perm(ulo, 0, 0x4441);
    uint32_t rbD = __byte_perm(ulo, uhi, 0x4543) & 0x00FF00FFu, gD = uhi & 0xFFu;
    if (BORDER) {
        const int cl = FLIP ? (W - 2 - sx) : sx;
        const bool t_ok = (unsigned)sy < (unsigned)H, b_ok = (unsigned)(sy + 1) < (unsigned)H;
        const bool l_ok = (unsigned)cl < (unsigned)W, r_ok = (unsigned)(cl + 1) < (unsigned)W;
        if (!(t_ok && l_ok)) { rbA = 0u; gA = 0u; }
        if (!(t_ok && r_ok)) { rbB = 0u; gB = 0u; }
        if (!(b_ok && l_ok)) { rbC = 0u; gC = 0u; }
        if (!(b_ok && r_ok)) { rbD = 0u; gD = 0u; }
    }
    // weights of the left / right SOURCE pixel (mirrored view: the left source pixel is view column sx+1)
    const uint32_t wl = FLIP ? fx : 32u - fx, wr = FLIP ? 32u - fx : fx;
    warp_emit<HAS_U8, HAS_NORM, BF16, DW, DH>(out, rbA, gA, rbB, gB, rbC, gC, rbD, gD, wl, wr, fy, lut32);
}

template <bool HAS_U8, bool HAS_NORM, bool BF16, bool FLIP, bool BORDER, int DW, int DH>
__device__ __forceinline__ void warp_rows_staged(WarpOut<HAS_U8, HAS_NORM, BF16, DW, DH>& out, uint32_t x0s, uint32_t y0s, int r0, int r1,
                                                 int ad, int bd, uint32_t K, int rowpitch, int H, int W, uint32_t lut32) {
    int r = r0;
    // four independent pixels per step, unrolled for instruction-level parallelism
    for (; r + 3 * WS_GROUP_WARPS < r1; r += 4 * WS_GROUP_WARPS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rr = r + k * WS_GROUP_WARPS;
            warp_pixel_staged<HAS_U8, HAS_NORM, BF16, FLIP, BORDER, DW, DH>(out, (int)lds32(x0s + 4u * rr), (int)lds32(y0s + 4u * rr), ad, bd, K,
                                                                    rowpitch, H, W, lut32);
            out.advance();
        }
    }
    {
        for (; r < r1; r += WS_GROUP_WARPS) {
            warp_pixel_staged<HAS_U8, HAS_NORM, BF16, FLIP, BORDER, DW, DH>(out, (int)lds32(x0s + 4u * r), (int)lds32(y0s + 4u * r), ad, bd, K,
                                                                    rowpitch, H, W, lut32);
            out.advance();
        }
    }
}

// ---- ring structures ------------------------------------------------------------------------------------
struct WarpBand {               // one pipeline item: rows [r0, r1) of a tile.  The planner leaves the copy plan and the tap-address
                                // constant ready to use: what four consumer warps would each recompute per item is computed once.
    int r0, r1;                 // rows of the tile
    int rowpitch, mode;         // staged bytes per row; WS_MODE_*
    int nb16, cshift;           // copy plan: 16-byte chunks per row; lanes per row = 1 << cshift
    int nrows, d0;              // staged rows that lie inside the image; stage offset of the first copied byte
    const uint8_t* g0;          // global address of the first copied byte
    int koff;                   // staged byte address of source pixel (sx, sy) = stage base + koff + sy*rowpitch + (flip ? -3 : 3)*sx
    int by0, A0;                // first staged source row, source byte offset of staged byte 0 (TMA path)
    int cls, ncopy;             // TMA: width class of the band's tensor map, 8-row boxes to copy
    int pad_;
};
struct WarpTileDesc {           // written by a planner warp, read by all consumer threads
    int b, x0, y0, H, W, flip, nbands;    // nbands == 0: end of this CTA's tile list
    int pad;
    long long pitch;
    const uint8_t* src;
    WarpBand band[WT_TH / 8];   // 64 / 32 / 16 / 8-row bands
    int ad[WT_TW], bd[WT_TW], X0[WT_TH], Y0[WT_TH];
};

__device__ __forceinline__ void group_bar(int g) {   // the WS_GROUP_WARPS*32 threads of consumer group g only
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(WS_GROUP_WARPS * 32) : "memory");
}

template <bool HAS_U8, bool HAS_NORM, bool BF16, bool TMA, int DW, int DH>
__global__ void __launch_bounds__(WS_THREADS, 2) warp_affine_ws_kernel(WarpArgs a, int B) {
    const int kdw = DW > 0 ? DW : a.dw, kdh = DW > 0 ? DH : a.dh;      // configured output size: dw / dh fold into immediates
    extern __shared__ __align__(128) uint8_t s_dyn[];          // WS_STAGES * WS_STAGE_ALLOC
    __shared__ float s_lut[768];
    __shared__ WarpTileDesc s_desc[WS_DESC];
    __shared__ __align__(8) uint64_t s_dfull[WS_DESC], s_dempty[WS_DESC], s_full[WS_STAGES], s_free[WS_STAGES];

    // warp roles: with WS_CFG_PLANNER_LOW the planners take the LOWEST warp ids (the issue arbiter prefers high warp ids,
    // and the planners should never win a slot against a blending warp)
    const int tid = threadIdx.x, lane = tid & 31, wrp_raw = tid >> 5;
    const bool is_planner = WS_CFG_PLANNER_LOW ? wrp_raw < WS_PLANNER_WARPS : wrp_raw >= WS_CONSUMER_WARPS;
    const int wrp = WS_CFG_PLANNER_LOW ? (is_planner ? WS_CONSUMER_WARPS + wrp_raw : wrp_raw - WS_PLANNER_WARPS) : wrp_raw;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WS_DESC; ++s) {
            mbar_init(smem_addr(&s_dfull[s]), 1);
            mbar_init(smem_addr(&s_dempty[s]), WS_GROUP_WARPS);     // the warps of the group that owns the tile
        }
#pragma unroll
        for (int s = 0; s < WS_STAGES; ++s) {
            mbar_init(smem_addr(&s_full[s]), TMA ? 1 : WS_GROUP_WARPS * 32);   // TMA: the issuing lane (+ tx bytes); LDGSTS: one async arrival per lane
            mbar_init(smem_addr(&s_free[s]), WS_GROUP_WARPS);                  // every warp of the group is done reading the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (HAS_NORM)
        for (int i = tid; i < 768; i += WS_THREADS) s_lut[i] = a.lut[i];
    __syncthreads();

    const int tiles_x = (kdw + WT_TW - 1) / WT_TW, tiles_y = (kdh + WT_TH - 1) / WT_TH;
    const int tps = tiles_x * tiles_y;                                  // 64-row tiles per sample
    constexpr int TH_SMALL = WT_TH / 2;                                  // the tiles handed out last are half as high: the launch ends on short items
    const int tps_small = tiles_x * ((kdh + TH_SMALL - 1) / TH_SMALL);
    const int B_big = WS_CFG_TAIL_DIV > 0 ? B - (B + WS_CFG_TAIL_DIV - 1) / max(WS_CFG_TAIL_DIV, 1) : B;   // samples cut into full-height tiles; the last 1 / TAIL_DIV of them into half-height tiles
    const int64_t ntiles_big = (int64_t)B_big * tps;
    const int64_t ntiles = ntiles_big + (int64_t)(B - B_big) * tps_small;
    // tile i of this CTA is global tile blockIdx.x + i*gridDim.x: every CTA sees a mix of samples, so
    // heavy samples (strong down-scaling -> big boxes, more bands) do not pile up on one CTA
    const int n_my = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (is_planner) {
        // ====================================== PLANNERS ========================================
        // Planner warp p describes the tiles p, p+P, ... of this CTA: OpenCV's fixed-point column / row
        // terms, the source box of each band, its staging mode and copy plan.  Descriptors go through a
        // WS_DESC-deep ring, several tiles ahead of the consumers; planners never touch pixel data.
        const int pw = wrp - WS_CONSUMER_WARPS;
        for (int i = pw;; i += WS_PLANNER_WARPS) {
            const int slot = i % WS_DESC;
            WarpTileDesc& D = s_desc[slot];
            mbar_wait_backoff(smem_addr(&s_dempty[slot]), ((uint32_t)(i / WS_DESC) & 1u) ^ 1u);
            // Tile heights.  A 64-row tile halves the per-item hand-off cost per pixel (one descriptor, one copy plan, one barrier
            // round for 2048 pixels), but with ~10 of them per consumer group the launch would end with groups idling for up to a
            // whole tile; so (WS_CFG_TAIL_DIV) the tiles of the last samples can be half as high and are handed out after all the others.
            int64_t t;
            int th = WT_TH;                                 // height of this tile
            if (WS_CFG_DYNAMIC) {
                unsigned int claimed = 0, small_t = 0;
                if (lane == 0) {
                    claimed = atomicAdd(a.tile_counter, 1u);
                    if ((int64_t)claimed >= ntiles_big) small_t = atomicAdd(a.tile_counter + 2, 1u);
                }
                t = (int64_t)__shfl_sync(0xffffffffu, claimed, 0);
                if (t >= ntiles_big) {
                    t = ntiles_big + (int64_t)__shfl_sync(0xffffffffu, small_t, 0);
                    th = TH_SMALL;
                }
            } else {
                t = i < n_my ? (int64_t)blockIdx.x + (int64_t)i * gridDim.x : ntiles;
                if (t >= ntiles_big) th = TH_SMALL;
            }
            if (t >= ntiles) {
                // end marker of THIS planner: the consumers skip its later slots (the other planner of the group may still
                // hold tiles it claimed earlier)
                if (lane == 0) { D.nbands = 0; mbar_arrive(smem_addr(&s_dfull[slot])); }
                break;
            }
            int b, tx, ty;
            if (th == WT_TH) {
                b = (int)(t / tps);
                const int rem = (int)(t - (int64_t)b * tps);
                ty = rem / tiles_x; tx = rem - ty * tiles_x;
            } else {
                const int64_t ts = t - ntiles_big;
                const int bs = (int)(ts / tps_small), rem = (int)(ts - (int64_t)bs * tps_small);
                b = B_big + bs; ty = rem / tiles_x; tx = rem - ty * tiles_x;
            }
            const int x0 = tx * WT_TW, y0 = ty * th;
            const int H = a.src_h[b], W = a.src_w[b];
            const int64_t pitch = a.src_pitch[b];
            const uint8_t* src = a.src_base + a.src_off[b];
            const bool flip = a.flip && a.flip[b];
            const bool bulk_ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)pitch) & 15) == 0 && pitch >= (int64_t)W * 3;
            double Minv[6];
            if (a.M) {
                invert_affine(a.M + 6 * b, Minv);          // every lane (uniform values)
            } else {
                double Mf[6];
                affine_from_csr(a.center[2 * b], a.center[2 * b + 1], a.scale[2 * b], a.scale_f32, a.rot[b], kdw, kdh, Mf);
                invert_affine(Mf, Minv);
            }
            const double yy = (double)min(y0 + lane, kdh - 1), xx = (double)min(x0 + lane, kdw - 1);
            const int X0l = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Minv[1], yy), Minv[2]), 1024.0)) + ROUND_DELTA;
            const int Y0l = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Minv[4], yy), Minv[5]), 1024.0)) + ROUND_DELTA;
            const int adl = __double2int_rn(__dmul_rn(__dmul_rn(Minv[0], xx), 1024.0));
            const int bdl = __double2int_rn(__dmul_rn(__dmul_rn(Minv[3], xx), 1024.0));
            D.ad[lane] = adl; D.bd[lane] = bdl; D.X0[lane] = X0l; D.Y0[lane] = Y0l;
            int X0h = X0l, Y0h = Y0l;                       // rows 32..63 of the tile (WT_TH == 64)
            if (WT_TH > 32) {
                const double yh = (double)min(y0 + 32 + lane, kdh - 1);
                X0h = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Minv[1], yh), Minv[2]), 1024.0)) + ROUND_DELTA;
                Y0h = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(Minv[4], yh), Minv[5]), 1024.0)) + ROUND_DELTA;
                D.X0[(32 + lane) % WT_TH] = X0h; D.Y0[(32 + lane) % WT_TH] = Y0h;
            }
            auto row_term = [&](int lo_v, int hi_v, int r) {       // value of tile row r (uniform r)
                const int vl = __shfl_sync(0xffffffffu, lo_v, r & 31), vh = __shfl_sync(0xffffffffu, hi_v, r & 31);
                return r < 32 ? vl : vh;
            };
            const int adL = __shfl_sync(0xffffffffu, adl, 0), adR = __shfl_sync(0xffffffffu, adl, 31);
            const int bdL = __shfl_sync(0xffffffffu, bdl, 0), bdR = __shfl_sync(0xffffffffu, bdl, 31);
            int rpp = th, nb = 0;                          // rows per band: 64, 32, 16 or 8
            for (int band = 0; band < th && y0 + band < kdh;) {
                // box of rows [band, band+rpp): X and Y are monotone in x and y -> extremes at the corners
                const int rA = band, rB = band + rpp - 1;
                const int X0A = row_term(X0l, X0h, rA), X0B = row_term(X0l, X0h, rB);
                const int Y0A = row_term(Y0l, Y0h, rA), Y0B = row_term(Y0l, Y0h, rB);
                const int sx0 = sat16((X0A + adL) >> AB_BITS), sx1 = sat16((X0A + adR) >> AB_BITS);
                const int sx2 = sat16((X0B + adL) >> AB_BITS), sx3 = sat16((X0B + adR) >> AB_BITS);
                const int sy0 = sat16((Y0A + bdL) >> AB_BITS), sy1 = sat16((Y0A + bdR) >> AB_BITS);
                const int sy2 = sat16((Y0B + bdL) >> AB_BITS), sy3 = sat16((Y0B + bdR) >> AB_BITS);
                const int bx0 = min(min(sx0, sx1), min(sx2, sx3)), bx1 = max(max(sx0, sx1), max(sx2, sx3)) + 1;
                const int by0 = min(min(sy0, sy1), min(sy2, sy3)), by1 = max(max(sy0, sy1), max(sy2, sy3)) + 1;
                const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
                const int c_lo = flip ? (W - 1 - bx1) : bx0;      // ascending SOURCE columns
                const int A0 = floor16(3 * c_lo), A1 = floor16(3 * (c_lo + bw) + 15);
                int rowpitch = A1 - A0, cls = 0, ncopy = 0;
                bool fits;
                if (TMA) {
                    // row pitch rounded up to a width class (one tensor map per class), rows to whole 8-row boxes
                    const int need = rowpitch >> 4;
                    while (cls < TM_NCLS && c_tm_cls[cls] < need) ++cls;
                    ncopy = (bh + TM_ROWS - 1) / TM_ROWS;
                    rowpitch = cls < TM_NCLS ? 16 * c_tm_cls[cls] : rowpitch;
                    fits = cls < TM_NCLS && (int64_t)rowpitch * ncopy * TM_ROWS <= WS_STAGE_BYTES;
                } else {
                    fits = (int64_t)rowpitch * bh <= WS_STAGE_BYTES;
                }
                if (bulk_ok && !fits && rpp > 8) { rpp >>= 1; continue; }   // retry this band with fewer rows
                int mode = WS_MODE_DIRECT;
                if (bulk_ok && fits)
                    mode = (by0 >= 0 && by1 < H && c_lo >= 0 && c_lo + bw <= W) ? WS_MODE_STAGED : WS_MODE_BORDER;
                if (lane == 0) {
                    // rows / bytes outside the image are simply not copied (BORDER mode masks those taps)
                    const int cs = max(A0, 0), ce = min(A1, (int)pitch);
                    const int nb16 = mode == WS_MODE_DIRECT ? 0 : (max(ce - cs, 0) >> 4);
                    const int ry_lo = max(0, -by0), ry_hi = min(bh, H - by0);
                    WarpBand wb;
                    wb.r0 = band; wb.r1 = min(min(band + rpp, th), kdh - y0);
                    wb.rowpitch = rowpitch; wb.mode = mode;
                    wb.nb16 = nb16; wb.cshift = nb16 <= 8 ? 3 : (nb16 <= 16 ? 4 : (nb16 <= 32 ? 5 : 6));
                    wb.nrows = ry_hi - ry_lo; wb.d0 = cs - A0 + ry_lo * rowpitch;
                    wb.g0 = src + (int64_t)(by0 + ry_lo) * pitch + cs;
                    wb.koff = -(by0 * rowpitch + A0) + (flip ? 3 * (W - 2) : 0);
                    wb.by0 = by0; wb.A0 = A0; wb.cls = cls; wb.ncopy = mode == WS_MODE_DIRECT ? 0 : ncopy; wb.pad_ = 0;
                    D.band[nb] = wb;
                }
                ++nb;
                band += rpp;
            }
            if (lane == 0) {
                D.b = b; D.x0 = x0; D.y0 = y0; D.H = H; D.W = W; D.flip = flip ? 1 : 0; D.nbands = nb;
                D.pitch = pitch; D.src = src;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(&s_dfull[slot]));   // release: publishes the descriptor
        }
    } else {

    // =========================================== CONSUMERS ===========================================
    // WS_GROUPS independent groups of WS_GROUP_WARPS warps; group g owns the tiles g, g+G, ... of this CTA
    // and runs a WS_GSTAGES-stage cp.async pipeline over their bands: all threads of the group issue the
    // 16-byte copies of the next band (the copy completion arrives on full[stage]), then blend the current
    // one.  Many warps share the copy issue (a single warp's LDGSTS queue is shallow), and the groups
    // overlap each other's barrier / descriptor latency.
    const int64_t plane = (int64_t)kdh * kdw;
    const uint32_t lut32 = smem_addr(s_lut);
    const uint32_t dyn_base = smem_addr(s_dyn), full_base = smem_addr(&s_full[0]), free_base = smem_addr(&s_free[0]);      // shared-window addresses once, not per item
    static_assert(WS_PLANNER_WARPS >= WS_GROUPS, "every consumer group needs an end marker on its own tile sequence");
    const int grp = wrp / WS_GROUP_WARPS, gw = wrp - grp * WS_GROUP_WARPS;
    const int ctid = gw * 32 + lane;                      // 0..127 within the group
    int pt = grp, pb = 0, n_pref = 0;                     // prefetch cursor: tile, band, item count
    int cb = 0, n_comp = 0;                               // compute cursor: band of tile q0, item count
    bool pref_done = false;
    // tiles whose descriptor the prefetch cursor has found valid and the compute cursor has not finished yet (the prefetch runs
    // WS_GSTAGES - 1 = 1 item ahead, so at most two): the compute side takes its tiles from here instead of waiting on the
    // descriptor barrier and re-reading nbands a second time
    int q0 = -1, q1 = -1, q0n = 0, q1n = 0, pnb = 0;      // tile slots, their band counts; band count of the prefetch cursor's tile
    static_assert(WS_GSTAGES == 2, "the tile FIFO below holds two entries");
    const CUtensorMap* last_tm = nullptr;
    // slot i is written by planner i % WS_PLANNER_WARPS; a planner that ran out of tiles leaves ONE end marker and stops,
    // so each cursor remembers which of the group's planners are finished and steps over their slots
    constexpr uint32_t ALL_PLANNERS = (1u << WS_PLANNER_WARPS) - 1u;
    uint32_t gmask = 0;
    for (int pl_ = grp; pl_ < WS_PLANNER_WARPS; pl_ += WS_GROUPS) gmask |= 1u << pl_;
    uint32_t pdone = ALL_PLANNERS & ~gmask;
    auto next_slot = [&](int i, uint32_t done) {          // next slot of this group whose planner is alive (caller checks done != ALL)
        do { i += WS_GROUPS; } while (done & (1u << (i & (WS_PLANNER_WARPS - 1))));
        return i;
    };

    auto prefetch_one = [&]() {
        if (pref_done) return;
        int slot = (pt & (WS_DESC - 1));
        if (pb == 0) {
            for (;;) {
                mbar_wait(smem_addr(&s_dfull[slot]), ((uint32_t)pt >> WS_DESC_LOG2) & 1u);
                pnb = s_desc[slot].nbands;
                if (pnb != 0) break;
                pdone |= 1u << (pt & (WS_PLANNER_WARPS - 1));
                if (pdone == ALL_PLANNERS) { pref_done = true; return; }
                pt = next_slot(pt, pdone);
                slot = (pt & (WS_DESC - 1));
            }
            if (q0 < 0) { q0 = pt; q0n = pnb; } else { q1 = pt; q1n = pnb; }
        }
        const WarpTileDesc& D = s_desc[slot];
        const int stage = grp * WS_GSTAGES + (n_pref & (WS_GSTAGES - 1));
        const WarpBand& Bd = D.band[pb];
        if (TMA) {
            // warp 0 of the group issues the band's boxes, one per lane; the other warps go straight back to blending
            if (gw == 0) {
                const uint32_t mbar = smem_addr(&s_full[stage]);
                if (Bd.ncopy > 0) {
                    const CUtensorMap* tm = a.tmaps + (size_t)D.b * TM_NCLS + Bd.cls;
                    if (tm != last_tm && !(a.tma_flags & 1)) {          // maps are written by the kernel in front of this one (tensormap proxy: acquire once per map)
                        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
                        last_tm = tm;
                    }
                    const uint32_t box_bytes = (uint32_t)(TM_ROWS * Bd.rowpitch);
                    if (lane == 0) mbar_arrive_expect_tx(mbar, box_bytes * (uint32_t)Bd.ncopy);
                    __syncwarp();
                    const uint32_t d0 = smem_addr(s_dyn + (size_t)stage * WS_STAGE_ALLOC);
                    const int c0 = Bd.A0 >> 3;                     // 8-byte elements
                    for (int k = lane; k < Bd.ncopy; k += 32)
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                     ::"r"(d0 + (uint32_t)k * box_bytes), "l"(tm), "r"(mbar), "r"(c0), "r"(Bd.by0 + k * TM_ROWS) : "memory");
                } else if (lane == 0) {
                    mbar_arrive(mbar);             // DIRECT item: nothing to stage
                }
            }
            ++n_pref;
            if (++pb == pnb) { pb = 0; pt = next_slot(pt, pdone); }
            return;
        }
        // the stage was last read by item n_pref - WS_GSTAGES: every warp of the group arrives on s_free[stage] when it is done with
        // it (a group-wide bar.sync at the top of the loop did this before; the arrival now happens right after a warp's last
        // blend and the wait only here, after the copy set-up, so a slower warp is waited for under useful work)
        if (n_pref >= WS_GSTAGES) mbar_wait(free_base + 8u * (uint32_t)stage, (((uint32_t)n_pref >> WS_GSTAGES_LOG2) - 1u) & 1u);
        const int nb16 = WS_CFG_KNOCKOUT_COPY ? 0 : Bd.nb16;
        if (nb16 > 0) {
            // lanes spread over (row, chunk): the group's 128 threads cover 128 >> cshift rows per pass
            const int cshift = Bd.cshift;
            const int ci = ctid & ((1 << cshift) - 1), rsub = ctid >> cshift, rstep = (WS_GROUP_WARPS * 32) >> cshift;
            if (ci < nb16) {
                // everything the loop needs in registers first: the "memory" clobber of the copy instruction would otherwise
                // re-read the band from shared memory every iteration (ncu: 10 % of the kernel's stall samples sat on that
                // load -> compare -> branch chain, 13 instructions per copy)
                const int64_t pitch = D.pitch;
                const int rowpitch = Bd.rowpitch;
                const uint8_t* g = Bd.g0 + (int64_t)rsub * pitch + 16 * ci;
                uint32_t d = dyn_base + (uint32_t)(stage * WS_STAGE_ALLOC) + (uint32_t)(Bd.d0 + rsub * rowpitch + 16 * ci);
                const int64_t gstep = (int64_t)rstep * pitch;
                const uint32_t dstep = (uint32_t)(rstep * rowpitch);
                int left = Bd.nrows - rsub;                              // rows still to copy at or below this thread's first row
                // two copies per iteration (their loop control shared), then a possible last one
                const uint8_t* g2 = g + gstep;
                uint32_t d2 = d + dstep;
                for (; left > rstep; left -= 2 * rstep, g += 2 * gstep, g2 += 2 * gstep, d += 2 * dstep, d2 += 2 * dstep) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d2), "l"(g2) : "memory");
                }
                if (left > 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
            }
        }
        // arrives on full[stage] once all of this lane's copies have landed
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_base + 8u * (uint32_t)stage) : "memory");
        ++n_pref;
        if (++pb == pnb) { pb = 0; pt = next_slot(pt, pdone); }
    };

#pragma unroll 1
    for (int k = 0; k < WS_GSTAGES - 1; ++k) prefetch_one();

    for (;;) {
        if (q0 < 0) break;             // no tile left: the prefetch cursor has met the end markers of all planners of this group
        const int slot = (q0 & (WS_DESC - 1));
        const WarpTileDesc& D = s_desc[slot];
        if (TMA) group_bar(grp);       // (TMA path: one warp issues the boxes, the group still meets here)
        prefetch_one();
        const int stage = grp * WS_GSTAGES + (n_comp & (WS_GSTAGES - 1));
        const WarpBand& Bd = D.band[cb];
        const int mode = Bd.mode;
        const int x = D.x0 + lane, y0 = D.y0, r0 = Bd.r0 + gw, r1 = Bd.r1;
        const int H = D.H, W = D.W;
        const bool flip = D.flip != 0;
        const int64_t sample_off = (int64_t)D.b * plane;
        if (x < kdw && r0 < r1) {
            const int ad = D.ad[lane], bd = D.bd[lane];
            const uint32_t x0s = smem_addr(&D.X0[0]), y0s = smem_addr(&D.Y0[0]);
            WarpOut<HAS_U8, HAS_NORM, BF16, DW, DH> out;
            out.init(a, sample_off, (int64_t)(y0 + r0) * kdw + x, plane, WS_GROUP_WARPS);
            // the descriptor reads and the cursor set-up above do not need the staged bytes: wait for them only now
            mbar_wait(full_base + 8u * (uint32_t)stage, ((uint32_t)n_comp >> WS_GSTAGES_LOG2) & 1u);
            if (mode != WS_MODE_DIRECT) {
                const int rowpitch = Bd.rowpitch;
                const uint32_t K = dyn_base + (uint32_t)(stage * WS_STAGE_ALLOC) + (uint32_t)Bd.koff;
                if (mode == WS_MODE_STAGED) {
                    if (flip) warp_rows_staged<HAS_U8, HAS_NORM, BF16, true, false, DW, DH>(out, x0s, y0s, r0, r1, ad, bd, K, rowpitch, H, W, lut32);
                    else warp_rows_staged<HAS_U8, HAS_NORM, BF16, false, false, DW, DH>(out, x0s, y0s, r0, r1, ad, bd, K, rowpitch, H, W, lut32);
                } else {
                    if (flip) warp_rows_staged<HAS_U8, HAS_NORM, BF16, true, true, DW, DH>(out, x0s, y0s, r0, r1, ad, bd, K, rowpitch, H, W, lut32);
                    else warp_rows_staged<HAS_U8, HAS_NORM, BF16, false, true, DW, DH>(out, x0s, y0s, r0, r1, ad, bd, K, rowpitch, H, W, lut32);
                }
            } else {
                const uint8_t* src = D.src;
                const int64_t pitch = D.pitch;
                for (int r = r0; r < r1; r += WS_GROUP_WARPS) {
                    const int X = ((int)lds32(x0s + 4u * r) + ad) >> (AB_BITS - INTER_BITS);
                    const int Y = ((int)lds32(y0s + 4u * r) + bd) >> (AB_BITS - INTER_BITS);
                    const uint32_t fx = X & 31, fy = Y & 31;
                    const int sx = sat16(X >> INTER_BITS), sy = sat16(Y >> INTER_BITS);
                    const bool y0ok = (unsigned)sy < (unsigned)H, y1ok = (unsigned)(sy + 1) < (unsigned)H;
                    const bool x0ok = (unsigned)sx < (unsigned)W, x1ok = (unsigned)(sx + 1) < (unsigned)W;
                    const int cx0 = flip ? (W - 1 - sx) : sx, cx1 = flip ? (W - 2 - sx) : (sx + 1);
                    const uint8_t* q0 = src + (int64_t)sy * pitch;
                    const uint8_t* q1 = q0 + pitch;
                    uint32_t rbA = 0, gA = 0, rbB = 0, gB = 0, rbC = 0, gC = 0, rbD = 0, gD = 0;
                    if (y0ok && x0ok) { const uint8_t* q = q0 + 3 * cx0; rbA = __ldg(q) | (__ldg(q + 2) << 16); gA = __ldg(q + 1); }
                    if (y0ok && x1ok) { const uint8_t* q = q0 + 3 * cx1; rbB = __ldg(q) | (__ldg(q + 2) << 16); gB = __ldg(q + 1); }
                    if (y1ok && x0ok) { const uint8_t* q = q1 + 3 * cx0; rbC = __ldg(q) | (__ldg(q + 2) << 16); gC = __ldg(q + 1); }
                    if (y1ok && x1ok) { const uint8_t* q = q1 + 3 * cx1; rbD = __ldg(q) | (__ldg(q + 2) << 16); gD = __ldg(q + 1); }
                    warp_emit<HAS_U8, HAS_NORM, BF16, DW, DH>(out, rbA, gA, rbB, gB, rbC, gC, rbD, gD, 32u - fx, fx, fy, lut32);
                    out.advance();
                }
            }
        }
        if (!TMA) {                    // this warp is done with the stage
            __syncwarp();
            if (lane == 0) mbar_arrive(free_base + 8u * (uint32_t)stage);
        }
        ++n_comp;
        if (++cb == q0n) {
            cb = 0; q0 = q1; q0n = q1n; q1 = -1;
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(&s_dempty[slot]));   // this warp no longer reads the descriptor
        }
    }
    }   // consumers
    if (WS_CFG_DYNAMIC) {
        // the last CTA to finish puts the tile counter back to zero for the next launch that uses this ring entry (every
        // planner has made its final, out-of-range claim by then), so no memset sits between the matrices and this kernel
        __syncthreads();
        if (tid == 0 && atomicAdd(a.tile_counter + 1, 1u) == gridDim.x - 1) {
            a.tile_counter[0] = 0u;
            a.tile_counter[1] = 0u;
            a.tile_counter[2] = 0u;
        }
    }
}

// One warp per (sample, width class): template -> shared memory, tensormap.replace the fields that depend on the sample and
// the class, then publish to global memory with the tensormap-proxy release fence.
__global__ void __launch_bounds__(32 * TM_NCLS)
tmap_build_kernel(const __grid_constant__ CUtensorMap tmpl, const uint8_t* __restrict__ src_base, const int64_t* __restrict__ src_off,
                  const int32_t* __restrict__ src_h, const int64_t* __restrict__ src_pitch, CUtensorMap* __restrict__ out, int B) {
    __shared__ __align__(128) CUtensorMap sm[TM_NCLS];
    const int b = blockIdx.x, cls = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    reinterpret_cast<uint32_t*>(&sm[cls])[lane] = reinterpret_cast<const uint32_t*>(&tmpl)[lane];
    __syncwarp();
    const uint8_t* src = src_base + src_off[b];
    const int64_t pitch = src_pitch[b];
    const bool ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)pitch) & 15) == 0 && pitch >= 16;   // unaligned sources never use their maps
    if (lane == 0 && ok) {
        const uint32_t m = smem_addr(&sm[cls]);
        asm volatile("tensormap.replace.tile.global_address.shared::cta.b1024.b64 [%0], %1;" ::"r"(m), "l"(src) : "memory");
        asm volatile("tensormap.replace.tile.global_dim.shared::cta.b1024.b32 [%0], 0, %1;" ::"r"(m), "r"((uint32_t)(pitch >> 3)) : "memory");
        asm volatile("tensormap.replace.tile.global_dim.shared::cta.b1024.b32 [%0], 1, %1;" ::"r"(m), "r"((uint32_t)src_h[b]) : "memory");
        asm volatile("tensormap.replace.tile.global_stride.shared::cta.b1024.b64 [%0], 0, %1;" ::"r"(m), "l"((uint64_t)pitch) : "memory");
        asm volatile("tensormap.replace.tile.box_dim.shared::cta.b1024.b32 [%0], 0, %1;" ::"r"(m), "r"((uint32_t)(2 * c_tm_cls[cls])) : "memory");
    }
    __syncwarp();
    CUtensorMap* g = out + (size_t)b * TM_NCLS + cls;
    asm volatile("tensormap.cp_fenceproxy.global.shared::cta.tensormap::generic.release.gpu.sync.aligned [%0], [%1], 128;"
                 ::"l"(g), "r"(smem_addr(&sm[cls])) : "memory");
}

// The template: a 2-D tensor of 8-byte elements {128, 64} with a row stride of 1024 bytes and box {8, 8}; every sample- or
// class-dependent field is replaced on the device.  cuTensorMapEncodeTiled comes from the driver at run time.
static const CUtensorMap* tmap_template() {
    static CUtensorMap tmpl;
    static int state = 0;                        // 0: not tried, 1: ok, -1: unavailable
    static std::atomic<bool> guard{false};
    while (guard.exchange(true)) {}
    if (state == 0) {
        state = -1;
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn && qr == cudaDriverEntryPointSuccess) {
            const cuuint64_t dims[2] = {128, 64}, strides[1] = {1024};
            const cuuint32_t box[2] = {8, TM_ROWS}, estr[2] = {1, 1};
            void* dummy = nullptr;
            if (cudaMalloc(&dummy, 65536) == cudaSuccess) {
                if (reinterpret_cast<EncodeFn>(fn)(&tmpl, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, dummy, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    state = 1;
                cudaFree(dummy);
            }
        }
        cudaGetLastError();
    }
    guard.store(false);
    return state == 1 ? &tmpl : nullptr;
}

template <bool HAS_U8, bool HAS_NORM, bool BF16>
static int launch_warp_tile(const WarpArgs& a_in, int B, cudaStream_t s) {
    const size_t smem = (size_t)WS_STAGES * WS_STAGE_ALLOC;
    ADVMIX_CUDA_OK(ensure_dyn_smem(warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 0, 0>, (int)smem));
    ADVMIX_CUDA_OK(ensure_dyn_smem(warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, true, 0, 0>, (int)smem));
    ADVMIX_CUDA_OK(ensure_dyn_smem(warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 192, 256>, (int)smem));
    ADVMIX_CUDA_OK(ensure_dyn_smem(warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 256, 256>, (int)smem));
    ADVMIX_CUDA_OK(ensure_dyn_smem(warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 512, 512>, (int)smem));
    WarpArgs a = a_in;
    a.tmaps = nullptr;
    { static const int flags = getenv("ADVMIX_TMA_FLAGS") ? atoi(getenv("ADVMIX_TMA_FLAGS")) : 0; a.tma_flags = flags; }
    {
        // tensor maps of this launch: a slice of a per-device ring (launches on different streams may overlap), dedicated slices
        // for launches captured into CUDA graphs (the pointer is frozen into the graph), like the tile counters below
        // Opt-in (ADVMIX_WARP_TMA=1): measured on B200, 256 crops per launch (profiles/r2_ncu_warp_tma_vs_ldgsts.txt):
        // LDGSTS staging 82.4 us; TMA with 8-row boxes 106 us, 16-row boxes 102 us (95 us without the per-map proxy fence), 4-row
        // boxes and 28 width classes 111 us.  The TMA kernel executes 9 % fewer instructions but its warps wait longer for a
        // stage (issue-active 48 % vs 66 %, barrier stall 1.8 vs 0.6): box padding (width classes, whole 8/16-row boxes) adds
        // bytes and bands, and each box has a fixed cost the 16-byte LDGSTS copies, issued by 128 lanes at once, do not pay.
        static const bool use_tma = getenv("ADVMIX_WARP_TMA") != nullptr && getenv("ADVMIX_WARP_NO_TMA") == nullptr;
        const CUtensorMap* tmpl = use_tma ? tmap_template() : nullptr;
        constexpr int TM_RING = 4, TM_CAPTURED = 16, TM_MAXB = 512;      // 20 slices x 512 samples x 14 maps x 128 B = 18 MB per device
        if (tmpl && B <= TM_MAXB) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            ADVMIX_CUDA_OK(cudaStreamIsCapturing(s, &cap));
            static std::atomic<unsigned int> next_slice{0}, next_cap_slice{0};
            // cudaMalloc is illegal while capturing: the first eager (warm-up) call allocates, a capture only looks the buffer up
            CUtensorMap* ring = reinterpret_cast<CUtensorMap*>(cached_buffer(
                "warp_tensor_maps", cap != cudaStreamCaptureStatusNone ? 0 : (size_t)(TM_RING + TM_CAPTURED) * TM_MAXB * TM_NCLS * sizeof(CUtensorMap)));
            if (ring) {
                unsigned int slice;
                if (cap != cudaStreamCaptureStatusNone) {
                    slice = TM_RING + next_cap_slice.fetch_add(1);
                    if (slice >= (unsigned)(TM_RING + TM_CAPTURED)) slice = UINT_MAX;      // out of dedicated slices: LDGSTS staging for this graph node
                } else {
                    slice = next_slice.fetch_add(1) % TM_RING;
                }
                if (slice != UINT_MAX) {
                    CUtensorMap* maps = ring + (size_t)slice * TM_MAXB * TM_NCLS;
                    tmap_build_kernel<<<B, 32 * TM_NCLS, 0, s>>>(*tmpl, a.src_base, a.src_off, a.src_h, a.src_pitch, maps, B);
                    ADVMIX_LAUNCH_OK();
                    a.tmaps = maps;
                }
            }
        }
    }
    if (WS_CFG_DYNAMIC) {
        // One {tiles claimed, CTAs finished} pair per launch; the kernel leaves both at 0.  Eager launches take the next
        // entry of a 256-entry ring (launches on different streams may overlap; more than 256 warp launches in flight at
        // once are not supported).  A launch recorded into a CUDA graph freezes its pointer into the graph and replays
        // it for the life of the graph, so captured launches get dedicated entries that are never handed out again
        // (1024 per process).  The table must exist before a capture starts (cudaMalloc is illegal while capturing):
        // run the op once eagerly first, as any warm-up does.
        constexpr int RING = 256, CAPTURED = 1024;
        static const unsigned int zeros[4 * (RING + CAPTURED)] = {};        // {claimed 64-row tiles, finished CTAs, claimed 32-row tiles, -} per entry
        static std::atomic<unsigned int> next_counter{0}, next_captured{0};
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        ADVMIX_CUDA_OK(cudaStreamIsCapturing(s, &cap));
        unsigned int* ring = const_cast<unsigned int*>(reinterpret_cast<const unsigned int*>(cached_table("warp_tile_counters4", zeros, sizeof(zeros))));
        if (!ring) return fail(ADVMIX_ERR_CUDA, "warp_affine: tile-counter table unavailable%s",
                               cap != cudaStreamCaptureStatusNone ? " (first call inside a stream capture: run the op once eagerly before capturing)" : "");
        if (cap != cudaStreamCaptureStatusNone) {
            const unsigned int k = next_captured.fetch_add(1);
            if (k >= (unsigned)CAPTURED) return fail(ADVMIX_ERR_UNSUPPORTED, "warp_affine: more than %d launches captured into CUDA graphs", CAPTURED);
            a.tile_counter = ring + 4 * (RING + k);
        } else {
            a.tile_counter = ring + 4 * (next_counter.fetch_add(1) % RING);
        }
    }
    const int tiles_y = (a.dh + WT_TH - 1) / WT_TH;
    const int tiles_x = (a.dw + WT_TW - 1) / WT_TW;
    const int grid = (int)std::min<int64_t>((int64_t)B * tiles_y * tiles_x, 2 * sm_count());
    // the configured output sizes (COCO 192x256, MPII 256x256, bottom-up 512x512) get kernels with dw / dh as compile-time constants
    if (a.tmaps) warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, true, 0, 0><<<grid, WS_THREADS, smem, s>>>(a, B);
    else if (a.dw == 192 && a.dh == 256) warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 192, 256><<<grid, WS_THREADS, smem, s>>>(a, B);
    else if (a.dw == 256 && a.dh == 256) warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 256, 256><<<grid, WS_THREADS, smem, s>>>(a, B);
    else if (a.dw == 512 && a.dh == 512) warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 512, 512><<<grid, WS_THREADS, smem, s>>>(a, B);
    else warp_affine_ws_kernel<HAS_U8, HAS_NORM, BF16, false, 0, 0><<<grid, WS_THREADS, smem, s>>>(a, B);
    return ADVMIX_OK;
}

// ---- get_affine_transform (transforms.py:69-101), batched ---------------------------
// One warp per CTA: the kernel is a chain of float64 latencies (cv2's LU solve per sample); eight small CTAs on eight SMs
// finish in 12 us where two 128-thread CTAs took 18 us.
__global__ void affine_matrices_kernel(const float* __restrict__ center, const double* __restrict__ scale,
                                       const double* __restrict__ rot, double* __restrict__ M, int B,
                                       int out_w, int out_h, int scale_f32) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double m[6];
    affine_from_csr(center[2 * b], center[2 * b + 1], scale[2 * b], scale_f32, rot[b], out_w, out_h, m);
#pragma unroll
    for (int k = 0; k < 6; ++k) M[6 * b + k] = m[k];
}

// ---- fliplr_joints + affine_transform ----------------------------------------------
struct JointsCsr {           // optional in-kernel get_affine_transform (M == nullptr); M_out receives the matrices
    const float* center;
    const double* scale;
    const double* rot;
    double* M_out;
    int scale_f32, out_w, out_h;
};

__global__ void joints_kernel(const double* __restrict__ jin, const double* __restrict__ vin,
                              const uint8_t* __restrict__ flip, const int32_t* __restrict__ src_w,
                              const int32_t* __restrict__ perm, const double* __restrict__ M, JointsCsr csr,
                              double* __restrict__ jout, double* __restrict__ vout, int B, int J,
                              const int32_t* __restrict__ rec = nullptr) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * J) return;
    const int b = t / J, j = t - b * J;
    const int64_t rb = rec ? rec[b] : b;               // row of the record table (advmix_joints_flip_affine_rec) or batch index
    double mloc[6];
    if (!M) {
        affine_from_csr(csr.center[2 * b], csr.center[2 * b + 1], csr.scale[2 * b], csr.scale_f32, csr.rot[b], csr.out_w, csr.out_h, mloc);
        if (j == 0 && csr.M_out)
            for (int k = 0; k < 6; ++k) csr.M_out[6 * b + k] = mloc[k];
    } else {
        for (int k = 0; k < 6; ++k) mloc[k] = M[6 * b + k];
    }
    double x, y, z, v0, v1, v2;
    if (flip && flip[b]) {
        const int s = perm ? perm[j] : j;
        const double* p = jin + (rb * J + s) * 3;
        const double* q = vin + (rb * J + s) * 3;
        v0 = q[0]; v1 = q[1]; v2 = q[2];
        // joints[:,0] = width - joints[:,0] - 1 ; then joints*joints_vis
        x = __dmul_rn(__dsub_rn(__dsub_rn((double)src_w[b], p[0]), 1.0), v0);
        y = __dmul_rn(p[1], v1);
        z = __dmul_rn(p[2], v2);
    } else {
        const double* p = jin + (rb * J + j) * 3;
        const double* q = vin + (rb * J + j) * 3;
        x = p[0]; y = p[1]; z = p[2];
        v0 = q[0]; v1 = q[1]; v2 = q[2];
    }
    if (v0 > 0.0) {
        const double* m = mloc;
        const double nx = fma(m[0], x, fma(m[1], y, m[2]));
        const double ny = fma(m[3], x, fma(m[4], y, m[5]));
        x = nx; y = ny;
    }
    double* o = jout + (int64_t)t * 3;
    o[0] = x; o[1] = y; o[2] = z;
    double* w = vout + (int64_t)t * 3;
    w[0] = v0; w[1] = v1; w[2] = v2;
}

// ---- ToTensor + Normalize ------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_kernel(const uint8_t* __restrict__ in, void* __restrict__ out,
                                                        const float* __restrict__ lut_g, int64_t groups,
                                                        int64_t plane, int dtype) {
    __shared__ float lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = g * 4;
        const int64_t b = pix / plane, r = pix - b * plane;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(in + pix * 3);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        uint8_t v[4][3] = {{(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16)},
                           {(uint8_t)(w0 >> 24), (uint8_t)w1, (uint8_t)(w1 >> 8)},
                           {(uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24), (uint8_t)w2},
                           {(uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)}};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f0 = lut[c * 256 + v[0][c]], f1 = lut[c * 256 + v[1][c]];
            const float f2 = lut[c * 256 + v[2][c]], f3 = lut[c * 256 + v[3][c]];
            const int64_t o = (b * 3 + c) * plane + r;
            if (dtype == ADVMIX_F32) {
                st_stream_f4(reinterpret_cast<float*>(out) + o, make_float4(f0, f1, f2, f3));
            } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(f0, f1), hi = __floats2bfloat162_rn(f2, f3);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + o) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
        }
    }
}

__global__ void normalize_kernel_scalar(const uint8_t* __restrict__ in, void* __restrict__ out,
                                        const float* __restrict__ lut, int64_t npix, int64_t plane, int dtype) {
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
         pix += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = pix / plane, r = pix - b * plane;
        for (int c = 0; c < 3; ++c) {
            const float f = lut[c * 256 + in[pix * 3 + c]];
            const int64_t o = (b * 3 + c) * plane + r;
            if (dtype == ADVMIX_F32) reinterpret_cast<float*>(out)[o] = f;
            else reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(f);
        }
    }
}


// Zero-copy gather of the source bytes this step's crops read: the SMs read the pinned (UVA-mapped) host
// buffer directly with 16-byte loads and write the same bytes at the same offsets of the device buffer.
// One launch for the whole batch.  Per sample the host passes the bounding box and the (convex) source
// quadrilateral of the crop; every row copies only the quad's extent over that row (+- margin), so a rotated
// crop does not drag its whole bounding box across PCIe.
struct H2DBox { int64_t off, pitch; int32_t row_lo, row_hi, byte_lo, byte_hi; float qx[4], qy[4]; };
constexpr int H2D_ROWS = 16;       // rows per CTA, two per warp
#ifndef H2D_ALIGN
#define H2D_ALIGN 16               // alignment of a row's byte range (a multiple of 16, at most the row pitch alignment)
#endif
constexpr float H2D_MARGIN = 3.0f; // pixels around the quad: bilinear taps + fixed-point rounding
__global__ void __launch_bounds__(256)
h2d_boxes_kernel(const uint8_t* __restrict__ host, uint8_t* __restrict__ dev, const H2DBox* __restrict__ boxes,
                 unsigned long long* __restrict__ bytes_out) {
    const H2DBox bx = boxes[blockIdx.y];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    unsigned long long sent = 0;
    for (int r = bx.row_lo + blockIdx.x * H2D_ROWS + wrp; r < min(bx.row_hi, bx.row_lo + (blockIdx.x + 1) * H2D_ROWS); r += 8) {
        // x-extent of the quad over the slab y in [r - m, r + m]: vertices inside the slab and edge crossings
        const float y0 = (float)r - H2D_MARGIN, y1 = (float)r + H2D_MARGIN;
        float xmin = 3.0e38f, xmax = -3.0e38f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float ax = bx.qx[k], ay = bx.qy[k], cx = bx.qx[(k + 1) & 3], cy = bx.qy[(k + 1) & 3];
            if (ay >= y0 && ay <= y1) { xmin = fminf(xmin, ax); xmax = fmaxf(xmax, ax); }
            const float dy = cy - ay;
            if (dy != 0.0f) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float t = ((e ? y1 : y0) - ay) / dy;
                    if (t >= 0.0f && t <= 1.0f) { const float x = ax + t * (cx - ax); xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); }
                }
            }
        }
        if (xmin > xmax) continue;                                         // the quad does not reach this row
        int b0 = (3 * ((int)floorf(xmin) - (int)H2D_MARGIN)) & ~(H2D_ALIGN - 1);
        int b1 = (3 * ((int)ceilf(xmax) + (int)H2D_MARGIN + 2) + H2D_ALIGN - 1) & ~(H2D_ALIGN - 1);
        b0 = max(b0, bx.byte_lo); b1 = min(b1, bx.byte_hi);
        if (b1 <= b0) continue;
        const int64_t o = bx.off + (int64_t)r * bx.pitch + b0;
        const int nch = (b1 - b0) >> 4;
        for (int c = lane; c < nch; c += 32)
            *reinterpret_cast<uint4*>(dev + o + 16 * c) = *reinterpret_cast<const uint4*>(host + o + 16 * c);
        if (lane == 0) sent += (unsigned long long)(b1 - b0);
    }
    if (bytes_out && sent) atomicAdd(bytes_out, sent);
}

}  // namespace advmix

using namespace advmix;

static int launch_warp_any(const WarpArgs& a, int B, bool u8, bool nm, bool bf, cudaStream_t st) {
    int rc;
    if (u8 && nm) rc = bf ? launch_warp_tile<true, true, true>(a, B, st) : launch_warp_tile<true, true, false>(a, B, st);
    else if (nm) rc = bf ? launch_warp_tile<false, true, true>(a, B, st) : launch_warp_tile<false, true, false>(a, B, st);
    else rc = launch_warp_tile<true, false, false>(a, B, st);
    if (rc) return rc;
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

extern "C" {

int advmix_warp_affine_u8c3(const uint8_t* src_base, const int64_t* src_off, const int32_t* src_h,
                            const int32_t* src_w, const int64_t* src_pitch, const uint8_t* flip_lr,
                            const double* M_fwd, uint8_t* dst_u8, void* dst_norm, const float* norm_lut,
                            int B, int dw, int dh, int norm_dtype, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && dw > 0 && dh > 0, "warp_affine: bad shape B=%d dw=%d dh=%d", B, dw, dh);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(src_off && src_h && src_w && src_pitch && M_fwd, "warp_affine: null argument");
    ADVMIX_REQUIRE(dst_u8 || dst_norm, "warp_affine: no output requested");
    ADVMIX_REQUIRE(!dst_norm || norm_lut, "warp_affine: dst_norm needs norm_lut");
    ADVMIX_REQUIRE(norm_dtype == ADVMIX_F32 || norm_dtype == ADVMIX_BF16, "warp_affine: bad dtype %d", norm_dtype);
    WarpArgs a{src_base, src_off, src_h, src_w, src_pitch, flip_lr, M_fwd, dst_u8, dst_norm, norm_lut, dw, dh, norm_dtype,
               nullptr, nullptr, nullptr, 0, nullptr};
    return launch_warp_any(a, B, dst_u8 != nullptr, dst_norm != nullptr, norm_dtype == ADVMIX_BF16, as_stream(stream));
}

int advmix_crop_csr_u8c3(const uint8_t* src_base, const int64_t* src_off, const int32_t* src_h, const int32_t* src_w,
                         const int64_t* src_pitch, const uint8_t* flip_lr, const float* center, const double* scale,
                         int scale_is_f32, const double* rot_deg, uint8_t* dst_u8, void* dst_norm, const float* norm_lut,
                         int B, int dw, int dh, int norm_dtype, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && dw > 0 && dh > 0, "crop_csr: bad shape B=%d dw=%d dh=%d", B, dw, dh);
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(src_off && src_h && src_w && src_pitch && center && scale && rot_deg, "crop_csr: null argument");
    ADVMIX_REQUIRE(dst_u8 || dst_norm, "crop_csr: no output requested");
    ADVMIX_REQUIRE(!dst_norm || norm_lut, "crop_csr: dst_norm needs norm_lut");
    ADVMIX_REQUIRE(norm_dtype == ADVMIX_F32 || norm_dtype == ADVMIX_BF16, "crop_csr: bad dtype %d", norm_dtype);
    WarpArgs a{src_base, src_off, src_h, src_w, src_pitch, flip_lr, nullptr, dst_u8, dst_norm, norm_lut, dw, dh, norm_dtype,
               center, scale, rot_deg, scale_is_f32, nullptr};
    return launch_warp_any(a, B, dst_u8 != nullptr, dst_norm != nullptr, norm_dtype == ADVMIX_BF16, as_stream(stream));
}

int advmix_joints_csr(const double* joints_in, const double* vis_in, const uint8_t* flip_lr, const int32_t* src_w,
                      const int32_t* flip_perm, const float* center, const double* scale, int scale_is_f32,
                      const double* rot_deg, double* M_fwd_out, double* joints_out, double* vis_out, int B, int J, int out_w,
                      int out_h, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0 && out_w > 0 && out_h > 0, "joints_csr: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(joints_in && vis_in && center && scale && rot_deg && joints_out && vis_out, "joints_csr: null argument");
    ADVMIX_REQUIRE(!flip_lr || src_w, "joints_csr: flip needs src_w");
    joints_kernel<<<ceil_div((long long)B * J, 128), 128, 0, as_stream(stream)>>>(
        joints_in, vis_in, flip_lr, src_w, flip_perm, nullptr, JointsCsr{center, scale, rot_deg, M_fwd_out, scale_is_f32, out_w, out_h},
        joints_out, vis_out, B, J);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_h2d_source_rows(const uint8_t* host_base_h, uint8_t* dev_base, const int64_t* off_h, const int64_t* pitch_h,
                           const int32_t* row_lo_h, const int32_t* row_hi_h, int B, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0, "h2d_source_rows: bad B");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(host_base_h && dev_base && off_h && pitch_h && row_lo_h && row_hi_h, "h2d_source_rows: null argument");
    cudaStream_t st = as_stream(stream);
    for (int b = 0; b < B; ++b) {
        if (row_hi_h[b] <= row_lo_h[b]) continue;
        const int64_t o = off_h[b] + (int64_t)row_lo_h[b] * pitch_h[b];
        ADVMIX_CUDA_OK(cudaMemcpyAsync(dev_base + o, host_base_h + o, (size_t)(row_hi_h[b] - row_lo_h[b]) * pitch_h[b],
                                       cudaMemcpyHostToDevice, st));
    }
    return ADVMIX_OK;
}


int advmix_h2d_source_boxes(const uint8_t* host_base, uint8_t* dev_base, const void* boxes, int B, int max_rows,
                            unsigned long long* bytes_out, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && max_rows >= 0, "h2d_source_boxes: bad B");
    if (B == 0 || max_rows == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(host_base && dev_base && boxes, "h2d_source_boxes: null argument");
    ADVMIX_REQUIRE(B <= 65535, "h2d_source_boxes: B <= 65535 per call");
    h2d_boxes_kernel<<<dim3(ceil_div(max_rows, H2D_ROWS), B), 256, 0, as_stream(stream)>>>(
        host_base, dev_base, reinterpret_cast<const H2DBox*>(boxes), bytes_out);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_affine_matrices(const float* center, const double* scale, int scale_is_f32, const double* rot_deg,
                           double* M_fwd, int B, int out_w, int out_h, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && out_w > 0 && out_h > 0, "affine_matrices: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(center && scale && rot_deg && M_fwd, "affine_matrices: null argument");
    affine_matrices_kernel<<<ceil_div(B, 32), 32, 0, as_stream(stream)>>>(center, scale, rot_deg, M_fwd, B, out_w, out_h, scale_is_f32);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_joints_flip_affine(const double* joints_in, const double* vis_in, const uint8_t* flip_lr,
                              const int32_t* src_w, const int32_t* flip_perm, const double* M_fwd,
                              double* joints_out, double* vis_out, int B, int J, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0, "joints_flip_affine: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(joints_in && vis_in && M_fwd && joints_out && vis_out, "joints_flip_affine: null argument");
    ADVMIX_REQUIRE(!flip_lr || src_w, "joints_flip_affine: flip needs src_w");
    joints_kernel<<<ceil_div((long long)B * J, 128), 128, 0, as_stream(stream)>>>(joints_in, vis_in, flip_lr, src_w, flip_perm,
                                                                                  M_fwd, JointsCsr{}, joints_out, vis_out, B, J);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_joints_flip_affine_rec(const double* rec_joints, const double* rec_vis, const int32_t* rec_idx, const uint8_t* flip_lr,
                                  const int32_t* src_w, const int32_t* flip_perm, const double* M_fwd, double* joints_out,
                                  double* vis_out, int B, int J, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && J > 0, "joints_flip_affine_rec: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(rec_joints && rec_vis && rec_idx && M_fwd && joints_out && vis_out, "joints_flip_affine_rec: null argument");
    ADVMIX_REQUIRE(!flip_lr || src_w, "joints_flip_affine_rec: flip needs src_w");
    joints_kernel<<<ceil_div((long long)B * J, 128), 128, 0, as_stream(stream)>>>(rec_joints, rec_vis, flip_lr, src_w, flip_perm,
                                                                                  M_fwd, JointsCsr{}, joints_out, vis_out, B, J, rec_idx);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

int advmix_normalize_u8c3(const uint8_t* in, void* out, const float* norm_lut, int B, int H, int W,
                          int norm_dtype, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0 && H > 0 && W > 0, "normalize: bad shape");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(in && out && norm_lut, "normalize: null argument");
    ADVMIX_REQUIRE(norm_dtype == ADVMIX_F32 || norm_dtype == ADVMIX_BF16, "normalize: bad dtype %d", norm_dtype);
    const int64_t plane = (int64_t)H * W, npix = plane * B;
    if (plane % 4 == 0) {
        const int64_t groups = npix / 4;
        int blocks = (int)std::min<int64_t>((groups + 255) / 256, (int64_t)sm_count() * 16);
        normalize_kernel<<<blocks, 256, 0, as_stream(stream)>>>(in, out, norm_lut, groups, plane, norm_dtype);
    } else {
        int blocks = (int)std::min<int64_t>((npix + 255) / 256, (int64_t)sm_count() * 16);
        normalize_kernel_scalar<<<blocks, 256, 0, as_stream(stream)>>>(in, out, norm_lut, npix, plane, norm_dtype);
    }
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"
