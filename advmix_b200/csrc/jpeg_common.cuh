// libjpeg(-turbo) integer building blocks shared by the jpeg_compression corruption and the JPEG decoder:
// jidctint's "islow" inverse DCT, the sample range limit, h2v2 fancy (triangle) up-sampling.
#pragma once
#include "common.cuh"

namespace advmix {

#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))


// one 1-D inverse pass; PASS2 adds the final descale (+3) - range limiting is done by the caller
template <bool PASS2>
__device__ __forceinline__ void idct8(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
    int z2 = d2, z3 = d6;
    int z1 = (z2 + z3) * FIX_0_541196100;
    int tmp2 = z1 + z3 * (-FIX_1_847759065), tmp3 = z1 + z2 * FIX_0_765366865;
    z2 = d0; z3 = d4;
    int tmp0 = (z2 + z3) << 13, tmp1 = (z2 - z3) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = d7; tmp1 = d5; tmp2 = d3; tmp3 = d1;
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * FIX_1_175875602;
    tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    const int sh = PASS2 ? 18 : 11;
    d0 = DESCALE(tmp10 + tmp3, sh); d7 = DESCALE(tmp10 - tmp3, sh);
    d1 = DESCALE(tmp11 + tmp2, sh); d6 = DESCALE(tmp11 - tmp2, sh);
    d2 = DESCALE(tmp12 + tmp1, sh); d5 = DESCALE(tmp12 - tmp1, sh);
    d3 = DESCALE(tmp13 + tmp0, sh); d4 = DESCALE(tmp13 - tmp0, sh);
}

__device__ __forceinline__ uint8_t range_limit(int x) {
    // libjpeg's sample_range_limit table (centred): index = x & RANGE_MASK
    const int i = x & 1023;
    if (i < 128) return (uint8_t)(i + 128);
    if (i < 512) return 255;
    if (i < 896) return 0;
    return (uint8_t)(i - 896);
}

#define ROWS8(F, a) F(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7])
#define COL8(F, a, c) F(a[c], a[8 + c], a[16 + c], a[24 + c], a[32 + c], a[40 + c], a[48 + c], a[56 + c])


__device__ __forceinline__ int up_h2v2(const uint8_t* __restrict__ C, int pitch, int ch, int cw, int y, int x) {
    const int cy = y >> 1, cx = x >> 1;
    const int ny = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
    const int nx = min(max((x & 1) ? cx + 1 : cx - 1, 0), cw - 1);
    const int thiscol = 3 * C[(size_t)cy * pitch + cx] + C[(size_t)ny * pitch + cx];
    const int othercol = 3 * C[(size_t)cy * pitch + nx] + C[(size_t)ny * pitch + nx];
    return (3 * thiscol + othercol + ((x & 1) ? 7 : 8)) >> 4;
}

__device__ __forceinline__ uint8_t clamp255(int v) { return (uint8_t)max(0, min(255, v)); }


}  // namespace advmix
