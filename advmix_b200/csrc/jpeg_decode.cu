// Row f1: baseline JPEG decode on the device, replacing cv2.imread at
// lib/dataset/JointsDataset.py:148 (and Image.open at tools/make_datasets.py:37) for the source images.
//
//   host  : advmix_jpeg_plan_h     - marker parser (SOF0/1, DQT, DHT, DRI, SOS), Huffman look-up tables,
//                                    output / workspace layout; touches only the file headers
//   device: jpeg_huffman_kernel    - entropy decode, one image per CTA (libjpeg's jdhuff.c bit for bit:
//                                    9-bit look-ahead table + maxcode walk, byte un-stuffing, RSTn handling)
//           jpeg_idct_kernel       - dequantise + jidctint "islow" IDCT + range limit, one thread per block
//           jpeg_color_kernel      - fancy (triangle) chroma up-sampling h2v2 / h2v1 + YCbCr->RGB|BGR
// The numeric pipeline is libjpeg(-turbo)'s integer decoder (JDCT_ISLOW, do_fancy_upsampling), which is
// what cv2.imread and PIL use, so decoded pixels are bit-identical to cv2.imdecode.
#include "jpeg_common.cuh"

#include <cstring>

namespace advmix {

struct HuffLut {
    uint16_t look[512];     // 9-bit look-ahead: (nbits << 8) | symbol, 0 = code longer than 9 bits
    int32_t maxcode[18];    // largest code of length l (-1 if none); [17] = sentinel
    int32_t valoff[17];     // huffval index of the first code of length l, minus that code
    uint8_t huffval[256];
    uint8_t pad[4];
};
static_assert(sizeof(HuffLut) == 1424, "HuffLut layout");

struct JpegPlan {
    int64_t file_off, file_len, out_off, out_pitch;                                          // 0
    int32_t width, height, ncomp, mcus_x, mcus_y, restart_interval, scan_off, scan_len;       // 32
    int32_t hs[4], vs[4], tq[4], td[4], ta[4];                                                // 64
    int32_t hmax, vmax, blocks_per_mcu, status;                                               // 144
    int64_t coef_off[4];        // int16 elements into the coefficient workspace                 160
    int64_t plane_off[4];       // bytes into the plane workspace                                192
    int32_t plane_w[4], plane_h[4];   // padded component planes (whole MCUs)                    224
    int32_t comp_w[4], comp_h[4];     // real component sizes ceil(W * hs / hmax)                256
    uint16_t quant[4][64];      // natural (row-major) order                                     288
    HuffLut huff[4];            // DC 0, DC 1, AC 0, AC 1                                        800
};
static_assert(sizeof(JpegPlan) == 800 + 4 * 1424, "JpegPlan layout");

enum { JPEG_OK = 0, JPEG_BAD = 1, JPEG_PROGRESSIVE = 2, JPEG_UNSUPPORTED = 3 };

static const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
__constant__ uint8_t c_zigzag[64];

// jpeg_make_d_derived_tbl (jdhuff.c)
static bool build_lut(const uint8_t* bits /*[17], index 1..16*/, const uint8_t* vals, int nvals, HuffLut* t) {
    memset(t, 0, sizeof(*t));
    char huffsize[257];
    unsigned huffcode[257];
    int p = 0;
    for (int l = 1; l <= 16; ++l) {
        if (p + bits[l] > 256) return false;
        for (int i = 0; i < bits[l]; ++i) huffsize[p++] = (char)l;
    }
    if (p != nvals) return false;
    huffsize[p] = 0;
    unsigned code = 0;
    int si = huffsize[0];
    for (int q = 0; huffsize[q];) {
        while (huffsize[q] == si) huffcode[q++] = code++;
        if (code > (1u << si)) return false;
        code <<= 1;
        ++si;
    }
    p = 0;
    for (int l = 1; l <= 16; ++l) {
        if (bits[l]) {
            t->valoff[l] = p - (int)huffcode[p];
            p += bits[l];
            t->maxcode[l] = (int)huffcode[p - 1];
        } else t->maxcode[l] = -1;
    }
    t->maxcode[17] = 0xFFFFF;
    p = 0;
    for (int l = 1; l <= 9; ++l)
        for (int i = 0; i < bits[l]; ++i, ++p) {
            const int look = (int)huffcode[p] << (9 - l);
            for (int c = 0; c < (1 << (9 - l)); ++c) t->look[look + c] = (uint16_t)((l << 8) | vals[p]);
        }
    memcpy(t->huffval, vals, nvals);
    return true;
}

static inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// Parses the headers of one file into `pl` (status != JPEG_OK on anything this decoder does not handle).
static void parse_one(const uint8_t* d, int64_t n, JpegPlan* pl) {
    pl->status = JPEG_BAD;
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return;
    int64_t p = 2;
    int comp_id[4] = {0, 0, 0, 0};
    bool have_sof = false, have_q[4] = {false, false, false, false}, have_h[4] = {false, false, false, false};
    bool adobe = false, jfif = false;
    int adobe_transform = 0;
    for (;;) {
        while (p < n && d[p] != 0xFF) ++p;
        while (p < n && d[p] == 0xFF) ++p;
        if (p >= n) return;
        const int m = d[p++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return;                                  // EOI before SOS
        if (p + 2 > n) return;
        const int len = be16(d + p);
        if (len < 2 || p + len > n) return;
        const uint8_t* s = d + p + 2;
        const int sl = len - 2;
        if (m == 0xDB) {                                        // DQT
            int o = 0;
            while (o < sl) {
                const int pq = s[o] >> 4, tq = s[o] & 15;
                ++o;
                if (tq > 3 || o + (pq ? 128 : 64) > sl) return;
                for (int i = 0; i < 64; ++i) {
                    pl->quant[tq][ZIGZAG[i]] = (uint16_t)(pq ? be16(s + o + 2 * i) : s[o + i]);
                }
                o += pq ? 128 : 64;
                have_q[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {                    // SOF0 / SOF1: sequential Huffman, 8 bit
            if (sl < 6 || s[0] != 8) { pl->status = JPEG_UNSUPPORTED; return; }
            pl->height = be16(s + 1); pl->width = be16(s + 3); pl->ncomp = s[5];
            if (pl->width <= 0 || pl->height <= 0) return;
            if ((pl->ncomp != 1 && pl->ncomp != 3) || sl < 6 + 3 * pl->ncomp) { pl->status = JPEG_UNSUPPORTED; return; }
            for (int c = 0; c < pl->ncomp; ++c) {
                comp_id[c] = s[6 + 3 * c];
                pl->hs[c] = s[7 + 3 * c] >> 4; pl->vs[c] = s[7 + 3 * c] & 15; pl->tq[c] = s[8 + 3 * c];
                if (pl->tq[c] > 3) return;
            }
            have_sof = true;
        } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            pl->status = m == 0xC2 ? JPEG_PROGRESSIVE : JPEG_UNSUPPORTED;   // progressive / lossless / arithmetic
            return;
        } else if (m == 0xC4) {                                 // DHT
            int o = 0;
            while (o + 17 <= sl) {
                const int tc = s[o] >> 4, th = s[o] & 15;
                uint8_t bits[17];
                bits[0] = 0;
                int cnt = 0;
                for (int i = 1; i <= 16; ++i) { bits[i] = s[o + i]; cnt += bits[i]; }
                o += 17;
                if (cnt > 256 || o + cnt > sl) return;
                if (tc > 1 || th > 1) { pl->status = JPEG_UNSUPPORTED; return; }
                if (!build_lut(bits, s + o, cnt, &pl->huff[tc * 2 + th])) return;
                have_h[tc * 2 + th] = true;
                o += cnt;
            }
        } else if (m == 0xDD) {                                 // DRI
            if (sl < 2) return;
            pl->restart_interval = be16(s);
        } else if (m == 0xE0) {
            if (sl >= 5 && !memcmp(s, "JFIF", 5)) jfif = true;
        } else if (m == 0xEE) {
            if (sl >= 12 && !memcmp(s, "Adobe", 5)) { adobe = true; adobe_transform = s[11]; }
        } else if (m == 0xDA) {                                 // SOS
            if (!have_sof || sl < 1) return;
            const int ns = s[0];
            if (ns != pl->ncomp || sl < 1 + 2 * ns + 3) { pl->status = JPEG_UNSUPPORTED; return; }   // one interleaved scan only
            for (int c = 0; c < ns; ++c) {
                if (s[1 + 2 * c] != comp_id[c]) { pl->status = JPEG_UNSUPPORTED; return; }
                pl->td[c] = s[2 + 2 * c] >> 4; pl->ta[c] = s[2 + 2 * c] & 15;
                if (pl->td[c] > 1 || pl->ta[c] > 1 || !have_h[pl->td[c]] || !have_h[2 + pl->ta[c]] || !have_q[pl->tq[c]]) return;
            }
            p += len;
            pl->scan_off = (int32_t)p;
            pl->scan_len = (int32_t)(n - p);
            break;
        }
        p += len;
    }
    // colour space (jdapimin.c default_decompress_parms): 3 components are YCbCr unless Adobe says RGB or the ids spell RGB
    if (pl->ncomp == 3) {
        bool ycc = true;
        if (!jfif && adobe) ycc = adobe_transform == 1;
        else if (!jfif && !adobe) ycc = !(comp_id[0] == 'R' && comp_id[1] == 'G' && comp_id[2] == 'B');
        if (!ycc) { pl->status = JPEG_UNSUPPORTED; return; }
        const bool chroma11 = pl->hs[1] == 1 && pl->vs[1] == 1 && pl->hs[2] == 1 && pl->vs[2] == 1;
        const bool ok = chroma11 && ((pl->hs[0] == 2 && pl->vs[0] == 2) || (pl->hs[0] == 2 && pl->vs[0] == 1) ||
                                     (pl->hs[0] == 1 && pl->vs[0] == 1));
        if (!ok) { pl->status = JPEG_UNSUPPORTED; return; }
    } else {
        pl->hs[0] = pl->vs[0] = 1;      // a single-component scan is never interleaved: MCU = one block
    }
    pl->hmax = pl->hs[0]; pl->vmax = pl->vs[0];
    pl->mcus_x = (pl->width + 8 * pl->hmax - 1) / (8 * pl->hmax);
    pl->mcus_y = (pl->height + 8 * pl->vmax - 1) / (8 * pl->vmax);
    pl->blocks_per_mcu = 0;
    for (int c = 0; c < pl->ncomp; ++c) {
        pl->blocks_per_mcu += pl->hs[c] * pl->vs[c];
        pl->plane_w[c] = pl->mcus_x * pl->hs[c] * 8;
        pl->plane_h[c] = pl->mcus_y * pl->vs[c] * 8;
        pl->comp_w[c] = (pl->width * pl->hs[c] + pl->hmax - 1) / pl->hmax;
        pl->comp_h[c] = (pl->height * pl->vs[c] + pl->vmax - 1) / pl->vmax;
    }
    pl->status = JPEG_OK;
}

// ---- entropy decoder -------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* base;     // file bytes
    int pos, end;            // next byte to fetch / one past the last byte
    uint64_t acc;            // bits, MSB first
    int nbits;
    bool marker;             // a marker (FF xx, xx != 00) was reached: feed zero bits like libjpeg

    __device__ __forceinline__ void fill() {      // nbits <= 32 on entry; afterwards >= 25 real or zero bits
        while (nbits <= 56) {
            uint32_t c = 0;
            if (!marker && pos < end) {
                c = base[pos];
                if (c == 0xFF) {
                    const uint32_t c2 = pos + 1 < end ? base[pos + 1] : 0xD9;
                    if (c2 == 0) pos += 2;                      // stuffed zero
                    else { marker = true; c = 0; }              // leave pos at the marker
                } else ++pos;
            }
            acc |= (uint64_t)c << (56 - nbits);
            nbits += 8;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(acc >> (64 - n)); }
    __device__ __forceinline__ void skip(int n) { acc <<= n; nbits -= n; }
    __device__ __forceinline__ int decode(const HuffLut& t) {
        if (nbits < 32) fill();
        const uint32_t e = t.look[peek(9)];
        if (e) { skip(e >> 8); return e & 255; }
        int l = 10;
        int code = (int)peek(10);
        while (code > t.maxcode[l]) { ++l; code = (int)peek(l); }
        if (l > 16) { skip(16); return 0; }                    // garbage input
        skip(l);
        return t.huffval[(code + t.valoff[l]) & 255];
    }
    __device__ __forceinline__ int receive_extend(int s) {     // s in 1..15
        if (nbits < 32) fill();
        const int v = (int)peek(s);
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    // RSTn: drop the padding bits, step over the marker, start clean
    __device__ __forceinline__ void restart() {
        acc = 0; nbits = 0;
        if (!marker) {                                          // the marker has not been fetched yet: find it
            while (pos + 1 < end && !(base[pos] == 0xFF && base[pos + 1] >= 0xD0 && base[pos + 1] <= 0xD7)) ++pos;
        }
        if (pos + 1 < end && base[pos] == 0xFF && base[pos + 1] >= 0xD0 && base[pos + 1] <= 0xD7) pos += 2;
        marker = false;
    }
};

// Sequential walk, one image per CTA (lane 0): used for files with restart intervals, whose RSTn handling
// follows jdhuff.c's process_restart.  Restart-free files go through jpeg_huffman_parallel_kernel below.
__global__ void __launch_bounds__(32)
jpeg_huffman_kernel(const uint8_t* __restrict__ files, const JpegPlan* __restrict__ plans, int16_t* __restrict__ coef) {
    __shared__ HuffLut s_h[4];
    const JpegPlan& pl = plans[blockIdx.x];
    if (pl.status != JPEG_OK || pl.restart_interval == 0) return;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(pl.huff);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_h);
        for (int i = threadIdx.x; i < (int)(sizeof(s_h) / 4); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    BitReader br{files + pl.file_off, pl.scan_off, pl.scan_off + pl.scan_len, 0, 0, false};
    int pred[3] = {0, 0, 0};
    int to_restart = pl.restart_interval;
    const int ncomp = pl.ncomp;
    for (int my = 0; my < pl.mcus_y; ++my)
        for (int mx = 0; mx < pl.mcus_x; ++mx) {
            if (pl.restart_interval) {
                if (to_restart == 0) { br.restart(); pred[0] = pred[1] = pred[2] = 0; to_restart = pl.restart_interval; }
                --to_restart;
            }
            for (int c = 0; c < ncomp; ++c) {
                const HuffLut& dc = s_h[pl.td[c]];
                const HuffLut& ac = s_h[2 + pl.ta[c]];
                const int hs = pl.hs[c], vs = pl.vs[c], bw = pl.plane_w[c] >> 3;
                for (int v = 0; v < vs; ++v)
                    for (int h = 0; h < hs; ++h) {
                        int16_t* blk = coef + pl.coef_off[c] + ((int64_t)(my * vs + v) * bw + (mx * hs + h)) * 64;
                        int s = br.decode(dc);
                        if (s) pred[c] += br.receive_extend(s & 15);
                        blk[0] = (int16_t)pred[c];
                        for (int k = 1; k < 64;) {
                            s = br.decode(ac);
                            const int r = s >> 4;
                            s &= 15;
                            if (s) {
                                k += r;
                                const int val = br.receive_extend(s);
                                if (k < 64) blk[k] = (int16_t)val;              // zig-zag order in memory (see jpeg_idct_kernel)
                                ++k;
                            } else {
                                if (r != 15) break;             // EOB
                                k += 16;
                            }
                        }
                    }
            }
        }
}


// ---- parallel entropy decoder (restart-free files) ---------------------------------------------------------
// A Huffman stream has no random access, but it self-synchronises: a decoder started at a wrong bit offset
// falls into step with the true symbol sequence after a while.  One CTA per image:
//   1. un-stuff the scan (FF 00 -> FF) into a clean byte stream, cut at the first marker (EOI);
//   2. cut it into <= 1024 sub-sequences; every thread decodes its sub-sequences from a guessed state
//      (bit 0 of the sub-sequence, block 0 of an MCU, DC symbol) and records the state it leaves with;
//   3. repeat: a sub-sequence whose predecessor's exit state differs from the state it started from is decoded
//      again from that exit state - the true state spreads from sub-sequence 0 and, thanks to
//      self-synchronisation, settles after a few rounds (at most #sub-sequences);
//   4. exclusive scan of the blocks completed per sub-sequence = the block index each one starts at; decode
//      once more, now writing the coefficients (DC as differences);
//   5. prefix-sum the DC differences per component in scan order.
// State = (bit position, block within the MCU, zig-zag index).
// 512 sub-sequences of >= 64 bytes per image, one per thread: measured best on B200 (1024 x 1024: more rounds of
// shorter sub-sequences, 2.9 ms per 256 images; 512 x 512: 2.6 ms; 256 x 256: 2.8 ms)
constexpr int JP_PAR_THREADS = 512, JP_MAX_SUBSEQ = 512;

struct CleanReader {
    const uint32_t* w;     // clean stream, 4-byte aligned, zero padded
    int bitpos;
    uint64_t acc;
    int have, next;
    __device__ __forceinline__ static uint32_t be(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
    __device__ __forceinline__ void init(int pos) {
        bitpos = pos;
        const int wi = pos >> 5, sh = pos & 31;
        acc = (((uint64_t)be(w[wi]) << 32) | be(w[wi + 1])) << sh;
        have = 64 - sh;
        next = wi + 2;
    }
    __device__ __forceinline__ void ensure() {
        if (have < 32) { acc |= (uint64_t)be(w[next++]) << (32 - have); have += 32; }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(acc >> (64 - n)); }
    __device__ __forceinline__ void skip(int n) { acc <<= n; have -= n; bitpos += n; }
    __device__ __forceinline__ int decode(const HuffLut& t) {
        ensure();
        const uint32_t e = t.look[peek(9)];
        if (e) { skip(e >> 8); return e & 255; }
        int l = 10;
        int code = (int)peek(10);
        while (code > t.maxcode[l]) { ++l; code = (int)peek(l); }
        if (l > 16) { skip(16); return 0; }
        skip(l);
        return t.huffval[(code + t.valoff[l]) & 255];
    }
    __device__ __forceinline__ int receive_extend(int s) {
        ensure();
        const int v = (int)peek(s);
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
};

struct McuMap {            // block-in-MCU -> component and position inside the MCU (chroma is always 1x1 here)
    int bpm, ny, hs0;      // blocks per MCU, luma blocks per MCU, luma blocks per MCU row
    __device__ __forceinline__ int comp(int j) const { return j < ny ? 0 : j - ny + 1; }
    __device__ __forceinline__ int bv(int j) const { return j < ny ? j / hs0 : 0; }
    __device__ __forceinline__ int bh(int j) const { return j < ny ? j % hs0 : 0; }
};

__device__ __forceinline__ uint64_t pack_state(int bitpos, int bi, int k) { return ((uint64_t)(uint32_t)bitpos << 16) | (uint32_t)(bi << 8) | (uint32_t)k; }

// Decodes from `state` until the bit position reaches end_bit; returns the exit state, adds completed blocks to
// nblocks.  WRITE: stores coefficients; `blk` is the scan-order index of the block the state is inside.
// (Measured and rejected, twice: staging each block in a per-thread shared-memory row and flushing it as 16-byte chunks
// up to the last non-zero coefficient instead of one scattered 2-byte store per coefficient - 0.95 -> 1.00 ms per 256
// images: the extra shared-memory traffic and the second code path for blocks that straddle two sub-sequences cost more
// than the saved L2 transactions.)
template <bool WRITE>
__device__ __forceinline__ uint64_t decode_subseq(const JpegPlan& pl, const HuffLut* s_h, const uint8_t* s_zz, const McuMap& mm,
                                                  const uint32_t* clean, uint64_t state, int end_bit, int total_bits,
                                                  int& nblocks, int blk, int16_t* __restrict__ coef) {
    CleanReader br;
    br.w = clean;
    br.init((int)(state >> 16));
    int bi = (int)(state >> 8) & 255, k = (int)state & 255;
    const int limit = min(end_bit, total_bits);
    // WRITE: position of the current block, advanced incrementally (one division per sub-sequence, not per block)
    int16_t* dst = nullptr;
    int mx = 0, my = 0;
    auto block_ptr = [&]() -> int16_t* {
        if (my >= pl.mcus_y) return nullptr;                                   // garbage beyond the image
        const int c = mm.comp(bi);
        return coef + pl.coef_off[c] + ((int64_t)(my * pl.vs[c] + mm.bv(bi)) * (pl.plane_w[c] >> 3) + (mx * pl.hs[c] + mm.bh(bi))) * 64;
    };
    if (WRITE) {
        const int mcu = blk / mm.bpm;                        // blk % bpm == bi for a true state
        my = mcu / pl.mcus_x; mx = mcu - my * pl.mcus_x;
        dst = block_ptr();
    }
    // table ids per block of the MCU, packed 2 bits each (DC id | AC id << 1), so the loop does not index arrays
    uint32_t tabs = 0;
    for (int j = 0; j < mm.bpm; ++j) { const int c = mm.comp(j); tabs |= (uint32_t)(pl.td[c] | (pl.ta[c] << 1)) << (2 * j); }
    while (br.bitpos < limit) {
        // one symbol: DC (k == 0) and AC share the path - a DC symbol is a run-0 symbol of the DC table
        const uint32_t tb = tabs >> (2 * bi);
        const HuffLut& t = k == 0 ? s_h[tb & 1] : s_h[2 + ((tb >> 1) & 1)];
        br.ensure();
        const uint32_t win = br.peek(32);                   // code (<= 16 bits) + value bits (<= 15) of the fast path fit
        const uint32_t e = t.look[win >> 23];
        int sym, nb;
        if (e) { nb = e >> 8; sym = e & 255; }
        else {
            nb = 10;
            int code = (int)(win >> 22);
            while (code > t.maxcode[nb]) { ++nb; code = (int)(win >> (32 - nb)); }
            sym = nb > 16 ? 0 : t.huffval[(code + t.valoff[nb]) & 255];
            nb = min(nb, 16);
        }
        const int r = k == 0 ? 0 : sym >> 4, sz = sym & 15;
        int v = 0;
        if (nb + sz <= 32) {
            if (sz) {
                v = (int)((win << nb) >> (32 - sz));
                v = v < (1 << (sz - 1)) ? v - (1 << sz) + 1 : v;
            }
            br.skip(nb + sz);
        } else {                                            // 16-bit code + long value: two steps
            br.skip(nb);
            v = br.receive_extend(sz);
        }
        if (sz) {
            k += r;
            if (WRITE && dst && k < 64) dst[k] = (int16_t)v;    // zig-zag order in memory (jpeg_idct_kernel undoes it)
            ++k;
        } else if (k == 0) k = 1;                           // zero DC difference
        else k = r == 15 ? k + 16 : 64;                     // ZRL / EOB
        if (k >= 64) {
            k = 0;
            ++nblocks;
            if (++bi == mm.bpm) {
                bi = 0;
                if (WRITE && ++mx == pl.mcus_x) { mx = 0; ++my; }
            }
            if (WRITE) dst = block_ptr();
        }
    }
    return pack_state(br.bitpos, bi, k);
}

__global__ void __launch_bounds__(JP_PAR_THREADS)
jpeg_huffman_parallel_kernel(const uint8_t* __restrict__ files, const JpegPlan* __restrict__ plans, uint8_t* __restrict__ clean_all,
                             int16_t* __restrict__ coef) {
    __shared__ HuffLut s_h[4];
    __shared__ uint64_t s_start[JP_MAX_SUBSEQ], s_exit[2][JP_MAX_SUBSEQ];
    __shared__ int s_cnt[JP_MAX_SUBSEQ];
    __shared__ int s_scan[JP_PAR_THREADS];
    __shared__ int s_marker;
    __shared__ uint8_t s_zz[64];
    const JpegPlan& pl = plans[blockIdx.x];
    if (pl.status != JPEG_OK || pl.restart_interval != 0) return;
    const int tid = threadIdx.x;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(pl.huff);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_h);
        for (int i = tid; i < (int)(sizeof(s_h) / 4); i += JP_PAR_THREADS) dst[i] = src[i];
    }
    if (tid < 64) s_zz[tid] = c_zigzag[tid];
    McuMap mm;
    mm.ny = pl.hs[0] * pl.vs[0]; mm.hs0 = pl.hs[0]; mm.bpm = mm.ny + (pl.ncomp == 3 ? 2 : 0);

    // ---- 1. un-stuff ------------------------------------------------------------------------------------
    const uint8_t* raw = files + pl.file_off + pl.scan_off;
    const int n = pl.scan_len;
    uint8_t* clean = clean_all + pl.file_off;               // file offsets are 16-byte aligned
    if (tid == 0) s_marker = n;
    __syncthreads();
    // 16 raw bytes per thread and tile, fetched as aligned words (plus one byte of context on either side);
    // a tile that contains the first marker (FF followed by anything but 00) ends the stream
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(raw) & ~(uintptr_t)3);
    const int mis = (int)(reinterpret_cast<uintptr_t>(raw) & 3);          // raw[i] = byte (i + mis) of rw
    int base_out = 0;
    for (int t0 = 0; t0 < n; t0 += JP_PAR_THREADS * 16) {
        const int lo = t0 + tid * 16;
        uint32_t w[6];                                                   // bytes lo-4 .. lo+19 relative to the word grid
        const int w0 = (lo + mis) >> 2;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int wi = w0 - 1 + j;
            w[j] = (wi >= 0 && wi * 4 < n + mis + 4 && lo < n + 16) ? __ldg(rw + wi) : 0u;
        }
        const int sh = (lo + mis) & 3;
        auto byte_at = [&](int j) -> uint32_t {                          // raw[lo + j], j in -1..16
            const int q = j + sh + 4;
            return (w[q >> 2] >> (8 * (q & 3))) & 255u;
        };
        uint32_t keepmask = 0;
        int first_marker = n;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int i = lo + j;
            const uint32_t b = byte_at(j), prev = byte_at(j - 1), nxt = byte_at(j + 1);
            if (i < n) {
                if (b == 0xFF && i + 1 < n && nxt != 0 && first_marker == n) first_marker = i;
                if (!(b == 0 && i > 0 && prev == 0xFF)) keepmask |= 1u << j;
            }
        }
        if (first_marker < n) atomicMin(&s_marker, first_marker);
        __syncthreads();
        const int m = s_marker;
        if (lo + 16 > m) keepmask &= m > lo ? (1u << (m - lo)) - 1u : 0u;   // nothing at or after the marker
        const int keep = __popc(keepmask);
        // block exclusive scan of `keep`
        int v = keep;
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, off); if ((tid & 31) >= off) v += o; }
        if ((tid & 31) == 31) s_scan[tid >> 5] = v;
        __syncthreads();
        if (tid < 32) {
            int ws = tid < JP_PAR_THREADS / 32 ? s_scan[tid] : 0;
            for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, ws, off); if (tid >= off) ws += o; }
            s_scan[32 + tid] = ws;
        }
        __syncthreads();
        int o = base_out + v - keep + ((tid >> 5) ? s_scan[32 + (tid >> 5) - 1] : 0);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (keepmask & (1u << j)) clean[o++] = (uint8_t)byte_at(j);
        base_out += s_scan[32 + JP_PAR_THREADS / 32 - 1];
        __syncthreads();
        if (m < t0 + JP_PAR_THREADS * 16) break;
    }
    const int nclean = base_out;
    for (int i = tid; i < 32; i += JP_PAR_THREADS) clean[nclean + i] = 0;      // zero padding for the word reader
    __syncthreads();
    const uint32_t* cw = reinterpret_cast<const uint32_t*>(clean);
    const int total_bits = nclean * 8;

    // ---- 2. first pass from guessed states -----------------------------------------------------------------
    int S = (nclean + JP_MAX_SUBSEQ - 1) / JP_MAX_SUBSEQ;
    S = max(64, (S + 3) & ~3);                               // bytes per sub-sequence
    const int nsub = max(1, (nclean + S - 1) / S);
    for (int i = tid; i < nsub; i += JP_PAR_THREADS) {
        const uint64_t st = pack_state(i * S * 8, 0, 0);
        int cnt = 0;
        s_start[i] = st;
        s_exit[0][i] = decode_subseq<false>(pl, s_h, s_zz, mm, cw, st, (i + 1) * S * 8, total_bits, cnt, 0, nullptr);
        s_cnt[i] = cnt;
    }
    __syncthreads();
    // ---- 3. propagate exit states until every sub-sequence started from its predecessor's exit ---------------
    int cur = 0;
    for (int round = 0; round < nsub; ++round) {
        int changed = 0;
        for (int i = tid; i < nsub; i += JP_PAR_THREADS) {
            uint64_t ex = s_exit[cur][i];
            if (i > 0) {
                const uint64_t inc = s_exit[cur][i - 1];
                if (inc != s_start[i]) {
                    int cnt = 0;
                    ex = decode_subseq<false>(pl, s_h, s_zz, mm, cw, inc, (i + 1) * S * 8, total_bits, cnt, 0, nullptr);
                    s_start[i] = inc;
                    s_cnt[i] = cnt;
                    changed = 1;
                }
            }
            s_exit[cur ^ 1][i] = ex;
        }
        cur ^= 1;
        if (!__syncthreads_or(changed)) break;
    }
    // ---- 4. block index of every sub-sequence (exclusive scan), then the writing pass --------------------------
    {
        const int per = (nsub + JP_PAR_THREADS - 1) / JP_PAR_THREADS;
        const int lo = tid * per, hi = min(lo + per, nsub);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += s_cnt[i];
        int v = sum;
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, off); if ((tid & 31) >= off) v += o; }
        if ((tid & 31) == 31) s_scan[tid >> 5] = v;
        __syncthreads();
        if (tid < 32) {
            int w = tid < JP_PAR_THREADS / 32 ? s_scan[tid] : 0;
            for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, off); if (tid >= off) w += o; }
            s_scan[32 + tid] = w;
        }
        __syncthreads();
        int run = v - sum + ((tid >> 5) ? s_scan[32 + (tid >> 5) - 1] : 0);
        for (int i = lo; i < hi; ++i) { const int c = s_cnt[i]; s_cnt[i] = run; run += c; }
        __syncthreads();
    }
    for (int i = tid; i < nsub; i += JP_PAR_THREADS) {
        int cnt = 0;
        decode_subseq<true>(pl, s_h, s_zz, mm, cw, s_start[i], (i + 1) * S * 8, total_bits, cnt, s_cnt[i], coef);
    }
    __syncthreads();
    // ---- 5. DC prediction: inclusive scan of the differences per component, scan order ---------------------------
    for (int c = 0; c < pl.ncomp; ++c) {
        const int bpc = pl.hs[c] * pl.vs[c], nblk = pl.mcus_x * pl.mcus_y * bpc, bw = pl.plane_w[c] >> 3;
        auto dc_ptr = [&](int t) -> int16_t* {
            const int mcu = t / bpc, j = t - mcu * bpc, my = mcu / pl.mcus_x, mx = mcu - my * pl.mcus_x;
            const int v = j / pl.hs[c], h = j - v * pl.hs[c];
            return coef + pl.coef_off[c] + ((int64_t)(my * pl.vs[c] + v) * bw + (mx * pl.hs[c] + h)) * 64;
        };
        const int per = (nblk + JP_PAR_THREADS - 1) / JP_PAR_THREADS;
        const int lo = tid * per, hi = min(lo + per, nblk);
        int sum = 0;
        for (int t = lo; t < hi; ++t) sum += *dc_ptr(t);
        int v = sum;
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, v, off); if ((tid & 31) >= off) v += o; }
        if ((tid & 31) == 31) s_scan[tid >> 5] = v;
        __syncthreads();
        if (tid < 32) {
            int w = tid < JP_PAR_THREADS / 32 ? s_scan[tid] : 0;
            for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, off); if (tid >= off) w += o; }
            s_scan[32 + tid] = w;
        }
        __syncthreads();
        int run = v - sum + ((tid >> 5) ? s_scan[32 + (tid >> 5) - 1] : 0);
        for (int t = lo; t < hi; ++t) { int16_t* p = dc_ptr(t); run += *p; *p = (int16_t)run; }
        __syncthreads();
    }
}

// ---- dequantise + IDCT: one thread per 8x8 block of any component --------------------------------------
#ifndef JPEG_IDCT_MIN_CTAS
#define JPEG_IDCT_MIN_CTAS 6    // 80 registers (96 bytes of spills) and 6 CTAs per SM: 165 us instead of 184 us per 256 images (8: 194 us)
#endif
__global__ void __launch_bounds__(128, JPEG_IDCT_MIN_CTAS)
jpeg_idct_kernel(const JpegPlan* __restrict__ plans, const int16_t* __restrict__ coef, uint8_t* __restrict__ planes) {
    const JpegPlan& pl = plans[blockIdx.y];
    if (pl.status != JPEG_OK) return;
    int nblk[3], total = 0;
    for (int c = 0; c < pl.ncomp; ++c) { nblk[c] = (pl.plane_w[c] >> 3) * (pl.plane_h[c] >> 3); total += nblk[c]; }
    for (int t = blockIdx.x * 128 + threadIdx.x; t < total; t += gridDim.x * 128) {
        int c = 0, b = t;
        while (b >= nblk[c]) { b -= nblk[c]; ++c; }
        const int bw = pl.plane_w[c] >> 3;
        const int16_t* src = coef + pl.coef_off[c] + (int64_t)b * 64;
        const uint16_t* q = pl.quant[pl.tq[c]];
        // the entropy decoders store a block in zig-zag order (its non-zero coefficients then sit in the first one or two
        // 32-byte sectors instead of four); the fully unrolled loop turns the permutation into register names
        constexpr uint8_t ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        int d[64];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + 8 * k);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                d[ZZ[8 * k + 2 * j]] = (int)(int16_t)(w[j] & 0xFFFF) * (int)q[ZZ[8 * k + 2 * j]];
                d[ZZ[8 * k + 2 * j + 1]] = (int)(int16_t)(w[j] >> 16) * (int)q[ZZ[8 * k + 2 * j + 1]];
            }
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) COL8(idct8<false>, d, cc);
#pragma unroll
        for (int r = 0; r < 8; ++r) ROWS8(idct8<true>, (d + 8 * r));
        const int pitch = pl.plane_w[c];
        uint8_t* base = planes + pl.plane_off[c] + (int64_t)(b / bw) * 8 * pitch + (b % bw) * 8;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint2 v;
            v.x = range_limit(d[r * 8]) | (range_limit(d[r * 8 + 1]) << 8) | (range_limit(d[r * 8 + 2]) << 16) | (range_limit(d[r * 8 + 3]) << 24);
            v.y = range_limit(d[r * 8 + 4]) | (range_limit(d[r * 8 + 5]) << 8) | (range_limit(d[r * 8 + 6]) << 16) | (range_limit(d[r * 8 + 7]) << 24);
            *reinterpret_cast<uint2*>(base + (int64_t)r * pitch) = v;
        }
    }
}

// h2v1 fancy up-sampling (jdsample.c): 3/4 nearer + 1/4 further, rounding 1 (even) / 2 (odd output column)
__device__ __forceinline__ int up_h2v1(const uint8_t* __restrict__ C, int pitch, int cw, int y, int x) {
    const int cx = x >> 1;
    const int nx = min(max((x & 1) ? cx + 1 : cx - 1, 0), cw - 1);
    const int cur = C[(size_t)y * pitch + cx];
    if (nx == cx) return cur;                                   // first / last column are copied
    return (3 * cur + C[(size_t)y * pitch + nx] + ((x & 1) ? 2 : 1)) >> 2;
}

// One thread per 4 horizontally adjacent pixels: luma as one 32-bit load, 12 output bytes as three 32-bit stores
// (output rows are 16-byte aligned: out_off % 256 == 0, out_pitch % 16 == 0).
__global__ void __launch_bounds__(256)
jpeg_color_kernel(const JpegPlan* __restrict__ plans, const uint8_t* __restrict__ planes, uint8_t* __restrict__ out, int bgr) {
    const JpegPlan& pl = plans[blockIdx.y];
    if (pl.status != JPEG_OK) return;
    const int W = pl.width, H = pl.height;
    const uint8_t* Y = planes + pl.plane_off[0];
    const uint8_t* Cb = planes + pl.plane_off[1];
    const uint8_t* Cr = planes + pl.plane_off[2];
    uint8_t* dst = out + pl.out_off;
    const int mode = pl.ncomp == 1 ? 0 : (pl.hs[0] == 2 ? (pl.vs[0] == 2 ? 3 : 2) : 1);   // gray, 4:4:4, 4:2:2, 4:2:0
    const int qw = (W + 3) >> 2;                              // groups of 4 pixels per row
    const int pw0 = pl.plane_w[0], pw1 = pl.plane_w[1], ch = pl.comp_h[1], cw = pl.comp_w[1];
    const int64_t ngrp = (int64_t)H * qw;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < ngrp; t += (int64_t)gridDim.x * 256) {
        const int y = (int)((uint32_t)t / (uint32_t)qw), x0 = (int)((uint32_t)t - (uint32_t)y * (uint32_t)qw) * 4;
        const uint32_t y4 = *reinterpret_cast<const uint32_t*>(Y + (size_t)y * pw0 + x0);      // plane width is a multiple of 8
        uint8_t px[12];
        if (mode == 3 && x0 + 4 <= W) {
            // 4:2:0, whole group inside the image: the 4 pixels share chroma columns cx0-1 .. cx0+2 of rows cy and ny, so
            // each plane costs 6 loads (two of them 16-bit) instead of 16, and the vertical sums 3*near + far are formed
            // once per column.  Same integers as up_h2v2 (clamped neighbour index == libjpeg's first / last column rule).
            const int cy = y >> 1, nyr = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
            const int cx0 = x0 >> 1, im1 = max(cx0 - 1, 0), i2 = min(cx0 + 2, cw - 1);
            int cbv[4], crv[4];
            {
                const uint8_t* a = Cb + (size_t)cy * pw1;
                const uint8_t* b = Cb + (size_t)nyr * pw1;
                const uint32_t a01 = *reinterpret_cast<const uint16_t*>(a + cx0), b01 = *reinterpret_cast<const uint16_t*>(b + cx0);
                cbv[0] = 3 * a[im1] + b[im1]; cbv[1] = 3 * (int)(a01 & 255u) + (int)(b01 & 255u);
                cbv[2] = 3 * (int)(a01 >> 8) + (int)(b01 >> 8); cbv[3] = 3 * a[i2] + b[i2];
            }
            {
                const uint8_t* a = Cr + (size_t)cy * pw1;
                const uint8_t* b = Cr + (size_t)nyr * pw1;
                const uint32_t a01 = *reinterpret_cast<const uint16_t*>(a + cx0), b01 = *reinterpret_cast<const uint16_t*>(b + cx0);
                crv[0] = 3 * a[im1] + b[im1]; crv[1] = 3 * (int)(a01 & 255u) + (int)(b01 & 255u);
                crv[2] = 3 * (int)(a01 >> 8) + (int)(b01 >> 8); crv[3] = 3 * a[i2] + b[i2];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cur = 1 + (j >> 1), nbr = (j & 1) ? cur + 1 : cur - 1, bias = (j & 1) ? 7 : 8;
                const int cb = ((3 * cbv[cur] + cbv[nbr] + bias) >> 4) - 128, cr = ((3 * crv[cur] + crv[nbr] + bias) >> 4) - 128;
                const int yv = (y4 >> (8 * j)) & 255;
                const int r = max(0, min(255, yv + ((91881 * cr + 32768) >> 16)));
                const int g = max(0, min(255, yv + ((-22554 * cb + 32768 - 46802 * cr) >> 16)));
                const int b = max(0, min(255, yv + ((116130 * cb + 32768) >> 16)));
                px[3 * j] = (uint8_t)(bgr ? b : r); px[3 * j + 1] = (uint8_t)g; px[3 * j + 2] = (uint8_t)(bgr ? r : b);
            }
        } else
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = min(x0 + j, W - 1);                 // the last group of a row may run past W: recompute the last pixel
            const int yy = (y4 >> (8 * j)) & 255;
            int r = yy, g = yy, b = yy;
            if (mode != 0) {
                int cb, cr;
                if (mode == 3) {
                    cb = up_h2v2(Cb, pw1, ch, cw, y, x);
                    cr = up_h2v2(Cr, pw1, ch, cw, y, x);
                } else if (mode == 2) {
                    cb = up_h2v1(Cb, pw1, cw, y, x);
                    cr = up_h2v1(Cr, pw1, cw, y, x);
                } else {
                    cb = Cb[(size_t)y * pw1 + x];
                    cr = Cr[(size_t)y * pw1 + x];
                }
                cb -= 128; cr -= 128;
                const int yv = x0 + j < W ? yy : (int)Y[(size_t)y * pw0 + x];
                r = max(0, min(255, yv + ((91881 * cr + 32768) >> 16)));
                g = max(0, min(255, yv + ((-22554 * cb + 32768 - 46802 * cr) >> 16)));
                b = max(0, min(255, yv + ((116130 * cb + 32768) >> 16)));
            }
            px[3 * j] = (uint8_t)(bgr ? b : r); px[3 * j + 1] = (uint8_t)g; px[3 * j + 2] = (uint8_t)(bgr ? r : b);
        }
        uint8_t* o = dst + (int64_t)y * pl.out_pitch + 3 * x0;
        if (x0 + 4 <= W) {
            uint32_t* o32 = reinterpret_cast<uint32_t*>(o);    // 3*x0 is a multiple of 4
            o32[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
            o32[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
            o32[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
        } else {
            for (int j = 0; j < 3 * (W - x0); ++j) o[j] = px[j];
        }
    }
}

}  // namespace advmix

using namespace advmix;

extern "C" {

size_t advmix_jpeg_plan_stride(void) { return sizeof(JpegPlan); }

int advmix_jpeg_plan_h(const uint8_t* files_h, const int64_t* off_h, const int64_t* len_h, int B, void* plans_h,
                       int64_t* out_bytes, int64_t* coef_elems, int64_t* plane_bytes) {
    ADVMIX_REQUIRE(B >= 0, "jpeg_plan: bad B");
    ADVMIX_REQUIRE(B == 0 || (files_h && off_h && len_h && plans_h), "jpeg_plan: null argument");
    JpegPlan* pl = reinterpret_cast<JpegPlan*>(plans_h);
    int64_t out = 0, ce = 0, pb = 0;
    int bad = 0;
    for (int b = 0; b < B; ++b) {
        JpegPlan& p = pl[b];
        memset(&p, 0, sizeof(p));
        p.file_off = off_h[b]; p.file_len = len_h[b];
        if (len_h[b] > 0x7fffffff) { p.status = JPEG_UNSUPPORTED; ++bad; continue; }
        parse_one(files_h + off_h[b], len_h[b], &p);
        if (p.status != JPEG_OK) { ++bad; continue; }
        p.out_pitch = ((int64_t)p.width * 3 + 15) & ~(int64_t)15;
        p.out_off = out;
        out += (p.out_pitch * p.height + 255) & ~(int64_t)255;
        for (int c = 0; c < p.ncomp; ++c) {
            const int64_t px = (int64_t)p.plane_w[c] * p.plane_h[c];
            p.coef_off[c] = ce; ce += px;
            p.plane_off[c] = pb; pb += (px + 15) & ~(int64_t)15;
        }
    }
    if (out_bytes) *out_bytes = out;
    if (coef_elems) *coef_elems = ce;
    if (plane_bytes) *plane_bytes = pb;
    return bad ? fail(ADVMIX_ERR_UNSUPPORTED, "jpeg_plan: %d of %d files are not baseline YCbCr/gray JPEGs this decoder handles "
                                              "(see the per-image status field)", bad, B)
               : ADVMIX_OK;
}

int advmix_jpeg_decode(const uint8_t* files, const void* plans, int B, int max_blocks, int max_pixels, uint8_t* out,
                       void* workspace, size_t ws_bytes, int64_t coef_elems, int64_t plane_bytes, int64_t files_bytes,
                       int any_restart, int bgr, advmix_stream_t stream) {
    ADVMIX_REQUIRE(B >= 0, "jpeg_decode: bad B");
    if (B == 0) return ADVMIX_OK;
    ADVMIX_REQUIRE(files && plans && out && workspace, "jpeg_decode: null argument");
    ADVMIX_REQUIRE(B <= 65535, "jpeg_decode: B <= 65535 per call");
    const size_t coef_b = ((size_t)coef_elems * 2 + 255) & ~(size_t)255;
    const size_t plane_b = ((size_t)plane_bytes + 255) & ~(size_t)255;
    const size_t need = coef_b + plane_b + (size_t)files_bytes + 64;
    if (ws_bytes < need) return fail(ADVMIX_ERR_WORKSPACE, "jpeg_decode: workspace %zu < %zu bytes", ws_bytes, need);
    cudaStream_t st = as_stream(stream);
    if (first_use_on_device(&c_zigzag)) ADVMIX_CUDA_OK(cudaMemcpyToSymbol(c_zigzag, ZIGZAG, 64));   // __constant__ symbols are per device
    int16_t* coef = reinterpret_cast<int16_t*>(workspace);
    uint8_t* planes = reinterpret_cast<uint8_t*>(workspace) + coef_b;
    const JpegPlan* pl = reinterpret_cast<const JpegPlan*>(plans);
    ADVMIX_CUDA_OK(cudaMemsetAsync(coef, 0, (size_t)coef_elems * 2, st));
    uint8_t* clean = planes + plane_b;                       // un-stuffed scans, same offsets as the files
    jpeg_huffman_parallel_kernel<<<B, JP_PAR_THREADS, 0, st>>>(files, pl, clean, coef);
    ADVMIX_LAUNCH_OK();
    if (any_restart) {
        jpeg_huffman_kernel<<<B, 32, 0, st>>>(files, pl, coef);
        ADVMIX_LAUNCH_OK();
    }
    const int cap = std::max(1, (sm_count() * 16 + B - 1) / B);
    jpeg_idct_kernel<<<dim3(std::min(ceil_div(max_blocks, 128), cap), B), 128, 0, st>>>(pl, coef, planes);
    ADVMIX_LAUNCH_OK();
    jpeg_color_kernel<<<dim3(std::min(ceil_div(max_pixels / 4 + 1, 256), cap), B), 256, 0, st>>>(pl, planes, out, bgr);
    ADVMIX_LAUNCH_OK();
    return ADVMIX_OK;
}

}  // extern "C"
