// K8: PNG decode on sm_100a, one warp per image:  chunk walk -> zlib/DEFLATE inflate (stored, fixed and dynamic Huffman,
// RFC 1950/1951) -> scanline un-filter (None/Sub/Up/Average/Paeth, PNG spec section 9) -> [H, W, C] samples in file
// (RGB) order, native endianness.
//
// Replaces `cv2.imdecode(np.frombuffer(value), IMREAD_UNCHANGED)` + the BGR->RGB reorder of
// CompressedImageCodec.decode (petastorm/codecs.py:102-116): cv2 returns BGR and the codec flips it back, so the net
// result is the PNG's own RGB order, which is what this kernel writes.
//
// Work split inside the warp: the DEFLATE bit stream is inherently serial, so lane 0 decodes symbols (canonical
// Huffman decode from per-warp tables in shared memory); stored blocks and the final sample copy are warp-wide vector
// copies; Up/None filters are lane-parallel, Sub/Average/Paeth run one lane per byte-channel.  Parallelism comes from
// the number of images: a row-group holds 10^2..10^4 images and every SM keeps 16+ image-warps resident.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"
#include "dev_util.cuh"
#include "kernels.h"

namespace pst {

namespace {

constexpr int kMaxBits = 15;
constexpr int kMaxLCodes = 286;
constexpr int kMaxDCodes = 30;
constexpr int kFixLCodes = 288;

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                                        35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                         1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                         9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Huff {
    uint16_t count[kMaxBits + 1];
    uint16_t symbol[kFixLCodes];
};
struct HuffD {
    uint16_t count[kMaxBits + 1];
    uint16_t symbol[kMaxDCodes];
};

constexpr int kLutL = 10;      // primary bits of the literal/length lookup table
constexpr int kLutD = 8;       // primary bits of the distance lookup table

struct WarpState {
    Huff lencode;
    HuffD distcode;
    // one-step decode tables: entry = (symbol << 4) | code length for codes of <= kLut* bits (index = the next stream
    // bits, LSB first); 0 = longer code, decoded bit by bit from the canonical tables above
    uint16_t lut_l[1 << kLutL];
    uint16_t lut_d[1 << kLutD];
    uint16_t lengths[kMaxLCodes + kMaxDCodes + 2];
    uint8_t palette[256 * 3];
    // broadcast slots
    int32_t err;
    int32_t have_plte;
};

// IDAT-spanning byte source: the zlib stream is the concatenation of all IDAT chunk payloads
struct Src {
    const uint8_t *p;        // next byte
    const uint8_t *seg_end;  // end of current IDAT payload
    const uint8_t *blob_end;
    bool eof;
};

__device__ __forceinline__ uint32_t be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

// advance to the next IDAT payload; sets eof when none is left
__device__ void src_next_segment(Src &s) {
    const uint8_t *q = s.seg_end + 4;  // skip CRC of the current chunk
    for (;;) {
        if (q + 8 > s.blob_end) { s.eof = true; return; }
        uint32_t len = be32(q);
        uint32_t typ = be32(q + 4);
        const uint8_t *data = q + 8;
        if (data + len + 4 > s.blob_end) { s.eof = true; return; }
        if (typ == 0x49444154u) {  // IDAT
            s.p = data;
            s.seg_end = data + len;
            if (len == 0) { q = data + 4; continue; }
            return;
        }
        if (typ == 0x49454e44u) { s.eof = true; return; }  // IEND
        q = data + len + 4;
    }
}
__device__ __forceinline__ int src_byte(Src &s) {
    if (s.p == s.seg_end) {
        if (s.eof) return -1;
        src_next_segment(s);
        if (s.eof) return -1;
    }
    return *s.p++;
}

struct Bits {
    uint64_t buf;
    int cnt;
};
__device__ __forceinline__ bool bits_need(Bits &b, Src &s, int n) {
    while (b.cnt < n) {
        int c = src_byte(s);
        if (c < 0) return false;
        b.buf |= (uint64_t)c << b.cnt;
        b.cnt += 8;
    }
    return true;
}
__device__ __forceinline__ int bits_get(Bits &b, Src &s, int n) {
    if (n == 0) return 0;
    if (!bits_need(b, s, n)) return -1;
    int v = (int)(b.buf & ((1ull << n) - 1));
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

// fills the bit buffer from the current IDAT payload without the per-byte segment checks (up to 7 bytes at once)
__device__ __forceinline__ void bits_refill_fast(Bits &b, Src &s) {
    while (b.cnt <= 56 && s.p < s.seg_end) {
        b.buf |= (uint64_t)(*s.p++) << b.cnt;
        b.cnt += 8;
    }
}

// one-step tables from the canonical (count / symbol) representation: canonical codes are handed out in order of
// (length, symbol); DEFLATE packs them MSB first into an LSB-first bit stream, hence the bit reversal
template <typename H>
__device__ void lut_build(const H &h, uint16_t *lut, int primary) {
    for (int i = 0; i < (1 << primary); i++) lut[i] = 0;
    int code = 0, index = 0;
    for (int len = 1; len <= primary; len++) {
        const int count = h.count[len];
        for (int k = 0; k < count; k++) {
            const uint32_t rev = __brev((uint32_t)(code + k)) >> (32 - len);
            const uint16_t e = (uint16_t)((h.symbol[index + k] << 4) | len);
            for (uint32_t j = rev; j < (1u << primary); j += 1u << len) lut[j] = e;
        }
        code = (code + count) << 1;
        index += count;
    }
}

// canonical Huffman decode, one bit at a time (count/symbol representation)
template <typename H>
__device__ int huff_decode(Bits &b, Src &s, const H &h) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= kMaxBits; len++) {
        if (b.cnt == 0 && !bits_need(b, s, 1)) return -1;
        code |= (int)(b.buf & 1);
        b.buf >>= 1;
        b.cnt--;
        int count = h.count[len];
        if (code - count < first) return h.symbol[index + (code - first)];
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -2;
}

// build count/symbol tables from code lengths; returns <0 if over-subscribed, >0 if incomplete, 0 if complete
template <typename H>
__device__ int huff_build(H &h, const uint16_t *length, int n) {
    for (int len = 0; len <= kMaxBits; len++) h.count[len] = 0;
    for (int sym = 0; sym < n; sym++) h.count[length[sym]]++;
    if (h.count[0] == n) return 0;
    int left = 1;
    for (int len = 1; len <= kMaxBits; len++) {
        left <<= 1;
        left -= h.count[len];
        if (left < 0) return left;
    }
    uint16_t offs[kMaxBits + 1];
    offs[1] = 0;
    for (int len = 1; len < kMaxBits; len++) offs[len + 1] = offs[len] + h.count[len];
    for (int sym = 0; sym < n; sym++)
        if (length[sym] != 0) h.symbol[offs[length[sym]]++] = (uint16_t)sym;
    return left;
}

// table-driven decode: one lookup for codes of <= `primary` bits, the bit-serial walk for the rare longer ones
template <typename H>
__device__ __forceinline__ int huff_decode_lut(Bits &b, Src &s, const H &h, const uint16_t *lut, int primary) {
    if (b.cnt < kMaxBits) bits_refill_fast(b, s);
    const uint16_t e = lut[b.buf & ((1u << primary) - 1)];
    const int len = e & 15;
    if (e != 0 && len <= b.cnt) {
        b.buf >>= len;
        b.cnt -= len;
        return e >> 4;
    }
    return huff_decode(b, s, h);      // long code, or the tail of a segment (bits_need crosses IDAT boundaries)
}

__device__ __forceinline__ int paeth(int a, int b, int c) {
    int p = a + b - c;
    int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

// One Huffman-coded block (type 1 fixed / type 2 dynamic), decoded by a single lane: table set-up, then
// literal/length + distance symbols through the lookup tables until end-of-block.  Returns 0 or 1 (corrupt).
__device__ int deflate_huffman_block(Bits &b, Src &s, uint8_t *raw, int64_t raw_total, WarpState &ws, int64_t &op, int type) {
    int err = 0;
    do {
        if (type == 1) {
            int sym = 0;
            for (; sym < 144; sym++) ws.lengths[sym] = 8;
            for (; sym < 256; sym++) ws.lengths[sym] = 9;
            for (; sym < 280; sym++) ws.lengths[sym] = 7;
            for (; sym < kFixLCodes; sym++) ws.lengths[sym] = 8;
            huff_build(ws.lencode, ws.lengths, kFixLCodes);
            for (sym = 0; sym < kMaxDCodes; sym++) ws.lengths[sym] = 5;
            huff_build(ws.distcode, ws.lengths, kMaxDCodes);
        } else {
            int nlen = bits_get(b, s, 5), ndist = bits_get(b, s, 5), ncode = bits_get(b, s, 4);
            if (nlen < 0 || ndist < 0 || ncode < 0) { err = 1; break; }
            nlen += 257; ndist += 1; ncode += 4;
            if (nlen > kMaxLCodes || ndist > kMaxDCodes) { err = 1; break; }
            int idx = 0;
            for (; idx < ncode; idx++) {
                int v = bits_get(b, s, 3);
                if (v < 0) { err = 1; break; }
                ws.lengths[c_clen_order[idx]] = (uint16_t)v;
            }
            if (err) break;
            for (; idx < 19; idx++) ws.lengths[c_clen_order[idx]] = 0;
            if (huff_build(ws.lencode, ws.lengths, 19) != 0) { err = 1; break; }
            idx = 0;
            while (idx < nlen + ndist) {
                int sym = huff_decode(b, s, ws.lencode);
                if (sym < 0) { err = 1; break; }
                if (sym < 16) ws.lengths[idx++] = (uint16_t)sym;
                else {
                    int len = 0, rep;
                    if (sym == 16) {
                        if (idx == 0) { err = 1; break; }
                        len = ws.lengths[idx - 1];
                        rep = bits_get(b, s, 2);
                        if (rep < 0) { err = 1; break; }
                        rep += 3;
                    } else if (sym == 17) {
                        rep = bits_get(b, s, 3);
                        if (rep < 0) { err = 1; break; }
                        rep += 3;
                    } else {
                        rep = bits_get(b, s, 7);
                        if (rep < 0) { err = 1; break; }
                        rep += 11;
                    }
                    if (idx + rep > nlen + ndist) { err = 1; break; }
                    while (rep--) ws.lengths[idx++] = (uint16_t)len;
                }
            }
            if (err) break;
            if (ws.lengths[256] == 0) { err = 1; break; }
            int e1 = huff_build(ws.lencode, ws.lengths, nlen);
            if (e1 < 0 || (e1 > 0 && nlen - ws.lencode.count[0] != 1)) { err = 1; break; }
            int e2 = huff_build(ws.distcode, ws.lengths + nlen, ndist);
            if (e2 < 0 || (e2 > 0 && ndist - ws.distcode.count[0] != 1)) { err = 1; break; }
        }
        lut_build(ws.lencode, ws.lut_l, kLutL);
        lut_build(ws.distcode, ws.lut_d, kLutD);
        // literal/length + distance codes
        for (;;) {
            int sym = huff_decode_lut(b, s, ws.lencode, ws.lut_l, kLutL);
            if (sym < 0) { err = 1; break; }
            if (sym < 256) {
                if (op >= raw_total) { err = 1; break; }
                raw[op++] = (uint8_t)sym;
            } else if (sym == 256) {
                break;
            } else {
                sym -= 257;
                if (sym >= 29) { err = 1; break; }
                int eb = bits_get(b, s, c_len_extra[sym]);
                if (eb < 0) { err = 1; break; }
                int len = c_len_base[sym] + eb;
                int ds = huff_decode_lut(b, s, ws.distcode, ws.lut_d, kLutD);
                if (ds < 0 || ds >= 30) { err = 1; break; }
                int de = bits_get(b, s, c_dist_extra[ds]);
                if (de < 0) { err = 1; break; }
                int64_t dist = (int64_t)c_dist_base[ds] + de;
                if (dist > op || op + len > raw_total) { err = 1; break; }
                for (int i = 0; i < len; i++, op++) raw[op] = raw[op - dist];
            }
        }
    } while (0);
    return err;
}

__device__ __forceinline__ const uint8_t *shfl_ptr(const uint8_t *p, int src) {
    const uint64_t v = (uint64_t)(uintptr_t)p;
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return (const uint8_t *)(uintptr_t)(((uint64_t)hi << 32) | lo);
}
__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)((uint64_t)v >> 32), src);
    return (int64_t)(((uint64_t)hi << 32) | lo);
}

// DEFLATE blocks (RFC 1951) from the bit reader into raw[op ..), at most raw_total bytes; returns 0 or 1 (corrupt),
// uniformly across the warp.  Called by ALL 32 lanes of a warp: lane 0 owns the bit reader, the source cursor and `op`
// (the other lanes' copies are ignored); the bit stream itself is serial, but
//   * stored blocks (what zlib emits for incompressible data, e.g. every scanline of a noise image) are moved by the
//     whole warp with 16-byte vector copies, and
//   * Huffman blocks are decoded by lane 0 through one-step lookup tables (10 bits literal/length, 8 bits distance).
// `ws` holds the tables of the warp.  The source must lie in a buffer with 16 bytes of slack on both sides (the arena).
__device__ int deflate_blocks(Bits &b, Src &s, uint8_t *raw, int64_t raw_total, WarpState &ws, int64_t &op, int lane) {
    int err = 0;
    int last = 0;
    while (!err && !last) {
        int type = 0;
        if (lane == 0) {
            last = bits_get(b, s, 1);
            type = bits_get(b, s, 2);
            if (last < 0 || type < 0) err = 1;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        type = __shfl_sync(0xffffffffu, type, 0);
        err = __shfl_sync(0xffffffffu, err, 0);
        if (err) break;
        if (type == 0) {
            // stored: drop the bits up to the byte boundary, LEN / NLEN, then raw bytes
            int len = 0;
            if (lane == 0) {
                const int drop = b.cnt & 7;
                b.buf >>= drop;
                b.cnt -= drop;
                const int l = bits_get(b, s, 16), nl = bits_get(b, s, 16);
                if (l < 0 || nl < 0 || (l ^ 0xffff) != nl || op + l > raw_total) err = 1;
                else {
                    len = l;
                    // the table-driven decoder reads ahead: whole bytes still in the bit buffer are the first bytes of
                    // the block; what follows comes straight from the source
                    while (len > 0 && b.cnt >= 8) {
                        raw[op++] = (uint8_t)(b.buf & 0xff);
                        b.buf >>= 8;
                        b.cnt -= 8;
                        len--;
                    }
                }
            }
            err = __shfl_sync(0xffffffffu, err, 0);
            if (err) break;
            len = __shfl_sync(0xffffffffu, len, 0);
            while (len > 0) {
                const uint8_t *p = nullptr;
                int take = 0;
                if (lane == 0) {
                    if (s.p == s.seg_end) {
                        src_next_segment(s);
                        if (s.eof) err = 1;
                    }
                    if (!err) {
                        const int64_t avail = s.seg_end - s.p;
                        take = avail > len ? len : (int)avail;
                        p = s.p;
                        s.p += take;
                    }
                }
                err = __shfl_sync(0xffffffffu, err, 0);
                if (err) break;
                take = __shfl_sync(0xffffffffu, take, 0);
                p = shfl_ptr(p, 0);
                const int64_t o = shfl_i64(op, 0);
                coop_copy(raw + o, p, take, lane, 32);
                if (lane == 0) op += take;
                len -= take;
            }
            __syncwarp();          // later back-references of lane 0 may read what the other lanes just wrote
            continue;
        }
        if (type == 3) { err = 1; break; }
        if (lane == 0) {
            err = deflate_huffman_block(b, s, raw, raw_total, ws, op, type);
        }
        __syncwarp();
        err = __shfl_sync(0xffffffffu, err, 0);
    }
    return err;
}

// PNG filter type 1 (Sub): recon[i] = filt[i] + recon[i - bpp] (mod 256) is a prefix sum per byte channel.  Every lane
// scans its own run of 8 pixels, a warp scan adds the totals of the lanes in front of it, and a running carry links the
// 256-pixel chunks of a scanline - instead of one lane per channel walking the whole line.
template <int BPP>
__device__ __forceinline__ void unfilter_sub(uint8_t *cur, int64_t nbytes, int lane) {
    constexpr int K = 8;
    const int64_t npix = nbytes / BPP;
    uint32_t carry[BPP];
#pragma unroll
    for (int c = 0; c < BPP; c++) carry[c] = 0;
    for (int64_t p0 = 0; p0 < npix; p0 += 32 * K) {
        const int64_t my0 = p0 + (int64_t)lane * K;
        const int cnt = (int)max((int64_t)0, min((int64_t)K, npix - my0));
        uint8_t vals[K * BPP];
        uint32_t loc[BPP];
#pragma unroll
        for (int c = 0; c < BPP; c++) loc[c] = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
#pragma unroll
            for (int c = 0; c < BPP; c++) {
                if (k < cnt) loc[c] = (loc[c] + cur[(my0 + k) * BPP + c]) & 255u;
                vals[k * BPP + c] = (uint8_t)loc[c];
            }
        }
#pragma unroll
        for (int c = 0; c < BPP; c++) {
            uint32_t incl = loc[c];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t add = (carry[c] + incl - loc[c]) & 255u;
            carry[c] = (carry[c] + __shfl_sync(0xffffffffu, incl, 31)) & 255u;
#pragma unroll
            for (int k = 0; k < K; k++)
                if (k < cnt) cur[(my0 + k) * BPP + c] = (uint8_t)((vals[k * BPP + c] + add) & 255u);
        }
    }
}

constexpr int kPngWarpsPerBlock = 2;

__global__ void __launch_bounds__(32 * kPngWarpsPerBlock)
k_png_batch(const uint8_t *__restrict__ base, const int64_t *__restrict__ offs, const int32_t *__restrict__ lens,
            const int64_t *__restrict__ row_idx, int64_t n, int height, int width, int channels, int sample_bytes,
            uint8_t *__restrict__ dst, uint8_t *__restrict__ work, int64_t work_per_image, int32_t *status) {
    __shared__ WarpState wstate[kPngWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpState &ws = wstate[warp];
    const int64_t img = (int64_t)blockIdx.x * kPngWarpsPerBlock + warp;
    if (img >= n) return;
    const int64_t r = row_idx ? row_idx[img] : img;
    const uint8_t *blob = base + offs[r];
    const int32_t blen = lens[r];
    const int64_t stride = (int64_t)width * channels * sample_bytes;  // bytes per reconstructed scanline
    const int64_t out_bytes = stride * height;
    uint8_t *out = dst + img * out_bytes;
    uint8_t *raw = work + img * work_per_image;  // filtered scanlines: height * (1 + file_stride)

    // ---- lane 0: header walk; then the whole warp inflates (lane 0 owns the bit reader)
    int ctype = 0, depth = 0;
    int64_t file_stride = 0;  // bytes per scanline in the file (palette images: 1 byte per pixel)
    Src s;
    s.p = s.seg_end = s.blob_end = nullptr;
    s.eof = false;
    Bits b;
    b.buf = 0;
    b.cnt = 0;
    if (lane == 0) {
        int err = 0;
        ws.have_plte = 0;
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        if (blen < 8 + 25) err = DE_PNG_CORRUPT;
        for (int i = 0; i < 8 && !err; i++)
            if (blob[i] != sig[i]) err = DE_PNG_CORRUPT;
        s.blob_end = blob + blen;
        if (!err) {
            const uint8_t *q = blob + 8;
            if (be32(q) != 13 || be32(q + 4) != 0x49484452u) err = DE_PNG_CORRUPT;  // IHDR
            else {
                uint32_t w = be32(q + 8), h = be32(q + 12);
                depth = q[16];
                ctype = q[17];
                int interlace = q[20];
                int fch = ctype == 0 ? 1 : (ctype == 2 ? 3 : (ctype == 3 ? 1 : (ctype == 4 ? 2 : 4)));
                int och = ctype == 3 ? 3 : fch;  // palette expands to RGB
                if (q[18] != 0 || q[19] != 0) err = DE_PNG_CORRUPT;
                else if (interlace != 0 || (depth != 8 && depth != 16) || ctype == 4 || ctype == 6 ||
                         (ctype == 3 && depth != 8))
                    err = DE_PNG_UNSUPPORTED;
                else if ((int)w != width || (int)h != height || och != channels || depth / 8 != sample_bytes)
                    err = DE_PNG_UNSUPPORTED;
                file_stride = (int64_t)w * fch * (depth / 8);
                // walk chunks up to the first IDAT, picking up PLTE / tRNS
                const uint8_t *c = q + 8 + 13 + 4;
                bool found = false;
                while (!err && c + 8 <= s.blob_end) {
                    uint32_t len = be32(c), typ = be32(c + 4);
                    const uint8_t *data = c + 8;
                    if (data + len + 4 > s.blob_end) { err = DE_PNG_CORRUPT; break; }
                    if (typ == 0x504c5445u) {  // PLTE
                        if (len > 768 || len % 3) { err = DE_PNG_CORRUPT; break; }
                        for (uint32_t i = 0; i < len; i++) ws.palette[i] = data[i];
                        for (uint32_t i = len; i < 768; i++) ws.palette[i] = 0;
                        ws.have_plte = 1;
                    } else if (typ == 0x74524e53u) {  // tRNS -> cv2 would add an alpha channel
                        err = DE_PNG_UNSUPPORTED;
                        break;
                    } else if (typ == 0x49444154u) {
                        s.p = data;
                        s.seg_end = data + len;
                        found = true;
                        break;
                    } else if (typ == 0x49454e44u) {
                        break;
                    }
                    c = data + len + 4;
                }
                if (!err && !found) err = DE_PNG_CORRUPT;
                if (!err && ctype == 3 && !ws.have_plte) err = DE_PNG_CORRUPT;
            }
        }
        if (!err) {
            int cmf = bits_get(b, s, 8), flg = bits_get(b, s, 8);
            if (cmf < 0 || flg < 0 || (cmf & 15) != 8 || ((cmf << 8) + flg) % 31 != 0 || (flg & 0x20)) err = DE_PNG_CORRUPT;
        }
        ws.err = err;
    }
    __syncwarp();
    int err = ws.err;
    file_stride = __shfl_sync(0xffffffffu, (int)file_stride, 0);
    if (!err) {
        const int64_t raw_total = (int64_t)height * (1 + file_stride);
        int64_t op = 0;
        if (deflate_blocks(b, s, raw, raw_total, ws, op, lane)) err = DE_PNG_CORRUPT;
        else if (shfl_i64(op, 0) != raw_total) err = DE_PNG_CORRUPT;
    }
    if (err) {
        if (lane == 0) {
            if (atomicCAS(status, 0, err) == 0) { status[1] = (int)img; status[2] = blen; }
        }
        // deterministic output for a failed image
        for (int64_t i = lane; i < out_bytes; i += 32) out[i] = 0;
        return;
    }
    ctype = __shfl_sync(0xffffffffu, ctype, 0);
    depth = __shfl_sync(0xffffffffu, depth, 0);
    const int bpp = ctype == 3 ? 1 : channels * sample_bytes;
    const int64_t line = 1 + file_stride;

    // ---- un-filter in place (raw rows), lanes cooperate
    for (int y = 0; y < height; y++) {
        uint8_t *cur = raw + (int64_t)y * line + 1;
        const uint8_t *prior = y ? raw + (int64_t)(y - 1) * line + 1 : nullptr;
        const int ft = raw[(int64_t)y * line];
        if (ft == 0) {
        } else if (ft == 2) {
            if (prior)
                for (int64_t i = lane; i < file_stride; i += 32) cur[i] = (uint8_t)(cur[i] + prior[i]);
        } else if (ft == 1) {
            if (bpp == 3) unfilter_sub<3>(cur, file_stride, lane);
            else if (bpp == 1) unfilter_sub<1>(cur, file_stride, lane);
            else if (bpp == 6) unfilter_sub<6>(cur, file_stride, lane);
            else unfilter_sub<2>(cur, file_stride, lane);
        } else if (ft == 3) {
            if (lane < bpp) {
                int a = 0;
                for (int64_t i = lane; i < file_stride; i += bpp) {
                    int b = prior ? prior[i] : 0;
                    a = (cur[i] + ((a + b) >> 1)) & 255;
                    cur[i] = (uint8_t)a;
                }
            }
        } else if (ft == 4) {
            if (lane < bpp) {
                int a = 0, c = 0;
                for (int64_t i = lane; i < file_stride; i += bpp) {
                    int b = prior ? prior[i] : 0;
                    a = (cur[i] + paeth(a, b, c)) & 255;
                    cur[i] = (uint8_t)a;
                    c = b;
                }
            }
        } else {
            if (lane == 0 && atomicCAS(status, 0, DE_PNG_CORRUPT) == 0) { status[1] = (int)img; status[2] = -ft; }
        }
        __syncwarp();
    }

    // ---- samples -> dst
    if (ctype == 3) {
        for (int64_t p = lane; p < (int64_t)width * height; p += 32) {
            int y = (int)(p / width), x = (int)(p % width);
            int idx = raw[(int64_t)y * line + 1 + x];
            out[p * 3 + 0] = ws.palette[idx * 3 + 0];
            out[p * 3 + 1] = ws.palette[idx * 3 + 1];
            out[p * 3 + 2] = ws.palette[idx * 3 + 2];
        }
    } else if (sample_bytes == 1) {
        for (int y = 0; y < height; y++) {
            const uint8_t *srow = raw + (int64_t)y * line + 1;
            uint8_t *drow = out + (int64_t)y * stride;
            for (int64_t i = lane; i < stride; i += 32) drow[i] = srow[i];
        }
    } else {
        // 16-bit samples are big-endian in the file
        for (int y = 0; y < height; y++) {
            const uint8_t *srow = raw + (int64_t)y * line + 1;
            uint16_t *drow = reinterpret_cast<uint16_t *>(out + (int64_t)y * stride);
            for (int64_t i = lane; i < stride / 2; i += 32) drow[i] = (uint16_t)((srow[2 * i] << 8) | srow[2 * i + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Parquet pages compressed with GZIP (RFC 1952 members around a DEFLATE stream): one warp per page, lane 0 inflates
// into the page image in the arena scratch, the other lanes only move the uncompressed level bytes of V2 pages.
// Replaces the gzip decompression inside Arrow C++ `piece.read` (petastorm/arrow_reader_worker.py:358).  The bit
// stream is serial, so the parallelism is the number of pages of the row-group.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kPngWarpsPerBlock)
k_gzip_pages(uint8_t *__restrict__ arena, const DevPage *__restrict__ pages, const int32_t *__restrict__ list, int n_list,
             int32_t *status) {
    __shared__ WarpState wstate[kPngWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpState &ws = wstate[warp];
    const int li = blockIdx.x * kPngWarpsPerBlock + warp;
    if (li >= n_list) return;
    const int pi = list[li];
    const DevPage pg = pages[pi];
    const uint8_t *src = arena + pg.src_off;
    uint8_t *dst = arena + pg.img_off;
    int64_t src_n = pg.comp_size, dst_n = pg.uncomp_size;
    if (pg.kind == PK_DATA_V2) {
        const int64_t lv = (int64_t)pg.def_bytes + pg.rep_bytes;
        for (int64_t i = lane; i < lv; i += 32) dst[i] = src[i];
        src += lv; dst += lv; src_n -= lv; dst_n -= lv;
    }
    if (src_n <= 0) return;
    // the whole warp walks the members; lane 0 owns the cursor (its err / op are broadcast where the flow depends on them)
    Src s;
    s.p = src;
    s.seg_end = s.blob_end = src + src_n;
    s.eof = true;                       // a single segment: the reader never looks for a next one
    int err = 0;
    int64_t op = 0;
    for (;;) {                          // members may be concatenated
        if (shfl_i64(op, 0) >= dst_n) break;
        Bits b;
        b.buf = 0;
        b.cnt = 0;
        if (lane == 0) {
            do {
                if (s.seg_end - s.p < 18) { err = 1; break; }
                const uint8_t *h = s.p;
                const int flg = h[3];
                if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || (flg & 0xe0)) { err = 2; break; }
                s.p += 10;
                if (flg & 4) {                  // FEXTRA
                    if (s.seg_end - s.p < 2) { err = 1; break; }
                    const int xlen = s.p[0] | (s.p[1] << 8);
                    s.p += 2;
                    if (s.seg_end - s.p < xlen) { err = 1; break; }
                    s.p += xlen;
                }
                for (int f = 8; f <= 16 && !err; f <<= 1) {   // FNAME, FCOMMENT: zero-terminated
                    if (!(flg & f)) continue;
                    while (s.p < s.seg_end && *s.p) s.p++;
                    if (s.p >= s.seg_end) err = 1; else s.p++;
                }
                if (!err && (flg & 2)) {        // FHCRC
                    if (s.seg_end - s.p < 2) err = 1; else s.p += 2;
                }
            } while (0);
        }
        err = __shfl_sync(0xffffffffu, err, 0);
        if (err) break;
        const int64_t before = shfl_i64(op, 0);
        if (deflate_blocks(b, s, dst, dst_n, ws, op, lane)) { err = 3; break; }
        if (lane == 0) {
            // the bit reader may have read ahead: whole bytes left in its buffer belong to the trailer
            s.p -= b.cnt >> 3;
            if (s.seg_end - s.p < 8) err = 1;      // CRC32 (not verified), ISIZE
            else {
                const uint32_t isize = (uint32_t)s.p[4] | ((uint32_t)s.p[5] << 8) | ((uint32_t)s.p[6] << 16) | ((uint32_t)s.p[7] << 24);
                if (isize != (uint32_t)(op - before)) err = 4;
                s.p += 8;
            }
        }
        err = __shfl_sync(0xffffffffu, err, 0);
        if (err) break;
    }
    if (!err && shfl_i64(op, 0) != dst_n) err = 5;
    if (lane == 0 && err && atomicCAS(status, 0, (int)DE_GZIP_CORRUPT) == 0) { status[1] = pi; status[2] = err; }
}

// ---------------------------------------------------------------------------------------------------------------
// CompressedNdarrayCodec blobs (np.savez_compressed: a ZIP archive whose first member `arr.npy` is deflated,
// petastorm/codecs.py:181-198): one warp per blob inflates the first member into dst[i * member_bytes ..).  The .npy
// images are then turned into the batch tensor by k_npy_batch, exactly like NdarrayCodec values.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kPngWarpsPerBlock)
k_zip_inflate_batch(const uint8_t *__restrict__ base, const int64_t *__restrict__ offs, const int32_t *__restrict__ lens,
                    const int64_t *__restrict__ row_idx, int64_t n, int64_t member_bytes, uint8_t *__restrict__ dst,
                    int32_t *status) {
    __shared__ WarpState wstate[kPngWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpState &ws = wstate[warp];
    const int64_t i = (int64_t)blockIdx.x * kPngWarpsPerBlock + warp;
    if (i >= n) return;
    const int64_t r = row_idx ? row_idx[i] : i;
    const uint8_t *blob = base + offs[r];
    const int64_t blen = lens[r];
    uint8_t *out = dst + i * member_bytes;
    int err = 0;
    // (every lane reads the same few header bytes: the flow below stays uniform across the warp)
    auto u16 = [&](int o) { return (int)blob[o] | ((int)blob[o + 1] << 8); };
    if (blen < 30 || blob[0] != 'P' || blob[1] != 'K' || blob[2] != 3 || blob[3] != 4) err = 1;
    else if (u16(6) & 1) err = 2;                            // encrypted
    if (!err) {
        const int method = u16(8);
        const int64_t data_off = 30 + (int64_t)u16(26) + u16(28);
        if (data_off > blen) err = 1;
        else if (method == 8) {
            Src s;
            s.p = blob + data_off;
            s.seg_end = s.blob_end = blob + blen;
            s.eof = true;
            Bits b;
            b.buf = 0;
            b.cnt = 0;
            int64_t op = 0;
            if (deflate_blocks(b, s, out, member_bytes, ws, op, lane)) err = 3;
            else if (shfl_i64(op, 0) != member_bytes) err = 4;
        } else if (method == 0) {
            if (data_off + member_bytes > blen) err = 4;
            else coop_copy(out, blob + data_off, member_bytes, lane, 32);
        } else {
            err = 5;
        }
    }
    if (lane == 0 && err && atomicCAS(status, 0, (int)DE_ZIP_CORRUPT) == 0) { status[1] = (int)i; status[2] = err; }
}

}  // namespace

int64_t png_work_bytes(int height, int width, int channels, int sample_bytes) {
    int64_t line = 1 + (int64_t)width * channels * sample_bytes;
    return ((int64_t)height * line + 63) / 64 * 64;
}

cudaError_t launch_png_batch(const uint8_t *base, const int64_t *offs, const int32_t *lens, const int64_t *row_idx,
                             int64_t n, int height, int width, int channels, int sample_bytes, uint8_t *dst,
                             uint8_t *work, int32_t *status, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + kPngWarpsPerBlock - 1) / kPngWarpsPerBlock;
    k_png_batch<<<(unsigned)blocks, 32 * kPngWarpsPerBlock, 0, s>>>(base, offs, lens, row_idx, n, height, width, channels,
                                                                   sample_bytes, dst, work,
                                                                   png_work_bytes(height, width, channels, sample_bytes),
                                                                   status);
    return cudaGetLastError();
}

cudaError_t launch_zip_inflate_batch(const uint8_t *base, const int64_t *offs, const int32_t *lens,
                                     const int64_t *row_idx, int64_t n, int64_t member_bytes, uint8_t *dst,
                                     int32_t *status, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + kPngWarpsPerBlock - 1) / kPngWarpsPerBlock;
    k_zip_inflate_batch<<<(unsigned)blocks, 32 * kPngWarpsPerBlock, 0, s>>>(base, offs, lens, row_idx, n, member_bytes,
                                                                           dst, status);
    return cudaGetLastError();
}

cudaError_t launch_gzip_pages(uint8_t *arena, const DevPage *pages, const int32_t *list, int n, int32_t *status,
                              cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_gzip_pages<<<(n + kPngWarpsPerBlock - 1) / kPngWarpsPerBlock, 32 * kPngWarpsPerBlock, 0, s>>>(arena, pages, list, n,
                                                                                                     status);
    return cudaGetLastError();
}

}  // namespace pst
