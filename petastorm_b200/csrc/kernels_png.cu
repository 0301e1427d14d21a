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

struct WarpState {
    Huff lencode;
    HuffD distcode;
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

__device__ __forceinline__ int paeth(int a, int b, int c) {
    int p = a + b - c;
    int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

// DEFLATE blocks (RFC 1951) from the bit reader into raw[op ..), at most raw_total bytes; returns 0 or 1 (corrupt).
// One lane; `ws` holds the Huffman tables of its warp.
__device__ int deflate_blocks(Bits &b, Src &s, uint8_t *raw, int64_t raw_total, WarpState &ws, int64_t &op) {
    int err = 0;
    int last = 0;
    while (!err && !last) {
        last = bits_get(b, s, 1);
        int type = bits_get(b, s, 2);
        if (last < 0 || type < 0) { err = 1; break; }
        if (type == 0) {
            // stored: drop bits to the byte boundary, LEN / NLEN, then raw bytes
            b.buf = 0;
            b.cnt = 0;  // bits_get refills byte-wise, so a partial byte is all that can be buffered... see note
            int l0 = src_byte(s), l1 = src_byte(s), n0 = src_byte(s), n1 = src_byte(s);
            if (l0 < 0 || l1 < 0 || n0 < 0 || n1 < 0) { err = 1; break; }
            int len = l0 | (l1 << 8);
            if ((len ^ 0xffff) != (n0 | (n1 << 8))) { err = 1; break; }
            if (op + len > raw_total) { err = 1; break; }
            while (len > 0) {
                if (s.p == s.seg_end) {
                    src_next_segment(s);
                    if (s.eof) { err = 1; break; }
                }
                int64_t take = s.seg_end - s.p;
                if (take > len) take = len;
                for (int64_t i = 0; i < take; i++) raw[op + i] = s.p[i];
                s.p += take;
                op += take;
                len -= (int)take;
            }
            continue;
        }
        if (type == 3) { err = 1; break; }
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
        // literal/length + distance codes
        for (;;) {
            int sym = huff_decode(b, s, ws.lencode);
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
                int ds = huff_decode(b, s, ws.distcode);
                if (ds < 0 || ds >= 30) { err = 1; break; }
                int de = bits_get(b, s, c_dist_extra[ds]);
                if (de < 0) { err = 1; break; }
                int64_t dist = (int64_t)c_dist_base[ds] + de;
                if (dist > op || op + len > raw_total) { err = 1; break; }
                for (int i = 0; i < len; i++, op++) raw[op] = raw[op - dist];
            }
        }
    }
    return err;
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

    // ---- lane 0: header walk + inflate
    int ctype = 0, depth = 0;
    int64_t file_stride = 0;  // bytes per scanline in the file (palette images: 1 byte per pixel)
    if (lane == 0) {
        int err = 0;
        ws.have_plte = 0;
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        if (blen < 8 + 25) err = DE_PNG_CORRUPT;
        for (int i = 0; i < 8 && !err; i++)
            if (blob[i] != sig[i]) err = DE_PNG_CORRUPT;
        Src s;
        s.blob_end = blob + blen;
        s.eof = false;
        s.p = s.seg_end = nullptr;
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
        ws.err = err;
        // hand the source cursor to the inflate loop below through registers (lane 0 only)
        if (!err) {
            const int64_t raw_total = (int64_t)height * (1 + file_stride);
            Bits b;
            b.buf = 0;
            b.cnt = 0;
            int cmf = bits_get(b, s, 8), flg = bits_get(b, s, 8);
            if (cmf < 0 || flg < 0 || (cmf & 15) != 8 || ((cmf << 8) + flg) % 31 != 0 || (flg & 0x20)) err = DE_PNG_CORRUPT;
            int64_t op = 0;
            if (!err && deflate_blocks(b, s, raw, raw_total, ws, op)) err = DE_PNG_CORRUPT;
            if (!err && op != raw_total) err = DE_PNG_CORRUPT;
            ws.err = err;
        }
    }
    __syncwarp();
    const int err = ws.err;
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
    file_stride = __shfl_sync(0xffffffffu, file_stride, 0);
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
            if (lane < bpp) {
                int a = 0;
                for (int64_t i = lane; i < file_stride; i += bpp) {
                    a = (cur[i] + a) & 255;
                    cur[i] = (uint8_t)a;
                }
            }
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
    if (lane != 0 || src_n <= 0) return;
    Src s;
    s.p = src;
    s.seg_end = s.blob_end = src + src_n;
    s.eof = true;                       // a single segment: the reader never looks for a next one
    int err = 0;
    int64_t op = 0;
    while (!err && op < dst_n) {        // members may be concatenated
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
        if (err) break;
        Bits b;
        b.buf = 0;
        b.cnt = 0;
        const int64_t before = op;
        if (deflate_blocks(b, s, dst, dst_n, ws, op)) { err = 3; break; }
        if (s.seg_end - s.p < 8) { err = 1; break; }      // CRC32 (not verified), ISIZE
        const uint32_t isize = (uint32_t)s.p[4] | ((uint32_t)s.p[5] << 8) | ((uint32_t)s.p[6] << 16) | ((uint32_t)s.p[7] << 24);
        if (isize != (uint32_t)(op - before)) { err = 4; break; }
        s.p += 8;
    }
    if (!err && op != dst_n) err = 5;
    if (err && atomicCAS(status, 0, (int)DE_GZIP_CORRUPT) == 0) { status[1] = pi; status[2] = err; }
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
    if (i >= n || lane != 0) return;
    const int64_t r = row_idx ? row_idx[i] : i;
    const uint8_t *blob = base + offs[r];
    const int64_t blen = lens[r];
    uint8_t *out = dst + i * member_bytes;
    int err = 0;
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
            if (deflate_blocks(b, s, out, member_bytes, ws, op)) err = 3;
            else if (op != member_bytes) err = 4;
        } else if (method == 0) {
            if (data_off + member_bytes > blen) err = 4;
            else for (int64_t k = 0; k < member_bytes; k++) out[k] = blob[data_off + k];
        } else {
            err = 5;
        }
    }
    if (err && atomicCAS(status, 0, (int)DE_ZIP_CORRUPT) == 0) { status[1] = (int)i; status[2] = err; }
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
