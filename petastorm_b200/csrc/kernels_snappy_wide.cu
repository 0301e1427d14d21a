// k_snappy_frag -- raw-Snappy decompress of one 64 KiB fragment (or one whole single-fragment page) per CTA, "wide":
// all 256 threads of the CTA work on the same 1 KiB window of the compressed stream at a time.
//
// The three-warp pipeline of k_snappy_pages (kernels_decode.cu) retires one 32-element batch per ~3000 cycles: its cost
// is fixed overhead (hand-overs, shuffles, header loads) amortised over 32 elements.  Here a window carries ~256
// elements (a C2 int64 page has 4 bytes of stream per element), so the same fixed costs are paid once per window:
//
//   1. tables   every thread turns 4 byte positions of the window into "the element that would start here"
//               {stream bytes, output bytes, 1} (256-entry tag lookup), then five rounds of pointer doubling
//               T2k[p] = Tk[p] + Tk[p + stream(Tk[p])] give the same for 2, 4, ... 32 consecutive elements.  T1, T4, T16
//               and T32 are kept.  A slow-path tag (literal with length bytes, copy with a 4-byte offset), a position
//               behind the window or behind the stream is {0,0,0}, which stops every chain that reaches it.
//   2. walk     thread 0 follows T32 from the known start: one shared-memory load per 32 elements, recording a hop
//               {stream position, output position, elements} until the window ends, a slow-path element is reached or the
//               chunk is full (<= 8 KiB of output, <= 512 elements - the ring holds 16 KiB).
//   3. expand   eight threads per hop find the start of every element: at most one T16 step, three T4 steps and four T1
//               steps each.
//   4. decode   one thread per element reads its tag bytes, validates, and moves its literal into the output ring.
//   5. copies   back-references ring -> ring in dependency rounds: a copy runs once every element in front of its source
//               is complete (the first pending one always is).  Typical streams need two or three rounds.
//   6. flush    the ring is written through to HBM at the end of every chunk with 16-byte vector copies (output
//               positions are biased so that ring and HBM agree modulo 16).
//
// A source that left the ring is read back from HBM (everything in front of the current chunk has been flushed).  Long
// literals bypass the element machinery: all threads copy them global -> global (and mirror them into the ring when they
// are short enough for later back-references to find them there).  Streams the reference compressor never emits - a
// back-reference into an earlier fragment - raise the page flag; the serial fallback launch of k_snappy_pages redoes the
// page as one stream, exactly as before.
//
// Replaces Snappy inside Arrow C++ `piece.read` (petastorm/arrow_reader_worker.py:358, py_dict_reader_worker.py:267).
// Format: google/snappy format_description.txt.  Algorithmic bytes: stored bytes read + image bytes written.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"
#include "dev_util.cuh"
#include "kernels.h"

namespace pst {

constexpr int kWThreads = 256;
constexpr int kWWin = 1024;                 // stream positions per window (4 per thread)
constexpr int kWPad = 64;                   // zero entries behind the window: a fast-path element is <= 62 bytes long
constexpr int kWRing = 16384;
constexpr uint32_t kWRingMask = kWRing - 1;
constexpr uint32_t kWChunkOut = 8192;       // output bytes per chunk (a hop adds <= 2 KiB)
constexpr int kWMaxEl = 512;                // elements per chunk
constexpr int kWMaxHops = 64;
constexpr uint32_t kWMirror = 4096;         // bypassed literals up to this length are mirrored into the ring

// table entry: stream bytes (bits 0-10) | output bytes (bits 11-22) | elements (bits 23-28) of a run of consecutive
// fast-path elements; the fields of two runs add without carries (32 elements: <= 1984 / 2048 / 32)
__device__ __forceinline__ uint32_t ent_used(uint32_t e) { return e & 0x7ffu; }
__device__ __forceinline__ uint32_t ent_made(uint32_t e) { return (e >> 11) & 0xfffu; }
__device__ __forceinline__ uint32_t ent_count(uint32_t e) { return e >> 23; }

struct WideShared {
    uint8_t ring[kWRing + 16];
    uint32_t t1[kWWin + kWPad];
    uint32_t t4[kWWin + kWPad];
    uint32_t t16[kWWin + kWPad];
    uint32_t t32[kWWin + kWPad];
    uint32_t tmp[kWWin + kWPad];
    uint32_t lut[256];
    uint32_t hop_ip[kWMaxHops];
    uint32_t hop_op[kWMaxHops];
    uint16_t hop_first[kWMaxHops];
    uint16_t hop_cnt[kWMaxHops];
    uint32_t e_ip[kWMaxEl];
    uint32_t e_op[kWMaxEl];
    uint8_t e_done[kWMaxEl];    // element complete in the ring (literals after the decode phase, copies once they ran)
    // control words written by thread 0 (or with atomics) between barriers
    uint32_t ip, op;            // stream / output position behind the current chunk
    uint32_t n_hops, n_el;
    uint32_t special;           // a slow-path element follows the chunk
    uint32_t sp_kind, sp_src, sp_len;
    uint32_t err;               // 0 ok, 1.. corrupt (detail), 0x100 cross-fragment reference
};

__global__ void __launch_bounds__(kWThreads)
k_snappy_frag(uint8_t *__restrict__ arena, const DevPage *__restrict__ pages, const SnFrag *__restrict__ frags,
              int n_frags, const uint32_t *__restrict__ frag_pos, uint32_t *page_flag, int32_t *status) {
    extern __shared__ __align__(16) uint8_t wide_smem[];
    WideShared &sh = *reinterpret_cast<WideShared *>(wide_smem);
    const int tid = threadIdx.x;
    const int li = blockIdx.x;
    if (li >= n_frags) return;
    const int pi = frags[li].page;
    const int frag_k = frags[li].k;
    const DevPage pg = pages[pi];
    if (pg.multi_slot >= 0 && *(volatile uint32_t *)&page_flag[pg.multi_slot] != 0) return;   // the serial launch takes it

    const uint8_t *src = arena + pg.src_off;
    uint8_t *dst = arena + pg.img_off;
    uint32_t src_n = (uint32_t)pg.comp_size;
    uint32_t dst_n = (uint32_t)pg.uncomp_size;
    // V2 data pages: the level bytes are stored uncompressed in front of the compressed values
    if (pg.kind == PK_DATA_V2) {
        const uint32_t lv = (uint32_t)(pg.def_bytes + pg.rep_bytes);
        if (frag_k == 0) coop_copy(dst, src, lv, tid, kWThreads);
        src += lv; dst += lv; src_n -= lv; dst_n -= lv;
    }
    const uint32_t full_n = dst_n;                      // the length the stream's preamble must announce
    if (pg.nfrag > 1) {
        const uint32_t c0 = frag_pos[pg.frag_first + frag_k], c1 = frag_pos[pg.frag_first + frag_k + 1];
        src += c0;
        src_n = c1 - c0;
        dst += (uint32_t)frag_k * (uint32_t)kSnappyFragment;
        dst_n = min((uint32_t)kSnappyFragment, full_n - (uint32_t)frag_k * (uint32_t)kSnappyFragment);
    }
    // output positions are biased by the misalignment of `dst`: position p lives at ring[p & mask] and at dst[p]
    const uint32_t bias = (uint32_t)((uintptr_t)dst & 15);
    dst -= bias;
    dst_n += bias;
    if ((int32_t)src_n <= 0) return;
    // stream addressing: `gin` is the 16-byte aligned base, positions are 32-bit offsets from it
    const uint8_t *gin = src - ((uintptr_t)src & 15);
    const uint32_t in_begin = (uint32_t)((uintptr_t)src & 15);
    const uint32_t in_end = in_begin + src_n;

    // ---- set-up: tag lookup table, zero pads, preamble
    {
        const uint32_t kind = tid & 3, t6 = tid >> 2;
        uint32_t used = 0, made = 0;
        if (kind == 0) { if (t6 < 60) { used = t6 + 2; made = t6 + 1; } }
        else if (kind == 1) { used = 2; made = (t6 & 7) + 4; }
        else if (kind == 2) { used = 3; made = t6 + 1; }
        sh.lut[tid] = used ? (used | (made << 11) | (1u << 23)) : 0u;
    }
    if (tid < kWPad) {
        sh.t1[kWWin + tid] = 0; sh.t4[kWWin + tid] = 0; sh.t16[kWWin + tid] = 0; sh.t32[kWWin + tid] = 0;
        sh.tmp[kWWin + tid] = 0;
    }
    if (tid == 0) {
        uint32_t ip = in_begin, err = 0;
        if (frag_k == 0) {      // varint uncompressed length
            uint64_t ulen = 0;
            int shift = 0;
            for (;;) {
                if (ip >= in_end || shift > 35) { err = 1; break; }
                const uint8_t b = gin[ip++];
                ulen |= (uint64_t)(b & 0x7f) << shift;
                if (!(b & 0x80)) break;
                shift += 7;
            }
            if (!err && ulen != (uint64_t)full_n) err = 2;
        }
        sh.ip = ip;
        sh.op = bias;
        sh.err = err;
    }
    __syncthreads();

    uint32_t flushed = bias;          // output positions below this are in HBM
    uint32_t valid_from = bias;       // output positions below this are not in the ring (bypassed literal)
    uint32_t tab_ws = 0xffffffffu;    // window the tables describe: [tab_ws, tab_ws + kWWin)
    uint32_t pf_ws = 0xffffffffu, pf_word = 0;   // stream word of this thread for the window expected next (prefetched)

    for (;;) {
        if (sh.err) break;
        const uint32_t ip0 = sh.ip;
        if (ip0 >= in_end) break;
        // ---- 1. tables (only when the position left the window they describe)
        if (ip0 < tab_ws || ip0 >= tab_ws + (uint32_t)kWWin) {
            tab_ws = ip0 & ~3u;
            {
                const uint32_t pos = tab_ws + 4u * (uint32_t)tid;
                const uint32_t word = pf_ws == tab_ws ? pf_word
                                                      : (pos < in_end ? __ldg(reinterpret_cast<const uint32_t *>(gin + pos)) : 0u);
                uint32_t e[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    e[q] = sh.lut[(word >> (8 * q)) & 0xffu];
                    if (pos + (uint32_t)q >= in_end) e[q] = 0;
                }
                *reinterpret_cast<uint4 *>(&sh.t1[4 * tid]) = make_uint4(e[0], e[1], e[2], e[3]);
            }
            __syncthreads();
            // pointer doubling: 2, 4, 8, 16, 32 elements (an entry of 0 re-reads itself and stays 0)
#define PST_DOUBLE(FROM, TO)                                                  \
            {                                                                 \
                uint32_t e[4], f[4];                                          \
                _Pragma("unroll") for (int q = 0; q < 4; q++) e[q] = FROM[tid + kWThreads * q];                         \
                _Pragma("unroll") for (int q = 0; q < 4; q++) f[q] = FROM[tid + kWThreads * q + ent_used(e[q])];        \
                _Pragma("unroll") for (int q = 0; q < 4; q++) TO[tid + kWThreads * q] = e[q] + f[q];                    \
            }                                                                 \
            __syncthreads();
            PST_DOUBLE(sh.t1, sh.tmp)
            PST_DOUBLE(sh.tmp, sh.t4)
            PST_DOUBLE(sh.t4, sh.tmp)
            PST_DOUBLE(sh.tmp, sh.t16)
            PST_DOUBLE(sh.t16, sh.t32)
#undef PST_DOUBLE
        }
        // ---- 2. walk (thread 0)
        if (tid == 0) {
            uint32_t ip = ip0, op = sh.op, n_hops = 0, n_el = 0, special = 0;
            const uint32_t op0 = op;
            while (n_hops < (uint32_t)kWMaxHops && ip < in_end) {
                const uint32_t rel = ip - tab_ws;
                if (rel >= (uint32_t)kWWin) break;                  // the next window continues
                const uint32_t e = sh.t32[rel];
                const uint32_t cnt = ent_count(e);
                if (cnt == 0) { special = 1; break; }               // slow-path element at ip
                const uint32_t made = ent_made(e);
                if (n_hops && (op - op0 + made > kWChunkOut || n_el + cnt > (uint32_t)kWMaxEl)) break;   // chunk full
                sh.hop_ip[n_hops] = ip;
                sh.hop_op[n_hops] = op;
                sh.hop_first[n_hops] = (uint16_t)n_el;
                sh.hop_cnt[n_hops] = (uint16_t)cnt;
                n_hops++;
                n_el += cnt;
                ip += ent_used(e);
                op += made;
            }
            if (op > dst_n) sh.err = 8;
            sh.ip = ip;
            sh.op = op;
            sh.n_hops = n_hops;
            sh.n_el = n_el;
            sh.special = special;
        }
        __syncthreads();
        if (sh.err) break;
        const uint32_t n_hops = sh.n_hops, n_el = sh.n_el, op_end = sh.op, special = sh.special;
        {   // the window that follows starts where the walk stopped (unless a slow-path element moves it): fetch its
            // bytes now, they are needed after this chunk's copies
            const uint32_t nws = sh.ip & ~3u;
            if (nws != pf_ws && (nws < tab_ws || nws >= tab_ws + (uint32_t)kWWin)) {
                const uint32_t pos = nws + 4u * (uint32_t)tid;
                pf_word = pos < in_end ? __ldg(reinterpret_cast<const uint32_t *>(gin + pos)) : 0u;
                pf_ws = nws;
            }
        }

        if (n_el) {
            // ---- 3. expand: eight threads per hop, four elements each
            for (uint32_t gi = tid; gi < n_hops * 8u; gi += kWThreads) {
                const uint32_t h = gi >> 3, g = gi & 7u;
                const uint32_t cnt = sh.hop_cnt[h];
                if (4u * g >= cnt) continue;
                uint32_t ip = sh.hop_ip[h], op = sh.hop_op[h];
                if (g & 4u) {
                    const uint32_t e = sh.t16[ip - tab_ws];
                    ip += ent_used(e);
                    op += ent_made(e);
                }
                for (uint32_t k = 0; k < (g & 3u); k++) {
                    const uint32_t e = sh.t4[ip - tab_ws];
                    ip += ent_used(e);
                    op += ent_made(e);
                }
                const uint32_t base = (uint32_t)sh.hop_first[h] + 4u * g;
                for (uint32_t k = 0; k < 4u && 4u * g + k < cnt; k++) {
                    sh.e_ip[base + k] = ip;
                    sh.e_op[base + k] = op;
                    const uint32_t e = sh.t1[ip - tab_ws];
                    ip += ent_used(e);
                    op += ent_made(e);
                }
            }
            __syncthreads();
            // ---- 4. decode + literals (one thread per element, two passes at most)
            uint32_t c_d[2], c_a[2], c_len[2], c_j[2];
            bool c_pend[2] = {false, false};
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const uint32_t j = (uint32_t)tid + (uint32_t)s * kWThreads;
                if (j >= n_el) continue;
                const uint32_t ip = sh.e_ip[j], d = sh.e_op[j];
                const uint32_t tag = gin[ip];
                const uint32_t kind = tag & 3u, t6 = tag >> 2;
                const uint32_t used = ent_used(sh.lut[tag]);
                uint32_t len, a = 0;
                if (kind == 0) len = t6 + 1;
                else if (kind == 1) { len = (t6 & 7u) + 4u; a = ((tag >> 5) << 8) | gin[ip + 1]; }
                else { len = t6 + 1; a = (uint32_t)gin[ip + 1] | ((uint32_t)gin[ip + 2] << 8); }
                sh.e_done[j] = (uint8_t)(kind == 0);
                if (used == 0 || ip + used > in_end || d + len > dst_n || (kind != 0 && a == 0)) {
                    atomicMax(&sh.err, 8u);
                    continue;
                }
                if (kind == 0) {
                    const uint8_t *lp = gin + ip + 1;
                    for (uint32_t i = 0; i < len; i++) sh.ring[(d + i) & kWRingMask] = lp[i];
                } else {
                    if (a > d - bias) {          // source in front of this stream's first output byte
                        atomicMax(&sh.err, (frag_k == 0) ? 8u : 0x100u);
                        continue;
                    }
                    c_d[s] = d; c_a[s] = a; c_len[s] = len; c_j[s] = j; c_pend[s] = true;
                }
            }
            __syncthreads();
            if (sh.err) break;
            // ---- 5. back-references in dependency rounds.  A copy is ready once every element that produces a byte of
            // its source range is complete: elements of earlier chunks always are; inside the chunk the producing elements
            // are found by binary search over the (sorted) output positions and their done flags are checked.  The first
            // pending copy is always ready, so every round makes progress; typical streams need two or three rounds.
            uint32_t c_lo[2], c_hi[2];
            const uint32_t chunk_op0 = sh.e_op[0];
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (!c_pend[s]) continue;
                const uint32_t sp = c_d[s] - c_a[s];
                const uint32_t src_end = min(sp + c_len[s], c_d[s]);       // bytes >= d are produced by the copy itself
                if (src_end <= chunk_op0) { c_lo[s] = 1; c_hi[s] = 0; continue; }      // empty range: nothing to wait for
                // largest index whose output position is <= q
                auto find = [&](uint32_t q) {
                    uint32_t lo = 0, hi = c_j[s];          // the element itself starts at d > q
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (sh.e_op[mid] <= q) lo = mid; else hi = mid;
                    }
                    return lo;
                };
                c_lo[s] = find(max(sp, chunk_op0));
                c_hi[s] = find(src_end - 1);
            }
            for (;;) {
                bool still = false;
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    if (!c_pend[s]) continue;
                    bool ready = true;
                    for (uint32_t k = c_lo[s]; k <= c_hi[s]; k++)
                        if (!sh.e_done[k]) { ready = false; break; }
                    if (!ready) { still = true; continue; }
                    __threadfence_block();                 // the producers' ring bytes before their done flags
                    const uint32_t d = c_d[s], len = c_len[s];
                    const uint32_t sp = d - c_a[s];
                    if (sp >= valid_from && sp + (uint32_t)kWRing >= op_end) {
                        // sequential: an overlapping copy (offset < length) re-reads its own bytes
                        for (uint32_t i = 0; i < len; i++)
                            sh.ring[(d + i) & kWRingMask] = sh.ring[(sp + i) & kWRingMask];
                    } else {
                        // the source left the ring (or was bypassed): it lies in front of this chunk, hence in HBM
                        for (uint32_t i = 0; i < len; i++) {
                            const uint32_t q = sp + i;
                            sh.ring[(d + i) & kWRingMask] = q < flushed ? dst[q] : sh.ring[q & kWRingMask];
                        }
                    }
                    __threadfence_block();
                    sh.e_done[c_j[s]] = 1;
                    c_pend[s] = false;
                }
                if (!__syncthreads_or(still ? 1 : 0)) break;
            }
            // ---- 6. flush the chunk
            {
                uint32_t f = flushed;
                while (f < op_end) {
                    const uint32_t r = f & kWRingMask;
                    const uint32_t nn = min(op_end - f, (uint32_t)kWRing - r);
                    coop_copy(dst + f, sh.ring + r, nn, tid, kWThreads);
                    f += nn;
                }
                flushed = op_end;
            }
            __syncthreads();
        }
        // ---- slow-path element behind the chunk
        if (special) {
            if (tid == 0) {
                const uint32_t ip = sh.ip, op = sh.op;
                const uint32_t tag = gin[ip];
                const uint32_t t6 = tag >> 2;
                uint32_t err = 0;
                if ((tag & 3u) == 3u) {                       // copy with a 4-byte offset
                    if (ip + 5 > in_end) err = 5;
                    else {
                        sh.sp_kind = 1;
                        sh.sp_src = (uint32_t)gin[ip + 1] | ((uint32_t)gin[ip + 2] << 8) | ((uint32_t)gin[ip + 3] << 16) |
                                    ((uint32_t)gin[ip + 4] << 24);
                        sh.sp_len = t6 + 1;
                        if (sh.sp_src == 0 || op + t6 + 1 > dst_n) err = 8;
                        else if (sh.sp_src > op - bias) err = (frag_k == 0) ? 8u : 0x100u;
                        sh.ip = ip + 5;
                    }
                } else if ((tag & 3u) == 0 && t6 >= 60) {     // literal with 1..4 length bytes
                    const uint32_t nb = t6 - 59;
                    if (ip + 1 + nb > in_end) err = 4;
                    else {
                        uint32_t v = 0;
                        for (uint32_t i = 0; i < nb; i++) v |= (uint32_t)gin[ip + 1 + i] << (8 * i);
                        const uint32_t len = v + 1, p0 = ip + 1 + nb;
                        if (len == 0 || len > in_end - p0 || len > dst_n - op) err = 4;
                        else {
                            sh.sp_kind = 0;
                            sh.sp_src = p0;
                            sh.sp_len = len;
                            sh.ip = p0 + len;
                        }
                    }
                } else {
                    err = 6;      // a fast-path tag with a zero table entry: the stream ends inside an element
                }
                if (err) sh.err = err;
            }
            __syncthreads();
            if (sh.err) break;
            const uint32_t op = sh.op, len = sh.sp_len;
            if (sh.sp_kind == 0) {
                // long literal: straight to HBM; short ones are mirrored into the ring for later back-references
                coop_copy(dst + op, gin + sh.sp_src, len, tid, kWThreads);
                if (len <= kWMirror) {
                    const uint8_t *lp = gin + sh.sp_src;
                    for (uint32_t i = tid; i < len; i += kWThreads) sh.ring[(op + i) & kWRingMask] = lp[i];
                } else {
                    valid_from = op + len;
                }
            } else {
                // copy-4: one thread, source from HBM or ring byte by byte (sequential: it may overlap itself)
                if (tid == 0) {
                    const uint32_t sp = op - sh.sp_src;
                    for (uint32_t i = 0; i < len; i++) {
                        const uint32_t q = sp + i;
                        uint8_t v;
                        if (q >= op) v = sh.ring[q & kWRingMask];                       // its own output
                        else if (q < valid_from || q + (uint32_t)kWRing < op + len) v = dst[q];
                        else v = sh.ring[q & kWRingMask];
                        sh.ring[(op + i) & kWRingMask] = v;
                        dst[op + i] = v;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) sh.op = op + len;
            flushed = op + len;
            __syncthreads();
        }
    }
    __syncthreads();
    const uint32_t err = sh.err;
    if (tid == 0) {
        if (err == 0x100u) {
            // a back-reference into an earlier fragment: legal Snappy, just not what the reference compressor emits
            if (pg.multi_slot >= 0) *(volatile uint32_t *)&page_flag[pg.multi_slot] = 2;
            else report_error(status, DE_SNAPPY_CORRUPT, pi, 8);
        } else if (err) {
            report_error(status, DE_SNAPPY_CORRUPT, pi, (int)err);
        } else if (sh.op != dst_n || sh.ip != in_end) {
            report_error(status, DE_SNAPPY_CORRUPT, pi, 9);
        }
    }
}

cudaError_t configure_snappy_wide() {
    cudaError_t e = cudaFuncSetAttribute(k_snappy_frag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WideShared));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_snappy_frag, cudaFuncAttributePreferredSharedMemoryCarveout,
                                (int)cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_snappy_frag(uint8_t *arena, const DevPage *pages, const SnFrag *frags, int n_frags,
                               const uint32_t *frag_pos, uint32_t *page_flag, int32_t *status, cudaStream_t s) {
    if (n_frags <= 0) return cudaSuccess;
    k_snappy_frag<<<n_frags, kWThreads, sizeof(WideShared), s>>>(arena, pages, frags, n_frags, frag_pos, page_flag, status);
    return cudaGetLastError();
}

}  // namespace pst
