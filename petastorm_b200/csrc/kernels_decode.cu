// sm_100a decode kernels for Parquet pages resident in HBM.
//
//   k_snappy_index   K2a compressed offset of every 64 KiB output boundary of a Snappy page (builder warps + walker);
//                        k_snappy_index_cluster: the same on a four-CTA cluster per page (opt-in, big pages)
//   k_snappy_pages   K2  raw-Snappy decompress, one CTA (parser, placement and copy warp) per 64 KiB fragment
//   k_ba_dict_index  K4  BYTE_ARRAY dictionary entry index ({offset,len} per entry)
//   k_decode_pages   K3/K4/K5/K6  levels (RLE/bit-packed hybrid) + PLAIN / dictionary values + validity,
//                    one CTA (256 threads) per data page
//
// They replace what Arrow C++ does inside `piece.read(columns=...)`
// (reference call sites petastorm/arrow_reader_worker.py:358, petastorm/py_dict_reader_worker.py:267).
// Formats follow the public specifications: google/snappy format_description.txt, parquet-format Encodings.md.
//
// All of this is byte/integer work bound by HBM bandwidth, not by math: no tensor cores.  The design rules applied are
// coalesced 16-byte accesses (the planner places every value section 16B-aligned, see host_api.cpp), shared-memory
// staging of run tables, and work items small enough (64 KiB Snappy fragments, pages) that a 256 MB row-group
// (~2900 pages, ~6300 fragments) fills the 148 SMs several times over.
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"
#include "dev_util.cuh"
#include "kernels.h"

namespace pst {

// ---------------------------------------------------------------------------------------------------------------
// K2  Snappy (raw block format, google/snappy format_description.txt).  One CTA = one 64 KiB fragment (or, in the serial
// fallback launch, one whole page) = three warps working as a pipeline over batches of <= 32 elements:
//
//   warp P (parser)     finds where the elements START -- the inherently serial part, every tag position depends on the
//                       previous element.  Tag bytes come from a shared-memory staging window filled by cp.async
//                       (LDGSTS); for every 768 bytes of input all 32 lanes build successor tables (element length for
//                       every byte position, then the lengths of the four elements that follow each position), so
//                       the chain costs one LDS + one dp4a per FOUR elements.  Output: <= 8 hop records per batch.
//   warp A (placement)  decodes one element per lane, assigns output positions with a warp scan, validates, and moves
//                       literal bytes staging -> ring.
//   warp B (copy)       resolves back-references ring -> ring in dependency rounds (a copy runs once every element in
//                       front of its source range is complete) and writes the ring through to HBM in >= 2 KiB
//                       pieces (positions are biased so that ring and HBM agree modulo 16: plain 16-byte copies).
//
// The ring holds the most recent 8 KiB of output in shared memory, so the many tiny copies of a match-heavy stream
// never pay a global store -> L2 -> global load round trip.  Neighbouring warps hand batches over by a rendezvous on
// one named barrier per pair (double-buffered slots): P parses batch b+1 while A places batch b and B copies batch
// b-1.  Literals >= 1 KiB bypass staging and ring (one vectorised global -> global copy by B).  A back-reference that
// reaches outside the ring, or into a bypassed literal, is served from the output already written to HBM, every lane
// fetching its own source.  23 KB of shared memory and 64 registers per thread: ten CTAs per SM.
//
// Why this shape: measured on B200 (profiles/r1_snappy_v2_ring.txt ... r1_final_three_stage.txt) a lone warp retires
// ~1 dependent instruction per ~5.5 cycles and a taken branch costs ~15, so the cost of a stream is (instructions on
// the serial chain) x 5.5 cycles x elements; a 1 MiB page of a C2 int64 column has 2.3e5 elements.  Hence: shortest
// possible serial chain (tables), the rest of the per-element work spread over 32 lanes and three warps, and 64 KiB
// fragments (k_snappy_index) so that a page is 16 CTAs instead of one.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane);   // defined with the page decoder below

constexpr int kSnappyThreads = 96;      // parser warp, placement warp, copy warp
#ifndef PST_RING
#define PST_RING 8192
#endif
constexpr int kRing = PST_RING;          // power of two; with staging and tables ~23 KiB of shared memory per CTA (9 CTAs/SM).
// (History: in round 1 an 8 KiB ring lost to 16 KiB - 1.66 against 1.38 ms per C2 row-group - because 8 % of the
// back-references of the int64 pages then reach behind the ring and went through a serial path with a flush in front.
// Since every lane fetches such a source from HBM on its own (warp B) the occupancy wins: 0.99 ms with 16 KiB and
// 7 CTAs/SM, 0.83 ms with 8 KiB and 9 CTAs/SM.)
constexpr uint32_t kRingMask = kRing - 1;
constexpr int kStage = 8192;          // input staging window: 4 chunks of 2 KiB (power of two)
constexpr uint32_t kStageMask = kStage - 1;
constexpr int kChunkShift = 11;
constexpr uint32_t kChunk = 1u << kChunkShift;
constexpr int kBatchOps = 32;
constexpr uint32_t kBatchIn = 2048;   // >= input span of a batch without its last element (32 elements x <= 62 bytes)
// (output per batch is bounded by the two limits above: 32 copies x 64 B + < 2 KiB of literals)
constexpr uint32_t kBigLiteral = 1024;
constexpr uint32_t kFlushBytes = PST_RING >= 16384 ? 4096 : 2048;   // kFlushBytes + 2 * kMaxBatchOut <= kRing: A never overwrites unflushed bytes
constexpr uint32_t kLookahead = kBatchIn + kBigLiteral + 8;   // staged bytes a batch may touch past its start

constexpr int kHops = kBatchOps / 4;   // the parser advances four elements per step ("hop")
struct SnBatch {
    uint2 hop[kHops];         // {input offset of the hop's first tag byte, the byte lengths of its <= 4 elements}
    uint32_t n;               // hop records in use
    uint32_t last;            // no batch follows
    uint32_t err;             // parser error code (0 = ok)
    uint32_t big_len;         // length of a bypassed big literal
    uint32_t rare_a;          // the element after the last hop needed the slow path (copy-4 / literal with a length
    uint32_t rare_h;          //   suffix): its decoded {source or offset, length | kind << 24}; rare_h == 0 when absent
    uint32_t pad_[2];
};
static_assert(sizeof(SnBatch) == 96, "SnBatch layout");
constexpr uint32_t kHdrOff = kHops * 8;
// successor tables of the parser (see warp P)
#ifndef PST_TABW
#define PST_TABW 768
#endif
constexpr int kTabW = PST_TABW;           // input positions covered by one table build (multiple of 256; 768: 10 CTAs/SM)
constexpr int kTabPad = 64;           // zero entries behind the window: an element is at most 61 bytes long


__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// Named barriers with IMMEDIATE ids: with a register id ptxas reserves all 16 hardware barriers for the CTA, and the
// 64 barriers of an SM then cap residency at 4 CTAs (launch__occupancy_limit_barriers in profiles/r1_snappy_v8).
template <int ID>
__device__ __forceinline__ void bar_sync_imm() { asm volatile("bar.sync %0, 64;\n" ::"n"(ID) : "memory"); }
template <int ID>
__device__ __forceinline__ void bar_arrive_imm() { asm volatile("bar.arrive %0, 64;\n" ::"n"(ID) : "memory"); }
template <int BASE, int N>
__device__ __forceinline__ void named_bar_sync(int k) {      // barrier BASE + k, k in [0, N), 64 participants
    if (k == 0) bar_sync_imm<BASE>();
    if (N > 1 && k == 1) bar_sync_imm<BASE + 1>();
    if (N > 2 && k == 2) bar_sync_imm<BASE + 2>();
    if (N > 3 && k == 3) bar_sync_imm<BASE + 3>();
}
template <int BASE, int N>
__device__ __forceinline__ void named_bar_arrive(int k) {
    if (k == 0) bar_arrive_imm<BASE>();
    if (N > 1 && k == 1) bar_arrive_imm<BASE + 1>();
    if (N > 2 && k == 2) bar_arrive_imm<BASE + 2>();
    if (N > 3 && k == 3) bar_arrive_imm<BASE + 3>();
}
// explicit shared-space accesses with 32-bit addresses: keeps generic->shared address conversions (S2R + LEA per
// access, ~10% of the parser's instructions in profiles/r1_snappy_v3_parser_executor.txt) out of the serial chain
// 32-bit shared-space address of a shared-memory object, computed once by an opaque asm so that the compiler cannot
// re-materialise it (S2R SR_CgaCtaId + LEA, ~25 cycles) inside the serial loops
__device__ __forceinline__ uint32_t shared_addr(const void *p) {
    uint32_t a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }\n" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t x) {
    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(addr), "r"(x) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};\n" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// `len` >= 1 bytes from position sp of one power-of-two shared-memory buffer to position d of another (or the same), in
// stream order (a back-reference with distance < length re-reads its own output).  A rolled loop with masked addresses:
// the variants with running addresses, grouped loads or word-wise moves all measured slower on the 3-5 byte elements of
// numeric columns (profiles/experiments/r2k_lean_byte_moves.cu.txt), and unrolling costs
// instruction-cache footprint, which this kernel is sensitive to.
__device__ __forceinline__ void smem_bytes(uint32_t src_s, uint32_t smask, uint32_t sp, uint32_t dst_s, uint32_t dmask,
                                           uint32_t d, uint32_t len) {
#pragma unroll 1
    for (uint32_t i = 0; i < len; i++) sts_u8(dst_s + ((d + i) & dmask), lds_u8(src_s + ((sp + i) & smask)));
}
// Hand-over between the stages is a RENDEZVOUS on one named barrier per pair of warps: when P and A meet, P has finished
// writing batch b+1 and A has finished reading batch b (the slots are double-buffered), then both move on.  Two named
// barriers per CTA (plus barrier 0) instead of eight keep the barrier file of the SM (64) from limiting residency.
constexpr int kBarPA = 1;
constexpr int kBarAB = 2;
// output bytes of one batch: 32 elements of <= 64 bytes plus a staged long literal (< kBigLiteral)
constexpr uint32_t kMaxBatchOut = kBatchOps * 64 + kBigLiteral;
static_assert(kFlushBytes + 2 * kMaxBatchOut <= (uint32_t)kRing, "warp A must never overwrite unflushed ring bytes");

struct SnExec {                 // warp A -> warp B: the back-references of one batch, positions already assigned
    uint32_t d[kBatchOps];      // output position
    uint32_t a[kBatchOps];      // distance to the source
    uint32_t len[kBatchOps];    // 0 = this lane has no back-reference
    uint32_t dst_end;           // output position behind the batch's elements
    uint32_t big_len;           // bypassed literal that follows the batch (0 = none) ...
    uint32_t big_src;           // ... and its input offset
    uint32_t last;
    uint32_t failed;
    uint32_t dst_begin;         // output position in front of the batch
    uint32_t pad_[2];
};
static_assert(sizeof(SnExec) == 416, "SnExec layout");
constexpr uint32_t kExecHdr = kBatchOps * 12;

// ---- code that runs rarely is kept OUT of line.  The three warps of a CTA execute three different loops and seven CTAs
// share an SM: the decoder turned out to be sensitive to its instruction-cache footprint (r2l captures: a variant that
// executed 6 % fewer instructions but was 10 KB longer lost 24 % to `no_instructions` stalls).
// ring -> HBM write-through of output positions [from, to), to - from <= kRing; ring and dst agree modulo 16.
__device__ __noinline__ void snappy_ring_flush(uint8_t *dst, uint32_t ring_s, uint32_t from, uint32_t to) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t body0 = min((from + 15u) & ~15u, to), body1 = max(to & ~15u, body0);
    if (from + lane < body0) dst[from + lane] = (uint8_t)lds_u8(ring_s + ((from + lane) & kRingMask));      // < 16 bytes
    if (body1 + lane < to) dst[body1 + lane] = (uint8_t)lds_u8(ring_s + ((body1 + lane) & kRingMask));      // < 16 bytes
#pragma unroll 2
    for (uint32_t q = body0 + 16u * lane; q < body1; q += 512u) {
        const uint4 v = lds_v4(ring_s + (q & kRingMask));
        *reinterpret_cast<uint4 *>(dst + q) = v;
    }
}
__device__ __noinline__ void snappy_warp_copy_cold(uint8_t *dst, const uint8_t *src, uint32_t n) {
    coop_copy(dst, src, n, (int)(threadIdx.x & 31u), 32);
}

#ifndef PST_SNAPPY_MIN_CTAS
#define PST_SNAPPY_MIN_CTAS 10
#endif
__global__ void __launch_bounds__(kSnappyThreads, PST_SNAPPY_MIN_CTAS)
k_snappy_pages(uint8_t *__restrict__ arena, const DevPage *__restrict__ pages, const SnFrag *__restrict__ frags,
               int n_frags, const int32_t *__restrict__ multi_list, int n_multi, const uint32_t *__restrict__ frag_pos,
               uint32_t *page_flag, int32_t *status, int serial_mode) {
    // one allocation: the vector copies may read up to 15 bytes past the end of the ring (into the padding)
    __shared__ __align__(16) uint8_t smem_all[kRing + 16 + kStage];
    __shared__ __align__(16) SnBatch batches[2];
    __shared__ __align__(16) SnExec execs[2];
    __shared__ __align__(16) uint32_t quad_tab[kTabW + kTabPad];
    __shared__ __align__(16) uint8_t step_tab[kTabW + kTabPad];
    __shared__ __align__(16) uint8_t step_lut[256];
    __shared__ volatile uint32_t abort_flag;
    uint8_t *const ring = smem_all;
    uint8_t *const stage = smem_all + kRing + 16;
    // Work item.  serial_mode 0: one (page, fragment) pair of the plan's fragment list; fragments of a page whose
    // index pass raised its flag are skipped.  serial_mode 1 (fallback launch): one multi-fragment page, decoded as a
    // single stream, only if its flag is raised.
    const int li = blockIdx.x;
    int pi, frag_k = 0;
    if (serial_mode) {
        if (li >= n_multi) return;
        pi = multi_list[li];
    } else {
        if (li >= n_frags) return;
        pi = frags[li].page;
        frag_k = frags[li].k;
    }
    const DevPage pg = pages[pi];
    const bool fragmented = !serial_mode && pg.nfrag > 1;
    if (pg.multi_slot >= 0) {
        const uint32_t flag = *(volatile uint32_t *)&page_flag[pg.multi_slot];
        if (serial_mode ? flag == 0 : flag != 0) return;
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool is_parser = warp == 0;

    const uint8_t *src = arena + pg.src_off;
    uint8_t *dst = arena + pg.img_off;
    uint32_t src_n = (uint32_t)pg.comp_size;
    uint32_t dst_n = (uint32_t)pg.uncomp_size;
    if (threadIdx.x == 0) abort_flag = 0;
    // bytes an element occupies in the stream as a function of its tag; 0 = slow path (copy with 4-byte offset,
    // literal with a length suffix)
    for (int t = threadIdx.x; t < 256; t += kSnappyThreads) {
        const int kind = t & 3, t6 = t >> 2;
        step_lut[t] = (uint8_t)(kind == 0 ? (t6 < 60 ? t6 + 2 : 0) : kind == 1 ? 2 : kind == 2 ? 3 : 0);
    }
    for (int t = threadIdx.x; t < kTabPad; t += kSnappyThreads) {
        step_tab[kTabW + t] = 0;
        quad_tab[kTabW + t] = 0;
    }

    // V2 data pages: the level bytes are stored uncompressed in front of the compressed values
    if (pg.kind == PK_DATA_V2) {
        uint32_t lv = (uint32_t)(pg.def_bytes + pg.rep_bytes);
        if (warp == 1 && frag_k == 0) snappy_warp_copy_cold(dst, src, lv);
        src += lv; dst += lv; src_n -= lv; dst_n -= lv;
    }
    const uint32_t full_n = dst_n;            // the length the stream's preamble must announce
    const bool has_preamble = frag_k == 0;
    if (fragmented) {
        // [c0, c1) of the compressed values produce output bytes [k * 64 KiB, ...) (positions from k_snappy_index)
        const uint32_t c0 = frag_pos[pg.frag_first + frag_k], c1 = frag_pos[pg.frag_first + frag_k + 1];
        src += c0;
        src_n = c1 - c0;
        dst += (uint32_t)frag_k * (uint32_t)kSnappyFragment;
        dst_n = min((uint32_t)kSnappyFragment, full_n - (uint32_t)frag_k * (uint32_t)kSnappyFragment);
    }
    // Output positions are biased by the misalignment of `dst`: position p lives at ring[p & mask] and at dst[p], so
    // ring and HBM agree modulo 16 and the write-through is a plain 16-byte copy (no funnel shifts).
    const uint32_t bias = (uint32_t)((uintptr_t)dst & 15);
    dst -= bias;
    dst_n += bias;
    __syncthreads();
    if ((int32_t)src_n <= 0) return;

    // input addressing: `gin` is the 16-byte aligned base, positions are 32-bit offsets from it
    const uint8_t *gin = src - ((uintptr_t)src & 15);
    const uint32_t in_begin = (uint32_t)((uintptr_t)src & 15);
    const uint32_t in_end = in_begin + src_n;
    const uint32_t in_end16 = (in_end + 15) & ~15u;

    if (is_parser) {
        // ============================================ warp P =====================================================
        uint32_t ip = in_begin;
        uint32_t first_err = 0;
        const uint32_t stage_s = shared_addr(stage);
        const uint32_t batches_s = shared_addr(&batches[0]);
        const uint32_t quad_s = shared_addr(&quad_tab[0]), step_s = shared_addr(&step_tab[0]);
        const uint32_t lut_s = shared_addr(&step_lut[0]);
        uint32_t tab_w0 = 0, tab_end = 0;             // input window [tab_w0, tab_end) the tables describe
        if (has_preamble) {   // varint uncompressed length (a handful of bytes, read straight from global)
            uint64_t ulen = 0;
            int shift = 0;
            for (;;) {
                if (ip >= in_end || shift > 35) { first_err = 1; break; }
                uint8_t b = gin[ip++];
                ulen |= (uint64_t)(b & 0x7f) << shift;
                if (!(b & 0x80)) break;
                shift += 7;
            }
            if (!first_err && ulen != (uint64_t)full_n) first_err = 2;
        }
        uint32_t issued_end = ip & ~(kChunk - 1);   // input bytes below this have been requested
        uint32_t ready_end = issued_end;            // input bytes below this are resident in `stage`
        uint32_t keep_from = ip;                    // start of the previous batch: X may still read literals from there
        bool restart = false;
        for (uint32_t b = 0;; b++) {
            const int s = b & 1;
            if (restart) {   // after a bypassed literal the staging window moves (A is done with it: second rendezvous)
                issued_end = ready_end = ip & ~(kChunk - 1);
                keep_from = ip;
                tab_w0 = tab_end = 0;
                restart = false;
            }
            const uint32_t bt_s = batches_s + (uint32_t)s * (uint32_t)sizeof(SnBatch);
            // lane 0 decides (the flag may flip while the lanes read it): keeps the warp's control flow uniform
            const bool stop = __shfl_sync(0xffffffffu, (int)(abort_flag != 0 || first_err != 0), 0) != 0;
            if (!stop && ip < in_end && ip + kLookahead > ready_end && ready_end < in_end16) {
                // refill: every chunk that does not overwrite [keep_from, ...) ; 4 chunk slots
                const uint32_t target = ((keep_from >> kChunkShift) + 4) << kChunkShift;
                while (issued_end < target && issued_end < in_end16) {
                    uint8_t *sdst = stage + (issued_end & kStageMask);
#pragma unroll
                    for (int k = 0; k < (int)kChunk / 16 / 32; k++) {
                        uint32_t o = (uint32_t)(lane + 32 * k) * 16;
                        if (issued_end + o < in_end16) cp_async16(sdst + o, gin + issued_end + o);
                    }
                    issued_end += kChunk;
                }
                cp_async_commit();
                cp_async_wait_all();
                __syncwarp();
                ready_end = issued_end;
            }
            uint32_t n = 0, err = first_err, big = 0, big_len = 0, rare_a = 0, rare_h = 0;
            const uint32_t batch_start = ip;
            // ---- successor tables.  Finding where the elements START is the only serial part of Snappy: the position
            // of a tag depends on the element before it.  Instead of decoding tag after tag on that chain (~130 cycles
            // per element for a lone lane), all 32 lanes first compute, for EVERY byte position of a kTabW-byte window, how
            // long an element starting there would be (step_tab, via a 256-entry lookup of the tag byte) and from that
            // the lengths of the four consecutive elements that follow each position (quad_tab).  The chain is then
            // one shared-memory load + one dp4a per FOUR elements.  Positions that start a slow-path element, lie
            // behind the window or behind the end of the stream have length 0, which stops the walk there.
            bool covered = !stop && ip < in_end && ip >= tab_w0 && ip < tab_end;
            if (!stop && ip < in_end && !covered) {
                const uint32_t tag0 = lds_u8(stage_s + (ip & kStageMask));
                if (lds_u8(lut_s + tag0) != 0) {   // a window that starts with a slow-path element is not worth a build
                    const uint32_t w0 = ip & ~3u;
                    // (rolled: this runs once per ~9 batches and straight-line code would only evict the hot loops from the
                    // instruction cache)
#pragma unroll 1
                    for (int h = 0; h < kTabW / 128; h += 2) {
                        uint32_t word[2], e[8];
#pragma unroll
                        for (int q = 0; q < 2; q++)
                            word[q] = lds_u32(stage_s + ((w0 + 4u * ((uint32_t)lane + 32u * (uint32_t)(h + q))) & kStageMask));
#pragma unroll
                        for (int q = 0; q < 8; q++) e[q] = lds_u8(lut_s + ((word[q >> 2] >> (8 * (q & 3))) & 0xffu));
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const uint32_t wi = (uint32_t)lane + 32u * (uint32_t)(h + q);
                            uint32_t st = e[4 * q] | (e[4 * q + 1] << 8) | (e[4 * q + 2] << 16) | (e[4 * q + 3] << 24);
                            const int32_t nv = (int32_t)(in_end - (w0 + 4u * wi));
                            if (nv < 4) st = nv <= 0 ? 0u : (st & ((1u << (8 * nv)) - 1u));
                            sts_u32(step_s + 4u * wi, st);
                        }
                    }
                    __syncwarp();
#pragma unroll 1
                    for (int k0 = 0; k0 < kTabW / 32; k0 += 8) {
                        uint32_t s1[8], s2[8], s3[8], s4[8], pp[8];
                        const uint32_t p0 = (uint32_t)lane + 32u * (uint32_t)k0;
#pragma unroll
                        for (int q = 0; q < 8; q++) s1[q] = lds_u8(step_s + p0 + 32u * q);
#pragma unroll
                        for (int q = 0; q < 8; q++) { pp[q] = p0 + 32u * q + s1[q]; s2[q] = lds_u8(step_s + pp[q]); }
#pragma unroll
                        for (int q = 0; q < 8; q++) { pp[q] += s2[q]; s3[q] = lds_u8(step_s + pp[q]); }   // zeros propagate:
#pragma unroll
                        for (int q = 0; q < 8; q++) { pp[q] += s3[q]; s4[q] = lds_u8(step_s + pp[q]); }   // a 0 re-reads itself
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            sts_u32(quad_s + 4u * (p0 + 32u * q), s1[q] | (s2[q] << 8) | (s3[q] << 16) | (s4[q] << 24));
                    }
                    __syncwarp();
                    tab_w0 = w0;
                    tab_end = w0 + kTabW;
                    covered = true;
                }
            }
            if (lane == 0 && !stop && ip < in_end) {
                uint32_t rec_s = bt_s;
                const uint32_t rec_end = bt_s + kHops * 8;
                bool at_stop = !covered;      // the walk ended on a position whose element length is 0
                if (covered) {
                    const uint32_t qbase = quad_s - (tab_w0 << 2);
                    for (;;) {
                        const uint32_t q = lds_u32(qbase + (ip << 2));
                        sts_v2(rec_s, ip, q);
                        ip = __dp4a(q, 0x01010101u, ip);          // ip += the four element lengths
                        if (q < 0x01000000u) {                    // fewer than four elements: the walk stops here
                            if (q != 0) rec_s += 8;
                            at_stop = true;
                            break;
                        }
                        rec_s += 8;
                        if (rec_s == rec_end) break;
                    }
                }
                n = (rec_s - bt_s) >> 3;
                if (at_stop && ip < in_end && (!covered || ip < tab_end) && n < (uint32_t)kHops) {
                    // slow-path element (behind a window end the next batch builds a new table instead)
#define IN(p) lds_u8(stage_s + ((p) & kStageMask))
                    const uint32_t tag = IN(ip);
                    const uint32_t t6 = tag >> 2;
                    if ((tag & 3) == 3) {
                        rare_a = IN(ip + 1) | (IN(ip + 2) << 8) | (IN(ip + 3) << 16) | (IN(ip + 4) << 24);
                        rare_h = (t6 + 1) | (1u << 24);
                        ip += 5;
                    } else if ((tag & 3) == 0 && t6 >= 60) {
                        const uint32_t nb = t6 - 59;
                        uint32_t v = 0;
#pragma unroll 1
                        for (uint32_t i = 0; i < nb; i++) v |= IN(ip + 1 + i) << (8 * i);
                        const uint32_t len = v + 1;
                        const uint32_t p0 = ip + 1 + nb;
                        if (p0 > in_end || len > in_end - p0) {
                            err = 4;
                        } else {
                            rare_a = p0;
                            if (len >= kBigLiteral) {
                                rare_h = 2u << 24;
                                big = 1;
                                big_len = len;
                            } else {
                                rare_h = len | (3u << 24);   // kind 3: staged literal with explicit length
                            }
                            ip = p0 + len;
                        }
                    } else {
                        err = 6;   // unreachable: a fast-path tag has a non-zero length inside the window
                    }
#undef IN
                }
                if (!err && ip > in_end) err = 5;   // a slow-path element ran past the end of the stream
            }
            n = __shfl_sync(0xffffffffu, n, 0);
            ip = __shfl_sync(0xffffffffu, ip, 0);
            err = __shfl_sync(0xffffffffu, err, 0);
            big = __shfl_sync(0xffffffffu, big, 0);
            big_len = __shfl_sync(0xffffffffu, big_len, 0);
            const bool last = stop || err != 0 || ip >= in_end;
            if (lane == 0) {
                sts_v4(bt_s + kHdrOff, n, last ? 1u : 0u, err, big_len);
                sts_v2(bt_s + kHdrOff + 16, rare_a, rare_h);
            }
            keep_from = batch_start;
            __syncwarp();
            __threadfence_block();
            bar_sync_imm<kBarPA>();                 // hand-over: A starts on this batch, its previous one is finished
            if (last) return;
            if (big) {
                bar_sync_imm<kBarPA>();             // A has moved this batch's literals out of the staging window
                restart = true;
            }
        }
    } else if (warp == 1) {
        // ============================================ warp A =====================================================
        // decodes the 32 elements of a batch (one per lane), assigns output positions with a warp scan, validates,
        // moves the literal bytes staging -> ring and hands the back-references to warp B
        const uint32_t ring_s = shared_addr(ring), stage_s = shared_addr(stage), batches_s = shared_addr(&batches[0]);
        const uint32_t execs_s = shared_addr(&execs[0]);
        uint32_t dst0 = bias;         // output position of the next batch
        bool failed = false;
        for (uint32_t b = 0;; b++) {
            const int s = b & 1;
            bar_sync_imm<kBarPA>();
            const uint32_t bt_s = batches_s + (uint32_t)s * (uint32_t)sizeof(SnBatch);
            const uint32_t ex_s = execs_s + (uint32_t)s * (uint32_t)sizeof(SnExec);
            const uint4 hdr = lds_v4(bt_s + kHdrOff);
            const uint32_t n = hdr.x, last = hdr.y, perr = hdr.z, big_len = hdr.w;
            if (perr && !failed) {
                failed = true;
                if (lane == 0) report_error(status, DE_SNAPPY_CORRUPT, pi, (int)perr);
            }
            // every lane decodes its own element from the staged bytes (the parser only located the hops): lane l
            // takes element l & 3 of hop l >> 2; the slow-path element, if any, goes to the lane after the last hop
            uint32_t kind = 0, a = 0, len = 0;
            bool have = false, overrun = false;
            const uint2 rare = lds_v2(bt_s + kHdrOff + 16);
            if (!failed) {
                const int h = lane >> 2, j = lane & 3;
                if (h < (int)n) {
                    const uint2 hp = lds_v2(bt_s + h * 8);
                    if ((hp.y >> (8 * j)) & 0xffu) {
                        const uint32_t pos = __dp4a(hp.y & ((1u << (8 * j)) - 1u), 0x01010101u, hp.x);
                        const uint32_t tag = lds_u8(stage_s + (pos & kStageMask));
                        const uint32_t b1 = lds_u8(stage_s + ((pos + 1) & kStageMask));
                        const uint32_t b2 = lds_u8(stage_s + ((pos + 2) & kStageMask));
                        const uint32_t t6 = tag >> 2;
                        kind = tag & 3;
                        have = true;
                        overrun = pos + ((hp.y >> (8 * j)) & 0xffu) > in_end;
                        if (kind == 0) { len = t6 + 1; a = pos + 1; }
                        else if (kind == 1) { len = (t6 & 7) + 4; a = ((tag >> 5) << 8) | b1; }
                        else { len = t6 + 1; a = b1 | (b2 << 8); kind = 1; }
                    }
                } else if (lane == 4 * (int)n && rare.y != 0) {
                    a = rare.x;
                    kind = rare.y >> 24;           // 1 copy-4, 2 bypassed big literal, 3 long staged literal
                    len = kind == 2 ? 0u : (rare.y & 0xffffffu);
                    if (kind == 3) kind = 0;
                    have = true;
                }
            }
            const uint32_t incl = warp_incl_scan(len, lane);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t d = dst0 + incl - len;          // output position of this lane's element
            const bool is_copy = have && kind == 1;
            const bool has_big = !failed && (rare.y >> 24) == 2;
            if (!failed) {
                const bool cross = is_copy && a > d - bias;   // source in front of this stream's first output byte
                const bool bad = overrun || (d + len > dst_n) || (is_copy && a == 0) || (cross && frag_k == 0);
                if (__any_sync(0xffffffffu, bad)) {
                    failed = true;
                    if (lane == 0) report_error(status, DE_SNAPPY_CORRUPT, pi, 8);
                } else if (__any_sync(0xffffffffu, cross)) {
                    // a back-reference into an earlier fragment: legal Snappy, just not what the reference compressor
                    // emits.  The page is handed to the serial fallback launch, which rewrites its whole image.
                    failed = true;
                    if (lane == 0) *(volatile uint32_t *)&page_flag[pg.multi_slot] = 2;
                } else if (has_big && dst0 + total + big_len > dst_n) {
                    failed = true;
                    if (lane == 0) report_error(status, DE_SNAPPY_CORRUPT, pi, 4);
                } else if (last && dst0 + total + (has_big ? big_len : 0u) != dst_n) {
                    failed = true;
                    if (lane == 0) report_error(status, DE_SNAPPY_CORRUPT, pi, 9);
                }
                if (failed) abort_flag = 1;
            }
            // (the rendezvous with B below keeps this warp exactly one batch ahead of B, which is what B's "source still in
            // the ring" test assumes)
            if (!failed) {
                // ---- literals: staging -> ring.  Short ones per lane, longer ones by the whole warp.
                const bool is_lit = have && kind == 0;
                if (is_lit && len <= 16) smem_bytes(stage_s, kStageMask, a, ring_s, kRingMask, d, len);
                uint32_t longs = __ballot_sync(0xffffffffu, is_lit && len > 16);
                while (longs) {
                    const int l = __ffs(longs) - 1;
                    longs &= longs - 1;
                    const uint32_t bl = __shfl_sync(0xffffffffu, len, l);
                    const uint32_t bd = __shfl_sync(0xffffffffu, d, l);
                    const uint32_t ba = __shfl_sync(0xffffffffu, a, l);
                    for (uint32_t i = lane; i < bl; i += 32)
                        sts_u8(ring_s + ((bd + i) & kRingMask), lds_u8(stage_s + ((ba + i) & kStageMask)));
                }
            }
            sts_u32(ex_s + 4u * lane, d);
            sts_u32(ex_s + 128u + 4u * lane, a);
            sts_u32(ex_s + 256u + 4u * lane, (is_copy && !failed) ? len : 0u);
            if (lane == 0) {
                sts_v4(ex_s + kExecHdr, dst0 + total, has_big && !failed ? big_len : 0u, rare.x, last);
                sts_v2(ex_s + kExecHdr + 16, failed ? 1u : 0u, dst0);
            }
            if (!failed) dst0 += total + (has_big ? big_len : 0u);
            __syncwarp();
            __threadfence_block();
            if (big_len != 0 && !last) bar_sync_imm<kBarPA>();   // P may move the staging window now
            bar_sync_imm<kBarAB>();                 // hand-over: B starts on this batch, its previous one is finished
            if (last) return;
            if (has_big && !failed) {
                // the bypassed literal moves the output position by an arbitrary amount, so the next batch's literals
                // would land on ring slots warp B may still read: wait until B is done with this batch
                bar_sync_imm<kBarAB>();
            }
        }
    } else {
        // ============================================ warp B =====================================================
        // back-references ring -> ring in dependency rounds, write-through of the ring to HBM, bypassed literals
        const uint32_t ring_s = shared_addr(ring), execs_s = shared_addr(&execs[0]);
        uint32_t flushed = bias;      // output positions below this are already in global memory
        uint32_t valid_from = bias;      // output positions below this are not in the ring (bypassed literal)
        auto flush_to = [&](uint32_t t) {
            if (flushed < t) {
                snappy_ring_flush(dst, ring_s, flushed, t);
                flushed = t;
            }
        };
        for (uint32_t b = 0;; b++) {
            const int s = b & 1;
            bar_sync_imm<kBarAB>();
            const uint32_t ex_s = execs_s + (uint32_t)s * (uint32_t)sizeof(SnExec);
            const uint4 hdr = lds_v4(ex_s + kExecHdr);
            const uint32_t dst_end = hdr.x, big_len = hdr.y, big_src = hdr.z, last = hdr.w;
            const uint2 hdr2 = lds_v2(ex_s + kExecHdr + 16);
            const bool failed = hdr2.x != 0;
            const uint32_t dst_begin = hdr2.y;
            if (!failed) {
                const uint32_t d = lds_u32(ex_s + 4u * lane), a = lds_u32(ex_s + 128u + 4u * lane);
                const uint32_t len = lds_u32(ex_s + 256u + 4u * lane);
                const bool is_copy = len != 0;
                const uint32_t sp = d - a;                                   // source position
                const uint32_t src_end = min(sp + len, d);                   // bytes >= d are produced by the lane itself
                // warp A may already be placing the literals of the next batch: the oldest kMaxBatchOut bytes of the
                // ring are not trusted
                const bool in_ring = sp >= valid_from && dst_end - sp <= (uint32_t)kRing - kMaxBatchOut;
                // the common case first: a source that ends in front of the batch depends on nothing in it
                const bool early = is_copy && in_ring && src_end <= dst_begin;
                if (early) smem_bytes(ring_s, kRingMask, d - a, ring_s, kRingMask, d, len);
                // a source outside the ring that this warp has already written to HBM depends on nothing either: every lane
                // fetches its own.  (5 % of the C2 back-references.  The first version handled them one lane at a time in
                // the dependency rounds with a flush in front - 3.8 % of the kernel's instructions and most of this warp's
                // batch-to-batch variance; fetching them in warp A like literals was 6 % slower than this, A being the
                // busiest of the three warps.)
                const bool far_done = is_copy && !in_ring && sp + len <= flushed;
                if (far_done) {
                    // (loads first: one L2 round trip for the first eight bytes instead of one per byte)
                    const uint8_t *g = dst + sp;
                    uint32_t v[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = (uint32_t)k < len ? (uint32_t)__ldcg(g + k) : 0u;
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if ((uint32_t)k < len) sts_u8(ring_s + ((d + (uint32_t)k) & kRingMask), v[k]);
#pragma unroll 1
                    for (uint32_t i = 8; i < len; i++) sts_u8(ring_s + ((d + i) & kRingMask), (uint32_t)__ldcg(g + i));
                }
                uint32_t pending = __ballot_sync(0xffffffffu, is_copy && !early && !far_done);
                __syncwarp();
                while (pending) {
                    const int first = __ffs(pending) - 1;
                    const uint32_t done_pos = __shfl_sync(0xffffffffu, d, first);   // everything below is complete
                    const bool mine = (pending >> lane) & 1;
                    bool ready = mine && (lane == first || src_end <= done_pos);
                    if (ready && !in_ring && lane != first) ready = false;       // far sources wait for their turn
                    const bool far_first = __shfl_sync(0xffffffffu, (int)(ready && !in_ring), first) != 0;
                    if (far_first) {
                        // the source left the ring (or was bypassed): complete output below done_pos goes to HBM first
                        flush_to(done_pos);
                        __syncwarp();
                        if (lane == first) {
#pragma unroll 1
                            for (uint32_t i = 0; i < len; i++) {
                                const uint32_t q = sp + i;
                                ring[(d + i) & kRingMask] = q < done_pos ? dst[q] : ring[q & kRingMask];
                            }
                        }
                    }
                    // sequential per lane: an overlapping copy (offset < length) re-reads its own bytes
                    if (ready && in_ring) smem_bytes(ring_s, kRingMask, d - a, ring_s, kRingMask, d, len);
                    pending &= ~__ballot_sync(0xffffffffu, ready);
                    __syncwarp();
                }
                uint32_t dst0 = dst_end;
                if (big_len) {      // bypassed big literal (always behind the last element of its batch)
                    flush_to(dst0);
                    snappy_warp_copy_cold(dst + dst0, gin + big_src, big_len);
                    dst0 += big_len;
                    flushed = dst0;
                    valid_from = dst0;
                    __syncwarp();
                }
                if (last || dst0 - flushed >= kFlushBytes) {
                    flush_to(dst0);
                    __syncwarp();
                }
            }
            if (last) return;
            if (big_len != 0 && !failed) bar_sync_imm<kBarAB>();   // tell A that the ring may be reused out of order
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2a  Snappy fragment index.  The reference compressor (snappy::RawCompress, which Arrow's Parquet writer calls once
// per page) works on 64 KiB blocks of input and restarts its match window at every block, so a page is a concatenation
// of independently decodable fragments -- but the stream does not say where they start.  This kernel finds out:
// it walks the element chain of a page (multi-fragment pages only) WITHOUT producing output and records the
// compressed offset at which the output position reaches every multiple of 64 KiB.  k_snappy_pages then decodes all
// fragments of all pages in parallel.  A page whose elements straddle a 64 KiB boundary, or whose stream is damaged,
// gets its flag raised and is decoded as one stream by the serial fallback launch (which also reports the error).
//
// The walk is the serial part, so it is made as short as possible: twelve builder warps compute, for every byte position
// of a 1 KiB input window, {bytes consumed, bytes produced} by the NEXT 64 ELEMENTS starting there (lookup of the tag
// byte, then six rounds of pointer doubling T2[p] = T[p] + T[p + consumed(T[p])] in shared memory).  The walker warp
// then needs one shared-memory load and two adds per 64 elements.  Positions holding a slow-path tag, behind the window
// or behind the stream have {0, 0}, which stalls the chain there.
// ---------------------------------------------------------------------------------------------------------------
// What bounds it (r2t/r2w captures): the doubling rounds keep the SM's shared-memory pipe ~90 % busy (their second load
// is data dependent: ~3.5-way bank conflicts), the walker executes ~140 dependent instructions per window, and both
// come to ~1.3 k cycles per window - a 1 MiB page (950 windows; the sixteen dictionary pages of the C2 int64 columns)
// takes 0.65 ms whether there are four builders (round 1: 0.84 ms with named barriers and 4 rounds) or twelve, 5 rounds
// or 6.  Pages are launched longest first so that these do not start in the second wave.  The hand-over uses mbarriers
// in shared memory (one full / one empty barrier per builder): named barriers are limited to 16 per CTA.
constexpr int kIdxBuilders = 12;
constexpr int kIdxThreads = 32 * (kIdxBuilders + 1);
constexpr int kIdxW = 1024;
constexpr int kIdxPad = 64;
#ifndef PST_IDX_ROUNDS
#define PST_IDX_ROUNDS 6
#endif
constexpr int kIdxRounds = PST_IDX_ROUNDS;          // a table entry covers 2^kIdxRounds elements: the walker (the serial part) does one
                                       // shared-memory load per 64 elements; the twelve builders absorb the extra rounds
constexpr size_t kIdxSmemBytes = (size_t)2 * kIdxBuilders * (kIdxW + kIdxPad) * sizeof(uint32_t);

__device__ __forceinline__ void idx_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void idx_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void idx_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

// thread-block cluster primitives (the big-page variant of the index kernel)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t cta_addr, uint32_t rank) {      // shared::cta -> shared::cluster
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(a) : "r"(cta_addr), "r"(rank));
    return a;
}
__device__ __forceinline__ void cluster_sts_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;\n" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_mbar_arrive(uint32_t cluster_bar) {          // release: publishes the caller's stores
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void cluster_mbar_arrive_relaxed(uint32_t cluster_bar) {  // a pure "I am done reading" signal
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_bar) : "memory");
}
// Waits on a LOCAL barrier whose arrivals come from other CTAs of the cluster.  A thread suspended in try_wait is not
// woken by a remote arrival before its time limit runs out (r2w capture: builders and walker each waited ~6 us per
// hand-over with the default limit and the cluster kernel was no faster than one SM), so the builders pass a short limit
// and the walker - alone on its SM - polls with test_wait.
__device__ __forceinline__ void cluster_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(200u)       // suspend-time limit in ns
            : "memory");
    }
}
__device__ __forceinline__ void cluster_mbar_poll(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// tag -> (bytes consumed * 4) | (bytes produced << 16); 0 = slow path (copy with a 4-byte offset, literal with a length suffix)
__device__ __forceinline__ uint32_t idx_lut_entry(uint32_t t) {
    const uint32_t kind = t & 3, t6 = t >> 2;
    uint32_t used = 0, made = 0;
    if (kind == 0) { if (t6 < 60) { used = t6 + 2; made = t6 + 1; } }
    else if (kind == 1) { used = 2; made = (t6 & 7) + 4; }
    else if (kind == 2) { used = 3; made = t6 + 1; }
    return (used << 2) | (made << 16);
}

// One window of the successor table (one warp): every position of [w0, w0 + kIdxW) as if a tag started there, then
// kIdxRounds rounds of pointer doubling between the warp's two buffers.  The last round is written to `final_s`: the
// warp's own `tab` buffer (REMOTE false; the rounds are arranged to end there) or a table slot of the walker CTA of the
// cluster (REMOTE true, a shared::cluster address).
template <bool REMOTE>
__device__ __forceinline__ void idx_build_window(uint32_t tab_s, uint32_t tmp_s, uint32_t lut_s, const uint8_t *gin,
                                                 uint32_t w0, uint32_t in_end, uint32_t in_end16, uint32_t final_s,
                                                 int lane) {
    // (the asm accessors are volatile, i.e. executed in program order: loads are issued in batches of eight before
    // their results are used, otherwise every position would pay the full LDS latency)
    {
        uint32_t word[kIdxW / 128];
#pragma unroll
        for (int k = 0; k < kIdxW / 128; k++) {
            const uint32_t pos = w0 + 4u * ((uint32_t)lane + 32u * (uint32_t)k);
            word[k] = pos < in_end16 ? __ldg(reinterpret_cast<const uint32_t *>(gin + pos)) : 0u;
        }
#pragma unroll
        for (int h = 0; h < kIdxW / 128; h += 2) {
            uint32_t e[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
                e[q] = lds_u32(lut_s + (((word[h + (q >> 2)] >> (8 * (q & 3))) & 0xffu) << 2));
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t wi = (uint32_t)lane + 32u * (uint32_t)(h + q);
                const int32_t nv = (int32_t)(in_end - (w0 + 4u * wi));      // stream bytes in this word
                if (nv < 4) {
                    if (nv < 1) e[4 * q] = 0;
                    if (nv < 2) e[4 * q + 1] = 0;
                    if (nv < 3) e[4 * q + 2] = 0;
                    e[4 * q + 3] = 0;
                }
                sts_v4((kIdxRounds & 1 ? tmp_s : tab_s) + 16u * wi, e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
            }
        }
    }
    __syncwarp();
    // 2, 4, ... 2^kIdxRounds elements by pointer doubling (the low half of an entry is already a byte offset); the
    // buffers alternate and the last round lands in `tab` (or in the remote slot)
    uint32_t from_s = kIdxRounds & 1 ? tmp_s : tab_s, to_s = kIdxRounds & 1 ? tab_s : tmp_s;
#pragma unroll 1
    for (int r = 0; r < kIdxRounds; r++) {
        const bool remote = REMOTE && r == kIdxRounds - 1;
#pragma unroll 1
        for (int k0 = 0; k0 < kIdxW / 32; k0 += 8) {
            uint32_t e[8], f[8];
            const uint32_t a0 = from_s + 4u * ((uint32_t)lane + 32u * (uint32_t)k0);
#pragma unroll
            for (int q = 0; q < 8; q++) e[q] = lds_u32(a0 + 128u * q);
#pragma unroll
            for (int q = 0; q < 8; q++) f[q] = lds_u32(a0 + 128u * q + (e[q] & 0xffffu));   // e == 0 re-reads itself
            if (remote) {
#pragma unroll
                for (int q = 0; q < 8; q++) cluster_sts_u32(a0 - from_s + final_s + 128u * q, e[q] + f[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; q++) sts_u32(a0 - from_s + to_s + 128u * q, e[q] + f[q]);
            }
        }
        __syncwarp();
        const uint32_t t = from_s; from_s = to_s; to_s = t;
    }
}

// The walker's state and its pass over one window: follows the chain through the window's table (`base` is the table's
// shared-memory address minus 4 * the window's first position) and records fragment boundaries.  The whole warp runs
// it on identical values (every load is a broadcast) and lane 0 writes the results: that way the elements around a
// fragment boundary, which have to be taken one at a time from the stream itself, can be fetched by all lanes at once.
struct IdxWalk {
    uint32_t ip, op, next_b, k, flag;
};
__device__ __forceinline__ void idx_walk_window(IdxWalk &w, uint32_t base, uint32_t wend, uint32_t lut_s,
                                                const uint8_t *gin, uint32_t in_begin, uint32_t in_end,
                                                uint32_t *my_pos, uint32_t nfrag, int lane) {
    uint32_t ip = w.ip, op = w.op, next_b = w.next_b, k = w.k, flag = w.flag;
    for (;;) {
        // the chain: one LDS + two adds per 2^kIdxRounds elements
        uint32_t a = base + (ip << 2), e, adv;
        // far from the next fragment boundary (an entry produces at most kIdxEntryOut bytes) the chain needs no
        // test per hop: LDS -> mask -> add, four hops per round; a zero entry is a fixed point, so a stall
        // inside the round shows in its last entry and neither `a` nor `op` moved past it
        constexpr uint32_t kIdxEntryOut = 64u << kIdxRounds;
        while (next_b - op > 4u * kIdxEntryOut) {
            const uint32_t e1 = lds_u32(a);
            a += e1 & 0xffffu;
            const uint32_t e2 = lds_u32(a);
            a += e2 & 0xffffu;
            const uint32_t e3 = lds_u32(a);
            a += e3 & 0xffffu;
            const uint32_t e4 = lds_u32(a);
            a += e4 & 0xffffu;
            op += (e1 >> 16) + (e2 >> 16) + (e3 >> 16) + (e4 >> 16);
            if ((e4 & 0xffffu) == 0) break;
        }
        for (;;) {
            bool out = false;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                e = lds_u32(a);
                adv = e & 0xffffu;
                if (adv == 0 || op + (e >> 16) >= next_b) { out = true; break; }
                a += adv;
                op += e >> 16;
            }
            if (out) break;
        }
        ip = (a - base) >> 2;
        if (adv == 0) {
            if (ip >= wend) break;
            // slow-path element: its length is in the stream, not in the tag
            const uint32_t tag = gin[ip];
            uint32_t used, made;
            if ((tag & 3) == 3) { used = 5; made = (tag >> 2) + 1; }
            else if ((tag & 3) == 0 && (tag >> 2) >= 60) {
                const uint32_t nb = (tag >> 2) - 59;
                uint32_t v = 0;
                for (uint32_t i = 0; i < nb && ip + 1 + i < in_end; i++) v |= (uint32_t)gin[ip + 1 + i] << (8 * i);
                made = v + 1;
                used = 1 + nb + made;
                if (made > in_end - ip || used > in_end - ip) { flag = 1; break; }
            } else { flag = 1; break; }
            if (op == next_b && k < nfrag) {
                if (lane == 0) my_pos[k] = ip - in_begin;
                k++;
                next_b += kSnappyFragment;
            } else if (op < next_b && made > next_b - op) { flag = 1; break; }
            ip += used;
            op += made;
            if (ip >= wend) break;      // a long literal usually leaves the window (and several more)
            continue;
        }
        const uint32_t op2 = op + (e >> 16);
        if (op2 >= next_b) {
            // a fragment boundary lies in (or right behind) these elements: take them one at a time.  Their tags come from
            // the stream in HBM - read serially that was ~30 dependent loads of ~0.3 us per boundary, a quarter of the
            // time of a 1 MiB page (r2w capture) - so the warp fetches 256 bytes at once (8 per lane) and the tags are
            // picked out of the lanes' registers.
            const uint32_t hop_end = ip + (adv >> 2);
            uint32_t st0 = ip, st_lo = 0, st_hi = 0;
            bool staged = false;
            while (ip < hop_end) {
                if (!staged || ip - st0 >= 256u) {
                    st0 = ip;
                    st_lo = st_hi = 0;
                    const uint32_t q = st0 + 8u * (uint32_t)lane;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (q + i < in_end) st_lo |= (uint32_t)gin[q + i] << (8 * i);
                        if (q + 4 + i < in_end) st_hi |= (uint32_t)gin[q + 4 + i] << (8 * i);
                    }
                    staged = true;
                }
                const uint32_t o = ip - st0;
                const uint32_t wlo = __shfl_sync(0xffffffffu, st_lo, (int)(o >> 3));
                const uint32_t whi = __shfl_sync(0xffffffffu, st_hi, (int)(o >> 3));
                const uint32_t tag = (((o & 4u) ? whi : wlo) >> (8u * (o & 3u))) & 0xffu;
                const uint32_t one = lds_u32(lut_s + (tag << 2));
                const uint32_t made = one >> 16;
                if (one == 0) { flag = 1; break; }   // cannot happen: the table came from the same bytes
                if (op == next_b && k < nfrag) {
                    if (lane == 0) my_pos[k] = ip - in_begin;
                    k++;
                    next_b += kSnappyFragment;
                } else if (op < next_b && made > next_b - op) { flag = 1; break; }
                ip += (one & 0xffffu) >> 2;
                op += made;
            }
            if (flag) break;
            continue;
        }
        ip += adv >> 2;
        op = op2;
    }
    w.ip = ip; w.op = op; w.next_b = next_b; w.k = k; w.flag = flag;
}
// the stream's preamble: varint uncompressed length, which must match the page header
__device__ __forceinline__ void idx_walk_preamble(IdxWalk &w, const uint8_t *gin, uint32_t in_end, uint32_t dst_n) {
    uint64_t ulen = 0;
    int shift = 0;
    for (;;) {
        if (w.ip >= in_end || shift > 35) { w.flag = 1; break; }
        const uint8_t b = gin[w.ip++];
        ulen |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
    }
    if (ulen != (uint64_t)dst_n) w.flag = 1;
}

__global__ void __launch_bounds__(kIdxThreads)
k_snappy_index(const uint8_t *__restrict__ arena, const DevPage *__restrict__ pages,
               const int32_t *__restrict__ multi_list, int n_multi, uint32_t *__restrict__ frag_pos,
               uint32_t *__restrict__ page_flag) {
    extern __shared__ __align__(16) uint8_t idx_smem[];
    uint32_t (*tab)[kIdxW + kIdxPad] = reinterpret_cast<uint32_t (*)[kIdxW + kIdxPad]>(idx_smem);
    uint32_t (*tmp)[kIdxW + kIdxPad] = tab + kIdxBuilders;
    __shared__ __align__(8) uint64_t bar_full[kIdxBuilders], bar_empty[kIdxBuilders];
    __shared__ uint32_t lut[256];                  // tag -> (bytes consumed * 4) | (bytes produced << 16); 0 = slow path
    __shared__ volatile uint32_t walker_ip;        // lower bound of the walker's position (lets builders skip windows)
    __shared__ volatile uint32_t give_up;
    const int li = blockIdx.x;
    if (li >= n_multi) return;
    const int pi = multi_list[li];
    const DevPage pg = pages[pi];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const uint8_t *src = arena + pg.src_off;
    uint32_t src_n = (uint32_t)pg.comp_size;
    uint32_t dst_n = (uint32_t)pg.uncomp_size;
    if (pg.kind == PK_DATA_V2) {
        const uint32_t lv = (uint32_t)(pg.def_bytes + pg.rep_bytes);
        src += lv; src_n -= lv; dst_n -= lv;
    }
    const uint8_t *gin = src - ((uintptr_t)src & 15);
    const uint32_t in_begin = (uint32_t)((uintptr_t)src & 15);
    const uint32_t in_end = in_begin + src_n;
    const uint32_t in_end16 = (in_end + 15) & ~15u;

    for (int t = threadIdx.x; t < 256; t += kIdxThreads) lut[t] = idx_lut_entry((uint32_t)t);
    for (int t = threadIdx.x; t < kIdxBuilders * kIdxPad; t += kIdxThreads) {
        tab[t / kIdxPad][kIdxW + t % kIdxPad] = 0;
        tmp[t / kIdxPad][kIdxW + t % kIdxPad] = 0;
    }
    if (threadIdx.x == 0) {
        walker_ip = in_begin;
        give_up = 0;
        for (int b = 0; b < kIdxBuilders; b++) {
            idx_mbar_init((uint32_t)__cvta_generic_to_shared(&bar_full[b]), 1);
            idx_mbar_init((uint32_t)__cvta_generic_to_shared(&bar_empty[b]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const uint32_t nwin = (in_end + kIdxW - 1) / kIdxW;      // windows over positions [0, in_end), aligned to kIdxW

    if (warp < kIdxBuilders) {
        // ============================================ builders ===================================================
        const uint32_t tab_s = shared_addr(&tab[warp][0]), tmp_s = shared_addr(&tmp[warp][0]);
        const uint32_t lut_s = shared_addr(&lut[0]);
        const uint32_t my_full = (uint32_t)__cvta_generic_to_shared(&bar_full[warp]);
        const uint32_t my_empty = (uint32_t)__cvta_generic_to_shared(&bar_empty[warp]);
        uint32_t use = 0;          // how often this builder's table slot has been filled
        for (uint32_t j = (uint32_t)warp; j < nwin; j += kIdxBuilders, use++) {
            if (use) idx_mbar_wait(my_empty, (use - 1) & 1u);      // the walker is done with the previous window of the slot
            const uint32_t w0 = j * kIdxW;
            const bool skip = __shfl_sync(0xffffffffu, (int)(give_up != 0 || w0 + kIdxW <= walker_ip), 0) != 0;
            if (!skip) idx_build_window<false>(tab_s, tmp_s, lut_s, gin, w0, in_end, in_end16, tab_s, lane);
            __syncwarp();
            if (lane == 0) idx_mbar_arrive(my_full);       // release: the table of window j is complete
        }
    } else {
        // ============================================ walker =====================================================
        const uint32_t lut_s = shared_addr(&lut[0]);
        IdxWalk w{in_begin, 0u, (uint32_t)kSnappyFragment, 1u, 0u};
        uint32_t *const my_pos = frag_pos + pg.frag_first;
        if (lane == 0) my_pos[0] = 0;
        idx_walk_preamble(w, gin, in_end, dst_n);
        if (lane == 0 && w.flag) give_up = 1;
        // (addresses once, slot and phase by counting: the per-window overhead of this loop is on the serial path)
        const uint32_t tab_s0 = shared_addr(&tab[0][0]), full_s0 = (uint32_t)__cvta_generic_to_shared(&bar_full[0]);
        const uint32_t empty_s0 = (uint32_t)__cvta_generic_to_shared(&bar_empty[0]);
        uint32_t b = 0, phase = 0;
        for (uint32_t j = 0; j < nwin; j++) {
            idx_mbar_wait(full_s0 + 8u * b, phase);
            const uint32_t wend = min((j + 1) * (uint32_t)kIdxW, in_end);
            if (!w.flag && w.ip < wend) {
                idx_walk_window(w, tab_s0 + b * (uint32_t)sizeof(tab[0]) - ((j * (uint32_t)kIdxW) << 2), wend, lut_s, gin,
                                in_begin, in_end, my_pos, (uint32_t)pg.nfrag, lane);
                if (lane == 0 && w.flag) give_up = 1;
            }
            if (lane == 0) walker_ip = w.ip;
            __syncwarp();
            if (lane == 0 && j + kIdxBuilders < nwin) idx_mbar_arrive(empty_s0 + 8u * b);
            if (++b == (uint32_t)kIdxBuilders) { b = 0; phase ^= 1u; }
        }
        const uint32_t ip = w.ip, op = w.op, k = w.k;
        uint32_t flag = w.flag;
        if (lane == 0) {
            if (!flag && (ip != in_end || op != dst_n || k != (uint32_t)pg.nfrag)) flag = 1;
            if (!flag) my_pos[pg.nfrag] = src_n;
            page_flag[pg.multi_slot] = flag;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K2a for BIG pages: the same index, one thread-block CLUSTER of four CTAs (four SMs) per page.
//
// In k_snappy_index the twelve builders and the walker share the shared-memory pipe of one SM: the doubling rounds keep it
// ~90 % busy (the data-dependent second load of a round is a ~3.5-way bank conflict), so the walker's chain - five
// dependent shared-memory loads per window - waits in the same queue and a 1 MiB page (950 windows; C2 has sixteen of
// them, the int64 dictionary pages) takes 0.65 ms however many builders there are (r2t capture: 36 % of the builders'
// samples are waits for the walker, the walker is never idle).  Here the walker has an SM to itself: CTA 0 of the cluster
// holds one table slot per builder (36 x 4.25 KiB) and runs the walker warp; CTAs 1-3 run twelve builders each, which do
// their doubling rounds in their own shared memory and write the LAST round straight into their slot in CTA 0 through
// distributed shared memory (st.shared::cluster), then arrive on the slot's mbarrier in CTA 0
// (mbarrier.arrive.release.cluster); the walker hands a slot back by arriving on the builder's own barrier in the
// builder's CTA.  Pages below kIdxBigPage stay with the single-CTA kernel, which spends fewer SM-cycles per window.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kIdxCluster = 4;
constexpr int kIdxSlots = (kIdxCluster - 1) * kIdxBuilders;
constexpr size_t kIdxClusterSmemBytes = (size_t)kIdxSlots * (kIdxW + kIdxPad) * sizeof(uint32_t);
static_assert(kIdxClusterSmemBytes >= kIdxSmemBytes, "builder CTAs use the front of the same allocation for tab/tmp");

__global__ void __cluster_dims__(kIdxCluster, 1, 1) __launch_bounds__(kIdxThreads)
k_snappy_index_cluster(const uint8_t *__restrict__ arena, const DevPage *__restrict__ pages,
                       const int32_t *__restrict__ list, int n_list, uint32_t *__restrict__ frag_pos,
                       uint32_t *__restrict__ page_flag) {
    extern __shared__ __align__(16) uint8_t idx_smem[];
    typedef uint32_t Table[kIdxW + kIdxPad];
    Table *const slots = reinterpret_cast<Table *>(idx_smem);          // CTA 0: kIdxSlots final tables
    Table *const tab = slots;                                          // CTAs 1..: the builders' two working buffers
    Table *const tmp = slots + kIdxBuilders;
    __shared__ __align__(8) uint64_t bar_full[kIdxSlots];              // used in CTA 0 (remote arrivals from the builders)
    __shared__ __align__(8) uint64_t bar_empty[kIdxBuilders];          // used in CTAs 1.. (remote arrivals from the walker)
    __shared__ uint32_t lut[256];
    const uint32_t rank = cluster_ctarank();
    const int li = (int)(blockIdx.x / kIdxCluster);                    // the grid is exactly n_list clusters
    const int pi = list[min(li, n_list - 1)];
    const DevPage pg = pages[pi];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const uint8_t *src = arena + pg.src_off;
    uint32_t src_n = (uint32_t)pg.comp_size;
    uint32_t dst_n = (uint32_t)pg.uncomp_size;
    if (pg.kind == PK_DATA_V2) {
        const uint32_t lv = (uint32_t)(pg.def_bytes + pg.rep_bytes);
        src += lv; src_n -= lv; dst_n -= lv;
    }
    const uint8_t *gin = src - ((uintptr_t)src & 15);
    const uint32_t in_begin = (uint32_t)((uintptr_t)src & 15);
    const uint32_t in_end = in_begin + src_n;
    const uint32_t in_end16 = (in_end + 15) & ~15u;

    for (int t = threadIdx.x; t < 256; t += kIdxThreads) lut[t] = idx_lut_entry((uint32_t)t);
    // the zero entries behind every window (the builders never write them)
    const int n_tables = rank == 0 ? kIdxSlots : 2 * kIdxBuilders;
    for (int t = threadIdx.x; t < n_tables * kIdxPad; t += kIdxThreads) slots[t / kIdxPad][kIdxW + t % kIdxPad] = 0;
    if (threadIdx.x == 0) {
        // full: every lane of the builder arrives for its own stores (a release by one lane would not cover the others')
        for (int b = 0; b < kIdxSlots; b++) idx_mbar_init((uint32_t)__cvta_generic_to_shared(&bar_full[b]), 32);
        for (int b = 0; b < kIdxBuilders; b++) idx_mbar_init((uint32_t)__cvta_generic_to_shared(&bar_empty[b]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    cluster_sync_all();       // barriers and pads of every CTA are in place before anybody reaches across
    const uint32_t nwin = (in_end + kIdxW - 1) / kIdxW;

    if (rank != 0 && warp < kIdxBuilders) {
        // ============================================ builders (CTAs 1..3) =======================================
        const uint32_t g = (rank - 1u) * (uint32_t)kIdxBuilders + (uint32_t)warp;        // slot in CTA 0
        const uint32_t tab_s = shared_addr(&tab[warp][0]), tmp_s = shared_addr(&tmp[warp][0]);
        const uint32_t lut_s = shared_addr(&lut[0]);
        const uint32_t slot_c = cluster_map(shared_addr(&slots[g][0]), 0);               // my slot, in CTA 0
        const uint32_t full_c = cluster_map((uint32_t)__cvta_generic_to_shared(&bar_full[g]), 0);
        const uint32_t my_empty = (uint32_t)__cvta_generic_to_shared(&bar_empty[warp]);
        uint32_t use = 0;
        for (uint32_t j = g; j < nwin; j += kIdxSlots, use++) {
            if (use) cluster_mbar_wait(my_empty, (use - 1) & 1u);      // the walker is done with the slot's previous window
            idx_build_window<true>(tab_s, tmp_s, lut_s, gin, j * kIdxW, in_end, in_end16, slot_c, lane);
            cluster_mbar_arrive(full_c);                               // all 32 lanes, each releasing its own remote stores
        }
    } else if (rank == 0 && warp == kIdxBuilders) {
        // ============================================ walker (CTA 0) =============================================
        const uint32_t lut_s = shared_addr(&lut[0]);
        IdxWalk w{in_begin, 0u, (uint32_t)kSnappyFragment, 1u, 0u};
        uint32_t *const my_pos = frag_pos + pg.frag_first;
        if (lane == 0) my_pos[0] = 0;
        idx_walk_preamble(w, gin, in_end, dst_n);
        const uint32_t slots_s0 = shared_addr(&slots[0][0]), full_s0 = (uint32_t)__cvta_generic_to_shared(&bar_full[0]);
        const uint32_t empty_s0 = (uint32_t)__cvta_generic_to_shared(&bar_empty[0]);
        uint32_t g = 0, gb = 0, grank = 1, phase = 0;      // slot, its builder's warp and CTA, phase of the slot's barrier
        for (uint32_t j = 0; j < nwin; j++) {
            cluster_mbar_poll(full_s0 + 8u * g, phase);
            const uint32_t wend = min((j + 1) * (uint32_t)kIdxW, in_end);
            if (!w.flag && w.ip < wend)
                idx_walk_window(w, slots_s0 + g * (uint32_t)sizeof(Table) - ((j * (uint32_t)kIdxW) << 2), wend, lut_s, gin,
                                in_begin, in_end, my_pos, (uint32_t)pg.nfrag, lane);
            __syncwarp();
            // hand the slot back to its builder (relaxed: the walk above has consumed every value it loaded from the slot,
            // and a release here would put a GPU-scope membar on the walker's chain)
            if (lane == 0 && j + kIdxSlots < nwin) cluster_mbar_arrive_relaxed(cluster_map(empty_s0 + 8u * gb, grank));
            if (++gb == (uint32_t)kIdxBuilders) { gb = 0; grank++; }
            if (++g == (uint32_t)kIdxSlots) { g = 0; gb = 0; grank = 1; phase ^= 1u; }
        }
        if (lane == 0) {
            uint32_t flag = w.flag;
            if (!flag && (w.ip != in_end || w.op != dst_n || w.k != (uint32_t)pg.nfrag)) flag = 1;
            if (!flag) my_pos[pg.nfrag] = src_n;
            page_flag[pg.multi_slot] = flag;
        }
    }
    cluster_sync_all();       // nobody leaves while its shared memory may still be written or read by a peer
}

// ---------------------------------------------------------------------------------------------------------------
// K4 (BYTE_ARRAY dictionaries): {offset, len} of every entry of a PLAIN dictionary page.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_ba_dict_index(uint8_t *__restrict__ arena, const DevPage *__restrict__ pages,
                                const DevCol *__restrict__ cols, const int32_t *__restrict__ list, int n_list,
                                int32_t *status) {
    int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_list) return;
    int pi = list[li];
    const DevPage pg = pages[pi];
    const DevCol col = cols[pg.col];
    const uint8_t *img = arena + pg.img_off;
    BaDictEntry *idx = reinterpret_cast<BaDictEntry *>(arena + col.dict_index_off);
    int64_t pos = 0;
    for (int i = 0; i < col.dict_count; i++) {
        if (pos + 4 > pg.uncomp_size) { report_error(status, DE_BYTE_ARRAY_CORRUPT, pi, i); return; }
        uint32_t len = ld_u32_chain(img + pos);
        pos += 4;
        if (pos + len > (int64_t)pg.uncomp_size) { report_error(status, DE_BYTE_ARRAY_CORRUPT, pi, i); return; }
        idx[i].off = pg.img_off + pos;
        idx[i].len = (int32_t)len;
        idx[i].pad = 0;
        pos += len;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// RLE / bit-packed hybrid stream reader (parquet Encodings.md "RLE/Bit-Packing Hybrid").
// Thread 0 walks the run headers (inherently serial) and writes a run table into shared memory; then all threads of
// the CTA expand table entries in parallel (bit extraction for bit-packed runs).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDecThreads = 256;
constexpr int kTile = 1024;       // values per tile (4 x 256 threads); sizes the shared-memory staging (~32 KB per CTA)
constexpr int kRunCap = 256;      // run-table entries per scan round

struct HybridCursor {
    const uint8_t *p;       // next unread header byte
    const uint8_t *end;
    const uint8_t *bp_ptr;  // current bit-packed run: first byte of the run
    uint32_t bp_index;      // values already consumed inside the current bit-packed run
    uint32_t remaining;     // values left in the current run
    uint32_t rle_value;
    int32_t bw;
    int32_t kind;           // 0 none, 1 RLE, 2 bit-packed
    int32_t error;
};

struct RunEntry {
    const uint8_t *ptr;     // bit-packed: run data; RLE: unused
    uint32_t out_start;     // first output index (relative to this fill)
    uint32_t count;
    uint32_t value_or_index;  // RLE: the value; bit-packed: index of the first value inside the run
    uint32_t is_rle;
};

struct RunTable {
    RunEntry e[kRunCap];
    int32_t n;
    uint32_t filled;        // values covered by e[0..n)
};

__device__ __forceinline__ uint32_t extract_bits(const uint8_t *base, uint64_t bitoff, int bw) {
    const uint8_t *addr = base + (bitoff >> 3);
    uint32_t mis = (uint32_t)((uintptr_t)addr & 3);
    const uint32_t *a = reinterpret_cast<const uint32_t *>(addr - mis);
    uint32_t s = mis * 8 + (uint32_t)(bitoff & 7);  // <= 31
    uint32_t w0 = a[0];
    uint32_t w1 = (s + bw > 32) ? a[1] : 0u;
    uint64_t v = (((uint64_t)w1 << 32) | w0) >> s;
    uint32_t mask = bw >= 32 ? 0xffffffffu : ((1u << bw) - 1u);
    return (uint32_t)v & mask;
}

// thread 0 only: append run entries until `want` values are covered or the table is full
__device__ void hybrid_scan(HybridCursor &c, RunTable &t, uint32_t want) {
    int n = 0;
    uint32_t filled = 0;
    while (filled < want && n < kRunCap) {
        if (c.remaining == 0) {
            // next run header (ULEB128)
            uint32_t h = 0;
            int shift = 0;
            for (;;) {
                if (c.p >= c.end || shift > 28) { c.error = 1; t.n = n; t.filled = filled; return; }
                uint8_t b = *c.p++;
                h |= (uint32_t)(b & 0x7f) << shift;
                if (!(b & 0x80)) break;
                shift += 7;
            }
            if (h & 1) {
                uint32_t groups = h >> 1;
                c.kind = 2;
                c.remaining = groups * 8;
                c.bp_ptr = c.p;
                c.bp_index = 0;
                c.p += (uint64_t)groups * c.bw;
                // a writer may pad the final group past the end of the section; clamp happens via `want`
            } else {
                c.kind = 1;
                c.remaining = h >> 1;
                int nb = (c.bw + 7) >> 3;
                uint32_t v = 0;
                for (int i = 0; i < nb; i++) {
                    if (c.p >= c.end) { c.error = 1; t.n = n; t.filled = filled; return; }
                    v |= (uint32_t)(*c.p++) << (8 * i);
                }
                c.rle_value = v;
            }
            if (c.remaining == 0) continue;  // empty run (legal, useless)
        }
        uint32_t take = min(c.remaining, want - filled);
        RunEntry &e = t.e[n++];
        e.out_start = filled;
        e.count = take;
        if (c.kind == 1) {
            e.is_rle = 1;
            e.value_or_index = c.rle_value;
            e.ptr = nullptr;
        } else {
            e.is_rle = 0;
            e.value_or_index = c.bp_index;
            e.ptr = c.bp_ptr;
            c.bp_index += take;
        }
        c.remaining -= take;
        filled += take;
    }
    t.n = n;
    t.filled = filled;
}

// Warp version of hybrid_scan for long index streams (dictionary fast path of k_decode_pages).  Walking the headers is
// a chain of dependent loads from HBM/L2 - one per run, ~0.2 us each - and the writers emit long sequences of IDENTICAL
// bit-packed runs (Arrow: 64 groups = 512 values per header byte).  So the warp speculates: once a bit-packed header is
// known, lane l reads the header that would follow l runs of the same shape; the leading lanes that find the same header
// again are runs whose position is thereby proven, and all of them enter the table at the price of two loads.  Anything
// else (RLE runs, the odd last run) is decoded one header at a time as before.  All lanes run the same control flow on the
// same cursor values; lane 0 writes the cursor back.
__device__ __forceinline__ bool hybrid_header(const uint8_t *q, const uint8_t *end, uint32_t &h, uint32_t &hlen) {
    h = 0;
    hlen = 0;
    int shift = 0;
    for (;;) {
        if (q + hlen >= end || shift > 28) return false;
        const uint8_t b = q[hlen++];
        h |= (uint32_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return true;
        shift += 7;
    }
}
__device__ void hybrid_scan_warp(HybridCursor &c, RunTable &t, uint32_t want, int lane) {
    const uint8_t *p = c.p, *const end = c.end, *bp_ptr = c.bp_ptr;
    uint32_t bp_index = c.bp_index, remaining = c.remaining, rle_value = c.rle_value;
    const int bw = c.bw;
    int kind = c.kind, n = 0, error = 0;
    uint32_t filled = 0;
    while (filled < want && n < kRunCap) {
        if (remaining == 0) {
            uint32_t h0, hlen0;
            if (!hybrid_header(p, end, h0, hlen0)) { error = 1; break; }
            const uint32_t count0 = min(h0 >> 1, 1u << 27) * 8u;      // (clamped: a damaged header must not wrap the arithmetic)
            if ((h0 & 1u) && count0 != 0) {
                // ---- a sequence of identical bit-packed runs
                const uint64_t stride = (uint64_t)hlen0 + (uint64_t)(h0 >> 1) * (uint64_t)bw;
                const uint8_t *q = p + (uint64_t)lane * stride;
                uint32_t h = 0, hlen = 0;
                const bool same = lane == 0 || (q < end && hybrid_header(q, end, h, hlen) && h == h0 && hlen == hlen0);
                const uint32_t miss = ~__ballot_sync(0xffffffffu, same);
                uint32_t k = miss ? (uint32_t)(__ffs(miss) - 1) : 32u;                  // proven runs
                k = min(k, (uint32_t)(kRunCap - n));
                k = min(k, (want - filled + count0 - 1) / count0);                       // runs needed
                if ((uint32_t)lane < k) {
                    RunEntry &e = t.e[n + lane];
                    e.ptr = q + hlen0;
                    e.out_start = filled + (uint32_t)lane * count0;
                    e.count = min(count0, want - e.out_start);
                    e.value_or_index = 0;
                    e.is_rle = 0;
                }
                const uint32_t got = min(k * count0, want - filled);
                const uint32_t last_take = got - (k - 1) * count0;                       // values taken from the last run
                n += (int)k;
                filled += got;
                p += (uint64_t)k * stride;
                kind = 2;
                bp_ptr = p - (stride - hlen0);
                bp_index = last_take;
                remaining = count0 - last_take;
                continue;
            }
            p += hlen0;
            if (h0 & 1u) continue;                       // empty bit-packed run (legal, useless)
            kind = 1;
            remaining = h0 >> 1;
            const int nb = (bw + 7) >> 3;
            uint32_t v = 0;
            if (p + nb > end) { error = 1; break; }
            for (int i = 0; i < nb; i++) v |= (uint32_t)p[i] << (8 * i);
            p += nb;
            rle_value = v;
            if (remaining == 0) continue;
        }
        const uint32_t take = min(remaining, want - filled);
        if (lane == 0) {
            RunEntry &e = t.e[n];
            e.out_start = filled;
            e.count = take;
            e.is_rle = kind == 1 ? 1u : 0u;
            e.value_or_index = kind == 1 ? rle_value : bp_index;
            e.ptr = kind == 1 ? nullptr : bp_ptr;
        }
        n++;
        if (kind != 1) bp_index += take;
        remaining -= take;
        filled += take;
    }
    __syncwarp();
    if (lane == 0) {
        c.p = p; c.bp_ptr = bp_ptr; c.bp_index = bp_index; c.remaining = remaining; c.rle_value = rle_value;
        c.kind = kind;
        if (error) c.error = 1;
        t.n = n;
        t.filled = filled;
    }
}

// All threads: produce exactly `want` values of the stream into dst[0..want) (shared memory, uint32).
// Returns false (uniformly) on a corrupt stream.
__device__ bool hybrid_fill(HybridCursor &c, RunTable &t, uint32_t *dst, uint32_t want) {
    uint32_t done = 0;
    while (done < want) {
        __syncthreads();
        if (threadIdx.x == 0) hybrid_scan(c, t, want - done);
        __syncthreads();
        if (c.error || t.filled == 0) return false;
        const int n = t.n;
        const uint32_t filled = t.filled;
        const int bw = c.bw;
        int ei = 0;
        for (uint32_t i = threadIdx.x; i < filled; i += kDecThreads) {
            while (ei + 1 < n && t.e[ei + 1].out_start <= i) ei++;
            const RunEntry &e = t.e[ei];
            uint32_t v;
            if (e.is_rle) v = e.value_or_index;
            else v = extract_bits(e.ptr, (uint64_t)(e.value_or_index + (i - e.out_start)) * bw, bw);
            dst[done + i] = v;
        }
        done += filled;
    }
    __syncthreads();
    return true;
}

__device__ __forceinline__ void cursor_init(HybridCursor &c, const uint8_t *p, const uint8_t *end, int bw) {
    c.p = p; c.end = end; c.bp_ptr = p; c.bp_index = 0; c.remaining = 0; c.rle_value = 0; c.bw = bw; c.kind = 0; c.error = 0;
}

__device__ __forceinline__ int bits_for(int max_level) {
    int b = 0;
    while ((1 << b) <= max_level) b++;
    return max_level == 0 ? 0 : b;
}

// block-wide exclusive scan of one 32-bit value per (thread, item) laid out item-major: index = k*kDecThreads + tid
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

struct DecodeShared {
    HybridCursor def_c, rep_c, idx_c;
    RunTable table;
    uint32_t s_def[kTile];
    uint32_t s_rep[kTile];
    uint32_t s_idx[kTile];      // dictionary indices / boolean RLE values of the valid entries of the tile
    uint32_t s_rank[kTile];     // exclusive rank among valid entries (tile-relative)
    int64_t s_ba_off[kTile];    // PLAIN BYTE_ARRAY: arena offset of each valid value
    int32_t s_ba_len[kTile];
    uint32_t warp_sums[kDecThreads / 32];
    const uint8_t *val_ptr;
    const uint8_t *val_end;
    int64_t ba_pos;             // PLAIN BYTE_ARRAY cursor (bytes into the value section)
    uint32_t tile_valid;
    int32_t all_valid;
    int32_t fail;
};

template <int W>
__device__ __forceinline__ void store_fixed(uint8_t *dst, const uint8_t *src, bool aligned) {
    if (W == 4) {
        uint32_t v = aligned ? *reinterpret_cast<const uint32_t *>(src) : ld_u32_unaligned(src);
        *reinterpret_cast<uint32_t *>(dst) = v;
    } else if (W == 8) {
        uint64_t v;
        if (aligned) v = *reinterpret_cast<const uint64_t *>(src);
        else v = (uint64_t)ld_u32_unaligned(src) | ((uint64_t)ld_u32_unaligned(src + 4) << 32);
        *reinterpret_cast<uint64_t *>(dst) = v;
    }
}

__device__ __forceinline__ void copy_small(uint8_t *dst, const uint8_t *src, int w) {
    for (int i = 0; i < w; i++) dst[i] = src[i];
}

__global__ void __launch_bounds__(kDecThreads)
k_decode_pages(uint8_t *__restrict__ arena, uint8_t *__restrict__ out, const DevCol *__restrict__ cols,
               const DevPage *__restrict__ pages, const int32_t *__restrict__ list, int n_list, int32_t *status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    DecodeShared &sh = *reinterpret_cast<DecodeShared *>(smem_raw);

    int li = blockIdx.x;
    if (li >= n_list) return;
    const int pi = list[li];
    const DevPage pg = pages[pi];
    const DevCol col = cols[pg.col];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const uint8_t *img = arena + pg.img_off;
    const uint8_t *img_end = img + pg.uncomp_size;
    const int W = col.width;
    const uint32_t nvals = (uint32_t)pg.num_values;
    const int64_t first = pg.first_value;

    // ---- sections (thread 0), cursors
    if (tid == 0) {
        sh.fail = 0;
        sh.all_valid = 0;
        const uint8_t *p = img;
        const uint8_t *rep_p = nullptr, *rep_e = nullptr, *def_p = nullptr, *def_e = nullptr;
        int rep_bw = bits_for(col.max_rep), def_bw = bits_for(col.max_def);
        if (pg.kind == PK_DATA_V2) {
            rep_p = p; rep_e = p + pg.rep_bytes; p = rep_e;
            def_p = p; def_e = p + pg.def_bytes; p = def_e;
        } else {
            if (col.max_rep > 0) {
                if (pg.rep_enc == ENC_RLE) {
                    if (p + 4 > img_end) sh.fail = 1;
                    else { uint32_t l = ld_u32_unaligned(p); rep_p = p + 4; rep_e = rep_p + l; p = rep_e; }
                } else sh.fail = 2;  // BIT_PACKED levels with max level > 0: deprecated, never written by parquet-mr>=1.x/arrow
            }
            if (col.max_def > 0 && !sh.fail) {
                if (pg.def_enc == ENC_RLE) {
                    if (p + 4 > img_end) sh.fail = 1;
                    else { uint32_t l = ld_u32_unaligned(p); def_p = p + 4; def_e = def_p + l; p = def_e; }
                } else sh.fail = 2;
            }
        }
        if (p > img_end) sh.fail = 1;
        cursor_init(sh.rep_c, rep_p, rep_e, rep_bw);
        cursor_init(sh.def_c, def_p, def_e, def_bw);
        sh.val_ptr = p;
        sh.val_end = img_end;
        sh.ba_pos = 0;
        // all-valid shortcut: the whole page is one RLE run of max_def
        if (!sh.fail && col.max_def > 0 && def_p && def_e > def_p) {
            const uint8_t *q = def_p;
            uint32_t h = 0; int shift = 0; bool ok = true;
            for (;;) {
                if (q >= def_e || shift > 28) { ok = false; break; }
                uint8_t b = *q++;
                h |= (uint32_t)(b & 0x7f) << shift;
                if (!(b & 0x80)) break;
                shift += 7;
            }
            if (ok && !(h & 1) && (h >> 1) >= nvals) {
                int nb = (def_bw + 7) >> 3;
                uint32_t v = 0;
                for (int i = 0; i < nb && q < def_e; i++) v |= (uint32_t)(*q++) << (8 * i);
                if ((int)v == col.max_def) sh.all_valid = 1;
            }
        }
        if (col.max_def == 0) sh.all_valid = 1;
        // dictionary / boolean-RLE index stream
        bool dict_enc = pg.encoding == ENC_PLAIN_DICTIONARY || pg.encoding == ENC_RLE_DICTIONARY;
        if (!sh.fail && dict_enc) {
            if (p >= img_end) { if (nvals) cursor_init(sh.idx_c, p, p, 0); }
            else { int bw = p[0]; if (bw > 32) sh.fail = 1; cursor_init(sh.idx_c, p + 1, img_end, bw); }
            if (col.dict_img_off < 0) sh.fail = 3;
        } else if (!sh.fail && pg.encoding == ENC_RLE) {  // boolean RLE: 4-byte length prefix, bit width 1
            if (p + 4 > img_end) sh.fail = 1;
            else { uint32_t l = ld_u32_unaligned(p); const uint8_t *e2 = p + 4 + l; cursor_init(sh.idx_c, p + 4, e2 < img_end ? e2 : img_end, 1); }
        }
    }
    __syncthreads();
    if (sh.fail) {
        if (tid == 0) report_error(status, sh.fail == 2 ? DE_UNSUPPORTED_ENCODING : (sh.fail == 3 ? DE_DICT_INDEX_RANGE : DE_LEVELS_CORRUPT), pi, sh.fail);
        return;
    }

    uint8_t *o_values = out + col.values_off;
    uint8_t *o_valid = col.valid_off >= 0 ? out + col.valid_off : nullptr;
    uint8_t *o_rep = col.rep_off >= 0 ? out + col.rep_off : nullptr;
    uint8_t *o_def = col.def_off >= 0 ? out + col.def_off : nullptr;
    const uint8_t *val_ptr = sh.val_ptr;
    const bool all_valid = sh.all_valid != 0;
    const bool dict_enc = pg.encoding == ENC_PLAIN_DICTIONARY || pg.encoding == ENC_RLE_DICTIONARY;

    // ---- fast path: PLAIN fixed-width, no nulls, flat  ->  one vectorised copy
    if (pg.encoding == ENC_PLAIN && all_valid && col.max_rep == 0 && col.ptype != PST_BOOLEAN_T && W > 0) {
        int64_t nbytes = (int64_t)nvals * W;
        if (val_ptr + nbytes > img_end) { if (tid == 0) report_error(status, DE_PAGE_OVERRUN, pi, 1); return; }
        coop_copy(o_values + first * W, val_ptr, nbytes, tid, kDecThreads);
        if (o_valid) coop_fill(o_valid + first, 1, nvals, tid, kDecThreads);
        return;
    }

    const uint8_t *dict = col.dict_img_off >= 0 ? arena + col.dict_img_off : nullptr;
    const BaDictEntry *ba_dict = col.dict_index_off >= 0 ? reinterpret_cast<const BaDictEntry *>(arena + col.dict_index_off) : nullptr;

    // ---- fast path: dictionary-encoded 4/8-byte values of a flat page without nulls (what pyarrow writes for the first
    // ~1 MiB of distinct values of every column before it falls back to PLAIN).  Thread 0 scans up to 256 run headers of
    // the index stream at once (a 20,000-value page is ~40 bit-packed runs), then every thread extracts its indices
    // straight from the page image and gathers from the dictionary: two block barriers per page instead of ~10 per
    // 1024-value tile, no staging of indices or ranks through shared memory.
    if (dict_enc && all_valid && col.max_rep == 0 && (W == 4 || W == 8) && dict != nullptr) {
        uint32_t done = 0;
        const uint32_t dict_count = (uint32_t)col.dict_count;
        while (done < nvals) {
            __syncthreads();
            if (warp == 0) hybrid_scan_warp(sh.idx_c, sh.table, nvals - done, lane);
            __syncthreads();
            if (sh.idx_c.error || sh.table.filled == 0) {
                if (tid == 0) report_error(status, DE_LEVELS_CORRUPT, pi, 12);
                return;
            }
            const int n = sh.table.n;
            const uint32_t filled = sh.table.filled;
            const int bw = sh.idx_c.bw;
            // four values per thread and round: the index words come from HBM and the dictionary entries from L2/HBM, and
            // with one value in flight per thread 74 % of the warp samples were waits on these two loads (r2t capture)
            int ei = 0;
            bool range_err = false;
            for (uint32_t i0 = tid; i0 < filled; i0 += 4 * kDecThreads) {
                uint32_t di[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)u * kDecThreads;
                    di[u] = 0;
                    if (i < filled) {
                        while (ei + 1 < n && sh.table.e[ei + 1].out_start <= i) ei++;
                        const RunEntry &e = sh.table.e[ei];
                        di[u] = e.is_rle ? e.value_or_index
                                         : extract_bits(e.ptr, (uint64_t)(e.value_or_index + (i - e.out_start)) * bw, bw);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (di[u] >= dict_count) { range_err = true; di[u] = 0; }
                const int64_t row = first + done + i0;
                if (W == 4) {
                    uint32_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = reinterpret_cast<const uint32_t *>(dict)[di[u]];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (i0 + (uint32_t)u * kDecThreads < filled)
                            reinterpret_cast<uint32_t *>(o_values)[row + u * kDecThreads] = v[u];
                } else {
                    uint64_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = reinterpret_cast<const uint64_t *>(dict)[di[u]];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (i0 + (uint32_t)u * kDecThreads < filled)
                            reinterpret_cast<uint64_t *>(o_values)[row + u * kDecThreads] = v[u];
                }
            }
            if (range_err) report_error(status, DE_DICT_INDEX_RANGE, pi, 0);
            done += filled;
        }
        if (o_valid) coop_fill(o_valid + first, 1, nvals, tid, kDecThreads);
        return;
    }
    const bool val_aligned = W > 0 && (((uintptr_t)val_ptr) % (W == 8 ? 8 : 4)) == 0;

    uint32_t rank_base = 0;  // valid values consumed before this tile
    for (uint32_t base = 0; base < nvals; base += kTile) {
        const uint32_t tn = min((uint32_t)kTile, nvals - base);
        // levels of the tile
        if (col.max_rep > 0) {
            if (!hybrid_fill(sh.rep_c, sh.table, sh.s_rep, tn)) { if (tid == 0) report_error(status, DE_LEVELS_CORRUPT, pi, 10); return; }
        }
        if (!all_valid) {
            if (!hybrid_fill(sh.def_c, sh.table, sh.s_def, tn)) { if (tid == 0) report_error(status, DE_LEVELS_CORRUPT, pi, 11); return; }
        }
        // ranks: exclusive scan of validity, tile laid out item-major (i = k*256 + tid); without nulls the rank of an
        // entry is its position and the scan (two block barriers per 256 entries) is skipped
        uint32_t carry = all_valid ? tn : 0;
        for (uint32_t k = 0; !all_valid && k < tn; k += kDecThreads) {
            uint32_t i = k + tid;
            uint32_t v = (i < tn) ? (all_valid ? 1u : (sh.s_def[i] == (uint32_t)col.max_def ? 1u : 0u)) : 0u;
            uint32_t incl = warp_incl_scan(v, lane);
            if (lane == 31) sh.warp_sums[warp] = incl;
            __syncthreads();
            uint32_t woff = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kDecThreads / 32; w++) {
                uint32_t s = sh.warp_sums[w];
                if (w < warp) woff += s;
                total += s;
            }
            if (i < tn) sh.s_rank[i] = carry + woff + incl - v;
            carry += total;
            __syncthreads();
        }
        const uint32_t tile_valid = carry;

        // value-side staging for the valid entries of the tile
        if (dict_enc || pg.encoding == ENC_RLE) {
            if (tile_valid) {
                if (!hybrid_fill(sh.idx_c, sh.table, sh.s_idx, tile_valid)) { if (tid == 0) report_error(status, DE_LEVELS_CORRUPT, pi, 12); return; }
            }
        } else if (col.ptype == PST_BYTE_ARRAY_T) {
            // PLAIN BYTE_ARRAY: length-prefixed values; positions are a serial chain
            if (tid == 0) {
                int64_t pos = sh.ba_pos;
                const int64_t lim = sh.val_end - val_ptr;
                for (uint32_t r = 0; r < tile_valid; r++) {
                    if (pos + 4 > lim) { sh.fail = 1; break; }
                    uint32_t len = ld_u32_chain(val_ptr + pos);
                    pos += 4;
                    if (pos + (int64_t)len > lim) { sh.fail = 1; break; }
                    sh.s_ba_off[r] = (val_ptr - arena) + pos;
                    sh.s_ba_len[r] = (int32_t)len;
                    pos += len;
                }
                sh.ba_pos = pos;
            }
            __syncthreads();
            if (sh.fail) { if (tid == 0) report_error(status, DE_BYTE_ARRAY_CORRUPT, pi, 13); return; }
        }

        // emit
        for (uint32_t i = tid; i < tn; i += kDecThreads) {
            const int64_t row = first + base + i;
            const bool valid = all_valid || sh.s_def[i] == (uint32_t)col.max_def;
            const uint32_t r = all_valid ? i : sh.s_rank[i];   // tile-relative rank
            const uint64_t gr = (uint64_t)rank_base + r;  // page-relative rank
            if (o_valid) o_valid[row] = valid ? 1 : 0;
            if (o_rep) { o_rep[row] = (uint8_t)sh.s_rep[i]; o_def[row] = (uint8_t)(all_valid ? col.max_def : sh.s_def[i]); }
            if (col.ptype == PST_BYTE_ARRAY_T) {
                int64_t off = 0; int32_t len = 0;
                if (valid) {
                    if (dict_enc) {
                        uint32_t di = sh.s_idx[r];
                        if (di >= (uint32_t)col.dict_count) { report_error(status, DE_DICT_INDEX_RANGE, pi, (int)di); }
                        else { off = ba_dict[di].off; len = ba_dict[di].len; }
                    } else { off = sh.s_ba_off[r]; len = sh.s_ba_len[r]; }
                }
                reinterpret_cast<int64_t *>(o_values)[row] = off;
                reinterpret_cast<int32_t *>(out + col.lens_off)[row] = len;
            } else if (col.ptype == PST_BOOLEAN_T) {
                uint8_t v = 0;
                if (valid) {
                    if (pg.encoding == ENC_RLE) v = (uint8_t)(sh.s_idx[r] & 1);
                    else {
                        const uint8_t *bp = val_ptr + (gr >> 3);
                        if (bp >= img_end) { report_error(status, DE_PAGE_OVERRUN, pi, 2); }
                        else v = (*bp >> (gr & 7)) & 1;
                    }
                }
                o_values[row] = v;
            } else {
                uint8_t *d = o_values + row * W;
                if (!valid) {
                    if (W == 4) *reinterpret_cast<uint32_t *>(d) = 0;
                    else if (W == 8) *reinterpret_cast<uint64_t *>(d) = 0;
                    else for (int b = 0; b < W; b++) d[b] = 0;
                } else {
                    const uint8_t *s;
                    bool al;
                    if (dict_enc) {
                        uint32_t di = sh.s_idx[r];
                        if (di >= (uint32_t)col.dict_count) { report_error(status, DE_DICT_INDEX_RANGE, pi, (int)di); di = 0; }
                        s = dict + (uint64_t)di * W;
                        al = true;  // dictionary images are placed 16B aligned
                    } else {
                        s = val_ptr + gr * W;
                        al = val_aligned;
                        if (s + W > img_end) { report_error(status, DE_PAGE_OVERRUN, pi, 3); continue; }
                    }
                    if (W == 4) store_fixed<4>(d, s, al);
                    else if (W == 8) store_fixed<8>(d, s, al);
                    else copy_small(d, s, W);
                }
            }
        }
        rank_base += tile_valid;
        __syncthreads();
    }
    // per-column count of level entries below max_def (nulls), read by the host together with the error word
    if (tid == 0 && nvals > rank_base) atomicAdd(&status[8 + pg.col], (int32_t)(nvals - rank_base));
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
cudaError_t configure_decode_kernels() {
    // the Snappy kernels are latency-bound warps: as many CTAs per SM as their shared memory allows
    cudaError_t e = cudaFuncSetAttribute(k_snappy_pages, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_snappy_index, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_snappy_index, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIdxSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_snappy_index_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIdxClusterSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_decode_pages, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeShared));
}

cudaError_t launch_snappy_index(uint8_t *arena, const DevPage *pages, const int32_t *multi_list, int n_multi,
                                uint32_t *frag_pos, uint32_t *page_flag, cudaStream_t s) {
    if (n_multi <= 0) return cudaSuccess;
    k_snappy_index<<<n_multi, kIdxThreads, kIdxSmemBytes, s>>>(arena, pages, multi_list, n_multi, frag_pos, page_flag);
    return cudaGetLastError();
}
cudaError_t launch_snappy_index_cluster(uint8_t *arena, const DevPage *pages, const int32_t *list, int n_list,
                                        uint32_t *frag_pos, uint32_t *page_flag, cudaStream_t s) {
    if (n_list <= 0) return cudaSuccess;
    k_snappy_index_cluster<<<n_list * kIdxCluster, kIdxThreads, kIdxClusterSmemBytes, s>>>(arena, pages, list, n_list,
                                                                                           frag_pos, page_flag);
    return cudaGetLastError();
}

cudaError_t launch_snappy(uint8_t *arena, const DevPage *pages, const SnFrag *frags, int n_frags,
                          const int32_t *multi_list, int n_multi, const uint32_t *frag_pos, uint32_t *page_flag,
                          int32_t *status, int serial_mode, cudaStream_t s) {
    const int n = serial_mode ? n_multi : n_frags;
    if (n <= 0) return cudaSuccess;
    k_snappy_pages<<<n, kSnappyThreads, 0, s>>>(arena, pages, frags, n_frags, multi_list, n_multi, frag_pos, page_flag,
                                                 status, serial_mode);
    return cudaGetLastError();
}

cudaError_t launch_ba_dict_index(uint8_t *arena, const DevPage *pages, const DevCol *cols, const int32_t *list, int n,
                                 int32_t *status, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_ba_dict_index<<<(n + 31) / 32, 32, 0, s>>>(arena, pages, cols, list, n, status);
    return cudaGetLastError();
}

cudaError_t launch_decode_pages(uint8_t *arena, uint8_t *out, const DevCol *cols, const DevPage *pages,
                                const int32_t *list, int n, int32_t *status, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    size_t smem = sizeof(DecodeShared);
    k_decode_pages<<<n, kDecThreads, smem, s>>>(arena, out, cols, pages, list, n, status);
    return cudaGetLastError();
}

}  // namespace pst
