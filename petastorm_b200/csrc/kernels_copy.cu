// k_copy_tiles -- PLAIN fixed-width values of flat pages without nulls are a byte range of the page image; moving them
// to their place in the column tensor is a pure copy, and it is the bulk of a row-group (all float/int pages of the C2
// workload after the dictionary fallback).  The planner turns those pages into <= 64 KiB tiles (CopyTile); this kernel
// moves them with the Blackwell bulk-copy engine instead of through registers:
//
//   warp 0   one elected lane drives a ring of kCopyStages x 16 KiB shared-memory slots:
//            cp.async.bulk global -> shared (completion on an mbarrier, complete_tx::bytes), then
//            cp.async.bulk shared -> global (bulk_group), kCopyStages loads in flight per CTA.  SASS: UBLKCP.
//   warp 1-3 the odd jobs of the same tiles: the < 16 byte tail of an aligned tile, tiles whose source or destination is
//            not 16-byte aligned (vector copy with funnel shifts, coop_copy), and the validity bytes of columns that
//            carry a validity array.
//
// Replaces the value copy Arrow C++ does inside piece.read for PLAIN pages (petastorm/arrow_reader_worker.py:358,
// petastorm/py_dict_reader_worker.py:267).  Algorithmic bytes: nbytes read + nbytes written per tile (+ nvalid).
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"
#include "dev_util.cuh"
#include "kernels.h"

namespace pst {

constexpr int kCopyThreads = 128;
constexpr int kCopyStages = 4;
constexpr uint32_t kCopyChunk = 16384;
constexpr int kCopyCtasPerSm = 3;      // 3 x 64 KiB of staging per SM

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// A tile takes the bulk path when source and destination are 16-byte aligned (the planner phases page images so that
// value sections are; the destination is when the page starts on a multiple of 16 / width values).
__device__ __forceinline__ bool tile_is_bulk(const uint8_t *src, const uint8_t *dst, int32_t nbytes) {
    return ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0 && nbytes >= 16;
}

__global__ void __launch_bounds__(kCopyThreads)
k_copy_tiles(const uint8_t *__restrict__ arena, uint8_t *__restrict__ out, const CopyTile *__restrict__ tiles, int n_tiles) {
    extern __shared__ __align__(128) uint8_t stage_mem[];          // kCopyStages x kCopyChunk
    __shared__ __align__(8) uint64_t full_bar[kCopyStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kCopyStages; s++) mbar_init(smem_u32(&full_bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        if (lane != 0) return;
        // ---- bulk-copy driver.  The chunks of all aligned tiles of this CTA form one sequence; chunk k lives in slot
        // k % kCopyStages.  `head` walks the sequence for loads, `tail` for stores, kCopyStages chunks apart.
        struct Cursor {
            int tile;            // current tile index (strided by gridDim.x)
            uint32_t off;        // bytes of the tile's body already handed out
        };
        auto body_of = [&](int t) -> uint32_t {
            const CopyTile tl = tiles[t];
            return tile_is_bulk(arena + tl.src_off, out + tl.dst_off, tl.nbytes) ? ((uint32_t)tl.nbytes & ~15u) : 0u;
        };
        auto advance = [&](Cursor &c) {       // move to the next chunk; tile == n_tiles (or beyond) when exhausted
            while (c.tile < n_tiles) {
                if (c.off < body_of(c.tile)) return;
                c.tile += gridDim.x;
                c.off = 0;
            }
        };
        Cursor head{(int)blockIdx.x, 0}, tail{(int)blockIdx.x, 0};
        advance(head);
        advance(tail);
        uint32_t k_load = 0, k_store = 0;
        auto issue_load = [&]() {
            const CopyTile tl = tiles[head.tile];
            const uint32_t body = (uint32_t)tl.nbytes & ~15u;
            const uint32_t n = min(kCopyChunk, body - head.off);
            const int s = (int)(k_load % kCopyStages);
            const uint32_t bar = smem_u32(&full_bar[s]);
            mbar_expect_tx(bar, n);
            bulk_g2s(smem_u32(stage_mem + (size_t)s * kCopyChunk), arena + tl.src_off + head.off, n, bar);
            head.off += n;
            k_load++;
            advance(head);
        };
        for (int s = 0; s < kCopyStages && head.tile < n_tiles; s++) issue_load();
        while (tail.tile < n_tiles) {
            const CopyTile tl = tiles[tail.tile];
            const uint32_t body = (uint32_t)tl.nbytes & ~15u;
            const uint32_t n = min(kCopyChunk, body - tail.off);
            const int s = (int)(k_store % kCopyStages);
            mbar_wait(smem_u32(&full_bar[s]), (k_store / kCopyStages) & 1u);
            bulk_s2g(out + tl.dst_off + tail.off, smem_u32(stage_mem + (size_t)s * kCopyChunk), n);
            bulk_commit();
            tail.off += n;
            k_store++;
            advance(tail);
            if (head.tile < n_tiles) {
                bulk_wait_read0();          // the slot just stored from is the one the next load lands in
                issue_load();
            }
        }
        bulk_wait_all();
        return;
    }

    // ---- warps 1..3: tails, unaligned tiles, validity bytes
    const int tid = threadIdx.x - 32, nthr = kCopyThreads - 32;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const CopyTile tl = tiles[t];
        const uint8_t *src = arena + tl.src_off;
        uint8_t *dst = out + tl.dst_off;
        if (tile_is_bulk(src, dst, tl.nbytes)) {
            const int32_t body = tl.nbytes & ~15;
            if (tid < tl.nbytes - body) dst[body + tid] = src[body + tid];
        } else {
            coop_copy(dst, src, tl.nbytes, tid, nthr);
        }
        if (tl.valid_off >= 0) coop_fill(out + tl.valid_off, 1, tl.nvalid, tid, nthr);
    }
}

cudaError_t configure_copy_kernel() {
    return cudaFuncSetAttribute(k_copy_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, kCopyStages * (int)kCopyChunk);
}

cudaError_t launch_copy_tiles(const uint8_t *arena, uint8_t *out, const CopyTile *tiles, int n_tiles, int sm_count,
                              cudaStream_t s) {
    if (n_tiles <= 0) return cudaSuccess;
    const int grid = n_tiles < sm_count * kCopyCtasPerSm ? n_tiles : sm_count * kCopyCtasPerSm;
    k_copy_tiles<<<grid, kCopyThreads, kCopyStages * kCopyChunk, s>>>(arena, out, tiles, n_tiles);
    return cudaGetLastError();
}

}  // namespace pst
