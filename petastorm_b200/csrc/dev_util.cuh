// Device helpers shared by the decode and copy kernels: error word, unaligned loads, cooperative byte copies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pst {

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void report_error(int32_t *status, int code, int page, int detail) {
    if (atomicCAS(status, 0, code) == 0) {
        status[1] = page;
        status[2] = detail;
    }
}

__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Little-endian 32-bit value at any address as two aligned word loads + a funnel shift: ONE memory latency on a serial
// chain (length prefixes of BYTE_ARRAY values) instead of four byte loads.  Reads the aligned words that contain
// [p, p+4): up to 3 bytes in front of p and behind p+3 (the planner leaves that slack around every page image).
__device__ __forceinline__ uint32_t ld_u32_chain(const uint8_t *p) {
    const uint32_t mis = (uint32_t)((uintptr_t)p & 3);
    const uint32_t *a = reinterpret_cast<const uint32_t *>(p - mis);
    const uint32_t w0 = a[0];
    const uint32_t w1 = mis ? a[1] : 0u;
    return __funnelshift_r(w0, w1, mis * 8);
}

// 16 bytes starting `m` bytes into the 32-byte window {a, b}
__device__ __forceinline__ uint4 extract16(uint4 a, uint4 b, uint32_t m) {
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t q = m >> 2, s = (m & 3) * 8;
    uint4 r;
    // q is uniform across the cooperating threads -> no divergence; switch keeps w[] in registers
    switch (q) {
        case 0: r.x = __funnelshift_r(w[0], w[1], s); r.y = __funnelshift_r(w[1], w[2], s);
                r.z = __funnelshift_r(w[2], w[3], s); r.w = __funnelshift_r(w[3], w[4], s); break;
        case 1: r.x = __funnelshift_r(w[1], w[2], s); r.y = __funnelshift_r(w[2], w[3], s);
                r.z = __funnelshift_r(w[3], w[4], s); r.w = __funnelshift_r(w[4], w[5], s); break;
        case 2: r.x = __funnelshift_r(w[2], w[3], s); r.y = __funnelshift_r(w[3], w[4], s);
                r.z = __funnelshift_r(w[4], w[5], s); r.w = __funnelshift_r(w[5], w[6], s); break;
        default: r.x = __funnelshift_r(w[3], w[4], s); r.y = __funnelshift_r(w[4], w[5], s);
                 r.z = __funnelshift_r(w[5], w[6], s); r.w = __funnelshift_r(w[6], w[7], s); break;
    }
    return r;
}

// Cooperative byte copy by `nthr` threads (tid in [0,nthr)); src and dst must not overlap.  Destination-aligned
// 16-byte stores; source read as aligned 16-byte words and funnel-shifted, so any relative alignment runs at
// vector width.  May read up to 15 bytes past src+n and before src (inside the same aligned 16B words) -- the planner
// leaves that slack around every page.
__device__ __forceinline__ void coop_copy(uint8_t *dst, const uint8_t *src, int64_t n, int tid, int nthr) {
    if (n <= 0) return;
    if (n < 64) {
        for (int64_t i = tid; i < n; i += nthr) dst[i] = src[i];
        return;
    }
    int64_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    for (int64_t i = tid; i < head; i += nthr) dst[i] = src[i];
    uint8_t *d = dst + head;
    const uint8_t *s = src + head;
    int64_t body = (n - head) >> 4;
    uint32_t m = (uint32_t)((uintptr_t)s & 15);
    const uint4 *s16 = reinterpret_cast<const uint4 *>(s - m);
    uint4 *d16 = reinterpret_cast<uint4 *>(d);
    if (m == 0) {
        int64_t k = tid;
        for (; k + 3 * (int64_t)nthr < body; k += 4 * (int64_t)nthr) {
            uint4 v0 = s16[k], v1 = s16[k + nthr], v2 = s16[k + 2 * nthr], v3 = s16[k + 3 * nthr];
            d16[k] = v0; d16[k + nthr] = v1; d16[k + 2 * nthr] = v2; d16[k + 3 * nthr] = v3;
        }
        for (; k < body; k += nthr) d16[k] = s16[k];
    } else {
        int64_t k = tid;
        for (; k + (int64_t)nthr < body; k += 2 * (int64_t)nthr) {
            uint4 a0 = s16[k], b0 = s16[k + 1], a1 = s16[k + nthr], b1 = s16[k + nthr + 1];
            d16[k] = extract16(a0, b0, m);
            d16[k + nthr] = extract16(a1, b1, m);
        }
        for (; k < body; k += nthr) d16[k] = extract16(s16[k], s16[k + 1], m);
    }
    int64_t done = head + (body << 4);
    for (int64_t i = done + tid; i < n; i += nthr) dst[i] = src[i];
}

__device__ __forceinline__ void coop_fill(uint8_t *dst, uint8_t v, int64_t n, int tid, int nthr) {
    if (n <= 0) return;
    int64_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head > n) head = n;
    for (int64_t i = tid; i < head; i += nthr) dst[i] = v;
    int64_t body = (n - head) >> 4;
    uint32_t w = 0x01010101u * v;
    uint4 v4 = make_uint4(w, w, w, w);
    uint4 *d16 = reinterpret_cast<uint4 *>(dst + head);
    for (int64_t k = tid; k < body; k += nthr) d16[k] = v4;
    for (int64_t i = head + (body << 4) + tid; i < n; i += nthr) dst[i] = v;
}


}  // namespace pst
