// sm_100a column post-processing kernels: null handling, casts, row gather, NdarrayCodec payload copy, predicates
// (in_set, MD5 pseudorandom split), stream compaction, TransformSpec normalise, NGram windows, dtype sanitise.
// Each kernel names the reference function it replaces; all are HBM-bound element-wise / gather work.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"
#include "kernels.h"

namespace pst {

constexpr int kThreads = 256;
static inline int grid_for(int64_t n, int per_block, int cap = 148 * 16) {
    int64_t g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

__device__ __forceinline__ void report_err(int32_t *status, int code, int a, int b) {
    if (atomicCAS(status, 0, code) == 0) {
        status[1] = a;
        status[2] = b;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K6: nulls.  pandas semantics of `column.to_pandas()` (petastorm/arrow_reader_worker.py:55-57): integer columns with
// nulls become float64 with NaN; float columns keep their type with NaN at nulls.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_int_to_f64_nan(const T *__restrict__ v, const uint8_t *__restrict__ valid, int64_t n,
                                 double *__restrict__ out) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = valid[i] ? (double)v[i] : nan;
}
__global__ void k_f32_nan(const float *__restrict__ v, const uint8_t *__restrict__ valid, int64_t n,
                          float *__restrict__ out) {
    const float nan = __int_as_float(0x7fc00000);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = valid[i] ? v[i] : nan;
}
__global__ void k_f64_nan(const double *__restrict__ v, const uint8_t *__restrict__ valid, int64_t n,
                          double *__restrict__ out) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = valid[i] ? v[i] : nan;
}

cudaError_t launch_nullable_to_f64(const void *values, const uint8_t *valid, int64_t n, int ptype, int bits,
                                   int is_unsigned, double *out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int g = grid_for(n, kThreads);
    if (ptype == PST_INT32_T) {
        // INT32 storage of (u)int8/16/32 logical types: the stored int32 already holds the logical value, except
        // uint32 which is stored as its two's-complement bit pattern
        if (is_unsigned && bits == 32)
            k_int_to_f64_nan<uint32_t><<<g, kThreads, 0, s>>>((const uint32_t *)values, valid, n, out);
        else
            k_int_to_f64_nan<int32_t><<<g, kThreads, 0, s>>>((const int32_t *)values, valid, n, out);
    } else if (ptype == PST_INT64_T) {
        if (is_unsigned)
            k_int_to_f64_nan<uint64_t><<<g, kThreads, 0, s>>>((const uint64_t *)values, valid, n, out);
        else
            k_int_to_f64_nan<int64_t><<<g, kThreads, 0, s>>>((const int64_t *)values, valid, n, out);
    } else if (ptype == PST_FLOAT_T) {
        k_f32_nan<<<g, kThreads, 0, s>>>((const float *)values, valid, n, (float *)out);
    } else if (ptype == PST_DOUBLE_T) {
        k_f64_nan<<<g, kThreads, 0, s>>>((const double *)values, valid, n, out);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// INT32 storage -> narrow logical integer type (parquet stores int8/uint8/int16/uint16 as INT32)
template <typename T>
__global__ void k_narrow(const int32_t *__restrict__ src, int64_t n, T *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (T)src[i];
}
cudaError_t launch_narrow_int32(const int32_t *src, int64_t n, int bits, void *dst, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int g = grid_for(n, kThreads);
    if (bits == 8) k_narrow<uint8_t><<<g, kThreads, 0, s>>>(src, n, (uint8_t *)dst);
    else if (bits == 16) k_narrow<uint16_t><<<g, kThreads, 0, s>>>(src, n, (uint16_t *)dst);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// K13 row gather (also K15 window gather): dst[i, 0:copy_bytes] = src[idx[i]*src_stride : +copy_bytes]
// Small rows: one thread per 4/8/16-byte row; big rows: the row is cut in 32 KiB slices, one CTA per slice.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_gather_small(const T *__restrict__ src, const int64_t *__restrict__ idx, int64_t n, T *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[idx[i]];
}

constexpr int64_t kSlice = 32768;

__device__ __forceinline__ void block_copy_any(uint8_t *d, const uint8_t *s, int64_t n) {
    // 16B vector path when both sides share 16B alignment, else 4B, else bytes
    uintptr_t a = (uintptr_t)d | (uintptr_t)s;
    if ((a & 15) == 0) {
        int64_t nv = n >> 4;
        const uint4 *s4 = (const uint4 *)s;
        uint4 *d4 = (uint4 *)d;
        for (int64_t k = threadIdx.x; k < nv; k += blockDim.x) d4[k] = s4[k];
        for (int64_t k = (nv << 4) + threadIdx.x; k < n; k += blockDim.x) d[k] = s[k];
    } else if ((a & 3) == 0) {
        int64_t nv = n >> 2;
        const uint32_t *s4 = (const uint32_t *)s;
        uint32_t *d4 = (uint32_t *)d;
        for (int64_t k = threadIdx.x; k < nv; k += blockDim.x) d4[k] = s4[k];
        for (int64_t k = (nv << 2) + threadIdx.x; k < n; k += blockDim.x) d[k] = s[k];
    } else {
        for (int64_t k = threadIdx.x; k < n; k += blockDim.x) d[k] = s[k];
    }
}

__global__ void k_gather_big(const uint8_t *__restrict__ src, const int64_t *__restrict__ idx, int64_t n,
                             int64_t src_stride, int64_t copy_bytes, int64_t slices, uint8_t *__restrict__ dst) {
    for (int64_t w = blockIdx.x; w < n * slices; w += gridDim.x) {
        int64_t row = w / slices, sl = w % slices;
        int64_t b0 = sl * kSlice;
        int64_t nb = min(kSlice, copy_bytes - b0);
        const uint8_t *s = src + idx[row] * src_stride + b0;
        uint8_t *d = dst + row * copy_bytes + b0;
        block_copy_any(d, s, nb);
    }
}

static cudaError_t gather_impl(const uint8_t *src, const int64_t *idx, int64_t n, int64_t src_stride, int64_t copy_bytes,
                               uint8_t *dst, cudaStream_t s) {
    if (n <= 0 || copy_bytes <= 0) return cudaSuccess;
    bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    if (src_stride == copy_bytes && aligned && (copy_bytes == 1 || copy_bytes == 2 || copy_bytes == 4 || copy_bytes == 8 || copy_bytes == 16)) {
        int g = grid_for(n, kThreads);
        switch (copy_bytes) {
            case 1: k_gather_small<uint8_t><<<g, kThreads, 0, s>>>((const uint8_t *)src, idx, n, (uint8_t *)dst); break;
            case 2: k_gather_small<uint16_t><<<g, kThreads, 0, s>>>((const uint16_t *)src, idx, n, (uint16_t *)dst); break;
            case 4: k_gather_small<uint32_t><<<g, kThreads, 0, s>>>((const uint32_t *)src, idx, n, (uint32_t *)dst); break;
            case 8: k_gather_small<uint64_t><<<g, kThreads, 0, s>>>((const uint64_t *)src, idx, n, (uint64_t *)dst); break;
            default: k_gather_small<uint4><<<g, kThreads, 0, s>>>((const uint4 *)src, idx, n, (uint4 *)dst); break;
        }
        return cudaGetLastError();
    }
    int64_t slices = (copy_bytes + kSlice - 1) / kSlice;
    int64_t work = n * slices;
    int g = (int)(work < 148 * 32 ? work : 148 * 32);
    int threads = copy_bytes >= 4096 ? 256 : (copy_bytes >= 512 ? 128 : 32);
    k_gather_big<<<g, threads, 0, s>>>(src, idx, n, src_stride, copy_bytes, slices, dst);
    return cudaGetLastError();
}

cudaError_t launch_gather_rows(const uint8_t *src, const int64_t *idx, int64_t n_out, int64_t row_bytes, uint8_t *dst,
                               cudaStream_t s) {
    return gather_impl(src, idx, n_out, row_bytes, row_bytes, dst, s);
}
// NGram windows over a scalar column (rows of 1..16 bytes): one thread per (window, timestep) element.  Consecutive
// threads write consecutive elements of `dst` and read consecutive rows of `src` (windows of neighbouring starts
// overlap), so both sides coalesce; a CTA-per-window copy would move 64 bytes per warp.
template <typename T>
__global__ void k_ngram_small(const T *__restrict__ src, const int64_t *__restrict__ starts, int64_t n_windows,
                              int length, T *__restrict__ dst) {
    const int64_t total = n_windows * (int64_t)length;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = i / length;
        const int t = (int)(i - w * length);
        dst[i] = src[starts[w] + t];
    }
}

cudaError_t launch_ngram_gather(const uint8_t *src, const int64_t *starts, int64_t n_windows, int length,
                                int64_t row_bytes, uint8_t *dst, cudaStream_t s) {
    if (n_windows <= 0 || length <= 0 || row_bytes <= 0) return cudaSuccess;
    const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & (uintptr_t)(row_bytes - 1)) == 0;
    if (aligned && (row_bytes == 1 || row_bytes == 2 || row_bytes == 4 || row_bytes == 8 || row_bytes == 16)) {
        const int g = grid_for(n_windows * (int64_t)length, kThreads);
        switch (row_bytes) {
            case 1: k_ngram_small<uint8_t><<<g, kThreads, 0, s>>>((const uint8_t *)src, starts, n_windows, length, (uint8_t *)dst); break;
            case 2: k_ngram_small<uint16_t><<<g, kThreads, 0, s>>>((const uint16_t *)src, starts, n_windows, length, (uint16_t *)dst); break;
            case 4: k_ngram_small<uint32_t><<<g, kThreads, 0, s>>>((const uint32_t *)src, starts, n_windows, length, (uint32_t *)dst); break;
            case 8: k_ngram_small<uint64_t><<<g, kThreads, 0, s>>>((const uint64_t *)src, starts, n_windows, length, (uint64_t *)dst); break;
            default: k_ngram_small<uint4><<<g, kThreads, 0, s>>>((const uint4 *)src, starts, n_windows, length, (uint4 *)dst); break;
        }
        return cudaGetLastError();
    }
    return gather_impl(src, starts, n_windows, row_bytes, row_bytes * length, dst, s);
}

// ---------------------------------------------------------------------------------------------------------------
// K7 NdarrayCodec: value = .npy blob (magic, version, header dict, payload).  The host parsed the first blob's header
// (np.load semantics, petastorm/codecs.py:155-157); every other blob must carry byte-identical header bytes (same
// dtype/shape/order) -- verified here -- and its payload is copied to dst[i].
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_npy_batch(const uint8_t *__restrict__ base, const int64_t *__restrict__ offs,
                            const int32_t *__restrict__ lens, const int64_t *__restrict__ row_idx, int64_t n,
                            int64_t data_off, int64_t payload, int64_t slices, uint8_t *__restrict__ dst,
                            int32_t *status) {
    const int64_t r0 = row_idx ? row_idx[0] : 0;
    const uint8_t *h0 = base + offs[r0];
    for (int64_t w = blockIdx.x; w < n * slices; w += gridDim.x) {
        int64_t i = w / slices, sl = w % slices;
        int64_t r = row_idx ? row_idx[i] : i;
        const uint8_t *blob = base + offs[r];
        if (sl == 0) {
            // header check (data_off bytes, <= a few hundred) + length check
            int bad = 0;
            if ((int64_t)lens[r] != data_off + payload) bad = 1;
            else
                for (int64_t k = threadIdx.x; k < data_off; k += blockDim.x)
                    if (blob[k] != h0[k]) bad = 1;
            if (bad) report_err(status, DE_NPY_HEADER_MISMATCH, (int)i, lens[r]);
        }
        if ((int64_t)lens[r] < data_off + payload) continue;
        int64_t b0 = sl * kSlice;
        int64_t nb = min(kSlice, payload - b0);
        block_copy_any(dst + i * payload + b0, blob + data_off + b0, nb);
    }
}

cudaError_t launch_npy_batch(const uint8_t *base, const int64_t *offs, const int32_t *lens, const int64_t *row_idx,
                             int64_t n, int64_t data_off, int64_t payload_bytes, uint8_t *dst, int32_t *status,
                             cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int64_t slices = payload_bytes > 0 ? (payload_bytes + kSlice - 1) / kSlice : 1;
    int g = (int)(n * slices < 148 * 32 ? n * slices : 148 * 32);
    int threads = payload_bytes >= 4096 ? 256 : (payload_bytes >= 512 ? 128 : 32);
    k_npy_batch<<<g, threads, 0, s>>>(base, offs, lens, row_idx, n, data_off, payload_bytes, slices, dst, status);
    return cudaGetLastError();
}

// first `k` bytes of every blob -> dst[n, k] (zero padded): lets the host read .npy / PNG headers of a whole
// row-group with one small D2H when field shapes are variable
__global__ void k_blob_prefix(const uint8_t *__restrict__ base, const int64_t *__restrict__ offs,
                              const int32_t *__restrict__ lens, int64_t n, int k, uint8_t *__restrict__ dst) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * k; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = t / k;
        int b = (int)(t % k);
        dst[t] = b < lens[i] ? base[offs[i] + b] : 0;
    }
}
cudaError_t launch_blob_prefix(const uint8_t *base, const int64_t *offs, const int32_t *lens, int64_t n, int k,
                               uint8_t *dst, cudaStream_t s) {
    if (n <= 0 || k <= 0) return cudaSuccess;
    k_blob_prefix<<<grid_for(n * k, kThreads), kThreads, 0, s>>>(base, offs, lens, n, k, dst);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// K11 predicates
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t load_key(const void *keys, int key_bytes, int is_unsigned, int64_t i) {
    switch (key_bytes) {
        case 1: return is_unsigned ? (int64_t)((const uint8_t *)keys)[i] : (int64_t)((const int8_t *)keys)[i];
        case 2: return is_unsigned ? (int64_t)((const uint16_t *)keys)[i] : (int64_t)((const int16_t *)keys)[i];
        case 4: return is_unsigned ? (int64_t)((const uint32_t *)keys)[i] : (int64_t)((const int32_t *)keys)[i];
        default: return ((const int64_t *)keys)[i];
    }
}

// in_set (petastorm/predicates.py:44-55): membership by binary search in a sorted int64 set
__global__ void k_mask_in_set(const void *__restrict__ keys, int key_bytes, int is_unsigned, int64_t n,
                              const int64_t *__restrict__ set_sorted, int64_t set_n, uint8_t *__restrict__ mask) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = load_key(keys, key_bytes, is_unsigned, i);
        int64_t lo = 0, hi = set_n;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (set_sorted[mid] < k) lo = mid + 1; else hi = mid;
        }
        mask[i] = (lo < set_n && set_sorted[lo] == k) ? 1 : 0;
    }
}
cudaError_t launch_mask_in_set(const void *keys, int key_bytes, int key_unsigned, int64_t n, const int64_t *set_sorted,
                               int64_t set_n, uint8_t *mask, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_mask_in_set<<<grid_for(n, kThreads), kThreads, 0, s>>>(keys, key_bytes, key_unsigned, n, set_sorted, set_n, mask);
    return cudaGetLastError();
}

// in_pseudorandom_split (petastorm/predicates.py:39-41,144-182):
//   bucket = int(md5(str(value).encode()).hexdigest(), 16) % sys.maxsize ; keep iff lo <= bucket < hi
// MD5 per RFC 1321, single 64-byte block (the decimal string of an int64 is at most 20 characters).
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }

__device__ void md5_short(const uint8_t *msg, int len, uint32_t digest[4]) {
    const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8,
        0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340,
        0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87,
        0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
        0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039,
        0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92,
        0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb,
        0xeb86d391};
    const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20,
                       5, 9, 14, 20, 5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                       6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t M[16];
#pragma unroll
    for (int i = 0; i < 16; i++) M[i] = 0;
    for (int i = 0; i < len; i++) M[i >> 2] |= (uint32_t)msg[i] << (8 * (i & 3));
    M[len >> 2] |= 0x80u << (8 * (len & 3));
    M[14] = (uint32_t)len * 8;
    uint32_t a = 0x67452301, b = 0xefcdab89, c = 0x98badcfe, d = 0x10325476;
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        f = f + a + K[i] + M[g];
        a = d; d = c; c = b;
        b = b + rotl32(f, S[i]);
    }
    digest[0] = a + 0x67452301; digest[1] = b + 0xefcdab89; digest[2] = c + 0x98badcfe; digest[3] = d + 0x10325476;
}

__device__ __forceinline__ uint64_t bswap64(uint64_t v) {
    return ((uint64_t)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123);
}

__global__ void k_mask_md5_split(const void *__restrict__ keys, int key_bytes, int is_unsigned, int64_t n,
                                 uint64_t lo, uint64_t hi, uint8_t *__restrict__ mask) {
    const uint64_t M = 0x7fffffffffffffffULL;  // sys.maxsize
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = load_key(keys, key_bytes, is_unsigned, i);
        // str(k)
        uint8_t buf[24];
        int len = 0;
        uint64_t u = k < 0 ? (uint64_t)0 - (uint64_t)k : (uint64_t)k;
        if (key_bytes == 8 && is_unsigned) u = (uint64_t)k;
        uint8_t tmp[20];
        int nd = 0;
        do { tmp[nd++] = (uint8_t)('0' + (u % 10)); u /= 10; } while (u);
        if (k < 0 && !(key_bytes == 8 && is_unsigned)) buf[len++] = '-';
        while (nd) buf[len++] = tmp[--nd];
        uint32_t dg[4];
        md5_short(buf, len, dg);
        // hexdigest is the 16 digest bytes in order -> big-endian 128-bit integer
        uint64_t hi64 = bswap64((uint64_t)dg[0] | ((uint64_t)dg[1] << 32));
        uint64_t lo64 = bswap64((uint64_t)dg[2] | ((uint64_t)dg[3] << 32));
        // N = hi64 * 2^64 + lo64 ; 2^63 == 1 (mod M)  =>  2^64 == 2
        uint64_t a = hi64 % M, b = lo64 % M;
        uint64_t r = ((a * 2) % M + b) % M;
        mask[i] = (r >= lo && r < hi) ? 1 : 0;
    }
}
cudaError_t launch_mask_md5_split(const void *keys, int key_bytes, int key_unsigned, int64_t n, double lo, double hi,
                                  uint8_t *mask, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    // thresholds arrive as exact integers encoded in doubles' integer range by the host (see native.py); the host
    // passes ceil() of the python floats so that integer comparison reproduces python's exact int/float comparison
    auto to_u64 = [](double d) -> uint64_t {
        if (d <= 0) return 0;
        if (d >= 18446744073709551615.0) return ~0ULL;
        return (uint64_t)d;
    };
    k_mask_md5_split<<<grid_for(n, kThreads), kThreads, 0, s>>>(keys, key_bytes, key_unsigned, n, to_u64(lo), to_u64(hi), mask);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// stream compaction: mask -> ascending row indices (three small kernels: per-block counts, scan, scatter)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCompactBlock = 1024;  // elements per block (256 threads x 4)

__global__ void k_compact_count(const uint8_t *__restrict__ mask, int64_t n, int64_t *__restrict__ block_counts) {
    __shared__ int s[kThreads / 32];
    int64_t base = (int64_t)blockIdx.x * kCompactBlock;
    int c = 0;
    for (int k = 0; k < 4; k++) {
        int64_t i = base + k * kThreads + threadIdx.x;
        if (i < n && mask[i]) c++;
    }
    for (int d = 16; d; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; w++) t += s[w];
        block_counts[blockIdx.x] = t;
    }
}
__global__ void k_compact_scan(int64_t *__restrict__ block_counts, int64_t nblocks, int64_t *__restrict__ total) {
    // single block; serial over chunks of 1024 with a block scan inside
    __shared__ int64_t s[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += 1024) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < nblocks ? block_counts[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            int64_t t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_counts[i] = carry + s[threadIdx.x] - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void k_compact_scatter(const uint8_t *__restrict__ mask, int64_t n, const int64_t *__restrict__ block_offs,
                                  int64_t *__restrict__ out_idx) {
    __shared__ int warp_tot[4][kThreads / 32];
    int64_t base = (int64_t)blockIdx.x * kCompactBlock;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t off = block_offs[blockIdx.x];
    // ascending order: element index = base + k*256 + tid  -> process k sequentially
    int flags[4];
    int incl[4];
    for (int k = 0; k < 4; k++) {
        int64_t i = base + k * kThreads + threadIdx.x;
        int f = (i < n && mask[i]) ? 1 : 0;
        flags[k] = f;
        int v = f;
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        incl[k] = v;
        if (lane == 31) warp_tot[k][warp] = v;
    }
    __syncthreads();
    int running = 0;
    for (int k = 0; k < 4; k++) {
        int woff = 0, tot = 0;
        for (int w = 0; w < kThreads / 32; w++) {
            int t = warp_tot[k][w];
            if (w < warp) woff += t;
            tot += t;
        }
        if (flags[k]) out_idx[off + running + woff + incl[k] - 1] = base + k * kThreads + threadIdx.x;
        running += tot;
    }
}
int64_t compact_tmp_bytes(int64_t n) { return ((n + kCompactBlock - 1) / kCompactBlock + 1) * 8; }
cudaError_t launch_mask_compact(const uint8_t *mask, int64_t n, int64_t *out_idx, int64_t *count, void *tmp,
                                cudaStream_t s) {
    if (n <= 0) return cudaMemsetAsync(count, 0, 8, s);
    int64_t nb = (n + kCompactBlock - 1) / kCompactBlock;
    int64_t *bc = (int64_t *)tmp;
    k_compact_count<<<(int)nb, kThreads, 0, s>>>(mask, n, bc);
    k_compact_scan<<<1, 1024, 0, s>>>(bc, nb, count);
    k_compact_scatter<<<(int)nb, kThreads, 0, s>>>(mask, n, bc, out_idx);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// K12 TransformSpec normalise: out = ((float)x - mean) / std cast to the output dtype; IEEE fp32 subtract + divide
// (no fast-math, no reciprocal) so the result is bit-identical to numpy's float32 arithmetic.
// dtype codes: 0=u8 1=f16 2=f32 3=i32 4=i16 5=u16 6=f64
// ---------------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v) { return (T)v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

template <typename S, typename D>
__global__ void k_normalize(const S *__restrict__ src, int64_t n, float mean, float stddev, D *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = to_f32<S>(src[i]);
        float y = __fdiv_rn(__fsub_rn(x, mean), stddev);
        dst[i] = from_f32<D>(y);
    }
}
// vectorised f16 -> f16 / f32 (8 elements per thread), the C4 shape
__global__ void k_normalize_h8(const uint4 *__restrict__ src, int64_t nvec, float mean, float stddev,
                               uint4 *__restrict__ dst_h, float4 *__restrict__ dst_f) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v = src[i];
        const __half *h = reinterpret_cast<const __half *>(&v);
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; k++) y[k] = __fdiv_rn(__fsub_rn(__half2float(h[k]), mean), stddev);
        if (dst_h) {
            uint4 o;
            __half *oh = reinterpret_cast<__half *>(&o);
#pragma unroll
            for (int k = 0; k < 8; k++) oh[k] = __float2half_rn(y[k]);
            dst_h[i] = o;
        } else {
            dst_f[2 * i] = make_float4(y[0], y[1], y[2], y[3]);
            dst_f[2 * i + 1] = make_float4(y[4], y[5], y[6], y[7]);
        }
    }
}

template <typename S>
static cudaError_t normalize_dispatch_dst(const S *src, int64_t n, float mean, float stddev, void *dst, int dst_dtype,
                                          cudaStream_t s) {
    int g = grid_for(n, kThreads);
    switch (dst_dtype) {
        case 1: k_normalize<S, __half><<<g, kThreads, 0, s>>>(src, n, mean, stddev, (__half *)dst); break;
        case 2: k_normalize<S, float><<<g, kThreads, 0, s>>>(src, n, mean, stddev, (float *)dst); break;
        case 6: k_normalize<S, double><<<g, kThreads, 0, s>>>(src, n, mean, stddev, (double *)dst); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_normalize(const void *src, int src_dtype, int64_t n, float mean, float stddev, void *dst,
                             int dst_dtype, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    if (src_dtype == 1 && (dst_dtype == 1 || dst_dtype == 2) && (n % 8) == 0 &&
        (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        int64_t nvec = n / 8;
        k_normalize_h8<<<grid_for(nvec, kThreads), kThreads, 0, s>>>((const uint4 *)src, nvec, mean, stddev,
                                                                      dst_dtype == 1 ? (uint4 *)dst : nullptr,
                                                                      dst_dtype == 2 ? (float4 *)dst : nullptr);
        return cudaGetLastError();
    }
    switch (src_dtype) {
        case 0: return normalize_dispatch_dst((const uint8_t *)src, n, mean, stddev, dst, dst_dtype, s);
        case 1: return normalize_dispatch_dst((const __half *)src, n, mean, stddev, dst, dst_dtype, s);
        case 2: return normalize_dispatch_dst((const float *)src, n, mean, stddev, dst, dst_dtype, s);
        case 3: return normalize_dispatch_dst((const int32_t *)src, n, mean, stddev, dst, dst_dtype, s);
        case 4: return normalize_dispatch_dst((const int16_t *)src, n, mean, stddev, dst, dst_dtype, s);
        case 5: return normalize_dispatch_dst((const uint16_t *)src, n, mean, stddev, dst, dst_dtype, s);
        case 6: return normalize_dispatch_dst((const double *)src, n, mean, stddev, dst, dst_dtype, s);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K15 NGram (petastorm/ngram.py:225-270): start i is a window iff rows i..i+L-1 exist and every consecutive timestamp
// gap is <= delta; unsorted timestamps inside any candidate window are an error (NotImplementedError upstream).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_ngram_valid(const int64_t *__restrict__ ts, int64_t n, int length, int64_t delta,
                              uint8_t *__restrict__ ok, int32_t *status) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t good = 0;
        if (i + length <= n) {
            good = 1;
            int64_t prev = ts[i];
            for (int j = 1; j < length; j++) {
                int64_t cur = ts[i + j];
                if (cur < prev) report_err(status, DE_NGRAM_UNSORTED, (int)i, j);
                if (cur - prev > delta) good = 0;
                prev = cur;
            }
        }
        ok[i] = good;
    }
}
cudaError_t launch_ngram_valid_starts(const int64_t *ts, int64_t n, int length, int64_t delta, uint8_t *ok,
                                      int32_t *status, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_ngram_valid<<<grid_for(n, kThreads), kThreads, 0, s>>>(ts, n, length, delta, ok, status);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// K16 dtype sanitise (petastorm/pytorch.py:40-70): uint16->int32, uint32->int64, bool->uint8
// ---------------------------------------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void k_cast(const S *__restrict__ src, int64_t n, D *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (D)src[i];
}
cudaError_t launch_sanitize(const void *src, int64_t n, int kind, void *dst, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int g = grid_for(n, kThreads);
    switch (kind) {
        case 0: k_cast<uint16_t, int32_t><<<g, kThreads, 0, s>>>((const uint16_t *)src, n, (int32_t *)dst); break;
        case 1: k_cast<uint32_t, int64_t><<<g, kThreads, 0, s>>>((const uint32_t *)src, n, (int64_t *)dst); break;
        case 2: k_cast<uint8_t, uint8_t><<<g, kThreads, 0, s>>>((const uint8_t *)src, n, (uint8_t *)dst); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// list columns: check that a repeated column holds uniform-length, fully-defined lists, i.e. that the level entries
// are exactly n_rows x L elements (the only shape `np.vstack(list_of_lists)` accepts --
// petastorm/arrow_reader_worker.py:68-76).  flags[0] = 1 if not uniform.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_list_uniform(const uint8_t *__restrict__ rep, const uint8_t *__restrict__ def, int64_t n, int max_def,
                               int64_t L, int64_t *__restrict__ flags) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool start = rep[i] == 0;
        bool expect = (L > 0) && (i % L == 0);
        if (start != expect || def[i] != max_def) flags[0] = 1;
    }
}
cudaError_t launch_list_uniform(const uint8_t *rep, const uint8_t *def, int64_t n, int max_def, int64_t L,
                                int64_t *flags, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_list_uniform<<<grid_for(n, kThreads), kThreads, 0, s>>>(rep, def, n, max_def, L, flags);
    return cudaGetLastError();
}

}  // namespace pst
