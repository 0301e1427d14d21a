// Launchers of the device kernels (definitions in kernels_decode.cu / kernels_ops.cu / kernels_png.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"

namespace pst {

constexpr int PST_BOOLEAN_T = 0;
constexpr int PST_INT32_T = 1;
constexpr int PST_INT64_T = 2;
constexpr int PST_INT96_T = 3;
constexpr int PST_FLOAT_T = 4;
constexpr int PST_DOUBLE_T = 5;
constexpr int PST_BYTE_ARRAY_T = 6;
constexpr int PST_FLBA_T = 7;

cudaError_t configure_decode_kernels();
// Snappy runs as up to three launches: fragment index of the multi-fragment pages, all fragments in parallel
// (serial_mode 0), serial fallback for pages the first two flagged (serial_mode 1)
cudaError_t launch_snappy_index(uint8_t *arena, const DevPage *pages, const int32_t *multi_list, int n_multi,
                                uint32_t *frag_pos, uint32_t *page_flag, cudaStream_t s);
// the same for big pages: one cluster of four CTAs per page (walker CTA + three builder CTAs, tables through DSMEM)
cudaError_t launch_snappy_index_cluster(uint8_t *arena, const DevPage *pages, const int32_t *list, int n_list,
                                        uint32_t *frag_pos, uint32_t *page_flag, cudaStream_t s);
cudaError_t launch_snappy(uint8_t *arena, const DevPage *pages, const SnFrag *frags, int n_frags,
                          const int32_t *multi_list, int n_multi, const uint32_t *frag_pos, uint32_t *page_flag,
                          int32_t *status, int serial_mode, cudaStream_t s);
cudaError_t launch_ba_dict_index(uint8_t *arena, const DevPage *pages, const DevCol *cols, const int32_t *list, int n,
                                 int32_t *status, cudaStream_t s);
cudaError_t launch_decode_pages(uint8_t *arena, uint8_t *out, const DevCol *cols, const DevPage *pages,
                                const int32_t *list, int n, int32_t *status, cudaStream_t s);

// ---- kernels_copy.cu: PLAIN value tiles -> out with the bulk-copy engine (cp.async.bulk + mbarrier)
cudaError_t configure_copy_kernel();
cudaError_t launch_copy_tiles(const uint8_t *arena, uint8_t *out, const CopyTile *tiles, int n_tiles, int sm_count,
                              cudaStream_t s);

// ---- kernels_ops.cu
cudaError_t launch_nullable_to_f64(const void *values, const uint8_t *valid, int64_t n, int ptype, int bits,
                                   int is_unsigned, double *out, cudaStream_t s);
cudaError_t launch_narrow_int32(const int32_t *src, int64_t n, int bits, void *dst, cudaStream_t s);
cudaError_t launch_gather_rows(const uint8_t *src, const int64_t *idx, int64_t n_out, int64_t row_bytes, uint8_t *dst,
                               cudaStream_t s);
cudaError_t launch_npy_batch(const uint8_t *base, const int64_t *offs, const int32_t *lens, const int64_t *row_idx,
                             int64_t n, int64_t data_off, int64_t payload_bytes, uint8_t *dst, int32_t *status,
                             cudaStream_t s);
cudaError_t launch_blob_prefix(const uint8_t *base, const int64_t *offs, const int32_t *lens, int64_t n, int k,
                               uint8_t *dst, cudaStream_t s);
cudaError_t launch_mask_in_set(const void *keys, int key_bytes, int key_unsigned, int64_t n, const int64_t *set_sorted,
                               int64_t set_n, uint8_t *mask, cudaStream_t s);
cudaError_t launch_mask_md5_split(const void *keys, int key_bytes, int key_unsigned, int64_t n, double lo, double hi,
                                  uint8_t *mask, cudaStream_t s);
int64_t compact_tmp_bytes(int64_t n);
cudaError_t launch_mask_compact(const uint8_t *mask, int64_t n, int64_t *out_idx, int64_t *count, void *tmp,
                                cudaStream_t s);
cudaError_t launch_normalize(const void *src, int src_dtype, int64_t n, float mean, float stddev, void *dst,
                             int dst_dtype, cudaStream_t s);
cudaError_t launch_ngram_valid_starts(const int64_t *ts, int64_t n, int length, int64_t delta, uint8_t *ok,
                                      int32_t *status, cudaStream_t s);
cudaError_t launch_ngram_gather(const uint8_t *src, const int64_t *starts, int64_t n_windows, int length,
                                int64_t row_bytes, uint8_t *dst, cudaStream_t s);
cudaError_t launch_sanitize(const void *src, int64_t n, int kind, void *dst, cudaStream_t s);
cudaError_t launch_list_uniform(const uint8_t *rep, const uint8_t *def, int64_t n, int max_def, int64_t L,
                                int64_t *flags, cudaStream_t s);

// ---- kernels_png.cu (DEFLATE users: PNG images, GZIP-compressed pages)
cudaError_t launch_zip_inflate_batch(const uint8_t *base, const int64_t *offs, const int32_t *lens,
                                     const int64_t *row_idx, int64_t n, int64_t member_bytes, uint8_t *dst,
                                     int32_t *status, cudaStream_t s);
cudaError_t launch_gzip_pages(uint8_t *arena, const DevPage *pages, const int32_t *list, int n, int32_t *status,
                              cudaStream_t s);
int64_t png_work_bytes(int height, int width, int channels, int sample_bytes);
cudaError_t launch_png_batch(const uint8_t *base, const int64_t *offs, const int32_t *lens, const int64_t *row_idx,
                             int64_t n, int height, int width, int channels, int sample_bytes, uint8_t *dst,
                             uint8_t *work, int32_t *status, cudaStream_t s);

}  // namespace pst
