/*
 * pst_b200.h -- C-ABI of libpst_b200.so: the B200-native Parquet row-group -> decoded device tensors path.
 *
 * The reference (uber/petastorm 0.13.1) has NO FFI of its own: it is pure Python and every byte of decode work
 * happens inside third-party native libraries that it calls.  This header is therefore the
 * boundary a maintainer of the reference would bind (ctypes) to replace those calls; each entry point cites the
 * reference call site it replaces (paths relative to /root/reference).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C, no C++/torch types cross the boundary; handles are opaque pointers freed by the matching *_close /
 *     *_destroy; every function returning int returns 0 on success, non-zero on failure, and the failure text is
 *     available (thread-local) from pst_last_error().
 *   - device memory is OWNED BY THE CALLER (the Python host allocates it from the torch caching allocator and passes
 *     raw addresses + a cudaStream_t as uintptr_t).  All device work is enqueued on the stream passed in; nothing
 *     here synchronises the device unless documented.
 *   - host-only entry points (pst_file_*, pst_plan_create/…_info) work on a machine without a GPU.
 */
#ifndef PST_B200_H_
#define PST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PST_ABI_VERSION 2

typedef struct pst_file pst_file;
typedef struct pst_plan pst_plan;
typedef struct pst_ctx pst_ctx;

/* ---- parquet physical types / encodings / codecs (values = parquet.thrift enums) ---- */
enum { PST_BOOLEAN = 0, PST_INT32 = 1, PST_INT64 = 2, PST_INT96 = 3, PST_FLOAT = 4, PST_DOUBLE = 5,
       PST_BYTE_ARRAY = 6, PST_FIXED_LEN_BYTE_ARRAY = 7 };
enum { PST_CODEC_NONE = 0, PST_CODEC_SNAPPY = 1, PST_CODEC_GZIP = 2 };

const char *pst_last_error(void);
int pst_abi_version(void);
/* 1 when the library was built with its CUDA kernels (always, for the shipped .so) */
int pst_has_cuda(void);

/* ------------------------------------------------------------------------------------------------------------------
 * File + footer.  Replaces pq.ParquetFile(fs.open(piece.path)) + metadata access:
 *   petastorm/arrow_reader_worker.py:172, petastorm/py_dict_reader_worker.py:146,
 *   petastorm/etl/dataset_metadata.py:340-353 (footer walk), :356-385 (_common_metadata key/values).
 * The file is mmapped read-only; the thrift-compact FileMetaData is parsed on the host.
 * ------------------------------------------------------------------------------------------------------------------ */
int pst_file_open(const char *path, pst_file **out);
void pst_file_close(pst_file *f);
int pst_file_num_row_groups(const pst_file *f);
int64_t pst_file_num_rows(const pst_file *f);
int64_t pst_file_row_group_num_rows(const pst_file *f, int rg);
/* number of leaf columns */
int pst_file_num_columns(const pst_file *f);
/* JSON description of the schema tree leaves: name/path, physical type, converted+logical type, repetition,
 * max definition / repetition level, type_length, precision/scale.  Pointer valid until pst_file_close. */
int pst_file_schema_json(const pst_file *f, const char **json, size_t *len);
/* value of a footer key/value entry (e.g. b'dataset-toolkit.unischema.v1'); returns 1 if the key is absent */
int pst_file_kv_metadata(const pst_file *f, const char *key, const uint8_t **val, size_t *len);
int pst_file_num_kv(const pst_file *f);
int pst_file_kv_at(const pst_file *f, int i, const char **key, size_t *klen, const uint8_t **val, size_t *vlen);

typedef struct pst_chunk_info {
    int32_t physical_type;
    int32_t codec;
    int64_t num_values;
    int64_t data_page_offset;
    int64_t dictionary_page_offset; /* -1 if none */
    int64_t total_compressed_size;
    int64_t total_uncompressed_size;
    int64_t start_offset; /* first byte of the chunk in the file */
} pst_chunk_info;
int pst_file_chunk_info(const pst_file *f, int rg, int col, pst_chunk_info *out);

/* ------------------------------------------------------------------------------------------------------------------
 * Row-group plan (host): page-header walk of the requested column chunks + HBM layout.
 * Replaces the host half of piece.read(columns=…) -- petastorm/arrow_reader_worker.py:358,
 * petastorm/py_dict_reader_worker.py:267 (Arrow C++ page reader).
 *
 * Layout it decides (all offsets 16-byte aligned, see DESIGN.md "HBM layout"):
 *   raw     : page payloads exactly as stored in the file (compressed or not) + the page/column tables
 *   scratch : decompressed image of every compressed page
 *   out     : one dense region per requested column (+ validity bytes, + (offset,len) pairs for BYTE_ARRAY)
 * ------------------------------------------------------------------------------------------------------------------ */
int pst_plan_create(const pst_file *f, int rg, const int *cols, int ncols, pst_plan **out);
void pst_plan_destroy(pst_plan *p);

typedef struct pst_plan_info {
    int64_t num_rows;
    int64_t raw_bytes;      /* size of the raw region = arena[0:raw_bytes] (payloads + tables); this crosses PCIe */
    int64_t arena_bytes;    /* raw region + decompression scratch: the device allocation the caller must provide */
    int64_t out_bytes;      /* size of the output region */
    int64_t payload_bytes;  /* sum of page payload bytes as stored in the file */
    int64_t uncompressed_bytes;
    int32_t num_pages;
    int32_t num_columns;
    int32_t num_compressed_pages;
    int32_t num_index_pages; /* Snappy pages whose 64 KiB fragment boundaries the device has to find (k_snappy_index) */
    int32_t num_unwrapped_pages;  /* literal-only Snappy pages (incompressible data): the staging copy drops the framing
                                     and the device receives them as uncompressed page images */
    int32_t num_copy_tiles;       /* <= 64 KiB work items of k_copy_tiles (PLAIN fixed-width pages without nulls) */
    int32_t num_decode_pages;     /* data pages that go through the general page decoder k_decode_pages */
    int32_t num_snappy_fragments; /* work items of k_snappy_pages */
    int32_t num_host_indexed_pages; /* multi-fragment Snappy pages of literal-dominated streams (blob columns): the planner
                                       walks their few tags itself, the others go through k_snappy_index */
    int32_t num_cluster_index_pages;/* of the pages the device indexes (num_index_pages): those of >= 256 KiB stored bytes,
                                       indexed by one four-CTA cluster each (k_snappy_index_cluster) */
} pst_plan_info;
int pst_plan_get_info(const pst_plan *p, pst_plan_info *out);

typedef struct pst_plan_column {
    int32_t column;         /* leaf column index in the file */
    int32_t physical_type;
    int32_t type_length;    /* bytes per value in `values` (1 for BOOLEAN, 12 for INT96, FLBA length, 0 for BYTE_ARRAY) */
    int32_t max_def;
    int32_t max_rep;
    int32_t has_dictionary;
    int64_t num_values;     /* level entries in the chunk (== rows for flat columns) */
    int64_t values_off;     /* offset in out region: num_values * type_length bytes; BYTE_ARRAY: int64 arena offsets */
    int64_t lens_off;       /* BYTE_ARRAY only: int32 byte lengths[num_values] (else -1) */
    int64_t valid_off;      /* offset in out region of num_values validity bytes; -1 when max_def == 0, and for a flat
                               column whose chunk statistics promise null_count == 0 (the decoder still counts nulls in
                               d_status[8 + slot]; a non-zero count for such a column means the statistics lied) */
    int64_t rep_off;        /* offset in out region of num_values repetition-level bytes, or -1 when max_rep == 0 */
    int64_t def_off;        /* offset in out region of num_values definition-level bytes (only when max_rep > 0), else -1 */
} pst_plan_column;
int pst_plan_get_column(const pst_plan *p, int i, pst_plan_column *out);
/* Introspection of the plan (host only; used by the CPU tests of the planner and by bench.py's traffic accounting). */
typedef struct pst_plan_page {
    int32_t column_slot;    /* plan column the page belongs to */
    int32_t kind;           /* 0 DATA_PAGE, 2 DICTIONARY_PAGE, 3 DATA_PAGE_V2 */
    int32_t encoding;       /* parquet Encoding of the values */
    int32_t codec;          /* codec the DEVICE sees: a literal-only Snappy page is delivered uncompressed (0) */
    int32_t flags;          /* 1 all-valid (planner read the levels), 2 values go through k_copy_tiles, 4 unwrapped */
    int32_t stored_bytes;   /* payload bytes in the file */
    int32_t image_bytes;    /* uncompressed page image bytes */
    int32_t num_values;
    int32_t first_value;
    int32_t fragments;      /* Snappy work items of the page (0 when the device does not decompress it) */
    int64_t src_off;        /* arena offset of the payload in the raw region */
    int64_t img_off;        /* arena offset of the page image (== src_off for pages the device sees uncompressed) */
} pst_plan_page;
int pst_plan_get_page(const pst_plan *p, int i, pst_plan_page *out);
typedef struct pst_copy_tile {
    int64_t src_off;        /* arena offset */
    int64_t dst_off;        /* out offset */
    int64_t valid_off;      /* out offset of the validity bytes to set to 1, or -1 */
    int32_t nbytes;
    int32_t nvalid;
} pst_copy_tile;
int pst_plan_get_copy_tile(const pst_plan *p, int i, pst_copy_tile *out);

/* Host helper: write the raw-region image (page payloads [first_page,last_page) at their planned offsets, plus the
 * tables when first_page == 0) into dst[0:raw_bytes].  Used by the staging threads and by CPU tests of the planner. */
int pst_plan_fill_raw(const pst_plan *p, uint8_t *dst, int64_t first_page, int64_t last_page);

/* ------------------------------------------------------------------------------------------------------------------
 * Context: device, pinned staging ring / pinned row-group cache, host copy threads.
 * ------------------------------------------------------------------------------------------------------------------ */
int pst_ctx_create(int device, int64_t pinned_cache_bytes, int copy_threads, pst_ctx **out);
void pst_ctx_destroy(pst_ctx *c);
/* Changes the budget of the pinned row-group cache of a live context; shrinking below the bytes in use drops every cached
 * row-group (after a device synchronisation: copies in flight may still read them). */
int pst_ctx_set_pinned_cache_bytes(pst_ctx *c, int64_t nbytes);
/* JSON counters: bytes staged, bytes H2D, pages decoded, cache hits… (Reader.diagnostics, petastorm/reader.py:701-703) */
int pst_ctx_stats_json(pst_ctx *c, char *buf, size_t cap);

/* Stage a plan's raw region into pinned host memory (parallel memcpy from the mmap; kept in the pinned row-group
 * cache when it fits the budget) and enqueue ONE cudaMemcpyAsync host->device into d_arena[0:raw_bytes] on `stream`. */
int pst_plan_upload(pst_ctx *c, pst_plan *p, uint64_t d_arena, uint64_t stream);

/* Enqueue the device decode of a plan whose raw region is resident at d_arena:
 *   K2 snappy -> scratch, value tiles of PLAIN pages without nulls -> out (bulk-copy engine), then for the other
 *   pages K3 RLE/bit-packed levels, K4 PLAIN, K5 dictionary gather, K6 validity.
 * After the stream reaches this point, `out` holds the columns described by pst_plan_get_column.
 * d_status: device int32[8] {error_code, page, detail, ...} written by the kernels (all zero == ok); the caller
 * zeroes it before the call and checks it after synchronising.  Returns the number of kernels launched in
 * *launches (may be NULL). */
int pst_plan_decode(pst_ctx *c, pst_plan *p, uint64_t d_arena, uint64_t d_out, uint64_t d_status, uint64_t stream,
                    int *launches);

/* Measurement aid (bench.py roofline): the same launches with CUDA events between them on `stream`; synchronises and
 * writes the device milliseconds of {snappy fragment index, snappy fragments, snappy serial fallback, byte-array
 * dictionary index, value tile copy, page decode} to ms6[0..5]. */
int pst_plan_decode_timed(pst_ctx *c, pst_plan *p, uint64_t d_arena, uint64_t d_out, uint64_t d_status, uint64_t stream,
                          float *ms6);

/* ------------------------------------------------------------------------------------------------------------------
 * Column post-processing kernels (all async on `stream`; pointers are device addresses).
 * ------------------------------------------------------------------------------------------------------------------ */

/* K6: dense fixed-width values + validity bytes -> float64 with NaN at nulls.  Reproduces pandas' nullable-int ->
 * float64 promotion of `column.to_pandas()` (petastorm/arrow_reader_worker.py:55-57).  src_kind: physical type,
 * is_unsigned/bit_width describe the logical integer type. */
int pst_nullable_to_f64(uint64_t values, uint64_t valid, int64_t n, int physical_type, int bit_width, int is_unsigned,
                        uint64_t out_f64, uint64_t stream);

/* K16/K4 cast: int32 storage -> int8/uint8/int16/uint16 logical types (parquet stores them as INT32) */
int pst_narrow_int32(uint64_t src_i32, int64_t n, int bit_width, uint64_t dst, uint64_t stream);

/* K13: row gather out[i] = src[idx[i]] for rows of `row_bytes` bytes.  Replaces table.take(indices)
 * (petastorm/arrow_reader_worker.py:361-371) and data_frame.sample (petastorm/py_dict_reader_worker.py:269-270). */
int pst_gather_rows(uint64_t src, uint64_t idx_i64, int64_t n_out, int64_t row_bytes, uint64_t dst, uint64_t stream);

/* K7: NdarrayCodec.  offs/lens = int64 arena offsets + int32 byte lengths per row (output of the BYTE_ARRAY decode,
 * see pst_plan_column), base = the arena; row_idx_i64 (may be 0) selects/permutes rows.  Each value is a .npy blob
 * whose payload starts `data_off` bytes in and is `payload_bytes` long; the payloads are copied into a dense
 * [n, payload_bytes] tensor.  Replaces np.load(BytesIO(value)) -- petastorm/codecs.py:155-157.
 * The header bytes of EVERY blob are compared on-device with the first selected row's header; d_status[0] != 0 if any
 * differs (ragged shapes then take the per-group path on the host side). */
int pst_npy_batch(uint64_t base, uint64_t offs_i64, uint64_t lens_i32, uint64_t row_idx_i64, int64_t n,
                  int64_t data_off, int64_t payload_bytes, uint64_t dst, uint64_t d_status, uint64_t stream);

/* CompressedNdarrayCodec values (np.savez_compressed ZIP archives): the first member of every selected blob is
 * inflated (DEFLATE) or copied (stored) into dst[i * member_bytes ..); d_status[0] != 0 when a blob is not a ZIP
 * archive or its member does not have exactly member_bytes bytes.  The .npy images then go through pst_npy_batch.
 * Replaces np.load(BytesIO(value))['arr'] -- petastorm/codecs.py:196-198. */
int pst_zip_inflate_batch(uint64_t base, uint64_t offs_i64, uint64_t lens_i32, uint64_t row_idx_i64, int64_t n,
                          int64_t member_bytes, uint64_t dst, uint64_t d_status, uint64_t stream);

/* first k bytes of every BYTE_ARRAY value -> dst[n, k] (zero padded).  Lets the host read the .npy / PNG headers of a
 * whole row-group with one small D2H when a field's shape is variable (None dimensions in the Unischema). */
int pst_blob_prefix(uint64_t base, uint64_t offs_i64, uint64_t lens_i32, int64_t n, int k, uint64_t dst,
                    uint64_t stream);

/* K8: PNG (zlib inflate: stored / fixed / dynamic Huffman; filters None/Sub/Up/Average/Paeth; 8/16-bit gray, RGB,
 * palette) one image per warp.  Replaces cv2.imdecode + BGR->RGB reorder -- petastorm/codecs.py:102-116.
 * dst is [n, height, width, channels] of `sample_bytes`-byte samples in RGB order, native endianness.
 * d_work: device scratch, n * pst_png_work_bytes(...) bytes.  d_status[0]!=0 on malformed / unsupported streams. */
int64_t pst_png_work_bytes(int height, int width, int channels, int sample_bytes);
int pst_png_batch(uint64_t base, uint64_t offs_i64, uint64_t lens_i32, uint64_t row_idx_i64, int64_t n, int height,
                  int width, int channels, int sample_bytes, uint64_t dst, uint64_t d_work, uint64_t d_status,
                  uint64_t stream);

/* K9: JPEG through nvJPEG (library) -> interleaved RGB u8 [n, height, width, 3].  `host_blobs`/`host_lens` are host
 * pointers to the n bitstreams (nvJPEG parses Huffman tables on the host).  Replaces cv2.imdecode for '.jpeg'. */
int pst_jpeg_available(void);  /* 1 when libnvjpeg could be dlopen'ed */
int pst_jpeg_backend(void);    /* nvjpegBackend_t in use, -1 before the first call */
int pst_jpeg_batch(pst_ctx *c, const uint8_t *const *host_blobs, const size_t *host_lens, int64_t n, int height,
                   int width, uint64_t dst, uint64_t stream);
/* The same with the bitstreams left where the Parquet decode put them: image i is the `host_lens[i]` bytes at DEVICE
 * address base + host_offs[i] (the host only holds the 12 bytes of (offset, length) per image).  Uses nvJPEG's
 * device-bitstream backends: the hardware JPEG engines (NVJPEG_BACKEND_HARDWARE_DEVICE) when this GPU/driver exposes
 * them, else GPU-assisted Huffman (NVJPEG_BACKEND_GPU_HYBRID_DEVICE).  pst_jpeg_device_backend() creates that handle on
 * first use and returns its nvjpegBackend_t, or -1 when neither exists (callers then stage the blobs on the host). */
int pst_jpeg_device_backend(void);
int pst_jpeg_batch_device(pst_ctx *c, uint64_t base, const int64_t *host_offs, const int32_t *host_lens, int64_t n,
                          int height, int width, uint64_t dst, uint64_t stream);

/* K11: predicate masks.  in_set on an integer key column (petastorm/predicates.py:44-55): mask[i] = key[i] in set.
 * `set_sorted` is a sorted device array of int64. */
int pst_mask_in_set_i64(uint64_t keys, int key_bytes, int key_unsigned, int64_t n, uint64_t set_sorted, int64_t set_n,
                        uint64_t mask_u8, uint64_t stream);
/* in_pseudorandom_split (petastorm/predicates.py:144-182): bucket = int(md5(str(v)).hexdigest(),16) % sys.maxsize,
 * keep iff lo <= bucket < hi.  Keys are the decimal string of an integer column (computed on device). */
int pst_mask_md5_split_i64(uint64_t keys, int key_bytes, int key_unsigned, int64_t n, double bucket_lo,
                           double bucket_hi, uint64_t mask_u8, uint64_t stream);
/* mask -> ascending int64 row indices; count written to d_count (int64).  d_tmp needs pst_compact_tmp_bytes(n). */
int64_t pst_compact_tmp_bytes(int64_t n);
int pst_mask_compact(uint64_t mask_u8, int64_t n, uint64_t out_idx_i64, uint64_t d_count, uint64_t d_tmp,
                     uint64_t stream);

/* K12: TransformSpec normalise: out = ((float)x - mean) / std, cast to out dtype (petastorm/transform.py:27-57 with
 * the fixed arithmetic order of SURVEY 8(a11)).  dtype codes: 0=u8 1=f16 2=f32 3=i32 4=i16 5=u16 6=f64 */
int pst_normalize(uint64_t src, int src_dtype, int64_t n, float mean, float stddev, uint64_t dst, int dst_dtype,
                  uint64_t stream);

/* K15: NGram window validity (petastorm/ngram.py:225-270).  ts: int64[n] sorted timestamps of one row-group.
 * start_ok[i]=1 iff rows i..i+length-1 exist and every consecutive gap <= delta_threshold; status[0]=1 if ts is not
 * sorted inside any candidate window.  The non-overlap rule (timestamp_overlap=False) is sequential and applied on the
 * compacted candidate list by pst_ngram_no_overlap. */
int pst_ngram_valid_starts(uint64_t ts_i64, int64_t n, int length, int64_t delta_threshold, uint64_t start_ok_u8,
                           uint64_t d_status, uint64_t stream);
/* window gather: out[w, t, :] = src[start[w] + t, :] for t < length; rows are row_bytes wide */
int pst_ngram_gather(uint64_t src, uint64_t starts_i64, int64_t n_windows, int length, int64_t row_bytes, uint64_t dst,
                     uint64_t stream);

/* K16: dtype sanitise of petastorm.pytorch._sanitize_pytorch_types (petastorm/pytorch.py:40-70):
 * kind 0: uint16->int32, 1: uint32->int64, 2: bool->uint8 (copy) */
int pst_sanitize(uint64_t src, int64_t n, int kind, uint64_t dst, uint64_t stream);

/* list<primitive> columns: verifies on-device that the level entries of a repeated column form n_rows lists of exactly
 * list_len fully-defined elements -- the only shape np.vstack(list_of_lists) accepts
 * (petastorm/arrow_reader_worker.py:68-76).  d_flags: int64[1], set to 1 when not uniform. */
int pst_list_uniform(uint64_t rep_u8, uint64_t def_u8, int64_t n, int max_def, int64_t list_len, uint64_t d_flags,
                     uint64_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PST_B200_H_ */
