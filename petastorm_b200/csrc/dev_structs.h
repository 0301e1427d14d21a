// POD tables shared between the host planner (plan.cpp) and the device kernels (kernels.cu).
// They are written by the host into the tail of the plan's `raw` region and read by every decode kernel.
#pragma once
#include <cstdint>

namespace pst {

enum PageKind : uint8_t { PK_DATA_V1 = 0, PK_DICT = 2, PK_DATA_V2 = 3 };

// parquet.thrift Encoding
enum Enc : uint8_t {
    ENC_PLAIN = 0, ENC_PLAIN_DICTIONARY = 2, ENC_RLE = 3, ENC_BIT_PACKED = 4, ENC_DELTA_BINARY_PACKED = 5,
    ENC_DELTA_LENGTH_BYTE_ARRAY = 6, ENC_DELTA_BYTE_ARRAY = 7, ENC_RLE_DICTIONARY = 8, ENC_BYTE_STREAM_SPLIT = 9
};

// error codes written to d_status[0]
enum DevErr : int32_t {
    DE_OK = 0, DE_SNAPPY_CORRUPT = 1, DE_LEVELS_CORRUPT = 2, DE_UNSUPPORTED_ENCODING = 3, DE_DICT_INDEX_RANGE = 4,
    DE_PAGE_OVERRUN = 5, DE_NPY_HEADER_MISMATCH = 6, DE_PNG_CORRUPT = 7, DE_PNG_UNSUPPORTED = 8,
    DE_NGRAM_UNSORTED = 9, DE_BYTE_ARRAY_CORRUPT = 10, DE_GZIP_CORRUPT = 11,
    DE_ZIP_CORRUPT = 12
};

constexpr int32_t kSnappyFragment = 65536;

// DevPage::flags
enum PageFlags : uint8_t {
    PF_ALL_VALID = 1,       // the planner read the definition levels: one RLE run of max_def covering the page
    PF_COPY = 2,            // PLAIN fixed-width flat all-valid page: its values go to `out` through k_copy_tiles
    PF_UNWRAPPED = 4,       // stored as literal-only Snappy; the staging copy dropped the framing (device sees NONE)
};

// One work item of the tile copy kernel: nbytes of value bytes arena[src_off..] -> out[dst_off..], plus `nvalid`
// validity bytes (value 1) at out[valid_off..] when the column carries a validity array (valid_off >= 0).
struct CopyTile {
    int64_t src_off;
    int64_t dst_off;
    int64_t valid_off;
    int32_t nbytes;
    int32_t nvalid;
};
static_assert(sizeof(CopyTile) == 32, "CopyTile must be 32 bytes");
constexpr int32_t kCopyTileBytes = 65536;

struct SnFrag {             // one work item of the fragment decode kernel
    int32_t page;           // page table index
    int32_t k;              // fragment ordinal inside the page
};

struct DevPage {            // 64 bytes
    int64_t src_off;        // arena offset of the payload as stored in the file
    int64_t img_off;        // arena offset of the uncompressed page image (== src_off for uncompressed pages)
    int32_t comp_size;      // stored payload bytes
    int32_t uncomp_size;    // uncompressed page image bytes
    int32_t num_values;     // level entries in this page (incl. nulls)
    int32_t first_value;    // index of this page's first level entry inside the column chunk
    int32_t def_bytes;      // V2: byte length of the definition levels; V1: -1 (length prefix is in the stream)
    int32_t rep_bytes;      // V2: byte length of the repetition levels; V1: -1
    int16_t col;            // plan column slot
    uint8_t kind;           // PageKind
    uint8_t encoding;       // Enc of the values
    uint8_t codec;          // 0 none, 1 snappy, 2 gzip
    uint8_t def_enc;        // V1: Enc of definition levels
    uint8_t rep_enc;        // V1: Enc of repetition levels
    uint8_t v2_compressed;  // V2: values section compressed?
    int16_t page_ordinal;   // ordinal in file order inside the chunk (diagnostics, saturates)
    uint8_t flags;          // PageFlags (decided by the host planner)
    uint8_t pad_;
    // Snappy pages: the compressed values are decoded as `nfrag` independent fragments of kSnappyFragment output bytes
    // (the Snappy compressor restarts its match window every 64 KiB, see kernels_decode.cu)
    int32_t frag_first;     // index of this page's first entry in the plan's fragment-position table (nfrag + 1 entries)
    int32_t nfrag;          // fragments of this page (>= 1 for a compressed page, 0 otherwise)
    int32_t multi_slot;     // index into the plan's multi-fragment page list / flag array, -1 when nfrag <= 1
};
static_assert(sizeof(DevPage) == 64, "DevPage must be 64 bytes");

struct BaDictEntry {        // one entry of a BYTE_ARRAY dictionary: where its bytes lie in the arena
    int64_t off;
    int32_t len;
    int32_t pad;
};
static_assert(sizeof(BaDictEntry) == 16, "BaDictEntry must be 16 bytes");

struct DevCol {             // 96 bytes
    int32_t ptype;          // parquet physical type
    int32_t width;          // bytes per decoded value (BOOLEAN 1, INT96 12, FLBA n); BYTE_ARRAY: 0
    int32_t max_def;
    int32_t max_rep;
    int64_t num_values;     // level entries in the chunk
    int64_t values_off;     // out region offsets (see pst_plan_column)
    int64_t valid_off;
    int64_t rep_off;
    int64_t def_off;
    int64_t lens_off;       // BYTE_ARRAY: int32 lens[num_values] (values_off holds int64 arena offsets)
    int64_t dict_img_off;   // arena offset of the uncompressed dictionary page image, -1 if none
    int32_t dict_count;     // entries in the dictionary
    int32_t dict_page;      // page table index of the dictionary page, -1 if none
    int64_t dict_index_off; // BYTE_ARRAY dictionaries: arena offset of {int64 off, int32 len, int32 pad}[dict_count]
    int64_t pad_;
};
static_assert(sizeof(DevCol) == 96, "DevCol must be 96 bytes");

}  // namespace pst
