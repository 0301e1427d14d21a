// Internal host-side objects behind the opaque C handles.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "dev_structs.h"
#include "parquet_meta.h"

struct pst_file {
    std::string path;
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t size = 0;
    int64_t mtime_ns = 0;
    pst::FileMeta meta;
    std::string schema_json;
};

struct HostPage {
    pst::DevPage d;          // what the device sees
    int64_t file_off = 0;    // payload position in the file
    int32_t stored_size = 0; // payload bytes as stored in the file
    int64_t copy_v0 = -1;    // >= 0: PF_COPY page, byte offset of the value section inside the page image
    // literal-only Snappy page (PF_UNWRAPPED): {file offset, length} of the byte ranges that make up the page image
    std::vector<std::pair<int64_t, int32_t>> segs;
};

struct pst_plan {
    const pst_file *file = nullptr;
    int rg = 0;
    std::vector<int> cols;
    std::vector<pst::DevCol> dcols;
    std::vector<HostPage> pages;
    std::vector<int32_t> compressed_pages;   // page indices needing decompression
    std::vector<int32_t> data_pages;         // page indices of data pages
    std::vector<int32_t> ba_dict_pages;      // BYTE_ARRAY dictionary pages needing an entry index
    std::vector<pst::SnFrag> snappy_frags;   // (page, fragment) work items of the Snappy fragment kernel
    std::vector<int32_t> multi_pages;        // compressed pages with more than one fragment (indexed first)
    std::vector<int32_t> gzip_pages;         // page indices compressed with GZIP
    std::vector<int32_t> index_pages;        // multi-fragment pages whose fragment positions the device has to find,
                                             // longest first
    int32_t index_big_count = 0;             // the first index_big_count of them go to the cluster variant of the kernel
    // BYTE_ARRAY dictionaries the planner indexed itself (pages the device sees uncompressed): entries per plan column
    std::vector<std::pair<int, std::vector<pst::BaDictEntry>>> host_dict_index;
    std::vector<pst::CopyTile> copy_tiles;   // work items of k_copy_tiles (PF_COPY pages, <= 64 KiB each)
    int64_t unwrapped_pages = 0;             // literal-only Snappy pages delivered as uncompressed images
    int64_t host_indexed_pages = 0;          // multi-fragment Snappy pages whose fragment positions the planner found
    int64_t copy_tiles_off = 0;
    std::vector<uint32_t> frag_pos_host;     // fragment-position table as far as the host knows it (literal-only pages)
    int64_t frag_pos_count = 0;              // entries of the fragment-position table (sum of nfrag + 1)
    int64_t num_rows = 0;
    int64_t payload_bytes = 0;
    int64_t uncompressed_bytes = 0;
    // arena layout
    int64_t tables_off = 0;      // offset of the tables inside the raw region
    int64_t cols_off = 0, pages_off = 0, comp_list_off = 0, data_list_off = 0, dict_list_off = 0;
    int64_t frag_list_off = 0, multi_list_off = 0, gzip_list_off = 0, index_list_off = 0;   // tables (raw region)
    int64_t frag_pos_off = 0, page_flag_off = 0;     // raw region too, completed / raised by the device
    int64_t raw_bytes = 0;
    int64_t scratch_off = 0;     // == align(raw_bytes)
    int64_t arena_bytes = 0;
    int64_t out_bytes = 0;
    std::vector<uint8_t> tables; // host image of the tables (copied to raw[tables_off:])
    uint64_t cache_key = 0;
};

namespace pst {
void set_error(const std::string &msg);
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
// Decompress at most `want` leading bytes of a raw-snappy stream (host side peek). Returns bytes produced.
size_t snappy_peek(const uint8_t *src, size_t n, uint8_t *out, size_t want);
}  // namespace pst
