// Host-side Parquet metadata model: FileMetaData (footer) and PageHeader, parsed from thrift-compact bytes.
// Follows the public parquet-format specification (parquet.thrift); replaces the footer/page-header handling of
// Arrow C++ that the reference reaches through pq.ParquetFile / piece.read
// (petastorm/arrow_reader_worker.py:172,358; petastorm/py_dict_reader_worker.py:146,267;
//  petastorm/etl/dataset_metadata.py:340-353).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace pst {

struct SchemaElement {
    std::string name;
    int32_t type = -1;            // physical type, -1 for groups
    int32_t type_length = 0;
    int32_t repetition = 0;       // 0 required, 1 optional, 2 repeated
    int32_t num_children = 0;
    int32_t converted_type = -1;
    int32_t scale = 0, precision = 0;
    // logical type (union): kind = field id in LogicalType (1 STRING, 3 LIST, 5 DECIMAL, 6 DATE, 7 TIME, 8 TIMESTAMP,
    // 10 INTEGER, ...), 0 if absent
    int32_t logical_kind = 0;
    int32_t logical_unit = 0;     // TIME/TIMESTAMP: 1 millis, 2 micros, 3 nanos
    bool logical_utc = false;
    int32_t int_bits = 0;         // INTEGER
    bool int_signed = true;
};

struct LeafColumn {
    int schema_index = 0;                 // index into FileMeta::schema
    std::vector<std::string> path;        // path_in_schema
    int max_def = 0, max_rep = 0;
    int top_index = 0;                    // index of the top-level field this leaf belongs to
};

struct ColumnChunkMeta {
    int32_t type = 0;
    int32_t codec = 0;
    int64_t num_values = 0;
    int64_t total_uncompressed_size = 0;
    int64_t total_compressed_size = 0;
    int64_t data_page_offset = 0;
    int64_t dictionary_page_offset = -1;
    int64_t file_offset = 0;
    std::vector<int32_t> encodings;
    std::string file_path;
    bool has_null_count = false;
    int64_t null_count = 0;
    int64_t start_offset() const {
        // parquet-mr before 1.10 wrote dictionary_page_offset = 0 for "absent"
        if (dictionary_page_offset > 0 && dictionary_page_offset < data_page_offset) return dictionary_page_offset;
        return data_page_offset;
    }
};

struct RowGroupMeta {
    std::vector<ColumnChunkMeta> columns;
    int64_t total_byte_size = 0;
    int64_t num_rows = 0;
};

struct FileMeta {
    int32_t version = 0;
    int64_t num_rows = 0;
    std::vector<SchemaElement> schema;
    std::vector<RowGroupMeta> row_groups;
    std::vector<std::pair<std::string, std::string>> kv;
    std::string created_by;
    std::vector<LeafColumn> leaves;
};

struct PageHeader {
    int32_t type = -1;  // 0 DATA_PAGE, 1 INDEX_PAGE, 2 DICTIONARY_PAGE, 3 DATA_PAGE_V2
    int32_t uncompressed_page_size = 0;
    int32_t compressed_page_size = 0;
    int32_t num_values = 0;
    int32_t encoding = 0;
    int32_t def_encoding = 3;
    int32_t rep_encoding = 3;
    // v2
    int32_t num_nulls = 0, num_rows = 0;
    int32_t def_bytes = 0, rep_bytes = 0;
    bool is_compressed = true;
    size_t header_size = 0;  // bytes consumed by the header itself
};

// Parses the thrift FileMetaData in [p, p+n). Throws std::runtime_error on malformed input.
void parse_file_meta(const uint8_t *p, size_t n, FileMeta &out);
// Parses one PageHeader starting at p (at most n bytes available).
void parse_page_header(const uint8_t *p, size_t n, PageHeader &out);
// JSON rendering of the leaf columns (consumed by the Python host to build a Unischema).
std::string schema_json(const FileMeta &m);

}  // namespace pst
