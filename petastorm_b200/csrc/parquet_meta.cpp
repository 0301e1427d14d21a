#include "parquet_meta.h"

#include <sstream>

#include "thrift_compact.h"

namespace pst {
namespace {

using R = CompactReader;

template <typename F>
void for_struct(R &r, F &&on_field) {
    int16_t last = 0;
    for (;;) {
        R::Field f = r.field(last);
        if (f.type == CT_STOP) return;
        if (!on_field(f)) r.skip(f.type);
    }
}

void parse_logical_type(R &r, SchemaElement &se) {
    // union LogicalType: exactly one field set; the field id identifies the kind.
    for_struct(r, [&](R::Field f) {
        se.logical_kind = f.id;
        if (f.type != CT_STRUCT) return false;
        if (f.id == 5) {  // DECIMAL {1: scale, 2: precision}
            for_struct(r, [&](R::Field g) {
                if (g.id == 1 && g.type == CT_I32) { se.scale = (int32_t)r.zigzag(); return true; }
                if (g.id == 2 && g.type == CT_I32) { se.precision = (int32_t)r.zigzag(); return true; }
                return false;
            });
            return true;
        }
        if (f.id == 7 || f.id == 8) {  // TIME / TIMESTAMP {1: isAdjustedToUTC, 2: unit(union)}
            for_struct(r, [&](R::Field g) {
                if (g.id == 1 && (g.type == CT_TRUE || g.type == CT_FALSE)) { se.logical_utc = g.type == CT_TRUE; return true; }
                if (g.id == 2 && g.type == CT_STRUCT) {
                    for_struct(r, [&](R::Field u) { se.logical_unit = u.id; return false; });
                    return true;
                }
                return false;
            });
            return true;
        }
        if (f.id == 10) {  // INTEGER {1: bitWidth(i8), 2: isSigned}
            for_struct(r, [&](R::Field g) {
                if (g.id == 1 && g.type == CT_BYTE) { se.int_bits = (int8_t)r.byte(); return true; }
                if (g.id == 2 && (g.type == CT_TRUE || g.type == CT_FALSE)) { se.int_signed = g.type == CT_TRUE; return true; }
                return false;
            });
            return true;
        }
        return false;
    });
}

void parse_schema_element(R &r, SchemaElement &se) {
    for_struct(r, [&](R::Field f) {
        switch (f.id) {
            case 1: if (f.type == CT_I32) { se.type = (int32_t)r.zigzag(); return true; } break;
            case 2: if (f.type == CT_I32) { se.type_length = (int32_t)r.zigzag(); return true; } break;
            case 3: if (f.type == CT_I32) { se.repetition = (int32_t)r.zigzag(); return true; } break;
            case 4: if (f.type == CT_BINARY) { se.name = r.binary(); return true; } break;
            case 5: if (f.type == CT_I32) { se.num_children = (int32_t)r.zigzag(); return true; } break;
            case 6: if (f.type == CT_I32) { se.converted_type = (int32_t)r.zigzag(); return true; } break;
            case 7: if (f.type == CT_I32) { se.scale = (int32_t)r.zigzag(); return true; } break;
            case 8: if (f.type == CT_I32) { se.precision = (int32_t)r.zigzag(); return true; } break;
            case 10: if (f.type == CT_STRUCT) { parse_logical_type(r, se); return true; } break;
            default: break;
        }
        return false;
    });
}

void parse_statistics(R &r, ColumnChunkMeta &c) {
    for_struct(r, [&](R::Field f) {
        if (f.id == 3 && f.type == CT_I64) { c.null_count = r.zigzag(); c.has_null_count = true; return true; }
        return false;
    });
}

void parse_column_meta(R &r, ColumnChunkMeta &c) {
    for_struct(r, [&](R::Field f) {
        switch (f.id) {
            case 1: if (f.type == CT_I32) { c.type = (int32_t)r.zigzag(); return true; } break;
            case 2:
                if (f.type == CT_LIST) {
                    R::ListHeader h = r.list();
                    for (uint32_t i = 0; i < h.size; i++) c.encodings.push_back((int32_t)r.zigzag());
                    return true;
                }
                break;
            case 4: if (f.type == CT_I32) { c.codec = (int32_t)r.zigzag(); return true; } break;
            case 5: if (f.type == CT_I64) { c.num_values = r.zigzag(); return true; } break;
            case 6: if (f.type == CT_I64) { c.total_uncompressed_size = r.zigzag(); return true; } break;
            case 7: if (f.type == CT_I64) { c.total_compressed_size = r.zigzag(); return true; } break;
            case 9: if (f.type == CT_I64) { c.data_page_offset = r.zigzag(); return true; } break;
            case 11: if (f.type == CT_I64) { c.dictionary_page_offset = r.zigzag(); return true; } break;
            case 12: if (f.type == CT_STRUCT) { parse_statistics(r, c); return true; } break;
            default: break;
        }
        return false;
    });
}

void parse_column_chunk(R &r, ColumnChunkMeta &c) {
    for_struct(r, [&](R::Field f) {
        if (f.id == 1 && f.type == CT_BINARY) { c.file_path = r.binary(); return true; }
        if (f.id == 2 && f.type == CT_I64) { c.file_offset = r.zigzag(); return true; }
        if (f.id == 3 && f.type == CT_STRUCT) { parse_column_meta(r, c); return true; }
        return false;
    });
}

void parse_row_group(R &r, RowGroupMeta &g) {
    for_struct(r, [&](R::Field f) {
        if (f.id == 1 && f.type == CT_LIST) {
            R::ListHeader h = r.list();
            g.columns.resize(h.size);
            for (uint32_t i = 0; i < h.size; i++) parse_column_chunk(r, g.columns[i]);
            return true;
        }
        if (f.id == 2 && f.type == CT_I64) { g.total_byte_size = r.zigzag(); return true; }
        if (f.id == 3 && f.type == CT_I64) { g.num_rows = r.zigzag(); return true; }
        return false;
    });
}

// Depth-first walk assigning definition / repetition levels to the leaves.
void build_leaves(FileMeta &m) {
    m.leaves.clear();
    if (m.schema.empty()) return;
    struct Frame { int remaining; int def; int rep; };
    std::vector<Frame> stack;
    std::vector<std::string> path;
    stack.push_back({m.schema[0].num_children, 0, 0});
    int top = -1;
    for (size_t i = 1; i < m.schema.size(); i++) {
        while (!stack.empty() && stack.back().remaining == 0) {
            stack.pop_back();
            if (!path.empty()) path.pop_back();
        }
        if (stack.empty()) throw std::runtime_error("parquet schema tree is malformed");
        stack.back().remaining--;
        if (stack.size() == 1) top++;
        const SchemaElement &se = m.schema[i];
        int def = stack.back().def + (se.repetition != 0 ? 1 : 0);
        int rep = stack.back().rep + (se.repetition == 2 ? 1 : 0);
        path.push_back(se.name);
        if (se.num_children > 0 && se.type < 0) {
            stack.push_back({se.num_children, def, rep});
        } else {
            LeafColumn lc;
            lc.schema_index = (int)i;
            lc.path = path;
            lc.max_def = def;
            lc.max_rep = rep;
            lc.top_index = top;
            m.leaves.push_back(lc);
            path.pop_back();
        }
    }
}

void json_escape(std::ostringstream &os, const std::string &s) {
    os << '"';
    for (unsigned char ch : s) {
        switch (ch) {
            case '"': os << "\\\""; break;
            case '\\': os << "\\\\"; break;
            case '\n': os << "\\n"; break;
            case '\r': os << "\\r"; break;
            case '\t': os << "\\t"; break;
            default:
                if (ch < 0x20) {
                    char buf[8];
                    snprintf(buf, sizeof buf, "\\u%04x", ch);
                    os << buf;
                } else {
                    os << ch;
                }
        }
    }
    os << '"';
}

}  // namespace

void parse_file_meta(const uint8_t *p, size_t n, FileMeta &out) {
    R r(p, n);
    for_struct(r, [&](R::Field f) {
        switch (f.id) {
            case 1: if (f.type == CT_I32) { out.version = (int32_t)r.zigzag(); return true; } break;
            case 2:
                if (f.type == CT_LIST) {
                    R::ListHeader h = r.list();
                    out.schema.resize(h.size);
                    for (uint32_t i = 0; i < h.size; i++) parse_schema_element(r, out.schema[i]);
                    return true;
                }
                break;
            case 3: if (f.type == CT_I64) { out.num_rows = r.zigzag(); return true; } break;
            case 4:
                if (f.type == CT_LIST) {
                    R::ListHeader h = r.list();
                    out.row_groups.resize(h.size);
                    for (uint32_t i = 0; i < h.size; i++) parse_row_group(r, out.row_groups[i]);
                    return true;
                }
                break;
            case 5:
                if (f.type == CT_LIST) {
                    R::ListHeader h = r.list();
                    for (uint32_t i = 0; i < h.size; i++) {
                        std::string k, v;
                        for_struct(r, [&](R::Field g) {
                            if (g.id == 1 && g.type == CT_BINARY) { k = r.binary(); return true; }
                            if (g.id == 2 && g.type == CT_BINARY) { v = r.binary(); return true; }
                            return false;
                        });
                        out.kv.emplace_back(std::move(k), std::move(v));
                    }
                    return true;
                }
                break;
            case 6: if (f.type == CT_BINARY) { out.created_by = r.binary(); return true; } break;
            default: break;
        }
        return false;
    });
    build_leaves(out);
    for (const RowGroupMeta &g : out.row_groups)
        if (g.columns.size() != out.leaves.size())
            throw std::runtime_error("row group column count does not match the schema leaf count");
}

void parse_page_header(const uint8_t *p, size_t n, PageHeader &out) {
    R r(p, n);
    out = PageHeader();
    for_struct(r, [&](R::Field f) {
        switch (f.id) {
            case 1: if (f.type == CT_I32) { out.type = (int32_t)r.zigzag(); return true; } break;
            case 2: if (f.type == CT_I32) { out.uncompressed_page_size = (int32_t)r.zigzag(); return true; } break;
            case 3: if (f.type == CT_I32) { out.compressed_page_size = (int32_t)r.zigzag(); return true; } break;
            case 5:  // DataPageHeader
                if (f.type == CT_STRUCT) {
                    for_struct(r, [&](R::Field g) {
                        if (g.type != CT_I32) return false;
                        if (g.id == 1) { out.num_values = (int32_t)r.zigzag(); return true; }
                        if (g.id == 2) { out.encoding = (int32_t)r.zigzag(); return true; }
                        if (g.id == 3) { out.def_encoding = (int32_t)r.zigzag(); return true; }
                        if (g.id == 4) { out.rep_encoding = (int32_t)r.zigzag(); return true; }
                        return false;
                    });
                    return true;
                }
                break;
            case 7:  // DictionaryPageHeader
                if (f.type == CT_STRUCT) {
                    for_struct(r, [&](R::Field g) {
                        if (g.type != CT_I32) return false;
                        if (g.id == 1) { out.num_values = (int32_t)r.zigzag(); return true; }
                        if (g.id == 2) { out.encoding = (int32_t)r.zigzag(); return true; }
                        return false;
                    });
                    return true;
                }
                break;
            case 8:  // DataPageHeaderV2
                if (f.type == CT_STRUCT) {
                    for_struct(r, [&](R::Field g) {
                        if (g.id == 7 && (g.type == CT_TRUE || g.type == CT_FALSE)) { out.is_compressed = g.type == CT_TRUE; return true; }
                        if (g.type != CT_I32) return false;
                        switch (g.id) {
                            case 1: out.num_values = (int32_t)r.zigzag(); return true;
                            case 2: out.num_nulls = (int32_t)r.zigzag(); return true;
                            case 3: out.num_rows = (int32_t)r.zigzag(); return true;
                            case 4: out.encoding = (int32_t)r.zigzag(); return true;
                            case 5: out.def_bytes = (int32_t)r.zigzag(); return true;
                            case 6: out.rep_bytes = (int32_t)r.zigzag(); return true;
                            default: return false;
                        }
                    });
                    return true;
                }
                break;
            default: break;
        }
        return false;
    });
    out.header_size = r.consumed();
    if (out.type < 0) throw std::runtime_error("page header without a type");
}

std::string schema_json(const FileMeta &m) {
    std::ostringstream os;
    os << "{\"num_rows\":" << m.num_rows << ",\"created_by\":";
    json_escape(os, m.created_by);
    os << ",\"top_level\":[";
    // top-level fields (children of the root), in order
    {
        size_t i = 1;
        bool first = true;
        // walk subtrees
        while (i < m.schema.size()) {
            const SchemaElement &se = m.schema[i];
            if (!first) os << ',';
            first = false;
            os << "{\"name\":";
            json_escape(os, se.name);
            os << ",\"repetition\":" << se.repetition << ",\"num_children\":" << se.num_children
               << ",\"converted_type\":" << se.converted_type << ",\"logical_kind\":" << se.logical_kind << "}";
            // skip the subtree
            size_t todo = (size_t)se.num_children;
            i++;
            while (todo > 0 && i < m.schema.size()) {
                todo += (size_t)m.schema[i].num_children;
                todo--;
                i++;
            }
        }
    }
    os << "],\"leaves\":[";
    for (size_t k = 0; k < m.leaves.size(); k++) {
        const LeafColumn &lc = m.leaves[k];
        const SchemaElement &se = m.schema[lc.schema_index];
        if (k) os << ',';
        os << "{\"index\":" << k << ",\"name\":";
        json_escape(os, lc.path.empty() ? se.name : lc.path[0]);
        os << ",\"path\":[";
        for (size_t j = 0; j < lc.path.size(); j++) {
            if (j) os << ',';
            json_escape(os, lc.path[j]);
        }
        os << "],\"top_index\":" << lc.top_index << ",\"physical_type\":" << se.type
           << ",\"type_length\":" << se.type_length << ",\"repetition\":" << se.repetition
           << ",\"converted_type\":" << se.converted_type << ",\"logical_kind\":" << se.logical_kind
           << ",\"logical_unit\":" << se.logical_unit << ",\"logical_utc\":" << (se.logical_utc ? "true" : "false")
           << ",\"int_bits\":" << se.int_bits << ",\"int_signed\":" << (se.int_signed ? "true" : "false")
           << ",\"scale\":" << se.scale << ",\"precision\":" << se.precision << ",\"max_def\":" << lc.max_def
           << ",\"max_rep\":" << lc.max_rep << "}";
    }
    os << "]}";
    return os.str();
}

}  // namespace pst
