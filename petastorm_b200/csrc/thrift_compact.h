// Minimal Thrift *compact protocol* reader -- just enough to parse parquet.thrift FileMetaData and PageHeader.
// Written against the Apache Thrift compact-protocol specification (thrift/doc/specs/thrift-compact-protocol.md);
// this replaces the thrift deserialisation that Arrow C++ performs inside `pq.ParquetFile(...)`
// (reference call sites: petastorm/arrow_reader_worker.py:172, petastorm/py_dict_reader_worker.py:146).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace pst {

struct ThriftError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum CType : uint8_t {
    CT_STOP = 0, CT_TRUE = 1, CT_FALSE = 2, CT_BYTE = 3, CT_I16 = 4, CT_I32 = 5, CT_I64 = 6, CT_DOUBLE = 7,
    CT_BINARY = 8, CT_LIST = 9, CT_SET = 10, CT_MAP = 11, CT_STRUCT = 12
};

class CompactReader {
public:
    CompactReader(const uint8_t *p, size_t n) : p_(p), end_(p + n), begin_(p) {}

    size_t consumed() const { return static_cast<size_t>(p_ - begin_); }

    uint8_t byte() {
        need(1);
        return *p_++;
    }
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        for (;;) {
            uint8_t b = byte();
            v |= static_cast<uint64_t>(b & 0x7f) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
            if (shift > 63) throw ThriftError("varint too long");
        }
        return v;
    }
    int64_t zigzag() {
        uint64_t v = varint();
        return static_cast<int64_t>((v >> 1) ^ (~(v & 1) + 1));
    }
    // Starts a struct scope: returns previous field id to restore on leave.
    struct Field {
        int16_t id;
        uint8_t type;  // CType; CT_STOP at end of struct
    };
    // Reads the next field header inside a struct. `last_id` is the running field id of this struct scope.
    Field field(int16_t &last_id) {
        uint8_t h = byte();
        if (h == 0) return {0, CT_STOP};
        uint8_t type = h & 0x0f;
        uint8_t delta = h >> 4;
        int16_t id;
        if (delta == 0)
            id = static_cast<int16_t>(zigzag());
        else
            id = static_cast<int16_t>(last_id + delta);
        last_id = id;
        return {id, type};
    }
    struct ListHeader {
        uint32_t size;
        uint8_t type;
    };
    ListHeader list() {
        uint8_t h = byte();
        uint32_t size = h >> 4;
        uint8_t type = h & 0x0f;
        if (size == 15) size = static_cast<uint32_t>(varint());
        return {size, type};
    }
    std::string binary() {
        uint64_t n = varint();
        need(n);
        std::string s(reinterpret_cast<const char *>(p_), n);
        p_ += n;
        return s;
    }
    // zero-copy view of a binary field
    void binary_view(const uint8_t *&ptr, size_t &len) {
        uint64_t n = varint();
        need(n);
        ptr = p_;
        len = n;
        p_ += n;
    }
    bool bool_value(uint8_t field_type) { return field_type == CT_TRUE; }
    // bool element inside a list: one byte (1 = true, anything else false)
    bool list_bool() { return byte() == 1; }

    void skip(uint8_t type) {
        switch (type) {
            case CT_TRUE:
            case CT_FALSE:
                return;
            case CT_BYTE:
                byte();
                return;
            case CT_I16:
            case CT_I32:
            case CT_I64:
                varint();
                return;
            case CT_DOUBLE:
                need(8);
                p_ += 8;
                return;
            case CT_BINARY: {
                uint64_t n = varint();
                need(n);
                p_ += n;
                return;
            }
            case CT_LIST:
            case CT_SET: {
                ListHeader h = list();
                for (uint32_t i = 0; i < h.size; i++) skip_element(h.type);
                return;
            }
            case CT_MAP: {
                uint64_t n = varint();
                if (n == 0) return;
                uint8_t kv = byte();
                for (uint64_t i = 0; i < n; i++) {
                    skip_element(kv >> 4);
                    skip_element(kv & 0x0f);
                }
                return;
            }
            case CT_STRUCT: {
                int16_t last = 0;
                for (;;) {
                    Field f = field(last);
                    if (f.type == CT_STOP) break;
                    skip(f.type);
                }
                return;
            }
            default:
                throw ThriftError("unknown thrift compact type " + std::to_string(type));
        }
    }

private:
    // element inside a collection: booleans occupy one byte there
    void skip_element(uint8_t type) {
        if (type == CT_TRUE || type == CT_FALSE)
            byte();
        else
            skip(type);
    }
    void need(uint64_t n) {
        if (static_cast<uint64_t>(end_ - p_) < n) throw ThriftError("thrift: truncated input");
    }
    const uint8_t *p_;
    const uint8_t *end_;
    const uint8_t *begin_;
};

}  // namespace pst
