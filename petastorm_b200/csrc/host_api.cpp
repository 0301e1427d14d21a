// Host-only half of the C-ABI: file mapping, footer, row-group plan (page walk + HBM layout). No CUDA in here.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <cstring>
#include <stdexcept>

#include "../../include/pst_b200.h"
#include "host_state.h"

namespace pst {
static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

// --------------------------------------------------------------------------------------------------------------
// Snappy "peek": decode only the first `want` bytes of a raw snappy block (format_description.txt of google/snappy).
// Used by the planner to learn the byte length of the level sections of a compressed V1 data page so that the
// value section can be placed 16-byte aligned in HBM.
// --------------------------------------------------------------------------------------------------------------
size_t snappy_peek(const uint8_t *src, size_t n, uint8_t *out, size_t want) {
    size_t ip = 0;
    // preamble: uncompressed length varint
    uint64_t ulen = 0;
    int shift = 0;
    for (;;) {
        if (ip >= n) return 0;
        uint8_t b = src[ip++];
        ulen |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
        if (shift > 35) return 0;
    }
    if (want > ulen) want = (size_t)ulen;
    size_t op = 0;
    while (op < want && ip < n) {
        uint8_t tag = src[ip++];
        size_t len, offset = 0;
        switch (tag & 3) {
            case 0: {
                len = (tag >> 2) + 1;
                if (len > 60) {
                    size_t nb = len - 60;
                    if (ip + nb > n) return op;
                    len = 0;
                    for (size_t i = 0; i < nb; i++) len |= (size_t)src[ip + i] << (8 * i);
                    len += 1;
                    ip += nb;
                }
                size_t take = std::min(len, want - op);
                if (ip + take > n) return op;
                memcpy(out + op, src + ip, take);
                op += take;
                ip += len;
                continue;
            }
            case 1:
                if (ip >= n) return op;
                len = ((tag >> 2) & 7) + 4;
                offset = ((size_t)(tag >> 5) << 8) | src[ip++];
                break;
            case 2:
                if (ip + 2 > n) return op;
                len = (tag >> 2) + 1;
                offset = src[ip] | ((size_t)src[ip + 1] << 8);
                ip += 2;
                break;
            default:
                if (ip + 4 > n) return op;
                len = (tag >> 2) + 1;
                offset = src[ip] | ((size_t)src[ip + 1] << 8) | ((size_t)src[ip + 2] << 16) | ((size_t)src[ip + 3] << 24);
                ip += 4;
                break;
        }
        if (offset == 0 || offset > op) return op;
        for (size_t i = 0; i < len && op < want; i++, op++) out[op] = out[op - offset];
    }
    return op;
}
}  // namespace pst

using namespace pst;

#define PST_TRY try {
#define PST_CATCH(ret)                      \
    }                                       \
    catch (const std::exception &e) {       \
        pst::set_error(e.what());           \
        return ret;                         \
    }

extern "C" {

const char *pst_last_error(void) { return pst::g_last_error.c_str(); }
int pst_abi_version(void) { return PST_ABI_VERSION; }

int pst_file_open(const char *path, pst_file **out) {
    PST_TRY
    *out = nullptr;
    int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) throw std::runtime_error(std::string("cannot open ") + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) {
        ::close(fd);
        throw std::runtime_error(std::string("fstat failed for ") + path);
    }
    size_t size = (size_t)st.st_size;
    if (size < 12) {
        ::close(fd);
        throw std::runtime_error(std::string(path) + " is too small to be a parquet file");
    }
    void *map = mmap(nullptr, size, PROT_READ, MAP_SHARED, fd, 0);
    const int mmap_errno = errno;
    ::close(fd);   // the mapping outlives the descriptor: an open file costs no fd (datasets have 10^4..10^5 files)
    fd = -1;
    if (map == MAP_FAILED)
        throw std::runtime_error(std::string("mmap failed for ") + path + ": " + strerror(mmap_errno));
    std::unique_ptr<pst_file> f(new pst_file());
    f->path = path;
    f->fd = -1;
    f->map = static_cast<const uint8_t *>(map);
    f->size = size;
    f->mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    auto fail = [&](const std::string &m) {
        munmap(map, size);
        f->map = nullptr;
        throw std::runtime_error(std::string(path) + ": " + m);
    };
    if (memcmp(f->map + size - 4, "PAR1", 4) != 0 || memcmp(f->map, "PAR1", 4) != 0) {
        if (memcmp(f->map + size - 4, "PARE", 4) == 0) fail("encrypted parquet footers are not supported");
        fail("missing PAR1 magic");
    }
    uint32_t flen;
    memcpy(&flen, f->map + size - 8, 4);
    if ((size_t)flen + 12 > size) fail("footer length exceeds the file size");
    try {
        parse_file_meta(f->map + size - 8 - flen, flen, f->meta);
    } catch (const std::exception &e) {
        fail(std::string("footer parse failed: ") + e.what());
    }
    f->schema_json = schema_json(f->meta);
    *out = f.release();
    return 0;
    PST_CATCH(1)
}

void pst_file_close(pst_file *f) {
    if (!f) return;
    if (f->map) munmap(const_cast<uint8_t *>(f->map), f->size);
    if (f->fd >= 0) ::close(f->fd);
    delete f;
}

int pst_file_num_row_groups(const pst_file *f) { return (int)f->meta.row_groups.size(); }
int64_t pst_file_num_rows(const pst_file *f) { return f->meta.num_rows; }
int64_t pst_file_row_group_num_rows(const pst_file *f, int rg) {
    if (rg < 0 || rg >= (int)f->meta.row_groups.size()) return -1;
    return f->meta.row_groups[rg].num_rows;
}
int pst_file_num_columns(const pst_file *f) { return (int)f->meta.leaves.size(); }

int pst_file_schema_json(const pst_file *f, const char **json, size_t *len) {
    *json = f->schema_json.c_str();
    *len = f->schema_json.size();
    return 0;
}

int pst_file_num_kv(const pst_file *f) { return (int)f->meta.kv.size(); }
int pst_file_kv_at(const pst_file *f, int i, const char **key, size_t *klen, const uint8_t **val, size_t *vlen) {
    if (i < 0 || i >= (int)f->meta.kv.size()) {
        set_error("kv index out of range");
        return 1;
    }
    *key = f->meta.kv[i].first.data();
    *klen = f->meta.kv[i].first.size();
    *val = reinterpret_cast<const uint8_t *>(f->meta.kv[i].second.data());
    *vlen = f->meta.kv[i].second.size();
    return 0;
}
int pst_file_kv_metadata(const pst_file *f, const char *key, const uint8_t **val, size_t *len) {
    for (const auto &kv : f->meta.kv)
        if (kv.first == key) {
            *val = reinterpret_cast<const uint8_t *>(kv.second.data());
            *len = kv.second.size();
            return 0;
        }
    *val = nullptr;
    *len = 0;
    return 1;
}

int pst_file_chunk_info(const pst_file *f, int rg, int col, pst_chunk_info *out) {
    if (rg < 0 || rg >= (int)f->meta.row_groups.size() || col < 0 || col >= (int)f->meta.leaves.size()) {
        set_error("row group / column out of range");
        return 1;
    }
    const ColumnChunkMeta &c = f->meta.row_groups[rg].columns[col];
    out->physical_type = c.type;
    out->codec = c.codec;
    out->num_values = c.num_values;
    out->data_page_offset = c.data_page_offset;
    out->dictionary_page_offset = c.dictionary_page_offset > 0 ? c.dictionary_page_offset : -1;
    out->total_compressed_size = c.total_compressed_size;
    out->total_uncompressed_size = c.total_uncompressed_size;
    out->start_offset = c.start_offset();
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------------------------
static int level_bits(int max_level) {
    int b = 0;
    while ((1 << b) <= max_level) b++;
    return max_level == 0 ? 0 : b;
}

// What the planner learns from the first bytes of a data page image: the byte offset of the value section (-1 when it
// cannot be learned) and whether the definition levels are one RLE run of max_def that covers the page (no nulls).
struct LevelPeek {
    int64_t values_off = -1;
    bool all_valid = false;
};

// One RLE-hybrid section [p, p+len): true iff it is a single RLE run of `value` covering >= num_values entries
static bool single_rle_run(const uint8_t *p, size_t len, size_t have, int bit_width, uint32_t value, int64_t num_values) {
    if (len > have) len = have;
    size_t q = 0;
    uint64_t h = 0;
    for (int shift = 0;; shift += 7) {
        if (q >= len || shift > 35) return false;
        const uint8_t b = p[q++];
        h |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
    }
    if (h & 1) return false;
    if ((int64_t)(h >> 1) < num_values) return false;
    const int nb = (bit_width + 7) >> 3;
    if (q + (size_t)nb > len) return false;
    uint32_t v = 0;
    for (int i = 0; i < nb; i++) v |= (uint32_t)p[q + i] << (8 * i);
    return v == value;
}

static LevelPeek v1_level_peek(const uint8_t *payload, size_t n, int codec, int max_rep, int max_def, int rep_enc,
                               int def_enc, int num_values) {
    LevelPeek r;
    if (max_rep == 0 && max_def == 0) {
        r.values_off = 0;
        r.all_valid = true;
        return r;
    }
    uint8_t head[512];
    size_t have;
    if (codec == PST_CODEC_NONE) {
        have = std::min(n, sizeof head);
        memcpy(head, payload, have);
    } else if (codec == PST_CODEC_SNAPPY) {
        have = snappy_peek(payload, n, head, sizeof head);
    } else {
        return r;
    }
    size_t pos = 0;
    bool def_single = false;
    auto section = [&](int max_level, int enc, bool is_def) -> bool {
        if (max_level == 0) return true;
        if (enc == ENC_RLE) {
            if (pos + 4 > have) return false;
            uint32_t len;
            memcpy(&len, head + pos, 4);
            if (is_def && pos + 4 < have)
                def_single = single_rle_run(head + pos + 4, len, have - (pos + 4), level_bits(max_level),
                                            (uint32_t)max_level, num_values);
            pos += 4 + (size_t)len;
            return true;
        }
        if (enc == ENC_BIT_PACKED) {
            pos += ((size_t)num_values * level_bits(max_level) + 7) / 8;
            return true;
        }
        return false;
    };
    if (!section(max_rep, rep_enc, false)) return r;
    if (pos > have && max_def > 0 && def_enc == ENC_RLE) return r;  // cannot see the def length prefix
    if (!section(max_def, def_enc, true)) return r;
    r.values_off = (int64_t)pos;
    r.all_valid = max_def == 0 || def_single;
    return r;
}

// Incompressible data leaves the Snappy compressor as literal elements only (one per 64 KiB block).  Such a stream is
// not compressed at all, just framed: the planner records where the literal bytes lie in the file and the staging copy
// (pst_plan_fill_raw) lays them down back to back, so the device receives an ordinary uncompressed page image and no
// Snappy work item exists for the page.  Returns false for a stream with any back-reference (or an unreasonable number
// of literals); segs receives {offset inside s, length} of every literal.
static bool literal_segments(const uint8_t *s, int64_t n, int64_t ulen, std::vector<std::pair<int64_t, int32_t>> &segs) {
    constexpr size_t kMaxSegs = 64;
    int64_t ip = 0;
    uint64_t v = 0;
    for (int shift = 0;; shift += 7) {
        if (ip >= n || shift > 35) return false;
        const uint8_t b = s[ip++];
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
    }
    if ((int64_t)v != ulen) return false;
    int64_t op = 0;
    while (ip < n) {
        const uint8_t tag = s[ip];
        if ((tag & 3) != 0) return false;
        const int t6 = tag >> 2;
        int64_t len, hdr;
        if (t6 < 60) {
            len = t6 + 1;
            hdr = 1;
        } else {
            const int nb = t6 - 59;
            if (ip + 1 + nb > n) return false;
            len = 0;
            for (int i = 0; i < nb; i++) len |= (int64_t)s[ip + 1 + i] << (8 * i);
            len += 1;
            hdr = 1 + nb;
        }
        if (len > ulen - op || ip + hdr + len > n || segs.size() >= kMaxSegs) return false;
        segs.emplace_back(ip + hdr, (int32_t)len);
        ip += hdr + len;
        op += len;
    }
    return op == ulen;
}

// Fragment positions of a literal-dominated stream, found on the host.  Blob columns (images, tensors, .npy payloads) are
// practically incompressible: their pages are a handful of long literals with a rare back-reference, and pyarrow puts a
// whole column chunk of large values into ONE page (the page size is only checked per write batch) - a 134 MB page of
// 1 MiB tensors has ~2,500 elements.  Walking those tags here costs microseconds, while the device-side index kernel
// would walk the page as one serial chain.  The walk gives up after `max_elements` elements (a compressible stream: the
// device indexes it) or when an element straddles a 64 KiB output boundary (k_snappy_index flags such pages for the
// serial fallback).  pos[0..nfrag] receives the compressed offsets (same convention as k_snappy_index).
static bool host_fragment_index(const uint8_t *s, int64_t n, int64_t ulen, int nfrag, uint32_t *pos, int64_t max_elements) {
    int64_t ip = 0;
    uint64_t v = 0;
    for (int shift = 0;; shift += 7) {
        if (ip >= n || shift > 35) return false;
        const uint8_t b = s[ip++];
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
    }
    if ((int64_t)v != ulen) return false;
    int64_t op = 0, next_b = kSnappyFragment, elements = 0;
    int k = 1;
    pos[0] = 0;
    while (ip < n) {
        if (++elements > max_elements) return false;
        // a stream that compresses shows it at once: after 64 elements a literal-dominated stream (>= 256 stored bytes per
        // element on average, which is what max_elements encodes) has consumed >= 16 KiB; anything under 4 KiB will not
        // recover, and walking it to the budget cost 1 ms of the 2.9 ms a C2 row-group takes to plan
        if (elements == 64 && ip < 4096) return false;
        if (op == next_b) {
            if (k >= nfrag) return false;
            pos[k++] = (uint32_t)ip;
            next_b += kSnappyFragment;
        }
        const uint8_t tag = s[ip];
        int64_t used, made;
        switch (tag & 3) {
            case 0: {
                const int t6 = tag >> 2;
                if (t6 < 60) {
                    made = t6 + 1;
                    used = 1 + made;
                } else {
                    const int nb = t6 - 59;
                    if (ip + 1 + nb > n) return false;
                    made = 0;
                    for (int i = 0; i < nb; i++) made |= (int64_t)s[ip + 1 + i] << (8 * i);
                    made += 1;
                    used = 1 + nb + made;
                }
                break;
            }
            case 1: used = 2; made = ((tag >> 2) & 7) + 4; break;
            case 2: used = 3; made = (tag >> 2) + 1; break;
            default: used = 5; made = (tag >> 2) + 1; break;
        }
        if (made > next_b - op) return false;      // straddles a fragment boundary
        ip += used;
        op += made;
        if (ip > n || op > ulen) return false;
    }
    if (ip != n || op != ulen || k != nfrag) return false;
    pos[nfrag] = (uint32_t)n;
    return true;
}

int pst_plan_create(const pst_file *f, int rg, const int *cols, int ncols, pst_plan **out) {
    PST_TRY
    *out = nullptr;
    if (rg < 0 || rg >= (int)f->meta.row_groups.size()) throw std::runtime_error("row group index out of range");
    if (ncols <= 0) throw std::runtime_error("a plan needs at least one column");
    const RowGroupMeta &g = f->meta.row_groups[rg];
    std::unique_ptr<pst_plan> p(new pst_plan());
    p->file = f;
    p->rg = rg;
    p->num_rows = g.num_rows;
    p->cols.assign(cols, cols + ncols);
    p->dcols.resize(ncols);

    int64_t raw_cur = 0;      // cursor inside the raw region
    int64_t scratch_cur = 0;  // cursor inside the scratch region (relative; rebased after raw size is known)
    std::vector<int64_t> scratch_rel;  // per page: relative scratch offset or -1
    std::vector<int64_t> dict_index_rel(ncols, -1);
    std::vector<char> stats_no_nulls(ncols, 0);   // the chunk statistics promise null_count == 0

    for (int slot = 0; slot < ncols; slot++) {
        int col = cols[slot];
        if (col < 0 || col >= (int)f->meta.leaves.size()) throw std::runtime_error("column index out of range");
        const ColumnChunkMeta &c = g.columns[col];
        const LeafColumn &leaf = f->meta.leaves[col];
        const SchemaElement &se = f->meta.schema[leaf.schema_index];
        if (!c.file_path.empty()) throw std::runtime_error("column chunks stored in external files are not supported");
        if (c.codec != PST_CODEC_NONE && c.codec != PST_CODEC_SNAPPY && c.codec != PST_CODEC_GZIP)
            throw std::runtime_error("unsupported compression codec " + std::to_string(c.codec) +
                                     " (supported: UNCOMPRESSED, SNAPPY, GZIP) in column " + se.name);
        DevCol &dc = p->dcols[slot];
        memset(&dc, 0, sizeof dc);
        dc.ptype = c.type;
        switch (c.type) {
            case PST_BOOLEAN: dc.width = 1; break;
            case PST_INT32: case PST_FLOAT: dc.width = 4; break;
            case PST_INT64: case PST_DOUBLE: dc.width = 8; break;
            case PST_INT96: dc.width = 12; break;
            case PST_FIXED_LEN_BYTE_ARRAY: dc.width = se.type_length; break;
            case PST_BYTE_ARRAY: dc.width = 0; break;
            default: throw std::runtime_error("unknown physical type");
        }
        dc.max_def = leaf.max_def;
        dc.max_rep = leaf.max_rep;
        dc.num_values = c.num_values;
        dc.dict_img_off = -1;
        dc.dict_page = -1;
        dc.dict_index_off = -1;
        dc.values_off = dc.valid_off = dc.rep_off = dc.def_off = dc.lens_off = -1;
        stats_no_nulls[slot] = c.has_null_count && c.null_count == 0;

        int64_t off = c.start_offset();
        int64_t chunk_end = off + c.total_compressed_size;
        int64_t seen = 0;
        int ordinal = 0;
        while (seen < c.num_values) {
            if (off < 0 || (size_t)off >= f->size || (c.total_compressed_size > 0 && off >= chunk_end))
                throw std::runtime_error("column chunk of " + se.name + " ended before all values were found");
            PageHeader h;
            parse_page_header(f->map + off, std::min<size_t>(f->size - (size_t)off, 1 << 20), h);
            int64_t payload = off + (int64_t)h.header_size;
            if (h.compressed_page_size < 0 || (size_t)(payload + h.compressed_page_size) > f->size)
                throw std::runtime_error("page payload exceeds the file size");
            off = payload + h.compressed_page_size;
            if (h.type == 1) continue;  // INDEX_PAGE
            HostPage hp;
            memset(&hp.d, 0, sizeof hp.d);
            hp.d.multi_slot = -1;
            hp.file_off = payload;
            DevPage &d = hp.d;
            d.comp_size = h.compressed_page_size;
            d.uncomp_size = h.uncompressed_page_size;
            d.num_values = h.num_values;
            d.col = (int16_t)slot;
            d.encoding = (uint8_t)h.encoding;
            d.codec = (uint8_t)c.codec;
            d.page_ordinal = (int16_t)std::min(ordinal++, 32767);
            d.def_bytes = d.rep_bytes = -1;
            int64_t values_off_in_image = -1;
            bool all_valid = false;
            if (h.type == 2) {
                d.kind = PK_DICT;
                if (h.encoding != ENC_PLAIN && h.encoding != ENC_PLAIN_DICTIONARY)
                    throw std::runtime_error("dictionary page with unsupported encoding");
                if (dc.dict_page >= 0) throw std::runtime_error("more than one dictionary page in a column chunk");
                dc.dict_page = (int32_t)p->pages.size();
                dc.dict_count = h.num_values;
                values_off_in_image = 0;
            } else if (h.type == 0) {
                d.kind = PK_DATA_V1;
                d.def_enc = (uint8_t)h.def_encoding;
                d.rep_enc = (uint8_t)h.rep_encoding;
                d.first_value = (int32_t)seen;
                seen += h.num_values;
                const LevelPeek lp = v1_level_peek(f->map + payload, (size_t)h.compressed_page_size, c.codec, leaf.max_rep,
                                                   leaf.max_def, h.rep_encoding, h.def_encoding, h.num_values);
                values_off_in_image = lp.values_off;
                all_valid = lp.all_valid;
            } else if (h.type == 3) {
                d.kind = PK_DATA_V2;
                d.def_bytes = h.def_bytes;
                d.rep_bytes = h.rep_bytes;
                d.v2_compressed = h.is_compressed ? 1 : 0;
                d.first_value = (int32_t)seen;
                seen += h.num_values;
                values_off_in_image = (int64_t)h.def_bytes + h.rep_bytes;
                if (values_off_in_image > h.compressed_page_size || values_off_in_image > h.uncompressed_page_size)
                    throw std::runtime_error("V2 page level sections exceed the page size");
                if (!h.is_compressed) d.codec = PST_CODEC_NONE;
                all_valid = leaf.max_def == 0 ||
                            (h.num_nulls == 0 && h.def_bytes > 0 &&
                             single_rle_run(f->map + payload + h.rep_bytes, (size_t)h.def_bytes, (size_t)h.def_bytes,
                                            level_bits(leaf.max_def), (uint32_t)leaf.max_def, h.num_values));
            } else {
                throw std::runtime_error("unknown page type " + std::to_string(h.type));
            }
            if (d.kind != PK_DICT) {
                switch (h.encoding) {
                    case ENC_PLAIN: case ENC_PLAIN_DICTIONARY: case ENC_RLE_DICTIONARY: break;
                    case ENC_RLE:
                        if (c.type != PST_BOOLEAN) throw std::runtime_error("RLE value encoding on a non-boolean column");
                        break;
                    default:
                        throw std::runtime_error("unsupported value encoding " + std::to_string(h.encoding) +
                                                 " in column " + se.name);
                }
                if (all_valid) d.flags |= PF_ALL_VALID;
            }
            // placement: payload as stored goes into raw; phase chosen so the value section is 16B aligned
            int64_t phase = 0;
            if (values_off_in_image > 0) phase = (16 - values_off_in_image % 16) % 16;
            hp.stored_size = h.compressed_page_size;
            bool compressed = d.codec != PST_CODEC_NONE && d.comp_size > 0;
            if (compressed && d.codec == PST_CODEC_SNAPPY) {
                // the compressed stream covers the whole image (V1) or everything behind the level bytes (V2)
                const int64_t lv = d.kind == PK_DATA_V2 ? (int64_t)d.def_bytes + d.rep_bytes : 0;
                std::vector<std::pair<int64_t, int32_t>> segs;
                if (literal_segments(f->map + payload + lv, (int64_t)d.comp_size - lv, (int64_t)d.uncomp_size - lv, segs)) {
                    // framed, not compressed: the staging copy lays the literal bytes down back to back
                    if (lv > 0) hp.segs.emplace_back(payload, (int32_t)lv);
                    for (const auto &sg : segs) hp.segs.emplace_back(payload + lv + sg.first, sg.second);
                    d.codec = PST_CODEC_NONE;
                    d.comp_size = d.uncomp_size;
                    d.flags |= PF_UNWRAPPED;
                    compressed = false;
                    p->unwrapped_pages++;
                }
            }
            if (compressed) {
                d.src_off = align_up(raw_cur, 16);
                raw_cur = d.src_off + d.comp_size;
                int64_t rel = align_up(scratch_cur, 16) + phase;
                scratch_rel.push_back(rel);
                scratch_cur = rel + d.uncomp_size + 16;  // +16: vector-store slack
                if (d.codec == PST_CODEC_GZIP) {
                    p->gzip_pages.push_back((int32_t)p->pages.size());
                } else {
                    p->compressed_pages.push_back((int32_t)p->pages.size());
                    const int64_t values_uncomp = (int64_t)d.uncomp_size -
                                                  (d.kind == PK_DATA_V2 ? (int64_t)d.def_bytes + d.rep_bytes : 0);
                    d.nfrag = (int32_t)std::max<int64_t>(1, (values_uncomp + kSnappyFragment - 1) / kSnappyFragment);
                    d.frag_first = (int32_t)p->frag_pos_count;
                    p->frag_pos_count += d.nfrag + 1;
                    p->frag_pos_host.resize((size_t)p->frag_pos_count, 0u);
                    if (d.nfrag > 1) {
                        d.multi_slot = (int32_t)p->multi_pages.size();
                        p->multi_pages.push_back((int32_t)p->pages.size());
                        // literal-dominated streams (>= 256 stored bytes per element on average) are indexed right
                        // here; everything else by k_snappy_index
                        const int64_t lv = (int64_t)d.uncomp_size - values_uncomp;   // V2: uncompressed level bytes
                        if (host_fragment_index(f->map + payload + lv, (int64_t)d.comp_size - lv, values_uncomp, d.nfrag,
                                                p->frag_pos_host.data() + d.frag_first,
                                                std::max<int64_t>(64, ((int64_t)d.comp_size - lv) / 256)))
                            p->host_indexed_pages++;
                        else
                            p->index_pages.push_back((int32_t)p->pages.size());
                    }
                    for (int32_t k = 0; k < d.nfrag; k++)
                        p->snappy_frags.push_back(SnFrag{(int32_t)p->pages.size(), k});
                }
            } else {
                d.src_off = align_up(raw_cur, 16) + phase;
                raw_cur = d.src_off + d.comp_size + 16;  // +16: vector-load slack
                d.img_off = d.src_off;
                scratch_rel.push_back(-1);
            }
            // PLAIN fixed-width values of a flat page without nulls are a plain byte range of the page image: they go
            // to `out` through the tile copy kernel (k_copy_tiles) instead of the general page decoder
            hp.copy_v0 = -1;
            if (d.kind != PK_DICT && all_valid && values_off_in_image >= 0 && h.encoding == ENC_PLAIN &&
                leaf.max_rep == 0 && dc.width > 0 && c.type != PST_BOOLEAN && d.codec != PST_CODEC_GZIP &&
                values_off_in_image + (int64_t)h.num_values * dc.width <= (int64_t)d.uncomp_size) {
                hp.copy_v0 = values_off_in_image;
                d.flags |= PF_COPY;
            }
            if (d.kind != PK_DICT && !(d.flags & PF_COPY)) p->data_pages.push_back((int32_t)p->pages.size());
            p->payload_bytes += hp.stored_size;
            p->uncompressed_bytes += d.uncomp_size;
            p->pages.push_back(hp);
        }
        if (seen != c.num_values) throw std::runtime_error("page value counts do not add up to the chunk's num_values");
        if (dc.dict_page >= 0 && c.type == PST_BYTE_ARRAY) {
            // {offset, length} of every dictionary entry: a serial walk over the length prefixes.  When the device
            // receives the page uncompressed (stored that way, or literal-only Snappy) the planner walks it right here
            // and ships the table with the raw region; otherwise k_ba_dict_index does it after decompression.
            const HostPage &dp = p->pages[dc.dict_page];
            std::vector<BaDictEntry> entries;
            bool on_host = dp.d.codec == PST_CODEC_NONE && dc.dict_count <= 65536;
            if (on_host) {
                auto image_byte = [&](int64_t off) -> int {        // byte `off` of the page image, -1 behind its end
                    if (dp.segs.empty()) return off < dp.d.comp_size ? f->map[dp.file_off + off] : -1;
                    for (const auto &sg : dp.segs) {
                        if (off < sg.second) return f->map[sg.first + off];
                        off -= sg.second;
                    }
                    return -1;
                };
                int64_t pos = 0;
                entries.reserve((size_t)dc.dict_count);
                for (int i = 0; i < dc.dict_count && on_host; i++) {
                    uint32_t len = 0;
                    for (int b = 0; b < 4; b++) {
                        const int v = image_byte(pos + b);
                        if (v < 0) { on_host = false; break; }
                        len |= (uint32_t)v << (8 * b);
                    }
                    pos += 4;
                    if (!on_host || pos + (int64_t)len > (int64_t)dp.d.uncomp_size) { on_host = false; break; }
                    entries.push_back(BaDictEntry{dp.d.img_off + pos, (int32_t)len, 0});
                    pos += len;
                }
            }
            if (on_host) {
                p->host_dict_index.emplace_back(slot, std::move(entries));      // placed with the tables below
            } else {      // (a malformed dictionary is left to the device kernel, which reports it)
                int64_t rel = align_up(scratch_cur, 16);
                dict_index_rel[slot] = rel;
                scratch_cur = rel + (int64_t)dc.dict_count * 16;
                p->ba_dict_pages.push_back(dc.dict_page);
            }
        }
    }

    // Launch order of the Snappy work items: the better a page compressed, the more elements its stream has per output
    // byte and the longer its fragments take; those go first so that the tail of the grid is made of short items.
    auto ratio_less = [&](int32_t a, int32_t b) {
        const DevPage &x = p->pages[a].d, &y = p->pages[b].d;
        return (int64_t)x.comp_size * y.uncomp_size < (int64_t)y.comp_size * x.uncomp_size;
    };
    std::stable_sort(p->snappy_frags.begin(), p->snappy_frags.end(),
                     [&](const SnFrag &a, const SnFrag &b) { return ratio_less(a.page, b.page); });
    {   // same order for the index kernel; multi_slot follows the sorted list
        std::stable_sort(p->multi_pages.begin(), p->multi_pages.end(), ratio_less);
        // the index kernel spends one CTA per page and its time goes with the page's compressed size: longest first, so
        // that the launch does not end on a 1 MiB dictionary page that started in the second wave
        std::stable_sort(p->index_pages.begin(), p->index_pages.end(), [&](int32_t a, int32_t b) {
            return p->pages[a].d.comp_size > p->pages[b].d.comp_size;
        });
        // pages of 256 KiB and more (dictionary pages, mostly) get four SMs each: k_snappy_index_cluster
        p->index_big_count = 0;
        for (int32_t pi : p->index_pages)
            if (p->pages[pi].d.comp_size >= 256 * 1024) p->index_big_count++;
        for (size_t i = 0; i < p->multi_pages.size(); i++) p->pages[p->multi_pages[i]].d.multi_slot = (int32_t)i;
    }

    // ---- tables at the tail of the raw region
    int64_t npages = (int64_t)p->pages.size();
    p->tables_off = align_up(raw_cur, 256);
    p->cols_off = p->tables_off;
    p->pages_off = align_up(p->cols_off + (int64_t)sizeof(DevCol) * ncols, 64);
    p->comp_list_off = align_up(p->pages_off + (int64_t)sizeof(DevPage) * npages, 16);
    p->data_list_off = align_up(p->comp_list_off + 4 * (int64_t)p->compressed_pages.size(), 16);
    p->dict_list_off = align_up(p->data_list_off + 4 * (int64_t)p->data_pages.size(), 16);
    p->frag_list_off = align_up(p->dict_list_off + 4 * (int64_t)p->ba_dict_pages.size(), 16);
    p->multi_list_off = align_up(p->frag_list_off + (int64_t)sizeof(SnFrag) * (int64_t)p->snappy_frags.size(), 16);
    p->gzip_list_off = align_up(p->multi_list_off + 4 * (int64_t)p->multi_pages.size(), 16);
    p->index_list_off = align_up(p->gzip_list_off + 4 * (int64_t)p->gzip_pages.size(), 16);
    // fragment positions (host-filled for literal-only pages, written by k_snappy_index for the others) and page flags
    // (zero; raised on the device): part of the raw image so that every upload resets them
    p->copy_tiles_off = align_up(p->index_list_off + 4 * (int64_t)p->index_pages.size(), 32);
    // (the tile list itself is built below, once the out layout is known; its size only depends on the pages)
    int64_t n_tiles = 0;
    for (const HostPage &hp : p->pages)
        if (hp.copy_v0 >= 0) {
            const int64_t w = p->dcols[hp.d.col].width, per = (kCopyTileBytes / w) * w;
            n_tiles += ((int64_t)hp.d.num_values * w + per - 1) / per;
        }
    int64_t dict_tab_cur = align_up(p->copy_tiles_off + (int64_t)sizeof(CopyTile) * n_tiles, 16);
    std::vector<int64_t> host_dict_off(p->host_dict_index.size());
    for (size_t i = 0; i < p->host_dict_index.size(); i++) {
        host_dict_off[i] = dict_tab_cur;
        p->dcols[p->host_dict_index[i].first].dict_index_off = dict_tab_cur;
        dict_tab_cur += (int64_t)sizeof(BaDictEntry) * (int64_t)p->host_dict_index[i].second.size();
    }
    p->frag_pos_off = align_up(dict_tab_cur, 16);
    p->page_flag_off = align_up(p->frag_pos_off + 4 * p->frag_pos_count, 16);
    p->raw_bytes = align_up(p->page_flag_off + 4 * (int64_t)p->multi_pages.size(), 256);
    p->scratch_off = p->raw_bytes;
    p->arena_bytes = align_up(p->scratch_off + scratch_cur + 256, 256);

    for (size_t i = 0; i < p->pages.size(); i++)
        if (scratch_rel[i] >= 0) p->pages[i].d.img_off = p->scratch_off + scratch_rel[i];

    // ---- out region
    int64_t out_cur = 0;
    for (int slot = 0; slot < ncols; slot++) {
        DevCol &dc = p->dcols[slot];
        int64_t n = dc.num_values;
        if (dc.dict_page >= 0) dc.dict_img_off = p->pages[dc.dict_page].d.img_off;
        if (dict_index_rel[slot] >= 0) dc.dict_index_off = p->scratch_off + dict_index_rel[slot];
        out_cur = align_up(out_cur, 256);
        dc.values_off = out_cur;
        if (dc.ptype == PST_BYTE_ARRAY) {
            out_cur += 8 * n;
            out_cur = align_up(out_cur, 256);
            dc.lens_off = out_cur;
            out_cur += 4 * n;
        } else {
            out_cur += (int64_t)dc.width * n;
        }
        // A flat column whose chunk statistics say null_count == 0 carries no validity array: every page is all-valid
        // (the page decoder still counts level entries below max_def, and the host raises if the statistics lied).
        if (dc.max_def > 0 && !(stats_no_nulls[slot] && dc.max_rep == 0)) {
            out_cur = align_up(out_cur, 256);
            dc.valid_off = out_cur;
            out_cur += n;
        }
        if (dc.max_rep > 0) {
            out_cur = align_up(out_cur, 256);
            dc.rep_off = out_cur;
            out_cur += n;
            out_cur = align_up(out_cur, 256);
            dc.def_off = out_cur;
            out_cur += n;
        }
    }
    p->out_bytes = align_up(out_cur + 16, 256);

    // ---- tile list of the copy kernel
    for (const HostPage &hp : p->pages) {
        if (hp.copy_v0 < 0) continue;
        const DevCol &dc = p->dcols[hp.d.col];
        const int64_t w = dc.width, per = (kCopyTileBytes / w) * w, total = (int64_t)hp.d.num_values * w;
        for (int64_t a = 0; a < total; a += per) {
            CopyTile t;
            t.src_off = hp.d.img_off + hp.copy_v0 + a;
            t.dst_off = dc.values_off + (int64_t)hp.d.first_value * w + a;
            t.nbytes = (int32_t)std::min(per, total - a);
            t.valid_off = dc.valid_off >= 0 ? dc.valid_off + hp.d.first_value + a / w : -1;
            t.nvalid = dc.valid_off >= 0 ? t.nbytes / (int32_t)w : 0;
            p->copy_tiles.push_back(t);
        }
    }
    if ((int64_t)p->copy_tiles.size() != n_tiles) throw std::runtime_error("internal error: copy tile count");

    // ---- host image of the tables
    p->tables.assign((size_t)(p->raw_bytes - p->tables_off), 0);
    uint8_t *t = p->tables.data();
    memcpy(t + (p->cols_off - p->tables_off), p->dcols.data(), sizeof(DevCol) * ncols);
    for (int64_t i = 0; i < npages; i++)
        memcpy(t + (p->pages_off - p->tables_off) + i * (int64_t)sizeof(DevPage), &p->pages[i].d, sizeof(DevPage));
    if (!p->compressed_pages.empty())
        memcpy(t + (p->comp_list_off - p->tables_off), p->compressed_pages.data(), 4 * p->compressed_pages.size());
    if (!p->data_pages.empty())
        memcpy(t + (p->data_list_off - p->tables_off), p->data_pages.data(), 4 * p->data_pages.size());
    if (!p->ba_dict_pages.empty())
        memcpy(t + (p->dict_list_off - p->tables_off), p->ba_dict_pages.data(), 4 * p->ba_dict_pages.size());
    if (!p->snappy_frags.empty())
        memcpy(t + (p->frag_list_off - p->tables_off), p->snappy_frags.data(), sizeof(SnFrag) * p->snappy_frags.size());
    if (!p->multi_pages.empty())
        memcpy(t + (p->multi_list_off - p->tables_off), p->multi_pages.data(), 4 * p->multi_pages.size());
    if (!p->gzip_pages.empty())
        memcpy(t + (p->gzip_list_off - p->tables_off), p->gzip_pages.data(), 4 * p->gzip_pages.size());
    if (!p->index_pages.empty())
        memcpy(t + (p->index_list_off - p->tables_off), p->index_pages.data(), 4 * p->index_pages.size());
    if (!p->copy_tiles.empty())
        memcpy(t + (p->copy_tiles_off - p->tables_off), p->copy_tiles.data(), sizeof(CopyTile) * p->copy_tiles.size());
    for (size_t i = 0; i < p->host_dict_index.size(); i++)
        if (!p->host_dict_index[i].second.empty())
            memcpy(t + (host_dict_off[i] - p->tables_off), p->host_dict_index[i].second.data(),
                   sizeof(BaDictEntry) * p->host_dict_index[i].second.size());
    if (!p->frag_pos_host.empty())
        memcpy(t + (p->frag_pos_off - p->tables_off), p->frag_pos_host.data(), 4 * p->frag_pos_host.size());

    // cache key: file identity + row group + column set
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](uint64_t v) {
        for (int i = 0; i < 8; i++) {
            key ^= (v >> (8 * i)) & 0xff;
            key *= 1099511628211ull;
        }
    };
    for (unsigned char ch : f->path) mix(ch);
    mix((uint64_t)f->size);
    mix((uint64_t)f->mtime_ns);
    mix((uint64_t)rg);
    for (int c : p->cols) mix((uint64_t)c);
    p->cache_key = key;

    *out = p.release();
    return 0;
    PST_CATCH(1)
}

void pst_plan_destroy(pst_plan *p) { delete p; }

int pst_plan_get_info(const pst_plan *p, pst_plan_info *out) {
    out->num_rows = p->num_rows;
    out->raw_bytes = p->raw_bytes;
    out->arena_bytes = p->arena_bytes;
    out->out_bytes = p->out_bytes;
    out->payload_bytes = p->payload_bytes;
    out->uncompressed_bytes = p->uncompressed_bytes;
    out->num_pages = (int32_t)p->pages.size();
    out->num_columns = (int32_t)p->cols.size();
    out->num_compressed_pages = (int32_t)(p->compressed_pages.size() + p->gzip_pages.size());
    out->num_index_pages = (int32_t)p->index_pages.size();
    out->num_unwrapped_pages = (int32_t)p->unwrapped_pages;
    out->num_host_indexed_pages = (int32_t)p->host_indexed_pages;
    out->num_cluster_index_pages = p->index_big_count;
    out->num_copy_tiles = (int32_t)p->copy_tiles.size();
    out->num_decode_pages = (int32_t)p->data_pages.size();
    out->num_snappy_fragments = (int32_t)p->snappy_frags.size();
    return 0;
}

int pst_plan_get_column(const pst_plan *p, int i, pst_plan_column *out) {
    if (i < 0 || i >= (int)p->dcols.size()) {
        set_error("plan column out of range");
        return 1;
    }
    const DevCol &dc = p->dcols[i];
    out->column = p->cols[i];
    out->physical_type = dc.ptype;
    out->type_length = dc.width;
    out->max_def = dc.max_def;
    out->max_rep = dc.max_rep;
    out->has_dictionary = dc.dict_page >= 0;
    out->num_values = dc.num_values;
    out->values_off = dc.values_off;
    out->lens_off = dc.lens_off;
    out->valid_off = dc.valid_off;
    out->rep_off = dc.rep_off;
    out->def_off = dc.def_off;
    return 0;
}

int pst_plan_get_page(const pst_plan *p, int i, pst_plan_page *out) {
    if (i < 0 || i >= (int)p->pages.size()) {
        set_error("plan page out of range");
        return 1;
    }
    const HostPage &hp = p->pages[i];
    out->column_slot = hp.d.col;
    out->kind = hp.d.kind;
    out->encoding = hp.d.encoding;
    out->codec = hp.d.codec;
    out->flags = hp.d.flags;
    out->stored_bytes = hp.stored_size;
    out->image_bytes = hp.d.uncomp_size;
    out->num_values = hp.d.num_values;
    out->first_value = hp.d.first_value;
    out->fragments = hp.d.nfrag;
    out->src_off = hp.d.src_off;
    out->img_off = hp.d.img_off;
    return 0;
}

int pst_plan_get_copy_tile(const pst_plan *p, int i, pst_copy_tile *out) {
    if (i < 0 || i >= (int)p->copy_tiles.size()) {
        set_error("copy tile out of range");
        return 1;
    }
    const CopyTile &t = p->copy_tiles[i];
    out->src_off = t.src_off;
    out->dst_off = t.dst_off;
    out->valid_off = t.valid_off;
    out->nbytes = t.nbytes;
    out->nvalid = t.nvalid;
    return 0;
}

// Copies the plan's raw region image (payloads at their planned offsets + tables) into `dst` (raw_bytes bytes).
// Host-only helper used by the staging path and by CPU tests of the planner.
int pst_plan_fill_raw(const pst_plan *p, uint8_t *dst, int64_t first_page, int64_t last_page) {
    int64_t n = (int64_t)p->pages.size();
    if (first_page < 0) first_page = 0;
    if (last_page > n) last_page = n;
    for (int64_t i = first_page; i < last_page; i++) {
        const HostPage &hp = p->pages[i];
        if (hp.segs.empty()) {
            memcpy(dst + hp.d.src_off, p->file->map + hp.file_off, (size_t)hp.d.comp_size);
        } else {   // literal-only Snappy page: the literal bytes back to back = the uncompressed page image
            uint8_t *o = dst + hp.d.src_off;
            for (const auto &sg : hp.segs) {
                memcpy(o, p->file->map + sg.first, (size_t)sg.second);
                o += sg.second;
            }
        }
    }
    if (first_page == 0) memcpy(dst + p->tables_off, p->tables.data(), p->tables.size());
    return 0;
}

}  // extern "C"
