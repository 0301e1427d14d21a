"""TEST INFRASTRUCTURE: a slow numpy interpreter of a plan's raw-region image (page payloads + device tables).

It follows the same tables the CUDA kernels read (``petastorm_b200/csrc/dev_structs.h``) so that the host planner
(thrift footer, page walk, HBM layout) can be checked against pyarrow on a machine without a GPU.  It is NOT used by
the product and is not an oracle for the kernels themselves (the GPU parity tests compare kernel output with pyarrow).
"""
import struct

import numpy as np
import pyarrow as pa

DEVPAGE = struct.Struct('<qqiiiiiihBBBBBBi12x')
DEVCOL = struct.Struct('<iiiiqqqqqqqiiqq')


def _snappy(buf, n):
    return pa.Codec('snappy').decompress(bytes(buf), decompressed_size=n).to_pybytes()


def _decompress(codec, buf, n):
    if codec == 2:
        import zlib
        out = zlib.decompress(bytes(buf), 31)
        assert len(out) == n
        return out
    return _snappy(buf, n)


def _hybrid(buf, pos, end, bw, count):
    """RLE/bit-packed hybrid -> list of `count` ints (parquet Encodings.md)."""
    out = []
    while len(out) < count:
        h, shift = 0, 0
        while True:
            b = buf[pos]
            pos += 1
            h |= (b & 0x7f) << shift
            if not b & 0x80:
                break
            shift += 7
        if h & 1:
            groups = h >> 1
            nbytes = groups * bw
            bits = int.from_bytes(bytes(buf[pos:pos + nbytes]), 'little')
            pos += nbytes
            mask = (1 << bw) - 1
            for i in range(groups * 8):
                if len(out) >= count:
                    break
                out.append((bits >> (i * bw)) & mask)
        else:
            run = h >> 1
            nb = (bw + 7) // 8
            v = int.from_bytes(bytes(buf[pos:pos + nb]), 'little')
            pos += nb
            out.extend([v] * min(run, count - len(out)))
    return out, pos


def _bits_for(level):
    return 0 if level == 0 else int(level).bit_length()


def read_tables(plan):
    """(host copy of the raw region as uint8 array, DevCol tuples, DevPage tuples, arena offset of the page table)."""
    info = plan.info
    arena = np.zeros(info.arena_bytes, dtype=np.uint8)
    plan.fill_raw(arena.ctypes.data)
    # table offsets are not exported through the C-ABI; recover them from the layout rules: cols at a 256-aligned
    # tables_off behind the payloads, pages at align64(cols_end)
    ncols, npages = info.num_columns, info.num_pages
    for t in range(0, info.raw_bytes, 256):
        cols_end = t + DEVCOL.size * ncols
        pages_off = (cols_end + 63) // 64 * 64
        if pages_off + DEVPAGE.size * npages > info.raw_bytes:
            break
        ok = True
        for i in range(ncols):
            c = DEVCOL.unpack_from(arena, t + i * DEVCOL.size)
            pc = plan.cols[i]
            if c[0] != pc.physical_type or c[2] != pc.max_def or c[4] != pc.num_values or c[5] != pc.values_off:
                ok = False
                break
        if ok:
            tables_off = t
            break
    else:
        raise AssertionError('tables not found')
    cols = [DEVCOL.unpack_from(arena, tables_off + i * DEVCOL.size) for i in range(ncols)]
    pages_off = (tables_off + DEVCOL.size * ncols + 63) // 64 * 64
    pages = [DEVPAGE.unpack_from(arena, pages_off + i * DEVPAGE.size) for i in range(npages)]
    return arena, cols, pages, pages_off


def decode_plan(plan, native):
    """Returns {slot: dict(values=..., valid=..., rep=..., defs=...)} decoded on the CPU from the raw image."""
    arena, cols, pages, _ = read_tables(plan)
    result = {}
    images = {}
    for pi, pg in enumerate(pages):
        (src_off, img_off, comp, uncomp, nvals, first, def_bytes, rep_bytes, col, kind, enc, codec, def_enc, rep_enc,
         v2c, ordinal) = pg
        payload = arena[src_off:src_off + comp]
        if codec in (1, 2) and comp > 0:
            if kind == 3:
                lv = def_bytes + rep_bytes
                img = bytes(payload[:lv]) + _decompress(codec, payload[lv:], uncomp - lv)
            else:
                img = _decompress(codec, payload, uncomp)
        else:
            img = bytes(payload)
        assert len(img) == uncomp, (pi, len(img), uncomp)
        images[pi] = img
    for slot, c in enumerate(cols):
        (ptype, width, max_def, max_rep, nv, values_off, valid_off, rep_off, def_off, lens_off, dict_img_off,
         dict_count, dict_page, dict_index_off, _pad) = c
        dict_vals = None
        if dict_page >= 0:
            dimg = images[dict_page]
            dict_vals = _plain(dimg, 0, ptype, width, dict_count)
        vals = [None] * nv
        reps = [0] * nv
        defs = [max_def] * nv
        for pi, pg in enumerate(pages):
            if pg[8] != slot or pg[9] == 2:
                continue
            (src_off, img_off, comp, uncomp, nvals, first, def_bytes, rep_bytes, col, kind, enc, codec, def_enc,
             rep_enc, v2c, ordinal) = pg
            img = images[pi]
            pos = 0
            rl = dl = None
            if kind == 3:
                if max_rep:
                    rl, _ = _hybrid(img, 0, rep_bytes, _bits_for(max_rep), nvals)
                if max_def:
                    dl, _ = _hybrid(img, rep_bytes, rep_bytes + def_bytes, _bits_for(max_def), nvals)
                pos = rep_bytes + def_bytes
            else:
                if max_rep:
                    ln = struct.unpack_from('<I', img, pos)[0]
                    rl, _ = _hybrid(img, pos + 4, pos + 4 + ln, _bits_for(max_rep), nvals)
                    pos += 4 + ln
                if max_def:
                    ln = struct.unpack_from('<I', img, pos)[0]
                    dl, _ = _hybrid(img, pos + 4, pos + 4 + ln, _bits_for(max_def), nvals)
                    pos += 4 + ln
            nvalid = nvals if dl is None else sum(1 for d in dl if d == max_def)
            if enc in (2, 8):
                bw = img[pos]
                idx, _ = _hybrid(img, pos + 1, len(img), bw, nvalid)
                pv = [dict_vals[i] for i in idx]
            elif enc == 3:
                ln = struct.unpack_from('<I', img, pos)[0]
                pv, _ = _hybrid(img, pos + 4, pos + 4 + ln, 1, nvalid)
                pv = [bool(v) for v in pv]
            else:
                pv = _plain(img, pos, ptype, width, nvalid)
            it = iter(pv)
            for i in range(nvals):
                d = max_def if dl is None else dl[i]
                defs[first + i] = d
                if rl is not None:
                    reps[first + i] = rl[i]
                vals[first + i] = next(it) if d == max_def else None
        result[slot] = dict(values=vals, defs=defs, reps=reps, ptype=ptype)
    return result


def _plain(img, pos, ptype, width, count):
    if ptype == 0:
        bits = np.unpackbits(np.frombuffer(img, dtype=np.uint8, offset=pos), bitorder='little')
        return [bool(b) for b in bits[:count]]
    if ptype in (1, 2, 4, 5):
        dt = {1: '<i4', 2: '<i8', 4: '<f4', 5: '<f8'}[ptype]
        return list(np.frombuffer(img, dtype=dt, count=count, offset=pos))
    if ptype == 6:
        out = []
        for _ in range(count):
            ln = struct.unpack_from('<I', img, pos)[0]
            out.append(bytes(img[pos + 4:pos + 4 + ln]))
            pos += 4 + ln
        return out
    return [bytes(img[pos + i * width:pos + (i + 1) * width]) for i in range(count)]
