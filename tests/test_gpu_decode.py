"""GPU parity of the page-decode kernels (Snappy, RLE/bit-packed hybrid, PLAIN, dictionary, validity) against Arrow C++
(pyarrow), the decoder the reference reaches through ``piece.read`` (petastorm/arrow_reader_worker.py:358).
Bit-exact comparison; everything goes through the C-ABI (petastorm_b200.native)."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

pytestmark = pytest.mark.gpu


def _decode_file(path, columns=None):
    import torch
    from petastorm_b200 import rowgroup
    dec = rowgroup.RowGroupDecoder()
    f = rowgroup.open_file(path)
    leaves = list(range(f.num_columns)) if columns is None else columns
    out = []
    for rg in range(f.num_row_groups):
        d = dec.decode(path, rg, leaves)
        d.check()
        torch.cuda.synchronize()
        out.append(d)
    return f, out


def _compare_flat(path, **kw):
    from petastorm_b200 import rowgroup
    f, decoded = _decode_file(path)
    pf = pq.ParquetFile(path)
    for rg, d in enumerate(decoded):
        tbl = pf.read_row_group(rg)
        assert d.num_rows == tbl.num_rows
        for slot in range(f.num_columns):
            leaf = f.schema['leaves'][slot]
            col = d.column(slot)
            exp = tbl.column(leaf['name'])
            if len(leaf['path']) != 1:
                continue
            exp_valid = np.asarray(exp.is_valid().to_numpy(zero_copy_only=False))
            if pa.types.is_null(exp.type):
                assert col.valid is not None and not col.valid.cpu().numpy().any()
                continue
            if col.valid is not None:
                got_valid = col.valid.cpu().numpy().astype(bool)
                np.testing.assert_array_equal(got_valid, exp_valid, err_msg='validity of %s rg %d' % (leaf['name'], rg))
            else:
                assert exp_valid.all()
            if col.physical_type == rowgroup.BYTE_ARRAY:
                got = rowgroup.gather_blobs_to_host(col)
                expl = exp.to_pylist()
                expl = [e.encode() if isinstance(e, str) else e for e in expl]
                assert got == expl, 'column %s rg %d' % (leaf['name'], rg)
                continue
            got = col.values.cpu().numpy()
            if pa.types.is_boolean(exp.type):
                e = np.asarray(exp.fill_null(False).to_numpy(zero_copy_only=False)).astype(np.uint8)
            elif pa.types.is_timestamp(exp.type) or pa.types.is_date(exp.type):
                e = np.asarray(exp.cast(pa.int64() if pa.types.is_timestamp(exp.type) else pa.int32())
                               .fill_null(0).to_numpy(zero_copy_only=False))
            else:
                e = np.asarray(exp.fill_null(0).to_numpy(zero_copy_only=False))
                if e.dtype.itemsize < got.dtype.itemsize or e.dtype.kind == 'u':
                    # int8/int16/uint* logical types are stored as INT32/INT64
                    e = e.astype(got.dtype) if e.dtype.itemsize <= got.dtype.itemsize else e.view(got.dtype)
            sel = exp_valid
            assert got.dtype.itemsize == e.dtype.itemsize, (leaf['name'], got.dtype, e.dtype)
            np.testing.assert_array_equal(got.view(e.dtype)[sel], e[sel], err_msg='values of %s rg %d' % (leaf['name'], rg))
            assert (got[~sel] == 0).all()


def _mixed_table(n, seed=1234):
    rng = np.random.default_rng(seed)
    cols = {}
    for i in range(3):
        cols['f%02d' % i] = rng.standard_normal(n).astype(np.float32)
    cols['d0'] = rng.standard_normal(n)
    for i in range(2):
        cols['i%02d' % i] = rng.integers(0, 2 ** 40, n, dtype=np.int64)
    cols['lowcard'] = rng.integers(0, 100, n, dtype=np.int64)
    cols['runs'] = np.repeat(rng.integers(0, 5, (n + 99) // 100, dtype=np.int32), 100)[:n]
    cols['b'] = rng.integers(0, 2, n).astype(bool)
    cols['i8'] = rng.integers(-128, 127, n).astype(np.int8)
    cols['u16'] = rng.integers(0, 65535, n).astype(np.uint16)
    cols['u32'] = rng.integers(0, 2 ** 32 - 1, n).astype(np.uint32)
    t = pa.table(cols)
    t = t.append_column('nullable_i32', pa.array([None if i % 7 == 0 else i for i in range(n)], type=pa.int32()))
    t = t.append_column('nullable_f64', pa.array([None if i % 3 == 0 else i * 0.5 for i in range(n)], type=pa.float64()))
    t = t.append_column('nullable_b', pa.array([None if i % 5 == 0 else bool(i & 1) for i in range(n)], type=pa.bool_()))
    t = t.append_column('mostly_null', pa.array([i if i % 997 == 0 else None for i in range(n)], type=pa.int64()))
    t = t.append_column('all_null', pa.array([None] * n, type=pa.float32()))
    t = t.append_column('s', pa.array(['str%d' % (i % 1000) for i in range(n)]))
    t = t.append_column('s_unique', pa.array([None if i % 11 == 0 else 'value-%d-%s' % (i, 'x' * (i % 17)) for i in range(n)]))
    t = t.append_column('bin', pa.array([bytes([i % 256]) * (i % 40) for i in range(n)], type=pa.binary()))
    t = t.append_column('ts', pa.array(np.arange(n, dtype=np.int64) * 1000, type=pa.timestamp('us')))
    return t


@pytest.mark.parametrize('compression', ['snappy', 'none', 'gzip'])
@pytest.mark.parametrize('version', ['1.0', '2.0'])
def test_mixed_columns(tmp_path, compression, version):
    t = _mixed_table(150003)
    path = str(tmp_path / 'mixed.parquet')
    pq.write_table(t, path, compression=compression, row_group_size=100000, data_page_version=version)
    _compare_flat(path)


def test_small_pages_no_dictionary(tmp_path):
    t = _mixed_table(20011, seed=7)
    path = str(tmp_path / 'small.parquet')
    pq.write_table(t, path, compression='snappy', row_group_size=7000, data_page_size=512, use_dictionary=False)
    _compare_flat(path)


def test_required_columns_uncompressed(tmp_path):
    n = 123457
    rng = np.random.default_rng(3)
    schema = pa.schema([pa.field('a', pa.float32(), nullable=False), pa.field('b', pa.int64(), nullable=False),
                        pa.field('c', pa.int32(), nullable=False), pa.field('d', pa.bool_(), nullable=False)])
    t = pa.table({'a': rng.standard_normal(n).astype(np.float32), 'b': rng.integers(-2 ** 62, 2 ** 62, n),
                  'c': rng.integers(-2 ** 31, 2 ** 31 - 1, n).astype(np.int32), 'd': rng.integers(0, 2, n).astype(bool)},
                 schema=schema)
    path = str(tmp_path / 'req.parquet')
    pq.write_table(t, path, compression='none', use_dictionary=False, row_group_size=50000)
    _compare_flat(path)


def test_empty_and_tiny_row_groups(tmp_path):
    t = _mixed_table(5, seed=11)
    path = str(tmp_path / 'tiny.parquet')
    pq.write_table(t, path, compression='snappy', row_group_size=2)
    _compare_flat(path)
    t0 = _mixed_table(1, seed=2)
    path0 = str(tmp_path / 'one.parquet')
    pq.write_table(t0, path0, compression='snappy')
    _compare_flat(path0)


def test_compressible_snappy_long_matches(tmp_path):
    # long back-references, overlapping copies (offset < length) and multi-MB literals
    n = 400000
    rng = np.random.default_rng(5)
    t = pa.table({'zeros': np.zeros(n, dtype=np.int64), 'ramp': (np.arange(n) // 3).astype(np.int32),
                  'pattern': np.tile(np.arange(7, dtype=np.int64), n // 7 + 1)[:n],
                  'noise': rng.integers(0, 2 ** 63 - 1, n, dtype=np.int64)})
    path = str(tmp_path / 'comp.parquet')
    pq.write_table(t, path, compression='snappy', use_dictionary=False, data_page_size=8 << 20)
    _compare_flat(path)


def test_c2_shape(tmp_path):
    """BASELINE.json configs[1] shape: 64 float32 + 16 int64 flat columns, Snappy, pyarrow defaults."""
    n = 400000
    rng = np.random.default_rng(1234)
    cols = {}
    for i in range(64):
        cols['f%02d' % i] = rng.standard_normal(n).astype(np.float32)
    for i in range(16):
        cols['i%02d' % i] = rng.integers(0, 2 ** 40, n, dtype=np.int64)
    path = str(tmp_path / 'c2.parquet')
    pq.write_table(pa.table(cols), path, compression='snappy', row_group_size=n)
    _compare_flat(path)
    # latency mode: the 1 MiB dictionary pages of the int64 columns are indexed by four-CTA clusters (k_snappy_index_cluster)
    import os
    from petastorm_b200 import native, rowgroup
    f = rowgroup.open_file(path)
    plan = native.Plan(f, 0, list(range(f.num_columns)))
    assert plan.info.num_cluster_index_pages >= 16
    assert plan.info.num_index_pages > plan.info.num_cluster_index_pages
    os.environ['PST_IDX_CLUSTER'] = '1'
    try:
        _compare_flat(path)
    finally:
        del os.environ['PST_IDX_CLUSTER']


def test_list_column_levels(tmp_path):
    n = 5000
    rng = np.random.default_rng(9)
    lists = [rng.integers(0, 100, 4).tolist() for _ in range(n)]
    t = pa.table({'id': np.arange(n), 'l': pa.array(lists, type=pa.list_(pa.int32()))})
    path = str(tmp_path / 'list.parquet')
    pq.write_table(t, path, compression='snappy')
    f, decoded = _decode_file(path)
    leaf = [i for i, l in enumerate(f.schema['leaves']) if l['name'] == 'l'][0]
    col = decoded[0].column(leaf)
    assert col.max_rep == 1
    rep = col.rep.cpu().numpy()
    defs = col.defs.cpu().numpy()
    vals = col.values.cpu().numpy()
    assert (rep.reshape(n, 4)[:, 0] == 0).all() and (rep.reshape(n, 4)[:, 1:] == 1).all()
    assert (defs == col.max_def).all()
    np.testing.assert_array_equal(vals.reshape(n, 4), np.array(lists, dtype=np.int32))


def test_corrupt_page_is_reported(tmp_path):
    import torch
    from petastorm_b200 import rowgroup
    n = 50000
    rng = np.random.default_rng(1)
    path = str(tmp_path / 'c.parquet')
    pq.write_table(pa.table({'a': rng.integers(0, 1000, n)}), path, compression='snappy', use_dictionary=False)
    dec = rowgroup.RowGroupDecoder()
    plan = dec.plan(path, 0, [0])
    arena = dec.upload(plan, private=True)
    torch.cuda.synchronize()
    arena[100:4000] = 0xff  # trash the first page's snappy stream
    d = dec.decode_resident(plan, arena)
    with pytest.raises(rowgroup.DeviceDecodeError):
        d.check()


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7f
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _literals(data, chunk=60):
    out = bytearray()
    for i in range(0, len(data), chunk):
        piece = data[i:i + chunk]
        out.append((len(piece) - 1) << 2)
        out += piece
    return bytes(out)


def _encode_periodic(img, period, head, wide_offsets):
    """A valid raw-Snappy stream for `img` (periodic with `period` behind its first `head` bytes): literals for the
    head, then 64-byte copies at distance `period` -- 4-byte offsets when `wide_offsets` (a distance the reference
    compressor never produces), 2-byte offsets otherwise."""
    out = bytearray(_varint(len(img)))
    first = min(65536, head)
    out += _literals(img[:first])                      # 1092 x 60 + 16: ends exactly on the 64 KiB boundary
    out += _literals(img[first:head])
    pos = head
    while pos < len(img):
        n = min(64, len(img) - pos)
        if wide_offsets:
            out.append(((n - 1) << 2) | 3)
            out += period.to_bytes(4, 'little')
        else:
            out.append(((n - 1) << 2) | 2)
            out += period.to_bytes(2, 'little')
        pos += n
    return bytes(out)


@pytest.mark.parametrize('case', ['cross_fragment_copies', 'straddling_elements'])
def test_snappy_streams_the_reference_compressor_never_emits(tmp_path, case):
    """The parallel fragment decode relies on the 64 KiB block structure of snappy::RawCompress output.  Streams
    without it are still valid Snappy: back-references that reach into an earlier 64 KiB block (found by the fragment
    kernel, flag 2) and elements that straddle a block boundary (found by the index kernel, flag 1) must take the
    serial fallback and decode to the same values."""
    import torch
    from petastorm_b200 import rowgroup
    from plan_emulator import read_tables
    n = 400000
    period_values, head, wide = (10000, 80000, True) if case == 'cross_fragment_copies' else (4000, 32010, False)
    vals = (np.arange(n, dtype=np.int64) % period_values) * 7919
    schema = pa.schema([pa.field('a', pa.int64(), nullable=False)])
    path = str(tmp_path / 'p.parquet')
    pq.write_table(pa.table({'a': vals}, schema=schema), path, compression='snappy', use_dictionary=False,
                   data_page_size=1 << 20)
    dec = rowgroup.RowGroupDecoder()
    plan = dec.plan(path, 0, [0])
    host, _, pages, pages_off = read_tables(plan)
    arena = dec.upload(plan, private=True)
    torch.cuda.synchronize()
    codec = pa.Codec('snappy')
    patched = 0
    for i, pg in enumerate(pages):
        src_off, _, comp, uncomp = pg[0], pg[1], pg[2], pg[3]
        if pg[11] != 1 or uncomp <= 2 * 65536:
            continue
        img = codec.decompress(bytes(host[src_off:src_off + comp]), decompressed_size=uncomp).to_pybytes()
        stream = _encode_periodic(img, period_values * 8, head, wide)
        assert codec.decompress(stream, decompressed_size=uncomp).to_pybytes() == img
        assert len(stream) <= comp
        arena[src_off:src_off + len(stream)] = torch.frombuffer(bytearray(stream), dtype=torch.uint8).to(arena.device)
        size_field = torch.tensor([len(stream)], dtype=torch.int32).view(torch.uint8).to(arena.device)
        arena[pages_off + i * 64 + 16:pages_off + i * 64 + 20] = size_field
        patched += 1
    assert patched >= 2
    d = dec.decode_resident(plan, arena)
    d.check()
    np.testing.assert_array_equal(d.column(0).values.cpu().numpy(), vals)
