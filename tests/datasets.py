"""Deterministic synthetic datasets shared by the tests, the golden-vector generator (oracle/make_golden.py) and the
parity checks.  Everything is seeded; the same call produces the same decoded values on every machine (the golden
digests are over *decoded* values, so they do not depend on parquet byte layout)."""
import hashlib
import os
from decimal import Decimal

import numpy as np


def _schemas():
    from petastorm_b200 import spark_types as T
    from petastorm_b200.codecs import CompressedImageCodec, NdarrayCodec, ScalarCodec
    from petastorm_b200.unischema import Unischema, UnischemaField
    hello = Unischema('HelloWorldSchema', [
        UnischemaField('id', np.int32, (), ScalarCodec(T.IntegerType()), False),
        UnischemaField('image1', np.uint8, (128, 256, 3), CompressedImageCodec('png'), False),
        UnischemaField('array_4d', np.uint8, (None, 128, 30, None), NdarrayCodec(), False),
    ])
    test = Unischema('TestSchema', [
        UnischemaField('partition_key', np.str_, (), ScalarCodec(T.StringType()), False),
        UnischemaField('id', np.int64, (), ScalarCodec(T.LongType()), False),
        UnischemaField('id2', np.int32, (), ScalarCodec(T.ShortType()), False),
        UnischemaField('id_float', np.float64, (), ScalarCodec(T.DoubleType()), False),
        UnischemaField('id_odd', np.bool_, (), ScalarCodec(T.BooleanType()), False),
        UnischemaField('python_primitive_uint8', np.uint8, (), ScalarCodec(T.ShortType()), False),
        UnischemaField('image_png', np.uint8, (32, 16, 3), CompressedImageCodec('png'), False),
        UnischemaField('matrix', np.float32, (32, 16, 3), NdarrayCodec(), False),
        UnischemaField('decimal', Decimal, (), ScalarCodec(T.DecimalType(10, 9)), False),
        UnischemaField('matrix_uint16', np.uint16, (32, 16, 3), NdarrayCodec(), False),
        UnischemaField('matrix_uint32', np.uint32, (32, 16, 3), NdarrayCodec(), False),
        UnischemaField('matrix_string', np.bytes_, (None, None,), NdarrayCodec(), False),
        UnischemaField('empty_matrix_string', np.bytes_, (None,), NdarrayCodec(), False),
        UnischemaField('matrix_nullable', np.uint16, (32, 16, 3), NdarrayCodec(), True),
        UnischemaField('sensor_name', np.str_, (1,), NdarrayCodec(), False),
        UnischemaField('string_array_nullable', np.str_, (None,), NdarrayCodec(), True),
        UnischemaField('integer_nullable', np.int32, (), ScalarCodec(T.IntegerType()), True),
    ])
    series = Unischema('SeriesSchema', [UnischemaField('ts', np.int64, (), ScalarCodec(T.LongType()), False)] +
                       [UnischemaField('c%02d' % i, np.float32, (), ScalarCodec(T.FloatType()), False) for i in range(12)])
    tensor = Unischema('TensorSchema', [
        UnischemaField('key', np.int32, (), ScalarCodec(T.IntegerType()), False),
        UnischemaField('tensor', np.float16, (8, 16, 16), NdarrayCodec(), False),
    ])
    tensor_c4 = Unischema('TensorC4Schema', [
        UnischemaField('key', np.int32, (), ScalarCodec(T.IntegerType()), False),
        UnischemaField('tensor', np.float16, (32, 128, 128), NdarrayCodec(), False),
    ])
    imagenet = Unischema('ImagenetSchema', [
        UnischemaField('label', np.int32, (), ScalarCodec(T.IntegerType()), False),
        UnischemaField('image', np.uint8, (224, 224, 3), CompressedImageCodec('jpeg', 80), False),
    ])
    return dict(hello=hello, test=test, series=series, tensor=tensor, tensor_c4=tensor_c4, imagenet=imagenet)


def schema(name):
    return _schemas()[name]


def hello_rows(n, seed=1234):
    """examples/hello_world/petastorm_dataset/generate_petastorm_dataset.py:29-51 shapes (C1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:128, 0:256]
    for i in range(n):
        if i % 3 == 2:   # a compressible image exercises dynamic-Huffman blocks, noise exercises stored blocks
            img = np.stack([(xx + i) % 256, (yy * 2 + i) % 256, (xx + yy) % 256], -1).astype(np.uint8)
        else:
            img = rng.integers(0, 255, (128, 256, 3), dtype=np.uint8)
        yield {'id': np.int32(i), 'image1': img, 'array_4d': rng.integers(0, 255, (4, 128, 30, 3), dtype=np.uint8)}


def test_rows(n, seed=99):
    """Modelled on petastorm/tests/test_common.py:71-94 (_randomize_row)."""
    rng = np.random.default_rng(seed)
    for i in range(n):
        yield {
            'partition_key': 'p_%d' % (i % 4),
            'id': np.int64(i),
            'id2': np.int32(i % 2),
            'id_float': np.float64(i),
            'id_odd': np.bool_(i % 2),
            'python_primitive_uint8': np.uint8(i % 255),
            'image_png': rng.integers(0, 255, (32, 16, 3), dtype=np.uint8),
            'matrix': rng.random((32, 16, 3)).astype(np.float32),
            'decimal': Decimal(int(rng.integers(0, 255))) / Decimal(100),
            'matrix_uint16': rng.integers(0, 2 ** 16 - 1, (32, 16, 3)).astype(np.uint16),
            'matrix_uint32': rng.integers(0, 2 ** 32 - 1, (32, 16, 3)).astype(np.uint32),
            'matrix_string': np.asarray([[b'abc%d' % i, b'd'], [b'', b'xyz']]).astype(np.bytes_),
            'empty_matrix_string': np.asarray([], dtype=np.bytes_),
            'matrix_nullable': None if i % 3 == 0 else rng.integers(0, 2 ** 16 - 1, (32, 16, 3)).astype(np.uint16),
            'sensor_name': np.asarray(['test_sensor']),
            'string_array_nullable': None if i % 5 == 0 else np.asarray(['a%d' % i, 'bc'][:1 + i % 2]),
            'integer_nullable': None if i % 2 else np.int32(i),
        }


def series_rows(n, gap_every=37, seed=5):
    rng = np.random.default_rng(seed)
    ts = 0
    for i in range(n):
        ts += 1 if i % gap_every else 10
        row = {'ts': np.int64(ts)}
        for c in range(12):
            row['c%02d' % c] = np.float32(rng.standard_normal())
        yield row


def tensor_rows(n, seed=8):
    rng = np.random.default_rng(seed)
    for i in range(n):
        yield {'key': np.int32(i), 'tensor': rng.standard_normal((8, 16, 16)).astype(np.float16)}


def tensor_c4_rows(n, seed=18):
    """C4 at its real shape: one value = 128-byte .npy header + 1 MiB of float16 (larger than a data page, 17 Snappy
    fragments per value)."""
    rng = np.random.default_rng(seed)
    for i in range(n):
        yield {'key': np.int32(i), 'tensor': rng.standard_normal((32, 128, 128), dtype=np.float32).astype(np.float16)}


def imagenet_rows(n, seed=3):
    import cv2
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:224, 0:224].astype(np.float32)
    for i in range(n):
        a, b, c = rng.uniform(0.5, 3, 3)
        img = np.stack([127 + 120 * np.sin(xx / (20 * a) + i), 127 + 120 * np.cos(yy / (25 * b)),
                        127 + 100 * np.sin((xx + yy) / (30 * c))], -1)
        img = cv2.GaussianBlur(img.astype(np.float32), (0, 0), 2) + rng.normal(0, 3, img.shape)
        yield {'label': np.int32(rng.integers(0, 1000)), 'image': np.clip(img, 0, 255).astype(np.uint8)}


def build(kind, out_dir, n, **kw):
    """Materialise dataset `kind` under out_dir; returns its file:// url."""
    from petastorm_b200.etl.dataset_writer import write_petastorm_dataset
    s = schema(kind)
    rows = {'hello': hello_rows, 'test': test_rows, 'series': series_rows, 'tensor': tensor_rows,
            'tensor_c4': tensor_c4_rows, 'imagenet': imagenet_rows}[kind](n)
    write_petastorm_dataset(out_dir, s, rows, **kw)
    return 'file://' + out_dir


def flat_table(n, seed=1234, ncols_f=8, ncols_i=2, with_extras=True):
    """C2-like plain parquet columns (+ a few logical types)."""
    import pyarrow as pa
    rng = np.random.default_rng(seed)
    cols = {}
    for i in range(ncols_f):
        cols['f%02d' % i] = rng.standard_normal(n).astype(np.float32)
    for i in range(ncols_i):
        cols['i%02d' % i] = rng.integers(0, 2 ** 40, n, dtype=np.int64)
    t = pa.table(cols)
    if with_extras:
        t = t.append_column('key', pa.array(np.arange(n, dtype=np.int32)))
        t = t.append_column('small', pa.array(rng.integers(-100, 100, n).astype(np.int8)))
        t = t.append_column('u16', pa.array(rng.integers(0, 60000, n).astype(np.uint16)))
        t = t.append_column('flag', pa.array(rng.integers(0, 2, n).astype(bool)))
        t = t.append_column('nullable_int', pa.array([None if i % 4 == 0 else i for i in range(n)], type=pa.int64()))
        t = t.append_column('nullable_float', pa.array([None if i % 6 == 0 else i / 3 for i in range(n)], type=pa.float32()))
        t = t.append_column('name', pa.array(['n%d' % (i % 50) for i in range(n)]))
        t = t.append_column('vec', pa.array([rng.integers(0, 9, 3).tolist() for _ in range(n)], type=pa.list_(pa.int32())))
    return t


def write_flat(out_dir, n, files=2, row_group_size=None, partitioned=False, **kw):
    import pyarrow.parquet as pq
    os.makedirs(out_dir, exist_ok=True)
    per = n // files
    for f in range(files):
        t = flat_table(per, seed=1234 + f, **kw)
        d = os.path.join(out_dir, 'part=%d' % f) if partitioned else out_dir
        os.makedirs(d, exist_ok=True)
        pq.write_table(t, os.path.join(d, 'data-%03d.parquet' % f), compression='snappy',
                       row_group_size=row_group_size or per)
    return 'file://' + out_dir


# ---------------------------------------------------------------------------------------------------------------------
# value digests (golden vectors are stored as digests to keep tests/golden small)
# ---------------------------------------------------------------------------------------------------------------------
def digest(value):
    """Stable text digest of a decoded value: type family + dtype + shape + content hash."""
    if value is None:
        return 'None'
    if hasattr(value, 'cpu') and hasattr(value, 'numpy'):
        value = value.cpu().numpy()
    if not isinstance(value, np.ndarray) and hasattr(value, 'to_numpy'):
        # pandas extension arrays (pandas 3 hands string columns back as ArrowStringArray in the predicate path)
        value = np.asarray(value.to_numpy(), dtype=object).astype(np.str_)
    if isinstance(value, Decimal):
        return 'Decimal:' + str(value.normalize())
    if isinstance(value, np.ndarray):
        a = np.ascontiguousarray(value)
        if a.dtype.kind in 'SUO':
            body = repr(a.tolist()).encode()
            return 'arr:{}:{}:{}'.format(a.dtype.kind, a.shape, hashlib.sha1(body).hexdigest()[:16])
        return 'arr:{}:{}:{}'.format(a.dtype.str, a.shape, hashlib.sha1(a.tobytes()).hexdigest()[:16])
    if isinstance(value, (np.generic,)):
        if value.dtype.kind in 'SU':
            return 'np:{}:{!r}'.format(value.dtype.kind, value.item())
        if value.dtype.kind == 'M':
            return 'np:M:' + str(value)
        return 'np:{}:{!r}'.format(value.dtype.str, value.item())
    if isinstance(value, (bytes, bytearray)):
        return 'bytes:' + hashlib.sha1(bytes(value)).hexdigest()[:16]
    return '{}:{!r}'.format(type(value).__name__, value)


def flat_transform(df):
    """TransformSpec.func of the batch-reader scenarios: receives the row-group as a pandas DataFrame
    (petastorm/arrow_reader_worker.py:247-251), edits a column, adds a scalar and a 2-D field."""
    df['f00'] = df['f00'] * np.float32(2)
    df['sum01'] = (df['i00'] + df['i01']).astype(np.int64)
    df['mat'] = df['key'].map(lambda k: np.full((2, 3), k, dtype=np.float32) + np.arange(6, dtype=np.float32).reshape(2, 3))
    return df


def flat_transform_drop(df):
    """Same, for the predicate path: upstream applies func WITHOUT the removed-fields post-processing there
    (arrow_reader_worker.py:342-345), so the function itself returns exactly the transformed schema's columns."""
    df = flat_transform(df.copy())
    del df['i01']
    del df['name']
    del df['mat']     # upstream does not ravel multi-dimensional fields on the predicate path (from_pandas would fail)
    return df


def flat_transform_select(df):
    """selected_fields narrows the schema, not the data: the function returns exactly the selected columns (a result
    with other columns is a ValueError upstream, arrow_reader_worker.py:261-268)."""
    return flat_transform(df.copy())[['sum01', 'mat', 'key']]


FLAT_TRANSFORM_FIELDS = ['key', 'f00', 'i00', 'i01', 'name']
FLAT_TRANSFORM_EDITS = [('f00', np.float32, (), False), ('sum01', np.int64, (), False), ('mat', np.float32, (2, 3), False)]


def ngram_column_digests(windows, offsets, names):
    """{"<offset>.<field>": sha1 over the values of that field at that offset across all windows} + the window count.
    `windows` is an iterable of {offset: row-like}; used for NGram scenarios too large for per-window digests."""
    cols = {(o, n): [] for o in offsets for n in names}
    count = 0
    for w in windows:
        count += 1
        for o in offsets:
            item = w[o]
            d = item._asdict() if hasattr(item, '_asdict') else item
            for n in names:
                cols[(o, n)].append(d[n])
    out = {'windows': count}
    for (o, n), vals in sorted(cols.items()):
        out['%d.%s' % (o, n)] = digest(np.asarray(vals))
    return out


def digest_row(row):
    d = row._asdict() if hasattr(row, '_asdict') else dict(row)
    return {k: digest(v) for k, v in sorted(d.items())}
