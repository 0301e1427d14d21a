"""CPU tests of the host-side logic (no GPU): schema objects, predicates, NGram, ventilator, pools, shuffling buffers,
planner (thrift footer + page walk + HBM layout) against pyarrow, dataset discovery, sharding broadcast over gloo.
Modelled on the reference's unit tests."""
import hashlib
import io
import os
import sys
import threading
import time
from decimal import Decimal

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import datasets
from helpers import GOLDEN


# ---------------------------------------------------------------------------------------------------------------------
# unischema / transform  (petastorm/tests/test_unischema.py, test_transform.py)
# ---------------------------------------------------------------------------------------------------------------------
def _schema():
    from petastorm_b200.codecs import NdarrayCodec, ScalarCodec
    from petastorm_b200 import spark_types as T
    from petastorm_b200.unischema import Unischema, UnischemaField
    return Unischema('S', [UnischemaField('int_field', np.int8, (), ScalarCodec(T.ByteType()), False),
                           UnischemaField('string_field', np.str_, (), ScalarCodec(T.StringType()), False),
                           UnischemaField('other_string_field', np.str_, (), ScalarCodec(T.StringType()), False),
                           UnischemaField('mat', np.float32, (3, None), NdarrayCodec(), True)])


def test_unischema_views_and_equality():
    from petastorm_b200.codecs import ScalarCodec
    from petastorm_b200 import spark_types as T
    from petastorm_b200.unischema import UnischemaField
    s = _schema()
    assert list(s.fields.keys()) == ['int_field', 'string_field', 'other_string_field', 'mat']
    assert s.int_field.numpy_dtype is np.int8
    v = s.create_schema_view([s.int_field, 'other.*$'])
    assert set(v.fields.keys()) == {'int_field', 'other_string_field'}
    # lookup is by name: codec/shape of the argument are ignored (unischema.py:235-237)
    v2 = s.create_schema_view([UnischemaField('mat', np.int64, (), None, False)])
    assert v2.fields['mat'].shape == (3, None)
    with pytest.raises(ValueError):
        s.create_schema_view([UnischemaField('nope', np.int8, (), None, False)])
    with pytest.raises(ValueError):
        s.create_schema_view([42])
    # full-match regex; legacy prefix semantics only warn
    with pytest.warns(UserWarning):
        v3 = s.create_schema_view(['string'])
    assert len(v3.fields) == 0
    a = UnischemaField('a', np.int32, (), ScalarCodec(T.IntegerType()), False)
    b = UnischemaField('a', np.int32, (), None, False)
    assert a == b and hash(a) == hash(b) and a != UnischemaField('a', np.int64, (), None, False)
    nt = s.make_namedtuple(int_field=1, string_field='x', other_string_field='y', mat=None)
    assert nt.int_field == 1 and nt.mat is None
    assert type(nt) is type(s.make_namedtuple(int_field=2, string_field='', other_string_field='', mat=None))


def test_field_order_switch(monkeypatch):
    from petastorm_b200 import unischema
    from petastorm_b200.unischema import Unischema, UnischemaField
    fields = [UnischemaField('b', np.int32, ()), UnischemaField('a', np.int32, ())]
    assert list(Unischema('o1', fields).fields.keys()) == ['b', 'a']
    monkeypatch.setattr(unischema, '_UNISCHEMA_FIELD_ORDER', 'alphabetical')
    assert list(Unischema('o2', fields).fields.keys()) == ['a', 'b']


def test_transform_schema():
    from petastorm_b200.transform import TransformSpec, edit_field, transform_schema
    s = _schema()
    t = transform_schema(s, TransformSpec(lambda x: x, edit_fields=[edit_field('mat', np.float16, (3, 4), False),
                                                                   ('new', np.int32, (), False)],
                                           removed_fields=['string_field']))
    assert list(t.fields.keys()) == ['int_field', 'other_string_field', 'mat', 'new']
    assert t.fields['mat'].codec is None and t.fields['mat'].numpy_dtype is np.float16
    t2 = transform_schema(s, TransformSpec(selected_fields=['mat', 'int_field']))
    assert list(t2.fields.keys()) == ['mat', 'int_field']
    with pytest.raises(ValueError):
        TransformSpec(removed_fields=['a'], selected_fields=['b'])
    with pytest.warns(UserWarning):
        transform_schema(s, TransformSpec(removed_fields=['ghost']))


def test_infer_schema_from_parquet(tmp_path):
    from petastorm_b200 import native
    from petastorm_b200.unischema import Unischema
    path = str(tmp_path / 'f.parquet')
    t = datasets.flat_table(50)
    t = t.append_column('dec', pa.array([Decimal('1.50')] * 50, type=pa.decimal128(10, 2)))
    t = t.append_column('ts', pa.array(np.arange(50), type=pa.timestamp('ms')))
    t = t.append_column('raw', pa.array([b'x'] * 50, type=pa.binary()))
    pq.write_table(t, path)
    with pytest.warns(UserWarning, match='u16'):
        s = Unischema.from_parquet_schema(native.ParquetFile(path).schema, [('part', np.int64)])
    f = s.fields
    assert list(f.keys())[0] == 'part' and 'u16' not in f
    assert f['f00'].numpy_dtype is np.float32 and f['i00'].numpy_dtype is np.int64 and f['small'].numpy_dtype is np.int8
    assert f['flag'].numpy_dtype is np.bool_ and f['name'].numpy_dtype is np.str_ and f['raw'].numpy_dtype is np.bytes_
    assert f['dec'].numpy_dtype is Decimal and f['ts'].numpy_dtype is np.datetime64
    assert f['vec'].shape == (None,) and f['vec'].numpy_dtype is np.int32
    assert f['nullable_int'].nullable and all(x.codec is None for x in f.values())


# ---------------------------------------------------------------------------------------------------------------------
# predicates  (petastorm/tests/test_predicates.py)
# ---------------------------------------------------------------------------------------------------------------------
def test_predicates_host_semantics():
    from petastorm_b200.predicates import (PredicateBase, _string_to_bucket, in_intersection, in_lambda, in_negate,
                                           in_pseudorandom_split, in_reduce, in_set)
    assert in_set({1, 2}, 'a').do_include({'a': np.int32(2)}) and not in_set({1, 2}, 'a').do_include({'a': 3})
    assert in_set({'x'}, 'a').get_fields() == {'a'}
    assert in_intersection([1, 5], 'l').do_include({'l': [5, 9]}) and not in_intersection([1], 'l').do_include({'l': [2]})
    with pytest.raises(ValueError):
        in_intersection([1], 'l').do_include({'l': 3})
    assert in_lambda(['a', 'b'], lambda a, b: a + b == 3).do_include({'a': 1, 'b': 2})
    assert in_lambda(['a'], lambda a, st: a in st, state_arg={7}).do_include({'a': 7})
    with pytest.raises(ValueError):
        in_lambda('a', lambda a: True)
    assert in_negate(in_set({1}, 'a')).do_include({'a': 2})
    with pytest.raises(ValueError):
        in_negate('nope')
    both = in_reduce([in_set({1, 2}, 'a'), in_set({2, 3}, 'b')], all)
    assert both.get_fields() == {'a', 'b'} and both.do_include({'a': 2, 'b': 2}) and not both.do_include({'a': 1, 'b': 1})
    # known answer of the md5 bucket (petastorm/predicates.py:39-41)
    assert _string_to_bucket('abc', 1000) == int(hashlib.md5(b'abc').hexdigest(), 16) % 1000
    with pytest.raises(ValueError):
        in_pseudorandom_split([0.5, 0.5], 2, 'k')
    splits = [in_pseudorandom_split([0.3, 0.4, 0.3], i, 'k') for i in range(3)]
    for key in range(200):
        assert sum(p.do_include({'k': key}) for p in splits) == 1
    assert issubclass(in_set, PredicateBase)


# ---------------------------------------------------------------------------------------------------------------------
# NGram  (petastorm/tests/test_ngram.py, test_ngram_end_to_end.py)
# ---------------------------------------------------------------------------------------------------------------------
def test_ngram_host_known_answers():
    from petastorm_b200.ngram import NGram
    from petastorm_b200.unischema import Unischema, UnischemaField
    s = Unischema('T', [UnischemaField('id', np.int64, ()), UnischemaField('v', np.float32, ()),
                        UnischemaField('w', np.float32, ())])
    ng = NGram({-1: [s.id, s.v], 0: [s.id, s.w]}, delta_threshold=4, timestamp_field=s.id)
    assert ng.length == 2
    rows = [{'id': i, 'v': float(i), 'w': -float(i)} for i in [0, 3, 8, 10, 11, 20, 30]]
    res = ng.form_ngram(rows, s)
    # documentation example petastorm/ngram.py:54-83: (0,3), (8,10), (10,11)
    assert [(r[-1]['id'], r[0]['id']) for r in res] == [(0, 3), (8, 10), (10, 11)]
    assert set(res[0][-1].keys()) == {'id', 'v'} and set(res[0][0].keys()) == {'id', 'w'}
    ng2 = NGram({0: [s.id], 1: [s.id], 2: [s.id]}, delta_threshold=1, timestamp_field=s.id, timestamp_overlap=False)
    rows = [{'id': i} for i in range(8)]
    assert [r[0]['id'] for r in ng2.form_ngram(rows, s)] == [0, 3]
    with pytest.raises(NotImplementedError):
        ng.form_ngram([{'id': 5, 'v': 0, 'w': 0}, {'id': 3, 'v': 0, 'w': 0}], s)
    for bad in (dict(fields=None, delta_threshold=1, timestamp_field=s.id),
                dict(fields={0: s.id}, delta_threshold=1, timestamp_field=s.id),
                dict(fields={0: [s.id]}, delta_threshold='x', timestamp_field=s.id),
                dict(fields={0: [s.id]}, delta_threshold=1, timestamp_field=None),
                dict(fields={0: [s.id]}, delta_threshold=1, timestamp_field=s.id, timestamp_overlap=None)):
        with pytest.raises(ValueError):
            NGram(**bad)
    ng3 = NGram({0: ['i.*'], 1: ['v']}, 1, 'id')
    ng3.resolve_regex_field_names(s)
    assert ng3.get_field_names_at_timestep(0) == ['id'] and ng3.timestamp_field.name == 'id'
    assert ng3.get_field_names_at_timestep(5) == []
    assert set(f.name for f in ng3.get_field_names_at_all_timesteps()) == {'id', 'v'}


# ---------------------------------------------------------------------------------------------------------------------
# ventilator / pool  (workers_pool/tests/test_ventilator.py, test_workers_pool.py)
# ---------------------------------------------------------------------------------------------------------------------
def test_ventilator_order_epochs_and_backpressure():
    from petastorm_b200.workers_pool.ventilator import ConcurrentVentilator
    got = []
    items = [{'x': i} for i in range(10)]
    v = ConcurrentVentilator(lambda x: got.append(x), items, iterations=2, randomize_item_order=True, random_seed=5,
                             max_ventilation_queue_size=100)
    v.start()
    while not v.completed():
        time.sleep(0.01)
    v.stop()
    exp = list(np.random.default_rng(5).permutation(10))
    assert got == [int(i) for i in exp] * 2          # one permutation per start(), repeated every epoch
    # back-pressure: never more than max_ventilation_queue_size unprocessed items
    got2 = []
    v2 = ConcurrentVentilator(lambda x: got2.append(x), items, iterations=1, max_ventilation_queue_size=3)
    v2.start()
    time.sleep(0.1)
    assert len(got2) == 3
    for _ in range(3):
        v2.processed_item()
    time.sleep(0.1)
    assert len(got2) == 6
    with pytest.raises(NotImplementedError):
        v2.reset()
    v2.stop()
    with pytest.raises(ValueError):
        ConcurrentVentilator(print, items, iterations=0)
    with pytest.raises(ValueError):
        ConcurrentVentilator(print, [1, 2])
    v3 = ConcurrentVentilator(print, [], iterations=None)
    assert v3.completed()


class _StubWorker(object):
    def __init__(self, worker_id, publish, args):
        self.publish = publish
        self.args = args

    def process(self, value, fail=False, many=1):
        if fail:
            raise ValueError('worker failure %d' % value)
        for k in range(many):
            self.publish(value * 10 + k)

    def shutdown(self):
        pass


@pytest.mark.parametrize('sync', [False, True])
def test_gpu_pool_protocol(sync):
    from petastorm_b200.workers_pool import EmptyResultError
    from petastorm_b200.workers_pool.gpu_pool import GpuPool
    from petastorm_b200.workers_pool.ventilator import ConcurrentVentilator
    pool = GpuPool(synchronous=sync, results_queue_size=2)
    vent = ConcurrentVentilator(pool.ventilate, [{'value': i, 'many': 2} for i in range(6)], iterations=1,
                                max_ventilation_queue_size=4)
    pool.start(_StubWorker, None, vent)
    out = []
    with pytest.raises(EmptyResultError):
        while True:
            out.append(pool.get_results())
    assert out == [i * 10 + k for i in range(6) for k in range(2)]      # FIFO, in ventilation order
    pool.stop()
    pool.join()
    with pytest.raises(RuntimeError):
        pool.start(_StubWorker, None, None)


def test_gpu_pool_worker_exception_and_stop_with_full_queue():
    from petastorm_b200.workers_pool.gpu_pool import GpuPool
    pool = GpuPool(results_queue_size=1)
    pool.start(_StubWorker, None, None)
    pool.ventilate(value=1, fail=True)
    with pytest.raises(ValueError, match='worker failure 1'):
        pool.get_results()
    # stop() must not dead-lock while the worker is blocked on a full results queue
    pool2 = GpuPool(results_queue_size=1)
    pool2.start(_StubWorker, None, None)
    pool2.ventilate(value=1, many=50)
    time.sleep(0.2)
    t0 = time.time()
    pool2.stop()
    pool2.join()
    assert time.time() - t0 < 5


# ---------------------------------------------------------------------------------------------------------------------
# shuffling buffers  (petastorm/tests/test_shuffling_buffer.py)
# ---------------------------------------------------------------------------------------------------------------------
def test_shuffling_buffer_state_machines():
    import torch
    from petastorm_b200.reader_impl.pytorch_shuffling_buffer import (BatchedNoopShufflingBuffer,
                                                                     BatchedRandomShufflingBuffer)
    from petastorm_b200.reader_impl.shuffling_buffer import NoopShufflingBuffer, RandomShufflingBuffer
    b = NoopShufflingBuffer()
    assert not b.can_retrieve() and b.can_add()
    b.add_many([1, 2, 3])
    assert [b.retrieve() for _ in range(3)] == [1, 2, 3] and b.size == 0
    r = RandomShufflingBuffer(10, 3)
    r.add_many([1, 2])
    assert not r.can_retrieve()
    r.add_many([3, 4, 5])
    assert r.can_retrieve() and r.size == 5
    got = [r.retrieve() for _ in range(3)]
    assert not r.can_retrieve()
    with pytest.raises(RuntimeError):
        r.retrieve()
    r.finish()
    while r.can_retrieve():
        got.append(r.retrieve())
    assert sorted(got) == [1, 2, 3, 4, 5]
    with pytest.raises(RuntimeError):
        r.add_many([1])
    r2 = RandomShufflingBuffer(2, 1, extra_capacity=2)
    r2.add_many([1, 2])
    assert not r2.can_add()
    with pytest.raises(RuntimeError):
        r2.add_many([3])
    r3 = RandomShufflingBuffer(2, 1, extra_capacity=2)
    with pytest.raises(RuntimeError):
        r3.add_many(list(range(5)))
    # batched variants on CPU tensors: stream through and compare multisets
    for buf in (BatchedNoopShufflingBuffer(batch_size=4), BatchedRandomShufflingBuffer(10, 3, batch_size=4)):
        seen = []
        k = 0
        for chunk in range(6):
            if buf.can_add():
                ids = torch.arange(k, k + 7)
                buf.add_many([ids, ids.float().unsqueeze(1).repeat(1, 3)])
                k += 7
            while buf.can_retrieve():
                a, bm = buf.retrieve()
                assert torch.equal(bm[:, 0].long(), a)
                seen.extend(a.tolist())
        buf.finish()
        while buf.can_retrieve():
            a, _ = buf.retrieve()
            seen.extend(a.tolist())
        assert sorted(seen) == list(range(k))


def test_sanitize_and_collate_tables():
    """petastorm/tests/test_pytorch_dataloader.py:84-149"""
    import torch
    from petastorm_b200.pytorch import _sanitize_pytorch_types, decimal_friendly_collate
    row = {'u16': np.zeros(3, np.uint16), 'u32': np.zeros(3, np.uint32), 'b': np.zeros(3, np.bool_),
           'bs': np.bool_(True), 'i8': np.zeros(3, np.int8), 'f': np.zeros(3, np.float32),
           't16': torch.zeros(2, dtype=torch.uint16), 'tb': torch.zeros(2, dtype=torch.bool)}
    _sanitize_pytorch_types(row)
    assert row['u16'].dtype == np.int32 and row['u32'].dtype == np.int64 and row['b'].dtype == np.uint8
    assert row['bs'].dtype == np.uint8 and row['i8'].dtype == np.int8 and row['f'].dtype == np.float32
    assert row['t16'].dtype == torch.int32 and row['tb'].dtype == torch.uint8
    for bad in ({'s': np.array(['a'])}, {'s': np.array([b'a'])}, {'o': np.array([None], dtype=object)}, {'n': None}):
        with pytest.raises(TypeError):
            _sanitize_pytorch_types(bad)
    out = decimal_friendly_collate([{'a': np.int32(1), 'd': Decimal('1.5'), 's': 'x', 'l': [1, Decimal(2)]},
                                    {'a': np.int32(2), 'd': Decimal('2.5'), 's': 'y', 'l': [3, Decimal(4)]}])
    assert out['a'].tolist() == [1, 2] and out['d'] == [Decimal('1.5'), Decimal('2.5')] and out['s'] == ['x', 'y']
    assert out['l'][0].tolist() == [1, 3] and list(out['l'][1]) == [Decimal(2), Decimal(4)]


# ---------------------------------------------------------------------------------------------------------------------
# native host side: footer, planner, dataset discovery, legacy schema
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('compression,version,dictionary', [('snappy', '1.0', True), ('none', '2.0', True),
                                                            ('snappy', '2.0', False), ('gzip', '1.0', True),
                                                            ('gzip', '2.0', False)])
def test_planner_tables_decode_like_pyarrow(tmp_path, compression, version, dictionary):
    """Thrift footer parse + page-header walk + HBM layout: the raw image and device tables of a plan, interpreted by a
    slow numpy emulator, must reproduce pyarrow's values for every column (nulls, dictionaries, booleans, strings)."""
    import plan_emulator
    from petastorm_b200 import native
    t = datasets.flat_table(3001, with_extras=True)
    path = str(tmp_path / 'p.parquet')
    pq.write_table(t, path, compression=compression, data_page_version=version, use_dictionary=dictionary,
                   row_group_size=1700, data_page_size=4096)
    f = native.ParquetFile(path)
    pf = pq.ParquetFile(path)
    assert f.num_rows == 3001 and f.num_row_groups == pf.metadata.num_row_groups
    for rg in range(f.num_row_groups):
        plan = native.Plan(f, rg, list(range(f.num_columns)))
        info = plan.info
        assert info.num_rows == pf.metadata.row_group(rg).num_rows
        assert info.raw_bytes % 256 == 0 and info.arena_bytes >= info.raw_bytes
        res = plan_emulator.decode_plan(plan, native)
        tbl = pf.read_row_group(rg)
        for slot, leaf in enumerate(f.schema['leaves']):
            got = res[slot]['values']
            if leaf['max_rep']:
                exp = [x for row in tbl.column(leaf['name']).to_pylist() for x in row]
                got = [g.item() for g in got]
                assert got == exp
                continue
            exp = tbl.column(leaf['name']).to_pylist()
            if res[slot]['ptype'] == 6:
                got = [g.decode() if g is not None else None for g in got]
            got = [None if g is None else (g.item() if hasattr(g, 'item') else g) for g in got]
            assert got == exp, leaf['name']


def test_footer_metadata_matches_pyarrow(tmp_path):
    from petastorm_b200 import native
    path = str(tmp_path / 'm.parquet')
    t = datasets.flat_table(100).replace_schema_metadata({'k1': 'v1', 'bin': b'\x00\x01\xff'})
    pq.write_table(t, path, row_group_size=30)
    f = native.ParquetFile(path)
    md = pq.ParquetFile(path).metadata
    kv = f.key_value_metadata()
    assert kv[b'k1'] == b'v1' and kv[b'bin'] == b'\x00\x01\xff'
    assert f.num_row_groups == md.num_row_groups == 4
    for rg in range(4):
        assert f.row_group_num_rows(rg) == md.row_group(rg).num_rows
        for c in range(f.num_columns):
            ci, m = f.chunk_info(rg, c), md.row_group(rg).column(c)
            assert (ci.num_values, ci.total_compressed_size, ci.total_uncompressed_size, ci.data_page_offset) == \
                (m.num_values, m.total_compressed_size, m.total_uncompressed_size, m.data_page_offset)
    leaf = {l['name']: l for l in f.schema['leaves']}
    assert leaf['vec']['max_rep'] == 1 and leaf['vec']['max_def'] == 3 and leaf['f00']['max_def'] == 1


def test_dataset_discovery_and_legacy_schema():
    from petastorm_b200.codecs import CompressedImageCodec, NdarrayCodec, ScalarCodec
    from petastorm_b200.etl import dataset_metadata as dm
    for version in ('0.4.0', '0.7.6'):
        ds = dm.ParquetDataset(os.path.join(GOLDEN, 'legacy', version))
        assert len(ds.pieces) == 10 and ds.partitions.partition_names == {'partition_key'}
        assert ds.partitions.dtype_of('partition_key') is np.str_
        rgs = dm.load_row_groups(ds)
        assert len(rgs) == 10 and [p.path for p in rgs] == sorted(p.path for p in rgs)
        assert rgs[0].partition_keys == [('partition_key', 'p_0')]
        s = dm.get_schema(ds)
        assert isinstance(s.fields['image_png'].codec, CompressedImageCodec)
        assert s.fields['image_png'].codec.image_codec == 'png' and s.fields['image_png'].shape == (32, 16, 3)
        assert isinstance(s.fields['matrix'].codec, NdarrayCodec) and isinstance(s.fields['id'].codec, ScalarCodec)
        assert s.fields['decimal'].numpy_dtype is Decimal and s.fields['partition_key'].numpy_dtype is np.str_


def test_dataset_without_metadata_and_errors(tmp_path):
    from petastorm_b200.errors import PetastormMetadataError
    from petastorm_b200.etl import dataset_metadata as dm
    url = datasets.write_flat(str(tmp_path / 'flat'), 300, files=3, row_group_size=40, partitioned=True)
    ds = dm.ParquetDataset(url[7:])
    assert ds.common_metadata is None and ds.partitions.dtype_of('part') is np.int64
    rgs = dm.load_row_groups(ds)
    assert len(rgs) == 9 and rgs[0].row_group == 0 and rgs[2].row_group == 2
    with pytest.raises(PetastormMetadataError):
        dm.get_schema(ds)
    s = dm.infer_or_load_unischema(ds)
    assert list(s.fields.keys())[0] == 'part' and s.fields['part'].numpy_dtype is np.int64
    with pytest.raises(IOError):
        dm.ParquetDataset(str(tmp_path / 'missing'))
    from petastorm_b200.fs_utils import get_filesystem_and_path_or_paths
    with pytest.raises(ValueError, match='local files only'):
        get_filesystem_and_path_or_paths('hdfs://nn/path')
    with pytest.raises(ValueError):
        get_filesystem_and_path_or_paths(['file:///a', 's3://b'])


def test_restricted_unpickler_refuses_foreign_globals():
    import pickle
    from petastorm_b200.etl.legacy import restricted_loads
    with pytest.raises(pickle.UnpicklingError):
        restricted_loads(pickle.dumps(os.system))
    with pytest.raises(pickle.UnpicklingError):
        restricted_loads(b"cbuiltins\neval\n(S'1'\ntR.")
    assert restricted_loads(pickle.dumps({'a': (1, 2)})) == {'a': (1, 2)}


def test_written_dataset_is_readable_by_its_own_metadata(tmp_path):
    """Write side: the pickled Unischema uses the reference's module names."""
    from petastorm_b200.etl import dataset_metadata as dm
    url = datasets.build('tensor', str(tmp_path / 'ds'), 10, row_group_rows=4)
    ds = dm.ParquetDataset(url[7:])
    raw = ds.common_metadata[dm.UNISCHEMA_KEY]
    assert b'cpetastorm.unischema\n' in raw and b'petastorm_b200' not in raw
    s = dm.get_schema(ds)
    assert list(s.fields.keys()) == ['key', 'tensor'] and s.fields['tensor'].shape == (8, 16, 16)
    assert len(dm.load_row_groups(ds)) == 3


def test_npy_header_and_codecs():
    from petastorm_b200.codecs import (CompressedImageCodec, CompressedNdarrayCodec, NdarrayCodec, ScalarCodec,
                                       _is_compliant_shape, parse_npy_header)
    from petastorm_b200 import spark_types as T
    from petastorm_b200.unischema import UnischemaField
    for arr in (np.zeros((32, 128, 128), np.float16), np.arange(6, dtype=np.int64).reshape(2, 3),
                np.asarray([b'ab', b'c']), np.zeros((0,), np.float32)):
        m = io.BytesIO()
        np.save(m, arr)
        dtype, shape, fortran, off = parse_npy_header(m.getvalue())
        assert dtype == arr.dtype and shape == arr.shape and not fortran and off % 64 == 0
        assert len(m.getvalue()) == off + arr.nbytes
    f = UnischemaField('m', np.float32, (2, None), NdarrayCodec(), False)
    a = np.ones((2, 5), np.float32)
    np.testing.assert_array_equal(NdarrayCodec().decode(f, NdarrayCodec().encode(f, a)), a)
    np.testing.assert_array_equal(CompressedNdarrayCodec().decode(f, CompressedNdarrayCodec().encode(f, a)), a)
    with pytest.raises(ValueError):
        NdarrayCodec().encode(f, a.astype(np.float64))
    with pytest.raises(ValueError):
        NdarrayCodec().encode(f, np.ones((3, 5), np.float32))
    img = UnischemaField('i', np.uint8, (4, 6, 3), CompressedImageCodec('png'), False)
    x = np.random.default_rng(0).integers(0, 255, (4, 6, 3), dtype=np.uint8)
    np.testing.assert_array_equal(CompressedImageCodec('png').decode(img, CompressedImageCodec('png').encode(img, x)), x)
    assert str(CompressedImageCodec('jpeg', 70)) == "CompressedImageCodec('jpeg', 70)"
    assert ScalarCodec(T.IntegerType()).decode(UnischemaField('s', np.int32, ()), 5) == np.int32(5)
    assert _is_compliant_shape((1, 2, 3), (1, None, 3)) and not _is_compliant_shape((1, 2), (1,))


# ---------------------------------------------------------------------------------------------------------------------
# multi-process: row-group owner table broadcast, world_size 2 over gloo
# ---------------------------------------------------------------------------------------------------------------------
def _shard_worker(rank, world, port, url, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from petastorm_b200 import sharding
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    kwargs = sharding.sharded_reader_kwargs(url)
    r, w, owners = sharding.broadcast_row_group_assignment(9)
    # a rank whose listing differs from rank 0's: every rank raises before the table broadcast (no hang)
    try:
        sharding.broadcast_row_group_assignment(9 + rank)
        mismatch = 'no error'
    except RuntimeError as e:
        mismatch = 'raised' if 'same dataset' in str(e) else repr(e)
    with open(os.path.join(out_dir, 'rank%d.txt' % rank), 'w') as f:
        f.write('%r|%r|%r|%s' % (kwargs, (r, w), owners.tolist(), mismatch))
    dist.destroy_process_group()


def test_shard_assignment_broadcast_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    url = datasets.write_flat(str(tmp_path / 'flat'), 300, files=3, row_group_size=40)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_shard_worker, args=(2, port, url, str(tmp_path)), nprocs=2, join=True)
    a = open(str(tmp_path / 'rank0.txt')).read().split('|')
    b = open(str(tmp_path / 'rank1.txt')).read().split('|')
    assert a[0] == "{'cur_shard': 0, 'shard_count': 2}" and b[0] == "{'cur_shard': 1, 'shard_count': 2}"
    assert a[2] == b[2] == repr([i % 2 for i in range(9)])
    assert a[3] == b[3] == 'raised'
    from petastorm_b200 import sharding
    assert sharding.shard_order(9, 2, 1) == [1, 3, 5, 7]
    assert sorted(sharding.shard_order(9, 2, 0, seed=3)) == [0, 2, 4, 6, 8]


def test_partition_filters_follow_legacy_pyarrow_semantics(tmp_path):
    """``filters`` prune FILES by hive partition key only (what the reference's legacy pq.ParquetDataset did,
    petastorm/reader.py:430-433; reference tests test_end_to_end.py:917-937)."""
    from petastorm_b200.etl import dataset_metadata as dm
    url = datasets.write_flat(str(tmp_path / 'flat'), 300, files=3, row_group_size=40, partitioned=True)
    path = url[7:]
    parts = lambda ds: sorted({int(dict(p.partition_keys)['part']) for p in ds.pieces})
    assert parts(dm.ParquetDataset(path)) == [0, 1, 2]
    assert parts(dm.ParquetDataset(path, filters=[('part', '=', 1)])) == [1]
    assert parts(dm.ParquetDataset(path, filters=[('part', '!=', 1)])) == [0, 2]
    assert parts(dm.ParquetDataset(path, filters=[('part', '>=', 1), ('part', '<', 2)])) == [1]
    assert parts(dm.ParquetDataset(path, filters=[[('part', '=', 0)], [('part', '=', 2)]])) == [0, 2]
    assert parts(dm.ParquetDataset(path, filters=[('part', 'in', {0, 2})])) == [0, 2]
    assert parts(dm.ParquetDataset(path, filters=[('part', 'not in', [0])])) == [1, 2]
    assert parts(dm.ParquetDataset(path, filters=[('part', '=', '2')])) == [2]          # value type decides the cast
    assert parts(dm.ParquetDataset(path, filters=[('f0', '>', 1e9)])) == [0, 1, 2]      # not a partition key: ignored
    assert parts(dm.ParquetDataset(path, filters=[('part', '=', 7)])) == []
    with pytest.raises(ValueError):
        dm.ParquetDataset(path, filters=[('part', '~', 1)])
    with pytest.raises(TypeError):
        dm.ParquetDataset(path, filters=[('part', 'in', 1)])


def test_planner_resolves_literal_only_snappy_pages_on_the_host(tmp_path):
    """Incompressible pages leave snappy::RawCompress as literal elements only (one per 64 KiB block): the stream is
    framed, not compressed.  The planner records where the literal bytes lie and the staging copy lays them down back to
    back, so the device receives an ordinary uncompressed page image; only compressible pages are left for the Snappy
    kernels.  The raw image + the copy tiles, emulated here with numpy, must reproduce the column values."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from petastorm_b200 import native
    n = 300000
    rng = np.random.default_rng(0)
    path = str(tmp_path / 'x.parquet')
    cols = {'noise': rng.integers(0, 2 ** 63 - 1, n, dtype=np.int64),
            'narrow': rng.integers(0, 2 ** 20, n, dtype=np.int64)}
    pq.write_table(pa.table(cols), path, compression='snappy', use_dictionary=False, data_page_size=1 << 20)
    f = native.ParquetFile(path)
    plan = native.Plan(f, 0, [0])
    noise = plan.info
    narrow = native.Plan(f, 0, [1]).info
    assert noise.num_unwrapped_pages >= 2 and noise.num_compressed_pages == 0 and noise.num_snappy_fragments == 0
    assert narrow.num_compressed_pages >= 2 and narrow.num_index_pages == narrow.num_compressed_pages
    assert narrow.num_unwrapped_pages == 0
    # every page of `noise` is PLAIN without nulls: all values travel through copy tiles, none through the page decoder
    assert noise.num_decode_pages == 0 and noise.num_copy_tiles >= noise.num_unwrapped_pages
    assert plan.cols[0].valid_off == -1          # chunk statistics: null_count == 0 -> no validity array
    raw = np.zeros(noise.arena_bytes, dtype=np.uint8)
    plan.fill_raw(raw.ctypes.data)
    out = np.zeros(noise.out_bytes, dtype=np.uint8)
    for i in range(noise.num_copy_tiles):
        t = plan.copy_tile(i)
        assert t.src_off % 16 == 0 and t.nbytes <= 65536 and t.valid_off == -1
        out[t.dst_off:t.dst_off + t.nbytes] = raw[t.src_off:t.src_off + t.nbytes]
    pc = plan.cols[0]
    np.testing.assert_array_equal(out[pc.values_off:pc.values_off + 8 * n].view(np.int64), cols['noise'])


def test_planner_copy_tiles_v2_pages_and_validity_arrays(tmp_path):
    """DATA_PAGE_V2 (levels stored uncompressed in front of the values), columns without statistics (validity array
    kept: the tiles carry the validity bytes to set) and a column with real nulls (general page decoder)."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from petastorm_b200 import native
    n = 70000
    rng = np.random.default_rng(1)
    vals = rng.standard_normal(n).astype(np.float32)
    holes = vals.astype(object)
    holes[::7] = None
    table = pa.table({'x': vals, 'h': pa.array(list(holes), type=pa.float32())})
    for version, stats, codec in (('2.0', True, 'snappy'), ('1.0', False, 'none'), ('2.0', False, 'snappy')):
        path = str(tmp_path / 'v{}_{}_{}.parquet'.format(version[0], int(stats), codec))
        pq.write_table(table, path, compression=codec, use_dictionary=False, data_page_version=version,
                       write_statistics=stats, data_page_size=64 << 10)
        f = native.ParquetFile(path)
        plan = native.Plan(f, 0, [0, 1])
        info = plan.info
        x, h = plan.cols
        assert (x.valid_off == -1) == stats and h.valid_off >= 0
        raw = np.zeros(info.arena_bytes, dtype=np.uint8)
        plan.fill_raw(raw.ctypes.data)
        out = np.zeros(info.out_bytes, dtype=np.uint8)
        pages = [plan.page(i) for i in range(info.num_pages)]
        assert all(pg.flags & 2 for pg in pages if pg.column_slot == 0)          # x: all through copy tiles
        assert not any(pg.flags & 2 for pg in pages if pg.column_slot == 1)      # h has nulls: page decoder
        assert all(pg.codec == 0 for pg in pages if pg.column_slot == 0)         # random floats: literal-only or stored
        covered = 0
        for i in range(info.num_copy_tiles):
            t = plan.copy_tile(i)
            out[t.dst_off:t.dst_off + t.nbytes] = raw[t.src_off:t.src_off + t.nbytes]
            covered += t.nbytes
            if not stats:
                assert t.valid_off >= 0 and t.nvalid * 4 == t.nbytes
                out[t.valid_off:t.valid_off + t.nvalid] = 1
        assert covered == 4 * n
        np.testing.assert_array_equal(out[x.values_off:x.values_off + 4 * n].view(np.float32), vals)
        if not stats:
            assert out[x.valid_off:x.valid_off + n].all()


def _index_fixture_pieces():
    pieces = []
    for p in range(6):
        rows = []
        for i in range(5):
            k = p * 5 + i
            rows.append({'id': np.int64(k), 'tag': 'tag_%d' % (k % 4), 'sensor': None if p % 2 else 's%d' % (k % 3),
                         'vec': np.array([k % 3, 7 + p], dtype=np.int32)})
        pieces.append(rows)
    return pieces


def test_rowgroup_indexers_match_the_reference_classes():
    """tests/golden/rowgroup_index.json was pickled by the reference's own SingleFieldIndexer / FieldNotNullIndexer
    (oracle/make_golden_index.py): it must un-pickle onto this package's classes, answer the same look-ups, and this
    package's indexers must build the same index from the same rows and pickle it under the reference's module name."""
    import json
    import pickletools
    from petastorm_b200.etl.legacy import restricted_loads
    from petastorm_b200.etl.rowgroup_indexers import (FieldNotNullIndexer, SingleFieldIndexer,
                                                      pickle_indexers_reference_compatible)
    from petastorm_b200.selectors import IntersectIndexSelector, SingleIndexSelector, UnionIndexSelector
    fx = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'rowgroup_index.json')))
    ref = restricted_loads(bytes.fromhex(fx['pickle_hex']))
    mine = {ix.index_name: ix for ix in [SingleFieldIndexer('by_tag', 'tag'), SingleFieldIndexer('by_id', 'id'),
                                         SingleFieldIndexer('by_vec', 'vec'), FieldNotNullIndexer('has_sensor', 'sensor')]}
    for piece_index, rows in enumerate(_index_fixture_pieces()):
        for ix in mine.values():
            ix.build_index(rows, piece_index)
    for index_dict in (ref, mine, restricted_loads(pickle_indexers_reference_compatible(mine))):
        assert isinstance(index_dict['by_tag'], SingleFieldIndexer)
        assert isinstance(index_dict['has_sensor'], FieldNotNullIndexer)
        lk = fx['lookups']
        for v, exp in lk['by_tag'].items():
            assert sorted(index_dict['by_tag'].get_row_group_indexes(v)) == exp
        for v, exp in lk['by_id'].items():
            assert sorted(index_dict['by_id'].get_row_group_indexes(int(v))) == exp
        for v, exp in lk['by_vec'].items():
            assert sorted(index_dict['by_vec'].get_row_group_indexes(int(v))) == exp
        assert sorted(index_dict['has_sensor'].get_row_group_indexes()) == lk['has_sensor']
        for name, exp in lk['indexed_values'].items():
            assert sorted(map(str, index_dict[name].indexed_values)) == exp
        assert index_dict['by_tag'].column_names == ['tag'] and index_dict['by_tag'].index_name == 'by_tag'
        # selectors: pure set algebra over the index
        assert SingleIndexSelector('by_tag', ['tag_0']).select_row_groups(index_dict) == set(lk['by_tag']['tag_0'])
        both = IntersectIndexSelector([SingleIndexSelector('by_id', [7]), SingleIndexSelector('by_tag', ['tag_3'])])
        assert both.select_row_groups(index_dict) == set(lk['by_id']['7']) & set(lk['by_tag']['tag_3'])
        either = UnionIndexSelector([SingleIndexSelector('by_id', [0]), SingleIndexSelector('by_id', [29])])
        assert either.select_row_groups(index_dict) == set(lk['by_id']['0']) | set(lk['by_id']['29'])
    # the pickle names the reference's module, never this package
    ops = [str(arg) for op, arg, _ in pickletools.genops(pickle_indexers_reference_compatible(mine)) if op.name == 'GLOBAL']
    assert any(o.startswith('petastorm.etl.rowgroup_indexers ') for o in ops)
    assert not any('petastorm_b200' in o for o in ops)
    with pytest.raises(TypeError):
        mine['by_tag'] + mine['has_sensor']


def test_update_common_metadata_keeps_the_other_keys(tmp_path):
    """Role of utils.add_to_dataset_metadata (petastorm/utils.py:88-130): adding the row-group index must not lose the
    unischema / row-groups-per-file entries, and the cached footer of the rewritten file must not be served again."""
    from petastorm_b200.etl import dataset_metadata as dm
    from petastorm_b200.etl.rowgroup_indexers import SingleFieldIndexer, pickle_indexers_reference_compatible
    url = datasets.build('test', str(tmp_path / 't'), 12, row_group_rows=5)
    path = url[len('file://'):]
    ds = dm.ParquetDataset(path)
    before = dict(ds.common_metadata)
    with pytest.raises(dm.PetastormMetadataError):
        dm.get_row_group_indexes(ds)
    ix = SingleFieldIndexer('by_id', 'id')
    ix.build_index([{'id': np.int32(3)}, {'id': np.int32(4)}], 0)
    ix.build_index([{'id': np.int32(4)}], 1)
    dm.update_common_metadata(ds, {dm.ROWGROUPS_INDEX_KEY: pickle_indexers_reference_compatible({'by_id': ix})})
    for fresh in (ds, dm.ParquetDataset(path)):
        after = fresh.common_metadata
        for k, v in before.items():
            if k != b'ARROW:schema':         # pyarrow's own copy of the schema embeds the key-value metadata
                assert after[k] == v
        loaded = dm.get_row_group_indexes(fresh)
        assert sorted(loaded['by_id'].get_row_group_indexes(4)) == [0, 1]
        assert sorted(loaded['by_id'].get_row_group_indexes(3)) == [0]
    assert list(dm.get_schema(dm.ParquetDataset(path)).fields) == list(dm.get_schema(ds).fields)


def test_weighted_sampling_reader_contract():
    """petastorm/tests/test_weighted_sampling_reader.py: mixing by probability, validation of mismatching readers,
    stop when the first reader is exhausted."""
    from collections import namedtuple
    from petastorm_b200.weighted_sampling_reader import WeightedSamplingReader

    class FakeSchema(object):
        def __init__(self, names):
            self.fields = {n: None for n in names}

    class FakeReader(object):
        def __init__(self, value, n, names=('x',), batched=False, ngram=None):
            self._value, self._left = value, n
            self.schema, self.batched_output, self.ngram = FakeSchema(names), batched, ngram
            self.last_row_consumed = False
            self.stopped = self.joined = False

        def __iter__(self):
            return self

        def __next__(self):
            if self._left == 0:
                self.last_row_consumed = True
                raise StopIteration
            self._left -= 1
            return self._value

        def stop(self):
            self.stopped = True

        def join(self):
            self.joined = True

    np.random.seed(4)
    a, b = FakeReader('a', 10 ** 6), FakeReader('b', 10 ** 6)
    with WeightedSamplingReader([a, b], [1, 3]) as mixed:      # not normalised on purpose
        draws = [next(mixed) for _ in range(4000)]
    assert a.stopped and a.joined and b.stopped and b.joined
    assert 0.70 < draws.count('b') / 4000.0 < 0.80
    short = WeightedSamplingReader([FakeReader('a', 3), FakeReader('b', 10 ** 6)], [0.5, 0.5])
    assert len(list(short)) >= 3 and short.last_row_consumed
    with pytest.raises(ValueError, match='Two or more'):
        WeightedSamplingReader([a], [1.0])
    with pytest.raises(ValueError, match='same length'):
        WeightedSamplingReader([a, b], [1.0])
    with pytest.raises(ValueError, match='batched_output'):
        WeightedSamplingReader([a, FakeReader('c', 1, batched=True)], [0.5, 0.5])
    with pytest.raises(ValueError, match='same schema'):
        WeightedSamplingReader([a, FakeReader('c', 1, names=('y',))], [0.5, 0.5])
    with pytest.raises(ValueError, match='ngram'):
        WeightedSamplingReader([a, FakeReader('c', 1, ngram=namedtuple('N', 'a')(1))], [0.5, 0.5])


def test_local_disk_cache_validation_and_eviction(tmp_path):
    """petastorm/local_disk_cache.py:23-82: the shard-capacity sanity check, and least-recently-stored eviction.  The
    values a reader stores are host images of decoded row-groups; the eviction logic is exercised here through the
    file layer with a fake payload."""
    from petastorm_b200 import gpu_workers
    from petastorm_b200.local_disk_cache import LocalDiskCache
    with pytest.raises(ValueError, match='size_limit_bytes / shards'):
        LocalDiskCache(str(tmp_path / 'c0'), 1000, 100, shards=6)
    LocalDiskCache(str(tmp_path / 'c1'), 1000, 100, shards=6, eviction_policy='none')     # no check without eviction
    cache = LocalDiskCache(str(tmp_path / 'c2'), 40000, 100, shards=3)
    real_to, real_from = gpu_workers.to_host_payload, gpu_workers.from_host_payload
    gpu_workers.to_host_payload = lambda v: v
    gpu_workers.from_host_payload = lambda p, device=None: p
    try:
        calls = []
        for k in range(8):
            assert cache.get('key%d' % k, lambda k=k: calls.append(k) or {'blob': b'x' * 9000, 'k': k})['k'] == k
        assert calls == list(range(8)) and cache.volume() <= 40000 + 9100
        assert cache.get('key7', lambda: calls.append('again'))['k'] == 7 and calls[-1] == 7      # hit
        assert cache.get('key0', lambda: calls.append('refill') or {'k': 0})['k'] == 0 and calls[-1] == 'refill'  # evicted
        assert cache.hits == 1 and cache.misses == 9
    finally:
        gpu_workers.to_host_payload, gpu_workers.from_host_payload = real_to, real_from
    cache.cleanup()
    assert not os.path.exists(str(tmp_path / 'c2'))


def test_planner_indexes_blob_pages_on_the_host(tmp_path):
    """Blob columns (1 MiB tensors, jpeg images) are practically incompressible and pyarrow writes a whole column chunk of
    large values as ONE page: the planner walks the few Snappy tags of such literal-dominated pages itself
    (num_host_indexed_pages) instead of leaving a multi-megabyte page to the serial walk of k_snappy_index, and indexes the
    BYTE_ARRAY dictionary of pages the device sees uncompressed."""
    import glob
    from petastorm_b200 import native
    url = datasets.build('tensor_c4', str(tmp_path / 'c4'), 8, row_group_rows=8)
    f = native.ParquetFile(glob.glob(url[7:] + '/*.parquet')[0])
    info = native.Plan(f, 0, list(range(f.num_columns))).info
    assert info.num_snappy_fragments >= 8 * 16          # the tensor page still decompresses on the device, in parallel
    assert info.num_index_pages == 0 and info.num_host_indexed_pages >= 1
    # a compressible multi-fragment page is left to the device index
    import pyarrow as pa
    import pyarrow.parquet as pq
    path = str(tmp_path / 'narrow.parquet')
    pq.write_table(pa.table({'narrow': np.random.default_rng(0).integers(0, 2 ** 20, 300000, dtype=np.int64)}), path,
                   compression='snappy', use_dictionary=False, data_page_size=1 << 20)
    info = native.Plan(native.ParquetFile(path), 0, [0]).info
    assert info.num_index_pages >= 2 and info.num_host_indexed_pages == 0


def test_planner_counts_the_pages_of_the_cluster_index(tmp_path):
    """Pages of >= 256 KiB stored bytes among those the device has to index (the 1 MiB dictionary pages pyarrow writes for
    int64 columns of more than 131,072 distinct values) are the ones the opt-in four-CTA cluster variant of the index
    kernel takes (PST_IDX_CLUSTER=1); the float32 dictionary pages of random values are literal-only Snappy and never
    reach the device index at all."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from petastorm_b200 import native
    n = 300000
    rng = np.random.default_rng(7)
    path = str(tmp_path / 'c2like.parquet')
    pq.write_table(pa.table({'f0': rng.standard_normal(n).astype(np.float32), 'f1': rng.standard_normal(n).astype(np.float32),
                             'i0': rng.integers(0, 2 ** 40, n, dtype=np.int64),
                             'i1': rng.integers(0, 2 ** 40, n, dtype=np.int64)}), path, compression='snappy',
                   row_group_size=n)
    info = native.Plan(native.ParquetFile(path), 0, [0, 1, 2, 3]).info
    assert info.num_cluster_index_pages == 2                     # one dictionary page per int64 column
    assert info.num_index_pages > info.num_cluster_index_pages   # their data pages stay with the single-CTA kernel
    assert info.num_unwrapped_pages >= 2                         # the float columns never see the Snappy kernels


def test_row_packer_views_are_the_columns_again():
    """The device loaders pack the narrow columns of a row-group into one uint8 [n, R] buffer and hand batches out as
    typed strided views of it (one as_strided per field): values, dtypes and shapes must survive, also for a batch that
    is a slice of the buffer (non-zero storage offset) and for the wide columns that stay separate."""
    import torch
    from petastorm_b200.pytorch import _RowPacker
    n = 37
    g = torch.Generator().manual_seed(3)
    cols = {'ts': torch.arange(n * 16, dtype=torch.int64).reshape(n, 16),
            'a': torch.randn(n, 16, generator=g), 'b': torch.randn(n, 16, 2, generator=g),
            'c': torch.randint(0, 100, (n,), dtype=torch.int32, generator=g),
            'img': torch.randn(n, 3, 64, 64, generator=g),
            'u': torch.randint(0, 255, (n, 5), dtype=torch.uint8, generator=g)}
    packer = _RowPacker(cols)
    assert [v[0] for v in packer.small] == ['ts', 'a', 'b', 'c', 'u'] and packer.big == ['img']
    assert packer.row_bytes % 16 == 0
    packed = packer.pack(cols)
    assert len(packed) == 2 and packed[0].dtype == torch.uint8 and packed[0].shape == (n, packer.row_bytes)
    for batch, expect in ((packed, cols), ([t[5:20] for t in packed], {k: v[5:20] for k, v in cols.items()})):
        out = packer.unpack(batch)
        assert list(out) == list(cols)
        for k in cols:
            assert out[k].dtype == cols[k].dtype and out[k].shape == expect[k].shape
            assert torch.equal(out[k], expect[k]), k
    # a single narrow column is not worth a packed buffer
    assert _RowPacker({'x': cols['a'], 'img': cols['img']}).small == []
