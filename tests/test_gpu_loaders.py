"""GPU tests of the PyTorch loaders over petastorm_b200 readers (model: petastorm/tests/test_pytorch_dataloader.py):
value parity with the oracle, ordered when shuffling is off, as a multiset keyed by id when a shuffling buffer with the
global RNG is involved (petastorm/reader_impl/shuffling_buffer.py:162)."""
import os

import numpy as np
import pytest

import datasets
from helpers import load_schema, oracle_specs, to_host
from oracle import port

pytestmark = pytest.mark.gpu

NUMERIC_FIELDS = ['id', 'id2', 'id_float', 'id_odd', 'python_primitive_uint8', 'image_png', 'matrix', 'matrix_uint16',
                  'matrix_uint32']


@pytest.fixture(scope='module')
def synth(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp('loaders'))
    return {
        'test': datasets.build('test', os.path.join(tmp, 'test'), 40, row_group_rows=6, partition_by='partition_key'),
        'hello': datasets.build('hello', os.path.join(tmp, 'hello'), 12, row_group_rows=5),
        'flat': datasets.write_flat(os.path.join(tmp, 'flat'), 600, files=2, row_group_size=100),
        'series': datasets.build('series', os.path.join(tmp, 'series'), 200, row_group_rows=80),
    }


def _expected_rows(url, fields):
    specs = oracle_specs(load_schema(url[len('file://'):]))
    specs = {k: v for k, v in specs.items() if k in fields}
    rows = port.read_rows(url, specs)
    return {int(r['id']): port.sanitize_pytorch_types(dict(r)) for r in rows}


def _check_batches(batches, expected, batch_size, ordered_ids=None):
    import torch
    seen = []
    for b in batches:
        n = len(b['id'])
        assert n <= batch_size
        ids = to_host(b['id']).tolist()
        for k, i in enumerate(ids):
            for name, ev in expected[i].items():
                gv = to_host(b[name][k])
                np.testing.assert_array_equal(gv, ev, err_msg=name)
                assert np.asarray(gv).dtype == np.asarray(ev).dtype, name
        seen.extend(ids)
    assert sorted(seen) == sorted(expected.keys())
    if ordered_ids is not None:
        assert seen == ordered_ids
    return seen


@pytest.mark.parametrize('capacity', [0, 3, 11, 1000])
def test_dataloader_row_reader(synth, capacity):
    from petastorm_b200 import make_reader
    from petastorm_b200.pytorch import DataLoader
    url = synth['test']
    exp = _expected_rows(url, NUMERIC_FIELDS)
    with DataLoader(make_reader(url, schema_fields=NUMERIC_FIELDS, shuffle_row_groups=False), batch_size=7,
                    shuffling_queue_capacity=capacity) as loader:
        batches = list(loader)
    order = [int(r['id']) for r in port.read_rows(url, {'id': oracle_specs(load_schema(url[7:]))['id']})]
    seen = _check_batches(batches, exp, 7, ordered_ids=order if capacity == 0 else None)
    assert [len(b['id']) for b in batches[:-1]] == [7] * (len(batches) - 1)
    assert loader.device_batched is True     # whole row-groups through the device shuffling buffer, no per-row Python
    import torch
    assert all(v.is_cuda for v in batches[0].values())
    assert batches[0]['image_png'].is_cuda and batches[0]['matrix_uint16'].dtype == torch.int32
    assert batches[0]['matrix_uint32'].dtype == torch.int64 and batches[0]['id_odd'].dtype == torch.uint8
    if capacity >= 11:
        assert seen != order  # decorrelated (probability of identity is negligible)


def test_dataloader_custom_collate_takes_the_row_loop(synth):
    """A user collate_fn sees exactly what upstream hands it: a list of row dicts (petastorm/pytorch.py:131-248)."""
    from petastorm_b200 import make_reader
    from petastorm_b200.pytorch import DataLoader
    url = synth['test']
    seen = []

    def collate(rows):
        assert isinstance(rows, list) and isinstance(rows[0], dict)
        seen.append(len(rows))
        return [int(r['id']) for r in rows]

    with DataLoader(make_reader(url, schema_fields=['id', 'matrix'], shuffle_row_groups=False), batch_size=9,
                    collate_fn=collate) as loader:
        ids = sum(list(loader), [])
    order = [int(r['id']) for r in port.read_rows(url, {'id': oracle_specs(load_schema(url[7:]))['id']})]
    assert loader.device_batched is False and ids == order and seen == [9, 9, 9, 9, 4]


def test_dataloader_batch_reader_device_path(synth):
    """DataLoader over make_batch_reader (upstream transposes every row-group into per-row tuples, pytorch.py:207-216):
    same batches, formed on the device."""
    import torch
    from petastorm_b200 import make_batch_reader
    from petastorm_b200.pytorch import DataLoader
    url = synth['flat']
    cols = ['key', 'f00', 'i00', 'small', 'flag']
    exp = port.read_batches(url, columns=cols)
    exp = {k: np.concatenate([e[k] for e in exp]) for k in cols}
    with DataLoader(make_batch_reader(url, schema_fields=cols, shuffle_row_groups=False), batch_size=128) as loader:
        batches = list(loader)
    assert loader.device_batched is True
    assert [len(b['key']) for b in batches] == [128] * 4 + [88]
    for k in cols:
        got = torch.cat([b[k] for b in batches]).cpu().numpy()
        e = exp[k].astype(np.uint8) if exp[k].dtype == np.bool_ else exp[k]
        np.testing.assert_array_equal(got, e)
        assert got.dtype == e.dtype


def test_dataloader_rejects_strings_and_nulls(synth):
    from petastorm_b200 import make_reader
    from petastorm_b200.pytorch import DataLoader
    with DataLoader(make_reader(synth['test'], schema_fields=['id', 'sensor_name'], shuffle_row_groups=False)) as loader:
        with pytest.raises(TypeError, match='string'):
            next(iter(loader))
    with DataLoader(make_reader(synth['test'], schema_fields=['id', 'integer_nullable'], shuffle_row_groups=False)) as loader:
        with pytest.raises(TypeError, match='nullable'):
            list(loader)


@pytest.mark.parametrize('capacity', [0, 20])
def test_batched_dataloader_row_reader(synth, capacity):
    import torch
    from petastorm_b200 import make_reader
    from petastorm_b200.pytorch import BatchedDataLoader
    url = synth['hello']
    specs = oracle_specs(load_schema(url[7:]))
    exp = {int(r['id']): r for r in port.read_rows(url, specs)}
    with BatchedDataLoader(make_reader(url, shuffle_row_groups=False, num_epochs=2), batch_size=5,
                           shuffling_queue_capacity=capacity) as loader:
        batches = list(loader)
    seen = []
    for b in batches:
        assert isinstance(b['image1'], torch.Tensor) and b['image1'].is_cuda and b['id'].is_cuda
        ids = b['id'].cpu().tolist()
        for k, i in enumerate(ids):
            np.testing.assert_array_equal(b['image1'][k].cpu().numpy(), exp[i]['image1'])
            np.testing.assert_array_equal(b['array_4d'][k].cpu().numpy(), exp[i]['array_4d'])
        seen.extend(ids)
    assert sorted(seen) == sorted(list(exp.keys()) * 2)
    if capacity == 0:
        assert seen == list(range(12)) * 2 and all(len(b['id']) == 5 for b in batches[:-1])


def test_batched_dataloader_batch_reader_c2_shape(synth):
    """C2 consumer: make_batch_reader row-groups re-batched to a fixed batch size on the device."""
    import torch
    from petastorm_b200 import make_batch_reader
    from petastorm_b200.pytorch import BatchedDataLoader
    url = synth['flat']
    cols = ['key', 'f00', 'f07', 'i00', 'i01', 'small', 'flag']
    exp = port.read_batches(url, columns=cols)
    exp = {k: np.concatenate([e[k] for e in exp]) for k in cols}
    with BatchedDataLoader(make_batch_reader(url, schema_fields=cols, shuffle_row_groups=False), batch_size=64) as loader:
        batches = list(loader)
    assert [len(b['key']) for b in batches] == [64] * 9 + [24]
    for k in cols:
        got = torch.cat([b[k] for b in batches]).cpu().numpy()
        e = exp[k].astype(np.uint8) if exp[k].dtype == np.bool_ else exp[k]
        np.testing.assert_array_equal(got, e)
    # with a shuffling queue: same multiset of rows
    with BatchedDataLoader(make_batch_reader(url, schema_fields=cols, shuffle_row_groups=False), batch_size=64,
                           shuffling_queue_capacity=150) as loader:
        got = torch.cat([b['key'] for b in loader]).cpu().numpy()
    assert sorted(got.tolist()) == sorted(exp['key'].tolist()) and got.tolist() != exp['key'].tolist()


def test_in_mem_loader_and_reiteration_rules(synth):
    import torch
    from petastorm_b200 import make_reader
    from petastorm_b200.pytorch import BatchedDataLoader, DataLoader, InMemBatchedDataLoader
    url = synth['hello']
    loader = InMemBatchedDataLoader(make_reader(url, shuffle_row_groups=False, num_epochs=None), batch_size=4,
                                    num_epochs=3, seed=7, rows_capacity=10, shuffle=True)
    ids = [b['id'].cpu().tolist() for b in loader]
    assert len(ids) == 9
    for epoch in range(3):
        g = torch.Generator()
        g.manual_seed(7 + epoch)
        perm = torch.randperm(10, generator=g).tolist()
        assert sum(ids[epoch * 3:(epoch + 1) * 3], []) == perm
    with pytest.raises(RuntimeError):
        list(loader)
    # a second full pass resets the reader; a nested pass raises (petastorm/pytorch.py:109-128)
    dl = DataLoader(make_reader(url, schema_fields=['id'], shuffle_row_groups=False), batch_size=5)
    first = [b['id'].tolist() for b in dl]
    second = [b['id'].tolist() for b in dl]
    assert first == second and sum(first, []) == list(range(12))
    it = iter(dl)
    next(it)
    with pytest.raises(RuntimeError, match='must finish a full pass'):
        next(iter(dl))
    dl.reader.stop()
    dl.reader.join()


def test_ngram_through_dataloader(synth):
    from petastorm_b200 import make_reader
    from petastorm_b200.ngram import NGram
    from petastorm_b200.pytorch import DataLoader
    url = synth['series']
    schema = load_schema(url[7:])
    ng = NGram({0: [schema.ts, schema.c00], 1: [schema.ts, schema.c01]}, delta_threshold=1, timestamp_field=schema.ts)
    specs = oracle_specs(schema)
    exp = port.read_rows(url, specs, ngram=dict(fields={0: ['ts', 'c00'], 1: ['ts', 'c01']}, ts='ts', delta=1))
    with DataLoader(make_reader(url, schema_fields=ng, shuffle_row_groups=False), batch_size=16,
                    shuffling_queue_capacity=50) as loader:
        got = []
        for b in loader:
            for k in range(len(b[0]['ts'])):
                got.append((int(b[0]['ts'][k]), float(b[0]['c00'][k]), int(b[1]['ts'][k]), float(b[1]['c01'][k])))
    want = [(int(w[0]['ts']), float(w[0]['c00']), int(w[1]['ts']), float(w[1]['c01'])) for w in exp]
    assert sorted(got) == sorted(want)
