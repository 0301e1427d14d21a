"""Pins the oracle (oracle/port.py, a CPU restatement) to the reference: its output must reproduce the golden vectors
that the *unmodified* reference generated (oracle/make_golden.py), on the reference's own checked-in legacy datasets
and on the synthetic scenarios.  CPU only."""
import os

import numpy as np
import pytest

import datasets
from helpers import GOLDEN, golden, load_schema, oracle_specs
from oracle import port

LEGACY_VERSIONS = ['0.4.0', '0.4.3', '0.5.1', '0.6.0', '0.7.0', '0.7.6']


@pytest.mark.parametrize('version', LEGACY_VERSIONS)
def test_port_reproduces_reference_on_legacy_datasets(version):
    d = os.path.join(GOLDEN, 'legacy', version)
    rows = port.read_rows('file://' + d, oracle_specs(load_schema(d)))
    rows.sort(key=lambda r: int(r['id']))
    got = [datasets.digest_row(r) for r in rows]
    exp = golden('legacy_expected.json')[version]
    assert len(got) == len(exp) == 100
    assert got == exp


@pytest.fixture(scope='module')
def synth(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp('synth'))
    urls = {
        'hello': datasets.build('hello', os.path.join(tmp, 'hello'), 12, row_group_rows=5),
        'test': datasets.build('test', os.path.join(tmp, 'test'), 40, row_group_rows=6, partition_by='partition_key'),
        'series': datasets.build('series', os.path.join(tmp, 'series'), 200, row_group_rows=80),
        'tensor': datasets.build('tensor', os.path.join(tmp, 'tensor'), 30, row_group_rows=8),
        'flat': datasets.write_flat(os.path.join(tmp, 'flat'), 600, files=2, row_group_size=100),
    }
    return urls


def _specs(url):
    return oracle_specs(load_schema(url[len('file://'):]))


def test_port_rows_match_reference(synth):
    exp = golden('synthetic_expected.json')
    assert [datasets.digest_row(r) for r in port.read_rows(synth['hello'], _specs(synth['hello']))] == exp['hello_rows']
    s = _specs(synth['test'])
    assert [datasets.digest_row(r) for r in port.read_rows(synth['test'], s)] == exp['test_rows']
    ids = lambda rows: [int(r['id']) for r in rows]  # noqa: E731
    assert ids(port.read_rows(synth['test'], s, shuffle_row_groups=True, seed=42)) == exp['test_ids_shuffle_row_groups_seed42']
    assert ids(port.read_rows(synth['test'], s, shuffle_rows=True, seed=7)) == exp['test_ids_shuffle_rows_seed7']
    assert ids(port.read_rows(synth['test'], s, shuffle_row_drop_partitions=3)) == exp['test_ids_drop_partitions_3']
    assert ids(port.read_rows(synth['test'], s, cur_shard=1, shard_count=3)) == exp['test_ids_shard_1_of_3']
    assert ids(port.read_rows(synth['test'], s, num_epochs=2)) == exp['test_ids_two_epochs']


def test_port_predicates_match_reference(synth):
    from petastorm_b200.predicates import in_pseudorandom_split, in_set
    exp = golden('synthetic_expected.json')
    s = _specs(synth['test'])
    ids = lambda rows: [int(r['id']) for r in rows]  # noqa: E731
    assert ids(port.read_rows(synth['test'], s, predicate=in_set({3, 7, 8, 21, 39, 1000}, 'id'))) == exp['test_ids_in_set']
    assert ids(port.read_rows(synth['test'], s, predicate=in_pseudorandom_split([0.3, 0.4, 0.3], 1, 'id'))) == \
        exp['test_ids_pseudorandom_split']


def test_port_ngram_and_transform_match_reference(synth):
    from petastorm_b200.predicates import in_set
    exp = golden('synthetic_expected.json')
    s = _specs(synth['series'])
    fields = {k: ['ts', 'c00', 'c11'] if k % 2 == 0 else ['ts', 'c05'] for k in range(4)}
    for overlap in (True, False):
        res = port.read_rows(synth['series'], s, ngram=dict(fields=fields, ts='ts', delta=1, overlap=overlap))
        got = [{str(k): datasets.digest_row(v) for k, v in sorted(item.items())} for item in res]
        assert got == exp['series_ngram_overlap_%s' % overlap]

    def norm(row):
        row['tensor'] = ((row['tensor'].astype(np.float32) - np.float32(0.25)) / np.float32(1.5)).astype(np.float16)
        return row

    res = port.read_rows(synth['tensor'], _specs(synth['tensor']), predicate=in_set(set(range(0, 30, 2)), 'key'),
                         transform_func=norm)
    assert [datasets.digest_row(r) for r in res] == exp['tensor_even_normalized']


def test_port_batches_match_reference(synth):
    exp = golden('synthetic_expected.json')
    # the reference's schema inference drops uint16 (petastorm/unischema.py:467-502 has no uint16 branch)
    cols = [c for c in exp['flat_batches'][0].keys()]
    got = port.read_batches(synth['flat'], columns=cols)
    assert [datasets.digest_row(r) for r in got] == exp['flat_batches']
    got = port.read_batches(synth['flat'], columns=cols, shuffle_rows=True, seed=11, shuffle_row_groups=True)
    assert [[int(k) for k in r['key']] for r in got] == exp['flat_keys_shuffled_seed11']
    got = port.read_batches(synth['flat'], columns=cols, shuffle_row_drop_partitions=2)
    assert [[int(k) for k in r['key']] for r in got] == exp['flat_keys_drop_partitions_2']


def test_port_batch_transform_matches_reference(synth):
    """TransformSpec on the batch reader (arrow_reader_worker.py:247-277): func over the DataFrame of a row-group,
    removed / selected fields."""
    exp = golden('synthetic_expected.json')
    cols = datasets.FLAT_TRANSFORM_FIELDS
    got = port.read_batches(synth['flat'], columns=cols, transform_func=datasets.flat_transform,
                            removed_fields=['i01', 'name'])
    assert [datasets.digest_row(r) for r in got] == exp['flat_transform_removed']
    got = port.read_batches(synth['flat'], columns=cols, transform_func=datasets.flat_transform_select)
    assert [datasets.digest_row(r) for r in got] == exp['flat_transform_selected']
    got = port.read_batches(synth['flat'], columns=cols, removed_fields=['i01', 'name'])
    assert [datasets.digest_row(r) for r in got] == exp['flat_transform_only_removed']


def test_port_config_shapes_match_reference(tmp_path):
    """C4 (1 MiB float16 tensors, predicate + normalise) and C5 (120k rows, NGram 16 over 13 fields) at config shape."""
    from petastorm_b200.predicates import in_set
    exp = golden('synthetic_expected.json')
    url = datasets.build('tensor_c4', str(tmp_path / 'c4'), 64, row_group_rows=16)
    specs = _specs(url)

    def norm(row):
        row['tensor'] = ((row['tensor'].astype(np.float32) - np.float32(0.25)) / np.float32(1.5)).astype(np.float16)
        return row

    res = port.read_rows(url, specs, predicate=in_set(set(range(0, 64, 2)), 'key'), transform_func=norm)
    assert [datasets.digest_row(r) for r in res] == exp['tensor_c4_even_normalized']
    url = datasets.build('series', str(tmp_path / 'c5'), 120000, row_group_rows=60000)
    specs = _specs(url)
    names = list(specs.keys())
    res = port.read_rows(url, specs, ngram=dict(fields={k: names for k in range(16)}, ts='ts', delta=1))
    assert datasets.ngram_column_digests(res, range(16), names) == exp['series_big_ngram16']
