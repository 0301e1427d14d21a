"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Row-group index fixture made with the *unmodified* reference indexer classes
(petastorm/etl/rowgroup_indexers.py:21-128) in this container:

    python oracle/make_golden_index.py        ->  tests/golden/rowgroup_index.json

The file holds the pickle the reference stores under ``dataset-toolkit.rowgroups_index.v1`` (hex) for a fixed set of
fake decoded rows, and the look-ups the reference's own classes answer on it.
"""
import json
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402


def fake_pieces():
    import numpy as np
    pieces = []
    for p in range(6):
        rows = []
        for i in range(5):
            k = p * 5 + i
            rows.append({'id': np.int64(k), 'tag': 'tag_%d' % (k % 4), 'sensor': None if p % 2 else 's%d' % (k % 3),
                         'vec': np.array([k % 3, 7 + p], dtype=np.int32)})
        pieces.append(rows)
    return pieces


def main():
    assert refshim.available(), 'needs /root/reference'
    refshim.activate()
    from petastorm.etl.rowgroup_indexers import FieldNotNullIndexer, SingleFieldIndexer
    indexers = [SingleFieldIndexer('by_tag', 'tag'), SingleFieldIndexer('by_id', 'id'),
                SingleFieldIndexer('by_vec', 'vec'), FieldNotNullIndexer('has_sensor', 'sensor')]
    for piece_index, rows in enumerate(fake_pieces()):
        for ix in indexers:
            ix.build_index(rows, piece_index)
    index_dict = {ix.index_name: ix for ix in indexers}
    blob = pickle.dumps(index_dict, protocol=2)
    lookups = {
        'by_tag': {v: sorted(index_dict['by_tag'].get_row_group_indexes(v)) for v in ['tag_0', 'tag_3']},
        'by_id': {str(v): sorted(index_dict['by_id'].get_row_group_indexes(v)) for v in [0, 7, 29]},
        'by_vec': {str(v): sorted(index_dict['by_vec'].get_row_group_indexes(v)) for v in [0, 2, 9, 12]},
        'has_sensor': sorted(index_dict['has_sensor'].get_row_group_indexes()),
        'indexed_values': {name: sorted(map(str, index_dict[name].indexed_values)) for name in index_dict},
    }
    out = os.path.join(ROOT, 'tests', 'golden', 'rowgroup_index.json')
    with open(out, 'w') as f:
        json.dump({'pickle_hex': blob.hex(), 'lookups': lookups}, f, indent=0, sort_keys=True)
    print('wrote', out, len(blob), 'bytes')


if __name__ == '__main__':
    main()
