"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates the golden vectors under tests/golden/ by running the *unmodified*
reference (/root/reference, through oracle/refshim.py) in this container.  Commit its outputs; re-run only when the
fixtures in tests/datasets.py change:

    python oracle/make_golden.py

Outputs
  tests/golden/legacy/<version>/...        the reference's six checked-in parquet-mr datasets (binary test DATA copied
                                           from petastorm/tests/data/legacy; they are the reference's on-disk fixtures)
  tests/golden/legacy_expected.json        digest of every decoded field of every row, as decoded by the reference's
                                           own PyDictReaderWorker
  tests/golden/synthetic_expected.json     same for the synthetic scenarios of tests/datasets.py (orders under seeds,
                                           predicates, NGram windows, TransformSpec, batch reader columns)
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import refshim  # noqa: E402

LEGACY_VERSIONS = ['0.4.0', '0.4.3', '0.5.1', '0.6.0', '0.7.0', '0.7.6']


def main():
    assert refshim.available(), 'needs /root/reference'
    refshim.activate()
    import numpy as np
    import datasets
    from petastorm.ngram import NGram
    from petastorm.predicates import in_lambda, in_pseudorandom_split, in_set
    from petastorm.transform import TransformSpec

    golden = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(golden, exist_ok=True)

    # ---- legacy datasets -----------------------------------------------------------------------------------------
    legacy_out = {}
    for v in LEGACY_VERSIONS:
        src = os.path.join(refshim.REFERENCE, 'petastorm', 'tests', 'data', 'legacy', v)
        dst = os.path.join(golden, 'legacy', v)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns('*.crc', '.*'))
        schema = refshim.load_reference_unischema(src)
        rows = refshim.reference_rows('file://' + src, schema)
        rows.sort(key=lambda r: int(r.id))
        legacy_out[v] = [datasets.digest_row(r) for r in rows]
        print('legacy', v, len(rows), 'rows')
    with open(os.path.join(golden, 'legacy_expected.json'), 'w') as f:
        json.dump(legacy_out, f, indent=0, sort_keys=True)

    # ---- synthetic scenarios ---------------------------------------------------------------------------------------
    out = {}
    tmp = tempfile.mkdtemp(prefix='pst_golden_')
    try:
        url = datasets.build('hello', os.path.join(tmp, 'hello'), 12, row_group_rows=5)
        schema = refshim.load_reference_unischema(os.path.join(tmp, 'hello'))
        out['hello_rows'] = [datasets.digest_row(r) for r in refshim.reference_rows(url, schema)]

        url = datasets.build('test', os.path.join(tmp, 'test'), 40, row_group_rows=6, partition_by='partition_key')
        schema = refshim.load_reference_unischema(os.path.join(tmp, 'test'))
        out['test_rows'] = [datasets.digest_row(r) for r in refshim.reference_rows(url, schema)]
        out['test_ids_shuffle_row_groups_seed42'] = [int(r.id) for r in refshim.reference_rows(
            url, schema, shuffle_row_groups=True, seed=42)]
        out['test_ids_shuffle_rows_seed7'] = [int(r.id) for r in refshim.reference_rows(
            url, schema, shuffle_rows=True, seed=7)]
        out['test_ids_drop_partitions_3'] = [int(r.id) for r in refshim.reference_rows(url, schema, drop_partitions=3)]
        out['test_ids_in_set'] = [int(r.id) for r in refshim.reference_rows(
            url, schema, predicate=in_set({3, 7, 8, 21, 39, 1000}, 'id'))]
        out['test_ids_pseudorandom_split'] = [int(r.id) for r in refshim.reference_rows(
            url, schema, predicate=in_pseudorandom_split([0.3, 0.4, 0.3], 1, 'id'))]
        out['test_ids_partition_predicate'] = sorted(int(r.id) for r in refshim.reference_rows(
            url, schema, predicate=in_set({'p_2'}, 'partition_key')))
        out['test_ids_shard_1_of_3'] = [int(r.id) for r in refshim.reference_rows(url, schema, cur_shard=1, shard_count=3)]
        out['test_ids_two_epochs'] = [int(r.id) for r in refshim.reference_rows(url, schema, num_epochs=2)]

        url = datasets.build('series', os.path.join(tmp, 'series'), 200, row_group_rows=80)
        schema = refshim.load_reference_unischema(os.path.join(tmp, 'series'))
        fields = {k: [schema.ts, schema.c00, schema.c11] if k % 2 == 0 else [schema.ts, schema.c05] for k in range(4)}
        for overlap in (True, False):
            ng = NGram(fields, delta_threshold=1, timestamp_field=schema.ts, timestamp_overlap=overlap)
            res = refshim.reference_rows(url, schema, ngram=ng)
            out['series_ngram_overlap_%s' % overlap] = [
                {str(k): datasets.digest_row(v) for k, v in sorted(item.items())} for item in res]

        url = datasets.build('tensor', os.path.join(tmp, 'tensor'), 30, row_group_rows=8)
        schema = refshim.load_reference_unischema(os.path.join(tmp, 'tensor'))

        def norm(row):
            row['tensor'] = ((row['tensor'].astype(np.float32) - np.float32(0.25)) / np.float32(1.5)).astype(np.float16)
            return row

        res = refshim.reference_rows(url, schema, predicate=in_set(set(range(0, 30, 2)), 'key'),
                                     transform_spec=TransformSpec(norm))
        out['tensor_even_normalized'] = [datasets.digest_row(r) for r in res]

        url = datasets.write_flat(os.path.join(tmp, 'flat'), 600, files=2, row_group_size=100)
        from petastorm.unischema import Unischema, UnischemaField
        import pyarrow.parquet as pq
        from oracle import port
        pieces = port.list_pieces(url)
        arrow_schema = pq.ParquetFile(pieces[0][0]).schema_arrow
        from petastorm.unischema import _numpy_and_codec_from_arrow_type
        import pyarrow as pa
        fs = []
        for name in arrow_schema.names:
            t = arrow_schema.field(name).type
            try:
                _numpy_and_codec_from_arrow_type(t)
            except ValueError:
                continue  # Unischema.from_arrow_schema(omit_unsupported_fields=True) drops it (unischema.py:343-349)
            fs.append(UnischemaField(name, _numpy_and_codec_from_arrow_type(t), (None,) if pa.types.is_list(t) else (),
                                     None, arrow_schema.field(name).nullable))
        flat_schema = Unischema('inferred_schema', fs)
        res = refshim.reference_batches(url, flat_schema)
        out['flat_batches'] = [datasets.digest_row(r) for r in res]
        res = refshim.reference_batches(url, flat_schema, shuffle_rows=True, seed=11, shuffle_row_groups=True)
        out['flat_keys_shuffled_seed11'] = [[int(k) for k in r.key] for r in res]
        res = refshim.reference_batches(url, flat_schema, drop_partitions=2)
        out['flat_keys_drop_partitions_2'] = [[int(k) for k in r.key] for r in res]
        # predicate on the batch reader: bool columns make the reference itself crash on pandas 3
        # (`other_data_frame[erase_mask] = None`, arrow_reader_worker.py:331), so the scenario selects other columns
        view = flat_schema.create_schema_view([flat_schema.fields[n] for n in ('key', 'f00', 'i00', 'nullable_int', 'name')])
        res = refshim.reference_batches(url, view, predicate=in_lambda(['key'], lambda key: key % 3 == 0))
        out['flat_predicate_mod3'] = [datasets.digest_row(r) for r in res]
        # TransformSpec on the batch reader (arrow_reader_worker.py:247-277): func over the pandas DataFrame of a
        # row-group, edited / added fields (one of them 2-D), removed_fields or selected_fields
        tview = flat_schema.create_schema_view([flat_schema.fields[n] for n in datasets.FLAT_TRANSFORM_FIELDS])
        spec = TransformSpec(datasets.flat_transform, edit_fields=datasets.FLAT_TRANSFORM_EDITS,
                             removed_fields=['i01', 'name'])
        out['flat_transform_removed'] = [datasets.digest_row(r) for r in refshim.reference_batches(url, tview, transform_spec=spec)]
        spec = TransformSpec(datasets.flat_transform_select, edit_fields=datasets.FLAT_TRANSFORM_EDITS,
                             selected_fields=['sum01', 'mat', 'key'])
        out['flat_transform_selected'] = [datasets.digest_row(r) for r in refshim.reference_batches(url, tview, transform_spec=spec)]
        spec = TransformSpec(datasets.flat_transform_drop, edit_fields=datasets.FLAT_TRANSFORM_EDITS[:2],
                             removed_fields=['i01', 'name'])
        res = refshim.reference_batches(url, tview, transform_spec=spec,
                                        predicate=in_lambda(['key'], lambda key: key % 3 == 0))
        out['flat_transform_predicate'] = [datasets.digest_row(r) for r in res]
        spec = TransformSpec(removed_fields=['i01', 'name'])          # no func: only drops columns
        out['flat_transform_only_removed'] = [datasets.digest_row(r) for r in refshim.reference_batches(url, tview, transform_spec=spec)]

        # ---- config-shape scenarios (C3/C4/C5 were tested at toy sizes only) ---------------------------
        # C4: float16 (32,128,128) = 1 MiB values, 64 rows in 4 row-groups, in_set on the key + normalise
        url = datasets.build('tensor_c4', os.path.join(tmp, 'tensor_c4'), 64, row_group_rows=16)
        schema = refshim.load_reference_unischema(os.path.join(tmp, 'tensor_c4'))
        res = refshim.reference_rows(url, schema, predicate=in_set(set(range(0, 64, 2)), 'key'),
                                     transform_spec=TransformSpec(norm))
        out['tensor_c4_even_normalized'] = [datasets.digest_row(r) for r in res]
        out['tensor_c4_rows'] = [datasets.digest_row(r) for r in refshim.reference_rows(url, schema)]
        # C3: 256 jpeg images in 4 row-groups, row-group order under a seed (labels; pixels are compared within the
        # nvJPEG tolerance against the oracle's cv2 decode, not through a golden)
        url = datasets.build('imagenet', os.path.join(tmp, 'imagenet'), 256, row_group_rows=64)
        schema = refshim.load_reference_unischema(os.path.join(tmp, 'imagenet'))
        out['imagenet_labels_shuffle_row_groups_seed5'] = [int(r.label) for r in refshim.reference_rows(
            url, schema, shuffle_row_groups=True, seed=5)]
        # C5: 120k rows in 2 row-groups with a gap every 37 rows, NGram length 16 over all 13 fields
        url = datasets.build('series', os.path.join(tmp, 'series_big'), 120000, row_group_rows=60000)
        schema = refshim.load_reference_unischema(os.path.join(tmp, 'series_big'))
        names = list(schema.fields.keys())
        ng = NGram({k: list(schema.fields.values()) for k in range(16)}, delta_threshold=1, timestamp_field=schema.ts)
        res = refshim.reference_rows(url, schema, ngram=ng)
        out['series_big_ngram16'] = datasets.ngram_column_digests(res, range(16), names)
        out['series_big_ngram16_first'] = [{str(k): datasets.digest_row(v) for k, v in sorted(item.items())}
                                           for item in res[:5]]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    with open(os.path.join(golden, 'synthetic_expected.json'), 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print('synthetic scenarios:', sorted(out.keys()))


if __name__ == '__main__':
    main()
