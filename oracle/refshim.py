"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Runs the *unmodified* reference (uber/petastorm 0.13.1 under /root/reference)
in this container: its own ``ArrowReaderWorker`` / ``PyDictReaderWorker`` / ``ConcurrentVentilator`` / ``DummyPool`` /
codecs / NGram, given duck-typed ``piece`` / ``dataset`` objects in place of the legacy pyarrow classes that pyarrow 24
removed.  Only the ``Reader.__init__`` facade is restated (``oracle/port.py``).

Usable only where /root/reference exists (this container): it pins ``oracle/port.py`` and generates ``tests/golden``.
Run with ``PYTHONPATH=oracle/shims:/root/reference`` (``activate()`` arranges that in-process).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'petastorm'))


def activate():
    """Put the shims and the reference on sys.path and apply the numpy/pyarrow aliases (idempotent)."""
    shims = os.path.join(HERE, 'shims')
    for p in (REFERENCE, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib.util
    if 'sitecustomize_pst' not in sys.modules:
        spec = importlib.util.spec_from_file_location('sitecustomize_pst', os.path.join(shims, 'sitecustomize.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules['sitecustomize_pst'] = mod
    import petastorm  # noqa: F401  pylint: disable=unused-import
    return petastorm


class _Partitions(object):
    def __init__(self, names):
        self.partition_names = set(names)
        self._names = list(names)

    def __bool__(self):
        return bool(self._names)

    def __len__(self):
        return len(self._names)


class Piece(object):
    """Stands in for ``pq.ParquetDatasetPiece`` (path, row_group, partition_keys, read)."""

    def __init__(self, path, row_group, partition_keys=()):
        self.path = path
        self.row_group = row_group
        self.partition_keys = list(partition_keys)

    def read(self, columns=None, partitions=None, **kwargs):
        import pyarrow as pa
        import pyarrow.parquet as pq
        cols = None if columns is None else [c for c in columns]
        table = pq.ParquetFile(self.path).read_row_group(self.row_group, columns=cols)
        # legacy pyarrow appended the partition values of the piece as (dictionary) columns
        for name, value in self.partition_keys:
            try:
                arr = pa.array([int(value)] * table.num_rows, type=pa.int64())
            except ValueError:
                arr = pa.array([value] * table.num_rows, type=pa.string())
            table = table.append_column(name, arr)
        return table


class _FS(object):
    def open(self, path, mode='rb'):
        return open(path, mode)


class Dataset(object):
    def __init__(self, partition_names):
        self.partitions = _Partitions(partition_names) if partition_names else None
        self.fs = _FS()


def _worker_classes():
    activate()
    from petastorm.arrow_reader_worker import ArrowReaderWorker
    from petastorm.py_dict_reader_worker import PyDictReaderWorker

    class ShimArrowWorker(ArrowReaderWorker):
        _pnames = ()

        def process(self, *a, **k):
            if not self._dataset:
                self._dataset = Dataset(self._pnames)
            return super(ShimArrowWorker, self).process(*a, **k)

    class ShimPyDictWorker(PyDictReaderWorker):
        _pnames = ()

        def process(self, *a, **k):
            if not self._dataset:
                self._dataset = Dataset(self._pnames)
            return super(ShimPyDictWorker, self).process(*a, **k)

    return ShimArrowWorker, ShimPyDictWorker


def load_reference_unischema(dataset_dir):
    """The Unischema the reference itself un-pickles from ``_common_metadata``."""
    activate()
    import pyarrow.parquet as pq
    from petastorm.etl.legacy import depickle_legacy_package_name_compatible
    md = pq.read_metadata(os.path.join(dataset_dir, '_common_metadata')).metadata
    return depickle_legacy_package_name_compatible(md[b'dataset-toolkit.unischema.v1'])


def _run(worker_cls, url, schema, transformed_schema, ngram, shuffle_row_groups, seed, shuffle_rows, predicate,
         transform_spec, num_epochs, cur_shard, shard_count, drop_partitions):
    from petastorm.cache import NullCache
    from petastorm.workers_pool.dummy_pool import DummyPool
    from petastorm.workers_pool.ventilator import ConcurrentVentilator
    from petastorm.workers_pool import EmptyResultError
    from oracle import port
    pieces_raw = port.list_pieces(url)
    pnames = port.partition_names(pieces_raw)
    pieces = [Piece(p, rg, keys) for p, rg, keys in pieces_raw]
    worker_cls._pnames = tuple(pnames)
    indexes = port.shard_indexes(range(len(pieces)), len(pieces), cur_shard, shard_count, seed)
    items = [{'piece_index': i, 'worker_predicate': predicate, 'shuffle_row_drop_partition': (p, drop_partitions)}
             for i in indexes for p in range(drop_partitions)]
    pool = DummyPool()
    vent = ConcurrentVentilator(pool.ventilate, items, iterations=num_epochs, randomize_item_order=shuffle_row_groups,
                                random_seed=seed, max_ventilation_queue_size=4)
    pool.start(worker_cls, (None, url, schema, ngram, pieces, NullCache(), transform_spec, transformed_schema, None,
                            shuffle_rows, seed, False), ventilator=vent)
    reader = worker_cls.new_results_queue_reader()
    out = []
    try:
        while True:
            out.append(reader.read_next(pool, transformed_schema, ngram))
    except StopIteration:
        pass
    pool.stop()
    pool.join()
    return out


def reference_batches(url, schema, shuffle_row_groups=False, seed=None, shuffle_rows=False, predicate=None,
                      transform_spec=None, num_epochs=1, cur_shard=None, shard_count=None, drop_partitions=1):
    """namedtuples exactly as ``next(make_batch_reader(...))`` of the reference would produce (dummy pool)."""
    arrow_cls, _ = _worker_classes()
    from petastorm.transform import transform_schema
    tschema = transform_schema(schema, transform_spec) if transform_spec else schema
    return _run(arrow_cls, url, schema, tschema, None, shuffle_row_groups, seed, shuffle_rows, predicate,
                transform_spec, num_epochs, cur_shard, shard_count, drop_partitions)


def reference_rows(url, schema, ngram=None, shuffle_row_groups=False, seed=None, shuffle_rows=False, predicate=None,
                   transform_spec=None, num_epochs=1, cur_shard=None, shard_count=None, drop_partitions=1):
    """namedtuples (or NGram dicts) exactly as ``next(make_reader(...))`` of the reference would produce."""
    _, row_cls = _worker_classes()
    from petastorm.transform import transform_schema
    tschema = transform_schema(schema, transform_spec) if transform_spec else schema
    return _run(row_cls, url, schema, tschema, ngram, shuffle_row_groups, seed, shuffle_rows, predicate,
                transform_spec, num_epochs, cur_shard, shard_count, drop_partitions)
