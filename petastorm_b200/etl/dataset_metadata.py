"""Dataset discovery and Petastorm metadata, read side.

Covers what the reference gets from the (removed) legacy ``pq.ParquetDataset`` plus
``petastorm/etl/dataset_metadata.py:244-418``: enumerate parquet files (hive ``key=value`` directories become partition
columns), read ``_common_metadata`` / ``_metadata`` key-values, split files into one piece per row-group (three
strategies, pieces sorted by path for a stable order) and load or infer the Unischema.  Footers are parsed by
``libpst_b200.so`` (thrift-compact, host only); no pyarrow on this path.
"""
import json
import logging
import os
from concurrent import futures

import numpy as np

from petastorm_b200 import rowgroup
from petastorm_b200.errors import PetastormMetadataError, PetastormMetadataGenerationError  # noqa: F401
from petastorm_b200.etl.legacy import depickle_legacy_package_name_compatible
from petastorm_b200.unischema import Unischema

logger = logging.getLogger(__name__)

ROW_GROUPS_PER_FILE_KEY = b'dataset-toolkit.num_row_groups_per_file.v1'
UNISCHEMA_KEY = b'dataset-toolkit.unischema.v1'
ROWGROUPS_INDEX_KEY = b'dataset-toolkit.rowgroups_index.v1'


class RowGroupPiece(object):
    """One row-group of one file (stands where ``pq.ParquetDatasetPiece`` stood)."""

    __slots__ = ('path', 'row_group', 'partition_keys')

    def __init__(self, path, row_group, partition_keys=()):
        self.path = path
        self.row_group = row_group
        self.partition_keys = list(partition_keys)  # [(name, string value)]

    def __repr__(self):
        return 'RowGroupPiece({!r}, {}, {})'.format(self.path, self.row_group, self.partition_keys)


class PartitionSet(object):
    """Hive partition levels of a dataset: ordered names, sorted distinct values, inferred numpy type."""

    def __init__(self):
        self.levels = []  # [(name, sorted values as strings, numpy dtype)]

    @property
    def partition_names(self):
        return set(name for name, _, _ in self.levels)

    def __bool__(self):
        return bool(self.levels)

    __nonzero__ = __bool__

    def __len__(self):
        return len(self.levels)

    def dtype_of(self, name):
        for n, _, dt in self.levels:
            if n == name:
                return dt
        raise KeyError(name)


def _is_data_file(name):
    return not (name.startswith('_') or name.startswith('.') or name.endswith('.crc'))


class ParquetDataset(object):
    """A directory (or explicit list) of parquet files on the local filesystem."""

    def __init__(self, path_or_paths, filters=None):
        self.paths = path_or_paths
        self.pieces = []  # one per file: RowGroupPiece(path, None, partition_keys)
        self.partitions = PartitionSet()
        self.common_metadata_path = None
        self.metadata_path = None
        self.base_path = None
        if isinstance(path_or_paths, (list, tuple)):
            for p in path_or_paths:
                if os.path.isdir(p):
                    raise ValueError('A list of urls must point to parquet files, not directories: {}'.format(p))
                if not os.path.exists(p):
                    raise IOError('Path does not exist: {}'.format(p))
                self.pieces.append(RowGroupPiece(p, None, []))
            self.base_path = os.path.dirname(path_or_paths[0]) if path_or_paths else None
        else:
            if not os.path.exists(path_or_paths):
                raise IOError('Passed non-file path: {}'.format(path_or_paths))
            if os.path.isdir(path_or_paths):
                self.base_path = path_or_paths
                self._walk(path_or_paths, [])
                cm = os.path.join(path_or_paths, '_common_metadata')
                md = os.path.join(path_or_paths, '_metadata')
                self.common_metadata_path = cm if os.path.exists(cm) else None
                self.metadata_path = md if os.path.exists(md) else None
            else:
                self.base_path = os.path.dirname(path_or_paths)
                self.pieces.append(RowGroupPiece(path_or_paths, None, []))
        self.pieces.sort(key=lambda p: p.path)
        self._finish_partitions()
        if filters:
            self.pieces = filter_pieces(self.pieces, filters)
        self._common_kv = None
        self._metadata_file = None

    def _walk(self, directory, keys):
        entries = sorted(os.listdir(directory))
        for e in entries:
            full = os.path.join(directory, e)
            if os.path.isdir(full):
                if '=' in e and not e.startswith(('_', '.')):
                    k, v = e.split('=', 1)
                    self._walk(full, keys + [(k, v)])
                elif not e.startswith(('_', '.')):
                    self._walk(full, keys)
            elif _is_data_file(e):
                self.pieces.append(RowGroupPiece(full, None, keys))

    def _finish_partitions(self):
        names = []
        values = {}
        for p in self.pieces:
            for k, v in p.partition_keys:
                if k not in values:
                    names.append(k)
                    values[k] = set()
                values[k].add(v)
        for name in names:
            vals = values[name]
            try:
                ints = sorted(int(v) for v in vals)
                self.partitions.levels.append((name, [str(i) for i in ints], np.int64))
            except ValueError:
                self.partitions.levels.append((name, sorted(vals), np.str_))

    @property
    def common_metadata(self):
        """dict of key/value metadata of ``_common_metadata`` (bytes -> bytes) or None."""
        if self.common_metadata_path is None:
            return None
        if self._common_kv is None:
            self._common_kv = rowgroup.open_file(self.common_metadata_path).key_value_metadata()
        return self._common_kv

    @property
    def metadata(self):
        """native.ParquetFile of the ``_metadata`` summary file or None."""
        if self.metadata_path is None:
            return None
        if self._metadata_file is None:
            self._metadata_file = rowgroup.open_file(self.metadata_path)
        return self._metadata_file

    def first_file(self):
        if not self.pieces:
            raise IOError('No parquet files found in {}'.format(self.paths))
        return rowgroup.open_file(self.pieces[0].path)


_FILTER_OPS = {
    '=': lambda p, f: p == f, '==': lambda p, f: p == f, '!=': lambda p, f: p != f,
    '<': lambda p, f: p < f, '>': lambda p, f: p > f, '<=': lambda p, f: p <= f, '>=': lambda p, f: p >= f,
    'in': lambda p, f: p in f, 'not in': lambda p, f: p not in f,
}


def filter_pieces(pieces, filters):
    """``filters`` of ``make_reader`` / ``make_batch_reader`` with the semantics the reference gets from the legacy
    ``pq.ParquetDataset(filters=...)`` it builds (petastorm/reader.py:430-433): a list of ``(column, op, value)``
    tuples (AND) or a list of such lists (OR of ANDs); only hive partition keys are tested - a predicate on any other
    column accepts every file -, the partition value (a directory-name string) is cast to the type of the filter
    value, and files whose partition keys fail every conjunction are dropped."""
    if not isinstance(filters, (list, tuple)) or not filters:
        raise ValueError('filters must be a non-empty List[Tuple] or List[List[Tuple]]')
    dnf = [list(filters)] if isinstance(filters[0], tuple) or (isinstance(filters[0], list) and filters[0] and
                                                                not isinstance(filters[0][0], (list, tuple))) \
        else [list(c) for c in filters]
    for conj in dnf:
        for f in conj:
            if len(f) != 3 or f[1] not in _FILTER_OPS:
                raise ValueError('"{}" is not a valid filter: expected (column, op, value) with op in {}'.format(
                    f, sorted(_FILTER_OPS)))
            if f[1] in ('in', 'not in') and not isinstance(f[2], (set, list, tuple, frozenset)):
                raise TypeError("'{}' object is not a collection".format(type(f[2]).__name__))

    def accepts(piece, f):
        column, op, value = f
        for k, v in piece.partition_keys:
            if k != column:
                continue
            if op in ('in', 'not in'):
                values = list(value)
                if not values:
                    return op == 'not in'
                cast = type(values[0])
                if not _FILTER_OPS[op](cast(v), set(values)):
                    return False
            elif not _FILTER_OPS[op](type(value)(v), value):
                return False
        return True

    return [p for p in pieces if any(all(accepts(p, f) for f in conj) for conj in dnf)]


def _footer_split(piece):
    f = rowgroup.open_file(piece.path)
    return [RowGroupPiece(piece.path, rg, piece.partition_keys) for rg in range(f.num_row_groups)]


def load_row_groups(dataset):
    """One :class:`RowGroupPiece` per row-group, in path order (petastorm/etl/dataset_metadata.py:244-353).

    Strategy: (1) no ``_common_metadata`` -> read every footer; (2) the ``num_row_groups_per_file`` JSON written by
    ``materialize_dataset``; a ``_common_metadata`` without that key is a PetastormMetadataError like upstream."""
    common = dataset.common_metadata
    if common is None:
        with futures.ThreadPoolExecutor(max_workers=8) as pool:
            parts = list(pool.map(_footer_split, dataset.pieces))
        return [p for sub in parts for p in sub]
    if ROW_GROUPS_PER_FILE_KEY not in common:
        if dataset.metadata is not None and dataset.metadata.num_row_groups > 0:
            # a parquet summary file with row-group information: counts per file come from the footers themselves
            with futures.ThreadPoolExecutor(max_workers=8) as pool:
                parts = list(pool.map(_footer_split, dataset.pieces))
            return [p for sub in parts for p in sub]
        raise PetastormMetadataError(
            'Could not find row group metadata in _common_metadata file.'
            ' Use materialize_dataset(..) in petastorm.etl.dataset_metadata.py to generate'
            ' this file in your ETL code.'
            ' You can generate it on an existing dataset using petastorm-generate-metadata.py')
    per_file = json.loads(common[ROW_GROUPS_PER_FILE_KEY].decode())
    base = dataset.base_path
    out = []
    for piece in sorted(dataset.pieces, key=lambda p: p.path):
        key = os.path.relpath(piece.path, base)
        if key == '.':
            continue
        if key not in per_file:
            raise PetastormMetadataError('File {} is not listed in the dataset metadata ({})'.format(
                key, ROW_GROUPS_PER_FILE_KEY.decode()))
        out.extend(RowGroupPiece(piece.path, rg, piece.partition_keys) for rg in range(per_file[key]))
    return out


def get_schema(dataset):
    """Unischema stored by ``materialize_dataset`` (petastorm/etl/dataset_metadata.py:356-385)."""
    common = dataset.common_metadata
    if common is None:
        raise PetastormMetadataError(
            'Could not find _common_metadata file. Use materialize_dataset(..) in'
            ' petastorm.etl.dataset_metadata.py to generate this file in your ETL code.'
            ' You can generate it on an existing dataset using petastorm-generate-metadata.py')
    if UNISCHEMA_KEY not in common:
        raise PetastormMetadataError(
            'Could not find the unischema in the dataset common metadata file.'
            ' Please provide or generate dataset with the unischema attached.'
            ' Common Metadata file might not be generated properly.'
            ' Make sure to use materialize_dataset(..) in petastorm.etl.dataset_metadata to'
            ' properly generate this file in your ETL code.'
            ' You can generate it on an existing dataset using petastorm-generate-metadata.py')
    return depickle_legacy_package_name_compatible(common[UNISCHEMA_KEY])


def get_schema_from_dataset_url(dataset_url_or_urls, hdfs_driver='libhdfs3', storage_options=None, filesystem=None):
    from petastorm_b200.fs_utils import get_filesystem_and_path_or_paths
    _, path_or_paths = get_filesystem_and_path_or_paths(dataset_url_or_urls, hdfs_driver, storage_options, filesystem)
    return get_schema(ParquetDataset(path_or_paths))


def infer_or_load_unischema(dataset):
    """Stored Unischema if present, else inferred from the parquet schema of the first file
    (petastorm/etl/dataset_metadata.py:410-418, petastorm/unischema.py:302-353)."""
    try:
        return get_schema(dataset)
    except PetastormMetadataError:
        logger.info('Failed loading Unischema from metadata in %s. Assuming the dataset was not created with '
                    'Petastorm. Will try to construct from native Parquet schema.', dataset.paths)
    first = dataset.first_file()
    partition_fields = [(name, dt) for name, _, dt in dataset.partitions.levels]
    return Unischema.from_parquet_schema(first.schema, partition_fields)


def update_common_metadata(dataset, key_values):
    """Adds / replaces key-value pairs of the dataset's ``_common_metadata`` (what ``utils.add_to_dataset_metadata``
    does for the reference, petastorm/utils.py:95-130).  ``key_values``: bytes -> bytes."""
    import pyarrow.parquet as pq
    if dataset.common_metadata_path is None:
        raise PetastormMetadataError('The dataset has no _common_metadata file to update')
    arrow_schema = pq.read_schema(dataset.common_metadata_path)
    meta = dict(arrow_schema.metadata or {})
    meta.update(key_values)
    pq.write_metadata(arrow_schema.with_metadata(meta), dataset.common_metadata_path)
    rowgroup.forget_file(dataset.common_metadata_path)
    dataset._common_kv = None  # pylint: disable=protected-access


def get_row_group_indexes(dataset):
    """Pickled value -> row-group indexes built by ``build_rowgroup_index`` (petastorm/etl/rowgroup_indexing.py:136-158)."""
    common = dataset.common_metadata
    if common is None or ROWGROUPS_INDEX_KEY not in common:
        raise PetastormMetadataError('Row-group indexes are not available in the dataset metadata')
    from petastorm_b200.etl.legacy import restricted_loads
    return restricted_loads(common[ROWGROUPS_INDEX_KEY])
