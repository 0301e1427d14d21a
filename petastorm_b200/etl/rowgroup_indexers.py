"""Row-group indexers: value -> set of row-group ordinals, stored pickled in ``_common_metadata`` under
``dataset-toolkit.rowgroups_index.v1`` and consumed by the selectors (``petastorm_b200/selectors.py``).

Read-side contract of ``petastorm/etl/rowgroup_indexers.py:21-128``: datasets indexed by the reference unpickle onto
these classes (``etl/legacy.py`` maps the module name), and ``build_rowgroup_index`` below writes pickles the reference
can load in turn.  The attribute names (``_index_name``, ``_column_name``, ``_index_data``) are part of that pickle
format.
"""
import collections

import numpy as np


class RowGroupIndexerBase(object):
    """index_name / column_names / indexed_values / get_row_group_indexes / build_index, and ``a + b`` to merge the
    partial indexes of two sets of row-groups."""

    def __init__(self, index_name, index_field):
        self._index_name = index_name
        self._column_name = index_field

    @property
    def index_name(self):
        return self._index_name

    @property
    def column_names(self):
        return [self._column_name]

    def _check_mergeable(self, other):
        if type(other) is not type(self):
            raise TypeError('Make sure Spark map function return the same indexer type')
        if self._column_name != other._column_name:  # pylint: disable=protected-access
            raise ValueError('Make sure indexers in Spark map function index the same fields')

    def _column(self, decoded_rows):
        values = [row[self._column_name] for row in decoded_rows]
        if not values:
            raise ValueError("Cannot build index for empty rows, column '{}'".format(self._column_name))
        return values


class SingleFieldIndexer(RowGroupIndexerBase):
    """Every distinct value of one field (array-valued fields: every element) -> the row-groups that contain it."""

    def __init__(self, index_name, index_field):
        super(SingleFieldIndexer, self).__init__(index_name, index_field)
        self._index_data = collections.defaultdict(set)

    def __add__(self, other):
        self._check_mergeable(other)
        for key, pieces in other._index_data.items():  # pylint: disable=protected-access
            self._index_data[key].update(pieces)
        return self

    @property
    def indexed_values(self):
        return list(self._index_data.keys())

    def get_row_group_indexes(self, value_key):
        return self._index_data[value_key]

    def build_index(self, decoded_rows, piece_index):
        for value in self._column(decoded_rows):
            if value is None:
                continue
            if isinstance(value, np.ndarray):
                for element in value.flatten().tolist():
                    self._index_data[element].add(piece_index)
            else:
                self._index_data[value.item() if isinstance(value, np.generic) else value].add(piece_index)
        return self._index_data


class FieldNotNullIndexer(RowGroupIndexerBase):
    """The row-groups in which a field is not null at least once."""

    def __init__(self, index_name, index_field):
        super(FieldNotNullIndexer, self).__init__(index_name, index_field)
        self._index_data = set()

    def __add__(self, other):
        self._check_mergeable(other)
        self._index_data.update(other._index_data)  # pylint: disable=protected-access
        return self

    @property
    def indexed_values(self):
        return ['Field is Not Null']

    def get_row_group_indexes(self, value_key=None):
        return self._index_data

    def build_index(self, decoded_rows, piece_index):
        if any(value is not None for value in self._column(decoded_rows)):
            self._index_data.add(piece_index)
        return self._index_data


def pickle_indexers_reference_compatible(index_dict):
    """``pickle.dumps({index_name: indexer})`` (protocol 2, like the reference writes) with this module spelled
    ``petastorm.etl.rowgroup_indexers`` so that the reference un-pickles the result onto its own classes."""
    import pickle
    data = pickle.dumps(index_dict, protocol=2)
    return data.replace(b'cpetastorm_b200.etl.rowgroup_indexers\n', b'cpetastorm.etl.rowgroup_indexers\n')


def build_rowgroup_index(dataset_url, indexers, reader_factory=None):
    """Builds the indexes by reading the indexed columns of every row-group (the reference does this with a Spark job,
    ``petastorm/etl/rowgroup_indexing.py:35-158``) and stores them in the dataset's ``_common_metadata``."""
    from petastorm_b200 import make_reader
    from petastorm_b200.etl import dataset_metadata as dm
    from petastorm_b200.fs_utils import get_filesystem_and_path_or_paths
    _, path = get_filesystem_and_path_or_paths(dataset_url)
    dataset = dm.ParquetDataset(path)
    pieces = dm.load_row_groups(dataset)
    schema = dm.get_schema(dataset)
    columns = sorted({c for ix in indexers for c in ix.column_names})
    for c in columns:
        if c not in schema.fields:
            raise ValueError('Indexed field {} is not part of the dataset schema'.format(c))
    factory = reader_factory or make_reader
    from petastorm_b200 import rowgroup
    piece_rows = [rowgroup.open_file(p.path).row_group_num_rows(p.row_group) for p in pieces]
    with factory(dataset_url, schema_fields=[schema.fields[c] for c in columns], shuffle_row_groups=False,
                 num_epochs=1, output='numpy') as reader:
        it = iter(reader)
        for piece_index, n in enumerate(piece_rows):
            rows = [next(it)._asdict() for _ in range(n)]
            if rows:
                for ix in indexers:
                    ix.build_index(rows, piece_index)
    index_dict = {ix.index_name: ix for ix in indexers}
    dm.update_common_metadata(dataset, {dm.ROWGROUPS_INDEX_KEY: pickle_indexers_reference_compatible(index_dict)})
    return index_dict
