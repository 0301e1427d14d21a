"""Schema objects of the read path: :class:`UnischemaField`, :class:`Unischema` and helpers.

API mirror of ``petastorm/unischema.py`` (field tuple ``:50-85``, schema ``:174-356``, regex matching ``:437-464``,
parquet-type -> numpy map ``:467-502``) written for the B200 pipeline: the arrow-schema inference works from the
footer JSON produced by ``libpst_b200.so`` (no pyarrow dataset objects), and every field knows how its column is laid
out on the device.  Spark rendering (``as_spark_schema``) is out of scope (no JVM) and raises.

Pickle compatibility: datasets store a pickled ``petastorm.unischema.Unischema``; the restricted unpickler in
:mod:`petastorm_b200.etl.legacy` maps those globals onto the classes here, so attribute names (``_name``, ``_fields``)
are kept.
"""
import re
import warnings
from collections import OrderedDict, namedtuple
from decimal import Decimal

import numpy as np

# 'preserve_input_order' (default) or 'alphabetical' (legacy) - same switch as petastorm/unischema.py:33-36
_UNISCHEMA_FIELD_ORDER = 'preserve_input_order'


def _alphabetical():
    return _UNISCHEMA_FIELD_ORDER.lower() == 'alphabetical'


_FieldBase = namedtuple('UnischemaField', ['name', 'numpy_dtype', 'shape', 'codec', 'nullable'])


class UnischemaField(_FieldBase):
    """One field: ``(name, numpy_dtype, shape, codec=None, nullable=False)``.

    ``shape`` uses ``None`` for variable dimensions.  Equality and hashing ignore the codec, like the reference
    (petastorm/unischema.py:39-47,71-85): codec instances change identity when unpickled."""

    __slots__ = ()

    def __new__(cls, name, numpy_dtype, shape, codec=None, nullable=False):
        return super(UnischemaField, cls).__new__(cls, name, numpy_dtype, shape, codec, nullable)

    def _identity(self):
        return (self.name, self.numpy_dtype, self.shape, self.nullable)

    def __eq__(self, other):
        try:
            return self._identity() == (other.name, other.numpy_dtype, other.shape, other.nullable)
        except AttributeError:
            return False

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(self._identity())


class _NamedtupleCache(object):
    """One namedtuple class per (schema name, field names) so that types compare equal across calls
    (petastorm/unischema.py:88-111)."""
    _store = {}

    @staticmethod
    def get(parent_schema_name, field_names):
        names = sorted(field_names) if _alphabetical() else list(field_names)
        key = ' '.join([parent_schema_name] + names)
        cls = _NamedtupleCache._store.get(key)
        if cls is None:
            cls = namedtuple('{}_view'.format(parent_schema_name), names)
            _NamedtupleCache._store[key] = cls
        return cls


class Unischema(object):
    """An ordered collection of :class:`UnischemaField`; fields are also reachable as attributes."""

    def __init__(self, name, fields):
        self._name = name
        fields = list(fields)
        if _alphabetical():
            fields.sort(key=lambda f: f.name)
        self._fields = OrderedDict((f.name, f) for f in fields)
        for f in fields:
            if f.name in self.__dict__ or hasattr(type(self), f.name):
                warnings.warn('Can not create dynamic property {} because it conflicts with an existing property of '
                              'Unischema'.format(f.name))
            else:
                setattr(self, f.name, f)

    @property
    def fields(self):
        return self._fields

    def create_schema_view(self, fields):
        """Sub-schema from UnischemaField objects and/or full-match regex strings (petastorm/unischema.py:199-240).
        Field objects are looked up *by name*: the stored codec/shape wins over whatever the caller passed."""
        patterns = [f for f in fields if isinstance(f, str)]
        objects = [f for f in fields if isinstance(f, tuple)]
        if len(patterns) + len(objects) != len(fields):
            raise ValueError('Elements of "fields" must be either a string (regular expressions) or '
                             'an instance of UnischemaField class.')
        unknown = set(f.name for f in objects) - set(self._fields.keys())
        if unknown:
            raise ValueError('field {} does not belong to the schema {}'.format(unknown, self))
        view = [self._fields[f.name] for f in objects] + match_unischema_fields(self, patterns)
        return Unischema('{}_view'.format(self._name), view)

    def _get_namedtuple(self):
        return _NamedtupleCache.get(self._name, self._fields.keys())

    def make_namedtuple(self, **kargs):
        """Namedtuple instance of this schema (petastorm/unischema.py:283-297)."""
        return self._get_namedtuple()(**kargs)

    def make_namedtuple_tf(self, *args, **kargs):
        return self._get_namedtuple()(*args, **kargs)

    def as_spark_schema(self):
        raise NotImplementedError('Spark rendering of a Unischema needs pyspark/JVM and is outside the B200 read path')

    def __str__(self):
        rows = ''.join("  {}('{}', {}, {}, {}, {}),\n".format(type(f).__name__, f.name,
                                                             getattr(f.numpy_dtype, '__name__', f.numpy_dtype),
                                                             f.shape, f.codec, f.nullable)
                       for f in self._fields.values())
        return '{}({}, [\n{}])'.format(type(self).__name__, self._name, rows)

    def __getattr__(self, item):
        # only reached for missing attributes; keeps unpickled (init-less) instances usable
        if item in ('_fields', '_name'):
            raise AttributeError(item)
        fields = self.__dict__.get('_fields')
        if fields is not None and item in fields:
            return fields[item]
        raise AttributeError(item)

    # ------------------------------------------------------------------------------------------------------------
    @classmethod
    def from_parquet_schema(cls, file_schema, partition_fields=(), omit_unsupported_fields=True):
        """Infer a Unischema from a plain Parquet file - the job of ``Unischema.from_arrow_schema``
        (petastorm/unischema.py:302-353): partition columns first, no codecs, list columns get shape ``(None,)``.

        :param file_schema: the ``schema`` dict of :class:`petastorm_b200.native.ParquetFile` (footer JSON)
        :param partition_fields: iterable of ``(name, numpy_dtype)`` of hive partition keys
        """
        fields = [UnischemaField(name, dtype, (), None, False) for name, dtype in partition_fields]
        tops = file_schema['top_level']
        leaves_by_top = {}
        for leaf in file_schema['leaves']:
            leaves_by_top.setdefault(leaf['top_index'], []).append(leaf)
        for ti, top in enumerate(tops):
            leaves = leaves_by_top.get(ti, [])
            name = top['name']
            if top['num_children'] == 0:
                leaf = leaves[0]
                shape = ()
                nullable = leaf['repetition'] == 1
            else:
                is_list = top['converted_type'] == 3 or top['logical_kind'] == 3
                if not is_list or len(leaves) != 1 or leaves[0]['max_rep'] != 1:
                    warnings.warn('[ARROW-1644] Ignoring unsupported structure for field %r' % name)
                    continue
                leaf = leaves[0]
                shape = (None,)
                nullable = top['repetition'] == 1
            try:
                np_type = numpy_dtype_of_leaf(leaf)
            except ValueError:
                if omit_unsupported_fields:
                    warnings.warn('Column %r has an unsupported field type. Ignoring...' % name)
                    continue
                raise
            fields.append(UnischemaField(name, np_type, shape, None, nullable))
        return cls('inferred_schema', fields)


def match_unischema_fields(schema, field_regex):
    """Fields whose *entire* name matches one of the patterns (petastorm/unischema.py:437-464); warns when the legacy
    prefix-match semantics would have selected a different set."""
    if not field_regex:
        return []
    full, legacy = [], set()
    seen = set()
    for pattern in field_regex:
        for name, field in schema.fields.items():
            if re.fullmatch(pattern, name) and name not in seen:
                seen.add(name)
                full.append(field)
            if re.match(pattern, name):
                legacy.add(name)
    if seen != legacy:
        diff = sorted((seen | legacy) - (seen & legacy))
        warnings.warn('schema_fields behavior has changed. Now, regular expression pattern must match the entire '
                      'field name. The change in the behavior affects the following fields: {}'.format(', '.join(diff)))
    return full


# parquet ConvertedType values
_CT_UTF8, _CT_DECIMAL, _CT_DATE = 0, 5, 6
_CT_TIME_MILLIS, _CT_TIME_MICROS, _CT_TS_MILLIS, _CT_TS_MICROS = 7, 8, 9, 10
_CT_UINT8, _CT_UINT16, _CT_UINT32, _CT_UINT64, _CT_INT8, _CT_INT16, _CT_INT32, _CT_INT64 = 11, 12, 13, 14, 15, 16, 17, 18


def integer_logical_type(leaf):
    """(bits, signed) of an INT32/INT64 leaf, honouring converted/logical INTEGER annotations."""
    ct = leaf['converted_type']
    if leaf.get('logical_kind') == 10 and leaf.get('int_bits'):
        return leaf['int_bits'], bool(leaf['int_signed'])
    table = {_CT_UINT8: (8, False), _CT_UINT16: (16, False), _CT_UINT32: (32, False), _CT_UINT64: (64, False),
             _CT_INT8: (8, True), _CT_INT16: (16, True), _CT_INT32: (32, True), _CT_INT64: (64, True)}
    if ct in table:
        return table[ct]
    return (32 if leaf['physical_type'] == 1 else 64), True


def numpy_dtype_of_leaf(leaf):
    """numpy type of a parquet leaf, the counterpart of ``_numpy_and_codec_from_arrow_type``
    (petastorm/unischema.py:467-502) expressed on parquet physical + logical types."""
    pt = leaf['physical_type']
    ct = leaf['converted_type']
    lk = leaf.get('logical_kind', 0)
    if ct == _CT_DECIMAL or lk == 5:
        return Decimal
    if pt == 0:
        return np.bool_
    if pt in (1, 2):
        if ct == _CT_DATE or lk == 6 or ct in (_CT_TS_MILLIS, _CT_TS_MICROS) or lk == 8:
            return np.datetime64
        if ct in (_CT_TIME_MILLIS, _CT_TIME_MICROS) or lk == 7:
            raise ValueError('time-of-day columns are not supported')
        bits, signed = integer_logical_type(leaf)
        if not signed and bits > 8:
            # the reference's arrow -> numpy map has no uint16/uint32/uint64 branch: such columns are reported as
            # unsupported and dropped from the inferred schema (petastorm/unischema.py:467-502, :343-349)
            raise ValueError('Cannot auto-create unischema due to unsupported column type uint{}'.format(bits))
        return {(8, True): np.int8, (8, False): np.uint8, (16, True): np.int16, (16, False): np.uint16,
                (32, True): np.int32, (32, False): np.uint32, (64, True): np.int64, (64, False): np.uint64}[(bits, signed)]
    if pt == 3:
        return np.datetime64  # INT96 legacy timestamps
    if pt == 4:
        return np.float32
    if pt == 5:
        return np.float64
    if pt == 6:
        if ct == _CT_UTF8 or lk == 1:
            return np.str_
        return np.bytes_
    if pt == 7:
        if lk == 15:
            raise ValueError('Cannot auto-create unischema due to unsupported column type halffloat')
        return np.bytes_
    raise ValueError('Cannot auto-create unischema due to unsupported column type {}'.format(pt))


def insert_explicit_nulls(unischema, row_dict):
    """Adds ``None`` for missing nullable fields; missing non-nullable fields raise (petastorm/unischema.py:409-424)."""
    for name, field in unischema.fields.items():
        if name not in row_dict:
            if field.nullable:
                row_dict[name] = None
            else:
                raise ValueError('Field {} is not found in the row_dict, but is not nullable.'.format(name))
