"""Restricted un-pickling of the Unischema stored in ``_common_metadata``.

Role of ``petastorm/etl/legacy.py:22-79``: datasets carry ``pickle.dumps(Unischema)`` under the key
``dataset-toolkit.unischema.v1``; the pickle references ``petastorm.unischema.*``, ``petastorm.codecs.*`` (and the
package's two historical names), ``pyspark.sql.types.*``, numpy scalar types (including names removed in numpy 2:
``string_``, ``unicode_``), ``collections.OrderedDict``, ``decimal.Decimal`` and the py2 spellings ``copy_reg`` /
``__builtin__``.  Everything is resolved from an allow-list onto this package's classes; anything else is refused.
"""
import builtins
import copyreg
import importlib
import io
import pickle

import numpy as np

_LEGACY_PACKAGES = ('petastorm', 'av.experimental.deepdrive.dataset_toolkit', 'av.ml.dataset_toolkit')
_OWN_MODULES = {'unischema': 'petastorm_b200.unischema', 'codecs': 'petastorm_b200.codecs',
                'sequence': 'petastorm_b200.ngram', 'ngram': 'petastorm_b200.ngram',
                'transform': 'petastorm_b200.transform', 'predicates': 'petastorm_b200.predicates'}
_SAFE_BUILTINS = {'tuple', 'list', 'dict', 'set', 'frozenset', 'int', 'float', 'str', 'bytes', 'bool', 'object',
                  'complex', 'slice', 'range', 'bytearray'}
_PY2_BUILTINS = {'long': int, 'unicode': str, 'basestring': str}
_NUMPY_RENAMED = {'string_': np.bytes_, 'unicode_': np.str_, 'str': np.str_, 'float_': np.float64, 'bool8': np.bool_,
                  'object_': np.object_, 'int0': np.intp, 'uint0': np.uintp, 'bool': np.bool_}


def _restore_namedtuple(name, fields, value):
    """``pyspark.serializers._restore``: pyspark hijacks namedtuple pickling, so UnischemaField instances written from a
    Spark driver are stored as ``_restore('UnischemaField', field_names, values)``."""
    if name == 'UnischemaField':
        from petastorm_b200.unischema import UnischemaField
        return UnischemaField(**dict(zip(fields, value)))
    import collections
    return collections.namedtuple(name, fields)(*value)


class RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == 'pyspark.serializers' and name == '_restore':
            return _restore_namedtuple
        for pkg in _LEGACY_PACKAGES:
            if module == pkg + '.etl.rowgroup_indexers':      # row-group indexes (dataset-toolkit.rowgroups_index.v1)
                return getattr(importlib.import_module('petastorm_b200.etl.rowgroup_indexers'), name)
            if module.startswith(pkg + '.'):
                sub = module[len(pkg) + 1:].split('.')[0]
                if sub in _OWN_MODULES:
                    return getattr(importlib.import_module(_OWN_MODULES[sub]), name)
                raise pickle.UnpicklingError("global '%s.%s' is forbidden" % (module, name))
        root = module.split('.')[0]
        if root == 'pyspark':
            from petastorm_b200 import spark_types
            return spark_types.resolve(name)
        if root == 'numpy':
            if module == 'numpy' and name in _NUMPY_RENAMED and not hasattr(np, name):
                return _NUMPY_RENAMED[name]
            target = module.replace('numpy.core', 'numpy._core') if module.startswith('numpy.core') else module
            try:
                mod = importlib.import_module(target)
            except ImportError:
                mod = importlib.import_module(module)
            return getattr(mod, name)
        if module in ('collections', 'decimal'):
            return getattr(importlib.import_module(module), name)
        if module in ('builtins', '__builtin__'):
            if name in _PY2_BUILTINS:
                return _PY2_BUILTINS[name]
            if name in _SAFE_BUILTINS:
                return getattr(builtins, name)
            raise pickle.UnpicklingError("global '%s.%s' is forbidden" % (module, name))
        if module == '_codecs' and name == 'encode':
            import _codecs          # protocol-2 pickles of numpy scalars / bytes written by python 3 go through it
            return _codecs.encode
        if module in ('copy_reg', 'copyreg'):
            if name in ('_reconstructor', '__newobj__', '__newobj_ex__'):
                return getattr(copyreg, name)
            raise pickle.UnpicklingError("global '%s.%s' is forbidden" % (module, name))
        raise pickle.UnpicklingError("global '%s.%s' is forbidden" % (module, name))


def restricted_loads(data):
    try:
        return RestrictedUnpickler(io.BytesIO(data)).load()
    except UnicodeDecodeError:
        return RestrictedUnpickler(io.BytesIO(data), encoding='latin1').load()


def depickle_legacy_package_name_compatible(pickled_string):
    """Unpickle a stored Unischema whatever package name it was written under (petastorm/etl/legacy.py:54-79)."""
    return restricted_loads(pickled_string)
