"""Row predicates (API of petastorm/predicates.py:27-182) with device implementations.

Every predicate keeps the reference's per-row contract ``do_include(values) -> bool`` (``values`` maps field name to a
decoded value) - that is what a user-defined ``PredicateBase`` implements and what the host fallback for *user code*
calls.  The built-in predicates additionally implement ``device_mask(columns)``: given ``{field: CUDA tensor of the
row-group column}`` they return a uint8 CUDA mask computed by hand-written kernels (``in_set`` -> binary search,
``in_pseudorandom_split`` -> MD5 on device, ``in_negate`` / ``in_reduce`` -> mask algebra).  ``device_mask`` returns
``None`` when a predicate (or one of its children) has no device form for the given column types; the worker then
evaluates ``do_include`` row by row on the host, exactly like the reference.
"""
import abc
import collections.abc
import hashlib
import sys

import numpy as np


class PredicateBase(abc.ABC):
    """Base class for row predicates."""

    @abc.abstractmethod
    def get_fields(self):
        """Set of field names the predicate needs."""

    @abc.abstractmethod
    def do_include(self, values):
        """True if the row should be kept."""

    def device_mask(self, columns):  # pylint: disable=unused-argument
        """uint8 CUDA mask over the row-group, or None if only the host form exists."""
        return None


def _string_to_bucket(string, bucket_num):
    """md5(string) as an integer modulo ``bucket_num`` (petastorm/predicates.py:39-41)."""
    return int(hashlib.md5(string.encode('utf-8')).hexdigest(), 16) % bucket_num


def _is_int_tensor(t):
    import torch
    return hasattr(t, 'is_cuda') and t.is_cuda and t.dim() == 1 and t.dtype in (
        torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8, torch.uint16, torch.uint32)


class in_set(PredicateBase):
    """``values[field] in inclusion_values``."""

    def __init__(self, inclusion_values, predicate_field):
        self._inclusion_values = set(inclusion_values)
        self._predicate_field = predicate_field
        self._device_set = None
        self._sorted_ints = None

    def __getstate__(self):
        # pickle-compatible with the reference class (two attributes); the device caches are rebuilt on demand
        return {'_inclusion_values': self._inclusion_values, '_predicate_field': self._predicate_field}

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._device_set = None
        self._sorted_ints = None

    def get_fields(self):
        return {self._predicate_field}

    def do_include(self, values):
        return values[self._predicate_field] in self._inclusion_values

    def device_mask(self, columns):
        import torch
        from petastorm_b200 import device_ops
        col = columns.get(self._predicate_field)
        if col is None or not _is_int_tensor(col):
            return None
        if self._sorted_ints is None:      # once per predicate: the set may hold millions of keys
            try:
                ints = sorted(int(v) for v in self._inclusion_values
                              if isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)))
            except (TypeError, ValueError):
                ints = []
            # non-integer members: python equality semantics differ, stay on the host form
            self._sorted_ints = False if len(ints) != len(self._inclusion_values) else \
                np.asarray([v for v in ints if -2 ** 63 <= v < 2 ** 63], dtype=np.int64)
        if self._sorted_ints is False:
            return None
        if self._device_set is None or self._device_set.device != col.device:
            self._device_set = torch.from_numpy(self._sorted_ints).to(col.device)
        return device_ops.mask_in_set(col.contiguous(), self._device_set)


class in_intersection(PredicateBase):
    """The list-valued field shares at least one element with ``inclusion_values``."""

    def __init__(self, inclusion_values, _predicate_field):
        self._inclusion_values = list(inclusion_values)
        self._predicate_field = _predicate_field

    def get_fields(self):
        return {self._predicate_field}

    def do_include(self, values):
        value = values[self._predicate_field]
        if not isinstance(value, collections.abc.Iterable):
            raise ValueError('Predicate field should have iterable type')
        return bool(np.isin(value, self._inclusion_values).any())


class in_lambda(PredicateBase):
    """User function over a list of fields; optional ``state_arg`` is appended to the arguments when not None."""

    def __init__(self, predicate_fields, predicate_func, state_arg=None):
        if not isinstance(predicate_fields, list):
            raise ValueError('Predicate fields should be a list')
        self._predicate_fields = predicate_fields
        self._predicate_func = predicate_func
        self._state_arg = state_arg

    def get_fields(self):
        return set(self._predicate_fields)

    def do_include(self, values):
        args = [values[f] for f in self._predicate_fields]
        if self._state_arg is not None:
            args.append(self._state_arg)
        return self._predicate_func(*args)


class in_negate(PredicateBase):
    """Logical not of another predicate."""

    def __init__(self, predicate):
        if not isinstance(predicate, PredicateBase):
            raise ValueError('Predicate is nor derived from PredicateBase')
        self._predicate = predicate

    def get_fields(self):
        return self._predicate.get_fields()

    def do_include(self, values):
        return not self._predicate.do_include(values)

    def device_mask(self, columns):
        m = self._predicate.device_mask(columns)
        return None if m is None else (m ^ 1)


class in_reduce(PredicateBase):
    """Combine predicates with a reduction such as ``all`` or ``any``."""

    def __init__(self, predicate_list, reduce_func):
        if not all(isinstance(p, PredicateBase) for p in predicate_list):
            raise ValueError('Predicate is nor derived from PredicateBase')
        self._predicate_list = predicate_list
        self._reduce_func = reduce_func

    def get_fields(self):
        fields = set()
        for p in self._predicate_list:
            fields |= p.get_fields()
        return fields

    def do_include(self, values):
        return self._reduce_func([p.do_include(values) for p in self._predicate_list])

    def device_mask(self, columns):
        if self._reduce_func not in (all, any) or not self._predicate_list:
            return None
        masks = [p.device_mask(columns) for p in self._predicate_list]
        if any(m is None for m in masks):
            return None
        out = masks[0]
        for m in masks[1:]:
            out = (out & m) if self._reduce_func is all else (out | m)
        return out


class in_pseudorandom_split(PredicateBase):
    """Deterministic dataset split: hash the key with MD5 into ``[0, sys.maxsize)`` and keep the rows whose bucket
    falls into the ``subset_index``-th fraction (petastorm/predicates.py:144-182)."""

    def __init__(self, fraction_list, subset_index, predicate_field):
        if subset_index >= len(fraction_list):
            raise ValueError('subset_index is out of range')
        self._predicate_field = predicate_field
        highs = [sum(fraction_list[:i + 1]) for i in range(len(fraction_list))]
        low = highs[subset_index - 1] if subset_index else 0
        self._bucket_low = low * (sys.maxsize - 1)
        self._bucket_high = highs[subset_index] * (sys.maxsize - 1)

    def get_fields(self):
        return {self._predicate_field}

    def do_include(self, values):
        if self._predicate_field not in values.keys():
            raise ValueError('Tested values does not have split key: %s' % self._predicate_field)
        bucket = _string_to_bucket(str(values[self._predicate_field]), sys.maxsize)
        return self._bucket_low <= bucket < self._bucket_high

    def device_mask(self, columns):
        from petastorm_b200 import device_ops
        col = columns.get(self._predicate_field)
        if col is None or not _is_int_tensor(col):
            return None  # string keys are hashed on the host (they are host objects anyway)
        return device_ops.mask_md5_split(col.contiguous(), self._bucket_low, self._bucket_high)
