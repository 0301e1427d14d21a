"""``TransformSpec`` and the schema it implies (API of petastorm/transform.py:19-89).

On the B200 path a transform is applied to a whole decoded row-group on the device:

* :class:`DeviceTransform` subclasses (e.g. :class:`Normalize`) run hand-written kernels;
* a plain callable marked with ``device=True`` (``TransformSpec(func, ..., device=True)``) receives a dict of batched
  CUDA tensors - the batch-reader convention of the reference, where ``func`` gets the whole row-group
  (petastorm/arrow_reader_worker.py:250-251);
* any other callable is *user* host code and gets exactly what the reference would pass (a row dict for
  ``make_reader``, a pandas DataFrame for ``make_batch_reader``).
"""
import warnings

from petastorm_b200.unischema import Unischema, UnischemaField


def edit_field(name, numpy_dtype, shape, nullable=False):
    """4-tuple ``(name, numpy_dtype, shape, is_nullable)`` for ``TransformSpec.edit_fields``."""
    return name, numpy_dtype, shape, nullable


class TransformSpec(object):
    def __init__(self, func=None, edit_fields=None, removed_fields=None, selected_fields=None, device=None):
        if removed_fields is not None and selected_fields is not None:
            raise ValueError('User can only specify one of removed_fields and selected_fields in TransformSpec.')
        self.func = func
        self.edit_fields = edit_fields or []
        self.removed_fields = removed_fields or []
        self.selected_fields = selected_fields
        # device=None: decide from the callable (DeviceTransform instances run on the device)
        self.device = isinstance(func, DeviceTransform) if device is None else bool(device)


def transform_schema(schema, transform_spec):
    """Post-transform schema: drop edited+removed fields, append edited fields (codec ``None``), then optionally
    select/reorder (petastorm/transform.py:60-89)."""
    removed = set(transform_spec.removed_fields)
    unknown = removed - set(schema.fields.keys())
    if unknown:
        warnings.warn('remove_fields specified some field names that are not part of the schema. '
                      'These field names will be ignored "{}". '.format(', '.join(unknown)))
    dropped = {e[0] for e in transform_spec.edit_fields} | removed
    fields = [f for name, f in schema.fields.items() if name not in dropped]
    fields.extend(UnischemaField(name=e[0], numpy_dtype=e[1], shape=e[2], codec=None, nullable=e[3])
                  for e in transform_spec.edit_fields)
    if transform_spec.selected_fields is not None:
        wanted = list(transform_spec.selected_fields)
        unknown = set(wanted) - set(f.name for f in fields)
        if unknown:
            warnings.warn('selected_fields specified some field names that are not part of the schema. '
                          'These field names will be ignored "{}". '.format(', '.join(unknown)))
        fields = sorted((f for f in fields if f.name in wanted), key=lambda f: wanted.index(f.name))
    return Unischema(schema._name + '_transformed', fields)


class DeviceTransform(object):
    """A transform with a device implementation: called with ``{field: batched CUDA tensor}``, returns the same."""

    def __call__(self, columns):
        raise NotImplementedError


class Normalize(DeviceTransform):
    """``out = ((float32)x - mean) / std`` cast to ``out_dtype`` for the given fields (K12 kernel; bit-identical to the
    numpy expression ``((x.astype(np.float32) - mean) / std).astype(out_dtype)`` a reference user would write)."""

    def __init__(self, fields, mean, std, out_dtype='float32'):
        self.fields = [fields] if isinstance(fields, str) else list(fields)
        self.mean = float(mean)
        self.std = float(std)
        self.out_dtype = str(out_dtype)

    def __call__(self, columns):
        import torch
        from petastorm_b200 import device_ops
        if isinstance(columns, dict) and self.fields and not hasattr(columns[self.fields[0]], 'is_cuda'):
            # row-dict (host) form: same arithmetic in numpy so that host- and device-side users agree bit for bit
            import numpy as np
            for name in self.fields:
                columns[name] = ((columns[name].astype(np.float32) - np.float32(self.mean)) /
                                 np.float32(self.std)).astype(self.out_dtype)
            return columns
        tdt = getattr(torch, self.out_dtype)
        for name in self.fields:
            columns[name] = device_ops.normalize(columns[name].contiguous(), self.mean, self.std, tdt)
        return columns
