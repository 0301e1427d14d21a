"""PyTorch loaders over a Reader (API of petastorm/pytorch.py: ``DataLoader`` :131-256, ``BatchedDataLoader`` :259-370,
``InMemBatchedDataLoader`` :437-501, ``decimal_friendly_collate`` :73-95, ``_sanitize_pytorch_types`` :40-70).

With a petastorm_b200 reader the rows are already device tensors, so the loaders differ from upstream in *where* data
lives, not in what they yield:

* :class:`DataLoader` and :class:`BatchedDataLoader` consume whole decoded row-groups: a device-resident batched
  shuffling buffer (``torch.randperm`` + the row-gather kernel) hands out ``{field: tensor[batch_size, ...]}``; NGram
  windows are gathered by ``pst_ngram_gather`` into ``[W, L, ...]`` tensors.  ``DataLoader`` falls back to the
  reference's row loop (``collate_fn`` over a list of row dicts) for a custom ``collate_fn`` or host-valued fields.
* :class:`InMemBatchedDataLoader` decodes once into HBM and reshuffles per epoch with ``torch.Generator(seed+epoch)``.
"""
import collections.abc
import decimal
import logging
import re

import numpy as np
import torch
from torch.utils.data.dataloader import default_collate

from petastorm_b200.reader_impl.pytorch_shuffling_buffer import (BatchedNoopShufflingBuffer,
                                                                 BatchedRandomShufflingBuffer)
from petastorm_b200.reader_impl.shuffling_buffer import NoopShufflingBuffer, RandomShufflingBuffer

logger = logging.getLogger(__name__)
_string_classes = (str, bytes)

_TORCH_PROMOTIONS = {torch.uint16: torch.int32, torch.uint32: torch.int64, torch.bool: torch.uint8}


def _sanitize_pytorch_types(row_as_dict):
    """In-place dtype promotions PyTorch needs: uint16->int32, uint32->int64, bool->uint8; string/object arrays and
    ``None`` are errors (petastorm/pytorch.py:40-70).  CUDA tensors are promoted by the K16 kernel."""
    for name, value in row_as_dict.items():
        if isinstance(value, torch.Tensor):
            if value.dtype in _TORCH_PROMOTIONS:
                if value.is_cuda:
                    from petastorm_b200 import device_ops
                    row_as_dict[name] = device_ops.sanitize(value)
                else:
                    row_as_dict[name] = value.to(_TORCH_PROMOTIONS[value.dtype])
        elif isinstance(value, np.ndarray):
            if value.dtype == np.uint16:
                row_as_dict[name] = value.astype(np.int32)
            elif value.dtype == np.uint32:
                row_as_dict[name] = value.astype(np.int64)
            elif value.dtype == np.bool_:
                row_as_dict[name] = value.astype(np.uint8)
            elif re.search('[SaUO]', value.dtype.str):
                raise TypeError('Pytorch does not support arrays of string or object classes. '
                                'Found in field {}.'.format(name))
        elif isinstance(value, np.bool_):
            row_as_dict[name] = np.uint8(value)
        elif value is None:
            raise TypeError('Pytorch does not support nullable fields. Found None in {}'.format(name))
        elif isinstance(value, list) and any(v is None for v in value):
            raise TypeError('Pytorch does not support nullable fields. Found None in {}'.format(name))


def decimal_friendly_collate(batch):
    """``default_collate`` that passes ``decimal.Decimal`` and strings through as lists (petastorm/pytorch.py:73-95)."""
    first = batch[0]
    if isinstance(first, decimal.Decimal):
        return batch
    if isinstance(first, collections.abc.Mapping):
        return {key: decimal_friendly_collate([d[key] for d in batch]) for key in first}
    if isinstance(first, _string_classes):
        return batch
    if isinstance(first, collections.abc.Sequence):
        return [decimal_friendly_collate(samples) for samples in zip(*batch)]
    return default_collate(batch)


_PARALLEL_ITER_ERROR = "You must finish a full pass of Petastorm DataLoader before making another pass from the \
beginning.If you do need to terminate early and restart from beginning, please re-create the reader and the data \
loader."


class LoaderBase(object):
    """Re-iteration rules of petastorm/pytorch.py:103-128: a second pass resets the reader, a concurrent pass or a pass
    after a failed one raises."""

    def __init__(self):
        self._in_iter = None
        self._error = None

    def __iter__(self):
        if self._error is not None:
            raise RuntimeError('Cannot start a new iteration because last time iteration failed with error {err}.'
                               .format(err=repr(self._error)))
        if self._in_iter is not None and self._in_iter == True:  # noqa: E712
            raise RuntimeError(_PARALLEL_ITER_ERROR)
        if self._in_iter is not None:
            self.reader.reset()
            logger.warning('Start a new pass of Petastorm DataLoader, reset underlying Petastorm reader to position 0.')
        self._in_iter = True
        try:
            for batch in self._iter_impl():
                yield batch
        except Exception as e:
            self._error = e
            logger.error('Iteration on Petastorm DataLoader raise error: %s', repr(e))
            raise
        finally:
            self._in_iter = False

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.reader.stop()
        self.reader.join()


def _is_tensor_column(value):
    from petastorm_b200.gpu_workers import ScalarColumn
    return isinstance(value, (torch.Tensor, ScalarColumn))


def _as_tensor(value):
    from petastorm_b200.gpu_workers import ScalarColumn
    return value.tensor if isinstance(value, ScalarColumn) else value


class _DeviceGroups(object):
    """Whole decoded row-groups of a petastorm_b200 reader as ``{field: tensor[n, ...]}`` (plain rows / batch reader) or
    :class:`~petastorm_b200.gpu_workers.NGramColumns` (``[W, L, ...]`` per field) - what the device-batched paths of
    :class:`DataLoader` and :class:`BatchedDataLoader` consume instead of per-row namedtuples.

    ``available()`` peeks at the first row-group: the device path needs a petastorm_b200 reader and tensor-valued
    columns only (strings, Decimals, nulls and ragged arrays keep the row-at-a-time path of the reference)."""

    def __init__(self, reader):
        self.reader = reader
        self._qr = getattr(reader, '_results_queue_reader', None)
        self._pool = getattr(reader, '_workers_pool', None)

    def available(self):
        from petastorm_b200.gpu_workers import GpuNGramWindows
        if self._qr is None or not hasattr(self._qr, 'peek_rowgroup') or getattr(self._qr, '_output', 'torch') != 'torch':
            return False
        first = self._qr.peek_rowgroup(self._pool)
        if first is None:
            return True            # no data at all: either path yields nothing
        if isinstance(first, GpuNGramWindows):
            names = set(first.ngram.get_field_names_at_all_timesteps_names())
            return all(_is_tensor_column(v) for k, v in first.rows.columns.items() if k in names)
        return all(_is_tensor_column(v) for v in first.values())

    def __iter__(self):
        from petastorm_b200.gpu_workers import GpuNGramWindows
        while True:
            try:
                cols = self._qr.read_next_rowgroup(self._pool, raw=True)
            except StopIteration:
                self.reader.last_row_consumed = True
                return
            if isinstance(cols, GpuNGramWindows):
                self._raise_for_host_values(dict(cols.rows.columns))
            if not all(_is_tensor_column(v) for v in cols.values()):
                self._raise_for_host_values(cols)
            yield cols

    @staticmethod
    def _raise_for_host_values(cols):
        """A later row-group brought values without a tensor form (a null, a ragged array): report it the way the
        row path would (``_sanitize_pytorch_types`` raises for None and string/object arrays)."""
        n = min(len(v) for v in cols.values()) if cols else 0
        for i in range(n):
            row = {}
            for k, v in cols.items():
                item = v[i]
                row[k] = item.cpu().numpy() if isinstance(item, torch.Tensor) else item
            _sanitize_pytorch_types(row)
        raise TypeError('Pytorch does not support the values of this row-group (ragged or non-numeric arrays): '
                        'fields {}'.format([k for k, v in cols.items() if not _is_tensor_column(v)]))


class _NGramBatch(dict):
    """``{offset: {field: tensor[batch, ...]}}`` over the ``[batch, L, ...]`` field tensors of a batch of NGram windows.
    The per-offset dicts are views (no copies) and are only built when an offset is looked at: an NGram of length 16
    over 13 fields would otherwise cost 208 tensor objects per batch whether the training loop reads them or not."""

    def __init__(self, fields, timesteps):
        super(_NGramBatch, self).__init__()
        self._fields = fields
        self._timesteps = timesteps
        self._base = min(timesteps)

    def __missing__(self, offset):
        if offset not in self._timesteps:
            raise KeyError(offset)
        item = {name: self._fields[name][:, offset - self._base] for name in self._timesteps[offset]
                if name in self._fields}
        dict.__setitem__(self, offset, item)
        return item

    def __contains__(self, offset):
        return offset in self._timesteps

    def __iter__(self):
        return iter(self._timesteps)

    def __len__(self):
        return len(self._timesteps)

    def keys(self):
        return self._timesteps.keys()

    def values(self):
        return [self[k] for k in self._timesteps]

    def items(self):
        return [(k, self[k]) for k in self._timesteps]

    def get(self, offset, default=None):
        return self[offset] if offset in self._timesteps else default

    def __repr__(self):
        return repr(dict(self.items()))

    @property
    def windows(self):
        """``{field: tensor[batch, L, ...]}`` - the whole windows, one tensor per field."""
        return self._fields


def _nest_ngram_batch(keys, values, timesteps):
    """``[W, L, ...]`` field tensors of a batch -> ``{offset: {field: tensor[W, ...]}}`` (lazy views, no copies)."""
    return _NGramBatch(dict(zip(keys, values)), timesteps)


class DataLoader(LoaderBase):
    """``petastorm.pytorch.DataLoader`` (petastorm/pytorch.py:131-256): batches of ``batch_size`` rows, optional
    shuffling queue, ``collate_fn`` to merge rows, last batch possibly partial.

    Over a petastorm_b200 reader with the default ``collate_fn`` and tensor-valued fields the loader never sees a row:
    whole decoded row-groups go into a device-resident shuffling buffer (``torch.randperm`` + the row-gather kernel)
    and batches are sliced / gathered on the device - the result is what ``decimal_friendly_collate`` returns for the
    same rows (``{field: tensor[batch, ...]}``, promoted dtypes), with every field on the device.  A custom
    ``collate_fn``, host-valued fields (strings, Decimals) or a foreign reader take the reference's row loop.
    NGram readers (an extension: upstream's loader has no NGram support) yield ``{offset: {field: tensor[batch, ...]}}``.
    """

    def __init__(self, reader, batch_size=1, collate_fn=decimal_friendly_collate, shuffling_queue_capacity=0):
        super(DataLoader, self).__init__()
        self.reader = reader
        self.batch_size = batch_size
        self.collate_fn = collate_fn
        self._batch_acc = []
        self.shuffling_queue_capacity = shuffling_queue_capacity
        self._in_iter = None
        self.device_batched = None     # set by the first pass: True when the device path was taken

    def _iter_impl(self):
        groups = _DeviceGroups(self.reader)
        if self.collate_fn is decimal_friendly_collate and groups.available():
            self.device_batched = True
            return _iter_device_batches(groups, self.batch_size, self.shuffling_queue_capacity, None)
        self.device_batched = False
        return self._iter_rows()

    def _iter_rows(self):
        keys = None
        if self.shuffling_queue_capacity > 0:
            self._shuffling_buffer = RandomShufflingBuffer(self.shuffling_queue_capacity,
                                                           min_after_retrieve=self.shuffling_queue_capacity - 1,
                                                           extra_capacity=100000000)
        else:
            self._shuffling_buffer = NoopShufflingBuffer()
        for row in self.reader:
            if isinstance(row, dict):
                # NGram window {offset: namedtuple} (extension: upstream's torch loader has no NGram support)
                row_as_dict = {offset: item._asdict() for offset, item in row.items()}
                for item in row_as_dict.values():
                    _sanitize_pytorch_types(item)
            else:
                row_as_dict = row._asdict()
                _sanitize_pytorch_types(row_as_dict)
            keys = row_as_dict.keys()
            if not self.reader.batched_output:
                self._shuffling_buffer.add_many([row_as_dict])
            else:
                # a batched reader hands out whole row-groups: transpose to per-row tuples (upstream :207-216)
                self._shuffling_buffer.add_many(list(zip(*(row_as_dict[k] for k in keys))))
            for batch in self._yield_batches(keys):
                yield batch
        self._shuffling_buffer.finish()
        for batch in self._yield_batches(keys):
            yield batch
        if self._batch_acc:
            yield self.collate_fn(self._batch_acc)
            self._batch_acc = []

    def _yield_batches(self, keys):
        while self._shuffling_buffer.can_retrieve():
            item = self._shuffling_buffer.retrieve()
            if not isinstance(item, dict):
                item = dict(zip(keys, item))
            self._batch_acc.append(item)
            if len(self._batch_acc) == self.batch_size:
                yield self.collate_fn(self._batch_acc)
                self._batch_acc = []


class _RowPacker(object):
    """Packs the narrow columns of a row-group (scalars, short vectors, NGram windows of scalars: <= ``SMALL`` bytes per
    row) into ONE ``uint8 [n, R]`` tensor, so that the shuffling buffer moves a row with one gather instead of one per
    field (a C5 window has 13 fields; the buffer is launch-bound otherwise).  A batch is split back into typed views
    of the packed batch buffer (no copies): narrow fields come out as strided views ``[batch, ...]``.  Wide columns
    (images, tensors) stay separate tensors."""

    SMALL = 256

    def __init__(self, columns):
        self.keys = list(columns.keys())
        self.small, self.big = [], []
        off = 0
        for k in self.keys:
            t = columns[k]
            item = t.element_size()
            nb = item * int(np.prod(t.shape[1:], dtype=np.int64))
            if 0 < nb <= self.SMALL:
                off = (off + item - 1) // item * item
                self.small.append((k, off, nb, t.dtype, tuple(t.shape[1:])))
                off += nb
            else:
                self.big.append(k)
        if len(self.small) < 2:          # nothing to gain: keep every column as it is
            self.big = self.keys
            self.small = []
        self.row_bytes = (off + 15) // 16 * 16

    def pack(self, columns):
        """list of buffer columns: [packed narrow fields (if any)] + wide columns"""
        out = []
        if self.small:
            n = columns[self.small[0][0]].shape[0]
            packed = torch.empty((n, self.row_bytes), dtype=torch.uint8, device=columns[self.small[0][0]].device)
            for k, off, nb, dtype, shape in self.small:
                t = columns[k]
                if t.dtype != dtype or tuple(t.shape[1:]) != shape:
                    raise TypeError('field {} changed from {}{} to {}{} between row-groups'.format(
                        k, dtype, shape, t.dtype, tuple(t.shape[1:])))
                packed[:, off:off + nb] = t.contiguous().reshape(n, -1).view(torch.uint8)
            out.append(packed)
        out.extend(columns[k] for k in self.big)
        return out

    def unpack(self, batch):
        """Typed views of the packed batch, one ``as_strided`` per field over one reinterpretation of the buffer per dtype
        (this runs once per batch on the consumer's thread; slicing + ``view`` + ``reshape`` per field was 72 us of a
        170 us batch of C5 windows)."""
        res = {}
        rest = batch
        if self.small:
            packed, rest = batch[0], batch[1:]
            n = packed.shape[0]
            typed = {}
            for k, off, nb, dtype, shape, item, strides in self._views():
                base = typed.get(dtype)
                if base is None:
                    base = typed[dtype] = packed.view(dtype)
                res[k] = torch.as_strided(base, (n,) + shape, (self.row_bytes // item,) + strides,
                                          base.storage_offset() + off // item)
        res.update(zip(self.big, rest))
        return {k: res[k] for k in self.keys}

    def _views(self):
        views = getattr(self, '_view_plan', None)
        if views is None:
            views = []
            for k, off, nb, dtype, shape in self.small:
                item = torch.empty(0, dtype=dtype).element_size()
                strides, acc = [], 1
                for d in reversed(shape):
                    strides.append(acc)
                    acc *= d
                views.append((k, off, nb, dtype, shape, item, tuple(reversed(strides))))
            self._view_plan = views
        return views


def _iter_device_batches(groups, batch_size, shuffling_queue_capacity, transform_fn):
    """Row-groups -> batches on the device: sanitise dtypes per column, pack the narrow columns, feed the batched
    shuffling buffer, emit ``{field: tensor[batch, ...]}`` (nested per offset for NGram windows)."""
    from petastorm_b200.gpu_workers import NGramColumns
    if shuffling_queue_capacity > 0:
        min_after = shuffling_queue_capacity - 1
        buf = BatchedRandomShufflingBuffer(min_after + batch_size, min_after_retrieve=min_after,
                                           extra_capacity=100000000, batch_size=batch_size)
    else:
        buf = BatchedNoopShufflingBuffer(batch_size=batch_size)
    packer, timesteps = None, None

    def drain():
        while buf.can_retrieve():
            batch = packer.unpack(buf.retrieve())
            if timesteps is not None:
                yield _nest_ngram_batch(list(batch.keys()), list(batch.values()), timesteps)
            else:
                yield batch

    for cols in groups:
        if isinstance(cols, NGramColumns):
            timesteps = cols.timesteps
        cols = {k: _as_tensor(v) for k, v in cols.items()}
        _sanitize_pytorch_types(cols)
        if packer is None:
            packer = _RowPacker(cols)
        buf.add_many(packer.pack(cols))
        for batch in drain():
            yield batch
    buf.finish()
    for batch in drain():
        yield batch


def _as_batch_tensor(value, transform_fn, batched):
    """Column value of a row / row-group -> tensor with a leading row dimension."""
    if isinstance(value, torch.Tensor):
        return value if batched else value.unsqueeze(0)
    return transform_fn(value) if batched else transform_fn([value])


class BatchedDataLoader(LoaderBase):
    """Batched loader (petastorm/pytorch.py:259-370): tensors in, ``{field: tensor[batch, ...]}`` out, shuffling through a
    batched buffer.  Over a petastorm_b200 reader whole decoded row-groups feed the device buffer directly."""

    def __init__(self, reader, batch_size=1, transform_fn=None, shuffling_queue_capacity=0):
        super(BatchedDataLoader, self).__init__()
        self.reader = reader
        self.batch_size = batch_size
        self.transform_fn = transform_fn or torch.as_tensor
        self._batch_acc = []
        self.shuffling_queue_capacity = shuffling_queue_capacity
        self._in_iter = None
        self.device_batched = None

    def _iter_impl(self):
        groups = _DeviceGroups(self.reader)
        if groups.available():
            self.device_batched = True
            return _iter_device_batches(groups, self.batch_size, self.shuffling_queue_capacity, self.transform_fn)
        self.device_batched = False
        return self._iter_rows()

    def _iter_rows(self):
        keys = None
        if self.shuffling_queue_capacity > 0:
            min_after = self.shuffling_queue_capacity - 1
            self._shuffling_buffer = BatchedRandomShufflingBuffer(min_after + self.batch_size,
                                                                  min_after_retrieve=min_after,
                                                                  extra_capacity=100000000,
                                                                  batch_size=self.batch_size)
        else:
            self._shuffling_buffer = BatchedNoopShufflingBuffer(batch_size=self.batch_size)
        for row in self.reader:
            row_as_dict = row._asdict()
            keys = row_as_dict.keys()
            _sanitize_pytorch_types(row_as_dict)
            for k, v in row_as_dict.items():
                row_as_dict[k] = _as_batch_tensor(v, self.transform_fn, self.reader.batched_output)
            self._shuffling_buffer.add_many(row_as_dict.values())
            for batch in self._yield_batches(keys):
                yield batch
        self._shuffling_buffer.finish()
        for batch in self._yield_batches(keys):
            yield batch

    def _yield_batches(self, keys):
        while self._shuffling_buffer.can_retrieve():
            batch = self._shuffling_buffer.retrieve()
            if not isinstance(batch, dict):
                batch = dict(zip(keys, batch))
            yield batch


def _load_rows_into_mem(reader, transform_fn, rows_capacity):
    """Up to ``rows_capacity`` rows into pre-allocated tensors, then stop the reader (petastorm/pytorch.py:373-434)."""
    n_rows = 0
    buffer = None
    keys = None
    for row in reader:
        row_as_dict = row._asdict()
        _sanitize_pytorch_types(row_as_dict)
        for k, v in row_as_dict.items():
            row_as_dict[k] = _as_batch_tensor(v, transform_fn, reader.batched_output)
        if not keys:
            keys = row_as_dict.keys()
        items = list(row_as_dict.values())
        take = min(len(items[0]), rows_capacity - n_rows)
        if buffer is None:
            buffer = [torch.empty((rows_capacity,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device) for v in items]
        for i, v in enumerate(items):
            buffer[i][n_rows:n_rows + take] = v[:take]
        n_rows += take
        if n_rows >= rows_capacity:
            break
    reader.stop()
    reader.join()
    if buffer is not None and n_rows < rows_capacity:
        buffer = [b[:n_rows] for b in buffer]
    return keys, buffer


class InMemBatchedDataLoader(object):
    """Loads up to ``rows_capacity`` rows once (into HBM with a petastorm_b200 reader) and serves ``num_epochs``
    epochs, reshuffled per epoch with ``torch.Generator().manual_seed(seed + epoch)`` like upstream (:469-493)."""

    def __init__(self, reader, batch_size=1, transform_fn=None, num_epochs=1, seed=0, rows_capacity=1024,
                 shuffle=False):
        self._batch_size = batch_size
        self._num_epochs = num_epochs
        self._seed = seed
        self._shuffle = shuffle
        self._in_iter = False
        self._keys, self._buffer = _load_rows_into_mem(reader, transform_fn or torch.as_tensor, rows_capacity)

    def __iter__(self):
        if self._in_iter:
            raise RuntimeError("InMemBatchedDataLoader couldn't be used multiple times, please\
                    specify total number of epochs using num_epochs in constructor.")
        self._in_iter = True
        from petastorm_b200.reader_impl.pytorch_shuffling_buffer import _take
        size = len(self._buffer[0])
        for epoch in range(self._num_epochs):
            if self._shuffle:
                g = torch.Generator()
                g.manual_seed(self._seed + epoch)
                indices = torch.randperm(size, generator=g)
            else:
                indices = torch.arange(size)
            for i in range(0, size, self._batch_size):
                idx = indices[i:i + self._batch_size].to(self._buffer[0].device)
                yield dict(zip(self._keys, [_take(v, idx) for v in self._buffer]))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        pass
