"""Row-group workers of the B200 path.

:class:`GpuArrowWorker` is the counterpart of ``ArrowReaderWorker`` (petastorm/arrow_reader_worker.py:117-393, batch
reader: one result per row-group, columns as arrays) and :class:`GpuPyDictWorker` of ``PyDictReaderWorker``
(petastorm/py_dict_reader_worker.py:100-286, row reader: codecs decoded, optional NGram).  Both keep the reference's
plug-in seam: ``Worker(worker_id, publish_func, args)``, ``process(piece_index, worker_predicate,
shuffle_row_drop_partition)``, ``staticmethod new_results_queue_reader()``.

What differs is *where* the work happens: ``piece.read`` becomes plan + H2D + CUDA decode
(:mod:`petastorm_b200.rowgroup`), ``table.take`` / ``DataFrame.sample`` become a device gather, codecs become batched
kernels, predicates become device masks + stream compaction, and results stay in HBM as torch tensors (views of the
row-group's output buffer).  Values that cannot be tensors (strings, Decimals, dates) are materialised on the host
after the device did the decompression / level / dictionary work - the reference DataLoader refuses them anyway.
"""
import hashlib
from decimal import Decimal

import numpy as np
import torch

from petastorm_b200 import device_ops, native, rowgroup
from petastorm_b200.cache import NullCache
from petastorm_b200.codecs import (CompressedImageCodec, CompressedNdarrayCodec, NdarrayCodec, ScalarCodec,
                                   parse_npy_header)
from petastorm_b200.errors import DecodeFieldError
from petastorm_b200.rowgroup import BOOLEAN, BYTE_ARRAY, FIXED_LEN_BYTE_ARRAY, INT32, INT64, INT96
from petastorm_b200.unischema import integer_logical_type
from petastorm_b200.workers_pool import EmptyResultError
from petastorm_b200.workers_pool.worker_base import WorkerBase

_TORCH_OF_NUMPY = {np.dtype('uint8'): torch.uint8, np.dtype('int8'): torch.int8, np.dtype('int16'): torch.int16,
                   np.dtype('uint16'): torch.uint16, np.dtype('int32'): torch.int32, np.dtype('uint32'): torch.uint32,
                   np.dtype('int64'): torch.int64, np.dtype('uint64'): torch.uint64, np.dtype('float16'): torch.float16,
                   np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64, np.dtype('bool'): torch.bool}


#: test hook: force the host-staged JPEG path even when nvJPEG can read device bitstreams
_FORCE_HOST_JPEG = [False]


class ScalarColumn(object):
    """A numeric scalar column of the row reader: the device tensor (for the batched loaders) plus a lazily fetched
    host copy that serves the per-row numpy scalars the reference hands out (``field.numpy_dtype(value)``)."""

    __slots__ = ('tensor', 'np_type', '_host')

    def __init__(self, tensor, np_type, host=None):
        self.tensor = tensor
        self.np_type = np_type
        self._host = host

    def _values(self):
        if self._host is None:
            self._host = self.tensor.cpu().numpy()
        return self._host

    def __len__(self):
        return int(self.tensor.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return ScalarColumn(self.tensor[i], self.np_type, None if self._host is None else self._host[i])
        v = self._values()[i]
        return self.np_type(v) if self.np_type is not None else v

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def take(self, dev_index):
        return ScalarColumn(device_ops.gather_rows(self.tensor.contiguous(), dev_index), self.np_type)


def _npify(value):
    if isinstance(value, ScalarColumn):
        return value._values()  # pylint: disable=protected-access
    return value.cpu().numpy() if isinstance(value, torch.Tensor) else value


class WorkerOptions(object):
    """Extra, GPU-specific worker arguments appended after the reference's 12-tuple."""

    def __init__(self, partitions=None, output='torch', device=None):
        self.partitions = partitions      # etl.dataset_metadata.PartitionSet
        self.output = output              # 'torch' (device tensors) or 'numpy' (host arrays, drop-in types)
        self.device = device


class _RawRowGroup(object):
    """Decoded (still raw) columns of one row-group for a set of schema field names."""

    def __init__(self, piece, decoded, file_schema, name_to_slot, partition_values, num_rows):
        self.piece = piece
        self.decoded = decoded
        self.file_schema = file_schema
        self.name_to_slot = name_to_slot
        self.partition_values = partition_values
        self.num_rows = num_rows


class _GpuWorkerBase(WorkerBase):
    def __init__(self, worker_id, publish_func, args):
        super(_GpuWorkerBase, self).__init__(worker_id, publish_func, args)
        self._dataset_path = args[1]
        self._schema = args[2]
        self._ngram = args[3]
        self._split_pieces = args[4]
        self._local_cache = args[5]
        self._transform_spec = args[6]
        self._transformed_schema = args[7]
        self._shuffle_rows = args[9]
        self._random_seed = args[10]
        self._options = args[12] if len(args) > 12 and args[12] is not None else WorkerOptions()
        self._rng = np.random.default_rng(self._random_seed)
        self._decoder = None
        self._post_streams = {}
        self.rows_decoded = 0
        self.payload_bytes = 0
        self.t_issue = self.t_wait = self.t_build = 0.0   # host seconds: issuing device work / waiting / building columns

    # ---- device decode of the requested fields ------------------------------------------------------------------
    def _get_decoder(self):
        if self._decoder is None:
            self._decoder = rowgroup.RowGroupDecoder(self._options.device)
        return self._decoder

    def _field_slots(self, piece, field_names):
        """(open file, {field name: plan slot}, [leaf column ids]) of the leaf columns behind `field_names`
        (partition columns are served from the piece's keys, not from the file)."""
        pfile = rowgroup.open_file(piece.path)
        partition_names = self._options.partitions.partition_names if self._options.partitions else set()
        leaves = pfile.schema['leaves']
        wanted = [n for n in field_names if n not in partition_names]
        name_to_slot, leaf_ids = {}, []
        by_name = getattr(pfile, '_leaves_by_name', None)
        if by_name is None:
            by_name = {}
            for l in leaves:
                by_name.setdefault(l['path'][0], []).append(l['index'])
            pfile._leaves_by_name = by_name  # pylint: disable=protected-access
        for name in wanted:
            ids = by_name.get(name, [])
            if not ids:
                raise ValueError('Field {} was not found in the file {}'.format(name, piece.path))
            if len(ids) > 1:
                raise ValueError('Field {} maps to a nested parquet structure that is not supported'.format(name))
            name_to_slot[name] = len(leaf_ids)
            leaf_ids.append(ids[0])
        return pfile, name_to_slot, leaf_ids

    def _read_raw(self, piece, field_names):
        """plan + H2D + device decode of the leaf columns behind `field_names` (partition columns excluded)."""
        import time
        t0 = time.perf_counter()
        dec = self._get_decoder()
        pfile, name_to_slot, leaf_ids = self._field_slots(piece, field_names)
        num_rows = pfile.row_group_num_rows(piece.row_group)
        decoded = None
        if leaf_ids:
            decoded = dec.decode(piece.path, piece.row_group, leaf_ids)
            self.payload_bytes += decoded.plan.info.payload_bytes
        pvals = {k: v for k, v in piece.partition_keys if k in field_names}
        self.t_issue += time.perf_counter() - t0
        return _RawRowGroup(piece, decoded, pfile.schema, name_to_slot, pvals, num_rows)

    def _stream_of(self, raw):
        """The stream post-processing of a decoded row-group runs on.  Not the decode stream: by the time a row-group
        is finalised, the issuing thread has already queued later row-groups (H2D + ~15 ms of decode) on that stream
        and anything appended there - in particular the `done` event the consumer waits for - would sit behind them.
        ``decoded.wait()`` (``_build_columns``) orders this stream after the row-group's own decode."""
        return self._thread_post_stream()

    def _thread_post_stream(self):
        """One post-processing stream per host thread that resolves row-groups (consumer, resolver threads): the
        blocking host reads of one resolution then never wait for another one's kernels."""
        import threading
        streams = self._post_streams
        ident = threading.get_ident()
        stream = streams.get(ident)
        if stream is None:
            device = self._get_decoder().device
            torch.cuda.set_device(device)          # resolver threads start on device 0
            stream = streams[ident] = torch.cuda.Stream(device)
        return stream

    # ---- row selection ------------------------------------------------------------------------------------------
    def _row_order(self, num_rows, shuffle_row_drop_partition, ngram_length=0):
        """Host int64 array of source rows in output order, or None for "all rows, natural order".

        Combines the within-row-group shuffle (petastorm/arrow_reader_worker.py:361-371: seeded ``Generator`` kept per
        worker, or global ``np.random`` when the seed is None/0) with the drop-partition slice
        ``floor(arange(n) / (n / min(n, P))) == this`` (``:386-391``; the row reader additionally borrows
        ``length-1`` rows of the next partition for NGrams, petastorm/py_dict_reader_worker.py:276-285)."""
        order = None
        if self._shuffle_rows and num_rows:
            order = self._shuffle_order(num_rows)
        this_partition, num_partitions = shuffle_row_drop_partition
        if num_partitions > 1 and num_rows:
            pidx = np.floor(np.arange(num_rows) / (float(num_rows) / min(num_rows, num_partitions)))
            if ngram_length > 1:
                nxt = np.where(pidx >= this_partition + 1)[0]
                if nxt.size:
                    pidx[nxt[0:ngram_length - 1]] = this_partition
            keep = np.nonzero(pidx == this_partition)[0]
            order = keep if order is None else order[keep]
        return order

    def _shuffle_order(self, num_rows):
        if self._random_seed is not None and self._random_seed != 0:
            return self._rng.permutation(num_rows)
        return np.random.permutation(num_rows)

    def _cache_key(self, piece, piece_index):
        path = self._dataset_path
        path_str = ','.join(path) if isinstance(path, list) else path
        return '{}:{}:{}'.format(hashlib.md5(path_str.encode('utf-8')).hexdigest(), piece.path, piece_index)

    def _check_cache_usage(self, worker_predicate, shuffle_row_drop_partition):
        if not isinstance(self._local_cache, NullCache):
            if worker_predicate:
                raise RuntimeError('Local cache is not supported together with predicates, '
                                   'unless the dataset is partitioned by the column the predicate operates on.')
            if shuffle_row_drop_partition[1] != 1:
                raise RuntimeError('Local cache is not supported together with shuffle_row_drop_partitions > 1')

    @staticmethod
    def _validate_predicate_fields(worker_predicate, schema):
        predicate_fields = set(worker_predicate.get_fields())
        if not predicate_fields:
            raise ValueError('At least one field name must be returned by predicate\'s get_field() method')
        all_names = set(schema.fields.keys())
        invalid = predicate_fields - all_names
        if invalid:
            raise ValueError('At least some column names requested by the predicate ({}) '
                             'are not valid schema names: ({})'.format(', '.join(invalid), ', '.join(all_names)))
        return predicate_fields, all_names

    @property
    def diagnostics(self):
        d = {'rows_decoded': self.rows_decoded, 'payload_bytes': self.payload_bytes,
             'host_seconds': {'issue': round(self.t_issue, 4), 'wait': round(self.t_wait, 4),
                              'build': round(self.t_build, 4),
                              'plan_upload_decode': [round(x, 4) for x in (self._decoder.host_seconds
                                                                           if self._decoder else (0, 0, 0))]}}
        if self._decoder is not None:
            d['gpu_launches'] = self._decoder.launches
            d['h2d_bytes'] = self._decoder.h2d_bytes
            d['hbm_cache_hits'] = self._decoder.hbm_cache_hits
            d.update(self._decoder.ctx.stats())
        return d

    # ---- column materialisation helpers (shared) ----------------------------------------------------------------
    def _partition_column(self, raw, name, count):
        dtype = self._options.partitions.dtype_of(name)
        value = raw.partition_values[name]
        if dtype is np.int64:
            return np.full(count, int(value), dtype=np.int64)
        return np.full(count, str(value))      # numpy infers the string width (dtype=np.str_ alone would mean '<U1')

    @staticmethod
    def _leaf_of(raw, name):
        slot = raw.name_to_slot[name]
        col = raw.decoded.column(slot)
        return slot, col, raw.file_schema['leaves'][col.leaf]

    @staticmethod
    def _numeric_tensor(col, leaf):
        """Typed device tensor of a fixed-width column, nulls still zero-filled."""
        pt = col.physical_type
        v = col.values
        if pt == BOOLEAN:
            return v.view(torch.bool)
        if pt in (INT32, INT64):
            bits, signed = integer_logical_type(leaf)
            if pt == INT32 and bits in (8, 16):
                return device_ops.narrow_int32(v, {(8, True): torch.int8, (8, False): torch.uint8,
                                                   (16, True): torch.int16, (16, False): torch.uint16}[(bits, signed)])
            if not signed:
                return v.view(torch.uint32 if pt == INT32 else torch.uint64)
            return v
        return v

    @staticmethod
    def _decimal_from_bytes(raw_bytes, scale):
        return Decimal(int.from_bytes(raw_bytes, 'big', signed=True)).scaleb(-scale)

    def _host_objects(self, col, leaf, order):
        """Host materialisation of the values that have no tensor form: list of str / bytes / Decimal / None."""
        pt = col.physical_type
        is_decimal = leaf['converted_type'] == 5 or leaf['logical_kind'] == 5
        if pt == BYTE_ARRAY:
            blobs = rowgroup.gather_blobs_to_host(col, order)
            if is_decimal:
                return [None if b is None else self._decimal_from_bytes(b, leaf['scale']) for b in blobs]
            if leaf['converted_type'] == 0 or leaf['logical_kind'] == 1:
                return [None if b is None else b.decode('utf-8') for b in blobs]
            return blobs
        vals = col.values.cpu().numpy()
        valid = col.valid.cpu().numpy().astype(bool) if col.valid is not None else np.ones(len(vals), dtype=bool)
        idx = np.arange(len(vals)) if order is None else np.asarray(order)
        vals, valid = vals[idx], valid[idx]
        if pt == FIXED_LEN_BYTE_ARRAY and not is_decimal:
            out = np.empty(len(vals), dtype=object)
            out[:] = [v.tobytes() for v in vals]
            out[~valid] = None
            return list(out)
        # Decimals are Python objects (one per cell, like upstream); the unscaled integers are assembled with numpy
        if pt == FIXED_LEN_BYTE_ARRAY:
            width = vals.shape[1] if vals.ndim == 2 else 0
            if 0 < width <= 8:
                be = np.zeros((len(vals), 8), dtype=np.uint8)
                be[:, 8 - width:] = vals
                be[:, :8 - width] = np.where(vals[:, :1] & 0x80, 0xff, 0)       # sign extension
                ints = be.view('>i8').ravel().tolist()
            else:
                ints = [int.from_bytes(v.tobytes(), 'big', signed=True) for v in vals]
        else:               # INT32 / INT64 decimals
            ints = vals.tolist()
        scale = leaf['scale']
        return [Decimal(i).scaleb(-scale) if ok else None for i, ok in zip(ints, valid.tolist())]

    @staticmethod
    def _datetime_array(col, leaf, order, null_count):
        """Host datetime64 array (timestamps keep their unit like pandas>=2, dates become datetime64[D])."""
        vals = col.values.cpu().numpy()
        if col.physical_type == INT96:
            # legacy impala timestamps: 8 bytes nanos-of-day + 4 bytes julian day
            nanos = vals[:, :8].copy().view('<i8').ravel()
            days = vals[:, 8:].copy().view('<i4').ravel().astype(np.int64)
            out = ((days - 2440588) * 86400 * 10 ** 9 + nanos).astype('datetime64[ns]')
        elif leaf['converted_type'] == 6 or leaf['logical_kind'] == 6:
            out = vals.astype('datetime64[D]')
        else:
            unit = {1: 'ms', 2: 'us', 3: 'ns'}.get(leaf['logical_unit'])
            if unit is None:
                unit = 'ms' if leaf['converted_type'] == 9 else 'us'
            out = vals.astype('datetime64[{}]'.format(unit))
        if null_count:
            out = out.copy()
            out[~col.valid.cpu().numpy().astype(bool)] = np.datetime64('NaT')
        return out if order is None else out[order]


def _is_temporal(leaf):
    return (leaf['physical_type'] == INT96 or leaf['converted_type'] in (6, 9, 10) or leaf['logical_kind'] in (6, 8))


def _is_host_only(leaf):
    pt = leaf['physical_type']
    is_decimal = leaf['converted_type'] == 5 or leaf['logical_kind'] == 5
    return pt in (BYTE_ARRAY, FIXED_LEN_BYTE_ARRAY) or is_decimal


# =====================================================================================================================
# batch reader worker
# =====================================================================================================================
class GpuBatch(object):
    """One decoded row-group of the batch reader: ``columns`` maps field name -> CUDA tensor (or host numpy array for
    strings / decimals / datetimes).  ``wait()`` orders the consumer stream after the decode."""

    def __init__(self, columns, num_rows, keepalive):
        self.columns = columns
        self.num_rows = num_rows
        self._keepalive = keepalive

    def wait(self):
        for k in self._keepalive:
            if k is not None:
                k.wait()
        _record_on_current_stream(self.columns)


class GpuArrowResultsQueueReader(object):
    """``read_next`` of the batch reader (petastorm/arrow_reader_worker.py:89-114): one namedtuple per row-group."""

    def __init__(self, output='torch'):
        self._output = output
        self._peeked = None

    @property
    def batched_output(self):
        return True

    def _next_batch(self, workers_pool):
        if self._peeked is not None:
            batch, self._peeked = self._peeked, None
            return batch
        batch = workers_pool.get_results()
        if isinstance(batch, PendingRowGroup):
            batch = batch.resolve()
        batch.wait()
        return batch

    def peek_rowgroup(self, workers_pool):
        """Columns of the next row-group without consuming it (the loaders look at the value types to pick the
        device-batched path); None at the end of the data."""
        try:
            if self._peeked is None:
                self._peeked = self._next_batch(workers_pool)
            return self._peeked.columns
        except EmptyResultError:
            return None

    def read_next_rowgroup(self, workers_pool, raw=False):
        """Whole next row-group as ``{field: column}`` - entry point of the device-batched loaders."""
        try:
            return dict(self._next_batch(workers_pool).columns)
        except EmptyResultError:
            raise StopIteration

    def read_next(self, workers_pool, schema, ngram):
        try:
            assert not ngram, 'ArrowReader does not support ngrams for now'
            batch = self._next_batch(workers_pool)
            cols = batch.columns
            if self._output == 'numpy':
                cols = {k: _npify(v) for k, v in cols.items()}
            return schema.make_namedtuple(**cols)
        except EmptyResultError:
            raise StopIteration


class GpuArrowWorker(_GpuWorkerBase):
    def __init__(self, worker_id, publish_func, args):
        super(GpuArrowWorker, self).__init__(worker_id, publish_func, args)
        if self._ngram:
            raise NotImplementedError('ngrams are not supported by ArrowReaderWorker')

    @staticmethod
    def new_results_queue_reader():
        return GpuArrowResultsQueueReader()

    def process(self, piece_index, worker_predicate, shuffle_row_drop_partition):
        piece = self._split_pieces[piece_index]
        self._check_cache_usage(worker_predicate, shuffle_row_drop_partition)
        if worker_predicate:
            batch = self._load_rows_with_predicate(piece, worker_predicate, shuffle_row_drop_partition)
        elif isinstance(self._local_cache, NullCache):
            # asynchronous: issue the device work now, resolve on the consumer side
            pending = self._issue_rows(piece, shuffle_row_drop_partition)
            if pending.num_rows:
                self.publish_func(pending)
            return
        else:
            batch = self._local_cache.get(self._cache_key(piece, piece_index),
                                          lambda: self._load_rows(piece, shuffle_row_drop_partition))
        if batch is not None and batch.num_rows:
            self.rows_decoded += batch.num_rows
            self.publish_func(batch)

    # ---- column -> batch value (semantics of convert_arrow_table_to_numpy_dict, arrow_reader_worker.py:31-86) -----
    def _materialize(self, raw, name, field, order):
        count = raw.num_rows if order is None else len(order)
        if name in raw.partition_values:
            return self._partition_column(raw, name, count)
        slot, col, leaf = self._leaf_of(raw, name)
        nulls = raw.decoded.null_counts[slot]
        dev_order = None if order is None else torch.from_numpy(np.ascontiguousarray(order)).to(col.arena.device)
        if col.max_rep > 0:
            return self._list_column(raw, name, field, col, leaf, nulls, dev_order)
        if _is_temporal(leaf):
            return self._datetime_array(col, leaf, order, nulls)
        if _is_host_only(leaf):
            objs = self._host_objects(col, leaf, order)
            arr = np.empty(len(objs), dtype=object)
            arr[:] = objs
            if leaf['physical_type'] == BYTE_ARRAY and (leaf['converted_type'] == 0 or leaf['logical_kind'] == 1):
                return arr.astype(np.str_)  # arrow_reader_worker.py:66-67 (None becomes 'None', as upstream)
            return arr
        t = self._numeric_tensor(col, leaf)
        if nulls:
            pt = col.physical_type
            if pt == BOOLEAN:
                # pandas yields an object column of True/False/None (no tensor form); built with numpy, not per row
                vals = t.cpu().numpy().astype(bool)
                valid = col.valid.cpu().numpy().astype(bool)
                arr = vals.astype(object)
                arr[~valid] = None
                return arr if order is None else arr[order]
            bits, signed = integer_logical_type(leaf) if pt in (INT32, INT64) else (0, True)
            t = device_ops.nulls_to_nan(col.values, col.valid, pt, bits, not signed)
        if dev_order is not None:
            t = device_ops.gather_rows(t.contiguous(), dev_order)
        return t

    def _list_column(self, raw, name, field, col, leaf, nulls, dev_order):
        n_rows = raw.num_rows
        n = col.num_values
        if nulls or n_rows == 0 or n % n_rows != 0 or \
                not device_ops.list_is_uniform(col.rep, col.defs, col.max_def, n // n_rows):
            raise RuntimeError('Length of all values in column \'{}\' are expected to be the same length.'.format(name))
        t = self._numeric_tensor(col, leaf).view(n_rows, n // n_rows)
        shape = self._schema_for_shapes().fields[name].shape if name in self._schema_for_shapes().fields else ()
        if len(shape) > 1:
            t = t.reshape((n_rows,) + tuple(shape))
        if dev_order is not None:
            t = device_ops.gather_rows(t.contiguous(), dev_order)
        return t

    def _schema_for_shapes(self):
        return self._transformed_schema if self._transform_spec else self._schema

    def _build_columns(self, raw, names, order):
        import time
        t0 = time.perf_counter()
        if raw.decoded is not None:
            raw.decoded.check()   # host sync: surfaces corrupt pages and delivers the per-column null counts
            raw.decoded.wait()
        t1 = time.perf_counter()
        out = {}
        for name in names:
            out[name] = self._materialize(raw, name, self._schema.fields[name], order)
        self.t_wait += t1 - t0
        self.t_build += time.perf_counter() - t1
        return out

    def _issue_rows(self, piece, shuffle_row_drop_partition):
        names = [f.name for f in self._schema.fields.values()]
        raw = self._read_raw(piece, names)                                  # plan + H2D + decode kernels, all async
        order = self._row_order(raw.num_rows, shuffle_row_drop_partition)   # RNG drawn here: order of issue is fixed
        count = raw.num_rows if order is None else len(order)

        def finalize():
            with torch.cuda.stream(self._stream_of(raw)):
                cols = self._build_columns(raw, names, order)
                if self._transform_spec:
                    cols = self._apply_transform(cols, count)
                done = torch.cuda.Event()
                done.record()
            self.rows_decoded += count
            return GpuBatch(cols, count, [_EventWaiter(done), raw.decoded])

        return PendingRowGroup(finalize, count)

    def _load_rows(self, piece, shuffle_row_drop_partition):
        return self._issue_rows(piece, shuffle_row_drop_partition).resolve()

    def _apply_transform(self, cols, count):
        spec = self._transform_spec
        if spec.func:
            if spec.device:
                cols = spec.func(cols)
            else:
                cols = _host_dataframe_transform(spec.func, cols, self._get_decoder().device)
        for name in set(cols.keys()) & set(spec.removed_fields):
            del cols[name]
        got, want = set(cols.keys()), set(f.name for f in self._transformed_schema.fields.values())
        if got != want:
            raise ValueError('Transformed result columns ({rc}) do not match required schema columns({sc})'
                             .format(rc=','.join(got), sc=','.join(want)))
        return {name: cols[name] for name in self._transformed_schema.fields.keys()}

    def _load_rows_with_predicate(self, piece, worker_predicate, shuffle_row_drop_partition):
        predicate_fields, all_names = self._validate_predicate_fields(worker_predicate, self._schema)
        other_names = all_names - predicate_fields
        decoder = self._get_decoder()
        # 1. predicate columns first
        raw_p = self._read_raw(piece, predicate_fields)
        order = self._row_order(raw_p.num_rows, shuffle_row_drop_partition)
        with torch.cuda.stream(self._stream_of(raw_p)):
            pcols = self._build_columns(raw_p, sorted(predicate_fields), order)
            mask = worker_predicate.device_mask(pcols)
            if mask is None:
                mask = _host_vector_predicate(worker_predicate, pcols, decoder.device)
            keep = device_ops.mask_to_indices(mask.contiguous())  # syncs: the count decides the early exit
            if keep.numel() == 0:
                return None
            keep_host = keep.cpu().numpy()
            sel = keep_host if order is None else order[keep_host]
            # 2. the other columns, only for matching rows
            cols = {name: _take(pcols[name], keep, keep_host) for name in pcols}
            raw_o = None
            if other_names:
                raw_o = self._read_raw(piece, other_names)
                cols.update(self._build_columns(raw_o, sorted(other_names), sel))
            cols = {name: cols[name] for name in self._schema.fields.keys()}
            if int(keep.numel()) < (raw_p.num_rows if order is None else len(order)):
                # pandas quirk kept for parity: upstream blanks the rejected rows with None before filtering
                # (arrow_reader_worker.py:324,331), which upcasts every integer column to float64
                for name, v in cols.items():
                    if isinstance(v, torch.Tensor) and not v.is_floating_point() and v.dtype != torch.bool:
                        cols[name] = v.to(torch.float64)
            if self._transform_spec:
                # upstream applies func without the removed-fields post-processing here (arrow_reader_worker.py:342-345)
                spec = self._transform_spec
                if spec.device:
                    cols = spec.func(cols)
                else:
                    cols = _host_dataframe_transform(spec.func, cols, decoder.device)
            done = torch.cuda.Event()
            done.record()
        return GpuBatch(cols, int(keep.numel()), [_EventWaiter(done), raw_p.decoded, raw_o.decoded if raw_o else None])


class PendingRowGroup(object):
    """A row-group whose H2D copy and decode kernels have been *issued* by the pool thread but whose host-visible part
    (error word, null counts, column views, codec / post-processing launches) is resolved by the consumer.  Publishing
    these handles instead of finished batches keeps several row-groups in flight: the pool thread never blocks on the
    device, so PCIe transfer, decode and the consumer overlap (the results queue bounds the depth)."""

    def __init__(self, finalize, num_rows):
        self._finalize = finalize
        self._result = None
        self._done = False
        self._future = None
        self.num_rows = num_rows

    def resolve_ahead(self, executor):
        """Start the resolution on a resolver thread of the pool (codec launches and their host-side waits then
        overlap the consumer's work on the previous row-group); ``resolve`` waits for it."""
        if self._future is None and not self._done:
            self._future = executor.submit(self._run)

    def _run(self):
        if not self._done:
            self._result = self._finalize()
            self._finalize = None
            self._done = True
        return self._result

    def resolve(self):
        if self._future is not None:
            return self._future.result()
        return self._run()


def _record_on_current_stream(columns):
    """Columns were allocated on side streams; tell the caching allocator that the consumer's stream uses them."""
    stream = torch.cuda.current_stream()
    for v in columns.values():
        t = v.tensor if isinstance(v, ScalarColumn) else v
        if isinstance(t, torch.Tensor) and t.is_cuda:
            t.record_stream(stream)


class _EventWaiter(object):
    def __init__(self, event):
        self._event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self._event)


def _take(value, dev_index, host_index):
    if isinstance(value, torch.Tensor):
        return device_ops.gather_rows(value.contiguous(), dev_index)
    return value[host_index]


def _host_vector_predicate(predicate, cols, device):
    """User predicate without a device form on the batch reader: upstream passes a pandas DataFrame and expects a
    boolean Series (petastorm/arrow_reader_worker.py:315-318)."""
    import pandas as pd
    frame = pd.DataFrame({k: list(_npify(v)) if _npify(v).ndim > 1 else _npify(v) for k, v in cols.items()})
    res = predicate.do_include(frame)
    mask = np.asarray(res, dtype=bool)
    if mask.ndim == 0:
        mask = np.full(len(frame), bool(mask))
    return torch.from_numpy(mask.astype(np.uint8)).to(device)


def _host_dataframe_transform(func, cols, device):
    """Opaque user ``TransformSpec.func`` on the batch reader: it receives a pandas DataFrame like upstream
    (petastorm/arrow_reader_worker.py:247-277) and the returned columns go back to the device."""
    import pandas as pd
    data = {}
    for k, v in cols.items():
        a = _npify(v)
        data[k] = list(a) if a.ndim > 1 else a
    out = func(pd.DataFrame(data))
    res = {}
    for k in out.columns:
        s = out[k]
        a = s.values
        if a.dtype == object and len(a) and isinstance(a[0], np.ndarray):
            a = np.stack(list(a))
        if isinstance(a, np.ndarray) and a.dtype in _TORCH_OF_NUMPY:
            res[k] = torch.from_numpy(np.ascontiguousarray(a)).to(device)
        else:
            res[k] = np.asarray(a)
    return res


# =====================================================================================================================
# row reader worker
# =====================================================================================================================
class GpuRowGroupRows(object):
    """Decoded row-group of the row reader.  ``columns[name]`` is a per-row indexable: a CUDA tensor ``[n, ...]``,
    a host numpy array, or a python list (strings, Decimals, ragged arrays, ``None`` for nulls)."""

    def __init__(self, columns, num_rows, keepalive):
        self.columns = columns
        self.num_rows = num_rows
        self._keepalive = keepalive

    def wait(self):
        for k in self._keepalive:
            if k is not None:
                k.wait()
        _record_on_current_stream(self.columns)

    def row(self, i, output):
        out = {}
        for name, col in self.columns.items():
            v = col[i]
            if output == 'numpy' and isinstance(v, torch.Tensor):
                v = v.cpu().numpy()
            out[name] = v
        return out


class NGramColumns(dict):
    """Windows of one row-group in device form: ``self[name]`` is a tensor ``[W, L, ...]`` (window, timestep, value
    shape) for every field of the NGram; ``timesteps`` maps each offset of the NGram to the field names it carries
    (petastorm/ngram.py:259-264 projects each timestep to its own field list)."""

    def __init__(self, columns, timesteps):
        super(NGramColumns, self).__init__(columns)
        self.timesteps = timesteps


class GpuNGramWindows(object):
    """NGram result of one row-group: window starts + the per-row columns they index into."""

    def __init__(self, rows, starts, ngram, starts_dev=None):
        self.rows = rows
        self.starts = starts          # python list of start rows
        self.starts_dev = starts_dev  # the same as a CUDA int64 tensor (None when the windows were formed on the host)
        self.ngram = ngram
        self.num_rows = len(starts)

    def wait(self):
        self.rows.wait()

    def window_columns(self, first=0):
        """:class:`NGramColumns` of the windows ``first:`` - one window-gather kernel (``pst_ngram_gather``) per field,
        or None when a field has no tensor form (strings, nulls, ragged arrays)."""
        ng = self.ngram
        base = ng.base_key
        timesteps = {base + k: ng.get_field_names_at_timestep(base + k) for k in range(ng.length)}
        names = [n for n in self.rows.columns if any(n in f for f in timesteps.values())]
        cols = {}
        for name in names:
            v = self.rows.columns[name]
            t = v.tensor if isinstance(v, ScalarColumn) else v
            if not isinstance(t, torch.Tensor):
                return None
            cols[name] = t
        starts = self.starts_dev
        if starts is None:
            device = next(iter(cols.values())).device if cols else None
            starts = torch.tensor(self.starts, dtype=torch.int64, device=device)
        else:
            starts.record_stream(torch.cuda.current_stream(starts.device))   # allocated on the worker's side stream
        if first:
            starts = starts[first:]
        return NGramColumns({name: device_ops.ngram_gather(t.contiguous(), starts, ng.length)
                             for name, t in cols.items()}, timesteps)


class GpuPyDictResultsQueueReader(object):
    """``read_next`` of the row reader (petastorm/py_dict_reader_worker.py:64-97): one namedtuple per row, or one
    ``{offset: namedtuple}`` per NGram window; a row-group is buffered and handed out row by row."""

    def __init__(self, output='torch'):
        import threading
        self._lock = threading.Lock()
        self._current = None
        self._next_index = 0
        self._output = output

    @property
    def batched_output(self):
        return False

    @staticmethod
    def _next_group(workers_pool):
        """Next non-empty decoded row-group (resolving handles of work issued ahead by the pool thread)."""
        while True:
            group = workers_pool.get_results()
            if isinstance(group, PendingRowGroup):
                group = group.resolve()
                if group is None:
                    continue
            group.wait()
            return group

    def peek_rowgroup(self, workers_pool):
        """Columns of the (rest of the) current row-group without consuming it: ``{field: column}`` for plain rows,
        the :class:`GpuNGramWindows` for an NGram reader; None at the end of the data."""
        try:
            with self._lock:
                while self._current is None or self._next_index >= self._current.num_rows:
                    self._current = self._next_group(workers_pool)
                    self._next_index = 0
                return self._current if isinstance(self._current, GpuNGramWindows) else self._current.columns
        except EmptyResultError:
            return None

    def read_next_rowgroup(self, workers_pool, raw=False):
        """Whole (rest of the) current row-group as ``{field: column}`` with device tensors where they exist - the
        entry point of the device-batched loaders (no per-row namedtuples).  An NGram reader hands out
        :class:`NGramColumns` (``[W, L, ...]`` per field), or the :class:`GpuNGramWindows` itself when a field has no
        tensor form.  ``raw=True`` keeps :class:`ScalarColumn` wrappers (per-row numpy scalars on demand)."""
        try:
            with self._lock:
                if self._current is not None and self._next_index < self._current.num_rows:
                    cur, start = self._current, self._next_index
                else:
                    cur, start = self._next_group(workers_pool), 0
                self._current = None
            if isinstance(cur, GpuNGramWindows):
                cols = cur.window_columns(start)
                if cols is None:
                    # hand the group back: the caller walks it window by window through read_next
                    with self._lock:
                        self._current, self._next_index = cur, start
                    return cur
                return cols
            cols = {k: (v[start:] if start else v) for k, v in cur.columns.items()}
            if raw:
                return cols
            return {k: (v.tensor if isinstance(v, ScalarColumn) else v) for k, v in cols.items()}
        except EmptyResultError:
            raise StopIteration

    def read_next(self, workers_pool, schema, ngram):
        try:
            with self._lock:
                while self._current is None or self._next_index >= self._current.num_rows:
                    self._current = self._next_group(workers_pool)
                    self._next_index = 0
                i = self._next_index
                self._next_index += 1
                cur = self._current
            if ngram:
                start = cur.starts[i]
                base = ngram.base_key
                item = {}
                for k in range(ngram.length):
                    names = ngram.get_field_names_at_timestep(base + k)
                    row = cur.rows.row(start + k, self._output)
                    view = ngram.get_schema_at_timestep(schema, base + k)
                    item[base + k] = view.make_namedtuple(**{n: row[n] for n in row if n in names})
                return item
            return schema.make_namedtuple(**cur.row(i, self._output))
        except EmptyResultError:
            raise StopIteration


class GpuPyDictWorker(_GpuWorkerBase):
    @staticmethod
    def new_results_queue_reader():
        return GpuPyDictResultsQueueReader()

    def process(self, piece_index, worker_predicate, shuffle_row_drop_partition):
        piece = self._split_pieces[piece_index]
        self._check_cache_usage(worker_predicate, shuffle_row_drop_partition)
        if worker_predicate:
            # the predicate columns are decoded and evaluated here (the count decides whether the payload is read at
            # all: upstream's early exit, py_dict_reader_worker.py:234-236); the payload columns of the matching rows
            # are issued asynchronously and resolved by the consumer, like a row-group without a predicate
            pending = self._load_rows_with_predicate(piece, worker_predicate, shuffle_row_drop_partition, defer=True)
            if pending is not None:
                self.publish_func(pending)
            return
        elif isinstance(self._local_cache, NullCache):
            pending = self._issue_rows(piece, shuffle_row_drop_partition)   # asynchronous, resolved by the consumer
            if pending.num_rows:
                self.publish_func(pending)
            return
        else:
            rows = self._local_cache.get(self._cache_key(piece, piece_index),
                                         lambda: self._load_rows(piece, shuffle_row_drop_partition))
        result = self._finish_rows(rows)
        if result is not None:
            self.publish_func(result)

    def _finish_rows(self, rows):
        """NGram formation + accounting of a decoded row-group; None when nothing is left to publish."""
        if rows is None or rows.num_rows == 0:
            return None
        result = self._form_ngram(rows) if self._ngram else rows
        if result.num_rows == 0:
            return None
        self.rows_decoded += result.num_rows
        return result

    # ---- shuffle: DataFrame.sample(frac=1, random_state=seed) (py_dict_reader_worker.py:269-270) ----------------
    def _shuffle_order(self, num_rows):
        # pandas' sample(frac=1, random_state=s) draws RandomState(s).permutation(n) (choice without replacement of
        # all n rows); with seed None it uses the global numpy state
        if self._random_seed is None:
            return np.random.permutation(num_rows)
        return np.random.RandomState(self._random_seed).permutation(num_rows)

    # ---- per-field decode (utils.decode_row semantics, petastorm/utils.py:52-85) --------------------------------
    def _decode_field(self, raw, name, field, order):
        count = raw.num_rows if order is None else len(order)
        if name in raw.partition_values:
            value = field.numpy_dtype(raw.partition_values[name]) if field.numpy_dtype else raw.partition_values[name]
            return [value] * count
        slot, col, leaf = self._leaf_of(raw, name)
        nulls = raw.decoded.null_counts[slot]
        try:
            codec = field.codec
            if isinstance(codec, (NdarrayCodec, CompressedImageCodec, CompressedNdarrayCodec)) and \
                    col.physical_type == BYTE_ARRAY:
                return self._decode_blobs(col, field, codec, order, nulls)
            return self._decode_scalars(col, leaf, field, order, nulls)
        except DecodeFieldError:
            raise
        except Exception as e:  # pylint: disable=broad-except
            raise DecodeFieldError('Decoding field "{}" failed'.format(name)).with_traceback(e.__traceback__)

    def _decode_scalars(self, col, leaf, field, order, nulls):
        """ScalarCodec / codec-less scalar: ``field.numpy_dtype(value)`` per row, ``None`` at nulls."""
        np_type = field.numpy_dtype
        if _is_temporal(leaf):
            vals = list(self._datetime_array(col, leaf, order, nulls))
        elif _is_host_only(leaf):
            vals = self._host_objects(col, leaf, order)
        else:
            t = self._numeric_tensor(col, leaf)
            cast_ok = np_type is not None and isinstance(np_type, type) and issubclass(np_type, np.generic)
            if not nulls and (field.codec is None or isinstance(field.codec, ScalarCodec)) and cast_ok and \
                    np.dtype(np_type) in _TORCH_OF_NUMPY:
                if order is not None:
                    t = device_ops.gather_rows(t.contiguous(), torch.from_numpy(np.ascontiguousarray(order)).to(t.device))
                if _TORCH_OF_NUMPY[np.dtype(np_type)] != t.dtype:
                    t = t.to(_TORCH_OF_NUMPY[np.dtype(np_type)])  # e.g. ShortType storage of a uint8 field
                return ScalarColumn(t, np_type)
            arr = t.cpu().numpy()
            if order is not None:
                arr = arr[order]
            vals = list(arr)
            if nulls:
                valid = col.valid.cpu().numpy().astype(bool)
                if order is not None:
                    valid = valid[order]
                vals = [v if ok else None for v, ok in zip(vals, valid)]
        cast = np_type is not None and isinstance(np_type, type) and issubclass(np_type, (np.generic, Decimal))
        if field.codec is not None or cast:
            if np_type is Decimal:
                return [None if v is None else Decimal(v) for v in vals]
            return [None if v is None else np_type(v) for v in vals]
        return vals

    def _decode_blobs(self, col, field, codec, order, nulls):
        """NdarrayCodec / CompressedImageCodec / CompressedNdarrayCodec columns -> per-row indexable."""
        device = col.arena.device
        n_all = col.num_values
        src_rows = np.arange(n_all) if order is None else np.asarray(order)
        if nulls:
            valid = col.valid.cpu().numpy().astype(bool)
            present = valid[src_rows]
        else:
            present = np.ones(len(src_rows), dtype=bool)
        live = src_rows[present]
        live_dev = None
        if order is not None or nulls:
            live_dev = torch.from_numpy(np.ascontiguousarray(live.astype(np.int64))).to(device)
        if isinstance(codec, CompressedNdarrayCodec):
            dense = self._decode_zipnpy(col, field, codec, live, live_dev)
        elif isinstance(codec, NdarrayCodec):
            dense = self._decode_npy(col, field, live, live_dev)
        elif codec.image_codec == 'png':
            dense = self._decode_png(col, field, live, live_dev)
        elif codec.image_codec in ('jpeg', 'jpg'):
            dense = self._decode_jpeg(col, field, live)
        else:
            raise ValueError('unsupported image codec {}'.format(codec.image_codec))
        if present.all():
            return dense
        out, k = [], 0
        for ok in present:
            if ok:
                out.append(dense[k])
                k += 1
            else:
                out.append(None)
        return out

    def _decode_zipnpy(self, col, field, codec, live, live_dev):
        """CompressedNdarrayCodec: the first blob is opened on the host to learn the .npy header and size of the
        member; all blobs are then inflated on the device (one warp each) and decoded like NdarrayCodec values.
        Ragged shapes, string dtypes or anything unexpected fall back to the reference's host decode."""
        if len(live) == 0:
            return []
        import io
        import zipfile
        try:
            first = rowgroup.gather_blobs_to_host(col, live[:1])[0]
            with zipfile.ZipFile(io.BytesIO(first)) as z:
                member = z.read(z.infolist()[0])
            dtype, shape, fortran, data_off = parse_npy_header(member[:min(len(member), 1024)])
            ok = dtype in _TORCH_OF_NUMPY and not fortran and dtype.byteorder in ('=', '<', '|')
        except Exception:  # pylint: disable=broad-except
            ok = False
        if ok:
            payload = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
            if data_off + payload == len(member):
                dense_blobs, st = device_ops.zip_inflate_batch(col, len(member), live_dev)
                if int(device_ops.to_host(st)[0]) == 0:
                    out, status = device_ops.npy_batch(dense_blobs, data_off, payload, _TORCH_OF_NUMPY[dtype], shape)
                    if int(device_ops.to_host(status)[0]) == 0:
                        return out
        blobs = rowgroup.gather_blobs_to_host(col, live)
        return [codec.decode(field, b) for b in blobs]

    def _decode_npy(self, col, field, live, live_dev):
        if len(live) == 0:
            return []
        device = col.arena.device
        heads = device_ops.to_host(device_ops.blob_prefix(col, 256)).numpy()   # one small D2H for all headers of the row-group
        groups = {}
        for pos, r in enumerate(live):
            b = heads[r].tobytes()
            hl = _npy_header_len(b)
            groups.setdefault(b[:hl] if 0 < hl <= len(b) else None, []).append(pos)
        if len(groups) == 1 and None not in groups:
            dtype, shape, fortran, data_off = parse_npy_header(next(iter(groups)))
            if dtype in _TORCH_OF_NUMPY and not fortran and dtype.byteorder in ('=', '<', '|'):
                payload = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
                out, status = device_ops.npy_batch(col, data_off, payload, _TORCH_OF_NUMPY[dtype], shape, live_dev)
                if int(device_ops.to_host(status)[0]) == 0:
                    return out
        # ragged shapes / string dtypes / fortran order: decode each group; string arrays have no tensor form
        result = [None] * len(live)
        for head, positions in groups.items():
            rows = live[positions]
            ok = False
            if head is not None:    # None: header longer than the prefix -> host decode
                try:
                    dtype, shape, fortran, data_off = parse_npy_header(head)
                    ok = dtype in _TORCH_OF_NUMPY and not fortran and dtype.byteorder in ('=', '<', '|')
                except Exception:  # pylint: disable=broad-except
                    ok = False
            if ok:
                payload = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
                idx = torch.from_numpy(np.ascontiguousarray(rows.astype(np.int64))).to(device)
                out, status = device_ops.npy_batch(col, data_off, payload, _TORCH_OF_NUMPY[dtype], shape, idx)
                if int(device_ops.to_host(status)[0]) == 0:
                    for k, p in enumerate(positions):
                        result[p] = out[k]
                    continue
            blobs = rowgroup.gather_blobs_to_host(col, rows)
            for p, b in zip(positions, blobs):
                result[p] = NdarrayCodec().decode(field, b)
        return result

    def _decode_png(self, col, field, live, live_dev):
        if len(live) == 0:
            return []
        device = col.arena.device
        np_dtype = np.dtype(field.numpy_dtype)
        shape = field.shape
        if shape and None not in shape and len(shape) in (2, 3) and np_dtype in (np.dtype('uint8'), np.dtype('uint16')):
            ch = shape[2] if len(shape) == 3 else 1
            out, status = device_ops.png_batch(col, shape[0], shape[1], ch, _TORCH_OF_NUMPY[np_dtype], live_dev)
            st = device_ops.to_host(status).tolist()
            if st[0] == 0:
                return out
            if st[0] == 7:
                raise ValueError('corrupt PNG stream in row {} of the row-group'.format(st[1]))
            # geometry differs from the schema (or unsupported variant): take the per-header path below
        heads = device_ops.to_host(device_ops.blob_prefix(col, 33)).numpy()
        groups = {}
        for pos, r in enumerate(live):
            h = heads[r]
            w, hgt = int.from_bytes(h[16:20].tobytes(), 'big'), int.from_bytes(h[20:24].tobytes(), 'big')
            groups.setdefault((hgt, w, int(h[24]), int(h[25])), []).append(pos)
        result = [None] * len(live)
        for (hgt, w, depth, ctype), positions in groups.items():
            ch = {0: 1, 2: 3, 3: 3}.get(ctype)
            if ch is None or depth not in (8, 16):
                raise ValueError('Unexpected image dimensions. Supported dimensions are (H, W) or (H, W, 3).')
            idx = torch.from_numpy(np.ascontiguousarray(live[positions].astype(np.int64))).to(device)
            out, status = device_ops.png_batch(col, hgt, w, ch, torch.uint8 if depth == 8 else torch.uint16, idx)
            st = device_ops.to_host(status).tolist()
            if st[0] != 0:
                raise ValueError('PNG decode failed (code {}) in row {}'.format(st[0], st[1]))
            for k, p in enumerate(positions):
                result[p] = out[k]
        return result

    def _decode_jpeg(self, col, field, live):
        """JPEG column -> ``[n, H, W, 3]`` RGB through nvJPEG.  The bitstreams stay in HBM (the Parquet decode left every
        blob in the arena): the host reads the 12 bytes of (offset, length) per image and the first bytes of every
        stream (SOF marker: geometry), and hands nvJPEG device pointers.  Without a device-bitstream backend the blobs
        are staged through the host (the GPU still decodes)."""
        if len(live) == 0:
            return []
        live = np.asarray(live)
        if device_ops.jpeg_device_backend() < 0 or _FORCE_HOST_JPEG[0]:
            return self._decode_jpeg_host_staged(col, field, live)
        offs = device_ops.to_host(col.offs).numpy()[live]
        lens = device_ops.to_host(col.lens).numpy()[live]
        # geometry of every stream from its SOF marker (nvJPEG writes what the stream says: the output buffers must fit)
        hw = _jpeg_sizes(device_ops.to_host(device_ops.blob_prefix(col, 1024)).numpy()[live])
        shape = field.shape
        if shape and None not in shape and len(shape) == 3 and shape[2] == 3:
            bad = np.nonzero((hw[:, 0] != shape[0]) | (hw[:, 1] != shape[1]))[0]
            if len(bad):
                raise ValueError('JPEG {} is {}x{}, expected {}x{}'.format(int(bad[0]), int(hw[bad[0], 0]),
                                                                          int(hw[bad[0], 1]), shape[0], shape[1]))
        uniq = np.unique(hw, axis=0)
        dims = {(int(h), int(w)): np.nonzero((hw[:, 0] == h) & (hw[:, 1] == w))[0] for h, w in uniq}
        try:
            if len(dims) == 1:
                (h, w), _ = next(iter(dims.items()))
                return device_ops.jpeg_batch_device(col, offs, lens, h, w)
            result = [None] * len(live)
            for (h, w), pos in dims.items():
                out = device_ops.jpeg_batch_device(col, offs[pos], lens[pos], h, w)
                for k, p in enumerate(pos):
                    result[p] = out[k]
            return result
        except native.NativeLibraryError:
            # e.g. a progressive stream the hardware engines refuse: the host-staged backend reports precise errors
            return self._decode_jpeg_host_staged(col, field, live)

    def _decode_jpeg_host_staged(self, col, field, live):
        blobs = rowgroup.gather_blobs_to_host(col, live)
        shape = field.shape
        if shape and None not in shape and len(shape) == 3 and shape[2] == 3:
            return device_ops.jpeg_batch(blobs, shape[0], shape[1], col.arena.device)
        # variable geometry: group by the SOF dimensions
        dims = [_jpeg_size(b) for b in blobs]
        result = [None] * len(blobs)
        for d in set(dims):
            pos = [i for i, x in enumerate(dims) if x == d]
            out = device_ops.jpeg_batch([blobs[i] for i in pos], d[0], d[1], col.arena.device)
            for k, p in enumerate(pos):
                result[p] = out[k]
        return result

    # ---- row-group loads ----------------------------------------------------------------------------------------
    def _decode_all(self, raw, names, order):
        if raw.decoded is not None:
            raw.decoded.check()
            raw.decoded.wait()
        return {name: self._decode_field(raw, name, self._schema.fields[name], order) for name in names}

    def _issue_rows(self, piece, shuffle_row_drop_partition):
        names = [f.name for f in self._schema.fields.values()]
        raw = self._read_raw(piece, names)
        order = self._row_order(raw.num_rows, shuffle_row_drop_partition,
                                self._ngram.length if self._ngram else 0)
        count = raw.num_rows if order is None else len(order)

        def finalize():
            return self._finish_rows(self._decode_issued(raw, names, order, count))

        return PendingRowGroup(finalize, count)

    def _decode_issued(self, raw, names, order, count):
        with torch.cuda.stream(self._stream_of(raw)):
            cols = self._decode_all(raw, names, order)
            if self._transform_spec:
                cols = self._apply_transform(cols, count)
            done = torch.cuda.Event()
            done.record()
        return GpuRowGroupRows(cols, count, [_EventWaiter(done), raw.decoded])

    def _load_rows(self, piece, shuffle_row_drop_partition):
        names = [f.name for f in self._schema.fields.values()]
        raw = self._read_raw(piece, names)
        order = self._row_order(raw.num_rows, shuffle_row_drop_partition,
                                self._ngram.length if self._ngram else 0)
        return self._decode_issued(raw, names, order, raw.num_rows if order is None else len(order))

    def _apply_transform(self, cols, count):
        spec = self._transform_spec
        if spec.func:
            if spec.device:
                scalar_types = {k: v.np_type for k, v in cols.items() if isinstance(v, ScalarColumn)}
                cols = spec.func({k: (v.tensor if isinstance(v, ScalarColumn) else v) for k, v in cols.items()})
                for k, np_type in scalar_types.items():   # untouched scalar columns keep their per-row numpy form
                    if isinstance(cols.get(k), torch.Tensor) and cols[k].dim() == 1 and \
                            _TORCH_OF_NUMPY.get(np.dtype(np_type)) == cols[k].dtype:
                        cols[k] = ScalarColumn(cols[k], np_type)
            else:
                # opaque user code: row dicts of host values, exactly what upstream passes
                # (petastorm/py_dict_reader_worker.py:38-52)
                rows = [spec.func({k: _npify(v[i]) if not isinstance(v, ScalarColumn) else v[i] for k, v in cols.items()})
                        for i in range(count)]
                keys = list(rows[0].keys()) if rows else list(cols.keys())
                cols = {k: [r[k] for r in rows] for k in keys}
        for name in spec.removed_fields:
            cols.pop(name, None)
        return cols

    def _load_rows_with_predicate(self, piece, worker_predicate, shuffle_row_drop_partition, defer=False):
        """Two-phase read (petastorm/py_dict_reader_worker.py:216-262): predicate columns, mask, then the other columns
        of the matching rows only.  With ``defer`` phase two is only *issued* (H2D + page decode queued on a decode
        stream) and a :class:`PendingRowGroup` is returned whose resolution builds the columns - the payload transfer
        of this row-group then overlaps the consumer's work on the previous one.  None when nothing matches."""
        predicate_fields, all_names = self._validate_predicate_fields(worker_predicate, self._schema)
        # partition columns ride along with every read upstream (legacy pyarrow appended them to each piece.read), so
        # they stay in the second read's field set; _read_raw serves them from the piece's partition keys
        other_names = all_names - predicate_fields
        raw_p = self._read_raw(piece, predicate_fields)
        order = self._row_order(raw_p.num_rows, shuffle_row_drop_partition,
                                self._ngram.length if self._ngram else 0)
        with torch.cuda.stream(self._stream_of(raw_p)):
            pcols = self._decode_all(raw_p, sorted(predicate_fields), order)
            count = raw_p.num_rows if order is None else len(order)
            mask = self._predicate_mask(worker_predicate, raw_p, pcols, order, count)
            keep_host = np.nonzero(mask)[0]
            if len(keep_host) == 0:
                return None
            sel = keep_host if order is None else np.asarray(order)[keep_host]
            cols = {}
            for name, v in pcols.items():
                if isinstance(v, ScalarColumn):
                    cols[name] = v.take(torch.from_numpy(keep_host.astype(np.int64)).to(v.tensor.device))
                elif isinstance(v, torch.Tensor):
                    cols[name] = device_ops.gather_rows(v.contiguous(),
                                                        torch.from_numpy(keep_host.astype(np.int64)).to(v.device))
                else:
                    cols[name] = [v[i] for i in keep_host]
            phase1 = torch.cuda.Event()
            phase1.record()
        raw_o = self._read_raw(piece, other_names) if other_names else None    # issued, not waited for

        def finalize():
            with torch.cuda.stream(self._stream_of(raw_p)):
                if raw_o is not None:
                    cols.update(self._decode_all(raw_o, sorted(other_names), sel))
                out = self._apply_transform(cols, len(keep_host)) if self._transform_spec else cols
                done = torch.cuda.Event()
                done.record()
            return GpuRowGroupRows(out, len(keep_host),
                                   [_EventWaiter(done), raw_p.decoded, raw_o.decoded if raw_o else None])

        if defer:
            return PendingRowGroup(lambda: self._finish_rows(finalize()), len(keep_host))
        return finalize()

    def _predicate_mask(self, predicate, raw, pcols, order, count):
        """Boolean host mask of the rows to keep.  Device form when the predicate has one for these columns."""
        dev_cols = {}
        for name in pcols:
            if name in raw.partition_values:
                continue
            slot, col, leaf = self._leaf_of(raw, name)
            if raw.decoded.null_counts[slot] == 0 and not _is_host_only(leaf) and not _is_temporal(leaf) and \
                    col.max_rep == 0 and col.physical_type in (INT32, INT64):
                t = self._numeric_tensor(col, leaf)
                if order is not None:
                    t = device_ops.gather_rows(t.contiguous(), torch.from_numpy(np.ascontiguousarray(order)).to(t.device))
                dev_cols[name] = t
        if len(dev_cols) == len(pcols):
            mask = predicate.device_mask(dev_cols)
            if mask is not None:
                return mask.cpu().numpy().astype(bool)
        # user-defined / host-only predicate: row-by-row like upstream (py_dict_reader_worker.py:232)
        return np.array([bool(predicate.do_include({k: (v[i] if isinstance(v, ScalarColumn) else _npify(v[i]))
                                                      for k, v in pcols.items()}))
                         for i in range(count)], dtype=bool)

    # ---- NGram --------------------------------------------------------------------------------------------------
    def _form_ngram(self, rows):
        """Window starts of a decoded row-group (petastorm/ngram.py:225-270).  The timestamp column was produced on the
        post-processing stream, so the validity kernel and the compaction run there too."""
        ts_name = self._ngram.timestamp_field.name
        ts = rows.columns[ts_name]
        if isinstance(ts, ScalarColumn):
            ts = ts.tensor
        device = self._get_decoder().device
        ts_list = [_npify(v) for v in ts] if not isinstance(ts, torch.Tensor) else None
        starts_dev = None
        with torch.cuda.stream(self._thread_post_stream()):
            if ts_list is not None and len(ts_list) and isinstance(ts_list[0], (np.integer, int)) and \
                    all(v is not None for v in ts_list):
                ts_dev = torch.tensor(np.asarray(ts_list, dtype=np.int64), device=device)
                starts_dev = self._ngram.window_starts_device(ts_dev)
            elif isinstance(ts, torch.Tensor) and not ts.is_floating_point():
                starts_dev = self._ngram.window_starts_device(ts.to(torch.int64))
            if starts_dev is not None:
                starts = starts_dev.cpu().tolist()
        if starts_dev is None:
            starts = self._ngram.window_starts_host(ts_list if ts_list is not None else list(ts.cpu().numpy()))
        return GpuNGramWindows(rows, starts, self._ngram, starts_dev)


# ---- host images of decoded row-groups (LocalDiskCache) -------------------------------------------------------------------
def _host_value(v):
    if isinstance(v, ScalarColumn):
        return ('scalar', v._values(), v.np_type)  # pylint: disable=protected-access
    if isinstance(v, torch.Tensor):
        return ('tensor', v.cpu().numpy())
    if isinstance(v, list) and any(isinstance(x, torch.Tensor) for x in v):
        return ('list', [x.cpu().numpy() if isinstance(x, torch.Tensor) else x for x in v], True)
    return ('host', v)


def _device_value(item, device):
    kind = item[0]
    if kind == 'scalar':
        return ScalarColumn(torch.from_numpy(item[1]).to(device), item[2], item[1])
    if kind == 'tensor':
        return torch.from_numpy(item[1]).to(device)
    if kind == 'list':
        return [torch.from_numpy(x).to(device) if isinstance(x, np.ndarray) and x.dtype in _TORCH_OF_NUMPY else x
                for x in item[1]]
    return item[1]


def to_host_payload(value):
    """Picklable host image of a decoded row-group (:class:`GpuBatch` / :class:`GpuRowGroupRows`; None stays None)."""
    if value is None:
        return None
    if isinstance(value, PendingRowGroup):
        value = value.resolve()
    value.wait()
    torch.cuda.current_stream().synchronize()
    kind = 'batch' if isinstance(value, GpuBatch) else 'rows'
    return {'kind': kind, 'num_rows': value.num_rows, 'columns': {k: _host_value(v) for k, v in value.columns.items()}}


def from_host_payload(payload, device=None):
    """Inverse of :func:`to_host_payload`: the columns go back to the (current) device with one H2D copy each."""
    if payload is None:
        return None
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
    cols = {k: _device_value(v, device) for k, v in payload['columns'].items()}
    cls = GpuBatch if payload['kind'] == 'batch' else GpuRowGroupRows
    return cls(cols, payload['num_rows'], [])


def _npy_header_len(prefix):
    """Total header length (magic + version + length field + dict) of a .npy blob, 0 if it is not one."""
    if prefix[:6] != b'\x93NUMPY':
        return 0
    if prefix[6] == 1:
        return int.from_bytes(prefix[8:10], 'little') + 10
    return int.from_bytes(prefix[8:12], 'little') + 12


def _jpeg_sizes(heads):
    """(height, width) of n JPEG streams from their first bytes (uint8 [n, k]).  Streams written by one encoder share
    their header layout, so the SOF position of the first stream is tried on all of them at once; the rest is parsed
    one by one."""
    n = len(heads)
    out = np.zeros((n, 2), dtype=np.int64)
    todo = np.arange(n)
    if n:
        b = heads[0].tobytes()
        i = 2
        pos = -1
        while i + 9 < len(b):
            if b[i] != 0xFF:
                i += 1
                continue
            m = b[i + 1]
            if m in (0xC0, 0xC1, 0xC2):
                pos = i
                break
            if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
                i += 2
                continue
            i += 2 + int.from_bytes(b[i + 2:i + 4], 'big')
        if pos >= 0 and pos + 9 <= heads.shape[1]:
            same = (heads[:, pos] == 0xFF) & np.isin(heads[:, pos + 1], (0xC0, 0xC1, 0xC2)) & \
                (heads[:, :pos] == heads[0, :pos]).all(axis=1)
            h16 = heads[:, pos + 5].astype(np.int64) * 256 + heads[:, pos + 6]
            w16 = heads[:, pos + 7].astype(np.int64) * 256 + heads[:, pos + 8]
            out[same, 0], out[same, 1] = h16[same], w16[same]
            todo = np.nonzero(~same)[0]
    for i in todo:
        out[i] = _jpeg_size(heads[i].tobytes())
    return out


def _jpeg_size(blob):
    """(height, width) from the SOF marker of a JPEG stream."""
    i = 2
    n = len(blob)
    while i + 9 < n:
        if blob[i] != 0xFF:
            i += 1
            continue
        marker = blob[i + 1]
        if marker in (0xC0, 0xC1, 0xC2):
            return (int.from_bytes(blob[i + 5:i + 7], 'big'), int.from_bytes(blob[i + 7:i + 9], 'big'))
        if marker in (0xD8, 0x01) or 0xD0 <= marker <= 0xD7:
            i += 2
            continue
        i += 2 + int.from_bytes(blob[i + 2:i + 4], 'big')
    raise ValueError('no SOF marker in JPEG stream')
