"""Row-group decode on the device: plan (host) -> H2D of raw page bytes -> CUDA decode -> typed column tensors.

This is the B200 replacement of ``piece.read(columns=...)`` (Arrow C++) that both reference workers call:
``petastorm/arrow_reader_worker.py:358`` and ``petastorm/py_dict_reader_worker.py:267``.

Ownership rule: every decoded row-group owns its HBM (``arena`` = raw page bytes + decompression scratch, ``out`` =
decoded columns) through ordinary torch tensors; column tensors handed to users are views that keep ``out`` alive, so
nothing is recycled under the user (the reference never reuses output buffers either - SURVEY 8b).
"""
import collections
import os
import threading
from ctypes import byref, c_int

import numpy as np
import time

import torch

from petastorm_b200 import native

# parquet physical types
BOOLEAN, INT32, INT64, INT96, FLOAT, DOUBLE, BYTE_ARRAY, FIXED_LEN_BYTE_ARRAY = range(8)

_DEVICE_ERRORS = {
    1: 'corrupt Snappy stream', 2: 'corrupt definition/repetition levels or value stream',
    3: 'unsupported page encoding', 4: 'dictionary index out of range', 5: 'page values overrun the page',
    6: 'NdarrayCodec blobs in one batch do not share the same .npy header', 7: 'corrupt PNG stream',
    8: 'unsupported PNG variant (interlaced / alpha / bit depth / unexpected geometry)',
    9: 'NGram assumes that the data is sorted by the timestamp field which is not the case',
    10: 'corrupt BYTE_ARRAY page', 11: 'corrupt GZIP page',
}


class DeviceDecodeError(RuntimeError):
    pass


_ctx_lock = threading.Lock()
_contexts = {}
_files = collections.OrderedDict()

#: default budget of the pinned row-group cache (host RAM kept page-locked so that re-reading a row-group in a later
#: epoch is one cudaMemcpyAsync); override with set_pinned_cache_bytes() before the first reader is created
_pinned_cache_bytes = [4 << 30]


def set_pinned_cache_bytes(nbytes):
    """Budget of the pinned row-group cache, for contexts created later and for the live ones (shrinking drops what
    they hold)."""
    _pinned_cache_bytes[0] = int(nbytes)
    with _ctx_lock:
        for ctx in _contexts.values():
            ctx.set_pinned_cache_bytes(int(nbytes))


#: budget of the HBM-resident raw row-group cache (0 = off).  180 GB of HBM3e hold the *encoded* bytes of datasets far
#: larger than what a training job re-reads per epoch (the nominal C2 dataset is 43 GB encoded): with the cache on, the
#: raw region of a row-group stays in HBM after its first decode and later epochs decode it again without touching PCIe.
_hbm_cache_bytes = [0]


def set_hbm_cache_bytes(nbytes):
    _hbm_cache_bytes[0] = int(nbytes)


def get_context(device=None):
    """Process-wide native context of a CUDA device (created lazily)."""
    if not torch.cuda.is_available():
        raise native.NativeLibraryError('petastorm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    if device is None:
        device = torch.cuda.current_device()
    with _ctx_lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = native.Context(device, _pinned_cache_bytes[0], -1)
            _contexts[device] = ctx
        return ctx


OPEN_FILE_CACHE_ENTRIES = 1024


def open_file(path):
    """Cached :class:`native.ParquetFile` (mmap + parsed footer) - host only.  The cache is an LRU: an evicted file is
    unmapped once the last plan that refers to it is gone (plans keep their file alive)."""
    with _ctx_lock:
        f = _files.get(path)
        if f is None:
            f = native.ParquetFile(path)
            _files[path] = f
            if len(_files) > OPEN_FILE_CACHE_ENTRIES:
                _files.popitem(last=False)
        else:
            _files.move_to_end(path)
        return f


def forget_file(path):
    """Drops one file from the open-file cache (it was rewritten).  The mapping itself is released when the last plan
    or decoded row-group that refers to it is gone (they hold a reference): closing it here would leave them with a
    dangling ``pst_file``."""
    with _ctx_lock:
        _files.pop(path, None)


def forget_files():
    with _ctx_lock:
        _files.clear()


_TORCH_OF_PHYSICAL = {INT32: torch.int32, INT64: torch.int64, FLOAT: torch.float32, DOUBLE: torch.float64,
                      BOOLEAN: torch.uint8}


class DecodedColumn(object):
    """Device-resident decode result of one leaf column of one row-group."""

    __slots__ = ('leaf', 'physical_type', 'num_values', 'values', 'valid', 'offs', 'lens', 'rep', 'defs', 'max_def',
                 'max_rep', 'arena')

    def null_count(self):
        if self.valid is None:
            return 0
        return int(self.num_values - int(self.valid.sum().item()))


# diagnostics: set to a list to collect (host time, [before H2D, after H2D, after decode] timing events) per row-group
TRACE = None


class DecodedRowGroup(object):
    def __init__(self, plan, arena, out, status, stream, device):
        self.plan = plan
        self.arena = arena
        self.out = out
        self.status = status
        self.stream = stream
        self.device = device
        self.num_rows = plan.info.num_rows
        self._checked = False
        self.owns_arena = any(pc.physical_type == BYTE_ARRAY for pc in plan.cols)
        self.null_counts = None  # per plan column, filled by check()
        with torch.cuda.stream(stream):
            # the status words travel to pinned host memory on the decode stream, so check() needs no device copy
            self.host_status = torch.empty(status.shape, dtype=status.dtype, pin_memory=True)
            self.host_status.copy_(status, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record(stream)

    def wait(self, stream=None):
        """Make `stream` (default: the current stream) wait for the decode; cheap, no host sync."""
        stream = stream or torch.cuda.current_stream(self.device)
        stream.wait_event(self.event)
        # the buffers were allocated on the decode stream: tell the caching allocator about the consumer stream
        self.out.record_stream(stream)
        if self.owns_arena:
            self.arena.record_stream(stream)

    def check(self):
        """Host-synchronises with the decode and raises if a kernel reported a malformed page."""
        if self._checked:
            return
        # poll instead of cudaEventSynchronize: a host thread parked inside a blocking CUDA wait delays CUDA calls
        # issued by the row-group issuing thread (measured: 6 ms per cudaEventRecord), which stalls the pipeline
        spins = 0
        while not self.event.query():
            spins += 1
            time.sleep(0 if spins < 50 else 0.0001)
        st = self.host_status.tolist()
        self._checked = True
        self.null_counts = st[8:]
        if st[0] != 0:
            raise DeviceDecodeError('{} (file {}, row-group {}, page table entry {}, detail {})'.format(
                _DEVICE_ERRORS.get(st[0], 'device decode error %d' % st[0]), self.plan.file.path,
                self.plan.row_group, st[1], st[2]))
        for slot, pc in enumerate(self.plan.cols):
            # a flat nullable column planned without a validity array (chunk statistics: null_count == 0) must not
            # contain a single level entry below max_def
            if pc.max_def > 0 and pc.max_rep == 0 and pc.valid_off < 0 and self.null_counts[slot] != 0:
                raise DeviceDecodeError('column chunk statistics of leaf {} claim null_count == 0 but its pages hold {} '
                                        'nulls (file {}, row-group {})'.format(pc.column, self.null_counts[slot],
                                                                               self.plan.file.path, self.plan.row_group))

    def column(self, slot):
        pc = self.plan.cols[slot]
        n = pc.num_values
        col = DecodedColumn()
        col.leaf = pc.column
        col.physical_type = pc.physical_type
        col.num_values = n
        col.max_def = pc.max_def
        col.max_rep = pc.max_rep
        col.arena = self.arena
        col.values = col.offs = col.lens = col.valid = col.rep = col.defs = None
        out = self.out
        if pc.physical_type == BYTE_ARRAY:
            col.offs = out[pc.values_off:pc.values_off + 8 * n].view(torch.int64)
            col.lens = out[pc.lens_off:pc.lens_off + 4 * n].view(torch.int32)
        elif pc.physical_type in _TORCH_OF_PHYSICAL:
            dt = _TORCH_OF_PHYSICAL[pc.physical_type]
            col.values = out[pc.values_off:pc.values_off + pc.type_length * n].view(dt)
        else:  # INT96 / FIXED_LEN_BYTE_ARRAY: raw bytes [n, width]
            col.values = out[pc.values_off:pc.values_off + pc.type_length * n].view(n, pc.type_length)
        if pc.valid_off >= 0:
            col.valid = out[pc.valid_off:pc.valid_off + n]
        if pc.rep_off >= 0:
            col.rep = out[pc.rep_off:pc.rep_off + n]
            col.defs = out[pc.def_off:pc.def_off + n]
        return col


class RowGroupDecoder(object):
    """Issues plan -> upload -> decode for row-groups on side streams of one device."""

    NUM_STREAMS = int(os.environ.get('PST_DECODE_STREAMS', '5'))

    def __init__(self, device=None):
        self.ctx = get_context(device)
        self.device = torch.device('cuda', self.ctx.device)
        # consecutive row-groups go to different streams: a row-group's decode is latency-bound (~15 ms for 256 MB, the
        # serial Snappy streams) while its H2D copy takes ~5 ms, so ~4 decodes must be in flight to keep PCIe busy
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.NUM_STREAMS)]
        # one recyclable arena per stream: work on a stream is ordered, so the H2D copy of the next row-group may
        # overwrite the arena as soon as it is *enqueued* behind the previous decode (numeric-only plans; BYTE_ARRAY
        # columns keep pointing into their arena and get a private one)
        self._stream_arena = {}
        self._plans = collections.OrderedDict()
        self._hbm_cache = {}          # (path, row-group, columns) -> [arena, last-use event]
        self._hbm_cache_used = 0
        self.hbm_cache_hits = 0
        self._next_stream = 0
        self.stream = self.streams[0]
        self.launches = 0
        self.h2d_bytes = 0
        self.host_seconds = [0.0, 0.0, 0.0]  # plan / upload / decode-issue (diagnostics)

    PLAN_CACHE_ENTRIES = 128

    def plan(self, path, row_group, leaf_columns):
        """Plans are immutable (page walk + HBM layout of one row-group / column set): later epochs reuse them."""
        key = (path, row_group, tuple(leaf_columns))
        plan = self._plans.get(key)
        if plan is None:
            plan = native.Plan(open_file(path), row_group, leaf_columns)
            self._plans[key] = plan
            if len(self._plans) > self.PLAN_CACHE_ENTRIES:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return plan

    def next_stream(self):
        s = self.streams[self._next_stream]
        self._next_stream = (self._next_stream + 1) % len(self.streams)
        self.stream = s
        return s

    def upload(self, plan, stream=None, private=False):
        """H2D of a plan's raw region into an arena; returns the arena tensor (async on `stream`).  `private=True`
        forces a dedicated arena (the caller wants to keep the raw bytes resident)."""
        stream = stream or self.stream
        recyclable = not private and all(pc.physical_type != BYTE_ARRAY for pc in plan.cols)
        key = stream.cuda_stream
        arena = self._stream_arena.get(key) if recyclable else None
        if arena is None or arena.numel() < plan.info.arena_bytes:
            with torch.cuda.stream(stream):
                arena = torch.empty(plan.info.arena_bytes + (plan.info.arena_bytes >> 3 if recyclable else 0),
                                    dtype=torch.uint8, device=self.device)
            if recyclable:
                self._stream_arena[key] = arena
        native.check(native.lib.pst_plan_upload(self.ctx.handle, plan.handle, arena.data_ptr(),
                                                stream.cuda_stream), 'pst_plan_upload')
        self.h2d_bytes += plan.info.raw_bytes
        return arena

    def decode_resident(self, plan, arena, stream=None):
        """Device decode of a plan whose raw region is already resident in `arena` (no PCIe traffic)."""
        stream = stream or self.stream
        with torch.cuda.stream(stream):
            out = torch.empty(plan.info.out_bytes, dtype=torch.uint8, device=self.device)
            status = torch.zeros(8 + len(plan.cols), dtype=torch.int32, device=self.device)
        nl = c_int(0)
        native.check(native.lib.pst_plan_decode(self.ctx.handle, plan.handle, arena.data_ptr(), out.data_ptr(),
                                                status.data_ptr(), stream.cuda_stream, byref(nl)), 'pst_plan_decode')
        self.launches += nl.value
        return DecodedRowGroup(plan, arena, out, status, stream, self.device)

    def decode(self, path, row_group, leaf_columns):
        t0 = time.perf_counter()
        plan = self.plan(path, row_group, leaf_columns)
        stream = self.next_stream()
        t1 = time.perf_counter()
        if TRACE is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record(stream)
        entry = None
        if _hbm_cache_bytes[0] > 0:
            key = (path, row_group, tuple(leaf_columns))
            entry = self._hbm_cache.get(key)
            if entry is not None:
                arena = entry[0]
                stream.wait_event(entry[1])      # the previous decode of this arena (scratch is rewritten)
                self.hbm_cache_hits += 1
            elif self._hbm_cache_used + plan.info.arena_bytes <= _hbm_cache_bytes[0]:
                arena = self.upload(plan, stream, private=True)
                entry = self._hbm_cache[key] = [arena, None]
                self._hbm_cache_used += plan.info.arena_bytes
            else:
                arena = self.upload(plan, stream)
        else:
            arena = self.upload(plan, stream)
        if TRACE is not None:
            ev[1].record(stream)
        t2 = time.perf_counter()
        d = self.decode_resident(plan, arena, stream)
        if entry is not None:
            entry[1] = d.event
        if TRACE is not None:
            ev[2].record(stream)
            TRACE.append((time.perf_counter(), ev))
        t3 = time.perf_counter()
        self.host_seconds[0] += t1 - t0
        self.host_seconds[1] += t2 - t1
        self.host_seconds[2] += t3 - t2
        return d


def gather_blobs_to_host(col, row_indices=None):
    """BYTE_ARRAY column -> list of python ``bytes`` (or None) on the host.

    Used for the values that cannot be tensors (strings, decimals) - the reference DataLoader rejects those as well
    (petastorm/pytorch.py:64-66) - after the device did the decompress / level / dictionary work."""
    offs = col.offs.cpu().numpy()
    lens = col.lens.cpu().numpy()
    valid = col.valid.cpu().numpy() if col.valid is not None else None
    n = len(offs)
    if n == 0:
        return []
    idx = np.arange(n) if row_indices is None else np.asarray(row_indices)
    lo = int(offs[idx].min()) if len(idx) else 0
    hi = int((offs[idx] + lens[idx]).max()) if len(idx) else 0
    # one contiguous D2H of the span that holds the values (values of a page are contiguous; dictionaries small)
    span = col.arena[lo:hi].cpu().numpy().tobytes() if hi > lo else b''
    res = []
    for i in idx:
        if valid is not None and not valid[i]:
            res.append(None)
        else:
            o = int(offs[i]) - lo
            res.append(span[o:o + int(lens[i])])
    return res
