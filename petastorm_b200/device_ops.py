"""Torch-tensor front-ends of the post-processing kernels exported by libpst_b200.so.

Every function enqueues hand-written CUDA work on the *current* torch stream and returns device tensors; none of them
falls back to torch/ATen math for the operation it names.
"""
import math

import numpy as np
import torch

from petastorm_b200.native import lib, check

_NORM_DTYPE = {torch.uint8: 0, torch.float16: 1, torch.float32: 2, torch.int32: 3, torch.int16: 4, torch.uint16: 5,
               torch.float64: 6}


#: C-ABI kernel-launching calls made through this module (bench.py's gpu_launches); a counter, not a lock-step log
LAUNCH_CALLS = [0]


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream(device_index=None):
    """Handle of the current torch stream (of ``device_index`` when the caller knows it: the raw lookup costs a fraction
    of a microsecond, ``torch.cuda.current_stream()`` ~15 - this runs once per kernel launch, and the per-batch calls of
    the loaders are host-bound)."""
    LAUNCH_CALLS[0] += 1
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device() if device_index is None else device_index)
    return torch.cuda.current_stream().cuda_stream


def to_host(t):
    """Device tensor -> pinned host tensor WITHOUT a blocking CUDA call: an asynchronous copy on the current stream and a
    polled event.  ``tensor.cpu()`` / ``.item()`` park the calling thread inside the driver, and a thread parked there
    delays the CUDA calls of the row-group issuing thread (6 ms per ``cudaEventRecord`` measured):
    every host read on the consumer / resolver side of the readers goes through here."""
    import time
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    spins = 0
    while not ev.query():
        spins += 1
        time.sleep(0 if spins < 50 else 0.0001)
    return host


def _status(device):
    return torch.zeros(8, dtype=torch.int32, device=device)


def gather_rows(src, index):
    """out[i] = src[index[i]] along dim 0 (K13; ``table.take`` / ``DataFrame.sample`` of the reference workers)."""
    assert src.is_cuda and src.is_contiguous()
    if index.dtype != torch.int64 or index.device != src.device or not index.is_contiguous():
        index = index.to(device=src.device, dtype=torch.int64).contiguous()
    n = index.numel()
    inner = src.shape[1:]
    out = torch.empty((n,) + tuple(inner), dtype=src.dtype, device=src.device)
    if n == 0:
        return out
    row_bytes = src.element_size() * math.prod(inner)
    check(lib.pst_gather_rows(src.data_ptr(), index.data_ptr(), n, row_bytes, out.data_ptr(), _stream(src.device.index)),
          'gather_rows')
    return out


def nulls_to_nan(values, valid, physical_type, bit_width=0, is_unsigned=False):
    """pandas null semantics on device: int -> float64 with NaN, float keeps dtype with NaN (K6)."""
    n = values.numel()
    from petastorm_b200 import rowgroup
    out_dtype = torch.float32 if physical_type == rowgroup.FLOAT else torch.float64
    out = torch.empty(n, dtype=out_dtype, device=values.device)
    check(lib.pst_nullable_to_f64(values.data_ptr(), valid.data_ptr(), n, physical_type, bit_width,
                                  1 if is_unsigned else 0, out.data_ptr(), _stream()), 'nullable_to_f64')
    return out


def narrow_int32(values, torch_dtype):
    """INT32 storage -> int8/uint8/int16/uint16 logical type."""
    bits = torch.empty(0, dtype=torch_dtype).element_size() * 8
    out = torch.empty(values.numel(), dtype=torch_dtype, device=values.device)
    check(lib.pst_narrow_int32(values.data_ptr(), values.numel(), bits, out.data_ptr(), _stream()), 'narrow_int32')
    return out


def mask_in_set(keys, sorted_set):
    """mask[i] = keys[i] in set (``in_set``, petastorm/predicates.py:44-55)."""
    n = keys.numel()
    mask = torch.empty(n, dtype=torch.uint8, device=keys.device)
    unsigned = keys.dtype in (torch.uint8, torch.uint16, torch.uint32, torch.uint64)
    check(lib.pst_mask_in_set_i64(keys.data_ptr(), keys.element_size(), 1 if unsigned else 0, n, sorted_set.data_ptr(),
                                  sorted_set.numel(), mask.data_ptr(), _stream()), 'mask_in_set')
    return mask


def mask_md5_split(keys, bucket_low, bucket_high):
    """``in_pseudorandom_split`` on an integer key column: md5(str(key)) bucket in [low, high) (predicates.py:144-182).

    bucket_low/high are the reference's python floats; python compares the exact int bucket with them, which for an
    integer bucket is the same as comparing with ceil(float) as an integer."""
    n = keys.numel()
    mask = torch.empty(n, dtype=torch.uint8, device=keys.device)
    unsigned = keys.dtype in (torch.uint8, torch.uint16, torch.uint32, torch.uint64)
    lo = float(math.ceil(bucket_low))
    hi = float(math.ceil(bucket_high))
    check(lib.pst_mask_md5_split_i64(keys.data_ptr(), keys.element_size(), 1 if unsigned else 0, n, lo, hi,
                                     mask.data_ptr(), _stream()), 'mask_md5_split')
    return mask


def mask_to_indices(mask):
    """Ascending row indices where mask != 0 (stream compaction).  Synchronises to learn the count."""
    n = mask.numel()
    idx = torch.empty(n, dtype=torch.int64, device=mask.device)
    count = torch.zeros(1, dtype=torch.int64, device=mask.device)
    tmp = torch.empty(max(int(lib.pst_compact_tmp_bytes(n)), 8), dtype=torch.uint8, device=mask.device)
    check(lib.pst_mask_compact(mask.data_ptr(), n, idx.data_ptr(), count.data_ptr(), tmp.data_ptr(), _stream()),
          'mask_compact')
    k = int(count.item())
    return idx[:k]


def normalize(src, mean, std, out_dtype=torch.float32):
    """((float32)x - mean) / std cast to out_dtype, IEEE fp32 arithmetic (K12)."""
    assert src.is_cuda and src.is_contiguous()
    out = torch.empty(src.shape, dtype=out_dtype, device=src.device)
    check(lib.pst_normalize(src.data_ptr(), _NORM_DTYPE[src.dtype], src.numel(), float(mean), float(std),
                            out.data_ptr(), _NORM_DTYPE[out_dtype], _stream()), 'normalize')
    return out


def ngram_valid_starts(ts, length, delta_threshold):
    """uint8 mask of window starts + status tensor (K15, petastorm/ngram.py:225-270)."""
    ts = ts.contiguous()
    ok = torch.empty(ts.numel(), dtype=torch.uint8, device=ts.device)
    status = _status(ts.device)
    check(lib.pst_ngram_valid_starts(ts.data_ptr(), ts.numel(), int(length), int(delta_threshold), ok.data_ptr(),
                                     status.data_ptr(), _stream()), 'ngram_valid_starts')
    return ok, status


def ngram_gather(src, starts, length):
    """out[w, t] = src[starts[w] + t] for t < length."""
    assert src.is_contiguous()
    starts = starts.to(torch.int64).contiguous()
    w = starts.numel()
    out = torch.empty((w, length) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if w == 0:
        return out
    row_bytes = src.element_size() * int(np.prod(src.shape[1:], dtype=np.int64)) if src.dim() > 1 else src.element_size()
    check(lib.pst_ngram_gather(src.data_ptr(), starts.data_ptr(), w, int(length), row_bytes, out.data_ptr(), _stream()),
          'ngram_gather')
    return out


def sanitize(src):
    """``_sanitize_pytorch_types`` promotions on device (petastorm/pytorch.py:40-70)."""
    if src.dtype == torch.uint16:
        kind, dt = 0, torch.int32
    elif src.dtype == torch.uint32:
        kind, dt = 1, torch.int64
    elif src.dtype == torch.bool:
        kind, dt = 2, torch.uint8
    else:
        return src
    src = src.contiguous()
    out = torch.empty(src.shape, dtype=dt, device=src.device)
    check(lib.pst_sanitize(src.data_ptr(), src.numel(), kind, out.data_ptr(), _stream()), 'sanitize')
    return out


def npy_batch(col, data_off, payload_bytes, torch_dtype, shape, row_index=None):
    """NdarrayCodec payloads of a BYTE_ARRAY column -> [n, *shape] tensor (K7)."""
    n = col.num_values if row_index is None else row_index.numel()
    dev = col.offs.device
    out = torch.empty((n,) + tuple(shape), dtype=torch_dtype, device=dev)
    status = _status(dev)
    if n:
        assert out.numel() * out.element_size() == n * payload_bytes
        check(lib.pst_npy_batch(col.arena.data_ptr(), col.offs.data_ptr(), col.lens.data_ptr(),
                                0 if row_index is None else row_index.data_ptr(), n, data_off, payload_bytes,
                                out.data_ptr(), status.data_ptr(), _stream()), 'npy_batch')
    return out, status


class _DenseBlobs(object):
    """Fixed-size blobs stored back to back in one device tensor, addressed like a BYTE_ARRAY column."""

    def __init__(self, data, n, blob_bytes):
        self.arena = data
        self.num_values = n
        self.offs = torch.arange(n, dtype=torch.int64, device=data.device) * blob_bytes
        self.lens = torch.full((n,), blob_bytes, dtype=torch.int32, device=data.device)


def zip_inflate_batch(col, member_bytes, row_index=None):
    """First member of every (selected) ZIP blob of a BYTE_ARRAY column -> _DenseBlobs of .npy images + status."""
    n = col.num_values if row_index is None else row_index.numel()
    dev = col.offs.device
    out = torch.empty((n, member_bytes), dtype=torch.uint8, device=dev)
    status = _status(dev)
    if n:
        check(lib.pst_zip_inflate_batch(col.arena.data_ptr(), col.offs.data_ptr(), col.lens.data_ptr(),
                                        0 if row_index is None else row_index.data_ptr(), n, member_bytes,
                                        out.data_ptr(), status.data_ptr(), _stream()), 'zip_inflate_batch')
    return _DenseBlobs(out, n, member_bytes), status


def blob_prefix(col, k):
    """uint8 [n, k] tensor holding the first k bytes of every value of a BYTE_ARRAY column."""
    n = col.num_values
    out = torch.empty((n, k), dtype=torch.uint8, device=col.offs.device)
    if n:
        check(lib.pst_blob_prefix(col.arena.data_ptr(), col.offs.data_ptr(), col.lens.data_ptr(), n, k, out.data_ptr(),
                                  _stream()), 'blob_prefix')
    return out


def png_batch(col, height, width, channels, torch_dtype, row_index=None):
    """PNG blobs of a BYTE_ARRAY column -> [n, H, W, C] (or [n, H, W]) tensor in RGB order (K8)."""
    n = col.num_values if row_index is None else row_index.numel()
    dev = col.offs.device
    sb = torch.empty(0, dtype=torch_dtype).element_size()
    shape = (n, height, width) + ((channels,) if channels > 1 else ())
    out = torch.empty(shape, dtype=torch_dtype, device=dev)
    status = _status(dev)
    if n:
        work = torch.empty(n * int(lib.pst_png_work_bytes(height, width, channels, sb)), dtype=torch.uint8, device=dev)
        check(lib.pst_png_batch(col.arena.data_ptr(), col.offs.data_ptr(), col.lens.data_ptr(),
                                0 if row_index is None else row_index.data_ptr(), n, height, width, channels, sb,
                                out.data_ptr(), work.data_ptr(), status.data_ptr(), _stream()), 'png_batch')
    return out, status


def list_is_uniform(rep, defs, max_def, list_len):
    flags = torch.zeros(1, dtype=torch.int64, device=rep.device)
    check(lib.pst_list_uniform(rep.data_ptr(), defs.data_ptr(), rep.numel(), max_def, list_len, flags.data_ptr(),
                               _stream()), 'list_uniform')
    return int(flags.item()) == 0


def jpeg_batch(host_blobs, height, width, device):
    """list of python bytes (JPEG streams) -> uint8 [n, H, W, 3] RGB through nvJPEG (K9)."""
    import ctypes
    n = len(host_blobs)
    out = torch.empty((n, height, width, 3), dtype=torch.uint8, device=device)
    if n == 0:
        return out
    ptrs = (ctypes.c_void_p * n)()
    lens = (ctypes.c_size_t * n)()
    keep = []
    for i, b in enumerate(host_blobs):
        buf = ctypes.create_string_buffer(b, len(b)) if not isinstance(b, ctypes.Array) else b
        keep.append(buf)
        ptrs[i] = ctypes.cast(buf, ctypes.c_void_p)
        lens[i] = len(b)
    check(lib.pst_jpeg_batch(None, ptrs, lens, n, height, width, out.data_ptr(), _stream()), 'jpeg_batch')
    # nvJPEG may still be reading the host bitstreams asynchronously: keep them alive until the stream drained
    torch.cuda.current_stream().synchronize()
    del keep
    return out


def jpeg_batch_device(col, host_offs, host_lens, height, width):
    """JPEG streams of a BYTE_ARRAY column decoded where they are: ``host_offs`` / ``host_lens`` are the (arena offset,
    length) of the selected values as host numpy arrays (int64 / int32); the bitstreams never leave HBM (K9 through
    nvJPEG's device-bitstream backend)."""
    n = len(host_offs)
    out = torch.empty((n, height, width, 3), dtype=torch.uint8, device=col.arena.device)
    if n == 0:
        return out
    host_offs = np.ascontiguousarray(host_offs, dtype=np.int64)
    host_lens = np.ascontiguousarray(host_lens, dtype=np.int32)
    check(lib.pst_jpeg_batch_device(None, col.arena.data_ptr(), host_offs.ctypes.data, host_lens.ctypes.data, n, height,
                                    width, out.data_ptr(), _stream()), 'jpeg_batch_device')
    return out


def jpeg_device_backend():
    """nvjpegBackend_t of the device-bitstream handle (5 hardware engines, 4 GPU-assisted Huffman), -1 when unavailable."""
    return int(lib.pst_jpeg_device_backend())
