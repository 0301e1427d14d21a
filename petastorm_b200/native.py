"""ctypes binding of ``libpst_b200.so`` (C-ABI declared in ``include/pst_b200.h``).

This module is the only place that touches the shared library.  There is deliberately NO fallback: if the library is
missing, importing :mod:`petastorm_b200.native` raises, and every GPU entry point raises if CUDA is unavailable.
"""
import ctypes
import json
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint8, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PST_B200_LIB') or os.path.join(_HERE, 'libpst_b200.so')   # override: kernel experiments


class NativeLibraryError(RuntimeError):
    """libpst_b200.so is missing or a call into it failed."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            'petastorm_b200: the CUDA extension {} has not been built. Run `python -c "import __graft_entry__ as g; '
            'g.build()"` (or `make -C petastorm_b200/csrc`). There is no CPU fallback.'.format(LIB_PATH))
    # make sure the CUDA runtime torch ships is the one bound (same SONAME libcudart.so.12)
    try:
        import torch  # noqa: F401  pylint: disable=unused-import,import-outside-toplevel
    except Exception:  # pragma: no cover
        pass
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()


class ChunkInfo(Structure):
    _fields_ = [('physical_type', c_int32), ('codec', c_int32), ('num_values', c_int64),
                ('data_page_offset', c_int64), ('dictionary_page_offset', c_int64),
                ('total_compressed_size', c_int64), ('total_uncompressed_size', c_int64), ('start_offset', c_int64)]


class PlanInfo(Structure):
    _fields_ = [('num_rows', c_int64), ('raw_bytes', c_int64), ('arena_bytes', c_int64), ('out_bytes', c_int64),
                ('payload_bytes', c_int64), ('uncompressed_bytes', c_int64), ('num_pages', c_int32),
                ('num_columns', c_int32), ('num_compressed_pages', c_int32), ('num_index_pages', c_int32),
                ('num_unwrapped_pages', c_int32), ('num_copy_tiles', c_int32), ('num_decode_pages', c_int32),
                ('num_snappy_fragments', c_int32), ('num_host_indexed_pages', c_int32), ('num_cluster_index_pages', c_int32)]


class PlanColumn(Structure):
    _fields_ = [('column', c_int32), ('physical_type', c_int32), ('type_length', c_int32), ('max_def', c_int32),
                ('max_rep', c_int32), ('has_dictionary', c_int32), ('num_values', c_int64), ('values_off', c_int64),
                ('lens_off', c_int64), ('valid_off', c_int64), ('rep_off', c_int64), ('def_off', c_int64)]


class PlanPage(Structure):
    _fields_ = [('column_slot', c_int32), ('kind', c_int32), ('encoding', c_int32), ('codec', c_int32),
                ('flags', c_int32), ('stored_bytes', c_int32), ('image_bytes', c_int32), ('num_values', c_int32),
                ('first_value', c_int32), ('fragments', c_int32), ('src_off', c_int64), ('img_off', c_int64)]


class CopyTile(Structure):
    _fields_ = [('src_off', c_int64), ('dst_off', c_int64), ('valid_off', c_int64), ('nbytes', c_int32),
                ('nvalid', c_int32)]


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_sig('pst_last_error', c_char_p)
_sig('pst_abi_version', c_int)
_sig('pst_has_cuda', c_int)
_sig('pst_file_open', c_int, c_char_p, POINTER(c_void_p))
_sig('pst_file_close', None, c_void_p)
_sig('pst_file_num_row_groups', c_int, c_void_p)
_sig('pst_file_num_rows', c_int64, c_void_p)
_sig('pst_file_row_group_num_rows', c_int64, c_void_p, c_int)
_sig('pst_file_num_columns', c_int, c_void_p)
_sig('pst_file_schema_json', c_int, c_void_p, POINTER(c_char_p), POINTER(c_size_t))
_sig('pst_file_kv_metadata', c_int, c_void_p, c_char_p, POINTER(POINTER(c_uint8)), POINTER(c_size_t))
_sig('pst_file_num_kv', c_int, c_void_p)
_sig('pst_file_kv_at', c_int, c_void_p, c_int, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_void_p),
     POINTER(c_size_t))
_sig('pst_file_chunk_info', c_int, c_void_p, c_int, c_int, POINTER(ChunkInfo))
_sig('pst_plan_create', c_int, c_void_p, c_int, POINTER(c_int), c_int, POINTER(c_void_p))
_sig('pst_plan_destroy', None, c_void_p)
_sig('pst_plan_get_info', c_int, c_void_p, POINTER(PlanInfo))
_sig('pst_plan_get_column', c_int, c_void_p, c_int, POINTER(PlanColumn))
_sig('pst_plan_fill_raw', c_int, c_void_p, c_void_p, c_int64, c_int64)
_sig('pst_plan_get_page', c_int, c_void_p, c_int, POINTER(PlanPage))
_sig('pst_plan_get_copy_tile', c_int, c_void_p, c_int, POINTER(CopyTile))
_sig('pst_ctx_create', c_int, c_int, c_int64, c_int, POINTER(c_void_p))
_sig('pst_ctx_destroy', None, c_void_p)
_sig('pst_ctx_stats_json', c_int, c_void_p, c_char_p, c_size_t)
_sig('pst_ctx_set_pinned_cache_bytes', c_int, c_void_p, c_int64)
_sig('pst_plan_upload', c_int, c_void_p, c_void_p, c_uint64, c_uint64)
_sig('pst_plan_decode', c_int, c_void_p, c_void_p, c_uint64, c_uint64, c_uint64, c_uint64, POINTER(c_int))
_sig('pst_plan_decode_timed', c_int, c_void_p, c_void_p, c_uint64, c_uint64, c_uint64, c_uint64, POINTER(c_float))
_sig('pst_nullable_to_f64', c_int, c_uint64, c_uint64, c_int64, c_int, c_int, c_int, c_uint64, c_uint64)
_sig('pst_narrow_int32', c_int, c_uint64, c_int64, c_int, c_uint64, c_uint64)
_sig('pst_gather_rows', c_int, c_uint64, c_uint64, c_int64, c_int64, c_uint64, c_uint64)
_sig('pst_npy_batch', c_int, c_uint64, c_uint64, c_uint64, c_uint64, c_int64, c_int64, c_int64, c_uint64, c_uint64,
     c_uint64)
_sig('pst_zip_inflate_batch', c_int, c_uint64, c_uint64, c_uint64, c_uint64, c_int64, c_int64, c_uint64, c_uint64, c_uint64)
_sig('pst_blob_prefix', c_int, c_uint64, c_uint64, c_uint64, c_int64, c_int, c_uint64, c_uint64)
_sig('pst_png_work_bytes', c_int64, c_int, c_int, c_int, c_int)
_sig('pst_png_batch', c_int, c_uint64, c_uint64, c_uint64, c_uint64, c_int64, c_int, c_int, c_int, c_int, c_uint64,
     c_uint64, c_uint64, c_uint64)
_sig('pst_jpeg_available', c_int)
_sig('pst_jpeg_backend', c_int)
_sig('pst_jpeg_batch', c_int, c_void_p, POINTER(c_void_p), POINTER(c_size_t), c_int64, c_int, c_int, c_uint64, c_uint64)
_sig('pst_jpeg_device_backend', c_int)
_sig('pst_jpeg_batch_device', c_int, c_void_p, c_uint64, c_void_p, c_void_p, c_int64, c_int, c_int, c_uint64, c_uint64)
_sig('pst_mask_in_set_i64', c_int, c_uint64, c_int, c_int, c_int64, c_uint64, c_int64, c_uint64, c_uint64)
_sig('pst_mask_md5_split_i64', c_int, c_uint64, c_int, c_int, c_int64, c_double, c_double, c_uint64, c_uint64)
_sig('pst_compact_tmp_bytes', c_int64, c_int64)
_sig('pst_mask_compact', c_int, c_uint64, c_int64, c_uint64, c_uint64, c_uint64, c_uint64)
_sig('pst_normalize', c_int, c_uint64, c_int, c_int64, c_float, c_float, c_uint64, c_int, c_uint64)
_sig('pst_ngram_valid_starts', c_int, c_uint64, c_int64, c_int, c_int64, c_uint64, c_uint64, c_uint64)
_sig('pst_ngram_gather', c_int, c_uint64, c_uint64, c_int64, c_int, c_int64, c_uint64, c_uint64)
_sig('pst_sanitize', c_int, c_uint64, c_int64, c_int, c_uint64, c_uint64)
_sig('pst_list_uniform', c_int, c_uint64, c_uint64, c_int64, c_int, c_int64, c_uint64, c_uint64)

#: every symbol include/pst_b200.h declares (checked by tests/test_abi.py)
EXPORTED = [
    'pst_last_error', 'pst_abi_version', 'pst_has_cuda', 'pst_file_open', 'pst_file_close',
    'pst_file_num_row_groups', 'pst_file_num_rows', 'pst_file_row_group_num_rows', 'pst_file_num_columns',
    'pst_file_schema_json', 'pst_file_kv_metadata', 'pst_file_num_kv', 'pst_file_kv_at', 'pst_file_chunk_info',
    'pst_plan_create', 'pst_plan_destroy', 'pst_plan_get_info', 'pst_plan_get_column', 'pst_plan_fill_raw',
    'pst_plan_get_page', 'pst_plan_get_copy_tile',
    'pst_ctx_create', 'pst_ctx_destroy', 'pst_ctx_stats_json', 'pst_ctx_set_pinned_cache_bytes', 'pst_plan_upload', 'pst_plan_decode', 'pst_plan_decode_timed',
    'pst_nullable_to_f64', 'pst_narrow_int32', 'pst_gather_rows', 'pst_npy_batch', 'pst_blob_prefix', 'pst_zip_inflate_batch', 'pst_png_work_bytes',
    'pst_png_batch', 'pst_jpeg_available', 'pst_jpeg_backend', 'pst_jpeg_batch', 'pst_jpeg_device_backend', 'pst_jpeg_batch_device', 'pst_mask_in_set_i64',
    'pst_mask_md5_split_i64', 'pst_compact_tmp_bytes', 'pst_mask_compact', 'pst_normalize',
    'pst_ngram_valid_starts', 'pst_ngram_gather', 'pst_sanitize', 'pst_list_uniform',
]


def last_error():
    msg = lib.pst_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(rc, what='libpst_b200 call'):
    if rc != 0:
        raise NativeLibraryError('{} failed: {}'.format(what, last_error()))


class ParquetFile(object):
    """An mmapped Parquet file with its parsed footer (``pst_file``)."""

    def __init__(self, path):
        self.path = path
        self._h = c_void_p()
        check(lib.pst_file_open(os.fsencode(path), byref(self._h)), 'open {}'.format(path))
        js, n = c_char_p(), c_size_t()
        check(lib.pst_file_schema_json(self._h, byref(js), byref(n)))
        self.schema = json.loads(ctypes.string_at(js, n.value).decode('utf-8'))
        self.num_row_groups = lib.pst_file_num_row_groups(self._h)
        self.num_rows = lib.pst_file_num_rows(self._h)
        self.num_columns = lib.pst_file_num_columns(self._h)

    @property
    def handle(self):
        return self._h

    def row_group_num_rows(self, rg):
        return lib.pst_file_row_group_num_rows(self._h, rg)

    def key_value_metadata(self):
        out = {}
        for i in range(lib.pst_file_num_kv(self._h)):
            k, kl, v, vl = c_void_p(), c_size_t(), c_void_p(), c_size_t()
            check(lib.pst_file_kv_at(self._h, i, byref(k), byref(kl), byref(v), byref(vl)))
            out[ctypes.string_at(k, kl.value)] = ctypes.string_at(v, vl.value)
        return out

    def chunk_info(self, rg, col):
        ci = ChunkInfo()
        check(lib.pst_file_chunk_info(self._h, rg, col, byref(ci)))
        return ci

    def close(self):
        if self._h:
            lib.pst_file_close(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


class Plan(object):
    """Row-group plan: page table + HBM layout for a set of leaf columns (``pst_plan``)."""

    def __init__(self, pfile, row_group, columns):
        self.file = pfile
        self.row_group = row_group
        self.columns = list(columns)
        arr = (c_int * len(self.columns))(*self.columns)
        self._h = c_void_p()
        check(lib.pst_plan_create(pfile.handle, row_group, arr, len(self.columns), byref(self._h)),
              'plan row-group {} of {}'.format(row_group, pfile.path))
        self.info = PlanInfo()
        check(lib.pst_plan_get_info(self._h, byref(self.info)))
        self.cols = []
        for i in range(len(self.columns)):
            pc = PlanColumn()
            check(lib.pst_plan_get_column(self._h, i, byref(pc)))
            self.cols.append(pc)

    @property
    def handle(self):
        return self._h

    def page(self, i):
        pg = PlanPage()
        check(lib.pst_plan_get_page(self._h, i, byref(pg)))
        return pg

    def copy_tile(self, i):
        t = CopyTile()
        check(lib.pst_plan_get_copy_tile(self._h, i, byref(t)))
        return t

    def fill_raw(self, buf_address):
        check(lib.pst_plan_fill_raw(self._h, c_void_p(buf_address), 0, 1 << 62))

    def close(self):
        if self._h:
            lib.pst_plan_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass


class Context(object):
    """Device context: pinned staging ring / pinned row-group cache / copy threads (``pst_ctx``)."""

    def __init__(self, device=0, pinned_cache_bytes=0, copy_threads=-1):
        import torch  # pylint: disable=import-outside-toplevel
        if not torch.cuda.is_available():
            raise NativeLibraryError('petastorm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.device = device
        self._h = c_void_p()
        check(lib.pst_ctx_create(device, pinned_cache_bytes, copy_threads, byref(self._h)), 'pst_ctx_create')

    @property
    def handle(self):
        return self._h

    def set_pinned_cache_bytes(self, nbytes):
        check(lib.pst_ctx_set_pinned_cache_bytes(self._h, int(nbytes)), 'pst_ctx_set_pinned_cache_bytes')

    def stats(self):
        buf = ctypes.create_string_buffer(1024)
        check(lib.pst_ctx_stats_json(self._h, buf, 1024))
        return json.loads(buf.value.decode())

    def close(self):
        if self._h:
            lib.pst_ctx_destroy(self._h)
            self._h = c_void_p()
