#!/usr/bin/env python
"""Probe of the Blackwell hardware decompression engine (cuMemBatchDecompressAsync, CUDA >= 12.8) on raw-Snappy Parquet
page payloads.  Measurement aid only (not part of the product path): prints whether the engine is exposed on this box,
whether torch-allocated memory is decompress-capable, correctness against the host, and throughput for the page shapes
of the C2 row-group (160 KB int64 pages, 1 MiB dictionary pages, incompressible float pages)."""
import ctypes
import json
import sys
import time

import numpy as np
import pyarrow as pa
import torch


class Params(ctypes.Structure):
    _fields_ = [('srcNumBytes', ctypes.c_size_t), ('dstNumBytes', ctypes.c_size_t), ('dstActBytes', ctypes.c_void_p),
                ('src', ctypes.c_void_p), ('dst', ctypes.c_void_p), ('algo', ctypes.c_uint32),
                ('padding', ctypes.c_ubyte * 20)]


def main():
    import subprocess
    if len(sys.argv) == 1:
        variants = [('torch:pinned', 'runs', 1), ('torch:device', 'runs', 1), ('torch:pinned', 'i64_160k', 4),
                    ('torch:pinned', 'i64_160k', 4, 3, 0), ('torch:pinned', 'i64_160k', 4, 0, 9),
                    ('torch:pinned', 'i64_160k', 560), ('torch:pinned', 'i64_1m', 16),
                    ('torch:pinned', 'f32_80k', 2240), ('torch:pinned', 'runs', 256),
                    ('torch:pinned', 'i64_160k', 2800), ('torch:device', 'i64_160k', 560)]
        for v in variants:
            r = subprocess.run([sys.executable, __file__] + [str(x) for x in v], capture_output=True, text=True,
                               timeout=120)
            print(r.stdout.strip()[-1500:], flush=True)
            if r.returncode != 0:
                print('  rc', r.returncode, r.stderr.strip()[-600:], flush=True)
        return 0
    out = {}
    mem_kind, params_kind = sys.argv[1].split(':')
    one = [sys.argv[2]] + [int(x) for x in sys.argv[3:]]
    cu = ctypes.CDLL('libcuda.so.1')
    cu.cuMemAlloc_v2.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t]
    cu.cuMemcpyHtoD_v2.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
    cu.cuMemcpyDtoH_v2.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t]
    cu.cuMemsetD8_v2.argtypes = [ctypes.c_uint64, ctypes.c_ubyte, ctypes.c_size_t]
    cu.cuPointerGetAttribute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64]

    class Raw(object):
        """device buffer from cuMemAlloc (documented as decompress-capable)"""
        def __init__(self, nbytes, host=None):
            self.n = nbytes
            p = ctypes.c_uint64(0)
            rc = cu.cuMemAlloc_v2(ctypes.byref(p), max(nbytes, 256))
            assert rc == 0, rc
            self.ptr = p.value
            if host is not None:
                assert cu.cuMemcpyHtoD_v2(self.ptr, host.ctypes.data, nbytes) == 0
            else:
                assert cu.cuMemsetD8_v2(self.ptr, 0, max(nbytes, 256)) == 0
        def data_ptr(self):
            return self.ptr
        def numpy(self):
            h = np.empty(self.n, dtype=np.uint8)
            rc = cu.cuMemcpyDtoH_v2(h.ctypes.data, self.ptr, self.n)
            assert rc == 0, 'DtoH rc %d' % rc
            return h

    class Tor(object):
        def __init__(self, nbytes, host=None):
            self.t = torch.from_numpy(host).to('cuda:0') if host is not None else \
                torch.zeros(nbytes, dtype=torch.uint8, device='cuda:0')
        def data_ptr(self):
            return self.t.data_ptr()
        def numpy(self):
            return self.t.cpu().numpy()
    Buf = Raw if mem_kind == 'cumem' else Tor
    torch.cuda.init()
    dev = torch.device('cuda', 0)
    torch.zeros(1, device=dev)
    val = ctypes.c_int(0)
    rc = cu.cuDeviceGetAttribute(ctypes.byref(val), 136, 0)
    out['algo_mask_rc'], out['algo_mask'] = rc, val.value
    rc = cu.cuDeviceGetAttribute(ctypes.byref(val), 137, 0)
    out['max_len_rc'], out['max_len'] = rc, val.value
    if not hasattr(cu, 'cuMemBatchDecompressAsync'):
        out['error'] = 'driver has no cuMemBatchDecompressAsync'
        print(json.dumps(out))
        return
    fn = cu.cuMemBatchDecompressAsync
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]

    rng = np.random.default_rng(7)

    def make_pages(kind, n_pages):
        pages = []
        for _ in range(n_pages):
            if kind == 'i64_160k':
                raw = rng.integers(0, 2 ** 40, 20000, dtype=np.int64).tobytes()
            elif kind == 'i64_1m':
                raw = rng.integers(0, 2 ** 40, 131072, dtype=np.int64).tobytes()
            elif kind == 'f32_80k':
                raw = rng.standard_normal(20000).astype(np.float32).tobytes()
            else:  # highly compressible
                raw = (np.arange(40000, dtype=np.int32) // 64).tobytes()
            comp = pa.compress(raw, codec='snappy', asbytes=True)
            pages.append((raw, comp))
        return pages

    def run(kind, n_pages, src_mis=0, dst_mis=0, reps=5):
        pages = make_pages(kind, n_pages)
        comp_total = sum(len(c) + 64 for _, c in pages) + 64
        raw_total = sum(len(r) + 64 for r, _ in pages) + 64
        host = np.zeros(comp_total, dtype=np.uint8)
        srcs, dsts, pos, dpos = [], [], src_mis, dst_mis
        for r, c in pages:
            host[pos:pos + len(c)] = np.frombuffer(c, dtype=np.uint8)
            srcs.append(pos)
            dsts.append(dpos)
            pos = (pos + len(c) + 63) // 64 * 64 + src_mis
            dpos = (dpos + len(r) + 63) // 64 * 64 + dst_mis
        d_src = Buf(comp_total, host)
        d_dst = Buf(raw_total)
        d_act = Buf(4 * n_pages)
        cap = ctypes.c_int(0)
        cu.cuPointerGetAttribute(ctypes.byref(cap), 21, d_src.data_ptr())
        print(json.dumps({'mem': mem_kind, 'capable_attr': cap.value}), flush=True)
        # the descriptor array is consumed by the GPU after the call returns: keep it in pinned host or device memory
        pbuf = torch.zeros(64 * n_pages, dtype=torch.uint8).pin_memory()
        arr = (Params * n_pages).from_address(pbuf.data_ptr())
        for i, (r, c) in enumerate(pages):
            arr[i].srcNumBytes = len(c)
            arr[i].dstNumBytes = len(r)
            arr[i].dstActBytes = d_act.data_ptr() + 4 * i
            arr[i].src = d_src.data_ptr() + srcs[i]
            arr[i].dst = d_dst.data_ptr() + dsts[i]
            arr[i].algo = 2
        stream = torch.cuda.current_stream().cuda_stream
        err = ctypes.c_size_t(0)
        arr_arg = ctypes.c_void_p(pbuf.data_ptr())
        if params_kind == 'device':
            dbuf = pbuf.to('cuda:0')
            torch.cuda.synchronize()
            arr_arg = ctypes.c_void_p(dbuf.data_ptr())
        res_extra = {'params': params_kind}
        res = {'kind': kind, 'pages': n_pages, 'src_mis': src_mis, 'dst_mis': dst_mis, 'capable': cap.value,
               'comp_bytes': sum(len(c) for _, c in pages), 'raw_bytes': sum(len(r) for r, _ in pages)}
        rc = fn(arr_arg, n_pages, 0, ctypes.byref(err), stream)
        torch.cuda.synchronize()
        res['rc'] = rc
        res.update(res_extra)
        if rc != 0:
            res['err_index'] = err.value
            return res
        got = d_dst.numpy()
        ok = all(got[dsts[i]:dsts[i] + len(r)].tobytes() == r for i, (r, _) in enumerate(pages))
        res['correct'] = bool(ok)
        res['act_ok'] = bool((d_act.numpy().view(np.uint32) == np.array([len(r) for r, _ in pages])).all())
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            fn(arr_arg, n_pages, 0, ctypes.byref(err), stream)
            host_ms = (time.perf_counter() - t0) * 1e3
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res['ms'] = best
        res['submit_host_ms'] = host_ms
        res['out_gbps'] = res['raw_bytes'] / best / 1e6
        return res

    try:
        res = run(*one)
    except Exception as e:  # pylint: disable=broad-except
        res = {'args': one, 'exception': repr(e)[:300]}
    res['mem'] = mem_kind
    res.update(out)
    print(json.dumps(res), flush=True)


if __name__ == '__main__':
    sys.exit(main())
