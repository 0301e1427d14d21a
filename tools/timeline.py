#!/usr/bin/env python
"""Device timeline of pipelined upload+decode at the RowGroupDecoder level (CUDA events on the decode streams)."""
import sys
import time
import torch
sys.path.insert(0, '.')
import bench
from petastorm_b200 import rowgroup

def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    url = bench.ensure_dataset(8)
    import glob
    files = sorted(glob.glob(url.replace('file://', '') + '/*.parquet'))
    dec = rowgroup.RowGroupDecoder()
    leaves = list(range(rowgroup.open_file(files[0]).num_columns))
    # warm the pinned cache and arenas
    for rep in range(2):
        for f in files:
            dec.decode(f, 0, leaves).check()
    torch.cuda.synchronize()
    N = 24
    recs = []
    inflight = []
    ref = torch.cuda.Event(enable_timing=True)
    ref.record()
    t0 = time.perf_counter()
    for i in range(N):
        f = files[i % len(files)]
        plan = dec.plan(f, 0, leaves)
        s = dec.next_stream()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(s)
        arena = dec.upload(plan, s)
        e1.record(s)
        d = dec.decode_resident(plan, arena, s)
        e2.record(s)
        recs.append((e0, e1, e2))
        inflight.append(d)
        if len(inflight) >= depth:
            inflight.pop(0).check()
    for d in inflight:
        d.check()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print('depth %d: %.2f ms per row-group' % (depth, wall / N * 1e3))
    for i, (e0, e1, e2) in enumerate(recs):
        print('rg %2d  h2d %8.2f .. %8.2f   decode .. %8.2f' % (i, ref.elapsed_time(e0), ref.elapsed_time(e1), ref.elapsed_time(e2)))


if __name__ == '__main__':
    main()
