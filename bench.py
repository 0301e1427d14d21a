#!/usr/bin/env python
"""bench.py -- headline benchmark of the Parquet -> GPU tensor read hot path (BASELINE.json configs[1], "C2").

Workload (config.workload): make_batch_reader over plain Parquet with 64 float32 + 16 int64 flat columns, Snappy,
pyarrow defaults otherwise (v1 pages, dictionary-then-PLAIN fallback, nullable columns), ~256 MB row-groups
(666,667 rows), consumer re-batches to 4096 rows.  A stated subset of the nominal 100 M rows is materialised
(`--row-groups`, default 8 = 5.3 M rows = 2 GB decoded) and cycled with num_epochs.

One *step* = one row-group through the hot path (plan -> raw bytes -> CUDA decode -> 4096-row batches).

  value        samples/s with the raw column-chunk bytes already resident in HBM (decode kernels only), CUDA events
  e2e          samples/s through the public API (make_batch_reader) from HOST buffers: pinned host -> H2D -> decode ->
               D2H of one 4096-row batch of one column per step; wall clock between device synchronisations
  roofline     HBM roofline of the dominant decode kernel, timed live with CUDA events on its launching stream
  cpu_baseline the oracle's restatement of the reference ProcessPool + ArrowReaderWorker path on this box's cores
               (bounded sample), rank 0 / N=1 only
  --impl reference   times that CPU path alone with the same metric/config keys

Multi-GPU (torchrun, one rank per GPU): rank 0 builds the row-group owner table, one NCCL broadcast distributes it,
then every rank decodes only its own row-groups (weak scaling: every rank does `steps` row-groups).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_PER_GROUP = 666667          # 64*4 + 16*8 = 384 B/row -> 256 MB decoded per row-group
N_F32, N_I64 = 64, 16
ROW_BYTES = N_F32 * 4 + N_I64 * 8
BATCH = 4096
DATA_DIR = os.environ.get('PST_BENCH_DIR', '/tmp/pst_bench_c2')


# ---------------------------------------------------------------------------------------------------------------------
# synthetic data (np.random.default_rng(1234); float32 ~ N(0,1), int64 ~ U[0, 2^40))
# ---------------------------------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the C2 row-group, from the committed ncu --set full capture
NCU_DRAM_SOURCE = None   # filled in from the committed ncu --set full capture of this round (profiles/r2_*.txt)
NCU_DRAM_BYTES_PER_LAUNCH = {}


def plan_algorithmic_bytes(plan):
    """Bytes each decode kernel has to move for one launch on this plan (one row-group), from the planner's page table:
    index reads the stored bytes of the multi-fragment Snappy pages; fragments read stored + write image bytes of every
    Snappy page; the tile copy reads and writes the value bytes of PLAIN pages without nulls; the page decoder reads
    the images of the remaining data pages (+ their dictionaries) and writes their values."""
    info = plan.info
    res = dict.fromkeys(['k_snappy_index', 'k_snappy_pages', 'k_snappy_pages(serial fallback)', 'k_ba_dict_index',
                         'k_copy_tiles', 'k_decode_pages'], 0)
    dict_bytes = {}
    for i in range(info.num_pages):
        pg = plan.page(i)
        if pg.fragments > 1:
            res['k_snappy_index'] += pg.stored_bytes
        if pg.fragments >= 1:
            res['k_snappy_pages'] += pg.stored_bytes + pg.image_bytes
        if pg.kind == 2:
            dict_bytes[pg.column_slot] = pg.image_bytes
        elif not pg.flags & 2:
            width = max(plan.cols[pg.column_slot].type_length, 1)
            res['k_decode_pages'] += pg.image_bytes + pg.num_values * width
    res['k_decode_pages'] += sum(dict_bytes.values())
    for i in range(info.num_copy_tiles):
        t = plan.copy_tile(i)
        res['k_copy_tiles'] += 2 * t.nbytes + t.nvalid
    return res


def _write_one(args):
    path, seed, rows = args
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    rng = np.random.default_rng(seed)
    cols = {}
    for i in range(N_F32):
        cols['f%02d' % i] = rng.standard_normal(rows, dtype=np.float32)
    for i in range(N_I64):
        cols['i%02d' % i] = rng.integers(0, 2 ** 40, rows, dtype=np.int64)
    pq.write_table(pa.table(cols), path + '.tmp', compression='snappy', row_group_size=rows)
    os.replace(path + '.tmp', path)
    return path


def ensure_dataset(n_groups, rows=ROWS_PER_GROUP):
    import multiprocessing as mp
    os.makedirs(DATA_DIR, exist_ok=True)
    todo = []
    for g in range(n_groups):
        p = os.path.join(DATA_DIR, 'part-%05d-r%d.parquet' % (g, rows))
        if not os.path.exists(p):
            todo.append((p, 1234 + g, rows))
    # stale files of another size would change the row-group list
    keep = set('part-%05d-r%d.parquet' % (g, rows) for g in range(n_groups))
    for name in os.listdir(DATA_DIR):
        if name not in keep:
            os.remove(os.path.join(DATA_DIR, name))
    if todo:
        with mp.get_context('spawn').Pool(min(len(todo), 16)) as pool:
            pool.map(_write_one, todo)
    return 'file://' + DATA_DIR


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler (profiling recipe: nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.active = True

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            if self.active:              # only the timed regions count
                self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            parts = [p.strip() for p in s.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port of the reference's ProcessPool + ArrowReaderWorker path
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_throughput(url, steps, warmup, workers, pool='process'):
    """rows/s of the CPU reader (oracle/port.py: process_pool_batches or thread_pool_batches) over `steps` row-groups
    after `warmup` row-groups (child start-up excluded by the warm-up, like petastorm/benchmark/throughput.py:68-90)."""
    from oracle import port
    total = warmup + steps
    n_groups = len(port.list_pieces(url))
    epochs = (total + n_groups - 1) // n_groups
    make = port.process_pool_batches if pool == 'process' else port.thread_pool_batches
    gen = make(url, workers, epochs=epochs)
    rows = 0
    t0 = None
    for k, batch in enumerate(gen):
        if k == warmup:
            t0 = time.perf_counter()
        if k >= warmup:
            # consumer side of the reference: re-batch to 4096 (views) -- arrow_reader_worker.py:97-111 + loader
            n = len(batch['f00'])
            for s in range(0, n, BATCH):
                _ = batch['f00'][s:s + BATCH]
            rows += n
        if k + 1 >= total:
            break
    dt = time.perf_counter() - t0
    gen.close()
    return rows / dt, rows, dt


def cpu_reference_best(url, steps, warmup, cores, process_workers):
    """The reference's reader in its three relevant pool configurations, each on a bounded sample; the fastest one is
    the baseline.  Returns (best rows/s, description dict)."""
    import pyarrow
    variants = []
    for pool, workers in (('thread', 10), ('thread', max(1, min(cores // 2, 64))), ('process', process_workers)):
        n = max(steps, workers if pool == 'process' else 192)   # ~10 s of CPU work per variant on a 128-core host
        v, rows, dt = cpu_reference_throughput(url, n, max(warmup, workers if pool == 'process' else 8), workers, pool)
        variants.append({'pool': pool, 'workers': workers, 'samples_per_sec': v, 'rows': rows, 'seconds': round(dt, 2)})
    best = max(variants, key=lambda x: x['samples_per_sec'])
    sample = ('%d rows in %.1f s, oracle/port.py restatement of the reference reader on its %s pool with %d workers '
              '(ParquetFile.read_row_group + take%s), pyarrow %s; best of thread x10 (reference default), '
              'thread x cores/2, process x%d' %
              (best['rows'], best['seconds'], best['pool'], best['workers'],
               ' + Arrow-IPC back to the consumer' if best['pool'] == 'process' else ', tables handed over in-process',
               pyarrow.__version__, process_workers))
    return best, variants, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=64)
    ap.add_argument('--warmup', type=int, default=8)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--row-groups', type=int, default=8, help='row-groups materialised per GPU box (cycled)')
    ap.add_argument('--rows-per-group', type=int, default=ROWS_PER_GROUP)
    ap.add_argument('--cpu-workers', type=int, default=0, help='processes of the CPU reference arm (0 = all cores, max 64)')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rows_pg = args.rows_per_group
    n_groups = max(args.row_groups, 2 * world)    # every rank cycles over at least two distinct row-groups
    cores = os.cpu_count() or 1
    cpu_workers = args.cpu_workers or min(cores, 64)

    workload = ('C2 make_batch_reader: {}xfloat32 + {}xint64 flat columns, Snappy, {} rows/row-group (~{} MB decoded), '
                'batch {}; {} row-groups materialised ({} rows) and cycled').format(
                    N_F32, N_I64, rows_pg, rows_pg * ROW_BYTES // 2 ** 20, BATCH, n_groups, n_groups * rows_pg)
    config = {'workload': workload, 'rows_per_row_group': rows_pg, 'row_groups_materialised': n_groups,
              'batch': BATCH, 'compression': 'snappy', 'parallelism': 'row-group shards, %d rank(s)' % world,
              'l2_policy': 'inputs larger than L2: each step reads a distinct ~%d MB arena' % (rows_pg * ROW_BYTES // 2 ** 20)}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == 'reference':
        if rank != 0:
            return
        url = ensure_dataset(n_groups, rows_pg)
        # three pool configurations of the reference on bounded samples; the fastest is the baseline
        best, variants, sample = cpu_reference_best(url, args.steps, args.warmup, cores, cpu_workers)
        value = best['samples_per_sec']
        line = {'impl': 'reference', 'metric': 'samples_per_sec', 'value': value, 'unit': 'samples/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': 1e3 * rows_pg / value,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': config, 'delivered_gbps': value * ROW_BYTES / 1e9,
                'cpu_baseline': {'value': value, 'unit': 'samples/s',
                                 'cores': cores if best['pool'] == 'thread' else best['workers'], 'kind': 'port',
                                 'sample': sample, 'variants': variants, 'host_cores': cores},
                'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    # ----------------------------------------------------------------------------------------------------- B200 arm
    import petastorm_b200  # noqa: F401  (first: sets CUDA_DEVICE_MAX_CONNECTIONS before the CUDA context exists)
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # stdout carries the JSON line: keep NCCL's version banner (printf to stdout at NCCL_DEBUG=VERSION or WARN) off
        # it unless somebody asked for a more verbose level on purpose
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION', 'WARN'):
            os.environ['NCCL_DEBUG'] = 'NONE'
        dist.init_process_group('nccl', device_id=dev)
    if rank == 0:
        url = ensure_dataset(n_groups, rows_pg)
    if world > 1:
        dist.barrier()
    url = 'file://' + DATA_DIR

    from petastorm_b200 import make_batch_reader, rowgroup, sharding
    from petastorm_b200.etl import dataset_metadata as dm
    rowgroup.set_pinned_cache_bytes(8 << 30)

    shard_kwargs = sharding.sharded_reader_kwargs(url)          # one NCCL broadcast of the owner table
    pieces = dm.load_row_groups(dm.ParquetDataset(DATA_DIR))
    mine = [i for i in range(len(pieces)) if world == 1 or i % world == rank]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) value: raw bytes resident in HBM, decode kernels only -------------------------------------------------
    dec = rowgroup.RowGroupDecoder(local_rank)
    leaves = list(range(N_F32 + N_I64))
    plans, arenas = [], []
    for i in mine:
        p = dec.plan(pieces[i].path, pieces[i].row_group, leaves)
        plans.append(p)
        arenas.append(dec.upload(p, private=True))
    torch.cuda.synchronize()
    payload = sum(p.info.payload_bytes for p in plans) / len(plans)   # encoded bytes E per row-group
    stream = dec.streams[0]

    def resident_step(k):
        j = k % len(plans)
        # consecutive row-groups on different streams, exactly like the readers (RowGroupDecoder.decode)
        d = dec.decode_resident(plans[j], arenas[j], dec.streams[k % len(dec.streams)])
        col = d.column(0).values
        # consumer side: 4096-row batches are views of the row-group tensors (no copy)
        nb = (col.numel() + BATCH - 1) // BATCH
        return d, nb

    for k in range(args.warmup):
        d, _ = resident_step(k)
    barrier()
    launches0 = dec.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keep = []
    barrier()
    e0.record(stream)
    for other in dec.streams[1:]:
        other.wait_event(e0)                     # every decode stream starts after the start event
    last = {}
    for k in range(args.steps):
        d, _ = resident_step(args.warmup + k)
        keep = [d]
        last[(args.warmup + k) % len(dec.streams)] = d.event
    for ev in last.values():
        stream.wait_event(ev)                    # the stop event waits for the last row-group of every stream
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    sampler.active = False
    keep[0].check()
    gpu_launches = dec.launches - launches0
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * args.steps * rows_pg / (dev_ms / 1e3)

    # ---- (2) roofline of the dominant kernel, per-kernel CUDA events on the launching stream ---------------------
    from ctypes import c_float
    from petastorm_b200 import native
    ms_acc = np.zeros(6)
    reps = max(4, min(args.steps, 8))
    for k in range(reps):
        j = k % len(plans)
        out = torch.empty(plans[j].info.out_bytes, dtype=torch.uint8, device=dev)
        status = torch.zeros(8 + len(leaves), dtype=torch.int32, device=dev)
        ms6 = (c_float * 6)()
        native.check(native.lib.pst_plan_decode_timed(dec.ctx.handle, plans[j].handle, arenas[j].data_ptr(),
                                                      out.data_ptr(), status.data_ptr(), stream.cuda_stream, ms6))
        ms_acc += np.array(list(ms6))
    ms_avg = ms_acc / reps
    names = ['k_snappy_index', 'k_snappy_pages', 'k_snappy_pages(serial fallback)', 'k_ba_dict_index', 'k_copy_tiles',
             'k_decode_pages']
    dom = int(np.argmax(ms_avg))
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    # algorithmic bytes per launch (one launch = one row-group), from the plan's own page table
    algo = plan_algorithmic_bytes(plans[0])
    algo_by_kernel = [algo[n] for n in names]
    algo_bytes = algo_by_kernel[dom]
    achieved = algo_bytes / (ms_avg[dom] / 1e3) / 1e9
    decode_ms = float(ms_avg.sum())
    per_kernel = {n: {'ms': float(m), 'algorithmic_bytes': int(a), 'gbps': (a / (m / 1e3) / 1e9) if m > 0 else None,
                      'frac': (a / (m / 1e3) / 1e9 / peak) if m > 0 else None}
                  for n, m, a in zip(names, ms_avg, algo_by_kernel)}
    roofline = {'bound': 'hbm', 'kernel': names[dom], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': NCU_DRAM_BYTES_PER_LAUNCH.get(names[dom]), 'peak_source': peak_src,
                'traffic_source': NCU_DRAM_SOURCE,
                'algorithmic_bytes_per_launch': algo_bytes,
                'kernel_ms': {n: float(m) for n, m in zip(names, ms_avg)},
                'per_kernel': per_kernel,
                'minimal_bytes_E_plus_D': int(payload + rows_pg * ROW_BYTES),
                'whole_decode_ms_serialised': decode_ms,
                'whole_decode_frac': (payload + rows_pg * ROW_BYTES) / (decode_ms / 1e3) / 1e9 / peak}
    del arenas, plans, keep, d
    torch.cuda.empty_cache()

    # ---- (3) e2e through the public API from host buffers ----------------------------------------------------------
    total = args.warmup + args.steps
    epochs = (total + len(mine) - 1) // len(mine) + 1
    host_buf = torch.empty(BATCH, dtype=torch.int64).pin_memory()
    reader = make_batch_reader(url, shuffle_row_groups=False, num_epochs=epochs, device=local_rank, **shard_kwargs)
    it = iter(reader)
    h2d0 = d2h = 0
    for k in range(args.warmup):
        b = next(it)
        host_buf.copy_(b.i00[:BATCH], non_blocking=True)
    barrier()
    h2d0 = reader.diagnostics['h2d_bytes']
    consumed = []
    if os.environ.get('PST_TRACE'):
        rowgroup.TRACE = []
        consumed.append(time.perf_counter())
    rows_e2e = 0
    sampler.active = True
    t0 = time.perf_counter()
    for k in range(args.steps):
        b = next(it)                              # one decoded row-group (namedtuple of CUDA tensors)
        n = b.f00.shape[0]
        for s in range(0, n, BATCH):              # consumer re-batches to 4096 (views)
            _ = b.f00[s:s + BATCH]
        host_buf.copy_(b.i00[:BATCH], non_blocking=True)   # device -> host read of the step's result
        torch.cuda.current_stream().synchronize()
        if rowgroup.TRACE is not None:
            consumed.append(time.perf_counter())
        d2h += host_buf.numel() * 8
        rows_e2e += n
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()          # sampled over both timed regions (HBM-resident decode and end-to-end reader)
    diag = reader.diagnostics
    h2d = diag['h2d_bytes'] - h2d0
    reader.stop()
    reader.join()
    if rowgroup.TRACE:
        tr = rowgroup.TRACE
        torch.cuda.synchronize()
        h0, e0 = tr[0]
        for h, ev in tr:
            sys.stderr.write('issue %8.2f  h2d %8.2f .. %8.2f  decode .. %8.2f\n' % (
                (h - h0) * 1e3, e0[0].elapsed_time(ev[0]), e0[0].elapsed_time(ev[1]), e0[0].elapsed_time(ev[2])))
        sys.stderr.write('consumed at: ' + ' '.join('%.1f' % ((c - h0) * 1e3) for c in consumed) + '\n')
    tw = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw.item())
    e2e_value = world * rows_e2e / wall

    # ---- (4) CPU baseline on this box's cores (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        best, variants, sample = cpu_reference_best(url, 16, 8, cores, cpu_workers)
        # thread pools: the workers call Arrow C++ with use_threads=True, so its own pool (all host cores) decodes
        cpu = {'value': best['samples_per_sec'], 'unit': 'samples/s',
               'cores': cores if best['pool'] == 'thread' else best['workers'], 'kind': 'port',
               'sample': sample, 'variants': variants, 'host_cores': cores}

    if rank == 0:
        line = {'metric': 'samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': config, 'delivered_gbps': value * ROW_BYTES / 1e9, 'roofline': roofline,
                'cpu_baseline': cpu,
                'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d // max(args.steps, 1),
                        'd2h_bytes_per_step': d2h // max(args.steps, 1), 'delivered_gbps': e2e_value * ROW_BYTES / 1e9,
                        'ms_per_step': 1e3 * wall / args.steps,
                        'h2d_gbps': world * h2d / wall / 1e9, 'pinned_cache_hits': diag.get('pinned_cache_hits'),
                        'host_seconds_total': diag.get('host_seconds')},
                'gpu_launches': gpu_launches, 'clocks': clocks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
