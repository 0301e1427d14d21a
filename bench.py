#!/usr/bin/env python
"""bench.py -- benchmarks of the Parquet -> GPU tensor read hot path on the five BASELINE.json configurations.

  --workload c2 (default, the headline: BASELINE.json configs[1])
        make_batch_reader over plain Parquet, 64 float32 + 16 int64 flat columns, Snappy, pyarrow defaults (v1 pages,
        dictionary-then-PLAIN fallback, nullable columns), ~256 MB row-groups (666,667 rows), consumer re-batches to 4096
  --workload c1   HelloWorldSchema (id int32 + 128x256x3 png + uint8 (4,128,30,3) NdarrayCodec), make_reader
  --workload c3   ImageNet shape: 224x224x3 jpeg (quality 80) + int32 label, make_reader(shuffle_row_groups) +
                  DataLoader(batch 256)
  --workload c4   NdarrayCodec float16 (32,128,128) tensor + int32 key, TransformSpec normalise + in_set predicate on key
  --workload c5   NGram (length 16) over a 12 x float32 time series, DataLoader(shuffling_queue_capacity=100000)

A stated subset of each nominal dataset is materialised (`--row-groups`) and cycled with num_epochs.  One *step* = one
row-group through the hot path.  Every line carries

  value        samples/s with the raw column-chunk bytes already resident in HBM when the timed region starts
               (c2: plan -> decode kernels per row-group on the decode streams, CUDA events; c1/c3/c4/c5: the public
               reader + loader pipeline over the decoder's HBM-resident raw row-group cache, CUDA events)
  e2e          samples/s through the public API from HOST buffers: pinned host -> H2D -> decode -> D2H of a small result
               per step; `e2e.cold` repeats it with the pinned row-group cache and the plan cache disabled
               (page cache -> pinned staging -> H2D every step)
  roofline     HBM roofline of the dominant kernel, timed live with CUDA events on its launching stream
  cpu_baseline the oracle's restatement of the reference reader (oracle/port.py, kind "port": ParquetFile.read_row_group /
               cv2 / numpy on a thread or process pool - NOT the reference's ZeroMQ ProcessPool) on this box's cores,
               bounded sample, rank 0 / N=1 only
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

BENCH_DIR = os.environ.get('PST_BENCH_DIR', '/tmp/pst_bench')


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
        self.active = False

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


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return json.load(open(path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ncu --set full DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernels, from the
# committed capture of this round
NCU_DRAM_SOURCE = ('profiles/r2end_c2_decode_kernels.txt (ncu --set full, one launch of each kernel on a C2 row-group)')
NCU_DRAM_BYTES_PER_LAUNCH = {'k_snappy_index': 76608000 + 2254848, 'k_snappy_pages': 81706752 + 39542528,
                             'k_copy_tiles': 172147968 + 121758464, 'k_decode_pages': 127348480 + 54201856}


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


# ---------------------------------------------------------------------------------------------------------------------
# synthetic datasets (np.random.default_rng(1234 + part)); one file = one row-group
# ---------------------------------------------------------------------------------------------------------------------
C2_ROWS, C2_F32, C2_I64, C2_BATCH = 666667, 64, 16, 4096
C2_ROW_BYTES = C2_F32 * 4 + C2_I64 * 8


def _gen_c2(args):
    path, seed, rows = args
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    rng = np.random.default_rng(seed)
    cols = {}
    for i in range(C2_F32):
        cols['f%02d' % i] = rng.standard_normal(rows, dtype=np.float32)
    for i in range(C2_I64):
        cols['i%02d' % i] = rng.integers(0, 2 ** 40, rows, dtype=np.int64)
    pq.write_table(pa.table(cols), path + '.tmp', compression='snappy', row_group_size=rows)
    os.replace(path + '.tmp', path)
    return 1


def _schemas():
    import numpy as np
    from petastorm_b200 import spark_types as T
    from petastorm_b200.codecs import CompressedImageCodec, NdarrayCodec, ScalarCodec
    from petastorm_b200.unischema import Unischema, UnischemaField
    return {
        'c1': Unischema('HelloWorldSchema', [
            UnischemaField('id', np.int32, (), ScalarCodec(T.IntegerType()), False),
            UnischemaField('image1', np.uint8, (128, 256, 3), CompressedImageCodec('png'), False),
            UnischemaField('array_4d', np.uint8, (4, 128, 30, 3), NdarrayCodec(), False)]),
        'c3': Unischema('ImagenetSchema', [
            UnischemaField('label', np.int32, (), ScalarCodec(T.IntegerType()), False),
            UnischemaField('image', np.uint8, (224, 224, 3), CompressedImageCodec('jpeg', 80), False)]),
        'c4': Unischema('TensorSchema', [
            UnischemaField('key', np.int32, (), ScalarCodec(T.IntegerType()), False),
            UnischemaField('tensor', np.float16, (32, 128, 128), NdarrayCodec(), False)]),
        'c5': Unischema('SeriesSchema', [UnischemaField('ts', np.int64, (), ScalarCodec(T.LongType()), False)] + [
            UnischemaField('c%02d' % i, np.float32, (), ScalarCodec(T.FloatType()), False) for i in range(12)]),
    }


def _gen_row_part(args):
    """One part file (= one row-group) of a make_reader workload; rows [first, first + rows)."""
    kind, path, seed, first, rows = args
    import io
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    sys.path.insert(0, ROOT)
    from petastorm_b200.etl.dataset_writer import arrow_schema_of
    rng = np.random.default_rng(seed)
    schema = _schemas()[kind]
    if kind == 'c1':
        import cv2
        ids, imgs, arrs = [], [], []
        for i in range(rows):
            img = rng.integers(0, 255, (128, 256, 3), dtype=np.uint8)
            _, enc = cv2.imencode('.png', img[:, :, (2, 1, 0)])
            buf = io.BytesIO()
            np.save(buf, rng.integers(0, 255, (4, 128, 30, 3), dtype=np.uint8))
            ids.append(first + i)
            imgs.append(enc.tobytes())
            arrs.append(buf.getvalue())
        cols = {'id': np.asarray(ids, dtype=np.int32), 'image1': imgs, 'array_4d': arrs}
    elif kind == 'c3':
        import cv2
        yy, xx = np.mgrid[0:224, 0:224].astype(np.float32)
        labels, imgs = [], []
        for i in range(rows):
            a, b, c = rng.uniform(0.5, 3, 3)
            img = np.stack([127 + 120 * np.sin(xx / (20 * a) + i), 127 + 120 * np.cos(yy / (25 * b)),
                            127 + 100 * np.sin((xx + yy) / (30 * c))], -1)
            img = cv2.GaussianBlur(img.astype(np.float32), (0, 0), 2) + rng.normal(0, 3, img.shape)
            img = np.clip(img, 0, 255).astype(np.uint8)
            _, enc = cv2.imencode('.jpeg', img[:, :, (2, 1, 0)], [int(cv2.IMWRITE_JPEG_QUALITY), 80])
            labels.append(int(rng.integers(0, 1000)))
            imgs.append(enc.tobytes())
        cols = {'label': np.asarray(labels, dtype=np.int32), 'image': imgs}
    elif kind == 'c4':
        blobs = []
        for i in range(rows):
            buf = io.BytesIO()
            np.save(buf, rng.standard_normal((32, 128, 128), dtype=np.float32).astype(np.float16))
            blobs.append(buf.getvalue())
        cols = {'key': np.arange(first, first + rows, dtype=np.int32), 'tensor': blobs}
    else:  # c5: monotone timestamps with a gap > delta_threshold every 1000 rows
        idx = np.arange(first, first + rows, dtype=np.int64)
        cols = {'ts': idx + 5 * (idx // 1000)}
        for c in range(12):
            cols['c%02d' % c] = rng.standard_normal(rows, dtype=np.float32)
    table = pa.Table.from_pydict({n: cols[n] for n in schema.fields.keys()}, schema=arrow_schema_of(schema))
    pq.write_table(table, path + '.tmp', compression='snappy', row_group_size=rows)
    os.replace(path + '.tmp', path)
    return 1


def ensure_dataset(kind, n_groups, rows):
    """Materialise `n_groups` part files of `rows` rows under BENCH_DIR/<kind>; returns the file:// url."""
    import multiprocessing as mp
    d = os.path.join(BENCH_DIR, kind)
    os.makedirs(d, exist_ok=True)
    names = ['part-%05d-r%d.parquet' % (g, rows) for g in range(n_groups)]
    for name in os.listdir(d):
        if name not in names and not name.startswith('_'):
            os.remove(os.path.join(d, name))
    todo = [g for g in range(n_groups) if not os.path.exists(os.path.join(d, names[g]))]
    if todo:
        if kind == 'c2':
            jobs = [(os.path.join(d, names[g]), 1234 + g, rows) for g in todo]
            fn = _gen_c2
        else:
            jobs = [(kind, os.path.join(d, names[g]), 1234 + g, g * rows, rows) for g in todo]
            fn = _gen_row_part
        with mp.get_context('spawn').Pool(min(len(jobs), 16)) as pool:
            pool.map(fn, jobs)
    if kind != 'c2':
        from petastorm_b200.etl.dataset_writer import write_common_metadata
        write_common_metadata(d, _schemas()[kind], {n: 1 for n in names})
    return 'file://' + d


# ---------------------------------------------------------------------------------------------------------------------
# workload descriptions
# ---------------------------------------------------------------------------------------------------------------------
class Workload(object):
    key = None
    rows_per_group = 0
    decoded_row_bytes = 0        # D per delivered sample
    value_steps_min = 0          # row-groups of the HBM-resident leg: at least this many, for a timed region of >= 0.25 s
    batch = 1
    delivered_fraction = 1.0     # delivered samples / stored rows (predicate, NGram window yield)

    def __init__(self, rows_per_group=None):
        if rows_per_group:
            self.rows_per_group = rows_per_group

    def describe(self, n_groups, world):
        raise NotImplementedError

    # make_reader / loader of the public API ---------------------------------------------------------------------
    def reader(self, url, epochs, device, shard_kwargs):
        raise NotImplementedError

    def loader(self, reader):
        raise NotImplementedError

    @staticmethod
    def batch_rows(batch):
        windows = getattr(batch, 'windows', None)      # NGram batch of the device loaders: {field: tensor[batch, L, ...]}
        if windows is not None:
            return int(next(iter(windows.values())).shape[0])
        first = batch[next(iter(batch))]
        if isinstance(first, dict):
            first = first[next(iter(first))]
        return int(first.shape[0])

    def d2h_source(self, batch):
        raise NotImplementedError

    # reference arm ----------------------------------------------------------------------------------------------
    def reference_batches(self, url, workers, pool, epochs):
        raise NotImplementedError


def _specs(kind):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import oracle_specs
    return oracle_specs(_schemas()[kind])


class C1(Workload):
    key, rows_per_group, batch = 'c1', 256, 64
    value_steps_min = 128
    decoded_row_bytes = 4 + 128 * 256 * 3 + 4 * 128 * 30 * 3

    def describe(self, n_groups, world):
        return ('C1 HelloWorldSchema make_reader: id int32 + image1 128x256x3 png (noise, cv2-written) + array_4d uint8 '
                '(4,128,30,3) NdarrayCodec, Snappy, {} rows/row-group, BatchedDataLoader batch {}; {} row-groups '
                'materialised ({} rows of the nominal 10k) and cycled').format(self.rows_per_group, self.batch, n_groups,
                                                                                n_groups * self.rows_per_group)

    def reader(self, url, epochs, device, shard_kwargs):
        from petastorm_b200 import make_reader
        return make_reader(url, shuffle_row_groups=False, num_epochs=epochs, device=device, **shard_kwargs)

    def loader(self, reader):
        from petastorm_b200.pytorch import BatchedDataLoader
        return BatchedDataLoader(reader, batch_size=self.batch)

    def d2h_source(self, batch):
        return batch['id']

    def reference_batches(self, url, workers, pool, epochs):
        from oracle import port
        return port.reference_dataloader(port.pool_rows(url, _specs('c1'), workers, pool, epochs), self.batch)


class C3(Workload):
    key, rows_per_group, batch = 'c3', 1024, 256
    value_steps_min = 48
    decoded_row_bytes = 224 * 224 * 3 + 4

    def describe(self, n_groups, world):
        return ('C3 ImageNet shape make_reader(shuffle_row_groups=True) + DataLoader(batch {}): image 224x224x3 '
                'CompressedImageCodec(jpeg, q80) + int32 label, {} rows/row-group; {} row-groups materialised ({} '
                'images of the nominal 1.28 M) and cycled').format(self.batch, self.rows_per_group, n_groups,
                                                                    n_groups * self.rows_per_group)

    def reader(self, url, epochs, device, shard_kwargs):
        from petastorm_b200 import make_reader
        return make_reader(url, shuffle_row_groups=True, seed=17, num_epochs=epochs, device=device, **shard_kwargs)

    def loader(self, reader):
        from petastorm_b200.pytorch import DataLoader
        return DataLoader(reader, batch_size=self.batch)

    def d2h_source(self, batch):
        return batch['label']

    def reference_batches(self, url, workers, pool, epochs):
        from oracle import port
        rows = port.pool_rows(url, _specs('c3'), workers, pool, epochs, shuffle_row_groups=True, seed=17)
        return port.reference_dataloader(rows, self.batch)


class C4(Workload):
    key, rows_per_group, batch = 'c4', 128, 32
    value_steps_min = 192
    decoded_row_bytes = 32 * 128 * 128 * 2 + 4
    delivered_fraction = 0.5
    MEAN, STD = 0.25, 1.5

    def describe(self, n_groups, world):
        return ('C4 make_reader: key int32 + tensor float16 (32,128,128) NdarrayCodec (1 MiB/row), TransformSpec '
                'normalise ((x - {}) / {} in fp32 -> float16) + predicate in_set(even keys) (50 % of the rows delivered), '
                '{} rows/row-group, BatchedDataLoader batch {}; {} row-groups materialised ({} rows of the nominal 2 M) '
                'and cycled').format(self.MEAN, self.STD, self.rows_per_group, self.batch, n_groups,
                                     n_groups * self.rows_per_group)

    def reader(self, url, epochs, device, shard_kwargs):
        from petastorm_b200 import make_reader
        from petastorm_b200.predicates import in_set
        from petastorm_b200.transform import Normalize, TransformSpec
        return make_reader(url, shuffle_row_groups=False, num_epochs=epochs, device=device,
                           predicate=in_set(range(0, 1 << 22, 2), 'key'),
                           transform_spec=TransformSpec(Normalize('tensor', self.MEAN, self.STD, 'float16')),
                           **shard_kwargs)

    def loader(self, reader):
        from petastorm_b200.pytorch import BatchedDataLoader
        return BatchedDataLoader(reader, batch_size=self.batch)

    def d2h_source(self, batch):
        return batch['key']

    def reference_batches(self, url, workers, pool, epochs):
        from oracle import port
        rows = port.pool_rows(url, _specs('c4'), workers, pool, epochs, predicate=port.InSet(range(0, 1 << 22, 2), 'key'),
                              transform_func=port.NormalizeRow('tensor', self.MEAN, self.STD, 'float16'))
        return port.reference_dataloader(rows, self.batch)


class C5(Workload):
    key, rows_per_group, batch = 'c5', 131072, 1024
    value_steps_min = 16
    LENGTH, CAPACITY = 16, 100000
    decoded_row_bytes = 16 * (12 * 4 + 8)
    delivered_fraction = (1000 - 15) / 1000.0

    def describe(self, n_groups, world):
        return ('C5 make_reader(schema_fields=NGram) + DataLoader(batch {}, shuffling_queue_capacity={}): ts int64 '
                '(gap > delta_threshold every 1000 rows) + 12 x float32, NGram length {} with all 13 fields at every '
                'timestep, delta_threshold 1, {} rows/row-group; {} row-groups materialised ({} rows of the nominal '
                '50 M) and cycled').format(self.batch, self.CAPACITY, self.LENGTH, self.rows_per_group, n_groups,
                                           n_groups * self.rows_per_group)

    def ngram(self):
        from petastorm_b200.ngram import NGram
        schema = _schemas()['c5']
        fields = {k: list(schema.fields.values()) for k in range(self.LENGTH)}
        return NGram(fields, delta_threshold=1, timestamp_field=schema.fields['ts'])

    def reader(self, url, epochs, device, shard_kwargs):
        from petastorm_b200 import make_reader
        return make_reader(url, schema_fields=self.ngram(), shuffle_row_groups=False, num_epochs=epochs, device=device,
                           **shard_kwargs)

    def loader(self, reader):
        from petastorm_b200.pytorch import DataLoader
        return DataLoader(reader, batch_size=self.batch, shuffling_queue_capacity=self.CAPACITY)

    def d2h_source(self, batch):
        return batch[0]['ts']

    def reference_batches(self, url, workers, pool, epochs):
        from oracle import port
        names = list(_schemas()['c5'].fields.keys())
        ngram = dict(fields={k: names for k in range(self.LENGTH)}, ts='ts', delta=1)
        rows = port.pool_rows(url, _specs('c5'), workers, pool, epochs, ngram=ngram)
        return port.reference_dataloader(rows, self.batch, self.CAPACITY)


ROW_WORKLOADS = {'c1': C1, 'c3': C3, 'c4': C4, 'c5': C5}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (oracle/port.py)
# ---------------------------------------------------------------------------------------------------------------------
def _arrow_threads(cores):
    """torchrun exports OMP_NUM_THREADS=1, which also caps Arrow's CPU pool: the reference arm gets the host's cores
    back so that it runs the same at N=1 and N>1."""
    import pyarrow as pa
    os.environ.pop('OMP_NUM_THREADS', None)
    pa.set_cpu_count(max(1, cores))
    pa.set_io_thread_count(max(8, min(cores, 64)))
    return pa.cpu_count()


def c2_cpu_rate(url, steps, warmup, workers, pool):
    """rows/s of the CPU batch reader (oracle/port.py process_pool_batches / thread_pool_batches) over `steps`
    row-groups after `warmup` row-groups (child start-up excluded, like petastorm/benchmark/throughput.py:68-90)."""
    from oracle import port
    total = warmup + steps
    n_groups = len(port.list_pieces(url))
    epochs = (total + n_groups - 1) // n_groups
    make = port.process_pool_batches if pool == 'process' else port.thread_pool_batches
    gen = make(url, workers, epochs=epochs)
    rows, t0 = 0, None
    for k, batch in enumerate(gen):
        if k == warmup:
            t0 = time.perf_counter()
        if k >= warmup:
            n = len(batch['f00'])
            for s in range(0, n, C2_BATCH):      # consumer side: re-batch to 4096 (views)
                _ = batch['f00'][s:s + C2_BATCH]
            rows += n
        if k + 1 >= total:
            break
    dt = time.perf_counter() - t0
    gen.close()
    return rows / dt, rows, dt


def c2_cpu_best(url, steps, warmup, cores, process_workers):
    import pyarrow
    variants = []
    for pool, workers in (('thread', 10), ('thread', max(1, min(cores // 2, 64))), ('process', process_workers)):
        n = max(steps, workers if pool == 'process' else 96)
        v, rows, dt = c2_cpu_rate(url, n, max(warmup, workers if pool == 'process' else 8), workers, pool)
        variants.append({'pool': pool, 'workers': workers, 'samples_per_sec': v, 'rows': rows, 'seconds': round(dt, 2)})
    best = max(variants, key=lambda x: x['samples_per_sec'])
    sample = ('%d rows in %.1f s; oracle/port.py restatement (kind "port") of the reference batch reader on a %s pool '
              'with %d workers (ParquetFile.read_row_group + take%s), NOT the reference\'s ZeroMQ ProcessPool + '
              'ventilator; pyarrow %s with %d Arrow CPU threads; best of thread x10 (reference default), thread x '
              'cores/2, process x%d' %
              (best['rows'], best['seconds'], best['pool'], best['workers'],
               ' + Arrow-IPC back to the consumer' if best['pool'] == 'process' else ', tables handed over in-process',
               pyarrow.__version__, pyarrow.cpu_count(), process_workers))
    return best, variants, sample


def row_cpu_best(w, url, cores, budget_groups, budget_seconds=15.0):
    """The reference row reader (+ DataLoader) on a thread pool x10 (reference default) and a process pool, each on a
    bounded sample (`budget_groups` row-groups or `budget_seconds`, whichever comes first); the faster one is the
    baseline.  C5 skips the thread pool: NGram formation is pure Python (GIL-bound), ten row-groups in flight on ten
    threads deliver their first window only after minutes."""
    variants = []
    pools = (('thread', 10), ('process', max(2, min(cores // 2, 32))))
    if w.key == 'c5':
        pools = pools[1:]
    for pool, workers in pools:
        rows = 0
        first_t = None
        arrivals = []          # (time, rows) of every counted batch
        gen = w.reference_batches(url, workers, pool, 2)
        limit = budget_groups * w.rows_per_group * w.delivered_fraction
        for batch in gen:
            n = Workload.batch_rows(batch)
            now = time.perf_counter()
            if first_t is None:
                first_t = now       # pool start-up + first batch excluded like the warm-up
                continue
            rows += n
            arrivals.append((now, n))
            if rows >= limit or now - first_t > budget_seconds:
                break
        dt = time.perf_counter() - first_t
        gen.close()
        # Row-groups of 131 k rows (C5) reach the consumer in bursts - a worker needs tens of seconds of pure Python per
        # row-group - so a short sample can consist of one burst behind a long wait.  The baseline is the better of the
        # plain rate and the rate inside the longest run of batches without a gap of more than 2 s: an upper bound of
        # what the CPU path sustains, i.e. the conservative choice for the ratio.
        rate, used_rows, used_dt = rows / max(dt, 1e-9), rows, dt
        run_start, run_rows, prev_t = first_t, 0, first_t
        for t, n in arrivals + [(float('inf'), 0)]:
            if t - prev_t > 2.0:
                span = prev_t - run_start
                if run_rows >= 2048 and span > 0.5 and run_rows / span > rate:
                    rate, used_rows, used_dt = run_rows / span, run_rows, span
                run_start, run_rows = t, 0
            else:
                run_rows += n
            prev_t = t
        variants.append({'pool': pool, 'workers': workers, 'samples_per_sec': rate, 'rows': used_rows,
                         'seconds': round(used_dt, 2), 'sample_rows': rows, 'sample_seconds': round(dt, 2)})
    best = max(variants, key=lambda x: x['samples_per_sec'])
    sample = ('%d samples in %.1f s (of a sample of %d in %.1f s: the longest run of batches without a gap of more than 2 s '
              'when that is faster); oracle/port.py restatement (kind "port") of PyDictReaderWorker + '
              'petastorm.pytorch.DataLoader on a %s pool with %d workers (pq.read_row_group -> to_pandas -> codec decode '
              'per row -> row loop + default_collate), NOT the reference\'s ZeroMQ ProcessPool; best of thread x10 '
              '(reference default) and a process pool' % (best['rows'], best['seconds'], best['sample_rows'],
                                                          best['sample_seconds'], best['pool'], best['workers']))
    return best, variants, sample


# ---------------------------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=64)
    ap.add_argument('--warmup', type=int, default=8)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--row-groups', type=int, default=8, help='row-groups materialised per GPU box (cycled)')
    ap.add_argument('--rows-per-group', type=int, default=0)
    ap.add_argument('--cpu-workers', type=int, default=0, help='processes of the CPU reference arm (0 = all cores, max 64)')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--skip-cold', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    return args


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cores = os.cpu_count() or 1
    if args.workload == 'c2':
        fn, fargs = run_c2, (args, rank, local_rank, world, cores)
    else:
        fn, fargs = run_rows, (args, ROW_WORKLOADS[args.workload](args.rows_per_group or None), rank, local_rank, world, cores)
    prof_path = os.environ.get('PST_BENCH_PROFILE')      # diagnostics: cProfile of the consumer (main) thread
    if prof_path:
        import cProfile
        prof = cProfile.Profile()
        try:
            return prof.runcall(fn, *fargs)
        finally:
            prof.dump_stats(prof_path)
    return fn(*fargs)


def _dist_setup(local_rank, world):
    import petastorm_b200  # noqa: F401  (first: sets CUDA_DEVICE_MAX_CONNECTIONS before the CUDA context exists)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # stdout carries the JSON line: keep NCCL's version banner off it unless a verbose level was asked for
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION', 'WARN'):
            os.environ['NCCL_DEBUG'] = 'NONE'
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    return torch, dist, dev, barrier, max_over_ranks, sum_over_ranks


# =====================================================================================================================
# C2: make_batch_reader
# =====================================================================================================================
def run_c2(args, rank, local_rank, world, cores):
    rows_pg = args.rows_per_group or C2_ROWS
    n_groups = max(args.row_groups, 2 * world)    # every rank cycles over at least two distinct row-groups
    cpu_workers = args.cpu_workers or min(cores, 64)
    workload = ('C2 make_batch_reader: {}xfloat32 + {}xint64 flat columns, Snappy, {} rows/row-group (~{} MB decoded), '
                'batch {}; {} row-groups materialised ({} rows of the nominal 100 M) and cycled').format(
                    C2_F32, C2_I64, rows_pg, rows_pg * C2_ROW_BYTES // 2 ** 20, C2_BATCH, n_groups, n_groups * rows_pg)
    config = {'workload': workload, 'rows_per_row_group': rows_pg, 'row_groups_materialised': n_groups,
              'batch': C2_BATCH, 'compression': 'snappy', 'parallelism': 'row-group shards, %d rank(s)' % world,
              'l2_policy': 'inputs larger than L2: each step reads a distinct ~%d MB arena' % (rows_pg * C2_ROW_BYTES // 2 ** 20)}

    if args.impl == 'reference':
        if rank != 0:
            return
        url = ensure_dataset('c2', n_groups, rows_pg)
        arrow_threads = _arrow_threads(cores)
        best, variants, sample = c2_cpu_best(url, min(args.steps, 64), args.warmup, cores, cpu_workers)
        value = best['samples_per_sec']
        line = {'impl': 'reference', 'metric': 'samples_per_sec', 'value': value, 'unit': 'samples/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * rows_pg / value,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': config, 'delivered_gbps': value * C2_ROW_BYTES / 1e9,
                'cpu_baseline': {'value': value, 'unit': 'samples/s',
                                 'cores': cores if best['pool'] == 'thread' else best['workers'], 'kind': 'port',
                                 'sample': sample, 'variants': variants, 'host_cores': cores,
                                 'arrow_cpu_threads': arrow_threads},
                'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    torch, dist, dev, barrier, max_over_ranks, sum_over_ranks = _dist_setup(local_rank, world)
    import numpy as np
    if rank == 0:
        ensure_dataset('c2', n_groups, rows_pg)
    barrier()
    data_dir = os.path.join(BENCH_DIR, 'c2')
    url = 'file://' + data_dir

    from petastorm_b200 import make_batch_reader, rowgroup, sharding
    from petastorm_b200.etl import dataset_metadata as dm
    rowgroup.set_pinned_cache_bytes(8 << 30)
    shard_kwargs = sharding.sharded_reader_kwargs(url)          # one NCCL broadcast of the owner table
    pieces = dm.load_row_groups(dm.ParquetDataset(data_dir))
    mine = [i for i in range(len(pieces)) if world == 1 or i % world == rank]

    # ---- (1) value: raw bytes resident in HBM, decode kernels only -------------------------------------------------
    dec = rowgroup.RowGroupDecoder(local_rank)
    leaves = list(range(C2_F32 + C2_I64))
    plans, arenas = [], []
    for i in mine:
        p = dec.plan(pieces[i].path, pieces[i].row_group, leaves)
        plans.append(p)
        arenas.append(dec.upload(p, private=True))
    torch.cuda.synchronize()
    payload = sum(p.info.payload_bytes for p in plans) / len(plans)   # encoded bytes E per row-group
    payload_raw = plans[0].info.raw_bytes
    stream = dec.streams[0]
    # a timed region of >= 0.25 s: the step count of the HBM-resident leg is scaled up (a step stays one row-group)
    value_steps = max(args.steps, 192)

    def resident_step(k):
        j = k % len(plans)
        # consecutive row-groups on different streams, exactly like the readers (RowGroupDecoder.decode)
        return dec.decode_resident(plans[j], arenas[j], dec.streams[k % len(dec.streams)])

    for k in range(args.warmup):
        d = resident_step(k)
    barrier()
    launches0 = dec.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.active = True
    e0.record(stream)
    for other in dec.streams[1:]:
        other.wait_event(e0)                     # every decode stream starts after the start event
    last, keep = {}, None
    for k in range(value_steps):
        d = resident_step(args.warmup + k)
        col = d.column(0).values
        nb = (col.numel() + C2_BATCH - 1) // C2_BATCH       # consumer: 4096-row batches are views (no copy)
        keep = (d, nb)
        last[(args.warmup + k) % len(dec.streams)] = d.event
    for ev in last.values():
        stream.wait_event(ev)                    # the stop event waits for the last row-group of every stream
    e1.record(stream)
    barrier()
    sampler.active = False
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    keep[0].check()
    gpu_launches = (dec.launches - launches0) * args.steps // value_steps
    value = world * value_steps * rows_pg / (dev_ms / 1e3)

    # ---- (2) roofline of the dominant kernel, per-kernel CUDA events on the launching stream ---------------------
    from ctypes import c_float
    from petastorm_b200 import native
    ms_acc = np.zeros(6)
    reps = 8
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
    peak, peak_src = hbm_peak()
    algo = plan_algorithmic_bytes(plans[0])
    algo_by_kernel = [algo[n] for n in names]
    achieved = algo_by_kernel[dom] / (ms_avg[dom] / 1e3) / 1e9
    decode_ms = float(ms_avg.sum())
    per_kernel = {n: {'ms': float(m), 'algorithmic_bytes': int(a), 'gbps': (a / (m / 1e3) / 1e9) if m > 0 else None,
                      'frac': (a / (m / 1e3) / 1e9 / peak) if m > 0 else None}
                  for n, m, a in zip(names, ms_avg, algo_by_kernel)}
    min_bytes = payload + rows_pg * C2_ROW_BYTES
    roofline = {'bound': 'hbm', 'kernel': names[dom], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': NCU_DRAM_BYTES_PER_LAUNCH.get(names[dom]), 'peak_source': peak_src,
                'traffic_source': NCU_DRAM_SOURCE, 'algorithmic_bytes_per_launch': int(algo_by_kernel[dom]),
                'kernel_ms': {n: float(m) for n, m in zip(names, ms_avg)}, 'per_kernel': per_kernel,
                'minimal_bytes_E_plus_D': int(min_bytes), 'whole_decode_ms_serialised': decode_ms,
                'whole_decode_frac_serialised': min_bytes / (decode_ms / 1e3) / 1e9 / peak,
                'whole_decode_frac_overlapped': min_bytes / (dev_ms / value_steps / 1e3) / 1e9 / peak}
    del arenas, plans, keep, d
    torch.cuda.empty_cache()

    # ---- (3) e2e through the public API from host buffers ----------------------------------------------------------
    def e2e_leg(cold):
        """`steps` row-groups through make_batch_reader; `cold` disables the pinned row-group cache and the plan cache:
        every step walks the page headers again and copies page cache -> pinned staging ring -> H2D."""
        if cold:
            rowgroup.set_pinned_cache_bytes(0)
            rowgroup.RowGroupDecoder.PLAN_CACHE_ENTRIES = 0
        steps = args.steps if not cold else min(args.steps, 32)
        inflight = 8
        total = args.warmup + steps + inflight      # the pipeline runs ahead: never let it drain inside the region
        epochs = (total + len(mine) - 1) // len(mine) + 1
        host_buf = torch.empty(C2_BATCH, dtype=torch.int64).pin_memory()
        reader = make_batch_reader(url, shuffle_row_groups=False, num_epochs=epochs, device=local_rank, **shard_kwargs)
        it = iter(reader)
        for k in range(args.warmup):
            b = next(it)
            host_buf.copy_(b.i00[:C2_BATCH], non_blocking=True)
        barrier()
        sampler.active = True
        diag0 = reader.diagnostics
        rows_e2e = d2h = 0
        t0 = time.perf_counter()
        for k in range(steps):
            b = next(it)                              # one decoded row-group (namedtuple of CUDA tensors)
            n = b.f00.shape[0]
            for s in range(0, n, C2_BATCH):           # consumer re-batches to 4096 (views)
                _ = b.f00[s:s + C2_BATCH]
            host_buf.copy_(b.i00[:C2_BATCH], non_blocking=True)   # device -> host read of the step's result
            torch.cuda.current_stream().synchronize()
            d2h += host_buf.numel() * 8
            rows_e2e += n
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        sampler.active = False
        diag = reader.diagnostics
        reader.stop()
        reader.join()
        barrier()
        wall = max_over_ranks(wall)
        v = world * rows_e2e / wall
        # the reader runs up to `inflight` row-groups ahead of the consumer, so the H2D bytes it ISSUED inside the window
        # need not equal steps x raw region; h2d_bytes_per_step is the raw region every step has to move
        return {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': int(payload_raw), 'd2h_bytes_per_step': d2h // steps,
                'delivered_gbps': v * C2_ROW_BYTES / 1e9, 'ms_per_step': 1e3 * wall / steps,
                'h2d_gbps': world * steps * payload_raw / wall / 1e9,
                'h2d_bytes_issued_in_window': int(diag['h2d_bytes'] - diag0['h2d_bytes']),
                'pinned_cache_hits': diag.get('pinned_cache_hits', 0) - diag0.get('pinned_cache_hits', 0),
                'pinned_cache_misses': diag.get('pinned_cache_misses', 0) - diag0.get('pinned_cache_misses', 0),
                'steps': steps, 'host_seconds_total': diag.get('host_seconds')}

    e2e = e2e_leg(False)
    e2e['note'] = ('steady state: the 8 GiB pinned row-group cache holds the materialised subset and plans are cached, '
                   'so a step is one cudaMemcpyAsync from pinned memory + decode; the CPU arm re-reads and re-parses '
                   'its files every step (page cache).  e2e.cold times the miss path.')
    if not args.skip_cold:
        e2e['cold'] = e2e_leg(True)
    clocks = sampler.stop()

    # ---- (4) CPU baseline on this box's cores (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        arrow_threads = _arrow_threads(cores)
        best, variants, sample = c2_cpu_best(url, 16, 8, cores, cpu_workers)
        cpu = {'value': best['samples_per_sec'], 'unit': 'samples/s',
               'cores': cores if best['pool'] == 'thread' else best['workers'], 'kind': 'port',
               'sample': sample, 'variants': variants, 'host_cores': cores, 'arrow_cpu_threads': arrow_threads}

    if rank == 0:
        line = {'metric': 'samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_ms / value_steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': config, 'delivered_gbps': value * C2_ROW_BYTES / 1e9, 'value_steps_timed': value_steps,
                'value_region_ms': dev_ms, 'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e,
                'gpu_launches': gpu_launches, 'clocks': clocks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# =====================================================================================================================
# C1 / C3 / C4 / C5: make_reader + loader
# =====================================================================================================================
def run_rows(args, w, rank, local_rank, world, cores):
    # four distinct row-groups per rank: with two, consecutive decodes of the HBM-resident leg wait for each other (the
    # same cached arena is decoded again two steps later) and the 8-GPU `value` of C3 came out below its `e2e`
    n_groups = max(args.row_groups, 4 * world)
    rows_pg = w.rows_per_group
    delivered_pg = rows_pg * w.delivered_fraction
    config = {'workload': w.describe(n_groups, world), 'rows_per_row_group': rows_pg,
              'row_groups_materialised': n_groups, 'batch': w.batch, 'compression': 'snappy',
              'parallelism': 'row-group shards, %d rank(s)' % world,
              'samples': 'delivered samples (after the predicate / valid NGram windows)',
              'l2_policy': 'inputs larger than L2 in aggregate: the materialised row-groups are cycled, each step '
                           'decodes a different one'}

    if args.impl == 'reference':
        if rank != 0:
            return
        url = ensure_dataset(w.key, n_groups, rows_pg)
        _arrow_threads(cores)
        best, variants, sample = row_cpu_best(w, url, cores, max(2, min(args.steps, 4)))
        value = best['samples_per_sec']
        line = {'impl': 'reference', 'metric': 'samples_per_sec', 'value': value, 'unit': 'samples/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * delivered_pg / value,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': config, 'delivered_gbps': value * w.decoded_row_bytes / 1e9,
                'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': best['workers'], 'kind': 'port',
                                 'sample': sample, 'variants': variants, 'host_cores': cores},
                'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    torch, dist, dev, barrier, max_over_ranks, sum_over_ranks = _dist_setup(local_rank, world)
    if rank == 0:
        ensure_dataset(w.key, n_groups, rows_pg)
    barrier()
    data_dir = os.path.join(BENCH_DIR, w.key)
    url = 'file://' + data_dir
    from petastorm_b200 import device_ops, rowgroup, sharding
    rowgroup.set_pinned_cache_bytes(8 << 30)
    shard_kwargs = sharding.sharded_reader_kwargs(url)          # one NCCL broadcast of the owner table
    groups_mine = n_groups // world if world > 1 else n_groups
    sampler = ClockSampler(local_rank)
    sampler.start()

    def run_leg(steps, warmup, resident):
        """`steps` row-groups worth of delivered samples through reader + loader.  Timed on the device (CUDA events on
        the consumer stream around the region) and on the host (wall clock between synchronisations)."""
        rowgroup.set_hbm_cache_bytes((64 << 30) if resident else 0)
        total = max(warmup, groups_mine + 2) + steps + 8
        epochs = (total + groups_mine - 1) // groups_mine + 2
        reader = w.reader(url, epochs, local_rank, shard_kwargs)
        loader = w.loader(reader)
        host_buf = torch.empty(w.batch, dtype=torch.int64).pin_memory()
        it = iter(loader)
        seen = 0
        target_warm = (warmup if not resident else max(warmup, groups_mine + 2)) * delivered_pg
        while seen < target_warm:             # resident: the first epoch fills the HBM cache (untimed)
            seen += w.batch_rows(next(it))
        barrier()
        diag0 = reader.diagnostics
        calls0 = device_ops.LAUNCH_CALLS[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active = True
        e0.record()
        t0 = time.perf_counter()
        rows = d2h = 0
        target = steps * delivered_pg
        while rows < target:
            b = next(it)
            n = w.batch_rows(b)
            src = w.d2h_source(b)
            host_buf[:n].copy_(src if src.dtype == torch.int64 else src.to(torch.int64), non_blocking=True)
            if not resident:
                torch.cuda.current_stream().synchronize()      # the step's result is read on the host
            d2h += n * 8
            rows += n
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        sampler.active = False
        dev_ms = e0.elapsed_time(e1)
        diag = reader.diagnostics
        reader.stop()
        reader.join()
        barrier()
        return {'rows': rows, 'wall': max_over_ranks(wall), 'dev_ms': max_over_ranks(dev_ms), 'd2h': d2h,
                'rows_all': sum_over_ranks(rows),
                'h2d': diag['h2d_bytes'] - diag0['h2d_bytes'],
                'launches': diag['gpu_launches'] - diag0['gpu_launches'] + device_ops.LAUNCH_CALLS[0] - calls0,
                'hbm_hits': diag.get('hbm_cache_hits', 0) - diag0.get('hbm_cache_hits', 0),
                'device_batched': getattr(loader, 'device_batched', None)}

    # ---- (1) value: raw bytes resident in HBM ------------------------------------------------------------------------
    value_steps = max(args.steps, w.value_steps_min)
    r = run_leg(value_steps, args.warmup, resident=True)
    value = r['rows_all'] / (r['dev_ms'] / 1e3)
    gpu_launches = r['launches'] * args.steps // value_steps
    value_info = {'steps_timed': value_steps, 'hbm_cache_hits': r['hbm_hits'], 'h2d_bytes_in_region': r['h2d'],
                  'device_batched_loader': r['device_batched'], 'region_ms': r['dev_ms']}

    # ---- (2) e2e from host buffers -------------------------------------------------------------------------------------
    r2 = run_leg(args.steps, args.warmup, resident=False)
    e2e_value = r2['rows_all'] / r2['wall']
    steps_done = max(r2['rows'] / delivered_pg, 1e-9)
    e2e = {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': int(r2['h2d'] / steps_done),
           'd2h_bytes_per_step': int(r2['d2h'] / steps_done), 'delivered_gbps': e2e_value * w.decoded_row_bytes / 1e9,
           'ms_per_step': 1e3 * r2['wall'] / steps_done, 'h2d_gbps': world * r2['h2d'] / r2['wall'] / 1e9,
           'device_batched_loader': r2['device_batched']}

    # ---- (3) roofline of the dominant kernel --------------------------------------------------------------------------
    roofline = row_roofline(w, url, local_rank, torch)
    clocks = sampler.stop()

    # ---- (4) CPU baseline ------------------------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        _arrow_threads(cores)
        best, variants, sample = row_cpu_best(w, url, cores, 2)
        cpu = {'value': best['samples_per_sec'], 'unit': 'samples/s', 'cores': best['workers'], 'kind': 'port',
               'sample': sample, 'variants': variants, 'host_cores': cores}

    if rank == 0:
        line = {'metric': 'samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['dev_ms'] / max(r['rows'] / delivered_pg, 1e-9),
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': config, 'delivered_gbps': value * w.decoded_row_bytes / 1e9, 'value_info': value_info,
                'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': gpu_launches, 'clocks': clocks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time_op(torch, fn, reps=6):
    """Average device milliseconds of `fn()` (CUDA events on the current stream, after two warm-up calls)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def row_roofline(w, url, device, torch):
    """Device time of the codec / post-processing kernels of one decoded row-group of the workload, each timed alone
    with CUDA events on its launching stream, with the algorithmic bytes of one launch; the slowest is the dominant one."""
    import numpy as np
    from petastorm_b200 import device_ops, make_reader
    from petastorm_b200.gpu_workers import ScalarColumn
    peak, peak_src = hbm_peak()
    reader = make_reader(url, reader_pool_type='dummy', shuffle_row_groups=False, num_epochs=1, device=device)
    worker = reader._workers_pool._worker                           # pylint: disable=protected-access
    piece = worker._split_pieces[0]                                  # pylint: disable=protected-access
    names = list(worker._schema.fields.keys())                       # pylint: disable=protected-access
    raw = worker._read_raw(piece, names)                             # pylint: disable=protected-access
    raw.decoded.check()
    raw.decoded.wait()
    n = raw.num_rows
    kernels = {}

    def col_of(name):
        return worker._leaf_of(raw, name)[1]                         # pylint: disable=protected-access

    def blob_bytes(col):
        return int(col.lens.sum().item())

    if w.key == 'c1':
        img, arr = col_of('image1'), col_of('array_4d')
        ms = _time_op(torch, lambda: device_ops.png_batch(img, 128, 256, 3, torch.uint8))
        kernels['k_png_batch'] = (ms, blob_bytes(img) + 3 * n * 128 * 256 * 3)   # blob + raw image written, read, RGB out
        payload = 4 * 128 * 30 * 3
        ms = _time_op(torch, lambda: device_ops.npy_batch(arr, 128, payload, torch.uint8, (4, 128, 30, 3)))
        kernels['k_npy_batch'] = (ms, 2 * n * payload)
    elif w.key == 'c3':
        img = col_of('image')
        field = worker._schema.fields['image']                       # pylint: disable=protected-access
        ms = _time_op(torch, lambda: worker._decode_jpeg(img, field, np.arange(n)), reps=3)   # pylint: disable=protected-access
        kernels['nvjpeg batched decode (library) incl. bitstream staging'] = (ms, blob_bytes(img) + n * 224 * 224 * 3)
    elif w.key == 'c4':
        t = col_of('tensor')
        payload = 32 * 128 * 128 * 2
        dense = device_ops.npy_batch(t, 128, payload, torch.float16, (32, 128, 128))[0]
        ms = _time_op(torch, lambda: device_ops.npy_batch(t, 128, payload, torch.float16, (32, 128, 128)))
        kernels['k_npy_batch'] = (ms, 2 * n * payload)
        ms = _time_op(torch, lambda: device_ops.normalize(dense, C4.MEAN, C4.STD, torch.float16))
        kernels['k_normalize_h8'] = (ms, 2 * n * payload)
    else:
        ng = w.ngram()
        cols = worker._decode_all(raw, names, None)                  # pylint: disable=protected-access
        ts = cols['ts'].tensor if isinstance(cols['ts'], ScalarColumn) else cols['ts']
        starts = ng.window_starts_device(ts.to(torch.int64))
        nw = int(starts.numel())
        f = cols['c00'].tensor if isinstance(cols['c00'], ScalarColumn) else cols['c00']
        ms = _time_op(torch, lambda: ng.window_starts_device(ts.to(torch.int64)))
        kernels['k_ngram_valid + compaction'] = (ms, n * 8 + n + nw * 8)
        ms = _time_op(torch, lambda: device_ops.ngram_gather(f.contiguous(), starts, ng.length))
        kernels['k_ngram_small (one of 13 fields)'] = (ms, n * 4 + nw * 8 + nw * ng.length * 4)
    # the Parquet page decode of the row-group itself (all plan kernels together)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec = worker._get_decoder()                                      # pylint: disable=protected-access
    plan = raw.decoded.plan
    arena = dec.upload(plan, private=True)
    torch.cuda.synchronize()
    d = dec.decode_resident(plan, arena, dec.streams[0])
    torch.cuda.synchronize()
    e0.record(dec.streams[0])
    for _ in range(4):
        d = dec.decode_resident(plan, arena, dec.streams[0])
    e1.record(dec.streams[0])
    torch.cuda.synchronize()
    kernels['parquet page decode (k_snappy_* + k_copy_tiles + k_decode_pages)'] = (
        e0.elapsed_time(e1) / 4, plan.info.payload_bytes + plan.info.uncompressed_bytes)
    del d
    reader.stop()
    reader.join()
    dom = max(kernels, key=lambda k: kernels[k][0])
    per = {k: {'ms': ms, 'algorithmic_bytes': int(b), 'gbps': b / (ms / 1e3) / 1e9, 'frac': b / (ms / 1e3) / 1e9 / peak}
           for k, (ms, b) in kernels.items()}
    return {'bound': 'hbm', 'kernel': dom, 'achieved': per[dom]['gbps'], 'peak': peak, 'unit': 'GB/s',
            'frac': per[dom]['frac'], 'traffic': None, 'peak_source': peak_src,
            'algorithmic_bytes_per_launch': per[dom]['algorithmic_bytes'], 'per_kernel': per,
            'rows_per_launch': int(n)}


if __name__ == '__main__':
    main()
