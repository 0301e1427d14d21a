"""GpuPool: the worker pool of the B200 path.

It speaks the reference's pool protocol (``start / ventilate / get_results / stop / join / workers_count /
diagnostics`` - canonical minimal implementation petastorm/workers_pool/dummy_pool.py:20-91, threaded one
thread_pool.py:109-263) but there is no process boundary and no serialisation: one host thread per pool issues the
row-group work (plan, staging copy, ``cudaMemcpyAsync``, kernel launches) on a side CUDA stream while the consumer
thread drains finished row-groups, so PCIe transfer, device decode and the consumer overlap.  Results are published in
ventilation order (single issuing thread), which makes seeded runs reproducible.

``synchronous=True`` (``reader_pool_type='dummy'``) runs everything inside ``get_results`` on the caller's thread.
"""
import os
import queue
import sys
import threading

from petastorm_b200.workers_pool import EmptyResultError, VentilatedItemProcessedMessage

_POLL = 0.005


class WorkerBase(object):
    """What a pool expects of a worker class (the reference's plug-in seam #1): it is constructed as
    ``worker_class(worker_id, publish_func, args)``, receives every ventilated item through ``process(**item)`` -
    publishing zero or more results with ``self.publish_func(result)`` - and is told to release its resources with
    ``shutdown()`` when the pool is joined."""

    def __init__(self, worker_id, publish_func, args):
        self.worker_id, self.publish_func, self.args = worker_id, publish_func, args

    def process(self, *args, **kargs):
        raise NotImplementedError('{} does not implement process()'.format(type(self).__name__))

    def shutdown(self):
        """Nothing to release by default."""


class WorkerTerminationRequested(Exception):
    """Raised inside a worker thread when the pool is being stopped."""


class GpuPool(object):
    def __init__(self, workers_count=1, results_queue_size=3, synchronous=False, device=None, resolvers=0):
        # one issuing thread keeps row-groups in order; `workers_count` only sizes the ventilation window
        self.workers_count = max(1, int(workers_count)) if not synchronous else 1
        self._results_queue_size = max(1, int(results_queue_size))
        self._synchronous = synchronous
        self._device = device
        self._ventilator_queue = None
        self._results_queue = None
        self._worker = None
        self._thread = None
        self._ventilator = None
        self._stop_event = threading.Event()
        self._ventilated_items = 0
        self._ventilated_items_processed = 0
        self._count_lock = threading.Lock()   # get_results() may be called from several consumer threads
        self._started = False
        self._sync_results = []
        # resolver threads: row-groups whose device work was issued are *resolved* (host-visible part: error word, codec
        # launches, host reads) ahead of the consumer, in parallel, and handed out in ventilation order.  Measured (r2g):
        # with a JPEG field (nvJPEG spends ~5 ms of host time per 1024 images) two resolvers lift C3 from 99 k to 153 k
        # images/s; for the other codecs they gain nothing and cost the C1 reader a factor of ten end to end (their polling
        # and short blocking reads starve the issuing thread), hence off unless the reader asks for them.
        self._resolvers = None
        self._resolver_count = 0 if synchronous else int(os.environ.get('PST_RESOLVERS', resolvers))

    # ---- protocol -----------------------------------------------------------------------------------------------
    def start(self, worker_class, worker_args=None, ventilator=None):
        if self._started:
            raise RuntimeError('GpuPool({}) can not be reused! Create a new object'.format(self.workers_count))
        self._started = True
        self._ventilator_queue = queue.Queue()
        # every row-group occupies two entries (its result and its processed marker)
        self._results_queue = queue.Queue(2 * self._results_queue_size + 1)
        publish = self._sync_results.append if self._synchronous else self._stop_aware_put
        self._worker = worker_class(0, publish, worker_args)
        if not self._synchronous:
            if self._resolver_count > 0:
                from concurrent.futures import ThreadPoolExecutor
                self._resolvers = ThreadPoolExecutor(self._resolver_count, thread_name_prefix='pst-gpu-resolve')
            self._thread = threading.Thread(target=self._worker_loop, name='pst-gpu-issue', daemon=True)
            self._thread.start()
        if ventilator:
            self._ventilator = ventilator
            self._ventilator.start()

    def ventilate(self, *args, **kargs):
        self._ventilated_items += 1
        self._ventilator_queue.put((args, kargs))

    def all_done(self):
        return (self._ventilated_items == self._ventilated_items_processed and
                (self._ventilator is None or self._ventilator.completed()))

    def get_results(self):
        if self._synchronous:
            return self._get_results_sync()
        while True:
            if self.all_done() and self._results_queue.empty():
                # re-check after the emptiness test: the worker publishes results before the processed marker
                if self.all_done():
                    raise EmptyResultError()
            try:
                result = self._results_queue.get(timeout=_POLL)
            except queue.Empty:
                continue
            if isinstance(result, VentilatedItemProcessedMessage):
                with self._count_lock:
                    self._ventilated_items_processed += 1
                if self._ventilator:
                    self._ventilator.processed_item()
                continue
            if isinstance(result, _WorkerFailure):
                self.stop()
                self.join()
                raise result.exc.with_traceback(result.tb)
            return result

    def _get_results_sync(self):
        if self._sync_results:
            return self._sync_results.pop(0)
        while self._ventilator_queue.qsize() or (self._ventilator and not self._ventilator.completed()):
            try:
                args, kargs = self._ventilator_queue.get(timeout=0.1)
            except queue.Empty:
                continue
            self._worker.process(*args, **kargs)
            self._ventilated_items_processed += 1
            if self._ventilator:
                self._ventilator.processed_item()
            if self._sync_results:
                return self._sync_results.pop(0)
        raise EmptyResultError()

    def stop(self):
        if self._ventilator:
            self._ventilator.stop()
        self._stop_event.set()

    def join(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self._resolvers is not None:
            self._resolvers.shutdown(wait=True, cancel_futures=True)
            self._resolvers = None
        if self._worker is not None:
            self._worker.shutdown()

    @property
    def diagnostics(self):
        d = {'output_queue_size': self._results_queue.qsize() if self._results_queue is not None else 0,
             'items_ventilated': self._ventilated_items, 'items_processed': self._ventilated_items_processed}
        worker_diag = getattr(self._worker, 'diagnostics', None)
        if worker_diag:
            d.update(worker_diag)
        return d

    # ---- worker thread ------------------------------------------------------------------------------------------
    def _stop_aware_put(self, data):
        """Blocking put into the bounded results queue that gives up when the pool is stopped (so ``stop()`` cannot
        dead-lock against a full queue - petastorm/workers_pool/thread_pool.py:242-256)."""
        while True:
            if self._stop_event.is_set():
                raise WorkerTerminationRequested()
            try:
                self._results_queue.put(data, timeout=_POLL)
                break
            except queue.Full:
                continue
        # once it has a slot in the (bounded) results queue the row-group may be resolved ahead of the consumer
        resolve_ahead = getattr(data, 'resolve_ahead', None)
        if resolve_ahead is not None and self._resolvers is not None:
            resolve_ahead(self._resolvers)

    def _worker_loop(self):
        prof_path = os.environ.get('PST_POOL_PROFILE')      # diagnostics: cProfile of the issuing thread, dumped on exit
        prof = None
        if prof_path:
            import cProfile
            prof = cProfile.Profile()
            try:
                prof.enable()
            except ValueError:      # Python >= 3.12: one profiler per process (e.g. the consumer thread is being profiled)
                prof = None
        try:
            self._worker_loop_body()
        finally:
            if prof is not None:
                prof.disable()
                prof.dump_stats('{}.{}'.format(prof_path, threading.get_ident()))

    def _worker_loop_body(self):
        if self._device is not None:
            import torch
            torch.cuda.set_device(self._device)
        while not self._stop_event.is_set():
            try:
                args, kargs = self._ventilator_queue.get(timeout=_POLL)
            except queue.Empty:
                continue
            try:
                self._worker.process(*args, **kargs)
                self._stop_aware_put(VentilatedItemProcessedMessage())
            except WorkerTerminationRequested:
                return
            except Exception as e:  # pylint: disable=broad-except
                sys.stderr.write('GpuPool worker failed: {!r}\n'.format(e))
                try:
                    self._stop_aware_put(_WorkerFailure(e, sys.exc_info()[2]))
                except WorkerTerminationRequested:
                    pass
                return


class _WorkerFailure(object):
    def __init__(self, exc, tb):
        self.exc = exc
        self.tb = tb
