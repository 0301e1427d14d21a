"""Ventilator: feeds work items (row-group reads) to a pool from a background thread, with back-pressure, epochs and
an optional seeded shuffle of the item order (semantics of petastorm/workers_pool/ventilator.py:55-174).

Behaviours kept on purpose because seeded runs must reproduce the reference's order:
* the item order is permuted once per :meth:`start` (so once per ``Reader.reset``), not per epoch;
* ``random_seed`` ``None`` or ``0`` means *unseeded* (global ``np.random``), any other value draws from one
  ``np.random.default_rng(seed)`` stream that lives as long as the ventilator.
"""
import threading
from abc import ABC, abstractmethod
import numpy as np

_VENTILATION_INTERVAL = 0.01


class Ventilator(ABC):
    def __init__(self, ventilate_fn):
        self._ventilate_fn = ventilate_fn

    @abstractmethod
    def start(self):
        """Begin ventilating; the pool must already accept items."""

    @abstractmethod
    def processed_item(self):
        """Pool callback: one ventilated item has been fully processed."""

    @abstractmethod
    def completed(self):
        """True once nothing more will ever be ventilated."""

    @abstractmethod
    def stop(self):
        """Stop ventilating."""


class ConcurrentVentilator(Ventilator):
    def __init__(self, ventilate_fn, items_to_ventilate, iterations=1, randomize_item_order=False, random_seed=None,
                 max_ventilation_queue_size=None, ventilation_interval=_VENTILATION_INTERVAL):
        super(ConcurrentVentilator, self).__init__(ventilate_fn)
        if iterations is not None and (not isinstance(iterations, int) or iterations < 1):
            raise ValueError('iterations must be positive integer or None')
        if not isinstance(items_to_ventilate, list) or any(not isinstance(i, dict) for i in items_to_ventilate):
            raise ValueError('items_to_ventilate must be a list of dicts')
        self._items = items_to_ventilate
        self._iterations = iterations
        self._iterations_remaining = iterations
        self._randomize = randomize_item_order
        self._seed = random_seed
        self._rng = np.random.default_rng(random_seed)
        self._max_queue = max_ventilation_queue_size or len(items_to_ventilate)
        self._interval = ventilation_interval
        self._cursor = 0
        self._thread = None
        self._ventilated = 0
        self._processed = 0
        self._stop_requested = False
        # back-pressure wake-up: the reference polls every `ventilation_interval` (10 ms) while the window is full,
        # which is longer than a whole row-group decode here, so `processed_item` wakes the thread instead
        self._window = threading.Condition()

    def start(self):
        self._thread = threading.Thread(target=self._run, name='pst-ventilator', daemon=True)
        self._thread.start()

    def processed_item(self):
        with self._window:
            self._processed += 1
            self._window.notify()

    def completed(self):
        return self._stop_requested or self._iterations_remaining == 0 or not self._items

    def reset(self):
        """Restart from the first epoch; only legal once everything was ventilated."""
        if not self.completed():
            raise NotImplementedError('Reseting ventilator while ventilating is not supported.')
        self._iterations_remaining = self._iterations
        self.start()

    def _run(self):
        if self._randomize:
            order = (self._rng if (self._seed is not None and self._seed != 0) else np.random).permutation(len(self._items))
            self._items = [self._items[i] for i in order]
        while not self.completed():
            with self._window:
                if self._ventilated - self._processed >= self._max_queue:
                    self._window.wait(self._interval)
                    continue
            self._ventilate_fn(**self._items[self._cursor])
            self._cursor += 1
            self._ventilated += 1
            if self._cursor >= len(self._items):
                self._cursor = 0
                if self._iterations_remaining is not None:
                    self._iterations_remaining -= 1

    def stop(self):
        self._stop_requested = True
        with self._window:
            self._window.notify()
        if self._thread is not None:
            self._thread.join()
            self._thread = None
