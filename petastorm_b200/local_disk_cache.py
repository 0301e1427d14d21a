"""``LocalDiskCache``: keeps decoded row-groups on a local file system (``cache_type='local-disk'`` of ``make_reader`` /
``make_batch_reader``; API of petastorm/local_disk_cache.py:23-82, which wraps ``diskcache.FanoutCache``).

The reference uses it to avoid re-reading (and re-decoding) a remote store in later epochs.  Here a cached value is the
host image of a decoded row-group (numpy arrays / python objects, pickled to one file per key under ``path``, spread
over ``shards`` directories); a hit costs one file read + one H2D copy of the *decoded* columns and no page decode at
all.  Eviction is least-recently-stored (files by modification time) once ``size_limit_bytes`` is exceeded - like the
reference's default ``eviction_policy``; ``eviction_policy='none'`` stops storing when the cache is full.
"""
import hashlib
import os
import pickle
import shutil
import threading

from petastorm_b200.cache import CacheBase


class LocalDiskCache(CacheBase):
    def __init__(self, path, size_limit_bytes, expected_row_size_bytes, shards=6, cleanup=True, **settings):
        """:param path: directory of the cache (created)
        :param size_limit_bytes: disk budget; may be exceeded by one value while it is being stored
        :param expected_row_size_bytes: approximate size of a row, used for the same sanity check as upstream
        :param shards: number of sub-directories the files are spread over
        :param cleanup: remove the directory in :meth:`cleanup`
        :param settings: ``eviction_policy`` ('least-recently-stored' (default) or 'none'); other diskcache settings of
          the reference are accepted and ignored"""
        self._eviction_policy = settings.get('eviction_policy', 'least-recently-stored')
        if self._eviction_policy != 'none' and size_limit_bytes / shards < 5 * expected_row_size_bytes:
            raise ValueError('Condition \'size_limit_bytes / shards < 5 * expected_row_size_bytes\' needs to hold, '
                             'otherwise, newly added cached values might end up being immediately evicted.')
        self._cleanup = cleanup
        self._path = path
        self._size_limit_bytes = int(size_limit_bytes)
        self._shards = max(1, int(shards))
        self._lock = threading.Lock()
        self.hits = 0
        self.misses = 0
        for s in range(self._shards):
            os.makedirs(os.path.join(path, '%03d' % s), exist_ok=True)

    def _file_of(self, key):
        digest = hashlib.md5(str(key).encode('utf-8')).hexdigest()
        return os.path.join(self._path, '%03d' % (int(digest[:8], 16) % self._shards), digest + '.rowgroup')

    def _files(self):
        out = []
        for s in range(self._shards):
            d = os.path.join(self._path, '%03d' % s)
            for name in os.listdir(d):
                if name.endswith('.rowgroup'):
                    p = os.path.join(d, name)
                    try:
                        st = os.stat(p)
                    except OSError:
                        continue
                    out.append((st.st_mtime_ns, st.st_size, p))
        return out

    def volume(self):
        """Bytes currently stored."""
        return sum(size for _, size, _ in self._files())

    def get(self, key, fill_cache_func):
        from petastorm_b200 import gpu_workers
        path = self._file_of(key)
        try:
            with open(path, 'rb') as f:
                payload = pickle.load(f)
            self.hits += 1
            return gpu_workers.from_host_payload(payload)
        except (OSError, EOFError, pickle.UnpicklingError):
            pass
        self.misses += 1
        value = fill_cache_func()
        payload = gpu_workers.to_host_payload(value)
        with self._lock:
            if self._eviction_policy == 'none' and self.volume() >= self._size_limit_bytes:
                return value
            tmp = path + '.tmp%d' % threading.get_ident()
            with open(tmp, 'wb') as f:
                pickle.dump(payload, f, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(tmp, path)
            if self._eviction_policy != 'none':
                files = sorted(self._files())
                total = sum(size for _, size, _ in files)
                for _, size, p in files:
                    if total <= self._size_limit_bytes or p == path:
                        break
                    try:
                        os.remove(p)
                        total -= size
                    except OSError:
                        pass
        return value

    def cleanup(self):
        if self._cleanup:
            shutil.rmtree(self._path, ignore_errors=True)
