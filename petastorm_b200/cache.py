"""Cache protocol of the readers (petastorm/cache.py:21-39): ``get(key, fill_cache_func)`` of decoded row-groups.
:class:`NullCache` is the default (the GPU path already keeps raw row-groups in the page cache, a pinned host cache and -
optionally - an HBM-resident cache, see :mod:`petastorm_b200.rowgroup`); :class:`petastorm_b200.local_disk_cache.
LocalDiskCache` is ``cache_type='local-disk'``; custom ``CacheBase`` implementations are honoured by the workers."""
import abc


class CacheBase(abc.ABC):
    @abc.abstractmethod
    def get(self, key, fill_cache_func):
        """Value for ``key``; on a miss call ``fill_cache_func()``, store and return its result."""


class NullCache(CacheBase):
    """Never stores anything."""

    def get(self, key, fill_cache_func):
        return fill_cache_func()

    def cleanup(self):
        pass
