"""Cache protocol of the readers (petastorm/cache.py:21-39).  The GPU path reads local files through the page cache
and a pinned row-group cache inside libpst_b200.so, so only the null cache is provided; custom ``CacheBase``
implementations are honoured by the workers."""
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
