"""``WeightedSamplingReader``: mixes the output of two or more readers, picking the reader of every ``next()`` with a
configurable probability (API and behaviour of petastorm/weighted_sampling_reader.py:20-115).

It quacks like a :class:`~petastorm_b200.reader.Reader` (``batched_output``, ``ngram``, ``schema``, ``last_row_consumed``,
``stop`` / ``join``, context manager), so it can be handed to the PyTorch loaders; over petastorm_b200 readers the samples
it hands out are the device tensors of the underlying readers - nothing is copied.  Iteration stops as soon as one of
the readers runs out of data.  The draw uses the global ``np.random`` state, like upstream.
"""
import numpy as np


class WeightedSamplingReader(object):
    def __init__(self, readers, probabilities):
        if len(readers) <= 1:
            raise ValueError('Two or more readers must be specified. Got {}.'.format(len(readers)))
        if len(readers) != len(probabilities):
            raise ValueError('readers and probabilities are expected to be lists of the same length')
        self._readers = readers
        # probabilities that do not add up to one are normalised
        self._cum_prob = np.cumsum(np.asarray(probabilities, dtype=np.float64) / np.sum(probabilities))
        first = readers[0]
        for other in readers[1:]:
            if first.batched_output != other.batched_output:
                raise ValueError('All readers passed to WeightedSamplingReader should have the same value of '
                                 '"batched_output" attribute')
            if set(first.schema.fields.keys()) != set(other.schema.fields.keys()):
                raise ValueError('All readers passed to WeightedSamplingReader should have the same schema')
            both = first.ngram is not None and other.ngram is not None
            if (first.ngram is None) != (other.ngram is None) or (both and first.ngram != other.ngram):
                raise ValueError('All readers passed to WeightedSamplingReader should have the same ngram spec')
        self.batched_output = first.batched_output
        self.ngram = first.ngram
        self.schema = first.schema

    def __len__(self):
        return sum(len(reader) for reader in self._readers)

    def __iter__(self):
        return self

    def __next__(self):
        r = np.random.random()
        return next(self._readers[int(np.where(r < self._cum_prob)[0][0])])

    def next(self):
        return self.__next__()

    @property
    def last_row_consumed(self):
        return any(r.last_row_consumed for r in self._readers)

    def reset(self):
        for reader in self._readers:
            reader.reset()

    def stop(self):
        for reader in self._readers:
            reader.stop()

    def join(self):
        for reader in self._readers:
            reader.join()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.stop()
        self.join()
