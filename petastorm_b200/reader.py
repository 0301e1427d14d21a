"""``make_reader`` / ``make_batch_reader`` / :class:`Reader` - the public entry points of the read path.

Signatures, defaults, validation and error behaviour follow ``petastorm/reader.py`` (``make_reader`` ``:60-206``,
``make_batch_reader`` ``:209-352``, ``Reader`` ``:355-730``) so that existing call sites keep working; what runs
underneath is the B200 pipeline (:mod:`petastorm_b200.gpu_workers` on a :class:`~petastorm_b200.workers_pool.gpu_pool.GpuPool`).

Arguments that only make sense for the CPU pools (``workers_count``, ``results_queue_size``, ``pyarrow_serialize``,
``zmq_copy_buffers``, ``hdfs_driver``) are accepted; ``reader_pool_type`` is validated like upstream ('thread' and
'process' both select the asynchronous GPU pool, 'dummy' the synchronous one).  Two keyword additions:
``output='torch'|'numpy'`` (device tensors - the default - or host numpy values with the reference's exact types) and
``device`` (CUDA ordinal, default: current device).
"""
import collections.abc
import logging
import os
import random
import warnings

from petastorm_b200.cache import NullCache
from petastorm_b200.errors import NoDataAvailableError, PetastormMetadataError
from petastorm_b200.etl import dataset_metadata
from petastorm_b200.fs_utils import get_filesystem_and_path_or_paths, normalize_dataset_url_or_urls
from petastorm_b200.gpu_workers import GpuArrowWorker, GpuPyDictWorker, WorkerOptions
from petastorm_b200.ngram import NGram
from petastorm_b200.predicates import PredicateBase
from petastorm_b200.selectors import RowGroupSelectorBase
from petastorm_b200.transform import transform_schema
from petastorm_b200.workers_pool.gpu_pool import GpuPool
from petastorm_b200.workers_pool.ventilator import ConcurrentVentilator

logger = logging.getLogger(__name__)

# no more than workers * (1 + this) row-groups are in flight (petastorm/reader.py:43-45)
_VENTILATE_EXTRA_ROWGROUPS = 3

LOCAL_DISK_CACHE = 'local-disk'
NULL_CACHE = 'null'


def _make_cache(cache_type, cache_location, cache_size_limit, cache_row_size_estimate, cache_extra_settings):
    if cache_type is None or cache_type == NULL_CACHE:
        return NullCache()
    if cache_type == LOCAL_DISK_CACHE:
        from petastorm_b200.local_disk_cache import LocalDiskCache
        return LocalDiskCache(cache_location, cache_size_limit, cache_row_size_estimate, **(cache_extra_settings or {}))
    raise ValueError('Unknown cache_type: {}'.format(cache_type))


def _resolve_device(device):
    """CUDA ordinal the reader works on.  ``None`` means the *caller's* current device (``torch.cuda.set_device`` of a
    DDP rank): the current device is per host thread, so it has to be read here and not on the pool's issuing thread,
    where it would always be 0."""
    import torch
    if device is None:
        if not torch.cuda.is_available():
            from petastorm_b200 import native
            raise native.NativeLibraryError('petastorm_b200 needs a CUDA device (B200, sm_100a); there is no CPU '
                                            'fallback')
        return torch.cuda.current_device()
    if isinstance(device, torch.device):
        return device.index if device.index is not None else torch.cuda.current_device()
    return int(device)


def _make_pool(reader_pool_type, workers_count, results_queue_size, device, resolvers=0):
    if reader_pool_type in ('thread', 'process'):
        # row-groups in flight: one being consumed, one or two decoding (~3 ms), one copying (~5 ms), one queued behind
        # it so that the PCIe copy engine never idles
        pool = GpuPool(workers_count=1, results_queue_size=min(max(int(results_queue_size), 1), 6), device=device,
                       resolvers=resolvers)
        pool.max_in_flight = int(os.environ.get('PST_MAX_IN_FLIGHT', '6'))
        return pool
    if reader_pool_type == 'dummy':
        return GpuPool(synchronous=True, device=device)
    raise ValueError('Unknown reader_pool_type: {}'.format(reader_pool_type))


def make_reader(dataset_url,
                schema_fields=None,
                reader_pool_type='thread', workers_count=10, pyarrow_serialize=False, results_queue_size=50,
                seed=None, shuffle_rows=False,
                shuffle_row_groups=True, shuffle_row_drop_partitions=1,
                predicate=None,
                rowgroup_selector=None,
                num_epochs=1,
                cur_shard=None, shard_count=None, shard_seed=None,
                cache_type=NULL_CACHE, cache_location=None, cache_size_limit=None,
                cache_row_size_estimate=None, cache_extra_settings=None,
                hdfs_driver='libhdfs3',
                transform_spec=None,
                filters=None,
                storage_options=None,
                zmq_copy_buffers=True,
                filesystem=None,
                convert_early_to_numpy=False,
                output='torch', device=None):
    """Reader over a *Petastorm* dataset (one with a stored Unischema): yields one namedtuple per row with codecs
    decoded on the device.  See :func:`make_batch_reader` for plain Parquet stores."""
    device = _resolve_device(device)
    dataset_url_or_urls = normalize_dataset_url_or_urls(dataset_url)
    filesystem, dataset_path = get_filesystem_and_path_or_paths(dataset_url_or_urls, hdfs_driver,
                                                                storage_options=storage_options, filesystem=filesystem)
    cache = _make_cache(cache_type, cache_location, cache_size_limit, cache_row_size_estimate, cache_extra_settings)
    resolvers = 0
    try:
        stored = dataset_metadata.get_schema(dataset_metadata.ParquetDataset(dataset_path))
        # nvJPEG's batched decode has milliseconds of host-side work per row-group: resolve those row-groups ahead of the
        # consumer on two threads (see GpuPool)
        from petastorm_b200.codecs import CompressedImageCodec
        if any(isinstance(f.codec, CompressedImageCodec) and f.codec.image_codec in ('jpeg', 'jpg')
               for f in stored.fields.values()):
            resolvers = 2
    except PetastormMetadataError:
        warnings.warn('Currently make_reader supports reading only Petastorm datasets. '
                      'To read from a non-Petastorm Parquet store use make_batch_reader')
    if reader_pool_type == 'process' and pyarrow_serialize:
        warnings.warn('pyarrow_serializer was deprecated and will be removed in future versions. '
                      'The argument no longer has any effect.')
    reader_pool = _make_pool(reader_pool_type, workers_count, results_queue_size, device, resolvers)
    try:
        return Reader(filesystem, dataset_path,
                      worker_class=GpuPyDictWorker, is_batched_reader=False,
                      schema_fields=schema_fields, reader_pool=reader_pool, shuffle_rows=shuffle_rows, seed=seed,
                      shuffle_row_groups=shuffle_row_groups,
                      shuffle_row_drop_partitions=shuffle_row_drop_partitions, predicate=predicate,
                      rowgroup_selector=rowgroup_selector, num_epochs=num_epochs, cur_shard=cur_shard,
                      shard_count=shard_count, shard_seed=shard_seed, cache=cache, transform_spec=transform_spec,
                      filters=filters, convert_early_to_numpy=convert_early_to_numpy, output=output, device=device)
    except PetastormMetadataError as e:
        logger.error('Unexpected exception: %s', str(e))
        raise RuntimeError('make_reader has failed. If you were trying to open a Parquet store that was not '
                           'created using Petastorm materialize_dataset and it contains only scalar columns, '
                           'you may use make_batch_reader to read it.\n'
                           'Inner exception: %s', str(e))


def make_batch_reader(dataset_url_or_urls,
                      schema_fields=None,
                      reader_pool_type='thread', workers_count=10,
                      results_queue_size=50,
                      seed=None, shuffle_rows=False,
                      shuffle_row_groups=True, shuffle_row_drop_partitions=1,
                      predicate=None,
                      rowgroup_selector=None,
                      num_epochs=1,
                      cur_shard=None, shard_count=None, shard_seed=None,
                      cache_type='null', cache_location=None, cache_size_limit=None,
                      cache_row_size_estimate=None, cache_extra_settings=None,
                      hdfs_driver='libhdfs3',
                      transform_spec=None,
                      filters=None,
                      storage_options=None,
                      zmq_copy_buffers=True,
                      filesystem=None,
                      convert_early_to_numpy=False,
                      output='torch', device=None):
    """Reader over a plain Parquet store (native scalar / list-of-primitive columns): yields one namedtuple of column
    arrays per row-group; re-batching is the loader's job."""
    device = _resolve_device(device)
    dataset_url_or_urls = normalize_dataset_url_or_urls(dataset_url_or_urls)
    filesystem, dataset_path_or_paths = get_filesystem_and_path_or_paths(
        dataset_url_or_urls, hdfs_driver, storage_options=storage_options, filesystem=filesystem)
    try:
        dataset_metadata.get_schema(dataset_metadata.ParquetDataset(dataset_path_or_paths))
        warnings.warn('Please use make_reader (instead of \'make_batch_dataset\' function to read this dataset. '
                      'You may get unexpected results. '
                      'Currently make_batch_reader supports reading only Parquet stores that contain '
                      'standard Parquet data types and do not require petastorm decoding.')
    except PetastormMetadataError:
        pass
    cache = _make_cache(cache_type, cache_location, cache_size_limit, cache_row_size_estimate, cache_extra_settings)
    reader_pool = _make_pool(reader_pool_type, workers_count, results_queue_size, device)
    return Reader(filesystem, dataset_path_or_paths,
                  schema_fields=schema_fields, worker_class=GpuArrowWorker, reader_pool=reader_pool, seed=seed,
                  shuffle_rows=shuffle_rows, shuffle_row_groups=shuffle_row_groups,
                  shuffle_row_drop_partitions=shuffle_row_drop_partitions, predicate=predicate,
                  rowgroup_selector=rowgroup_selector, num_epochs=num_epochs, cur_shard=cur_shard,
                  shard_count=shard_count, shard_seed=shard_seed, cache=cache, transform_spec=transform_spec,
                  is_batched_reader=True, filters=filters, convert_early_to_numpy=convert_early_to_numpy,
                  output=output, device=device)


class Reader(object):
    """Iterator over a dataset.

    :ivar last_row_consumed: True once the last row was returned.
    """

    def __init__(self, pyarrow_filesystem, dataset_path, schema_fields=None,
                 seed=None, shuffle_rows=False, shuffle_row_groups=True,
                 shuffle_row_drop_partitions=1,
                 predicate=None, rowgroup_selector=None, reader_pool=None, num_epochs=1,
                 cur_shard=None, shard_count=None, cache=None, worker_class=None,
                 transform_spec=None, is_batched_reader=False, filters=None, shard_seed=None,
                 convert_early_to_numpy=False, output='torch', device=None):
        self.num_epochs = num_epochs
        if not (isinstance(schema_fields, collections.abc.Iterable) or isinstance(schema_fields, NGram)
                or schema_fields is None):
            raise ValueError('Fields must be either None, an iterable collection of Unischema fields '
                             'or an NGram object.')
        if output not in ('torch', 'numpy'):
            raise ValueError("output must be 'torch' or 'numpy'")
        self.is_batched_reader = is_batched_reader
        # fail at construction (not at the first next()) when there is no CUDA device: no CPU fallback exists
        from petastorm_b200 import rowgroup
        device = _resolve_device(device)
        self.device = device
        rowgroup.get_context(device)

        # 1. open the dataset
        self.dataset = dataset_metadata.ParquetDataset(dataset_path, filters=filters)
        stored_schema = dataset_metadata.infer_or_load_unischema(self.dataset)

        if isinstance(schema_fields, NGram):
            self.ngram = schema_fields
            self.ngram.resolve_regex_field_names(stored_schema)
        else:
            self.ngram = None

        worker_class = worker_class or GpuPyDictWorker
        self._results_queue_reader = worker_class.new_results_queue_reader()
        if hasattr(self._results_queue_reader, '_output'):
            self._results_queue_reader._output = output  # pylint: disable=protected-access

        if self.ngram and not self.ngram.timestamp_overlap and shuffle_row_drop_partitions > 1:
            raise NotImplementedError('Using timestamp_overlap=False is not implemented with'
                                      ' shuffle_options.shuffle_row_drop_partitions > 1')

        self.cache = cache or NullCache()
        self._workers_pool = reader_pool or GpuPool(device=device)

        if self.ngram:
            fields = self.ngram.get_field_names_at_all_timesteps()
        else:
            fields = schema_fields if isinstance(schema_fields, collections.abc.Iterable) else None

        storage_schema = stored_schema.create_schema_view(fields) if fields else stored_schema
        if len(storage_schema.fields) == 0:
            raise RuntimeError("No fields matching the criteria '{}' were found in the dataset {}.".format(
                fields, dataset_path))
        self.schema = transform_schema(storage_schema, transform_spec) if transform_spec else storage_schema

        # 2. all row-groups
        row_groups = dataset_metadata.load_row_groups(self.dataset)

        # 3. filter them: partition-level predicate, selector, shard
        _shard_seed = seed
        if shard_seed:
            warnings.warn('shard_seed was deprecated and will be removed in future versions. '
                          'Use seed to apply randomization effects on sharding row groups.')
            _shard_seed = shard_seed
        filtered_row_group_indexes, worker_predicate = self._filter_row_groups(
            self.dataset, row_groups, predicate, rowgroup_selector, cur_shard, shard_count, _shard_seed)

        # 4. ventilator
        normalized_drop = self._normalize_shuffle_options(shuffle_row_drop_partitions, row_groups)
        self.ventilator = self._create_ventilator(filtered_row_group_indexes, shuffle_row_groups, normalized_drop,
                                                  self.num_epochs, worker_predicate,
                                                  getattr(self._workers_pool, 'max_in_flight', None) or
                                                  self._workers_pool.workers_count * (1 + _VENTILATE_EXTRA_ROWGROUPS),
                                                  seed)

        # 5. start the pool; the worker receives the reference's 12-tuple plus the GPU options
        options = WorkerOptions(partitions=self.dataset.partitions, output=output, device=device)
        self._workers_pool.start(worker_class, (pyarrow_filesystem, dataset_path, storage_schema, self.ngram, row_groups,
                                                self.cache, transform_spec, self.schema, filters, shuffle_rows, seed,
                                                convert_early_to_numpy, options),
                                 ventilator=self.ventilator)
        logger.debug('Workers pool started')
        self.last_row_consumed = False
        self.stopped = False

    def reset(self):
        """Start over once every sample of all epochs was consumed; raises ``NotImplementedError`` mid-iteration."""
        if not self.last_row_consumed:
            raise NotImplementedError('Currently do not support resetting a reader while in the middle of iteration. '
                                      'You can call reset only after all samples were consumed.')
        self.last_row_consumed = False
        self.ventilator.reset()

    @property
    def batched_output(self):
        return self._results_queue_reader.batched_output

    # ---- row-group filtering (petastorm/reader.py:533-652) ------------------------------------------------------
    def _filter_row_groups(self, dataset, row_groups, predicate, rowgroup_selector, cur_shard, shard_count, seed):
        indexes, worker_predicate = self._apply_predicate_to_row_groups(dataset, row_groups, predicate)
        if rowgroup_selector:
            indexes = self._apply_row_group_selector(dataset, rowgroup_selector, indexes)
        if cur_shard is not None or shard_count is not None:
            indexes = self._partition_row_groups(dataset, row_groups, shard_count, cur_shard, indexes, seed)
        if not indexes:
            warnings.warn('No matching data is available for loading after rowgroup '
                          'selector were applied and the data was sharded.')
        return indexes, worker_predicate

    def _partition_row_groups(self, dataset, row_groups, shard_count, cur_shard, filtered_row_group_indexes, seed):
        """Shard = row-groups whose *index value* is congruent to ``cur_shard`` modulo ``shard_count``; a seed only
        permutes the visiting order (petastorm/reader.py:573-597).  Every rank evaluates this rule locally;
        :func:`petastorm_b200.sharding.broadcast_row_group_assignment` makes the ranks agree through one NCCL
        broadcast."""
        if not shard_count or not isinstance(cur_shard, int) or not isinstance(shard_count, int):
            raise ValueError('partition and num_partitions must be ints and both specified to use partitioning')
        if shard_count is not None and len(row_groups) < shard_count:
            raise NoDataAvailableError('Number of row-groups in the dataset must be greater or equal to the number of '
                                       'requested shards. Otherwise, some of the shards will end up being empty.')
        if seed is not None:
            random.Random(seed).shuffle(filtered_row_group_indexes)
        return [i for i in filtered_row_group_indexes if i % shard_count == cur_shard]

    def _apply_row_group_selector(self, dataset, rowgroup_selector, filtered_row_group_indexes):
        if not isinstance(rowgroup_selector, RowGroupSelectorBase):
            raise ValueError('rowgroup_selector parameter is expected to be derived from RowGroupSelectorBase')
        available = dataset_metadata.get_row_group_indexes(dataset)
        required = rowgroup_selector.get_index_names()
        if not set(required).issubset(set(available.keys())):
            raise ValueError('Some of required indexes {} are not available in {}'.format(required,
                                                                                         list(available.keys())))
        selected = rowgroup_selector.select_row_groups(available)
        return [idx for idx in filtered_row_group_indexes if idx in selected]

    def _apply_predicate_to_row_groups(self, dataset, row_groups, predicate):
        if not predicate:
            return list(range(len(row_groups))), None
        if not isinstance(predicate, PredicateBase):
            raise ValueError('predicate parameter is expected to be derived from PredicateBase')
        predicate_fields = predicate.get_fields()
        partition_names = dataset.partitions.partition_names if dataset.partitions else set()
        if set(predicate_fields) == partition_names:
            assert len(partition_names) == 1, 'Datasets with only a single partition level supported at the moment'
            kept = []
            for piece_index, piece in enumerate(row_groups):
                name, value = piece.partition_keys[0]
                value = self.schema.fields[name].numpy_dtype(value)  # typed per the schema (reader.py:640-641)
                if predicate.do_include({name: value}):
                    kept.append(piece_index)
            return kept, None
        return list(range(len(row_groups))), predicate

    @staticmethod
    def _normalize_shuffle_options(shuffle_row_drop_partitions, row_groups):
        """Never ask for more drop partitions than the largest row-group has rows (petastorm/reader.py:654-664)."""
        if shuffle_row_drop_partitions > 1 and row_groups:
            from petastorm_b200 import rowgroup
            largest = 1
            for piece in row_groups:
                largest = max(largest, rowgroup.open_file(piece.path).row_group_num_rows(piece.row_group))
            return min(shuffle_row_drop_partitions, largest)
        return shuffle_row_drop_partitions

    def _create_ventilator(self, row_group_indexes, shuffle_row_groups, shuffle_row_drop_partitions, num_epochs,
                           worker_predicate, max_ventilation_queue_size, seed):
        items = [{'piece_index': piece_index,
                  'worker_predicate': worker_predicate,
                  'shuffle_row_drop_partition': (part, shuffle_row_drop_partitions)}
                 for piece_index in row_group_indexes for part in range(shuffle_row_drop_partitions)]
        return ConcurrentVentilator(self._workers_pool.ventilate, items, iterations=num_epochs,
                                    max_ventilation_queue_size=max_ventilation_queue_size,
                                    randomize_item_order=shuffle_row_groups, random_seed=seed)

    # ---- life-cycle ---------------------------------------------------------------------------------------------
    def stop(self):
        self._workers_pool.stop()
        self.stopped = True

    def join(self):
        self._workers_pool.join()

    def cleanup_cache(self):
        cleanup = getattr(self.cache, 'cleanup', None)
        if cleanup and not isinstance(self.cache, NullCache):
            try:
                cleanup()
            except (OSError, IOError, AttributeError) as e:
                print('Error cleaning cache: {}'.format(e))

    @property
    def diagnostics(self):
        return self._workers_pool.diagnostics

    def __iter__(self):
        return self

    def __next__(self):
        if self.stopped:
            raise RuntimeError('Trying to read a sample after a reader created by '
                               'make_reader/make_batch_reader has stopped. This may happen if the '
                               'make_reader/make_batch_reader context manager has exited but you try to '
                               'fetch a sample from it anyway')
        try:
            return self._results_queue_reader.read_next(self._workers_pool, self.schema, self.ngram)
        except StopIteration:
            self.last_row_consumed = True
            raise

    def next(self):
        return self.__next__()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.stop()
        self.join()
        self.cleanup_cache()
