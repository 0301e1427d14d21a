"""URL handling for the readers (role of petastorm/fs_utils.py:179-218).

The B200 path mmaps *local* files, so only ``file://`` URLs (and plain paths) resolve; ``hdfs://``, ``s3://``,
``gs://`` raise a clear error instead of silently taking a slow path (out of scope)."""
from urllib.parse import urlparse


def normalize_dir_url(dataset_url):
    if dataset_url is None or not isinstance(dataset_url, str):
        raise ValueError('directory url must be a string')
    dataset_url = dataset_url[:-1] if dataset_url.endswith('/') and len(dataset_url) > 1 else dataset_url
    return dataset_url


def normalize_dataset_url_or_urls(dataset_url_or_urls):
    if isinstance(dataset_url_or_urls, list):
        if not dataset_url_or_urls:
            raise ValueError('dataset url list must be non-empty.')
        return [normalize_dir_url(url) for url in dataset_url_or_urls]
    return normalize_dir_url(dataset_url_or_urls)


def _local_path(url):
    parsed = urlparse(url)
    if parsed.scheme in ('', 'file'):
        if parsed.scheme == 'file' and parsed.netloc not in ('', 'localhost'):
            raise ValueError('file:// urls must not carry a host: {}'.format(url))
        return parsed.path if parsed.scheme == 'file' else url
    raise ValueError('petastorm_b200 reads local files only (file:// urls): the GPU path mmaps row-groups into pinned '
                     'host memory. Unsupported scheme "{}" in {}'.format(parsed.scheme, url))


def get_dataset_path(parsed_url):
    return parsed_url.path


def get_filesystem_and_path_or_paths(url_or_urls, hdfs_driver='libhdfs3', storage_options=None, filesystem=None):
    """(filesystem, path or list of paths).  ``filesystem`` is always ``None`` (local mmap I/O)."""
    if filesystem is not None:
        raise ValueError('a custom filesystem object cannot be used: the GPU path mmaps local files')
    if isinstance(url_or_urls, list):
        schemes = set(urlparse(u).scheme for u in url_or_urls)
        if len(schemes) > 1:
            raise ValueError('The dataset url list must contain url with the same scheme.')
        return None, [_local_path(u) for u in url_or_urls]
    return None, _local_path(url_or_urls)
