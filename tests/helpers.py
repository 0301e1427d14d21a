"""Shared test helpers."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def oracle_specs(schema):
    """Field descriptions the oracle port understands, from a petastorm_b200 Unischema."""
    from petastorm_b200.codecs import CompressedImageCodec, CompressedNdarrayCodec, NdarrayCodec, ScalarCodec
    specs = {}
    for name, f in schema.fields.items():
        if isinstance(f.codec, CompressedImageCodec):
            kind = 'image'
        elif isinstance(f.codec, CompressedNdarrayCodec):
            kind = 'compressed_ndarray'
        elif isinstance(f.codec, NdarrayCodec):
            kind = 'ndarray'
        elif isinstance(f.codec, ScalarCodec):
            kind = 'scalar'
        else:
            kind = None
        specs[name] = {'codec': kind, 'dtype': f.numpy_dtype}
    return specs


def load_schema(dataset_dir):
    from petastorm_b200.etl import dataset_metadata as dm
    return dm.get_schema(dm.ParquetDataset(dataset_dir))


def to_host(value):
    if hasattr(value, 'cpu') and hasattr(value, 'numpy'):
        return value.cpu().numpy()
    return value
