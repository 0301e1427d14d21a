"""Write side without Spark: materialise a Petastorm dataset with pyarrow.

The reference writes datasets from a Spark job (``materialize_dataset``, petastorm/etl/dataset_metadata.py:52-132 +
``dict_to_spark_row``, petastorm/unischema.py:359-406); there is no JVM here, so fixtures and synthetic benchmark data
are produced by this module instead (SURVEY 8f "next" #2).  The result has the same on-disk contract:

* rows are encoded with the field codecs (``codec.encode``), scalars stored in native parquet types;
* ``_common_metadata`` carries ``dataset-toolkit.unischema.v1`` (a protocol-2 pickle whose globals are spelled
  ``petastorm.unischema.*`` / ``petastorm.codecs.*`` / ``pyspark.sql.types.*`` so that the *reference* can read the
  dataset too) and ``dataset-toolkit.num_row_groups_per_file.v1``.

Not part of the GPU hot path (pyarrow is only used here, on the write side).
"""
import json
import os
import pickle
from decimal import Decimal

import numpy as np

from petastorm_b200 import spark_types
from petastorm_b200.codecs import ScalarCodec
from petastorm_b200.etl.dataset_metadata import ROW_GROUPS_PER_FILE_KEY, UNISCHEMA_KEY
from petastorm_b200.unischema import insert_explicit_nulls


def pickle_unischema_reference_compatible(schema):
    """Protocol-2 pickle of a Unischema with module names rewritten to the reference's package layout."""
    data = pickle.dumps(schema, protocol=2)
    data = data.replace(b'cpetastorm_b200.spark_types\n', b'cpyspark.sql.types\n')
    data = data.replace(b'cpetastorm_b200.unischema\n', b'cpetastorm.unischema\n')
    data = data.replace(b'cpetastorm_b200.codecs\n', b'cpetastorm.codecs\n')
    return data


def _arrow_type_of(field):
    import pyarrow as pa
    if field.codec is not None and not isinstance(field.codec, ScalarCodec):
        return pa.binary()
    if isinstance(field.codec, ScalarCodec):
        tname = type(field.codec.spark_dtype()).__name__
        table = {'ByteType': pa.int8(), 'ShortType': pa.int16(), 'IntegerType': pa.int32(), 'LongType': pa.int64(),
                 'FloatType': pa.float32(), 'DoubleType': pa.float64(), 'BooleanType': pa.bool_(),
                 'StringType': pa.string(), 'BinaryType': pa.binary()}
        if tname == 'DecimalType':
            st = field.codec.spark_dtype()
            return pa.decimal128(st.precision, st.scale)
        return table[tname]
    dt = field.numpy_dtype
    if dt is Decimal:
        return pa.decimal128(38, 18)
    if dt in (np.str_,):
        return pa.string()
    if dt in (np.bytes_,):
        return pa.binary()
    return pa.from_numpy_dtype(np.dtype(dt))


def encode_row(schema, row):
    """dict of python/numpy values -> dict of storable cells (role of ``dict_to_spark_row``)."""
    row = dict(row)
    insert_explicit_nulls(schema, row)
    if set(row.keys()) != set(schema.fields.keys()):
        raise ValueError('Dictionary fields \n{}\n do not match schema fields \n{}'.format(
            '\n'.join(sorted(row.keys())), '\n'.join(schema.fields.keys())))
    out = {}
    for name, value in row.items():
        field = schema.fields[name]
        if value is None:
            if not field.nullable:
                raise ValueError('Field {} is not "nullable", but got passes a None value'.format(name))
            out[name] = None
        elif field.codec is not None:
            enc = field.codec.encode(field, value)
            out[name] = bytes(enc) if isinstance(enc, (bytearray, memoryview)) else enc
        else:
            out[name] = value.tolist() if isinstance(value, np.generic) else value
    return out


def arrow_schema_of(schema, partition_by=None):
    """The pyarrow schema a dataset with this Unischema is stored with (codec fields are binary blobs)."""
    import pyarrow as pa
    names = [n for n in schema.fields.keys() if n != partition_by]
    return pa.schema([pa.field(n, _arrow_type_of(schema.fields[n]), nullable=bool(schema.fields[n].nullable))
                      for n in names])


def write_common_metadata(output_dir, schema, row_groups_per_file, partition_by=None):
    """``_common_metadata`` of a Petastorm dataset: the pickled Unischema and the row-group count of every file
    (``{path relative to output_dir: num_row_groups}``) - what ``materialize_dataset`` adds after the Spark job
    (petastorm/etl/dataset_metadata.py:194-241).  Lets part files be written independently (e.g. in parallel)."""
    import pyarrow.parquet as pq
    meta = {UNISCHEMA_KEY: pickle_unischema_reference_compatible(schema),
            ROW_GROUPS_PER_FILE_KEY: json.dumps(row_groups_per_file).encode()}
    pq.write_metadata(arrow_schema_of(schema, partition_by).with_metadata(meta),
                      os.path.join(output_dir, '_common_metadata'))


def write_petastorm_dataset(output_dir, schema, rows, rows_per_file=None, row_group_rows=None, compression='snappy',
                            partition_by=None, data_page_size=None, use_dictionary=True):
    """Encode `rows` (iterable of dicts) and write them as a Petastorm dataset under `output_dir`.

    :param rows_per_file: rows per parquet file (default: everything in one file per partition)
    :param row_group_rows: rows per row-group inside a file
    :param partition_by: optional field name; its values become hive ``name=value`` directories
    :return: list of written file paths
    """
    import pyarrow as pa
    import pyarrow.parquet as pq
    os.makedirs(output_dir, exist_ok=True)
    names = [n for n in schema.fields.keys() if n != partition_by]
    arrow_schema = arrow_schema_of(schema, partition_by)
    buckets = {}
    for row in rows:
        enc = encode_row(schema, row)
        key = str(enc[partition_by]) if partition_by else ''
        buckets.setdefault(key, []).append(enc)
    written = []
    per_file = {}
    for key in sorted(buckets.keys()):
        enc_rows = buckets[key]
        directory = os.path.join(output_dir, '{}={}'.format(partition_by, key)) if partition_by else output_dir
        os.makedirs(directory, exist_ok=True)
        step = rows_per_file or max(len(enc_rows), 1)
        for fi, start in enumerate(range(0, len(enc_rows), step)):
            chunk = enc_rows[start:start + step]
            table = pa.Table.from_pydict({n: [r[n] for r in chunk] for n in names}, schema=arrow_schema)
            path = os.path.join(directory, 'part-{:05d}.parquet'.format(fi))
            kwargs = {}
            if data_page_size:
                kwargs['data_page_size'] = data_page_size
            pq.write_table(table, path, compression=compression, row_group_size=row_group_rows or len(chunk) or 1,
                           use_dictionary=use_dictionary, **kwargs)
            per_file[os.path.relpath(path, output_dir)] = pq.ParquetFile(path).metadata.num_row_groups
            written.append(path)
    write_common_metadata(output_dir, schema, per_file, partition_by)
    return written


def spark_type_for(numpy_dtype):
    """Convenience for building ``ScalarCodec`` fields without pyspark."""
    table = {np.int8: spark_types.ByteType, np.uint8: spark_types.ShortType, np.int16: spark_types.ShortType,
             np.uint16: spark_types.IntegerType, np.int32: spark_types.IntegerType, np.int64: spark_types.LongType,
             np.float32: spark_types.FloatType, np.float64: spark_types.DoubleType, np.str_: spark_types.StringType,
             np.bytes_: spark_types.BinaryType, np.bool_: spark_types.BooleanType}
    return table[numpy_dtype]()
