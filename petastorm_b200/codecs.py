"""Field codecs.  Class names and instance attributes match ``petastorm/codecs.py`` because codec instances are
pickled into datasets (petastorm/codecs.py:20-21): ``CompressedImageCodec._image_codec/_quality``,
``ScalarCodec._spark_type``.

Two decode surfaces:

* ``decode(unischema_field, value)`` - the reference's per-value contract (petastorm/codecs.py:36-55).  It is kept for
  API compatibility and for values that cannot live on the device (strings, decimals); it runs on the host.
* ``decode_batch(field, column, rows)`` - the B200 path: a whole row-group column (device BYTE_ARRAY references) is
  decoded by one kernel launch into a ``[n, *shape]`` tensor.  This is what the readers use.

``encode`` (write side, SURVEY 8f "next") is provided for fixtures and uses numpy / OpenCV like the reference.
"""
import ast
import struct
from io import BytesIO

import numpy as np

try:
    import cv2
    OPENCV_AVAILABLE = True
except ImportError:  # pragma: no cover
    OPENCV_AVAILABLE = False


class DataframeColumnCodec(object):
    """Abstract codec (petastorm/codecs.py:36-55)."""

    def encode(self, unischema_field, value):
        raise RuntimeError('Abstract method was called')

    def decode(self, unischema_field, value):
        raise RuntimeError('Abstract method was called')

    def spark_dtype(self):
        raise RuntimeError('Abstract method was called')


def _is_compliant_shape(a, b):
    """Shapes match, ``None``/0 dimensions are wildcards (petastorm/codecs.py:274-294)."""
    if len(a) != len(b):
        return False
    return all(not (x and y) or x == y for x, y in zip(a, b))


def _check_ndarray(unischema_field, value):
    expected = unischema_field.numpy_dtype
    if not isinstance(value, np.ndarray):
        raise ValueError('Unexpected type of {} feature. Expected ndarray of {}. Got {}'.format(
            unischema_field.name, expected, type(value)))
    if expected != value.dtype.type:
        raise ValueError('Unexpected type of {} feature. Expected {}. Got {}'.format(
            unischema_field.name, expected, value.dtype))
    if not _is_compliant_shape(value.shape, unischema_field.shape):
        raise ValueError('Unexpected dimensions of {} feature. Expected {}. Got {}'.format(
            unischema_field.name, unischema_field.shape, value.shape))


def parse_npy_header(blob):
    """(dtype, shape, fortran_order, data_offset) of a ``.npy`` blob - the header format of numpy.lib.format
    v1.0/2.0/3.0: magic ``\\x93NUMPY``, version, little-endian header length, python-literal dict."""
    if len(blob) < 10 or bytes(blob[:6]) != b'\x93NUMPY':
        raise ValueError('not a .npy blob')
    major = blob[6]
    if major == 1:
        hlen = struct.unpack_from('<H', blob, 8)[0]
        start = 10
    elif major in (2, 3):
        hlen = struct.unpack_from('<I', blob, 8)[0]
        start = 12
    else:
        raise ValueError('unsupported .npy version {}'.format(major))
    text = bytes(blob[start:start + hlen]).decode('latin1' if major < 3 else 'utf8')
    meta = ast.literal_eval(text)
    return np.dtype(meta['descr']), tuple(meta['shape']), bool(meta['fortran_order']), start + hlen


class NdarrayCodec(DataframeColumnCodec):
    """ndarray <-> ``.npy`` bytes (petastorm/codecs.py:133-171)."""

    def encode(self, unischema_field, value):
        _check_ndarray(unischema_field, value)
        memfile = BytesIO()
        np.save(memfile, value)
        return bytearray(memfile.getvalue())

    def decode(self, unischema_field, value):
        return np.load(BytesIO(value))

    def spark_dtype(self):
        from petastorm_b200 import spark_types
        return spark_types.BinaryType()

    def __str__(self):
        return '{}()'.format(type(self).__name__)


class CompressedNdarrayCodec(DataframeColumnCodec):
    """ndarray <-> ``.npz`` (zip+deflate) bytes (petastorm/codecs.py:174-212)."""

    def encode(self, unischema_field, value):
        _check_ndarray(unischema_field, value)
        memfile = BytesIO()
        np.savez_compressed(memfile, arr=value)
        return bytearray(memfile.getvalue())

    def decode(self, unischema_field, value):
        return np.load(BytesIO(value))['arr']

    def spark_dtype(self):
        from petastorm_b200 import spark_types
        return spark_types.BinaryType()

    def __str__(self):
        return '{}()'.format(type(self).__name__)


class CompressedImageCodec(DataframeColumnCodec):
    """png / jpeg compressed images (petastorm/codecs.py:58-130)."""

    def __init__(self, image_codec='png', quality=80):
        self._image_codec = '.' + image_codec
        self._quality = quality

    @property
    def image_codec(self):
        return self._image_codec[1:]

    def encode(self, unischema_field, value):
        assert OPENCV_AVAILABLE, 'CompressedImageCodec.encode requires opencv-python'
        if unischema_field.numpy_dtype != value.dtype:
            raise ValueError('Unexpected type of {} feature, expected {}, got {}'.format(
                unischema_field.name, unischema_field.numpy_dtype, value.dtype))
        if not _is_compliant_shape(value.shape, unischema_field.shape):
            raise ValueError('Unexpected dimensions of {} feature, expected {}, got {}'.format(
                unischema_field.name, unischema_field.shape, value.shape))
        if value.ndim == 2:
            bgr_or_gray = value
        elif value.ndim == 3 and value.shape[2] == 3:
            bgr_or_gray = value[:, :, (2, 1, 0)]  # OpenCV wants BGR
        else:
            raise ValueError('Unexpected image dimensions. Supported dimensions are (H, W) or (H, W, 3). '
                             'Got {}'.format(value.shape))
        _, contents = cv2.imencode(self._image_codec, bgr_or_gray, [int(cv2.IMWRITE_JPEG_QUALITY), self._quality])
        return bytearray(contents)

    def decode(self, unischema_field, value):
        """Host decode of ONE value - reference contract; the readers use the batched device kernels instead."""
        assert OPENCV_AVAILABLE, 'CompressedImageCodec.decode requires opencv-python'
        img = cv2.imdecode(np.frombuffer(value, dtype=np.uint8), cv2.IMREAD_UNCHANGED)
        if img.ndim == 2:
            return img
        if img.ndim == 3 and img.shape[2] == 3:
            return img[:, :, (2, 1, 0)]
        raise ValueError('Unexpected image dimensions. Supported dimensions are (H, W) or (H, W, 3). '
                         'Got {}'.format(img.shape))

    def spark_dtype(self):
        from petastorm_b200 import spark_types
        return spark_types.BinaryType()

    def __str__(self):
        return "{}('{}', {})".format(type(self).__name__, self.image_codec, self._quality)


class ScalarCodec(DataframeColumnCodec):
    """Scalars stored in a native parquet type (petastorm/codecs.py:215-271)."""

    def __init__(self, spark_type):
        self._spark_type = spark_type

    def encode(self, unischema_field, value):
        unsized = isinstance(value, np.ndarray) and value.shape == ()
        if not unsized and hasattr(value, '__len__') and not isinstance(value, str):
            raise TypeError("Expected a scalar as a value for field '{}'. Got a non-numpy type'{}'".format(
                unischema_field.name, type(value)))
        if unischema_field.shape:
            raise ValueError("The shape field of unischema_field '%s' must be an empty tuple (i.e. '()' to indicate a "
                             "scalar. However, the actual shape is %s" % (unischema_field.name, unischema_field.shape))
        tname = type(self._spark_type).__name__
        if tname in ('ByteType', 'ShortType', 'IntegerType', 'LongType'):
            return int(value)
        if tname in ('FloatType', 'DoubleType'):
            return float(value)
        if tname == 'BooleanType':
            return bool(value)
        if tname == 'StringType':
            if not isinstance(value, str):
                raise ValueError('Expected a string value for field {}. Got type {}'.format(
                    unischema_field.name, type(value)))
            return str(value)
        return value

    def decode(self, unischema_field, value):
        return unischema_field.numpy_dtype(value)

    def spark_dtype(self):
        return self._spark_type

    def __str__(self):
        return '{}({}())'.format(type(self).__name__, type(self._spark_type).__name__)
