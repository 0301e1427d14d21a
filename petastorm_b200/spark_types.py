"""Light stand-ins for ``pyspark.sql.types`` classes.

Petastorm pickles its Unischema - including the ``ScalarCodec(spark_type)`` instances - into ``_common_metadata``
(petastorm/codecs.py:20-21).  Reading such a dataset must not require pyspark/JVM, so the restricted unpickler
(:mod:`petastorm_b200.etl.legacy`) resolves ``pyspark.sql.types.X`` to the classes below.  They only carry the state
the pickle holds; the read path never asks Spark anything.
"""


class DataType(object):
    def __repr__(self):
        return '{}()'.format(type(self).__name__)

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __hash__(self):
        return hash(type(self).__name__)

    @classmethod
    def typeName(cls):
        return cls.__name__[:-4].lower()


def _make(name):
    return type(name, (DataType,), {'__module__': __name__})


_NAMES = ['ByteType', 'ShortType', 'IntegerType', 'LongType', 'FloatType', 'DoubleType', 'StringType', 'BinaryType',
          'BooleanType', 'TimestampType', 'DateType', 'NullType', 'ArrayType', 'MapType', 'StructType', 'StructField',
          'AtomicType', 'NumericType', 'IntegralType', 'FractionalType']
for _n in _NAMES:
    globals()[_n] = _make(_n)


class DecimalType(DataType):
    def __init__(self, precision=10, scale=0):
        self.precision = precision
        self.scale = scale
        self.hasPrecisionInfo = True

    def __repr__(self):
        return 'DecimalType({},{})'.format(getattr(self, 'precision', '?'), getattr(self, 'scale', '?'))


def resolve(name):
    """Class for ``pyspark.sql.types.<name>`` (unknown names get a fresh stub)."""
    cls = globals().get(name)
    if cls is None:
        cls = _make(name)
        globals()[name] = cls
    return cls
