"""Stub Spark SQL types (state holders only)."""


class DataType(object):
    def __repr__(self):
        return type(self).__name__ + '()'

    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __hash__(self):
        return hash(type(self).__name__)


for _n in ['ByteType', 'ShortType', 'IntegerType', 'LongType', 'FloatType', 'DoubleType', 'StringType', 'BinaryType',
           'BooleanType', 'TimestampType', 'DateType', 'NullType', 'ArrayType', 'MapType', 'StructType', 'StructField']:
    globals()[_n] = type(_n, (DataType,), {})


class DecimalType(DataType):
    def __init__(self, precision=10, scale=0):
        self.precision = precision
        self.scale = scale
        self.hasPrecisionInfo = True
