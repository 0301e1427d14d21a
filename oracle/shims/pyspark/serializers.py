"""pyspark hijacks namedtuple pickling; datasets written from a Spark driver store UnischemaField this way."""
import collections


def _restore(name, fields, value):
    if name == 'UnischemaField':
        from petastorm.unischema import UnischemaField
        return UnischemaField(*value)
    return collections.namedtuple(name, fields)(*value)
