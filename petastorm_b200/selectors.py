"""Row-group selectors (API of petastorm/selectors.py:20-100): pure set algebra over a value -> row-group index that
was stored in the dataset metadata by ``build_rowgroup_index`` (a Spark job, out of scope here)."""
import abc


class RowGroupSelectorBase(abc.ABC):
    @abc.abstractmethod
    def get_index_names(self):
        """Names of the indexes this selector needs."""

    @abc.abstractmethod
    def select_row_groups(self, index_dict):
        """Set of row-group ordinals to read."""


class SingleIndexSelector(RowGroupSelectorBase):
    """Row-groups that contain any of ``values_list`` according to one index."""

    def __init__(self, index_name, values_list):
        self._index_name = index_name
        self._values_to_select = values_list

    def get_index_names(self):
        return [self._index_name]

    def select_row_groups(self, index_dict):
        indexer = index_dict[self._index_name]
        picked = set()
        for value in self._values_to_select:
            picked |= set(indexer.get_row_group_indexes(value))
        return picked


class _Composite(RowGroupSelectorBase):
    def __init__(self, selectors):
        self._selectors = list(selectors)

    def get_index_names(self):
        names = set()
        for s in self._selectors:
            names |= set(s.get_index_names())
        return list(names)


class IntersectIndexSelector(_Composite):
    def select_row_groups(self, index_dict):
        sets = [s.select_row_groups(index_dict) for s in self._selectors]
        return set.intersection(*sets) if sets else set()


class UnionIndexSelector(_Composite):
    def select_row_groups(self, index_dict):
        out = set()
        for s in self._selectors:
            out |= s.select_row_groups(index_dict)
        return out
