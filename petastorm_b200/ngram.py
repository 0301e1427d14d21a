"""NGram: sliding windows of consecutive rows of one row-group (API of petastorm/ngram.py:102-339).

``fields`` maps consecutive integer offsets to lists of fields; a window of ``length`` rows starting at row ``i`` is
emitted iff every consecutive ``timestamp_field`` gap inside it is ``<= delta_threshold`` (rows must be sorted by the
timestamp inside a row-group, otherwise ``NotImplementedError``), and - with ``timestamp_overlap=False`` - its first
timestamp is greater than the last timestamp of the previously emitted window.  Windows never span row-groups.

Device form: :meth:`NGram.window_starts_device` computes the gap test for all candidate starts with one kernel
(K15, ``pst_ngram_valid_starts``) and the readers gather the windows with ``pst_ngram_gather``.
"""
import numbers

import numpy as np

from petastorm_b200.unischema import UnischemaField, match_unischema_fields


class NGram(object):
    def __init__(self, fields, delta_threshold, timestamp_field, timestamp_overlap=True):
        self._validate(fields, delta_threshold, timestamp_field, timestamp_overlap)
        self._fields = fields
        self._delta_threshold = delta_threshold
        self._timestamp_field = timestamp_field
        self.timestamp_overlap = timestamp_overlap

    @staticmethod
    def _validate(fields, delta_threshold, timestamp_field, timestamp_overlap):
        if fields is None or not isinstance(fields, dict):
            raise ValueError('Fields must be set and must be a dictionary.')
        for value in fields.values():
            if not isinstance(value, list):
                raise ValueError('Each field value must be a list of unischema fields/regular expression(s)')
            for f in value:
                if not isinstance(f, (UnischemaField, str)):
                    raise ValueError('All field values must be of type UnischemaField/regular expression')
        if delta_threshold is None or not isinstance(delta_threshold, numbers.Number):
            raise ValueError('delta_threshold must be a number.')
        if timestamp_field is None or not isinstance(timestamp_field, (UnischemaField, str)):
            raise ValueError('timestamp_field must be set and must be of type UnischemaField or regular expression')
        if timestamp_overlap is None or not isinstance(timestamp_overlap, bool):
            raise ValueError('timestamp_overlap must be set and must be of type bool')

    # ---- properties ---------------------------------------------------------------------------------------------
    @property
    def length(self):
        return max(self._fields.keys()) - min(self._fields.keys()) + 1

    @property
    def fields(self):
        return self._fields

    @property
    def delta_threshold(self):
        return self._delta_threshold

    @property
    def timestamp_field(self):
        return self._timestamp_field

    @property
    def base_key(self):
        return min(self._fields.keys())

    # ---- schema helpers -----------------------------------------------------------------------------------------
    def convert_fields(self, unischema, field_list):
        """UnischemaField objects stay, strings are full-match regexes over the schema (petastorm/ngram.py:303-326)."""
        patterns = [f for f in field_list if isinstance(f, str)]
        objects = [f for f in field_list if isinstance(f, tuple)]
        if len(patterns) + len(objects) != len(field_list):
            raise ValueError('"Elements of fields"/"timestamp field" must be either a string (regular expressions) or'
                             ' an instance of UnischemaField class.')
        return objects + match_unischema_fields(unischema, patterns)

    def resolve_regex_field_names(self, schema):
        self._fields = {k: self.convert_fields(schema, v) for k, v in self._fields.items()}
        ts = self.convert_fields(schema, [self._timestamp_field])
        if len(ts) > 1:
            raise ValueError('timestamp_field was matched to more than one unischema field')
        self._timestamp_field = ts[0]

    def get_field_names_at_timestep(self, timestep):
        if timestep not in self._fields:
            return []
        return [f.name for f in self._fields[timestep]]

    def get_field_names_at_all_timesteps(self):
        return list({f for fields in self._fields.values() for f in fields})

    def get_field_names_at_all_timesteps_names(self):
        """Names (not field objects) of every field that appears at any timestep."""
        return sorted({f.name for fields in self._fields.values() for f in fields})

    def get_schema_at_timestep(self, schema, timestep):
        names = self.get_field_names_at_timestep(timestep)
        return schema.create_schema_view([schema.fields[n] for n in schema.fields if n in names])

    def make_namedtuple(self, schema, ngram_as_dicts):
        return {t: self.get_schema_at_timestep(schema, t).make_namedtuple(**row) for t, row in ngram_as_dicts.items()}

    # ---- window formation ---------------------------------------------------------------------------------------
    def _apply_no_overlap(self, starts, ts_host):
        """Sequential rule of ``timestamp_overlap=False`` (petastorm/ngram.py:248-253,266-268) on candidate starts."""
        keep, prev_end = [], None
        for s in starts:
            if prev_end is not None and ts_host[s] <= prev_end:
                continue
            keep.append(s)
            prev_end = ts_host[s + self.length - 1]
        return keep

    def window_starts_device(self, ts_tensor):
        """Ascending start rows (CUDA int64 tensor) of all valid windows of a row-group whose timestamp column is the
        CUDA int64 tensor ``ts_tensor``.  Raises NotImplementedError if the timestamps are not sorted."""
        import torch
        from petastorm_b200 import device_ops
        if ts_tensor.numel() < self.length:
            return torch.empty(0, dtype=torch.int64, device=ts_tensor.device)
        delta = self._delta_threshold
        # gap <= delta on integers: floor(delta) is equivalent for any real threshold
        delta_i = int(np.floor(float(delta)))
        ok, status = device_ops.ngram_valid_starts(ts_tensor.to(torch.int64), self.length, delta_i)
        starts = device_ops.mask_to_indices(ok)
        if int(status[0].item()) != 0:
            raise NotImplementedError('NGram assumes that the data is sorted by {0} field which is not the case'
                                      .format(self._timestamp_field.name))
        if not self.timestamp_overlap and starts.numel():
            ts_host = ts_tensor.cpu().numpy()
            kept = self._apply_no_overlap(starts.cpu().tolist(), ts_host)
            starts = torch.tensor(kept, dtype=torch.int64, device=ts_tensor.device)
        return starts

    def window_starts_host(self, ts):
        """Same as :meth:`window_starts_device` for a host sequence of timestamps (any comparable type)."""
        n = len(ts)
        length = self.length
        starts = []
        for i in range(n - length + 1):
            w = ts[i:i + length]
            if any(w[k] > w[k + 1] for k in range(length - 1)):
                raise NotImplementedError('NGram assumes that the data is sorted by {0} field which is not the case'
                                          .format(self._timestamp_field.name))
            if all(w[k + 1] - w[k] <= self._delta_threshold for k in range(length - 1)):
                starts.append(i)
        if not self.timestamp_overlap:
            starts = self._apply_no_overlap(starts, ts)
        return starts

    def form_ngram(self, data, schema):
        """Reference-compatible host form: ``data`` is a list of row dicts of one row-group; returns a list of
        ``{offset: {field: value}}`` (petastorm/ngram.py:225-270)."""
        ts_name = self._timestamp_field.name
        starts = self.window_starts_host([row[ts_name] for row in data])
        base = self.base_key
        out = []
        for s in starts:
            item = {}
            for k in range(self.length):
                names = self.get_field_names_at_timestep(base + k)
                row = data[s + k]
                item[base + k] = {name: row[name] for name in row if name in names}
            out.append(item)
        return out

    def __eq__(self, other):
        if set(self.fields.keys()) != set(other.fields.keys()):
            return False
        return all(set(self.fields[k]) == set(other.fields[k]) for k in self.fields)

    def __ne__(self, other):
        return not self == other
