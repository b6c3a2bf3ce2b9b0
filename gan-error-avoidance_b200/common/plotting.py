"""Loss history container — the data half of the reference's common/plotting.py (:12-57, :116-185).

The reference pickles a ``History`` object into ``net_archive/*_state.pt`` (g_lis/main.py:350-357) and
unpickles it on resume (:341).  Checkpoint compatibility therefore needs classes of the same names
(``common.plotting.History`` / ``LineGroup`` / ``Line``) carrying the same attributes (``line_groups``;
``group_name, lines, increasing, xlim``; ``xs, ys, counts, datetimes, last_index`` as growing numpy arrays).
The matplotlib plotter (``LossPlotter``) is host-side visualisation and out of this path's scope.
"""
import pickle
import time
from collections import OrderedDict

import numpy as np

GROWTH_BY = 500


class Line(object):
    """One curve: parallel arrays that grow in blocks of GROWTH_BY; entries [0, last_index] are valid."""

    _DTYPES = (("xs", np.int32), ("ys", np.float32), ("counts", np.uint16), ("datetimes", np.uint64))

    def __init__(self):
        for name, dt in self._DTYPES:
            setattr(self, name, np.zeros(GROWTH_BY, dtype=dt))
        self.last_index = -1

    def _view(self, name):
        return getattr(self, name)[:self.last_index + 1]

    def get_xs(self):
        return self._view("xs")

    def get_ys(self):
        return self._view("ys")

    def get_counts(self):
        return self._view("counts")

    def get_datetimes(self):
        return self._view("datetimes")

    def append(self, x, y, average=False):
        """New point (x, y); with ``average`` a repeated x folds into a running mean of its ys."""
        if x is None or y is None:
            raise ValueError("Line.append needs x and y")
        now = int(time.time() * 1000)
        i = self.last_index
        if average and i >= 0 and self.xs[i] == x:
            n = int(self.counts[i])
            self.ys[i] = (self.ys[i] * n + y) / (n + 1)
            self.counts[i] = n + 1
            self.datetimes[i] = now
            return
        if i + 1 == self.xs.shape[0]:
            for name, dt in self._DTYPES:
                setattr(self, name, np.concatenate([getattr(self, name), np.zeros(GROWTH_BY, dtype=dt)]))
        i += 1
        self.xs[i], self.ys[i], self.counts[i], self.datetimes[i] = x, y, 1, now
        self.last_index = i


class LineGroup(object):
    def __init__(self, group_name, line_names, increasing=True):
        self.group_name = group_name
        self.lines = OrderedDict((name, Line()) for name in line_names)
        self.increasing = increasing
        self.xlim = (None, None)

    def get_line_names(self):
        return list(self.lines.keys())

    def get_line_xs(self):
        return [line.get_xs() for line in self.lines.values()]

    def get_line_ys(self):
        return [line.get_ys() for line in self.lines.values()]

    def get_max_x(self):
        return max([int(line.get_xs().max()) if line.last_index > -1 else 0 for line in self.lines.values()] or [0])


class History(object):
    def __init__(self):
        self.line_groups = OrderedDict()

    @staticmethod
    def from_string(s):
        """Unpickle a history; accepts python-2 pickles of the reference's own classes (latin1 strings)."""
        if isinstance(s, str):
            s = s.encode("latin1")
        return pickle.loads(s, encoding="latin1")

    def to_string(self):
        return pickle.dumps(self, protocol=2)     # the highest protocol python 2.7 reads

    def add_group(self, group_name, line_names, increasing=True):
        grp = self.line_groups.get(group_name)
        if grp is None:
            self.line_groups[group_name] = LineGroup(group_name, line_names, increasing=increasing)
            return
        grp.increasing = increasing
        for name in line_names:
            grp.lines.setdefault(name, Line())

    def add_value(self, group_name, line_name, x, y, average=False):
        self.line_groups[group_name].lines[line_name].append(x, y, average=average)

    def get_group_names(self):
        return list(self.line_groups.keys())

    def get_groups_increasing(self):
        return [g.increasing for g in self.line_groups.values()]

    def get_max_x(self):
        return max([g.get_max_x() for g in self.line_groups.values()] or [0])

    def get_recent_average(self, group_name, line_name, nb_points):
        return float(np.average(self.line_groups[group_name].lines[line_name].get_ys()[-nb_points:]))
