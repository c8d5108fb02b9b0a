"""grid1 (src/hrweno_grids.f90) -- host-side set-up, kept on the host as in the reference.

Only the rounding-relevant parts are mirrored: edges/center/width exactly as
grid1_linear (:76-79), grid1_bilinear (:122-131), grid1_log (:171-175),
grid1_geometric (:222-225) and grid1_compute (:242-247) evaluate them.
"""
from __future__ import annotations

import numpy as np


class grid1:
    def __init__(self):
        self.ncells = 0
        self.scale = ""
        self.name = ""
        self.edges = self.center = self.width = self.left = self.right = None

    def _compute(self, xedges, name):  # grids.f90:232-250
        self.ncells = len(xedges) - 1
        self.edges = np.ascontiguousarray(xedges, dtype=np.float64)
        self.left = self.edges[:-1]
        self.right = self.edges[1:]
        self.center = (self.left + self.right) / 2
        self.width = self.right - self.left
        self.name = name

    def linear(self, xmin, xmax, ncells, name=""):  # grids.f90:41-84
        if xmax <= xmin:
            raise ValueError("Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin.")
        if ncells < 1:
            raise ValueError("Invalid input 'ncells'. Valid range: ncells > 1.")
        rx = (xmax - xmin) / ncells
        self.scale = "linear"
        self._compute(xmin + rx * np.arange(ncells + 1, dtype=np.float64), name)
        return self

    def bilinear(self, xmin, xcross, xmax, ncells, name=""):  # grids.f90:86-136
        if xcross <= xmin:
            raise ValueError("Invalid input 'xmin', 'xcross'. Valid range: xcross > xmin.")
        if xmax <= xcross:
            raise ValueError("Invalid input 'xcross', 'xmax'. Valid range: xmax > xcross.")
        if min(ncells) < 1:
            raise ValueError("Invalid input 'ncells'. Valid range: ncells(i) >= 1.")
        rx1 = (xcross - xmin) / ncells[0]
        rx2 = (xmax - xcross) / ncells[1]
        e1 = xmin + rx1 * np.arange(ncells[0] + 1, dtype=np.float64)
        e2 = xcross + rx2 * np.arange(1, ncells[1] + 1, dtype=np.float64)
        self.scale = "bilinear"
        self._compute(np.concatenate([e1, e2]), name)
        return self

    def log(self, xmin, xmax, ncells, name=""):  # grids.f90:138-180
        if xmin <= 0.0:
            raise ValueError("Invalid input 'xmin'. Valid range: xmin > 0.")
        if not xmax > xmin:
            raise ValueError("Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin.")
        if ncells < 1:
            raise ValueError("Invalid input 'ncells'. Valid range: ncells > 1.")
        rx = np.log(xmax / xmin) / ncells
        self.scale = "log"
        self._compute(np.exp(np.log(xmin) + rx * np.arange(ncells + 1, dtype=np.float64)), name)
        return self

    def geometric(self, xmin, xmax, ratio, ncells, name=""):  # grids.f90:182-230
        if xmax <= xmin:
            raise ValueError("Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin")
        if ratio <= 0.0:
            raise ValueError("Invalid input 'ratio'. Valid range: ratio > 0")
        if ncells < 1:
            raise ValueError("Invalid input 'ncells'. Valid range: ncells > 1")
        a = (xmax - xmin) / (float(_powi(ratio, np.array([ncells]))[0]) - 1.0)
        self.scale = "geometric"
        self._compute(xmin + a * (_powi(ratio, np.arange(ncells + 1)) - 1), name)
        return self


def _powi(x, m):
    """real**integer as gfortran evaluates `ratio**i` (grids.f90:223-225): binary exponentiation, low bit first
    (libgcc __powidf2 / libgfortran pow_r8_i4), vectorised over the non-negative integer exponents m"""
    n = np.asarray(m, dtype=np.int64).copy()
    y = np.where(n & 1, float(x), 1.0)
    xx = float(x)
    n >>= 1
    while np.any(n):
        xx = xx * xx
        y = np.where(n & 1, y * xx, y)
        n >>= 1
    return y
