"""grid1 (src/hrweno_grids.f90) -- host-side set-up, kept on the host as in the reference.

Only the rounding-relevant parts are mirrored: edges/center/width exactly as
grid1_linear (:76-79), grid1_bilinear (:122-131), grid1_log (:171-175),
grid1_geometric (:222-225) and grid1_compute (:242-247) evaluate them.
"""
from __future__ import annotations

import numpy as np


class grid1:
    """`grid1()` evaluates in real64; `grid1(np.float32)` is the type of the reference's REAL32 build (hrweno_kinds.F90:9-10):
    every scalar and array of the set-up is a float32 and every operation rounds to float32."""

    def __init__(self, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("grid1: real64 or real32")
        self.ncells = 0
        self.scale = ""
        self.name = ""
        self.edges = self.center = self.width = self.left = self.right = None

    def _compute(self, xedges, name):  # grids.f90:232-250
        self.ncells = len(xedges) - 1
        self.edges = np.ascontiguousarray(xedges, dtype=self.dtype)
        self.left = self.edges[:-1]
        self.right = self.edges[1:]
        self.center = (self.left + self.right) / self.dtype.type(2)
        self.width = self.right - self.left
        self.name = name

    def linear(self, xmin, xmax, ncells, name=""):  # grids.f90:41-84
        if xmax <= xmin:
            raise ValueError("Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin.")
        if ncells < 1:
            raise ValueError("Invalid input 'ncells'. Valid range: ncells > 1.")
        T = self.dtype.type
        xmin, xmax = T(xmin), T(xmax)
        rx = (xmax - xmin) / T(ncells)
        self.scale = "linear"
        self._compute(xmin + rx * np.arange(ncells + 1, dtype=self.dtype), name)
        return self

    def bilinear(self, xmin, xcross, xmax, ncells, name=""):  # grids.f90:86-136
        if xcross <= xmin:
            raise ValueError("Invalid input 'xmin', 'xcross'. Valid range: xcross > xmin.")
        if xmax <= xcross:
            raise ValueError("Invalid input 'xcross', 'xmax'. Valid range: xmax > xcross.")
        if min(ncells) < 1:
            raise ValueError("Invalid input 'ncells'. Valid range: ncells(i) >= 1.")
        T = self.dtype.type
        xmin, xcross, xmax = T(xmin), T(xcross), T(xmax)
        rx1 = (xcross - xmin) / T(ncells[0])
        rx2 = (xmax - xcross) / T(ncells[1])
        e1 = xmin + rx1 * np.arange(ncells[0] + 1, dtype=self.dtype)
        e2 = xcross + rx2 * np.arange(1, ncells[1] + 1, dtype=self.dtype)
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
        T = self.dtype.type
        xmin, xmax = T(xmin), T(xmax)
        rx = np.log(xmax / xmin) / T(ncells)
        self.scale = "log"
        self._compute(np.exp(np.log(xmin) + rx * np.arange(ncells + 1, dtype=self.dtype)), name)
        return self

    def geometric(self, xmin, xmax, ratio, ncells, name=""):  # grids.f90:182-230
        if xmax <= xmin:
            raise ValueError("Invalid input 'xmin', 'xmax'. Valid range: xmax > xmin")
        if ratio <= 0.0:
            raise ValueError("Invalid input 'ratio'. Valid range: ratio > 0")
        if ncells < 1:
            raise ValueError("Invalid input 'ncells'. Valid range: ncells > 1")
        T = self.dtype.type
        xmin, xmax, ratio = T(xmin), T(xmax), T(ratio)
        a = (xmax - xmin) / (_powi(ratio, np.array([ncells]))[0] - T(1))
        self.scale = "geometric"
        self._compute(xmin + a * (_powi(ratio, np.arange(ncells + 1)) - T(1)), name)
        return self


def _powi(x, m):
    """real**integer as gfortran evaluates `ratio**i` (grids.f90:223-225): binary exponentiation, low bit first
    (libgcc __powidf2 / libgfortran pow_r8_i4), vectorised over the non-negative integer exponents m"""
    T = type(x) if isinstance(x, np.floating) else np.float64  # the kind of x: products round to it
    n = np.asarray(m, dtype=np.int64).copy()
    xx = T(x)
    y = np.where(n & 1, xx, T(1)).astype(T)
    n >>= 1
    while np.any(n):
        xx = xx * xx
        y = np.where(n & 1, y * xx, y).astype(T)
        n >>= 1
    return y
