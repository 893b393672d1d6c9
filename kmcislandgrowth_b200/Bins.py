"""Drop-in for the reference's Bins.py: the cell list lives on the GPU.

`PutInBins` returns a `DeviceBins` handle; the reference driver treats that object as opaque and
only hands it back to Bins.* / Energy.* (SURVEY.md §8b), which is all the handle needs to support.
It also pickles (the reference ships `(positions, substrate_bins)` through multiprocessing.Pool,
2DGrowth.py:121-122) and can be indexed `bins[nx][ny]` like the reference's ragged lists.
"""
import os

import numpy as np

from . import runtime


def _gv():
    return runtime.constants()


class DeviceBins(object):
    """Bins.py:24-42 -- atoms bucketed by BinIndex, column-major [nx][ny], input order kept."""

    def __init__(self, R):
        self.R = np.ascontiguousarray(R, dtype=np.float64).copy()
        self._h = None
        self._key = None

    # -- device side (lazy, per process, rebuilt when the constants change)
    def handle(self):
        ctx = runtime.context()
        key = (os.getpid(), id(ctx), ctx.params.key())
        if self._h is None or self._key != key:
            self._h = ctx.bins_create(self.R)
            self._key = key
        return self._h

    def __reduce__(self):
        return (DeviceBins, (self.R,))

    def __len__(self):
        cell, start, sxy, sidx = runtime.context().bins_export(self.handle())
        gv = _gv()
        occ = np.nonzero(np.diff(start[:gv.nbins_x * gv.nbins_y + 1]))[0]
        return int(occ.max() // gv.nbins_y + 1) if occ.size else 0

    def __getitem__(self, nx):
        """column nx as a list of per-row lists [x, y, x, y, ...] (the reference's bins[nx])"""
        gv = _gv()
        if nx < 0 or nx >= gv.nbins_x:
            raise IndexError(nx)
        cell, start, sxy, sidx = runtime.context().bins_export(self.handle())
        col = []
        for ny in range(gv.nbins_y):
            c = nx * gv.nbins_y + ny
            col.append(list(sxy[2 * start[c]:2 * start[c + 1]]))
        while col and not col[-1]:
            col.pop()
        return col

    def cell_of_atom(self):
        return runtime.context().bins_export(self.handle())[0]


def BinIndex(x, y):
    """Bins.py:9-22 -- (int((x+L/2)/bin_size), int((y-Dmin)/bin_size)), evaluated on the device with
    IEEE add/divide and truncation, bit-identical to the reference."""
    out = runtime.context().bin_index(np.array([float(x), float(y)]))
    return (int(out[0, 0]), int(out[0, 1]))


def PutInBins(R):
    """Bins.py:24-42"""
    return DeviceBins(R)


def NearbyBins(x, y):
    """Bins.py:44-70 -- 3x3 stencil, every index wrapped once, seam column appended (host arithmetic on
    the device-computed BinIndex; integers only)."""
    gv = _gv()
    (nx, ny) = BinIndex(x, y)
    nearby_x = [nx - 1, nx, nx + 1]
    nearby_y = [ny - 1, ny, ny + 1]
    for i in range(3):
        if nearby_x[i] < 0:
            nearby_x[i] += gv.nbins_x
        if nearby_x[i] >= gv.nbins_x:
            nearby_x[i] -= gv.nbins_x
        if nearby_y[i] < 0:
            nearby_y[i] += gv.nbins_y
        if nearby_y[i] >= gv.nbins_y:
            nearby_y[i] -= gv.nbins_y
    if nx == 0:
        nearby_x.append(gv.nbins_x - 2)
    if nx == gv.nbins_x - 2:
        nearby_x.append(0)
    return (nearby_x, nearby_y)


def _as_bins(R_bins):
    return R_bins if isinstance(R_bins, DeviceBins) else DeviceBins(R_bins)


def NearbyAtoms(x, y, R_bins):
    """Bins.py:72-84 -- atoms of the stencil cells, in the reference's order."""
    out, offs = runtime.context().nearby_atoms(_as_bins(R_bins).handle(), np.array([float(x), float(y)]))
    return out


def NearestNeighbors(adatoms, substrate_bins):
    """Bins.py:86-111 -- per adatom the POSITIONS of adatoms within 1.2 r_a (itself included) and of
    substrate atoms within 1.2 r_as, as two object arrays of flat [x,y,...] arrays."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    n = adatoms.size // 2
    na_out = np.empty(n, dtype=object)
    ns_out = np.empty(n, dtype=object)
    if n == 0:
        return (na_out, ns_out)
    na, ns, (ax, ao), (sx, so) = runtime.context().neighbors(adatoms, _as_bins(substrate_bins).handle(), lists=True)
    for i in range(n):
        na_out[i] = ax[2 * ao[i]:2 * ao[i + 1]].copy()
        ns_out[i] = sx[2 * so[i]:2 * so[i + 1]].copy()
    return (na_out, ns_out)


def BondCounts(adatoms, substrate_bins):
    """(n_a incl. self, n_s) per adatom without materialising the lists (fast path of SurfaceAtoms)."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)
    return runtime.context().neighbors(adatoms, _as_bins(substrate_bins).handle())
