"""Drop-in for the reference's Periodic.py (same function names, arguments and return values),
computed by libkmcb200 on the GPU.  Reference lines are cited per function."""
import numpy as np

from . import runtime


def _gv():
    return runtime.constants()


def PutInBox(x):
    """Periodic.py:8-20 -- wrap one x coordinate into (-L/2, L/2] (Python float modulo, then fold)."""
    R = np.array([float(x), 0.0])
    runtime.context().put_all_in_box(R)
    return float(R[0])


def PutAllInBox(R):
    """Periodic.py:22-25 -- wraps the x entries of [x0,y0,x1,y1,...] IN PLACE and returns the same array."""
    if isinstance(R, np.ndarray) and R.dtype == np.float64 and R.flags.c_contiguous:
        if R.size:
            runtime.context().put_all_in_box(R)
        return R
    tmp = np.ascontiguousarray(R, dtype=np.float64).copy()
    if tmp.size:
        runtime.context().put_all_in_box(tmp)
    try:
        R[:] = tmp            # keep the reference's in-place contract for lists / other dtypes
    except TypeError:
        return tmp
    return R


def Displacement(Ri, Rj):
    """Periodic.py:27-36 -- vector(s) from Ri to Rj, x entries min-image wrapped."""
    Ri = np.asarray(Ri, dtype=np.float64)
    Rj = np.asarray(Rj, dtype=np.float64)
    if Ri.size == 0:
        return np.zeros(0)
    return runtime.context().displacement(Ri, Rj)


def Distance(Ri, Rj):
    """Periodic.py:38-48"""
    return float(runtime.context().distances(np.asarray(Ri, dtype=np.float64), np.asarray(Rj, dtype=np.float64))[0, 0])


def Displacements(x, y, Rjs):
    """Periodic.py:50-52 -- displacements from (x, y) to every atom of Rjs."""
    Rjs = np.asarray(Rjs, dtype=np.float64)
    d = np.array([x, y] * (len(Rjs) // 2), dtype=np.float64)
    return Displacement(d, Rjs)


def Distances(a, b):
    """Periodic.py:54-62 -- len(a)/2 x len(b)/2 matrix of min-image distances."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    na, nb = a.size // 2, b.size // 2
    if na == 0:
        return np.array([])
    if nb == 0:
        return np.zeros((na, 0))
    return runtime.context().distances(a, b)
