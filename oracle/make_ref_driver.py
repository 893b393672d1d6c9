#!/usr/bin/env python3
"""Write the reference's OWN driver, mechanically made runnable on Python 3, into the git-ignored baseline/_ref/
(test infrastructure; nothing of it is committed).

    python oracle/make_ref_driver.py            # needs /root/reference (the build container)

Files written: baseline/_ref/Constants.py (the reference's constants module, unchanged but for `len(x)/2`) and
baseline/_ref/RefGrowth.py (the functions of 2DGrowth.py, lines 1-303: print statements, the matplotlib import and the
forked Pool(1) replaced as documented in oracle/make_golden.py).  tests/test_dropin.py puts dropin/ AHEAD of that
directory on sys.path, so the reference driver's `import Periodic / Bins / Energy` (2DGrowth.py:12-14) resolve to the
B200 implementation while `DepositAdatom`, `HopAtom`, `Relaxation`, `HoppingRates`, ... are the reference's own code.
baseline/_ref travels to the GPU box with the repository snapshot (/root/reference does not exist there).
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)


def make(dst=None):
    import make_golden
    if not os.path.isdir(make_golden.REF):
        return None
    dst = dst or os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(dst, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="kmc_ref_py3_")
    try:
        make_golden.convert_reference(tmp)
        shutil.copyfile(os.path.join(tmp, "Constants.py"), os.path.join(dst, "Constants.py"))
        shutil.copyfile(os.path.join(tmp, "Growth.py"), os.path.join(dst, "RefGrowth.py"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return dst


if __name__ == "__main__":
    out = make()
    print("reference driver written to", out if out else "(skipped: no reference checkout)")
