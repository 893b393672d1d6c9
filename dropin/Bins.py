"""Drop-in shim: put this directory on sys.path AHEAD of the reference checkout and the reference
driver's `import Bins` resolves to the B200 implementation."""
from kmcislandgrowth_b200.Bins import *  # noqa: F401,F403
from kmcislandgrowth_b200.Bins import _gv  # noqa: F401
