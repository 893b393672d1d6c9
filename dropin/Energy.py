"""Drop-in shim: put this directory on sys.path AHEAD of the reference checkout and the reference
driver's `import Energy` resolves to the B200 implementation."""
from kmcislandgrowth_b200.Energy import *  # noqa: F401,F403
from kmcislandgrowth_b200.Energy import _gv  # noqa: F401
