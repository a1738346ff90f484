"""``pi_mpc.mppi`` shim: re-exports the B200 engine's MPPI class."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from mppi_playground_b200.mppi import MPPI  # noqa: E402

__all__ = ["MPPI"]
