"""Drop-in ``pi_mpc`` package (mirrors src/pi_mpc/__init__.py:1-4 of the reference).

Put this directory's parent (``mppi_playground_b200/compat``) ahead of the
reference's ``src`` on PYTHONPATH and ``from pi_mpc.mppi import MPPI`` /
``from pi_mpc import MPPI`` in example/*.py picks up the B200 engine.
"""
from .mppi import MPPI

__all__ = ["MPPI"]
