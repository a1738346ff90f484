"""mppi_playground_b200 - B200-native MPPI rollout engine behind the
``pi_mpc.MPPI`` surface of kohonda/mppi_playground.

    from mppi_playground_b200 import MPPI          # same ctor / forward as pi_mpc.mppi.MPPI

or, to run the reference's examples unchanged, put ``mppi_playground_b200/compat``
first on PYTHONPATH: it provides a ``pi_mpc`` package that re-exports this class.

The package holds only what the solve path needs: ``csrc/`` (CUDA kernels and
the C ABI of include/mppi_b200.h), ``_capi`` (ctypes binding), ``mppi`` (the
host-side mirror of the reference class) and ``models`` (callable -> device
model resolution). Importing it does not load CUDA; constructing ``MPPI`` does
and raises if libmppi_b200.so or a CUDA device is missing - there is no CPU path.
"""
from .mppi import MPPI, shard_bounds  # noqa: F401
from .models import (CartpoleContinuousModel, CartpoleModel, GoalInDangerZoneModel, MountainCarModel, Navigation2DModel, PendulumModel,  # noqa: F401
                     RacingModel, RacingReferencePath, racing_reference_path)

__all__ = ["MPPI", "PendulumModel", "CartpoleModel", "MountainCarModel", "CartpoleContinuousModel", "GoalInDangerZoneModel", "Navigation2DModel", "RacingModel",
           "racing_reference_path", "RacingReferencePath", "shard_bounds"]
