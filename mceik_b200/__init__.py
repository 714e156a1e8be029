"""mceik_b200 -- B200 (sm_100a) implementation of mceik's forward-model / location hot path.

The product is the CUDA library ``mceik_b200/lib/libmceik_b200.so`` (C ABI in
``include/mceik_b200.h``).  This package is the thin host-side mirror of the reference's
operator interface for that path:

* :mod:`mceik_b200.eikonal` -- ``eikonal3d_serial_driver``, ``eikonal3d_initialize/solve/finalize``
  (reference fsm3d.f90) and the batched :class:`~mceik_b200.eikonal.EikonalSolver`.
* :mod:`mceik_b200.locate`  -- ``locate_l2_gridSearch__double64/float64``, ``locate_minLoc*``
  (reference locate.c), ``locate3d_gridsearch__double64/float64`` (gridsearch.f90),
  ``locate3d_initialize/gridsearch/finalize`` (locate.f90) and the batched
  :class:`~mceik_b200.locate.Locator`.
* :mod:`mceik_b200.sharding` -- how fields and events are partitioned over one-process-per-GPU ranks.

PyTorch is used only for device memory, streams and ``torch.distributed``.
"""
from ._lib import MceikError, kernel_launch_count, load  # noqa: F401
from .context import Context  # noqa: F401

__all__ = ["Context", "MceikError", "kernel_launch_count", "load"]
