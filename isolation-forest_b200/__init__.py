"""isolation_forest_b200 -- B200-native isolation-forest engine behind the linkedin/isolation-forest surface.

The directory is called ``isolation-forest_b200`` (repo naming contract); it is imported under the module
name ``isolation_forest_b200`` through ``__graft_entry__.load_package()``.

Only the hot path lives here: ``csrc/`` (hand-written sm_100a kernels + the C ABI of include/ifb200.h,
built into ``libifb200.so``), ``_native.py`` (ctypes binding) and the host-side mirror of the reference's
Estimator / Model interface.
"""
from . import _native  # noqa: F401
from .estimators import (ExtendedIsolationForest, ExtendedIsolationForestModel, IllegalArgumentException,  # noqa: F401
                         IllegalStateException, IsolationForest, IsolationForestModel, Scored)

__all__ = ["_native", "IsolationForest", "IsolationForestModel", "ExtendedIsolationForest",
           "ExtendedIsolationForestModel", "IllegalArgumentException", "IllegalStateException", "Scored"]
