"""netphys_b200 -- B200-native rigid-body hot path of dougfrazer/netphys behind the engine's own API.

The product is the C-ABI shared library ``netphys_b200/libnetphys_b200.so`` (CUDA, sm_100a; sources in
``netphys_b200/csrc``, interface in ``include/netphys_b200.h``).  This package is the thin ctypes view of it
used by tests/ and bench.py, plus synthetic scenes.  There is no CPU implementation here: importing works
without a GPU (so the ABI can be inspected), creating a ``World`` without one raises.
"""
from . import scenes  # noqa: F401
from .lib import (World, CollisionData, StaticPhysicsData, BodyState, Collision, SnapshotObject, Counters,  # noqa: F401
                  NetphysError, load_library, library_path, device_count, DECLARED_SYMBOLS)
