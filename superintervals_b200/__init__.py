"""superintervals_b200 -- B200-native (sm_100a CUDA) batch interval-overlap queries.

Drop-in for the build/count/search path of kcleal/superintervals:
  * C ABI:   include/c_superintervals.h + include/superintervals_b200.h
             (libsuperintervals_b200.so, built in-tree)
  * C++:     include/superintervals.hpp  (si::IntervalMap<S,T>)
  * Python:  IntervalMap (host buffers, reference Python API) and
             DeviceIndex (device-resident tensors)
There is no CPU fallback: importing the query classes requires the built library.
"""
from ._lib import SuperIntervalsError, build_library, lib  # noqa: F401
from .intervalmap import IntervalMap  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # DeviceIndex pulls in torch; load it lazily so the host API works without it
    if name == "DeviceIndex":
        from .device import DeviceIndex
        return DeviceIndex
    raise AttributeError(name)
