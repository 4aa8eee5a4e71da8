"""garmentnets_b200 -- B200-native (sm_100a) implementation of the GarmentNets dense-inference hot path.

The compute lives in ``lib/libgarmentnets_b200.so`` (C-ABI, ``include/garmentnets_b200.h``); this package is the
Python host side that mirrors the reference's ``components.*`` call surface on top of it.
"""
from ._lib import GarmentNetsB200Error, LIB_PATH, load as load_library  # noqa: F401

__version__ = "0.1.0"
