"""nafwebsod_b200 -- B200 (sm_100a) implementation of NA-fWebSOD's per-proposal head.

The product is ``libnawsod.so`` (hand-written CUDA behind the C ABI of include/nawsod.h);
this package is the thin host side that mirrors the reference's operator and head-builder
interface for that path.  See DESIGN.md and INTEGRATION.md.
"""
from . import _lib  # noqa: F401
from ._lib import LIB_PATH, set_tuning  # noqa: F401

__all__ = ["_lib", "ops", "LIB_PATH", "set_tuning"]


def __getattr__(name):
    # ops / heads / dp / test_time / roi_data / loader / torch_ops / conv_body import torch; load them lazily so `import nafwebsod_b200` stays cheap
    if name in ("ops", "heads", "dp", "test_time", "roi_data", "loader", "torch_ops", "conv_body"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
