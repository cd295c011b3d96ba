"""dcl_b200 -- B200-native sliding-window ClsWiseFormer inference.

Host side of the C ABI in ``include/dcl_b200.h``.  The arithmetic lives in ``libdcl_b200.so``
(hand-written sm_100a CUDA, built from ``csrc/``); this package only marshals pointers.  There
is no CPU or PyTorch fallback: importing works anywhere, every compute call raises
``DclError`` unless the library is built and a CUDA device is present.

The directory name is fixed by the project layout and is not a Python identifier; import the
package as ``dcl_b200`` (alias module at the repository root) or through ``importlib``.
"""
import os as _os

# A volume call keeps three patch graphs in flight on separate streams, each with six coupler branches.  With the
# driver's default of 8 hardware work queues two of those streams can end up sharing a queue (false serialisation:
# roughly one process in four then ran 33-42 instead of 25.5 ms per volume); 32 queues is the maximum.  Must be set
# before the CUDA context is created; an explicit setting by the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._native import (DclError, Precision, StitchMode, abi_version, build_library, library_path, load_library,
                      weight_catalogue, workspace_bytes)
from .engine import Engine, patch_starts, reference_starts
from . import sharded
from . import volio

__all__ = ["DclError", "Precision", "StitchMode", "Engine", "abi_version", "build_library", "library_path",
           "load_library", "weight_catalogue", "workspace_bytes", "patch_starts", "reference_starts", "sharded", "volio"]
