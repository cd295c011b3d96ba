"""Importable alias of the package directory (whose mandated name contains hyphens)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200")
sys.modules[__name__] = _pkg
