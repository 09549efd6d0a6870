"""Importable alias for the hyphenated package directory camera-intrinsic-calibration-rs_b200/."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("camera-intrinsic-calibration-rs_b200")
sys.modules[__name__] = _pkg
