"""Importable alias of the `slam-robot_b200` package (its directory name is not an identifier)."""
import importlib
import sys

_pkg = importlib.import_module("slam-robot_b200")
sys.modules[__name__] = _pkg
