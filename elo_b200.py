"""Importable alias of the package directory ``efficientlo-net_b200/`` (whose name has a hyphen)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.modules[__name__] = importlib.import_module("efficientlo-net_b200")
