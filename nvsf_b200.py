"""Import shim: ``import nvsf_b200`` -> the package directory
``selfsupervised-nvsf_b200/`` (whose name is not a valid Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("selfsupervised-nvsf_b200")
sys.modules[__name__] = _pkg
