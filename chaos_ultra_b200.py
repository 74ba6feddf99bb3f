"""Import shim: the package directory is named ``chaos-ultra_b200`` (project name with its hyphen),
which ``import`` cannot spell.  ``import chaos_ultra_b200`` loads that package and aliases it."""
import importlib
import sys

_pkg = importlib.import_module("chaos-ultra_b200")
sys.modules[__name__] = _pkg
