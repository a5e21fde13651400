"""Importable alias of the package directory ``dynamic-diffuse-global-illumination-minecraft_b200``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("dynamic-diffuse-global-illumination-minecraft_b200")

capi = _pkg.capi
sharding = _pkg.sharding
RVPT = _pkg.RVPT
Camera = _pkg.Camera
DDGIError = _pkg.DDGIError
probe_row_shard = _pkg.probe_row_shard
