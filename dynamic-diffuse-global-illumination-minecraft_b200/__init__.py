"""B200-native DDGI probe-field engine: the per-frame probe update and the per-pixel
8-probe-cage sample of helenl9098/Dynamic-Diffuse-Global-Illumination-Minecraft as
hand-written sm_100a CUDA behind a C-ABI (include/ddgi.h).

The directory name is not a Python identifier; import it with
``importlib.import_module("dynamic-diffuse-global-illumination-minecraft_b200")`` or through
the ``ddgi_b200`` alias module at the repository root.
"""
from . import capi, sharding
from .rvpt import RVPT, Camera, DDGIError, probe_row_shard

__all__ = ["capi", "sharding", "RVPT", "Camera", "DDGIError", "probe_row_shard"]
