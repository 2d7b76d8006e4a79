"""inferno-autoscaler_b200 — B200-native Analyze -> Optimize hot path of the
Workload-Variant-Autoscaler (llm-d-incubation/inferno-autoscaler).

The directory name carries a hyphen (fixed by the project layout); import it through the
repo-root helper:

    import wva_import; wva = wva_import.load()      # module name: inferno_autoscaler_b200

Sub-modules: abi (ctypes mirror of include/wva_b200.h), image (SystemSpec -> SoA image),
synth (BASELINE.json workloads), binding (libwva_b200.so loader + Context).  There is no
CPU implementation in this package: binding raises when the CUDA library is missing.
"""
from . import abi, image, synth  # noqa: F401
from .image import SystemImage  # noqa: F401

__all__ = ["abi", "image", "synth", "SystemImage"]
