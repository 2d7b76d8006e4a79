"""CPU-only: the C-ABI library loads and exports every symbol include/wva_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "wva_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wva_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(wva):
    import __graft_entry__ as g
    g.build_cuda()
    from inferno_autoscaler_b200 import binding
    lib = ctypes.CDLL(binding.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(binding.EXPORTS) == declared
    assert lib.wva_abi_version() == wva.abi.ABI_VERSION


def test_no_cpu_fallback_in_product_package():
    """The product package must not import or call the oracle (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "inferno-autoscaler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "libwva_oracle" not in text and "wvao_" not in text, f


def test_ctx_create_fails_loudly_without_gpu(wva):
    import torch
    if torch.cuda.is_available():
        return
    from inferno_autoscaler_b200 import binding
    try:
        binding.Context(0)
    except binding.WvaError as e:
        assert e.code == wva.abi.ECUDA
    else:
        raise AssertionError("Context() must fail without a CUDA device")
