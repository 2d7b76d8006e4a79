"""The C-ABI from plain C99 -- the way a cgo preamble compiles it (inferno-autoscaler_b200/host/cabi_test.c):
header is valid C, struct layouts are the ones cgo mirrors, and (GPU) the *_arrays entry points that the Go
binding calls (no struct of pointers crosses the boundary: cgo's pointer-passing rule) equal the struct forms."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "inferno-autoscaler_b200", "host", "cabi_test")


def _build():
    import __graft_entry__ as g
    g.build_cuda(); g.build_host()
    assert os.path.exists(EXE)


def test_cabi_compiles_as_c99_and_fails_loudly_without_gpu():
    import torch
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    if not torch.cuda.is_available():
        assert "layout ok; no device" in out.stdout


@pytest.mark.gpu
def test_cabi_arrays_forms_equal_struct_forms():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cabi_test: ok" in out.stdout, out.stdout


def test_go_binding_builds_no_struct_of_go_pointers():
    """go/internal/native must not pass a Go-allocated struct of Go pointers to C (cgocheck)."""
    src = open(os.path.join(ROOT, "go", "internal", "native", "native.go")).read()
    code = "\n".join(l for l in src.splitlines() if not l.strip().startswith("//"))
    assert "C.wva_alloc_soa{" not in code
    assert "C.wva_system_soa{" not in code.replace("unsafe.Sizeof(C.wva_system_soa{})", "")
    for sym in ("wva_system_upload_arrays", "wva_analyze_pairs_arrays", "wva_solve_arrays"):
        assert "C." + sym in code
