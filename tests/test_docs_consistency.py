"""CPU: the documents name what the code contains -- every kernel of csrc/ appears in DESIGN.md's kernel
section, every C-ABI function of include/wva_b200.h in INTEGRATION.md or the header's own comments, and
the profiles README lists the files that exist."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*p):
    return open(os.path.join(ROOT, *p)).read()


def test_every_kernel_is_described_in_design():
    src = (_read("inferno-autoscaler_b200", "csrc", "wva_kernels.cuh") + _read("inferno-autoscaler_b200", "csrc", "wva_b200.cu") +
           _read("inferno-autoscaler_b200", "csrc", "wva_grid_scan.cuh") + _read("inferno-autoscaler_b200", "csrc", "wva_greedy_scan.cuh"))
    kernels = set(re.findall(r"__global__[^;{]*?\b(k_[a-z0-9_]+)\s*\(", src, flags=re.S))
    assert len(kernels) >= 30
    design = _read("DESIGN.md")
    missing = sorted(k for k in kernels if k not in design)
    assert not missing, "kernels not mentioned in DESIGN.md: %s" % missing


def test_profiles_readme_lists_existing_files():
    readme = _read("profiles", "README.md")
    names = set(re.findall(r"`([A-Za-z0-9_./]+\.(?:json|csv|txt))`", readme))
    present = set(os.listdir(os.path.join(ROOT, "profiles")))
    for n in names:
        if "*" in n or "/" in n or "..." in n:
            continue
        assert n in present, "profiles/README.md names %s which is not in profiles/" % n
    # and nothing measured sits there undocumented
    for f in present - {"README.md"}:
        stem = f.rsplit(".", 1)[0]
        assert f in readme or stem in readme or any(stem.startswith(p) for p in ("bench_r01_n", "bench_r02_multi_")), f
