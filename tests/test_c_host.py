"""The drop-in boundary from a non-Python host: a C program using only the six reference symbols links against
libnbody_b200.so (CPU: compile + link; GPU: run it)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_host", "host.c")
EXE = os.path.join(ROOT, "tests", "c_host", "host")


def build():
    libdir = os.path.join(ROOT, "rust_exp_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L", libdir,
                           "-lnbody_b200", f"-Wl,-rpath,{libdir}"])
    return EXE


def test_c_host_compiles_and_links_against_the_six_symbols():
    exe = build()
    out = subprocess.check_output(["nm", "-u", exe], text=True)
    used = sorted(l.split()[-1] for l in out.splitlines() if " nb" in l)
    assert used == sorted(["nb_draw", "nb_num_particles", "nb_random_disk", "nb_stable_orbits", "nb_step_barnes_hut",
                           "nb_step_brute_force"])


@pytest.mark.gpu
def test_c_host_runs_the_reference_frame_loop():
    exe = build()
    p = subprocess.run([exe, "5"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["cross_ok"] == 1 and d["lit_pixels"] > 1000 and d["after_random_disk"] == 2000


# ---- the reference's link contract: ONE static archive, host module defines nb_* itself (INTEGRATION.md option A) ----
SHIM_SRC = os.path.join(ROOT, "tests", "c_host", "shim_host.c")
SHIM_EXE = os.path.join(ROOT, "tests", "c_host", "shim_host")
CUDA_LIB = os.environ.get("CUDA_LIB_DIR", "/usr/local/cuda/lib64")


def build_shim():
    libdir = os.path.join(ROOT, "rust_exp_b200")
    assert os.path.exists(os.path.join(libdir, "libnbody_b200.a")), "libnbody_b200.a is not built (make -C rust_exp_b200/csrc)"
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SHIM_SRC, "-o", SHIM_EXE,
                           os.path.join(libdir, "libnbody_b200.a"), "-L", CUDA_LIB, "-lcudart_static", "-lstdc++", "-lm", "-ldl",
                           "-lrt", "-lpthread"])
    return SHIM_EXE


def test_static_archive_links_under_a_host_that_defines_the_nb_symbols():
    """No duplicate-symbol error: the archive's unprefixed aliases live in nb_alias.o, which is not pulled in."""
    exe = build_shim()
    out = subprocess.check_output(["nm", "--defined-only", exe], text=True)
    defined = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in ("nb_num_particles", "nb_stable_orbits", "nb_step_barnes_hut", "nb_draw", "b200_nb_step_barnes_hut", "b200_nb_draw"):
        assert n in defined
    assert "nb_set_particles" not in defined   # only nb_alias.o defines it, and nb_alias.o was not linked
    # no dynamic dependency on the shared library: this binary carries the archive
    assert "libnbody_b200" not in subprocess.check_output(["ldd", exe], text=True)


@pytest.mark.gpu
def test_static_shim_host_runs_and_step_calls_are_synchronous():
    exe = build_shim()
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["cross_ok"] == 1 and d["lit_pixels"] > 1000 and d["particles"] == 10000
    # the wall-clock timer around the step call sees the GPU time of the step (> 50 us), not a 5 us enqueue
    assert d["last_step_wall_ms"] > 0.05
