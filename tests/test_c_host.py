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
