"""Reference vectors: outputs of the UNMODIFIED Rust reference (rs-src/nbody.rs) on fixed inputs, produced by a
maintainer with tools/ref_vectors/nbody_vectors.rs (see its header; the reference cannot be built in this image).

While tests/golden/reference_vectors/ is empty these tests SKIP and parity stays "unpinned" (DESIGN.md section 2).
Once the .bin files are committed there, the oracle (CPU) and the EXACT GPU mode must reproduce every one of them
bit for bit -- that pins the oracle to the reference itself.

Independently of that, the CPU part below always checks that the harness inputs are reproducible and that the CASES
table of the Rust file and of make_inputs.py agree, so the harness cannot rot unnoticed."""
import importlib.util
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RV = os.path.join(ROOT, "tools", "ref_vectors")
GOLD = os.path.join(ROOT, "tests", "golden", "reference_vectors")


def _mk():
    spec = importlib.util.spec_from_file_location("make_inputs", os.path.join(RV, "make_inputs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _have(name):
    return os.path.exists(os.path.join(GOLD, name + ".bin"))


def _load(name):
    return np.fromfile(os.path.join(GOLD, name + ".bin"), dtype="<f4").reshape(-1, 5)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_harness_case_tables_agree_and_inputs_are_deterministic():
    m = _mk()
    rs = open(os.path.join(RV, "nbody_vectors.rs")).read()
    rust_cases = re.findall(r'\("(\w+)", "(\w+)", (-?[\d.]+), ([\d.]+), (\d+), (\d+)\)', rs)
    assert [(a, b, float(c), float(d), int(e), int(f)) for a, b, c, d, e, f in rust_cases] == [tuple(c) for c in m.CASES]
    a, b = m.inputs(), m.inputs()
    assert all(np.array_equal(bits(a[k]), bits(b[k])) for k in a)
    assert {c[0] for c in m.CASES} <= set(a)


def _cases():
    return _mk().CASES


@pytest.mark.parametrize("case", _cases(), ids=[c[1] for c in _cases()])
def test_oracle_reproduces_reference_vector(oracle, case):
    inp, out, theta, dt, steps, nthreads = case
    if not _have(out):
        pytest.skip("no reference vector committed yet (tools/ref_vectors/nbody_vectors.rs, needs cargo): parity unpinned")
    oracle.set_particles(_mk().inputs()[inp])
    for _ in range(steps):
        oracle.step_brute_force(dt) if theta < 0 else oracle.step_barnes_hut(theta, dt, nthreads)
    assert np.array_equal(bits(oracle.get_particles()), bits(_load(out)))


def test_oracle_force_law_reproduces_reference_kat(oracle):
    if not os.path.exists(os.path.join(GOLD, "force_kat.bin")):
        pytest.skip("no reference vector committed yet: parity unpinned")
    k = _mk().inputs()["force_kat_pairs"]
    ref = np.fromfile(os.path.join(GOLD, "force_kat.bin"), dtype="<f4").reshape(-1, 2)
    got = np.array([oracle.force(k[i, :2], k[i, 4], k[i + 1, :2], k[i + 1, 4]) for i in range(0, len(k), 2)], dtype=np.float32)
    assert np.array_equal(bits(got), bits(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=[c[1] for c in _cases()])
def test_gpu_exact_mode_reproduces_reference_vector(fresh, case):
    from rust_exp_b200 import binding

    inp, out, theta, dt, steps, nthreads = case
    if not _have(out):
        pytest.skip("no reference vector committed yet: parity unpinned")
    fresh.set_mode(binding.MODE_EXACT)
    fresh.set_particles(_mk().inputs()[inp])
    for _ in range(steps):
        fresh.step_brute_force(dt) if theta < 0 else fresh.step_barnes_hut(theta, dt, nthreads)
    assert np.array_equal(bits(fresh.get_particles()), bits(_load(out)))
