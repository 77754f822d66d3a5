"""The C++ host layer above the C ABI (include/towr_b200_ifopt.hpp: ifopt Component API + towr::NlpFormulation
mirror), driven like towr/test/hopper_example.cc by tests/cpp/ifopt_shim_test.cc."""
import os
import subprocess

import numpy as np
import pytest

import towr_b200 as tb
from towr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "ifopt_shim_test")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "ifopt_shim_test.cc"),
                           "-o", EXE, "-L" + libdir, "-ltowr_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])


def test_c_abi_from_plain_c99():
    """include/towr_b200.h is a C header: compile and drive it from C99."""
    exe = os.path.join(os.path.dirname(EXE), "c_abi_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", os.path.join(ROOT, "tests", "cpp", "c_abi_test.c"), "-o", exe,
                           "-L" + libdir, "-ltowr_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_cpp_shim_structure_and_block_slicing():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


def test_ifopt_conformance():
    """The Gpu* views compiled against a verbatim restatement of ifopt 2.0's class declarations (external-ifopt mode of the
    header): concrete, reference signatures, driven through ifopt's own Composite / ConstraintSet / CostTerm methods."""
    exe = os.path.join(os.path.dirname(EXE), "ifopt_conformance")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "ifopt_conformance.cc"), "-o", exe,
                           "-L" + libdir, "-ltowr_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_shim_evaluation_matches_python_binding():
    _build()
    out = subprocess.run([EXE, "--gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
    vals = dict(line.split() for line in out.stdout.splitlines() if " " in line)
    p = tb.Problem(tb.make_formulation("hopper").to_spec())
    X = np.tile(p.GetVariableValues(), (3, 1))
    res = p.batch(3).eval_host(X)
    (_, r0, nr), = [c for c in p.constraint_sets() if c[0] == "dynamic"]
    assert np.isclose(float(vals["dynamic_g_sum"]), res["g"][1, r0:r0 + nr].sum(), rtol=1e-13, atol=1e-9)
    assert np.isclose(float(vals["jac_sum"]), res["jac"][1].sum(), rtol=1e-13)
    assert int(vals["status"]) == 0
