"""The C++ host side (include/minimcmc.hpp) mirrors the crate's API; tests/cpp/reference_style_tests.cpp restates a
selection of the reference's own tests against it.  Compiled with g++ here; the host-only part runs on CPU, the full
program on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = tmp_path / "reference_style_tests"
    lib_dir = os.path.join(ROOT, "mini_mcmc_b200")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "reference_style_tests.cpp"), "-o", str(exe),
                    "-L", lib_dir, "-l:libminimcmc.so", f"-Wl,-rpath,{lib_dir}"], check=True)
    return str(exe)


def test_cpp_host_builds_and_host_only_part_passes(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "--host-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_reference_style_tests(cuda_device, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
