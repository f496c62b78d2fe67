"""The C ABI from plain C: examples/c_abi_gas_cell.c compiles against include/pyrad_b200.h with gcc, links the in-tree
shared library, and -- on a GPU box -- runs the gas-cell call sequence without Python in the loop.  Without a device
the program must fail loudly with PRB_ERR_NODEVICE (exit code 77), never compute anything."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = os.path.join(str(tmp_path), "c_abi_gas_cell")
    libdir = os.path.join(ROOT, "pyrad_b200")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_gas_cell.c"), "-o", exe, "-L", libdir, "-lpyrad_b200",
           "-Wl,-rpath," + libdir, "-lm"]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_c_example_compiles_links_and_refuses_to_run_without_a_device(tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 77 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    out = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "c_abi_gas_cell ok" in out.stdout and "pairs" in out.stdout
    assert "2 rows in one pass" in out.stdout                      # the grouped API from C pointer tables
