import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _compile_example(tmp_path, name, extra=()):
    """examples/<name>.c compiled with gcc (C, not C++) against include/oqb.h and linked to liboqb_b200.so."""
    import shutil
    import subprocess
    import oqb_b200  # noqa: F401
    from oqb_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    _lib.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(str(tmp_path), name)
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", name + ".c"),
           "-L", libdir, "-loqb_b200", f"-Wl,-rpath,{libdir}", "-lm", *extra, "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


@pytest.fixture
def c_demo_exe(tmp_path):
    return _compile_example(tmp_path, "c_abi_demo")


@pytest.fixture
def c_liouv_exe(tmp_path):
    """examples/c_liouv_demo.c: the AME RHS end to end through C only (oqb_liouv_*)."""
    return _compile_example(tmp_path, "c_liouv_demo")


@pytest.fixture
def c_comm_exe(tmp_path):
    """examples/c_comm_demo.c: one process, one context per GPU, NCCL ensemble reduce through the C ABI."""
    return _compile_example(tmp_path, "c_comm_demo", extra=("-lpthread",))
