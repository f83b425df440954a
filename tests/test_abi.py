"""The C-ABI library loads and exports every symbol include/yacht_gpu.h declares (no compute here)."""
import ctypes
import os
import re
import subprocess

import pytest

from yacht_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "yacht_gpu.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ygpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_fails_loudly_without_gpu():
    lib = _lib.load_library()
    if lib.ygpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.YgpuError, match="no CPU fallback"):
        _lib.GpuContext(0)
    exe = os.path.join(ROOT, "yacht_b200", "run_yacht_train_core")
    cp = subprocess.run([exe, "nofile", "/tmp", "/tmp/x"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert cp.returncode != 0 and "no CPU path" in cp.stderr


def test_core_cli_argument_errors():
    # reference main.cpp:172-182,430-436: bad values -> message, "Usage: <prog> -h", exit code 1
    exe = os.path.join(ROOT, "yacht_b200", "run_yacht_train_core")
    for bad in (["-t", "0"], ["-p", "0"], ["-c", "1.5"], ["-c", "-0.1"]):
        cp = subprocess.run([exe] + bad + ["a", "b", "c"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert cp.returncode == 1 and "Usage:" in cp.stdout
    cp = subprocess.run([exe, "a", "b"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert cp.returncode == 1
    cp = subprocess.run([exe, "-h"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert cp.returncode == 0 and "containment_threshold" in cp.stdout


def test_product_never_imports_oracle():
    # the product path must not route through the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "yacht_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, fn)
