import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ORACLE_HOST_LIB = os.path.join(ROOT, "oracle", "_build", "libkml_host_oracle.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle_kml.so")
CUDA_LIB = os.path.join(ROOT, "karamelo_b200", "lib", "libkml.so")
CUDA_HOST_LIB = os.path.join(ROOT, "karamelo_b200", "lib", "libkml_host.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle behind the host driver (test infrastructure; built on demand with g++)."""
    if not os.path.exists(ORACLE_HOST_LIB) or not os.path.exists(ORACLE_LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    from karamelo_b200.api import load_host_library
    return load_host_library(ORACLE_HOST_LIB)


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; never falls back to anything else."""
    from karamelo_b200.api import load_host_library
    return load_host_library(None)
