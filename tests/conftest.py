import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "arrow-h264_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def built_host_libs():
    """Host-only artefacts the CPU suite needs: the generator and the oracle restatement (and, where
    /root/reference exists, the reference's own Decoder).  The CUDA engine is built by __graft_entry__.build()."""
    pkg = os.path.join(ROOT, "arrow-h264_b200")
    if not os.path.exists(os.path.join(pkg, "libh264synth.so")):
        subprocess.check_call(["make", "-C", pkg, "libh264synth.so"])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_port.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    yield
