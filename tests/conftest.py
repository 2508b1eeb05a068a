import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

FIXTURES = os.path.join(ROOT, "tests", "golden", "fixtures")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build whatever is missing (the prebuilt .so files travel to the GPU box; make is then a no-op)
    need = [os.path.join(ROOT, "jpeg_rust_b200", "lib", "libjpgpu.so"),
            os.path.join(ROOT, "jpeg_rust_b200", "lib", "libjpgenc.so"),
            os.path.join(ROOT, "oracle", "liboracle.so"),
            os.path.join(ROOT, "tests", "sim", "libjpsim.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-s", "-C", ROOT, "all"])


def fixture_bytes(name):
    with open(os.path.join(FIXTURES, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def fixtures():
    return {n: fixture_bytes(n) for n in ("lena.jpeg", "2x2-chroma.jpeg", "lena-bw.jpeg", "huff_simple0.jpg")}
