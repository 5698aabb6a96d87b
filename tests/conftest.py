import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Make sure the oracle and the CUDA library are built (no-op when up to date)."""
    import subprocess
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    if not (ROOT / "pedfac_b200" / "lib" / "libpedfac_cuda.so").exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "pedfac_b200" / "csrc")], check=True)
