import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
  sys.path.insert(0, str(ROOT))


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def g1():
  import mjlab_b200
  return mjlab_b200.load_model("g1_flat")


@pytest.fixture(scope="session")
def go1():
  import mjlab_b200
  return mjlab_b200.load_model("go1_flat")
