import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
  from oracle import build
  build.build()
  from oracle import ref_int
  return ref_int.lib()


@pytest.fixture(scope="session")
def cuda_lib():
  """The C-ABI library; GPU tests must run on the CUDA path or fail loudly."""
  import torch
  from snnquantprune_b200 import _lib
  assert torch.cuda.is_available(), "GPU tests need a CUDA device"
  lib = _lib.lib()
  assert lib.snnqp_device_ok() == 1, lib.snnqp_last_error().decode()
  return lib
