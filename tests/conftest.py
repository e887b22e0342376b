import sys
from pathlib import Path

import pytest

import os

ROOT = Path(__file__).resolve().parents[1]
# The reference's baseline_model.pt does not travel with this repository; the hierarchical tasks refuse to start without it
# unless told otherwise (controller.make_controller).  The tests check shapes, wiring and parity of the env layer BELOW
# the controller, so they opt in to the seeded random-init controller; test_host_layer_cpu checks the refusal itself.
os.environ.setdefault("ACS_ALLOW_RANDOM_CONTROLLER", "1")
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
