import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bidirectional-pathtracing_b200"))

ROOT_DIR = ROOT
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_scene(name):
    from bdpt import flatscene as fs
    return fs.load(os.path.join(GOLDEN, "scenes", name + ".fscene.gz"))


ALL_SCENES = ["CBspheres_lambertian", "CBspheres", "CBgems", "CBspheres_lambertian_occluded", "CBempty",
              "CBbunny", "CBcoil"]
