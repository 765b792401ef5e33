import os

import numpy as np
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
    if name == "CBbunny_microfacet_cu":
        # reference dae/sky/CBbunny_microfacet_cu.dae: byte for byte CBbunny.dae's geometry, camera and light with the
        # bunny's material (index 0) a copper MicrofacetBSDF (eta, k, alpha from the file's CGL profile) — derived here
        # instead of committing the 3.5 MB geometry twice; tests/test_host_loader.py checks the derivation against the
        # .dae where the reference tree is mounted
        sc = fs.load(os.path.join(GOLDEN, "scenes", "CBbunny.fscene.gz"))
        mats = sc.materials.copy()
        m = mats[0]
        m["type"] = fs.MAT_MICROFACET
        m["reflectance"] = 0.0
        m["eta"] = [0.33228, 1.0162, 1.2474]
        m["k"] = [3.1646, 2.5785, 2.4603]
        m["alpha"] = np.float32(0.05)
        mats[0] = m
        return fs.FlatScene(sc.camera, mats, sc.lights, sc.prims)
    return fs.load(os.path.join(GOLDEN, "scenes", name + ".fscene.gz"))


ALL_SCENES = ["CBspheres_lambertian", "CBspheres", "CBgems", "CBspheres_lambertian_occluded", "CBempty",
              "CBbunny", "CBcoil"]


def env_scene(kind="sphere", radiance=(1.0, 1.0, 1.0)):
    """Scenes with an environment light (BDPT_LIGHT_ENV; the reference's BDPT aborts on it: an extension, see
    include/bdpt_flatscene.h).  "sphere": ONE diffuse sphere (rho = 0.8) alone under the environment — convex, so the
    radiance it reflects under a constant environment L is exactly rho * L; "cornell": CBspheres with the environment
    added to its area light.  The camera's screen distance is shortened so that a 48 x 36 frame sees what 480 x 360
    does (the field of view depends on the resolution)."""
    from bdpt import flatscene as fs
    cam_scale = 10.0
    if kind == "sphere":
        sc = load_scene("CBspheres_lambertian")
        prims = sc.prims[sc.prims["type"] == 1][:1].copy()
        lights = np.zeros(1, fs.LIGHT_DT)
    else:
        sc = load_scene("CBspheres")
        prims = sc.prims
        lights = np.concatenate([sc.lights, np.zeros(1, fs.LIGHT_DT)])
    lights[-1]["type"] = fs.LIGHT_ENV
    lights[-1]["radiance"] = radiance
    cam = sc.camera.copy()
    cam["screen_dist"] = cam["screen_dist"] / cam_scale
    return fs.FlatScene(cam, sc.materials, lights, prims)
