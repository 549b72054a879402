"""Shared fixtures.  GPU tests are marked @pytest.mark.gpu and call the CUDA path through the C ABI;
everything else runs on CPU (oracle vs golden vectors, host logic, library symbol checks)."""
from __future__ import annotations

import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    p = importlib.import_module("raytracer-in-cpp_b200")
    p.build.build_all()
    return p


@pytest.fixture(scope="session")
def capi(pkg):
    return pkg.capi


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build_port()
    return O


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory, pkg):
    """Synthetic OBJ/MTL scenes, generated once per session (deterministic generators)."""
    d = tmp_path_factory.mktemp("scenes")
    return str(d)


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLD, name + ".npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())

    def __getitem__(self, k):
        return self.z[k]

    def has_scene(self):
        return "verts" in self.z.files


def load_golden(name):
    return Golden(name)


# which golden file carries the baked scene arrays for each case
SCENE_OF = {
    "cube_point_1000": "cube_point_1000",
    "cube_area_640x360": "cube_area_640x360",
    "cube_rot_500x400": "cube_rot_500x400",
    "cube_area_d3_g4_1920x1080_s9": "cube_point_1000",
    "dodge_point_1000": "dodge_point_1000",
    "dodge_area_rot_400x300": "dodge_point_1000",
    "gallery_small_point_320x240": "gallery_small_point_320x240",
    "gallery_area_200x150": "gallery_area_200x150",
    "gallery_area_d2_g4_200x150": "gallery_area_200x150",
    "gallery_point_d0_320x240": "gallery_area_200x150",
    "hf32_point_256x144": "hf32_point_256x144",
}
# cases whose scene is regenerated from the deterministic generators and loaded by the product loader
GENERATED = {
    "hf224_point_3840x2160_s24": ("write_heightfield", (224,)),
    "hf224_area_d5_g4_3840x2160_s24": ("write_heightfield", (224,)),
    "hf224m_area_d5_g4_3840x2160_s48": ("write_heightfield", (224, 1234, 4)),
    "hf707_point_1920x1080_s20": ("write_heightfield", (707,)),
    "hf707_point_7680x4320_s80": ("write_heightfield", (707,)),
}

_scene_cache = {}


def scene_arrays(case, pkg, scene_dir):
    """(verts, fnormals, vnormals, mat_id, mats) for a golden case."""
    if case in SCENE_OF:
        g = load_golden(SCENE_OF[case])
        return g["verts"], g["fnormals"], g["vnormals"], g["mat_id"], g["mats"]
    gen, args = GENERATED[case]
    key = (gen, args)
    if key not in _scene_cache:
        path = os.path.join(scene_dir, f"{gen}_{'_'.join(map(str, args))}.obj")
        if not os.path.exists(path):
            getattr(pkg.scenes, gen)(path, *args)
        mesh = pkg.capi.Mesh(path)
        _scene_cache[key] = mesh.arrays()
        mesh.close()
    return _scene_cache[key]


def case_params(g):
    m = g.meta
    md = m["max_depth"] if m["max_depth"] is not None else -1
    grid = tuple(m["grid"]) if m["grid"] else (5, 5)
    return dict(w=m["w"], h=m["h"], area=m["area"], point=m["point"], max_depth=md, grid=grid)
