#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/ref_oracle, built from /root/reference by `make -C oracle ref`) headless.

The reference ships no golden vectors of its own, so these files are the pin:
for each case the reference's float RGB, primary face id and hit parameter t on a strided pixel
subset, plus (for small scenes) the baked scene the reference actually traced.

Run from the repository root, in the build container (needs /root/reference):
    python tests/golden/make_golden.py [case ...]
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

scenes = importlib.import_module("raytracer-in-cpp_b200.scenes")
GOLD = os.path.join(ROOT, "tests", "golden")

# name -> dict(scene=('ref', file) | ('gen', generator, args), w, h, area, point, stride, lights, cam_rot,
#              cam_trans, max_depth, grid, keep_scene)
CASES = {
    "cube_point_1000": dict(scene=("ref", "cube.obj"), w=1000, h=1000, area=0, point=1, stride=7, keep_scene=True),
    "cube_area_640x360": dict(scene=("ref", "cube.obj"), w=640, h=360, area=1, point=0, stride=3, keep_scene=True),
    "cube_rot_500x400": dict(scene=("ref", "cube.obj"), w=500, h=400, area=0, point=1, stride=3,
                             lights=[(-1, 1, 1), (2, 0.5, 1.5)], cam_rot=(-0.4, 0.7), cam_trans=(0.2, 0.1, -0.3),
                             keep_scene=True),
    "dodge_point_1000": dict(scene=("ref", "dodgeColorTest.obj"), w=1000, h=1000, area=0, point=1, stride=7,
                             keep_scene=True),
    "dodge_area_rot_400x300": dict(scene=("ref", "dodgeColorTest.obj"), w=400, h=300, area=1, point=0, stride=3,
                                   lights=[(-1, 1, 1), (0.5, 2, 1.5)], cam_rot=(0.3, -0.5),
                                   cam_trans=(0.1, -0.2, 0.3), keep_scene=False),
    "gallery_small_point_320x240": dict(scene=("gen", "write_gallery", (1,)), w=320, h=240, area=0, point=1,
                                        stride=2, lights=[(-1, 1.2, 1.5)], keep_scene=True),
    "gallery_area_200x150": dict(scene=("gen", "write_gallery", (3,)), w=200, h=150, area=1, point=0, stride=1,
                                 lights=[(-1, 1.2, 1.5), (1.5, 1.0, 1.0)], keep_scene=True),
    "gallery_area_d2_g4_200x150": dict(scene=("gen", "write_gallery", (3,)), w=200, h=150, area=1, point=0,
                                       stride=1, lights=[(-1, 1.2, 1.5)], max_depth=2, grid=(4, 4),
                                       keep_scene=False),
    "gallery_point_d0_320x240": dict(scene=("gen", "write_gallery", (3,)), w=320, h=240, area=0, point=1,
                                     stride=2, lights=[(-1, 1.2, 1.5)], max_depth=0, keep_scene=False),
    "hf32_point_256x144": dict(scene=("gen", "write_heightfield", (32,)), w=256, h=144, area=0, point=1, stride=1,
                               keep_scene=True),
    # BASELINE configs at full size, strided subsets (the reference needs minutes-hours per full frame)
    "hf224_point_3840x2160_s24": dict(scene=("gen", "write_heightfield", (224,)), w=3840, h=2160, area=0, point=1,
                                      stride=24, keep_scene=False),
    "hf707_point_1920x1080_s20": dict(scene=("gen", "write_heightfield", (707,)), w=1920, h=1080, area=0, point=1,
                                      stride=20, keep_scene=False),
    "hf707_point_7680x4320_s80": dict(scene=("gen", "write_heightfield", (707,)), w=7680, h=4320, area=0, point=1,
                                      stride=80, keep_scene=False),
    # BASELINE configs[3] as bench.py runs it, minus the analytic spheres (which the reference does not have):
    # 100 352 triangles, 3840x2160, 4x4 area light, depth cap 5 -- and the same with a mirror (illum 4) field so
    # that the depth-5 recursion is actually exercised on 100 k triangles
    "hf224_area_d5_g4_3840x2160_s24": dict(scene=("gen", "write_heightfield", (224,)), w=3840, h=2160, area=1, point=0,
                                           stride=24, max_depth=5, grid=(4, 4), keep_scene=False),
    "hf224m_area_d5_g4_3840x2160_s48": dict(scene=("gen", "write_heightfield", (224, 1234, 4)), w=3840, h=2160, area=1,
                                            point=0, stride=48, max_depth=5, grid=(4, 4), keep_scene=False),
    "cube_area_d3_g4_1920x1080_s9": dict(scene=("ref", "cube.obj"), w=1920, h=1080, area=1, point=0, stride=9,
                                         max_depth=3, grid=(4, 4), keep_scene=False),
}


def scene_path(spec, tmp):
    kind = spec[0]
    if kind == "ref":
        return os.path.join(O.REF_SCENES, spec[1])
    gen, args = spec[1], spec[2]
    p = os.path.join(tmp, f"{gen}_{'_'.join(map(str, args))}.obj")
    if not os.path.exists(p):
        getattr(scenes, gen)(p, *args)
    return p


def make(name, tmp):
    c = CASES[name]
    obj = scene_path(c["scene"], tmp)
    out = os.path.join(tmp, name + ".bin")
    dump = os.path.join(tmp, name + "_scene.bin")
    log = O.run_ref(obj, out, c["w"], c["h"], c["area"], c["point"], c["stride"], dump_scene=dump,
                    lights=c.get("lights"), cam_rot=c.get("cam_rot"), cam_trans=c.get("cam_trans"),
                    max_depth=c.get("max_depth"), grid=c.get("grid"))
    r = O.load_render_dump(out)
    s = O.load_scene_dump(dump)
    meta = dict(name=name, scene=list(c["scene"][:2]) + [list(c["scene"][2])] if c["scene"][0] == "gen" else list(c["scene"]),
                w=c["w"], h=c["h"], area=c["area"], point=c["point"], stride=c["stride"],
                max_depth=c.get("max_depth"), grid=c.get("grid"), n_faces=s.n_faces,
                ref_stdout=json.loads(log.strip().splitlines()[-1]), patched=bool(c.get("max_depth") is not None or c.get("grid")))
    arrays = dict(rgb=r.rgb, face=r.face, t=r.t, pxy=r.pxy,
                  # camera / lights / light model as the reference saw them
                  eye=s.eye, view_inv=s.view_inv, view=s.view, viewport=s.viewport,
                  cam=np.array([s.fovy, s.aspect, s.pscale], np.float32), lights=s.lights,
                  light_color=s.light_color, light_radius=np.float32(s.light_radius),
                  root_min=s.root_min, root_max=s.root_max, centroid=s.centroid,
                  radius=np.float32(s.radius), norm_scale=np.float32(s.norm_scale),
                  octree_stats=s.octree_stats, mats=s.mats, model_matrix=s.model_matrix,
                  meta=np.frombuffer(json.dumps(meta).encode(), np.uint8))
    if c.get("keep_scene"):
        arrays.update(verts=s.verts, fnormals=s.fnormals, vnormals=s.vnormals, mat_id=s.mat_id,
                      vertex_ids=s.vertex_ids, obj_verts=s.obj_verts)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
    print(name, "pixels", len(r.face), "ref render_s %.2f" % r.render_s, "faces", s.n_faces, flush=True)


def main():
    if not O.have_ref():
        O.build_ref()
    names = sys.argv[1:] or list(CASES)
    with tempfile.TemporaryDirectory(prefix="rt_golden_") as tmp:
        for n in names:
            make(n, tmp)


if __name__ == "__main__":
    main()
