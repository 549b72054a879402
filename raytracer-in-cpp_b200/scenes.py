"""Synthetic scene generators (OBJ + MTL text in the dialect the reference's loader accepts).

The BASELINE.json configs name synthetic scenes by shape only ("1M-triangle mesh", "100k-triangle
+ 1k-sphere scene"); SURVEY.md section 8(d) fixes the generators (height-field grid, numpy
default_rng seeds) so that the reference, the oracle and the CUDA path all load the identical file.

File dialect constraints (reference loader: tucano/utils/objimporter.hpp:83-284, mtlIO.hpp:45-125):
  * faces must be triangles, "f a b c" (1-based vertex ids);
  * MTL tokens are split on single spaces; material names must not contain spaces;
  * "usemtl <name>" must match the MTL name byte for byte (no trailing blanks / CR).
"""
from __future__ import annotations

import io
import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Material:
    name: str
    kd: tuple = (0.5, 0.5, 0.5)
    ks: tuple = (1.0, 1.0, 1.0)
    ns: float = 10.0
    ni: float = 0.0
    illum: int = 2


def write_obj(path: str, verts: np.ndarray, groups: list, materials: list) -> None:
    """groups: list of (material_name, faces[int, 3] zero-based)."""
    base = os.path.splitext(os.path.basename(path))[0]
    mtl_path = os.path.splitext(path)[0] + ".mtl"
    with open(mtl_path, "w") as f:
        for m in materials:
            f.write(f"newmtl {m.name}\n")
            f.write("Ka 0.0 0.0 0.0\n")
            f.write("Kd %.9g %.9g %.9g\n" % tuple(m.kd))
            f.write("Ks %.9g %.9g %.9g\n" % tuple(m.ks))
            f.write("Ns %.9g\n" % m.ns)
            f.write("Ni %.9g\n" % m.ni)
            f.write("illum %d\n" % m.illum)
            f.write("\n")
    buf = io.StringIO()
    buf.write(f"mtllib {base}.mtl\n")
    np.savetxt(buf, np.asarray(verts, np.float32), fmt="v %.9g %.9g %.9g")
    for name, faces in groups:
        buf.write(f"usemtl {name}\n")
        np.savetxt(buf, np.asarray(faces, np.int64) + 1, fmt="f %d %d %d")
    with open(path, "w") as f:
        f.write(buf.getvalue())


# ----------------------------------------------------------------------------------------------
# primitive builders
# ----------------------------------------------------------------------------------------------
def heightfield_mesh(n: int, seed: int = 1234, noise: float = 0.02):
    """(n+1)^2 vertex grid over [-1,1]^2, z = 0.15 sin(5x) cos(4y) + noise*N(0,1); 2 n^2 triangles
    wound so that normals face +z (towards the default camera at (0,0,2))."""
    rng = np.random.default_rng(seed)
    lin = np.linspace(-1.0, 1.0, n + 1)
    x, y = np.meshgrid(lin, lin, indexing="xy")
    z = 0.15 * np.sin(5 * x) * np.cos(4 * y) + noise * rng.standard_normal(x.shape)
    verts = np.stack([x, y, z], -1).reshape(-1, 3).astype(np.float32)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a = idx[:-1, :-1].ravel(); b = idx[:-1, 1:].ravel(); c = idx[1:, 1:].ravel(); d = idx[1:, :-1].ravel()
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)
    # interleave the two triangles of each cell so that neighbouring faces have neighbouring ids
    faces = faces.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    return verts, faces


def icosphere(subdiv: int, radius: float = 1.0, center=(0.0, 0.0, 0.0)):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
         (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
         (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = v[i] + v[j]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for (a, b, c) in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    verts = (np.array(v) * radius + np.array(center)).astype(np.float32)
    return verts, np.array(f, np.int64)


def quad(p0, p1, p2, p3):
    """Two triangles p0 p1 p2 / p0 p2 p3."""
    return np.array([p0, p1, p2, p3], np.float32), np.array([(0, 1, 2), (0, 2, 3)], np.int64)


class SceneBuilder:
    def __init__(self):
        self.verts = []
        self.nv = 0
        self.groups = []
        self.materials = []

    def material(self, m: Material):
        self.materials.append(m)
        return m.name

    def add(self, mat_name, verts, faces):
        self.verts.append(np.asarray(verts, np.float32))
        self.groups.append((mat_name, np.asarray(faces, np.int64) + self.nv))
        self.nv += len(verts)

    def write(self, path):
        write_obj(path, np.concatenate(self.verts, 0), self.groups, self.materials)
        return path

    @property
    def n_faces(self):
        return int(sum(len(f) for _, f in self.groups))


# ----------------------------------------------------------------------------------------------
# named scenes
# ----------------------------------------------------------------------------------------------
def write_heightfield(path: str, n: int, seed: int = 1234, illum: int = 2) -> int:
    """SURVEY.md 8(d) C3/C5 scene: single material Kd .7 Ks .5 Ns 50.  n=707 -> 999 698 triangles,
    n=224 -> 100 352 triangles."""
    verts, faces = heightfield_mesh(n, seed)
    mat = Material("field", kd=(0.7, 0.7, 0.7), ks=(0.5, 0.5, 0.5), ns=50.0, illum=illum)
    write_obj(path, verts, [("field", faces)], [mat])
    return len(faces)


def write_gallery(path: str, sphere_subdiv: int = 2) -> int:
    """A small room exercising every branch of the reference's material switch
    (src/flyscene.cpp:712-760): diffuse floor/wall (illum 2), mirror (3, 4), Fresnel mirror (5),
    refractive (6, 7) and pass-through glass (9) objects.  sphere_subdiv=3 gives > 1000 faces so the
    reference octree actually splits."""
    b = SceneBuilder()
    b.material(Material("floor", kd=(0.6, 0.6, 0.55), ks=(0.2, 0.2, 0.2), ns=20, illum=2))
    b.material(Material("wall", kd=(0.3, 0.5, 0.7), ks=(0.1, 0.1, 0.1), ns=5, illum=1))
    b.material(Material("mirror3", kd=(0.2, 0.2, 0.2), ks=(1, 1, 1), ns=80, illum=3))
    b.material(Material("mirror4", kd=(0.7, 0.1, 0.1), ks=(1, 1, 1), ns=30, illum=4))
    b.material(Material("fresnel5", kd=(0.1, 0.6, 0.2), ks=(0.8, 0.8, 0.8), ns=40, ni=1.5, illum=5))
    b.material(Material("refract6", kd=(0.5, 0.5, 0.1), ks=(0.6, 0.6, 0.6), ns=25, ni=1.3, illum=6))
    b.material(Material("refract7", kd=(0.4, 0.1, 0.5), ks=(0.9, 0.9, 0.9), ns=60, ni=1.2, illum=7))
    b.material(Material("glass9", kd=(0.9, 0.9, 1.0), ks=(1, 1, 1), ns=100, illum=9))
    # floor (y = -1) and back wall (z = -1.5), both facing the camera side
    b.add("floor", *quad((-2, -1, 1.2), (2, -1, 1.2), (2, -1, -1.5), (-2, -1, -1.5)))
    b.add("wall", *quad((-2, -1, -1.5), (2, -1, -1.5), (2, 1.5, -1.5), (-2, 1.5, -1.5)))
    # a tilted mirror panel on the left
    b.add("mirror3", *quad((-1.9, -0.9, -1.2), (-1.0, -0.9, -0.2), (-1.0, 1.0, -0.2), (-1.9, 1.0, -1.2)))
    # a glass pane in front of the centre
    b.add("glass9", *quad((-0.5, -0.6, 0.6), (0.5, -0.6, 0.6), (0.5, 0.4, 0.6), (-0.5, 0.4, 0.6)))
    for name, cx, cy, cz, r in [("mirror4", -0.3, -0.55, -0.4, 0.45), ("fresnel5", 0.9, -0.6, -0.2, 0.4),
                                ("refract6", 0.15, -0.75, 0.25, 0.22), ("refract7", 1.3, 0.3, -0.9, 0.35)]:
        v, f = icosphere(sphere_subdiv, r, (cx, cy, cz))
        b.add(name, v, f)
    b.write(path)
    return b.n_faces


def sphere_cloud(n: int = 1000, seed: int = 4321):
    """centres uniform in the unit ball, r in U[0.01, 0.03], materials alternate."""
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((n, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    p *= rng.random((n, 1)) ** (1.0 / 3.0)
    r = rng.uniform(0.01, 0.03, (n, 1))
    return np.concatenate([p, r], 1).astype(np.float32)
