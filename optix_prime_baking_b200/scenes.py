"""Procedural scenes for the BASELINE.json configs.

Host-side input generation only (numpy); nothing here is on the bake path.  All scenes are
functions of integer seeds, so the CPU oracle and the CUDA path see byte-identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class Mesh:
    vertices: np.ndarray                 # (nV, 3) float32
    tris: np.ndarray                     # (nT, 3) uint32
    normals: Optional[np.ndarray] = None  # (nV, 3) float32 or None

    def __post_init__(self):
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32)
        self.tris = np.ascontiguousarray(self.tris, dtype=np.uint32)
        if self.normals is not None:
            self.normals = np.ascontiguousarray(self.normals, dtype=np.float32)

    @property
    def bbox(self):
        if getattr(self, "_bbox", None) is None:
            self._bbox = (self.vertices.min(axis=0), self.vertices.max(axis=0)) if len(self.vertices) else \
                (np.zeros(3, np.float32), np.zeros(3, np.float32))
        return self._bbox


@dataclass
class Instance:
    mesh_index: int
    xform: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))  # row-major 4x4
    storage_identifier: int = 0


@dataclass
class Scene:
    meshes: List[Mesh]
    instances: List[Instance]

    def world_bbox(self):
        lo = np.full(3, np.inf, dtype=np.float64)
        hi = np.full(3, -np.inf, dtype=np.float64)
        for inst in self.instances:
            m = self.meshes[inst.mesh_index]
            blo, bhi = m.bbox
            corners = np.array([[x, y, z] for x in (blo[0], bhi[0]) for y in (blo[1], bhi[1])
                                for z in (blo[2], bhi[2])], dtype=np.float64)
            w = corners @ inst.xform[:3, :3].T.astype(np.float64) + inst.xform[:3, 3].astype(np.float64)
            lo = np.minimum(lo, w.min(axis=0))
            hi = np.maximum(hi, w.max(axis=0))
        return lo.astype(np.float32), hi.astype(np.float32)

    @property
    def num_triangles(self):
        return sum(len(self.meshes[i.mesh_index].tris) for i in self.instances)


def tea(rounds: int, v0, v1):
    """Vectorised tea<N> (random.h) on uint32 arrays."""
    v0 = np.array(v0, dtype=np.uint32, copy=True)
    v1 = np.array(v1, dtype=np.uint32, copy=True)
    v0, v1 = np.broadcast_arrays(v0, v1)
    v0 = v0.copy()
    v1 = v1.copy()
    s0 = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            s0 = np.uint32((int(s0) + 0x9E3779B9) & 0xFFFFFFFF)
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s0) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s0) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0


def _grid_tris(nu: int, nv: int) -> np.ndarray:
    """Two CCW (seen from +normal of a (u,v) right-handed patch) triangles per cell of an
    (nu+1) x (nv+1) vertex grid, row-major in u then v."""
    i, j = np.meshgrid(np.arange(nu, dtype=np.uint32), np.arange(nv, dtype=np.uint32), indexing="ij")
    a = (i * (nv + 1) + j).ravel()
    b = a + np.uint32(1)
    c = a + np.uint32(nv + 1)
    d = c + np.uint32(1)
    t = np.empty((2 * nu * nv, 3), dtype=np.uint32)
    t[0::2] = np.stack([a, b, d], axis=1)
    t[1::2] = np.stack([a, d, c], axis=1)
    return t


def vertex_normals(vertices: np.ndarray, tris: np.ndarray) -> np.ndarray:
    v = vertices.astype(np.float64)
    fn = np.cross(v[tris[:, 1]] - v[tris[:, 0]], v[tris[:, 2]] - v[tris[:, 0]])
    n = np.zeros_like(v)
    for k in range(3):
        for c in range(3):
            n[:, c] += np.bincount(tris[:, k], weights=fn[:, c], minlength=len(v))
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    return (n / ln).astype(np.float32)


def uv_sphere(stacks: int = 200, slices: int = 200, radius: float = 1.0, center=(0.0, 0.0, 0.0),
              displace: float = 0.0, seed: int = 0) -> Mesh:
    """UV sphere with single pole vertices: 2*slices*(stacks-1) triangles, outward CCW."""
    th = np.pi * np.arange(1, stacks, dtype=np.float64) / stacks          # polar, excl. poles
    ph = 2.0 * np.pi * np.arange(slices, dtype=np.float64) / slices
    st, sp = np.meshgrid(th, ph, indexing="ij")
    dirs = np.stack([np.sin(st) * np.cos(sp), np.cos(st), np.sin(st) * np.sin(sp)], axis=-1).reshape(-1, 3)
    dirs = np.concatenate([[[0.0, 1.0, 0.0]], dirs, [[0.0, -1.0, 0.0]]], axis=0)
    r = np.full(len(dirs), radius, dtype=np.float64)
    if displace != 0.0:
        h = tea(4, np.uint32(seed), np.arange(len(dirs), dtype=np.uint32)).astype(np.float64) / 2.0 ** 32
        # smooth-ish low-frequency bumps plus a little hash noise
        r = r * (1.0 + displace * (0.6 * np.sin(5 * dirs[:, 0]) * np.cos(4 * dirs[:, 1] + dirs[:, 2]) + 0.4 * (h - 0.5)))
    verts = dirs * r[:, None] + np.asarray(center, dtype=np.float64)
    tris = []
    ring = lambda k: 1 + k * slices  # first vertex of ring k (k = 0..stacks-2)
    j = np.arange(slices, dtype=np.int64)
    jn = (j + 1) % slices
    tris.append(np.stack([np.zeros_like(j), ring(0) + jn, ring(0) + j], axis=1))            # north cap
    for k in range(stacks - 2):
        a, b = ring(k) + j, ring(k) + jn
        c, d = ring(k + 1) + j, ring(k + 1) + jn
        quad = np.empty((2 * slices, 3), dtype=np.int64)
        quad[0::2] = np.stack([a, b, d], axis=1)
        quad[1::2] = np.stack([a, d, c], axis=1)
        tris.append(quad)
    south = len(dirs) - 1
    tris.append(np.stack([np.full_like(j, south), ring(stacks - 2) + j, ring(stacks - 2) + jn], axis=1))
    tris = np.concatenate(tris, axis=0).astype(np.uint32)
    verts32 = verts.astype(np.float32)
    normals = dirs.astype(np.float32) if displace == 0.0 else vertex_normals(verts32, tris)
    return Mesh(verts32, tris, normals)


def value_noise(x: np.ndarray, z: np.ndarray, seed: int, octaves: int = 4) -> np.ndarray:
    """Value noise from tea<4>(seed + octave, cell) lattice hashes, smoothstep-interpolated."""
    out = np.zeros_like(x, dtype=np.float64)
    amp, freq = 1.0, 1.0
    for o in range(octaves):
        fx, fz = x * freq, z * freq
        ix, iz = np.floor(fx).astype(np.int64), np.floor(fz).astype(np.int64)
        tx, tz = fx - ix, fz - iz
        tx = tx * tx * (3 - 2 * tx)
        tz = tz * tz * (3 - 2 * tz)

        def lat(i, j):
            cell = ((i & 0xFFFF) << 16 | (j & 0xFFFF)).astype(np.uint32)
            return tea(4, np.uint32(seed + o), cell).astype(np.float64) / 2.0 ** 32

        v = (lat(ix, iz) * (1 - tx) + lat(ix + 1, iz) * tx) * (1 - tz) + \
            (lat(ix, iz + 1) * (1 - tx) + lat(ix + 1, iz + 1) * tx) * tz
        out += amp * v
        amp *= 0.5
        freq *= 2.0
    return out


def heightfield(n: int = 708, seed: int = 1, size: float = 10.0, height: float = 1.5,
                base_freq: float = 6.0, warp: float = 0.0) -> Mesh:
    """(n x n)-cell terrain: 2*n*n triangles, up = +Y.  warp > 0 makes the grid spacing (and so
    the triangle areas) non-uniform (config 3, SURVEY §7 'area distribution quirk').
    Large grids (config 3: 27 s of numpy) are cached as .npy files in a scratch directory, so that the
    ranks of one job and back-to-back runs on one box generate them once."""
    if n >= 2000:
        import os
        import tempfile
        d = os.environ.get("AOBAKE_SCENE_CACHE", os.path.join(tempfile.gettempdir(), "aobake_scene_cache"))
        stem = os.path.join(d, f"hf_v1_{n}_{seed}_{size}_{height}_{base_freq}_{warp}")
        try:
            return Mesh(np.load(stem + "_v.npy"), np.load(stem + "_t.npy"), np.load(stem + "_n.npy"))
        except (OSError, ValueError):
            pass
        m = _heightfield(n, seed, size, height, base_freq, warp)
        try:
            os.makedirs(d, exist_ok=True)
            for suffix, arr in (("_v.npy", m.vertices), ("_t.npy", m.tris), ("_n.npy", m.normals)):
                tmp = f"{stem}{suffix}.{os.getpid()}.tmp"
                with open(tmp, "wb") as f:
                    np.save(f, arr)
                os.replace(tmp, stem + suffix)   # atomic: a concurrent reader sees the old state or the whole file
        except OSError:
            pass
        return m
    return _heightfield(n, seed, size, height, base_freq, warp)


def _heightfield(n: int, seed: int, size: float, height: float, base_freq: float, warp: float) -> Mesh:
    u = np.arange(n + 1, dtype=np.float64) / n
    if warp > 0.0:
        u = u + warp * np.sin(2 * np.pi * u) / (2 * np.pi)
    # u-major rows (x), v-minor (z); winding chosen so that face normals point +Y
    xx, zz = np.meshgrid(u, u, indexing="ij")
    yy = height * (value_noise(xx * base_freq, zz * base_freq, seed) / 1.875 - 0.5)
    verts = np.stack([xx * size, yy, zz * size], axis=-1).reshape(-1, 3).astype(np.float32)
    tris = _grid_tris(n, n)
    return Mesh(verts, tris, vertex_normals(verts, tris))


def ground_plane(bbox_min, bbox_max, upaxis: int = 1, scale_factor: float = 100.0,
                 offset_factor: float = 0.03) -> Mesh:
    """make_ground_plane (main.cpp): a 2-triangle blocker quad under the scene."""
    bbox_min = np.asarray(bbox_min, dtype=np.float32)
    bbox_max = np.asarray(bbox_max, dtype=np.float32)
    axis, flip = upaxis % 3, upaxis >= 3
    a1, a2 = (axis + 1) % 3, (axis + 2) % 3
    ext = np.float32((bbox_max - bbox_min).max())
    h = bbox_max[axis] + np.float32(offset_factor) * ext if flip else bbox_min[axis] - np.float32(offset_factor) * ext
    c1 = np.float32(0.5) * (bbox_min[a1] + bbox_max[a1])
    c2 = np.float32(0.5) * (bbox_min[a2] + bbox_max[a2])
    h1 = np.float32(0.5) * np.float32(scale_factor) * (bbox_max[a1] - bbox_min[a1])
    h2 = np.float32(0.5) * np.float32(scale_factor) * (bbox_max[a2] - bbox_min[a2])
    v = np.zeros((4, 3), dtype=np.float32)
    for i, (s1, s2) in enumerate([(-1, -1), (1, -1), (1, 1), (-1, 1)]):
        v[i, axis] = h
        v[i, a1] = c1 + np.float32(s1) * h1
        v[i, a2] = c2 + np.float32(s2) * h2
    t = np.array([[0, 2, 1], [0, 3, 2]] if flip else [[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    return Mesh(v, t, None)


def ground_blockers(scene: Scene, upaxis: int = 1, scale_factor: float = 100.0,
                    offset_factor: float = 0.03) -> Scene:
    lo, hi = scene.world_bbox()
    return Scene([ground_plane(lo, hi, upaxis, scale_factor, offset_factor)], [Instance(0)])


def _rotation(seed: int, i: int) -> np.ndarray:
    h = [int(tea(4, np.uint32(seed + k), np.uint32(i))) / 2.0 ** 32 for k in range(3)]
    # uniform random rotation from three uniforms (Shoemake)
    u1, u2, u3 = h
    q = np.array([np.sqrt(1 - u1) * np.sin(2 * np.pi * u2), np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
                  np.sqrt(u1) * np.sin(2 * np.pi * u3), np.sqrt(u1) * np.cos(2 * np.pi * u3)])
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


# ---- the five BASELINE.json configs (scaled by `scale` in tests) ----------------------
def config1_sphere(stacks: int = 200, slices: int = 200):
    """c1: unit sphere resting on y = -1 + default ground blocker."""
    scene = Scene([uv_sphere(stacks, slices)], [Instance(0)])
    return scene, ground_blockers(scene)


def config2_heightfield(n: int = 708, seed: int = 1):
    scene = Scene([heightfield(n, seed)], [Instance(0)])
    return scene, Scene([], [])


def config3_bigmesh(n: int = 3163, seed: int = 3, with_ground: bool = False):
    scene = Scene([heightfield(n, seed, size=40.0, height=3.0, base_freq=24.0, warp=0.6)], [Instance(0)])
    return scene, (ground_blockers(scene) if with_ground else Scene([], []))


def config4_instanced(grid: int = 10, stacks: int = 158, slices: int = 158, seed: int = 4,
                      with_ground: bool = False):
    mesh = uv_sphere(stacks, slices, displace=0.15, seed=seed)
    lo, hi = mesh.bbox
    spacing = 2.5 * float((hi - lo).max())
    insts = []
    for i in range(grid ** 3):
        ix, iy, iz = i % grid, (i // grid) % grid, i // (grid * grid)
        m = np.eye(4, dtype=np.float64)
        m[:3, :3] = _rotation(seed * 1000, i)
        m[:3, 3] = [ix * spacing, iy * spacing, iz * spacing]
        insts.append(Instance(0, m.astype(np.float32), storage_identifier=i))
    scene = Scene([mesh], insts)
    return scene, (ground_blockers(scene) if with_ground else Scene([], []))


def default_distances(scene: Scene, offset_scale: float = 0.01, maxdist_scale: float = 10.0):
    """-d / -m defaults of the sample CLI: fractions of the max scene extent."""
    lo, hi = scene.world_bbox()
    ext = float((hi - lo).max())
    return np.float32(offset_scale * ext), np.float32(maxdist_scale * ext)
