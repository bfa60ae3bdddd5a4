"""ctypes binding of oracle/libao_oracle.so — TEST INFRASTRUCTURE (the checker), never the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from optix_prime_baking_b200.ctypes_types import PackedScene, SampleBuffers
from optix_prime_baking_b200.scenes import Scene

_ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
_LIB = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", _ORACLE_DIR], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ORACLE_DIR, "libao_oracle.so")
        src = os.path.join(_ORACLE_DIR, "ao_oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build_oracle()
        L = C.CDLL(path)
        L.ao_oracle_tea.restype = C.c_uint32
        L.ao_oracle_tea.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.ao_oracle_lcg.restype = C.c_uint32
        L.ao_oracle_rnd.restype = C.c_float
        L.ao_oracle_halton.restype = C.c_float
        L.ao_oracle_halton.argtypes = [C.c_uint32, C.c_uint32]
        L.ao_oracle_sincos2pi.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ao_oracle_distribute_samples.restype = C.c_uint64
        L.ao_oracle_distribute_samples.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ao_oracle_triangle_counts.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ao_oracle_sample_instances.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ao_oracle_tracer_create.restype = C.c_void_p
        L.ao_oracle_tracer_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ao_oracle_tracer_destroy.argtypes = [C.c_void_p]
        L.ao_oracle_tracer_is_two_level.argtypes = [C.c_void_p]
        L.ao_oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.ao_oracle_ray_margin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ao_oracle_generate_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.ao_oracle_generate_rays_for.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.ao_oracle_compute_ao.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_float,
                                           C.c_float, C.c_void_p, C.c_void_p]
        L.ao_oracle_filter_area.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ao_oracle_filter_least_squares.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                                     C.c_double, C.c_int, C.c_void_p, C.c_int]
        L.ao_oracle_ls_residual.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int]
        L.ao_oracle_ls_residual.restype = C.c_double
        L.ao_oracle_instance_areas.argtypes = [C.c_void_p, C.c_void_p]
        L.ao_oracle_make_ground_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                                  C.c_void_p, C.c_void_p]
        L.ao_oracle_affine_inverse.argtypes = [C.c_void_p, C.c_void_p]
        L.ao_oracle_set_traversal.argtypes = [C.c_int]
        L.ao_oracle_set_traversal.restype = C.c_int
        _LIB = L
    return _LIB


TRAVERSAL_AUTO, TRAVERSAL_BINARY, TRAVERSAL_WIDE = 0, 1, 2


def set_traversal(mode: int = TRAVERSAL_AUTO) -> int:
    """Traversal of the oracle's triangle BVHs (process-wide): 0 = auto, 1 = binary scalar, 2 = 8-wide AVX2.  Returns the
    kind in effect on this CPU (1 or 2); both give the same answer for every ray."""
    return int(lib().ao_oracle_set_traversal(int(mode)))


def tea(rounds, v0, v1):
    return lib().ao_oracle_tea(rounds, v0 & 0xFFFFFFFF, v1 & 0xFFFFFFFF)


def lcg_stream(seed, n):
    s = C.c_uint32(seed)
    return [lib().ao_oracle_lcg(C.byref(s)) for _ in range(n)]


def halton(i, base):
    return lib().ao_oracle_halton(i, base)


def sincos2pi(u):
    c, s = C.c_float(), C.c_float()
    lib().ao_oracle_sincos2pi(C.c_float(u), C.byref(c), C.byref(s))
    return c.value, s.value


def _vertex_out(scene: Scene):
    arrs = [np.zeros(len(scene.meshes[i.mesh_index].vertices), dtype=np.float32) for i in scene.instances]
    ptrs = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
    return arrs, ptrs


class Oracle:
    """Object wrapper: one scene (+ blockers), mirrors the bake API call order."""

    def __init__(self, scene: Scene, blockers: Scene | None = None, mode: int = 0):
        self.scene, self.blockers = scene, blockers
        self.ps = PackedScene(scene)
        self.pb = PackedScene(blockers) if blockers is not None and len(blockers.instances) else None
        self.mode = mode
        self._tracer = None
        self.per_instance = None
        self.samples = None

    def instance_areas(self):
        out = np.zeros(len(self.scene.instances), dtype=np.float64)
        lib().ao_oracle_instance_areas(self.ps.ref(), out.ctypes.data)
        return out

    def distribute_samples(self, min_per_tri: int, requested: int):
        per = np.zeros(max(len(self.scene.instances), 1), dtype=np.uint64)
        total = lib().ao_oracle_distribute_samples(self.ps.ref(), min_per_tri, requested, per.ctypes.data)
        self.per_instance = per[:len(self.scene.instances)]
        return int(total), self.per_instance

    def triangle_counts(self, inst, n_samples, min_per_tri):
        nT = len(self.scene.meshes[self.scene.instances[inst].mesh_index].tris)
        counts = np.zeros(nT, dtype=np.uint64)
        rc = lib().ao_oracle_triangle_counts(self.ps.ref(), inst, n_samples, min_per_tri, counts.ctypes.data)
        assert rc == 0
        return counts

    def sample_instances(self, per_instance, min_per_tri: int) -> SampleBuffers:
        per = np.ascontiguousarray(per_instance, dtype=np.uint64)
        self.per_instance = per
        sb = SampleBuffers(int(per.sum()))
        rc = lib().ao_oracle_sample_instances(self.ps.ref(), per.ctypes.data, min_per_tri, sb.ref())
        assert rc == 0, rc
        self.samples = sb
        return sb

    @property
    def tracer(self):
        if self._tracer is None:
            self._tracer = lib().ao_oracle_tracer_create(self.ps.ref(), self.pb.ref() if self.pb else None, self.mode)
        return self._tracer

    def is_two_level(self):
        return bool(lib().ao_oracle_tracer_is_two_level(self.tracer))

    def trace_rays(self, rays: np.ndarray, brute: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hit = np.zeros(len(rays), dtype=np.uint8)
        lib().ao_oracle_trace_rays(self.tracer, rays.ctypes.data, len(rays), hit.ctypes.data, 1 if brute else 0)
        return hit

    def ray_margin(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = np.zeros(len(rays), dtype=np.float32)
        lib().ao_oracle_ray_margin(self.tracer, rays.ctypes.data, len(rays), out.ctypes.data)
        return out

    def generate_rays(self, samples: SampleBuffers, begin, end, rays_per_sample, offset, maxdist):
        q = lib().ao_oracle_sqrt_rays(rays_per_sample)
        out = np.zeros((end - begin, q * q, 8), dtype=np.float32)
        lib().ao_oracle_generate_rays(samples.ref(), begin, end, rays_per_sample, float(offset), float(maxdist),
                                      out.ctypes.data)
        return out

    def generate_rays_for(self, samples: SampleBuffers, k, global_index, rays_per_sample, offset, maxdist):
        """Rays of local sample k with the RNG streams of `global_index` (subset parity checks)."""
        q = lib().ao_oracle_sqrt_rays(rays_per_sample)
        out = np.zeros((q * q, 8), dtype=np.float32)
        lib().ao_oracle_generate_rays_for(samples.ref(), k, global_index, rays_per_sample, float(offset), float(maxdist),
                                          out.ctypes.data)
        return out

    def compute_ao(self, samples: SampleBuffers, rays_per_sample, offset, maxdist, begin=0, end=None):
        end = samples.n if end is None else end
        ao = np.zeros(end - begin, dtype=np.float32)
        hits = np.zeros(end - begin, dtype=np.uint32)
        lib().ao_oracle_compute_ao(self.tracer, samples.ref(), begin, end, rays_per_sample, float(offset),
                                   float(maxdist), ao.ctypes.data, hits.ctypes.data)
        return ao, hits

    def filter_area(self, samples: SampleBuffers, ao: np.ndarray, per_instance=None):
        per = np.ascontiguousarray(self.per_instance if per_instance is None else per_instance, dtype=np.uint64)
        arrs, ptrs = _vertex_out(self.scene)
        ao = np.ascontiguousarray(ao, dtype=np.float32)
        lib().ao_oracle_filter_area(self.ps.ref(), per.ctypes.data, samples.ref(), ao.ctypes.data, ptrs)
        return arrs

    def filter_least_squares(self, samples: SampleBuffers, ao: np.ndarray, weight=0.1, tol=1e-10,
                             max_iter=20000, per_instance=None, energy=0):
        """energy 0: (A1+A2) |grad jump|^2 (the default); 1: the scale-free round-1 form."""
        per = np.ascontiguousarray(self.per_instance if per_instance is None else per_instance, dtype=np.uint64)
        arrs, ptrs = _vertex_out(self.scene)
        ao = np.ascontiguousarray(ao, dtype=np.float32)
        iters = lib().ao_oracle_filter_least_squares(self.ps.ref(), per.ctypes.data, samples.ref(), ao.ctypes.data,
                                                     float(weight), float(tol), int(max_iter), ptrs, int(energy))
        assert iters >= 0
        self.ls_iterations = iters
        return arrs

    def ls_residual(self, samples: SampleBuffers, ao: np.ndarray, vertex_x, weight=0.1, per_instance=None, energy=0) -> float:
        """|b - (M + wR) x| / |b| of a candidate solution (one array per instance) under the oracle's operator."""
        per = np.ascontiguousarray(self.per_instance if per_instance is None else per_instance, dtype=np.uint64)
        xs = [np.ascontiguousarray(v, dtype=np.float32) for v in vertex_x]
        ptrs = (C.c_void_p * max(len(xs), 1))(*[a.ctypes.data for a in xs])
        ao = np.ascontiguousarray(ao, dtype=np.float32)
        return float(lib().ao_oracle_ls_residual(self.ps.ref(), per.ctypes.data, samples.ref(), ao.ctypes.data, float(weight), ptrs, int(energy)))

    def close(self):
        if self._tracer is not None:
            lib().ao_oracle_tracer_destroy(self._tracer)
            self._tracer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
