"""The CLI (tools/aobake_cli.cpp; SURVEY §8f next rows: OBJ loader, -i grid instancing, result
writer, reference flag set): builds with plain g++ on CPU; on the GPU it bakes an OBJ written
here and its raw dump must equal the Python API's result for the same scene."""
import os
import struct
import subprocess

import numpy as np
import pytest

from optix_prime_baking_b200 import scenes
from optix_prime_baking_b200.scenes import Instance, Scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "optix_prime_baking_b200")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build_cli(tmp_path):
    from optix_prime_baking_b200 import build
    build.build()
    exe = str(tmp_path / "aobake_cli")
    res = subprocess.run([GXX, "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "aobake_cli.cpp"),
                          "-o", exe, "-L", LIBDIR, "-laobake", f"-Wl,-rpath,{LIBDIR}"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def write_obj(path, mesh):
    with open(path, "w") as f:
        for v in mesh.vertices:
            f.write(f"v {v[0]:.9g} {v[1]:.9g} {v[2]:.9g}\n")
        for n in mesh.normals:
            f.write(f"vn {n[0]:.9g} {n[1]:.9g} {n[2]:.9g}\n")
        for t in mesh.tris:
            a, b, c = (int(x) + 1 for x in t)
            f.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")


def read_raw(path):
    with open(path, "rb") as f:
        ni, nvt = struct.unpack("<QQ", f.read(16))
        recs = [struct.unpack("<QQQ", f.read(24)) for _ in range(ni)]
        data = np.frombuffer(f.read(4 * nvt), dtype=np.float32)
    return recs, data


def test_cli_builds_and_rejects_bad_flags(tmp_path):
    exe = build_cli(tmp_path)
    assert subprocess.run([exe, "--cpu"], capture_output=True).returncode == 2      # no CPU path, loudly
    assert subprocess.run([exe, "--bogus"], capture_output=True).returncode == 2
    assert subprocess.run([exe, "-f", "/nonexistent.obj"], capture_output=True).returncode == 1


def test_cli_rejects_malformed_obj_files(tmp_path):
    """face indices were used before they were validated.  A zero / out-of-range vertex index,
    or a normal index past the vn count, must be `could not load` (exit 1), not a heap overrun."""
    exe = build_cli(tmp_path)
    head = "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\n"
    for k, face in enumerate(["f 0//1 1//1 2//1", "f 1//1 2//1 9//1", "f 1//1 2//1 3//7", "f 1//1 2//1 -9//1", "f 1//-5 2//1 3//1"]):
        path = tmp_path / f"bad{k}.obj"
        path.write_text(head + face + "\n")
        res = subprocess.run([exe, "-f", str(path), "--no_viewer"], capture_output=True, text=True)
        assert res.returncode == 1, (face, res.returncode, res.stderr)


@pytest.mark.gpu
def test_cli_bake_matches_python_api(tmp_path):
    from optix_prime_baking_b200 import api
    exe = build_cli(tmp_path)
    mesh = scenes.uv_sphere(24, 24)
    obj, raw = str(tmp_path / "s.obj"), str(tmp_path / "s.raw")
    write_obj(obj, mesh)
    res = subprocess.run([exe, "-f", obj, "-o", raw, "-i", "3", "-r", "64", "--no_least_squares"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "Total samples" in res.stderr and "Compute AO" in res.stderr
    recs, data = read_raw(raw)
    assert [r[0] for r in recs] == [0, 1, 2] and all(r[2] == len(mesh.vertices) for r in recs)
    # the same scene through the Python API
    lo, hi = mesh.bbox
    insts = []
    for i in range(3):
        xf = np.eye(4, dtype=np.float32)
        g = (i % 2, (i // 2) % 2, i // 4)
        for k in range(3):
            xf[k, 3] = np.float32(1.1) * (hi[k] - lo[k]) * np.float32(g[k])
        insts.append(Instance(0, xf, i))
    scene = Scene([mesh], insts)
    blockers = scenes.ground_blockers(scene)
    off, maxd = scenes.default_distances(scene)
    with api.Baker() as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(3, 0)
        bk.sample_instances(per, 3, download=False)
        bk.compute_ao(64, off, maxd, download=False)
        want = np.concatenate(bk.map_ao_to_vertices(api.FILTER_AREA_BASED))
    # the CLI derives the world box (hence offset/maxdist/ground height) in fp32, the Python helper in
    # fp64: a ray in a million may flip, moving one vertex by ~1e-3
    d = np.abs(data - want)
    assert d.max() < 5e-3 and d.mean() < 1e-5


@pytest.mark.gpu
def test_cli_multi_gpu_matches_single_gpu(tmp_path):
    """--gpus 2: C++ threads + the native NCCL exchange; the AO is bit-identical, so the vertex dumps agree to fp64-atomics rounding."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = build_cli(tmp_path)
    outs = []
    for gpus in (1, 2):
        raw = str(tmp_path / f"g{gpus}.raw")
        res = subprocess.run([exe, "-o", raw, "-i", "2", "-r", "64", "-s", "200000", "--no_least_squares", "--gpus", str(gpus)],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr
        outs.append(read_raw(raw)[1])
    assert np.abs(outs[0] - outs[1]).max() < 1e-6
