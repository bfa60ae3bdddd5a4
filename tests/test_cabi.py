"""The C-ABI boundary: libaobake.so loads and exports every symbol include/aobake.h declares;
the C++ bake:: shim compiles and links with plain g++ (CPU), and runs on the GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "optix_prime_baking_b200", "libaobake.so")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aobake.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aobake_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from optix_prime_baking_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in aobake.h but not exported"
    from optix_prime_baking_b200 import api
    assert sorted(api.EXPORTS) == names


def test_pure_host_entry_points(lib_path):
    """Entry points that need no device: defaults, ground plane, error text, create without a GPU."""
    from optix_prime_baking_b200 import api
    p = api.default_params()
    assert p.device == 0 and p.cg_max_iterations > 0 and p.cg_tolerance > 0
    v, t = api.make_ground_plane([-1, -1, -1], [1, 1, 1], 1, 100.0, 0.03)
    assert v.shape == (4, 3) and abs(v[0, 1] - (-1 - 0.06)) < 1e-6 and abs(abs(v[:, 0]).max() - 100.0) < 1e-4
    with pytest.raises(api.AoBakeError):
        api.make_ground_plane([0, 0, 0], [1, 1, 1], 9)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(api.AoBakeError) as ei:       # loud failure, never a CPU fallback
            api.Baker()
        assert "no CUDA device" in str(ei.value) or ei.value.code in (2, 5)


def _compile_cpp(tmp_path):
    exe = str(tmp_path / "test_bake_api")
    cmd = [GXX, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_bake_api.cpp"),
           "-o", exe, "-L", os.path.dirname(LIB), "-laobake", f"-Wl,-rpath,{os.path.dirname(LIB)}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_cpp_shim_compiles_and_links(lib_path, tmp_path):
    _compile_cpp(tmp_path)


@pytest.mark.gpu
def test_cpp_shim_runs(lib_path, tmp_path):
    exe = _compile_cpp(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "bake_api ok" in res.stdout


def _compile(tmp_path, src, name):
    exe = str(tmp_path / name)
    cmd = [GXX, "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", src),
           "-o", exe, "-L", os.path.dirname(LIB), "-laobake", f"-Wl,-rpath,{os.path.dirname(LIB)}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_cpp_distributed_compiles(lib_path, tmp_path):
    _compile(tmp_path, "test_distributed.cpp", "test_distributed")


@pytest.mark.gpu
def test_cpp_distributed_runs_without_python(lib_path, tmp_path):
    """Multi-GPU bake driven from C++ threads only: sharding + NCCL all-reduce inside libaobake.so."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = _compile(tmp_path, "test_distributed.cpp", "test_distributed")
    res = subprocess.run([exe, str(min(n, 4))], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "distributed ok" in res.stdout


def test_ctypes_mirrors_match_the_c_compiler(tmp_path):
    """sizeof/offsetof of every struct in aobake.h as gcc sees them == the ctypes mirrors."""
    import ctypes as C
    from optix_prime_baking_b200 import api
    from optix_prime_baking_b200 import ctypes_types as ct
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "aobake.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(AoMesh),offsetof(AoMesh,num_triangles),offsetof(AoMesh,bbox_min),sizeof(AoInstance),offsetof(AoInstance,storage_identifier),'
                   'offsetof(AoInstance,mesh_index),sizeof(AoScene),sizeof(AoSampleInfo),sizeof(AoSamples),sizeof(AoBakeParams),offsetof(AoBakeParams,refill_below),'
                   'sizeof(AoTimings),offsetof(AoTimings,rays_traced),sizeof(AoStats));return 0;}\n')
    exe = str(tmp_path / "layout")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    assert subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], capture_output=True).returncode == 0
    got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(ct.AoMesh), ct.AoMesh.num_triangles.offset, ct.AoMesh.bbox_min.offset, C.sizeof(ct.AoInstance),
            ct.AoInstance.storage_identifier.offset, ct.AoInstance.mesh_index.offset, C.sizeof(ct.AoScene), C.sizeof(ct.AoSampleInfo),
            C.sizeof(ct.AoSamples), C.sizeof(api.AoBakeParams), api.AoBakeParams.refill_below.offset, C.sizeof(api.AoTimings),
            api.AoTimings.rays_traced.offset, C.sizeof(api.AoStats)]
    assert got == want


def test_missing_nccl_is_a_status_not_a_crash(lib_path):
    """NcclApi::load() once built its message from two dlerror() calls and dereferenced NULL.
    With the NCCL library unresolvable, aobake_comm_unique_id must return AOBAKE_ERR_COMM and a message
    (needs no GPU: binding NCCL is a dlopen)."""
    code = ("import ctypes, sys\n"
            f"L = ctypes.CDLL({lib_path!r})\n"
            "L.aobake_last_error.restype = ctypes.c_char_p\n"
            "L.aobake_last_error.argtypes = [ctypes.c_void_p]\n"
            "buf = ctypes.create_string_buffer(128)\n"
            "rc = L.aobake_comm_unique_id(buf)\n"
            "msg = L.aobake_last_error(None).decode()\n"
            "print(rc, msg)\n"
            "sys.exit(0 if rc == 7 and 'dlopen' in msg else 1)\n")
    env = dict(os.environ, AOBAKE_NCCL_LIB="/nonexistent/libnccl-missing.so")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
