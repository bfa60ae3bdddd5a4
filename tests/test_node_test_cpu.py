"""Property tests of the two box tests of the traversal kernels (intersect_node8 in fp32 and
intersect_node8_h2 in packed fp16), run through the CPU emulation of the same source (tests/emu):

  conservative — every child box that the exact ray (evaluated in fp64 on the de-quantised box)
  touches inside [tmin, tmax] must be reported, for nodes of any size at any distance, thin nodes,
  steep and grazing rays, rays that start inside the box, short rays.  A bit missing here is a
  triangle the GPU would never test, i.e. a silent false miss — the traversal-level fuzz tests
  only see those that change a final answer.

The fp16 emulation itself (round-to-nearest-even from a double, subnormals, saturation) is pinned
against numpy's float16."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    E = C.CDLL(os.path.join(d, "libaob_emu.so"))
    E.emu_node_test.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    for f in ("emu_h2_fma", "emu_h2_min", "emu_h2_lane_sum"):
        getattr(E, f).restype = C.c_uint32
    E.emu_h2_pack_sat.restype = C.c_uint32
    E.emu_h2_pack_sat.argtypes = [C.c_float, C.c_float]
    E.emu_h2_fma.argtypes = [C.c_uint32] * 3
    E.emu_h2_min.argtypes = [C.c_uint32] * 2
    E.emu_h2_lane_sum.argtypes = [C.c_uint32]
    return E


def _h(bits):
    return np.array([bits], dtype=np.uint16).view(np.float16)[0]


def _bits(x):
    return int(np.array([x], dtype=np.float16).view(np.uint16)[0])


def test_fp16_emulation_matches_numpy_float16(emu):
    rng = np.random.default_rng(0)
    # conversions: normals, subnormals, ties, overflow (saturating), signed zero
    vals = np.concatenate([rng.normal(size=2000) * 10.0 ** rng.uniform(-9, 5, 2000), [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9,
                           2.0 ** -24, 2.0 ** -25, 3 * 2.0 ** -25, 2.0 ** -14, 2.0 ** -14 - 2.0 ** -26, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11]]).astype(np.float32)
    for v in vals:
        got = emu.emu_h2_pack_sat(float(v), float(-v))
        with np.errstate(over="ignore"):
            want_hi = np.float16(v)
        if np.isinf(want_hi):
            want_hi = np.float16(65504.0) * (1 if v > 0 else -1)          # cvt.rn.satfinite
        assert got >> 16 == _bits(want_hi), v
        assert (got & 0xffff) == (_bits(want_hi) ^ 0x8000), v
    # fma: one rounding of the exact a*b + c (numpy: exact in float64 for these operands, then one rounding)
    a = rng.integers(0, 256, 4000).astype(np.uint16)                     # subnormal bytes q * 2^-24, as the node test feeds them
    b = np.clip(rng.normal(size=4000) * 10.0 ** rng.uniform(-3, 4.4, 4000), -65000, 65000).astype(np.float16)
    c = (rng.normal(size=4000) * 10.0 ** rng.uniform(-7, 4, 4000)).astype(np.float16)
    for k in range(4000):
        ah, bh, ch = _h(a[k]), b[k], c[k]
        exact = float(ah) * float(bh) + float(ch)
        with np.errstate(over="ignore"):
            want = np.float16(exact)
        got = emu.emu_h2_fma(int(a[k]) | (int(a[k]) << 16), _bits(bh) | (_bits(-bh) << 16), _bits(ch) | (_bits(ch) << 16))
        assert (got & 0xffff) == _bits(want), (k, ah, bh, ch)
    # min and the lane sum
    for _ in range(2000):
        x, y = np.float16(rng.normal() * 100), np.float16(rng.normal() * 100)
        assert emu.emu_h2_min(_bits(x), _bits(y)) & 0xffff == _bits(min(x, y))
        s = emu.emu_h2_lane_sum(_bits(x) | (_bits(y) << 16))
        assert (s & 0xffff) == (s >> 16) == _bits(np.float16(float(x) + float(y)))
    assert emu.emu_h2_lane_sum(0x7c00 | (0xfc00 << 16)) & 0x7fff > 0x7c00      # inf - inf = NaN (reads as a hit, see aob_bvh.cuh)


def _make_node(rng, scale, center, thin):
    """One Node8 record (80 bytes) with 8 internal child slots on random sub-boxes of the grid."""
    ext = scale * np.array([1.0, thin[0], thin[1]])[rng.permutation(3)]
    p = (center - 0.5 * ext).astype(np.float32)
    e = np.zeros(3, dtype=np.uint8)
    for a in range(3):
        s = np.float32(ext[a] / 255.0 * 1.000001)
        b = int(np.array([s], dtype=np.float32).view(np.uint32)[0])
        e[a] = ((b >> 23) & 0xff) + (1 if b & 0x7fffff else 0)
    em = max(int(e.max()), 24)
    e = np.maximum(e, em - 14).astype(np.uint8)                          # the builder's spread rule (kExpSpread)
    q = np.zeros((3, 8, 2), dtype=np.uint8)
    for a in range(3):
        lo = rng.integers(0, 255, 8)
        hi = np.minimum(255, lo + rng.integers(0, 160, 8) * (rng.random(8) < 0.85))     # some slabs of zero thickness
        q[a, :, 0], q[a, :, 1] = lo, hi
    node = np.zeros(80, dtype=np.uint8)
    node[0:12] = p.view(np.uint8)
    node[12:15] = e
    node[15] = 0xff                                                      # all slots internal
    node[24:32] = [0x20 | (24 + k) for k in range(8)]
    node[32:80] = q.reshape(-1)
    scale_f = np.array([np.array([int(x) << 23], dtype=np.uint32).view(np.float32)[0] for x in e], dtype=np.float64)
    lo = p.astype(np.float64)[:, None] + q[:, :, 0].astype(np.float64) * scale_f[:, None]
    hi = p.astype(np.float64)[:, None] + q[:, :, 1].astype(np.float64) * scale_f[:, None]
    return node, lo, hi                                                  # lo/hi: (3, 8) exact child boxes


def _rays_at(rng, lo, hi, n):
    """Rays aimed at (or just past) the children of the node, from far and near, steep and grazing."""
    blo, bhi = lo.min(axis=1), hi.max(axis=1)
    size = float((bhi - blo).max())
    k = rng.integers(0, 8, n)
    u = rng.random((n, 3))
    overshoot = np.where(rng.random((n, 1)) < 0.5, 0.0, 0.02) * rng.normal(size=(n, 3))       # half the targets sit just outside
    target = lo[:, k].T + u * (hi[:, k] - lo[:, k]).T + overshoot * size
    snap = rng.random(n) < 0.3                                           # on a face, an edge or a corner of the child box
    face = np.where(rng.random((n, 3)) < 0.5, lo[:, k].T, hi[:, k].T)
    pick = rng.random((n, 3)) < 0.5
    target = np.where(snap[:, None] & pick, face, target)
    d = rng.normal(size=(n, 3))
    graze = rng.random(n) < 0.35
    d[graze, rng.integers(0, 3)] *= 10.0 ** rng.uniform(-3.5, -1, graze.sum())              # down to |d_a| ~ 3e-4 > 2^-12
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dist = size * 10.0 ** rng.uniform(-2, 4, n)
    dist[rng.random(n) < 0.1] = 0.0                                      # origin inside / on the box
    o = target - d * dist[:, None]
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3], rays[:, 4:7] = o, d
    rays[:, 7] = np.where(rng.random(n) < 0.3, dist * rng.uniform(0.5, 1.5, n) + 1e-30, dist * 1e3 + size * 1e3)
    return rays


def _exact_hits(rays, lo, hi):
    """fp64 slab test of the fp32 ray against the exact child boxes -> (n, 8) bool."""
    o = rays[:, 0:3].astype(np.float64)[:, :, None]
    d = rays[:, 4:7].astype(np.float64)[:, :, None]
    tmin, tmax = rays[:, 3].astype(np.float64)[:, None], rays[:, 7].astype(np.float64)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (lo[None] - o) / d, (hi[None] - o) / d
    tn, tf = np.minimum(t1, t2), np.maximum(t1, t2)
    par = d == 0.0
    inside = (o >= lo[None]) & (o <= hi[None])
    tn = np.where(par, np.where(inside, -np.inf, np.inf), tn)
    tf = np.where(par, np.where(inside, np.inf, -np.inf), tf)
    return np.maximum(tn.max(axis=1), tmin) <= np.minimum(tf.min(axis=1), tmax)


@pytest.mark.parametrize("variant", [0, 1])                             # 0 = packed fp16, 1 = fp32
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_node_tests_never_cull_a_box_the_exact_ray_touches(emu, variant, seed):
    rng = np.random.default_rng(seed)
    checked = false_pos = exact_total = 0
    for _ in range(250):
        scale = 10.0 ** rng.uniform(-4, 4)
        center = rng.normal(size=3) * scale * 10.0 ** rng.uniform(-1, 3)                     # up to 1000 node sizes from the origin
        thin = 10.0 ** rng.uniform(-5, 0, 2) * (rng.random(2) < 0.7) + 1e-30 * 0             # flat and needle-like nodes, exactly flat too
        node, lo, hi = _make_node(rng, scale, center, thin)
        rays = _rays_at(rng, lo, hi, 2000)
        masks = np.zeros(len(rays), dtype=np.uint32)
        wide = np.zeros(len(rays), dtype=np.uint8)
        emu.emu_node_test(node.ctypes.data, rays.ctypes.data, len(rays), variant, masks.ctypes.data, wide.ctypes.data)
        ok = wide == 0 if variant == 0 else np.ones(len(rays), dtype=bool)                   # the fp16 test's precondition
        got = ((masks[:, None] >> (24 + np.arange(8))[None, :]) & 1).astype(bool)
        want = _exact_hits(rays, lo, hi)
        missed = want & ~got & ok[:, None]
        assert not missed.any(), (variant, seed, np.argwhere(missed)[:5], node.tolist())
        checked += int(ok.sum())
        exact_total += int((want & ok[:, None]).sum())
        false_pos += int((got & ~want & ok[:, None]).sum())
    assert checked > 400000 and exact_total > 200000
    # Padding is not free, but it must stay bounded.  Extra boxes reported per box truly hit, on this
    # adversarial set (half the rays aimed 2 % of a node size past a child, nodes up to 1000 sizes from
    # the origin, rays from up to 10^4 sizes away): 0.29 for fp32, 0.58 for fp16.  On AO rays the fp16
    # test visits 1-4 % more nodes (profiles/emu_node_test_compare.py).
    assert false_pos < (0.8 if variant == 0 else 0.4) * exact_total
