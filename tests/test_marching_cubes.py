"""Marching cubes + ggm lookup: CUDA vs the sequential C oracle (bit-exact faces / vertices), plus topological
invariants of the oracle itself (parity vs scikit-image is unpinned: SURVEY.md section 8c)."""
from collections import Counter

import numpy as np
import pytest
import torch

from oracle import postproc


def _sphere(n, r=0.7, inside_high=True):
    z, y, x = np.meshgrid(*[np.linspace(-1, 1, n)] * 3, indexing="ij")
    d = np.sqrt(x * x + y * y + z * z)
    return ((r - d) if inside_high else (d - r)).astype(np.float32) + 0.5


def _torus(n):
    z, y, x = np.meshgrid(*[np.linspace(-1, 1, n)] * 3, indexing="ij")
    return (0.25 - np.sqrt((np.sqrt(x * x + y * y) - 0.6) ** 2 + z * z)).astype(np.float32) + 0.5


def _noise(shape, seed):
    # smooth-ish random field with many ambiguous faces
    import scipy.ndimage as ni
    v = np.random.default_rng(seed).normal(size=shape).astype(np.float32)
    return (ni.gaussian_filter(v, 0.7) * 3 + 0.5).astype(np.float32)


def _topology(verts, faces):
    und, dirc = Counter(), Counter()
    for a, b, c in faces:
        for e in ((a, b), (b, c), (c, a)):
            und[(min(e), max(e))] += 1
            dirc[e] += 1
    return len(verts) - len(und) + len(faces), set(und.values()), max(dirc.values())


def test_oracle_sphere_is_closed_manifold():
    v, f, n, val = postproc.marching_cubes(_sphere(40), 0.5, (1 / 39,) * 3)
    euler, shared, dmax = _topology(v, f)
    assert euler == 2 and shared == {2} and dmax == 1
    assert v.dtype == np.float64 and f.dtype == np.int32
    assert np.abs(np.linalg.norm(v - 0.5, axis=1) - 0.35).max() < 0.02
    # first vertex belongs to the first intersected cell of a axis0->axis1->axis2 scan: vertex ids grow with axis0
    assert np.all(np.diff(np.floor(v[:, 0] * 39 + 1e-6)[np.argsort(np.arange(len(v)))]) >= -1)


def test_oracle_torus_euler_zero():
    v, f, _, _ = postproc.marching_cubes(_torus(48), 0.5)
    euler, shared, dmax = _topology(v, f)
    assert euler == 0 and shared == {2} and dmax == 1


def test_oracle_noise_is_watertight_inside():
    vol = _noise((20, 21, 22), 1)
    v, f, _, _ = postproc.marching_cubes(vol, 0.5)
    und = Counter()
    for a, b, c in f:
        for e in ((a, b), (b, c), (c, a)):
            und[(min(e), max(e))] += 1
    # edges with a single incident face may only lie on the volume boundary
    for (a, b), k in und.items():
        assert k in (1, 2)
        if k == 1:
            on_border = lambda p: np.any(p == 0) or p[0] == 19 or p[1] == 20 or p[2] == 21
            assert on_border(v[a]) and on_border(v[b])


def test_oracle_winding_and_errors():
    vol = _sphere(24)
    v, f_asc, _, _ = postproc.marching_cubes(vol, 0.5, gradient_direction="ascent")
    _, f_desc, _, _ = postproc.marching_cubes(vol, 0.5, gradient_direction="descent")
    assert np.array_equal(f_asc, f_desc[:, ::-1])
    p = v[f_desc]
    nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert (np.sum(nrm * (p.mean(1) - 11.5), 1) > 0).all()  # descent: object greater than exterior -> outward normals
    with pytest.raises(ValueError):
        postproc.marching_cubes(vol, 5.0)
    with pytest.raises(ValueError):
        postproc.marching_cubes(vol, -5.0)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["sphere", "torus", "noise", "noise_big", "slab", "ragged"])
@pytest.mark.parametrize("direction", ["ascent", "descent"])
def test_mc_matches_oracle_bit_exact(dev, case, direction):
    from garmentnets_b200 import ops
    vol = {"sphere": lambda: _sphere(33), "torus": lambda: _torus(40), "noise": lambda: _noise((17, 19, 23), 2),
           "noise_big": lambda: _noise((64, 64, 64), 3), "slab": lambda: _sphere(16, 2.5),
           "ragged": lambda: _noise((2, 5, 9), 4)}[case]()
    if case == "slab":
        vol[:, :, :8] = 0.0  # surface cut by a plane that coincides with grid points (values exactly at / below level)
        vol[:, :, 8:] = 1.0
        vol[3, 4, 8] = 0.5
    n = vol.shape[-1]
    spacing = (1 / (n - 1),) * 3
    ref = postproc.predict_tail(vol, 0.5, 0.5, direction)
    ggm = ops.gaussian_gradient_magnitude(torch.from_numpy(vol).to(dev), 0.5)
    verts, faces, normals, values, ggm_at = ops.marching_cubes(torch.from_numpy(vol).to(dev), 0.5, spacing, direction, ggm)
    assert np.array_equal(faces.cpu().numpy(), ref["faces"])  # topology bit-exact, including vertex numbering
    assert np.array_equal(verts.cpu().numpy(), ref["verts"])
    assert np.array_equal(values.cpu().numpy(), ref["volume_value"])
    assert np.array_equal(normals.cpu().numpy(), ref["normals"])
    assert np.abs(ggm_at.cpu().numpy() - ref["volume_gradient_magnitude"]).max() <= 1e-6


def _heavy_noise(shape, seed):
    """heavy-tailed signed noise: many body-diagonal configurations whose interior test says "tunnel"."""
    rng = np.random.default_rng(seed)
    return (rng.uniform(0.02, 1.0, shape) ** 3 * rng.choice([-1.0, 1.0], shape) + 0.5).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_mc33_tunnel_cells_match_oracle_bit_exact(dev, seed):
    """MC33 interior test + tunnel tilings on the device: a field with hundreds of tunnel cells (4.1.2, 6.1.2, 7.4.2,
    10.1.2, 12.1.2, 13.5.2 all occur), faces / vertices / normals / values bit-exact against the oracle."""
    import ctypes
    from garmentnets_b200 import ops
    vol = _heavy_noise((40, 41, 43), seed)
    ref = postproc.predict_tail(vol, 0.5, 0.5, "ascent")
    lib = postproc._lib()
    lib.mc_oracle_tunnel_cells.restype = ctypes.c_int64
    assert int(lib.mc_oracle_tunnel_cells()) > 100
    vt = torch.from_numpy(vol).to(dev)
    verts, faces, normals, values, _ = ops.marching_cubes(vt, 0.5, (1 / 42,) * 3, "ascent")
    assert np.array_equal(faces.cpu().numpy(), ref["faces"])
    assert np.array_equal(verts.cpu().numpy(), ref["verts"])
    assert np.array_equal(values.cpu().numpy(), ref["volume_value"])
    assert np.array_equal(normals.cpu().numpy(), ref["normals"])


@pytest.mark.gpu
def test_mc_batch32_at_128_cubed_bit_exact(dev):
    """BASELINE.json configs[2] size of the marching-cubes stage: 32 volumes of 128^3 in one batched call, every mesh
    bit-exact against the oracle (smooth blobs + a noise component so that ambiguous / tunnel cells occur)."""
    import scipy.ndimage as ni
    from garmentnets_b200 import ops
    rng = np.random.default_rng(11)
    base = np.stack([_sphere(128, 0.55 + 0.01 * i) for i in range(4)])
    vols = np.empty((32, 128, 128, 128), np.float32)
    for i in range(32):
        noise = ni.gaussian_filter(rng.normal(size=(128, 128, 128)).astype(np.float32), 1.0)
        vols[i] = base[i % 4] + (0.6 + 0.1 * (i % 5)) * noise
    vt = torch.from_numpy(vols).to(dev)
    res = ops.marching_cubes_batch(vt, 0.5, (1 / 127,) * 3, "ascent", None)
    for i in range(32):
        v, f, n, val = postproc.marching_cubes(vols[i], 0.5, (1 / 127,) * 3, "ascent")
        verts, faces, normals, values, _ = res[i]
        assert np.array_equal(faces.cpu().numpy(), f), i
        assert np.array_equal(verts.cpu().numpy(), v.astype(np.float32)), i
        assert np.array_equal(values.cpu().numpy(), val), i


@pytest.mark.gpu
def test_mc_errors(dev):
    from garmentnets_b200 import ops
    vol = torch.from_numpy(_sphere(16)).to(dev)
    with pytest.raises(ValueError):
        ops.marching_cubes(vol, 9.0)
    flat = torch.zeros((8, 8, 8), device=dev)
    flat[0, 0, 0] = 1.0  # level inside the range but exactly at a value -> still produces a tiny surface
    v, f, *_ = ops.marching_cubes(flat, 0.5)
    assert len(v) == 3 and len(f) == 1


@pytest.mark.gpu
def test_mc_full_size_properties(dev):
    """128^3 (BASELINE.json size): size-independent properties -- closed orientable manifold, Euler characteristic."""
    from garmentnets_b200 import ops
    vol = torch.from_numpy(_torus(128)).to(dev)
    verts, faces, *_ = ops.marching_cubes(vol, 0.5, (1 / 127,) * 3)
    f = faces.cpu().numpy().astype(np.int64)
    V = len(verts)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    directed = e[:, 0] * V + e[:, 1]
    assert len(np.unique(directed)) == len(directed)  # consistently oriented
    und = np.minimum(e[:, 0], e[:, 1]) * V + np.maximum(e[:, 0], e[:, 1])
    _, counts = np.unique(und, return_counts=True)
    assert np.all(counts == 2)  # watertight
    assert V - len(counts) + len(f) == 0  # torus
    assert f.max() == V - 1 and f.min() == 0
    # first-use numbering: the first reference to vertex k precedes the first reference to vertex k+1
    first = np.full(V, len(f) * 3, dtype=np.int64)
    flat = f[:, ::-1].reshape(-1)  # 'ascent' stores (c,b,a); creation order is a,b,c
    np.minimum.at(first, flat, np.arange(len(flat)))
    assert np.all(np.diff(first) > 0)


@pytest.mark.gpu
def test_mc_batch_equals_single_and_reports_errors(dev):
    """One-synchronisation batch API == per-volume API; per-volume errors are returned, not raised."""
    from garmentnets_b200 import ops
    vols = np.stack([_sphere(24), _torus(24), np.full((24, 24, 24), 0.1, np.float32), _noise((24, 24, 24), 7)])
    vt = torch.from_numpy(vols).to(dev)
    ggm = ops.gaussian_gradient_magnitude_batched(vt, 0.5)
    for i in range(4):
        assert torch.equal(ggm[i], ops.gaussian_gradient_magnitude(vt[i], 0.5))
    res = ops.marching_cubes_batch(vt, 0.5, (1 / 23,) * 3, "ascent", ggm)
    assert isinstance(res[2], ValueError)
    for i in (0, 1, 3):
        single = ops.marching_cubes(vt[i], 0.5, (1 / 23,) * 3, "ascent", ggm[i])
        for a, b in zip(res[i], single):
            assert torch.equal(a, b)
