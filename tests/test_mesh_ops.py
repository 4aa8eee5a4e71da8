"""SURVEY.md section 8f on the device: largest connected component after hole removal (ref eval.py:538-546) and
area-weighted surface sampling (ref common/geometry_util.py:160-223, eval.py:222-243).  The sampling is pinned to the
reference's own functions (tests/golden/mesh_sampling.npz, oracle/make_golden_mesh_sampling.py); the components to a
scipy.sparse.csgraph restatement of the igl calls (igl is not installable offline: parity unpinned, integer semantics)."""
import os

import numpy as np
import pytest
import torch

from oracle import postproc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_sampling.npz")


def _random_mesh(rng, n_verts, n_faces, n_islands):
    """faces drawn inside `n_islands` disjoint vertex ranges (so several components), plus unused vertices."""
    bounds = np.sort(rng.choice(np.arange(1, n_verts), n_islands - 1, replace=False))
    lo = np.concatenate([[0], bounds])
    hi = np.concatenate([bounds, [n_verts]])
    faces = []
    for _ in range(n_faces):
        k = rng.integers(0, n_islands)
        if hi[k] - lo[k] < 3:
            continue
        faces.append(rng.choice(np.arange(lo[k], hi[k]), 3, replace=False))
    return np.asarray(faces, np.int32).reshape(-1, 3)


def test_oracle_components_small_known_answer():
    faces = np.array([[0, 1, 2], [2, 3, 4], [6, 7, 8]], np.int32)   # {0..4}, {5}, {6,7,8}, {9}
    n, idx, sizes, mask = postproc.connected_components(faces, 10)
    assert n == 4 and sizes.tolist() == [5, 1, 3, 1]
    assert idx.tolist() == [0, 0, 0, 0, 0, 1, 2, 2, 2, 3]
    assert mask.tolist() == [True] * 5 + [False] * 5


def test_oracle_doublearea_matches_golden_probabilities():
    """the golden face choices were produced by the reference with this doublearea: re-deriving them with numpy's own
    RandomState reproduces the reference's selection (pins the restated inverse-CDF reading of RandomState.choice)."""
    z = np.load(GOLDEN)
    p = postproc.doublearea(z["verts"], z["faces"])
    p = p / np.sum(p)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    rs = np.random.RandomState(seed=0)
    fi = cdf.searchsorted(rs.random_sample(10000), side="right")
    assert np.array_equal(fi, z["fi_0"])
    uv = rs.uniform(0, 1, size=(10000, 2))
    bad = uv.sum(1) >= 1
    uv[bad] = 1 - uv[bad]
    assert np.array_equal(uv, z["bc_0"][:, :2])


@pytest.mark.gpu
def test_components_match_oracle_batch(dev):
    from garmentnets_b200.common import geometry_util as G
    rng = np.random.default_rng(0)
    meshes = [(_random_mesh(rng, 500, 700, 6), 500), (_random_mesh(rng, 64, 10, 3), 64), (np.zeros((0, 3), np.int32), 5),
              (_random_mesh(rng, 3000, 9000, 2), 3000)]
    vptr = np.concatenate([[0], np.cumsum([n for _, n in meshes])])
    fptr = np.concatenate([[0], np.cumsum([len(f) for f, _ in meshes])])
    faces = torch.from_numpy(np.concatenate([f for f, _ in meshes])).to(dev)
    mask, labels, summary = G.connected_components_batch(faces, vptr, fptr)
    mask, labels = mask.cpu().numpy(), labels.cpu().numpy()
    for b, (f, n) in enumerate(meshes):
        num_cc, idx, sizes, ref_mask = postproc.connected_components(f, n)
        got = labels[vptr[b]:vptr[b + 1]]
        # label = lowest vertex of the component; relabelled in ascending order it is igl's / scipy's numbering
        _, relabel = np.unique(got, return_inverse=True)
        assert np.array_equal(relabel, idx), b
        assert np.array_equal(mask[vptr[b]:vptr[b + 1]], ref_mask), b
        assert summary[b, 0] == num_cc and summary[b, 1] == sizes.max()
    # single-mesh igl-shaped API
    f, n = meshes[0]
    num_cc, cc_idxs, cc_sizes = G.connected_components(torch.from_numpy(f).to(dev), n)
    ref = postproc.connected_components(f, n)
    assert num_cc == ref[0] and np.array_equal(cc_idxs.cpu().numpy(), ref[1]) and np.array_equal(cc_sizes.cpu().numpy(), ref[2])
    assert np.array_equal(G.largest_component_mask(torch.from_numpy(f).to(dev), n).cpu().numpy(), ref[3])


@pytest.mark.gpu
def test_largest_component_of_a_marching_cubes_mesh(dev):
    """eval.py:532-546 end to end on the device: hole removal by a per-vertex threshold, largest component, re-indexing."""
    from garmentnets_b200 import ops
    from garmentnets_b200.common import geometry_util as G
    from garmentnets_b200.common.marching_cubes_util import delete_invalid_verts
    n = 48
    z, y, x = np.meshgrid(*[np.linspace(-1, 1, n)] * 3, indexing="ij")
    vol = (np.maximum(0.35 - np.sqrt((x + 0.4) ** 2 + y ** 2 + z ** 2), 0.2 - np.sqrt((x - 0.5) ** 2 + y ** 2 + z ** 2)) + 0.5)
    vol = vol.astype(np.float32)
    verts, faces, _, _, _ = ops.marching_cubes(torch.from_numpy(vol).to(dev), 0.5, (1 / (n - 1),) * 3, "ascent")
    on = verts[:, 2] > 0.3            # cut both spheres open
    sv, sf = delete_invalid_verts(verts, faces, on)
    mask = G.largest_component_mask(sf, sv.shape[0])
    cv, cf = delete_invalid_verts(sv, sf, mask)
    ref_v, ref_f = postproc.delete_invalid_verts(verts.cpu().numpy(), faces.cpu().numpy(), on.cpu().numpy())
    _, _, _, ref_mask = postproc.connected_components(ref_f, len(ref_v))
    ref_cv, ref_cf = postproc.delete_invalid_verts(ref_v, ref_f, ref_mask)
    assert np.array_equal(cv.cpu().numpy(), ref_cv) and np.array_equal(cf.cpu().numpy(), ref_cf)
    assert 0 < len(ref_cv) < len(ref_v)       # two components, the big sphere's cap survives
    # axis 2 is x: the big sphere sits at x = 0.3 (unit coordinates), the small one at 0.75; the cut big sphere wins
    assert 0.3 < float(cv[:, 2].mean()) < 0.55 and float(cv[:, 2].max()) < 0.6


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n", [(0, 10000), (7, 257)])
def test_mesh_sampling_matches_reference_functions(dev, seed, n):
    from garmentnets_b200.common import geometry_util as G
    z = np.load(GOLDEN)
    verts, faces = torch.from_numpy(z["verts"]).to(dev), torch.from_numpy(z["faces"]).to(dev)
    bc, fi = G.mesh_sample_barycentric(verts, faces, num_samples=n, seed=seed)
    assert fi.dtype == faces.dtype and bc.dtype == torch.float64
    assert np.array_equal(fi.cpu().numpy(), z[f"fi_{seed}"])          # same faces drawn
    assert np.array_equal(bc.cpu().numpy(), z[f"bc_{seed}"])          # same barycentric coordinates, bit for bit
    pts = G.barycentric_interpolation(bc, verts, faces[fi.long()])
    assert pts.dtype == torch.float32 and np.array_equal(pts.cpu().numpy(), z[f"pts_{seed}"])
    fld = G.interpolate_on_faces(bc, fi.long(), faces, torch.from_numpy(z["field"]).to(dev))
    assert np.array_equal(fld.cpu().numpy(), z[f"fld_{seed}"])


@pytest.mark.gpu
def test_mesh_sampling_float64_vertices_and_given_areas(dev):
    from garmentnets_b200.common import geometry_util as G
    z = np.load(GOLDEN)
    v64, faces = torch.from_numpy(z["v64"]).to(dev), torch.from_numpy(z["faces"]).to(dev)
    bc, fi = G.mesh_sample_barycentric(v64, faces, num_samples=500, seed=3)
    assert np.array_equal(fi.cpu().numpy(), z["fi_64"]) and np.array_equal(bc.cpu().numpy(), z["bc_64"])
    pts = G.barycentric_interpolation(bc, v64, faces[fi.long()])
    assert pts.dtype == torch.float64 and np.array_equal(pts.cpu().numpy(), z["pts_64"])
    areas = torch.from_numpy(postproc.doublearea(z["v64"], z["faces"])).to(dev)
    bc2, fi2 = G.mesh_sample_barycentric(v64, faces, num_samples=500, seed=3, face_areas=areas)
    assert torch.equal(fi2, fi) and torch.equal(bc2, bc)
