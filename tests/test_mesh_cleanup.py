"""Mesh clean-up after marching cubes (ref common/marching_cubes_util.py:5-52): the CUDA
compaction against the numpy restatement of the reference, bit for bit (integer work)."""
import numpy as np
import pytest
import torch

from oracle import postproc


def test_oracle_delete_invalid_verts_known_answer():
    """Hand-checked case: two triangles share an edge, vertex 3 is off the surface -> only face 0 survives, vertex ids
    are compacted in ascending order."""
    verts = np.arange(15, dtype=np.float32).reshape(5, 3)
    faces = np.array([[0, 2, 4], [2, 3, 4]], dtype=np.int32)
    on = np.array([True, True, True, False, True])
    v, f = postproc.delete_invalid_verts(verts, faces, on)
    assert np.array_equal(v, verts[[0, 2, 4]])
    assert np.array_equal(f, np.array([[0, 1, 2]], dtype=np.int32)) and f.dtype == np.int32


@pytest.mark.gpu
@pytest.mark.parametrize("V,F,p_on,seed", [(50, 120, 0.8, 0), (5000, 12000, 0.5, 1), (200000, 400000, 0.9, 2), (1000, 3000, 0.0, 3),
                                           (1000, 3000, 1.0, 4), (7, 1, 1.0, 5)])
def test_delete_invalid_verts_matches_reference(dev, V, F, p_on, seed):
    from garmentnets_b200.common.marching_cubes_util import delete_invalid_verts
    rng = np.random.default_rng(seed)
    verts = rng.normal(size=(V, 3)).astype(np.float32)
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int32)
    on = rng.random(V) < p_on
    v_ref, f_ref = postproc.delete_invalid_verts(verts, faces, on)
    v, f = delete_invalid_verts(torch.from_numpy(verts).to(dev), torch.from_numpy(faces).to(dev), torch.from_numpy(on).to(dev))
    assert f.dtype == torch.int32 and v.dtype == torch.float32
    assert np.array_equal(v.cpu().numpy(), v_ref)
    assert np.array_equal(f.cpu().numpy().reshape(-1, 3), f_ref.reshape(-1, 3))


@pytest.mark.gpu
def test_delete_invalid_verts_batch_ragged(dev):
    """Packed batch with local vertex ids, an empty sample, a sample whose faces all die and a large one."""
    from garmentnets_b200.common.marching_cubes_util import delete_invalid_verts_batch
    rng = np.random.default_rng(9)
    Vs, Fs = [300, 0, 40, 70000], [900, 0, 100, 150000]
    faces, ons, refs = [], [], []
    for i, (V, F) in enumerate(zip(Vs, Fs)):
        f = rng.integers(0, max(V, 1), size=(F, 3)).astype(np.int32)
        on = rng.random(V) < (0.0 if i == 2 else 0.7)
        faces.append(f)
        ons.append(on)
        refs.append(postproc.delete_invalid_verts(np.arange(V, dtype=np.int64), f, on))
    vptr = np.concatenate([[0], np.cumsum(Vs)]).astype(np.int64)
    fptr = np.concatenate([[0], np.cumsum(Fs)]).astype(np.int64)
    keep, vf, nv, nf = delete_invalid_verts_batch(torch.from_numpy(np.concatenate(faces)).to(dev), vptr, fptr,
                                                  torch.from_numpy(np.concatenate(ons)).to(dev))
    keep, vf = keep.cpu().numpy(), vf.cpu().numpy()
    for b, (kept_ref, f_ref) in enumerate(refs):
        assert np.array_equal(keep[nv[b]:nv[b + 1]] - vptr[b], kept_ref), b       # surviving local vertex ids, ascending
        assert np.array_equal(vf[nf[b]:nf[b + 1]], f_ref.reshape(-1, 3)), b
    assert nv[-1] == len(keep) and nf[-1] == len(vf)


@pytest.mark.gpu
def test_wnf_to_mesh_matches_oracle(dev):
    """Whole helper (ggm -> marching cubes -> threshold -> clean-up) on a winding-number-like field: two nested sheets,
    one of which has a weak gradient and is removed by the threshold."""
    from garmentnets_b200.common.marching_cubes_util import wnf_to_mesh
    n = 40
    z, y, x = np.meshgrid(*([np.linspace(-1, 1, n)] * 3), indexing="ij")
    r = np.sqrt(x * x + y * y + (z * 1.3) ** 2)
    wnf = (1.0 / (1.0 + np.exp((r - 0.45) * 60.0)) + 0.56 * np.exp(-((r - 0.85) / 0.25) ** 2)).astype(np.float32)
    v_ref, f_ref = postproc.wnf_to_mesh(wnf, 0.5, 0.25)
    v, f = wnf_to_mesh(torch.from_numpy(wnf).to(dev), 0.5, 0.25)
    assert 100 < len(f_ref) < 10000   # the steep inner sheet survives, the weak outer sheets are removed (20180 raw faces)
    assert np.array_equal(f.cpu().numpy(), f_ref)
    assert np.array_equal(v.cpu().numpy(), v_ref.astype(np.float32))


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_cleanup.npz"))


def test_oracle_pinned_to_reference_function():
    """tests/golden/mesh_cleanup.npz was produced by the reference's own delete_invalid_verts
    (oracle/make_golden_mesh_cleanup.py executes its unmodified source): the restatement must reproduce it bit for bit."""
    g = _golden()
    for i in range(int(g["cases"])):
        v, f = postproc.delete_invalid_verts(g[f"verts{i}"], g[f"faces{i}"], g[f"on{i}"])
        assert np.array_equal(v, g[f"valid_verts{i}"]) and np.array_equal(np.asarray(f).reshape(-1, 3), g[f"valid_faces{i}"])


@pytest.mark.gpu
def test_cuda_against_reference_golden(dev):
    from garmentnets_b200.common.marching_cubes_util import delete_invalid_verts
    g = _golden()
    for i in range(int(g["cases"])):
        v, f = delete_invalid_verts(torch.from_numpy(g[f"verts{i}"]).to(dev), torch.from_numpy(g[f"faces{i}"]).to(dev),
                                    torch.from_numpy(g[f"on{i}"]).to(dev))
        assert np.array_equal(v.cpu().numpy(), g[f"valid_verts{i}"])
        assert np.array_equal(f.cpu().numpy().reshape(-1, 3), g[f"valid_faces{i}"])
