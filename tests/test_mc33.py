"""Marching cubes 33 structure of the oracle (CPU; the CUDA kernel is compared bit-for-bit with this oracle in
tests/test_marching_cubes.py): face test + interior test, sub-cases 4.1.2 / 6.1.2 / 7.4.2 / 10.1.2 / 12.1.2 / 13.5.2 (tunnels)
and the extra centre vertices, exhaustively over all 256 cube codes x random ambiguity resolutions on single cells:

* every tiling is an oriented 2-manifold with boundary, its boundary lies in the cube faces and no other edge does
  (so two cells sharing a face can never disagree or collide);
* Euler characteristic = (#contour loops) - 2 (#tunnels): caps are discs, tunnels are annuli;
* the interior decision equals the TRUE connectivity of the trilinear interpolant inside the cell, established by brute
  force on a sampled 33^3 lattice -- this pins the interior test to the mathematics (Chernyaev's MC33), independently of
  both implementations;
* multi-cell noise volumes are watertight away from the volume boundary, tunnel cells included.

PARITY UNPINNED vs scikit-image: its Lewiner tables cannot be recited offline; what is checked here is the topology those
tables encode (ref predict.py:172-177)."""
import collections
import ctypes

import numpy as np
import pytest

from oracle import postproc

CORNER = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EC = [(0, 1), (1, 2), (3, 2), (0, 3), (4, 5), (5, 6), (7, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
FC = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (3, 7, 6, 2), (0, 4, 7, 3), (1, 2, 6, 5)]
FLT_EPS = float(np.finfo(np.float32).eps)
CIDX = {c: i for i, c in enumerate(CORNER)}


def _eb(a, b):
    return next(e for e, (p, q) in enumerate(EC) if {p, q} == {a, b})


FE = [[_eb(f[i], f[(i + 1) % 4]) for i in range(4)] for f in FC]


def cell_volume(vals):
    a = np.zeros((2, 2, 2), np.float32)
    for c, (x, y, z) in enumerate(CORNER):
        a[z, y, x] = vals[c]
    return a


# ---- independent restatement of the combinatorics (expected loop / tunnel structure) ---------------------------------
def ambiguous_faces(idx):
    am = 0
    for f in range(6):
        s = [(idx >> c) & 1 for c in FC[f]]
        if s[0] == s[2] and s[1] == s[3] and s[0] != s[1]:
            am |= 1 << f
    return am


def face_bits(idx, v):
    fb = 0
    am = ambiguous_faces(idx)
    for f in range(6):
        if (am >> f) & 1:
            a, b, c, d = [float(v[x]) for x in FC[f]]
            pp, nn = (a * c, b * d) if a > 0 else (b * d, a * c)
            if pp - nn > -FLT_EPS:
                fb |= 1 << f
    return fb


def loops_and_regions(idx, fb):
    am = ambiguous_faces(idx)
    succ = {}
    par = list(range(8))

    def find(x):
        while par[x] != x:
            x = par[x]
        return x

    for a, b in EC:
        if ((idx >> a) & 1) == ((idx >> b) & 1):
            par[find(a)] = find(b)
    for f in range(6):
        s = [(idx >> c) & 1 for c in FC[f]]
        if sum(s) in (0, 4):
            continue
        fe = FE[f]
        if not (am >> f) & 1:
            i0 = next(i for i in range(4) if s[i] and not s[(i + 3) & 3])
            j0 = next(i for i in range(4) if s[i] and not s[(i + 1) & 3])
            succ[fe[j0]] = fe[(i0 + 3) & 3]
        elif (fb >> f) & 1:
            for n in range(4):
                if not s[n]:
                    succ[fe[(n + 3) & 3]] = fe[n]
            pos = [c for c in FC[f] if (idx >> c) & 1]
            par[find(pos[0])] = find(pos[1])
        else:
            for p in range(4):
                if s[p]:
                    succ[fe[p]] = fe[(p + 3) & 3]
            neg = [c for c in FC[f] if not (idx >> c) & 1]
            par[find(neg[0])] = find(neg[1])
    seen, loops = set(), []
    for e0 in range(12):
        if e0 in succ and e0 not in seen:
            poly, e = [], e0
            while e not in seen:
                seen.add(e)
                poly.append(e)
                e = succ[e]
            loops.append(poly)
    regions = collections.defaultdict(list)
    for c in range(8):
        regions[find(c)].append(c)
    return loops, list(regions.values())


def manhattan(p, q):
    return sum(abs(CORNER[p][i] - CORNER[q][i]) for i in range(3))


def tunnel_candidates(idx, fb):
    """[(loop a, loop b, [(p, q) ...])] for the annular regions MC33 tests."""
    loops, regions = loops_and_regions(idx, fb)
    touch = [[i for i, l in enumerate(loops) if any(EC[e][0] in r or EC[e][1] in r for e in l)] for r in regions]
    annuli = [(r, ls) for r, ls in zip(regions, touch) if len(ls) == 2]
    out = []
    for r, ls in annuli:
        nb = [next(r2 for r2, ls2 in zip(regions, touch) if r2 is not r and l in ls2) for l in ls]
        pairs = sorted((p, q) for p in nb[0] for q in nb[1])
        body = [pq for pq in pairs if manhattan(*pq) == 3]
        face = [pq for pq in pairs if manhattan(*pq) == 2]
        if body:
            out.append((ls[0], ls[1], body))
        elif len(annuli) == 2 and face:
            out.append((ls[0], ls[1], face))
    return loops, out


def trilinear_joined(v, p, q, N=33):
    """Brute force: are corners p and q (same sign) in one connected component of that sign inside the closed cell?"""
    from scipy import ndimage
    g = np.linspace(0.0, 1.0, N)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    F = np.zeros_like(X)
    for c, (cx, cy, cz) in enumerate(CORNER):
        F += float(v[c]) * (X if cx else 1 - X) * (Y if cy else 1 - Y) * (Z if cz else 1 - Z)
    sigma = 1.0 if v[p] > 0 else -1.0
    lab, _ = ndimage.label(sigma * F > 0)
    ip = tuple(x * (N - 1) for x in CORNER[p])
    iq = tuple(x * (N - 1) for x in CORNER[q])
    return lab[ip] > 0 and lab[ip] == lab[iq]


# ---- mesh checks ------------------------------------------------------------------------------------------------------
def mesh_topology(verts, faces):
    """(directed edges all unique, boundary edges, euler characteristic)."""
    de = collections.Counter()
    for t in faces:
        assert len(set(t.tolist())) == 3
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            de[(int(a), int(b))] += 1
    assert max(de.values()) == 1, "an oriented edge is used twice (non-manifold or inconsistent orientation)"
    boundary = [(a, b) for (a, b) in de if (b, a) not in de]
    und = {frozenset(e) for e in de}
    chi = len(np.unique(faces)) - len(und) + len(faces)
    return boundary, und, chi


def in_one_face(p, q, size=1.0):
    return any((p[i] == q[i]) and p[i] in (0.0, size) for i in range(3))


def random_cell(rng, idx):
    mag = rng.uniform(0.02, 1.0, 8) ** rng.choice([1, 3])
    return np.where([(idx >> c) & 1 for c in range(8)], mag, -mag).astype(np.float32)


def test_single_cells_all_codes_manifold_and_euler():
    rng = np.random.default_rng(0)
    seen_tunnel_sizes = collections.Counter()
    centre_vertex_cells = 0
    for idx in range(1, 255):
        for _ in range(24):
            v = random_cell(rng, idx)
            fb = face_bits(idx, v.astype(np.float64))
            loops, cands = tunnel_candidates(idx, fb)
            verts, faces, _, _ = postproc.marching_cubes(cell_volume(v), 0.0, (1.0, 1.0, 1.0), "descent")
            verts = verts[:, ::-1]  # (axis0, axis1, axis2) = (z, y, x) -> x, y, z
            boundary, und, chi = mesh_topology(verts, faces)
            # boundary = the contour segments: as many as there are cube-edge crossings, all inside cube faces
            assert len(boundary) == sum(len(l) for l in loops)
            for a, b in boundary:
                assert in_one_face(verts[a], verts[b])
            # no other edge lies in a cube face
            bset = {frozenset(e) for e in boundary}
            for e in und - bset:
                a, b = tuple(e)
                assert not in_one_face(verts[a], verts[b]), (idx, fb)
            ntun = (len(loops) - chi) // 2
            assert chi == len(loops) - 2 * ntun and ntun in (0, 1)
            assert ntun == 0 or cands, (idx, fb)
            n_edge_verts = sum(len(l) for l in loops)
            centre_vertex_cells += len(verts) > n_edge_verts
            if ntun:
                seen_tunnel_sizes[len(faces)] += 1
    # 4.1.2 (6), 6.1.2 (7 = tube 7 | tube + ...), two-fan tubes 10.1.2 / 12.1.2 (12), 7.4.2 (13)
    assert {6, 7, 12, 13} <= set(seen_tunnel_sizes), seen_tunnel_sizes
    assert centre_vertex_cells > 0


def test_interior_decision_equals_trilinear_topology():
    """Targeted at the tunnel-capable configurations: the oracle's choice (read off the Euler characteristic of its mesh)
    against brute-force connectivity of the trilinear interpolant."""
    rng = np.random.default_rng(1)
    stats = collections.Counter()
    trials = 0
    codes = list(range(1, 255))
    while trials < 700:
        idx = int(rng.choice(codes)) if trials % 4 else int(rng.choice([165, 90]))   # every 4th: case 13
        v = random_cell(rng, idx)
        fb = face_bits(idx, v.astype(np.float64))
        loops, cands = tunnel_candidates(idx, fb)
        if not cands:
            continue
        trials += 1
        truth = any(trilinear_joined(v.astype(np.float64), p, q) for _, _, pairs in cands for p, q in pairs[:1])
        verts, faces, _, _ = postproc.marching_cubes(cell_volume(v), 0.0, (1.0, 1.0, 1.0), "descent")
        _, _, chi = mesh_topology(verts, faces)
        got = (len(loops) - chi) // 2 == 1
        stats[(truth, got)] += 1
        assert truth == got, (idx, fb, v.tolist())
    assert stats[(True, True)] >= 15 and stats[(False, False)] >= 200, stats


def test_thirteen_five_both_kinds_of_tunnel_occur():
    """Case 13.5 has two nested annuli; either one can carry the tunnel (13.5.2 and its sign-inverted form)."""
    rng = np.random.default_rng(2)
    kinds = set()
    for _ in range(30000):
        idx = int(rng.choice([165, 90]))
        v = random_cell(rng, idx)
        fb = face_bits(idx, v.astype(np.float64))
        if bin(fb).count("1") != 3:
            continue
        loops, cands = tunnel_candidates(idx, fb)
        if len(cands) != 2:
            continue
        verts, faces, _, _ = postproc.marching_cubes(cell_volume(v), 0.0, (1.0, 1.0, 1.0), "descent")
        _, _, chi = mesh_topology(verts, faces)
        if chi == len(loops) - 2:
            sign = 1 if v[cands[0][2][0][0]] > 0 else -1
            joined_first = any(trilinear_joined(v.astype(np.float64), p, q) for p, q in cands[0][2][:1])
            kinds.add((sign, joined_first))
            assert len(faces) == 14       # two-fan tube (9 + 4) + one cap
        if len(kinds) >= 2:
            break
    assert len(kinds) >= 2, kinds


def _tunnel_cells():
    lib = postproc._lib()
    lib.mc_oracle_tunnel_cells.restype = ctypes.c_int64
    return int(lib.mc_oracle_tunnel_cells())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_noise_volumes_are_watertight_with_tunnels(seed):
    """Cells sharing a face agree on it for every resolution of the ambiguities, tunnel tilings included: heavy-tailed noise
    (many body-diagonal configurations) gives a mesh that is closed away from the volume boundary."""
    rng = np.random.default_rng(seed)
    n = 14
    vol = (rng.uniform(0.02, 1.0, (n, n, n)) ** 3 * rng.choice([-1.0, 1.0], (n, n, n))).astype(np.float32)
    verts, faces, _, _ = postproc.marching_cubes(vol, 0.0, (1.0, 1.0, 1.0), "descent")
    assert _tunnel_cells() > 0
    boundary, und, chi = mesh_topology(verts, faces)
    for a, b in boundary:
        assert in_one_face(verts[a], verts[b], size=float(n - 1)), "open edge inside the volume"


# ---- the CUDA kernels' look-up tables against the oracle, on the host (no GPU needed) -------------------------------
def _host_tiling(vals, level=0.0):
    from garmentnets_b200 import _lib
    lib = _lib.load()
    cv = (ctypes.c_float * 8)(*[float(x) for x in vals])
    code, ntri, nvert, ncen = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    tri = (ctypes.c_uint8 * 42)()
    order = (ctypes.c_uint8 * 14)()
    cen_n = (ctypes.c_uint8 * 2)()
    cen_loop = (ctypes.c_uint8 * 24)()
    ptr = lambda o: ctypes.cast(ctypes.byref(o) if not isinstance(o, ctypes.Array) else o, ctypes.c_void_p)
    rc = lib.gnb_mc_cell_tiling_host(ptr(cv), ctypes.c_float(level), ptr(code), ptr(ntri), ptr(tri), ptr(nvert), ptr(order),
                                     ptr(ncen), ptr(cen_n), ptr(cen_loop))
    assert rc == 0
    t = np.frombuffer(tri, np.uint8)[:3 * ntri.value].reshape(-1, 3).astype(np.int64)
    o = np.frombuffer(order, np.uint8)[:nvert.value].astype(np.int64)
    return code.value, t, o, ncen.value


def test_kernel_tables_equal_the_oracle_for_every_configuration():
    """Table-driven kernel tiling (built on the host by marching_cubes.cu::build_tables and selected by the host build of
    the kernel's face / interior tests) == the oracle's run-time tracing, triangle for triangle and in the same order,
    for every cube code x face decisions x tunnel decision reached by random cells (incl. heavy-tailed values)."""
    rng = np.random.default_rng(3)
    seen = set()
    tunnels = 0
    for rep in range(40):
        for idx in range(1, 255):
            v = random_cell(rng, idx)
            code, tri, order, ncen = _host_tiling(v)
            assert code & 255 == idx
            verts, faces, _, _ = postproc.marching_cubes(cell_volume(v), 0.0, (1.0, 1.0, 1.0), "descent")
            rank = {int(e): k for k, e in enumerate(order)}
            want = np.array([[rank[int(e)] for e in t] for t in tri], np.int64).reshape(-1, 3)
            assert len(verts) == len(order)
            assert np.array_equal(faces.astype(np.int64), want), (idx, code, v.tolist())
            seen.add(code)
            tunnels += (code >> 14) != 0
    assert tunnels > 30
    assert len({c & 0x3FFF for c in seen}) > 450     # most of the 656 (code, face decision) configurations were reached


def test_kernel_tables_case13_tunnels():
    rng = np.random.default_rng(4)
    kinds = set()
    for _ in range(40000):
        idx = int(rng.choice([165, 90]))
        v = random_cell(rng, idx)
        code, tri, order, ncen = _host_tiling(v)
        if code >> 14:
            verts, faces, _, _ = postproc.marching_cubes(cell_volume(v), 0.0, (1.0, 1.0, 1.0), "descent")
            rank = {int(e): k for k, e in enumerate(order)}
            want = np.array([[rank[int(e)] for e in t] for t in tri], np.int64)
            assert np.array_equal(faces.astype(np.int64), want)
            kinds.add((idx, code >> 14))
            if len(kinds) == 4:
                break
    assert len(kinds) >= 3, kinds
