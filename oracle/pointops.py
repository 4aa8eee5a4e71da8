"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the third-party point-set ops the reference calls.

PARITY UNPINNED: torch_cluster 1.5.9 (`fps`, `radius`, `knn`) and torch_geometric 1.7.2 (`PointConv`,
`knn_interpolate`, `global_max_pool`) are not vendored in /root/reference nor installed here; these functions restate
their published semantics at the reference's call sites:
    components/pointnet2.py:26      idx = fps(pos, batch, ratio)
    components/pointnet2.py:28-29   row, col = radius(pos, pos[idx], r, batch, batch[idx], max_num_neighbors=64)
    components/pointnet2.py:31      PointConv(nn)(x, (pos, pos[idx]), edge_index)
    components/pointnet2.py:49      global_max_pool(x, batch)
    components/pointnet2.py:72      knn_interpolate(x, pos, pos_skip, batch, batch_skip, k)

All distances are fp32 with individually rounded products and sums, ((dx*dx + dy*dy) + dz*dz), which numpy float32
arithmetic gives and the CUDA kernels reproduce with __fmul_rn/__fadd_rn.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def sqdist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """fp32 ((dx*dx + dy*dy) + dz*dz); a [...,3], b broadcastable."""
    d = (a.astype(F32) - b.astype(F32)).astype(F32)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return ((dx * dx + dy * dy).astype(F32) + dz * dz).astype(F32)


def batch_to_ptr(batch: np.ndarray, num_graphs: int | None = None) -> np.ndarray:
    if num_graphs is None:
        num_graphs = int(batch.max()) + 1 if batch.size else 0
    counts = np.bincount(batch, minlength=num_graphs)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


def fps_counts(ptr: np.ndarray, ratio: float) -> np.ndarray:
    """torch_cluster fps: out sizes = ceil(deg.float() * ratio)."""
    n = np.diff(ptr).astype(F32)
    return np.ceil(n * F32(ratio)).astype(np.int64)


def fps(pos: np.ndarray, ptr: np.ndarray, ratio: float, start: np.ndarray | None = None) -> np.ndarray:
    """Farthest point sampling per cloud.  start[b] = local index of the first sample (the reference draws it at
    random, components/pointnet2.py:26 default random_start=True; it is injected here).  Ties -> lowest index."""
    pos = np.ascontiguousarray(pos, dtype=F32)
    counts = fps_counts(ptr, ratio)
    out = []
    for b in range(len(ptr) - 1):
        p = pos[ptr[b]:ptr[b + 1]]
        n, m = len(p), int(counts[b])
        if n == 0 or m == 0:
            continue
        cur = 0 if start is None else int(min(max(int(start[b]), 0), n - 1))
        dist = np.full(n, np.inf, dtype=F32)
        sel = np.empty(m, dtype=np.int64)
        sel[0] = cur
        for s in range(1, m):
            dist = np.minimum(dist, sqdist(p, p[cur]))
            cur = int(np.argmax(dist))  # first maximum
            sel[s] = cur
        out.append(sel + ptr[b])
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def ball_query(x: np.ndarray, y: np.ndarray, ptr_x: np.ndarray, ptr_y: np.ndarray, r: float, K: int = 64):
    """First K points (index order) of the same cloud with d2 < float32(r*r).  Returns nbr i64[M,K] (-1 padded), cnt."""
    x = np.ascontiguousarray(x, dtype=F32)
    y = np.ascontiguousarray(y, dtype=F32)
    r2 = F32(float(r) * float(r))
    M = len(y)
    nbr = np.full((M, K), -1, dtype=np.int64)
    cnt = np.zeros(M, dtype=np.int32)
    for b in range(len(ptr_x) - 1):
        xs = x[ptr_x[b]:ptr_x[b + 1]]
        for q in range(ptr_y[b], ptr_y[b + 1]):
            hit = np.nonzero(sqdist(xs, y[q]) < r2)[0][:K]
            nbr[q, :len(hit)] = hit + ptr_x[b]
            cnt[q] = len(hit)
    return nbr, cnt


def radius_pairs(nbr: np.ndarray, cnt: np.ndarray):
    """(row, col) as torch_cluster.radius returns them: row = query index, col = point index."""
    row = np.repeat(np.arange(len(cnt), dtype=np.int64), cnt)
    col = np.concatenate([nbr[i, :c] for i, c in enumerate(cnt)]) if len(cnt) else np.zeros(0, np.int64)
    return row, col.astype(np.int64)


def knn(x: np.ndarray, y: np.ndarray, ptr_x: np.ndarray, ptr_y: np.ndarray, k: int):
    """k nearest x (same cloud) per y, ascending d2, ties -> lower index.  idx -1 / d2 inf padded."""
    x = np.ascontiguousarray(x, dtype=F32)
    y = np.ascontiguousarray(y, dtype=F32)
    Ny = len(y)
    idx = np.full((Ny, k), -1, dtype=np.int64)
    d2 = np.full((Ny, k), np.inf, dtype=F32)
    for b in range(len(ptr_x) - 1):
        xs = x[ptr_x[b]:ptr_x[b + 1]]
        if len(xs) == 0:
            continue
        for q in range(ptr_y[b], ptr_y[b + 1]):
            d = sqdist(xs, y[q])
            order = np.argsort(d, kind="stable")[:k]
            idx[q, :len(order)] = order + ptr_x[b]
            d2[q, :len(order)] = d[order]
    return idx, d2


def knn_interpolate(feat: np.ndarray, idx: np.ndarray, d2: np.ndarray) -> np.ndarray:
    """PyG 1.7.2 knn_interpolate: w = 1/clamp(d2, 1e-16); y = sum(x*w)/sum(w), sums in rank order, fp32."""
    feat = np.ascontiguousarray(feat, dtype=F32)
    Ny, k = idx.shape
    num = np.zeros((Ny, feat.shape[1]), dtype=F32)
    den = np.zeros((Ny, 1), dtype=F32)
    first = np.ones(Ny, dtype=bool)
    for p in range(k):
        valid = idx[:, p] >= 0
        w = (F32(1.0) / np.maximum(d2[:, p], F32(1e-16))).astype(F32)[:, None]
        t = (feat[np.where(valid, idx[:, p], 0)] * w).astype(F32)
        init = valid & first
        acc = valid & ~first
        num[init] = t[init]
        den[init] = w[init]
        num[acc] = (num[acc] + t[acc]).astype(F32)
        den[acc] = (den[acc] + w[acc]).astype(F32)
        first &= ~valid
    out = np.zeros_like(num)
    ok = ~first
    out[ok] = (num[ok] / den[ok]).astype(F32)
    return out


def pointconv_edges(nbr: np.ndarray, cnt: np.ndarray):
    """PyG 1.7.2 PointConv(add_self_loops=True) on edge_index=[col,row]: remove_self_loops drops edges whose POINT
    index equals the CENTROID index (two different index spaces), add_self_loops appends (i,i) for i < M.
    Net effect: edges(i) = ballquery(i) U {flat point i}.  Returns (offs i64[M+1], src i64[E]) neighbours first,
    self loop last."""
    M = len(cnt)
    lists = []
    for i in range(M):
        js = [int(j) for j in nbr[i, :cnt[i]] if int(j) != i]
        js.append(i)
        lists.append(js)
    offs = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    src = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if M else np.zeros(0, np.int64)
    return offs, src


def pointconv_edge_features(x_feat, pos_x, pos_y, offs, src):
    """message input cat[x_j, pos_j - pos_i] (PyG PointConv.message)."""
    pos_x = np.ascontiguousarray(pos_x, dtype=F32)
    pos_y = np.ascontiguousarray(pos_y, dtype=F32)
    dst = np.repeat(np.arange(len(offs) - 1), np.diff(offs))
    rel = (pos_x[src] - pos_y[dst]).astype(F32)
    if x_feat is None:
        return rel
    return np.concatenate([np.ascontiguousarray(x_feat, dtype=F32)[src], rel], axis=1)


def segment_max(rows: np.ndarray, offs: np.ndarray) -> np.ndarray:
    out = np.zeros((len(offs) - 1, rows.shape[1]), dtype=F32)
    for i in range(len(offs) - 1):
        if offs[i + 1] > offs[i]:
            out[i] = rows[offs[i]:offs[i + 1]].max(axis=0)
    return out


def scatter(src: np.ndarray, index: np.ndarray, dim_size: int, reduce: str) -> np.ndarray:
    """torch_scatter.scatter(src [C,N], index [N], dim=-1, dim_size, reduce); empty slots are 0.
    ref networks/conv_implicit_wnf.py:92-94, components/gridding.py:32-35."""
    src = np.ascontiguousarray(src, dtype=F32)
    C, N = src.shape
    out = np.zeros((C, dim_size), dtype=F32)
    if N == 0:
        return out
    if reduce in ("max", "min"):
        fill = -np.inf if reduce == "max" else np.inf
        tmp = np.full((C, dim_size), fill, dtype=F32)
        (np.maximum if reduce == "max" else np.minimum).at(tmp, (slice(None), index), src)
        touched = np.zeros(dim_size, dtype=bool)
        touched[index] = True
        out[:, touched] = tmp[:, touched]
        return out
    np.add.at(out, (slice(None), index), src)
    if reduce == "mean":
        cnt = np.bincount(index, minlength=dim_size).astype(F32)
        out = (out / np.maximum(cnt, 1)[None, :]).astype(F32)
    return out
