"""Host-side multi-rank logic on CPU: world_size-2 gloo processes (sharding, metric all-gather, padded mesh gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from garmentnets_b200 import dist as gd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = gd.shard_range(7, rank, world)
        recs = gd.gather_metrics({"n_volumes": hi - lo, "elapsed_ms": 10.0 * (rank + 1), "sum_verts": 100 * (rank + 1),
                                  "sum_faces": 200 * (rank + 1), "checksum": float(rank)})
        summary = gd.summarize(recs)
        g = torch.Generator().manual_seed(rank)
        meshes = []
        for i in range(2):
            V, F = 5 + 3 * rank + i, 4 + rank + 2 * i
            meshes.append({"verts": torch.rand(V, 3, generator=g), "warp_field": torch.rand(V, 3, generator=g),
                           "faces": torch.randint(0, V, (F, 3), generator=g, dtype=torch.int32)})
        allm = gd.gather_meshes(meshes)
        ok = all(torch.equal(allm[rank][i][k], meshes[i][k]) for i in range(2) for k in ("verts", "faces", "warp_field"))
        # packed form (what predict() leaves on the device): one payload per dtype for the whole batch
        import numpy as np
        vptr = np.concatenate([[0], np.cumsum([m["verts"].shape[0] for m in meshes])]).astype(np.int64)
        fptr = np.concatenate([[0], np.cumsum([m["faces"].shape[0] for m in meshes])]).astype(np.int64)
        packed = {"verts": torch.cat([m["verts"] for m in meshes]), "faces": torch.cat([m["faces"] for m in meshes])}
        gp = gd.gather_packed(packed, torch.cat([m["warp_field"] for m in meshes]), vptr, fptr)
        for r in range(world):
            for i in range(2):
                v0, v1, f0, f1 = gp["vptr"][r][i], gp["vptr"][r][i + 1], gp["fptr"][r][i], gp["fptr"][r][i + 1]
                ok = ok and torch.equal(gp["verts"][r][v0:v1], allm[r][i]["verts"])
                ok = ok and torch.equal(gp["warp_field"][r][v0:v1], allm[r][i]["warp_field"])
                ok = ok and torch.equal(gp["faces"][r][f0:f1], allm[r][i]["faces"])
        shapes = [[tuple(m["verts"].shape) + tuple(m["faces"].shape) for m in lst] for lst in allm]
        q.put((rank, (lo, hi), summary, ok, shapes))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            blocks = [gd.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, s0, ok0, sh0), (r1, b1, s1, ok1, sh1) = res
    assert b0 == (0, 4) and b1 == (4, 7)
    assert s0 == s1  # every rank sees the same whole-job summary
    assert s0["n_volumes"] == 7 and s0["elapsed_ms"] == 20.0 and s0["sum_verts"] == 300 and s0["sum_faces"] == 600
    assert abs(s0["volumes_per_s"] - 7 / 0.02) < 1e-9  # units of all ranks / max time over ranks
    assert ok0 and ok1 and sh0 == sh1
    assert sh0 == [[(5, 3, 4, 3), (6, 3, 6, 3)], [(8, 3, 5, 3), (9, 3, 7, 3)]]


def test_numa_binding_reads_sysfs(tmp_path):
    """bind_to_gpu_numa_node's sysfs walk on a fake tree: the process ends up on the node's CPUs that it was allowed to use;
    a node of -1 (single-socket hosts, VMs) or a missing entry leaves the affinity alone."""
    import os
    from garmentnets_b200 import dist as gd

    assert gd._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    before = os.sched_getaffinity(0)
    some = sorted(before)[: max(1, len(before) // 2)]
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    (node / "cpulist").write_text(",".join(str(c) for c in some) + "\n")
    try:
        assert gd._bind_to_numa_node_of("0000:1b:00.0", str(tmp_path)) == 1
        assert os.sched_getaffinity(0) == set(some)
    finally:
        os.sched_setaffinity(0, before)
    (dev / "numa_node").write_text("-1\n")
    assert gd._bind_to_numa_node_of("0000:1b:00.0", str(tmp_path)) is None
    assert gd._bind_to_numa_node_of("0000:ff:00.0", str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before
