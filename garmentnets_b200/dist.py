"""Multi-GPU plumbing: the path shards by independent units -- every stage is per cloud / per
volume (BatchNorm in eval mode, GroupNorm per sample) -- so ranks own contiguous blocks of clouds, weights are
replicated and there is NO data-path collective.  The only communication is one all-gather of a fixed-size per-rank
record for the final metrics and, when the caller wants every mesh on rank 0 (BASELINE config 4), one padded
all-gather of the variable-length meshes.  Works with NCCL (GPU tensors) and gloo (CPU tensors, used by the tests).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

RECORD_FIELDS = ("n_volumes", "elapsed_ms", "sum_verts", "sum_faces", "checksum")


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` units owned by ``rank``; blocks differ in size by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of /sys/devices/system/node/nodeN/cpulist)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int, sysfs: str = "/sys") -> Optional[int]:
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned staging buffer is
    allocated (first touch then places the pinned pages on that node, so the device -> host mesh transfers of a rank
    do not cross the socket interconnect).  One process per GPU: with eight ranks on a two-socket host an unbound rank
    has an even chance of staging through the far socket.  Returns the node, or None when the platform does not say
    (no sysfs entry, node -1, an affinity mask that excludes the node): the process is then left as it is."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = props.pci_bus_id
        if not isinstance(bus, str):   # torch reports domain / bus / device as integers
            bus = f"{int(props.pci_domain_id):04x}:{int(bus):02x}:{int(props.pci_device_id):02x}.0"
    except Exception:
        return None
    return _bind_to_numa_node_of(bus.lower(), sysfs)


def _bind_to_numa_node_of(pci_bus_id: str, sysfs: str = "/sys") -> Optional[int]:
    try:
        with open(os.path.join(sysfs, "bus/pci/devices", pci_bus_id, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")) as f:
            cpus = set(_parse_cpulist(f.read()))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_metrics(record: Dict[str, float]) -> List[Dict[str, float]]:
    """One all-gather of the fixed-size per-rank record; every rank gets the list of all records."""
    world = dist.get_world_size()
    mine = torch.tensor([float(record.get(k, 0.0)) for k in RECORD_FIELDS], dtype=torch.float64, device=_device())
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [dict(zip(RECORD_FIELDS, t.tolist())) for t in out]


def summarize(records: Sequence[Dict[str, float]]) -> Dict[str, float]:
    """Whole-job numbers: units summed over ranks, time = max over ranks (never a wall clock)."""
    n = sum(r["n_volumes"] for r in records)
    t = max(r["elapsed_ms"] for r in records)
    return {"n_volumes": n, "elapsed_ms": t, "volumes_per_s": n / (t / 1e3) if t > 0 else 0.0,
            "sum_verts": sum(r["sum_verts"] for r in records), "sum_faces": sum(r["sum_faces"] for r in records),
            "checksum": sum(r["checksum"] for r in records)}


def gather_meshes(meshes: Sequence[Dict[str, torch.Tensor]]) -> List[List[Dict[str, torch.Tensor]]]:
    """All-gather of this rank's list of meshes ({"verts" f32[V,3], "faces" i32[F,3], "warp_field" f32[V,3]}).
    Two collectives: the per-sample (V, F) counts, then ONE padded buffer per rank.  Returns, on every rank, the
    per-rank lists in rank order.  Every rank must pass the same number of meshes."""
    world = dist.get_world_size()
    dev = _device()
    n = len(meshes)
    counts = torch.tensor([[m["verts"].shape[0], m["faces"].shape[0]] for m in meshes], dtype=torch.int64,
                          device=dev).reshape(n, 2)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    vmax = int(max(int(c[:, 0].max()) if n else 0 for c in all_counts))
    fmax = int(max(int(c[:, 1].max()) if n else 0 for c in all_counts))
    # one float32 payload per rank: [n, vmax, 6] (verts | warp) and one int32 payload [n, fmax, 3]
    vbuf = torch.zeros((n, vmax, 6), dtype=torch.float32, device=dev)
    fbuf = torch.zeros((n, fmax, 3), dtype=torch.int32, device=dev)
    for i, m in enumerate(meshes):
        V, F = m["verts"].shape[0], m["faces"].shape[0]
        vbuf[i, :V, :3] = m["verts"].to(dev)
        vbuf[i, :V, 3:] = m["warp_field"].to(dev)
        fbuf[i, :F] = m["faces"].to(dev)
    all_v = [torch.empty_like(vbuf) for _ in range(world)]
    all_f = [torch.empty_like(fbuf) for _ in range(world)]
    dist.all_gather(all_v, vbuf)
    dist.all_gather(all_f, fbuf)
    out = []
    for r in range(world):
        lst = []
        for i in range(n):
            V, F = int(all_counts[r][i, 0]), int(all_counts[r][i, 1])
            lst.append({"verts": all_v[r][i, :V, :3], "warp_field": all_v[r][i, :V, 3:], "faces": all_f[r][i, :F]})
        out.append(lst)
    return out


def _all_gather_rows(out: torch.Tensor, mine: torch.Tensor) -> None:
    """out[r] = rank r's ``mine`` (same shape on every rank): one collective.  NCCL takes the stacked output tensor; gloo
    (CPU tests) wants the flat form."""
    dist.all_gather_into_tensor(out.view(-1), mine.reshape(-1))


def gather_packed(packed: Dict[str, torch.Tensor], warp_field: torch.Tensor, vptr, fptr) -> Dict[str, object]:
    """All-gather of the PACKED meshes of a whole batch (the layout ``ConvImplicitWNFPipeline.predict`` leaves on the
    device: verts f32[sum V,3], faces i32[sum F,3] with per-sample local ids, warp_field f32[sum V,3], host row offsets
    ``vptr`` / ``fptr`` i64[B+1]).  Three collectives for any batch size: the offsets of every rank (one small
    all-gather, one host read to size the payloads), then ONE ``all_gather_into_tensor`` per dtype over buffers padded to
    the largest rank (float32 [Vmax, 6] = verts | warp field, int32 [Fmax, 3]).  Returns ``{"vptr": [world][B+1],
    "fptr": ..., "verts": [world] views, "warp_field": ..., "faces": ..., "bytes": payload bytes received}``.
    BASELINE.json configs[3] ("NCCL gather of meshes")."""
    world = dist.get_world_size()
    dev = _device()
    offs = torch.cat([torch.as_tensor(vptr, dtype=torch.int64), torch.as_tensor(fptr, dtype=torch.int64)]).to(dev)
    all_offs = torch.empty((world, offs.numel()), dtype=torch.int64, device=dev)
    _all_gather_rows(all_offs, offs)
    all_offs = all_offs.cpu()
    nb = len(vptr)
    vtot, ftot = all_offs[:, nb - 1], all_offs[:, 2 * nb - 1]
    vmax, fmax = int(vtot.max()), int(ftot.max())
    V, F = int(vptr[-1]), int(fptr[-1])
    vbuf = torch.empty((vmax, 6), dtype=torch.float32, device=dev)
    fbuf = torch.empty((fmax, 3), dtype=torch.int32, device=dev)
    vbuf[:V, :3] = packed["verts"].to(dev)
    vbuf[:V, 3:] = warp_field.to(dev)
    fbuf[:F] = packed["faces"].to(dev)
    all_v = torch.empty((world, vmax, 6), dtype=torch.float32, device=dev)
    all_f = torch.empty((world, fmax, 3), dtype=torch.int32, device=dev)
    _all_gather_rows(all_v, vbuf)
    _all_gather_rows(all_f, fbuf)
    return {"vptr": [all_offs[r, :nb].numpy() for r in range(world)], "fptr": [all_offs[r, nb:].numpy() for r in range(world)],
            "verts": [all_v[r, :int(vtot[r]), :3] for r in range(world)],
            "warp_field": [all_v[r, :int(vtot[r]), 3:] for r in range(world)],
            "faces": [all_f[r, :int(ftot[r])] for r in range(world)],
            "bytes": int(all_v.numel() * 4 + all_f.numel() * 4 + all_offs.numel() * 8)}
