"""Multi-GPU plumbing: the forward has no exchange step, so images shard over ranks and the only
collective is one all-gather of the packed per-image records (58.6 KB/img). One process per GPU
(torchrun / torch.distributed for rendezvous only); the all-gather itself is `dirb200_allgather_records`
(NCCL over NVLink) on the compute stream. The reference has no distributed code to mirror."""
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous balanced split of n images: the first n % world ranks take one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def padded_shard(n: int, world: int) -> int:
    """ncclAllGather needs equal contributions: every rank sends ceil(n / world) records (tail zero-padded)."""
    return (n + world - 1) // world


def broadcast_bytes(payload, src: int = 0, group=None) -> bytes:
    """Rank `src` passes bytes, the others None; everybody gets src's bytes (used for the 128-byte NCCL id)."""
    obj = [payload]
    dist.broadcast_object_list(obj, src=src, group=group)
    return obj[0]


def assemble(gathered: torch.Tensor, n: int, world: int) -> torch.Tensor:
    """(world * padded, R) rank-major gather result -> (n, R) in original image order (pads dropped)."""
    pad = padded_shard(n, world)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        parts.append(gathered[r * pad:r * pad + (hi - lo)])
    return torch.cat(parts, 0)


def init_nccl(model):
    """Create the library's communicator for an initialised torch.distributed job."""
    model.init_nccl(dist.get_rank(), dist.get_world_size(), lambda b: broadcast_bytes(b))


def forward_sharded(model, img_global: torch.Tensor):
    """Every rank holds the same (n,3,256,256) batch (or at least its own slice of it); each runs its shard and all
    ranks return the reference's outs_list for all n images (aux maps stay local and are not gathered)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = img_global.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    pad = padded_shard(n, world)
    if hi > lo:
        local = model.run_raw(img_global[lo:hi])["record"]
    else:  # fewer images than ranks: this rank has no work but must still enter the all-gather
        from . import capi
        local = torch.zeros(0, capi.RECORD_FLOATS, device=model._device())
    if local.shape[0] < pad:
        local = torch.cat([local, local.new_zeros(pad - local.shape[0], local.shape[1])], 0)
    full = assemble(model.allgather_records(local.contiguous()), n, world)
    return model.unpack_record(full)


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and, by first touch, every pinned host buffer it allocates afterwards) to the CPU cores of the
    NUMA node its GPU hangs off. With one process per GPU and all ranks left on node 0, eight concurrent H2D streams
    share one socket's memory controllers and cross the socket interconnect (measured round 1: 0.74 end-to-end scaling
    at 8 GPUs with the kernels themselves at 0.97). Returns a small report dict, or None when sysfs has no answer."""
    import os

    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(f"{base}/local_cpulist") as f:
            cpulist = f.read().strip()
        node = None
        if os.path.exists(f"{base}/numa_node"):
            with open(f"{base}/numa_node") as f:
                node = int(f.read().strip())
        cpus = set()
        for part in cpulist.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {"gpu": bdf, "numa_node": node, "bound": False, "why": "no local cpu is in this process' cpuset"}
        # several ranks share these cores (always the case on a single-node VM): give each rank its own slice, so that
        # eight Python main threads, their NCCL proxy threads and the copy-engine interrupt handlers do not migrate
        # over each other
        nlocal = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        ordered = sorted(cpus)
        share = len(ordered) // max(nlocal, 1)
        if nlocal > 1 and share >= 2:
            mine = set(ordered[local_rank * share:(local_rank + 1) * share])
        else:
            mine = cpus
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(len(mine), 8)))
        return {"gpu": bdf, "numa_node": node, "bound": True, "cpus": len(mine), "local_cpus": len(cpus),
                "of_allowed": len(allowed)}
    except (OSError, ValueError, AttributeError, RuntimeError) as e:
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}
