"""Sharding of independent convolution tiles (multi-view deconvolution blocks, BASELINE config 4) over
one-process-per-GPU ranks.  There is no collective on the data path: every rank convolves its own
tiles through the C ABI; torch.distributed is used only for the barrier and for reducing timings
(max over ranks).  Pure host logic -- testable with the gloo backend on CPUs."""
import os


def assign_tiles(n_tiles, world_size, rank):
    """Round-robin assignment (tile b -> rank b mod P): balanced to within one tile for any P."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return list(range(rank, n_tiles, world_size))


def rank_info():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def parse_cpulist(text):
    """'0-3,8-11' (sysfs cpulist format) -> [0, 1, 2, 3, 8, 9, 10, 11]"""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(pci_bus_id, sysfs="/sys"):
    """NUMA node of the GPU with PCI address 'dddd:bb:dd.f' (sysfs), or None when the platform does not say"""
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", pci_bus_id.lower(), "numa_node")).read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa_node(local_rank, sysfs="/sys"):
    """One process per GPU: run this rank's host threads (and therefore first-touch its pinned staging buffers) on the
    CPUs of the NUMA node its GPU hangs off, so that the 2 x volume bytes of every host-pointer call do not cross the
    socket interconnect and the ranks spread over all memory controllers.  Returns (node, n_cpus) or None when
    nothing was changed (no sysfs information, single node, affinity not permitted)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus_id = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None
    node = gpu_numa_node(bus_id, sysfs)
    if node is None:
        return None
    try:
        cpus = parse_cpulist(open(os.path.join(sysfs, f"devices/system/node/node{node}/cpulist")).read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
    except (OSError, ValueError, AttributeError):
        return None
    return node, len(cpus)


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def convolve_tiles(tiles, im_dim, kernel, kernel_dim, dev, stream=0):
    """Convolve this rank's device-resident tiles in place, stream-ordered, no host sync.
    `tiles` is a list of torch CUDA tensors (one flat fp32 tensor per tile)."""
    from . import api
    for t in tiles:
        api.convolve_device_async(t, im_dim, kernel, kernel_dim, dev, stream)


def convolve_host_tiles(tiles, im_dim, kernel, kernel_dim, dev):
    """Convolve this rank's HOST tiles (numpy arrays or pinned torch tensors) in place through the pipelined
    batch entry point: one PSF spectrum, upload / convolution / download of neighbouring tiles overlap
    (BASELINE config 4: 64 blocks of 384^3 dealt round-robin over the ranks, see assign_tiles)."""
    from . import api
    if tiles:
        api.convolve_batch(tiles, im_dim, kernel, kernel_dim, dev)
