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
