"""world_size-2 test of the multi-GPU host logic on CPUs (gloo): tile sharding and the
max-over-ranks timing reduction used by bench.py.  No collective touches tile data."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_tiles, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fourierconvolutioncudalib_b200 import tiles
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = tiles.assign_tiles(n_tiles, world, rank)
    # every rank "processes" its tiles; pretend rank r needs (r+1) ms per tile
    elapsed = (rank + 1) * 1.0 * len(mine)
    worst = tiles.max_over_ranks(elapsed)
    total = tiles.sum_over_ranks(len(mine))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((gathered, worst, total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_tiles", [64, 7])
def test_tiles_shard_exactly_once_and_timing_is_max_over_ranks(n_tiles):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_tiles, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, worst, total = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(t for g in gathered for t in g)
    assert flat == list(range(n_tiles))                       # every tile exactly once
    assert max(len(g) for g in gathered) - min(len(g) for g in gathered) <= 1
    assert total == n_tiles
    assert worst == max((r + 1) * len(g) for r, g in enumerate(gathered))


def test_assign_tiles_single_rank_and_errors():
    sys.path.insert(0, ROOT)
    from fourierconvolutioncudalib_b200 import tiles
    assert tiles.assign_tiles(5, 1, 0) == [0, 1, 2, 3, 4]
    assert tiles.assign_tiles(3, 8, 5) == []
    with pytest.raises(ValueError):
        tiles.assign_tiles(3, 2, 2)


def test_numa_helpers(tmp_path):
    """host-side pieces of tiles.bind_to_gpu_numa_node (sysfs parsing; no GPU, no affinity change)"""
    sys.path.insert(0, ROOT)
    from fourierconvolutioncudalib_b200 import tiles
    assert tiles.parse_cpulist("0-3,8-11\n") == [0, 1, 2, 3, 8, 9, 10, 11]
    assert tiles.parse_cpulist("5") == [5] and tiles.parse_cpulist("") == []
    dev = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    assert tiles.gpu_numa_node("0000:1B:00.0", str(tmp_path)) == 1
    (dev / "numa_node").write_text("-1\n")
    assert tiles.gpu_numa_node("0000:1b:00.0", str(tmp_path)) is None      # platform does not say
    assert tiles.gpu_numa_node("0000:99:00.0", str(tmp_path)) is None
    assert tiles.bind_to_gpu_numa_node(0, str(tmp_path)) is None           # no CUDA device here: nothing changes
