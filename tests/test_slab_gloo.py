"""world_size-2 test (gloo, CPU) of the slab-decomposed path's HOST logic: slab ownership, buffer contracts and
the all-to-all exchange of fourierconvolutioncudalib_b200/slab.py (SlabConvolver + DistExchange).  The pass-level
kernels are replaced by a float64 torch.fft model that honours the documented buffer layouts
(include/fcb200_ext.h: z-slab spectrum [nzl][d1][xcp], exchange buffer [P][nzl][nyl][xcp], y-slab spectrum
[d2][nyl][xcp]); the result must equal the oracle's convolution of the whole volume (oracle/fc_oracle.py)."""
import os
import socket
import sys

import numpy as np
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


class ModelPasses:
    """CPU model of fcb200_slab_xy_forward / _z_fused / _yx_inverse on float32 buffers holding interleaved
    complex values (the on-device pair-planar order is internal to the kernels and irrelevant to the exchange)."""

    def __init__(self, world):
        self.world = world

    def spectrum_pitch(self, d0):
        return ((d0 // 2 + 1) + 3) & ~3

    def _c(self, buf, shape):
        return torch.view_as_complex(buf.view(*shape, 2))

    def xy_forward(self, real_slab, zslab, send, im_dim, nzl, nzp, nyl, dev, st):
        d0, d1, d2 = im_dim
        xc, xcp, P = d0 // 2 + 1, self.spectrum_pitch(d0), -(-d1 // nyl)
        spec = torch.fft.fft(torch.fft.rfft(real_slab.view(nzl, d1, d0).double(), dim=2), dim=1)   # [nzl][d1][xc]
        out = self._c(send, (P, nzp, nyl, xcp))
        out.fill_(float("nan"))          # pad planes / rows of ragged blocks must never reach the result
        for p in range(P):
            rows = min(nyl, d1 - p * nyl)
            out[p, :nzl, :rows, :xc] = spec[:, p * nyl:p * nyl + rows, :].to(torch.complex64)

    def z_fused(self, yslab, H, im_dim, nyl, dev, st):
        d0, d1, d2 = im_dim
        xcp = self.spectrum_pitch(d0)
        y = self._c(yslab[:d2 * nyl * xcp * 2], (d2, nyl, xcp))     # the buffer holds P*nzp >= d2 planes
        h = self._c(H[:d2 * nyl * xcp * 2], (d2, nyl, xcp))
        z = torch.fft.fft(y.to(torch.complex128), dim=0) * h.to(torch.complex128) / float(d0 * d1 * d2)
        y.copy_((torch.fft.ifft(z, dim=0) * d2).to(torch.complex64))       # unnormalised inverse, like the kernels

    def yx_inverse(self, recv, zslab, real_slab, im_dim, nzl, nzp, nyl, dev, st):
        d0, d1, d2 = im_dim
        xc, xcp, P = d0 // 2 + 1, self.spectrum_pitch(d0), -(-d1 // nyl)
        blocks = self._c(recv, (P, nzp, nyl, xcp)).to(torch.complex128)
        spec = torch.cat([blocks[p, :nzl] for p in range(P)], dim=1)[:, :d1, :xc]                   # [nzl][d1][xc]
        out = torch.fft.irfft(torch.fft.ifft(spec, dim=1) * d1, n=d0, dim=2) * d0                   # unnormalised
        real_slab.copy_(out.reshape(-1).float())


def _worker(rank, world, port, im_dim, k_dim, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import fc_oracle
    from fourierconvolutioncudalib_b200 import slab
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d0, d1, d2 = im_dim
    rng = np.random.default_rng(77)                      # same volume on every rank
    im = rng.random(d0 * d1 * d2, dtype=np.float32)
    k = rng.random(int(np.prod(k_dim)), dtype=np.float32)
    passes = ModelPasses(world)
    conv = slab.SlabConvolver(im_dim, k_dim, rank, world, 0, slab.DistExchange(), passes=passes, device="cpu")
    # PSF spectrum of this rank's ky slab, from the oracle's placement
    S = fc_oracle.place_psf(k, k_dim, im_dim).reshape(d2, d1, d0)
    Hfull = torch.fft.fft(torch.fft.fft(torch.fft.rfft(torch.from_numpy(S), dim=2), dim=1), dim=0)
    xc = d0 // 2 + 1
    h = torch.view_as_complex(conv.H[:d2 * conv.nyl * conv.xcp * 2].view(d2, conv.nyl, conv.xcp, 2))
    h.zero_()
    h[:, :conv.ny_here, :xc] = Hfull[:, rank * conv.nyl:rank * conv.nyl + conv.ny_here, :].to(torch.complex64)
    mine = torch.from_numpy(im.copy())
    my_slab = conv.slab_of(mine).clone()
    assert my_slab.numel() == conv.nzl * d1 * d0
    conv.convolve(my_slab)
    padded = torch.zeros(conv.nzp * d1 * d0)            # all_gather needs equal sizes: ragged slabs are padded
    padded[:my_slab.numel()] = my_slab
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    if rank == 0:
        got = torch.cat(parts).numpy()[:d0 * d1 * d2]
        want = fc_oracle.convolve_inplace_ref(im, im_dim, k, k_dim)
        out.put((float(np.abs(got - want).max() / np.abs(want).max()),
                 float(np.linalg.norm(got - want) / np.linalg.norm(want))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("im_dim,k_dim", [((16, 12, 8), (3, 5, 3)), ((10, 8, 6), (3, 3, 3)),
                                          ((12, 9, 7), (3, 3, 3)), ((8, 5, 6), (3, 3, 1))])     # ragged: 9, 7, 5 over 2 ranks
def test_two_rank_slab_schedule_matches_oracle(im_dim, k_dim):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, im_dim, k_dim, out)) for r in range(world)]
    for p in procs:
        p.start()
    max_err, l2 = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert max_err <= 1e-4 and l2 <= 1e-5


def test_slab_ownership_ragged_and_rejected_extents():
    sys.path.insert(0, ROOT)
    from fourierconvolutioncudalib_b200 import slab
    # 1125 planes / 2160 rows over 8 ranks (BASELINE config 5, caller-padded): pitch 141 / 270, last rank 138 planes
    convs = [slab.SlabConvolver((16, 2160, 1125), (3, 3, 3), r, 8, 0, None, passes=ModelPasses(8), device="meta")
             for r in range(8)]
    assert [c.nzp for c in convs] == [141] * 8 and [c.nzl for c in convs] == [141] * 7 + [138]
    assert [c.nyl for c in convs] == [270] * 8 and sum(c.ny_here for c in convs) == 2160
    assert sum(c.nzl for c in convs) == 1125
    with pytest.raises(ValueError):      # 3 rows over 4 ranks: the last rank would own nothing
        slab.SlabConvolver((16, 3, 8), (3, 3, 3), 0, 4, 0, None, passes=ModelPasses(4), device="cpu")
    with pytest.raises(ValueError):      # 9 planes over 8 ranks: pitch 2 -> ranks 5..7 would own nothing
        slab.SlabConvolver((16, 16, 9), (3, 3, 3), 0, 8, 0, None, passes=ModelPasses(8), device="cpu")
