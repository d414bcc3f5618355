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

    def xy_forward(self, real_slab, zslab, send, im_dim, nzl, nyl, dev, st):
        d0, d1, d2 = im_dim
        xc, xcp, P = d0 // 2 + 1, self.spectrum_pitch(d0), d1 // nyl
        spec = torch.fft.fft(torch.fft.rfft(real_slab.view(nzl, d1, d0).double(), dim=2), dim=1)   # [nzl][d1][xc]
        out = self._c(send, (P, nzl, nyl, xcp))
        out.zero_()
        for p in range(P):
            out[p, :, :, :xc] = spec[:, p * nyl:(p + 1) * nyl, :].to(torch.complex64)

    def z_fused(self, yslab, H, im_dim, nyl, dev, st):
        d0, d1, d2 = im_dim
        xcp = self.spectrum_pitch(d0)
        y = self._c(yslab, (d2, nyl, xcp))
        h = self._c(H, (d2, nyl, xcp))
        z = torch.fft.fft(y.to(torch.complex128), dim=0) * h.to(torch.complex128) / float(d0 * d1 * d2)
        y.copy_((torch.fft.ifft(z, dim=0) * d2).to(torch.complex64))       # unnormalised inverse, like the kernels

    def yx_inverse(self, recv, zslab, real_slab, im_dim, nzl, nyl, dev, st):
        d0, d1, d2 = im_dim
        xc, xcp, P = d0 // 2 + 1, self.spectrum_pitch(d0), d1 // nyl
        blocks = self._c(recv, (P, nzl, nyl, xcp)).to(torch.complex128)
        spec = torch.cat([blocks[p] for p in range(P)], dim=1)[:, :, :xc]                           # [nzl][d1][xc]
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
    h = torch.view_as_complex(conv.H.view(d2, conv.nyl, conv.xcp, 2))
    h.zero_()
    h[:, :, :xc] = Hfull[:, rank * conv.nyl:(rank + 1) * conv.nyl, :].to(torch.complex64)
    mine = torch.from_numpy(im.copy())
    my_slab = conv.slab_of(mine).clone()
    conv.convolve(my_slab)
    parts = [torch.empty_like(my_slab) for _ in range(world)]
    dist.all_gather(parts, my_slab)
    if rank == 0:
        got = torch.cat(parts).numpy()
        want = fc_oracle.convolve_inplace_ref(im, im_dim, k, k_dim)
        out.put((float(np.abs(got - want).max() / np.abs(want).max()),
                 float(np.linalg.norm(got - want) / np.linalg.norm(want))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("im_dim,k_dim", [((16, 12, 8), (3, 5, 3)), ((10, 8, 6), (3, 3, 3))])
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


def test_slab_rejects_indivisible_extents():
    sys.path.insert(0, ROOT)
    from fourierconvolutioncudalib_b200 import slab
    with pytest.raises(ValueError):
        slab.SlabConvolver((16, 9, 8), (3, 3, 3), 0, 2, 0, None, passes=ModelPasses(2), device="cpu")
