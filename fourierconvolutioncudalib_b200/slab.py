"""Slab-decomposed 3D FFT convolution of ONE large volume over P GPUs (BASELINE config 5).

One process per GPU.  Rank g owns z planes [g*nz/P, (g+1)*nz/P) of the real volume.  Per call:

  x+y forward on the z slab  (C ABI: fcb200_slab_xy_forward; the y pass writes the send buffer)
  all-to-all                 (z slabs -> ky slabs)
  fused z pass               (forward z, x PSF spectrum slab x 1/N, inverse z: fcb200_slab_z_fused)
  all-to-all                 (ky slabs -> z slabs)
  y+x inverse                (fcb200_slab_yx_inverse; the y pass reads the receive buffer)

The exchange is the only collective; it is pluggable so the same code runs with torch.distributed
(NCCL, `DistExchange`) or with emulated ranks inside one process (`LocalExchange`, used by the
single-GPU tests).  Result == the single-GPU convolution3DfftCUDAInPlace result up to fp32 round-off
(the arithmetic per pencil is identical; only the kernel variant of the y pass differs).
"""
import ctypes

import numpy as np

from . import api


def _ints(seq):
    return (ctypes.c_int * len(seq))(*[int(v) for v in seq])


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class DistExchange:
    """all-to-all over torch.distributed (NCCL): block p of `send` goes to rank p"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group

    def __call__(self, send, recv):
        self.dist.all_to_all_single(recv, send, group=self.group)


class SlabConvolver:
    def __init__(self, im_dim, kernel_dim, rank, world, dev, exchange):
        import torch
        self.torch = torch
        self.im_dim = tuple(int(v) for v in im_dim)
        self.k_dim = tuple(int(v) for v in kernel_dim)
        d0, d1, d2 = self.im_dim
        if d2 % world or d1 % world:
            raise ValueError("slab mode needs imDim[1] and imDim[2] divisible by the number of ranks")
        self.rank, self.world, self.dev = rank, world, dev
        self.nzl, self.nyl = d2 // world, d1 // world
        self.xcp = api.spectrum_pitch(d0)
        self.exchange = exchange
        device = torch.device(f"cuda:{dev}")
        n_spec = self.nzl * d1 * self.xcp * 2            # floats of one slab-sized complex buffer
        self.zslab = torch.empty(n_spec, dtype=torch.float32, device=device)
        self.buf_a = torch.empty(n_spec, dtype=torch.float32, device=device)   # send, later receive
        self.buf_b = torch.empty(n_spec, dtype=torch.float32, device=device)   # y-slab spectrum
        self.H = torch.empty(n_spec, dtype=torch.float32, device=device)       # PSF spectrum, y slab
        self._psf_ready_for = None
        self._lib = api._load()

    def slab_of(self, volume_flat):
        """this rank's part of a full flat [d2][d1][d0] host/torch array (helper for tests/benchmarks)"""
        d0, d1, d2 = self.im_dim
        n = self.nzl * d1 * d0
        return volume_flat[self.rank * n:(self.rank + 1) * n]

    def prepare_psf(self, kernel_dev, stream=0):
        """PSF spectrum of this rank's ky slab (call again when the PSF changes)"""
        torch = self.torch
        lib = self._lib
        n = lib.fcb200_slab_psf_scratch_elems(_ints(self.im_dim), _ints(self.k_dim), self.dev)
        scratch = torch.empty(max(int(n), 1) * 2, dtype=torch.float32, device=self.H.device)
        self.H.zero_()     # planes without taps are never read by the z pass, but keep the buffer defined
        lib.fcb200_slab_psf(_p(kernel_dev), _ints(self.k_dim), _ints(self.im_dim), self.rank * self.nyl, self.nyl,
                            _p(self.H), _p(scratch), self.dev, ctypes.c_void_p(int(stream)))
        torch.cuda.current_stream().synchronize() if stream == 0 else None
        self._scratch = scratch      # keep alive until the stream has consumed it
        self._psf_ready_for = kernel_dev.data_ptr()

    def convolve(self, real_slab, stream=0):
        """in-place convolution of this rank's z slab (flat torch CUDA tensor [nzl][d1][d0])"""
        lib = self._lib
        st = ctypes.c_void_p(int(stream))
        dims = _ints(self.im_dim)
        lib.fcb200_slab_xy_forward(_p(real_slab), _p(self.zslab), _p(self.buf_a), dims, self.nzl, self.nyl, self.dev, st)
        self.exchange(self.buf_a, self.buf_b)          # buf_b = [d2][nyl][xcp]
        lib.fcb200_slab_z_fused(_p(self.buf_b), _p(self.H), dims, self.nyl, self.dev, st)
        self.exchange(self.buf_b, self.buf_a)          # buf_a = [P][nzl][nyl][xcp]
        lib.fcb200_slab_yx_inverse(_p(self.buf_a), _p(self.zslab), _p(real_slab), dims, self.nzl, self.nyl, self.dev, st)


class LocalExchange:
    """Emulates the all-to-all between `world` SlabConvolvers that live in ONE process (tests on one
    GPU).  Ranks run their steps in lock-step through run_lockstep()."""

    def __init__(self, world):
        self.world = world
        self.pending = []

    def __call__(self, send, recv):
        self.pending.append((send, recv))

    def flush(self):
        assert len(self.pending) == self.world
        P = self.world
        n = self.pending[0][0].numel() // P
        for dst in range(P):
            for src in range(P):
                self.pending[dst][1][src * n:(src + 1) * n].copy_(self.pending[src][0][dst * n:(dst + 1) * n])
        self.pending = []


def run_lockstep(convolvers, slabs, exchange):
    """Drive `world` emulated ranks through one convolution on a single GPU."""
    import ctypes as ct
    st = ct.c_void_p(0)
    lib = convolvers[0]._lib
    for c, s in zip(convolvers, slabs):
        lib.fcb200_slab_xy_forward(_p(s), _p(c.zslab), _p(c.buf_a), _ints(c.im_dim), c.nzl, c.nyl, c.dev, st)
    for c in convolvers:
        exchange(c.buf_a, c.buf_b)
    exchange.flush()
    for c in convolvers:
        lib.fcb200_slab_z_fused(_p(c.buf_b), _p(c.H), _ints(c.im_dim), c.nyl, c.dev, st)
    for c in convolvers:
        exchange(c.buf_b, c.buf_a)
    exchange.flush()
    for c, s in zip(convolvers, slabs):
        lib.fcb200_slab_yx_inverse(_p(c.buf_a), _p(c.zslab), _p(s), _ints(c.im_dim), c.nzl, c.nyl, c.dev, st)
