"""Slab-decomposed 3D FFT convolution of ONE large volume over P GPUs (BASELINE config 5).

One process per GPU.  Rank g owns z planes [g*nzp, min(nz, (g+1)*nzp)) of the real volume, nzp = ceil(nz/P), and
after the exchange the ky rows [g*nyl, min(ny, (g+1)*nyl)), nyl = ceil(ny/P): the extents need not be divisible
by P (ragged slabs; exchange blocks keep the pitch nzp x nyl, the last rank's pad planes / rows are carried along
and never read back).  Per call:

  x+y forward on the z slab  (C ABI: fcb200_slab_xy_forward; the y pass writes the send buffer)
  all-to-all                 (z slabs -> ky slabs)
  fused z pass               (forward z, x PSF spectrum slab x 1/N, inverse z: fcb200_slab_z_fused)
  all-to-all                 (ky slabs -> z slabs)
  y+x inverse                (fcb200_slab_yx_inverse; the y pass reads the receive buffer)

`PeerSlabConvolver` is the B200-native form of the same schedule: no all-to-all at all.  The y pass and the
last stage of the fused z pass store every output row straight into the owning rank's buffer over
NVLink / NVSwitch peer memory (CUDA IPC mappings, fcb200_slab_xy_forward_peer / fcb200_slab_z_fused_peer), so
the transfer overlaps the butterflies tile by tile; the phases are separated by a stream-ordered barrier only.

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
    return ctypes.c_void_p(t if isinstance(t, int) else t.data_ptr())


class DistExchange:
    """all-to-all over torch.distributed (NCCL): block p of `send` goes to rank p"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group

    def __call__(self, send, recv):
        self.dist.all_to_all_single(recv, send, group=self.group)


class CPasses:
    """the pass-level C-ABI entry points (include/fcb200_ext.h: fcb200_slab_*) -- the product backend"""

    def __init__(self):
        self.lib = api._load()

    def spectrum_pitch(self, d0):
        return api.spectrum_pitch(d0)

    def xy_forward(self, real_slab, zslab, send, im_dim, nzl, nzp, nyl, dev, st):
        self.lib.fcb200_slab_xy_forward(_p(real_slab), _p(zslab), _p(send), _ints(im_dim), nzl, nzp, nyl, dev, st)

    def z_fused(self, yslab, H, im_dim, nyl, dev, st):
        self.lib.fcb200_slab_z_fused(_p(yslab), _p(H), _ints(im_dim), nyl, dev, st)

    def yx_inverse(self, recv, zslab, real_slab, im_dim, nzl, nzp, nyl, dev, st):
        self.lib.fcb200_slab_yx_inverse(_p(recv), _p(zslab), _p(real_slab), _ints(im_dim), nzl, nzp, nyl, dev, st)


class SlabConvolver:
    def __init__(self, im_dim, kernel_dim, rank, world, dev, exchange, passes=None, device=None):
        """passes / device: the tests' CPU model of the pass-level entry points (tests/test_slab_gloo.py) runs the
        same orchestration, buffers and exchange on CPU tensors; the product always uses CPasses on cuda:dev."""
        import torch
        self.torch = torch
        self.passes = passes if passes is not None else CPasses()
        self.im_dim = tuple(int(v) for v in im_dim)
        self.k_dim = tuple(int(v) for v in kernel_dim)
        d0, d1, d2 = self.im_dim
        self.rank, self.world, self.dev = rank, world, dev
        # block pitches (equal on all ranks) and what this rank really owns
        self.nzp, self.nyl = -(-d2 // world), -(-d1 // world)
        if (world - 1) * self.nzp >= d2 or (world - 1) * self.nyl >= d1:
            raise ValueError("slab mode: every rank must own at least one z plane and one ky row "
                             f"(imDim {self.im_dim} over {world} ranks)")
        self.nzl = min(self.nzp, d2 - rank * self.nzp)
        self.ny_here = min(self.nyl, d1 - rank * self.nyl)
        self.xcp = self.passes.spectrum_pitch(d0)
        self.exchange = exchange
        device = torch.device(device if device is not None else f"cuda:{dev}")
        n_spec = world * self.nzp * self.nyl * self.xcp * 2     # floats of one exchange-sized complex buffer
        self.zslab = torch.zeros(n_spec, dtype=torch.float32, device=device)
        self.buf_a = torch.zeros(n_spec, dtype=torch.float32, device=device)   # send, later receive
        self.buf_b = torch.zeros(n_spec, dtype=torch.float32, device=device)   # y-slab spectrum [P*nzp][nyl][xcp]
        self.H = torch.zeros(n_spec, dtype=torch.float32, device=device)       # PSF spectrum, y slab
        self._psf_ready_for = None
        self._lib = self.passes.lib if isinstance(self.passes, CPasses) else None

    def slab_of(self, volume_flat):
        """this rank's part of a full flat [d2][d1][d0] host/torch array (helper for tests/benchmarks)"""
        d0, d1, d2 = self.im_dim
        plane = d1 * d0
        z0 = self.rank * self.nzp
        return volume_flat[z0 * plane:(z0 + self.nzl) * plane]

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
        ps = self.passes
        st = ctypes.c_void_p(int(stream))
        ps.xy_forward(real_slab, self.zslab, self.buf_a, self.im_dim, self.nzl, self.nzp, self.nyl, self.dev, st)
        self.exchange(self.buf_a, self.buf_b)          # buf_b = [P*nzp][nyl][xcp], planes 0..d2-1 valid
        ps.z_fused(self.buf_b, self.H, self.im_dim, self.nyl, self.dev, st)
        self.exchange(self.buf_b, self.buf_a)          # buf_a = [P][nzp][nyl][xcp]
        ps.yx_inverse(self.buf_a, self.zslab, real_slab, self.im_dim, self.nzl, self.nzp, self.nyl, self.dev, st)


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
        lib.fcb200_slab_xy_forward(_p(s), _p(c.zslab), _p(c.buf_a), _ints(c.im_dim), c.nzl, c.nzp, c.nyl, c.dev, st)
    for c in convolvers:
        exchange(c.buf_a, c.buf_b)
    exchange.flush()
    for c in convolvers:
        lib.fcb200_slab_z_fused(_p(c.buf_b), _p(c.H), _ints(c.im_dim), c.nyl, c.dev, st)
    for c in convolvers:
        exchange(c.buf_b, c.buf_a)
    exchange.flush()
    for c, s in zip(convolvers, slabs):
        lib.fcb200_slab_yx_inverse(_p(c.buf_a), _p(c.zslab), _p(s), _ints(c.im_dim), c.nzl, c.nzp, c.nyl, c.dev, st)


class _RawBuffer:
    """cudaMalloc'ed device buffer (whole allocation => shareable through a CUDA IPC handle)"""

    def __init__(self, lib, nbytes, dev):
        self.lib, self.dev, self.nbytes = lib, dev, int(nbytes)
        self.ptr = int(lib.fcb200_device_malloc(self.nbytes, dev))

    def data_ptr(self):
        return self.ptr

    def ipc_handle(self):
        h = ctypes.create_string_buffer(64)
        self.lib.fcb200_ipc_get_handle(ctypes.c_void_p(self.ptr), h)
        return h.raw

    def free(self):
        if self.ptr:
            self.lib.fcb200_device_free(ctypes.c_void_p(self.ptr), self.dev)
            self.ptr = 0


class DistBarrier:
    """stream-ordered barrier: a 1-element all-reduce on the current stream (NCCL orders it after the kernels
    already queued on that stream and holds back the ones queued after it)"""

    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.flag = torch.zeros(1, dtype=torch.float32, device=device)

    def __call__(self):
        self.dist.all_reduce(self.flag, group=self.group)


class PeerSlabConvolver(SlabConvolver):
    """Slab convolution whose exchange is fused into the FFT kernels (peer stores over NVLink).

    Buffers: `yslab` [P*nzp][nyl][xcp] (written by every rank's y pass) and `recv` [P][nzp][nyl][xcp] (written by
    every rank's fused z pass) are the two buffers the peers map.  Call `connect_ipc()` (one process per GPU) or
    `connect_local()` (all ranks emulated in one process) before `convolve()`."""

    def __init__(self, im_dim, kernel_dim, rank, world, dev, barrier=None, raw=True):
        super().__init__(im_dim, kernel_dim, rank, world, dev, exchange=None)
        torch = self.torch
        nbytes = self.buf_a.numel() * 4
        if raw:     # IPC needs whole cudaMalloc allocations, not slices of torch's caching allocator
            self.buf_a = _RawBuffer(self._lib, nbytes, dev)
            self.buf_b = _RawBuffer(self._lib, nbytes, dev)
        self.barrier = barrier if barrier is not None else (lambda: None)
        self.peer_yslab = self.peer_recv = None
        self._mapped = []

    def _table(self, ptrs):
        torch = self.torch
        return torch.tensor([int(p) for p in ptrs], dtype=torch.int64, device=self.H.device)

    def connect_local(self, convolvers):
        """all ranks live in this process (single-GPU emulation, or one process driving several GPUs with
        peer access enabled)"""
        self.peer_yslab = self._table([c.buf_b.data_ptr() for c in convolvers])
        self.peer_recv = self._table([c.buf_a.data_ptr() for c in convolvers])

    def connect_ipc(self, group=None):
        """exchange CUDA IPC handles over torch.distributed and map every peer's two buffers"""
        import torch.distributed as dist
        mine = (self.buf_b.ipc_handle(), self.buf_a.ipc_handle())
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        ys, rv = [], []
        for r, (hy, hr) in enumerate(handles):
            if r == self.rank:
                ys.append(self.buf_b.data_ptr())
                rv.append(self.buf_a.data_ptr())
                continue
            py = int(self._lib.fcb200_ipc_open_handle(hy, self.dev))
            pr = int(self._lib.fcb200_ipc_open_handle(hr, self.dev))
            self._mapped += [py, pr]
            ys.append(py)
            rv.append(pr)
        self.peer_yslab, self.peer_recv = self._table(ys), self._table(rv)

    def close(self):
        for p in self._mapped:
            self._lib.fcb200_ipc_close_handle(ctypes.c_void_p(p), self.dev)
        self._mapped = []
        for b in (self.buf_a, self.buf_b):
            if isinstance(b, _RawBuffer):
                b.free()

    # the three phases, separately callable so that emulated ranks can run them in lock-step
    def phase_forward(self, real_slab, st):
        self._lib.fcb200_slab_xy_forward_peer(_p(real_slab), _p(self.zslab), _p(self.peer_yslab), _ints(self.im_dim),
                                              self.nzl, self.nzp, self.nyl, self.rank, self.dev, st)

    def phase_z(self, st):
        self._lib.fcb200_slab_z_fused_peer(_p(self.buf_b), _p(self.H), _p(self.peer_recv), _ints(self.im_dim), self.nzp,
                                           self.nyl, self.rank, self.dev, st)

    def phase_inverse(self, real_slab, st):
        self._lib.fcb200_slab_yx_inverse(_p(self.buf_a), _p(self.zslab), _p(real_slab), _ints(self.im_dim), self.nzl,
                                         self.nzp, self.nyl, self.dev, st)

    def convolve(self, real_slab, stream=0):
        st = ctypes.c_void_p(int(stream))
        self.barrier()                 # every rank has finished reading the buffers of the previous call
        self.phase_forward(real_slab, st)
        self.barrier()                 # all y-pass rows have landed in my y slab
        self.phase_z(st)
        self.barrier()                 # all planes have landed in my receive buffer
        self.phase_inverse(real_slab, st)


def run_lockstep_peer(convolvers, slabs):
    """Drive `world` emulated PeerSlabConvolvers through one convolution on a single GPU (same stream: the
    phases are ordered by the stream itself)."""
    st = ctypes.c_void_p(0)
    for c, s in zip(convolvers, slabs):
        c.phase_forward(s, st)
    for c in convolvers:
        c.phase_z(st)
    for c, s in zip(convolvers, slabs):
        c.phase_inverse(s, st)
