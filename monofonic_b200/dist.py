"""Slab decomposition over the GPUs of one node: one process per GPU, launched by torchrun.

The reference splits every Grid_FFT into FFTW-MPI slabs (src/grid_fft.cc:120-128) and lets FFTW-MPI do the
all-to-all inside each 3-D FFT.  Here the library does the exchange itself: the FFT pass that precedes it stores
straight into the peer GPUs' receive buffers (CUDA IPC mapping, NVLink/NVSwitch), closed by a device-side flag
barrier.  The host side only has to (1) pick the device, (2) carry the 64-byte IPC handles between the ranks
once - done here with torch.distributed - and (3) reduce timings.  After `init()` every Grid_FFT /
OrszagConvolver / Particles / LPTPipeline created in this process is a slab; the pipeline code is unchanged.

`staged=True` runs the same slab algorithm with the exchange delivered by `torch.distributed.all_to_all_single`
(gloo on CPU for the emulation tests; NCCL on GPUs as the library-collective baseline the fused path is
compared with)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import grid as G

_keep = {}


class _DevBuf:
    """raw device pointer -> torch tensor (via __cuda_array_interface__)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def _make_exchange_cb(lib, world, device):
    import torch
    import torch.distributed as dist

    on_gpu = lib.mfb_build_info() == b"cuda sm_100a"

    def view(ptr, nbytes):
        if on_gpu:
            return torch.as_tensor(_DevBuf(ptr, nbytes), device=f"cuda:{device}")
        return torch.from_numpy(np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(ptr)))

    def cb(send, recv, chunk_bytes, user):
        n = int(chunk_bytes) * world
        s, r = view(send, n), view(recv, n)
        try:
            dist.all_to_all_single(r, s)
        except RuntimeError:
            # backends without all-to-all (older gloo): pairwise exchange
            reqs = []
            me = dist.get_rank()
            r[me * chunk_bytes:(me + 1) * chunk_bytes] = s[me * chunk_bytes:(me + 1) * chunk_bytes]
            for p in range(world):
                if p != me:
                    reqs.append(dist.isend(s[p * chunk_bytes:(p + 1) * chunk_bytes], p))
                    reqs.append(dist.irecv(r[p * chunk_bytes:(p + 1) * chunk_bytes], p))
            for q in reqs:
                q.wait()
        if on_gpu:
            torch.cuda.synchronize()

    return L.EXCHANGE_FN(cb)


def pick_nslots(nbytes, n, world, free_bytes, lpt_order=3, particle_bytes_per_cell=32):
    """how many exchange buffers to put into the peer-mapped block.  The 3-D FFTs need 2 (a ping-pong pair); the shared-transform
    source stage uses ALL of them as slots and wants 10 at 3LPT (the six X buffers of the last stage + its four targets in ONE
    shared-field loop; 8: that stage as two multi-product loops; 7: one target per loop; 6, 5: targets split over several
    loops) and 4 at 2LPT.  The largest count is taken that leaves room for everything else of an LPT run - 8 grids, the
    particle arrays (which also serve as the chunk workspace of that stage) and a safety margin; below the minimum the
    block keeps 2 buffers and the memory goes to the per-product workspace."""
    if lpt_order < 2:
        return 2
    cells = n[0] * n[1] * n[2] // world
    fixed = 8 * 8 * cells * (n[2] + 2) // n[2] + particle_bytes_per_cell * cells + (3 << 30)
    for k in ((10, 8, 7, 6, 5) if lpt_order > 2 else (4,)):
        if k * nbytes + fixed <= free_bytes:
            return k
    return 2


def init(device: int, grid_res, with_conv=True, staged=False, lib_path=None, nslots=None, lpt_order=3):
    """mfb_init + slab setup for the current torch.distributed world (no-op split for world size 1).
    grid_res: the largest grid (int or 3-tuple) that will be created; sizes the exchange buffers.
    nslots: exchange buffers in the peer-mapped block (None: pick_nslots from the free device memory)."""
    import torch.distributed as dist

    lib = G.init(device, lib_path)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return lib
    n = (grid_res,) * 3 if isinstance(grid_res, int) else tuple(grid_res)
    nbytes = lib.mfb_dist_xbuf_bytes((C.c_int * 3)(*n), int(with_conv), world)
    if nslots is None:
        nslots = 2
        if with_conv:
            free = 64 << 30
            if lib.mfb_build_info() == b"cuda sm_100a":
                import torch

                free = torch.cuda.mem_get_info(device)[0]
            if staged:
                free //= 2      # staged mode keeps a staging copy of every buffer
            nslots = pick_nslots(nbytes, n, world, free, lpt_order)
            # every rank must carve the same block
            t = [None] * world
            dist.all_gather_object(t, int(nslots))
            nslots = min(t)
    if staged:
        cb = _make_exchange_cb(lib, world, device)
        _keep["cb"] = cb  # the library holds the raw pointer: keep the thunk alive
        L.check(lib.mfb_dist_init_staged_slots(rank, world, nbytes, int(nslots), cb, None), lib)
    else:
        h = C.create_string_buffer(64)
        L.check(lib.mfb_dist_init_slots(rank, world, nbytes, int(nslots), h), lib)
        handles = [None] * world
        dist.all_gather_object(handles, h.raw)
        blob = C.create_string_buffer(b"".join(handles), 64 * world)
        L.check(lib.mfb_dist_connect(blob), lib)
        dist.barrier()
    return lib


def shutdown():
    lib = G._state["lib"]
    if lib is not None:
        lib.mfb_dist_shutdown()
    _keep.clear()


def check():
    lib = G._lib()
    L.check(lib.mfb_dist_check(), lib)


def gather_real(g: "G.Grid_FFT") -> np.ndarray:
    """all ranks: the global real-space array [n0][n1][n2] of a slab grid (tests / small outputs)"""
    import torch
    import torch.distributed as dist

    loc = torch.from_numpy(g.get_real())
    if not g.is_distributed():
        return loc.numpy()
    parts = [torch.empty_like(loc) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, loc)
    return torch.cat(parts, 0).numpy()


def gather_kspace(g: "G.Grid_FFT") -> np.ndarray:
    """all ranks: the global half-spectrum in the single-GPU layout [kx][ky][kz]"""
    import torch
    import torch.distributed as dist

    loc = g.get_kspace()
    if not g.is_distributed():
        return loc
    t = torch.from_numpy(np.ascontiguousarray(loc).view(np.float64))
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    full = torch.cat(parts, 0).numpy().view(np.complex128)  # [ky][kx][kz]
    return np.ascontiguousarray(np.transpose(full, (1, 0, 2)))


def gather_particles(p: dict) -> dict:
    """concatenate per-rank particle arrays in the single-GPU ordering: for every lattice shift the slabs
    in rank order (particle_generator.hh:150-169 stores shift iload at iload*N_loc + cell)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    out = {}
    for k, v in p.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        out[k] = [x.numpy() for x in parts]
    return out
