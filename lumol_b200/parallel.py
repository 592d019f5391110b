"""Multi-GPU plumbing on the host side: one process per GPU, NCCL inside the library.

``torch.distributed`` is only the bootstrap (it carries the 128-byte NCCL unique id from rank 0 to the other
ranks and provides the barrier / max-over-ranks for timing); every data-path collective (position all-gather,
rho(k) / energy / virial all-reduce) is issued by ``liblumol_cuda.so`` on its own stream.
"""

import ctypes

from . import _ffi


def owned_range(n, rank, world):
    """Atoms [lo, hi) owned by ``rank``: equal contiguous blocks of ceil(n / world) atoms (the last may be short).
    Mirrors ``Context::owned_range`` in csrc/context.hpp, which the all-gather block size relies on."""
    chunk = (n + world - 1) // world
    lo = min(n, chunk * rank)
    hi = min(n, lo + chunk)
    return lo, hi


def exchange_unique_id(rank, make_id, broadcast):
    """Rank 0 creates the id with ``make_id()`` (128 bytes); ``broadcast(bytes_or_None)`` returns rank 0's bytes on
    every rank.  Split out so the CPU tests can drive it with a gloo group."""
    payload = make_id() if rank == 0 else None
    data = broadcast(payload)
    if len(data) != 128:
        raise ValueError("a NCCL unique id is 128 bytes")
    return bytes(data)


def init_communicator(device, rank, world):
    """Create the library's NCCL communicator for ``device`` (a ``DeviceSystem``) using torch.distributed as the
    out-of-band channel."""
    import torch
    import torch.distributed as dist

    lib, ctx = device.lib, device.ctx

    def make_id():
        buffer = (ctypes.c_uint8 * 128)()
        _ffi.check(None, lib.lumol_cuda_comm_unique_id(buffer))
        return bytes(buffer)

    def broadcast(payload):
        tensor = torch.zeros(128, dtype=torch.uint8)
        if payload is not None:
            tensor = torch.tensor(list(payload), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            tensor = tensor.cuda()
        dist.broadcast(tensor, 0)
        return bytes(tensor.cpu().tolist())

    unique_id = exchange_unique_id(rank, make_id, broadcast)
    buffer = (ctypes.c_uint8 * 128)(*unique_id)
    _ffi.check(ctx, lib.lumol_cuda_comm_init(ctx, world, rank, buffer))


def max_over_ranks(value, world):
    """Largest ``value`` over the ranks (device-side timings are reported as the max)."""
    if world == 1:
        return value
    import torch
    import torch.distributed as dist

    tensor = torch.tensor([value], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        tensor = tensor.cuda()
    dist.all_reduce(tensor, op=dist.ReduceOp.MAX)
    return float(tensor.item())
