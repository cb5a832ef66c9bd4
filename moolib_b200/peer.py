"""One-process-per-GPU plumbing for the HP-A kernels: exchange the opaque mb_ar_handle blobs over an existing
torch.distributed process group (gloo or nccl) and import every peer.  torch.distributed is only the control plane
here -- the gradient bytes never go through it (the reference's control plane is its Group/Broker RPC,
src/group.h:330-491; in the moolib-API host layer the same blobs travel over that instead).
"""
import torch
import torch.distributed as dist

from . import _lib


def gather_blobs(blob: bytes, group=None, device=None):
    """all_gather of a fixed-size byte string; returns the list indexed by rank."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cpu") if backend == "gloo" else torch.device(device or f"cuda:{torch.cuda.current_device()}")
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def connect_context(ctx, group=None, device=None):
    """Export this rank's handle, gather everyone's, import all peers.  Collective over `group`."""
    blobs = gather_blobs(ctx.export(), group, device)
    rank = dist.get_rank(group)
    for r, b in enumerate(blobs):
        if r != rank:
            ctx.import_peer(r, b)
    dist.barrier(group)
    return ctx


def make_context(max_bytes, nslots=1, group=None, device=None):
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    ctx = _lib.ArContext(rank, world, dev, max_bytes, nslots)
    return connect_context(ctx, group, dev)


def shard_range(n_items, rank, world):
    """Contiguous block partition of n_items work units over `world` ranks (weak scaling keeps per-rank work fixed;
    this is for strong-scaled sweeps)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
