"""Multi-GPU sharding of a batch of independent worlds.

Worlds are independent pure function calls (reference: src/Physics.elm:522-525), so a batch shards by contiguous
world-index range, one process per GPU, with NO data-path collective while stepping; the only collective is the
final gather of the integrated state (`gather_body_state`, NCCL on GPUs / gloo in the CPU tests).  One giant scene is
not split (north star): "replicas only".
"""
import numpy as np

from . import abi


def shard_range(n_worlds, rank, world_size):
    """World-index range [w0, w1) of `rank`: world w lives on rank floor(w * world_size / n_worlds)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    w0 = (n_worlds * rank + world_size - 1) // world_size
    w1 = (n_worlds * (rank + 1) + world_size - 1) // world_size
    return w0, w1


def owner_of(world, n_worlds, world_size):
    return (world * world_size) // n_worlds


def gather_body_state(local, n_worlds_total, device="cpu", group=None):
    """All-gather the in/out state of every rank's shard (shards may be ragged).  `local` is this rank's abi.Bodies.
    Returns {field: ndarray over ALL bodies in world order} on every rank."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    counts = torch.zeros(ws, dtype=torch.int64, device=device)
    mine = torch.tensor([local.n_bodies], dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, mine, group=group) if device != "cpu" else dist.all_gather(list(counts.split(1)), mine, group=group)
    counts = [int(c) for c in counts.cpu()]
    cap = max(max(counts), 1)
    out = {}
    for name in abi.BODY_INOUT:
        a = getattr(local, name).reshape(local.n_bodies, -1)
        width = a.shape[1]
        send = torch.zeros((cap, width), dtype=torch.float64, device=device)
        send[:local.n_bodies] = torch.from_numpy(np.ascontiguousarray(a)).to(device)
        recv = [torch.empty_like(send) for _ in range(ws)]
        dist.all_gather(recv, send, group=group)
        full = np.concatenate([r[:counts[k]].cpu().numpy() for k, r in enumerate(recv)], axis=0)
        out[name] = full if width > 1 else full.reshape(-1)
    return out


def gather_state_device(engine, n_bodies, group=None, send=None, recv=None):
    """The final state gather of a sharded batch with no host staging: the 22 doubles per body a step writes (transform3d 7,
    velocity 3, angular velocity 3, invInertiaWorld 9; src/Internal/SolverBody.elm:97-225) are packed on the device by
    ep_pack_state_device straight into the send buffer of ONE ncclAllGather (torch.distributed.all_gather_into_tensor on the
    current CUDA stream).  Every rank must hold the same number of bodies (equal shards) or pass `n_bodies` = the largest
    shard.  Returns (recv [world_size, n_bodies, 22] on the device, send); pass them back in to reuse the buffers."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    if send is None:
        send = torch.zeros((n_bodies, 22), dtype=torch.float64, device="cuda")
        recv = torch.empty((ws * n_bodies, 22), dtype=torch.float64, device="cuda")
    engine.pack_state_device(send.data_ptr(), torch.cuda.current_stream().cuda_stream)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv.view(ws, n_bodies, 22), send
