"""Multi-GPU plumbing: the path shards over independent sequences,
so there is no data-path collective -- only a placement rule and one gather of the
per-sequence trajectories at the end of a run (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_sequences(n_sequences: int, world: int, rank: int):
    """Sequence s runs on rank s % world (cyclic; equal-length sequences stay balanced)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return [s for s in range(n_sequences) if s % world == rank]


def owner_of(sequence: int, world: int) -> int:
    return sequence % world


def gather_trajectories(local, n_sequences: int, world: int, rank: int, device=None):
    """local: float32 [n_local][frames][12] for shard_sequences(...) in order (numpy or torch).
    Returns float32 numpy [n_sequences][frames][12] on every rank, ordered by sequence index.
    Ranks may own different numbers of sequences; shorter shards are zero-padded for the
    collective and the padding is dropped afterwards."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(local, dtype=torch.float32)
    if device is not None:
        t = t.to(device)
    frames = t.shape[1] if t.ndim == 3 else 0
    per_rank = (n_sequences + world - 1) // world
    pad = torch.zeros((per_rank, frames, 12), dtype=torch.float32, device=t.device)
    if t.shape[0]:
        pad[: t.shape[0]] = t
    if world == 1 or not dist.is_initialized():
        allr = pad.unsqueeze(0)
    else:
        allr = torch.empty((world, per_rank, frames, 12), dtype=torch.float32, device=t.device)
        dist.all_gather_into_tensor(allr.view(world * per_rank, frames, 12), pad)
    allr = allr.cpu().numpy()
    out = np.zeros((n_sequences, frames, 12), dtype=np.float32)
    for r in range(world):
        for k, s in enumerate(shard_sequences(n_sequences, world, r)):
            out[s] = allr[r, k]
    return out
