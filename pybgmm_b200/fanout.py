"""Multi-GPU fan-out of the hot path: one process per GPU, no data-path collective.

A single exact chain does not shard (every draw depends on the globally updated statistics), so the two fan-outs
are (1) independent chains, one per rank, and (2) disjoint contiguous row shards, one independent CRP/pCRP per
rank with its own labels.  The only communication is the gather of assignments at the end, done with
torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(N, world_size, rank):
    """Contiguous row range [lo, hi) of `rank` (sizes differ by at most one row)."""
    base, rem = divmod(int(N), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def offset_labels(z_per_rank, K_per_rank):
    """Concatenate per-shard assignments into one global labelling: shard g's label k becomes k + sum_{h<g} K_h
    (unassigned -1 stays -1)."""
    out, off = [], 0
    for z, K in zip(z_per_rank, K_per_rank):
        z = np.asarray(z, dtype=np.int64)
        out.append(np.where(z >= 0, z + off, -1))
        off += int(K)
    return np.concatenate(out)


def gather_assignments(z_local, K_local, group=None):
    """All-gather the per-rank assignment vectors (torch int64 tensors, equal length on every rank, on the device of
    the backend: CUDA for NCCL, CPU for gloo).  Returns (list of per-rank tensors, list of per-rank K)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    outs = [torch.empty_like(z_local) for _ in range(world)]
    dist.all_gather(outs, z_local, group=group)
    k = torch.tensor([int(K_local)], dtype=torch.int64, device=z_local.device)
    ks = [torch.empty_like(k) for _ in range(world)]
    dist.all_gather(ks, k, group=group)
    return outs, [int(t.item()) for t in ks]


def chain_assignments_tensor(chain, device):
    """Relabelled assignments of a `_lib.Chain` as a torch int64 tensor on `device`, written by the engine directly
    into the tensor's storage (no host round trip)."""
    import torch
    z = torch.empty(chain.N, dtype=torch.int64, device=device)
    chain.assignments_to_device(z.data_ptr())
    torch.cuda.synchronize(device)
    return z
