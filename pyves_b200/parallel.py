"""Walker sharding and the one collective of the path: the sum of the basis moments.

Walkers are independent between coefficient updates, so rank r of R owns the contiguous block of
global walker ids ``shard(N, r, R)`` and Philox counters use global ids (trajectories do not depend on
R).  Per update every rank contributes ``[sum_w, sum_w f (K), sum_w f f (K*K, optional)]`` in FP64 to a
single ``all_reduce(SUM)`` (NCCL over NVLink on GPUs, gloo in the CPU tests); every rank then applies
the identical deterministic optimiser step, so no broadcast of coefficients is needed.
"""
import torch


def dist_info(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard(n_total, rank, world):
    """Walkers [lo, hi) owned by ``rank``: contiguous blocks, remainder to the low ranks."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_moments(sum_w, sum_wf, sum_wff=None, group=None):
    """Sum the moment tensors over all ranks with ONE all_reduce; returns new tensors."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sum_w, sum_wf, sum_wff
    K = sum_wf.numel()
    parts = [sum_w.reshape(-1), sum_wf.reshape(-1)] + ([sum_wff.reshape(-1)] if sum_wff is not None else [])
    flat = torch.cat(parts)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat[:1], flat[1:1 + K], (flat[1 + K:].reshape(K, K) if sum_wff is not None else None)


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node), so that the
    pinned host buffers it allocates afterwards (first touch) and the threads that feed the copies sit next to that GPU's
    PCIe root.  With one process per GPU all moving walker state through host memory at once, remote-node buffers make the
    ranks collide on the inter-socket link.  Returns the core list, or None when NVML / affinity control is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
