"""Multi-GPU plumbing for the two ways the path shards.

* Independent jobs (serving, BASELINE config 5): jobs are assigned to ranks round-robin; every rank
  runs its jobs on its own GPU with no data-path collective.  ``torch.distributed`` (NCCL on GPUs,
  gloo in CPU tests) is used only for barriers, the max-over-ranks timing and gathering small results.
* One large canvas split in row strips (config 4, ``tiled.TiledTransfer``): halo rows travel as peer
  stores inside libst2, the Gram / loss / dot-product sums are all-reduced with NCCL; ``strip_bounds``
  fixes the partition.
"""
import torch
import torch.distributed as dist


def shard_jobs(n_jobs, world_size, rank):
    """Indices of the jobs rank ``rank`` runs (round-robin: job j -> rank j % world_size)."""
    if not 0 <= rank < world_size:
        raise ValueError('rank %d outside world of %d' % (rank, world_size))
    return list(range(rank, n_jobs, world_size))


def strip_bounds(height, world_size, align=32):
    """Row strips for spatial tiling: boundaries at multiples of ``align`` (32 = 2^5 rows keep the
    windows of all five 2x2/2 ceil-mode pools, pool1 .. pool5, inside one strip), sizes as equal as that allows.  Returns [(row0, row1)] per rank;
    trailing ranks may get empty strips on tiny canvases."""
    blocks = (height + align - 1) // align
    out, start = [], 0
    for r in range(world_size):
        nb = blocks // world_size + (1 if r < blocks % world_size else 0)
        end = min(height, start + nb * align)
        out.append((start, end))
        start = end
    return out


def all_max(value, device=None):
    """Max of a Python float over all ranks (device-timed milliseconds -> slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_objects(obj):
    """List of every rank's picklable ``obj`` on every rank (per-job results of a serving run)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def run_jobs(jobs, step_fn, world_size=1, rank=0):
    """Run ``step_fn(job_index, job)`` for this rank's share of ``jobs``; returns {index: result}."""
    return {j: step_fn(j, jobs[j]) for j in shard_jobs(len(jobs), world_size, rank)}


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers
    allocated afterwards (first touch) are local to the GPU's PCIe root: with one process per GPU the
    host<->device copies of all ranks otherwise share whatever node the launcher happened to start them on.
    Returns the node number, or None when the topology is not exposed (then nothing is changed)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
