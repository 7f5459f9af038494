"""One process per GPU (torch.distributed): the frequency tables are built once on rank 0 and
replicated with a single broadcast of the device image over NCCL/NVLink; subjects are sharded
by contiguous input ranges, there is no collective on the per-subject path, and rank 0
concatenates the per-rank output rows in input order.  `.miss` / `.problem` rows
carry global line indices, so every shard is given the index of its first line.

torch is used for the process group and the broadcast only (plumbing)."""
import ctypes as C

from . import _lib
from .networkx_graph import Graph

FILE_KEYS = ("umug", "umug_pops", "pmug", "pmug_pops", "miss", "problem")


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of rank `rank` out of `world` over n input lines."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DeviceView(object):
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def bind_to_device_numa_node(device):
    """Pins this process to the CPUs next to GPU `device` (sysfs local_cpulist of its PCI function) so
    that the pinned staging buffers it allocates afterwards are first-touched on that NUMA node.  One
    process per GPU on a multi-socket box otherwise funnels every host<->device copy through whichever
    socket the scheduler happened to pick.  Returns the CPU set used, or None if nothing was changed."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def broadcast_graph(graph, config, device, src=0):
    """Replicates the tables of rank `src` to every rank's GPU.  On `src`, `graph` is a built
    Graph; elsewhere it may be None.  Returns this rank's Graph."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = torch.device("cuda", device)
    meta = [None]
    if rank == src:
        lib = graph.lib
        nbytes = C.c_int64()
        _lib.check(lib.grimb_tables_image_size(graph.handle, C.byref(nbytes)), "grimb_tables_image_size", lib)
        meta = [(nbytes.value, graph.alleles)]
    dist.broadcast_object_list(meta, src=src)
    nbytes, alleles = meta[0]
    if rank == src:
        ptr = C.c_void_p()
        _lib.check(lib.grimb_tables_image_ptr(graph.handle, C.byref(ptr)), "grimb_tables_image_ptr", lib)
        img = torch.as_tensor(_DeviceView(ptr.value, nbytes), device=dev)
    else:
        img = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(img, src)          # the one collective: ncclBroadcast of the table image
    torch.cuda.synchronize(dev)
    if rank == src:
        return graph
    g = Graph(config, device=device)
    g._set_dictionaries(alleles)
    lib = g.lib
    h = C.c_void_p()
    _lib.check(lib.grimb_tables_from_image(C.c_void_p(img.data_ptr()), nbytes, device, C.byref(h)),
               "grimb_tables_from_image", lib)
    g.handle = h
    return g


def impute_lines_sharded(imputation, lines, rank, world, gather):
    """Imputes this rank's contiguous share of `lines`; `gather(obj)` must return the list of all
    ranks' objects on rank 0 (None elsewhere).  Returns the concatenated file rows on rank 0."""
    lo, hi = shard_range(len(lines), rank, world)
    mine = imputation.impute_lines(lines[lo:hi], first_index=lo)
    parts = gather(mine)
    if parts is None:
        return None
    out = {k: [] for k in FILE_KEYS}
    for part in parts:            # rank order == input order
        for k in FILE_KEYS:
            out[k].extend(part[k])
    return out
