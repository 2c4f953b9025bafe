"""
Host-memory placement for one-process-per-GPU runs.

Every rank streams its CSR batches out of pinned host memory (~150 MB per 100 000 positions of
BASELINE config 3).  On a two-socket box the pinned pages of a rank land on the NUMA node of the
core that touched them first; a rank whose pages live on the other socket pulls every batch over
the inter-socket link, and with eight ranks that link -- not PCIe, not the GPU -- sets the
end-to-end rate.  ``bind_rank_to_gpu_node`` pins the calling process (and with it every later
first-touch allocation, cudaHostAlloc included) to the cores of the NUMA node the GPU hangs off,
and asks the kernel to prefer that node for new pages.  It reads sysfs only, needs no libnuma,
and is a no-op (with the reason reported) where the topology is not visible or has one node.

Call it once per rank, before the first pinned allocation (bench.py and
readers.TranscriptPipeline do).  The reference has no equivalent: its workers share GPUs through
a queue (run.py:104-131) and move dense tensors with ``.to(device)`` from pageable memory.
"""
import ctypes
import os

_SYS_set_mempolicy = 238      # x86_64
_MPOL_PREFERRED = 1


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11}"""
    cpus = set()
    for part in text.strip().split(','):
        part = part.strip()
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-', 1)
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def gpu_pci_address(device_index):
    """'dddd:bb:dd.f' of a CUDA device (as sysfs spells it), or None."""
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def gpu_numa_node(pci_address, sysfs='/sys'):
    """(node, cpus of that node) for a PCI device; node is None where sysfs does not say."""
    if not pci_address:
        return None, set()
    base = os.path.join(sysfs, 'bus', 'pci', 'devices', pci_address)
    node = _read(os.path.join(base, 'numa_node'))
    cpus = _read(os.path.join(base, 'local_cpulist'))
    try:
        node = int(node) if node is not None else None
    except ValueError:
        node = None
    if node is not None and node < 0:
        node = None
    if node is not None:
        node_cpus = _read(os.path.join(sysfs, 'devices', 'system', 'node', f'node{node}', 'cpulist'))
        if node_cpus:
            cpus = node_cpus
    return node, (parse_cpulist(cpus) if cpus else set())


def n_numa_nodes(sysfs='/sys'):
    online = _read(os.path.join(sysfs, 'devices', 'system', 'node', 'online'))
    return len(parse_cpulist(online)) if online else 1


def prefer_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}); returns 0 or the errno (containers may refuse)."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(_SYS_set_mempolicy, _MPOL_PREFERRED, ctypes.byref(mask),
                          ctypes.c_ulong(ctypes.sizeof(mask) * 8))
        return 0 if rc == 0 else ctypes.get_errno()
    except Exception:
        return -1


def bind_rank_to_gpu_node(device_index, sysfs='/sys', pci_address=None, apply=True):
    """Restricts the calling process to the cores next to GPU ``device_index`` and prefers that
    node's memory.  Returns a dict describing what was found and done (goes into bench.py's JSON
    line).  Never raises."""
    info = {"applied": False, "node": None, "nodes": None, "cpus": None, "mempolicy": None, "why": None}
    try:
        info["nodes"] = n_numa_nodes(sysfs)
        addr = pci_address or gpu_pci_address(device_index)
        node, cpus = gpu_numa_node(addr, sysfs)
        info["node"] = node
        if node is None:
            info["why"] = "sysfs does not name the GPU's NUMA node"
            return info
        if info["nodes"] <= 1:
            info["why"] = "one NUMA node"
            return info
        allowed = os.sched_getaffinity(0)
        local = sorted(allowed & cpus)
        if not local:
            info["why"] = "no allowed core on the GPU's node"
            return info
        info["cpus"] = len(local)
        if apply:
            os.sched_setaffinity(0, local)
            info["mempolicy"] = prefer_node(node)
            info["applied"] = True
    except Exception as e:  # topology files differ between kernels; placement is an optimisation
        info["why"] = f"{type(e).__name__}: {e}"
    return info
