"""NUMA placement of a site process next to its GPU.

An 8xB200 box has two CPU sockets; GPUs 0-3 hang off one, 4-7 off the other.  A rank that runs (and pins its staging
buffers) on the far socket pays a cross-socket hop for every host->device DMA and every launch, and eight unpinned
ranks with default-sized intra-op thread pools oversubscribe the cores.  ``pin_to_gpu`` is called once per process,
before the first pinned allocation: it restricts the process to the CPUs of the GPU's NUMA node (first-touch then
places pinned pages there too) and sizes torch's intra-op pool to the rank's share of that node.
"""
import os as _os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def _bus_id(index):
    import torch
    pr = torch.cuda.get_device_properties(index)
    dom, bus, dev = getattr(pr, 'pci_domain_id', 0), getattr(pr, 'pci_bus_id', None), getattr(pr, 'pci_device_id', 0)
    if bus is None:
        return None
    return f'{dom:04x}:{bus:02x}:{dev:02x}.0'


def gpu_numa_cpus(index):
    """(numa_node, set of cpu ids local to GPU ``index``) from sysfs, NVML as fall-back; (None, None) if unknown."""
    try:
        bus = _bus_id(index)
    except Exception:
        bus = None
    if bus:
        base = f'/sys/bus/pci/devices/{bus}'
        try:
            node = int(open(f'{base}/numa_node').read().strip())
            cpus = _parse_cpulist(open(f'{base}/local_cpulist').read())
            if node < 0:      # single-socket / virtualised: whatever sysfs calls local
                node = 0
            if cpus:
                return node, cpus
        except Exception:
            pass
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()) if bus else pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (_os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        return None, cpus or None
    except Exception:
        return None, None


def pin_to_gpu(index, ranks_per_node=None, max_threads=8):
    """Bind this process to the CPUs local to GPU ``index``.  Returns a small report dict (also useful in logs);
    never raises - on boxes where the topology cannot be read the process is left alone."""
    report = {'gpu': index, 'numa_node': None, 'cpus': None, 'threads': None, 'pinned': False}
    if _os.environ.get('COINN_NO_AFFINITY') == '1' or not hasattr(_os, 'sched_setaffinity'):
        return report
    node, cpus = gpu_numa_cpus(index)
    try:
        allowed = _os.sched_getaffinity(0)
    except Exception:
        allowed = None
    if cpus and allowed:
        cpus = cpus & allowed
    if not cpus:
        return report
    try:
        _os.sched_setaffinity(0, cpus)
        report.update(numa_node=node, cpus=len(cpus), pinned=True)
    except Exception:
        return report
    try:
        import torch
        share = ranks_per_node or 4
        n = max(1, min(int(max_threads), len(cpus) // max(share, 1)))
        torch.set_num_threads(n)
        report['threads'] = n
    except Exception:
        pass
    return report
