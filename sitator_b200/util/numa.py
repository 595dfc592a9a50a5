"""NUMA placement of a rank's host buffers (frame-sharded runs, one process per GPU).

The trajectory enters through page-locked host memory and the labels leave the same way; with eight ranks copying
at once the copies are bound by host memory / inter-socket bandwidth unless every rank's buffers live on the NUMA
node its GPU hangs off.  ``bind_to_gpu_node`` restricts the calling process to the CPUs local to the GPU *before* the
buffers are allocated (first touch then places the pages on that node).  Linux only; does nothing where the
information is missing.  No reference counterpart (the reference is a single CPU process).
"""
import os


def _pci_bus_id(device_index):
    import torch
    props = torch.cuda.get_device_properties(device_index)
    try:
        return "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    except AttributeError:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_info(device_index):
    """{'bus': ..., 'numa_node': int or None, 'local_cpus': count} for a CUDA device."""
    bus = _pci_bus_id(device_index)
    out = {"bus": bus, "numa_node": None, "local_cpus": 0}
    if bus is None:
        return out
    base = "/sys/bus/pci/devices/" + bus
    try:
        out["numa_node"] = int(open(base + "/numa_node").read().strip())
    except Exception:
        pass
    try:
        out["local_cpus"] = len(_parse_cpulist(open(base + "/local_cpulist").read()))
    except Exception:
        pass
    return out


def bind_to_gpu_node(device_index):
    """Restrict this process to the CPUs local to the GPU (intersection with the current affinity mask).
    Returns True if the affinity was changed."""
    bus = _pci_bus_id(device_index)
    if bus is None or not hasattr(os, "sched_setaffinity"):
        return False
    try:
        local = _parse_cpulist(open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read())
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target or target == allowed:
            return False
        os.sched_setaffinity(0, target)
        return True
    except Exception:
        return False
