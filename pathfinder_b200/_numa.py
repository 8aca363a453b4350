"""Bind a rank to the host NUMA node its GPU hangs off.

One process per GPU moves ~75 MB per ELBO step over PCIe (packed trajectories up, fits and tables down).
With N ranks on a two-socket host and no placement, the page-locked staging buffers of several ranks land
on one socket and every copy crosses the inter-socket link; first-touch after `sched_setaffinity` to the
GPU's node keeps each rank's buffers local.  Pure host plumbing: nothing here touches the compute path,
and every failure (no sysfs entry, cgroup-restricted CPU set, single-node host) leaves the process as it was.
"""
from __future__ import annotations

import os


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


def device_numa_node(device: int):
    """NUMA node of CUDA device `device` (PCI bus id -> sysfs), or None when the host does not say."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(device)
        bus_id = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
    except Exception:
        return None
    try:
        with open(f"/sys/bus/pci/devices/{bus_id}/numa_node") as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_device_numa(device: int):
    """Restrict this process to the CPUs of the device's NUMA node (intersected with the CPUs it may already
    use).  Returns {"node": k, "cpus": count} on success, None when nothing was changed."""
    node = device_numa_node(device)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if len(target) < 4 or target == allowed:  # leave a starved or single-node CPU set alone
            return None
        os.sched_setaffinity(0, target)
    except (OSError, ValueError, AttributeError):
        return None
    return {"node": node, "cpus": len(target)}
