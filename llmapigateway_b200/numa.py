"""Host-side placement: run the process (and so its first-touched / pinned pages) on the CPUs of the NUMA node the GPU hangs off.
Pinned staging buffers that live on the far socket cost a cross-socket hop on every H2D/D2H byte (SURVEY 8(b) ownership of the
staging buffers).  Pure /sys reads and sched_setaffinity: no numactl needed."""
from __future__ import annotations

import os
from pathlib import Path


def _parse_cpulist(text: str) -> set[int]:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(device: int = 0) -> int | None:
    """NUMA node of CUDA device `device` (from its PCI bus id), None when the platform does not say."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node")
        node = int(path.read_text().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device: int = 0) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with what it may already use).  Returns what was
    done, for the record."""
    node = gpu_numa_node(device)
    info = {"gpu_numa_node": node, "bound": False}
    if node is None:
        return info
    try:
        cpus = _parse_cpulist(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text())
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as ex:
        info["error"] = repr(ex)
    return info
