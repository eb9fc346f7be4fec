"""Host-side placement for the end-to-end (host buffer) path: bind a rank to the CPUs of its GPU's NUMA node
BEFORE it allocates pinned memory, so cudaHostAlloc's pages land on the socket the GPU's PCIe root hangs off.
Round 1's end-to-end numbers scaled 1 / 0.98 / 0.66 / 0.56 at 1 / 2 / 4 / 8 GPUs with ranks and their pinned
buffers placed wherever the launcher happened to start them (GPU0-3 sit on node 0, GPU4-7 on node 1)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path


def parse_cpulist(text: str) -> list[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]"""
    cpus: list[int] = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def normalise_pci_bus_id(bus_id: str) -> str:
    """nvidia-smi prints '00000000:1B:00.0'; sysfs wants '0000:1b:00.0'."""
    b = bus_id.strip().lower()
    dom, rest = b.split(":", 1)
    return f"{int(dom, 16):04x}:{rest}"


def gpu_numa_node(index: int, sysfs: str = "/sys", pci_bus_id: str | None = None) -> int | None:
    """NUMA node of a GPU, or None if unknown.  pci_bus_id ('0000:1b:00.0', e.g. built from
    torch.cuda.get_device_properties) wins; otherwise nvidia-smi is asked for physical GPU `index`."""
    try:
        if pci_bus_id is None:
            out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
            if not out:
                return None
            pci_bus_id = out[0]
        node = int(Path(sysfs, "bus/pci/devices", normalise_pci_bus_id(pci_bus_id), "numa_node").read_text().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node: int, sysfs: str = "/sys") -> list[int]:
    try:
        return parse_cpulist(Path(sysfs, f"devices/system/node/node{node}/cpulist").read_text())
    except Exception:
        return []


def bind_to_gpu_node(index: int, sysfs: str = "/sys", pci_bus_id: str | None = None) -> dict:
    """Restrict this process to the CPUs of GPU `index`'s NUMA node (intersected with its current affinity).
    Returns {"node": n | None, "cpus": count, "bound": bool}; never raises."""
    info = {"node": None, "cpus": len(os.sched_getaffinity(0)), "bound": False}
    node = gpu_numa_node(index, sysfs, pci_bus_id)
    if node is None:
        return info
    info["node"] = node
    cpus = set(node_cpus(node, sysfs)) & set(os.sched_getaffinity(0))
    if not cpus:
        return info
    try:
        os.sched_setaffinity(0, cpus)
        info["cpus"], info["bound"] = len(cpus), True
    except Exception:
        pass
    return info
