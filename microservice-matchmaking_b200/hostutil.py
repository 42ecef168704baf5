"""Host-side helpers of the ingest path: NUMA placement of the calling process next to its GPU.

The end-to-end rate of the search stage is bounded by host <-> device copies of the player columns;
with several ranks per box every rank must stream from the memory of the socket its GPU hangs off, otherwise
the copies of 8 ranks cross the inter-socket link."""
import os


def _pci_bus_id(device):
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        return f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    except Exception:
        pass
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", f"--id={int(device)}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        return out[-12:] if len(out) >= 12 else None  # 00000000:1B:00.0 -> 0000:1b:00.0
    except Exception:
        return None


def gpu_numa_node(device):
    """NUMA node of a CUDA device from sysfs, or None when it cannot be told (single-socket boxes report -1)."""
    bus = _pci_bus_id(device)
    if not bus:
        return None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _cpus_of_node(node):
    cpus = []
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa(device):
    """Pin this process to the cores of the GPU's NUMA node (first-touch then places pinned buffers there).
    Returns a dict describing what was done (recorded in the bench line)."""
    node = gpu_numa_node(device)
    info = {"gpu": int(device), "numa_node": node, "bound": False}
    if node is None:
        return info
    try:
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in _cpus_of_node(node) if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
    except Exception as exc:  # containers without the sysfs files / CAP_SYS_NICE: run unbound
        info["error"] = repr(exc)
    return info
