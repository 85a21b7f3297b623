"""Host-side placement for one-process-per-GPU deployments: run the process (and therefore first-touch its pinned
staging memory) on the NUMA node the GPU's PCIe root hangs off.  With 8 ranks feeding 8 GPUs from host memory the
copies otherwise cross the socket interconnect and share one node's DRAM (round 1, 16 GB per rank and step:
292 ms at 1-2 GPUs, 560 ms at 4, 696 ms at 8).  Pure host logic (sysfs + sched_setaffinity); every step is optional and
reports what it did instead of failing."""
import os
from typing import Dict, Optional, Set


def _parse_cpulist(s: str) -> Set[int]:
    out: Set[int] = set()
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        else:
            out.add(int(part))
    return out


def pci_address(device: int) -> Optional[str]:
    """'dddd:bb:dd.0' of CUDA device `device` (torch device properties)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def numa_node_of(pci: str, sysfs: str = "/sys") -> Optional[int]:
    try:
        v = int(open(os.path.join(sysfs, "bus/pci/devices", pci, "numa_node")).read().strip())
        return v if v >= 0 else None
    except Exception:
        return None


def cpus_of_node(node: int, sysfs: str = "/sys") -> Set[int]:
    try:
        return _parse_cpulist(open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")).read())
    except Exception:
        return set()


class bind_to_device_numa:
    """Context manager: restrict the process to the CPUs of the GPU's NUMA node (intersected with the CPUs it may
    already use), restore the previous mask on exit.  `.info` says what happened."""

    def __init__(self, device: int, sysfs: str = "/sys", pci: Optional[str] = None):
        self.device, self.sysfs, self.pci = device, sysfs, pci
        self.info: Dict[str, object] = {"bound": False}
        self._old = None

    def __enter__(self):
        try:
            pci = self.pci or pci_address(self.device)
            node = numa_node_of(pci, self.sysfs) if pci else None
            self.info.update({"pci": pci, "numa_node": node})
            if node is None:
                self.info["why"] = "no NUMA node reported for the device"
                return self
            allowed = os.sched_getaffinity(0)
            want = cpus_of_node(node, self.sysfs) & allowed
            if not want:
                self.info["why"] = "none of the node's CPUs is available to this process"
                return self
            self._old = allowed
            os.sched_setaffinity(0, want)
            self.info.update({"bound": True, "cpus": len(want)})
        except Exception as ex:  # noqa: BLE001
            self.info["why"] = repr(ex)[:120]
        return self

    def __exit__(self, *exc):
        if self._old is not None:
            try:
                os.sched_setaffinity(0, self._old)
            except Exception:
                pass
        return False
