"""Host placement of a rank next to its GPU.

The host-buffer path of the solver is bound by device->host copies (80 B per cell-hour of FP64 results).  With one
process per GPU every rank pins its own result buffers; where those pages live decides whether eight concurrent DMA
streams each write into the memory of the socket their GPU hangs off, or all into one node across the inter-socket
link.  `bind_to_gpu(local_rank)` restricts the calling process to the CPUs sysfs lists as local to the GPU
(`/sys/bus/pci/devices/<bus id>/local_cpulist`, intersected with the CPUs the process may use) and asks the kernel to
prefer that node for its future allocations (set_mempolicy(MPOL_PREFERRED)); pinned allocations made AFTER the call —
torch's pin_memory(), the library's staging slots — are then first-touched on that node.  Everything is best effort:
containers often hide the topology or forbid set_mempolicy, and the function reports what it could do instead of failing.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

_SYS_SET_MEMPOLICY = 238  # x86-64
_MPOL_PREFERRED = 1


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(s: Optional[str]) -> List[int]:
    out: List[int] = []
    if not s:
        return out
    for tok in s.split(","):
        tok = tok.strip()
        if not tok:
            continue
        if "-" in tok:
            a, b = tok.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(tok))
    return out


def gpu_bus_id(device: int) -> Optional[str]:
    try:
        import torch

        props = torch.cuda.get_device_properties(device)
        return f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    except Exception:
        return None


def gpu_locality(device: int) -> Dict[str, object]:
    """{bus, node, cpus}: NUMA node and local CPUs of a CUDA device as sysfs reports them (node -1 / [] if hidden)."""
    bus = gpu_bus_id(device)
    node, cpus = -1, []
    if bus:
        base = f"/sys/bus/pci/devices/{bus}"
        n = _read(base + "/numa_node")
        if n is not None:
            try:
                node = int(n)
            except ValueError:
                node = -1
        cpus = parse_cpulist(_read(base + "/local_cpulist"))
    return {"bus": bus, "node": node, "cpus": cpus}


def bind_to_gpu(device: int, set_memory_policy: bool = True) -> Dict[str, object]:
    """Bind the calling process next to CUDA device `device`; returns a report (never raises)."""
    loc = gpu_locality(device)
    rep: Dict[str, object] = {"bus": loc["bus"], "node": loc["node"], "cpus_bound": 0, "mempolicy": "unchanged"}
    try:
        allowed = os.sched_getaffinity(0)
        want = allowed & set(loc["cpus"])  # type: ignore[arg-type]
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            rep["cpus_bound"] = len(want)
        elif want:
            rep["cpus_bound"] = len(want)
    except (AttributeError, OSError) as exc:
        rep["affinity_error"] = repr(exc)
    node = int(loc["node"])  # type: ignore[arg-type]
    if set_memory_policy and node >= 0:
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            rc = libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_PREFERRED, mask, ctypes.c_ulong(1024))
            rep["mempolicy"] = f"preferred node {node}" if rc == 0 else f"refused (errno {ctypes.get_errno()})"
        except Exception as exc:  # pragma: no cover - platform dependent
            rep["mempolicy"] = f"unavailable ({exc!r})"
    return rep
