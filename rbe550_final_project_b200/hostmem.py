"""Host-side placement for the host-buffer entry points (pv_check_states_host*, pv_check_edges_host): on a multi-socket
box a rank's pinned rows should live in the memory of the socket its GPU hangs off.  With one process per GPU and no
placement every rank's pinned buffers land wherever the launcher happened to start it -- usually all on one socket, whose
memory controllers then feed all PCIe links (measured on an 8-GPU box: 24 GB/s per GPU with 8 ranks against 54 GB/s with
4; profiles/r2_notes.md).  `bind_to_gpu_numa` moves the calling process onto the CPUs NVML reports as local to the GPU
(and, where the kernel allows it, prefers that node for new pages), so buffers pinned AFTERWARDS are local.

Not on the reference's path (planning.py is one process, one robot); used by bench.py at N > 1 and by callers of
distributed.* that feed host buffers."""
import ctypes
import os


def _pci_bus_id(device_index: int) -> str:
    import torch

    p = torch.cuda.get_device_properties(device_index)
    return f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"


def _node_of_cpu(cpu: int):
    base = f"/sys/devices/system/cpu/cpu{cpu}"
    try:
        for name in os.listdir(base):
            if name.startswith("node") and name[4:].isdigit():
                return int(name[4:])
    except OSError:
        pass
    return None


def bind_to_gpu_numa(device_index: int, prefer_memory: bool = True) -> dict:
    """Restrict the calling process to the CPUs local to CUDA device `device_index` (intersected with the CPUs it may use
    already) and prefer that NUMA node for its new pages.  Returns what was done; never raises: on a box without NVML, with
    one node, or with a cpuset that excludes the GPU's CPUs it changes nothing and says why."""
    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(_pci_bus_id(device_index).encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        info.update(gpu_local_cpus=len(local), allowed_cpus=len(allowed), usable=len(cpus))
        if not cpus or cpus == allowed:
            info["why"] = "no usable local CPUs" if not cpus else "already local (one node, or the launcher placed the rank)"
            return info
        os.sched_setaffinity(0, cpus)
        info["bound"] = True
        node = _node_of_cpu(min(cpus))
        info["node"] = node
        if prefer_memory and node is not None and node < 64:
            # set_mempolicy(MPOL_PREFERRED, {node}): new pages from the GPU's node while it has room (x86-64 syscall 238,
            # aarch64 237); CPU affinity alone already gives first-touch-local pages, this also covers migrated threads
            try:
                nr = {"x86_64": 238, "aarch64": 237}.get(os.uname().machine)
                if nr is not None:
                    mask = ctypes.c_ulong(1 << node)
                    rc = ctypes.CDLL(None, use_errno=True).syscall(nr, 1, ctypes.byref(mask), ctypes.c_ulong(65))
                    info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
            except Exception as exc:  # noqa: BLE001
                info["mempolicy"] = repr(exc)
    except Exception as exc:  # noqa: BLE001  (placement is an optimisation, never a failure)
        info["why"] = repr(exc)
    return info
