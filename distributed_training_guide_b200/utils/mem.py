"""Memory statistics for the log record (reference ``get_mem_stats``, ``01-...:248-257``)."""
from __future__ import annotations

import torch


def get_mem_stats(device=None):
    device = torch.device(device) if device is not None else None
    if device is None or device.type != "cuda" or not torch.cuda.is_available():
        try:
            import psutil

            vm = psutil.virtual_memory()
            rss = psutil.Process().memory_info().rss
            total, cur = vm.total, rss
        except Exception:  # pragma: no cover
            total, cur = 0, 0
        g = 1e-9
        return {"total_gb": g * total, "curr_alloc_gb": g * cur, "peak_alloc_gb": g * cur,
                "curr_resv_gb": g * cur, "peak_resv_gb": g * cur}
    mem = torch.cuda.memory_stats(device)
    props = torch.cuda.get_device_properties(device)
    extra = 0
    try:  # symmetric (NVLink-visible) buffers live outside the caching allocator
        from ..parallel import symm

        extra = symm.allocated_bytes()
    except Exception:
        pass
    return {
        "total_gb": 1e-9 * props.total_memory,
        "curr_alloc_gb": 1e-9 * (mem["allocated_bytes.all.current"] + extra),
        "peak_alloc_gb": 1e-9 * (mem["allocated_bytes.all.peak"] + extra),
        "curr_resv_gb": 1e-9 * (mem["reserved_bytes.all.current"] + extra),
        "peak_resv_gb": 1e-9 * (mem["reserved_bytes.all.peak"] + extra),
    }


def reset_peak(device=None):
    device = torch.device(device) if device is not None else None
    if device is not None and device.type == "cuda":
        torch.cuda.reset_peak_memory_stats(device)
