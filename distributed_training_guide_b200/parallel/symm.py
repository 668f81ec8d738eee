"""NVLink symmetric memory: buffers every rank of a group can address directly.

This is the substrate the reference does not have (it only ever calls NCCL):
each rank ``cudaMalloc``s the same size outside the caching allocator, exports a CUDA-IPC handle,
handles are exchanged once through ``torch.distributed`` (NCCL/gloo is used for this bootstrap
only), and every rank maps its peers' buffers.  Collective *kernels* (``csrc/comm.cu``,
``csrc/fused_tp.cu``) then load/store peer memory over NVLink 5 / NVSwitch and synchronise with
device-side epoch flags in a symmetric signal pad — no host round trips, no NCCL on the hot path.

``SymmGroup`` also runs on a single rank (N = 1: same kernels, no peers), which is how the
single-GPU chapter shares the fused optimizer path.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .. import _ext

_LIVE_GROUPS = []


def allocated_bytes() -> int:
    c = _ext.load(False)
    return int(c.symm_allocated_bytes()) if c is not None and hasattr(c, "symm_allocated_bytes") else 0


def post_mortem(timeout_s: float = 5.0):
    """Signal-pad state of every live group, read on a side stream with a bounded wait so it also works while a
    kernel of this process is spinning on the device (a hung run's last words; used by bench.py's watchdog)."""
    lines = []
    for i, g in enumerate(list(_LIVE_GROUPS)):
        try:
            lines.append(f"symm group {i} (rank {g.rank}/{g.world}): " + g.describe_pads(timeout_s=timeout_s))
        except Exception as e:  # pragma: no cover - diagnostics only
            lines.append(f"symm group {i}: unavailable ({e!r})")
    return lines


class SymmBuffer:
    """One symmetric allocation: ``local`` (this rank's memory as a tensor) + every rank's base pointer
    (+ the NVSwitch multicast address of the allocation when it was bound to one, else 0)."""

    def __init__(self, local: torch.Tensor, ptrs: List[int], raw: torch.Tensor, mc_ptr: int = 0, handle=None):
        self.local = local
        self.ptrs = ptrs
        self._raw = raw  # keeps the allocation alive
        self.mc_ptr = int(mc_ptr or 0)
        self._handle = handle

    def elem_offset_of(self, view: torch.Tensor) -> int:
        off = view.data_ptr() - self.local.data_ptr()
        assert off >= 0 and off % view.element_size() == 0
        return off // view.element_size()


class SymmGroup:
    def __init__(self, device: torch.device, pg=None, ranks: Optional[List[int]] = None,
                 comm_blocks: Optional[int] = None):
        self.C = _ext.load(required=True)
        self.device = torch.device(device)
        self.pg = pg
        if dist.is_initialized() and (pg is not None or ranks is None):
            self.world = dist.get_world_size(pg)
            self.rank = dist.get_rank(pg)
        else:
            self.world, self.rank = 1, 0
        if self.world not in (1, 2, 4, 8):
            raise ValueError(f"symmetric collectives support 1/2/4/8 ranks, got {self.world}")
        if comm_blocks is None:
            # CTAs per collective kernel.  One rank: the fused kernel is a pure HBM-bound AdamW (14 B/element),
            # it needs the whole chip to reach memory bandwidth.  Several ranks: enough CTAs to keep
            # ~1 MB of 16-byte NVLink loads in flight (~2 us latency x ~800 GB/s) without starving the
            # tensor-core kernels it overlaps with.
            comm_blocks = int(os.environ.get("DTG_COMM_BLOCKS", 256 if self.world == 1 else 96))
        self.comm_blocks = min(comm_blocks, int(self.C.SYMM_MAX_CHANNELS))
        self._peer_handles = []
        self.epoch = 0
        # EXPERIMENTAL (DTG_NVLS=1): allocate through torch.distributed._symmetric_memory so every buffer is also
        # bound to an NVSwitch multicast address; the bucket kernels then reduce in the switch (comm_nvls.cu).
        self.nvls = bool(os.environ.get("DTG_NVLS")) and self.world > 1
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.pads = self.alloc_bytes(int(self.C.SYMM_PAD_BYTES))
        self.pad_ptrs = self.pads.ptrs
        _LIVE_GROUPS.append(self)

    # -- allocation -----------------------------------------------------------------------------
    def _alloc_bytes_multicast(self, nbytes: int) -> SymmBuffer:
        import torch.distributed._symmetric_memory as symm_mem

        pg = self.pg if self.pg is not None else dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(pg.group_name)
        except Exception:
            pass  # newer torch enables groups on demand
        raw = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        hdl = symm_mem.rendezvous(raw, pg)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("DTG_NVLS=1 but this system exposes no NVSwitch multicast (NVLS) address")
        raw.zero_()
        return SymmBuffer(raw, ptrs, raw, mc_ptr=mc, handle=hdl)

    def alloc_bytes(self, nbytes: int) -> SymmBuffer:
        if self.nvls:
            return self._alloc_bytes_multicast(nbytes)
        raw, handle = self.C.symm_alloc(int(nbytes), self.device.index or 0)
        if self.world == 1:
            ptrs = [raw.data_ptr()]
        else:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=self.pg)
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(raw.data_ptr())
                else:
                    p = self.C.symm_open(h, self.device.index or 0)
                    self._peer_handles.append(p)
                    ptrs.append(p)
            dist.barrier(group=self.pg)
        return SymmBuffer(raw, ptrs, raw)

    def alloc(self, numel: int, dtype: torch.dtype) -> SymmBuffer:
        esize = torch.empty((), dtype=dtype).element_size()
        b = self.alloc_bytes(numel * esize)
        b.local = b._raw[: numel * esize].view(dtype)
        return b

    def allocator(self, registry: dict):
        """An ``alloc(n, dtype) -> tensor`` callable for ``flat.build_groups`` that records the
        SymmBuffer of each returned tensor in ``registry[data_ptr]``."""

        def alloc(n, dtype):
            b = self.alloc(n, dtype)
            registry[b.local.data_ptr()] = b
            return b.local

        return alloc

    # -- collectives ------------------------------------------------------------------------------
    def _epochs(self, n: int = 2) -> int:
        e = self.epoch + 1
        self.epoch += n
        return e

    def allreduce_scale_(self, buf: SymmBuffer, elem_off: int, n: int, scale: float, blocks: Optional[int] = None):
        if self.nvls and buf.mc_ptr:
            self.C.comm_nvls_allreduce_scale(buf.mc_ptr, self.pad_ptrs, elem_off, n, scale, self.rank, self._epochs(2),
                                             self.err, blocks or self.comm_blocks)
            return
        self.C.comm_allreduce_scale(buf.ptrs, self.pad_ptrs, elem_off, n, scale, self.rank, self._epochs(2), self.err,
                                    blocks or self.comm_blocks)

    def rs_adamw_(self, grads: SymmBuffer, params: Optional[SymmBuffer], param_local, m, v, push_params: bool,
                  elem_off: int, n: int, hyper, step: int, grad_scale: float, blocks: Optional[int] = None):
        lr, b1, b2, eps, wd = hyper
        if self.nvls and push_params and grads.mc_ptr and params is not None and params.mc_ptr:
            self.C.comm_nvls_rs_adamw(grads.mc_ptr, params.mc_ptr, params.ptrs[self.rank], m, v, self.pad_ptrs, elem_off,
                                      n, lr, b1, b2, eps, wd, step, grad_scale, self.rank, self._epochs(2), self.err,
                                      blocks or self.comm_blocks)
            return
        self.C.comm_rs_adamw(grads.ptrs, params.ptrs if params is not None else [], param_local, m, v, push_params,
                             self.pad_ptrs, elem_off, n, lr, b1, b2, eps, wd, step, grad_scale, self.rank,
                             self._epochs(2), self.err, blocks or self.comm_blocks)

    def allgather_(self, shards: SymmBuffer, full: torch.Tensor, shard_off: int, per: int, barrier: bool = True,
                   blocks: Optional[int] = None, copy_engine: bool = False):
        """``full[r*per:(r+1)*per] = shard of rank r``.  ``copy_engine=True``: a 1-warp barrier kernel + N async
        peer copies (zero SM time: the right choice for a prefetch that runs under GEMMs)."""
        self.C.comm_allgather(shards.ptrs, full, self.pad_ptrs, shard_off, per, self.rank, self._epochs(1), self.err,
                              barrier, 0 if copy_engine else (blocks or self.comm_blocks))

    def reduce_scatter_(self, grads: SymmBuffer, out: torch.Tensor, elem_off: int, n: int, scale: float,
                        blocks: Optional[int] = None):
        self.C.comm_reduce_scatter(grads.ptrs, out, self.pad_ptrs, elem_off, n, scale, self.rank, self._epochs(2),
                                   self.err, blocks or self.comm_blocks)

    def barrier_(self):
        self.C.comm_barrier(self.pad_ptrs, self.rank, self._epochs(1), self.err)

    def check(self):
        """Raise if a device-side barrier timed out (a peer died or diverged)."""
        v = int(self.err.item())
        if v:
            raise RuntimeError(f"NVLink barrier timed out waiting for rank {v - 1} (group rank {self.rank}); "
                               + self.describe_pads())

    def describe_pads(self, timeout_s: float = 5.0) -> str:
        """Signal-pad state for a post-mortem: per peer, the range of epochs this rank has received over its
        channels, next to the epoch this rank's host has issued.  A peer stuck at a lower epoch never launched
        (or never finished) the matching collective.  The read runs on its own stream and gives up after
        ``timeout_s`` so it cannot itself hang behind a spinning kernel."""
        import time

        try:
            ch = int(self.C.SYMM_MAX_CHANNELS)
            mr = int(self.C.SYMM_PAD_BYTES) // 4 // ch           # uint32 [channels][max ranks]
            words = self.pads.local[: ch * mr * 4].view(torch.int32)
            host = torch.empty(ch * mr + 1, dtype=torch.int32, pin_memory=True)
            side = torch.cuda.Stream(device=self.device)
            done = torch.cuda.Event()
            with torch.cuda.stream(side):
                host[: ch * mr].copy_(words, non_blocking=True)
                host[ch * mr:].copy_(self.err, non_blocking=True)
                done.record(side)
            t0 = time.time()
            while not done.query():
                if time.time() - t0 > timeout_s:
                    return f"(pad read did not complete within {timeout_s:.0f} s; issued locally: {self.epoch})"
                time.sleep(0.01)
            tab = host[: ch * mr].view(ch, mr)[: self.comm_blocks, : self.world]
            seen = ", ".join(f"rank {p}: {int(tab[:, p].min())}..{int(tab[:, p].max())}" for p in range(self.world))
            return (f"epochs received per peer (min..max over {self.comm_blocks} channels): {seen}; "
                    f"issued locally: {self.epoch}; error flag: {int(host[ch * mr])}")
        except Exception as e:  # pragma: no cover - best effort diagnostics
            return f"(pad state unavailable: {e})"

    def close(self):
        for p in self._peer_handles:
            try:
                self.C.symm_close(p)
            except Exception:
                pass
        self._peer_handles = []
