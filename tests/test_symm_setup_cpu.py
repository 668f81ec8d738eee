"""Symmetric-arena set-up protocol without a GPU: the store rendezvous and the SCM_RIGHTS exchange of allocation
file descriptors (``parallel/symm.py: _StoreComm, _VmmChunk``) run in real processes against a stand-in for the
CUDA VMM chunk whose 'allocations' are memfds carrying the owner's rank."""
import os

from dist_utils import run_distributed


class _FakeVmm:
    def __init__(self, device, nbytes, world, rank, want_mc):
        self.world, self.rank, self.want_mc = world, rank, want_mc
        self.nbytes = (nbytes + 4095) // 4096 * 4096
        self.seen, self.mc_seen, self.calls = {}, None, []

    def size(self):
        return self.nbytes

    def _fd(self, text):
        fd = os.memfd_create("fake-vmm")
        os.write(fd, text.encode())
        return fd

    def export_fd(self):
        return self._fd(f"mem-of-{self.rank}")

    def mc_create_export(self):
        assert self.rank == 0
        self.mc_seen = "mc-of-0"
        return self._fd("mc-of-0")

    def _read(self, fd):
        return os.pread(fd, 64, 0).decode()  # (the duplicates share one file offset: positional read)

    def import_peer(self, peer, fd):
        self.seen[peer] = self._read(fd)

    def mc_import(self, fd):
        self.mc_seen = self._read(fd)

    def map_all(self):
        self.calls.append("map")

    def mc_add_device(self):
        self.calls.append("add")

    def mc_bind_and_map(self):
        self.calls.append("bind")

    def base(self):
        return 1 << 40

    def mc_base(self):
        return 1 << 41

    def release(self):
        pass


def _setup(rank, world, multicast):
    import types

    import torch
    import torch.distributed as dist

    from distributed_training_guide_b200.parallel import symm

    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = types.SimpleNamespace(C=types.SimpleNamespace(VmmChunk=_FakeVmm), comm=symm._StoreComm(None), world=world,
                              rank=rank, multicast=multicast, device=torch.device("cpu"), _n_chunks=0,
                              token="t%d" % os.getppid())
    out = []
    for i in range(2):                     # two chunks in a row: socket names and store keys must not collide
        g._n_chunks = i
        ch = symm._VmmChunk(g, 10000)
        assert ch.size == 12288
        assert ch.chunk.seen == {p: f"mem-of-{p}" for p in range(world) if p != rank}, ch.chunk.seen
        if multicast:
            assert ch.chunk.mc_seen == "mc-of-0" and ch.chunk.calls == ["map", "add", "bind"] and ch.mc_base == 1 << 41
        else:
            assert ch.chunk.mc_seen is None and ch.chunk.calls == ["map"] and ch.mc_base == 0
        assert ch.bases == [(1 << 40) + r * 12288 for r in range(world)]
        out.append(sorted(ch.chunk.seen))
    assert g.comm.all_gather(rank * 10) == [r * 10 for r in range(world)]
    return out


def test_fd_exchange_and_store_rendezvous_four_ranks():
    for multicast in (False, True):
        res = run_distributed(_setup, world=4, args=(multicast,), timeout=120)
        assert len(res) == 4
