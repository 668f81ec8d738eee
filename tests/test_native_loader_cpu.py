"""C++ token loader: covers the file exactly once per epoch across ranks, reshuffles per epoch."""
import numpy as np
import pytest
import torch

from distributed_training_guide_b200 import _ext

pytestmark = pytest.mark.skipif(not _ext.available(), reason="extension not built")


def test_token_loader_partition_and_epochs(tmp_path):
    C = _ext.load(True)
    S, B, world = 16, 2, 2
    n_chunks = 20
    toks = (np.arange(n_chunks * S) % 60000).astype(np.uint16)
    path = tmp_path / "toks.bin"
    toks.tofile(path)
    firsts = {0: [], 1: []}
    for ep in (0, 1):
        seen = []
        for r in range(world):
            ld = C.TokenLoader(str(path), 2, S, B, r, world, 7, 3, False)
            ld.set_epoch(ep)
            assert ld.num_batches() == n_chunks // world // B
            for _ in range(ld.num_batches()):
                b = ld.next().clone()
                assert b.shape == (B, S) and b.dtype == torch.int64
                for row in b:
                    assert int(row[0]) % S == 0 and torch.equal(row, row[0] + torch.arange(S))  # a whole chunk
                    seen.append(int(row[0]) // S)
            firsts[ep].append(seen[-1])
        assert len(seen) == len(set(seen)) == (n_chunks // world // B) * B * world  # disjoint across ranks
    assert firsts[0] != firsts[1] or True  # permutations differ between epochs (checked below)
    a = C.TokenLoader(str(path), 2, S, B, 0, 1, 7, 3, False)
    e0 = [int(a.next()[0, 0]) for _ in range(3)]
    a.set_epoch(1)
    e1 = [int(a.next()[0, 0]) for _ in range(3)]
    assert e0 != e1


def test_trainer_uses_native_loader(tmp_path):
    from types import SimpleNamespace

    from distributed_training_guide_b200.models import get_config
    from distributed_training_guide_b200.utils import data as D

    cfg = get_config("debug-llama")
    toks = np.random.randint(0, cfg.vocab_size, size=64 * 40).astype(np.uint16)
    path = tmp_path / "c.bin"
    toks.tofile(path)
    args = SimpleNamespace(dataset_name=str(path), dataset_subset=None, model_name="debug-llama", seq_length=64, seed=0,
                           batch_size=4)
    ds = D.load_and_preprocess_data(args, cfg)
    dl = D.build_dataloader(ds, 4, seed=0)
    assert isinstance(dl, D.NativeTokenLoader) and len(dl) == 10
    batch = next(iter(dl))
    assert batch["input_ids"].shape == (4, 64) and batch["input_ids"].max() < cfg.vocab_size


def test_token_loader_restarts_and_teardown(tmp_path):
    """Abandoning an epoch half way (set_epoch while the producer thread is ahead) and destroying a loader whose
    producer is mid-epoch must neither deadlock nor crash."""
    C = _ext.load(True)
    S, B = 64, 4
    toks = (np.arange(S * 997) % 50000).astype(np.uint16)
    path = tmp_path / "stress.bin"
    toks.tofile(path)
    ld = C.TokenLoader(str(path), 2, S, B, 0, 2, 3, 4, False)
    for ep in range(60):
        ld.set_epoch(ep)
        n = ld.num_batches()
        for _ in range(n if ep % 3 else n // 2):
            b = ld.next()
            assert b.shape == (B, S)
    ld2 = C.TokenLoader(str(path), 2, S, B, 1, 2, 3, 4, False)
    ld2.next()
    del ld2
