"""Native token loader on a GPU: the pinned ring must not be refilled before the asynchronous H2D copy out of a
slot has executed, even when the host runs far ahead of the device (slots were recycled by call count)."""
import numpy as np
import pytest
import torch

from distributed_training_guide_b200 import _ext
from distributed_training_guide_b200.utils import data as D

pytestmark = pytest.mark.gpu


def test_ring_slot_is_guarded_until_its_h2d_copy_ran(tmp_path):
    S, B, n_chunks = 512, 4, 64
    toks = (np.arange(n_chunks * S) % 50000).astype(np.uint16)
    path = tmp_path / "toks.bin"
    toks.tofile(path)
    dl = D.NativeTokenLoader(str(path), S, 50000, B, depth=2)   # the smallest ring: every slot is reused at once
    dev = torch.device("cuda:0")
    got = []
    for batch in dl:
        torch.cuda._sleep(int(2e8))            # ~0.1 s of device work in front of every copy: the host runs ahead
        got.append(D.to_device(batch, dev)["input_ids"])
    torch.cuda.synchronize()
    assert len(got) == n_chunks // B
    seen = []
    for b in got:
        b = b.cpu()
        for row in b:
            first = int(row[0])
            # every row is one whole, unmixed chunk of the file
            assert torch.equal(row, (first + torch.arange(S)) % 50000), "a ring slot was overwritten before its copy ran"
            seen.append(first)
    assert len(set(seen)) == len(seen) == n_chunks
