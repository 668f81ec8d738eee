"""Spawn helpers for multi-process CPU (gloo) tests."""
import os
import socket
import traceback

import torch
import torch.multiprocessing as mp


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _to_plain(x):
    """Tensors -> numpy so results survive the sender exiting (torch's shared-memory handles do not)."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().float().numpy()
    if isinstance(x, dict):
        return {k: _to_plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_plain(v) for v in x)
    return x


def _worker(rank, world, port, fn, args, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    try:
        out = fn(rank, world, *args)
        q.put((rank, "ok", _to_plain(out)))
    except Exception:
        q.put((rank, "err", traceback.format_exc()))
    finally:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


def run_distributed(fn, world=2, args=(), timeout=300):
    """Runs ``fn(rank, world, *args)`` in ``world`` processes; returns [result of rank 0, rank 1, ...]."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, status, out = q.get(timeout=timeout)
            if status != "ok":
                raise AssertionError(f"rank {rank} failed:\n{out}")
            results[rank] = out
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    return [results[r] for r in range(world)]


def update_rel_err(init, got, ref):
    """Relative L2 error of the *update* (final - initial weights) of ``got`` against ``ref``.  Max-abs differences of
    the weights themselves are dominated by sign flips of near-zero gradients under Adam (each costs 2*lr) and say
    little; a wrong number of optimizer steps or a wrong accumulation moves this metric to ~1, a correct engine sits
    at a few percent (bf16 rounding)."""
    import numpy as np

    keys = sorted(ref)
    f = lambda d: np.concatenate([np.asarray(d[k], dtype=np.float64).ravel() for k in keys])  # noqa: E731
    i, g, r = f(init), f(got), f(ref)
    return float(np.linalg.norm((g - i) - (r - i)) / max(np.linalg.norm(r - i), 1e-30))


def initial_weights(model_name="debug-llama", **kw):
    """The seeded initial weights every engine starts from (init is a function of seed + parameter name)."""
    import torch

    from distributed_training_guide_b200.engine import TrainEngine

    torch.manual_seed(0)
    eng = TrainEngine.create(model_name, parallelism="single", batch_size=1, seq_length=32, device="cpu", lr=1e-3, **kw)
    return {k: v.detach().float().clone().numpy() for k, v in eng.model.state_dict().items()}
