"""The GPU branches of the FSDP / DDP engines (streams, events, fused-kernel calls) executed on the CPU with the CUDA
runtime and the NVLink kernels replaced by functional stand-ins: catches logic and bookkeeping errors in code that the
gloo tests (torch.distributed fallback path) never reach.  Single process (world = 1)."""
import contextlib
from types import SimpleNamespace
from unittest import mock

import numpy as np
import pytest
import torch

from dist_utils import update_rel_err


class _FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 0.0


class _FakeStream:
    def wait_event(self, ev):
        pass

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass


class _FakeC:
    """Stand-ins for the extension entry points the engines call directly."""

    def adamw_flat(self, p, g, m, v, lr, b1, b2, eps, wd, step, grad_scale):
        from distributed_training_guide_b200.ops import reference as ref

        ref.adamw_step(p, g, m, v, lr, b1, b2, eps, wd, step, grad_scale)


class _FakeSymm:
    """world = 1: the 'collectives' act on the local buffers; semantics of parallel/symm.py."""

    def __init__(self):
        self.C = _FakeC()
        self.device = torch.device("cpu")
        self.calls = []

    def allgather_(self, shards, full, shard_off, per, barrier=True, blocks=None, copy_engine=False):
        self.calls.append("allgather")
        full[:per].copy_(shards[shard_off:shard_off + per])

    def reduce_scatter_(self, grads, out, elem_off, n, scale, blocks=None):
        self.calls.append("reduce_scatter")
        out.copy_((grads[elem_off:elem_off + n].float() * scale).to(out.dtype))

    def rs_adamw_(self, grads, params, param_local, m, v, push_params, elem_off, n, hyper, step, grad_scale, blocks=None):
        from distributed_training_guide_b200.ops import reference as ref

        self.calls.append("rs_adamw")
        lr, b1, b2, eps, wd = hyper
        target = params[elem_off:elem_off + n] if push_params else param_local
        ref.adamw_step(target, grads[elem_off:elem_off + n], m, v, lr, b1, b2, eps, wd, step, grad_scale)

    def check(self):
        pass


def _patches():
    return [mock.patch("torch.cuda.Event", _FakeEvent), mock.patch("torch.cuda.Stream", lambda *a, **k: _FakeStream()),
            mock.patch("torch.cuda.stream", lambda s: contextlib.nullcontext()),
            mock.patch("torch.cuda.current_stream", lambda *a, **k: _FakeStream()),
            mock.patch("torch.cuda.synchronize", lambda *a, **k: None)]


def _run(engine_kind, kernel_path, K, steps=3, model_name="debug-llama"):
    from distributed_training_guide_b200.engine import TrainEngine

    with contextlib.ExitStack() as es:
        for p in _patches():
            es.enter_context(p)
        torch.manual_seed(0)
        eng = TrainEngine.create(model_name, parallelism=engine_kind, batch_size=2, seq_length=32, device="cpu",
                                 lr=1e-3)
        eng.model.eval()
        e = eng.strategy.engine
        init = {k: v.detach().float().clone().numpy() for k, v in eng.model.state_dict().items()}
        fake = _FakeSymm()
        if kernel_path:
            e.use_kernels, e.symm = True, fake
            e.comm_stream, e._done = _FakeStream(), _FakeEvent()
            if engine_kind == "fsdp":
                e.ag_done, e.rs_done, e.slot_free = {}, {}, [None] * len(e.full_slots)
                e.ag_copy_engine = True
                e._symm_of = {t.data_ptr(): t for t in [s.param for s in e.shards] + list(e.grad_slots)
                              + [e.embed.grad, e.head.grad]}
            else:
                e.registry = {t.data_ptr(): t for g in e.groups for t in (g.param, g.grad)}
        s = eng.strategy
        for i in range(steps):
            for m in range(K):
                gen = torch.Generator().manual_seed(100 * i + m)
                ids = torch.randint(0, eng.config.vocab_size, (2, 32), generator=gen)
                s.pre_step(eng.model)
                out = eng.model(**s.prepare_batch({"input_ids": ids, "labels": ids.clone()}))
                with s.grad_sync(eng.model, enabled=(m == K - 1)):
                    s.backward(eng.model, out.loss / K)
            eng.optimizer.step()
            eng.lr_scheduler.step()
            eng.optimizer.zero_grad()
        sd = e.full_state_dict() if engine_kind == "fsdp" else eng.model.state_dict()
        return init, {k: v.detach().float().clone().numpy() for k, v in sd.items()}, fake.calls


@pytest.mark.parametrize("model_name", ["debug-llama", "debug-gpt2"])
def test_fsdp_kernel_branch_matches_fallback_branch(model_name):
    for K in (1, 2):
        init, want, _ = _run("fsdp", False, K, model_name=model_name)
        _, got, calls = _run("fsdp", True, K, model_name=model_name)
        assert "allgather" in calls and ("rs_adamw" in calls if K == 1 else "reduce_scatter" in calls), calls
        assert update_rel_err(init, got, want) < 0.05, K
        assert all(np.isfinite(v).all() for v in got.values())


@pytest.mark.parametrize("model_name", ["debug-llama", "debug-gpt2"])
def test_single_gpu_engine_kernel_branch_matches_plain_optimizer(model_name):
    """Chapter 01 on a GPU runs AdamW per bucket inside backward (the N = 1 case of the ZeRO-1 bucket kernel); GPT-2
    takes the same route with autograd-accumulated gradients and a tied lm_head."""
    from distributed_training_guide_b200.engine import TrainEngine
    from distributed_training_guide_b200.parallel.ddp import DataParallelEngine

    def run(kernel_path, K):
        with contextlib.ExitStack() as es:
            for p in _patches():
                es.enter_context(p)
            torch.manual_seed(0)
            eng = TrainEngine.create(model_name, parallelism="single", batch_size=2, seq_length=32, device="cpu", lr=1e-3)
            eng.model.eval()  # GPT-2: no dropout, the two runs must see the same function
            init = {k: v.detach().float().clone().numpy() for k, v in eng.model.state_dict().items()}
            fake = _FakeSymm()
            if kernel_path:
                st = eng.strategy
                reg = {t.data_ptr(): t for g in st.groups for t in (g.param, g.grad)}
                st.engine = DataParallelEngine(eng.model, st.groups, eng.optimizer, symm=fake, registry=reg, zero1=True,
                                               world_size=1, rank=0)
            s = eng.strategy
            for i in range(3):
                for m in range(K):
                    gen = torch.Generator().manual_seed(100 * i + m)
                    ids = torch.randint(0, eng.config.vocab_size, (2, 32), generator=gen)
                    out = eng.model(**s.prepare_batch({"input_ids": ids, "labels": ids.clone()}))
                    with s.grad_sync(eng.model, enabled=(m == K - 1)):
                        s.backward(eng.model, out.loss / K)
                eng.optimizer.step()
                eng.lr_scheduler.step()
                eng.optimizer.zero_grad()
            return init, {k: v.detach().float().clone().numpy() for k, v in eng.model.state_dict().items()}, fake.calls

    for K in (1, 2):
        init, want, _ = run(False, K)
        _, got, calls = run(True, K)
        assert calls.count("rs_adamw") == 3 * 4, calls  # 3 optimizer steps x buckets (embed, 2 layers, head)
        assert update_rel_err(init, got, want) < 0.05, K
