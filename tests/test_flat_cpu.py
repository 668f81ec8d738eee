"""Flat parameter groups (parallel/flat.py): layout invariants the kernels rely on."""
import pytest
import torch

from distributed_training_guide_b200.models import build_model, get_config
from distributed_training_guide_b200.parallel.flat import ALIGN, build_groups


@pytest.mark.parametrize("model_name,world", [("debug-llama", 1), ("debug-llama-gqa", 2), ("debug-llama", 8),
                                              ("debug-gpt2", 4)])
def test_layout_invariants(model_name, world):
    cfg = get_config(model_name)
    torch.manual_seed(0)
    model = build_model(cfg, dtype=torch.float32, device="cpu")
    before = {k: v.clone() for k, v in model.state_dict().items()}
    groups = build_groups(model, "cpu", torch.float32, world_size=world)
    names = [g.name for g in groups]
    assert names[0] == "embed" and names[-1] in ("head", f"layer{cfg.num_hidden_layers - 1}")
    covered = set()
    for g in groups:
        # every view is 16-byte aligned (TMA / 128-bit loads), nothing overlaps, shards are equal and aligned
        ends = []
        for n, p, off, shape in zip(g.names, g.params, g.offsets, g.shapes):
            assert off % ALIGN == 0 and tuple(p.shape) == shape
            assert p.data_ptr() == g.param.data_ptr() + off * g.param.element_size()      # the parameter IS the view
            assert p.grad is not None and p.grad.data_ptr() == g.grad.data_ptr() + off * g.grad.element_size()
            ends.append((off, off + p.numel()))
            covered.add(n)
        ends.sort()
        assert all(a[1] <= b[0] for a, b in zip(ends, ends[1:]))
        assert g.padded_numel >= g.numel and g.padded_numel % (ALIGN * world) == 0
        lo_hi = [g.shard_range(r, world) for r in range(world)]
        assert lo_hi[0][0] == 0 and lo_hi[-1][1] == g.padded_numel
        assert all(a[1] == b[0] and (a[1] - a[0]) % ALIGN == 0 for a, b in zip(lo_hi, lo_hi[1:]))
    assert covered == {n for n, _ in model.named_parameters()}
    # values survived the move into the flat buffers
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k


def test_fused_projections_are_adjacent_views():
    cfg = get_config("debug-llama-gqa")
    model = build_model(cfg, dtype=torch.float32, device="cpu")
    build_groups(model, "cpu", torch.float32, world_size=2)
    layer = model.model.layers[0]
    att, mlp = layer.self_attn, layer.mlp
    qkv = layer._fused["qkv"].data
    gu = layer._fused["gate_up"].data
    assert qkv.shape[0] == att.q_proj.weight.shape[0] + att.k_proj.weight.shape[0] + att.v_proj.weight.shape[0]
    assert qkv.data_ptr() == att.q_proj.weight.data_ptr() and qkv.is_contiguous()
    assert torch.equal(qkv, torch.cat([att.q_proj.weight, att.k_proj.weight, att.v_proj.weight]))
    assert gu.data_ptr() == mlp.gate_proj.weight.data_ptr()
    assert torch.equal(gu, torch.cat([mlp.gate_proj.weight, mlp.up_proj.weight]))
    # writing through the fused gradient view lands in the per-parameter gradients
    layer._fused["qkv"]._dtg_grad.fill_(3.0)
    assert float(att.k_proj.weight.grad.min()) == 3.0 and float(att.v_proj.weight.grad.max()) == 3.0
