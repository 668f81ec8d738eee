"""Pretrained-weight loading (reference chapter 05: rank-0 load + distribute): a Hugging Face style directory
(config.json + sharded safetensors + index) is loaded into the replicated engine and into 2-rank FSDP shards."""
import json

import pytest
import torch

from dist_utils import run_distributed


def _write_hf_dir(path, two_files=True):
    from safetensors.torch import save_file

    from distributed_training_guide_b200.models import build_model, get_config
    from distributed_training_guide_b200.models.configs import to_hf_config_dict

    cfg = get_config("debug-llama")
    torch.manual_seed(1234)
    model = build_model(cfg, dtype=torch.float32, device="cpu")
    sd = {k: torch.randn_like(v) * 0.05 for k, v in model.state_dict().items()}
    path.mkdir(parents=True, exist_ok=True)
    (path / "config.json").write_text(json.dumps(to_hf_config_dict(cfg)))
    keys = sorted(sd)
    half = len(keys) // 2 if two_files else len(keys)
    files = {"model-00001-of-00002.safetensors": keys[:half], "model-00002-of-00002.safetensors": keys[half:]}
    weight_map = {}
    for fname, ks in files.items():
        if ks:
            save_file({k: sd[k].contiguous() for k in ks}, str(path / fname))
            weight_map.update({k: fname for k in ks})
    (path / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": weight_map}))
    return sd


def test_replicated_load(tmp_path):
    from types import SimpleNamespace

    from distributed_training_guide_b200.models import build_model, get_config
    from distributed_training_guide_b200.tools.load_hf import find_checkpoint, maybe_load_pretrained

    sd = _write_hf_dir(tmp_path / "m")
    assert find_checkpoint(str(tmp_path / "m")) == str(tmp_path / "m")
    cfg = get_config(str(tmp_path / "m"))
    model = build_model(cfg, dtype=torch.float32, device="cpu")
    args = SimpleNamespace(model_name=str(tmp_path / "m"), pretrained="require")
    assert maybe_load_pretrained(args, model=model)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    with pytest.raises(FileNotFoundError):
        maybe_load_pretrained(SimpleNamespace(model_name=str(tmp_path), pretrained="require"), model=model)
    assert not maybe_load_pretrained(SimpleNamespace(model_name="debug-llama", pretrained="auto"), model=model)


def _fsdp_load(rank, world, model_dir):
    from types import SimpleNamespace

    from distributed_training_guide_b200.parallel.strategies import FullyShardedDataParallel
    from distributed_training_guide_b200.models import get_config

    args = SimpleNamespace(model_name=model_dir, pretrained=None, chapter="05-training-llama-405b", seed=0, device="cpu",
                           cpu_offload=False, checkpoint_activations=False, prefetch_layers=True, batch_size=1,
                           seq_length=32, local_rank=None)
    st = FullyShardedDataParallel(args)
    st.setup(args)
    model = st.build_model(args, get_config(model_dir))
    full = st.engine.full_state_dict()
    return {k: v.float().cpu() for k, v in full.items()} if rank == 0 else {}


def test_fsdp_load_distributes_rank0_checkpoint(tmp_path):
    sd = _write_hf_dir(tmp_path / "m")
    res = run_distributed(_fsdp_load, world=2, args=(str(tmp_path / "m"),), timeout=240)
    got = res[0]
    assert set(got) == set(sd)
    for k, v in sd.items():
        ref = v.to(torch.bfloat16).float()
        assert torch.allclose(torch.as_tensor(got[k]), ref, atol=0, rtol=0), k
