"""Pre-fetch what a run needs into $HF_HOME before launching many ranks (so 64 processes do not
hammer the hub / shared storage at start-up):

    python download.py -m meta-llama/Llama-3.1-405B [--skip-model] [--dataset Skylion007/openwebtext]

With --skip-model only config + tokenizer are fetched (random-init / architecture-only runs).
On a box without network this falls back to the embedded config registry and says so.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-m", "--model-name", default="meta-llama/Llama-3.1-405B")
    ap.add_argument("--skip-model", action="store_true")
    ap.add_argument("--dataset", default=None)
    args = ap.parse_args()
    try:
        from transformers import AutoConfig, AutoTokenizer

        AutoConfig.from_pretrained(args.model_name)
        AutoTokenizer.from_pretrained(args.model_name)
        if not args.skip_model:
            from huggingface_hub import snapshot_download

            snapshot_download(args.model_name, allow_patterns=["*.safetensors", "*.json"])
        if args.dataset:
            import datasets

            datasets.load_dataset(args.dataset)
        print(f"cached {args.model_name} under {os.environ.get('HF_HOME', '~/.cache/huggingface')}")
    except Exception as e:
        from distributed_training_guide_b200.models import get_config

        cfg = get_config(args.model_name)
        print(f"hub unreachable ({type(e).__name__}); embedded architecture config will be used: {cfg.num_parameters() / 1e9:.2f} B parameters")


if __name__ == "__main__":
    main()
