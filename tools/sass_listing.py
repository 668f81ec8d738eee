"""Per-kernel SASS evidence under profiles/sass/: for every named hot kernel of the extension, the mnemonic counts that
identify its Blackwell-native paths and an excerpt of the instructions themselves (cuobjdump -sass of the shipped
distributed_training_guide_b200/_C.so; runs without a GPU).

    python tools/sass_listing.py
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "distributed_training_guide_b200", "_C.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KEY = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "SYNCS", "LDGMC",
       "STGMC", "REDG", "RED.E", "ATOMG", "MUFU.EX2", "HMMA", "LDG.E.128", "STG.E.128", "LD.E", "ST.E", "MEMBAR", "FENCE",
       "CCTL", "ERRBAR", "NANOSLEEP"]
# (file stem, regex on the demangled name, what to look for)
KERNELS = [
    ("gemm_plain_2cta", r"gemm_bf16_kernel<true, true, 2, 0, 0, 0>", "tcgen05.mma cta_group::2 (UTCHMMA.2CTA), TMA loads, TMEM loads"),
    ("gemm_allgather_commcta", r"gemm_bf16_kernel<true, true, 2, 3, 0, 0>", "A_MODE 3: UBLKCP bulk copies peer -> smem -> local + flags, inside the GEMM"),
    ("gemm_fsdp_gather", r"gemm_bf16_kernel<true, true, 2, 0, 3, 0>", "B_MODE 3: gather warp (UBLKCP), chunk counters (RED / LD.ACQUIRE), inside the GEMM"),
    ("gemm_fsdp_gather_dgrad", r"gemm_bf16_kernel<true, false, 2, 0, 3, 0>", "B_MODE 3, dgrad form"),
    ("gemm_reduce_scatter_push", r"gemm_bf16_kernel<true, true, 2, 0, 0, 1>", "C_MODE 1: epilogue stores rows into the owner's staging slot (peer STG)"),
    ("gemm_wgrad_kgather", r"gemm_bf16_kernel<false, false, 2, 0, 2, 0>", "K-gathered wgrad: per-peer TMA descriptors"),
    ("attn_fwd_v1", r"attn_fwd_kernel", "SS-form QK and PV, P through shared memory"),
    ("attn_fwd_v2", r"attn_fwd2_kernel", "two query tiles, TS-form PV (A from TMEM), STTM of P, polynomial exp2"),
    ("attn_bwd_kv_ts", r"attn_bwd_kernel<true, true>", "KV pass, P^T/dS^T in TMEM (TS-form gradient MMAs)"),
    ("attn_bwd_q_ts", r"attn_bwd_kernel<false, true>", "Q pass, dS in TMEM"),
    ("attn_bwd_kv_ss", r"attn_bwd_kernel<true, false>", "KV pass, P^T/dS^T through shared memory"),
    ("rmsnorm_fwd", r"rmsnorm_fwd_kernel", "128-bit loads/stores, fused residual add"),
    ("rmsnorm_bwd", r"rmsnorm_bwd_kernel", ""),
    ("rope_inplace", r"rope_inplace_kernel", ""),
    ("swiglu_fwd", r"swiglu_fwd_kernel", ""),
    ("swiglu_bwd", r"swiglu_bwd_kernel", ""),
    ("cross_entropy", r"ce_row_kernel", "one CTA per row: online max/sum, dlogits written in place over the logits"),
    ("embedding_fwd", r"embedding_fwd_kernel", ""),
    ("adamw_flat", r"adamw_flat_kernel", ""),
    ("embedding_bwd_sorted", r"embedding_bwd_sorted_kernel", "deterministic, no atomics"),
    ("zero1_bucket_nr8", r"rs_adamw_kernel<8, __nv_bfloat16, true>", "peer LD (pull) + AdamW + peer ST (push), device barrier"),
    ("fsdp_bucket_nr8", r"rs_adamw_kernel<8, __nv_bfloat16, false>", ""),
    ("ddp_allreduce_nr8", r"allreduce_scale_kernel<8>", ""),
    ("nvls_zero1_bucket", r"nvls_rs_adamw_kernel<__nv_bfloat16, true>", "LDGMC ...ADD.BF16x8 = multimem.ld_reduce, multicast store = multimem.st"),
    ("nvls_fsdp_bucket", r"nvls_rs_adamw_kernel<__nv_bfloat16, false>", ""),
    ("nvls_allreduce", r"nvls_allreduce_scale_kernel", ""),
    ("tp_reduce_mc", r"tp_reduce_mc_kernel", "GEMM -> reduce-scatter, reduce half in the switch (LDGMC)"),
    ("vocab_parallel_ce", r"vp_ce_grad_kernel", "peer loads of the ranks' softmax statistics"),
    ("tp_embed_fwd", r"tp_embed_fwd_kernel", "hidden-parallel embedding: lookup pushed to the owner (all-to-all fused)"),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).strip())
    names = list(funcs)
    nice = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    index = ["# SASS listings per kernel (`cuobjdump -sass distributed_training_guide_b200/_C.so`, sm_100a)", "",
             "`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = TMA tensor load,",
             "`UBLKCP` = cp.async.bulk, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier, `LDGMC…ADD` = multimem.ld_reduce,",
             "`HMMA` would be the legacy mma.sync path (absent).", "", "| file | kernel | instructions | key mnemonics |", "|---|---|---|---|"]
    for stem, pat, note in KERNELS:
        hits = [(m, n) for m, n in zip(names, nice) if re.search(pat, n)]
        if not hits:
            continue
        mangled, pretty = hits[0]
        body = funcs[mangled]
        counts = {k: sum(1 for l in body if re.search(r"(?<![A-Z])" + re.escape(k), l)) for k in KEY}
        counts = {k: v for k, v in counts.items() if v}
        md = [f"# {pretty.split('(')[0]}", "", f"{note}" if note else "", "",
              f"{len(body)} SASS instructions; mnemonic counts: " + ", ".join(f"`{k}` {v}" for k, v in counts.items()), "",
              "```"]
        shown, last = 0, -10
        for i, l in enumerate(body):
            if any(re.search(r"(?<![A-Z])" + re.escape(k), l) for k in KEY[:13]) and shown < 60:
                if i - last > 2:
                    md.append("    ...")
                for j in range(max(last + 1, i - 1), min(len(body), i + 2)):
                    md.append(body[j])
                last = min(len(body), i + 2) - 1
                shown += 1
        md += ["```", ""]
        with open(os.path.join(OUT, stem + ".md"), "w") as fp:
            fp.write("\n".join(md))
        index.append(f"| [{stem}.md]({stem}.md) | `{pretty.split('(')[0][:80]}` | {len(body)} | "
                     + ", ".join(f"{k} {v}" for k, v in list(counts.items())[:8]) + " |")
    with open(os.path.join(OUT, "README.md"), "w") as fp:
        fp.write("\n".join(index) + "\n")
    print(f"wrote {len(index) - 8} listings to {OUT}")


if __name__ == "__main__":
    main()
