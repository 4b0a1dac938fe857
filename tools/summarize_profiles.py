#!/usr/bin/env python
"""Turns the raw ncu outputs under gpurun_out/ into the small, tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <round tag, e.g. r01>"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def launches(path_csv, out_md, steps_hint):
    rows = list(csv.reader(open(path_csv, errors="replace")))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    idx = {h: i for i, h in enumerate(hdr)}
    seq = []
    for r in rows[start:]:
        if len(r) < len(hdr):
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1000 if u.startswith("n") else (v * 1000 if u.startswith("m") else v)
        seq.append((r[idx["Kernel Name"]], v))
    tot, cnt = collections.Counter(), collections.Counter()
    for n, v in seq:
        k = n.split("(")[0].replace("void ", "").replace("dino::", "")
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    with open(out_md, "w") as f:
        f.write(f"# ncu launch list summary ({tag})\n\n")
        f.write("Command: `DINO_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs` (tools/r02_profile.sh)\n")
        f.write(f"(ViT-L/14, 518x518, batch 64, classify; {len(seq)} launches captured = {steps_hint}). Per-launch times are serialised and\n"
                "cold-cache under ncu: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share | mean us |\n|---|---:|---:|---:|---:|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k}` | {cnt[k]} | {v / 1000:.2f} | {100 * v / total:.1f} % | {v / cnt[k]:.1f} |\n")
        f.write(f"| **all** | {len(seq)} | {total / 1000:.2f} | 100 % | |\n")
        f.write("\nTemplate arguments of `gemm_f16_tcgen05<BN, EPI>`: EPI 0 = qkv (+bias -> fp16), 1 = fc1 (+bias, GELU -> fp16), 2 = o-proj / fc2\n"
                "(+bias, LayerScale, residual TMA reduce-add), 4 = patch embedding (+bias +pos-embed, token scatter).\n")


def raw_metrics(rep, out_md, title, want):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out_md, "w") as f:
        f.write(f"# {title} ({tag})\n\n`ncu --set full --clock-control none --import-source on` (one GPU, ViT-L/14 518^2 batch 64 step via tools/profile_step.py).\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']]}  (grid {r[idx['launch__grid_size']]}, block {r[idx['launch__block_size']]})\n\n| metric | value |\n|---|---|\n")
            for w in want:
                if w in idx:
                    f.write(f"| {w} | {r[idx[w]]} {units[idx[w]]} |\n")
            f.write("\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

if os.path.exists(os.path.join(G, f"{tag}_launches_bench.csv")):
    launches(os.path.join(G, f"{tag}_launches_bench.csv"), os.path.join(P, f"{tag}_launches_bench.md"), "about 10 forward passes (warm-up, timed, end-to-end) + weight upload; capture capped by -c")
for name, title in (("gemm", "GEMM kernels (gemm_f16_tcgen05): qkv, o-proj, fc1, fc2 of one block"), ("attn", "Attention kernel"),
                    ("ln", "LayerNorm kernel (norm1 / norm2 of one block)")):
    rep = os.path.join(G, f"{tag}_prof_{name}.ncu-rep")
    if os.path.exists(rep):
        raw_metrics(rep, os.path.join(P, f"{tag}_ncu_{name}.md"), title, WANT)
print(os.listdir(P))
