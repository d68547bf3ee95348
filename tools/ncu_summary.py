#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into profiles/.

    python tools/ncu_summary.py gpurun_out/r7_full.ncu-rep profiles/r1_s2_full [--traffic]

Writes <out>.md (per-kernel table: duration, DRAM bytes, throughput percentages, shared-memory
wavefronts, occupancy, the stall-reason histogram and the hottest SASS lines) and, with
--traffic, updates profiles/traffic.json (dram bytes per launch per kernel tag, which
bench.py reports as roofline.traffic).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts (LSU)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("smsp__cycles_active.avg", "SMSP cycles active"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]

TAGS = [
    (r"k_pipe_tcg<(0|false), 0>", "pipe_gather_fwd"),
    (r"k_pipe_tcg<(0|false), 2>", "pipe_gather_fwd_mse"),
    (r"k_pipe_tcg<(1|true), 1>", "pipe_gather_bwd"),
    (r"k_pipe_gather<64, 64, (0|false), 0>", "pipe_gather_fwd"),
    (r"k_pipe_gather<64, 64, (0|false), 2>", "pipe_gather_fwd_mse"),
    (r"k_pipe_gather<64, 64, (1|true), 1>", "pipe_gather_bwd"),
    (r"k_pipe_tn_reduce", "pipe_tn_reduce"),
    (r"k_pipe_tn<", "pipe_tn"),
    (r"k_mse_graph", "mse_graph"),
    (r"k_aggregate<", "aggregate"),
    (r"k_tc_rows", "tc_rows"),
    (r"k_step", "optimiser_step"),
    (r"k_finalize", "finalize_step"),
    (r"k_agg_tc128", "agg_tc_fwd"),
]


def tag_of(name):
    plain = re.sub(r"\((int|bool|long long|unsigned int)\)", "", name)
    for pat, tag in TAGS:
        if re.search(pat, plain):
            return tag
    return re.sub(r"\(.*", "", plain).split("::")[-1].strip()


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v * mult


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit.lower(), 1.0)


def ncu(args):
    return subprocess.run(["ncu"] + args, check=True, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    want_traffic = "--traffic" in sys.argv
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, body = raw[0], raw[1], raw[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary of `{os.path.basename(rep)}`", "",
             "Read with `ncu -i <rep> --page raw --csv` / `--page source --csv` "
             "(tools/ncu_summary.py).  Durations under ncu are cold-cache and serialised: "
             "compare shares and byte counts, not absolute times.", ""]
    traffic = {}
    for k, r in enumerate(body):
        name = r[col["Kernel Name"]]
        tag = tag_of(name)
        lines += [f"## launch {k}: `{tag}`", "", f"`{name[:160]}`", "", "| metric | value |", "|---|---|"]
        rd = wr = dur = None
        seen = set()
        for key, label in RAW_KEYS:
            if key not in col or label in seen:
                continue
            v, u = r[col[key]], units[col[key]]
            if v == "":
                continue
            seen.add(label)
            lines.append(f"| {label} (`{key}`) | {v} {u} |")
            if key == "dram__bytes_read.sum":
                rd = to_bytes(v, u)
            if key == "dram__bytes_write.sum":
                wr = to_bytes(v, u)
            if key == "gpu__time_duration.sum":
                dur = to_us(v, u)
        if rd is not None and wr is not None and dur:
            gbps = (rd + wr) / (dur * 1e-6) / 1e9
            lines.append(f"| **DRAM traffic / launch** | {(rd + wr) / 1e6:.1f} MB -> {gbps:.0f} GB/s "
                         f"under ncu ({dur:.1f} us) |")
            traffic.setdefault(tag, []).append(rd + wr)
        lines.append("")
    # source page: stall histogram + hottest instructions per kernel
    try:
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    except subprocess.CalledProcessError:
        src = []
    starts = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    for n, s0 in enumerate(starts):
        end = (starts[n + 1] - 1) if n + 1 < len(starts) else len(src)
        kname = src[s0 - 1][1] if s0 > 0 and len(src[s0 - 1]) > 1 else f"kernel {n}"
        h = src[s0]
        rows = [r for r in src[s0 + 1:end] if len(r) == len(h)]
        si, so = h.index("# Samples"), h.index("Source")
        stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        agg = {h[i]: 0 for i in stall_cols}
        tot = 0
        for r in rows:
            if r[si].isdigit():
                tot += int(r[si])
            for i in stall_cols:
                if r[i].isdigit():
                    agg[h[i]] += int(r[i])
        lines += [f"## stall sampling, launch {n}: `{tag_of(kname)}`", "",
                  f"{tot} warp samples over {len(rows)} SASS instructions.", "",
                  "| stall reason | samples | share |", "|---|---|---|"]
        for name, c in sorted(agg.items(), key=lambda x: -x[1])[:8]:
            if c:
                lines.append(f"| {name} | {c} | {100.0 * c / max(tot, 1):.1f} % |")
        lines += ["", "Hottest instructions:", "", "| # | samples | SASS | top stalls |", "|---|---|---|---|"]
        top = sorted(((int(r[si]), i) for i, r in enumerate(rows) if r[si].isdigit()), reverse=True)[:14]
        for c, i in sorted(top, key=lambda x: x[1]):
            r = rows[i]
            st = sorted(((int(r[j]), h[j]) for j in stall_cols if r[j].isdigit() and int(r[j]) > 0),
                        reverse=True)[:2]
            lines.append(f"| {i} | {c} | `{r[so].strip()[:60]}` | "
                         + ", ".join(f"{nm} {v}" for v, nm in st) + " |")
        lines.append("")
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    open(out + ".md", "w").write("\n".join(lines) + "\n")
    print("wrote", out + ".md")
    if want_traffic:
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        cur = json.load(open(tp)) if os.path.exists(tp) else {}
        for tag, vals in traffic.items():
            cur[tag] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
                        "source": os.path.basename(out) + ".md"}
        json.dump(cur, open(tp, "w"), indent=1, sort_keys=True)
        print("updated", tp)


if __name__ == "__main__":
    main()
