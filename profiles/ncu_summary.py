#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` report into markdown: headline metrics per captured launch, the
stall-reason mix and the instructions that collect the most warp-stall samples (needs -lineinfo; SASS view).

    python profiles/ncu_summary.py gpurun_out/gemm_r1i.ncu-rep profiles/ncu_gemm_r1i.md "title"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst, title):
    raw = ncu_csv(rep, "raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    md = [f"# {title}", "", f"source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`; times under the "
          "profiler are cold-cache and serialised -- never bench values)", ""]
    names = [r[hdr.index("Kernel Name")].split("(")[0][-60:] for r in rows]
    md += ["| metric | " + " | ".join(f"launch {i}" for i in range(len(rows))) + " | unit |", "|---|" + "---|" * (len(rows) + 1)]
    for m in METRICS:
        if m in hdr:
            k = hdr.index(m)
            md.append(f"| {m} | " + " | ".join(r[k] for r in rows) + f" | {units[k]} |")
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    md += ["", "kernels: " + "; ".join(f"launch {i} = `{n}`" for i, n in enumerate(names)), "",
           "## warps stalled per issued instruction, by reason", "",
           "| reason | " + " | ".join(f"launch {i}" for i in range(len(rows))) + " |", "|---|" + "---|" * len(rows)]
    for h in sorted(stall, key=lambda h: -float(rows[0][hdr.index(h)] or 0)):
        k = hdr.index(h)
        if max(float(r[k] or 0) for r in rows) < 0.05:
            continue
        short = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
        md.append(f"| {short} | " + " | ".join(f"{float(r[k] or 0):.2f}" for r in rows) + " |")
    src = ncu_csv(rep, "source")
    hi = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if hi:
        h = src[hi[0]]
        end = hi[1] - 1 if len(hi) > 1 else len(src)
        body = [r for r in src[hi[0] + 1:end] if len(r) == len(h)]
        iS, iSrc, iE = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[iS]) for r in body if r[iS].isdigit())
        md += ["", f"## instructions with the most warp-stall samples (launch 0, {tot} samples)", "",
               "| samples | executed | SASS | stall reasons |", "|---|---|---|---|"]
        for r in sorted(body, key=lambda r: -int(r[iS]) if r[iS].isdigit() else 0)[:14]:
            why = " ".join(f"{c[6:]}={r[h.index(c)]}" for c in cols if r[h.index(c)] not in ("0", ""))
            md.append(f"| {r[iS]} | {r[iE]} | `{r[iSrc].strip()[:70]}` | {why} |")
    with open(dst, "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md[:30]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
