#!/usr/bin/env python
"""Turn an ncu launch list (csv, `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`)
into a per-family summary (markdown) + profiles/roofline_traffic.json used by bench.py's `roofline.traffic`.

    python profiles/summarize.py gpurun_out/launches_r1e.csv r1e
"""
import csv
import json
import os
import sys
from collections import OrderedDict

FAMILY = [("sepconv_mid", "sepconv_mid"), ("conv1_tc", "conv1"), ("pad_copy", "subsample"), ("sepconv2d_fused", "sepconv_fused"), ("sepconv_fused", "sepconv_fused"), ("gemm_tcgen05_2cta", "gemm_pointwise"), ("conv3x3_is", "gemm_conv2"), ("mc_head_fused", "head_fused"), ("gemm_tcgen05_kernel<32", "gemm_conv2"), ("gemm_tcgen05_kernel", "gemm_pointwise"),
          ("depthwise3x3", "depthwise"), ("maxpool_add", "maxpool_add"), ("subsample2", "subsample"), ("conv1_kernel", "conv1"),
          ("tile_stats", "tile_stats"), ("gap_kernel", "gap"), ("mc_expand", "mc_expand"), ("head_final", "head_final"),
          ("group_kahan", "threshold"), ("roc_", "threshold"), ("seg_bounds", "threshold"), ("group_apply", "threshold"),
          ("tile_process", "threshold"), ("validate_kernel", "threshold"), ("DeviceRadixSort", "threshold(cub)"),
          ("DeviceScan", "threshold(cub)")]


def fam(name):
    for key, f in FAMILY:
        if key in name:
            return f
    return "other(torch)"


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)


def main(path, tag):
    lines = [l for l in open(path) if not l.startswith("==")]
    byid = OrderedDict()
    for x in csv.DictReader(lines):
        d = byid.setdefault(x["ID"], {"name": x["Kernel Name"]})
        v, u = float(x["Metric Value"]), x["Metric Unit"]
        if x["Metric Name"].startswith("gpu__time"):
            d["us"] = to_us(v, u)
        elif "read" in x["Metric Name"]:
            d["rd"] = to_bytes(v, u)
        elif "write" in x["Metric Name"]:
            d["wr"] = to_bytes(v, u)
    L = list(byid.values())
    idx = [i for i, d in enumerate(L) if "tile_stats" in d["name"]]
    seq = L[idx[-2]:idx[-1]] if len(idx) > 1 else L[idx[-1]:]
    agg = OrderedDict()
    for d in seq:
        a = agg.setdefault(fam(d["name"]), {"us": 0.0, "n": 0, "rd": 0.0, "wr": 0.0})
        a["us"] += d.get("us", 0)
        a["n"] += 1
        a["rd"] += d.get("rd", 0)
        a["wr"] += d.get("wr", 0)
    total = sum(a["us"] for a in agg.values())
    out = [f"# ncu launch list summary `{tag}` -- one backbone+head micro-batch (cold-cache, serialised: compare SHARES)", "",
           f"source: `{os.path.basename(path)}`; {len(seq)} launches, {total:.0f} us", "",
           "| family | launches | time us | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|"]
    traffic = {}
    for f, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.append(f"| {f} | {a['n']} | {a['us']:.1f} | {100 * a['us'] / total:.1f} % | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} |")
        traffic[f] = {"dram_bytes_per_launch": (a["rd"] + a["wr"]) / max(1, a["n"]), "launches": a["n"],
                      "share_of_step": a["us"] / total}
    # per-launch list in execution order (kernel, grid-independent): lets a reader map every GEMM to its layer
    out += ["", "## launches in execution order", "", "| # | kernel | us | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|"]
    for i, d in enumerate(seq):
        short = d["name"].split("(")[0].replace("void ", "").replace("bq::", "")
        out.append(f"| {i} | `{short[:60]}` | {d.get('us', 0):.1f} | {d.get('rd', 0) / 1e6:.1f} | {d.get('wr', 0) / 1e6:.1f} |")
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, f"launches_{tag}.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    with open(os.path.join(here, "roofline_traffic.json"), "w") as f:
        json.dump({"source": os.path.basename(path), "tag": tag, **traffic}, f, indent=1)
    print("\n".join(out[:24]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
