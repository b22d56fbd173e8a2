#!/usr/bin/env python
"""Per-kernel time and DRAM bytes of ONE step from an ncu launch list, and the step totals bench.py reports as
`roofline.traffic`.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file launches.csv python tools/stage_sweep.py --steps 1 --warmup 2 default
    python tools/ncu_traffic.py launches.csv [--json profiles/traffic.json] [--source "..."]

The list holds warm-up steps too: the LAST occurrence of every kernel (and of every repeated kernel position inside a
step) is the measured step. Kernels are grouped into the stages bench.py times (`stage_ms`).
"""
import argparse
import csv
import json
import re

STAGE_OF = (("k_hist", "filter+planes+partition"), ("k_scan", "filter+planes+partition"), ("k_scatter", "filter+planes+partition"),
            ("k_publish_m", "filter+planes+partition"), ("k_init", "filter+planes+partition"), ("k_dedup", "dedup"), ("k_find_long_runs", "dedup"), ("k_plan", "plan"),
            ("k_pileup", "pileup"), ("k_totals", "totals+median"), ("k_base_totals", "totals+median"), ("k_median", "totals+median"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json")
    ap.add_argument("--source", default="")
    ap.add_argument("--steps-in-list", type=int, default=3)
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv, errors="replace")))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = {}
    for r in rows[i + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        lid = int(r[col["ID"]])
        full = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
        name = re.sub(r"<.*", "", full).split("::")[-1] + ("<compact>" if "<1" in full or "<(bool)1" in full else "")
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        launches.setdefault(lid, {"name": name})[r[col["Metric Name"]]] = v * scale
    ids = sorted(launches)
    per_step = len(ids) // a.steps_in_list
    step = [launches[k] for k in ids[-per_step:]]          # the last step of the list
    stages, total_b, total_us = {}, 0.0, 0.0
    print(f"{'kernel':44s} {'us':>9s} {'DRAM read MB':>13s} {'DRAM write MB':>14s}")
    for L in step:
        us = L.get("gpu__time_duration.sum", 0.0)
        rd, wr = L.get("dram__bytes_read.sum", 0.0), L.get("dram__bytes_write.sum", 0.0)
        print(f"{L['name'][:44]:44s} {us:9.1f} {rd / 1e6:13.1f} {wr / 1e6:14.1f}")
        st = next((s for k, s in STAGE_OF if L["name"].startswith(k)), "other")
        stages[st] = stages.get(st, 0.0) + rd + wr
        total_b += rd + wr
        total_us += us
    print(f"{'one step, ' + str(len(step)) + ' launches':44s} {total_us:9.1f} {'':13s} {total_b / 1e6:14.1f} MB in all")
    for st, b in stages.items():
        print(f"  stage {st:28s} {b / 1e6:10.1f} MB")
    if a.json:
        json.dump({"step_dram_bytes": int(total_b), "stages": {k: int(v) for k, v in stages.items()},
                   "launches_per_step": len(step), "serialised_step_us": total_us, "source": a.source,
                   "config": "C2: 2000 cells x 20M records"}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
