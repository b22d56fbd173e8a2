#!/usr/bin/env python
"""Developer tool for one GPU call: reads a tools/stage_sweep.py log, and when a variant whose results are the same is
faster than `default` by more than --margin, copies its library over the default one (the macro defaults in the sources
are then set to match by hand). Usage: adopt_fastest.py sweep.log [--margin 0.01] [--only name,name]"""
import argparse, json, os, shutil
ap = argparse.ArgumentParser()
ap.add_argument("log"); ap.add_argument("--margin", type=float, default=0.01); ap.add_argument("--only", default="")
a = ap.parse_args()
rows = [json.loads(l) for l in open(a.log) if l.startswith("{")]
base = next(r for r in rows if r["variant"] == "default")
only = set(a.only.split(",")) if a.only else None
best = base
for r in rows:
    if r["check"] != "same" or (only and r["variant"] not in only):
        continue
    if r["ms"] < best["ms"] and r["ms"] < base["ms"] * (1 - a.margin):
        best = r
here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mgatk2_b200")
if best is not base:
    shutil.copyfile(os.path.join(here, f"libmgatk2_b200_{best['variant']}.so"), os.path.join(here, "libmgatk2_b200.so"))
print(json.dumps({"adopted": best["variant"], "ms": best["ms"], "default_ms": base["ms"]}))
