"""The numbers bench.py quotes from profiles/ can be recomputed from what is committed there: `traffic.json` (read into
`roofline.traffic`) is exactly what tools/ncu_traffic.py makes of the committed ncu launch list, the stage shares of that
list agree with the `stage_ms` of the committed bench line, and the algorithmic bytes of the bench configuration are the
2.61 GB of SURVEY.md section 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_traffic_json_follows_from_the_launch_list(tmp_path):
    out = tmp_path / "t.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), os.path.join(PROF, "r2_launches_c2_20M.csv"),
                    "--json", str(out)], check=True, capture_output=True)
    got, want = json.load(open(out)), json.load(open(os.path.join(PROF, "traffic.json")))
    for k in ("step_dram_bytes", "stages", "launches_per_step"):
        assert got[k] == want[k], k
    assert abs(got["serialised_step_us"] - want["serialised_step_us"]) < 1e-6


def test_launch_list_shares_agree_with_the_bench_line():
    line = json.loads(open(os.path.join(PROF, "r2_bench_c2_1gpu.json")).read().strip().splitlines()[-1])
    stage_ms = line["roofline"]["stage_ms"]
    txt = open(os.path.join(PROF, "r2_launches_c2_20M.txt")).read().splitlines()
    us = {}
    for row in txt[1:]:
        parts = row.split()
        if row.startswith("one step"):
            break
        us[parts[0]] = us.get(parts[0], 0.0) + float(parts[1])
    total_ncu = sum(us.values())
    big = {"filter+planes+partition": "k_scatter_planes<compact>", "dedup": "k_dedup<compact>", "pileup": "k_pileup_main<compact>"}
    step_ms = line["ms_per_step"]
    assert abs(sum(stage_ms.values()) - step_ms) < 0.02 * step_ms              # the stages are the step
    for stage, kernel in big.items():
        share_ncu = us[kernel] / total_ncu                                       # serialised, cold caches: shares, not absolutes
        share_bench = stage_ms[stage] / step_ms
        assert abs(share_ncu - share_bench) < 0.06, (stage, share_ncu, share_bench)
    r = line["roofline"]
    assert abs(r["achieved"] - r["algorithmic_bytes_per_step"] / (step_ms * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9


def test_algorithmic_bytes_of_the_bench_configuration():
    line = json.loads(open(os.path.join(PROF, "r2_bench_c2_1gpu.json")).read().strip().splitlines()[-1])
    # 94 B per 2x50 bp record + 22 B per (cell, position) + 32 B per cell
    assert line["roofline"]["algorithmic_bytes_per_step"] == 20_000_000 * 94 + 2000 * 16569 * 22 + 2000 * 32
