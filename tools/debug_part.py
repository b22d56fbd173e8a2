import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mgatk2_b200._lib import ParamsC
from mgatk2_b200.engine import PileupEngine
from mgatk2_b200.synth import synth_batch
from oracle.oracle import make_params, run_oracle
b = synth_batch(20, 10_000_000, "stress150", seed=77)
q = b.split_on_start_borders(5)[0]
q = q.take(np.arange(q.n_records))
n = int(os.environ.get("DBG_N", q.n_records))
q = q.slice(0, n)
eng = PileupEngine(0)
p = make_params(20, 20, 30, 10, 0, 0.8, 0, max_read_extent=q.max_read_extent())
lp = ParamsC(*[getattr(p, f) for f, _ in p._fields_])
ora = run_oracle(q, p, n_threads=8, dense=False)
db = eng.upload(q)
for rep in range(int(os.environ.get("DBG_REPS", 3))):
    do = eng.alloc_device_outputs(20, 16569, q.n_records, overflow_capacity=1 << 16, max_read_extent=q.max_read_extent())
    eng.run_device(db, lp, do)
    r = eng.download(do, lp)
    print(rep, q.n_records, "oracle", ora.stats["filtered_reads"], ora.stats["stage1_reads"], "device", r.stats["filtered_reads"], r.stats["stage1_reads"],
          int(ora.cell_qc["sum_depth"].sum()), int(r.cell_qc["sum_depth"].sum()),
          "cells differing", np.nonzero(r.cell_qc["n_reads"] != ora.cell_qc["n_reads"])[0].tolist(), flush=True)
