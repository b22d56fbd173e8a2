#!/usr/bin/env python
"""Developer sweep on the GPU box: one synthetic batch (bench shape), many library variants / environment settings.

    python tools/stage_sweep.py [--records N] [--cells C] [--params run] VARIANT ...

A VARIANT is `name[:lib=libmgatk2_b200_x.so][:ENV=value]...`; the library file is looked up next to the default one.
Prints per-stage device times (mean of --steps runs after --warmup) and checks the counters and the per-cell QC rows of
every variant against the first one, so a variant that is fast because it is wrong shows up at once.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="*", default=["default"])
    ap.add_argument("--records", type=int, default=20_000_000)
    ap.add_argument("--cells", type=int, default=2000)
    ap.add_argument("--profile", default="atac50")
    ap.add_argument("--params", default="run")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()

    import torch
    from bench import BASE_SEED, CONFIG_INDEX, default_params
    from mgatk2_b200 import _lib
    from mgatk2_b200.synth import synth_batch

    batch = synth_batch(a.cells, a.records, a.profile, seed=BASE_SEED + CONFIG_INDEX)
    extent = batch.max_read_extent()
    params = default_params(a.cells, extent, a.params)
    default_lib = _lib.LIB_PATH
    ref = None
    for spec in a.variants:
        parts = spec.split(":")
        name, env, lib = parts[0], {}, default_lib
        for kv in parts[1:]:
            k, v = kv.split("=", 1)
            if k == "lib":
                lib = os.path.join(os.path.dirname(default_lib), v)
            else:
                env[k] = v
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        _lib._lib, _lib.LIB_PATH = None, lib            # a fresh ctypes instance of that file
        from mgatk2_b200.engine import PileupEngine
        eng = PileupEngine(0)
        dbatch = eng.upload(batch)
        dout = eng.alloc_device_outputs(a.cells, 16569, batch.n_records, max_read_extent=extent)
        for _ in range(a.warmup):
            eng.run_device(dbatch, params, dout)
        torch.cuda.synchronize()
        acc, tot = {}, []
        for _ in range(a.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run_device(dbatch, params, dout)
            e1.record()
            torch.cuda.synchronize()
            tot.append(e0.elapsed_time(e1))
            for k, v in eng.stage_times().items():
                acc[k] = acc.get(k, 0.0) + v / a.steps
        res = eng.download(dout, params)
        sig = (dict(res.stats), res.cell_qc.copy(), res.base_totals.copy(), int(res.planes.astype(np.uint64).sum()))
        ok = "ref"
        if ref is None:
            ref = sig
        else:
            same = (all(sig[0][k] == ref[0][k] for k in ref[0]) and np.array_equal(sig[1], ref[1])
                    and np.array_equal(sig[2], ref[2]) and sig[3] == ref[3])
            ok = "same" if same else "DIFFERENT"
        print(json.dumps({"variant": name, "ms": round(float(np.median(tot)), 4),
                          "stages": {k: round(v, 4) for k, v in acc.items()}, "check": ok}), flush=True)
        eng.close()
        del dbatch, dout, eng
        torch.cuda.empty_cache()
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


if __name__ == "__main__":
    main()
