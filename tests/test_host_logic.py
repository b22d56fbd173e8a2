"""CPU-only checks of the host side: batch packing, subsetting, synthetic generator invariants, config mirror,
barcode sharding (including a world_size-2 gloo run that recombines per-rank results)."""
import os

import numpy as np

from mgatk2_b200.batch import ReadBatch
from mgatk2_b200.sharding import assign_cells, combine_columns, combine_stats, shard_batch
from mgatk2_b200.synth import make_whitelist, synth_batch


def test_record_round_trip_and_take():
    recs = [dict(pos=5, flag=99, mapq=60, seq="ACGTNACGT", qual=[1, 2, 3, 4, 5, 6, 7, 8, 9], cigar=[(4, 2), (0, 7)], tlen=-7, bc_idx=3),
            dict(pos=9, flag=16, mapq=0, seq="", qual=[], cigar=[], tlen=0, bc_idx=-1),
            dict(pos=9, flag=0, mapq=255, seq="RYKM=", qual=[40] * 5, cigar=[(0, 2), (1, 1), (2, 4), (0, 2)], tlen=2 ** 31 - 1, bc_idx=0)]
    b = ReadBatch.from_records(recs)
    assert b.n_records == 3 and len(b.blob) % 16 == 0 and b.is_sorted()
    for i, r in enumerate(recs):
        got = b.record(i)
        assert got == {k: r[k] for k in got}
    assert b.reference_span().tolist() == [7, 0, 8] and b.max_read_extent() == 9
    sub = b.take(np.array([2, 0]))
    assert sub.record(0) == b.record(2) and sub.record(1) == b.record(0)
    assert sub.blob_off.tolist() == [0, 2] and len(sub.blob) == 64


def test_synth_invariants():
    b = synth_batch(50, 20_000, "stress150", seed=3)
    assert b.n_records == 20_000 and b.is_sorted() and len(b.blob) % 16 == 0
    assert (b.bc_idx >= -2).all() and (b.bc_idx < 50).all() and (b.pos >= 0).all() and (b.pos < 16569).all()
    # cigars consume exactly l_seq query bases
    q = np.zeros(b.n_records, np.int64)
    for k in range(int(b.n_cigar.max())):
        idx, w = b.cigar_words(k)
        q[idx] += np.where(np.isin(w & 0xF, (0, 1, 4, 7, 8)), w >> 4, 0)
    assert (q == b.l_seq).all()
    b2 = synth_batch(50, 20_000, "stress150", seed=3)
    assert all(np.array_equal(getattr(b, f), getattr(b2, f)) for f in ("pos", "tlen", "flag", "blob"))
    assert len(set(make_whitelist(500))) == 500


def test_config_mirror_defaults():
    from mgatk2_b200 import PipelineConfig
    cfg = PipelineConfig()
    assert (cfg.quality.min_baseq, cfg.quality.min_mapq, cfg.quality.min_distance_from_end) == (20, 30, 5)
    assert cfg.quality.max_strand_bias == 0.9 and cfg.mito_length == 16569 and cfg.barcode_tag == "CB"
    assert cfg.dedup.mode == 0 and PipelineConfig(use_fragment_length_dedup=False).dedup.mode == 1
    assert PipelineConfig(skip_deduplication=True).dedup.mode == 2
    p = cfg.to_params(7, 50)
    assert (p.n_cells, p.max_read_extent, p.min_distance_from_end, p.flags) == (7, 50, 5, 0)


def test_sharding_partitions_every_record_once():
    b = synth_batch(37, 30_000, "atac50", seed=5)
    counts = np.bincount(b.bc_idx[b.bc_idx >= 0], minlength=37)
    for owner in (assign_cells(37, 4), assign_cells(37, 4, counts)):
        total = 0
        for r in range(4):
            sub, cols, unowned = shard_batch(b, owner, r)
            assert sub.is_sorted() and (sub.bc_idx >= 0).all() and sub.bc_idx.max() < len(cols)
            assert (owner[cols] == r).all()
            total += sub.n_records + unowned
        assert total == b.n_records
    loads = np.bincount(assign_cells(37, 4, counts), weights=counts, minlength=4)
    assert loads.max() / loads.mean() < 1.15


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import make_params, run_oracle
    n_cells = 23
    b = synth_batch(n_cells, 12_000, "atac50", seed=9)
    owner = assign_cells(n_cells, world)
    sub, cols, unowned = shard_batch(b, owner, rank)
    res = run_oracle(sub, make_params(len(cols)))          # stands in for this rank's GPU: same per-shard contract
    totals = torch.from_numpy(res.base_totals.copy())
    dist.all_reduce(totals)                                 # the only cross-rank reduction (reference-allele vote)
    gathered = [None] * world
    dist.all_gather_object(gathered, (cols, res.cell_qc, res.stats, unowned))
    if rank == 0:
        full = run_oracle(b, make_params(n_cells))
        qc = combine_columns(n_cells, [g[0] for g in gathered], [g[1] for g in gathered])
        stats = combine_stats([g[2] for g in gathered], sum(g[3] for g in gathered))
        ok = bool(np.array_equal(qc, full.cell_qc) and np.array_equal(totals.numpy(), full.base_totals)
                  and all(stats[k] == full.stats[k] for k in ("total_reads", "filtered_reads", "dup_with_length",
                                                              "dup_position_only", "stage1_reads")))
        open(tmp, "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_gloo_recombination(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_column_plan_and_routing():
    """dispatch.py on the CPU: columns only for observed barcodes, contiguous device ranges balanced by record count,
    routing keeps file order and renumbers bc_idx per device; every usable record lands on exactly one device."""
    from mgatk2_b200 import dispatch
    from mgatk2_b200.synth import synth_batch
    n_wl = 500
    batch = synth_batch(40, 6000, "atac50", seed=3)
    batch.bc_idx = np.where(batch.bc_idx >= 0, batch.bc_idx * 7 + 11, batch.bc_idx).astype(np.int32)   # spread over the list
    ok = dispatch.usable(batch, n_wl)
    counts = np.bincount(batch.bc_idx[ok], minlength=n_wl)
    for n_dev in (1, 2, 3, 8):
        plan = dispatch.plan_columns(n_wl, counts, n_dev)
        assert plan.columns.tolist() == np.nonzero(counts)[0].tolist()
        assert plan.cuts[0] == 0 and plan.cuts[-1] == plan.n_columns and len(plan.cuts) == n_dev + 1
        assert all(a <= b for a, b in zip(plan.cuts[:-1], plan.cuts[1:]))
        loads = [int(counts[plan.columns[a:b]].sum()) for a, b in zip(plan.cuts[:-1], plan.cuts[1:])]
        assert max(loads) <= counts.sum() / n_dev + counts.max()
        subs, unowned = dispatch.route(batch, plan, n_wl)
        if n_dev == 1:
            assert unowned == 0 and subs[0].n_records == batch.n_records
            np.testing.assert_array_equal(subs[0].bc_idx >= 0, ok)
            continue
        assert unowned == int((~ok).sum()) and sum(s.n_records for s in subs) == int(ok.sum())
        for d, s in enumerate(subs):
            glob = plan.columns[s.bc_idx + plan.cuts[d]]
            want = np.nonzero(ok & (plan.device_of(plan.local_of[np.clip(batch.bc_idx, 0, n_wl - 1)]) == d))[0]
            np.testing.assert_array_equal(glob, batch.bc_idx[want])
            np.testing.assert_array_equal(s.pos, batch.pos[want])
            assert s.is_sorted()
    # nothing observed / whitelist columns for a stream whose barcodes are not known in advance
    assert dispatch.plan_columns(n_wl, np.zeros(n_wl, np.int64), 4).n_columns == 0
    assert dispatch.plan_columns(7, None, 2).columns.tolist() == list(range(7))
