"""CPU tests (gloo, world_size 2) of the multi-rank logic of the sharded filter.

The device kernels are exercised on GPUs by tests/sharded_check.py; here two gloo ranks each
hold one shard of a synthetic weighted population and run the cross-rank protocol with the
host restatements in particle_filters_b200/sharded.py:
 all-gather of partial records -> identical normaliser / ESS on every rank;
 all-gather of fixed-point shard totals -> exclusive scan -> per-rank output ranges and the
 cross-shard migration plan, checked against the oracle's ancestors.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist

    from oracle import systematic_ancestors
    from particle_filters_b200.sharded import (combine_partial_records, frac_bits_for, migration_plan,
                                               outputs_below, shard_offsets)

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 1 << 14
        rng = np.random.default_rng(11)          # same stream on both ranks: the global truth
        x = rng.standard_normal(N)
        logw = -0.5 * (0.8 - np.exp(0.5 * x)) ** 2 - np.log(N)
        u = 0.6180339887
        n = N // world
        sl = slice(rank * n, (rank + 1) * n)
        # 1. partial record of this shard (what k1 + k2_reduce_partials emit), shift c = 0
        m = logw[sl].max()
        e = np.exp(logw[sl] - m)
        rec = torch.tensor([m, e.sum(), (e * e).sum(), (e * x[sl]).sum(), (e * x[sl] ** 2).sum()], dtype=torch.float64)
        allrec = [torch.zeros_like(rec) for _ in range(world)]
        dist.all_gather(allrec, rec)
        comb = combine_partial_records(np.stack([r.numpy() for r in allrec]))
        log_norm = comb[0] + np.log(comb[1])
        ess = comb[1] ** 2 / comb[2]
        mean = comb[3] / comb[1]
        # truth from the whole population
        M = logw.max()
        E = np.exp(logw - M)
        assert abs(log_norm - (M + np.log(E.sum()))) < 1e-12
        assert abs(ess - E.sum() ** 2 / (E * E).sum()) < 1e-8 * N
        assert abs(mean - (E * x).sum() / E.sum()) < 1e-12
        # 2. fixed-point shard totals -> exclusive scan
        F = frac_bits_for(N)
        w = np.exp(logw - log_norm)
        qfix = [int(round(float(v) * N * (1 << F))) for v in w[sl]]
        tot = torch.tensor([sum(qfix)], dtype=torch.int64)
        alltot = [torch.zeros_like(tot) for _ in range(world)]
        dist.all_gather(alltot, tot)
        totals = [int(t.item()) for t in alltot]
        off = shard_offsets(totals)
        assert off[0] == 0 and off[rank] == sum(totals[:rank])
        # 3. my output range and the migration plan vs the oracle's ancestors
        anc = systematic_ancestors(w / w.sum(), u)
        lo = outputs_below(off[rank], u, N)
        hi = N if rank == world - 1 else outputs_below(off[rank] + totals[rank], u, N)
        mine = np.nonzero((anc >= rank * n) & (anc < (rank + 1) * n))[0]
        assert abs(lo - mine.min()) <= 1 and abs(hi - (mine.max() + 1)) <= 1      # ties only
        plan = migration_plan(totals, u, N)
        truth = np.zeros_like(plan)
        for s in range(world):
            for d in range(world):
                truth[s, d] = np.sum((anc // n == s) & (np.arange(N) // n == d))
        assert plan.sum() == N
        assert np.abs(plan - truth).max() <= 1
        # each rank reaches the same plan (everything derives from all-gathered integers)
        plans = [None] * world
        dist.all_gather_object(plans, plan.tolist())
        assert plans[0] == plans[1]
        q.put((rank, "ok", int(plan[0, 1] + plan[1, 0])))
    except Exception as exc:  # pragma: no cover
        q.put((rank, f"FAIL: {exc!r}", -1))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_two_gloo_ranks():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, crossing in results:
        assert status == "ok", f"rank {rank}: {status}"
    assert results[0][2] == results[1][2] and results[0][2] >= 0


def _fused_worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist

    from oracle import systematic_ancestors
    from particle_filters_b200.sharded import (contribution_row, migration_plan, outputs_below, quantised_weights_f32,
                                               shard_offsets, shard_totals_from_contributions)

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # One fused step seen from rank `rank`: it holds shard `rank` of the step-(t-1) population, resamples it (its
        # outputs are the indices [lo, hi) of the global resampled set, which may belong to the other rank), pushes
        # them through the "functor" (here: new log-weights drawn per GLOBAL output index, like Philox) and publishes its
        # row of the contribution table.  Check: column sums = the true per-shard totals of the step-t weights.
        N = 1 << 13
        n = N // world
        rng = np.random.default_rng(5)
        lw_prev = (-0.5 * (rng.standard_normal(N) * 1.7) ** 2 - np.log(N)).astype(np.float32)   # bootstrap: <= -log N
        a_prev = quantised_weights_f32(lw_prev, N)
        totals_prev = [int(sum(a_prev[r * n:(r + 1) * n])) for r in range(world)]
        u = 0.3819660113
        total = sum(totals_prev)
        off = shard_offsets(totals_prev)
        lo = 0 if rank == 0 else outputs_below(off[rank], u, N, f32=True, total=total)
        hi = N if rank == world - 1 else outputs_below(off[rank] + totals_prev[rank], u, N, f32=True, total=total)
        # the step-t log-weight of global output index i depends on i alone (Philox is keyed by the global index)
        lw_new_all = (-0.5 * (np.random.default_rng(99).standard_normal(N) * 1.3) ** 2 - np.log(N)).astype(np.float32)
        a_new_all = quantised_weights_f32(lw_new_all, N)
        row = contribution_row(a_new_all[lo:hi], lo, N, world)
        table = [None] * world
        dist.all_gather_object(table, row)
        totals_new = shard_totals_from_contributions(table)
        truth = [int(sum(a_new_all[r * n:(r + 1) * n])) for r in range(world)]
        assert totals_new == truth, (totals_new, truth)
        # the output ranges of the ranks tile [0, N) and agree with the migration plan and with the oracle's ancestors
        ranges = [None] * world
        dist.all_gather_object(ranges, (lo, hi))
        assert ranges[0][0] == 0 and ranges[-1][1] == N and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        plan = migration_plan(totals_prev, u, N, f32=True)
        assert plan.sum() == N and plan[rank].sum() == hi - lo
        w = np.array([float(v) for v in a_prev]) / float(total)
        anc = systematic_ancestors(w, u)
        mine = np.nonzero((anc >= rank * n) & (anc < (rank + 1) * n))[0]
        assert abs(lo - mine.min()) <= 1 and abs(hi - (mine.max() + 1)) <= 1                 # round-off ties only
        q.put((rank, "ok", int(plan.sum() - np.trace(plan))))
    except Exception as exc:  # pragma: no cover
        q.put((rank, f"FAIL: {exc!r}", -1))
    finally:
        dist.destroy_process_group()


def test_fused_sharded_contribution_table_two_gloo_ranks():
    """Host restatement of what the fused sharded step exchanges (one slot per launch): every rank's row of the G x G
    table of fixed-point weight contributions; its column sums are the shard totals of the next resampling."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 150)
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, crossing in results:
        assert status == "ok", f"rank {rank}: {status}"
    assert results[0][2] == results[1][2] and results[0][2] > 0        # some particles do cross the shard boundary


def test_combine_partial_records_matches_flat_reduction():
    from particle_filters_b200.sharded import combine_partial_records

    rng = np.random.default_rng(0)
    lw = rng.standard_normal(1000) * 5
    x = rng.standard_normal(1000)
    recs = []
    for chunk in np.array_split(np.arange(1000), 7):
        m = lw[chunk].max()
        e = np.exp(lw[chunk] - m)
        recs.append([m, e.sum(), (e * e).sum(), (e * x[chunk]).sum()])
    recs.append([-np.inf, 0.0, 0.0, 0.0])        # an empty / all-zero-weight shard
    c = combine_partial_records(np.array(recs))
    M = lw.max()
    E = np.exp(lw - M)
    np.testing.assert_allclose(c, [M, E.sum(), (E * E).sum(), (E * x).sum()], rtol=1e-13)


def test_outputs_below_agrees_with_searchsorted():
    from particle_filters_b200.sharded import frac_bits_for, outputs_below

    N, u = 5000, 0.123
    F = frac_bits_for(N)
    rng = np.random.default_rng(2)
    for c in rng.random(200):
        C = int(c * N * (1 << F))
        want = int(np.searchsorted((u + np.arange(N)) / N, c, side="left"))
        assert abs(outputs_below(C, u, N) - want) <= 1
