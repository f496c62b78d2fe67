"""CPU tests of the multi-GPU host logic: pair-count-balanced, tile-aligned wavenumber chunks, per-rank line
subsets and the padded all-gather -- including a real world_size-2 run over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import physics as ph
from pyrad_b200 import distributed as pd
from pyrad_b200 import partition as pt
from pyrad_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_cost_equals_reference_pair_count():
    rng = np.random.default_rng(0)
    n_total = 20000
    idx = np.sort(rng.integers(-600, n_total + 600, 3000))
    for windows in ([500], [3, 50, 500], [1], [2, 700, 700]):
        cost = pt.block_pair_cost(idx, n_total, windows, block=1024)
        assert cost.sum() == sum(ph.pair_count(idx, n_total, w) for w in windows)
        b = 5
        brute = 0
        for w in windows:
            wm = max(w - 2, 0)
            lo = np.maximum(idx - wm, b * 1024)
            hi = np.minimum(idx + wm, min((b + 1) * 1024, n_total) - 1)
            brute += np.maximum(hi - lo + 1, 0).sum()
        assert cost[b] == brute


def test_chunks_are_aligned_contiguous_and_balanced():
    rng = np.random.default_rng(1)
    n_total = 1_000_000
    # strongly non-uniform line density: a band head
    nu = np.sort(np.concatenate([rng.uniform(0, 1000, 20000), rng.normal(650, 15, 80000)]))
    idx = pt.line_index(nu, 0.0, 0.001)
    cost = pt.block_pair_cost(idx, n_total, [5000, 2000, 300])
    for world in (2, 4, 8):
        ch = pt.balanced_chunks(cost, n_total, world)
        assert ch[0][0] == 0 and ch[-1][1] == n_total
        assert all(a[1] == b[0] for a, b in zip(ch[:-1], ch[1:]))
        assert all(a % pt.ALIGN == 0 for a, _ in ch)
        per = [cost[a // pt.ALIGN:(b + pt.ALIGN - 1) // pt.ALIGN].sum() for a, b in ch]
        assert max(per) <= 1.15 * cost.sum() / world + cost.max()
        eq = pt.equal_chunks(n_total, world)
        per_eq = [cost[a // pt.ALIGN:(b + pt.ALIGN - 1) // pt.ALIGN].sum() for a, b in eq]
        assert max(per) <= max(per_eq)


def test_time_cost_weights_classes_and_points():
    rng = np.random.default_rng(2)
    n_total = 200_000
    idx = np.sort(rng.integers(0, n_total, 50_000))
    wide = pt.block_time_cost(idx, n_total, [5000])
    assert np.allclose(wide, pt.block_pair_cost(idx, n_total, [5000]) / 4.9e12)
    # a window of one sample: hardly any pairs, the per-point term carries the cost
    tiny = pt.block_time_cost(idx, n_total, [1] * 10)
    assert tiny.sum() > 10 * n_total * 1e-11 and tiny.min() > 0
    mixed = pt.block_time_cost(idx, n_total, [5000, 300, 1])
    assert np.allclose(mixed, pt.block_time_cost(idx, n_total, [5000]) + pt.block_time_cost(idx, n_total, [300]) +
                       pt.block_time_cost(idx, n_total, [1]))


def test_feedback_rebalancing_moves_work_off_the_slow_rank():
    rng = np.random.default_rng(4)
    n_total = 2_000_000
    idx = np.sort(rng.integers(0, n_total, 400_000))
    cost = pt.block_time_cost(idx, n_total, [5000, 800, 60])
    chunks = pt.balanced_chunks(cost, n_total, 4)
    # the model is blind to a cost that grows with wavenumber: rank r really takes (1 + 0.1 r) x its prediction
    true = cost * (1 + 0.3 * np.arange(len(cost)) / len(cost))
    per = lambda ch, c: [c[a // pt.ALIGN:(b + pt.ALIGN - 1) // pt.ALIGN].sum() for a, b in ch]
    measured = per(chunks, true)
    new, corrected = pt.rebalanced_chunks(cost, chunks, measured, n_total)
    assert new[0][0] == 0 and new[-1][1] == n_total and all(a[1] == b[0] for a, b in zip(new[:-1], new[1:]))
    assert max(per(new, true)) < max(measured)                       # the slowest rank got lighter
    assert max(per(new, true)) <= 1.03 * np.sum(true) / 4 + true.max()
    plan = pd.ShardPlan(np.sort(rng.uniform(0, 2000, 50_000)), 0.0, 0.001, n_total, [5000, 60], 1, 4)
    again = plan.rebalanced([1.0, 1.3, 1.0, 1.0])
    assert again is not plan and again.chunks != plan.chunks and again.chunks[1][1] - again.chunks[1][0] < plan.chunks[1][1] - plan.chunks[1][0]
    assert pd.ShardPlan(np.zeros(0), 0.0, 0.001, 8192, [5], 0, 1).rebalanced([1.0]).chunks == [(0, 8192)]


def test_line_subset_reaches_every_window():
    ln = synth.make_lines(5000, 0.0, 100.0, 5)
    idx = pt.line_index(ln["nu"], 0.0, 0.01)
    l0, l1 = pt.lines_for_chunk(ln["nu"], 0.0, 0.01, 4096, 8192, 498)
    inside = np.nonzero((idx >= 4096 - 498) & (idx <= 8191 + 498))[0]
    # every line that reaches the chunk, from a start rounded down to a multiple of four lines (K2's staging chunks
    # then fall where they do in the unsharded run: bitwise shard invariance)
    assert l0 % 4 == 0 and inside[0] - 3 <= l0 <= inside[0] and l1 == inside[-1] + 1


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ln = synth.make_lines(800, 598.0, 642.0, 21)
        rmin, rmax, res, P, T = 600.0, 640.0, 0.0025, 300.0, 250
        cutoff = ph.layer_cutoff(P)
        n_total = ph.grid_len(rmin, rmax, res)
        W = ph.window_len(cutoff, res)
        plan = pd.ShardPlan(ln["nu"], rmin, res, n_total, [W], rank, world)
        sub = plan.subset(ln)
        # stand-in for the CUDA engine: the oracle evaluates ONLY this rank's chunk from ONLY its line subset
        pts = np.arange(plan.i_begin, plan.i_end)
        local = ph.cross_section_at(pts, sub, T, P, 4e-4, 43.98983, 250.0, 286.09, rmin, rmax, res, cutoff)
        g = pd.all_gather_spectra(torch.from_numpy(local), plan, dist)
        full = pd.assemble(g, plan).numpy()
        ref = ph.cross_section(ln, T, P, 4e-4, 43.98983, 250.0, 286.09, rmin, rmax, res, cutoff)
        ok = full.shape == ref.shape and np.allclose(full, ref, rtol=1e-12, atol=0)
        q.put((rank, bool(ok), plan.chunks))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_all_gather_assembles_the_unsharded_spectrum():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2]                       # both ranks derived the same plan


def test_farfield_work_accounting_matches_brute_force():
    """partition.farfield_work restates the far-field kernel's integer class tests (level 1: spans, level 2: domains of
    four spans); checked against a direct count."""
    rng = np.random.default_rng(5)
    idx = np.sort(rng.integers(-300, 6000, 900))

    def far(f, first, length, wm):
        last = first + length - 1
        full = last - wm <= f <= first + wm
        return full and (f < first + (length - 1) // 2 - length or f > first + length // 2 + length)

    for (a, b, window, span, l2) in ((0, 4096, 1400, 256, 0), (4096, 5500, 1400, 256, 0), (0, 4096, 700, 128, 0),
                                     (0, 1000, 200, 128, 0), (0, 4096, 2600, 256, 4), (2048, 5500, 2600, 128, 4),
                                     (0, 4096, 2600, 128, 8)):
        wm = window - 2
        exact, nodes = pt.farfield_work(idx, a, b, window, span, level2_spans=l2, level2_min_domains=0)
        want_exact = want_nodes = 0
        for first in range(a, b, span):
            pts = min(first + span - 1, b - 1) - first + 1
            dfirst = a + ((first - a) // (span * l2)) * span * l2 if l2 else 0
            for f in idx:
                covered = sum(1 for i in range(first, first + pts) if abs(i - f) <= wm)
                if l2 and far(f, dfirst, span * l2, wm):
                    assert far(f, first, span, wm)                       # level 2 is a subset of level 1
                    want_nodes += 16 if first == dfirst else 0           # once per domain
                elif far(f, first, span, wm):
                    want_nodes += 16
                else:
                    want_exact += covered
        assert (exact, nodes) == (want_exact, want_nodes), (a, b, window, span, l2)
    assert pt.farfield_work(idx, 0, 4096, 1400, 256)[1] > 0
    assert pt.farfield_work(idx, 0, 4096, 2600, 256, level2_spans=4, level2_min_domains=0)[1] < pt.farfield_work(idx, 0, 4096, 2600, 256, level2_spans=0)[1]
    # the kernel's own rule: no level 2 below four domain lengths of window
    assert pt.farfield_work(idx, 0, 4096, 2600, 256) == pt.farfield_work(idx, 0, 4096, 2600, 256, level2_spans=0)


def test_farfield_cost_model_tracks_the_far_field_work():
    """block_time_cost(farfield=True) prices the wide classes by what the far-field kernel evaluates: per block it stays
    within a small factor of the exact accounting (partition.farfield_work) and far below the all-pairs cost."""
    rng = np.random.default_rng(11)
    n = 65536
    idx = np.sort(rng.integers(-6000, n + 6000, 40000))
    for window in (5000, 1400):
        exact_cost = pt.block_time_cost(idx, n, [window])
        far_cost = pt.block_time_cost(idx, n, [window], farfield=True)
        assert far_cost.shape == exact_cost.shape and np.all(far_cost > 0)
        assert far_cost.sum() < (0.4 if window == 5000 else 0.7) * exact_cost.sum()
        span = 256 if window - 2 >= 1024 else 128
        c_pair = pt._class_cost(window - 2)[0]
        for b in (2, 7, 12):
            ex, nodes = pt.farfield_work(idx, b * 4096, (b + 1) * 4096, window, span)
            want = c_pair * (ex + 2.0 * nodes)
            assert 0.6 * want <= far_cost[b] <= 1.6 * want, (window, b, far_cost[b], want)
    # level 2 (cfg5-like window): the far term shrinks again
    assert pt.block_time_cost(idx, n, [25000], farfield=True).sum() < 0.1 * pt.block_time_cost(idx, n, [25000]).sum()


def test_shard_plans_on_random_inputs_are_partitions_of_the_grid():
    """Random grids (ten points to millions), line lists (empty, one line, tens of thousands), window sets, world sizes
    (also more ranks than tiles) and both cost models: every rank derives the same chunks; they are contiguous,
    tile-aligned, cover [0, n_total) exactly and may be empty but never overlap."""
    rng = np.random.default_rng(0)
    for _ in range(150):
        world = int(rng.choice([1, 2, 3, 4, 8]))
        n_total = int(np.exp(rng.uniform(np.log(10), np.log(3e6))))
        res = float(rng.choice([0.1, 0.01, 0.001]))
        rmin = float(rng.choice([0.0, 600.0]))
        nu = np.sort(rng.uniform(max(rmin - 5, 0), rmin + n_total * res + 5, int(rng.choice([0, 1, 5, 1000, 50000]))))
        wins = [int(np.exp(rng.uniform(0, np.log(30000)))) for _ in range(int(rng.integers(1, 5)))]
        ff = bool(rng.random() < 0.5)
        plans = [pd.ShardPlan(nu, rmin, res, n_total, wins, r, world, farfield=ff) for r in range(world)]
        ch = plans[0].chunks
        assert all(p.chunks == ch for p in plans) and len(ch) == world
        assert ch[0][0] == 0 and ch[-1][1] == n_total
        assert all(a[1] == b[0] for a, b in zip(ch[:-1], ch[1:])) and all(a <= b for a, b in ch)
        assert all(a % pt.ALIGN == 0 or a == n_total for a, _ in ch)
