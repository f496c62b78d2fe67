"""Multi-GPU check (run under torchrun, one rank per GPU): the peer-memory gather fused into the compute
kernels must deliver, on EVERY rank, exactly what a separate NCCL all-gather of the ranks' local spectra
delivers -- for the single-layer fused K2 epilogue and for the multi-layer K3 fold, over several steps
(double buffering).  Prints 'peer_check ok' on rank 0; any mismatch raises."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrad_b200 import distributed as pd           # noqa: E402
from pyrad_b200 import engine as eng               # noqa: E402
from pyrad_b200 import workloads                   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    e = eng.Engine(local)
    torch.cuda.set_stream(torch.cuda.ExternalStream(e.stream))
    for n_layers, top, variant, split in ((1, 2.0, eng.K2_CLASSED, 0), (6, 45.0, eng.K2_CLASSED, 0),
                                          (1, 2.0, eng.K2_FARFIELD, 1), (6, 45.0, eng.K2_FARFIELD, 1)):
        # (the far-field variant with line-range parts: the configuration bench.py strong-scales with)
        e.set_k2_variant(variant, 0)
        e.set_option(eng.OPT_SPLIT_TILES, split)
        w = workloads.atmosphere(n_layers=n_layers, n_lines=20000, rmin=600.0, rmax=700.0, res=0.001, top_km=top)
        sp = w["species"]
        n_total = eng.grid_len(w["range_min"], w["range_max"], w["res"])
        win = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
        plan = pd.ShardPlan(w["lines"]["nu"], w["range_min"], w["res"], n_total, win, rank, world)
        e.upload_lines(plan.subset(w["lines"]), n_groups=len(sp))
        e.set_grid(w["range_min"], w["res"], n_total, plan.i_begin, plan.i_end)
        qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
        args = (w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp], win,
                w["t_surface"], w["range_max"])
        nc = plan.i_end - plan.i_begin
        # reference: local results + NCCL all-gather
        e.atmosphere(*args)
        rp, tp = e.atmosphere_result_dev()
        ref_r = pd.all_gather_spectra(pd.device_tensor(rp, nc), plan, dist).clone()
        ref_t = pd.all_gather_spectra(pd.device_tensor(tp, nc), plan, dist).clone()
        torch.cuda.synchronize()
        pd.connect_peers(e, rank, world, plan.max_chunk, dist)
        for step in range(4):
            e.atmosphere(*args)
            g_r, g_t = pd.gathered_spectra(e)
            for r, (a, b) in enumerate(plan.chunks):
                assert torch.equal(g_r[r, : b - a], ref_r[r, : b - a]), (n_layers, variant, step, rank, r, "radiance")
                assert torch.equal(g_t[r, : b - a], ref_t[r, : b - a]), (n_layers, variant, step, rank, r, "transmittance")
        full = pd.assemble(g_t, plan)
        assert full.numel() == n_total
        dist.barrier()
        e.peer_disconnect()
    # random columns (PEER_FUZZ=<cases>, same seed on every rank): grids from a fraction of a tile (ranks with EMPTY chunks)
    # to hundreds of tiles, 1-10 layers, both variants, line-range parts on / off; the gathered spectrum must equal the
    # NCCL all-gather on every rank AND, assembled, the unsharded run of the same column bit for bit
    n_fuzz = int(os.environ.get("PEER_FUZZ", "0"))
    for case in range(n_fuzz):
        rng = np.random.default_rng(77 + case)
        n_layers = int(rng.integers(1, 11))
        res = float(rng.choice([0.01, 0.002, 0.001]))
        n_grid = int(np.exp(rng.uniform(np.log(500.0), np.log(600000.0))))
        rmin = float(rng.choice([2.5, 600.0, 2349.0]))
        variant = eng.K2_FARFIELD if rng.random() < 0.5 else eng.K2_CLASSED
        split = int(rng.random() < 0.5)
        e.set_k2_variant(variant, 0)
        e.set_option(eng.OPT_SPLIT_TILES, split)
        w = workloads.atmosphere(n_layers=n_layers, n_lines=int(rng.integers(50, 30000)), rmin=rmin, rmax=rmin + n_grid * res,
                                 res=res, top_km=float(rng.uniform(2.0, 60.0)), seed=300 + case)
        sp = w["species"]
        n_total = eng.grid_len(w["range_min"], w["range_max"], w["res"])
        win = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
        qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
        args = (w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp], win,
                w["t_surface"], w["range_max"])
        e.upload_lines(w["lines"], n_groups=len(sp))
        e.set_grid(w["range_min"], w["res"], n_total)
        e.atmosphere(*args)
        whole_r, whole_t = np.empty(n_total, dtype=np.float32), np.empty(n_total, dtype=np.float32)
        e.atmosphere_read_f32(whole_r, whole_t)
        plan = pd.ShardPlan(w["lines"]["nu"], w["range_min"], w["res"], n_total, win, rank, world, farfield=variant == eng.K2_FARFIELD)
        nc = plan.i_end - plan.i_begin
        e.upload_lines(plan.subset(w["lines"]), n_groups=len(sp))
        e.set_grid(w["range_min"], w["res"], n_total, plan.i_begin, plan.i_end)
        pd.connect_peers(e, rank, world, plan.max_chunk, dist)
        for step in range(2):
            e.atmosphere(*args)
            g_r, g_t = pd.gathered_spectra(e)
        full_t = pd.assemble(g_t, plan).cpu().numpy()
        full_r = pd.assemble(g_r, plan).cpu().numpy()
        info = (case, n_layers, res, n_total, plan.chunks, variant, split)
        assert full_t.shape == (n_total,), info
        assert np.array_equal(full_t, whole_t), info
        assert np.array_equal(full_r, whole_r, equal_nan=True), info
        dist.barrier()
        e.peer_disconnect()
        if rank == 0:
            print("fuzz case %d ok: layers %d n %d chunks %s variant %d split %d" % (case, n_layers, n_total, plan.chunks, variant, split), flush=True)
    e.set_k2_variant(eng.K2_CLASSED, 0)
    e.set_option(eng.OPT_SPLIT_TILES, 0)
    if rank == 0:
        print("peer_check ok: world %d" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
