"""Development aid (torchrun, one rank per GPU): where the host-buffer e2e step of the cfg2 cell spends its time when N
ranks share one host -- full call, upload only (results stay on the device), resident inputs with zero-copy results,
resident inputs with a D2H copy of the results."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from pyrad_b200 import engine as eng, workloads
import bench


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    aff = bench.pin_rank_to_gpu_numa(local) if os.environ.get("PROBE_PIN", "1") == "1" else "off"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = workloads.cfg2_shard(rank, world)
    sp = w["species"]
    e = eng.Engine(local)
    T, P = w["T"], w["P"]
    win = eng.window_len(w["cutoff"], w["res"])
    n_chunk = w["i_end"] - w["i_begin"]
    host = {}
    keep = []
    for k, v in w["lines"].items():
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory(); keep.append(t); host[k] = t.numpy()
    h_rad = torch.empty(n_chunk, dtype=torch.float32).pin_memory(); h_tr = torch.empty(n_chunk, dtype=torch.float32).pin_memory()
    cell = e.gas_cell_host_call(host, len(sp), w["range_min"], w["res"], w["n_total"], w["i_begin"], w["i_end"], w["depth_cm"], T, P,
                                w["conc"], [s.molmass for s in sp], [s.q(T) for s in sp], [s.q296 for s in sp], win, 288.0, w["range_max"])
    cell()                                                           # (sets the engine's group count for atmosphere_call)
    col = e.atmosphere_call([w["depth_cm"]], [T], [P], [w["conc"]], [s.molmass for s in sp], [[s.q(T) for s in sp]],
                            [s.q296 for s in sp], [win], 288.0, w["range_max"])

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        e.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        e.synchronize()
        dt = (time.perf_counter() - t0) / reps * 1e3
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {}
    e.set_result_host(h_rad.numpy(), h_tr.numpy())
    res["full (upload + zero-copy results)"] = timed(cell)
    e.set_result_host()
    res["upload only (results stay on device)"] = timed(cell)
    cell()                                                           # lines resident from here on
    e.set_result_host(h_rad.numpy(), h_tr.numpy())
    res["resident inputs, zero-copy results"] = timed(col)
    e.set_result_host()
    res["resident inputs, results on device"] = timed(col)

    def col_copy():
        col(); e.atmosphere_read_f32(h_rad.numpy(), h_tr.numpy())
    res["resident inputs, D2H copy of results"] = timed(col_copy)

    def full_copy():
        cell(); e.atmosphere_read_f32(h_rad.numpy(), h_tr.numpy())
    res["upload + D2H copy of results"] = timed(full_copy)
    if rank == 0:
        print("N=%d affinity %s" % (world, aff))
        for k, v in res.items():
            print("  %-42s %.3f ms (max over ranks)" % (k, v))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
