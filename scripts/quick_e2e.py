"""e2e timing breakdown of the pipelined gas cell (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyrad_b200 import engine as eng, workloads
w = workloads.cfg2_shard(0, 1)
sp = w["species"]; e = eng.Engine(0)
T, P = w["T"], w["P"]; win = eng.window_len(w["cutoff"], w["res"])
mol = [s.molmass for s in sp]; q296 = [s.q296 for s in sp]; qt = [s.q(T) for s in sp]
host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in w["lines"].items()}
lines = {k: v.numpy() for k, v in host.items()}
n = w["i_end"] - w["i_begin"]
h_rad = torch.empty(n, dtype=torch.float32).pin_memory(); h_tr = torch.empty(n, dtype=torch.float32).pin_memory()
def pipelined():
    e.gas_cell_host(lines, len(sp), w["range_min"], w["res"], w["n_total"], w["i_begin"], w["i_end"], w["depth_cm"], T, P, w["conc"], mol, qt, q296, win, w["t_surface"], w["range_max"])
def separate():
    e.upload_lines(lines, len(sp)); e.set_grid(w["range_min"], w["res"], w["n_total"], w["i_begin"], w["i_end"])
    e.atmosphere([w["depth_cm"]], [T], [P], [w["conc"]], mol, [qt], q296, [win], w["t_surface"], w["range_max"])
def resident():
    e.atmosphere([w["depth_cm"]], [T], [P], [w["conc"]], mol, [qt], q296, [win], w["t_surface"], w["range_max"])
for host_dst in (False, True):
    if host_dst: e.set_result_host(h_rad.numpy(), h_tr.numpy())
    for name, fn in (("pipelined", pipelined), ("separate", separate), ("resident", resident)):
        for _ in range(3): fn()
        t0 = time.perf_counter()
        for _ in range(20): fn()
        print("host_dst=%s %-10s %.3f ms" % (host_dst, name, (time.perf_counter() - t0) / 20 * 1e3), flush=True)
