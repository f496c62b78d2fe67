"""Per-layer K1/K2 timing of the cfg4 atmosphere (development aid)."""
import os, sys, time, json, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyrad_b200 import engine as eng, workloads, partition as pt
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler

def main():
    L = int(os.environ.get("QA_LAYERS", 100)); nl = int(os.environ.get("QA_LINES", 5000000))
    w = workloads.atmosphere(n_layers=L, n_lines=nl)
    sp = w["species"]
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    e = eng.Engine(0)
    e.upload_lines(w["lines"], len(sp)); e.set_grid(w["range_min"], w["res"], n)
    win = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
    qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
    e.set_timing(True)
    if os.environ.get("QA_NOTMAFOLD"):
        e.set_option(eng.OPT_FOLD_TMA, 0)
    if os.environ.get("QA_OLDNARROW"):
        e.set_option(eng.OPT_POINT_KERNEL, 0)
    if os.environ.get("QA_NARROW"):
        e.set_narrow_threshold(int(os.environ["QA_NARROW"]))
    if os.environ.get("QA_VARIANT"):
        e.set_k2_variant(int(os.environ["QA_VARIANT"]), 0)
    if os.environ.get("QA_NOBATCH"):
        e.set_option(eng.OPT_BATCH_LAYERS, 0)      # per-layer launches: exact per-layer timings
    args = (w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp], win, w["t_surface"], w["range_max"])
    e.atmosphere(*args)
    cs = ClockSampler(0); cs.start()
    t0 = time.time(); e.atmosphere(*args); e.atmosphere(*args); dt = (time.time() - t0) / 2
    clocks = cs.stop()
    k1, k2 = e.atmosphere_layer_timing(L)
    idx = pt.line_index(w["lines"]["nu"], w["range_min"], w["res"])
    print("total %.1f ms  %s  clocks %s" % (dt * 1e3, e.atmosphere_timing(), clocks))
    for l in range(L):
        if l < 12 or l % 10 == 0 or os.environ.get("QA_ALL"):
            pairs = pt.block_pair_cost(idx, n, [win[l]]).sum()
            print("layer %3d P=%8.3f W=%5d k1 %.3f ms k2 %.3f ms pairs %.3e  %.3e pairs/s" % (l, w["P"][l], win[l], k1[l], k2[l], pairs, pairs / (k2[l] * 1e-3)))

if __name__ == "__main__":
    main()
