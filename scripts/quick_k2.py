"""Quick K2 timing sweep (development aid, not the bench): variants x points-per-thread on a cfg2-like cell."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyrad_b200 import engine as eng, workloads

def main():
    n_lines = int(os.environ.get("QK_LINES", 500000))
    rmax = float(os.environ.get("QK_RMAX", 3000.0))
    w = workloads.cfg2(n_lines, rmax)
    if os.environ.get("QK_P"):
        w["P"] = float(os.environ["QK_P"])
        w["T"] = int(os.environ.get("QK_T", 250))
        w["cutoff"] = w["P"] / 1013.25 * 5
    e = eng.Engine(0)
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    t0 = time.time()
    e.upload_lines(w["lines"], len(w["species"]))
    e.set_grid(w["range_min"], w["res"], n)
    print("upload+grid %.3fs n=%d lines=%d" % (time.time() - t0, n, e.n_lines), flush=True)
    sp = w["species"]
    T, P = w["T"], w["P"]
    win = eng.window_len(w["cutoff"], w["res"])
    stream = torch.cuda.ExternalStream(e.stream)
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    wts = [eng.number_density_weight(c, P, T) for c in w["conc"]]
    res = {}
    ref = None
    only = os.environ.get("QK_ONLY")
    combos = [(v, p) for v in (1, 0) for p in (4, 8, 16)]
    if only:
        combos = [tuple(int(x) for x in c.split(":")) for c in only.split(",")]
    for variant, ppt in combos:
        if True:
            if ppt < 0:
                e.set_k2_variant(variant, 0); e.set_narrow_threshold(1 << 20)
            else:
                e.set_k2_variant(variant, ppt); e.set_narrow_threshold(0)
            t0 = time.time()
            e.layer_prepass(T, P, w["conc"], [s.molmass for s in sp], [s.q(T) for s in sp], [s.q296 for s in sp], win, wts)
            e.synchronize()
            tk1 = time.time() - t0
            pairs = e.pair_count()
            for _ in range(2):
                e.line_sum_dev(out.data_ptr(), eng.OUT_F64)
            e.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            reps = 3
            ev[0].record(stream)
            for _ in range(reps):
                e.line_sum_dev(out.data_ptr(), eng.OUT_F64)
            ev[1].record(stream)
            e.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / reps
            o = out.cpu().numpy()
            if ref is None:
                ref = o
            d = np.abs(o - ref) / np.maximum(np.abs(ref), 1e-300)
            print("variant %d ppt %2d: %.3f ms  %.3e pairs/s  (k1 %.1f ms, pairs %.3e) maxrel-vs-first %.2e" %
                  (variant, ppt, ms, pairs / ms * 1e3, tk1 * 1e3, pairs, d.max()), flush=True)
            res["v%d_p%d" % (variant, ppt)] = {"ms": ms, "pairs_per_s": pairs / ms * 1e3}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/quick_k2.json", "w"), indent=1)

if __name__ == "__main__":
    main()
