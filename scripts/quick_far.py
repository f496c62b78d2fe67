"""Development aid: exact (PRB_K2_CLASSED) versus far-field (PRB_K2_FARFIELD) K2 on cfg2 and on the cfg4 atmosphere --
times and differences.  Usage: [PRB_LIB=...] python scripts/quick_far.py <tag>"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyrad_b200 import engine as eng, workloads


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "x"
    variants = [1, 2] if os.environ.get("QF_FAR", "1") == "1" else [1]
    res = {}
    e = eng.Engine(0)
    stream = torch.cuda.ExternalStream(e.stream)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    # ---- cfg2, K2 alone
    w = workloads.cfg2()
    sp = w["species"]
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    e.upload_lines(w["lines"], len(sp)); e.set_grid(w["range_min"], w["res"], n)
    T, P = w["T"], w["P"]
    win = eng.window_len(w["cutoff"], w["res"])
    wts = [eng.number_density_weight(c, P, T) for c in w["conc"]]
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    ref = None
    for v in variants:
        e.set_k2_variant(v, 0)
        e.layer_prepass(T, P, w["conc"], [s.molmass for s in sp], [s.q(T) for s in sp], [s.q296 for s in sp], win, wts)
        pairs = e.pair_count()
        ms = []
        for i in range(3 + 10):
            with torch.cuda.stream(stream):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                e.line_sum_dev(out.data_ptr(), eng.OUT_F64)
                b.record(stream)
            e.synchronize(); torch.cuda.synchronize()
            if i >= 3:
                ms.append(a.elapsed_time(b))
        o = out.cpu().numpy()
        if ref is None:
            ref = o
        d = np.abs(o - ref) / np.maximum(np.abs(ref), 1e-40 * np.abs(ref).max())
        res["cfg2_v%d" % v] = {"k2_ms": float(np.mean(ms)), "k2_ms_min": float(np.min(ms)), "pairs_per_s": pairs / np.mean(ms) * 1e3,
                               "max_rel_vs_exact": float(d.max())}
        print(tag, "cfg2 variant", v, res["cfg2_v%d" % v], flush=True)
    # ---- cfg4 atmosphere
    if os.environ.get("QF_ATM", "1") == "1":
        w = workloads.atmosphere(n_layers=100, n_lines=5_000_000)
        sp = w["species"]
        n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
        e.upload_lines(w["lines"], len(sp)); e.set_grid(w["range_min"], w["res"], n)
        winl = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
        qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
        args = (w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp], winl,
                w["t_surface"], w["range_max"])
        e.set_timing(True)
        ref = None
        for v in variants:
            e.set_k2_variant(v, 0)
            e.atmosphere(*args); e.synchronize()
            ts = []
            for _ in range(3):
                e.atmosphere(*args); e.synchronize()
                ts.append(dict(e.atmosphere_timing()))
            rad = np.empty(n, dtype=np.float32); tr = np.empty(n, dtype=np.float32)
            e.atmosphere_read_f32(rad, tr)
            if ref is None:
                ref = (rad.copy(), tr.copy())
            ok = np.isfinite(ref[0]) & (ref[0] != 0)
            fin = np.isfinite(ref[1]) & np.isfinite(tr)      # next to 0 cm-1 the reference's negative Doppler widths give inf
            res["atm_v%d" % v] = {"k2_ms": float(np.median([t["k2_ms"] for t in ts])), "k1_ms": float(np.median([t["k1_ms"] for t in ts])),
                                  "max_abs_T_vs_exact": float(np.abs(tr[fin] - ref[1][fin]).max()),
                                  "max_rel_rad_vs_exact": float((np.abs(rad - ref[0])[ok] / np.abs(ref[0][ok])).max())}
            print(tag, "atm variant", v, res["atm_v%d" % v], flush=True)
    # ---- cfg5 stress sweep (W = 25 000), whole call
    if os.environ.get("QF_STRESS", "1") == "1":
        w = workloads.cfg5()
        sp = w["species"]
        n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
        e.upload_lines(w["lines"], len(sp)); e.set_grid(w["range_min"], w["res"], n)
        T, P = w["T"], w["P"]
        args = ([w["depth_cm"]], [T], [P], [w["conc"]], [s.molmass for s in sp], [[s.q(T) for s in sp]], [s.q296 for s in sp],
                [eng.window_len(w["cutoff"], w["res"])], 288.0, w["range_max"])
        ref = None
        for v in variants:
            e.set_k2_variant(v, 0)
            e.atmosphere(*args); e.synchronize()
            ms = []
            for _ in range(3):
                with torch.cuda.stream(stream):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    e.atmosphere(*args)
                    b.record(stream)
                e.synchronize(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            rad = np.empty(n, dtype=np.float32); tr = np.empty(n, dtype=np.float32)
            e.atmosphere_read_f32(rad, tr)
            if ref is None:
                ref = tr.copy()
            fin = np.isfinite(ref) & np.isfinite(tr)
            res["cfg5_v%d" % v] = {"ms": float(np.median(ms)), "max_abs_T_vs_exact": float(np.abs(tr[fin] - ref[fin]).max())}
            print(tag, "cfg5 variant", v, res["cfg5_v%d" % v], flush=True)
    e.set_k2_variant(1, 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/quick_far_%s.json" % tag, "w"), indent=1)


if __name__ == "__main__":
    main()
