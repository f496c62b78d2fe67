"""Development aid: ONE far-field K2 launch of a chosen workload, for ncu.  Usage: python scripts/prof_far.py cfg5|cfg2|atm0 [variant]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyrad_b200 import engine as eng, workloads


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    e = eng.Engine(0)
    if what == "cfg2":
        w = workloads.cfg2()
    elif what == "cfg5":
        w = workloads.cfg5()
    elif what.startswith("p"):
        w = workloads.cfg5(cutoff=5.0 * float(what[1:]) / 1013.25)   # a cfg4 layer of that pressure [hPa] on the cfg4 line density
        w["P"], w["T"] = float(what[1:]), 230
    else:
        w = workloads.cfg5(cutoff=5.0 * 971.9 / 1013.25)          # the widest cfg4 layer's window on the cfg4 line density
    sp = w["species"]
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    e.upload_lines(w["lines"], len(sp)); e.set_grid(w["range_min"], w["res"], n)
    T, P = w["T"], w["P"]
    win = eng.window_len(w["cutoff"], w["res"])
    wts = [eng.number_density_weight(c, P, T) for c in w["conc"]]
    e.set_k2_variant(variant, 0)
    e.layer_prepass(T, P, w["conc"], [s.molmass for s in sp], [s.q(T) for s in sp], [s.q296 for s in sp], win, wts)
    out = e.line_sum()
    print(what, variant, float(out.sum()))


if __name__ == "__main__":
    main()
