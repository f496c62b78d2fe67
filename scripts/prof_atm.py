"""Development aid: ONE cfg4 column (prb_atmosphere, 100 layers) for ncu.  Usage: python scripts/prof_atm.py [variant] [layers]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyrad_b200 import engine as eng, workloads


def main():
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    w = workloads.atmosphere(n_layers=layers)
    sp = w["species"]
    e = eng.Engine(0)
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    e.upload_lines(w["lines"], len(sp)); e.set_grid(w["range_min"], w["res"], n)
    win = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
    qt = np.array([[s.q(T) for s in sp] for T in w["T"]])
    e.set_k2_variant(variant, 0)
    e.atmosphere(w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp], win,
                 w["t_surface"], w["range_max"])
    rad, tr = e.atmosphere_read()
    print("atm", variant, layers, float(np.nansum(rad)))


if __name__ == "__main__":
    main()
