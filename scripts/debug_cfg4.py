"""Development aid: cfg4 column on the GPU, spectra and the k-matrix columns at the test's sample points -> gpurun_out/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyrad_b200 import engine as eng, workloads, distributed as pd
from tests import helpers as H


def main():
    w = workloads.atmosphere()
    sp = w["species"]
    e = eng.Engine(0)
    n = H.engine_setup(e, w)
    win = [eng.window_len(c, w["res"]) for c in w["cutoff"]]
    qt = np.array([[s.q(t) for s in sp] for t in w["T"]])
    e.atmosphere(w["depth_cm"], w["T"], w["P"], w["conc"], [s.molmass for s in sp], qt, [s.q296 for s in sp], win,
                 w["t_surface"], w["range_max"])
    rad, tr = e.atmosphere_read()
    pts = H.boundary_points(n, 200, 44, n_tiles=16)
    kp, ld = e.atmosphere_kmatrix_dev()
    km = pd.device_tensor(kp, ld * 100).view(100, ld)[:, torch.as_tensor(pts, device="cuda")].cpu().numpy()
    np.savez("gpurun_out/debug_cfg4.npz", pts=pts, rad=rad[pts], tr=tr[pts], kmat=km)


if __name__ == "__main__":
    main()
