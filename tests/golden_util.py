"""Loading the committed golden vectors (generated from the REAL reference by tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CELL_CASES = ["cell_co2_1atm", "cell_lowp", "cell_tiny_window", "cell_fine_grid", "cell_highp_dynres"]
XSC_CASES = ["xsc_native_res", "xsc_coarse_res"]
LINE_KEYS = ("nu", "sw", "a", "elower", "gamma_air", "gamma_self", "delta_air", "n_air")


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def lines_of(g, group, cutoff=None, rmin=None, rmax=None):
    """Line columns of isotopologue `group` as written to the reference's data tree; optionally only the
    lines the reference keeps (effMin < nu < effMax strictly, pyradUtilities.py:437-438)."""
    ln = {k: np.asarray(g["lines%d_%s" % (group, k)]) for k in LINE_KEYS}
    if cutoff is not None:
        lo, hi = max(rmin - cutoff, 0), rmax + cutoff
        m = (ln["nu"] > lo) & (ln["nu"] < hi)
        ln = {k: v[m] for k, v in ln.items()}
    return ln


def files_of(g, prefix):
    """{file name: bytes} of a folder stored by make_golden._pack_files."""
    names = [str(n) for n in g[prefix + "_names"]]
    blob = g[prefix + "_blob"].tobytes()
    offs = g[prefix + "_offsets"]
    return {n: blob[int(offs[i]):int(offs[i + 1])] for i, n in enumerate(names)}
