"""Ingestion timing (development aid): PRB_INGEST_TIMING=1 python scripts/quick_ingest.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyrad_b200 import engine as eng, synth
n = int(os.environ.get("QI_LINES", 500000))
ln = synth.make_lines(n, 0.0, 3000.0, 5)
rows = ["2,1,%r,%r,1.0,%r,%r,%r,%r,%r" % tuple(float(ln[k][j]) for k in ("nu", "sw", "elower", "gamma_air", "gamma_self", "delta_air", "n_air")) for j in range(n)]
text = ("\n".join(rows) + "\n").encode()
e = eng.Engine(0)
for i in range(3):
    t0 = time.perf_counter(); k = e.ingest_csv(text, -1.0, 1e9); dt = time.perf_counter() - t0
    print("ingest %d rows, %.1f MB: %.2f ms (%.2f GB/s, %.2e lines/s)" % (k, len(text) / 1e6, dt * 1e3, len(text) / dt / 1e9, k / dt), flush=True)
