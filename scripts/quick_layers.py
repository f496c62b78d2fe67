"""Development aid: K2 alone on single cfg4-density layers of given pressures [hPa], exact and far-field."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyrad_b200 import engine as eng, workloads


def main():
    tag = sys.argv[1]
    pressures = [float(x) for x in sys.argv[2:]] or [560.0, 250.0, 100.0, 60.0, 30.0]
    e = eng.Engine(0)
    stream = torch.cuda.ExternalStream(e.stream)
    base = workloads.cfg5(cutoff=5.0)
    sp = base["species"]
    n = eng.grid_len(base["range_min"], base["range_max"], base["res"])
    e.upload_lines(base["lines"], len(sp)); e.set_grid(base["range_min"], base["res"], n)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    for P in pressures:
        T = 230
        win = eng.window_len(5.0 * P / 1013.25, base["res"])
        wts = [eng.number_density_weight(c, P, T) for c in base["conc"]]
        row = []
        for v in (1, 2):
            e.set_k2_variant(v, 0)
            e.layer_prepass(T, P, base["conc"], [s.molmass for s in sp], [s.q(T) for s in sp], [s.q296 for s in sp], win, wts)
            ms = []
            for i in range(2 + 5):
                with torch.cuda.stream(stream):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream); e.line_sum_dev(out.data_ptr(), eng.OUT_F32); b.record(stream)
                e.synchronize(); torch.cuda.synchronize()
                if i >= 2:
                    ms.append(a.elapsed_time(b))
            row.append(float(np.median(ms)))
        print("%s P=%7.1f W=%5d  exact %.3f ms  far %.3f ms" % (tag, P, win, row[0], row[1]), flush=True)
    e.set_k2_variant(1, 0)


if __name__ == "__main__":
    main()
