"""Development aid: the bench's headline step (cfg2, resident inputs, L2 flushed) timed per stage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyrad_b200 import engine as eng, workloads


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    w = workloads.cfg2_shard(0, 1)
    sp = w["species"]
    e = eng.Engine(0)
    e.upload_lines(w["lines"], n_groups=len(sp))
    e.set_grid(w["range_min"], w["res"], w["n_total"], w["i_begin"], w["i_end"])
    T, P = w["T"], w["P"]
    win = eng.window_len(w["cutoff"], w["res"])
    ext = torch.cuda.ExternalStream(e.stream)
    torch.cuda.set_stream(ext)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    call = e.atmosphere_call([w["depth_cm"]], [T], [P], [w["conc"]], [s.molmass for s in sp], [[s.q(T) for s in sp]],
                             [s.q296 for s in sp], [win], w["t_surface"], w["range_max"])
    for timing in (False, True):
        e.set_timing(timing)
        for _ in range(5):
            flush.zero_(); call()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        k1 = k2 = 0.0
        for a, b in evs:
            flush.zero_()
            a.record(ext); call(); b.record(ext)
            if timing:
                t = e.atmosphere_timing(); k1 += t["k1_ms"]; k2 += t["k2_ms"]
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        print("timing=%s steps %d: mean %.4f median %.4f min %.4f max %.4f ms  k1 %.4f k2 %.4f" %
              (timing, steps, ms.mean(), np.median(ms), ms.min(), ms.max(), k1 / steps, k2 / steps))


if __name__ == "__main__":
    main()
