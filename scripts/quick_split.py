"""Development aid: one rank's share of the cfg5 stress sweep at N = 8 (306 tiles of 2048 points on 296 CTA slots) with
and without PRB_OPT_SPLIT_TILES, exact and far-field K2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyrad_b200 import engine as eng, workloads, distributed as pd


def main():
    w = workloads.cfg5()
    sp = w["species"]
    e = eng.Engine(0)
    stream = torch.cuda.ExternalStream(e.stream)
    n = eng.grid_len(w["range_min"], w["range_max"], w["res"])
    win = eng.window_len(w["cutoff"], w["res"])
    rank = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    far_plan = len(sys.argv) > 2 and sys.argv[2] == "far"
    plan = pd.ShardPlan(w["lines"]["nu"], w["range_min"], w["res"], n, [win], rank, 8, farfield=far_plan)
    print("rank", rank, "plan", "far-field" if far_plan else "exact", "chunks", [(b - a + 2047) // 2048 for a, b in plan.chunks])
    e.upload_lines(plan.subset(w["lines"]), len(sp)); e.set_grid(w["range_min"], w["res"], n, plan.i_begin, plan.i_end)
    T, P = w["T"], w["P"]
    args = ([w["depth_cm"]], [T], [P], [w["conc"]], [s.molmass for s in sp], [[s.q(T) for s in sp]], [s.q296 for s in sp],
            [win], 288.0, w["range_max"])
    print("chunk", plan.i_begin, plan.i_end, "tiles", (plan.i_end - plan.i_begin + 2047) // 2048)
    ref = {}
    e.set_timing(True)
    for variant in (1, 2):
        for split in (0, 1):
            e.set_k2_variant(variant, 0); e.set_option(eng.OPT_SPLIT_TILES, split)
            e.atmosphere(*args); e.synchronize()
            ms = []
            for _ in range(5):
                with torch.cuda.stream(stream):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream); e.atmosphere(*args); b.record(stream)
                e.synchronize(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            rad, tr = e.atmosphere_read()
            ref.setdefault(variant, tr)
            print("variant", variant, "split", split, "ms %.3f" % np.median(ms), e.atmosphere_timing(), "max |dT| vs unsplit %.2e" % np.nanmax(np.abs(tr - ref[variant])))
    e.set_option(eng.OPT_SPLIT_TILES, 0); e.set_k2_variant(1, 0)


if __name__ == "__main__":
    main()
