"""Development aid: per-call times of the mirror's cold getTransmittance with the page-locked result pool."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyrad_b200 import classes as C, engine as eng, workloads


def main():
    w = workloads.cfg2()
    root = tempfile.mkdtemp()
    for s in w["species"]:
        d = os.path.join(root, "data", str(s.global_iso)); os.makedirs(d)
        open(os.path.join(d, "params.pyr"), "w").write("# params\n%d,%s,%d,1,0.99,%r,1,%r\n" % (s.global_iso, s.name, s.mol_id, float(s.q296), float(s.molmass)))
    C.DATA_ROOT, C.BASE_RESOLUTION, C.Layer.hasAtmosphere = root, w["res"], False
    layer = C.Layer(w["depth_cm"], w["T"], w["P"], w["range_min"], w["range_max"], dynamicResolution=False)
    for g, (s, c) in enumerate(zip(w["species"], w["conc"])):
        m = C.Molecule(s.name, layer, concentration=c); layer.append(m)
        m[0].setLines({k: np.array(v) for k, v in w["per_group_lines"][g].items()}, {int(w["T"]): s.q(w["T"])})
    pool = C.engine().result_pool
    tr = None
    for i in range(8):
        t0 = time.perf_counter()
        C.resetCrossSection(layer); C._RESIDENT_KEY = None
        tr = C.getTransmittance(layer)
        print("call %d: %.2f ms  pool free %s held %.0f MB  pinned=%s" % (i, (time.perf_counter() - t0) * 1e3,
              {k >> 20: len(v) for k, v in pool.free.items()}, pool.held / 2**20, not tr.flags.owndata))
    for i in range(3):
        t0 = time.perf_counter()
        rows = [C.getCrossSection(m[0]) for m in layer]
        print("rows %d: %.2f ms" % (i, (time.perf_counter() - t0) * 1e3))
        C.resetCrossSection(layer); C._RESIDENT_KEY = None
        tr = C.getTransmittance(layer)
        del rows


if __name__ == "__main__":
    main()
