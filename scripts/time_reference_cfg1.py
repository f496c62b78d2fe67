"""The REAL, UNMODIFIED reference timed on BASELINE cfg1: Layer(10 cm, 296 K, 1013 hPa, 500-800 cm-1)
+ addMolecule('co2', ppm=400) on the seeded synthetic data tree, getTransmittance(layer) -- i.e. Isotope.createCrossSection's
per-element Python loop over all ~50 k lines (pyradClasses.py:361-407).  Only runs where /root/reference is mounted (the build
container); the result is committed as tests/golden/reference_cfg1_timing.json and quoted by bench.py's cfg1 object, labelled
with the machine it was measured on.  Usage: python scripts/time_reference_cfg1.py [n_lines]"""
import json, os, platform, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import physics as ph, ref_harness as rh
from pyrad_b200 import synth, workloads


def main():
    n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
    w = workloads.cfg1(n_lines)
    sp = w["species"][0]
    work = tempfile.mkdtemp(prefix="ref_cfg1_")
    rh.seed_workdir(work)
    ln = w["per_group_lines"][0]
    rh.write_params(work, sp.global_iso, sp.name, sp.mol_id, 1, 0.99, sp.q296, 1, sp.molmass)
    rh.write_q_table(work, sp.global_iso, range(100, 501), [sp.q(t) for t in range(100, 501)])
    rh.write_line_segments(work, sp.global_iso, sp.mol_id, 1, ln, int(ln["nu"].min() / 100) * 100, ln["nu"].max() + 101)
    ref = rh.load_reference(work)
    C = ref.classes
    os.chdir(work)
    with rh.quiet():
        t0 = time.perf_counter()
        layer = C.Layer(w["depth_cm"], w["T"], w["P"], w["range_min"], w["range_max"])
        mol = layer.addMolecule("co2", ppm=w["conc"][0] * 1e6)
        t_read = time.perf_counter() - t0
        t0 = time.perf_counter()
        tr = np.asarray(C.getTransmittance(layer))
        t_compute = time.perf_counter() - t0
    kept = len(mol[0])
    n = ph.grid_len(w["range_min"], w["range_max"], w["res"])
    lines_kept = {k: v[(ln["nu"] > layer.effectiveRangeMin) & (ln["nu"] < layer.effectiveRangeMax)] for k, v in ln.items()}
    pairs = ph.pair_count(ph.line_index(lines_kept["nu"], w["range_min"], w["res"]), n, ph.window_len(w["cutoff"], w["res"]))
    # the oracle on the same inputs, for the record (parity of the timed run itself)
    sig = ph.cross_section(lines_kept, w["T"], w["P"], w["conc"][0], sp.molmass, sp.q(w["T"]), sp.q296, w["range_min"],
                           w["range_max"], w["res"], w["cutoff"])
    t_ref = ph.transmittance(ph.abs_coef(sig, w["conc"][0], w["P"], w["T"]), w["depth_cm"])
    out = {"workload": "cfg1: CO2 cell 10 cm, 296 K, 1013 hPa, 400 ppm, 500-800 cm-1 @ 0.01 cm-1, %d synthetic lines (%d kept)" % (n_lines, kept),
           "what": "real unmodified bschrag620/PyRad: Layer + addMolecule (file read) then getTransmittance(layer) "
                   "(Isotope.createCrossSection, pyradClasses.py:361-407)",
           "read_lines_s": t_read, "get_transmittance_s": t_compute, "pairs": int(pairs),
           "pairs_per_s": pairs / t_compute, "cores": 1,
           "machine": "%s, %d logical CPUs (build container, not the GPU box)" % (platform.processor() or platform.machine(), os.cpu_count()),
           "python": platform.python_version(), "numpy": np.__version__,
           "max_abs_T_diff_oracle_vs_reference": float(np.abs(tr - t_ref).max())}
    print(json.dumps(out, indent=1))
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_cfg1_timing.json")
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
