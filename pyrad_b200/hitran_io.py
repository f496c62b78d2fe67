"""Readers for the reference's on-disk formats (cwd-relative ``data/`` tree), so the host mirror in
``classes.py`` can run on the very same files the reference reads.  No network access here: a missing
segment file is an error, not a download.

Formats (reference pyradUtilities.py):
  data/<globalIso>/<segment>.pyr   HITRAN-online CSV rows  molec,iso,nu,sw,a,elower,gamma_air,gamma_self,
                                   delta_air,n_air  (readHitranOnlineFile :421-448, gatherData :173-189)
  data/<globalIso>/q<globalIso>.txt  rows "T Q"               (readQFile :451-461)
  data/<globalIso>/params.pyr      one CSV row                (readMolParams :464-477)
  data/xsc/<NAME>/<file>.txt       two whitespace columns     (returnXscFileContents :680-696)
"""
import os
import re

import numpy as np

NULL_TAG = "#/null/#"
LINE_COLUMNS = ("nu", "sw", "a", "elower", "gamma_air", "gamma_self", "delta_air", "n_air")


def data_dir(root=None):
    return os.path.join(root if root is not None else os.getcwd(), "data")


def _rows(path):
    """Lines of a data file with leading '#' comment rows dropped; None when absent / tagged empty."""
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        rows = f.readlines()
    if not rows or NULL_TAG in rows[0]:
        return None
    while rows and rows[0].startswith("#") and len(rows) > 1:
        rows.pop(0)
    return rows


def gather_lines(global_iso, wave_min, wave_max, root=None):
    """All lines with wave_min < nu < wave_max (strict, as the reference) from the 100 cm-1 segment files,
    as SoA float64 arrays ascending in nu.  Duplicate wavenumbers collapse, last one wins (the reference
    keys a dict by nu)."""
    seg = int(wave_min / 100) * 100
    found = {}
    while seg < wave_max:
        path = os.path.join(data_dir(root), str(global_iso), "%d.pyr" % seg)
        if not os.path.isfile(path):
            raise FileNotFoundError("%s is missing (pyrad_b200 never downloads; seed the data tree first)" % path)
        rows = _rows(path)
        for row in rows or ():
            c = row.split(",")
            if len(c) < 10 or row.startswith("#"):
                continue
            nu = float(c[2])
            if wave_min < nu < wave_max:
                found[nu] = (float(c[3]), float(c[4]), float(c[5]), float(c[6]), float(c[7]), float(c[8]), float(c[9]))
        seg += 100
    nus = list(found.keys())                      # insertion order == file order (ascending), like the reference
    vals = np.array([found[n] for n in nus], dtype=np.float64).reshape(len(nus), 7)
    out = {"nu": np.array(nus, dtype=np.float64)}
    for j, k in enumerate(LINE_COLUMNS[1:]):
        out[k] = np.ascontiguousarray(vals[:, j]) if len(nus) else np.zeros(0)
    return out


def read_q_table(global_iso, root=None):
    path = os.path.join(data_dir(root), str(global_iso), "q%s.txt" % global_iso)
    q = {}
    with open(path) as f:
        for row in f:
            c = row.split()
            if len(c) >= 2:
                q[int(c[0])] = float(c[1])
    return q


def read_mol_params(global_iso, root=None):
    path = os.path.join(data_dir(root), str(global_iso), "params.pyr")
    rows = _rows(path)
    if not rows:
        raise FileNotFoundError(path)
    c = rows[0].split(",")
    return {"globalIso": int(c[0]), "shortName": c[1], "molNum": int(c[2]), "isoN": int(c[3]),
            "abundance": float(c[4]), "q296": float(c[5]), "gj": int(c[6]), "molmass": float(c[7])}


_RE = {
    "TEMP": re.compile(r"[0-9.]*(?=K)"),
    "PRESSURE": re.compile(r"[0-9.]*(?=Torr)"),
    "MOLECULE_SHORT_NAME": re.compile(r"^[A-Za-z0-9]*"),
    "RANGE": re.compile(r"(?<=_)[0-9.]*-[0-9.]*(?=_)"),
    "RES": re.compile(r"(?<=_)[0-9]{1,}.[0-9]{1,}(?=_)"),
}


def parse_xsc_filename(filename):
    """TEMP / PRESSURE[Torr] / RANGE / RES fields out of an xsc file name (parseXscFileName :611-641)."""
    stem = re.sub(".txt", "", filename)
    out = {}
    for k, rx in _RE.items():
        m = rx.search(stem)
        out[k] = m.group(0) if m else False
    return out


def read_xsc_table(name, filename, root=None):
    path = os.path.join(data_dir(root), "xsc", name, filename)
    rows = _rows(path)
    if rows is None:
        raise FileNotFoundError(path)
    wn, xs = [], []
    for row in rows:
        parts = re.split("[ ]+", row.strip())
        if len(parts) == 2:
            try:
                a, b = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            wn.append(a)
            xs.append(b)
    return np.array(wn), np.array(xs)
