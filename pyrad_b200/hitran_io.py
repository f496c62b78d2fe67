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


def segment_paths(global_iso, wave_min, wave_max, root=None):
    """The 100 cm-1 segment files gatherData walks (pyradUtilities.py:173-187): int(min/100)*100, +100, ... < max."""
    seg = int(wave_min / 100) * 100
    out = []
    while seg < wave_max:
        out.append(os.path.join(data_dir(root), str(global_iso), "%d.pyr" % seg))
        seg += 100
    return out


def gather_text(global_iso, wave_min, wave_max, root=None):
    """Raw bytes of the segment files, concatenated in segment order, for the device parser
    (prb_ingest_hitran_csv).  A file that is absent is an error (pyrad_b200 never downloads); a file tagged
    empty by the reference (NULL_TAG in its first row, openReturnLines :98) contributes nothing."""
    parts = []
    for path in segment_paths(global_iso, wave_min, wave_max, root):
        if not os.path.isfile(path):
            raise FileNotFoundError("%s is missing (pyrad_b200 never downloads; seed the data tree first)" % path)
        with open(path, "rb") as f:
            blob = f.read()
        head = blob.split(b"\n", 1)[0]
        if not blob or NULL_TAG.encode() in head:
            continue
        if not blob.endswith(b"\n"):
            blob += b"\n"
        parts.append(blob)
    return b"".join(parts)


def gather_lines(global_iso, wave_min, wave_max, root=None, engine=None):
    """All lines with wave_min < nu < wave_max (strict, as the reference) from the 100 cm-1 segment files, as SoA
    float64 arrays in file order; duplicate wavenumbers collapse, last one wins (the reference keys a dict by nu).
    The text is parsed ON THE DEVICE (K5, exact decimal -> double conversion); the columns come back for the
    host-side Line objects and stay resident as the engine's line list."""
    if engine is None:
        raise ValueError("gather_lines needs the CUDA engine (there is no host parser in this package)")
    engine.ingest_csv(gather_text(global_iso, wave_min, wave_max, root), wave_min, wave_max)
    return engine.download_lines()


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


def read_xsc_table(name, filename, root=None, engine=None):
    """(wavenumber, cross section) of an xsc table file (returnXscFileContents :680-696), parsed on the device."""
    path = os.path.join(data_dir(root), "xsc", name, filename)
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    if engine is None:
        raise ValueError("read_xsc_table needs the CUDA engine (there is no host parser in this package)")
    with open(path, "rb") as f:
        blob = f.read()
    head = blob.split(b"\n", 1)[0]
    if not blob or NULL_TAG.encode() in head:
        raise FileNotFoundError(path)
    return engine.parse_xsc_text(blob)
