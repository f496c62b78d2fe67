"""The reference's xsc FILE utilities on top of the CUDA engine, same names and
argument meaning as pyradUtilities.py so the download-time callers (unzipFile :258-272, pyradInteractive.py:515-517)
can be pointed here:

  changeResXscFile(filepath)   :515-534   table -> np.interp onto arange(min, max, BASE_RESOLUTION) -> rewritten file
  changeResFolder(folderName)  :505-512
  writeXscFile(...)            :537-546   text format of a pyrad-adjusted table
  returnMatchingXsc(...)       :601-608
  mergeXsc(folder)             :549-597   files of equal T and P summed onto the union range, rewritten

What runs where: the table text is parsed on the device (prb_parse_xsc_text, K5), the re-gridding is numpy's interp
arithmetic on the device (prb_xsc_place, interp = 1), the aligned placement of every member table and their sum in
listing order are device work too (prb_xsc_place, interp = 0; prb_layer_stream with unit weights adds the rows in
the reference's order starting from 0.0).  The host only lists folders, parses file NAMES, and turns the finished
float64 columns into text.  The written files are byte-identical to the reference's (tests/golden/xsc_files.npz).
There is no CPU path: every function needs the engine.

Deviations: a table row that is not exactly two numbers is skipped (the reference's `except` branch raises
NameError on its own undefined `targetFile`, :691); partially overlapping merges cannot occur (the target range is the
union), and a member whose samples are not on the target's 0.01 grid raises ValueError as the reference's
list.index does.
"""
import os
import re

import numpy as np

from . import classes as _cls
from . import hitran_io as _io

_RX = {
    "BROADENER": re.compile(r"(?<=_)[A-Za-z0-9]*(?=_[0-9]*_[0-9]*$)"),
    "ID": re.compile(r"(?<=_)[0-9]*_[0-9]*$"),
}


def parseXscFileName(file):
    """parseXscFileName :611-641 -- hitran_io's fields plus BROADENER, ID and the two file-name spellings."""
    stem = re.sub(".txt", "", file)
    out = _io.parse_xsc_filename(file)
    b = _RX["BROADENER"].search(stem)
    i = _RX["ID"].search(stem)
    out["BROADENER"] = b.group(0) if b and b.group(0) else ""
    out["ID"] = i.group(0).replace("_", "-") if i else False
    out["SHORT_FILENAME"] = stem
    out["LONG_FILENAME"] = stem + ".txt"
    return out


def _xsc_dir(folder):
    return os.path.join(_io.data_dir(_cls.DATA_ROOT), "xsc", folder)


def _read_table(filepath):
    with open(filepath, "rb") as f:
        blob = f.read()
    return _cls.engine().parse_xsc_text(blob)


def writeXscFile(wavenumbers, crossSection, rangeMin, rangeMax, temp, pressure, molName, pathToFolder, broadener, i):
    """:537-546.  The header comment carries no newline, so the first row is glued to it -- kept, because the
    reference's reader then drops that row and downstream results depend on it."""
    filename = "%s_%sK-%sTorr_%s-%s_%s_%s_%s.txt" % (molName, temp, pressure, rangeMin, rangeMax, _cls.BASE_RESOLUTION,
                                                   broadener, i.replace("-", "_"))
    w = np.asarray(wavenumbers, dtype=np.float64)
    c = np.asarray(crossSection, dtype=np.float64)
    rows = ["%s     %s\n" % wc for wc in zip(w, c)]            # '%s' of numpy float64 scalars, as the reference prints them
    with open(os.path.join(pathToFolder, filename), "wb") as f:
        f.write(("# pyrad adjusted cross-section file" + "".join(rows)).encode("utf-8"))
    return filename


def changeResXscFile(filepath):
    """:515-534.  Returns the new file's name (the reference prints it)."""
    props = parseXscFileName(os.path.basename(filepath))
    rmin, rmax = (float(v) for v in props["RANGE"].split("-"))
    wn, xs = _read_table(filepath)
    if wn.size == 0:
        raise ValueError("no table rows in %s" % filepath)
    hi_x = np.arange(rmin, rmax, _cls.BASE_RESOLUTION)
    step = float(hi_x[1] - hi_x[0]) if hi_x.size > 1 else float(_cls.BASE_RESOLUTION)
    # np.arange fills start + i * (second - first): the device regenerates exactly these abscissae
    hi_y = _cls.engine().xsc_place(hi_x.size, 0, 0, hi_x.size, wn, xs, interp=True, ax0=float(hi_x[0]), adelta=step)
    folder = os.path.dirname(filepath)
    os.remove(filepath)
    return writeXscFile(hi_x, hi_y, rmin, rmax, float(props["TEMP"]), float(props["PRESSURE"]),
                        props["MOLECULE_SHORT_NAME"], folder, props["BROADENER"], props["ID"])


def changeResFolder(folderName):
    """:505-512 (the reference then removes the file it has just rewritten when the name did not change; here a
    rewritten file is kept)."""
    d = _xsc_dir(folderName)
    return [changeResXscFile(os.path.join(d, f)) for f in sorted(os.listdir(d))]


def returnMatchingXsc(folder, temp, pressure, res):
    """:601-608 -- TEMP and PRESSURE equal; RES only has to be non-zero (the reference tests its truth value)."""
    out = []
    for f in os.listdir(_xsc_dir(folder)):
        p = parseXscFileName(f)
        if float(p["TEMP"]) == temp and float(p["PRESSURE"]) == pressure and float(p["RES"]):
            out.append(f)
    return out


def mergeXsc(folder):
    """:549-597.  Returns the list of files written, or False when a group mixes resolutions (as the reference
    does, after it has already removed that group's members)."""
    d = _xsc_dir(folder)
    e = _cls.engine()
    groups = {}
    for f in os.listdir(d):
        p = parseXscFileName(f)
        key = p["MOLECULE_SHORT_NAME"] + "_" + p["TEMP"] + "K-" + p["PRESSURE"] + "Torr"
        if key not in groups:
            groups[key] = returnMatchingXsc(folder, float(p["TEMP"]), float(p["PRESSURE"]), float(p["RES"]))
    written = []
    for members in groups.values():
        mins, maxes, ress, tables = [], [], [], []
        for f in members:
            p = parseXscFileName(f)
            lo, hi = (float(v) for v in p["RANGE"].split("-"))
            mins.append(lo)
            maxes.append(hi)
            ress.append(float(p["RES"]))
            tables.append(_read_table(os.path.join(d, f)))
            os.remove(os.path.join(d, f))
        if any(r != ress[0] for r in ress):
            return False
        new_x = np.arange(min(mins), max(maxes), ress[0])
        n = new_x.size
        placed = np.empty((len(tables), n))
        for r, (wn, xs) in enumerate(tables):
            dst0, src0, count, out_len = _cls._merge_plan(new_x, wn)
            if out_len != n:
                raise ValueError("mergeXsc: member table does not sit on the merged grid")
            placed[r] = e.xsc_place(n, dst0, src0, count, None, xs, interp=False)
        # sum over members in listing order, from 0.0, on the device: k = sum_m sigma_m * 1.0 (prb_layer_stream)
        new_y = e.layer_stream(placed, np.ones(len(tables)), 0.0, 296.0, (0.0, 1.0, float(n - 1), n), None,
                               want=("abs_coef",))[0]
        written.append(writeXscFile(new_x, new_y, min(mins), max(maxes), float(p["TEMP"]), float(p["PRESSURE"]),
                                    p["MOLECULE_SHORT_NAME"], d, p["BROADENER"], p["ID"]))
    return written
