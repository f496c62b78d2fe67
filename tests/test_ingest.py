"""HITRAN-online CSV ingestion.  CPU: the number parser compiled into the library
(host run of the device code) against Python's float(); the oracle's reader against the golden kept-line lists
the REAL reference produced.  GPU: the device parser (k5_*) against the oracle's reader, bit for bit."""
import ctypes as C
import os
import random
import struct

import numpy as np
import pytest

from oracle import physics as ph
from oracle import ref_harness as rh
from pyrad_b200 import _lib
from pyrad_b200 import hitran_io
from pyrad_b200 import synth
from tests import golden_util as G


def _parse(lib, s):
    b = s.encode()
    v = C.c_double()
    rc = lib.prb_debug_parse_double(b, len(b), C.byref(v))
    return rc, v.value


def test_number_parser_matches_python_float():
    lib = _lib.load()
    rng = random.Random(7)
    cases = ["0", "-0.0", "1.", "  .5 ", "+3.25e-2", "1e-400", "1e400", "4.9e-324", "2.4703282292062327e-324",
             "2.4703282292062328e-324", "9007199254740993", "1.7976931348623157e308", "1.7976931348623159e308",
             "123456789012345678", "0.000001234E-25", "1E23", "8.5E22", "9007199254740992.5E0", "667.661000",
             "3.000E-19", "1.234E-30\r", "0.0712", "-.004500", "5000.000001"]
    for _ in range(20000):
        k = rng.random()
        if k < 0.35:
            cases.append("%dE%d" % (rng.randint(0, 10 ** rng.randint(1, 19) - 1), rng.randint(-340, 310)))
        elif k < 0.7:
            cases.append("%.*E" % (rng.randint(0, 17), rng.uniform(-1, 1) * 10.0 ** rng.randint(-320, 308)))
        elif k < 0.85:
            cases.append(repr(rng.uniform(0, 5000)))
        else:
            cases.append("%.17g" % struct.unpack("d", struct.pack("Q", rng.getrandbits(62)))[0])
    for s in cases:
        rc, v = _parse(lib, s)
        ref = float(s)
        assert rc == 0, s
        assert v == ref and np.signbit(v) == np.signbit(ref), (s, v, ref)
    for s in ["", "abc", "1e", "--1", "1.2.3", "#", "nan", "inf", "1_0", "0x10", "1 2"]:
        assert _parse(lib, s)[0] == -7, s                      # PRB_ERR_PARSE: fail loudly, as float() raises


def test_number_parser_long_literals():
    """More than 19 significant digits (float() parses them; a HITRAN field padded with zeros is one): the first 19 digits
    and "were the rest zero" decide the result whenever both ends of that interval round to the same double -- always
    for trailing zeros.  Whatever is accepted equals float(); the rare undecided literal is refused, never guessed."""
    lib = _lib.load()
    rng = random.Random(11)
    for s in ["12345678901234567890", "667.661000000000000000000000", "1.00000000000000000000000000E-5",
              "0.000000000000000000000000000001234567890123456789012345", "123456789012345678901234567890E-20",
              "9007199254740993.00000000000000", "5000.0000010000000000000000000"]:
        rc, v = _parse(lib, s)
        assert rc == 0 and v == float(s), (s, v)
    accepted = 0
    for _ in range(20000):
        digits = "".join(rng.choice("0123456789") for _ in range(rng.randint(20, 40)))
        cut = rng.randint(0, len(digits))
        s = (digits[:cut] + "." + digits[cut:] if rng.random() < 0.7 else digits) + ("E%d" % rng.randint(-330, 290) if rng.random() < 0.5 else "")
        rc, v = _parse(lib, s)
        if rc == 0:
            accepted += 1
            ref = float(s)
            assert v == ref or (np.isinf(ref) and np.isinf(v)), (s, v, ref)
    assert accepted > 19000                                     # undecided (refused) literals are rare


def test_number_parser_property_any_decimal_literal():
    """Property test (hypothesis): every decimal literal with at most 19 significant digits converts to exactly the
    double Python's float() returns -- sign of zero, subnormals, overflow to inf included."""
    from hypothesis import given, settings, strategies as st
    lib = _lib.load()
    literal = st.from_regex(r"[ ]{0,2}[+-]?(([0-9]{1,12}(\.[0-9]{0,7})?)|(\.[0-9]{1,12}))([eE][+-]?[0-9]{1,3})?[ ]{0,2}", fullmatch=True)

    @settings(max_examples=3000, deadline=None)
    @given(literal)
    def check(s):
        rc, v = _parse(lib, s)
        ref = float(s)
        assert rc == 0, s
        assert (v == ref and np.signbit(v) == np.signbit(ref)) or (np.isinf(ref) and np.isinf(v) and (v > 0) == (ref > 0)), (s, v, ref)

    check()


@pytest.mark.parametrize("name", G.CELL_CASES)
def test_oracle_reader_keeps_what_the_reference_kept(name, tmp_path):
    g = G.load(name)
    cutoff = ph.layer_cutoff(float(g["P"]))
    lo, hi = ph.effective_range(float(g["range_min"]), float(g["range_max"]), cutoff)
    for i in range(len(g["conc"])):
        ln = G.lines_of(g, i)
        rh.write_line_segments(str(tmp_path), 900 + i, 2, 1, ln, int(ln["nu"].min() / 100) * 100, ln["nu"].max() + 101)
        text = hitran_io.gather_text(900 + i, lo, hi, str(tmp_path)).decode()
        rows = [r for r in text.splitlines() if not r.startswith("#")]
        got = ph.read_hitran_online_rows(rows, lo, hi)
        np.testing.assert_array_equal(got["nu"], g["kept_nu_%d" % i])


def _csv(lines, mol=2, iso=1, fmt=repr):
    rows = ["%d,%d,%s" % (mol, iso, ",".join(fmt(float(lines[k][j])) for k in
                                              ("nu", "sw", "a", "elower", "gamma_air", "gamma_self", "delta_air", "n_air")))
            for j in range(len(lines["nu"]))]
    return rows


@pytest.mark.gpu
def test_device_ingest_matches_oracle_reader(engine):
    ln = synth.make_lines(5000, 480.0, 830.0, 123)
    rows = _csv(ln)
    # HITRAN's own fixed formats, duplicates (last wins), comment rows, CRLF, a row without trailing newline
    rows[10] = rows[10].replace(repr(float(ln["sw"][10])), "%.3E" % ln["sw"][10])
    rows.insert(200, rows[199].rsplit(",", 1)[0] + ",0.123")           # same nu, different n_air: replaces row 199
    rows.insert(1000, "# a comment row in the middle")
    rows[2000] = rows[2000] + "\r"
    text = ("# header\n" + "\n".join(rows)).encode()                   # no trailing newline
    lo, hi = 500.0, 800.0
    ref = ph.read_hitran_online_rows([r for r in text.decode().split("\n") if not r.startswith("#")], lo, hi)
    n = engine.ingest_csv(text, lo, hi)
    got = engine.download_lines()
    assert n == len(ref["nu"]) and 0 < n < 5000
    for k in hitran_io.LINE_COLUMNS:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)       # bit exact: same float() values, same order
    # the list is live on the device: the line sum runs on it without another upload
    engine.set_grid(500.0, 0.01, 30000)
    engine.layer_prepass(296, 1013.25, [4e-4], [43.98983], [286.09], [286.09], 500)
    sig = engine.line_sum()
    sigma_ref = ph.cross_section(ref, 296, 1013.25, 4e-4, 43.98983, 286.09, 286.09, 500.0, 800.0, 0.01, ph.layer_cutoff(1013.25))
    err = np.abs(sig - sigma_ref) / np.maximum(np.abs(sigma_ref), 1e-40 * np.abs(sigma_ref).max())
    assert err.max() <= 1e-5


@pytest.mark.gpu
def test_device_ingest_edge_cases(engine):
    assert engine.ingest_csv(b"", 0.0, 10.0) == 0
    assert engine.ingest_csv(b"# only a comment\n", 0.0, 10.0) == 0
    one = b"2,1,5.5,1e-20,1.0,100.0,0.07,0.09,-0.002,0.7"
    assert engine.ingest_csv(one, 0.0, 10.0) == 1
    assert engine.ingest_csv(one, 5.5, 10.0) == 0                      # strict bounds
    assert engine.ingest_csv(one, 0.0, 5.5) == 0
    # out-of-range rows are not parsed beyond their wavenumber (the reference never touches their other cells)
    assert engine.ingest_csv(b"2,1,50.0,#,#,#,#,#,#,#\n" + one + b"\n", 0.0, 10.0) == 1
    with pytest.raises(_lib.EngineError) as ei:
        engine.ingest_csv(b"2,1,5.5,#,1.0,100.0,0.07,0.09,-0.002,0.7\n", 0.0, 10.0)
    assert ei.value.code == -7 and "row 0" in str(ei.value)
    with pytest.raises(_lib.EngineError):
        engine.ingest_csv(b"2,1,5.5,1e-20\n", 0.0, 10.0)               # too few cells
    # descending wavenumbers: sorted on the device (the reference's reader takes any order)
    assert engine.ingest_csv(one + b"\n" + one.replace(b"5.5", b"4.5") + b"\n", 0.0, 10.0) == 2
    np.testing.assert_array_equal(engine.download_lines()["nu"], [4.5, 5.5])


@pytest.mark.gpu
def test_device_ingest_takes_rows_in_any_order(engine):
    """readHitranOnlineFile keys a dict by wavenumber (pyradUtilities.py:421-448): the rows may come in any order and a
    later row replaces an earlier one of the same wavenumber WHEREVER it stands.  The device path sorts the kept rows
    (stable radix sort) and keeps the last row of every wavenumber: the same set of lines, ascending."""
    ln = synth.make_lines(4000, 480.0, 830.0, 321)
    rows = _csv(ln)
    rng = np.random.default_rng(3)
    dup = [rows[i].rsplit(",", 1)[0] + ",0.%03d" % i for i in (7, 1500, 3999)]      # same nu, new n_air
    order = rng.permutation(len(rows))
    shuffled = [rows[i] for i in order] + dup                                        # the replacements come last
    text = ("\n".join(shuffled) + "\n").encode()
    lo, hi = 500.0, 800.0
    ref = ph.read_hitran_online_rows(shuffled, lo, hi)                               # dict semantics, insertion order
    asc = np.argsort(ref["nu"], kind="stable")
    n = engine.ingest_csv(text, lo, hi)
    got = engine.download_lines()
    assert n == len(ref["nu"]) and np.all(np.diff(got["nu"]) > 0)
    for k in hitran_io.LINE_COLUMNS:
        np.testing.assert_array_equal(got[k], ref[k][asc], err_msg=k)
    inside = [i for i in (7, 1500, 3999) if lo < ln["nu"][i] < hi]
    assert inside and all(got["n_air"][np.searchsorted(got["nu"], ln["nu"][i])] == float("0.%03d" % i) for i in inside)


@pytest.mark.gpu
def test_device_xsc_table_parse_matches_oracle_reader(engine):
    fx, fy = synth.make_xsc_table(800.0, 860.0, 0.02, 9)
    rows = ["%r     %r" % (float(a), float(b)) for a, b in zip(fx, fy)]
    rows[5] = "  " + rows[5] + "   "                     # stripped
    rows[9] = rows[9].replace("     ", " ")              # any run of spaces separates
    rows[11] = "%.6f\t%.6e" % (fx[11], fy[11])           # a tab is not a separator: float() fails, row skipped
    rows[20] = rows[20] + " 7"                           # three tokens: skipped
    rows[30] = "garbage here"                            # not numbers: skipped
    rows.insert(40, "")                                  # blank row: skipped
    rows.insert(50, "# comment")                         # skipped
    rows[60] = "%.4E %.3E\r" % (fx[58], fy[58])          # exponent formats, CRLF
    text = ("# header line\n" + "\n".join(rows)).encode()
    ref_rows = text.decode().split("\n")[1:]
    wn_ref, xs_ref = ph.read_xsc_rows(ref_rows)
    wn, xs = engine.parse_xsc_text(text)
    assert 0 < len(wn_ref) < len(rows)
    np.testing.assert_array_equal(wn, wn_ref)
    np.testing.assert_array_equal(xs, xs_ref)
    assert engine.parse_xsc_text(b"")[0].size == 0 and engine.parse_xsc_text(b"# nothing\n\n")[0].size == 0
