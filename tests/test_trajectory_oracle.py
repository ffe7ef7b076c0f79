"""Trajectory ingest, CPU side: the oracle parser against the committed golden frames (which came from the reference's
own output.dat by the reference's formulas), the decimal conversion of the device parser (compiled for the host by
tests/host_check/parse_check.cpp) against Python's correctly rounded float() on a million tokens, and the per-strand
reversal map."""

import ctypes as C
import random
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from mythos_b200.input import trajectory as jd_traj
from oracle import trajectory_oracle as to
from tests.golden_cases import load_case

ROOT = Path(__file__).resolve().parent.parent
HEAD = ROOT / "tests" / "golden" / "traj_dna1_simple_helix_head.dat"


def test_oracle_parser_reproduces_the_golden_frames():
    case = load_case("dna1_simple_helix")
    ts, bs, es, states = to.read_text(HEAD.read_text(), case["strand_counts"].tolist(), is_5p_3p=False)
    assert states.shape == (5, 16, 15) and ts.tolist() == [100.0, 200.0, 300.0, 400.0, 500.0][: len(ts)] or len(ts) == 5
    np.testing.assert_array_equal(bs, np.full((5, 3), 20.0))
    c, q = to.rigid_bodies(states)
    np.testing.assert_array_equal(c, case["center"][:5])
    np.testing.assert_allclose(q, case["quat"][:5], rtol=0, atol=1e-15)


def test_destination_rows():
    assert jd_traj.destination_rows([3, 2], False) is None
    assert jd_traj.destination_rows([3, 2], True).tolist() == [2, 1, 0, 4, 3]


@pytest.fixture(scope="module")
def parse_lib():
    out = ROOT / "tests" / "_build" / "libparse_check.so"
    out.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(out), str(ROOT / "tests" / "host_check" / "parse_check.cpp")], check=True)
    return C.CDLL(str(out))


def _run(lib, tokens):
    table = jd_traj.pow5_table()
    buf = b"\0".join(t.encode() for t in tokens) + b"\0"
    out, ok = np.zeros(len(tokens)), np.zeros(len(tokens), dtype=np.int32)
    lib.parse_check_tokens(buf, C.c_long(len(buf)), len(tokens), C.c_void_p(table.ctypes.data), C.c_void_p(out.ctypes.data), C.c_void_p(ok.ctypes.data))
    return out, ok


def test_decimal_conversion_is_correctly_rounded(parse_lib):
    rng = random.Random(7)
    toks = ["0", "-0", "0.0", "1", "-1.5", "4.35", "0.1", "0.30000000000000004", "9007199254740993", "9007199254740992.5",
            "1e22", "1e23", "8.5e-23", "1.7976931348623157e64", "2.2250738585072014e-40", "123456789012345678", "9999999999999999999",
            "0.000001", "5e-324".replace("324", "40"), "17.000000000000004", "2.5000000000000004e+00"]
    for _ in range(250000):  # shortest round-trip representations of random doubles (what str(float) writes)
        x = struct.unpack("d", struct.pack("Q", rng.getrandbits(64)))[0]
        if x != x or abs(x) > 1e15 or abs(x) < 1e-12:
            x = rng.uniform(-30, 30)
        toks.append(repr(x))
    for _ in range(250000):  # random digit strings up to 19 digits, point anywhere, optional exponent
        s = str(rng.randint(1, 10 ** rng.randint(1, 19) - 1))
        pos = rng.randint(0, len(s))
        toks.append(("-" if rng.random() < 0.5 else "") + s[:pos] + "." + s[pos:] + (f"e{rng.randint(-25, 10):+d}" if rng.random() < 0.5 else ""))
    for _ in range(250000):  # what oxDNA writes: 15 significant digits
        toks.append("%.15g" % rng.uniform(-50, 50))
        toks.append("%.15g" % (rng.uniform(-1, 1) * 10 ** rng.randint(-9, 3)))
    out, ok = _run(parse_lib, toks)
    assert ok.all(), [t for t, k in zip(toks, ok) if not k][:5]
    want = np.array([float(t) for t in toks])
    same = (out == want) & (np.signbit(out) == np.signbit(want))
    assert same.all(), [(t, o, w) for t, o, w, s in zip(toks, out, want, same) if not s][:5]


def test_decimal_conversion_refuses_what_it_cannot_decide(parse_lib):
    out, ok = _run(parse_lib, ["12345678901234567890.5", "1e200", "2.2250738585072014e-64", "abc", "", "1.5e", "--1", "1.2.3"])
    assert not ok.any()
