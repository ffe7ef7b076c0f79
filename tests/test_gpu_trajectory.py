"""Trajectory ingest on the GPU: the device parser against the oracle parser (np.fromstring-style
conversion, the reference's axes -> quaternion formulas) on the head of the reference's own dna1/simple-helix/output.dat,
on mythos-style files (str(float): 16-17 digits), with 5'->3' strand reversal, ragged strands, CRLF and a missing final
newline; malformed files raise; the parsed frames feed the energy function and reproduce oxDNA's energies."""

from pathlib import Path

import numpy as np
import pytest
import torch

from mythos_b200.input import trajectory as jd_traj
from oracle import trajectory_oracle as to
from tests.golden_cases import TOL, load_case
from tests.product_cases import energy_fn_of

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
HEAD = ROOT / "tests" / "golden" / "traj_dna1_simple_helix_head.dat"


def _check(path, strand_lengths, is_5p_3p, dtype=torch.float64):
    traj = jd_traj.from_file(path, strand_lengths, is_5p_3p=is_5p_3p, dtype=dtype)
    ts, bs, es, states = to.read_text(Path(path).read_text(), strand_lengths, is_5p_3p=is_5p_3p)
    c, q = to.rigid_bodies(states)
    assert traj.center.is_cuda and traj.center.shape == c.shape and traj.quat.shape == q.shape
    np.testing.assert_array_equal(traj.times, ts)
    np.testing.assert_array_equal(traj.energies, es)
    np.testing.assert_array_equal(traj.box_size, bs[0])
    if dtype == torch.float64:
        np.testing.assert_array_equal(traj.center.cpu().numpy(), c)  # correctly rounded: bit-identical to np.fromstring
        np.testing.assert_allclose(traj.quat.cpu().numpy(), q, rtol=0, atol=2e-15)  # (device vs glibc atan2 / asin / sincos)
    else:
        np.testing.assert_allclose(traj.center.cpu().numpy(), c, rtol=1e-6)
        np.testing.assert_allclose(traj.quat.cpu().numpy(), q, rtol=0, atol=1e-6)
    return traj


def test_reference_file_head_matches_oracle_and_golden_energies():
    case = load_case("dna1_simple_helix")
    traj = _check(HEAD, case["strand_counts"].tolist(), False)
    np.testing.assert_array_equal(traj.center.cpu().numpy(), case["center"][:5])
    efn = energy_fn_of(case)
    terms = efn.compute_terms_frames(traj.state_rigid_body).cpu().numpy()
    want = case["golden_terms_per_nt"][:5]
    for k in range(terms.shape[1]):
        np.testing.assert_allclose(np.around(terms[:, k] / 16, 6), want[:, k], atol=TOL["dna1"][k], rtol=1e-7)
    _check(HEAD, case["strand_counts"].tolist(), False, dtype=torch.float32)


def _write(path, states, times, box, newline="\n", final_newline=True, fmt=str):
    lines = []
    for t, s in zip(times, states):
        lines += [f"t = {t}", f"b = {box[0]} {box[1]} {box[2]}", "E = -1.5 -2.25 0.75"]
        lines += [" ".join(fmt(x) for x in row) for row in s]
    text = newline.join(lines) + (newline if final_newline else "")
    Path(path).write_text(text, newline="")


@pytest.mark.parametrize("variant", ["repr", "g15", "crlf_nofinal"])
def test_written_files_round_trip(tmp_path, variant):
    rng = np.random.default_rng(5)
    strands = [7, 3, 12]
    n = sum(strands)
    F = 300 if variant == "repr" else 9  # 300 states x 25 lines x ~290 B = 2.2 MB: many 64 KiB chunks of the line index
    states = rng.normal(0, 3, (F, n, 15))
    a1 = rng.normal(size=(F, n, 3))
    a1 /= np.linalg.norm(a1, axis=-1, keepdims=True)
    a3 = np.cross(a1, rng.normal(size=(F, n, 3)))
    a3 /= np.linalg.norm(a3, axis=-1, keepdims=True)
    states[:, :, 3:6], states[:, :, 6:9] = a1, a3
    states[0, 0, 0], states[0, 1, 1], states[0, 2, 2] = 1.25e-7, -3.0e-11, 123456.789012345  # exponent notation, small numbers
    p = tmp_path / "traj.dat"
    if variant == "repr":
        _write(p, states, np.arange(F) * 100, (40.0, 40.0, 40.0))
    elif variant == "g15":
        _write(p, states, np.arange(F) * 100, (40, 40, 40), fmt=lambda x: "%.15g" % x)
    else:
        _write(p, states, np.arange(F) * 100, (40.0, 40.0, 40.0), newline="\r\n", final_newline=False)
    for is53 in (True, False):
        _check(p, strands, is53)


def test_malformed_files_raise(tmp_path):
    rng = np.random.default_rng(1)
    states = rng.normal(size=(2, 4, 15))
    p = tmp_path / "t.dat"
    _write(p, states, [0, 1], (9.0, 9.0, 9.0))
    with pytest.raises(ValueError):  # wrong strand lengths
        jd_traj.from_file(p, [5])
    with pytest.raises(ValueError):
        jd_traj.from_file(p, [2])  # 12 lines = 2 x (3 + 3)?  header lines land on nucleotide rows
    rows = p.read_text().split("\n")
    rows[4] = "abc " + rows[4].split(" ", 1)[1]  # a nucleotide line that does not start with a number
    (tmp_path / "bad.dat").write_text("\n".join(rows))
    with pytest.raises(ValueError):
        jd_traj.from_file(tmp_path / "bad.dat", [4])
    many = p.read_text().split("\n")
    many[3] = "1.234567890123456789012345 " + many[3].split(" ", 1)[1]  # more than 19 significant digits
    (tmp_path / "many.dat").write_text("\n".join(many))
    with pytest.raises(ValueError):
        jd_traj.from_file(tmp_path / "many.dat", [4])
    with pytest.raises(FileNotFoundError):
        jd_traj.from_file(tmp_path / "absent.dat", [4])
    _write(p, states, [0, 1], (9.0, 9.0, 9.0))
    txt = p.read_text().replace("b = 9.0 9.0 9.0", "b = 8.0 9.0 9.0", 1)
    (tmp_path / "box.dat").write_text(txt)
    with pytest.raises(ValueError, match=jd_traj.ERR_FIXED_BOX_SIZE):
        jd_traj.from_file(tmp_path / "box.dat", [4])
