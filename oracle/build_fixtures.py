"""Build tests/golden/*.npz from the reference's own oxDNA-standalone fixtures.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference exists):

    python oracle/build_fixtures.py

The GPU box has no /root/reference, so the golden inputs (frames converted to centre + quaternion exactly as
the reference's parser does) and golden outputs (oxDNA's per-term energies) are committed as small .npz files.

What is restated here (relative to /root/reference):
  topology parsing    mythos/input/topology.py:193-327 (classic and new format, 3'->5' internal order,
                      new-format strand reversal at :291)
  trajectory parsing  mythos/input/trajectory.py:273-320 (15 floats per nucleotide, is_5p_3p reversal)
  axes -> quaternion  mythos/input/trajectory.py:163-175, mythos/utils/math.py:9-65
  golden columns      mythos/energy/dna1/tests/test_integration.py:21-40 (split_energy.dat, skiprows=1)
  seq-dep weights     mythos/input/sequence_dependence.py:12-51
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/data/test-data")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
NT_IDX = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}


def read_topology(path: Path):
    lines = path.read_text().strip().splitlines()
    head = lines[0].split()
    seq, counts, is_end, nt_type, circ = [], [], [], [], []
    if len(head) == 2:  # classic: strand_id base 3' 5'
        rows = [ln.split() for ln in lines[1:]]
        ids = np.array([int(r[0]) for r in rows])
        for sid in range(1, int(head[1]) + 1):
            sel = [r for r in rows if int(r[0]) == sid]
            bases = [r[1] for r in sel]
            circular = int(sel[-1][3]) != -1
            n = len(bases)
            counts.append(n)
            circ.append(circular)
            seq += bases
            e = [0] * n
            if not circular:
                e[0] = e[-1] = 1
            is_end += e
            nt_type += [1 if "T" in bases else 2 if "U" in bases else 0] * n
        new_format = False
        del ids
    else:  # new: "ACGT id=1 type=DNA circular=false", written 5'->3'
        for ln in lines[1:]:
            nts = ln.split()[0]
            n = len(nts)
            counts.append(n)
            circular = "circular=true" in ln
            circ.append(circular)
            seq += list(nts[::-1])
            e = [0] * n
            if not circular:
                e[0] = e[-1] = 1
            is_end += e
            nt_type += [1 if "type=DNA" in ln else 2 if "type=RNA" in ln else 0] * n
        new_format = True
    return {
        "seq": np.array([NT_IDX[s] for s in seq], dtype=np.int32),
        "strand_counts": np.array(counts, dtype=np.int32),
        "circular": np.array(circ, dtype=np.int32),
        "is_end": np.array(is_end, dtype=np.int32),
        "nt_type": np.array(nt_type, dtype=np.int32),
        "new_format": new_format,
    }


def read_trajectory(path: Path, strand_counts, is_5p_3p: bool):
    n = int(sum(strand_counts))
    bounds = np.concatenate([[0], np.cumsum(strand_counts)])
    frames, boxes, cur = [], [], []
    for ln in path.read_text().splitlines():
        if not ln.strip():
            continue
        c = ln[0]
        if c == "t" or c == "E":
            continue
        if c == "b":
            boxes.append(np.array(ln.split("=")[1].split(), dtype=np.float64))
            continue
        cur.append(np.array(ln.split(), dtype=np.float64))
        if len(cur) == n:
            st = np.stack(cur)
            if is_5p_3p:
                st = np.concatenate([st[s:e][::-1] for s, e in zip(bounds[:-1], bounds[1:])])
            frames.append(st)
            cur = []
    return np.stack(frames), boxes[0]


def axes_to_quaternion(a1, a3):
    """x=a1, z=a3, y=a3 x a1 -> Tait-Bryan ZYX angles -> quaternion (w,x,y,z)."""
    y = np.cross(a3, a1)
    psi = np.arctan2(a1[..., 1], a1[..., 0])
    theta = np.arcsin(-np.clip(a1[..., 2], -1, 1))
    phi = np.arctan2(y[..., 2], a3[..., 2])
    sp, cp = np.sin(0.5 * psi), np.cos(0.5 * psi)
    st, ct = np.sin(0.5 * theta), np.cos(0.5 * theta)
    sf, cf = np.sin(0.5 * phi), np.cos(0.5 * phi)
    q0 = sp * st * sf + cp * ct * cf
    q1 = -sp * st * cf + sf * cp * ct
    q2 = sp * ct * sf + cp * st * cf
    q3 = sp * ct * cf - cp * st * sf
    return np.stack([q0, q1, q2, q3], -1)


def read_seq_dep(path: Path):
    kv = {}
    for ln in path.read_text().splitlines():
        s = ln.strip().replace(" ", "")
        if s:
            k, v = s.split("=")
            kv[k] = float(v.replace("f", ""))
    stack = np.zeros((4, 4))
    for i, a in enumerate("ACGT"):
        for j, b in enumerate("ACGT"):
            stack[i, j] = kv[f"STCK_{a}_{b}"]
    hb = np.zeros((4, 4))
    at = kv.get("HYDR_A_T", kv.get("HYDR_T_A"))
    gc = kv.get("HYDR_G_C", kv.get("HYDR_C_G"))
    hb[0, 3] = hb[3, 0] = at
    hb[2, 1] = hb[1, 2] = gc
    return {"ss_stack_weights": stack, "ss_hb_weights": hb, "eps_stack_kt_coeff": kv["STCK_FACT_EPS"]}


# (name, model, dir, is_5p_3p as the reference test passes it, topology file, extras)
CASES = [
    ("dna1_simple_helix", "dna1", "dna1/simple-helix", False, "generated.top", {}),
    ("dna1_simple_coax", "dna1", "dna1/simple-coax", False, "generated.top", {}),
    ("dna1_seq_dep", "dna1", "dna1/simple-helix-seq-dep", False, "generated.top", {"seq_dep": "seq_dep.dat"}),
    ("dna2_simple_helix", "dna2", "dna2/simple-helix", False, "generated.top", {"hce": 0, "salt": 0.5}),
    ("dna2_simple_coax", "dna2", "dna2/simple-coax", False, "generated.top", {"hce": 0, "salt": 0.5}),
    ("dna2_half_charged", "dna2", "dna2/simple-helix-half-charged-ends", False, "generated.top", {"hce": 1, "salt": 0.5}),
    ("rna2_helix_12bp", "rna2", "rna2/simple-helix-12bp", False, "generated.top", {"hce": 0, "salt": 1.0}),
    ("rna2_simple_coax", "rna2", "rna2/simple-coax", False, "generated.top", {"hce": 0, "salt": 1.0}),
    ("na1_helix_dna_dna", "na1", "na1/simple-helix-dna-dna", True, "generated.top", {"hce": 0, "salt": 0.5}),
    ("na1_helix_rna_rna", "na1", "na1/simple-helix-rna-rna", True, "generated.top", {"hce": 0, "salt": 0.5}),
    ("na1_helix_dna_rna", "na1", "na1/simple-helix-dna-rna", True, "generated.top", {"hce": 0, "salt": 0.5}),
    ("na1_helix_rna_dna", "na1", "na1/simple-helix-rna-dna", True, "generated.top", {"hce": 0, "salt": 0.5}),
    ("na1_coax_dna", "na1", "na1/simple-coax-dna-dna-dna", True, "generated.top", {"hce": 0, "salt": 0.5}),
    ("na1_coax_rna", "na1", "na1/simple-coax-rna-rna-rna", True, "generated.top", {"hce": 0, "salt": 0.5}),
]


def main() -> int:
    if not REF.exists():
        print("reference data not present; fixtures can only be rebuilt in the build container", file=sys.stderr)
        return 1
    OUT.mkdir(parents=True, exist_ok=True)
    for name, model, d, is53, topf, extra in CASES:
        base = REF / d
        top = read_topology(base / topf)
        traj, box = read_trajectory(base / "output.dat", top["strand_counts"], is53)
        split = np.loadtxt(base / "split_energy.dat", skiprows=1)
        total = np.loadtxt(base / "energy.dat", skiprows=1)
        nf = min(len(traj), len(split))
        traj, split, total = traj[:nf], split[:nf], total[:nf]
        center = traj[:, :, 0:3]
        quat = axes_to_quaternion(traj[:, :, 3:6], traj[:, :, 6:9])
        terms = np.zeros((nf, 8))
        terms[:, : split.shape[1] - 1] = split[:, 1:]
        payload = {
            "model": model,
            "center": center,
            "quat": quat,
            "seq": top["seq"],
            "strand_counts": top["strand_counts"],
            "circular": top["circular"],
            "is_end": top["is_end"],
            "nt_type": top["nt_type"],
            "file_box": box,
            "golden_terms_per_nt": terms,  # oxDNA split_energy.dat columns 1.., 6 dp, per nucleotide
            "golden_potential_per_nt": total[:, 1],  # energy.dat column 1 (potential energy per nucleotide)
            "t_kelvin": 296.15,
            "half_charged_ends": int(extra.get("hce", 0)),
            "salt_conc": float(extra.get("salt", 0.5)),
            "source": d,
        }
        if "seq_dep" in extra:
            payload.update(read_seq_dep(base / extra["seq_dep"]))
        np.savez_compressed(OUT / f"{name}.npz", **payload)
        print(f"{name}: frames={nf} N={center.shape[1]} box={box}")

    return 0


if __name__ == "__main__":
    raise SystemExit(main())
