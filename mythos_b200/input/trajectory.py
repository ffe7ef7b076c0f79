"""oxDNA trajectory ingest on the device (interface of ``mythos/input/trajectory.py:192-320``).

``from_file(path, strand_lengths, is_5p_3p=True)`` reads the file's bytes once into pinned host memory, copies them to the
GPU and parses them there (``csrc/trajectory.cu``): line index -> one thread per line -> centres ``(F,N,3)`` and
quaternions ``(F,N,4)`` in the internal 3'->5' order, ready for ``energy_fn.map`` / ``DiffTReObjective``.  The reference
parses the same text line by line in Python and converts every state with numpy (``NucleotideState.quaternions``).
Numbers are converted with correct rounding, so centres are bit-identical to ``np.fromstring``'s; a number the device
conversion cannot decide (more than 19 significant digits, decimal exponent beyond +-64) raises -- there is no host parser
behind this one.
"""

from __future__ import annotations

import ctypes as C
import dataclasses as dc
from functools import lru_cache
from pathlib import Path

import numpy as np
import torch

from mythos_b200 import _lib
from mythos_b200.rigid_body import Quaternion, RigidBody
from mythos_b200.simulators.io import SimulatorTrajectory

ERR_TRAJECTORY_FILE_NOT_FOUND = "Trajectory file not found: {}"
ERR_FIXED_BOX_SIZE = "Only trajecories in a fixed box size are supported"
POW5_MIN, POW5_MAX = -64, 64


@lru_cache(maxsize=1)
def pow5_table() -> np.ndarray:
    """(129, 2) uint64: 128-bit truncated, normalised 5^q for q in [-64, 64] as {high, low} words -- the table of the
    Eisel-Lemire conversion in ``csrc/parse_decimal.cuh`` (negative powers are reciprocals rounded up)."""
    rows = []
    for q in range(POW5_MIN, POW5_MAX + 1):
        if q >= 0:
            v = 5**q
            sh = v.bit_length() - 128
            t = v >> sh if sh > 0 else v << (-sh)
        else:
            d = 5 ** (-q)
            b = d.bit_length() + 127
            t = (1 << b) // d + (1 if (1 << b) % d else 0)
            while t.bit_length() > 128:
                t >>= 1
        rows.append(((t >> 64) & 0xFFFFFFFFFFFFFFFF, t & 0xFFFFFFFFFFFFFFFF))
    return np.array(rows, dtype=np.uint64)


_POW5_DEV: dict = {}


def _pow5_on(device) -> torch.Tensor:
    key = str(device)
    if key not in _POW5_DEV:
        _POW5_DEV[key] = torch.from_numpy(pow5_table().view(np.int64).copy()).to(device)
    return _POW5_DEV[key]


def destination_rows(strand_lengths, is_5p_3p: bool) -> np.ndarray | None:
    """Row r of a state in the file -> nucleotide index in the internal order: per-strand reversal for 5'->3' files
    (``trajectory.py:289-293``), identity (None) otherwise."""
    if not is_5p_3p:
        return None
    dest, start = [], 0
    for n in strand_lengths:
        dest += list(range(start + n - 1, start - 1, -1))
        start += n
    return np.asarray(dest, dtype=np.int32)


@dc.dataclass(frozen=True)
class Trajectory:
    """Parsed trajectory, resident on the device (fields of the reference's ``Trajectory`` that the energy path uses)."""

    n_nucleotides: int
    strand_lengths: list
    times: np.ndarray  # (F)
    energies: np.ndarray  # (F,3)
    box_size: np.ndarray  # (3)
    center: torch.Tensor  # (F,N,3) device
    quat: torch.Tensor  # (F,N,4) device

    @property
    def state_rigid_body(self) -> RigidBody:
        return RigidBody(self.center, Quaternion(self.quat))

    def __len__(self) -> int:
        return int(self.center.shape[0])

    def slice(self, key) -> "Trajectory":
        key = slice(key, key + 1) if isinstance(key, int) else key
        return dc.replace(self, times=self.times[key], energies=self.energies[key], center=self.center[key], quat=self.quat[key])

    def to_simulator_trajectory(self, kT: float | None = None) -> SimulatorTrajectory:
        temp = None if kT is None else torch.full((len(self),), float(kT), dtype=self.center.dtype, device=self.center.device)
        return SimulatorTrajectory(center=self.center, orientation=Quaternion(self.quat), temperature=temp)


def parse_bytes(text: torch.Tensor, strand_lengths, *, is_5p_3p: bool = True, dtype=torch.float64):
    """Device parse of a uint8 CUDA tensor holding the file -> (center, quat, times, box, energies) with the last three
    on the host.  Raises on malformed files."""
    _lib.require_cuda(text, "trajectory bytes")
    dev = text.device
    n = int(sum(strand_lengths))
    n_bytes = int(text.numel())
    if n_bytes == 0:
        raise ValueError("empty trajectory file")
    if text.data_ptr() % 16:
        text = text.clone()
    lib = _lib.lib()
    ws_bytes = int(lib.mythos_b200_traj_workspace_bytes(n_bytes))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    n_lines_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        _lib.check(lib.mythos_b200_traj_index(st, text.data_ptr(), n_bytes, ws.data_ptr(), ws_bytes, None, 0, n_lines_dev.data_ptr()), "traj_index")
        n_lines = int(n_lines_dev.item())  # the one host read the output shapes need
        per = n + 3
        n_frames = n_lines // per
        if n_frames == 0 or n_lines % per:
            # a trailing blank line is tolerated; anything else means the strand lengths do not describe this file
            tail_ok = n_lines % per == 1 and n_frames > 0 and bytes(text[-1:].cpu().numpy()) in (b"\n",) and n_bytes >= 2 and bytes(text[-2:-1].cpu().numpy()) == b"\n"
            if not tail_ok:
                raise ValueError(f"trajectory has {n_lines} lines, not a multiple of {per} (= {n} nucleotides + 3 header lines per state)")
        line_start = torch.empty(n_lines + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.mythos_b200_traj_index(st, text.data_ptr(), n_bytes, ws.data_ptr(), ws_bytes, line_start.data_ptr(), n_lines + 1,
                                              n_lines_dev.data_ptr()), "traj_index")
        dest = destination_rows(strand_lengths, is_5p_3p)
        dest_dev = None if dest is None else torch.from_numpy(dest).to(dev)
        center = torch.empty((n_frames, n, 3), dtype=dtype, device=dev)
        quat = torch.empty((n_frames, n, 4), dtype=dtype, device=dev)
        times = torch.empty(n_frames, dtype=torch.float64, device=dev)
        box = torch.empty((n_frames, 3), dtype=torch.float64, device=dev)
        energies = torch.empty((n_frames, 3), dtype=torch.float64, device=dev)
        status = torch.zeros(4, dtype=torch.int32, device=dev)
        a = _lib.TrajArgs()
        a.text, a.n_bytes, a.line_start, a.n_lines = text.data_ptr(), n_bytes, line_start.data_ptr(), n_lines
        a.n, a.n_frames, a.dest, a.pow5 = n, n_frames, _lib.ptr(dest_dev), _pow5_on(dev).data_ptr()
        a.center, a.quat, a.times, a.box, a.energies, a.status = (center.data_ptr(), quat.data_ptr(), times.data_ptr(), box.data_ptr(),
                                                                   energies.data_ptr(), status.data_ptr())
        fn = getattr(lib, f"mythos_b200_traj_parse_{_lib.suffix(dtype)}")
        _lib.check(fn(st, C.byref(a)), "traj_parse")
        bad_numbers, bad_lines = status[:2].tolist()
    if bad_lines:
        raise ValueError(f"trajectory: {bad_lines} header lines are not where {n} nucleotide lines per state put them "
                         "(wrong strand_lengths or a truncated state)")
    if bad_numbers:
        raise ValueError(f"trajectory: {bad_numbers} lines hold numbers the device parser cannot convert exactly "
                         "(malformed, more than 19 significant digits, or a decimal exponent beyond +-64)")
    return center, quat, times.cpu().numpy(), box.cpu().numpy(), energies.cpu().numpy()


def from_file(path, strand_lengths, *, is_5p_3p: bool = True, n_processes: int = 1, device=None, dtype=torch.float64) -> Trajectory:
    """Parse an oxDNA trajectory file on the GPU (``trajectory.py:192-246``; ``n_processes`` is accepted and ignored: the
    whole file is parsed by one kernel)."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(ERR_TRAJECTORY_FILE_NOT_FOUND.format(path))
    if not torch.cuda.is_available():
        raise _lib.MythosB200Error("trajectory ingest runs on the GPU (no host parser)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    raw = np.fromfile(path, dtype=np.uint8)
    host = torch.from_numpy(raw).pin_memory()
    text = host.to(dev, non_blocking=True)
    center, quat, times, box, energies = parse_bytes(text, strand_lengths, is_5p_3p=is_5p_3p, dtype=dtype)
    if not np.all(box == box[0]):
        raise ValueError(ERR_FIXED_BOX_SIZE)
    return Trajectory(n_nucleotides=int(sum(strand_lengths)), strand_lengths=list(strand_lengths), times=times, energies=energies,
                      box_size=box[0], center=center, quat=quat)
