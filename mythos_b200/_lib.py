"""ctypes binding of the C-ABI in ``include/mythos_b200.h`` (the same symbols the XLA-FFI shim adapts).

Host code stays in Python (as in the reference); this module is the only place that touches the shared
object.  There is no fallback: if ``libmythos_b200.so`` is missing or a call is made without a CUDA device
the error is raised to the caller.
"""

from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache
from pathlib import Path

import torch

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("MYTHOS_B200_LIB", PKG / "libmythos_b200.so"))  # env override: kernel-variant experiments

N_TERMS = 8
TERM_NAMES = (
    "fene",
    "bonded_excluded_volume",
    "stacking",
    "unbonded_excluded_volume",
    "hydrogen_bonding",
    "cross_stacking",
    "coaxial_stacking",
    "debye",
)
BONDED_TERMS, UNBONDED_TERMS, ALL_TERMS = 0x07, 0xF8, 0xFF
FLAG_ACCUMULATE = 0x1
FLAG_GENERIC_KERNEL = 0x2
FLAG_LIST_KERNEL = 0x4
NL_ROWS = 0x1
NL_TAG_SUPPORTS = 0x2
NL_WARP_SLOTS = 0x4
NL_REUSE_EXCLUSIONS = 0x8
NL_PACKED_SLOTS = 0x10
FLAG_TAGGED_PAIRS = 0x8
MAX_BANKS = 3
STATUS = {0: "MB_OK", 1: "MB_EINVAL_SHAPE", 2: "MB_EINVAL_MODEL", 3: "MB_ECAPACITY", 4: "MB_ECUDA"}


class MythosB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


class FlavourGeom(C.Structure):
    _fields_ = [
        ("back", C.c_double * 3),
        ("back_stack", C.c_double),
        ("stack", C.c_double),
        ("base", C.c_double),
        ("stack3", C.c_double * 2),
        ("stack5", C.c_double * 2),
        ("p3", C.c_double * 3),
        ("p5", C.c_double * 3),
        ("use_back_stack", C.c_int32),
        ("_pad", C.c_int32),
    ]


class BankForms(C.Structure):
    _fields_ = [("stack_form", C.c_int32), ("cross_form", C.c_int32), ("coax_form", C.c_int32), ("has_debye", C.c_int32)]


class Model(C.Structure):
    _fields_ = [
        ("n_banks", C.c_int32),
        ("half_charged_ends", C.c_int32),
        ("geom", FlavourGeom * 2),
        ("forms", BankForms * MAX_BANKS),
        ("box", C.c_double * 3),
    ]


class EnergyArgs(C.Structure):
    _fields_ = [
        ("model", C.POINTER(Model)),
        ("n", C.c_int32),
        ("n_frames", C.c_int32),
        ("center", C.c_void_p),
        ("quat", C.c_void_p),
        ("seq", C.c_void_p),
        ("nt_type", C.c_void_p),
        ("nt_type_stack", C.c_void_p),
        ("is_end", C.c_void_p),
        ("bonded", C.c_void_p),
        ("n_bonded", C.c_int32),
        ("pairs", C.c_void_p),
        ("pair_capacity", C.c_int64),
        ("pair_frame_stride", C.c_int64),
        ("params", C.c_void_p),
        ("cot", C.c_void_p),
        ("term_mask", C.c_uint32),
        ("flags", C.c_uint32),
        ("terms", C.c_void_p),
        ("d_center", C.c_void_p),
        ("d_quat", C.c_void_p),
        ("d_params", C.c_void_p),
        ("d_params_frame_stride", C.c_int64),
        ("pair_count", C.c_void_p),
        ("all_pairs_cutoff", C.c_double),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("pair_split", C.c_void_p),
        ("observables", C.c_void_p),
        ("observables_out", C.c_void_p),
        ("pseq", C.c_void_p),
    ]


class Pseq(C.Structure):
    _fields_ = [("pmarg", C.c_void_p), ("bp_of", C.c_void_p), ("within", C.c_void_p), ("same_w_stack", C.c_void_p),
                ("same_w_hb", C.c_void_p), ("d_pmarg", C.c_void_p), ("d_same_w_stack", C.c_void_p), ("d_same_w_hb", C.c_void_p),
                ("terms", C.c_uint32), ("_pad", C.c_uint32)]


class TrajArgs(C.Structure):
    _fields_ = [("text", C.c_void_p), ("n_bytes", C.c_int64), ("line_start", C.c_void_p), ("n_lines", C.c_int64), ("n", C.c_int32),
                ("n_frames", C.c_int32), ("dest", C.c_void_p), ("pow5", C.c_void_p), ("center", C.c_void_p), ("quat", C.c_void_p),
                ("times", C.c_void_p), ("box", C.c_void_p), ("energies", C.c_void_p), ("status", C.c_void_p)]


class ObservableSpec(C.Structure):
    _fields_ = [("base_pairs", C.c_void_p), ("n_base_pairs", C.c_int32), ("n_quartets", C.c_int32), ("quartets", C.c_void_p),
                ("sigma_backbone", C.c_double)]


N_OBS = 4
OBS_PROPELLER, OBS_RISE, OBS_PITCH_ANGLE, OBS_DIAMETER = range(4)


class NlArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("n_frames", C.c_int32),
        ("center", C.c_void_p),
        ("bonded", C.c_void_p),
        ("n_bonded", C.c_int32),
        ("box", C.c_double * 3),
        ("r_cutoff", C.c_double),
        ("dr_threshold", C.c_double),
        ("pairs", C.c_void_p),
        ("capacity", C.c_int64),
        ("count", C.c_void_p),
        ("overflow", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("flags", C.c_uint32),
        ("_pad", C.c_uint32),
        ("max_row", C.c_void_p),
        ("tag_bits", C.c_uint32),
        ("_pad2", C.c_uint32),
        ("append_count", C.c_void_p),
        ("lane_slots", C.c_int32),
        ("_pad3", C.c_int32),
        ("slot_base", C.c_int64),
        ("slot_width", C.c_int64),
        ("reference", C.c_void_p),
        ("move_threshold", C.c_double),
        ("rebuilds", C.c_void_p),
    ]


class LangevinArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("center", C.c_void_p),
        ("quat", C.c_void_p),
        ("p_center", C.c_void_p),
        ("p_quat", C.c_void_p),
        ("d_center", C.c_void_p),
        ("d_quat", C.c_void_p),
        ("dt", C.c_double),
        ("kT", C.c_double),
        ("gamma_center", C.c_double),
        ("gamma_quat", C.c_double),
        ("mass", C.c_double),
        ("inertia", C.c_double * 3),
        ("box", C.c_double * 3),
        ("seed", C.c_uint64),
        ("step", C.c_uint64),
        ("noise", C.c_void_p),
        ("phase", C.c_int32),
        ("advance_step", C.c_int32),
        ("step_ptr", C.c_void_p),
        ("traj_center", C.c_void_p),
        ("traj_quat", C.c_void_p),
        ("traj_rows", C.c_int64),
        ("zero_forces", C.c_int32),
        ("_pad2", C.c_int32),
    ]


class LangevinAdjointArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("phase", C.c_int32)] + [(k, C.c_void_p) for k in (
        "center", "quat", "p_center", "p_quat", "d_center", "d_quat", "noise", "lam_center", "lam_quat", "lam_p_center", "lam_p_quat",
        "lam_force_center", "lam_force_quat")] + [(k, C.c_double) for k in ("dt", "kT", "gamma_center", "gamma_quat", "mass")] + [
        ("inertia", C.c_double * 3), ("seed", C.c_uint64), ("step", C.c_uint64)]


class WeightsArgs(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32),
        ("beta", C.c_void_p),
        ("e_new", C.c_void_p),
        ("e_ref", C.c_void_p),
        ("weights", C.c_void_p),
        ("sums", C.c_void_p),
    ]


_SIGNATURES = {
    "mythos_b200_energy_f64": (C.c_int, [C.c_void_p, C.POINTER(EnergyArgs)]),
    "mythos_b200_energy_f32": (C.c_int, [C.c_void_p, C.POINTER(EnergyArgs)]),
    "mythos_b200_frame_kernel_fits": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "mythos_b200_energy_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "mythos_b200_backbone_sites_f64": (C.c_int, [C.c_void_p, C.POINTER(Model), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "mythos_b200_backbone_sites_f32": (C.c_int, [C.c_void_p, C.POINTER(Model), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "mythos_b200_support_points_f64": (C.c_int, [C.c_void_p, C.POINTER(Model), C.c_int32, C.c_int32] + [C.c_void_p] * 6),
    "mythos_b200_support_points_f32": (C.c_int, [C.c_void_p, C.POINTER(Model), C.c_int32, C.c_int32] + [C.c_void_p] * 6),
    "mythos_b200_nl_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "mythos_b200_nl_conditional_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "mythos_b200_nl_build_f64": (C.c_int, [C.c_void_p, C.POINTER(NlArgs)]),
    "mythos_b200_nl_build_f32": (C.c_int, [C.c_void_p, C.POINTER(NlArgs)]),
    "mythos_b200_langevin_f64": (C.c_int, [C.c_void_p, C.POINTER(LangevinArgs)]),
    "mythos_b200_langevin_f32": (C.c_int, [C.c_void_p, C.POINTER(LangevinArgs)]),
    "mythos_b200_langevin_adjoint_f64": (C.c_int, [C.c_void_p, C.POINTER(LangevinAdjointArgs)]),
    "mythos_b200_langevin_adjoint_f32": (C.c_int, [C.c_void_p, C.POINTER(LangevinAdjointArgs)]),
    "mythos_b200_weights_neff_f64": (C.c_int, [C.c_void_p, C.POINTER(WeightsArgs)]),
    "mythos_b200_weights_neff_f32": (C.c_int, [C.c_void_p, C.POINTER(WeightsArgs)]),
    "mythos_b200_fma_peak_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mythos_b200_fma_peak_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mythos_b200_special_rate_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mythos_b200_special_rate_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mythos_b200_observables_f64": (C.c_int, [C.c_void_p, C.POINTER(Model), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ObservableSpec), C.c_void_p]),
    "mythos_b200_observables_f32": (C.c_int, [C.c_void_p, C.POINTER(Model), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ObservableSpec), C.c_void_p]),
    "mythos_b200_traj_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "mythos_b200_traj_index": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_void_p]),
    "mythos_b200_traj_parse_f64": (C.c_int, [C.c_void_p, C.POINTER(TrajArgs)]),
    "mythos_b200_traj_parse_f32": (C.c_int, [C.c_void_p, C.POINTER(TrajArgs)]),
    "mythos_b200_theta_tape_forward": (C.c_int, [C.c_void_p] * 4),
    "mythos_b200_theta_tape_vjp": (C.c_int, [C.c_void_p] * 5),
    "mythos_b200_abi_version": (C.c_int, []),
    "mythos_b200_param_count": (C.c_int, []),
    "mythos_b200_param_name": (C.c_char_p, [C.c_int]),
    "mythos_b200_param_index": (C.c_int, [C.c_char_p]),
    "mythos_b200_last_error": (C.c_char_p, []),
    "mythos_b200_sizeof_model": (C.c_size_t, []),
    "mythos_b200_sizeof_energy_args": (C.c_size_t, []),
    "mythos_b200_sizeof_nl_args": (C.c_size_t, []),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    """Load the shared object (built by ``mythos_b200.build``); raises if it is absent."""
    if not LIB_PATH.exists():
        raise MythosB200Error(
            f"{LIB_PATH} not found: build it with `python -m mythos_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this path."
        )
    handle = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    if (handle.mythos_b200_sizeof_model() != C.sizeof(Model) or handle.mythos_b200_sizeof_energy_args() != C.sizeof(EnergyArgs)
            or handle.mythos_b200_sizeof_nl_args() != C.sizeof(NlArgs)):
        raise MythosB200Error("ctypes struct layout does not match include/mythos_b200.h")
    return handle


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().mythos_b200_last_error().decode()
        err = MythosB200Error(f"{what}: {STATUS.get(status, status)}: {msg}")
        err.status = status
        raise err


@lru_cache(maxsize=1)
def param_names() -> tuple[str, ...]:
    """Kernel-level parameter names of one bank, in bank order ("<term>.<reference name>")."""
    handle = lib()
    names = []
    i = 0
    while (s := handle.mythos_b200_param_name(i)) is not None:
        names.append(s.decode())
        i += 1
    return tuple(names)


@lru_cache(maxsize=1)
def param_count() -> int:
    return int(lib().mythos_b200_param_count())


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def current_stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise MythosB200Error(
            f"{what} must live on a CUDA device: the oxDNA energy path runs only as sm_100a kernels (no CPU fallback)"
        )


def suffix(dtype: torch.dtype) -> str:
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise MythosB200Error(f"unsupported dtype {dtype}: float32 or float64")
