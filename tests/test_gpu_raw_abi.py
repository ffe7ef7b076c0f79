"""The C-ABI stands alone: ``mythos_b200_energy_f64`` driven with raw ``cudaMalloc`` buffers through ctypes -- no torch
tensor, no torch stream, no allocator of ours anywhere in the call -- on the reference's own dna1 golden frames
(``data/test-data/dna1/simple-helix``), checked against oxDNA's ``split_energy.dat`` and against the oracle's autograd
forces.  This is what an XLA custom-call handler (or any other host) does with the library."""

import ctypes as C

import numpy as np
import pytest

from mythos_b200 import _lib
from tests.golden_cases import TOL, load_case, theta_for
from tests.product_cases import energy_fn_of

pytestmark = pytest.mark.gpu


class _Cuda:
    def __init__(self):
        _lib.lib()  # loads libmythos_b200.so and with it the CUDA runtime it links
        self.rt = C.CDLL("libcudart.so.12")
        self.rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        self.rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        self.rt.cudaFree.argtypes = [C.c_void_p]
        self.rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]
        self.rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
        self.rt.cudaStreamDestroy.argtypes = [C.c_void_p]
        self.live = []

    def ok(self, code):
        assert code == 0, f"CUDA runtime error {code}"

    def upload(self, arr: np.ndarray) -> C.c_void_p:
        arr = np.ascontiguousarray(arr)
        p = C.c_void_p()
        self.ok(self.rt.cudaMalloc(C.byref(p), max(arr.nbytes, 8)))
        self.ok(self.rt.cudaMemcpy(p, arr.ctypes.data, arr.nbytes, 1))
        self.live.append(p)
        return p

    def zeros(self, nbytes: int) -> C.c_void_p:
        p = C.c_void_p()
        self.ok(self.rt.cudaMalloc(C.byref(p), max(nbytes, 8)))
        self.ok(self.rt.cudaMemset(p, 0, max(nbytes, 8)))
        self.live.append(p)
        return p

    def download(self, p, shape, dtype=np.float64) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        self.ok(self.rt.cudaMemcpy(out.ctypes.data, p, out.nbytes, 2))
        return out

    def free_all(self):
        for p in self.live:
            self.rt.cudaFree(p)
        self.live = []


def test_energy_f64_from_raw_cuda_buffers():
    import torch  # only for the oracle's autograd below and to pack the host-side parameter vector

    from mythos_b200.energy import model as kmodel
    from oracle import oxdna_oracle as orc

    case = load_case("dna1_simple_helix")
    efn = energy_fn_of(case)
    plan = kmodel.plan_for(efn.energy_fns)
    bank = plan.params_vector().detach().numpy().astype(np.float64)
    F, n = 4, case["center"].shape[1]
    center, quat = case["center"][:F].astype(np.float64), case["quat"][:F].astype(np.float64)
    pairs = np.asarray(case["pairs"], dtype=np.int32).reshape(2, -1)
    bonded = np.asarray(case["bonded"], dtype=np.int32).reshape(-1, 2)

    cu = _Cuda()
    stream = C.c_void_p()
    cu.ok(cu.rt.cudaStreamCreate(C.byref(stream)))
    try:
        a = _lib.EnergyArgs()
        a.model = C.pointer(plan.model)
        a.n, a.n_frames = n, F
        a.center, a.quat = cu.upload(center), cu.upload(quat)
        a.seq = cu.upload(np.asarray(case["seq"], dtype=np.int32))
        a.bonded, a.n_bonded = cu.upload(bonded), bonded.shape[0]
        a.pairs, a.pair_capacity, a.pair_frame_stride = cu.upload(pairs), pairs.shape[1], 0
        a.params = cu.upload(bank)
        a.term_mask = plan.term_mask
        a.flags = 0
        a.terms = cu.zeros(F * 8 * 8)
        a.d_center, a.d_quat = cu.zeros(F * n * 3 * 8), cu.zeros(F * n * 4 * 8)
        a.d_params = cu.zeros(bank.size * 8)
        need = _lib.lib().mythos_b200_energy_workspace_bytes(n, F, pairs.shape[1], 8)
        a.workspace, a.workspace_bytes = cu.zeros(int(need)), int(need)
        assert _lib.lib().mythos_b200_energy_f64(stream, C.byref(a)) == 0, _lib.lib().mythos_b200_last_error()
        cu.ok(cu.rt.cudaStreamSynchronize(stream))
        terms = cu.download(a.terms, (F, 8))
        d_center, d_quat = cu.download(a.d_center, (F, n, 3)), cu.download(a.d_quat, (F, n, 4))
        d_params = cu.download(a.d_params, (bank.size,))
    finally:
        cu.rt.cudaStreamDestroy(stream)
        cu.free_all()

    # oxDNA's own per-nucleotide split energies (6 decimals), the reference's tolerances
    want = case["golden_terms_per_nt"][:F]
    for k in range(want.shape[1]):
        np.testing.assert_allclose(np.around(terms[:, k] / n, 6), want[:, k], atol=TOL["dna1"][k], rtol=1e-7)
    # forces / dE/dquat of frame 0 against the oracle's autograd
    params = orc.init_all("dna1", theta_for(case))
    ct, qt = torch.tensor(center[0], requires_grad=True), torch.tensor(quat[0], requires_grad=True)
    e = orc.energy_terms("dna1", ct, qt, case["seq"], case["bonded"], case["pairs"], params, box=20.0).sum()
    gc, gq = torch.autograd.grad(e, [ct, qt])
    np.testing.assert_allclose(d_center[0], gc.numpy(), rtol=1e-6, atol=1e-8 * float(gc.abs().max()))
    np.testing.assert_allclose(d_quat[0], gq.numpy(), rtol=1e-6, atol=1e-8 * float(gq.abs().max()))
    assert np.isfinite(d_params).all() and np.abs(d_params).sum() > 0


def test_bad_arguments_are_refused_without_touching_the_device():
    a = _lib.EnergyArgs()
    assert _lib.lib().mythos_b200_energy_f64(None, C.byref(a)) != 0
    assert _lib.lib().mythos_b200_last_error()
