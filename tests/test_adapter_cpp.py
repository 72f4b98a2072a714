"""The C++ host adapter (include/relion_b200_adapter.hpp): AccProjector / AccBackprojector / MlDeviceBundle /
MlOptimiserCuda with the reference's names and call sequence (src/acc/cuda/cuda_ml_optimiser.h:18-144), driven through
tests/cpp/adapter_shim.cpp the way src/ml_optimiser.cpp:3577-3869 drives the reference's objects."""
import ctypes as C
import os

import numpy as np
import pytest

from relion_b200 import capi
from relion_b200.estep import marshal_model, marshal_sampling, marshal_pool, make_pool_out
from relion_b200.workload import make_workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "cpp", "libadapter_shim.so")


def _shim():
    if not os.path.exists(SHIM):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(SHIM)
    lib.adapter_estep.restype = C.c_int
    return lib


def _run(wl, device=0, skip_maximization=False, nr_threads=3):
    lib = _shim()
    K = wl.model.nr_classes
    mm, ms, mp = marshal_model(wl.model), marshal_sampling(wl.sampling), marshal_pool(wl.pool)
    out = make_pool_out(wl.pool.n_particles, wl.model.ori_size // 2 + 1, K, wl.sampling.n_dir)
    refs = [np.ascontiguousarray(v if v.ndim == 3 else v[None], np.complex128) for v in wl.refs]   # MlModel::PPref is double
    z, y, x = refs[0].shape
    init = -((y - 1) // 2)
    bz, by, bx = wl.bp_shape if len(wl.bp_shape) == 3 else (1,) + tuple(wl.bp_shape)
    ref_dims = np.array([[x, y, z, init, init if z > 1 else 0, wl.r_max]] * K, np.int32)
    bp_dims = np.array([[bx, by, bz, -((by - 1) // 2), -((bz - 1) // 2) if bz > 1 else 0, wl.r_max]] * K, np.int32)
    pp = (C.c_void_p * K)(*[r.ctypes.data for r in refs])
    bps = [[np.zeros((bz, by, bx), np.float32) for _ in range(K)] for _ in range(3)]
    ptrs = [(C.c_void_p * K)(*[a.ctypes.data for a in arrs]) for arrs in bps]
    err = C.create_string_buffer(1024)
    rc = lib.adapter_estep(device, C.byref(mm.struct), C.byref(ms.struct), K, pp, ref_dims.ctypes.data_as(C.c_void_p),
                           bp_dims.ctypes.data_as(C.c_void_p), C.c_float(wl.padding_factor), C.byref(mp.struct), C.byref(out.struct),
                           int(skip_maximization), nr_threads, ptrs[0], ptrs[1], ptrs[2], err, len(err))
    return rc, err.value.decode(), out.result, bps


def test_adapter_without_gpu_throws_relion_error():
    """No CPU fallback: MlDeviceBundle::setDevice throws (the reference's HANDLE_ERROR -> REPORT_ERROR) when there is no device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    wl = make_workload(ori_size=16, healpix_order=1, n_particles=1, seed=3)
    rc, msg, _, _ = _run(wl)
    assert rc == -1
    assert "relion_b200" in msg and "no CUDA device" in msg and "relion_b200_adapter.hpp" in msg, msg


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(ori_size=32, healpix_order=1, n_particles=10, nr_classes=2, seed=51, snr=0.3),
                                dict(ori_size=32, healpix_order=2, n_particles=6, nr_classes=1, seed=52, snr=0.2, local_search=True)])
def test_adapter_estep_equals_python_host_mirror_and_oracle(device, kw):
    from oracle.bindings import Oracle, Projector, Backprojector
    wl = make_workload(**kw)
    rc, msg, res, bps = _run(wl)
    assert rc == 0, msg
    # same library behind both host sides: identical decisions, sums equal up to the order of the atomics
    device.set_model(wl.model); device.set_sampling(wl.sampling)
    for k, v in enumerate(wl.refs):
        device.set_reference(k, v.astype(np.complex128), wl.r_max, wl.padding_factor)
        device.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)
    py = device.expectation_some_particles(wl.pool)
    for f in ("best_ihidden_over", "nr_significant_coarse", "n_fine_samples"):
        assert np.array_equal(res.particles[f], py.particles[f]), f
    np.testing.assert_allclose(res.particles["dLL_nolog"], py.particles["dLL_nolog"], rtol=1e-6)
    np.testing.assert_allclose(res.wsum_pdf_class, py.wsum_pdf_class, rtol=1e-6)
    for k in range(wl.model.nr_classes):
        gre, gim, gw = device.bp_get(k)
        for a, b in ((bps[0][k], gre), (bps[1][k], gim), (bps[2][k], gw)):
            assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-12)
    # and against the oracle
    o = Oracle("port")
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    obp = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
    st, ores, _ = o.estep_pool(wl.model, wl.sampling, refs, obp, wl.pool, num_threads=0, exact_threshold=True)
    assert st == 0
    assert np.mean(res.particles["best_ihidden_over"] == ores.particles["best_ihidden_over"]) >= 0.995
    np.testing.assert_allclose(res.particles["dLL_nolog"], ores.particles["dLL_nolog"], rtol=1e-4)
    for k in range(wl.model.nr_classes):
        assert np.abs(bps[2][k] - obp[k].weight).max() <= 5e-3 * np.abs(obp[k].weight).max()


@pytest.mark.gpu
def test_adapter_error_behaviour_on_the_device():
    """A bad model is reported like the reference reports CUDA errors: RelionError with the library's message."""
    wl = make_workload(ori_size=16, healpix_order=1, n_particles=1, seed=3)
    wl.model.current_size = 18            # > ori_size
    rc, msg, _, _ = _run(wl)
    assert rc == -1 and "current_size" in msg, msg
