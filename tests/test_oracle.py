"""CPU-only tests: the oracle against the reference's golden vectors, host logic, and the C-ABI surface.

* tests/golden/estep_reference_*.npz were produced by the reference's own ALTCPU kernels compiled from
  /root/reference (tests/golden/make_golden.py).  The restated kernels must reproduce them.
* tests/ctf.cpp:5-10 of the reference is the only known-answer test in its tree (CTF value 0.59154).
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from relion_b200 import sampling as smp
from relion_b200 import synth
from relion_b200.workload import make_workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _check_against(d, g, rtol=5e-5):
    for f in g["particles"].dtype.names:
        a, b = d["particles"][f], g["particles"][f]
        if a.dtype.kind in "iu":
            assert np.array_equal(a, b), f
        else:
            np.testing.assert_allclose(a, b, rtol=1e-3 if "weight" in f or f in ("pmax", "wsum_XA", "wsum_AA") else rtol, err_msg=f)
    assert np.array_equal(d["coarse_significant"], g["coarse_significant"])
    assert np.array_equal(d["fine_ihidden_over"], g["fine_ihidden_over"])
    m = g["coarse_diff2"] > -1e30
    assert np.array_equal(m, d["coarse_diff2"] > -1e30)
    np.testing.assert_allclose(d["coarse_diff2"][m], g["coarse_diff2"][m], rtol=rtol)
    np.testing.assert_allclose(d["fine_diff2"], g["fine_diff2"], rtol=rtol)
    np.testing.assert_allclose(d["coarse_weights"], g["coarse_weights"], rtol=2e-3, atol=1e-30)
    np.testing.assert_allclose(d["fine_weights"], g["fine_weights"], rtol=2e-3, atol=1e-30)
    for k in ("wdiff2s_parts", "wdiff2s_AA", "wdiff2s_XA", "shells", "pdf_direction", "pdf_class"):
        np.testing.assert_allclose(d[k], g[k], rtol=1e-3, atol=1e-5 * max(np.abs(g[k]).max(), 1e-30), err_msg=k)
    for k in g.files:
        if k.startswith("bp") and k.endswith("_sample"):
            assert np.abs(d[k] - g[k]).max() <= 1e-4 * max(np.abs(g[k]).max(), 1e-30), k
        if k.startswith("bp") and k.endswith("_abs"):
            np.testing.assert_allclose(d[k], g[k], rtol=1e-4, err_msg=k)


@pytest.mark.parametrize("name", ["global_k2", "local_k1", "window_k1", "cc_global_k1", "cc_window_k1", "mag_local_k1", "box_global_k1"])
def test_port_matches_reference_golden(name):
    mod = _cases()
    g = np.load(os.path.join(GOLD, f"estep_reference_{name}.npz"))
    d = mod.run_case("port", mod.CASES[name])
    _check_against(d, g)


@pytest.mark.parametrize("name", ["global_k2", "cc_global_k1", "mag_local_k1"])
def test_compiled_reference_matches_golden(name):
    from oracle.bindings import have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/librefkernels.so not built (needs /root/reference)")
    mod = _cases()
    g = np.load(os.path.join(GOLD, f"estep_reference_{name}.npz"))
    d = mod.run_case("reference", mod.CASES[name])
    _check_against(d, g, rtol=1e-6)


def test_euler_matrices_with_left_and_right_matrices_port_matches_compiled_reference():
    """cpu_kernel_make_eulers_3D<invert, doL, doR> (cpu_kernels/helper.cpp:744-875) for the four (doL, doR) instantiations:
    inverse(L A R) as adjugate / determinant with L, the plain transpose without it (R alone included)."""
    from oracle.bindings import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/librefkernels.so not built (needs /root/reference)")
    port, ref = Oracle("port"), Oracle("reference")
    rng = np.random.default_rng(31)
    rot, tilt, psi = rng.uniform(-180, 180, 50), rng.uniform(0, 180, 50), rng.uniform(0, 360, 50)
    L = np.array([[1.03, 0.012, 0.0], [-0.008, 0.96, 0.0], [0.0, 0.0, 1.0]]) * 0.8
    c, s = np.cos(0.3), np.sin(0.3)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    for ml, mr in ((None, None), (L, None), (None, R), (L, R)):
        a, b = port.make_eulers(rot, tilt, psi, ml, mr), ref.make_eulers(rot, tilt, psi, ml, mr)
        np.testing.assert_allclose(a, b, rtol=0, atol=3e-7)
        if ml is not None:
            A = synth.inverse_euler_f32(rot, tilt, psi).reshape(-1, 3, 3).astype(np.float64).transpose(0, 2, 1)
            want = np.linalg.inv(ml @ A @ (np.eye(3) if mr is None else mr)).reshape(-1, 9)
            np.testing.assert_allclose(a, want, atol=2e-6)


def test_cc_kernels_port_matches_compiled_reference():
    """Kernel level: the restated cross-correlation kernels against the reference's own diff2_CC_coarse_2D / diff2_CC_fine_2D
    (cpu_kernels/diff2.h:611-742, 904-1050) on random inputs, including a projector whose r_max cuts the window."""
    from oracle.bindings import Oracle, Projector, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/librefkernels.so not built (needs /root/reference)")
    port, ref = Oracle("port"), Oracle("reference")
    wl = make_workload(ori_size=32, n_particles=2, seed=12)
    rng = np.random.default_rng(4)
    for n, r_max in ((32, wl.r_max), (20, wl.r_max), (32, 9)):
        pj = Projector(wl.refs[0], r_max, wl.padding_factor)
        xs = n // 2 + 1
        O, T = 9, 7
        eul = synth.inverse_euler_f32(rng.uniform(-180, 180, O), rng.uniform(0, 180, O), rng.uniform(0, 360, O))
        tx = (-2 * np.pi * rng.uniform(-4, 4, T) / 32).astype(np.float32)
        ty = (-2 * np.pi * rng.uniform(-4, 4, T) / 32).astype(np.float32)
        re = rng.standard_normal((n, xs)).astype(np.float32)
        im = rng.standard_normal((n, xs)).astype(np.float32)
        corr = rng.uniform(0.5, 2.0, (n, xs)).astype(np.float32)
        a = port.diff2_coarse(pj, n, eul, tx, ty, re, im, corr, cc=True)
        b = ref.diff2_coarse(pj, n, eul, tx, ty, re, im, corr, cc=True)
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
        rot_idx = np.repeat(np.arange(O), T); trans_idx = np.tile(np.arange(T), O)
        job_idx = np.arange(0, O * T, T); job_num = np.full(O, T)
        a = port.diff2_cc_fine(pj, n, eul, tx, ty, re, im, corr, rot_idx, trans_idx, job_idx, job_num)
        b = ref.diff2_cc_fine(pj, n, eul, tx, ty, re, im, corr, rot_idx, trans_idx, job_idx, job_num)
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
        # the two kernels compute the same quantity (fine = coarse on the same orientation / translation grid)
        np.testing.assert_allclose(a.reshape(O, T), port.diff2_coarse(pj, n, eul, tx, ty, re, im, corr, cc=True), rtol=2e-5, atol=2e-6)


def test_ctf_known_answer():
    # tests/ctf.cpp:5-10: setValues(10000, 12000, 90, 300, 2.7, 0.1, 0, 1, 0); getCTF(10, 10) == Approx(0.59154)
    c = synth.CTF(10000.0, 12000.0, 90.0, 300.0, 2.7, 0.1, 0.0, 1.0, 0.0)
    assert float(c.get_ctf(10.0, 10.0)) == pytest.approx(0.59154, rel=1e-4)


def test_sampling_tables():
    # HEALPix: 12*4^order pixels, unit vectors cover the sphere evenly (SURVEY.md §8 sizes)
    for order, ndir, npsi in ((1, 48, 12), (2, 192, 24), (3, 768, 48)):
        s = smp.make_sampling(order, 5.0, 2.0, oversampling=1, build_oversampled=(order < 3))
        assert s.n_dir == ndir and s.n_psi == npsi
        v = np.stack([np.sin(np.radians(s.tilt)) * np.cos(np.radians(s.rot)),
                      np.sin(np.radians(s.tilt)) * np.sin(np.radians(s.rot)), np.cos(np.radians(s.tilt))], 1)
        assert np.abs(v.mean(0)).max() < 1e-9
        assert s.n_trans == 21 and s.n_over_trans == 4 and s.n_over_rot == 8
    assert smp.make_sampling(1, 3.0, 2.0).n_trans == 9
    assert smp.make_sampling(1, 5.0, 1.0).n_trans == 81
    s = smp.make_sampling(1, 3.0, 2.0)
    # children of a coarse pixel lie closer to it than to any other coarse pixel
    d = smp._direction(s.rot, s.tilt)
    g = (5 * s.n_psi + 3) * 8
    child = smp._direction(s.over_rot[g:g + 8], s.over_tilt[g:g + 8])
    assert np.all(np.argmax(child @ d.T, axis=1) == 5)
    # oversampled translations average back to the coarse one
    np.testing.assert_allclose(s.over_trans_x.reshape(-1, 4).mean(1), s.trans_x, atol=1e-12)
    # nested <-> xyf round trip
    ip = np.arange(12 * 16)
    x, y, f = smp.nest2xyf(2, ip)
    assert np.array_equal(smp.xyf2nest(2, x, y, f), ip)


def test_local_search_lists():
    s = smp.make_sampling(2, 3.0, 1.0)
    di, dp, pi_, pp = smp.select_nonzero_prior(s, 30.0, 60.0, 100.0, 7.5, 7.5, 7.5)
    assert 0 < len(di) < s.n_dir and 0 < len(pi_) < s.n_psi
    assert dp.sum() == pytest.approx(1.0) and pp.sum() == pytest.approx(1.0)
    # far-away prior with a tiny sigma still selects the nearest direction
    di, dp, _, _ = smp.select_nonzero_prior(s, 30.0, 60.0, 100.0, 0.01, 0.01, 0.01)
    assert len(di) == 1 and dp[0] == 1.0


def test_reference_ft_central_slice_matches_projection_fft():
    # a central slice of PPref equals the normalised 2D FFT of the real-space projection
    n = 24
    vol = synth.make_phantom(n, n_blobs=10, seed=3)
    data, r_max = synth.reference_ft(vol, padding_factor=2.0, do_gridding=False)
    sl = synth.project_numpy(data, r_max, 2.0, np.eye(3), n)
    proj = vol.sum(axis=0)                       # project along z (identity orientation)
    F = np.fft.rfft2(np.fft.ifftshift(proj)) / (n * n)
    M = synth.mresol(n)
    m = (M >= 0) & (M < n // 2 - 1)
    assert np.abs(sl[m] - F[m]).max() <= 2e-3 * np.abs(F[m]).max()


def test_significance_exact_vs_sequential():
    from oracle.bindings import Oracle
    o = Oracle("port")
    rng = np.random.default_rng(0)
    differ = 0
    for i in range(200):
        n = int(rng.integers(2, 3000))
        w = np.exp(rng.normal(0, 4, n)).astype(np.float32)
        w[rng.random(n) < 0.1] = 0
        a = o.significance(w, 0.999, 0, True, exact=False)
        b = o.significance(w, 0.999, 0, True, exact=True)
        assert a["n_filtered"] == b["n_filtered"]
        assert abs(a["threshold_idx"] - b["threshold_idx"]) <= 1      # only rounding-edge cases may move by one
        differ += a["threshold_idx"] != b["threshold_idx"]
        np.testing.assert_allclose(a["sum_weight"], b["sum_weight"], rtol=1e-5)
    assert differ <= 10
    # maxsig caps the count
    w = np.arange(1, 101, dtype=np.float32)
    a = o.significance(w, 0.5, 7, True, exact=False)
    assert a["n_filtered"] - a["threshold_idx"] == 7 and a["significant_weight"] == 94.0


def test_library_exports_every_declared_symbol():
    from relion_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "relion_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rb_ctx"}
    assert declared == set(capi.PROTOTYPES), declared ^ set(capi.PROTOTYPES)
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rb_version() == 100
    # struct layouts agree with the C header (compile a probe with the host compiler)
    src = '#include "relion_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(rb_sampling), sizeof(rb_model), sizeof(rb_particles), sizeof(rb_particle_out), sizeof(rb_pool_out), sizeof(rb_weights_out));return 0;}'
    exe = os.path.join(ROOT, "tests", "_probe_sizes")
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src.encode(), check=True)
    sizes = list(map(int, subprocess.run([exe], capture_output=True, check=True).stdout.split()))
    os.remove(exe)
    assert sizes == [ctypes.sizeof(t) for t in (capi.rb_sampling, capi.rb_model, capi.rb_particles, capi.rb_particle_out,
                                                 capi.rb_pool_out, capi.rb_weights_out)]


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from relion_b200 import capi
    from relion_b200.estep import MlDeviceBundle
    with pytest.raises(capi.RelionB200Error) as e:
        MlDeviceBundle(0)
    assert e.value.status == capi.RB_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "relion_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                t = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|oracle/|liboracle|librefkernels|oracle_kernels\.h|estep_driver", t):
                    if "must not reference anything under oracle/" in t or "no reference to oracle/" in t:
                        t2 = re.sub(r".*(must not reference anything under|no reference to) oracle/.*", "", t)
                        if not re.search(r"(from|import)\s+oracle|oracle/|liboracle|librefkernels|oracle_kernels\.h|estep_driver", t2):
                            continue
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_reconstruct_recovers_phantom():
    """oracle/reconstruct.py (restated BackProjector::reconstruct, skip_gridding) turns the oracle's accumulators back
    into the phantom the particles were projected from: FSC ~ 1 at low resolution, correct absolute scale."""
    from relion_b200 import synth
    from relion_b200.workload import make_workload
    from oracle.bindings import Oracle, Projector, Backprojector
    from oracle import reconstruct as rc
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=200, nr_classes=1, seed=31, snr=1.0)
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
    st, _, _ = Oracle("port").estep_pool(wl.model, wl.sampling, refs, bps, wl.pool, num_threads=0, exact_threshold=True)
    assert st == 0
    m = rc.reconstruct(bps[0].real, bps[0].imag, bps[0].weight, 32, wl.r_max, 2.0)
    vol = synth.make_phantom(32, n_blobs=40, seed=1993)
    f = rc.fsc(m, vol)
    assert f[:5].min() > 0.95, f
    assert 0.8 < (m * vol).sum() / (vol * vol).sum() < 1.2
    assert abs(rc.fsc(vol, vol) - 1).max() < 1e-12


def test_backproject_posed_port_matches_numpy_restatement():
    """Two independent restatements of BackProjector::backproject2Dto3D (compiled C++ used as the CPU baseline of the
    reconstruct workload, numpy used by the GPU parity test) agree to rounding."""
    from oracle.bindings import backproject_posed
    from oracle.backproject_posed import backproject2Dto3D
    from relion_b200 import synth
    rng = np.random.default_rng(5)
    n, count, r_max = 20, 9, 9
    xs = n // 2 + 1
    pad = synth.pad_size_for(r_max, 2.0)
    shape = (pad, pad, pad // 2 + 1)
    F = (rng.standard_normal((count, n, xs)) + 1j * rng.standard_normal((count, n, xs))).astype(np.complex64)
    W = rng.uniform(-0.1, 1.0, (count, n, xs)).astype(np.float32)
    eul = synth.inverse_euler_f32(rng.uniform(-180, 180, count), rng.uniform(0, 180, count), rng.uniform(0, 360, count))
    re, im, w = backproject_posed(shape, F, W, eul, r_max)
    data = np.zeros(shape, np.complex128); weight = np.zeros(shape, np.float64)
    for i in range(count):
        backproject2Dto3D(data, weight, F[i], eul[i].reshape(3, 3).astype(np.float64), W[i], r_max, 2.0)
    assert weight.max() > 0
    np.testing.assert_allclose(re, data.real, rtol=0, atol=1e-12 * np.abs(data).max())
    np.testing.assert_allclose(im, data.imag, rtol=0, atol=1e-12 * np.abs(data).max())
    np.testing.assert_allclose(w, weight, rtol=0, atol=1e-12 * weight.max())


def _inplane_eulers(psi_deg):
    from relion_b200 import synth
    psi = np.asarray(psi_deg, np.float64)
    return synth.inverse_euler_f32(np.zeros_like(psi), np.zeros_like(psi), psi)


def test_2d_embedding_matches_reference_2d_kernels():
    """2D references (2D classification, BASELINE config #1) are handed to the projection-type kernels as a two-plane
    volume whose second plane is zero.  With in-plane rotations zp == 0, so project3Dmodel must reproduce the reference's
    own project2Dmodel bit for bit, and diff2_coarse<REF3D> its <REF2D> instantiation to summation-order rounding.
    The 2D back-projection has its own kernel (no circle bound on x): port vs the reference's backproject2D."""
    import ctypes as C
    from oracle.bindings import Oracle, Projector, Backprojector, have_reference, REF_LIB
    from relion_b200 import synth
    if not have_reference():
        pytest.skip("oracle/_ref/librefkernels.so not built")
    ref = Oracle("reference"); port = Oracle("port")
    lib = C.CDLL(REF_LIB)
    f32p = C.POINTER(C.c_float)
    fp = lambda a: a.ctypes.data_as(f32p)
    rng = np.random.default_rng(3)
    n = 24; xs = n // 2 + 1; r_max = n // 2
    img = synth.make_phantom_2d(n, seed=4)
    data, _ = synth.reference_ft_2d(img, padding_factor=2.0)
    data = data.astype(np.complex64)
    pad, mx = data.shape
    inity = -((pad - 1) // 2)
    P = Projector(data, r_max, 2.0)
    assert P.vol.shape == (2, pad, mx) and not P.vol[1].any()
    eul = _inplane_eulers(rng.uniform(0, 360, 7))
    for e in eul:
        want_re = np.zeros((n, xs), np.float32); want_im = np.zeros((n, xs), np.float32)
        lib.refk2d_project(fp(np.ascontiguousarray(data).view(np.float32)), mx, pad, inity, r_max, C.c_float(2.0), xs, n, fp(e), fp(want_re), fp(want_im))
        for orc in (ref, port):
            got = orc.project(P, n, e)
            assert np.array_equal(got.real, want_re) and np.array_equal(got.imag, want_im), orc.kind
    # coarse diff2
    O, T = 300, 5
    eul = _inplane_eulers(rng.uniform(0, 360, O))
    tx = (-2 * np.pi * rng.uniform(-3, 3, T) / n).astype(np.float32); ty = (-2 * np.pi * rng.uniform(-3, 3, T) / n).astype(np.float32)
    re = rng.standard_normal((n, xs)).astype(np.float32); im = rng.standard_normal((n, xs)).astype(np.float32)
    corr = rng.uniform(0.5, 2.0, (n, xs)).astype(np.float32)
    want = np.zeros((O, T), np.float32)
    lib.refk2d_diff2_coarse(fp(np.ascontiguousarray(data).view(np.float32)), mx, pad, inity, r_max, C.c_float(2.0), xs, n, fp(eul), C.c_ulong(O),
                            fp(tx), fp(ty), C.c_ulong(T), fp(re), fp(im), fp(corr), fp(want))
    for orc in (ref, port):
        np.testing.assert_allclose(orc.diff2_coarse(P, n, eul, tx, ty, re, im, corr), want, rtol=2e-6)
    # back-projection: restated 2D kernel against the reference's
    O, T = 9, 4
    eul = _inplane_eulers(rng.uniform(0, 360, O))
    weights = rng.random((O, T)).astype(np.float32); weights[rng.random((O, T)) < 0.4] = np.finfo(np.float32).min
    ctfs = rng.uniform(-1, 1, (n, xs)).astype(np.float32)
    tx, ty = tx[:T], ty[:T]
    b_ref = Backprojector((pad, mx), r_max, 2.0); b_port = Backprojector((pad, mx), r_max, 2.0)
    ref.backproject(b_ref, n, eul, tx, ty, re, im, weights, corr, ctfs, 2.5, 0.3)
    port.backproject(b_port, n, eul, tx, ty, re, im, weights, corr, ctfs, 2.5, 0.3)
    assert np.abs(b_ref.weight).max() > 0
    for a, b in ((b_port.real, b_ref.real), (b_port.imag, b_ref.imag), (b_port.weight, b_ref.weight)):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-6 * np.abs(b).max())


def test_2d_classification_pool_oracle():
    """The oracle driver on a 2D-classification pool (K = 3 classes, psi-only sampling): classes and in-plane angles of
    clean particles are recovered; port and reference-compiled kernels agree."""
    from relion_b200.workload import make_workload
    from oracle.bindings import Oracle, Projector, Backprojector, have_reference
    wl = make_workload(ori_size=32, n_particles=16, nr_classes=3, seed=5, snr=0.5, ref_dim=2, psi_step=12.0)
    assert wl.sampling.n_dir == 1 and wl.sampling.n_over_rot == 2 and wl.bp_shape == (67, 34)
    out = {}
    for kind in ["port"] + (["reference"] if have_reference() else []):
        refs = [Projector(v, wl.r_max, 2.0) for v in wl.refs]
        bps = [Backprojector(wl.bp_shape, wl.r_max, 2.0) for _ in wl.refs]
        st, res, _ = Oracle(kind).estep_pool(wl.model, wl.sampling, refs, bps, wl.pool, num_threads=0)
        assert st == 0
        out[kind] = (res.particles, bps)
        assert np.mean(res.particles["best_class"] == wl.truth["cls"]) == 1.0
        assert np.mean(res.particles["best_ipsi"] == wl.truth["ipsi"]) >= 0.9
    if "reference" in out:
        a, b = out["port"], out["reference"]
        assert np.array_equal(a[0]["best_ihidden_over"], b[0]["best_ihidden_over"])
        for k in range(3):
            np.testing.assert_allclose(a[1][k].weight, b[1][k].weight, rtol=0, atol=1e-4 * np.abs(b[1][k].weight).max())


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference times the reference's CPU kernels (oracle/_ref, else the port) and prints the contract's JSON
    line; under torchrun only rank 0 works."""
    import json
    import sys
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "8"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "particles/s" and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    env["RANK"] = "1"; env["WORLD_SIZE"] = "2"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_sgd_backprojection_port_matches_compiled_reference():
    """Gradient refinement: the restated residual back-projection against the reference's own CpuKernels::backproject3D_SGD
    (src/acc/cpu/cpu_kernels/BP.h:757-1047, compiled in oracle/_ref) on random weights, a random reference and random poses."""
    import ctypes as C
    from oracle.bindings import Oracle, Projector, Backprojector, _fp
    from relion_b200 import synth
    try:
        ref = Oracle("reference")
    except Exception as exc:                                            # no compiled reference on this machine
        pytest.skip(str(exc))
    port = Oracle("port")
    rng = np.random.default_rng(77)
    n, r_max, pf, O, T = 24, 10, 2.0, 5, 7
    xs = n // 2 + 1
    pad = synth.pad_size_for(r_max, pf)
    shape = (pad, pad, pad // 2 + 1)
    vol = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    eul = synth.inverse_euler_f32(rng.uniform(-180, 180, O), rng.uniform(0, 180, O), rng.uniform(0, 360, O)).astype(np.float32)
    img = (rng.standard_normal((n, xs)) + 1j * rng.standard_normal((n, xs))).astype(np.complex64)
    re, im = np.ascontiguousarray(img.real), np.ascontiguousarray(img.imag)
    tx = (rng.uniform(-3, 3, T) * 2 * np.pi / n).astype(np.float32); ty = (rng.uniform(-3, 3, T) * 2 * np.pi / n).astype(np.float32)
    w = rng.uniform(0, 1, (O, T)).astype(np.float32)
    minvs2 = rng.uniform(0.5, 2, (n, xs)).astype(np.float32); ctfs = rng.uniform(-1, 1, (n, xs)).astype(np.float32)
    outs = []
    for orc in (ref, port):
        proj = Projector(vol, r_max, pf)
        bp = Backprojector(shape, r_max, pf)
        orc.K.backproject_sgd(C.byref(bp.struct), C.byref(proj.struct), xs, n, _fp(re), _fp(im), _fp(tx), _fp(ty), _fp(w), _fp(minvs2), _fp(ctfs),
                              T, 0.3, 2.5, _fp(eul), O)
        outs.append((bp.real.copy(), bp.imag.copy(), bp.weight.copy()))
    assert np.abs(outs[0][2]).max() > 0
    for a, b in zip(outs[0], outs[1]):
        assert np.abs(a - b).max() <= 2e-5 * np.abs(a).max()


def test_sgd_backprojection_2d_port_matches_compiled_reference():
    """Gradient refinement of 2D references (RELION 4's default 2D classification): the restated residual back-projection against
    the reference's own CpuKernels::backproject2D_SGD (src/acc/cpu/cpu_kernels/BP.h:1047-1262)."""
    import ctypes as C
    from oracle.bindings import Oracle, Projector, Backprojector, _fp
    from relion_b200 import synth
    try:
        ref = Oracle("reference")
    except Exception as exc:
        pytest.skip(str(exc))
    port = Oracle("port")
    rng = np.random.default_rng(78)
    n, r_max, pf, O, T = 24, 10, 2.0, 6, 5
    xs = n // 2 + 1
    pad = synth.pad_size_for(r_max, pf)
    shape2 = (pad, pad // 2 + 1)
    img2 = (rng.standard_normal(shape2) + 1j * rng.standard_normal(shape2)).astype(np.complex64)
    psi = rng.uniform(0, 360, O)
    eul = synth.inverse_euler_f32(np.zeros(O), np.zeros(O), psi).astype(np.float32)
    img = (rng.standard_normal((n, xs)) + 1j * rng.standard_normal((n, xs))).astype(np.complex64)
    re, im = np.ascontiguousarray(img.real), np.ascontiguousarray(img.imag)
    tx = (rng.uniform(-3, 3, T) * 2 * np.pi / n).astype(np.float32); ty = (rng.uniform(-3, 3, T) * 2 * np.pi / n).astype(np.float32)
    w = rng.uniform(0, 1, (O, T)).astype(np.float32)
    minvs2 = rng.uniform(0.5, 2, (n, xs)).astype(np.float32); ctfs = rng.uniform(-1, 1, (n, xs)).astype(np.float32)
    outs = []
    for orc in (ref, port):
        proj = Projector(img2, r_max, pf)
        bp = Backprojector(shape2, r_max, pf)
        orc.K.backproject2d_sgd(C.byref(bp.struct), C.byref(proj.struct), xs, n, _fp(re), _fp(im), _fp(tx), _fp(ty), _fp(w), _fp(minvs2), _fp(ctfs),
                                T, 0.3, 2.5, _fp(eul), O)
        outs.append((bp.real.copy(), bp.imag.copy(), bp.weight.copy()))
    assert np.abs(outs[0][2]).max() > 0
    for a, b in zip(outs[0], outs[1]):
        assert np.abs(a - b).max() <= 2e-5 * np.abs(a).max()


def test_pre_shift_is_the_sampled_translation_of_skip_align():
    """rb_particles.pre_shift (the particle's own fractional offset applied to the transforms once) against the reference's way of
    doing --skip_align: that offset as the particle's ONE sampled translation (src/ml_optimiser.cpp:4196-4225,
    acc_ml_optimiser_impl.h:3752-3756), through the compiled reference kernels where they are built.  One particle per pool, since the
    sampled translation is per particle."""
    import copy
    from oracle.bindings import Oracle, Projector, Backprojector, have_reference
    from relion_b200.workload import make_skip_align_workload
    from relion_b200.estep import ParticlePool
    o = Oracle("reference" if have_reference() else "port")
    wl = make_skip_align_workload(n_particles=4, nr_classes=2, seed=81, snr=0.5, ori_size=24)
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    P = wl.pool.n_particles
    for p in range(P):
        sl = slice(p, p + 1)
        base = dict(Fimg=wl.pool.Fimg[sl], Fimg_nomask=wl.pool.Fimg_nomask[sl], Fctf=wl.pool.Fctf[sl], group_id=wl.pool.group_id[sl],
                    optics_group=wl.pool.optics_group[sl], highres_Xi2=wl.pool.highres_Xi2[sl], old_offset=np.zeros((1, 2)),
                    prior_offset=np.zeros((1, 2)), dir_off=np.array([0, 1], np.int32), dir_idx=np.array([p], np.int32), dir_prior=np.ones(1),
                    psi_off=np.array([0, 1], np.int32), psi_idx=np.array([p], np.int32), psi_prior=np.ones(1))
        got = []
        for mode in ("pre_shift", "sampled"):
            s = copy.copy(wl.sampling)
            pool = ParticlePool(**base)
            if mode == "pre_shift":
                pool.pre_shift = wl.pool.pre_shift[sl]
            else:
                s.trans_x = wl.pool.pre_shift[sl, 0].copy(); s.trans_y = wl.pool.pre_shift[sl, 1].copy()
                s.over_trans_x, s.over_trans_y = s.trans_x.copy(), s.trans_y.copy()
            bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
            st, out, _ = o.estep_pool(wl.model, s, refs, bps, pool, num_threads=1)
            assert st == 0
            got.append((out, bps))
        (a, ba), (b, bb) = got
        assert a.particles["best_class"][0] == b.particles["best_class"][0]
        for f in ("dLL_nolog", "min_diff2", "sum_weight", "wsum_sigma2_offset", "wsum_norm_correction", "wsum_XA", "wsum_AA"):
            np.testing.assert_allclose(a.particles[f], b.particles[f], rtol=2e-5, atol=1e-6, err_msg=f)
        np.testing.assert_allclose(a.wsum_sigma2_noise, b.wsum_sigma2_noise, rtol=1e-3, atol=1e-5 * np.abs(b.wsum_sigma2_noise).max())
        for x, y in zip(ba, bb):
            for nm in ("real", "imag", "weight"):
                u, v = getattr(x, nm), getattr(y, nm)
                assert np.abs(u - v).max() <= 2e-5 * max(np.abs(v).max(), 1e-30), nm


def test_ctypes_mirrors_have_the_layout_of_the_header(tmp_path):
    """Every struct of include/relion_b200.h that relion_b200/capi.py mirrors: same size, same field names in the same order at the
    same offsets (a C program compiled against the header prints them).  A field added on one side only would otherwise shift every
    pointer behind it silently."""
    import ctypes as C
    import subprocess
    from relion_b200 import capi
    names = ["rb_sampling", "rb_model", "rb_particles", "rb_raw_particles", "rb_posed_raw", "rb_particle_out", "rb_pool_out", "rb_weights_out"]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "relion_b200.h"', 'int main(void) {']
    for n in names:
        st = getattr(capi, n)
        src.append(f'printf("{n} size %zu\\n", sizeof({n}));')
        for f in st._fields_:
            src.append(f'printf("{n} {f[0]} %zu\\n", offsetof({n}, {f[0]}));')
    src += ['return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), str(c), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    want = {}
    for line in out:
        if line.strip():
            s, f, v = line.split()
            want[(s, f)] = int(v)
    hdr = open(os.path.join(ROOT, "include", "relion_b200.h")).read()
    for n in names:
        st = getattr(capi, n)
        assert C.sizeof(st) == want[(n, "size")], n
        for f in st._fields_:
            assert getattr(st, f[0]).offset == want[(n, f[0])], (n, f[0])
        # no member of the C struct is missing from the mirror: the sizes match and the last mirrored field ends the struct
        last = st._fields_[-1]
        assert getattr(st, last[0]).offset + C.sizeof(last[1]) + 8 > C.sizeof(st), n
        assert f"}} {n};" in hdr


def test_every_environment_switch_is_documented():
    """INTEGRATION.md section 5 lists every RB_* variable the library or its Python host reads."""
    import glob
    import re
    seen = set()
    for f in glob.glob(os.path.join(ROOT, "relion_b200", "csrc", "*.cu*")) + glob.glob(os.path.join(ROOT, "relion_b200", "*.py")):
        seen |= set(re.findall(r'"(RB_[A-Z0-9_]+)"', open(f).read()))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(v for v in seen if v not in doc)
    assert not missing, missing
