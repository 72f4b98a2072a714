"""Pins the numpy restatements that check the rows AROUND the hot path (SURVEY.md 8 f1-f3, and the posed back-projection of
a12) against the REFERENCE's own host code, compiled from /root/reference into oracle/_ref/librefrecon.so (Projector,
BackProjector, softMaskOutsideMap, FourierTransformer; oracle/ref_recon.cpp) and oracle/_ref/librefkernels.so (the ALTCPU
image-preparation helpers of src/acc/cpu/cpu_kernels/helper.cpp; oracle/ref_kernels.cpp).

  relion_b200.synth.reference_ft / reference_ft_2d   ==  Projector::computeFourierTransformMap
  oracle.reconstruct.reconstruct                     ==  BackProjector::reconstruct (skip_gridding branch)
  oracle.reconstruct.symmetrise                      ==  BackProjector::symmetrise (point groups)
  oracle.reconstruct.update_ssnr                     ==  BackProjector::updateSSNRarrays
  oracle.reconstruct.soft_mask_outside_map / fsc     ==  softMaskOutsideMap / getFSC
  oracle.bindings.backproject_posed                  ==  BackProjector::set2DFourierTransform (backproject2Dto3D)
  oracle.prepare.*                                   ==  cpu_translate2D, softMaskBackgroundValue + cosineFilter, powerClass,
                                                         CenterFFT + FourierTransformer + windowFourierTransform
The -m gpu tests compare the device code (rb_set_reference_from_map, rb_reconstruct, rb_bp_symmetrise, rb_update_ssnr,
rb_pool_prepare, rb_backproject_posed) with those restatements, so this file is what ties them to the reference.
Machines without the compiled libraries (no /root/reference) run the same restatements against tests/golden/host_golden.npz,
which tools/make_host_golden.py wrote from the compiled reference.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reconstruct as rc            # noqa: E402
from oracle import prepare as prep              # noqa: E402
from oracle import refhost                      # noqa: E402
from relion_b200 import synth                   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "host_golden.npz")
needs_ref = pytest.mark.skipif(not refhost.available(), reason="oracle/_ref/librefrecon.so not built (needs /root/reference)")


def _fwd(ainv):
    """Euler matrices A3D of the inverse matrices (rotations: the transpose); BackProjector inverts A itself."""
    return np.ascontiguousarray(np.transpose(ainv, (0, 2, 1)))


def _accumulators(ori=24, current=24, n_img=40, seed=3, pf=2.0):
    """Accumulators of a small posed back-projection (through the REFERENCE's BackProjector when it is there)."""
    rng = np.random.default_rng(seed)
    vol = synth.make_phantom(ori, n_blobs=12, seed=seed)
    data, r_max = synth.reference_ft(vol, current, pf)
    n = current
    rot, tilt, psi = rng.uniform(0, 360, n_img), np.degrees(np.arccos(rng.uniform(-1, 1, n_img))), rng.uniform(0, 360, n_img)
    eul = synth.inverse_euler_f32(rot, tilt, psi).reshape(n_img, 3, 3).astype(np.float64)
    F = np.stack([synth.project_numpy(data, r_max, pf, eul[i], n) for i in range(n_img)])
    W = rng.uniform(0.2, 1.0, F.shape)
    # pixels exactly ON the sphere |k| = r_max are in or out of backproject2Dto3D's `r2_3D > max_r2` test by the last bit of the
    # rotated coordinates (|A k|^2 == max_r2 in exact arithmetic): give them no weight, so that both sides skip them
    iy = np.arange(n); ky = np.where(iy < n // 2 + 1, iy, iy - n)[:, None]; kx = np.arange(n // 2 + 1)[None, :]
    W = W * ((kx * kx + ky * ky) < r_max * r_max)
    return vol, F * W, W, eul, r_max          # eul: the INVERSE Euler matrices (what the projection / the restatement take)


@needs_ref
@pytest.mark.parametrize("ori,cur", [(32, 32), (32, 24), (20, 12)])
def test_reference_ft_is_projector_compute_fourier_transform_map(ori, cur):
    vol = synth.make_phantom(ori, n_blobs=15, seed=ori + cur)
    want, (sz, sy), r_max, _ = refhost.ft_map(vol, cur, 2.0, data_dim=2)
    got, r = synth.reference_ft(vol, cur, 2.0)
    assert r == r_max and got.shape == want.shape and sz == sy == -((got.shape[0] - 1) // 2)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@needs_ref
def test_reference_ft_2d_is_projector_compute_fourier_transform_map():
    img = synth.make_phantom_2d(32, n_blobs=9, seed=4)
    want, (_, sy), r_max, _ = refhost.ft_map(img, 24, 2.0, data_dim=2)
    got, r = synth.reference_ft_2d(img, 24, 2.0)
    assert r == r_max and got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@needs_ref
def test_posed_backprojection_is_backprojector_set2dfouriertransform():
    from oracle.bindings import backproject_posed
    _, F, W, eul, r_max = _accumulators()
    want = refhost.backproject(F, _fwd(eul), W, 24, 24, 2.0)
    got = backproject_posed(want[0].shape, F, W, eul.reshape(-1, 9), r_max, 2.0)
    for a, b in zip(got, want):
        # the restatement takes fp32 images / matrices (it is the checker of the fp32 device kernel)
        assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()


@needs_ref
@pytest.mark.parametrize("with_tau2", [False, True])
def test_reconstruct_is_backprojector_reconstruct(with_tau2):
    ori = 24
    _, F, W, eul, r_max = _accumulators(ori, ori)
    re, im, w = refhost.backproject(F, _fwd(eul), W, ori, ori, 2.0)
    tau2 = np.linspace(4.0, 0.05, ori // 2 + 1) if with_tau2 else None
    want = refhost.reconstruct(re, im, w, ori, ori, 2.0, tau2=tau2, tau2_fudge=2.0, minres_map=2)
    got = rc.reconstruct(re, im, w, ori, r_max, 2.0, tau2=tau2, tau2_fudge=2.0, minres_map=2)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()


@needs_ref
@pytest.mark.parametrize("sym", ["C2", "D2", "C3"])
def test_symmetrise_is_backprojector_symmetrise(sym):
    ori = 20
    _, F, W, eul, r_max = _accumulators(ori, ori, n_img=25, seed=8)
    re, im, w = refhost.backproject(F, _fwd(eul), W, ori, ori, 2.0)
    R = refhost.sym_matrices(sym)
    assert len(R) >= 1
    want = refhost.symmetrise(re, im, w, ori, ori, sym, 2.0)
    got = rc.symmetrise(re, im, w, r_max, 2.0, R)
    for a, b in zip(got, want):
        assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()


@needs_ref
@pytest.mark.parametrize("sym,nr_asu,twist,rise", [("C1", 5, 22.03, 1.408), ("C1", 4, -166.7, 0.9), ("C2", 3, 30.0, 0.0), ("D2", 1, 10.0, 2.0)])
def test_symmetrise_with_helical_symmetry_is_backprojector_symmetrise(sym, nr_asu, twist, rise):
    """applyHelicalSymmetry between the Hermitian fold and the point group: odd and even numbers of asymmetrical units,
    no rise (no phase ramp), nr_asu = 1 (no helical part)."""
    ori = 20
    _, F, W, eul, r_max = _accumulators(ori, ori, n_img=25, seed=9)
    re, im, w = refhost.backproject(F, _fwd(eul), W, ori, ori, 2.0)
    R = refhost.sym_matrices(sym) if sym != "C1" else np.zeros((0, 3, 3))
    want = refhost.symmetrise_helical(re, im, w, ori, ori, sym, nr_asu, twist, rise, 2.0)
    got = rc.symmetrise(re, im, w, r_max, 2.0, R, helical=(nr_asu, twist, rise, ori))
    for a, b in zip(got, want):
        assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()
    if nr_asu > 1:
        plain = rc.symmetrise(re, im, w, r_max, 2.0, R)
        assert np.abs(got[2] - plain[2]).max() > 0.1 * np.abs(plain[2]).max()


@needs_ref
@pytest.mark.parametrize("with_fsc,whole", [(False, False), (True, False), (True, True)])
def test_update_ssnr_is_backprojector_update_ssnr_arrays(with_fsc, whole):
    ori = 24
    _, F, W, eul, r_max = _accumulators(ori, ori)
    _, _, w = refhost.backproject(F, _fwd(eul), W, ori, ori, 2.0)
    ns = ori // 2 + 1
    tau2 = np.linspace(3.0, 0.02, ns)
    fsc = np.clip(np.linspace(1.0, -0.05, ns), -1, 1)
    avg = np.linspace(1.0, 0.4, ns)
    want = refhost.update_ssnr(w, ori, ori, 2.0, 2.0, tau2, fsc=fsc, avgctf2=avg, update_tau2_with_fsc=with_fsc, is_whole_instead_of_half=whole)
    got = rc.update_ssnr(w, ori, r_max, 2.0, 2.0, tau2, fsc=fsc, avgctf2=avg, update_tau2_with_fsc=with_fsc, is_whole_instead_of_half=whole)
    for a, b, name in zip(got, want, ("tau2", "sigma2", "data_vs_prior", "fourier_coverage")):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12, err_msg=name)


@needs_ref
def test_soft_mask_and_fsc_are_the_references():
    rng = np.random.default_rng(5)
    v = rng.standard_normal((20, 20, 20))
    want = refhost.soft_mask(v, 10.0, 3.0)
    got = rc.soft_mask_outside_map(v, 3.0)
    assert np.abs(got - want).max() <= 1e-12
    a, b = synth.make_phantom(20, 10, seed=1), synth.make_phantom(20, 10, seed=1) + 0.2 * rng.standard_normal((20, 20, 20))
    np.testing.assert_allclose(rc.fsc(a, b), refhost.fsc(a, b), rtol=1e-10, atol=1e-12)


@needs_ref
@pytest.mark.parametrize("n,cs", [(32, 32), (32, 20), (30, 18)])
def test_image_transform_is_centerfft_fouriertransform_window(n, cs):
    rng = np.random.default_rng(n + cs)
    img = rng.standard_normal((n, n))
    want = refhost.image_ft(img, cs)
    got, _ = prep.normalize_and_transform(img, cs)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@needs_ref
@pytest.mark.parametrize("with_tau2", [False, True])
def test_reconstruct_with_iterative_gridding_is_backprojector_reconstruct(with_tau2):
    """--dont_skip_gridding: the restated Pipe & Menon iteration (oracle/reconstruct.py, max_iter_preweight = 10) against the
    reference's own BackProjector::reconstruct(skip_gridding = false), src/backprojector.cpp:1577-1700."""
    ori, cur = 24, 24
    vol, F, W, eul, r_max = _accumulators(ori, cur, n_img=60, seed=9)
    re, im, w = refhost.backproject(F, _fwd(eul), W, ori, cur, 2.0)
    tau2 = np.linspace(3.0, 0.05, ori // 2 + 1) if with_tau2 else None
    want = refhost.reconstruct(re, im, w, ori, cur, 2.0, tau2=tau2, tau2_fudge=1.5, minres_map=1, skip_gridding=False, max_iter_preweight=10)
    got = rc.reconstruct(re, im, w, ori, r_max, 2.0, tau2=tau2, tau2_fudge=1.5, minres_map=1, max_iter_preweight=10)
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    skip = rc.reconstruct(re, im, w, ori, r_max, 2.0, tau2=tau2, tau2_fudge=1.5, minres_map=1)
    assert np.abs(got - skip).max() > 1e-3 * np.abs(want).max()          # the two branches really differ


@needs_ref
@pytest.mark.parametrize("n,shift", [(32, (0.0, 0.0)), (32, (2.3, -1.7)), (24, (-0.4, 3.1))])
def test_posed_particle_preparation_is_the_reference_transform_and_shift(n, shift):
    """oracle.backproject_posed.prepare_particle (the checker of rb_backproject_posed_raw) against the reference's own
    CenterFFT + FourierTransformer + shiftImageInFourierTransform (src/fftw.cpp:874-918): centring by sign in Fourier space equals
    the real-space CenterFFT for even boxes, the shift has the reference's sign; only the DC term differs (set to zero)."""
    from oracle.backproject_posed import prepare_particle
    rng = np.random.default_rng(n)
    img = rng.standard_normal((n, n))
    want = refhost.image_ft(img, n, shift)
    got, w = prepare_particle(img, shift, None)
    assert got[0, 0] == 0 and np.all(w == 1.0)
    want[0, 0] = 0
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@needs_ref
def test_image_preparation_helpers_are_the_altcpu_kernels():
    rng = np.random.default_rng(11)
    n, cs = 32, 20
    img = rng.standard_normal((n, n)).astype(np.float32)
    for dx, dy in ((0, 0), (3, -2), (-5, 4)):
        want = refhost.prep_translate2d(img, dx, dy)
        got = prep.translate_and_norm(img.astype(np.float64), dx, dy, 1.0)
        assert np.array_equal(got.astype(np.float32), want)
    for radius, width in ((11.0, 3.0), (-1.0, 5.0), (9.5, 2.0)):
        want, bg = refhost.prep_soft_mask(img, radius, width)
        got, gbg = prep.soft_mask(img.astype(np.float64), radius, width)
        assert abs(gbg - bg) <= 2e-5 * max(1.0, abs(bg))
        assert np.abs(got - want).max() <= 2e-5
    F = (rng.standard_normal((n, n // 2 + 1)) + 1j * rng.standard_normal((n, n // 2 + 1))).astype(np.complex64)
    ws, wxi = refhost.prep_power_class(F, cs)
    gs, gxi = prep.power_class(F.astype(np.complex128), cs)
    np.testing.assert_allclose(gs, ws, rtol=2e-5)
    assert abs(gxi - wxi) <= 2e-5 * wxi


def test_restatements_match_the_committed_reference_outputs():
    """The same restatements against outputs of the compiled reference stored by tools/make_host_golden.py."""
    g = np.load(GOLDEN)
    vol = g["vol"]
    got, _ = synth.reference_ft(vol, int(g["ft_current"]), 2.0)
    assert np.abs(got - g["ft_data"]).max() <= 1e-12 * np.abs(g["ft_data"]).max()
    ori = vol.shape[0]
    r_max = int(g["r_max"])
    tau2 = g["tau2"]
    got = rc.reconstruct(g["bp_re"], g["bp_im"], g["bp_w"], ori, r_max, 2.0, tau2=tau2, tau2_fudge=2.0, minres_map=2)
    assert np.abs(got - g["recon"]).max() <= 1e-10 * np.abs(g["recon"]).max()
    got = rc.symmetrise(g["bp_re"], g["bp_im"], g["bp_w"], r_max, 2.0, g["sym_R"])
    for a, name in zip(got, ("sym_re", "sym_im", "sym_w")):
        assert np.abs(a - g[name]).max() <= 1e-10 * np.abs(g[name]).max()
    got = rc.update_ssnr(g["bp_w"], ori, r_max, 2.0, 2.0, tau2, fsc=g["fsc"], update_tau2_with_fsc=True)
    for a, name in zip(got, ("ssnr_tau2", "ssnr_sigma2", "ssnr_dvp", "ssnr_cov")):
        np.testing.assert_allclose(a, g[name], rtol=1e-10, atol=1e-12, err_msg=name)
    from oracle.bindings import backproject_posed
    got = backproject_posed(g["bp_re"].shape, g["img_F"], g["img_W"], g["img_A"].reshape(-1, 9), r_max, 2.0)
    for a, name in zip(got, ("bp_re", "bp_im", "bp_w")):
        assert np.abs(a - g[name]).max() <= 2e-6 * np.abs(g[name]).max()
    got, _ = prep.normalize_and_transform(g["raw_img"].astype(np.float64), int(g["raw_cs"]))
    assert np.abs(got - g["raw_ft"]).max() <= 1e-12 * np.abs(g["raw_ft"]).max()
    got, _ = prep.soft_mask(g["raw_img"].astype(np.float64), 11.0, 3.0)
    assert np.abs(got - g["raw_masked"]).max() <= 2e-5


@needs_ref
def test_compiled_helper_cpp_matches_the_restated_helpers():
    """exponentiate_weights_fine and cpu_kernel_make_eulers_3D: the `reference` kernel table now holds the functions compiled
    from src/acc/cpu/cpu_kernels/helper.cpp itself; the restatements of oracle/port_kernels.cpp must reproduce them."""
    from oracle.bindings import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/librefkernels.so not built")
    rng = np.random.default_rng(2)
    rot, tilt, psi = rng.uniform(-180, 180, 200), rng.uniform(0, 180, 200), rng.uniform(-180, 180, 200)
    a, b = Oracle("port").make_eulers(rot, tilt, psi), Oracle("reference").make_eulers(rot, tilt, psi)
    assert np.abs(a - b).max() <= 2e-7
