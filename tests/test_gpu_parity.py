"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Tolerances (fp32 kernels, sums of O(1e3) terms, different association order and FMA contraction):
  projection values               1e-5 of the slice maximum
  diff2 / wavg sums               2e-5 relative
  weights after expf              2e-4 relative (CUDA expf vs glibc expf, <= 2 ulp, amplified by exp(50-max))
  back-projected volumes          1e-4 of the volume maximum (fp32 atomics, order dependent)
  significance selection          bit-exact against the oracle's exact-arithmetic rule on the same weights
  log-likelihood                  1e-4 relative (north_star)
  max-posterior pose              >= 99.5 % of particles (north_star)
"""
import numpy as np
import pytest

from relion_b200 import synth
from relion_b200.workload import make_workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    from oracle.bindings import Oracle
    return Oracle("port")


def _setup(device, wl, n_acc=None):
    device.set_model(wl.model)
    device.set_sampling(wl.sampling)
    for k, v in enumerate(wl.refs):
        device.set_reference(k, v, wl.r_max, wl.padding_factor)
    for k in range(n_acc or len(wl.refs)):
        device.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)


def _stage_inputs(wl, n, seed, n_orient, n_trans):
    rng = np.random.default_rng(seed)
    xs = n // 2 + 1
    eul = synth.inverse_euler_f32(rng.uniform(-180, 180, n_orient), rng.uniform(0, 180, n_orient), rng.uniform(0, 360, n_orient))
    tx = (-2 * np.pi * rng.uniform(-4, 4, n_trans) / wl.model.ori_size).astype(np.float32)
    ty = (-2 * np.pi * rng.uniform(-4, 4, n_trans) / wl.model.ori_size).astype(np.float32)
    re = rng.standard_normal((n, xs)).astype(np.float32)
    im = rng.standard_normal((n, xs)).astype(np.float32)
    corr = rng.uniform(0.5, 2.0, (n, xs)).astype(np.float32)
    corr[0, 0] = 0
    return eul, tx, ty, re, im, corr


@pytest.mark.parametrize("n,r_max_cut", [(32, False), (20, False), (32, True)])
def test_project(device, oracle, n, r_max_cut):
    from oracle.bindings import Projector
    wl = make_workload(ori_size=32, n_particles=2, seed=11)
    r_max = 10 if r_max_cut else wl.r_max
    device.set_reference(0, wl.refs[0], r_max, wl.padding_factor)
    ref = Projector(wl.refs[0], r_max, wl.padding_factor)
    eul = _stage_inputs(wl, n, 1, 9, 1)[0]
    got = device.project(0, n, eul)
    for i in range(len(eul)):
        want = oracle.project(ref, n, eul[i])
        assert np.abs(got[i] - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("flags", [dict(), dict(do_scale_correction=False), dict(do_ctf_correction=False), dict(refs_are_ctf_corrected=False)])
def test_image_corrections_stage(device, flags):
    """pixel_correction and corr_img (acc_ml_optimiser_impl.h:1251-1268, buildCorrImage
    acc_helper_functions_impl.h:164-196) with Minvsigma2 = 1 / (sigma2_fudge sigma2_noise[ires]) on the Mresol pixels, DC
    excluded (src/ml_optimiser.cpp:6868-6879), at the coarse window, restated in numpy from those lines."""
    wl = make_workload(ori_size=40, current_size=32, healpix_order=1, n_particles=6, seed=14, snr=0.3, nr_groups=3)
    for k, v in flags.items():
        setattr(wl.model, k, v)
    wl.model.scale_correction = np.array([0.8, 1.0, 1.3])
    _setup(device, wl)
    device.pool_upload(0, wl.pool)
    device.estep_slot(0)
    m = wl.model
    nc, cs = m.coarse_size, m.current_size
    xs = nc // 2 + 1
    iy = np.arange(nc); yy = np.where(iy < xs, iy, iy - nc)[:, None]; xx = np.arange(xs)[None, :]
    ires = np.rint(np.sqrt((xx * xx + yy * yy).astype(np.float64))).astype(int)
    mresol = (ires < xs) & ~((xx == 0) & (yy < 0))
    for p in range(wl.pool.n_particles):
        F = synth.window_ft(np.asarray(wl.pool.Fimg[p]), nc)
        ctf = synth.window_ft(np.asarray(wl.pool.Fctf[p]), nc).real if m.do_ctf_correction else np.ones((nc, xs))
        s2 = np.atleast_2d(m.sigma2_noise)[wl.pool.optics_group[p]]
        minv = np.where(mresol & (ires > 0), 1.0 / (m.sigma2_fudge * s2[np.minimum(ires, len(s2) - 1)]), 0.0)
        scale = m.scale_correction[wl.pool.group_id[p]] if m.do_scale_correction else 1.0
        pc = np.full((nc, xs), 1.0 / scale)
        corr = minv.copy()
        if m.do_ctf_correction and m.refs_are_ctf_corrected:
            pc = np.where(np.abs(ctf) > 1e-8, pc / np.where(ctf == 0, 1, ctf), pc)
            corr = corr * ctf * ctf
        if m.do_scale_correction:
            corr = corr * scale * scale
        got = device.debug_prepared_coarse_image(0, p, nc)
        sel = mresol
        want = F * pc
        assert np.abs(got[..., 0][sel] - want.real[sel]).max() <= 2e-6 * np.abs(want[sel]).max()
        assert np.abs(got[..., 1][sel] - want.imag[sel]).max() <= 2e-6 * np.abs(want[sel]).max()
        np.testing.assert_allclose(got[..., 2][sel], 0.5 * corr[sel], rtol=2e-6, atol=1e-30)
        assert not got[..., 2][~sel].any()


def test_coarse_euler_matrices_stage(device):
    """the coarse-pass Euler matrices the device builds in rb_set_sampling against the reference's own
    cpu_kernel_make_eulers_3D<invert = true> (src/acc/cpu/cpu_kernels/helper.cpp, compiled in oracle/_ref; the restated port
    where the compiled reference is absent): fp32, [n_dir][n_psi][9]."""
    from oracle.bindings import Oracle
    try:
        orc = Oracle("reference")
    except Exception:
        orc = Oracle("port")
    wl = make_workload(ori_size=32, healpix_order=2, n_particles=2, seed=5, snr=0.3)
    _setup(device, wl)
    s = wl.sampling
    got = device.debug_coarse_eulers(s.n_dir, s.n_psi)
    rot = np.repeat(np.asarray(s.rot, np.float32), s.n_psi); tilt = np.repeat(np.asarray(s.tilt, np.float32), s.n_psi)
    psi = np.tile(np.asarray(s.psi, np.float32), s.n_dir)
    want = orc.make_eulers(rot, tilt, psi).reshape(s.n_dir, s.n_psi, 9)
    assert np.abs(got - want).max() <= 2e-6


@pytest.mark.parametrize("n,O,T,r_max_cut", [(14, 37, 9, False), (32, 5, 21, False), (32, 256 + 3, 30, False), (24, 8, 5, True), (16, 1, 1, False)])
def test_diff2_coarse_stage(device, oracle, n, O, T, r_max_cut):
    from oracle.bindings import Projector
    wl = make_workload(ori_size=32, n_particles=2, seed=12)
    r_max = 7 if r_max_cut else wl.r_max
    device.set_reference(0, wl.refs[0], r_max, wl.padding_factor)
    ref = Projector(wl.refs[0], r_max, wl.padding_factor)
    eul, tx, ty, re, im, corr = _stage_inputs(wl, n, 2, O, T)
    init = np.full((O, T), 3.25, np.float32)
    got = device.diff2_coarse(0, n, eul, tx, ty, re, im, corr, init=init)
    want = oracle.diff2_coarse(ref, n, eul, tx, ty, re, im, corr, init=init)
    np.testing.assert_allclose(got, want, rtol=2e-5)


@pytest.mark.parametrize("n,r_max_cut", [(32, False), (18, False), (32, True)])
def test_diff2_fine_stage(device, oracle, n, r_max_cut):
    from oracle.bindings import Projector
    wl = make_workload(ori_size=32, n_particles=2, seed=13)
    r_max = 9 if r_max_cut else wl.r_max
    device.set_reference(0, wl.refs[0], r_max, wl.padding_factor)
    ref = Projector(wl.refs[0], r_max, wl.padding_factor)
    O, T = 11, 36
    eul, tx, ty, re, im, corr = _stage_inputs(wl, n, 3, O, T)
    rng = np.random.default_rng(5)
    # jobs: runs of <= 4 consecutive translations of one orientation (makeJobsForDiff2Fine)
    rot_idx, trans_idx, job_idx, job_num = [], [], [], []
    for o in range(O):
        t = 0
        while t < T:
            if rng.random() < 0.5:
                t += 1
                continue
            ln = int(min(rng.integers(1, 5), T - t))
            job_idx.append(len(rot_idx)); job_num.append(ln)
            for j in range(ln):
                rot_idx.append(o); trans_idx.append(t + j)
            t += ln + 1
    got = device.diff2_fine(0, n, eul, tx, ty, re, im, corr, 1.5, rot_idx, trans_idx, job_idx, job_num)
    want = oracle.diff2_fine(ref, n, eul, tx, ty, re, im, corr, 1.5, rot_idx, trans_idx, job_idx, job_num)
    np.testing.assert_allclose(got, want, rtol=2e-5)


@pytest.mark.parametrize("seed", range(12))
def test_convert_weights(device, oracle, seed):
    rng = np.random.default_rng(100 + seed)
    no, nt = int(rng.integers(1, 400)), int(rng.integers(1, 30))
    spread = [0.5, 3.0, 30.0, 300.0][seed % 4]
    d2 = (1000.0 + spread * rng.random((no, nt))).astype(np.float32)
    if seed % 3 == 0:
        d2[rng.random((no, nt)) < 0.2] = np.finfo(np.float32).min        # orientations never computed
        d2[0, 0] = 1000.0
    pdf = rng.random(no)
    pdf[rng.random(no) < 0.1] = 0.0
    pdf[0] = 0.5
    po = np.where(pdf > 0, np.log(np.where(pdf > 0, pdf, 1.0)), 0.0).astype(np.float32)
    oz = (pdf == 0).astype(np.uint8)
    pt = (-rng.random(nt)).astype(np.float32)
    tz = np.zeros(nt, np.uint8)
    maxsig = 5 if seed % 5 == 4 else 0
    w, sig, out = device.convert_weights(d2, po, oz, pt, tz, 0.999, maxsig)
    ref = oracle.convert_weights_coarse(d2, po, oz, pt, tz, 0.999, maxsig, exact=True)
    np.testing.assert_allclose(w, ref["weights"], rtol=2e-4, atol=0)
    assert out.min_diff2 == np.float32(ref["min_diff2"])
    # the selector itself, on the GPU's own weights: bit-exact
    own = oracle.significance(w, 0.999, maxsig, True, exact=True)
    assert out.n_nonzero == own["n_filtered"]
    assert out.nr_significant == own["n_filtered"] - own["threshold_idx"]
    assert np.float32(out.significant_weight) == np.float32(own["significant_weight"])
    assert np.float32(out.sum_weight) == np.float32(own["sum_weight"])
    assert np.array_equal(sig.astype(bool), w >= np.float32(own["significant_weight"]))
    assert int(out.max_index) == int(np.argmax(w))


def test_convert_weights_ties_and_edges(device, oracle):
    # many identical weights: the threshold falls inside a run of ties
    d2 = np.full((50, 4), 1000.0, np.float32)
    po = np.zeros(50, np.float32); oz = np.zeros(50, np.uint8); pt = np.zeros(4, np.float32); tz = np.zeros(4, np.uint8)
    for frac in (0.999, 0.5, 0.9, 1.0, 0.0):
        w, sig, out = device.convert_weights(d2, po, oz, pt, tz, frac, 0)
        own = oracle.significance(w, frac, 0, True, exact=True)
        assert out.nr_significant == own["n_filtered"] - own["threshold_idx"], frac
        assert np.float32(out.significant_weight) == np.float32(own["significant_weight"])
    # a single sample
    w, sig, out = device.convert_weights(np.array([[7.0]], np.float32), po[:1], oz[:1], pt[:1], tz[:1], 0.999, 0)
    assert sig[0, 0] == 1


def test_wavg_and_backproject_stage(device, oracle):
    from oracle.bindings import Projector, Backprojector
    wl = make_workload(ori_size=32, n_particles=2, seed=14)
    n, O, T = 32, 6, 8
    device.set_model(wl.model)
    device.set_reference(0, wl.refs[0], wl.r_max, wl.padding_factor)
    device.bp_init(0, wl.bp_shape, wl.r_max, wl.padding_factor)
    ref = Projector(wl.refs[0], wl.r_max, wl.padding_factor)
    bp = Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor)
    eul, tx, ty, re, im, corr = _stage_inputs(wl, n, 4, O, T)
    rng = np.random.default_rng(6)
    weights = rng.random((O, T)).astype(np.float32)
    weights[rng.random((O, T)) < 0.5] = np.finfo(np.float32).min
    ctfs = rng.uniform(-1, 1, (n, n // 2 + 1)).astype(np.float32)
    minvs2 = corr
    got = device.wavg(0, n, eul, tx, ty, re, im, weights, ctfs, 3.7, 0.2)
    want = oracle.wavg(ref, n, eul, tx, ty, re, im, weights, ctfs, 3.7, 0.2)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=2e-5, atol=1e-6 * np.abs(w).max())
    device.backproject(0, n, eul, tx, ty, re, im, weights, minvs2, ctfs, 3.7, 0.2)
    oracle.backproject(bp, n, eul, tx, ty, re, im, weights, minvs2, ctfs, 3.7, 0.2)
    gre, gim, gw = device.bp_get(0)
    for g, w in ((gre, bp.real), (gim, bp.imag), (gw, bp.weight)):
        assert np.abs(w).max() > 0
        assert np.abs(g - w).max() <= 1e-4 * np.abs(w).max()


@pytest.mark.parametrize("n,O,T,r_max_cut", [(14, 37, 9, False), (32, 5, 21, False), (32, 256 + 3, 30, False), (24, 8, 5, True)])
def test_diff2_cc_coarse_stage(device, oracle, n, O, T, r_max_cut):
    """First-iteration cross-correlation criterion, coarse kernel (diff2_CC_coarse, cpu_kernels/diff2.h:611-742)."""
    from oracle.bindings import Projector
    wl = make_workload(ori_size=32, n_particles=2, seed=12)
    r_max = 7 if r_max_cut else wl.r_max
    device.set_reference(0, wl.refs[0], r_max, wl.padding_factor)
    ref = Projector(wl.refs[0], r_max, wl.padding_factor)
    eul, tx, ty, re, im, corr = _stage_inputs(wl, n, 2, O, T)
    corr[0, 0] = 0.7          # the CC weight does not vanish at the origin
    init = np.full((O, T), 0.25, np.float32)
    got = device.diff2_coarse(0, n, eul, tx, ty, re, im, corr, init=init, cc=True)
    want = oracle.diff2_coarse(ref, n, eul, tx, ty, re, im, corr, init=init, cc=True)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want - 0.25).max()


@pytest.mark.parametrize("n,r_max_cut", [(32, False), (18, False), (32, True)])
def test_diff2_cc_fine_stage(device, oracle, n, r_max_cut):
    """First-iteration cross-correlation criterion, fine kernel (diff2_CC_fine, cpu_kernels/diff2.h:904-1050)."""
    from oracle.bindings import Projector
    wl = make_workload(ori_size=32, n_particles=2, seed=13)
    r_max = 9 if r_max_cut else wl.r_max
    device.set_reference(0, wl.refs[0], r_max, wl.padding_factor)
    ref = Projector(wl.refs[0], r_max, wl.padding_factor)
    O, T = 11, 36
    eul, tx, ty, re, im, corr = _stage_inputs(wl, n, 3, O, T)
    corr[0, 0] = 1.3
    rng = np.random.default_rng(5)
    rot_idx, trans_idx, job_idx, job_num = [], [], [], []
    for o in range(O):
        t = 0
        while t < T:
            if rng.random() < 0.5:
                t += 1
                continue
            ln = int(min(rng.integers(1, 5), T - t))
            job_idx.append(len(rot_idx)); job_num.append(ln)
            for j in range(ln):
                rot_idx.append(o); trans_idx.append(t + j)
            t += ln + 1
    got = device.diff2_cc_fine(0, n, eul, tx, ty, re, im, corr, rot_idx, trans_idx, job_idx, job_num)
    want = oracle.diff2_cc_fine(ref, n, eul, tx, ty, re, im, corr, rot_idx, trans_idx, job_idx, job_num)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def _compare_pool(device, oracle, wl, pose_frac=0.995, exact_threshold=True, num_threads=0, n_acc=None):
    from oracle.bindings import Projector, Backprojector
    from oracle.parity import classify_significance
    n_acc = n_acc or len(wl.refs)
    _setup(device, wl, n_acc)
    res = device.expectation_some_particles(wl.pool)
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in range(n_acc)]
    st, ores, _ = oracle.estep_pool(wl.model, wl.sampling, refs, bps, wl.pool, num_threads=num_threads, exact_threshold=exact_threshold)
    assert st == 0
    g, o = res.particles, ores.particles
    agree = np.mean(g["best_ihidden_over"] == o["best_ihidden_over"])
    assert agree >= pose_frac, agree
    np.testing.assert_allclose(g["min_diff2_coarse"], o["min_diff2_coarse"], rtol=2e-5)
    # Significant-pose counts (north_star: "agree exactly whenever the reference's diff2 are identical").  The GPU's own
    # coarse weights go through the REFERENCE's rule (fp32 sequential scan, acc_helper_functions.h:226-232): the count must
    # be the same, or the cumulative sum must sit on the threshold within the rounding of an fp32 scan (oracle/parity.py).
    if not wl.model.do_cc:
        sig = classify_significance(device, 0, g, oracle, wl.model.adaptive_fraction, wl.model.maximum_significants)
        assert sig["mismatch"] == 0, sig
        assert sig["rounding_edge"] <= max(1, 0.1 * len(g)), sig
    # against the oracle's own run the weights differ in the last bits (expf, summation order), so a count may move by one
    # where its threshold is such an edge: those particles are masked out of the weighted-sum comparisons below
    same = g["nr_significant_coarse"] == o["nr_significant_coarse"]
    assert same.mean() >= 0.9, (g["nr_significant_coarse"], o["nr_significant_coarse"])
    ok = same & (g["n_fine_samples"] == o["n_fine_samples"]) & (g["best_ihidden_over"] == o["best_ihidden_over"])
    np.testing.assert_allclose(g["dLL_nolog"], o["dLL_nolog"], rtol=1e-4)
    np.testing.assert_allclose(g["pmax"][ok], o["pmax"][ok], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(g["sum_weight"][ok], o["sum_weight"][ok], rtol=2e-3)
    np.testing.assert_allclose(g["sumw"][ok], o["sumw"][ok], rtol=1e-4)
    assert np.mean(g["n_bp_orient"][ok] == o["n_bp_orient"][ok]) >= 0.9, (g["n_bp_orient"], o["n_bp_orient"])
    np.testing.assert_allclose(g["wsum_sigma2_offset"][ok], o["wsum_sigma2_offset"][ok], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(g["wsum_norm_correction"][ok], o["wsum_norm_correction"][ok], rtol=2e-3)
    np.testing.assert_allclose(g["wsum_XA"][ok], o["wsum_XA"][ok], rtol=5e-3, atol=1e-3 * np.abs(o["wsum_XA"]).max())
    np.testing.assert_allclose(g["wsum_AA"][ok], o["wsum_AA"][ok], rtol=5e-3)
    sh = np.abs(ores.wsum_sigma2_noise).max()
    assert np.abs(res.wsum_sigma2_noise[ok] - ores.wsum_sigma2_noise[ok]).max() <= 5e-3 * sh
    if wl.model.prior_offset_class is not None:
        # thr_wsum_prior_offsetx/y_class (:2847-2851): sums of weight x offset; the tolerance is relative to the largest sum
        scale = max(np.abs(ores.wsum_prior_offset_class).max(), wl.model.pixel_size * 1e-3 * len(g))
        assert np.abs(res.wsum_prior_offset_class - ores.wsum_prior_offset_class).max() <= 2e-3 * scale + (0 if ok.all() else scale)
    if ok.all():
        np.testing.assert_allclose(res.wsum_pdf_class, ores.wsum_pdf_class, rtol=1e-4)
        assert np.abs(res.wsum_pdf_direction - ores.wsum_pdf_direction).max() <= 2e-3
        for k in range(n_acc):
            gre, gim, gw = device.bp_get(k)
            for a, b in ((gre, bps[k].real), (gim, bps[k].imag), (gw, bps[k].weight)):
                assert np.abs(a - b).max() <= 5e-3 * max(np.abs(b).max(), 1e-12)
    return res, ores


@pytest.mark.parametrize("stack", ["1", "0"])
def test_pool_global_search(device, oracle, monkeypatch, stack):
    """Two classes through the tensor-core contraction: class rows stacked along M (default) or one contraction per class."""
    monkeypatch.setenv("RB_GEMM_STACK", stack)
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=12, nr_classes=2, seed=21, snr=0.3)
    _compare_pool(device, oracle, wl)


@pytest.mark.parametrize("fused", ["1", "0"])
def test_pool_local_search(device, oracle, monkeypatch, fused):
    """Local searches; fused=1: projection + tcgen05 contraction per (particle, orientation tile) (k_coarse_fused),
    fused=0: the SIMT coarse kernel."""
    monkeypatch.setenv("RB_COARSE_FUSED", fused)
    wl = make_workload(ori_size=32, healpix_order=2, n_particles=10, nr_classes=1, seed=22, snr=0.2, local_search=True)
    _compare_pool(device, oracle, wl)


def test_pool_local_search_fused_two_classes_many_orientations(device, oracle, monkeypatch):
    """More than one 128-orientation tile per particle and two classes through the fused kernel; wide prior."""
    monkeypatch.setenv("RB_COARSE_FUSED", "1")
    wl = make_workload(ori_size=32, healpix_order=2, n_particles=6, nr_classes=2, seed=28, snr=0.2, local_search=True, sigma_ang=25.0)
    assert (np.diff(wl.pool.dir_off) * np.diff(wl.pool.psi_off)).max() > 128
    _compare_pool(device, oracle, wl)


def test_pool_global_search_through_fused_kernel(device, oracle, monkeypatch):
    """The fused kernel also handles identity orientation lists (global search) when forced."""
    monkeypatch.setenv("RB_COARSE_GEMM", "0")
    monkeypatch.setenv("RB_COARSE_FUSED", "2")
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=5, nr_classes=2, seed=29, snr=0.3)
    _compare_pool(device, oracle, wl)


@pytest.mark.parametrize("kw", [dict(ori_size=32, healpix_order=1, n_particles=12, nr_classes=1, seed=41, snr=0.3),
                                dict(ori_size=40, current_size=28, healpix_order=1, n_particles=8, nr_classes=1, seed=42, snr=0.1),
                                dict(ori_size=32, n_particles=16, nr_classes=1, seed=43, snr=0.5, ref_dim=2, psi_step=12.0)])
@pytest.mark.parametrize("coarse", ["gemm", "simt"])
def test_pool_firstiter_cc(device, oracle, monkeypatch, kw, coarse):
    """--firstiter_cc / --always_cc: cross-correlation kernels in both passes, weight one for the best pose
    (acc_ml_optimiser_impl.h:1164, 1287, 2012-2071, 3571); coarse pass through the tensor-core contraction and the SIMT kernel."""
    monkeypatch.setenv("RB_COARSE_GEMM", "2" if coarse == "gemm" else "0")
    wl = make_workload(do_cc=True, **kw)
    if kw.get("ref_dim") == 2:
        wl.model.bp_circle_bound = False
    res, ores = _compare_pool(device, oracle, wl)
    g = res.particles
    assert np.all(g["nr_significant_coarse"] == 1)
    assert np.all(g["n_fine_samples"] == wl.sampling.n_over_rot * wl.sampling.n_over_trans)
    assert np.all(g["sum_weight"] == 1.0) and np.all(g["pmax"] == 1.0)
    assert np.all(g["dLL_nolog"] > 0)           # = -min_diff2: the best normalised cross-correlation is positive


@pytest.mark.parametrize("fused", ["1", "0"])
def test_pool_always_cc_local_search(device, oracle, monkeypatch, fused):
    """--always_cc late in a refinement: the cross-correlation criterion with LOCAL searches, through the fused projection +
    tcgen05 kernel (CC epilogue: -cross / sqrt(norm)) and through the SIMT coarse kernel."""
    monkeypatch.setenv("RB_COARSE_FUSED", fused)
    wl = make_workload(do_cc=True, ori_size=32, healpix_order=2, n_particles=10, nr_classes=1, seed=47, snr=0.3, local_search=True)
    res, _ = _compare_pool(device, oracle, wl)
    assert np.all(res.particles["nr_significant_coarse"] == 1)
    assert np.all(res.particles["sum_weight"] == 1.0)


@pytest.mark.parametrize("band", ["1", "0"])
@pytest.mark.parametrize("local", [False, True])
def test_pool_gradient_refinement_backprojection(device, oracle, monkeypatch, band, local):
    """do_grad (SGD / VDAM): the back-projection accumulates the weighted residual sum_t w_t (X_t - CTF A)
    (cuda_kernel_backproject3D_SGD, BP.cuh:406-656) and, with the pseudo half-sets that gradient refinement switches on,
    particle p goes into accumulator class + (p % 2) * K (acc_ml_optimiser_impl.h:3395-3400): 2 K accumulators, through the
    band-major store kernel and the orientation-major one."""
    monkeypatch.setenv("RB_BAND", band)
    wl = make_workload(ori_size=32, healpix_order=2 if local else 1, n_particles=10, nr_classes=2, seed=61, snr=0.3, local_search=local)
    wl.model.do_grad = True
    wl.pool.bp_offset = (np.arange(wl.pool.n_particles) % 2 * wl.model.nr_classes).astype(np.int32)
    res, ores = _compare_pool(device, oracle, wl, n_acc=2 * wl.model.nr_classes)
    # both halves received something, and the residual accumulators differ from the weighted-image ones
    for k in range(2 * wl.model.nr_classes):
        assert np.abs(device.bp_get(k)[2]).max() > 0


def test_pool_gradient_refinement_2d_classification(device, oracle):
    """RELION 4's default 2D classification is a gradient (VDAM) refinement: 2D references, residual back-projection
    (cuda_kernel_backproject2D_SGD, BP.cuh:659-826), pseudo half-sets -> 2 K 2D accumulators."""
    wl = make_workload(ori_size=32, n_particles=24, nr_classes=3, seed=63, snr=0.2, ref_dim=2, psi_step=10.0)
    wl.model.do_grad = True
    wl.model.bp_circle_bound = False
    wl.pool.bp_offset = (np.arange(wl.pool.n_particles) % 2 * wl.model.nr_classes).astype(np.int32)
    _compare_pool(device, oracle, wl, n_acc=2 * wl.model.nr_classes)
    for k in range(2 * wl.model.nr_classes):
        assert np.abs(device.bp_get(k)[2]).max() > 0


@pytest.mark.parametrize("flags", [dict(do_map=False), dict(do_scale_correction=False), dict(do_ctf_correction=False),
                                   dict(do_map=False, do_scale_correction=False, do_ctf_correction=False)])
def test_pool_optimiser_flags(device, oracle, flags):
    """--no_map (all-ones Minvsigma2 in the back-projection, :3110-3115), no --scale, no --ctf."""
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=8, nr_classes=1, seed=61, snr=0.3)
    for k, v in flags.items():
        setattr(wl.model, k, v)
    _compare_pool(device, oracle, wl)


@pytest.mark.parametrize("coarse", ["gemm", "simt"])
def test_pool_global_search_with_zero_prior_directions(device, oracle, monkeypatch, coarse):
    """Directions with pdf_direction == 0 are never computed and stay at lowest() in Mweight (:3849, helper.cuh:39-42)."""
    monkeypatch.setenv("RB_COARSE_GEMM", "2" if coarse == "gemm" else "0")
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=8, nr_classes=2, seed=81, snr=0.3)
    pd = np.array(wl.model.pdf_direction, dtype=np.float64, copy=True)
    pd[0, 1::3] = 0.0
    pd[1, ::4] = 0.0
    wl.model.pdf_direction = pd
    _compare_pool(device, oracle, wl)


def test_pool_reduced_current_size(device, oracle):
    wl = make_workload(ori_size=40, current_size=28, healpix_order=1, n_particles=6, seed=23, snr=0.3)
    _compare_pool(device, oracle, wl)


def test_pool_low_snr_many_significant(device, oracle):
    wl = make_workload(ori_size=24, healpix_order=1, n_particles=6, seed=24, snr=0.01, adaptive_fraction=0.999)
    res, _ = _compare_pool(device, oracle, wl, pose_frac=0.8)
    assert res.particles["nr_significant_coarse"].max() > 5


def test_pool_single_particle_and_skip_maximization(device, oracle):
    wl = make_workload(ori_size=24, healpix_order=1, n_particles=1, seed=25)
    _setup(device, wl)
    r1 = device.expectation_some_particles(wl.pool, skip_maximization=True)
    assert r1.particles["wsum_norm_correction"][0] == 0.0
    assert np.all(device.bp_get(0)[2] == 0)
    r2 = device.expectation_some_particles(wl.pool)
    assert r2.particles["best_ihidden_over"][0] == r1.particles["best_ihidden_over"][0]
    assert r2.particles["wsum_norm_correction"][0] > 0


def test_pool_capacity_error(device, oracle, monkeypatch):
    from relion_b200.capi import RelionB200Error, RB_ERR_CAPACITY
    wl = make_workload(ori_size=24, healpix_order=1, n_particles=4, seed=26, snr=0.01)
    _setup(device, wl)
    monkeypatch.setenv("RB_FINE_SAMPLE_CAP", "8")
    with pytest.raises(RelionB200Error) as e:
        device.expectation_some_particles(wl.pool)
    assert e.value.status == RB_ERR_CAPACITY
    monkeypatch.delenv("RB_FINE_SAMPLE_CAP")
    device.expectation_some_particles(wl.pool)


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 96), (200, 300, 840), (1, 1, 1), (640, 1000, 3030), (129, 257, 33)])
def test_gemm_tf32x3(device, M, N, K):
    """The tcgen05 contraction behind the global-search coarse pass, held to FP32-equivalent accuracy: 3xTF32 splitting
    must land within 2e-6 of sum|a||b| of the float64 product, or within 4x the error of a sequential fp32 dot product
    for long K (plain TF32 would be ~1e-3)."""
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32) * rng.uniform(0.1, 10.0, (M, 1)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    got = device.gemm_tf32x3(A, B)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    bound = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    err = np.abs(got - want) / bound
    # yardstick: a sequential fp32 dot product of the same data (the reference's kernels sum in fp32, in pixel order)
    seq = np.zeros((min(M, 64), min(N, 64)), np.float32)
    for k in range(K):
        seq += A[:64, k:k + 1] * B[None, :64, k]
    err_seq = np.abs(seq - want[:64, :64]) / bound[:64, :64]
    assert err.max() <= max(2e-6, 4 * err_seq.max()), (err.max(), err_seq.max())
    assert err.mean() <= max(3e-7, 2 * err_seq.mean()), (err.mean(), err_seq.mean())


def test_pool_global_search_simt_and_tensor_paths_agree(device, oracle, monkeypatch):
    """Global search: the SIMT coarse kernel (RB_COARSE_GEMM=0) and the tensor-core contraction (RB_COARSE_GEMM=2) must
    both match the oracle, and select the same significant coarse samples as each other."""
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=12, nr_classes=2, seed=27, snr=0.3)
    monkeypatch.setenv("RB_COARSE_GEMM", "0")
    r0, _ = _compare_pool(device, oracle, wl)
    monkeypatch.setenv("RB_COARSE_GEMM", "2")
    r1, _ = _compare_pool(device, oracle, wl)
    np.testing.assert_allclose(r0.particles["min_diff2_coarse"], r1.particles["min_diff2_coarse"], rtol=2e-5)
    assert np.mean(r0.particles["nr_significant_coarse"] == r1.particles["nr_significant_coarse"]) >= 0.9
    assert np.array_equal(r0.particles["best_ihidden_over"], r1.particles["best_ihidden_over"])
    # the multi-CTA weight conversion used for large orientation grids must reproduce the one-CTA-per-particle kernel
    monkeypatch.setenv("RB_WEIGHTS_LARGE", "1")
    r2, _ = _compare_pool(device, oracle, wl)
    for key in ("nr_significant_coarse", "best_ihidden_over", "n_fine_samples", "min_diff2_coarse", "significant_weight_coarse", "sum_weight_coarse"):
        assert np.array_equal(r1.particles[key], r2.particles[key]), key


@pytest.mark.parametrize("maxsig,fraction,snr", [(0, 0.999, 0.3), (7, 0.999, 0.01), (0, 0.0, 0.3), (0, 1.0, 0.05), (3, 0.5, 0.01)])
def test_large_weight_conversion_matches_single_cta(device, monkeypatch, maxsig, fraction, snr):
    """The multi-CTA coarse weight conversion (histogram over the top 13 bits + seeded radix descent) against the
    one-CTA-per-particle kernel on the same pool: identical thresholds, counts, sums and fine-pass lists, including the
    --maxsig branch and the 'nothing crosses the threshold' branch (adaptive_fraction 0)."""
    wl = make_workload(ori_size=24, healpix_order=1, n_particles=9, nr_classes=2, seed=60 + maxsig, snr=snr, adaptive_fraction=fraction)
    wl.model.maximum_significants = maxsig
    res = []
    for large in ("100000000", "1"):
        monkeypatch.setenv("RB_WEIGHTS_LARGE", large)
        _setup(device, wl)
        res.append(device.expectation_some_particles(wl.pool).particles)
    for key in ("nr_significant_coarse", "significant_weight_coarse", "sum_weight_coarse", "min_diff2_coarse", "n_fine_orient",
                "n_fine_samples", "best_ihidden_over", "sum_weight", "pmax"):
        assert np.array_equal(res[0][key], res[1][key]), (key, res[0][key], res[1][key])
    if maxsig:
        assert res[1]["nr_significant_coarse"].max() <= maxsig


@pytest.mark.parametrize("n,count,r_max", [(32, 24, 16), (24, 7, 9), (16, 300, 8)])
def test_backproject_posed_against_reference_algorithm(device, n, count, r_max):
    """BASELINE config #2: rb_backproject_posed against the restated BackProjector::backproject2Dto3D (float64) on random
    poses; images already CTF-multiplied, weights ctf^2 with some non-positive entries (skipped pixels)."""
    from oracle.backproject_posed import backproject2Dto3D
    rng = np.random.default_rng(n * 100 + count)
    xs = n // 2 + 1
    pad = synth.pad_size_for(r_max, 2.0)
    shape = (pad, pad, pad // 2 + 1)
    device.bp_init(0, shape, r_max, 2.0)
    F = (rng.standard_normal((count, n, xs)) + 1j * rng.standard_normal((count, n, xs))).astype(np.complex64)
    W = rng.uniform(-0.1, 1.0, (count, n, xs)).astype(np.float32)
    eul = synth.inverse_euler_f32(rng.uniform(-180, 180, count), rng.uniform(0, 180, count), rng.uniform(0, 360, count))
    device.backproject_posed(0, n, F, W, eul)
    gre, gim, gw = device.bp_get(0)
    data = np.zeros(shape, np.complex128); weight = np.zeros(shape, np.float64)
    for i in range(count):
        backproject2Dto3D(data, weight, F[i], eul[i].reshape(3, 3).astype(np.float64), W[i], r_max, 2.0)
    assert np.abs(weight).max() > 0
    for got, want in ((gre, data.real), (gim, data.imag), (gw, weight)):
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    # the staged (device-resident) entry points do the same
    device.bp_clear(0)
    device.bp_posed_stage(n, F, W, eul)
    device.bp_posed_run(0)
    gre2, _, gw2 = device.bp_get(0)
    assert np.abs(gw2 - weight).max() <= 2e-5 * np.abs(weight).max()
    assert np.abs(gre2 - data.real).max() <= 2e-5 * np.abs(data.real).max()


@pytest.mark.parametrize("n,count,with_ctf", [(32, 80, True), (24, 7, True), (32, 70, False)])
def test_backproject_posed_from_raw_images(device, n, count, with_ctf):
    """rb_backproject_posed_raw (transform, CenterFFTbySign, origin shift, CTF, DC removal on the device) against the prepared
    entry point fed with the restated preparation of Reconstructor::backprojectOneParticle (oracle/backproject_posed.py),
    and against the restated backproject2Dto3D itself."""
    from oracle.backproject_posed import backproject2Dto3D, prepare_particle
    rng = np.random.default_rng(1000 + n + count)
    xs = n // 2 + 1
    r_max, angpix = n // 2 - 1, 1.7
    pad = synth.pad_size_for(r_max, 2.0)
    shape = (pad, pad, pad // 2 + 1)
    imgs = rng.standard_normal((count, n, n)).astype(np.float32)
    shift = rng.uniform(-3.0, 3.0, (count, 2)); shift[0] = 0.0
    eul = synth.inverse_euler_f32(rng.uniform(-180, 180, count), rng.uniform(0, 180, count), rng.uniform(0, 360, count))
    ctf = None
    cimgs = [None] * count
    if with_ctf:
        defU = rng.uniform(8000, 22000, count); defV = defU + rng.uniform(-600, 600, count); ang = rng.uniform(0, 180, count)
        bfac = rng.uniform(0, 80, count)
        ctf = dict(defU=defU, defV=defV, defAngle=ang, Bfac=bfac, kV=[300.0], Cs=[2.7], Q0=[0.1])
        cimgs = [synth.CTF(defU[i], defV[i], ang[i], Bfac=bfac[i]).fftw_image(n, n, angpix) for i in range(count)]
    device.bp_init(0, shape, r_max, 2.0)
    device.backproject_posed_raw(0, imgs, eul, shift=shift, ctf=ctf, pixel_size=angpix)
    got = device.bp_get(0)
    prepared = [prepare_particle(imgs[i], shift[i], cimgs[i]) for i in range(count)]
    F = np.stack([p[0] for p in prepared]).astype(np.complex64); W = np.stack([p[1] for p in prepared]).astype(np.float32)
    device.bp_clear(0)
    device.backproject_posed(0, n, F, W, eul)
    want = device.bp_get(0)
    assert np.abs(want[2]).max() > 0
    for a, b in zip(got, want):
        assert np.abs(a - b).max() <= 3e-5 * np.abs(b).max()
    if count <= 10:
        data = np.zeros(shape, np.complex128); weight = np.zeros(shape, np.float64)
        for i in range(count):
            backproject2Dto3D(data, weight, prepared[i][0], eul[i].reshape(3, 3).astype(np.float64), prepared[i][1], r_max, 2.0)
        for a, b in zip(got, (data.real, data.imag, weight)):
            assert np.abs(a - b).max() <= 3e-5 * np.abs(b).max()


@pytest.mark.parametrize("coarse", ["gemm", "simt"])
def test_pool_2d_classification(device, oracle, monkeypatch, coarse):
    """BASELINE config #1 regime: 2D references (project2Dmodel / backproject2D), K classes, psi-only sampling with
    2 oversampled psi per coarse one.  The library embeds a 2D reference in a two-plane volume (rb_set_reference with
    mdlZ == 1); the oracle's 2D back-projection is the reference's backproject2D (no circle bound)."""
    monkeypatch.setenv("RB_COARSE_GEMM", "2" if coarse == "gemm" else "0")
    wl = make_workload(ori_size=32, n_particles=24, nr_classes=3, seed=33, snr=0.2, ref_dim=2, psi_step=10.0)
    wl.model.bp_circle_bound = False
    res, ores = _compare_pool(device, oracle, wl)
    assert np.mean(res.particles["best_class"] == wl.truth["cls"]) >= 0.9
    gre, gim, gw = device.bp_get(0)
    assert gw.shape == wl.bp_shape


def test_class2d_per_class_prior_offsets(device, oracle):
    """2D references carry their own centre of the translation prior (mymodel.prior_offset_class,
    acc_ml_optimiser_impl.h:2100-2104, 2673-2677): per-class pdf_offset in the fine-pass weights, the first class' block in
    the coarse pass (:2187-2196), per-class wsum_sigma2_offset and the wsum_prior_offset_class sums (:2847-2851).  A tight
    sigma2_offset makes the prior matter."""
    wl = make_workload(ori_size=32, n_particles=24, nr_classes=3, seed=35, snr=0.3, ref_dim=2, psi_step=10.0)
    wl.model.prior_offset_class = np.array([[0.8, -0.6], [-1.3, 0.4], [0.2, 1.7]])
    wl.model.sigma2_offset = 2.0
    wl.model.bp_circle_bound = False          # the reference's backproject2D has no circle bound (see test_pool_2d_classification)
    res, ores = _compare_pool(device, oracle, wl, pose_frac=0.95)
    assert np.abs(ores.wsum_prior_offset_class).max() > 0
    # and the prior is really the per-class one: with the particle's own (zero) centre the sums differ
    wl.model.prior_offset_class = None
    _setup(device, wl)
    plain = device.expectation_some_particles(wl.pool)
    assert np.abs(plain.particles["wsum_sigma2_offset"] - res.particles["wsum_sigma2_offset"]).max() > 1e-3
    assert not np.abs(plain.wsum_prior_offset_class).any()


@pytest.mark.parametrize("ori,cur,local", [(32, 32, False), (40, 28, False), (32, 32, True)])
def test_pool_prepare_matches_reference_algorithm(device, ori, cur, local):
    """getFourierTransformsAndCtfs on the device (rb_pool_prepare) against its numpy restatement
    (oracle/prepare.py): transforms of the unmasked and the zero-masked image, CTF image, highres_Xi2, power spectrum;
    then the E-step on the device-prepared slot against the E-step on the oracle-prepared pool."""
    from relion_b200.workload import raw_pool_from
    from oracle.prepare import prepared_pool as oracle_prepared_pool
    wl = make_workload(ori_size=ori, current_size=cur, healpix_order=2 if local else 1, n_particles=9, seed=70 + ori + cur, snr=0.3,
                       local_search=local, nr_groups=1)
    raw = raw_pool_from(wl, seed=3)
    _setup(device, wl)
    power = device.pool_prepare(0, raw)
    F, F0, Cc, xi2 = device.pool_download(0, cur)
    pool, opower = oracle_prepared_pool(wl, raw)
    for got, want in ((F, pool.Fimg), (F0, pool.Fimg_nomask)):
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    assert np.abs(F - F0).max() > 1e-4 * np.abs(F0).max()          # the mask did something
    np.testing.assert_allclose(Cc, pool.Fctf, rtol=0, atol=2e-6)
    np.testing.assert_allclose(xi2, pool.highres_Xi2, rtol=1e-4, atol=1e-9 * max(np.abs(pool.highres_Xi2).max(), 1.0))
    np.testing.assert_allclose(power, opower, rtol=1e-4, atol=1e-6 * max(opower.max(), 1e-30))
    if cur < ori:
        assert xi2.min() > 0
    got = device.estep_slot(0)
    for k in range(wl.model.nr_classes):
        device.bp_clear(k)
    want = device.expectation_some_particles(pool)
    assert np.array_equal(got.particles["best_ihidden_over"], want.particles["best_ihidden_over"])
    assert np.mean(got.particles["nr_significant_coarse"] == want.particles["nr_significant_coarse"]) >= 0.8
    np.testing.assert_allclose(got.particles["dLL_nolog"], want.particles["dLL_nolog"], rtol=1e-4)


def test_pool_prepare_beam_tilt_and_mtf_factor(device):
    """Beam-tilt demodulation and MTF division (ObservationModel::demodulatePhase / divideByMtf, obs_model.cpp:528-626, applied
    to both transforms at acc_ml_optimiser_impl.h:535-536) as one complex factor image per optics group, multiplied in after
    windowing: conj(phase correction) * avgMTF / MTF."""
    from relion_b200.workload import raw_pool_from
    from oracle.prepare import prepared_pool as oracle_prepared_pool
    ori, cur = 40, 28
    wl = make_workload(ori_size=ori, current_size=cur, healpix_order=1, n_particles=8, seed=97, snr=0.3, nr_groups=1)
    raw = raw_pool_from(wl, seed=4)
    rng = np.random.default_rng(5)
    xs = cur // 2 + 1
    phase = np.exp(-1j * rng.uniform(-0.6, 0.6, (1, cur, xs)))            # conj of a tilt phase ramp stand-in
    mtf = rng.uniform(0.4, 1.0, (1, cur, xs)); avg = rng.uniform(0.4, 1.0, (1, cur, xs))
    factor = (phase * avg / mtf).astype(np.complex64)
    raw.og_fourier_factor = factor
    _setup(device, wl)
    device.pool_prepare(0, raw)
    F, F0, _, xi2 = device.pool_download(0, cur)
    pool, _ = oracle_prepared_pool(wl, raw)
    for got, want in ((F, pool.Fimg * factor[0]), (F0, pool.Fimg_nomask * factor[0])):
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    np.testing.assert_allclose(xi2, pool.highres_Xi2, rtol=1e-4)       # the power beyond the current size is taken before the correction


def test_pool_prepare_noise_filled_mask(device):
    """RELION's default soft mask (no --zero_mask): the edge blends into a noise image with the model's noise spectrum
    (makeNoiseImage + cosineFilter, utilities_impl.h:231-371, acc_ml_optimiser_impl.h:355-400, 660-668).  The random numbers
    cannot match RELION's (its curand and CPU generators do not match each other either), so: (1) given the device's own noise
    images, the masked transforms must equal the restated blend exactly; (2) the noise images must have the model's spectrum;
    (3) same seed -> same image, other seed -> other image."""
    from relion_b200.workload import raw_pool_from
    from oracle import prepare as prep
    ori = 48
    wl = make_workload(ori_size=ori, healpix_order=1, n_particles=24, seed=91, snr=0.3, nr_groups=1)
    # a coloured noise spectrum, so that the shape is tested and not just the level
    nshell = ori // 2 + 1
    wl.model.sigma2_noise = np.atleast_2d(1e-3 * (1.0 + 3.0 * np.exp(-np.arange(nshell) / 6.0)))
    wl.model.sigma2_fudge = 1.3
    raw = raw_pool_from(wl, seed=5)
    raw.noise_seed = 1234 + np.arange(raw.n_particles, dtype=np.int64)
    raw.noise_seed[3] = raw.noise_seed[2]                      # two particles with the same seed
    _setup(device, wl)
    device.pool_prepare(0, raw)
    F, F0, _, _ = device.pool_download(0, ori)
    noise = device.debug_prep_noise(raw.n_particles, ori)
    # (1) exact blend
    for p in range(raw.n_particles):
        dx, dy = (int(np.sign(v) * np.floor(abs(v) + 0.5)) for v in raw.old_offset[p])
        t = prep.translate_and_norm(np.asarray(raw.images[p], np.float64), dx, dy, 1.0 if raw.norm_factor is None else raw.norm_factor[p])
        masked = prep.noise_mask(t, noise[p].astype(np.float64), raw.mask_radius, raw.width_mask_edge)
        want, _ = prep.normalize_and_transform(masked, ori)
        assert np.abs(F[p] - want).max() <= 2e-5 * np.abs(want).max(), p
    assert np.abs(F - F0).max() > 1e-4 * np.abs(F0).max()
    # (3) seeds
    assert np.array_equal(noise[2], noise[3]) and not np.array_equal(noise[0], noise[1])
    # (2) spectrum: mean |FT|^2 per shell over the particles against 2 sigma2_fudge sigma2_noise (columns 0 < x < n/2)
    spec = np.sqrt(wl.model.sigma2_fudge * wl.model.sigma2_noise[0])
    ires, expect = prep.noise_image_shell_power(ori, spec)
    P2 = np.zeros(ires.shape)
    for p in range(raw.n_particles):
        if p == 3:
            continue
        c = np.roll(noise[p].astype(np.float64), (-(ori // 2), -(ori // 2)), axis=(0, 1))
        P2 += np.abs(np.fft.rfft2(c) / (ori * ori)) ** 2
    P2 /= raw.n_particles - 1
    inner = np.zeros(ires.shape, bool); inner[:, 1:ori // 2] = True
    for shell in range(4, ori // 2 - 1):
        sel = inner & (ires == shell)
        assert abs(P2[sel].mean() / expect[sel].mean() - 1.0) < 0.25, (shell, P2[sel].mean(), expect[sel].mean())
    assert abs(noise.mean()) < 0.05 * noise.std()
    # and the E-step runs on the noise-masked slot
    res = device.estep_slot(0)
    assert np.all(res.particles["sum_weight"] > 0)


def test_estep_from_star_and_mrc_files_equals_in_memory_pool(device, tmp_path):
    """particles written as two MRC stacks + a RELION 3.1 STAR file, streamed back through the native
    feed (rb_feed_*) and ParticleSet.pool, must give the same prepared slot and the same E-step result as the in-memory
    RawParticlePool they were written from - bit for bit, the inputs are identical."""
    from relion_b200 import particle_io, star
    from relion_b200.workload import raw_pool_from
    wl = make_workload(ori_size=32, current_size=24, healpix_order=1, n_particles=10, seed=91, snr=0.3, nr_groups=1)
    raw = raw_pool_from(wl, seed=5)
    raw.norm_factor = np.round(raw.norm_factor, 3)                                      # survive the text round trip exactly
    raw.old_offset = np.round(raw.old_offset, 2)
    P, n, ps = raw.n_particles, wl.model.ori_size, wl.model.pixel_size
    images = np.asarray(raw.images, np.float32)
    half = P // 2
    particle_io.write_mrc(str(tmp_path / "micA.mrcs"), images[:half], ps)
    particle_io.write_mrc(str(tmp_path / "micB.mrcs"), images[half:], ps)
    optics = star.StarTable("optics", {"rlnOpticsGroupName": ["opticsGroup1"], "rlnOpticsGroup": [1], "rlnVoltage": [300.0],
                                       "rlnSphericalAberration": [2.7], "rlnAmplitudeContrast": [0.1], "rlnImagePixelSize": [float(ps)],
                                       "rlnImageSize": [n], "rlnImageDimensionality": [2]})
    avg_norm = 0.97
    names = ["%06d@%s" % ((i if i < half else i - half) + 1, "micA.mrcs" if i < half else "micB.mrcs") for i in range(P)]
    parts = star.StarTable("particles", {
        "rlnImageName": names, "rlnGroupName": ["group1"] * P, "rlnOpticsGroup": [1] * P,
        "rlnDefocusU": [float(v) for v in raw.ctf_defU], "rlnDefocusV": [float(v) for v in raw.ctf_defV],
        "rlnDefocusAngle": [float(v) for v in raw.ctf_defAngle],
        "rlnOriginXAngst": [float(v) * ps for v in raw.old_offset[:, 0]], "rlnOriginYAngst": [float(v) * ps for v in raw.old_offset[:, 1]],
        "rlnNormCorrection": [avg_norm / float(v) for v in raw.norm_factor]})
    star.write_star(str(tmp_path / "particles.star"), [optics, parts])
    pset = particle_io.ParticleSet.read(str(tmp_path / "particles.star"))
    assert len(pset) == P and pset.image_size() == n
    _setup(device, wl)
    device.pool_prepare(0, raw)
    want_prep = device.pool_download(0, wl.model.current_size)
    want = device.estep_slot(0)
    for k in range(wl.model.nr_classes):
        device.bp_clear(k)
    feed = particle_io.ParticleFeed(image_size=n, max_particles=P, depth=2, n_threads=2)
    chunks = list(pset.stream(feed, pool_size=P, avg_norm_correction=avg_norm, mask_radius=raw.mask_radius, width_mask_edge=raw.width_mask_edge))
    assert len(chunks) == 1
    # the generator released the buffer when it finished; stream again and use the pool while it is valid
    for ids, pool in pset.stream(feed, pool_size=P, avg_norm_correction=avg_norm, mask_radius=raw.mask_radius, width_mask_edge=raw.width_mask_edge):
        np.testing.assert_array_equal(np.asarray(pool.images), images[ids])
        np.testing.assert_allclose(pool.norm_factor, raw.norm_factor[ids], rtol=1e-5)
        np.testing.assert_allclose(pool.old_offset, raw.old_offset[ids], atol=1e-5)
        np.testing.assert_allclose(pool.ctf_defU, raw.ctf_defU[ids], atol=1e-5)
        # remove the last-digit rounding of the text file: from here on the inputs are identical
        pool.norm_factor, pool.old_offset = raw.norm_factor[ids], raw.old_offset[ids]
        pool.ctf_defU, pool.ctf_defV, pool.ctf_defAngle = raw.ctf_defU[ids], raw.ctf_defV[ids], raw.ctf_defAngle[ids]
        device.pool_prepare(1, pool)
        got_prep = device.pool_download(1, wl.model.current_size)
        got = device.estep_slot(1)
    feed.close()
    for a, b in zip(got_prep, want_prep):
        np.testing.assert_array_equal(a, b)
    assert np.array_equal(got.particles["best_ihidden_over"], want.particles["best_ihidden_over"])
    np.testing.assert_array_equal(got.particles["dLL_nolog"], want.particles["dLL_nolog"])
    np.testing.assert_allclose(got.wsum_sigma2_noise, want.wsum_sigma2_noise, rtol=2e-6)     # float atomics: order of the adds


@pytest.mark.parametrize("ref_dim,with_fsc,whole", [(3, False, False), (3, True, False), (3, True, True), (2, False, False)])
def test_update_ssnr_on_device(device, ref_dim, with_fsc, whole):
    """rb_update_ssnr (BackProjector::updateSSNRarrays: sigma2, FSC-based tau2, data_vs_prior, Fourier
    coverage) on the device accumulator against the numpy restatement applied to the downloaded weights."""
    from oracle import reconstruct as rc
    ori = 32
    if ref_dim == 3:
        wl = make_workload(ori_size=ori, healpix_order=1, n_particles=40, seed=95, snr=0.5)
    else:
        wl = make_workload(ori_size=ori, n_particles=40, seed=96, snr=0.5, ref_dim=2, psi_step=30.0, nr_classes=2)
    _setup(device, wl)
    device.expectation_some_particles(wl.pool)
    _, _, w = device.bp_get(0)
    ns = ori // 2 + 1
    rng = np.random.default_rng(7)
    tau2 = 1e-3 / (1.0 + np.arange(ns)) ** 2
    tau2[5] = 0.0                                                   # the "use small value instead" branch
    fsc = np.clip(1.0 - np.arange(ns) / ns + 0.05 * rng.standard_normal(ns), -0.1, 1.0) if with_fsc else None
    avgctf2 = rng.uniform(0.3, 1.0, ns) if with_fsc else None
    got = device.update_ssnr(0, ori, tau2, tau2_fudge=2.0, fsc=fsc, avgctf2=avgctf2, update_tau2_with_fsc=with_fsc,
                             is_whole_instead_of_half=whole)
    want = rc.update_ssnr(w, ori, wl.r_max, wl.padding_factor, 2.0, tau2, fsc=fsc, avgctf2=avgctf2,
                          update_tau2_with_fsc=with_fsc, is_whole_instead_of_half=whole)
    for g, x, name in zip(got, want, ("tau2", "sigma2", "data_vs_prior", "fourier_coverage")):
        np.testing.assert_allclose(g, x, rtol=1e-9, atol=1e-300, err_msg=name)
    assert got[1].max() > 0 and (got[3] > 0).any()                  # something was measured
    if with_fsc:
        assert not np.allclose(got[0], tau2)                        # tau2 was replaced by the FSC-based estimate


@pytest.mark.parametrize("ori,cur,with_tau2", [(32, 32, False), (32, 32, True), (40, 28, True)])
def test_reconstruct_on_device(device, ori, cur, with_tau2):
    """rb_reconstruct (BackProjector::reconstruct, skip_gridding, + windowToOridimRealSpace +
    griddingCorrect on the device) against the numpy restatement applied to the same accumulator."""
    from oracle import reconstruct as rc
    wl = make_workload(ori_size=ori, current_size=cur, healpix_order=1, n_particles=60, seed=80 + ori, snr=0.5)
    _setup(device, wl)
    device.expectation_some_particles(wl.pool)
    tau2 = None
    if with_tau2:
        tau2 = 1e-3 / (1.0 + np.arange(ori // 2 + 1)) ** 2
        tau2[-2:] = 0.0                                        # the tau2 <= 0 branch (:1483-1487)
    got = device.reconstruct(0, ori, tau2=tau2, tau2_fudge=2.0, minres_map=2)
    gre, gim, gw = device.bp_get(0)
    want = rc.reconstruct(gre, gim, gw, ori, wl.r_max, wl.padding_factor, tau2=tau2, tau2_fudge=2.0, minres_map=2)
    assert np.abs(want).max() > 0
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
    f = rc.fsc(got.astype(np.float64), want)
    assert f.min() >= 0.9999, f


@pytest.mark.parametrize("ori,cur,with_tau2", [(32, 32, False), (32, 32, True), (40, 28, True)])
def test_reconstruct_with_iterative_gridding_on_device(device, ori, cur, with_tau2):
    """--dont_skip_gridding: rb_reconstruct_gridding (ten Pipe & Menon iterations in double, cuFFT Z2D / D2Z of the padded volume)
    against the numpy restatement that tests/test_reference_host.py pins on BackProjector::reconstruct(skip_gridding = false)."""
    from oracle import reconstruct as rc
    wl = make_workload(ori_size=ori, current_size=cur, healpix_order=1, n_particles=60, seed=83 + ori, snr=0.5)
    _setup(device, wl)
    device.expectation_some_particles(wl.pool)
    tau2 = 1e-3 / (1.0 + np.arange(ori // 2 + 1)) ** 2 if with_tau2 else None
    got = device.reconstruct(0, ori, tau2=tau2, tau2_fudge=2.0, minres_map=2, max_iter_preweight=10)
    gre, gim, gw = device.bp_get(0)
    want = rc.reconstruct(gre, gim, gw, ori, wl.r_max, wl.padding_factor, tau2=tau2, tau2_fudge=2.0, minres_map=2, max_iter_preweight=10)
    assert np.abs(want).max() > 0
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
    assert rc.fsc(got.astype(np.float64), want).min() >= 0.9999
    # and it is not the skip_gridding result
    plain = device.reconstruct(0, ori, tau2=tau2, tau2_fudge=2.0, minres_map=2)
    assert np.abs(got - plain).max() > 1e-3 * np.abs(want).max()


@pytest.mark.parametrize("ori,cur", [(32, 32), (40, 28)])
def test_reference_from_map_on_device(device, oracle, ori, cur):
    """rb_set_reference_from_map (Projector::computeFourierTransformMap on the device): central slices of the
    device-made reference against slices of the numpy restatement (synth.reference_ft, float64); power spectrum."""
    from oracle.bindings import Projector
    vol = synth.make_phantom(ori, n_blobs=30, seed=9)
    ps = device.set_reference_from_map(0, vol, current_size=cur)
    data, r_max = synth.reference_ft(vol, current_size=cur)
    ref = Projector(data.astype(np.complex64), r_max, 2.0)
    rng = np.random.default_rng(1)
    eul = synth.inverse_euler_f32(rng.uniform(-180, 180, 6), rng.uniform(0, 180, 6), rng.uniform(0, 360, 6))
    got = device.project(0, cur, eul)
    for i in range(len(eul)):
        want = oracle.project(ref, cur, eul[i])
        assert np.abs(got[i] - want).max() <= 2e-5 * np.abs(want).max()
    # radial power spectrum of the reference (:497-568)
    pad = data.shape[0]; h = (pad - 1) // 2
    k = np.arange(-h, h + 1); kx = np.arange(pad // 2 + 1)
    kz, ky, kxx = np.meshgrid(k, k, kx, indexing="ij")
    r2 = kz * kz + ky * ky + kxx * kxx
    inr = r2 <= int(np.floor(r_max * 2.0 + 0.5)) ** 2
    ires = np.floor(np.sqrt(r2[inr]) / 2.0 + 0.5).astype(int)
    want_ps = np.bincount(ires, weights=0.5 * np.abs(data[inr]) ** 2, minlength=ori // 2 + 1) / np.maximum(np.bincount(ires, minlength=ori // 2 + 1), 1)
    np.testing.assert_allclose(ps[:len(want_ps)], want_ps[:len(ps)], rtol=2e-4, atol=1e-9 * want_ps.max())


def test_refinement_iterations_stay_on_the_device(device):
    """The three 'next' rows chained around the hot path: map -> reference (rb_set_reference_from_map) -> E-step over a pool
    of RAW images (rb_pool_prepare + rb_estep_slot) -> map (rb_reconstruct), three times, starting from a blurred phantom.
    The reconstruction must converge towards the phantom the particles were projected from (FSC at low resolution)."""
    from oracle import reconstruct as rc
    from relion_b200.workload import raw_pool_from
    ori = 32
    wl = make_workload(ori_size=ori, healpix_order=1, n_particles=300, seed=90, snr=1.0, nr_groups=1)
    raw = raw_pool_from(wl, seed=4, max_old_offset=0.0, mask_radius=0.45 * ori)
    raw.norm_factor[:] = 1.0
    truth = synth.make_phantom(ori, n_blobs=40, seed=1993)
    # blurred start: keep only the lowest shells
    F = np.fft.fftn(truth)
    f = np.fft.fftfreq(ori) * ori
    kz, ky, kx = np.meshgrid(f, f, f, indexing="ij")
    start = np.real(np.fft.ifftn(F * np.exp(-(kz ** 2 + ky ** 2 + kx ** 2) / (2 * 2.0 ** 2))))
    device.set_model(wl.model)
    device.set_sampling(wl.sampling)
    device.bp_init(0, wl.bp_shape, wl.r_max, wl.padding_factor)
    cur = start
    fscs = []
    for it in range(3):
        device.set_reference_from_map(0, cur)
        device.bp_clear(0)
        device.pool_prepare(0, raw, want_power=False)
        res = device.estep_slot(0)
        cur = device.reconstruct(0, ori).astype(np.float64)
        fscs.append(rc.fsc(cur, truth))
    assert fscs[0][1:4].min() > 0.9, fscs[0]
    assert fscs[-1][1:6].min() > 0.95, fscs[-1]
    assert fscs[-1][4:8].mean() >= fscs[0][4:8].mean() - 0.02, (fscs[0], fscs[-1])
    assert np.mean(res.particles["best_idir"] == wl.truth["idir"]) > 0.5      # 30-degree grid: neighbours / pole degeneracy allowed


@pytest.mark.parametrize("group", ["C1", "D2", "C3"])
def test_symmetrise_on_device(device, group):
    """rb_bp_symmetrise (enforceHermitianSymmetry + applyPointGroupSymmetry on the device accumulator) against the numpy
    restatement; D2 = three two-fold axes (BASELINE config #5), C3 exercises the interpolation."""
    from oracle import reconstruct as rc
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=40, seed=95, snr=0.5)
    _setup(device, wl)
    device.expectation_some_particles(wl.pool)
    gre, gim, gw = device.bp_get(0)
    rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    rots = {"C1": [], "D2": [np.diag([1.0, -1, -1]), np.diag([-1.0, 1, -1]), np.diag([-1.0, -1, 1])],
            "C3": [rz(2 * np.pi / 3), rz(4 * np.pi / 3)]}[group]
    device.bp_symmetrise(0, rots)
    sre, sim, sw = device.bp_get(0)
    wre, wim, ww = rc.symmetrise(gre, gim, gw, wl.r_max, wl.padding_factor, rots)
    for got, want in ((sre, wre), (sim, wim), (sw, ww)):
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    assert np.abs(sw - gw).max() > 0


@pytest.mark.parametrize("nr_asu,twist,rise,with_c2", [(5, 22.03, 1.408, False), (4, -166.7, 0.9, True), (3, 30.0, 0.0, False)])
def test_symmetrise_with_helical_symmetry_on_device(device, nr_asu, twist, rise, with_c2):
    """rb_bp_symmetrise_helical (applyHelicalSymmetry between the Hermitian fold and the point group) against the numpy
    restatement, itself pinned against BackProjector::symmetrise(nr_helical_asu, twist, rise) in tests/test_reference_host.py."""
    from oracle import reconstruct as rc
    wl = make_workload(ori_size=32, healpix_order=1, n_particles=40, seed=96, snr=0.5)
    _setup(device, wl)
    device.expectation_some_particles(wl.pool)
    gre, gim, gw = device.bp_get(0)
    rots = [np.diag([-1.0, -1, 1])] if with_c2 else []
    device.bp_symmetrise(0, rots, helical=(nr_asu, twist, rise, wl.model.ori_size))
    sre, sim, sw = device.bp_get(0)
    wre, wim, ww = rc.symmetrise(gre, gim, gw, wl.r_max, wl.padding_factor, rots, helical=(nr_asu, twist, rise, wl.model.ori_size))
    for got, want in ((sre, wre), (sim, wim), (sw, ww)):
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    pre, pim, pw = rc.symmetrise(gre, gim, gw, wl.r_max, wl.padding_factor, rots)
    assert np.abs(ww - pw).max() > 0.1 * np.abs(pw).max()


@pytest.mark.parametrize("mode", ["skip_align", "skip_rotate"])
def test_pool_only_classify(device, oracle, mode):
    """--skip_align / --skip_rotate (rb_model.do_skip_rotate): every particle keeps its own orientation (one-entry lists into sampling
    tables that hold the pool's orientations), the orientation prior is pdf_class, and with --skip_align the particle's fractional
    offset is applied once as rb_particles.pre_shift instead of being its only sampled translation."""
    from relion_b200.workload import make_skip_align_workload
    wl = make_skip_align_workload(skip_rotate_only=(mode == "skip_rotate"), n_particles=14, nr_classes=3, seed=140, snr=0.3)
    res, ores = _compare_pool(device, oracle, wl)
    assert np.mean(res.particles["best_class"] == wl.truth["cls"]) >= 0.9
    assert np.array_equal(res.particles["best_itrans"], wl.truth["itrans"])
    # the prior is pdf_class: other class weights move the class sums
    np.testing.assert_allclose(res.wsum_pdf_class, ores.wsum_pdf_class, rtol=1e-4)
    if mode == "skip_align":
        plain = make_skip_align_workload(n_particles=14, nr_classes=3, seed=140, snr=0.3)
        plain.pool.pre_shift = None
        other = device.expectation_some_particles(plain.pool)
        assert np.abs(other.particles["dLL_nolog"] - res.particles["dLL_nolog"]).max() > 0.1


def test_references_ending_inside_the_window_need_ref_max_r(device):
    """A pool whose references end inside the image window changes the pixel sets of the fine pass (rb_model.ref_max_r): without
    the field the library refuses the pool instead of summing rows the reference skips; so does a left matrix with 2D references."""
    from relion_b200.capi import RelionB200Error
    wl = make_workload(ori_size=40, ref_box=32, healpix_order=1, n_particles=3, seed=130, snr=0.3)
    assert wl.model.ref_max_r == wl.r_max == 16
    wl.model.ref_max_r = 0
    _setup(device, wl)
    with pytest.raises(RelionB200Error) as e:
        device.expectation_some_particles(wl.pool)
    assert "ref_max_r" in str(e.value)
    wl.model.ref_max_r = 16
    device.set_model(wl.model); device.set_sampling(wl.sampling)
    device.expectation_some_particles(wl.pool)
    w2 = make_workload(ori_size=32, n_particles=3, nr_classes=1, seed=131, snr=0.5, ref_dim=2, psi_step=12.0)
    w2.pool.mat_left = np.eye(3) * 1.1
    _setup(device, w2)
    with pytest.raises(RelionB200Error) as e:
        device.expectation_some_particles(w2.pool)
    assert "3D references" in str(e.value)


_MAG_L = [[1.03, 0.012, 0.0], [-0.008, 0.96, 0.0], [0.0, 0.0, 1.0]]
_BODY_R = [[0.9553364891, -0.2955202067, 0.0], [0.2955202067, 0.9553364891, 0.0], [0.0, 0.0, 1.0]]


@pytest.mark.parametrize("case", ["mag_local", "mag_global", "left_only_local", "right_only_global", "bigger_box_global", "bigger_box_local",
                                  "smaller_box_global", "wrong_scale_global", "bigger_box_local_orientation_major", "bigger_box_cc_global"])
def test_pool_with_left_and_right_matrices(device, oracle, monkeypatch, case):
    """rb_particles.mat_left / mat_right (MBL / MBR: orientation matrices inverse(L A R) in the coarse pass, the fine pass and the
    store stage; cuda_kernel_make_eulers_3D<invert, doL, doR>, generateEulerMatrices(..., L, R)): anisotropic magnification with a
    body rotation, each matrix alone, and optics groups whose box differs from the references' (applyScaleDifference).  A bigger
    image box: the references end inside the image window and the fine pass skips the rows beyond (rb_model.ref_max_r); a
    deliberately wrong scale: the projection leaves the reference for the outer image shells, which must then read zero."""
    kw = dict(ori_size=32, n_particles=12, seed=120, snr=0.2)
    if case.endswith("orientation_major"):
        monkeypatch.setenv("RB_BAND", "0")                      # k_diff2_fine / k_store instead of the band-major kernels
        case = case[:-len("_orientation_major")]
    if "_cc_" in case:
        kw.update(do_cc=True)                                   # the cross-correlation kernels have the same row rule (diff2.h:657-666)
    if case.endswith("local"):
        kw.update(healpix_order=2, local_search=True)
    else:
        kw.update(healpix_order=1)
    if case.startswith("mag"):
        kw.update(mat_left=_MAG_L, mat_right=_BODY_R)
    elif case.startswith("left_only"):
        kw.update(mat_left=_MAG_L)
    elif case.startswith("right_only"):
        kw.update(mat_right=_BODY_R)
    elif case.startswith("bigger_box"):
        kw.update(ori_size=40, ref_box=32)
    elif case.startswith("smaller_box"):
        kw.update(ori_size=32, ref_box=40)
    else:
        kw.update(ori_size=40, ref_box=32, mat_left=np.eye(3) * 0.8)
    wl = make_workload(**kw)
    res, ores = _compare_pool(device, oracle, wl)
    # the matrices matter: without them the pool has another likelihood
    plain = make_workload(**kw)
    plain.pool.mat_left = plain.pool.mat_right = None
    other = device.expectation_some_particles(plain.pool)
    assert np.abs(other.particles["dLL_nolog"] - res.particles["dLL_nolog"]).max() > (0.01 if wl.model.do_cc else 0.1)
    # ... and a pool with the matrices again gets the coarse matrices rebuilt
    for k in range(len(wl.refs)):
        device.bp_clear(k)
    again = device.expectation_some_particles(wl.pool)
    assert np.array_equal(again.particles["best_ihidden_over"], res.particles["best_ihidden_over"])
    np.testing.assert_allclose(again.particles["dLL_nolog"], res.particles["dLL_nolog"], rtol=1e-6)


@pytest.mark.parametrize("local", [True, False])
def test_pool_128px_against_reference_kernels(device, local):
    """A mid-size pool (128-px box, several 128-orientation tiles, hundreds of K-blocks in the tensor-core kernels) against
    RELION's own compiled ALTCPU kernels driven by the restated orchestration (oracle kind 'reference')."""
    from oracle.bindings import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/librefkernels.so not built")
    if local:
        wl = make_workload(ori_size=128, healpix_order=3, offset_range=3.0, offset_step=1.0, n_particles=6, seed=101, snr=0.1,
                           local_search=True, pixel_size=2.0, n_blobs=60)
    else:
        wl = make_workload(ori_size=128, current_size=64, healpix_order=2, offset_range=5.0, offset_step=2.0, n_particles=6, seed=102,
                           snr=0.1, pixel_size=2.0, n_blobs=60)
    res, ores = _compare_pool(device, Oracle("reference"), wl)
    assert np.array_equal(res.particles["best_ihidden_over"], ores.particles["best_ihidden_over"])


def test_pipelined_slots_equal_sequential_calls(device):
    """launch(i) ; fetch(i-1) ; upload(i+1) over two device slots (rb_estep_slot_nocopy / rb_estep_fetch / rb_pool_upload)
    must give exactly the per-particle results of one rb_estep_pool call per pool."""
    wls = [make_workload(ori_size=32, healpix_order=1, n_particles=7 + i, seed=110 + i, snr=0.3) for i in range(4)]
    _setup(device, wls[0])
    seq = [device.expectation_some_particles(w.pool).particles for w in wls]
    for k in range(wls[0].model.nr_classes):
        device.bp_clear(k)
    got = [None] * len(wls)
    device.pool_upload(0, wls[0].pool)
    for i in range(len(wls)):
        device.estep_slot_nocopy(i % 2)
        if i >= 1:
            got[i - 1] = device.estep_fetch((i - 1) % 2).particles
        if i + 1 < len(wls):
            device.pool_upload((i + 1) % 2, wls[i + 1].pool)
    got[-1] = device.estep_fetch((len(wls) - 1) % 2).particles
    for a, b in zip(seq, got):
        for key in ("best_ihidden_over", "nr_significant_coarse", "n_fine_samples", "min_diff2_coarse", "sum_weight", "dLL_nolog"):
            assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("local", [False, True])
def test_pool_many_translations(device, oracle, local):
    """81 coarse translations (offset range 5, step 1): more than the 32-translation tiles of the fused kernel (local searches
    take three passes of 32 / 32 / 17 translations through it) and more than one 64-entry prior table of the multi-CTA weight
    conversion."""
    wl = make_workload(ori_size=32, healpix_order=2 if local else 1, offset_range=5.0, offset_step=1.0, n_particles=5, seed=120,
                       snr=0.3, local_search=local)
    assert wl.sampling.n_trans == 81
    _compare_pool(device, oracle, wl)


@pytest.mark.parametrize("local", [False, True])
def test_pool_class_with_zero_prior_is_skipped(device, oracle, local):
    """A class whose pdf_class is zero is never evaluated (acc_ml_optimiser_impl.h:1069): its Mweight entries stay at lowest()
    through every coarse path, no particle may pick it, and its accumulator stays empty."""
    wl = make_workload(ori_size=32, healpix_order=2 if local else 1, n_particles=8, nr_classes=3, seed=121, snr=0.3, local_search=local)
    wl.model.pdf_class = np.array([0.5, 0.0, 0.5])
    res, _ = _compare_pool(device, oracle, wl, pose_frac=0.0)
    assert not np.any(res.particles["best_class"] == 1)
    assert np.all(device.bp_get(1)[2] == 0)


@pytest.mark.parametrize("kw", [dict(ori_size=32, healpix_order=2, n_particles=10, nr_classes=1, seed=140, snr=0.2, local_search=True),
                                dict(ori_size=32, healpix_order=1, n_particles=9, nr_classes=2, seed=141, snr=0.3),
                                dict(ori_size=40, current_size=28, healpix_order=1, n_particles=7, nr_classes=1, seed=142, snr=0.05)])
@pytest.mark.parametrize("flags", [dict(), dict(do_map=False)])
def test_band_major_path_equals_orientation_major_path(device, oracle, monkeypatch, kw, flags):
    """The band-major fine pass / store stage (kernels_band.cu: projection per radial band into band-ordered slices, streaming
    diff2, store by (band, orientation chunk)) against the orientation-major kernels (RB_BAND=0) on the same pool: same
    fine-pass lists and poses, sums equal up to the summation order; both against the oracle.  --no_map adds the extra
    x = 0 half column to the store-stage pixel set."""
    wl = make_workload(**kw)
    for k, v in flags.items():
        setattr(wl.model, k, v)
    out = {}
    for band in ("0", "1"):
        monkeypatch.setenv("RB_BAND", band)
        res, _ = _compare_pool(device, oracle, wl, pose_frac=0.9 if kw["snr"] < 0.1 else 0.995)
        out[band] = (res, [device.bp_get(k) for k in range(wl.model.nr_classes)])
    a, b = out["0"][0].particles, out["1"][0].particles
    for key in ("nr_significant_coarse", "n_fine_orient", "n_fine_samples", "best_ihidden_over", "n_bp_orient"):
        assert np.array_equal(a[key], b[key]), key
    for key in ("min_diff2", "sum_weight", "pmax", "dLL_nolog", "wsum_norm_correction", "wsum_XA", "wsum_AA", "sumw", "wsum_sigma2_offset"):
        np.testing.assert_allclose(a[key], b[key], rtol=2e-4, atol=1e-6 * max(np.abs(a[key]).max(), 1e-30), err_msg=key)
    np.testing.assert_allclose(out["0"][0].wsum_sigma2_noise, out["1"][0].wsum_sigma2_noise, rtol=1e-3, atol=1e-5 * np.abs(out["0"][0].wsum_sigma2_noise).max())
    for ka, kb in zip(out["0"][1], out["1"][1]):
        for x, y in zip(ka, kb):
            assert np.abs(x - y).max() <= 1e-4 * max(np.abs(x).max(), 1e-12)


@pytest.mark.parametrize("cache_slices", [0, 5])
def test_pool_store_stage_without_slice_cache(device, oracle, monkeypatch, cache_slices):
    """The store stage re-gathers the reference for fine orientations whose slice did not fit the cache the fine pass fills
    (k_store<SLICED=false>); with the default 4 GiB budget that path never runs, so force it: no cache at all, and a cache of
    five slices (both kernels contribute to the same accumulators)."""
    wl = make_workload(ori_size=32, healpix_order=2, n_particles=8, nr_classes=1, seed=130, snr=0.2, local_search=True)
    npf = wl.model.current_size * (wl.model.current_size // 2 + 1)
    monkeypatch.setenv("RB_SLICE_CACHE_BYTES", str(cache_slices * npf * 8))
    monkeypatch.setenv("RB_BAND_ROUNDS", "100000")          # band-major path: as many rounds of a few slices as the pool needs
    res, _ = _compare_pool(device, oracle, wl)
    assert res.particles["n_fine_orient"].sum() > cache_slices


def test_two_device_bundles_in_one_process(device):
    """RELION drives several GPUs from the threads of one process (one MlDeviceBundle per device): contexts on different
    devices must not share per-device state (kernel attributes, streams, buffers).  On a box with one GPU the second bundle
    lives on the same device: two independent contexts (own streams, buffers, models) interleaved in one process."""
    import torch
    from relion_b200.estep import MlDeviceBundle
    wl = make_workload(ori_size=64, healpix_order=2, n_particles=12, nr_classes=1, seed=71, snr=0.2, local_search=True)
    wg = make_workload(ori_size=32, healpix_order=1, n_particles=8, nr_classes=2, seed=72, snr=0.3)
    second = MlDeviceBundle(1 if torch.cuda.device_count() >= 2 else 0)
    try:
        out = {}
        for name, dev, w in (("local0", device, wl), ("local1", second, wl), ("global1", second, wg), ("global0", device, wg)):
            _setup(dev, w)
            out[name] = (dev.expectation_some_particles(w.pool), [dev.bp_get(k) for k in range(w.model.nr_classes)])
        for a, b in (("local0", "local1"), ("global0", "global1")):
            ra, rb = out[a][0].particles, out[b][0].particles
            for f in ("best_ihidden_over", "nr_significant_coarse", "n_fine_samples"):
                assert np.array_equal(ra[f], rb[f]), (a, b, f)
            np.testing.assert_allclose(ra["dLL_nolog"], rb["dLL_nolog"], rtol=1e-6)
            for ka, kb in zip(out[a][1], out[b][1]):
                for x, y in zip(ka, kb):
                    assert np.abs(x - y).max() <= 1e-5 * max(np.abs(y).max(), 1e-12)
    finally:
        second.close()


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_halfmap_fsc_against_reference(device, kind):
    """north_star: the reconstructed half-maps have FSC >= 0.995 against the reference's maps at every shell to
    Nyquist.  Two random half-sets go through the CUDA E-step and through the CPU oracle (RELION's own ALTCPU kernels
    when oracle/_ref is there); each accumulator pair is reconstructed by the same restated BackProjector::reconstruct
    (oracle/reconstruct.py) and compared shell by shell."""
    from oracle.bindings import Oracle, Projector, Backprojector, have_reference
    from oracle import reconstruct as rc
    if kind == "reference" and not have_reference():
        pytest.skip("oracle/_ref/librefkernels.so not built")
    orc = Oracle(kind)
    for half in range(2):
        wl = make_workload(ori_size=32, healpix_order=1, n_particles=150, nr_classes=1, seed=40 + half, snr=0.5)
        _setup(device, wl)
        device.expectation_some_particles(wl.pool)
        gre, gim, gw = device.bp_get(0)
        refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
        bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
        st, _, _ = orc.estep_pool(wl.model, wl.sampling, refs, bps, wl.pool, num_threads=0, exact_threshold=(kind == "port"))
        assert st == 0
        ours = rc.reconstruct(gre, gim, gw, wl.model.ori_size, wl.r_max, wl.padding_factor)
        want = rc.reconstruct(bps[0].real, bps[0].imag, bps[0].weight, wl.model.ori_size, wl.r_max, wl.padding_factor)
        f = rc.fsc(ours, want)
        assert f.shape[0] == wl.model.ori_size // 2 + 1
        assert f.min() >= 0.995, (half, f)
