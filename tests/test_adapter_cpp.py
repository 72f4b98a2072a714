"""The C++ host adapter (include/relion_b200_adapter.hpp): AccProjector / AccBackprojector / MlDeviceBundle(MlOptimiser *) /
MlOptimiserCuda(MlOptimiser *, MlDeviceBundle *, const char *) with the reference's names, constructor signatures and call
sequence (src/acc/cuda/cuda_ml_optimiser.h:18-144), the host bookkeeping of storeWeightedSums in C++ and the NCCL twin of
MlOptimiserMpi::combineAllWeightedSums.

 * compile check against the REAL reference headers (tests/cpp/relion_callsites.cpp, CPU, needs /root/reference);
 * tests/cpp/estep_multi_gpu (C++, no Python in the loop) on the mock MlOptimiser of tests/cpp/mock_relion: one-GPU E-step
   against the same particles split over two ranks + combineAllWeightedSums (NCCL when the box has two GPUs, host sum
   otherwise), then its metadata rows and weighted sums against the Python host mirror.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from relion_b200.workload import make_workload, raw_pool_from

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "estep_multi_gpu")

# exp_metadata columns (src/ml_optimiser.h:51-80)
ROT, TILT, PSI, XOFF, YOFF, ZOFF, CLASS, DLL, PMAX, NR_SIGN, NORM = range(11)
DEFU, DEFV, DEFANG, BFAC, KFAC, PHASE = range(11, 17)
NCOL = 25


def _exe():
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    return EXE


def _dump(path, arrays):
    """name -> array file read by tests/cpp/estep_multi_gpu.cpp (dtype codes 0 float64, 1 float32, 2 int32)."""
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            code = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int32): 2}[a.dtype]
            nb = name.encode()
            f.write(struct.pack("<i", len(nb))); f.write(nb)
            f.write(struct.pack("<ii", code, a.ndim)); f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(a.tobytes())


def _workload_arrays(wl, raw, avg_norm=1.0):
    s, m = wl.sampling, wl.model
    N = raw.n_particles
    sc = lambda v: np.array([float(v)], np.float64)
    md = np.zeros((N, NCOL), np.float64)
    md[:, 17:25] = 999.0                                                       # no priors given
    md[:, ROT], md[:, TILT], md[:, PSI] = wl.truth["rot"], wl.truth["tilt"], wl.truth["psi"]
    md[:, XOFF:YOFF + 1] = raw.old_offset
    md[:, NORM] = avg_norm / raw.norm_factor
    md[:, DEFU], md[:, DEFV], md[:, DEFANG] = raw.ctf_defU, raw.ctf_defV, raw.ctf_defAngle
    md[:, KFAC] = 1.0
    ps = float(m.pixel_size)
    refs = np.stack([np.ascontiguousarray(v, np.complex128) for v in wl.refs])
    a = {
        "nr_classes": sc(m.nr_classes), "ori_size": sc(m.ori_size), "pixel_size": sc(ps), "nr_groups": sc(len(m.scale_correction)),
        "sigma2_offset": sc(m.sigma2_offset), "avg_norm_correction": sc(avg_norm), "local_search": sc(raw.dir_off is not None),
        "sigma2_ang": sc(4.0), "r_max": sc(wl.r_max), "healpix_order": sc(s.healpix_order), "n_over_rot": sc(s.n_over_rot),
        "n_over_trans": sc(s.n_over_trans), "coarse_size": sc(m.coarse_size), "current_size": sc(m.current_size),
        "adaptive_fraction": sc(m.adaptive_fraction), "particle_diameter": sc(2.0 * raw.mask_radius * ps), "width_mask_edge": sc(raw.width_mask_edge),
        "refs": refs.view(np.float64).reshape(refs.shape + (2,)),
        "sigma2_noise": np.asarray(m.sigma2_noise, np.float64).reshape(-1), "scale_correction": np.asarray(m.scale_correction, np.float64),
        "pdf_class": np.asarray(m.pdf_class, np.float64), "data_vs_prior_class": np.asarray(m.data_vs_prior_class, np.float64),
        "rot": np.asarray(s.rot, np.float64), "tilt": np.asarray(s.tilt, np.float64), "psi": np.asarray(s.psi, np.float64),
        "over_rot": np.asarray(s.over_rot, np.float64), "over_tilt": np.asarray(s.over_tilt, np.float64), "over_psi": np.asarray(s.over_psi, np.float64),
        "trans_x": np.asarray(s.trans_x, np.float64) * ps, "trans_y": np.asarray(s.trans_y, np.float64) * ps,
        "over_trans_x": np.asarray(s.over_trans_x, np.float64) * ps, "over_trans_y": np.asarray(s.over_trans_y, np.float64) * ps,
        "metadata": md, "images": np.asarray(raw.images, np.float32), "group_id": np.asarray(raw.group_id, np.int32),
    }
    if raw.dir_off is not None:
        a.update(dir_off=np.asarray(raw.dir_off, np.int32), dir_idx=np.asarray(raw.dir_idx, np.int32), dir_prior=np.asarray(raw.dir_prior, np.float64),
                 psi_off=np.asarray(raw.psi_off, np.int32), psi_idx=np.asarray(raw.psi_idx, np.int32), psi_prior=np.asarray(raw.psi_prior, np.float64))
    return a, md


def test_adapter_compiles_against_the_reference_headers():
    """RELION's own call sites keep compiling: the adapter + the three accelerator call sites of src/ml_optimiser.cpp against the
    real src/ml_optimiser.h (declaration-only stubs for fftw3.h / tiffio.h / png.h, nothing linked)."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present (GPU box)")
    cpp = os.path.join(ROOT, "tests", "cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-I" + os.path.join(cpp, "relion_stubs"), "-I/root/reference",
                        "-I" + os.path.join(ROOT, "include"), os.path.join(cpp, "relion_callsites.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_adapter_without_gpu_throws_relion_error(tmp_path):
    """No CPU fallback: MlDeviceBundle::setDevice throws (the reference's HANDLE_ERROR -> REPORT_ERROR) when there is no device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    wl = make_workload(ori_size=16, healpix_order=1, n_particles=2, seed=3, nr_groups=1)
    raw = raw_pool_from(wl, seed=1)
    arrays, _ = _workload_arrays(wl, raw)
    _dump(str(tmp_path / "w.bin"), arrays)
    r = subprocess.run([_exe(), str(tmp_path / "w.bin"), str(tmp_path / "o.bin"), "1"], capture_output=True, text=True)
    assert r.returncode == 1
    assert "relion_b200" in r.stderr and "no CUDA device" in r.stderr and "relion_b200_adapter.hpp" in r.stderr, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(ori_size=32, healpix_order=1, n_particles=21, nr_classes=2, seed=51, snr=0.3, nr_groups=3),
                                dict(ori_size=40, current_size=28, healpix_order=2, n_particles=14, nr_classes=1, seed=52, snr=0.2, local_search=True, nr_groups=2)])
def test_cpp_estep_reduction_and_bookkeeping(device, tmp_path, kw):
    """tests/cpp/estep_multi_gpu: 1-rank run == 2-rank run + combineAllWeightedSums (inside the program), and its metadata rows /
    weighted sums == the Python host mirror on the same raw pool (same library underneath, bookkeeping done twice)."""
    import torch
    from relion_b200 import parallel
    wl = make_workload(**kw)
    raw = raw_pool_from(wl, seed=7)
    arrays, md0 = _workload_arrays(wl, raw, avg_norm=0.95)
    _dump(str(tmp_path / "w.bin"), arrays)
    r = subprocess.run([_exe(), str(tmp_path / "w.bin"), str(tmp_path / "o.bin"), "2", "8"], capture_output=True, text=True)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    if torch.cuda.device_count() >= 2:
        assert "NCCL" in r.stdout
    buf = open(tmp_path / "o.bin", "rb").read()
    nm = struct.unpack_from("<q", buf, 0)[0]
    md = np.frombuffer(buf, np.float64, nm, 8).reshape(-1, NCOL)
    npk = struct.unpack_from("<q", buf, 8 + 8 * nm)[0]
    pack = np.frombuffer(buf, np.float64, npk, 16 + 8 * nm)

    # the Python host mirror on the same pool
    m, s = wl.model, wl.sampling
    device.set_model(m); device.set_sampling(s)
    for k, v in enumerate(wl.refs):
        device.set_reference(k, v.astype(np.complex128), wl.r_max, wl.padding_factor)
        device.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)
    power = device.pool_prepare(0, raw)
    res = device.estep_slot(0)
    p = res.particles
    if raw.dir_off is not None:
        gd = np.array([raw.dir_idx[raw.dir_off[i] + p["best_idir"][i]] for i in range(len(p))])
        gp = np.array([raw.psi_idx[raw.psi_off[i] + p["best_ipsi"][i]] for i in range(len(p))])
    else:
        gd, gp = p["best_idir"], p["best_ipsi"]
    g = (gd * s.n_psi + gp) * s.n_over_rot + p["best_iover_rot"]
    np.testing.assert_array_equal(md[:, ROT], np.asarray(s.over_rot)[g])
    np.testing.assert_array_equal(md[:, TILT], np.asarray(s.over_tilt)[g])
    np.testing.assert_array_equal(md[:, PSI], np.asarray(s.over_psi)[g])
    it = p["best_itrans"] * s.n_over_trans + p["best_iover_trans"]
    ps = float(m.pixel_size)
    rnd = np.where(raw.old_offset > 0, np.floor(raw.old_offset + 0.5), -np.floor(-raw.old_offset + 0.5))
    np.testing.assert_allclose(md[:, XOFF], rnd[:, 0] + np.asarray(s.over_trans_x)[it] * ps / ps, rtol=0, atol=1e-12)
    np.testing.assert_allclose(md[:, YOFF], rnd[:, 1] + np.asarray(s.over_trans_y)[it] * ps / ps, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(md[:, CLASS], p["best_class"] + 1)
    np.testing.assert_array_equal(md[:, NR_SIGN], p["nr_significant_coarse"])
    np.testing.assert_allclose(md[:, PMAX], p["pmax"], rtol=1e-6)
    # dLL = log(sum_weight) - min_diff2 - logsigma2 (acc_ml_optimiser_impl.h:3556-3574)
    from relion_b200.synth import mresol
    ir = mresol(m.current_size)
    sig = np.asarray(m.sigma2_noise, np.float64).reshape(-1)
    logsigma2 = np.log(2 * np.pi * sig[ir[ir > 0]]).sum()
    np.testing.assert_allclose(md[:, DLL], p["dLL_nolog"] - logsigma2, rtol=1e-6)
    # norm correction (:3505-3538): sqrt(2 * (wsum_norm + power beyond the current size)) times the old value / avg
    hi = power[:, m.current_size // 2 + 1:].astype(np.float64).sum(axis=1)
    np.testing.assert_allclose(md[:, NORM], (md0[:, NORM] / 0.95) * np.sqrt(2.0 * (p["wsum_norm_correction"] + hi)), rtol=2e-4)
    # packed weighted sums: MlWsumModel::pack order, LL first, then ave_Pmax, sigma2_offset, avg_norm_correction
    np.testing.assert_allclose(pack[0], (p["dLL_nolog"] - logsigma2).sum(), rtol=1e-6)
    np.testing.assert_allclose(pack[1], p["pmax"].astype(np.float64).sum(), rtol=1e-6)
    np.testing.assert_allclose(pack[2], p["wsum_sigma2_offset"].sum(), rtol=1e-5)
    np.testing.assert_allclose(pack[3], md[:, NORM].sum(), rtol=1e-9)
    nshell = m.ori_size // 2 + 1
    s2 = res.wsum_sigma2_noise.astype(np.float64).sum(axis=0)
    s2[m.current_size // 2 + 1:] += power[:, m.current_size // 2 + 1:].astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(pack[7:7 + nshell], s2, rtol=2e-4, atol=1e-9 * s2.max())


@pytest.mark.gpu
@pytest.mark.parametrize("img_box,model_box", [(40, 32), (32, 40)])
def test_cpp_adapter_optics_group_with_its_own_box(device, tmp_path, img_box, model_box):
    """An optics group whose box differs from the model's (same pixel size) through the C++ adapter: MlDeviceBundle::setGeometry loads
    the group's image geometry (its box as the library's ori_size, sigma2_noise gathered at ROUND(remap * ires), rb_model.ref_max_r when
    the references end inside the window), the pool carries mat_left = applyScaleDifference(I), the weighted sigma2 sums are
    scattered back onto the model's shells (acc_ml_optimiser_impl.h:3615-3625) and logsigma2 reads the remapped shells (:3556-3566).
    Checked against the Python host mirror, which is given the same image-geometry model directly."""
    from relion_b200.synth import mresol
    wl = make_workload(ori_size=img_box, ref_box=model_box, healpix_order=1, n_particles=13, nr_classes=1, seed=61, snr=0.3, nr_groups=2)
    m, s = wl.model, wl.sampling
    remap = model_box / img_box                                       # (ori * pixel) / (my_image_size * my_pixel_size)
    ns_img, ns_mod = img_box // 2 + 1, model_box // 2 + 1
    rnd = lambda v: np.floor(np.asarray(v, np.float64) + 0.5).astype(np.int64)
    # the model's own spectrum (on its shells) and what an image of this group sees of it
    sig_img0 = np.asarray(m.sigma2_noise, np.float64).reshape(-1)
    sig_mod = np.interp(np.arange(ns_mod) / remap, np.arange(ns_img), sig_img0)
    gather = rnd(remap * np.arange(ns_img))
    assert gather.max() < ns_mod
    m.sigma2_noise = sig_mod[gather][None, :]
    dvp_mod = np.zeros((1, ns_mod)); dvp_mod[:, : min(ns_mod, ns_img)] = np.asarray(m.data_vs_prior_class)[:, : min(ns_mod, ns_img)]
    dvp_img = np.zeros((1, ns_img)); dvp_img[:, : min(ns_mod, ns_img)] = dvp_mod[:, : min(ns_mod, ns_img)]
    m.data_vs_prior_class = dvp_img
    raw = raw_pool_from(wl, seed=8)
    assert raw.mat_left is not None and abs(raw.mat_left[0, 0] - img_box / model_box) < 1e-12
    arrays, md0 = _workload_arrays(wl, raw, avg_norm=0.95)
    arrays["model_ori_size"] = np.array([float(model_box)])
    arrays["sigma2_noise"] = sig_mod
    arrays["data_vs_prior_class"] = dvp_mod
    _dump(str(tmp_path / "w.bin"), arrays)
    r = subprocess.run([_exe(), str(tmp_path / "w.bin"), str(tmp_path / "o.bin"), "2", "5"], capture_output=True, text=True)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    buf = open(tmp_path / "o.bin", "rb").read()
    nm = struct.unpack_from("<q", buf, 0)[0]
    md = np.frombuffer(buf, np.float64, nm, 8).reshape(-1, NCOL)
    npk = struct.unpack_from("<q", buf, 8 + 8 * nm)[0]
    pack = np.frombuffer(buf, np.float64, npk, 16 + 8 * nm)

    assert m.ref_max_r == (wl.r_max if wl.r_max < img_box // 2 else 0)
    device.set_model(m); device.set_sampling(s)
    device.set_reference(0, wl.refs[0].astype(np.complex128), wl.r_max, wl.padding_factor)
    device.bp_init(0, wl.bp_shape, wl.r_max, wl.padding_factor)
    power = device.pool_prepare(0, raw)
    res = device.estep_slot(0)
    p = res.particles
    g = (p["best_idir"] * s.n_psi + p["best_ipsi"]) * s.n_over_rot + p["best_iover_rot"]
    np.testing.assert_array_equal(md[:, ROT], np.asarray(s.over_rot)[g])
    np.testing.assert_array_equal(md[:, TILT], np.asarray(s.over_tilt)[g])
    np.testing.assert_array_equal(md[:, PSI], np.asarray(s.over_psi)[g])
    np.testing.assert_array_equal(md[:, CLASS], p["best_class"] + 1)
    np.testing.assert_array_equal(md[:, NR_SIGN], p["nr_significant_coarse"])
    np.testing.assert_allclose(md[:, PMAX], p["pmax"], rtol=1e-6)
    ir = mresol(m.current_size)
    logsigma2 = np.log(2 * np.pi * sig_mod[rnd(remap * ir[ir > 0])]).sum()
    np.testing.assert_allclose(md[:, DLL], p["dLL_nolog"] - logsigma2, rtol=1e-6)
    np.testing.assert_allclose(pack[0], (p["dLL_nolog"] - logsigma2).sum(), rtol=1e-6)
    # wsum_model.sigma2_noise lives on the MODEL's shells: image shell i goes to ROUND(i * remap)
    s2_img = res.wsum_sigma2_noise.astype(np.float64).sum(axis=0)
    s2_img[m.current_size // 2 + 1:] += power[:, m.current_size // 2 + 1:].astype(np.float64).sum(axis=0)
    s2_mod = np.zeros(ns_mod)
    for i in range(ns_img):
        if gather[i] < ns_mod:
            s2_mod[gather[i]] += s2_img[i]
    np.testing.assert_allclose(pack[7:7 + ns_mod], s2_mod, rtol=2e-4, atol=1e-9 * s2_mod.max())
    # the scale matters: the same pool without it lands somewhere else
    raw.mat_left = None
    device.pool_prepare(0, raw)
    other = device.estep_slot(0)
    assert np.abs(other.particles["dLL_nolog"] - p["dLL_nolog"]).max() > 0.1


@pytest.mark.gpu
def test_cpp_adapter_skip_align(device, tmp_path):
    """--skip_align through the C++ adapter: RELION refills the sampling object with the pool's own orientations and fractional
    offsets before every pool (src/ml_optimiser.cpp:4180-4225); the adapter reloads the sampling tables per pool, hands every particle
    the one-entry lists of its row, passes its translation as rb_particles.pre_shift and looks the metadata row up by the row index
    (acc_ml_optimiser_impl.h:3752-3766, 2874-2876).  Pools of 5 particles so that rows restart; checked against the Python mirror."""
    from relion_b200.workload import make_skip_align_workload
    from relion_b200.synth import mresol
    wl = make_skip_align_workload(n_particles=13, nr_classes=3, seed=71, snr=0.3, nr_groups=2)
    m, s = wl.model, wl.sampling
    raw = raw_pool_from(wl, seed=9)
    rnd = np.where(raw.old_offset > 0, np.floor(raw.old_offset + 0.5), -np.floor(-raw.old_offset + 0.5))
    raw.pre_shift = raw.old_offset - rnd
    arrays, md0 = _workload_arrays(wl, raw, avg_norm=0.95)
    arrays["skip_align"] = np.array([1.0]); arrays["local_search"] = np.array([0.0])
    for k in ("dir_off", "dir_idx", "dir_prior", "psi_off", "psi_idx", "psi_prior"):
        arrays.pop(k, None)
    _dump(str(tmp_path / "w.bin"), arrays)
    r = subprocess.run([_exe(), str(tmp_path / "w.bin"), str(tmp_path / "o.bin"), "2", "5"], capture_output=True, text=True)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    buf = open(tmp_path / "o.bin", "rb").read()
    nm = struct.unpack_from("<q", buf, 0)[0]
    md = np.frombuffer(buf, np.float64, nm, 8).reshape(-1, NCOL)
    npk = struct.unpack_from("<q", buf, 8 + 8 * nm)[0]
    pack = np.frombuffer(buf, np.float64, npk, 16 + 8 * nm)

    device.set_model(m); device.set_sampling(s)
    for k, v in enumerate(wl.refs):
        device.set_reference(k, v.astype(np.complex128), wl.r_max, wl.padding_factor)
        device.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)
    device.pool_prepare(0, raw)
    res = device.estep_slot(0)
    p = res.particles
    # the poses are the particles' own; the class is what the E-step decides
    np.testing.assert_array_equal(md[:, ROT], wl.truth["rot"])
    np.testing.assert_array_equal(md[:, PSI], wl.truth["psi"])
    np.testing.assert_allclose(md[:, XOFF], raw.old_offset[:, 0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(md[:, YOFF], raw.old_offset[:, 1], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(md[:, CLASS], p["best_class"] + 1)
    assert np.mean(p["best_class"] == wl.truth["cls"]) >= 0.8
    np.testing.assert_allclose(md[:, PMAX], p["pmax"], rtol=1e-6)
    ir = mresol(m.current_size)
    sig = np.asarray(m.sigma2_noise, np.float64).reshape(-1)
    logsigma2 = np.log(2 * np.pi * sig[ir[ir > 0]]).sum()
    np.testing.assert_allclose(md[:, DLL], p["dLL_nolog"] - logsigma2, rtol=1e-6)
    np.testing.assert_allclose(pack[0], (p["dLL_nolog"] - logsigma2).sum(), rtol=1e-6)
    np.testing.assert_allclose(pack[2], p["wsum_sigma2_offset"].sum(), rtol=1e-5)
