"""Seeded E-step workloads (references + sampling + model state + particle pool).

Used by tests/, __graft_entry__.smoke() and bench.py so that the CUDA path, the CPU oracle and the
reference arm all see byte-identical inputs.  Pure host code (numpy).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Optional

import numpy as np

from . import sampling as smp
from . import synth
from .estep import ModelParams, ParticlePool


@dataclasses.dataclass
class Workload:
    name: str
    model: ModelParams
    sampling: smp.Sampling
    refs: list                 # list of complex64 [Z, Y, X] padded Fourier volumes, one per class
    r_max: int
    padding_factor: float
    pool: ParticlePool
    truth: dict                # true class / angles / shifts of every particle
    bp_shape: tuple


def coarse_size_for(ori_size: int, pixel_size: float, angular_step: float, particle_diameter: float, current_size: int) -> int:
    """image_coarse_size for adaptive_oversampling > 0 (src/ml_optimiser.cpp:5757-5767), 3D reference."""
    rotated_distance = (angular_step / 360.0) * math.pi * particle_diameter
    coarse_resolution = rotated_distance / 1.2
    c = 2 * int(math.ceil(pixel_size * ori_size / coarse_resolution))
    return max(2, min(current_size, c))


def make_workload(name: str = "tiny", *, ori_size: int = 32, current_size: Optional[int] = None,
                  healpix_order: int = 1, offset_range: float = 3.0, offset_step: float = 2.0,
                  n_particles: int = 8, nr_classes: int = 1, snr: float = 0.5, seed: int = 1993,
                  local_search: bool = False, sigma_ang: Optional[float] = None,
                  pixel_size: float = 2.0, particle_diameter: Optional[float] = None,
                  nr_groups: int = 2, adaptive_fraction: float = 0.999, coarse_size: Optional[int] = None,
                  projector: Optional[Callable] = None, n_blobs: int = 40, refs_override=None,
                  ref_seed: int = 1993, ref_dim: int = 3, psi_step: float = 6.0, do_cc: bool = False,
                  mat_left=None, mat_right=None, ref_box: Optional[int] = None) -> Workload:
    """Build a complete, seeded E-step problem.

    projector(vol_complex64, r_max, pf, eulers[n,9] float32, n) -> [n_img, n, n//2+1] complex: how noise-free
    slices are made (default: the float64 numpy projector; bench.py passes the CUDA projector for 256 px).
    """
    rng = np.random.default_rng(seed)
    current_size = current_size or ori_size
    pf = 2.0
    # ref_box: the references live in their own box (an optics group whose box `ori_size` differs from the model's `ref_box`, same
    # pixel size): mat_left then defaults to ObservationModel::applyScaleDifference(I) = (image box * pixel) / (model box * pixel),
    # so that image pixel i reads the reference at i * ref_box / ori_size
    ref_size = ref_box or ori_size
    if ref_box is not None and mat_left is None:
        mat_left = np.eye(3) * (ori_size / ref_box)
    # ---- references -------------------------------------------------------------------------
    if refs_override is not None:
        refs, r_max = refs_override
    else:
        refs = []
        r_max = None
        for k in range(nr_classes):
            if ref_dim == 2:
                img = synth.make_phantom_2d(ori_size, n_blobs=n_blobs, seed=ref_seed + 17 * k)
                data, r_max = synth.reference_ft_2d(img, current_size=current_size, padding_factor=pf)
            else:
                vol = synth.make_phantom(ref_size, n_blobs=n_blobs, seed=ref_seed + 17 * k)
                data, r_max = synth.reference_ft(vol, current_size=min(current_size, ref_size) if ref_box is None else ref_size,
                                                 padding_factor=pf)
            refs.append(data.astype(np.complex64))
    # ---- sampling ---------------------------------------------------------------------------
    if ref_dim == 2:
        assert not local_search, "2D classification searches psi globally"
        s = smp.make_sampling_2d(psi_step, offset_range, offset_step, oversampling=1)
        ang_step = s.psi_step
    else:
        s = smp.make_sampling(healpix_order, offset_range, offset_step, oversampling=1)
        ang_step = smp.angular_sampling(healpix_order)
    diameter = particle_diameter or 0.7 * ori_size * pixel_size
    if coarse_size is None:
        coarse_size = coarse_size_for(ori_size, pixel_size, ang_step, diameter, current_size)
    # ---- true poses ---------------------------------------------------------------------------
    P = n_particles
    cls = rng.integers(0, nr_classes, P)
    # true orientations: a random fine-grid orientation (so that the maximum is well defined)
    idir = rng.integers(0, s.n_dir, P)
    ipsi = rng.integers(0, s.n_psi, P)
    io = rng.integers(0, s.n_over_rot, P)
    g = (idir * s.n_psi + ipsi) * s.n_over_rot + io
    rot, tilt, psi = s.over_rot[g], s.over_tilt[g], s.over_psi[g]
    it = rng.integers(0, s.n_trans * s.n_over_trans, P)
    shifts = np.stack([s.over_trans_x[it], s.over_trans_y[it]], axis=1)
    eul = synth.inverse_euler_f32(rot, tilt, psi)
    if mat_left is not None or mat_right is not None:          # inverse(L A R) (generateEulerMatrices, acc_helper_functions_impl.h:248-255)
        L = np.eye(3) if mat_left is None else np.asarray(mat_left, np.float64).reshape(3, 3)
        R = np.eye(3) if mat_right is None else np.asarray(mat_right, np.float64).reshape(3, 3)
        eul = np.stack([np.linalg.inv(L @ e.reshape(3, 3).astype(np.float64).T @ R).reshape(9) for e in eul]).astype(np.float32)
    # ---- noise-free slices ----------------------------------------------------------------------
    n = current_size
    if projector is None:
        slices = np.empty((P, n, n // 2 + 1), np.complex128)
        for p in range(P):
            A_inv = eul[p].reshape(3, 3).astype(np.float64)
            slices[p] = synth.project_numpy(refs[cls[p]], r_max, pf, A_inv, n)
    else:
        slices = np.empty((P, n, n // 2 + 1), np.complex64)
        for k in range(nr_classes):
            sel = np.nonzero(cls == k)[0]
            if len(sel):
                slices[sel] = projector(k, eul[sel], n)
    parts = synth.make_particles(slices, ori_size, pixel_size, snr, seed + 1, rot, tilt, psi, shifts)
    # ---- model state -----------------------------------------------------------------------------
    nshell = ori_size // 2 + 1
    sigma2 = np.tile(parts.sigma2_noise[None, :], (1, 1))
    scale = 1.0 + 0.05 * rng.standard_normal(nr_groups)
    pdf_class = np.full(nr_classes, 1.0 / nr_classes)
    pdf_direction = np.full((nr_classes, s.n_dir), 1.0 / s.n_dir)
    dvp = np.zeros((nr_classes, nshell))
    dvp[:, : max(2, nshell // 2)] = 10.0
    model = ModelParams(nr_classes=nr_classes, ori_size=ori_size, coarse_size=coarse_size, current_size=current_size,
                        pixel_size=pixel_size, sigma2_noise=sigma2, scale_correction=scale, pdf_class=pdf_class,
                        pdf_direction=None if local_search else pdf_direction, data_vs_prior_class=dvp,
                        sigma2_offset=(offset_range * pixel_size / 1.5) ** 2, adaptive_fraction=adaptive_fraction,
                        do_cc=do_cc, ref_max_r=int(r_max) if r_max < current_size // 2 else 0)
    # ---- pool ------------------------------------------------------------------------------------
    group = rng.integers(0, nr_groups, P).astype(np.int32)
    pool = ParticlePool(Fimg=parts.Fimg, Fimg_nomask=parts.Fimg_nomask, Fctf=parts.Fctf, group_id=group,
                        optics_group=np.zeros(P, np.int32), highres_Xi2=parts.highres_Xi2,
                        old_offset=np.zeros((P, 2)), prior_offset=np.zeros((P, 2)),
                        mat_left=None if mat_left is None else np.asarray(mat_left, np.float64).reshape(3, 3),
                        mat_right=None if mat_right is None else np.asarray(mat_right, np.float64).reshape(3, 3))
    if local_search:
        sig = sigma_ang if sigma_ang is not None else 2.0 * ang_step / 2.0   # 2 x oversampled step (src/ml_optimiser.cpp:2316-2326)
        doff, poff = [0], [0]
        di, dp, pi_, pp = [], [], [], []
        for p in range(P):
            # prior centre: the true orientation perturbed by ~half a coarse step
            pr = rot[p] + rng.normal(0, 0.3 * ang_step)
            pt = float(np.clip(tilt[p] + rng.normal(0, 0.3 * ang_step), 0.0, 180.0))
            pq = (psi[p] + rng.normal(0, 0.3 * ang_step)) % 360.0
            a, b, c, d = smp.select_nonzero_prior(s, pr, pt, pq, sig, sig, sig)
            di.append(a); dp.append(b); pi_.append(c); pp.append(d)
            doff.append(doff[-1] + len(a)); poff.append(poff[-1] + len(c))
        pool.dir_off = np.array(doff, np.int32); pool.psi_off = np.array(poff, np.int32)
        pool.dir_idx = np.concatenate(di).astype(np.int32); pool.dir_prior = np.concatenate(dp)
        pool.psi_idx = np.concatenate(pi_).astype(np.int32); pool.psi_prior = np.concatenate(pp)
    pad = refs[0].shape[0]
    bp_shape = (pad, pad // 2 + 1) if refs[0].ndim == 2 else (pad, pad, pad // 2 + 1)
    truth = dict(cls=cls, rot=rot, tilt=tilt, psi=psi, shifts=shifts, idir=idir, ipsi=ipsi, iover_rot=io, itrans_over=it,
                 ctf_params=parts.ctf_params)
    return Workload(name, model, s, refs, r_max, pf, pool, truth, bp_shape)


def raw_pool_from(wl: Workload, seed: int = 0, mask_radius: Optional[float] = None, width_mask_edge: float = 3.0,
                  max_old_offset: float = 2.4):
    """A RawParticlePool (real-space images + CTF parameters + metadata) for the image-preparation path: the images are the
    inverse transforms of the workload's unmasked Fourier particles, the old offsets random non-integers (they get rounded and
    applied as integer shifts), the norm factors close to 1."""
    from .estep import RawParticlePool
    rng = np.random.default_rng(seed)
    P = wl.pool.n_particles
    F0 = np.asarray(wl.pool.Fimg_nomask.numpy() if hasattr(wl.pool.Fimg_nomask, "numpy") else wl.pool.Fimg_nomask)
    F0 = F0.view(np.complex64).reshape(P, wl.model.current_size, wl.model.current_size // 2 + 1)
    images = synth.raw_images_from_ft(F0, wl.model.ori_size)
    cp = wl.truth["ctf_params"]
    raw = RawParticlePool(images=images, old_offset=rng.uniform(-max_old_offset, max_old_offset, (P, 2)), prior_offset=np.zeros((P, 2)),
                          group_id=wl.pool.group_id, optics_group=wl.pool.optics_group,
                          ctf_defU=cp[:, 0].copy(), ctf_defV=cp[:, 1].copy(), ctf_defAngle=cp[:, 2].copy(),
                          og_kV=np.array([300.0]), og_Cs=np.array([2.7]), og_Q0=np.array([0.1]),
                          norm_factor=rng.uniform(0.9, 1.1, P),
                          mask_radius=(0.42 * wl.model.ori_size if mask_radius is None else mask_radius), width_mask_edge=width_mask_edge,
                          dir_off=wl.pool.dir_off, dir_idx=wl.pool.dir_idx, dir_prior=wl.pool.dir_prior,
                          psi_off=wl.pool.psi_off, psi_idx=wl.pool.psi_idx, psi_prior=wl.pool.psi_prior,
                          mat_left=wl.pool.mat_left, mat_right=wl.pool.mat_right)
    return raw



def make_skip_align_workload(*, ori_size: int = 32, n_particles: int = 12, nr_classes: int = 3, seed: int = 7, snr: float = 0.3,
                             skip_rotate_only: bool = False, offset_range: float = 3.0, offset_step: float = 1.0,
                             pixel_size: float = 2.0, nr_groups: int = 2, n_blobs: int = 40) -> Workload:
    """--skip_align (only classify; skip_rotate_only: --skip_rotate, translations still searched): the sampling tables hold the
    POOL's orientations, particle p uses entry p of the direction and psi tables through one-entry lists
    (MlOptimiser::expectationSomeParticles, src/ml_optimiser.cpp:4180-4225: sampling.addOneOrientation / addOneTranslation per
    particle; acc_ml_optimiser_impl.h:3752-3766: idir = ipsi = itrans = row of the particle), no oversampling
    (src/ml_optimiser.cpp:2382-2389), the orientation prior is pdf_class.  With --skip_align the particle's own fractional offset is
    its only translation: here rb_particles.pre_shift, the sampling holds (0, 0)."""
    rng = np.random.default_rng(seed)
    pf = 2.0
    P = n_particles
    refs = []
    for k in range(nr_classes):
        vol = synth.make_phantom(ori_size, n_blobs=n_blobs, seed=seed + 17 * k)
        data, r_max = synth.reference_ft(vol, current_size=ori_size, padding_factor=pf)
        refs.append(data.astype(np.complex64))
    rot, tilt, psi = rng.uniform(-180, 180, P), rng.uniform(0, 180, P), rng.uniform(-180, 180, P)
    if skip_rotate_only:
        base = smp.make_sampling(1, offset_range, offset_step, oversampling=0, build_oversampled=False)
        tx, ty = np.asarray(base.trans_x, np.float64), np.asarray(base.trans_y, np.float64)
        it = rng.integers(0, len(tx), P)
        shifts = np.stack([tx[it], ty[it]], axis=1)
        pre_shift = None
        old_offset = np.zeros((P, 2))
    else:
        tx, ty = np.zeros(1), np.zeros(1)
        it = np.zeros(P, np.int64)
        shifts = rng.uniform(-0.5, 0.5, (P, 2))              # the fractional part of the old offset
        pre_shift = shifts.copy()
        old_offset = np.zeros((P, 2))                        # the rounded part, already applied to the image
    s = smp.Sampling(healpix_order=1, psi_step=0.0, offset_range=offset_range, offset_step=offset_step, oversampling=0,
                     rot=rot.copy(), tilt=tilt.copy(), psi=psi.copy(), trans_x=tx, trans_y=ty)
    # the "oversampled" tables of oversampling order 0 are the coarse ones: entry (d, q) = (rot[d], tilt[d], psi[q])
    s.over_rot = np.repeat(rot, P); s.over_tilt = np.repeat(tilt, P); s.over_psi = np.tile(psi, P)
    s.over_trans_x, s.over_trans_y = tx.copy(), ty.copy()
    cls = rng.integers(0, nr_classes, P)
    eul = synth.inverse_euler_f32(rot, tilt, psi)
    n = ori_size
    slices = np.empty((P, n, n // 2 + 1), np.complex128)
    for p in range(P):
        slices[p] = synth.project_numpy(refs[cls[p]], r_max, pf, eul[p].reshape(3, 3).astype(np.float64), n)
    parts = synth.make_particles(slices, ori_size, pixel_size, snr, seed + 1, rot, tilt, psi, shifts)
    nshell = ori_size // 2 + 1
    pdf_class = rng.uniform(0.5, 1.5, nr_classes); pdf_class /= pdf_class.sum()
    dvp = np.zeros((nr_classes, nshell)); dvp[:, : max(2, nshell // 2)] = 10.0
    model = ModelParams(nr_classes=nr_classes, ori_size=ori_size, coarse_size=ori_size, current_size=ori_size, pixel_size=pixel_size,
                        sigma2_noise=np.tile(parts.sigma2_noise[None, :], (1, 1)), scale_correction=1.0 + 0.05 * rng.standard_normal(nr_groups),
                        pdf_class=pdf_class, pdf_direction=None, data_vs_prior_class=dvp,
                        sigma2_offset=(offset_range * pixel_size / 1.5) ** 2, do_skip_rotate=True)
    one = np.arange(P + 1, dtype=np.int32)
    pool = ParticlePool(Fimg=parts.Fimg, Fimg_nomask=parts.Fimg_nomask, Fctf=parts.Fctf, group_id=rng.integers(0, nr_groups, P).astype(np.int32),
                        optics_group=np.zeros(P, np.int32), highres_Xi2=parts.highres_Xi2, old_offset=old_offset, prior_offset=np.zeros((P, 2)),
                        dir_off=one, dir_idx=np.arange(P, dtype=np.int32), dir_prior=np.ones(P),
                        psi_off=one.copy(), psi_idx=np.arange(P, dtype=np.int32), psi_prior=np.ones(P), pre_shift=pre_shift)
    pad = refs[0].shape[0]
    truth = dict(cls=cls, rot=rot, tilt=tilt, psi=psi, shifts=shifts, itrans=it, ctf_params=parts.ctf_params)
    return Workload("skip_align", model, s, refs, r_max, pf, pool, truth, (pad, pad, pad // 2 + 1))
