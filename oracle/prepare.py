"""CPU restatement (numpy, float64) of the image preparation that precedes the E-step kernels.

TEST INFRASTRUCTURE ONLY (the checker for rb_pool_prepare, SURVEY.md §8f "next" row 1).

Follows getFourierTransformsAndCtfs for 2D images, one body, no helix / tomo / beam tilt / MTF, zero-masking
(/root/reference/src/acc/acc_ml_optimiser_impl.h:11-1010):
  old offsets rounded                                   :216  (my_old_offset.selfROUND)
  TranslateAndNormCorrect                               src/acc/utilities_impl.h:374-436, cpu_translate2D helper.cpp:256-282
  normalizeAndTransformImage                            src/acc/utilities_impl.h:438-486
      runCenterFFT(forward = false)                     src/acc/acc_helper_functions.h:519-552, centerFFT_2D src/fftw.h:406-437
      forward FFT, scaled by 1 / (n*n)
      windowFourierTransform2, shrinking branch         src/acc/acc_helper_functions_impl.h:2235-2330, helper.cuh:954-1000
  Fimg_nomask = transform of the unmasked image         :520-538
  softMaskBackgroundValue + cosineFilter (zero mask)    helper.cpp:117-253, call :610-680
  powerClass (spectrum of the full-size masked transform, highres_Xi2 beyond the current size)   helper.h:468-540, call :705-772
  CTF::getFftwImage on the current-size window          src/ctf.h:184-256 via relion_b200.synth.CTF (known answer tests/ctf.cpp)
Parity: unpinned by reference tests (the reference holds no vectors for image preparation).
"""
from __future__ import annotations

import math

import numpy as np

from relion_b200 import synth


def translate_and_norm(img: np.ndarray, dx: int, dy: int, norm: float) -> np.ndarray:
    """out[y+dy, x+dx] = norm * img[y, x] where the target is inside the box; everything else stays zero."""
    n = img.shape[0]
    out = np.zeros_like(img, dtype=np.float64)
    ys, xs = np.mgrid[0:n, 0:n]
    yp, xp = ys + dy, xs + dx
    ok = (yp >= 0) & (xp >= 0) & (yp < n) & (xp < n)
    out[yp[ok], xp[ok]] = img[ys[ok], xs[ok]] * norm
    return out


def normalize_and_transform(img: np.ndarray, current_size: int):
    """(current-size windowed transform, full-size transform) of a real-space image, RELION conventions."""
    n = img.shape[0]
    centred = np.roll(img, (-(n // 2), -(n // 2)), axis=(0, 1))          # runCenterFFT(forward=false): shift by -n/2
    F = np.fft.rfft2(centred) / float(n * n)
    return synth.window_ft(F, current_size), F


def soft_mask(img: np.ndarray, radius: float, cosine_width: float):
    """Zero-masking with the background value of the soft edge (softMaskBackgroundValue + cosineFilter)."""
    n = img.shape[0]
    if radius < 0:
        radius = n / 2.0
    radius_p = radius + cosine_width
    c = np.arange(n) - n // 2
    y, x = np.meshgrid(c, c, indexing="ij")
    r = np.sqrt((x * x + y * y).astype(np.float64))
    rc = np.where(r > radius_p, 1.0, np.where(r < radius, 0.0, 0.5 + 0.5 * np.cos((radius_p - r) / cosine_width * np.pi)))
    outside = r >= radius
    s, sbg = rc[outside].sum(), (rc * img)[outside].sum()
    bg = sbg / s
    return np.where(r < radius, img, img * (1.0 - rc) + bg * rc), bg


def noise_mask(img: np.ndarray, noise: np.ndarray, radius: float, cosine_width: float):
    """Noise-filled soft mask (RELION's default, !do_zero_mask): cosineFilter with the noise image as the fill value
    (/root/reference/src/acc/cpu/cpu_kernels/helper.cpp:228-247, acc_ml_optimiser_impl.h:660-668)."""
    n = img.shape[0]
    if radius < 0:
        radius = n / 2.0
    radius_p = radius + cosine_width
    c = np.arange(n) - n // 2
    y, x = np.meshgrid(c, c, indexing="ij")
    r = np.sqrt((x * x + y * y).astype(np.float64))
    rc = np.where(r > radius_p, 1.0, np.where(r < radius, 0.0, 0.5 + 0.5 * np.cos((radius_p - r) / cosine_width * np.pi)))
    return np.where(r < radius, img, img * (1.0 - rc) + noise * rc)


def noise_image_shell_power(n: int, spectrum: np.ndarray):
    """Expected mean |F|^2 per shell of the 1/N-normalised transform of a noise image made like makeNoiseImage
    (/root/reference/src/acc/utilities_impl.h:231-371): independent complex normals (variance spectrum[ires]^2 per component) on
    the half transform, zero beyond the last shell.  Columns 0 < x < n/2 keep their value (2 spectrum^2); the self-conjugate
    columns x = 0 and x = n/2 are symmetrised by the inverse real transform."""
    xf = n // 2 + 1
    iy = np.arange(n); y = np.where(iy >= xf, iy - n, iy)
    x = np.arange(xf)
    ires = np.rint(np.sqrt((x[None, :] ** 2 + y[:, None] ** 2).astype(np.float64))).astype(int)
    scale = np.where(ires < min(xf, spectrum.shape[0]), spectrum[np.minimum(ires, spectrum.shape[0] - 1)], 0.0)
    return ires, 2.0 * scale ** 2


def power_class(F_full: np.ndarray, current_size: int):
    """(spectrum [n/2+1], highres_Xi2): |F|^2 per shell of the full-size transform; Xi2 = power in shells >= current/2+1."""
    n = F_full.shape[0]
    xdim = n // 2 + 1
    iy = np.arange(n)
    y = np.where(iy < xdim, iy, iy - n)[:, None]
    x = np.arange(xdim)[None, :]
    ires = np.floor(np.sqrt((x * x + y * y).astype(np.float64)) + 0.5).astype(np.int64)
    ok = (ires < xdim) & ~((x == 0) & (y < 0))
    p = np.abs(F_full) ** 2
    spectrum = np.bincount(ires[ok], weights=p[ok], minlength=xdim)[:xdim]
    xi2 = p[ok & (ires >= current_size // 2 + 1)].sum()
    return spectrum, float(xi2)


def prepare_particle(img, old_offset, norm_factor, ctf: synth.CTF | None, ori_size, current_size, pixel_size,
                     mask_radius, width_mask_edge):
    """One particle: returns dict(Fimg, Fimg_nomask, Fctf, highres_Xi2, power_img, old_offset_rounded)."""
    rnd = lambda v: int(v + 0.5) if v > 0 else int(v - 0.5)                               # ROUND (src/macros.h:197), :216
    dx, dy = rnd(old_offset[0]), rnd(old_offset[1])
    t = translate_and_norm(np.asarray(img, np.float64), dx, dy, norm_factor)
    F_nomask, _ = normalize_and_transform(t, current_size)
    masked, _ = soft_mask(t, mask_radius, width_mask_edge)
    Fimg, F_full = normalize_and_transform(masked, current_size)
    if current_size < ori_size:
        spectrum, xi2 = power_class(F_full, current_size)
    else:
        spectrum, xi2 = np.zeros(ori_size // 2 + 1), 0.0
    Fctf = ctf.fftw_image(current_size, ori_size, pixel_size) if ctf is not None else np.ones(Fimg.shape)
    return dict(Fimg=Fimg, Fimg_nomask=F_nomask, Fctf=Fctf, highres_Xi2=xi2, power_img=spectrum, old_offset=(float(dx), float(dy)))


def prepared_pool(wl, raw):
    """The ParticlePool this restatement makes from a RawParticlePool (tests, bench cpu leg), plus power_img [P, n/2+1]."""
    from relion_b200.estep import ParticlePool
    P = raw.n_particles
    cs, n = wl.model.current_size, wl.model.ori_size
    xs = cs // 2 + 1
    Fimg = np.empty((P, cs, xs), np.complex64); F0 = np.empty_like(Fimg); Fctf = np.empty((P, cs, xs), np.float32)
    xi2 = np.zeros(P); old = np.zeros((P, 2)); power = np.zeros((P, n // 2 + 1))
    for p in range(P):
        ctf = synth.CTF(raw.ctf_defU[p], raw.ctf_defV[p], raw.ctf_defAngle[p], kV=raw.og_kV[raw.optics_group[p]],
                        Cs=raw.og_Cs[raw.optics_group[p]], Q0=raw.og_Q0[raw.optics_group[p]])
        r = prepare_particle(raw.images[p], raw.old_offset[p], raw.norm_factor[p], ctf, n, cs, wl.model.pixel_size,
                                  raw.mask_radius, raw.width_mask_edge)
        Fimg[p], F0[p], Fctf[p], xi2[p], old[p], power[p] = r["Fimg"], r["Fimg_nomask"], r["Fctf"], r["highres_Xi2"], r["old_offset"], r["power_img"]
    pool = ParticlePool(Fimg=Fimg, Fimg_nomask=F0, Fctf=Fctf, group_id=raw.group_id, optics_group=raw.optics_group, highres_Xi2=xi2,
                        old_offset=old, prior_offset=raw.prior_offset, dir_off=raw.dir_off, dir_idx=raw.dir_idx, dir_prior=raw.dir_prior,
                        psi_off=raw.psi_off, psi_idx=raw.psi_idx, psi_prior=raw.psi_prior)
    return pool, power
