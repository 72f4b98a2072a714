"""CPU restatement (numpy, float64) of the reconstruction that turns back-projection accumulators into a map.

TEST INFRASTRUCTURE ONLY (the checker): nothing under relion_b200/ imports this.  It exists so that the parity
tests can apply north_star's criterion "the reconstructed half-maps have FSC >= 0.995 against the reference's maps
at every shell to Nyquist" to the accumulators the CUDA path and the CPU oracle produce from the same particles.

Follows, in /root/reference:
  BackProjector::reconstruct, default `skip_gridding` branch        src/backprojector.cpp:1379-1575
      (skip_gridding is the default: src/ml_optimiser.cpp:586 `--dont_skip_gridding` turns it off)
  Projector::decenter                                                src/projector.h:249-260
  MAP regularisation term (1/tau2 added to the weights)              src/backprojector.cpp:1463-1507
  BackProjector::windowToOridimRealSpace                             src/backprojector.cpp:2530-2665
  windowFourierTransform (shrinking branch)                          src/fftw.h:850-856
  CenterFFTbySign                                                    src/fftw.h:390-403
  softMaskOutsideMap (radius = xsize/2, cosine_width = 3)            src/mask.cpp:43-96
  Projector::griddingCorrect (TRILINEAR: divide by sinc^2)           src/projector.cpp:595-628
  getFSC                                                             src/fftw.cpp:481-513
Parity: unpinned by reference tests (the reference has none for reconstruct); both maps of an FSC comparison go
through this same code, so the comparison measures the accumulators, not this restatement.
"""
from __future__ import annotations

import math

import numpy as np


def _fftw_freq(n: int) -> np.ndarray:
    """FOR_ALL_ELEMENTS_IN_FFTW_TRANSFORM: kp = k < n/2+1 ? k : k - n."""
    k = np.arange(n)
    return np.where(k < n // 2 + 1, k, k - n)


def decenter(centered: np.ndarray, max_r2: int) -> np.ndarray:
    """Projector-centred [pad, pad, pad//2+1] (y,z origin at (pad-1)//2) -> FFTW order, zero beyond max_r2."""
    pad = centered.shape[0]
    h = (pad - 1) // 2
    f = _fftw_freq(pad)
    kz, ky, kx = np.meshgrid(f, f, np.arange(pad // 2 + 1), indexing="ij")
    out = centered[kz + h, ky + h, kx]
    return np.where(kz * kz + ky * ky + kx * kx <= max_r2, out, 0)


def soft_mask_outside_map(vol: np.ndarray, cosine_width: float = 3.0) -> np.ndarray:
    n = vol.shape[0]
    c = np.arange(n) - n // 2
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    r = np.sqrt((x * x + y * y + z * z).astype(np.float64))
    radius = n / 2.0
    radius_p = radius + cosine_width
    rc = np.where(r > radius_p, 1.0, np.where(r < radius, 0.0, 0.5 + 0.5 * np.cos(np.pi * (radius_p - r) / cosine_width)))
    bg = float((rc * vol).sum() / rc.sum())
    return (1.0 - rc) * vol + rc * bg


def gridding_correct(vol: np.ndarray, ori_size: int, padding_factor: float) -> np.ndarray:
    n = vol.shape[0]
    c = np.arange(n) - n // 2
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    rval = np.sqrt((x * x + y * y + z * z).astype(np.float64)) / (ori_size * padding_factor)
    sinc = np.ones_like(rval)
    nz = rval > 0
    sinc[nz] = np.sin(np.pi * rval[nz]) / (np.pi * rval[nz])
    return vol / (sinc * sinc)


def helical_operators(nr_helical_asu: int, helical_twist: float, helical_rise: float, ori_size: int, padding_factor: float):
    """The operators of applyHelicalSymmetry (src/backprojector.cpp:2186-2197, :2284-2288): for hh in [-n/2, n/2 + n%2) without 0
    the rotation about Z by hh * -twist degrees (rotation3DMatrix 'Z', src/transformations.cpp:111-117, entries below 1e-6 zeroed
    by setSmallValuesToZero) and the phase ramp zshift = hh * rise / (-ori_size * padding_factor) in cycles per voxel of z."""
    if nr_helical_asu < 2:
        return np.zeros((0, 3, 3)), np.zeros(0)
    h_min = -(nr_helical_asu // 2) if nr_helical_asu >= 0 else 0
    h_max = -h_min + nr_helical_asu % 2
    Rs, zs = [], []
    for hh in range(h_min, h_max):
        if hh == 0:
            continue
        a = math.radians(hh * -helical_twist)
        c, s = math.cos(a), math.sin(a)
        c = 0.0 if abs(c) < 1e-6 else c
        s = 0.0 if abs(s) < 1e-6 else s
        Rs.append([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        zs.append(hh * helical_rise / (-ori_size * padding_factor) if abs(helical_rise) > 0 else 0.0)
    return np.array(Rs, np.float64).reshape(-1, 3, 3), np.array(zs, np.float64)


def symmetrise(real: np.ndarray, imag: np.ndarray, weight: np.ndarray, r_max: int, padding_factor: float, rotations=(), helical=None):
    """BackProjector::symmetrise (src/backprojector.cpp:2136-2480): enforceHermitianSymmetry, applyHelicalSymmetry (when
    helical = (nr_helical_asu, twist in degrees, rise in pixels, ori_size) is given) and applyPointGroupSymmetry on centred
    [Z, Y, X] arrays; returns new (real, imag, weight) float64."""
    data = real.astype(np.float64) + 1j * imag.astype(np.float64)
    w = weight.astype(np.float64).copy()
    Z, Y, X = data.shape
    hz, hy = (Z - 1) // 2, (Y - 1) // 2
    # enforceHermitianSymmetry (:2148-2165)
    for iz in range(-hz, Z - hz):
        for iy in range(0 if iz < 0 else 1, Y - hy):
            if -iz + hz >= Z or -iy + hy < 0:
                continue
            a, b = (iz + hz, iy + hy, 0), (-iz + hz, -iy + hy, 0)
            fs = data[a] + np.conj(data[b])
            data[a], data[b] = fs, np.conj(fs)
            sw = w[a] + w[b]
            w[a] = w[b] = sw
    passes = []
    if helical is not None:
        passes.append(helical_operators(helical[0], helical[1], helical[2], helical[3], padding_factor))
    rotations = np.asarray(rotations, np.float64).reshape(-1, 3, 3)
    passes.append((rotations, np.zeros(len(rotations))))
    for rotations, zshifts in passes:
        if not len(rotations):
            continue
        rr = int(math.floor(r_max * padding_factor + 0.5))
        kz, ky, kx = np.meshgrid(np.arange(Z) - hz, np.arange(Y) - hy, np.arange(X), indexing="ij")
        inside = (kx * kx + ky * ky + kz * kz) <= rr * rr
        sum_d, sum_w = data.copy(), w.copy()
        x, y, z = kx[inside].astype(np.float64), ky[inside].astype(np.float64), kz[inside].astype(np.float64)
        for R, zshift in zip(rotations, zshifts):
            xp = x * R[0, 0] + y * R[0, 1] + z * R[0, 2]
            yp = x * R[1, 0] + y * R[1, 1] + z * R[1, 2]
            zp = x * R[2, 0] + y * R[2, 1] + z * R[2, 2]
            neg = xp < 0
            xp = np.where(neg, -xp, xp); yp = np.where(neg, -yp, yp); zp = np.where(neg, -zp, zp)
            x0 = np.floor(xp).astype(int); fx = xp - x0
            y0 = np.floor(yp).astype(int); fy = yp - y0; y0 += hy
            z0 = np.floor(zp).astype(int); fz = zp - z0; z0 += hz
            vd = np.zeros(x.shape, np.complex128); vw = np.zeros(x.shape)
            for dz, wz in ((0, 1 - fz), (1, fz)):
                for dy, wy in ((0, 1 - fy), (1, fy)):
                    for dx, wx in ((0, 1 - fx), (1, fx)):
                        vd += data[z0 + dz, y0 + dy, x0 + dx] * (wz * wy * wx)
                        vw += w[z0 + dz, y0 + dy, x0 + dx] * (wz * wy * wx)
            vd = np.where(neg, np.conj(vd), vd)
            if zshift != 0.0:                                                     # phase ramp of the helical rise (:2284-2296)
                vd = vd * np.exp(2j * np.pi * z * zshift)
            sum_d[inside] += vd
            sum_w[inside] += vw
        data, w = sum_d, sum_w
    return data.real, data.imag, w


def ftblob_ratio(rval: np.ndarray, blob_radius: float = 1.9, alpha: float = 15.0, nr_elem: int = 10000) -> np.ndarray:
    """tab_ftblob(rval) / tab_ftblob(0) of BackProjector (src/backprojector.h:115: TabFtBlob::initialise(blob_radius * 2, alpha,
    order 0, 10000), src/tabfuncs.cpp:95-121: nearest-below table of kaiser_Fourier_value, zero beyond 0.5;
    src/funcs.cpp:244-251 with bessi1_5(x) = sqrt(2 / (pi x)) (cosh x - sinh x / x), numerical_recipes.cpp:366-369)."""
    a = blob_radius * 2.0
    sampling = 0.5 / nr_elem

    def value(w):
        arg = 2.0 * np.pi * a * w
        sigma = np.sqrt(np.abs(alpha * alpha - arg * arg))
        sigma = np.where(sigma == 0, 1e-300, sigma)
        i15 = np.sqrt(2.0 / (np.pi * sigma)) * (np.cosh(sigma) - np.sinh(sigma) / sigma)
        j15 = np.sqrt(2.0 / (np.pi * sigma)) * (np.sin(sigma) / sigma - np.cos(sigma))
        return np.where(arg > alpha, j15, i15) / sigma ** 1.5

    idx = (np.abs(rval) / sampling).astype(np.int64)
    tab = value(np.arange(nr_elem) * sampling)
    return np.where(idx >= nr_elem, 0.0, tab[np.minimum(idx, nr_elem - 1)]) / tab[0]


def reconstruct(real: np.ndarray, imag: np.ndarray, weight: np.ndarray, ori_size: int, r_max: int,
                padding_factor: float = 2.0, tau2: np.ndarray | None = None, tau2_fudge: float = 1.0,
                minres_map: int = 0, max_iter_preweight: int = 0) -> np.ndarray:
    """BackProjector::reconstruct for a 3D reference built from 2D images: [ori, ori, ori] float64.  max_iter_preweight = 0:
    the default skip_gridding branch; > 0: the iterative gridding of Pipe & Menon (--dont_skip_gridding,
    /root/reference/src/backprojector.cpp:1577-1700, convoluteBlobRealSpace :2483-2528)."""
    data = real.astype(np.float64) + 1j * imag.astype(np.float64)
    pad = data.shape[0]
    rr = int(math.floor(r_max * padding_factor + 0.5))
    max_r2 = rr * rr
    f = _fftw_freq(pad)
    kz, ky, kx = np.meshgrid(f, f, np.arange(pad // 2 + 1), indexing="ij")
    r2 = kz * kz + ky * ky + kx * kx
    Fweight = decenter(weight.astype(np.float64), max_r2)
    if tau2 is not None:                                                           # :1463-1507
        oversampling_correction = padding_factor ** 3
        ires = np.floor(np.sqrt(r2.astype(np.float64)) / padding_factor + 0.5).astype(np.int64)
        ires_c = np.minimum(ires, len(tau2) - 1)
        t = np.asarray(tau2, np.float64)[ires_c]
        invtau2 = np.where(t > 0, 1.0 / (oversampling_correction * tau2_fudge * np.where(t > 0, t, 1.0)),
                           np.where(Fweight > 1e-20, 1.0 / (0.001 * np.where(Fweight > 1e-20, Fweight, 1.0)), 0.0))
        Fweight = np.where((r2 < max_r2) & (ires >= minres_map), Fweight + invtau2, Fweight)
    Fconv = decenter(data, max_r2)
    if max_iter_preweight > 0:
        # Fnewweight starts as 1 inside the sphere; every iteration divides it by |blob-convolution of (Fnewweight Fweight)|
        Fnew = (r2 < max_r2).astype(np.float64)
        kp = np.arange(pad); kp = np.where(kp < pad // 2, kp, kp - pad)          # padhdim = pad_size / 2 (:2487, :2512-2514)
        kk, ii, jj = np.meshgrid(kp, kp, kp, indexing="ij")
        blob = ftblob_ratio(np.sqrt((kk * kk + ii * ii + jj * jj).astype(np.float64)) / (ori_size * padding_factor))
        n3 = float(pad) ** 3
        for _ in range(max_iter_preweight):
            M = np.fft.irfftn(Fnew * Fweight, s=(pad,) * 3, axes=(0, 1, 2)) * n3           # unnormalised inverse transform
            conv = np.fft.rfftn(M * blob, axes=(0, 1, 2)) / n3                              # forward transform divides by N
            Fnew = np.where(r2 < max_r2, Fnew / np.maximum(1e-6, np.abs(conv)), Fnew)     # Eq. [14] of Pipe & Menon (:1633-1648)
        Fconv = Fconv * Fnew
    else:
        # radial average of the weights / 1000 as the floor of the divisor (:1513-1573)
        round_max_r2 = int(math.floor(r_max * padding_factor * r_max * padding_factor + 0.5))
        iresf = np.floor(np.sqrt(r2.astype(np.float64)) / padding_factor).astype(np.int64)
        inside = r2 < round_max_r2
        radavg = np.bincount(iresf[inside], weights=Fweight[inside], minlength=r_max)[:r_max]
        counter = np.bincount(iresf[inside], minlength=r_max)[:r_max].astype(np.float64)
        radavg = radavg / (1000.0 * np.maximum(counter, 1.0))
        w = np.maximum(Fweight, radavg[np.minimum(iresf, r_max - 1)])
        Fconv = np.where(w != 0, Fconv / np.where(w != 0, w, 1.0), Fconv)

    # windowToOridimRealSpace
    padoridim = int(math.floor(padding_factor * ori_size + 0.5))
    padoridim += padoridim % 2
    fo = _fftw_freq(padoridim)
    xo = padoridim // 2 + 1
    xin = pad // 2 + 1
    if xo > xin:                                               # enlarging branch: zero-pad
        Fin = np.zeros((padoridim, padoridim, xo), np.complex128)
        Fin[np.ix_(f % padoridim, f % padoridim, np.arange(xin))] = Fconv
    else:
        Fin = Fconv[np.ix_(fo % pad, fo % pad, np.arange(xo))]
    k = np.arange(padoridim)
    sign = 1 - 2 * ((k[:, None, None] ^ k[None, :, None] ^ np.arange(xo)[None, None, :]) & 1)
    Fin = Fin * sign                                           # CenterFFTbySign
    M = np.fft.irfftn(Fin, s=(padoridim,) * 3, axes=(0, 1, 2)) * float(padoridim) ** 3   # RELION's inverse transform is unnormalised
    o = padoridim // 2 - ori_size // 2
    M = M[o:o + ori_size, o:o + ori_size, o:o + ori_size]
    M = M / (padding_factor ** 3 * ori_size)                   # normfft, ref_dim 3 / data_dim 2
    M = soft_mask_outside_map(M)
    return gridding_correct(M, ori_size, padding_factor)


def fsc(map1: np.ndarray, map2: np.ndarray) -> np.ndarray:
    """getFSC: shells 0..n/2 (fsc[i] = sum conj(z1) z2 / sqrt(sum|z1|^2 sum|z2|^2) over round(|k|) == i)."""
    n = map1.shape[0]
    F1, F2 = np.fft.rfftn(map1), np.fft.rfftn(map2)
    f = _fftw_freq(n)
    kz, ky, kx = np.meshgrid(f, f, np.arange(n // 2 + 1), indexing="ij")
    idx = np.floor(np.sqrt((kz * kz + ky * ky + kx * kx).astype(np.float64)) + 0.5).astype(np.int64)
    ok = idx < n // 2 + 1
    num = np.bincount(idx[ok], weights=(np.conj(F1) * F2).real[ok], minlength=n // 2 + 1)
    d1 = np.bincount(idx[ok], weights=(np.abs(F1) ** 2)[ok], minlength=n // 2 + 1)
    d2 = np.bincount(idx[ok], weights=(np.abs(F2) ** 2)[ok], minlength=n // 2 + 1)
    return num / np.sqrt(np.maximum(d1 * d2, 1e-300))


def update_ssnr(weight: np.ndarray, ori_size: int, r_max: int, padding_factor: float, tau2_fudge: float, tau2: np.ndarray,
                fsc: np.ndarray | None = None, avgctf2: np.ndarray | None = None, update_tau2_with_fsc: bool = False,
                is_whole_instead_of_half: bool = False):
    """BackProjector::updateSSNRarrays (src/backprojector.cpp:1041-1204) on a centred weight array [Z, Y, X] (3D) or
    [Y, X] (2D): returns (tau2, sigma2, data_vs_prior, fourier_coverage), each [ori_size/2 + 1] float64."""
    w = np.asarray(weight, np.float64)
    ns = ori_size // 2 + 1
    rr = int(math.floor(r_max * padding_factor + 0.5))
    max_r2 = rr * rr
    if w.ndim == 3:
        Z, Y, X = w.shape
        kz, ky, kx = np.meshgrid(np.arange(Z) - (Z - 1) // 2, np.arange(Y) - (Y - 1) // 2, np.arange(X), indexing="ij")
        r2 = kz * kz + ky * ky + kx * kx
        oc = padding_factor ** 3
    else:
        Y, X = w.shape
        ky, kx = np.meshgrid(np.arange(Y) - (Y - 1) // 2, np.arange(X), indexing="ij")
        r2 = ky * ky + kx * kx
        oc = padding_factor ** 2
    inside = r2 < max_r2
    ires = np.floor(np.sqrt(r2[inside].astype(np.float64)) / padding_factor + 0.5).astype(np.int64)      # ROUND
    wi = w[inside]
    s = np.bincount(ires, weights=oc * wi, minlength=ns)[:ns]
    counter = np.bincount(ires, minlength=ns)[:ns].astype(np.float64)
    if np.any((s > 0) & (s <= 1e-20)):
        raise ValueError("unexpectedly small, yet non-zero sigma2 value")
    sigma2 = np.where(s > 1e-20, counter / np.where(s > 1e-20, s, 1.0), 0.0)
    tau2 = np.array(tau2, np.float64, copy=True)
    dvp = np.zeros(ns)
    if update_tau2_with_fsc:
        f = np.maximum(0.001, np.asarray(fsc, np.float64))
        if is_whole_instead_of_half:
            f = np.sqrt(2.0 * f / (f + 1.0))
        f = np.minimum(0.999, f)
        ssnr = f / (1.0 - f) * tau2_fudge
        tau2 = ssnr * sigma2
        dvp = ssnr.copy()
    if np.any(tau2 < 0):
        raise ValueError("Negative values encountered for tau2 spectrum")
    t = tau2[ires]
    with np.errstate(divide="ignore", invalid="ignore"):
        invtau2 = np.where(t > 0, 1.0 / (oc * tau2_fudge * np.where(t > 0, t, 1.0)), 1.0 / (0.001 * wi))
        if avgctf2 is not None:
            a = np.asarray(avgctf2, np.float64)[ires]
            invtau2 = np.where((t > 0) & (a > 0), invtau2 / np.where(a > 0, a, 1.0), invtau2)
        ratio = wi / invtau2
    ratio = np.where(np.isnan(ratio), 0.0, ratio)              # weight 0 with tau2 0: 0 / inf
    cov = np.bincount(ires, weights=(ratio >= 1.0).astype(np.float64), minlength=ns)[:ns]
    if not update_tau2_with_fsc:
        dsum = np.bincount(ires, weights=ratio, minlength=ns)[:ns]
        idx = np.arange(ns)
        dvp = np.where(idx > r_max, 0.0, np.where(counter < 0.001, 999.0, dsum / np.maximum(counter, 1e-300)))
    cov = np.where(counter > 0, cov / np.maximum(counter, 1.0), cov)
    return tau2, sigma2, dvp, cov
