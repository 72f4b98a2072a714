"""CPU restatement (numpy, float64) of RELION's posed back-projection, the path relion_reconstruct runs per particle.

TEST INFRASTRUCTURE ONLY (the checker for rb_backproject_posed / BASELINE config #2).

Follows /root/reference/src/backprojector.cpp:55-357 (BackProjector::backproject2Dto3D, TRILINEAR branch, no Ewald
sphere, no magnification matrix) as called by Reconstructor::backprojectOneParticle (src/reconstructor.cpp:328-744:
F2D already multiplied by the CTF, weight image Fctf = ctf^2, DC component zeroed at :716).
Parity: unpinned by reference tests (the reference holds no vectors for reconstruct).
"""
from __future__ import annotations

import math

import numpy as np


def backproject2Dto3D(data: np.ndarray, weight: np.ndarray, f2d: np.ndarray, A_inv: np.ndarray, Mweight: np.ndarray,
                      r_max: int, padding_factor: float = 2.0) -> None:
    """Adds one image into data (complex128 [Z, Y, X]) / weight (float64 [Z, Y, X]) in place.

    A_inv: the INVERTED 3x3 orientation matrix (what Ainv = A.inv() is at :76-77, before the padding scale).
    """
    Ainv = np.asarray(A_inv, np.float64) * padding_factor
    rr = int(math.floor(r_max * padding_factor + 0.5))
    max_r2 = rr * rr
    zdim, ydim, xdim = data.shape
    starty, startz = -((ydim - 1) // 2), -((zdim - 1) // 2)          # STARTINGY / STARTINGZ of the centred array
    s, sh = f2d.shape
    AtA_xx = float((Ainv[:, 0] ** 2).sum()); AtA_xy = float((Ainv[:, 0] * Ainv[:, 1]).sum()); AtA_yy = float((Ainv[:, 1] ** 2).sum())
    for i in range(s):
        if i < sh:
            y, first_allowed_x = i, 0
        else:
            y, first_allowed_x = i - s, 1
        discr = AtA_xy * AtA_xy * y * y - AtA_xx * (AtA_yy * y * y - max_r2)                 # :118-128
        if discr < 0.0:
            continue
        d = math.sqrt(discr) / AtA_xx
        q = -AtA_xy * y / AtA_xx
        first_x = max(int(math.ceil(q - d)), first_allowed_x)
        last_x = min(int(math.floor(q + d)), sh - 1)
        if last_x < first_x:
            continue
        x = np.arange(first_x, last_x + 1, dtype=np.float64)
        val = f2d[i, first_x:last_x + 1].astype(np.complex128)
        w = Mweight[i, first_x:last_x + 1].astype(np.float64)
        xp = Ainv[0, 0] * x + Ainv[0, 1] * y
        yp = Ainv[1, 0] * x + Ainv[1, 1] * y
        zp = Ainv[2, 0] * x + Ainv[2, 1] * y
        ok = (w > 0.0) & (xp * xp + yp * yp + zp * zp <= max_r2)
        neg = xp < 0
        xp = np.where(neg, -xp, xp); yp = np.where(neg, -yp, yp); zp = np.where(neg, -zp, zp)
        val = np.where(neg, np.conj(val), val)
        x0 = np.floor(xp).astype(np.int64); fx = xp - x0
        y0 = np.floor(yp).astype(np.int64); fy = yp - y0; y0 -= starty
        z0 = np.floor(zp).astype(np.int64); fz = zp - z0; z0 -= startz
        ok &= (x0 >= 0) & (x0 + 1 < xdim) & (y0 >= 0) & (y0 + 1 < ydim) & (z0 >= 0) & (z0 + 1 < zdim)   # :213-218
        if not ok.any():
            continue
        x0, y0, z0, fx, fy, fz, val, w = (a[ok] for a in (x0, y0, z0, fx, fy, fz, val, w))
        for dz, wz in ((0, 1.0 - fz), (1, fz)):
            for dy, wy in ((0, 1.0 - fy), (1, fy)):
                for dx, wx in ((0, 1.0 - fx), (1, fx)):
                    dd = wz * wy * wx
                    np.add.at(data, (z0 + dz, y0 + dy, x0 + dx), dd * val)
                    np.add.at(weight, (z0 + dz, y0 + dy, x0 + dx), dd * w)


def prepare_particle(img: np.ndarray, shift, ctf_image: np.ndarray | None, ctf_premultiplied: bool = False):
    """What Reconstructor::backprojectOneParticle does to one 2D particle before backproject2Dto3D
    (/root/reference/src/reconstructor.cpp:428-745, no Ewald sphere / FOM / subtraction): F2D = FFT(img) / N with
    CenterFFTbySign (src/fftw.h:390-403), shiftImageInFourierTransform (src/fftw.cpp:874-918: x e^{-2 pi i (x tx + y ty) / n}),
    F2D *= Fctf unless the data are premultiplied, Fctf = Fctf^2, F2D(0, 0) = 0.  Returns (F2D complex128, Fctf^2 float64)."""
    n = img.shape[0]
    xs = n // 2 + 1
    F = np.fft.rfft2(np.asarray(img, np.float64)) / float(n * n)
    iy = np.arange(n)[:, None]; x = np.arange(xs)[None, :]
    F = F * np.where(((iy ^ x) & 1) != 0, -1.0, 1.0)
    y = np.where(iy < xs, iy, iy - n)
    tx, ty = (0.0, 0.0) if shift is None else (float(shift[0]), float(shift[1]))
    if abs(tx / n) >= 1e-6 or abs(ty / n) >= 1e-6:
        F = F * np.exp(-2j * np.pi * (x * tx + y * ty) / n)
    c = np.ones((n, xs)) if ctf_image is None else np.asarray(ctf_image, np.float64)
    if not ctf_premultiplied:
        F = F * c
    F[0, 0] = 0.0
    return F, c * c
