"""Host-side orientation / translation tables for the E-step harness.

In a RELION build these lists come from RELION's own `HealpixSampling`
(/root/reference/src/healpix_sampling.cpp) and are handed to the library through `rb_set_sampling`
(include/relion_b200.h).  This module restates just enough of that host logic to drive the tests and
bench.py without RELION:

* HEALPix NESTED pixel -> (z, phi)          Healpix_Base::pix2ang_z_phi / nest2xyf (vendored
                                            src/Healpix_2.15a; published algorithm, Gorski et al. 2005)
* coarse grid                               HealpixSampling::setOrientations      (healpix_sampling.cpp:477-560)
* oversampled orientations                  getOrientations / pushbackOversampledPsiAngles (:1832-1960)
* translations and their oversampling       setTranslations (:291-445), getTranslationsInPixel (:1724-1830)
* prior-selected local-search lists         selectOrientationsWithNonZeroPriorProbability (:695-1010),
                                            C1 symmetry, no bimodal/helical branches

Symmetry pruning, tilt limits, helices and random perturbation are out of scope.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])


def _compress_bits(v: np.ndarray) -> np.ndarray:
    """Keep the even bits of v (inverse of bit interleaving)."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def _spread_bits(v: np.ndarray) -> np.ndarray:
    v = v & 0x00000000FFFFFFFF
    v = (v | (v << 16)) & 0x0000FFFF0000FFFF
    v = (v | (v << 8)) & 0x00FF00FF00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v << 2)) & 0x3333333333333333
    v = (v | (v << 1)) & 0x5555555555555555
    return v


def nest2xyf(order: int, ipix: np.ndarray):
    ipix = np.asarray(ipix, dtype=np.int64)
    npface = 1 << (2 * order)
    face = ipix >> (2 * order)
    p = ipix & (npface - 1)
    return _compress_bits(p), _compress_bits(p >> 1), face


def xyf2nest(order: int, x, y, face):
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    return (np.asarray(face, dtype=np.int64) << (2 * order)) + _spread_bits(x) + (_spread_bits(y) << 1)


def xyf2ang(order: int, ix, iy, face):
    """(z, phi) of pixel (ix, iy, face) — Healpix_Base::pix2ang_z_phi, NEST branch."""
    nside = 1 << order
    nl4 = 4 * nside
    npix = 12 * nside * nside
    fact2 = 4.0 / npix
    fact1 = (nside << 1) * fact2
    ix = np.asarray(ix, dtype=np.int64)
    iy = np.asarray(iy, dtype=np.int64)
    face = np.asarray(face, dtype=np.int64)
    jr = _JRLL[face] * nside - ix - iy - 1
    north = jr < nside
    south = jr > 3 * nside
    nr = np.where(north, jr, np.where(south, nl4 - jr, nside))
    z = np.where(north, 1.0 - nr * nr * fact2, np.where(south, nr * nr * fact2 - 1.0, (2 * nside - jr) * fact1))
    kshift = np.where(north | south, 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > nl4, jp - nl4, jp)
    jp = np.where(jp < 1, jp + nl4, jp)
    phi = (jp - (kshift + 1) * 0.5) * ((math.pi / 2) / nr)
    return z, phi


def _direction_from_zphi(z, phi):
    """rot, tilt in degrees; rot wrapped to [-180, 180] (HealpixSampling::getDirectionFromHealPix + checkDirection)."""
    rot = np.degrees(phi)
    tilt = np.degrees(np.arccos(np.clip(z, -1.0, 1.0)))
    rot = np.where(rot > 180.0, rot - 360.0, rot)
    return rot, tilt


def angular_sampling(order: int) -> float:
    """HealpixSampling::getAngularSampling for 3D: 360 / (6 * nside)."""
    return 360.0 / (6 * (1 << order))


@dataclasses.dataclass
class Sampling:
    healpix_order: int
    psi_step: float
    offset_range: float          # pixels
    offset_step: float           # pixels
    oversampling: int = 1        # adaptive_oversampling
    rot: np.ndarray = None       # [n_dir]
    tilt: np.ndarray = None
    psi: np.ndarray = None       # [n_psi]
    over_rot: np.ndarray = None  # [n_dir*n_psi*n_over_rot]
    over_tilt: np.ndarray = None
    over_psi: np.ndarray = None
    trans_x: np.ndarray = None   # [n_trans] pixels
    trans_y: np.ndarray = None
    over_trans_x: np.ndarray = None  # [n_trans*n_over_trans]
    over_trans_y: np.ndarray = None
    is_3d: bool = True           # False: 2D references, psi-only sampling (HealpixSampling with is_3D == false)

    @property
    def n_dir(self):
        return len(self.rot)

    @property
    def n_psi(self):
        return len(self.psi)

    @property
    def n_trans(self):
        return len(self.trans_x)

    @property
    def n_over_rot(self):
        # oversamplingFactorOrientations (src/healpix_sampling.cpp:1644-1651): 8^order in 3D, 2^order in 2D
        return (8 if self.is_3d else 2) ** self.oversampling

    @property
    def n_over_trans(self):
        return 4 ** self.oversampling


def make_sampling(healpix_order: int, offset_range: float, offset_step: float, oversampling: int = 1,
                  psi_step: float | None = None, build_oversampled: bool = True) -> Sampling:
    """Coarse grid + oversampled tables in the layout rb_set_sampling expects."""
    s = Sampling(healpix_order, psi_step or angular_sampling(healpix_order), offset_range, offset_step, oversampling)
    npix = 12 * (1 << (2 * healpix_order))
    ipix = np.arange(npix, dtype=np.int64)
    x, y, f = nest2xyf(healpix_order, ipix)
    z, phi = xyf2ang(healpix_order, x, y, f)
    s.rot, s.tilt = _direction_from_zphi(z, phi)
    nr_psi = int(math.ceil(360.0 / s.psi_step))
    s.psi_step = 360.0 / nr_psi
    s.psi = np.arange(nr_psi, dtype=np.float64) * s.psi_step

    # translations (setTranslations, 2D branch)
    maxp = int(math.ceil(offset_range / offset_step))
    tx, ty = [], []
    for ix in range(-maxp, maxp + 1):
        for iy in range(-maxp, maxp + 1):
            xo, yo = ix * offset_step, iy * offset_step
            if xo * xo + yo * yo < offset_range * offset_range + 0.001:
                tx.append(xo)
                ty.append(yo)
    s.trans_x = np.array(tx, dtype=np.float64)
    s.trans_y = np.array(ty, dtype=np.float64)
    nov = 2 ** oversampling
    if oversampling == 0:
        s.over_trans_x, s.over_trans_y = s.trans_x.copy(), s.trans_y.copy()
    else:
        sub = -0.5 * offset_step + (0.5 + np.arange(nov)) * offset_step / nov
        ox = s.trans_x[:, None, None] + sub[None, :, None] + 0 * sub[None, None, :]
        oy = s.trans_y[:, None, None] + 0 * sub[None, :, None] + sub[None, None, :]
        s.over_trans_x = ox.reshape(-1).copy()
        s.over_trans_y = oy.reshape(-1).copy()

    if build_oversampled:
        if oversampling == 0:
            s.over_rot = np.repeat(s.rot, nr_psi)
            s.over_tilt = np.repeat(s.tilt, nr_psi)
            s.over_psi = np.tile(s.psi, npix)
        else:
            # getOrientations: the fact x fact children of the coarse pixel on the fine NESTED grid, row (j) major,
            # each followed by its oversampled psi values
            fo = healpix_order + oversampling
            fact = 1 << oversampling
            jj, ii = np.meshgrid(np.arange(fact), np.arange(fact), indexing="ij")
            cx = (fact * x[:, None, None] + ii[None]).reshape(npix, -1)
            cy = (fact * y[:, None, None] + jj[None]).reshape(npix, -1)
            cz, cphi = xyf2ang(fo, cx, cy, f[:, None])
            crot, ctilt = _direction_from_zphi(cz, cphi)                       # [npix, fact*fact]
            opsi = s.psi[:, None] - 0.5 * s.psi_step + (0.5 + np.arange(nov))[None, :] * s.psi_step / nov  # [n_psi, nov]
            nd = fact * fact
            s.over_rot = np.broadcast_to(crot[:, None, :, None], (npix, nr_psi, nd, nov)).reshape(-1).copy()
            s.over_tilt = np.broadcast_to(ctilt[:, None, :, None], (npix, nr_psi, nd, nov)).reshape(-1).copy()
            s.over_psi = np.broadcast_to(opsi[None, :, None, :], (npix, nr_psi, nd, nov)).reshape(-1).copy()
    return s


def make_sampling_2d(psi_step: float, offset_range: float, offset_step: float, oversampling: int = 1) -> Sampling:
    """Psi-only sampling of 2D classification (HealpixSampling::setOrientations / getOrientations, 2D branches,
    src/healpix_sampling.cpp:471-480, 1832-1870): one direction (rot = tilt = 0), n_over_rot = 2^oversampling psi values."""
    s3 = make_sampling(0, offset_range, offset_step, oversampling, psi_step=psi_step, build_oversampled=False)
    s = Sampling(-1, s3.psi_step, offset_range, offset_step, oversampling, is_3d=False)
    s.rot = np.zeros(1); s.tilt = np.zeros(1); s.psi = s3.psi
    s.trans_x, s.trans_y, s.over_trans_x, s.over_trans_y = s3.trans_x, s3.trans_y, s3.over_trans_x, s3.over_trans_y
    nov = 2 ** oversampling
    if oversampling == 0:
        s.over_psi = s.psi.copy()
    else:
        s.over_psi = (s.psi[:, None] - 0.5 * s.psi_step + (0.5 + np.arange(nov))[None, :] * s.psi_step / nov).reshape(-1).copy()
    s.over_rot = np.zeros_like(s.over_psi); s.over_tilt = np.zeros_like(s.over_psi)
    return s


def _direction(rot, tilt):
    """Euler_angles2direction (src/euler.cpp)."""
    a, b = np.radians(rot), np.radians(tilt)
    return np.stack([np.sin(b) * np.cos(a), np.sin(b) * np.sin(a), np.cos(b)], axis=-1)


def _gaussian1d(x, sigma):
    return np.exp(-0.5 * (x / sigma) ** 2) / math.sqrt(2 * math.pi * sigma * sigma)


def select_nonzero_prior(s: Sampling, prior_rot: float, prior_tilt: float, prior_psi: float,
                         sigma_rot: float, sigma_tilt: float, sigma_psi: float, sigma_cutoff: float = 3.0):
    """selectOrientationsWithNonZeroPriorProbability for C1, sigma_rot > 0 and sigma_tilt > 0.

    Returns (dir_idx int32, dir_prior float64, psi_idx int32, psi_prior float64).
    """
    pd = _direction(prior_rot, prior_tilt)
    dots = np.clip(_direction(s.rot, s.tilt) @ pd, -1.0, 1.0)
    diffang = np.degrees(np.arccos(dots))
    biggest = max(sigma_rot, sigma_tilt)
    sel = np.nonzero(diffang < sigma_cutoff * biggest)[0]
    if len(sel) == 0:
        sel = np.array([int(np.argmin(diffang))])
        dprior = np.array([1.0])
    else:
        dprior = _gaussian1d(diffang[sel], biggest)
        dprior = dprior / dprior.sum()
    diffpsi = np.abs(s.psi - prior_psi)
    diffpsi = np.where(diffpsi > 180.0, np.abs(diffpsi - 360.0), diffpsi)
    psel = np.nonzero(diffpsi < sigma_cutoff * sigma_psi)[0]
    if len(psel) == 0:
        psel = np.array([int(np.argmin(diffpsi))])
        pprior = np.array([1.0])
    else:
        pprior = _gaussian1d(diffpsi[psel], sigma_psi)
        pprior = pprior / pprior.sum()
    return sel.astype(np.int32), dprior.astype(np.float64), psel.astype(np.int32), pprior.astype(np.float64)


def euler_matrix(rot, tilt, psi):
    """Euler_angles2matrix (src/euler.cpp), ZYZ, degrees; returns A (not inverted)."""
    a, b, g = np.radians(rot), np.radians(tilt), np.radians(psi)
    ca, sa, cb, sb, cg, sg = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(g), np.sin(g)
    cc, cs, sc, ss = cb * ca, cb * sa, sb * ca, sb * sa
    A = np.empty(np.shape(a) + (3, 3))
    A[..., 0, 0] = cg * cc - sg * sa
    A[..., 0, 1] = cg * cs + sg * ca
    A[..., 0, 2] = -cg * sb
    A[..., 1, 0] = -sg * cc - cg * sa
    A[..., 1, 1] = -sg * cs + cg * ca
    A[..., 1, 2] = sg * sb
    A[..., 2, 0] = sc
    A[..., 2, 1] = ss
    A[..., 2, 2] = cb
    return A
