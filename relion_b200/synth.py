"""Seeded synthetic inputs for the E-step tests and bench.py.

Mirrors what `relion_project --ctf --add_noise --white_noise` produces from a phantom
(/root/reference/src/apps/project.cpp:62-140) and what RELION's host code derives from it before the
E-step, restated from the cited reference code:

* reference volume  -> padded Fourier volume `PPref`  Projector::computeFourierTransformMap
                                                       (src/projector.cpp:116-592, griddingCorrect :595-628)
* CTF image                                            CTF::initialise / getCTF / getFftwImage
                                                       (src/ctf.cpp:211-261, src/ctf.h:184-256); known answer
                                                       tests/ctf.cpp:5-10 (0.59154)
* central slice (numpy, float64, for small tests)      Projector::project (src/projector.cpp:630-797)
* noise model: sigma2_noise[ires] = variance of the real or imaginary part of a Fourier component.

Host-side harness code: numpy only, no GPU.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from .sampling import euler_matrix


# --------------------------------------------------------------------------------------------------
# phantom and its padded Fourier transform
# --------------------------------------------------------------------------------------------------
def make_phantom(n: int, n_blobs: int = 60, seed: int = 1993, radius_frac: float = 0.32) -> np.ndarray:
    """Asymmetric sum of Gaussian blobs inside a sphere of radius radius_frac*n; [n, n, n] float64, origin at n//2."""
    rng = np.random.default_rng(seed)
    vol = np.zeros((n, n, n), np.float64)
    R = radius_frac * n
    o = n // 2
    for _ in range(n_blobs):
        while True:
            p = rng.uniform(-R, R, 3)
            if np.linalg.norm(p) < R:
                break
        s = rng.uniform(0.02, 0.06) * n
        a = rng.uniform(0.5, 1.5)
        # evaluate each blob inside its +-5 sigma box only (exp(-12.5) ~ 4e-6 outside)
        h = int(math.ceil(5 * s))
        lo = [max(0, int(round(p[i])) + o - h) for i in range(3)]      # x, y, z
        hi = [min(n, int(round(p[i])) + o + h + 1) for i in range(3)]
        gx = np.arange(lo[0], hi[0]) - o - p[0]
        gy = np.arange(lo[1], hi[1]) - o - p[1]
        gz = np.arange(lo[2], hi[2]) - o - p[2]
        ex, ey, ez = (np.exp(-g * g / (2 * s * s)) for g in (gx, gy, gz))
        vol[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] += a * ez[:, None, None] * ey[None, :, None] * ex[None, None, :]
    return vol


def pad_size_for(r_max: int, padding_factor: float = 2.0) -> int:
    """Projector::initialiseData (src/projector.cpp:70): pad_size = 2*(ROUND(pf*r_max)+1)+1."""
    return 2 * (int(math.floor(padding_factor * r_max + 0.5)) + 1) + 1


def reference_ft(vol: np.ndarray, current_size: int | None = None, padding_factor: float = 2.0,
                 do_gridding: bool = True):
    """computeFourierTransformMap for a 3D reference used with 2D images.

    Returns (data complex128 [pad, pad, pad//2+1] with y,z origin at (pad-1)//2 and x origin 0, r_max).
    """
    ori = vol.shape[0]
    r_max = min((current_size if current_size else ori) // 2, ori // 2)
    padori = int(math.floor(padding_factor * ori + 0.5))
    padori += padori % 2
    pf = padori / ori
    v = vol.astype(np.float64).copy()
    if do_gridding:
        # griddingCorrect, TRILINEAR: divide by sinc^2(r / (ori*pf))
        c = np.arange(ori) - ori // 2
        z, y, x = np.meshgrid(c, c, c, indexing="ij")
        r = np.sqrt(x * x + y * y + z * z)
        rval = r / (ori * pf)
        sinc = np.ones_like(rval)
        nz = rval > 0
        sinc[nz] = np.sin(np.pi * rval[nz]) / (np.pi * rval[nz])
        v /= sinc * sinc
    Mpad = np.zeros((padori,) * 3, np.float64)
    o = padori // 2 - ori // 2
    Mpad[o:o + ori, o:o + ori, o:o + ori] = v
    F = None
    if padori >= 256:
        # data generation only: use the GPU's FFT for the 512^3 transform when one is there
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.from_numpy(np.fft.ifftshift(Mpad)).cuda()
                F = (torch.fft.rfftn(t) / float(padori) ** 3).cpu().numpy()
                del t
        except Exception:
            F = None
    if F is None:
        F = np.fft.rfftn(np.fft.ifftshift(Mpad)) / float(padori) ** 3      # normalised forward FFT (src/fftw.cpp:333-360)
    normfft = pf * pf * pf * ori                                            # 3D reference, 2D data (:147-163)
    pad = pad_size_for(r_max, pf)
    h = (pad - 1) // 2
    max_r2 = int(math.floor(r_max * pf + 0.5)) ** 2
    k = np.arange(-h, h + 1)
    kx = np.arange(0, pad // 2 + 1)
    # the data array reaches one sample beyond round(r_max*pf) in every direction (room for the trilinear
    # neighbour); those samples lie outside the FFT grid when r_max is Nyquist and are zero anyway
    kin = (k >= -(padori // 2 - 1)) & (k <= padori // 2)   # FFTW frequencies of an even-sized transform
    kxin = kx <= padori // 2
    sub = np.zeros((pad, pad, pad // 2 + 1), np.complex128)
    sub[np.ix_(kin, kin, kxin)] = F[np.ix_(k[kin] % padori, k[kin] % padori, kx[kxin])] * normfft
    kz, ky, kxx = np.meshgrid(k, k, kx, indexing="ij")
    sub[(kz * kz + ky * ky + kxx * kxx) > max_r2] = 0
    return np.ascontiguousarray(sub), r_max


def make_phantom_2d(n: int, n_blobs: int = 25, seed: int = 1993, radius_frac: float = 0.32) -> np.ndarray:
    """Asymmetric sum of 2D Gaussian blobs; [n, n] float64, origin at n//2 (a 2D class average)."""
    rng = np.random.default_rng(seed)
    c = np.arange(n) - n // 2
    y, x = np.meshgrid(c, c, indexing="ij")
    img = np.zeros((n, n), np.float64)
    R = radius_frac * n
    for _ in range(n_blobs):
        while True:
            p = rng.uniform(-R, R, 2)
            if np.linalg.norm(p) < R:
                break
        sg = rng.uniform(0.02, 0.06) * n
        img += rng.uniform(0.5, 1.5) * np.exp(-((x - p[0]) ** 2 + (y - p[1]) ** 2) / (2 * sg * sg))
    return img


def reference_ft_2d(img: np.ndarray, current_size: int | None = None, padding_factor: float = 2.0):
    """computeFourierTransformMap for a 2D reference used with 2D images (src/projector.cpp:116-592, ref_dim == 2,
    data_dim == 2: normfft = pf^2).  Returns (data complex128 [pad, pad//2+1], y origin at (pad-1)//2, r_max)."""
    ori = img.shape[0]
    r_max = min((current_size if current_size else ori) // 2, ori // 2)
    padori = int(math.floor(padding_factor * ori + 0.5))
    padori += padori % 2
    pf = padori / ori
    c = np.arange(ori) - ori // 2
    y, x = np.meshgrid(c, c, indexing="ij")
    rval = np.sqrt(x * x + y * y) / (ori * pf)
    sinc = np.ones_like(rval)
    nz = rval > 0
    sinc[nz] = np.sin(np.pi * rval[nz]) / (np.pi * rval[nz])
    v = img.astype(np.float64) / (sinc * sinc)                              # griddingCorrect, TRILINEAR (:595-628)
    Mpad = np.zeros((padori, padori), np.float64)
    o = padori // 2 - ori // 2
    Mpad[o:o + ori, o:o + ori] = v
    F = np.fft.rfft2(np.fft.ifftshift(Mpad)) / float(padori) ** 2
    normfft = pf * pf
    pad = pad_size_for(r_max, pf)
    h = (pad - 1) // 2
    max_r2 = int(math.floor(r_max * pf + 0.5)) ** 2
    k = np.arange(-h, h + 1)
    kx = np.arange(0, pad // 2 + 1)
    kin = (k >= -(padori // 2 - 1)) & (k <= padori // 2)
    kxin = kx <= padori // 2
    sub = np.zeros((pad, pad // 2 + 1), np.complex128)
    sub[np.ix_(kin, kxin)] = F[np.ix_(k[kin] % padori, kx[kxin])] * normfft
    ky, kxx = np.meshgrid(k, kx, indexing="ij")
    sub[(ky * ky + kxx * kxx) > max_r2] = 0
    return np.ascontiguousarray(sub), r_max


# --------------------------------------------------------------------------------------------------
# CTF
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class CTF:
    defU: float
    defV: float
    defAng: float
    kV: float = 300.0
    Cs: float = 2.7
    Q0: float = 0.1
    Bfac: float = 0.0
    scale: float = 1.0
    phase_shift: float = 0.0

    def __post_init__(self):
        local_Cs = self.Cs * 1e7
        local_kV = self.kV * 1e3
        az = math.radians(self.defAng)
        self.lam = 12.2643247 / math.sqrt(local_kV * (1.0 + local_kV * 0.978466e-6))
        self.K1 = math.pi / 2 * 2 * self.lam
        self.K2 = math.pi / 2 * local_Cs * self.lam ** 3
        self.K3 = math.atan(self.Q0 / math.sqrt(1 - self.Q0 * self.Q0))
        self.K4 = -self.Bfac / 4.0
        self.K5 = math.radians(self.phase_shift)
        ca, sa = math.cos(az), math.sin(az)
        Q = np.array([[ca, sa], [-sa, ca]])
        D = np.array([[-self.defU, 0.0], [0.0, -self.defV]])
        A = Q.T @ D @ Q
        self.Axx, self.Axy, self.Ayy = A[0, 0], A[0, 1], A[1, 1]

    def get_ctf(self, X, Y):
        X = np.asarray(X, np.float64)
        Y = np.asarray(Y, np.float64)
        u2 = X * X + Y * Y
        gamma = self.K1 * (self.Axx * X * X + 2.0 * self.Axy * X * Y + self.Ayy * Y * Y) + self.K2 * u2 * u2 - self.K5 - self.K3
        r = -np.sin(gamma) * np.exp(self.K4 * u2) * self.scale
        small = np.abs(r) < 1e-8
        return np.where(small, np.where(r < 0, -1e-8, 1e-8), r)

    def fftw_image(self, n: int, ori_size: int, angpix: float):
        """CTF::getFftwImage on a window of size n of an ori_size box: [n, n//2+1] float64."""
        xs = ori_size * angpix
        iy = np.arange(n)
        ip = np.where(iy < n // 2 + 1, iy, iy - n)
        jp = np.arange(n // 2 + 1)
        return self.get_ctf(jp[None, :] / xs, ip[:, None] / xs)


# --------------------------------------------------------------------------------------------------
# numpy Fourier-slice projection (float64) for small tests and data generation
# --------------------------------------------------------------------------------------------------
def project_numpy(data: np.ndarray, r_max: int, padding_factor: float, A_inv: np.ndarray, n: int) -> np.ndarray:
    """Central slice [n, n//2+1] complex128 of the padded volume `data` (Projector::project, trilinear); a 2D
    `data` [pad, pad//2+1] is rotated in plane (Projector::rotate2D) through the same code with a zero second plane."""
    if data.ndim == 2:
        pad2 = data.shape[0]
        emb = np.zeros((pad2, pad2, data.shape[1]), data.dtype)
        emb[(pad2 - 1) // 2] = data                       # the z = 0 plane of a centred volume
        data = emb
    pad = data.shape[0]
    init = -((pad - 1) // 2)
    xs = n // 2 + 1
    iy = np.arange(n)
    y = np.where(iy < xs, iy, iy - n).astype(np.float64)[:, None]
    x = np.arange(xs, dtype=np.float64)[None, :]
    my_r_max = min(r_max, xs - 1)
    Ai = A_inv * padding_factor
    xp = Ai[0, 0] * x + Ai[0, 1] * y
    yp = Ai[1, 0] * x + Ai[1, 1] * y
    zp = Ai[2, 0] * x + Ai[2, 1] * y
    inside = (x * x + y * y) <= my_r_max * my_r_max
    inside &= (xp * xp + yp * yp + zp * zp) <= (my_r_max * padding_factor) ** 2 * (1 + 1e-6)   # a scaling A_inv can leave the reference
    neg = xp < 0
    xp = np.where(neg, -xp, xp); yp = np.where(neg, -yp, yp); zp = np.where(neg, -zp, zp)
    x0 = np.floor(xp).astype(np.int64); fx = xp - x0
    y0 = np.floor(yp).astype(np.int64); fy = yp - y0
    z0 = np.floor(zp).astype(np.int64); fz = zp - z0
    yi = np.clip(y0 - init, 0, pad - 2); zi = np.clip(z0 - init, 0, pad - 2); xi = np.clip(x0, 0, data.shape[2] - 2)
    out = np.zeros((n, xs), np.complex128)
    for dz, wz in ((0, 1 - fz), (1, fz)):
        for dy, wy in ((0, 1 - fy), (1, fy)):
            for dx, wx in ((0, 1 - fx), (1, fx)):
                out += data[zi + dz, yi + dy, xi + dx] * (wz * wy * wx)
    out = np.where(neg, np.conj(out), out)
    return np.where(inside, out, 0)


def mresol(n: int) -> np.ndarray:
    """Mresol_fine / Mresol_coarse for window n (src/ml_optimiser.cpp:5784-5811): [n, n//2+1] int, -1 = excluded."""
    xs = n // 2 + 1
    iy = np.arange(n)
    ip = np.where(iy < xs, iy, iy - n)[:, None]
    jp = np.arange(xs)[None, :]
    ires = np.floor(np.sqrt((ip * ip + jp * jp).astype(np.float64)) + 0.5).astype(np.int64)
    ok = (ires < xs) & ~((jp == 0) & (ip < 0))
    return np.where(ok, ires, -1)


def window_ft(a: np.ndarray, nout: int) -> np.ndarray:
    """windowFourierTransform, shrinking branch (src/fftw.h:850-856) on [..., n, n//2+1]."""
    nin = a.shape[-2]
    if nin == nout:
        return a
    xo = nout // 2 + 1
    i = np.arange(nout)
    ip = np.where(i < xo, i, i - nout)
    return a[..., ip % nin, :xo]


# --------------------------------------------------------------------------------------------------
# particle stacks
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SyntheticParticles:
    Fimg: np.ndarray          # [P, n, n//2+1] complex64
    Fimg_nomask: np.ndarray
    Fctf: np.ndarray          # [P, n, n//2+1] float32
    rot: np.ndarray
    tilt: np.ndarray
    psi: np.ndarray
    shift: np.ndarray         # [P, 2] true shifts in pixels
    sigma2_noise: np.ndarray  # [ori_size//2+1]
    highres_Xi2: np.ndarray   # [P]
    ctf_params: np.ndarray = None   # [P, 3] defU, defV, defAngle (300 kV, Cs 2.7 mm, Q0 0.1)


def phase_shift_image(n: int, ori_size: int, sx: float, sy: float) -> np.ndarray:
    """exp(i*(x*tx + y*ty)) with tx = -2*pi*sx/ori_size: the factor the E-step kernels apply for a shift (sx, sy)."""
    xs = n // 2 + 1
    iy = np.arange(n)
    y = np.where(iy < xs, iy, iy - n).astype(np.float64)[:, None]
    x = np.arange(xs, dtype=np.float64)[None, :]
    return np.exp(1j * (-2 * np.pi / ori_size) * (x * sx + y * sy))


def make_particles(slices: np.ndarray, ori_size: int, angpix: float, snr: float, seed: int,
                   rot, tilt, psi, shifts, defocus_range=(10000.0, 30000.0), nomask_extra: float = 0.05) -> SyntheticParticles:
    """Turn noise-free central slices [P, n, n//2+1] into CTF-modulated, shifted, noisy particle FTs.

    The slice of particle p is what the kernels would compute for its true orientation; the particle is
    X = CTF * conj-shift(slice) + noise, so that applying the search translation equal to `shifts[p]`
    re-aligns it with the reference.
    """
    rng = np.random.default_rng(seed)
    P, n, xs = slices.shape
    Fctf = np.empty((P, n, xs), np.float32)
    Fn = np.empty((P, n, xs), np.complex64)
    Fn0 = np.empty((P, n, xs), np.complex64)
    ctf_params = np.zeros((P, 3), np.float64)
    for p in range(P):
        d = rng.uniform(*defocus_range)
        ctf_params[p] = (d, d + rng.uniform(-500.0, 500.0), rng.uniform(0.0, 180.0))
        ctf = CTF(*ctf_params[p])
        c = ctf.fftw_image(n, ori_size, angpix)
        Fctf[p] = c
        Fn[p] = slices[p] * c * np.conj(phase_shift_image(n, ori_size, shifts[p, 0], shifts[p, 1]))
    M = mresol(n)
    valid = M > 0
    nsamp = min(P, 64)
    signal_power = float(np.mean(np.abs(Fn[:nsamp][:, valid].astype(np.complex128)) ** 2))
    s2 = signal_power / (2.0 * snr) if snr > 0 else 1.0        # variance of re and of im
    sd = math.sqrt(s2)
    for p0 in range(0, P, 64):                                  # chunked: keeps the 256-px pools within a few GB
        p1 = min(P, p0 + 64)
        shp = (p1 - p0, n, xs)
        noise = (rng.standard_normal(shp, dtype=np.float32) + 1j * rng.standard_normal(shp, dtype=np.float32)) * np.float32(sd)
        extra = (rng.standard_normal(shp, dtype=np.float32) + 1j * rng.standard_normal(shp, dtype=np.float32)) * np.float32(sd * nomask_extra)
        Fn[p0:p1] += noise
        Fn0[p0:p1] = Fn[p0:p1] + extra
    sigma2 = np.full(ori_size // 2 + 1, s2, np.float64)
    xi2 = np.zeros(P, np.float64)
    if n < ori_size:
        # power between the current window and Nyquist that the windowed images no longer carry
        npix_hi = math.pi / 2 * ((ori_size / 2) ** 2 - (n / 2) ** 2)
        xi2 = rng.uniform(0.9, 1.1, P) * 2.0 * s2 * npix_hi
    return SyntheticParticles(Fn, Fn0, Fctf,
                              np.asarray(rot, np.float64), np.asarray(tilt, np.float64), np.asarray(psi, np.float64),
                              np.asarray(shifts, np.float64), sigma2, xi2, ctf_params)


def inverse_euler_f32(rot, tilt, psi) -> np.ndarray:
    """[n, 9] float32 inverted (transposed) ZYZ matrices, the layout the kernels take."""
    A = euler_matrix(np.asarray(rot, np.float64), np.asarray(tilt, np.float64), np.asarray(psi, np.float64))
    return np.ascontiguousarray(np.swapaxes(A, -1, -2).reshape(-1, 9).astype(np.float32))


def raw_images_from_ft(F: np.ndarray, ori_size: int) -> np.ndarray:
    """Real-space particle images [P, n, n] float32 whose (unmasked) RELION transform, windowed to F's size, is F
    (up to the Hermitian symmetrisation of the x = 0 column): the inverse of normalizeAndTransformImage
    (zero-pad to ori_size, unnormalised inverse FFT, origin back to the box centre)."""
    P, cs, xs = F.shape
    n = ori_size
    out = np.empty((P, n, n), np.float32)
    iy = np.arange(cs)
    ip = np.where(iy < xs, iy, iy - cs)
    for p in range(P):
        Fp = np.zeros((n, n // 2 + 1), np.complex128)
        Fp[ip % n, :xs] = F[p]
        out[p] = np.roll(np.fft.irfft2(Fp, s=(n, n)) * float(n * n), (n // 2, n // 2), axis=(0, 1))
    return out
