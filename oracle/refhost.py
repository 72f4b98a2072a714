"""ctypes bindings of oracle/_ref/librefrecon.so: the REFERENCE's own host classes (Projector, BackProjector, softMaskOutsideMap,
FourierTransformer), compiled from /root/reference by `make -C oracle refrecon` (oracle/ref_recon.cpp lists what is wrapped).

TEST INFRASTRUCTURE ONLY: imported by tests/test_reference_host.py and tools/make_host_golden.py, which pin the numpy
restatements (oracle/reconstruct.py, oracle/prepare.py, relion_b200/synth.py) that the -m gpu tests of the rows f1-f3 compare
the device code with.  Nothing under relion_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "librefrecon.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available() -> bool:
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.refrec_last_error.restype = C.c_char_p
    return _lib


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference host code failed: " + _load().refrec_last_error().decode())


def ft_map(vol: np.ndarray, current_size: int = -1, padding_factor: float = 2.0, data_dim: int = 2, do_gridding: bool = True):
    """Projector::computeFourierTransformMap -> (data complex128 [Z, Y, X] (or [Y, X]), (startZ, startY), r_max, power_spectrum)."""
    lib = _load()
    v = _f64(vol)
    ori, dim = v.shape[0], v.ndim
    pad = 2 * (int(np.floor(padding_factor * ori + 0.5)) // 2 + 2) + 1 + 4
    cap = pad * pad * (pad if dim == 3 else 1)
    dims = np.zeros(6, np.int32)
    out = np.zeros(2 * cap, np.float64)
    ps = np.zeros(ori // 2 + 1, np.float64)
    _check(lib.refrec_ft_map(_p(v), ori, dim, int(current_size), C.c_double(padding_factor), int(data_dim), int(do_gridding),
                             _p(dims, C.c_int), _p(out), C.c_longlong(cap), _p(ps)))
    Z, Y, X = (int(d) for d in dims[:3])
    data = out[:2 * Z * Y * X].view(np.complex128).reshape((Z, Y, X) if dim == 3 else (Y, X)).copy()
    return data, (int(dims[3]), int(dims[4])), int(dims[5]), ps


def bp_dims(ori_size: int, ref_dim: int, current_size: int, padding_factor: float = 2.0):
    d = np.zeros(4, np.int32)
    _check(_load().refrec_bp_dims(ori_size, ref_dim, current_size, C.c_double(padding_factor), _p(d, C.c_int)))
    return (int(d[0]), int(d[1]), int(d[2])), int(d[3])


def reconstruct(real, imag, weight, ori_size: int, current_size: int, padding_factor: float = 2.0, tau2=None, tau2_fudge: float = 1.0,
                minres_map: int = -1, skip_gridding: bool = True, max_iter_preweight: int = 10, normalise: float = 1.0):
    """BackProjector::reconstruct on centred accumulators [Z, Y, X] (3D) or [Y, X] (2D) -> real-space map [ori]^dim."""
    re, im, w = _f64(real), _f64(imag), _f64(weight)
    dim = re.ndim
    out = np.zeros((ori_size,) * dim, np.float64)
    t2 = _f64(tau2) if tau2 is not None else np.zeros(ori_size // 2 + 1)
    _check(_load().refrec_reconstruct(_p(re), _p(im), _p(w), ori_size, dim, current_size, C.c_double(padding_factor), int(skip_gridding),
                                      max_iter_preweight, int(tau2 is not None), _p(t2), len(t2), C.c_double(tau2_fudge),
                                      C.c_double(normalise), minres_map, _p(out)))
    return out


def symmetrise(real, imag, weight, ori_size: int, current_size: int, sym: str, padding_factor: float = 2.0):
    re, im, w = (np.array(a, np.float64, copy=True, order="C") for a in (real, imag, weight))
    _check(_load().refrec_symmetrise(_p(re), _p(im), _p(w), ori_size, re.ndim, current_size, C.c_double(padding_factor), sym.encode()))
    return re, im, w


def symmetrise_helical(real, imag, weight, ori_size: int, current_size: int, sym: str, nr_helical_asu: int, helical_twist: float,
                       helical_rise: float, padding_factor: float = 2.0):
    """BackProjector::symmetrise(nr_helical_asu, helical_twist [deg], helical_rise [pixels]) of the compiled reference."""
    re, im, w = (np.array(a, np.float64, copy=True, order="C") for a in (real, imag, weight))
    _check(_load().refrec_symmetrise_helical(_p(re), _p(im), _p(w), ori_size, re.ndim, current_size, C.c_double(padding_factor),
                                             sym.encode(), int(nr_helical_asu), C.c_double(helical_twist), C.c_double(helical_rise)))
    return re, im, w


def sym_matrices(sym: str) -> np.ndarray:
    """The R matrices of SymList::get_matrices for point group `sym`: [nsym, 3, 3]."""
    R = np.zeros(9 * 256, np.float64)
    n = _load().refrec_sym_matrices(sym.encode(), _p(R), 256)
    if n < 0:
        _check(1)
    return R[:9 * n].reshape(n, 3, 3).copy()


def update_ssnr(weight, ori_size: int, current_size: int, padding_factor: float, tau2_fudge: float, tau2, fsc=None, avgctf2=None,
                update_tau2_with_fsc: bool = False, is_whole_instead_of_half: bool = False):
    w = _f64(weight)
    ns = ori_size // 2 + 1
    t2 = np.array(tau2, np.float64, copy=True)
    s2, dvp, cov = np.zeros(ns), np.zeros(ns), np.zeros(ns)
    f, a = _f64(fsc), _f64(avgctf2)
    _check(_load().refrec_update_ssnr(_p(w), ori_size, w.ndim, current_size, C.c_double(padding_factor), C.c_double(tau2_fudge), _p(t2),
                                      _p(s2), _p(dvp), _p(cov), _p(f), _p(a), int(update_tau2_with_fsc), int(is_whole_instead_of_half)))
    return t2, s2, dvp, cov


def backproject(imgs: np.ndarray, A: np.ndarray, weights, ori_size: int, current_size: int, padding_factor: float = 2.0):
    """BackProjector::set2DFourierTransform over images [N, n, n/2+1] complex with the particles' Euler matrices A [N, 3, 3]
    (the call inverts them itself) -> (re, im, w) centred."""
    F = np.ascontiguousarray(imgs, np.complex128)
    N, n = F.shape[0], F.shape[1]
    (Z, Y, X), _ = bp_dims(ori_size, 3, current_size, padding_factor)
    re, im, w = (np.zeros((Z, Y, X), np.float64) for _ in range(3))
    Am = _f64(A)
    W = _f64(weights)
    _check(_load().refrec_backproject(_p(F.view(np.float64)), _p(Am), _p(W), N, n, ori_size, current_size, C.c_double(padding_factor),
                                      _p(re), _p(im), _p(w)))
    return re, im, w


def soft_mask(vol: np.ndarray, radius: float, cosine_width: float) -> np.ndarray:
    v = np.array(vol, np.float64, copy=True, order="C")
    _check(_load().refrec_soft_mask(_p(v), v.shape[0], v.ndim, C.c_double(radius), C.c_double(cosine_width)))
    return v


def image_ft(img: np.ndarray, current_size: int, shift=(0.0, 0.0)) -> np.ndarray:
    """CenterFFT + FourierTransform + windowFourierTransform (+ shiftImageInFourierTransform) -> [cs, cs/2+1] complex128."""
    v = _f64(img)
    out = np.zeros((current_size, current_size // 2 + 1), np.complex128)
    _check(_load().refrec_image_ft(_p(v), v.shape[0], current_size, C.c_double(shift[0]), C.c_double(shift[1]), _p(out.view(np.float64))))
    return out


def fsc(m1: np.ndarray, m2: np.ndarray) -> np.ndarray:
    a, b = _f64(m1), _f64(m2)
    out = np.zeros(a.shape[0] // 2 + 1)
    _check(_load().refrec_fsc(_p(a), _p(b), a.shape[0], _p(out)))
    return out


# ---- the ALTCPU image-preparation helpers compiled into oracle/_ref/librefkernels.so (oracle/ref_kernels.cpp, refk_prep_*) ----
_klib = None
_fp32 = C.POINTER(C.c_float)


def _kernels():
    global _klib
    if _klib is None:
        _klib = C.CDLL(os.path.join(_HERE, "_ref", "librefkernels.so"))
    return _klib


def prep_translate2d(img: np.ndarray, dx: int, dy: int) -> np.ndarray:
    a = np.ascontiguousarray(img, np.float32)
    out = np.zeros_like(a)
    _kernels().refk_prep_translate2d(a.ctypes.data_as(_fp32), out.ctypes.data_as(_fp32), a.shape[0], int(dx), int(dy))
    return out


def prep_soft_mask(img: np.ndarray, radius: float, cosine_width: float):
    a = np.array(img, np.float32, copy=True, order="C")
    bg = C.c_float(0)
    _kernels().refk_prep_soft_mask(a.ctypes.data_as(_fp32), a.shape[0], C.c_float(radius), C.c_float(cosine_width), C.byref(bg))
    return a, float(bg.value)


def prep_power_class(F: np.ndarray, current_size: int):
    a = np.ascontiguousarray(F, np.complex64)
    n = a.shape[0]
    spec = np.zeros(n // 2 + 1, np.float32)
    xi2 = C.c_float(0)
    _kernels().refk_prep_power_class(a.view(np.float32).ctypes.data_as(_fp32), n, int(current_size), spec.ctypes.data_as(_fp32), C.byref(xi2))
    return spec.astype(np.float64), float(xi2.value)
