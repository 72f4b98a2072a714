"""TEST INFRASTRUCTURE ONLY — ctypes bindings of the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs may import this.
Two kernel providers share one E-step driver (oracle/estep_driver.cpp):

  kind="port"       oracle/liboracle.so            restated kernels (oracle/port_kernels.cpp)
  kind="reference"  oracle/_ref/librefkernels.so   the reference's own ALTCPU kernels compiled from
                                                   /root/reference by oracle/Makefile (prebuilt .so travels
                                                   to the GPU box; absent => RuntimeError, never a silent swap)
  kind="refcuda"    oracle/_ref/librefcuda.so      the reference's own CUDA kernels (diff2.cuh, wavg.cuh, BP.cuh, texture
                                                   projector) compiled for sm_100 (oracle/ref_cuda_kernels.cu): the TIMING
                                                   baseline "reference --gpu path" of bench.py; needs a GPU, one host thread
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from relion_b200 import capi
from relion_b200.estep import (ModelParams, ParticlePool, marshal_model, marshal_pool, marshal_sampling,
                               make_pool_out, _ptr)

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "librefkernels.so")
REFCUDA_LIB = os.path.join(HERE, "_ref", "librefcuda.so")

f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_ubyte)
ulp = C.POINTER(C.c_ulong)


class ok_projector(C.Structure):
    _fields_ = [("mdl", f32p), ("mdlX", C.c_int), ("mdlY", C.c_int), ("mdlZ", C.c_int),
                ("mdlInitY", C.c_int), ("mdlInitZ", C.c_int), ("mdlMaxR", C.c_int), ("padding_factor", C.c_float)]


class ok_backprojector(C.Structure):
    _fields_ = [("real", f32p), ("imag", f32p), ("weight", f32p),
                ("mdlX", C.c_int), ("mdlY", C.c_int), ("mdlZ", C.c_int),
                ("mdlInitY", C.c_int), ("mdlInitZ", C.c_int), ("maxR", C.c_int), ("padding_factor", C.c_float),
                ("sync", C.c_void_p)]


PP = C.POINTER(ok_projector)
BP = C.POINTER(ok_backprojector)


class ok_kernel_table(C.Structure):
    _fields_ = [
        ("kind", C.c_char_p),
        ("make_eulers_3d", C.CFUNCTYPE(None, f32p, f32p, f32p, f32p, C.c_ulong, f32p, f32p)),
        ("project", C.CFUNCTYPE(None, PP, C.c_int, C.c_int, f32p, f32p, f32p)),
        ("diff2_coarse", C.CFUNCTYPE(None, PP, C.c_int, C.c_int, f32p, C.c_ulong, f32p, f32p, C.c_ulong, f32p, f32p, f32p, f32p)),
        ("diff2_fine", C.CFUNCTYPE(None, PP, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, C.c_float,
                                   C.c_ulong, C.c_ulong, C.c_ulong, ulp, ulp, ulp, ulp, f32p)),
        ("weights_exponent_coarse", C.CFUNCTYPE(None, f32p, u8p, f32p, u8p, f32p, C.c_float, C.c_ulong, C.c_ulong, C.c_size_t)),
        ("exponentiate", C.CFUNCTYPE(None, f32p, C.c_float, C.c_size_t)),
        ("exponentiate_weights_fine", C.CFUNCTYPE(None, f32p, u8p, f32p, u8p, f32p, C.c_float, C.c_ulong, C.c_ulong,
                                                  ulp, ulp, ulp, ulp, C.c_long)),
        ("collect2jobs", C.CFUNCTYPE(None, C.c_int, f32p, f32p, f32p, f32p, C.c_float, C.c_float,
                                     C.c_ulong, C.c_ulong, C.c_ulong, C.c_ulong, f32p, f32p, f32p, f32p, ulp, ulp, ulp, ulp)),
        ("wavg", C.CFUNCTYPE(None, PP, C.c_int, C.c_int, f32p, C.c_ulong, f32p, f32p, f32p, f32p, f32p, f32p,
                             f32p, f32p, f32p, C.c_ulong, C.c_float, C.c_float, C.c_float)),
        ("backproject", C.CFUNCTYPE(None, BP, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                    C.c_ulong, C.c_float, C.c_float, f32p, C.c_ulong)),
        ("bp_sync_alloc", C.CFUNCTYPE(C.c_void_p, C.c_int, C.c_int)),
        ("bp_sync_free", C.CFUNCTYPE(None, C.c_void_p)),
        ("backproject2d", C.CFUNCTYPE(None, BP, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                      C.c_ulong, C.c_float, C.c_float, f32p, C.c_ulong)),
        ("diff2_cc_coarse", C.CFUNCTYPE(None, PP, C.c_int, C.c_int, f32p, C.c_ulong, f32p, f32p, C.c_ulong, f32p, f32p, f32p, f32p)),
        ("diff2_cc_fine", C.CFUNCTYPE(None, PP, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p,
                                      C.c_ulong, C.c_ulong, C.c_ulong, ulp, ulp, ulp, ulp, f32p)),
        ("backproject_sgd", C.CFUNCTYPE(None, BP, PP, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                        C.c_ulong, C.c_float, C.c_float, f32p, C.c_ulong)),
        ("backproject2d_sgd", C.CFUNCTYPE(None, BP, PP, C.c_int, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                          C.c_ulong, C.c_float, C.c_float, f32p, C.c_ulong)),
    ]


class ok_debug(C.Structure):
    _fields_ = [("particle", C.c_int), ("coarse_diff2", f32p), ("coarse_weights", f32p), ("coarse_significant", u8p),
                ("fine_capacity", C.c_int64), ("fine_count", C.c_int64), ("fine_ihidden_over", C.POINTER(C.c_int64)),
                ("fine_diff2", f32p), ("fine_weights", f32p),
                ("wdiff2s_parts", f32p), ("wdiff2s_AA", f32p), ("wdiff2s_XA", f32p)]


def build(ref: bool = True):
    """make port (+ ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref and os.path.isdir("/root/reference/src/acc"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_libs = {}


def have_reference() -> bool:
    return os.path.exists(REF_LIB)


def have_refcuda() -> bool:
    return os.path.exists(REFCUDA_LIB)


def _load(kind: str):
    if kind in _libs:
        return _libs[kind]
    if not os.path.exists(PORT_LIB):
        build(ref=False)
    port = C.CDLL(PORT_LIB)
    port.portk_kernel_table.restype = C.POINTER(ok_kernel_table)
    port.oracle_estep_pool.restype = C.c_int
    port.oracle_estep_pool.argtypes = [C.POINTER(ok_kernel_table), C.POINTER(capi.rb_model), C.POINTER(capi.rb_sampling),
                                       PP, BP, C.POINTER(capi.rb_particles), C.POINTER(capi.rb_pool_out),
                                       C.c_uint, C.c_int, C.c_int, C.POINTER(ok_debug)]
    port.oracle_significance.restype = C.c_int64
    port.oracle_significance.argtypes = [f32p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_int, f32p, f32p, C.POINTER(C.c_int64)]
    f64p = C.POINTER(C.c_double)
    port.oracle_backproject_posed.restype = C.c_int
    port.oracle_backproject_posed.argtypes = [f64p, f64p, f64p, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_double]
    if kind == "port":
        table = port.portk_kernel_table()
    elif kind == "reference":
        if not os.path.exists(REF_LIB):
            raise RuntimeError(f"{REF_LIB} missing: build it in the dev container with `make -C oracle ref`")
        ref = C.CDLL(REF_LIB)
        ref.refk_kernel_table.restype = C.POINTER(ok_kernel_table)
        table = ref.refk_kernel_table()
        _libs["_ref_handle"] = ref
    elif kind == "refcuda":
        if not os.path.exists(REFCUDA_LIB):
            raise RuntimeError(f"{REFCUDA_LIB} missing: build it in the dev container with `make -C oracle refcuda`")
        rc = C.CDLL(REFCUDA_LIB)
        rc.refcuda_kernel_table.restype = C.POINTER(ok_kernel_table)
        rc.refcuda_timers.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int]
        rc.refcuda_bp_download.argtypes = [BP]
        table = rc.refcuda_kernel_table()
        _libs["_refcuda_handle"] = rc
    else:
        raise ValueError(kind)
    _libs[kind] = (port, table)
    return _libs[kind]


def _fp(a):
    return a.ctypes.data_as(f32p)


class Projector:
    """ok_projector over a complex64 [Z, Y, X] volume."""

    def __init__(self, vol: np.ndarray, r_max: int, padding_factor: float = 2.0):
        vol = np.asarray(vol)
        self.is_2d = vol.ndim == 2
        if self.is_2d:
            # 2D reference: two planes, the second one zero, z origin 0.  With in-plane rotations zp == 0 and the 3D
            # kernels reduce exactly to project2Dmodel (tests/test_oracle.py pins this against the reference's 2D kernels)
            vol = np.stack([vol, np.zeros_like(vol)])
        self.vol = np.ascontiguousarray(vol, dtype=np.complex64)
        z, y, x = self.vol.shape
        init = -((y - 1) // 2)
        self.struct = ok_projector(_fp(self.vol.view(np.float32)), x, y, z, init, 0 if self.is_2d else init, r_max, padding_factor)


class Backprojector:
    def __init__(self, shape_zyx, r_max: int, padding_factor: float = 2.0):
        """shape (Z, Y, X), or (Y, X) for the 2D accumulator of 2D classification (mdlZ == 1 -> backproject2D)."""
        self.is_2d = len(shape_zyx) == 2
        z, y, x = ((1,) + tuple(shape_zyx)) if self.is_2d else shape_zyx
        self.real = np.zeros(shape_zyx, np.float32)
        self.imag = np.zeros(shape_zyx, np.float32)
        self.weight = np.zeros(shape_zyx, np.float32)
        init = -((y - 1) // 2)
        self.struct = ok_backprojector(_fp(self.real), _fp(self.imag), _fp(self.weight), x, y, z, init, 0 if self.is_2d else init,
                                       r_max, padding_factor, None)


class Oracle:
    """CPU oracle on one kernel provider."""

    def __init__(self, kind: str = "port"):
        self.kind = kind
        self.lib, self.table = _load(kind)
        self.K = self.table.contents

    # ---- whole E-step -------------------------------------------------------------------------
    def estep_pool(self, model: ModelParams, sampling, refs, bps, pool: ParticlePool, skip_maximization=False,
                   num_threads=0, exact_threshold=False, debug_particle=None, fine_capacity=1 << 20):
        mm, ms, mp = marshal_model(model), marshal_sampling(sampling), marshal_pool(pool)
        out = make_pool_out(pool.n_particles, model.ori_size // 2 + 1, model.nr_classes, sampling.n_dir)
        K = model.nr_classes
        parr = (ok_projector * K)(*[r.struct for r in refs])
        nb = len(bps)                # K, or 2 K with the pseudo half-sets of gradient refinement (pool.bp_offset)
        barr = (ok_backprojector * nb)(*[b.struct for b in bps])
        syncs = []
        if num_threads != 1:
            for k in range(nb):
                s = self.K.bp_sync_alloc(barr[k].mdlY, barr[k].mdlZ)
                barr[k].sync = s
                syncs.append(s)
        dbg = None
        dbg_arrays = None
        if debug_particle is not None:
            p = debug_particle
            if pool.dir_off is not None:
                nd = int(pool.dir_off[p + 1] - pool.dir_off[p]); npsi = int(pool.psi_off[p + 1] - pool.psi_off[p])
            else:
                nd, npsi = sampling.n_dir, sampling.n_psi
            nc = K * nd * npsi * sampling.n_trans
            npf = model.current_size * (model.current_size // 2 + 1)
            dbg_arrays = dict(
                coarse_diff2=np.zeros(nc, np.float32), coarse_weights=np.zeros(nc, np.float32),
                coarse_significant=np.zeros(nc, np.uint8), fine_ihidden_over=np.zeros(fine_capacity, np.int64),
                fine_diff2=np.zeros(fine_capacity, np.float32), fine_weights=np.zeros(fine_capacity, np.float32),
                wdiff2s_parts=np.zeros(npf, np.float32), wdiff2s_AA=np.zeros(K * npf, np.float32), wdiff2s_XA=np.zeros(K * npf, np.float32))
            a = dbg_arrays
            dbg = ok_debug(p, _fp(a["coarse_diff2"]), _fp(a["coarse_weights"]), a["coarse_significant"].ctypes.data_as(u8p),
                           fine_capacity, 0, a["fine_ihidden_over"].ctypes.data_as(C.POINTER(C.c_int64)),
                           _fp(a["fine_diff2"]), _fp(a["fine_weights"]), _fp(a["wdiff2s_parts"]), _fp(a["wdiff2s_AA"]), _fp(a["wdiff2s_XA"]))
        st = self.lib.oracle_estep_pool(self.table, C.byref(mm.struct), C.byref(ms.struct), parr, barr, C.byref(mp.struct),
                                        C.byref(out.struct), 1 if skip_maximization else 0, num_threads, int(exact_threshold),
                                        C.byref(dbg) if dbg is not None else None)
        for s in syncs:
            self.K.bp_sync_free(s)
        if self.kind == "refcuda":
            for k in range(nb):
                _libs["_refcuda_handle"].refcuda_bp_download(C.byref(barr[k]))
        if dbg is not None:
            n = int(dbg.fine_count)
            dbg_arrays["fine_count"] = n
            for k in ("fine_ihidden_over", "fine_diff2", "fine_weights"):
                dbg_arrays[k] = dbg_arrays[k][:n]
        return st, out.result, dbg_arrays

    def cuda_timers(self, reset=False):
        """kind="refcuda": summed device time (ms) and launch counts of the reference's CUDA kernels since the last reset."""
        ms = (C.c_double * 4)(); n = (C.c_long * 4)()
        _libs["_refcuda_handle"].refcuda_timers(ms, n, 1 if reset else 0)
        names = ("coarse", "fine", "wavg", "backproject")
        return {k: float(ms[i]) for i, k in enumerate(names)}, {k: int(n[i]) for i, k in enumerate(names)}

    def release_cuda(self):
        _libs["_refcuda_handle"].refcuda_release()

    def significance(self, weights, adaptive_fraction=0.999, maxsig=0, filter_zero=True, exact=False):
        w = np.ascontiguousarray(weights, np.float32)
        s = C.c_float(); g = C.c_float(); nf = C.c_int64()
        idx = self.lib.oracle_significance(_fp(w), w.size, adaptive_fraction, maxsig, int(filter_zero), int(exact),
                                           C.byref(s), C.byref(g), C.byref(nf))
        return dict(threshold_idx=int(idx), sum_weight=float(s.value), significant_weight=float(g.value), n_filtered=int(nf.value))

    # ---- kernel-level calls (stage parity) ------------------------------------------------------
    def make_eulers(self, rot, tilt, psi, mat_left=None, mat_right=None):
        a = np.ascontiguousarray(rot, np.float32); b = np.ascontiguousarray(tilt, np.float32); g = np.ascontiguousarray(psi, np.float32)
        out = np.zeros((len(a), 9), np.float32)
        L = None if mat_left is None else np.ascontiguousarray(mat_left, np.float32).reshape(9)
        R = None if mat_right is None else np.ascontiguousarray(mat_right, np.float32).reshape(9)
        self.K.make_eulers_3d(_fp(a), _fp(b), _fp(g), _fp(out), len(a), None if L is None else _fp(L), None if R is None else _fp(R))
        return out

    def project(self, ref: Projector, n, euler9):
        e = np.ascontiguousarray(euler9, np.float32)
        xs = n // 2 + 1
        re = np.zeros((n, xs), np.float32); im = np.zeros((n, xs), np.float32)
        self.K.project(C.byref(ref.struct), xs, n, _fp(e), _fp(re), _fp(im))
        return re + 1j * im

    def diff2_coarse(self, ref: Projector, n, eulers, tx, ty, re, im, corr, init=None, cc=False):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32); corr = np.ascontiguousarray(corr, np.float32)
        out = np.zeros((e.shape[0], len(tx)), np.float32) if init is None else np.ascontiguousarray(init, np.float32).copy()
        (self.K.diff2_cc_coarse if cc else self.K.diff2_coarse)(C.byref(ref.struct), n // 2 + 1, n, _fp(e), e.shape[0], _fp(tx), _fp(ty), len(tx), _fp(re), _fp(im), _fp(corr), _fp(out))
        return out

    def diff2_cc_fine(self, ref: Projector, n, eulers, tx, ty, re, im, corr, rot_idx, trans_idx, job_idx, job_num):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32); corr = np.ascontiguousarray(corr, np.float32)
        ri = np.ascontiguousarray(rot_idx, np.uint64); ti = np.ascontiguousarray(trans_idx, np.uint64)
        ji = np.ascontiguousarray(job_idx, np.uint64); jn = np.ascontiguousarray(job_num, np.uint64)
        out = np.zeros(len(ri), np.float32)
        up = lambda a: a.ctypes.data_as(ulp)
        self.K.diff2_cc_fine(C.byref(ref.struct), n // 2 + 1, n, _fp(e), _fp(tx), _fp(ty), _fp(re), _fp(im), _fp(corr),
                             e.shape[0], len(tx), len(ji), up(ri), up(ti), up(ji), up(jn), _fp(out))
        return out

    def diff2_fine(self, ref: Projector, n, eulers, tx, ty, re, im, corr, sum_init, rot_idx, trans_idx, job_idx, job_num):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32); corr = np.ascontiguousarray(corr, np.float32)
        ri = np.ascontiguousarray(rot_idx, np.uint64); ti = np.ascontiguousarray(trans_idx, np.uint64)
        ji = np.ascontiguousarray(job_idx, np.uint64); jn = np.ascontiguousarray(job_num, np.uint64)
        out = np.zeros(len(ri), np.float32)
        up = lambda a: a.ctypes.data_as(ulp)
        self.K.diff2_fine(C.byref(ref.struct), n // 2 + 1, n, _fp(e), _fp(tx), _fp(ty), _fp(re), _fp(im), _fp(corr), float(sum_init),
                          e.shape[0], len(tx), len(ji), up(ri), up(ti), up(ji), up(jn), _fp(out))
        return out

    def convert_weights_coarse(self, diff2, pdf_o, pdf_oz, pdf_t, pdf_tz, adaptive_fraction=0.999, maxsig=0, exact=False):
        """weights_exponent_coarse + exponentiate + significance, as convertAllSquaredDifferencesToWeights(0)."""
        w = np.ascontiguousarray(diff2, np.float32).copy()
        no, nt = w.shape
        po = np.ascontiguousarray(pdf_o, np.float32); pt = np.ascontiguousarray(pdf_t, np.float32)
        oz = np.ascontiguousarray(pdf_oz, np.uint8); tz = np.ascontiguousarray(pdf_tz, np.uint8)
        computed = w > np.finfo(np.float32).min
        mn = np.float32(w[computed].min())
        self.K.weights_exponent_coarse(_fp(po), oz.ctypes.data_as(u8p), _fp(pt), tz.ctypes.data_as(u8p), _fp(w), mn, no, nt, w.size)
        mx = np.float32(w.max())
        self.K.exponentiate(_fp(w), np.float32(50.0) - mx, w.size)
        sig = self.significance(w, adaptive_fraction, maxsig, True, exact)
        sig["min_diff2"] = float(mn)
        sig["weights"] = w
        sig["significant"] = (w >= np.float32(sig["significant_weight"])).astype(np.uint8)
        sig["nr_significant"] = sig["n_filtered"] - sig["threshold_idx"]
        return sig

    def wavg(self, ref: Projector, n, eulers, tx, ty, re, im, weights, ctfs, weight_norm, sig_w):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32)
        w = np.ascontiguousarray(weights, np.float32); c = np.ascontiguousarray(ctfs, np.float32)
        np_ = n * (n // 2 + 1)
        parts = np.zeros(np_, np.float32); AA = np.zeros(np_, np.float32); XA = np.zeros(np_, np.float32)
        self.K.wavg(C.byref(ref.struct), n // 2 + 1, n, _fp(e), e.shape[0], _fp(re), _fp(im), _fp(tx), _fp(ty), _fp(w), _fp(c),
                    _fp(parts), _fp(AA), _fp(XA), len(tx), float(weight_norm), float(sig_w), 1.0)
        return parts, AA, XA

    def backproject(self, bp: Backprojector, n, eulers, tx, ty, re, im, weights, minvsigma2, ctfs, weight_norm, sig_w):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32)
        w = np.ascontiguousarray(weights, np.float32); c = np.ascontiguousarray(ctfs, np.float32); mi = np.ascontiguousarray(minvsigma2, np.float32)
        fn = self.K.backproject2d if bp.is_2d else self.K.backproject
        fn(C.byref(bp.struct), n // 2 + 1, n, _fp(re), _fp(im), _fp(tx), _fp(ty), _fp(w), _fp(mi), _fp(c),
           len(tx), float(sig_w), float(weight_norm), _fp(e), e.shape[0])


def backproject_posed(shape_zyx, F2D, Fctf, eulers, r_max, padding_factor=2.0, out=None):
    """Compiled restatement of BackProjector::backproject2Dto3D (oracle/estep_driver.cpp:oracle_backproject_posed), one
    thread, double accumulators like relion_reconstruct: returns (real, imag, weight) float64 [Z, Y, X]."""
    port, _ = _load("port")
    z, y, x = shape_zyx
    if out is None:
        re = np.zeros((z, y, x), np.float64); im = np.zeros_like(re); w = np.zeros_like(re)
    else:
        re, im, w = out                                      # accumulate into caller-owned volumes (timing without the allocation)
    F = np.ascontiguousarray(F2D, np.complex64); W = np.ascontiguousarray(Fctf, np.float32)
    e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
    f64p = C.POINTER(C.c_double)
    st = port.oracle_backproject_posed(re.ctypes.data_as(f64p), im.ctypes.data_as(f64p), w.ctypes.data_as(f64p), x, y, z,
                                       F.view(np.float32).ctypes.data_as(f32p), _fp(W), _fp(e), F.shape[1], F.shape[0], r_max, padding_factor)
    assert st == 0
    return re, im, w
