"""Host-side mirror of RELION's accelerator boundary for the expectation step.

`MlDeviceBundle` plays the role of the reference's MlDeviceBundle + MlOptimiserCuda
(/root/reference/src/acc/cuda/cuda_ml_optimiser.h:18-144): per-device projectors
(`set_reference` = AccProjector::setMdlDim/initMdl), back-projectors (`bp_init`/`bp_get` =
AccBackprojector::setMdlDim/initMdl/getMdlData) and the E-step over a pool of particles
(`expectation_some_particles` = doThreadExpectationSomeParticles for the whole pool at once).
Everything numerical happens in the C-ABI library (relion_b200/capi.py); this file only marshals
numpy / torch host buffers and keeps them alive across the asynchronous calls.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional

import numpy as np

from . import capi
from .sampling import Sampling


def _ptr(a, ctype):
    """ctypes pointer to a contiguous numpy array or (pinned) CPU torch tensor, or NULL."""
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    if hasattr(a, "data_ptr"):  # torch tensor
        assert a.is_contiguous() and a.device.type == "cpu"
        return C.cast(a.data_ptr(), C.POINTER(ctype))
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.POINTER(ctype))


def _fptr(a):
    """float* to a float32 / complex64 numpy array or a CPU torch tensor (pinned host staging buffers)."""
    if hasattr(a, "data_ptr"):
        return _ptr(a, C.c_float)
    a = np.ascontiguousarray(a)
    assert a.dtype in (np.float32, np.complex64), a.dtype
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


@dataclasses.dataclass
class ModelParams:
    """What the E-step reads from MlModel / MlOptimiser (SURVEY.md §8b data contract)."""
    nr_classes: int
    ori_size: int
    coarse_size: int
    current_size: int
    pixel_size: float
    sigma2_noise: np.ndarray                 # [nr_optics_groups, ori_size/2+1]
    scale_correction: np.ndarray             # [nr_groups]
    pdf_class: np.ndarray                    # [K]
    pdf_direction: Optional[np.ndarray]      # [K, n_dir]
    data_vs_prior_class: Optional[np.ndarray] = None  # [K, ori_size/2+1]
    sigma2_offset: float = 9.0
    offset_range: float = 0.0
    sigma2_fudge: float = 1.0
    adaptive_fraction: float = 0.999
    maximum_significants: int = 0
    do_ctf_correction: bool = True
    refs_are_ctf_corrected: bool = True
    do_scale_correction: bool = True
    do_map: bool = True
    ctf_premultiplied: bool = False
    bp_circle_bound: bool = True
    do_cc: bool = False                      # (iter == 1 and do_firstiter_cc) or do_always_cc
    prior_offset_class: Optional[np.ndarray] = None   # [K, 2] pixels: mymodel.prior_offset_class (2D references), None for 3D
    do_grad: bool = False                    # gradient (SGD / VDAM) refinement: residual back-projection
    ref_max_r: int = 0                       # the references' r_max when smaller than current_size / 2 (rb_model.ref_max_r)
    do_skip_rotate: bool = False             # --skip_align / --skip_rotate: one-entry orientation lists, prior = pdf_class


@dataclasses.dataclass
class ParticlePool:
    """One pool of particles after image preparation (getFourierTransformsAndCtfs outputs)."""
    Fimg: np.ndarray                # [P, n, n/2+1] complex64 (masked), n = current_size
    Fimg_nomask: np.ndarray         # [P, n, n/2+1] complex64
    Fctf: Optional[np.ndarray]      # [P, n, n/2+1] float32
    group_id: np.ndarray            # [P] int32
    optics_group: np.ndarray        # [P] int32
    highres_Xi2: np.ndarray         # [P] float64
    old_offset: np.ndarray          # [P, 2] float64 (pixels)
    prior_offset: np.ndarray        # [P, 2] float64
    dir_off: Optional[np.ndarray] = None
    dir_idx: Optional[np.ndarray] = None
    dir_prior: Optional[np.ndarray] = None
    psi_off: Optional[np.ndarray] = None
    psi_idx: Optional[np.ndarray] = None
    psi_prior: Optional[np.ndarray] = None
    bp_offset: Optional[np.ndarray] = None   # [P] int32: accumulator = class + bp_offset (pseudo half-sets of gradient refinement)
    pre_shift: Optional[np.ndarray] = None   # [P, 2] pixels: per-particle translation on top of the sampled ones (--skip_align)
    mat_left: Optional[np.ndarray] = None    # [3, 3] MBL: orientation matrices become inverse(mat_left A mat_right) (magnification / scale
    mat_right: Optional[np.ndarray] = None   # [3, 3] MBR   difference of the optics group, body matrices)

    @property
    def n_particles(self):
        return int(self.Fimg.shape[0])


_OUT_DTYPE = np.dtype([
    ("best_ihidden_over", np.int64),
    ("best_class", np.int32), ("best_idir", np.int32), ("best_ipsi", np.int32),
    ("best_iover_rot", np.int32), ("best_itrans", np.int32), ("best_iover_trans", np.int32),
    ("nr_significant_coarse", np.int32), ("n_fine_orient", np.int32), ("n_fine_samples", np.int32),
    ("min_diff2_coarse", np.float32), ("sum_weight_coarse", np.float32), ("significant_weight_coarse", np.float32),
    ("min_diff2", np.float32), ("max_weight", np.float32), ("sum_weight", np.float32),
    ("significant_weight", np.float32), ("pmax", np.float32), ("n_bp_orient", np.int32),
    ("dLL_nolog", np.float64), ("wsum_norm_correction", np.float64),
    ("wsum_XA", np.float64), ("wsum_AA", np.float64), ("sumw", np.float64), ("wsum_sigma2_offset", np.float64),
], align=True)
assert _OUT_DTYPE.itemsize == C.sizeof(capi.rb_particle_out), (_OUT_DTYPE.itemsize, C.sizeof(capi.rb_particle_out))


@dataclasses.dataclass
class PoolResult:
    particles: np.ndarray           # structured array, one row per particle (rb_particle_out)
    wsum_sigma2_noise: np.ndarray   # [P, ori_size/2+1] float32
    wsum_pdf_direction: np.ndarray  # [K, n_dir] float64
    wsum_pdf_class: np.ndarray      # [K] float64
    wsum_prior_offset_class: Optional[np.ndarray] = None   # [K, 2] float64, Angstrom (2D references with prior_offset_class)


class _Marshalled:
    """ctypes structs plus the host arrays they point into (kept alive together)."""

    def __init__(self):
        self.keep = []

    def hold(self, a):
        self.keep.append(a)
        return a


def marshal_sampling(s: Sampling):
    m = _Marshalled()
    st = capi.rb_sampling()
    st.n_dir, st.n_psi = s.n_dir, s.n_psi
    st.rot = _ptr(m.hold(_f64(s.rot)), C.c_double)
    st.tilt = _ptr(m.hold(_f64(s.tilt)), C.c_double)
    st.psi = _ptr(m.hold(_f64(s.psi)), C.c_double)
    st.n_over_rot = s.n_over_rot
    st.over_rot = _ptr(m.hold(_f64(s.over_rot)), C.c_double)
    st.over_tilt = _ptr(m.hold(_f64(s.over_tilt)), C.c_double)
    st.over_psi = _ptr(m.hold(_f64(s.over_psi)), C.c_double)
    st.n_trans = s.n_trans
    st.trans_x = _ptr(m.hold(_f64(s.trans_x)), C.c_double)
    st.trans_y = _ptr(m.hold(_f64(s.trans_y)), C.c_double)
    st.n_over_trans = s.n_over_trans
    st.over_trans_x = _ptr(m.hold(_f64(s.over_trans_x)), C.c_double)
    st.over_trans_y = _ptr(m.hold(_f64(s.over_trans_y)), C.c_double)
    m.struct = st
    return m


def marshal_model(p: ModelParams):
    m = _Marshalled()
    st = capi.rb_model()
    st.nr_classes, st.ori_size, st.coarse_size, st.current_size = p.nr_classes, p.ori_size, p.coarse_size, p.current_size
    st.pixel_size = p.pixel_size
    s2 = m.hold(_f64(np.atleast_2d(p.sigma2_noise)))
    st.nr_optics_groups = s2.shape[0]
    st.sigma2_noise = _ptr(s2, C.c_double)
    sc = m.hold(_f64(np.atleast_1d(p.scale_correction)))
    st.nr_groups = sc.shape[0]
    st.scale_correction = _ptr(sc, C.c_double)
    st.pdf_class = _ptr(m.hold(_f64(p.pdf_class)), C.c_double)
    st.pdf_direction = _ptr(m.hold(_f64(p.pdf_direction)), C.c_double)
    st.data_vs_prior_class = _ptr(m.hold(_f64(p.data_vs_prior_class)), C.c_double)
    st.sigma2_offset, st.offset_range, st.sigma2_fudge = p.sigma2_offset, p.offset_range, p.sigma2_fudge
    st.adaptive_fraction, st.maximum_significants = p.adaptive_fraction, p.maximum_significants
    st.do_ctf_correction = int(p.do_ctf_correction)
    st.refs_are_ctf_corrected = int(p.refs_are_ctf_corrected)
    st.do_scale_correction = int(p.do_scale_correction)
    st.do_map = int(p.do_map)
    st.ctf_premultiplied = int(p.ctf_premultiplied)
    st.bp_circle_bound = int(p.bp_circle_bound)
    st.do_cc = int(p.do_cc)
    st.do_grad = int(p.do_grad)
    st.ref_max_r = int(p.ref_max_r)
    st.do_skip_rotate = int(p.do_skip_rotate)
    if p.prior_offset_class is not None:
        poc = m.hold(_f64(np.asarray(p.prior_offset_class).reshape(p.nr_classes, 2)))
        st.prior_offset_class = _ptr(poc, C.c_double)
    m.struct = st
    return m


@dataclasses.dataclass
class RawParticlePool:
    """One pool of raw particle images + metadata, what getFourierTransformsAndCtfs starts from (rb_raw_particles)."""
    images: np.ndarray              # [P, n, n] float32 real space (numpy or pinned torch tensor)
    old_offset: np.ndarray          # [P, 2]
    prior_offset: np.ndarray        # [P, 2]
    group_id: np.ndarray
    optics_group: np.ndarray
    ctf_defU: Optional[np.ndarray] = None
    ctf_defV: Optional[np.ndarray] = None
    ctf_defAngle: Optional[np.ndarray] = None
    ctf_Bfac: Optional[np.ndarray] = None
    ctf_scale: Optional[np.ndarray] = None
    ctf_phase_shift: Optional[np.ndarray] = None
    og_kV: Optional[np.ndarray] = None
    og_Cs: Optional[np.ndarray] = None
    og_Q0: Optional[np.ndarray] = None
    norm_factor: Optional[np.ndarray] = None
    mask_radius: float = -1.0
    width_mask_edge: float = 5.0
    dir_off: Optional[np.ndarray] = None
    dir_idx: Optional[np.ndarray] = None
    dir_prior: Optional[np.ndarray] = None
    psi_off: Optional[np.ndarray] = None
    psi_idx: Optional[np.ndarray] = None
    psi_prior: Optional[np.ndarray] = None
    bp_offset: Optional[np.ndarray] = None
    noise_seed: Optional[np.ndarray] = None   # [P] int64 random_seed + part_id: noise-filled soft mask; None: zero mask
    og_fourier_factor: Optional[np.ndarray] = None   # [nog, cs, cs/2+1] complex64: conj(beam-tilt phase) * avgMTF / MTF per optics group
    mat_left: Optional[np.ndarray] = None
    mat_right: Optional[np.ndarray] = None
    pre_shift: Optional[np.ndarray] = None

    @property
    def n_particles(self):
        return int(self.group_id.shape[0])


def marshal_raw_pool(pool: RawParticlePool):
    m = _Marshalled()
    st = capi.rb_raw_particles()
    st.n_particles = pool.n_particles
    img = pool.images if hasattr(pool.images, "data_ptr") else np.ascontiguousarray(pool.images, np.float32)
    st.image_size = int(img.shape[-1])
    st.images = _ptr(m.hold(img), C.c_float)
    for name in ("norm_factor", "old_offset", "prior_offset", "ctf_defU", "ctf_defV", "ctf_defAngle", "ctf_Bfac", "ctf_scale",
                 "ctf_phase_shift", "og_kV", "og_Cs", "og_Q0", "dir_prior", "psi_prior"):
        setattr(st, name, _ptr(m.hold(_f64(getattr(pool, name))), C.c_double))
    for name in ("group_id", "optics_group", "dir_off", "dir_idx", "psi_off", "psi_idx", "bp_offset"):
        setattr(st, name, _ptr(m.hold(_i32(getattr(pool, name))), C.c_int))
    st.mask_radius, st.width_mask_edge = float(pool.mask_radius), float(pool.width_mask_edge)
    if pool.og_fourier_factor is not None:
        st.og_fourier_factor = _ptr(m.hold(np.ascontiguousarray(pool.og_fourier_factor, dtype=np.complex64).view(np.float32)), C.c_float)
    if pool.noise_seed is not None:
        st.noise_seed = m.hold(np.ascontiguousarray(pool.noise_seed, dtype=np.int64)).ctypes.data_as(C.POINTER(C.c_int64))
    st.mat_left = _ptr(m.hold(_f64(pool.mat_left)), C.c_double)
    st.mat_right = _ptr(m.hold(_f64(pool.mat_right)), C.c_double)
    st.pre_shift = _ptr(m.hold(_f64(pool.pre_shift)), C.c_double)
    m.struct = st
    return m


def marshal_pool(pool: ParticlePool):
    m = _Marshalled()
    st = capi.rb_particles()
    st.n_particles = pool.n_particles

    def img(a, dtype):
        if a is None:
            return None
        if hasattr(a, "data_ptr"):
            return m.hold(a)
        return m.hold(np.ascontiguousarray(a, dtype=dtype))

    st.Fimg = _ptr(img(pool.Fimg, np.complex64), C.c_float)
    st.Fimg_nomask = _ptr(img(pool.Fimg_nomask, np.complex64), C.c_float)
    st.Fctf = _ptr(img(pool.Fctf, np.float32), C.c_float)
    st.group_id = _ptr(m.hold(_i32(pool.group_id)), C.c_int)
    st.optics_group = _ptr(m.hold(_i32(pool.optics_group)), C.c_int)
    st.highres_Xi2 = _ptr(m.hold(_f64(pool.highres_Xi2)), C.c_double)
    st.old_offset = _ptr(m.hold(_f64(pool.old_offset)), C.c_double)
    st.prior_offset = _ptr(m.hold(_f64(pool.prior_offset)), C.c_double)
    st.dir_off = _ptr(m.hold(_i32(pool.dir_off)), C.c_int)
    st.dir_idx = _ptr(m.hold(_i32(pool.dir_idx)), C.c_int)
    st.dir_prior = _ptr(m.hold(_f64(pool.dir_prior)), C.c_double)
    st.psi_off = _ptr(m.hold(_i32(pool.psi_off)), C.c_int)
    st.psi_idx = _ptr(m.hold(_i32(pool.psi_idx)), C.c_int)
    st.psi_prior = _ptr(m.hold(_f64(pool.psi_prior)), C.c_double)
    st.bp_offset = _ptr(m.hold(_i32(pool.bp_offset)), C.c_int)
    st.mat_left = _ptr(m.hold(_f64(pool.mat_left)), C.c_double)
    st.mat_right = _ptr(m.hold(_f64(pool.mat_right)), C.c_double)
    st.pre_shift = _ptr(m.hold(_f64(pool.pre_shift)), C.c_double)
    m.struct = st
    return m


def make_pool_out(n_particles: int, nshell: int, nr_classes: int, n_dir: int):
    m = _Marshalled()
    parts = m.hold(np.zeros(n_particles, dtype=_OUT_DTYPE))
    shells = m.hold(np.zeros((n_particles, nshell), dtype=np.float32))
    pdir = m.hold(np.zeros((nr_classes, n_dir), dtype=np.float64))
    pcls = m.hold(np.zeros(nr_classes, dtype=np.float64))
    st = capi.rb_pool_out()
    st.particles = C.cast(parts.ctypes.data, C.POINTER(capi.rb_particle_out))
    st.wsum_sigma2_noise = _ptr(shells, C.c_float)
    st.wsum_pdf_direction = _ptr(pdir, C.c_double)
    st.wsum_pdf_class = _ptr(pcls, C.c_double)
    poff = m.hold(np.zeros((nr_classes, 2), dtype=np.float64))
    st.wsum_prior_offset_class = _ptr(poff, C.c_double)
    m.struct = st
    m.result = PoolResult(parts, shells, pdir, pcls, poff)
    return m


class MlDeviceBundle:
    """One per GPU (cuda_ml_optimiser.h:18-76)."""

    def __init__(self, device: int = 0):
        self.lib = capi.load_library()
        h = C.c_void_p()
        capi.check(self.lib, self.lib.rb_ctx_create(device, C.byref(h)))
        self.ctx = h
        self.device_id = device
        self.model: Optional[ModelParams] = None
        self.sampling: Optional[Sampling] = None
        self._keep = {}

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.rb_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- projectors / back-projectors ---------------------------------------------------------
    def set_reference(self, iclass: int, vol: np.ndarray, r_max: int, padding_factor: float = 2.0):
        """vol: complex [Z, Y, X] padded Fourier volume (MlModel::PPref[k].data layout), or [Y, X] for a 2D reference."""
        if vol.ndim == 2:
            vol = vol[None]
        z, y, x = vol.shape
        init = -((y - 1) // 2)
        if vol.dtype == np.complex128:
            v = np.ascontiguousarray(vol)
            st = self.lib.rb_set_reference(self.ctx, iclass, _ptr(v.view(np.float64), C.c_double), x, y, z, init, init, r_max, padding_factor)
        else:
            v = np.ascontiguousarray(vol, dtype=np.complex64)
            st = self.lib.rb_set_reference_f32(self.ctx, iclass, _ptr(v.view(np.float32), C.c_float), x, y, z, init, init, r_max, padding_factor)
        capi.check(self.lib, st)

    def set_reference_from_map(self, iclass: int, vol: np.ndarray, current_size: int = 0, padding_factor: float = 2.0):
        """rb_set_reference_from_map: computeFourierTransformMap on the device; returns the radial power spectrum [ori/2+1]."""
        v = np.ascontiguousarray(vol, np.float32)
        ori = v.shape[0]
        ps = np.zeros(ori // 2 + 1, np.float64)
        capi.check(self.lib, self.lib.rb_set_reference_from_map(self.ctx, iclass, _ptr(v, C.c_float), ori, int(current_size), float(padding_factor),
                                                                _ptr(ps, C.c_double)))
        return ps

    def bp_init(self, iclass: int, shape_zyx, r_max: int, padding_factor: float = 2.0):
        self._keep[("bp_2d", iclass)] = len(shape_zyx) == 2
        if len(shape_zyx) == 2:
            shape_zyx = (1,) + tuple(shape_zyx)
        z, y, x = shape_zyx
        init = -((y - 1) // 2)
        capi.check(self.lib, self.lib.rb_bp_init(self.ctx, iclass, x, y, z, init, init, r_max, padding_factor))
        self._keep[("bp_shape", iclass)] = (z, y, x)

    def bp_clear(self, iclass: int):
        capi.check(self.lib, self.lib.rb_bp_clear(self.ctx, iclass))

    def bp_get(self, iclass: int):
        z, y, x = self._keep[("bp_shape", iclass)]    # z == 1: 2D accumulator, the library returns its [Y][X] plane
        re = np.empty((z, y, x), np.float32)
        im = np.empty_like(re)
        w = np.empty_like(re)
        capi.check(self.lib, self.lib.rb_bp_get(self.ctx, iclass, _ptr(re, C.c_float), _ptr(im, C.c_float), _ptr(w, C.c_float)))
        if self._keep.get(("bp_2d", iclass)):
            return re[0], im[0], w[0]
        return re, im, w

    def bp_symmetrise(self, iclass: int, rotations=None, helical=None):
        """rb_bp_symmetrise: Hermitian symmetry of the x = 0 plane + point-group mates; rotations [nsym, 3, 3] (None: C1).
        helical = (nr_helical_asu, twist [deg], rise [pixels], ori_size): rb_bp_symmetrise_helical (applyHelicalSymmetry in between)."""
        r = None if rotations is None or len(rotations) == 0 else np.ascontiguousarray(rotations, np.float64).reshape(-1, 9)
        n = 0 if r is None else r.shape[0]
        if helical is None:
            capi.check(self.lib, self.lib.rb_bp_symmetrise(self.ctx, iclass, _ptr(r, C.c_double), n))
        else:
            capi.check(self.lib, self.lib.rb_bp_symmetrise_helical(self.ctx, iclass, _ptr(r, C.c_double), n, int(helical[0]),
                                                                   float(helical[1]), float(helical[2]), int(helical[3])))

    def reconstruct(self, iclass: int, ori_size: int, tau2=None, tau2_fudge: float = 1.0, minres_map: int = 0,
                    max_iter_preweight: int = 0, normalise: float = 1.0) -> np.ndarray:
        """rb_reconstruct: BackProjector::reconstruct on the device; [ori, ori, ori] float32.  max_iter_preweight > 0: the iterative
        gridding branch (rb_reconstruct_gridding), else the default skip_gridding branch."""
        out = np.empty((ori_size,) * 3, np.float32)
        t = _f64(tau2)
        if max_iter_preweight > 0:
            capi.check(self.lib, self.lib.rb_reconstruct_gridding(self.ctx, iclass, ori_size, _ptr(t, C.c_double), 0 if t is None else len(t),
                                                                  float(tau2_fudge), int(minres_map), int(max_iter_preweight), float(normalise),
                                                                  _ptr(out, C.c_float)))
            return out
        capi.check(self.lib, self.lib.rb_reconstruct(self.ctx, iclass, ori_size, _ptr(t, C.c_double), 0 if t is None else len(t),
                                                     float(tau2_fudge), int(minres_map), _ptr(out, C.c_float)))
        return out

    def update_ssnr(self, iclass: int, ori_size: int, tau2, tau2_fudge: float = 1.0, fsc=None, avgctf2=None,
                    update_tau2_with_fsc: bool = False, is_whole_instead_of_half: bool = False):
        """rb_update_ssnr: BackProjector::updateSSNRarrays on the device accumulator.
        Returns (tau2, sigma2, data_vs_prior, fourier_coverage), each [ori_size/2 + 1] float64."""
        ns = ori_size // 2 + 1
        t = np.array(tau2, np.float64, copy=True)
        if t.shape != (ns,):
            raise ValueError(f"tau2 must have {ns} shells")
        sigma2, dvp, cov = np.zeros(ns), np.zeros(ns), np.zeros(ns)
        f, a = _f64(fsc), _f64(avgctf2)
        capi.check(self.lib, self.lib.rb_update_ssnr(self.ctx, iclass, ori_size, float(tau2_fudge), _ptr(t, C.c_double), _ptr(sigma2, C.c_double),
                                                     _ptr(dvp, C.c_double), _ptr(cov, C.c_double), _ptr(f, C.c_double), _ptr(a, C.c_double),
                                                     1 if update_tau2_with_fsc else 0, 1 if is_whole_instead_of_half else 0))
        return t, sigma2, dvp, cov

    def bp_device_tensor(self, iclass: int):
        """The interleaved (re, im, weight, 0) accumulator as a torch CUDA tensor sharing the library's memory
        (for torch.distributed.all_reduce over NCCL — replaces MlOptimiserMpi::combineAllWeightedSums)."""
        import torch
        p = C.c_void_p()
        n = C.c_size_t()
        capi.check(self.lib, self.lib.rb_bp_device_buffer(self.ctx, iclass, C.byref(p), C.byref(n)))

        class _Wrap:
            pass

        wrap = _Wrap()
        wrap.__cuda_array_interface__ = {"shape": (n.value,), "typestr": "<f4", "data": (p.value, False), "version": 2}
        return torch.as_tensor(wrap, device=f"cuda:{self.device_id}")

    def sync_all_backprojects(self):
        capi.check(self.lib, self.lib.rb_sync(self.ctx))

    # ---- per-iteration state --------------------------------------------------------------------
    def set_model(self, p: ModelParams):
        m = marshal_model(p)
        capi.check(self.lib, self.lib.rb_set_model(self.ctx, C.byref(m.struct)))
        self.model = p
        if self.sampling is not None and p.pdf_direction is not None:
            pd = _f64(p.pdf_direction)
            capi.check(self.lib, self.lib.rb_set_pdf_direction(self.ctx, _ptr(pd, C.c_double)))

    def set_sampling(self, s: Sampling):
        assert self.model is not None, "set_model first (translations are scaled by ori_size)"
        m = marshal_sampling(s)
        capi.check(self.lib, self.lib.rb_set_sampling(self.ctx, C.byref(m.struct)))
        self.sampling = s
        if self.model.pdf_direction is not None:
            pd = _f64(self.model.pdf_direction)
            assert pd.shape == (self.model.nr_classes, s.n_dir), pd.shape
            capi.check(self.lib, self.lib.rb_set_pdf_direction(self.ctx, _ptr(pd, C.c_double)))

    # ---- the E-step -----------------------------------------------------------------------------
    def _new_out(self, n_particles):
        return make_pool_out(n_particles, self.model.ori_size // 2 + 1, self.model.nr_classes, self.sampling.n_dir)

    def expectation_some_particles(self, pool: ParticlePool, skip_maximization: bool = False) -> PoolResult:
        """H2D of the pool, all E-step stages, D2H of the per-particle results (rb_estep_pool)."""
        mp = marshal_pool(pool)
        out = self._new_out(pool.n_particles)
        st = self.lib.rb_estep_pool(self.ctx, C.byref(mp.struct), C.byref(out.struct), 1 if skip_maximization else 0)
        capi.check(self.lib, st)
        return out.result

    def pool_upload(self, slot: int, pool: ParticlePool):
        mp = marshal_pool(pool)
        self._keep[("pool", slot)] = mp   # host buffers must outlive the asynchronous copy
        capi.check(self.lib, self.lib.rb_pool_upload(self.ctx, slot, C.byref(mp.struct)))
        self._keep[("pool_n", slot)] = pool.n_particles

    def pool_prepare(self, slot: int, raw: RawParticlePool, want_power: bool = True):
        """rb_pool_prepare: getFourierTransformsAndCtfs on the device; returns power_img [P, n/2+1] (or None)."""
        mp = marshal_raw_pool(raw)
        self._keep[("pool", slot)] = mp
        n = int(mp.struct.image_size)
        power = np.zeros((raw.n_particles, n // 2 + 1), np.float32) if want_power else None
        capi.check(self.lib, self.lib.rb_pool_prepare(self.ctx, slot, C.byref(mp.struct), _ptr(power, C.c_float)))
        self._keep[("pool_n", slot)] = raw.n_particles
        return power

    def debug_prepared_coarse_image(self, slot: int, particle: int, coarse_size: int):
        """Prepared coarse-window image of a particle (rb_debug_prepared_coarse_image, test hook): [nc, nc/2+1, 4]."""
        out = np.empty((coarse_size, coarse_size // 2 + 1, 4), np.float32)
        capi.check(self.lib, self.lib.rb_debug_prepared_coarse_image(self.ctx, slot, particle, _ptr(out, C.c_float)))
        return out

    def debug_coarse_eulers(self, n_dir: int, n_psi: int):
        """Coarse-pass Euler matrices built on the device by set_sampling (rb_debug_coarse_eulers, test hook): [n_dir, n_psi, 9]."""
        out = np.empty((n_dir, n_psi, 9), np.float32)
        capi.check(self.lib, self.lib.rb_debug_coarse_eulers(self.ctx, _ptr(out, C.c_float), out.size))
        return out

    def debug_prep_noise(self, n_particles: int, n: int):
        """Noise images of the last pool_prepare with noise_seed (rb_debug_prep_noise, test hook)."""
        out = np.empty((n_particles, n, n), np.float32)
        capi.check(self.lib, self.lib.rb_debug_prep_noise(self.ctx, n_particles, n, _ptr(out, C.c_float)))
        return out

    def pool_download(self, slot: int, current_size: int):
        P = self._keep[("pool_n", slot)]
        xs = current_size // 2 + 1
        F = np.empty((P, current_size, xs), np.complex64); F0 = np.empty_like(F)
        Cc = np.ones((P, current_size, xs), np.float32); xi2 = np.zeros(P, np.float64)
        capi.check(self.lib, self.lib.rb_pool_download(self.ctx, slot, _ptr(F.view(np.float32), C.c_float), _ptr(F0.view(np.float32), C.c_float),
                                                       _ptr(Cc, C.c_float), _ptr(xi2, C.c_double)))
        return F, F0, Cc, xi2

    def estep_slot(self, slot: int, skip_maximization: bool = False) -> PoolResult:
        out = self._new_out(self._keep[("pool_n", slot)])
        capi.check(self.lib, self.lib.rb_estep_slot(self.ctx, slot, C.byref(out.struct), 1 if skip_maximization else 0))
        return out.result

    def estep_slot_nocopy(self, slot: int, skip_maximization: bool = False):
        capi.check(self.lib, self.lib.rb_estep_slot_nocopy(self.ctx, slot, 1 if skip_maximization else 0))

    def estep_fetch(self, slot: int) -> PoolResult:
        out = self._new_out(self._keep[("pool_n", slot)])
        capi.check(self.lib, self.lib.rb_estep_fetch(self.ctx, slot, C.byref(out.struct)))
        return out.result

    def debug_coarse_weights(self, slot: int, particle: int) -> np.ndarray:
        """Coarse-pass weights of one particle of a slot after its E-step (rb_debug_coarse_weights, test hook)."""
        n = C.c_longlong()
        capi.check(self.lib, self.lib.rb_debug_coarse_weights(self.ctx, slot, particle, None, 0, C.byref(n)))
        w = np.empty(n.value, np.float32)
        capi.check(self.lib, self.lib.rb_debug_coarse_weights(self.ctx, slot, particle, _ptr(w, C.c_float), n.value, C.byref(n)))
        return w

    def stage_ms(self, name: str) -> float:
        return float(self.lib.rb_stage_ms(self.ctx, name.encode()))

    def timer_start(self):
        capi.check(self.lib, self.lib.rb_timer_start(self.ctx))

    def timer_stop(self) -> float:
        ms = C.c_double()
        capi.check(self.lib, self.lib.rb_timer_stop(self.ctx, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self.lib.rb_launch_count(self.ctx))

    # ---- stage-level entry points (parity tests) --------------------------------------------------
    def project(self, iclass, img_size, eulers):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        out = np.empty((e.shape[0], img_size, img_size // 2 + 1), np.complex64)
        capi.check(self.lib, self.lib.rb_project(self.ctx, iclass, img_size, _ptr(e, C.c_float), e.shape[0], _ptr(out.view(np.float32), C.c_float)))
        return out

    def diff2_coarse(self, iclass, img_size, eulers, tx, ty, re, im, corr, init=None, cc=False):
        """runDiff2KernelCoarse; cc=True: the first-iteration cross-correlation kernel (rb_diff2_cc_coarse)."""
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32); corr = np.ascontiguousarray(corr, np.float32)
        out = np.zeros((e.shape[0], len(tx)), np.float32) if init is None else np.ascontiguousarray(init, np.float32).copy()
        fn = self.lib.rb_diff2_cc_coarse if cc else self.lib.rb_diff2_coarse
        capi.check(self.lib, fn(self.ctx, iclass, img_size, _ptr(e, C.c_float), e.shape[0],
                                                      _ptr(tx, C.c_float), _ptr(ty, C.c_float), len(tx),
                                                      _ptr(re, C.c_float), _ptr(im, C.c_float), _ptr(corr, C.c_float), _ptr(out, C.c_float)))
        return out

    def gemm_tf32x3(self, A, B):
        """C = A @ B.T on the tcgen05 tensor cores with 3xTF32 splitting (rb_gemm_tf32x3); A [M, K], B [N, K] float32."""
        A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
        assert A.ndim == 2 and B.ndim == 2 and A.shape[1] == B.shape[1]
        out = np.empty((A.shape[0], B.shape[0]), np.float32)
        capi.check(self.lib, self.lib.rb_gemm_tf32x3(self.ctx, _ptr(A, C.c_float), _ptr(B, C.c_float), A.shape[0], B.shape[0], A.shape[1],
                                                     _ptr(out, C.c_float)))
        return out

    def diff2_cc_fine(self, iclass, img_size, eulers, tx, ty, re, im, corr, rot_idx, trans_idx, job_idx, job_num):
        """runDiff2KernelFine with the first-iteration cross-correlation criterion (rb_diff2_cc_fine)."""
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32); corr = np.ascontiguousarray(corr, np.float32)
        ri = np.ascontiguousarray(rot_idx, np.uint64); ti = np.ascontiguousarray(trans_idx, np.uint64)
        ji = np.ascontiguousarray(job_idx, np.uint64); jn = np.ascontiguousarray(job_num, np.uint64)
        out = np.zeros(len(ri), np.float32)
        capi.check(self.lib, self.lib.rb_diff2_cc_fine(self.ctx, iclass, img_size, _ptr(e, C.c_float), e.shape[0],
                                                       _ptr(tx, C.c_float), _ptr(ty, C.c_float), len(tx),
                                                       _ptr(re, C.c_float), _ptr(im, C.c_float), _ptr(corr, C.c_float),
                                                       _ptr(ri, C.c_uint64), _ptr(ti, C.c_uint64), _ptr(ji, C.c_uint64), _ptr(jn, C.c_uint64),
                                                       len(ji), _ptr(out, C.c_float), len(ri)))
        return out

    def diff2_fine(self, iclass, img_size, eulers, tx, ty, re, im, corr, sum_init, rot_idx, trans_idx, job_idx, job_num):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32); corr = np.ascontiguousarray(corr, np.float32)
        ri = np.ascontiguousarray(rot_idx, np.uint64); ti = np.ascontiguousarray(trans_idx, np.uint64)
        ji = np.ascontiguousarray(job_idx, np.uint64); jn = np.ascontiguousarray(job_num, np.uint64)
        out = np.zeros(len(ri), np.float32)
        capi.check(self.lib, self.lib.rb_diff2_fine(self.ctx, iclass, img_size, _ptr(e, C.c_float), e.shape[0],
                                                    _ptr(tx, C.c_float), _ptr(ty, C.c_float), len(tx),
                                                    _ptr(re, C.c_float), _ptr(im, C.c_float), _ptr(corr, C.c_float), float(sum_init),
                                                    _ptr(ri, C.c_uint64), _ptr(ti, C.c_uint64), _ptr(ji, C.c_uint64), _ptr(jn, C.c_uint64),
                                                    len(ji), _ptr(out, C.c_float), len(ri)))
        return out

    def convert_weights(self, diff2, pdf_o, pdf_oz, pdf_t, pdf_tz, adaptive_fraction=0.999, maxsig=0, filter_zero=True):
        w = np.ascontiguousarray(diff2, np.float32).copy()
        no, nt = w.shape
        po = np.ascontiguousarray(pdf_o, np.float32); pt = np.ascontiguousarray(pdf_t, np.float32)
        oz = np.ascontiguousarray(pdf_oz, np.uint8); tz = np.ascontiguousarray(pdf_tz, np.uint8)
        sig = np.zeros(w.shape, np.uint8)
        out = capi.rb_weights_out()
        capi.check(self.lib, self.lib.rb_convert_weights(self.ctx, _ptr(w, C.c_float), no, nt, _ptr(po, C.c_float), _ptr(oz, C.c_ubyte),
                                                         _ptr(pt, C.c_float), _ptr(tz, C.c_ubyte), adaptive_fraction, maxsig,
                                                         int(filter_zero), _ptr(sig, C.c_ubyte), C.byref(out)))
        return w, sig, out

    def wavg(self, iclass, img_size, eulers, tx, ty, re, im, weights, ctfs, weight_norm, sig_w):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32)
        w = np.ascontiguousarray(weights, np.float32); c = np.ascontiguousarray(ctfs, np.float32)
        np_ = img_size * (img_size // 2 + 1)
        parts = np.zeros(np_, np.float32); AA = np.zeros(np_, np.float32); XA = np.zeros(np_, np.float32)
        capi.check(self.lib, self.lib.rb_wavg(self.ctx, iclass, img_size, _ptr(e, C.c_float), e.shape[0], _ptr(tx, C.c_float), _ptr(ty, C.c_float), len(tx),
                                              _ptr(re, C.c_float), _ptr(im, C.c_float), _ptr(w, C.c_float), _ptr(c, C.c_float),
                                              float(weight_norm), float(sig_w), _ptr(parts, C.c_float), _ptr(AA, C.c_float), _ptr(XA, C.c_float)))
        return parts, AA, XA

    def backproject(self, iclass, img_size, eulers, tx, ty, re, im, weights, minvsigma2, ctfs, weight_norm, sig_w):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        tx = np.ascontiguousarray(tx, np.float32); ty = np.ascontiguousarray(ty, np.float32)
        re = np.ascontiguousarray(re, np.float32); im = np.ascontiguousarray(im, np.float32)
        w = np.ascontiguousarray(weights, np.float32); c = np.ascontiguousarray(ctfs, np.float32); mi = np.ascontiguousarray(minvsigma2, np.float32)
        capi.check(self.lib, self.lib.rb_backproject(self.ctx, iclass, img_size, _ptr(e, C.c_float), e.shape[0], _ptr(tx, C.c_float), _ptr(ty, C.c_float), len(tx),
                                                     _ptr(re, C.c_float), _ptr(im, C.c_float), _ptr(w, C.c_float), _ptr(mi, C.c_float), _ptr(c, C.c_float),
                                                     float(weight_norm), float(sig_w)))

    def backproject_posed(self, iclass, img_size, F2D, Fctf, eulers):
        """Reconstructor::backprojectOneParticle for a batch (rb_backproject_posed): F2D [n, s, s/2+1] complex64 (numpy or a
        pinned torch tensor viewed as float32), Fctf [n, s, s/2+1] float32, eulers [n, 9] inverted matrices."""
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        capi.check(self.lib, self.lib.rb_backproject_posed(self.ctx, iclass, img_size, e.shape[0], _fptr(F2D), _fptr(Fctf), _ptr(e, C.c_float)))

    def backproject_posed_raw(self, iclass, images, eulers, shift=None, ctf=None, pixel_size=1.0, ctf_premultiplied=False):
        """rb_backproject_posed_raw: real-space images [n, s, s] float32 (numpy or pinned torch tensor); the transform, centring,
        origin shift, CTF and DC removal of Reconstructor::backprojectOneParticle run on the device.  ctf: dict with defU, defV,
        defAngle (per image) and kV, Cs, Q0 (scalars or per optics group), optional Bfac, scale, phase_shift, optics_group."""
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        keep = [e]
        st = capi.rb_posed_raw()
        st.n_images = e.shape[0]
        st.image_size = int(images.shape[-1])
        st.images = _fptr(images)
        st.eulers = _ptr(e, C.c_float)

        def dbl(a):
            if a is None:
                return _ptr(None, C.c_double)
            a = np.ascontiguousarray(np.atleast_1d(a), np.float64)
            keep.append(a)
            return _ptr(a, C.c_double)

        st.shift = dbl(None if shift is None else np.asarray(shift, np.float64).reshape(-1, 2))
        if ctf is not None:
            st.ctf_defU, st.ctf_defV, st.ctf_defAngle = dbl(ctf["defU"]), dbl(ctf["defV"]), dbl(ctf["defAngle"])
            st.ctf_Bfac, st.ctf_scale, st.ctf_phase_shift = dbl(ctf.get("Bfac")), dbl(ctf.get("scale")), dbl(ctf.get("phase_shift"))
            st.og_kV, st.og_Cs, st.og_Q0 = dbl(ctf["kV"]), dbl(ctf["Cs"]), dbl(ctf["Q0"])
            if ctf.get("optics_group") is not None:
                og = np.ascontiguousarray(ctf["optics_group"], np.int32)
                keep.append(og)
                st.optics_group = _ptr(og, C.c_int)
        st.pixel_size = float(pixel_size)
        st.ctf_premultiplied = int(ctf_premultiplied)
        capi.check(self.lib, self.lib.rb_backproject_posed_raw(self.ctx, iclass, C.byref(st)))

    def bp_posed_stage(self, img_size, F2D, Fctf, eulers):
        e = np.ascontiguousarray(eulers, np.float32).reshape(-1, 9)
        capi.check(self.lib, self.lib.rb_bp_posed_stage(self.ctx, img_size, e.shape[0], _fptr(F2D), _fptr(Fctf), _ptr(e, C.c_float)))

    def bp_posed_run(self, iclass):
        capi.check(self.lib, self.lib.rb_bp_posed_run(self.ctx, iclass))
