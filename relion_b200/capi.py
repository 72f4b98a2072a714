"""ctypes binding of the C-ABI in include/relion_b200.h (the drop-in boundary).

The shared library is built in-tree (relion_b200/librelion_b200.so, see relion_b200/csrc/Makefile and
__graft_entry__.build()).  There is no Python or CPU fallback: if the library is missing, loading
fails loudly, and rb_ctx_create() fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librelion_b200.so")

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)
c_u64_p = C.POINTER(C.c_uint64)

RB_OK = 0
RB_ERR_CUDA = -1
RB_ERR_ARG = -2
RB_ERR_STATE = -3
RB_ERR_CAPACITY = -4
RB_ERR_TRANSLIM = -5
RB_ERR_NO_SIGNIFICANT = -6
RB_ERR_SUMWEIGHT_ZERO = -7
RB_ERR_PMAX = -8


class rb_sampling(C.Structure):
    _fields_ = [
        ("n_dir", C.c_int), ("n_psi", C.c_int),
        ("rot", c_double_p), ("tilt", c_double_p), ("psi", c_double_p),
        ("n_over_rot", C.c_int),
        ("over_rot", c_double_p), ("over_tilt", c_double_p), ("over_psi", c_double_p),
        ("n_trans", C.c_int),
        ("trans_x", c_double_p), ("trans_y", c_double_p),
        ("n_over_trans", C.c_int),
        ("over_trans_x", c_double_p), ("over_trans_y", c_double_p),
    ]


class rb_model(C.Structure):
    _fields_ = [
        ("nr_classes", C.c_int), ("ori_size", C.c_int), ("coarse_size", C.c_int), ("current_size", C.c_int),
        ("pixel_size", C.c_double),
        ("nr_optics_groups", C.c_int), ("sigma2_noise", c_double_p),
        ("nr_groups", C.c_int), ("scale_correction", c_double_p),
        ("pdf_class", c_double_p), ("pdf_direction", c_double_p), ("data_vs_prior_class", c_double_p),
        ("sigma2_offset", C.c_double), ("offset_range", C.c_double), ("sigma2_fudge", C.c_double),
        ("adaptive_fraction", C.c_double), ("maximum_significants", C.c_int),
        ("do_ctf_correction", C.c_int), ("refs_are_ctf_corrected", C.c_int), ("do_scale_correction", C.c_int),
        ("do_map", C.c_int), ("ctf_premultiplied", C.c_int), ("bp_circle_bound", C.c_int),
        ("do_cc", C.c_int),
        ("prior_offset_class", c_double_p),
        ("do_grad", C.c_int),
        ("do_skip_rotate", C.c_int),
        ("ref_max_r", C.c_int),
    ]


class rb_particles(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int),
        ("Fimg", c_float_p), ("Fimg_nomask", c_float_p), ("Fctf", c_float_p),
        ("group_id", c_int_p), ("optics_group", c_int_p),
        ("highres_Xi2", c_double_p), ("old_offset", c_double_p), ("prior_offset", c_double_p),
        ("dir_off", c_int_p), ("dir_idx", c_int_p), ("dir_prior", c_double_p),
        ("psi_off", c_int_p), ("psi_idx", c_int_p), ("psi_prior", c_double_p),
        ("bp_offset", c_int_p),
        ("pre_shift", c_double_p),
        ("mat_left", c_double_p), ("mat_right", c_double_p),
    ]


class rb_raw_particles(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int), ("image_size", C.c_int),
        ("images", c_float_p), ("norm_factor", c_double_p), ("old_offset", c_double_p), ("prior_offset", c_double_p),
        ("group_id", c_int_p), ("optics_group", c_int_p),
        ("ctf_defU", c_double_p), ("ctf_defV", c_double_p), ("ctf_defAngle", c_double_p),
        ("ctf_Bfac", c_double_p), ("ctf_scale", c_double_p), ("ctf_phase_shift", c_double_p),
        ("og_kV", c_double_p), ("og_Cs", c_double_p), ("og_Q0", c_double_p),
        ("mask_radius", C.c_double), ("width_mask_edge", C.c_double),
        ("dir_off", c_int_p), ("dir_idx", c_int_p), ("dir_prior", c_double_p),
        ("psi_off", c_int_p), ("psi_idx", c_int_p), ("psi_prior", c_double_p),
        ("bp_offset", c_int_p),
        ("og_fourier_factor", c_float_p),
        ("noise_seed", C.POINTER(C.c_int64)),
        ("mat_left", c_double_p), ("mat_right", c_double_p),
        ("pre_shift", c_double_p),
        ("noise_sigma2", c_double_p),
    ]


class rb_posed_raw(C.Structure):
    _fields_ = [
        ("n_images", C.c_int), ("image_size", C.c_int),
        ("images", c_float_p), ("eulers", c_float_p), ("shift", c_double_p),
        ("ctf_defU", c_double_p), ("ctf_defV", c_double_p), ("ctf_defAngle", c_double_p),
        ("ctf_Bfac", c_double_p), ("ctf_scale", c_double_p), ("ctf_phase_shift", c_double_p),
        ("optics_group", c_int_p),
        ("og_kV", c_double_p), ("og_Cs", c_double_p), ("og_Q0", c_double_p),
        ("pixel_size", C.c_double), ("ctf_premultiplied", C.c_int),
    ]


class rb_particle_out(C.Structure):
    _fields_ = [
        ("best_ihidden_over", C.c_int64),
        ("best_class", C.c_int), ("best_idir", C.c_int), ("best_ipsi", C.c_int),
        ("best_iover_rot", C.c_int), ("best_itrans", C.c_int), ("best_iover_trans", C.c_int),
        ("nr_significant_coarse", C.c_int), ("n_fine_orient", C.c_int), ("n_fine_samples", C.c_int),
        ("min_diff2_coarse", C.c_float), ("sum_weight_coarse", C.c_float), ("significant_weight_coarse", C.c_float),
        ("min_diff2", C.c_float), ("max_weight", C.c_float), ("sum_weight", C.c_float),
        ("significant_weight", C.c_float), ("pmax", C.c_float), ("n_bp_orient", C.c_int),
        ("dLL_nolog", C.c_double), ("wsum_norm_correction", C.c_double),
        ("wsum_XA", C.c_double), ("wsum_AA", C.c_double), ("sumw", C.c_double), ("wsum_sigma2_offset", C.c_double),
    ]


class rb_pool_out(C.Structure):
    _fields_ = [
        ("particles", C.POINTER(rb_particle_out)),
        ("wsum_sigma2_noise", c_float_p),
        ("wsum_pdf_direction", c_double_p),
        ("wsum_pdf_class", c_double_p),
        ("wsum_prior_offset_class", c_double_p),
    ]


class rb_weights_out(C.Structure):
    _fields_ = [
        ("min_diff2", C.c_float), ("max_weight", C.c_float), ("max_index", C.c_int64),
        ("sum_weight", C.c_float), ("significant_weight", C.c_float),
        ("nr_significant", C.c_int), ("n_nonzero", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/relion_b200.h declares
PROTOTYPES = {
    "rb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rb_ctx_destroy": (None, [C.c_void_p]),
    "rb_last_error": (C.c_char_p, []),
    "rb_version": (C.c_int, []),
    "rb_sync": (C.c_int, [C.c_void_p]),
    "rb_launch_count": (C.c_longlong, [C.c_void_p]),
    "rb_stage_ms": (C.c_double, [C.c_void_p, C.c_char_p]),
    "rb_timer_start": (C.c_int, [C.c_void_p]),
    "rb_timer_stop": (C.c_int, [C.c_void_p, c_double_p]),
    "rb_set_reference": (C.c_int, [C.c_void_p, C.c_int, c_double_p] + [C.c_int] * 6 + [C.c_double]),
    "rb_set_reference_f32": (C.c_int, [C.c_void_p, C.c_int, c_float_p] + [C.c_int] * 6 + [C.c_double]),
    "rb_set_reference_from_map": (C.c_int, [C.c_void_p, C.c_int, c_float_p, C.c_int, C.c_int, C.c_double, c_double_p]),
    "rb_bp_init": (C.c_int, [C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_double]),
    "rb_bp_clear": (C.c_int, [C.c_void_p, C.c_int]),
    "rb_bp_get": (C.c_int, [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p]),
    "rb_bp_symmetrise": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int]),
    "rb_bp_symmetrise_helical": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]),
    "rb_reconstruct": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_double, C.c_int, c_float_p]),
    "rb_reconstruct_gridding": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, c_float_p]),
    "rb_update_ssnr": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p,
                                 c_double_p, c_double_p, C.c_int, C.c_int]),
    "rb_bp_device_buffer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "rb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rb_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "rb_comm_create_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    "rb_comm_destroy": (None, [C.c_void_p]),
    "rb_comm_size": (C.c_int, [C.c_void_p]),
    "rb_comm_rank": (C.c_int, [C.c_void_p]),
    "rb_comm_handle": (C.c_void_p, [C.c_void_p]),
    "rb_comm_group_start": (C.c_int, []),
    "rb_comm_group_end": (C.c_int, []),
    "rb_bp_allreduce": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb_bp_allreduce_nccl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "rb_wsum_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_size_t]),
    "rb_set_sampling": (C.c_int, [C.c_void_p, C.POINTER(rb_sampling)]),
    "rb_set_model": (C.c_int, [C.c_void_p, C.POINTER(rb_model)]),
    "rb_set_pdf_direction": (C.c_int, [C.c_void_p, c_double_p]),
    "rb_estep_pool": (C.c_int, [C.c_void_p, C.POINTER(rb_particles), C.POINTER(rb_pool_out), C.c_uint]),
    "rb_pool_upload": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(rb_particles)]),
    "rb_pool_prepare": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(rb_raw_particles), c_float_p]),
    "rb_pool_download": (C.c_int, [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p, c_double_p]),
    "rb_mrc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "rb_mrc_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [c_float_p]),
    "rb_mrc_read_images": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.c_int, c_float_p]),
    "rb_mrc_close": (None, [C.c_void_p]),
    "rb_mrc_write": (C.c_int, [C.c_char_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_float]),
    "rb_feed_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "rb_feed_submit": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int)]),
    "rb_feed_wait": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(c_float_p)]),
    "rb_feed_release": (C.c_int, [C.c_void_p, C.c_int]),
    "rb_feed_destroy": (None, [C.c_void_p]),
    "rb_estep_slot": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(rb_pool_out), C.c_uint]),
    "rb_estep_slot_nocopy": (C.c_int, [C.c_void_p, C.c_int, C.c_uint]),
    "rb_estep_fetch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(rb_pool_out)]),
    "rb_debug_coarse_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_longlong, C.POINTER(C.c_longlong)]),
    "rb_debug_prep_noise": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "rb_debug_coarse_eulers": (C.c_int, [C.c_void_p, c_float_p, C.c_longlong]),
    "rb_debug_prepared_coarse_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "rb_project": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p]),
    "rb_diff2_coarse": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, c_float_p, C.c_int,
                                  c_float_p, c_float_p, c_float_p, c_float_p]),
    "rb_diff2_cc_coarse": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, c_float_p, C.c_int,
                                     c_float_p, c_float_p, c_float_p, c_float_p]),
    "rb_diff2_cc_fine": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, c_float_p, C.c_int,
                                   c_float_p, c_float_p, c_float_p,
                                   c_u64_p, c_u64_p, c_u64_p, c_u64_p, C.c_int, c_float_p, C.c_int]),
    "rb_gemm_tf32x3": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, c_float_p]),
    "rb_diff2_fine": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, c_float_p, C.c_int,
                                c_float_p, c_float_p, c_float_p, C.c_float,
                                c_u64_p, c_u64_p, c_u64_p, c_u64_p, C.c_int, c_float_p, C.c_int]),
    "rb_convert_weights": (C.c_int, [C.c_void_p, c_float_p, C.c_int64, C.c_int, c_float_p, c_ubyte_p, c_float_p, c_ubyte_p,
                                     C.c_double, C.c_int, C.c_int, c_ubyte_p, C.POINTER(rb_weights_out)]),
    "rb_wavg": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, c_float_p, C.c_int,
                          c_float_p, c_float_p, c_float_p, c_float_p, C.c_float, C.c_float,
                          c_float_p, c_float_p, c_float_p]),
    "rb_backproject": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, c_float_p, C.c_int,
                                 c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, C.c_float, C.c_float]),
    "rb_backproject_posed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p]),
    "rb_backproject_posed_raw": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(rb_posed_raw)]),
    "rb_bp_posed_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p, c_float_p, c_float_p]),
    "rb_bp_posed_run": (C.c_int, [C.c_void_p, C.c_int]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen the in-tree library and bind every prototype.  Raises if it was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"relion_b200: native library {p} not found. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class RelionB200Error(RuntimeError):
    """Raised for a non-zero rb_status (the C++ adapter throws RelionError in the same place)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[rb_status {status}] {message}")
        self.status = status


def check(lib, status: int):
    if status != RB_OK:
        raise RelionB200Error(status, lib.rb_last_error().decode("utf-8", "replace"))
