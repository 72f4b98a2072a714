"""Generates tests/golden/estep_reference_*.npz from the REFERENCE's own kernels.

Run in the dev container (needs /root/reference to build oracle/_ref/librefkernels.so):

    python tests/golden/make_golden.py

The fixtures pin the oracle: tests/test_oracle.py checks the restated kernels (oracle/port_kernels.cpp)
against them on every machine, including ones where /root/reference does not exist.  Inputs are not
stored — they are regenerated from the seeds recorded in each file by relion_b200.workload.make_workload.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bindings import Oracle, Projector, Backprojector, build  # noqa: E402
from relion_b200.workload import make_workload  # noqa: E402

CASES = {
    "global_k2": dict(ori_size=24, healpix_order=1, n_particles=4, nr_classes=2, seed=5, snr=0.3),
    "local_k1": dict(ori_size=32, healpix_order=2, n_particles=3, nr_classes=1, seed=6, snr=0.2, local_search=True),
    "window_k1": dict(ori_size=40, current_size=28, healpix_order=1, n_particles=3, nr_classes=1, seed=8, snr=0.3),
    # first-iteration cross-correlation criterion (--firstiter_cc): the reference's diff2_CC_coarse / diff2_CC_fine kernels
    "cc_global_k1": dict(ori_size=24, healpix_order=1, n_particles=4, nr_classes=1, seed=9, snr=0.3, do_cc=True),
    "cc_window_k1": dict(ori_size=40, current_size=28, healpix_order=1, n_particles=3, nr_classes=1, seed=10, snr=0.1, do_cc=True),
    # MBL / MBR of the pool (cpu_kernel_make_eulers_3D<true, doL, doR>, generateEulerMatrices(..., L, R)): an anisotropic
    # magnification matrix with a body rotation, and an optics group whose box differs from the model's (scale difference)
    "mag_local_k1": dict(ori_size=32, healpix_order=2, n_particles=3, nr_classes=1, seed=21, snr=0.2, local_search=True,
                         mat_left=[[1.03, 0.012, 0.0], [-0.008, 0.96, 0.0], [0.0, 0.0, 1.0]],
                         mat_right=[[0.9553364891, -0.2955202067, 0.0], [0.2955202067, 0.9553364891, 0.0], [0.0, 0.0, 1.0]]),
    "box_global_k1": dict(ori_size=32, ref_box=24, healpix_order=1, n_particles=4, nr_classes=1, seed=22, snr=0.3),
}


def run_case(kind, kw):
    wl = make_workload(**kw)
    o = Oracle(kind)
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
    st, out, dbg = o.estep_pool(wl.model, wl.sampling, refs, bps, wl.pool, num_threads=1, debug_particle=0)
    assert st == 0
    d = dict(particles=out.particles, shells=out.wsum_sigma2_noise, pdf_direction=out.wsum_pdf_direction,
             pdf_class=out.wsum_pdf_class,
             coarse_diff2=dbg["coarse_diff2"][::7].copy(), coarse_weights=dbg["coarse_weights"][::7].copy(),
             coarse_significant=dbg["coarse_significant"].copy(),
             fine_ihidden_over=dbg["fine_ihidden_over"], fine_diff2=dbg["fine_diff2"], fine_weights=dbg["fine_weights"],
             wdiff2s_parts=dbg["wdiff2s_parts"], wdiff2s_AA=dbg["wdiff2s_AA"], wdiff2s_XA=dbg["wdiff2s_XA"])
    for k, b in enumerate(bps):
        for nm in ("real", "imag", "weight"):
            a = getattr(b, nm)
            d[f"bp{k}_{nm}_sum"] = np.array(a.astype(np.float64).sum())
            d[f"bp{k}_{nm}_abs"] = np.array(np.abs(a).astype(np.float64).sum())
            d[f"bp{k}_{nm}_sample"] = a.reshape(-1)[::97].copy()
    return d


if __name__ == "__main__":
    build(ref=True)
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        d = run_case("reference", kw)
        np.savez_compressed(os.path.join(here, f"estep_reference_{name}.npz"), **d)
        print(name, "written", {k: v.shape for k, v in d.items() if hasattr(v, "shape") and v.ndim}.__len__(), "arrays")
