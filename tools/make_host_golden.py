#!/usr/bin/env python
"""Writes tests/golden/host_golden.npz: outputs of the REFERENCE's own host classes (oracle/_ref/librefrecon.so, built from
/root/reference by `make -C oracle refrecon`) on one small seeded case, for tests/test_reference_host.py on machines without
the reference tree.  Run from the repo root in the dev container:  python tools/make_host_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refhost          # noqa: E402
from relion_b200 import synth       # noqa: E402


def main():
    assert refhost.available(), "build oracle/_ref/librefrecon.so first (make -C oracle refrecon)"
    ori, cur, pf, n_img = 16, 16, 2.0, 24
    rng = np.random.default_rng(7)
    vol = synth.make_phantom(ori, n_blobs=8, seed=7)
    ft, _, r_max, _ = refhost.ft_map(vol, cur, pf, data_dim=2)
    rot, tilt, psi = rng.uniform(0, 360, n_img), np.degrees(np.arccos(rng.uniform(-1, 1, n_img))), rng.uniform(0, 360, n_img)
    A = synth.inverse_euler_f32(rot, tilt, psi).reshape(n_img, 3, 3).astype(np.float64)
    W = rng.uniform(0.2, 1.0, (n_img, cur, cur // 2 + 1)).astype(np.float32)
    iy = np.arange(cur); ky = np.where(iy < cur // 2 + 1, iy, iy - cur)[:, None]; kx = np.arange(cur // 2 + 1)[None, :]
    W = (W * ((kx * kx + ky * ky) < r_max * r_max)).astype(np.float32)       # no weight exactly on |k| = r_max (rounding-dependent there)
    F = (np.stack([synth.project_numpy(ft, r_max, pf, A[i], cur) for i in range(n_img)]) * W).astype(np.complex64)
    re, im, w = refhost.backproject(F.astype(np.complex128), np.ascontiguousarray(np.transpose(A, (0, 2, 1))), W.astype(np.float64), ori, cur, pf)   # A: inverse matrices
    ns = ori // 2 + 1
    tau2 = np.linspace(4.0, 0.05, ns)
    fsc = np.clip(np.linspace(1.0, -0.05, ns), -1, 1)
    recon = refhost.reconstruct(re, im, w, ori, cur, pf, tau2=tau2, tau2_fudge=2.0, minres_map=2)
    R = refhost.sym_matrices("D2")
    sre, sim, sw = refhost.symmetrise(re, im, w, ori, cur, "D2", pf)
    t2, s2, dvp, cov = refhost.update_ssnr(w, ori, cur, pf, 2.0, tau2, fsc=fsc, update_tau2_with_fsc=True)
    raw = rng.standard_normal((32, 32)).astype(np.float32)
    raw_ft = refhost.image_ft(raw, 20)
    raw_masked, _ = refhost.prep_soft_mask(raw, 11.0, 3.0)
    out = os.path.join(ROOT, "tests", "golden", "host_golden.npz")
    np.savez_compressed(out, vol=vol, ft_current=cur, ft_data=ft, r_max=r_max, img_F=F, img_W=W, img_A=A.astype(np.float32),
                        bp_re=re, bp_im=im, bp_w=w, tau2=tau2, fsc=fsc, recon=recon, sym_R=R, sym_re=sre, sym_im=sim, sym_w=sw,
                        ssnr_tau2=t2, ssnr_sigma2=s2, ssnr_dvp=dvp, ssnr_cov=cov, raw_img=raw, raw_cs=20, raw_ft=raw_ft, raw_masked=raw_masked)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
