"""Time rb_reconstruct / rb_reconstruct_gridding at the 256-px sizes (515^3 padded volume) on a posed back-projection."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from relion_b200.estep import MlDeviceBundle
from relion_b200 import synth

n, r_max, pf, count = 256, 128, 2.0, 256
xs = n // 2 + 1
rng = np.random.default_rng(3)
pad = synth.pad_size_for(r_max, pf)
dev = MlDeviceBundle(0)
dev.bp_init(0, (pad, pad, pad // 2 + 1), r_max, pf)
F = (rng.standard_normal((count, n, xs)) + 1j * rng.standard_normal((count, n, xs))).astype(np.complex64)
W = np.ones((count, n, xs), np.float32)
eul = synth.inverse_euler_f32(rng.uniform(-180, 180, count), np.degrees(np.arccos(rng.uniform(-1, 1, count))), rng.uniform(0, 360, count))
dev.backproject_posed(0, n, F, W, eul)
tau2 = 1e-3 / (1.0 + np.arange(n // 2 + 1)) ** 2
for it in (0, 10):
    t0 = time.perf_counter()
    vol = dev.reconstruct(0, n, tau2=tau2, tau2_fudge=2.0, minres_map=2, max_iter_preweight=it)
    print("max_iter_preweight", it, "seconds", round(time.perf_counter() - t0, 3), "finite", bool(np.isfinite(vol).all()), "max", float(np.abs(vol).max()))
