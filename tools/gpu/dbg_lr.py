import os, sys, numpy as np
sys.path.insert(0, ".")
from relion_b200.workload import make_workload
from relion_b200.estep import MlDeviceBundle
from oracle.bindings import Oracle, Projector, Backprojector
kw = dict(ori_size=40, ref_box=32, n_particles=4, seed=120, snr=0.2, healpix_order=1)
wl = make_workload(**kw)
dev = MlDeviceBundle(0)
dev.set_model(wl.model); dev.set_sampling(wl.sampling)
for k, v in enumerate(wl.refs):
    dev.set_reference(k, v, wl.r_max, wl.padding_factor)
    dev.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)
res = dev.expectation_some_particles(wl.pool)
o = Oracle("port")
refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
st, ores, _ = o.estep_pool(wl.model, wl.sampling, refs, bps, wl.pool, num_threads=0)
for f in ("min_diff2_coarse", "sum_weight_coarse", "min_diff2", "sum_weight", "dLL_nolog", "best_ihidden_over", "nr_significant_coarse"):
    print(f, res.particles[f], ores.particles[f])
print("r_max", wl.r_max, "refs", wl.refs[0].shape, "model", wl.model.ori_size, wl.model.current_size, wl.model.coarse_size)
