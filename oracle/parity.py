"""TEST INFRASTRUCTURE ONLY - parity helpers shared by the GPU tests and bench.py's in-run `parity` block.

classify_significance
    north_star: "the significant-pose counts agree exactly whenever the reference's diff2 are identical".  The CUDA
    selector applies the reference's rule (thresholdIdx = first i with cum[i] <= (1 - f) * sum < cum[i + 1] over the
    ascending-sorted non-zero weights, /root/reference/src/acc/acc_helper_functions.h:226-232 after sortOnHost / scanOnHost,
    src/acc/utilities.h:383-398) in exact arithmetic, while the reference scans in fp32, in an order that depends on the
    build (sequential loop, `omp simd inscan`, or CUB's tree on the GPU).  This function feeds the GPU's OWN weights to the
    reference's rule (oracle `exact=False`: the sequential fp32 scan of the ALTCPU build) and classifies every particle:
      "equal"          same count
      "rounding_edge"  counts differ, and every cumulative sum between the two threshold indices lies within the
                       worst-case rounding error of an fp32 running sum of that length around the threshold: the two
                       rules differ only by the rounding of the scan, any summation order could land on either side
      "mismatch"       anything else: a real disagreement, tests fail on it
"""
from __future__ import annotations

import numpy as np


def classify_one(weights: np.ndarray, nr_sig_gpu: int, oracle, adaptive_fraction: float, maxsig: int = 0):
    """Returns (label, nr_sig_reference_rule)."""
    w = np.ascontiguousarray(weights, np.float32).ravel()
    ref = oracle.significance(w, adaptive_fraction, maxsig, True, exact=False)
    n_ref = ref["n_filtered"] - ref["threshold_idx"]
    if n_ref == nr_sig_gpu:
        return "equal", n_ref
    nz = np.sort(w[w > 0].astype(np.float64))
    n = nz.size
    if n == 0:
        return "mismatch", n_ref
    cum = np.cumsum(nz)
    total = cum[-1]
    thr = (1.0 - adaptive_fraction) * total
    i_gpu, i_ref = n - nr_sig_gpu, n - n_ref
    lo, hi = sorted((i_gpu, i_ref))
    if lo < 0 or hi > n:
        return "mismatch", n_ref
    eps = 2.0 ** -24
    # worst-case error of an fp32 running sum of i + 1 terms, plus the threshold's own error (it derives from the fp32 total)
    idx = np.arange(max(lo - 1, 0), min(hi + 1, n))
    bound = (idx + 1) * eps * cum[idx] + (1.0 - adaptive_fraction) * n * eps * total + eps * thr
    if np.all(np.abs(cum[idx] - thr) <= bound + 1e-300):
        return "rounding_edge", n_ref
    return "mismatch", n_ref


def classify_significance(device, slot: int, result_particles, oracle, adaptive_fraction: float, maxsig: int = 0, particles=None):
    """Classify every particle of a slot (or the listed ones).  Returns dict(equal=, rounding_edge=, mismatch=, labels=[...])."""
    P = len(result_particles)
    sel = range(P) if particles is None else particles
    out = {"equal": 0, "rounding_edge": 0, "mismatch": 0, "labels": []}
    for p in sel:
        w = device.debug_coarse_weights(slot, p)
        label, _ = classify_one(w, int(result_particles["nr_significant_coarse"][p]), oracle, adaptive_fraction, maxsig)
        out[label] += 1
        out["labels"].append(label)
    return out


def pool_range(pool, a: int, b: int, current_size: int):
    """Particles [a, b) of a ParticlePool (per-particle prior lists cut and re-based accordingly)."""
    from relion_b200.estep import ParticlePool
    as_np = lambda x, dt: np.asarray(x.numpy() if hasattr(x, "numpy") else x).view(dt) if x is not None else None
    P = int(pool.group_id.shape[0])
    size, xs = current_size, current_size // 2 + 1
    F = as_np(pool.Fimg, np.complex64).reshape(P, size, xs)
    F0 = as_np(pool.Fimg_nomask, np.complex64).reshape(P, size, xs)
    Cc = as_np(pool.Fctf, np.float32)
    sub = ParticlePool(Fimg=np.ascontiguousarray(F[a:b]), Fimg_nomask=np.ascontiguousarray(F0[a:b]),
                       Fctf=None if Cc is None else np.ascontiguousarray(Cc.reshape(P, size, xs)[a:b]),
                       group_id=pool.group_id[a:b], optics_group=pool.optics_group[a:b], highres_Xi2=pool.highres_Xi2[a:b],
                       old_offset=pool.old_offset[a:b], prior_offset=pool.prior_offset[a:b])
    if pool.dir_off is not None:
        d0, d1, p0, p1 = pool.dir_off[a], pool.dir_off[b], pool.psi_off[a], pool.psi_off[b]
        sub.dir_off = (pool.dir_off[a:b + 1] - d0).astype(np.int32); sub.psi_off = (pool.psi_off[a:b + 1] - p0).astype(np.int32)
        sub.dir_idx = pool.dir_idx[d0:d1]; sub.dir_prior = pool.dir_prior[d0:d1]
        sub.psi_idx = pool.psi_idx[p0:p1]; sub.psi_prior = pool.psi_prior[p0:p1]
    return sub


def pool_slice(pool, n: int, current_size: int):
    """The first n particles of a ParticlePool (per-particle prior lists cut accordingly)."""
    from relion_b200.estep import ParticlePool
    as_np = lambda a, dt: np.asarray(a.numpy() if hasattr(a, "numpy") else a).view(dt) if a is not None else None
    P = int(pool.group_id.shape[0])
    size, xs = current_size, current_size // 2 + 1
    F = as_np(pool.Fimg, np.complex64).reshape(P, size, xs)
    F0 = as_np(pool.Fimg_nomask, np.complex64).reshape(P, size, xs)
    Cc = as_np(pool.Fctf, np.float32)
    sub = ParticlePool(Fimg=np.ascontiguousarray(F[:n]), Fimg_nomask=np.ascontiguousarray(F0[:n]),
                       Fctf=None if Cc is None else np.ascontiguousarray(Cc.reshape(P, size, xs)[:n]),
                       group_id=pool.group_id[:n], optics_group=pool.optics_group[:n], highres_Xi2=pool.highres_Xi2[:n],
                       old_offset=pool.old_offset[:n], prior_offset=pool.prior_offset[:n])
    if pool.dir_off is not None:
        sub.dir_off = pool.dir_off[:n + 1]; sub.psi_off = pool.psi_off[:n + 1]
        sub.dir_idx = pool.dir_idx[:pool.dir_off[n]]; sub.dir_prior = pool.dir_prior[:pool.dir_off[n]]
        sub.psi_idx = pool.psi_idx[:pool.psi_off[n]]; sub.psi_prior = pool.psi_prior[:pool.psi_off[n]]
    return sub


def parity_block(device, wl, oracle_kind: str, n: int, slot: int = 0):
    """Run the first n particles of wl.pool through the CUDA E-step (slot `slot`, accumulators cleared before and after) and
    through the CPU oracle of `oracle_kind`; compare what the north star names.  Returns a JSON-able dict."""
    import os
    from oracle.bindings import Oracle, Projector, Backprojector
    orc = Oracle(oracle_kind)
    K = wl.model.nr_classes
    sub = pool_slice(wl.pool, n, wl.model.current_size)
    for k in range(K):
        device.bp_clear(k)
    device.pool_upload(slot, sub)
    res = device.estep_slot(slot)
    device.sync_all_backprojects()
    sig = classify_significance(device, slot, res.particles, orc, wl.model.adaptive_fraction, wl.model.maximum_significants)
    acc = [device.bp_get(k) for k in range(K)]
    for k in range(K):
        device.bp_clear(k)
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
    st, ores, _ = orc.estep_pool(wl.model, wl.sampling, refs, bps, sub, num_threads=os.cpu_count() or 1, exact_threshold=False)
    assert st == 0, st
    g, o = res.particles, ores.particles
    pose = float(np.mean(g["best_ihidden_over"] == o["best_ihidden_over"]))
    nsig_equal = float(np.mean(g["nr_significant_coarse"] == o["nr_significant_coarse"]))
    ll = float(np.max(np.abs(g["dLL_nolog"] - o["dLL_nolog"]) / np.maximum(np.abs(o["dLL_nolog"]), 1e-30)))
    bp_rel = 0.0
    for k in range(K):
        for a, b in zip(acc[k], (bps[k].real, bps[k].imag, bps[k].weight)):
            m = float(np.abs(b).max())
            if m > 0:
                bp_rel = max(bp_rel, float(np.abs(a - b).max()) / m)
    return {"oracle": oracle_kind, "particles": int(n), "pose_agree": round(pose, 5), "nsig_equal_frac": round(nsig_equal, 5),
            "nsig_on_gpu_weights": {k: sig[k] for k in ("equal", "rounding_edge", "mismatch")},
            "ll_rel_max": float("%.3g" % ll), "bp_rel_max": float("%.3g" % bp_rel)}
