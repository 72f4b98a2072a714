#!/usr/bin/env python
"""bench.py — particles/s of the E-step hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--pool P]

A "step" is one pass of the whole hot path (coarse diff2 -> weights/significance -> fine diff2 -> weights ->
weighted sums + back-projection) over one pool of P synthetic particles.  Headline workload:
`refine3d_256_local` = late 3D auto-refine iteration at 256 px (C1, K=1, HEALPix order 4 coarse -> order 5
oversampled, local angular searches with sigma = 2 x oversampled step, offset range 3 / step 1 px, full-size
256-px window, 515^3 padded reference and accumulator), the regime BASELINE.json's metric is quoted on.

  value  device-resident throughput: the pool is already in HBM; K steps timed with CUDA events on the stream
         the kernels run on; max over ranks.  One NCCL all-reduce of the back-projection accumulator closes
         the timed region when N > 1 (the per-iteration reduction that replaces MPI).
  e2e    same metric through the public C-ABI call (rb_estep_pool) with pinned HOST buffers: H2D of the
         pool and D2H of the results are inside the timed region, every step.
  roofline      dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle on a bounded sample of the same pool, all host cores (rank 0, N == 1).

`--impl reference` times the reference's own CPU implementation of the path (oracle/_ref: RELION's ALTCPU
kernels compiled from /root/reference; else the restated port) on all host threads, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particles/sec per E-step (3D refine, 256 px)"
UNIT = "particles/s"

WORKLOADS = {
    # name: (make_workload kwargs, default pool size)
    "refine3d_256_local": (dict(ori_size=256, healpix_order=4, offset_range=3.0, offset_step=1.0, nr_classes=1, snr=0.05,
                                local_search=True, pixel_size=1.0, n_blobs=200, nr_groups=8), 512),
    "refine3d_128_local": (dict(ori_size=128, healpix_order=3, offset_range=3.0, offset_step=1.0, nr_classes=1, snr=0.05,
                                local_search=True, pixel_size=2.0, n_blobs=100, nr_groups=8), 512),
    "refine3d_128_global": (dict(ori_size=128, current_size=64, healpix_order=2, offset_range=5.0, offset_step=2.0, nr_classes=1,
                                 snr=0.05, pixel_size=2.0, n_blobs=100, nr_groups=8), 256),
    # BASELINE config #4 regime: 3D classification, global search at HEALPix order 3 (36 864 orientations), 21 translations,
    # coarse window 54 px: the coarse cross term is the dense contraction that runs on the tensor cores
    "class3d_256_global": (dict(ori_size=256, current_size=128, healpix_order=3, offset_range=5.0, offset_step=2.0, nr_classes=4,
                                snr=0.05, pixel_size=1.0, n_blobs=200, nr_groups=8), 256),
    # BASELINE config #4 itself: K = 8 (the per-iteration reduction then carries eight accumulators)
    "class3d_256_global_k8": (dict(ori_size=256, current_size=128, healpix_order=3, offset_range=5.0, offset_step=2.0, nr_classes=8,
                                   snr=0.05, pixel_size=1.0, n_blobs=200, nr_groups=8), 256),
    # BASELINE config #5 sizing: 400-px box, 803^3 padded reference / accumulator (16.6 GB expanded reference per class)
    "refine3d_400_local": (dict(ori_size=400, healpix_order=4, offset_range=3.0, offset_step=1.0, nr_classes=1, snr=0.05,
                                local_search=True, pixel_size=1.0, n_blobs=200, nr_groups=8), 256),
    # BASELINE config #1: 2D classification, K = 10, 64-px particles, psi step 6 deg, offset range 5 / step 2
    "class2d_64": (dict(ori_size=64, nr_classes=10, ref_dim=2, psi_step=6.0, offset_range=5.0, offset_step=2.0, snr=0.1,
                        pixel_size=3.0, n_blobs=25, nr_groups=8), 2000),
    # BASELINE config #3, state (i): first iteration of a 3D auto-refine with --firstiter_cc (cross-correlation criterion):
    # HEALPix order 2 global search (4 608 orientations), 21 translations, 30 A initial low-pass -> 40-px current size
    "refine3d_128_firstiter_cc": (dict(ori_size=128, current_size=40, healpix_order=2, offset_range=5.0, offset_step=2.0, nr_classes=1,
                                       snr=0.05, pixel_size=2.0, n_blobs=100, nr_groups=8, do_cc=True), 256),
    "tiny": (dict(ori_size=32, healpix_order=1, nr_classes=1, snr=0.3, n_blobs=20), 16),
}


def ncu_traffic(workload, pool, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this workload (profiles/traffic_r02.json, written by
    tools/sass_hash.py record), or None: the capture only counts while the kernel's SASS is the one that was profiled."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_hash
        return sass_hash.traffic(workload, pool, kernel)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name, pool, seed, projector=None):
    from relion_b200.workload import make_workload
    kw, dflt = WORKLOADS[name]
    P = pool or dflt
    return make_workload(name, n_particles=P, seed=seed, projector=projector, **kw), P


def build_device_workload(dev, name, P, seed=1993, host_refs=True, keep_refs=False):
    """The named workload with `P` particles, references already on `dev`; noise-free slices of 3D references come from the
    device projector (rb_project), 2D references are small enough for the numpy generator.  Model / sampling / accumulators
    are left to the caller (dev.set_model, dev.set_sampling, dev.bp_init).
    host_refs=False: the references are made on the device from the phantom maps (rb_set_reference_from_map) and wl.refs only
    carries their shape: much faster for the 400-px box, but no host copy for the CPU oracle.
    keep_refs=True (with host_refs=False): the references already on the device stay as they are."""
    from relion_b200.workload import make_workload
    from relion_b200 import synth
    kw, _ = WORKLOADS[name]
    if os.environ.get("RB_BENCH_COARSE_SIZE"):          # experiments only (cache-footprint studies of the coarse pass)
        kw = dict(kw, coarse_size=int(os.environ["RB_BENCH_COARSE_SIZE"]))
    if not host_refs and kw.get("ref_dim", 3) == 3:
        cur = kw.get("current_size") or kw["ori_size"]
        r_max = min(cur // 2, kw["ori_size"] // 2)
        pad = synth.pad_size_for(r_max, 2.0)
        for k in range(kw.get("nr_classes", 1)):
            if keep_refs:
                break
            vol = synth.make_phantom(kw["ori_size"], n_blobs=kw.get("n_blobs", 40), seed=1993 + 17 * k)
            dev.set_reference_from_map(k, vol, current_size=cur)
        shape_only = [np.broadcast_to(np.zeros(1, np.complex64), (pad, pad, pad // 2 + 1))] * kw.get("nr_classes", 1)
        return make_workload(name, n_particles=P, seed=seed, projector=lambda k, eul, n: dev.project(k, n, eul),
                             refs_override=(shape_only, r_max), **kw)
    if kw.get("ref_dim", 3) == 2:
        wl = make_workload(name, n_particles=P, seed=seed, **kw)
        wl.model.bp_circle_bound = False                     # backproject2D has no circle bound (BP.h:77)
        for k, v in enumerate(wl.refs):
            dev.set_reference(k, v, wl.r_max, 2.0)
        return wl
    refs = []
    r_max = None
    for k in range(kw.get("nr_classes", 1)):
        vol = synth.make_phantom(kw["ori_size"], n_blobs=kw.get("n_blobs", 40), seed=1993 + 17 * k)
        data, r_max = synth.reference_ft(vol, current_size=kw.get("current_size") or kw["ori_size"], padding_factor=2.0)
        refs.append(data.astype(np.complex64))
        dev.set_reference(k, refs[-1], r_max, 2.0)

    def projector(k, eul, n):
        return dev.project(k, n, eul)

    return make_workload(name, n_particles=P, seed=seed, projector=projector, refs_override=(refs, r_max), **kw)


def workload_config(name, P, world):
    """The part of `config` that names the workload: identical in the `ours` and `reference` arms (the driver compares them)."""
    from relion_b200 import sampling as smp
    if name == "reconstruct_256":
        return {"workload": name, "box": 256, "padding": 2.0, "particles_per_step_per_gpu": P, "poses": "uniform SO(3)",
                "l2_policy": "accumulator (1.1 GB) and image batch both larger than L2, no flush",
                "parallelism": f"particles sharded over {world} GPU(s), one accumulator per GPU"}
    kw, _ = WORKLOADS[name]
    return {"workload": name, "box": kw["ori_size"], "current_size": kw.get("current_size") or kw["ori_size"],
            "classes": kw.get("nr_classes", 1), "pool_particles_per_gpu": P,
            "healpix_order": kw.get("healpix_order"), "psi_step": kw.get("psi_step") if kw.get("ref_dim", 3) == 2 else None,
            "offset_range": kw.get("offset_range", 3.0), "offset_step": kw.get("offset_step", 2.0),
            "reference_dim": kw.get("ref_dim", 3), "oversampling": 1, "search": "local" if kw.get("local_search") else "global",
            "criterion": "cross-correlation" if kw.get("do_cc") else "gaussian",
            "l2_policy": "inputs larger than L2 (pool images + padded reference / accumulator >> 126 MB), no flush",
            "parallelism": f"particles sharded over {world} GPU(s), references replicated"}


def valid_pixels(n):
    from relion_b200.synth import mresol
    return int((mresol(n) >= 0).sum())


def stage_bytes(wl, res):
    """Algorithmic bytes per step of the three gather/scatter kernels."""
    p = res.particles
    s = wl.sampling
    T = s.n_trans
    npf, npc = valid_pixels(wl.model.current_size), valid_pixels(wl.model.coarse_size)
    K = wl.model.nr_classes
    if wl.pool.dir_off is not None:
        n_or = (np.diff(wl.pool.dir_off) * np.diff(wl.pool.psi_off)).astype(np.float64) * K
    else:
        n_or = np.full(len(p), K * s.n_dir * s.n_psi, np.float64)
    ofs = p["n_fine_orient"].astype(np.float64)
    sf = p["n_fine_samples"].astype(np.float64)
    coarse = (64.0 * n_or * npc + 12.0 * npc + 4.0 * n_or * T).sum()
    fine = (64.0 * ofs * npf + 12.0 * npf + 4.0 * sf).sum()
    # fused wavg + back-projection, SURVEY.md §8d figures: wavg gather 64 B + back-projection 204 B
    # (8 corners x 3 arrays x 4 B, x2 read-modify-write, + 12 B inputs) per pixel of every fine orientation
    # that holds at least one significant sample (n_bp_orient: the others are skipped by the reference kernels as well)
    obp = p["n_bp_orient"].astype(np.float64)
    store = ((64.0 + 204.0) * obp * npf).sum()
    # global searches: the coarse pass is the contraction [O x 2Np] . [2Np x P T] per class (+ the norm term [O x Np] . [Np x P])
    coarse_flops = (2.0 * (2 * npc) * n_or * T + 2.0 * npc * n_or).sum() if wl.pool.dir_off is None else 0.0
    return {"coarse": coarse, "fine": fine, "store": store, "coarse_flops": coarse_flops}


def build_rooflines(workload, P, wl, stage_ms, stages, by, peak, peak_src, tensor_coarse):
    """One roofline entry per hot kernel group; the headline `roofline` is the one the metric names (fine diff2, then the
    back-projection), the dominant kernel of the step follows in `roofline_kernels` with the bound it really has."""
    total = max(stage_ms["total"], 1e-9)
    band = stage_ms.get("fine_project", -1.0) > 0
    entries = []

    def hbm(kernel, stage, ms, note, traffic_kernels):
        ach = by[stage] / (ms * 1e-3) / 1e9
        tr = [ncu_traffic(workload, P, k) for k in traffic_kernels]
        return {"kernel": kernel, "stage": stage, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "frac_of_nominal_8000": round(ach / 8000.0, 4),
                "traffic": (sum(tr) if all(t is not None for t in tr) else None), "peak_source": peak_src,
                "ms": round(ms, 4), "share_of_step": round(ms / total, 3), "algorithmic_bytes_per_launch": by[stage], "note": note}

    if stage_ms["fine"] > 0:
        if band:
            entries.append(hbm("k_project_band + k_diff2_slices_sep", "fine", stage_ms["fine"],
                               "fine diff2 = band-major projection into slices + streaming diff2; algorithmic bytes 64 O_f Np + 12 Np + 4 S_f "
                               "(SURVEY 8d) over the whole fine stage; the reference cells are L2 hits after the first touch of a shell "
                               "(traffic); the projection is bound by L1 line visits (ncu: l1tex 78 %, issue 48 %), not by DRAM bytes; the "
                               "band-ordered image copies (0.25 ms) run on a side stream under the coarse pass and are not in this figure",
                               ["k_project_band", "k_diff2_slices_sep"]))
        else:
            entries.append(hbm("k_diff2_fine_async", "fine", stage_ms["fine"], "orientation-major fine kernel (cross-correlation criterion / RB_BAND=0)", ["k_diff2_fine_async"]))
    if stage_ms["store"] > 0:
        entries.append(hbm("k_store_band" if band else "k_store", "store", stage_ms["store"],
                           "wavg + back-projection; algorithmic bytes (64 + 204) O_bp Np (SURVEY 8d: gather + read-modify-write of 8 corners x 3 arrays); "
                           "band-major order keeps the accumulator shell in L2 (padded radius-sorted blocks), the red.global.add.v4.f32 stream runs at "
                           "the L2's reduction rate: the algorithmic model charges every corner a DRAM round trip that is no longer paid, so frac can "
                           "exceed 1 (traffic = the DRAM bytes actually moved)", ["k_store_band" if band else "k_store"]))
    if stage_ms["coarse"] > 0:
        if tensor_coarse:
            entries.append({"kernel": "k_gemm_tf32x3", "stage": "coarse", "bound": "tensor", "achieved": stages["coarse"]["tensor_bf16_equivalent_TFLOPs"],
                            "peak": stages["coarse"]["bf16_peak_TFLOPs"], "unit": "TFLOP/s", "frac": stages["coarse"]["frac_of_tensor_peak"],
                            "traffic": ncu_traffic(workload, P, "k_gemm_tf32x3"), "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained",
                            "ms": round(stage_ms["coarse"], 4), "share_of_step": round(stage_ms["coarse"] / total, 3),
                            "note": "achieved = tensor-pipe work in bf16-equivalent FLOPs (" + stages["coarse"]["products"] + " products per "
                                    "fp32-equivalent product, a tf32 product counted twice) over the coarse stage time (operand builders "
                                    "included); useful fp32-equivalent rate = tensor_TFLOPs_useful"})
        else:
            fused = wl.pool.dir_off is not None and os.environ.get("RB_COARSE_FUSED", "1") != "0" and not wl.model.do_cc and wl.sampling.n_trans <= 32
            ach = by["coarse"] / (stage_ms["coarse"] * 1e-3) / 1e9
            entries.append({"kernel": "k_coarse_fused" if fused else "k_diff2_coarse", "stage": "coarse", "bound": "l1tex",
                            "achieved": round(ach, 1), "unit": "GB/s (algorithmic gather bytes; NOT an HBM figure)", "peak": None, "frac": None,
                            "traffic": ncu_traffic(workload, P, "k_coarse_fused" if fused else "k_diff2_coarse"),
                            "ms": round(stage_ms["coarse"], 4), "share_of_step": round(stage_ms["coarse"] / total, 3),
                            "algorithmic_bytes_per_launch": by["coarse"],
                            "note": "local-search coarse pass: the coarse window of the reference stays in L2 (DRAM traffic ~10x below the "
                                    "algorithmic gather bytes), the kernel is bound by the L1's line throughput for divergent 16-byte gathers "
                                    "(ncu: l1tex throughput 76 %, 25.8 sectors per request, tensor pipe 3.5 %): no HBM or tensor roofline applies"})
    order = {"fine": 0, "store": 1, "coarse": 2}
    if tensor_coarse and stage_ms["coarse"] >= max(stage_ms["fine"], stage_ms["store"]):
        order = {"coarse": 0, "fine": 1, "store": 2}     # global searches: the contraction is the step
    entries.sort(key=lambda e: order[e["stage"]])
    head = dict(entries[0]) if entries else None
    return head, entries


def quick_workload(dev, name, P, steps=5, warmup=3, parity_sample=0):
    """Compact result of another BASELINE.json configuration on the same context (single GPU): device-resident and end-to-end
    particles/s, stage times, roofline entries, optionally parity against the CPU oracle on a few particles."""
    import torch
    kw, dflt = WORKLOADS[name]
    P = P or dflt
    t0 = time.time()
    wl = build_device_workload(dev, name, P, seed=1993, host_refs=parity_sample > 0 or kw.get("ref_dim", 3) == 2)
    dev.set_model(wl.model)
    dev.set_sampling(wl.sampling)
    for k in range(wl.model.nr_classes):
        dev.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pool = wl.pool
    pool.Fimg, pool.Fimg_nomask, pool.Fctf = pin(pool.Fimg.view(np.float32)), pin(pool.Fimg_nomask.view(np.float32)), pin(pool.Fctf)
    gen_s = time.time() - t0
    dev.pool_upload(0, pool)
    for _ in range(warmup):
        dev.estep_slot_nocopy(0)
    dev.sync_all_backprojects()
    res0 = dev.estep_fetch(0)
    stage_names = ["coarse", "weights_coarse", "fine_setup", "fine", "weights_fine", "store", "total", "fine_project", "fine_diff2", "store_band"]
    dev.timer_start()
    for _ in range(steps):
        dev.estep_slot_nocopy(0)
    ms = dev.timer_stop()
    stage_ms = {s: dev.stage_ms(s) for s in stage_names}
    # end to end: upload / launch / fetch pipelined over two slots, host buffers pinned
    for wslot in range(2):
        dev.pool_upload(wslot, pool)
        dev.estep_slot(wslot)
    dev.sync_all_backprojects()
    t1 = time.perf_counter()
    dev.pool_upload(0, pool)
    for i in range(steps):
        dev.estep_slot_nocopy(i % 2)
        if i >= 1:
            dev.estep_fetch((i - 1) % 2)
        if i + 1 < steps:
            dev.pool_upload((i + 1) % 2, pool)
    dev.estep_fetch((steps - 1) % 2)
    dev.sync_all_backprojects()
    e2e_s = time.perf_counter() - t1
    peak, peak_src = peaks()
    by = stage_bytes(wl, res0)
    stages = {s: {"ms": round(stage_ms[s], 4)} for s in stage_names}
    gemm_env = os.environ.get("RB_COARSE_GEMM", "1")
    tensor_coarse = by["coarse_flops"] > 0 and gemm_env != "0" and (gemm_env == "2" or wl.sampling.n_dir * wl.sampling.n_psi >= 32)
    if tensor_coarse:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        bf16_peak = float(d.get("bf16_tflops_sustained", 1400.0))
        mixed = int(os.environ.get("RB_GEMM_MODE", "3")) & 1
        units = 4 if mixed else 6
        tfs = by["coarse_flops"] / (stage_ms["coarse"] * 1e-3) / 1e12
        stages["coarse"].update({"tensor_TFLOPs_useful": round(tfs, 1), "tensor_bf16_equivalent_TFLOPs": round(units * tfs, 1),
                                 "bf16_peak_TFLOPs": round(bf16_peak, 1), "frac_of_tensor_peak": round(units * tfs / bf16_peak, 4),
                                 "products": "1 tf32 + 2 bf16" if mixed else "3 tf32"})
    _, entries = build_rooflines(name, P, wl, stage_ms, stages, by, peak, peak_src, tensor_coarse)
    out = {"config": workload_config(name, P, 1), "value": round(P * steps / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms / steps, 4),
           "e2e": round(P * steps / e2e_s, 1), "steps": steps, "stages_ms": {k: round(v, 3) for k, v in stage_ms.items() if v >= 0},
           "roofline_kernels": [{k: e[k] for k in ("kernel", "bound", "achieved", "unit", "peak", "frac", "ms", "share_of_step")} for e in entries],
           "datagen_s": round(gen_s, 1)}
    if parity_sample > 0:
        from oracle.parity import parity_block
        from oracle.bindings import have_reference
        try:
            out["parity"] = parity_block(dev, wl, "reference" if have_reference() else "port", min(P, parity_sample))
        except Exception as e:  # noqa: BLE001
            out["parity"] = {"error": repr(e)[:200]}
    return out


def quick_reconstruct(dev, P=1024, steps=3):
    """BASELINE config #2 in compact form: posed back-projection of P images at 256 px (device-resident and through rb_backproject_posed)."""
    import torch
    from relion_b200 import synth
    n, r_max, pf = 256, 128, 2.0
    xs = n // 2 + 1
    rng = np.random.default_rng(1)
    ctfs = np.stack([synth.CTF(d, d + 300.0, 30.0).fftw_image(n, n, 1.0) for d in rng.uniform(10000, 30000, 8)]).astype(np.float32)
    ci = rng.integers(0, 8, P)
    F = torch.empty((P, n, xs, 2), dtype=torch.float32).pin_memory()
    W = torch.empty((P, n, xs), dtype=torch.float32).pin_memory()
    Fn, Wn = F.numpy(), W.numpy()
    for p0 in range(0, P, 256):
        p1 = min(P, p0 + 256)
        c = ctfs[ci[p0:p1]]
        Fn[p0:p1] = rng.standard_normal((p1 - p0, n, xs, 2), dtype=np.float32) * c[..., None]
        Wn[p0:p1] = c * c
    Fn[:, 0, 0, :] = 0.0
    R = synth.inverse_euler_f32(rng.uniform(-180, 180, P), np.degrees(np.arccos(rng.uniform(-1, 1, P))), rng.uniform(0, 360, P))
    pad = synth.pad_size_for(r_max, pf)
    dev.bp_init(0, (pad, pad, pad // 2 + 1), r_max, pf)
    dev.bp_posed_stage(n, F, W, R)
    for _ in range(2):
        dev.bp_posed_run(0)
    dev.sync_all_backprojects()
    dev.timer_start()
    for _ in range(steps):
        dev.bp_posed_run(0)
    ms = dev.timer_stop()
    dev.backproject_posed(0, n, F, W, R)
    dev.sync_all_backprojects()
    t1 = time.perf_counter()
    for _ in range(steps):
        dev.backproject_posed(0, n, F, W, R)
    dev.sync_all_backprojects()
    e2e_s = time.perf_counter() - t1
    iy = np.arange(n); yy = np.where(iy < xs, iy, iy - n)[:, None]; xx = np.arange(xs)[None, :]
    inside = (xx * xx + yy * yy <= r_max * r_max) & ~((xx == 0) & (yy < 0))
    npr = float((inside[None] & (Wn[:64] > 0)).sum(axis=(1, 2)).mean())
    peak, _ = peaks()
    ach = 204.0 * npr * P * steps / (ms * 1e-3) / 1e9
    return {"config": workload_config("reconstruct_256", P, 1), "value": round(P * steps / (ms * 1e-3), 1), "unit": UNIT,
            "ms_per_step": round(ms / steps, 4), "e2e": round(P * steps / e2e_s, 1), "steps": steps,
            "roofline_kernels": [{"kernel": "k_posed_sort + k_posed_band", "bound": "hbm", "achieved": round(ach, 1), "unit": "GB/s", "peak": peak,
                                  "frac": round(ach / peak, 4)}]}


def other_workloads_block(dev, args):
    """BASELINE.json configs #1, #2, #4 (regime) and #5 (sizing) in the same driver-run line: compact, single GPU."""
    out = {}
    plan = [("class2d_64", lambda: quick_workload(dev, "class2d_64", 2000, parity_sample=64)),
            ("class3d_256_global", lambda: quick_workload(dev, "class3d_256_global", 256)),
            ("refine3d_400_local", lambda: quick_workload(dev, "refine3d_400_local", 256)),
            ("reconstruct_256", lambda: quick_reconstruct(dev))]
    for name, fn in plan:
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)[:300]}
    return out


def allreduce_parity_check(dev, comm, workload, wl, rank, world, per_rank=16):
    """Every rank builds the SAME small pool (fixed seed), works on its shard, the accumulators and the weighted sums are
    all-reduced over NCCL (rb_bp_allreduce / rb_wsum_allreduce); rank 0 then runs the whole pool alone and compares."""
    from relion_b200 import parallel
    from oracle.parity import pool_range
    n = per_rank * world
    chk = build_device_workload(dev, workload, n, seed=4242, host_refs=False, keep_refs=True) if WORKLOADS[workload][0].get("ref_dim", 3) == 3 else wl
    K = wl.model.nr_classes
    a, b = parallel.shard_range(n, rank, world)
    # the timed workload's model (noise spectra, group scales) is seeded per rank: the check runs under the check pool's own
    # model, which is the same on every rank, as the replicated model of a real run is
    dev.set_model(chk.model)
    dev.set_sampling(chk.sampling)
    for k in range(K):
        dev.bp_clear(k)
    dev.pool_upload(0, pool_range(chk.pool, a, b, wl.model.current_size))
    res = dev.estep_slot(0)
    sums = comm.all_reduce_wsums({"LL": np.array(res.particles["dLL_nolog"].sum()), "pmax": np.array(float(res.particles["pmax"].sum()))})
    comm.all_reduce_backprojectors()
    if rank != 0:
        dev.set_model(wl.model); dev.set_sampling(wl.sampling)
        return None
    tot = [dev.bp_get(k) for k in range(K)]
    for k in range(K):
        dev.bp_clear(k)
    dev.pool_upload(0, pool_range(chk.pool, 0, n, wl.model.current_size))
    one = dev.estep_slot(0)
    single = [dev.bp_get(k) for k in range(K)]
    rel = 0.0
    for k in range(K):
        for x, y in zip(tot[k], single[k]):
            m = float(np.abs(y).max())
            if m > 0:
                rel = max(rel, float(np.abs(x - y).max()) / m)
    ll1 = float(one.particles["dLL_nolog"].sum())
    dev.set_model(wl.model); dev.set_sampling(wl.sampling)
    ll_rel = abs(float(sums["LL"]) - ll1) / max(abs(ll1), 1e-30)
    return {"particles": n, "ranks": world, "bp_rel_max": float("%.3g" % rel), "ll_rel": float("%.3g" % ll_rel),
            "ok": bool(rel <= 1e-5 and ll_rel <= 1e-9), "note": "all-reduced accumulators / LL of the sharded run vs the same particles on rank 0 alone"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from relion_b200.estep import MlDeviceBundle
    from relion_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = MlDeviceBundle(local)

    # data generation (not timed): slices of the 256-px pool come from the device projector
    P = args.pool or WORKLOADS[args.workload][1]
    t0 = time.time()
    # every rank searches its own shard of the data set: same references, different particles (weak scaling)
    wl = build_device_workload(dev, args.workload, P, seed=1993 + 1000 * rank)
    gen_s = time.time() - t0
    dev.set_model(wl.model)
    dev.set_sampling(wl.sampling)
    for k in range(wl.model.nr_classes):
        dev.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)
    # the per-iteration reduction runs through the library's own NCCL communicator (rb_comm_*, C / NCCL on the library's
    # stream); torch.distributed only carries the unique id, the barrier and the max over ranks of the timings
    comm = parallel.DeviceComm(dev) if world > 1 else None

    # pinned host copies of the pool (what the RELION adapter would stage per pool)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hF, hF0, hC = pin(wl.pool.Fimg.view(np.float32)), pin(wl.pool.Fimg_nomask.view(np.float32)), pin(wl.pool.Fctf)
    pool = wl.pool
    pool.Fimg, pool.Fimg_nomask, pool.Fctf = hF, hF0, hC
    h2d = hF.numel() * 4 + hF0.numel() * 4 + hC.numel() * 4 + P * 80
    out_probe = dev._new_out(P)
    d2h = sum(a.nbytes for a in (out_probe.result.particles, out_probe.result.wsum_sigma2_noise,
                                 out_probe.result.wsum_pdf_direction, out_probe.result.wsum_pdf_class)) + P * 200

    def barrier():
        dev.sync_all_backprojects()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing -----------------------------------------------------------------
    dev.pool_upload(0, pool)
    dev.sync_all_backprojects()
    # clocks: nvidia-smi samples every 100 ms, a timed region of a few steps is shorter than that, so the sampler runs
    # from the start of the warm-up (kept under the same load for >= 1 s) to the end of the timed region
    sampler = ClockSampler(local)
    sampler.start()
    t_w = time.perf_counter()
    n_w = 0
    while n_w < args.warmup or time.perf_counter() - t_w < 1.0:
        dev.estep_slot_nocopy(0)
        dev.sync_all_backprojects()
        n_w += 1
    res0 = dev.estep_fetch(0)
    if world > 1:
        # warm the NCCL communicator / NVLink channels on the accumulator buffers (the first collective sets them up)
        parallel.all_reduce_backprojectors(dev, wl.model.nr_classes, comm)
    for k in range(wl.model.nr_classes):
        dev.bp_clear(k)
    launches0 = dev.launch_count()
    barrier()
    stage_names = ["coarse", "weights_coarse", "fine_setup", "fine", "weights_fine", "store", "total",
                   "fine_prep", "fine_project", "fine_diff2", "store_list", "store_band"]
    stage_ms = {s: 0.0 for s in stage_names}
    dev.timer_start()
    pools_done = args.steps
    if args.scaling == "strong":
        # fixed total work: ranks draw pool indices from one shared counter (what RELION's leader hands out over MPI,
        # src/ml_optimiser_mpi.cpp:1400-1690) and wait for each pool before asking for the next (dynamic load balance)
        store = dist.distributed_c10d._get_default_store() if world > 1 else None
        queue = parallel.PoolQueue(store, args.total_pools, iteration=1) if world > 1 else None
        pools_done = 0
        while True:
            i = queue.next() if queue is not None else (pools_done if pools_done < args.total_pools else None)
            if i is None:
                break
            dev.estep_slot_nocopy(0)
            dev.sync_all_backprojects()
            pools_done += 1
    else:
        for _ in range(args.steps):
            dev.estep_slot_nocopy(0)
    allreduce_ms = None
    if world > 1:
        dev.sync_all_backprojects()
        t_ar = time.perf_counter()
        parallel.all_reduce_backprojectors(dev, wl.model.nr_classes, comm)     # rb_bp_allreduce: same stream, complete on return
        allreduce_ms = (time.perf_counter() - t_ar) * 1e3
    ms = dev.timer_stop()
    for s in stage_names:   # stage events of the last timed step (all steps do identical work)
        stage_ms[s] = dev.stage_ms(s)
    barrier()
    # ---- N > 1: the reduced result equals a single-GPU run (same particles sharded over the ranks vs all on rank 0) ----
    allreduce_parity = None
    if world > 1:
        allreduce_parity = allreduce_parity_check(dev, comm, args.workload, wl, rank, world)
        for k in range(wl.model.nr_classes):
            dev.bp_clear(k)
        barrier()
    clocks = sampler.stop()
    launches = dev.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = P * world * args.steps / (ms_max / 1e3)
    pools_per_rank = None
    if args.scaling == "strong":
        value = P * args.total_pools / (ms_max / 1e3)
        c = torch.zeros(world, dtype=torch.int64, device=f"cuda:{local}")
        c[rank] = pools_done
        if world > 1:
            dist.all_reduce(c)
        pools_per_rank = [int(v) for v in c.tolist()]
        assert sum(pools_per_rank) == args.total_pools, pools_per_rank

    if args.kernels_only:
        # profiling runs (ncu): the device-resident timed region only
        if rank == 0:
            emit({"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                  "ms_per_step": round(ms_max / args.steps, 4), "stages": {k: round(v, 4) for k, v in stage_ms.items()},
                  "config": workload_config(args.workload, P, world), "scaling": args.scaling,
                  "strong_scaling": None if args.scaling != "strong" else {"total_pools": args.total_pools, "pools_per_rank": pools_per_rank,
                                                                            "timed_region_ms": round(ms_max, 3)},
                  "allreduce_ms": None if allreduce_ms is None else round(allreduce_ms, 3), "allreduce_parity": allreduce_parity,
                  "note": "--kernels-only (profiling run)"})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end-to-end through rb_estep_pool with host buffers ---------------------------------------
    # Every step: H2D of one pool from pinned host memory (rb_pool_upload), the whole E-step, D2H of the per-particle
    # results (rb_estep_slot).  Two device slots: the upload of pool i+1 overlaps the compute of pool i, as a RELION
    # adapter would do while its threads prepare the next pool.
    for wslot in range(2):   # warm both slots (first use allocates their buffers)
        dev.pool_upload(wslot, pool)
        dev.estep_slot(wslot)
    barrier()
    t1 = time.perf_counter()
    dev.pool_upload(0, pool)
    for i in range(args.steps):
        dev.estep_slot_nocopy(i % 2)                         # enqueue pool i
        if i >= 1:
            res = dev.estep_fetch((i - 1) % 2)               # D2H of pool i-1's results while pool i computes
        if i + 1 < args.steps:
            dev.pool_upload((i + 1) % 2, pool)               # H2D of pool i+1 while pool i computes
    res = dev.estep_fetch((args.steps - 1) % 2)
    dev.sync_all_backprojects()
    e2e_s = time.perf_counter() - t1
    t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = P * world * args.steps / float(t.item())

    # ---- end to end from RAW particle images: rb_pool_prepare (translate / norm / mask / FFT / CTF on the device) + E-step ----
    e2e_raw = None
    if True:
        from relion_b200.workload import raw_pool_from
        raw = raw_pool_from(wl, seed=5 + rank)
        raw.images = torch.from_numpy(raw.images).pin_memory()
        h2d_raw = raw.images.numel() * 4 + P * 120
        for wslot in range(2):
            dev.pool_prepare(wslot, raw, want_power=False)
            dev.estep_slot(wslot)
        barrier()
        t1 = time.perf_counter()
        dev.pool_prepare(0, raw, want_power=False)
        for i in range(args.steps):
            dev.estep_slot_nocopy(i % 2)
            if i >= 1:
                dev.estep_fetch((i - 1) % 2)
            if i + 1 < args.steps:
                dev.pool_prepare((i + 1) % 2, raw, want_power=False)
        dev.estep_fetch((args.steps - 1) % 2)
        dev.sync_all_backprojects()
        t = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_raw = {"value": round(P * world * args.steps / float(t.item()), 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d_raw),
                   "d2h_bytes_per_step": int(d2h), "note": "real-space images in, image preparation (getFourierTransformsAndCtfs) on the device"}

    # ---- end to end from MRC stacks on disk: native feed (reader threads -> page-locked buffers) + preparation + E-step ----
    e2e_files = None
    try:
        e2e_files = e2e_from_files(dev, wl, raw, args.steps, barrier, world, local, d2h)
    except Exception as e:  # noqa: BLE001  (no scratch space, ...): reported, not fatal
        e2e_files = {"unavailable": repr(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peak, peak_src = peaks()
    by = stage_bytes(wl, res0)
    dom = max(("coarse", "fine", "store"), key=lambda s: stage_ms[s])
    stages = {s: {"ms": round(stage_ms[s], 4)} for s in stage_names}
    for s in ("coarse", "fine", "store"):
        if stage_ms[s] > 0:
            stages[s]["algorithmic_GBps"] = round(by[s] / (stage_ms[s] * 1e-3) / 1e9, 1)
            stages[s]["frac_of_hbm_peak"] = round(by[s] / (stage_ms[s] * 1e-3) / 1e9 / peak, 4)
    gemm_env = os.environ.get("RB_COARSE_GEMM", "1")
    tensor_coarse = by["coarse_flops"] > 0 and gemm_env != "0" and (gemm_env == "2" or wl.sampling.n_dir * wl.sampling.n_psi >= 32)
    if tensor_coarse:
        # useful FLOPs: one fp32-equivalent product per operand pair.  The kernel executes three tensor-core products per
        # useful one: TF32 x TF32 for the main term and, by default (RB_GEMM_MODE bit 0), two bf16 x bf16 correction
        # products (otherwise two more TF32 ones).  A TF32 product occupies the tensor pipe twice as long as a bf16 one, so
        # the pipe time is counted in bf16-equivalent FLOPs: 2 + 1 + 1 = 4 per useful FLOP (6 for 3xTF32) and compared with
        # the measured sustained bf16 rate.
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        bf16_peak = float(d.get("bf16_tflops_sustained", 1400.0))
        mixed = int(os.environ.get("RB_GEMM_MODE", "3")) & 1
        units = 4 if mixed else 6
        tfs = by["coarse_flops"] / (stage_ms["coarse"] * 1e-3) / 1e12
        stages["coarse"].update({"tensor_TFLOPs_useful": round(tfs, 1), "tensor_TFLOPs_executed": round(3 * tfs, 1),
                                 "tensor_bf16_equivalent_TFLOPs": round(units * tfs, 1), "bf16_peak_TFLOPs": round(bf16_peak, 1),
                                 "frac_of_tensor_peak": round(units * tfs / bf16_peak, 4),
                                 "products": "1 tf32 + 2 bf16" if mixed else "3 tf32"})
    roofline, roofline_kernels = build_rooflines(args.workload, P, wl, stage_ms, stages, by, peak, peak_src, tensor_coarse)
    # ---- CPU baseline on a bounded sample (all host cores) ------------------------------------------
    # a first sample of --cpu-sample particles sizes a second one of about 12 s of CPU work (capped by the pool)
    # (rank 0 of a single-GPU run only: with N > 1 the other ranks' host threads would share the cores)
    if world > 1 or rank != 0:
        cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
               "sample": "not measured at N > 1 (see the N = 1 line)", "seconds": 0.0}
    else:
        cores = os.cpu_count() or 1
        n1 = args.cpu_sample if args.cpu_sample > 0 else min(P, max(8 * cores, 64))
        cpu = cpu_baseline(wl, sample=n1)
        if 0 < cpu["seconds"] < 6.0 and n1 < P:
            n2 = min(P, int(n1 * 12.0 / cpu["seconds"]))
            if n2 > n1:
                cpu = cpu_baseline(wl, sample=n2)

    # ---- parity inside the run: the first particles of the SAME pool through the CPU oracle (RELION's own kernels) -----
    parity = None
    if args.parity_sample > 0:
        from oracle.parity import parity_block
        from oracle.bindings import have_reference
        try:
            parity = parity_block(dev, wl, "reference" if have_reference() else "port", min(P, args.parity_sample))
            parity["ok"] = bool(parity["pose_agree"] >= 0.995 and parity["ll_rel_max"] <= 1e-4 and parity["nsig_on_gpu_weights"]["mismatch"] == 0)
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": repr(e)[:300]}

    ref_cuda = ref_cuda_block(wl, args.ref_cuda_sample, stage_ms, P) if args.ref_cuda_sample > 0 else None
    # the other BASELINE.json configurations, compact (they replace the model / references of the context: last)
    other = other_workloads_block(dev, args) if (world == 1 and args.other_workloads and args.workload == "refine3d_256_local") else None

    pr = res0.particles
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, P, world),
        "strong_scaling": None if args.scaling != "strong" else {"total_pools": args.total_pools, "pools_per_rank": pools_per_rank,
                                                                  "timed_region_ms": round(ms_max, 3),
                                                                  "note": "fixed total work handed out by a shared queue; one all-reduce closes the region"},
        "workload_stats": {"coarse_size": wl.model.coarse_size, "coarse_translations": wl.sampling.n_trans,
                           "mean_coarse_orientations": float(np.mean(np.diff(wl.pool.dir_off) * np.diff(wl.pool.psi_off))) if wl.pool.dir_off is not None else wl.sampling.n_dir * wl.sampling.n_psi,
                           "mean_fine_orientations": float(pr["n_fine_orient"].mean()), "mean_fine_samples": float(pr["n_fine_samples"].mean()),
                           "mean_bp_orientations": float(pr["n_bp_orient"].mean())},
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "e2e_from_raw_images": e2e_raw, "e2e_from_mrc_stacks": e2e_files,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_kernels": roofline_kernels, "stages": stages, "cpu_baseline": cpu,
        "parity": parity, "ref_cuda": ref_cuda, "allreduce_ms": None if allreduce_ms is None else round(allreduce_ms, 3),
        "allreduce_parity": allreduce_parity, "other_workloads": other, "datagen_s": round(gen_s, 1),
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity.get("ok", False):
        print("bench.py: parity check against the CPU oracle FAILED: " + json.dumps(parity), file=sys.stderr)
        sys.exit(3)


def e2e_from_files(dev, wl, raw, steps, barrier, world, local, d2h):
    """Particles read from two MRC stacks (page cache) by the native feed while the GPU works on the previous pool:
    STAR file -> ParticleSet -> rb_feed_* -> rb_pool_prepare -> E-step.  One step = one pool, as in the other legs."""
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from relion_b200 import particle_io, star
    P, n, ps = raw.n_particles, wl.model.ori_size, float(wl.model.pixel_size)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 4 * P * n * n * 4 else None
    tmp = tempfile.mkdtemp(prefix="rb_bench_", dir=base)
    try:
        images = raw.images.numpy() if hasattr(raw.images, "numpy") else np.asarray(raw.images)
        half = P // 2
        particle_io.write_mrc(os.path.join(tmp, "micA.mrcs"), images[:half], ps)
        particle_io.write_mrc(os.path.join(tmp, "micB.mrcs"), images[half:], ps)
        optics = star.StarTable("optics", {"rlnOpticsGroupName": ["opticsGroup1"], "rlnOpticsGroup": [1], "rlnVoltage": [300.0],
                                           "rlnSphericalAberration": [2.7], "rlnAmplitudeContrast": [0.1], "rlnImagePixelSize": [ps],
                                           "rlnImageSize": [n], "rlnImageDimensionality": [2]})
        names = ["%06d@%s" % ((i if i < half else i - half) + 1, "micA.mrcs" if i < half else "micB.mrcs") for i in range(P)]
        parts = star.StarTable("particles", {
            "rlnImageName": names, "rlnGroupName": ["group%d" % (int(g) + 1) for g in raw.group_id], "rlnOpticsGroup": [1] * P,
            "rlnDefocusU": [float(v) for v in raw.ctf_defU], "rlnDefocusV": [float(v) for v in raw.ctf_defV],
            "rlnDefocusAngle": [float(v) for v in raw.ctf_defAngle],
            "rlnOriginXAngst": [float(v) * ps for v in raw.old_offset[:, 0]], "rlnOriginYAngst": [float(v) * ps for v in raw.old_offset[:, 1]],
            "rlnNormCorrection": [1.0 / float(v) for v in raw.norm_factor]})
        star.write_star(os.path.join(tmp, "particles.star"), [optics, parts])
        pset = particle_io.ParticleSet.read(os.path.join(tmp, "particles.star"))
        # the STAR rows are sorted per micrograph (here: unchanged order); group ids follow first appearance: map to the workload's
        gmap = {nm: int(nm[5:]) - 1 for nm in pset.group_names}
        pset.group_id = np.array([gmap[pset.group_names[g]] for g in pset.group_id], np.int32)
        local_lists = None
        if raw.dir_off is not None:
            local_lists = (raw.dir_off, raw.dir_idx, raw.dir_prior, raw.psi_off, raw.psi_idx, raw.psi_prior)
        feed = particle_io.ParticleFeed(image_size=n, max_particles=P, depth=3, n_threads=8)
        # one id list over 2 warm-up pools + `steps` timed ones: the feed reads up to 3 pools ahead of the GPU
        all_ids = np.tile(np.arange(P, dtype=np.int64), steps + 2)
        t1 = None
        done = 0
        for i, (ids, pool) in enumerate(pset.stream(feed, pool_size=P, ids=all_ids, mask_radius=raw.mask_radius,
                                                    width_mask_edge=raw.width_mask_edge, local=local_lists)):
            if i == 2:
                dev.estep_fetch(0)
                dev.estep_fetch(1)
                dev.sync_all_backprojects()
                barrier()
                t1 = time.perf_counter()
            dev.pool_prepare(i % 2, pool, want_power=False)               # H2D from the feed's page-locked buffer + preparation
            dev.estep_slot_nocopy(i % 2)
            if i >= 3:
                dev.estep_fetch((i - 1) % 2)
            done = i
        dev.estep_fetch(done % 2)
        dev.sync_all_backprojects()
        t = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        feed.close()
        return {"value": round(P * world * steps / float(t.item()), 2), "unit": UNIT, "h2d_bytes_per_step": int(P * n * n * 4 + P * 120),
                "d2h_bytes_per_step": int(d2h), "reader_threads": 8,
                "note": "MRC stacks (page cache) -> native feed -> page-locked staging -> rb_pool_prepare -> E-step; STAR metadata via ParticleSet"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ref_cuda_block(wl, n, stage_ms, P):
    """The reference's OWN CUDA kernels (oracle/_ref/librefcuda.so: diff2.cuh, wavg.cuh, BP.cuh compiled for sm_100 from
    /root/reference, launched per particle with the reference's grid shapes) on the first n particles of the same pool, on this
    GPU: summed CUDA-event time per kernel family, next to our stage times per particle.  A reported baseline."""
    from oracle.bindings import Oracle, Projector, Backprojector, have_refcuda
    from oracle.parity import pool_slice
    if not have_refcuda():
        return {"unavailable": "oracle/_ref/librefcuda.so not built (needs /root/reference in the dev container)"}
    try:
        orc = Oracle("refcuda")
        n = min(n, wl.pool.n_particles)
        sub = pool_slice(wl.pool, n, wl.model.current_size)
        refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
        bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
        # first pass: texture upload, allocations, module load; second pass timed
        orc.estep_pool(wl.model, wl.sampling, refs, bps, pool_slice(wl.pool, min(n, 4), wl.model.current_size), num_threads=1)
        orc.cuda_timers(reset=True)
        t0 = time.perf_counter()
        st, _, _ = orc.estep_pool(wl.model, wl.sampling, refs, bps, sub, num_threads=1, exact_threshold=False)
        wall = time.perf_counter() - t0
        assert st == 0, st
        ms, launches = orc.cuda_timers(reset=True)
        orc.release_cuda()
        tot = sum(ms.values())
        ours = {"coarse": stage_ms["coarse"] / P, "fine": stage_ms["fine"] / P, "store": stage_ms["store"] / P, "total": stage_ms["total"] / P}
        refp = {"coarse": ms["coarse"] / n, "fine": ms["fine"] / n, "store": (ms["wavg"] + ms["backproject"]) / n, "total": tot / n}
        return {"particles": n, "coarse_ms": round(ms["coarse"], 3), "fine_ms": round(ms["fine"], 3), "wavg_ms": round(ms["wavg"], 3),
                "bp_ms": round(ms["backproject"], 3), "kernel_launches": launches,
                "kernel_ms_per_particle": round(tot / n, 4), "particles_per_s": round(n / (tot * 1e-3), 1),
                "particles_per_s_wall_incl_host_driver": round(n / wall, 1),
                "ours_ms_per_particle": {k: round(v, 5) for k, v in ours.items()},
                "speedup_vs_ref_cuda_kernels": {k: round(refp[k] / ours[k], 2) for k in ours if ours[k] > 0},
                "note": "reference CUDA kernels (texture projector, sm_100) launched one particle at a time as the reference does; "
                        "particles_per_s = particles / summed kernel time (host orchestration and copies excluded)"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:300]}


def cpu_baseline(wl, sample, kind=None, steps=1, warmup=0):
    """The CPU oracle on the first `sample` particles of the pool, one particle per OpenMP thread."""
    from oracle.bindings import Oracle, Projector, Backprojector, have_reference
    from relion_b200.estep import ParticlePool
    kind = kind or ("reference" if have_reference() else "port")
    o = Oracle(kind)
    n = min(sample, wl.pool.n_particles)
    p = wl.pool
    as_np = lambda a, dt: np.asarray(a.numpy() if hasattr(a, "numpy") else a).view(dt) if a is not None else None
    F = as_np(p.Fimg, np.complex64) if not np.iscomplexobj(p.Fimg) else p.Fimg
    F0 = as_np(p.Fimg_nomask, np.complex64) if not np.iscomplexobj(p.Fimg_nomask) else p.Fimg_nomask
    Cc = as_np(p.Fctf, np.float32)
    xs = wl.model.current_size // 2 + 1
    F = np.asarray(F).reshape(p.group_id.shape[0], wl.model.current_size, xs)
    F0 = np.asarray(F0).reshape(F.shape)
    sub = ParticlePool(Fimg=F[:n], Fimg_nomask=F0[:n], Fctf=np.asarray(Cc).reshape(F.shape)[:n], group_id=p.group_id[:n],
                       optics_group=p.optics_group[:n], highres_Xi2=p.highres_Xi2[:n], old_offset=p.old_offset[:n],
                       prior_offset=p.prior_offset[:n])
    if p.dir_off is not None:
        sub.dir_off = p.dir_off[:n + 1]; sub.psi_off = p.psi_off[:n + 1]
        sub.dir_idx = p.dir_idx[:p.dir_off[n]]; sub.dir_prior = p.dir_prior[:p.dir_off[n]]
        sub.psi_idx = p.psi_idx[:p.psi_off[n]]; sub.psi_prior = p.psi_prior[:p.psi_off[n]]
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
    cores = os.cpu_count() or 1
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        st, _, _ = o.estep_pool(wl.model, wl.sampling, refs, bps, sub, num_threads=cores)
        dt = time.perf_counter() - t0
        assert st == 0, st
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return {"value": round(n * len(times) / tot, 3), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} particles of the same pool per step, {len(times)} step(s), one particle per OpenMP thread", "seconds": round(tot, 2)}


def run_reference(args):
    """The reference's CPU implementation of the path on all host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.workload == "reconstruct_256":
        return run_reference_reconstruct(args)
    kw, dflt = WORKLOADS[args.workload]
    from relion_b200.workload import make_workload
    P = args.pool or dflt
    cores = os.cpu_count() or 1
    # one particle per OpenMP thread and local-search particles differ ~10x in cost: a step of `cores` particles would time
    # the slowest particle.  >= 8 particles per core (or the whole pool) lets the dynamic schedule even that out.
    n = args.cpu_sample if args.cpu_sample > 0 else min(P, max(8 * cores, 64))
    wl = make_workload(args.workload, n_particles=n, seed=1993, **kw)   # numpy projector: none of our kernels on this path
    cpu = cpu_baseline(wl, sample=n, steps=args.steps, warmup=min(args.warmup, 1))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(1e3 * n / cpu["value"], 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.workload, P, max(world, args.gpus)),
           "cpu_baseline": cpu,
           "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def run_reconstruct(args):
    """BASELINE config #2: posed back-projection only (relion_reconstruct's per-particle path), 256-px particles, pad 2."""
    import torch
    import torch.distributed as dist
    from relion_b200.estep import MlDeviceBundle
    from relion_b200 import parallel, synth

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n, r_max, pf = 256, 128, 2.0
    xs = n // 2 + 1
    P = args.pool or 2048
    rng = np.random.default_rng(1 + 1000 * rank)
    # synthetic posed particles: noise spectra times a CTF (values do not change the work), weights ctf^2, uniform SO(3) poses
    ctfs = np.stack([synth.CTF(d, d + 300.0, 30.0).fftw_image(n, n, 1.0) for d in rng.uniform(10000, 30000, 16)]).astype(np.float32)
    ci = rng.integers(0, 16, P)
    F = torch.empty((P, n, xs, 2), dtype=torch.float32).pin_memory()
    W = torch.empty((P, n, xs), dtype=torch.float32).pin_memory()
    Fn, Wn = F.numpy(), W.numpy()
    for p0 in range(0, P, 256):
        p1 = min(P, p0 + 256)
        c = ctfs[ci[p0:p1]]
        Fn[p0:p1] = rng.standard_normal((p1 - p0, n, xs, 2), dtype=np.float32) * c[..., None]
        Wn[p0:p1] = c * c
    Fn[:, 0, 0, :] = 0.0                                     # DC zeroed (src/reconstructor.cpp:716)
    u = rng.standard_normal((P, 4)); u /= np.linalg.norm(u, axis=1, keepdims=True)     # uniform rotations from unit quaternions
    a, b, c_, d = u.T
    R = np.stack([a * a + b * b - c_ * c_ - d * d, 2 * (b * c_ - a * d), 2 * (b * d + a * c_),
                  2 * (b * c_ + a * d), a * a - b * b + c_ * c_ - d * d, 2 * (c_ * d - a * b),
                  2 * (b * d - a * c_), 2 * (c_ * d + a * b), a * a - b * b - c_ * c_ + d * d], axis=1).astype(np.float32)
    pad = synth.pad_size_for(r_max, pf)
    shape = (pad, pad, pad // 2 + 1)
    dev = MlDeviceBundle(local)
    dev.bp_init(0, shape, r_max, pf)

    def barrier():
        dev.sync_all_backprojects(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    dev.bp_posed_stage(n, F, W, R)
    sampler = ClockSampler(local); sampler.start()
    t_w = time.perf_counter(); n_w = 0
    while n_w < args.warmup or time.perf_counter() - t_w < 1.0:
        dev.bp_posed_run(0); dev.sync_all_backprojects(); n_w += 1
    if world > 1:
        parallel.all_reduce_backprojectors(dev, 1)           # NCCL warm-up
        torch.cuda.synchronize()
    dev.bp_clear(0)
    launches0 = dev.launch_count()
    barrier()
    dev.timer_start()
    for _ in range(args.steps):
        dev.bp_posed_run(0)
    kernel_ms = dev.timer_stop()                             # the scatter kernels alone (roofline)
    ms = kernel_ms
    if world > 1:
        dev.timer_start()
        parallel.all_reduce_backprojectors(dev, 1)
        torch.cuda.synchronize()
        ms += dev.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = dev.launch_count() - launches0
    t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, kernel_ms = float(t[0].item()), float(t[1].item())
    value = P * world * args.steps / (ms_max / 1e3)
    # end to end: host buffers through rb_backproject_posed (chunked H2D overlapped with the scatter)
    dev.backproject_posed(0, n, F, W, R)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        dev.backproject_posed(0, n, F, W, R)
    dev.sync_all_backprojects()
    e2e_s = time.perf_counter() - t1
    t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = P * world * args.steps / float(t.item())
    # end to end from RAW images: real-space pixels + CTF parameters over PCIe, transform / shift / CTF on the device
    raw = torch.empty((P, n, n), dtype=torch.float32).pin_memory()
    rn = raw.numpy()
    for p0 in range(0, P, 256):
        rn[p0:min(P, p0 + 256)] = rng.standard_normal((min(P, p0 + 256) - p0, n, n), dtype=np.float32)
    dfs = rng.uniform(10000, 30000, P)
    ctfpar = dict(defU=dfs, defV=dfs + 300.0, defAngle=np.full(P, 30.0), kV=[300.0], Cs=[2.7], Q0=[0.1])
    shifts = rng.uniform(-3, 3, (P, 2))
    dev.backproject_posed_raw(0, raw, R, shift=shifts, ctf=ctfpar, pixel_size=1.0)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        dev.backproject_posed_raw(0, raw, R, shift=shifts, ctf=ctfpar, pixel_size=1.0)
    dev.sync_all_backprojects()
    t = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_raw_value = P * world * args.steps / float(t.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # pixels scattered per image: |(x, y)| <= r_max on the half image, minus the x = 0, y < 0 column, weight > 0
    iy = np.arange(n); yy = np.where(iy < xs, iy, iy - n)[:, None]; xx = np.arange(xs)[None, :]
    inside = (xx * xx + yy * yy <= r_max * r_max) & ~((xx == 0) & (yy < 0))
    npr = float((inside[None] & (Wn[:64] > 0)).sum(axis=(1, 2)).mean())
    peak, peak_src = peaks()
    bytes_per_launch = 204.0 * npr * P                       # (2 x 96 + 12) B per scattered pixel
    ach = bytes_per_launch * args.steps / (kernel_ms * 1e-3) / 1e9
    # CPU baseline: the compiled restatement of BackProjector::backproject2Dto3D, one thread as relion_reconstruct runs it
    from oracle.bindings import backproject_posed as cpu_bp
    ns = min(P, max(8, (args.cpu_sample or 32) * 16))
    acc = tuple(np.zeros(shape, np.float64) for _ in range(3))
    for a_ in acc:
        a_.fill(0.0)                                         # map the pages before timing
    Fs = np.ascontiguousarray(Fn[:ns].view(np.complex64)[..., 0])
    t0 = time.perf_counter()
    cpu_bp(shape, Fs, Wn[:ns], R[:ns], r_max, pf, out=acc)
    cpu_s = time.perf_counter() - t0
    cpu = {"value": round(ns / cpu_s, 2), "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{ns} images, double accumulators, one thread (Reconstructor::backprojectOneParticle loop)", "seconds": round(cpu_s, 2)}
    out = {"metric": "particles/sec, posed back-projection only (relion_reconstruct path, 256 px, pad 2)", "value": round(value, 1), "unit": UNIT,
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config("reconstruct_256", P, world),
           "workload_stats": {"accumulator": list(shape), "scattered_pixels_per_image": npr},
           "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(P * n * xs * 12 + P * 36), "d2h_bytes_per_step": 0},
           "e2e_from_raw_images": {"value": round(e2e_raw_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(P * n * n * 4 + P * (36 + 88)),
                                   "d2h_bytes_per_step": 0, "note": "rb_backproject_posed_raw: real-space images + CTF parameters in; "
                                   "FourierTransform, CenterFFTbySign, origin shift, CTF, DC removal (reconstructor.cpp:428-745) on the device"},
           "gpu_launches": int(launches), "clocks": clocks,
           "roofline": {"kernel": "k_posed_sort + k_posed_band", "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(ach / peak, 4), "traffic": ncu_traffic("reconstruct_256", P, "k_backproject_posed"), "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": bytes_per_launch},
           "cpu_baseline": cpu}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_reference_reconstruct(args):
    """Reference arm of the reconstruct workload: the compiled restatement of BackProjector::backproject2Dto3D on one thread
    (relion_reconstruct back-projects one particle after the other), bounded sample per step."""
    from relion_b200 import synth
    from oracle.bindings import backproject_posed as cpu_bp
    n, r_max, pf = 256, 128, 2.0
    xs = n // 2 + 1
    ns = max(8, (args.cpu_sample or 32) * 16)
    rng = np.random.default_rng(1)
    c = synth.CTF(20000.0, 20300.0, 30.0).fftw_image(n, n, 1.0).astype(np.float32)
    F = (rng.standard_normal((ns, n, xs)) + 1j * rng.standard_normal((ns, n, xs))).astype(np.complex64) * c
    W = np.broadcast_to(c * c, (ns, n, xs)).copy()
    R = synth.inverse_euler_f32(rng.uniform(-180, 180, ns), np.degrees(np.arccos(rng.uniform(-1, 1, ns))), rng.uniform(0, 360, ns))
    pad = synth.pad_size_for(r_max, pf)
    shape = (pad, pad, pad // 2 + 1)
    times = []
    acc = tuple(np.zeros(shape, np.float64) for _ in range(3))
    for a_ in acc:
        a_.fill(0.0)                                         # map the pages before timing
    for i in range(min(args.warmup, 1) + args.steps):
        t0 = time.perf_counter()
        cpu_bp(shape, F, W, R, r_max, pf, out=acc)
        if i >= min(args.warmup, 1):
            times.append(time.perf_counter() - t0)
    v = round(ns * len(times) / sum(times), 2)
    cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{ns} images per step, {len(times)} step(s), one thread, double accumulators"}
    emit(({"impl": "reference", "metric": "particles/sec, posed back-projection only (relion_reconstruct path, 256 px, pad 2)",
                      "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": round(1e3 * ns / v, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": workload_config("reconstruct_256", args.pool or 2048, max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus)),
                      "cpu_baseline": cpu, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="refine3d_256_local", choices=sorted(WORKLOADS) + ["reconstruct_256"])
    ap.add_argument("--pool", type=int, default=0, help="particles per pool per GPU (default: workload specific)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the bounded CPU sample (0: 8 per host core, at least 64, at most the pool)")
    ap.add_argument("--kernels-only", action="store_true", help="device-resident timed region only (for ncu)")
    ap.add_argument("--other-workloads", type=int, default=1, help="also run the other BASELINE.json configurations in compact form (N = 1 only)")
    ap.add_argument("--ref-cuda-sample", type=int, default=64, help="particles run through the reference's own CUDA kernels on this GPU (0: off)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): every rank runs --steps pools; strong: a FIXED number of pools "
                         "(--total-pools) is handed out to the ranks from a shared queue (parallel.PoolQueue), fetched results included")
    ap.add_argument("--total-pools", type=int, default=64, help="--scaling strong: pools of the whole job")
    ap.add_argument("--parity-sample", type=int, default=256, help="particles of the same pool checked against the CPU oracle inside the run (0: off)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    # stdout carries the ONE JSON line and nothing else: whatever libraries print on file descriptor 1 while the run lasts
    # (NCCL's version banner, OpenMP notices) goes to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        _dispatch(args)
    finally:
        sys.stdout.flush()
        os.dup2(json_fd, 1)
        os.close(json_fd)
        if _JSON_LINES:
            sys.stdout.write(_JSON_LINES[-1] + "\n")
            sys.stdout.flush()


_JSON_LINES = []


def emit(obj):
    """The run's result line (the last one emitted is the one printed on stdout)."""
    _JSON_LINES.append(json.dumps(obj))


def _dispatch(args):
    if args.workload == "reconstruct_256" and args.impl == "ours":
        run_reconstruct(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
