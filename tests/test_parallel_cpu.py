"""World-size-2 gloo test of the multi-rank host logic (particle sharding + weighted-sum reduction).

The device accumulators are reduced with the same torch.distributed call on the GPU (NCCL); here the CPU oracle
stands in for the device so that the test runs without a GPU: each rank runs the E-step on its shard, the sums are
all-reduced, and the result must equal a single-rank run over all particles.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from relion_b200 import parallel  # noqa: E402


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000, 12345):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    sums = {"LL": np.array(-3.5), "sigma2_noise": np.arange(12.0).reshape(2, 6), "pdf_class": np.array([0.25, 0.75])}
    out = parallel.unpack_wsums(parallel.pack_wsums(sums))
    assert set(out) == set(sums)
    for k in sums:
        np.testing.assert_array_equal(out[k], sums[k])


def _sub_pool(pool, a, b):
    from relion_b200.estep import ParticlePool
    return ParticlePool(Fimg=pool.Fimg[a:b], Fimg_nomask=pool.Fimg_nomask[a:b], Fctf=pool.Fctf[a:b], group_id=pool.group_id[a:b],
                        optics_group=pool.optics_group[a:b], highres_Xi2=pool.highres_Xi2[a:b], old_offset=pool.old_offset[a:b],
                        prior_offset=pool.prior_offset[a:b])


def _run_shard(wl, a, b):
    from oracle.bindings import Oracle, Projector, Backprojector
    o = Oracle("port")
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
    sub = _sub_pool(wl.pool, a, b)
    st, res, _ = o.estep_pool(wl.model, wl.sampling, refs, bps, sub, num_threads=1)
    assert st == 0
    nshell = wl.model.ori_size // 2 + 1
    logsigma2 = np.zeros(1)
    sums = parallel.fold_pool_result({}, res, sub.group_id, sub.optics_group, len(wl.model.scale_correction), 1,
                                     np.asarray(wl.model.scale_correction, np.float64), logsigma2)
    assert sums["sigma2_noise"].shape == (1, nshell)
    return sums, bps


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from relion_b200.workload import make_workload
    wl = make_workload(ori_size=24, healpix_order=1, n_particles=6, nr_classes=1, seed=31, snr=0.3)
    a, b = parallel.shard_range(wl.pool.n_particles, rank, world)
    sums, bps = _run_shard(wl, a, b)
    sums = parallel.all_reduce_wsums(sums)
    vol = torch.from_numpy(np.stack([bps[0].real, bps[0].imag, bps[0].weight]))
    dist.all_reduce(vol, op=dist.ReduceOp.SUM)     # what all_reduce_backprojectors does on the device tensors
    if rank == 0:
        q.put(({k: v for k, v in sums.items()}, vol.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums2, vol2 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from relion_b200.workload import make_workload
    wl = make_workload(ori_size=24, healpix_order=1, n_particles=6, nr_classes=1, seed=31, snr=0.3)
    sums1, bps1 = _run_shard(wl, 0, wl.pool.n_particles)
    for k in sums1:
        np.testing.assert_allclose(sums2[k], sums1[k], rtol=1e-9, atol=1e-12, err_msg=k)
    vol1 = np.stack([bps1[0].real, bps1[0].imag, bps1[0].weight])
    assert np.abs(vol2 - vol1).max() <= 1e-5 * np.abs(vol1).max()


def _half_worker(rank, world, port, q):
    import os
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from relion_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    groups = parallel.make_half_set_groups(world)
    half = parallel.half_set_of_rank(rank)
    # every rank contributes (rank + 1): the sum must stay inside the half-set
    t = parallel.all_reduce_tensor(torch.full((3,), float(rank + 1), dtype=torch.float64), groups[half])
    sums = parallel.all_reduce_wsums_group({"LL": np.array(float(rank + 1)), "pdf_class": np.full(2, 10.0 * (rank + 1))}, groups[half])
    # dynamic pool hand-out: 11 pools per half, every pool taken exactly once
    store = dist.TCPStore("127.0.0.1", port + 1, world, is_master=(rank == 0), timeout=__import__("datetime").timedelta(seconds=60))
    queue = parallel.PoolQueue(store, 11, half=half)
    mine = []
    while True:
        i = queue.next()
        if i is None:
            break
        mine.append(i)
    q.put((rank, half, t.tolist(), float(sums["LL"]), sums["pdf_class"].tolist(), mine))
    dist.barrier()
    dist.destroy_process_group()


def test_half_set_groups_and_pool_queue():
    """Four gloo ranks, two half-sets: reductions stay inside a half (splitC analogue), the shared counter hands every pool of a
    half to exactly one of its ranks."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_half_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(4))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, half, t, ll, pc, mine in out:
        want = sum(r + 1 for r in range(4) if r % 2 == half)
        assert t == [float(want)] * 3 and ll == float(want) and pc == [10.0 * want] * 2
    for half in (0, 1):
        taken = sorted(i for rank, h, _, _, _, mine in out if h == half for i in mine)
        assert taken == list(range(11))


def test_fold_pool_result_cc_has_no_logsigma2_term():
    """With the cross-correlation criterion dLL = -min_diff2 / nr_images (acc_ml_optimiser_impl.h:3571-3572): no logsigma2."""
    from oracle.bindings import Oracle, Projector, Backprojector
    from relion_b200.workload import make_workload
    wl = make_workload(ori_size=24, healpix_order=1, n_particles=4, nr_classes=1, seed=9, snr=0.3, do_cc=True)
    o = Oracle("port")
    refs = [Projector(v, wl.r_max, wl.padding_factor) for v in wl.refs]
    bps = [Backprojector(wl.bp_shape, wl.r_max, wl.padding_factor) for _ in wl.refs]
    st, res, _ = o.estep_pool(wl.model, wl.sampling, refs, bps, wl.pool, num_threads=1)
    assert st == 0
    logsigma2 = np.array([123.0])
    args = (res, wl.pool.group_id, wl.pool.optics_group, len(wl.model.scale_correction), 1, np.asarray(wl.model.scale_correction, np.float64), logsigma2)
    cc = parallel.fold_pool_result({}, *args, do_cc=True)
    gauss = parallel.fold_pool_result({}, *args)
    np.testing.assert_allclose(cc["LL"], -res.particles["min_diff2"].astype(np.float64).sum(), rtol=1e-6)
    np.testing.assert_allclose(cc["LL"] - gauss["LL"], 123.0 * wl.pool.n_particles, rtol=1e-12)
    assert cc["ave_Pmax"] == wl.pool.n_particles          # weight one for the best pose
