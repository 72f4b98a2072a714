"""Two ranks, one GPU each: the per-iteration reduction through the library's own NCCL communicator (rb_comm_*,
rb_bp_allreduce, rb_wsum_allreduce — relion_b200/csrc/comm.cu; replaces MlOptimiserMpi::combineAllWeightedSums,
/root/reference/src/ml_optimiser_mpi.cpp:2028-2185) must give what a single GPU gives on all particles.
Needs a box with >= 2 GPUs (gpurun --gpus 2); skipped otherwise.  The gloo twin of the host logic runs on CPU in
tests/test_parallel_cpu.py; bench.py --gpus N repeats this check inside every multi-GPU run (`allreduce_parity`)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

KW = dict(ori_size=32, healpix_order=2, n_particles=24, nr_classes=2, seed=77, snr=0.3, local_search=True, nr_groups=3)


def _setup(dev, wl):
    dev.set_model(wl.model)
    dev.set_sampling(wl.sampling)
    for k, v in enumerate(wl.refs):
        dev.set_reference(k, v, wl.r_max, wl.padding_factor)
        dev.bp_init(k, wl.bp_shape, wl.r_max, wl.padding_factor)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from relion_b200 import parallel
    from relion_b200.estep import MlDeviceBundle
    from relion_b200.workload import make_workload
    from oracle.parity import pool_range
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)        # plumbing only: the data path is the library's NCCL
    wl = make_workload(**KW)
    dev = MlDeviceBundle(rank)
    _setup(dev, wl)
    comm = parallel.DeviceComm(dev)
    a, b = parallel.shard_range(wl.pool.n_particles, rank, world)
    sub = pool_range(wl.pool, a, b, wl.model.current_size)
    res = dev.expectation_some_particles(sub)
    nshell = wl.model.ori_size // 2 + 1
    sums = parallel.fold_pool_result({}, res, sub.group_id, sub.optics_group, len(wl.model.scale_correction), 1,
                                     np.asarray(wl.model.scale_correction, np.float64), np.zeros(1))
    sums = comm.all_reduce_wsums(sums)
    comm.all_reduce_backprojectors()
    vols = [np.stack(dev.bp_get(k)) for k in range(wl.model.nr_classes)]
    if rank == 0:
        q.put(({k: v for k, v in sums.items()}, vols, res.particles["best_ihidden_over"].copy()))
    dist.barrier()
    comm.close()
    dev.close()
    dist.destroy_process_group()
    assert sums["sigma2_noise"].shape == (1, nshell)


def test_two_gpu_reduction_equals_one_gpu(device):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from relion_b200 import parallel
    from relion_b200.workload import make_workload
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums2, vols2, best0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    wl = make_workload(**KW)
    _setup(device, wl)
    res = device.expectation_some_particles(wl.pool)
    sums1 = parallel.fold_pool_result({}, res, wl.pool.group_id, wl.pool.optics_group, len(wl.model.scale_correction), 1,
                                      np.asarray(wl.model.scale_correction, np.float64), np.zeros(1))
    assert np.array_equal(best0, res.particles["best_ihidden_over"][:len(best0)])
    for k in sums1:
        if not k.startswith("_"):
            np.testing.assert_allclose(sums2[k], sums1[k], rtol=1e-5, atol=1e-9 * max(np.abs(sums1[k]).max(), 1e-30), err_msg=k)
    for k in range(wl.model.nr_classes):
        v1 = np.stack(device.bp_get(k))
        assert np.abs(vols2[k] - v1).max() <= 1e-5 * np.abs(v1).max()
