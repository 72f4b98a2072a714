"""Particle sharding and the per-iteration reduction across the GPUs of one box.

Replaces, for the E-step path only, what RELION does over MPI:
  * followers pull pools of particles from the leader (/root/reference/src/ml_optimiser_mpi.cpp:1400-1690)
        -> contiguous shards of the (already randomised) particle order, one per rank, no data-path collective;
  * MlOptimiserMpi::combineAllWeightedSums (:2028-2185): MlWsumModel::pack -> sum over workers -> unpack
        -> all_reduce(SUM) of (a) each class' back-projection accumulator, in place on the device (NCCL over
           NVLink), and (b) one fp64 vector with every other weighted sum (src/ml_model.cpp:1892-1957).
One process per GPU; torch.distributed is only the plumbing (backend "nccl" on GPUs, "gloo" in the CPU tests).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [first, last) of n_items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


@dataclasses.dataclass
class WsumPack:
    """The non-volume part of MlWsumModel::pack as one flat fp64 vector plus the layout to undo it."""
    vector: np.ndarray
    layout: List[Tuple[str, Tuple[int, ...]]]


def pack_wsums(sums: Dict[str, np.ndarray]) -> WsumPack:
    """sums: name -> array (LL, ave_Pmax, sigma2_offset, sigma2_noise[g,shell], sumw_group[g], wsum_signal_product[grp],
    wsum_reference_power[grp], pdf_direction[k,dir], pdf_class[k], ...).  Order is sorted by name so every rank agrees."""
    layout, parts = [], []
    for name in sorted(k for k in sums if not k.startswith("_")):
        a = np.asarray(sums[name], dtype=np.float64)
        layout.append((name, a.shape))
        parts.append(a.reshape(-1))
    return WsumPack(np.concatenate(parts) if parts else np.zeros(0), layout)


def unpack_wsums(pack: WsumPack) -> Dict[str, np.ndarray]:
    out, pos = {}, 0
    for name, shape in pack.layout:
        n = int(np.prod(shape)) if len(shape) else 1
        out[name] = pack.vector[pos:pos + n].reshape(shape).copy()
        pos += n
    return out


def fold_pool_result(sums: Dict[str, np.ndarray], result, group_id: np.ndarray, optics_group: np.ndarray,
                     nr_groups: int, nr_optics_groups: int, scale_correction: np.ndarray, logsigma2: np.ndarray,
                     do_cc: bool = False, current_size: int = 0, power_img: np.ndarray = None,
                     norm_correction: np.ndarray = None, avg_norm_correction: float = 1.0):
    """Host bookkeeping of storeWeightedSums after the kernels (acc_ml_optimiser_impl.h:3466-3657), in fp64: folds one pool's
    per-particle outputs into the running weighted sums of this rank.  The C++ adapter (include/relion_b200_adapter.hpp,
    MlOptimiserCuda::doThreadExpectationSomeParticles) does the same against MlWsumModel; this is its Python mirror.
      do_cc            first-iteration cross-correlation criterion: dLL = -min_diff2 without the logsigma2 term (:3571-3572)
      current_size     image_current_size; when it is smaller than the box, sigma2_noise (and the norm correction) beyond
                       current_size / 2 come from the particles' own power spectra (:3505-3515): pass power_img
                       ([P, ori_size/2+1], what rb_pool_prepare returns).  Without it the call fails instead of leaving zeros.
      norm_correction  rlnNormCorrection of the particles: accumulates avg_norm_correction (:3519-3538, :3652) and returns the
                       particles' new values in sums["_norm_correction_new"] (per pool, not reduced)
    Not covered here (the C++ adapter has them): sumw_ctf2 of CTF-premultiplied data, per-optics-group resampling (i_resam)."""
    p = result.particles
    nshell = result.wsum_sigma2_noise.shape[1]
    z = lambda *s: np.zeros(s, np.float64)
    for name, shape in (("LL", ()), ("ave_Pmax", ()), ("sigma2_offset", ()), ("avg_norm_correction", ()), ("sigma2_noise", (nr_optics_groups, nshell)),
                        ("sumw_group", (nr_optics_groups,)), ("wsum_signal_product", (nr_groups,)),
                        ("wsum_reference_power", (nr_groups,)), ("pdf_direction", result.wsum_pdf_direction.shape),
                        ("pdf_class", result.wsum_pdf_class.shape)):
        sums.setdefault(name, z(*shape))
    shells = result.wsum_sigma2_noise.astype(np.float64)
    wsum_norm = p["wsum_norm_correction"].astype(np.float64)
    first_hi = current_size // 2 + 1 if current_size else nshell
    if first_hi < nshell:
        if power_img is None:
            raise ValueError("fold_pool_result: current_size < ori_size needs power_img (sigma2_noise beyond the current size, "
                             "acc_ml_optimiser_impl.h:3505-3515)")
        hi = np.asarray(power_img, np.float64)[:, first_hi:]
        shells = shells.copy()
        shells[:, first_hi:] += hi
        wsum_norm = wsum_norm + hi.sum(axis=1)
    dll = p["dLL_nolog"] if do_cc else p["dLL_nolog"] - logsigma2[optics_group]
    sums["LL"] = sums["LL"] + dll.sum()
    sums["ave_Pmax"] = sums["ave_Pmax"] + p["pmax"].astype(np.float64).sum()
    sums["sigma2_offset"] = sums["sigma2_offset"] + p["wsum_sigma2_offset"].sum()
    np.add.at(sums["sigma2_noise"], optics_group, shells)
    np.add.at(sums["sumw_group"], optics_group, p["sumw"])
    sc = scale_correction[group_id]
    np.add.at(sums["wsum_signal_product"], group_id, p["wsum_XA"] / sc)            # :3550-3554
    np.add.at(sums["wsum_reference_power"], group_id, p["wsum_AA"] / (sc * sc))
    sums["pdf_direction"] = sums["pdf_direction"] + result.wsum_pdf_direction
    sums["pdf_class"] = sums["pdf_class"] + result.wsum_pdf_class
    if getattr(result, "wsum_prior_offset_class", None) is not None:                # 2D references (:3639-3642)
        sums["prior_offset_class"] = sums.get("prior_offset_class", 0.0) + result.wsum_prior_offset_class
    if norm_correction is not None:
        new = (np.asarray(norm_correction, np.float64) / avg_norm_correction) * np.sqrt(2.0 * wsum_norm)      # :3525-3531
        sums["avg_norm_correction"] = sums["avg_norm_correction"] + new.sum()
        sums["_norm_correction_new"] = new
    return sums


def all_reduce_wsums(sums: Dict[str, np.ndarray], device=None) -> Dict[str, np.ndarray]:
    """Sum the small weighted sums over all ranks (one fp64 all_reduce)."""
    import torch
    import torch.distributed as dist
    pack = pack_wsums(sums)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return unpack_wsums(pack)
    t = torch.from_numpy(pack.vector.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    pack.vector = t.cpu().numpy()
    return unpack_wsums(pack)


class DeviceComm:
    """NCCL communicator behind the C-ABI (rb_comm_*, relion_b200/csrc/comm.cu): the reduction itself is C / NCCL on the
    library's own stream; torch.distributed (any backend) only carries the 128-byte unique id.  `ranks`: the members (default:
    every rank).  EVERY rank of the job constructs the object with the same `ranks` (the id broadcast is job-wide); ranks
    that are not members get an inert object.  Half-set communicators: make_half_set_comms()."""

    def __init__(self, bundle, ranks=None):
        import ctypes as C
        import torch.distributed as dist
        from . import capi
        self.bundle, self.lib = bundle, bundle.lib
        world = dist.get_world_size()
        self.ranks = list(range(world)) if ranks is None else list(ranks)
        me = dist.get_rank()
        self.handle = None
        uid = (C.c_ubyte * 128)()
        if me == self.ranks[0]:
            capi.check(self.lib, self.lib.rb_comm_unique_id(uid))
        box = [bytes(uid)]
        # every rank of the job takes part in the broadcast (simple and collective-safe); only members create the communicator
        dist.broadcast_object_list(box, src=self.ranks[0])
        if me not in self.ranks:
            return
        uid = (C.c_ubyte * 128).from_buffer_copy(box[0])
        h = C.c_void_p()
        capi.check(self.lib, self.lib.rb_comm_create(bundle.ctx, len(self.ranks), self.ranks.index(me), uid, C.byref(h)))
        self.handle = h

    def all_reduce_backprojectors(self):
        """Every class' accumulator summed over the ranks, in place, ordered on the library's stream and complete on return."""
        from . import capi
        capi.check(self.lib, self.lib.rb_bp_allreduce(self.bundle.ctx, self.handle))

    def all_reduce_wsums(self, sums: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        import ctypes as C
        from . import capi
        pack = pack_wsums(sums)
        v = np.ascontiguousarray(pack.vector, np.float64)
        capi.check(self.lib, self.lib.rb_wsum_allreduce(self.bundle.ctx, self.handle, v.ctypes.data_as(C.POINTER(C.c_double)), v.size))
        pack.vector = v
        return unpack_wsums(pack)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rb_comm_destroy(self.handle)
            self.handle = None


def make_half_set_comms(bundle, world_size: int):
    """One DeviceComm per half-set (the analogue of splitC, /root/reference/src/mpi.cpp:79); returns the calling rank's own."""
    import torch.distributed as dist
    mine = None
    for h in (0, 1):
        c = DeviceComm(bundle, [r for r in range(world_size) if half_set_of_rank(r) == h])
        if half_set_of_rank(dist.get_rank()) == h:
            mine = c
    return mine


def all_reduce_backprojectors(bundle, nr_classes: int, comm: "DeviceComm" = None):
    """In-place sum of every class' device accumulator over all ranks.  With a DeviceComm: NCCL from C on the library's
    stream (rb_bp_allreduce).  Without: torch.distributed on torch's stream, followed by a device synchronisation - the
    library's streams are non-blocking and do not order against torch's, so a later reconstruct / bp_get / symmetrise could
    otherwise read the accumulator before the reduction has finished."""
    if comm is not None:
        comm.all_reduce_backprojectors()
        return
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    bundle.sync_all_backprojects()
    for k in range(nr_classes):
        t = bundle.bp_device_tensor(k)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if torch.cuda.is_available():
        torch.cuda.synchronize(bundle.device_id)


# --------------------------------------------------------------------------------------------------
# gold-standard half-sets and dynamic pool hand-out
# --------------------------------------------------------------------------------------------------
def half_set_of_rank(rank: int) -> int:
    """Which random half-set a rank works on: alternating, as MlOptimiserMpi assigns followers
    (/root/reference/src/ml_optimiser_mpi.cpp:125-160: odd followers half 1, even followers half 2); 0-based here."""
    return rank % 2


def make_half_set_groups(world_size: int):
    """One process group per half-set (the analogue of splitC, src/mpi.cpp:79): the back-projection accumulators and the
    weighted sums of a half are summed within its group only.  Every rank must call this (new_group is collective).
    Returns [group_half0, group_half1]; with world_size == 1 both halves live on the one rank and no group is needed."""
    import torch.distributed as dist
    if world_size < 2:
        return [None, None]
    return [dist.new_group([r for r in range(world_size) if half_set_of_rank(r) == h]) for h in (0, 1)]


def all_reduce_tensor(t, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_backprojectors_group(bundle, nr_classes: int, group=None):
    """all_reduce_backprojectors restricted to a half-set group."""
    import torch
    bundle.sync_all_backprojects()
    for k in range(nr_classes):
        all_reduce_tensor(bundle.bp_device_tensor(k), group)
    if torch.cuda.is_available():
        torch.cuda.synchronize(bundle.device_id)     # the library's streams do not order against torch's


def all_reduce_wsums_group(sums: Dict[str, np.ndarray], group=None, device=None) -> Dict[str, np.ndarray]:
    import torch
    pack = pack_wsums(sums)
    t = torch.from_numpy(pack.vector.copy())
    if device is not None:
        t = t.to(device)
    all_reduce_tensor(t, group)
    pack.vector = t.cpu().numpy()
    return unpack_wsums(pack)


class PoolQueue:
    """Dynamic hand-out of pools to ranks: particles of local searches differ ~10x in cost, so ranks take the next pool from
    a shared counter instead of owning a fixed range (what RELION's leader does over MPI, src/ml_optimiser_mpi.cpp:1400-1690;
    here an atomic add on a torch.distributed TCPStore, no data-path collective).  One counter per half-set."""

    def __init__(self, store, n_pools: int, half: int = 0, iteration: int = 0):
        self.store, self.n_pools = store, int(n_pools)
        self.key = f"rb_pool_queue_{iteration}_{half}"

    def next(self):
        """Index of the next pool for the calling rank, or None when the half-set is exhausted."""
        i = int(self.store.add(self.key, 1)) - 1
        return i if i < self.n_pools else None
