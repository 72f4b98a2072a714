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
    for name in sorted(sums):
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
                     do_cc: bool = False):
    """Host bookkeeping of storeWeightedSums after the kernels (acc_ml_optimiser_impl.h:3546-3657), in fp64:
    folds one pool's per-particle outputs into the running weighted sums of this rank.
    do_cc (first-iteration cross-correlation criterion): dLL = -min_diff2 without the logsigma2 term (:3571-3572)."""
    p = result.particles
    nshell = result.wsum_sigma2_noise.shape[1]
    z = lambda *s: np.zeros(s, np.float64)
    for name, shape in (("LL", ()), ("ave_Pmax", ()), ("sigma2_offset", ()), ("sigma2_noise", (nr_optics_groups, nshell)),
                        ("sumw_group", (nr_optics_groups,)), ("wsum_signal_product", (nr_groups,)),
                        ("wsum_reference_power", (nr_groups,)), ("pdf_direction", result.wsum_pdf_direction.shape),
                        ("pdf_class", result.wsum_pdf_class.shape)):
        sums.setdefault(name, z(*shape))
    dll = p["dLL_nolog"] if do_cc else p["dLL_nolog"] - logsigma2[optics_group]
    sums["LL"] = sums["LL"] + dll.sum()
    sums["ave_Pmax"] = sums["ave_Pmax"] + p["pmax"].astype(np.float64).sum()
    sums["sigma2_offset"] = sums["sigma2_offset"] + p["wsum_sigma2_offset"].sum()
    np.add.at(sums["sigma2_noise"], optics_group, result.wsum_sigma2_noise.astype(np.float64))
    np.add.at(sums["sumw_group"], optics_group, p["sumw"])
    sc = scale_correction[group_id]
    np.add.at(sums["wsum_signal_product"], group_id, p["wsum_XA"] / sc)            # :3550-3554
    np.add.at(sums["wsum_reference_power"], group_id, p["wsum_AA"] / (sc * sc))
    sums["pdf_direction"] = sums["pdf_direction"] + result.wsum_pdf_direction
    sums["pdf_class"] = sums["pdf_class"] + result.wsum_pdf_class
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


def all_reduce_backprojectors(bundle, nr_classes: int):
    """In-place sum of every class' device accumulator over all ranks (NCCL)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    bundle.sync_all_backprojects()
    for k in range(nr_classes):
        t = bundle.bp_device_tensor(k)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


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
    bundle.sync_all_backprojects()
    for k in range(nr_classes):
        all_reduce_tensor(bundle.bp_device_tensor(k), group)


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
