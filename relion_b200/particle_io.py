"""Particle I/O and the metadata feed of the E-step driver.

Host-side mirror of
  Experiment::read                          /root/reference/src/exp_model.cpp:820-1060   particle STAR -> particles, groups, optics groups
  MlOptimiser::getMetaAndImageDataSubset    /root/reference/src/ml_optimiser.cpp:10285-10552   pool -> exp_metadata [P][25] + images
  MlOptimiser::setMetaDataSubset            /root/reference/src/ml_optimiser.cpp:10554-10640   results -> particle table
  Image<T>::readMRC / writeMRC              /root/reference/src/rwMRC.h (native: csrc/io_mrc.cpp, C-ABI rb_mrc_* / rb_feed_*)

`ParticleSet.pool(first, last)` yields the RawParticlePool rb_pool_prepare consumes; `ParticleFeed` keeps the next pools'
images streaming from their MRC stacks into page-locked buffers while the GPU works.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi, star

# exp_metadata columns (src/ml_optimiser.h:51-80)
METADATA_ROT, METADATA_TILT, METADATA_PSI, METADATA_XOFF, METADATA_YOFF, METADATA_ZOFF = 0, 1, 2, 3, 4, 5
METADATA_CLASS, METADATA_DLL, METADATA_PMAX, METADATA_NR_SIGN, METADATA_NORM = 6, 7, 8, 9, 10
METADATA_CTF_DEFOCUS_U, METADATA_CTF_DEFOCUS_V, METADATA_CTF_DEFOCUS_ANGLE = 11, 12, 13
METADATA_CTF_BFACTOR, METADATA_CTF_KFACTOR, METADATA_CTF_PHASE_SHIFT = 14, 15, 16
METADATA_ROT_PRIOR, METADATA_TILT_PRIOR, METADATA_PSI_PRIOR = 17, 18, 19
METADATA_XOFF_PRIOR, METADATA_YOFF_PRIOR, METADATA_ZOFF_PRIOR = 20, 21, 22
METADATA_PSI_PRIOR_FLIP_RATIO, METADATA_ROT_PRIOR_FLIP_RATIO = 23, 24
METADATA_LINE_LENGTH = 25
PRIOR_UNSET = 999.0


# ---------------------------------------------------------------------------------------------
# MRC files
# ---------------------------------------------------------------------------------------------
class MrcStack:
    """Read access to an .mrc / .mrcs file through the native reader (rb_mrc_*)."""

    def __init__(self, path: str):
        self.lib = capi.load_library()
        h = C.c_void_p()
        capi.check(self.lib, self.lib.rb_mrc_open(os.fsencode(path), C.byref(h)))
        self.handle = h
        self.path = path
        nx, ny, nz, mode, ps = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_float()
        capi.check(self.lib, self.lib.rb_mrc_info(h, C.byref(nx), C.byref(ny), C.byref(nz), C.byref(mode), C.byref(ps)))
        self.nx, self.ny, self.nz, self.mode, self.pixel_size = nx.value, ny.value, nz.value, mode.value, ps.value

    def read(self, indices: Optional[Sequence[int]] = None) -> np.ndarray:
        """Images `indices` (0-based; default: all) as float32 [count, ny, nx]."""
        idx = np.arange(self.nz, dtype=np.int64) if indices is None else np.ascontiguousarray(indices, np.int64)
        out = np.empty((idx.size, self.ny, self.nx), np.float32)
        capi.check(self.lib, self.lib.rb_mrc_read_images(self.handle, idx.ctypes.data_as(C.POINTER(C.c_longlong)), int(idx.size),
                                                         out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rb_mrc_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def write_mrc(path: str, data: np.ndarray, pixel_size: float = 1.0) -> None:
    """float32 stack [nz, ny, nx] (or one image [ny, nx]) with the header Image<T>::writeMRC produces."""
    lib = capi.load_library()
    a = np.ascontiguousarray(data, np.float32)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("write_mrc expects [nz, ny, nx] or [ny, nx]")
    capi.check(lib, lib.rb_mrc_write(os.fsencode(path), a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[2], a.shape[1], a.shape[0],
                                     float(pixel_size)))


def decompose_image_name(name: str) -> Tuple[int, str]:
    """"000012@Extract/job007/mic.mrcs" -> (11, path): FileName::decompose, image numbers in file names count from 1."""
    if "@" in name:
        num, path = name.split("@", 1)
        return int(num) - 1, path
    return 0, name


class ParticleFeed:
    """Prefetching reader (rb_feed_*): submit() the next pool, wait() for its images, release() the buffer when uploaded."""

    def __init__(self, image_size: int, max_particles: int, depth: int = 3, n_threads: int = 4):
        self.lib = capi.load_library()
        h = C.c_void_p()
        capi.check(self.lib, self.lib.rb_feed_create(int(image_size), int(max_particles), int(depth), int(n_threads), C.byref(h)))
        self.handle = h
        self.image_size, self.max_particles, self.depth = int(image_size), int(max_particles), int(depth)
        self._counts = {}

    def submit(self, paths: Sequence[str], indices: Sequence[int]) -> int:
        n = len(paths)
        arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        idx = np.ascontiguousarray(indices, np.int64)
        ticket = C.c_int()
        capi.check(self.lib, self.lib.rb_feed_submit(self.handle, arr, idx.ctypes.data_as(C.POINTER(C.c_longlong)), n, C.byref(ticket)))
        self._counts[ticket.value] = n
        return ticket.value

    def wait(self, ticket: int) -> np.ndarray:
        """[P, n, n] float32 view of the staging buffer (valid until release(ticket))."""
        p = C.POINTER(C.c_float)()
        capi.check(self.lib, self.lib.rb_feed_wait(self.handle, int(ticket), C.byref(p)))
        n = self._counts[ticket]
        return np.ctypeslib.as_array(p, shape=(n, self.image_size, self.image_size))

    def release(self, ticket: int) -> None:
        capi.check(self.lib, self.lib.rb_feed_release(self.handle, int(ticket)))
        self._counts.pop(ticket, None)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rb_feed_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


# ---------------------------------------------------------------------------------------------
# particle sets
# ---------------------------------------------------------------------------------------------
_PIPELINE_RE = re.compile(r"^(?:.*/)?[A-Za-z0-9_]+/job\d+/(.+)$")


def group_name_of(micrograph: str) -> str:
    """decomposePipelineFileName: the micrograph name without its "<JobType>/jobNNN/" prefix (exp_model.cpp:936-940)."""
    m = _PIPELINE_RE.match(micrograph)
    return m.group(1) if m else micrograph


@dataclass
class OpticsGroups:
    names: List[str]
    pixel_size: np.ndarray
    image_size: np.ndarray
    kV: np.ndarray
    Cs: np.ndarray
    Q0: np.ndarray


class ParticleSet:
    """A particle STAR file (RELION >= 3.1: data_optics + data_particles) laid out the way Experiment::read does it:
    particles sorted on rlnMicrographName, one noise / scale group per rlnGroupName (else per micrograph), optics groups by
    rlnOpticsGroup."""

    def __init__(self, optics: star.StarTable, particles: star.StarTable, root: str = ""):
        if particles.has("rlnMicrographName"):
            particles = particles.sorted_by("rlnMicrographName")                         # exp_model.cpp:899-901
        self.table = particles
        self.optics_table = optics
        self.root = root
        P = len(particles)
        og_ids = optics.column("rlnOpticsGroup", np.int64)
        self.optics = OpticsGroups(
            names=[str(v) for v in optics.column("rlnOpticsGroupName", default="opticsGroup")],
            pixel_size=optics.column("rlnImagePixelSize", np.float64), image_size=optics.column("rlnImageSize", np.int64),
            kV=optics.column("rlnVoltage", np.float64, default=300.0), Cs=optics.column("rlnSphericalAberration", np.float64, default=2.7),
            Q0=optics.column("rlnAmplitudeContrast", np.float64, default=0.1))
        lookup = {int(g): i for i, g in enumerate(og_ids)}
        try:
            self.optics_group = np.array([lookup[int(g)] for g in particles.column("rlnOpticsGroup", np.int64, default=int(og_ids[0]))], np.int32)
        except KeyError as e:
            raise star.StarError(f"particle refers to optics group {e.args[0]} that data_optics does not define") from None
        # groups (exp_model.cpp:927-960): by name, ids in order of first appearance
        if particles.has("rlnGroupName"):
            names = [str(v) for v in particles.column("rlnGroupName")]
        elif particles.has("rlnMicrographName"):
            names = [group_name_of(str(v)) for v in particles.column("rlnMicrographName")]
        else:
            names = ["group"] * P
        self.group_names: List[str] = []
        ids = {}
        gid = np.empty(P, np.int32)
        for i, nm in enumerate(names):
            if nm not in ids:
                ids[nm] = len(self.group_names)
                self.group_names.append(nm)
            gid[i] = ids[nm]
        self.group_id = gid
        self.random_subset = particles.column("rlnRandomSubset", np.int32, default=0)
        dec = [decompose_image_name(str(v)) for v in particles.column("rlnImageName")]
        self.image_index = np.array([d[0] for d in dec], np.int64)
        self.image_path = [os.path.join(root, d[1]) if root and not os.path.isabs(d[1]) else d[1] for d in dec]

    @classmethod
    def read(cls, path: str, root: Optional[str] = None) -> "ParticleSet":
        tables = star.read_star(path)
        if "particles" not in tables or "optics" not in tables:
            raise star.StarError(f"{path}: expected data_optics and data_particles blocks (RELION >= 3.1 particle file)")
        return cls(tables["optics"], tables["particles"], root=os.path.dirname(os.path.abspath(path)) if root is None else root)

    def __len__(self) -> int:
        return len(self.table)

    @property
    def n_groups(self) -> int:
        return len(self.group_names)

    def image_size(self) -> int:
        return int(self.optics.image_size[self.optics_group[0]])

    def half_set(self, subset: int) -> np.ndarray:
        """Indices of the particles of random half `subset` (1 or 2), the split gold-standard refinement keeps apart."""
        return np.nonzero(self.random_subset == subset)[0]

    # -- getMetaAndImageDataSubset, metadata half (ml_optimiser.cpp:10420-10520) --
    def metadata(self, ids: Sequence[int], do_ctf_correction: bool = True) -> np.ndarray:
        t = self.table
        ids = np.asarray(ids, np.int64)
        md = np.zeros((ids.size, METADATA_LINE_LENGTH), np.float64)
        ps = self.optics.pixel_size[self.optics_group[ids]]

        def col(label, default):
            return t.column(label, np.float64, default=default)[ids]

        md[:, METADATA_ROT], md[:, METADATA_TILT], md[:, METADATA_PSI] = col("rlnAngleRot", 0.0), col("rlnAngleTilt", 0.0), col("rlnAnglePsi", 0.0)
        md[:, METADATA_XOFF] = col("rlnOriginXAngst", 0.0) / ps
        md[:, METADATA_YOFF] = col("rlnOriginYAngst", 0.0) / ps
        md[:, METADATA_CLASS] = col("rlnClassNumber", 0.0)
        md[:, METADATA_DLL], md[:, METADATA_PMAX] = col("rlnLogLikeliContribution", 0.0), col("rlnMaxValueProbDistribution", 0.0)
        md[:, METADATA_NR_SIGN] = col("rlnNrOfSignificantSamples", 0.0)
        md[:, METADATA_NORM] = col("rlnNormCorrection", 1.0)
        for k, label in ((METADATA_ROT_PRIOR, "rlnAngleRotPrior"), (METADATA_TILT_PRIOR, "rlnAngleTiltPrior"), (METADATA_PSI_PRIOR, "rlnAnglePsiPrior"),
                         (METADATA_PSI_PRIOR_FLIP_RATIO, "rlnAnglePsiFlipRatio")):
            md[:, k] = col(label, PRIOR_UNSET)
        md[:, METADATA_XOFF_PRIOR] = col("rlnOriginXPriorAngst", PRIOR_UNSET * 1.0) / (ps if t.has("rlnOriginXPriorAngst") else 1.0)
        md[:, METADATA_YOFF_PRIOR] = col("rlnOriginYPriorAngst", PRIOR_UNSET * 1.0) / (ps if t.has("rlnOriginYPriorAngst") else 1.0)
        md[:, METADATA_ZOFF_PRIOR] = PRIOR_UNSET
        if do_ctf_correction:
            du = col("rlnDefocusU", 0.0)
            md[:, METADATA_CTF_DEFOCUS_U] = du
            md[:, METADATA_CTF_DEFOCUS_V] = col("rlnDefocusV", 0.0) if t.has("rlnDefocusV") else du
            md[:, METADATA_CTF_DEFOCUS_ANGLE] = col("rlnDefocusAngle", 0.0)
            md[:, METADATA_CTF_BFACTOR] = col("rlnCtfBfactor", 0.0)
            md[:, METADATA_CTF_KFACTOR] = col("rlnCtfScalefactor", 1.0)
            md[:, METADATA_CTF_PHASE_SHIFT] = col("rlnPhaseShift", 0.0)
        return md

    def pool(self, ids: Sequence[int], images: np.ndarray, avg_norm_correction: float = 1.0, do_norm_correction: bool = True,
             mask_radius: float = -1.0, width_mask_edge: float = 5.0, local=None):
        """The RawParticlePool of particles `ids` (sorted positions): what the acc driver's stage 1 starts from.
        norm_factor = avg_norm_correction / rlnNormCorrection (acc_ml_optimiser_impl.h:438-447); translation priors that
        are unset (999) become 0 (:56-58).  `local` optionally carries the per-particle orientation lists
        (dir_off, dir_idx, dir_prior, psi_off, psi_idx, psi_prior) of a local search."""
        from .estep import RawParticlePool
        ids = np.asarray(ids, np.int64)
        md = self.metadata(ids)
        prior = md[:, [METADATA_XOFF_PRIOR, METADATA_YOFF_PRIOR]].copy()
        prior[np.abs(prior - PRIOR_UNSET) < 0.01] = 0.0
        norm = md[:, METADATA_NORM]
        norm_factor = np.where(do_norm_correction & (norm > 0), avg_norm_correction / np.where(norm > 0, norm, 1.0), 1.0)
        kw = {}
        if local is not None:
            kw = dict(zip(("dir_off", "dir_idx", "dir_prior", "psi_off", "psi_idx", "psi_prior"), local))
        return RawParticlePool(images=images, old_offset=md[:, [METADATA_XOFF, METADATA_YOFF]].copy(), prior_offset=prior,
                               group_id=self.group_id[ids].copy(), optics_group=self.optics_group[ids].copy(),
                               ctf_defU=md[:, METADATA_CTF_DEFOCUS_U].copy(), ctf_defV=md[:, METADATA_CTF_DEFOCUS_V].copy(),
                               ctf_defAngle=md[:, METADATA_CTF_DEFOCUS_ANGLE].copy(), ctf_Bfac=md[:, METADATA_CTF_BFACTOR].copy(),
                               ctf_scale=md[:, METADATA_CTF_KFACTOR].copy(), ctf_phase_shift=md[:, METADATA_CTF_PHASE_SHIFT].copy(),
                               og_kV=self.optics.kV, og_Cs=self.optics.Cs, og_Q0=self.optics.Q0, norm_factor=norm_factor,
                               mask_radius=mask_radius, width_mask_edge=width_mask_edge, **kw)

    def pools(self, pool_size: int, ids: Optional[Sequence[int]] = None):
        """Consecutive id ranges of at most pool_size particles (the nr_pool loop of expectationSomeParticles)."""
        ids = np.arange(len(self), dtype=np.int64) if ids is None else np.asarray(ids, np.int64)
        for a in range(0, ids.size, pool_size):
            yield ids[a:a + pool_size]

    def stream(self, feed: ParticleFeed, pool_size: int, ids: Optional[Sequence[int]] = None, **pool_kw):
        """Generator of (ids, RawParticlePool) with the following pools' images already being read: keeps feed.depth - 1
        pools in flight, hands each buffer back once the consumer asks for the next pool."""
        chunks = list(self.pools(pool_size, ids))
        tickets: List[Tuple[int, np.ndarray]] = []
        nxt = 0

        def top_up():
            nonlocal nxt
            while nxt < len(chunks) and len(tickets) < feed.depth:
                c = chunks[nxt]
                tickets.append((feed.submit([self.image_path[i] for i in c], self.image_index[c]), c))
                nxt += 1

        top_up()
        while tickets:
            ticket, c = tickets[0]
            images = feed.wait(ticket)
            try:
                yield c, self.pool(c, images, **pool_kw)
            finally:
                feed.release(ticket)
                tickets.pop(0)
            top_up()

    # -- setMetaDataSubset (ml_optimiser.cpp:10554-10640): results back into the table --
    def update(self, ids: Sequence[int], *, rot, tilt, psi, xoff, yoff, class_number, dLL, pmax, nr_significant,
               norm_correction=None) -> None:
        """Write a pool's assignments back: angles (degrees), offsets (pixels -> Angstrom), class (1-based), dLL, Pmax,
        number of significant samples and, when given, the new norm correction."""
        t = self.table
        ids = np.asarray(ids, np.int64)
        ps = self.optics.pixel_size[self.optics_group[ids]]

        def put(label, values, default):
            col = t.columns.get(label)
            if col is None:
                col = [default] * len(t)
                t.columns[label] = col
            for i, v in zip(ids, values):
                col[int(i)] = v

        put("rlnAngleRot", [float(v) for v in rot], 0.0)
        put("rlnAngleTilt", [float(v) for v in tilt], 0.0)
        put("rlnAnglePsi", [float(v) for v in psi], 0.0)
        put("rlnOriginXAngst", [float(v) for v in np.asarray(xoff, np.float64) * ps], 0.0)
        put("rlnOriginYAngst", [float(v) for v in np.asarray(yoff, np.float64) * ps], 0.0)
        put("rlnClassNumber", [int(v) for v in class_number], 0)
        put("rlnLogLikeliContribution", [float(v) for v in dLL], 0.0)
        put("rlnMaxValueProbDistribution", [float(v) for v in pmax], 0.0)
        put("rlnNrOfSignificantSamples", [int(v) for v in nr_significant], 0)
        if norm_correction is not None:
            put("rlnNormCorrection", [float(v) for v in norm_correction], 1.0)

    def write(self, path: str) -> None:
        star.write_star(path, [self.optics_table, self.table])
