"""Particle I/O: MRC stacks, STAR tables, the metadata feed.  Host-only code: runs without a GPU.

The MRC cases are built byte by byte from the format the reference reads (src/rwMRC.h:82-283), not with our own writer,
so that reader and writer are pinned independently; the STAR cases use the layout RELION 3.1+ writes
(src/metadata_table.cpp:1366-1519) and the parsing rules of readStarLoop / readStarList / nextTokenInSTAR.
"""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from relion_b200 import capi, particle_io, star  # noqa: E402


def _mrc_bytes(data, mode, big_endian=False, nsymbt=0, pixel_size=1.25):
    """An MRC file image: 1024-byte header (+ nsymbt bytes of extended header) + data, in either byte order."""
    nz, ny, nx = data.shape
    e = ">" if big_endian else "<"
    hdr = struct.pack(e + "10i6f3i3f2i", nx, ny, nz, mode, 0, 0, 0, nx, ny, nz, pixel_size * nx, pixel_size * ny, pixel_size * nz,
                      90.0, 90.0, 90.0, 1, 2, 3, float(data.min()), float(data.max()), float(data.mean()), 0, nsymbt)
    hdr += b"\0" * (4 * 25) + struct.pack(e + "3f", 0.0, 0.0, 0.0) + b"MAP " + (b"\x11\x11\0\0" if big_endian else b"\x44\x41\0\0")
    hdr += struct.pack(e + "fi", 1.0, 0) + b"\0" * 800
    assert len(hdr) == 1024
    dt = {0: "i1", 1: "i2", 2: "f4", 6: "u2", 12: "f2"}[mode]
    return hdr + b"\xab" * nsymbt + data.astype(e + dt if dt != "i1" else dt).tobytes()


@pytest.mark.parametrize("mode,big", [(2, False), (2, True), (1, False), (1, True), (6, False), (0, False), (12, False), (12, True)])
def test_mrc_reader_modes_and_byte_order(tmp_path, mode, big):
    rng = np.random.default_rng(mode * 2 + big)
    if mode == 2:
        data = rng.standard_normal((5, 6, 8)).astype(np.float32)
    elif mode == 12:
        data = np.concatenate([rng.standard_normal(235) * 10, [0.0, -0.0, 6.1e-5, 5.96e-8, 65504.0]]).astype(np.float16).reshape(5, 6, 8)   # + denormals
    elif mode == 0:
        data = rng.integers(-128, 128, (5, 6, 8)).astype(np.int8)                     # SIGNED since RELION 3.1 (rwMRC.h:201)
    elif mode == 1:
        data = rng.integers(-32768, 32768, (5, 6, 8)).astype(np.int16)
    else:
        data = rng.integers(0, 65536, (5, 6, 8)).astype(np.uint16)
    p = tmp_path / "stack.mrcs"
    p.write_bytes(_mrc_bytes(data, mode, big_endian=big, nsymbt=160 if mode == 1 else 0))
    with particle_io.MrcStack(str(p)) as m:
        assert (m.nx, m.ny, m.nz, m.mode) == (8, 6, 5, mode)
        assert abs(m.pixel_size - 1.25) < 1e-6
        got = m.read()
        np.testing.assert_array_equal(got, data.astype(np.float32))
        np.testing.assert_array_equal(m.read([4, 0, 4]), data.astype(np.float32)[[4, 0, 4]])


def test_mrc_reader_refuses_what_the_reference_refuses(tmp_path):
    data = np.zeros((2, 4, 4), np.float32)
    for mode in (3, 4, 101, 7):
        p = tmp_path / f"m{mode}.mrc"
        raw = bytearray(_mrc_bytes(data, 2))
        raw[12:16] = struct.pack("<i", mode)
        p.write_bytes(bytes(raw))
        with pytest.raises(capi.RelionB200Error):
            particle_io.MrcStack(str(p))
    p = tmp_path / "short.mrcs"
    p.write_bytes(_mrc_bytes(data, 2)[:-8])                                            # truncated data
    with pytest.raises(capi.RelionB200Error, match="truncated"):
        particle_io.MrcStack(str(p))
    with pytest.raises(capi.RelionB200Error):
        particle_io.MrcStack(str(tmp_path / "missing.mrcs"))
    p = tmp_path / "ok.mrcs"
    p.write_bytes(_mrc_bytes(data, 2))
    with particle_io.MrcStack(str(p)) as m:
        with pytest.raises(capi.RelionB200Error, match="exceeds stack size"):          # rwMRC.h:176
            m.read([2])


def test_mrc_writer_header_and_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    data = rng.standard_normal((3, 10, 12)).astype(np.float32)
    p = str(tmp_path / "out.mrcs")
    particle_io.write_mrc(p, data, pixel_size=0.83)
    raw = open(p, "rb").read()
    assert len(raw) == 1024 + data.nbytes
    nx, ny, nz, mode, _, _, _, mx, my, mz = struct.unpack("<10i", raw[:40])
    assert (nx, ny, nz, mode, mx, my, mz) == (12, 10, 3, 2, 12, 10, 3)
    a, b, c, al, be, ga = struct.unpack("<6f", raw[40:64])
    np.testing.assert_allclose([a, b, c], [0.83 * 12, 0.83 * 10, 0.83 * 3], rtol=1e-6)
    assert (al, be, ga) == (90.0, 90.0, 90.0)
    assert struct.unpack("<3i", raw[64:76]) == (1, 2, 3)
    amin, amax, amean = struct.unpack("<3f", raw[76:88])
    np.testing.assert_allclose([amin, amax, amean], [data.min(), data.max(), data.mean()], rtol=1e-5, atol=1e-6)
    assert struct.unpack("<i", raw[92:96])[0] == 0 and raw[208:212] == b"MAP " and raw[212:214] == b"\x44\x41"
    np.testing.assert_allclose(struct.unpack("<f", raw[216:220])[0], data.std(ddof=1), rtol=1e-5)
    with particle_io.MrcStack(p) as m:
        np.testing.assert_array_equal(m.read(), data)
        assert abs(m.pixel_size - 0.83) < 1e-6


STAR_31 = """
# version 30001

data_optics

loop_
_rlnOpticsGroupName #1
_rlnOpticsGroup #2
_rlnVoltage #3
_rlnSphericalAberration #4
_rlnAmplitudeContrast #5
_rlnImagePixelSize #6
_rlnImageSize #7
_rlnImageDimensionality #8
opticsGroup1            1   300.000000     2.700000     0.100000     1.250000           16            2
opticsGroup2            2   200.000000     2.000000     0.070000     2.500000           16            2


# version 30001

data_particles

loop_
_rlnImageName #1
_rlnMicrographName #2
_rlnOpticsGroup #3
_rlnDefocusU #4
_rlnDefocusV #5
_rlnDefocusAngle #6
_rlnOriginXAngst #7
_rlnOriginYAngst #8
_rlnAngleRot #9
_rlnAngleTilt #10
_rlnAnglePsi #11
_rlnRandomSubset #12
_rlnNormCorrection #13
_rlnUnknownThing #14
000002@stackB.mrcs MotionCorr/job002/Movies/micB.mrc            2 21000.000000 20500.000000    35.000000     2.500000    -5.000000    10.000000    20.000000    30.000000            2     0.800000 'quoted value'
000001@stackA.mrcs MotionCorr/job002/Movies/micA.mrc            1 11000.000000 10500.000000    15.000000     1.250000    -2.500000   -10.000000    40.000000   130.000000            1     1.250000 x # trailing comment
000001@stackB.mrcs MotionCorr/job002/Movies/micB.mrc            2 22000.000000 21500.000000    45.000000     0.000000     0.000000     0.000000     0.000000     0.000000            1     1.000000 y
000003@stackA.mrcs MotionCorr/job002/Movies/micA.mrc            1 12000.000000 11500.000000    25.000000 1.000000e-04     3.750000    50.000000    60.000000    70.000000            2     1.000000 z

"""


def test_star_parser_follows_the_reference_rules(tmp_path):
    tables = star.parse_star(STAR_31)
    assert [t.name for t in tables] == ["optics", "particles"]
    opt, par = tables
    assert opt.version == 30001 and len(opt) == 2 and len(par) == 4
    assert par.labels()[0] == "rlnImageName" and par.labels()[-1] == "rlnUnknownThing"
    assert par.columns["rlnOpticsGroup"] == [2, 1, 2, 1]                               # integers stay integers
    assert par.columns["rlnOriginXAngst"][3] == pytest.approx(1e-4)
    assert par.columns["rlnUnknownThing"] == ["quoted value", "x", "y", "z"]           # quotes group, '#' ends the line
    with pytest.raises(star.StarError, match="more columns"):
        star.parse_star("data_x\n\nloop_\n_rlnA #1\n_rlnB #2\n_rlnC #3\n1 2 3 4\n")
    with pytest.raises(star.StarError, match="fewer columns"):
        star.parse_star("data_x\n\nloop_\n_rlnA #1\n_rlnB #2\n_rlnC #3\n1 2\n")
    assert star.parse_star("data_x\n\nloop_\n_rlnA #1\n_rlnB #2\nfn_mtf\n")[0].columns["rlnB"] == [""]   # two-column exception
    with pytest.raises(star.StarError, match="CR\\+LF"):
        star.parse_star("data_x\r\n_rlnA 1\r\n")
    # list block followed by a loop block, as in model.star
    t = star.parse_star("data_model_general\n\n_rlnReferenceDimensionality 3\n_rlnCurrentResolution 7.5\n\ndata_model_classes\n\nloop_\n_rlnReferenceImage #1\nclass001.mrc\n")
    assert t[0].is_list and t[0].value("rlnReferenceDimensionality") == 3 and t[0].value("rlnCurrentResolution") == 7.5
    assert not t[1].is_list and t[1].columns["rlnReferenceImage"] == ["class001.mrc"]


def test_star_writer_matches_the_reference_layout_and_roundtrips(tmp_path):
    tables = star.parse_star(STAR_31)
    txt = star.format_star(tables)
    lines = txt.split("\n")
    assert lines[1] == "# version 50001" and lines[3] == "data_optics" and lines[5] == "loop_ " and lines[6] == "_rlnOpticsGroupName #1 "
    # getValueToString: %12.6f, %12.5f when negative, scientific below 1e-3 (metadata_table.cpp:252-278)
    row = [ln for ln in lines if ln.startswith("000003@stackA.mrcs")][0]
    assert "12000.000000" in row and "1.000000e-04" in row and "    3.750000" in row
    row = [ln for ln in lines if ln.startswith("000001@stackA.mrcs")][0]
    assert "    -2.50000 " in row and "   -10.00000 " in row
    again = star.parse_star(txt)
    for a, b in zip(tables, again):
        assert a.name == b.name and a.columns == b.columns
    p = str(tmp_path / "x.star")
    star.write_star(p, tables)
    assert star.read_star(p, "particles").columns == tables[1].columns
    assert star.read_star(p, "").name == "optics"
    with pytest.raises(KeyError):
        star.read_star(p, "nothing")


def _write_dataset(tmp_path, n=16):
    rng = np.random.default_rng(11)
    A = rng.standard_normal((3, n, n)).astype(np.float32)
    B = rng.standard_normal((2, n, n)).astype(np.float32)
    particle_io.write_mrc(str(tmp_path / "stackA.mrcs"), A, 1.25)
    particle_io.write_mrc(str(tmp_path / "stackB.mrcs"), B, 2.5)
    p = tmp_path / "particles.star"
    p.write_text(STAR_31)
    return str(p), A, B


def test_particle_set_layout_follows_experiment_read(tmp_path):
    path, A, B = _write_dataset(tmp_path)
    ps = particle_io.ParticleSet.read(path)
    # sorted on micrograph name (stable): micA rows first, in file order
    assert [os.path.basename(p) for p in ps.image_path] == ["stackA.mrcs", "stackA.mrcs", "stackB.mrcs", "stackB.mrcs"]
    assert ps.image_index.tolist() == [0, 2, 1, 0]                                      # names count from 1
    assert ps.group_names == ["Movies/micA.mrc", "Movies/micB.mrc"] and ps.group_id.tolist() == [0, 0, 1, 1]
    assert ps.optics_group.tolist() == [0, 0, 1, 1] and ps.image_size() == 16
    assert ps.half_set(1).tolist() == [0, 3] and ps.half_set(2).tolist() == [1, 2]
    md = ps.metadata(np.arange(4))
    assert md.shape == (4, particle_io.METADATA_LINE_LENGTH)
    # offsets in pixels of the particle's optics group (ml_optimiser.cpp:10427-10431)
    np.testing.assert_allclose(md[:, particle_io.METADATA_XOFF], [1.0, 1e-4 / 1.25, 1.0, 0.0])
    np.testing.assert_allclose(md[:, particle_io.METADATA_YOFF], [-2.0, 3.0, -2.0, 0.0])
    np.testing.assert_allclose(md[:, particle_io.METADATA_PSI], [130.0, 70.0, 30.0, 0.0])
    np.testing.assert_allclose(md[:, particle_io.METADATA_CTF_DEFOCUS_U], [11000.0, 12000.0, 21000.0, 22000.0])
    assert np.all(md[:, particle_io.METADATA_CTF_KFACTOR] == 1.0) and np.all(md[:, particle_io.METADATA_CTF_BFACTOR] == 0.0)
    for k in (particle_io.METADATA_ROT_PRIOR, particle_io.METADATA_XOFF_PRIOR, particle_io.METADATA_PSI_PRIOR_FLIP_RATIO):
        assert np.all(md[:, k] == 999.0)                                               # unset priors (:10452-10487)
    np.testing.assert_allclose(md[:, particle_io.METADATA_NORM], [1.25, 1.0, 0.8, 1.0])


def test_feed_streams_pools_in_order_and_reuses_buffers(tmp_path):
    path, A, B = _write_dataset(tmp_path)
    ps = particle_io.ParticleSet.read(path)
    feed = particle_io.ParticleFeed(image_size=16, max_particles=3, depth=2, n_threads=3)
    want = np.stack([A[0], A[2], B[1], B[0]])
    seen = []
    for _ in range(3):                                                                  # several epochs through two buffers
        for ids, pool in ps.stream(feed, pool_size=3, avg_norm_correction=2.0):
            np.testing.assert_array_equal(np.asarray(pool.images), want[ids])
            np.testing.assert_allclose(pool.norm_factor, 2.0 / np.array([1.25, 1.0, 0.8, 1.0])[ids])
            np.testing.assert_allclose(pool.prior_offset, 0.0)                         # 999 -> 0 (acc_ml_optimiser_impl.h:56-58)
            assert pool.group_id.tolist() == ps.group_id[ids].tolist()
            np.testing.assert_allclose(pool.og_kV, [300.0, 200.0])
            seen.append(ids.tolist())
    assert seen == [[0, 1, 2], [3]] * 3
    # all buffers busy -> RB_ERR_STATE; a wrong-sized stack or a missing image -> error at wait()
    t1 = feed.submit([ps.image_path[0]], [0])
    t2 = feed.submit([ps.image_path[0]], [1])
    with pytest.raises(capi.RelionB200Error) as e:
        feed.submit([ps.image_path[0]], [2])
    assert e.value.status == capi.RB_ERR_STATE
    feed.wait(t1), feed.release(t1), feed.release(t2)
    t3 = feed.submit([ps.image_path[0], ps.image_path[2]], [0, 7])
    with pytest.raises(capi.RelionB200Error, match="exceeds stack size"):
        feed.wait(t3)
    feed.release(t3)
    particle_io.write_mrc(str(tmp_path / "small.mrcs"), np.zeros((1, 8, 8), np.float32))
    t4 = feed.submit([str(tmp_path / "small.mrcs")], [0])
    with pytest.raises(capi.RelionB200Error, match="incorrect image size"):
        feed.wait(t4)
    feed.release(t4)
    with pytest.raises(capi.RelionB200Error):
        feed.submit([ps.image_path[0]] * 4, [0] * 4)                                    # more than max_particles
    feed.close()


def test_results_go_back_into_the_particle_table(tmp_path):
    path, _, _ = _write_dataset(tmp_path)
    ps = particle_io.ParticleSet.read(path)
    ps.update([1, 2], rot=[1.0, 2.0], tilt=[3.0, 4.0], psi=[5.0, 6.0], xoff=[2.0, -1.0], yoff=[0.5, 0.25], class_number=[2, 1],
              dLL=[-1234.5, -2345.6], pmax=[0.9, 0.1], nr_significant=[12, 345], norm_correction=[1.1, 0.9])
    out = str(tmp_path / "run_it001_data.star")
    ps.write(out)
    again = particle_io.ParticleSet.read(out, root=str(tmp_path))
    t = again.table
    np.testing.assert_allclose(t.column("rlnOriginXAngst", np.float64), [1.25, 2.0 * 1.25, -1.0 * 2.5, 0.0])     # pixels -> Angstrom
    np.testing.assert_allclose(t.column("rlnAnglePsi", np.float64), [130.0, 5.0, 6.0, 0.0])
    assert t.column("rlnClassNumber", np.int64).tolist() == [0, 2, 1, 0]
    assert t.column("rlnNrOfSignificantSamples", np.int64).tolist() == [0, 12, 345, 0]
    np.testing.assert_allclose(t.column("rlnNormCorrection", np.float64), [1.25, 1.1, 0.9, 1.0])
    assert again.group_id.tolist() == ps.group_id.tolist() and again.image_index.tolist() == ps.image_index.tolist()


def test_feed_keeps_a_bounded_number_of_stacks_open(tmp_path, monkeypatch):
    """Particle sets span thousands of per-micrograph stacks: the feed keeps at most RB_FEED_MAX_OPEN of them open (least
    recently used closed first; the reference keeps one, src/ml_optimiser.cpp:10369-10377), so the descriptor count stays
    bounded while every image still arrives."""
    monkeypatch.setenv("RB_FEED_MAX_OPEN", "3")
    n, per, nstacks = 8, 4, 20
    rng = np.random.default_rng(3)
    stacks = []
    for s in range(nstacks):
        imgs = rng.standard_normal((per, n, n)).astype(np.float32)
        path = str(tmp_path / ("mic%03d.mrcs" % s))
        particle_io.write_mrc(path, imgs, 1.0)
        stacks.append((path, imgs))
    fds0 = len(os.listdir("/proc/self/fd"))
    feed = particle_io.ParticleFeed(image_size=n, max_particles=per * 2, depth=2, n_threads=3)
    peak = 0
    for rep in range(2):                                   # second sweep re-opens evicted stacks
        for s in range(0, nstacks, 2):
            paths = [stacks[s][0]] * per + [stacks[s + 1][0]] * per
            idx = list(range(per)) * 2
            t = feed.submit(paths, idx)
            got = feed.wait(t)
            np.testing.assert_array_equal(got[:per], stacks[s][1])
            np.testing.assert_array_equal(got[per:], stacks[s + 1][1])
            feed.release(t)
            peak = max(peak, len(os.listdir("/proc/self/fd")) - fds0)
    feed.close()
    assert peak <= 3 + 3, peak                             # the cache limit plus stacks still held by the reader threads
    assert len(os.listdir("/proc/self/fd")) <= fds0 + 1
