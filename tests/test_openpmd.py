"""Checkpoint layout: record / attribute contract of the reference's openPMD plugin and a
bit-exact dump -> restore round trip (picongpu_b200/openpmd.py).  CPU: the tree and the JSON container on synthetic
rank parts; GPU: a run that is interrupted, dumped, restored into fresh contexts and continued equals the run that was
not interrupted."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from picongpu_b200 import openpmd  # noqa: E402
from picongpu_b200 import param as prm  # noqa: E402

LOCAL = (16, 16, 8)


def _params(rank, world, **kw):
    return prm.khi_params(grid=LOCAL, devices=(1, world, 1), rank_pos=(0, rank, 0), **kw)


def _synthetic_part(p, seed, slides=0, step=7):
    rng = np.random.default_rng(seed)
    n = p.grid
    off = [n[d] * p.rank_pos[d] for d in range(3)]
    tot = [off[0], off[1] + slides * n[1], off[2]]
    part = dict(grid=tuple(n), rank_pos=tuple(p.rank_pos), devices=tuple(p.devices), offset=tuple(off), total_offset=tuple(tot), slides=slides, step=step,
                E=rng.standard_normal((3, n[2], n[1], n[0])).astype(np.float32), B=rng.standard_normal((3, n[2], n[1], n[0])).astype(np.float32), species=[])
    for s, sp in enumerate(p.species):
        m = 100 + 37 * s + 11 * seed
        cell = np.sort(rng.integers(0, n[0] * n[1] * n[2], m)).astype(np.int64)
        tci = np.stack([cell % n[0] + tot[0], (cell // n[0]) % n[1] + tot[1], cell // (n[0] * n[1]) + tot[2]]).astype(np.int32)
        part["species"].append(dict(name=sp.name, position=rng.random((3, m)).astype(np.float32), positionOffset=tci,
                                    momentum=rng.standard_normal((3, m)).astype(np.float32), weighting=(1.0 + rng.random(m)).astype(np.float32), cell=cell.astype(np.int32)))
    return part


def test_checkpoint_tree_follows_the_reference_contract(tmp_path):
    world = 2
    ps = [_params(r, world) for r in range(world)]
    parts = [_synthetic_part(ps[r], seed=r) for r in range(world)]
    name = openpmd.write(ps[0], parts[::-1], str(tmp_path))  # rank order must not matter
    assert os.path.basename(name) == "checkpoint_7.json"
    raw = json.load(open(name))
    # openPMD-api JSON backend: groups carry "attributes", attributes are {"datatype", "value"}, datasets "datatype" + "data"
    assert set(raw) == {"attributes", "data", "platform_byte_widths"}
    a = raw["attributes"]
    assert a["openPMD"] == {"datatype": "STRING", "value": "1.1.0"} and a["openPMDextension"]["value"] == 1  # ED-PIC
    assert a["basePath"]["value"] == "/data/%T/" and a["meshesPath"]["value"] == "fields/" and a["particlesPath"]["value"] == "particles/"
    assert a["iterationEncoding"]["value"] == "fileBased" and "%T" in a["iterationFormat"]["value"]
    assert a["picongpuIOVersionMajor"]["value"] == 3  # plugins/common/openPMDVersion.def:47
    it = raw["data"]["7"]
    ia = it["attributes"]
    p = ps[0]
    assert ia["iteration"]["value"] == 7 and ia["sim_slides"]["value"] == 0  # read first by the restart (openPMDWriter.x.cpp:1292-1295)
    assert ia["dt"] == {"datatype": "FLOAT", "value": float(np.float32(p.dt))} and ia["timeUnitSI"]["value"] == p.unit_time
    assert abs(ia["time"]["value"] - 7 * p.dt) < 1e-6
    for k in ("unit_energy", "unit_length", "unit_speed", "unit_time", "unit_mass", "unit_charge", "unit_efield", "unit_bfield"):
        assert ia[k] == {"datatype": "DOUBLE", "value": getattr(p, k)}
    assert [ia[k]["value"] for k in ("cell_width", "cell_height", "cell_depth")] == [float(np.float32(c)) for c in p.cell_size]
    # meshes: F[z][y][x] of the global domain
    E = it["fields"]["E"]
    assert E["attributes"]["axisLabels"]["value"] == ["z", "y", "x"] and E["attributes"]["geometry"]["value"] == "cartesian"
    assert E["attributes"]["dataOrder"]["value"] == "C" and E["attributes"]["gridUnitSI"]["value"] == p.unit_length
    assert E["attributes"]["unitDimension"]["value"] == [1.0, 1.0, -3.0, -1.0, 0.0, 0.0, 0.0]  # V/m
    assert it["fields"]["B"]["attributes"]["unitDimension"]["value"] == [0.0, 1.0, -2.0, -1.0, 0.0, 0.0, 0.0]  # T
    assert E["x"]["datatype"] == "FLOAT" and np.array(E["x"]["data"]).shape == (LOCAL[2], LOCAL[1] * world, LOCAL[0])
    assert E["x"]["attributes"]["position"]["value"] == [0.5, 0.0, 0.0] and it["fields"]["B"]["x"]["attributes"]["position"]["value"] == [0.0, 0.5, 0.5]
    assert E["y"]["attributes"]["unitSI"]["value"] == p.unit_efield
    # species: the frame attributes minus multiMask / localCellIdx plus totalCellIdx (WriteSpecies.hpp:205-215)
    e = it["particles"]["e"]
    assert {"position", "positionOffset", "momentum", "weighting", "mass", "charge", "particlePatches"} <= set(e)
    assert e["position"]["x"]["datatype"] == "FLOAT" and e["positionOffset"]["y"]["datatype"] == "INT"
    cs = [float(np.float32(c)) * p.unit_length for c in p.cell_size]
    assert e["position"]["y"]["attributes"]["unitSI"]["value"] == cs[1] and e["positionOffset"]["z"]["attributes"]["unitSI"]["value"] == cs[2]
    assert e["momentum"]["x"]["attributes"]["unitSI"]["value"] == p.unit_mass * p.unit_speed
    assert e["momentum"]["attributes"]["macroWeighted"]["value"] == 1 and e["momentum"]["attributes"]["weightingPower"]["value"] == 1.0
    assert e["position"]["attributes"]["macroWeighted"]["value"] == 0 and e["position"]["attributes"]["weightingPower"]["value"] == 0.0
    assert e["momentum"]["attributes"]["unitDimension"]["value"] == [1.0, 1.0, -1.0, 0.0, 0.0, 0.0, 0.0]
    assert e["weighting"]["datatype"] == "FLOAT" and e["weighting"]["attributes"]["unitSI"]["value"] == 1.0
    n_e = sum(q["species"][0]["weighting"].shape[0] for q in parts)
    assert len(e["weighting"]["data"]) == n_e
    # constant records: no dataset, `value` and `shape`; electron charge is negative, ion mass 1836 electron masses
    i = it["particles"]["i"]
    assert "data" not in e["charge"] and e["charge"]["attributes"]["shape"]["value"] == [n_e]
    assert e["charge"]["attributes"]["value"]["value"] * p.unit_charge == pytest.approx(prm.ELECTRON_CHARGE_SI * p.typical_num_particles_per_macro, rel=1e-6)
    assert i["mass"]["attributes"]["value"]["value"] / e["mass"]["attributes"]["value"]["value"] == pytest.approx(1836.152672, rel=1e-6)
    # patches: one per rank, ordered along y, in cells of the total domain
    pt = e["particlePatches"]
    assert pt["offset"]["y"]["data"] == [0, LOCAL[1]] and pt["extent"]["y"]["data"] == [LOCAL[1]] * 2 and pt["extent"]["x"]["data"] == [LOCAL[0]] * 2
    assert pt["numParticles"]["data"] == [q["species"][0]["weighting"].shape[0] for q in parts]
    assert pt["numParticlesOffset"]["data"] == [0, parts[0]["species"][0]["weighting"].shape[0]]


@pytest.mark.parametrize("slides", [0, 3])
def test_checkpoint_round_trip_is_bit_exact(tmp_path, slides):
    world = 3
    ps = [_params(r, world, periodic=(1, 0, 1), moving_window=1) for r in range(world)]
    parts = [_synthetic_part(ps[r], seed=5 + r, slides=slides, step=1234) for r in range(world)]
    tree = openpmd.read(openpmd.write(ps[0], parts, str(tmp_path)))
    for r in range(world):
        back = openpmd.rank_part(tree, ps[r])
        assert back["step"] == 1234 and back["slides"] == slides
        assert np.array_equal(back["E"], parts[r]["E"]) and np.array_equal(back["B"], parts[r]["B"])
        for s in range(2):
            a, b = back["species"][s], parts[r]["species"][s]
            for k in ("position", "momentum", "weighting", "cell"):
                assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), (r, s, k)
    # a rank whose local domain is not one of the patches cannot restart (LoadSpecies.hpp:251-253)
    other = prm.khi_params(grid=(16, 24, 8), devices=(1, 2, 1), rank_pos=(0, 1, 0), periodic=(1, 0, 1))
    with pytest.raises(ValueError):
        openpmd.rank_part(tree, other)
    bad = json.loads(json.dumps(openpmd._to_jsonable(tree)))
    bad["attributes"]["picongpuIOVersionMajor"]["value"] = 2
    with pytest.raises(ValueError):
        openpmd.check_restart_compatibility(bad)


def _gloo_worker(rank, world, port, outdir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = _params(rank, world, periodic=(1, 0, 1), moving_window=1)
    part = _synthetic_part(p, seed=20 + rank, slides=1, step=40)
    name = openpmd.write_distributed(p, part, outdir, rank, world)
    back = openpmd.rank_part(openpmd.read(name), p)
    ok = np.array_equal(back["E"], part["E"]) and np.array_equal(back["B"], part["B"])
    for a, b in zip(back["species"], part["species"]):
        ok = ok and all(np.array_equal(a[k], b[k]) for k in ("position", "momentum", "weighting", "cell"))
    open(os.path.join(outdir, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_checkpoint_two_ranks_gloo(tmp_path):
    """One process per rank: parts gathered on rank 0, one file, every rank finds its own patch again."""
    import torch.multiprocessing as mp

    world = 2
    port = 29500 + (os.getpid() % 2000) + 41
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert [open(os.path.join(str(tmp_path), "ok%d" % r)).read() for r in range(world)] == ["1", "1"]
    tree = openpmd.read(openpmd.file_name(str(tmp_path), 40))
    assert tree["data"]["40"]["particles"]["e"]["particlePatches"]["offset"]["y"]["data"].tolist() == [LOCAL[1], 2 * LOCAL[1]]  # one slide


@pytest.mark.gpu
def test_restart_continues_the_run(orc, tmp_path):
    """KHI plasma with a laser entering (the source depends on the restored step counter): 12 steps in one go against
    6 steps, checkpoint, restore into a fresh context, 6 more steps."""
    import util
    from picongpu_b200 import picstep

    kw = dict(periodic=(1, 0, 1), absorber_kind=1, absorber_cells=((0, 0), (6, 6), (0, 0)), absorber_strength=((0, 0), (1e-3, 1e-3), (0, 0)))
    p = util.make_params((16, 48, 8), **kw)
    p.laser = prm.plane_wave_laser(p, a0=0.5, pulse_duration_si=2e-15, ramp_init=2.0, offset_ymin=8)
    _, e, i = util.khi_ic(orc, p, ppc_dim=(2, 2, 1))

    def fresh():
        s = picstep.Simulation(p, exact=True)
        return s

    ref = fresh()
    for name, sp in (("e", e), ("i", i)):
        ref.upload_particles(name, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    ref.step(6)
    ref.sync()
    name = openpmd.write(p, [openpmd.collect(ref)], str(tmp_path))
    ref.step(6)
    ref.sync()
    tree = openpmd.read(name)
    assert set(tree["data"]) == {"6"}
    new = fresh()
    part = openpmd.restore(new, tree)
    assert new.step_index == 6 and part["species"][0]["weighting"].shape[0] == new.particle_count("e")
    new.step(6)
    new.sync()
    amp = abs(p.laser["amplitude"])
    for f in (picstep.FIELD_E, picstep.FIELD_B):
        a, b = ref.download_field(f), new.download_field(f)
        assert np.abs(a).max() > 0.05 * amp / (1.0 if f == picstep.FIELD_E else p.c)
        # the restored particles come back in the dumped order; arrivals of the following steps are appended in the
        # order of a global atomic, so the deposition sums may differ in their last bits
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()
    for sp in ("e", "i"):
        a, b = ref.download_particles(sp), new.download_particles(sp)
        assert a[2].shape == b[2].shape
        ka, kb = np.lexsort((a[2], a[3])), np.lexsort((b[2], b[3]))
        assert np.array_equal(a[3][ka], b[3][kb]) and np.array_equal(a[2][ka], b[2][kb])
        assert np.abs(np.sort(a[1][0]) - np.sort(b[1][0])).max() <= 2e-6 * np.abs(a[1][0]).max()
    ref.close()
    new.close()
