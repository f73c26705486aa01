"""GPU test of the N>1 path through the C ABI (N-GPU result == 1-GPU result): two ranks, one
B200 each, 1-D slab decomposition in y, guard exchange + particle migration over NCCL send/recv (comm.cu), must
reproduce the single-domain oracle run on the same global grid.  Needs two GPUs (`gpurun --gpus 2`); skipped on a
one-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from picongpu_b200 import param as prm  # noqa: E402
from picongpu_b200 import picstep  # noqa: E402

import util  # noqa: E402
from test_multirank_cpu import _kick  # noqa: E402

pytestmark = pytest.mark.gpu
LOCAL = (16, 32, 8)  # four supercell layers along the split axis: BORDER (2) and CORE (2) areas both exist


OPEN = dict(periodic=(1, 0, 1), current_interpolation=1, absorber_kind=1, absorber_cells=((0, 0), (6, 6), (0, 0)),
            absorber_strength=((0, 0), (0.05, 0.05), (0, 0)))  # LWFA-like boundaries: open + absorbing along the split axis


# examples/LaserWakefield's boundaries and source: nothing periodic, PML on every outer face, a GaussianPulse (three
# Laguerre modes) entering through a Huygens surface five cells below the rank boundary, so that the pulse crosses it
_P0 = prm.khi_params(grid=LOCAL)
LWFA = dict(periodic=(0, 0, 0), absorber_kind=2, absorber_cells=((4, 4), (6, 6), (2, 2)), pml=prm.pml_params(_P0),
            laser=prm.gaussian_pulse_laser(_P0, a0=0.5, pulse_duration_si=3e-15, w0_si=0.4e-6, pulse_init=1.0, focus_position_si=(0.0, 3.5e-6, 0.0),
                                           polarisation="circular", position=((5, -5), (26, -8), (2, -2)), modes=(0.7, 0.2, 0.1)))


def _worker(rank, world, port, steps, fused, outdir, kw):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import nvidia.nccl as _n

        os.environ.setdefault("PICSTEP_NCCL_LIB", os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2"))
    except Exception:
        pass
    from oracle import picoracle as orc

    orc.lib().orc_set_num_threads(2)
    p = prm.khi_params(grid=LOCAL, devices=(1, world, 1), rank_pos=(0, rank, 0), **kw)
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    sim = picstep.Simulation(p, device=rank, exact=False, unfused=not fused)
    box = [sim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim.comm_init(box[0], rank, world)
    for name, sp in (("e", e), ("i", i)):
        sim.upload_particles(name, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    n0 = sim.particle_count("e")
    sim.step(steps)
    sim.sync()
    E, B = sim.download_field(picstep.FIELD_E), sim.download_field(picstep.FIELD_B)
    pe = sim.download_particles("e")
    pi = sim.download_particles("i")
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), E=o.interior(E), B=o.interior(B), ne=pe[2].shape[0], ni=pi[2].shape[0], n0=n0,
             eux=np.sort(pe[1][0]), iux=np.sort(pi[1][0]), ecell=np.sort(pe[3]), gauss=sim.gauss_residual())
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fused,variant", [(True, "periodic"), (False, "periodic"), (True, "open"), (True, "lwfa")])
def test_two_gpu_slab_decomposition_equals_single_domain(orc, tmp_path, fused, variant):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, steps = 2, (16 if variant == "lwfa" else 6)
    kw = {"open": OPEN, "lwfa": LWFA}.get(variant, {})
    port = 29500 + (os.getpid() % 2000) + (7 if fused else 0) + (13 if kw else 0) + (17 if variant == "lwfa" else 0)
    mp.spawn(_worker, args=(world, port, steps, fused, str(tmp_path), kw), nprocs=world, join=True)
    p = prm.khi_params(grid=(LOCAL[0], LOCAL[1] * world, LOCAL[2]), **kw)
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    E, B, J = o.field(), o.field(), o.field()
    for _ in range(steps):
        if kw:
            o.step_open(E, B, J, [e, i])
        else:
            o.step(E, B, J, [e, i])
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]
    Eg = np.concatenate([r[0]["E"], r[1]["E"]], axis=2)
    Bg = np.concatenate([r[0]["B"], r[1]["B"]], axis=2)
    _, escale = util.khi_scales(p, 1)
    if variant == "lwfa":
        amp = abs(p.laser["amplitude"])
        escale = max(escale, amp)
        # the pulse has crossed from rank 0 (which holds the Huygens surface) into rank 1
        assert np.abs(r[1]["E"]).max() > 0.03 * amp
    # production build vs oracle, tolerance as in test_khi_100_steps_vs_oracle (2e-5 of the per-species drive scale)
    assert np.abs(Eg - o.interior(E)).max() / escale < 2e-5
    assert np.abs(Bg - o.interior(B)).max() / escale < 2e-5
    assert int(r[0]["ne"]) + int(r[1]["ne"]) == e["w"].shape[0]
    assert int(r[0]["ni"]) + int(r[1]["ni"]) == i["w"].shape[0]
    # particles really crossed the slab boundary (and, with open faces, some were absorbed)
    assert int(r[0]["ne"]) != int(r[0]["n0"]) or int(r[1]["ne"]) != int(r[1]["n0"])
    if kw:
        assert int(r[0]["ne"]) + int(r[1]["ne"]) < int(r[0]["n0"]) + int(r[1]["n0"])
    for key, ref in (("eux", e["mom"][0]), ("iux", i["mom"][0])):
        a, b = np.sort(np.concatenate([r[0][key], r[1][key]])), np.sort(ref)
        assert np.abs(a - b).max() / np.abs(b).max() < 5e-6
    if kw:
        return  # the Gauss check below assumes a closed (periodic) box
    # Gauss residual stays at round-off on both ranks (charge conserving deposition across the rank boundary)
    q = 25.0 * abs(p.base_charge) * p.typical_num_particles_per_macro
    assert max(float(r[0]["gauss"]), float(r[1]["gauss"])) / q < 1e-4


# ---- moving window: GridController::slide + Simulation::slide ------------------------------------------------------
MW = dict(periodic=(1, 0, 1), moving_window=1, absorber_kind=1, absorber_cells=((0, 0), (6, 6), (0, 0)),
          absorber_strength=((0, 0), (0.05, 0.05), (0, 0)))


def _new_slab(orc, p_top):
    """Plasma of the slab that enters the window: the KHI recipe with another seed, kept two cells away from the
    slab faces (the fresh top rank has no guard data from its lower neighbour before the first exchange)."""
    _, e, i = util.khi_ic(orc, p_top, seed=77)
    out = []
    for sp in (e, i):
        cy = (sp["cell"] // p_top.grid[0]) % p_top.grid[1]
        keep = (cy >= 2) & (cy < p_top.grid[1] - 2)
        out.append(dict(massRatio=sp["massRatio"], chargeRatio=sp["chargeRatio"], pos=np.ascontiguousarray(sp["pos"][:, keep]),
                        mom=np.ascontiguousarray(sp["mom"][:, keep]), w=np.ascontiguousarray(sp["w"][keep]), cell=np.ascontiguousarray(sp["cell"][keep])))
    return out


def _mw_worker(rank, world, port, steps, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import nvidia.nccl as _n

        os.environ.setdefault("PICSTEP_NCCL_LIB", os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2"))
    except Exception:
        pass
    from oracle import picoracle as orc

    orc.lib().orc_set_num_threads(2)
    p = prm.khi_params(grid=LOCAL, devices=(1, world, 1), rank_pos=(0, rank, 0), **MW)
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    sim = picstep.Simulation(p, device=rank, exact=False)
    box = [sim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sim.comm_init(box[0], rank, world)
    for name, sp in (("e", e), ("i", i)):
        sim.upload_particles(name, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    sim.step(steps)
    was_reset = sim.slide()
    assert was_reset == (rank == 0)  # the lowest rank becomes the top of the window
    assert sim.p.rank_pos[1] == (rank - 1) % world
    if was_reset:
        assert sim.particle_count("e") == 0 and np.abs(sim.download_field(picstep.FIELD_E)).max() == 0.0
        p_top = prm.khi_params(grid=LOCAL, devices=(1, world, 1), rank_pos=(0, world - 1, 0), **MW)
        for name, sp in zip(("e", "i"), _new_slab(orc, p_top)):
            sim.upload_particles(name, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    sim.step(steps)
    sim.sync()
    E, B = sim.download_field(picstep.FIELD_E), sim.download_field(picstep.FIELD_B)
    pe = sim.download_particles("e")
    np.savez(os.path.join(outdir, "mw%d.npz" % rank), E=o.interior(E), B=o.interior(B), ne=pe[2].shape[0], eux=np.sort(pe[1][0]), pos=sim.p.rank_pos[1])
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_moving_window_slide(orc, tmp_path):
    """Slide once between two blocks of steps on two GPUs; the single-domain oracle emulates the slide by shifting the
    global arrays down by one local domain and dropping / adding the particles of the leaving / entering slab."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, steps = 2, 4
    port = 29500 + (os.getpid() % 2000) + 29
    mp.spawn(_mw_worker, args=(world, port, steps, str(tmp_path)), nprocs=world, join=True)
    ny = LOCAL[1]
    p = prm.khi_params(grid=(LOCAL[0], ny * world, LOCAL[2]), **MW)
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    E, B, J = o.field(), o.field(), o.field()
    sps = [e, i]
    for _ in range(steps):
        o.step_open(E, B, J, sps)
    g = p.guard_cells
    for F in (E, B):  # shift down by one local domain; the entering slab is empty
        F[:, :, : F.shape[2] - ny, :] = F[:, :, ny:, :].copy()
        F[:, :, g[1] + ny * (world - 1):, :] = 0.0
    p_top = prm.khi_params(grid=LOCAL, devices=(1, world, 1), rank_pos=(0, world - 1, 0), **MW)
    fresh = _new_slab(orc, p_top)
    for sp, new in zip(sps, fresh):
        n = p.grid
        cx, cy, cz = sp["cell"] % n[0], (sp["cell"] // n[0]) % n[1], sp["cell"] // (n[0] * n[1])
        keep = cy >= ny
        lc = new["cell"]
        lx, ly, lz = lc % LOCAL[0], (lc // LOCAL[0]) % LOCAL[1], lc // (LOCAL[0] * LOCAL[1])
        gcell_new = (lx + n[0] * ((ly + ny * (world - 1)) + n[1] * lz)).astype(np.int32)
        gcell_old = (cx + n[0] * ((cy - ny) + n[1] * cz)).astype(np.int32)[keep]
        sp["pos"] = np.ascontiguousarray(np.concatenate([sp["pos"][:, keep], new["pos"]], axis=1))
        sp["mom"] = np.ascontiguousarray(np.concatenate([sp["mom"][:, keep], new["mom"]], axis=1))
        sp["w"] = np.ascontiguousarray(np.concatenate([sp["w"][keep], new["w"]]))
        sp["cell"] = np.ascontiguousarray(np.concatenate([gcell_old, gcell_new]))
    for _ in range(steps):
        o.step_open(E, B, J, sps)
    r = {int(d["pos"]): d for d in (np.load(os.path.join(str(tmp_path), "mw%d.npz" % k)) for k in range(world))}
    Eg = np.concatenate([r[0]["E"], r[1]["E"]], axis=2)
    Bg = np.concatenate([r[0]["B"], r[1]["B"]], axis=2)
    _, escale = util.khi_scales(p, 1)
    assert np.abs(Eg - o.interior(E)).max() / escale < 2e-5
    assert np.abs(Bg - o.interior(B)).max() / escale < 2e-5
    assert int(r[0]["ne"]) + int(r[1]["ne"]) == sps[0]["w"].shape[0]
    a, b = np.sort(np.concatenate([r[0]["eux"], r[1]["eux"]])), np.sort(sps[0]["mom"][0])
    assert np.abs(a - b).max() / np.abs(b).max() < 5e-6
