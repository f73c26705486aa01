#!/usr/bin/env python
"""bench.py — headline benchmark of the PIC step (BASELINE.json: macro-particle updates/s and ms/PIC step).

  python bench.py --gpus N --steps K --warmup W            our arm (libpicstep.so, device resident + e2e)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (restated oracle, OpenMP)

Workload (config.workload): KelvinHelmholtz 3D, 256^3 cells per GPU, 25 electrons + 25 ions per cell, TSC / Boris /
Esirkepov / Yee, periodic, weak-scaled in y (`-d 1 N 1`) like share/picongpu/examples/KelvinHelmholtz/etc/picongpu/
8_bench.cfg.  One "step" is one full PIC step (current reset, push + re-sort/migration of both species, field
update, deposition, J reduction, field update).  One macro-particle update = gather + push + move + deposit of one
macro particle for one step.  Inputs are synthetic (device-side KHI generator, Philox seed 42).

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from picongpu_b200 import param as prm  # noqa: E402

# algorithmic bytes
BYTES_PER_UPDATE = 54.0  # read 30 B + write 24 B per macro-particle update (fused gather+push+move+deposit)
BYTES_PER_CELL = 168.0  # E,B gather 24 + J 36 + 2x B-half 72 + E update 36
# per launch of the two particle kernels
PUSH_BYTES_PER_PARTICLE = 30.0 + 24.0 + 4.0  # read pos,mom,w,cell; write pos,mom; write re-sort key
DEPOSIT_BYTES_PER_PARTICLE = 30.0  # read pos,mom,w,cell
DEPOSIT_BYTES_PER_CELL = 24.0  # J read-modify-write, 3 components
FUSED_BYTES_PER_CELL = 24.0 + 24.0  # fused run kernel: E,B tile read + J read-modify-write


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


VARIANT = {"shape": "TSC", "pusher": "Boris", "current": "Esirkepov", "solver": "Yee", "interp": "none", "config": "khi"}

# BASELINE.json configurations.  khi = C2 (the headline, C1 at 64^3); the others are the same kernels with other
# template arguments / boundaries and are benchmark lines kept under profiles/, not the driver's default:
#   thermal   C5  share/picongpu/benchmarks/Thermal: electrons only, uniform, T = 17.5 * 511 keV (relativistic: c dt / dx = 0.5,
#                 every particle crosses cells all the time), Boris as the north star says (the benchmark default is HigueraCary)
#   lwfa_like C3  LaserWakefield's kernels and boundaries (CIC, open + absorbing y, cell sizes with c dt / dy = 0.94) on the KHI
#                 plasma -- WITHOUT the laser (incident-field source not built): it measures the step, not wakefield physics
#   foil_like C4  PQS + Lehe + Binomial current smoothing (the "high-order deposition stress test") on the KHI plasma
LWFA_MOVE_POINT = 0.14
CONFIGS = {
    "khi": {},
    "thermal": dict(shape="TSC", pusher="Boris", current="Esirkepov", solver="Yee", interp="none"),
    "lwfa_like": dict(shape="CIC", pusher="Boris", current="Esirkepov", solver="Yee", interp="none"),
    "foil_like": dict(shape="PQS", pusher="Boris", current="Esirkepov", solver="Lehe", interp="binomial"),
    # C3 with its laser and its moving window: examples/LaserWakefield's grid (2048 cells along y at N = 8), boundaries
    # (nothing periodic, the default PML of 12 cells on every outer face), CIC / Boris / Esirkepov / Yee and the example's
    # GaussianPulse (a0 = 8, 0.8 um, 5 fs, W0 = 4.25 um, circular, focus 46.2 um inside, PULSE_INIT 15) through the YMin
    # Huygens surface; cold electron + ion plasma; the window starts to slide once the pulse is in completely
    # (--windowMovePoint 0.14 instead of the example's 0.9, so that 900 steps hold two slides: tools/campaign.sh) and the
    # slab that enters is re-initialised; N >= 2 GPUs
    "lwfa": dict(shape="CIC", pusher="Boris", current="Esirkepov", solver="Yee", interp="none"),
    # the slow-path cliff: relativistic electrons (the Thermal plasma) on LaserWakefield's cells, c dt / dy = 0.94 -- most
    # trajectories are too wide for the four-node window of the fused kernel and are deposited with global atomics
    "lwfa_hot": dict(shape="CIC", pusher="Boris", current="Esirkepov", solver="Yee", interp="none"),
}


def n_species():
    return 1 if VARIANT["config"] in ("thermal", "lwfa_hot") else 2


def workload_name(grid, ppc, n):
    v = VARIANT
    extra = "" if v["interp"] == "none" else "_Binomial"
    if v["config"] in ("thermal", "lwfa_hot"):
        return "%s_%dx%dx%d_per_gpu_%dppc_electrons_T17.5mc2_%s_%s_%s_%s%s_periodic_d1x%dx1" % (
            "Thermal3D" if v["config"] == "thermal" else "ThermalOnLaserWakefieldCells3D_cdt_over_dy_0.94",
            grid[0], grid[1], grid[2], ppc, v["shape"], v["pusher"], v["current"], v["solver"], extra, n)
    head = {"khi": "KelvinHelmholtz3D", "lwfa_like": "LaserWakefieldLike3D_noLaser_KHIplasma_open_y", "foil_like": "FoilLCTLike3D_KHIplasma",
            "lwfa": "LaserWakefield3D_a0_8_GaussianPulse_PML_movingWindow_coldPlasma"}[v["config"]]
    return "%s_%dx%dx%d_per_gpu_%d+%dppc_%s_%s_%s_%s%s_%s_d1x%dx1" % (
        head, grid[0], grid[1], grid[2], ppc, ppc, v["shape"], v["pusher"], v["current"], v["solver"], extra,
        {"lwfa_like": "absorbing_y", "lwfa": "absorbing_xyz"}.get(v["config"], "periodic"), n)


def variant_kwargs():
    v = VARIANT
    kw = dict(shape=prm.SHAPE_NAMES[v["shape"]], pusher=prm.PUSHER_NAMES[v["pusher"]], current_solver=prm.CURRENT_NAMES[v["current"]],
              field_solver=prm.SOLVER_NAMES[v["solver"]], current_interpolation=1 if v["interp"] == "binomial" else 0)
    if v["config"] in ("lwfa_like", "lwfa"):
        # share/picongpu/examples/LaserWakefield/include/picongpu/param/simulation.param: dt = 1.39e-16 s, cells 0.1772 um x
        # 0.4430e-7 m x 0.1772 um (c dt / dy = 0.94); --periodic 1 0 1, exponential absorber on the open axis
        kw.update(periodic=(1, 0, 1), absorber_kind=1, delta_t_si=1.39e-16, cell_si=(0.1772e-6, 0.4430e-7, 0.1772e-6))
    if v["config"] == "lwfa":
        # etc/picongpu/8.cfg: no --periodic (all axes open), -m; fieldAbsorber default = pml, NUM_CELLS 12 everywhere
        kw.update(moving_window=1, periodic=(0, 0, 0), absorber_kind=2)
    if v["config"] == "lwfa_hot":
        kw.update(delta_t_si=1.39e-16, cell_si=(0.1772e-6, 0.4430e-7, 0.1772e-6), base_density_si=1.0e25)
    return kw


def make_params(grid, **kw):
    if VARIANT["config"] in ("thermal", "lwfa_hot"):
        return prm.thermal_params(grid=grid, **kw)
    p = prm.khi_params(grid=grid, **kw)
    if VARIANT["config"] == "lwfa":
        # examples/LaserWakefield/include/picongpu/param/incidentField.param (the defaults of gaussian_pulse_laser) and
        # param/fieldAbsorber.param (the defaults of pml_params)
        p.laser = prm.gaussian_pulse_laser(p)
        p.pml = prm.pml_params(p)
    return p


# -------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated (oracle/picoracle.cpp), OpenMP over all host cores
# -------------------------------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may use (cgroup / affinity aware), never the OMP_NUM_THREADS=1 torchrun exports."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(steps, warmup, grid=(64, 64, 64)):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from oracle import picoracle

    # torch.distributed.run exports OMP_NUM_THREADS=1 to every worker: set the team size explicitly
    picoracle.lib().orc_set_num_threads(host_cores())
    p = prm.khi_params(grid=grid, **variant_kwargs())
    o, e, i = util.khi_ic(picoracle, p)
    E, B, J = o.field(), o.field(), o.field()
    npart = 2 * e["w"].shape[0]
    for _ in range(warmup):
        o.step(E, B, J, [e, i])
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(E, B, J, [e, i])
    dt = (time.perf_counter() - t0) / steps
    return npart / dt, dt * 1e3, o.L.orc_num_threads(), npart


def run_reference(args):
    """The reference's CPU path of the same step (OpenMP restatement, all host cores).  The metric is a RATE of
    macro-particle updates, so the bounded 64^3 sample is comparable with the 256^3-per-GPU GPU arm; the host of an
    N-GPU box is ONE CPU arm whatever N is, so the line is n_gpus independent and says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 3))
    rate, ms, cores, npart = cpu_oracle_rate(steps, warm)
    sample = "KelvinHelmholtz 64x64x64, 25+25 ppc (%d macro particles), %d timed steps of the restated reference step on %d host threads" % (npart, steps, cores)
    line = {
        "impl": "reference",
        "metric": "macro_particle_updates_per_s",
        "value": rate,
        "unit": "updates/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warm,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.grid, args.ppc, args.gpus), "sample": sample,
                   "n_gpus_independent": "the whole host (all %d cores) is one CPU arm for every N; the value does not grow with --gpus" % cores},
        "cpu_baseline": {"value": rate, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "reference binary not buildable here (needs Boost+MPI); OpenMP restatement of the same arithmetic"},
        "e2e": {"value": rate, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# -------------------------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------------------------
def multi_gpu_parity_check(world, rank, local, dev):
    """N-rank CUDA step against the single-domain oracle, before the timed region (the oracle as CHECKER only):
    KHI 16 x 32*N x 8 global, slabs in y, 6 steps, particles kicked in y/z so they cross the rank boundaries
    (pack / NCCL send-recv / append kernels, E/B/J guard exchange).  Returns the dict stored in checks."""
    import torch
    import torch.distributed as dist

    from picongpu_b200 import picstep

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from oracle import picoracle
    from test_multirank_cpu import _kick

    steps, local_grid = 6, (16, 32, 8)  # four supercell layers along y per rank: two BORDER layers and a CORE area
    picoracle.lib().orc_set_num_threads(2 if world > 1 else host_cores())
    p = prm.khi_params(grid=local_grid, devices=(1, world, 1), rank_pos=(0, rank, 0))
    o, e, i = util.khi_ic(picoracle, p)
    _kick(p, e)
    _kick(p, i)
    sim = picstep.Simulation(p, device=local, exact=False)
    if world > 1:
        uid = sim.comm_unique_id() if rank == 0 else bytes(128)
        t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        sim.comm_init(bytes(t.cpu().tolist()), rank, world)
    for name, sp in (("e", e), ("i", i)):
        sim.upload_particles(name, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    n0 = [sim.particle_count("e"), sim.particle_count("i")]
    sim.step(steps)
    sim.sync()
    n1 = [sim.particle_count("e"), sim.particle_count("i")]
    mine = torch.from_numpy(np.stack([o.interior(sim.download_field(picstep.FIELD_E)), o.interior(sim.download_field(picstep.FIELD_B))])).to(dev)
    cnt = torch.tensor(n0 + n1, device=dev, dtype=torch.int64)
    sim.close()
    if world > 1:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        cnts = [torch.empty_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
    else:
        parts, cnts = [mine], [cnt]
    if rank != 0:
        return None
    G = np.concatenate([t.cpu().numpy() for t in parts], axis=3)  # [E|B][3][z][y*N][x]
    cn = np.stack([t.cpu().numpy() for t in cnts])
    pg = prm.khi_params(grid=(local_grid[0], local_grid[1] * world, local_grid[2]))
    og, eg, ig = util.khi_ic(picoracle, pg)
    _kick(pg, eg)
    _kick(pg, ig)
    E, B, J = og.field(), og.field(), og.field()
    for _ in range(steps):
        og.step(E, B, J, [eg, ig])
    _, escale = util.khi_scales(pg, 1)
    # per-rank particle counts of the oracle's final state
    ny = local_grid[1]
    ref_cnt = []
    for sp in (eg, ig):
        cy = (sp["cell"] // pg.grid[0]) % pg.grid[1]
        ref_cnt.append(np.bincount(cy // ny, minlength=world))
    counts_equal = bool(np.array_equal(cn[:, 2], ref_cnt[0]) and np.array_equal(cn[:, 3], ref_cnt[1]))
    migrated = int(np.abs(cn[:, 2] - cn[:, 0]).sum() + np.abs(cn[:, 3] - cn[:, 1]).sum())
    return {"workload": "KHI 16x%dx8 global on %d rank(s), %d steps, kicked particles, production build vs single-domain oracle" % (ny * world, world, steps),
            "dE": float(np.abs(G[0] - og.interior(E)).max() / escale), "dB": float(np.abs(G[1] - og.interior(B)).max() / escale),
            "scale": "per-species drive field of one step (util.khi_scales)", "counts_equal": counts_equal,
            "count_mismatch": int(np.abs(cn[:, 2] - ref_cnt[0]).sum() + np.abs(cn[:, 3] - ref_cnt[1]).sum()),
            "migrated": migrated, "particles": int(cn[:, 2:].sum())}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from picongpu_b200 import picstep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=dev)
        try:
            import nvidia.nccl as _n

            os.environ.setdefault("PICSTEP_NCCL_LIB", os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2"))
        except Exception:
            pass

    parity = None
    if not args.no_parity:
        try:
            parity = multi_gpu_parity_check(world, rank, local, dev)
        except Exception as ex:  # pragma: no cover
            parity = {"error": str(ex)[:300]}
    grid = tuple(args.grid)
    if args.scaling == "strong":
        if grid[1] % (8 * world) or grid[1] // world < 16:
            raise SystemExit("strong scaling: the global y extent must split into >= 2 supercells per GPU")
        grid = (grid[0], grid[1] // world, grid[2])
    p = make_params(grid, devices=(1, world, 1), rank_pos=(0, rank, 0), **variant_kwargs())
    sim = picstep.Simulation(p, device=local, exact=False)
    if world > 1:
        if rank == 0:
            uid = sim.comm_unique_id()
        else:
            uid = bytes(128)
        t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        sim.comm_init(bytes(t.cpu().tolist()), rank, world)
    ppc_dim = {25: (5, 5, 1), 16: (4, 4, 1), 9: (3, 3, 1), 8: (2, 2, 2), 4: (2, 2, 1), 3: (3, 1, 1), 2: (2, 1, 1), 1: (1, 1, 1)}[args.ppc]
    names = [sp.name for sp in p.species]

    def init(sm):
        if VARIANT["config"] in ("thermal", "lwfa_hot"):
            sm.init_thermal("e", args.ppc)
        elif VARIANT["config"] == "lwfa":
            sm.init_thermal("e", args.ppc, temperature_keV=1.0e-3, seed=42 + 2 * sm.slides)
            sm.init_thermal("i", args.ppc, temperature_keV=1.0e-3, seed=43 + 2 * sm.slides)
        else:
            sm.init_khi(ppc_dim=ppc_dim)

    init(sim)
    ncell = grid[0] * grid[1] * grid[2]
    npart = sum(sim.particle_count(nm) for nm in names)
    stream = torch.cuda.ExternalStream(sim.stream(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        sim.sync()

    # ---- device resident: W warm-up + K timed steps ------------------------------------------------------------
    sim.step(args.warmup)
    barrier()
    sim.slow_path_counts()  # reset
    sim.stage_times(True)  # reset + enable asynchronous per-stage events
    l0 = sim.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    slides = 0
    if VARIANT["config"] == "lwfa":
        # Simulation::runOneStep + movingWindowCheck (Simulation.hpp:522-593): step by step, slide when the window passes a
        # local-domain border; the rank that becomes the top of the window starts empty and gets fresh plasma
        for k in range(args.steps):
            step = sim.step_index
            do_slide, _ = picstep.moving_window_info(p.global_grid[1], p.grid[1], p.cell_size[1], p.c * p.dt, LWFA_MOVE_POINT, step)
            sim.step(1)
            if do_slide:
                slides += 1
                if sim.slide():
                    init(sim)
    else:
        sim.step(args.steps)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    sampler.stop_flag = True
    sampler.join()
    ov = sim.overlap_times() if world > 1 else None
    stage = sim.stage_times(False)
    launches = sim.launch_count() - l0
    if world > 1:
        tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        cnt = torch.tensor([float(sum(sim.particle_count(nm) for nm in names))], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        npart_total = float(cnt.item())
    else:
        npart_total = float(sum(sim.particle_count(nm) for nm in names))
    ms_step = ms_total / args.steps
    value = npart_total / (ms_step * 1e-3)

    wide, plane = sim.slow_path_counts()
    slow_path = {"wide_trajectories_per_update": wide / (npart * args.steps), "pqs_plane_trajectories_per_update": plane / (npart * args.steps),
                 "note": "fraction of macro-particle updates of this rank whose current went through the global-atomic path of the fused kernel"}
    # size independent properties at the benchmark size: particle conservation (periodic), Gauss residual at round-off
    checks = {"particles_conserved": bool(abs(npart_total - npart * world) < 0.5)}
    if VARIANT["config"] == "lwfa":
        checks["window_slides_in_timed_region"] = slides
        checks["laser_field_energy"] = float(sim.field_energy().sum())
    if VARIANT["config"] in ("lwfa_like", "lwfa"):
        checks["particles_conserved"] = None  # open y faces absorb particles
        checks["particles_left"] = npart_total / (npart * world)
    checks["multi_gpu_vs_oracle"] = parity
    try:
        gr = sim.gauss_residual()
        checks["gauss_residual_over_cell_charge"] = gr / (float(args.ppc) * abs(p.base_charge) * p.real_particles_per_cell / args.ppc)
        if n_species() == 1 or VARIANT["config"] == "lwfa":
            # electrons without a neutralising species / electrons and ions seeded independently: rho != eps0 div E = 0
            # from the start, by construction (the reference has no Poisson solve at start-up either)
            checks["gauss_residual_over_cell_charge"] = None
    except Exception as ex:  # pragma: no cover
        checks["gauss_error"] = str(ex)

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------
    peak, peak_kind = measured_peaks()
    nspec = n_species()
    fused = stage["deposit"] == 0.0  # picstep_step fast path: gather+push+move+deposit in one kernel (runKernel)
    if fused:
        per_launch = {
            "run": (stage["push"] / (args.steps * nspec), (npart / nspec) * BYTES_PER_UPDATE + ncell * FUSED_BYTES_PER_CELL, "runKernel<%s,%s,fused>" % (VARIANT["shape"], VARIANT["pusher"])),
        }
    else:
        per_launch = {
            "deposit": (stage["deposit"] / (args.steps * nspec), (npart / nspec) * DEPOSIT_BYTES_PER_PARTICLE + ncell * DEPOSIT_BYTES_PER_CELL, "depositKernel<%s,%s>" % (VARIANT["shape"], VARIANT["current"])),
            "push": (stage["push"] / (args.steps * nspec), (npart / nspec) * PUSH_BYTES_PER_PARTICLE + ncell * 24.0, "pushKernel<%s,%s>" % (VARIANT["shape"], VARIANT["pusher"])),
        }
    dom = max(per_launch, key=lambda k: per_launch[k][0])
    ms_k, bytes_k, kname = per_launch[dom]
    achieved = bytes_k / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
    traffic = None
    smem_wavefronts = inst_executed = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel at the default workload, from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if kname in tj and grid == (256, 256, 256) and args.ppc == 25:
            traffic = float(tj[kname]["dram_bytes_per_launch"])
            smem_wavefronts = tj[kname].get("smem_wavefronts_per_launch")
            inst_executed = tj[kname].get("inst_executed_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_kind + " copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
                "ms_per_launch": ms_k, "algorithmic_bytes_per_launch": bytes_k,
                "share_of_step": ms_k * nspec / ms_step if ms_step > 0 else None}
    # The limiter of the dominant kernel is not HBM but the SM's shared-memory / LSU data pipe (one 128-byte wavefront per
    # cycle and SM) together with the issue slots (4 warp instructions per cycle and SM): both reported next to the HBM
    # fraction.  Wavefronts and instructions per launch come from the committed ncu capture of the same workload
    # (profiles/traffic.json), the duration and the SM clock from this run.
    clk = sampler.summary()
    limiter = None
    if smem_wavefronts and inst_executed and clk.get("sm_mhz") and ms_k > 0:
        cyc = ms_k * 1e-3 * clk["sm_mhz"] * 1e6 * 148
        limiter = {"bound": "smem_pipe", "kernel": kname, "achieved": smem_wavefronts / cyc, "peak": 1.0, "unit": "wavefronts/cycle/SM",
                   "frac": smem_wavefronts / cyc, "wavefronts_per_update": smem_wavefronts / (npart / nspec),
                   "issue": {"achieved": inst_executed / cyc, "peak": 4.0, "unit": "warp instructions/cycle/SM", "frac": inst_executed / cyc / 4.0,
                             "warp_instructions_per_update": inst_executed / (npart / nspec)},
                   "source": "wavefronts / instructions: ncu --set full (profiles/traffic.json); cycles: this run's CUDA-event duration x sampled SM clock x 148 SMs"}
    step_bytes = npart * BYTES_PER_UPDATE + ncell * BYTES_PER_CELL
    step_roofline = {"bound": "hbm", "achieved": step_bytes / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step_per_gpu": step_bytes,
                     "roofline_ms_per_step": step_bytes / (peak * 1e9) * 1e3}

    # ---- end to end: host buffers in, host buffers out, every step ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            # the whole state of every rank has to sit in pinned host memory: shrink the e2e domain in y until all
            # ranks of this node fit into the host RAM (8 x 27 GB does not fit a 196 GB host)
            import psutil

            avail = torch.tensor([float(psutil.virtual_memory().available)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(avail, op=dist.ReduceOp.MIN)
            budget = 0.6 * float(avail.item()) / world
            egrid = list(grid)
            state_bytes = lambda g: g[0] * g[1] * g[2] * (2 * args.ppc * 32 + 24) * 1.02
            while state_bytes(egrid) > budget and egrid[1] % 32 == 0 and egrid[1] > 32:
                egrid[1] //= 2
            if tuple(egrid) != grid:
                sim.close()
                p = make_params(tuple(egrid), devices=(1, world, 1), rank_pos=(0, rank, 0), **variant_kwargs())
                sim = picstep.Simulation(p, device=local, exact=False)
                if world > 1:
                    uid = sim.comm_unique_id() if rank == 0 else bytes(128)
                    t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
                    dist.broadcast(t, 0)
                    sim.comm_init(bytes(t.cpu().tolist()), rank, world)
                init(sim)
                sim.step(3)
                sim.sync()
                stream = torch.cuda.ExternalStream(sim.stream(), device=dev)
            e2e = run_e2e(sim, p, args, stream, world, local, dev)
            e2e["workload"] = workload_name(tuple(egrid), args.ppc, world)
            if tuple(egrid) != grid:
                e2e["note"] = "domain halved in y until the pinned host copies of all ranks fit the host RAM"
        except Exception as ex:  # pragma: no cover
            e2e = {"value": None, "unit": "updates/s", "error": str(ex)[:200]}
    sim.close()

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, cms, cores, cn = cpu_oracle_rate(steps=3, warmup=1)
        cpu = {"value": rate, "unit": "updates/s", "cores": cores, "kind": "port",
               "sample": "KelvinHelmholtz 64x64x64, 25+25 ppc (%d macro particles), 3 timed steps, %.0f ms/step" % (cn, cms)}

    if rank == 0:
        line = {
            "metric": "macro_particle_updates_per_s",
            "value": value,
            "unit": "updates/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(grid, args.ppc, world), "macro_particles_per_gpu": npart,
                       "cells_per_gpu": ncell, "l2_policy": "inputs (%.1f GB of particle data per GPU) are far larger than the 126 MB L2" % (npart * 30 / 1e9),
                       "parallelism": "domain decomposition 1x%dx1" % world},
            "updates_per_s_per_gpu": value / world,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "stage_note": "spans per stage on the stream the stage runs on; at N=1 the re-sort (migrate) runs on a second stream next to the following push / field update, so its span overlaps theirs and the spans do not add up to ms_per_step",
            "roofline": roofline,
            "step_roofline": step_roofline,
            "limiter_roofline": limiter,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "slow_path": slow_path,
            "overlap": None if ov is None else {"exchange_ms_after_border": ov[0], "core_ms_after_border": ov[1], "steps": ov[2], "exchange_hidden": bool(ov[0] < ov[1]),
                                                "note": "rank 0, device time per step from 'BORDER area pushed' until the migration + J guard exchange is complete (second stream) / until the last CORE kernel is complete (compute stream)"},
            "clocks": sampler.summary(),
            "checks": checks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


def run_e2e(sim, p, args, stream, world, local, dev):
    """Same step through picstep_step_host: pinned host state -> H2D -> step -> D2H of E,B + energies, every step."""
    import torch
    import torch.distributed as dist

    from picongpu_b200 import picstep

    sp = []
    for name in [sp.name for sp in p.species]:
        n = sim.particle_count(name)
        arrs = [torch.empty((3, n), dtype=torch.float32).pin_memory(), torch.empty((3, n), dtype=torch.float32).pin_memory(),
                torch.empty((n,), dtype=torch.float32).pin_memory(), torch.empty((n,), dtype=torch.int32).pin_memory()]
        sim.download_particles(name, out=tuple(t.numpy() for t in arrs))  # straight into the pinned buffers
        sp.append(arrs)
    E = torch.from_numpy(sim.download_field(picstep.FIELD_E)).pin_memory()
    B = torch.from_numpy(sim.download_field(picstep.FIELD_B)).pin_memory()
    sp_np = [tuple(t.numpy() for t in arrs) for arrs in sp]
    h2d = sum(t.numel() * t.element_size() for arrs in sp for t in arrs) + E.numel() * 4 + B.numel() * 4
    d2h = E.numel() * 4 + B.numel() * 4 + 4 * 8
    steps = max(1, min(args.steps, args.e2e_steps))
    npart = sum(a[2].shape[0] for a in sp_np)
    sim.step_host(E.numpy(), B.numpy(), sp_np)  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step_host(E.numpy(), B.numpy(), sp_np)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if world > 1:
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    return {"value": npart * world / dt, "unit": "updates/s", "ms_per_step": dt * 1e3, "steps": steps,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "api": "picstep_step_host (pinned host buffers -> H2D -> one PIC step -> D2H of E,B and energies)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, nargs=3, default=[256, 256, 256], help="cells per GPU")
    ap.add_argument("--ppc", type=int, default=25, help="macro particles per cell and species")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank CUDA-vs-oracle check before the timed region")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --grid is the grid per GPU (default, what the driver runs); strong: --grid is the GLOBAL grid, split in y over the GPUs")
    # other BASELINE.json configurations are the same kernels with other template arguments (defaults = the headline)
    ap.add_argument("--shape", default=None, choices=sorted(prm.SHAPE_NAMES))
    ap.add_argument("--pusher", default=None, choices=sorted(prm.PUSHER_NAMES))
    ap.add_argument("--current", default=None, choices=["Esirkepov", "EmZ"])
    ap.add_argument("--solver", default=None, choices=sorted(prm.SOLVER_NAMES))
    ap.add_argument("--interp", default=None, choices=["none", "binomial"])
    ap.add_argument("--config", default="khi", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: the KHI headline)")
    args = ap.parse_args()
    VARIANT.update(config=args.config)
    VARIANT.update(CONFIGS[args.config])  # the configuration's kernels ...
    VARIANT.update({k: v for k, v in dict(shape=args.shape, pusher=args.pusher, current=args.current, solver=args.solver, interp=args.interp).items() if v is not None})  # ... unless given explicitly
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.config != "khi":  # the CPU arm and the N-rank oracle check are the KHI headline's
        args.no_cpu = True
        args.no_parity = True
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
