"""CPU tests of the N>1 path: the domain-decomposition scheme the CUDA exchange implements (axis-sequenced guard
passes with the margins of picstep_exchange_widths, neighbours from picstep_neighbor_ranks, particle migration
records carrying the receiver's cell coordinates) is run on the oracle with two `gloo` ranks and must reproduce
the single-domain oracle run on the same global grid.  (N-GPU result == 1-GPU result.)"""
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

FE, FB, FJ = picstep.FIELD_E, picstep.FIELD_B, picstep.FIELD_J


@pytest.fixture(scope="module", autouse=True)
def built():
    from picongpu_b200 import build

    build.build_all()


def test_axis_passes_equal_full_wrap(orc):
    """x, y, z passes over the full padded extent == the 26-direction periodic wrap (copy and add variants)."""
    p = util.make_params((16, 24, 8))
    o = orc.Oracle(p)
    rng = np.random.RandomState(0)
    F = rng.normal(size=o.field().shape).astype(np.float32)
    A, Bf = F.copy(), F.copy()
    o.guard_copy(A)
    g = p.guard_cells
    for a in range(3):
        o.halo_axis(Bf, a, g[a], g[a], add=False)
    assert np.array_equal(A, Bf)
    # narrow widths: identical inside the exchanged margins
    C1 = F.copy()
    w = [picstep.exchange_widths(p.shape, p.field_solver, p.lehe_dir, FE, a) for a in range(3)]
    for a in range(3):
        o.halo_axis(C1, a, w[a][0], w[a][1], add=False)
    n = p.grid
    sl = tuple(slice(g[a] - w[a][0], g[a] + n[a] + w[a][1]) for a in (2, 1, 0))
    assert np.array_equal(C1[(slice(None),) + sl], A[(slice(None),) + sl])
    # J reduction: deposit-like data confined to the current margins
    wj = [picstep.exchange_widths(p.shape, p.field_solver, p.lehe_dir, FJ, a) for a in range(3)]
    J = np.zeros_like(F)
    slj = tuple(slice(g[a] - wj[a][0], g[a] + n[a] + wj[a][1]) for a in (2, 1, 0))
    J[(slice(None),) + slj] = F[(slice(None),) + slj]
    J1, J2 = J.copy(), J.copy()
    o.guard_add(J1)
    for a in range(3):
        o.halo_axis(J2, a, wj[a][0], wj[a][1], add=True)
    assert np.allclose(o.interior(J1), o.interior(J2), rtol=1e-6, atol=1e-6)


# ---- two-rank run -------------------------------------------------------------------------------------------------
def _exchange_planes(F, axis_np, send_lo, send_hi, recv_lo, recv_hi, lo_rank, hi_rank, add):
    """send_*/recv_* are slices along the split axis (numpy axis `axis_np` of F[c,z,y,x])."""

    def take(sl):
        idx = [slice(None)] * 4
        idx[axis_np] = sl
        return torch.from_numpy(np.ascontiguousarray(F[tuple(idx)]))

    def put(sl, t):
        idx = [slice(None)] * 4
        idx[axis_np] = sl
        if add:
            F[tuple(idx)] += t.numpy()
        else:
            F[tuple(idx)] = t.numpy()

    bl, bh = take(recv_lo).clone(), take(recv_hi).clone()
    sl_, sh_ = take(send_lo), take(send_hi)
    # same posting order as comm.cu: sends (lower, upper), receives (upper, lower); a missing neighbour (-1, open
    # boundary) is skipped on both sides
    ops = []
    if lo_rank >= 0:
        ops.append(dist.P2POp(dist.isend, sl_, lo_rank))
    if hi_rank >= 0:
        ops.append(dist.P2POp(dist.isend, sh_, hi_rank))
        ops.append(dist.P2POp(dist.irecv, bh, hi_rank))
    if lo_rank >= 0:
        ops.append(dist.P2POp(dist.irecv, bl, lo_rank))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    if lo_rank >= 0:
        put(recv_lo, bl)
    if hi_rank >= 0:
        put(recv_hi, bh)


def _field_exchange(o, p, F, field, lo_rank, hi_rank, mode=None, width=None):
    add = (field == FJ) if mode is None else (mode == "add")
    g, n = p.guard_cells, p.grid
    for a in range(3):
        lo, up = picstep.exchange_widths(p.shape, p.field_solver, p.lehe_dir, field, a)
        if width is not None:
            lo = up = width
        if not p.periodic[a] and p.devices[a] == 1:
            continue  # open axis inside one rank: no exchange
        if p.wrap[a]:
            o.halo_axis(F, a, lo, up, add=add)
        else:
            ax = 3 - a  # numpy axis of F[c,z,y,x]
            G, N = g[a], n[a]
            if not add:
                _exchange_planes(F, ax, slice(G, G + up), slice(G + N - lo, G + N), slice(G - lo, G), slice(G + N, G + N + up), lo_rank, hi_rank, False)
            else:
                _exchange_planes(F, ax, slice(G - lo, G), slice(G + N, G + N + up), slice(G, G + up), slice(G + N - lo, G + N), lo_rank, hi_rank, True)


def _migrate(p, sp, cell3, lo_rank, hi_rank):
    """Particles whose new cell left the slab travel with the receiver's cell coordinate (ParticlesBase.kernel:707-938)."""
    n = p.grid
    a = 1
    y = cell3[a]
    out_lo, out_hi = y < 0, y >= n[a]
    stay = ~(out_lo | out_hi)

    def pack(m, shift):
        c3 = cell3[:, m].copy()
        c3[a] += shift
        for d in (0, 2):
            c3[d] %= n[d]
        cell = (c3[0] + n[0] * (c3[1] + n[1] * c3[2])).astype(np.float32)  # exact for small grids
        return torch.from_numpy(np.concatenate([sp["pos"][:, m], sp["mom"][:, m], sp["w"][None, m], cell[None]]).astype(np.float32))

    s_lo, s_hi = pack(out_lo, n[a]), pack(out_hi, -n[a])
    cnt = torch.tensor([s_lo.shape[1], s_hi.shape[1]])
    rc_hi, rc_lo = torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long)

    def both(send_lo_t, send_hi_t, recv_hi_t, recv_lo_t):
        # particles leaving through an open face (neighbour -1) are absorbed: nothing is sent or received there
        ops = []
        if lo_rank >= 0:
            ops.append(dist.P2POp(dist.isend, send_lo_t, lo_rank))
        if hi_rank >= 0:
            ops.append(dist.P2POp(dist.isend, send_hi_t, hi_rank))
            ops.append(dist.P2POp(dist.irecv, recv_hi_t, hi_rank))
        if lo_rank >= 0:
            ops.append(dist.P2POp(dist.irecv, recv_lo_t, lo_rank))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    both(cnt[0:1].clone(), cnt[1:2].clone(), rc_hi, rc_lo)
    r_hi, r_lo = torch.zeros((8, int(rc_hi)), dtype=torch.float32), torch.zeros((8, int(rc_lo)), dtype=torch.float32)
    both(s_lo.contiguous(), s_hi.contiguous(), r_hi, r_lo)
    rec = np.concatenate([r_lo.numpy(), r_hi.numpy()], axis=1)
    sp["pos"] = np.ascontiguousarray(np.concatenate([sp["pos"][:, stay], rec[0:3]], axis=1))
    sp["mom"] = np.ascontiguousarray(np.concatenate([sp["mom"][:, stay], rec[3:6]], axis=1))
    sp["w"] = np.ascontiguousarray(np.concatenate([sp["w"][stay], rec[6]]))
    sp["cell"] = np.ascontiguousarray(np.concatenate([sp["cell"][stay], rec[7].astype(np.int32)]))
    return int(out_lo.sum() + out_hi.sum())


def _kick(p, sp):
    """Deterministic y/z momentum kick (a function of the GLOBAL cell index only) so that particles actually cross
    the slab boundary within a few steps; identical on every decomposition."""
    n, off, G = p.grid, p.global_offset, p.global_grid
    c = sp["cell"]
    gx, gy, gz = c % n[0] + off[0], (c // n[0]) % n[1] + off[1], c // (n[0] * n[1]) + off[2]
    gid = (gx + G[0] * (gy + G[1] * gz)).astype(np.float64)
    mass = np.float32(p.base_mass) * np.float32(sp["massRatio"]) * sp["w"]
    sp["mom"][1] += (0.5 * np.sin(1.7 * gid)).astype(np.float32) * mass
    sp["mom"][2] += (0.3 * np.cos(0.9 * gid)).astype(np.float32) * mass


def _worker(rank, world, port, steps, outdir, kw=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import picoracle as orc

    orc.lib().orc_set_num_threads(2)
    kw = kw or {}
    p = prm.khi_params(grid=(16, 16, 8), devices=(1, world, 1), rank_pos=(0, rank, 0), **kw)
    assert p.wrap == (1, 0, 1)
    lo_rank, hi_rank = picstep.neighbor_ranks(p.devices, p.periodic, rank, 1)
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    E, B, J = o.field(), o.field(), o.field()
    migrated = 0
    for _ in range(steps):
        J[:] = 0
        for sp in (e, i):
            _, cell3 = o.push(sp["massRatio"], sp["chargeRatio"], E, B, sp["pos"], sp["mom"], sp["w"], sp["cell"], want_cell3=True)
            migrated += _migrate(p, sp, cell3, lo_rank, hi_rank)
        o.update_b_half(E, B)
        _field_exchange(o, p, B, FB, lo_rank, hi_rank)
        o.update_e(E, B)
        for sp in (e, i):
            o.deposit(sp["massRatio"], sp["chargeRatio"], J, sp["pos"], sp["mom"], sp["w"], sp["cell"])
        _field_exchange(o, p, J, FJ, lo_rank, hi_rank)
        if p.current_interpolation == 1:  # FieldJ "receive" exchange: one guard cell := neighbour border
            _field_exchange(o, p, J, FJ, lo_rank, hi_rank, mode="copy", width=1)
        o.add_current(E, J)
        o.absorb(E)
        _field_exchange(o, p, E, FE, lo_rank, hi_rank)
        o.update_b_half(E, B)
        o.absorb(B)
        _field_exchange(o, p, B, FB, lo_rank, hi_rank)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), E=o.interior(E), B=o.interior(B), ne=e["w"].shape[0], ni=i["w"].shape[0],
             migrated=migrated, ew=np.sort(e["mom"][0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_slab_decomposition_equals_single_domain(orc, tmp_path, world):
    """world = 3: lower and upper neighbour of a rank are different ranks (with two periodic ranks they coincide)."""
    steps = 4
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, steps, str(tmp_path)), nprocs=world, join=True)
    # single domain reference on the global grid 16 x (16 * world) x 8
    p = prm.khi_params(grid=(16, 16 * world, 8))
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    E, B, J = o.field(), o.field(), o.field()
    for _ in range(steps):
        o.step(E, B, J, [e, i])
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]
    Eg = np.concatenate([x["E"] for x in r], axis=2)
    Bg = np.concatenate([x["B"] for x in r], axis=2)
    _, escale = util.khi_scales(p, 1)
    assert np.abs(Eg - o.interior(E)).max() / escale < 1e-5
    assert np.abs(Bg - o.interior(B)).max() / escale < 1e-5
    assert sum(int(x["ne"]) for x in r) == e["w"].shape[0]
    assert sum(int(x["ni"]) for x in r) == i["w"].shape[0]
    assert sum(int(x["migrated"]) for x in r) > 100
    # same particles (momentum multiset) on both decompositions; J is summed in a different order per
    # decomposition, so E and with it the momenta agree to fp32 round-off, not bit for bit
    a, b = np.sort(np.concatenate([x["ew"] for x in r])), np.sort(e["mom"][0])
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-6


OPEN = dict(periodic=(1, 0, 1), current_interpolation=1, absorber_kind=1, absorber_cells=((0, 0), (6, 6), (0, 0)),
            absorber_strength=((0, 0), (0.05, 0.05), (0, 0)))


def test_two_rank_open_boundary_equals_single_domain(orc, tmp_path):
    """Same decomposition with an open, absorbing split axis and the Binomial filter: missing neighbours are skipped,
    leavers through the open faces are absorbed, the J guard gets the neighbour's border for the filter."""
    world, steps = 2, 4
    port = 29500 + (os.getpid() % 2000) + 11
    mp.spawn(_worker, args=(world, port, steps, str(tmp_path), OPEN), nprocs=world, join=True)
    p = prm.khi_params(grid=(16, 32, 8), **OPEN)
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    n0 = e["w"].shape[0]
    E, B, J = o.field(), o.field(), o.field()
    sps = [e, i]
    for _ in range(steps):
        o.step_open(E, B, J, sps)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]
    Eg = np.concatenate([r[0]["E"], r[1]["E"]], axis=2)
    Bg = np.concatenate([r[0]["B"], r[1]["B"]], axis=2)
    _, escale = util.khi_scales(p, 1)
    assert np.abs(Eg - o.interior(E)).max() / escale < 1e-5
    assert np.abs(Bg - o.interior(B)).max() / escale < 1e-5
    assert int(r[0]["ne"]) + int(r[1]["ne"]) == sps[0]["w"].shape[0] < n0
    a, b = np.sort(np.concatenate([r[0]["ew"], r[1]["ew"]])), np.sort(sps[0]["mom"][0])
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-6


# ---- moving window: rank rotation (GridController::slide) on the CPU emulation of the exchange scheme -----------------
MW = dict(periodic=(1, 0, 1), moving_window=1, absorber_kind=1, absorber_cells=((0, 0), (6, 6), (0, 0)),
          absorber_strength=((0, 0), (0.05, 0.05), (0, 0)))
LOCAL = (16, 16, 8)


def _new_slab(orc, p_top):
    """Plasma of the slab that enters the window: KHI recipe with another seed, kept two cells away from the slab faces
    (the fresh top rank has no guard data from its lower neighbour before the first exchange)."""
    _, e, i = util.khi_ic(orc, p_top, seed=77)
    out = []
    for sp in (e, i):
        cy = (sp["cell"] // p_top.grid[0]) % p_top.grid[1]
        keep = (cy >= 2) & (cy < p_top.grid[1] - 2)
        out.append(dict(massRatio=sp["massRatio"], chargeRatio=sp["chargeRatio"], pos=np.ascontiguousarray(sp["pos"][:, keep]),
                        mom=np.ascontiguousarray(sp["mom"][:, keep]), w=np.ascontiguousarray(sp["w"][keep]), cell=np.ascontiguousarray(sp["cell"][keep])))
    return out


def _one_step(o, p, E, B, J, species, lo_rank, hi_rank):
    J[:] = 0
    for sp in species:
        _, cell3 = o.push(sp["massRatio"], sp["chargeRatio"], E, B, sp["pos"], sp["mom"], sp["w"], sp["cell"], want_cell3=True)
        _migrate(p, sp, cell3, lo_rank, hi_rank)
    o.update_b_half(E, B)
    _field_exchange(o, p, B, FB, lo_rank, hi_rank)
    o.update_e(E, B)
    for sp in species:
        if sp["w"].shape[0]:
            o.deposit(sp["massRatio"], sp["chargeRatio"], J, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    _field_exchange(o, p, J, FJ, lo_rank, hi_rank)
    o.add_current(E, J)
    o.absorb(E)
    _field_exchange(o, p, E, FE, lo_rank, hi_rank)
    o.update_b_half(E, B)
    o.absorb(B)
    _field_exchange(o, p, B, FB, lo_rank, hi_rank)


def _mw_worker(rank, world, port, steps, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import picoracle as orc

    orc.lib().orc_set_num_threads(2)
    pos, slides = rank, 0
    p = prm.khi_params(grid=LOCAL, devices=(1, world, 1), rank_pos=(0, pos, 0), **MW)
    lo_rank, hi_rank = picstep.window_neighbors(world, 0, pos, slides)
    assert (lo_rank, hi_rank) == picstep.neighbor_ranks(p.devices, p.periodic, rank, 1)
    o, e, i = util.khi_ic(orc, p)
    _kick(p, e)
    _kick(p, i)
    species = [e, i]
    E, B, J = o.field(), o.field(), o.field()
    for _ in range(steps):
        _one_step(o, p, E, B, J, species, lo_rank, hi_rank)
    # slide: every rank moves one position down, the lowest becomes the (empty) top of the window
    slides += 1
    pos = (pos - 1) % world
    p = prm.khi_params(grid=LOCAL, devices=(1, world, 1), rank_pos=(0, pos, 0), **MW)
    o = orc.Oracle(p)  # the open faces moved with the positions
    lo_rank, hi_rank = picstep.window_neighbors(world, 0, pos, slides)
    if pos == world - 1:
        E[:], B[:], J[:] = 0, 0, 0
        species = _new_slab(orc, p)
    for _ in range(steps):
        _one_step(o, p, E, B, J, species, lo_rank, hi_rank)
    np.savez(os.path.join(outdir, "mw%d.npz" % rank), E=o.interior(E), B=o.interior(B), ne=species[0]["w"].shape[0], ew=np.sort(species[0]["mom"][0]), pos=pos)
    dist.barrier()
    dist.destroy_process_group()


def test_window_neighbors_follow_the_slides():
    """After k slides the rank at position q is (q + k) mod n; the window ends have no neighbour."""
    n = 4
    for k in range(9):
        rank_at = [(q + k) % n for q in range(n)]
        for q in range(n):
            lo, hi = picstep.window_neighbors(n, 0, q, k)
            assert lo == (rank_at[q - 1] if q > 0 else -1) and hi == (rank_at[q + 1] if q < n - 1 else -1)
    assert picstep.window_neighbors(3, 1, 0, 0) == (2, 1)


def test_two_rank_moving_window_slide_equals_shifted_single_domain(orc, tmp_path):
    world, steps = 2, 3
    port = 29500 + (os.getpid() % 2000) + 23
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
    for sp, new in zip(sps, _new_slab(orc, p_top)):
        n = p.grid
        cx, cy, cz = sp["cell"] % n[0], (sp["cell"] // n[0]) % n[1], sp["cell"] // (n[0] * n[1])
        keep = cy >= ny
        lc = new["cell"]
        lx, ly, lz = lc % LOCAL[0], (lc // LOCAL[0]) % LOCAL[1], lc // (LOCAL[0] * LOCAL[1])
        gnew = (lx + n[0] * ((ly + ny * (world - 1)) + n[1] * lz)).astype(np.int32)
        gold = (cx + n[0] * ((cy - ny) + n[1] * cz)).astype(np.int32)[keep]
        sp["pos"] = np.ascontiguousarray(np.concatenate([sp["pos"][:, keep], new["pos"]], axis=1))
        sp["mom"] = np.ascontiguousarray(np.concatenate([sp["mom"][:, keep], new["mom"]], axis=1))
        sp["w"] = np.ascontiguousarray(np.concatenate([sp["w"][keep], new["w"]]))
        sp["cell"] = np.ascontiguousarray(np.concatenate([gold, gnew]))
    for _ in range(steps):
        o.step_open(E, B, J, sps)
    r = {int(d["pos"]): d for d in (np.load(os.path.join(str(tmp_path), "mw%d.npz" % k)) for k in range(world))}
    Eg = np.concatenate([r[0]["E"], r[1]["E"]], axis=2)
    Bg = np.concatenate([r[0]["B"], r[1]["B"]], axis=2)
    _, escale = util.khi_scales(p, 1)
    assert np.abs(Eg - o.interior(E)).max() / escale < 1e-5
    assert np.abs(Bg - o.interior(B)).max() / escale < 1e-5
    assert int(r[0]["ne"]) + int(r[1]["ne"]) == sps[0]["w"].shape[0]
    a, b = np.sort(np.concatenate([r[0]["ew"], r[1]["ew"]])), np.sort(sps[0]["mom"][0])
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-6
