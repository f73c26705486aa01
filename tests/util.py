"""Shared helpers for the parity tests: seeded synthetic fields / particles and canonical particle ordering."""
import numpy as np

from picongpu_b200 import param as prm


def make_params(grid=(16, 16, 8), **kw):
    return prm.khi_params(grid=grid, **kw)


def smooth_fields(p, seed=0, amp=1.0):
    """Random smooth periodic E,B (a few Fourier modes) on the padded grid, guards filled periodically.
    Magnitudes are O(amp * m_e c / (e dt)) so that the push is strongly field dependent."""
    rng = np.random.RandomState(seed)
    n = p.grid
    g = p.guard_cells
    z, y, x = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    out = []
    for _ in range(2):
        F = np.zeros((3, n[2], n[1], n[0]), np.float64)
        for c in range(3):
            for _m in range(4):
                k = rng.randint(0, 3, 3)
                ph = rng.uniform(0, 2 * np.pi)
                F[c] += rng.normal() * np.sin(2 * np.pi * (k[0] * x / n[0] + k[1] * y / n[1] + k[2] * z / n[2]) + ph)
            F[c] += 0.1 * rng.normal(size=F[c].shape)
        F *= amp
        out.append(pad_periodic(F.astype(np.float32), g))
    return out[0], out[1]


def pad_periodic(F, g):
    return np.ascontiguousarray(np.pad(F, ((0, 0), (g[2], g[2]), (g[1], g[1]), (g[0], g[0])), mode="wrap"))


def random_particles(p, ppc=4, seed=1, thermal=0.3, species_mass_ratio=1.0, tag_weights=True):
    """ppc particles per cell at random in-cell positions with a relativistic-ish momentum spread.
    Weights are made pairwise distinct (tag_weights) so particles can be matched across re-sorts."""
    rng = np.random.RandomState(seed)
    n = p.grid
    ncell = n[0] * n[1] * n[2]
    npart = ncell * ppc
    cell = np.repeat(np.arange(ncell, dtype=np.int32), ppc)
    rng.shuffle(cell)
    pos = rng.uniform(0, 1, (3, npart)).astype(np.float32)
    pos = np.minimum(pos, np.float32(1.0 - 2.0**-24))
    w0 = p.typical_num_particles_per_macro
    if tag_weights:
        w = (w0 * (1.0 + np.arange(npart) * 2.0**-20)).astype(np.float32)
        assert len(np.unique(w)) == npart
    else:
        w = np.full(npart, w0, np.float32)
    mass = np.float32(p.base_mass) * np.float32(species_mass_ratio) * w
    mom = (rng.normal(size=(3, npart)) * thermal).astype(np.float32) * mass * np.float32(p.c)
    return np.ascontiguousarray(pos), np.ascontiguousarray(mom.astype(np.float32)), w, cell


def order_by_weight(pos, mom, w, cell):
    o = np.argsort(w, kind="stable")
    return pos[:, o], mom[:, o], w[o], cell[o]


def canonical_order(pos, mom, w, cell):
    """Lexicographic order on (cell, pos bits, mom bits, w bits): aligns two bit-identical multisets."""
    keys = [w.view(np.uint32)]
    for a in (mom[2], mom[1], mom[0], pos[2], pos[1], pos[0]):
        keys.append(np.ascontiguousarray(a).view(np.uint32))
    keys.append(cell)
    o = np.lexsort(keys)
    return pos[:, o], mom[:, o], w[o], cell[o]


def khi_ic(orc, p, seed=42, ppc_dim=(5, 5, 1)):
    """KelvinHelmholtz initial condition from the oracle's generator: (electrons, ions) as dicts."""
    import ctypes as C

    o = orc.Oracle(p)
    ppc = ppc_dim[0] * ppc_dim[1] * ppc_dim[2]
    ncell = p.grid[0] * p.grid[1] * p.grid[2]
    npart = ncell * ppc
    mk3 = lambda: np.zeros((3, npart), np.float32)
    e = dict(massRatio=1.0, chargeRatio=1.0, pos=mk3(), mom=mk3(), w=np.zeros(npart, np.float32), cell=np.zeros(npart, np.int32))
    i = dict(massRatio=1836.152672, chargeRatio=-1.0, pos=mk3(), mom=mk3(), w=np.zeros(npart, np.float32), cell=np.zeros(npart, np.int32))
    o.L.orc_khi_init(C.byref(o.p), np.array(p.global_grid, np.int32), np.array(p.global_offset, np.int32),
                     np.array(ppc_dim, np.int32), p.real_particles_per_cell, 1836.152672, 1.021, 0.0005, p.ev_pic, seed,
                     e["pos"], e["mom"], e["w"], e["cell"], i["pos"], i["mom"], i["w"], i["cell"])
    return o, e, i


def khi_scales(p, steps=1, ppc=25, beta_gamma=0.2):
    """Error scales for the KelvinHelmholtz start: electron and ion drift currents cancel down to the thermal noise,
    so deviations are measured against the per-species current n*q*v/V (and the E it would drive in `steps` steps),
    not against max|J_net|."""
    V = float(np.prod(np.float64(p.cell_size)))
    jscale = ppc * 1.0 * beta_gamma / V
    escale = p.dt / p.eps0 * jscale * steps
    return jscale, escale


# ---- one-to-one particle matching across a run (tests/test_gpu_c1.py, test_khi_100_steps_vs_oracle) ----------------
def tag_weights(p, sp):
    """w -> the float `tag` ulps above w: tag = ((x0 * 25 + j) * 8 + y0 % 8) * 4 + z0 % 4 with (x0, y0, z0) the start
    cell and j the index inside the cell.  Two particles share a tag only if their start cells differ by a multiple
    of 8 in y or of 4 in z — thermal motion over 100 steps does not bridge that."""
    n = p.grid
    cell = sp["cell"]
    order = np.argsort(cell, kind="stable")
    sorted_cell = cell[order]
    first = np.searchsorted(sorted_cell, sorted_cell, side="left")
    j = np.empty(cell.shape[0], np.int64)
    j[order] = np.arange(cell.shape[0]) - first
    assert j.max() < 25
    x0, y0, z0 = cell % n[0], (cell // n[0]) % n[1], cell // (n[0] * n[1])
    tag = ((x0.astype(np.int64) * 25 + j) * 8 + (y0 % 8)) * 4 + (z0 % 4)
    sp["w"] = (sp["w"].view(np.uint32) + tag.astype(np.uint32)).view(np.float32).copy()


def match_key(p, sp_or_arrays, w0_bits):
    """(tag, block of the START cell): the tag holds y0 % 8 and z0 % 4, and a particle moves far less than half a block
    (4 cells in y, 2 in z) in 100 steps, so the block it started in follows from where it is now."""
    pos, w, cell = sp_or_arrays
    n = p.grid
    tag = (w.view(np.uint32) - np.uint32(w0_bits)).astype(np.int64)
    assert tag.min() >= 0 and tag.max() < 64 * 25 * 8 * 4
    Y = ((cell // n[0]) % n[1]) + pos[1].astype(np.float64)
    Z = (cell // (n[0] * n[1])) + pos[2].astype(np.float64)
    yb = np.rint((Y - ((tag >> 2) & 7) - 0.5) / 8.0).astype(np.int64) % (n[1] // 8)
    zb = np.rint((Z - (tag & 3) - 0.5) / 4.0).astype(np.int64) % (n[2] // 4)
    return (tag << 16) | (yb << 8) | zb


def global_pos(p, pos, cell):
    n = p.grid
    c3 = np.stack([cell % n[0], (cell // n[0]) % n[1], cell // (n[0] * n[1])]).astype(np.float64)
    return c3 + pos.astype(np.float64)


def permuted_copy(species, seed=123):
    """The same particles stored in another order: only the fp32 summation order of the deposition changes."""
    rng = np.random.RandomState(seed)
    out = []
    for sp in species:
        c = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sp.items()}
        perm = rng.permutation(c["w"].shape[0])
        for k in ("pos", "mom"):
            c[k] = np.ascontiguousarray(c[k][:, perm])
        for k in ("w", "cell"):
            c[k] = np.ascontiguousarray(c[k][perm])
        out.append(c)
    return out
