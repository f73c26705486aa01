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
