import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import util
from picongpu_b200 import picstep, param as prm
FE, FB, FJ = picstep.FIELD_E, picstep.FIELD_B, picstep.FIELD_J

def run(thermal, amp, ppc=6, grid=(16, 16, 8), shape=prm.SHAPE_TSC):
    p = util.make_params(grid, shape=shape)
    E, B = util.smooth_fields(p, seed=5, amp=amp)
    pos, mom, w, cell = util.random_particles(p, ppc=ppc, seed=9, thermal=thermal)
    res = []
    for unfused in (False, True):
        s = picstep.Simulation(p, device=0, exact=True, unfused=unfused)
        s.upload_field(FE, E); s.upload_field(FB, B)
        s.upload_particles("e", pos, mom, w, cell)
        s.current_reset()
        s.step(1)
        res.append(s.download_field(FJ))
        s.close()
    Ja, Jb = res
    g, n = p.guard_cells, p.grid
    Ja, Jb = (x[:, g[2]:g[2] + n[2], g[1]:g[1] + n[1], g[0]:g[0] + n[0]] for x in (Ja, Jb))
    d = np.abs(Ja - Jb)
    print("thermal", thermal, "amp", amp, "max|J|", np.abs(Jb).max(), "maxdiff", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "sumJa", Ja.sum(axis=(1,2,3)), "sumJb", Jb.sum(axis=(1,2,3)))
    nz = np.argwhere(d > 1e-4 * np.abs(Jb).max())
    print("  cells with diff:", len(nz), "of", d.size, "first", nz[:8].tolist())

run(0.01, 0.0)
run(0.1, 0.0)
run(0.4, 0.0)
run(0.4, 0.05)
run(0.1, 0.0, grid=(32, 32, 16))
