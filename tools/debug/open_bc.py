"""debug: where do GPU and oracle differ for open boundaries (run on the GPU box)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from oracle import picoracle as orc
from picongpu_b200 import picstep

for periodic, interp in (((0, 1, 1), 0), ((0, 1, 1), 1), ((1, 0, 1), 1), ((1, 1, 1), 1)):
    p = util.make_params((16, 16, 8), periodic=periodic, current_interpolation=interp, absorber_kind=1,
                         absorber_cells=((6, 6), (5, 7), (3, 3)), absorber_strength=((0.05, 0.05), (0.1, 0.02), (0.2, 0.2)))
    o, e, i = util.khi_ic(orc, p)
    rng = np.random.RandomState(11)
    e["mom"] += (rng.normal(size=e["mom"].shape) * 0.3).astype(np.float32) * (np.float32(p.base_mass) * e["w"] * np.float32(p.c))
    s = picstep.Simulation(p, device=0, exact=True)
    for name, sp in (("e", e), ("i", i)):
        s.upload_particles(name, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    E, B, J = o.field(), o.field(), o.field()
    for st in range(3):
        o.step_open(E, B, J, [e, i])
        s.step(1)
        s.sync()
        out = []
        for nm, f, ref in (("E", 0, E), ("B", 1, B), ("J", 2, J)):
            g = s.download_field(f)
            d = np.abs(o.interior(g) - o.interior(ref))
            k = np.unravel_index(d.argmax(), d.shape)
            out.append("%s %.2e@%s/max %.2e" % (nm, d.max(), k, np.abs(o.interior(ref)).max()))
        print(periodic, interp, "step", st, " | ".join(out), "n", s.particle_count("e"), e["w"].shape[0])
    s.close()
