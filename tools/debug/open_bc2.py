"""debug: compare J including guards after one step (run on the GPU box)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from oracle import picoracle as orc
from picongpu_b200 import picstep
periodic, interp = (0, 1, 1), 1
p = util.make_params((16, 16, 8), periodic=periodic, current_interpolation=interp, absorber_kind=0)
o, e, i = util.khi_ic(orc, p)
rng = np.random.RandomState(11)
e["mom"] += (rng.normal(size=e["mom"].shape) * 0.3).astype(np.float32) * (np.float32(p.base_mass) * e["w"] * np.float32(p.c))
for unf in (False, True):
    s = picstep.Simulation(p, device=0, exact=True, unfused=unf)
    e2 = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in e.items()}
    i2 = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in i.items()}
    for name, sp in (("e", e2), ("i", i2)):
        s.upload_particles(name, sp["pos"], sp["mom"], sp["w"], sp["cell"])
    E, B, J = o.field(), o.field(), o.field()
    o.step_open(E, B, J, [e2, i2])
    s.step(1); s.sync()
    Jg = s.download_field(2)
    d = np.abs(Jg - J)
    print("unfused", unf, "max diff J full", d.max(), "max J", np.abs(J).max())
    idx = np.argwhere(d > 1e-5)
    print("n cells differing", len(idx))
    if len(idx):
        print("x range", idx[:, 3].min(), idx[:, 3].max(), "y range", idx[:, 2].min(), idx[:, 2].max(), "z range", idx[:, 1].min(), idx[:, 1].max())
        for k in idx[:10]:
            print(tuple(k), Jg[tuple(k)], J[tuple(k)])
    Eg = s.download_field(0)
    d = np.abs(Eg - E); print("E diff", d.max(), np.argwhere(d > 1e-7)[:5])
    s.close()
