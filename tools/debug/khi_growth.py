"""KHI growth rate run (share/picongpu/tests/KHI_growthRate): prints the growth rates per B component"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from picongpu_b200 import param as prm, picstep
p = prm.khi_params(grid=(192, 512, 12), delta_t_si=1.79e-16 * 0.86, cell_si=(9.34635e-8 * 0.86,) * 3)
s = picstep.Simulation(p, device=0, exact=False)
s.init_khi()
g, n = p.guard_cells, p.grid
en = []
t0 = time.time()
for k in range(301):
    B = s.download_field(picstep.FIELD_B)
    I = B[:, g[2]:g[2] + n[2], g[1]:g[1] + n[1], g[0]:g[0] + n[0]].astype(np.float64)
    en.append([(I[c] ** 2).sum() for c in range(3)])
    s.step(10)
s.sync()
print("wall", time.time() - t0)
en = np.array(en)
gamma = 1.021
omega = np.sqrt(1e25 * prm.ELECTRON_CHARGE_SI ** 2 / (prm.EPS0_SI if hasattr(prm, "EPS0_SI") else 8.8541878128e-12) / gamma / prm.ELECTRON_MASS_SI)
t = np.arange(301) * 10 * p.delta_t_si * omega
theory = 1 / (8 ** 0.5 * gamma)
for c in range(3):
    f = en[:, c]
    G = 0.5 * np.log(f[3:] / f[1:-2]) / (t[3:] - t[1:-2])
    print("B%s: max growth %.4f at t=%.1f (theory %.4f, diff %.1f %%), energy first/last %.3e %.3e" % ("xyz"[c], np.nanmax(G), t[1 + np.nanargmax(G)], theory, 100 * (theory - np.nanmax(G)) / np.nanmax(G), f[1], f[-1]))
np.save(os.path.join(ROOT, "gpurun_out", "khi_growth_en.npy"), en)
