"""Generate tests/golden/current_deposition.npz from the reference's own Python Esirkepov implementation.

Run in the build container (needs /root/reference):  python tests/golden/make_current_deposition_golden.py

It imports share/picongpu/tests/CurrentDeposition/lib/python/test/CurrentDeposition/{grid_class,
assignment_and_W_func}.py UNMODIFIED and stores, for a list of one-particle moves, the three
current-deposition-vector grids W_x, W_y, W_z (grid_class.py:95-135, assignment_and_W_func.py:94-117).
The reference's `current_density_field` is NOT used for y/z: its loops `range(start[2], start[2])`
are empty (grid_class.py:184,192).  The tests rebuild J from W with the recursion the x branch states
(J[i] = factor * W[i] + J[i-1], grid_class.py:176-181) applied along the matching axis.
"""
import os
import sys

import numpy as np

REF = "/root/reference/share/picongpu/tests/CurrentDeposition/lib/python/test/CurrentDeposition"
sys.path.insert(0, REF)
from grid_class import grid  # noqa: E402

rng = np.random.RandomState(1234)
cases = []
# the reference test: one electron, beta=0.999, directions (1,0,0),(1,1,0),(1,1,1); KHI-like dt/dx
dx = 1.7417
for order in (1, 2, 3):
    for direction in ((1, 0, 0), (1, 1, 0), (1, 1, 1), (-1, 1, -1)):
        d = np.array(direction, float)
        d = d / np.linalg.norm(d)
        delta = 0.999 * d / dx
        for start in ((0.5, 0.5, 0.5), (0.93, 0.07, 0.6), (0.02, 0.98, 0.51)):
            cases.append((order, np.array(start), delta))
    for _ in range(6):
        start = rng.uniform(0, 1, 3)
        delta = rng.uniform(-0.57, 0.57, 3)
        cases.append((order, start, delta))

orders, starts, ends, offs2, Wx, Wy, Wz = [], [], [], [], [], [], []
for order, start, delta in cases:
    g = grid(order)
    gx, gy, gz = g.create_grid()
    end_abs = start + delta
    off2 = np.floor(end_abs).astype(int)
    pos2 = end_abs - off2
    s, e = g.particle_step(start, pos2, np.zeros(3, int), off2)
    wx, wy, wz = g.current_deposition_field(s, e, gx, gy, gz)
    n = 7  # pad all orders to the largest minimal grid (order 3 -> 7)
    pad = lambda a: np.pad(np.array(a), ((0, n - g.num_cells),) * 3)
    orders.append(order)
    starts.append(start)
    ends.append(pos2)
    offs2.append(off2)
    Wx.append(pad(wx))
    Wy.append(pad(wy))
    Wz.append(pad(wz))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "current_deposition.npz")
np.savez_compressed(out, order=np.array(orders), start=np.array(starts), end=np.array(ends), off2=np.array(offs2),
                    Wx=np.array(Wx), Wy=np.array(Wy), Wz=np.array(Wz))
print("wrote", out, len(orders), "cases")
