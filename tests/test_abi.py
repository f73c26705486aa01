"""CPU-side checks of the product boundary: the C-ABI library loads, exports every symbol include/picstep.h
declares, fails loudly without a GPU (no CPU fallback), and its host-only helpers follow the reference."""
import ctypes as C
import os
import re

import pytest
import torch

from picongpu_b200 import param as prm
from picongpu_b200 import picstep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    from picongpu_b200 import build

    build.build_all()


def _declared():
    txt = open(os.path.join(ROOT, "include", "picstep.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|const char\*)\s+(picstep_[a-z_0-9]+)\s*\(", txt, re.M)))


@pytest.mark.parametrize("exact", [False, True])
def test_library_exports_every_declared_symbol(exact):
    L = picstep.load(exact)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert sorted(picstep.SYMBOLS) == names
    v = L.picstep_version().decode()
    assert "sm_100a" in v and (("exact" in v) == exact)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = prm.khi_params(grid=(16, 16, 8))
    with pytest.raises(picstep.PicstepError, match="no CUDA device|CUDA"):
        picstep.Simulation(p)


def test_neighbor_ranks_cartesian():
    """x-fastest rank linearisation, periodic wrap through the topology (CommunicatorMPI.cpp:70-111)."""
    assert picstep.neighbor_ranks((1, 4, 1), (1, 1, 1), 0, 1) == (3, 1)
    assert picstep.neighbor_ranks((1, 4, 1), (1, 1, 1), 3, 1) == (2, 0)
    assert picstep.neighbor_ranks((1, 4, 1), (1, 0, 1), 0, 1) == (-1, 1)
    assert picstep.neighbor_ranks((1, 4, 1), (1, 0, 1), 3, 1) == (2, -1)
    assert picstep.neighbor_ranks((2, 2, 2), (1, 1, 1), 5, 0) == (4, 4)
    assert picstep.neighbor_ranks((2, 2, 2), (1, 1, 1), 5, 2) == (1, 1)
    assert picstep.neighbor_ranks((1, 2, 1), (1, 1, 1), 1, 1) == (0, 0)


def test_exchange_widths_follow_margins():
    """E/B: max(interpolation, solver) margins (EMFieldBase.x.cpp:58-110); J: current solver margins."""
    E, J = picstep.FIELD_E, picstep.FIELD_J
    assert picstep.exchange_widths(prm.SHAPE_TSC, prm.SOLVER_YEE, 1, E, 0) == (1, 2)
    assert picstep.exchange_widths(prm.SHAPE_CIC, prm.SOLVER_YEE, 1, E, 1) == (1, 1)
    assert picstep.exchange_widths(prm.SHAPE_PQS, prm.SOLVER_YEE, 1, E, 2) == (2, 2)
    assert picstep.exchange_widths(prm.SHAPE_PCS, prm.SOLVER_YEE, 1, E, 2) == (2, 3)
    assert picstep.exchange_widths(prm.SHAPE_CIC, prm.SOLVER_LEHE, 1, E, 1) == (1, 2)
    assert picstep.exchange_widths(prm.SHAPE_TSC, prm.SOLVER_YEE, 1, J, 0) == (2, 3)
    assert picstep.exchange_widths(prm.SHAPE_CIC, prm.SOLVER_YEE, 1, J, 0) == (1, 2)
    assert picstep.exchange_widths(prm.SHAPE_PCS, prm.SOLVER_YEE, 1, J, 0) == (3, 4)


def test_unit_system_khi():
    """PIC units of the KelvinHelmholtz example (simulation.unitless:420-480): dt = c = 1, dx = 1.7417."""
    p = prm.khi_params()
    assert p.dt == 1.0 and p.c == 1.0
    assert abs(p.cell_size[0] - 9.34635e-8 / (1.79e-16 * 2.99792458e8)) < 1e-6
    assert abs(p.base_mass * p.typical_num_particles_per_macro - 1.0) < 1e-6
    assert abs(p.base_charge * p.typical_num_particles_per_macro + 1.0) < 1e-6
    assert p.cfl_ok()
    # omega_pe * dt for n0 = 1e25 / m^3
    wpe = (1e25 * 1.602176634e-19**2 / (9.1093837139e-31 * 8.8541878128e-12)) ** 0.5
    n_pic = p.real_particles_per_cell / (p.cell_size[0] ** 3) / p.typical_num_particles_per_macro
    wpe_pic = (n_pic * (p.base_charge * p.typical_num_particles_per_macro) ** 2 / (p.eps0 * p.base_mass * p.typical_num_particles_per_macro)) ** 0.5
    assert abs(wpe_pic - wpe * 1.79e-16) / (wpe * 1.79e-16) < 1e-4


def test_moving_window_schedule():
    """MovingWindow::getCurrentSlideInfo (MovingWindow.hpp:44-170): hand-computed answers for 4 GPUs x 16 cells,
    dy = 1, c*dt = 0.5, movePoint 0.5 (window starts to move in step 47, slides in steps 79, 111, 143, ...), and a
    straight restatement of the formulas over a range of steps."""
    import math

    G, L, dy, cdt, mp = 64, 16, 1.0, 0.5, 0.5
    info = lambda st: picstep.moving_window_info(G, L, dy, cdt, mp, st)
    assert info(0) == (False, 0) and info(46) == (False, 0) and info(47) == (False, 0)
    assert info(49) == (False, 1) and info(78) == (False, 15)
    assert info(79) == (True, 0) and info(80) == (False, 0) and info(81) == (False, 1)
    slides = [st for st in range(400) if info(st)[0]]
    assert slides[:4] == [79, 111, 143, 175]

    def ref(st, G, L, dy, cdt, mp):
        win = G - L
        start = math.ceil(win * (1.0 - mp))
        first_slide = math.ceil((G - start) * dy / cdt) - 1
        first_move = math.ceil((win - start) * dy / cdt) - 1
        if first_move > st:
            return False, 0
        pos = math.floor(cdt * st / dy) + start
        nxt = math.floor(cdt * (st + 1) / dy) + start
        return (first_slide <= st and (nxt % L) < (pos % L)), nxt % L

    for (G, L, dy, cdt, mp) in ((64, 16, 1.0, 0.5, 0.5), (2048, 256, 1.7417, 1.0, 0.9), (96, 32, 0.8, 0.45, 0.0)):
        for st in range(0, 3000, 7):
            assert picstep.moving_window_info(G, L, dy, cdt, mp, st) == ref(st, G, L, dy, cdt, mp)


def test_params_struct_layout_matches_the_header(tmp_path):
    """The ctypes mirror of `picstep_params` (picongpu_b200/picstep.py) has the size and field offsets the C compiler
    gives the struct in include/picstep.h (a plain C translation unit must be able to include the header)."""
    import subprocess

    fields = [f[0] for f in picstep.Params._fields_]
    src = tmp_path / "layout.c"
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "picstep.h"', "int main(void){",
             'printf("%zu\\n", sizeof(picstep_params));']
    lines += ['printf("%%zu\\n", offsetof(picstep_params, %s));' % f for f in fields]
    lines += ["return 0;}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == C.sizeof(picstep.Params)
    for f, off in zip(fields, out[1:]):
        assert getattr(picstep.Params, f).offset == off, f
