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


def test_laser_and_pml_unit_conversions():
    """Unitless incident-field and PML parameters of examples/LaserWakefield (incidentField.param, simulation.param;
    BaseParam.hpp:43-180, GaussianPulse.hpp:93-110, fieldAbsorber.unitless:72-104) against their SI definitions computed
    independently: a0 = e E0 / (m_e c omega), omega dt, Rayleigh length, TIME_SHIFT, focus on the transversal centre,
    sigma_opt = 0.8 (m + 1) / (Z0 dx)."""
    import math

    dt, cell = 1.39e-16, (0.1772e-6, 0.4430e-7, 0.1772e-6)
    p = prm.khi_params(grid=(192, 2048, 192), delta_t_si=dt, cell_si=cell, periodic=(0, 0, 0))
    las = prm.gaussian_pulse_laser(p)
    lam, c = 0.8e-6, prm.SPEED_OF_LIGHT_SI
    omega_si = 2.0 * math.pi * c / lam
    e0_si = abs(las["amplitude"]) * p.unit_efield
    a0 = abs(prm.ELECTRON_CHARGE_SI) * e0_si / (prm.ELECTRON_MASS_SI * c * omega_si)
    assert abs(a0 - 8.0) < 1e-5 and las["amplitude"] > 0  # UNITCONV_A0_to_Amplitude_SI = -2 pi / lambda * m_e c^2 / q_e, q_e < 0
    assert abs(las["omega"] - omega_si * dt) < 1e-6 * omega_si * dt
    assert abs(las["wave_length"] * p.unit_length - lam) < 1e-6 * lam
    w0_si = 5.0e-6 / 1.17741
    assert abs(las["w0"] * p.unit_length - w0_si) < 1e-6 * w0_si
    z_r = math.pi * las["w0"] ** 2 / las["wave_length"] * p.unit_length
    assert abs(z_r - math.pi * w0_si**2 / lam) < 1e-5 * z_r  # 70.8 um
    assert abs(las["pulse_duration"] - 5.0e-15 / dt) < 1e-5 * (5.0e-15 / dt)
    assert abs(las["time_shift"] + 0.5 * 15.0 * 5.0e-15 / dt) < 1e-4 * (37.5e-15 / dt)  # the pulse peak enters after 37.5 fs
    assert abs(las["focus_position"][1] * p.unit_length - 4.62e-5) < 1e-10 and las["focus_origin_center"] == (1, 0, 1)
    assert las["polarisation"] == 1 and las["position"] == ((16, -16),) * 3 and las["modes"] == (1.0,)
    # plane wave of the same example family: RAMP_INIT, plateau
    pw = prm.plane_wave_laser(p, a0=1.5, pulse_duration_si=10.615e-15 / 4.0, nofocus_constant_si=13.34e-15, ramp_init=20.6146)
    assert abs(pw["nofocus_constant"] - 13.34e-15 / dt) < 1e-5 * (13.34e-15 / dt) and abs(pw["ramp_init"] - 20.6146) < 1e-5
    # PML: SIGMA_OPT_SI = 0.8 (order + 1) / (Z0 cell), normalised by eps0 and the unit of time; alpha 0.2 S/m
    pm = prm.pml_params(p)
    z0 = prm.MUE0_SI * c
    eps0 = 1.0 / (prm.MUE0_SI * c * c)
    for d in range(3):
        assert abs(pm["sigma_max"][d] - 0.8 * 5.0 / (z0 * cell[d]) / eps0 * dt) < 1e-5 * pm["sigma_max"][d]
        assert abs(pm["alpha_max"][d] - 0.2 / eps0 * dt) < 1e-5 * pm["alpha_max"][d]
    assert pm["kappa_max"] == (1.0, 1.0, 1.0) and pm["sigma_kappa_grading_order"] == 4.0 and pm["alpha_grading_order"] == 1.0
    # ExpRampWithPrepulse: ordering check of the reference's static_assert
    with pytest.raises(ValueError):
        prm.exp_ramp_with_prepulse_laser(p, time_points_si=(-100e-15, -300e-15, -50e-15))


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


def test_plain_c_caller(tmp_path):
    """A C99 program (no C++, no Python) includes include/picstep.h, loads libpicstep.so and calls through the ABI: the
    version string comes back, and without a CUDA device picstep_create refuses with PICSTEP_ERR_NOGPU and says why."""
    import subprocess

    src = tmp_path / "caller.c"
    src.write_text(r"""
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "picstep.h"
typedef const char* (*version_fn)(void);
typedef int (*create_fn)(const picstep_params*, picstep_ctx**);
typedef const char* (*error_fn)(const picstep_ctx*);
int main(int argc, char** argv)
{
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if(!h) { printf("dlopen: %s\n", dlerror()); return 2; }
    version_fn version = (version_fn) dlsym(h, "picstep_version");
    create_fn create = (create_fn) dlsym(h, "picstep_create");
    error_fn last_error = (error_fn) dlsym(h, "picstep_last_error");
    if(!version || !create || !last_error) return 3;
    picstep_params p;
    memset(&p, 0, sizeof p);
    p.grid[0] = p.grid[1] = 16; p.grid[2] = 8;
    p.supercell[0] = p.supercell[1] = 8; p.supercell[2] = 4;
    p.guard_supercells[0] = p.guard_supercells[1] = p.guard_supercells[2] = 1;
    p.devices[0] = p.devices[1] = p.devices[2] = 1;
    p.periodic[0] = p.periodic[1] = p.periodic[2] = 1;
    p.shape = PICSTEP_SHAPE_TSC; p.pusher = PICSTEP_PUSHER_BORIS;
    p.current_solver = PICSTEP_CURRENT_ESIRKEPOV; p.field_solver = PICSTEP_SOLVER_YEE;
    picstep_ctx* ctx = NULL;
    int rc = create(&p, &ctx);
    printf("%s|%d|%s\n", version(), rc, rc ? last_error(NULL) : "created");
    return 0;
}
""")
    exe = tmp_path / "caller"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-ldl"], check=True)
    out = subprocess.run([str(exe), picstep.lib_path(False)], capture_output=True, text=True, check=True).stdout.strip()
    version, rc, msg = out.split("|")
    assert version.startswith("picstep") and "sm_100a" in version
    import torch

    if not torch.cuda.is_available():
        assert int(rc) == 5 and "no CPU fallback" in msg  # PICSTEP_ERR_NOGPU
