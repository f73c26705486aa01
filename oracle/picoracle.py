"""ctypes binding of the CPU oracle (oracle/libpicoracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package picongpu_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class OrcParams(C.Structure):
    _fields_ = [
        ("n", C.c_int * 3),
        ("sc", C.c_int * 3),
        ("g", C.c_int * 3),
        ("cell", C.c_float * 3),
        ("dt", C.c_float),
        ("c", C.c_float),
        ("eps0", C.c_float),
        ("mue0", C.c_float),
        ("base_mass", C.c_float),
        ("base_charge", C.c_float),
        ("shape", C.c_int),
        ("pusher", C.c_int),
        ("current", C.c_int),
        ("solver", C.c_int),
        ("lehe_dir", C.c_int),
        ("wrap", C.c_int * 3),
        ("current_interp", C.c_int),
        ("open", (C.c_int * 2) * 3),
        ("absorber_cells", (C.c_int * 2) * 3),
        ("absorber_strength", (C.c_float * 2) * 3),
    ]


LASER_MAX_MODES = 8


class OrcLaser(C.Structure):
    """Incident field (PlaneWave or GaussianPulse) on the YMin Huygens surface (oracle/picoracle.cpp: struct OrcLaser)"""

    _fields_ = [
        ("polarisation", C.c_int),
        ("offset_ymin", C.c_int),
        ("amplitude", C.c_float),
        ("omega", C.c_float),
        ("pulse_duration", C.c_float),
        ("nofocus_constant", C.c_float),
        ("ramp_init", C.c_float),
        ("phase", C.c_float),
        ("pol", C.c_float * 3),
        ("time_delay", C.c_float),
        ("global_y_offset", C.c_int),
        ("profile", C.c_int),
        ("position", (C.c_int * 2) * 3),
        ("global_size", C.c_int * 3),
        ("periodic", C.c_int * 3),
        ("w0", C.c_float),
        ("wave_length", C.c_float),
        ("time_shift", C.c_float),
        ("focus_position", C.c_float * 3),
        ("focus_origin_center", C.c_int * 3),
        ("tilt", C.c_float * 2),
        ("n_modes", C.c_int),
        ("modes", C.c_float * LASER_MAX_MODES),
        ("mode_phases", C.c_float * LASER_MAX_MODES),
        ("w0_axis", C.c_float * 2),
        ("profile_params", C.c_float * 16),
    ]


def laser_position(las):
    """POSITION[3][2] of a laser dict: 'position' if given, else {offset_ymin, -offset_ymin} on every axis"""
    if las.get("position") is not None:
        return tuple((int(a), int(b)) for a, b in las["position"])
    o = int(las["offset_ymin"])
    return ((o, -o), (o, -o), (o, -o))


def make_laser(cfg):
    """cfg.laser: dict with the profile parameters in PIC units (picongpu_b200.param.plane_wave_laser /
    gaussian_pulse_laser) or None"""
    g = (lambda k, dflt=None: cfg.get(k, dflt)) if isinstance(cfg, dict) else (lambda k, dflt=None: getattr(cfg, k, dflt))
    las = g("laser")
    if not las:
        return None
    L = OrcLaser()
    L.polarisation = int(las["polarisation"])
    L.offset_ymin = int(las["offset_ymin"])
    for k in ("amplitude", "omega", "pulse_duration", "nofocus_constant", "ramp_init", "phase", "time_delay"):
        setattr(L, k, float(las.get(k, 0.0)))
    for d in range(3):
        L.pol[d] = float(las["pol"][d])
    go = g("global_offset", (0, 0, 0))
    L.global_y_offset = int(go[1])
    L.profile = int(las.get("profile", 0))
    pos = laser_position(las)
    gg = g("global_grid") or g("grid")
    per = g("periodic", (1, 0, 1))
    for d in range(3):
        L.position[d][0], L.position[d][1] = pos[d]
        L.global_size[d] = int(gg[d])
        L.periodic[d] = int(per[d])
        L.focus_position[d] = float(las.get("focus_position", (0.0, 0.0, 0.0))[d])
        L.focus_origin_center[d] = int(las.get("focus_origin_center", (0, 0, 0))[d])
    if L.position[1][0] != L.offset_ymin:
        raise ValueError("laser: position[1][0] must equal offset_ymin")
    L.w0 = float(las.get("w0", 0.0))
    L.wave_length = float(las.get("wave_length", 0.0))
    L.time_shift = float(las.get("time_shift", 0.0))
    L.tilt[0], L.tilt[1] = (float(v) for v in las.get("tilt", (0.0, 0.0)))
    modes = list(las.get("modes", (1.0,)))
    phases = list(las.get("mode_phases", (0.0,) * len(modes)))
    if not 1 <= len(modes) <= LASER_MAX_MODES or len(phases) != len(modes):
        raise ValueError("laser: 1..8 Laguerre modes with one phase each")
    L.n_modes = len(modes)
    for m in range(len(modes)):
        L.modes[m], L.mode_phases[m] = float(modes[m]), float(phases[m])
    L.w0_axis[0], L.w0_axis[1] = (float(v) for v in las.get("w0_axis", (0.0, 0.0)))
    for k, v in enumerate(las.get("profile_params", ())):
        L.profile_params[k] = float(v)
    return L


class OrcPml(C.Structure):
    """PML absorber parameters (oracle/picoracle.cpp: struct OrcPml)"""

    _fields_ = [
        ("thickness", (C.c_int * 2) * 3),
        ("sigmaMax", C.c_float * 3),
        ("kappaMax", C.c_float * 3),
        ("alphaMax", C.c_float * 3),
        ("sigmaKappaGradingOrder", C.c_float),
        ("alphaGradingOrder", C.c_float),
    ]


def make_pml(cfg):
    """absorber_kind 2: thickness from absorber_cells at the open faces, parameters from cfg.pml (param.pml_params)"""
    g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    if not (_has(cfg, "absorber_kind") and int(g("absorber_kind")) == 2):
        return None
    M = OrcPml()
    pm = g("pml")
    for d in range(3):
        for sd in range(2):
            M.thickness[d][sd] = int(g("absorber_cells")[d][sd]) if int(g("open")[d][sd]) else 0
        M.sigmaMax[d] = float(pm["sigma_max"][d])
        M.kappaMax[d] = float(pm["kappa_max"][d])
        M.alphaMax[d] = float(pm["alpha_max"][d])
    if _has(cfg, "moving_window") and int(g("moving_window")):
        M.thickness[1][1] = 0
    M.sigmaKappaGradingOrder = float(pm["sigma_kappa_grading_order"])
    M.alphaGradingOrder = float(pm["alpha_grading_order"])
    return M


class OrcSpecies(C.Structure):
    _fields_ = [
        ("massRatio", C.c_float),
        ("chargeRatio", C.c_float),
        ("np", C.c_int64),
        ("pos", C.c_void_p),
        ("mom", C.c_void_p),
        ("w", C.c_void_p),
        ("cell", C.c_void_p),
    ]


def build(force=False):
    """Compile oracle/picoracle.cpp -> oracle/libpicoracle.so (g++, OpenMP if available)."""
    so = os.path.join(_HERE, "libpicoracle.so")
    src = os.path.join(_HERE, "picoracle.cpp")
    if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(src):
        return so
    base = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-o", so, src]
    for cxx, omp in (("/usr/bin/g++", ["-fopenmp"]), ("g++", ["-fopenmp"]), ("/usr/bin/g++", []), ("g++", [])):
        try:
            r = subprocess.run([cxx] + omp + base, capture_output=True, text=True)
        except FileNotFoundError:
            continue
        if r.returncode == 0:
            return so
    raise RuntimeError("could not build oracle: " + r.stderr)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    P = C.POINTER(OrcParams)
    L.orc_shape_array.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, f32p]
    L.orc_shape_eval.argtypes = [C.c_int, C.c_int, C.c_float]
    L.orc_shape_eval.restype = C.c_float
    L.orc_shape_unit_test.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p]
    L.orc_move_particle.argtypes = [i32p, f32p, C.c_int, f32p, C.POINTER(C.c_int)]
    L.orc_move_particle.restype = C.c_int
    L.orc_size_last_frame.argtypes = [C.c_uint, C.c_uint]
    L.orc_size_last_frame.restype = C.c_uint
    L.orc_lehe_coeff.argtypes = [P, C.c_int, f32p]
    L.orc_push_one.argtypes = [P, C.c_float, C.c_float, C.c_float, f32p, f32p, f32p, f32p]
    L.orc_gather.argtypes = [P, f32p, f32p, C.c_int64, f32p, i32p, f32p, f32p]
    L.orc_push.argtypes = [P, C.c_float, C.c_float, f32p, f32p, C.c_int64, f32p, f32p, f32p, i32p, C.c_void_p, C.c_void_p]
    L.orc_deposit.argtypes = [P, C.c_float, C.c_float, f32p, C.c_int64, f32p, f32p, f32p, i32p]
    L.orc_deposit_one.argtypes = [P, f32p, i32p, f32p, f32p, C.c_float]
    L.orc_guard_copy.argtypes = [P, f32p]
    L.orc_guard_add.argtypes = [P, f32p]
    L.orc_halo_axis.argtypes = [P, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_update_b_half.argtypes = [P, f32p, f32p]
    L.orc_update_e.argtypes = [P, f32p, f32p]
    L.orc_add_current.argtypes = [P, f32p, f32p]
    L.orc_field_energy.argtypes = [P, f32p, f32p, f64p]
    L.orc_particle_energy.argtypes = [P, C.c_float, C.c_int64, f32p, f32p, f64p]
    L.orc_charge_density.argtypes = [P, C.c_float, f32p, C.c_int64, f32p, f32p, i32p]
    L.orc_gauss_residual.argtypes = [P, f32p, f32p]
    L.orc_gauss_residual.restype = C.c_double
    L.orc_step.argtypes = [P, f32p, f32p, f32p, C.c_int, C.POINTER(OrcSpecies)]
    L.orc_khi_init.argtypes = [P, i32p, i32p, i32p, C.c_float, C.c_float, C.c_double, C.c_double, C.c_double, C.c_uint32,
                               f32p, f32p, f32p, i32p, f32p, f32p, f32p, i32p]
    L.orc_absorb.argtypes = [P, f32p]
    L.orc_incident_update.argtypes = [P, C.POINTER(OrcLaser), f32p, C.c_int, C.c_float]
    L.orc_update_e_pml.argtypes = [P, C.POINTER(OrcPml), f32p, f32p, f32p]
    L.orc_update_b_half_pml.argtypes = [P, C.POINTER(OrcPml), f32p, f32p, f32p, C.c_int]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    _LIB = L
    return L


def make_params(cfg):
    """cfg: a picongpu_b200.param.SimParams-like object or dict with the same field names."""
    g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    p = OrcParams()
    for d in range(3):
        p.n[d] = int(g("grid")[d])
        p.sc[d] = int(g("supercell")[d])
        p.g[d] = int(g("supercell")[d]) * int(g("guard_supercells")[d])
        p.cell[d] = float(g("cell_size")[d])
        p.wrap[d] = int(g("wrap")[d]) if _has(cfg, "wrap") else 1
    p.dt = float(g("dt"))
    p.c = float(g("c"))
    p.eps0 = float(g("eps0"))
    p.mue0 = float(g("mue0"))
    p.base_mass = float(g("base_mass"))
    p.base_charge = float(g("base_charge"))
    p.shape = int(g("shape"))
    p.pusher = int(g("pusher"))
    p.current = int(g("current_solver"))
    p.solver = int(g("field_solver"))
    p.lehe_dir = int(g("lehe_dir"))
    p.current_interp = int(g("current_interpolation")) if _has(cfg, "current_interpolation") else 0
    absorbing = _has(cfg, "absorber_kind") and int(g("absorber_kind")) == 1
    for d in range(3):
        for sd in range(2):
            p.open[d][sd] = int(g("open")[d][sd]) if _has(cfg, "open") else 0
            p.absorber_cells[d][sd] = int(g("absorber_cells")[d][sd]) if absorbing else 0
            p.absorber_strength[d][sd] = float(g("absorber_strength")[d][sd]) if absorbing else 0.0
    if _has(cfg, "moving_window") and int(g("moving_window")):
        p.absorber_cells[1][1] = 0  # no absorber on the +y side while the window slides (Exponential.hpp:97-101)
    return p


def _has(cfg, k):
    return (k in cfg) if isinstance(cfg, dict) else hasattr(cfg, k)


class Oracle:
    """Thin object wrapper: holds OrcParams and exposes the stage calls on numpy arrays."""

    def __init__(self, cfg):
        self.p = make_params(cfg)
        self.L = lib()
        self.n = tuple(self.p.n)
        self.g = tuple(self.p.g)
        self.N = tuple(self.p.n[d] + 2 * self.p.g[d] for d in range(3))
        self.laser = make_laser(cfg)
        self.pml = make_pml(cfg)
        if self.pml is not None:  # convolutional fields psi_yx, zx, xy, zy, xz, yz of E and B
            self.psiE = np.zeros((6, self.N[2], self.N[1], self.N[0]), np.float32)
            self.psiB = np.zeros((6, self.N[2], self.N[1], self.N[0]), np.float32)
        self.step_index = 0  # step_open counts the steps (the incident field is a function of time)

    def field(self):
        return np.zeros((3, self.N[2], self.N[1], self.N[0]), np.float32)

    def interior(self, F):
        g, n = self.g, self.n
        return F[..., g[2]:g[2] + n[2], g[1]:g[1] + n[1], g[0]:g[0] + n[0]]

    def gather(self, E, B, pos, cell):
        npart = pos.shape[1]
        Eo = np.empty((3, npart), np.float32)
        Bo = np.empty((3, npart), np.float32)
        self.L.orc_gather(C.byref(self.p), E, B, npart, pos, cell, Eo, Bo)
        return Eo, Bo

    def push(self, mr, cr, E, B, pos, mom, w, cell, want_mask=False, want_cell3=False):
        npart = pos.shape[1]
        mask = np.zeros(npart, np.uint8) if want_mask else None
        cell3 = np.zeros((3, npart), np.int32) if want_cell3 else None
        self.L.orc_push(C.byref(self.p), mr, cr, E, B, npart, pos, mom, w, cell,
                        mask.ctypes.data if want_mask else None, cell3.ctypes.data if want_cell3 else None)
        return mask, cell3

    def deposit(self, mr, cr, J, pos, mom, w, cell):
        self.L.orc_deposit(C.byref(self.p), mr, cr, J, pos.shape[1], pos, mom, w, cell)

    def guard_copy(self, F):
        self.L.orc_guard_copy(C.byref(self.p), F)

    def guard_add(self, F):
        self.L.orc_guard_add(C.byref(self.p), F)

    def halo_axis(self, F, axis, lo, up, add=False):
        self.L.orc_halo_axis(C.byref(self.p), F, F.shape[0], axis, lo, up, 1 if add else 0)

    def update_b_half(self, E, B, first_half=False):
        """first_half: updateBFirstHalf (advances the PML's psiB; FDTDBase.hpp:200-211), else updateBSecondHalf"""
        if self.pml is not None:
            self.L.orc_update_b_half_pml(C.byref(self.p), C.byref(self.pml), E, B, self.psiB, 1 if first_half else 0)
        else:
            self.L.orc_update_b_half(C.byref(self.p), E, B)

    def update_e(self, E, B):
        if self.pml is not None:
            self.L.orc_update_e_pml(C.byref(self.p), C.byref(self.pml), E, B, self.psiE)
        else:
            self.L.orc_update_e(C.byref(self.p), E, B)

    def add_current(self, E, J):
        self.L.orc_add_current(C.byref(self.p), E, J)

    def step(self, E, B, J, species):
        """species: list of dicts(massRatio, chargeRatio, pos, mom, w, cell) updated in place."""
        arr = (OrcSpecies * len(species))()
        for i, s in enumerate(species):
            arr[i].massRatio = s["massRatio"]
            arr[i].chargeRatio = s["chargeRatio"]
            arr[i].np = s["pos"].shape[1]
            for k in ("pos", "mom", "w", "cell"):
                assert s[k].flags["C_CONTIGUOUS"]
                setattr(arr[i], k, s[k].ctypes.data)
        self.L.orc_step(C.byref(self.p), E, B, J, len(species), arr)

    def absorb(self, F):
        self.L.orc_absorb(C.byref(self.p), F)

    def incident_update(self, F, updated_is_e, step):
        """incidentField::Solver::updateE / updateBHalf at the (fractional) step `step`; no-op without a laser"""
        if self.laser is not None:
            self.L.orc_incident_update(C.byref(self.p), C.byref(self.laser), F, 1 if updated_is_e else 0, float(step))

    def step_open(self, E, B, J, species, background_j=None):
        """One PIC step for a single domain with any mix of periodic and open (absorbing) axes, composed of the stage
        calls in the order of Simulation::runOneStep (Simulation.hpp:526-541).  Particles that leave through an open
        face are deleted before the current deposition (Particles.tpp:322-368: push, shift, applyBoundary).
        species dicts are updated (arrays are REPLACED when particles were absorbed).  Exchange passes are done
        axis by axis over the periodic axes only (orc_halo_axis).  `background_j(J, step)`: optional FieldBackgroundJ."""
        periodic = [bool(self.p.wrap[d]) for d in range(3)]
        g = self.g
        J[...] = 0
        for s in species:
            self.push(s["massRatio"], s["chargeRatio"], E, B, s["pos"], s["mom"], s["w"], s["cell"])
            keep = s["cell"] >= 0
            if not keep.all():
                s["pos"] = np.ascontiguousarray(s["pos"][:, keep])
                s["mom"] = np.ascontiguousarray(s["mom"][:, keep])
                s["w"] = np.ascontiguousarray(s["w"][keep])
                s["cell"] = np.ascontiguousarray(s["cell"][keep])

        def copy_guards(F, width=None):
            for a in range(3):
                if periodic[a]:
                    w = g[a] if width is None else width
                    self.halo_axis(F, a, w, w, add=False)

        # update_beforeCurrent (FDTDBase.hpp:97-121) with the incident field source (:108-117)
        self.update_b_half(E, B)
        self.incident_update(B, False, self.step_index)
        copy_guards(B)
        self.incident_update(E, True, self.step_index + 0.5)
        self.update_e(E, B)
        if background_j is not None:  # stage::CurrentBackground (Simulation.hpp:538): J += FieldBackgroundJ(cell, step)
            background_j(J, self.step_index)
        for s in species:
            if s["w"].shape[0]:
                self.deposit(s["massRatio"], s["chargeRatio"], J, s["pos"], s["mom"], s["w"], s["cell"])
        for a in range(3):
            if periodic[a]:
                self.halo_axis(J, a, g[a], g[a], add=True)
        if self.p.current_interp == 1:
            copy_guards(J, 1)
        self.add_current(E, J)
        self.absorb(E)
        self.incident_update(B, False, self.step_index + 1.0)  # FDTDBase.hpp:161-166
        copy_guards(E)
        self.update_b_half(E, B, first_half=True)
        self.absorb(B)
        copy_guards(B)
        self.step_index += 1

    def field_energy(self, E, B):
        out = np.zeros(2, np.float64)
        self.L.orc_field_energy(C.byref(self.p), E, B, out)
        return out

    def particle_energy(self, mr, mom, w):
        out = np.zeros(2, np.float64)
        self.L.orc_particle_energy(C.byref(self.p), mr, mom.shape[1], mom, w, out)
        return out

    def charge_density(self, cr, rho, pos, w, cell):
        self.L.orc_charge_density(C.byref(self.p), cr, rho, pos.shape[1], pos, w, cell)

    def gauss_residual(self, E, species):
        rho3 = np.zeros((3,) + tuple(reversed(self.N)), np.float32)  # only component 0 used
        for s in species:
            self.charge_density(s["chargeRatio"], rho3[0], s["pos"], s["w"], s["cell"])
        self.guard_add(rho3)
        return self.L.orc_gauss_residual(C.byref(self.p), E, rho3[0])
