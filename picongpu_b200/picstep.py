"""ctypes binding of libpicstep.so (include/picstep.h) and a thin host-side mirror of the reference's simulation
loop (`Simulation::runOneStep`, include/picongpu/simulation/control/Simulation.hpp:522-542).

There is deliberately no CPU path here: if the CUDA library is missing or no GPU is present every entry point
raises.  The CPU oracle lives in oracle/ and is only used by tests and the bench's cpu_baseline.
"""
import ctypes as C
import os

import numpy as np

from . import param as prm

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

FIELD_E, FIELD_B, FIELD_J = 0, 1, 2
REDUCE_FIELD_ENERGY, REDUCE_PARTICLE_ENERGY, REDUCE_GAUSS, REDUCE_PARTICLE_COUNT, REDUCE_SLOW_PATH = 0, 1, 2, 3, 4
STAGES = ["current_reset", "push", "migrate", "field_before", "deposit", "add_current", "field_after"]

# every symbol include/picstep.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "picstep_version", "picstep_last_error", "picstep_create", "picstep_destroy", "picstep_species_add", "picstep_species_set_policy",
    "picstep_fields_upload", "picstep_fields_download", "picstep_fields_upload_soa", "picstep_fields_download_soa",
    "picstep_particles_upload", "picstep_particles_count", "picstep_particles_download", "picstep_supercell_counts",
    "picstep_init_khi", "picstep_init_thermal", "picstep_current_reset", "picstep_push", "picstep_migrate",
    "picstep_field_update_before_current", "picstep_deposit", "picstep_add_current",
    "picstep_field_update_after_current", "picstep_field_exchange", "picstep_step", "picstep_step_host",
    "picstep_sync", "picstep_reduce", "picstep_debug_gather", "picstep_comm_unique_id", "picstep_comm_init",
    "picstep_launch_count", "picstep_stage_times", "picstep_overlap_times", "picstep_stream", "picstep_neighbor_ranks",
    "picstep_exchange_widths", "picstep_slide", "picstep_moving_window_info", "picstep_window_neighbors",
]


class PicstepError(RuntimeError):
    pass


class Params(C.Structure):
    """struct picstep_params"""

    _fields_ = [
        ("grid", C.c_int32 * 3),
        ("supercell", C.c_int32 * 3),
        ("guard_supercells", C.c_int32 * 3),
        ("cell_size", C.c_float * 3),
        ("dt", C.c_float),
        ("c", C.c_float),
        ("eps0", C.c_float),
        ("mue0", C.c_float),
        ("base_mass", C.c_float),
        ("base_charge", C.c_float),
        ("shape", C.c_int32),
        ("pusher", C.c_int32),
        ("current_solver", C.c_int32),
        ("field_solver", C.c_int32),
        ("lehe_dir", C.c_int32),
        ("periodic", C.c_int32 * 3),
        ("devices", C.c_int32 * 3),
        ("rank_pos", C.c_int32 * 3),
        ("device", C.c_int32),
        ("flags", C.c_int32),
        ("current_interpolation", C.c_int32),
        ("absorber_kind", C.c_int32),
        ("absorber_cells", (C.c_int32 * 2) * 3),
        ("absorber_strength", (C.c_float * 2) * 3),
        ("moving_window", C.c_int32),
        ("laser_enabled", C.c_int32),
        ("laser_polarisation", C.c_int32),
        ("laser_offset_ymin", C.c_int32),
        ("laser_amplitude", C.c_float),
        ("laser_omega", C.c_float),
        ("laser_pulse_duration", C.c_float),
        ("laser_nofocus_constant", C.c_float),
        ("laser_ramp_init", C.c_float),
        ("laser_phase", C.c_float),
        ("laser_pol_dir", C.c_float * 3),
        ("laser_time_delay", C.c_float),
        ("pml_sigma_max", C.c_float * 3),
        ("pml_kappa_max", C.c_float * 3),
        ("pml_alpha_max", C.c_float * 3),
        ("pml_sigma_kappa_grading_order", C.c_float),
        ("pml_alpha_grading_order", C.c_float),
        ("laser_profile", C.c_int32),
        ("laser_position", (C.c_int32 * 2) * 3),
        ("laser_w0", C.c_float),
        ("laser_wave_length", C.c_float),
        ("laser_time_shift", C.c_float),
        ("laser_focus_position", C.c_float * 3),
        ("laser_focus_origin_center", C.c_int32 * 3),
        ("laser_tilt", C.c_float * 2),
        ("laser_n_modes", C.c_int32),
        ("laser_modes", C.c_float * 8),
        ("laser_mode_phases", C.c_float * 8),
        ("laser_w0_axis", C.c_float * 2),
        ("laser_profile_params", C.c_float * 16),
    ]


def lib_path(exact=False):
    # PICSTEP_LIB: another build of the production library (A/B timing of kernel variants on one box)
    if not exact and os.environ.get("PICSTEP_LIB"):
        return os.environ["PICSTEP_LIB"]
    return os.path.join(_HERE, "libpicstep_exact.so" if exact else "libpicstep.so")


def load(exact=False):
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    key = bool(exact)
    if key in _LIBS:
        return _LIBS[key]
    path = lib_path(exact)
    if not os.path.exists(path):
        raise PicstepError(
            "%s is missing: build it with `python -m picongpu_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % path
        )
    L = C.CDLL(path)
    vp, i32, i64, u32, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
    fp = C.POINTER(C.c_float)
    L.picstep_version.restype = C.c_char_p
    L.picstep_last_error.restype = C.c_char_p
    L.picstep_last_error.argtypes = [vp]
    L.picstep_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.picstep_destroy.argtypes = [vp]
    L.picstep_species_add.argtypes = [vp, C.c_char_p, f32, f32, i64, C.POINTER(i32)]
    L.picstep_species_set_policy.argtypes = [vp, i32, i32, i32, i32]
    for n in ("picstep_fields_upload", "picstep_fields_download", "picstep_fields_upload_soa", "picstep_fields_download_soa"):
        getattr(L, n).argtypes = [vp, i32, vp]
    L.picstep_particles_upload.argtypes = [vp, i32, i64, vp, vp, vp, vp]
    L.picstep_particles_count.argtypes = [vp, i32, C.POINTER(i64)]
    L.picstep_particles_download.argtypes = [vp, i32, i64, vp, vp, vp, vp, C.POINTER(i64)]
    L.picstep_supercell_counts.argtypes = [vp, i32, vp]
    L.picstep_init_khi.argtypes = [vp, C.POINTER(i32 * 3), f32, C.c_double, C.c_double, C.c_double, u32]
    L.picstep_init_thermal.argtypes = [vp, i32, i32, f32, C.c_double, C.c_double, u32]
    L.picstep_current_reset.argtypes = [vp]
    L.picstep_push.argtypes = [vp, i32, u32]
    L.picstep_migrate.argtypes = [vp, i32]
    L.picstep_field_update_before_current.argtypes = [vp, u32]
    L.picstep_deposit.argtypes = [vp, i32]
    L.picstep_add_current.argtypes = [vp]
    L.picstep_field_update_after_current.argtypes = [vp, u32]
    L.picstep_field_exchange.argtypes = [vp, i32]
    L.picstep_step.argtypes = [vp, u32, u32]
    L.picstep_step_host.argtypes = [vp, u32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.picstep_sync.argtypes = [vp]
    L.picstep_slide.argtypes = [vp, C.POINTER(i32)]
    L.picstep_reduce.argtypes = [vp, i32, i32, C.POINTER(C.c_double)]
    L.picstep_debug_gather.argtypes = [vp, i32, i64, vp]
    L.picstep_comm_unique_id.argtypes = [vp]
    L.picstep_comm_init.argtypes = [vp, vp, i32, i32]
    L.picstep_launch_count.argtypes = [vp, C.POINTER(i64)]
    L.picstep_stage_times.argtypes = [vp, i32, fp]
    L.picstep_overlap_times.argtypes = [vp, fp]
    L.picstep_stream.argtypes = [vp, C.POINTER(vp)]
    L.picstep_neighbor_ranks.argtypes = [C.POINTER(i32 * 3), C.POINTER(i32 * 3), i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.picstep_exchange_widths.argtypes = [i32, i32, i32, i32, i32, C.POINTER(i32 * 2)]
    _LIBS[key] = L
    return L


def to_c_params(p, device=0, flags=0):
    cp = Params()
    for d in range(3):
        cp.grid[d] = p.grid[d]
        cp.supercell[d] = p.supercell[d]
        cp.guard_supercells[d] = p.guard_supercells[d]
        cp.cell_size[d] = p.cell_size[d]
        cp.periodic[d] = p.periodic[d]
        cp.devices[d] = p.devices[d]
        cp.rank_pos[d] = p.rank_pos[d]
    cp.dt, cp.c, cp.eps0, cp.mue0 = p.dt, p.c, p.eps0, p.mue0
    cp.base_mass, cp.base_charge = p.base_mass, p.base_charge
    cp.shape, cp.pusher, cp.current_solver, cp.field_solver, cp.lehe_dir = (
        p.shape, p.pusher, p.current_solver, p.field_solver, p.lehe_dir)
    cp.device = device
    cp.flags = flags
    cp.current_interpolation = int(getattr(p, "current_interpolation", 0))
    cp.absorber_kind = int(getattr(p, "absorber_kind", 0))
    cp.moving_window = int(getattr(p, "moving_window", 0))
    pm = getattr(p, "pml", None)
    if pm:
        for d in range(3):
            cp.pml_sigma_max[d], cp.pml_kappa_max[d], cp.pml_alpha_max[d] = pm["sigma_max"][d], pm["kappa_max"][d], pm["alpha_max"][d]
        cp.pml_sigma_kappa_grading_order = pm["sigma_kappa_grading_order"]
        cp.pml_alpha_grading_order = pm["alpha_grading_order"]
    las = getattr(p, "laser", None)
    if las:
        cp.laser_enabled = 1
        cp.laser_polarisation = int(las["polarisation"])
        cp.laser_offset_ymin = int(las["offset_ymin"])
        cp.laser_amplitude, cp.laser_omega = las["amplitude"], las["omega"]
        cp.laser_pulse_duration, cp.laser_nofocus_constant = las["pulse_duration"], las["nofocus_constant"]
        cp.laser_ramp_init, cp.laser_phase, cp.laser_time_delay = las["ramp_init"], las["phase"], las["time_delay"]
        for d in range(3):
            cp.laser_pol_dir[d] = las["pol"][d]
        cp.laser_profile = int(las.get("profile", 0))
        if las.get("position") is not None:
            for d in range(3):
                cp.laser_position[d][0], cp.laser_position[d][1] = (int(v) for v in las["position"][d])
        cp.laser_w0, cp.laser_wave_length = las.get("w0", 0.0), las.get("wave_length", 0.0)
        cp.laser_time_shift = las.get("time_shift", 0.0)
        for d in range(3):
            cp.laser_focus_position[d] = las.get("focus_position", (0.0, 0.0, 0.0))[d]
            cp.laser_focus_origin_center[d] = int(las.get("focus_origin_center", (0, 0, 0))[d])
        cp.laser_tilt[0], cp.laser_tilt[1] = las.get("tilt", (0.0, 0.0))
        modes = las.get("modes", (1.0,))
        phases = las.get("mode_phases", (0.0,) * len(modes))
        if len(modes) > 8 or len(phases) != len(modes):
            raise ValueError("laser: at most 8 Laguerre modes, one phase each")
        cp.laser_n_modes = len(modes)
        for m in range(len(modes)):
            cp.laser_modes[m], cp.laser_mode_phases[m] = modes[m], phases[m]
        cp.laser_w0_axis[0], cp.laser_w0_axis[1] = las.get("w0_axis", (0.0, 0.0))
        for k, v in enumerate(las.get("profile_params", ())):
            cp.laser_profile_params[k] = v
    for d in range(3):
        for sd in range(2):
            cp.absorber_cells[d][sd] = int(getattr(p, "absorber_cells", ((0, 0),) * 3)[d][sd])
            cp.absorber_strength[d][sd] = float(getattr(p, "absorber_strength", ((0.0, 0.0),) * 3)[d][sd])
    return cp


def moving_window_info(global_cells, local_cells, cell_size, c_dt, move_point, step, exact=False):
    """MovingWindow::getCurrentSlideInfo: (slide during this step?, window offset in the first GPU after it)."""
    L = load(exact)
    L.picstep_moving_window_info.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    sl, off = C.c_int32(0), C.c_int32(0)
    rc = L.picstep_moving_window_info(global_cells, local_cells, cell_size, c_dt, move_point, step, C.byref(sl), C.byref(off))
    if rc:
        raise PicstepError("picstep_moving_window_info: invalid argument")
    return bool(sl.value), off.value


def window_neighbors(n_ranks, periodic, position, slides, exact=False):
    """Ranks of the lower / upper neighbour of the rank at `position` after `slides` slides of the moving window."""
    L = load(exact)
    lo, hi = C.c_int32(-1), C.c_int32(-1)
    if L.picstep_window_neighbors(n_ranks, int(periodic), position, slides, C.byref(lo), C.byref(hi)):
        raise PicstepError("picstep_window_neighbors: invalid argument")
    return lo.value, hi.value


def neighbor_ranks(devices, periodic, rank, axis, exact=False):
    L = load(exact)
    lo, hi = C.c_int32(-1), C.c_int32(-1)
    rc = L.picstep_neighbor_ranks((C.c_int32 * 3)(*devices), (C.c_int32 * 3)(*periodic), rank, axis, C.byref(lo), C.byref(hi))
    if rc:
        raise PicstepError("picstep_neighbor_ranks: invalid argument")
    return lo.value, hi.value


def exchange_widths(shape, field_solver, lehe_dir, field, axis, exact=False):
    L = load(exact)
    out = (C.c_int32 * 2)()
    rc = L.picstep_exchange_widths(shape, field_solver, lehe_dir, field, axis, C.byref(out))
    if rc:
        raise PicstepError("picstep_exchange_widths: invalid argument")
    return out[0], out[1]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Simulation:
    """Host driver around one picstep context (one GPU / one rank).

    Mirrors the reference's objects by name: `fieldE/B/J` accessors, species by name, and `run_one_step`
    issuing the stage calls in the order of Simulation::runOneStep.
    """

    def __init__(self, params, device=0, exact=False, atomic_deposit=False, cell_deposit=False, unfused=False, no_fdtd_tma=False):
        self.p = params
        self.L = load(exact)
        self.ctx = C.c_void_p()
        # picstep_params.flags: bit0 reference-strategy atomic deposit, bit1 warp-per-cell deposit kernel,
        # bit2 picstep_step without push+deposit fusion, bit4 Yee update with the one-thread-per-cell kernels (no TMA bricks)
        cp = to_c_params(params, device, (1 if atomic_deposit else 0) | (2 if cell_deposit else 0) | (4 if unfused else 0) | (16 if no_fdtd_tma else 0))
        rc = self.L.picstep_create(C.byref(cp), C.byref(self.ctx))
        if rc:
            raise PicstepError("picstep_create failed (%d): %s" % (rc, self.L.picstep_last_error(None).decode()))
        self.species = {}
        self.step_index = 0
        self.slides = 0  # number of moving-window slides so far
        N = params.padded
        self.field_shape = (3, N[2], N[1], N[0])
        for s in params.species:
            self.add_species(s.name, s.mass_ratio, s.charge_ratio)

    # -- plumbing ------------------------------------------------------------------------------------------------
    def _chk(self, rc, what):
        if rc:
            raise PicstepError("%s failed (%d): %s" % (what, rc, self.L.picstep_last_error(self.ctx).decode()))

    def close(self):
        if self.ctx:
            self.L.picstep_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- species / particles -------------------------------------------------------------------------------------
    def add_species(self, name, mass_ratio, charge_ratio, capacity=0):
        sid = C.c_int32(-1)
        self._chk(self.L.picstep_species_add(self.ctx, name.encode(), mass_ratio, charge_ratio, capacity, C.byref(sid)), "species_add")
        self.species[name] = sid.value
        return sid.value

    def set_policy(self, species, shape=-1, pusher=-1, current_solver=-1):
        """shape<> / particlePusher<> / current<> flags of one species (default: the values of the parameters)."""
        self._chk(self.L.picstep_species_set_policy(self.ctx, self._sid(species), shape, pusher, current_solver), "species_set_policy")

    def _sid(self, s):
        return self.species[s] if isinstance(s, str) else int(s)

    def upload_particles(self, species, pos, mom, w, cell):
        pos = np.ascontiguousarray(pos, np.float32)
        mom = np.ascontiguousarray(mom, np.float32)
        w = np.ascontiguousarray(w, np.float32)
        cell = np.ascontiguousarray(cell, np.int32)
        n = w.shape[0]
        assert pos.shape == (3, n) and mom.shape == (3, n) and cell.shape == (n,)
        self._chk(self.L.picstep_particles_upload(self.ctx, self._sid(species), n, _ptr(pos), _ptr(mom), _ptr(w), _ptr(cell)), "particles_upload")

    def particle_count(self, species):
        n = C.c_int64(0)
        self._chk(self.L.picstep_particles_count(self.ctx, self._sid(species), C.byref(n)), "particles_count")
        return n.value

    def download_particles(self, species, out=None):
        """Particles in frame-run order.  `out` = (pos[3,n], mom[3,n], w[n], cell[n]) pre-allocated C-contiguous
        arrays (e.g. views of pinned memory) to download into; allocated here otherwise."""
        n = self.particle_count(species)
        if out is None:
            pos = np.empty((3, n), np.float32)
            mom = np.empty((3, n), np.float32)
            w = np.empty(n, np.float32)
            cell = np.empty(n, np.int32)
        else:
            pos, mom, w, cell = out
            assert pos.shape == (3, n) and mom.shape == (3, n) and w.shape == (n,) and cell.shape == (n,)
            assert pos.dtype == np.float32 and mom.dtype == np.float32 and w.dtype == np.float32 and cell.dtype == np.int32
            assert all(a.flags["C_CONTIGUOUS"] for a in (pos, mom, w, cell))
        got = C.c_int64(0)
        self._chk(self.L.picstep_particles_download(self.ctx, self._sid(species), n, _ptr(pos), _ptr(mom), _ptr(w), _ptr(cell), C.byref(got)), "particles_download")
        return pos, mom, w, cell

    def supercell_counts(self, species):
        nsc = self.p.num_supercells
        out = np.zeros(nsc[0] * nsc[1] * nsc[2], np.int64)
        self._chk(self.L.picstep_supercell_counts(self.ctx, self._sid(species), _ptr(out)), "supercell_counts")
        return out.reshape(nsc[2], nsc[1], nsc[0])

    def init_khi(self, ppc_dim=(5, 5, 1), gamma=1.021, temperature_keV=0.0005, seed=42):
        p = self.p
        self._chk(self.L.picstep_init_khi(self.ctx, C.byref((C.c_int32 * 3)(*ppc_dim)), p.real_particles_per_cell, gamma, temperature_keV, p.ev_pic, seed), "init_khi")

    def init_thermal(self, species, ppc, temperature_keV=17.5 * 510.998950, seed=42):
        """Uniform warm plasma (Thermal benchmark: `temperature = 17.5 * 510.998950` keV, particle.param:77)."""
        p = self.p
        self._chk(self.L.picstep_init_thermal(self.ctx, self._sid(species), ppc, p.real_particles_per_cell, temperature_keV, p.ev_pic, seed), "init_thermal")

    # -- fields --------------------------------------------------------------------------------------------------
    def upload_field(self, field, soa):
        soa = np.ascontiguousarray(soa, np.float32)
        assert soa.shape == self.field_shape
        self._chk(self.L.picstep_fields_upload_soa(self.ctx, field, _ptr(soa)), "fields_upload_soa")

    def download_field(self, field):
        out = np.empty(self.field_shape, np.float32)
        self._chk(self.L.picstep_fields_download_soa(self.ctx, field, _ptr(out)), "fields_download_soa")
        return out

    def upload_field_aos(self, field, aos):
        aos = np.ascontiguousarray(aos, np.float32)
        assert aos.shape == self.field_shape[1:] + (3,)
        self._chk(self.L.picstep_fields_upload(self.ctx, field, _ptr(aos)), "fields_upload")

    def download_field_aos(self, field):
        out = np.empty(self.field_shape[1:] + (3,), np.float32)
        self._chk(self.L.picstep_fields_download(self.ctx, field, _ptr(out)), "fields_download")
        return out

    # -- stages (names follow the reference's stage functors) ---------------------------------------------------------
    def current_reset(self):
        self._chk(self.L.picstep_current_reset(self.ctx), "current_reset")

    def push(self, species):
        self._chk(self.L.picstep_push(self.ctx, self._sid(species), self.step_index), "push")

    def migrate(self, species):
        self._chk(self.L.picstep_migrate(self.ctx, self._sid(species)), "migrate")

    def field_update_before_current(self):
        self._chk(self.L.picstep_field_update_before_current(self.ctx, self.step_index), "field_update_before_current")

    def deposit(self, species):
        self._chk(self.L.picstep_deposit(self.ctx, self._sid(species)), "deposit")

    def add_current(self):
        self._chk(self.L.picstep_add_current(self.ctx), "add_current")

    def field_update_after_current(self):
        self._chk(self.L.picstep_field_update_after_current(self.ctx, self.step_index), "field_update_after_current")

    def field_exchange(self, field):
        self._chk(self.L.picstep_field_exchange(self.ctx, field), "field_exchange")

    def run_one_step(self):
        """Simulation::runOneStep: CurrentReset, ParticlePush (+migration), update_beforeCurrent, CurrentDeposition,
        CurrentInterpolationAndAdditionToEMF, update_afterCurrent."""
        self.current_reset()
        for s in self.species.values():
            self.push(s)
            self.migrate(s)
        self.field_update_before_current()
        for s in self.species.values():
            self.deposit(s)
        self.add_current()
        self.field_update_after_current()
        self.step_index += 1

    def step(self, n=1):
        """n steps through the fused C entry point picstep_step()."""
        self._chk(self.L.picstep_step(self.ctx, self.step_index, n), "step")
        self.step_index += n

    def slide(self):
        """GridController::slide + Simulation::slide: returns True when this rank became the (empty) top of the window."""
        r = C.c_int32(0)
        self._chk(self.L.picstep_slide(self.ctx, C.byref(r)), "slide")
        self.slides += 1
        n = self.p.devices[1]
        self.p.rank_pos = (self.p.rank_pos[0], (self.p.rank_pos[1] - 1 + n) % n, self.p.rank_pos[2])
        return bool(r.value)

    def step_host(self, E, B, species_arrays):
        """One step through HOST buffers (upload E,B + all species, step, download E,B + energies).
        species_arrays: list of (pos, mom, w, cell) in species order; E,B are updated in place."""
        ns = len(species_arrays)
        n = (C.c_int64 * ns)(*[a[2].shape[0] for a in species_arrays])
        mk = lambda k: (C.c_void_p * ns)(*[a[k].ctypes.data for a in species_arrays])
        pos, mom, w, cell = mk(0), mk(1), mk(2), mk(3)
        en = (C.c_double * 4)()
        self._chk(self.L.picstep_step_host(self.ctx, self.step_index, _ptr(E), _ptr(B), ns, n, pos, mom, w, cell, en), "step_host")
        self.step_index += 1
        return np.array(list(en))

    def sync(self):
        self._chk(self.L.picstep_sync(self.ctx), "sync")

    # -- diagnostics ---------------------------------------------------------------------------------------------
    def reduce(self, what, species=0):
        out = (C.c_double * 2)()
        self._chk(self.L.picstep_reduce(self.ctx, what, self._sid(species) if what not in (REDUCE_GAUSS, REDUCE_SLOW_PATH) else 0, out), "reduce")
        return np.array([out[0], out[1]])

    def field_energy(self):
        return self.reduce(REDUCE_FIELD_ENERGY)

    def particle_energy(self, species):
        return self.reduce(REDUCE_PARTICLE_ENERGY, species)

    def slow_path_counts(self):
        """(wide trajectories, PQS one-plane trajectories) deposited through global atomics since the last call."""
        r = self.reduce(REDUCE_SLOW_PATH)
        return int(r[0]), int(r[1])

    def gauss_residual(self):
        return float(self.reduce(REDUCE_GAUSS)[0])

    def debug_gather(self, species):
        n = self.particle_count(species)
        out = np.empty((6, n), np.float32)
        self._chk(self.L.picstep_debug_gather(self.ctx, self._sid(species), n, _ptr(out)), "debug_gather")
        return out[:3], out[3:]

    def launch_count(self):
        n = C.c_int64(0)
        self._chk(self.L.picstep_launch_count(self.ctx, C.byref(n)), "launch_count")
        return n.value

    def stage_times(self, enable=True):
        ms = (C.c_float * 7)()
        self._chk(self.L.picstep_stage_times(self.ctx, 1 if enable else 0, ms), "stage_times")
        return dict(zip(STAGES, list(ms)))

    def overlap_times(self):
        """(exchange ms, CORE ms, steps): device time per step from 'BORDER pushed' to 'exchange complete' / 'CORE complete'"""
        out = (C.c_float * 3)()
        self._chk(self.L.picstep_overlap_times(self.ctx, out), "overlap_times")
        return float(out[0]), float(out[1]), int(out[2])

    def stream(self):
        s = C.c_void_p()
        self._chk(self.L.picstep_stream(self.ctx, C.byref(s)), "stream")
        return s.value

    # -- multi GPU -----------------------------------------------------------------------------------------------
    def comm_init(self, unique_id, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._chk(self.L.picstep_comm_init(self.ctx, buf, rank, nranks), "comm_init")

    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.L.picstep_comm_unique_id(buf)
        if rc:
            raise PicstepError("picstep_comm_unique_id failed: %s" % self.L.picstep_last_error(None).decode())
        return bytes(buf)
