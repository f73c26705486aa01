"""Host-side mirror of the reference's compile-time `.param` / `.unitless` configuration for the PIC step.

The reference selects pusher / shape / current solver / field solver through type aliases in
`include/picongpu/param/{species,fieldSolver,memory,simulation}.param` and derives the PIC unit system in
`include/picongpu/unitless/simulation.unitless:420-480`.  Here the same names are runtime values that are
handed to the C ABI (`include/picstep.h`, `picstep_params`).
"""
from dataclasses import dataclass, field
import math

import numpy as np

# enums shared with include/picstep.h
SHAPE_NGP, SHAPE_CIC, SHAPE_TSC, SHAPE_PQS, SHAPE_PCS = range(5)
PUSHER_BORIS, PUSHER_VAY, PUSHER_HIGUERA_CARY = range(3)
CURRENT_ESIRKEPOV, CURRENT_EMZ = range(2)
SOLVER_YEE, SOLVER_LEHE = range(2)

SHAPE_NAMES = {"NGP": 0, "CIC": 1, "TSC": 2, "PQS": 3, "PCS": 4}
PUSHER_NAMES = {"Boris": 0, "Vay": 1, "HigueraCary": 2}
CURRENT_NAMES = {"Esirkepov": 0, "EmZ": 1, "EZ": 1}
SOLVER_NAMES = {"Yee": 0, "Lehe": 1}

# include/picongpu/param/physicalConstants.param:25-50
SPEED_OF_LIGHT_SI = 2.99792458e8
MUE0_SI = 1.25663706127e-6
ELECTRON_MASS_SI = 9.1093837139e-31
ELECTRON_CHARGE_SI = -1.602176634e-19
EV_SI = 1.602176634e-19


def _f32(x):
    return float(np.float32(x))


@dataclass
class Species:
    """A species definition: `Particles<Name, Flags, Attributes>` with massRatio<> / chargeRatio<> flags
    (include/picongpu/param/speciesAttributes.param:195-256)."""

    name: str
    mass_ratio: float
    charge_ratio: float


@dataclass
class SimParams:
    # --- memory.param / grid ---
    grid: tuple  # local cells per device (no guard), multiple of supercell
    supercell: tuple = (8, 8, 4)  # SuperCellSize, memory.param:51
    guard_supercells: tuple = (1, 1, 1)  # GuardSize, memory.param:73
    # --- simulation.param (SI) ---
    delta_t_si: float = 1.79e-16
    cell_si: tuple = (9.34635e-8, 9.34635e-8, 9.34635e-8)
    base_density_si: float = 1.0e25
    typical_ppc: int = 25
    # --- species.param / fieldSolver.param ---
    shape: int = SHAPE_TSC
    pusher: int = PUSHER_BORIS
    current_solver: int = CURRENT_ESIRKEPOV
    field_solver: int = SOLVER_YEE
    lehe_dir: int = 1
    # --- --currentInterpolation (0 none, 1 binomial); fieldAbsorber.param (0 none, 1 exponential; NUM_CELLS and
    #     exponential::STRENGTH per [axis][negative, positive], only used at non-periodic outer boundaries) ---
    current_interpolation: int = 0
    absorber_kind: int = 0  # 0 none, 1 exponential, 2 PML
    absorber_cells: tuple = ((12, 12), (12, 12), (12, 12))
    absorber_strength: tuple = ((1.0e-3, 1.0e-3), (1.0e-3, 1.0e-3), (1.0e-3, 1.0e-3))
    moving_window: int = 0  # -m: sliding window along y (needs a non-periodic y axis)
    # fieldAbsorber.param pml:: values (dict from pml_params(), PIC units); used with absorber_kind = 2
    pml: dict = None
    # incidentField.param: PlaneWave / GaussianPulse profile on YMin (dict from plane_wave_laser() / gaussian_pulse_laser(),
    # PIC units) or None = profiles::None
    laser: dict = None
    # --- runtime (-d, --periodic) ---
    periodic: tuple = (1, 1, 1)
    devices: tuple = (1, 1, 1)
    rank_pos: tuple = (0, 0, 0)
    species: list = field(default_factory=list)

    def __post_init__(self):
        for d in range(3):
            if self.grid[d] % self.supercell[d]:
                raise ValueError("grid must be a multiple of the supercell size (DomainAdjuster)")
            if self.grid[d] // self.supercell[d] < 2 and self.devices[d] > 1:
                raise ValueError("at least 2 supercells per split axis")
        # unit system: simulation.unitless:420-480 (all in float_64, cast to float_X on use)
        self.unit_time = self.delta_t_si
        self.unit_speed = SPEED_OF_LIGHT_SI
        self.unit_length = self.unit_time * self.unit_speed
        cell_vol_si = self.cell_si[0] * self.cell_si[1] * self.cell_si[2]
        self.typical_num_particles_per_macro = self.base_density_si * cell_vol_si / float(self.typical_ppc)
        self.unit_mass = ELECTRON_MASS_SI * self.typical_num_particles_per_macro
        self.unit_charge = -1.0 * ELECTRON_CHARGE_SI * self.typical_num_particles_per_macro
        self.unit_energy = self.unit_mass * self.unit_length**2 / self.unit_time**2
        self.unit_efield = 1.0 / (self.unit_time**2 / self.unit_mass / self.unit_length * self.unit_charge)
        self.unit_bfield = self.unit_mass / (self.unit_time * self.unit_charge)
        # PIC-unit float_X values (simulation.unitless:34-110)
        self.cell_size = tuple(_f32(c / self.unit_length) for c in self.cell_si)
        self.dt = _f32(self.delta_t_si / self.unit_time)
        self.c = _f32(SPEED_OF_LIGHT_SI / self.unit_speed)
        self.base_mass = _f32(ELECTRON_MASS_SI / self.unit_mass)
        self.base_charge = _f32(ELECTRON_CHARGE_SI / self.unit_charge)
        self.mue0 = _f32(MUE0_SI / self.unit_length / self.unit_mass * self.unit_charge * self.unit_charge)
        self.eps0 = _f32(1.0 / self.mue0 / self.c / self.c)
        self.ev_pic = EV_SI / self.unit_energy
        # densities
        self.real_particles_per_cell = _f32(
            _f32(self.base_density_si * self.unit_length**3)
            * _f32(np.float32(self.cell_size[0]) * np.float32(self.cell_size[1]) * np.float32(self.cell_size[2]))
        )
        self.wrap = tuple(1 if (self.periodic[d] and self.devices[d] == 1) else 0 for d in range(3))
        # open[d] = (lower, upper): this rank's face is a non-periodic outer boundary of the global domain
        self.open = tuple((int(not self.periodic[d] and self.rank_pos[d] == 0),
                           int(not self.periodic[d] and self.rank_pos[d] == self.devices[d] - 1)) for d in range(3))

    @property
    def guard_cells(self):
        return tuple(self.supercell[d] * self.guard_supercells[d] for d in range(3))

    @property
    def padded(self):
        g = self.guard_cells
        return tuple(self.grid[d] + 2 * g[d] for d in range(3))

    @property
    def num_supercells(self):
        return tuple(self.grid[d] // self.supercell[d] for d in range(3))

    @property
    def global_grid(self):
        return tuple(self.grid[d] * self.devices[d] for d in range(3))

    @property
    def global_offset(self):
        return tuple(self.grid[d] * self.rank_pos[d] for d in range(3))

    def cfl_ok(self):
        """Yee CFL: c*dt <= 1/sqrt(sum 1/dx^2) (include/picongpu/fields/MaxwellSolver/CFLChecker.hpp)."""
        s = sum(1.0 / (c * c) for c in self.cell_size)
        return self.c * self.dt <= 1.0 / math.sqrt(s)


def khi_params(grid=(64, 64, 64), **kw):
    """share/picongpu/examples/KelvinHelmholtz/include/picongpu/param/*: Boris, Yee, Esirkepov, TSC, 25 ppc."""
    p = SimParams(grid=tuple(grid), **kw)
    p.species = [Species("e", 1.0, 1.0), Species("i", 1836.152672, -1.0)]
    return p


def thermal_params(grid=(64, 64, 64), **kw):
    """share/picongpu/benchmarks/Thermal/include/picongpu/param/*: electrons only, uniform warm plasma."""
    kw.setdefault("delta_t_si", 9.65531e-14)
    kw.setdefault("cell_si", (5.78918e-5,) * 3)
    kw.setdefault("base_density_si", 1.0e20)
    p = SimParams(grid=tuple(grid), **kw)
    p.species = [Species("e", 1.0, 1.0)]
    return p


def plane_wave_laser(p, a0=1.0, wavelength_si=0.8e-6, pulse_duration_si=5.0e-15, nofocus_constant_si=0.0, ramp_init=8.0, phase=0.0,
                     polarisation="linear", pol_dir=(1.0, 0.0, 0.0), offset_ymin=16, time_delay_si=0.0):
    """incidentField.param for a `profiles::PlaneWave<>` entering through YMin (include/picongpu/fields/incidentField/
    profiles/PlaneWave.def:36-60, BaseParam.hpp): the SI parameters converted to the unitless values of
    `PlaneWaveUnitless` / `BaseParamUnitless` (float_X).  AMPLITUDE_SI = a0 * (-2 pi / wavelength * m_e c^2 / e) as in
    examples/LaserWakefield/include/picongpu/param/incidentField.param."""
    amplitude_si = a0 * (-2.0 * math.pi / wavelength_si * ELECTRON_MASS_SI * SPEED_OF_LIGHT_SI**2 / ELECTRON_CHARGE_SI)
    wave_length = _f32(wavelength_si / p.unit_length)
    f = _f32(np.float32(p.c) / np.float32(wave_length))
    return dict(
        polarisation=0 if polarisation == "linear" else 1,
        offset_ymin=int(offset_ymin),
        amplitude=_f32(amplitude_si / p.unit_efield),
        omega=_f32(np.float32(2.0 * math.pi) * np.float32(f)),
        pulse_duration=_f32(pulse_duration_si / p.unit_time),
        nofocus_constant=_f32(nofocus_constant_si / p.unit_time),
        ramp_init=_f32(ramp_init),
        phase=_f32(phase),
        pol=tuple(_f32(v) for v in pol_dir),
        time_delay=_f32(time_delay_si / p.unit_time),
    )


def gaussian_pulse_laser(p, a0=8.0, wavelength_si=0.8e-6, pulse_duration_si=5.0e-15, w0_si=5.0e-6 / 1.17741, pulse_init=15.0,
                         focus_position_si=(0.0, 4.62e-5, 0.0), focus_origin_center=(1, 0, 1), phase=0.0, polarisation="circular",
                         pol_dir=(1.0, 0.0, 0.0), position=((16, -16), (16, -16), (16, -16)), time_delay_si=0.0,
                         tilt_deg=(0.0, 0.0), modes=(1.0,), mode_phases=None):
    """incidentField.param for a `profiles::GaussianPulse<Params, GaussianPulseEnvelope<Params>>` (with a tilt:
    `PulseFrontTilt`) entering through YMin: the SI parameters converted as in GaussianPulseUnitless / BaseParamUnitless
    (profiles/GaussianPulse.hpp:93-110, BaseParam.hpp:43-180).  The defaults are the values of
    examples/LaserWakefield/include/picongpu/param/incidentField.param (a0 = 8, circular, focus 46.2 um behind the
    y boundary on the transversal centre of the box)."""
    amplitude_si = a0 * (-2.0 * math.pi / wavelength_si * ELECTRON_MASS_SI * SPEED_OF_LIGHT_SI**2 / ELECTRON_CHARGE_SI)
    wave_length = _f32(wavelength_si / p.unit_length)
    f = _f32(np.float32(p.c) / np.float32(wave_length))
    pulse_duration = _f32(pulse_duration_si / p.unit_time)
    pi_f = float(np.float32(math.pi))
    modes = tuple(_f32(m) for m in modes)
    return dict(
        profile=1,
        polarisation=0 if polarisation == "linear" else 1,
        offset_ymin=int(position[1][0]),
        position=tuple((int(a), int(b)) for a, b in position),
        amplitude=_f32(amplitude_si / p.unit_efield),
        omega=_f32(np.float32(2.0 * math.pi) * np.float32(f)),
        wave_length=wave_length,
        pulse_duration=pulse_duration,
        # GaussianPulseEnvelope::TIME_SHIFT = -0.5_X * PULSE_INIT (float_64) * PULSE_DURATION (float_X)
        time_shift=_f32(-0.5 * float(pulse_init) * pulse_duration),
        w0=_f32(w0_si / p.unit_length),
        focus_position=tuple(_f32(v / p.unit_length) for v in focus_position_si),
        focus_origin_center=tuple(int(v) for v in focus_origin_center),
        tilt=tuple(_f32(t * pi_f / 180.0) for t in tilt_deg),
        modes=modes,
        mode_phases=tuple(_f32(v) for v in (mode_phases if mode_phases is not None else (0.0,) * len(modes))),
        nofocus_constant=0.0,
        ramp_init=0.0,
        phase=_f32(phase),
        pol=tuple(_f32(v) for v in pol_dir),
        time_delay=_f32(time_delay_si / p.unit_time),
    )


def _separable_base(p, profile, a0, amplitude_si, wavelength_si, pulse_duration_si, w0_axis_si, focus_position_si, focus_origin_center, phase,
                    polarisation, pol_dir, position, time_delay_si, nofocus_constant_si):
    """BaseTransversalGaussianParamUnitless (profiles/BaseParam.hpp:43-203) of the separable profiles"""
    if amplitude_si is None:
        amplitude_si = a0 * (-2.0 * math.pi / wavelength_si * ELECTRON_MASS_SI * SPEED_OF_LIGHT_SI**2 / ELECTRON_CHARGE_SI)
    wave_length = _f32(wavelength_si / p.unit_length)
    f = _f32(np.float32(p.c) / np.float32(wave_length))
    return dict(
        profile=profile, polarisation=0 if polarisation == "linear" else 1, offset_ymin=int(position[1][0]),
        position=tuple((int(a), int(b)) for a, b in position), amplitude=_f32(amplitude_si / p.unit_efield),
        omega=_f32(np.float32(2.0 * math.pi) * np.float32(f)), wave_length=wave_length, pulse_duration=_f32(pulse_duration_si / p.unit_time),
        w0_axis=tuple(_f32(w / p.unit_length) for w in w0_axis_si), focus_position=tuple(_f32(v / p.unit_length) for v in focus_position_si),
        focus_origin_center=tuple(int(v) for v in focus_origin_center), nofocus_constant=_f32(nofocus_constant_si / p.unit_time), ramp_init=0.0,
        phase=_f32(phase), pol=tuple(_f32(v) for v in pol_dir), time_delay=_f32(time_delay_si / p.unit_time), profile_params=())


_SEPARABLE_DEFAULTS = dict(a0=1.0, amplitude_si=None, wavelength_si=0.8e-6, pulse_duration_si=5.0e-15, w0_axis_si=(4.246e-6, 4.246e-6),
                           focus_position_si=(0.0, 0.0, 0.0), focus_origin_center=(1, 0, 1), phase=0.0, polarisation="linear", pol_dir=(1.0, 0.0, 0.0),
                           position=((16, -16), (16, -16), (16, -16)), time_delay_si=0.0, nofocus_constant_si=0.0)


def wavepacket_laser(p, pulse_init=20.0, **kw):
    """`profiles::Wavepacket<>` on YMin (profiles/Wavepacket.def:36-72, Wavepacket.hpp:52-73): Gaussian in time with an
    optional plateau and Gaussian transversally.  INIT_TIME = PULSE_INIT (float_64) * PULSE_DURATION + LASER_NOFOCUS_CONSTANT."""
    a = dict(_SEPARABLE_DEFAULTS, nofocus_constant_si=7.0 * 5.0e-15)
    a.update(kw)
    las = _separable_base(p, 2, **a)
    las["profile_params"] = (_f32(float(pulse_init) * las["pulse_duration"] + las["nofocus_constant"]),)
    return las


def polynom_laser(p, **kw):
    """`profiles::Polynom<>` on YMin (profiles/Polynom.def:36-52, Polynom.hpp:112-136): the amplitude rises for half of
    PULSE_DURATION with a polynomial of fifth order and falls symmetrically."""
    a = dict(_SEPARABLE_DEFAULTS)
    a.update(kw)
    return _separable_base(p, 3, **a)


def exp_ramp_with_prepulse_laser(p, int_ratio_prepulse=0.0, int_ratio_points=(1.0e-8, 1.0e-4, 1.0e-4), time_prepulse_si=-950.0e-15,
                                 time_peakpulse_si=0.0, time_points_si=(-1000.0e-15, -300.0e-15, -100.0e-15), prepulse_duration_si=None,
                                 ramp_init=16.0, **kw):
    """`profiles::ExpRampWithPrepulse<>` on YMin (profiles/ExpRampWithPrepulse.def:36-110, ExpRampWithPrepulse.hpp:52-124):
    Gaussian main pulse with plateau, preceded by two exponential ramps through three (time, intensity ratio) points and
    an optional Gaussian prepulse.  time_start_init = TIME_POINT_1 - 0.5 * RAMP_INIT (float_64) * PULSE_DURATION."""
    a = dict(_SEPARABLE_DEFAULTS, amplitude_si=1.0e6 if "a0" not in kw else None)
    a.update(kw)
    las = _separable_base(p, 4, **a)
    t = [_f32(v / p.unit_time) for v in time_points_si]
    pre = _f32((prepulse_duration_si if prepulse_duration_si is not None else a["pulse_duration_si"]) / p.unit_time)
    end_upramp = float(np.float32(_f32(time_peakpulse_si / p.unit_time)) - np.float32(0.5) * np.float32(las["nofocus_constant"]))
    if not (t[0] < t[1] < t[2] < end_upramp):
        raise ValueError("TIME_POINT_1/2/3 and the beginning of the plateau should be in ascending order")
    las["profile_params"] = (_f32(t[0] - 0.5 * float(ramp_init) * las["pulse_duration"]), _f32(time_prepulse_si / p.unit_time),
                             _f32(time_peakpulse_si / p.unit_time), t[0], t[1], t[2], pre, _f32(int_ratio_prepulse)) + tuple(_f32(v) for v in int_ratio_points)
    return las


def pml_params(p, sigma_kappa_grading_order=4.0, sigma_opt_multiplier=1.0, kappa_max=(1.0, 1.0, 1.0), alpha_grading_order=1.0,
               alpha_max_si=(0.2, 0.2, 0.2)):
    """include/picongpu/param/fieldAbsorber.param:98-158 (pml:: defaults) converted as in
    unitless/fieldAbsorber.unitless:72-104: SIGMA_OPT_SI = 0.8 (order + 1) / (Z0 * cell size), normalised by eps0 and
    the unit of time."""
    eps0_si = 1.0 / (MUE0_SI * SPEED_OF_LIGHT_SI**2)
    z0_si = MUE0_SI * SPEED_OF_LIGHT_SI
    sigma_max_si = [0.8 * (sigma_kappa_grading_order + 1.0) / (z0_si * c) * sigma_opt_multiplier for c in p.cell_si]
    return dict(
        sigma_max=tuple(_f32(s / eps0_si * p.unit_time) for s in sigma_max_si),
        kappa_max=tuple(_f32(k) for k in kappa_max),
        alpha_max=tuple(_f32(a / eps0_si * p.unit_time) for a in alpha_max_si),
        sigma_kappa_grading_order=_f32(sigma_kappa_grading_order),
        alpha_grading_order=_f32(alpha_grading_order),
    )
