/* picstep.h — C ABI of libpicstep.so: the B200-native (sm_100a) implementation of PIConGPU's core PIC step.
 *
 * The reference has no runtime plugin ABI for this path: pusher / shape / current solver / field solver are
 * compile-time policy types selected in `.param` files and driven by `Simulation::runOneStep`
 * (include/picongpu/simulation/control/Simulation.hpp:522-542).  This header turns that surface into plain C:
 * every entry point names the reference interface it replaces.  All functions return a status code (0 = OK),
 * never throw, take an opaque context bound to ONE CUDA device (+ one NCCL rank), are stream ordered and are not
 * thread safe per context (same convention as the reference: one host thread per rank).
 * The caller owns every host array it passes; the library owns all device memory.
 *
 * There is NO CPU fallback: picstep_create() fails with PICSTEP_ERR_NOGPU when no CUDA device is present.
 */
#ifndef PICSTEP_H
#define PICSTEP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    typedef struct picstep_ctx picstep_ctx;

    enum picstep_status
    {
        PICSTEP_OK = 0,
        PICSTEP_ERR_INVALID = 1, /* bad argument / unsupported configuration */
        PICSTEP_ERR_CUDA = 2, /* CUDA runtime error, see picstep_last_error() */
        PICSTEP_ERR_CAPACITY = 3, /* particle / exchange buffer too small */
        PICSTEP_ERR_COMM = 4, /* NCCL error or communicator missing */
        PICSTEP_ERR_NOGPU = 5 /* no CUDA device: the product path has no CPU fallback */
    };

    /* particles::shapes::{NGP,CIC,TSC,PQS,PCS} (include/picongpu/particles/shapes/) */
    enum picstep_shape
    {
        PICSTEP_SHAPE_NGP = 0,
        PICSTEP_SHAPE_CIC = 1,
        PICSTEP_SHAPE_TSC = 2,
        PICSTEP_SHAPE_PQS = 3,
        PICSTEP_SHAPE_PCS = 4
    };
    /* particles::pusher::{Boris,Vay,HigueraCary} (include/picongpu/unitless/pusher.unitless:43-86) */
    enum picstep_pusher
    {
        PICSTEP_PUSHER_BORIS = 0,
        PICSTEP_PUSHER_VAY = 1,
        PICSTEP_PUSHER_HIGUERA_CARY = 2
    };
    /* currentSolver::{Esirkepov,EmZ} (include/picongpu/fields/currentDeposition/) */
    enum picstep_current_solver
    {
        PICSTEP_CURRENT_ESIRKEPOV = 0,
        PICSTEP_CURRENT_EMZ = 1
    };
    /* fields::maxwellSolver::{Yee,Lehe<dir>} (include/picongpu/fields/MaxwellSolver/) */
    enum picstep_field_solver
    {
        PICSTEP_SOLVER_YEE = 0,
        PICSTEP_SOLVER_LEHE = 1
    };
    /* fields::currentInterpolation::{None,Binomial} (include/picongpu/fields/currentInterpolation/) */
    enum picstep_current_interpolation
    {
        PICSTEP_CURRENT_INTERPOLATION_NONE = 0,
        PICSTEP_CURRENT_INTERPOLATION_BINOMIAL = 1
    };
    /* incidentField.param profile on YMin (include/picongpu/fields/incidentField/profiles/profiles.def) */
    enum picstep_laser_profile
    {
        PICSTEP_LASER_PLANE_WAVE = 0, /* profiles::PlaneWave (profiles/PlaneWave.hpp) */
        PICSTEP_LASER_GAUSSIAN_PULSE = 1, /* profiles::GaussianPulse / PulseFrontTilt (profiles/GaussianPulse.hpp) */
        PICSTEP_LASER_WAVEPACKET = 2, /* profiles::Wavepacket (profiles/Wavepacket.hpp) */
        PICSTEP_LASER_POLYNOM = 3, /* profiles::Polynom (profiles/Polynom.hpp) */
        PICSTEP_LASER_EXP_RAMP_WITH_PREPULSE = 4 /* profiles::ExpRampWithPrepulse (profiles/ExpRampWithPrepulse.hpp) */
    };
    /* fields::absorber::Absorber::Kind (include/picongpu/fields/absorber/Absorber.hpp) */
    enum picstep_absorber
    {
        PICSTEP_ABSORBER_NONE = 0,
        PICSTEP_ABSORBER_EXPONENTIAL = 1,
        PICSTEP_ABSORBER_PML = 2
    };
    /* FieldE / FieldB / FieldJ, as named through DataConnector ("E","B","J") */
    enum picstep_field
    {
        PICSTEP_FIELD_E = 0,
        PICSTEP_FIELD_B = 1,
        PICSTEP_FIELD_J = 2
    };
    /* picstep_reduce() selectors */
    enum picstep_reduction
    {
        PICSTEP_REDUCE_FIELD_ENERGY = 0, /* out[0]=B energy, out[1]=E energy   (plugins/EnergyFields.x.cpp:198-233) */
        PICSTEP_REDUCE_PARTICLE_ENERGY = 1, /* out[0]=kinetic, out[1]=total, needs `species` (EnergyParticles.x.cpp:100-131) */
        PICSTEP_REDUCE_GAUSS = 2, /* out[0]=max|eps0 div E - rho|*V     (plugins/ChargeConservation.tpp:181-259) */
        PICSTEP_REDUCE_PARTICLE_COUNT = 3, /* out[0]=number of macro particles of `species` */
        /* out[0]=trajectories the fused kernel deposited through its global-atomic path since the last call (too wide for
         * the four-node window), out[1]=PQS trajectories that sent one plane of nodes that way; resets the counters */
        PICSTEP_REDUCE_SLOW_PATH = 4
    };

    /* The `.param` surface of the hot path as runtime values (all physical values in PIC units, i.e. what
     * `sim.pic.get*()` returns, include/picongpu/unitless/simulation.unitless:34-110). */
    typedef struct picstep_params
    {
        int32_t grid[3]; /* local cells without guard (-g / gridDist), multiple of supercell */
        int32_t supercell[3]; /* SuperCellSize  (param/memory.param:51); only 8x8x4 is compiled in */
        int32_t guard_supercells[3]; /* GuardSize      (param/memory.param:73); only 1x1x1 */
        float cell_size[3]; /* sim.pic.getCellSize() */
        float dt; /* sim.pic.getDt() */
        float c; /* sim.pic.getSpeedOfLight() */
        float eps0; /* sim.pic.getEps0() */
        float mue0; /* sim.pic.getMue0() */
        float base_mass; /* sim.pic.getBaseMass() */
        float base_charge; /* sim.pic.getBaseCharge() */
        int32_t shape; /* picstep_shape            (species.param: UsedParticleShape) */
        int32_t pusher; /* picstep_pusher           (species.param: UsedParticlePusher) */
        int32_t current_solver; /* picstep_current_solver   (species.param: UsedParticleCurrentSolver) */
        int32_t field_solver; /* picstep_field_solver     (fieldSolver.param:57) */
        int32_t lehe_dir; /* Cherenkov free direction of Lehe */
        int32_t periodic[3]; /* --periodic */
        int32_t devices[3]; /* -d : ranks per axis (one axis may be > 1) */
        int32_t rank_pos[3]; /* this rank's coordinate in the device grid */
        int32_t device; /* CUDA device ordinal */
        int32_t flags; /* bit0: deposit with the reference-strategy per-particle atomic kernel (cross-check);
                        * bit1: deposit with the warp-per-cell kernel (all shapes, EmZ) instead of the run kernel;
                        * bit2: picstep_step() runs push and deposit as separate kernels (no fusion);
                        * bit3: picstep_step() keeps the re-sort on the main stream (no overlap with the next kernel);
                        * bit4: Yee update with the one-thread-per-cell kernels instead of the TMA-staged bricks */
        /* --currentInterpolation none|binomial (simulation/stage/CurrentInterpolationAndAdditionToEMF.hpp:60-92,
         * fields/currentInterpolation/Binomial.hpp:41-112) */
        int32_t current_interpolation; /* picstep_current_interpolation */
        /* field absorber at non-periodic outer boundaries (param/fieldAbsorber.param:44-72,
         * fields/absorber/exponential/Exponential.kernel:45-118): thickness in cells and strength per
         * [axis][0 = negative side, 1 = positive side]; particles crossing such a boundary are absorbed
         * (particles/boundary/Absorbing.hpp:50-92, offset 0) */
        int32_t absorber_kind; /* picstep_absorber */
        int32_t absorber_cells[3][2];
        float absorber_strength[3][2];
        /* -m / --moving: the sliding window is active (simulation/control/MovingWindow.hpp): y must be non-periodic,
         * the +y absorber is switched off (Exponential.hpp:97-101) and picstep_slide() may be called */
        int32_t moving_window;
        /* incidentField.param (fields/incidentField/Solver.hpp:547-575, called from FDTDBase.hpp:108-117,161-166): a
         * `profiles::PlaneWave<>` (profiles/PlaneWave.hpp) or GaussianPulse (see laser_profile below) entering through the
         * YMin Huygens surface, Yee solver.  All values are the profile's unitless parameters
         * (PlaneWaveUnitless / BaseParamUnitless, PIC units).  The source is switched off once the moving window has
         * slid (Solver.hpp: "After the sliding window started moving, does nothing for the y boundaries"). */
        int32_t laser_enabled; /* 0: profiles::None on all boundaries */
        int32_t laser_polarisation; /* PolarisationType: 0 Linear, 1 Circular */
        int32_t laser_offset_ymin; /* POSITION[1][0]: cells between the global y-min boundary and the surface */
        float laser_amplitude; /* AMPLITUDE */
        float laser_omega; /* w = 2 pi c / WAVE_LENGTH */
        float laser_pulse_duration; /* PULSE_DURATION */
        float laser_nofocus_constant; /* LASER_NOFOCUS_CONSTANT (plateau) */
        float laser_ramp_init; /* RAMP_INIT */
        float laser_phase; /* LASER_PHASE */
        float laser_pol_dir[3]; /* POLARISATION_DIRECTION (unit, orthogonal to y) */
        float laser_time_delay; /* TIME_DELAY */
        /* absorber_kind = PICSTEP_ABSORBER_PML: convolutional PML (fields/absorber/pml/Pml.kernel, hook FDTDBase.hpp:244-298),
         * thickness = absorber_cells; the pml:: values of param/fieldAbsorber.param:98-158 as normalised in
         * unitless/fieldAbsorber.unitless:72-104 */
        float pml_sigma_max[3]; /* NORMALIZED_SIGMA_MAX */
        float pml_kappa_max[3]; /* KAPPA_MAX */
        float pml_alpha_max[3]; /* NORMALIZED_ALPHA_MAX */
        float pml_sigma_kappa_grading_order; /* SIGMA_KAPPA_GRADING_ORDER */
        float pml_alpha_grading_order; /* ALPHA_GRADING_ORDER */
        /* incidentField.param, continued: profile kind and the values of a `profiles::GaussianPulse<Params,
         * GaussianPulseEnvelope<Params>>` (profiles/GaussianPulse.hpp:93-346; with a non-zero tilt it is
         * `profiles::PulseFrontTilt`) entering through YMin.  The Huygens surface of this profile ends at POSITION on the
         * transversal axes (Solver.hpp:209-258), which may be periodic or not; the PlaneWave profile spans a periodic
         * transversal axis completely and ends at POSITION on a non-periodic one. */
        int32_t laser_profile; /* picstep_laser_profile */
        int32_t laser_position[3][2]; /* POSITION[axis][min, max]; max <= 0 counts from the upper boundary;
                                         [1][0] == laser_offset_ymin.  All zero: {offset, -offset} on every axis */
        float laser_w0; /* W0 */
        float laser_wave_length; /* WAVE_LENGTH */
        float laser_time_shift; /* GaussianPulseEnvelope::TIME_SHIFT = -0.5 * PULSE_INIT * PULSE_DURATION */
        float laser_focus_position[3]; /* FOCUS_POSITION_{X,Y,Z} */
        int32_t laser_focus_origin_center[3]; /* FOCUS_ORIGIN_* == Origin::Center (Functors.hpp:267-291) */
        float laser_tilt[2]; /* TILT_AXIS_1, TILT_AXIS_2 in radian */
        int32_t laser_n_modes; /* laguerreModes.size(), 1..8 (0 = one mode of weight 1) */
        float laser_modes[8]; /* laguerreModes */
        float laser_mode_phases[8]; /* laguerrePhases */
        /* separable profiles with a Gaussian transversal envelope (Wavepacket, Polynom, ExpRampWithPrepulse:
         * BaseTransversalGaussianParamUnitless, profiles/BaseParam.hpp:186-203; Functors.hpp:481-533) */
        float laser_w0_axis[2]; /* W0_AXIS_1, W0_AXIS_2 */
        /* Wavepacket: [0] INIT_TIME.  ExpRampWithPrepulse: [0] time_start_init, [1] TIME_PREPULSE, [2] TIME_PEAKPULSE,
         * [3..5] TIME_1..3, [6] PREPULSE_DURATION, [7] INT_RATIO_PREPULSE, [8..10] INT_RATIO_POINT_1..3 */
        float laser_profile_params[16];
    } picstep_params;

    /* library / build information: returns e.g. "picstep sm_100a fmad=on" */
    const char* picstep_version(void);
    /* last error text of a context (ctx may be NULL: error of the last failed picstep_create) */
    const char* picstep_last_error(const picstep_ctx* ctx);

    /* Simulation::init (Simulation.hpp:287-434): allocate E,B,J with guards and the kernel work buffers */
    int picstep_create(const picstep_params* params, picstep_ctx** out);
    int picstep_destroy(picstep_ctx* ctx);

    /* `Particles<Name, Flags, Attributes>` with massRatio<> / chargeRatio<> (speciesDefinition.param).
     * `capacity` = number of particle slots to reserve per buffer (0: sized at the first upload * 1.25). */
    int picstep_species_add(picstep_ctx* ctx, const char* name, float mass_ratio, float charge_ratio, int64_t capacity, int32_t* species_id);
    /* Per-species policy flags `shape<>`, `particlePusher<>`, `current<>` of the species definition
     * (include/picongpu/param/speciesAttributes.param:195-256; examples/KelvinHelmholtz/include/picongpu/param/
     * speciesDefinition.param:64-70).  Default: the values of picstep_params; a negative argument keeps the current
     * value.  Shapes whose lower interpolation margins agree modulo four cells can be mixed (NGP | CIC, TSC | PQS, PCS). */
    int picstep_species_set_policy(picstep_ctx* ctx, int32_t species, int32_t shape, int32_t pusher, int32_t current_solver);

    /* Field transfer in the reference's own layout: AoS float3, x fastest, guards included
     * (dims = grid + 2*supercell*guard_supercells; include/pmacc/memory/buffers/DeviceBuffer.hpp:111-119). */
    int picstep_fields_upload(picstep_ctx* ctx, int32_t field, const float* aos_with_guards);
    int picstep_fields_download(picstep_ctx* ctx, int32_t field, float* aos_with_guards);
    /* Same, component-major SoA [3][Nz][Ny][Nx] (the library's internal layout, no transposition). */
    int picstep_fields_upload_soa(picstep_ctx* ctx, int32_t field, const float* soa_with_guards);
    int picstep_fields_download_soa(picstep_ctx* ctx, int32_t field, float* soa_with_guards);

    /* Particle transfer as flat SoA arrays: pos[3][n] (in-cell, [0,1)), mom[3][n], weighting[n] and the local cell
     * index cell[n] = cx + grid.x*(cy + grid.y*cz).  Upload replaces the species' content and builds the
     * supercell-resident frame runs.  Download returns particles in frame-run order (supercell-major, cell-minor). */
    int picstep_particles_upload(picstep_ctx* ctx, int32_t species, int64_t n, const float* pos, const float* mom, const float* weighting, const int32_t* cell);
    int picstep_particles_count(picstep_ctx* ctx, int32_t species, int64_t* n);
    int picstep_particles_download(picstep_ctx* ctx, int32_t species, int64_t capacity, float* pos, float* mom, float* weighting, int32_t* cell, int64_t* n);
    /* SuperCell bookkeeping (include/pmacc/particles/memory/dataTypes/SuperCell.hpp:32-118): per supercell
     * (x fastest, no guard) the number of particles; frames = ceil(n/256), last frame size = ((n-1)%256)+1. */
    int picstep_supercell_counts(picstep_ctx* ctx, int32_t species, int64_t* counts);

    /* Synthetic KelvinHelmholtz initial condition generated on the device (bench input; same recipe and Philox
     * stream as the oracle's orc_khi_init): species 0 = electrons, 1 = ions; ppc = ppc_dim[0]*[1]*[2] each. */
    int picstep_init_khi(picstep_ctx* ctx, const int32_t* ppc_dim, float real_particles_per_cell, double gamma_drift, double temperature_keV, double ev_pic, uint32_t seed);
    /* Synthetic uniform warm plasma of one species generated on the device (bench input for the Thermal benchmark,
     * share/picongpu/benchmarks/Thermal/include/picongpu/param/{particle,speciesInitialization}.param: `ppc` particles
     * per cell at random in-cell positions, Maxwellian momenta of the given temperature, no drift). */
    int picstep_init_thermal(picstep_ctx* ctx, int32_t species, int32_t ppc, float real_particles_per_cell, double temperature_keV, double ev_pic, uint32_t seed);

    /* ---- stage calls, in the order of Simulation::runOneStep (Simulation.hpp:526-541) ---- */
    /* CurrentReset              (simulation/stage/CurrentReset.hpp:45-52) */
    int picstep_current_reset(picstep_ctx* ctx);
    /* ParticlePush: species->update(step): gather + push + move, marks leavers
     *                           (particles/Particles.tpp:322-368, Particles.kernel:170-316) */
    int picstep_push(picstep_ctx* ctx, int32_t species, uint32_t step);
    /* shiftBetweenSupercells + asyncCommunication(species): re-sort across supercells and migrate between ranks
     *                           (pmacc/particles/ParticlesBase.hpp:217-237, AsyncCommunicationImpl.hpp:46-60) */
    int picstep_migrate(picstep_ctx* ctx, int32_t species);
    /* fields::Solver::update_beforeCurrent (fields/MaxwellSolver/FDTD/FDTDBase.hpp:97-121) */
    int picstep_field_update_before_current(picstep_ctx* ctx, uint32_t step);
    /* CurrentDeposition         (simulation/stage/CurrentDeposition.x.cpp:105-116, fields/FieldJ.kernel:52-142) */
    int picstep_deposit(picstep_ctx* ctx, int32_t species);
    /* CurrentInterpolationAndAdditionToEMF: J guard reduction + E += -dt/eps0 * J
     *                           (simulation/stage/CurrentInterpolationAndAdditionToEMF.hpp:99-148) */
    int picstep_add_current(picstep_ctx* ctx);
    /* fields::Solver::update_afterCurrent (FDTDBase.hpp:151-183) */
    int picstep_field_update_after_current(picstep_ctx* ctx, uint32_t step);
    /* guard exchange of one field: E/B guards := neighbour border (GridBuffer::asyncCommunication,
     * pmacc/memory/buffers/GridBuffer.hpp:472-483); for J: border += neighbour guard (FieldJ.x.cpp:156-174) */
    int picstep_field_exchange(picstep_ctx* ctx, int32_t field);

    /* Moving window.  picstep_moving_window_info restates MovingWindow::getCurrentSlideInfo
     * (simulation/control/MovingWindow.hpp:44-170): for the step about to be computed, does the window slide
     * (*do_slide) and where does it start inside the first GPU afterwards (*offset_first_gpu, in cells).
     * global_cells / local_cells are the y extents, c_dt = speed of light * dt, cell_size in PIC units, move_point as
     * --windowMovePoint.  Host-only, no context needed. */
    int picstep_moving_window_info(int32_t global_cells, int32_t local_cells, double cell_size, double c_dt, double move_point, uint32_t step, int32_t* do_slide, int32_t* offset_first_gpu);
    /* GridController::slide + Simulation::slide (pmacc/mappings/simulation/GridController.hpp:166-176,
     * pmacc/communication/CommunicatorMPI.cpp:156-170, simulation/control/Simulation.hpp:581-593): all ranks call it
     * together; every rank moves one position down in y, neighbours / open faces / absorber follow, and the rank
     * that becomes the top of the window is reset (fields zero, no particles; *was_reset = 1) so that the caller can
     * initialise the new slab with picstep_particles_upload. */
    int picstep_slide(picstep_ctx* ctx, int32_t* was_reset);

    /* runOneStep x n : all stages, all species, device resident */
    int picstep_step(picstep_ctx* ctx, uint32_t first_step, uint32_t n);
    /* One step through HOST buffers (used for the end-to-end measurement): uploads E,B (SoA) and every species
     * from the given host arrays, runs one step, downloads E,B and the field/particle energies.
     * species arrays: pos/mom/weighting/cell pointers per species, n[s] particles each. */
    int picstep_step_host(picstep_ctx* ctx, uint32_t step, float* E_soa, float* B_soa, int32_t n_species, const int64_t* n, const float* const* pos, const float* const* mom, const float* const* weighting, const int32_t* const* cell, double* energies4);
    /* wait for all queued device work of this context */
    int picstep_sync(picstep_ctx* ctx);

    /* plugins' reductions used as parity observables; see picstep_reduction */
    int picstep_reduce(picstep_ctx* ctx, int32_t what, int32_t species, double* out);

    /* parity hook (no reference counterpart as a call, the arithmetic is FieldToParticleInterpolation.hpp:97-124):
     * E.x,E.y,E.z,B.x,B.y,B.z interpolated to every particle, out[6][capacity], frame-run order */
    int picstep_debug_gather(picstep_ctx* ctx, int32_t species, int64_t capacity, float* out);

    /* ---- multi GPU (replaces pmacc::CommunicatorMPI, include/pmacc/communication/CommunicatorMPI.cpp:70-149) ---- */
    /* 128-byte NCCL unique id, created on rank 0 and broadcast by the caller (torch.distributed / MPI / file) */
    int picstep_comm_unique_id(void* id128);
    int picstep_comm_init(picstep_ctx* ctx, const void* id128, int32_t rank, int32_t nranks);

    /* ---- measurement helpers ---- */
    /* number of kernels this context launched since creation (bench.py's gpu_launches) */
    int picstep_launch_count(picstep_ctx* ctx, int64_t* n);
    /* accumulated device time per stage in ms since the last call with reset!=0; names via picstep_stage_name.
     * stages: 0 current_reset 1 push 2 migrate 3 field_before 4 deposit 5 add_current 6 field_after */
    int picstep_stage_times(picstep_ctx* ctx, int32_t enable, float* ms7);
    /* Overlap evidence of the decomposed step (CORE / BORDER areas): out3[0] = mean device time per step from "BORDER area
     * pushed" until the exchange of leaving particles and J guard strips is complete (second stream), out3[1] = the same
     * until the last CORE kernel is complete (compute stream), out3[2] = number of steps measured since the last call.
     * Recorded while picstep_stage_times is enabled.  The exchange is hidden when out3[0] < out3[1]. */
    int picstep_overlap_times(picstep_ctx* ctx, float* out3);
    /* the CUDA stream (cudaStream_t) all work of the context is queued on */
    int picstep_stream(picstep_ctx* ctx, void** stream);

    /* ---- host-side pure functions (no GPU needed): domain decomposition used by the exchange ---- */
    /* neighbour ranks of `rank` along `axis` in a devices[3] grid, x fastest (CommunicatorMPI.cpp:70-111);
     * -1 where the boundary is not periodic */
    int picstep_neighbor_ranks(const int32_t* devices, const int32_t* periodic, int32_t rank, int32_t axis, int32_t* lower, int32_t* upper);
    /* the same for the split axis of a sliding window: ranks keep their identity (NCCL / MPI rank) while their
     * positions rotate, after `slides` calls of picstep_slide the rank at position q is (q + slides) mod n_ranks
     * (CommunicatorMPI::slide, pmacc/communication/CommunicatorMPI.cpp:156-170); -1 where the window ends */
    int picstep_window_neighbors(int32_t n_ranks, int32_t periodic, int32_t position, int32_t slides, int32_t* lower, int32_t* upper);
    /* guard widths actually exchanged for E/B (max of interpolation and solver margins,
     * fields/EMFieldBase.x.cpp:58-110) and J (current solver margins, FieldJ.x.cpp:78-118): out[0]=lower, out[1]=upper */
    int picstep_exchange_widths(int32_t shape, int32_t field_solver, int32_t lehe_dir, int32_t field, int32_t axis, int32_t* out2);

#ifdef __cplusplus
}
#endif
#endif /* PICSTEP_H */
