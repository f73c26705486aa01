// picoracle — CPU restatement of PIConGPU's core PIC step (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// This file is the parity oracle for picongpu_b200.  It restates, in plain C++17 (fp32, no FMA
// contraction: build with -ffp-contract=off), the arithmetic of the reference's hot path so that the
// CUDA kernels can be compared against it on identical inputs.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Parity status: PINNED by the reference's own known-answer tests (tests/test_oracle_golden.py):
//   * share/picongpu/unit/MoveParticle.cpp:98-202   (7 cell/supercell crossing scenarios)
//   * share/picongpu/unit/shape.cpp:176-207         (partition of unity, mt19937(42), on/off support)
//   * include/pmacc/test/particles/memory/SuperCell.hpp:69-98 (last-frame arithmetic)
//   * share/picongpu/tests/CurrentDeposition (Python Esirkepov reference imported from the reference
//     tree to generate tests/golden/current_deposition.npz)
//   * share/picongpu/tests/Pusher/README.rst        (gyro radius / phase drift bounds)
//   * share/picongpu/tests/PusherScaling/README.rst (phase lag ~ dt^2: exponent 2 +- 0.1, std <= 0.05)
//   * share/picongpu/tests/FieldAbsorber (the PML acceptance test: radiating wire, small box against a reflection-free
//     box, lib/python/test/FieldAbsorber/validate.py bound 1e-4; tests/test_field_absorber_acceptance.py)
// Restated without a golden vector in the reference (checked by their own known answers in
// tests/test_oracle_golden.py and by construction against the cited source): the Higuera-Cary pusher
// (gyration test of share/picongpu/tests/Pusher applies), the Binomial current interpolation (delta
// response = 1-2-1 tensor weights / 64), the exponential absorber (attenuation profile), the open-boundary step.
// PARITY UNPINNED against reference outputs: the incident-field source (Huygens surface, PlaneWave / GaussianPulse /
// Wavepacket / Polynom / ExpRampWithPrepulse profiles) -- the reference tests it by compiling only
// (share/picongpu/tests/compileLaser), there is no fixture to pin against.  What the tests hold instead: independent
// float64 formulas on the Yee positions (textbook complex Gaussian beam with scipy's Laguerre polynomials, the
// documented envelopes), and the defining properties in vacuum (requested amplitude, focus position and waist, nothing
// behind the total-field / scattered-field surface beyond the truncation of the beam).
// The coupled step as a whole is additionally pinned on the GPU by the reference's acceptance test
// share/picongpu/tests/KHI_growthRate (tests/test_gpu_parity.py::test_khi_growth_rate_reference_acceptance).
// The reference binary itself cannot be built in this image (needs Boost + MPI), see DESIGN.md.
//
// All paths cited below are relative to /root/reference/include/ unless stated otherwise
// (P/ = picongpu/, M/ = pmacc/).
//
// Data conventions (shared with the CUDA library):
//   * local domain n[3] cells, guard g[3] cells per side (GuardSize * SuperCellSize), padded N = n + 2g
//   * fields are SoA: F[c][z][y][x], x fastest, index ((z*Ny + y)*Nx + x), local cell (0,0,0) at (g0,g1,g2)
//   * particles: pos[3][np] in-cell [0,1), mom[3][np], w[np], cell[np] = cx + n0*(cy + n1*cz)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <vector>
#ifdef _OPENMP
#    include <omp.h>
#endif

extern "C"
{
    struct OrcParams
    {
        int n[3]; // local cells (no guard)
        int sc[3]; // SuperCellSize (P/param/memory.param:51)
        int g[3]; // guard cells per side
        float cell[3]; // sim.pic.getCellSize()
        float dt; // sim.pic.getDt()
        float c; // sim.pic.getSpeedOfLight()
        float eps0; // sim.pic.getEps0()
        float mue0; // sim.pic.getMue0()
        float base_mass; // sim.pic.getBaseMass()
        float base_charge; // sim.pic.getBaseCharge()
        int shape; // 0 NGP, 1 CIC, 2 TSC, 3 PQS, 4 PCS
        int pusher; // 0 Boris, 1 Vay, 2 HigueraCary
        int current; // 0 Esirkepov, 1 EmZ
        int solver; // 0 Yee, 1 Lehe
        int lehe_dir; // Cherenkov-free direction for Lehe
        int wrap[3]; // 1: periodic wrap inside this domain; 0: leaving particles get cell coordinate -1 / n
        int current_interp; // 0 None, 1 Binomial (--currentInterpolation)
        int open[3][2]; // [axis][lower, upper]: face is a non-periodic outer boundary (no neighbour, no exchange)
        int absorber_cells[3][2]; // exponential absorber thickness, NUM_CELLS (P/param/fieldAbsorber.param:53-57)
        float absorber_strength[3][2]; // exponential::STRENGTH (fieldAbsorber.param:68-72)
    };
}

namespace
{
    using f32 = float;

    // ---------------------------------------------------------------------------------------------
    // Shapes: P/particles/shapes/{NGP,CIC,TSC,PQS,PCS}.hpp
    // ---------------------------------------------------------------------------------------------
    constexpr int MAXS = 6; // support+1 of PCS

    inline int shapeSupport(int shape)
    {
        return shape + 1;
    }

    // TSC.hpp:47-64
    inline f32 tsc_r1(f32 x)
    {
        f32 const sq = x * x;
        return 0.75f - sq;
    }
    inline f32 tsc_r2(f32 x)
    {
        f32 const tmp = 3.0f / 2.0f - x;
        f32 const sq = tmp * tmp;
        return 0.5f * sq;
    }
    // PQS.hpp:47-66
    inline f32 pqs_r1(f32 x)
    {
        f32 const sq = x * x;
        f32 const tr = sq * x;
        return 1.0f / 6.0f * (4.0f - 6.0f * sq + 3.0f * tr);
    }
    inline f32 pqs_r2(f32 x)
    {
        f32 const tmp = 2.0f - x;
        f32 const tr = tmp * tmp * tmp;
        return 1.0f / 6.0f * tr;
    }
    // PCS.hpp:47-77
    inline f32 pcs_r1(f32 x)
    {
        f32 const sq = x * x;
        return 115.f / 192.f + sq * (-5.f / 8.f + 1.0f / 4.0f * sq);
    }
    inline f32 pcs_r2(f32 x)
    {
        return 1.f / 96.f * (55.f + 4.f * x * (5.f - 2.f * x * (15.f + 2.f * x * (-5.f + x))));
    }
    inline f32 pcs_r3(f32 x)
    {
        f32 const tmp = 5.f - 2.f * x;
        f32 const sq = tmp * tmp;
        f32 const bi = sq * sq;
        return 1.f / 384.f * bi;
    }

    /** detail::<Shape>::shapeArray — values on the support points, particle on support.
     * NGP.hpp:58-65, CIC.hpp:58-69, TSC.hpp:76-88, PQS.hpp:77-90, PCS.hpp:88-102 */
    inline void shapeArrayOnSupport(int shape, f32 x, f32* v)
    {
        switch(shape)
        {
        case 0:
            v[0] = 1.0f;
            break;
        case 1:
            v[0] = 1.0f - x;
            v[1] = x;
            break;
        case 2:
            v[0] = tsc_r2(std::fabs(-1.f - x));
            v[1] = tsc_r1(std::fabs(x));
            v[2] = 1.0f - (v[0] + v[1]);
            break;
        case 3:
            v[0] = pqs_r2(std::fabs(-1.f - x));
            v[1] = pqs_r1(x);
            v[3] = pqs_r2(2.f - x);
            v[2] = 1.0f - (v[0] + v[1] + v[3]);
            break;
        default:
            v[0] = pcs_r3(std::fabs(-2.f - x));
            v[1] = pcs_r2(std::fabs(-1.f - x));
            v[2] = pcs_r1(std::fabs(x));
            v[4] = pcs_r3(2.f - x);
            v[3] = 1.0f - (v[0] + v[1] + v[2] + v[4]);
            break;
        }
    }

    /** ChargeAssignment::shapeArray(x, isOutOfRange) — support+1 values, shifted by one slot if the
     * particle sits in the neighbouring assignment cell (e.g. TSC.hpp:140-153). */
    inline void shapeArrayOffSupport(int shape, f32 xx, bool isOutOfRange, f32* v)
    {
        int const supp = shapeSupport(shape);
        f32 const x = isOutOfRange ? xx - 1.0f : xx;
        f32 t[MAXS];
        shapeArrayOnSupport(shape, x, t);
        // identical select chain as the reference, written as a loop from the top slot downwards
        v[supp] = isOutOfRange ? t[supp - 1] : 0.0f;
        for(int i = supp - 1; i >= 1; --i)
            v[i] = isOutOfRange ? t[i - 1] : t[i];
        v[0] = isOutOfRange ? 0.0f : t[0];
    }

    /** ChargeAssignmentOnSupport::operator()(x): NGP.hpp:127-133, CIC.hpp:141-147, TSC.hpp:161-185,
     * PQS.hpp:163-186, PCS.hpp:190-214 */
    inline f32 shapeEvalOnSupport(int shape, f32 x)
    {
        f32 const a = std::fabs(x);
        switch(shape)
        {
        case 0:
            return 1.0f;
        case 1:
            return 1.0f - a;
        case 2:
        {
            f32 const r1 = tsc_r1(a), r2 = tsc_r2(a);
            return a < 0.5f ? r1 : r2;
        }
        case 3:
        {
            f32 const r1 = pqs_r1(a), r2 = pqs_r2(a);
            return a < 1.0f ? r1 : r2;
        }
        default:
        {
            f32 const r1 = pcs_r1(a), r2 = pcs_r2(a), r3 = pcs_r3(a);
            f32 r = r3;
            if(a < 0.5f)
                r = r1;
            else if(a < 1.5f)
                r = r2;
            return r;
        }
        }
    }

    /** ChargeAssignment::operator()(x): NGP.hpp:80-95, CIC.hpp:88-106, TSC.hpp:108-131, PQS.hpp:110-133,
     * PCS.hpp:123-148 */
    inline f32 shapeEval(int shape, f32 x)
    {
        f32 const a = std::fabs(x);
        switch(shape)
        {
        case 0:
            return f32(-0.5f <= x && x < 0.5f);
        case 1:
            return a < 1.0f ? 1.0f - a : 0.0f;
        case 2:
        {
            f32 const r1 = tsc_r1(a), r2 = tsc_r2(a);
            f32 r = 0.0f;
            if(a < 0.5f)
                r = r1;
            else if(a < 1.5f)
                r = r2;
            return r;
        }
        case 3:
        {
            f32 const r1 = pqs_r1(a), r2 = pqs_r2(a);
            f32 r = 0.0f;
            if(a < 1.0f)
                r = r1;
            else if(a < 2.0f)
                r = r2;
            return r;
        }
        default:
        {
            f32 const on = shapeEvalOnSupport(4, a);
            return a < 2.5f ? on : 0.0f;
        }
        }
    }

    // begin offsets: ChargeAssignment(OnSupport)::begin
    inline int shapeBegin(int shape)
    {
        static int const b[5] = {0, 0, -1, -1, -2};
        return b[shape];
    }

    // ---------------------------------------------------------------------------------------------
    // Domain helpers
    // ---------------------------------------------------------------------------------------------
    struct Dom
    {
        int n[3], g[3], N[3], sc[3], nsc[3];
        int64_t vol;
        explicit Dom(OrcParams const& p)
        {
            for(int d = 0; d < 3; ++d)
            {
                n[d] = p.n[d];
                g[d] = p.g[d];
                N[d] = n[d] + 2 * g[d];
                sc[d] = p.sc[d];
                nsc[d] = n[d] / sc[d];
            }
            vol = int64_t(N[0]) * N[1] * N[2];
        }
        inline int64_t idx(int x, int y, int z) const
        {
            return (int64_t(z) * N[1] + y) * N[0] + x;
        }
    };

    struct Field3
    {
        f32* c[3];
        Field3(f32* base, int64_t vol)
        {
            c[0] = base;
            c[1] = base + vol;
            c[2] = base + 2 * vol;
        }
    };

    // ---------------------------------------------------------------------------------------------
    // Field -> particle interpolation
    // P/algorithms/FieldToParticleInterpolation.hpp:97-124, AssignedTrilinearInterpolation.hpp:54-86,
    // ShiftCoordinateSystem.hpp:54-79, P/fields/YeeCell.hpp:70-130
    // ---------------------------------------------------------------------------------------------
    f32 const fieldPosE[3][3] = {{0.5f, 0.f, 0.f}, {0.f, 0.5f, 0.f}, {0.f, 0.f, 0.5f}};
    f32 const fieldPosB[3][3] = {{0.f, 0.5f, 0.5f}, {0.5f, 0.f, 0.5f}, {0.5f, 0.5f, 0.f}};

    inline f32 interpolateComponent(
        Dom const& D,
        f32 const* F,
        int shape,
        int const cellG[3], // particle cell incl. guard offset
        f32 const pos[3],
        f32 const fpos[3])
    {
        int const supp = shapeSupport(shape);
        bool const isEven = (supp % 2) == 0;
        int const begin = -supp / 2 + (supp + 1) % 2;
        int const end = begin + supp - 1;
        f32 S[3][MAXS];
        int base[3];
        for(int d = 0; d < 3; ++d)
        {
            f32 const v_pos = pos[d] - fpos[d] - 0.5f;
            int shift;
            if(isEven)
                shift = v_pos >= -0.5f ? 0 : -1;
            else
                shift = v_pos >= 0.0f ? 1 : 0;
            f32 const p = v_pos - f32(shift) + 0.5f;
            base[d] = cellG[d] + shift;
            // shapes::Cached<ChargeAssignmentOnSupport>(pos, true) -> shapeArray(pos, false)
            shapeArrayOnSupport(shape, p, S[d]);
        }
        f32 result_z = 0.0f;
        for(int z = begin; z <= end; ++z)
        {
            f32 result_y = 0.0f;
            for(int y = begin; y <= end; ++y)
            {
                f32 result_x = 0.0f;
                for(int x = begin; x <= end; ++x)
                    result_x += F[D.idx(base[0] + x, base[1] + y, base[2] + z)] * S[0][x - begin];
                result_y += result_x * S[1][y - begin];
            }
            result_z += result_y * S[2][z - begin];
        }
        return result_z;
    }

    // ---------------------------------------------------------------------------------------------
    // Pushers.  P/particles/pusher/particlePusherBoris.hpp:42-91, particlePusherVay.hpp:43-112,
    // P/algorithms/Gamma.hpp:30-38, Velocity.hpp:28-38 (CPU backend: rsqrt == 1/sqrt)
    // ---------------------------------------------------------------------------------------------
    inline f32 l2norm2(f32 const v[3])
    {
        f32 tmp = v[0] * v[0];
        tmp += v[1] * v[1];
        tmp += v[2] * v[2];
        return tmp;
    }
    inline void cross(f32 const a[3], f32 const b[3], f32 r[3])
    {
        r[0] = a[1] * b[2] - a[2] * b[1];
        r[1] = a[2] * b[0] - a[0] * b[2];
        r[2] = a[0] * b[1] - a[1] * b[0];
    }
    inline f32 gammaF(OrcParams const& P, f32 const mom[3], f32 mass)
    {
        f32 const fMom2 = l2norm2(mom);
        f32 const c2 = P.c * P.c;
        f32 const m2_c2_reci = 1.0f / (mass * mass * c2);
        return std::sqrt(1.0f + fMom2 * m2_c2_reci);
    }
    inline void velocityF(OrcParams const& P, f32 const mom[3], f32 mass0, f32 vel[3])
    {
        f32 const rc2 = f32(1. / double(P.c) / double(P.c)); // getMue0Eps0()
        f32 const m0_2 = mass0 * mass0;
        f32 const fMom2 = l2norm2(mom);
        f32 const t = 1.0f / std::sqrt(m0_2 + fMom2 * rc2);
        for(int d = 0; d < 3; ++d)
            vel[d] = t * mom[d];
    }

    inline void pushBoris(OrcParams const& P, f32 mass, f32 charge, f32 const E[3], f32 const B[3], f32 mom[3], f32 pos[3])
    {
        f32 const QoM = charge / mass;
        f32 const dt = P.dt;
        f32 mom_minus[3], t[3], s[3], tmp[3], mom_prime[3], mom_plus[3], vel[3];
        for(int d = 0; d < 3; ++d)
            mom_minus[d] = mom[d] + 0.5f * charge * E[d] * dt;
        f32 const gamma_reci = 1.0f / gammaF(P, mom_minus, mass);
        for(int d = 0; d < 3; ++d)
            t[d] = 0.5f * QoM * B[d] * gamma_reci * dt;
        f32 const sfac = 1.0f / (1.0f + l2norm2(t));
        for(int d = 0; d < 3; ++d)
            s[d] = 2.0f * t[d] * sfac;
        cross(mom_minus, t, tmp);
        for(int d = 0; d < 3; ++d)
            mom_prime[d] = mom_minus[d] + tmp[d];
        cross(mom_prime, s, tmp);
        for(int d = 0; d < 3; ++d)
            mom_plus[d] = mom_minus[d] + tmp[d];
        for(int d = 0; d < 3; ++d)
            mom[d] = mom_plus[d] + 0.5f * charge * E[d] * dt;
        velocityF(P, mom, mass, vel);
        for(int d = 0; d < 3; ++d)
            pos[d] += (vel[d] * dt) / P.cell[d];
    }

    inline void pushVay(OrcParams const& P, f32 mass, f32 charge, f32 const E[3], f32 const B[3], f32 mom[3], f32 pos[3])
    {
        f32 const dt = P.dt;
        f32 const factor = f32(0.5 * double(charge) * double(dt)); // `0.5 * charge * deltaT` promotes to double
        f32 vel0[3], cr[3], mom0[3], momp[3];
        velocityF(P, mom, mass, vel0);
        cross(vel0, B, cr);
        for(int d = 0; d < 3; ++d)
            mom0[d] = mom[d] + factor * (E[d] + cr[d]);
        for(int d = 0; d < 3; ++d)
            momp[d] = mom0[d] + factor * E[d];
        f32 const gamma_prime = gammaF(P, momp, mass);
        // sqrt_Vay = precision64Bit (P/param/pusher.param:62)
        double tau[3];
        for(int d = 0; d < 3; ++d)
            tau[d] = double(factor / mass * B[d]);
        double dotpt = double(momp[0]) * tau[0];
        dotpt += double(momp[1]) * tau[1];
        dotpt += double(momp[2]) * tau[2];
        double const u_star = dotpt / double(P.c * mass);
        double tau2 = tau[0] * tau[0];
        tau2 += tau[1] * tau[1];
        tau2 += tau[2] * tau[2];
        double const sigma = double(gamma_prime * gamma_prime) - tau2;
        double const gamma_plus = std::sqrt(0.5 * (sigma + std::sqrt(sigma * sigma + 4.0 * (tau2 + u_star * u_star))));
        f32 t[3];
        for(int d = 0; d < 3; ++d)
            t[d] = f32(tau[d] * double(1.0f / gamma_plus));
        f32 const s = 1.0f / (1.0f + l2norm2(t));
        f32 dpt = momp[0] * t[0];
        dpt += momp[1] * t[1];
        dpt += momp[2] * t[2];
        cross(momp, t, cr);
        for(int d = 0; d < 3; ++d)
            mom[d] = s * (momp[d] + dpt * t[d] + cr[d]);
        f32 vel[3];
        velocityF(P, mom, mass, vel);
        for(int d = 0; d < 3; ++d)
            pos[d] += (vel[d] * dt) / P.cell[d];
    }

    // P/particles/pusher/particlePusherHigueraCary.hpp:45-145 (sqrt_HigueraCary = precision64Bit, P/param/pusher.param:71;
    // Gamma<> computes in float_X, P/unitless/pusher.unitless:55)
    inline void pushHigueraCary(OrcParams const& P, f32 mass, f32 charge, f32 const E[3], f32 const B[3], f32 mom[3], f32 pos[3])
    {
        f32 const dt = P.dt;
        f32 half_e[3], mm32[3];
        double mom_minus[3], tau[3];
        for(int d = 0; d < 3; ++d)
        {
            half_e[d] = 0.5f * charge * E[d] * dt;
            mm32[d] = mom[d] + half_e[d];
            mom_minus[d] = double(mm32[d]);
            tau[d] = double(0.5f * B[d] * charge * dt / mass);
        }
        double const gamma_minus = double(gammaF(P, mm32, mass));
        double tau2 = tau[0] * tau[0];
        tau2 += tau[1] * tau[1];
        tau2 += tau[2] * tau[2];
        double const sigma = gamma_minus * gamma_minus - tau2;
        double dotpt = mom_minus[0] * tau[0];
        dotpt += mom_minus[1] * tau[1];
        dotpt += mom_minus[2] * tau[2];
        double const u_star = dotpt / double(mass * P.c);
        double const gamma_plus = std::sqrt(0.5 * (sigma + std::sqrt(sigma * sigma + 4.0 * (tau2 + u_star * u_star))));
        double t[3];
        for(int d = 0; d < 3; ++d)
            t[d] = tau[d] / gamma_plus;
        double t2 = t[0] * t[0];
        t2 += t[1] * t[1];
        t2 += t[2] * t[2];
        double const sfac = 1.0 / (1.0 + t2);
        double dmt = mom_minus[0] * t[0];
        dmt += mom_minus[1] * t[1];
        dmt += mom_minus[2] * t[2];
        double cr[3] = {mom_minus[1] * t[2] - mom_minus[2] * t[1], mom_minus[2] * t[0] - mom_minus[0] * t[2], mom_minus[0] * t[1] - mom_minus[1] * t[0]};
        double mom_plus[3];
        for(int d = 0; d < 3; ++d)
            mom_plus[d] = sfac * (mom_minus[d] + dmt * t[d] + cr[d]);
        double const cr2[3] = {mom_plus[1] * t[2] - mom_plus[2] * t[1], mom_plus[2] * t[0] - mom_plus[0] * t[2], mom_plus[0] * t[1] - mom_plus[1] * t[0]};
        for(int d = 0; d < 3; ++d)
        {
            f32 const mom_diff = half_e[d] + f32(cr2[d]);
            mom[d] = f32(mom_plus[d]) + mom_diff;
        }
        f32 vel[3];
        velocityF(P, mom, mass, vel);
        for(int d = 0; d < 3; ++d)
            pos[d] += (vel[d] * dt) / P.cell[d];
    }

    // ---------------------------------------------------------------------------------------------
    // moveParticle: P/particles/MoveParticle.hpp:48-160
    // ---------------------------------------------------------------------------------------------
    /** @param localCell in/out cell coordinates inside the supercell
     *  @param dirOut per-dimension cell crossing direction (before masking to supercell crossings)
     *  @return multiMask (1 = stays in supercell, >=2 leaves into direction multiMask-1) */
    inline int moveParticle(int const sc[3], f32 const newPos[3], f32 posOut[3], int localCell[3], int dirOut[3])
    {
        f32 const shift = 0.5f;
        int dir[3];
        for(int i = 0; i < 3; ++i)
        {
            f32 pos = newPos[i] - shift;
            f32 moveDir = 0.0f;
            if(pos < -0.5f)
                moveDir = -1.0f;
            if(pos >= 0.5f)
                moveDir = 1.0f;
            pos -= moveDir;
            posOut[i] = pos + shift;
            dir[i] = int(moveDir);
            dirOut[i] = dir[i];
        }
        int newMultimask = 1;
        if(dir[0] != 0 || dir[1] != 0 || dir[2] != 0)
        {
            for(int i = 0; i < 3; ++i)
                localCell[i] += dir[i];
            for(int i = 0; i < 3; ++i)
                dir[i] = uint32_t(localCell[i]) >= uint32_t(sc[i]) ? dir[i] : 0;
            for(int i = 0; i < 3; ++i)
                localCell[i] -= dir[i] * sc[i];
            uint32_t exchangeType = 1;
            for(int i = 0; i < 3; ++i)
            {
                newMultimask += (dir[i] == -1 ? 2 : dir[i]) * int(exchangeType);
                exchangeType *= 3;
            }
        }
        return newMultimask;
    }

    // ---------------------------------------------------------------------------------------------
    // Current deposition.  P/fields/FieldJ.kernel:110-142, currentDeposition/Esirkepov/Esirkepov.hpp:62-242,
    // relayPoint.hpp:48-63, EmZ/EmZ.hpp:66-155, EmZ/DepositCurrent.hpp:35-119,
    // PermutatedFieldValueAccess.hpp:77-99
    // ---------------------------------------------------------------------------------------------
    inline f32 relayPoint(bool isEven, int& i_1, int& i_2, f32 x_1, f32 x_2)
    {
        if(isEven)
        {
            i_1 = int(std::floor(x_1));
            i_2 = int(std::floor(x_2));
            return i_1 == i_2 ? x_2 : f32(std::max(i_1, i_2));
        }
        i_1 = int(std::floor(x_1 + 0.5f));
        i_2 = int(std::floor(x_2 + 0.5f));
        return i_1 == i_2 ? x_2 : f32(i_1 + i_2) / 2.0f;
    }

    /** Accumulator: adds `val` to component `comp` of J at cell (base + off). */
    struct JAcc
    {
        Dom const* D;
        f32* J[3];
        inline void add(int comp, int x, int y, int z, f32 val) const
        {
            J[comp][D->idx(x, y, z)] += val;
        }
    };

    /** Esirkepov::cptCurrent1D (Esirkepov.hpp:147-242) for one rotated direction.
     * perm: rotated axis r -> original axis perm[r]; component = perm[2]. */
    inline void esirkepovCpt1D(
        OrcParams const& P,
        JAcc const& acc,
        int shape,
        int const status[3], // rotated
        int const baseCell[3], // original coordinates (already shifted by gridShift)
        f32 const pos0[3], // rotated
        f32 const pos1[3],
        int const perm[3],
        f32 cellEdgeLength,
        f32 charge)
    {
        if(pos0[2] == pos1[2])
            return;
        int const supp = shapeSupport(shape);
        int const begin = shapeBegin(shape);
        int const end = begin + supp;
        f32 S0[3][MAXS], S1[3][MAXS];
        for(int r = 0; r < 3; ++r)
        {
            bool const startIn = (status[r] & 2) != 0;
            bool const endIn = (status[r] & 4) != 0;
            // shapes::Cached<ChargeAssignment>(pos, isInBase) -> shapeArray(pos, !isInBase)
            shapeArrayOffSupport(shape, pos0[r], !startIn, S0[r]);
            shapeArrayOffSupport(shape, pos1[r], !endIn, S1[r]);
        }
        f32 const vol = P.cell[0] * P.cell[1] * P.cell[2];
        f32 const currentSurfaceDensity = charge * (1.0f / f32(vol * P.dt)) * cellEdgeLength;
        int const leaveI = status[0] & 1, leaveJ = status[1] & 1, leaveK = status[2] & 1;
        for(int i = begin; i < end + 1; ++i)
            if(i < end + leaveI)
            {
                f32 const s0i = S0[0][i - begin];
                f32 const dsi = S1[0][i - begin] - s0i;
                for(int j = begin; j < end + 1; ++j)
                    if(j < end + leaveJ)
                    {
                        f32 const s0j = S0[1][j - begin];
                        f32 const dsj = S1[1][j - begin] - s0j;
                        f32 const tmp = -currentSurfaceDensity
                            * (s0i * s0j + 0.5f * (dsi * s0j + s0i * dsj) + (1.0f / 3.0f) * dsj * dsi);
                        f32 accumulated_J = 0.0f;
                        for(int k = begin; k < end; ++k)
                            if(k < end + leaveK - 1)
                            {
                                f32 const W = (S1[2][k - begin] - S0[2][k - begin]) * tmp;
                                accumulated_J += W;
                                int o[3];
                                o[perm[0]] = i;
                                o[perm[1]] = j;
                                o[perm[2]] = k;
                                acc.add(perm[2], baseCell[0] + o[0], baseCell[1] + o[1], baseCell[2] + o[2], accumulated_J);
                            }
                    }
            }
    }

    inline void depositEsirkepov(
        OrcParams const& P,
        JAcc const& acc,
        int const cellG[3],
        f32 const pos[3],
        f32 const vel[3],
        f32 charge)
    {
        int const shape = P.shape;
        bool const isEven = (shapeSupport(shape) % 2) == 0;
        f32 p0[3], p1[3];
        int status[3] = {0, 0, 0};
        int base[3];
        for(int d = 0; d < 3; ++d)
        {
            f32 const deltaPos = vel[d] * P.dt / P.cell[d];
            p0[d] = pos[d] - deltaPos;
            p1[d] = pos[d];
            int iStart, iEnd;
            relayPoint(isEven, iStart, iEnd, p0[d], p1[d]);
            int const gridShift = iStart < iEnd ? iStart : iEnd;
            status[d] |= (gridShift == iStart) ? 2 : 0;
            status[d] |= (gridShift == iEnd) ? 4 : 0;
            status[d] |= (iStart != iEnd) ? 1 : 0;
            p0[d] -= f32(gridShift);
            p1[d] -= f32(gridShift);
            base[d] = cellG[d] + gridShift;
        }
        {
            int const perm[3] = {1, 2, 0};
            int const st[3] = {status[1], status[2], status[0]};
            f32 const a0[3] = {p0[1], p0[2], p0[0]}, a1[3] = {p1[1], p1[2], p1[0]};
            esirkepovCpt1D(P, acc, shape, st, base, a0, a1, perm, P.cell[0], charge);
        }
        {
            int const perm[3] = {2, 0, 1};
            int const st[3] = {status[2], status[0], status[1]};
            f32 const a0[3] = {p0[2], p0[0], p0[1]}, a1[3] = {p1[2], p1[0], p1[1]};
            esirkepovCpt1D(P, acc, shape, st, base, a0, a1, perm, P.cell[1], charge);
        }
        {
            int const perm[3] = {0, 1, 2};
            esirkepovCpt1D(P, acc, shape, status, base, p0, p1, perm, P.cell[2], charge);
        }
    }

    /** emz::DepositCurrent<...,DIM3>::cptCurrent1D (EmZ/DepositCurrent.hpp:77-118) */
    inline void emzCpt1D(
        JAcc const& acc,
        int shape,
        int const baseCell[3],
        f32 const pos0[3],
        f32 const pos1[3],
        int const perm[3],
        f32 currentSurfaceDensity)
    {
        if(pos0[2] == pos1[2])
            return;
        int const supp = shapeSupport(shape);
        int const begin = shapeBegin(shape);
        int const end = begin + supp;
        f32 S0[3][MAXS], S1[3][MAXS];
        for(int r = 0; r < 3; ++r)
        {
            shapeArrayOnSupport(shape, pos0[r], S0[r]);
            shapeArrayOnSupport(shape, pos1[r], S1[r]);
        }
        for(int i = begin; i < end; ++i)
        {
            f32 const s0i = S0[0][i - begin];
            f32 const dsi = S1[0][i - begin] - s0i;
            for(int j = begin; j < end; ++j)
            {
                f32 const s0j = S0[1][j - begin];
                f32 const dsj = S1[1][j - begin] - s0j;
                f32 const tmp
                    = -currentSurfaceDensity * (s0i * s0j + 0.5f * (dsi * s0j + s0i * dsj) + (1.0f / 3.0f) * dsj * dsi);
                f32 accumulated_J = 0.0f;
                for(int k = begin; k < end - 1; ++k)
                {
                    f32 const W = (S1[2][k - begin] - S0[2][k - begin]) * tmp;
                    accumulated_J += W;
                    int o[3];
                    o[perm[0]] = i;
                    o[perm[1]] = j;
                    o[perm[2]] = k;
                    acc.add(perm[2], baseCell[0] + o[0], baseCell[1] + o[1], baseCell[2] + o[2], accumulated_J);
                }
            }
        }
    }

    inline void emzDeposit3(
        OrcParams const& P,
        JAcc const& acc,
        int shape,
        int const base[3],
        f32 const p0[3],
        f32 const p1[3],
        f32 chargeDensity)
    {
        {
            int const perm[3] = {1, 2, 0};
            f32 const a0[3] = {p0[1], p0[2], p0[0]}, a1[3] = {p1[1], p1[2], p1[0]};
            emzCpt1D(acc, shape, base, a0, a1, perm, P.cell[0] * chargeDensity / P.dt);
        }
        {
            int const perm[3] = {2, 0, 1};
            f32 const a0[3] = {p0[2], p0[0], p0[1]}, a1[3] = {p1[2], p1[0], p1[1]};
            emzCpt1D(acc, shape, base, a0, a1, perm, P.cell[1] * chargeDensity / P.dt);
        }
        {
            int const perm[3] = {0, 1, 2};
            emzCpt1D(acc, shape, base, p0, p1, perm, P.cell[2] * chargeDensity / P.dt);
        }
    }

    inline void depositEmZ(
        OrcParams const& P,
        JAcc const& acc,
        int const cellG[3],
        f32 const posEnd[3],
        f32 const vel[3],
        f32 charge)
    {
        int const shape = P.shape;
        bool const isEven = (shapeSupport(shape) % 2) == 0;
        f32 posStart[3], relay[3];
        int shiftStart[3], shiftEnd[3];
        for(int d = 0; d < 3; ++d)
        {
            f32 const deltaPos = (vel[d] * P.dt) / P.cell[d];
            posStart[d] = posEnd[d] - deltaPos;
            relay[d] = relayPoint(isEven, shiftStart[d], shiftEnd[d], posStart[d], posEnd[d]);
        }
        f32 const chargeDensity = charge / (P.cell[0] * P.cell[1] * P.cell[2]);
        f32 l0[3], l1[3];
        int base[3];
        for(int d = 0; d < 3; ++d)
        {
            l0[d] = posStart[d] - f32(shiftStart[d]);
            l1[d] = relay[d] - f32(shiftStart[d]);
            base[d] = cellG[d] + shiftStart[d];
        }
        emzDeposit3(P, acc, shape, base, l0, l1, chargeDensity);
        bool const two = shiftStart[0] != shiftEnd[0] || shiftStart[1] != shiftEnd[1] || shiftStart[2] != shiftEnd[2];
        if(two)
        {
            for(int d = 0; d < 3; ++d)
            {
                l1[d] = posEnd[d] - f32(shiftEnd[d]);
                l0[d] = relay[d] - f32(shiftEnd[d]);
                base[d] = cellG[d] + shiftEnd[d];
            }
            emzDeposit3(P, acc, shape, base, l0, l1, chargeDensity);
        }
    }

    // ---------------------------------------------------------------------------------------------
    // Field solver.  P/fields/MaxwellSolver/FDTD/FDTDBase.kernel:51-124, differentiation/Curl.hpp:84-90,
    // ForwardDerivative.hpp:58-63, BackwardDerivative.hpp:58-63, Lehe/Derivative.hpp:66-236
    // ---------------------------------------------------------------------------------------------
    struct LeheCoeff
    {
        f32 alpha, delta, betaDir1, betaDir2;
    };

    inline LeheCoeff leheCoeff(OrcParams const& P, int dir0)
    {
        // Lehe/Derivative.hpp:94-111 (float_64 arithmetic on the float_X PIC-unit values), :134-137 (fp32 betas)
        int const dir1 = (dir0 + 1) % 3, dir2 = (dir0 + 2) % 3;
        LeheCoeff r;
        double const stepRatio = double(P.cell[dir0] / (P.c * P.dt));
        double const coeff = stepRatio * std::sin(1.5707963267948966 * double(P.c) * double(P.dt) / double(P.cell[dir0]));
        r.delta = f32(0.25 * (1.0 - coeff * coeff));
        double const sr1 = double(P.cell[dir0] / P.cell[dir1]);
        double const sr2 = double(P.cell[dir0] / P.cell[dir2]);
        double const b1 = 0.125 * sr1 * sr1, b2 = 0.125 * sr2 * sr2;
        r.alpha = f32(1.0 - 2.0 * b1 - 2.0 * b2 - 3.0 * double(r.delta));
        f32 const s1 = P.cell[dir0] / P.cell[dir1], s2 = P.cell[dir0] / P.cell[dir2];
        r.betaDir1 = 0.125f * s1 * s1;
        r.betaDir2 = 0.125f * s2 * s2;
        return r;
    }

    /** derivative along `dir` of all three components at (x,y,z); mode 0 forward, 1 backward, 2 Lehe */
    inline void derivative(
        OrcParams const& P,
        Dom const& D,
        Field3 const& F,
        int mode,
        LeheCoeff const& lc,
        int dir,
        int x,
        int y,
        int z,
        f32 out[3])
    {
        auto at = [&](int c, int dx, int dy, int dz) { return F.c[c][D.idx(x + dx, y + dy, z + dz)]; };
        int e[3] = {0, 0, 0};
        e[dir] = 1;
        auto fwd = [&](int c, int ox, int oy, int oz)
        { return (at(c, ox + e[0], oy + e[1], oz + e[2]) - at(c, ox, oy, oz)) / P.cell[dir]; };
        if(mode == 0)
        {
            for(int c = 0; c < 3; ++c)
                out[c] = fwd(c, 0, 0, 0);
        }
        else if(mode == 1)
        {
            for(int c = 0; c < 3; ++c)
                out[c] = (at(c, 0, 0, 0) - at(c, -e[0], -e[1], -e[2])) / P.cell[dir];
        }
        else
        {
            int const cf = P.lehe_dir;
            if(dir == cf)
            {
                int const dir1 = (dir + 1) % 3, dir2 = (dir + 2) % 3;
                int u1[3] = {0, 0, 0}, u2[3] = {0, 0, 0};
                u1[dir1] = 1;
                u2[dir2] = 1;
                for(int c = 0; c < 3; ++c)
                {
                    f32 r = lc.alpha * fwd(c, 0, 0, 0) + lc.betaDir1 * fwd(c, u1[0], u1[1], u1[2]);
                    r = r + lc.betaDir1 * fwd(c, -u1[0], -u1[1], -u1[2]);
                    r = r + lc.betaDir2 * fwd(c, u2[0], u2[1], u2[2]);
                    r = r + lc.betaDir2 * fwd(c, -u2[0], -u2[1], -u2[2]);
                    r = r + lc.delta * (at(c, 2 * e[0], 2 * e[1], 2 * e[2]) - at(c, -e[0], -e[1], -e[2])) / P.cell[dir];
                    out[c] = r;
                }
            }
            else
            {
                f32 const beta = 0.125f;
                f32 const alpha = 1.0f - 2.0f * beta;
                int u[3] = {0, 0, 0};
                u[cf] = 1;
                for(int c = 0; c < 3; ++c)
                {
                    f32 r = alpha * fwd(c, 0, 0, 0) + beta * fwd(c, u[0], u[1], u[2]);
                    r = r + beta * fwd(c, -u[0], -u[1], -u[2]);
                    out[c] = r;
                }
            }
        }
    }

    inline void curlAt(
        OrcParams const& P,
        Dom const& D,
        Field3 const& F,
        int mode,
        LeheCoeff const* lc,
        int x,
        int y,
        int z,
        f32 out[3])
    {
        f32 dx[3], dy[3], dz[3];
        derivative(P, D, F, mode, lc[0], 0, x, y, z, dx);
        derivative(P, D, F, mode, lc[1], 1, x, y, z, dy);
        derivative(P, D, F, mode, lc[2], 2, x, y, z, dz);
        out[0] = dy[2] - dz[1];
        out[1] = dz[0] - dx[2];
        out[2] = dx[1] - dy[0];
    }

    // ---------------------------------------------------------------------------------------------
    // Counter based RNG for the IC generator: Philox4x32-10 (Salmon et al. 2011), documented layout:
    // key = (seed, species), counter = (particle index lo, hi, stream, 0).
    // ---------------------------------------------------------------------------------------------
    inline void philox4x32_10(uint32_t ctr[4], uint32_t const key_in[2])
    {
        uint32_t key[2] = {key_in[0], key_in[1]};
        for(int r = 0; r < 10; ++r)
        {
            uint64_t const p0 = uint64_t(0xD2511F53u) * ctr[0];
            uint64_t const p1 = uint64_t(0xCD9E8D57u) * ctr[2];
            uint32_t const n0 = uint32_t(p1 >> 32) ^ ctr[1] ^ key[0];
            uint32_t const n1 = uint32_t(p1);
            uint32_t const n2 = uint32_t(p0 >> 32) ^ ctr[3] ^ key[1];
            uint32_t const n3 = uint32_t(p0);
            ctr[0] = n0;
            ctr[1] = n1;
            ctr[2] = n2;
            ctr[3] = n3;
            key[0] += 0x9E3779B9u;
            key[1] += 0xBB67AE85u;
        }
    }
    inline f32 u01(uint32_t r)
    {
        return (f32(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
    }
} // namespace

extern "C"
{
    // ---------------------------------------------------------------------------------------------
    // Unit-level entry points (used by the golden-vector tests)
    // ---------------------------------------------------------------------------------------------
    void orc_shape_array(int shape, int onSupport, float x, int isOutOfRange, float* out)
    {
        if(onSupport)
            shapeArrayOnSupport(shape, x, out);
        else
            shapeArrayOffSupport(shape, x, isOutOfRange != 0, out);
    }

    float orc_shape_eval(int shape, int onSupport, float x)
    {
        return onSupport ? shapeEvalOnSupport(shape, x) : shapeEval(shape, x);
    }

    /** The reference's unit::shape test body (share/picongpu/unit/shape.cpp:128-141,176-207):
     * positions from std::mt19937(42) + uniform_real_distribution<>(0,1), sum of shape(g - p). */
    void orc_shape_unit_test(int shape, int onSupport, int numValues, float* positions, float* sums)
    {
        std::mt19937 mt(42.0);
        std::uniform_real_distribution<> dist(0.0, 1.0);
        int const supp = shapeSupport(shape);
        bool const isEven = supp % 2 == 0;
        int const begin = shapeBegin(shape);
        int const end = onSupport ? begin + supp - 1 : begin + supp;
        for(int n = 0; n < numValues; ++n)
        {
            f32 const pos = f32(dist(mt));
            positions[n] = pos;
            f32 res = 0.0f;
            for(int g = begin; g <= end; ++g)
            {
                f32 p = pos;
                if(onSupport)
                {
                    f32 const v_pos = pos - 0.5f;
                    int s;
                    if(isEven)
                        s = v_pos >= -0.5f ? 0 : -1;
                    else
                        s = v_pos >= 0.0f ? 1 : 0;
                    p = v_pos - f32(s) + 0.5f;
                }
                res += onSupport ? shapeEvalOnSupport(shape, f32(g) - p) : shapeEval(shape, f32(g) - p);
            }
            sums[n] = res;
        }
    }

    /** moveParticle on one particle; localCellIdx linearised x-fastest in the supercell. */
    int orc_move_particle(int const* sc, float const* newPos, int localCellIdx, float* posOut, int* localCellIdxOut)
    {
        int lc[3] = {localCellIdx % sc[0], (localCellIdx / sc[0]) % sc[1], localCellIdx / (sc[0] * sc[1])};
        int dir[3];
        int const mask = moveParticle(sc, newPos, posOut, lc, dir);
        *localCellIdxOut = lc[0] + sc[0] * (lc[1] + sc[1] * lc[2]);
        return mask;
    }

    /** SuperCell::getSizeLastFrame (M/particles/memory/dataTypes/SuperCell.hpp) */
    unsigned orc_size_last_frame(unsigned numParticles, unsigned frameSize)
    {
        return numParticles ? ((numParticles - 1u) % frameSize + 1u) : 0u;
    }

    void orc_lehe_coeff(OrcParams const* P, int dir, float* out4)
    {
        LeheCoeff const c = leheCoeff(*P, dir);
        out4[0] = c.alpha;
        out4[1] = c.delta;
        out4[2] = c.betaDir1;
        out4[3] = c.betaDir2;
    }

    /** Boris/Vay momentum+position update for one particle in given fields (T/Pusher known-answer test). */
    void orc_push_one(OrcParams const* P, float massRatio, float chargeRatio, float w, float const* E, float const* B, float* mom, float* pos)
    {
        f32 const mass = (P->base_mass * massRatio) * w;
        f32 const charge = (P->base_charge * chargeRatio) * w;
        if(P->pusher == 0)
            pushBoris(*P, mass, charge, E, B, mom, pos);
        else if(P->pusher == 1)
            pushVay(*P, mass, charge, E, B, mom, pos);
        else
            pushHigueraCary(*P, mass, charge, E, B, mom, pos);
    }

    // ---------------------------------------------------------------------------------------------
    // Stage-level entry points
    // ---------------------------------------------------------------------------------------------

    /** Interpolate E and B to the particles only (for gather parity tests). Eout/Bout are [3][np]. */
    void orc_gather(OrcParams const* Pp, float* E, float* B, int64_t np, float const* pos, int32_t const* cell, float* Eout, float* Bout)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        Field3 const FE(E, D.vol), FB(B, D.vol);
#pragma omp parallel for schedule(static)
        for(int64_t i = 0; i < np; ++i)
        {
            int c = cell[i];
            int const cg[3] = {c % D.n[0] + D.g[0], (c / D.n[0]) % D.n[1] + D.g[1], c / (D.n[0] * D.n[1]) + D.g[2]};
            f32 const p[3] = {pos[i], pos[np + i], pos[2 * np + i]};
            for(int k = 0; k < 3; ++k)
            {
                Bout[k * np + i] = interpolateComponent(D, FB.c[k], P.shape, cg, p, fieldPosB[k]);
                Eout[k * np + i] = interpolateComponent(D, FE.c[k], P.shape, cg, p, fieldPosE[k]);
            }
        }
    }

    /** KernelMoveAndMarkParticles + PushParticlePerFrame (P/particles/Particles.kernel:170-316) followed by
     * the supercell re-assignment that KernelShiftParticles performs (M/particles/ParticlesBase.kernel:361-615),
     * expressed on a flat particle list: pos/mom updated in place, cell[] becomes the new cell (periodic wrap
     * where wrap[d]=1, else coordinate -1 / n[d] encoded in cellOut3), mask[] = multiMask from moveParticle.
     * cellOut3 (optional, [3][np]) receives unwrapped new cell coordinates. */
    void orc_push(
        OrcParams const* Pp,
        float massRatio,
        float chargeRatio,
        float* E,
        float* B,
        int64_t np,
        float* pos,
        float* mom,
        float const* w,
        int32_t* cell,
        uint8_t* mask,
        int32_t* cellOut3)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        Field3 const FE(E, D.vol), FB(B, D.vol);
#pragma omp parallel for schedule(static)
        for(int64_t i = 0; i < np; ++i)
        {
            int const c = cell[i];
            int cc[3] = {c % D.n[0], (c / D.n[0]) % D.n[1], c / (D.n[0] * D.n[1])};
            int const cg[3] = {cc[0] + D.g[0], cc[1] + D.g[1], cc[2] + D.g[2]};
            f32 p[3] = {pos[i], pos[np + i], pos[2 * np + i]};
            f32 m[3] = {mom[i], mom[np + i], mom[2 * np + i]};
            f32 Ef[3], Bf[3];
            for(int k = 0; k < 3; ++k)
            {
                Bf[k] = interpolateComponent(D, FB.c[k], P.shape, cg, p, fieldPosB[k]);
                Ef[k] = interpolateComponent(D, FE.c[k], P.shape, cg, p, fieldPosE[k]);
            }
            f32 const mass = (P.base_mass * massRatio) * w[i];
            f32 const charge = (P.base_charge * chargeRatio) * w[i];
            if(P.pusher == 0)
                pushBoris(P, mass, charge, Ef, Bf, m, p);
            else if(P.pusher == 1)
                pushVay(P, mass, charge, Ef, Bf, m, p);
            else
                pushHigueraCary(P, mass, charge, Ef, Bf, m, p);
            int lc[3] = {cc[0] % D.sc[0], cc[1] % D.sc[1], cc[2] % D.sc[2]};
            int dir[3];
            f32 pout[3];
            int const mm = moveParticle(D.sc, p, pout, lc, dir);
            for(int d = 0; d < 3; ++d)
            {
                cc[d] += dir[d];
                if(cellOut3)
                    cellOut3[d * np + i] = cc[d];
                if(P.wrap[d])
                    cc[d] = (cc[d] + D.n[d]) % D.n[d];
                pos[d * np + i] = pout[d];
                mom[d * np + i] = m[d];
            }
            bool const inside = cc[0] >= 0 && cc[0] < D.n[0] && cc[1] >= 0 && cc[1] < D.n[1] && cc[2] >= 0 && cc[2] < D.n[2];
            cell[i] = inside ? cc[0] + D.n[0] * (cc[1] + D.n[1] * cc[2]) : -1;
            if(mask)
                mask[i] = uint8_t(mm);
        }
    }

    /** KernelComputeCurrent + ComputePerFrame (P/fields/FieldJ.kernel:52-142): J += deposit(all particles).
     * Particles are processed supercell by supercell (ascending linear supercell index, ascending particle
     * index inside) in 8 checkerboard passes like the omp2b default StridedCachedSupercells
     * (Strategy.def:78-85, Deposit.hpp:63-90); contributions are added directly into J (including guards). */
    void orc_deposit(
        OrcParams const* Pp,
        float massRatio,
        float chargeRatio,
        float* J,
        int64_t np,
        float const* pos,
        float const* mom,
        float const* w,
        int32_t const* cell)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        int const nscTot = D.nsc[0] * D.nsc[1] * D.nsc[2];
        // CSR by supercell (stable)
        std::vector<int64_t> off(size_t(nscTot) + 1, 0);
        std::vector<int32_t> scOf(np);
        for(int64_t i = 0; i < np; ++i)
        {
            int const c = cell[i];
            int const cc[3] = {c % D.n[0], (c / D.n[0]) % D.n[1], c / (D.n[0] * D.n[1])};
            int const s = cc[0] / D.sc[0] + D.nsc[0] * (cc[1] / D.sc[1] + D.nsc[1] * (cc[2] / D.sc[2]));
            scOf[i] = s;
            off[size_t(s) + 1]++;
        }
        for(int s = 0; s < nscTot; ++s)
            off[size_t(s) + 1] += off[s];
        std::vector<int64_t> order(np), cur(off.begin(), off.end() - 1);
        for(int64_t i = 0; i < np; ++i)
            order[cur[scOf[i]]++] = i;

        JAcc acc;
        acc.D = &D;
        acc.J[0] = J;
        acc.J[1] = J + D.vol;
        acc.J[2] = J + 2 * D.vol;
        for(int pass = 0; pass < 8; ++pass)
        {
            int const px = pass & 1, py = (pass >> 1) & 1, pz = (pass >> 2) & 1;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
            for(int sz = pz; sz < D.nsc[2]; sz += 2)
                for(int sy = py; sy < D.nsc[1]; sy += 2)
                    for(int sx = px; sx < D.nsc[0]; sx += 2)
                    {
                        int const s = sx + D.nsc[0] * (sy + D.nsc[1] * sz);
                        for(int64_t q = off[s]; q < off[size_t(s) + 1]; ++q)
                        {
                            int64_t const i = order[q];
                            int const c = cell[i];
                            int const cg[3]
                                = {c % D.n[0] + D.g[0], (c / D.n[0]) % D.n[1] + D.g[1], c / (D.n[0] * D.n[1]) + D.g[2]};
                            f32 const p[3] = {pos[i], pos[np + i], pos[2 * np + i]};
                            f32 const m[3] = {mom[i], mom[np + i], mom[2 * np + i]};
                            f32 const charge = (P.base_charge * chargeRatio) * w[i];
                            f32 const mass = (P.base_mass * massRatio) * w[i];
                            f32 vel[3];
                            velocityF(P, m, mass, vel);
                            if(P.current == 0)
                                depositEsirkepov(P, acc, cg, p, vel, charge);
                            else
                                depositEmZ(P, acc, cg, p, vel, charge);
                        }
                    }
        }
    }

    /** Deposit one particle given explicit velocity (T/CurrentDeposition known-answer test). */
    void orc_deposit_one(OrcParams const* Pp, float* J, int const* cellCoord, float const* pos, float const* vel, float charge)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        JAcc acc;
        acc.D = &D;
        acc.J[0] = J;
        acc.J[1] = J + D.vol;
        acc.J[2] = J + 2 * D.vol;
        int const cg[3] = {cellCoord[0] + D.g[0], cellCoord[1] + D.g[1], cellCoord[2] + D.g[2]};
        if(P.current == 0)
            depositEsirkepov(P, acc, cg, pos, vel, charge);
        else
            depositEmZ(P, acc, cg, pos, vel, charge);
    }

    /** Periodic guard fill for E/B: own GUARD := opposite BORDER (what GridBuffer::asyncCommunication does
     * through the self-neighbour MPI topology, M/memory/buffers/GridBuffer.hpp:472-483). Full guard width. */
    void orc_guard_copy(OrcParams const* Pp, float* F)
    {
        Dom const D(*Pp);
        for(int c = 0; c < 3; ++c)
        {
            f32* f = F + c * D.vol;
#pragma omp parallel for schedule(static)
            for(int z = 0; z < D.N[2]; ++z)
                for(int y = 0; y < D.N[1]; ++y)
                    for(int x = 0; x < D.N[0]; ++x)
                    {
                        int const q[3] = {x, y, z};
                        int s[3];
                        bool guard = false;
                        for(int d = 0; d < 3; ++d)
                        {
                            int l = q[d] - D.g[d];
                            if(l < 0 || l >= D.n[d])
                                guard = true;
                            l = (l % D.n[d] + D.n[d]) % D.n[d];
                            s[d] = l + D.g[d];
                        }
                        if(guard)
                            f[D.idx(x, y, z)] = f[D.idx(s[0], s[1], s[2])];
                    }
        }
    }

    /** Periodic J guard reduction: BORDER += neighbour GUARD (M/fields/operations/AddExchangeToBorder.hpp:43-128
     * driven by FieldJ::asyncCommunication P/fields/FieldJ.x.cpp:156-174). Guards are left untouched. */
    void orc_guard_add(OrcParams const* Pp, float* F)
    {
        Dom const D(*Pp);
        for(int c = 0; c < 3; ++c)
        {
            f32* f = F + c * D.vol;
            // serial: several guard cells map onto the same border cell
            for(int z = 0; z < D.N[2]; ++z)
                for(int y = 0; y < D.N[1]; ++y)
                    for(int x = 0; x < D.N[0]; ++x)
                    {
                        int const q[3] = {x, y, z};
                        int s[3];
                        bool guard = false;
                        for(int d = 0; d < 3; ++d)
                        {
                            int l = q[d] - D.g[d];
                            if(l < 0 || l >= D.n[d])
                                guard = true;
                            l = (l % D.n[d] + D.n[d]) % D.n[d];
                            s[d] = l + D.g[d];
                        }
                        if(guard)
                            f[D.idx(s[0], s[1], s[2])] += f[D.idx(x, y, z)];
                    }
        }
    }

    /** One axis pass of the guard exchange for a single periodic rank along `axis`, spanning the full padded extent
     * of the two other axes (the 26 exchange directions of pmacc/type/Exchange.hpp:46-54 collapse into the passes
     * x, y, z).  add=0: guards [g-lo,g) and [g+n,g+n+up) := opposite border planes (E/B);
     * add=1: border planes += opposite guards (J, AddExchangeToBorder.hpp:43-128). */
    void orc_halo_axis(OrcParams const* Pp, float* F, int ncomp, int axis, int lo, int up, int add)
    {
        Dom const D(*Pp);
        int const g = D.g[axis], n = D.n[axis];
        // transverse extent: the whole padded axis, except beyond an open outer boundary (the reference has no
        // edge / corner exchange across a missing neighbour: Mask::getRelativeDirections, Exchange.hpp:46-54)
        int tlo[3], thi[3];
        for(int d = 0; d < 3; ++d)
        {
            tlo[d] = Pp->open[d][0] ? D.g[d] : 0;
            thi[d] = Pp->open[d][1] ? D.g[d] + D.n[d] : D.N[d];
        }
        tlo[axis] = 0;
        thi[axis] = 1;
        for(int c = 0; c < ncomp; ++c)
        {
            f32* f = F + c * D.vol;
            for(int z = tlo[2]; z < thi[2]; ++z)
                for(int y = tlo[1]; y < thi[1]; ++y)
                    for(int x = tlo[0]; x < thi[0]; ++x)
                    {
                        auto at = [&](int pl) -> f32&
                        {
                            int q[3] = {x, y, z};
                            q[axis] = pl;
                            return f[D.idx(q[0], q[1], q[2])];
                        };
                        if(!add)
                        {
                            for(int w = 0; w < lo; ++w)
                                at(g - lo + w) = at(g + n - lo + w);
                            for(int w = 0; w < up; ++w)
                                at(g + n + w) = at(g + w);
                        }
                        else
                        {
                            for(int w = 0; w < lo; ++w)
                                at(g + n - lo + w) += at(g - lo + w);
                            for(int w = 0; w < up; ++w)
                                at(g + w) += at(g + n + w);
                        }
                    }
        }
    }

    /** UpdateBHalfFunctor over CORE+BORDER: B -= curlE * 0.5 * dt (FDTDBase.kernel:115-121) */
    void orc_update_b_half(OrcParams const* Pp, float const* E, float* B)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        Field3 const FE(const_cast<float*>(E), D.vol);
        Field3 FB(B, D.vol);
        LeheCoeff lc[3] = {leheCoeff(P, 0), leheCoeff(P, 1), leheCoeff(P, 2)};
        int const mode = P.solver == 1 ? 2 : 0;
#pragma omp parallel for schedule(static) collapse(2)
        for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
            for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                {
                    f32 cu[3];
                    curlAt(P, D, FE, mode, lc, x, y, z, cu);
                    int64_t const i = D.idx(x, y, z);
                    for(int c = 0; c < 3; ++c)
                        FB.c[c][i] -= cu[c] * 0.5f * P.dt;
                }
    }

    /** UpdateEFunctor over CORE+BORDER: E += curlB * c^2 * dt (FDTDBase.kernel:74-81) */
    void orc_update_e(OrcParams const* Pp, float* E, float const* B)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        Field3 FE(E, D.vol);
        Field3 const FB(const_cast<float*>(B), D.vol);
        LeheCoeff lc[3] = {leheCoeff(P, 0), leheCoeff(P, 1), leheCoeff(P, 2)};
        f32 const c2 = P.c * P.c;
#pragma omp parallel for schedule(static) collapse(2)
        for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
            for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                {
                    f32 cu[3];
                    curlAt(P, D, FB, 1, lc, x, y, z, cu);
                    int64_t const i = D.idx(x, y, z);
                    for(int c = 0; c < 3; ++c)
                        FE.c[c][i] += cu[c] * c2 * P.dt;
                }
    }

    /** KernelAddCurrentDensity + currentInterpolation::None over CORE+BORDER: E += coeff * J,
     * coeff = -(1/eps0) * dt (FDTD.hpp:84-85, None.hpp:60-64) */
    void orc_add_current(OrcParams const* Pp, float* E, float const* J)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        f32 const coeff = -(1.0f / P.eps0) * P.dt;
        if(P.current_interp == 1)
        {
            // currentInterpolation::Binomial<DIM3> (P/fields/currentInterpolation/Binomial.hpp:62-110); the J guards
            // hold the neighbours' border values (FieldJ "receive" exchange, P/fields/FieldJ.x.cpp:118-141)
            f32 const M = 8.0f, S = 4.0f, Dw = 2.0f, T = 1.0f;
            f32 const inverseDivisor = 1.0f / (M + 6.0f * S + 12.0f * Dw + 8.0f * T);
#pragma omp parallel for schedule(static) collapse(2)
            for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
                for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                    for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                        for(int c = 0; c < 3; ++c)
                        {
                            f32 const* j = J + c * D.vol;
                            auto at = [&](int dx, int dy, int dz) { return j[D.idx(x + dx, y + dy, z + dz)]; };
                            f32 far = at(-1, -1, -1) + at(+1, -1, -1);
                            far = far + at(-1, +1, -1);
                            far = far + at(+1, +1, -1);
                            far = far + at(-1, -1, +1);
                            far = far + at(+1, -1, +1);
                            far = far + at(-1, +1, +1);
                            far = far + at(+1, +1, +1);
                            f32 edge = at(-1, -1, 0) + at(+1, -1, 0);
                            edge = edge + at(-1, +1, 0);
                            edge = edge + at(+1, +1, 0);
                            edge = edge + at(-1, 0, -1);
                            edge = edge + at(+1, 0, -1);
                            edge = edge + at(-1, 0, +1);
                            edge = edge + at(+1, 0, +1);
                            edge = edge + at(0, -1, -1);
                            edge = edge + at(0, +1, -1);
                            edge = edge + at(0, -1, +1);
                            edge = edge + at(0, +1, +1);
                            f32 face = at(-1, 0, 0) + at(+1, 0, 0);
                            face = face + at(0, -1, 0);
                            face = face + at(0, +1, 0);
                            face = face + at(0, 0, -1);
                            face = face + at(0, 0, +1);
                            f32 avg = T * far + Dw * edge;
                            avg = avg + S * face;
                            avg = avg + M * at(0, 0, 0);
                            avg *= inverseDivisor;
                            E[c * D.vol + D.idx(x, y, z)] += coeff * avg;
                        }
            return;
        }
#pragma omp parallel for schedule(static) collapse(2)
        for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
            for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                {
                    int64_t const i = D.idx(x, y, z);
                    for(int c = 0; c < 3; ++c)
                        E[c * D.vol + i] += coeff * J[c * D.vol + i];
                }
    }

    /** ExponentialImpl::run (P/fields/absorber/exponential/Exponential.hpp:67-108) + KernelAbsorbBorder
     * (Exponential.kernel:45-118): for every face without neighbour, in exchange-type order RIGHT(+x), LEFT(-x),
     * BOTTOM(+y), TOP(-y), BACK(+z), FRONT(-z): each guard cell walks inwards in steps of the guard width and damps
     * field(cell) *= exp(-strength * factor) while factor > 0; factor = thickness-1 at the outermost active cell.
     * The transverse range of a face is CORE+BORDER (ExchangeMappingMethods<GUARD>, ExchangeMappingMethods.hpp:60-100). */
    void orc_absorb(OrcParams const* Pp, float* F)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        for(int axis = 0; axis < 3; ++axis)
            for(int side = 1; side >= 0; --side) // positive direction first (exchange types 1,3,9 are the + faces)
            {
                int const thickness = P.absorber_cells[axis][side];
                if(!P.open[axis][side] || thickness == 0)
                    continue;
                f32 const strength = P.absorber_strength[axis][side];
                int const rel = side ? 1 : -1;
                int lo[3] = {D.g[0], D.g[1], D.g[2]}, hi[3] = {D.g[0] + D.n[0], D.g[1] + D.n[1], D.g[2] + D.n[2]};
                // guard supercell layer of this face
                lo[axis] = side ? D.g[axis] + D.n[axis] : 0;
                hi[axis] = side ? D.N[axis] : D.g[axis];
                for(int c = 0; c < 3; ++c)
                {
                    f32* f = F + c * D.vol;
#pragma omp parallel for schedule(static) collapse(2)
                    for(int z = lo[2]; z < hi[2]; ++z)
                        for(int y = lo[1]; y < hi[1]; ++y)
                            for(int x = lo[0]; x < hi[0]; ++x)
                            {
                                int cell[3] = {x, y, z};
                                while(true)
                                {
                                    cell[axis] += D.g[axis] * -rel;
                                    int factor;
                                    if(rel < 0)
                                        factor = D.g[axis] - cell[axis] + thickness - 1;
                                    else
                                        factor = D.g[axis] + cell[axis] - D.N[axis] + thickness;
                                    if(factor <= 0)
                                        break;
                                    f32 const a = std::exp(-strength * f32(factor));
                                    int64_t const i = D.idx(cell[0], cell[1], cell[2]);
                                    f[i] = f[i] * a;
                                }
                            }
                }
            }
    }

    // ---------------------------------------------------------------------------------------------
    // Metrics (parity observables named by the north star)
    // ---------------------------------------------------------------------------------------------
    /** EnergyFields (P/plugins/EnergyFields.x.cpp:198-233): out[0]=B energy, out[1]=E energy (PIC units) */
    void orc_field_energy(OrcParams const* Pp, float const* E, float const* B, double* out2)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        double sB = 0, sE = 0;
#pragma omp parallel for schedule(static) reduction(+ : sB, sE)
        for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
            for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                {
                    int64_t const i = D.idx(x, y, z);
                    for(int c = 0; c < 3; ++c)
                    {
                        sB += double(B[c * D.vol + i]) * double(B[c * D.vol + i]);
                        sE += double(E[c * D.vol + i]) * double(E[c * D.vol + i]);
                    }
                }
        double const V = double(P.cell[0]) * double(P.cell[1]) * double(P.cell[2]);
        out2[0] = sB * (0.5 / double(P.mue0) * V);
        out2[1] = sE * (double(P.eps0) * V * 0.5);
    }

    /** EnergyParticles (P/plugins/EnergyParticles.x.cpp:100-131, P/algorithms/KinEnergy.hpp:38-68):
     * out[0] = kinetic energy, out[1] = total energy */
    void orc_particle_energy(OrcParams const* Pp, float massRatio, int64_t np, float const* mom, float const* w, double* out2)
    {
        OrcParams const& P = *Pp;
        double ek = 0, et = 0;
#pragma omp parallel for schedule(static) reduction(+ : ek, et)
        for(int64_t i = 0; i < np; ++i)
        {
            f32 const m[3] = {mom[i], mom[np + i], mom[2 * np + i]};
            f32 const mom2 = l2norm2(m);
            f32 const mass = (P.base_mass * massRatio) * w[i];
            f32 const c2 = P.c * P.c;
            f32 const gamma = gammaF(P, m, mass);
            f32 kin;
            if(gamma < 1.005f)
                kin = mom2 / (2.0f * mass);
            else
                kin = (gamma - 1.0f) * mass * c2;
            ek += double(kin);
            et += double(std::sqrt(mom2 + mass * mass * c2) * P.c);
        }
        out2[0] = ek;
        out2[1] = et;
    }

    /** ChargeDensity on the cell origins (P/particles/particleToGrid/ComputeGridValuePerFrame.hpp:60-134,
     * derivedAttributes/ChargeDensity.hpp): rho += charge/V * prod_d S(offset_d - pos_d); rho has guards. */
    void orc_charge_density(OrcParams const* Pp, float chargeRatio, float* rho, int64_t np, float const* pos, float const* w, int32_t const* cell)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        int const supp = shapeSupport(P.shape);
        int const lo = supp / 2, up = (supp + 1) / 2;
        f32 const V = P.cell[0] * P.cell[1] * P.cell[2];
        for(int64_t i = 0; i < np; ++i)
        {
            int const c = cell[i];
            int const cg[3] = {c % D.n[0] + D.g[0], (c / D.n[0]) % D.n[1] + D.g[1], c / (D.n[0] * D.n[1]) + D.g[2]};
            f32 const charge = (P.base_charge * chargeRatio) * w[i];
            f32 const attr = charge / V;
            for(int oz = -lo; oz <= up; ++oz)
                for(int oy = -lo; oy <= up; ++oy)
                    for(int ox = -lo; ox <= up; ++ox)
                    {
                        f32 assign = 1.0f;
                        assign *= shapeEval(P.shape, f32(ox) - pos[i]);
                        assign *= shapeEval(P.shape, f32(oy) - pos[np + i]);
                        assign *= shapeEval(P.shape, f32(oz) - pos[2 * np + i]);
                        rho[D.idx(cg[0] + ox, cg[1] + oy, cg[2] + oz)] += assign * attr;
                    }
        }
    }

    // ---------------------------------------------------------------------------------------------
    // PML absorber (convolutional PML, [Taflove, Hagness] ch. 7): P/fields/absorber/pml/Pml.kernel:60-160 (relative depth,
    // graded sigma / kappa / alpha, coefficients b and c), :420-476 (UpdateEFunctor), :520-582 (UpdateBHalfFunctor; psiB
    // is advanced in the first half update only, FDTDBase.hpp:200-211), Pml.hpp:120-150 (local thickness: zero at faces
    // with a neighbour).  Yee curls.  psi: six planes yx, zx, xy, zy, xz, yz over the padded grid.
    // ---------------------------------------------------------------------------------------------
    struct OrcPml
    {
        int thickness[3][2]; // local thickness in cells per [axis][negative, positive]
        float sigmaMax[3], kappaMax[3], alphaMax[3]; // NORMALIZED_SIGMA_MAX, KAPPA_MAX, NORMALIZED_ALPHA_MAX
        float sigmaKappaGradingOrder, alphaGradingOrder;
    };

    static inline float pml_relative_depth(float cellIdx, float nNeg, float nPos, int numLocalDomainCells, int numGuardCells)
    {
        float const zeroBasedIdx = cellIdx - float(numGuardCells);
        if(zeroBasedIdx < nNeg)
            return (nNeg - zeroBasedIdx) / nNeg;
        float const zeroBasedRightPMLStart = float(numLocalDomainCells - 2 * numGuardCells) - nPos;
        if(zeroBasedIdx > zeroBasedRightPMLStart)
            return (zeroBasedIdx - zeroBasedRightPMLStart) / nPos;
        return 0.0f;
    }

    struct PmlCoeff
    {
        float kappa[3], b[3], c[3];
        bool inPml;
    };

    static inline PmlCoeff pml_coefficients(Dom const& D, OrcPml const& M, float const idx[3], float dt)
    {
        PmlCoeff q;
        float prod = 1.0f;
        for(int d = 0; d < 3; ++d)
        {
            float sigma = 0.0f, alpha = 0.0f;
            q.kappa[d] = 1.0f;
            float const depth = pml_relative_depth(idx[d], float(M.thickness[d][0]), float(M.thickness[d][1]), D.N[d], D.g[d]);
            if(depth != 0.0f)
            {
                float const sk = std::pow(depth, M.sigmaKappaGradingOrder);
                sigma = M.sigmaMax[d] * sk;
                q.kappa[d] = 1.0f + (M.kappaMax[d] - 1.0f) * sk;
                float const ag = std::pow(1.0f - depth, M.alphaGradingOrder);
                alpha = M.alphaMax[d] * ag;
            }
            q.b[d] = std::exp(-(sigma / q.kappa[d] + alpha) * dt);
            q.c[d] = 0.0f;
            float const denominator = q.kappa[d] * (sigma + alpha * q.kappa[d]);
            if(denominator != 0.0f)
                q.c[d] = sigma * (q.b[d] - 1.0f) / denominator;
        }
        prod = q.b[0] * q.b[1];
        prod = prod * q.b[2];
        q.inPml = prod != 1.0f;
        return q;
    }

    /** updateE with the PML functor: E += curl B c^2 dt outside the PML, the convolutional update inside */
    void orc_update_e_pml(OrcParams const* Pp, OrcPml const* Mp, float* E, float const* B, float* psi)
    {
        OrcParams const& P = *Pp;
        OrcPml const& M = *Mp;
        Dom const D(P);
        float const c2 = P.c * P.c;
        float const c2dt = c2 * P.dt;
        int64_t const sy = D.N[0], sz = int64_t(D.N[0]) * D.N[1];
        float const *bx = B, *by = B + D.vol, *bz = B + 2 * D.vol;
        float *ex = E, *ey = E + D.vol, *ez = E + 2 * D.vol;
        float *pyx = psi, *pzx = psi + D.vol, *pxy = psi + 2 * D.vol, *pzy = psi + 3 * D.vol, *pxz = psi + 4 * D.vol, *pyz = psi + 5 * D.vol;
#pragma omp parallel for schedule(static) collapse(2)
        for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
            for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                {
                    int64_t const i = D.idx(x, y, z);
                    // backward differences (CurlB of the Yee solver), d<comp>d<axis>
                    float const dBzdy = (bz[i] - bz[i - sy]) / P.cell[1], dBydz = (by[i] - by[i - sz]) / P.cell[2];
                    float const dBxdz = (bx[i] - bx[i - sz]) / P.cell[2], dBzdx = (bz[i] - bz[i - 1]) / P.cell[0];
                    float const dBydx = (by[i] - by[i - 1]) / P.cell[0], dBxdy = (bx[i] - bx[i - sy]) / P.cell[1];
                    float const idx[3] = {float(x), float(y), float(z)};
                    PmlCoeff const q = pml_coefficients(D, M, idx, P.dt);
                    if(q.inPml)
                    {
                        pyx[i] = q.b[0] * pyx[i] + q.c[0] * dBzdx;
                        pzx[i] = q.b[0] * pzx[i] + q.c[0] * dBydx;
                        pxy[i] = q.b[1] * pxy[i] + q.c[1] * dBzdy;
                        pzy[i] = q.b[1] * pzy[i] + q.c[1] * dBxdy;
                        pxz[i] = q.b[2] * pxz[i] + q.c[2] * dBydz;
                        pyz[i] = q.b[2] * pyz[i] + q.c[2] * dBxdz;
                        ex[i] += c2dt * (dBzdy / q.kappa[1] - dBydz / q.kappa[2] + pxy[i] - pxz[i]);
                        ey[i] += c2dt * (dBxdz / q.kappa[2] - dBzdx / q.kappa[0] + pyz[i] - pyx[i]);
                        ez[i] += c2dt * (dBydx / q.kappa[0] - dBxdy / q.kappa[1] + pzx[i] - pzy[i]);
                    }
                    else
                    {
                        ex[i] += (dBzdy - dBydz) * c2 * P.dt;
                        ey[i] += (dBxdz - dBzdx) * c2 * P.dt;
                        ez[i] += (dBydx - dBxdy) * c2 * P.dt;
                    }
                }
    }

    /** updateBHalf with the PML functor; updatePsi: first half update of the step (FDTDBase.hpp:200-211) */
    void orc_update_b_half_pml(OrcParams const* Pp, OrcPml const* Mp, float const* E, float* B, float* psi, int updatePsi)
    {
        OrcParams const& P = *Pp;
        OrcPml const& M = *Mp;
        Dom const D(P);
        float const halfDt = 0.5f * P.dt;
        int64_t const sy = D.N[0], sz = int64_t(D.N[0]) * D.N[1];
        float const *ex = E, *ey = E + D.vol, *ez = E + 2 * D.vol;
        float *bx = B, *by = B + D.vol, *bz = B + 2 * D.vol;
        float *pyx = psi, *pzx = psi + D.vol, *pxy = psi + 2 * D.vol, *pzy = psi + 3 * D.vol, *pxz = psi + 4 * D.vol, *pyz = psi + 5 * D.vol;
#pragma omp parallel for schedule(static) collapse(2)
        for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
            for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                {
                    int64_t const i = D.idx(x, y, z);
                    // forward differences (CurlE of the Yee solver)
                    float const dEzdy = (ez[i + sy] - ez[i]) / P.cell[1], dEydz = (ey[i + sz] - ey[i]) / P.cell[2];
                    float const dExdz = (ex[i + sz] - ex[i]) / P.cell[2], dEzdx = (ez[i + 1] - ez[i]) / P.cell[0];
                    float const dEydx = (ey[i + 1] - ey[i]) / P.cell[0], dExdy = (ex[i + sy] - ex[i]) / P.cell[1];
                    float const idx[3] = {0.5f + float(x), 0.5f + float(y), 0.5f + float(z)};
                    PmlCoeff const q = pml_coefficients(D, M, idx, P.dt);
                    if(q.inPml)
                    {
                        if(updatePsi)
                        {
                            pyx[i] = q.b[0] * pyx[i] + q.c[0] * dEzdx;
                            pzx[i] = q.b[0] * pzx[i] + q.c[0] * dEydx;
                            pxy[i] = q.b[1] * pxy[i] + q.c[1] * dEzdy;
                            pzy[i] = q.b[1] * pzy[i] + q.c[1] * dExdy;
                            pxz[i] = q.b[2] * pxz[i] + q.c[2] * dEydz;
                            pyz[i] = q.b[2] * pyz[i] + q.c[2] * dExdz;
                        }
                        bx[i] += halfDt * (dEydz / q.kappa[2] - dEzdy / q.kappa[1] + pxz[i] - pxy[i]);
                        by[i] += halfDt * (dEzdx / q.kappa[0] - dExdz / q.kappa[2] + pyx[i] - pyz[i]);
                        bz[i] += halfDt * (dExdy / q.kappa[1] - dEydx / q.kappa[0] + pzy[i] - pzx[i]);
                    }
                    else
                    {
                        bx[i] -= (dEzdy - dEydz) * halfDt;
                        by[i] -= (dExdz - dEzdx) * halfDt;
                        bz[i] -= (dEydx - dExdy) * halfDt;
                    }
                }
    }

    // ---------------------------------------------------------------------------------------------
    // Incident field (laser) through the YMin Huygens surface: P/fields/incidentField/Solver.hpp:190-395 (updateField),
    // Solver.kernel:101-404 (UpdateFunctor for the Yee solver: margin 1, one derivative coefficient = 1; kernel
    // :418-476 with the "last updated cell" rule), Functors.hpp (BaseFunctorE: getFocus / getOrigin / getCurrentTime /
    // getInternalCoordinates, BaseSeparableFunctorE::operator(), ApproximateIncidentB), profiles/PlaneWave.hpp:93-130
    // (getLongitudinal), profiles/GaussianPulse.hpp:186-346 (getValue with Laguerre modes, pulse-front tilt and the
    // GaussianPulseEnvelope), calculatePhaseVelocity.hpp + DispersionRelationSolver (Yee, propagation along y).
    // The surface spans a periodic transversal axis completely for the PlaneWave profile only
    // (MakePeriodicTransversalHuygensSurfaceContiguous, PlaneWave.def:71); otherwise it ends at POSITION.
    // ---------------------------------------------------------------------------------------------
    constexpr int ORC_LASER_MAX_MODES = 8;

    struct OrcLaser
    {
        int polarisation; // PolarisationType: 0 Linear, 1 Circular
        int offset_ymin; // POSITION[1][0]
        float amplitude, omega, pulse_duration, nofocus_constant, ramp_init, phase; // *Unitless
        float pol[3]; // POLARISATION_DIRECTION (unit, orthogonal to y)
        float time_delay; // TIME_DELAY
        int global_y_offset; // totalCellOffset[1] of this domain
        int profile; // 0 PlaneWave, 1 GaussianPulse<Params, GaussianPulseEnvelope> (with tilt: PulseFrontTilt), 2 Wavepacket, 3 Polynom, 4 ExpRampWithPrepulse
        int position[3][2]; // POSITION[axis][min, max] (max <= 0: counted from the upper boundary)
        int global_size[3]; // global domain cells
        int periodic[3];
        float w0, wave_length, time_shift; // W0, WAVE_LENGTH, GaussianPulseEnvelope::TIME_SHIFT
        float focus_position[3]; // FOCUS_POSITION_{X,Y,Z}
        int focus_origin_center[3]; // FOCUS_ORIGIN_* == Origin::Center
        float tilt[2]; // TILT_AXIS_1, TILT_AXIS_2 in radian
        int n_modes; // laguerreModes.size()
        float modes[ORC_LASER_MAX_MODES], mode_phases[ORC_LASER_MAX_MODES];
        // separable profiles with a Gaussian transversal envelope (BaseTransversalGaussianParamUnitless): profile 2
        // Wavepacket, 3 Polynom, 4 ExpRampWithPrepulse
        float w0_axis[2]; // W0_AXIS_1, W0_AXIS_2
        // Wavepacket: [0] INIT_TIME.  ExpRampWithPrepulse: [0] time_start_init, [1] TIME_PREPULSE, [2] TIME_PEAKPULSE,
        // [3..5] TIME_1..3, [6] PREPULSE_DURATION, [7] INT_RATIO_PREPULSE, [8..10] INT_RATIO_POINT_1..3
        float profile_params[16];
    };

    static float orc_laser_phase_velocity(OrcParams const& P, OrcLaser const& L)
    {
        // Yee dispersion relation along y: sin(w dt/2)/(c dt) = sin(k dy/2)/dy (DispersionRelationSolver.hpp), in fp64
        double const w = double(L.omega), dt = double(P.dt), c = double(P.c), dy = double(P.cell[1]);
        double const k = 2.0 / dy * std::asin(dy * std::sin(0.5 * w * dt) / (c * dt));
        return float(w / k / c);
    }

    // BaseFunctorE::getFocus / getOrigin (Functors.hpp:267-348) for DIR = (0, 1, 0)
    struct OrcLaserFrame
    {
        float focus[3], origin[3], axis1[3], axis2[3];
        float phaseVelocity;
    };

    static OrcLaserFrame orc_laser_frame(OrcParams const& P, OrcLaser const& L)
    {
        OrcLaserFrame F;
        float const direction[3] = {0.0f, 1.0f, 0.0f};
        for(int d = 0; d < 3; ++d)
        {
            F.focus[d] = L.focus_position[d];
            if(L.focus_origin_center[d])
                F.focus[d] += float(unsigned(L.global_size[d]) / 2u) * P.cell[d];
        }
        float originP = -std::numeric_limits<float>::infinity();
        for(int axis = 0; axis < 3; ++axis)
            if(std::abs(direction[axis]) > std::numeric_limits<float>::epsilon())
            {
                float const minPosition = (float(L.position[axis][0]) + 0.75f) * P.cell[axis];
                int const maxPositionIdx = (L.position[axis][1] > 0) ? L.position[axis][1] : L.global_size[axis] + L.position[axis][1];
                float const maxPosition = (float(maxPositionIdx) - 0.75f) * P.cell[axis];
                float const axisP = std::min((minPosition - F.focus[axis]) / direction[axis], (maxPosition - F.focus[axis]) / direction[axis]);
                originP = std::max(originP, axisP);
            }
        for(int d = 0; d < 3; ++d)
            F.origin[d] = F.focus[d] + originP * direction[d];
        for(int d = 0; d < 3; ++d)
            F.axis1[d] = L.pol[d];
        // getAxis2(): cross(DIR, POL_DIR)
        F.axis2[0] = direction[1] * L.pol[2] - direction[2] * L.pol[1];
        F.axis2[1] = direction[2] * L.pol[0] - direction[0] * L.pol[2];
        F.axis2[2] = direction[0] * L.pol[1] - direction[1] * L.pol[0];
        F.phaseVelocity = orc_laser_phase_velocity(P, L);
        return F;
    }

    static inline float orc_dot3(float const a[3], float const b[3])
    {
        float tmp = a[0] * b[0]; // pmacc Dot: Vector.tpp:87-94
        tmp += a[1] * b[1];
        tmp += a[2] * b[2];
        return tmp;
    }

    // PlaneWaveFunctorIncidentE::getLongitudinal (profiles/PlaneWave.hpp:93-130)
    static float orc_laser_longitudinal(OrcLaser const& L, float time, float phaseShift)
    {
        float envelope = L.amplitude;
        float const mue = 0.5f * L.ramp_init * L.pulse_duration;
        float const tau = L.pulse_duration * std::sqrt(2.0f);
        float const endUpramp = mue;
        float const startDownramp = mue + L.nofocus_constant;
        float integrationCorrectionFactor = 0.0f;
        if(time > startDownramp)
        {
            float const exponent = (time - startDownramp) / tau;
            envelope *= std::exp(-0.5f * exponent * exponent);
            integrationCorrectionFactor = (time - startDownramp) / (L.omega * tau * tau);
        }
        else if(time < endUpramp)
        {
            float const exponent = (time - endUpramp) / tau;
            envelope *= std::exp(-0.5f * exponent * exponent);
            integrationCorrectionFactor = (time - endUpramp) / (L.omega * tau * tau);
        }
        float const timeOszi = time - endUpramp;
        float const phase = L.omega * timeOszi + L.phase + phaseShift;
        return (std::sin(phase) + std::cos(phase) * integrationCorrectionFactor) * envelope;
    }

    // WavepacketFunctorIncidentE::getLongitudinal (profiles/Wavepacket.hpp:122-151)
    static float orc_wavepacket_longitudinal(OrcLaser const& L, float time, float phaseShift)
    {
        float const endUpramp = -0.5f * L.nofocus_constant, startDownramp = 0.5f * L.nofocus_constant;
        float const mue = 0.5f * L.profile_params[0];
        float const runTime = time - mue;
        float const tau = L.pulse_duration * std::sqrt(2.0f);
        float envelope = L.amplitude;
        float correctionFactor = 0.0f;
        if(runTime > startDownramp)
        {
            float const exponent = ((runTime - startDownramp) / L.pulse_duration / std::sqrt(2.0f));
            envelope *= std::exp(-0.5f * exponent * exponent);
            correctionFactor = (runTime - startDownramp) / (tau * tau * L.omega);
        }
        else if(runTime < endUpramp)
        {
            float const exponent = ((runTime - endUpramp) / L.pulse_duration / std::sqrt(2.0f));
            envelope *= std::exp(-0.5f * exponent * exponent);
            correctionFactor = (runTime - endUpramp) / (tau * tau * L.omega);
        }
        float const phase = L.omega * runTime + L.phase + phaseShift;
        return (std::sin(phase) + correctionFactor * std::cos(phase)) * envelope;
    }

    // PolynomFunctorIncidentE::getLongitudinal / polynomial (profiles/Polynom.hpp:112-136)
    static float orc_polynom_longitudinal(OrcLaser const& L, float time, float phaseShift)
    {
        float const riseTime = 0.5f * L.pulse_duration;
        float const tau = time / riseTime;
        float const phase = L.omega * (time - riseTime) + L.phase + phaseShift;
        float result = 0.0f;
        if(tau >= 0.0f && tau <= 1.0f)
            result = tau * tau * tau * (10.0f - 15.0f * tau + 6.0f * tau * tau);
        else if(tau > 1.0f && tau <= 2.0f)
            result = (2.0f - tau) * (2.0f - tau) * (2.0f - tau) * (4.0f - 9.0f * tau + 6.0f * tau * tau);
        float const amplitude = L.amplitude * result;
        return std::sin(phase) * amplitude;
    }

    // ExpRampWithPrepulseLongitudinal::getEnvelope + ExpRampWithPrepulseFunctorIncidentE::getLongitudinal
    // (profiles/ExpRampWithPrepulse.hpp:157-290)
    static float orc_exp_ramp_longitudinal(OrcLaser const& L, float time, float phaseShift)
    {
        float const* q = L.profile_params;
        float const time_start_init = q[0], TIME_PREPULSE = q[1], TIME_PEAKPULSE = q[2], TIME_1 = q[3], TIME_2 = q[4], TIME_3 = q[5];
        float const PREPULSE_DURATION = q[6];
        float const endUpramp = TIME_PEAKPULSE - 0.5f * L.nofocus_constant, startDownramp = TIME_PEAKPULSE + 0.5f * L.nofocus_constant;
        auto gauss = [](float t, float pulseDuration) {
            float const exponent = t / pulseDuration;
            return std::exp(-0.25f * exponent * exponent);
        };
        auto extrapolateExpo = [](float t1, float a1, float t2, float a2, float t) {
            float const log1 = (t2 - t) * std::log(a1);
            float const log2 = (t - t1) * std::log(a2);
            return std::exp((log1 + log2) / (t2 - t1));
        };
        float const runTime = time + time_start_init;
        float const phase = L.omega * runTime + L.phase + phaseShift;
        float const AMP_PREPULSE = std::sqrt(q[7]), AMP_1 = std::sqrt(q[8]), AMP_2 = std::sqrt(q[9]), AMP_3 = std::sqrt(q[10]);
        float env = 0.0f;
        bool const before_preupramp = runTime < time_start_init;
        bool const before_start = runTime < TIME_1;
        bool const before_peakpulse = runTime < endUpramp;
        bool const during_first_exp = (TIME_1 < runTime) && (runTime < TIME_2);
        bool const after_peakpulse = startDownramp <= runTime;
        if(before_preupramp)
            env = 0.0f;
        else if(before_start)
            env = AMP_1 * gauss(runTime - TIME_1, L.pulse_duration);
        else if(before_peakpulse)
        {
            float const ramp_when_peakpulse = extrapolateExpo(TIME_2, AMP_2, TIME_3, AMP_3, endUpramp);
            env += (1.0f - ramp_when_peakpulse) * gauss(runTime - endUpramp, L.pulse_duration);
            env += AMP_PREPULSE * gauss(runTime - TIME_PREPULSE, PREPULSE_DURATION);
            if(during_first_exp)
                env += extrapolateExpo(TIME_1, AMP_1, TIME_2, AMP_2, runTime);
            else
                env += extrapolateExpo(TIME_2, AMP_2, TIME_3, AMP_3, runTime);
        }
        else if(!after_peakpulse)
            env = 1.0f;
        else
            env = gauss(runTime - startDownramp, L.pulse_duration);
        return std::cos(phase) * L.amplitude * env;
    }

    static float orc_separable_longitudinal(OrcLaser const& L, float time, float phaseShift)
    {
        switch(L.profile)
        {
        case 2:
            return orc_wavepacket_longitudinal(L, time, phaseShift);
        case 3:
            return orc_polynom_longitudinal(L, time, phaseShift);
        case 4:
            return orc_exp_ramp_longitudinal(L, time, phaseShift);
        default:
            return orc_laser_longitudinal(L, time, phaseShift);
        }
    }

    // GaussianPulseFunctorIncidentE::simpleLaguerre (GaussianPulse.hpp:316-336)
    static float orc_simple_laguerre(unsigned n, float x)
    {
        if(n == 0)
            return 1.0f;
        unsigned currentN = 1;
        float laguerreNMinus1 = 1.0f;
        float laguerreN = 1.0f - x;
        float laguerreNPlus1 = 0.0f;
        while(currentN < n)
        {
            laguerreNPlus1 = ((2.0f * float(currentN) + 1.0f - x) * laguerreN - float(currentN) * laguerreNMinus1) / float(currentN + 1u);
            laguerreNMinus1 = laguerreN;
            laguerreN = laguerreNPlus1;
            currentN++;
        }
        return laguerreN;
    }

    // GaussianPulseFunctorIncidentE::getValue (GaussianPulse.hpp:208-308), `time` = getCurrentTime() >= 0, 3D
    static float orc_gaussian_pulse_value(OrcParams const& P, OrcLaser const& L, OrcLaserFrame const& F, float posIn[3], float time, float phaseShift)
    {
        float const pi = 3.14159265358979323846f;
        float const rayleighLength = pi * L.w0 * L.w0 / L.wave_length;
        float pos[3] = {posIn[0], posIn[1], posIn[2]};
        time += L.time_shift;
        float const focusRelativeToOrigin[3] = {F.focus[0] - F.origin[0], F.focus[1] - F.origin[1], F.focus[2] - F.origin[2]};
        float const axis0[3] = {0.0f, 1.0f, 0.0f};
        float const distanceFocusRelativeToOrigin = orc_dot3(focusRelativeToOrigin, axis0);
        float const focusPos = distanceFocusRelativeToOrigin - pos[0];
        float const w = L.w0 * std::sqrt(1.0f + (focusPos / rayleighLength) * (focusPos / rayleighLength));
        float const phase = L.omega * (time - focusPos / P.c) + L.phase + phaseShift;
        if(L.tilt[0] != 0.0f || L.tilt[1] != 0.0f)
        {
            float const tiltTimeShift = phase / L.omega + focusPos / P.c;
            float const tiltPositionShift = P.c * tiltTimeShift / orc_dot3(axis0, P.cell);
            pos[1] += std::tan(L.tilt[0]) * tiltPositionShift;
            pos[2] += std::tan(L.tilt[1]) * tiltPositionShift;
        }
        float const planeNoNormal[3] = {0.0f, 1.0f, 1.0f};
        float const q[3] = {pos[0] * planeNoNormal[0], pos[1] * planeNoNormal[1], pos[2] * planeNoNormal[2]};
        float transversalDistanceSquared = q[0] * q[0]; // l2norm2: Vector.tpp:104-110
        transversalDistanceSquared += q[1] * q[1];
        transversalDistanceSquared += q[2] * q[2];
        float const R_inv = -focusPos / (rayleighLength * rayleighLength + focusPos * focusPos);
        float const xi = std::atan(-focusPos / rayleighLength);
        float etrans = 0.0f;
        float const r2OverW2 = transversalDistanceSquared / w / w;
        float const r = 0.5f * transversalDistanceSquared * R_inv;
        float const twoPi = 6.28318530717958647692f; // Pi<float_X>::doubleValue
        for(int m = 0; m < L.n_modes; ++m)
            etrans += L.modes[m] * orc_simple_laguerre(unsigned(m), 2.0f * r2OverW2) * std::exp(-r2OverW2)
                * std::cos(twoPi / L.wave_length * focusPos - twoPi / L.wave_length * r + (2.0f * float(m) + 1.0f) * xi + phase + L.mode_phases[m]);
        float const shiftedTime = time - r / P.c;
        // GaussianPulseEnvelope::getEnvelope (GaussianPulse.hpp:343-348)
        float const exponent = shiftedTime / (2.0f * L.pulse_duration);
        etrans *= std::exp(-exponent * exponent);
        float etrans_norm = 0.0f;
        for(int m = 0; m < L.n_modes; ++m)
            etrans_norm += L.modes[m];
        float envelope = L.amplitude;
        envelope *= L.w0 / w;
        return envelope * etrans / etrans_norm;
    }

    // incident E at a (fractional) total cell index: BaseFunctorE::getCurrentTime / getInternalCoordinates +
    // BaseSeparableFunctorE::operator() (PlaneWave) or GaussianPulseFunctorIncidentE::operator()
    static void orc_laser_incident_e(OrcParams const& P, OrcLaser const& L, OrcLaserFrame const& F, float currentStep, float const idx[3], float out[3])
    {
        float const axis0[3] = {0.0f, 1.0f, 0.0f};
        float const shiftFromOrigin[3] = {idx[0] * P.cell[0] - F.origin[0], idx[1] * P.cell[1] - F.origin[1], idx[2] * P.cell[2] - F.origin[2]};
        float const distance = orc_dot3(shiftFromOrigin, axis0);
        float const timeDelay = distance / F.phaseVelocity + L.time_delay;
        float const time = currentStep * P.dt - timeDelay;
        out[0] = out[1] = out[2] = 0.0f;
        if(time < 0.0f)
            return;
        float a, b; // value with phase shift pi/2 (circular only) and 0
        if(L.profile != 1)
        {
            // BaseSeparableFunctorE::operator() (Functors.hpp:452-470); PlaneWave: transversal 1,
            // BaseSeparableTransversalGaussianFunctorE::getTransversal (:525-532) otherwise
            float transversal = 1.0f;
            if(L.profile != 0)
            {
                float const internalPosition[3] = {0.0f, orc_dot3(shiftFromOrigin, F.axis1), orc_dot3(shiftFromOrigin, F.axis2)};
                float const w0[3] = {1.0f, L.w0_axis[0], L.w0_axis[1]};
                float const r[3] = {internalPosition[0] / w0[0], internalPosition[1] / w0[1], internalPosition[2] / w0[2]};
                float r2 = r[0] * r[0];
                r2 += r[1] * r[1];
                r2 += r[2] * r[2];
                transversal = std::exp(-r2);
            }
            a = L.polarisation ? orc_separable_longitudinal(L, time, 1.57079632679489661923f) * transversal : 0.0f;
            b = orc_separable_longitudinal(L, time, 0.0f) * transversal;
        }
        else
        {
            float pos[3] = {orc_dot3(shiftFromOrigin, axis0), orc_dot3(shiftFromOrigin, F.axis1), orc_dot3(shiftFromOrigin, F.axis2)};
            a = L.polarisation ? orc_gaussian_pulse_value(P, L, F, pos, time, 1.57079632679489661923f) : 0.0f;
            b = orc_gaussian_pulse_value(P, L, F, pos, time, 0.0f);
        }
        if(L.polarisation == 0)
        {
            for(int d = 0; d < 3; ++d)
                out[d] = L.pol[d] * b;
        }
        else
        {
            float const rs2 = std::sqrt(2.0f);
            float const p1[3] = {L.pol[0] / rs2, L.pol[1] / rs2, L.pol[2] / rs2};
            // cross(axis0 = (0,1,0), p1)
            float const p2[3] = {1.0f * p1[2] - 0.0f * p1[1], 0.0f * p1[0] - 0.0f * p1[2], 0.0f * p1[1] - 1.0f * p1[0]};
            for(int d = 0; d < 3; ++d)
                out[d] = p1[d] * a + p2[d] * b;
        }
    }

    /** incidentField::Solver::updateE (updatedIsE = 1, uses B_inc = cross(dir, E_inc) / c) or ::updateBHalf (0, uses
     * E_inc) for a profile on YMin.  currentStep is fractional (FDTDBase.hpp:108-117,161-166).  The domain of this
     * oracle spans x and z completely (it is the last one along both). */
    void orc_incident_update(OrcParams const* Pp, OrcLaser const* Lp, float* F, int updatedIsE, float currentStep)
    {
        OrcParams const& P = *Pp;
        OrcLaser const& L = *Lp;
        Dom const D(P);
        // Solver.hpp:230-236: the updated plane in user (total) coordinates; E sits in the total-field region
        int const planeTotal = L.position[1][0] + 1 - (updatedIsE ? 0 : 1);
        int const yl = planeTotal - L.global_y_offset; // local, without guards
        if(yl < 0 || yl >= D.n[1])
            return;
        // transversal extent and "last updated cell" flags (Solver.hpp:209-258, Solver.kernel:458-466)
        int lo[3], hi[3];
        bool lastDomain[3];
        for(int d = 0; d < 3; d += 2)
        {
            int begin = L.position[d][0] + 1;
            int end = (L.position[d][1] > 0) ? L.position[d][1] : L.global_size[d] + L.position[d][1];
            lastDomain[d] = true;
            if(L.profile == 0 && L.periodic[d])
            {
                begin = 0;
                end = L.global_size[d];
                lastDomain[d] = false;
            }
            lo[d] = std::max(begin, 0);
            hi[d] = std::min(end, D.n[d]);
        }
        if(lo[0] >= hi[0] || lo[2] >= hi[2])
            return;
        OrcLaserFrame const frame = orc_laser_frame(P, L);
        float const c2 = P.c * P.c;
        float const curlCoefficient = updatedIsE ? P.dt * c2 : -(0.5f * P.dt);
        float const baseCoefficient = curlCoefficient / P.cell[1] * 1.0f; // direction +1
        // in-cell shifts (Solver.hpp:360-372): base shift -1 (E updated) / +1 (B updated) along y plus the Yee position of
        // the incident component: incident component 1 = x, 2 = z (Solver.kernel:222-223 for axis y)
        float const baseShift = updatedIsE ? -1.0f : 1.0f;
        float shift1[3], shift2[3];
        if(updatedIsE)
        {
            // incident field B: Bx at (0, .5, .5), Bz at (.5, .5, 0)   (YeeCell.hpp:70-130)
            shift1[0] = 0.0f, shift1[1] = baseShift + 0.5f, shift1[2] = 0.5f;
            shift2[0] = 0.5f, shift2[1] = baseShift + 0.5f, shift2[2] = 0.0f;
        }
        else
        {
            // incident field E: Ex at (.5, 0, 0), Ez at (0, 0, .5)
            shift1[0] = 0.5f, shift1[1] = baseShift + 0.0f, shift1[2] = 0.0f;
            shift2[0] = 0.0f, shift2[1] = baseShift + 0.0f, shift2[2] = 0.5f;
        }
        float* fx = F;
        float* fz = F + 2 * D.vol;
        int const y = yl + D.g[1];
#pragma omp parallel for schedule(static)
        for(int zl = lo[2]; zl < hi[2]; ++zl)
            for(int xl = lo[0]; xl < hi[0]; ++xl)
            {
                bool const lastX = lastDomain[0] && xl == hi[0] - 1, lastZ = lastDomain[2] && zl == hi[2] - 1;
                // Solver.kernel:318-325 with incidentComponent1 = x, incidentComponent2 = z
                bool const apply1 = updatedIsE ? !lastZ : !lastX;
                bool const apply2 = updatedIsE ? !lastX : !lastZ;
                // total cell index of the Huygens surface cell = the updated cell (margin 1)
                float const base[3] = {float(xl), float(planeTotal), float(zl)};
                float const i1[3] = {base[0] + shift1[0], base[1] + shift1[1], base[2] + shift1[2]};
                float const i2[3] = {base[0] + shift2[0], base[1] + shift2[1], base[2] + shift2[2]};
                float e1[3], e2[3], inc1, inc2;
                orc_laser_incident_e(P, L, frame, currentStep, i1, e1);
                orc_laser_incident_e(P, L, frame, currentStep, i2, e2);
                if(updatedIsE)
                {
                    // ApproximateIncidentB: cross((0,1,0), E) / c = (Ez, 0, -Ex) / c
                    inc1 = (1.0f * e1[2] - 0.0f * e1[1]) / P.c; // B_inc,x
                    inc2 = (0.0f * e2[1] - 1.0f * e2[0]) / P.c; // B_inc,z
                }
                else
                {
                    inc1 = e1[0];
                    inc2 = e2[2];
                }
                // Solver.kernel:318-337: result[dir1 = z] = +base * inc1, result[dir2 = x] = -base * inc2
                float rz = 0.0f, rx = 0.0f;
                if(apply1)
                    rz += 1.0f * inc1;
                if(apply2)
                    rx += 1.0f * inc2;
                rz *= baseCoefficient;
                rx *= -baseCoefficient;
                int64_t const i = D.idx(xl + D.g[0], y, zl + D.g[2]);
                fz[i] += rz;
                fx[i] += rx;
            }
    }

    /** ChargeConservation (P/plugins/ChargeConservation.tpp:122-136,205-259): max |div E * eps0 - rho| * V.
     * rho must already be guard-reduced. */
    double orc_gauss_residual(OrcParams const* Pp, float const* E, float const* rho)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        f32 const rw = 1.0f / P.cell[0], rh = 1.0f / P.cell[1], rd = 1.0f / P.cell[2];
        f32 mx = 0.0f;
        for(int z = D.g[2]; z < D.g[2] + D.n[2]; ++z)
            for(int y = D.g[1]; y < D.g[1] + D.n[1]; ++y)
                for(int x = D.g[0]; x < D.g[0] + D.n[0]; ++x)
                {
                    int64_t const i = D.idx(x, y, z);
                    f32 const div = (E[i] - E[D.idx(x - 1, y, z)]) * rw + (E[D.vol + i] - E[D.vol + D.idx(x, y - 1, z)]) * rh
                        + (E[2 * D.vol + i] - E[2 * D.vol + D.idx(x, y, z - 1)]) * rd;
                    f32 const dev = std::fabs(div * P.eps0 - rho[i]);
                    mx = std::max(mx, dev);
                }
        return double(mx * (P.cell[0] * P.cell[1] * P.cell[2]));
    }

    // ---------------------------------------------------------------------------------------------
    // Whole step for a single periodic domain: Simulation::runOneStep (P/simulation/control/Simulation.hpp:522-542)
    // ---------------------------------------------------------------------------------------------
    struct OrcSpecies
    {
        float massRatio, chargeRatio;
        int64_t np;
        float *pos, *mom, *w;
        int32_t* cell;
    };

    void orc_step(OrcParams const* Pp, float* E, float* B, float* J, int nSpecies, OrcSpecies* sp)
    {
        OrcParams const& P = *Pp;
        Dom const D(P);
        std::memset(J, 0, sizeof(float) * 3 * size_t(D.vol)); // CurrentReset (stage/CurrentReset.hpp:45-52)
        for(int s = 0; s < nSpecies; ++s) // ParticlePush (stage/ParticlePush.x.cpp:141-148)
            orc_push(Pp, sp[s].massRatio, sp[s].chargeRatio, E, B, sp[s].np, sp[s].pos, sp[s].mom, sp[s].w, sp[s].cell, nullptr, nullptr);
        // update_beforeCurrent (FDTDBase.hpp:97-121)
        orc_update_b_half(Pp, E, B);
        orc_guard_copy(Pp, B);
        orc_update_e(Pp, E, B);
        for(int s = 0; s < nSpecies; ++s) // CurrentDeposition (stage/CurrentDeposition.x.cpp:105-116)
            orc_deposit(Pp, sp[s].massRatio, sp[s].chargeRatio, J, sp[s].np, sp[s].pos, sp[s].mom, sp[s].w, sp[s].cell);
        orc_guard_add(Pp, J); // FieldJ::asyncCommunication (stage/CurrentInterpolationAndAdditionToEMF.hpp:99-148)
        if(P.current_interp == 1)
            orc_guard_copy(Pp, J); // fieldJrecv: GUARD := neighbour BORDER (FieldJ.x.cpp:118-141); all axes periodic here
        orc_add_current(Pp, E, J);
        // update_afterCurrent (FDTDBase.hpp:151-183)
        orc_guard_copy(Pp, E);
        orc_update_b_half(Pp, E, B);
        orc_guard_copy(Pp, B);
    }

    // ---------------------------------------------------------------------------------------------
    // KelvinHelmholtz initial condition (share/picongpu/examples/KelvinHelmholtz/include/picongpu/param/*):
    // Homogenous density, Quiet start 5x5x1 (QuietImpl.hpp:47-118, filled from the highest lattice index down),
    // ions derived from electrons, drift gamma=1.021 along +-x by global y quarter (Drift.hpp:56-80,
    // particleFilters.param), electron temperature 0.0005 keV (Temperature.hpp:63-87).
    // RNG: Philox4x32-10, key=(seed, 0), counter=(global particle index lo, hi, 0, 0); 3 normals per
    // particle via Box-Muller on (r0,r1) and (r2,r3). The reference's own RNG stream differs per backend
    // (P/param/random.param), so ICs are generated here once and shared by oracle and GPU.
    // ---------------------------------------------------------------------------------------------
    /** @param globalN global grid cells, offset of this domain in the global grid `globalOff`
     *  @param ppcDim particles per cell per dimension (5,5,1)
     *  @param ev2joule_pic sim.pic.conv().eV2Joule(1.0) in PIC energy units
     *  arrays sized n_cells*ppc; species 0 = electrons, 1 = ions */
    void orc_khi_init(
        OrcParams const* Pp,
        int const* globalN,
        int const* globalOff,
        int const* ppcDim,
        float realParticlesPerCell,
        float massRatioIon,
        double gammaDrift,
        double temperature_keV,
        double eV_pic,
        uint32_t seed,
        float* posE,
        float* momE,
        float* wE,
        int32_t* cellE,
        float* posI,
        float* momI,
        float* wI,
        int32_t* cellI)
    {
        OrcParams const& P = *Pp;
        int const ppc = ppcDim[0] * ppcDim[1] * ppcDim[2];
        int64_t const ncell = int64_t(P.n[0]) * P.n[1] * P.n[2];
        int64_t const np = ncell * ppc;
        f32 const weighting = realParticlesPerCell / f32(ppc);
        f32 spacing[3];
        for(int d = 0; d < 3; ++d)
            spacing[d] = 1.0f / f32(ppcDim[d]);
        double const beta = std::sqrt(1.0 - 1.0 / (gammaDrift * gammaDrift));
        f32 const massE = (P.base_mass * 1.0f) * weighting;
        f32 const massI = (P.base_mass * massRatioIon) * weighting;
        f32 const driftE = f32(gammaDrift * beta * double(massE) * double(P.c));
        f32 const driftI = f32(gammaDrift * beta * double(massI) * double(P.c));
        f32 const energy = f32(eV_pic * (temperature_keV * 1.0e3));
        f32 const macroEnergy = weighting * energy;
        f32 const stddev = std::sqrt(macroEnergy * massE);
#pragma omp parallel for schedule(static)
        for(int64_t c = 0; c < ncell; ++c)
        {
            int const cc[3] = {int(c % P.n[0]), int((c / P.n[0]) % P.n[1]), int(c / (int64_t(P.n[0]) * P.n[1]))};
            // RelativeGlobalDomainPosition filter on y (dimension 1)
            f32 const rel = f32(cc[1] + globalOff[1]) / f32(globalN[1]);
            f32 const sign = (rel >= 0.25f && rel < 0.75f) ? -1.0f : 1.0f;
            int64_t const gcell = int64_t(cc[0] + globalOff[0])
                + int64_t(globalN[0]) * (int64_t(cc[1] + globalOff[1]) + int64_t(globalN[1]) * int64_t(cc[2] + globalOff[2]));
            for(int k = 0; k < ppc; ++k)
            {
                int64_t const i = c * ppc + k;
                int const cur = ppc - 1 - k; // m_currentMacroParticles counts down
                int const ic[3] = {cur % ppcDim[0], (cur / ppcDim[0]) % ppcDim[1], cur / (ppcDim[0] * ppcDim[1])};
                for(int d = 0; d < 3; ++d)
                {
                    f32 const p = f32(ic[d]) * spacing[d] + spacing[d] * 0.5f;
                    posE[d * np + i] = p;
                    posI[d * np + i] = p;
                }
                wE[i] = weighting;
                wI[i] = weighting;
                cellE[i] = int32_t(c);
                cellI[i] = int32_t(c);
                uint64_t const gid = uint64_t(gcell) * uint64_t(ppc) + uint64_t(k);
                uint32_t ctr[4] = {uint32_t(gid), uint32_t(gid >> 32), 0u, 0u};
                uint32_t const key[2] = {seed, 0u};
                philox4x32_10(ctr, key);
                f32 const r0 = std::sqrt(-2.0f * std::log(u01(ctr[0])));
                f32 const r1 = std::sqrt(-2.0f * std::log(u01(ctr[2])));
                f32 const a0 = 6.283185307179586f * u01(ctr[1]);
                f32 const a1 = 6.283185307179586f * u01(ctr[3]);
                f32 const nrm[3] = {r0 * std::cos(a0), r0 * std::sin(a0), r1 * std::cos(a1)};
                momE[0 * np + i] = sign * driftE + nrm[0] * stddev;
                momE[1 * np + i] = 0.0f + nrm[1] * stddev;
                momE[2 * np + i] = 0.0f + nrm[2] * stddev;
                momI[0 * np + i] = sign * driftI;
                momI[1 * np + i] = 0.0f;
                momI[2 * np + i] = 0.0f;
            }
        }
    }

    int orc_num_threads()
    {
#ifdef _OPENMP
        return omp_get_max_threads();
#else
        return 1;
#endif
    }

    void orc_set_num_threads(int n)
    {
#ifdef _OPENMP
        omp_set_num_threads(n);
#else
        (void) n;
#endif
    }
}
