// common.cuh — shared device-side definitions of libpicstep (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace picstep
{
    // SuperCellSize 8x8x4 and 256-slot frames (reference: include/picongpu/param/memory.param:51-60)
    constexpr int SCX = 8, SCY = 8, SCZ = 4, SCVOL = SCX * SCY * SCZ;
    constexpr int FRAME_SLOTS = 256;

    // Re-sort keys: key = supercell_linear * 256 + localCellIdx for particles that stay on this rank.
    // bit31 marks a particle leaving the rank along the split axis, bit30 the side (0 lower, 1 upper); the low
    // 30 bits are then the key in the RECEIVER's key space.  KEY_DROP: particle left through an open boundary.
    constexpr uint32_t KEY_LEAVE = 0x80000000u;
    constexpr uint32_t KEY_UPPER = 0x40000000u;
    constexpr uint32_t KEY_MASK = 0x3FFFFFFFu;
    constexpr uint32_t KEY_DROP = 0xFFFFFFFFu;

#ifdef PICSTEP_EXACT
    // bit-exact build (-fmad=false): the reference's CPU backend computes rsqrt as 1/sqrt; IEEE division and sqrt
    __device__ __forceinline__ float ps_rsqrt(float x)
    {
        return 1.0f / sqrtf(x);
    }
    __device__ __forceinline__ float ps_div(float a, float b)
    {
        return a / b;
    }
    __device__ __forceinline__ float ps_sqrt(float x)
    {
        return sqrtf(x);
    }
#else
    // production build: rsqrt as the reference's CUDA backend (alpaka rsqrt -> ::rsqrtf); division and sqrt through
    // the SFU approximations (2 ulp) instead of the IEEE sequences with their slow-path calls -- the kernels are
    // issue bound and the error is far inside the 1e-5 parity tolerance (tests/test_gpu_parity.py)
    __device__ __forceinline__ float ps_rsqrt(float x)
    {
        return rsqrtf(x);
    }
    __device__ __forceinline__ float ps_div(float a, float b)
    {
        return __fdividef(a, b);
    }
    __device__ __forceinline__ float ps_sqrt(float x)
    {
        float r;
        asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
#endif

    struct DevParams
    {
        int n[3]; // local cells
        int g[3]; // guard cells per side
        int N[3]; // padded cells
        int nsc[3]; // local supercells
        int wrap[3]; // periodic wrap handled inside this rank
        int open[3]; // 1: non-periodic outer boundary on this axis (particles leaving are dropped)
        int split_axis; // axis decomposed over ranks, -1 if none
        int has_lower, has_upper; // neighbour ranks exist along split_axis
        long long vol; // N0*N1*N2
        float cell[3];
        float dt, c, eps0, mue0;
        int lehe_dir;
        // guard exchange passes span [tlo, thi) of the padded extent of a transverse axis: the whole extent where the
        // axis is periodic / has a neighbour rank on that side, only the active cells where it ends at an open boundary
        // (the reference has no edge / corner exchange across a missing neighbour, Mask::getRelativeDirections)
        int tlo[3], thi[3];
        // [0] trajectories deposited through the global-atomic path of the run kernel (too wide for the node window),
        // [1] PQS trajectories that sent one plane of nodes through global atomics; read and reset by
        // picstep_reduce(PICSTEP_REDUCE_SLOW_PATH)
        unsigned long long* stats;
    };

    // exponential field absorber: thickness per [axis][side] (0 where the face is not absorbing) and the tabulated
    // attenuation exp(-strength * factor), factor < ABS_MAX, in device memory
    constexpr int ABS_MAX = 256;
    struct AbsorberDev
    {
        int cells[3][2];
        float const* damp;
    };

    // Supercells covered by one launch of the run kernel: the layers first, first + stride, ... (count of them) along
    // `axis`, all supercells in the two other axes.  The whole domain is {2, 0, 1, nsc[2]}; with a domain decomposition
    // the BORDER area (the two layers that face the neighbour ranks, pmacc/type/Area.hpp:36-41,
    // AreaMappingMethods.hpp:41-120) is launched first and its exchange travels while the CORE area is computed.
    struct ScArea
    {
        int axis, first, stride, count;
    };

    // PlaneWave incident field on the YMin Huygens surface (fields.cu: incidentKernel), unitless profile parameters
    constexpr int LASER_MAX_MODES = 8;

    struct LaserDev
    {
        int polarisation, plane; // plane: padded-grid y index of the updated plane of this call
        int profile; // picstep_laser_profile
        float planeTotal; // its total (global) cell index along y
        float amplitude, omega, pulseDuration, nofocusConstant, rampInit, phase, timeDelay;
        float pol[3], axis2[3]; // internal axes 1 and 2 (axis 0 = propagation = +y)
        float origin[3], focus[3]; // BaseFunctorE::getOrigin / getFocus
        float phaseVelocity; // Yee numerical phase velocity along y, in units of c
        float currentTimeOrigin; // currentStep * dt (fractional step)
        float baseCoefficient; // curl coefficient / cellSize.y (direction +1)
        int updatedIsE;
        int lo[2], hi[2]; // updated cells along x and z (local = global: neither axis is split), [lo, hi)
        int lastDomain[2]; // the last cell of the range drops one component (Solver.kernel:318-325,458-466)
        // GaussianPulse
        float w0, waveLength, rayleighLength, timeShift, tanTilt[2];
        int tilted, nModes;
        float modes[LASER_MAX_MODES], modePhases[LASER_MAX_MODES];
        // Wavepacket / Polynom / ExpRampWithPrepulse (picstep.h: laser_w0_axis, laser_profile_params)
        float w0Axis[2], prm[16];
    };

    // convolutional PML (fields.cu: pmlUpdate*Kernel): local thickness per [axis][negative, positive], graded parameters,
    // psi = six planes yx, zx, xy, zy, xz, yz over the padded grid
    struct PmlDev
    {
        int thickness[3][2];
        float sigmaMax[3], kappaMax[3], alphaMax[3];
        float sigmaKappaGradingOrder, alphaGradingOrder;
        float* psi;
    };

    struct SpeciesDev
    {
        float* pos[3];
        float* mom[3];
        float* w;
        uint16_t* cell; // localCellIdx
        float mass_per_w; // getMass<Frame>()  = base_mass * massRatio
        float charge_per_w; // getCharge<Frame>() = base_charge * chargeRatio
    };

    struct Field3
    {
        float* c[3];
    };

    __device__ __forceinline__ long long fidx(DevParams const& P, int x, int y, int z)
    {
        return ((long long) z * P.N[1] + y) * P.N[0] + x;
    }

    // Lehe solver coefficients per differentiation direction (Lehe/Derivative.hpp:94-137)
    struct LeheCoeffs
    {
        float alpha[3], delta[3], beta1[3], beta2[3];
    };

    // arguments of the synthetic KelvinHelmholtz initial condition (init.cu)
    struct KhiArgs
    {
        int ppc[3];
        int globalN[3];
        int globalOff[3];
        float weighting, driftE, driftI, stddev;
        uint32_t seed;
    };

    // 32-byte migration record (KernelCopyGuardToExchange's "border frame", one particle per slot)
    struct __align__(16) MigRecord
    {
        float px, py, pz, ux;
        float uy, uz, w;
        uint32_t key;
    };
} // namespace picstep
