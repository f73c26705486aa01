// init.cu — synthetic KelvinHelmholtz initial condition generated on the device (bench input).
// Recipe: share/picongpu/examples/KelvinHelmholtz/include/picongpu/param/{particle,density,speciesInitialization}.param:
// homogeneous density, Quiet start ppc_x*ppc_y*ppc_z lattice (QuietImpl.hpp:47-118, filled from the highest lattice
// index down), ions cloned from electrons (Derive), drift +-x by global y quarter (Drift.hpp:56-80), electron
// temperature (Temperature.hpp:63-87).  RNG: Philox4x32-10, key (seed,0), counter (global particle id lo, hi, 0, 0);
// identical stream layout as the oracle's orc_khi_init (the transcendental functions differ in the last ulp).
// The particles are written directly in frame-run order, so no re-sort is needed afterwards.
#include "common.cuh"

namespace picstep
{
    __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
    {
#pragma unroll
        for(int r = 0; r < 10; ++r)
        {
            uint32_t const hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
            uint32_t const hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
            uint32_t const n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
            c[0] = n0;
            c[1] = n1;
            c[2] = n2;
            c[3] = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }

    __device__ __forceinline__ float u01(uint32_t r)
    {
        return (float(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
    }

    // one thread per cell, cells enumerated in key order (supercell major, localCellIdx minor)
    __global__ void __launch_bounds__(256) khiInitKernel(DevParams P, SpeciesDev E, SpeciesDev I, uint32_t* __restrict__ offE, uint32_t* __restrict__ offI, KhiArgs A)
    {
        int const ncell = P.nsc[0] * P.nsc[1] * P.nsc[2] * SCVOL;
        int const key = blockIdx.x * blockDim.x + threadIdx.x;
        int const ppc = A.ppc[0] * A.ppc[1] * A.ppc[2];
        if(key == 0)
        {
            offE[ncell] = uint32_t(ncell) * ppc;
            offI[ncell] = uint32_t(ncell) * ppc;
        }
        if(key >= ncell)
            return;
        offE[key] = uint32_t(key) * ppc;
        offI[key] = uint32_t(key) * ppc;
        int const sc = key / SCVOL, lc = key % SCVOL;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        int const cx = scx * SCX + lc % SCX, cy = scy * SCY + (lc / SCX) % SCY, cz = scz * SCZ + lc / (SCX * SCY);
        float const rel = float(cy + A.globalOff[1]) / float(A.globalN[1]);
        float const sign = (rel >= 0.25f && rel < 0.75f) ? -1.0f : 1.0f;
        unsigned long long const gcell = (unsigned long long) (cx + A.globalOff[0])
            + (unsigned long long) A.globalN[0] * ((unsigned long long) (cy + A.globalOff[1]) + (unsigned long long) A.globalN[1] * (unsigned long long) (cz + A.globalOff[2]));
        float const sp[3] = {1.0f / float(A.ppc[0]), 1.0f / float(A.ppc[1]), 1.0f / float(A.ppc[2])};
        for(int k = 0; k < ppc; ++k)
        {
            uint32_t const i = uint32_t(key) * ppc + k;
            int const cur = ppc - 1 - k;
            int const ic[3] = {cur % A.ppc[0], (cur / A.ppc[0]) % A.ppc[1], cur / (A.ppc[0] * A.ppc[1])};
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                // explicit roundings: identical bits with and without FMA contraction
                float const p = __fadd_rn(__fmul_rn(float(ic[d]), sp[d]), __fmul_rn(sp[d], 0.5f));
                E.pos[d][i] = p;
                I.pos[d][i] = p;
            }
            E.w[i] = A.weighting;
            I.w[i] = A.weighting;
            E.cell[i] = uint16_t(lc);
            I.cell[i] = uint16_t(lc);
            unsigned long long const gid = gcell * (unsigned long long) ppc + (unsigned long long) k;
            uint32_t ctr[4] = {uint32_t(gid), uint32_t(gid >> 32), 0u, 0u};
            philox4x32_10(ctr, A.seed, 0u);
            float const r0 = sqrtf(-2.0f * logf(u01(ctr[0])));
            float const r1 = sqrtf(-2.0f * logf(u01(ctr[2])));
            float const a0 = 6.283185307179586f * u01(ctr[1]);
            float const a1 = 6.283185307179586f * u01(ctr[3]);
            E.mom[0][i] = sign * A.driftE + (r0 * cosf(a0)) * A.stddev;
            E.mom[1][i] = 0.0f + (r0 * sinf(a0)) * A.stddev;
            E.mom[2][i] = 0.0f + (r1 * cosf(a1)) * A.stddev;
            I.mom[0][i] = sign * A.driftI;
            I.mom[1][i] = 0.0f;
            I.mom[2][i] = 0.0f;
        }
    }

    // Uniform warm plasma of ONE species (share/picongpu/benchmarks/Thermal: homogeneous density, `ppc` particles per cell
    // at random in-cell positions (startPosition::Random, RandomImpl.hpp), Maxwellian momenta with standard deviation
    // `stddev` per axis (Temperature.hpp:63-87)).  Philox counters: (global particle id, 0) momenta as in the KHI
    // generator, (global particle id, 1) positions.  Written directly in frame-run order.
    __global__ void __launch_bounds__(256) thermalInitKernel(DevParams P, SpeciesDev S, uint32_t* __restrict__ off, KhiArgs A)
    {
        int const ncell = P.nsc[0] * P.nsc[1] * P.nsc[2] * SCVOL;
        int const key = blockIdx.x * blockDim.x + threadIdx.x;
        int const ppc = A.ppc[0];
        if(key == 0)
            off[ncell] = uint32_t(ncell) * ppc;
        if(key >= ncell)
            return;
        off[key] = uint32_t(key) * ppc;
        int const sc = key / SCVOL, lc = key % SCVOL;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        int const cx = scx * SCX + lc % SCX, cy = scy * SCY + (lc / SCX) % SCY, cz = scz * SCZ + lc / (SCX * SCY);
        unsigned long long const gcell = (unsigned long long) (cx + A.globalOff[0])
            + (unsigned long long) A.globalN[0] * ((unsigned long long) (cy + A.globalOff[1]) + (unsigned long long) A.globalN[1] * (unsigned long long) (cz + A.globalOff[2]));
        for(int k = 0; k < ppc; ++k)
        {
            uint32_t const i = uint32_t(key) * ppc + k;
            unsigned long long const gid = gcell * (unsigned long long) ppc + (unsigned long long) k;
            uint32_t ctr[4] = {uint32_t(gid), uint32_t(gid >> 32), 1u, 0u};
            philox4x32_10(ctr, A.seed, 0u);
#pragma unroll
            for(int d = 0; d < 3; ++d)
                S.pos[d][i] = fminf(u01(ctr[d]), 1.0f - 1.0f / 16777216.0f);
            S.w[i] = A.weighting;
            S.cell[i] = uint16_t(lc);
            uint32_t c2[4] = {uint32_t(gid), uint32_t(gid >> 32), 0u, 0u};
            philox4x32_10(c2, A.seed, 0u);
            float const r0 = sqrtf(-2.0f * logf(u01(c2[0])));
            float const r1 = sqrtf(-2.0f * logf(u01(c2[2])));
            float const a0 = 6.283185307179586f * u01(c2[1]);
            float const a1 = 6.283185307179586f * u01(c2[3]);
            S.mom[0][i] = (r0 * cosf(a0)) * A.stddev;
            S.mom[1][i] = (r0 * sinf(a0)) * A.stddev;
            S.mom[2][i] = (r1 * cosf(a1)) * A.stddev;
        }
    }

    cudaError_t launchThermalInit(DevParams const& P, SpeciesDev S, uint32_t* off, KhiArgs const& A, cudaStream_t st)
    {
        int const ncell = P.nsc[0] * P.nsc[1] * P.nsc[2] * SCVOL;
        thermalInitKernel<<<(ncell + 255) / 256, 256, 0, st>>>(P, S, off, A);
        return cudaGetLastError();
    }

    cudaError_t launchKhiInit(DevParams const& P, SpeciesDev E, SpeciesDev I, uint32_t* offE, uint32_t* offI, KhiArgs const& A, cudaStream_t st)
    {
        int const ncell = P.nsc[0] * P.nsc[1] * P.nsc[2] * SCVOL;
        khiInitKernel<<<(ncell + 255) / 256, 256, 0, st>>>(P, E, I, offE, offI, A);
        return cudaGetLastError();
    }
} // namespace picstep
