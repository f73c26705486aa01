// deposit.cu — charge conserving current deposition (reference kernel K7: KernelComputeCurrent + ComputePerFrame,
// include/picongpu/fields/FieldJ.kernel:52-142; Esirkepov.hpp:62-242; EmZ.hpp:66-155, EmZ/DepositCurrent.hpp:35-119).
//
// Two implementations share the per-particle trajectory set-up:
//
//  * depositCellKernel (default) — exploits the cell-sorted frame runs.  Shared-memory fp32 atomicAdd is a CAS
//    loop on sm_100a (SASS: LDS + FADD + ATOMS.CAST.SPIN), so instead of 54..144 shared atomics per particle the
//    roles are transposed: a warp owns ONE cell at a time; in phase 1 each lane prepares the 1-D assignment
//    arrays of one particle (S0, DS per axis on the 5..7 point window around the cell); in phase 2 each lane owns
//    one transverse node (a,b) of that window and accumulates, over all particles of the cell, the current along
//    the third axis in REGISTERS.  The per-cell result is added with plain LDS/FADD/STS to a warp-private tile
//    (a warp owns one y-row of the supercell, so no other warp touches it); the eight private tiles are summed
//    and flushed once per supercell with red.global.add.f32.  No shared-memory atomics at all.
//
//  * depositAtomicKernel (flags bit0) — the reference's "CachedSupercells" strategy restated: thread per
//    particle, loop bounds and summation order exactly as Esirkepov.hpp:204-241, shared atomics, atomic flush.
//    Kept as cross-check and as the baseline the ncu profiles compare against.
#include "common.cuh"
#include "esirkepov.cuh"
#include "shapes.cuh"

namespace picstep
{
    __device__ __forceinline__ float norm2d(float x, float y, float z)
    {
        float t = x * x;
        t += y * y;
        t += z * z;
        return t;
    }

    template<int SHAPE>
    struct JTile
    {
        static constexpr int LO = CurrentMargin<SHAPE>::LO, UP = CurrentMargin<SHAPE>::UP;
        static constexpr int TX = SCX + LO + UP, TY = SCY + LO + UP, TZ = SCZ + LO + UP;
        static constexpr int TV = TX * TY * TZ;
    };

    // ------------------------------------------------------------------------------------------------------------
    // Reference-strategy kernel (thread per particle, shared atomics)
    // ------------------------------------------------------------------------------------------------------------
    /** One rotated 1-D pass of Esirkepov (Esirkepov.hpp:147-242).  R0,R1,R2: original axes of the rotated i,j,k. */
    template<int SHAPE, int R0, int R1, int R2>
    __device__ __forceinline__ void esirkepov1D(
        float* __restrict__ tile, // component R2 of the J tile
        int const base[3], // tile coordinates of the (grid-shifted) particle cell
        int const status[3],
        float const p0[3],
        float const p1[3],
        float currentSurfaceDensity)
    {
        using S = Shape<SHAPE>;
        using T = JTile<SHAPE>;
        if(p0[R2] == p1[R2])
            return;
        constexpr int begin = S::BEGIN, end = S::BEGIN + S::SUPP;
        float s0i[S::SUPP + 1], s1i[S::SUPP + 1], s0j[S::SUPP + 1], s1j[S::SUPP + 1], s0k[S::SUPP + 1], s1k[S::SUPP + 1];
        shapeOff<SHAPE>(p0[R0], !(status[R0] & 2), s0i);
        shapeOff<SHAPE>(p1[R0], !(status[R0] & 4), s1i);
        shapeOff<SHAPE>(p0[R1], !(status[R1] & 2), s0j);
        shapeOff<SHAPE>(p1[R1], !(status[R1] & 4), s1j);
        shapeOff<SHAPE>(p0[R2], !(status[R2] & 2), s0k);
        shapeOff<SHAPE>(p1[R2], !(status[R2] & 4), s1k);
        int const leaveI = status[R0] & 1, leaveJ = status[R1] & 1, leaveK = status[R2] & 1;
        int const stride[3] = {1, T::TX, T::TX * T::TY};
        int const origin = base[0] + T::TX * (base[1] + T::TY * base[2]);
#pragma unroll
        for(int i = begin; i < end + 1; ++i)
            if(i < end + leaveI)
            {
                float const a0 = s0i[i - begin];
                float const da = s1i[i - begin] - a0;
#pragma unroll
                for(int j = begin; j < end + 1; ++j)
                    if(j < end + leaveJ)
                    {
                        float const b0 = s0j[j - begin];
                        float const db = s1j[j - begin] - b0;
                        float const tmp = -currentSurfaceDensity * (a0 * b0 + 0.5f * (da * b0 + a0 * db) + (1.0f / 3.0f) * db * da);
                        float acc = 0.0f;
#pragma unroll
                        for(int k = begin; k < end; ++k)
                            if(k < end + leaveK - 1)
                            {
                                float const W = (s1k[k - begin] - s0k[k - begin]) * tmp;
                                acc += W;
                                atomicAdd(&tile[origin + i * stride[R0] + j * stride[R1] + k * stride[R2]], acc);
                            }
                    }
            }
    }

    /** emz::DepositCurrent::cptCurrent1D (EmZ/DepositCurrent.hpp:77-118): on-support segment, fixed bounds */
    template<int SHAPE, int R0, int R1, int R2>
    __device__ __forceinline__ void emz1D(float* __restrict__ tile, int const base[3], float const p0[3], float const p1[3], float currentSurfaceDensity)
    {
        using S = Shape<SHAPE>;
        using T = JTile<SHAPE>;
        if(p0[R2] == p1[R2])
            return;
        constexpr int begin = S::BEGIN;
        float s0i[S::SUPP], s1i[S::SUPP], s0j[S::SUPP], s1j[S::SUPP], s0k[S::SUPP], s1k[S::SUPP];
        S::on(p0[R0], s0i);
        S::on(p1[R0], s1i);
        S::on(p0[R1], s0j);
        S::on(p1[R1], s1j);
        S::on(p0[R2], s0k);
        S::on(p1[R2], s1k);
        int const stride[3] = {1, T::TX, T::TX * T::TY};
        int const origin = base[0] + T::TX * (base[1] + T::TY * base[2]);
#pragma unroll
        for(int i = 0; i < S::SUPP; ++i)
        {
            float const a0 = s0i[i];
            float const da = s1i[i] - a0;
#pragma unroll
            for(int j = 0; j < S::SUPP; ++j)
            {
                float const b0 = s0j[j];
                float const db = s1j[j] - b0;
                float const tmp = -currentSurfaceDensity * (a0 * b0 + 0.5f * (da * b0 + a0 * db) + (1.0f / 3.0f) * db * da);
                float acc = 0.0f;
#pragma unroll
                for(int k = 0; k < S::SUPP - 1; ++k)
                {
                    float const W = (s1k[k] - s0k[k]) * tmp;
                    acc += W;
                    atomicAdd(&tile[origin + (i + begin) * stride[R0] + (j + begin) * stride[R1] + (k + begin) * stride[R2]], acc);
                }
            }
        }
    }

    template<int SHAPE, int SOLVER>
    __device__ __forceinline__ void depositParticleAtomic(DevParams const& P, float* __restrict__ tile, int lx, int ly, int lz, float const pos[3], float const vel[3], float charge)
    {
        using S = Shape<SHAPE>;
        using T = JTile<SHAPE>;
        constexpr bool even = (S::SUPP % 2) == 0;
        int const l[3] = {lx + T::LO, ly + T::LO, lz + T::LO};
        float* const tx = tile;
        float* const ty = tile + T::TV;
        float* const tz = tile + 2 * T::TV;
        if constexpr(SOLVER == 0)
        {
            float p0[3], p1[3];
            int status[3], base[3];
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                float const dp = vel[d] * P.dt / P.cell[d];
                p0[d] = pos[d] - dp;
                p1[d] = pos[d];
                int iS, iE;
                relay<even>(iS, iE, p0[d], p1[d]);
                int const gs = iS < iE ? iS : iE;
                status[d] = (gs == iS ? 2 : 0) | (gs == iE ? 4 : 0) | (iS != iE ? 1 : 0);
                p0[d] -= float(gs);
                p1[d] -= float(gs);
                base[d] = l[d] + gs;
            }
            float const vol = P.cell[0] * P.cell[1] * P.cell[2];
            float const csd = charge * (1.0f / float(vol * P.dt));
            esirkepov1D<SHAPE, 1, 2, 0>(tx, base, status, p0, p1, csd * P.cell[0]);
            esirkepov1D<SHAPE, 2, 0, 1>(ty, base, status, p0, p1, csd * P.cell[1]);
            esirkepov1D<SHAPE, 0, 1, 2>(tz, base, status, p0, p1, csd * P.cell[2]);
        }
        else
        {
            float pS[3], rl[3];
            int sS[3], sE[3];
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                float const dp = (vel[d] * P.dt) / P.cell[d];
                pS[d] = pos[d] - dp;
                rl[d] = relay<even>(sS[d], sE[d], pS[d], pos[d]);
            }
            float const cd = charge / (P.cell[0] * P.cell[1] * P.cell[2]);
            float q0[3], q1[3];
            int base[3];
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                q0[d] = pS[d] - float(sS[d]);
                q1[d] = rl[d] - float(sS[d]);
                base[d] = l[d] + sS[d];
            }
            emz1D<SHAPE, 1, 2, 0>(tx, base, q0, q1, P.cell[0] * cd / P.dt);
            emz1D<SHAPE, 2, 0, 1>(ty, base, q0, q1, P.cell[1] * cd / P.dt);
            emz1D<SHAPE, 0, 1, 2>(tz, base, q0, q1, P.cell[2] * cd / P.dt);
            if(sS[0] != sE[0] || sS[1] != sE[1] || sS[2] != sE[2])
            {
#pragma unroll
                for(int d = 0; d < 3; ++d)
                {
                    q1[d] = pos[d] - float(sE[d]);
                    q0[d] = rl[d] - float(sE[d]);
                    base[d] = l[d] + sE[d];
                }
                emz1D<SHAPE, 1, 2, 0>(tx, base, q0, q1, P.cell[0] * cd / P.dt);
                emz1D<SHAPE, 2, 0, 1>(ty, base, q0, q1, P.cell[1] * cd / P.dt);
                emz1D<SHAPE, 0, 1, 2>(tz, base, q0, q1, P.cell[2] * cd / P.dt);
            }
        }
    }

    template<int SHAPE>
    __device__ __forceinline__ void flushTile(DevParams const& P, Field3 J, float const* __restrict__ tile, int scx, int scy, int scz)
    {
        using T = JTile<SHAPE>;
        int const ox = scx * SCX + P.g[0] - T::LO, oy = scy * SCY + P.g[1] - T::LO, oz = scz * SCZ + P.g[2] - T::LO;
        for(int i = threadIdx.x; i < 3 * T::TV; i += blockDim.x)
        {
            float const v = tile[i];
            if(v != 0.0f)
            {
                int const comp = i / T::TV;
                int const r = i % T::TV;
                int const x = r % T::TX, y = (r / T::TX) % T::TY, z = r / (T::TX * T::TY);
                atomicAdd(J.c[comp] + fidx(P, ox + x, oy + y, oz + z), v); // RED.E.ADD.F32
            }
        }
    }

    template<int SHAPE, int SOLVER>
    __global__ void __launch_bounds__(256) depositAtomicKernel(DevParams P, SpeciesDev S, Field3 J, uint32_t const* __restrict__ cellOff)
    {
        using T = JTile<SHAPE>;
        extern __shared__ float tile[];
        int const sc = blockIdx.x;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        uint32_t const p0 = cellOff[sc * SCVOL], p1 = cellOff[(sc + 1) * SCVOL];
        if(p0 == p1)
            return;
        for(int i = threadIdx.x; i < 3 * T::TV; i += blockDim.x)
            tile[i] = 0.0f;
        __syncthreads();
        float const rc2 = float(1.0 / double(P.c) / double(P.c));
        for(uint32_t i = p0 + threadIdx.x; i < p1; i += blockDim.x)
        {
            float const pos[3] = {S.pos[0][i], S.pos[1][i], S.pos[2][i]};
            float const ux = S.mom[0][i], uy = S.mom[1][i], uz = S.mom[2][i];
            float const w = S.w[i];
            int const lc = S.cell[i];
            float const mass = S.mass_per_w * w;
            float const charge = S.charge_per_w * w;
            float const t = ps_rsqrt(mass * mass + norm2d(ux, uy, uz) * rc2);
            float const vel[3] = {t * ux, t * uy, t * uz};
            depositParticleAtomic<SHAPE, SOLVER>(P, tile, lc % SCX, (lc / SCX) % SCY, lc / (SCX * SCY), pos, vel, charge);
        }
        __syncthreads();
        flushTile<SHAPE>(P, J, tile, scx, scy, scz);
    }

    // ------------------------------------------------------------------------------------------------------------
    // Cell-sorted kernel: warp per cell, lane groups per particle, register accumulation
    // ------------------------------------------------------------------------------------------------------------
    // Window of grid offsets, relative to the particle's (new) cell, that a trajectory ending in the cell can touch:
    //   wide   [-LO, UP]       with the current solver margins (Esirkepov.hpp:42-45), WN_W = LO + UP + 1 points;
    //   narrow [-LO+1, UP-1]   for odd supports when start and end point both have their assignment cell at offset
    //                          0 or +1, i.e. the particle moved less than half a cell -- the common case.
    template<int SHAPE>
    struct Win
    {
        using S = Shape<SHAPE>;
        static constexpr int WLO = CurrentMargin<SHAPE>::LO;
        static constexpr int WN = CurrentMargin<SHAPE>::LO + CurrentMargin<SHAPE>::UP + 1;
        static constexpr bool HAS_NARROW = (S::SUPP % 2 == 1) && (WN - 2 >= 4);
        static constexpr int WN_N = HAS_NARROW ? WN - 2 : WN;
        static constexpr int OFF_N = HAS_NARROW ? 1 : 0;
        // per particle(-segment) record in shared memory, laid out on the WIDE window:
        //   float2 {S0, DS} [3 axes][WN]   then   C[3][WN-1] = prefix sums of DS * (-currentSurfaceDensity)
        static constexpr int REC = 6 * WN + 3 * (WN - 1);
        // stride == 4 (mod 8) words: 16-byte aligned records, conflict free 128-bit lane-strided stores
        static constexpr int RECP = REC + ((4 - REC % 8) + 8) % 8;
        static constexpr int NREC = 33; // 32 records + one all-zero record for the tail of a pass
        // warp-private J tile: a warp owns one y-row of cells of the supercell
        static constexpr int PX = SCX + WN - 1, PY = WN, PZ = SCZ + WN - 1, PV = PX * PY * PZ;
    };

    /** Phase 1 helper: place the FR-entry frame arrays (S0, S1) of one axis at window index n0 of a zeroed record
     * and write the scaled prefix sums of DS for the first NC frame entries (accumulated_J recursion of
     * Esirkepov.hpp:223-236, factored: J_k = C_k * transverse weight). */
    template<int FR>
    __device__ __forceinline__ void placeFrame(float2* __restrict__ sd, float* __restrict__ cpre, int n0, float const* s0, float const* s1, int nc, float factor)
    {
        float run = 0.0f;
#pragma unroll
        for(int s = 0; s < FR; ++s)
        {
            float const ds = s1[s] - s0[s];
            sd[n0 + s] = make_float2(s0[s], ds);
            run += ds;
            if(s < nc)
                cpre[n0 + s] = run * factor;
        }
    }

    /** Phase 2: accumulate the records [0, nrec) of this warp into register accumulators and add them to the
     * warp-private tile.  WN_: window width handled (wide or narrow), OFF: index offset of that window inside the
     * wide-layout record.  Lane groups of G lanes work on one record each (PP records per pass); inside a group a
     * lane owns the transverse nodes (a = ah*AW + j, b), j < AW, and WN_-1 prefix values along the current axis. */
    template<int SHAPE, int WN_, int OFF>
    __device__ __forceinline__ void accumulateAndFlush(float const* __restrict__ myRecs, int nrec, float* __restrict__ myTile, int cxl, int cz, int lane)
    {
        using W = Win<SHAPE>;
        constexpr int AW = (WN_ == 6) ? 3 : 2;
        constexpr int NA = WN_ / AW; // a-groups
        constexpr int GUSED = NA * WN_;
        constexpr int G = GUSED <= 8 ? 8 : (GUSED <= 16 ? 16 : 32);
        constexpr int PP = 32 / G;
        constexpr int NK = WN_ - 1;
        static_assert(WN_ % AW == 0 && GUSED <= 32, "unsupported window width");
        int const g = lane % G, slot = lane / G;
        bool const valid = g < GUSED;
        int const ah = valid ? g / WN_ : 0, b = valid ? g % WN_ : 0;
        int const a0 = ah * AW;

        float accX[AW][NK], accY[AW][NK], accZ[AW][NK];
#pragma unroll
        for(int j = 0; j < AW; ++j)
#pragma unroll
            for(int k = 0; k < NK; ++k)
                accX[j][k] = accY[j][k] = accZ[j][k] = 0.0f;

        for(int base = 0; base < nrec; base += PP)
        {
            int const r = base + slot;
            float const* rec = myRecs + (r < nrec ? r : 32) * W::RECP; // record 32 is all zero
            float2 const* rx = reinterpret_cast<float2 const*>(rec) + OFF;
            float2 const* ry = rx + W::WN;
            float2 const* rz = rx + 2 * W::WN;
            float const* cxp = rec + 6 * W::WN + OFF;
            float const* cyp = cxp + (W::WN - 1);
            float const* czp = cyp + (W::WN - 1);
            float2 const zb = rz[b], yb = ry[b];
            // transverse weight S0i*S0j + 1/2 (DSi*S0j + S0i*DSj) + 1/3 DSi*DSj = S0i*(S0j + DSj/2) + DSi*(S0j/2 + DSj/3)
            float const zP = zb.x + 0.5f * zb.y, zQ = 0.5f * zb.x + (1.0f / 3.0f) * zb.y;
            float const yP = yb.x + 0.5f * yb.y, yQ = 0.5f * yb.x + (1.0f / 3.0f) * yb.y;
            float cx[NK], cy[NK], cz_[NK];
#pragma unroll
            for(int k = 0; k < NK; ++k)
            {
                cx[k] = cxp[k];
                cy[k] = cyp[k];
                cz_[k] = czp[k];
            }
#pragma unroll
            for(int j = 0; j < AW; ++j)
            {
                float2 const xa = rx[a0 + j], ya = ry[a0 + j];
                float const xP = xa.x + 0.5f * xa.y, xQ = 0.5f * xa.x + (1.0f / 3.0f) * xa.y;
                // Jx: (i,j) = (y,z) at node (a,b);  Jy: (z,x) at (x=a, z=b);  Jz: (x,y) at (a,b)
                float const tX = ya.x * zP + ya.y * zQ;
                float const tY = zb.x * xP + zb.y * xQ;
                float const tZ = xa.x * yP + xa.y * yQ;
#pragma unroll
                for(int k = 0; k < NK; ++k)
                {
                    accX[j][k] += cx[k] * tX;
                    accY[j][k] += cy[k] * tY;
                    accZ[j][k] += cz_[k] * tZ;
                }
            }
        }
        // combine the PP record slots (butterfly over the slot bits of the lane id)
#pragma unroll
        for(int o = G; o < 32; o <<= 1)
#pragma unroll
            for(int j = 0; j < AW; ++j)
#pragma unroll
                for(int k = 0; k < NK; ++k)
                {
                    accX[j][k] += __shfl_xor_sync(0xffffffffu, accX[j][k], o);
                    accY[j][k] += __shfl_xor_sync(0xffffffffu, accY[j][k], o);
                    accZ[j][k] += __shfl_xor_sync(0xffffffffu, accZ[j][k], o);
                }
        // add to the warp-private tile: plain read-modify-write, no atomics (within one instruction all active lanes
        // address distinct nodes, no other warp touches this tile); entry e is written by record slot e % PP.
        // window index n <-> private tile coordinate (cxl + n + OFF, n + OFF, cz + n + OFF) per axis
        if(valid)
        {
            int const tb = b + OFF;
#pragma unroll
            for(int j = 0; j < AW; ++j)
            {
                int const ta = a0 + j + OFF;
#pragma unroll
                for(int k = 0; k < NK; ++k)
                {
                    int const e = j * NK + k;
                    if(e % PP == slot)
                    {
                        int const tk = k + OFF;
                        myTile[(cxl + tk) + W::PX * (ta + W::PY * (cz + tb))] += accX[j][k]; // Jx: x=k, y=a, z=b
                        myTile[W::PV + (cxl + ta) + W::PX * (tk + W::PY * (cz + tb))] += accY[j][k]; // Jy: x=a, y=k, z=b
                        myTile[2 * W::PV + (cxl + ta) + W::PX * (tb + W::PY * (cz + tk))] += accZ[j][k]; // Jz: x=a, y=b, z=k
                    }
                }
            }
        }
    }

    template<int SHAPE, int SOLVER, int WARPS>
    __global__ void __launch_bounds__(WARPS * 32) depositCellKernel(DevParams P, SpeciesDev S, Field3 J, uint32_t const* __restrict__ cellOff)
    {
        using Sh = Shape<SHAPE>;
        using T = JTile<SHAPE>;
        using W = Win<SHAPE>;
        static_assert(SCY % WARPS == 0, "a CTA owns WARPS consecutive y-rows of one supercell");
        constexpr int PARTS = SCY / WARPS; // CTAs per supercell
        constexpr bool even = (Sh::SUPP % 2) == 0;
        constexpr int WN = W::WN;
        constexpr int NSEG = SOLVER == 0 ? 1 : 2; // EmZ: up to two on-support segments per particle
        constexpr int CHUNK = 32 / NSEG; // particles per phase-1 pass: 32 records per warp
        constexpr int FR = SOLVER == 0 ? Sh::SUPP + 1 : Sh::SUPP; // frame entries per axis

        extern __shared__ __align__(16) float smem[];
        float* tiles = smem; // WARPS * 3 * PV warp-private tiles
        float* recs = smem + WARPS * 3 * W::PV; // WARPS * NREC * RECP (16-byte aligned: PV*3*WARPS*4 bytes offset)

        int const sc = blockIdx.x / PARTS, part = blockIdx.x % PARTS;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        int const ly = part * WARPS + warp;
        for(int i = threadIdx.x; i < WARPS * 3 * W::PV; i += blockDim.x)
            tiles[i] = 0.0f;
        float* const myTile = tiles + warp * 3 * W::PV;
        float* const myRecs = recs + warp * W::NREC * W::RECP;
        for(int i = lane; i < W::RECP; i += 32)
            myRecs[32 * W::RECP + i] = 0.0f;
        __syncthreads();

        float const rc2 = float(1.0 / double(P.c) / double(P.c));
        float const vol = P.cell[0] * P.cell[1] * P.cell[2];

        for(int cz = 0; cz < SCZ; ++cz)
            for(int cxl = 0; cxl < SCX; ++cxl)
            {
                int const lc = cxl + SCX * (ly + SCY * cz);
                uint32_t const c0 = cellOff[sc * SCVOL + lc], c1 = cellOff[sc * SCVOL + lc + 1];
                for(uint32_t chunk = c0; chunk < c1; chunk += CHUNK)
                {
                    uint32_t const i = chunk + lane;
                    int const nIn = int(min(uint32_t(CHUNK), c1 - chunk));
                    bool narrowOk = true;
                    __syncwarp();
                    // ---- phase 1: lane = particle --------------------------------------------------------------
                    if(lane < CHUNK && i < c1)
                    {
                        float const pos[3] = {S.pos[0][i], S.pos[1][i], S.pos[2][i]};
                        float const ux = S.mom[0][i], uy = S.mom[1][i], uz = S.mom[2][i];
                        float const w = S.w[i];
                        float const mass = S.mass_per_w * w;
                        float const charge = S.charge_per_w * w;
                        float const t = ps_rsqrt(mass * mass + norm2d(ux, uy, uz) * rc2);
                        float const vel[3] = {t * ux, t * uy, t * uz};
                        float* rec = myRecs + lane * NSEG * W::RECP;
                        // zero the record(s), then place the frames at their window offset
                        {
                            float4* z4 = reinterpret_cast<float4*>(rec);
#pragma unroll
                            for(int q = 0; q < NSEG * W::RECP / 4; ++q)
                                z4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        if constexpr(SOLVER == 0)
                        {
                            float const csd = charge * (1.0f / float(vol * P.dt));
#pragma unroll
                            for(int d = 0; d < 3; ++d)
                            {
                                float const dp = vel[d] * P.dt / P.cell[d];
                                float const x0 = pos[d] - dp, x1 = pos[d];
                                int iS, iE;
                                relay<even>(iS, iE, x0, x1);
                                // Esirkepov.hpp:84-103: both points in the frame of gridShift = min(iS,iE), off-support
                                // arrays of SUPP+1 entries (same arithmetic, same bits as the reference)
                                int const gs = iS < iE ? iS : iE;
                                float const y0 = x0 - float(gs), y1 = x1 - float(gs);
                                float s0[FR], s1[FR];
                                shapeOff<SHAPE>(y0, gs != iS, s0);
                                shapeOff<SHAPE>(y1, gs != iE, s1);
                                float const f = (y0 == y1) ? 0.0f : -(csd * P.cell[d]);
                                int const leave = iS != iE ? 1 : 0;
                                placeFrame<FR>(reinterpret_cast<float2*>(rec) + WN * d, rec + 6 * WN + (WN - 1) * d, gs + Sh::BEGIN + W::WLO, s0, s1, Sh::SUPP - 1 + leave, f);
                                if(W::HAS_NARROW && (iS < 0 || iS > 1))
                                    narrowOk = false;
                            }
                        }
                        else
                        {
                            float pS[3], rl[3];
                            int sS[3], sE[3];
#pragma unroll
                            for(int d = 0; d < 3; ++d)
                            {
                                float const dp = (vel[d] * P.dt) / P.cell[d];
                                pS[d] = pos[d] - dp;
                                rl[d] = relay<even>(sS[d], sE[d], pS[d], pos[d]);
                                if(W::HAS_NARROW && (sS[d] < 0 || sS[d] > 1))
                                    narrowOk = false;
                            }
                            float const cd = charge / vol;
                            bool const two = sS[0] != sE[0] || sS[1] != sE[1] || sS[2] != sE[2];
                            float* rec2 = rec + W::RECP;
#pragma unroll
                            for(int d = 0; d < 3; ++d)
                            {
                                float const fd = -(P.cell[d] * cd / P.dt);
                                float const a0 = pS[d] - float(sS[d]), a1 = rl[d] - float(sS[d]);
                                float s0[FR], s1[FR];
                                Sh::on(a0, s0);
                                Sh::on(a1, s1);
                                placeFrame<FR>(reinterpret_cast<float2*>(rec) + WN * d, rec + 6 * WN + (WN - 1) * d, sS[d] + Sh::BEGIN + W::WLO, s0, s1, Sh::SUPP - 1, (a0 == a1) ? 0.0f : fd);
                                float const b0 = rl[d] - float(sE[d]), b1 = pos[d] - float(sE[d]);
                                Sh::on(b0, s0);
                                Sh::on(b1, s1);
                                placeFrame<FR>(reinterpret_cast<float2*>(rec2) + WN * d, rec2 + 6 * WN + (WN - 1) * d, sE[d] + Sh::BEGIN + W::WLO, s0, s1, Sh::SUPP - 1, (!two || b0 == b1) ? 0.0f : fd);
                            }
                        }
                    }
                    bool const allNarrow = W::HAS_NARROW && __all_sync(0xffffffffu, narrowOk);
                    __syncwarp();
                    // ---- phase 2: lane groups = records, lanes = transverse nodes, registers = current axis ----------
                    if(allNarrow)
                        accumulateAndFlush<SHAPE, W::WN_N, W::OFF_N>(myRecs, nIn * NSEG, myTile, cxl, cz, lane);
                    else
                        accumulateAndFlush<SHAPE, WN, 0>(myRecs, nIn * NSEG, myTile, cxl, cz, lane);
                }
            }
        __syncthreads();
        // ---- combine the warp-private tiles and flush once to global J (red.global.add.f32) --------------------------
        {
            int const ox = scx * SCX + P.g[0] - T::LO, oy = scy * SCY + P.g[1] - T::LO + part * WARPS, oz = scz * SCZ + P.g[2] - T::LO;
            constexpr int CY = WARPS + WN - 1; // y extent covered by this CTA
            constexpr int CV = T::TX * CY * T::TZ;
            for(int i = threadIdx.x; i < 3 * CV; i += blockDim.x)
            {
                int const comp = i / CV;
                int const r = i % CV;
                int const x = r % T::TX, y = (r / T::TX) % CY, z = r / (T::TX * CY);
                // CTA tile y = row + n  with row = warp index, n = window index in [0, WN)
                float v = 0.0f;
#pragma unroll
                for(int n = 0; n < WN; ++n)
                {
                    int const row = y - n;
                    if(row >= 0 && row < WARPS)
                        v += tiles[row * 3 * W::PV + comp * W::PV + x + W::PX * (n + W::PY * z)];
                }
                if(v != 0.0f)
                    atomicAdd(J.c[comp] + fidx(P, ox + x, oy + y, oz + z), v);
            }
        }
    }

    template<int SHAPE, int SOLVER>
    cudaError_t launchDepositT(bool atomicVariant, DevParams const& P, SpeciesDev const& S, Field3 J, uint32_t const* cellOff, cudaStream_t st)
    {
        int const nscTot = P.nsc[0] * P.nsc[1] * P.nsc[2];
        if(atomicVariant)
        {
            size_t const smem = sizeof(float) * 3 * JTile<SHAPE>::TV;
            cudaError_t e = cudaFuncSetAttribute(depositAtomicKernel<SHAPE, SOLVER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if(e != cudaSuccess)
                return e;
            depositAtomicKernel<SHAPE, SOLVER><<<nscTot, 256, smem, st>>>(P, S, J, cellOff);
        }
        else
        {
            constexpr int WARPS = 4;
            static_assert((WARPS * 3 * Win<SHAPE>::PV) % 4 == 0, "records must stay 16-byte aligned");
            size_t const smem = sizeof(float) * (WARPS * 3 * Win<SHAPE>::PV + WARPS * Win<SHAPE>::NREC * Win<SHAPE>::RECP);
            cudaError_t e = cudaFuncSetAttribute(depositCellKernel<SHAPE, SOLVER, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if(e != cudaSuccess)
                return e;
            depositCellKernel<SHAPE, SOLVER, WARPS><<<nscTot * (SCY / WARPS), WARPS * 32, smem, st>>>(P, S, J, cellOff);
        }
        return cudaGetLastError();
    }

    cudaError_t launchDeposit(int shape, int solver, bool atomicVariant, DevParams const& P, SpeciesDev const& S, Field3 J, uint32_t const* cellOff, cudaStream_t st)
    {
#define PS_CASE(SH, SO)                                                                                               \
    if(shape == SH && solver == SO)                                                                                   \
        return launchDepositT<SH, SO>(atomicVariant, P, S, J, cellOff, st);
        PS_CASE(0, 0)
        PS_CASE(1, 0)
        PS_CASE(2, 0)
        PS_CASE(3, 0)
        PS_CASE(4, 0)
        PS_CASE(1, 1)
        PS_CASE(2, 1)
        PS_CASE(3, 1)
        PS_CASE(4, 1)
#undef PS_CASE
        return cudaErrorInvalidValue;
    }
} // namespace picstep
