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

    // relayPoint.hpp:48-63 (only the two assignment-cell indices are needed by Esirkepov)
    template<bool EVEN>
    __device__ __forceinline__ float relay(int& i1, int& i2, float x1, float x2)
    {
        if constexpr(EVEN)
        {
            i1 = __float2int_rd(x1);
            i2 = __float2int_rd(x2);
            return i1 == i2 ? x2 : float(max(i1, i2));
        }
        else
        {
            i1 = __float2int_rd(x1 + 0.5f);
            i2 = __float2int_rd(x2 + 0.5f);
            return i1 == i2 ? x2 : float(i1 + i2) / 2.0f;
        }
    }

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
    // Cell-sorted kernel: warp per cell, lane per transverse node, register accumulation
    // ------------------------------------------------------------------------------------------------------------
    // Window of grid offsets, relative to the particle's (new) cell, that any trajectory ending in the cell can
    // touch: [-LO, UP] with the current solver margins (Esirkepov.hpp:42-45), WN = LO + UP + 1 points.
    template<int SHAPE>
    struct Win
    {
        static constexpr int WLO = CurrentMargin<SHAPE>::LO;
        static constexpr int WN = CurrentMargin<SHAPE>::LO + CurrentMargin<SHAPE>::UP + 1;
        // per particle(-segment) record in shared memory:
        //   float2 {S0, DS} [3 axes][WN]   then   C[3][WN-1] = prefix sums of DS * (-currentSurfaceDensity)
        static constexpr int REC = 6 * WN + 3 * (WN - 1);
        // record stride == 2 (mod 4) words: 8-byte aligned and conflict free for the 64-bit lane-strided stores
        static constexpr int RECP = REC + ((2 - REC % 4) + 4) % 4;
        // warp-private J tile: a warp owns one y-row of cells of the supercell
        static constexpr int PX = SCX + WN - 1, PY = WN, PZ = SCZ + WN - 1, PV = PX * PY * PZ;
    };

    /** Phase 1 helper: S0 and DS on the WN-point window of one axis, and the scaled prefix sums of DS.
     *  x0,x1: end points relative to their own assignment cell (on support); shift0/1: offset of that assignment
     *  cell from the particle cell.  Window index n <-> grid offset n - WLO. */
    template<int SHAPE>
    __device__ __forceinline__ void windowArrays(float2* __restrict__ sd, float* __restrict__ cpre, float x0, float x1, int shift0, int shift1, float factor)
    {
        using S = Shape<SHAPE>;
        using W = Win<SHAPE>;
        float a0[S::SUPP], a1[S::SUPP];
        S::on(x0, a0);
        S::on(x1, a1);
        float run = 0.0f;
#pragma unroll
        for(int n = 0; n < W::WN; ++n)
        {
            int const o = n - W::WLO;
            int const k0 = o - shift0 - S::BEGIN, k1 = o - shift1 - S::BEGIN;
            float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
            for(int s = 0; s < S::SUPP; ++s)
            {
                v0 = (k0 == s) ? a0[s] : v0;
                v1 = (k1 == s) ? a1[s] : v1;
            }
            float const ds = v1 - v0;
            sd[n] = make_float2(v0, ds);
            if(n < W::WN - 1)
            {
                run += ds; // accumulated_J recursion of Esirkepov.hpp:223-236, factored
                cpre[n] = run * factor;
            }
        }
    }

    template<int SHAPE, int SOLVER, int WARPS>
    __global__ void __launch_bounds__(WARPS * 32) depositCellKernel(DevParams P, SpeciesDev S, Field3 J, uint32_t const* __restrict__ cellOff)
    {
        using Sh = Shape<SHAPE>;
        using T = JTile<SHAPE>;
        using W = Win<SHAPE>;
        static_assert(WARPS == SCY, "one warp per y-row of the supercell");
        constexpr bool even = (Sh::SUPP % 2) == 0;
        constexpr int WN = W::WN;
        constexpr int NSEG = SOLVER == 0 ? 1 : 2; // EmZ: up to two on-support segments per particle
        static_assert(WN * WN <= 64, "transverse window must fit two nodes per lane");
        constexpr int NPL = (WN * WN + 31) / 32; // transverse nodes per lane

        extern __shared__ float smem[];
        float* tiles = smem; // WARPS * 3 * PV warp-private tiles
        float* recs = smem + WARPS * 3 * W::PV; // WARPS * 32 * RECP
        constexpr int CHUNK = 32 / NSEG; // particles per phase-1 pass: 32 records per warp

        int const sc = blockIdx.x;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        uint32_t const s0 = cellOff[sc * SCVOL], s1 = cellOff[(sc + 1) * SCVOL];
        if(s0 == s1)
            return;
        for(int i = threadIdx.x; i < WARPS * 3 * W::PV; i += blockDim.x)
            tiles[i] = 0.0f;
        __syncthreads();

        int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float* const myTile = tiles + warp * 3 * W::PV;
        float* const myRecs = recs + warp * 32 * W::RECP;
        float const rc2 = float(1.0 / double(P.c) / double(P.c));
        float const vol = P.cell[0] * P.cell[1] * P.cell[2];

        // transverse nodes owned by this lane
        int na[NPL], nb[NPL];
        bool nv[NPL];
#pragma unroll
        for(int q = 0; q < NPL; ++q)
        {
            int const node = lane + 32 * q;
            nv[q] = node < WN * WN;
            na[q] = nv[q] ? node % WN : 0;
            nb[q] = nv[q] ? node / WN : 0;
        }

        int const ly = warp;
        for(int cz = 0; cz < SCZ; ++cz)
            for(int cxl = 0; cxl < SCX; ++cxl)
            {
                int const lc = cxl + SCX * (ly + SCY * cz);
                uint32_t const c0 = cellOff[sc * SCVOL + lc], c1 = cellOff[sc * SCVOL + lc + 1];
                if(c0 == c1)
                    continue;
                // register accumulators: per owned transverse node WN-1 values along the current axis, 3 components
                float accX[NPL][WN - 1], accY[NPL][WN - 1], accZ[NPL][WN - 1];
#pragma unroll
                for(int q = 0; q < NPL; ++q)
#pragma unroll
                    for(int k = 0; k < WN - 1; ++k)
                        accX[q][k] = accY[q][k] = accZ[q][k] = 0.0f;

                for(uint32_t chunk = c0; chunk < c1; chunk += CHUNK)
                {
                    uint32_t const i = chunk + lane;
                    int const nIn = int(min(uint32_t(CHUNK), c1 - chunk));
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
                                // Esirkepov shifts both points by gridShift = min(iS,iE) and evaluates the
                                // off-support array, which is the on-support array of each point in its own
                                // assignment cell (shapeOff): same arithmetic, same bits.
                                int const gs = iS < iE ? iS : iE;
                                float const y0 = x0 - float(gs), y1 = x1 - float(gs);
                                float const f = (y0 == y1) ? 0.0f : -(csd * P.cell[d]);
                                windowArrays<SHAPE>(
                                    reinterpret_cast<float2*>(rec) + WN * d,
                                    rec + 6 * WN + (WN - 1) * d,
                                    gs != iS ? y0 - 1.0f : y0,
                                    gs != iE ? y1 - 1.0f : y1,
                                    iS,
                                    iE,
                                    f);
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
                            }
                            float const cd = charge / vol;
                            bool const two = sS[0] != sE[0] || sS[1] != sE[1] || sS[2] != sE[2];
                            float* rec2 = rec + W::RECP;
#pragma unroll
                            for(int d = 0; d < 3; ++d)
                            {
                                float const fd = -(P.cell[d] * cd / P.dt);
                                float const a0 = pS[d] - float(sS[d]), a1 = rl[d] - float(sS[d]);
                                windowArrays<SHAPE>(reinterpret_cast<float2*>(rec) + WN * d, rec + 6 * WN + (WN - 1) * d, a0, a1, sS[d], sS[d], (a0 == a1) ? 0.0f : fd);
                                float const b0 = rl[d] - float(sE[d]), b1 = pos[d] - float(sE[d]);
                                windowArrays<SHAPE>(reinterpret_cast<float2*>(rec2) + WN * d, rec2 + 6 * WN + (WN - 1) * d, b0, b1, sE[d], sE[d], (!two || b0 == b1) ? 0.0f : fd);
                            }
                        }
                    }
                    __syncwarp();
                    // ---- phase 2: lane = transverse node, loop over the particles of this chunk --------------------
                    for(int p = 0; p < nIn * NSEG; ++p)
                    {
                        float const* r = myRecs + p * W::RECP;
                        float2 const* rx = reinterpret_cast<float2 const*>(r);
                        float2 const* ry = rx + WN;
                        float2 const* rz = rx + 2 * WN;
                        float cx[WN - 1], cy[WN - 1], cz_[WN - 1];
#pragma unroll
                        for(int k = 0; k < WN - 1; ++k)
                        {
                            cx[k] = r[6 * WN + k];
                            cy[k] = r[6 * WN + (WN - 1) + k];
                            cz_[k] = r[6 * WN + 2 * (WN - 1) + k];
                        }
#pragma unroll
                        for(int q = 0; q < NPL; ++q)
                        {
                            int const a = na[q], b = nb[q];
                            float2 const xa = rx[a], ya = ry[a], yb = ry[b], zb = rz[b];
                            // transverse weights S0i*S0j + 1/2 (DSi*S0j + S0i*DSj) + 1/3 DSi*DSj, factored as
                            // S0i*(S0j + DSj/2) + DSi*(S0j/2 + DSj/3)
                            // Jx: (i,j) = (y,z) at node (a,b);  Jy: (z,x) at (x=a, z=b);  Jz: (x,y) at (a,b)
                            float const zP = zb.x + 0.5f * zb.y, zQ = 0.5f * zb.x + (1.0f / 3.0f) * zb.y;
                            float const xP = xa.x + 0.5f * xa.y, xQ = 0.5f * xa.x + (1.0f / 3.0f) * xa.y;
                            float const yP = yb.x + 0.5f * yb.y, yQ = 0.5f * yb.x + (1.0f / 3.0f) * yb.y;
                            float const tX = ya.x * zP + ya.y * zQ;
                            float const tY = zb.x * xP + zb.y * xQ;
                            float const tZ = xa.x * yP + xa.y * yQ;
#pragma unroll
                            for(int k = 0; k < WN - 1; ++k)
                            {
                                accX[q][k] += cx[k] * tX;
                                accY[q][k] += cy[k] * tY;
                                accZ[q][k] += cz_[k] * tZ;
                            }
                        }
                    }
                }
                // ---- per cell: add the register window to the warp-private tile (no atomics needed: within one
                // instruction all lanes address distinct nodes and no other warp touches this tile) ----------------
                __syncwarp();
#pragma unroll
                for(int q = 0; q < NPL; ++q)
                {
                    if(!nv[q])
                        continue;
                    int const a = na[q], b = nb[q];
#pragma unroll
                    for(int k = 0; k < WN - 1; ++k)
                    {
                        // window index n <-> private tile coordinate (cxl + n, n, cz + n) per axis
                        myTile[(cxl + k) + W::PX * (a + W::PY * (cz + b))] += accX[q][k]; // Jx: x=k, y=a, z=b
                        myTile[W::PV + (cxl + a) + W::PX * (k + W::PY * (cz + b))] += accY[q][k]; // Jy: x=a, y=k, z=b
                        myTile[2 * W::PV + (cxl + a) + W::PX * (b + W::PY * (cz + k))] += accZ[q][k]; // Jz: x=a, y=b, z=k
                    }
                }
                __syncwarp();
            }
        __syncthreads();
        // ---- combine the warp-private tiles and flush once to global J (red.global.add.f32) --------------------------
        {
            int const ox = scx * SCX + P.g[0] - T::LO, oy = scy * SCY + P.g[1] - T::LO, oz = scz * SCZ + P.g[2] - T::LO;
            for(int i = threadIdx.x; i < 3 * T::TV; i += blockDim.x)
            {
                int const comp = i / T::TV;
                int const r = i % T::TV;
                int const x = r % T::TX, y = (r / T::TX) % T::TY, z = r / (T::TX * T::TY);
                // block tile y = row + n  with row = warp index, n = window index in [0, WN)
                float v = 0.0f;
#pragma unroll
                for(int n = 0; n < WN; ++n)
                {
                    int const row = y - n;
                    if(row >= 0 && row < SCY)
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
            constexpr int WARPS = SCY;
            size_t const smem = sizeof(float) * (WARPS * 3 * Win<SHAPE>::PV + WARPS * 32 * Win<SHAPE>::RECP);
            cudaError_t e = cudaFuncSetAttribute(depositCellKernel<SHAPE, SOLVER, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if(e != cudaSuccess)
                return e;
            depositCellKernel<SHAPE, SOLVER, WARPS><<<nscTot, WARPS * 32, smem, st>>>(P, S, J, cellOff);
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
