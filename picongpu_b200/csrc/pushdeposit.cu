// pushdeposit.cu — the run kernel: charge conserving Esirkepov deposition over cell-sorted frame runs, optionally
// fused with gather + push + move (reference kernels K1 + K7: KernelMoveAndMarkParticles,
// include/picongpu/particles/Particles.kernel:170-316, and KernelComputeCurrent + Esirkepov,
// include/picongpu/fields/FieldJ.kernel:52-142, fields/currentDeposition/Esirkepov/Esirkepov.hpp:62-242).
//
// Why it looks the way it does (numbers measured on B200, see profiles/ and tools/microbench/lds_bcast.cu):
//   * shared-memory fp32 atomicAdd is a CAS loop (ATOMS.CAST.SPIN) and ATOMS costs ~2 cycles per lane, so the
//     reference's 54..144 shared atomics per particle cannot be the design.  Particles are cell sorted, so all
//     particles of a cell add to the same 4x4x4 node window: the sum over the particles of a cell is kept in
//     REGISTERS and written once per cell to a warp-private tile with plain LDS/FADD/STS.
//   * the shared memory return path delivers 256 B/cycle/SM (LDS.32 = 1, LDS.64 = 1, LDS.128 = 2 cycles per warp
//     instruction, independent of broadcast), so the per-particle record that phase 2 reads is laid out such that
//     a lane needs three 16-byte loads per record for twelve FMAs.
//
// One CTA (8 warps) per supercell; warp w owns the 32 consecutive cells [32w, 32w+32) of the supercell, i.e. a
// contiguous piece of the frame run.  It walks that piece in chunks of 32 particles, regardless of cell borders:
//   prologue: [FUSED] the E and the B tile of the supercell arrive by TMA (one 4-D box per field, tma.cuh) while
//     the CTA clears its private J tiles.
//   phase 1 (lane = particle): [FUSED: interpolate E,B from the shared tile, Boris/Vay/Higuera-Cary push, move,
//     emit the re-sort key and rank] then the 1-D assignment arrays of start and end point (same arithmetic as
//     Esirkepov.hpp:84-103) are turned into a record on the window [-1,2] around the anchor cell: per axis
//     {S0,DS}[4], [TSC, CIC, NGP: {P,Q}[4] = {S0+DS/2, S0/2+DS/3}[4]; PQS forms them in phase 2 so that two CTAs
//     still fit an SM] and C[3] = scaled prefix sums of DS (the accumulated_J recursion of Esirkepov.hpp:223-236,
//     factored out: J_k = C_k * transverse weight).
//   phase 2 (lane = (component, a-half, b-half), two records per pass): t(a,b) = S0_i(a) P_j(b) + DS_i(a) Q_j(b),
//     acc(a,b,k) += C_k t(a,b) with FFMA2 and broadcast scalar operands.  When the cell changes the two record
//     slots are combined with six SHFL and the 144 window values are added to the warp-private tile.
//   epilogue: the eight private tiles are summed row by row and flushed once per supercell with red.global.add.f32.
// Trajectories that do not fit the narrow window (more than half a cell per step for odd supports) are deposited by
// their own thread with global atomics, in the reference's loop order; a PQS particle that crosses a cell on one
// axis keeps the record and sends only the plane outside the window through global atomics.
// SOLVER = 1 is EmZ (EmZ.hpp:66-155): the trajectory is split at the relay point, each on-support segment is one
// record, the second segments are a second round of phase 2 over the same chunk.
// Anchor cell: the particle's cell (stand-alone deposit after the re-sort) or, FUSED, the cell it started in.
#include "common.cuh"
#include "esirkepov.cuh"
#include "pusher.cuh"
#include "shapes.cuh"
#include "tma.cuh"

namespace picstep
{
    template<int SHAPE>
    struct RunCfg
    {
        using Sh = Shape<SHAPE>;
        static constexpr int WN = 4; // narrow window: grid offsets -1..2 relative to the anchor cell
        static constexpr int WLO = 1;
        static constexpr int NK = WN - 1;
        static constexpr int FR = Sh::SUPP + 1; // entries of the off-support assignment arrays
        static constexpr int NMAX0 = WN - Sh::SUPP; // largest window index of frame entry 0 in a narrow record
        // record, per axis: {S0[0],S0[1],DS[0],DS[1]}, {S0[2],S0[3],DS[2],DS[3]}, {C[0],C[1],C[2],0}; the 33rd record is
        // all zero.  P = S0 + DS/2 and Q = S0/2 + DS/3 are formed in phase 2: the kernel is bound by the shared-memory
        // data pipe (profiles/), and the compact record costs 1.1 instead of 1.9 wavefronts per particle to write.
        // (Measured alternative, round 2: records that carry the transverse weights t(a,b) themselves, one record per
        // pass on 24 lanes -- 4 % fewer instructions, but 5.0 instead of 3.0 wavefronts per particle to read: 26.0
        // instead of 24.5 ms per launch.)
        static constexpr int COFF = 8, AXW = COFF + 4, RECW = 3 * AXW, NREC = 33;
        static constexpr int WARPS = 8, CELLS_PER_WARP = SCVOL / WARPS; // 32 cells: 8 x, 4 y, 1 z
        static constexpr int PX = SCX + WN - 1, PY = SCY / 2 + WN - 1, PZ = 1 + WN - 1, PV = PX * PY * PZ;
        // distance between the component planes of a private tile, padded: fewer bank conflicts of the per-cell flush
        // (tools/microbench/flush_banks.py; the 24 lanes of the three components cannot be made conflict free)
        static constexpr int PVC = PV + 4;
        static constexpr int TX = SCX + WN - 1, TY = SCY + WN - 1, TZ = SCZ + WN - 1, TV = TX * TY * TZ;
        static constexpr int EBW = Tile<SHAPE>::WORDS; // E/B tile words (FUSED): two 128-byte aligned TMA destinations
        static_assert(Sh::SUPP <= 4, "narrow window of 4 nodes needs a support of at most 4");
        static_assert((3 * PVC * WARPS) % 4 == 0 && RECW % 4 == 0, "records must stay 16-byte aligned");
    };

    template<int SHAPE, bool FUSED>
    constexpr size_t runSmemBytes()
    {
        using C = RunCfg<SHAPE>;
        return sizeof(float) * ((FUSED ? C::EBW : 0) + C::WARPS * 3 * C::PVC + C::WARPS * C::NREC * C::RECW);
    }

    template<int SHAPE, int PUSHER, bool FUSED, int SOLVER>
    __global__ void __launch_bounds__(256, 2) runKernel(
        DevParams P,
        SpeciesDev S, // attributes are read from here: slot j of the cell-sorted order lives at index inv[j] (j if inv is null)
        SpeciesDev D, // FUSED: the pushed attributes are written here at index j (the other buffer: lazy re-sort)
        uint32_t const* __restrict__ inv,
        Field3 E,
        Field3 B,
        Field3 J,
        uint32_t const* __restrict__ cellOff,
        uint32_t* __restrict__ cellCnt, // FUSED: per destination cell, number of particles arriving from other cells
        uint32_t* __restrict__ stayCnt, // FUSED: per cell, number of particles that stay in it
        uint32_t* __restrict__ key, // FUSED: destination supercell * 256 + cell (+ leave flags)
        uint32_t* __restrict__ rank, // FUSED: slot inside the destination cell: stayers first, then arrivals (bit 31)
        const __grid_constant__ TileMaps maps, // FUSED: TMA descriptors of E and B
        ScArea area) // the supercells this launch covers
    {
        using Sh = Shape<SHAPE>;
        using C = RunCfg<SHAPE>;
        using T = Tile<SHAPE>;
        constexpr bool even = (Sh::SUPP % 2) == 0;
        constexpr uint32_t FULL = 0xffffffffu;

        extern __shared__ __align__(128) float smem[];
        __shared__ uint64_t ebBar;
        float* const ebTile = smem;
        float* const tiles = smem + (FUSED ? C::EBW : 0);
        float* const recs = tiles + C::WARPS * 3 * C::PVC;

        int sc3[3];
        {
            int const a1 = (area.axis + 1) % 3, a2 = (area.axis + 2) % 3;
            int b = blockIdx.x;
            sc3[a1] = b % P.nsc[a1];
            b /= P.nsc[a1];
            sc3[a2] = b % P.nsc[a2];
            sc3[area.axis] = area.first + (b / P.nsc[a2]) * area.stride;
        }
        int const scx = sc3[0], scy = sc3[1], scz = sc3[2];
        int const sc = scx + P.nsc[0] * (scy + P.nsc[1] * scz);
        uint32_t const scBeg = cellOff[sc * SCVOL], scEnd = cellOff[(sc + 1) * SCVOL];
        if(scBeg == scEnd)
            return;
        int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float* const myTile = tiles + warp * 3 * C::PVC;
        float* const myRecs = recs + warp * C::NREC * C::RECW;

        if constexpr(FUSED)
        {
            // stage the E and B tiles: one TMA box (x, y, z, 3 components) per field, issued by one thread before the
            // CTA clears its J tiles and records, completion is awaited after the barrier
            if(threadIdx.x == 0)
            {
                int const ox = scx * SCX + P.g[0] - T::LO + maps.lead, oy = scy * SCY + P.g[1] - T::LO, oz = scz * SCZ + P.g[2] - T::LO;
                mbarInit(&ebBar, 1);
                mbarExpectTx(&ebBar, 2 * T::BYTES_PER_FIELD);
                tmaLoadTile(ebTile, &maps.B, ox, oy, oz, &ebBar);
                tmaLoadTile(ebTile + T::HALF, &maps.E, ox, oy, oz, &ebBar);
            }
        }
        {
            static_assert((C::WARPS * 3 * C::PVC) % 4 == 0, "tiles are cleared with 16-byte stores");
            float4* const t4 = reinterpret_cast<float4*>(tiles);
            for(int i = threadIdx.x; i < C::WARPS * 3 * C::PVC / 4; i += blockDim.x)
                t4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            for(int i = lane; i < C::RECW; i += 32) // record 32 stays all zero
                myRecs[32 * C::RECW + i] = 0.0f;
        }
        __syncthreads();
        if constexpr(FUSED)
            mbarWait(&ebBar, 0);

        float const rc2 = float(1.0 / double(P.c) / double(P.c));
        // displacement in cells = v dt / cellSize: the exact build divides as the reference does (MoveParticle's caller,
        // Esirkepov.hpp:84-90), the production build multiplies with the reciprocal of the (uniform) cell size
#ifdef PICSTEP_EXACT
#    define PS_DIV_CELL(a, d) ((a) / P.cell[d])
#else
        float const rcell[3] = {1.0f / P.cell[0], 1.0f / P.cell[1], 1.0f / P.cell[2]};
#    define PS_DIV_CELL(a, d) ((a) * rcell[d])
#endif
        float const vol = P.cell[0] * P.cell[1] * P.cell[2];
        float const* const tB = ebTile;
        float const* const tE = ebTile + T::HALF;

        // ---- phase 2 lane constants ------------------------------------------------------------------------------
        int const slot = lane >> 4, g = lane & 15;
        bool const p2active = g < 12;
        int const comp = p2active ? (g >> 2) : 0;
        int const ah = (g >> 1) & 1, bh = g & 1;
        int const ai = (comp + 1) % 3, aj = (comp + 2) % 3; // Jx: (i,j) = (y,z); Jy: (z,x); Jz: (x,y)
        // a lane reads {S0,DS} of axis i at nodes 2ah,2ah+1, {S0,DS} of axis j at nodes 2bh,2bh+1 and C of its component
        int const offSD = ai * C::AXW + 4 * ah;
        int const offPQ = aj * C::AXW + 4 * bh;
        int const offC = comp * C::AXW + C::COFF;
        auto strideOf = [](int a) { return a == 0 ? 1 : (a == 1 ? C::PX : C::PX * C::PY); };
        int const sC = strideOf(comp), sJ = strideOf(aj);
        int const laneTile = comp * C::PVC + (2 * ah + slot) * strideOf(ai) + 2 * bh * sJ;

        // accumulators of the lane's 2 x 2 x 3 nodes, packed over b: acc[a][k] = {J(a, b=0, k), J(a, b=1, k)}
        F2 acc[2][C::NK];
#pragma unroll
        for(int a = 0; a < 2; ++a)
#pragma unroll
            for(int k = 0; k < C::NK; ++k)
                acc[a][k] = F2(0.0f);

        int curCell = -1; // local cell index (0..255) the accumulators belong to
        uint32_t stayCarry = 0; // FUSED: stayers of curCell seen in earlier chunks (their rank in the re-sorted cell run)

        // adds the accumulators to the private tile and clears them
        auto flushCell = [&]()
        {
            __syncwarp();
            int const cx = curCell & (SCX - 1), cyl = (curCell >> 3) & (SCY / 2 - 1);
            int const cellBase = cx + C::PX * cyl + laneTile;
#pragma unroll
            for(int b = 0; b < 2; ++b)
#pragma unroll
                for(int k = 0; k < C::NK; ++k)
                {
                    float const a0 = b ? acc[0][k].y : acc[0][k].x, a1 = b ? acc[1][k].y : acc[1][k].x;
                    float const keep = slot ? a1 : a0;
                    float const send = slot ? a0 : a1;
                    float const v = keep + __shfl_xor_sync(FULL, send, 16);
                    if(p2active)
                        myTile[cellBase + b * sJ + k * sC] += v;
                }
#pragma unroll
            for(int a = 0; a < 2; ++a)
#pragma unroll
                for(int k = 0; k < C::NK; ++k)
                    acc[a][k] = F2(0.0f);
        };

        // record from the assignment values of start (S0) and end point (S1) on the window, the factors f and the
        // start value cm of the prefix sums (the node left of the window, SEMI only)
        auto storeRecord = [&](float* rec, float const(&S0)[3][C::WN], float const(&S1)[3][C::WN], float const(&f)[3], float const(&cm)[3])
        {
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                float4* r4 = reinterpret_cast<float4*>(rec + d * C::AXW);
                F2 DS[2];
#pragma unroll
                for(int j = 0; j < 2; ++j)
                {
                    F2 const s0p(S0[d][2 * j], S0[d][2 * j + 1]);
                    DS[j] = F2(S1[d][2 * j], S1[d][2 * j + 1]) - s0p;
                    r4[j] = make_float4(s0p.x, s0p.y, DS[j].x, DS[j].y);
                }
                float const c0 = cm[d] + DS[0].x, c1 = c0 + DS[0].y, c2 = c1 + DS[1].x;
                r4[C::COFF / 4] = make_float4(c0 * f[d], c1 * f[d], c2 * f[d], 0.0f);
            }
        };

        uint32_t const pBeg = cellOff[sc * SCVOL + warp * C::CELLS_PER_WARP];
        uint32_t const pEnd = cellOff[sc * SCVOL + (warp + 1) * C::CELLS_PER_WARP];

        // EmZ: record of one on-support segment (start and end values on the same SUPP nodes at window offset o)
        [[maybe_unused]] auto emzRecord = [&](float* rec, F2 const(&t)[3][Sh::SUPP], int const(&o)[3], float const(&f)[3])
        {
            float S0[3][C::WN], S1[3][C::WN];
#pragma unroll
            for(int d = 0; d < 3; ++d)
#pragma unroll
                for(int n = 0; n < C::WN; ++n)
                {
                    float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
                    for(int m = 0; m <= C::NMAX0; ++m)
                    {
                        int const sx = n - m;
                        if(sx >= 0 && sx < Sh::SUPP)
                        {
                            v0 = (o[d] == m) ? t[d][sx].x : v0;
                            v1 = (o[d] == m) ? t[d][sx].y : v1;
                        }
                    }
                    S0[d][n] = v0;
                    S1[d][n] = v1;
                }
            float const cm[3] = {0.0f, 0.0f, 0.0f};
            storeRecord(rec, S0, S1, f, cm);
        };

        // The particle attributes of a chunk are loaded one chunk ahead: the loads are issued before phase 2 of the
        // previous chunk, which does not touch global memory, so the HBM latency is hidden by it.
        // The attributes are addressed through the permutation of the previous step's re-sort (inv), whose entries
        // are loaded two chunks ahead.
        float pfx[3] = {0.f, 0.f, 0.f}, pfu[3] = {0.f, 0.f, 0.f}, pfw = 0.f;
        int pfc = -2;
        uint32_t pfSrc = 0;
        auto prefetchIdx = [&](uint32_t chunk_)
        {
            uint32_t const j = chunk_ + lane;
            pfSrc = j;
            if(inv != nullptr && j < pEnd)
                pfSrc = __ldcs(inv + j);
        };
        auto prefetch = [&](uint32_t chunk_)
        {
            uint32_t const j = chunk_ + lane;
            if(j < pEnd)
            {
                uint32_t const src = pfSrc;
#pragma unroll
                for(int d = 0; d < 3; ++d)
                {
                    pfx[d] = S.pos[d][src];
                    pfu[d] = S.mom[d][src];
                }
                pfw = S.w[src];
                pfc = S.cell[j];
            }
        };
        prefetchIdx(pBeg);
        prefetch(pBeg);
        prefetchIdx(pBeg + 32);

        for(uint32_t chunk = pBeg; chunk < pEnd; chunk += 32)
        {
            uint32_t const i = chunk + lane;
            bool const valid = i < pEnd;
            int lc = -2;
            bool useRec = false; // this lane wrote a record phase 2 has to add
            bool stays = false;
            uint32_t myRank = 0;
            // An arrival's rank is the return value of a global atomic.  Nothing may touch that register before the
            // store at the end of the chunk: an OR right behind the atomic made every warp with an arrival wait for the
            // L2 round trip there (ncu: 8.5 % of all stall samples on that one LOP3, long scoreboard).
            [[maybe_unused]] bool arrives = false;
            // EmZ: a trajectory that changes its assignment cell is split at the relay point; the second segment is
            // kept here while phase 2 adds the first one
            [[maybe_unused]] F2 tSeg2[3][Sh::SUPP];
            [[maybe_unused]] float fSeg2[3] = {0.0f, 0.0f, 0.0f};
            [[maybe_unused]] int oSeg2[3] = {0, 0, 0};
            [[maybe_unused]] bool twoSeg = false;
            __syncwarp(); // phase 2 of the previous chunk has finished reading the records
            // ---- phase 1: lane = particle -----------------------------------------------------------------------
            if(valid)
            {
                float x1[3] = {pfx[0], pfx[1], pfx[2]};
                float u[3] = {pfu[0], pfu[1], pfu[2]};
                float const w = pfw;
                lc = pfc;
                int const lx = lc % SCX, ly = (lc / SCX) % SCY, lz = lc / (SCX * SCY);
                float const mass = S.mass_per_w * w;
                float const charge = S.charge_per_w * w;
                int dir[3] = {0, 0, 0};
                bool deposit = true;
                float vel[3];
                [[maybe_unused]] float dpv[3] = {0.0f, 0.0f, 0.0f}; // FUSED: displacement of the move, in cells
                if constexpr(FUSED)
                {
                    float Bf[3], Ef[3];
                    gatherEB<SHAPE>(tB, tE, lx, ly, lz, x1[0], x1[1], x1[2], Ef, Bf);
                    pushMomentum<PUSHER>(P, rc2, mass, charge, Ef, Bf, u);
                    velocityOf(rc2, mass, u[0], u[1], u[2], vel[0], vel[1], vel[2]);
                    // moveParticle (MoveParticle.hpp:48-160)
#pragma unroll
                    for(int d = 0; d < 3; ++d)
                    {
                        dpv[d] = PS_DIV_CELL(vel[d] * P.dt, d);
                        float q = (x1[d] + dpv[d]) - 0.5f;
                        float mv = 0.0f;
                        if(q < -0.5f)
                            mv = -1.0f;
                        if(q >= 0.5f)
                            mv = 1.0f;
                        q -= mv;
                        x1[d] = q + 0.5f;
                        dir[d] = int(mv);
                        D.pos[d][i] = x1[d];
                        D.mom[d][i] = u[d];
                    }
                    D.w[i] = w;
                    // re-sort key (see pushKernel); fast path: the particle stays inside this supercell
                    int const nl[3] = {lx + dir[0], ly + dir[1], lz + dir[2]};
                    uint32_t k;
                    if(unsigned(nl[0]) < unsigned(SCX) && unsigned(nl[1]) < unsigned(SCY) && unsigned(nl[2]) < unsigned(SCZ))
                    {
                        k = uint32_t(sc) * SCVOL + uint32_t(nl[0] + SCX * (nl[1] + SCY * nl[2]));
                        if((dir[0] | dir[1] | dir[2]) == 0)
                            stays = true; // ranked per run of equal cells in phase 2
                        else
                        {
#ifdef PICSTEP_RANK_EAGER
                            myRank = 0x80000000u | atomicAdd(&cellCnt[k], 1u);
#else
                            myRank = atomicAdd(&cellCnt[k], 1u);
                            arrives = true;
#endif
                        }
                    }
                    else
                    {
                        int gc[3] = {scx * SCX + nl[0], scy * SCY + nl[1], scz * SCZ + nl[2]};
                        uint32_t flag = 0u;
                        bool drop = false;
#pragma unroll
                        for(int d = 0; d < 3; ++d)
                        {
                            if(gc[d] < 0 || gc[d] >= P.n[d])
                            {
                                bool const up = gc[d] >= P.n[d];
                                if(P.wrap[d])
                                    gc[d] += up ? -P.n[d] : P.n[d];
                                else if(d == P.split_axis && (up ? P.has_upper : P.has_lower))
                                {
                                    flag = KEY_LEAVE | (up ? KEY_UPPER : 0u);
                                    gc[d] += up ? -P.n[d] : P.n[d]; // coordinate in the receiver's local grid
                                }
                                else
                                    drop = true;
                            }
                        }
                        if(drop)
                        {
                            k = KEY_DROP;
                            deposit = false; // absorbed particles are deleted before the current deposition
                        }
                        else
                        {
                            int const dsc = (gc[0] >> 3) + P.nsc[0] * ((gc[1] >> 3) + P.nsc[1] * (gc[2] >> 2));
                            int const dlc = (gc[0] & 7) + SCX * ((gc[1] & 7) + SCY * (gc[2] & 3));
                            k = uint32_t(dsc) * SCVOL + uint32_t(dlc);
                            if(flag)
                                k |= flag;
                            else
                            {
#ifdef PICSTEP_RANK_EAGER
                                myRank = 0x80000000u | atomicAdd(&cellCnt[k], 1u);
#else
                                myRank = atomicAdd(&cellCnt[k], 1u);
                                arrives = true;
#endif
                            }
                        }
                    }
                    key[i] = k;
                }
                else
                    velocityOf(rc2, mass, u[0], u[1], u[2], vel[0], vel[1], vel[2]);

                if constexpr(SOLVER == 1)
                {
                    // EmZ (EmZ.hpp:66-155, EmZ/DepositCurrent.hpp:35-119): segment A = start -> relay point in the frame of
                    // the start point's assignment cell, segment B = relay point -> end in the frame of the end point's;
                    // each is an Esirkepov-type deposit with both points on the same support, i.e. one record each.
                    if(deposit)
                    {
                        float const csd = charge * (1.0f / float(vol * P.dt));
                        F2 tA[3][Sh::SUPP];
                        float fA[3];
                        int oA[3], cA[3], cB[3];
                        float a0[3], a1[3], b0[3], b1[3];
                        bool narrow = true, two = false;
#pragma unroll
                        for(int d = 0; d < 3; ++d)
                        {
                            float const dp = FUSED ? dpv[d] : PS_DIV_CELL(vel[d] * P.dt, d);
                            float const xe = x1[d], xs = xe - dp;
                            int iS, iE;
                            float const r = relay<even>(iS, iE, xs, xe);
                            a0[d] = xs - float(iS);
                            a1[d] = r - float(iS);
                            b0[d] = r - float(iE);
                            b1[d] = xe - float(iE);
                            Sh::on(F2(a0[d], a1[d]), tA[d]);
                            Sh::on(F2(b0[d], b1[d]), tSeg2[d]);
                            fA[d] = (a0[d] == a1[d]) ? 0.0f : -(csd * P.cell[d]);
                            fSeg2[d] = (b0[d] == b1[d]) ? 0.0f : -(csd * P.cell[d]);
                            oA[d] = iS + Sh::BEGIN + C::WLO + dir[d];
                            oSeg2[d] = iE + Sh::BEGIN + C::WLO + dir[d];
                            if(oA[d] < 0 || oA[d] > C::NMAX0 || oSeg2[d] < 0 || oSeg2[d] > C::NMAX0)
                                narrow = false;
                            two = two || iS != iE;
                            cA[d] = iS;
                            cB[d] = iE;
                        }
                        if(narrow)
                        {
                            useRec = true;
                            twoSeg = two;
                            emzRecord(myRecs + lane * C::RECW, tA, oA, fA);
                        }
                        else
                        {
                            // wide trajectory: the reference loops with global atomics, per segment
                            atomicAdd(P.stats, 1ull);
                            long long const sY = P.N[0], sZ = (long long) P.N[0] * P.N[1];
                            long long const oa = fidx(P, scx * SCX + P.g[0] + lx + dir[0] + cA[0], scy * SCY + P.g[1] + ly + dir[1] + cA[1], scz * SCZ + P.g[2] + lz + dir[2] + cA[2]);
                            emzSegmentGlobal<SHAPE>(J.c[0] + oa, J.c[1] + oa, J.c[2] + oa, sY, sZ, a0[0], a0[1], a0[2], a1[0], a1[1], a1[2], csd * P.cell[0], csd * P.cell[1], csd * P.cell[2]);
                            if(two)
                            {
                                long long const ob = fidx(P, scx * SCX + P.g[0] + lx + dir[0] + cB[0], scy * SCY + P.g[1] + ly + dir[1] + cB[1], scz * SCZ + P.g[2] + lz + dir[2] + cB[2]);
                                emzSegmentGlobal<SHAPE>(J.c[0] + ob, J.c[1] + ob, J.c[2] + ob, sY, sZ, b0[0], b0[1], b0[2], b1[0], b1[1], b1[2], csd * P.cell[0], csd * P.cell[1], csd * P.cell[2]);
                            }
                        }
                    }
                }
                else if(deposit)
                {
                    // Esirkepov.hpp:84-103: start and end point in the frame of gridShift = min(iS,iE); the on-support
                    // assignment values of both points are evaluated together (packed: .x start, .y end) and placed
                    // on the window: value s of the start point sits at window index o0 + s, o0 = iS + BEGIN + WLO + dir
                    float const csd = charge * (1.0f / float(vol * P.dt));
                    F2 t[3][Sh::SUPP];
                    float f[3], p0[3], p1[3];
                    int o0[3], o1[3], gs3[3], status[3];
                    // A support as wide as the window (PQS) leaves no room for a cell crossing.  SEMI: such a particle
                    // still goes through the records when it crosses on ONE axis: the window part as usual, the one
                    // plane of nodes outside the window (40 values) with global atomics instead of the whole particle
                    // (184); this took the PQS step from 161 to 120 ms at 256^3 (12 % of the KHI particles cross a cell per step).
                    constexpr bool SEMI = (Sh::SUPP == C::WN);
                    constexpr int MLO = SEMI ? -1 : 0, MHI = C::NMAX0 + (SEMI ? 1 : 0);
                    int ext[3] = {0, 0, 0}, nExt = 0;
                    bool narrow = true;
#pragma unroll
                    for(int d = 0; d < 3; ++d)
                    {
                        float const dp = FUSED ? dpv[d] : PS_DIV_CELL(vel[d] * P.dt, d);
                        float const xe = x1[d], xs = xe - dp;
                        int iS, iE;
                        relay<even>(iS, iE, xs, xe);
                        int const gs = iS < iE ? iS : iE;
                        float const y0 = xs - float(gs), y1 = xe - float(gs);
                        Sh::on(F2(gs != iS ? y0 - 1.0f : y0, gs != iE ? y1 - 1.0f : y1), t[d]);
                        f[d] = (y0 == y1) ? 0.0f : -(csd * P.cell[d]);
                        o0[d] = iS + Sh::BEGIN + C::WLO + dir[d];
                        o1[d] = iE + Sh::BEGIN + C::WLO + dir[d];
                        if(o0[d] < 0 || o0[d] > C::NMAX0 || o1[d] < 0 || o1[d] > C::NMAX0)
                        {
                            int const lo = o0[d] < o1[d] ? o0[d] : o1[d], hi = o0[d] < o1[d] ? o1[d] : o0[d];
                            if(!SEMI || lo < MLO || hi > MHI || (lo < 0 && hi > C::NMAX0))
                                narrow = false;
                            ext[d] = lo < 0 ? -1 : 1;
                            ++nExt;
                        }
                        p0[d] = y0;
                        p1[d] = y1;
                        gs3[d] = gs;
                        status[d] = (gs == iS ? 2 : 0) | (gs == iE ? 4 : 0) | (iS != iE ? 1 : 0);
                    }
                    float* const rec = myRecs + lane * C::RECW;
                    if(nExt > 1)
                        narrow = false;
                    // SEMI: assignment values of start / end point at the node outside the window, its axis and side
                    float sOut0 = 0.0f, sOut1 = 0.0f, fOut = 0.0f;
                    int axOut = -1, sideOut = 0;
                    if(narrow)
                    {
                        useRec = true;
                        float S0[3][C::WN], S1[3][C::WN];
#pragma unroll
                        for(int d = 0; d < 3; ++d)
                        {
                            if constexpr(SEMI)
                                if(ext[d] != 0)
                                {
                                    int const nout = ext[d] < 0 ? -1 : C::WN;
#pragma unroll
                                    for(int sx = 0; sx < Sh::SUPP; ++sx)
                                    {
                                        sOut0 = (nout - o0[d] == sx) ? t[d][sx].x : sOut0;
                                        sOut1 = (nout - o1[d] == sx) ? t[d][sx].y : sOut1;
                                    }
                                    axOut = d;
                                    sideOut = ext[d];
                                    fOut = f[d];
                                }
#pragma unroll
                            for(int n = 0; n < C::WN; ++n)
                            {
                                float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
                                for(int m = MLO; m <= MHI; ++m)
                                {
                                    int const s = n - m;
                                    if(s >= 0 && s < Sh::SUPP)
                                    {
                                        v0 = (o0[d] == m) ? t[d][s].x : v0;
                                        v1 = (o1[d] == m) ? t[d][s].y : v1;
                                    }
                                }
                                S0[d][n] = v0;
                                S1[d][n] = v1;
                            }
                        }
                        // prefix sums start at the node left of the window when that one carries weight
                        float cm[3];
#pragma unroll
                        for(int d = 0; d < 3; ++d)
                            cm[d] = (SEMI && ext[d] < 0) ? sOut1 - sOut0 : 0.0f;
                        storeRecord(rec, S0, S1, f, cm);
                        if constexpr(SEMI)
                            if(axOut >= 0)
                            {
                                atomicAdd(P.stats + 1, 1ull);
                                // the plane of nodes outside the window: J_A at the node next to it along A (16 values)
                                // and the two transverse components on the plane itself (2 x 12 values)
                                int const A = axOut, iA = (A + 1) % 3, jA = (A + 2) % 3;
                                float const dsOut = sOut1 - sOut0;
                                float const pOut = sOut0 + 0.5f * dsOut, qOut = 0.5f * sOut0 + (1.0f / 3.0f) * dsOut;
                                long long const strd[3] = {1, P.N[0], (long long) P.N[0] * P.N[1]};
                                long long const base = fidx(P, scx * SCX + P.g[0] + lx - C::WLO, scy * SCY + P.g[1] + ly - C::WLO, scz * SCZ + P.g[2] + lz - C::WLO);
                                auto jOf = [&](int c) { return c == 0 ? J.c[0] : (c == 1 ? J.c[1] : J.c[2]); };
                                // off 0: S0, 2: DS of the record; 8: P = S0 + DS/2, 10: Q = S0/2 + DS/3
                                auto at = [&](int d, int off, int n)
                                {
                                    float const* q = rec + d * C::AXW + (n >> 1) * 4 + (n & 1);
                                    float const s0 = q[0], ds = q[2];
                                    return off == 0 ? s0 : (off == 2 ? ds : (off == 8 ? s0 + 0.5f * ds : 0.5f * s0 + (1.0f / 3.0f) * ds));
                                };
                                int const nout = sideOut < 0 ? -1 : C::WN, kout = sideOut < 0 ? -1 : C::WN - 1;
                                float const cx = sideOut < 0 ? fOut * dsOut : -(fOut * dsOut);
                                float* const ja = jOf(A) + base + kout * strd[A];
#pragma unroll 1
                                for(int a = 0; a < C::WN; ++a)
#pragma unroll
                                    for(int b = 0; b < C::WN; ++b)
                                        redGlobal(ja + a * strd[iA] + b * strd[jA], cx * (at(iA, 0, a) * at(jA, 8, b) + at(iA, 2, a) * at(jA, 10, b)));
                                // component iA = (A+1)%3: its (i, j) axes are (jA, A), the outside node sits on j
                                float* const j1 = jOf(iA) + base + nout * strd[A];
                                // component jA = (A+2)%3: its (i, j) axes are (A, iA), the outside node sits on i
                                float* const j2 = jOf(jA) + base + nout * strd[A];
#pragma unroll 1
                                for(int a = 0; a < C::WN; ++a)
                                {
                                    float const t1 = at(jA, 0, a) * pOut + at(jA, 2, a) * qOut;
                                    float const t2 = sOut0 * at(iA, 8, a) + dsOut * at(iA, 10, a);
#pragma unroll
                                    for(int k = 0; k < C::NK; ++k)
                                    {
                                        redGlobal(j1 + a * strd[jA] + k * strd[iA], rec[iA * C::AXW + C::COFF + k] * t1);
                                        redGlobal(j2 + a * strd[iA] + k * strd[jA], rec[jA * C::AXW + C::COFF + k] * t2);
                                    }
                                }
                            }
                    }
                    else
                    {
                        // wide trajectory: reference loop with global atomics, in the frame of the particle's new cell
                        atomicAdd(P.stats, 1ull);
                        int const baseG[3]
                            = {scx * SCX + P.g[0] + lx + dir[0] + gs3[0], scy * SCY + P.g[1] + ly + dir[1] + gs3[1], scz * SCZ + P.g[2] + lz + dir[2] + gs3[2]};
                        long long const origin = fidx(P, baseG[0], baseG[1], baseG[2]);
                        esirkepovParticleGlobal<SHAPE>(J.c[0] + origin, J.c[1] + origin, J.c[2] + origin, P.N[0], (long long) P.N[0] * P.N[1], status[0] | (status[1] << 3) | (status[2] << 6), p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], csd * P.cell[0], csd * P.cell[1], csd * P.cell[2]);
                    }
                }
            }
            if(!useRec)
            {
                // nothing to add for this record in phase 2 (absorbed / wide trajectory / slot beyond the end of the
                // piece): C = 0.  The other words are stale but finite, except for a slot that has never been
                // written (first chunk of the piece): that one is cleared completely.
                float* const rec = myRecs + lane * C::RECW;
                if(chunk == pBeg)
                {
#pragma unroll
                    for(int q = 0; q < C::RECW / 4; ++q)
                        reinterpret_cast<float4*>(rec)[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
                else
                {
#pragma unroll
                    for(int d = 0; d < 3; ++d)
                        *reinterpret_cast<float4*>(rec + d * C::AXW + C::COFF) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
            }
            // ---- phase 2: the records of the chunk are added to the accumulators, which are flushed when the cell changes
            uint32_t const validMask = __ballot_sync(FULL, valid);
            int const n = __popc(validMask);
            int prev = __shfl_up_sync(FULL, lc, 1);
            if(lane == 0)
                prev = curCell;
            uint32_t const startMask = __ballot_sync(FULL, valid && lc != prev); // a new cell starts at this record
            if constexpr(FUSED)
            {
                // Stayers keep their relative order: rank = stayers of my cell before me.  Evaluated per lane from
                // the ballots: my run of equal cells starts at the highest start bit at or below my lane, or in an
                // earlier chunk (then stayCarry stayers precede this chunk).
                uint32_t const stayMask = __ballot_sync(FULL, stays);
                uint32_t const le = FULL >> (31 - lane); // bits 0..lane
                uint32_t const sBelow = startMask & le;
                int const lo = sBelow ? 31 - __clz(sBelow) : 0;
                uint32_t const incl = (sBelow ? 0u : stayCarry) + __popc(stayMask & le & (FULL << lo));
                if(stays)
                    myRank = incl - 1u;
                if(valid && lane < 31 && ((startMask >> (lane + 1)) & 1u)) // my cell ends with me
                    stayCnt[sc * SCVOL + lc] = incl;
                if(lane == 0 && (startMask & 1u) && curCell >= 0) // the cell left open by the previous chunk has ended
                    stayCnt[sc * SCVOL + curCell] = stayCarry;
                stayCarry = __shfl_sync(FULL, incl, n - 1);
            }
            prefetch(chunk + 32);
            prefetchIdx(chunk + 64);
            __syncwarp(); // records are visible
            // runs the records 0..n-1 of the chunk through the accumulators; sMask: a new cell starts at this record.
            // Two records per pass (one per half warp): t(a,b) = S0_i(a) P_j(b) + DS_i(a) Q_j(b), acc(a,b,k) += C_k t(a,b);
            // packed over b (FFMA2 with a broadcast scalar operand)
            auto runPasses = [&](uint32_t sMask)
            {
                auto pass = [&](float const* pSD, float const* pPQ, float const* pC)
                {
                    float4 const sd = *reinterpret_cast<float4 const*>(pSD); // {S0[a0], S0[a0+1], DS[a0], DS[a0+1]}
                    float4 const sj = *reinterpret_cast<float4 const*>(pPQ); // {S0[b0], S0[b0+1], DS[b0], DS[b0+1]}
                    float4 const c4 = *reinterpret_cast<float4 const*>(pC);
                    F2 const s0j(sj.x, sj.y), dsj(sj.z, sj.w);
                    F2 const pp = fma2(dsj, F2(0.5f), s0j); // P = S0 + DS/2
                    F2 const qq = fma2(dsj, F2(1.0f / 3.0f), s0j * F2(0.5f)); // Q = S0/2 + DS/3
                    F2 const t0 = fma2(F2(sd.z), qq, F2(sd.x) * pp);
                    F2 const t1 = fma2(F2(sd.w), qq, F2(sd.y) * pp);
                    acc[0][0] = fma2(F2(c4.x), t0, acc[0][0]);
                    acc[0][1] = fma2(F2(c4.y), t0, acc[0][1]);
                    acc[0][2] = fma2(F2(c4.z), t0, acc[0][2]);
                    acc[1][0] = fma2(F2(c4.x), t1, acc[1][0]);
                    acc[1][1] = fma2(F2(c4.y), t1, acc[1][1]);
                    acc[1][2] = fma2(F2(c4.z), t1, acc[1][2]);
                };
                float const* const zC = myRecs + 32 * C::RECW + offC; // C words of the all-zero record
                // records q, q+1 with a cell change at q (bit 0) and / or at q+1 (bit 1)
                auto passSlow = [&](float const* pSD, float const* pPQ, float const* pC, uint32_t bits, int q)
                {
                    if(bits & 1u)
                    {
                        if(curCell >= 0)
                            flushCell();
                        curCell = __shfl_sync(FULL, lc, q);
                    }
                    if(bits & 2u)
                    {
                        pass(pSD, pPQ, slot ? zC : pC); // record q closes its cell
                        flushCell();
                        curCell = __shfl_sync(FULL, lc, q + 1);
                        pass(pSD, pPQ, slot ? pC : zC); // record q+1 opens the next one
                    }
                    else
                        pass(pSD, pPQ, pC);
                };
                float const* rec = myRecs + slot * C::RECW;
                float const *pSD = rec + offSD, *pPQ = rec + offPQ, *pC = rec + offC;
                // groups of eight records (records beyond n carry C = 0): a group without a cell change is four passes of
                // straight line code, otherwise every pair of records is looked at
#pragma unroll 1
                for(int r = 0; r < n; r += 8)
                {
                    uint32_t const b8 = (sMask >> r) & 0xffu;
                    if(b8 == 0u)
                    {
#pragma unroll
                        for(int q = 0; q < 8; q += 2)
                            pass(pSD + q * C::RECW, pPQ + q * C::RECW, pC + q * C::RECW);
                    }
                    else
                    {
#pragma unroll
                        for(int q = 0; q < 8; q += 2)
                        {
                            uint32_t const b2 = (b8 >> q) & 3u;
                            if(b2 == 0u)
                                pass(pSD + q * C::RECW, pPQ + q * C::RECW, pC + q * C::RECW);
                            else
                                passSlow(pSD + q * C::RECW, pPQ + q * C::RECW, pC + q * C::RECW, b2, r + q);
                        }
                    }
                    pSD += 8 * C::RECW;
                    pPQ += 8 * C::RECW;
                    pC += 8 * C::RECW;
                }
            };
            runPasses(startMask);
            if constexpr(SOLVER == 1)
            {
                // second EmZ segment of the trajectories that changed their assignment cell: the same records again
                if(__ballot_sync(FULL, twoSeg))
                {
                    __syncwarp(); // the first round has read the records
                    float* const rec2 = myRecs + lane * C::RECW;
                    if(twoSeg)
                        emzRecord(rec2, tSeg2, oSeg2, fSeg2);
                    else
                    {
#pragma unroll
                        for(int d = 0; d < 3; ++d)
                            *reinterpret_cast<float4*>(rec2 + d * C::AXW + C::COFF) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    }
                    __syncwarp();
                    // the round starts over at record 0: the accumulators hold the last cell of the first round
                    int prev2 = __shfl_up_sync(FULL, lc, 1);
                    if(lane == 0)
                        prev2 = curCell;
                    runPasses(__ballot_sync(FULL, valid && lc != prev2));
                }
            }
            if(FUSED && valid)
            {
#ifdef PICSTEP_RANK_EAGER
                rank[i] = myRank;
#else
                rank[i] = arrives ? (0x80000000u | myRank) : myRank; // bit 31: arrival, ranked behind the stayers
#endif
            }
        }
        if(curCell >= 0)
        {
            flushCell();
            if(FUSED && lane == 0)
                stayCnt[sc * SCVOL + curCell] = stayCarry;
        }
        __syncthreads();
        // ---- combine the warp-private tiles and flush once to global J (red.global.add.f32) --------------------------
        // one thread per (component, z, y) row of the supercell's J tile: which private tiles overlap the row is
        // decided once per row, the TX values of the row are summed in registers (231 rows for TSC: one pass)
        {
            int const ox = scx * SCX + P.g[0] - C::WLO, oy = scy * SCY + P.g[1] - C::WLO, oz = scz * SCZ + P.g[2] - C::WLO;
            constexpr int ROWS = 3 * C::TZ * C::TY;
            for(int row = threadIdx.x; row < ROWS; row += blockDim.x)
            {
                int const cmp = row / (C::TZ * C::TY);
                int const r = row - cmp * (C::TZ * C::TY);
                int const z = r / C::TY, y = r - z * C::TY;
                float v[C::TX];
#pragma unroll
                for(int x = 0; x < C::TX; ++x)
                    v[x] = 0.0f;
#pragma unroll
                for(int zc = 0; zc < SCZ; ++zc)
                {
                    int const tz = z - zc;
#pragma unroll
                    for(int h = 0; h < 2; ++h)
                    {
                        int const ty = y - h * (SCY / 2);
                        if(tz >= 0 && tz < C::PZ && ty >= 0 && ty < C::PY)
                        {
                            float const* __restrict__ src = tiles + (zc * 2 + h) * 3 * C::PVC + cmp * C::PVC + C::PX * (ty + C::PY * tz);
#pragma unroll
                            for(int x = 0; x < C::TX; ++x)
                                v[x] += src[x];
                        }
                    }
                }
                float* const dst = (cmp == 0 ? J.c[0] : (cmp == 1 ? J.c[1] : J.c[2])) + fidx(P, ox, oy + y, oz + z);
#pragma unroll
                for(int x = 0; x < C::TX; ++x)
                    if(v[x] != 0.0f)
                        redGlobal(dst + x, v[x]);
            }
        }
    }

    template<int SHAPE, int PUSHER, bool FUSED, int SOLVER = 0>
    cudaError_t launchRunT(DevParams const& P, SpeciesDev const& S, SpeciesDev const& D, uint32_t const* inv, Field3 E, Field3 B, Field3 J, uint32_t const* cellOff, uint32_t* cellCnt, uint32_t* stayCnt, uint32_t* key, uint32_t* rank, TileMaps const& maps, ScArea const& area, cudaStream_t st)
    {
        int const nscTot = P.nsc[(area.axis + 1) % 3] * P.nsc[(area.axis + 2) % 3] * area.count;
        if(nscTot <= 0)
            return cudaSuccess;
        constexpr size_t smem = runSmemBytes<SHAPE, FUSED>();
        cudaError_t e = cudaFuncSetAttribute(runKernel<SHAPE, PUSHER, FUSED, SOLVER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if(e != cudaSuccess)
            return e;
        runKernel<SHAPE, PUSHER, FUSED, SOLVER><<<nscTot, 256, smem, st>>>(P, S, D, inv, E, B, J, cellOff, cellCnt, stayCnt, key, rank, maps, area);
        return cudaGetLastError();
    }

    bool runKernelSupports(int shape, int solver)
    {
        // Esirkepov: NGP..PQS (PCS: support 5 does not fit the 4-node window); EmZ: CIC..PQS (no EmZ for NGP)
        return (solver == 0 && shape >= 0 && shape <= 3) || (solver == 1 && shape >= 1 && shape <= 3);
    }

    /** stand-alone deposition of one species (picstep_deposit) */
    cudaError_t launchDepositRun(int shape, int solver, DevParams const& P, SpeciesDev const& S, Field3 J, uint32_t const* cellOff, cudaStream_t st)
    {
        Field3 none{};
        TileMaps const noMaps{};
        ScArea const whole{2, 0, 1, P.nsc[2]};
#define PS_CASE(SH, SO)                                                                                               \
    if(shape == SH && solver == SO)                                                                                   \
        return launchRunT<SH, 0, false, SO>(P, S, S, nullptr, none, none, J, cellOff, nullptr, nullptr, nullptr, nullptr, noMaps, whole, st);
        PS_CASE(0, 0)
        PS_CASE(1, 0)
        PS_CASE(2, 0)
        PS_CASE(3, 0)
        PS_CASE(1, 1)
        PS_CASE(2, 1)
        PS_CASE(3, 1)
#undef PS_CASE
        return cudaErrorInvalidValue;
    }

    /** fused gather + push + move + deposit of one species (picstep_step fast path) */
    cudaError_t launchPushDeposit(int shape, int pusher, int solver, DevParams const& P, SpeciesDev const& S, SpeciesDev const& D, uint32_t const* inv, Field3 E, Field3 B, Field3 J, uint32_t const* cellOff, uint32_t* cellCnt, uint32_t* stayCnt, uint32_t* key, uint32_t* rank, TileMaps const& maps, ScArea const& area, cudaStream_t st)
    {
#define PS_CASE(SH, PU)                                                                                               \
    if(shape == SH && pusher == PU && solver == 0)                                                                    \
        return launchRunT<SH, PU, true, 0>(P, S, D, inv, E, B, J, cellOff, cellCnt, stayCnt, key, rank, maps, area, st);   \
    if(shape == SH && pusher == PU && solver == 1 && SH >= 1)                                                         \
        return launchRunT<(SH >= 1 ? SH : 1), PU, true, 1>(P, S, D, inv, E, B, J, cellOff, cellCnt, stayCnt, key, rank, maps, area, st);
        PS_CASE(0, 0)
        PS_CASE(1, 0)
        PS_CASE(2, 0)
        PS_CASE(3, 0)
        PS_CASE(0, 1)
        PS_CASE(1, 1)
        PS_CASE(2, 1)
        PS_CASE(3, 1)
        PS_CASE(0, 2)
        PS_CASE(1, 2)
        PS_CASE(2, 2)
        PS_CASE(3, 2)
#undef PS_CASE
        return cudaErrorInvalidValue;
    }
} // namespace picstep
