// push.cu — gather + push + move (reference kernel K1: KernelMoveAndMarkParticles,
// include/picongpu/particles/Particles.kernel:170-316).
//
// One CTA per supercell.  The six E/B component tiles (supercell + interpolation margins) are staged in shared
// memory as SoA planes; every thread then owns particles of the supercell's frame run (contiguous, cell sorted,
// coalesced SoA loads), interpolates E and B on the Yee-staggered positions, applies Boris/Vay in registers,
// moves the particle and emits its re-sort key.  The per-destination-cell histogram for the re-sort is
// accumulated in shared memory and flushed with one global atomic per touched cell.
#include "common.cuh"
#include "pusher.cuh"
#include "tma.cuh"

#include <cstdio>
#include "shapes.cuh"

namespace picstep
{
    template<int SHAPE, int PUSHER>
    __global__ void __launch_bounds__(256) pushKernel(
        DevParams P,
        SpeciesDev S,
        Field3 E,
        Field3 B,
        uint32_t const* __restrict__ cellOff,
        uint32_t* __restrict__ cellCnt,
        uint32_t* __restrict__ key,
        const __grid_constant__ TileMaps maps)
    {
        using T = Tile<SHAPE>;
        extern __shared__ __align__(128) float tile[]; // [B0,B1,B2 | E0,E1,E2][TZ][TY][PX] followed by the 10x10x6 histogram
        __shared__ uint64_t tileBar;
        uint32_t* hist = reinterpret_cast<uint32_t*>(tile + T::WORDS);
        constexpr int HX = SCX + 2, HY = SCY + 2, HZ = SCZ + 2, HV = HX * HY * HZ;

        int const sc = blockIdx.x;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        uint32_t const p0 = cellOff[sc * SCVOL], p1 = cellOff[(sc + 1) * SCVOL];
        if(p0 == p1)
            return;

        for(int i = threadIdx.x; i < HV; i += blockDim.x)
            hist[i] = 0u;
        // stage the E and B tiles with one TMA box each (x, y, z, 3 components); the histogram is cleared meanwhile
        if(threadIdx.x == 0)
        {
            int const ox = scx * SCX + P.g[0] - T::LO + maps.lead, oy = scy * SCY + P.g[1] - T::LO, oz = scz * SCZ + P.g[2] - T::LO;
            mbarInit(&tileBar, 1);
            mbarExpectTx(&tileBar, 2 * T::BYTES_PER_FIELD);
            tmaLoadTile(tile, &maps.B, ox, oy, oz, &tileBar);
            tmaLoadTile(tile + T::HALF, &maps.E, ox, oy, oz, &tileBar);
        }
        __syncthreads();
        mbarWait(&tileBar, 0);

        float const rc2 = float(1.0 / double(P.c) / double(P.c));
        float const* tB = tile;
        float const* tE = tile + T::HALF;

        for(uint32_t i = p0 + threadIdx.x; i < p1; i += blockDim.x)
        {
            float px = S.pos[0][i], py = S.pos[1][i], pz = S.pos[2][i];
            float u[3] = {S.mom[0][i], S.mom[1][i], S.mom[2][i]};
            float const w = S.w[i];
            int const lc = S.cell[i];
            int const lx = lc % SCX, ly = (lc / SCX) % SCY, lz = lc / (SCX * SCY);

            float Bf[3], Ef[3];
            gatherEB<SHAPE>(tB, tE, lx, ly, lz, px, py, pz, Ef, Bf);
            float const mass = S.mass_per_w * w;
            float const charge = S.charge_per_w * w;
            pushMomentum<PUSHER>(P, rc2, mass, charge, Ef, Bf, u);
            float vx, vy, vz;
            velocityOf(rc2, mass, u[0], u[1], u[2], vx, vy, vz);
            float np[3] = {px + ps_div(vx * P.dt, P.cell[0]), py + ps_div(vy * P.dt, P.cell[1]), pz + ps_div(vz * P.dt, P.cell[2])};

            // moveParticle (MoveParticle.hpp:48-160): wrap to [0,1) with the +-0.5 shift trick, cell crossing
            int dir[3];
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                float q = np[d] - 0.5f;
                float mv = 0.0f;
                if(q < -0.5f)
                    mv = -1.0f;
                if(q >= 0.5f)
                    mv = 1.0f;
                q -= mv;
                np[d] = q + 0.5f;
                dir[d] = int(mv);
            }
            S.pos[0][i] = np[0];
            S.pos[1][i] = np[1];
            S.pos[2][i] = np[2];
            S.mom[0][i] = u[0];
            S.mom[1][i] = u[1];
            S.mom[2][i] = u[2];

            // re-sort key: destination supercell + cell (the reference encodes this as localCellIdx + multiMask and
            // resolves it in KernelShiftParticles, pmacc/particles/ParticlesBase.kernel:361-615)
            int const nl[3] = {lx + dir[0], ly + dir[1], lz + dir[2]};
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
            uint32_t k;
            if(drop)
                k = KEY_DROP;
            else
            {
                int const dsc = gc[0] / SCX + P.nsc[0] * (gc[1] / SCY + P.nsc[1] * (gc[2] / SCZ));
                int const dlc = gc[0] % SCX + SCX * (gc[1] % SCY + SCY * (gc[2] % SCZ));
                k = uint32_t(dsc) * SCVOL + uint32_t(dlc);
                if(flag)
                    k |= flag;
                else
                    atomicAdd(&hist[(nl[0] + 1) + HX * ((nl[1] + 1) + HY * (nl[2] + 1))], 1u);
            }
            key[i] = k;
        }
        __syncthreads();
        // flush the histogram: destination cell of histogram bin (hx,hy,hz) with the same wrap rules
        for(int i = threadIdx.x; i < HV; i += blockDim.x)
        {
            uint32_t const c = hist[i];
            if(c == 0u)
                continue;
            int gc[3] = {scx * SCX + i % HX - 1, scy * SCY + (i / HX) % HY - 1, scz * SCZ + i / (HX * HY) - 1};
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                if(gc[d] < 0)
                    gc[d] += P.n[d];
                else if(gc[d] >= P.n[d])
                    gc[d] -= P.n[d];
            }
            int const dsc = gc[0] / SCX + P.nsc[0] * (gc[1] / SCY + P.nsc[1] * (gc[2] / SCZ));
            int const dlc = gc[0] % SCX + SCX * (gc[1] % SCY + SCY * (gc[2] % SCZ));
            atomicAdd(&cellCnt[dsc * SCVOL + dlc], c);
        }
    }

    template<int SHAPE>
    size_t pushSmemBytes()
    {
        return sizeof(float) * Tile<SHAPE>::WORDS + sizeof(uint32_t) * (SCX + 2) * (SCY + 2) * (SCZ + 2);
    }

    template<int SHAPE, int PUSHER>
    cudaError_t launchPushT(DevParams const& P, SpeciesDev const& S, Field3 E, Field3 B, uint32_t const* cellOff, uint32_t* cellCnt, uint32_t* key, TileMaps const& maps, cudaStream_t st)
    {
        int const nscTot = P.nsc[0] * P.nsc[1] * P.nsc[2];
        size_t const smem = pushSmemBytes<SHAPE>();
        cudaError_t e = cudaFuncSetAttribute(pushKernel<SHAPE, PUSHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if(e != cudaSuccess)
            return e;
        pushKernel<SHAPE, PUSHER><<<nscTot, 256, smem, st>>>(P, S, E, B, cellOff, cellCnt, key, maps);
        return cudaGetLastError();
    }

    cudaError_t launchPush(int shape, int pusher, DevParams const& P, SpeciesDev const& S, Field3 E, Field3 B, uint32_t const* cellOff, uint32_t* cellCnt, uint32_t* key, TileMaps const& maps, cudaStream_t st)
    {
#define PS_CASE(SH, PU)                                                                                               \
    if(shape == SH && pusher == PU)                                                                                   \
        return launchPushT<SH, PU>(P, S, E, B, cellOff, cellCnt, key, maps, st);
        PS_CASE(0, 0)
        PS_CASE(1, 0)
        PS_CASE(2, 0)
        PS_CASE(3, 0)
        PS_CASE(4, 0)
        PS_CASE(0, 1)
        PS_CASE(1, 1)
        PS_CASE(2, 1)
        PS_CASE(3, 1)
        PS_CASE(4, 1)
        PS_CASE(0, 2)
        PS_CASE(1, 2)
        PS_CASE(2, 2)
        PS_CASE(3, 2)
        PS_CASE(4, 2)
#undef PS_CASE
        return cudaErrorInvalidValue;
    }

    /** box of the E/B tile of one supercell for `shape`: {PX, TY, TZ} and the margin below the supercell origin */
    void tileBox(int shape, int box[3], int* lo)
    {
#define PS_CASE(SH)                                                                                                   \
    if(shape == SH)                                                                                                   \
    {                                                                                                                 \
        box[0] = Tile<SH>::PX;                                                                                        \
        box[1] = Tile<SH>::TY;                                                                                        \
        box[2] = Tile<SH>::TZ;                                                                                        \
        *lo = Tile<SH>::LO;                                                                                           \
    }
        PS_CASE(0)
        PS_CASE(1)
        PS_CASE(2)
        PS_CASE(3)
        PS_CASE(4)
#undef PS_CASE
    }

    /** TMA descriptor of one field: 4-D tensor (x, y, z, component) over the SoA planes, box = one supercell tile.
     * cuTensorMapEncodeTiled is taken from the driver through the runtime (no link dependency on libcuda). */
    int makeTileMap(CUtensorMap* map, float* base, int const N[3], long long vol, int const box[3], char* err, size_t errLen)
    {
        using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, cuuint64_t const*, cuuint64_t const*, cuuint32_t const*, cuuint32_t const*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static Encode encode = nullptr;
        if(!encode)
        {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qr;
            if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn)
            {
                snprintf(err, errLen, "cuTensorMapEncodeTiled not available from the driver");
                return 1;
            }
            encode = reinterpret_cast<Encode>(fn);
        }
        cuuint64_t const dims[4] = {cuuint64_t(N[0]), cuuint64_t(N[1]), cuuint64_t(N[2]), 3};
        cuuint64_t const strides[3] = {cuuint64_t(N[0]) * 4, cuuint64_t(N[0]) * N[1] * 4, cuuint64_t(vol) * 4};
        cuuint32_t const bx[4] = {cuuint32_t(box[0]), cuuint32_t(box[1]), cuuint32_t(box[2]), 3};
        cuuint32_t const es[4] = {1, 1, 1, 1};
        CUresult const r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if(r != CUDA_SUCCESS)
        {
            snprintf(err, errLen, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
            return 1;
        }
        return 0;
    }

    // ---- gather only (parity test hook for FieldToParticleInterpolation) --------------------------------------
    template<int SHAPE>
    __global__ void __launch_bounds__(256) gatherKernel(DevParams P, SpeciesDev S, Field3 E, Field3 B, uint32_t const* __restrict__ cellOff, float* __restrict__ out, long long np)
    {
        using T = Tile<SHAPE>;
        extern __shared__ float tile[];
        int const sc = blockIdx.x;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        uint32_t const p0 = cellOff[sc * SCVOL], p1 = cellOff[(sc + 1) * SCVOL];
        if(p0 == p1)
            return;
        int const ox = scx * SCX + P.g[0] - T::LO, oy = scy * SCY + P.g[1] - T::LO, oz = scz * SCZ + P.g[2] - T::LO;
        constexpr int ROWS = T::TY * T::TZ;
        for(int i = threadIdx.x; i < 6 * ROWS * T::TX; i += blockDim.x)
        {
            int const x = i % T::TX;
            int const row = (i / T::TX) % ROWS;
            int const comp = i / (T::TX * ROWS);
            float const* src = comp < 3 ? B.c[comp] : E.c[comp - 3];
            tile[(comp < 3 ? comp * T::TV : T::HALF + (comp - 3) * T::TV) + row * T::PX + x] = __ldg(src + fidx(P, ox + x, oy + row % T::TY, oz + row / T::TY));
        }
        __syncthreads();
        for(uint32_t i = p0 + threadIdx.x; i < p1; i += blockDim.x)
        {
            float const px = S.pos[0][i], py = S.pos[1][i], pz = S.pos[2][i];
            int const lc = S.cell[i];
            int const lx = lc % SCX, ly = (lc / SCX) % SCY, lz = lc / (SCX * SCY);
            float Ef[3], Bf[3];
            gatherEB<SHAPE>(tile, tile + T::HALF, lx, ly, lz, px, py, pz, Ef, Bf);
#pragma unroll
            for(int k = 0; k < 3; ++k)
            {
                out[(long long) k * np + i] = Ef[k];
                out[(long long) (3 + k) * np + i] = Bf[k];
            }
        }
    }

    cudaError_t launchGather(int shape, DevParams const& P, SpeciesDev const& S, Field3 E, Field3 B, uint32_t const* cellOff, float* out, long long np, cudaStream_t st)
    {
        int const nscTot = P.nsc[0] * P.nsc[1] * P.nsc[2];
#define PS_CASE(SH)                                                                                                   \
    if(shape == SH)                                                                                                   \
    {                                                                                                                 \
        size_t const smem = sizeof(float) * Tile<SH>::WORDS;                                                          \
        cudaFuncSetAttribute(gatherKernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));               \
        gatherKernel<SH><<<nscTot, 256, smem, st>>>(P, S, E, B, cellOff, out, np);                                    \
        return cudaGetLastError();                                                                                    \
    }
        PS_CASE(0)
        PS_CASE(1)
        PS_CASE(2)
        PS_CASE(3)
        PS_CASE(4)
#undef PS_CASE
        return cudaErrorInvalidValue;
    }
} // namespace picstep
