// resort.cu — particle re-sort across supercells and rank migration (reference kernels K2-K5:
// KernelShiftParticles / KernelFillGaps / KernelCopyGuardToExchange / KernelInsertParticles / KernelDeleteParticles,
// include/pmacc/particles/ParticlesBase.kernel:61-938).
//
// The reference moves leaving particles between linked frame lists in 27 serial checkerboard passes.  Here the
// frame store is a set of supercell-resident SoA runs described by a prefix sum over cells, so the re-sort is a
// counting sort: the push kernel already histogrammed the destination cells, this file turns the histogram into
// run offsets (two-level exclusive scan) and compacts the particles into the second buffer (ballot / prefix-sum
// style slot claiming).  Particles that leave the rank are compacted into 32-byte records per neighbour.
#include "common.cuh"

namespace picstep
{
    // ---- two level exclusive scan over cell counts --------------------------------------------------------------
    // level 1: one CTA per supercell sums its 256 cell counters
    // (cnt2: second histogram that is added in, e.g. the stayers of the fused push kernel; all zero otherwise)
    __global__ void __launch_bounds__(256) supercellSumKernel(uint32_t const* __restrict__ cnt, uint32_t const* __restrict__ cnt2, uint32_t* __restrict__ scSum)
    {
        __shared__ uint32_t ws[8];
        uint32_t v = cnt[blockIdx.x * SCVOL + threadIdx.x] + cnt2[blockIdx.x * SCVOL + threadIdx.x];
#pragma unroll
        for(int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if((threadIdx.x & 31) == 0)
            ws[threadIdx.x >> 5] = v;
        __syncthreads();
        if(threadIdx.x == 0)
        {
            uint32_t s = 0;
#pragma unroll
            for(int i = 0; i < 8; ++i)
                s += ws[i];
            scSum[blockIdx.x] = s;
        }
    }

    // level 2: exclusive scan of the supercell sums by a single CTA (<= a few 100k entries), total -> scOff[n]
    __global__ void __launch_bounds__(1024) supercellScanKernel(uint32_t const* __restrict__ scSum, uint32_t* __restrict__ scOff, int n, uint32_t* __restrict__ total, uint32_t capacity, int* __restrict__ overflow)
    {
        __shared__ uint32_t ws[32];
        __shared__ uint32_t carry;
        if(threadIdx.x == 0)
            carry = 0;
        __syncthreads();
        for(int base = 0; base < n; base += 1024)
        {
            int const i = base + threadIdx.x;
            uint32_t const v = i < n ? scSum[i] : 0u;
            uint32_t x = v;
#pragma unroll
            for(int o = 1; o < 32; o <<= 1)
            {
                uint32_t const y = __shfl_up_sync(0xffffffffu, x, o);
                if((threadIdx.x & 31) >= o)
                    x += y;
            }
            if((threadIdx.x & 31) == 31)
                ws[threadIdx.x >> 5] = x;
            __syncthreads();
            if(threadIdx.x < 32)
            {
                uint32_t t = ws[threadIdx.x];
#pragma unroll
                for(int o = 1; o < 32; o <<= 1)
                {
                    uint32_t const y = __shfl_up_sync(0xffffffffu, t, o);
                    if(threadIdx.x >= o)
                        t += y;
                }
                ws[threadIdx.x] = t;
            }
            __syncthreads();
            uint32_t const warpBase = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0u;
            uint32_t const incl = carry + warpBase + x;
            if(i < n)
                scOff[i] = incl - v;
            __syncthreads();
            if(threadIdx.x == 1023)
                carry = incl;
            __syncthreads();
        }
        if(threadIdx.x == 0)
        {
            scOff[n] = carry;
            *total = carry;
            if(carry > capacity)
                *overflow = 1;
        }
    }

    // level 3: per supercell exclusive scan of its 256 counters + supercell base
    __global__ void __launch_bounds__(256) cellScanKernel(uint32_t const* __restrict__ cnt, uint32_t const* __restrict__ cnt2, uint32_t const* __restrict__ scOff, uint32_t* __restrict__ cellOff, int nsc)
    {
        __shared__ uint32_t ws[8];
        uint32_t const v = cnt[blockIdx.x * SCVOL + threadIdx.x] + cnt2[blockIdx.x * SCVOL + threadIdx.x];
        uint32_t x = v;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
            uint32_t const y = __shfl_up_sync(0xffffffffu, x, o);
            if((threadIdx.x & 31) >= o)
                x += y;
        }
        if((threadIdx.x & 31) == 31)
            ws[threadIdx.x >> 5] = x;
        __syncthreads();
        uint32_t wb = 0;
        for(int i = 0; i < (threadIdx.x >> 5); ++i)
            wb += ws[i];
        cellOff[blockIdx.x * SCVOL + threadIdx.x] = scOff[blockIdx.x] + wb + x - v;
        if(blockIdx.x == nsc - 1 && threadIdx.x == 255)
            cellOff[(long long) nsc * SCVOL] = scOff[nsc];
    }

    // ---- scatter: old run order -> new run order -----------------------------------------------------------------
    // Slot claiming inside a destination cell: lanes of a warp that target the same cell are ranked with
    // __match_any_sync / popc (warp-ballot prefix) and the leader claims the whole group with one atomic.
    __device__ __forceinline__ uint32_t claimSlot(uint32_t* cnt, uint32_t const* newOff, uint32_t k, bool valid)
    {
        uint32_t const active = __ballot_sync(0xffffffffu, valid);
        uint32_t dst = 0;
        if(valid)
        {
            uint32_t const peers = __match_any_sync(active, k);
            int const leader = __ffs(peers) - 1;
            uint32_t const rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
            uint32_t base = 0;
            if((threadIdx.x & 31) == leader)
                base = atomicSub(&cnt[k], uint32_t(__popc(peers)));
            base = __shfl_sync(peers, base, leader);
            // counters run down to zero, so they are clean for the next step without a memset
            dst = newOff[k] + base - uint32_t(__popc(peers)) + rank;
        }
        return dst;
    }

    __global__ void __launch_bounds__(256) scatterKernel(
        SpeciesDev src,
        SpeciesDev dst,
        uint32_t const* __restrict__ key,
        uint32_t const* __restrict__ nOld,
        uint32_t const* __restrict__ newOff,
        uint32_t* __restrict__ cnt,
        uint32_t capacity)
    {
        uint32_t const n = *nOld;
        for(uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
        {
            uint32_t const i = base + threadIdx.x;
            uint32_t k = KEY_DROP;
            if(i < n)
                k = key[i];
            bool const valid = (i < n) && !(k & KEY_LEAVE);
            uint32_t const d = claimSlot(cnt, newOff, k, valid);
            if(valid && d < capacity) // beyond the capacity: the scan has raised the overflow flag, nothing is written
            {
                dst.pos[0][d] = src.pos[0][i];
                dst.pos[1][d] = src.pos[1][i];
                dst.pos[2][d] = src.pos[2][i];
                dst.mom[0][d] = src.mom[0][i];
                dst.mom[1][d] = src.mom[1][i];
                dst.mom[2][d] = src.mom[2][i];
                dst.w[d] = src.w[i];
                dst.cell[d] = uint16_t(k & (SCVOL - 1));
            }
        }
    }

    // Ranked scatter (after the fused push+deposit kernel): every particle already knows its slot inside the
    // destination cell -- stayers keep their order (rank), arrivals follow them (bit 31: rank among the arrivals) --
    // so the permutation needs no atomics and no inter-thread communication: pure streaming with UNROLL
    // independent particles per thread in flight.
    template<int UNROLL>
    __global__ void __launch_bounds__(256) scatterRankedKernel(
        SpeciesDev src,
        SpeciesDev dst,
        uint32_t const* __restrict__ key,
        uint32_t const* __restrict__ rank,
        uint32_t const* __restrict__ nOld,
        uint32_t const* __restrict__ newOff,
        uint32_t const* __restrict__ stayCnt)
    {
        uint32_t const n = *nOld;
        for(uint32_t base = blockIdx.x * (256u * UNROLL); base < n; base += gridDim.x * (256u * UNROLL))
        {
            uint32_t k[UNROLL], r[UNROLL], d[UNROLL];
            bool ok[UNROLL];
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
            {
                uint32_t const i = base + u * 256u + threadIdx.x;
                k[u] = i < n ? __ldcs(key + i) : KEY_DROP;
                r[u] = i < n ? __ldcs(rank + i) : 0u;
                ok[u] = !(k[u] & KEY_LEAVE);
            }
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
                if(ok[u])
                {
                    d[u] = newOff[k[u]] + (r[u] & 0x7fffffffu);
                    if(r[u] >> 31)
                        d[u] += stayCnt[k[u]];
                }
#pragma unroll
            for(int a = 0; a < 7; ++a)
            {
                float v[UNROLL];
                float const* __restrict__ sp = a < 3 ? src.pos[a] : (a < 6 ? src.mom[a - 3] : src.w);
                float* __restrict__ dp = a < 3 ? dst.pos[a] : (a < 6 ? dst.mom[a - 3] : dst.w);
#pragma unroll
                for(int u = 0; u < UNROLL; ++u)
                    if(ok[u])
                        v[u] = __ldcs(sp + base + u * 256u + threadIdx.x);
#pragma unroll
                for(int u = 0; u < UNROLL; ++u)
                    if(ok[u])
                        dp[d[u]] = v[u];
            }
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
                if(ok[u])
                    dst.cell[d[u]] = uint16_t(k[u] & (SCVOL - 1));
        }
    }

    // ---- lazy re-sort (picstep_step fast path) ------------------------------------------------------------------
    // The fused kernel has written the pushed attributes in ITS processing order (index i) together with key/rank.
    // Instead of moving 28 B per particle into the new run order, only the permutation is materialised:
    // inv[slot in the new order] = i (4 B) and the new localCellIdx (2 B).  The next step's fused kernel reads the
    // attributes through inv (stayers keep their relative order, so the indirect loads stay almost fully coalesced)
    // and again writes them in processing order into the other buffer.  12 B read + 6 B written per particle
    // instead of 36 B + 30 B of the physical scatter.
    template<int UNROLL>
    __global__ void __launch_bounds__(256) invertRankedKernel(
        uint32_t const* __restrict__ key,
        uint32_t const* __restrict__ rank,
        uint32_t const* __restrict__ nOld,
        uint32_t const* __restrict__ newOff,
        uint32_t const* __restrict__ stayCnt,
        uint32_t* __restrict__ inv,
        uint16_t* __restrict__ cellOut,
        uint32_t capacity)
    {
        uint32_t const n = *nOld;
        for(uint32_t base = blockIdx.x * (256u * UNROLL); base < n; base += gridDim.x * (256u * UNROLL))
        {
            uint32_t k[UNROLL], r[UNROLL];
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
            {
                uint32_t const i = base + u * 256u + threadIdx.x;
                k[u] = i < n ? __ldcs(key + i) : KEY_DROP;
                r[u] = i < n ? __ldcs(rank + i) : 0u;
            }
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
                if(!(k[u] & KEY_LEAVE))
                {
                    uint32_t d = newOff[k[u]] + (r[u] & 0x7fffffffu);
                    if(r[u] >> 31)
                        d += stayCnt[k[u]];
                    if(d < capacity) // else: overflow flag already raised by the scan
                    {
                        inv[d] = base + u * 256u + threadIdx.x;
                        cellOut[d] = uint16_t(k[u] & (SCVOL - 1));
                    }
                }
        }
    }

    // lazy mode: received records are appended behind the nOld particles of the attribute buffer; their slots are the
    // last ones of their cell's run (the local particles occupy the front), cnt is cleared afterwards
    __global__ void __launch_bounds__(256) appendRecordsKernel(MigRecord const* __restrict__ rec, uint32_t nRec, uint32_t const* __restrict__ nOld, uint32_t appendOff, uint32_t capacity, SpeciesDev dst, uint32_t const* __restrict__ newOff, uint32_t* __restrict__ cnt, uint32_t* __restrict__ inv, int* __restrict__ overflow)
    {
        uint32_t const first = *nOld + appendOff;
        for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nRec; i += gridDim.x * blockDim.x)
        {
            MigRecord const r = rec[i];
            uint32_t const k = r.key & KEY_MASK;
            uint32_t const slot = newOff[k + 1] - 1u - atomicAdd(&cnt[k], 1u);
            uint32_t const d = first + i;
            if(d >= capacity || slot >= capacity)
            {
                *overflow = 1;
                continue;
            }
            dst.pos[0][d] = r.px;
            dst.pos[1][d] = r.py;
            dst.pos[2][d] = r.pz;
            dst.mom[0][d] = r.ux;
            dst.mom[1][d] = r.uy;
            dst.mom[2][d] = r.uz;
            dst.w[d] = r.w;
            inv[slot] = d;
            dst.cell[slot] = uint16_t(k & (SCVOL - 1));
        }
    }

    // materialise the run order: dst[j] = src[inv[j]] (used when a caller needs the sorted arrays themselves)
    __global__ void __launch_bounds__(256) gatherPermKernel(SpeciesDev src, SpeciesDev dst, uint32_t const* __restrict__ inv, uint32_t const* __restrict__ nNew)
    {
        uint32_t const n = *nNew;
        for(uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        {
            uint32_t const i = inv[j];
#pragma unroll
            for(int d = 0; d < 3; ++d)
            {
                dst.pos[d][j] = src.pos[d][i];
                dst.mom[d][j] = src.mom[d][i];
            }
            dst.w[j] = src.w[i];
            dst.cell[j] = src.cell[j];
        }
    }

    // ranked mode: received records fill their cell's run from the END (the local particles occupy the front);
    // cnt was cleared after the scan and is cleared again afterwards by clearRecordCountsKernel
    __global__ void __launch_bounds__(256) scatterRecordsBackKernel(MigRecord const* __restrict__ rec, uint32_t nRec, SpeciesDev dst, uint32_t const* __restrict__ newOff, uint32_t* __restrict__ cnt)
    {
        for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nRec; i += gridDim.x * blockDim.x)
        {
            MigRecord const r = rec[i];
            uint32_t const k = r.key & KEY_MASK;
            uint32_t const d = newOff[k + 1] - 1u - atomicAdd(&cnt[k], 1u);
            dst.pos[0][d] = r.px;
            dst.pos[1][d] = r.py;
            dst.pos[2][d] = r.pz;
            dst.mom[0][d] = r.ux;
            dst.mom[1][d] = r.uy;
            dst.mom[2][d] = r.uz;
            dst.w[d] = r.w;
            dst.cell[d] = uint16_t(k & (SCVOL - 1));
        }
    }
    __global__ void __launch_bounds__(256) clearRecordCountsKernel(MigRecord const* __restrict__ rec, uint32_t nRec, uint32_t* __restrict__ cnt)
    {
        for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nRec; i += gridDim.x * blockDim.x)
            cnt[rec[i].key & KEY_MASK] = 0u;
    }

    // received migration records -> new runs (KernelInsertParticles, ParticlesBase.kernel:846-938)
    __global__ void __launch_bounds__(256) scatterRecordsKernel(MigRecord const* __restrict__ rec, uint32_t nRec, SpeciesDev dst, uint32_t const* __restrict__ newOff, uint32_t* __restrict__ cnt, uint32_t capacity)
    {
        for(uint32_t base = blockIdx.x * blockDim.x; base < nRec; base += gridDim.x * blockDim.x)
        {
            uint32_t const i = base + threadIdx.x;
            MigRecord r;
            r.key = 0;
            if(i < nRec)
                r = rec[i];
            uint32_t const k = r.key & KEY_MASK;
            bool const valid = i < nRec;
            uint32_t const d = claimSlot(cnt, newOff, k, valid);
            if(valid && d < capacity)
            {
                dst.pos[0][d] = r.px;
                dst.pos[1][d] = r.py;
                dst.pos[2][d] = r.pz;
                dst.mom[0][d] = r.ux;
                dst.mom[1][d] = r.uy;
                dst.mom[2][d] = r.uz;
                dst.w[d] = r.w;
                dst.cell[d] = uint16_t(k & (SCVOL - 1));
            }
        }
    }

    __global__ void __launch_bounds__(256) countRecordsKernel(MigRecord const* __restrict__ rec, uint32_t nRec, uint32_t* __restrict__ cnt)
    {
        for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nRec; i += gridDim.x * blockDim.x)
            atomicAdd(&cnt[rec[i].key & KEY_MASK], 1u);
    }

    // ---- pack leavers (KernelCopyGuardToExchange, ParticlesBase.kernel:707-843) ----------------------------------
    // Only particles of the two supercell layers facing the split axis can leave; the launch covers exactly those
    // runs.  Warp-ballot compaction: one atomic per warp and side reserves the output range.
    __global__ void __launch_bounds__(256) packLeaversKernel(
        DevParams P,
        SpeciesDev S,
        uint32_t const* __restrict__ key,
        uint32_t const* __restrict__ cellOff,
        MigRecord* __restrict__ sendLo,
        MigRecord* __restrict__ sendHi,
        uint32_t* __restrict__ sendCnt, // [0]=lower, [1]=upper
        uint32_t capRec,
        int* __restrict__ overflow)
    {
        // blockIdx.x enumerates the border supercells: first the lower layer, then the upper layer
        int const a = P.split_axis;
        int const a1 = (a + 1) % 3, a2 = (a + 2) % 3;
        int const layer = P.nsc[a1] * P.nsc[a2];
        int b = blockIdx.x;
        int c[3];
        c[a] = (b < layer) ? 0 : P.nsc[a] - 1;
        if(b >= layer)
            b -= layer;
        c[a1] = b % P.nsc[a1];
        c[a2] = b / P.nsc[a1];
        int const sc = c[0] + P.nsc[0] * (c[1] + P.nsc[1] * c[2]);
        uint32_t const p0 = cellOff[sc * SCVOL], p1 = cellOff[(sc + 1) * SCVOL];
        int const lane = threadIdx.x & 31;
        for(uint32_t base = p0; base < p1; base += blockDim.x)
        {
            uint32_t const i = base + threadIdx.x;
            uint32_t k = 0;
            if(i < p1)
                k = key[i];
            bool const leave = (i < p1) && (k != KEY_DROP) && (k & KEY_LEAVE);
            bool const up = leave && (k & KEY_UPPER);
            bool const lo = leave && !up;
            uint32_t const mUp = __ballot_sync(0xffffffffu, up), mLo = __ballot_sync(0xffffffffu, lo);
            uint32_t bUp = 0, bLo = 0;
            if(lane == 0)
            {
                if(mLo)
                    bLo = atomicAdd(&sendCnt[0], uint32_t(__popc(mLo)));
                if(mUp)
                    bUp = atomicAdd(&sendCnt[1], uint32_t(__popc(mUp)));
            }
            bLo = __shfl_sync(0xffffffffu, bLo, 0);
            bUp = __shfl_sync(0xffffffffu, bUp, 0);
            if(leave)
            {
                uint32_t const below = (1u << lane) - 1u;
                uint32_t const slot = up ? bUp + __popc(mUp & below) : bLo + __popc(mLo & below);
                if(slot < capRec)
                {
                    MigRecord r;
                    r.px = S.pos[0][i];
                    r.py = S.pos[1][i];
                    r.pz = S.pos[2][i];
                    r.ux = S.mom[0][i];
                    r.uy = S.mom[1][i];
                    r.uz = S.mom[2][i];
                    r.w = S.w[i];
                    r.key = k & KEY_MASK;
                    (up ? sendHi : sendLo)[slot] = r;
                }
                else
                    *overflow = 1;
            }
        }
    }

    // ---- upload helpers ------------------------------------------------------------------------------------------
    // host cell index (cx + n0*(cy + n1*cz)) -> re-sort key, and histogram
    __global__ void __launch_bounds__(256) keysFromCellsKernel(DevParams P, int32_t const* __restrict__ cellIn, uint32_t n, uint32_t* __restrict__ key, uint32_t* __restrict__ cnt, int* __restrict__ bad)
    {
        for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        {
            int const c = cellIn[i];
            long long const ncell = (long long) P.n[0] * P.n[1] * P.n[2];
            if(c < 0 || c >= ncell)
            {
                *bad = 1;
                key[i] = KEY_DROP;
                continue;
            }
            int const cx = c % P.n[0], cy = (c / P.n[0]) % P.n[1], cz = c / (P.n[0] * P.n[1]);
            int const sc = cx / SCX + P.nsc[0] * (cy / SCY + P.nsc[1] * (cz / SCZ));
            int const lc = cx % SCX + SCX * (cy % SCY + SCY * (cz % SCZ));
            uint32_t const k = uint32_t(sc) * SCVOL + uint32_t(lc);
            key[i] = k;
            atomicAdd(&cnt[k], 1u);
        }
    }

    // frame-run order -> host cell index
    __global__ void __launch_bounds__(256) cellsFromRunsKernel(DevParams P, uint16_t const* __restrict__ lcArr, uint32_t const* __restrict__ cellOff, int32_t* __restrict__ cellOut)
    {
        int const sc = blockIdx.x;
        int const scx = sc % P.nsc[0], scy = (sc / P.nsc[0]) % P.nsc[1], scz = sc / (P.nsc[0] * P.nsc[1]);
        uint32_t const p0 = cellOff[sc * SCVOL], p1 = cellOff[(sc + 1) * SCVOL];
        for(uint32_t i = p0 + threadIdx.x; i < p1; i += blockDim.x)
        {
            int const lc = lcArr[i];
            int const cx = scx * SCX + lc % SCX, cy = scy * SCY + (lc / SCX) % SCY, cz = scz * SCZ + lc / (SCX * SCY);
            cellOut[i] = cx + P.n[0] * (cy + P.n[1] * cz);
        }
    }

    __global__ void __launch_bounds__(256) supercellCountsKernel(uint32_t const* __restrict__ cellOff, long long* __restrict__ out, int nsc)
    {
        int const s = blockIdx.x * blockDim.x + threadIdx.x;
        if(s < nsc)
            out[s] = (long long) cellOff[(s + 1) * SCVOL] - (long long) cellOff[s * SCVOL];
    }

    // ---- launchers -----------------------------------------------------------------------------------------------
    static inline int gridFor(uint32_t n, int block = 256, int maxBlocks = 148 * 16)
    {
        long long b = (n + block - 1) / block;
        if(b < 1)
            b = 1;
        if(b > maxBlocks)
            b = maxBlocks;
        return int(b);
    }

    cudaError_t launchScan(uint32_t const* cnt, uint32_t const* cnt2, uint32_t* scSum, uint32_t* scOff, uint32_t* cellOff, int nsc, uint32_t* total, uint32_t capacity, int* overflow, cudaStream_t st)
    {
        supercellSumKernel<<<nsc, 256, 0, st>>>(cnt, cnt2, scSum);
        supercellScanKernel<<<1, 1024, 0, st>>>(scSum, scOff, nsc, total, capacity, overflow);
        cellScanKernel<<<nsc, 256, 0, st>>>(cnt, cnt2, scOff, cellOff, nsc);
        return cudaGetLastError();
    }

    cudaError_t launchScatter(SpeciesDev src, SpeciesDev dst, uint32_t const* key, uint32_t const* nOld, uint32_t nOldUpper, uint32_t const* newOff, uint32_t* cnt, uint32_t capacity, cudaStream_t st)
    {
        scatterKernel<<<gridFor(nOldUpper), 256, 0, st>>>(src, dst, key, nOld, newOff, cnt, capacity);
        return cudaGetLastError();
    }

    cudaError_t launchScatterRanked(SpeciesDev src, SpeciesDev dst, uint32_t const* key, uint32_t const* rank, uint32_t const* nOld, uint32_t nOldUpper, uint32_t const* newOff, uint32_t const* stayCnt, cudaStream_t st)
    {
        constexpr int UNROLL = 4;
        long long blocks = (nOldUpper + 256ll * UNROLL - 1) / (256ll * UNROLL);
        if(blocks < 1)
            blocks = 1;
        if(blocks > 148 * 32)
            blocks = 148 * 32;
        scatterRankedKernel<UNROLL><<<int(blocks), 256, 0, st>>>(src, dst, key, rank, nOld, newOff, stayCnt);
        return cudaGetLastError();
    }

    cudaError_t launchInvertRanked(uint32_t const* key, uint32_t const* rank, uint32_t const* nOld, uint32_t nOldUpper, uint32_t const* newOff, uint32_t const* stayCnt, uint32_t* inv, uint16_t* cellOut, uint32_t capacity, cudaStream_t st)
    {
        constexpr int UNROLL = 4;
        long long blocks = (nOldUpper + 256ll * UNROLL - 1) / (256ll * UNROLL);
        if(blocks < 1)
            blocks = 1;
        if(blocks > 148 * 32)
            blocks = 148 * 32;
        invertRankedKernel<UNROLL><<<int(blocks), 256, 0, st>>>(key, rank, nOld, newOff, stayCnt, inv, cellOut, capacity);
        return cudaGetLastError();
    }

    cudaError_t launchAppendRecords(MigRecord const* rec, uint32_t nRec, uint32_t const* nOld, uint32_t appendOff, uint32_t capacity, SpeciesDev dst, uint32_t const* newOff, uint32_t* cnt, uint32_t* inv, int* overflow, cudaStream_t st)
    {
        if(nRec == 0)
            return cudaSuccess;
        appendRecordsKernel<<<gridFor(nRec), 256, 0, st>>>(rec, nRec, nOld, appendOff, capacity, dst, newOff, cnt, inv, overflow);
        return cudaGetLastError();
    }

    cudaError_t launchGatherPerm(SpeciesDev src, SpeciesDev dst, uint32_t const* inv, uint32_t const* nNew, uint32_t nUpper, cudaStream_t st)
    {
        gatherPermKernel<<<gridFor(nUpper), 256, 0, st>>>(src, dst, inv, nNew);
        return cudaGetLastError();
    }

    cudaError_t launchScatterRecordsBack(MigRecord const* rec, uint32_t nRec, SpeciesDev dst, uint32_t const* newOff, uint32_t* cnt, cudaStream_t st)
    {
        if(nRec == 0)
            return cudaSuccess;
        scatterRecordsBackKernel<<<gridFor(nRec), 256, 0, st>>>(rec, nRec, dst, newOff, cnt);
        return cudaGetLastError();
    }

    cudaError_t launchClearRecordCounts(MigRecord const* rec, uint32_t nRec, uint32_t* cnt, cudaStream_t st)
    {
        if(nRec == 0)
            return cudaSuccess;
        clearRecordCountsKernel<<<gridFor(nRec), 256, 0, st>>>(rec, nRec, cnt);
        return cudaGetLastError();
    }

    cudaError_t launchCountRecords(MigRecord const* rec, uint32_t nRec, uint32_t* cnt, cudaStream_t st)
    {
        if(nRec == 0)
            return cudaSuccess;
        countRecordsKernel<<<gridFor(nRec), 256, 0, st>>>(rec, nRec, cnt);
        return cudaGetLastError();
    }

    cudaError_t launchScatterRecords(MigRecord const* rec, uint32_t nRec, SpeciesDev dst, uint32_t const* newOff, uint32_t* cnt, uint32_t capacity, cudaStream_t st)
    {
        if(nRec == 0)
            return cudaSuccess;
        scatterRecordsKernel<<<gridFor(nRec), 256, 0, st>>>(rec, nRec, dst, newOff, cnt, capacity);
        return cudaGetLastError();
    }

    cudaError_t launchPackLeavers(DevParams const& P, SpeciesDev S, uint32_t const* key, uint32_t const* cellOff, MigRecord* lo, MigRecord* hi, uint32_t* sendCnt, uint32_t capRec, int* overflow, cudaStream_t st)
    {
        int const a = P.split_axis;
        int const layer = P.nsc[(a + 1) % 3] * P.nsc[(a + 2) % 3];
        packLeaversKernel<<<2 * layer, 256, 0, st>>>(P, S, key, cellOff, lo, hi, sendCnt, capRec, overflow);
        return cudaGetLastError();
    }

    cudaError_t launchKeysFromCells(DevParams const& P, int32_t const* cellIn, uint32_t n, uint32_t* key, uint32_t* cnt, int* bad, cudaStream_t st)
    {
        if(n == 0)
            return cudaSuccess;
        keysFromCellsKernel<<<gridFor(n), 256, 0, st>>>(P, cellIn, n, key, cnt, bad);
        return cudaGetLastError();
    }

    cudaError_t launchCellsFromRuns(DevParams const& P, uint16_t const* lc, uint32_t const* cellOff, int32_t* cellOut, cudaStream_t st)
    {
        cellsFromRunsKernel<<<P.nsc[0] * P.nsc[1] * P.nsc[2], 256, 0, st>>>(P, lc, cellOff, cellOut);
        return cudaGetLastError();
    }

    cudaError_t launchSupercellCounts(uint32_t const* cellOff, long long* out, int nsc, cudaStream_t st)
    {
        supercellCountsKernel<<<(nsc + 255) / 256, 256, 0, st>>>(cellOff, out, nsc);
        return cudaGetLastError();
    }
} // namespace picstep
