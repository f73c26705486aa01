// picstep.cu — C ABI (include/picstep.h) and the C++17 host driver that mirrors Simulation::runOneStep
// (reference: include/picongpu/simulation/control/Simulation.hpp:522-542).  All device work of a context is queued
// on one compute stream; guard/migration traffic between ranks goes through NCCL send/recv (comm.cu).
#include "../../include/picstep.h"
#include "common.cuh"
#include "tma.cuh"

#include <algorithm>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

namespace picstep
{
    // kernels (push.cu, deposit.cu, resort.cu, fields.cu, init.cu)
    cudaError_t launchPush(int, int, DevParams const&, SpeciesDev const&, Field3, Field3, uint32_t const*, uint32_t*, uint32_t*, TileMaps const&, cudaStream_t);
    void tileBox(int, int*, int*);
    int makeTileMap(CUtensorMap*, float*, int const*, long long, int const*, char*, size_t);
    cudaError_t launchGather(int, DevParams const&, SpeciesDev const&, Field3, Field3, uint32_t const*, float*, long long, cudaStream_t);
    cudaError_t launchDeposit(int, int, bool, DevParams const&, SpeciesDev const&, Field3, uint32_t const*, cudaStream_t);
    bool runKernelSupports(int, int);
    cudaError_t launchDepositRun(int, int, DevParams const&, SpeciesDev const&, Field3, uint32_t const*, cudaStream_t);
    cudaError_t launchPushDeposit(int, int, int, DevParams const&, SpeciesDev const&, SpeciesDev const&, uint32_t const*, Field3, Field3, Field3, uint32_t const*, uint32_t*, uint32_t*, uint32_t*, uint32_t*, TileMaps const&, ScArea const&, cudaStream_t);
    cudaError_t launchInvertRanked(uint32_t const*, uint32_t const*, uint32_t const*, uint32_t, uint32_t const*, uint32_t const*, uint32_t*, uint16_t*, uint32_t, cudaStream_t);
    cudaError_t launchAppendRecords(MigRecord const*, uint32_t, uint32_t const*, uint32_t, uint32_t, SpeciesDev, uint32_t const*, uint32_t*, uint32_t*, int*, cudaStream_t);
    cudaError_t launchGatherPerm(SpeciesDev, SpeciesDev, uint32_t const*, uint32_t const*, uint32_t, cudaStream_t);
    cudaError_t launchScan(uint32_t const*, uint32_t const*, uint32_t*, uint32_t*, uint32_t*, int, uint32_t*, uint32_t, int*, cudaStream_t);
    cudaError_t launchScatterRanked(SpeciesDev, SpeciesDev, uint32_t const*, uint32_t const*, uint32_t const*, uint32_t, uint32_t const*, uint32_t const*, cudaStream_t);
    cudaError_t launchScatterRecordsBack(MigRecord const*, uint32_t, SpeciesDev, uint32_t const*, uint32_t*, cudaStream_t);
    cudaError_t launchClearRecordCounts(MigRecord const*, uint32_t, uint32_t*, cudaStream_t);
    cudaError_t launchScatter(SpeciesDev, SpeciesDev, uint32_t const*, uint32_t const*, uint32_t, uint32_t const*, uint32_t*, uint32_t, cudaStream_t);
    cudaError_t launchCountRecords(MigRecord const*, uint32_t, uint32_t*, cudaStream_t);
    cudaError_t launchScatterRecords(MigRecord const*, uint32_t, SpeciesDev, uint32_t const*, uint32_t*, uint32_t, cudaStream_t);
    cudaError_t launchPackLeavers(DevParams const&, SpeciesDev, uint32_t const*, uint32_t const*, MigRecord*, MigRecord*, uint32_t*, uint32_t, int*, cudaStream_t);
    cudaError_t launchKeysFromCells(DevParams const&, int32_t const*, uint32_t, uint32_t*, uint32_t*, int*, cudaStream_t);
    cudaError_t launchCellsFromRuns(DevParams const&, uint16_t const*, uint32_t const*, int32_t*, cudaStream_t);
    cudaError_t launchSupercellCounts(uint32_t const*, long long*, int, cudaStream_t);
    void fdtdBox(int*);
    cudaError_t launchFdtdTma(int, bool, DevParams const&, Field3, Field3, CUtensorMap const&, int, cudaStream_t);
    cudaError_t launchIncident(DevParams const&, Field3, LaserDev const&, cudaStream_t);
    cudaError_t launchPmlUpdateE(DevParams const&, PmlDev const&, Field3, Field3, cudaStream_t);
    cudaError_t launchPmlUpdateBHalf(DevParams const&, PmlDev const&, Field3, Field3, bool, cudaStream_t);
    cudaError_t launchUpdateBHalf(int, DevParams const&, LeheCoeffs const&, Field3, Field3, cudaStream_t);
    cudaError_t launchUpdateE(DevParams const&, LeheCoeffs const&, Field3, Field3, cudaStream_t);
    cudaError_t launchAddCurrent(DevParams const&, Field3, Field3, bool, cudaStream_t);
    cudaError_t launchAbsorb(DevParams const&, Field3, AbsorberDev const&, cudaStream_t);
    cudaError_t launchHaloLocal(bool, DevParams const&, Field3, int, int, int, int, int, cudaStream_t);
    cudaError_t launchHaloPack(DevParams const&, Field3, int, int, int, int, float*, cudaStream_t);
    cudaError_t launchHaloUnpack(bool, DevParams const&, Field3, int, int, int, int, float const*, cudaStream_t);
    cudaError_t launchAosToSoa(float const*, Field3, long long, cudaStream_t);
    cudaError_t launchSoaToAos(Field3, float*, long long, cudaStream_t);
    cudaError_t launchFieldEnergy(DevParams const&, Field3, Field3, double*, cudaStream_t);
    cudaError_t launchParticleEnergy(DevParams const&, SpeciesDev, uint32_t const*, uint32_t const*, double*, cudaStream_t);
    cudaError_t launchChargeDensity(int, DevParams const&, SpeciesDev, uint32_t const*, float*, cudaStream_t);
    cudaError_t launchGaussResidual(DevParams const&, Field3, float const*, int*, cudaStream_t);
    cudaError_t launchKhiInit(DevParams const&, SpeciesDev, SpeciesDev, uint32_t*, uint32_t*, KhiArgs const&, cudaStream_t);
    cudaError_t launchThermalInit(DevParams const&, SpeciesDev, uint32_t*, KhiArgs const&, cudaStream_t);

    // NCCL transport (comm.cu)
    struct Comm;
    int commUniqueId(void* id128, std::string& err);
    int commInit(Comm** out, void const* id128, int rank, int nranks, std::string& err);
    void commDestroy(Comm*);
    // exchange with the lower / upper neighbour in one NCCL group; null pointers or zero counts are skipped
    int commSendRecv(Comm*, void const* sendLo, size_t nSendLo, void* recvLo, size_t nRecvLo, int rankLo, void const* sendHi, size_t nSendHi, void* recvHi, size_t nRecvHi, int rankHi, cudaStream_t, std::string& err);

    struct SpeciesHost
    {
        std::string name;
        float massRatio = 1, chargeRatio = 1;
        // per-species policies: shape<>, particlePusher<>, current<> flags of the species definition
        // (param/speciesAttributes.param:195-256); default = picstep_params
        int shape = 0, pusher = 0, current = 0;
        TileMaps tileMaps{}; // TMA descriptors of the E/B supercell tile of this species' shape
        int64_t capacity = 0;
        float* attr[2][7] = {}; // px,py,pz,ux,uy,uz,w per buffer
        uint16_t* cell[2] = {};
        int cur = 0;
        uint32_t* key = nullptr;
        uint32_t* cellOff[2] = {};
        uint32_t* cellCnt = nullptr; // per destination cell: histogram of the re-sort keys (ranked mode: arrivals only)
        uint32_t* stayCnt = nullptr; // per cell: particles that stay (written by the fused kernel, zero otherwise)
        uint32_t* rank = nullptr; // per particle: slot inside the destination cell (fused kernel)
        uint32_t uploadN = 0; // particle count of the upload in flight (uploadCopies -> uploadSort)
        bool ranked = false; // key/rank/stayCnt come from the fused kernel (which wrote the pushed attributes into buffer cur^1)
        uint32_t* inv = nullptr; // lazy re-sort: attribute index of slot j of the run order
        bool lazy = false; // attr[cur] is addressed through inv; cell[cur], cellOff[cur] are in run order
        std::vector<void*> raw; // cudaMalloc'ed blocks behind the skewed per-particle arrays
        uint32_t *scSum = nullptr, *scOff = nullptr;
        uint32_t* nDev = nullptr; // [2], indexed like cur
        uint32_t nUpper = 0; // host side upper bound of the particle count
        // migration
        MigRecord *sendLo = nullptr, *sendHi = nullptr, *recvLo = nullptr, *recvHi = nullptr;
        uint32_t capRec = 0;
        uint32_t* sendCnt = nullptr; // device [2]
    };

    constexpr int NSTAGE = 7;
} // namespace picstep

using namespace picstep;

struct picstep_ctx
{
    picstep_params prm{};
    DevParams P{};
    LeheCoeffs lehe{};
    TileMaps tileMaps{}; // TMA descriptors of E and B for the supercell tile of this shape
    alignas(64) CUtensorMap fdtdMap[2]; // TMA descriptors of E and B for the brick (+ halo) of the Yee update kernels
    bool fdtdTma = false;
    bool pmlOn = false; // absorber_kind == PML: the field update runs the PML functors (per-cell kernels)
    PmlDev pmlE{}, pmlB{}; // the same parameters with the convolutional fields psiE / psiB
    float* psiMem = nullptr; // 12 * vol floats
    int widthShape = 0; // the widest shape of any species: guard exchange margins follow it
    uint32_t* migPinned = nullptr; // pinned readback of the migration counts of all species (overlapped step)
    cudaEvent_t evBorder = nullptr, evComm = nullptr;
    std::vector<cudaEvent_t> evCore; // per species: its CORE launch is through (the re-sort may start on the second stream)
    // overlap evidence (picstep_overlap_times): per step, device time from "BORDER done" to "exchange done" (second stream)
    // and to "last CORE kernel done" (compute stream); timing events, resolved lazily
    struct OverlapSpan
    {
        cudaEvent_t border, comm, core;
    };
    std::vector<OverlapSpan> overlapSpans;
    double overlapMs[2] = {0.0, 0.0};
    int overlapSteps = 0;
    AbsorberDev absorber{}; // exponential absorber: thickness per face (0 = not absorbing) + attenuation table
    float* dampDev = nullptr;
    bool absorbing = false;
    int slides = 0; // number of picstep_slide calls so far (moving window)
    // picstep_step runs the re-sort / migration of a species on a second stream, next to the fused kernel of the next
    // species and the field update (they are bound by different units: HBM vs. shared memory / issue)
    cudaStream_t side = nullptr;
    // exchange of the decomposed step: highest priority, so that its small kernels (pack, NCCL) are scheduled as soon as a
    // CTA slot frees up although the CORE kernel's grid is still being dispatched (measured without: the exchange only
    // completed together with the CORE kernels)
    cudaStream_t commStream = nullptr;
    cudaEvent_t evFused = nullptr;
    cudaEvent_t evFlags = nullptr; // completion of the asynchronous error-flag readback (peekFlags)
    bool flagsPending = false;
    std::vector<cudaEvent_t> evMig;
    int device = 0;
    cudaStream_t stream = nullptr;
    float* fieldMem[3] = {}; // E,B,J : 3*vol floats each, fieldAlloc + tileMaps.lead floats
    float* fieldAlloc[3] = {}; // the cudaMalloc'ed blocks
    float* rho = nullptr; // vol floats (Gauss check scratch)
    float* aosTmp = nullptr; // 3*vol floats (layout conversion scratch)
    float* haloBuf[4] = {}; // sendLo, sendHi, recvLo, recvHi
    size_t haloBufFloats = 0;
    double* redBuf = nullptr; // device [4]
    int* flags = nullptr; // device [4]: 0 overflow, 1 bad cell index, 2 record overflow, 3 gauss max
    uint32_t* hostPinned = nullptr; // pinned [8] readback
    std::vector<SpeciesHost> species;
    std::string err;
    int64_t launches = 0;
    // multi GPU
    Comm* comm = nullptr;
    int rank = 0, nranks = 1, rankLo = -1, rankHi = -1;
    // stage timing: event pairs are recorded asynchronously and resolved when picstep_stage_times() is called,
    // so enabling it does not add host synchronisation to the step
    struct Span
    {
        int stage;
        cudaEvent_t a, b;
    };
    bool timing = false;
    std::vector<Span> spans;
    std::vector<cudaEvent_t> evPool;
    float stageMs[NSTAGE] = {};
};

static std::string g_createErr;

namespace
{
    int fail(picstep_ctx* c, int code, std::string const& msg)
    {
        if(c)
            c->err = msg;
        else
            g_createErr = msg;
        return code;
    }

#define CU(ctx, call)                                                                                                 \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                      \
        if(e_ != cudaSuccess)                                                                                         \
            return fail(ctx, PICSTEP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while(0)

#define KL(ctx, nk, call)                                                                                             \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                      \
        if(e_ != cudaSuccess)                                                                                         \
            return fail(ctx, PICSTEP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
        (ctx)->launches += (nk);                                                                                      \
    } while(0)

    Field3 fieldOf(picstep_ctx* c, int f)
    {
        Field3 F;
        for(int k = 0; k < 3; ++k)
            F.c[k] = c->fieldMem[f] + (long long) k * c->P.vol;
        return F;
    }

    SpeciesDev devOf(picstep_ctx* c, SpeciesHost const& s, int buf)
    {
        SpeciesDev d;
        for(int k = 0; k < 3; ++k)
        {
            d.pos[k] = s.attr[buf][k];
            d.mom[k] = s.attr[buf][3 + k];
        }
        d.w = s.attr[buf][6];
        d.cell = s.cell[buf];
        d.mass_per_w = c->prm.base_mass * s.massRatio;
        d.charge_per_w = c->prm.base_charge * s.chargeRatio;
        return d;
    }

    int numCells(picstep_ctx* c)
    {
        return c->P.nsc[0] * c->P.nsc[1] * c->P.nsc[2] * SCVOL;
    }

    void freeSpeciesBuffers(SpeciesHost& s)
    {
        for(void* p : s.raw)
            cudaFree(p);
        s.raw.clear();
        for(int b = 0; b < 2; ++b)
        {
            for(int k = 0; k < 7; ++k)
                s.attr[b][k] = nullptr;
            s.cell[b] = nullptr;
        }
        s.key = nullptr;
        s.rank = nullptr;
        s.inv = nullptr;
        s.lazy = false;
        s.ranked = false;
        cudaFree(s.sendLo);
        cudaFree(s.sendHi);
        cudaFree(s.recvLo);
        cudaFree(s.recvHi);
        s.sendLo = s.sendHi = s.recvLo = s.recvHi = nullptr;
    }

    // The push and scatter kernels walk all per-particle arrays of a species in lockstep (same element index in up
    // to 18 streams).  cudaMalloc returns 2 MiB aligned blocks, so without a skew every stream would sit at the same
    // offset inside its page and the streams would camp on the same DRAM channels/banks (measured: the scatter ran
    // at 3.5 or 6.3 TB/s depending on the ping-pong direction).  Each array therefore starts at a different offset.
    template<class T>
    int allocSkewed(picstep_ctx* c, SpeciesHost& s, T** out, int64_t count)
    {
        size_t const skew = (size_t(s.raw.size()) * 37u % 61u + 1u) * 33u * 1024u; // multiples of 33 KiB, < 2 MiB
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, sizeof(T) * size_t(count) + skew);
        if(e != cudaSuccess)
            return fail(c, PICSTEP_ERR_CUDA, std::string("cudaMalloc(particle array): ") + cudaGetErrorString(e));
        s.raw.push_back(p);
        *out = reinterpret_cast<T*>(static_cast<char*>(p) + skew);
        return PICSTEP_OK;
    }

    int allocSpeciesBuffers(picstep_ctx* c, SpeciesHost& s, int64_t capacity)
    {
        freeSpeciesBuffers(s);
        if(capacity < 1024)
            capacity = 1024;
        if(capacity >= (int64_t(1) << 32) - 1)
            return fail(c, PICSTEP_ERR_CAPACITY, "species capacity must be < 2^32 particles per rank");
        s.capacity = capacity;
        for(int b = 0; b < 2; ++b)
        {
            for(int k = 0; k < 7; ++k)
                if(int rc = allocSkewed(c, s, &s.attr[b][k], capacity))
                    return rc;
            if(int rc = allocSkewed(c, s, &s.cell[b], capacity))
                return rc;
        }
        if(int rc = allocSkewed(c, s, &s.key, capacity))
            return rc;
        if(int rc = allocSkewed(c, s, &s.rank, capacity))
            return rc;
        if(int rc = allocSkewed(c, s, &s.inv, capacity))
            return rc;
        if(c->P.split_axis >= 0)
        {
            // exchange capacity: particles of one border supercell layer could at most all leave; reserve a
            // quarter of a layer's share of the capacity, at least 64k records (the reference uses fixed
            // BYTES_EXCHANGE_* sizes with a retry loop, include/picongpu/param/memory.param:82-104)
            int const a = c->P.split_axis;
            int64_t const layerShare = capacity / std::max(1, c->P.nsc[a]);
            s.capRec = uint32_t(std::max<int64_t>(65536, layerShare / 2));
            CU(c, cudaMalloc(&s.sendLo, sizeof(MigRecord) * s.capRec));
            CU(c, cudaMalloc(&s.sendHi, sizeof(MigRecord) * s.capRec));
            CU(c, cudaMalloc(&s.recvLo, sizeof(MigRecord) * s.capRec));
            CU(c, cudaMalloc(&s.recvHi, sizeof(MigRecord) * s.capRec));
        }
        return PICSTEP_OK;
    }

    // Grow the per-particle arrays of a species to newCap slots, keeping their contents (the reference grows its frame
    // heap through mallocMC, ParticlesBox.hpp:97-124; here a rank whose plasma gets denser than its first upload --
    // a density front entering an initially sparse slab -- re-allocates before the re-sort would overflow).
    int growSpeciesBuffers(picstep_ctx* c, SpeciesHost& s, int64_t newCap)
    {
        if(newCap <= s.capacity)
            return PICSTEP_OK;
        if(newCap >= (int64_t(1) << 32) - 1)
            return fail(c, PICSTEP_ERR_CAPACITY, "species capacity must be < 2^32 particles per rank");
        int64_t const oldCap = s.capacity;
        std::vector<void*> oldRaw;
        oldRaw.swap(s.raw);
        auto regrow = [&](auto** arr) -> int
        {
            auto* old = *arr;
            if(int rc = allocSkewed(c, s, arr, newCap))
                return rc;
            if(old && oldCap)
                CU(c, cudaMemcpyAsync(*arr, old, sizeof(**arr) * size_t(oldCap), cudaMemcpyDeviceToDevice, c->stream));
            return PICSTEP_OK;
        };
        int rc = PICSTEP_OK;
        for(int b = 0; b < 2 && !rc; ++b)
        {
            for(int k = 0; k < 7 && !rc; ++k)
                rc = regrow(&s.attr[b][k]);
            if(!rc)
                rc = regrow(&s.cell[b]);
        }
        if(!rc)
            rc = regrow(&s.key);
        if(!rc)
            rc = regrow(&s.rank);
        if(!rc)
            rc = regrow(&s.inv);
        cudaStreamSynchronize(c->stream);
        for(void* p : oldRaw)
            cudaFree(p);
        if(rc)
            return rc;
        s.capacity = newCap;
        return PICSTEP_OK;
    }

    // Lehe coefficients in fp64 -> fp32 (Lehe/Derivative.hpp:94-111); betas in fp32 (:134-137)
    void computeLehe(picstep_params const& p, LeheCoeffs& L)
    {
        for(int dir0 = 0; dir0 < 3; ++dir0)
        {
            int const dir1 = (dir0 + 1) % 3, dir2 = (dir0 + 2) % 3;
            double const stepRatio = double(p.cell_size[dir0] / (p.c * p.dt));
            double const coeff = stepRatio * std::sin(1.5707963267948966 * double(p.c) * double(p.dt) / double(p.cell_size[dir0]));
            L.delta[dir0] = float(0.25 * (1.0 - coeff * coeff));
            double const sr1 = double(p.cell_size[dir0] / p.cell_size[dir1]);
            double const sr2 = double(p.cell_size[dir0] / p.cell_size[dir2]);
            L.alpha[dir0] = float(1.0 - 2.0 * (0.125 * sr1 * sr1) - 2.0 * (0.125 * sr2 * sr2) - 3.0 * double(L.delta[dir0]));
            float const s1 = p.cell_size[dir0] / p.cell_size[dir1], s2 = p.cell_size[dir0] / p.cell_size[dir2];
            L.beta1[dir0] = 0.125f * s1 * s1;
            L.beta2[dir0] = 0.125f * s2 * s2;
        }
    }

    cudaEvent_t takeEvent(picstep_ctx* c)
    {
        if(!c->evPool.empty())
        {
            cudaEvent_t e = c->evPool.back();
            c->evPool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }

    struct StageTimer
    {
        picstep_ctx* c;
        picstep_ctx::Span sp{};
        StageTimer(picstep_ctx* ctx, int st) : c(ctx)
        {
            if(c->timing)
            {
                sp.stage = st;
                sp.a = takeEvent(c);
                sp.b = takeEvent(c);
                cudaEventRecord(sp.a, c->stream);
            }
        }
        ~StageTimer()
        {
            if(c->timing && sp.a)
            {
                cudaEventRecord(sp.b, c->stream);
                c->spans.push_back(sp);
            }
        }
    };

    // ---- guard exchange of one field along all axes ------------------------------------------------------------
    // mode -1: by field (E,B: guard := neighbour border; J: border += neighbour guard), 0: copy, 1: add;
    // width >= 0 overrides the margins of picstep_exchange_widths on both sides.
    // The pass along the split axis has three phases -- pack into the send buffers, NCCL send/recv, unpack (copy or
    // add) -- which exchangeField() queues back to back on the compute stream; the overlapped step (stepImpl) queues
    // the first two for J on the communication stream while the CORE area is still being computed.
    struct AxisExchange
    {
        int a, lo, up;
        bool add;
        size_t nSendLo, nSendHi, nRecvLo, nRecvHi; // floats
    };

    AxisExchange axisExchange(picstep_ctx* c, int f, int a, int mode, int width)
    {
        DevParams const& P = c->P;
        AxisExchange x{};
        int w[2];
        picstep_exchange_widths(c->widthShape, c->prm.field_solver, c->prm.lehe_dir, f, a, w);
        if(width >= 0)
            w[0] = w[1] = width;
        x.a = a;
        x.lo = w[0];
        x.up = w[1];
        x.add = mode < 0 ? (f == PICSTEP_FIELD_J) : (mode == 1);
        long long const plane = (long long) P.N[(a == 0) ? 1 : 0] * P.N[(a == 2) ? 1 : 2] * 3;
        if(!x.add)
        {
            // my lower border (up planes) -> lower neighbour's upper guard; my upper border (lo planes) -> upper neighbour's lower guard
            x.nSendLo = size_t(plane * x.up);
            x.nSendHi = size_t(plane * x.lo);
            x.nRecvLo = size_t(plane * x.lo); // from lower neighbour: its upper border -> my lower guard
            x.nRecvHi = size_t(plane * x.up);
        }
        else
        {
            // my lower guard (lo planes) -> lower neighbour adds to its upper border; my upper guard (up planes) -> upper neighbour
            x.nSendLo = size_t(plane * x.lo);
            x.nSendHi = size_t(plane * x.up);
            x.nRecvLo = size_t(plane * x.up); // lower neighbour's upper guard -> add to my lower border
            x.nRecvHi = size_t(plane * x.lo);
        }
        return x;
    }

    int axisPack(picstep_ctx* c, Field3 F, AxisExchange const& x, cudaStream_t st)
    {
        DevParams const& P = c->P;
        int const g = P.g[x.a], n = P.n[x.a];
        if(!x.add)
        {
            if(c->rankLo >= 0)
                KL(c, 1, launchHaloPack(P, F, 3, x.a, g, x.up, c->haloBuf[0], st));
            if(c->rankHi >= 0)
                KL(c, 1, launchHaloPack(P, F, 3, x.a, g + n - x.lo, x.lo, c->haloBuf[1], st));
        }
        else
        {
            if(c->rankLo >= 0)
                KL(c, 1, launchHaloPack(P, F, 3, x.a, g - x.lo, x.lo, c->haloBuf[0], st));
            if(c->rankHi >= 0)
                KL(c, 1, launchHaloPack(P, F, 3, x.a, g + n, x.up, c->haloBuf[1], st));
        }
        return PICSTEP_OK;
    }

    int axisComm(picstep_ctx* c, AxisExchange const& x, cudaStream_t st)
    {
        if(!c->comm)
            return fail(c, PICSTEP_ERR_COMM, "devices > 1 but picstep_comm_init was not called");
        int const rc = commSendRecv(c->comm, c->haloBuf[0], x.nSendLo * sizeof(float), c->haloBuf[2], x.nRecvLo * sizeof(float), c->rankLo, c->haloBuf[1], x.nSendHi * sizeof(float), c->haloBuf[3], x.nRecvHi * sizeof(float), c->rankHi, st, c->err);
        return rc ? PICSTEP_ERR_COMM : PICSTEP_OK;
    }

    int axisUnpack(picstep_ctx* c, Field3 F, AxisExchange const& x, cudaStream_t st)
    {
        DevParams const& P = c->P;
        int const g = P.g[x.a], n = P.n[x.a];
        if(!x.add)
        {
            if(c->rankLo >= 0)
                KL(c, 1, launchHaloUnpack(false, P, F, 3, x.a, g - x.lo, x.lo, c->haloBuf[2], st));
            if(c->rankHi >= 0)
                KL(c, 1, launchHaloUnpack(false, P, F, 3, x.a, g + n, x.up, c->haloBuf[3], st));
        }
        else
        {
            if(c->rankLo >= 0)
                KL(c, 1, launchHaloUnpack(true, P, F, 3, x.a, g, x.up, c->haloBuf[2], st));
            if(c->rankHi >= 0)
                KL(c, 1, launchHaloUnpack(true, P, F, 3, x.a, g + n - x.lo, x.lo, c->haloBuf[3], st));
        }
        return PICSTEP_OK;
    }

    // skipSplit: the pass along the split axis has been done already (overlapped J exchange)
    int exchangeField(picstep_ctx* c, int f, int mode = -1, int width = -1, bool skipSplit = false)
    {
        DevParams const& P = c->P;
        Field3 F = fieldOf(c, f);
        for(int a = 0; a < 3; ++a)
        {
            AxisExchange const x = axisExchange(c, f, a, mode, width);
            int const lo = x.lo, up = x.up;
            int const g = P.g[a], n = P.n[a];
            if(P.wrap[a])
            {
                if(!x.add)
                {
                    KL(c, 1, launchHaloLocal(false, P, F, 3, a, g + n - lo, g - lo, lo, c->stream)); // lower guard <- upper border
                    KL(c, 1, launchHaloLocal(false, P, F, 3, a, g, g + n, up, c->stream)); // upper guard <- lower border
                }
                else
                {
                    KL(c, 1, launchHaloLocal(true, P, F, 3, a, g - lo, g + n - lo, lo, c->stream)); // upper border += lower guard
                    KL(c, 1, launchHaloLocal(true, P, F, 3, a, g + n, g, up, c->stream)); // lower border += upper guard
                }
            }
            else if(a == P.split_axis && c->nranks > 1 && !skipSplit)
            {
                int rc = axisPack(c, F, x, c->stream);
                if(!rc)
                    rc = axisComm(c, x, c->stream);
                if(!rc)
                    rc = axisUnpack(c, F, x, c->stream);
                if(rc)
                    return rc;
            }
            // else: open boundary, guards stay as they are
        }
        return PICSTEP_OK;
    }

    int checkFlags(picstep_ctx* c)
    {
        CU(c, cudaMemcpyAsync(c->hostPinned, c->flags, sizeof(int) * 3, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        int const f[3] = {int(c->hostPinned[0]), int(c->hostPinned[1]), int(c->hostPinned[2])};
        if(f[0] | f[1] | f[2]) // report once: a later valid upload / step must not see a stale error
            CU(c, cudaMemsetAsync(c->flags, 0, sizeof(int) * 3, c->stream));
        if(f[0])
            return fail(c, PICSTEP_ERR_CAPACITY, "particle capacity exceeded during re-sort");
        if(f[1])
            return fail(c, PICSTEP_ERR_INVALID, "particle cell index out of range");
        if(f[2])
            return fail(c, PICSTEP_ERR_CAPACITY, "migration record buffer overflow");
        return PICSTEP_OK;
    }

    // non-blocking look at the error flags: the copy of the previous call has landed long ago (same stream, and the
    // host has queued n steps since); queues the next copy
    int peekFlags(picstep_ctx* c)
    {
        int rc = PICSTEP_OK;
        if(c->flagsPending && cudaEventQuery(c->evFlags) == cudaSuccess)
        {
            c->flagsPending = false;
            uint32_t const* f = c->hostPinned + 8;
            if(f[0] | f[1] | f[2])
                return checkFlags(c); // synchronises, reports and clears
        }
        if(!c->flagsPending)
        {
            CU(c, cudaMemcpyAsync(c->hostPinned + 8, c->flags, sizeof(int) * 3, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaEventRecord(c->evFlags, c->stream));
            c->flagsPending = true;
        }
        return rc;
    }

    // Everything that depends on where this rank sits in the device grid: neighbour ranks, which faces are outer
    // boundaries, transverse exchange ranges, absorber thickness per face and the attenuation table.  Called by
    // picstep_create and again by picstep_slide (the moving window rotates the rank positions along y).
    static void setupBoundaries(picstep_ctx* c, std::vector<float>& damp)
    {
        picstep_params const* p = &c->prm;
        DevParams& P = c->P;
        int const split = P.split_axis;
        c->rankLo = c->rankHi = -1;
        if(split >= 0) // NCCL ranks keep their identity when the window slides
            picstep_window_neighbors(p->devices[split], p->periodic[split], p->rank_pos[split], c->slides, &c->rankLo, &c->rankHi);
        P.has_lower = c->rankLo >= 0;
        P.has_upper = c->rankHi >= 0;
        damp.assign(size_t(6) * ABS_MAX, 1.0f);
        c->absorbing = false;
        for(int d = 0; d < 3; ++d)
        {
            // does a neighbour (possibly this rank itself through the periodic wrap) exist below / above along d?
            bool const nbLo = P.wrap[d] || (d == split && c->rankLo >= 0);
            bool const nbHi = P.wrap[d] || (d == split && c->rankHi >= 0);
            P.tlo[d] = nbLo ? 0 : P.g[d];
            P.thi[d] = nbHi ? P.N[d] : P.g[d] + P.n[d];
            bool const nb[2] = {nbLo, nbHi};
            for(int sd = 0; sd < 2; ++sd)
            {
                int cells = (p->absorber_kind == PICSTEP_ABSORBER_EXPONENTIAL && !nb[sd]) ? p->absorber_cells[d][sd] : 0;
                if(p->moving_window && d == 1 && sd == 1)
                    cells = 0; // the absorber on the +y side is off while the window slides (Exponential.hpp:97-101)
                c->absorber.cells[d][sd] = cells;
                c->absorbing = c->absorbing || cells > 1;
                // PML: local thickness, zero at faces with a neighbour and at +y while the window slides (Pml.hpp:120-150)
                int pmlCells = (p->absorber_kind == PICSTEP_ABSORBER_PML && !nb[sd]) ? p->absorber_cells[d][sd] : 0;
                if(p->moving_window && d == 1 && sd == 1)
                    pmlCells = 0;
                c->pmlE.thickness[d][sd] = c->pmlB.thickness[d][sd] = pmlCells;
                for(int f = 0; f < cells; ++f) // math::exp(-absorberStrength * float_X(factor)) (Exponential.kernel:107)
                    damp[size_t(2 * d + sd) * ABS_MAX + f] = std::exp(-p->absorber_strength[d][sd] * float(f));
            }
        }
    }

    // lazy -> physical run order: gather the attributes through inv into the other buffer (API calls and the
    // un-fused stage functions work on the sorted arrays themselves)
    int ensureSorted(picstep_ctx* c, SpeciesHost& s)
    {
        if(!s.lazy || s.capacity == 0)
            return PICSTEP_OK;
        int const nxt = s.cur ^ 1;
        int const nscTot = c->P.nsc[0] * c->P.nsc[1] * c->P.nsc[2];
        KL(c, 1, launchGatherPerm(devOf(c, s, s.cur), devOf(c, s, nxt), s.inv, s.nDev + s.cur, uint32_t(s.capacity), c->stream));
        CU(c, cudaMemcpyAsync(s.cellOff[nxt], s.cellOff[s.cur], sizeof(uint32_t) * (size_t(nscTot) * SCVOL + 1), cudaMemcpyDeviceToDevice, c->stream));
        CU(c, cudaMemcpyAsync(s.nDev + nxt, s.nDev + s.cur, sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        s.cur = nxt;
        s.lazy = false;
        return PICSTEP_OK;
    }

    // counting sort of species s from buffer `cur` (keys + histogram ready) into the other buffer
    int resortSpecies(picstep_ctx* c, SpeciesHost& s, uint32_t nRecLo, uint32_t nRecHi)
    {
        int const nscTot = c->P.nsc[0] * c->P.nsc[1] * c->P.nsc[2];
        int const nxt = s.cur ^ 1;
        if(nRecLo)
            KL(c, 1, launchCountRecords(s.recvLo, nRecLo, s.cellCnt, c->stream));
        if(nRecHi)
            KL(c, 1, launchCountRecords(s.recvHi, nRecHi, s.cellCnt, c->stream));
        KL(c, 3, launchScan(s.cellCnt, s.stayCnt, s.scSum, s.scOff, s.cellOff[nxt], nscTot, s.nDev + nxt, uint32_t(s.capacity), c->flags, c->stream));
        if(s.ranked)
        {
            // slots were assigned by the fused kernel, which also wrote the pushed attributes into buffer nxt in its
            // processing order: only the permutation (inv) and the new localCellIdx are materialised (lazy re-sort)
            size_t const cntBytes = sizeof(uint32_t) * size_t(nscTot) * SCVOL;
            KL(c, 1, launchInvertRanked(s.key, s.rank, s.nDev + s.cur, s.nUpper, s.cellOff[nxt], s.stayCnt, s.inv, s.cell[nxt], uint32_t(s.capacity), c->stream));
            CU(c, cudaMemsetAsync(s.cellCnt, 0, cntBytes, c->stream));
            CU(c, cudaMemsetAsync(s.stayCnt, 0, cntBytes, c->stream));
            c->launches += 2;
            if(nRecLo)
                KL(c, 1, launchAppendRecords(s.recvLo, nRecLo, s.nDev + s.cur, 0u, uint32_t(s.capacity), devOf(c, s, nxt), s.cellOff[nxt], s.cellCnt, s.inv, c->flags, c->stream));
            if(nRecHi)
                KL(c, 1, launchAppendRecords(s.recvHi, nRecHi, s.nDev + s.cur, nRecLo, uint32_t(s.capacity), devOf(c, s, nxt), s.cellOff[nxt], s.cellCnt, s.inv, c->flags, c->stream));
            if(nRecLo)
                KL(c, 1, launchClearRecordCounts(s.recvLo, nRecLo, s.cellCnt, c->stream));
            if(nRecHi)
                KL(c, 1, launchClearRecordCounts(s.recvHi, nRecHi, s.cellCnt, c->stream));
            s.lazy = true;
            s.ranked = false;
        }
        else
        {
            KL(c, 1, launchScatter(devOf(c, s, s.cur), devOf(c, s, nxt), s.key, s.nDev + s.cur, s.nUpper, s.cellOff[nxt], s.cellCnt, uint32_t(s.capacity), c->stream));
            if(nRecLo)
                KL(c, 1, launchScatterRecords(s.recvLo, nRecLo, devOf(c, s, nxt), s.cellOff[nxt], s.cellCnt, uint32_t(s.capacity), c->stream));
            if(nRecHi)
                KL(c, 1, launchScatterRecords(s.recvHi, nRecHi, devOf(c, s, nxt), s.cellOff[nxt], s.cellCnt, uint32_t(s.capacity), c->stream));
        }
        s.cur = nxt;
        s.nUpper = uint32_t(std::min<int64_t>(s.capacity, int64_t(s.nUpper) + nRecLo + nRecHi));
        return PICSTEP_OK;
    }
} // namespace

extern "C"
{
    const char* picstep_version(void)
    {
#ifdef PICSTEP_EXACT
        return "picstep 0.1 sm_100a exact(fmad=off)";
#else
        return "picstep 0.1 sm_100a fmad=on";
#endif
    }

    const char* picstep_last_error(const picstep_ctx* ctx)
    {
        return ctx ? ctx->err.c_str() : g_createErr.c_str();
    }

    int picstep_neighbor_ranks(const int32_t* devices, const int32_t* periodic, int32_t rank, int32_t axis, int32_t* lower, int32_t* upper)
    {
        if(!devices || !periodic || axis < 0 || axis > 2)
            return PICSTEP_ERR_INVALID;
        int const total = devices[0] * devices[1] * devices[2];
        if(rank < 0 || rank >= total)
            return PICSTEP_ERR_INVALID;
        int pos[3] = {rank % devices[0], (rank / devices[0]) % devices[1], rank / (devices[0] * devices[1])};
        auto lin = [&](int const* p) { return p[0] + devices[0] * (p[1] + devices[1] * p[2]); };
        int lo[3] = {pos[0], pos[1], pos[2]}, hi[3] = {pos[0], pos[1], pos[2]};
        lo[axis] -= 1;
        hi[axis] += 1;
        int rl = -1, rh = -1;
        if(lo[axis] >= 0)
            rl = lin(lo);
        else if(periodic[axis])
        {
            lo[axis] = devices[axis] - 1;
            rl = lin(lo);
        }
        if(hi[axis] < devices[axis])
            rh = lin(hi);
        else if(periodic[axis])
        {
            hi[axis] = 0;
            rh = lin(hi);
        }
        if(lower)
            *lower = rl;
        if(upper)
            *upper = rh;
        return PICSTEP_OK;
    }

    int picstep_exchange_widths(int32_t shape, int32_t field_solver, int32_t lehe_dir, int32_t field, int32_t axis, int32_t* out2)
    {
        if(shape < 0 || shape > 4 || !out2 || axis < 0 || axis > 2)
            return PICSTEP_ERR_INVALID;
        int const supp = shape + 1;
        if(field == PICSTEP_FIELD_J)
        {
            out2[0] = supp / 2 + 1 - (supp + 1) % 2; // Esirkepov.hpp:42-43
            out2[1] = (supp + 1) / 2 + 1;
        }
        else
        {
            int const glo = supp / 2, gup = (supp + 1) / 2; // FieldToParticleInterpolation.hpp:49-50
            int slo = 1, sup = 1; // Yee curls: backward / forward difference
            if(field_solver == PICSTEP_SOLVER_LEHE)
                sup = (axis == lehe_dir) ? 2 : 1; // Lehe/Derivative.hpp:77-91
            out2[0] = std::max(glo, slo);
            out2[1] = std::max(gup, sup);
        }
        return PICSTEP_OK;
    }

    int picstep_create(const picstep_params* p, picstep_ctx** out)
    {
        if(!p || !out)
            return fail(nullptr, PICSTEP_ERR_INVALID, "null argument");
        *out = nullptr;
        int ndev = 0;
        if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            return fail(nullptr, PICSTEP_ERR_NOGPU, "no CUDA device: libpicstep has no CPU fallback");
        if(p->device < 0 || p->device >= ndev)
            return fail(nullptr, PICSTEP_ERR_INVALID, "bad device ordinal");
        if(p->supercell[0] != SCX || p->supercell[1] != SCY || p->supercell[2] != SCZ)
            return fail(nullptr, PICSTEP_ERR_INVALID, "only SuperCellSize 8x8x4 is compiled in");
        if(p->guard_supercells[0] != 1 || p->guard_supercells[1] != 1 || p->guard_supercells[2] != 1)
            return fail(nullptr, PICSTEP_ERR_INVALID, "only GuardSize 1x1x1 is supported");
        if(p->shape < 0 || p->shape > 4 || p->pusher < 0 || p->pusher > 2 || p->current_solver < 0 || p->current_solver > 1 || p->field_solver < 0 || p->field_solver > 1 || p->lehe_dir < 0 || p->lehe_dir > 2)
            return fail(nullptr, PICSTEP_ERR_INVALID, "unknown shape / pusher / current solver / field solver");
        if(p->current_solver == PICSTEP_CURRENT_EMZ && p->shape == PICSTEP_SHAPE_NGP)
            return fail(nullptr, PICSTEP_ERR_INVALID, "EmZ needs at least CIC");
        if(p->laser_enabled)
        {
            if(p->field_solver != PICSTEP_SOLVER_YEE || p->periodic[1])
                return fail(nullptr, PICSTEP_ERR_INVALID, "the incident field source is built for the Yee solver and a non-periodic y axis");
            if(p->laser_profile < 0 || p->laser_profile > PICSTEP_LASER_EXP_RAMP_WITH_PREPULSE)
                return fail(nullptr, PICSTEP_ERR_INVALID, "unknown incident field profile");
            if(p->laser_profile >= PICSTEP_LASER_WAVEPACKET && (!(p->laser_w0_axis[0] > 0.0f) || !(p->laser_w0_axis[1] > 0.0f)))
                return fail(nullptr, PICSTEP_ERR_INVALID, "bad incident field parameters (W0_AXIS_1, W0_AXIS_2)");
            if(p->laser_profile == PICSTEP_LASER_EXP_RAMP_WITH_PREPULSE)
            {
                // static_assert of ExpRampWithPrepulsesLongitudinalUnitless (ExpRampWithPrepulse.hpp:84-88)
                float const* q = p->laser_profile_params;
                float const endUpramp = q[2] - 0.5f * p->laser_nofocus_constant;
                if(!((q[3] < q[4]) && (q[4] < q[5]) && (q[5] < endUpramp)))
                    return fail(nullptr, PICSTEP_ERR_INVALID, "The times in the parameters TIME_POINT_1/2/3 and the beginning of the plateau should be in ascending order");
            }
            bool anyPos = false;
            for(int d = 0; d < 3; ++d)
                anyPos = anyPos || p->laser_position[d][0] || p->laser_position[d][1];
            if(anyPos && p->laser_position[1][0] != p->laser_offset_ymin)
                return fail(nullptr, PICSTEP_ERR_INVALID, "laser_position[1][0] must equal laser_offset_ymin");
            if(p->laser_profile == PICSTEP_LASER_GAUSSIAN_PULSE)
            {
                if(!(p->laser_w0 > 0.0f) || !(p->laser_wave_length > 0.0f) || p->laser_n_modes < 0 || p->laser_n_modes > LASER_MAX_MODES)
                    return fail(nullptr, PICSTEP_ERR_INVALID, "bad GaussianPulse parameters (W0, WAVE_LENGTH, number of Laguerre modes)");
                // GaussianPulseFunctorIncidentE constructor: "Sum of laguerreModes can not be 0."
                float norm = 0.0f;
                for(int m = 0; m < p->laser_n_modes; ++m)
                    norm += p->laser_modes[m];
                if(p->laser_n_modes > 0 && !(std::abs(norm) > std::numeric_limits<float>::epsilon()))
                    return fail(nullptr, PICSTEP_ERR_INVALID, "Sum of laguerreModes can not be 0.");
            }
            // Solver.hpp:133-159 (checkRequirements): the surface keeps clear of the absorber and of the local domain border
            int const minOffset = p->absorber_kind ? p->absorber_cells[1][0] : 0;
            if(p->laser_offset_ymin < minOffset || p->laser_offset_ymin + 2 > p->grid[1])
                return fail(nullptr, PICSTEP_ERR_INVALID, "incident field POSITION[1][0] is too close to the boundary / outside of the first local domain");
            if(p->laser_polarisation < 0 || p->laser_polarisation > 1 || !(p->laser_omega > 0.0f) || !(p->laser_pulse_duration > 0.0f))
                return fail(nullptr, PICSTEP_ERR_INVALID, "bad incident field parameters");
        }
        if(p->absorber_kind == PICSTEP_ABSORBER_PML && p->field_solver != PICSTEP_SOLVER_YEE)
            return fail(nullptr, PICSTEP_ERR_INVALID, "the PML is built for the Yee solver");
        if(p->current_interpolation < 0 || p->current_interpolation > 1 || p->absorber_kind < 0 || p->absorber_kind > 2)
            return fail(nullptr, PICSTEP_ERR_INVALID, "unknown current interpolation / absorber kind");
        for(int d = 0; d < 3; ++d)
            for(int sd = 0; sd < 2; ++sd) // only faces that can absorb (non-periodic axes) are checked
                if(p->absorber_kind && !p->periodic[d] && (p->absorber_cells[d][sd] < 0 || p->absorber_cells[d][sd] >= ABS_MAX || p->absorber_cells[d][sd] > p->grid[d]))
                    return fail(nullptr, PICSTEP_ERR_INVALID, "absorber thickness must be in [0, min(255, local grid)]");
        int nsplit = 0, split = -1;
        for(int d = 0; d < 3; ++d)
        {
            if(p->grid[d] <= 0 || p->grid[d] % p->supercell[d])
                return fail(nullptr, PICSTEP_ERR_INVALID, "grid must be a positive multiple of the supercell size");
            if(p->devices[d] < 1 || p->rank_pos[d] < 0 || p->rank_pos[d] >= p->devices[d])
                return fail(nullptr, PICSTEP_ERR_INVALID, "bad devices / rank_pos");
            if(p->devices[d] > 1)
            {
                ++nsplit;
                split = d;
                if(p->grid[d] / p->supercell[d] < 2)
                    return fail(nullptr, PICSTEP_ERR_INVALID, "at least 2 supercells per rank along the split axis");
            }
        }
        if(nsplit > 1)
            return fail(nullptr, PICSTEP_ERR_INVALID, "only a 1-D domain decomposition (one axis with devices > 1) is supported");

        auto* c = new picstep_ctx();
        c->prm = *p;
        c->device = p->device;
        DevParams& P = c->P;
        for(int d = 0; d < 3; ++d)
        {
            P.n[d] = p->grid[d];
            P.g[d] = p->supercell[d] * p->guard_supercells[d];
            P.N[d] = P.n[d] + 2 * P.g[d];
            P.nsc[d] = P.n[d] / p->supercell[d];
            P.wrap[d] = (p->periodic[d] && p->devices[d] == 1) ? 1 : 0;
            P.open[d] = p->periodic[d] ? 0 : 1;
            P.cell[d] = p->cell_size[d];
        }
        P.split_axis = split;
        P.vol = (long long) P.N[0] * P.N[1] * P.N[2];
        P.dt = p->dt;
        P.c = p->c;
        P.eps0 = p->eps0;
        P.mue0 = p->mue0;
        P.lehe_dir = p->lehe_dir;
        c->nranks = p->devices[0] * p->devices[1] * p->devices[2];
        c->rank = p->rank_pos[0] + p->devices[0] * (p->rank_pos[1] + p->devices[1] * p->rank_pos[2]);
        if(p->moving_window && (p->periodic[1] || (split >= 0 && split != 1)))
        {
            delete c;
            return fail(nullptr, PICSTEP_ERR_INVALID, "the moving window needs a non-periodic y axis and a decomposition along y only");
        }
        std::vector<float> damp;
        setupBoundaries(c, damp);
        computeLehe(*p, c->lehe);
        c->widthShape = p->shape;
        if((long long) numCells(c) >= (1ll << 30))
        {
            delete c;
            return fail(nullptr, PICSTEP_ERR_INVALID, "more than 2^30 cells per rank");
        }

#define CUC(call)                                                                                                     \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                      \
        if(e_ != cudaSuccess)                                                                                         \
        {                                                                                                             \
            std::string m = std::string(#call) + ": " + cudaGetErrorString(e_);                                       \
            picstep_destroy(c);                                                                                       \
            return fail(nullptr, PICSTEP_ERR_CUDA, m);                                                                \
        }                                                                                                             \
    } while(0)
        CUC(cudaSetDevice(c->device));
        CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUC(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
        {
            int prLow = 0, prHigh = 0;
            CUC(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
            CUC(cudaStreamCreateWithPriority(&c->commStream, cudaStreamNonBlocking, prHigh));
        }
        CUC(cudaEventCreateWithFlags(&c->evFused, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&c->evFlags, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&c->evBorder, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&c->evComm, cudaEventDisableTiming));
        int tbox[3], tlo = 0;
        tileBox(p->shape, tbox, &tlo);
        c->tileMaps.lead = tlo & 3;
        for(int f = 0; f < 3; ++f)
        {
            // The fields start `lead` floats into their allocation so that the x origin of every supercell tile
            // (8k + 8 - margin) is a 16-byte aligned address: TMA box coordinates have to be (coordinate 0 of the
            // descriptor, which is anchored at the allocation, is then a multiple of four floats).
            CUC(cudaMalloc(&c->fieldAlloc[f], sizeof(float) * (3 * P.vol + 4)));
            c->fieldMem[f] = c->fieldAlloc[f] + c->tileMaps.lead;
            CUC(cudaMemsetAsync(c->fieldAlloc[f], 0, sizeof(float) * (3 * P.vol + 4), c->stream));
        }
        {
            char msg[128];
            for(int f = 0; f < 2; ++f)
                if(makeTileMap(f == PICSTEP_FIELD_E ? &c->tileMaps.E : &c->tileMaps.B, c->fieldAlloc[f], P.N, P.vol, tbox, msg, sizeof(msg)))
                {
                    picstep_destroy(c);
                    return fail(nullptr, PICSTEP_ERR_CUDA, msg);
                }
        }
        {
            // Yee update through TMA-staged bricks (fields.cu); the Lehe stencil and a driver that refuses the box
            // (e.g. a grid narrower than the box) keep the one-thread-per-cell kernels
            int fbox[3];
            fdtdBox(fbox);
            char msg[128];
            c->fdtdTma = p->field_solver == PICSTEP_SOLVER_YEE && !(p->flags & 16);
            for(int f = 0; f < 2 && c->fdtdTma; ++f)
                if(makeTileMap(&c->fdtdMap[f], c->fieldAlloc[f], P.N, P.vol, fbox, msg, sizeof(msg)))
                    c->fdtdTma = false;
        }
        if(p->absorber_kind == PICSTEP_ABSORBER_PML)
        {
            for(int d = 0; d < 3; ++d)
                if(c->pmlE.thickness[d][0] + c->pmlE.thickness[d][1] > P.n[d])
                {
                    picstep_destroy(c);
                    return fail(nullptr, PICSTEP_ERR_INVALID, "requested PML size exceeds the local domain");
                }
            c->pmlOn = true;
            c->fdtdTma = false; // the PML functors are per-cell kernels
            CUC(cudaMalloc(&c->psiMem, sizeof(float) * 12 * P.vol));
            CUC(cudaMemsetAsync(c->psiMem, 0, sizeof(float) * 12 * P.vol, c->stream));
            for(PmlDev* m : {&c->pmlE, &c->pmlB})
            {
                for(int d = 0; d < 3; ++d)
                {
                    m->sigmaMax[d] = p->pml_sigma_max[d];
                    m->kappaMax[d] = p->pml_kappa_max[d];
                    m->alphaMax[d] = p->pml_alpha_max[d];
                }
                m->sigmaKappaGradingOrder = p->pml_sigma_kappa_grading_order;
                m->alphaGradingOrder = p->pml_alpha_grading_order;
            }
            c->pmlE.psi = c->psiMem;
            c->pmlB.psi = c->psiMem + 6 * P.vol;
        }
        CUC(cudaMalloc(&c->dampDev, sizeof(float) * damp.size()));
        CUC(cudaMemcpy(c->dampDev, damp.data(), sizeof(float) * damp.size(), cudaMemcpyHostToDevice));
        c->absorber.damp = c->dampDev;
        CUC(cudaMalloc(&c->redBuf, sizeof(double) * 4));
        CUC(cudaMalloc(&c->P.stats, sizeof(unsigned long long) * 4));
        CUC(cudaMemsetAsync(c->P.stats, 0, sizeof(unsigned long long) * 4, c->stream));
        CUC(cudaMalloc(&c->flags, sizeof(int) * 4));
        CUC(cudaMemsetAsync(c->flags, 0, sizeof(int) * 4, c->stream));
        CUC(cudaMallocHost(&c->hostPinned, sizeof(double) * 8)); // 16 words: [0..7] readbacks, [8..10] peekFlags
        if(split >= 0)
        {
            long long const plane = (long long) P.N[(split == 0) ? 1 : 0] * P.N[(split == 2) ? 1 : 2] * 3;
            c->haloBufFloats = size_t(plane) * 4; // widest exchange: PCS current margin 4
            for(int b = 0; b < 4; ++b)
                CUC(cudaMalloc(&c->haloBuf[b], sizeof(float) * c->haloBufFloats));
        }
        CUC(cudaStreamSynchronize(c->stream));
#undef CUC
        *out = c;
        return PICSTEP_OK;
    }

    int picstep_destroy(picstep_ctx* c)
    {
        if(!c)
            return PICSTEP_OK;
        cudaSetDevice(c->device);
        if(c->stream)
            cudaStreamSynchronize(c->stream);
        for(auto& s : c->species)
        {
            freeSpeciesBuffers(s);
            cudaFree(s.cellOff[0]);
            cudaFree(s.cellOff[1]);
            cudaFree(s.cellCnt);
            cudaFree(s.stayCnt);
            cudaFree(s.scSum);
            cudaFree(s.scOff);
            cudaFree(s.nDev);
            cudaFree(s.sendCnt);
        }
        for(int f = 0; f < 3; ++f)
            cudaFree(c->fieldAlloc[f]);
        cudaFree(c->rho);
        cudaFree(c->aosTmp);
        for(int b = 0; b < 4; ++b)
            cudaFree(c->haloBuf[b]);
        cudaFree(c->redBuf);
        cudaFree(c->psiMem);
        cudaFree(c->P.stats);
        cudaFree(c->dampDev);
        cudaFree(c->flags);
        if(c->hostPinned)
            cudaFreeHost(c->hostPinned);
        if(c->migPinned)
            cudaFreeHost(c->migPinned);
        if(c->evBorder)
            cudaEventDestroy(c->evBorder);
        if(c->evComm)
            cudaEventDestroy(c->evComm);
        for(auto& sp : c->spans)
        {
            cudaEventDestroy(sp.a);
            cudaEventDestroy(sp.b);
        }
        for(auto e : c->evPool)
            cudaEventDestroy(e);
        if(c->comm)
            commDestroy(c->comm);
        if(c->stream)
            cudaStreamDestroy(c->stream);
        if(c->side)
            cudaStreamDestroy(c->side);
        if(c->commStream)
            cudaStreamDestroy(c->commStream);
        if(c->evFused)
            cudaEventDestroy(c->evFused);
        if(c->evFlags)
            cudaEventDestroy(c->evFlags);
        for(auto e : c->evMig)
            cudaEventDestroy(e);
        for(auto e : c->evCore)
            cudaEventDestroy(e);
        for(auto& o : c->overlapSpans)
        {
            cudaEventDestroy(o.border);
            cudaEventDestroy(o.comm);
            cudaEventDestroy(o.core);
        }
        delete c;
        return PICSTEP_OK;
    }

    int picstep_species_add(picstep_ctx* c, const char* name, float mass_ratio, float charge_ratio, int64_t capacity, int32_t* species_id)
    {
        if(!c || !name)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        SpeciesHost s;
        s.name = name;
        s.massRatio = mass_ratio;
        s.chargeRatio = charge_ratio;
        s.shape = c->prm.shape;
        s.pusher = c->prm.pusher;
        s.current = c->prm.current_solver;
        s.tileMaps = c->tileMaps;
        int const ncell = numCells(c);
        int const nsc = ncell / SCVOL;
        for(int b = 0; b < 2; ++b)
        {
            CU(c, cudaMalloc(&s.cellOff[b], sizeof(uint32_t) * (size_t(ncell) + 1)));
            CU(c, cudaMemsetAsync(s.cellOff[b], 0, sizeof(uint32_t) * (size_t(ncell) + 1), c->stream));
        }
        CU(c, cudaMalloc(&s.cellCnt, sizeof(uint32_t) * ncell));
        CU(c, cudaMemsetAsync(s.cellCnt, 0, sizeof(uint32_t) * ncell, c->stream));
        CU(c, cudaMalloc(&s.stayCnt, sizeof(uint32_t) * ncell));
        CU(c, cudaMemsetAsync(s.stayCnt, 0, sizeof(uint32_t) * ncell, c->stream));
        CU(c, cudaMalloc(&s.scSum, sizeof(uint32_t) * nsc));
        CU(c, cudaMalloc(&s.scOff, sizeof(uint32_t) * (nsc + 1)));
        CU(c, cudaMalloc(&s.nDev, sizeof(uint32_t) * 2));
        CU(c, cudaMemsetAsync(s.nDev, 0, sizeof(uint32_t) * 2, c->stream));
        CU(c, cudaMalloc(&s.sendCnt, sizeof(uint32_t) * 2));
        CU(c, cudaMemsetAsync(s.sendCnt, 0, sizeof(uint32_t) * 2, c->stream));
        // A rank of a decomposed run always takes part in the count / payload exchange of the migration (its neighbours
        // post their send / recv regardless), also when it starts empty: it gets a minimal buffer instead of none.
        if(capacity == 0 && c->P.split_axis >= 0)
            capacity = 4096;
        if(capacity > 0)
        {
            int const rc = allocSpeciesBuffers(c, s, capacity);
            if(rc)
                return rc;
        }
        c->species.push_back(s);
        if(species_id)
            *species_id = int32_t(c->species.size()) - 1;
        return PICSTEP_OK;
    }

    /* Per-species policies, the `shape<>`, `particlePusher<>` and `current<>` flags of a species definition
     * (param/speciesAttributes.param:195-256, e.g. examples/KelvinHelmholtz/.../speciesDefinition.param:64-70); a
     * negative value keeps the current one.  Shapes can be mixed as long as their lower interpolation margins agree
     * modulo four cells (NGP | CIC, TSC | PQS, PCS): the 16-byte alignment of the TMA tile origins is a property of
     * the field allocation, fixed by picstep_params.shape. */
    int picstep_species_set_policy(picstep_ctx* c, int32_t sp, int32_t shape, int32_t pusher, int32_t current_solver)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()))
            return PICSTEP_ERR_INVALID;
        SpeciesHost& s = c->species[sp];
        int const sh = shape < 0 ? s.shape : shape, pu = pusher < 0 ? s.pusher : pusher, cu = current_solver < 0 ? s.current : current_solver;
        if(sh > 4 || pu > 2 || cu > 1)
            return fail(c, PICSTEP_ERR_INVALID, "unknown shape / pusher / current solver");
        if(cu == PICSTEP_CURRENT_EMZ && sh == PICSTEP_SHAPE_NGP)
            return fail(c, PICSTEP_ERR_INVALID, "EmZ needs at least CIC");
        CU(c, cudaSetDevice(c->device));
        if(sh != s.shape)
        {
            int tbox[3], tlo = 0;
            tileBox(sh, tbox, &tlo);
            if((tlo & 3) != c->tileMaps.lead)
                return fail(c, PICSTEP_ERR_INVALID, "this shape's lower margin does not match the field alignment chosen by picstep_params.shape (mix NGP | CIC, TSC | PQS, PCS)");
            char msg[128];
            s.tileMaps.lead = c->tileMaps.lead;
            for(int f = 0; f < 2; ++f)
                if(makeTileMap(f == PICSTEP_FIELD_E ? &s.tileMaps.E : &s.tileMaps.B, c->fieldAlloc[f], c->P.N, c->P.vol, tbox, msg, sizeof(msg)))
                    return fail(c, PICSTEP_ERR_CUDA, msg);
        }
        s.shape = sh;
        s.pusher = pu;
        s.current = cu;
        c->widthShape = c->prm.shape;
        for(auto const& o : c->species)
            c->widthShape = std::max(c->widthShape, o.shape);
        return PICSTEP_OK;
    }

    // ---- fields ---------------------------------------------------------------------------------------------------
    int picstep_fields_upload_soa(picstep_ctx* c, int32_t f, const float* soa)
    {
        if(!c || f < 0 || f > 2 || !soa)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        CU(c, cudaMemcpyAsync(c->fieldMem[f], soa, sizeof(float) * 3 * c->P.vol, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        return PICSTEP_OK;
    }

    int picstep_fields_download_soa(picstep_ctx* c, int32_t f, float* soa)
    {
        if(!c || f < 0 || f > 2 || !soa)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        CU(c, cudaMemcpyAsync(soa, c->fieldMem[f], sizeof(float) * 3 * c->P.vol, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        return PICSTEP_OK;
    }

    int picstep_fields_upload(picstep_ctx* c, int32_t f, const float* aos)
    {
        if(!c || f < 0 || f > 2 || !aos)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        if(!c->aosTmp)
            CU(c, cudaMalloc(&c->aosTmp, sizeof(float) * 3 * c->P.vol));
        CU(c, cudaMemcpyAsync(c->aosTmp, aos, sizeof(float) * 3 * c->P.vol, cudaMemcpyHostToDevice, c->stream));
        KL(c, 1, launchAosToSoa(c->aosTmp, fieldOf(c, f), c->P.vol, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        return PICSTEP_OK;
    }

    int picstep_fields_download(picstep_ctx* c, int32_t f, float* aos)
    {
        if(!c || f < 0 || f > 2 || !aos)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        if(!c->aosTmp)
            CU(c, cudaMalloc(&c->aosTmp, sizeof(float) * 3 * c->P.vol));
        KL(c, 1, launchSoaToAos(fieldOf(c, f), c->aosTmp, c->P.vol, c->stream));
        CU(c, cudaMemcpyAsync(aos, c->aosTmp, sizeof(float) * 3 * c->P.vol, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        return PICSTEP_OK;
    }

    // ---- particles ------------------------------------------------------------------------------------------------
    // Upload of one species in two parts: the host-to-device copies (on any stream: picstep_step_host sends the next
    // species on the second stream while the previous one is being pushed) and the sort into frame runs.
    static int uploadCopies(picstep_ctx* c, int32_t sp, int64_t n, const float* pos, const float* mom, const float* w, const int32_t* cell, cudaStream_t st)
    {
        SpeciesHost& s = c->species[sp];
        if(n > s.capacity || s.capacity == 0)
        {
            int const rc = allocSpeciesBuffers(c, s, std::max<int64_t>(n + n / 4, 4096));
            if(rc)
                return rc;
        }
        // stage the unsorted input in the inactive buffer, build keys + histogram, then scatter into the active one
        s.lazy = false; // everything is overwritten
        int const stage = s.cur ^ 1;
        int const ncell = numCells(c);
        CU(c, cudaMemsetAsync(s.cellCnt, 0, sizeof(uint32_t) * ncell, st));
        s.uploadN = uint32_t(n);
        CU(c, cudaMemcpyAsync(s.nDev + stage, &s.uploadN, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        for(int k = 0; k < 3; ++k)
        {
            CU(c, cudaMemcpyAsync(s.attr[stage][k], pos + k * n, sizeof(float) * n, cudaMemcpyHostToDevice, st));
            CU(c, cudaMemcpyAsync(s.attr[stage][3 + k], mom + k * n, sizeof(float) * n, cudaMemcpyHostToDevice, st));
        }
        CU(c, cudaMemcpyAsync(s.attr[stage][6], w, sizeof(float) * n, cudaMemcpyHostToDevice, st));
        // the int32 host cell indices travel through the (not yet used) key array of the active buffer's pos.x
        int32_t* cellIn = reinterpret_cast<int32_t*>(s.attr[s.cur][0]);
        CU(c, cudaMemcpyAsync(cellIn, cell, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        return PICSTEP_OK;
    }

    static int uploadSort(picstep_ctx* c, int32_t sp)
    {
        SpeciesHost& s = c->species[sp];
        int const stage = s.cur ^ 1;
        int32_t* cellIn = reinterpret_cast<int32_t*>(s.attr[s.cur][0]);
        s.ranked = false;
        KL(c, 1, launchKeysFromCells(c->P, cellIn, s.uploadN, s.key, s.cellCnt, c->flags + 1, c->stream));
        s.cur = stage; // resortSpecies reads from `cur` and writes to the other one
        s.nUpper = s.uploadN;
        return resortSpecies(c, s, 0, 0);
    }

    static int uploadAsync(picstep_ctx* c, int32_t sp, int64_t n, const float* pos, const float* mom, const float* w, const int32_t* cell)
    {
        int const rc = uploadCopies(c, sp, n, pos, mom, w, cell, c->stream);
        return rc ? rc : uploadSort(c, sp);
    }

    int picstep_particles_upload(picstep_ctx* c, int32_t sp, int64_t n, const float* pos, const float* mom, const float* w, const int32_t* cell)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()) || n < 0 || (n > 0 && (!pos || !mom || !w || !cell)))
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        int rc = uploadAsync(c, sp, n, pos, mom, w, cell);
        if(rc)
            return rc;
        return checkFlags(c);
    }

    int picstep_particles_count(picstep_ctx* c, int32_t sp, int64_t* n)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()) || !n)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        SpeciesHost& s = c->species[sp];
        CU(c, cudaMemcpyAsync(c->hostPinned, s.nDev + s.cur, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        *n = int64_t(c->hostPinned[0]);
        s.nUpper = c->hostPinned[0];
        return PICSTEP_OK;
    }

    int picstep_particles_download(picstep_ctx* c, int32_t sp, int64_t capacity, float* pos, float* mom, float* w, int32_t* cell, int64_t* nOut)
    {
        int64_t n = 0;
        int rc = picstep_particles_count(c, sp, &n);
        if(rc)
            return rc;
        if(nOut)
            *nOut = n;
        if(n > capacity)
            return fail(c, PICSTEP_ERR_CAPACITY, "download buffer too small");
        if(n == 0)
            return PICSTEP_OK;
        SpeciesHost& s = c->species[sp];
        if(int rc2 = ensureSorted(c, s))
            return rc2;
        for(int k = 0; k < 3; ++k)
        {
            if(pos)
                CU(c, cudaMemcpyAsync(pos + k * capacity, s.attr[s.cur][k], sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
            if(mom)
                CU(c, cudaMemcpyAsync(mom + k * capacity, s.attr[s.cur][3 + k], sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
        }
        if(w)
            CU(c, cudaMemcpyAsync(w, s.attr[s.cur][6], sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
        if(cell)
        {
            int32_t* tmp = reinterpret_cast<int32_t*>(s.key); // key array is free between steps
            KL(c, 1, launchCellsFromRuns(c->P, s.cell[s.cur], s.cellOff[s.cur], tmp, c->stream));
            CU(c, cudaMemcpyAsync(cell, tmp, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
        }
        CU(c, cudaStreamSynchronize(c->stream));
        return PICSTEP_OK;
    }

    int picstep_supercell_counts(picstep_ctx* c, int32_t sp, int64_t* counts)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()) || !counts)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        SpeciesHost& s = c->species[sp];
        int const nsc = numCells(c) / SCVOL;
        long long* tmp = nullptr;
        CU(c, cudaMalloc(&tmp, sizeof(long long) * nsc));
        KL(c, 1, launchSupercellCounts(s.cellOff[s.cur], tmp, nsc, c->stream));
        CU(c, cudaMemcpyAsync(counts, tmp, sizeof(long long) * nsc, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFree(tmp);
        return PICSTEP_OK;
    }

    int picstep_init_khi(picstep_ctx* c, const int32_t* ppc_dim, float realPPC, double gammaDrift, double temperature_keV, double ev_pic, uint32_t seed)
    {
        if(!c || !ppc_dim || c->species.size() < 2)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        int const ppc = ppc_dim[0] * ppc_dim[1] * ppc_dim[2];
        int64_t const n = int64_t(numCells(c)) * ppc;
        for(int sp = 0; sp < 2; ++sp)
        {
            SpeciesHost& s = c->species[sp];
            if(n > s.capacity)
            {
                int const rc = allocSpeciesBuffers(c, s, n + n / 4);
                if(rc)
                    return rc;
            }
            s.nUpper = uint32_t(n);
            s.lazy = false;
            s.ranked = false;
            uint32_t const n32 = uint32_t(n);
            CU(c, cudaMemcpyAsync(s.nDev + s.cur, &n32, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        }
        SpeciesHost& e = c->species[0];
        SpeciesHost& i = c->species[1];
        float const weighting = realPPC / float(ppc);
        double const beta = std::sqrt(1.0 - 1.0 / (gammaDrift * gammaDrift));
        float const massE = (c->prm.base_mass * e.massRatio) * weighting;
        float const massI = (c->prm.base_mass * i.massRatio) * weighting;
        float const driftE = float(gammaDrift * beta * double(massE) * double(c->prm.c));
        float const driftI = float(gammaDrift * beta * double(massI) * double(c->prm.c));
        float const energy = float(ev_pic * (temperature_keV * 1.0e3));
        float const stddev = std::sqrt((weighting * energy) * massE);
        KhiArgs A;
        for(int d = 0; d < 3; ++d)
        {
            A.ppc[d] = ppc_dim[d];
            A.globalN[d] = c->prm.grid[d] * c->prm.devices[d];
            A.globalOff[d] = c->prm.grid[d] * c->prm.rank_pos[d];
        }
        A.weighting = weighting;
        A.driftE = driftE;
        A.driftI = driftI;
        A.stddev = stddev;
        A.seed = seed;
        KL(c, 1, launchKhiInit(c->P, devOf(c, e, e.cur), devOf(c, i, i.cur), e.cellOff[e.cur], i.cellOff[i.cur], A, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        return PICSTEP_OK;
    }

    int picstep_init_thermal(picstep_ctx* c, int32_t sp, int32_t ppc, float realPPC, double temperature_keV, double ev_pic, uint32_t seed)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()) || ppc < 1)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        SpeciesHost& s = c->species[sp];
        int64_t const n = int64_t(numCells(c)) * ppc;
        if(n > s.capacity)
        {
            // a relativistic thermal plasma piles up / thins out by a few per cent only: 12 % head room
            int const rc = allocSpeciesBuffers(c, s, n + n / 8);
            if(rc)
                return rc;
        }
        s.nUpper = uint32_t(n);
        s.lazy = false;
        s.ranked = false;
        uint32_t const n32 = uint32_t(n);
        CU(c, cudaMemcpyAsync(s.nDev + s.cur, &n32, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        KhiArgs A{};
        for(int d = 0; d < 3; ++d)
        {
            A.ppc[d] = d == 0 ? ppc : 1;
            A.globalN[d] = c->prm.grid[d] * c->prm.devices[d];
            A.globalOff[d] = c->prm.grid[d] * c->prm.rank_pos[d];
        }
        A.weighting = realPPC / float(ppc);
        float const mass = (c->prm.base_mass * s.massRatio) * A.weighting;
        float const energy = float(ev_pic * (temperature_keV * 1.0e3));
        A.stddev = std::sqrt((A.weighting * energy) * mass); // Temperature.hpp:75-80
        A.seed = seed;
        KL(c, 1, launchThermalInit(c->P, devOf(c, s, s.cur), s.cellOff[s.cur], A, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        return PICSTEP_OK;
    }

    // ---- stages ---------------------------------------------------------------------------------------------------
    int picstep_current_reset(picstep_ctx* c)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        StageTimer t(c, 0);
        CU(c, cudaMemsetAsync(c->fieldMem[PICSTEP_FIELD_J], 0, sizeof(float) * 3 * c->P.vol, c->stream));
        c->launches += 1;
        return PICSTEP_OK;
    }

    int picstep_push(picstep_ctx* c, int32_t sp, uint32_t)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()))
            return PICSTEP_ERR_INVALID;
        StageTimer t(c, 1);
        SpeciesHost& s = c->species[sp];
        if(s.capacity == 0)
            return PICSTEP_OK;
        s.ranked = false;
        if(int rc = ensureSorted(c, s))
            return rc;
        KL(c, 1, launchPush(s.shape, s.pusher, c->P, devOf(c, s, s.cur), fieldOf(c, PICSTEP_FIELD_E), fieldOf(c, PICSTEP_FIELD_B), s.cellOff[s.cur], s.cellCnt, s.key, s.tileMaps, c->stream));
        return PICSTEP_OK;
    }

    int picstep_migrate(picstep_ctx* c, int32_t sp)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()))
            return PICSTEP_ERR_INVALID;
        StageTimer t(c, 2);
        SpeciesHost& s = c->species[sp];
        if(s.capacity == 0)
            return PICSTEP_OK;
        uint32_t nRecLo = 0, nRecHi = 0;
        if(c->P.split_axis >= 0 && c->nranks > 1)
        {
            if(!c->comm)
                return fail(c, PICSTEP_ERR_COMM, "devices > 1 but picstep_comm_init was not called");
            CU(c, cudaMemsetAsync(s.sendCnt, 0, sizeof(uint32_t) * 2, c->stream));
            KL(c, 1, launchPackLeavers(c->P, devOf(c, s, s.ranked ? (s.cur ^ 1) : s.cur), s.key, s.cellOff[s.cur], s.sendLo, s.sendHi, s.sendCnt, s.capRec, c->flags + 2, c->stream));
            // counts first (fixed-size header), so receive buffers can never overflow silently
            CU(c, cudaMemcpyAsync(c->hostPinned, s.sendCnt, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, c->stream));
            uint32_t* cntDev = s.sendCnt; // reuse: [0],[1] send counts; receive counts land in scSum[0..1] scratch
            int rc = commSendRecv(c->comm, cntDev, sizeof(uint32_t), s.scSum, sizeof(uint32_t), c->rankLo, cntDev + 1, sizeof(uint32_t), s.scSum + 1, sizeof(uint32_t), c->rankHi, c->stream, c->err);
            if(rc)
                return PICSTEP_ERR_COMM;
            CU(c, cudaMemcpyAsync(c->hostPinned + 2, s.scSum, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaMemcpyAsync(c->hostPinned + 4, s.nDev + s.cur, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            uint32_t nSendLo = c->rankLo >= 0 ? c->hostPinned[0] : 0u, nSendHi = c->rankHi >= 0 ? c->hostPinned[1] : 0u;
            nRecLo = c->rankLo >= 0 ? c->hostPinned[2] : 0u;
            nRecHi = c->rankHi >= 0 ? c->hostPinned[3] : 0u;
            // A buffer overflow on either side must not leave the neighbour blocked in its payload exchange: the
            // exchange is still posted, with the counts clamped on BOTH sides the same way, and reported afterwards.
            bool const overflow = nSendLo > s.capRec || nSendHi > s.capRec || nRecLo > s.capRec || nRecHi > s.capRec;
            nSendLo = std::min(nSendLo, s.capRec);
            nSendHi = std::min(nSendHi, s.capRec);
            nRecLo = std::min(nRecLo, s.capRec);
            nRecHi = std::min(nRecHi, s.capRec);
            rc = commSendRecv(c->comm, s.sendLo, sizeof(MigRecord) * nSendLo, s.recvLo, sizeof(MigRecord) * nRecLo, c->rankLo, s.sendHi, sizeof(MigRecord) * nSendHi, s.recvHi, sizeof(MigRecord) * nRecHi, c->rankHi, c->stream, c->err);
            if(rc)
                return PICSTEP_ERR_COMM;
            if(overflow)
                return fail(c, PICSTEP_ERR_CAPACITY, "migration record buffer overflow");
            // the exact particle count is on the host now: it bounds the launches of the re-sort (instead of an
            // ever growing upper bound) and tells whether the arrivals still fit
            uint32_t const nNow = c->hostPinned[4];
            s.nUpper = nNow;
            if(int64_t(nNow) + nRecLo + nRecHi > s.capacity)
            {
                int64_t const want = int64_t(nNow) + nRecLo + nRecHi;
                if(int grc = growSpeciesBuffers(c, s, want + want / 4 + 4096))
                    return grc;
            }
        }
        return resortSpecies(c, s, nRecLo, nRecHi);
    }

    int picstep_field_exchange(picstep_ctx* c, int32_t f)
    {
        if(!c || f < 0 || f > 2)
            return PICSTEP_ERR_INVALID;
        return exchangeField(c, f);
    }

    /* incidentField::Solver::updateE (updatedIsE) / ::updateBHalf at the fractional step `currentStep`
     * (Solver.hpp:547-575; call sites FDTDBase.hpp:108-117,161-166).  PlaneWave on YMin only; nothing to do on a rank that
     * does not hold the updated plane, or once the moving window has slid. */
    static int incidentUpdate(picstep_ctx* c, bool updatedIsE, float currentStep)
    {
        picstep_params const& p = c->prm;
        if(!p.laser_enabled || c->slides > 0)
            return PICSTEP_OK;
        DevParams const& P = c->P;
        // Solver.hpp:230-236: E sits in the total-field region, B (scattered) one plane closer to the boundary
        int const planeTotal = p.laser_offset_ymin + 1 - (updatedIsE ? 0 : 1);
        int const yl = planeTotal - p.grid[1] * p.rank_pos[1];
        if(yl < 0 || yl >= P.n[1])
            return PICSTEP_OK;
        // POSITION (all zero: {offset, -offset} on every axis), global size, transversal extent of the surface
        int position[3][2], globalSize[3];
        bool anyPos = false;
        for(int d = 0; d < 3; ++d)
            anyPos = anyPos || p.laser_position[d][0] || p.laser_position[d][1];
        for(int d = 0; d < 3; ++d)
        {
            position[d][0] = anyPos ? p.laser_position[d][0] : p.laser_offset_ymin;
            position[d][1] = anyPos ? p.laser_position[d][1] : -p.laser_offset_ymin;
            globalSize[d] = p.grid[d] * p.devices[d];
        }
        LaserDev L{};
        for(int t = 0; t < 2; ++t)
        {
            // Solver.hpp:209-258: contiguous over a periodic transversal axis for the PlaneWave profile only
            int const d = 2 * t;
            int begin = position[d][0] + 1;
            int end = position[d][1] > 0 ? position[d][1] : globalSize[d] + position[d][1];
            L.lastDomain[t] = 1; // x and z are never split: this rank is the last one along both
            if(p.laser_profile == PICSTEP_LASER_PLANE_WAVE && p.periodic[d])
            {
                begin = 0;
                end = globalSize[d];
                L.lastDomain[t] = 0;
            }
            L.lo[t] = std::max(begin, 0);
            L.hi[t] = std::min(end, P.n[d]);
        }
        if(L.lo[0] >= L.hi[0] || L.lo[1] >= L.hi[1])
            return PICSTEP_OK;
        L.polarisation = p.laser_polarisation;
        L.profile = p.laser_profile;
        L.plane = yl + P.g[1];
        L.planeTotal = float(planeTotal);
        L.amplitude = p.laser_amplitude;
        L.omega = p.laser_omega;
        L.pulseDuration = p.laser_pulse_duration;
        L.nofocusConstant = p.laser_nofocus_constant;
        L.rampInit = p.laser_ramp_init;
        L.phase = p.laser_phase;
        L.timeDelay = p.laser_time_delay;
        for(int d = 0; d < 3; ++d)
            L.pol[d] = p.laser_pol_dir[d];
        {
            // BaseFunctorE::getFocus / getOrigin / getAxis2 (Functors.hpp:267-380) for DIR = (0, 1, 0), in float_X
            float const direction[3] = {0.0f, 1.0f, 0.0f};
            for(int d = 0; d < 3; ++d)
            {
                L.focus[d] = p.laser_focus_position[d];
                if(p.laser_focus_origin_center[d])
                    L.focus[d] += float(unsigned(globalSize[d]) / 2u) * P.cell[d];
            }
            float originP = -std::numeric_limits<float>::infinity();
            for(int axis = 0; axis < 3; ++axis)
                if(std::abs(direction[axis]) > std::numeric_limits<float>::epsilon())
                {
                    float const minPosition = (float(position[axis][0]) + 0.75f) * P.cell[axis];
                    int const maxPositionIdx = position[axis][1] > 0 ? position[axis][1] : globalSize[axis] + position[axis][1];
                    float const maxPosition = (float(maxPositionIdx) - 0.75f) * P.cell[axis];
                    float const axisP = std::min((minPosition - L.focus[axis]) / direction[axis], (maxPosition - L.focus[axis]) / direction[axis]);
                    originP = std::max(originP, axisP);
                }
            for(int d = 0; d < 3; ++d)
                L.origin[d] = L.focus[d] + originP * direction[d];
            L.axis2[0] = direction[1] * L.pol[2] - direction[2] * L.pol[1];
            L.axis2[1] = direction[2] * L.pol[0] - direction[0] * L.pol[2];
            L.axis2[2] = direction[0] * L.pol[1] - direction[1] * L.pol[0];
        }
        L.w0Axis[0] = p.laser_w0_axis[0];
        L.w0Axis[1] = p.laser_w0_axis[1];
        for(int k = 0; k < 16; ++k)
            L.prm[k] = p.laser_profile_params[k];
        if(p.laser_profile == PICSTEP_LASER_GAUSSIAN_PULSE)
        {
            // GaussianPulseUnitless (profiles/GaussianPulse.hpp:93-110)
            L.w0 = p.laser_w0;
            L.waveLength = p.laser_wave_length;
            L.rayleighLength = 3.14159265358979323846f * L.w0 * L.w0 / L.waveLength;
            L.timeShift = p.laser_time_shift;
            L.tilted = (p.laser_tilt[0] != 0.0f || p.laser_tilt[1] != 0.0f) ? 1 : 0;
            L.tanTilt[0] = std::tan(p.laser_tilt[0]);
            L.tanTilt[1] = std::tan(p.laser_tilt[1]);
            L.nModes = p.laser_n_modes > 0 ? p.laser_n_modes : 1;
            for(int m = 0; m < L.nModes; ++m)
            {
                L.modes[m] = p.laser_n_modes > 0 ? p.laser_modes[m] : 1.0f;
                L.modePhases[m] = p.laser_n_modes > 0 ? p.laser_mode_phases[m] : 0.0f;
            }
        }
        {
            // Yee dispersion relation along y (calculatePhaseVelocity.hpp, DispersionRelationSolver), fp64
            double const w = double(p.laser_omega), dt = double(P.dt), cc = double(P.c), dy = double(P.cell[1]);
            double const k = 2.0 / dy * std::asin(dy * std::sin(0.5 * w * dt) / (cc * dt));
            L.phaseVelocity = float(w / k / cc);
        }
        L.currentTimeOrigin = currentStep * P.dt;
        float const c2 = P.c * P.c;
        float const curlCoefficient = updatedIsE ? P.dt * c2 : -(0.5f * P.dt);
        L.baseCoefficient = curlCoefficient / P.cell[1] * 1.0f;
        L.updatedIsE = updatedIsE ? 1 : 0;
        KL(c, 1, launchIncident(P, fieldOf(c, updatedIsE ? PICSTEP_FIELD_E : PICSTEP_FIELD_B), L, c->stream));
        return PICSTEP_OK;
    }

    // B -= curl E * dt/2 (updateBFirstHalf / updateBSecondHalf)
    static int updateBHalf(picstep_ctx* c, bool firstHalf)
    {
        Field3 E = fieldOf(c, PICSTEP_FIELD_E), B = fieldOf(c, PICSTEP_FIELD_B);
        if(c->pmlOn) // psiB is advanced once per step, in the first half update (FDTDBase.hpp:200-211)
            KL(c, 1, launchPmlUpdateBHalf(c->P, c->pmlB, E, B, firstHalf, c->stream));
        else if(c->fdtdTma)
            KL(c, 1, launchFdtdTma(0, false, c->P, B, B, c->fdtdMap[PICSTEP_FIELD_E], c->tileMaps.lead, c->stream));
        else
            KL(c, 1, launchUpdateBHalf(c->prm.field_solver, c->P, c->lehe, E, B, c->stream));
        return PICSTEP_OK;
    }

    // update_beforeCurrent; addJ: the current term E += coeff * J rides in the E update kernel (J already reduced)
    static int fieldUpdateBeforeCurrent(picstep_ctx* c, bool addJ, uint32_t step)
    {
        StageTimer t(c, 3);
        Field3 E = fieldOf(c, PICSTEP_FIELD_E), B = fieldOf(c, PICSTEP_FIELD_B);
        int rc = updateBHalf(c, false); // updateBSecondHalf
        if(!rc)
            rc = incidentUpdate(c, false, float(step)); // B by half a step with E_inc at t = step
        if(rc)
            return rc;
        rc = exchangeField(c, PICSTEP_FIELD_B);
        if(!rc)
            rc = incidentUpdate(c, true, float(step) + 0.5f); // E with B_inc at t = step + 1/2, before the E update
        if(rc)
            return rc;
        if(c->pmlOn)
            KL(c, 1, launchPmlUpdateE(c->P, c->pmlE, E, B, c->stream));
        else if(c->fdtdTma)
            KL(c, 1, launchFdtdTma(1, addJ, c->P, E, fieldOf(c, PICSTEP_FIELD_J), c->fdtdMap[PICSTEP_FIELD_B], c->tileMaps.lead, c->stream));
        else
            KL(c, 1, launchUpdateE(c->P, c->lehe, E, B, c->stream));
        return PICSTEP_OK;
    }

    int picstep_field_update_before_current(picstep_ctx* c, uint32_t step)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        return fieldUpdateBeforeCurrent(c, false, step);
    }

    int picstep_deposit(picstep_ctx* c, int32_t sp)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()))
            return PICSTEP_ERR_INVALID;
        StageTimer t(c, 4);
        SpeciesHost& s = c->species[sp];
        if(s.capacity == 0)
            return PICSTEP_OK;
        if(int rc = ensureSorted(c, s))
            return rc;
        if(runKernelSupports(s.shape, s.current) && !(c->prm.flags & 3))
            KL(c, 1, launchDepositRun(s.shape, s.current, c->P, devOf(c, s, s.cur), fieldOf(c, PICSTEP_FIELD_J), s.cellOff[s.cur], c->stream));
        else
            KL(c, 1, launchDeposit(s.shape, s.current, (c->prm.flags & 1) != 0, c->P, devOf(c, s, s.cur), fieldOf(c, PICSTEP_FIELD_J), s.cellOff[s.cur], c->stream));
        return PICSTEP_OK;
    }

    /* fused ParticlePush + CurrentDeposition of one species (fast path of picstep_step): the current of the move
     * is deposited by the kernel that performs the move, from the cell the particle started in */
    static int pushDepositFused(picstep_ctx* c, int32_t sp, ScArea const& area)
    {
        StageTimer t(c, 1);
        SpeciesHost& s = c->species[sp];
        if(s.capacity == 0)
            return PICSTEP_OK;
        KL(c, 1, launchPushDeposit(s.shape, s.pusher, s.current, c->P, devOf(c, s, s.cur), devOf(c, s, s.cur ^ 1), s.lazy ? s.inv : nullptr, fieldOf(c, PICSTEP_FIELD_E), fieldOf(c, PICSTEP_FIELD_B), fieldOf(c, PICSTEP_FIELD_J), s.cellOff[s.cur], s.cellCnt, s.stayCnt, s.key, s.rank, s.tileMaps, area, c->stream));
        s.ranked = true;
        return PICSTEP_OK;
    }

    /* Overlapped step, communication-stream part (Simulation.hpp:537 `__setTransactionEvent(commEvent)` joins the
     * reference's asynchronous particle / field communication the same way): the BORDER area of every species has
     * been pushed, so the particles that leave the rank (only border supercells can lose any) are packed and exchanged
     * and the J guard strips along the split axis (only border-supercell particles deposit there) are sent while the
     * compute stream works on the CORE area.  The host waits ONCE, on the communication stream only, for the record
     * counts of all species -- the CORE kernels are queued by then, so the device does not idle. */
    static int exchangeBorderOverlapped(picstep_ctx* c, cudaStream_t st, std::vector<uint32_t>& nRecLo, std::vector<uint32_t>& nRecHi, AxisExchange const& xj)
    {
        int const ns = int(c->species.size());
        if(!c->comm)
            return fail(c, PICSTEP_ERR_COMM, "devices > 1 but picstep_comm_init was not called");
        if(!c->migPinned)
            CU(c, cudaMallocHost(&c->migPinned, sizeof(uint32_t) * 8 * 64));
        if(ns > 64)
            return fail(c, PICSTEP_ERR_INVALID, "more than 64 species");
        for(int i = 0; i < ns; ++i)
        {
            SpeciesHost& s = c->species[i];
            if(s.capacity == 0)
                continue;
            CU(c, cudaMemsetAsync(s.sendCnt, 0, sizeof(uint32_t) * 2, st));
            KL(c, 1, launchPackLeavers(c->P, devOf(c, s, s.cur ^ 1), s.key, s.cellOff[s.cur], s.sendLo, s.sendHi, s.sendCnt, s.capRec, c->flags + 2, st));
            if(commSendRecv(c->comm, s.sendCnt, sizeof(uint32_t), s.scSum, sizeof(uint32_t), c->rankLo, s.sendCnt + 1, sizeof(uint32_t), s.scSum + 1, sizeof(uint32_t), c->rankHi, st, c->err))
                return PICSTEP_ERR_COMM;
            CU(c, cudaMemcpyAsync(c->migPinned + 8 * i, s.sendCnt, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, st));
            CU(c, cudaMemcpyAsync(c->migPinned + 8 * i + 2, s.scSum, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, st));
            CU(c, cudaMemcpyAsync(c->migPinned + 8 * i + 4, s.nDev + s.cur, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        // J guard strips: pack and send before the host wait, they do not depend on the counts
        int rc = axisPack(c, fieldOf(c, PICSTEP_FIELD_J), xj, st);
        if(!rc)
            rc = axisComm(c, xj, st);
        if(rc)
            return rc;
        CU(c, cudaStreamSynchronize(st));
        bool overflow = false;
        for(int i = 0; i < ns; ++i)
        {
            SpeciesHost& s = c->species[i];
            if(s.capacity == 0)
                continue;
            uint32_t const* h = c->migPinned + 8 * i;
            uint32_t nSendLo = c->rankLo >= 0 ? h[0] : 0u, nSendHi = c->rankHi >= 0 ? h[1] : 0u;
            nRecLo[i] = c->rankLo >= 0 ? h[2] : 0u;
            nRecHi[i] = c->rankHi >= 0 ? h[3] : 0u;
            overflow = overflow || nSendLo > s.capRec || nSendHi > s.capRec || nRecLo[i] > s.capRec || nRecHi[i] > s.capRec;
            nSendLo = std::min(nSendLo, s.capRec);
            nSendHi = std::min(nSendHi, s.capRec);
            nRecLo[i] = std::min(nRecLo[i], s.capRec);
            nRecHi[i] = std::min(nRecHi[i], s.capRec);
            if(commSendRecv(c->comm, s.sendLo, sizeof(MigRecord) * nSendLo, s.recvLo, sizeof(MigRecord) * nRecLo[i], c->rankLo, s.sendHi, sizeof(MigRecord) * nSendHi, s.recvHi, sizeof(MigRecord) * nRecHi[i], c->rankHi, st, c->err))
                return PICSTEP_ERR_COMM;
            s.nUpper = h[4];
        }
        if(overflow)
            return fail(c, PICSTEP_ERR_CAPACITY, "migration record buffer overflow");
        return PICSTEP_OK;
    }

    // skipSplit: the J guard reduction along the split axis has been done already (overlapped step)
    static int addCurrentImpl(picstep_ctx* c, bool skipSplit)
    {
        StageTimer t(c, 5);
        int rc = exchangeField(c, PICSTEP_FIELD_J, -1, -1, skipSplit);
        if(rc)
            return rc;
        bool const binomial = c->prm.current_interpolation == PICSTEP_CURRENT_INTERPOLATION_BINOMIAL;
        if(binomial) // "receive" exchange of FieldJ: one guard cell := neighbour border (FieldJ.x.cpp:118-141, 156-174)
            if((rc = exchangeField(c, PICSTEP_FIELD_J, 0, 1)))
                return rc;
        KL(c, 1, launchAddCurrent(c->P, fieldOf(c, PICSTEP_FIELD_E), fieldOf(c, PICSTEP_FIELD_J), binomial, c->stream));
        return PICSTEP_OK;
    }

    int picstep_add_current(picstep_ctx* c)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        return addCurrentImpl(c, false);
    }

    int picstep_field_update_after_current(picstep_ctx* c, uint32_t step)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        StageTimer t(c, 6);
        Field3 E = fieldOf(c, PICSTEP_FIELD_E), B = fieldOf(c, PICSTEP_FIELD_B);
        if(c->absorbing) // exponentialImpl.run(E) (FDTDBase.hpp:153-158)
            KL(c, 1, launchAbsorb(c->P, E, c->absorber, c->stream));
        int rc = incidentUpdate(c, false, float(step) + 1.0f); // B by half a step with E_inc at t = step + 1 (FDTDBase.hpp:161-166)
        if(!rc)
            rc = exchangeField(c, PICSTEP_FIELD_E);
        if(rc)
            return rc;
        rc = updateBHalf(c, true); // updateBFirstHalf
        if(rc)
            return rc;
        if(c->absorbing) // exponentialImpl.run(B) (FDTDBase.hpp:175-179)
            KL(c, 1, launchAbsorb(c->P, B, c->absorber, c->stream));
        return exchangeField(c, PICSTEP_FIELD_B);
    }

    /* GridController::slide (pmacc/mappings/simulation/GridController.hpp:166-176, CommunicatorMPI.cpp:156-170) +
     * Simulation::slide (Simulation.hpp:581-593): every rank moves one position down in y, the lowest one becomes the
     * top of the window and starts empty (resetAll); ParticleInit of the new slab is the caller's upload. */
    int picstep_slide(picstep_ctx* c, int32_t* was_reset)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        if(!c->prm.moving_window)
            return fail(c, PICSTEP_ERR_INVALID, "picstep_slide needs picstep_params.moving_window");
        CU(c, cudaSetDevice(c->device));
        int const n = c->prm.devices[1];
        c->slides += 1;
        c->prm.rank_pos[1] = (c->prm.rank_pos[1] - 1 + n) % n;
        bool const reset = c->prm.rank_pos[1] == n - 1;
        std::vector<float> damp;
        setupBoundaries(c, damp);
        CU(c, cudaMemcpyAsync(c->dampDev, damp.data(), sizeof(float) * damp.size(), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream)); // damp is a local
        if(reset)
        {
            for(int f = 0; f < 3; ++f)
                CU(c, cudaMemsetAsync(c->fieldAlloc[f], 0, sizeof(float) * (3 * c->P.vol + 4), c->stream));
            if(c->psiMem)
                CU(c, cudaMemsetAsync(c->psiMem, 0, sizeof(float) * 12 * c->P.vol, c->stream));
            size_t const ncell = size_t(numCells(c));
            for(auto& s : c->species)
            {
                if(s.capacity == 0)
                    continue;
                CU(c, cudaMemsetAsync(s.nDev, 0, sizeof(uint32_t) * 2, c->stream));
                CU(c, cudaMemsetAsync(s.cellOff[s.cur], 0, sizeof(uint32_t) * (ncell + 1), c->stream));
                CU(c, cudaMemsetAsync(s.cellCnt, 0, sizeof(uint32_t) * ncell, c->stream));
                CU(c, cudaMemsetAsync(s.stayCnt, 0, sizeof(uint32_t) * ncell, c->stream));
                s.lazy = false;
                s.ranked = false;
                s.nUpper = 0;
            }
        }
        if(was_reset)
            *was_reset = reset ? 1 : 0;
        return PICSTEP_OK;
    }

    /* MovingWindow::getCurrentSlideInfo (simulation/control/MovingWindow.hpp:44-170), a pure function of the step:
     * does the window slide while `step` is computed, and the window offset inside the first GPU after the step */
    int picstep_moving_window_info(int32_t global_cells, int32_t local_cells, double cell_size, double c_dt, double move_point, uint32_t step, int32_t* do_slide, int32_t* offset_first_gpu)
    {
        if(global_cells <= 0 || local_cells <= 0 || global_cells % local_cells || global_cells < 2 * local_cells || !(cell_size > 0) || !(c_dt > 0))
            return PICSTEP_ERR_INVALID;
        if(do_slide)
            *do_slide = 0;
        if(offset_first_gpu)
            *offset_first_gpu = 0;
        uint32_t const windowSize = uint32_t(global_cells - local_cells);
        uint32_t const startCell = uint32_t(std::ceil(double(windowSize) * (1.0 - move_point)));
        uint32_t const firstSlideStep = uint32_t(std::ceil(double(uint32_t(global_cells) - startCell) * cell_size / c_dt) - 1);
        double const wayToFirstMove = double(windowSize - startCell) * cell_size;
        int32_t const firstMoveStep = int32_t(std::ceil(wayToFirstMove / c_dt) - 1);
        if(firstMoveStep <= int32_t(step))
        {
            uint32_t const passed = uint32_t(std::floor(c_dt * double(step) / cell_size));
            uint32_t const pos = passed + startCell;
            uint32_t const nextPassed = uint32_t(std::floor(c_dt * double(step + 1) / cell_size));
            uint32_t const nextPos = nextPassed + startCell;
            bool const endOfInitialGlobalDomain = firstSlideStep <= step;
            bool const passesBorder = (nextPos % uint32_t(local_cells)) < (pos % uint32_t(local_cells));
            if(endOfInitialGlobalDomain && passesBorder && do_slide)
                *do_slide = 1;
            if(offset_first_gpu)
                *offset_first_gpu = int32_t(nextPos % uint32_t(local_cells));
        }
        return PICSTEP_OK;
    }

    // hook(s): called before species s is pushed in the FIRST step (picstep_step_host finishes the upload of s there)
    static int stepImpl(picstep_ctx* c, uint32_t first, uint32_t n, std::function<int(int)> const& hook)
    {
        CU(c, cudaSetDevice(c->device));
        int const ns = int(c->species.size());
        // fast path: the deposition does not depend on the field update, so it is fused into the push kernel
        // Exception: the Binomial filter next to an open face reads J guard cells that no exchange overwrites (edge /
        // corner cells between an open and an exchanged axis).  They hold the un-folded deposits, which depend on the
        // cell a particle is deposited from, so the reference's order (move, wrap, then deposit) has to be kept there.
        bool staleGuardsRead = false;
        if(c->prm.current_interpolation == PICSTEP_CURRENT_INTERPOLATION_BINOMIAL)
        {
            bool anyOpen = false, anyExchange = false;
            for(int d = 0; d < 3; ++d)
            {
                anyOpen = anyOpen || c->P.tlo[d] != 0 || c->P.thi[d] != c->P.N[d];
                anyExchange = anyExchange || c->P.wrap[d] || (d == c->P.split_axis && c->nranks > 1);
            }
            staleGuardsRead = anyOpen && anyExchange;
        }
        bool allRun = true; // every species' (shape, current solver) pair has a fused kernel instantiation
        for(auto const& sp : c->species)
            allRun = allRun && runKernelSupports(sp.shape, sp.current);
        bool const fused = allRun && !(c->prm.flags & 7) && !staleGuardsRead;
        // Measured on B200 (KHI 256^3): 54.63 -> 54.41 ms/step only.  The fused kernel fills every SM (2 CTAs x 115 KB shared
        // memory, 60 K registers), so the re-sort kernels time-share instead of co-running; with a high-priority second
        // stream the re-sort finishes early but the step gets 0.5 ms longer.  Kept for one rank (no NCCL calls from two
        // streams), default priority.
        bool const overlap = fused && !(c->prm.flags & 8) && c->nranks == 1;
        // Several ranks: the BORDER area (the two supercell layers that face the neighbours) is pushed first; its
        // leaving particles and the J guard strips travel on the second stream while the CORE area is computed
        // (pmacc/type/Area.hpp:36-41; FDTDBase.hpp:107-120 and Simulation.hpp:537 overlap the same way).
        bool const commOverlap = fused && !(c->prm.flags & 8) && c->nranks > 1 && c->P.split_axis >= 0;
        std::vector<uint32_t> recLo(size_t(ns), 0u), recHi(size_t(ns), 0u);
        while(int(c->evMig.size()) < ns)
        {
            cudaEvent_t e;
            CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->evMig.push_back(e);
            CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->evCore.push_back(e);
        }
        std::vector<char> pendingMig(size_t(ns), 0);
        int rc = PICSTEP_OK;
        for(uint32_t it = 0; it < n; ++it)
        {
            uint32_t const step = first + it;
            rc = picstep_current_reset(c);
            bool jSplitDone = false;
            if(commOverlap && !rc)
            {
                int const a = c->P.split_axis;
                ScArea const border{a, 0, c->P.nsc[a] - 1, 2}, core{a, 1, 1, c->P.nsc[a] - 2};
                for(int s = 0; s < ns; ++s) // the re-sort of the previous step produced this species' run table
                    if(pendingMig[s])
                    {
                        CU(c, cudaStreamWaitEvent(c->stream, c->evMig[s], 0));
                        pendingMig[s] = false;
                    }
                for(int s = 0; s < ns && !rc; ++s)
                {
                    if(it == 0)
                        rc = hook(s);
                    if(!rc)
                        rc = pushDepositFused(c, s, border);
                }
                if(rc)
                    break;
                CU(c, cudaEventRecord(c->evBorder, c->stream));
                CU(c, cudaStreamWaitEvent(c->commStream, c->evBorder, 0));
                picstep_ctx::OverlapSpan span{};
                if(c->timing)
                {
                    span.border = takeEvent(c);
                    span.comm = takeEvent(c);
                    span.core = takeEvent(c);
                    cudaEventRecord(span.border, c->stream);
                }
                for(int s = 0; s < ns && !rc; ++s)
                {
                    rc = pushDepositFused(c, s, core);
                    CU(c, cudaEventRecord(c->evCore[s], c->stream));
                }
                if(rc)
                    break;
                if(c->timing)
                    cudaEventRecord(span.core, c->stream);
                AxisExchange const xj = axisExchange(c, PICSTEP_FIELD_J, a, -1, -1);
                rc = exchangeBorderOverlapped(c, c->commStream, recLo, recHi, xj);
                if(rc)
                    break;
                CU(c, cudaEventRecord(c->evComm, c->commStream));
                if(c->timing)
                {
                    cudaEventRecord(span.comm, c->commStream);
                    c->overlapSpans.push_back(span);
                }
                CU(c, cudaStreamWaitEvent(c->stream, c->evComm, 0));
                CU(c, cudaStreamWaitEvent(c->side, c->evComm, 0)); // the re-sort appends the received records
                {
                    // The re-sort of a species runs on the second stream as soon as its CORE launch is through: next to the
                    // CORE kernel of the following species (bound by shared memory, the re-sort by HBM) and to the field
                    // update.  The next step's kernels of this species wait for evMig.
                    cudaStream_t const mainStream = c->stream;
                    for(int s = 0; s < ns && !rc; ++s)
                    {
                        SpeciesHost& sp = c->species[s];
                        if(sp.capacity == 0)
                            continue;
                        CU(c, cudaStreamWaitEvent(c->side, c->evCore[s], 0));
                        c->stream = c->side;
                        {
                            StageTimer t(c, 2);
                            if(int64_t(sp.nUpper) + recLo[s] + recHi[s] > sp.capacity)
                            {
                                int64_t const want = int64_t(sp.nUpper) + recLo[s] + recHi[s];
                                rc = growSpeciesBuffers(c, sp, want + want / 4 + 4096);
                            }
                            if(!rc)
                                rc = resortSpecies(c, sp, recLo[s], recHi[s]);
                        }
                        c->stream = mainStream;
                        if(!rc)
                        {
                            CU(c, cudaEventRecord(c->evMig[s], c->side));
                            pendingMig[s] = true;
                        }
                    }
                }
                if(!rc)
                {
                    StageTimer t(c, 5);
                    rc = axisUnpack(c, fieldOf(c, PICSTEP_FIELD_J), xj, c->stream); // border += the neighbours' guard strips
                }
                jSplitDone = true;
            }
            for(int s = 0; s < ns && !rc && !commOverlap; ++s)
            {
                if(it == 0 && (rc = hook(s)))
                    break;
                if(!fused)
                {
                    rc = picstep_push(c, s, step);
                    if(!rc)
                        rc = picstep_migrate(c, s);
                    continue;
                }
                if(overlap && pendingMig[s]) // the re-sort of the previous step produced this species' run table
                {
                    CU(c, cudaStreamWaitEvent(c->stream, c->evMig[s], 0));
                    pendingMig[s] = false;
                }
                rc = pushDepositFused(c, s, ScArea{2, 0, 1, c->P.nsc[2]});
                if(rc)
                    break;
                if(overlap)
                {
                    CU(c, cudaEventRecord(c->evFused, c->stream));
                    CU(c, cudaStreamWaitEvent(c->side, c->evFused, 0));
                    cudaStream_t const mainStream = c->stream;
                    c->stream = c->side; // every launch of the re-sort goes to the second stream
                    rc = picstep_migrate(c, s);
                    c->stream = mainStream;
                    CU(c, cudaEventRecord(c->evMig[s], c->side));
                    pendingMig[s] = true;
                }
                else
                    rc = picstep_migrate(c, s);
            }
            // With the deposition fused into the push J is complete before the field update starts: its guard reduction
            // is done first and the current term rides in the E update kernel (FDTD.hpp:84-85: E += coeff * J after
            // E += curl B c^2 dt, the same two roundings in the same order) instead of a second pass over E.
            bool const fuseJ = fused && c->fdtdTma && c->prm.current_interpolation == PICSTEP_CURRENT_INTERPOLATION_NONE;
            if(!rc && fuseJ)
            {
                StageTimer t(c, 5);
                rc = exchangeField(c, PICSTEP_FIELD_J, -1, -1, jSplitDone);
            }
            if(!rc)
                rc = fieldUpdateBeforeCurrent(c, fuseJ, step);
            for(int s = 0; s < ns && !rc && !fused; ++s)
                rc = picstep_deposit(c, s);
            if(!rc && !fuseJ)
                rc = addCurrentImpl(c, jSplitDone);
            if(!rc)
                rc = picstep_field_update_after_current(c, step);
            if(rc)
                break;
        }
        for(int s = 0; s < ns; ++s) // join: everything that follows on the main stream sees the re-sorted species
            if(pendingMig[s])
                cudaStreamWaitEvent(c->stream, c->evMig[s], 0);
        return rc;
    }

    int picstep_window_neighbors(int32_t n, int32_t periodic, int32_t pos, int32_t slides, int32_t* lower, int32_t* upper)
    {
        if(n < 1 || pos < 0 || pos >= n || slides < 0 || !lower || !upper)
            return PICSTEP_ERR_INVALID;
        int const k = slides % n;
        *lower = (pos > 0 || periodic) ? ((pos - 1 + n) % n + k) % n : -1;
        *upper = (pos < n - 1 || periodic) ? ((pos + 1) % n + k) % n : -1;
        return PICSTEP_OK;
    }

    int picstep_step(picstep_ctx* c, uint32_t first, uint32_t n)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        int const rc = stepImpl(c, first, n, [](int) { return int(PICSTEP_OK); });
        // the device-side error flags (capacity, record overflow) are read without waiting for the steps: a flag that
        // is already up is reported here, anything later by the next call or picstep_sync
        return rc ? rc : peekFlags(c);
    }

    int picstep_step_host(picstep_ctx* c, uint32_t step, float* E, float* B, int32_t nSpecies, const int64_t* n, const float* const* pos, const float* const* mom, const float* const* w, const int32_t* const* cell, double* energies4)
    {
        if(!c || !E || !B || nSpecies != int(c->species.size()))
            return PICSTEP_ERR_INVALID;
        if(nSpecies < 1 || !n || !pos || !mom || !w || !cell)
            return fail(c, PICSTEP_ERR_INVALID, "picstep_step_host needs at least one species and its host arrays");
        for(int s = 0; s < nSpecies; ++s)
            if(n[s] < 0 || (n[s] > 0 && (!pos[s] || !mom[s] || !w[s] || !cell[s])))
                return fail(c, PICSTEP_ERR_INVALID, "picstep_step_host: null particle array");
        CU(c, cudaSetDevice(c->device));
        size_t const fbytes = sizeof(float) * 3 * c->P.vol;
        CU(c, cudaMemcpyAsync(c->fieldMem[PICSTEP_FIELD_E], E, fbytes, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->fieldMem[PICSTEP_FIELD_B], B, fbytes, cudaMemcpyHostToDevice, c->stream));
        // Species 0 is uploaded and sorted on the main stream; the copies of every further species go to the second
        // stream, so that they travel while the species before them is pushed, and are sorted right before their push.
        std::vector<cudaEvent_t> copied(size_t(nSpecies), nullptr);
        int rc = uploadCopies(c, 0, n[0], pos[0], mom[0], w[0], cell[0], c->stream);
        if(!rc)
        {
            // the copies of the next species start when these are through (both streams share the PCIe link; started
            // together they would each get half of it and the first push could not begin any earlier)
            CU(c, cudaEventRecord(c->evFused, c->stream));
            CU(c, cudaStreamWaitEvent(c->side, c->evFused, 0));
            rc = uploadSort(c, 0);
        }
        for(int s = 1; s < nSpecies && !rc; ++s)
        {
            rc = uploadCopies(c, s, n[s], pos[s], mom[s], w[s], cell[s], c->side);
            if(!rc && cudaEventCreateWithFlags(&copied[s], cudaEventDisableTiming) == cudaSuccess)
                cudaEventRecord(copied[s], c->side);
        }
        if(!rc)
            rc = stepImpl(c, step, 1, [&](int s) {
                if(s == 0)
                    return int(PICSTEP_OK);
                if(copied[s])
                    cudaStreamWaitEvent(c->stream, copied[s], 0);
                else
                    cudaStreamSynchronize(c->side);
                return uploadSort(c, s);
            });
        for(auto e : copied)
            if(e)
                cudaEventDestroy(e);
        if(rc)
            return rc;
        CU(c, cudaMemcpyAsync(E, c->fieldMem[PICSTEP_FIELD_E], fbytes, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(B, c->fieldMem[PICSTEP_FIELD_B], fbytes, cudaMemcpyDeviceToHost, c->stream));
        if(energies4)
        {
            rc = picstep_reduce(c, PICSTEP_REDUCE_FIELD_ENERGY, 0, energies4);
            double pe[2] = {0, 0}, acc[2] = {0, 0};
            for(int s = 0; s < nSpecies && !rc; ++s)
            {
                rc = picstep_reduce(c, PICSTEP_REDUCE_PARTICLE_ENERGY, s, pe);
                acc[0] += pe[0];
                acc[1] += pe[1];
            }
            energies4[2] = acc[0];
            energies4[3] = acc[1];
            if(rc)
                return rc;
        }
        return checkFlags(c);
    }

    /* Overlap evidence of the decomposed step: out[0] = mean device time per step from "BORDER area pushed" to "exchange
     * of leaving particles and J guard strips complete" (second stream), out[1] = the same to "last CORE kernel complete"
     * (compute stream), out[2] = steps measured.  Recorded while picstep_stage_times is enabled; the call synchronises. */
    int picstep_overlap_times(picstep_ctx* c, float* out3)
    {
        if(!c || !out3)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        CU(c, cudaStreamSynchronize(c->stream));
        CU(c, cudaStreamSynchronize(c->side));
        CU(c, cudaStreamSynchronize(c->commStream));
        for(auto& o : c->overlapSpans)
        {
            float a = 0.0f, b = 0.0f;
            if(cudaEventElapsedTime(&a, o.border, o.comm) == cudaSuccess && cudaEventElapsedTime(&b, o.border, o.core) == cudaSuccess)
            {
                c->overlapMs[0] += a;
                c->overlapMs[1] += b;
                c->overlapSteps += 1;
            }
            c->evPool.push_back(o.border);
            c->evPool.push_back(o.comm);
            c->evPool.push_back(o.core);
        }
        c->overlapSpans.clear();
        out3[0] = c->overlapSteps ? float(c->overlapMs[0] / c->overlapSteps) : 0.0f;
        out3[1] = c->overlapSteps ? float(c->overlapMs[1] / c->overlapSteps) : 0.0f;
        out3[2] = float(c->overlapSteps);
        c->overlapMs[0] = c->overlapMs[1] = 0.0;
        c->overlapSteps = 0;
        return PICSTEP_OK;
    }

    int picstep_sync(picstep_ctx* c)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        return checkFlags(c);
    }

    int picstep_reduce(picstep_ctx* c, int32_t what, int32_t sp, double* out)
    {
        if(!c || !out)
            return PICSTEP_ERR_INVALID;
        CU(c, cudaSetDevice(c->device));
        DevParams const& P = c->P;
        double const V = double(P.cell[0]) * double(P.cell[1]) * double(P.cell[2]);
        if(what == PICSTEP_REDUCE_FIELD_ENERGY)
        {
            CU(c, cudaMemsetAsync(c->redBuf, 0, sizeof(double) * 2, c->stream));
            KL(c, 1, launchFieldEnergy(P, fieldOf(c, PICSTEP_FIELD_E), fieldOf(c, PICSTEP_FIELD_B), c->redBuf, c->stream));
            CU(c, cudaMemcpyAsync(c->hostPinned, c->redBuf, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            double const* r = reinterpret_cast<double const*>(c->hostPinned);
            out[0] = r[0] * (0.5 / double(P.mue0) * V);
            out[1] = r[1] * (double(P.eps0) * V * 0.5);
            return PICSTEP_OK;
        }
        if(sp < 0 || sp >= int(c->species.size()))
            if(what != PICSTEP_REDUCE_GAUSS && what != PICSTEP_REDUCE_SLOW_PATH)
                return PICSTEP_ERR_INVALID;
        if(what == PICSTEP_REDUCE_PARTICLE_ENERGY)
        {
            SpeciesHost& s = c->species[sp];
            CU(c, cudaMemsetAsync(c->redBuf, 0, sizeof(double) * 2, c->stream));
            if(s.capacity)
                KL(c, 1, launchParticleEnergy(P, devOf(c, s, s.cur), s.nDev + s.cur, s.lazy ? s.inv : nullptr, c->redBuf, c->stream));
            CU(c, cudaMemcpyAsync(c->hostPinned, c->redBuf, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            double const* r = reinterpret_cast<double const*>(c->hostPinned);
            out[0] = r[0];
            out[1] = r[1];
            return PICSTEP_OK;
        }
        if(what == PICSTEP_REDUCE_SLOW_PATH)
        {
            CU(c, cudaMemcpyAsync(c->hostPinned, c->P.stats, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaMemsetAsync(c->P.stats, 0, sizeof(unsigned long long) * 2, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            unsigned long long const* r = reinterpret_cast<unsigned long long const*>(c->hostPinned);
            out[0] = double(r[0]);
            out[1] = double(r[1]);
            return PICSTEP_OK;
        }
        if(what == PICSTEP_REDUCE_PARTICLE_COUNT)
        {
            int64_t n = 0;
            int rc = picstep_particles_count(c, sp, &n);
            out[0] = double(n);
            return rc;
        }
        if(what == PICSTEP_REDUCE_GAUSS)
        {
            if(!c->rho)
                CU(c, cudaMalloc(&c->rho, sizeof(float) * 3 * P.vol));
            CU(c, cudaMemsetAsync(c->rho, 0, sizeof(float) * 3 * P.vol, c->stream));
            for(auto& s : c->species)
                if(int rc = ensureSorted(c, s))
                    return rc;
            for(auto& s : c->species)
                if(s.capacity)
                    KL(c, 1, launchChargeDensity(s.shape, P, devOf(c, s, s.cur), s.cellOff[s.cur], c->rho, c->stream));
            // guard reduction of rho with the J machinery: temporarily view rho as a 3-component field whose
            // components 1,2 are zero (FieldTmp::asyncCommunication in the reference)
            float* saved = c->fieldMem[PICSTEP_FIELD_J];
            c->fieldMem[PICSTEP_FIELD_J] = c->rho;
            int rc = exchangeField(c, PICSTEP_FIELD_J);
            c->fieldMem[PICSTEP_FIELD_J] = saved;
            if(rc)
                return rc;
            CU(c, cudaMemsetAsync(c->flags + 3, 0, sizeof(int), c->stream));
            KL(c, 1, launchGaussResidual(P, fieldOf(c, PICSTEP_FIELD_E), c->rho, c->flags + 3, c->stream));
            CU(c, cudaMemcpyAsync(c->hostPinned, c->flags + 3, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            float mx;
            std::memcpy(&mx, c->hostPinned, sizeof(float));
            out[0] = double(mx * (P.cell[0] * P.cell[1] * P.cell[2]));
            return PICSTEP_OK;
        }
        return PICSTEP_ERR_INVALID;
    }

    // ---- parity hook: gather only -------------------------------------------------------------------------------
    // out[6][n]: E.x,E.y,E.z,B.x,B.y,B.z at the particles, frame-run order (not part of the reference surface;
    // exported for the FieldToParticleInterpolation parity test)
    int picstep_debug_gather(picstep_ctx* c, int32_t sp, int64_t capacity, float* out)
    {
        if(!c || sp < 0 || sp >= int(c->species.size()) || !out)
            return PICSTEP_ERR_INVALID;
        int64_t n = 0;
        int rc = picstep_particles_count(c, sp, &n);
        if(rc)
            return rc;
        if(n > capacity)
            return fail(c, PICSTEP_ERR_CAPACITY, "gather buffer too small");
        SpeciesHost& s = c->species[sp];
        if(int rc2 = ensureSorted(c, s))
            return rc2;
        float* tmp = nullptr;
        CU(c, cudaMalloc(&tmp, sizeof(float) * 6 * std::max<int64_t>(n, 1)));
        KL(c, 1, launchGather(s.shape, c->P, devOf(c, s, s.cur), fieldOf(c, PICSTEP_FIELD_E), fieldOf(c, PICSTEP_FIELD_B), s.cellOff[s.cur], tmp, n, c->stream));
        for(int k = 0; k < 6; ++k)
            CU(c, cudaMemcpyAsync(out + k * capacity, tmp + k * n, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFree(tmp);
        return PICSTEP_OK;
    }

    // ---- multi GPU ------------------------------------------------------------------------------------------------
    int picstep_comm_unique_id(void* id128)
    {
        std::string err;
        int const rc = commUniqueId(id128, err);
        if(rc)
            g_createErr = err;
        return rc ? PICSTEP_ERR_COMM : PICSTEP_OK;
    }

    int picstep_comm_init(picstep_ctx* c, const void* id128, int32_t rank, int32_t nranks)
    {
        if(!c || !id128)
            return PICSTEP_ERR_INVALID;
        if(nranks != c->nranks || rank != c->rank)
            return fail(c, PICSTEP_ERR_INVALID, "rank / nranks do not match devices and rank_pos of the context");
        CU(c, cudaSetDevice(c->device));
        int const rc = commInit(&c->comm, id128, rank, nranks, c->err);
        return rc ? PICSTEP_ERR_COMM : PICSTEP_OK;
    }

    // ---- measurement ----------------------------------------------------------------------------------------------
    int picstep_launch_count(picstep_ctx* c, int64_t* n)
    {
        if(!c || !n)
            return PICSTEP_ERR_INVALID;
        *n = c->launches;
        return PICSTEP_OK;
    }

    int picstep_stage_times(picstep_ctx* c, int32_t enable, float* ms7)
    {
        if(!c)
            return PICSTEP_ERR_INVALID;
        if(!c->spans.empty())
        {
            CU(c, cudaSetDevice(c->device));
            CU(c, cudaStreamSynchronize(c->stream));
            CU(c, cudaStreamSynchronize(c->side)); // the re-sort spans are recorded on the second stream
            for(auto& sp : c->spans)
            {
                float ms = 0;
                cudaEventElapsedTime(&ms, sp.a, sp.b);
                c->stageMs[sp.stage] += ms;
                c->evPool.push_back(sp.a);
                c->evPool.push_back(sp.b);
            }
            c->spans.clear();
        }
        if(ms7)
            for(int i = 0; i < NSTAGE; ++i)
                ms7[i] = c->stageMs[i];
        for(int i = 0; i < NSTAGE; ++i)
            c->stageMs[i] = 0;
        c->timing = enable != 0;
        return PICSTEP_OK;
    }

    int picstep_stream(picstep_ctx* c, void** stream)
    {
        if(!c || !stream)
            return PICSTEP_ERR_INVALID;
        *stream = c->stream;
        return PICSTEP_OK;
    }
}
