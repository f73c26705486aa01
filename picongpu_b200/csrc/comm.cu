// comm.cu — NCCL transport for guard-cell and particle exchange between neighbouring sub-domains
// (replaces pmacc::CommunicatorMPI::startSend/startReceive, include/pmacc/communication/CommunicatorMPI.cpp:114-149).
//
// NCCL is loaded at run time with dlopen (the torch wheel ships libnccl.so.2; no link-time dependency, so the
// library also loads on machines without NCCL as long as devices == 1).  One ncclGroup per exchange phase: both
// directions of an axis travel together, the transfer is stream ordered behind the pack kernels.
#include "common.cuh"

#include <dlfcn.h>

#include <cstdlib>
#include <string>

namespace picstep
{
    namespace
    {
        struct ncclUniqueId_
        {
            char internal[128];
        };
        using ncclComm_t_ = void*;
        enum
        {
            ncclSuccess_ = 0
        };
        enum
        {
            ncclChar_ = 0
        };

        struct Api
        {
            void* handle = nullptr;
            int (*GetUniqueId)(ncclUniqueId_*) = nullptr;
            int (*CommInitRank)(ncclComm_t_*, int, ncclUniqueId_, int) = nullptr;
            int (*CommDestroy)(ncclComm_t_) = nullptr;
            int (*Send)(void const*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
            int (*Recv)(void*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
            int (*GroupStart)() = nullptr;
            int (*GroupEnd)() = nullptr;
            char const* (*GetErrorString)(int) = nullptr;
        };

        Api g_api;

        bool loadApi(std::string& err)
        {
            if(g_api.handle)
                return true;
            char const* names[4] = {std::getenv("PICSTEP_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
            void* h = nullptr;
            // prefer a copy that is already mapped into the process (e.g. by torch.distributed)
            for(int i = 0; i < 3 && !h; ++i)
                if(names[i])
                    h = dlopen(names[i], RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            for(int i = 0; i < 3 && !h; ++i)
                if(names[i])
                    h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
            if(!h)
            {
                err = std::string("cannot load NCCL (set PICSTEP_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
                return false;
            }
            Api a;
            a.handle = h;
#define SYM(field, name)                                                                                              \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                                                   \
    if(!a.field)                                                                                                      \
    {                                                                                                                 \
        err = std::string("NCCL symbol missing: ") + name;                                                            \
        return false;                                                                                                 \
    }
            SYM(GetUniqueId, "ncclGetUniqueId")
            SYM(CommInitRank, "ncclCommInitRank")
            SYM(CommDestroy, "ncclCommDestroy")
            SYM(Send, "ncclSend")
            SYM(Recv, "ncclRecv")
            SYM(GroupStart, "ncclGroupStart")
            SYM(GroupEnd, "ncclGroupEnd")
            SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
            g_api = a;
            return true;
        }
    } // namespace

    struct Comm
    {
        ncclComm_t_ comm = nullptr;
        int rank = 0, nranks = 1;
    };

    int commUniqueId(void* id128, std::string& err)
    {
        if(!loadApi(err))
            return 1;
        ncclUniqueId_ id;
        int const rc = g_api.GetUniqueId(&id);
        if(rc != ncclSuccess_)
        {
            err = std::string("ncclGetUniqueId: ") + g_api.GetErrorString(rc);
            return 1;
        }
        memcpy(id128, &id, sizeof(id));
        return 0;
    }

    int commInit(Comm** out, void const* id128, int rank, int nranks, std::string& err)
    {
        if(!loadApi(err))
            return 1;
        ncclUniqueId_ id;
        memcpy(&id, id128, sizeof(id));
        auto* c = new Comm();
        c->rank = rank;
        c->nranks = nranks;
        int const rc = g_api.CommInitRank(&c->comm, nranks, id, rank);
        if(rc != ncclSuccess_)
        {
            err = std::string("ncclCommInitRank: ") + g_api.GetErrorString(rc);
            delete c;
            return 1;
        }
        *out = c;
        return 0;
    }

    void commDestroy(Comm* c)
    {
        if(c)
        {
            if(c->comm && g_api.CommDestroy)
                g_api.CommDestroy(c->comm);
            delete c;
        }
    }

    int commSendRecv(Comm* c, void const* sendLo, size_t nSendLo, void* recvLo, size_t nRecvLo, int rankLo, void const* sendHi, size_t nSendHi, void* recvHi, size_t nRecvHi, int rankHi, cudaStream_t st, std::string& err)
    {
        int rc = g_api.GroupStart();
        // Per peer NCCL matches the k-th send with the k-th receive.  With two ranks on a periodic axis the lower and
        // the upper neighbour are the same peer: what I send downwards must arrive in its "from upper" buffer, so
        // sends are posted (lower, upper) and receives (upper, lower).
        if(rc == ncclSuccess_ && rankLo >= 0 && nSendLo)
            rc = g_api.Send(sendLo, nSendLo, ncclChar_, rankLo, c->comm, st);
        if(rc == ncclSuccess_ && rankHi >= 0 && nSendHi)
            rc = g_api.Send(sendHi, nSendHi, ncclChar_, rankHi, c->comm, st);
        if(rc == ncclSuccess_ && rankHi >= 0 && nRecvHi)
            rc = g_api.Recv(recvHi, nRecvHi, ncclChar_, rankHi, c->comm, st);
        if(rc == ncclSuccess_ && rankLo >= 0 && nRecvLo)
            rc = g_api.Recv(recvLo, nRecvLo, ncclChar_, rankLo, c->comm, st);
        int const rc2 = g_api.GroupEnd();
        if(rc == ncclSuccess_)
            rc = rc2;
        if(rc != ncclSuccess_)
        {
            err = std::string("NCCL send/recv: ") + g_api.GetErrorString(rc);
            return 1;
        }
        return 0;
    }
} // namespace picstep
