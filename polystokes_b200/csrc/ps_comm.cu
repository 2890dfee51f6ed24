// ps_comm.cu -- NCCL binding (product) / host-callback binding (emulation twin) of ps_comm.hpp.
#include "ps_comm.hpp"

#ifndef PS_EMULATE
#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only: every entry point is resolved with dlsym below
#include <mutex>
#endif

namespace ps {

#ifndef PS_EMULATE
namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommAbort) CommAbort = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    // prefer the copy already in the process (torch's bundled libnccl), then the system one
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw Error(std::string("ps_comm: cannot load libnccl.so.2: ") + dlerror());
#define PS_NCCL_SYM(field, sym) api.field = (decltype(api.field))dlsym(lib, sym); if (!api.field) throw Error(std::string("ps_comm: libnccl lacks ") + sym)
    PS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    PS_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    PS_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    PS_NCCL_SYM(CommAbort, "ncclCommAbort");
    PS_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    PS_NCCL_SYM(AllReduce, "ncclAllReduce");
    PS_NCCL_SYM(AllGather, "ncclAllGather");
    PS_NCCL_SYM(Send, "ncclSend");
    PS_NCCL_SYM(Recv, "ncclRecv");
    PS_NCCL_SYM(GroupStart, "ncclGroupStart");
    PS_NCCL_SYM(GroupEnd, "ncclGroupEnd");
#undef PS_NCCL_SYM
    api.lib = lib;
    return api;
}

void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw Error(std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(r));
}
#define PS_NCCL(x) nccl_check((x), #x)

struct NcclComm : Comm {
    ncclComm_t comm = nullptr;
    std::mutex m; bool aborted = false;
    ~NcclComm() override { std::lock_guard<std::mutex> lk(m); if (comm) { if (aborted) nccl().CommAbort(comm); else nccl().CommDestroy(comm); } }
    void abort() override { std::lock_guard<std::mutex> lk(m); if (comm && !aborted) { aborted = true; nccl().CommAbort(comm); comm = nullptr; } }
    void alive() { if (aborted || !comm) throw Error("the communicator was aborted (another rank of this handle failed)"); }
    void allreduce_sum(double* buf, int n, cudaStream_t st) override {
        alive();
        PS_NCCL(nccl().AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, comm, st));
        PS_COUNT_LAUNCH(1);
    }
    void allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        alive();
        PS_NCCL(nccl().AllGather(send, recv, bytes, ncclChar, comm, st));
        PS_COUNT_LAUNCH(1);
    }
    void sendrecv(int npeers, const int* peers, const void* const* sendBuf, const size_t* sendBytes, void* const* recvBuf, const size_t* recvBytes, cudaStream_t st) override {
        bool any = false;
        for (int i = 0; i < npeers; ++i) if (peers[i] >= 0 && (sendBytes[i] || recvBytes[i])) any = true;
        if (!any) return;
        alive();
        PS_NCCL(nccl().GroupStart());
        for (int i = 0; i < npeers; ++i) {
            if (peers[i] < 0) continue;
            if (sendBytes[i]) PS_NCCL(nccl().Send(sendBuf[i], sendBytes[i], ncclChar, peers[i], comm, st));
            if (recvBytes[i]) PS_NCCL(nccl().Recv(recvBuf[i], recvBytes[i], ncclChar, peers[i], comm, st));
        }
        PS_NCCL(nccl().GroupEnd());
        PS_COUNT_LAUNCH(1);
    }
};

}  // namespace

void nccl_unique_id(void* id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    PS_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
}

Comm* make_nccl_comm(int rank, int nranks, const void* id128) {
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NcclComm* c = new NcclComm;
    c->rank = rank; c->nranks = nranks;
    try { PS_NCCL(nccl().CommInitRank(&c->comm, nranks, id, rank)); }
    catch (...) { c->comm = nullptr; delete c; throw; }
    return c;
}

#else  // ---- PS_EMULATE: host callbacks ("device" memory is host memory in this build) ----

namespace {
struct CallbackComm : Comm {
    ps_allreduce_cb ar; ps_sendrecv_cb sr; void* ctx;
    void allreduce_sum(double* buf, int n, cudaStream_t) override { ar(ctx, buf, n); }
    void allgather(const void*, void*, size_t, cudaStream_t) override { throw Error("allgather: not available in the emulation twin"); }
    void sendrecv(int npeers, const int* peers, const void* const* sendBuf, const size_t* sendBytes, void* const* recvBuf, const size_t* recvBytes, cudaStream_t) override {
        sr(ctx, npeers, peers, sendBuf, sendBytes, recvBuf, recvBytes);
    }
};
}  // namespace

Comm* make_callback_comm(int rank, int nranks, ps_allreduce_cb ar, ps_sendrecv_cb sr, void* ctx) {
    CallbackComm* c = new CallbackComm;
    c->rank = rank; c->nranks = nranks; c->ar = ar; c->sr = sr; c->ctx = ctx;
    return c;
}
#endif

}  // namespace ps
