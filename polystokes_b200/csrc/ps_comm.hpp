// ps_comm.hpp -- the two collectives of the distributed CG: a grouped halo
// send/receive with the z-neighbours and a tiny in-place sum all-reduce, both enqueued on the solver's stream.
//
// Product build: NCCL (over NVLink / NVSwitch), resolved with dlopen at ps_comm_init time so that a single-GPU
// process never needs libnccl.  The library is picked up from the process if torch already loaded it
// (torch.distributed is only the plumbing that carries the 128-byte unique id between ranks).
// Test-only build (-DPS_EMULATE): host callbacks, so the N > 1 logic (partition, halo lists, distributed CG)
// runs under torch.distributed/gloo on machines without a GPU.
#pragma once
#include "ps_rt.hpp"

namespace ps {

struct Comm {
    int rank = 0, nranks = 1;
    virtual ~Comm() {}
    // may be called from ANOTHER host thread: makes the collectives in flight on this communicator give up (a rank of the same
    // process failed and will never join them).  The communicator is unusable afterwards.
    virtual void abort() {}
    // in-place sum of n doubles living in device memory
    virtual void allreduce_sum(double* buf, int n, cudaStream_t st) = 0;
    // every rank contributes `bytes` from send; recv receives nranks * bytes in rank order (device buffers)
    virtual void allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) = 0;
    // one grouped exchange: for every i send sendBytes[i] from sendBuf[i] to peers[i] and receive recvBytes[i] into recvBuf[i]
    virtual void sendrecv(int npeers, const int* peers, const void* const* sendBuf, const size_t* sendBytes,
                          void* const* recvBuf, const size_t* recvBytes, cudaStream_t st) = 0;
};

#ifndef PS_EMULATE
void nccl_unique_id(void* id128);
Comm* make_nccl_comm(int rank, int nranks, const void* id128);
#else
typedef void (*ps_allreduce_cb)(void* ctx, double* buf, int n);
typedef void (*ps_sendrecv_cb)(void* ctx, int npeers, const int* peers, const void* const* sendBuf, const size_t* sendBytes,
                               void* const* recvBuf, const size_t* recvBytes);
Comm* make_callback_comm(int rank, int nranks, ps_allreduce_cb ar, ps_sendrecv_cb sr, void* ctx);
#endif

}  // namespace ps
