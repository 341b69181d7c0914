// Inter-rank communication of the adjoint path: ghost-cell / ghost-face copies between pipeline stages and
// the all-reduce of the GMRES dot products.  Replaces OpenFOAM's Pstream processor-patch exchange and the
// PETSc VecScatter/MPI_Allreduce of the reference.  Product build: NCCL grouped
// ncclSend/ncclRecv on the solver's stream over NVLink; the test-only host build takes two callbacks so that
// the same host logic can be driven by torch.distributed/gloo on CPUs.
#pragma once
#include "backend.hpp"
#include "partition.hpp"
#include "views.hpp"

#if !defined(DAB_HOSTSIM) && defined(DAB_WITH_NCCL)
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace dab
{

#if !defined(DAB_HOSTSIM) && defined(DAB_WITH_NCCL)
// NCCL is bound at run time (dlopen of libnccl.so.2) so that a process that also imports torch shares torch's
// bundled NCCL instead of pulling a second copy in by link order.
struct NcclApi
{
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    static NcclApi& get()
    {
        static NcclApi a;
        if (!a.lib)
        {
            a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!a.lib) throw Error(std::string("cannot load libnccl.so.2: ") + dlerror());
            auto sym = [&](const char* n) {
                void* p = dlsym(a.lib, n);
                if (!p) throw Error(std::string("libnccl.so.2 lacks ") + n);
                return p;
            };
            a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
            a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
            a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
            a.Send = (decltype(a.Send))sym("ncclSend");
            a.Recv = (decltype(a.Recv))sym("ncclRecv");
            a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
            a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
            a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
            a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        }
        return a;
    }
};
#endif

extern "C"
{
    // per peer p: sendBufs[p] (sendCounts[p] doubles) goes to peers[p]; recvBufs[p] receives recvCounts[p] doubles
    typedef void (*dab_exchange_fn)(void* ctx, int nPeers, const int* peers, const double* const* sendBufs, const int* sendCounts,
                                    double* const* recvBufs, const int* recvCounts);
    typedef void (*dab_allreduce_fn)(void* ctx, double* buf, int n);
}

struct Comm
{
    int rank = 0, size = 1;
    dab_exchange_fn cbExchange = nullptr;
    dab_allreduce_fn cbAllreduce = nullptr;
    void* cbCtx = nullptr;
#if !defined(DAB_HOSTSIM) && defined(DAB_WITH_NCCL)
    ncclComm_t nccl = nullptr;
#endif
    bool active() const { return size > 1; }

    void initNccl(Backend& be, int rank_, int size_, const void* uid)
    {
        rank = rank_;
        size = size_;
#if !defined(DAB_HOSTSIM) && defined(DAB_WITH_NCCL)
        if (!uid) throw Error("dab_create: n_ranks > 1 needs an NCCL unique id");
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        NcclApi& N = NcclApi::get();
        ncclResult_t r = N.CommInitRank(&nccl, size, id, rank);
        if (r != ncclSuccess) throw Error(std::string("ncclCommInitRank: ") + N.GetErrorString(r));
        (void)be;
#else
        (void)be;
        (void)uid;
        if (!cbExchange) throw Error("multi-rank creation needs NCCL (product build) or communication callbacks (test build)");
#endif
    }

    void destroy()
    {
#if !defined(DAB_HOSTSIM) && defined(DAB_WITH_NCCL)
        if (nccl) NcclApi::get().CommDestroy(nccl);
        nccl = nullptr;
#endif
    }

    void exchange(Backend& be, const std::vector<int>& peers, const std::vector<const double*>& sendBufs, const std::vector<int>& sendCounts,
                  const std::vector<double*>& recvBufs, const std::vector<int>& recvCounts)
    {
        if (!active() || peers.empty()) return;
#if !defined(DAB_HOSTSIM) && defined(DAB_WITH_NCCL)
        NcclApi& N = NcclApi::get();
        N.GroupStart();
        for (size_t p = 0; p < peers.size(); p++)
        {
            if (sendCounts[p] > 0) N.Send(sendBufs[p], (size_t)sendCounts[p], ncclDouble, peers[p], nccl, be.stream);
            if (recvCounts[p] > 0) N.Recv(recvBufs[p], (size_t)recvCounts[p], ncclDouble, peers[p], nccl, be.stream);
        }
        ncclResult_t r = N.GroupEnd();
        if (r != ncclSuccess) throw Error(std::string("nccl halo exchange: ") + N.GetErrorString(r));
        be.launches++;
#else
        be.sync();
        cbExchange(cbCtx, (int)peers.size(), peers.data(), sendBufs.data(), sendCounts.data(), recvBufs.data(), recvCounts.data());
#endif
    }

    // in-place sum over ranks of n doubles in device memory
    void allreduceSum(Backend& be, double* dev, int n)
    {
        if (!active()) return;
#if !defined(DAB_HOSTSIM) && defined(DAB_WITH_NCCL)
        NcclApi& N = NcclApi::get();
        ncclResult_t r = N.AllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, nccl, be.stream);
        if (r != ncclSuccess) throw Error(std::string("ncclAllReduce: ") + N.GetErrorString(r));
        be.launches++;
#else
        be.sync();
        cbAllreduce(cbCtx, dev, n);
#endif
    }
};

// ---- pack / unpack kernels ---------------------------------------------------------------------------------
struct HaloPack
{
    const double* arr;
    int cellStride, compStride, ncomp;
    const int32_t* idx;     // [nSend] local indices to send (concatenated over peers)
    const int32_t* segOff;  // [nSend] start of the element's peer segment
    const int32_t* segCnt;  // [nSend] size of the element's peer segment
    int nSend, sumComp, compBase;
    double* buf;
    DAB_HD void operator()(int i) const
    {
        const int k = i / nSend, j = i - k * nSend;
        const int64_t o = (int64_t)segOff[j] * sumComp + (int64_t)(compBase + k) * segCnt[j] + (j - segOff[j]);
        buf[o] = arr[(int64_t)idx[j] * cellStride + (int64_t)k * compStride];
    }
};
struct HaloUnpack
{
    double* arr;
    int cellStride, compStride, ncomp;
    const int32_t* idx;     // [nRecv] local indices to fill
    const int32_t* segOff;
    const int32_t* segCnt;
    int nRecv, sumComp, compBase;
    const double* buf;
    DAB_HD void operator()(int i) const
    {
        const int k = i / nRecv, j = i - k * nRecv;
        const int64_t o = (int64_t)segOff[j] * sumComp + (int64_t)(compBase + k) * segCnt[j] + (j - segOff[j]);
        arr[(int64_t)idx[j] * cellStride + (int64_t)k * compStride] = buf[o];
    }
};

// one index set (cells or faces) exchanged with every peer
struct HaloSet
{
    int nSend = 0, nRecv = 0;
    std::vector<int> sendOffPeer, sendCntPeer, recvOffPeer, recvCntPeer; // per peer
    DevBuf<int32_t> dSendIdx, dSendSegOff, dSendSegCnt, dRecvIdx, dRecvSegOff, dRecvSegCnt;
    DevBuf<double> sendBuf, recvBuf;
    int capComp = 0;
#ifndef DAB_HOSTSIM
    cudaEvent_t evPack = nullptr, evDone = nullptr;
#endif
    bool pending = false;

    void build(Backend& be, const std::vector<std::vector<int32_t>>& send, const std::vector<std::vector<int32_t>>& recv)
    {
        std::vector<int32_t> si, so, sc, ri, ro, rc;
        for (size_t p = 0; p < send.size(); p++)
        {
            sendOffPeer.push_back((int)si.size());
            sendCntPeer.push_back((int)send[p].size());
            for (int32_t c : send[p])
            {
                si.push_back(c);
                so.push_back(sendOffPeer.back());
                sc.push_back((int32_t)send[p].size());
            }
            recvOffPeer.push_back((int)ri.size());
            recvCntPeer.push_back((int)recv[p].size());
            for (int32_t c : recv[p])
            {
                ri.push_back(c);
                ro.push_back(recvOffPeer.back());
                rc.push_back((int32_t)recv[p].size());
            }
        }
        nSend = (int)si.size();
        nRecv = (int)ri.size();
        dSendIdx.upload(be, si); dSendSegOff.upload(be, so); dSendSegCnt.upload(be, sc);
        dRecvIdx.upload(be, ri); dRecvSegOff.upload(be, ro); dRecvSegCnt.upload(be, rc);
    }
    void reserve(Backend& be, int sumComp)
    {
        if (sumComp <= capComp) return;
        sendBuf.alloc(be, (size_t)nSend * sumComp + 1, false);
        recvBuf.alloc(be, (size_t)nRecv * sumComp + 1, false);
        capComp = sumComp;
    }
};

struct HaloItem
{
    double* arr;
    int ncomp, cellStride, compStride;
};

struct Halo
{
    Comm* comm = nullptr;
    Backend* be = nullptr;
    std::vector<int> peers;
    HaloSet cells, faces;
    long exchanges = 0;

    void build(Backend& b, Comm& c, const HaloPlan& plan)
    {
        be = &b;
        comm = &c;
        peers = plan.peers;
        std::vector<std::vector<int32_t>> recvCells(plan.peers.size());
        for (size_t p = 0; p < plan.peers.size(); p++)
            for (int i = 0; i < plan.recvCellCount[p]; i++) recvCells[p].push_back(plan.recvCellStart[p] + i);
        cells.build(b, plan.sendCells, recvCells);
        faces.build(b, plan.sendFaces, plan.recvFaces);
    }

    void run(HaloSet& hs, const std::vector<HaloItem>& items)
    {
        if (!comm || !comm->active()) return;
        int sumComp = 0;
        for (const auto& it : items) sumComp += it.ncomp;
        hs.reserve(*be, sumComp);
        int base = 0;
        for (const auto& it : items)
        {
            be->launch(hs.nSend * it.ncomp, HaloPack{it.arr, it.cellStride, it.compStride, it.ncomp, hs.dSendIdx.p, hs.dSendSegOff.p,
                                                     hs.dSendSegCnt.p, hs.nSend, sumComp, base, hs.sendBuf.p});
            base += it.ncomp;
        }
        std::vector<const double*> sb;
        std::vector<double*> rb;
        std::vector<int> sc, rc;
        for (size_t p = 0; p < peers.size(); p++)
        {
            sb.push_back(hs.sendBuf.p + (size_t)hs.sendOffPeer[p] * sumComp);
            rb.push_back(hs.recvBuf.p + (size_t)hs.recvOffPeer[p] * sumComp);
            sc.push_back(hs.sendCntPeer[p] * sumComp);
            rc.push_back(hs.recvCntPeer[p] * sumComp);
        }
        comm->exchange(*be, peers, sb, sc, rb, rc);
        base = 0;
        for (const auto& it : items)
        {
            be->launch(hs.nRecv * it.ncomp, HaloUnpack{it.arr, it.cellStride, it.compStride, it.ncomp, hs.dRecvIdx.p, hs.dRecvSegOff.p,
                                                       hs.dRecvSegCnt.p, hs.nRecv, sumComp, base, hs.recvBuf.p});
            base += it.ncomp;
        }
        exchanges++;
    }
    void exchangeCells(const std::vector<HaloItem>& items) { run(cells, items); }
    void exchangeFaces(const std::vector<HaloItem>& items) { run(faces, items); }

    // asynchronous variant: pack on the compute stream, exchange + unpack on the communication stream; the caller
    // launches work that does not touch ghost slots in between and calls finish() before the work that does
    void start(HaloSet& hs, const std::vector<HaloItem>& items)
    {
        if (!comm || !comm->active()) return;
#ifdef DAB_HOSTSIM
        run(hs, items);
#else
        if (!hs.evPack)
        {
            cudaEventCreateWithFlags(&hs.evPack, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&hs.evDone, cudaEventDisableTiming);
        }
        int sumComp = 0;
        for (const auto& it : items) sumComp += it.ncomp;
        hs.reserve(*be, sumComp);
        int base = 0;
        for (const auto& it : items)
        {
            be->launch(hs.nSend * it.ncomp, HaloPack{it.arr, it.cellStride, it.compStride, it.ncomp, hs.dSendIdx.p, hs.dSendSegOff.p,
                                                     hs.dSendSegCnt.p, hs.nSend, sumComp, base, hs.sendBuf.p});
            base += it.ncomp;
        }
        cudaEventRecord(hs.evPack, be->stream);
        std::swap(be->stream, be->stream2); // everything below goes to the communication stream
        cudaStreamWaitEvent(be->stream, hs.evPack, 0);
        std::vector<const double*> sb;
        std::vector<double*> rb;
        std::vector<int> sc, rc;
        for (size_t p = 0; p < peers.size(); p++)
        {
            sb.push_back(hs.sendBuf.p + (size_t)hs.sendOffPeer[p] * sumComp);
            rb.push_back(hs.recvBuf.p + (size_t)hs.recvOffPeer[p] * sumComp);
            sc.push_back(hs.sendCntPeer[p] * sumComp);
            rc.push_back(hs.recvCntPeer[p] * sumComp);
        }
        comm->exchange(*be, peers, sb, sc, rb, rc);
        base = 0;
        for (const auto& it : items)
        {
            be->launch(hs.nRecv * it.ncomp, HaloUnpack{it.arr, it.cellStride, it.compStride, it.ncomp, hs.dRecvIdx.p, hs.dRecvSegOff.p,
                                                       hs.dRecvSegCnt.p, hs.nRecv, sumComp, base, hs.recvBuf.p});
            base += it.ncomp;
        }
        cudaEventRecord(hs.evDone, be->stream);
        std::swap(be->stream, be->stream2);
        hs.pending = true;
        exchanges++;
#endif
    }
    void finish()
    {
#ifndef DAB_HOSTSIM
        for (HaloSet* hs : {&cells, &faces})
            if (hs->pending)
            {
                cudaStreamWaitEvent(be->stream, hs->evDone, 0);
                hs->pending = false;
            }
#endif
    }
};

} // namespace dab
