// Domain decomposition of the host mesh: recursive coordinate bisection of the cell centres and
// extraction of one rank's sub-mesh with a one-cell ghost layer.  Replaces, for the adjoint hot path, the
// reference's `decomposePar` subprocess (scotch; reference dafoam/pyDAFoam.py:1454-1480, 597-604) and
// OpenFOAM's processor patches: cut faces stay *internal* faces of the local mesh whose other cell is a
// ghost cell appended after the owned cells, so the gather kernels run unchanged and every inter-rank
// dependency becomes a ghost-cell copy.
#pragma once
#include "mesh.hpp"
#include <algorithm>
#include <numeric>

namespace dab
{

struct HaloPlan
{
    std::vector<int> peers;                    // neighbour ranks, ascending
    std::vector<std::vector<int32_t>> sendCells; // per peer: owned local cells the peer holds as ghosts
    std::vector<int> recvCellStart, recvCellCount; // per peer: contiguous ghost slots [start, start+count)
    std::vector<std::vector<int32_t>> sendFaces; // per peer: local faces whose phi this rank owns and the peer needs
    std::vector<std::vector<int32_t>> recvFaces; // per peer: local (foreign) faces filled from the peer
};

struct Partition
{
    int rank = 0, nRanks = 1;
    int64_t nGlobalCells = 0;
    std::vector<int32_t> cellGlobal; // local (owned + ghost) -> global cell
    std::vector<int32_t> faceGlobal; // local face -> global face
    std::vector<uint8_t> faceOwned;  // 1 if the phi DOF of the local face belongs to this rank
    HaloPlan halo;
};

// recursive coordinate bisection: part[c] in [0, nParts)
inline void rcbPartition(const HostMesh& g, int nParts, std::vector<int>& part)
{
    part.assign(g.nC, 0);
    std::vector<int> idx(g.nC);
    std::iota(idx.begin(), idx.end(), 0);
    struct Job { int lo, hi, p0, np; };
    std::vector<Job> stack{{0, g.nC, 0, nParts}};
    while (!stack.empty())
    {
        Job j = stack.back();
        stack.pop_back();
        if (j.np == 1)
        {
            for (int i = j.lo; i < j.hi; i++) part[idx[i]] = j.p0;
            continue;
        }
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int i = j.lo; i < j.hi; i++)
            for (int k = 0; k < 3; k++)
            {
                mn[k] = std::min(mn[k], g.C[k][idx[i]]);
                mx[k] = std::max(mx[k], g.C[k][idx[i]]);
            }
        int ax = 0;
        for (int k = 1; k < 3; k++)
            if (mx[k] - mn[k] > mx[ax] - mn[ax]) ax = k;
        const int npL = j.np / 2, npR = j.np - npL;
        const int mid = j.lo + (int)((int64_t)(j.hi - j.lo) * npL / j.np);
        std::nth_element(idx.begin() + j.lo, idx.begin() + mid, idx.begin() + j.hi,
                         [&](int a, int b) { return g.C[ax][a] < g.C[ax][b] || (g.C[ax][a] == g.C[ax][b] && a < b); });
        stack.push_back({j.lo, mid, j.p0, npL});
        stack.push_back({mid, j.hi, j.p0 + npL, npR});
    }
}

// build the local mesh of `rank` from the global mesh (geometry and wall distance already computed)
inline void extractLocalMesh(const HostMesh& g, const std::vector<int>& part, int rank, int nRanks, HostMesh& l, Partition& P)
{
    P.rank = rank;
    P.nRanks = nRanks;
    P.nGlobalCells = g.nC;
    std::vector<int32_t> g2l(g.nC, -1);
    // owned cells: first the interior ones (no neighbour on another rank), then the ones along the cuts, each group
    // in global order -- kernels can then run the interior range while the ghost exchange is in flight
    std::vector<uint8_t> onCut(g.nC, 0);
    for (int f = 0; f < g.nIF; f++)
        if (part[g.own[f]] != part[g.nei[f]])
        {
            onCut[g.own[f]] = 1;
            onCut[g.nei[f]] = 1;
        }
    int nC = 0;
    for (int pass = 0; pass < 2; pass++)
    {
        for (int c = 0; c < g.nC; c++)
            if (part[c] == rank && onCut[c] == pass)
            {
                g2l[c] = nC++;
                P.cellGlobal.push_back(c);
            }
        if (pass == 0) l.nInterior = nC;
    }
    // ghost cells grouped by owning rank, then global id
    std::vector<std::pair<int, int>> ghosts; // (rank, global)
    for (int f = 0; f < g.nIF; f++)
    {
        const int a = g.own[f], b = g.nei[f];
        if (part[a] == rank && part[b] != rank) ghosts.emplace_back(part[b], b);
        if (part[b] == rank && part[a] != rank) ghosts.emplace_back(part[a], a);
    }
    std::sort(ghosts.begin(), ghosts.end());
    ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
    int nT = nC;
    for (auto& pr : ghosts)
    {
        if (P.halo.peers.empty() || P.halo.peers.back() != pr.first)
        {
            P.halo.peers.push_back(pr.first);
            P.halo.recvCellStart.push_back(nT);
            P.halo.recvCellCount.push_back(0);
        }
        g2l[pr.second] = nT++;
        P.cellGlobal.push_back(pr.second);
        P.halo.recvCellCount.back()++;
    }
    const int nPeers = (int)P.halo.peers.size();
    auto peerIndex = [&](int q) { return (int)(std::lower_bound(P.halo.peers.begin(), P.halo.peers.end(), q) - P.halo.peers.begin()); };
    // cells to send: owned cells adjacent to a cell of peer q (sorted by global id = local order)
    P.halo.sendCells.assign(nPeers, {});
    {
        std::vector<std::pair<int, int>> snd; // (peer, local cell)
        for (int f = 0; f < g.nIF; f++)
        {
            const int a = g.own[f], b = g.nei[f];
            if (part[a] == rank && part[b] != rank) snd.emplace_back(part[b], g2l[a]);
            if (part[b] == rank && part[a] != rank) snd.emplace_back(part[a], g2l[b]);
        }
        // the peer lists its ghosts by global id: send in the same order
        std::sort(snd.begin(), snd.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
            return a.first != b.first ? a.first < b.first : P.cellGlobal[a.second] < P.cellGlobal[b.second];
        });
        snd.erase(std::unique(snd.begin(), snd.end()), snd.end());
        for (auto& pr : snd) P.halo.sendCells[peerIndex(pr.first)].push_back(pr.second);
    }
    // faces: A (both owned), B (owner mine, neighbour ghost), D (owner ghost, neighbour mine), then boundary by patch
    std::vector<int> fa, fb, fd;
    for (int f = 0; f < g.nIF; f++)
    {
        const bool oa = part[g.own[f]] == rank, ob = part[g.nei[f]] == rank;
        if (oa && ob) fa.push_back(f);
        else if (oa) fb.push_back(f);
        else if (ob) fd.push_back(f);
    }
    std::vector<int> lf(fa);
    lf.insert(lf.end(), fb.begin(), fb.end());
    lf.insert(lf.end(), fd.begin(), fd.end());
    const int nIF = (int)lf.size();
    l.patches.clear();
    for (size_t p = 0; p < g.patches.size(); p++)
    {
        PatchDef pd = g.patches[p];
        const int start = (int)lf.size();
        for (int i = 0; i < g.patches[p].size; i++)
        {
            const int f = g.patches[p].start + i;
            if (part[g.own[f]] == rank) lf.push_back(f);
        }
        pd.start = start;
        pd.size = (int)lf.size() - start;
        l.patches.push_back(pd);
    }
    const int nF = (int)lf.size();
    P.faceGlobal.assign(lf.begin(), lf.end());
    P.faceOwned.assign(nF, 1);
    P.halo.sendFaces.assign(nPeers, {});
    P.halo.recvFaces.assign(nPeers, {});
    for (int i = 0; i < nIF; i++)
    {
        const int f = lf[i];
        const int ra = part[g.own[f]], rb = part[g.nei[f]];
        if (ra == rank && rb != rank) P.halo.sendFaces[peerIndex(rb)].push_back(i); // B: ascending global id
        if (ra != rank && rb == rank)
        {
            P.faceOwned[i] = 0;
            P.halo.recvFaces[peerIndex(ra)].push_back(i); // D: ascending global id
        }
    }
    // topology
    l.points = g.points;
    l.nP = g.nP;
    l.nF = nF; l.nIF = nIF; l.nBF = nF - nIF; l.nC = nC; l.nCtot = nT;
    l.own.resize(nF);
    l.nei.resize(nIF);
    l.fOff.assign(1, 0);
    l.fLab.clear();
    for (int i = 0; i < nF; i++)
    {
        const int f = lf[i];
        l.own[i] = g2l[g.own[f]];
        if (i < nIF) l.nei[i] = g2l[g.nei[f]];
        for (int q = g.fOff[f]; q < g.fOff[f + 1]; q++) l.fLab.push_back(g.fLab[q]);
        l.fOff.push_back((int32_t)l.fLab.size());
    }
    l.patchGeom = g.patchGeom;
    l.bPatch.assign(l.nBF, -1);
    for (size_t p = 0; p < l.patches.size(); p++)
        for (int i = 0; i < l.patches[p].size; i++) l.bPatch[l.patches[p].start - nIF + i] = (int32_t)p;
    // ELL cell -> faces for the owned cells
    std::vector<int> cnt(nC, 0);
    for (int i = 0; i < nF; i++)
    {
        if (l.own[i] < nC) cnt[l.own[i]]++;
        if (i < nIF && l.nei[i] < nC) cnt[l.nei[i]]++;
    }
    l.maxCF = nC ? *std::max_element(cnt.begin(), cnt.end()) : 0;
    l.cellFaces.assign((size_t)l.maxCF * nC, -1);
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int i = 0; i < nF; i++)
    {
        if (l.own[i] < nC) l.cellFaces[(size_t)cnt[l.own[i]]++ * nC + l.own[i]] = (i << 1);
        if (i < nIF && l.nei[i] < nC) l.cellFaces[(size_t)cnt[l.nei[i]]++ * nC + l.nei[i]] = (i << 1) | 1;
    }
    // geometry slices
    for (int k = 0; k < 3; k++)
    {
        l.Sf[k].resize(nF); l.Cf[k].resize(nF); l.corr[k].resize(nF); l.C[k].resize(nT);
    }
    l.magSf.resize(nF); l.w.resize(nF); l.delta.resize(nF); l.V.resize(nT); l.yWall.resize(nT);
    for (int i = 0; i < nF; i++)
    {
        const int f = lf[i];
        for (int k = 0; k < 3; k++) { l.Sf[k][i] = g.Sf[k][f]; l.Cf[k][i] = g.Cf[k][f]; l.corr[k][i] = g.corr[k][f]; }
        l.magSf[i] = g.magSf[f]; l.w[i] = g.w[f]; l.delta[i] = g.delta[f];
    }
    for (int c = 0; c < nT; c++)
    {
        const int gc = P.cellGlobal[c];
        for (int k = 0; k < 3; k++) l.C[k][c] = g.C[k][gc];
        l.V[c] = g.V[gc];
        l.yWall[c] = g.yWall[gc];
    }
}

} // namespace dab
