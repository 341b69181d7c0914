// Tile map of the CTA-resident product kernels (tile_kernels.hpp).
//
// A tile is a contiguous range of T owned cells [t*T, min((t+1)*T, nC)) of the mesh's own cell numbering -- so the tile's
// per-cell arrays are contiguous in HBM and load coalesced -- plus R halo rings (ring r = cells at face distance r from
// the tile).  Every tile gets a local numbering [tile cells | ring 1 | ring 2 | ...]; the cells of rings < R (the ones
// a kernel stage is evaluated on) get a local copy of their rows of the two ELL tables with the neighbour translated
// to the local index.  Nothing here assumes a structured mesh; how small the halo is depends on the cell numbering
// (tile-major numbering of the generator / a bandwidth-reducing renumbering gives compact tiles).
#pragma once
#include "mesh.hpp"
#include <vector>

namespace dab
{

struct TileMap
{
    int T = 0, nTiles = 0, R = 0, maxCF = 0;
    int ln = 0; // row stride of the local tables (>= max over tiles of the cells in rings < R), multiple of 32
    int ls = 0; // stride of gid (>= max over tiles of all local cells), multiple of 32
    int maxRun = 0, maxAll = 0;
    long sumAll = 0, sumRun = 0;
    std::vector<int32_t> cum; // [nTiles][R+1]: local cells up to and including ring r
    std::vector<int32_t> gid; // [nTiles][ls]: cell id of local cell l (-1 padding)
    std::vector<int32_t> tf;  // [nTiles][maxCF][ln]: face code (f<<1 | isNeighbour, -1 padding) of local cell l, slot k
    std::vector<int32_t> tn;  // [nTiles][maxCF][ln]: local index of the cell across that face (-1: boundary / padding)

    // sizes only (no tables): lets the caller test whether a tile size fits the compiled capacities
    static void measure(const HostMesh& hm, int T, int R, int& maxRun, int& maxAll, long& sumAll)
    {
        TileMap tmp;
        tmp.build(hm, T, R, false);
        maxRun = tmp.maxRun;
        maxAll = tmp.maxAll;
        sumAll = tmp.sumAll;
    }

    void build(const HostMesh& hm, int T_, int R_, bool tables = true)
    {
        T = T_;
        R = R_;
        maxCF = hm.maxCF;
        const int nC = hm.nC;
        nTiles = (nC + T - 1) / T;
        cum.assign((size_t)nTiles * (R + 1), 0);
        std::vector<int32_t> loc(hm.nCtot, -1);
        std::vector<std::vector<int32_t>> lists(tables ? nTiles : 0);
        std::vector<int32_t> cur;
        maxRun = maxAll = 0;
        sumAll = sumRun = 0;
        for (int t = 0; t < nTiles; t++)
        {
            const int c0 = t * T, c1 = std::min(nC, c0 + T);
            cur.clear();
            for (int c = c0; c < c1; c++)
            {
                loc[c] = c - c0;
                cur.push_back(c);
            }
            cum[(size_t)t * (R + 1)] = c1 - c0;
            size_t ringBegin = 0;
            for (int r = 1; r <= R; r++)
            {
                const size_t ringEnd = cur.size();
                for (size_t l = ringBegin; l < ringEnd; l++)
                {
                    const int g = cur[l];
                    if (g >= nC) continue; // a ghost cell (several ranks) has no table row
                    for (int k = 0; k < maxCF; k++)
                    {
                        const int n = hm.cellNbr[(size_t)k * nC + g];
                        if (n >= 0 && loc[n] < 0)
                        {
                            loc[n] = (int32_t)cur.size();
                            cur.push_back(n);
                        }
                    }
                }
                ringBegin = ringEnd;
                cum[(size_t)t * (R + 1) + r] = (int32_t)cur.size();
            }
            const int nRun = cum[(size_t)t * (R + 1) + R - 1], nAll = (int)cur.size();
            maxRun = std::max(maxRun, nRun);
            maxAll = std::max(maxAll, nAll);
            sumAll += nAll;
            sumRun += nRun;
            if (tables) lists[t] = cur;
            for (int g : cur) loc[g] = -1;
        }
        ln = (maxRun + 31) / 32 * 32;
        ls = (maxAll + 31) / 32 * 32;
        if (!tables) return;
        gid.assign((size_t)nTiles * ls, -1);
        tf.assign((size_t)nTiles * maxCF * ln, -1);
        tn.assign((size_t)nTiles * maxCF * ln, -1);
        for (int t = 0; t < nTiles; t++)
        {
            const std::vector<int32_t>& L = lists[t];
            for (size_t l = 0; l < L.size(); l++)
            {
                gid[(size_t)t * ls + l] = L[l];
                loc[L[l]] = (int32_t)l;
            }
            const int nRun = cum[(size_t)t * (R + 1) + R - 1];
            for (int l = 0; l < nRun; l++)
            {
                const int g = L[l];
                if (g >= nC) continue;
                for (int k = 0; k < maxCF; k++)
                {
                    const size_t o = ((size_t)t * maxCF + k) * ln + l;
                    tf[o] = hm.cellFaces[(size_t)k * nC + g];
                    const int n = hm.cellNbr[(size_t)k * nC + g];
                    tn[o] = n < 0 ? -1 : loc[n];
                }
            }
            for (int g : L) loc[g] = -1;
        }
    }
};

// device view of a tile map
struct TileView
{
    int T, nTiles, R, maxCF, ln, ls;
    const int32_t *cum, *gid, *tf, *tn;
};

} // namespace dab
