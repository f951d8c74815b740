// CPU stepping harness for open_pcc_metric_b200/csrc/pccm_core.cuh -- TEST INFRASTRUCTURE.
// Compiles the __host__ __device__ search / accumulator / eigen-solver code with g++ and
// runs it query by query over an index built with std::stable_sort, so that the exact
// logic the sm_100a kernels execute per thread can be checked against the oracle in the
// CPU-only test tier.  It is never loaded by the product package.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../open_pcc_metric_b200/csrc/pccm_core.cuh"

using namespace pccm;

static uint32_t g_short_row = 0;
#ifdef PCCM_COUNT
namespace pccm { Counters g_cnt; }
extern "C" void emul_counters(unsigned long long* out, int reset) { memcpy(out, &pccm::g_cnt, sizeof(pccm::g_cnt)); if (reset) memset(&pccm::g_cnt, 0, sizeof(pccm::g_cnt)); }
#endif

template <class K> struct Pack;
template <> struct Pack<KInt> {
    static uint4 make(const double* p, uint32_t idx) {
        uint4 r; r.x = (uint32_t)(int)p[0] | ((uint32_t)(int)p[1] << 16); r.y = (uint32_t)(int)p[2]; r.z = idx; r.w = 0; return r;
    }
};
template <> struct Pack<KF32> {
    static float4 make(const double* p, uint32_t idx) {
        float4 r; r.x = (float)p[0]; r.y = (float)p[1]; r.z = (float)p[2]; memcpy(&r.w, &idx, 4); return r;
    }
};
template <> struct Pack<KF64> {
    static RecF64 make(const double* p, uint32_t idx) { RecF64 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; r.idx = idx; return r; }
};

template <class K>
struct Index {
    RowGrid g;
    std::vector<typename K::Rec> recs;
    std::vector<uint32_t> row_start;
    uint32_t short_row = 0;
    void build(const double* pts, int64_t n, double cell) {
        double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int64_t i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], pts[3 * i + a]); mx[a] = std::max(mx[a], pts[3 * i + a]); }
        memset(&g, 0, sizeof g);
        g.n = (uint32_t)n;
        g.short_row = short_row;
        if (n == 0) { g.ny = g.nz = 1; g.h = g.inv_h = 1; row_start.assign(2, 0); return; }
        if (K::kind == KIND_INT) {
            int shift = (int)std::lround(std::log2(std::max(cell, 1.0)));
            g.shift = shift;
            g.iy0 = ((int)mn[1] >> shift) << shift;
            g.iz0 = ((int)mn[2] >> shift) << shift;
            g.ny = (((int)mx[1] - g.iy0) >> shift) + 1;
            g.nz = (((int)mx[2] - g.iz0) >> shift) + 1;
            g.h = (double)(1 << shift); g.inv_h = 1.0 / g.h; g.y0 = g.iy0; g.z0 = g.iz0;
        } else {
            double h = cell;
            g.ny = (int)std::floor((mx[1] - mn[1]) / h) + 1;
            g.nz = (int)std::floor((mx[2] - mn[2]) / h) + 1;
            g.h = h; g.inv_h = 1.0 / h; g.y0 = mn[1]; g.z0 = mn[2];
            double mag = 0;
            for (int a = 1; a < 3; ++a) mag = std::max(mag, std::max(std::fabs(mn[a]), std::fabs(mx[a])));
            g.slack = 1e-9 * (mag + h);
        }
        std::vector<uint32_t> order(n), row(n);
        for (int64_t i = 0; i < n; ++i) {
            order[i] = (uint32_t)i;
            int cy, cz;
            if (K::kind == KIND_INT) { cy = ((int)pts[3 * i + 1] - g.iy0) >> g.shift; cz = ((int)pts[3 * i + 2] - g.iz0) >> g.shift; }
            else {
                cy = (int)std::floor((pts[3 * i + 1] - g.y0) * g.inv_h); cz = (int)std::floor((pts[3 * i + 2] - g.z0) * g.inv_h);
                cy = std::min(std::max(cy, 0), g.ny - 1); cz = std::min(std::max(cz, 0), g.nz - 1);
            }
            row[i] = (uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy;
        }
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            if (row[a] != row[b]) return row[a] < row[b];
            double xa = pts[3 * a], xb = pts[3 * b];
            if (K::kind == KIND_F32) { xa = (float)xa; xb = (float)xb; }
            return xa < xb;
        });
        const size_t nrows = (size_t)g.ny * g.nz;
        row_start.assign(nrows + 1, 0);
        for (int64_t i = 0; i < n; ++i) row_start[row[i] + 1]++;
        for (size_t r = 0; r < nrows; ++r) row_start[r + 1] += row_start[r];
        recs.resize(n);
        for (int64_t i = 0; i < n; ++i) recs[i] = Pack<K>::make(pts + 3 * order[i], order[i]);
    }
};

template <class K>
static void nn_impl(const double* q, int64_t nq, const double* s, int64_t ns, double cell, int32_t* idx, double* d2) {
    Index<K> ix;
    ix.short_row = g_short_row;
    ix.build(s, ns, cell);
    for (int64_t i = 0; i < nq; ++i) {
        typename K::Q qq;
        qq.x = (typename K::C)q[3 * i]; qq.y = (typename K::C)q[3 * i + 1]; qq.z = (typename K::C)q[3 * i + 2];
        Best1<K> best;
        best.init();
        search<K>(ix.g, ix.row_start.data(), ix.recs.data(), qq, best);
        idx[i] = (int32_t)best.idx;
        d2[i] = K::d2_as_double(best.d2);
    }
}

template <class K>
static void knn_impl(const double* pts, int64_t n, int k, double cell, int32_t* idx, double* d2, double* normals) {
    Index<K> ix;
    ix.short_row = g_short_row;
    ix.build(pts, n, cell);
    std::vector<typename K::D> d2s(k);
    std::vector<uint32_t> idxs(k), poss(k);
    for (int64_t t = 0; t < n; ++t) {
        const typename K::Rec qr = ix.recs[t];
        const typename K::Q q = K::rec_q(qr);
        const uint32_t qidx = K::rec_idx(qr);
        TopK<K> acc;
        acc.init(d2s.data(), idxs.data(), poss.data(), 1, k);
        search<K>(ix.g, ix.row_start.data(), ix.recs.data(), q, acc);
        for (int j = 0; j < k; ++j) {
            idx[(size_t)qidx * k + j] = j < acc.count ? (int32_t)idxs[j] : -1;
            d2[(size_t)qidx * k + j] = j < acc.count ? K::d2_as_double(d2s[j]) : INFINITY;
        }
        if (normals) {
            double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < acc.count; ++j) {
                const typename K::Q nq = K::rec_q(ix.recs[poss[j]]);
                cumulant_add(cum, (double)nq.x, (double)nq.y, (double)nq.z);
            }
            normal_from_cumulants(cum, acc.count, normals + 3 * (size_t)qidx);
        }
    }
}

extern "C" void emul_set_short_row(uint32_t v) { g_short_row = v; }
extern "C" int emul_nn(int kind, const double* q, int64_t nq, const double* s, int64_t ns, double cell, int32_t* idx, double* d2) {
    if (kind == KIND_INT) nn_impl<KInt>(q, nq, s, ns, cell, idx, d2);
    else if (kind == KIND_F32) nn_impl<KF32>(q, nq, s, ns, cell, idx, d2);
    else nn_impl<KF64>(q, nq, s, ns, cell, idx, d2);
    return 0;
}
extern "C" int emul_knn_self(int kind, const double* pts, int64_t n, int k, double cell, int32_t* idx, double* d2, double* normals) {
    if (kind == KIND_INT) knn_impl<KInt>(pts, n, k, cell, idx, d2, normals);
    else if (kind == KIND_F32) knn_impl<KF32>(pts, n, k, cell, idx, d2, normals);
    else knn_impl<KF64>(pts, n, k, cell, idx, d2, normals);
    return 0;
}

// ------------------------------------------------------------------------------------------
// occupancy-brick path (pccm_vox.cuh): the same per-point build functions and per-query search
// functions the sm_100a kernels call, stepped sequentially in kernel order.
// ------------------------------------------------------------------------------------------
#include "../../open_pcc_metric_b200/csrc/pccm_vox.cuh"

struct VoxPair {
    std::vector<uint32_t> dirbits, dirpre, prank;
    std::vector<uint2> rows, vxyz, vkey;
    VoxView view[2];
};

static bool vox_build(const double* pts[2], const int64_t n[2], VoxPair& V) {
    VoxDims g[2];
    uint32_t ndirw[2], dir_off[2];
    for (int c = 0; c < 2; ++c) {
        int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {0, 0, 0};
        for (int64_t i = 0; i < n[c]; ++i)
            for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], (int)pts[c][3 * i + a]); mx[a] = std::max(mx[a], (int)pts[c][3 * i + a]); }
        g[c].obx = mn[0] >> 5; g[c].oby = mn[1] >> 3; g[c].obz = mn[2] >> 3;
        g[c].nbx = (mx[0] >> 5) - g[c].obx + 1; g[c].nby = (mx[1] >> 3) - g[c].oby + 1; g[c].nbz = (mx[2] >> 3) - g[c].obz + 1;
        const uint64_t bits = (uint64_t)g[c].nbx * g[c].nby * g[c].nbz;
        if (bits > (1ull << 30)) return false;       // the library keeps such pairs on the pencil path
        ndirw[c] = (uint32_t)((bits + 31) / 32);
    }
    dir_off[0] = 0; dir_off[1] = ndirw[0];
    const uint32_t nw = ndirw[0] + ndirw[1], n_total = (uint32_t)(n[0] + n[1]);
    V.dirbits.assign(nw, 0); V.dirpre.assign(nw + 1, 0);
    auto coords = [&](int c, int64_t i, int& x, int& y, int& z) { x = (int)pts[c][3 * i]; y = (int)pts[c][3 * i + 1]; z = (int)pts[c][3 * i + 2]; };
    for (int c = 0; c < 2; ++c)                                   // vx_mark_kernel
        for (int64_t i = 0; i < n[c]; ++i) { int x, y, z; coords(c, i, x, y, z); vx_mark_point(V.dirbits.data() + dir_off[c], vx_key(g[c], x, y, z)); }
    uint32_t run = 0;                                             // vx_dirsum_kernel + vx_dirscan_kernel
    for (uint32_t w = 0; w < nw; ++w) { V.dirpre[w] = run; run += (uint32_t)vx_popc(V.dirbits[w]); }
    V.dirpre[nw] = run;
    const uint32_t nblk0 = V.dirpre[ndirw[0]], nblk = run;
    V.rows.assign((size_t)(nblk + 1) * kVxRows, uint2{0u, 0u});
    V.vxyz.resize(n_total); V.vkey.resize(n_total);
    memset(V.vkey.data(), 0xff, (size_t)n_total * sizeof(uint2));
    V.prank.assign(n_total, kVxNone);
    for (int c = 0; c < 2; ++c)                                   // vx_fill_kernel
        for (int64_t i = 0; i < n[c]; ++i) {
            int x, y, z; coords(c, i, x, y, z);
            vx_fill_point(V.rows.data(), vx_slot_of_key(V.dirbits.data() + dir_off[c], V.dirpre.data() + dir_off[c], vx_key(g[c], x, y, z)), x, y, z);
        }
    run = 0;                                                      // vx_bricksum_kernel + vx_rowbase_kernel (the brick past the last one is empty)
    for (uint32_t s = 0; s <= nblk; ++s)
        for (int r = 0; r < kVxRows; ++r) { V.rows[(size_t)s * kVxRows + r].y = run; run += (uint32_t)vx_popc(V.rows[(size_t)s * kVxRows + r].x); }
    for (int c = 0; c < 2; ++c) {                                  // vx_place_kernel; arrival order of the atomics: odd indices
        std::vector<int64_t> order;                                 // downwards, then even ones upwards
        for (int64_t i = n[c]; i-- > 0;) if (i & 1) order.push_back(i);
        for (int64_t i = 0; i < n[c]; ++i) if (!(i & 1)) order.push_back(i);
        for (int64_t i : order) {
            int x, y, z; coords(c, i, x, y, z);
            const uint32_t slot = vx_slot_of_key(V.dirbits.data() + dir_off[c], V.dirpre.data() + dir_off[c], vx_key(g[c], x, y, z));
            V.prank[(c ? n[0] : 0) + i] = vx_place_point(V.rows.data(), V.vxyz.data(), V.vkey.data(), slot, x, y, z, 0u, (uint32_t)i);
        }
    }
    for (int c = 0; c < 2; ++c) {
        VoxView& W = V.view[c];
        W.g = g[c]; W.dirbits = V.dirbits.data() + dir_off[c]; W.dirpre = V.dirpre.data() + dir_off[c];
        W.rows = V.rows.data(); W.vxyz = V.vxyz.data(); W.vkey = V.vkey.data();
        W.prank = V.prank.data() + (c ? n[0] : 0);
        W.slot0 = c ? nblk0 : 0; W.nblk = c ? nblk - nblk0 : nblk0; W.n = (uint32_t)n[c];
        W.nblk_total = nblk; W.n_total = n_total;
    }
    return true;
}

// the staged window of one query brick: occupancy rows + rank bases of the centre column
struct VoxWindow {
    int sslot[27];
    bool any_brick;
    uint2 win[kVxRegRows];
    uint32_t rb[kVxRegRows];
};
static void vox_stage(const VoxView& S, int bx, int by, int bz, VoxWindow& W) {
    W.any_brick = false;
    for (int l = 0; l < 27; ++l) { W.sslot[l] = vx_slot(S, bx + l % 3 - 1, by + (l / 3) % 3 - 1, bz + l / 9 - 1); W.any_brick |= W.sslot[l] >= 0; }
    for (int r = 0; r < kVxRegRows; ++r) W.win[r] = vx_stage_row(S, W.sslot, r, W.rb[r]);
}

// stats: [0] decided by the 27-neighbourhood, [1] by the 125-neighbourhood, [2] undecided, [3] tail,
// [4] left to the pencil search.  Returns 1 when the brick grid exceeds the directory budget.
extern "C" int emul_vox_nn(const double* q, int64_t nq, const double* s, int64_t ns, int max_ring, int32_t* idx, double* d2, int64_t* stats) {
    const double* pts[2] = {q, s};
    const int64_t n[2] = {nq, ns};
    VoxPair V;
    if (!vox_build(pts, n, V)) return 1;
    Index<KInt> pencil;
    pencil.build(s, ns, 2.0);
    const VoxView& Q = V.view[0];
    const VoxView& S = V.view[1];
    for (int k = 0; k < 5; ++k) stats[k] = 0;
    for (int64_t i = 0; i < nq; ++i) { idx[i] = -2; d2[i] = -1; }
    std::vector<uint32_t> todo;
    // answer of the voxel at ranked position t (vres[]); every point of the voxel reads it through prank (epilogue)
    std::vector<int64_t> res_idx(V.vxyz.size(), -2), res_d2(V.vxyz.size(), -1);
    auto assign = [&](uint32_t t, uint32_t nidx, uint32_t nd2) { res_idx[t] = nidx; res_d2[t] = nd2; };
    // vx_search_kernel: one "warp" per query brick
    for (uint32_t lb = 0; lb < Q.nblk; ++lb) {
        const uint32_t slot = Q.slot0 + lb, t0 = vx_brick_begin(Q, slot), t1 = vx_brick_begin(Q, slot + 1);
        if (t0 >= t1) return -18;                        // every brick of the directory holds a voxel
        const uint2 first = Q.vxyz[t0];
        const int bx = (int)(first.x & 0xffffu) >> 5, by = (int)(first.x >> 16) >> 3, bz = (int)first.y >> 3;
        VoxWindow W;
        vox_stage(S, bx, by, bz, W);
        for (uint32_t t = t0; t < t1; ++t) {
            const uint2 qr = Q.vxyz[t];
            const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
            const int lx = qx & 31, ly = (qy & 7) + 2, lz = (qz & 7) + 2;
            if ((qx >> 5) != bx || (qy >> 3) != by || (qz >> 3) != bz) return -19;
            VxPick pk;
            uint32_t bd2 = kVxNone;
            int stage = -1;
            if (W.any_brick) {
                const uint32_t cand = vx_level27(vx_nb27(W.win, lx, ly, lz), bd2);
                if (cand) { vx_pick27(S, W.sslot, W.win, W.rb, lx, ly, lz, cand, pk); stage = 0; }
                else {
                    bd2 = vx_search125(S, W.sslot, W.win, W.rb, lx, ly, lz, pk);
                    if (bd2 < 9u) stage = 1;
                }
            }
            if (stage >= 0) {
                assign(t, pk.idx, bd2);
                if ((uint32_t)(pk.ex * pk.ex + pk.ey * pk.ey + pk.ez * pk.ez) != bd2) return -11;
                if (pk.idx >= (uint32_t)ns) return -12;
                if ((int)s[3 * pk.idx] != qx - pk.ex || (int)s[3 * pk.idx + 1] != qy - pk.ey || (int)s[3 * pk.idx + 2] != qz - pk.ez) return -12;
                stats[stage]++;
            } else {
                todo.push_back(t);
            }
        }
    }
    // vx_general_kernel: undecided voxels
    stats[2] = (int64_t)todo.size();
    for (uint32_t t : todo) {
        const uint2 qr = Q.vxyz[t];
        VxHit h;
        // the search kernel first scans the 27 neighbour bricks (ring 1), vx_general_kernel rings 0..2
        if (vx_search_general(S, (int)(qr.x & 0xffffu), (int)(qr.x >> 16), (int)qr.y, h, max_ring < 1 ? max_ring : 1) ||
            vx_search_general(S, (int)(qr.x & 0xffffu), (int)(qr.x >> 16), (int)qr.y, h, max_ring)) {
            if (S.vkey[h.rank].y != h.idx) return -13;
            assign(t, h.idx, h.d2);
        } else {                                    // vx_far_kernel
            KInt::Q qq; qq.x = (int)(qr.x & 0xffffu); qq.y = (int)(qr.x >> 16); qq.z = (int)qr.y;
            Best1<KInt> best;
            best.init();
            search<KInt>(pencil.g, pencil.row_start.data(), pencil.recs.data(), qq, best);
            assign(t, best.idx, best.d2);
            stats[4]++;
        }
    }
    // vx_epilogue_kernel: original order
    for (int64_t i = 0; i < nq; ++i) {
        const uint32_t t = Q.prank[i];
        const uint2 qr = Q.vxyz[t];
        if ((int)(qr.x & 0xffffu) != (int)q[3 * i] || (int)(qr.x >> 16) != (int)q[3 * i + 1] || (int)qr.y != (int)q[3 * i + 2]) return -15;
        if (Q.vkey[t].y > (uint32_t)i) return -16;               // the record holds the smallest index of its voxel
        if (Q.vkey[t].y != (uint32_t)i) stats[3]++;
        idx[i] = (int32_t)res_idx[t]; d2[i] = (double)res_d2[t];
    }
    if ((uint32_t)stats[3] != Q.n - vx_ndistinct(Q)) return -17;
    for (int64_t i = 0; i < nq; ++i) if (idx[i] == -2) return -14;   // every query exactly once
    return 0;
}

// Boundary distances on the brick index (vx_selfnn_kernel): distance of every point to its nearest OTHER
// point from the 26 / 124 voxels around it on the staged rows; voxels that hold several points answer 0.
// out[i] = squared distance, -1 where the staged rows cannot decide (the kernel scans whole bricks there,
// the library falls back to the pencil k-NN beyond 8 voxels).  Returns the number of undecided points.
extern "C" int64_t emul_vox_self(const double* pts_in, int64_t n_in, double* out) {
    const double* pts[2] = {pts_in, pts_in};
    const int64_t n[2] = {n_in, 1};          // second cloud: one point, unused
    VoxPair V;
    if (!vox_build(pts, n, V)) return -1;
    const VoxView& Q = V.view[0];
    const uint32_t nd = vx_ndistinct(Q);
    std::vector<uint8_t> dup(nd, 0);
    for (int64_t i = 0; i < n_in; ++i)       // vx_dupflag_kernel
        if (Q.vkey[Q.prank[i]].y != (uint32_t)i) dup[Q.prank[i]] = 1;
    std::vector<int64_t> vself(nd, -1);
    for (uint32_t lb = 0; lb < Q.nblk; ++lb) {
        const uint32_t slot = Q.slot0 + lb, t0 = vx_brick_begin(Q, slot), t1 = vx_brick_begin(Q, slot + 1);
        const uint2 first = Q.vxyz[t0];
        const int bx = (int)(first.x & 0xffffu) >> 5, by = (int)(first.x >> 16) >> 3, bz = (int)first.y >> 3;
        VoxWindow W;
        vox_stage(Q, bx, by, bz, W);
        for (uint32_t t = t0; t < t1; ++t) {
            if (dup[t]) { vself[t] = 0; continue; }
            const uint2 qr = Q.vxyz[t];
            const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
            const int lx = qx & 31, ly = (qy & 7) + 2, lz = (qz & 7) + 2;
            uint32_t bd2 = vx_self27(vx_nb27(W.win, lx, ly, lz));
            if (bd2 == kVxNone) bd2 = vx_self125(W.win, lx, ly, lz);
            if (bd2 < 9u) vself[t] = bd2;
        }
    }
    int64_t undecided = 0;
    for (int64_t i = 0; i < n_in; ++i) {     // vx_selfout_kernel
        out[i] = (double)vself[Q.prank[i]];
        if (out[i] < 0) ++undecided;
    }
    return undecided;
}
