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
