// pccm_core.cuh -- per-query search logic of the pencil-grid nearest-neighbour index.
//
// Everything here is __host__ __device__ so that the very same code that the
// sm_100a kernels run per thread can be stepped on the CPU by the logic tests
// (tests/emul/, test infrastructure only -- the product has no CPU path).
//
// Index ("pencil grid"): points are sorted by (row, x) where row = cz * ny + cy is
// the cell of the point in a uniform 2-D grid over (y, z) with cell size h, and x is
// the full-resolution coordinate.  A row is therefore a pencil of h x h cross-section
// running along x, stored as one contiguous x-sorted run of records;
// row_start[row] .. row_start[row + 1] delimits it.  A query walks pencils in
// Chebyshev rings around its own cell; inside a pencil it binary-searches its own x
// and sweeps outwards, stopping as soon as dx^2 + (distance to the pencil)^2 exceeds
// the current worst accepted distance.  Rings stop when the unvisited region is
// provably farther than the current answer, so results are exact.
//
// Replaces nanoflann's KD-tree descent behind Open3D's
// KDTreeFlann::search_knn_vector_3d (reference call site cloud_pair.py:22).
// Tie rule: smaller ORIGINAL index wins among equal squared distances.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PCCM_HD __host__ __device__ __forceinline__
#define PCCM_HDN __host__ __device__
#else
#define PCCM_HD inline
#define PCCM_HDN
#include <string.h>
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
#endif

namespace pccm {

#if defined(PCCM_COUNT) && !defined(__CUDA_ARCH__)
struct Counters { unsigned long long pencils_checked, pencils_visited, bsearch_steps, cands, rings, queries, ring0_cands, skipped_by_bound, offers, inserts, shifts; };
extern Counters g_cnt;
#define PCCM_CNT(x) (x)
#else
#define PCCM_CNT(x) ((void)0)
#endif

// ---- separately rounded double arithmetic (never contracted into FMA) ------------
PCCM_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PCCM_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
PCCM_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}

enum Kind : int { KIND_INT = 0, KIND_F32 = 1, KIND_F64 = 2 };

// Grid parameters of one indexed cloud (host fills, kernels read by value).
struct RowGrid {
    int32_t ny, nz;       // table dimensions (rows = ny * nz)
    int32_t shift;        // KIND_INT: h = 1 << shift
    int32_t iy0, iz0;     // KIND_INT: origin (multiples of h)
    double y0, z0;        // float kinds: origin
    double h, inv_h;      // float kinds: cell size
    double slack;         // float kinds: absolute safety margin on cell boundaries
    uint32_t n;           // number of points
    uint32_t short_row;   // rows up to this length are scanned linearly (no binary search)
};

struct alignas(16) RecF64 {
    double x, y, z;
    unsigned long long idx;
};

// ---- coordinate kinds -----------------------------------------------------------
// KInt: integer coordinates in [0, 32767].  Record = uint4 {x | y << 16, z, idx, rgba}.
struct KInt {
    typedef uint4 Rec;
    typedef int32_t C;    // coordinate / gap type
    typedef uint32_t D;   // squared distance type (max 3 * 32767^2 < 2^32)
    struct Q { C x, y, z; };
    static constexpr int kind = KIND_INT;
    static PCCM_HD D inf() { return 0xFFFFFFFFu; }
    static PCCM_HD C rec_x(const Rec* r) { return (C)(reinterpret_cast<const uint16_t*>(r)[0]); }
    static PCCM_HD uint32_t rec_idx(const Rec& r) { return r.z; }
    static PCCM_HD Q rec_q(const Rec& r) { Q q; q.x = (C)(r.x & 0xffffu); q.y = (C)(r.x >> 16); q.z = (C)(r.y & 0xffffu); return q; }
    static PCCM_HD D sq(C g) { return (D)(g * g); }
    static PCCM_HD D dist2(const Q& q, const Rec& r) {
        C dx = q.x - (C)(r.x & 0xffffu), dy = q.y - (C)(r.x >> 16), dz = q.z - (C)(r.y & 0xffffu);
        return (D)(dx * dx) + (D)(dy * dy) + (D)(dz * dz);
    }
    // lower bound of the squared distance given dx and the pencil bound B2
    static PCCM_HD D lbound(C dx, D B2) { return (D)(dx * dx) + B2; }
    static PCCM_HD int cell_y(const RowGrid& g, C v) { int c = (v - g.iy0) >> g.shift; return c < -1 ? -1 : (c > g.ny ? g.ny : c); }
    static PCCM_HD int cell_z(const RowGrid& g, C v) { int c = (v - g.iz0) >> g.shift; return c < -1 ? -1 : (c > g.nz ? g.nz : c); }
    // distance from v to the nearest coordinate of any cell with index >= c / <= c
    static PCCM_HD C gap_up_y(const RowGrid& g, C v, int c) { C d = (g.iy0 + (c << g.shift)) - v; return d > 0 ? d : 0; }
    static PCCM_HD C gap_dn_y(const RowGrid& g, C v, int c) { C d = v - (g.iy0 + ((c + 1) << g.shift) - 1); return d > 0 ? d : 0; }
    static PCCM_HD C gap_up_z(const RowGrid& g, C v, int c) { C d = (g.iz0 + (c << g.shift)) - v; return d > 0 ? d : 0; }
    static PCCM_HD C gap_dn_z(const RowGrid& g, C v, int c) { C d = v - (g.iz0 + ((c + 1) << g.shift) - 1); return d > 0 ? d : 0; }
    static PCCM_HD C gap_inf() { return 0x7fffffff; }
    // gaps never exceed 32767 + h, so m * m fits; worst == inf() (nothing found) is never exceeded
    static PCCM_HD bool gap_sq_gt(C m, D worst) { return (D)(m * m) > worst; }
    static PCCM_HD double d2_as_double(D d) { return (double)d; }
};

// Float kinds share geometry; they differ in the record only.
template <class RecT>
struct KFloatBase {
    typedef double C;
    typedef double D;
    struct Q { C x, y, z; };
    static PCCM_HD D inf() { return INFINITY; }
    static PCCM_HD D sq(C g) { return g * g; }
    static PCCM_HD D lbound(C dx, D B2) { return (dx * dx + B2) * 0.999999999999; }
    static PCCM_HD int cell_of(double v, double o, double inv_h, int n) {
        double c = floor((v - o) * inv_h);
        return c < -1.0 ? -1 : (c > (double)n ? n : (int)c);
    }
    static PCCM_HD int cell_y(const RowGrid& g, C v) { return cell_of(v, g.y0, g.inv_h, g.ny); }
    static PCCM_HD int cell_z(const RowGrid& g, C v) { return cell_of(v, g.z0, g.inv_h, g.nz); }
    static PCCM_HD C pos(C d) { return d > 0 ? d : 0; }
    static PCCM_HD C gap_up_y(const RowGrid& g, C v, int c) { return pos((g.y0 + c * g.h) - v - g.slack); }
    static PCCM_HD C gap_dn_y(const RowGrid& g, C v, int c) { return pos(v - (g.y0 + (c + 1) * g.h) - g.slack); }
    static PCCM_HD C gap_up_z(const RowGrid& g, C v, int c) { return pos((g.z0 + c * g.h) - v - g.slack); }
    static PCCM_HD C gap_dn_z(const RowGrid& g, C v, int c) { return pos(v - (g.z0 + (c + 1) * g.h) - g.slack); }
    static PCCM_HD C gap_inf() { return INFINITY; }
    static PCCM_HD bool gap_sq_gt(C m, D worst) { return m * m * 0.999999999999 > worst; }
    static PCCM_HD double d2_as_double(D d) { return d; }
    // nanoflann L2 order: ((dx*dx) + dy*dy) + dz*dz, each operation rounded separately
    static PCCM_HD D dist2_xyz(const Q& q, double px, double py, double pz) {
        double dx = dsub(q.x, px), dy = dsub(q.y, py), dz = dsub(q.z, pz);
        return dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
    }
};

// KF32: float32-representable coordinates.  Record = float4 {x, y, z, idx bits}.
struct KF32 : KFloatBase<float4> {
    typedef float4 Rec;
    static constexpr int kind = KIND_F32;
    static PCCM_HD C rec_x(const Rec* r) { return (C)(reinterpret_cast<const float*>(r)[0]); }
    static PCCM_HD uint32_t rec_idx(const Rec& r) {
#if defined(__CUDA_ARCH__)
        return __float_as_uint(r.w);
#else
        uint32_t u; memcpy(&u, &r.w, 4); return u;
#endif
    }
    static PCCM_HD Q rec_q(const Rec& r) { Q q; q.x = r.x; q.y = r.y; q.z = r.z; return q; }
    static PCCM_HD D dist2(const Q& q, const Rec& r) { return dist2_xyz(q, (double)r.x, (double)r.y, (double)r.z); }
};

// KF64: arbitrary float64 coordinates.  Record = {x, y, z, idx} (32 B).
struct KF64 : KFloatBase<RecF64> {
    typedef RecF64 Rec;
    static constexpr int kind = KIND_F64;
    static PCCM_HD C rec_x(const Rec* r) { return r->x; }
    static PCCM_HD uint32_t rec_idx(const Rec& r) { return (uint32_t)r.idx; }
    static PCCM_HD Q rec_q(const Rec& r) { Q q; q.x = r.x; q.y = r.y; q.z = r.z; return q; }
    static PCCM_HD D dist2(const Q& q, const Rec& r) { return dist2_xyz(q, r.x, r.y, r.z); }
};

// ---- record loads (read-only path on the device) -----------------------------------
PCCM_HD uint4 load_rec(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
PCCM_HD float4 load_rec(const float4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
PCCM_HD RecF64 load_rec(const RecF64* p) {
#if defined(__CUDA_ARCH__)
    const double2* d = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(d), b = __ldg(d + 1);
    RecF64 r; r.x = a.x; r.y = a.y; r.z = b.x; r.idx = (unsigned long long)__double_as_longlong(b.y);
    return r;
#else
    return *p;
#endif
}
PCCM_HD uint32_t load_u32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- accumulators -------------------------------------------------------------------
// Best-1 accumulator.
template <class K>
struct Best1 {
    typename K::D d2;
    uint32_t idx;   // original index of the best point
    uint32_t pos;   // its position in the sorted record array
    PCCM_HD void init() { d2 = K::inf(); idx = 0xFFFFFFFFu; pos = 0xFFFFFFFFu; }
    PCCM_HD typename K::D worst() const { return d2; }
    PCCM_HD void offer(typename K::D d, uint32_t i, uint32_t p) {
        if (d < d2 || (d == d2 && i < idx)) { d2 = d; idx = i; pos = p; }
    }
};

// Two smallest squared distances, values only (registers).  Enough for Open3D's
// ComputeNearestNeighborDistance = sqrt(second entry of the 2-NN result): ties and indices do
// not influence the value.
template <class K>
struct Best2Val {
    typename K::D m1, m2;
    int count;
    PCCM_HD void init() { m1 = K::inf(); m2 = K::inf(); count = 0; }
    PCCM_HD typename K::D worst() const { return m2; }
    PCCM_HD void offer(typename K::D d, uint32_t, uint32_t) {
        ++count;
        if (d < m1) { m2 = m1; m1 = d; }
        else if (d < m2) m2 = d;
    }
};

// Top-k accumulator over caller-provided strided storage (shared memory on the
// device: element j of this thread lives at base[j * stride]).  Kept sorted
// ascending by (d2, idx).
template <class K>
struct TopK {
    typename K::D* d2s;
    uint32_t* idxs;
    uint32_t* poss;
    int stride, k, count;
    typename K::D worst_d2;
    PCCM_HD void init(typename K::D* d, uint32_t* i, uint32_t* p, int stride_, int k_) {
        d2s = d; idxs = i; poss = p; stride = stride_; k = k_; count = 0; worst_d2 = K::inf();
    }
    PCCM_HD typename K::D worst() const { return worst_d2; }
    PCCM_HD void offer(typename K::D d, uint32_t i, uint32_t p) {
        int j;
        PCCM_CNT(g_cnt.offers++);
        if (count < k) {
            j = count++;
        } else {
            const int l = (k - 1) * stride;
            if (!(d < d2s[l] || (d == d2s[l] && i < idxs[l]))) return;
            j = k - 1;
        }
        while (j > 0) {
            const int a = (j - 1) * stride;
            typename K::D pd = d2s[a];
            uint32_t pi = idxs[a];
            if (pd < d || (pd == d && pi < i)) break;
            PCCM_CNT(g_cnt.shifts++);
            d2s[j * stride] = pd; idxs[j * stride] = pi; poss[j * stride] = poss[a];
            --j;
        }
        PCCM_CNT(g_cnt.inserts++);
        d2s[j * stride] = d; idxs[j * stride] = i; poss[j * stride] = p;
        if (count == k) worst_d2 = d2s[(k - 1) * stride];
    }
};

// ---- where the rows of the pencil table come from -----------------------------------------
// GlobalRows reads the table and the records from global memory (read-only path).  The query
// kernel substitutes a staged source that serves the rows around a tile from shared memory.
template <class K>
struct GlobalRows {
    const uint32_t* row_start;
    const typename K::Rec* recs;
    int ny;
    PCCM_HD const typename K::Rec* fetch(int yy, int zz, uint32_t& lo, uint32_t& hi) const {
        const uint32_t row = (uint32_t)zz * (uint32_t)ny + (uint32_t)yy;
        lo = load_u32(row_start + row);
        hi = load_u32(row_start + row + 1);
        return recs;
    }
    static PCCM_HD typename K::Rec load(const typename K::Rec* p) { return load_rec(p); }
    static PCCM_HD typename K::C xof(const typename K::Rec* p) { return K::rec_x(p); }
};

// ---- pencil visit ---------------------------------------------------------------------
template <class K, class Rows, class Acc>
PCCM_HD void visit_run(const typename K::Rec* __restrict__ recs, uint32_t lo, uint32_t hi,
                       const typename K::Q& q, typename K::D B2, Acc& acc, uint32_t short_row) {
    if (hi - lo <= short_row) {           // short pencil: one pass, no dependent search loads
        for (uint32_t i = lo; i < hi; ++i) {
            typename K::Rec r = Rows::load(recs + i);
            typename K::C dx = K::rec_q(r).x - q.x;
            if (K::lbound(dx, B2) > acc.worst()) continue;
            acc.offer(K::dist2(q, r), K::rec_idx(r), i);
        }
        return;
    }
    // first record with x >= q.x
    uint32_t a = lo, b = hi;
    while (a < b) {
        uint32_t m = (a + b) >> 1;
        PCCM_CNT(g_cnt.bsearch_steps++);
        if (Rows::xof(recs + m) < q.x) a = m + 1; else b = m;
    }
    for (uint32_t i = a; i < hi; ++i) {   // sweep towards +x
        typename K::Rec r = Rows::load(recs + i);
        typename K::C dx = K::rec_q(r).x - q.x;
        PCCM_CNT(g_cnt.cands++);
        if (K::lbound(dx, B2) > acc.worst()) break;
        acc.offer(K::dist2(q, r), K::rec_idx(r), i);
    }
    for (uint32_t i = a; i-- > lo;) {     // sweep towards -x
        typename K::Rec r = Rows::load(recs + i);
        typename K::C dx = q.x - K::rec_q(r).x;
        PCCM_CNT(g_cnt.cands++);
        if (K::lbound(dx, B2) > acc.worst()) break;
        acc.offer(K::dist2(q, r), K::rec_idx(r), i);
    }
}

// Exact search of query q in an indexed cloud whose rows are served by `rows`.
// true when nothing outside the Chebyshev ring-r square around (cy0, cz0) can tie or beat `worst`
template <class K>
PCCM_HD bool ring_done(const RowGrid& g, const typename K::Q& q, int cy0, int cz0, int r, typename K::D worst) {
    typedef typename K::C C;
    const int ylo = cy0 - r, yhi = cy0 + r, zlo = cz0 - r, zhi = cz0 + r;
    C m = K::gap_inf();
    bool open = false;
    if (ylo > 0)         { open = true; C t = K::gap_dn_y(g, q.y, ylo - 1); m = t < m ? t : m; }
    if (yhi < g.ny - 1)  { open = true; C t = K::gap_up_y(g, q.y, yhi + 1); m = t < m ? t : m; }
    if (zlo > 0)         { open = true; C t = K::gap_dn_z(g, q.z, zlo - 1); m = t < m ? t : m; }
    if (zhi < g.nz - 1)  { open = true; C t = K::gap_up_z(g, q.z, zhi + 1); m = t < m ? t : m; }
    return !open || K::gap_sq_gt(m, worst);
}

// r_begin > 0: rings 0 .. r_begin-1 were already visited by the caller (and ring_done was false).
template <class K, class Rows, class Acc>
PCCM_HD void search_rows(const RowGrid& g, const Rows& rows, const typename K::Q& q, Acc& acc, int r_begin = 0) {
    typedef typename K::C C;
    typedef typename K::D D;
    const int ny = g.ny, nz = g.nz;
    const int qcy = K::cell_y(g, q.y), qcz = K::cell_z(g, q.z);
    const int cy0 = qcy < 0 ? 0 : (qcy >= ny ? ny - 1 : qcy);
    const int cz0 = qcz < 0 ? 0 : (qcz >= nz ? nz - 1 : qcz);

    auto bound_y = [&](int c) -> D { return c > qcy ? K::sq(K::gap_up_y(g, q.y, c)) : (c < qcy ? K::sq(K::gap_dn_y(g, q.y, c)) : (D)0); };
    auto bound_z = [&](int c) -> D { return c > qcz ? K::sq(K::gap_up_z(g, q.z, c)) : (c < qcz ? K::sq(K::gap_dn_z(g, q.z, c)) : (D)0); };
    PCCM_CNT(g_cnt.queries++);
    auto visit = [&](int yy, int zz, D B2) {
        PCCM_CNT(g_cnt.pencils_checked++);
        if (B2 > acc.worst()) { PCCM_CNT(g_cnt.skipped_by_bound++); return; }
        uint32_t lo, hi;
        const typename K::Rec* base = rows.fetch(yy, zz, lo, hi);
        if (lo < hi) { PCCM_CNT(g_cnt.pencils_visited++); visit_run<K, Rows>(base, lo, hi, q, B2, acc, g.short_row); }
    };

    for (int r = r_begin;; ++r) {
        const int ylo = cy0 - r, yhi = cy0 + r, zlo = cz0 - r, zhi = cz0 + r;
        PCCM_CNT(g_cnt.rings++);
        if (r == 0) {
            visit(cy0, cz0, bound_y(cy0) + bound_z(cz0));
        } else {
            const int ya = ylo < 0 ? 0 : ylo, yb = yhi > ny - 1 ? ny - 1 : yhi;
            const int za = zlo + 1 < 0 ? 0 : zlo + 1, zb = zhi - 1 > nz - 1 ? nz - 1 : zhi - 1;
            for (int s = 0; s < 2; ++s) {            // the two full lines z = zlo, zhi
                const int zz = s ? zhi : zlo;
                if (zz < 0 || zz >= nz) continue;
                const D bz = bound_z(zz);
                if (bz > acc.worst()) continue;
                for (int yy = ya; yy <= yb; ++yy) visit(yy, zz, bound_y(yy) + bz);
            }
            for (int s = 0; s < 2; ++s) {            // the two open columns y = ylo, yhi
                const int yy = s ? yhi : ylo;
                if (yy < 0 || yy >= ny) continue;
                const D by = bound_y(yy);
                if (by > acc.worst()) continue;
                for (int zz = za; zz <= zb; ++zz) visit(yy, zz, by + bound_z(zz));
            }
        }
        // can the unvisited region still hold an equal or better point?
        if (ring_done<K>(g, q, cy0, cz0, r, acc.worst())) break;
    }
}

// Exact search of query q in the indexed cloud (g, row_start, recs) read from global memory.
template <class K, class Acc>
PCCM_HD void search(const RowGrid& g, const uint32_t* __restrict__ row_start,
                    const typename K::Rec* __restrict__ recs, const typename K::Q& q, Acc& acc) {
    GlobalRows<K> rows;
    rows.row_start = row_start; rows.recs = recs; rows.ny = g.ny;
    search_rows<K>(g, rows, q, acc);
}

// ---- Open3D EstimateNormals arithmetic ---------------------------------------------------
// utility::ComputeCovariance + geometry::FastEigen3x3 (Open3D 0.18.0, called through
// PointCloud.estimate_normals(), reference cloud_pair.py:61-64).  cum[9] holds the raw
// sums (x, y, z, xx, xy, xz, yy, yz, zz) over cnt neighbours.
PCCM_HD void cross3(const double* a, const double* b, double* o) {
    o[0] = dsub(dmul(a[1], b[2]), dmul(a[2], b[1]));
    o[1] = dsub(dmul(a[2], b[0]), dmul(a[0], b[2]));
    o[2] = dsub(dmul(a[0], b[1]), dmul(a[1], b[0]));
}
PCCM_HD double dot3(const double* a, const double* b) {
    return dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2]));
}

PCCM_HDN inline void eigenvector0(const double A[3][3], double eval0, double* out) {
    double row0[3] = {dsub(A[0][0], eval0), A[0][1], A[0][2]};
    double row1[3] = {A[0][1], dsub(A[1][1], eval0), A[1][2]};
    double row2[3] = {A[0][2], A[1][2], dsub(A[2][2], eval0)};
    double r0xr1[3], r0xr2[3], r1xr2[3];
    cross3(row0, row1, r0xr1);
    cross3(row0, row2, r0xr2);
    cross3(row1, row2, r1xr2);
    double d0 = dot3(r0xr1, r0xr1), d1 = dot3(r0xr2, r0xr2), d2 = dot3(r1xr2, r1xr2);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const double* v = imax == 0 ? r0xr1 : (imax == 1 ? r0xr2 : r1xr2);
    double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    out[0] = v[0] / s; out[1] = v[1] / s; out[2] = v[2] / s;
}

PCCM_HDN inline void eigenvector1(const double A[3][3], const double* evec0, double eval1, double* out) {
    double U[3], V[3];
    if (fabs(evec0[0]) > fabs(evec0[1])) {
        double inv_length = 1.0 / sqrt(dadd(dmul(evec0[0], evec0[0]), dmul(evec0[2], evec0[2])));
        U[0] = dmul(-evec0[2], inv_length); U[1] = 0; U[2] = dmul(evec0[0], inv_length);
    } else {
        double inv_length = 1.0 / sqrt(dadd(dmul(evec0[1], evec0[1]), dmul(evec0[2], evec0[2])));
        U[0] = 0; U[1] = dmul(evec0[2], inv_length); U[2] = dmul(-evec0[1], inv_length);
    }
    cross3(evec0, U, V);
    double AU[3], AV[3];
    for (int i = 0; i < 3; ++i) {
        const double a0 = i == 0 ? A[0][0] : (i == 1 ? A[0][1] : A[0][2]);
        const double a1 = i == 0 ? A[0][1] : (i == 1 ? A[1][1] : A[1][2]);
        const double a2 = i == 0 ? A[0][2] : (i == 1 ? A[1][2] : A[2][2]);
        AU[i] = dadd(dadd(dmul(a0, U[0]), dmul(a1, U[1])), dmul(a2, U[2]));
        AV[i] = dadd(dadd(dmul(a0, V[0]), dmul(a1, V[1])), dmul(a2, V[2]));
    }
    double m00 = dsub(dot3(U, AU), eval1);
    double m01 = dot3(U, AV);
    double m11 = dsub(dot3(V, AV), eval1);
    double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        double mx = a00 > a01 ? a00 : a01;
        if (mx > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1.0 / sqrt(dadd(1.0, dmul(m01, m01))); m01 = dmul(m01, m00); }
            else            { m00 /= m01; m01 = 1.0 / sqrt(dadd(1.0, dmul(m00, m00))); m00 = dmul(m00, m01); }
            for (int i = 0; i < 3; ++i) out[i] = dsub(dmul(m01, U[i]), dmul(m00, V[i]));
        } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
    } else {
        double mx = a11 > a01 ? a11 : a01;
        if (mx > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1.0 / sqrt(dadd(1.0, dmul(m01, m01))); m01 = dmul(m01, m11); }
            else            { m11 /= m01; m01 = 1.0 / sqrt(dadd(1.0, dmul(m11, m11))); m11 = dmul(m11, m01); }
            for (int i = 0; i < 3; ++i) out[i] = dsub(dmul(m11, U[i]), dmul(m01, V[i]));
        } else { out[0] = U[0]; out[1] = U[1]; out[2] = U[2]; }
    }
}

PCCM_HDN inline void fast_eigen_3x3(const double cov[3][3], double* out) {
    double max_coeff = cov[0][0];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (cov[i][j] > max_coeff) max_coeff = cov[i][j];
    if (max_coeff == 0) { out[0] = out[1] = out[2] = 0; return; }
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = cov[i][j] / max_coeff;
    double norm = dadd(dadd(dmul(A[0][1], A[0][1]), dmul(A[0][2], A[0][2])), dmul(A[1][2], A[1][2]));
    if (norm > 0) {
        double q = dadd(dadd(A[0][0], A[1][1]), A[2][2]) / 3.0;
        double b00 = dsub(A[0][0], q), b11 = dsub(A[1][1], q), b22 = dsub(A[2][2], q);
        double p = sqrt(dadd(dadd(dadd(dmul(b00, b00), dmul(b11, b11)), dmul(b22, b22)), dmul(norm, 2.0)) / 6.0);
        double c00 = dsub(dmul(b11, b22), dmul(A[1][2], A[1][2]));
        double c01 = dsub(dmul(A[0][1], b22), dmul(A[1][2], A[0][2]));
        double c02 = dsub(dmul(A[0][1], A[1][2]), dmul(b11, A[0][2]));
        double det = dadd(dsub(dmul(b00, c00), dmul(A[0][1], c01)), dmul(A[0][2], c02)) / dmul(dmul(p, p), p);
        double half_det = dmul(det, 0.5);
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        double angle = acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = dmul(cos(angle), 2.0);
        double beta0 = dmul(cos(dadd(angle, two_thirds_pi)), 2.0);
        double beta1 = -dadd(beta0, beta2);
        double e0 = dadd(q, dmul(p, beta0)), e1 = dadd(q, dmul(p, beta1)), e2 = dadd(q, dmul(p, beta2));
        double ev0[3], ev1[3], ev2[3];
        if (half_det >= 0) {
            eigenvector0(A, e2, ev2);
            if (e2 < e0 && e2 < e1) { out[0] = ev2[0]; out[1] = ev2[1]; out[2] = ev2[2]; return; }
            eigenvector1(A, ev2, e1, ev1);
            if (e1 < e0 && e1 < e2) { out[0] = ev1[0]; out[1] = ev1[1]; out[2] = ev1[2]; return; }
            cross3(ev1, ev2, out);
            return;
        }
        eigenvector0(A, e0, ev0);
        if (e0 < e1 && e0 < e2) { out[0] = ev0[0]; out[1] = ev0[1]; out[2] = ev0[2]; return; }
        eigenvector1(A, ev0, e1, ev1);
        if (e1 < e0 && e1 < e2) { out[0] = ev1[0]; out[1] = ev1[1]; out[2] = ev1[2]; return; }
        cross3(ev0, ev1, out);
        return;
    }
    double a00 = dmul(A[0][0], max_coeff), a11 = dmul(A[1][1], max_coeff), a22 = dmul(A[2][2], max_coeff);
    out[0] = out[1] = out[2] = 0;
    if (a00 < a11 && a00 < a22) out[0] = 1;
    else if (a11 < a00 && a11 < a22) out[1] = 1;
    else out[2] = 1;
}

// cum = raw sums over cnt neighbours -> unit normal (Open3D rules for degenerate cases).
PCCM_HDN inline void normal_from_cumulants(const double* cum, int cnt, double* nrm) {
    double cov[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (cnt >= 3) {
        double c[9];
        for (int j = 0; j < 9; ++j) c[j] = cum[j] / (double)cnt;
        cov[0][0] = dsub(c[3], dmul(c[0], c[0]));
        cov[1][1] = dsub(c[6], dmul(c[1], c[1]));
        cov[2][2] = dsub(c[8], dmul(c[2], c[2]));
        cov[0][1] = cov[1][0] = dsub(c[4], dmul(c[0], c[1]));
        cov[0][2] = cov[2][0] = dsub(c[5], dmul(c[0], c[2]));
        cov[1][2] = cov[2][1] = dsub(c[7], dmul(c[1], c[2]));
    }
    fast_eigen_3x3(cov, nrm);
    if (sqrt(dadd(dadd(dmul(nrm[0], nrm[0]), dmul(nrm[1], nrm[1])), dmul(nrm[2], nrm[2]))) == 0.0) {
        nrm[0] = 0; nrm[1] = 0; nrm[2] = 1;
    }
}

// cumulant update with one neighbour (separately rounded, sequential like the reference loop)
PCCM_HD void cumulant_add(double* c, double x, double y, double z) {
    c[0] = dadd(c[0], x); c[1] = dadd(c[1], y); c[2] = dadd(c[2], z);
    c[3] = dadd(c[3], dmul(x, x)); c[4] = dadd(c[4], dmul(x, y)); c[5] = dadd(c[5], dmul(x, z));
    c[6] = dadd(c[6], dmul(y, y)); c[7] = dadd(c[7], dmul(y, z)); c[8] = dadd(c[8], dmul(z, z));
}

// ---- colour / plane epilogue arithmetic ----------------------------------------------------
// metric.py:283-290: row-wise matmul(T, c); metric.py:329-333: (T c_q - T c_n)^2.
PCCM_HD void color_diff2(const double* T, const double* cq, const double* cn, double scale, double* d2, double* d2s) {
    for (int k = 0; k < 3; ++k) {
        double tq = dadd(dadd(dmul(T[3 * k], cq[0]), dmul(T[3 * k + 1], cq[1])), dmul(T[3 * k + 2], cq[2]));
        double tn = dadd(dadd(dmul(T[3 * k], cn[0]), dmul(T[3 * k + 1], cn[1])), dmul(T[3 * k + 2], cn[2]));
        double d = dsub(tq, tn);
        d2[k] = dmul(d, d);
        double ds = dmul(scale, d);
        d2s[k] = dmul(ds, ds);
    }
}
// metric.py:148-152: dot(E[i], normal[i]) then squared (metric.py:179)
PCCM_HD double plane_err2(const double* e, const double* n) {
    double pe = dadd(dadd(dmul(e[0], n[0]), dmul(e[1], n[1])), dmul(e[2], n[2]));
    return dmul(pe, pe);
}

}  // namespace pccm
