// pccm_vox.cuh -- occupancy-brick index for voxelised (KInt) clouds and the bit-scan nearest
// neighbour search on it.  Like pccm_core.cuh everything here is __host__ __device__: the
// sm_100a kernels (pccm_vox_kernels.cuh) and the CPU stepping harness (tests/emul/, test
// infrastructure) run the very same per-point / per-query functions.
//
// Replaces, for integer clouds, the KD-tree build + per-point descent behind
// KDTreeFlann(cloud) / search_knn_vector_3d(p, 1) (reference cloud_pair.py:65 and :22).
//
// Index.  Space is cut into bricks of 32 x 8 x 8 voxels (x, y, z).  A brick is 64 "rows"
// (one per (y, z)) of one 32-bit word: bit b of row r says voxel (32 bx + b, y, z) is occupied.
//   dirbits / dirpre : bitmap over the brick grid of the cloud's bounding box + exclusive
//                      popcount prefix  ->  slot of an occupied brick (slots follow (z, y, x))
//   masks[slot][64]  : the occupancy words
//   pre[slot][64]    : number of occupied voxels of the brick before row r
//   base[slot]       : rank of the brick's first voxel; rank = base + pre + popc(bits below)
//   recs[rank]       : {x | y << 16, z, rgb, idx} of the voxel's point with the SMALLEST original
//                      index (only that one can win under the tie rule; atomicMin on the 64-bit
//                      {idx, rgb} half of the record)
//   prank[i]         : rank of the voxel of input point i.  As queries the points of a voxel
//                      share ONE search (one lane per voxel); the per-point epilogue runs in the
//                      original order of the input and fetches its voxel's answer through prank.
// Two clouds of a pair share the arrays (cloud 1's slots, ranks and points continue cloud 0's).
//
// Search.  The nearest occupied voxel of a row to the query's x is two bit scans (CLZ on the
// bits at or below x, CLZ of the bit-reversed bits above): dx, hence dx^2 + dy^2 + dz^2 for the
// whole row, with no point ever loaded.  The 3 x 3 (then 5 x 5) rows around the query decide
// every query whose answer is closer than 2 (3) voxels -- exactly, because anything outside the
// scanned rows / the +-16 window in x is at least that far.  Only the voxels that TIE at the
// minimal distance are then looked up (rank -> record) to apply the smallest-index rule.
// Queries that stay undecided go to vx_search_general (brick rings, up to a ring limit) and, beyond
// that (clouds far apart, isolated outliers), to the pencil search of pccm_core.cuh.
#pragma once
#include "pccm_core.cuh"

#if !defined(__CUDACC__)
struct uint2 { uint32_t x, y; };
#endif

// Debug builds (-DPCCM_VX_DEBUG): every index the brick path computes is range-checked; violations are
// counted (pccm_debug_errors) instead of trapping, so one run reports all of them.
#if defined(PCCM_VX_DEBUG) && defined(__CUDACC__)
namespace pccm { __device__ unsigned int g_vx_errors = 0; }      // (the library is one translation unit)
#endif
#if defined(PCCM_VX_DEBUG) && defined(__CUDA_ARCH__)
#define VX_CHECK(cond) do { if (!(cond)) atomicAdd(&pccm::g_vx_errors, 1u); } while (0)
#else
#define VX_CHECK(cond) ((void)0)
#endif

namespace pccm {

// ---- bit helpers ---------------------------------------------------------------------------
PCCM_HD int vx_clz(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
PCCM_HD uint32_t vx_brev(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
PCCM_HD int vx_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
PCCM_HD int vx_ffs(uint32_t v) {   // 1-based position of the lowest set bit, 0 when none
#if defined(__CUDA_ARCH__)
    return __ffs((int)v);
#else
    return v ? __builtin_ctz(v) + 1 : 0;
#endif
}
PCCM_HD uint32_t vx_fshr(uint32_t lo, uint32_t hi, int s) {   // low word of (hi:lo) >> (s & 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, (uint32_t)s);
#else
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
PCCM_HD uint32_t vx_atomic_or(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicOr(p, v);
#else
    const uint32_t o = *p; *p = o | v; return o;
#endif
}
PCCM_HD uint32_t vx_atomic_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const uint32_t o = *p; *p = o + v; return o;
#endif
}
PCCM_HD unsigned long long vx_atomic_min64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return atomicMin(p, v);
#else
    const unsigned long long o = *p; if (v < o) *p = v; return o;
#endif
}
PCCM_HD uint32_t vx_ld32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
PCCM_HD uint32_t vx_ld16(const uint16_t* p) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ldg(p);
#else
    return *p;
#endif
}
PCCM_HD uint2 vx_ld64(const uint2* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- geometry ------------------------------------------------------------------------------
constexpr int kVxRows = 64;          // rows (words) per brick: 8 (y) x 8 (z)
constexpr int kVxRegY = 12;          // staged region: 8 + 2 * 2 rows in y and in z
constexpr int kVxRegRows = 144;
constexpr uint32_t kVxNone = 0xFFFFFFFFu;

struct VoxDims {                     // brick grid of one cloud (brick units)
    int32_t obx, oby, obz;           // origin = brick of the bounding-box minimum
    int32_t nbx, nby, nbz;
};

PCCM_HD uint32_t vx_key(const VoxDims& g, int x, int y, int z) {   // a point of the cloud itself: always inside
    return (uint32_t)((((z >> 3) - g.obz) * g.nby + ((y >> 3) - g.oby)) * g.nbx + ((x >> 5) - g.obx));
}
PCCM_HD int vx_row(int y, int z) { return ((z & 7) << 3) | (y & 7); }

// Query-time view of one indexed cloud.  dirbits / dirpre point at this cloud's directory; the
// other arrays are the pair's joint arrays.
struct VoxView {
    VoxDims g;
    const uint32_t* dirbits;
    const uint32_t* dirpre;
    const uint32_t* masks;
    const uint16_t* pre;
    const uint32_t* base;            // [nblk_total + 1]
    const uint4* recs;               // [n_total]
    const uint32_t* prank;           // [n] this cloud's points, original order -> rank of their voxel
    uint32_t slot0, nblk;            // this cloud's bricks are slots [slot0, slot0 + nblk)
    uint32_t n;                      // points of this cloud
    uint32_t nblk_total, n_total;
};

PCCM_HD int vx_slot(const VoxView& G, int bx, int by, int bz) {
    const int ix = bx - G.g.obx, iy = by - G.g.oby, iz = bz - G.g.obz;
    if ((unsigned)ix >= (unsigned)G.g.nbx || (unsigned)iy >= (unsigned)G.g.nby || (unsigned)iz >= (unsigned)G.g.nbz) return -1;
    const uint32_t key = (uint32_t)((iz * G.g.nby + iy) * G.g.nbx + ix);
    const uint32_t w = vx_ld32(G.dirbits + (key >> 5));
    const uint32_t bit = key & 31u;
    if (!((w >> bit) & 1u)) return -1;
    return (int)(vx_ld32(G.dirpre + (key >> 5)) + (uint32_t)vx_popc(w & ((1u << bit) - 1u)));
}
PCCM_HD uint32_t vx_rank(const VoxView& G, uint32_t slot, int r, int xbit) {
    VX_CHECK(slot < G.nblk_total && (unsigned)r < (unsigned)kVxRows && (unsigned)xbit < 32u);
    const uint32_t m = vx_ld32(G.masks + (size_t)slot * kVxRows + r);
    VX_CHECK((m >> xbit) & 1u);                       // the voxel we rank must be occupied
    return vx_ld32(G.base + slot) + vx_ld16(G.pre + (size_t)slot * kVxRows + r) + (uint32_t)vx_popc(m & ((1u << xbit) - 1u));
}
// positions of the cloud's records in the joint array
PCCM_HD uint32_t vx_ranked_begin(const VoxView& G) { return vx_ld32(G.base + G.slot0); }
PCCM_HD uint32_t vx_ndistinct(const VoxView& G) { return vx_ld32(G.base + G.slot0 + G.nblk) - vx_ld32(G.base + G.slot0); }

// ---- build, per point (the kernels call these once per input point, pass after pass) --------
PCCM_HD void vx_mark_point(uint32_t* dirbits, uint32_t key) {
    const uint32_t bit = 1u << (key & 31u);
    if (!(dirbits[key >> 5] & bit)) vx_atomic_or(dirbits + (key >> 5), bit);   // a stale read only costs a redundant atomic
}
PCCM_HD uint32_t vx_slot_of_key(const uint32_t* dirbits, const uint32_t* dirpre, uint32_t key) {
    const uint32_t w = dirbits[key >> 5];
    return dirpre[key >> 5] + (uint32_t)vx_popc(w & ((1u << (key & 31u)) - 1u));
}
PCCM_HD void vx_fill_point(uint32_t* masks, uint32_t slot, int x, int y, int z) {
    uint32_t* w = masks + (size_t)slot * kVxRows + vx_row(y, z);
    const uint32_t bit = 1u << (x & 31);
    if (!(*w & bit)) vx_atomic_or(w, bit);
}
// pass 3 (after the brick prefixes): rank of the point's voxel; voxel coordinates into the record;
// the smallest original index (with its colour) wins the record's {rgb, idx} half.  Records must
// be pre-filled with 0xFF.
PCCM_HD uint32_t vx_place_point(const uint32_t* masks, const uint16_t* pre, const uint32_t* base, uint4* recs,
                                uint32_t slot, int x, int y, int z, uint32_t rgb, uint32_t idx) {
    const int r = vx_row(y, z);
    const uint32_t m = masks[(size_t)slot * kVxRows + r];
    const uint32_t rank = base[slot] + pre[(size_t)slot * kVxRows + r] + (uint32_t)vx_popc(m & ((1u << (x & 31)) - 1u));
    VX_CHECK((m >> (x & 31)) & 1u);
    recs[rank].x = (uint32_t)x | ((uint32_t)y << 16);      // every point of the voxel writes the same two words
    recs[rank].y = (uint32_t)z;
    vx_atomic_min64(reinterpret_cast<unsigned long long*>(&recs[rank].z), ((unsigned long long)idx << 32) | rgb);
    return rank;
}

// ---- staged search (rows within 2 of the query, x within 16) ---------------------------------
// Region around query brick (bx, by, bz): rows y in [8 by - 2, 8 by + 10), z likewise; row i =
// rz * 12 + ry holds the 64 bits x in [32 bx - 16, 32 bx + 48) as {lo, hi}.
PCCM_HD uint2 vx_stage_row(const VoxView& S, const int* sslot, int i) {
    const int ry = i % kVxRegY, rz = i / kVxRegY;
    const int ny_i = ry < 2 ? 0 : (ry < 10 ? 1 : 2), nz_i = rz < 2 ? 0 : (rz < 10 ? 1 : 2);
    const int r_in = (((rz + 6) & 7) << 3) | ((ry + 6) & 7);
    const int* s3 = sslot + nz_i * 9 + ny_i * 3;
    const uint32_t l = s3[0] >= 0 ? vx_ld32(S.masks + (size_t)s3[0] * kVxRows + r_in) : 0u;
    const uint32_t c = s3[1] >= 0 ? vx_ld32(S.masks + (size_t)s3[1] * kVxRows + r_in) : 0u;
    const uint32_t r = s3[2] >= 0 ? vx_ld32(S.masks + (size_t)s3[2] * kVxRows + r_in) : 0u;
    uint2 w;
    w.x = (l >> 16) | (c << 16);
    w.y = (c >> 16) | (r << 16);
    return w;
}

// distance from the query's x to the nearest occupied voxel of a row, below-or-at (dd: 0..16,
// 17 = none in the window) and at-or-above (du: 0..15, 32 = none).  lx = x & 31.
PCCM_HD void vx_row_dists(const uint2 w, int lx, int& dd, int& du) {
    const uint32_t v = vx_fshr(w.x, w.y, lx);            // bit 16 = the query's own x
    dd = vx_clz(v & 0x1FFFFu) - 15;
    du = vx_clz(vx_brev(v >> 16));
}

#define VX_ROW(DY, DZ)                                                                      \
    {                                                                                       \
        int dd, du;                                                                         \
        vx_row_dists(win[(lz + (DZ)) * kVxRegY + (ly + (DY))], lx, dd, du);                 \
        const int dx = dd < du ? dd : du;                                                   \
        const uint32_t d2 = (uint32_t)(dx * dx + ((DY) * (DY) + (DZ) * (DZ)));               \
        if (d2 < bd2) { bd2 = d2; rows = 0u; }                                              \
        if (d2 == bd2) rows |= 1u << (((DZ) + 2) * 5 + (DY) + 2);                           \
    }

// the 3 x 3 rows: decides every query with best d2 < 4.  ly, lz in [2, 10): row of the query in the region.
PCCM_HD void vx_rows_inner(const uint2* win, int lx, int ly, int lz, uint32_t& bd2, uint32_t& rows) {
    VX_ROW(0, 0)
    VX_ROW(-1, 0) VX_ROW(1, 0) VX_ROW(0, -1) VX_ROW(0, 1)
    VX_ROW(-1, -1) VX_ROW(1, -1) VX_ROW(-1, 1) VX_ROW(1, 1)
}
// the eight rows around the query's own row (self nearest neighbour: the own row is handled apart)
PCCM_HD void vx_rows_ring1(const uint2* win, int lx, int ly, int lz, uint32_t& bd2, uint32_t& rows) {
    VX_ROW(-1, 0) VX_ROW(1, 0) VX_ROW(0, -1) VX_ROW(0, 1)
    VX_ROW(-1, -1) VX_ROW(1, -1) VX_ROW(-1, 1) VX_ROW(1, 1)
}
// the 16 rows of the 5 x 5 border: with them, every query with best d2 < 9.
PCCM_HD void vx_rows_outer(const uint2* win, int lx, int ly, int lz, uint32_t& bd2, uint32_t& rows) {
    VX_ROW(-2, 0) VX_ROW(2, 0) VX_ROW(0, -2) VX_ROW(0, 2)
    VX_ROW(-2, -1) VX_ROW(2, -1) VX_ROW(-2, 1) VX_ROW(2, 1) VX_ROW(-1, -2) VX_ROW(1, -2) VX_ROW(-1, 2) VX_ROW(1, 2)
    VX_ROW(-2, -2) VX_ROW(2, -2) VX_ROW(-2, 2) VX_ROW(2, 2)
}
#undef VX_ROW

struct VxPick {           // the chosen neighbour
    uint32_t idx, rgb, rank;
    int ex, ey, ez;       // query - neighbour
};

PCCM_HD void vx_cand(const VoxView& S, const int* sslot, int bx, int by, int bz, int cx, int cy, int cz,
                     int qx, int qy, int qz, VxPick& pk) {
    const int nb = ((cz >> 3) - bz + 1) * 9 + ((cy >> 3) - by + 1) * 3 + ((cx >> 5) - bx + 1);
    VX_CHECK((unsigned)nb < 27u && sslot[nb] >= 0);
    const uint32_t rank = vx_rank(S, (uint32_t)sslot[nb], vx_row(cy, cz), cx & 31);
    VX_CHECK(rank < S.n_total);
    const uint2 a = vx_ld64(reinterpret_cast<const uint2*>(S.recs + rank) + 1);   // {rgb, idx}
    if (a.y < pk.idx) { pk.idx = a.y; pk.rgb = a.x; pk.rank = rank; pk.ex = qx - cx; pk.ey = qy - cy; pk.ez = qz - cz; }
}

// among the voxels that tie at the minimal distance (rows = the rows that reach it), the one whose
// point has the smallest original index
PCCM_HD void vx_pick(const VoxView& S, const int* sslot, const uint2* win, int bx, int by, int bz,
                     int qx, int qy, int qz, uint32_t rows, VxPick& pk) {
    const int lx = qx & 31, ly = (qy & 7) + 2, lz = (qz & 7) + 2;
    pk.idx = kVxNone;
    while (rows) {
        const int b = vx_ffs(rows) - 1;
        rows &= rows - 1u;
        const int jz = b / 5, dz = jz - 2, dy = b - jz * 5 - 2;
        int dd, du;
        vx_row_dists(win[(lz + dz) * kVxRegY + (ly + dy)], lx, dd, du);
        const int dx = dd < du ? dd : du;
        if (dd == dx) vx_cand(S, sslot, bx, by, bz, qx - dx, qy + dy, qz + dz, qx, qy, qz, pk);
        if (du == dx && dx != 0) vx_cand(S, sslot, bx, by, bz, qx + dx, qy + dy, qz + dz, qx, qy, qz, pk);
    }
}

// ---- general search: brick rings, any distance (exact) --------------------------------------
// Sequential statement of what the kernels do warp-cooperatively (vx_warp_bricks in pccm_vox_kernels.cuh:
// ring 1 inside vx_search_kernel, rings 0..2 in vx_general_kernel); the CPU stepping harness runs this one.
struct VxHit {
    uint32_t d2, idx, rgb, rank;
    int cx, cy, cz;
};

PCCM_HD void vx_offer(const VoxView& S, uint32_t slot, int r, uint32_t d2, int cx, int cy, int cz, VxHit& h) {
    if (d2 > h.d2) return;
    const uint32_t rank = vx_rank(S, slot, r, cx & 31);
    const uint2 a = vx_ld64(reinterpret_cast<const uint2*>(S.recs + rank) + 1);
    if (d2 < h.d2 || a.y < h.idx) { h.d2 = d2; h.idx = a.y; h.rgb = a.x; h.rank = rank; h.cx = cx; h.cy = cy; h.cz = cz; }
}

// distance from word coordinate p (the query's x relative to the brick: ANY integer) to the nearest set bit
// of a row at or below / strictly above it; 40000 when there is none.  Shared by the sequential brick scan
// below and the warp-cooperative one of the kernels (vx_warp_bricks).
PCCM_HD void vx_row_nearest(uint32_t m, int p, int& dlo, int& dhi) {
    const uint32_t at_or_below = p >= 31 ? 0xFFFFFFFFu : (p < 0 ? 0u : ((2u << p) - 1u));
    const uint32_t above = p < 0 ? 0xFFFFFFFFu : (p >= 31 ? 0u : ~((2u << p) - 1u));
    const uint32_t ml = m & at_or_below, mh = m & above;
    dlo = ml ? p - (31 - vx_clz(ml)) : 40000;
    dhi = mh ? (vx_ffs(mh) - 1) - p : 40000;
}

PCCM_HD void vx_scan_brick(const VoxView& S, uint32_t slot, int bx, int by, int bz, int qx, int qy, int qz, VxHit& h) {
    const int X0 = bx << 5, Y0 = by << 3, Z0 = bz << 3;
    const int p = qx - X0;
    for (int zi = 0; zi < 8; ++zi) {
        const int dz = qz - (Z0 + zi);
        const uint32_t dz2 = (uint32_t)(dz * dz);
        if (dz2 > h.d2) continue;
        for (int yi = 0; yi < 8; ++yi) {
            const int dy = qy - (Y0 + yi);
            const uint32_t byz = dz2 + (uint32_t)(dy * dy);
            if (byz > h.d2) continue;
            const int r = (zi << 3) | yi;
            const uint32_t m = vx_ld32(S.masks + (size_t)slot * kVxRows + r);
            if (!m) continue;
            int dlo, dhi;
            vx_row_nearest(m, p, dlo, dhi);
            if (dlo < 40000) vx_offer(S, slot, r, byz + (uint32_t)(dlo * dlo), qx - dlo, Y0 + yi, Z0 + zi, h);
            if (dhi < 40000) vx_offer(S, slot, r, byz + (uint32_t)(dhi * dhi), qx + dhi, Y0 + yi, Z0 + zi, h);
        }
    }
}

PCCM_HD int vx_gap(int q, int lo, int hi) { return q < lo ? lo - q : (q > hi ? q - hi : 0); }
PCCM_HD int vx_imin(int a, int b) { return a < b ? a : b; }
PCCM_HD int vx_imax(int a, int b) { return a > b ? a : b; }
PCCM_HD int vx_iabs(int a) { return a < 0 ? -a : a; }

// Returns true when the answer is certified; false when ring `max_ring` was completed without
// certification (the caller hands the query to the pencil search, which prunes empty space).
PCCM_HD bool vx_search_general(const VoxView& S, int qx, int qy, int qz, VxHit& h, int max_ring) {
    h.d2 = kVxNone; h.idx = kVxNone; h.rgb = 0; h.rank = kVxNone; h.cx = h.cy = h.cz = 0;
    const int qbx = qx >> 5, qby = qy >> 3, qbz = qz >> 3;
    const int x0 = S.g.obx, x1 = S.g.obx + S.g.nbx - 1, y0 = S.g.oby, y1 = S.g.oby + S.g.nby - 1, z0 = S.g.obz, z1 = S.g.obz + S.g.nbz - 1;
    int rmax = vx_imax(vx_iabs(qbx - x0), vx_iabs(qbx - x1));
    rmax = vx_imax(rmax, vx_imax(vx_iabs(qby - y0), vx_iabs(qby - y1)));
    rmax = vx_imax(rmax, vx_imax(vx_iabs(qbz - z0), vx_iabs(qbz - z1)));
    for (int R = 0; R <= rmax; ++R) {
        if (R > max_ring) return false;
        const int zlo = vx_imax(qbz - R, z0), zhi = vx_imin(qbz + R, z1);
        const int ylo = vx_imax(qby - R, y0), yhi = vx_imin(qby + R, y1);
        const int xlo = vx_imax(qbx - R, x0), xhi = vx_imin(qbx + R, x1);
        for (int bz = zlo; bz <= zhi; ++bz) {
            const int gz = vx_gap(qz, bz << 3, (bz << 3) + 7);
            const uint32_t gz2 = (uint32_t)(gz * gz);
            if (gz2 > h.d2) continue;
            const bool fz = vx_iabs(bz - qbz) == R;
            for (int by = ylo; by <= yhi; ++by) {
                const int gy = vx_gap(qy, by << 3, (by << 3) + 7);
                const uint32_t gyz = gz2 + (uint32_t)(gy * gy);
                if (gyz > h.d2) continue;
                const bool face = fz || vx_iabs(by - qby) == R;
                // on a z / y face of the shell every x of the ring; elsewhere only the two x faces
                const int step = face ? 1 : (R == 0 ? 1 : 2 * R);
                for (int bx = face ? xlo : qbx - R; bx <= (face ? xhi : qbx + R); bx += step) {
                    if (bx < x0 || bx > x1) continue;
                    const int gx = vx_gap(qx, bx << 5, (bx << 5) + 31);
                    if (gyz + (uint32_t)(gx * gx) > h.d2) continue;
                    const int slot = vx_slot(S, bx, by, bz);
                    if (slot >= 0) vx_scan_brick(S, (uint32_t)slot, bx, by, bz, qx, qy, qz, h);
                }
            }
        }
        // every brick not visited yet is at Chebyshev ring >= R + 1: at least 8 R + 1 voxels away
        const uint32_t m = (uint32_t)(8 * R + 1);
        if (m > 65535u || m * m > h.d2) break;
    }
    return true;
}

}  // namespace pccm
