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
//   rows[slot][64]   : {occupancy word, rank of the first voxel of the row} (the rank is already global:
//                      brick base + rows before r); rank of a voxel = row rank + popc(bits below) -- ONE
//                      8-byte load.  The entry of the brick past the last one holds the number of
//                      distinct voxels.
//   vxyz[rank]       : {x | y << 16, z} of the voxel
//   vkey[rank]       : {rgb, idx} of the voxel's point with the SMALLEST original index (only that
//                      one can win under the tie rule; one 64-bit atomicMin per point)
//   prank[i]         : rank of the voxel of input point i.  As queries the points of a voxel
//                      share ONE search (one lane per voxel); the per-point epilogue runs in the
//                      original order of the input and fetches its voxel's answer through prank.
// Two clouds of a pair share the arrays (cloud 1's slots, ranks and points continue cloud 0's).
//
// Search.  A query's candidates closer than 2 voxels are the 27 bits of the 3 x 3 x 3 voxels around
// it: three bits of each of nine occupancy rows.  Those 27 bits, sorted by the four distance levels
// (0, 1, 2, 3), are the answer AND the list of voxels that tie at the minimum; only those are looked
// up (rank -> vkey) to apply the smallest-index rule.  Queries with an empty 27-neighbourhood
// (1-2 % on codec-like content) repeat the game on the 5 x 5 x 5 voxels (exact below distance 3),
// then on whole bricks (vx_search_general: brick rings, up to a ring limit) and, beyond that (clouds
// far apart, isolated outliers), go to the pencil search of pccm_core.cuh.
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
PCCM_HD uint32_t vx_fshr(uint32_t lo, uint32_t hi, int s) {   // low word of (hi:lo) >> s, 0 <= s <= 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(lo, hi, (uint32_t)s);
#else
    return s == 0 ? lo : (s >= 32 ? hi : (lo >> s) | (hi << (32 - s)));
#endif
}
PCCM_HD void vx_atomic_or(uint32_t* p, uint32_t v) {     // (no result: a reduction, not a round trip)
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
PCCM_HD void vx_atomic_min64(unsigned long long* p, unsigned long long v) {     // (no result: a reduction, not a round trip)
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
PCCM_HD uint32_t vx_ld32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// L2-fresh load for check-before-atomic patterns: an L1 hit could show a word as it was when the SM first read it,
// and every thread of that SM would then repeat an atomic that has long been done
PCCM_HD uint32_t vx_ld32_fresh(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
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
    const uint2* rows;               // [nblk_total + 1][64] {occupancy word, rank of the row's first voxel}
    const uint2* vxyz;               // [n_total]
    const uint2* vkey;               // [n_total]
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
    const uint2 m = vx_ld64(G.rows + (size_t)slot * kVxRows + r);
    VX_CHECK((m.x >> xbit) & 1u);                     // the voxel we rank must be occupied
    return m.y + (uint32_t)vx_popc(m.x & ((1u << xbit) - 1u));
}
// voxels of brick `slot` are the ranks [vx_brick_begin(slot), vx_brick_begin(slot + 1))
PCCM_HD uint32_t vx_brick_begin(const VoxView& G, uint32_t slot) { return vx_ld32(&G.rows[(size_t)slot * kVxRows].y); }
// positions of the cloud's records in the joint array
PCCM_HD uint32_t vx_ranked_begin(const VoxView& G) { return vx_brick_begin(G, G.slot0); }
PCCM_HD uint32_t vx_ndistinct(const VoxView& G) { return vx_brick_begin(G, G.slot0 + G.nblk) - vx_brick_begin(G, G.slot0); }

// ---- build, per point (the kernels call these once per input point, pass after pass) --------
PCCM_HD void vx_mark_point(uint32_t* dirbits, uint32_t key) {
    const uint32_t bit = 1u << (key & 31u);
#ifndef PCCM_MARK_MODE
#define PCCM_MARK_MODE 0
#endif
#if PCCM_MARK_MODE == 0
    if (!(dirbits[key >> 5] & bit)) vx_atomic_or(dirbits + (key >> 5), bit);   // a stale read only costs a redundant atomic
#elif PCCM_MARK_MODE == 1
    if (!(vx_ld32_fresh(dirbits + (key >> 5)) & bit)) vx_atomic_or(dirbits + (key >> 5), bit);
#else
    vx_atomic_or(dirbits + (key >> 5), bit);
#endif
}
PCCM_HD uint32_t vx_slot_of_key(const uint32_t* dirbits, const uint32_t* dirpre, uint32_t key) {
    const uint32_t w = dirbits[key >> 5];
    return dirpre[key >> 5] + (uint32_t)vx_popc(w & ((1u << (key & 31u)) - 1u));
}
PCCM_HD void vx_fill_point(uint2* rows, uint32_t slot, int x, int y, int z) {
    // unconditional: one reduction at the L2 per point costs no more than the load that could avoid it
    vx_atomic_or(&rows[(size_t)slot * kVxRows + vx_row(y, z)].x, 1u << (x & 31));
}
// pass 3 (after the row bases): rank of the point's voxel; voxel coordinates into vxyz; the smallest
// original index (with its colour) wins vkey.  vkey must be pre-filled with 0xFF.
PCCM_HD uint32_t vx_place_point(const uint2* rows, uint2* vxyz, uint2* vkey,
                                uint32_t slot, int x, int y, int z, uint32_t rgb, uint32_t idx) {
    const uint2 m = vx_ld64(rows + (size_t)slot * kVxRows + vx_row(y, z));
    const uint32_t rank = m.y + (uint32_t)vx_popc(m.x & ((1u << (x & 31)) - 1u));
    VX_CHECK((m.x >> (x & 31)) & 1u);
    // a reduction, not a round trip: nobody waits for the old value.  Every point of the voxel writes the voxel's
    // coordinates (the same 8 bytes) -- cheaper than learning from the atomic who came first
    vx_atomic_min64(reinterpret_cast<unsigned long long*>(vkey + rank), ((unsigned long long)idx << 32) | rgb);
    if (vxyz) {                                          // (null: the coordinates are written brick by brick, from the occupancy rows)
        uint2 c;
        c.x = (uint32_t)x | ((uint32_t)y << 16);
        c.y = (uint32_t)z;
        vxyz[rank] = c;
    }
    return rank;
}

// ---- staged search (rows within 2 of the query, x within 2) ----------------------------------
// Region around query brick (bx, by, bz): rows y in [8 by - 2, 8 by + 10), z likewise; row i =
// rz * 12 + ry holds the 64 bits x in [32 bx - 2, 32 bx + 62) as {lo, hi}: bit (lx + 2 + dx) of the
// pair is voxel x + dx for a query at x = 32 bx + lx.  rb[i] is the rank of the first voxel of the
// row's centre word (the brick column the query brick is in): a candidate inside that column is
// ranked from shared memory alone.
PCCM_HD uint2 vx_stage_row(const VoxView& S, const int* sslot, int i, uint32_t& rb) {
    const int ry = i % kVxRegY, rz = i / kVxRegY;
    const int ny_i = ry < 2 ? 0 : (ry < 10 ? 1 : 2), nz_i = rz < 2 ? 0 : (rz < 10 ? 1 : 2);
    const int r_in = (((rz + 6) & 7) << 3) | ((ry + 6) & 7);
    const int* s3 = sslot + nz_i * 9 + ny_i * 3;
    const uint32_t l = s3[0] >= 0 ? vx_ld32(&S.rows[(size_t)s3[0] * kVxRows + r_in].x) : 0u;
    uint2 cw;
    cw.x = 0u; cw.y = 0u;
    if (s3[1] >= 0) cw = vx_ld64(S.rows + (size_t)s3[1] * kVxRows + r_in);
    const uint32_t r = s3[2] >= 0 ? vx_ld32(&S.rows[(size_t)s3[2] * kVxRows + r_in].x) : 0u;
    const uint32_t c = cw.x;
    rb = cw.y;
    uint2 w;
    w.x = (l >> 30) | (c << 2);
    w.y = (c >> 30) | (r << 2);
    return w;
}
PCCM_HD uint32_t vx_centre_word(const uint2 w) { return (w.x >> 2) | (w.y << 30); }

// bits x-2 .. x+2 of window row (ly + dy, lz + dz) as bits 0 .. 4
PCCM_HD uint32_t vx_row5(const uint2* win, int lx, int ly, int lz, int dy, int dz) {
    const uint2 w = win[(lz + dz) * kVxRegY + (ly + dy)];
    return vx_fshr(w.x, w.y, lx) & 31u;
}

// The 27 voxels around the query: bit 9 (dz + 1) + 3 (dy + 1) + (dx + 1).
PCCM_HD uint32_t vx_nb27(const uint2* win, int lx, int ly, int lz) {
    uint32_t nb = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 9; ++j) {
        const uint2 w = win[(lz + j / 3 - 1) * kVxRegY + (ly + j % 3 - 1)];
        nb |= (vx_fshr(w.x, w.y, lx + 1) & 7u) << (3 * j);
    }
    return nb;
}
// distance levels of those 27 bits
constexpr uint32_t kVxL0 = 1u << 13;                                                        // the query's own voxel
constexpr uint32_t kVxL1 = (1u << 4) | (1u << 10) | (1u << 12) | (1u << 14) | (1u << 16) | (1u << 22);
constexpr uint32_t kVxL3 = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 8) | (1u << 18) | (1u << 20) | (1u << 24) | (1u << 26);
constexpr uint32_t kVxL2 = 0x07FFFFFFu & ~(kVxL0 | kVxL1 | kVxL3);
// the voxels at the minimal distance (0 when the neighbourhood is empty) and that squared distance
PCCM_HD uint32_t vx_level27(uint32_t nb, uint32_t& d2) {
    if (nb & kVxL0) { d2 = 0; return nb & kVxL0; }
    if (nb & kVxL1) { d2 = 1; return nb & kVxL1; }
    if (nb & kVxL2) { d2 = 2; return nb & kVxL2; }
    d2 = 3;
    return nb & kVxL3;
}

struct VxPick {           // the chosen neighbour
    uint32_t idx, rgb;
    int ex, ey, ez;       // query - neighbour
};

// one candidate voxel (cx, cy, cz) = query + (dx, dy, dz), |d*| <= 2: its {rgb, idx}; smallest idx wins
PCCM_HD void vx_cand(const VoxView& S, const int* sslot, const uint2* win, const uint32_t* rb,
                     int lx, int ly, int lz, int dx, int dy, int dz, VxPick& pk) {
    const int wi = (lz + dz) * kVxRegY + (ly + dy);
    const int cxl = lx + dx;                                  // x of the candidate relative to the query's brick column
    uint32_t rank;
    if ((unsigned)cxl < 32u) {                                // inside the centre column: ranked from the staged words
        const uint32_t c = vx_centre_word(win[wi]);
        VX_CHECK((c >> cxl) & 1u);
        rank = rb[wi] + (uint32_t)vx_popc(c & ((1u << cxl) - 1u));
    } else {                                                  // left / right brick column (x of the brick 0, 1 or 30, 31)
        const int ry = ly + dy, rz = lz + dz;
        const int ny_i = ry < 2 ? 0 : (ry < 10 ? 1 : 2), nz_i = rz < 2 ? 0 : (rz < 10 ? 1 : 2);
        const int nb = nz_i * 9 + ny_i * 3 + (cxl < 0 ? 0 : 2);
        VX_CHECK(sslot[nb] >= 0);
        rank = vx_rank(S, (uint32_t)sslot[nb], (((rz + 6) & 7) << 3) | ((ry + 6) & 7), cxl & 31);
    }
    VX_CHECK(rank < S.n_total);
    const uint2 a = vx_ld64(S.vkey + rank);                   // {rgb, idx}
    if (a.y < pk.idx) { pk.idx = a.y; pk.rgb = a.x; pk.ex = -dx; pk.ey = -dy; pk.ez = -dz; }
}

// among the voxels of a 27-bit candidate set the one whose point has the smallest original index
PCCM_HD void vx_pick27(const VoxView& S, const int* sslot, const uint2* win, const uint32_t* rb,
                       int lx, int ly, int lz, uint32_t cand, VxPick& pk) {
    pk.idx = kVxNone; pk.rgb = 0; pk.ex = pk.ey = pk.ez = 0;
    while (cand) {
        const int b = vx_ffs(cand) - 1;
        cand &= cand - 1u;
        const int j = b / 3, dx = b - 3 * j - 1, dz = j / 3 - 1, dy = j - 3 * (dz + 1) - 1;
        vx_cand(S, sslot, win, rb, lx, ly, lz, dx, dy, dz, pk);
    }
}

// The 5 x 5 x 5 voxels around a query as 125 bits in four words: row j = 5 (dz + 2) + (dy + 2) sits in word j / 6
// at bit 5 (j % 6), its five bits are dx = -2 .. 2.  The voxels at squared distance L are a compile-time mask.
#if defined(__CUDACC__)
#define PCCM_HDC __host__ __device__ constexpr
#else
#define PCCM_HDC constexpr
#endif
PCCM_HDC uint32_t vx_mask125(int L, int word) {
    uint32_t m = 0;
    for (int j = 0; j < 25; ++j) {
        if (j / 6 != word) continue;
        const int dy = j % 5 - 2, dz = j / 5 - 2;
        for (int dx = -2; dx <= 2; ++dx)
            if (dx * dx + dy * dy + dz * dz == L) m |= 1u << (5 * (j % 6) + dx + 2);
    }
    return m;
}
struct VxNb125 { uint32_t w[5]; };

PCCM_HD void vx_nb125(const uint2* win, int lx, int ly, int lz, VxNb125& nb) {
    nb.w[0] = nb.w[1] = nb.w[2] = nb.w[3] = nb.w[4] = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 25; ++j) nb.w[j / 6] |= vx_row5(win, lx, ly, lz, j % 5 - 2, j / 5 - 2) << (5 * (j % 6));
}
template <int L>
PCCM_HD bool vx_level125(const VxNb125& nb, VxNb125& cand) {
    constexpr uint32_t m0 = vx_mask125(L, 0), m1 = vx_mask125(L, 1), m2 = vx_mask125(L, 2), m3 = vx_mask125(L, 3), m4 = vx_mask125(L, 4);
    cand.w[0] = nb.w[0] & m0; cand.w[1] = nb.w[1] & m1; cand.w[2] = nb.w[2] & m2; cand.w[3] = nb.w[3] & m3; cand.w[4] = nb.w[4] & m4;
    return (cand.w[0] | cand.w[1] | cand.w[2] | cand.w[3] | cand.w[4]) != 0u;
}

// Minimal squared distance within those 125 voxels (exact when < 9: anything outside is 3+ voxels away in one
// axis) and the winner among the voxels that tie at it.  Returns kVxNone when nothing closer than 3 is there.
PCCM_HD uint32_t vx_search125(const VoxView& S, const int* sslot, const uint2* win, const uint32_t* rb,
                              int lx, int ly, int lz, VxPick& pk) {
    VxNb125 nb, cand;
    vx_nb125(win, lx, ly, lz, nb);
    pk.idx = kVxNone; pk.rgb = 0; pk.ex = pk.ey = pk.ez = 0;
    uint32_t best;
    if (vx_level125<0>(nb, cand)) best = 0;
    else if (vx_level125<1>(nb, cand)) best = 1;
    else if (vx_level125<2>(nb, cand)) best = 2;
    else if (vx_level125<3>(nb, cand)) best = 3;
    else if (vx_level125<4>(nb, cand)) best = 4;
    else if (vx_level125<5>(nb, cand)) best = 5;
    else if (vx_level125<6>(nb, cand)) best = 6;
    else if (vx_level125<8>(nb, cand)) best = 8;
    else return kVxNone;
    for (int k = 0; k < 5; ++k) {
        uint32_t c = cand.w[k];
        while (c) {
            const int b = vx_ffs(c) - 1;
            c &= c - 1u;
            const int j = 6 * k + b / 5, dx = b % 5 - 2;
            vx_cand(S, sslot, win, rb, lx, ly, lz, dx, j % 5 - 2, j / 5 - 2, pk);
        }
    }
    return best;
}

// ---- boundary distances (nearest OTHER voxel of the same cloud) on the staged rows --------------
// minimal squared distance over the 26 / 124 voxels around the query (own voxel excluded); kVxNone when empty
PCCM_HD uint32_t vx_self27(uint32_t nb) {
    if (nb & kVxL1) return 1u;
    if (nb & kVxL2) return 2u;
    if (nb & kVxL3) return 3u;
    return kVxNone;
}
PCCM_HD uint32_t vx_self125(const uint2* win, int lx, int ly, int lz) {
    uint32_t best = kVxNone;
    for (int j = 0; j < 25; ++j) {
        const int dy = j % 5 - 2, dz = j / 5 - 2;
        uint32_t v = vx_row5(win, lx, ly, lz, dy, dz);
        if (dy == 0 && dz == 0) v &= ~4u;
        if (!v) continue;
        const uint32_t dx2 = (v & 4u) ? 0u : ((v & 10u) ? 1u : 4u);
        const uint32_t d2 = dx2 + (uint32_t)(dy * dy + dz * dz);
        best = d2 < best ? d2 : best;
    }
    return best;
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
    const uint2 a = vx_ld64(S.vkey + rank);
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
            const uint32_t m = vx_ld32(&S.rows[(size_t)slot * kVxRows + r].x);
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
