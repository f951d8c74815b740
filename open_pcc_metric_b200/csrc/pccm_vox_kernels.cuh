// pccm_vox_kernels.cuh -- sm_100a kernels of the occupancy-brick path for voxelised pairs:
//   index build   vx_mark / vx_dircount / vx_fill / vx_brickpre / vx_place (+ the scans of pccm_kernels.cuh)
//   query stage   vx_search_kernel   one warp per query brick, one lane per VOXEL, search rows staged in shared
//                                    memory, tie look-ups, whole-brick scan for the rare undecided voxel
//                 vx_general_kernel  what is still undecided: one warp per voxel over 125 bricks
//                 vx_far_kernel      beyond 16 voxels: the pencil search (second round)
//                 vx_epilogue_kernel one lane per query POINT in input order: D1 / D2 / colour + reduction records
//   boundary      vx_dupflag / vx_selfnn / vx_selfout: distance to the nearest OTHER point of the same cloud
// Per-point / per-query logic shared with the CPU stepping harness: pccm_vox.cuh.
#pragma once
#include "pccm_kernels.cuh"
#include "pccm_vox.cuh"

namespace pccm {

// ------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------
struct VoxCloudBuild {
    const void* xyz;
    const void* rgb;
    int64_t stride, rgb_stride;
    int32_t dtype, rgb_dtype, rgb_in_rec;
    uint32_t n;
    VoxDims g;
    uint32_t dir_off;          // word offset of this cloud's directory in the joint dirbits / dirpre
};
struct VoxBuild {
    VoxCloudBuild c[2];
    int32_t nclouds;
    uint32_t n_total, ndirw_total, nblk_total;
    uint32_t* dirbits;         // [ndirw_total]
    uint32_t* dirpre;          // [ndirw_total + 1]
    uint32_t* masks;           // [nblk_total][64]
    uint16_t* pre;             // [nblk_total][64]
    uint32_t* base;            // [nblk_total + 1]
    uint4* recs;               // [n_total] 0xFF-filled
    uint32_t* prank;           // [n_total] rank of the voxel of input point i (cloud 1's points follow cloud 0's)
    uint2* packed;             // [n_total] scratch: {x | y << 16, z} of input point i (written by the fill pass)
    uint32_t* pslot;           // [n_total] scratch: brick slot of input point i
};

__device__ __forceinline__ void vx_point(const VoxBuild& B, uint32_t i, int& c, uint32_t& li, int& x, int& y, int& z) {
    c = (B.nclouds > 1 && i >= B.c[0].n) ? 1 : 0;
    li = i - (c ? B.c[0].n : 0u);
    const VoxCloudBuild& C = B.c[c];
    x = (int)load_coord(C.xyz, C.dtype, C.stride, li, 0);
    y = (int)load_coord(C.xyz, C.dtype, C.stride, li, 1);
    z = (int)load_coord(C.xyz, C.dtype, C.stride, li, 2);
}

// The per-point passes are chains of dependent loads (coordinates -> directory -> masks -> atomic):
// every thread carries kVxIlp points, stage by stage, so that their loads are in flight together.
constexpr int kVxIlp = 2;
__device__ __forceinline__ uint32_t vx_ilp_index(const VoxBuild& B, int k) {     // point k of this thread (>= n_total: none)
    const uint32_t per = (B.n_total + kVxIlp - 1) / kVxIlp;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    return t < per ? t + (uint32_t)k * per : 0xFFFFFFFFu;
}

__global__ void vx_mark_kernel(const __grid_constant__ VoxBuild B) {
    int c[kVxIlp], x[kVxIlp], y[kVxIlp], z[kVxIlp];
    bool on[kVxIlp];
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {
        const uint32_t i = vx_ilp_index(B, k);
        on[k] = i < B.n_total;
        uint32_t li;
        if (on[k]) vx_point(B, i, c[k], li, x[k], y[k], z[k]);
    }
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k)
        if (on[k]) vx_mark_point(B.dirbits + B.c[c[k]].dir_off, vx_key(B.c[c[k]].g, x[k], y[k], z[k]));
}

__global__ void vx_dircount_kernel(const uint32_t* __restrict__ dirbits, uint32_t nw, uint32_t* __restrict__ dirpre) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nw) dirpre[i] = i < nw ? (uint32_t)__popc(dirbits[i]) : 0u;
}

__global__ void vx_fill_kernel(const __grid_constant__ VoxBuild B) {
    int c[kVxIlp], x[kVxIlp], y[kVxIlp], z[kVxIlp];
    uint32_t idx[kVxIlp], slot[kVxIlp];
    bool on[kVxIlp];
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {
        idx[k] = vx_ilp_index(B, k);
        on[k] = idx[k] < B.n_total;
        uint32_t li;
        if (on[k]) vx_point(B, idx[k], c[k], li, x[k], y[k], z[k]);
    }
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k)
        if (on[k]) {
            const uint32_t off = B.c[c[k]].dir_off;
            slot[k] = vx_slot_of_key(B.dirbits + off, B.dirpre + off, vx_key(B.c[c[k]].g, x[k], y[k], z[k]));
        }
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k)
        if (on[k]) {
            vx_fill_point(B.masks, slot[k], x[k], y[k], z[k]);
            B.packed[idx[k]] = make_uint2((uint32_t)x[k] | ((uint32_t)y[k] << 16), (uint32_t)z[k]);   // later passes need not parse the input again
            B.pslot[idx[k]] = slot[k];
        }
}

// one warp per brick: exclusive prefix of the row popcounts, brick total -> base[slot] (scanned next)
__global__ void __launch_bounds__(256) vx_brickpre_kernel(const __grid_constant__ VoxBuild B) {
    __shared__ uint32_t s_tot[8];
    const uint32_t slot = (blockIdx.x * 256u + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = 0;
    if (slot < B.nblk_total) {
        const uint32_t* m = B.masks + (size_t)slot * kVxRows;
        const uint32_t c0 = (uint32_t)__popc(m[2 * lane]), c1 = (uint32_t)__popc(m[2 * lane + 1]);
        incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t ex = incl - c0 - c1;
        B.pre[(size_t)slot * kVxRows + 2 * lane] = (uint16_t)ex;
        B.pre[(size_t)slot * kVxRows + 2 * lane + 1] = (uint16_t)(ex + c0);
    }
    if (lane == 31) s_tot[warp] = incl;
    __syncthreads();
    // the block's eight brick totals leave as one full 32-byte sector (the entry past the last brick is zero)
    if (threadIdx.x < 8) {
        const uint32_t sl = blockIdx.x * 8u + threadIdx.x;
        if (sl <= B.nblk_total) B.base[sl] = s_tot[threadIdx.x];
    }
}

__device__ __forceinline__ uint32_t vx_point_rgb(const VoxCloudBuild& C, uint32_t li) {
    if (!C.rgb_in_rec) return 0u;   // colours that have already arrived ride in the record of the voxel's representative
    if (C.rgb_dtype == PCCM_U8) {
        const uint8_t* p = static_cast<const uint8_t*>(C.rgb) + (int64_t)li * C.rgb_stride;
        return p[0] | (p[1] << 8) | (p[2] << 16);
    }
    return (uint32_t)rint(load_coord(C.rgb, PCCM_F64, C.rgb_stride, li, 0) * 255.0) |
           ((uint32_t)rint(load_coord(C.rgb, PCCM_F64, C.rgb_stride, li, 1) * 255.0) << 8) |
           ((uint32_t)rint(load_coord(C.rgb, PCCM_F64, C.rgb_stride, li, 2) * 255.0) << 16);
}

__global__ void vx_place_kernel(const __grid_constant__ VoxBuild B) {
    uint32_t idx[kVxIlp], li[kVxIlp], slot[kVxIlp], rgba[kVxIlp];
    uint2 pk[kVxIlp];
    bool on[kVxIlp];
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {
        idx[k] = vx_ilp_index(B, k);
        on[k] = idx[k] < B.n_total;
        if (on[k]) {
            const int c = (B.nclouds > 1 && idx[k] >= B.c[0].n) ? 1 : 0;
            li[k] = idx[k] - (c ? B.c[0].n : 0u);
            pk[k] = B.packed[idx[k]];
            slot[k] = B.pslot[idx[k]];
            rgba[k] = vx_point_rgb(B.c[c], li[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k)
        if (on[k]) {
            VX_CHECK(slot[k] < B.nblk_total);
            const uint32_t rank = vx_place_point(B.masks, B.pre, B.base, B.recs, slot[k], (int)(pk[k].x & 0xffffu), (int)(pk[k].x >> 16), (int)pk[k].y, rgba[k], li[k]);
            VX_CHECK(rank < B.n_total && rank >= B.base[slot[k]] && rank < B.base[slot[k] + 1]);
            B.prank[idx[k]] = rank;
        }
}

// ------------------------------------------------------------------------------------
// query
// ------------------------------------------------------------------------------------
struct VxDir {
    VoxView q, s;
    CloudView qa, sa;          // attribute views (colours, normals) of the query / search cloud
    uint32_t flags;
    int32_t* idx_out;          // original query order, or null
    double* d2_out;
    uint32_t* todo;            // ranked positions of the voxels the staged search left undecided
    uint32_t* todo_count;
    uint32_t* far;             // ... and of those the brick rings left undecided (pencil search)
    uint32_t* far_count;
    RowGrid sgrid;             // pencil index of the search cloud (vx_far_kernel only)
    const uint4* srecs;
    const uint32_t* srow_start;
    uint32_t rec_off;          // first reduction record of this direction
    uint32_t ntiles;           // ceil(q.n / kVxEpiTile): reduction records of one vx_epilogue_kernel pass
};

struct VxParams {
    VxDir dir[2];
    int32_t ndirs;
    int32_t normals_mode;
    int32_t rank, world;       // this call handles the rank-th of `world` equal slices of each cloud's voxels
    double T[9];
    double color_scale;
    BlockPartial* partials;
    uint4* vres;               // [n_total] by ranked position: {d2, packed (query - neighbour), neighbour idx, neighbour rgb};
                               // d2 == kVxNone while the voxel is undecided.  Pencil-round answers set the top bit of
                               // the neighbour idx and put the neighbour's position in the pencil records into .y
    int32_t pass;              // vx_epilogue_kernel: 0 = brick answers, 1 = pencil-round answers only
};
constexpr uint32_t kVxFarBit = 0x80000000u;

__device__ __forceinline__ uint32_t vx_pack_e(int ex, int ey, int ez) {   // |e| <= 16 for every certified brick answer
    return (uint32_t)(ex + 128) | ((uint32_t)(ey + 128) << 8) | ((uint32_t)(ez + 128) << 16);
}

struct VxAcc {
    unsigned long long s1;
    uint32_t m1, cnt;
    double s2, m2, cs[3], cm[3];
    __device__ __forceinline__ void init() {
        s1 = 0; m1 = 0; cnt = 0; s2 = 0; m2 = -INFINITY;
        for (int k = 0; k < 3; ++k) { cs[k] = 0; cm[k] = -INFINITY; }
    }
};

// epilogue of ONE query point: D1 (+ per-point outputs), D2 with the other cloud's normals, colour
__device__ __forceinline__ void vx_epilogue(const VxParams& P, const VxDir& D, const CloudView& qa, const CloudView& sa,
                                            uint32_t qidx, uint32_t qrgb, uint32_t d2,
                                            int ex, int ey, int ez, uint32_t nidx, uint32_t nrgb, VxAcc& a) {
    a.s1 += d2;
    a.m1 = d2 > a.m1 ? d2 : a.m1;
    a.cnt++;
    if (D.idx_out) D.idx_out[qidx] = (int32_t)nidx;
    if (D.d2_out) D.d2_out[qidx] = (double)d2;
    if (D.flags & PCCM_EVAL_D2) {
        const double e[3] = {(double)ex, (double)ey, (double)ez};
        const uint32_t ni = P.normals_mode == PCCM_NORMALS_BY_NEIGHBOUR ? nidx : qidx;
        double nv[3] = {__ldg(sa.normals + 3 * (size_t)ni), __ldg(sa.normals + 3 * (size_t)ni + 1),
                        __ldg(sa.normals + 3 * (size_t)ni + 2)};
        const double pe = plane_err2(e, nv);
        a.s2 = dadd(a.s2, pe);
        a.m2 = fmax(a.m2, pe);
    }
    if (D.flags & PCCM_EVAL_COLOR) {
        double cq[3], cn[3], c2[3], c2s[3];
        load_color(qa, qidx, qrgb, cq);
        load_color(sa, nidx, nrgb, cn);
        color_diff2(P.T, cq, cn, P.color_scale, c2, c2s);
        for (int k = 0; k < 3; ++k) { a.cs[k] = dadd(a.cs[k], c2[k]); a.cm[k] = fmax(a.cm[k], c2s[k]); }
    }
}

// sum_d1 (unused by integer pairs) carries the number of query POINTS reduced: slices are cut by
// voxel, so the host cannot know it
__device__ __forceinline__ void vx_warp_record(const VxAcc& a, uint32_t flags, BlockPartial& r) {
    const unsigned full = 0xffffffffu;
    r.sum_d1_u64 = warp_sum_u64(a.s1);
    const uint32_t cnt = __reduce_add_sync(full, a.cnt);
    r.sum_d1 = (double)cnt;
    r.max_d1 = cnt ? (double)__reduce_max_sync(full, a.m1) : -INFINITY;
    r.sum_d2 = 0; r.max_d2 = -INFINITY;
    for (int k = 0; k < 3; ++k) { r.csum[k] = 0; r.cmax[k] = -INFINITY; }
    if (flags & (PCCM_EVAL_D2 | PCCM_EVAL_COLOR)) {
        const double s = warp_reduce4<false>(a.s2, a.cs[0], a.cs[1], a.cs[2]);
        const double m = warp_reduce4<true>(a.m2, a.cm[0], a.cm[1], a.cm[2]);
        r.sum_d2 = __shfl_sync(full, s, 0);  r.csum[0] = __shfl_sync(full, s, 8);
        r.csum[1] = __shfl_sync(full, s, 16); r.csum[2] = __shfl_sync(full, s, 24);
        r.max_d2 = __shfl_sync(full, m, 0);  r.cmax[0] = __shfl_sync(full, m, 8);
        r.cmax[1] = __shfl_sync(full, m, 16); r.cmax[2] = __shfl_sync(full, m, 24);
    }
}

// this rank's slice [t_lo, t_hi) of the query cloud's ranked positions
__device__ __forceinline__ void vx_slice(const VxParams& P, const VoxView& Q, uint32_t& t_lo, uint32_t& t_hi) {
    const uint32_t r0 = vx_ranked_begin(Q), nd = vx_ndistinct(Q);
    if (P.world == 1) { t_lo = r0; t_hi = r0 + nd; return; }      // (no 64-bit division on the common path)
    t_lo = r0 + (uint32_t)((unsigned long long)nd * (unsigned)P.rank / (unsigned)P.world);
    t_hi = r0 + (uint32_t)((unsigned long long)nd * (unsigned)(P.rank + 1) / (unsigned)P.world);
}

// Warp-cooperative exact search of one query over a list of occupied bricks (slots[k], ids[k]; id
// = position in a DIM^3 neighbourhood centred on the query's brick).  Lane l owns rows 2l and
// 2l+1 of every brick (one coalesced load of the 64 occupancy words).  Pass 1: minimal squared
// distance by bit scans only.  Pass 2 (when that distance is below `limit`, i.e. certified): the
// smallest original index among the voxels at that distance.  Returns the distance; rank is valid
// when it is below `limit`.
template <int DIM, bool SELF = false>
__device__ __forceinline__ uint32_t vx_warp_bricks(const VoxView& S, const int* slots, const int* ids, int n,
                                                   int qx, int qy, int qz, uint32_t limit, uint32_t& rank_out) {
    // SELF: the query is a voxel of S itself -- its own bit is cleared and only the distance is wanted
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int qbx = qx >> 5, qby = qy >> 3, qbz = qz >> 3;
    uint32_t best = kVxNone;
    constexpr int kBatch = 4;                      // bricks in flight per step (independent coalesced loads; 8 spills)
    for (int k0 = 0; k0 < n; k0 += kBatch) {
        uint2 mb[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
            mb[j] = k0 + j < n ? __ldg(reinterpret_cast<const uint2*>(S.masks + (size_t)slots[k0 + j] * kVxRows) + lane) : make_uint2(0u, 0u);
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            if (k0 + j >= n) break;
            const int b = ids[k0 + j];
            const int bx = qbx + b % DIM - DIM / 2, by = qby + (b / DIM) % DIM - DIM / 2, bz = qbz + b / (DIM * DIM) - DIM / 2;
            const int p = qx - (bx << 5);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                uint32_t m = k ? mb[j].y : mb[j].x;
                const int r = 2 * lane + k, dy = qy - ((by << 3) + (r & 7)), dz = qz - ((bz << 3) + (r >> 3));
                if (SELF && dy == 0 && dz == 0 && (unsigned)p < 32u) m &= ~(1u << p);
                if (!m) continue;
                int dlo, dhi;
                vx_row_nearest(m, p, dlo, dhi);
                const int dx = dlo < dhi ? dlo : dhi;
                const uint32_t d2 = (uint32_t)(dx * dx + dy * dy + dz * dz);
                best = d2 < best ? d2 : best;
            }
        }
    }
    best = __reduce_min_sync(full, best);
    if (SELF || best >= limit) return best;
    uint32_t bidx = kVxNone, brank = kVxNone;
    for (int k0 = 0; k0 < n; ++k0) {
        const int slot = slots[k0], b = ids[k0];
        const int bx = qbx + b % DIM - DIM / 2, by = qby + (b / DIM) % DIM - DIM / 2, bz = qbz + b / (DIM * DIM) - DIM / 2;
        const uint2 m2 = __ldg(reinterpret_cast<const uint2*>(S.masks + (size_t)slot * kVxRows) + lane);
        const int p = qx - (bx << 5);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t m = k ? m2.y : m2.x;
            if (!m) continue;
            const int r = 2 * lane + k, dy = qy - ((by << 3) + (r & 7)), dz = qz - ((bz << 3) + (r >> 3));
            int dlo, dhi;
            vx_row_nearest(m, p, dlo, dhi);
            const uint32_t byz = (uint32_t)(dy * dy + dz * dz);
            if (byz + (uint32_t)(dlo * dlo) == best) {
                const uint32_t rank = vx_rank(S, (uint32_t)slot, r, (p - dlo) & 31);
                const uint32_t i = __ldg(reinterpret_cast<const uint32_t*>(S.recs + rank) + 3);
                if (i < bidx) { bidx = i; brank = rank; }
            }
            if (byz + (uint32_t)(dhi * dhi) == best) {
                const uint32_t rank = vx_rank(S, (uint32_t)slot, r, (p + dhi) & 31);
                const uint32_t i = __ldg(reinterpret_cast<const uint32_t*>(S.recs + rank) + 3);
                if (i < bidx) { bidx = i; brank = rank; }
            }
        }
    }
    const uint32_t widx = __reduce_min_sync(full, bidx);
    const int src = __ffs((int)__ballot_sync(full, bidx == widx)) - 1;
    rank_out = __shfl_sync(full, brank, src);
    return best;
}

// SEARCH.  One warp = one brick of the query cloud.  The warp stages the search cloud's occupancy
// rows around the brick (12 x 12 rows x 64 bits, 1.1 KB) in its private slice of shared memory,
// then takes the brick's voxels 32 at a time, one lane per voxel: bit scans over the 3 x 3 (5 x 5)
// rows, rank look-ups only for the voxels that tie at the minimum.  Integer work only; the answer of
// every voxel goes to vres[] (16 bytes), undecided voxels to the todo list.
#ifndef PCCM_VX_THREADS
#define PCCM_VX_THREADS 128
#endif
constexpr int kVxThreads = PCCM_VX_THREADS;
constexpr int kVxWarps = kVxThreads / 32;

#ifndef PCCM_VX_MINBLOCKS
#define PCCM_VX_MINBLOCKS 10      // <= 48 registers: the rare whole-brick scan may spill, the staged search must not lose occupancy
#endif
__global__ void __launch_bounds__(kVxThreads, PCCM_VX_MINBLOCKS)
vx_search_kernel(const __grid_constant__ VxParams P) {
    __shared__ uint2 s_win[kVxWarps][kVxRegRows];
    __shared__ int s_slot[kVxWarps][28];
    __shared__ int s_occ[kVxWarps][2][28];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * kVxWarps + warp;
    const int d = (P.ndirs > 1 && gw >= P.dir[0].q.nblk) ? 1 : 0;
    const VxDir& D = P.dir[d];
    const uint32_t lb = gw - (d ? P.dir[0].q.nblk : 0u);
    if (lb >= D.q.nblk) return;
    const uint32_t slot = D.q.slot0 + lb;
    const uint4* __restrict__ qrecs = D.q.recs;
    const uint32_t b0 = __ldg(D.q.base + slot), b1 = __ldg(D.q.base + slot + 1);
    uint32_t t_lo, t_hi;
    vx_slice(P, D.q, t_lo, t_hi);
    const uint32_t t0 = max(b0, t_lo), t1 = min(b1, t_hi);
    if (t0 >= t1) return;
    uint2* win = s_win[warp];
    int* sslot = s_slot[warp];
    int* occ_slot = s_occ[warp][0];
    int* occ_id = s_occ[warp][1];
    const uint2 first = __ldg(reinterpret_cast<const uint2*>(qrecs + b0));
    const int bx = (int)(first.x & 0xffffu) >> 5, by = (int)(first.x >> 16) >> 3, bz = (int)first.y >> 3;
    int myslot = -1;
    if (lane < 27) {
        myslot = vx_slot(D.s, bx + lane % 3 - 1, by + (lane / 3) % 3 - 1, bz + lane / 9 - 1);
        sslot[lane] = myslot;
    }
    const unsigned occ = __ballot_sync(full, myslot >= 0);
    const bool any_brick = occ != 0u;
    const int nocc = __popc(occ);
    if (myslot >= 0) {                       // compacted list of the occupied neighbour bricks (undecided voxels scan them whole)
        const int k = __popc(occ & ((1u << lane) - 1u));
        occ_slot[k] = myslot;
        occ_id[k] = lane;
    }
    __syncwarp();
    if (any_brick) {
        for (int i = lane; i < kVxRegRows; i += 32) win[i] = vx_stage_row(D.s, sslot, i);
        __syncwarp();
    }
    for (uint32_t tb = t0; tb < t1; tb += 32) {
        const uint32_t t = tb + lane;
        const bool active = t < t1;
        const uint2 qr = __ldg(reinterpret_cast<const uint2*>(qrecs + (active ? t : t0)));
        const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
        uint32_t bd2 = kVxNone, rows = 0;
        bool done = false;
        if (any_brick) {
            const int lx = qx & 31, ly = (qy & 7) + 2, lz = (qz & 7) + 2;
            vx_rows_inner(win, lx, ly, lz, bd2, rows);
            done = bd2 < 4u;
            if (__any_sync(full, active && !done)) {
                vx_rows_outer(win, lx, ly, lz, bd2, rows);
                done = bd2 < 9u;
            }
        }
        uint4 r = make_uint4(kVxNone, 0u, 0u, 0u);
        if (active && done) {
            VxPick pk;
            vx_pick(D.s, sslot, win, bx, by, bz, qx, qy, qz, rows, pk);
            r = make_uint4(bd2, vx_pack_e(pk.ex, pk.ey, pk.ez), pk.idx, pk.rgb);
        }
        // voxels the 5 x 5 rows left undecided (nearest point 3+ voxels away; rare): the whole warp scans the
        // 27 neighbour bricks for one voxel at a time -- anything outside them is at least 9 voxels away
#ifndef PCCM_VX_INKERNEL
#define PCCM_VX_INKERNEL 1
#endif
        unsigned pend = (PCCM_VX_INKERNEL && any_brick) ? __ballot_sync(full, active && !done) : 0u;
        while (pend) {
            const int src = __ffs((int)pend) - 1;
            pend &= pend - 1u;
            const int sx = __shfl_sync(full, qx, src), sy = __shfl_sync(full, qy, src), sz = __shfl_sync(full, qz, src);
            uint32_t nrank = kVxNone;
            const uint32_t nd2 = vx_warp_bricks<3>(D.s, occ_slot, occ_id, nocc, sx, sy, sz, 81u, nrank);
            if (nd2 < 81u && lane == src) {
                const uint4 nr = __ldg(D.s.recs + nrank);
                r = make_uint4(nd2, vx_pack_e(qx - (int)(nr.x & 0xffffu), qy - (int)(nr.x >> 16), qz - (int)nr.y), nr.w, nr.z);
                done = true;
            }
        }
        if (active) P.vres[t] = r;
        const unsigned und = __ballot_sync(full, active && !done);
        if (und) {
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(D.todo_count, (uint32_t)__popc(und));
            pos = __shfl_sync(full, pos, 0);
            VX_CHECK(pos + (uint32_t)__popc(und) <= D.q.n);
            if (active && !done) D.todo[pos + __popc(und & ((1u << lane) - 1u))] = t;
        }
    }
}

// EPILOGUE.  Threads walk the query POINTS in the ORIGINAL order of the input (4 per thread, strided so
// that every load is coalesced): own colour, the other cloud's normal at the query index (quirk Q1)
// and the per-point outputs stream; only the voxel's 16-byte answer is a gather (through prank).
// Points that share a voxel share its answer.  A block is a fixed tile of kVxEpiTile points and
// writes one reduction record -> float sums do not depend on scheduling.  Multi-GPU slices are cut
// by voxel: a rank skips the points of voxels it did not search.
constexpr int kVxEpiThreads = 256;
#ifndef PCCM_VX_EPIPER
#define PCCM_VX_EPIPER 4
#endif
constexpr int kVxEpiPer = PCCM_VX_EPIPER;
constexpr int kVxEpiTile = kVxEpiThreads * kVxEpiPer;
__global__ void __launch_bounds__(kVxEpiThreads)
vx_epilogue_kernel(const __grid_constant__ VxParams P) {
    __shared__ double s_lut[256];
    const int d = (P.ndirs > 1 && blockIdx.x >= P.dir[0].ntiles) ? 1 : 0;
    const VxDir& D = P.dir[d];
    const uint32_t tile = blockIdx.x - (d ? P.dir[0].ntiles : 0u);
    s_lut[threadIdx.x] = D.qa.lut255[threadIdx.x];       // k / 255.0 table: shared memory instead of six global loads per point
    __syncthreads();
    CloudView qa = D.qa, sa = D.sa;
    qa.lut255 = s_lut; sa.lut255 = s_lut;
    uint32_t t_lo, t_hi;
    vx_slice(P, D.q, t_lo, t_hi);
    const uint32_t i0 = tile * kVxEpiTile + threadIdx.x;
    uint32_t rk[kVxEpiPer];
    uint4 v[kVxEpiPer];
#pragma unroll
    for (int j = 0; j < kVxEpiPer; ++j) {
        const uint32_t i = i0 + j * kVxEpiThreads;
        rk[j] = i < D.q.n ? __ldg(D.q.prank + i) : kVxNone;
        VX_CHECK(i >= D.q.n || rk[j] < D.q.n_total);
    }
#pragma unroll
    for (int j = 0; j < kVxEpiPer; ++j)
        v[j] = (rk[j] >= t_lo && rk[j] < t_hi) ? __ldg(P.vres + rk[j]) : make_uint4(kVxNone, 0u, 0u, 0u);
    VxAcc acc;
    acc.init();
#pragma unroll
    for (int j = 0; j < kVxEpiPer; ++j) {
        if (v[j].x == kVxNone) continue;
        const uint32_t i = i0 + j * kVxEpiThreads;
        const bool far = (v[j].z & kVxFarBit) != 0u;
        if (far != (P.pass == 1)) continue;
        int ex, ey, ez;
        uint32_t nrgb = v[j].w;
        if (!far) {
            ex = (int)(v[j].y & 0xffu) - 128; ey = (int)((v[j].y >> 8) & 0xffu) - 128; ez = (int)((v[j].y >> 16) & 0xffu) - 128;
        } else {                                  // pencil-round answer: any distance, coordinates from the records
            const uint2 qv = __ldg(reinterpret_cast<const uint2*>(D.q.recs + rk[j]));
            const uint4 nr = __ldg(D.srecs + v[j].y);               // {xy, z, idx, rgb}
            ex = (int)(qv.x & 0xffffu) - (int)(nr.x & 0xffffu); ey = (int)(qv.x >> 16) - (int)(nr.x >> 16); ez = (int)qv.y - (int)nr.y;
            nrgb = nr.w;
        }
        VX_CHECK((v[j].z & ~kVxFarBit) < D.s.n);
        vx_epilogue(P, D, qa, sa, i, 0u, v[j].x, ex, ey, ez, v[j].z & ~kVxFarBit, nrgb, acc);
    }
    BlockPartial r;
    vx_warp_record(acc, D.flags, r);
    __shared__ BlockPartial sm[kVxEpiThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockPartial o = sm[0];
        for (int w = 1; w < kVxEpiThreads / 32; ++w) partial_merge(o, sm[w]);
        P.partials[D.rec_off + (uint32_t)P.pass * D.ntiles + tile] = o;
    }
}

// Undecided voxels: one WARP each, over the 125 bricks of rings 0..2 around the query.  Lanes look
// the bricks up in the directory; every occupied brick is then scanned by the whole warp (lane l
// owns rows 2l and 2l+1: one coalesced load of the 64 occupancy words), first for the minimal
// distance only (bit scans, no record is touched), then again to fetch the index of the voxels
// that tie at that distance.  An answer closer than 17 voxels is certified (everything unvisited is
// at least that far) and written to vres[] before the epilogue kernel runs; the rest goes to the
// pencil search (second round).
__global__ void __launch_bounds__(128)
vx_general_kernel(const __grid_constant__ VxParams P) {
    __shared__ int s_slot[4][128];
    __shared__ int s_occ[4][128];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int* bslot = s_slot[threadIdx.x >> 5];
    int* bocc = s_occ[threadIdx.x >> 5];
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int d = 0; d < P.ndirs; ++d) {
        const VxDir& D = P.dir[d];
        const uint32_t ntodo = *D.todo_count;
        for (uint32_t w = gwarp; w < ntodo; w += nwarps) {
            const uint32_t t = D.todo[w];
            const uint2 qr = __ldg(reinterpret_cast<const uint2*>(D.q.recs + t));
            const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
            const int qbx = qx >> 5, qby = qy >> 3, qbz = qz >> 3;
            __syncwarp();
            // occupied bricks of the 5 x 5 x 5 neighbourhood, compacted: bocc[k] = brick number, bslot[k] = slot
            int nocc = 0;
            for (int b0 = 0; b0 < 125; b0 += 32) {
                const int b = b0 + lane;
                const int slot = b < 125 ? vx_slot(D.s, qbx + b % 5 - 2, qby + (b / 5) % 5 - 2, qbz + b / 25 - 2) : -1;
                const unsigned has = __ballot_sync(full, slot >= 0);
                if (slot >= 0) {
                    const int k = nocc + __popc(has & ((1u << lane) - 1u));
                    bslot[k] = slot;
                    bocc[k] = b;
                }
                nocc += __popc(has);
            }
            __syncwarp();
            uint32_t brank = kVxNone;
            const uint32_t best = vx_warp_bricks<5>(D.s, bslot, bocc, nocc, qx, qy, qz, 289u, brank);
            if (best >= 289u) {               // nothing certified within two brick rings
                if (lane == 0) D.far[atomicAdd(D.far_count, 1u)] = t;
                continue;
            }
            if (lane == 0) {
                const uint4 nr = __ldg(D.s.recs + brank);
                P.vres[t] = make_uint4(best, vx_pack_e(qx - (int)(nr.x & 0xffffu), qy - (int)(nr.x >> 16), qz - (int)nr.y), nr.w, nr.z);
            }
        }
    }
}

// What the brick rings could not certify (the nearest point is tens of voxels away: disjoint
// clouds, isolated outliers): exact pencil search, which skips empty space by its row table.
__global__ void __launch_bounds__(128)
vx_far_kernel(const __grid_constant__ VxParams P) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int d = 0; d < P.ndirs; ++d) {
        const VxDir& D = P.dir[d];
        const uint32_t nfar = *D.far_count;
        for (uint32_t w = gtid; w < nfar; w += stride) {
            const uint32_t t = D.far[w];
            const uint4 qr = __ldg(D.q.recs + t);
            KInt::Q q;
            q.x = (int)(qr.x & 0xffffu); q.y = (int)(qr.x >> 16); q.z = (int)qr.y;
            Best1<KInt> best;
            best.init();
            search<KInt>(D.sgrid, D.srow_start, D.srecs, q, best);
            P.vres[t] = make_uint4(best.d2, best.pos, best.idx | kVxFarBit, 0u);   // .y: position in the pencil records
        }
    }
}

// ------------------------------------------------------------------------------------
// distance of every point to its nearest OTHER point (compute_nearest_neighbor_distance,
// cloud_pair.py:108-109) on the brick index: the same row scans with the voxel's own bit cleared;
// a voxel that holds more than one point answers 0
// ------------------------------------------------------------------------------------
struct VxSelfParams {
    VoxView c;
    uint32_t n;                // points of the cloud
    uint32_t begin, end;       // slice of the cloud's points [begin, end) -> the same share of its voxels
    uint32_t* dupbits;         // [n_total / 32 + 1] by rank: voxel holds more than one point
    uint32_t* vself;           // [n_total] by rank: squared distance to the nearest other point (kVxNone: not decided here)
    double* minmax;            // per brick {min, max} of the distances (sqrt)
    uint32_t* undecided;       // voxels farther than 8 voxels from every other one (caller falls back to the pencil path)
    double* per_point;         // optional, original order
};

__global__ void vx_dupflag_kernel(const __grid_constant__ VxSelfParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t rank = __ldg(P.c.prank + i);
    if (__ldg(reinterpret_cast<const uint32_t*>(P.c.recs + rank) + 3) != i) atomicOr(P.dupbits + (rank >> 5), 1u << (rank & 31u));
}

__global__ void __launch_bounds__(kVxThreads)
vx_selfnn_kernel(const __grid_constant__ VxSelfParams P) {
    __shared__ uint2 s_win[kVxWarps][kVxRegRows];
    __shared__ int s_slot[kVxWarps][28];
    __shared__ int s_occ[kVxWarps][2][28];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lb = blockIdx.x * kVxWarps + warp;
    if (lb >= P.c.nblk) return;
    const uint32_t slot = P.c.slot0 + lb;
    const uint32_t b0 = __ldg(P.c.base + slot), b1 = __ldg(P.c.base + slot + 1);
    const uint32_t r0 = vx_ranked_begin(P.c), nd = vx_ndistinct(P.c);
    const uint32_t t_lo = r0 + (uint32_t)((unsigned long long)nd * P.begin / (P.n ? P.n : 1u));
    const uint32_t t_hi = r0 + (uint32_t)((unsigned long long)nd * P.end / (P.n ? P.n : 1u));
    const uint32_t t0 = max(b0, t_lo), t1 = min(b1, t_hi);
    uint32_t mn = kVxNone, mx = 0u;
    bool any = false;
    if (t0 < t1) {
        uint2* win = s_win[warp];
        int* sslot = s_slot[warp];
        int* occ_slot = s_occ[warp][0];
        int* occ_id = s_occ[warp][1];
        const uint2 first = __ldg(reinterpret_cast<const uint2*>(P.c.recs + b0));
        const int bx = (int)(first.x & 0xffffu) >> 5, by = (int)(first.x >> 16) >> 3, bz = (int)first.y >> 3;
        int myslot = -1;
        if (lane < 27) {
            myslot = vx_slot(P.c, bx + lane % 3 - 1, by + (lane / 3) % 3 - 1, bz + lane / 9 - 1);
            sslot[lane] = myslot;
        }
        const unsigned occ = __ballot_sync(full, myslot >= 0);
        const int nocc = __popc(occ);
        if (myslot >= 0) {
            const int k = __popc(occ & ((1u << lane) - 1u));
            occ_slot[k] = myslot;
            occ_id[k] = lane;
        }
        __syncwarp();
        for (int i = lane; i < kVxRegRows; i += 32) win[i] = vx_stage_row(P.c, sslot, i);
        __syncwarp();
        for (uint32_t tb = t0; tb < t1; tb += 32) {
            const uint32_t t = tb + lane;
            const bool active = t < t1;
            const uint2 qr = __ldg(reinterpret_cast<const uint2*>(P.c.recs + (active ? t : t0)));
            const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
            const int lx = qx & 31, ly = (qy & 7) + 2, lz = (qz & 7) + 2;
            // the voxel's own bit sits at bit 16 + lx of its row window: clear it in a private copy of that row
            const bool dup = active && ((__ldg(P.dupbits + (t >> 5)) >> (t & 31u)) & 1u);
            uint32_t bd2 = kVxNone, rows = 0;
            {
                uint2 own = win[lz * kVxRegY + ly];
                const int bit = 16 + lx;
                if (bit < 32) own.x &= ~(1u << bit); else own.y &= ~(1u << (bit - 32));
                int dd, du;
                vx_row_dists(own, lx, dd, du);
                const int dx = dd < du ? dd : du;
                bd2 = (uint32_t)(dx * dx);
            }
            uint32_t nb = kVxNone, nrows = 0;
            vx_rows_ring1(win, lx, ly, lz, nb, nrows);
            bd2 = nb < bd2 ? nb : bd2;
            bool done = bd2 < 4u;
            if (__any_sync(full, active && !done && !dup)) {
                vx_rows_outer(win, lx, ly, lz, bd2, rows);
                done = bd2 < 9u;
            }
            unsigned pend = __ballot_sync(full, active && !done && !dup);
            while (pend) {
                const int src = __ffs((int)pend) - 1;
                pend &= pend - 1u;
                const int sx = __shfl_sync(full, qx, src), sy = __shfl_sync(full, qy, src), sz = __shfl_sync(full, qz, src);
                uint32_t dummy = kVxNone;
                const uint32_t nd2 = vx_warp_bricks<3, true>(P.c, occ_slot, occ_id, nocc, sx, sy, sz, 81u, dummy);
                if (nd2 < 81u && lane == src) { bd2 = nd2; done = true; }
            }
            if (active) {
                const uint32_t v = dup ? 0u : (done ? bd2 : kVxNone);
                P.vself[t] = v;
                if (v != kVxNone) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; any = true; }
            }
            const unsigned und = __ballot_sync(full, active && !dup && !done);
            if (und && lane == 0) atomicAdd(P.undecided, (uint32_t)__popc(und));
        }
    }
    mn = __reduce_min_sync(full, mn);
    mx = __reduce_max_sync(full, mx);
    const bool wany = __any_sync(full, any);
    if (lane == 0) {
        P.minmax[2 * lb] = wany ? sqrt((double)mn) : INFINITY;
        P.minmax[2 * lb + 1] = wany ? sqrt((double)mx) : -INFINITY;
    }
}

__global__ void vx_selfout_kernel(const __grid_constant__ VxSelfParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t v = P.vself[__ldg(P.c.prank + i)];
    P.per_point[i] = sqrt((double)v);
}

}  // namespace pccm
