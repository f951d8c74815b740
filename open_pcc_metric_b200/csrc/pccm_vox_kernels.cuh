// pccm_vox_kernels.cuh -- sm_100a kernels of the occupancy-brick path for voxelised pairs.
//
// The whole pair evaluation is enqueued WITHOUT a host synchronisation: nothing the host would have to
// wait for (bounding boxes, number of occupied bricks, number of distinct voxels) is a kernel argument or
// an allocation size -- the kernels read those from a VoxPlan record in device memory that earlier kernels
// of the same stream wrote, arrays are sized by capacities, and a status word tells the host at the one
// synchronisation at the end (the result read-back) whether a capacity or an assumption (integer
// coordinates, 8-bit colours) did not hold; only then is anything repeated.
//
//   index build   stats_kernel (pccm_kernels.cuh: bounding box + classification + 8-byte packed coordinates)
//                 vx_mark_kernel      directory bit per occupied brick
//                 vx_dirsum_kernel    popcount sums per directory chunk
//                 vx_dirscan_kernel   directory prefix -> brick slots; writes the VoxPlan; zeroes the occupancy words
//                 vx_fill_kernel      occupancy bit per point
//                 vx_bricksum_kernel  voxels per chunk of bricks
//                 vx_rowbase_kernel   rank of the first voxel of every brick row
//                 vx_place_kernel     rank of every point; voxel coordinates and smallest index per voxel
//   query stage   vx_search_kernel    one warp per query brick, one lane per VOXEL: 27-bit neighbourhood from the
//                                     staged rows, tie look-ups, 125-bit neighbourhood and whole-brick scans for the
//                                     few voxels whose neighbourhood is empty
//                 vx_general_kernel   what is still undecided: one warp per voxel over 125 bricks
//                 vx_far_kernel       beyond 16 voxels: the pencil search (second round)
//                 vx_epilogue_kernel  one lane per query POINT in input order: D1 / D2 / colour + reduction records
//   boundary      vx_dupflag / vx_selfnn / vx_selfout: distance to the nearest OTHER point of the same cloud
// Per-point / per-query logic shared with the CPU stepping harness: pccm_vox.cuh.
#pragma once
#include "pccm_kernels.cuh"
#include "pccm_vox.cuh"

namespace pccm {

// ------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------
constexpr uint32_t kVxStNotInt = 1u;          // a coordinate is not an integer in [0, 32767]
constexpr uint32_t kVxStDirOverflow = 2u;     // the brick grids of the bounding boxes exceed the directory capacity
constexpr uint32_t kVxStBrickOverflow = 4u;   // more occupied bricks than the mask arrays hold
constexpr uint32_t kVxStRgbNotU8 = 8u;        // a float colour is not k / 255 (the 8-bit colour arrays are invalid)

constexpr int kVxDirChunk = 2048;             // directory words per scan block
constexpr int kVxBrickChunk = 64;             // bricks per row-base chunk (8 warps x 8 bricks)
constexpr int kVxMaxBrickChunks = 1 << 16;
constexpr int kVxSizeClasses = 16;            // bricks are listed by size class (voxels / 32, capped): the search takes the large ones first

// One pair split over the GPUs of a box: every rank holds both clouds, OWNS the queries of a slab of
// 8-voxel z layers (cut so that the slabs hold equal numbers of points) and indexes only what those queries can see:
// its slab plus a halo of kVxHaloBricks bricks.  The cuts come from the z histogram of the statistics pass, on the
// device, identically on every rank -- nothing is exchanged until the partial sums at the very end.
constexpr int kVxHaloBricks = 2;          // vx_general_kernel certifies answers closer than 17 voxels from bricks within +-2
struct ShardPlan {
    int32_t own_zlo, own_zhi;             // brick layers [own_zlo, own_zhi) whose voxels this rank queries
    int32_t need_zlo, need_zhi;           // brick layers [need_zlo, need_zhi] that are indexed
};

struct VoxPlan {
    VoxView view[2];
    ShardPlan shard;                          // (own = everything, need = everything when the pair is not split)
    uint32_t status;
    uint32_t nvox_total;                      // distinct voxels of both clouds
    uint32_t ndirw[2], dir_off[2], ndirw_total;
    uint32_t cap_dirw, cap_blk;
    int32_t mn[2][3], mx[2][3];               // integer bounding boxes
};

struct VoxBuildArgs {                         // everything the host knows when it enqueues the build
    uint32_t* sel;                            // split pairs: the input points inside this rank's slab, compacted by vx_mark_kernel
    uint32_t* sel_count;                      //   (unordered; fill and place walk this list instead of every point); null otherwise
    const ShardPlan* shard;                   // null: the whole pair is indexed and queried here
    int32_t full_need;                        // sharded, but index everything (the fallback when a query has to look beyond its halo)
    uint32_t mark_lo, mark_hi;                // vx_mark_kernel: this launch handles the thread slots [mark_lo, mark_hi) of the point passes
    const DevStats* stats[2];
    const uint2* packed[2];                   // {x | y << 16, z} of every input point (stats_kernel)
    const void* rgb[2];                       // colours for the voxel records: packed uchar4 arrays (or null)
    uint32_t n[2];
    uint32_t cap_dirw, cap_blk;
    uint32_t* dirbits;                        // [cap_dirw] zeroed (PCCM_DIR_BYTES: written by vx_dirsum_kernel)
    uint8_t* dirbytes;                        // PCCM_DIR_BYTES: [cap_dirw * 32] zeroed -- one byte per brick of the grids, set by plain stores
    uint32_t* dirpre;                         // [cap_dirw + 1]
    uint32_t* dirsums;                        // [cap_dirw / kVxDirChunk + 1]
    uint2* rows;                              // [cap_blk][64] {occupancy word, rank of the row's first voxel}
    uint32_t* bricksums;                      // [cap_blk / kVxBrickChunk + 1]
    uint32_t* bcursor;                        // [kVxSizeClasses] zeroed: bricks of each size class
    uint32_t* border;                         // [kVxSizeClasses][cap_blk] slots of the bricks of each class (arrival order)
    uint2* vxyz;                              // [n_total]
    uint2* vkey;                              // [n_total] (0xFF-filled by vx_mark_kernel)
    uint32_t* prank;                          // [n_total]
    uint32_t* pslot;                          // [n_total] scratch: brick slot of input point i
    uint32_t* bkey;                           // [cap_blk] directory key of every brick (written by its points in the fill pass), or null:
                                              //   with it the voxel coordinates are written by vx_rowbase_kernel -- in rank order, from
                                              //   the occupancy rows -- instead of by every point in the place pass (a scattered 8-byte store)
    VoxPlan* plan;
};

// The cuts: layer z belongs to rank k when k / world of all points (both clouds) lie in the layers below it.
__global__ void __launch_bounds__(1024) vx_shardplan_kernel(const uint32_t* __restrict__ zh0, const uint32_t* __restrict__ zh1,
                                                            int rank, int world, ShardPlan* out) {
    pdl_enter();
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_pre[kZHistBins];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int kPer = kZHistBins / 1024;
    unsigned long long v[kPer], mine = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) { v[k] = (unsigned long long)zh0[t * kPer + k] + zh1[t * kPer + k]; mine += v[k]; }
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long run = incl - mine;
    for (int w = 0; w < warp; ++w) run += s_w[w];
    unsigned long long total = 0;
    for (int w = 0; w < 32; ++w) total += s_w[w];
#pragma unroll
    for (int k = 0; k < kPer; ++k) { run += v[k]; s_pre[t * kPer + k] = run; }       // inclusive prefix per layer
    __syncthreads();
    // cut(k) = number of layers whose inclusive prefix is <= total * k / world
    const unsigned long long lo_t = total * (unsigned long long)rank / (unsigned long long)world;
    const unsigned long long hi_t = total * (unsigned long long)(rank + 1) / (unsigned long long)world;
    int clo = 0, chi = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        clo += s_pre[t * kPer + k] <= lo_t ? 1 : 0;
        chi += s_pre[t * kPer + k] <= hi_t ? 1 : 0;
    }
    clo = __reduce_add_sync(0xffffffffu, clo);
    chi = __reduce_add_sync(0xffffffffu, chi);
    __shared__ int s_c[2][32];
    if (lane == 0) { s_c[0][warp] = clo; s_c[1][warp] = chi; }
    __syncthreads();
    if (t == 0) {
        int a = 0, b = 0;
        for (int w = 0; w < 32; ++w) { a += s_c[0][w]; b += s_c[1][w]; }
        ShardPlan sp;
        sp.own_zlo = rank == 0 ? -(1 << 30) : a;
        sp.own_zhi = rank == world - 1 ? (1 << 30) : b;
        sp.need_zlo = rank == 0 ? -(1 << 30) : a - kVxHaloBricks;
        sp.need_zhi = rank == world - 1 ? (1 << 30) : b - 1 + kVxHaloBricks;
        *out = sp;
    }
}

// Brick grids of the two clouds from the device-side statistics.  Cloud 1's directory starts at a chunk
// boundary, so the number of bricks of cloud 0 is a sum of whole chunk sums.  Returns the status bits.
__device__ __forceinline__ ShardPlan vx_shard_of(const VoxBuildArgs& A) {
    ShardPlan sp;
    sp.own_zlo = -(1 << 30); sp.own_zhi = 1 << 30; sp.need_zlo = -(1 << 30); sp.need_zhi = 1 << 30;
    if (A.shard) {
        sp = *A.shard;
        if (A.full_need) { sp.need_zlo = -(1 << 30); sp.need_zhi = 1 << 30; }
    }
    return sp;
}
__device__ __forceinline__ uint32_t vx_plan_dims(const VoxBuildArgs& A, VoxDims g[2], uint32_t ndirw[2], uint32_t dir_off[2],
                                                 int32_t mn[2][3], int32_t mx[2][3], ShardPlan& sp) {
    uint32_t status = 0;
    unsigned long long words = 0;
    sp = vx_shard_of(A);
    for (int c = 0; c < 2; ++c) {
        const DevStats s = *A.stats[c];
        if (s.flags & (kDevNotInt | kDevNonFinite)) status |= kVxStNotInt;
        for (int a = 0; a < 3; ++a) { mn[c][a] = (int32_t)(0x7fffffffu - s.nmn[a]); mx[c][a] = (int32_t)s.mx[a]; }
        if (status) { ndirw[c] = 0; dir_off[c] = 0; g[c] = VoxDims{0, 0, 0, 1, 1, 1}; continue; }
        g[c].obx = mn[c][0] >> 5; g[c].oby = mn[c][1] >> 3;
        g[c].nbx = (mx[c][0] >> 5) - g[c].obx + 1;
        g[c].nby = (mx[c][1] >> 3) - g[c].oby + 1;
        // layers of this cloud inside the slab that is indexed here (an empty intersection keeps one layer with no bit set)
        const int zlo = max(mn[c][2] >> 3, sp.need_zlo), zhi = min(mx[c][2] >> 3, sp.need_zhi);
        g[c].obz = zlo;
        g[c].nbz = zhi >= zlo ? zhi - zlo + 1 : 1;
        const unsigned long long bits = (unsigned long long)g[c].nbx * (unsigned long long)g[c].nby * (unsigned long long)g[c].nbz;
        const unsigned long long w = (bits + 31ull) / 32ull;
        dir_off[c] = (uint32_t)words;
        ndirw[c] = (uint32_t)(w > 0xffffffffull ? 0xffffffffull : w);
        words += w;
        if (c == 0) words = (words + kVxDirChunk - 1) / kVxDirChunk * kVxDirChunk;
        if (words > (unsigned long long)A.cap_dirw) status |= kVxStDirOverflow;
    }
    return status;
}

struct VoxDimsSmem {
    ShardPlan sp;
    VoxDims g[2];
    uint32_t ndirw[2], dir_off[2];
    int32_t mn[2][3], mx[2][3];
    uint32_t status;
};
__device__ __forceinline__ void vx_block_dims(const VoxBuildArgs& A, VoxDimsSmem& S) {
    if (threadIdx.x == 0) S.status = vx_plan_dims(A, S.g, S.ndirw, S.dir_off, S.mn, S.mx, S.sp);
    __syncthreads();
}

// The per-point passes are chains of dependent accesses (coordinates -> directory -> masks -> atomic):
// every thread carries kVxIlp points, stage by stage, so that their loads are in flight together.
#ifndef PCCM_VX_ILP
#define PCCM_VX_ILP 2
#endif
#ifndef PCCM_DIR_BYTES
#define PCCM_DIR_BYTES 0
#endif
constexpr int kVxIlp = PCCM_VX_ILP;
__device__ __forceinline__ uint32_t vx_ilp_index(uint32_t n_total, int k) {     // point k of this thread (>= n_total: none)
    const uint32_t per = (n_total + kVxIlp - 1) / kVxIlp;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    return t < per ? t + (uint32_t)k * per : 0xFFFFFFFFu;
}
__device__ __forceinline__ uint2 vx_point(const VoxBuildArgs& A, uint32_t i, int& c, uint32_t& li) {
    c = i >= A.n[0] ? 1 : 0;
    li = i - (c ? A.n[0] : 0u);
    return __ldg(A.packed[c] + li);
}
#define VX_UNPACK(p, x, y, z) const int x = (int)((p).x & 0xffffu), y = (int)((p).x >> 16), z = (int)(p).y
__device__ __forceinline__ bool vx_in_slab(const ShardPlan& sp, const uint2& p) { const int bz = (int)p.y >> 3; return bz >= sp.need_zlo && bz <= sp.need_zhi; }

// Two launches: a small sample of the points first (its atomics meet little contention: thousands of points name the
// same brick, and read-modify-writes of one directory word queue up at the L2), then everybody else -- who now finds
// nearly every bit set on the first look (a cached look is enough: bits are only ever set).
__global__ void __launch_bounds__(256) vx_mark_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    __shared__ VoxDimsSmem S;
    __shared__ uint32_t s_wsum[8];
    __shared__ uint32_t s_base;
    const uint32_t n_total = A.n[0] + A.n[1];
    const uint32_t per = (n_total + kVxIlp - 1) / kVxIlp;
    const uint32_t t = A.mark_lo + blockIdx.x * blockDim.x + threadIdx.x;
    uint2 p[kVxIlp];
    int c[kVxIlp];
    bool on[kVxIlp];
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {                        // (the coordinates are on their way while thread 0 plans the grids)
        const uint32_t i = t < A.mark_hi && t < per ? t + (uint32_t)k * per : 0xFFFFFFFFu;
        on[k] = i < n_total;
        uint32_t li;
        if (on[k]) {
            p[k] = vx_point(A, i, c[k], li);
            if (!A.sel) A.vkey[i] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); // neutral element of the place pass's atomicMin
        }
    }
    vx_block_dims(A, S);
    if (S.status) return;
    bool in[kVxIlp];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {
        in[k] = on[k] && vx_in_slab(S.sp, p[k]);
        mine += in[k] ? 1u : 0u;
    }
    if (A.sel) {
        // split pair: the points of this rank's slab go to a compacted list (block-wise: one atomic per block), the
        // others get "no voxel" as their rank right here
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t off = incl - mine, total = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) off += s_wsum[w]; total += s_wsum[w]; }
        if (threadIdx.x == 0) s_base = total ? atomicAdd(A.sel_count, total) : 0u;
        __syncthreads();
        off += s_base;
#pragma unroll
        for (int k = 0; k < kVxIlp; ++k) {
            const uint32_t i = t + (uint32_t)k * per;
            if (in[k]) A.sel[off++] = i;
            else if (on[k]) A.prank[i] = kVxNone;
        }
    }
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k)
        if (in[k]) {
            VX_UNPACK(p[k], x, y, z);
#if PCCM_DIR_BYTES
            A.dirbytes[(size_t)S.dir_off[c[k]] * 32 + vx_key(S.g[c[k]], x, y, z)] = 1;
#else
            vx_mark_point(A.dirbits + S.dir_off[c[k]], vx_key(S.g[c[k]], x, y, z));
#endif
        }
}

template <int THREADS>
__device__ __forceinline__ uint32_t vx_block_sum_u32(uint32_t v, uint32_t* sm) {   // result in every thread
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) r += sm[w];
    return r;
}

// popcount sum of every chunk of kVxDirChunk directory words
__global__ void __launch_bounds__(256) vx_dirsum_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    __shared__ VoxDimsSmem S;
    __shared__ uint32_t sm[8];
    vx_block_dims(A, S);
    if (S.status) return;
    const uint32_t nw = S.dir_off[1] + S.ndirw[1];
    const uint32_t w0 = blockIdx.x * kVxDirChunk;
    if (w0 >= nw) return;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < kVxDirChunk / 256; ++k) {
        const uint32_t w = w0 + k * 256 + threadIdx.x;
        if (w < nw) {
#if PCCM_DIR_BYTES
            const uint4* b = reinterpret_cast<const uint4*>(A.dirbytes + (size_t)w * 32);
            const uint4 lo = __ldg(b), hi = __ldg(b + 1);
            auto bits4 = [](uint32_t v) { return ((v & 1u) | ((v >> 7) & 2u) | ((v >> 14) & 4u) | ((v >> 21) & 8u)); };   // bytes are 0 or 1
            const uint32_t m = bits4(lo.x) | (bits4(lo.y) << 4) | (bits4(lo.z) << 8) | (bits4(lo.w) << 12) |
                               (bits4(hi.x) << 16) | (bits4(hi.y) << 20) | (bits4(hi.z) << 24) | (bits4(hi.w) << 28);
            A.dirbits[w] = m;
            acc += (uint32_t)__popc(m);
#else
            acc += (uint32_t)__popc(A.dirbits[w]);
#endif
        }
    }
    acc = vx_block_sum_u32<256>(acc, sm);
    if (threadIdx.x == 0) A.dirsums[blockIdx.x] = acc;
}

// exclusive popcount prefix of the directory (dirpre), the VoxPlan, and the zeroing of the occupancy words of
// the bricks that exist (+ the empty one past the last)
__global__ void __launch_bounds__(256) vx_dirscan_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    __shared__ VoxDimsSmem S;
    __shared__ uint32_t sm[8];
    __shared__ uint32_t s_warp[8];
    vx_block_dims(A, S);
    const uint32_t nw = S.status ? 0u : S.dir_off[1] + S.ndirw[1];
    const uint32_t nchunks = (nw + kVxDirChunk - 1) / kVxDirChunk;
    const uint32_t split = S.status ? 0u : S.dir_off[1] / kVxDirChunk;     // chunks of cloud 0
    uint32_t before = 0, total = 0, first = 0;
    for (uint32_t j = threadIdx.x; j < nchunks; j += 256) {
        const uint32_t v = A.dirsums[j];
        total += v;
        if (j < blockIdx.x) before += v;
        if (j < split) first += v;
    }
    before = vx_block_sum_u32<256>(before, sm);
    total = vx_block_sum_u32<256>(total, sm);
    first = vx_block_sum_u32<256>(first, sm);
    uint32_t status = S.status;
    if (!status && (unsigned long long)total + 1ull > (unsigned long long)A.cap_blk) status |= kVxStBrickOverflow;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        VoxPlan P;
        for (int c = 0; c < 2; ++c) {
            VoxView& V = P.view[c];
            V.g = S.g[c];
            V.dirbits = A.dirbits + S.dir_off[c];
            V.dirpre = A.dirpre + S.dir_off[c];
            V.rows = A.rows; V.vxyz = A.vxyz; V.vkey = A.vkey;
            V.prank = A.prank + (c ? A.n[0] : 0u);
            V.slot0 = c ? first : 0u;
            V.nblk = c ? total - first : first;
            V.n = A.n[c];
            V.nblk_total = total;
            V.n_total = A.n[0] + A.n[1];
            P.ndirw[c] = S.ndirw[c]; P.dir_off[c] = S.dir_off[c];
            for (int a = 0; a < 3; ++a) { P.mn[c][a] = S.mn[c][a]; P.mx[c][a] = S.mx[c][a]; }
        }
        P.shard = S.sp;
        P.status = status;
        P.nvox_total = 0;
        P.ndirw_total = nw;
        P.cap_dirw = A.cap_dirw; P.cap_blk = A.cap_blk;
        *A.plan = P;
    }
    if (status) return;
    // dirpre of this block's chunk: thread t owns 8 consecutive words
    const uint32_t w0 = blockIdx.x * kVxDirChunk + threadIdx.x * 8;
    uint32_t cnt[8], mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        cnt[k] = (w0 + k < nw) ? (uint32_t)__popc(A.dirbits[w0 + k]) : 0u;
        mine += cnt[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t run = before + incl - mine;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (w0 + k <= nw) A.dirpre[w0 + k] = run;            // (entry nw = the number of bricks)
        run += cnt[k];
    }
    // occupancy words of bricks [0, total]: zero (the brick past the last one stays empty)
    const size_t n4 = ((size_t)total + 1) * kVxRows / 2;
    uint4* m4 = reinterpret_cast<uint4*>(A.rows);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) m4[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(256) vx_fill_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    const VoxPlan* __restrict__ P = A.plan;
    if (P->status) return;
    const uint32_t n_total = A.n[0] + A.n[1];
    uint2 p[kVxIlp];
    int c[kVxIlp];
    uint32_t idx[kVxIlp], slot[kVxIlp], key[kVxIlp];
    bool on[kVxIlp];
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {
        idx[k] = vx_ilp_index(n_total, k);
        on[k] = idx[k] < n_total;
        uint32_t li;
        if (on[k]) p[k] = vx_point(A, idx[k], c[k], li);
    }
    const ShardPlan sp = P->shard;
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {
        on[k] = on[k] && vx_in_slab(sp, p[k]);
        if (on[k]) {
            VX_UNPACK(p[k], x, y, z);
            const VoxView& V = P->view[c[k]];
            key[k] = vx_key(V.g, x, y, z);
            slot[k] = vx_slot_of_key(V.dirbits, V.dirpre, key[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k)
        if (on[k]) {
            VX_UNPACK(p[k], x, y, z);
            VX_CHECK(slot[k] < P->view[0].nblk_total);
            vx_fill_point(A.rows, slot[k], x, y, z);
            A.pslot[idx[k]] = slot[k];
            if (A.bkey) A.bkey[slot[k]] = key[k];
        }
}

// split pairs: the same two passes over the compacted list of this rank's points (grid-stride: its length is only
// known on the device)
__global__ void __launch_bounds__(256) vx_fill_sel_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    const VoxPlan* __restrict__ P = A.plan;
    if (P->status) return;
    const uint32_t nsel = *A.sel_count, stride = gridDim.x * blockDim.x;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nsel; w += stride) {
        const uint32_t i = A.sel[w];
        int c;
        uint32_t li;
        const uint2 p = vx_point(A, i, c, li);
        VX_UNPACK(p, x, y, z);
        const VoxView& V = P->view[c];
        const uint32_t key = vx_key(V.g, x, y, z);
        const uint32_t slot = vx_slot_of_key(V.dirbits, V.dirpre, key);
        VX_CHECK(slot < V.nblk_total);
        vx_fill_point(A.rows, slot, x, y, z);
        A.pslot[w] = slot;
        if (A.bkey) A.bkey[slot] = key;
        A.vkey[w] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);       // (there are at most nsel voxels)
    }
}
__global__ void __launch_bounds__(256) vx_place_sel_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    const VoxPlan* __restrict__ P = A.plan;
    if (P->status) return;
    const uint32_t nsel = *A.sel_count, stride = gridDim.x * blockDim.x;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nsel; w += stride) {
        const uint32_t i = A.sel[w];
        int c;
        uint32_t li;
        const uint2 p = vx_point(A, i, c, li);
        VX_UNPACK(p, x, y, z);
        const uint32_t rgba = A.rgb[c] ? (__ldg(static_cast<const uint32_t*>(A.rgb[c]) + li) & 0xffffffu) : 0u;
        A.prank[i] = vx_place_point(A.rows, A.bkey ? nullptr : A.vxyz, A.vkey, A.pslot[w], x, y, z, rgba, li);
    }
}

// voxels per chunk of kVxBrickChunk bricks (the brick past the last one counts as an empty brick).  A warp owns 8
// bricks of the chunk: one coalesced 256-byte load each, all eight in flight together.
__device__ __forceinline__ void vx_load_bricks8(const uint2* rows, uint32_t s0, uint32_t nb, int lane, uint2 m[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {              // lane l: rows 2l and 2l + 1 (one 16-byte load), occupancy words only
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (s0 + j < nb) q = __ldg(reinterpret_cast<const uint4*>(rows + (size_t)(s0 + j) * kVxRows) + lane);
        m[j] = make_uint2(q.x, q.z);
    }
}

__global__ void __launch_bounds__(256) vx_bricksum_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    __shared__ uint32_t sm[8];
    const VoxPlan* __restrict__ P = A.plan;
    if (P->status) return;
    const uint32_t nb = P->view[0].nblk_total + 1;
    const uint32_t nchunks = (nb + kVxBrickChunk - 1) / kVxBrickChunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        uint2 m[8];
        vx_load_bricks8(A.rows, ch * kVxBrickChunk + warp * 8, nb, lane, m);
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += (uint32_t)(__popc(m[j].x) + __popc(m[j].y));
        acc = vx_block_sum_u32<256>(acc, sm);
        if (threadIdx.x == 0) A.bricksums[ch] = acc;
    }
}

// rank of the first voxel of every row of every brick.  Block b owns a contiguous run of chunks: one strided
// sum of the chunk sums before its run, then chunk after chunk with a running offset.
__global__ void __launch_bounds__(256) vx_rowbase_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    __shared__ uint32_t sm[8];
    __shared__ uint32_t s_wtot[8];
    VoxPlan* P = A.plan;
    if (P->status) return;
    const uint32_t nb = P->view[0].nblk_total + 1;
    const uint32_t nchunks = (nb + kVxBrickChunk - 1) / kVxBrickChunk;
    const uint32_t c0 = (uint32_t)((unsigned long long)nchunks * blockIdx.x / gridDim.x);
    const uint32_t c1 = (uint32_t)((unsigned long long)nchunks * (blockIdx.x + 1) / gridDim.x);
    if (c0 >= c1) return;
    uint32_t off = 0;
    for (uint32_t j = threadIdx.x; j < c0; j += 256) off += A.bricksums[j];
    off = vx_block_sum_u32<256>(off, sm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ord_slot = kVxNone, ord_pos = 0;                // a brick of this lane waiting for its place in the size-class list
    size_t ord_at = 0;
    for (uint32_t ch = c0; ch < c1; ++ch) {
        const uint32_t s0 = ch * kVxBrickChunk + warp * 8;
        uint2 m[8];
        vx_load_bricks8(A.rows, s0, nb, lane, m);
        // inclusive prefix over the rows of each brick (lane l owns rows 2l, 2l+1); brick totals in lane 31
        uint32_t in2[8], wtot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t v = (uint32_t)(__popc(m[j].x) + __popc(m[j].y));
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            in2[j] = v;
        }
        uint32_t bbase[8];                                   // base of brick j inside this warp's run
        uint32_t mytot = 0;                                  // lane j < 8: voxels of brick j
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            bbase[j] = wtot;
            const uint32_t tot = __shfl_sync(0xffffffffu, in2[j], 31);
            wtot += tot;
            mytot = lane == j ? tot : mytot;
        }
        // list of the bricks by size class: the search kernel hands out the large bricks first, so that no warp
        // starts a 500-voxel brick when the others are about to finish (the order inside a class does not matter)
        // (the position comes back from the L2 while the chunk is finished: it is stored at the top of the next trip)
        if (ord_slot != kVxNone) A.border[ord_at + ord_pos] = ord_slot;
        ord_slot = kVxNone;
        const bool isbrick = lane < 8 && s0 + lane + 1 < nb;   // (entry nb - 1 is the end marker, not a brick)
        const unsigned bm = __ballot_sync(0xffffffffu, isbrick);
        if (isbrick) {
            // one atomic per size class present among the warp's eight bricks (neighbouring bricks are of similar size;
            // a few counters taking one atomic per brick of the pair serialise at the L2)
            const uint32_t cls = min(mytot >> 5, (uint32_t)kVxSizeClasses - 1u);
            const unsigned peers = __match_any_sync(bm, cls);
            const int leader = __ffs((int)peers) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(A.bcursor + cls, (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            ord_at = (size_t)cls * A.cap_blk;
            ord_pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            ord_slot = s0 + lane;
        }
        __syncthreads();                                     // (s_wtot of the previous chunk has been read)
        if (lane == 0) s_wtot[warp] = wtot;
        __syncthreads();
        uint32_t wbase = off, chunk_total = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) wbase += s_wtot[w];
            chunk_total += s_wtot[w];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (s0 + j >= nb) break;
            const uint32_t a = (uint32_t)__popc(m[j].x), b = (uint32_t)__popc(m[j].y);
            const uint32_t base = wbase + bbase[j] + in2[j] - a - b;
            reinterpret_cast<uint4*>(A.rows + (size_t)(s0 + j) * kVxRows)[lane] = make_uint4(m[j].x, base, m[j].y, base + a);
            if (A.bkey && s0 + j + 1 < nb) {
                // coordinates of the brick's voxels, in rank order (a lane's two rows are neighbours in rank, the lanes of
                // a warp follow each other: the stores of a brick fill consecutive sectors).  Measured against the whole
                // warp taking the non-empty rows one after the other, one lane per bit (a coalesced store per row): this
                // per-lane walk is faster (index build of the 10 M pair 1.53 vs 1.81 ms, of the 1 M pair 0.138 vs 0.195)
                const int cl = s0 + j >= P->view[1].slot0 ? 1 : 0;
                const VoxDims g = P->view[cl].g;
                const uint32_t key = __ldg(A.bkey + s0 + j);
                const uint32_t kx = key % (uint32_t)g.nbx, kyz = key / (uint32_t)g.nbx;
                const uint32_t x0 = (kx + (uint32_t)g.obx) << 5, y0 = (kyz % (uint32_t)g.nby + (uint32_t)g.oby) << 3, z0 = (kyz / (uint32_t)g.nby + (uint32_t)g.obz) << 3;
                uint32_t r = base;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t w = h ? m[j].y : m[j].x;
                    const uint32_t row = 2u * (uint32_t)lane + (uint32_t)h;
                    const uint32_t yz_lo = (y0 + (row & 7u)) << 16, zz = z0 + (row >> 3);
                    while (w) {
                        const uint32_t b = (uint32_t)__ffs((int)w) - 1u;
                        w &= w - 1u;
                        A.vxyz[r++] = make_uint2((x0 + b) | yz_lo, zz);
                    }
                }
            }
        }
        off += chunk_total;
    }
    if (ord_slot != kVxNone) A.border[ord_at + ord_pos] = ord_slot;
    if (c1 == nchunks && threadIdx.x == 0) P->nvox_total = off;
}

__global__ void __launch_bounds__(256) vx_place_kernel(const __grid_constant__ VoxBuildArgs A) {
    pdl_enter();
    const VoxPlan* __restrict__ P = A.plan;
    if (P->status) return;
    const uint32_t n_total = A.n[0] + A.n[1];
    const ShardPlan sp = P->shard;
    uint32_t idx[kVxIlp], li[kVxIlp], slot[kVxIlp], rgba[kVxIlp];
    uint2 p[kVxIlp];
    bool on[kVxIlp];
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k) {
        idx[k] = vx_ilp_index(n_total, k);
        on[k] = idx[k] < n_total;
        if (on[k]) {
            int c;
            p[k] = vx_point(A, idx[k], c, li[k]);
            if (!vx_in_slab(sp, p[k])) {                    // not indexed on this rank: no voxel, no query
                A.prank[idx[k]] = kVxNone;
                on[k] = false;
                continue;
            }
            slot[k] = A.pslot[idx[k]];
            rgba[k] = A.rgb[c] ? (__ldg(static_cast<const uint32_t*>(A.rgb[c]) + li[k]) & 0xffffffu) : 0u;   // colours that have already arrived ride in vkey
        }
    }
#pragma unroll
    for (int k = 0; k < kVxIlp; ++k)
        if (on[k]) {
            VX_UNPACK(p[k], x, y, z);
            VX_CHECK(slot[k] < P->view[0].nblk_total);
            const uint32_t rank = vx_place_point(A.rows, A.bkey ? nullptr : A.vxyz, A.vkey, slot[k], x, y, z, rgba[k], li[k]);
            VX_CHECK(rank < n_total);
            A.prank[idx[k]] = rank;
        }
}

// float64 colours -> uchar4, checking on the way that every channel is k / 255 (else the plan's status says so and
// the caller repeats the evaluation with float64 colour arrays)
__global__ void pack_rgb_u8_check_kernel(const void* rgb, int64_t stride, int64_t n, uchar4* out, uint32_t* status) {
    pdl_enter();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool bad = false;
    uint32_t v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double c = load_coord(rgb, PCCM_F64, stride, i, a);
        const double k = rint(c * 255.0);
        if (!(k >= 0.0 && k <= 255.0 && k / 255.0 == c)) bad = true;
        v[a] = (uint32_t)(k >= 0.0 && k <= 255.0 ? k : 0.0);
    }
    out[i] = make_uchar4((unsigned char)v[0], (unsigned char)v[1], (unsigned char)v[2], 0);
    if (bad) atomicOr(status, kVxStRgbNotU8);
}

// ------------------------------------------------------------------------------------
// query
// ------------------------------------------------------------------------------------
struct VxDir {
    int32_t qc, sc;            // which view of the plan is the query / the search cloud
    uint32_t nq;               // points of the query cloud
    const uint32_t* qprank;    // [nq]
    CloudView qa, sa;          // attribute views (colours, normals) of the query / search cloud
    uint32_t flags;
    int32_t* idx_out;          // original query order, or null
    double* d2_out;
    uint32_t* todo;            // ranked positions of the voxels the staged search left undecided
    uint32_t* far;             // ... and of those the brick rings left undecided (pencil search)
    RowGrid sgrid;             // pencil index of the search cloud (vx_far_kernel only)
    const uint4* srecs;
    const uint32_t* srow_start;
    uint32_t rec_off;          // first reduction record of this direction
    uint32_t ntiles;           // blocks (= reduction records) of this direction in one vx_epilogue_kernel pass
    uint32_t npt_tiles;        // ceil(nq / kVxEpiTile): tiles of points, dealt round-robin to those blocks
};

struct VxParams {
    const VoxPlan* plan;
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
    const uint32_t* bcursor;   // [kVxSizeClasses] bricks per size class and
    const uint32_t* border;    // [kVxSizeClasses][border_cap] their slots (vx_rowbase_kernel)
    uint32_t border_cap;
    uint32_t* counters;        // [0..1] undecided per direction, [2..3] far per direction, [4] brick ticket of the search kernel
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

// epilogue of ONE query point: D1 (+ per-point outputs), D2 with the other cloud's normals, colour.
__device__ __forceinline__ void vx_epilogue(const VxParams& P, const VxDir& D, const CloudView& qa, const CloudView& sa,
                                            uint32_t qidx, uint32_t qrgb, uint32_t d2,
                                            int ex, int ey, int ez, uint32_t nidx, uint32_t nrgb, VxAcc& a,
                                            const double* staged_normal = nullptr) {
    a.s1 += d2;
    a.m1 = d2 > a.m1 ? d2 : a.m1;
    a.cnt++;
    if (D.idx_out) D.idx_out[qidx] = (int32_t)nidx;
    if (D.d2_out) D.d2_out[qidx] = (double)d2;
    if (D.flags & PCCM_EVAL_D2) {
        const double e[3] = {(double)ex, (double)ey, (double)ez};
        double nv[3];
        const uint32_t ni = P.normals_mode == PCCM_NORMALS_BY_NEIGHBOUR ? nidx : qidx;
        if (staged_normal) { nv[0] = staged_normal[0]; nv[1] = staged_normal[1]; nv[2] = staged_normal[2]; }   // (shared memory: the tile's bulk copy)
        else { nv[0] = __ldg(sa.normals + 3 * (size_t)ni); nv[1] = __ldg(sa.normals + 3 * (size_t)ni + 1); nv[2] = __ldg(sa.normals + 3 * (size_t)ni + 2); }
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
// 2l+1 of every brick (one coalesced load of the 64 occupancy words).  The bricks are visited NEAREST
// FIRST (by the distance from the query to the brick's box, four bricks in flight per step) and the
// visit stops at the first brick whose box is farther than the best distance found: a voxel whose
// neighbour is 3 .. 8 voxels away touches a handful of the 27 bricks, not all of them.  Pass 1:
// minimal squared distance by bit scans only.  Pass 2 (when that distance is below `limit`, i.e.
// certified): the smallest original index among the voxels at that distance, in the bricks whose box
// is not farther than it.  Returns the distance; rank is valid when it is below `limit`.
template <int DIM, bool SELF = false>
__device__ __forceinline__ uint32_t vx_warp_bricks(const VoxView& S, const int* slots, const int* ids, int n,
                                                   int qx, int qy, int qz, uint32_t limit, uint32_t& rank_out) {
    // SELF: the query is a voxel of S itself -- its own bit is cleared and only the distance is wanted
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int qbx = qx >> 5, qby = qy >> 3, qbz = qz >> 3;
    uint32_t best = kVxNone;
    constexpr int kBatch = 4;                      // bricks in flight per step (independent coalesced loads; 8 spills)
    constexpr int kOwn = (DIM * DIM * DIM + 31) / 32;      // list entries per lane: entry k belongs to lane k % 32
    // squared distance from the query to the box of my entries (kVxNone: no entry)
    uint32_t gap[kOwn];
#pragma unroll
    for (int o = 0; o < kOwn; ++o) {
        const int k = o * 32 + lane;
        gap[o] = kVxNone;
        if (k < n) {
            const int b = ids[k];
            const int bx = qbx + b % DIM - DIM / 2, by = qby + (b / DIM) % DIM - DIM / 2, bz = qbz + b / (DIM * DIM) - DIM / 2;
            const int gx = vx_gap(qx, bx << 5, (bx << 5) + 31), gy = vx_gap(qy, by << 3, (by << 3) + 7), gz = vx_gap(qz, bz << 3, (bz << 3) + 7);
            gap[o] = (uint32_t)(gx * gx + gy * gy + gz * gz);
        }
    }
    uint32_t left[kOwn];                           // my entries not visited yet
#pragma unroll
    for (int o = 0; o < kOwn; ++o) left[o] = gap[o];
    for (;;) {
        // the (up to) kBatch nearest unvisited entries: key = gap << 8 | entry
        int sel[kBatch];
        int nsel = 0;
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            uint32_t key = kVxNone;
#pragma unroll
            for (int o = 0; o < kOwn; ++o)
                if (left[o] != kVxNone) { const uint32_t kk = (left[o] << 8) | (uint32_t)(o * 32 + lane); key = kk < key ? kk : key; }
            key = __reduce_min_sync(full, key);
            sel[j] = -1;
            if (key != kVxNone && (key >> 8) <= best) {          // (a box farther than the best cannot hold anything nearer or tied)
                sel[j] = (int)(key & 0xffu);
                ++nsel;
#pragma unroll
                for (int o = 0; o < kOwn; ++o)
                    if (o * 32 + lane == sel[j]) left[o] = kVxNone;
            }
        }
        if (!nsel) break;
        uint2 mb[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (sel[j] >= 0) q = __ldg(reinterpret_cast<const uint4*>(S.rows + (size_t)slots[sel[j]] * kVxRows) + lane);
            mb[j] = make_uint2(q.x, q.z);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            if (sel[j] < 0) continue;
            const int b = ids[sel[j]];
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
        best = __reduce_min_sync(full, best);
    }
    if (SELF || best >= limit) return best;
    uint32_t bidx = kVxNone, brank = kVxNone;
#pragma unroll
    for (int o = 0; o < kOwn; ++o) {
        unsigned cand = __ballot_sync(full, gap[o] <= best);       // entries whose box can hold a voxel at the minimal distance
        while (cand) {
            const int k0 = o * 32 + __ffs((int)cand) - 1;
            cand &= cand - 1u;
            const int slot = slots[k0], b = ids[k0];
            const int bx = qbx + b % DIM - DIM / 2, by = qby + (b / DIM) % DIM - DIM / 2, bz = qbz + b / (DIM * DIM) - DIM / 2;
            const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(S.rows + (size_t)slot * kVxRows) + lane);
            const uint2 m2 = make_uint2(q2.x, q2.z);
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
                    const uint32_t i = __ldg(&S.vkey[rank].y);
                    if (i < bidx) { bidx = i; brank = rank; }
                }
                if (byz + (uint32_t)(dhi * dhi) == best) {
                    const uint32_t rank = vx_rank(S, (uint32_t)slot, r, (p + dhi) & 31);
                    const uint32_t i = __ldg(&S.vkey[rank].y);
                    if (i < bidx) { bidx = i; brank = rank; }
                }
            }
        }
    }
    const uint32_t widx = __reduce_min_sync(full, bidx);
    const int src = __ffs((int)__ballot_sync(full, bidx == widx)) - 1;
    rank_out = __shfl_sync(full, brank, src);
    return best;
}

// rank of candidate bit b of a 27-neighbourhood (see vx_cand)
__device__ __forceinline__ uint32_t vx_cand_rank27(const VoxView& S, const int* sslot, const uint2* win, const uint32_t* rb,
                                                   int lx, int ly, int lz, int b) {
    const int j = b / 3, dx = b - 3 * j - 1, dz = j / 3 - 1, dy = j - 3 * (dz + 1) - 1;
    const int wi = (lz + dz) * kVxRegY + (ly + dy);
    const int cxl = lx + dx;
    if ((unsigned)cxl < 32u) {
        const uint32_t c = vx_centre_word(win[wi]);
        return rb[wi] + (uint32_t)__popc(c & ((1u << cxl) - 1u));
    }
    const int ry = ly + dy, rz = lz + dz;
    const int ny_i = ry < 2 ? 0 : (ry < 10 ? 1 : 2), nz_i = rz < 2 ? 0 : (rz < 10 ? 1 : 2);
    return vx_rank(S, (uint32_t)sslot[nz_i * 9 + ny_i * 3 + (cxl < 0 ? 0 : 2)], (((rz + 6) & 7) << 3) | ((ry + 6) & 7), cxl & 31);
}
__device__ __forceinline__ uint32_t vx_pack_e27(int b) {     // packed (query - neighbour) of candidate bit b
    const int j = b / 3, dx = b - 3 * j - 1, dz = j / 3 - 1, dy = j - 3 * (dz + 1) - 1;
    return vx_pack_e(-dx, -dy, -dz);
}

// SEARCH.  One warp = one brick of the query cloud at a time (bricks are handed out by a ticket counter).  The warp
// stages the search cloud's occupancy rows around the brick (12 x 12 rows x 64 bits) and the rank bases of the
// centre column in its private slice of shared memory, then takes the brick's voxels 32 at a time, one lane per
// voxel: 27-bit neighbourhood, distance level, and -- only when the evaluation needs the neighbour itself -- the
// look-up of the voxels that tie at the minimum, two at a time.  The few voxels with an empty neighbourhood are
// collected per brick and finished together: 125-bit neighbourhood, then whole-brick scans of the 27 bricks.
// Integer work only; the answer of every voxel goes to vres[] (16 bytes), undecided voxels to the todo list.
#ifndef PCCM_VX_THREADS
#define PCCM_VX_THREADS 128
#endif
constexpr int kVxThreads = PCCM_VX_THREADS;
constexpr int kVxWarps = kVxThreads / 32;

#ifndef PCCM_VX_MINBLOCKS
#define PCCM_VX_MINBLOCKS 10
#endif

constexpr int kVxPend = 64;
struct VxWarpSmem {
    uint2 win[kVxRegRows];
    uint32_t rb[kVxRegRows];
    int sslot[28];
    int occ_slot[28];
    int occ_id[28];
    uint16_t pend[kVxPend];
};

// the voxels of a brick that the 27-neighbourhood left open (lane l takes pend[l]): 125-neighbourhood, whole-brick
// scans of the 27 neighbour bricks, todo list
__device__ __forceinline__ void vx_finish_pending(const VxParams& P, int d, const VoxView& Q, const VoxView& S,
                                                  VxWarpSmem& W, uint32_t b0, int count, bool any_brick, int nocc) {
    const unsigned full = 0xffffffffu;
    const VxDir& D = P.dir[d];
    const int lane = threadIdx.x & 31;
    const bool active = lane < count;
    const uint32_t t = b0 + (active ? (uint32_t)W.pend[lane] : 0u);
    const uint2 qr = __ldg(Q.vxyz + t);
    const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
    uint4 r = make_uint4(kVxNone, 0u, 0u, 0u);
    bool done = false;
    if (active && any_brick) {
        VxPick pk;
        const uint32_t bd2 = vx_search125(S, W.sslot, W.win, W.rb, qx & 31, (qy & 7) + 2, (qz & 7) + 2, pk);
        if (bd2 < 9u) { r = make_uint4(bd2, vx_pack_e(pk.ex, pk.ey, pk.ez), pk.idx, pk.rgb); done = true; }
    }
    unsigned pend = any_brick ? __ballot_sync(full, active && !done) : 0u;
    while (pend) {                 // nearest point 3+ voxels away: the whole warp scans the 27 neighbour bricks for one voxel
        const int src = __ffs((int)pend) - 1;
        pend &= pend - 1u;
        const int sx = __shfl_sync(full, qx, src), sy = __shfl_sync(full, qy, src), sz = __shfl_sync(full, qz, src);
        uint32_t nrank = kVxNone;
        const uint32_t nd2 = vx_warp_bricks<3>(S, W.occ_slot, W.occ_id, nocc, sx, sy, sz, 81u, nrank);
        if (nd2 < 81u && lane == src) {          // anything outside the 27 bricks is at least 9 voxels away
            const uint2 nc = __ldg(S.vxyz + nrank), nk = __ldg(S.vkey + nrank);
            r = make_uint4(nd2, vx_pack_e(qx - (int)(nc.x & 0xffffu), qy - (int)(nc.x >> 16), qz - (int)nc.y), nk.y, nk.x);
            done = true;
        }
    }
    if (active) P.vres[t] = r;
    const unsigned und = __ballot_sync(full, active && !done);
    if (und) {
        uint32_t pos = 0;
        if (lane == 0) pos = atomicAdd(P.counters + d, (uint32_t)__popc(und));
        pos = __shfl_sync(full, pos, 0);
        VX_CHECK(pos + (uint32_t)__popc(und) <= D.nq);
        if (active && !done) D.todo[pos + __popc(und & ((1u << lane) - 1u))] = t;
    }
}

#if defined(PCCM_VX_TRACE)
// trace builds (tools/search_trace.py): per warp of the last vx_search_kernel launch {start, end (globaltimer ns), bricks, voxels}
constexpr int kVxTraceWarps = 16384;
__device__ unsigned long long g_vx_trace[kVxTraceWarps][4];
__device__ __forceinline__ unsigned long long vx_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

__global__ void __launch_bounds__(kVxThreads, PCCM_VX_MINBLOCKS)
vx_search_kernel(const __grid_constant__ VxParams P) {
    pdl_enter();
#if defined(PCCM_VX_TRACE)
    const unsigned long long tr_t0 = vx_now();
    unsigned long long tr_bricks = 0, tr_vox = 0;
    struct TraceOut {
        unsigned long long t0; unsigned long long* b; unsigned long long* v;
        __device__ ~TraceOut() {
            const uint32_t w = blockIdx.x * (uint32_t)kVxWarps + (threadIdx.x >> 5);
            if ((threadIdx.x & 31) == 0 && w < (uint32_t)kVxTraceWarps) { g_vx_trace[w][0] = t0; g_vx_trace[w][1] = vx_now(); g_vx_trace[w][2] = *b; g_vx_trace[w][3] = *v; }
        }
    } tr_out{tr_t0, &tr_bricks, &tr_vox};
#endif
    __shared__ VxWarpSmem s_w[kVxWarps];
    const unsigned full = 0xffffffffu;
    const VoxPlan* __restrict__ plan = P.plan;
    if (plan->status) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    VxWarpSmem& W = s_w[warp];
    // work list: every brick of the pair, largest size class first (bricks of a cloud that is not queried are skipped)
    uint32_t cls_cnt = 0, cls_incl = 0;
    if (lane < kVxSizeClasses) cls_cnt = __ldg(P.bcursor + (kVxSizeClasses - 1 - lane));
    cls_incl = cls_cnt;
#pragma unroll
    for (int o = 1; o < kVxSizeClasses; o <<= 1) {
        const uint32_t t = __shfl_up_sync(full, cls_incl, o);
        if (lane >= o) cls_incl += t;
    }
    const uint32_t nwork = __shfl_sync(full, cls_incl, kVxSizeClasses - 1);
#ifndef PCCM_VX_TICKET
#define PCCM_VX_TICKET 1          // bricks per ticket (measured: 4 per ticket costs 45 % -- the tail of the kernel is one ticket long)
#endif
    // The first brick of a warp is its own number (no ticket: thousands of warps asking one counter at once would all
    // stand in that queue before any of them starts); the others come from the ticket counter, which therefore counts
    // from the number of warps.  The next ticket is requested before the current brick is searched.
    const uint32_t nwarps = gridDim.x * (uint32_t)kVxWarps;
    uint32_t ticket = 0, gw = blockIdx.x * (uint32_t)kVxWarps + (uint32_t)warp, gw_end = gw + 1u;
    if (lane == 0) ticket = nwarps + atomicAdd(P.counters + 4, (uint32_t)PCCM_VX_TICKET);
    for (; gw < nwork; ++gw) {
        if (gw == gw_end) {
            gw = __shfl_sync(full, ticket, 0);
            if (gw >= nwork) break;
            gw_end = min(gw + (uint32_t)PCCM_VX_TICKET, nwork);
            if (lane == 0) ticket = nwarps + atomicAdd(P.counters + 4, (uint32_t)PCCM_VX_TICKET);   // the next ticket travels while these bricks are searched
        }
        const int ci = __popc(__ballot_sync(full, lane < kVxSizeClasses && gw >= cls_incl));      // classes before the ticket's
        const uint32_t cfirst = __shfl_sync(full, cls_incl - cls_cnt, ci);
        const uint32_t slot = __ldg(P.border + (size_t)(kVxSizeClasses - 1 - ci) * P.border_cap + (gw - cfirst));
        int d = 0;
        {
            const VoxView& Q0 = plan->view[P.dir[0].qc];
            if (slot - Q0.slot0 >= Q0.nblk) {
                if (P.ndirs < 2) continue;
                const VoxView& Q1 = plan->view[P.dir[1].qc];
                if (slot - Q1.slot0 >= Q1.nblk) continue;
                d = 1;
            }
        }
        const VxDir& D = P.dir[d];
        const VoxView& Q = plan->view[D.qc];
        const VoxView& S = plan->view[D.sc];
        const uint32_t b0 = vx_brick_begin(Q, slot), b1 = vx_brick_begin(Q, slot + 1);
        uint32_t t_lo, t_hi;
        vx_slice(P, Q, t_lo, t_hi);
        const uint32_t t0 = max(b0, t_lo), t1 = min(b1, t_hi);
        if (t0 >= t1) continue;
#if defined(PCCM_VX_TRACE)
        tr_bricks++; tr_vox += t1 - t0;
#endif
        const bool need_idx = D.idx_out != nullptr || (D.flags & PCCM_EVAL_COLOR) ||
                              ((D.flags & PCCM_EVAL_D2) && P.normals_mode == PCCM_NORMALS_BY_NEIGHBOUR);
        const bool need_e = (D.flags & PCCM_EVAL_D2) != 0;
        const uint2* __restrict__ qxyz = Q.vxyz;
        const uint2 first = __ldg(qxyz + b0);
        const int bx = (int)(first.x & 0xffffu) >> 5, by = (int)(first.x >> 16) >> 3, bz = (int)first.y >> 3;
        if (bz < plan->shard.own_zlo || bz >= plan->shard.own_zhi) {       // halo brick of a split pair: another rank's queries
            for (uint32_t t = t0 + lane; t < t1; t += 32) P.vres[t] = make_uint4(kVxNone, 0u, 0u, 0u);
            continue;
        }
        __syncwarp();                                        // (the previous brick's window is no longer read)
        int myslot = -1;
        if (lane < 27) {
            myslot = vx_slot(S, bx + lane % 3 - 1, by + (lane / 3) % 3 - 1, bz + lane / 9 - 1);
            W.sslot[lane] = myslot;
        }
        const unsigned occ = __ballot_sync(full, myslot >= 0);
        const bool any_brick = occ != 0u;
        const int nocc = __popc(occ);
        if (myslot >= 0) {                       // compacted list of the occupied neighbour bricks (whole-brick scans)
            const int k = __popc(occ & ((1u << lane) - 1u));
            W.occ_slot[k] = myslot;
            W.occ_id[k] = lane;
        }
        __syncwarp();
        if (any_brick) {
            for (int i = lane; i < kVxRegRows; i += 32) {
                uint32_t rb;
                W.win[i] = vx_stage_row(S, W.sslot, i, rb);
                W.rb[i] = rb;
            }
            __syncwarp();
        }
        int npend = 0;
        for (uint32_t tb = t0; tb < t1; tb += 32) {
            const uint32_t t = tb + lane;
            const bool active = t < t1;
            const uint2 qr = __ldg(qxyz + (active ? t : t0));
            const int lx = (int)(qr.x & 31u), ly = (int)((qr.x >> 16) & 7u) + 2, lz = (int)(qr.y & 7u) + 2;
            uint32_t bd2 = 0, cand = 0;
            if (any_brick) cand = vx_level27(vx_nb27(W.win, lx, ly, lz), bd2);
            const bool done = active && cand != 0u;
            if (done) {
                uint32_t bidx = kVxNone, brgb = 0u;
                int bb = __ffs((int)cand) - 1;
#ifndef PCCM_VX_SINGLE
#define PCCM_VX_SINGLE 0          // (measured: a separate one-candidate path costs 5 % -- two code paths diverge more than one loop)
#endif
                if (PCCM_VX_SINGLE && !(cand & (cand - 1u))) {
                    if (need_idx) {                          // one voxel at the minimum: its record, no tie to break
                        const uint2 k_a = __ldg(S.vkey + vx_cand_rank27(S, W.sslot, W.win, W.rb, lx, ly, lz, bb));
                        bidx = k_a.y; brgb = k_a.x;
                    }
                } else if (need_idx || need_e) {
                    // the voxels that tie at the minimum, two look-ups in flight per trip
                    uint32_t c2 = cand;
                    do {
                        const int b_a = __ffs((int)c2) - 1;
                        c2 &= c2 - 1u;
                        const int b_b = c2 ? __ffs((int)c2) - 1 : b_a;
                        c2 &= c2 - 1u;
                        const uint32_t r_a = vx_cand_rank27(S, W.sslot, W.win, W.rb, lx, ly, lz, b_a);
                        const uint32_t r_b = vx_cand_rank27(S, W.sslot, W.win, W.rb, lx, ly, lz, b_b);
                        VX_CHECK(r_a < S.n_total && r_b < S.n_total);
                        const uint2 k_a = __ldg(S.vkey + r_a), k_b = __ldg(S.vkey + r_b);
                        if (k_a.y < bidx) { bidx = k_a.y; brgb = k_a.x; bb = b_a; }
                        if (k_b.y < bidx) { bidx = k_b.y; brgb = k_b.x; bb = b_b; }
                    } while (c2);
                }
                P.vres[t] = make_uint4(bd2, vx_pack_e27(bb), bidx == kVxNone ? 0u : bidx, brgb);   // (no look-up: the index is not used)
            }
            const unsigned und = __ballot_sync(full, active && !done);
            if (und) {                                   // empty 27-neighbourhood (1-2 % of the voxels): collected per brick
                if (active && !done) W.pend[npend + __popc(und & ((1u << lane) - 1u))] = (uint16_t)(t - b0);
                npend += __popc(und);
                __syncwarp();
                if (npend >= 32) {
                    vx_finish_pending(P, d, Q, S, W, b0, 32, any_brick, nocc);
                    __syncwarp();
                    if (lane < npend - 32) W.pend[lane] = W.pend[32 + lane];
                    npend -= 32;
                    __syncwarp();
                }
            }
        }
        if (npend) vx_finish_pending(P, d, Q, S, W, b0, npend, any_brick, nocc);
    }
}

// EPILOGUE.  Threads walk the query POINTS in the ORIGINAL order of the input (coalesced): own colour, the other
// cloud's normal at the query index (quirk Q1) and the per-point outputs stream; only the voxel's 16-byte answer is a
// gather (through prank).  Points that share a voxel share its answer.  The grid is ONE resident wave of at most
// kVxEpiGrid blocks (a constant, not the SM count: the float sums must not depend on the GPU); the tiles of
// kVxEpiTile points of a direction are dealt round-robin to the direction's blocks, a thread keeps its sums in
// registers across its tiles and the block writes ONE reduction record at the end -> the warp / block folds are paid
// once per ~9 points of a thread instead of once per 2, and the sums depend on (n_a, n_b) only, not on scheduling.
// Multi-GPU slices are cut by voxel: a rank skips the points of voxels it did not search.
constexpr int kVxEpiThreads = 256;
#ifndef PCCM_VX_EPIPER
#define PCCM_VX_EPIPER 2
#endif
constexpr int kVxEpiPer = PCCM_VX_EPIPER;
constexpr int kVxEpiTile = kVxEpiThreads * kVxEpiPer;
#ifndef PCCM_EPI_MINBLOCKS
#define PCCM_EPI_MINBLOCKS 5         // (51 registers: the sums of a thread stay in registers; 6 blocks = 42 registers spill them)
#endif
#ifndef PCCM_VX_EPIGRID
#define PCCM_VX_EPIGRID 740            // 148 SMs x 5 resident blocks on a B200 (measured: 592 / 740 / 888 / 1480 / 1776 blocks within 1 us)
#endif
constexpr uint32_t kVxEpiGrid = PCCM_VX_EPIGRID;

// one query point and the 16-byte answer of its voxel (nothing to do when the voxel has none, or belongs to the other pass)
__device__ __forceinline__ void vx_epilogue_answer(const VxParams& P, const VxDir& D, const VoxPlan* __restrict__ plan,
                                                   const CloudView& qa, const CloudView& sa, uint32_t i, uint32_t rk, const uint4& v,
                                                   uint32_t qrgb, const double* staged_normal, VxAcc& acc) {
    if (v.x == kVxNone) return;
    const bool far = (v.z & kVxFarBit) != 0u;
    if (far != (P.pass == 1)) return;
    int ex, ey, ez;
    uint32_t nrgb = v.w;
    if (!far) {
        ex = (int)(v.y & 0xffu) - 128; ey = (int)((v.y >> 8) & 0xffu) - 128; ez = (int)((v.y >> 16) & 0xffu) - 128;
    } else {                                  // pencil-round answer: any distance, coordinates from the records
        const uint2 qv = __ldg(plan->view[D.qc].vxyz + rk);
        const uint4 nr = __ldg(D.srecs + v.y);               // {xy, z, idx, rgb}
        ex = (int)(qv.x & 0xffffu) - (int)(nr.x & 0xffffu); ey = (int)(qv.x >> 16) - (int)(nr.x >> 16); ez = (int)qv.y - (int)nr.y;
        nrgb = nr.w;
    }
    vx_epilogue(P, D, qa, sa, i, qrgb, v.x, ex, ey, ez, v.z & ~kVxFarBit, nrgb, acc, staged_normal);
}

// Bulk-copy staging (STAGED = true, the default; PCCM_EPI_TMA=0 selects the plain loads): the three arrays a tile reads
// in the input's own order -- prank, the other cloud's normals at the query index, the packed own colours -- are
// contiguous per tile, so ONE elected thread moves them into shared memory with cp.async.bulk (the 1-D form of the TMA
// engine: no tensor map), one tile ahead, completion counted in bytes on an mbarrier.  The threads then wait for one
// thing only, the gather of the voxel answers; the stream loads no longer queue behind it (they sat after the "does
// this point belong to my slice" branch) and cost no registers while in flight.  Sources are taken from the 16-byte
// boundary below a tile's first element to the one above its last (a 16-byte line never straddles a page).
struct VxEpiStage {
    alignas(16) unsigned char nrm[kVxEpiTile * 24 + 16];
    alignas(16) unsigned char rnk[kVxEpiTile * 4 + 16];
    alignas(16) unsigned char rgb[kVxEpiTile * 4 + 16];
};
struct VxEpiSrc {            // one staged array of a direction
    const unsigned char* base;   // null: not staged (absent, or not read in the input's order)
    uint32_t lead;               // bytes between the 16-byte boundary below the array and its first element
    uint32_t elem;               // bytes per point
    __device__ __forceinline__ void set(const void* p, uint32_t e) {
        base = static_cast<const unsigned char*>(p);
        lead = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u);
        elem = e;
    }
    __device__ __forceinline__ uint32_t issue(void* dst, uint32_t tile, uint32_t nq, unsigned long long* bar) const {
        if (!base) return 0u;
        const uint32_t first = tile * (uint32_t)kVxEpiTile, cnt = min((uint32_t)kVxEpiTile, nq - first);
        const uint32_t bytes = (lead + cnt * elem + 15u) & ~15u;
        bulk_g2s(dst, base + (size_t)first * elem - lead, bytes, bar);
        return bytes;
    }
};

// COMPACT (split pairs and voxel slices: this rank evaluates a fraction of the points): a tile only contributes the
// points whose voxel is indexed here -- their {index, rank} go to a queue in shared memory, in the tile's order (prefix
// sums, no atomics: the float sums stay reproducible), and the block works the queue off a full block of entries at a
// time.  The cost of a tile this rank owns nothing of is one coalesced read of its ranks and a block scan, not a trip
// through the epilogue with most lanes idle (measured on the 10 M pair split 8 ways: 204 -> 137 us per rank).
constexpr uint32_t kVxEpiQueue = 2 * kVxEpiTile;   // left-over entries (fewer than a block) + one tile; a power of two (ring)
static_assert((kVxEpiQueue & (kVxEpiQueue - 1u)) == 0u && kVxEpiQueue >= (uint32_t)kVxEpiThreads + (uint32_t)kVxEpiTile,
              "the epilogue's queue is a ring addressed by a mask");
template <bool STAGED, bool COMPACT>
__global__ void __launch_bounds__(kVxEpiThreads, PCCM_EPI_MINBLOCKS)
vx_epilogue_kernel(const __grid_constant__ VxParams P) {
    __shared__ double s_lut[256];
    __shared__ VxEpiStage s_stage[STAGED ? 2 : 1];
    __shared__ uint2 s_queue[COMPACT ? kVxEpiQueue : 1];
    __shared__ uint32_t s_wcnt[kVxEpiThreads / 32];
    __shared__ unsigned long long s_bar[2];
    pdl_launch();
    const int d = (P.ndirs > 1 && blockIdx.x >= P.dir[0].ntiles) ? 1 : 0;
    const VxDir& D = P.dir[d];
    const uint32_t blk = blockIdx.x - (d ? P.dir[0].ntiles : 0u);
    s_lut[threadIdx.x] = D.qa.lut255[threadIdx.x];       // k / 255.0 table: shared memory instead of six global loads per point
                                                         // (written when the context was created: read ahead of the wait)
    VxEpiSrc src_rnk{nullptr, 0u, 0u}, src_nrm{nullptr, 0u, 0u}, src_rgb{nullptr, 0u, 0u};
    if (STAGED) {
        src_rnk.set(D.qprank, 4u);
        if ((D.flags & PCCM_EVAL_D2) && P.normals_mode != PCCM_NORMALS_BY_NEIGHBOUR && D.sa.normals) src_nrm.set(D.sa.normals, 24u);
        if ((D.flags & PCCM_EVAL_COLOR) && D.qa.rgb_mode == 2) src_rgb.set(D.qa.rgb_u8, 4u);
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1u);
            mbar_init(&s_bar[1], 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    pdl_wait();
    const VoxPlan* __restrict__ plan = P.plan;
    if (plan->status) return;
    auto stage_tile = [&](uint32_t tile, int s) {          // (thread 0 only)
        uint32_t bytes = 0;
        // the arrive comes first in program order; the byte count of the three copies is known up front
        const uint32_t first = tile * (uint32_t)kVxEpiTile, cnt = min((uint32_t)kVxEpiTile, D.nq - first);
        if (src_rnk.base) bytes += (src_rnk.lead + cnt * 4u + 15u) & ~15u;
        if (src_nrm.base) bytes += (src_nrm.lead + cnt * 24u + 15u) & ~15u;
        if (src_rgb.base) bytes += (src_rgb.lead + cnt * 4u + 15u) & ~15u;
        mbar_expect_tx(&s_bar[s], bytes);
        src_rnk.issue(s_stage[s].rnk, tile, D.nq, &s_bar[s]);
        src_nrm.issue(s_stage[s].nrm, tile, D.nq, &s_bar[s]);
        src_rgb.issue(s_stage[s].rgb, tile, D.nq, &s_bar[s]);
    };
    __syncthreads();                                       // (table and barriers)
    if (STAGED && threadIdx.x == 0 && blk < D.npt_tiles) stage_tile(blk, 0);
    CloudView qa = D.qa, sa = D.sa;
    qa.lut255 = s_lut; sa.lut255 = s_lut;
    if (STAGED && src_rgb.base) qa.rgb_mode = 1;           // own colour: the packed word of the staged tile
    uint32_t t_lo = 0u, t_hi = kVxNone;                   // (kVxNone itself marks "no voxel")
    if (P.world > 1) vx_slice(P, plan->view[D.qc], t_lo, t_hi);
    VxAcc acc;
    acc.init();
    if (COMPACT) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t head = 0, tail = 0;                       // (the same in every thread)
        uint32_t rk[kVxEpiPer], nx[kVxEpiPer];
        auto load_ranks = [&](uint32_t tile, uint32_t* out) {
#pragma unroll
            for (int j = 0; j < kVxEpiPer; ++j) {
                const uint32_t i = tile * kVxEpiTile + threadIdx.x * kVxEpiPer + j;      // a thread's points are neighbours: queue order = input order
                out[j] = (tile < D.npt_tiles && i < D.nq) ? __ldg(D.qprank + i) : kVxNone;
            }
        };
        auto work_off = [&](uint32_t count) {              // entries head .. head + count (count <= block size)
            if (threadIdx.x < count) {                     // (two entries per thread, gathered together: 149 vs 138 us -- spills)
                const uint2 e = s_queue[(head + threadIdx.x) & (kVxEpiQueue - 1u)];
                const uint4 v = __ldg(P.vres + e.y);
                vx_epilogue_answer(P, D, plan, qa, sa, e.x, e.y, v, 0u, nullptr, acc);
            }
            head += count;
        };
        load_ranks(blk, nx);
        for (uint32_t tile = blk; tile < D.npt_tiles; tile += D.ntiles) {
#pragma unroll
            for (int j = 0; j < kVxEpiPer; ++j) rk[j] = nx[j];
            load_ranks(tile + D.ntiles, nx);               // (the next tile's ranks travel during this tile's scan)
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < kVxEpiPer; ++j) mine += (rk[j] >= t_lo && rk[j] < t_hi) ? 1u : 0u;
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) s_wcnt[warp] = incl;
            __syncthreads();
            uint32_t off = tail + incl - mine, total = 0;
#pragma unroll
            for (int w = 0; w < kVxEpiThreads / 32; ++w) { const uint32_t c = s_wcnt[w]; if (w < warp) off += c; total += c; }
#pragma unroll
            for (int j = 0; j < kVxEpiPer; ++j)
                if (rk[j] >= t_lo && rk[j] < t_hi)
                    s_queue[(off++) & (kVxEpiQueue - 1u)] = make_uint2(tile * kVxEpiTile + threadIdx.x * kVxEpiPer + j, rk[j]);
            tail += total;
            __syncthreads();
            while (tail - head >= (uint32_t)kVxEpiThreads) work_off(kVxEpiThreads);
        }
        work_off(tail - head);                               // (fewer than a block)
    } else {
    uint32_t it = 0;
    for (uint32_t tile = blk; tile < D.npt_tiles; tile += D.ntiles, ++it) {
    const int st = STAGED ? (int)(it & 1u) : 0;
    const uint32_t i0 = tile * kVxEpiTile + threadIdx.x;
    if (STAGED) {
        if (threadIdx.x == 0 && tile + D.ntiles < D.npt_tiles) stage_tile(tile + D.ntiles, st ^ 1);   // (its last readers passed the barrier below)
        mbar_wait(&s_bar[st], (it >> 1) & 1u);
    }
    uint32_t rk[kVxEpiPer];
    uint4 v[kVxEpiPer];
#pragma unroll
    for (int j = 0; j < kVxEpiPer; ++j) {
        const uint32_t i = i0 + j * kVxEpiThreads;
        if (STAGED) rk[j] = i < D.nq ? *reinterpret_cast<const uint32_t*>(s_stage[st].rnk + src_rnk.lead + (size_t)(threadIdx.x + j * kVxEpiThreads) * 4u) : kVxNone;
        else rk[j] = i < D.nq ? __ldg(D.qprank + i) : kVxNone;
    }
#pragma unroll
    for (int j = 0; j < kVxEpiPer; ++j)
        v[j] = (rk[j] >= t_lo && rk[j] < t_hi) ? __ldg(P.vres + rk[j]) : make_uint4(kVxNone, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < kVxEpiPer; ++j) {
        const uint32_t lp = threadIdx.x + j * kVxEpiThreads;         // position in the tile
        const double* sn = (STAGED && src_nrm.base) ? reinterpret_cast<const double*>(s_stage[st].nrm + src_nrm.lead + (size_t)lp * 24u) : nullptr;
        const uint32_t qrgb = (STAGED && src_rgb.base) ? *reinterpret_cast<const uint32_t*>(s_stage[st].rgb + src_rgb.lead + (size_t)lp * 4u) : 0u;
        vx_epilogue_answer(P, D, plan, qa, sa, i0 + j * kVxEpiThreads, rk[j], v[j], qrgb, sn, acc);
    }
    if (STAGED) __syncthreads();                  // everybody is done with this stage before it is refilled (two tiles from now)
    }
    }
    BlockPartial r;
    vx_warp_record(acc, D.flags, r);
    __shared__ BlockPartial sm[kVxEpiThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockPartial o = sm[0];
        for (int w = 1; w < kVxEpiThreads / 32; ++w) partial_merge(o, sm[w]);
        P.partials[D.rec_off + (uint32_t)P.pass * D.ntiles + blk] = o;
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
    pdl_enter();
    __shared__ int s_slot[4][128];
    __shared__ int s_occ[4][128];
    const unsigned full = 0xffffffffu;
    const VoxPlan* __restrict__ plan = P.plan;
    if (plan->status) return;
    const int lane = threadIdx.x & 31;
    int* bslot = s_slot[threadIdx.x >> 5];
    int* bocc = s_occ[threadIdx.x >> 5];
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int d = 0; d < P.ndirs; ++d) {
        const VxDir& D = P.dir[d];
        const VoxView& Q = plan->view[D.qc];
        const VoxView& S = plan->view[D.sc];
        const uint32_t ntodo = P.counters[d];
        for (uint32_t w = gwarp; w < ntodo; w += nwarps) {
            const uint32_t t = D.todo[w];
            const uint2 qr = __ldg(Q.vxyz + t);
            const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
            const int qbx = qx >> 5, qby = qy >> 3, qbz = qz >> 3;
            __syncwarp();
            // occupied bricks of the 5 x 5 x 5 neighbourhood, compacted: bocc[k] = brick number, bslot[k] = slot
            int nocc = 0;
            for (int b0 = 0; b0 < 125; b0 += 32) {
                const int b = b0 + lane;
                const int slot = b < 125 ? vx_slot(S, qbx + b % 5 - 2, qby + (b / 5) % 5 - 2, qbz + b / 25 - 2) : -1;
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
            const uint32_t best = vx_warp_bricks<5>(S, bslot, bocc, nocc, qx, qy, qz, 289u, brank);
            if (best >= 289u) {               // nothing certified within two brick rings
                if (lane == 0) D.far[atomicAdd(P.counters + 2 + d, 1u)] = t;
                continue;
            }
            if (lane == 0) {
                const uint2 nc = __ldg(S.vxyz + brank), nk = __ldg(S.vkey + brank);
                P.vres[t] = make_uint4(best, vx_pack_e(qx - (int)(nc.x & 0xffffu), qy - (int)(nc.x >> 16), qz - (int)nc.y), nk.y, nk.x);
            }
        }
    }
}

// What the brick rings could not certify (the nearest point is tens of voxels away: disjoint
// clouds, isolated outliers): exact pencil search, which skips empty space by its row table.
__global__ void __launch_bounds__(128)
vx_far_kernel(const __grid_constant__ VxParams P) {
    pdl_enter();
    const VoxPlan* __restrict__ plan = P.plan;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int d = 0; d < P.ndirs; ++d) {
        const VxDir& D = P.dir[d];
        const uint2* __restrict__ qxyz = plan->view[D.qc].vxyz;
        const uint32_t nfar = P.counters[2 + d];
        for (uint32_t w = gtid; w < nfar; w += stride) {
            const uint32_t t = D.far[w];
            const uint2 qr = __ldg(qxyz + t);
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
// cloud_pair.py:108-109) on the brick index: the 26 / 124 voxels around the voxel on the staged rows;
// a voxel that holds more than one point answers 0
// ------------------------------------------------------------------------------------
struct VxSelfParams {
    VoxView c;
    uint32_t n;                // points of the cloud
    uint32_t begin, end;       // slice of the cloud's points [begin, end) -> the same share of its voxels
    int32_t own_zlo, own_zhi;  // split pairs: only the voxels of these brick layers (the slice is then everything)
    uint32_t* dupbits;         // [n_total / 32 + 1] by rank: voxel holds more than one point
    uint32_t* vself;           // [n_total] by rank: squared distance to the nearest other point (kVxNone: not decided here)
    double* minmax;            // per brick {min, max} of the distances (sqrt)
    uint32_t* undecided;       // [0] count, [1..] ranks of the voxels farther than 8 voxels from every other one
    double* per_point;         // optional, original order
};

__global__ void vx_dupflag_kernel(const __grid_constant__ VxSelfParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t rank = __ldg(P.c.prank + i);
    if (rank == kVxNone) return;                       // (split pair: the point is outside this rank's slab)
    if (__ldg(&P.c.vkey[rank].y) != i) atomicOr(P.dupbits + (rank >> 5), 1u << (rank & 31u));
}

struct VxSelfSmem {
    uint2 win[kVxRegRows];
    int sslot[28];
    int occ_slot[28];
    int occ_id[28];
};
__global__ void __launch_bounds__(kVxThreads)
vx_selfnn_kernel(const __grid_constant__ VxSelfParams P) {
    __shared__ VxSelfSmem s_w[kVxWarps];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    VxSelfSmem& W = s_w[warp];
    const uint32_t lb = blockIdx.x * kVxWarps + warp;
    if (lb >= P.c.nblk) return;
    const uint32_t slot = P.c.slot0 + lb;
    const uint32_t b0 = vx_brick_begin(P.c, slot), b1 = vx_brick_begin(P.c, slot + 1);
    const uint32_t r0 = vx_ranked_begin(P.c), nd = vx_ndistinct(P.c);
    const uint32_t t_lo = r0 + (uint32_t)((unsigned long long)nd * P.begin / (P.n ? P.n : 1u));
    const uint32_t t_hi = r0 + (uint32_t)((unsigned long long)nd * P.end / (P.n ? P.n : 1u));
    const uint32_t t0 = max(b0, t_lo), t1 = min(b1, t_hi);
    uint32_t mn = kVxNone, mx = 0u;
    bool any = false;
    bool mine = t0 < t1;
    if (mine) {
        const int bz0 = (int)__ldg(&P.c.vxyz[b0].y) >> 3;
        mine = bz0 >= P.own_zlo && bz0 < P.own_zhi;
    }
    if (mine) {
        const uint2 first = __ldg(P.c.vxyz + b0);
        const int bx = (int)(first.x & 0xffffu) >> 5, by = (int)(first.x >> 16) >> 3, bz = (int)first.y >> 3;
        int myslot = -1;
        if (lane < 27) {
            myslot = vx_slot(P.c, bx + lane % 3 - 1, by + (lane / 3) % 3 - 1, bz + lane / 9 - 1);
            W.sslot[lane] = myslot;
        }
        const unsigned occ = __ballot_sync(full, myslot >= 0);
        const int nocc = __popc(occ);
        if (myslot >= 0) {
            const int k = __popc(occ & ((1u << lane) - 1u));
            W.occ_slot[k] = myslot;
            W.occ_id[k] = lane;
        }
        __syncwarp();
        for (int i = lane; i < kVxRegRows; i += 32) {
            uint32_t rb;
            W.win[i] = vx_stage_row(P.c, W.sslot, i, rb);
        }
        __syncwarp();
        for (uint32_t tb = t0; tb < t1; tb += 32) {
            const uint32_t t = tb + lane;
            const bool active = t < t1;
            const uint2 qr = __ldg(P.c.vxyz + (active ? t : t0));
            const int qx = (int)(qr.x & 0xffffu), qy = (int)(qr.x >> 16), qz = (int)qr.y;
            const int lx = qx & 31, ly = (qy & 7) + 2, lz = (qz & 7) + 2;
            const bool dup = active && ((__ldg(P.dupbits + (t >> 5)) >> (t & 31u)) & 1u);
            uint32_t bd2 = vx_self27(vx_nb27(W.win, lx, ly, lz));
            if (__any_sync(full, active && !dup && bd2 == kVxNone)) {
                if (bd2 == kVxNone) bd2 = vx_self125(W.win, lx, ly, lz);
            }
            bool done = bd2 < 9u;
            unsigned pend = __ballot_sync(full, active && !done && !dup);
            while (pend) {
                const int src = __ffs((int)pend) - 1;
                pend &= pend - 1u;
                const int sx = __shfl_sync(full, qx, src), sy = __shfl_sync(full, qy, src), sz = __shfl_sync(full, qz, src);
                uint32_t dummy = kVxNone;
                const uint32_t nd2 = vx_warp_bricks<3, true>(P.c, W.occ_slot, W.occ_id, nocc, sx, sy, sz, 81u, dummy);
                if (nd2 < 81u && lane == src) { bd2 = nd2; done = true; }
            }
            if (active) {
                const uint32_t v = dup ? 0u : (done ? bd2 : kVxNone);
                P.vself[t] = v;
                if (v != kVxNone) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; any = true; }
            }
            const unsigned und = __ballot_sync(full, active && !dup && !done);
            if (und) {
                uint32_t pos = 0;
                if (lane == 0) pos = atomicAdd(P.undecided, (uint32_t)__popc(und));
                pos = __shfl_sync(full, pos, 0);
                if (active && !dup && !done) P.undecided[1 + pos + __popc(und & ((1u << lane) - 1u))] = t;
            }
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

// The voxels vx_selfnn_kernel could not decide (no other voxel within 8): exact pencil search for the nearest
// OTHER point (2-NN: the voxel's own point is the first hit); their min / max go to extra slots of the min / max array.
struct VxSelfFarParams {
    VxSelfParams s;
    RowGrid grid;
    const uint4* recs;
    const uint32_t* row_start;
    double* minmax_extra;      // [2 * nblocks]
};
__global__ void __launch_bounds__(128)
vx_selffar_kernel(const __grid_constant__ VxSelfFarParams P) {
    const uint32_t n = P.s.undecided[0];
    double mn = INFINITY, mx = -INFINITY;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
        const uint32_t t = P.s.undecided[1 + w];
        const uint2 qr = __ldg(P.s.c.vxyz + t);
        KInt::Q q;
        q.x = (int)(qr.x & 0xffffu); q.y = (int)(qr.x >> 16); q.z = (int)qr.y;
        Best2Val<KInt> best;                 // the voxel holds exactly one point (itself, distance 0): the second one is the answer
        best.init();
        search<KInt>(P.grid, P.row_start, P.recs, q, best);
        P.s.vself[t] = best.m2;
        const double dd = sqrt((double)best.m2);
        mn = fmin(mn, dd); mx = fmax(mx, dd);
    }
    __shared__ double sm[4];
    const double bmn = block_min<128>(mn, sm);
    const double bmx = block_max<128>(mx, sm);
    if (threadIdx.x == 0) { P.minmax_extra[2 * blockIdx.x] = bmn; P.minmax_extra[2 * blockIdx.x + 1] = bmx; }
}

__global__ void vx_selfout_kernel(const __grid_constant__ VxSelfParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const uint32_t rank = __ldg(P.c.prank + i);
    P.per_point[i] = rank == kVxNone ? NAN : sqrt((double)P.vself[rank]);
}

}  // namespace pccm
