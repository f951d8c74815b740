// pccm_kernels.cuh -- sm_100a kernels of libpccm.so (index build, fused symmetric
// NN query with D1 / D2 / colour epilogues and block reductions, self k-NN with the
// boundary-distance and normal-estimation epilogues).  Host orchestration and the
// C ABI are in pccm_api.cu; the per-query search logic is in pccm_core.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pccm_core.cuh"
#include "../../include/pccm.h"

namespace pccm {

// Programmatic dependent launch (sm_90+ griddepcontrol): kernels of one evaluation are enqueued back to back on one
// stream; launched with the programmatic-serialisation attribute (launch_chain in pccm_api.cu) a kernel's blocks become
// resident while the previous kernel drains its last wave, and pdl_enter() -- the first statement of every such
// kernel -- holds them until that kernel has completed and its writes are visible.  Every chained kernel waits
// unconditionally, so completion stays transitive along the chain.  Both instructions are no-ops under an ordinary launch.
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_launch();
    pdl_wait();
}

// Bulk asynchronous copies (the 1-D form of the TMA engine: cp.async.bulk, no tensor map) with completion counted in
// bytes on a shared-memory mbarrier.  Used where a kernel streams contiguous arrays (stats_kernel: the float64 rows;
// vx_epilogue_kernel: ranks, normals, colours of a tile): one elected thread requests whole tiles, tiles ahead, and the
// loads cost the other threads neither registers nor issue slots while they travel.  Source, destination and size are
// multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

// Zeroing as a KERNEL of the chain instead of cudaMemsetAsync: a memset may be handed to a copy engine, where it queues
// behind the attribute uploads of the copy stream (measured end to end: the index build then waits ~2 ms for 96 MB
// of colours and normals it does not need), and a memset node also cuts the programmatic-launch chain.
__global__ void zero_words_kernel(uint32_t* __restrict__ p, size_t nwords) {
    pdl_enter();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}

// The read-back of an evaluation as ONE kernel of the chain: up to eight small device blocks (result records, counters,
// plan, statistics, colour flags) are copied word by word into the context's pinned host block (cudaMallocHost memory
// is addressable from the device) -- instead of as many cudaMemcpyAsync calls, each a stream operation of its own.
struct GatherArgs {
    const uint32_t* src[8];
    uint32_t* dst[8];
    uint32_t words[8];
    int n;
};
__global__ void __launch_bounds__(256) gather_to_host_kernel(const __grid_constant__ GatherArgs G) {
    pdl_enter();
    for (int k = 0; k < G.n; ++k)
        for (uint32_t i = threadIdx.x; i < G.words[k]; i += 256) G.dst[k][i] = G.src[k][i];
}

constexpr int kStatsThreads = 256;
constexpr int kQueryThreads = 128;
constexpr int kKnnThreads = 64;
#ifndef PCCM_STAGE
#define PCCM_STAGE 0   // 1 = stage the tile's pencil window in shared memory (exact; measured 15 % slower, profiles/README.md)
#endif


// ------------------------------------------------------------------------------------
// raw input access
// ------------------------------------------------------------------------------------
constexpr int kDtypeVRec = 100;   // internal source format: the brick index's voxel records, addressed through prank[] (pccm_vox.cuh)

__device__ __forceinline__ double load_coord(const void* base, int dtype, int64_t stride, int64_t i, int axis) {
    const char* p = static_cast<const char*>(base) + i * stride;
    switch (dtype) {
        case kDtypeVRec: {      // base = voxel records, i = the voxel's rank
            const uint32_t* u = reinterpret_cast<const uint32_t*>(p);
            return axis == 0 ? (double)(u[0] & 0xffffu) : (axis == 1 ? (double)(u[0] >> 16) : (double)u[1]);
        }
        case PCCM_F64: return reinterpret_cast<const double*>(p)[axis];
        case PCCM_F32: return (double)reinterpret_cast<const float*>(p)[axis];
        case PCCM_I32: return (double)reinterpret_cast<const int32_t*>(p)[axis];
        case PCCM_U16: return (double)reinterpret_cast<const uint16_t*>(p)[axis];
        default:       return (double)reinterpret_cast<const uint8_t*>(p)[axis];
    }
}

struct StatsPartial {
    double mn[3], mx[3];
    uint32_t not_int, not_f32, not_finite, rgb_not_u8;
};

// The same statistics as ONE record in device memory, for the kernels that plan the brick index without the
// host (pccm_vox_kernels.cuh): integer bounding box (the minimum as 0x7fffffff - min so that a zeroed record is the
// neutral element of atomicMax for every field) and the classification flags.
struct DevStats {
    uint32_t nmn[3], mx[3];
    uint32_t flags, pad;
};
constexpr uint32_t kDevNotInt = 1u, kDevNotF32 = 2u, kDevNonFinite = 4u;
constexpr int kZHistBins = 4096;          // 8-voxel layers of a 15-bit coordinate range

// K0: bounding box + classification of coordinates (and colours) in one pass; integer-valued coordinates are
// also written as 8-byte {x | y << 16, z} records -- the form every later pass of the brick index reads.
// F64ROWS: packed float64 rows (the form numpy / Open3D hand over): plain 8-byte loads instead of the per-element format
// switch.  Values that are integers in [0, 32767] -- every coordinate of voxelised content -- take the short way: one round
// trip through int, integer min / max; only other values pay for the float64 min / max and the float32 / finite checks.
// MODE 2 = F64ROWS with the rows staged by bulk copies: tiles of kStatsThreads rows (6 KB), kStatsStages - 1 tiles of a
// block in flight, no load instruction in the loop but the shared-memory reads (stride 24 bytes: conflict-free).
constexpr int kStatsStages = 4;
template <int MODE>
__global__ void __launch_bounds__(kStatsThreads, 4)
stats_kernel(const void* xyz, int dtype, int64_t stride, int64_t n,
             const void* rgb, int rgb_dtype, int64_t rgb_stride, StatsPartial* out, uint2* packed, DevStats* dev, uint32_t* zhist) {
    constexpr bool F64ROWS = MODE != 0;
    constexpr bool STAGED = MODE == 2;
    __shared__ __align__(16) unsigned char s_rows[STAGED ? kStatsStages : 1][STAGED ? kStatsThreads * 24 + 16 : 16];
    __shared__ unsigned long long s_bar[kStatsStages];
    if (STAGED && threadIdx.x == 0) {
        for (int k = 0; k < kStatsStages; ++k) mbar_init(&s_bar[k], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_enter();
    // zhist (sharded builds: one pair split over several GPUs by slabs of z): points per 8-voxel layer, kZHistBins
    // bins, accumulated per block in dynamic shared memory
    extern __shared__ uint32_t s_hist[];
    if (zhist) {
        for (int b = threadIdx.x; b < kZHistBins; b += kStatsThreads) s_hist[b] = 0;
        __syncthreads();
    }
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int imn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, imx[3] = {-1, -1, -1};        // over the values that are integers in [0, 32767]
    uint32_t not_int = 0, not_f32 = 0, not_fin = 0, rgb_bad = 0;
    auto point = [&](int64_t i, const double (&v3)[3]) {
        int i3[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = v3[a];
            const int iv = __double2int_rz(v);
            i3[a] = iv;
            if ((double)iv == v && (unsigned)iv <= 32767u) {         // the common case (voxelised content): one round trip through int
                imn[a] = min(imn[a], iv);
                imx[a] = max(imx[a], iv);
            } else {
                not_int = 1;
                mn[a] = fmin(mn[a], v);
                mx[a] = fmax(mx[a], v);
                if (!isfinite(v)) not_fin = 1;
                if ((double)(float)v != v) not_f32 = 1;
            }
        }
        if (packed) packed[i] = make_uint2(((uint32_t)i3[0] & 0xffffu) | ((uint32_t)i3[1] << 16), (uint32_t)i3[2]);
        if (zhist) atomicAdd(&s_hist[min(max(i3[2], 0) >> 3, kZHistBins - 1)], 1u);
        if (rgb != nullptr && rgb_dtype == PCCM_F64) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double c = load_coord(rgb, PCCM_F64, rgb_stride, i, a);
                double k = rint(c * 255.0);
                if (!(k >= 0.0 && k <= 255.0 && k / 255.0 == c)) rgb_bad = 1;
            }
        }
    };
    if (STAGED) {
        const unsigned char* base = static_cast<const unsigned char*>(xyz);
        const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(xyz) & 15u);        // (0 or 8: rows of float64)
        const int64_t ntiles = (n + kStatsThreads - 1) / kStatsThreads;
        auto request = [&](int64_t tile, int st) {           // (thread 0) rows of the tile, from the 16-byte line below the first to the one above the last
            const int64_t first = tile * kStatsThreads;
            const uint32_t cnt = (uint32_t)min((int64_t)kStatsThreads, n - first);
            const uint32_t bytes = (lead + cnt * 24u + 15u) & ~15u;
            mbar_expect_tx(&s_bar[st], bytes);
            bulk_g2s(s_rows[st], base + first * 24 - lead, bytes, &s_bar[st]);
        };
        __syncthreads();                                      // (barriers initialised)
        if (threadIdx.x == 0)
            for (int k = 0; k < kStatsStages - 1; ++k) {
                const int64_t tile = blockIdx.x + (int64_t)k * gridDim.x;
                if (tile < ntiles) request(tile, k);
            }
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = (int)(it % kStatsStages);
            if (threadIdx.x == 0) {                           // the stage consumed in the previous trip takes the tile kStatsStages - 1 ahead
                const int64_t ahead = tile + (int64_t)(kStatsStages - 1) * gridDim.x;
                if (ahead < ntiles) request(ahead, (int)((it + kStatsStages - 1) % kStatsStages));
            }
            mbar_wait(&s_bar[st], (it / kStatsStages) & 1u);
            const int64_t i = tile * kStatsThreads + threadIdx.x;
            if (i < n) {
                const double* row = reinterpret_cast<const double*>(s_rows[st] + lead + threadIdx.x * 24);
                const double v3[3] = {row[0], row[1], row[2]};
                point(i, v3);
            }
            __syncthreads();                                  // everybody has read the stage before it is requested again
        }
    } else {
    // two points of a thread are in flight (their loads are requested together): the pass is bound by memory latency
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t ib = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ib < n; ib += 2 * gstride) {
      double w3[2][3];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = ib + u * gstride;
        if (i >= n) continue;
        if (F64ROWS) {
            const double* row = static_cast<const double*>(xyz) + 3 * i;
            w3[u][0] = __ldg(row); w3[u][1] = __ldg(row + 1); w3[u][2] = __ldg(row + 2);
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a) w3[u][a] = load_coord(xyz, dtype, stride, i, a);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = ib + u * gstride;
        if (i >= n) continue;
        point(i, w3[u]);
      }
    }
    }
    __shared__ double s_mn[3][kStatsThreads / 32], s_mx[3][kStatsThreads / 32];
    __shared__ uint32_t s_flags[4];
    if (threadIdx.x < 4) s_flags[threadIdx.x] = 0;
    __syncthreads();
    if (zhist)
        for (int b = threadIdx.x; b < kZHistBins; b += kStatsThreads)
            if (s_hist[b]) atomicAdd(zhist + b, s_hist[b]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (imx[a] >= 0) { mn[a] = fmin(mn[a], (double)imn[a]); mx[a] = fmax(mx[a], (double)imx[a]); }
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], __shfl_down_sync(0xffffffffu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_down_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
    }
    if (not_int) atomicOr(&s_flags[0], 1u);
    if (not_f32) atomicOr(&s_flags[1], 1u);
    if (not_fin) atomicOr(&s_flags[2], 1u);
    if (rgb_bad) atomicOr(&s_flags[3], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        StatsPartial p;
        for (int a = 0; a < 3; ++a) {
            double lo = s_mn[a][0], hi = s_mx[a][0];
            for (int w = 1; w < kStatsThreads / 32; ++w) { lo = fmin(lo, s_mn[a][w]); hi = fmax(hi, s_mx[a][w]); }
            p.mn[a] = lo; p.mx[a] = hi;
        }
        p.not_int = s_flags[0]; p.not_f32 = s_flags[1]; p.not_finite = s_flags[2]; p.rgb_not_u8 = s_flags[3];
        out[blockIdx.x] = p;
        if (!p.not_int) {
            for (int a = 0; a < 3; ++a) {
                atomicMax(&dev->nmn[a], 0x7fffffffu - (uint32_t)(int)p.mn[a]);
                atomicMax(&dev->mx[a], (uint32_t)(int)p.mx[a]);
            }
        }
        const uint32_t f = (p.not_int ? kDevNotInt : 0u) | (p.not_f32 ? kDevNotF32 : 0u) | (p.not_finite ? kDevNonFinite : 0u);
        if (f) atomicOr(&dev->flags, f);
    }
}

// are all colour channels k / 255 for an integer k in [0, 255]?  (flag = 1 when not)
__global__ void rgb_classify_kernel(const void* rgb, int64_t stride, int64_t n, uint32_t* flag) {
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double c = load_coord(rgb, PCCM_F64, stride, i, a);
            const double k = rint(c * 255.0);
            if (!(k >= 0.0 && k <= 255.0 && k / 255.0 == c)) bad = true;
        }
    }
    if (bad) *flag = 1u;
}

// colours -> uchar4 (original order).  F64 input must have passed the k/255 test.
__global__ void pack_rgb_u8_kernel(const void* rgb, int rgb_dtype, int64_t stride, int64_t n, uchar4* out) {
    pdl_enter();
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uchar4 c;
    if (rgb_dtype == PCCM_U8) {
        const uint8_t* p = static_cast<const uint8_t*>(rgb) + i * stride;
        c = make_uchar4(p[0], p[1], p[2], 0);
    } else {
        c.x = (unsigned char)rint(load_coord(rgb, PCCM_F64, stride, i, 0) * 255.0);
        c.y = (unsigned char)rint(load_coord(rgb, PCCM_F64, stride, i, 1) * 255.0);
        c.z = (unsigned char)rint(load_coord(rgb, PCCM_F64, stride, i, 2) * 255.0);
        c.w = 0;
    }
    out[i] = c;
}

// packed 3-byte colours -> uchar4, four points per thread (three aligned 32-bit loads, one 16-byte store)
__global__ void pack_rgb_u8x4_kernel(const uint32_t* __restrict__ rgb, int64_t n, uchar4* __restrict__ out) {
    pdl_enter();
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;     // group of four points
    const int64_t i = 4 * q;
    if (i >= n) return;
    if (i + 4 <= n) {
        const uint32_t a = __ldg(rgb + 3 * q), b = __ldg(rgb + 3 * q + 1), c = __ldg(rgb + 3 * q + 2);
        uint4 o;
        o.x = a & 0xffffffu;
        o.y = (a >> 24) | ((b & 0xffffu) << 8);
        o.z = (b >> 16) | ((c & 0xffu) << 16);
        o.w = c >> 8;
        reinterpret_cast<uint4*>(out)[q] = o;
    } else {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(rgb);
        for (int64_t k = i; k < n; ++k) out[k] = make_uchar4(p[3 * k], p[3 * k + 1], p[3 * k + 2], 0);
    }
}

// strided rows of 3 (F64 / F32) -> packed double[n][3]
__global__ void pack_f64x3_kernel(const void* src, int dtype, int64_t stride, int64_t n, double* out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[3 * i + 0] = load_coord(src, dtype, stride, i, 0);
    out[3 * i + 1] = load_coord(src, dtype, stride, i, 1);
    out[3 * i + 2] = load_coord(src, dtype, stride, i, 2);
}

// ------------------------------------------------------------------------------------
// K1: sort keys + row histogram
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t flip_f32(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long flip_f64(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

template <class K> struct RecPack;
template <> struct RecPack<KInt> {
    static __device__ __forceinline__ uint4 make(double x, double y, double z, uint32_t idx, uint32_t rgba) {
        return make_uint4((uint32_t)(int)x | ((uint32_t)(int)y << 16), (uint32_t)(int)z, idx, rgba);
    }
};
template <> struct RecPack<KF32> {
    static __device__ __forceinline__ float4 make(double x, double y, double z, uint32_t idx, uint32_t) {
        return make_float4((float)x, (float)y, (float)z, __uint_as_float(idx));
    }
};
template <> struct RecPack<KF64> {
    static __device__ __forceinline__ RecF64 make(double x, double y, double z, uint32_t idx, uint32_t) {
        RecF64 r; r.x = x; r.y = y; r.z = z; r.idx = idx; return r;
    }
};

// ------------------------------------------------------------------------------------
// exclusive prefix sum of the row histogram -> pencil table (hand-written, three launches:
// per-tile sums, scan of the tile sums by one block, tile-local scan + offset).  In place.
// ------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                         // per thread
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 entries per block

__device__ __forceinline__ uint32_t block_exclusive_scan_u32(uint32_t v, uint32_t* total) {
    __shared__ uint32_t ws[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) { if (w < warp) base += ws[w]; tot += ws[w]; }
    if (total) *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ data, size_t n, uint32_t* __restrict__ sums) {
    const size_t base = (size_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const size_t i = base + (size_t)j * kScanThreads + threadIdx.x;
        if (i < n) s += data[i];
    }
    uint32_t tot;
    block_exclusive_scan_u32(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(uint32_t* sums, uint32_t nblocks) {
    uint32_t carry = 0;                                // one block walks the tile sums (<= 32k of them)
    for (uint32_t b0 = 0; b0 < nblocks; b0 += kScanThreads) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < nblocks ? sums[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan_u32(v, &tot);
        if (i < nblocks) sums[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ sums) {
    // thread t owns kScanItems CONSECUTIVE entries of the tile: local serial scan + block scan of the per-thread sums
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) { v[j] = base + j < n ? data[base + j] : 0u; s += v[j]; }
    uint32_t ex = block_exclusive_scan_u32(s, nullptr) + sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) { if (base + j < n) data[base + j] = ex; ex += v[j]; }
}

// short tables (brick directory, brick totals: <= kScanSmallMax entries): one block, one launch --
// thread t sums its contiguous segment, the block scans the 1024 sums, thread t writes its prefixes
constexpr int kScanSmallThreads = 1024;
constexpr size_t kScanSmallMax = 64 * 1024;
// POPC: the input is a bitmap (`bits`, n - 1 words) and entry i of the output is the number of set bits before
// word i (the brick directory: occupancy bits -> slot prefix, entry n - 1 = number of bricks); else in place.
template <bool POPC>
__global__ void __launch_bounds__(kScanSmallThreads) scan_small_kernel(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ bits) {
    __shared__ uint32_t ws[kScanSmallThreads / 32];
    const uint32_t per = (n + kScanSmallThreads - 1) / kScanSmallThreads;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    // eight predicated loads per trip: independent, all in flight together (a plain loop leaves a sequential
    // remainder -- 21 entries per thread took 18 us, one dependent miss after the other)
    uint32_t s = 0;
    for (uint32_t c = lo; c < hi; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = c + j < hi ? (POPC ? (c + j < n - 1 ? (uint32_t)__popc(bits[c + j]) : 0u) : data[c + j]) : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = ws[lane], in = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, in, o);
            if (lane >= o) in += u;
        }
        ws[lane] = in - v;
    }
    __syncthreads();
    uint32_t ex = ws[warp] + incl - s;
    for (uint32_t c = lo; c < hi; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = c + j < hi ? (POPC ? (c + j < n - 1 ? (uint32_t)__popc(bits[c + j]) : 0u) : data[c + j]) : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (c + j < hi) data[c + j] = ex;
            ex += v[j];
        }
    }
}

// ------------------------------------------------------------------------------------
// joint index build of the two clouds of a pair: one key pass, ONE sort (cloud id in the top
// key bit), one scan, one reorder -- half the launches of two separate builds
// ------------------------------------------------------------------------------------
struct PairRaw {
    const void* xyz[2];
    const void* rgb[2];
    int64_t stride[2], rgb_stride[2];
    int32_t dtype[2], rgb_dtype[2];
    int32_t rgb_in_rec[2];     // pack the colour into the record (KInt + 8-bit colours)
    uint32_t n[2];
    uint32_t table_off[2];     // offset of each cloud's row table inside the joint table
    RowGrid g[2];
    const uint32_t* vprank;    // kDtypeVRec sources: rank of the voxel of every point (cloud 1's points follow cloud 0's)
    uint32_t vprank_off[2];    // first entry of each cloud in vprank
};

// Row of the coordinate source that holds point li of cloud c (brick-index sources: the record of
// the point's voxel).
__device__ __forceinline__ uint32_t pair_src(const PairRaw& R, int c, uint32_t li) {
    if (R.dtype[c] != kDtypeVRec) return li;
    return __ldg(R.vprank + R.vprank_off[c] + li);
}

__device__ __forceinline__ uint32_t float_row(const RowGrid& g, double y, double z) {
    int cy = (int)floor((y - g.y0) * g.inv_h), cz = (int)floor((z - g.z0) * g.inv_h);
    cy = cy < 0 ? 0 : (cy >= g.ny ? g.ny - 1 : cy);
    cz = cz < 0 ? 0 : (cz >= g.nz ? g.nz - 1 : cz);
    return (uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy;
}
__device__ __forceinline__ uint32_t int_row(const RowGrid& g, int y, int z) {
    uint32_t cy = (uint32_t)((y - g.iy0) >> g.shift), cz = (uint32_t)((z - g.iz0) >> g.shift);
    return cz * (uint32_t)g.ny + cy;
}

// ------------------------------------------------------------------------------------
// hand-written replacement of the radix sort for the pair build (KInt): counting sort by row
// (the row histogram and its scan are needed for the pencil table anyway) + per-row sort by
// (x, original index).  Rows are short on voxelised clouds (a 2x2 tube crosses the surface a
// few times), so most of them fit one warp: bitonic network over lanes with shuffles.
// ------------------------------------------------------------------------------------
// Sort items: (x key, point index), compared as one number.  Integer and float32 coordinates fit a 64-bit word
// (32-bit order-preserving key | index); float64 x needs its 64 bits, so its items are a pair of words.
struct Item128 {
    unsigned long long k, i;
};
template <int KIND> struct RowItem {
    typedef unsigned long long T;
    static __device__ __forceinline__ T make(double x, uint32_t li) {
        const uint32_t key = KIND == KIND_INT ? (uint32_t)(int)x : flip_f32((float)x);
        return ((unsigned long long)key << 32) | li;
    }
    static __device__ __forceinline__ T inf() { return ~0ull; }
    static __device__ __forceinline__ bool less(T a, T b) { return a < b; }
    static __device__ __forceinline__ uint32_t index(T a) { return (uint32_t)a; }
    static __device__ __forceinline__ T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
};
template <> struct RowItem<KIND_F64> {
    typedef Item128 T;
    static __device__ __forceinline__ T make(double x, uint32_t li) { T t; t.k = flip_f64(x); t.i = li; return t; }
    static __device__ __forceinline__ T inf() { T t; t.k = ~0ull; t.i = ~0ull; return t; }
    static __device__ __forceinline__ bool less(const T& a, const T& b) { return a.k < b.k || (a.k == b.k && a.i < b.i); }
    static __device__ __forceinline__ uint32_t index(const T& a) { return (uint32_t)a.i; }
    static __device__ __forceinline__ T shfl_xor(const T& v, int m) {
        T t; t.k = __shfl_xor_sync(0xffffffffu, v.k, m); t.i = __shfl_xor_sync(0xffffffffu, v.i, m); return t;
    }
};

// pass 1: row of every point + its arrival rank inside the row (the histogram's atomicAdd)
template <int KIND>
__global__ void rowrank_pair_kernel(const __grid_constant__ PairRaw R, uint32_t* __restrict__ rowof,
                                    uint32_t* __restrict__ rank, uint32_t* table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n[0] + R.n[1]) return;
    const int c = i >= R.n[0];
    const uint32_t li = pair_src(R, c, i - (c ? R.n[0] : 0u));
    const double y = load_coord(R.xyz[c], R.dtype[c], R.stride[c], li, 1), z = load_coord(R.xyz[c], R.dtype[c], R.stride[c], li, 2);
    const uint32_t row = R.table_off[c] + (KIND == KIND_INT ? int_row(R.g[c], (int)y, (int)z) : float_row(R.g[c], y, z));
    rowof[i] = row;
    rank[i] = atomicAdd(table + row, 1u);
}

// pass 2 (after the exclusive scan of the table): scatter (x, idx) items into their row segment
template <int KIND>
__global__ void scatter_pair_kernel(const __grid_constant__ PairRaw R, const uint32_t* __restrict__ rowof,
                                    const uint32_t* __restrict__ rank, const uint32_t* __restrict__ table,
                                    typename RowItem<KIND>::T* __restrict__ items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n[0] + R.n[1]) return;
    const int c = i >= R.n[0];
    const uint32_t li = i - (c ? R.n[0] : 0u);
    items[table[rowof[i]] + rank[i]] = RowItem<KIND>::make(load_coord(R.xyz[c], R.dtype[c], R.stride[c], pair_src(R, c, li), 0), li);
}

// pass 3a: one warp per row, rows of <= 32 items: bitonic network over lanes (shuffles, registers
// only, no shared memory -> full occupancy); longer rows are queued.
constexpr int kRowSortThreads = 256;
template <int KIND> struct RowSortCap {          // items a warp / a block sorts in shared memory
    static constexpr uint32_t warp = KIND == KIND_F64 ? 256 : 512;
    static constexpr uint32_t block = KIND == KIND_F64 ? 2048 : 4096;
};
// (segments: row r of the pencil table is [seg_lo[r], seg_hi[r]) with seg_lo = table, seg_hi = table + 1; the buckets
// of split rows come with their own two arrays and a count that only the device knows)
template <int KIND>
__global__ void __launch_bounds__(kRowSortThreads)
rowsort_warp_kernel(const uint32_t* __restrict__ seg_lo, const uint32_t* __restrict__ seg_hi, uint32_t nrows, const uint32_t* nrows_dev,
                    typename RowItem<KIND>::T* __restrict__ items, uint32_t* __restrict__ long_rows, uint32_t* long_count) {
    typedef RowItem<KIND> I;
    const uint32_t row = (blockIdx.x * kRowSortThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= (nrows_dev ? min(nrows, *nrows_dev) : nrows)) return;
    const uint32_t lo = seg_lo[row], hi = seg_hi[row];
    const uint32_t len = hi - lo;
    if (len < 2) return;
    if (len > 32) {
        if (lane == 0) long_rows[atomicAdd(long_count, 1u)] = row;
        return;
    }
    typename I::T v = lane < (int)len ? items[lo + lane] : I::inf();
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const typename I::T o = I::shfl_xor(v, j);
            const bool up = ((lane & k) == 0);            // ascending block
            const bool lower = ((lane & j) == 0);
            const bool take_min = (up == lower);
            const bool o_less = I::less(o, v);
            v = (take_min == o_less) ? o : v;
        }
    }
    if (lane < (int)len) items[lo + lane] = v;
}

// pass 3b: queued rows of 33..RowSortCap::warp items: one warp each, network in the warp's
// private slice of shared memory (no block barrier); a fixed grid walks the queue.
template <int KIND>
__global__ void __launch_bounds__(kRowSortThreads)
rowsort_medium_kernel(const uint32_t* __restrict__ seg_lo, const uint32_t* __restrict__ seg_hi, typename RowItem<KIND>::T* __restrict__ items,
                      const uint32_t* __restrict__ long_rows, const uint32_t* __restrict__ long_count) {
    typedef RowItem<KIND> I;
    constexpr uint32_t kCap = RowSortCap<KIND>::warp;
    __shared__ typename I::T smw[kRowSortThreads / 32][kCap];
    const int lane = threadIdx.x & 31;
    typename I::T* sm = smw[threadIdx.x >> 5];
    const uint32_t count = *long_count;
    const uint32_t nwarps = gridDim.x * (kRowSortThreads / 32);
    for (uint32_t w = (blockIdx.x * kRowSortThreads + threadIdx.x) >> 5; w < count; w += nwarps) {
        const uint32_t row = long_rows[w];
        const uint32_t lo = seg_lo[row], len = seg_hi[row] - lo;
        if (len > kCap) continue;                          // block kernel
        uint32_t p2 = 64;
        while (p2 < len) p2 <<= 1;
        for (uint32_t i = lane; i < p2; i += 32) sm[i] = i < len ? items[lo + i] : I::inf();
        __syncwarp();
        for (uint32_t k = 2; k <= p2; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t t = lane; t < (p2 >> 1); t += 32) {
                    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
                    const uint32_t l = i | j;
                    const typename I::T a = sm[i], b = sm[l];
                    const bool up = (i & k) == 0;
                    if (I::less(b, a) == up) { sm[i] = b; sm[l] = a; }
                }
                __syncwarp();
            }
        for (uint32_t i = lane; i < len; i += 32) items[lo + i] = sm[i];
        __syncwarp();
    }
}

// pass 4: one block per long row: bitonic sort in shared memory (<= RowSortCap::block items) or,
// for degenerate inputs (e.g. thousands of duplicates of one voxel column), in global memory
template <int KIND>
__global__ void __launch_bounds__(kRowSortThreads)
rowsort_block_kernel(const uint32_t* __restrict__ seg_lo, const uint32_t* __restrict__ seg_hi, typename RowItem<KIND>::T* __restrict__ items,
                     const uint32_t* __restrict__ long_rows, const uint32_t* __restrict__ long_count, int allow_global) {
    typedef RowItem<KIND> I;
    constexpr uint32_t kCap = RowSortCap<KIND>::block;
    __shared__ typename I::T sm[kCap];
    const uint32_t count = *long_count;
    for (uint32_t w = blockIdx.x; w < count; w += gridDim.x) {
        const uint32_t row = long_rows[w];
        const uint32_t lo = seg_lo[row], len = seg_hi[row] - lo;
        if (len <= RowSortCap<KIND>::warp) continue;       // done by rowsort_medium_kernel
        if (len > kCap && !allow_global) continue;         // split into x buckets first (rowsplit_* kernels)
        uint32_t p2 = 64;
        while (p2 < len) p2 <<= 1;
        if (p2 <= kCap) {
            for (uint32_t i = threadIdx.x; i < p2; i += kRowSortThreads) sm[i] = i < len ? items[lo + i] : I::inf();
            __syncthreads();
            for (uint32_t k = 2; k <= p2; k <<= 1)
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    for (uint32_t i = threadIdx.x; i < p2; i += kRowSortThreads) {
                        const uint32_t l = i ^ j;
                        if (l > i) {
                            const typename I::T a = sm[i], b = sm[l];
                            const bool up = (i & k) == 0;
                            if (I::less(b, a) == up) { sm[i] = b; sm[l] = a; }
                        }
                    }
                    __syncthreads();
                }
            for (uint32_t i = threadIdx.x; i < len; i += kRowSortThreads) items[lo + i] = sm[i];
            __syncthreads();
        } else {
            // global-memory network in the all-ascending form (first sub-stage of every merge
            // pairs i with i ^ (k - 1)): the minimum always goes to the lower index, so the
            // virtual +inf padding beyond len never has to move and is simply skipped
            typename I::T* a = items + lo;
            for (uint32_t k = 2; k <= p2; k <<= 1) {
                for (uint32_t i = threadIdx.x; i < len; i += kRowSortThreads) {
                    const uint32_t l = i ^ (k - 1);
                    if (l > i && l < len) { const typename I::T x = a[i], y = a[l]; if (I::less(y, x)) { a[i] = y; a[l] = x; } }
                }
                __syncthreads();
                for (uint32_t j = k >> 2; j > 0; j >>= 1) {
                    for (uint32_t i = threadIdx.x; i < len; i += kRowSortThreads) {
                        const uint32_t l = i ^ j;
                        if (l > i && l < len) { const typename I::T x = a[i], y = a[l]; if (I::less(y, x)) { a[i] = y; a[l] = x; } }
                    }
                    __syncthreads();
                }
            }
        }
    }
}

// Rows longer than a block can sort in shared memory (a LiDAR ground plane puts tens of thousands of points into one
// pencil) are first cut into BUCKETS of x -- a second counting sort inside the row: bucket = position of the x key in
// the row's [min, max] key range, about kSplitTarget items each -- and the buckets are then sorted like rows.
constexpr uint32_t kSplitTarget = 256;
struct SplitRow {
    uint32_t row, lo, len, nb, boff;
    unsigned long long kmin, kmax;
};
template <int KIND> __device__ __forceinline__ unsigned long long item_key(const typename RowItem<KIND>::T& v);
template <> __device__ __forceinline__ unsigned long long item_key<KIND_INT>(const unsigned long long& v) { return v >> 32; }
template <> __device__ __forceinline__ unsigned long long item_key<KIND_F32>(const unsigned long long& v) { return v >> 32; }
template <> __device__ __forceinline__ unsigned long long item_key<KIND_F64>(const Item128& v) { return v.k; }

__device__ __forceinline__ uint32_t split_bucket(const SplitRow& S, unsigned long long key) {
    const double span = (double)(S.kmax - S.kmin) + 1.0;
    const uint32_t b = (uint32_t)((double)(key - S.kmin) / span * (double)S.nb);       // monotone in key
    return b < S.nb ? b : S.nb - 1;
}

// which rows are split, and where their buckets live (counters[0] = split rows, counters[1] = buckets)
template <int KIND>
__global__ void rowsplit_find_kernel(const uint32_t* __restrict__ table, uint32_t nrows, SplitRow* __restrict__ rows, uint32_t* counters) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const uint32_t lo = table[r], len = table[r + 1] - lo;
    if (len <= RowSortCap<KIND>::block) return;
    SplitRow S;
    S.row = r; S.lo = lo; S.len = len;
    S.nb = (len + kSplitTarget - 1) / kSplitTarget;
    S.boff = atomicAdd(counters + 1, S.nb);
    S.kmin = 0; S.kmax = 0;
    rows[atomicAdd(counters, 1u)] = S;
}
// per split row (one block each): key range; bucket of every item + its arrival rank; bucket boundaries; scatter
template <int KIND>
__global__ void __launch_bounds__(256)
rowsplit_kernel(SplitRow* __restrict__ rows, const uint32_t* __restrict__ counters, const typename RowItem<KIND>::T* __restrict__ items,
                typename RowItem<KIND>::T* __restrict__ items_alt, uint32_t* __restrict__ bucket_of, uint32_t* __restrict__ rank_of,
                uint32_t* __restrict__ bcount, uint32_t* __restrict__ bstart, uint32_t* __restrict__ bend) {
    __shared__ unsigned long long s_k[2][8];
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nsplit = counters[0];
    for (uint32_t w = blockIdx.x; w < nsplit; w += gridDim.x) {
        SplitRow S = rows[w];
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (uint32_t i = threadIdx.x; i < S.len; i += 256) {
            const unsigned long long k = item_key<KIND>(items[S.lo + i]);
            kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
            kmin = a < kmin ? a : kmin; kmax = b > kmax ? b : kmax;
        }
        __syncthreads();
        if (lane == 0) { s_k[0][warp] = kmin; s_k[1][warp] = kmax; }
        __syncthreads();
        for (int q = 0; q < 8; ++q) { kmin = s_k[0][q] < kmin ? s_k[0][q] : kmin; kmax = s_k[1][q] > kmax ? s_k[1][q] : kmax; }
        S.kmin = kmin; S.kmax = kmax;
        // buckets + arrival ranks (bcount is zero on entry)
        for (uint32_t i = threadIdx.x; i < S.len; i += 256) {
            const uint32_t b = split_bucket(S, item_key<KIND>(items[S.lo + i]));
            bucket_of[S.lo + i] = b;
            rank_of[S.lo + i] = atomicAdd(bcount + S.boff + b, 1u);
        }
        __syncthreads();
        // exclusive prefix of the bucket counts -> [bstart, bend) of every bucket, absolute positions
        if (threadIdx.x == 0) s_run = S.lo;
        __syncthreads();
        for (uint32_t b0 = 0; b0 < S.nb; b0 += 256) {
            const uint32_t b = b0 + threadIdx.x;
            const uint32_t c = b < S.nb ? bcount[S.boff + b] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_w[warp] = incl;
            __syncthreads();
            uint32_t base = s_run;
            for (int q = 0; q < warp; ++q) base += s_w[q];
            if (b < S.nb) { bstart[S.boff + b] = base + incl - c; bend[S.boff + b] = base + incl; }
            __syncthreads();
            if (threadIdx.x == 255) s_run = base + incl;
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < S.len; i += 256)
            items_alt[bstart[S.boff + bucket_of[S.lo + i]] + rank_of[S.lo + i]] = items[S.lo + i];
        __syncthreads();
    }
}
template <int KIND>
__global__ void __launch_bounds__(256)
rowsplit_copyback_kernel(const SplitRow* __restrict__ rows, const uint32_t* __restrict__ counters, typename RowItem<KIND>::T* __restrict__ items,
                         const typename RowItem<KIND>::T* __restrict__ items_alt) {
    const uint32_t nsplit = counters[0];
    for (uint32_t w = blockIdx.x; w < nsplit; w += gridDim.x) {
        const uint32_t lo = rows[w].lo, len = rows[w].len;
        for (uint32_t i = threadIdx.x; i < len; i += 256) items[lo + i] = items_alt[lo + i];
    }
}

// records from sorted (x, idx) items
template <class K>
__global__ void reorder_items_pair_kernel(const __grid_constant__ PairRaw R, const typename RowItem<K::kind>::T* __restrict__ items,
                                          typename K::Rec* __restrict__ recs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n[0] + R.n[1]) return;
    const int c = i >= R.n[0];
    const uint32_t src = RowItem<K::kind>::index(items[i]);
    uint32_t rgba = 0;
    if (R.rgb_in_rec[c]) {
        if (R.rgb_dtype[c] == PCCM_U8) {
            const uint8_t* p = static_cast<const uint8_t*>(R.rgb[c]) + (int64_t)src * R.rgb_stride[c];
            rgba = p[0] | (p[1] << 8) | (p[2] << 16);
        } else {
            rgba = (uint32_t)rint(load_coord(R.rgb[c], PCCM_F64, R.rgb_stride[c], src, 0) * 255.0) |
                   ((uint32_t)rint(load_coord(R.rgb[c], PCCM_F64, R.rgb_stride[c], src, 1) * 255.0) << 8) |
                   ((uint32_t)rint(load_coord(R.rgb[c], PCCM_F64, R.rgb_stride[c], src, 2) * 255.0) << 16);
        }
    }
    const uint32_t row = pair_src(R, c, src);       // where the coordinates of point src live
    recs[i] = RecPack<K>::make(load_coord(R.xyz[c], R.dtype[c], R.stride[c], row, 0), load_coord(R.xyz[c], R.dtype[c], R.stride[c], row, 1),
                               load_coord(R.xyz[c], R.dtype[c], R.stride[c], row, 2), src, rgba);
}

// ------------------------------------------------------------------------------------
// deterministic block reductions
// ------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* sm) {
    for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_down_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < THREADS / 32; ++w) r = dadd(r, sm[w]);
    return r;  // valid in thread 0
}
template <int THREADS>
__device__ __forceinline__ double block_max(double v, double* sm) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = -INFINITY;
    if (threadIdx.x == 0)
        for (int w = 0; w < THREADS / 32; ++w) r = fmax(r, sm[w]);
    return r;
}
template <int THREADS>
__device__ __forceinline__ double block_min(double v, double* sm) { return -block_max<THREADS>(-v, sm); }
template <int THREADS>
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < THREADS / 32; ++w) r += sm[w];
    return r;
}

// ------------------------------------------------------------------------------------
// K5: fused NN query + D1 / D2 / colour epilogue + block reduction
// ------------------------------------------------------------------------------------
struct CloudView {        // device-side view of an indexed cloud
    RowGrid grid;
    const void* recs;
    const uint32_t* row_start;
    int32_t rgb_mode;       // 0 none, 1 packed in the record (KInt), 2 rgb_u8 array, 3 rgb_f64 array
    const double* lut255;   // lut255[k] == (double)k / 255.0 (host-computed: same IEEE quotient, no device division)
    const uchar4* rgb_u8;   // original order, or null
    const double* rgb_f64;  // original order [n][3], or null
    const double* normals;  // original order [n][3], or null
};

struct BlockPartial {   // one reduced record (per warp while the kernel runs, per direction at the end)
    unsigned long long sum_d1_u64;
    double sum_d1, max_d1, sum_d2, max_d2, csum[3], cmax[3];
};

struct DirParams {
    CloudView q, s;
    uint32_t qbegin, qend;    // slice of the query cloud's sorted order
    uint32_t ntiles;          // ceil((qend - qbegin) / kQueryThreads)
    uint32_t flags;
    int32_t* idx_out;         // original query order, or null
    double* d2_out;
};

struct QueryParams {
    DirParams dir[2];
    int32_t ndirs;
    int32_t normals_mode;
    double T[9];
    double color_scale;
    BlockPartial* partials;   // one record per tile: direction d at [d * rec_stride + tile]
    uint32_t rec_stride;
    BlockPartial* chunks;     // [ndirs * kFinalChunks] scratch of the fold
    unsigned int* ticket;     // zero on entry; the last fold block resets it
    BlockPartial* out;        // [2]
};

__device__ __forceinline__ void load_color(const CloudView& c, uint32_t idx, uint32_t packed, double* out) {
    if (c.rgb_mode == 3) {
        out[0] = __ldg(c.rgb_f64 + 3 * (size_t)idx);
        out[1] = __ldg(c.rgb_f64 + 3 * (size_t)idx + 1);
        out[2] = __ldg(c.rgb_f64 + 3 * (size_t)idx + 2);
        return;
    }
    uint32_t p = packed;
    if (c.rgb_mode == 2) { uchar4 u = __ldg(c.rgb_u8 + idx); p = u.x | (u.y << 8) | (u.z << 16); }
    out[0] = c.lut255[p & 0xffu];             // generic loads: the table may sit in shared memory
    out[1] = c.lut255[(p >> 8) & 0xffu];
    out[2] = c.lut255[(p >> 16) & 0xffu];
}

template <class K> __device__ __forceinline__ uint32_t rec_rgba(const typename K::Rec&) { return 0; }
template <> __device__ __forceinline__ uint32_t rec_rgba<KInt>(const uint4& r) { return r.w; }

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void partial_init(BlockPartial& a) {
    a.sum_d1_u64 = 0; a.sum_d1 = 0; a.sum_d2 = 0; a.max_d1 = -INFINITY; a.max_d2 = -INFINITY;
    for (int k = 0; k < 3; ++k) { a.csum[k] = 0; a.cmax[k] = -INFINITY; }
}
__device__ __forceinline__ void partial_merge(BlockPartial& a, const BlockPartial& b) {
    a.sum_d1_u64 += b.sum_d1_u64;
    a.sum_d1 = dadd(a.sum_d1, b.sum_d1);
    a.sum_d2 = dadd(a.sum_d2, b.sum_d2);
    a.max_d1 = fmax(a.max_d1, b.max_d1);
    a.max_d2 = fmax(a.max_d2, b.max_d2);
    for (int k = 0; k < 3; ++k) { a.csum[k] = dadd(a.csum[k], b.csum[k]); a.cmax[k] = fmax(a.cmax[k], b.cmax[k]); }
}
__device__ __forceinline__ void partial_warp_reduce(BlockPartial& a, uint32_t flags) {
    a.sum_d1_u64 = warp_sum_u64(a.sum_d1_u64);
    a.sum_d1 = warp_sum(a.sum_d1);
    a.max_d1 = warp_max(a.max_d1);
    if (flags & PCCM_EVAL_D2) { a.sum_d2 = warp_sum(a.sum_d2); a.max_d2 = warp_max(a.max_d2); }
    if (flags & PCCM_EVAL_COLOR)
        for (int k = 0; k < 3; ++k) { a.csum[k] = warp_sum(a.csum[k]); a.cmax[k] = warp_max(a.cmax[k]); }
}

// Four values reduced over the 32 lanes with 6 exchanges instead of 20: at offsets 16 and 8 each
// lane hands half of its values to its partner and keeps the other half (so the number of live
// values halves while the number of lanes folded doubles), then a plain butterfly on the last
// one.  Result: every lane ends with the total of value (lane >> 3) & 3 ... the caller reads
// value j from lane 8 * j.  The pairing is fixed, so float sums are reproducible.
template <bool IS_MAX>
__device__ __forceinline__ double warp_reduce4(double v0, double v1, double v2, double v3) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    auto op = [](double a, double b) { return IS_MAX ? fmax(a, b) : dadd(a, b); };
    // offset 16: lanes with bit 4 clear keep (v0, v1), the others keep (v2, v3)
    const bool hi16 = (lane & 16) != 0;
    double s0 = hi16 ? v0 : v2, s1 = hi16 ? v1 : v3;          // what I give away
    double k0 = hi16 ? v2 : v0, k1 = hi16 ? v3 : v1;          // what I keep
    k0 = op(k0, __shfl_xor_sync(full, s0, 16));
    k1 = op(k1, __shfl_xor_sync(full, s1, 16));
    // offset 8: lanes with bit 3 clear keep k0, the others keep k1
    const bool hi8 = (lane & 8) != 0;
    double g = hi8 ? k0 : k1, k = hi8 ? k1 : k0;
    k = op(k, __shfl_xor_sync(full, g, 8));
    k = op(k, __shfl_xor_sync(full, k, 4));
    k = op(k, __shfl_xor_sync(full, k, 2));
    k = op(k, __shfl_xor_sync(full, k, 1));
    return k;   // lanes 0-7: v0, 8-15: v1, 16-23: v2, 24-31: v3
}

// Query-kernel variant: one lane = one query.  Integer clouds reduce D1 with three REDUX
// instructions (d2 < 2^32 is split in 16-bit halves so the 32-lane sums cannot overflow); the
// float64 sums / maxima (plane error + three colour channels) go through warp_reduce4.
template <class K>
__device__ __forceinline__ void query_warp_reduce(BlockPartial& a, uint32_t flags, uint32_t d2_int, bool active) {
    const unsigned full = 0xffffffffu;
    if (K::kind == KIND_INT) {
        const uint32_t v = active ? d2_int : 0u;
        const unsigned lo = __reduce_add_sync(full, v & 0xffffu), hi = __reduce_add_sync(full, v >> 16);
        a.sum_d1_u64 = (unsigned long long)lo + ((unsigned long long)hi << 16);
        const unsigned mx = __reduce_max_sync(full, v);
        a.max_d1 = __any_sync(full, active) ? (double)mx : -INFINITY;
        a.sum_d1 = 0;
    } else {
        a.sum_d1 = warp_sum(a.sum_d1);
        a.max_d1 = warp_max(a.max_d1);
    }
    if (flags & (PCCM_EVAL_D2 | PCCM_EVAL_COLOR)) {
        const double s = warp_reduce4<false>(a.sum_d2, a.csum[0], a.csum[1], a.csum[2]);
        const double m = warp_reduce4<true>(a.max_d2, a.cmax[0], a.cmax[1], a.cmax[2]);
        a.sum_d2 = __shfl_sync(full, s, 0);  a.csum[0] = __shfl_sync(full, s, 8);
        a.csum[1] = __shfl_sync(full, s, 16); a.csum[2] = __shfl_sync(full, s, 24);
        a.max_d2 = __shfl_sync(full, m, 0);  a.cmax[0] = __shfl_sync(full, m, 8);
        a.cmax[1] = __shfl_sync(full, m, 16); a.cmax[2] = __shfl_sync(full, m, 24);
    }
}

// ------------------------------------------------------------------------------------
// shared-memory staging of the pencils around a tile
//
// ncu on the first version: DRAM 5 %, issue slots 65 %, 17 cycles per issued instruction --
// the kernel waits on chains of DEPENDENT loads (table entry -> 4-5 binary-search probes ->
// sweep), each a 250-cycle L2 round trip because consecutive tiles land on different SMs.
// A tile's 128 queries are neighbours in (row, x) order, so the rows they can touch in rings
// 0 and 1 form, for each of <= kMaxSeg z-lines, ONE contiguous run of the record array.  The
// block copies those runs (16-byte coalesced loads, typically 0.6 k records = 10 KB) and the
// matching slice of the row table into shared memory once, and every probe of the search then
// costs a shared-memory access.  Rows outside the window (ring >= 2, oversized tiles) fall
// back to global memory transparently -- results are identical by construction.
// ------------------------------------------------------------------------------------
#ifndef PCCM_STAGE_BYTES
#define PCCM_STAGE_BYTES (24 * 1024)
#endif
constexpr int kMaxSeg = 4;
constexpr int kMaxSegRows = 96;

template <class K>
struct StagedRows {
    typedef typename K::Rec Rec;
    const uint32_t* row_start;   // global fallback
    const Rec* recs;
    int ny;
    int z0, nseg, ylo, yhi;      // window: zz in [z0, z0 + nseg), yy in [ylo, yhi]
    const uint32_t* tbl;         // shared: [nseg][kMaxSegRows + 1] absolute record positions
    const Rec* const* vbase;     // shared: vbase[k][pos] is the staged copy of recs[pos]
    __device__ __forceinline__ const Rec* fetch(int yy, int zz, uint32_t& lo, uint32_t& hi) const {
        const int k = zz - z0;
        if ((unsigned)k < (unsigned)nseg && yy >= ylo && yy <= yhi) {
            const uint32_t* t = tbl + k * (kMaxSegRows + 1) + (yy - ylo);
            lo = t[0];
            hi = t[1];
            return vbase[k];
        }
        const uint32_t row = (uint32_t)zz * (uint32_t)ny + (uint32_t)yy;
        lo = __ldg(row_start + row);
        hi = __ldg(row_start + row + 1);
        return recs;
    }
    static __device__ __forceinline__ Rec load(const Rec* p) { return *p; }            // generic: shared or global
    static __device__ __forceinline__ typename K::C xof(const Rec* p) { return K::rec_x(p); }
};

template <class K>
struct StageSmem {
    typename K::Rec recs[PCCM_STAGE_BYTES / sizeof(typename K::Rec)];
    uint32_t tbl[kMaxSeg][kMaxSegRows + 1];
    const typename K::Rec* vbase[kMaxSeg];
    uint32_t gstart[kMaxSeg], count[kMaxSeg], sbase[kMaxSeg];
    int red[4][kQueryThreads / 32];
    int z0, nseg, ylo, yhi;
};

// Block-wide: decide the window from the queries' cells, copy it.  Returns through `S`
// (nseg == 0 means "not staged": every fetch goes to global memory).
template <class K>
__device__ __forceinline__ void stage_window(const RowGrid& g, const uint32_t* __restrict__ row_start,
                                             const typename K::Rec* __restrict__ recs, bool active,
                                             const typename K::Q& q, StageSmem<K>& S) {
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ny = g.ny, nz = g.nz;
    int cy = K::cell_y(g, q.y), cz = K::cell_z(g, q.z);
    cy = cy < 0 ? 0 : (cy >= ny ? ny - 1 : cy);
    cz = cz < 0 ? 0 : (cz >= nz ? nz - 1 : cz);
    const int big = 0x7fffffff;
    const int ymn = __reduce_min_sync(full, active ? cy : big), ymx = __reduce_max_sync(full, active ? cy : -1);
    const int zmn = __reduce_min_sync(full, active ? cz : big), zmx = __reduce_max_sync(full, active ? cz : -1);
    if (lane == 0) { S.red[0][warp] = ymn; S.red[1][warp] = ymx; S.red[2][warp] = zmn; S.red[3][warp] = zmx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = S.red[0][0], b = S.red[1][0], c = S.red[2][0], d = S.red[3][0];
        for (int w = 1; w < kQueryThreads / 32; ++w) {
            a = min(a, S.red[0][w]); b = max(b, S.red[1][w]); c = min(c, S.red[2][w]); d = max(d, S.red[3][w]);
        }
        int nseg = 0;
        if (b >= 0) {   // at least one active query
            const int ylo = max(a - 1, 0), yhi = min(b + 1, ny - 1), z0 = max(c - 1, 0), z1 = min(d + 1, nz - 1);
            const int nrows = yhi - ylo + 1;
            nseg = z1 - z0 + 1;
            if (nseg > kMaxSeg || nrows > kMaxSegRows) nseg = 0;
            uint32_t total = 0;
            for (int k = 0; k < nseg; ++k) {
                const uint32_t first = (uint32_t)(z0 + k) * (uint32_t)ny + (uint32_t)ylo;
                const uint32_t gs = __ldg(row_start + first), ge = __ldg(row_start + first + nrows);
                S.gstart[k] = gs; S.count[k] = ge - gs; S.sbase[k] = total;
                total += ge - gs;
            }
            if (total > PCCM_STAGE_BYTES / sizeof(typename K::Rec)) nseg = 0;
            S.z0 = z0; S.ylo = ylo; S.yhi = yhi;
        }
        S.nseg = nseg;
    }
    __syncthreads();
    const int nseg = S.nseg;
    if (nseg == 0) return;
    const int nrows = S.yhi - S.ylo + 1;
    for (int k = 0; k < nseg; ++k) {
        const uint32_t first = (uint32_t)(S.z0 + k) * (uint32_t)ny + (uint32_t)S.ylo;
        for (int j = threadIdx.x; j <= nrows; j += kQueryThreads) S.tbl[k][j] = __ldg(row_start + first + j);
        const uint32_t gs = S.gstart[k], cnt = S.count[k], sb = S.sbase[k];
        for (uint32_t i = threadIdx.x; i < cnt; i += kQueryThreads) S.recs[sb + i] = load_rec(recs + gs + i);
        if (threadIdx.x == 0) S.vbase[k] = S.recs + sb - gs;   // virtual base (only dereferenced inside the copied range)
    }
    __syncthreads();
}

// PCCM_EVAL_TIE_AVERAGE: second walk of the search with the minimal distance known -- every point AT that distance
// contributes its plane error (its own normal) and its colour.
template <class K>
struct TieSum {
    typename K::D target;
    const CloudView* s;
    const typename K::Rec* srecs;
    typename K::Q q;
    uint32_t flags;
    uint32_t cnt;
    double pe_sum, c_sum[3];
    __device__ __forceinline__ typename K::D worst() const { return target; }
    __device__ __forceinline__ void offer(typename K::D d, uint32_t idx, uint32_t pos) {
        if (d != target) return;
        ++cnt;
        const typename K::Rec nr = load_rec(srecs + pos);
        if (flags & PCCM_EVAL_D2) {
            const typename K::Q nq = K::rec_q(nr);
            const double e[3] = {dsub((double)q.x, (double)nq.x), dsub((double)q.y, (double)nq.y), dsub((double)q.z, (double)nq.z)};
            const double nv[3] = {__ldg(s->normals + 3 * (size_t)idx), __ldg(s->normals + 3 * (size_t)idx + 1), __ldg(s->normals + 3 * (size_t)idx + 2)};
            pe_sum = dadd(pe_sum, plane_err2(e, nv));
        }
        if (flags & PCCM_EVAL_COLOR) {
            double cn[3];
            load_color(*s, idx, rec_rgba<K>(nr), cn);
            for (int k = 0; k < 3; ++k) c_sum[k] = dadd(c_sum[k], cn[k]);
        }
    }
};

// One launch covers both directions: blocks [0, dir[0].ntiles) serve direction 0, the rest
// direction 1; a block is one tile of kQueryThreads consecutive queries of the sorted order
// (the hardware block scheduler balances the very uneven tile costs).  Each warp reduces its 32
// queries with shuffles only -- no block barrier -- and writes ONE record at a position fixed
// by its tile, so the floating-point sums do not depend on scheduling.
#ifdef PCCM_QBLOCKS
#define PCCM_QUERY_BOUNDS __launch_bounds__(kQueryThreads, PCCM_QBLOCKS)
#else
#define PCCM_QUERY_BOUNDS __launch_bounds__(kQueryThreads)
#endif
template <class K>
__global__ void PCCM_QUERY_BOUNDS
pair_query_kernel(const __grid_constant__ QueryParams P) {
    typedef typename K::Rec Rec;
    typedef typename K::Q Q;
    constexpr int kWarps = kQueryThreads / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = (P.ndirs > 1 && blockIdx.x >= P.dir[0].ntiles) ? 1 : 0;
    const DirParams& D = P.dir[d];
    const uint32_t tile = blockIdx.x - (d ? P.dir[0].ntiles : 0u);
    const Rec* __restrict__ qrecs = static_cast<const Rec*>(D.q.recs);
    const Rec* __restrict__ srecs = static_cast<const Rec*>(D.s.recs);
    BlockPartial acc;
    partial_init(acc);
    const uint32_t t = D.qbegin + tile * kQueryThreads + threadIdx.x;
    const bool active = t < D.qend;
    const Rec qr = load_rec(qrecs + (active ? t : D.qbegin));
    const Q q = K::rec_q(qr);
    const uint32_t qidx = K::rec_idx(qr);
    Best1<K> best;
    best.init();
#if PCCM_STAGE
    __shared__ StageSmem<K> stg;
    stage_window<K>(D.s.grid, D.s.row_start, srecs, active, q, stg);
    if (active) {
        StagedRows<K> rows;
        rows.row_start = D.s.row_start; rows.recs = srecs; rows.ny = D.s.grid.ny;
        rows.z0 = stg.z0; rows.nseg = stg.nseg; rows.ylo = stg.ylo; rows.yhi = stg.yhi;
        rows.tbl = &stg.tbl[0][0]; rows.vbase = stg.vbase;
        search_rows<K>(D.s.grid, rows, q, best);
    }
#else
    if (active) search<K>(D.s.grid, D.s.row_start, srecs, q, best);
#endif
    if (active) {
        const double d1 = K::d2_as_double(best.d2);
        if (K::kind == KIND_INT) acc.sum_d1_u64 = (unsigned long long)best.d2;
        else acc.sum_d1 = d1;
        acc.max_d1 = d1;
        if (D.idx_out) D.idx_out[qidx] = (int32_t)best.idx;
        if (D.d2_out) D.d2_out[qidx] = d1;
        if ((D.flags & PCCM_EVAL_TIE_AVERAGE) && (D.flags & (PCCM_EVAL_D2 | PCCM_EVAL_COLOR))) {
            TieSum<K> ts;
            ts.target = best.d2; ts.s = &D.s; ts.srecs = srecs; ts.q = q; ts.flags = D.flags; ts.cnt = 0;
            ts.pe_sum = 0; ts.c_sum[0] = ts.c_sum[1] = ts.c_sum[2] = 0;
            search<K>(D.s.grid, D.s.row_start, srecs, q, ts);
            const double inv = 1.0 / (double)ts.cnt;           // (cnt >= 1: the best point itself)
            if (D.flags & PCCM_EVAL_D2) { acc.sum_d2 = dmul(ts.pe_sum, inv); acc.max_d2 = acc.sum_d2; }
            if (D.flags & PCCM_EVAL_COLOR) {
                double cq[3], cn[3] = {dmul(ts.c_sum[0], inv), dmul(ts.c_sum[1], inv), dmul(ts.c_sum[2], inv)};
                load_color(D.q, qidx, rec_rgba<K>(qr), cq);
                color_diff2(P.T, cq, cn, P.color_scale, acc.csum, acc.cmax);
            }
        } else if (D.flags & (PCCM_EVAL_D2 | PCCM_EVAL_COLOR)) {
            const Rec nr = load_rec(srecs + best.pos);
            if (D.flags & PCCM_EVAL_D2) {
                const Q nq = K::rec_q(nr);
                double e[3] = {dsub((double)q.x, (double)nq.x), dsub((double)q.y, (double)nq.y), dsub((double)q.z, (double)nq.z)};
                const uint32_t ni = P.normals_mode == PCCM_NORMALS_BY_NEIGHBOUR ? best.idx : qidx;
                double nv[3] = {__ldg(D.s.normals + 3 * (size_t)ni), __ldg(D.s.normals + 3 * (size_t)ni + 1),
                                __ldg(D.s.normals + 3 * (size_t)ni + 2)};
                acc.sum_d2 = plane_err2(e, nv);
                acc.max_d2 = acc.sum_d2;
            }
            if (D.flags & PCCM_EVAL_COLOR) {
                double cq[3], cn[3];
                load_color(D.q, qidx, rec_rgba<K>(qr), cq);
                load_color(D.s, best.idx, rec_rgba<K>(nr), cn);
                color_diff2(P.T, cq, cn, P.color_scale, acc.csum, acc.cmax);
            }
        }
    }
    // warps fold with shuffles as they finish; the block's single record is written once all
    // four are done (they could not retire earlier anyway: the block holds their resources)
    query_warp_reduce<K>(acc, D.flags, active ? (uint32_t)acc.sum_d1_u64 : 0u, active);
    __shared__ BlockPartial sm[kWarps];
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockPartial r = sm[0];
        for (int w = 1; w < kWarps; ++w) partial_merge(r, sm[w]);
        P.partials[(size_t)d * P.rec_stride + tile] = r;
    }
}

// K8: fixed-order fold of the per-warp records.  gridDim.x = ndirs * kFinalChunks; block
// (d, c) folds chunk c of direction d; the last block to finish folds the chunk results.
constexpr int kFinalThreads = 256;
constexpr int kFinalChunks = 32;

__device__ __forceinline__ void block_fold(BlockPartial& a, BlockPartial* sm, BlockPartial* dst) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    partial_warp_reduce(a, PCCM_EVAL_D2 | PCCM_EVAL_COLOR);
    __syncthreads();
    if (lane == 0) sm[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockPartial r = sm[0];
        for (int w = 1; w < kFinalThreads / 32; ++w) partial_merge(r, sm[w]);
        *dst = r;
    }
}

// a few hundred records per direction (the brick path's one-wave epilogue): one block per direction, no second phase
__global__ void __launch_bounds__(kFinalThreads) finalize_small_kernel(const __grid_constant__ QueryParams P) {
    pdl_enter();
    __shared__ BlockPartial sm[kFinalThreads / 32];
    const int d = blockIdx.x;
    const uint32_t nrec = P.dir[d].ntiles;
    const BlockPartial* in = P.partials + (size_t)d * P.rec_stride;
    BlockPartial a;
    partial_init(a);
    for (uint32_t i = threadIdx.x; i < nrec; i += kFinalThreads) partial_merge(a, in[i]);
    block_fold(a, sm, P.out + d);
}

__global__ void __launch_bounds__(kFinalThreads) finalize_kernel(const __grid_constant__ QueryParams P) {
    pdl_enter();
    __shared__ BlockPartial sm[kFinalThreads / 32];
    __shared__ bool is_last;
    const int d = blockIdx.x / kFinalChunks, c = blockIdx.x % kFinalChunks;
    const uint32_t nrec = P.dir[d].ntiles;
    const uint32_t per = (nrec + kFinalChunks - 1) / kFinalChunks;
    const uint32_t lo = c * per, hi = lo + per < nrec ? lo + per : nrec;
    const BlockPartial* in = P.partials + (size_t)d * P.rec_stride;
    BlockPartial a;
    partial_init(a);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += kFinalThreads) partial_merge(a, in[i]);
    block_fold(a, sm, P.chunks + blockIdx.x);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(P.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int dd = 0; dd < P.ndirs; ++dd) {
        partial_init(a);
        if (threadIdx.x < kFinalChunks) a = P.chunks[dd * kFinalChunks + threadIdx.x];
        block_fold(a, sm, P.out + dd);
    }
    if (threadIdx.x == 0) *P.ticket = 0;
}

// ------------------------------------------------------------------------------------
// K7: self k-NN with fused epilogues
// ------------------------------------------------------------------------------------
enum KnnMode : int { KNN_LIST = 0, KNN_BOUNDARY = 1, KNN_NORMALS = 2, KNN_NORMALS_FLAGGED = 3 };

struct KnnParams {
    CloudView c;
    uint32_t begin, end;   // range of sorted order
    int32_t k;
    int32_t mode;
    int32_t* idx_out;      // KNN_LIST: [n][k] original order
    double* d2_out;        // KNN_LIST: [n][k];  KNN_BOUNDARY: optional per-point sqrt distance [n]
    double* normals_out;   // KNN_NORMALS: [n][3] original order
    double* minmax;        // KNN_BOUNDARY: per-block {min, max}
    uint32_t* todo;        // KNN_NORMALS: sorted positions normals_int_kernel left to the list kernel
    uint32_t* todo_count;
};

// one point of the self k-NN: search + the epilogue selected by P.mode
template <class K>
__device__ __forceinline__ void knn_point(const KnnParams& P, uint32_t t, typename K::D* d2s, uint32_t* idxs, uint32_t* poss,
                                          double& bmin, double& bmax) {
    typedef typename K::Rec Rec;
    typedef typename K::Q Q;
    const int k = P.k;
    const Rec* __restrict__ recs = static_cast<const Rec*>(P.c.recs);
    const Rec qr = load_rec(recs + t);
    const Q q = K::rec_q(qr);
    const uint32_t qidx = K::rec_idx(qr);
    if (P.mode == KNN_BOUNDARY) {
        // ComputeNearestNeighborDistance: sqrt of the second entry of the 2-NN result, 0 when absent
        Best2Val<K> b2;
        b2.init();
        search<K>(P.c.grid, P.c.row_start, recs, q, b2);
        const double v = b2.count > 1 ? sqrt(K::d2_as_double(b2.m2)) : 0.0;
        bmin = v; bmax = v;
        if (P.d2_out) P.d2_out[qidx] = v;
        return;
    }
    TopK<K> acc;
    acc.init(d2s + threadIdx.x, idxs + threadIdx.x, poss + threadIdx.x, kKnnThreads, k);
    search<K>(P.c.grid, P.c.row_start, recs, q, acc);
    if (P.mode == KNN_LIST) {
        for (int j = 0; j < k; ++j) {
            const bool have = j < acc.count;
            P.idx_out[(size_t)qidx * k + j] = have ? (int32_t)idxs[j * kKnnThreads + threadIdx.x] : -1;
            P.d2_out[(size_t)qidx * k + j] = have ? K::d2_as_double(d2s[j * kKnnThreads + threadIdx.x]) : INFINITY;
        }
    } else {
        double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < acc.count; ++j) {
            const Rec nr = load_rec(recs + poss[j * kKnnThreads + threadIdx.x]);
            const Q nq = K::rec_q(nr);
            cumulant_add(cum, (double)nq.x, (double)nq.y, (double)nq.z);
        }
        double nv[3];
        normal_from_cumulants(cum, acc.count, nv);
        P.normals_out[3 * (size_t)qidx + 0] = nv[0];
        P.normals_out[3 * (size_t)qidx + 1] = nv[1];
        P.normals_out[3 * (size_t)qidx + 2] = nv[2];
    }
}

template <class K>
__global__ void __launch_bounds__(kKnnThreads)
knn_self_kernel(const __grid_constant__ KnnParams P) {
    typedef typename K::D D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    D* d2s = reinterpret_cast<D*>(smem_raw);
    uint32_t* idxs = reinterpret_cast<uint32_t*>(d2s + (size_t)P.k * kKnnThreads);
    uint32_t* poss = idxs + (size_t)P.k * kKnnThreads;
    double bmin = INFINITY, bmax = -INFINITY;
    if (P.mode == KNN_NORMALS_FLAGGED) {
        // a small grid walks the positions normals_int_kernel left over (usually none)
        const uint32_t nwork = *P.todo_count;
        for (uint32_t w = blockIdx.x * kKnnThreads + threadIdx.x; w < nwork; w += gridDim.x * kKnnThreads)
            knn_point<K>(P, P.todo[w], d2s, idxs, poss, bmin, bmax);
        return;
    }
    const uint32_t t = P.begin + blockIdx.x * kKnnThreads + threadIdx.x;
    if (t < P.end) knn_point<K>(P, t, d2s, idxs, poss, bmin, bmax);
    if (P.mode == KNN_BOUNDARY) {
        __shared__ double sm[kKnnThreads / 32];
        double mn = block_min<kKnnThreads>(bmin, sm);
        double mx = block_max<kKnnThreads>(bmax, sm);
        if (threadIdx.x == 0) { P.minmax[2 * blockIdx.x] = mn; P.minmax[2 * blockIdx.x + 1] = mx; }
    }
}

// ------------------------------------------------------------------------------------
// K7': normal estimation for integer (voxelised) clouds by counting selection
//
// The generic k-NN kernel keeps a (d2, index)-sorted list of k = 30 entries per thread; on
// voxel surfaces it performs ~70 insertions shifting ~600 entries per point -- the list, not
// the search, is the cost.  Squared distances of integer clouds are small integers, so the
// 30-NN set can be selected by COUNTING instead:
//   pass 1  histogram of d2 (64 bins) over the candidates of rings 0, 1, 2, ... until the
//           k-th smallest distance T is certified (every point with d2 <= T lies in the
//           visited rings);
//   pass 2  re-walk the window d2 <= T: points with d2 < T go straight into integer cumulants
//           (sums of x, y, z, xx, xy, ... are exact, so their order is irrelevant); among the
//           points with d2 == T the (k - count(d2 < T)) smallest ORIGINAL indices are kept in
//           a short sorted list -- exactly the set the (d2, index) ordering selects.
// The cumulants are the same exact integers the list-based kernel feeds to the eigen-solver,
// so the normals are bit-identical to knn_self_kernel<KInt> (tested).  A point whose k-th
// neighbour is farther than sqrt(63) (sparse data) is flagged with NaN and finished by the
// generic kernel.
// ------------------------------------------------------------------------------------
constexpr int kHistBins = 64;
constexpr int kNrmThreads = 64;
constexpr int kTieCap = 16;           // ties kept at the k-th distance (more -> generic kernel)

template <class F>
__device__ __forceinline__ void int_window_walk(const RowGrid& g, const uint32_t* __restrict__ row_start,
                                                const uint4* __restrict__ recs, const KInt::Q& q, int qcy, int qcz,
                                                int cy0, int cz0, int r, uint32_t limit, F&& f) {
    const int ny = g.ny, nz = g.nz;
    auto pencil = [&](int yy, int zz) {
        if (yy < 0 || yy >= ny || zz < 0 || zz >= nz) return;
        const uint32_t by = yy > qcy ? KInt::sq(KInt::gap_up_y(g, q.y, yy)) : (yy < qcy ? KInt::sq(KInt::gap_dn_y(g, q.y, yy)) : 0u);
        const uint32_t bz = zz > qcz ? KInt::sq(KInt::gap_up_z(g, q.z, zz)) : (zz < qcz ? KInt::sq(KInt::gap_dn_z(g, q.z, zz)) : 0u);
        const uint32_t B2 = by + bz;
        if (B2 > limit) return;
        const uint32_t row = (uint32_t)zz * (uint32_t)ny + (uint32_t)yy;
        uint32_t lo = __ldg(row_start + row);
        const uint32_t hi = __ldg(row_start + row + 1);
        if (lo >= hi) return;
        const int w = (int)floorf(sqrtf((float)(limit - B2)));
        const int xlo = q.x - w, xhi = q.x + w;
        uint32_t b = hi;
        while (lo < b) {                       // first record with x >= xlo
            const uint32_t m = (lo + b) >> 1;
            if (KInt::rec_x(recs + m) < xlo) lo = m + 1; else b = m;
        }
        for (uint32_t i = lo; i < hi; ++i) {
            const uint4 rec = __ldg(recs + i);
            if ((int)(rec.x & 0xffffu) > xhi) break;
            const uint32_t d2 = KInt::dist2(q, rec);
            if (d2 <= limit) f(rec, i, d2);
        }
    };
    if (r == 0) { pencil(cy0, cz0); return; }
    for (int dy = -r; dy <= r; ++dy) { pencil(cy0 + dy, cz0 - r); pencil(cy0 + dy, cz0 + r); }
    for (int dz = -r + 1; dz <= r - 1; ++dz) { pencil(cy0 - r, cz0 + dz); pencil(cy0 + r, cz0 + dz); }
}

__global__ void __launch_bounds__(kNrmThreads)
normals_int_kernel(const __grid_constant__ KnnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = P.k;
    uint8_t* H = smem_raw + threadIdx.x;                                                   // H[bin * T]
    uint32_t* Tidx = reinterpret_cast<uint32_t*>(smem_raw + kHistBins * kNrmThreads) + threadIdx.x;   // [kTieCap][T]
    uint32_t* Tpos = Tidx + (size_t)kTieCap * kNrmThreads;
    const uint4* __restrict__ recs = static_cast<const uint4*>(P.c.recs);
    const RowGrid& g = P.c.grid;
    const uint32_t t = P.begin + blockIdx.x * kNrmThreads + threadIdx.x;
    if (t >= P.end) return;
    const uint4 qr = __ldg(recs + t);
    const KInt::Q q = KInt::rec_q(qr);
    const uint32_t qidx = qr.z;
    const int ny = g.ny, nz = g.nz;
    const int qcy = KInt::cell_y(g, q.y), qcz = KInt::cell_z(g, q.z);
    const int cy0 = qcy < 0 ? 0 : (qcy >= ny ? ny - 1 : qcy);
    const int cz0 = qcz < 0 ? 0 : (qcz >= nz ? nz - 1 : qcz);
    for (int b = 0; b < kHistBins; ++b) H[b * kNrmThreads] = 0;

    // ---- pass 1: histogram until the k-th distance is certified ----
    uint32_t limit = kHistBins - 1, total = 0;
    int rfin = -1;
    uint32_t T = 0, nless = 0;
    bool resolved = false;
    for (int r = 0;; ++r) {
        int_window_walk(g, P.c.row_start, recs, q, qcy, qcz, cy0, cz0, r, limit, [&](const uint4&, uint32_t, uint32_t d2) {
            const uint8_t c = H[d2 * kNrmThreads];
            if (c != 255) H[d2 * kNrmThreads] = c + 1;
            ++total;
        });
        // distance to everything not yet visited
        const int ylo = cy0 - r, yhi = cy0 + r, zlo = cz0 - r, zhi = cz0 + r;
        int m = 0x7fffffff;
        bool open = false;
        if (ylo > 0)      { open = true; m = min(m, KInt::gap_dn_y(g, q.y, ylo - 1)); }
        if (yhi < ny - 1) { open = true; m = min(m, KInt::gap_up_y(g, q.y, yhi + 1)); }
        if (zlo > 0)      { open = true; m = min(m, KInt::gap_dn_z(g, q.z, zlo - 1)); }
        if (zhi < nz - 1) { open = true; m = min(m, KInt::gap_up_z(g, q.z, zhi + 1)); }
        const uint32_t gap2 = !open ? 0xFFFFFFFFu : (m > 65535 ? 0xFFFFFFFFu : (uint32_t)(m * m));
        if (total >= (uint32_t)k) {
            uint32_t cum = 0;
            int b = 0;
            for (; b < kHistBins; ++b) { const uint32_t c = H[b * kNrmThreads]; if (cum + c >= (uint32_t)k) break; cum += c; }
            if (b < kHistBins) {
                T = (uint32_t)b; nless = cum;
                limit = T;                                    // later rings only need d2 <= T
                if (T < gap2) { resolved = true; rfin = r; break; }
            }
        }
        if (!open) { rfin = r; break; }                       // whole table visited: fewer than k points within range
        if (gap2 > (uint32_t)(kHistBins - 1)) { rfin = r; break; }   // farther candidates cannot fall into the bins
    }
    if (!resolved || k - (int)nless > kTieCap) {              // sparse neighbourhood, n < k, or a huge tie class:
        P.normals_out[3 * (size_t)qidx] = __longlong_as_double(0x7ff8000000000001ll);   // the generic kernel finishes it
        P.todo[atomicAdd(P.todo_count, 1u)] = t;
        return;
    }
    // ---- pass 2: cumulants of d2 < T, smallest indices among d2 == T ----
    const int mties = k - (int)nless;
    int nt = 0;
    unsigned long long sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    uint32_t sx = 0, sy = 0, sz = 0;
    auto add = [&](const uint4& rec) {
        const uint32_t x = rec.x & 0xffffu, y = rec.x >> 16, z = rec.y & 0xffffu;
        sx += x; sy += y; sz += z;
        sxx += (unsigned long long)(x * x); sxy += (unsigned long long)(x * y); sxz += (unsigned long long)(x * z);
        syy += (unsigned long long)(y * y); syz += (unsigned long long)(y * z); szz += (unsigned long long)(z * z);
    };
    for (int r = 0; r <= rfin; ++r) {
        int_window_walk(g, P.c.row_start, recs, q, qcy, qcz, cy0, cz0, r, T, [&](const uint4& rec, uint32_t pos, uint32_t d2) {
            if (d2 < T) { add(rec); return; }
            const uint32_t idx = rec.z;
            int j;
            if (nt < mties) j = nt++;
            else { if (idx >= Tidx[(mties - 1) * kNrmThreads]) return; j = mties - 1; }
            while (j > 0 && Tidx[(j - 1) * kNrmThreads] > idx) {
                Tidx[j * kNrmThreads] = Tidx[(j - 1) * kNrmThreads];
                Tpos[j * kNrmThreads] = Tpos[(j - 1) * kNrmThreads];
                --j;
            }
            Tidx[j * kNrmThreads] = idx;
            Tpos[j * kNrmThreads] = pos;
        });
    }
    for (int j = 0; j < nt; ++j) add(__ldg(recs + Tpos[j * kNrmThreads]));
    double cum[9] = {(double)sx, (double)sy, (double)sz, (double)sxx, (double)sxy, (double)sxz, (double)syy, (double)syz, (double)szz};
    double nv[3];
    normal_from_cumulants(cum, k, nv);
    P.normals_out[3 * (size_t)qidx + 0] = nv[0];
    P.normals_out[3 * (size_t)qidx + 1] = nv[1];
    P.normals_out[3 * (size_t)qidx + 2] = nv[2];
}

__global__ void minmax_finalize_kernel(const double* in, uint32_t nb, double* out) {
    double mn = INFINITY, mx = -INFINITY;
    for (uint32_t i = threadIdx.x; i < nb; i += 256) { mn = fmin(mn, in[2 * i]); mx = fmax(mx, in[2 * i + 1]); }
    __shared__ double sm[8];
    double a = block_min<256>(mn, sm);
    double b = block_max<256>(mx, sm);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

// ------------------------------------------------------------------------------------
// minimal oriented bounding box sweep (the PSNR peak of the reference: cloud_pair.py:111-112 ->
// Open3D OrientedBoundingBox::CreateFromPointsMinimal).  The convex hull comes from Qhull on the
// host (as in Open3D); for EVERY hull triangle the hull vertices are expressed in the triangle's
// frame (u = b - a, w = u x (c - a), v = w x u, normalised) and the axis-aligned extent / volume
// of that box is reduced here: one block per triangle, F x V point rotations in total.
// ------------------------------------------------------------------------------------
constexpr int kObbThreads = 128;
__global__ void __launch_bounds__(kObbThreads)
obb_sweep_kernel(const double* __restrict__ verts, uint32_t nv, const double* __restrict__ tris, uint32_t nf,
                 double* __restrict__ vol_out, double* __restrict__ ext_out) {
    const uint32_t f = blockIdx.x;
    if (f >= nf) return;
    const double* t = tris + 9 * (size_t)f;
    const double a[3] = {t[0], t[1], t[2]};
    double u[3] = {dsub(t[3], a[0]), dsub(t[4], a[1]), dsub(t[5], a[2])};
    double v[3] = {dsub(t[6], a[0]), dsub(t[7], a[1]), dsub(t[8], a[2])};
    double w[3];
    cross3(u, v, w);
    cross3(w, u, v);
    const double lu = sqrt(dot3(u, u)), lv = sqrt(dot3(v, v)), lw = sqrt(dot3(w, w));
    for (int k = 0; k < 3; ++k) { u[k] /= lu; v[k] /= lv; w[k] /= lw; }
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = threadIdx.x; i < nv; i += kObbThreads) {
        const double p[3] = {dsub(__ldg(verts + 3 * (size_t)i), a[0]), dsub(__ldg(verts + 3 * (size_t)i + 1), a[1]),
                             dsub(__ldg(verts + 3 * (size_t)i + 2), a[2])};
        const double l[3] = {dot3(u, p), dot3(v, p), dot3(w, p)};   // R^T (p - a): the frame is orthonormal
        for (int k = 0; k < 3; ++k) { mn[k] = fmin(mn[k], l[k]); mx[k] = fmax(mx[k], l[k]); }
    }
    __shared__ double sm[kObbThreads / 32];
    double ext[3];
    for (int k = 0; k < 3; ++k) {
        const double lo = block_min<kObbThreads>(mn[k], sm);
        const double hi = block_max<kObbThreads>(mx[k], sm);
        ext[k] = dsub(hi, lo);
    }
    if (threadIdx.x == 0) {
        ext_out[3 * (size_t)f] = ext[0]; ext_out[3 * (size_t)f + 1] = ext[1]; ext_out[3 * (size_t)f + 2] = ext[2];
        vol_out[f] = dmul(dmul(ext[0], ext[1]), ext[2]);
    }
}

// ------------------------------------------------------------------------------------
// convex-hull prefilter for the minimal-OBB peak
//   extremes_kernel : arg-max of the projection on each of D directions (seed points);
//   outside_kernel  : keep only the points that are NOT strictly inside the seed polytope
//                     (n_f . p + d_f < -eps for every facet f) -- exact: a point strictly inside
//                     a convex subset of the hull cannot be a hull vertex.
// Qhull then runs on a few percent of the cloud.
// ------------------------------------------------------------------------------------
constexpr int kExtThreads = 128;
constexpr int kExtDirs = 16;          // directions per block (registers)
constexpr int kExtChunk = 16384;      // points per block

template <class K>
__global__ void __launch_bounds__(kExtThreads)
extremes_kernel(const typename K::Rec* __restrict__ recs, uint32_t first, uint32_t n, const double* __restrict__ dirs, int ndirs,
                double* __restrict__ pval, uint32_t* __restrict__ pidx) {
    const int d0 = blockIdx.y * kExtDirs;
    const uint32_t c0 = blockIdx.x * kExtChunk;
    double dx[kExtDirs], dy[kExtDirs], dz[kExtDirs], best[kExtDirs];
    uint32_t bi[kExtDirs];
#pragma unroll
    for (int j = 0; j < kExtDirs; ++j) {
        const int d = d0 + j < ndirs ? d0 + j : ndirs - 1;
        dx[j] = dirs[3 * d]; dy[j] = dirs[3 * d + 1]; dz[j] = dirs[3 * d + 2];
        best[j] = -INFINITY; bi[j] = 0xFFFFFFFFu;
    }
    const uint32_t end = c0 + kExtChunk < n ? c0 + kExtChunk : n;
    for (uint32_t i = c0 + threadIdx.x; i < end; i += kExtThreads) {
        const typename K::Rec r = load_rec(recs + first + i);
        const typename K::Q q = K::rec_q(r);
        const double x = (double)q.x, y = (double)q.y, z = (double)q.z;
        const uint32_t idx = K::rec_idx(r);
#pragma unroll
        for (int j = 0; j < kExtDirs; ++j) {
            const double v = dx[j] * x + dy[j] * y + dz[j] * z;
            if (v > best[j] || (v == best[j] && idx < bi[j])) { best[j] = v; bi[j] = idx; }
        }
    }
    __shared__ double sv[kExtThreads / 32][kExtDirs];
    __shared__ uint32_t si[kExtThreads / 32][kExtDirs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kExtDirs; ++j) {
        double v = best[j];
        uint32_t id = bi[j];
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, v, o);
            const uint32_t oi = __shfl_down_sync(0xffffffffu, id, o);
            if (ov > v || (ov == v && oi < id)) { v = ov; id = oi; }
        }
        if (lane == 0) { sv[warp][j] = v; si[warp][j] = id; }
    }
    __syncthreads();
    if (threadIdx.x < kExtDirs && d0 + (int)threadIdx.x < ndirs) {
        double v = sv[0][threadIdx.x];
        uint32_t id = si[0][threadIdx.x];
        for (int w = 1; w < kExtThreads / 32; ++w)
            if (sv[w][threadIdx.x] > v || (sv[w][threadIdx.x] == v && si[w][threadIdx.x] < id)) { v = sv[w][threadIdx.x]; id = si[w][threadIdx.x]; }
        pval[(size_t)blockIdx.x * ndirs + d0 + threadIdx.x] = v;
        pidx[(size_t)blockIdx.x * ndirs + d0 + threadIdx.x] = id;
    }
}

template <class K>
__global__ void outside_kernel(const typename K::Rec* __restrict__ recs, uint32_t first, uint32_t n,
                               const double* __restrict__ planes, int nf, double eps,
                               double* __restrict__ out_xyz, unsigned long long capacity, unsigned long long* counter) {
    extern __shared__ double sp[];   // planes
    for (int i = threadIdx.x; i < nf * 4; i += blockDim.x) sp[i] = planes[i];
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const typename K::Q q = K::rec_q(load_rec(recs + first + i));
    const double x = (double)q.x, y = (double)q.y, z = (double)q.z;
    bool inside = true;
    for (int f = 0; f < nf && inside; ++f)
        inside = sp[4 * f] * x + sp[4 * f + 1] * y + sp[4 * f + 2] * z + sp[4 * f + 3] < -eps;
    if (inside) return;
    const unsigned long long slot = atomicAdd(counter, 1ull);
    if (slot < capacity) { out_xyz[3 * slot] = x; out_xyz[3 * slot + 1] = y; out_xyz[3 * slot + 2] = z; }
}

}  // namespace pccm
