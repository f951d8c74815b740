// pccm_api.cu -- host orchestration + C ABI (include/pccm.h) of libpccm.so.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -shared ...
// There is no CPU path: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "pccm_kernels.cuh"
#include "pccm_vox_kernels.cuh"

using namespace pccm;

// --------------------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------------------
struct PendingTimer {
    cudaEvent_t a, b;
    double* dest;
};

struct pccm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;   // host->device copies of attributes that are only needed by the query epilogue
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_sync = nullptr;         // host waits that must not cover work enqueued after them
    // per-cloud events and pinned colour-flag words are recycled: creating them costs tens of microseconds
    // and a cloud lives for one evaluation
    std::vector<cudaEvent_t> cloud_events;
    std::vector<int> flag_slots;          // free 4-byte words of ctx->pinned (kFlagPinnedOffset)
    std::string err;
    int profiling = 0;
    pccm_timings tm{};
    std::vector<PendingTimer> pending;
    std::vector<cudaEvent_t> event_pool;
    // per-point outputs of the last pair_eval
    int32_t* pp_idx[2] = {nullptr, nullptr};
    double* pp_d2[2] = {nullptr, nullptr};
    int64_t pp_n[2] = {0, 0};
    const struct pccm_cloud* pp_owner[2] = {nullptr, nullptr};   // the pair those outputs belong to (dropped when either cloud goes)
    // small pinned + device scratch for result structs
    void* pinned = nullptr;
    void* dscratch = nullptr;
    static constexpr size_t kScratch = 1 << 16;
    static constexpr size_t kPinned = 3 << 16;      // results, flags, plan read-back + two statistics slots
    // device blocks of this context, recycled (see dalloc)
    std::multimap<size_t, void*> dev_free;            // released blocks by (bucketed) size
    std::unordered_map<void*, size_t> dev_live;       // blocks handed out -> their bucketed size
    size_t dev_cached = 0, dev_cache_cap = 0;         // bytes in dev_free / the most it may hold
    bool dev_cache = true;                            // PCCM_DEV_CACHE=0: every block straight from / to the driver's pool
    int sm_count = 148;
    int cell_override_shift = -1;   // debugging: PCCM_CELL_SHIFT
    double cell_scale = 1.0;        // debugging: PCCM_CELL_SCALE
    uint32_t short_row = 0;         // rows up to this length skip the binary search (PCCM_SHORT_ROW; measured: never a win)
    bool normals_counting = true;   // KInt normals by counting selection (PCCM_NORMALS_COUNTING=0: list-based kernel only)
    bool use_vox = true;            // KInt pairs: occupancy-brick index + bit-scan query (PCCM_VOX=0: pencil path only)
    bool eager_pencil = false;      // build the pencil index of brick-indexed pairs at once instead of on first use (PCCM_EAGER_PENCIL=1)
    int vx_search_blocks = 10;      // resident blocks per SM of the persistent brick search kernel (PCCM_VX_BLOCKS)
    struct HostNarrow* narrow = nullptr;   // host threads + pinned ring for narrowing float64 host arrays (created on first use)
    bool host_narrow = false;              // PCCM_HOST_NARROW=1: narrow float64 host arrays on the way up.  Off by default: measured on
                                           // the B200 boxes (PCIe 5 x16, 54 GB/s from pinned memory) eight host threads pack 24 MB
                                           // in 0.5 ms -- the time the link needs for the unpacked array; it pays on slower links only
    int host_threads = 0;                  // PCCM_HOST_THREADS (0 = min(hardware threads, 8))
    int shard_rank = 0, shard_world = 1;   // pccm_ctx_set_shard: pairs built from now on are split by z slabs over `world` ranks
    bool shard_sel = true;          // split pairs: fill / place walk a compacted list of the slab's points (PCCM_SHARD_SEL=0: every point)
    int vxyz_rows = -1;             // voxel coordinates written from the occupancy rows, in rank order (1), or by every point, scattered (0);
                                    // -1: by the number of points indexed here (PCCM_VXYZ_ROWS)
    bool stats_tma = true;          // statistics pass: float64 rows staged by cp.async.bulk + mbarrier (PCCM_STATS_TMA=0: plain loads)
    bool epi_compact = true;        // epilogue of a split pair / voxel slice works off a per-block queue of its own points (PCCM_EPI_COMPACT=0: every tile)
    bool epi_tma = true;            // epilogue streams staged by cp.async.bulk + mbarrier (PCCM_EPI_TMA=0: plain loads)
    bool pdl = true;                // programmatic dependent launch along the kernel chains of an evaluation (PCCM_PDL=0: ordinary launches)
    int mark_sample = 32;           // the brick directory is marked by 1 / mark_sample of the points first (PCCM_MARK_SAMPLE, 0 = one pass)
};

static thread_local std::string g_err;

static int fail(pccm_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_err = buf;
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, PCCM_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

// Launch of a kernel that begins with pdl_enter(): with the programmatic-serialisation attribute its blocks may become
// resident while the previous kernel of the stream is still draining (pccm_kernels.cuh, pdl_enter).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(pccm_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = ctx->pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// zero `bytes` (a multiple of 4, 4-byte aligned) on `st` with a kernel of the chain (see zero_words_kernel)
static inline cudaError_t dzero(pccm_ctx* ctx, void* p, size_t bytes, cudaStream_t st) {
    const size_t nwords = (bytes + 3) / 4;
    if (nwords == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((nwords + 1023) / 1024, (size_t)ctx->sm_count * 8);
    ctx->tm.total_launches++;
    return launch_chain(ctx, zero_words_kernel, blocks, 256, 0, st, static_cast<uint32_t*>(p), nwords);
}

struct StageTimer {
    pccm_ctx* ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    double* dest;
    bool on;
    cudaStream_t st;
    StageTimer(pccm_ctx* c, double* d, int level = 2, cudaStream_t s = nullptr) : ctx(c), dest(d), on(c->profiling >= level), st(s ? s : c->stream) {
        if (!on) return;
        a = get();
        b = get();
        cudaEventRecord(a, st);
    }
    ~StageTimer() {
        if (!on) return;
        cudaEventRecord(b, st);
        ctx->pending.push_back({a, b, dest});
    }
    cudaEvent_t get() {
        if (!ctx->event_pool.empty()) {
            cudaEvent_t e = ctx->event_pool.back();
            ctx->event_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

static void resolve_timers(pccm_ctx* ctx) {
    for (auto& p : ctx->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) *p.dest += ms;
        ctx->event_pool.push_back(p.a);
        ctx->event_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

// Device memory of a context is stream-ordered on ctx->stream (a block may be used by work enqueued after dalloc and must
// not be used by work enqueued after dfree) and RECYCLED inside the context: released blocks wait in a size-bucketed
// list and the next request of that bucket takes one -- a cloud pair allocates the same dozen sizes evaluation after
// evaluation.  The driver's own pool (cudaMallocAsync / cudaFreeAsync) gives the same semantics, but measured end to
// end it stalls the host for tens to hundreds of milliseconds now and then when an allocation meets frees of the
// previous evaluation that are still in flight (tools/e2e_bisect.py: steady 3.1 ms per pair with a device
// synchronisation between evaluations, p90 20 ms / max 490 ms without); it still backs the cache.
static size_t dev_bucket(size_t bytes) {
    if (bytes < 512) return 512;
    size_t g = 512;                                   // granularity: 1/8 of the largest power of two below the size
    while ((g << 4) <= bytes) g <<= 1;
    return (bytes + g - 1) / g * g;
}
static void dev_cache_trim(pccm_ctx* ctx, size_t keep) {
    while (ctx->dev_cached > keep && !ctx->dev_free.empty()) {
        auto it = std::prev(ctx->dev_free.end());     // largest first
        cudaFreeAsync(it->second, ctx->stream);
        ctx->dev_cached -= it->first;
        ctx->dev_free.erase(it);
    }
}
static cudaError_t dalloc_bytes(pccm_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    if (!ctx->dev_cache) return cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
    const size_t b = dev_bucket(bytes);
    auto it = ctx->dev_free.find(b);
    if (it != ctx->dev_free.end()) {
        *p = it->second;
        ctx->dev_cached -= b;
        ctx->dev_free.erase(it);
    } else {
        cudaError_t e = cudaMallocAsync(p, b, ctx->stream);
        if (e != cudaSuccess) {                       // out of memory: give the cached blocks back and try once more
            cudaGetLastError();
            dev_cache_trim(ctx, 0);
            cudaStreamSynchronize(ctx->stream);
            e = cudaMallocAsync(p, b, ctx->stream);
            if (e != cudaSuccess) { *p = nullptr; return e; }
        }
    }
    ctx->dev_live[*p] = b;
    return cudaSuccess;
}
template <class T>
static cudaError_t dalloc(pccm_ctx* ctx, T** p, size_t count) {
    return dalloc_bytes(ctx, reinterpret_cast<void**>(p), count * sizeof(T));
}
static void dfree(pccm_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->dev_live.find(p);
    if (it == ctx->dev_live.end()) { cudaFreeAsync(p, ctx->stream); return; }     // (not from the cache: PCCM_DEV_CACHE=0)
    const size_t b = it->second;
    ctx->dev_live.erase(it);
    ctx->dev_free.emplace(b, p);
    ctx->dev_cached += b;
    if (ctx->dev_cached > ctx->dev_cache_cap) dev_cache_trim(ctx, ctx->dev_cache_cap / 2);
}

// --------------------------------------------------------------------------------------
// cloud
// --------------------------------------------------------------------------------------
struct pccm_cloud {
    int64_t n = 0;
    // raw input, alive until the index is built
    const void* raw_xyz = nullptr;
    void* raw_owned = nullptr;
    int raw_dtype = PCCM_F64;
    int64_t raw_stride = 0;
    // raw colours until classified
    const void* raw_rgb = nullptr;
    void* raw_rgb_owned = nullptr;
    int raw_rgb_dtype = PCCM_F64;
    int64_t raw_rgb_stride = 0;
    // statistics
    StatsPartial* d_stats = nullptr;  // one device block: [stats_blocks] partials, [1] their fold (what the host reads, on demand),
                                      // the colour flag, the DevStats record and the packed coordinates
    DevStats* d_dev = nullptr;        // the same statistics as one integer record (brick-index planning on the device)
    uint2* packed = nullptr;          // {x | y << 16, z} per point, written by the statistics pass (integer-valued clouds)
    uint32_t* d_zhist = nullptr;      // sharded contexts: points per 8-voxel z layer (kZHistBins), in the same block
    bool stats_fetched = false;       // the statistics partials are (being) copied to the pinned slot stats_slot
    int stats_slot = 0;
    int stats_blocks = 0;
    bool stats_ready = false;
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    int data_kind = PCCM_KIND_F64;
    bool rgb_u8_ok = false;
    // colours travel on the copy stream (upload + the k/255 classification) beside the index build
    cudaEvent_t rgb_ready = nullptr;
    bool rgb_pending = false;
    uint32_t* d_rgbflag = nullptr;   // behind d_stats
    uint32_t* h_rgbflag = nullptr;   // pinned word (ctx->flag_slots)
    int flag_slot = -1;
    bool rgb_unclassified = false;   // float64 colours on the device that nobody has classified yet (done on first need)
    bool vox_rgb_done = false;       // brick index: colours packed into their original-order array
    bool rgb_spec = false;           // rgb_u8 was packed before the k/255 classification was known (d_rgbflag is read at the next settle point)
    bool vox_rgb_in_recs = false;    // brick index: the voxel records carry the representative's colour
    // attributes (original order)
    uchar4* rgb_u8 = nullptr;
    double* rgb_f64 = nullptr;
    double* normals = nullptr;
    bool normals_borrowed = false;   // caller's packed float64 device array (pccm_cloud_create, DEVICE)
    cudaEvent_t nrm_ready = nullptr; // normals still in flight on the copy stream
    bool nrm_pending = false;
    void* nrm_stage = nullptr;       // raw float32 / strided rows being converted on the copy stream
    bool has_colors = false, has_normals = false;
    // index
    int index_kind = -1;
    RowGrid grid{};
    void* recs = nullptr;          // record array (own, or the pair's joint array)
    uint32_t* row_start = nullptr; // pencil table; values are positions in `recs`
    uint32_t base = 0;             // position of this cloud's first record in `recs`
    struct SharedIndex* shared = nullptr;   // joint build: buffers owned by both clouds
    struct SharedVox* vox = nullptr;        // occupancy-brick index of the pair this cloud was built with (KInt)
    int vox_id = 0;                         // which of the pair's two views is this cloud
    bool rgb_in_rec = false;       // KInt records carry the 8-bit colour
    double cell_size = 0;
};

struct SharedIndex {
    void* recs = nullptr;
    uint32_t* table = nullptr;
    int refs = 0;
};

struct SharedVox {
    uint32_t *dirbits = nullptr, *dirpre = nullptr, *dirsums = nullptr, *bricksums = nullptr, *prank = nullptr;
    uint32_t *bcursor = nullptr, *border = nullptr;      // bricks by size class (vx_rowbase_kernel -> vx_search_kernel)
    uint2 *rows = nullptr, *vxyz = nullptr, *vkey = nullptr;
    unsigned char* arena = nullptr;  // ONE device allocation for all of the above (+ the plan)
    VoxPlan* dplan = nullptr;        // device: written by the build kernels, read by every brick kernel
    VoxPlan hplan{};                 // host copy, valid once the build has been settled
    bool pending = false;            // the build is enqueued but the host has not looked at its outcome yet
    bool sharded = false;            // built on a sharded context: only this rank's slab (+ halo) is indexed, only its layers are queried
    bool full_need = false;          // ... but the whole pair is indexed (fallback: some query had to look beyond the halo)
    struct HostNarrow* narrow = nullptr;   // host threads + pinned ring for narrowing float64 host arrays (created on first use)
    bool host_narrow = false;              // PCCM_HOST_NARROW=1: narrow float64 host arrays on the way up.  Off by default: measured on
                                           // the B200 boxes (PCIe 5 x16, 54 GB/s from pinned memory) eight host threads pack 24 MB
                                           // in 0.5 ms -- the time the link needs for the unpacked array; it pays on slower links only
    int host_threads = 0;                  // PCCM_HOST_THREADS (0 = min(hardware threads, 8))
    int shard_rank = 0, shard_world = 1;
    ShardPlan* dshard = nullptr;     // device: the cuts (in the arena)
    uint32_t cap_dirw = 0, cap_blk = 0;
    uint32_t n[2] = {0, 0};
    VoxView view[2];                 // = hplan.view
    struct pccm_cloud* owner[2] = {nullptr, nullptr};   // live clouds of the pair (the pencil index is built from vxyz on demand)
    double cell_size = 0;                                // what pccm_pair_build_index was asked for
    int force_kind = PCCM_KIND_AUTO;
    int refs = 0;
};

static void free_vox(pccm_ctx* ctx, SharedVox* v) {
    dfree(ctx, v->arena);          // every array of the index is a slice of this one block
    delete v;
}
static void release_vox(pccm_ctx* ctx, pccm_cloud* c) {
    SharedVox* v = c->vox;
    c->vox = nullptr;
    if (!v) return;
    v->owner[c->vox_id] = nullptr;
    if (--v->refs == 0) free_vox(ctx, v);
}

static constexpr int64_t kNarrowChunk = 1 << 17;       // host-side narrowing: rows per chunk
static int upload_narrowed(pccm_ctx* ctx, const double* src, int64_t n, int es, cudaStream_t s, unsigned char* d, bool* ok);
static int build_rowsort_kind(pccm_ctx* ctx, int kind, pccm_cloud* cl[2], const PairRaw& R);
static int vox_settle(pccm_ctx* ctx, pccm_cloud* c);
static int vox_fetch(pccm_ctx* ctx, SharedVox* v);
static int vox_adopt(pccm_ctx* ctx, SharedVox* v, bool* redo);

static size_t dtype_size(int dt) {
    switch (dt) {
        case PCCM_F64: return 8;
        case PCCM_F32: return 4;
        case PCCM_I32: return 4;
        case PCCM_U16: return 2;
        case PCCM_U8: return 1;
    }
    return 0;
}

static constexpr size_t kStatsPinnedOffset = 65536;   // bytes into ctx->pinned: two slots for the per-block statistics partials of a cloud
static constexpr size_t kStatsSlotBytes = 65536;
static constexpr size_t kFlagPinnedOffset = 20480;    // ... : 1024 colour-flag words

// the statistics partials of a cloud travel to the host only when host code needs them
static int stats_fetch(pccm_ctx* ctx, pccm_cloud* c, int slot) {
    if (c->stats_ready || c->n == 0) return PCCM_OK;
    char* h = static_cast<char*>(ctx->pinned) + kStatsPinnedOffset + (size_t)slot * kStatsSlotBytes;
    CK(cudaMemcpyAsync(h, c->d_stats, sizeof(StatsPartial) * (size_t)c->stats_blocks, cudaMemcpyDeviceToHost, ctx->stream));
    c->stats_fetched = true;
    c->stats_slot = slot;
    return PCCM_OK;
}
static int stats_adopt(pccm_ctx* ctx, pccm_cloud* c) {      // after the stream has been synchronised behind stats_fetch
    if (c->stats_ready) return PCCM_OK;
    if (c->n == 0) { c->data_kind = PCCM_KIND_INT; c->stats_ready = true; return PCCM_OK; }
    const StatsPartial* h = reinterpret_cast<const StatsPartial*>(static_cast<const char*>(ctx->pinned) + kStatsPinnedOffset + (size_t)c->stats_slot * kStatsSlotBytes);
    c->stats_fetched = false;
    bool not_int = false, not_f32 = false, not_fin = false;
    for (int a = 0; a < 3; ++a) { c->mn[a] = INFINITY; c->mx[a] = -INFINITY; }
    for (int b = 0; b < c->stats_blocks; ++b) {
        const StatsPartial& p = h[b];
        for (int a = 0; a < 3; ++a) { c->mn[a] = std::min(c->mn[a], p.mn[a]); c->mx[a] = std::max(c->mx[a], p.mx[a]); }
        not_int |= p.not_int != 0; not_f32 |= p.not_f32 != 0; not_fin |= p.not_finite != 0;
    }
    if (not_fin) return fail(ctx, PCCM_ERR_NONFINITE, "cloud has NaN or Inf coordinates");
    c->data_kind = !not_int ? PCCM_KIND_INT : (!not_f32 ? PCCM_KIND_F32 : PCCM_KIND_F64);
    c->stats_ready = true;
    return PCCM_OK;
}
static int ensure_stats(pccm_ctx* ctx, pccm_cloud* c) {
    if (c->stats_ready) return PCCM_OK;
    const int rc = stats_fetch(ctx, c, 0);
    if (rc) return rc;
    if (c->n) CK(cudaStreamSynchronize(ctx->stream));
    return stats_adopt(ctx, c);
}

// order the context stream after an in-flight normal upload (no-op when there is none)
static void wait_normals(pccm_ctx* ctx, pccm_cloud* c) {
    if (c->nrm_pending) {
        cudaStreamWaitEvent(ctx->stream, c->nrm_ready, 0);
        c->nrm_pending = false;
    }
    if (c->nrm_stage) {              // (after the wait: the free is ordered behind the conversion)
        dfree(ctx, c->nrm_stage);
        c->nrm_stage = nullptr;
    }
}

// colours in flight on the copy stream: wait (host: the classification decides the device format;
// context stream: later kernels read the uploaded rows)
static int ensure_colors(pccm_ctx* ctx, pccm_cloud* c) {
    if (!c->has_colors) return PCCM_OK;
    if (c->rgb_unclassified) {       // float64 colours handed over on the device: classify now
            uint32_t* h = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->pinned) + kFlagPinnedOffset + 4096);
        CK(cudaMemsetAsync(c->d_rgbflag, 0, sizeof(uint32_t), ctx->stream));
        rgb_classify_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(c->raw_rgb, c->raw_rgb_stride, c->n, c->d_rgbflag);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h, c->d_rgbflag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        c->rgb_u8_ok = *h == 0u;
        c->rgb_unclassified = false;
        return PCCM_OK;
    }
    if (!c->rgb_pending) return PCCM_OK;
    CK(cudaEventSynchronize(c->rgb_ready));
    CK(cudaStreamWaitEvent(ctx->stream, c->rgb_ready, 0));
    c->rgb_u8_ok = c->raw_rgb_dtype == PCCM_U8 || c->n == 0 || (c->h_rgbflag && *c->h_rgbflag == 0u);
    c->rgb_pending = false;
    return PCCM_OK;
}

// Colours of a cloud: HOST arrays are uploaded and classified (every channel == k / 255 ?) on the copy stream,
// so that the statistics pass, the index build and the search never wait for them; DEVICE arrays are used where
// they are (8-bit ones need nothing at all, float64 ones are classified when somebody needs to know).
static int attach_colors(pccm_ctx* ctx, pccm_cloud* c, const void* rgb, int dtype, int64_t stride, int mem_kind) {
    if (dtype != PCCM_F64 && dtype != PCCM_U8) return fail(ctx, PCCM_ERR_INVALID, "rgb must be F64 or U8");
    if (c->has_colors) return fail(ctx, PCCM_ERR_STATE, "cloud already has colours");
    if (c->index_kind >= 0 && !(c->vox && !c->recs && !c->vox_rgb_done))
        return fail(ctx, PCCM_ERR_STATE, "colours must be attached before the index is built");
    const size_t es = dtype_size(dtype);
    if (stride == 0) stride = (int64_t)(3 * es);
    if (stride < (int64_t)(3 * es) || (stride % (int64_t)es) != 0) return fail(ctx, PCCM_ERR_INVALID, "bad row stride %lld", (long long)stride);
    c->raw_rgb_dtype = dtype;
    c->raw_rgb_stride = stride;
    c->has_colors = true;
    if (c->n == 0) { c->rgb_u8_ok = true; return PCCM_OK; }
    if (mem_kind == PCCM_DEVICE) {
        c->raw_rgb = rgb;
        c->rgb_u8_ok = dtype == PCCM_U8;
        c->rgb_unclassified = dtype == PCCM_F64;
        return PCCM_OK;
    }
    cudaStream_t s = ctx->copy_stream ? ctx->copy_stream : ctx->stream;
    if (!c->rgb_ready) {
        if (!ctx->cloud_events.empty()) { c->rgb_ready = ctx->cloud_events.back(); ctx->cloud_events.pop_back(); }
        else CK(cudaEventCreateWithFlags(&c->rgb_ready, cudaEventDisableTiming));
    }
    if (dtype == PCCM_F64 && stride == 24 && c->n >= kNarrowChunk && ctx->host_narrow) {
        // 8-bit colours stored as k / 255.0: packed to 3 bytes per row by the host threads (which is also their classification)
        unsigned char* d8 = nullptr;
        CK(dalloc(ctx, &d8, (size_t)c->n * 3 + 16));
        if (s != ctx->stream) {
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CK(cudaStreamWaitEvent(s, ctx->ev_fork, 0));
        }
        bool narrowed = false;
        const int rcn = upload_narrowed(ctx, static_cast<const double*>(rgb), c->n, 1, s, d8, &narrowed);
        if (rcn) { dfree(ctx, d8); return rcn; }
        if (narrowed) {
            c->raw_rgb = d8; c->raw_rgb_owned = d8;
            c->raw_rgb_dtype = PCCM_U8; c->raw_rgb_stride = 3;
            CK(cudaEventRecord(c->rgb_ready, s));
            c->rgb_pending = true;
            return PCCM_OK;
        }
        CK(cudaEventRecord(ctx->ev_fork, s));                 // (the abandoned chunk copies still target d8: free it behind them)
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_fork, 0));
        dfree(ctx, d8);
    }
    if (dtype == PCCM_F64 && c->flag_slot < 0) {
        if (ctx->flag_slots.empty()) return fail(ctx, PCCM_ERR_STATE, "too many live clouds with colours in flight");
        c->flag_slot = ctx->flag_slots.back();
        ctx->flag_slots.pop_back();
        c->h_rgbflag = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->pinned) + kFlagPinnedOffset) + c->flag_slot;
    }
    unsigned char* d = nullptr;
    const size_t bytes = (size_t)c->n * (size_t)stride;
    CK(dalloc(ctx, &d, bytes));
    if (s != ctx->stream) {      // after the allocation above
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(s, ctx->ev_fork, 0));
    }
    CK(cudaMemcpyAsync(d, rgb, bytes, cudaMemcpyHostToDevice, s));
    c->raw_rgb = d;
    c->raw_rgb_owned = d;
    if (dtype == PCCM_F64) {
        CK(cudaMemsetAsync(c->d_rgbflag, 0, sizeof(uint32_t), s));
        rgb_classify_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(c->raw_rgb, stride, c->n, c->d_rgbflag);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c->h_rgbflag, c->d_rgbflag, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(c->rgb_ready, s));
    c->rgb_pending = true;
    return PCCM_OK;
}

static void launch_pack_u8(pccm_ctx* ctx, const pccm_cloud* c, uchar4* out) {
    const int threads = 256;
    if (c->raw_rgb_dtype == PCCM_U8 && c->raw_rgb_stride == 3 && (reinterpret_cast<uintptr_t>(c->raw_rgb) & 3u) == 0) {
        const int64_t groups = (c->n + 3) / 4;
        launch_chain(ctx, pack_rgb_u8x4_kernel, (int)((groups + threads - 1) / threads), threads, 0, ctx->stream, static_cast<const uint32_t*>(c->raw_rgb), c->n, out);
    } else {
        launch_chain(ctx, pack_rgb_u8_kernel, (int)((c->n + threads - 1) / threads), threads, 0, ctx->stream, c->raw_rgb, c->raw_rgb_dtype, c->raw_rgb_stride, c->n, out);
    }
}

// colours: uchar4 when every channel is k/255, else packed float64
static int finish_colors(pccm_ctx* ctx, pccm_cloud* c) {
    { const int rc0 = ensure_colors(ctx, c); if (rc0) return rc0; }
    if (!c->has_colors || c->rgb_u8 || c->rgb_f64 || c->n == 0) return PCCM_OK;
    const int threads = 256;
    const int blocks = (int)((c->n + threads - 1) / threads);
    if (c->raw_rgb_dtype == PCCM_U8 || c->rgb_u8_ok) {
        CK(dalloc(ctx, &c->rgb_u8, (size_t)c->n + 4));
        launch_pack_u8(ctx, c, c->rgb_u8);
    } else {
        CK(dalloc(ctx, &c->rgb_f64, (size_t)c->n * 3));
        pack_f64x3_kernel<<<blocks, threads, 0, ctx->stream>>>(c->raw_rgb, PCCM_F64, c->raw_rgb_stride, c->n, c->rgb_f64);
    }
    ctx->tm.total_launches++;
    CK(cudaGetLastError());
    dfree(ctx, c->raw_rgb_owned);
    c->raw_rgb_owned = nullptr;
    c->raw_rgb = nullptr;
    return PCCM_OK;
}

// --------------------------------------------------------------------------------------
// host-side narrowing of float64 HOST arrays (the form Open3D / numpy hand over)
//
// Voxelised coordinates are integers and 8-bit colours are k / 255: 24 bytes per row where 6 (uint16 x 3) or 3 (uchar x 3)
// say the same.  The end-to-end path is bound by the host->device copy, so the rows are narrowed WHILE CHECKING that
// nothing is lost -- by a few host threads, chunk by chunk, into a ring of pinned staging slots -- and every finished
// chunk's copy is enqueued at once (it overlaps the packing of the next chunks).  A value that does not fit stops the
// attempt and the array goes up unchanged.  Results are identical by construction: the device would have made the same
// conversion (stats_kernel / the colour classification) after the copy.
// --------------------------------------------------------------------------------------
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv, done_cv;
    std::function<void(int64_t)> job;
    int64_t next = 0, count = 0, finished = 0;
    uint64_t epoch = 0;
    bool stop = false;
    explicit HostPool(int nthreads) {
        for (int t = 0; t < nthreads; ++t)
            threads.emplace_back([this] {
                uint64_t seen = 0;
                for (;;) {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return stop || (epoch != seen && next < count); });
                    if (stop) return;
                    const uint64_t e = epoch;
                    while (next < count && epoch == e) {
                        const int64_t i = next++;
                        lk.unlock();
                        job(i);
                        lk.lock();
                        if (++finished == count) done_cv.notify_all();
                    }
                    seen = e;
                }
            });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
    // run fn(0..n-1) on the pool; the caller may poll `ready` flags the job sets -- returns after all have finished
    void start(int64_t n, std::function<void(int64_t)> fn) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(fn); next = 0; count = n; finished = 0; ++epoch;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        done_cv.wait(lk, [&] { return finished == count; });
    }
};

static constexpr int kNarrowSlots = 16;                // pinned staging ring (6 bytes per row and slot at most)

struct HostNarrow {
    HostPool* pool = nullptr;
    unsigned char* ring = nullptr;                     // pinned: kNarrowSlots x kNarrowChunk x 6 bytes
    cudaEvent_t slot_free[kNarrowSlots] = {};
    bool slot_used[kNarrowSlots] = {};
};

static bool narrow_chunk_u16_scalar(const double* src, int64_t n, uint16_t* dst) {
    bool ok = true;
    for (int64_t i = 0; i < n; ++i) {
        const double v = src[i];
        const int iv = (int)v;
        ok &= (double)iv == v && (unsigned)iv <= 32767u;
        dst[i] = (uint16_t)iv;
    }
    return ok;
}
static bool narrow_chunk_u8_scalar(const double* src, int64_t n, uint8_t* dst) {
    bool ok = true;
    for (int64_t i = 0; i < n; ++i) {
        const double c = src[i];
        const double k = std::nearbyint(c * 255.0);
        ok &= k >= 0.0 && k <= 255.0 && k / 255.0 == c;
        dst[i] = (uint8_t)(k >= 0.0 && k <= 255.0 ? k : 0.0);
    }
    return ok;
}
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
// four values per step (AVX2): the packing must run several times faster than the PCIe link it feeds
__attribute__((target("avx2"))) static bool narrow_chunk_u16_avx2(const double* src, int64_t n, uint16_t* dst) {
    __m256d bad = _mm256_setzero_pd();
    const __m256d lo = _mm256_set1_pd(0.0), hi = _mm256_set1_pd(32767.0);
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const __m256d v = _mm256_loadu_pd(src + i);
        const __m128i iv = _mm256_cvttpd_epi32(v);
        const __m256d back = _mm256_cvtepi32_pd(iv);
        // not equal after the round trip (also NaN), below 0 or above 32767
        bad = _mm256_or_pd(bad, _mm256_or_pd(_mm256_cmp_pd(back, v, _CMP_NEQ_UQ), _mm256_or_pd(_mm256_cmp_pd(v, lo, _CMP_LT_OQ), _mm256_cmp_pd(v, hi, _CMP_GT_OQ))));
        const __m128i p16 = _mm_packus_epi32(iv, iv);
        _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i), p16);
    }
    bool ok = _mm256_movemask_pd(bad) == 0;
    if (i < n) ok &= narrow_chunk_u16_scalar(src + i, n - i, dst + i);
    return ok;
}
__attribute__((target("avx2"))) static bool narrow_chunk_u8_avx2(const double* src, int64_t n, uint8_t* dst) {
    __m256d bad = _mm256_setzero_pd();
    const __m256d lo = _mm256_set1_pd(0.0), hi = _mm256_set1_pd(255.0), s255 = _mm256_set1_pd(255.0);
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const __m256d c = _mm256_loadu_pd(src + i);
        const __m256d k = _mm256_round_pd(_mm256_mul_pd(c, s255), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
        const __m256d q = _mm256_div_pd(k, s255);
        bad = _mm256_or_pd(bad, _mm256_or_pd(_mm256_cmp_pd(q, c, _CMP_NEQ_UQ), _mm256_or_pd(_mm256_cmp_pd(k, lo, _CMP_LT_OQ), _mm256_cmp_pd(k, hi, _CMP_GT_OQ))));
        const __m128i iv = _mm256_cvttpd_epi32(k);
        const __m128i p16 = _mm_packus_epi32(iv, iv);
        const __m128i p8 = _mm_packus_epi16(p16, p16);
        const int w = _mm_cvtsi128_si32(p8);
        memcpy(dst + i, &w, 4);
    }
    bool ok = _mm256_movemask_pd(bad) == 0;
    if (i < n) ok &= narrow_chunk_u8_scalar(src + i, n - i, dst + i);
    return ok;
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2");
#else
static const bool g_have_avx2 = false;
static bool narrow_chunk_u16_avx2(const double* s, int64_t n, uint16_t* d) { return narrow_chunk_u16_scalar(s, n, d); }
static bool narrow_chunk_u8_avx2(const double* s, int64_t n, uint8_t* d) { return narrow_chunk_u8_scalar(s, n, d); }
#endif
static bool narrow_chunk_u16(const double* src, int64_t rows, uint16_t* dst) {
    return g_have_avx2 ? narrow_chunk_u16_avx2(src, 3 * rows, dst) : narrow_chunk_u16_scalar(src, 3 * rows, dst);
}
static bool narrow_chunk_u8(const double* src, int64_t rows, uint8_t* dst) {
    return g_have_avx2 ? narrow_chunk_u8_avx2(src, 3 * rows, dst) : narrow_chunk_u8_scalar(src, 3 * rows, dst);
}

static HostNarrow* host_narrow_get(pccm_ctx* ctx);

// rows of 3 float64 (packed) -> rows of 3 uint16 / uint8 on the device.  *ok = false: some value does not fit (nothing
// usable was produced).  The copies run on `s`.
static int upload_narrowed(pccm_ctx* ctx, const double* src, int64_t n, int es, cudaStream_t s, unsigned char* d, bool* ok) {
    *ok = false;
    HostNarrow* H = host_narrow_get(ctx);
    if (!H) return PCCM_OK;
    const int64_t nchunks = (n + kNarrowChunk - 1) / kNarrowChunk;
    std::vector<std::atomic<int>> state((size_t)nchunks);       // 0 = not packed yet, 1 = packed, 2 = packed but a value did not fit
    for (auto& f : state) f.store(0, std::memory_order_relaxed);
    std::atomic<int64_t> released{std::min<int64_t>(nchunks, kNarrowSlots)};     // chunks whose slot may be written
    std::atomic<bool> bad{false};
    for (int k = 0; k < kNarrowSlots; ++k)
        if (H->slot_used[k]) { cudaEventSynchronize(H->slot_free[k]); H->slot_used[k] = false; }
    H->pool->start(nchunks, [&](int64_t i) {
        while (i >= released.load(std::memory_order_acquire)) {
            if (bad.load(std::memory_order_relaxed)) { state[(size_t)i].store(2, std::memory_order_release); return; }
            std::this_thread::yield();
        }
        const int64_t r0 = i * kNarrowChunk, rows = std::min(kNarrowChunk, n - r0);
        unsigned char* slot = H->ring + (size_t)(i % kNarrowSlots) * kNarrowChunk * 6;
        const bool fits = bad.load(std::memory_order_relaxed) ? false
                          : (es == 2 ? narrow_chunk_u16(src + 3 * r0, rows, reinterpret_cast<uint16_t*>(slot))
                                     : narrow_chunk_u8(src + 3 * r0, rows, slot));
        if (!fits) bad.store(true, std::memory_order_relaxed);
        state[(size_t)i].store(fits ? 1 : 2, std::memory_order_release);
    });
    cudaError_t err = cudaSuccess;
    for (int64_t i = 0; i < nchunks; ++i) {
        int st;
        while ((st = state[(size_t)i].load(std::memory_order_acquire)) == 0) std::this_thread::yield();
        if (st == 2 || bad.load(std::memory_order_relaxed)) { bad.store(true); released.store(nchunks, std::memory_order_release); break; }
        const int64_t r0 = i * kNarrowChunk, rows = std::min(kNarrowChunk, n - r0);
        const int k = (int)(i % kNarrowSlots);
        if (err == cudaSuccess) err = cudaMemcpyAsync(d + (size_t)r0 * 3 * es, H->ring + (size_t)k * kNarrowChunk * 6, (size_t)rows * 3 * es, cudaMemcpyHostToDevice, s);
        if (err == cudaSuccess) err = cudaEventRecord(H->slot_free[k], s);
        H->slot_used[k] = true;
        if (i + kNarrowSlots < nchunks) {               // the chunk that will reuse this slot may go once this copy has landed
            if (err == cudaSuccess) err = cudaEventSynchronize(H->slot_free[k]);
            H->slot_used[k] = false;
            released.store(i + kNarrowSlots + 1, std::memory_order_release);
        }
    }
    H->pool->wait();
    if (err != cudaSuccess) return fail(ctx, PCCM_ERR_CUDA, "narrowed upload: %s", cudaGetErrorString(err));
    *ok = !bad.load();
    return PCCM_OK;
}

static int upload_rows(pccm_ctx* ctx, const void* src, int dtype, int64_t n, int64_t* stride, int mem_kind,
                       const void** dev, void** owned) {
    const size_t es = dtype_size(dtype);
    if (es == 0) return fail(ctx, PCCM_ERR_INVALID, "bad dtype %d", dtype);
    if (*stride == 0) *stride = (int64_t)(3 * es);
    if (*stride < (int64_t)(3 * es) || (*stride % (int64_t)es) != 0) return fail(ctx, PCCM_ERR_INVALID, "bad row stride %lld", (long long)*stride);
    *owned = nullptr;
    if (mem_kind == PCCM_DEVICE) {
        *dev = src;
        return PCCM_OK;
    }
    unsigned char* d = nullptr;
    const size_t bytes = (size_t)n * (size_t)*stride;
    CK(dalloc(ctx, &d, bytes));
    CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = d;
    *owned = d;
    return PCCM_OK;
}

// --------------------------------------------------------------------------------------
// C ABI: context
// --------------------------------------------------------------------------------------
extern "C" int pccm_version(void) { return PCCM_VERSION; }

extern "C" const char* pccm_last_error(const pccm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int pccm_ctx_create(int device, void* stream, pccm_ctx** out) {
    pccm_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, PCCM_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, PCCM_ERR_CUDA, "no CUDA device (%s); libpccm has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, PCCM_ERR_INVALID, "device %d out of range (%d)", device, count);
    ctx = new pccm_ctx();
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, PCCM_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); }
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(nullptr, PCCM_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        ctx->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->copy_stream = nullptr;
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) { ctx->copy_stream = nullptr; }
    if (cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, PCCM_ERR_CUDA, "event creation failed");
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = (size_t)16 << 30;
        ctx->dev_cache_cap = total_b / 4;             // released blocks kept for reuse: at most a quarter of the device
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    for (int k = 1023; k >= 0; --k) ctx->flag_slots.push_back(k);
    if (cudaMallocHost(&ctx->pinned, pccm_ctx::kPinned) != cudaSuccess || cudaMalloc(&ctx->dscratch, pccm_ctx::kScratch) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, PCCM_ERR_CUDA, "scratch allocation failed");
    }
    cudaMemset(ctx->dscratch, 0, pccm_ctx::kScratch);
    {
        double lut[256];
        for (int k = 0; k < 256; ++k) lut[k] = (double)k / 255.0;   // the quotient Open3D / numpy store for 8-bit colours
        cudaMemcpy(static_cast<char*>(ctx->dscratch) + 32768, lut, sizeof lut, cudaMemcpyHostToDevice);
    }
    if (const char* s = getenv("PCCM_CELL_SHIFT")) ctx->cell_override_shift = atoi(s);
    if (const char* s = getenv("PCCM_SHORT_ROW")) ctx->short_row = (uint32_t)atoi(s);
    if (const char* s = getenv("PCCM_VOX")) ctx->use_vox = atoi(s) != 0;
    if (const char* s = getenv("PCCM_EAGER_PENCIL")) ctx->eager_pencil = atoi(s) != 0;
    if (const char* s = getenv("PCCM_VX_BLOCKS")) ctx->vx_search_blocks = std::max(1, atoi(s));
    if (const char* s = getenv("PCCM_PDL")) ctx->pdl = atoi(s) != 0;
    if (const char* s = getenv("PCCM_EPI_TMA")) ctx->epi_tma = atoi(s) != 0;
    if (const char* s = getenv("PCCM_VXYZ_ROWS")) ctx->vxyz_rows = atoi(s);
    if (const char* s = getenv("PCCM_STATS_TMA")) ctx->stats_tma = atoi(s) != 0;
    if (const char* s = getenv("PCCM_EPI_COMPACT")) ctx->epi_compact = atoi(s) != 0;
    if (const char* s = getenv("PCCM_DEV_CACHE")) ctx->dev_cache = atoi(s) != 0;
    if (const char* s = getenv("PCCM_MARK_SAMPLE")) ctx->mark_sample = std::max(0, atoi(s));
    if (const char* s = getenv("PCCM_SHARD_SEL")) ctx->shard_sel = atoi(s) != 0;
    if (const char* s = getenv("PCCM_HOST_NARROW")) ctx->host_narrow = atoi(s) != 0;
    if (const char* s = getenv("PCCM_HOST_THREADS")) ctx->host_threads = std::max(0, atoi(s));
    if (const char* s = getenv("PCCM_NORMALS_COUNTING")) ctx->normals_counting = atoi(s) != 0;
    if (const char* s = getenv("PCCM_CELL_SCALE")) ctx->cell_scale = atof(s);
    *out = ctx;
    return PCCM_OK;
}

extern "C" int pccm_ctx_destroy(pccm_ctx* ctx) {
    if (!ctx) return PCCM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    resolve_timers(ctx);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    for (auto e : ctx->cloud_events) cudaEventDestroy(e);
    if (ctx->narrow) {
        delete ctx->narrow->pool;
        for (auto e : ctx->narrow->slot_free) if (e) cudaEventDestroy(e);
        if (ctx->narrow->ring) cudaFreeHost(ctx->narrow->ring);
        delete ctx->narrow;
    }
    for (int d = 0; d < 2; ++d) { dfree(ctx, ctx->pp_idx[d]); dfree(ctx, ctx->pp_d2[d]); }
    dev_cache_trim(ctx, 0);
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(ctx->pinned);
    cudaFree(ctx->dscratch);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PCCM_OK;
}

static HostNarrow* host_narrow_get(pccm_ctx* ctx) {
    if (ctx->narrow) return ctx->narrow;
    HostNarrow* H = new HostNarrow();
    if (cudaMallocHost(reinterpret_cast<void**>(&H->ring), (size_t)kNarrowSlots * kNarrowChunk * 6) != cudaSuccess) { cudaGetLastError(); delete H; return nullptr; }
    for (auto& e : H->slot_free)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    int nt = ctx->host_threads > 0 ? ctx->host_threads : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    H->pool = new HostPool(nt);
    ctx->narrow = H;
    return H;
}

extern "C" int pccm_ctx_synchronize(pccm_ctx* ctx) {
    if (!ctx) return fail(nullptr, PCCM_ERR_INVALID, "ctx is NULL");
    CK(cudaStreamSynchronize(ctx->stream));
    return PCCM_OK;
}
extern "C" int pccm_ctx_set_profiling(pccm_ctx* ctx, int level) {
    if (!ctx) return fail(nullptr, PCCM_ERR_INVALID, "ctx is NULL");
    ctx->profiling = level;
    return PCCM_OK;
}
extern "C" int pccm_ctx_set_shard(pccm_ctx* ctx, int rank, int world) {
    if (!ctx) return fail(nullptr, PCCM_ERR_INVALID, "ctx is NULL");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, PCCM_ERR_INVALID, "bad rank/world %d/%d", rank, world);
    ctx->shard_rank = rank;
    ctx->shard_world = world;
    return PCCM_OK;
}
extern "C" int pccm_ctx_reset_timings(pccm_ctx* ctx) {
    if (!ctx) return fail(nullptr, PCCM_ERR_INVALID, "ctx is NULL");
    CK(cudaStreamSynchronize(ctx->stream));
    resolve_timers(ctx);
    ctx->tm = pccm_timings{};
    return PCCM_OK;
}
extern "C" int pccm_ctx_get_timings(pccm_ctx* ctx, pccm_timings* out) {
    if (!ctx || !out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaStreamSynchronize(ctx->stream));
    resolve_timers(ctx);
    *out = ctx->tm;
    return PCCM_OK;
}

// --------------------------------------------------------------------------------------
// C ABI: cloud
// --------------------------------------------------------------------------------------
extern "C" int pccm_cloud_destroy(pccm_ctx* ctx, pccm_cloud* c) {
    if (!ctx || !c) return PCCM_OK;
    cudaSetDevice(ctx->device);
    if (ctx->pp_owner[0] == c || ctx->pp_owner[1] == c) ctx->pp_owner[0] = ctx->pp_owner[1] = nullptr;   // per-point outputs of this pair are void
    wait_normals(ctx, c);
    if (c->nrm_ready) cudaEventDestroy(c->nrm_ready);
    if (c->rgb_pending) {
        cudaStreamWaitEvent(ctx->stream, c->rgb_ready, 0);   // the frees below are ordered on the context stream
        cudaEventSynchronize(c->rgb_ready);                  // ... and the pinned flag word goes back to the pool
    }
    if (c->rgb_ready) ctx->cloud_events.push_back(c->rgb_ready);
    if (c->flag_slot >= 0) ctx->flag_slots.push_back(c->flag_slot);
    dfree(ctx, c->raw_owned);
    dfree(ctx, c->raw_rgb_owned);
    dfree(ctx, c->d_stats);              // (the packed coordinates live in the same block)
    dfree(ctx, c->rgb_u8);
    dfree(ctx, c->rgb_f64);
    if (!c->normals_borrowed) dfree(ctx, c->normals);
    release_vox(ctx, c);
    if (c->shared) {
        if (--c->shared->refs == 0) {
            dfree(ctx, c->shared->recs);
            dfree(ctx, c->shared->table);
            delete c->shared;
        }
    } else {
        dfree(ctx, c->recs);
        dfree(ctx, c->row_start);
    }
    if (c->stats_fetched) cudaStreamSynchronize(ctx->stream);   // (a copy into the pinned slot is in flight: rare)
    delete c;
    return PCCM_OK;
}

static int set_normals_impl(pccm_ctx* ctx, pccm_cloud* c, const void* normals, int dtype, int64_t stride, int mem_kind,
                            bool may_borrow = false) {
    if (dtype != PCCM_F64 && dtype != PCCM_F32) return fail(ctx, PCCM_ERR_INVALID, "normals must be F64 or F32");
    const bool packed_f64 = dtype == PCCM_F64 && (stride == 0 || stride == 24);
    if (c->normals_borrowed) { c->normals = nullptr; c->normals_borrowed = false; }
    if (packed_f64 && mem_kind == PCCM_DEVICE && may_borrow) {
        // packed float64 rows already on the device: use them in place (caller keeps them alive)
        if (c->normals) dfree(ctx, c->normals);
        c->normals = const_cast<double*>(static_cast<const double*>(normals));
        c->normals_borrowed = true;
        c->has_normals = true;
        return PCCM_OK;
    }
    if (!c->normals) CK(dalloc(ctx, &c->normals, (size_t)c->n * 3));
    wait_normals(ctx, c);
    if (packed_f64) {   // already in the device format: one copy, no repack
        if (c->n) {
            if (mem_kind == PCCM_HOST && may_borrow && ctx->copy_stream) {
                // at cloud creation: the normals are first read by the query epilogue, so their
                // upload runs on the copy stream while statistics, sort and table build proceed
                if (!c->nrm_ready) CK(cudaEventCreateWithFlags(&c->nrm_ready, cudaEventDisableTiming));
                CK(cudaEventRecord(ctx->ev_fork, ctx->stream));            // after the allocation above
                CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
                CK(cudaMemcpyAsync(c->normals, normals, (size_t)c->n * 24, cudaMemcpyHostToDevice, ctx->copy_stream));
                CK(cudaEventRecord(c->nrm_ready, ctx->copy_stream));
                c->nrm_pending = true;
            } else {
                CK(cudaMemcpyAsync(c->normals, normals, (size_t)c->n * 24, mem_kind == PCCM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        c->has_normals = true;
        return PCCM_OK;
    }
    if (c->n && mem_kind == PCCM_HOST && may_borrow && ctx->copy_stream) {
        // float32 / strided host normals at cloud creation: upload and conversion on the copy stream as well
        const size_t es = dtype_size(dtype);
        if (stride == 0) stride = (int64_t)(3 * es);
        if (stride < (int64_t)(3 * es) || (stride % (int64_t)es) != 0) return fail(ctx, PCCM_ERR_INVALID, "bad row stride %lld", (long long)stride);
        unsigned char* d = nullptr;
        CK(dalloc(ctx, &d, (size_t)c->n * (size_t)stride));
        if (!c->nrm_ready) CK(cudaEventCreateWithFlags(&c->nrm_ready, cudaEventDisableTiming));
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));            // after the allocations
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
        CK(cudaMemcpyAsync(d, normals, (size_t)c->n * (size_t)stride, cudaMemcpyHostToDevice, ctx->copy_stream));
        pack_f64x3_kernel<<<(int)((c->n + 255) / 256), 256, 0, ctx->copy_stream>>>(d, dtype, stride, c->n, c->normals);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->nrm_ready, ctx->copy_stream));
        c->nrm_pending = true;
        // the staging buffer is freed on the context stream: order that free after the conversion without
        // making the context stream wait now -- keep it until the normals are first used
        c->nrm_stage = d;
        c->has_normals = true;
        return PCCM_OK;
    }
    const void* dev = nullptr;
    void* owned = nullptr;
    int rc = upload_rows(ctx, normals, dtype, c->n, &stride, mem_kind, &dev, &owned);
    if (rc) return rc;
    if (c->n) {
        pack_f64x3_kernel<<<(int)((c->n + 255) / 256), 256, 0, ctx->stream>>>(dev, dtype, stride, c->n, c->normals);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
    }
    dfree(ctx, owned);
    c->has_normals = true;
    return PCCM_OK;
}

extern "C" int pccm_cloud_create(pccm_ctx* ctx, const void* xyz, int xyz_dtype, int64_t n, int64_t xyz_stride,
                                 const void* rgb, int rgb_dtype, int64_t rgb_stride,
                                 const void* normals, int nrm_dtype, int64_t nrm_stride,
                                 int mem_kind, pccm_cloud** out) {
    if (!ctx || !out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (n < 0 || n > 0x7fffffffLL) return fail(ctx, PCCM_ERR_INVALID, "n=%lld out of range", (long long)n);
    if (n > 0 && !xyz) return fail(ctx, PCCM_ERR_INVALID, "xyz is NULL");
    if (xyz_dtype == PCCM_U8 || dtype_size(xyz_dtype) == 0) return fail(ctx, PCCM_ERR_INVALID, "bad xyz dtype %d", xyz_dtype);
    if (rgb && rgb_dtype != PCCM_F64 && rgb_dtype != PCCM_U8) return fail(ctx, PCCM_ERR_INVALID, "rgb must be F64 or U8");
    CK(cudaSetDevice(ctx->device));
    pccm_cloud* c = new pccm_cloud();
    c->n = n;
    int rc = PCCM_OK;
    {
        StageTimer t(ctx, &ctx->tm.upload_ms);
        c->raw_dtype = xyz_dtype;
        c->raw_stride = xyz_stride;
        bool narrowed = false;
        if (n >= kNarrowChunk && mem_kind == PCCM_HOST && xyz_dtype == PCCM_F64 && (xyz_stride == 0 || xyz_stride == 24) && ctx->host_narrow) {
            unsigned char* d = nullptr;
            CK(dalloc(ctx, &d, (size_t)n * 6));
            rc = upload_narrowed(ctx, static_cast<const double*>(xyz), n, 2, ctx->stream, d, &narrowed);
            if (narrowed) { c->raw_xyz = d; c->raw_owned = d; c->raw_dtype = PCCM_U16; c->raw_stride = 6; }
            else dfree(ctx, d);
        }
        if (n && !rc && !narrowed) rc = upload_rows(ctx, xyz, xyz_dtype, n, &c->raw_stride, mem_kind, &c->raw_xyz, &c->raw_owned);
    }
    if (rc) { pccm_cloud_destroy(ctx, c); return rc; }
    if (n) {
        StageTimer t(ctx, &ctx->tm.stats_ms);
        c->stats_blocks = (int)std::min<int64_t>((n + 2 * kStatsThreads - 1) / (2 * kStatsThreads), (int64_t)ctx->sm_count * 4);
        // ONE device block per cloud: partials, their fold, the colour flag, the DevStats record and -- for integer-capable
        // inputs -- the packed 8-byte coordinates the brick index is built from
        const bool want_packed = ctx->use_vox && xyz_dtype != PCCM_F32;
        const bool want_hist = want_packed && ctx->shard_world > 1;
        const size_t hist_bytes = want_hist ? (size_t)kZHistBins * sizeof(uint32_t) : 0;
        const size_t head = ((size_t)c->stats_blocks + 2) * sizeof(StatsPartial) + hist_bytes;
        unsigned char* blockp = nullptr;
        cudaError_t e = dalloc(ctx, &blockp, head + (want_packed ? (size_t)n * sizeof(uint2) : 0));
        if (e == cudaSuccess) {
            c->d_stats = reinterpret_cast<StatsPartial*>(blockp);
            c->d_rgbflag = reinterpret_cast<uint32_t*>(c->d_stats + c->stats_blocks);
            c->d_dev = reinterpret_cast<DevStats*>(c->d_stats + c->stats_blocks + 1);
            if (want_packed) c->packed = reinterpret_cast<uint2*>(blockp + head);
            if (want_hist) c->d_zhist = reinterpret_cast<uint32_t*>(c->d_stats + c->stats_blocks + 2);
            e = dzero(ctx, c->d_stats + c->stats_blocks, 2 * sizeof(StatsPartial) + hist_bytes, ctx->stream);
        }
        if (e != cudaSuccess) { pccm_cloud_destroy(ctx, c); return fail(ctx, PCCM_ERR_CUDA, "stats alloc: %s", cudaGetErrorString(e)); }
        const bool f64rows = c->raw_dtype == PCCM_F64 && c->raw_stride == 24 && (reinterpret_cast<uintptr_t>(c->raw_xyz) & 7u) == 0;
        if (f64rows)
            launch_chain(ctx, ctx->stats_tma ? stats_kernel<2> : stats_kernel<1>, c->stats_blocks, kStatsThreads, hist_bytes, ctx->stream, c->raw_xyz, c->raw_dtype, c->raw_stride, n,
                         nullptr, PCCM_F64, 0, c->d_stats, c->packed, c->d_dev, c->d_zhist);   // colours are classified apart
        else
            launch_chain(ctx, stats_kernel<0>, c->stats_blocks, kStatsThreads, hist_bytes, ctx->stream, c->raw_xyz, c->raw_dtype, c->raw_stride, n,
                         nullptr, PCCM_F64, 0, c->d_stats, c->packed, c->d_dev, c->d_zhist);
        ctx->tm.total_launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) { pccm_cloud_destroy(ctx, c); return fail(ctx, PCCM_ERR_CUDA, "stats launch: %s", cudaGetErrorString(e)); }
    }
    {   // attributes: needed by the epilogue only -> copy stream, behind the coordinates
        StageTimer t(ctx, &ctx->tm.upload_ms);
        if (rgb) rc = attach_colors(ctx, c, rgb, rgb_dtype, rgb_stride, mem_kind);
        if (!rc && normals) rc = set_normals_impl(ctx, c, normals, nrm_dtype, nrm_stride, mem_kind, true);
    }
    if (rc) { pccm_cloud_destroy(ctx, c); return rc; }
    *out = c;
    return PCCM_OK;
}

extern "C" int pccm_cloud_attach(pccm_ctx* ctx, pccm_cloud* c, const void* rgb, int rgb_dtype, int64_t rgb_stride,
                                 const void* normals, int nrm_dtype, int64_t nrm_stride, int mem_kind) {
    if (!ctx || !c) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, &ctx->tm.upload_ms);
    int rc = PCCM_OK;
    if (rgb) rc = attach_colors(ctx, c, rgb, rgb_dtype, rgb_stride, mem_kind);
    if (!rc && normals) rc = set_normals_impl(ctx, c, normals, nrm_dtype, nrm_stride, mem_kind, true);
    return rc;
}

extern "C" int pccm_cloud_info_get(pccm_ctx* ctx, pccm_cloud* c, pccm_cloud_info* out) {
    if (!ctx || !c || !out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    int rc = vox_settle(ctx, c);
    if (!rc) rc = ensure_stats(ctx, c);
    if (!rc && !c->rgb_spec) rc = ensure_colors(ctx, c);
    if (!rc && c->rgb_spec && c->vox) {      // 8-bit colours were assumed: find out
        rc = vox_fetch(ctx, c->vox);
        if (!rc) { CK(cudaStreamSynchronize(ctx->stream)); bool redo = false; rc = vox_adopt(ctx, c->vox, &redo); }
    }
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->n = c->n;
    out->data_kind = c->data_kind;
    out->index_kind = c->index_kind;
    out->has_colors = c->has_colors;
    out->colors_u8 = c->rgb_u8 != nullptr || (c->has_colors && (c->raw_rgb_dtype == PCCM_U8 || c->rgb_u8_ok) && !c->rgb_f64);
    out->has_normals = c->has_normals;
    out->indexed = c->index_kind >= 0;
    out->ny = c->grid.ny;
    out->nz = c->grid.nz;
    out->cell_size = c->cell_size;
    out->sharded = c->vox && c->vox->sharded ? c->vox->shard_world : 0;
    for (int a = 0; a < 3; ++a) { out->aabb_min[a] = c->mn[a]; out->aabb_max[a] = c->mx[a]; }
    return PCCM_OK;
}

extern "C" int pccm_cloud_set_normals(pccm_ctx* ctx, pccm_cloud* c, const void* normals, int nrm_dtype, int64_t nrm_stride, int mem_kind) {
    if (!ctx || !c || (!normals && c->n)) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    int rc = set_normals_impl(ctx, c, normals, nrm_dtype, nrm_stride, mem_kind);
    if (rc) return rc;
    if (mem_kind == PCCM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return PCCM_OK;
}

static int copy_out(pccm_ctx* ctx, void* dst, const void* src_dev, size_t bytes, int mem_kind) {
    if (bytes == 0) return PCCM_OK;
    CK(cudaMemcpyAsync(dst, src_dev, bytes, mem_kind == PCCM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    if (mem_kind == PCCM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return PCCM_OK;
}

extern "C" int pccm_cloud_get_normals(pccm_ctx* ctx, pccm_cloud* c, double* out, int mem_kind) {
    if (!ctx || !c || (!out && c->n)) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    if (!c->has_normals) return fail(ctx, PCCM_ERR_STATE, "cloud has no normals");
    CK(cudaSetDevice(ctx->device));
    wait_normals(ctx, c);
    return copy_out(ctx, out, c->normals, (size_t)c->n * 3 * sizeof(double), mem_kind);
}

// --------------------------------------------------------------------------------------
// index build
// --------------------------------------------------------------------------------------
static int bits_for(uint64_t v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

static int exclusive_scan(pccm_ctx* ctx, uint32_t* data, size_t count) {
    if (count <= kScanSmallMax) {
        scan_small_kernel<false><<<1, kScanSmallThreads, 0, ctx->stream>>>(data, (uint32_t)count, nullptr);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        return PCCM_OK;
    }
    const uint32_t nblocks = (uint32_t)((count + kScanTile - 1) / kScanTile);
    uint32_t* sums = nullptr;
    CK(dalloc(ctx, &sums, (size_t)nblocks));
    scan_tile_sums_kernel<<<nblocks, kScanThreads, 0, ctx->stream>>>(data, count, sums);
    scan_sums_kernel<<<1, kScanThreads, 0, ctx->stream>>>(sums, nblocks);
    scan_apply_kernel<<<nblocks, kScanThreads, 0, ctx->stream>>>(data, count, sums);
    ctx->tm.total_launches += 3;
    CK(cudaGetLastError());
    dfree(ctx, sums);
    return PCCM_OK;
}

static constexpr int64_t kMaxRows = 1ll << 26;

// choose h from the point density: surface-like clouds have spacing ~ sqrt(area / n)
static double auto_cell(const pccm_cloud* c) {
    double ex = c->mx[0] - c->mn[0] + 1e-300, ey = c->mx[1] - c->mn[1] + 1e-300, ez = c->mx[2] - c->mn[2] + 1e-300;
    double area = 2.0 * (ex * ey + ey * ez + ex * ez);
    double spacing = std::sqrt(area / (double)std::max<int64_t>(c->n, 1)) / 1.4;
    return 2.0 * spacing;
}

// grid of one cloud for a coordinate kind (cell size 0 = automatic)
static void choose_grid(pccm_ctx* ctx, const pccm_cloud* c, int kind, double cell_size, RowGrid& g, int& xbits) {
    double h = cell_size > 0 ? cell_size : auto_cell(c) * ctx->cell_scale;
    if (kind == PCCM_KIND_INT) {
        int shift = (int)std::lround(std::log2(std::max(h, 1.0)));
        if (cell_size <= 0 && ctx->cell_override_shift >= 0) shift = ctx->cell_override_shift;
        shift = std::max(0, std::min(shift, 15));
        for (;; ++shift) {
            g.shift = shift;
            g.iy0 = ((int)c->mn[1] >> shift) << shift;
            g.iz0 = ((int)c->mn[2] >> shift) << shift;
            g.ny = (((int)c->mx[1] - g.iy0) >> shift) + 1;
            g.nz = (((int)c->mx[2] - g.iz0) >> shift) + 1;
            if ((int64_t)g.ny * g.nz <= kMaxRows) break;
        }
        g.h = (double)(1 << g.shift); g.inv_h = 1.0 / g.h; g.y0 = g.iy0; g.z0 = g.iz0;
        xbits = std::max(1, bits_for((uint64_t)c->mx[0]));
    } else {
        double ey = c->mx[1] - c->mn[1], ez = c->mx[2] - c->mn[2];
        if (!(h > 0) || !std::isfinite(h)) h = 1.0;
        for (;;) {
            double ny = std::floor(ey / h) + 1, nz = std::floor(ez / h) + 1;
            if (ny * nz <= (double)kMaxRows) { g.ny = (int)ny; g.nz = (int)nz; break; }
            h *= 1.5;
        }
        g.h = h; g.inv_h = 1.0 / h; g.y0 = c->mn[1]; g.z0 = c->mn[2];
        double mag = 0;
        for (int a = 1; a < 3; ++a) mag = std::max(mag, std::max(std::fabs(c->mn[a]), std::fabs(c->mx[a])));
        g.slack = 1e-9 * (mag + h);
    }
}

extern "C" int pccm_cloud_build_index(pccm_ctx* ctx, pccm_cloud* c, double cell_size, int force_kind) {
    if (!ctx || !c) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    int rc = vox_settle(ctx, c);
    if (!rc) rc = ensure_stats(ctx, c);
    if (rc) return rc;
    rc = finish_colors(ctx, c);
    if (rc) return rc;
    int kind = c->data_kind;
    if (force_kind != PCCM_KIND_AUTO) {
        if (force_kind < c->data_kind || force_kind > PCCM_KIND_F64)
            return fail(ctx, PCCM_ERR_INVALID, "force_kind %d not allowed for data kind %d", force_kind, c->data_kind);
        kind = force_kind;
    }
    if (c->index_kind >= 0) {
        if (c->index_kind == kind) return PCCM_OK;
        return fail(ctx, PCCM_ERR_STATE, "cloud already indexed with kind %d", c->index_kind);
    }
    if (c->n && !c->raw_xyz) return fail(ctx, PCCM_ERR_STATE, "raw coordinates already released");
    const uint32_t n = (uint32_t)c->n;
    RowGrid g{};
    g.n = n;
    g.short_row = ctx->short_row;
    if (n == 0) {
        g.ny = g.nz = 1; g.h = g.inv_h = 1; g.shift = 0;
        c->grid = g; c->index_kind = kind; c->cell_size = 1;
        CK(dalloc(ctx, &c->row_start, 2));
        CK(cudaMemsetAsync(c->row_start, 0, 2 * sizeof(uint32_t), ctx->stream));
        return PCCM_OK;
    }
    int xbits = 0;
    choose_grid(ctx, c, kind, cell_size, g, xbits);
    c->cell_size = g.h;
    // the pair build with an empty partner: counting sort by row + per-row sort networks (no library sort)
    PairRaw R{};
    R.g[0] = g;
    R.g[1].ny = R.g[1].nz = 0;
    R.n[0] = n; R.n[1] = 0;
    R.xyz[0] = c->raw_xyz; R.dtype[0] = c->raw_dtype; R.stride[0] = c->raw_stride;
    R.table_off[0] = 0; R.table_off[1] = (uint32_t)((size_t)g.ny * g.nz);
    R.rgb_in_rec[0] = kind == PCCM_KIND_INT && c->rgb_u8 != nullptr;
    R.rgb[0] = c->rgb_u8; R.rgb_dtype[0] = PCCM_U8; R.rgb_stride[0] = sizeof(uchar4);
    pccm_cloud* cl[2] = {c, nullptr};
    return build_rowsort_kind(ctx, kind, cl, R);
}

// Index build without a radix sort, for every coordinate kind and for one cloud or a pair: counting sort by row
// (histogram ranks + scan) and a hand-written per-row sort by (x, index) -- 64-bit items for integer and float32
// coordinates, 128-bit ones for float64.  cl[1] may be null (single cloud; R.n[1] == 0).
template <class K> struct RecOf;
template <> struct RecOf<KInt> { typedef uint4 T; };
template <> struct RecOf<KF32> { typedef float4 T; };
template <> struct RecOf<KF64> { typedef RecF64 T; };

template <class K>
static int build_rowsort(pccm_ctx* ctx, pccm_cloud* cl[2], const PairRaw& R) {
    constexpr int KIND = K::kind;
    typedef typename RowItem<KIND>::T Item;
    const uint32_t n = R.n[0] + R.n[1];
    const size_t nrows = (size_t)R.g[0].ny * R.g[0].nz + (R.n[1] || cl[1] ? (size_t)R.g[1].ny * R.g[1].nz : 0);
    const size_t ntab = nrows + 1;
    const int threads = 256, blocks = (int)((n + threads - 1) / threads);
    SharedIndex* sh = new SharedIndex();
    uint32_t *rowof = nullptr, *rank = nullptr, *long_rows = nullptr;
    Item* items = nullptr;
    CK(dalloc(ctx, &sh->table, ntab));
    CK(dalloc(ctx, &rowof, (size_t)n));
    CK(dalloc(ctx, &rank, (size_t)n));
    CK(dalloc(ctx, &items, (size_t)n));
    CK(dalloc(ctx, &long_rows, nrows + 1));   // [0] = count, [1..] = row ids
    {
        StageTimer t(ctx, &ctx->tm.keys_ms);
        CK(cudaMemsetAsync(sh->table, 0, ntab * sizeof(uint32_t), ctx->stream));
        CK(cudaMemsetAsync(long_rows, 0, sizeof(uint32_t), ctx->stream));
        if (n) rowrank_pair_kernel<KIND><<<blocks, threads, 0, ctx->stream>>>(R, rowof, rank, sh->table);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
    }
    int rc;
    {
        StageTimer t(ctx, &ctx->tm.table_ms);
        rc = exclusive_scan(ctx, sh->table, ntab);
    }
    if (!rc && n) {
        StageTimer t(ctx, &ctx->tm.sort_ms);
        scatter_pair_kernel<KIND><<<blocks, threads, 0, ctx->stream>>>(R, rowof, rank, sh->table, items);
        const uint32_t wblocks = (uint32_t)((nrows * 32 + kRowSortThreads - 1) / kRowSortThreads);
        const int medium_grid = ctx->sm_count * (KIND == KIND_F64 ? 7 : 14);
        rowsort_warp_kernel<KIND><<<wblocks, kRowSortThreads, 0, ctx->stream>>>(sh->table, sh->table + 1, (uint32_t)nrows, nullptr, items, long_rows + 1, long_rows);
        rowsort_medium_kernel<KIND><<<medium_grid, kRowSortThreads, 0, ctx->stream>>>(sh->table, sh->table + 1, items, long_rows + 1, long_rows);
        rowsort_block_kernel<KIND><<<ctx->sm_count * 2, kRowSortThreads, 0, ctx->stream>>>(sh->table, sh->table + 1, items, long_rows + 1, long_rows, 0);
        ctx->tm.total_launches += 4;
        // rows longer than a block sorts in shared memory: cut into x buckets (a second counting sort inside the row), then
        // the same three kernels over the buckets.  Everything is sized by upper bounds; the counts stay on the device.
        const uint32_t max_split = n / RowSortCap<KIND>::block + 1;
        if (n > RowSortCap<KIND>::block) {
            const uint32_t max_buckets = n / kSplitTarget + max_split + 1;
            SplitRow* srows = nullptr;
            uint32_t *bmeta = nullptr, *long2 = nullptr;        // bmeta: [0..1] counters, then bcount / bstart / bend
            Item* items_alt = nullptr;
            CK(dalloc(ctx, &srows, (size_t)max_split));
            CK(dalloc(ctx, &bmeta, 2 + 3 * (size_t)max_buckets));
            CK(dalloc(ctx, &long2, (size_t)max_buckets + 1));
            CK(dalloc(ctx, &items_alt, (size_t)n));
            CK(cudaMemsetAsync(bmeta, 0, (2 + (size_t)max_buckets) * sizeof(uint32_t), ctx->stream));
            CK(cudaMemsetAsync(long2, 0, sizeof(uint32_t), ctx->stream));
            uint32_t *bcount = bmeta + 2, *bstart = bcount + max_buckets, *bend = bstart + max_buckets;
            rowsplit_find_kernel<KIND><<<(unsigned)((nrows + 255) / 256), 256, 0, ctx->stream>>>(sh->table, (uint32_t)nrows, srows, bmeta);
            rowsplit_kernel<KIND><<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(srows, bmeta, items, items_alt, rowof, rank, bcount, bstart, bend);
            rowsplit_copyback_kernel<KIND><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(srows, bmeta, items, items_alt);
            const uint32_t bblocks = (uint32_t)(((size_t)max_buckets * 32 + kRowSortThreads - 1) / kRowSortThreads);
            rowsort_warp_kernel<KIND><<<bblocks, kRowSortThreads, 0, ctx->stream>>>(bstart, bend, max_buckets, bmeta + 1, items, long2 + 1, long2);
            rowsort_medium_kernel<KIND><<<medium_grid, kRowSortThreads, 0, ctx->stream>>>(bstart, bend, items, long2 + 1, long2);
            rowsort_block_kernel<KIND><<<ctx->sm_count * 2, kRowSortThreads, 0, ctx->stream>>>(bstart, bend, items, long2 + 1, long2, 1);
            ctx->tm.total_launches += 6;
            dfree(ctx, srows); dfree(ctx, bmeta); dfree(ctx, long2); dfree(ctx, items_alt);
        }
        CK(cudaGetLastError());
    }
    if (!rc) {
        StageTimer t(ctx, &ctx->tm.reorder_ms);
        typename RecOf<K>::T* r = nullptr;
        CK(dalloc(ctx, &r, (size_t)n));
        if (n) reorder_items_pair_kernel<K><<<blocks, threads, 0, ctx->stream>>>(R, items, r);
        sh->recs = r;
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
    }
    dfree(ctx, rowof); dfree(ctx, rank); dfree(ctx, items); dfree(ctx, long_rows);
    if (rc) { dfree(ctx, sh->table); delete sh; return rc; }
    sh->refs = 0;
    for (int c = 0; c < 2; ++c) {
        pccm_cloud* p = cl[c];
        if (!p) continue;           // pencil index built on demand from brick records: the partner may be gone
        sh->refs++;
        p->shared = sh;
        p->recs = sh->recs;
        p->row_start = sh->table + R.table_off[c];
        p->base = c ? R.n[0] : 0u;
        p->grid = R.g[c];
        p->cell_size = R.g[c].h;
        p->index_kind = KIND;
        p->rgb_in_rec = R.rgb_in_rec[c] != 0;
        p->packed = nullptr;
        dfree(ctx, p->raw_owned); p->raw_owned = nullptr; p->raw_xyz = nullptr;
        if (p->rgb_in_rec) { dfree(ctx, p->raw_rgb_owned); p->raw_rgb_owned = nullptr; p->raw_rgb = nullptr; }
    }
    return PCCM_OK;
}
static int build_rowsort_kind(pccm_ctx* ctx, int kind, pccm_cloud* cl[2], const PairRaw& R) {
    if (kind == PCCM_KIND_INT) return build_rowsort<KInt>(ctx, cl, R);
    if (kind == PCCM_KIND_F32) return build_rowsort<KF32>(ctx, cl, R);
    return build_rowsort<KF64>(ctx, cl, R);
}
static int build_pair_rowsort(pccm_ctx* ctx, pccm_cloud* cl[2], const PairRaw& R) { return build_rowsort<KInt>(ctx, cl, R); }

// Occupancy-brick index of a KInt pair (pccm_vox.cuh), enqueued WITHOUT waiting for anything: the bounding
// boxes, the number of occupied bricks and the number of distinct voxels stay on the device (VoxPlan); arrays are
// sized by capacities.  Whether the assumptions held (integer coordinates, capacities) is looked at when the host
// synchronises anyway -- the result read-back of pccm_pair_eval, or the first call that needs host-side knowledge
// of the index (vox_settle) -- and only then is the build repeated with exact capacities or handed to the pencil path.
static constexpr size_t kVoxPinnedOffset = 12288;   // bytes into ctx->pinned: VoxPlan + colour flags + counters
static constexpr uint64_t kVoxMaxDirBits = 1ull << 30;
static constexpr uint32_t kVoxDefaultDirWords = 1u << 20;

static uint32_t pow2_ceil(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

static int finish_colors(pccm_ctx* ctx, pccm_cloud* c);

// 8-bit colour array of a cloud without a host wait: float64 colours still being classified on the copy stream are
// packed speculatively (the kernel re-checks every channel and raises the cloud's colour flag, read back at the
// next settle point).  Returns with c->rgb_u8 set when the colours are (assumed) 8-bit.
static int colors_u8_async(pccm_ctx* ctx, pccm_cloud* c) {
    if (!c->has_colors || c->rgb_u8 || c->rgb_f64 || c->n == 0) return PCCM_OK;
    if (c->rgb_pending && cudaEventQuery(c->rgb_ready) == cudaSuccess) { const int rc = ensure_colors(ctx, c); if (rc) return rc; }
    if (!c->rgb_pending && !c->rgb_unclassified) return finish_colors(ctx, c);        // classification known: the ordinary path
    if (c->rgb_pending) CK(cudaStreamWaitEvent(ctx->stream, c->rgb_ready, 0));        // device-side wait for the upload + classification
    CK(dalloc(ctx, &c->rgb_u8, (size_t)c->n + 4));
    const int threads = 256, blocks = (int)((c->n + threads - 1) / threads);
    if (c->raw_rgb_dtype == PCCM_U8) launch_pack_u8(ctx, c, c->rgb_u8);
    else launch_chain(ctx, pack_rgb_u8_check_kernel, blocks, threads, 0, ctx->stream, c->raw_rgb, c->raw_rgb_stride, c->n, c->rgb_u8, c->d_rgbflag);
    ctx->tm.total_launches++;
    CK(cudaGetLastError());
    c->rgb_spec = true;                                        // the raw colours stay until the flag has been read
    c->rgb_unclassified = false;
    return PCCM_OK;
}

static int vox_enqueue(pccm_ctx* ctx, pccm_cloud* cl[2], double cell_size, int force_kind, uint32_t cap_dirw, uint32_t cap_blk,
                       bool full_need = false) {
    StageTimer t(ctx, &ctx->tm.vox_build_ms);
    const uint32_t n_total = (uint32_t)(cl[0]->n + cl[1]->n);
    SharedVox* v = new SharedVox();
    v->cap_dirw = cap_dirw; v->cap_blk = cap_blk;
    v->cell_size = cell_size; v->force_kind = force_kind;
    auto bail = [&](int rc) { free_vox(ctx, v); return rc; };
#define CKV(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return bail(fail(ctx, PCCM_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_))); \
    } while (0)
    const uint32_t ndirblocks = cap_dirw / kVxDirChunk + 1;
    const uint32_t nbrickchunks = cap_blk / kVxBrickChunk + 2;
    // one allocation, sliced (a stream-ordered allocation costs a microsecond or two of host time, and the host is what
    // feeds the GPU here); the scratch slot list lives at the end and is simply never read again
    size_t off = 0;
    auto slice = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_plan = slice(sizeof(VoxPlan));
    const size_t o_shard = slice(sizeof(ShardPlan));
    const size_t o_selcnt = slice(256);             // [0] selected points, [16 ..] bricks per size class (directly in front of
                                                    // the directory bits: one zeroing pass clears all of it)
    const size_t o_dirbits = slice((size_t)cap_dirw * 4), o_dirpre = slice(((size_t)cap_dirw + 1) * 4), o_dirsums = slice((size_t)ndirblocks * 4);
    const size_t o_dirbytes = slice(PCCM_DIR_BYTES ? (size_t)cap_dirw * 32 : 0);
    const size_t o_rows = slice((size_t)cap_blk * kVxRows * 8);
    const size_t o_bricksums = slice((size_t)nbrickchunks * 4);
    const size_t o_border = slice((size_t)kVxSizeClasses * cap_blk * 4);
    const size_t o_vxyz = slice((size_t)n_total * 8), o_vkey = slice((size_t)n_total * 8), o_prank = slice((size_t)n_total * 4);
    const size_t o_pslot = slice((size_t)n_total * 4);
    // Scattered 8-byte stores of the coordinates are cheap while the voxel arrays sit in the L2 and ~0.8 ms of DRAM
    // read-modify-writes on a 10 M + 10 M pair; the row walk costs a few microseconds per million voxels either way.
    const bool sharded_here = ctx->shard_world > 1 && cl[0]->d_zhist && cl[1]->d_zhist && !full_need;
    const size_t indexed_here = sharded_here ? (size_t)n_total / (size_t)ctx->shard_world : (size_t)n_total;
    const bool vxyz_rows = ctx->vxyz_rows >= 0 ? ctx->vxyz_rows != 0 : indexed_here >= ((size_t)4 << 20);
    const size_t o_bkey = slice(vxyz_rows ? (size_t)cap_blk * 4 : 0);
    const bool sharded = ctx->shard_world > 1 && cl[0]->d_zhist && cl[1]->d_zhist;
    const bool use_sel = sharded && !full_need && ctx->shard_sel;
    const size_t o_sel = slice(use_sel ? (size_t)n_total * 4 : 0);
    CKV(dalloc(ctx, &v->arena, off));
    v->dplan = reinterpret_cast<VoxPlan*>(v->arena + o_plan);
    v->dirbits = reinterpret_cast<uint32_t*>(v->arena + o_dirbits);
    v->dirpre = reinterpret_cast<uint32_t*>(v->arena + o_dirpre);
    v->dirsums = reinterpret_cast<uint32_t*>(v->arena + o_dirsums);
    v->rows = reinterpret_cast<uint2*>(v->arena + o_rows);
    v->bricksums = reinterpret_cast<uint32_t*>(v->arena + o_bricksums);
    v->bcursor = reinterpret_cast<uint32_t*>(v->arena + o_selcnt) + 16;
    v->border = reinterpret_cast<uint32_t*>(v->arena + o_border);
    v->vxyz = reinterpret_cast<uint2*>(v->arena + o_vxyz);
    v->vkey = reinterpret_cast<uint2*>(v->arena + o_vkey);
    v->prank = reinterpret_cast<uint32_t*>(v->arena + o_prank);
    uint32_t* pslot = reinterpret_cast<uint32_t*>(v->arena + o_pslot);
    v->sharded = sharded;
    v->full_need = full_need;
    v->shard_rank = ctx->shard_rank; v->shard_world = ctx->shard_world;
    if (v->sharded) {
        v->dshard = reinterpret_cast<ShardPlan*>(v->arena + o_shard);
        launch_chain(ctx, vx_shardplan_kernel, 1, 1024, 0, ctx->stream, cl[0]->d_zhist, cl[1]->d_zhist, v->shard_rank, v->shard_world, v->dshard);
        ctx->tm.total_launches++;
    }
#if PCCM_DIR_BYTES
    CKV(cudaMemsetAsync(v->arena + o_dirbytes, 0, (size_t)cap_dirw * 32, ctx->stream));
#else
    (void)o_dirbytes;
    CKV(dzero(ctx, v->arena + o_selcnt, (o_dirbits - o_selcnt) + (size_t)cap_dirw * sizeof(uint32_t), ctx->stream));
#endif
    VoxBuildArgs A{};
    A.shard = v->dshard;
    A.full_need = full_need ? 1 : 0;
    A.sel = use_sel ? reinterpret_cast<uint32_t*>(v->arena + o_sel) : nullptr;
    A.sel_count = use_sel ? reinterpret_cast<uint32_t*>(v->arena + o_selcnt) : nullptr;
    for (int c = 0; c < 2; ++c) {
        pccm_cloud* p = cl[c];
        A.stats[c] = p->d_dev;
        A.packed[c] = p->packed;
        A.n[c] = v->n[c] = (uint32_t)p->n;
        // colours that are already on the device ride in the voxel records; colours still in flight are read
        // from the colour arrays by the epilogue -- the build never waits for them
        A.rgb[c] = nullptr;
        p->vox_rgb_in_recs = false;
        if (p->has_colors && !p->rgb_f64 && (!p->rgb_pending || cudaEventQuery(p->rgb_ready) == cudaSuccess)) {
            const int rcc = colors_u8_async(ctx, p);
            if (rcc) return bail(rcc);
            if (p->rgb_u8) { A.rgb[c] = p->rgb_u8; p->vox_rgb_in_recs = true; }
        }
    }
    A.cap_dirw = cap_dirw; A.cap_blk = cap_blk;
    A.dirbytes = v->arena + o_dirbytes;
    A.dirbits = v->dirbits; A.dirpre = v->dirpre; A.dirsums = v->dirsums; A.rows = v->rows;
    A.bcursor = v->bcursor; A.border = v->border;
    A.bricksums = v->bricksums; A.vxyz = v->vxyz; A.vkey = v->vkey; A.prank = v->prank; A.pslot = pslot; A.plan = v->dplan;
    A.bkey = vxyz_rows ? reinterpret_cast<uint32_t*>(v->arena + o_bkey) : nullptr;
    const int threads = 256;
    const int blocks_ilp = (int)(((n_total + kVxIlp - 1) / kVxIlp + threads - 1) / threads);
    const int brick_grid = (int)std::min<uint32_t>((uint32_t)ctx->sm_count * 8u, nbrickchunks);
    {
        const uint32_t per = (n_total + kVxIlp - 1) / kVxIlp;
        const uint32_t sample = ctx->mark_sample > 0 && per > 65536u ? per / (uint32_t)ctx->mark_sample : 0u;
        if (sample) {
            A.mark_lo = 0; A.mark_hi = sample;
            launch_chain(ctx, vx_mark_kernel, (sample + threads - 1) / threads, threads, 0, ctx->stream, A);
            ctx->tm.total_launches++;
        }
        A.mark_lo = sample; A.mark_hi = per;
        launch_chain(ctx, vx_mark_kernel, (per - sample + threads - 1) / threads, threads, 0, ctx->stream, A);
    }
    launch_chain(ctx, vx_dirsum_kernel, ndirblocks, threads, 0, ctx->stream, A);
    launch_chain(ctx, vx_dirscan_kernel, ndirblocks, threads, 0, ctx->stream, A);
    const int sel_grid = ctx->sm_count * 8;
    if (use_sel) launch_chain(ctx, vx_fill_sel_kernel, sel_grid, threads, 0, ctx->stream, A);
    else launch_chain(ctx, vx_fill_kernel, blocks_ilp, threads, 0, ctx->stream, A);
    launch_chain(ctx, vx_bricksum_kernel, brick_grid, threads, 0, ctx->stream, A);
    launch_chain(ctx, vx_rowbase_kernel, brick_grid, threads, 0, ctx->stream, A);
    if (use_sel) launch_chain(ctx, vx_place_sel_kernel, sel_grid, threads, 0, ctx->stream, A);
    else launch_chain(ctx, vx_place_kernel, blocks_ilp, threads, 0, ctx->stream, A);
    ctx->tm.total_launches += 7;
    CKV(cudaGetLastError());
#undef CKV
    v->pending = true;
    v->refs = 2;
    for (int c = 0; c < 2; ++c) {
        pccm_cloud* p = cl[c];
        release_vox(ctx, p);
        p->vox = v;
        p->vox_id = c;
        v->owner[c] = p;
        p->index_kind = PCCM_KIND_INT;      // assumed; vox_settle corrects it when the coordinates turn out not to be integers
        p->rgb_in_rec = false;              // (pencil records; decided when that index is built)
        p->vox_rgb_done = false;
    }
    return PCCM_OK;
}

// enqueue the read-back of everything the host needs to judge a pending build (the caller synchronises)
static int vox_fetch(pccm_ctx* ctx, SharedVox* v) {
    char* pin = static_cast<char*>(ctx->pinned) + kVoxPinnedOffset;
    CK(cudaMemcpyAsync(pin, v->dplan, sizeof(VoxPlan), cudaMemcpyDeviceToHost, ctx->stream));
    for (int c = 0; c < 2; ++c) {
        pccm_cloud* p = v->owner[c];
        uint32_t* flag = reinterpret_cast<uint32_t*>(pin + sizeof(VoxPlan)) + c;
        *flag = 0;
        if (p && p->rgb_spec) CK(cudaMemcpyAsync(flag, p->d_rgbflag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return PCCM_OK;
}

// the same as entries of ONE gather launch (see gather_to_host_kernel)
static void gather_add(GatherArgs& G, void* host_dst, const void* dev_src, size_t bytes) {
    G.src[G.n] = static_cast<const uint32_t*>(dev_src);
    G.dst[G.n] = static_cast<uint32_t*>(host_dst);
    G.words[G.n] = (uint32_t)((bytes + 3) / 4);
    G.n++;
}
static void vox_fetch_list(pccm_ctx* ctx, SharedVox* v, GatherArgs& G) {
    char* pin = static_cast<char*>(ctx->pinned) + kVoxPinnedOffset;
    gather_add(G, pin, v->dplan, sizeof(VoxPlan));
    for (int c = 0; c < 2; ++c) {
        pccm_cloud* p = v->owner[c];
        uint32_t* flag = reinterpret_cast<uint32_t*>(pin + sizeof(VoxPlan)) + c;
        *flag = 0;
        if (p && p->rgb_spec) gather_add(G, flag, p->d_rgbflag, sizeof(uint32_t));
    }
}

static int pair_build_classic(pccm_ctx* ctx, pccm_cloud* a, pccm_cloud* b, double cell_size, int force_kind, bool allow_vox);

// After the stream has been synchronised behind vox_fetch: adopt the plan.  *redo is set when the evaluation that
// was enqueued together with the build must be repeated (the index was rebuilt or the colour format changed).
static int vox_adopt(pccm_ctx* ctx, SharedVox* v, bool* redo) {
    *redo = false;
    const char* pin = static_cast<const char*>(ctx->pinned) + kVoxPinnedOffset;
    const bool was_pending = v->pending;
    if (was_pending) {
        memcpy(&v->hplan, pin, sizeof(VoxPlan));
        v->pending = false;
        v->view[0] = v->hplan.view[0]; v->view[1] = v->hplan.view[1];
    }
    pccm_cloud* cl[2] = {v->owner[0], v->owner[1]};
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(pin + sizeof(VoxPlan));
    int rc = PCCM_OK;
    for (int c = 0; c < 2 && !rc; ++c) {
        pccm_cloud* p = cl[c];
        if (!p) continue;
        if (p->rgb_spec) {                   // speculative 8-bit colours: keep or replace by the float64 array
            p->rgb_spec = false;
            p->rgb_pending = false;
            p->rgb_u8_ok = flags[c] == 0u;
            if (!p->rgb_u8_ok) {
                dfree(ctx, p->rgb_u8); p->rgb_u8 = nullptr;
                p->vox_rgb_in_recs = false;
                rc = finish_colors(ctx, p);
                *redo = true;
            } else {
                dfree(ctx, p->raw_rgb_owned); p->raw_rgb_owned = nullptr; p->raw_rgb = nullptr;
            }
        }
    }
    if (rc || !was_pending) return rc;
    // statistics of the clouds: an integer pair (the plan says so) is fully described by the plan's bounding boxes;
    // only when the coordinates turned out not to be integers do the per-block partials travel (a second wait)
    for (int c = 0; c < 2 && !rc; ++c) {
        pccm_cloud* p = cl[c];
        if (!p) continue;
        if (!(v->hplan.status & kVxStNotInt) && !p->stats_ready && !p->stats_fetched && p->n) {
            for (int a = 0; a < 3; ++a) { p->mn[a] = (double)v->hplan.mn[c][a]; p->mx[a] = (double)v->hplan.mx[c][a]; }
            p->data_kind = PCCM_KIND_INT;
            p->stats_ready = true;
        } else {
            rc = p->stats_fetched ? stats_adopt(ctx, p) : ensure_stats(ctx, p);
        }
    }
    if (rc) {                                    // (NaN / Inf coordinates): no index
        for (int c = 0; c < 2; ++c)
            if (cl[c]) { release_vox(ctx, cl[c]); cl[c]->index_kind = -1; }
        return rc;
    }
    const uint32_t st = v->hplan.status;
    if (st == 0) {
        for (int c = 0; c < 2; ++c) {
            pccm_cloud* p = cl[c];
            if (!p) continue;
            if (v->sharded && !v->full_need) continue;      // (a query that has to look beyond the halo needs the whole pair again)
            p->packed = nullptr;
            dfree(ctx, p->raw_owned); p->raw_owned = nullptr; p->raw_xyz = nullptr;   // vxyz holds every voxel, prank every point
        }
        return PCCM_OK;
    }
    // the assumptions did not hold: detach the index and rebuild
    *redo = true;
    const double cell = v->cell_size;
    const int fk = v->force_kind;
    const bool full_need = v->full_need;
    const int shard_rank = v->shard_rank, shard_world = v->sharded ? v->shard_world : 1;
    uint32_t cap_dirw = v->cap_dirw, cap_blk = v->cap_blk;
    bool retry_vox = !(st & kVxStNotInt) && cl[0] && cl[1];
    if (retry_vox && (st & kVxStDirOverflow)) {
        uint64_t words = 0;
        for (int c = 0; c < 2; ++c) {
            const pccm_cloud* p = cl[c];
            const uint64_t bits = (uint64_t)(((int)p->mx[0] >> 5) - ((int)p->mn[0] >> 5) + 1) * (uint64_t)(((int)p->mx[1] >> 3) - ((int)p->mn[1] >> 3) + 1) *
                                  (uint64_t)(((int)p->mx[2] >> 3) - ((int)p->mn[2] >> 3) + 1);
            if (bits > kVoxMaxDirBits) retry_vox = false;
            words += (bits + 31) / 32;
            if (c == 0) words = (words + kVxDirChunk - 1) / kVxDirChunk * kVxDirChunk;
        }
        cap_dirw = (uint32_t)std::min<uint64_t>(words + kVxDirChunk, 0xffffffffull);
        cap_blk = (uint32_t)std::min<uint64_t>((uint64_t)cl[0]->n + (uint64_t)cl[1]->n + 2, 0xffffffffull);      // (the brick count is not known yet)
    } else if (retry_vox && (st & kVxStBrickOverflow)) {
        cap_blk = v->hplan.view[0].nblk_total + 2;
    }
    for (int c = 0; c < 2; ++c)
        if (cl[c]) { release_vox(ctx, cl[c]); cl[c]->index_kind = -1; }
    if (retry_vox) {
        const int sr = ctx->shard_rank, sw = ctx->shard_world;      // the rebuilt index is split the way the first one was
        ctx->shard_rank = shard_rank; ctx->shard_world = shard_world;
        rc = vox_enqueue(ctx, cl, cell, fk, cap_dirw, cap_blk, full_need);
        ctx->shard_rank = sr; ctx->shard_world = sw;
        if (rc) return rc;
        SharedVox* nv = cl[0]->vox;
        rc = vox_fetch(ctx, nv);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        bool again = false;
        return vox_adopt(ctx, nv, &again);       // exact capacities: cannot overflow again
    }
    if (cl[0] && cl[1]) return pair_build_classic(ctx, cl[0], cl[1], cell, fk, false);
    pccm_cloud* only = cl[0] ? cl[0] : cl[1];
    return only ? pccm_cloud_build_index(ctx, only, cell, fk) : PCCM_OK;
}

// Make the host's knowledge of a cloud's index current (no-op unless a brick build is pending).
static int vox_settle(pccm_ctx* ctx, pccm_cloud* c) {
    if (!c->vox || !c->vox->pending) return PCCM_OK;
    SharedVox* v = c->vox;
    int rc = vox_fetch(ctx, v);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    bool redo = false;
    return vox_adopt(ctx, v, &redo);
}

// Colours of a brick-indexed cloud, on first use: one packed array in original order (uchar4 when every
// channel is k / 255, float64 otherwise) -- the epilogue streams through it for the query's own colour
// and, when the voxel records do not carry colours (they were still uploading at build time), gathers
// the neighbour's colour from it.
static int vox_colors(pccm_ctx* ctx, pccm_cloud* c) {
    if (!c->vox || c->vox_rgb_done || !c->has_colors) return PCCM_OK;
    const int rc = colors_u8_async(ctx, c);
    if (rc) return rc;
    c->vox_rgb_done = true;
    return PCCM_OK;
}

// A split pair whose queries (or a k-NN / hull call) have to look beyond the slab this rank indexed: index the whole
// pair after all (this rank still queries its own layers only).  Needs both clouds alive (their coordinates were kept).
static int vox_make_full(pccm_ctx* ctx, pccm_cloud* c) {
    SharedVox* v = c->vox;
    if (!v || !v->sharded || v->full_need) return PCCM_OK;
    pccm_cloud* cl[2] = {v->owner[0], v->owner[1]};
    if (!cl[0] || !cl[1]) return fail(ctx, PCCM_ERR_STATE, "a split pair needs both clouds alive to widen its index");
    const double cell = v->cell_size;
    const int fk = v->force_kind;
    const uint32_t cap_dirw = kVoxDefaultDirWords;
    const uint32_t cap_blk = std::max(1u << 16, (v->n[0] + v->n[1]) / 8) + 2;
    const int sr = ctx->shard_rank, sw = ctx->shard_world;
    ctx->shard_rank = v->shard_rank; ctx->shard_world = v->shard_world;
    for (int k = 0; k < 2; ++k) { release_vox(ctx, cl[k]); cl[k]->index_kind = -1; }
    int rc = vox_enqueue(ctx, cl, cell, fk, cap_dirw, cap_blk, true);
    ctx->shard_rank = sr; ctx->shard_world = sw;
    if (rc) return rc;
    return vox_settle(ctx, c);
}

// Pencil index of a brick-indexed cloud, built on first use (self k-NN, normals, hull prefilter,
// far queries) for both clouds of the pair from the brick records.
static int ensure_pencil(pccm_ctx* ctx, pccm_cloud* c) {
    { int rc = vox_settle(ctx, c); if (!rc) rc = vox_make_full(ctx, c); if (rc) return rc; }
    if (c->recs || c->row_start || !c->vox) return PCCM_OK;
    SharedVox* v = c->vox;
    pccm_cloud* cl[2] = {v->owner[0], v->owner[1]};
    PairRaw R{};
    for (int k = 0; k < 2; ++k) {
        RowGrid g{};
        g.short_row = ctx->short_row;
        if (cl[k] && cl[k]->recs) cl[k] = nullptr;    // (cannot happen: both are built together)
        if (cl[k]) {                                   // the pencil records carry the colours: their format must be final
            int rc = vox_colors(ctx, cl[k]);
            if (!rc && cl[k]->rgb_spec) {
                rc = vox_fetch(ctx, v);
                if (!rc) { CK(cudaStreamSynchronize(ctx->stream)); bool redo = false; rc = vox_adopt(ctx, v, &redo); }
            }
            if (rc) return rc;
        }
        if (cl[k]) {
            int xb = 0;
            g.n = (uint32_t)cl[k]->n;
            choose_grid(ctx, cl[k], PCCM_KIND_INT, v->cell_size, g, xb);
            R.n[k] = (uint32_t)cl[k]->n;
            // 8-bit colours ride in the pencil records: read from the packed colour array vox_colors made
            R.rgb_in_rec[k] = cl[k]->has_colors && cl[k]->rgb_u8 != nullptr;
            R.rgb[k] = cl[k]->rgb_u8; R.rgb_dtype[k] = PCCM_U8; R.rgb_stride[k] = sizeof(uchar4);
        } else {
            g.ny = g.nz = 1; g.h = g.inv_h = 1;
            R.n[k] = 0;
        }
        R.g[k] = g;
        R.xyz[k] = v->vxyz; R.dtype[k] = kDtypeVRec; R.stride[k] = sizeof(uint2);
    }
    R.table_off[0] = 0;
    R.table_off[1] = (uint32_t)((size_t)R.g[0].ny * R.g[0].nz);
    R.vprank = v->prank;
    R.vprank_off[0] = 0; R.vprank_off[1] = v->n[0];
    return build_pair_rowsort(ctx, cl, R);
}

extern "C" int pccm_pair_build_index(pccm_ctx* ctx, pccm_cloud* a, pccm_cloud* b, double cell_size, int force_kind) {
    if (!ctx || !a || !b) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    { int rc = vox_settle(ctx, a); if (!rc) rc = vox_settle(ctx, b); if (rc) return rc; }
    const uint64_t n_pair = (uint64_t)a->n + (uint64_t)b->n;
    bool spec = ctx->use_vox && a != b && a->n && b->n && a->index_kind < 0 && b->index_kind < 0 && a->packed && b->packed &&
                n_pair <= 0x7fffffffull && (force_kind == PCCM_KIND_AUTO || force_kind == PCCM_KIND_INT);
    if (spec && a->stats_ready && b->stats_ready && std::max(a->data_kind, b->data_kind) != PCCM_KIND_INT) spec = false;
    if (spec) {
        // (statistics the host already knows decide at once; otherwise the build is enqueued on the assumption that
        // both clouds are voxelised -- the statistics kernels of the stream say so on the device)
    }
    if (spec) {
        pccm_cloud* cl[2] = {a, b};
        const uint32_t n_total = (uint32_t)n_pair;
        const uint32_t cap_dirw = std::min(kVoxDefaultDirWords, std::max(1u << 12, pow2_ceil(n_total / (PCCM_DIR_BYTES ? 8 : 2))));
        const uint32_t cap_blk = n_total <= (1u << 16) ? n_total + 2 : std::max(1u << 16, n_total / 8) + 2;
        return vox_enqueue(ctx, cl, cell_size, force_kind, cap_dirw, cap_blk);
    }
    return pair_build_classic(ctx, a, b, cell_size, force_kind, true);
}

// The synchronous build: waits for the statistics, then the pencil index of both clouds in joint launches (integer
// pairs whose statistics were known before the brick build could be enqueued still get the brick index).
static int pair_build_classic(pccm_ctx* ctx, pccm_cloud* a, pccm_cloud* b, double cell_size, int force_kind, bool allow_vox) {
    int rc = ensure_stats(ctx, a);
    if (!rc) rc = ensure_stats(ctx, b);
    if (rc) return rc;
    int kind = std::max(a->data_kind, b->data_kind);
    if (force_kind != PCCM_KIND_AUTO) {
        if (force_kind < kind || force_kind > PCCM_KIND_F64) return fail(ctx, PCCM_ERR_INVALID, "force_kind %d not allowed for data kind %d", force_kind, kind);
        kind = force_kind;
    }
    pccm_cloud* cl[2] = {a, b};
    const bool separate = a == b || a->n == 0 || b->n == 0 || a->index_kind >= 0 || b->index_kind >= 0 ||
                          (uint64_t)a->n + (uint64_t)b->n > 0x7fffffffull;
    if (separate) {
        rc = pccm_cloud_build_index(ctx, a, cell_size, kind);
        if (!rc && a != b) rc = pccm_cloud_build_index(ctx, b, cell_size, kind);
        return rc;
    }
    (void)allow_vox;
    PairRaw R{};
    for (int c = 0; c < 2; ++c) {
        pccm_cloud* p = cl[c];
        if (!p->raw_xyz) return fail(ctx, PCCM_ERR_STATE, "raw coordinates already released");
        RowGrid g{};
        g.n = (uint32_t)p->n;
        g.short_row = ctx->short_row;
        int xb = 0;
        choose_grid(ctx, p, kind, cell_size, g, xb);
        R.g[c] = g;
        R.n[c] = (uint32_t)p->n;
        R.xyz[c] = p->raw_xyz; R.dtype[c] = p->raw_dtype; R.stride[c] = p->raw_stride;
    }
    R.table_off[0] = 0;
    R.table_off[1] = (uint32_t)((size_t)R.g[0].ny * R.g[0].nz);
    for (int c = 0; c < 2; ++c) {
        pccm_cloud* p = cl[c];
        rc = ensure_colors(ctx, p);
        if (rc) return rc;
        // 8-bit colours ride in the KInt record; every other combination keeps a colour array
        R.rgb_in_rec[c] = kind == PCCM_KIND_INT && p->has_colors && (p->raw_rgb_dtype == PCCM_U8 || p->rgb_u8_ok) && p->raw_rgb != nullptr;
        R.rgb[c] = p->raw_rgb; R.rgb_dtype[c] = p->raw_rgb_dtype; R.rgb_stride[c] = p->raw_rgb_stride;
        if (!R.rgb_in_rec[c]) { rc = finish_colors(ctx, p); if (rc) return rc; }
    }
    for (int c = 0; c < 2; ++c) cl[c]->packed = nullptr;
    return build_rowsort_kind(ctx, kind, cl, R);
}

// --------------------------------------------------------------------------------------
// queries
// --------------------------------------------------------------------------------------
static CloudView view_of(const pccm_cloud* c) {
    CloudView v;
    v.grid = c->grid;
    v.recs = c->recs;
    v.row_start = c->row_start;
    v.rgb_u8 = c->rgb_u8;
    v.rgb_f64 = c->rgb_f64;
    v.rgb_mode = !c->has_colors ? 0 : (c->rgb_in_rec ? 1 : (c->rgb_u8 ? 2 : 3));
    v.lut255 = nullptr;
    v.normals = c->normals;
    return v;
}

static constexpr size_t kTicketOffset = 4096;   // bytes into ctx->dscratch (zeroed at creation)
static constexpr size_t kChunksOffset = 16384;  // 64 BlockPartial records of the fold
static constexpr size_t kLutOffset = 32768;     // double[256]: k / 255.0

// One launch covers every requested direction; the per-direction reduced records land in
// ctx->dscratch[0..1] and are copied to ctx->pinned.
static int launch_query(pccm_ctx* ctx, int kind, QueryParams& P) {
    uint32_t max_tiles = 0, total_tiles = 0;
    for (int d = 0; d < P.ndirs; ++d) {
        P.dir[d].ntiles = (P.dir[d].qend - P.dir[d].qbegin + kQueryThreads - 1) / kQueryThreads;
        max_tiles = std::max(max_tiles, P.dir[d].ntiles);
        total_tiles += P.dir[d].ntiles;
    }
    P.rec_stride = max_tiles;
    BlockPartial* partials = nullptr;
    CK(dalloc(ctx, &partials, (size_t)P.rec_stride * 2 + 1));
    P.partials = partials;
    for (int d = 0; d < P.ndirs; ++d)
        P.dir[d].q.lut255 = P.dir[d].s.lut255 = reinterpret_cast<const double*>(static_cast<char*>(ctx->dscratch) + kLutOffset);
    P.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(ctx->dscratch) + kTicketOffset);
    P.out = static_cast<BlockPartial*>(ctx->dscratch);
    P.chunks = reinterpret_cast<BlockPartial*>(static_cast<char*>(ctx->dscratch) + kChunksOffset);
    if (total_tiles) {
        StageTimer t(ctx, &ctx->tm.query_ms, 1);
        if (kind == PCCM_KIND_INT) pair_query_kernel<KInt><<<total_tiles, kQueryThreads, 0, ctx->stream>>>(P);
        else if (kind == PCCM_KIND_F32) pair_query_kernel<KF32><<<total_tiles, kQueryThreads, 0, ctx->stream>>>(P);
        else pair_query_kernel<KF64><<<total_tiles, kQueryThreads, 0, ctx->stream>>>(P);
        ctx->tm.query_launches++;
        ctx->tm.total_launches++;
    }
    CK(cudaGetLastError());
    {
        StageTimer t(ctx, &ctx->tm.finalize_ms);
        launch_chain(ctx, finalize_kernel, P.ndirs * kFinalChunks, kFinalThreads, 0, ctx->stream, P);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->pinned, ctx->dscratch, 2 * sizeof(BlockPartial), cudaMemcpyDeviceToHost, ctx->stream));
    }
    dfree(ctx, partials);
    return PCCM_OK;
}

// Brick path of a symmetric evaluation: staged bit-scan search (one warp per query brick; it also
// finishes most undecided voxels itself), brick-ring search for the rest, per-point epilogue in
// the original order, common fold.  Works on a build that is still pending: nothing here needs to know
// the outcome of the build, the ONE synchronisation is the result read-back, and the build is judged
// there (*redo = the evaluation must be repeated: the index was rebuilt or the colour format changed).
// Voxels even the brick rings cannot certify (nearest point tens of voxels away) are finished by the
// pencil search in a second round.  Results land where launch_query puts them.
struct VoxScratch {           // freed on every exit path
    pccm_ctx* ctx;
    unsigned char* block = nullptr;
    ~VoxScratch() { dfree(ctx, block); }
};

static int launch_vox_query(pccm_ctx* ctx, int ndirs, pccm_cloud* qc[2], pccm_cloud* sc[2], QueryParams& Q, int rank, int world, bool* redo) {
    *redo = false;
    SharedVox* v = qc[0]->vox;
    VxParams P{};
    P.plan = v->dplan;
    P.ndirs = ndirs;
    P.normals_mode = Q.normals_mode;
    P.rank = v->sharded ? 0 : rank;           // (a split pair is cut by layers of z on the device, not by voxel ranks)
    P.world = v->sharded ? 1 : world;
    memcpy(P.T, Q.T, sizeof P.T);
    P.color_scale = Q.color_scale;
    const double* lut = reinterpret_cast<const double*>(static_cast<char*>(ctx->dscratch) + kLutOffset);
    const uint32_t n_total = v->n[0] + v->n[1];
    // one scratch block: [0..7] counters (undecided, far per direction, brick ticket), the four voxel lists, the voxel
    // answers, the reduction records
    uint32_t max_tiles = 0;
    for (int d = 0; d < ndirs; ++d) max_tiles = std::max(max_tiles, (v->n[qc[d]->vox_id] + kVxEpiTile - 1) / kVxEpiTile);
    const size_t todo_bytes = ((2 * (size_t)n_total + 8) * sizeof(uint32_t) + 255) & ~(size_t)255;
    const size_t vres_bytes = ((size_t)n_total * sizeof(uint4) + 255) & ~(size_t)255;
    const size_t part_bytes = ((size_t)max_tiles * 4 + 1) * sizeof(BlockPartial);
    VoxScratch sx{ctx};
    CK(dalloc(ctx, &sx.block, todo_bytes + vres_bytes + part_bytes));
    uint32_t* todo = reinterpret_cast<uint32_t*>(sx.block);
    uint4* vres = reinterpret_cast<uint4*>(sx.block + todo_bytes);
    BlockPartial* partials = reinterpret_cast<BlockPartial*>(sx.block + todo_bytes + vres_bytes);
    CK(dzero(ctx, todo, 8 * sizeof(uint32_t), ctx->stream));
    uint32_t rec_stride = 0, ntiles = 0, all_tiles = 0;
    for (int d = 0; d < ndirs; ++d) {
        VxDir& D = P.dir[d];
        D.qc = qc[d]->vox_id; D.sc = sc[d]->vox_id;
        D.nq = v->n[D.qc];
        D.qprank = v->prank + (D.qc ? v->n[0] : 0u);
        D.qa = Q.dir[d].q; D.sa = Q.dir[d].s;
        D.flags = Q.dir[d].flags;
        D.idx_out = Q.dir[d].idx_out; D.d2_out = Q.dir[d].d2_out;
        D.todo = todo + 8 + (d ? (size_t)v->n[P.dir[0].qc] : 0);
        D.far = todo + 8 + n_total + (d ? (size_t)v->n[P.dir[0].qc] : 0);
        D.npt_tiles = (D.nq + kVxEpiTile - 1) / kVxEpiTile;
        all_tiles += D.npt_tiles;
    }
    for (int d = 0; d < ndirs; ++d) {
        // blocks of the epilogue: one resident wave, shared out by the directions' sizes (a function of the two point
        // counts only -- float sums are reproducible on any GPU); small pairs keep one tile per block
        VxDir& D = P.dir[d];
        D.ntiles = all_tiles <= kVxEpiGrid ? D.npt_tiles
                                           : std::max<uint32_t>(1u, (uint32_t)((uint64_t)kVxEpiGrid * D.npt_tiles / all_tiles));
        rec_stride = std::max(rec_stride, 2u * D.ntiles);
        ntiles += D.ntiles;
    }
    for (int d = 0; d < ndirs; ++d) P.dir[d].rec_off = (uint32_t)d * rec_stride;
    P.partials = partials; P.vres = vres; P.counters = todo;
    P.bcursor = v->bcursor; P.border = v->border; P.border_cap = v->cap_blk;
    // a rank that evaluates a fraction of the points (split pair, voxel slice) queues them per block instead of walking every tile
    const bool compact = ctx->epi_compact && (v->sharded || world > 1);
    void (*epilogue_kernel)(VxParams) = compact ? vx_epilogue_kernel<false, true>
                                                : (ctx->epi_tma ? vx_epilogue_kernel<true, false> : vx_epilogue_kernel<false, false>);
    {
        StageTimer stage(ctx, &ctx->tm.query_ms, 1);     // the whole query stage (level 1: ONE event pair, so that the three
                                                         // kernels stay chained; the per-kernel split needs level 2)
        {
            StageTimer t(ctx, &ctx->tm.vox_search_ms, 2);
            launch_chain(ctx, vx_search_kernel, ctx->sm_count * ctx->vx_search_blocks, kVxThreads, 0, ctx->stream, P);
            ctx->tm.query_launches++;
        }
        {
            StageTimer t(ctx, &ctx->tm.vox_tail_ms, 2);
            launch_chain(ctx, vx_general_kernel, ctx->sm_count * 4, 128, 0, ctx->stream, P);
        }
        // The search reads coordinates only.  Colours and normals that are still being uploaded on the copy stream are
        // waited for HERE, so that the search runs underneath their copies and only the epilogue is left behind the last byte.
        for (int d = 0; d < ndirs; ++d) {
            if (P.dir[d].flags & PCCM_EVAL_COLOR) { const int rcc = vox_colors(ctx, qc[d]); if (rcc) return rcc; }
            if (P.dir[d].flags & PCCM_EVAL_D2) { wait_normals(ctx, qc[d]); wait_normals(ctx, sc[d]); }
        }
        for (int d = 0; d < ndirs; ++d) {
            VxDir& D = P.dir[d];
            if (P.dir[d].flags & PCCM_EVAL_COLOR) { const int rcc = vox_colors(ctx, sc[d]); if (rcc) return rcc; }
            // own colour: streamed from the packed array; neighbour colour: in the voxel answer when the records carry it
            D.qa.rgb_mode = !qc[d]->has_colors ? 0 : (qc[d]->rgb_u8 ? 2 : 3);
            D.sa.rgb_mode = !sc[d]->has_colors ? 0 : (sc[d]->vox_rgb_in_recs ? 1 : (sc[d]->rgb_u8 ? 2 : 3));
            D.qa.rgb_u8 = qc[d]->rgb_u8; D.qa.rgb_f64 = qc[d]->rgb_f64;
            D.sa.rgb_u8 = sc[d]->rgb_u8; D.sa.rgb_f64 = sc[d]->rgb_f64;
            D.qa.lut255 = D.sa.lut255 = lut;
        }
        {
            StageTimer t(ctx, &ctx->tm.vox_epilogue_ms, 2);
            launch_chain(ctx, epilogue_kernel, ntiles, kVxEpiThreads, 0, ctx->stream, P);
        }
        ctx->tm.total_launches += 3;
    }
    CK(cudaGetLastError());
    // common fold: same record layout as the pencil path
    Q.rec_stride = rec_stride;
    Q.partials = partials;
    Q.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(ctx->dscratch) + kTicketOffset);
    Q.out = static_cast<BlockPartial*>(ctx->dscratch);
    Q.chunks = reinterpret_cast<BlockPartial*>(static_cast<char*>(ctx->dscratch) + kChunksOffset);
    uint32_t* hcnt = reinterpret_cast<uint32_t*>(static_cast<char*>(ctx->pinned) + kVoxPinnedOffset + sizeof(VoxPlan) + 16);
    auto fold = [&](uint32_t passes, bool copy) -> int {
        StageTimer t(ctx, &ctx->tm.finalize_ms);
        uint32_t most = 0;
        for (int d = 0; d < ndirs; ++d) { Q.dir[d].ntiles = passes * P.dir[d].ntiles; most = std::max(most, Q.dir[d].ntiles); }
        if (most <= 2048) launch_chain(ctx, finalize_small_kernel, Q.ndirs, kFinalThreads, 0, ctx->stream, Q);
        else launch_chain(ctx, finalize_kernel, Q.ndirs * kFinalChunks, kFinalThreads, 0, ctx->stream, Q);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        if (copy) CK(cudaMemcpyAsync(ctx->pinned, ctx->dscratch, 2 * sizeof(BlockPartial), cudaMemcpyDeviceToHost, ctx->stream));
        return PCCM_OK;
    };
    int rc = fold(1, false);
    if (rc) return rc;
    {   // everything the host reads at its one wait, in one launch: results, counters, plan, statistics, colour flags
        GatherArgs G{};
        gather_add(G, ctx->pinned, ctx->dscratch, 2 * sizeof(BlockPartial));
        gather_add(G, hcnt, todo, 8 * sizeof(uint32_t));
        vox_fetch_list(ctx, v, G);
        CK(launch_chain(ctx, gather_to_host_kernel, 1, 256, 0, ctx->stream, G));
        ctx->tm.total_launches++;
    }
    CK(cudaStreamSynchronize(ctx->stream));          // the one host wait of build + evaluation
    rc = vox_adopt(ctx, v, redo);
    if (rc || *redo) return rc;
    ctx->tm.vox_undecided = (int64_t)hcnt[0] + hcnt[1];
    ctx->tm.vox_far = (int64_t)hcnt[2] + hcnt[3];
    ctx->tm.vox_tail = (int64_t)n_total - (int64_t)v->hplan.nvox_total;
    if (hcnt[2] + hcnt[3] > 0 && v->sharded && !v->full_need) {
        // a query of this rank has to look beyond the halo: index the whole pair, evaluate again
        rc = vox_make_full(ctx, qc[0]);
        *redo = true;
        return rc;
    }
    if (hcnt[2] + hcnt[3] > 0) {
        // second round: pencil search for the far voxels, the epilogue of their points, fold again
        for (int d = 0; d < ndirs; ++d) {
            rc = ensure_pencil(ctx, sc[d]);
            if (rc) return rc;
            VxDir& D = P.dir[d];
            D.sgrid = sc[d]->grid;
            D.srecs = static_cast<const uint4*>(sc[d]->recs);
            D.srow_start = sc[d]->row_start;
        }
        P.pass = 1;
        {
            StageTimer t(ctx, &ctx->tm.vox_tail_ms, 2);
            launch_chain(ctx, vx_far_kernel, ctx->sm_count * 8, 128, 0, ctx->stream, P);
            launch_chain(ctx, epilogue_kernel, ntiles, kVxEpiThreads, 0, ctx->stream, P);
            ctx->tm.total_launches += 2;
            CK(cudaGetLastError());
        }
        rc = fold(2, true);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return PCCM_OK;
}

static int check_pair(pccm_ctx* ctx, pccm_cloud* q, pccm_cloud* s) {
    if (q->index_kind < 0 || s->index_kind < 0) return fail(ctx, PCCM_ERR_STATE, "both clouds must be indexed (pccm_cloud_build_index)");
    if (q->index_kind != s->index_kind)
        return fail(ctx, PCCM_ERR_STATE, "clouds indexed with different kinds (%d vs %d); rebuild with force_kind", q->index_kind, s->index_kind);
    return PCCM_OK;
}

extern "C" int pccm_nn(pccm_ctx* ctx, pccm_cloud* query, pccm_cloud* search, int32_t* idx_out, double* d2_out, int mem_kind) {
    if (!ctx || !query || !search) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    int rc = check_pair(ctx, query, search);
    if (rc) return rc;
    if (query->n == 0) return PCCM_OK;
    if (search->n == 0) return fail(ctx, PCCM_ERR_INDEX, "search cloud is empty (reference: IndexError at cloud_pair.py:23)");
    const uint32_t nq = (uint32_t)query->n;
    int32_t* d_idx = nullptr;
    double* d_d2 = nullptr;
    if (idx_out) { if (mem_kind == PCCM_DEVICE) d_idx = idx_out; else CK(dalloc(ctx, &d_idx, (size_t)nq)); }
    if (d2_out) { if (mem_kind == PCCM_DEVICE) d_d2 = d2_out; else CK(dalloc(ctx, &d_d2, (size_t)nq)); }
    QueryParams P{};
    for (int attempt = 0; attempt < 3; ++attempt) {
        P = QueryParams{};
        P.ndirs = 1;
        P.dir[0].q = view_of(query); P.dir[0].s = view_of(search);
        P.dir[0].qbegin = query->base; P.dir[0].qend = query->base + nq; P.dir[0].flags = 0;
        P.dir[0].idx_out = d_idx; P.dir[0].d2_out = d_d2;
        P.normals_mode = 0; P.color_scale = 1;
        if (ctx->use_vox && query->vox && query->vox == search->vox && query != search && query->index_kind == PCCM_KIND_INT) {
            pccm_cloud* qc[2] = {query, nullptr};
            pccm_cloud* sc[2] = {search, nullptr};
            bool redo = false;
            rc = launch_vox_query(ctx, 1, qc, sc, P, 0, 1, &redo);
            if (!rc && redo) continue;               // the pending build was replaced: evaluate on what it became
        } else {
            rc = ensure_pencil(ctx, query);
            if (!rc) rc = ensure_pencil(ctx, search);
            if (!rc) rc = check_pair(ctx, query, search);
            if (!rc) {
                P.dir[0].q = view_of(query); P.dir[0].s = view_of(search);
                P.dir[0].qbegin = query->base; P.dir[0].qend = query->base + nq;
                rc = launch_query(ctx, query->index_kind, P);
            }
        }
        break;
    }
    if (mem_kind == PCCM_HOST) {
        if (!rc && idx_out) rc = copy_out(ctx, idx_out, d_idx, (size_t)nq * sizeof(int32_t), PCCM_HOST);
        if (!rc && d2_out) rc = copy_out(ctx, d2_out, d_d2, (size_t)nq * sizeof(double), PCCM_HOST);
        dfree(ctx, d_idx); dfree(ctx, d_d2);
    }
    return rc;
}

extern "C" int pccm_pair_eval(pccm_ctx* ctx, pccm_cloud* a, pccm_cloud* b, uint32_t flags,
                              const double* color_matrix, double color_scale, int normals_mode,
                              int rank, int world, pccm_pair_result* out) {
    if (!ctx || !a || !b || !out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, PCCM_ERR_INVALID, "bad rank/world %d/%d", rank, world);
    CK(cudaSetDevice(ctx->device));
    int rc = check_pair(ctx, a, b);
    if (rc) return rc;
    if (a->n == 0 || b->n == 0) return fail(ctx, PCCM_ERR_INDEX, "empty cloud (reference: IndexError at cloud_pair.py:23)");
    if ((flags & PCCM_EVAL_D2) && (!a->has_normals || !b->has_normals))
        return fail(ctx, PCCM_ERR_STATE, "D2 needs normals on both clouds (set or estimate them)");
    if (flags & PCCM_EVAL_COLOR) {
        if (!a->has_colors || !b->has_colors) return fail(ctx, PCCM_ERR_STATE, "colour metrics need colours on both clouds");
        if (!color_matrix) return fail(ctx, PCCM_ERR_INVALID, "color_matrix is NULL");
    }
    pccm_cloud* cl[2] = {a, b};
    // metric.py:148-152 indexes the OTHER cloud's normals with the query index: a direction
    // whose search cloud is shorter than its query cloud raises IndexError in the reference.
    uint32_t dflags[2] = {flags, flags};
    for (int d = 0; d < 2; ++d)
        if ((flags & PCCM_EVAL_D2) && normals_mode == PCCM_NORMALS_BY_QUERY_INDEX && cl[1 - d]->n < cl[d]->n)
            dflags[d] &= ~(uint32_t)PCCM_EVAL_D2;
    if (flags & PCCM_EVAL_PERPOINT) {
        for (int d = 0; d < 2; ++d) {
            const uint64_t n = (uint64_t)cl[d]->n;
            if (ctx->pp_n[d] != cl[d]->n) {
                dfree(ctx, ctx->pp_idx[d]); dfree(ctx, ctx->pp_d2[d]);
                ctx->pp_idx[d] = nullptr; ctx->pp_d2[d] = nullptr; ctx->pp_n[d] = 0;
                CK(dalloc(ctx, &ctx->pp_idx[d], (size_t)n));
                CK(dalloc(ctx, &ctx->pp_d2[d], (size_t)n));
                ctx->pp_n[d] = cl[d]->n;
            }
            if (world > 1) {  // entries outside this rank's slice stay -1 / NaN
                CK(cudaMemsetAsync(ctx->pp_idx[d], 0xff, n * sizeof(int32_t), ctx->stream));
                CK(cudaMemsetAsync(ctx->pp_d2[d], 0xff, n * sizeof(double), ctx->stream));
            }
        }
        ctx->pp_owner[0] = a; ctx->pp_owner[1] = b;
    }
    QueryParams P{};
    bool vox = false;
    pccm_cloud* sc[2] = {b, a};
    for (int attempt = 0; attempt < 3; ++attempt) {
        vox = ctx->use_vox && a->vox && a->vox == b->vox && a != b && a->index_kind == PCCM_KIND_INT && !(flags & PCCM_EVAL_TIE_AVERAGE);
        // attributes still on their way up (copy stream) are waited for where they are first read: at once on the pencil
        // path (one fused kernel), between the search and the epilogue on the brick path (launch_vox_query)
        if (!vox) {
            if (flags & PCCM_EVAL_D2) { wait_normals(ctx, a); wait_normals(ctx, b); }
            rc = ensure_pencil(ctx, a);
            if (!rc) rc = ensure_pencil(ctx, b);
            if (!rc) rc = check_pair(ctx, a, b);
            if (rc) return rc;
        }
        P = QueryParams{};
        P.ndirs = 2;
        P.normals_mode = normals_mode;
        if (color_matrix) memcpy(P.T, color_matrix, sizeof P.T);
        P.color_scale = color_scale;
        for (int d = 0; d < 2; ++d) {
            const uint64_t n = (uint64_t)cl[d]->n;
            DirParams& D = P.dir[d];
            D.q = view_of(cl[d]); D.s = view_of(cl[1 - d]);
            D.qbegin = cl[d]->base + (uint32_t)(n * (uint64_t)rank / (uint64_t)world);
            D.qend = cl[d]->base + (uint32_t)(n * (uint64_t)(rank + 1) / (uint64_t)world);
            D.flags = dflags[d];
            if (flags & PCCM_EVAL_PERPOINT) { D.idx_out = ctx->pp_idx[d]; D.d2_out = ctx->pp_d2[d]; }
        }
        if (vox) {
            bool redo = false;
            rc = launch_vox_query(ctx, 2, cl, sc, P, rank, world, &redo);
            if (!rc && redo) continue;               // the pending build was replaced: evaluate on what it became
        } else {
            rc = launch_query(ctx, a->index_kind, P);
        }
        break;
    }
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    const BlockPartial* r = static_cast<const BlockPartial*>(ctx->pinned);
    memset(out, 0, sizeof *out);
    for (int d = 0; d < 2; ++d) {
        pccm_dir_result& o = out->dir[d];
        o.n = vox ? (int64_t)r[d].sum_d1 : (int64_t)(P.dir[d].qend - P.dir[d].qbegin);   // brick path: slices are cut by voxel, the kernels count the points
        o.n_total = cl[d]->n;
        o.d1_exact_int = a->index_kind == PCCM_KIND_INT;
        o.d2_valid = (dflags[d] & PCCM_EVAL_D2) != 0;
        o.sum_d1_u64 = r[d].sum_d1_u64;
        o.sum_d1 = o.d1_exact_int ? (double)r[d].sum_d1_u64 : r[d].sum_d1;
        o.max_d1 = r[d].max_d1;
        o.sum_d2 = r[d].sum_d2; o.max_d2 = r[d].max_d2;
        for (int k = 0; k < 3; ++k) { o.color_sum[k] = r[d].csum[k]; o.color_max[k] = r[d].cmax[k]; }
    }
    return PCCM_OK;
}

extern "C" int pccm_pair_get(pccm_ctx* ctx, int which, int direction, void* out, int mem_kind) {
    if (!ctx || !out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    if (direction < 0 || direction > 1) return fail(ctx, PCCM_ERR_INVALID, "direction must be 0 or 1");
    if (!ctx->pp_idx[direction] || !ctx->pp_owner[0] || !ctx->pp_owner[1])
        return fail(ctx, PCCM_ERR_STATE, "no per-point results: call pccm_pair_eval with PCCM_EVAL_PERPOINT (and keep both clouds alive)");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->pp_n[direction];
    if (which == PCCM_GET_IDX) return copy_out(ctx, out, ctx->pp_idx[direction], n * sizeof(int32_t), mem_kind);
    if (which == PCCM_GET_D2) return copy_out(ctx, out, ctx->pp_d2[direction], n * sizeof(double), mem_kind);
    return fail(ctx, PCCM_ERR_INVALID, "bad selector %d", which);
}

// --------------------------------------------------------------------------------------
// self k-NN family
// --------------------------------------------------------------------------------------
static int launch_knn(pccm_ctx* ctx, pccm_cloud* c, KnnParams& P) {
    const uint32_t cnt = P.end - P.begin;
    if (cnt == 0) return PCCM_OK;
    uint32_t nblocks = (cnt + kKnnThreads - 1) / kKnnThreads;
    if (P.mode == KNN_NORMALS_FLAGGED) nblocks = std::min(nblocks, (uint32_t)ctx->sm_count * 2u);   // walks the todo list
    const size_t dsz = c->index_kind == PCCM_KIND_INT ? sizeof(uint32_t) : sizeof(double);
    const size_t smem = (size_t)P.k * kKnnThreads * (dsz + 2 * sizeof(uint32_t));
    if (smem > 200 * 1024) return fail(ctx, PCCM_ERR_UNSUPPORTED, "k=%d too large", P.k);
    StageTimer t(ctx, &ctx->tm.knn_ms, 1);
    if (c->index_kind == PCCM_KIND_INT) {
        CK(cudaFuncSetAttribute(knn_self_kernel<KInt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_self_kernel<KInt><<<nblocks, kKnnThreads, smem, ctx->stream>>>(P);
    } else if (c->index_kind == PCCM_KIND_F32) {
        CK(cudaFuncSetAttribute(knn_self_kernel<KF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_self_kernel<KF32><<<nblocks, kKnnThreads, smem, ctx->stream>>>(P);
    } else {
        CK(cudaFuncSetAttribute(knn_self_kernel<KF64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_self_kernel<KF64><<<nblocks, kKnnThreads, smem, ctx->stream>>>(P);
    }
    ctx->tm.knn_launches++;
    ctx->tm.total_launches++;
    CK(cudaGetLastError());
    return PCCM_OK;
}

// pccm_self_nn_minmax on the brick index.  Voxels that have no other voxel within their 27 neighbour bricks (isolated
// by more than 8 voxels) are finished one by one with the pencil search -- the slice keeps its share of the voxels
// whatever path its voxels take, so the slices of several ranks always cover every point exactly once.
struct SelfScratch {
    pccm_ctx* ctx;
    uint32_t *scratch = nullptr, *vself = nullptr, *und = nullptr;
    double *mm = nullptr, *d_pp = nullptr;
    bool own_pp = false;
    ~SelfScratch() { dfree(ctx, scratch); dfree(ctx, vself); dfree(ctx, und); dfree(ctx, mm); if (own_pp) dfree(ctx, d_pp); }
};

static int vox_self_nn(pccm_ctx* ctx, pccm_cloud* c, int64_t begin, int64_t end, double* min_out, double* max_out,
                       double* per_point, int mem_kind) {
    SharedVox* v = c->vox;
    VxSelfParams P{};
    P.c = v->view[c->vox_id];
    P.n = (uint32_t)c->n;
    P.begin = (uint32_t)begin; P.end = (uint32_t)end;
    P.own_zlo = v->hplan.shard.own_zlo; P.own_zhi = v->hplan.shard.own_zhi;
    if (v->sharded) { P.begin = 0; P.end = P.n; }          // split pairs: this rank's layers instead of a slice of the points
    const uint32_t n_total = P.c.n_total, nwords = (n_total + 31u) / 32u;
    const uint32_t far_blocks = (uint32_t)ctx->sm_count * 2u;
    SelfScratch sx{ctx};
    CK(dalloc(ctx, &sx.scratch, (size_t)nwords + 1));       // dupbits
    CK(dalloc(ctx, &sx.vself, (size_t)n_total));
    CK(dalloc(ctx, &sx.und, (size_t)P.c.n + 1));             // [0] count, then the ranks of the undecided voxels
    CK(dalloc(ctx, &sx.mm, ((size_t)P.c.nblk + far_blocks) * 2));
    if (per_point) {
        if (mem_kind == PCCM_DEVICE) sx.d_pp = per_point;
        else { CK(dalloc(ctx, &sx.d_pp, (size_t)c->n)); sx.own_pp = true; }
    }
    CK(dzero(ctx, sx.scratch, ((size_t)nwords + 1) * sizeof(uint32_t), ctx->stream));
    CK(dzero(ctx, sx.und, sizeof(uint32_t), ctx->stream));
    P.undecided = sx.und; P.dupbits = sx.scratch; P.vself = sx.vself; P.minmax = sx.mm; P.per_point = sx.d_pp;
    uint32_t* hund = reinterpret_cast<uint32_t*>(static_cast<double*>(ctx->pinned) + 1024 + 2);
    {
        StageTimer t(ctx, &ctx->tm.knn_ms, 1);
        vx_dupflag_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, ctx->stream>>>(P);
        vx_selfnn_kernel<<<(P.c.nblk + kVxWarps - 1) / kVxWarps, kVxThreads, 0, ctx->stream>>>(P);
        ctx->tm.knn_launches++;
        ctx->tm.total_launches += 2;
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(hund, sx.und, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint32_t nmm = P.c.nblk;
    if (*hund > 0u && v->sharded && !v->full_need) {        // an isolated voxel: its nearest neighbour may lie outside the slab
        const int rc = vox_make_full(ctx, c);
        if (rc) return rc;
        return vox_self_nn(ctx, c, begin, end, min_out, max_out, per_point, mem_kind);
    }
    if (*hund > 0u) {
        const int rc = ensure_pencil(ctx, c);
        if (rc) return rc;
        VxSelfFarParams F{};
        F.s = P;
        F.grid = c->grid; F.recs = static_cast<const uint4*>(c->recs); F.row_start = c->row_start;
        F.minmax_extra = sx.mm + 2 * (size_t)P.c.nblk;
        StageTimer t(ctx, &ctx->tm.knn_ms, 1);
        vx_selffar_kernel<<<far_blocks, 128, 0, ctx->stream>>>(F);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        nmm += far_blocks;
    }
    minmax_finalize_kernel<<<1, 256, 0, ctx->stream>>>(sx.mm, nmm, static_cast<double*>(ctx->dscratch) + 1024);
    ctx->tm.total_launches++;
    if (per_point) {
        vx_selfout_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, ctx->stream>>>(P);
        ctx->tm.total_launches++;
    }
    CK(cudaGetLastError());
    double* hmm = static_cast<double*>(ctx->pinned) + 1024;
    CK(cudaMemcpyAsync(hmm, static_cast<double*>(ctx->dscratch) + 1024, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    int rc = PCCM_OK;
    if (per_point && mem_kind == PCCM_HOST) rc = copy_out(ctx, per_point, sx.d_pp, (size_t)c->n * sizeof(double), PCCM_HOST);
    CK(cudaStreamSynchronize(ctx->stream));
    *min_out = hmm[0];
    *max_out = hmm[1];
    return rc;
}

static int check_range(pccm_ctx* ctx, pccm_cloud* c, int64_t begin, int64_t end) {
    if (c->index_kind < 0) return fail(ctx, PCCM_ERR_STATE, "cloud is not indexed");
    if (begin < 0 || end < begin || end > c->n) return fail(ctx, PCCM_ERR_INVALID, "bad range [%lld, %lld)", (long long)begin, (long long)end);
    return ensure_pencil(ctx, c);
}

extern "C" int pccm_knn_self(pccm_ctx* ctx, pccm_cloud* c, int k, int32_t* idx_out, double* d2_out, int mem_kind) {
    if (!ctx || !c || !idx_out || !d2_out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    if (k < 1) return fail(ctx, PCCM_ERR_INVALID, "k must be >= 1");
    CK(cudaSetDevice(ctx->device));
    int rc = check_range(ctx, c, 0, c->n);
    if (rc) return rc;
    if (c->n == 0) return PCCM_OK;
    const size_t cnt = (size_t)c->n * (size_t)k;
    int32_t* d_idx = idx_out;
    double* d_d2 = d2_out;
    if (mem_kind == PCCM_HOST) { CK(dalloc(ctx, &d_idx, cnt)); CK(dalloc(ctx, &d_d2, cnt)); }
    KnnParams P{};
    P.c = view_of(c); P.begin = c->base; P.end = c->base + (uint32_t)c->n; P.k = k; P.mode = KNN_LIST;
    P.idx_out = d_idx; P.d2_out = d_d2;
    rc = launch_knn(ctx, c, P);
    if (mem_kind == PCCM_HOST) {
        if (!rc) rc = copy_out(ctx, idx_out, d_idx, cnt * sizeof(int32_t), PCCM_HOST);
        if (!rc) rc = copy_out(ctx, d2_out, d_d2, cnt * sizeof(double), PCCM_HOST);
        dfree(ctx, d_idx); dfree(ctx, d_d2);
    }
    return rc;
}

extern "C" int pccm_self_nn_minmax(pccm_ctx* ctx, pccm_cloud* c, int64_t begin, int64_t end,
                                   double* min_out, double* max_out, double* per_point, int mem_kind) {
    if (!ctx || !c || !min_out || !max_out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    if (c->index_kind < 0) return fail(ctx, PCCM_ERR_STATE, "cloud is not indexed");
    if (begin < 0 || end < begin || end > c->n) return fail(ctx, PCCM_ERR_INVALID, "bad range [%lld, %lld)", (long long)begin, (long long)end);
    if (end == begin) { *min_out = INFINITY; *max_out = -INFINITY; return PCCM_OK; }
    { const int rcs = vox_settle(ctx, c); if (rcs) return rcs; }
    if (ctx->use_vox && c->vox && c->n >= 2 && (!per_point || (begin == 0 && end == c->n))) {
        // brick index: the pair search's neighbourhood bits with the voxel's own bit cleared; isolated voxels are
        // finished with the pencil search inside (the pencil index is only built when there are any)
        return vox_self_nn(ctx, c, begin, end, min_out, max_out, per_point, mem_kind);
    }
    int rc = check_range(ctx, c, begin, end);
    if (rc) return rc;
    const uint32_t nblocks = (uint32_t)((end - begin + kKnnThreads - 1) / kKnnThreads);
    double* mm = nullptr;
    double* d_pp = nullptr;
    CK(dalloc(ctx, &mm, (size_t)nblocks * 2));
    if (per_point) { if (mem_kind == PCCM_DEVICE) d_pp = per_point; else CK(dalloc(ctx, &d_pp, (size_t)c->n)); }
    KnnParams P{};
    P.c = view_of(c); P.begin = c->base + (uint32_t)begin; P.end = c->base + (uint32_t)end; P.k = 2; P.mode = KNN_BOUNDARY;
    P.d2_out = d_pp; P.minmax = mm;
    rc = launch_knn(ctx, c, P);
    if (!rc) {
        minmax_finalize_kernel<<<1, 256, 0, ctx->stream>>>(mm, nblocks, static_cast<double*>(ctx->dscratch) + 1024);
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(static_cast<double*>(ctx->pinned) + 1024, static_cast<double*>(ctx->dscratch) + 1024, 2 * sizeof(double),
                           cudaMemcpyDeviceToHost, ctx->stream));
        if (per_point && mem_kind == PCCM_HOST) rc = copy_out(ctx, per_point, d_pp, (size_t)c->n * sizeof(double), PCCM_HOST);
        CK(cudaStreamSynchronize(ctx->stream));
        *min_out = (static_cast<double*>(ctx->pinned) + 1024)[0];
        *max_out = (static_cast<double*>(ctx->pinned) + 1024)[1];
    }
    if (per_point && mem_kind == PCCM_HOST) dfree(ctx, d_pp);
    dfree(ctx, mm);
    return rc;
}

extern "C" int pccm_estimate_normals(pccm_ctx* ctx, pccm_cloud* c, int k, int64_t begin, int64_t end) {
    if (!ctx || !c) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    if (k < 1) return fail(ctx, PCCM_ERR_INVALID, "k must be >= 1");
    CK(cudaSetDevice(ctx->device));
    int rc = check_range(ctx, c, begin, end);
    if (rc) return rc;
    wait_normals(ctx, c);
    if (c->normals_borrowed) { c->normals = nullptr; c->normals_borrowed = false; }
    if (!c->normals) {
        // zero-filled so that ranks estimating disjoint slices can combine buffers by summation
        CK(dalloc(ctx, &c->normals, (size_t)c->n * 3));
        CK(cudaMemsetAsync(c->normals, 0, (size_t)c->n * 3 * sizeof(double), ctx->stream));
    }
    KnnParams P{};
    uint32_t* todo_buf = nullptr;
    P.c = view_of(c); P.begin = c->base + (uint32_t)begin; P.end = c->base + (uint32_t)end; P.k = k; P.mode = KNN_NORMALS;
    P.normals_out = c->normals;
    if (c->index_kind == PCCM_KIND_INT && k <= 255 && ctx->normals_counting && end > begin) {
        // voxelised clouds: counting selection (no per-thread sorted list), then the generic
        // kernel only for the points it flagged (sparse neighbourhoods, fewer than k points)
        const uint32_t cnt = P.end - P.begin;
        const size_t smem = (size_t)kHistBins * kNrmThreads + (size_t)kTieCap * kNrmThreads * 2 * sizeof(uint32_t);
        uint32_t* todo = nullptr;
        CK(dalloc(ctx, &todo, (size_t)cnt + 1));           // [0] = count, [1..] = sorted positions
        CK(cudaMemsetAsync(todo, 0, sizeof(uint32_t), ctx->stream));
        P.todo = todo + 1;
        P.todo_count = todo;
        todo_buf = todo;
        {
            StageTimer t(ctx, &ctx->tm.knn_ms, 1);
            CK(cudaFuncSetAttribute(normals_int_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            normals_int_kernel<<<(cnt + kNrmThreads - 1) / kNrmThreads, kNrmThreads, smem, ctx->stream>>>(P);
            ctx->tm.knn_launches++;
            ctx->tm.total_launches++;
            CK(cudaGetLastError());
        }
        P.mode = KNN_NORMALS_FLAGGED;
    }
    rc = launch_knn(ctx, c, P);
    dfree(ctx, todo_buf);
    if (!rc) c->has_normals = true;
    return rc;
}

// --------------------------------------------------------------------------------------
// minimal-OBB sweep
// --------------------------------------------------------------------------------------
extern "C" int pccm_obb_sweep(pccm_ctx* ctx, const double* hull_vertices, int64_t nv, const double* triangles, int64_t nf,
                              double* vol_out, double* ext_out) {
    if (!ctx || !hull_vertices || !triangles || !vol_out || !ext_out) return fail(ctx, PCCM_ERR_INVALID, "NULL argument");
    if (nv <= 0 || nf <= 0 || nv > 0x7fffffffLL || nf > 0x7fffffffLL) return fail(ctx, PCCM_ERR_INVALID, "bad sizes");
    CK(cudaSetDevice(ctx->device));
    double *dv = nullptr, *dt = nullptr, *dvol = nullptr, *dext = nullptr;
    CK(dalloc(ctx, &dv, (size_t)nv * 3));
    CK(dalloc(ctx, &dt, (size_t)nf * 9));
    CK(dalloc(ctx, &dvol, (size_t)nf));
    CK(dalloc(ctx, &dext, (size_t)nf * 3));
    CK(cudaMemcpyAsync(dv, hull_vertices, (size_t)nv * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dt, triangles, (size_t)nf * 72, cudaMemcpyHostToDevice, ctx->stream));
    obb_sweep_kernel<<<(unsigned)nf, kObbThreads, 0, ctx->stream>>>(dv, (uint32_t)nv, dt, (uint32_t)nf, dvol, dext);
    ctx->tm.total_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(vol_out, dvol, (size_t)nf * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ext_out, dext, (size_t)nf * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx, dv); dfree(ctx, dt); dfree(ctx, dvol); dfree(ctx, dext);
    return PCCM_OK;
}

// --------------------------------------------------------------------------------------
// convex-hull prefilter
// --------------------------------------------------------------------------------------
extern "C" int pccm_cloud_extremes(pccm_ctx* ctx, pccm_cloud* c, const double* dirs, int ndirs, int32_t* idx_out) {
    if (!ctx || !c || !dirs || !idx_out || ndirs < 1) return fail(ctx, PCCM_ERR_INVALID, "bad argument");
    if (c->index_kind < 0) return fail(ctx, PCCM_ERR_STATE, "cloud is not indexed");
    if (c->n == 0) return fail(ctx, PCCM_ERR_INVALID, "empty cloud");
    CK(cudaSetDevice(ctx->device));
    { const int rc = ensure_pencil(ctx, c); if (rc) return rc; }
    const uint32_t n = (uint32_t)c->n;
    const uint32_t nchunks = (n + kExtChunk - 1) / kExtChunk;
    double *ddirs = nullptr, *pval = nullptr;
    uint32_t* pidx = nullptr;
    CK(dalloc(ctx, &ddirs, (size_t)ndirs * 3));
    CK(dalloc(ctx, &pval, (size_t)nchunks * ndirs));
    CK(dalloc(ctx, &pidx, (size_t)nchunks * ndirs));
    CK(cudaMemcpyAsync(ddirs, dirs, (size_t)ndirs * 24, cudaMemcpyHostToDevice, ctx->stream));
    const dim3 grid(nchunks, (ndirs + kExtDirs - 1) / kExtDirs);
    if (c->index_kind == PCCM_KIND_INT) extremes_kernel<KInt><<<grid, kExtThreads, 0, ctx->stream>>>(static_cast<const uint4*>(c->recs), c->base, n, ddirs, ndirs, pval, pidx);
    else if (c->index_kind == PCCM_KIND_F32) extremes_kernel<KF32><<<grid, kExtThreads, 0, ctx->stream>>>(static_cast<const float4*>(c->recs), c->base, n, ddirs, ndirs, pval, pidx);
    else extremes_kernel<KF64><<<grid, kExtThreads, 0, ctx->stream>>>(static_cast<const RecF64*>(c->recs), c->base, n, ddirs, ndirs, pval, pidx);
    ctx->tm.total_launches++;
    CK(cudaGetLastError());
    std::vector<double> hv((size_t)nchunks * ndirs);
    std::vector<uint32_t> hi((size_t)nchunks * ndirs);
    CK(cudaMemcpyAsync(hv.data(), pval, hv.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hi.data(), pidx, hi.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < ndirs; ++d) {
        double bv = -INFINITY;
        uint32_t bi = 0xFFFFFFFFu;
        for (uint32_t k = 0; k < nchunks; ++k) {
            const double v = hv[(size_t)k * ndirs + d];
            const uint32_t i = hi[(size_t)k * ndirs + d];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        idx_out[d] = (int32_t)bi;
    }
    dfree(ctx, ddirs); dfree(ctx, pval); dfree(ctx, pidx);
    return PCCM_OK;
}

extern "C" int pccm_cloud_outside_hull(pccm_ctx* ctx, pccm_cloud* c, const double* planes, int nf, double eps,
                                       int64_t capacity, double* xyz_out, int64_t* count_out) {
    if (!ctx || !c || !planes || !count_out || nf < 1 || capacity < 0 || (capacity && !xyz_out)) return fail(ctx, PCCM_ERR_INVALID, "bad argument");
    if (c->index_kind < 0) return fail(ctx, PCCM_ERR_STATE, "cloud is not indexed");
    if (nf * 32 > 200 * 1024) return fail(ctx, PCCM_ERR_UNSUPPORTED, "too many facets (%d)", nf);
    CK(cudaSetDevice(ctx->device));
    { const int rc = ensure_pencil(ctx, c); if (rc) return rc; }
    const uint32_t n = (uint32_t)c->n;
    double *dpl = nullptr, *dout = nullptr;
    unsigned long long* dcnt = nullptr;
    CK(dalloc(ctx, &dpl, (size_t)nf * 4));
    CK(dalloc(ctx, &dout, (size_t)std::max<int64_t>(capacity, 1) * 3));
    CK(dalloc(ctx, &dcnt, 1));
    CK(cudaMemcpyAsync(dpl, planes, (size_t)nf * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(dcnt, 0, 8, ctx->stream));
    const size_t smem = (size_t)nf * 32;
    const int threads = 256, blocks = (int)((n + threads - 1) / threads);
    if (n) {
        if (c->index_kind == PCCM_KIND_INT) {
            CK(cudaFuncSetAttribute(outside_kernel<KInt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            outside_kernel<KInt><<<blocks, threads, smem, ctx->stream>>>(static_cast<const uint4*>(c->recs), c->base, n, dpl, nf, eps, dout, (unsigned long long)capacity, dcnt);
        } else if (c->index_kind == PCCM_KIND_F32) {
            CK(cudaFuncSetAttribute(outside_kernel<KF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            outside_kernel<KF32><<<blocks, threads, smem, ctx->stream>>>(static_cast<const float4*>(c->recs), c->base, n, dpl, nf, eps, dout, (unsigned long long)capacity, dcnt);
        } else {
            CK(cudaFuncSetAttribute(outside_kernel<KF64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            outside_kernel<KF64><<<blocks, threads, smem, ctx->stream>>>(static_cast<const RecF64*>(c->recs), c->base, n, dpl, nf, eps, dout, (unsigned long long)capacity, dcnt);
        }
        ctx->tm.total_launches++;
        CK(cudaGetLastError());
    }
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, dcnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *count_out = (int64_t)cnt;
    const int64_t take = std::min<int64_t>((int64_t)cnt, capacity);
    if (take > 0) {
        CK(cudaMemcpyAsync(xyz_out, dout, (size_t)take * 24, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    dfree(ctx, dpl); dfree(ctx, dout); dfree(ctx, dcnt);
    return PCCM_OK;
}

// trace builds (-DPCCM_VX_TRACE): per-warp {start ns, end ns, bricks, voxels} of the last search launch; returns the warps copied
extern "C" int pccm_debug_trace(unsigned long long* out, int max_warps) {
#if defined(PCCM_VX_TRACE)
    cudaDeviceSynchronize();
    const int n = std::min(max_warps, (int)pccm::kVxTraceWarps);
    if (cudaMemcpyFromSymbol(out, pccm::g_vx_trace, (size_t)n * 4 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    return n;
#else
    (void)out; (void)max_warps;
    return -1;
#endif
}

// debug builds (-DPCCM_VX_DEBUG): number of range-check violations the brick kernels have counted so far
extern "C" int pccm_debug_errors(void) {
#if defined(PCCM_VX_DEBUG)
    unsigned int v = 0;
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(&v, pccm::g_vx_errors, sizeof v) != cudaSuccess) return -1;
    return (int)v;
#else
    return -1;
#endif
}
