/*
 * pccm.h -- C ABI of libpccm.so, the B200 (sm_100a) replacement for the one
 * data-parallel hot path of aaletov/open-pcc-metric.
 *
 * The reference has no FFI: its only seam is the Python class CloudPair
 * (open_pcc_metric/cloud_pair.py:45-124), which calls Open3D's C++ through
 * pybind once per point.  Each entry point below names the reference code (and the
 * Open3D call under it) that it replaces.  Plain pointers and sizes only; no C++
 * or torch types cross this boundary.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns PCCM_OK (0) or a negative pccm_status; the message
 *    for the last failure is pccm_last_error(ctx) (ctx may be NULL);
 *  - all device work is ordered on the context's CUDA stream; functions that
 *    fill HOST outputs synchronise that stream before returning;
 *  - a pccm_ctx is not thread safe: one context per (thread, device);
 *  - caller owns every buffer it passes; pageable HOST inputs are copied before return,
 *    PINNED host inputs are copied asynchronously (cudaMemcpyAsync): keep them unchanged
 *    until a call that returns results or pccm_ctx_synchronize;
 *    DEVICE coordinates / colours are read until the index is built; packed float64
 *    DEVICE normals given to pccm_cloud_create are used in place and must outlive
 *    the cloud;
 *  - there is NO CPU fallback: without a CUDA device pccm_ctx_create fails.
 */
#ifndef PCCM_H_
#define PCCM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCCM_VERSION 100 /* 0.1.0 */

typedef enum pccm_status {
    PCCM_OK = 0,
    PCCM_ERR_INVALID = -1,   /* bad argument */
    PCCM_ERR_CUDA = -2,      /* CUDA runtime failure (message holds cudaGetErrorString) */
    PCCM_ERR_STATE = -3,     /* call order (index not built, normals missing, ...) */
    PCCM_ERR_INDEX = -4,     /* the reference would raise IndexError here (quirk Q1, empty search cloud) */
    PCCM_ERR_NONFINITE = -5, /* NaN / Inf coordinate */
    PCCM_ERR_UNSUPPORTED = -6
} pccm_status;

typedef enum pccm_dtype { PCCM_F64 = 0, PCCM_F32 = 1, PCCM_I32 = 2, PCCM_U16 = 3, PCCM_U8 = 4 } pccm_dtype;
typedef enum pccm_mem { PCCM_HOST = 0, PCCM_DEVICE = 1 } pccm_mem;
/* coordinate kind of an index: INT = integers in [0, 32767] (exact uint32 d2),
 * F32 = float32-representable (stored as float, evaluated in float64),
 * F64 = arbitrary float64. */
typedef enum pccm_kind { PCCM_KIND_AUTO = -1, PCCM_KIND_INT = 0, PCCM_KIND_F32 = 1, PCCM_KIND_F64 = 2 } pccm_kind;

enum {
    PCCM_EVAL_D2 = 1,        /* point-to-plane error (metric.py:124-179, point_to_plane=True) */
    PCCM_EVAL_COLOR = 2,     /* colour error on matched points (metric.py:302-333, 389-426) */
    PCCM_EVAL_PERPOINT = 4,  /* keep per-point idx / d2 for pccm_pair_get */
    PCCM_EVAL_TIE_AVERAGE = 8 /* opt-in, beyond the reference (which keeps whichever tied neighbour its KD-tree visited first,
                                cloud_pair.py:22-23): plane error and colour of a query are AVERAGED over every point of the search
                                cloud at the minimal distance (MPEG pc_error style) -- the plane error of each tied point with that
                                point's own normal, the colour as the mean colour of the tied points.  D1, maxima of D1 and the
                                per-point outputs are unaffected.  Runs on the pencil index. */
};
enum {
    PCCM_NORMALS_BY_QUERY_INDEX = 0, /* reference behaviour, metric.py:130 + :148-152 (quirk Q1) */
    PCCM_NORMALS_BY_NEIGHBOUR = 1    /* normal of the matched point (MPEG pc_error style, opt-in) */
};
enum { PCCM_GET_IDX = 0, PCCM_GET_D2 = 1 };

typedef struct pccm_ctx pccm_ctx;
typedef struct pccm_cloud pccm_cloud;

typedef struct pccm_cloud_info {
    int64_t n;
    int32_t data_kind;    /* pccm_kind the coordinates allow */
    int32_t index_kind;   /* pccm_kind of the built index, -1 if none */
    int32_t has_colors, colors_u8, has_normals, indexed;
    int32_t ny, nz;       /* pencil table dimensions */
    int32_t sharded;      /* 0, or the number of ranks the pair this cloud was indexed with is split over (pccm_ctx_set_shard) */
    int32_t reserved;
    double cell_size;
    double aabb_min[3], aabb_max[3];
} pccm_cloud_info;

/* One direction of a symmetric evaluation = one pass of get_neighbour_cloud
 * (cloud_pair.py:10-42) plus the reductions metric.py applies to its outputs. */
typedef struct pccm_dir_result {
    int64_t n;              /* queries reduced by THIS call (a rank's share) */
    int64_t n_total;        /* size of the query cloud */
    uint64_t sum_d1_u64;    /* exact sum of squared NN distances when d1_exact_int */
    int32_t d1_exact_int;   /* 1 for PCCM_KIND_INT pairs */
    int32_t d2_valid;       /* 0 when D2 was requested but the reference would raise IndexError in this
                               direction (BY_QUERY_INDEX with a shorter search cloud, quirk Q1) */
    double sum_d1;          /* sum of squared NN distances   (GeoMSE numerator, metric.py:226-228) */
    double max_d1;          /* max of squared NN distances   (GeoHausdorffDistance, metric.py:366) */
    double sum_d2;          /* sum of squared plane errors   (point_to_plane=True) */
    double max_d2;
    double color_sum[3];    /* sum of (T c_q - T c_nn)^2 per channel     (ColorMSE, metric.py:333) */
    double color_max[3];    /* max of (scale * (T c_q - T c_nn))^2       (ColorHausdorffDistance, metric.py:421-426) */
} pccm_dir_result;

typedef struct pccm_pair_result {
    pccm_dir_result dir[2]; /* [0] = left: iterate A, search B; [1] = right */
} pccm_pair_result;

/* Device time per stage, accumulated between pccm_ctx_reset_timings calls when
 * profiling is enabled (CUDA events on the context stream). */
typedef struct pccm_timings {
    double upload_ms, stats_ms, keys_ms, sort_ms, table_ms, reorder_ms;
    double query_ms, finalize_ms, knn_ms;
    double vox_build_ms;       /* occupancy-brick index of integer pairs */
    double vox_tail_ms;        /* brick-ring search of the voxels the search kernel left undecided (+ the pencil round, when one is needed) */
    double vox_search_ms;      /* staged bit-scan search kernel of the brick path (query_ms covers the whole query stage) */
    double vox_epilogue_ms;    /* per-point epilogue kernel of the brick path (D1 / D2 / colour + reduction records) */
    int64_t query_launches, knn_launches;
    int64_t total_launches;    /* kernels of this library (hand-written, sm_100a) */
    int64_t library_launches;  /* kernels of a library (none: every sort and scan is hand-written; the field is kept and stays 0) */
    int64_t vox_undecided;     /* queries of the last brick-path evaluation that needed the general search */
    int64_t vox_far;           /* ... of which the pencil search had to finish (nearest point tens of voxels away) */
    int64_t vox_tail;          /* points that share a voxel with a smaller index (they reuse that voxel's search), last evaluation */
} pccm_timings;

int pccm_version(void);
const char* pccm_last_error(const pccm_ctx* ctx);

/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for a private stream */
int pccm_ctx_create(int device, void* stream, pccm_ctx** out);
int pccm_ctx_destroy(pccm_ctx* ctx);
int pccm_ctx_synchronize(pccm_ctx* ctx);
/* level 0 = off, 1 = time the query stage (one event pair) / the k-NN kernels only, 2 = time every stage and every
 * kernel of the query stage (the event pairs keep those kernels from overlapping their launches) */
int pccm_ctx_set_profiling(pccm_ctx* ctx, int level);
int pccm_ctx_reset_timings(pccm_ctx* ctx);
/* One pair over several GPUs (one context per GPU, the same calls on every rank, both clouds given to every rank):
 * integer pairs built from now on are split by slabs of z that hold equal numbers of points -- this rank indexes its
 * slab (+ a halo of two bricks) and evaluates the queries of its slab only; pccm_pair_eval / pccm_self_nn_minmax then
 * return this rank's partial sums (pccm_dir_result.n counts its points) whatever rank / range they are called with,
 * and the caller adds the ranks up (sums) / takes the extremes (max, min).  The cuts are computed on the device from a
 * histogram the statistics pass takes along; nothing is exchanged inside the library.  Other pairs (float kinds) keep
 * the rank / world arguments of pccm_pair_eval.  Set before the clouds are created; (0, 1) switches it off. */
int pccm_ctx_set_shard(pccm_ctx* ctx, int rank, int world);
int pccm_ctx_get_timings(pccm_ctx* ctx, pccm_timings* out);

/* Replaces building an o3d.geometry.PointCloud + KDTreeFlann input (cloud_pair.py:59,65).
 * xyz: n rows of 3 values, row stride in bytes (0 = packed).  rgb: NULL, PCCM_F64 in
 * [0,1] (Open3D convention) or PCCM_U8.  normals: NULL, PCCM_F64 or PCCM_F32. */
int pccm_cloud_create(pccm_ctx* ctx, const void* xyz, int xyz_dtype, int64_t n, int64_t xyz_stride,
                      const void* rgb, int rgb_dtype, int64_t rgb_stride,
                      const void* normals, int nrm_dtype, int64_t nrm_stride,
                      int mem_kind, pccm_cloud** out);
/* Colours and / or normals for a cloud created without them (NULL = leave as is); same formats and
 * ownership rules as pccm_cloud_create.  Lets a caller submit the COORDINATES of both clouds of a pair
 * first: the index build needs nothing else, and attributes uploaded afterwards travel on the copy
 * stream beside it (they are first read by the epilogue).  Colours must be attached before
 * pccm_cloud_build_index (a brick-indexed pair accepts them until its first colour evaluation). */
int pccm_cloud_attach(pccm_ctx* ctx, pccm_cloud* cloud, const void* rgb, int rgb_dtype, int64_t rgb_stride,
                      const void* normals, int nrm_dtype, int64_t nrm_stride, int mem_kind);
int pccm_cloud_destroy(pccm_ctx* ctx, pccm_cloud* cloud);
int pccm_cloud_info_get(pccm_ctx* ctx, pccm_cloud* cloud, pccm_cloud_info* out);

/* Replaces o3d.geometry.KDTreeFlann(cloud) (cloud_pair.py:65): builds the pencil-grid
 * index.  cell_size 0 = automatic; force_kind = PCCM_KIND_AUTO or a kind >= data_kind
 * (both clouds of a pair must share one kind). */
int pccm_cloud_build_index(pccm_ctx* ctx, pccm_cloud* cloud, double cell_size, int force_kind);

/* Builds the indices of BOTH clouds of a pair (the two KDTreeFlann builds of cloud_pair.py:65) in joint launches.
 * Integer pairs (every coordinate an integer in [0, 32767]: voxelised content) get the occupancy-brick index -- a
 * directory of occupied 32 x 8 x 8-voxel bricks, one occupancy word per brick row, one record per distinct voxel --
 * enqueued WITHOUT a host synchronisation (the outcome is judged at the result read-back of pccm_pair_eval); other
 * pairs get the pencil-grid index (counting sort by row + per-row sorting networks, no library sort).  Results are the
 * same as with two pccm_cloud_build_index calls with the common kind. */
int pccm_pair_build_index(pccm_ctx* ctx, pccm_cloud* a, pccm_cloud* b, double cell_size, int force_kind);

int pccm_cloud_set_normals(pccm_ctx* ctx, pccm_cloud* cloud, const void* normals, int nrm_dtype,
                           int64_t nrm_stride, int mem_kind);
int pccm_cloud_get_normals(pccm_ctx* ctx, pccm_cloud* cloud, double* out, int mem_kind);

/* Replaces PointCloud.estimate_normals() with Open3D defaults (cloud_pair.py:61-64):
 * k-NN (self included) covariance + analytic eigenvector of the smallest eigenvalue.
 * [begin, end) is a range of the cloud's SORTED order (0, n for the whole cloud);
 * results land in the cloud's normal buffer at original indices. */
int pccm_estimate_normals(pccm_ctx* ctx, pccm_cloud* cloud, int k, int64_t begin, int64_t end);

/* k nearest neighbours of every point in its own cloud, rows sorted by (d2, index),
 * self included (what KDTreeFlann.search_knn_vector_3d(p, k) returns for p in the cloud).
 * idx_out[n*k] (-1 padded when n < k), d2_out[n*k] (inf padded). */
int pccm_knn_self(pccm_ctx* ctx, pccm_cloud* cloud, int k, int32_t* idx_out, double* d2_out, int mem_kind);

/* Replaces PointCloud.compute_nearest_neighbor_distance() + np.min/np.max
 * (cloud_pair.py:108-109, metric.py:182-188): distance of each point to its nearest
 * OTHER point.  per_point may be NULL.  [begin, end) as above (on the brick index: the same
 * share of the cloud's occupied voxels; any partition of [0, n) into ranges covers every point once). */
int pccm_self_nn_minmax(pccm_ctx* ctx, pccm_cloud* cloud, int64_t begin, int64_t end,
                        double* min_out, double* max_out, double* per_point, int mem_kind);

/* Replaces one get_neighbour_cloud pass (cloud_pair.py:10-42): per query point the
 * original index and squared distance of its nearest neighbour in `search`.
 * Outputs are in the query cloud's ORIGINAL order; either may be NULL. */
int pccm_nn(pccm_ctx* ctx, pccm_cloud* query, pccm_cloud* search, int32_t* idx_out, double* d2_out, int mem_kind);

/* Replaces CloudPair.__init__'s two passes (cloud_pair.py:67-78) fused with the
 * per-point arithmetic and reductions of metric.py (see pccm_dir_result).
 * color_matrix: row-major 3x3 applied to both colours (identity for "rgb");
 * color_scale: factor inside color_max (255 for rgb, quirk Q6; else 1).
 * rank/world: this call reduces the rank-th of `world` equal contiguous slices of
 * each query cloud's sorted order (0, 1 = everything; integer pairs on the brick index:
 * of its occupied voxels -- points that share a voxel stay together, pccm_dir_result.n
 * counts the points actually reduced); partial results from all ranks add up (sums) /
 * max (maxima). */
int pccm_pair_eval(pccm_ctx* ctx, pccm_cloud* a, pccm_cloud* b, uint32_t flags,
                   const double* color_matrix, double color_scale, int normals_mode,
                   int rank, int world, pccm_pair_result* out);

/* Per-point products of the last pccm_pair_eval called with PCCM_EVAL_PERPOINT:
 * which = PCCM_GET_IDX (int32[n]) or PCCM_GET_D2 (double[n]); direction 0 = left. */
int pccm_pair_get(pccm_ctx* ctx, int which, int direction, void* out, int mem_kind);

/* Facet sweep of Open3D's OrientedBoundingBox::CreateFromPointsMinimal (the PSNR peak,
 * cloud_pair.py:111-112): for each of nf hull triangles (9 doubles: a, b, c) the extent and
 * volume of the axis-aligned box of the nv hull vertices in the triangle's frame.  The convex
 * hull itself is computed by the caller (Qhull on the host, as Open3D does).  HOST buffers. */
int pccm_obb_sweep(pccm_ctx* ctx, const double* hull_vertices, int64_t nv, const double* triangles, int64_t nf,
                   double* vol_out, double* ext_out);

/* Convex-hull prefilter for that peak.  pccm_cloud_extremes: original index of the arg-max of
 * dirs[d] . p for each of ndirs directions (row-major [ndirs][3]).  pccm_cloud_outside_hull:
 * the points that are NOT strictly inside the convex polytope { p : n_f . p + d_f < -eps for all
 * f } given as planes[nf][4] (e.g. the hull of the extremes) -- a superset of the hull vertices
 * of the whole cloud.  count_out receives the number found; at most `capacity` are written. */
int pccm_cloud_extremes(pccm_ctx* ctx, pccm_cloud* cloud, const double* dirs, int ndirs, int32_t* idx_out);
int pccm_cloud_outside_hull(pccm_ctx* ctx, pccm_cloud* cloud, const double* planes, int nf, double eps,
                            int64_t capacity, double* xyz_out, int64_t* count_out);

#ifdef __cplusplus
}
#endif
#endif /* PCCM_H_ */
