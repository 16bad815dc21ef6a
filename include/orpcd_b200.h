/*
 * orpcd_b200.h -- C ABI of the B200 (sm_100a) registration hot path.
 *
 * The reference (dpolimeni/Multi-scale-pointcloud-registration, package or_pcd)
 * has no FFI of its own: its plugin API is Python abstract classes
 * (Optimizer/iOptimizer.py:7-21, Preprocessor/iProcessBlock.py:6-12) whose
 * concrete classes call Open3D 0.18.0.  Each entry point below replaces one of
 * those Open3D call sites; the Python classes under
 * multi-scale-pointcloud-registration_b200/or_pcd_b200/ bind them with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller; the
 *     library never allocates or frees across this boundary (workspace sizes
 *     come from the *_workspace_bytes() queries);
 *   - all work is enqueued on `stream` (a cudaStream_t); calls return after
 *     enqueueing unless stated otherwise;
 *   - return value: ORPCD_OK or a negative ORPCD_ERR_* code; no exceptions;
 *   - point clouds are float64, C-contiguous (N,3) as numpy hands them to the
 *     reference; symmetric 3x3 matrices are stored as 6 doubles
 *     (xx, xy, xz, yy, yz, zz); 4x4 transforms are 16 doubles, row-major, in
 *     Open3D's column convention (x' = R x + t);
 *   - neighbour ties break towards the lowest index (BASELINE.json north_star).
 *   - re-entrant per (device, stream); no global mutable state.
 */
#ifndef ORPCD_B200_H
#define ORPCD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *orpcd_stream_t; /* cudaStream_t */

#define ORPCD_OK 0
#define ORPCD_ERR_INVALID (-1)   /* bad argument                          */
#define ORPCD_ERR_WORKSPACE (-2) /* workspace too small / misaligned      */
#define ORPCD_ERR_CUDA (-3)      /* a CUDA runtime call or launch failed  */

#define ORPCD_MAX_KNN 512

/* Library / build identification. */
int orpcd_version(void);
/* Compute capability the kernels were compiled for (100 for sm_100a). */
int orpcd_compiled_arch(void);
const char *orpcd_error_string(int code);

/* ------------------------------------------------------------------------
 * Exact k-nearest neighbours, float64 distances ((dx*dx + dy*dy) + dz*dz,
 * no FMA contraction), results sorted by (distance, index).
 * Replaces open3d.geometry.KDTreeFlann.SearchKNN as used inside
 *   PointCloud.remove_statistical_outlier   (Preprocessor/Outliers/sor.py:69-71)
 *   PointCloud.estimate_normals             (Optimizer/fastGlobalOptimizer.py:118-127)
 *   registration_generalized_icp            (Optimizer/generalizedICP.py:59-70)
 * d_points : [batch][n][3]  (batch stride = points_stride doubles)
 * d_queries: [batch][m][3]  (batch stride = queries_stride doubles)
 * d_idx    : [batch][m][k] int32, -1 where fewer than k points exist
 * d_d2     : [batch][m][k] float64, +inf where fewer than k points exist
 * 1 <= k <= ORPCD_MAX_KNN.
 * d_workspace (orpcd_knn_exact_workspace_bytes, may be 0 bytes / NULL): with few
 * queries the point range of each query is split over several blocks and the
 * per-split lists are merged (exact: the (d2, index) order is total); without
 * a workspace the search runs unsplit.
 * ---------------------------------------------------------------------- */
size_t orpcd_knn_exact_workspace_bytes(int n, int m, int batch, int k);
int orpcd_knn_exact(const double *d_points, int n, long long points_stride,
                    const double *d_queries, int m, long long queries_stride,
                    int batch, int k, int *d_idx, double *d_d2, void *d_workspace,
                    size_t workspace_bytes, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * k-NN of a moved copy of a cloud, re-ranked from the base cloud's lists.
 * The Aligner registers `attempts` rigid images of one source cloud
 * (Aligner/Aligner.py:179-185) and Open3D rebuilds the 20-NN graph of each;
 * here the kc > k nearest neighbours of the BASE cloud (orpcd_knn_exact once)
 * are re-ranked with exact float64 distances of the moved coordinates.  A
 * point whose k-th re-ranked distance is not strictly below
 * d_factor[b] * d_cand_bound[i] (the least a non-candidate can reach) gets an
 * exact scan, so the output always equals orpcd_knn_exact on the moved cloud.
 * d_cand_idx  : [n][kc] neighbour lists of the base cloud (self query)
 * d_cand_bound: [n] base d2 of the last candidate (orpcd_knn_candidate_bound)
 * d_factor    : [batch] lower bound of moved_d2 / base_d2 (rigid: 1 - 1e-9,
 *               per-axis scale s: min(s)^2 * (1 - 1e-9))
 * ---------------------------------------------------------------------- */
int orpcd_knn_candidate_bound(const int *d_cand_idx, const double *d_cand_d2, int n, int kc,
                              double *d_bound, orpcd_stream_t stream);
size_t orpcd_knn_rerank_workspace_bytes(int n, int batch);
int orpcd_knn_rerank(const int *d_cand_idx, const double *d_cand_bound, int kc,
                     const double *d_points, int n, long long points_stride, int batch,
                     const double *d_factor, int k, int *d_idx, double *d_d2, void *d_workspace,
                     size_t workspace_bytes, const int *d_order, const double *d_cand_d2,
                     orpcd_stream_t stream);
/* d_order (may be NULL) [n]: order in which the points are processed, e.g. the base cloud's spatial order
 * (orpcd_spatial_order): neighbouring threads then gather the same candidates and most gathers hit L1.
 * d_cand_d2 (may be NULL) [n][kc]: the base distances of the candidates (ascending, as orpcd_knn_exact returned
 * them): the scan of a point stops once the next candidate's lower bound factor x base d2 exceeds the k-th moved
 * distance.  Neither changes the results. */
size_t orpcd_spatial_order_workspace_bytes(int n);
int orpcd_spatial_order(const double *d_points, int n, void *d_workspace, size_t workspace_bytes,
                        int *d_order, orpcd_stream_t stream);   /* d_order[pos] = index of the point at Hilbert position pos */

/* ------------------------------------------------------------------------
 * Statistical outlier removal mask.
 * Replaces PointCloud.remove_statistical_outlier(nb_neighbors, std_ratio)
 * (Preprocessor/Outliers/sor.py:69-71).
 * d_mask  : [n] uint8 (1 = inlier)
 * d_avg   : [n] float64 mean neighbour distance per point (may be NULL)
 * d_stats : [4] float64 {cloud_mean, std_dev, threshold, n_inliers}
 * ---------------------------------------------------------------------- */
size_t orpcd_sor_workspace_bytes(int n, int nb_neighbors);
int orpcd_sor_mask(const double *d_points, int n, int nb_neighbors, double std_ratio,
                   void *d_workspace, size_t workspace_bytes, unsigned char *d_mask,
                   double *d_avg, double *d_stats, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Per-point covariance -> normal (closed-form symmetric eigen-solve) ->
 * GICP covariance R diag(eps,1,1) R^T, for `batch` clouds of n points.
 * Replaces PointCloud.estimate_normals(KNN(knn) | Hybrid(radius, knn)) and
 * InitializePointCloudForGeneralizedICP inside registration_generalized_icp
 * (Optimizer/generalizedICP.py:59-70; fastGlobalOptimizer.py:118-127).
 * radius <= 0 selects the pure k-NN search.  Any output may be NULL.
 * d_normals  : [batch][n][3]
 * d_cov6     : [batch][n][6]  neighbourhood covariance
 * d_gicp_cov6: [batch][n][6]  (requires gicp_eps > 0)
 * d_gicp_axis: [batch][n][3]  unit axis a with d_gicp_cov6 = I - (1-eps) a a^T
 * d_nn_idx   : [batch][n][knn] int32 neighbour lists actually used (-1 padded)
 * orpcd_normals_cov_from_knn takes neighbour lists computed elsewhere
 * (orpcd_knn_exact / orpcd_knn_rerank) instead of searching itself.
 * ---------------------------------------------------------------------- */
size_t orpcd_normals_workspace_bytes(int n, int batch, int knn);
int orpcd_normals_cov(const double *d_points, int n, long long points_stride, int batch,
                      int knn, double radius, double gicp_eps, void *d_workspace,
                      size_t workspace_bytes, double *d_normals, double *d_cov6,
                      double *d_gicp_cov6, double *d_gicp_axis, int *d_nn_idx,
                      orpcd_stream_t stream);
int orpcd_normals_cov_from_knn(const double *d_points, int n, long long points_stride, int batch,
                               int knn, const int *d_knn_idx, const double *d_knn_d2, double radius,
                               double gicp_eps, double *d_normals, double *d_cov6,
                               double *d_gicp_cov6, double *d_gicp_axis, int *d_nn_idx,
                               const int *d_order /* may be NULL, see orpcd_knn_rerank */, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * target * (sx, sy, sz) for a batch of scale candidates
 * (Aligner/Aligner.py:221-222, `target * new_scale_factors`).
 * d_scales: [n_scales][3]; d_out: [n_scales][n][3].
 * ---------------------------------------------------------------------- */
int orpcd_apply_scale_batch(const double *d_points, int n, const double *d_scales,
                            int n_scales, double *d_out, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * source @ R_u + t_u for a batch of random starts (row-vector convention of
 * Aligner/Aligner.py:183-185, `np.dot(source, rotation) + translation`).
 * d_rot: [n_units][3][3] row-major; d_trans: [n_units][3]; d_out: [n_units][n][3].
 * Same arithmetic as numpy's expression (OpenBLAS dgemm: x R0j, then two fused
 * multiply-adds over k, then one rounded add of t_j), so the clouds are bit-identical
 * to the host's; the Aligner verifies that once before relying on it.
 * ---------------------------------------------------------------------- */
int orpcd_apply_rigid_batch(const double *d_points, int n, const double *d_rot,
                            const double *d_trans, int n_units, double *d_out,
                            orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Batched Generalized-ICP: n_units independent inner registrations
 * (one per multi-start x scale candidate) advanced together, one kernel
 * launch per Gauss-Newton iteration for all of them (update, chunk-culled
 * FP32 1-NN sweep with certified float64 re-rank, Jacobian sums, 6x6 solve).
 * Replaces registration_generalized_icp(source, target, max_dist, I,
 * TransformationEstimationForGeneralizedICP(), ICPConvergenceCriteria(
 * max_iteration)) at Optimizer/generalizedICP.py:59-70, called once per
 * start from Aligner/Aligner.py:188-190.
 *
 * Unit u registers source set d_unit_src[u] onto target set d_unit_tgt[u].
 * ---------------------------------------------------------------------- */
typedef struct {
    int n_units;
    int n_src;            /* points per source cloud              */
    int n_tgt;            /* points per target cloud              */
    int n_src_sets;
    int n_tgt_sets;
    int max_iterations;   /* ICPConvergenceCriteria.max_iteration  */
    int target_splits;    /* dense sweeps only: splits of the target table, 0 = automatic */
    int poll_interval;    /* launch-per-round modes (sweep_mode != 0) only: >0: every that many
                             iterations the call waits on the stream and stops early once every
                             unit is done.  The default persistent kernel ends by itself. */
    int profile;          /* !=0 (diagnostic): CUDA events around the dominant kernel (persistent
                             kernel; launch-per-round modes: every round / sweep launch) on the
                             stream; the call then synchronises and reads the device counters
                             back (fills the * entries of h_profile) */
    int cov_mode;         /* 0: d_*_cov hold 3x3 covariances, 6 doubles per point, and
                                (Cs+Ct)^-1/2 comes from a Jacobi eigen-solve;
                             1: d_*_cov hold the unit AXIS a of each GICP plane
                                covariance C = I - (1-eps) a a^T (3 doubles per point,
                                d_gicp_axis of orpcd_normals_cov) and (Cs+Ct)^-1 is
                                evaluated in closed form                          */
    double max_correspondence_distance;
    double relative_fitness; /* 1e-6 in the reference's call       */
    double relative_rmse;    /* 1e-6                               */
    double gicp_epsilon;     /* eps of diag(eps,1,1); used by cov_mode 1 (Open3D: 1e-3) */
    int sweep_q;             /* tuning (modes 1..4): queries per thread (4|8), 0 = auto       */
    int sweep_tj;            /* tuning (modes 1..4): targets per inner step (2|4|8), 0 = auto */
    int sweep_smem_pad;      /* tuning (modes 1..3): unused dynamic smem bytes, 0             */
    int sweep_mode;          /* 0 = default: ONE persistent launch for all rounds of all units --
                                warps take (unit, patch) items from a work queue, each item is a
                                fused round (update, chunk-culled sweep, certificate, exact
                                re-rank, Jacobian sums), the warp closing a unit's round solves
                                and enqueues the unit's next round;
                                5 = the same fused round, one launch per round (lock step);
                                4 = culled sweep with separate update / resolve / fix /
                                reduce kernels; 1..3 = dense (brute-force) sweep + separate
                                kernels: 1 all packed FFMA2, 2 FFMA2 + scalar FFMA, 3 scalar */
    int src_rigid_copies;    /* != 0: every source set is a rigid image of ONE cloud with the same point
                                order (the Aligner's random starts, Aligner/Aligner.py:179-185): the spatial
                                (Hilbert) order of set 0 is reused for all of them instead of sorting each */
    int half_patch_from_round; /* fused rounds: a unit's work items are 64-point patches below this round and
                                32-point patches from it on (shorter per-round latency for the stragglers of a
                                batch); 0 = default (round 32 for clouds of at most 16384 points, never for larger
                                ones), < 0 = never.  Depends on the round index and the cloud size only: results are
                                deterministic and independent of the other units of the batch */
} orpcd_gicp_params;

size_t orpcd_gicp_batch_workspace_bytes(const orpcd_gicp_params *p);

/* d_src_points [n_src_sets][n_src][3], d_src_cov [n_src_sets][n_src][6 | 3],
 * d_tgt_points [n_tgt_sets][n_tgt][3], d_tgt_cov [n_tgt_sets][n_tgt][6 | 3]  (see cov_mode),
 * d_out_T     [n_units][16],
 * d_out_stats [n_units][4]  {fitness, inlier_rmse, iterations, n_correspondences},
 * d_out_status[n_units]     0 = ok, 1 = no correspondence (inlier_rmse == 0:
 *                           the reference raises ValueError, generalizedICP.py:77-81),
 * d_trace (may be NULL) [n_units][max_iterations+1][2] (fitness, rmse) per Evaluate,
 * h_profile (HOST, may be NULL) [8]: {kernels launched, ms inside the round kernel
 *   (dense modes: the 1-NN sweep kernel)*, (query,target) pairs evaluated*, rounds
 *   issued (persistent kernel: rounds that ran*), queries that needed the exact float64 search*, target splits, launches
 *   with work*, unit-rounds*};
 *   entries marked * are filled only when p->profile != 0.                    */
int orpcd_gicp_batch(const orpcd_gicp_params *p, const double *d_src_points,
                     const double *d_src_cov, const double *d_tgt_points,
                     const double *d_tgt_cov, const int *d_unit_src, const int *d_unit_tgt,
                     void *d_workspace, size_t workspace_bytes, double *d_out_T,
                     double *d_out_stats, int *d_out_status, double *d_trace,
                     double *h_profile, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Batched point-to-point / point-to-plane ICP with Open3D's loop and stopping
 * rule (|d fitness| < relative_fitness and |d rmse| < relative_rmse, or
 * max_iterations).  Replaces registration_icp(source, target, threshold, init,
 * TransformationEstimationPointToPoint() | PointToPlane(), ICPConvergenceCriteria(
 * max_iteration)) at Aligner/Aligner.py:353-360 (Aligner.refine_registration).
 * point_to_plane == 0: Umeyama/Kabsch step from the correspondences;
 * point_to_plane != 0: Gauss-Newton step on (vs - vt).nt with d_tgt_normals
 *                      [n_tgt_sets][n_tgt][3] (Open3D raises without target normals).
 * d_init_T (may be NULL) [n_units][16] initial transforms, Open3D column convention.
 * Only the size, distance and convergence fields of `p` are read; workspace as
 * orpcd_gicp_batch_workspace_bytes(p) with p->cov_mode == 1.  Outputs as orpcd_gicp_batch.
 * ---------------------------------------------------------------------- */
int orpcd_icp_batch(const orpcd_gicp_params *p, int point_to_plane, const double *d_src_points,
                    const double *d_tgt_points, const double *d_tgt_normals, const int *d_unit_src,
                    const int *d_unit_tgt, const double *d_init_T, void *d_workspace,
                    size_t workspace_bytes, double *d_out_T, double *d_out_stats, int *d_out_status,
                    orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * EvaluateRegistration for a batch of units: source set unit_src[u] moved by
 * d_T[u] (NULL = identity), nearest target of every source point within
 * max_correspondence_distance -> fitness, inlier_rmse, correspondence count.
 * Replaces registration::EvaluateRegistration as used at the end of
 * registration_fgr_based_on_feature_matching (Optimizer/fastGlobalOptimizer.py:166-174)
 * and o3d...evaluate_registration.  Uses the same exact 1-NN kernels as
 * orpcd_gicp_batch (FP32 certified filter + float64 re-rank).  Only the size and
 * distance fields of `p` are read; workspace as orpcd_gicp_batch_workspace_bytes(p)
 * with p->cov_mode == 1.
 * d_out_stats [n_units][4] {fitness, inlier_rmse, 0, n_correspondences};
 * d_out_status[n_units] 1 = empty correspondence set (the reference raises Warning,
 * fastGlobalOptimizer.py:181-188); d_out_corr (may be NULL) [n_units][n_src] target
 * index per source point or -1.
 * ---------------------------------------------------------------------- */
int orpcd_evaluate_batch(const orpcd_gicp_params *p, const double *d_src_points,
                         const double *d_tgt_points, const int *d_unit_src, const int *d_unit_tgt,
                         const double *d_T, void *d_workspace, size_t workspace_bytes,
                         double *d_out_stats, int *d_out_status, int *d_out_corr,
                         orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * FPFH descriptors, 33 float64 per point, point-major [batch][n][33] (the memory
 * order of Open3D's 33 x N column-major Feature::data_).
 * Replaces compute_fpfh_feature(cloud, KDTreeSearchParamHybrid(radius, knn))
 * (Optimizer/fastGlobalOptimizer.py:130-143).  Neighbour lists come from
 * orpcd_knn_exact / orpcd_knn_rerank (self query, width list_k >= knn, sorted);
 * radius <= 0 selects the pure k-NN search.  d_spfh is the intermediate SPFH
 * table, also an output: [batch][n][3][12] float64, each of the three 11-bin
 * histograms followed by a zero (288-byte rows, 32-byte aligned histograms:
 * the FPFH pass gathers them with 16-byte loads); 32-byte aligned base.
 * d_fpfh [batch][n][33].
 * ---------------------------------------------------------------------- */
int orpcd_fpfh(const double *d_points, const double *d_normals, int n, long long points_stride,
               int batch, const int *d_nn_idx, const double *d_nn_d2, int list_k, int knn,
               double radius, double *d_spfh, double *d_fpfh, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Nearest neighbour in 33-D feature space (the kd-tree searches of FGR's
 * InitialMatching, FastGlobalRegistration.cpp; fastGlobalOptimizer.py:166-174).
 * Features are first packed (orpcd_feature_pack) into TF32 hi/lo operand tiles for
 * the tcgen05 tensor-core filter; orpcd_feature_match then returns, for pair p and
 * every row of query set d_pair_q[p], the exact float64 nearest row of database set
 * d_pair_db[p] (nanoflann summation order, ties -> lowest index).
 * mode 0: tensor-core filter (the eight smallest 32-row chunk minima + the ninth as a bound)
 * + certified float64 re-score; mode 1: float64 scan only.
 * Distances are translation invariant, so the filter runs on features centred on
 * d_center[33] (NULL = no centring; e.g. the mean feature of one set from
 * orpcd_feature_center): smaller norms give a tighter certificate.  The same vector
 * must be given to the pack of both sides and to the match.  Exact distances are
 * always taken on the original values.
 * d_packed_* : orpcd_feature_pack_bytes(n) bytes per set, 128-byte aligned
 * d_set_maxnorm: [n_sets] float, max |f|^2 of each set (written by the pack)
 * d_row_mask (may be NULL) [n_pairs][n_query] bytes: rows with a zero are not resolved (d_nn = -1).  FGR's cross
 * check only reads the reverse table at the images of the forward one (ji[ij[i]] == i), so the reverse match is
 * asked for those rows only;
 * d_nn [n_pairs][n_query] int32; d_nn_d2 (may be NULL) float64;
 * d_stats (may be NULL) int[2]: rows that needed the exact scan, chunks re-scored;
 * d_debug_v (may be NULL) [n_pairs][n_query][ceil(n_db/128)*128] filter values.
 * ---------------------------------------------------------------------- */
size_t orpcd_feature_pack_bytes(int n);
int orpcd_feature_center(const double *d_features, int n, double *d_center, orpcd_stream_t stream);
int orpcd_feature_pack(const double *d_features, int n, int n_sets, const double *d_center,
                       void *d_packed, float *d_set_maxnorm, orpcd_stream_t stream);
size_t orpcd_feature_match_workspace_bytes(int n_query, int n_pairs);
int orpcd_feature_match(const double *d_feat_q, const void *d_packed_q, int n_query,
                        const double *d_feat_db, const void *d_packed_db,
                        const float *d_db_maxnorm, int n_db, const double *d_center,
                        const int *d_pair_q, const int *d_pair_db, int n_pairs, int mode, void *d_workspace,
                        size_t workspace_bytes, const unsigned char *d_row_mask, int *d_nn, double *d_nn_d2,
                        int *d_stats, float *d_debug_v, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Batched Fast Global Registration after feature matching: per unit
 * NormalizePointCloud, cross check of the two nearest-neighbour tables,
 * tuple test (Philox4x32-10 stream keyed per unit), graduated non-convexity
 * Geman-McClure solve, transform back to the original scale and inversion.
 * Replaces the body of registration_fgr_based_on_feature_matching
 * (Optimizer/fastGlobalOptimizer.py:158-174) except the final
 * EvaluateRegistration (orpcd_evaluate_batch).
 * d_ij [n_units][n_feat_src]: nearest target feature of every source feature;
 * d_ji [n_units][n_feat_tgt]: nearest source feature of every target feature;
 * d_out_T [n_units][16] source -> target, row-major, column convention;
 * d_out_info [n_units][4] {cross-checked pairs, pairs after the tuple test,
 *                          tuple trials run, swapped}.
 * ---------------------------------------------------------------------- */
typedef struct {
    int n_units;
    int n_src;             /* points per source cloud                         */
    int n_tgt;             /* points per target cloud                         */
    int n_feat_src;        /* rows of the source feature table (= n_src)      */
    int n_feat_tgt;        /* rows of the target feature table (the reference passes the
                              SOURCE's features twice: then n_feat_tgt = n_src) */
    int iteration_number;  /* FastGlobalRegistrationOption.iteration_number   */
    int decrease_mu;
    int max_tuple_count;   /* Open3D default 1000                             */
    int tuple_test;        /* Open3D default true                             */
    double division_factor;
    double tuple_scale;
    double max_correspondence_distance;
} orpcd_fgr_params;

/* Keys of the per-unit tuple-test streams: key[u] = tuple_key(seed, source cloud, target cloud) of
 * or_pcd_b200/Optimizer/fastGlobalOptimizer.py (hash of the seed, the cloud sizes and the bit patterns of
 * each cloud's first and last point).  Stands in for Open3D's unseeded global std::mt19937
 * (FastGlobalRegistration.cpp AdvancedMatching; Optimizer/fastGlobalOptimizer.py:166-174). */
int orpcd_fgr_tuple_keys(const double *d_src_points, int n_src, const double *d_tgt_points, int n_tgt,
                         const int *d_unit_src, const int *d_unit_tgt, int n_units,
                         unsigned long long seed, unsigned long long *d_keys, orpcd_stream_t stream);

size_t orpcd_fgr_batch_workspace_bytes(const orpcd_fgr_params *p);
int orpcd_fgr_batch(const orpcd_fgr_params *p, const double *d_src_points,
                    const double *d_tgt_points, const int *d_unit_src, const int *d_unit_tgt,
                    const int *d_ij, const int *d_ji, const unsigned long long *d_tuple_keys,
                    void *d_workspace, size_t workspace_bytes, double *d_out_T, int *d_out_info,
                    orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Farthest point sampling (PointNet++), float64, bit-exact with the reference's
 * numpy/scipy loop (Preprocessor/Downsamplers/farthestDownsampler.py:33-54):
 * d_out_indices[0] = first_index (the reference draws it with np.random.randint),
 * then sample_size-1 steps of dist = min(dist, |p - cloud|), next = first argmax.
 * d_dist_workspace: [n] float64, needed when n > 16384 (may be NULL otherwise).
 * ---------------------------------------------------------------------- */
int orpcd_farthest_point_sample(const double *d_points, int n, int first_index, int sample_size,
                                double *d_dist_workspace, int *d_out_indices, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Voxel-grid down-sampling: one output point per occupied voxel, the average of
 * the voxel's points taken in index order.  Replaces PointCloud.voxel_down_sample
 * (Preprocessor/Downsamplers/voxelDownsampler.py:83,95).  Open3D emits the voxels
 * in std::unordered_map order (unspecified); here they come in lexicographic
 * (ix, iy, iz) order of the voxel index, so parity is as a set / after sorting.
 * d_out_points [n][3] (first *d_out_count rows valid), d_out_first_index (may be
 * NULL) [n] lowest point index of each voxel, d_out_count int32[1]: number of
 * voxels; -1 when Open3D itself would fail ("voxel_size is too small": voxel_size * INT_MAX
 * < extent); -2 when the grid is too fine for THIS implementation's 64-bit (dense cell id,
 * point index) sort word -- dim_x dim_y dim_z must stay below 2^(63 - ceil(log2 n)); Open3D hashes
 * the occupied cells and has no such limit.
 * ---------------------------------------------------------------------- */
size_t orpcd_voxel_downsample_workspace_bytes(int n);
int orpcd_voxel_downsample(const double *d_points, int n, double voxel_size, void *d_workspace,
                           size_t workspace_bytes, double *d_out_points, int *d_out_first_index,
                           int *d_out_count, orpcd_stream_t stream);

/* ------------------------------------------------------------------------
 * Roofline helper: one nearest-neighbour sweep (the dominant kernel of
 * orpcd_gicp_batch) over `n_units` x n_src queries against n_tgt targets,
 * exactly as one Gauss-Newton iteration issues it, for timing in isolation.
 * Uses a workspace prepared by orpcd_gicp_batch on the same parameters.
 * ---------------------------------------------------------------------- */
int orpcd_gicp_sweep_only(const orpcd_gicp_params *p, void *d_workspace,
                          size_t workspace_bytes, orpcd_stream_t stream);

/* FP32 FFMA pipe micro-benchmark (roofline denominator):
 * launches blocks x 256 threads, each doing iters x 16 dependent-chain FFMAs on
 * 8 independent accumulators.  FLOPs = blocks*256*iters*16*8*2.  d_sink: [blocks*256]. */
int orpcd_ffma_burn(int blocks, int iters, float *d_sink, orpcd_stream_t stream);
/* Same with packed FFMA2 (fma.rn.f32x2): FLOPs = blocks*256*iters*16*8*4. */
int orpcd_ffma2_burn(int blocks, int iters, float *d_sink, orpcd_stream_t stream);
/* Float64 DFMA: FLOPs = blocks*256*iters*16*8*2.  d_sink: [blocks*256] doubles. */
int orpcd_dfma_burn(int blocks, int iters, double *d_sink, orpcd_stream_t stream);
/* Issue-slot probe: FFMA2 chains interleaved with alu_per_fma (0..2) three-input FMNMX per FFMA2. */
int orpcd_mix_burn(int blocks, int iters, int alu_per_fma, float *d_sink, orpcd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ORPCD_B200_H */
