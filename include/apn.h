/*
 * apn.h — C ABI of libapn_sm100.so: the B200 (sm_100a) point-cloud render path of
 * Articulated-Point-NeRF.
 *
 * Every entry point takes raw device pointers, sizes and a cudaStream_t (passed as void*);
 * nothing allocates, nothing synchronises the host unless stated.  All floating point data
 * is fp32, all indices int32 unless stated.  Return value: 0 on success, negative on error;
 * apn_last_error() returns the message of the last failing call on this thread.
 *
 * Each group cites the reference interface it replaces (paths relative to the reference
 * repository root).
 */
#ifndef APN_H_
#define APN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APN_K 8            /* neighbours per sample (lib/temporalpoints.py:42) */
#define APN_C 128          /* feature channels / MLP width (configs/nerf/default.py:60) */
#define APN_POS_FREQS 10   /* posbase_pe  (lib/tineuvox.py:96) */
#define APN_VIEW_FREQS 4   /* viewbase_pe (lib/tineuvox.py:96) */
#define APN_PE_POS 63      /* 3 + 3*2*10 */
#define APN_PE_VIEW 27     /* 3 + 3*2*4  */

typedef void* apn_stream_t; /* cudaStream_t */

int apn_version(void);
const char* apn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long apn_launch_count(void);
/* NVTX range push / pop (nvtx3) for the host-side call groups; names follow the reference's torch.profiler.record_function
 * ranges (lib/temporalpoints.py:421,433,439,452,496,503,546,552,611,629,634,653; lib/pointwarper.py:217,230,241).
 * Return the nesting level NVTX reports (negative without a tool attached). */
int apn_range_push(const char* name);
int apn_range_pop(void);

/* ---------------------------------------------------------------------------------------
 * K0  Pose chain.  Replaces lib/pointwarper.py:217-236: TransformNet (lib/pointwarper.py:5-37; 256-wide, 4 hidden
 * layers, bias-free output layer), 4-parameter Rodrigues (:118-143), per-joint pivot transforms and the kinematic
 * chain product (:145-193), forward and backward, one single-CTA launch each.
 *   w5 / b4: HOST arrays of device pointers to the Linear weights (torch layout) / biases
 *   parent_node (J): chain parent of node i (-1 for the root; must be < i)   pivot (J): joint rotated about
 *   sibling (J): node whose rotation node i uses (:232)   rot_mask (J) uint8 or NULL: rotation frozen to identity
 *   bone_T (J,4,4)  global_t (3)  thetas (J)  saved: apn_pose_saved_bytes(J) bytes kept for the backward (or NULL)
 * ------------------------------------------------------------------------------------- */
size_t apn_pose_saved_bytes(int J);
int apn_pose_fwd(const float* t_embed, int t_dim, const float* const* w5, const float* const* b4, const float* joints,
                 const int32_t* parent_node, const int32_t* pivot, const int32_t* sibling, const uint8_t* rot_mask, int J,
                 float* bone_T, float* global_t, float* thetas, void* saved, apn_stream_t stream);
/* d_w5 / d_b4 / d_joints are overwritten; d_global_t / d_thetas may be NULL */
int apn_pose_bwd(const float* t_embed, int t_dim, const float* const* w5, const float* const* b4, const float* joints,
                 const int32_t* parent_node, const int32_t* pivot, const int32_t* sibling, const uint8_t* rot_mask, int J,
                 const void* saved, const float* d_bone_T, const float* d_global_t, const float* d_thetas,
                 float* const* d_w5, float* const* d_b4, float* d_joints, apn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K1  Linear blend skinning.
 * Replaces lib/temporalpoints.py:401-414 (get_weights: softmax(raw/max(eps,theta)) + merge),
 * lib/pointwarper.py:241-266 (blend of bone 4x4s, point transform, + global_t),
 * lib/temporalpoints.py:569 (torch.inverse of the blended frame; only [:3,:3] is consumed,
 * lib/temporalpoints.py:478) and lib/temporalpoints.py:424 (min/max of the warped cloud).
 *   raw_w (N,J)  theta_weight (1)  merge_rules (J) int32 or NULL (= identity)
 *   bone_T (J,4,4) row-major  xyz (N,3)  global_t (3)
 *   xyz_out (N,3)  ginv_out (N,9) = inverse(G[:3,:3])  w_out (N,J) merged weights or NULL
 *   g_out (N,4,4) blended frames (what PointWarper.forward(get_frames=True) returns) or NULL
 *   theta_weight == NULL: raw_w already holds the final (soft-maxed, merged) weights, as passed
 *   to PointWarper.forward (lib/pointwarper.py:213)
 *   bbox (6) = min xyz, max xyz of xyz_out (initialised by the call)
 * ------------------------------------------------------------------------------------- */
int apn_lbs_fwd(const float* raw_w, const float* theta_weight, float eps, const int32_t* merge_rules,
                const float* bone_T, const float* xyz, const float* global_t, int N, int J,
                float* xyz_out, float* ginv_out, float* w_out, float* g_out, float* bbox, apn_stream_t stream);

size_t apn_lbs_bwd_workspace_bytes(int N, int J);
/* d_raw (N,J), d_theta (1), d_bone_T (J,4,4; last row 0), d_global_t (3) are overwritten.
 * d_w (N,J) is the extra gradient arriving on the merged weights (regularisers) or NULL;
 * d_g (N,4,4) the gradient arriving on g_out or NULL.
 * Without merge rules and d_g, and for J <= 80, the two contractions of the backward run on the tensor cores
 * (mma.sync TF32 with the 3xTF32 split: fp32-class, within 1e-4 of the fp32 path); otherwise on the CUDA cores. */
int apn_lbs_bwd(const float* raw_w, const float* theta_weight, float eps, const int32_t* merge_rules,
                const float* bone_T, const float* xyz, int N, int J, const float* ginv,
                const float* d_xyz, const float* d_ginv, const float* d_w, const float* d_g,
                float* d_raw, float* d_theta, float* d_bone_T, float* d_global_t,
                void* workspace, size_t workspace_bytes, apn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K2  Uniform multi-level grid over the warped cloud + exact 8-NN of ray samples.
 * Replaces lib/temporalpoints.py:373-399 (sample_ray -> render_utils_cuda.sample_pts_on_rays,
 * lib/cuda/render_utils_kernel.cu:12-236), lib/temporalpoints.py:434-437 (KeOps
 * Kmin_argKmin(K=8) brute force) and lib/temporalpoints.py:440-444 (radius rule).
 * Neighbour contract: d2 = (dx*dx + dy*dy) + dz*dz in fp32 without FMA, the 8 smallest
 * ascending by (d2, point index); a sample is kept iff it lies inside the padded bbox and
 * its 8th d2 <= query_radius (the reference compares the SQUARED distance with 0.01).
 * ------------------------------------------------------------------------------------- */
size_t apn_grid_workspace_bytes(int N, int cell_capacity);
/* Builds the grid in `grid` (opaque blob of apn_grid_workspace_bytes). bbox (6) comes from
 * apn_lbs_fwd.  cell_hint ~ 1.5x mean point spacing; bbox_pad = query_radius (0.01). */
int apn_grid_build(const float* xyz, const float* bbox, int N, float query_radius, float bbox_pad,
                   float cell_hint, int cell_capacity, void* grid, size_t grid_bytes, apn_stream_t stream);
/* copies the 64-float/int descriptor of the grid to host memory (synchronises the stream) */
int apn_grid_describe(const void* grid, float* host_out64, apn_stream_t stream);

/* exclusive scan of n int32 -> out[0..n] (out[n] = total); ws >= apn_scan_workspace_bytes(n) */
size_t apn_scan_workspace_bytes(int n);
int apn_exclusive_scan_i32(const int32_t* in, int32_t* out, int n, void* ws, size_t ws_bytes, apn_stream_t stream);

/* pass 1: number of candidate samples per ray (inside the padded bbox and with >= 8 points
 * in the coarse neighbourhood); pass 2 (fill=1) writes (ray, step) of every candidate at
 * cand_base[ray]+k. */
int apn_ray_candidates(const float* rays_o, const float* rays_d, int R, float near, float far, float stepdist,
                       const void* grid, int fill, int32_t* cand_count, const int32_t* cand_base,
                       int32_t* cand_ray, int32_t* cand_step, apn_stream_t stream);
/* exact 8-NN of every candidate; keep[i] = 1 iff the sample survives the radius rule.
 * nn_idx (n_cand,8) original point indices ascending by (d2, index); nn_d2 (n_cand,8) or NULL. */
int apn_knn(const float* rays_o, const float* rays_d, float near, float far, float stepdist, const void* grid,
            const int32_t* cand_ray, const int32_t* cand_step, int n_cand,
            int32_t* nn_idx, float* nn_d2, int32_t* keep, apn_stream_t stream);
/* The same contract as apn_knn, served by the cell-sorted search: the candidates are radix-sorted by the grid leaf that holds
 * them, a warp's 32 queries share one neighbourhood whose points are staged in shared memory, and every lane scans the
 * stage against its own query with a register top-8 (csrc/grid_knn.cu, knn_sorted_kernel).  Bit-identical results.
 * workspace: apn_knn_sorted_workspace_bytes(n_cand) bytes (sort keys / values + radix-sort temporaries). */
size_t apn_knn_sorted_workspace_bytes(int n_cand);
int apn_knn_sorted(const float* rays_o, const float* rays_d, float near, float far, float stepdist, const void* grid,
                   const int32_t* cand_ray, const int32_t* cand_step, int n_cand,
                   int32_t* nn_idx, float* nn_d2, int32_t* keep, void* workspace, size_t workspace_bytes, apn_stream_t stream);
/* order-preserving compaction of the kept candidates (kept_pos = exclusive scan of keep).
 * ray_start (R+1): first kept sample of each ray. */
int apn_compact_samples(const float* rays_o, const float* rays_d, float near, float far, float stepdist,
                        const void* grid, const int32_t* cand_ray, const int32_t* cand_step, const int32_t* cand_base,
                        const int32_t* keep, const int32_t* kept_pos, const int32_t* nn_idx_cand, int n_cand, int R,
                        float* pts, int32_t* ray_id, int32_t* step_id, int32_t* nn_idx, int32_t* ray_start,
                        apn_stream_t stream);
/* Sync-free sampling stage: count -> scan -> fill -> exact 8-NN -> scan -> compact in one call with every length kept on
 * the device (the reference reads its sample count back, lib/cuda/render_utils_kernel.cu:205-206: N_steps.sum().item()).
 * The candidate list (cand_cap) and the sample arrays (m_cap rows of pts / ray_id / step_id / nn_idx) have fixed capacities;
 * counts (5 x int32, device): [0] candidates used, [1] samples kept = the m_dev of apn_agg_inputs, [2] flags (1 grid
 * overflow | 2 candidate list truncated | 4 sample arrays truncated), [3] candidates found, [4] samples found.
 * Same results as apn_ray_candidates + apn_knn + apn_compact_samples whenever counts[2] == 0; capturable in a CUDA graph. */
size_t apn_sample_knn_static_workspace_bytes(int R, int cand_cap);
int apn_sample_knn_static(const float* rays_o, const float* rays_d, int R, float near, float far, float stepdist,
                          const void* grid, int cand_cap, int m_cap, void* workspace, size_t workspace_bytes,
                          float* pts, int32_t* ray_id, int32_t* step_id, int32_t* nn_idx, int32_t* ray_start,
                          int32_t* counts, apn_stream_t stream);
/* brute-force-free k-NN of arbitrary query points against the grid (init self-k-NN,
 * lib/temporalpoints.py:104-111; chamfer K=1, :747-751). max_d2 <= 0: unbounded. */
int apn_knn_points(const float* query, int n_query, const void* grid, int k, int32_t* nn_idx, float* nn_d2,
                   apn_stream_t stream);
/* batched exact nearest neighbour, K = 1, dim 2 or 3 (batch chamfer loss, lib/temporalpoints.py:783-787: KeOps
 * D_ij.argKmin(dim=2, K=1) / argKmin(dim=1, K=1) over (B, N, M); run.py:659-690).  query (B, n_query, dim),
 * target (B, n_target, dim) -> nn_idx (B, n_query), index into that batch item's targets; ties -> lowest index. */
int apn_nn1_batched(const float* query, const float* target, int n_batch, int n_query, int n_target, int dim,
                    int32_t* nn_idx, apn_stream_t stream);
/* Camera rays on the device: get_rays_of_a_view (lib/tineuvox.py:675-738, mode='center', ndc=False) in one launch from the
 * camera (K 3x3 and c2w (3|4)x4, HOST pointers, passed to the kernel by value) — for n consecutive pixels starting at
 * first_pixel (row-major), or for the pixels listed in pixel_ids (device, n entries: a rank's tiles, a training batch).
 * rays_o / rays_d / viewdirs (n,3); viewdirs may be NULL. */
int apn_rays_of_a_view(const float* K_host9, const float* c2w_host, int c2w_rows, int H, int W, int inverse_y, int flip_x,
                       int flip_y, const int32_t* pixel_ids, long long first_pixel, int n,
                       float* rays_o, float* rays_d, float* viewdirs, apn_stream_t stream);
/* Stage-2 regulariser losses with their gradients (run.py:633-657; lib/temporalpoints.py:714-725), one launch over the
 * canonical points and their static neighbourhood nn_i (N,K) (lib/temporalpoints.py:104-110):
 *   losses3[0] = weight_arap * sum_ik |nn_dist_ik - sqrt(|xyz_i - xyz_n|^2 + eps)|               (get_arap_loss)
 *   losses3[1] = weight_tv * mean_ikj |w_ij - w_nj|                                              (get_neighbour_weight_tv_loss)
 *   losses3[2] = weight_sparsity * -mean_ij [w log(w+eps) + (1-w) log(1-w+eps)]                  (get_weight_sparsity_loss)
 * xyz (N,3) warped cloud, w (N,J) merged skinning weights.  d_xyz (N,3) is ACCUMULATED into (it already holds the
 * render-loss gradient), d_w (N,J) is OVERWRITTEN (zeroed first); both feed apn_lbs_bwd.  A zero weight skips its term
 * (and its pointers may then be NULL). */
int apn_point_regularisers(const float* xyz, const float* w, const int32_t* nn_i, const float* nn_dist, int N, int K, int J,
                           float eps, float weight_arap, float weight_tv, float weight_sparsity, float* d_xyz, float* d_w,
                           float* losses3, apn_stream_t stream);
/* Pose-side regularisers in one launch:
 *   losses2[0] = weight_transformation_reg * (sum|global_t| + sum|thetas|) / J     (get_transformation_regularisation_loss,
 *                lib/temporalpoints.py:797-800) -> d_thetas (J), d_global_t (3) OVERWRITTEN; skipped if d_thetas NULL
 *   losses2[1] = weight_joint_chamfer * sum_j min_s |joints_j - skeleton_s|^2       (get_joint_chamfer_loss, :731-733; nearest
 *                by the k-NN contract, ties -> lowest index) -> d_joints (J,3) OVERWRITTEN; skipped if d_joints NULL */
int apn_pose_regularisers(const float* thetas, const float* global_t, const float* joints, const float* skeleton, int J, int S,
                          float weight_transformation_reg, float weight_joint_chamfer, float* d_thetas, float* d_global_t,
                          float* d_joints, float* losses2, apn_stream_t stream);
/* time embedding of the pose-network input: poc_fre of the scalar time (lib/tineuvox.py:872-878; lib/temporalpoints.py:546-550):
 * out (1 + 2 n_freq) = [t, sin(t f_i)..., cos(t f_i)...] */
int apn_time_embed(const float* t, const float* freqs, int n_freq, float* out, apn_stream_t stream);
/* render loss of the stage-2 loop and its gradient (run.py:617-621): loss (1) = weight * mean((pred - target)^2),
 * grad (n) = (pred - target) * 2 weight / n; fixed summation order */
int apn_mse_loss_grad(const float* pred, const float* target, int n, float weight, float* loss, float* grad,
                      apn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K3  Aggregation: gather + positional encoding + feature MLP + inverse-distance reduce + heads.
 * Replaces lib/temporalpoints.py:446-515 (aggregate_pts after the k-NN), lib/tineuvox.py:65-88
 * (RGBNet), :158 (densitynet), :396-400/646-670 (activate_density / Raw2Alpha ->
 * lib/cuda/render_utils_kernel.cu:358-428), :872-878 (poc_fre).
 * ------------------------------------------------------------------------------------- */
typedef struct apn_mlp_weights {
  /* feat_net: Linear(d_in,128)+LeakyReLU, 2x[Linear(128,128)+LeakyReLU], Linear(128,128)+LeakyReLU
   * (lib/temporalpoints.py:123-130). torch Linear layout: W (out,in) row-major. */
  const float* w[4];
  const float* b[4];
  const float* density_w;   /* (1,128) */
  const float* density_b;   /* (1)     */
  const float* rgb_feat_w;  /* (128,128) rgbnet.feature_linears */
  const float* rgb_feat_b;  /* (128) */
  const float* rgb_v0_w;    /* (64,155)  rgbnet.views_linears.0 */
  const float* rgb_v0_b;    /* (64) */
  const float* rgb_v2_w;    /* (3,64)    rgbnet.views_linears.2 */
  const float* rgb_v2_b;    /* (3) */
} apn_mlp_weights;

typedef struct apn_agg_inputs {
  int M;                       /* kept samples */
  int N;                       /* points */
  int d_in;                    /* 191, or 255 with a 64-wide pose embedding */
  const float* pts;            /* (M,3) sample positions */
  const int32_t* nn_idx;       /* (M,8) */
  const int32_t* ray_id;       /* (M) */
  const float* xyz;            /* (N,3) warped cloud */
  const float* ginv;           /* (N,9) */
  const float* feat;           /* (N,128) canonical_feat */
  const float* pose_emb;       /* (d_in-191) or NULL */
  const float* viewdirs;       /* (R,3) */
  /* direct branch (lib/temporalpoints.py:459-470) */
  const float* canonical_alpha; /* (N) */
  const float* canonical_rgbs;  /* (N,3) */
  const float* direct_eps;      /* (N) */
  float mean_min_distance;
  float eps;                   /* 1e-6 */
  float act_shift;
  float interval;              /* stepsize * voxel_size_ratio */
  /* Optional device-side sample count (NULL: M is exact).  When given, M is the CAPACITY of the per-sample arrays and
   * the kernels read the true count min(*m_dev, M) on the device: the host never reads a count back, so a whole
   * training step can be enqueued ahead of the GPU and captured as a CUDA graph (the reference synchronises on its
   * sample count, lib/cuda/render_utils_kernel.cu:205-206).  Honoured by the tensor-core entry points
   * (apn_aggregate_fwd_tc, apn_aggregate_bwd_tc); rows >= *m_dev of every per-sample output are left untouched. */
  const int32_t* m_dev;
} apn_agg_inputs;

typedef struct apn_agg_outputs {
  float* alpha;         /* (M) */
  float* rgb;           /* (M,3) after sigmoid */
  float* alpha_direct;  /* (M) */
  float* rgb_direct;    /* (M,3) */
  float* idw;           /* (M,8) normalised inverse-distance weights (also used by render_weights) */
  /* saved for backward (may be NULL in inference): */
  float* x0;            /* (8M, ld0) MLP input rows, ld0 = round_up(d_in, 4) */
  float* act[4];        /* (8M,128) post-activation of each feat_net layer */
  float* h;             /* (M,128) reduced feature */
  float* exp_d;         /* (M) exp(density+shift) */
  float* fv;            /* (M,160) [rgb feature 128 | view PE 27 | pad] */
  float* v0;            /* (M,64) hidden of views_linears */
} apn_agg_outputs;

size_t apn_aggregate_scratch_bytes(int M, int d_in);
/* fp32 CUDA-core path (parity mode; also the training path). If out->x0 etc. are NULL the
 * activations live in `scratch` only. */
int apn_aggregate_fwd(const apn_agg_inputs* in, const apn_mlp_weights* w, const apn_agg_outputs* out,
                      void* scratch, size_t scratch_bytes, apn_stream_t stream);

typedef struct apn_agg_grads {
  /* incoming */
  const float* d_alpha;  /* (M) */
  const float* d_rgb;    /* (M,3) */
  /* outgoing; all ACCUMULATED into (caller zeroes) */
  float* d_xyz;          /* (N,3) */
  float* d_ginv;         /* (N,9) */
  float* d_feat;         /* (N,128) */
  float* d_pose_emb;     /* (d_in-191) or NULL */
  float* d_w[4];
  float* d_b[4];
  float* d_density_w;
  float* d_density_b;
  float* d_rgb_feat_w;
  float* d_rgb_feat_b;
  float* d_rgb_v0_w;
  float* d_rgb_v0_b;
  float* d_rgb_v2_w;
  float* d_rgb_v2_b;
} apn_agg_grads;

size_t apn_aggregate_bwd_scratch_bytes(int M, int d_in);
int apn_aggregate_bwd(const apn_agg_inputs* in, const apn_mlp_weights* w, const apn_agg_outputs* saved,
                      const apn_agg_grads* g, void* scratch, size_t scratch_bytes, apn_stream_t stream);

/* tcgen05 (5th-gen tensor core, accumulators in tensor memory) fused inference path: same contract as
 * apn_aggregate_fwd for alpha / rgb / alpha_direct / rgb_direct / idw; nothing is saved for backward.
 * precision: 0 = fp16 operands, fp32 accumulate (one MMA per K step);
 *            1 = split operands x = hi + lo (two fp16 terms), hi*hi + hi*lo + lo*hi in fp32: fp32-class (parity mode).
 * packed_weights: apn_aggregate_tc_weights_bytes(d_in) bytes filled by apn_aggregate_tc_pack_weights (re-pack after
 * every weight update); scratch >= apn_aggregate_tc_scratch_bytes(M). */
size_t apn_aggregate_tc_weights_bytes(int d_in);
int apn_aggregate_tc_pack_weights(const apn_mlp_weights* w, int d_in, void* packed, apn_stream_t stream);
/* point_table (N,128) = canonical_feat * W0[:, 63:191]^T in exact fp32: the feature columns of feat_net's first
 * layer are a gather of per-point rows, so their product with W0 is tabulated per point (rebuild when
 * canonical_feat or feat_net.0.weight change) and added back in the layer-0 epilogue. */
int apn_aggregate_tc_point_table(const float* feat, const float* w0, int d_in, int N, float* point_table,
                                 apn_stream_t stream);
size_t apn_aggregate_tc_scratch_bytes(int M);
/* tape == NULL: inference (scratch required).  tape != NULL (training, precision 1 only): the kernel also records
 * what the tensor-core backward needs (post-activations, PE operand, LeakyReLU masks; apn_aggregate_tc_tape_bytes(M)
 * bytes, 1 KiB aligned) and out->h / exp_d / fv / v0 must be given (kept by the caller for the backward). */
size_t apn_aggregate_tc_tape_bytes(int M);
int apn_aggregate_fwd_tc(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_weights,
                         const float* point_table, const apn_agg_outputs* out, int precision, void* tape,
                         size_t tape_bytes, void* scratch, size_t scratch_bytes, apn_stream_t stream);

/* Tensor-core backward of the same op (replaces autograd through lib/temporalpoints.py:446-515): split-fp16 dgrad chain
 * + wgrad with accumulators in tensor memory; heads in fp32.  Same gradient contract as apn_aggregate_bwd (all
 * outgoing buffers accumulated into; caller zeroes).  d_in = 191, or up to 256 with a pose embedding (lib/temporalpoints.py:
 * 483-490,571-576): the embedding is the same vector for every row, so its weight columns act as a layer-0 bias in the
 * forward and its gradients follow from the layer-0 bias gradient (d_e = W0[:,191:]^T db0, dW0[:,191:] += db0 (x) e).
 * packed_bwd: apn_aggregate_tc_bwd_weights_bytes() bytes filled by apn_aggregate_tc_pack_weights_bwd. */
size_t apn_aggregate_tc_bwd_weights_bytes(void);
int apn_aggregate_tc_pack_weights_bwd(const apn_mlp_weights* w, int d_in, void* packed_bwd, apn_stream_t stream);
size_t apn_aggregate_tc_bwd_scratch_bytes(int M, int N);
int apn_aggregate_bwd_tc(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_bwd,
                         const apn_agg_outputs* saved, const void* tape, const apn_agg_grads* g, void* scratch,
                         size_t scratch_bytes, apn_stream_t stream);
/* The same backward in two parts for data-parallel training (not in the reference, which is single-GPU): phase 1 runs
 * everything up to g->d_feat (heads, density, dgrad chain, d_xyz / d_ginv, d_feat = dP W0_feat) — the gradient of the point
 * features, ~90 % of the bytes a rank exchanges, is final when it returns and its all-reduce can start; phase 2 adds the
 * feat_net weight gradients, the point-table weight gradient and the pose-embedding gradient from the SAME scratch buffer
 * (untouched in between).  phase 0 == apn_aggregate_bwd_tc.  Phase 1 followed by phase 2 gives the same results.
 * A second split for the one-GPU step: phase 3 stops as soon as d_xyz / d_ginv are final (heads, density, dgrad chain) — the
 * LBS / pose backward can start —, phase 4 adds EVERY parameter gradient (d_feat and the weight gradients); the caller may run
 * phase 4 on another stream beside the LBS / pose backward.  Phase 3 followed by phase 4 gives the same results as phase 0.
 * Phase 4 = phase 5 (d_feat = dP W0_feat only, entirely on the caller's stream) + phase 6 (the weight gradients), in any order
 * or on two streams: they write disjoint buffers and only read what phase 3 left in the scratch. */
int apn_aggregate_bwd_tc_phase(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_bwd,
                               const apn_agg_outputs* saved, const void* tape, const apn_agg_grads* g, void* scratch,
                               size_t scratch_bytes, int phase, apn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K4  Ray compositing.
 * Replaces lib/temporalpoints.py:611-677: alpha > thres mask, Alphas2Weights
 * (lib/tineuvox.py:627-643 -> lib/cuda/render_utils_kernel.cu:431-561, early stop at T<1e-3),
 * weight > thres mask, torch_scatter.segment_coo sums for rgb / depth, + alphainv_last*bg.
 *   ray_start (R+1) offsets into the ray-major sample list
 *   T_save (M) and n_used (R) are written when non-NULL (needed by the backward)
 *   the per-sample arrays (alpha, rgb, step_id, T_save, d_alpha, d_rgb) must be 16-byte aligned: the kernels move them with
 *   16-byte vector accesses and, in the backward, with TMA bulk copies of each warp's 32-ray span
 * ------------------------------------------------------------------------------------- */
int apn_composite_fwd(const float* alpha, const float* rgb, const int32_t* step_id, const float* extra, int n_extra,
                      const int32_t* ray_start, int R, float thres, float bg,
                      float* rgb_marched, float* alphainv_last, float* depth, float* extra_marched,
                      float* T_save, int32_t* n_used, apn_stream_t stream);
int apn_composite_bwd(const float* alpha, const float* rgb, const int32_t* step_id, const int32_t* ray_start, int R,
                      float thres, float bg, const float* T_save, const int32_t* n_used, const float* alphainv_last,
                      const float* d_rgb_marched, const float* d_alphainv_last, const float* d_depth,
                      float* d_alpha, float* d_rgb, apn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K4b  Adam.  Replaces adam_upd_cuda.{adam_upd, masked_adam_upd, adam_upd_with_perlr}
 * (lib/cuda/adam_upd.cpp:79-86, lib/cuda/adam_upd_kernel.cu:9-132) called by
 * lib/masked_adam.py:39-72.  Multi-tensor: one launch for a list of tensors.
 *   mode 0 = plain, 1 = skip where grad==0, 2 = per-element lr (perlr != NULL)
 * ------------------------------------------------------------------------------------- */
typedef struct apn_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  const float* perlr;   /* mode 2 only */
  long long numel;
  float step_size;      /* lr*sqrt(1-b2^t)/(1-b1^t), computed in float as adam_upd_kernel.cu:72 */
  int mode;
} apn_adam_tensor;
float apn_adam_step_size(int step, float beta1, float beta2, float lr);
/* `tensors` is a HOST array of n_tensors descriptors (copied into kernel parameters in chunks). */
int apn_adam_multi(const apn_adam_tensor* tensors, int n_tensors, float beta1, float beta2, float eps,
                   apn_stream_t stream);
/* The same update with the per-tensor step sizes in DEVICE memory (step_sizes_dev[n_tensors]; the host refreshes them
 * before a captured CUDA graph is replayed — kernel arguments are frozen at capture, lib/masked_adam.py:62 changes the
 * bias-corrected step size every iteration) and an optional device-side skip word: *skip_dev != 0 makes the launch a
 * no-op (the step's sample workspace overflowed, apn_sample_knn_static counts[2], and the step is re-run). */
int apn_adam_multi_dev(const apn_adam_tensor* tensors, int n_tensors, float beta1, float beta2, float eps,
                       const float* step_sizes_dev, const int32_t* skip_dev, apn_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Reference-compatible single ops (the pybind surface of lib/cuda/render_utils.cpp:144-155).
 * int64 ids as in the reference.  sample_pts_on_rays is two-phase because the total is
 * data dependent (the reference does N_steps.sum().item(), render_utils_kernel.cu:206).
 * ------------------------------------------------------------------------------------- */
int apn_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min, const float* xyz_max,
                       float near, float far, int R, float* t_min, float* t_max, apn_stream_t stream);
int apn_infer_n_samples(const float* t_min, const float* t_max, float stepdist, int R, int64_t* n_samples,
                        apn_stream_t stream);
int apn_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min, int R,
                            float* rays_start, float* rays_dir, apn_stream_t stream);
/* n_cumsum (R) inclusive cumsum of n_samples; total = n_cumsum[R-1] known to the caller */
int apn_sample_pts_on_rays_fill(const float* rays_start, const float* rays_dir, const float* xyz_min,
                                const float* xyz_max, const int64_t* n_cumsum, int R, long long total, float stepdist,
                                float* pts, uint8_t* mask_outbbox, int64_t* ray_id, int64_t* step_id,
                                apn_stream_t stream);
int apn_raw2alpha(const float* density, float shift, float interval, long long n, float* exp_d, float* alpha,
                  apn_stream_t stream);
int apn_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval, long long n, float* grad,
                           apn_stream_t stream);
/* weight/T/alphainv_last/i_start/i_end must be pre-filled 0/1/1/0/0 as render_utils_kernel.cu:478-482 */
int apn_alpha2weight(const float* alpha, const int64_t* ray_id, long long n_pts, int n_rays, float* weight, float* T,
                     float* alphainv_last, int64_t* i_start, int64_t* i_end, apn_stream_t stream);
int apn_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* alphainv_last,
                              const int64_t* i_start, const int64_t* i_end, int n_rays, const float* grad_weights,
                              const float* grad_last, float* grad, apn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* APN_H_ */
