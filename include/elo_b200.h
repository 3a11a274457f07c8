/*
 * elo_b200.h -- C ABI of libelo_b200.so, the B200 (sm_100a) implementation of EfficientLO-Net's
 * projection-aware point-cloud hot path.  Plain pointers and sizes only; no torch / TensorFlow
 * types.  Every pointer is a DEVICE pointer unless the function name ends in `_host`.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, as the reference uses).
 * All functions return 0 on success, a positive cudaError_t if the CUDA runtime failed, or a
 * negative ELO_ERR_* code; elo_last_error() gives the message (thread-local).
 * Nothing here allocates device memory, synchronises, or reads results back on the host, so
 * every device-pointer entry point is CUDA-graph capturable.
 *
 * Reference interfaces replaced (file:line relative to the reference repository):
 *   elo_fused_conv_select_k  <-  FusedConvSelectKLauncher  tf_ops/2d_conv_select_k/fused_conv.cpp:73,
 *                                                           tf_ops/2d_conv_select_k/fused_conv_g.cu:215
 *   elo_fused_conv_random_k  <-  FusedConvRandomKLauncher  tf_ops/2d_conv_random_k/fused_conv.cpp:73,
 *                                                           tf_ops/2d_conv_random_k/fused_conv_g.cu:162
 *   (+ the zero-fill of fused_conv.cpp:154-166, which the kernels here do themselves: every output
 *    element is written, callers need not clear the buffers.)
 */
#ifndef ELO_B200_H
#define ELO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ELO_OK 0
#define ELO_ERR_INVALID_ARGUMENT (-1) /* the conditions fused_conv.cpp:79-100 rejects, null pointers */
#define ELO_ERR_UNSUPPORTED (-2)      /* legal for the reference op but outside this build's limits */

const char *elo_last_error(void);
int elo_version(void);
long long elo_launch_count(void); /* kernels launched by this library in this process so far */

/*
 * Projection-aware neighbour search.  Same parameter list and meaning as the reference Launchers,
 * plus the stream.  Shapes (C-contiguous):
 *   xyz1 (B,H,W,3) f32   xyz2 (B,small_h,small_w,3) f32   idx_n2 (B,npoints,2) i32 [h,w]
 *   random_hw (kH*kW) i32
 *   selected_bhw_idx (B,npoints,K,3) i32 [b,h,w]          selected_mask (B,npoints,K,1) f32
 *   valid_idx, valid_in_dis_idx (B,npoints,kH*kW,1) f32   -- may be NULL (no caller in the model
 *   reads them, utils/pointnet_util.py:49,106,197,272); they are 83 % of the op's bytes.
 * Limits: kH*kW <= 5000 and K <= 5000 (the reference's local-array size, fused_conv_g.cu:42-43),
 * H, W, small_h, small_w < 32768.
 */
int elo_fused_conv_select_k(int batch_size, int H, int W, int npoints, int kernel_size_H,
                            int kernel_size_W, int K, int flag_copy, float distance, int stride_h,
                            int stride_w, const float *xyz1, const float *xyz2, const int *idx_n2,
                            const int *random_hw, int *selected_bhw_idx, float *valid_idx,
                            float *valid_in_dis_idx, float *selected_mask, int small_h, int small_w,
                            void *stream);

int elo_fused_conv_random_k(int batch_size, int H, int W, int npoints, int kernel_size_H,
                            int kernel_size_W, int K, int flag_copy, float distance, int stride_h,
                            int stride_w, const float *xyz1, const float *xyz2, const int *idx_n2,
                            const int *random_hw, int *selected_bhw_idx, float *valid_idx,
                            float *valid_in_dis_idx, float *selected_mask, int small_h, int small_w,
                            void *stream);


/* ------------------------------------------------------------------------------------------------
 * Fused blocks.  These have no single native counterpart in the reference: each replaces the chain
 * of stock TensorFlow kernels that one Python block of utils/pointnet_util.py / model_util.py builds
 * around the custom op (gather_nd, concat, 1x1 conv + bias + batch-norm + ReLU per layer, mask,
 * reduce_max / softmax / reduce_sum).  Descriptors are plain C structs; tensors are fp32,
 * C-contiguous, channel-last, on the device.
 *
 * Weights are packed by efficientlo-net_b200/packing.py with the inference batch-norm folded in.
 * FFMA engine: per layer, rows 0..Cin-1 = W'[k][0..Cout), row Cin = bias, zero rows up to a multiple of
 * (2048 / Cout); layers back to back in execution order.  Tensor-core engine: per layer ceil(Cin / R)
 * chunks of R = 2048 / Cout k-rows, each chunk [hi | lo] tf32 halves in the K-major core-matrix order
 * [R/4][Cout][4]; after the last chunk of the last layer, the biases of all layers in order.
 */
/* Which engine runs the per-group MLPs of the fused blocks below:
 *   1 (default)  tcgen05 tensor cores, tf32 x 3 split (fp32-grade accuracy), 128-row tiles, activations in TMEM;
 *   0            fp32 FFMA out of shared memory (the first implementation; kept as a cross-check).
 * The packed `weights` a descriptor carries must match the engine (packing.pack_stream_tc / pack_stream). */
int elo_set_mlp_engine(int engine);
int elo_get_mlp_engine(void);
/* How the tensor-core kernels cut a call into 128-row tiles (results are identical):
 *   0 (default)  latency: a call that needs fewer tiles than there are SMs is spread over all of them (fewer
 *                rows per tile: shorter gather / pooling phases, the shortest single forward);
 *   1            throughput: full tiles, as few CTAs as the work needs -- less SM time per forward, which is
 *                what counts when several independent forwards are in flight on different streams.
 * Read when a forward is launched (or captured into a CUDA graph). */
int elo_set_tile_policy(int policy);
int elo_get_tile_policy(void);
/* Work decomposition of the two stand-alone index ops (results are identical, bit for bit):
 *   0 (default)  by size: calls with at least 2 x 64 queries per SM, K <= 32 and distance^2 < 1e10 take the
 *                tile-staged thread-per-query kernel (fused_conv_tiled.cu), everything else one warp per query;
 *   1            tiled whenever K <= 32 and distance^2 < 1e10;   2   always one warp per query. */
int elo_set_index_kernel(int which);
/* The tile-staged select-K kernel has a form with one extra warp per CTA, the store warp, that writes the CTA's
 * count rows (valid_idx / valid_in_dis_idx) while the query warps walk their windows.  It is used for windows of
 * at least `min_cells` cells (default 128: it wins on 7x25, loses a little on 5x15); 0 = always, a huge value =
 * never.  Results are identical either way. */
int elo_set_store_warp_min_cells(int min_cells);
int elo_get_store_warp_min_cells(void);
/* How the tile-staged index kernel brings a CTA's neighbourhood of the xyz2 grid into shared memory:
 *   0 (default)  bulk copies through the TMA engine (cp.async.bulk, one or two per tile row), spread in place to the
 *                padded (x, y, z, empty) float4 grid; needs xyz2 on a 16-byte boundary, else falls back to 1;
 *   1            plain loads by the CTA's threads.
 * Results are identical either way. */
int elo_set_tile_staging(int mode);
int elo_get_tile_staging(void);
int elo_get_index_kernel(void);
/* Programmatic dependent launch between the kernels of this library (default on; environment ELO_PDL=0
 * turns it off): a kernel's prologue -- barrier / tensor-memory set-up, weight prefetch -- overlaps the
 * tail of the kernel before it.  Results are identical either way. */
int elo_set_pdl(int on);
int elo_get_pdl(void);
/* Debug: device buffer of 64 int64; CTA (0,0) of every tensor-core kernel launched afterwards writes phase
 * timestamps (ns, %globaltimer) into it: compute thread 0 -> [0,32), MMA thread -> [32,64).  NULL = off. */
int elo_set_time_log(long long *device_buf);

typedef struct {
    int kernel_size_H, kernel_size_W, K;
    float distance;
    int stride_h, stride_w;   /* query (h,w) -> window centre (h/stride_h, w/stride_w) in the searched grid */
    int small_h, small_w;     /* extent of the searched grid */
    const int *random_hw;     /* (kernel_size_H*kernel_size_W) scan order */
} elo_window;

typedef struct {
    int H, W;                    /* query image */
    int out_h, out_w;            /* queries per sample: cells (i*q_stride_h, j*q_stride_w) */
    int q_stride_h, q_stride_w;  /* 1,1 = every pixel (get_hw_idx); >1 = strided centres (get_selected_idx) */
} elo_queries;

/* Batched neighbour search: nspec independent select-K / random-K searches in one launch, each writing
 * the compact table the fused blocks take as `nbr`: out_nbr (B, out_h*out_w, K) int32 = linear cell
 * (h * small_w + w) of the k-th selected neighbour in the searched grid, -1 where selected_mask is 0.
 * Same selection, bit for bit, as elo_fused_conv_{select,random}_k with flag_copy = 0. */
#define ELO_MAX_SEARCH 16
typedef struct {
    int select;               /* 1 = select-K, 0 = random-K */
    int batch_size;
    elo_queries queries;
    elo_window window;
    const float *xyz1;        /* (B, H, W, 3) query image */
    const float *xyz2;        /* (B, small_h, small_w, 3) searched grid */
    int *out_nbr;
    long long query_begin, query_end;   /* linear queries [begin, end) of the B*out_h*out_w to run (row bands of one
                                           pair: rows [r0, r1) are queries [r0*out_w, r1*out_w)); 0, 0 = all.  Rows of
                                           out_nbr outside the range are not touched. */
} elo_search_desc;
int elo_multi_search(const elo_search_desc *specs, int nspec, void *stream);

/* set-conv (utils/pointnet_util.py:179-250 down_conv, mlp2 = None) and the first half of set-upconv
 * (:254-298): random-K neighbours of each query in (xyz2, feat2); rows [q_k - p, feat2_k] -> 2..3
 * layers (each 64 or 128 wide) -> * mask -> max over K -> out (B, out_h*out_w, cout[last]).
 * nsets = 2 runs two parameter sets (feat2 / scan order / weights / out) over the same geometry in
 * one launch -- the model's two up_conv calls per level (pwclo_model.py:247-251). */
typedef struct {
    int batch_size;           /* samples per parameter set */
    elo_queries queries;
    int nsets;
    int set_batch_offset[2];  /* first sample of each set in the tensors (0,0 for the two up_conv calls;
                                 0,B when the two frames of the siamese pyramid are stacked as (2B,...)) */
    elo_window window[2];
    int feat_channels;        /* channels of feat2, multiple of 4 */
    int num_layers;
    int cout[3];
    const float *xyz1;        /* (B, H, W, 3) query image */
    const float *xyz2;        /* (B, small_h, small_w, 3) searched grid */
    const float *feat2[2];    /* (B, small_h, small_w, feat_channels) */
    const float *weights[2];
    float *out[2];
    int *dbg_nbr[2];          /* optional (B, n, K): selected linear cell of the searched grid, -1 = masked */
    const int *nbr[2];        /* optional: tables from elo_multi_search; when given, the kernel does not search */
    long long query_begin, query_end;   /* queries [begin, end) of each set's batch_size*out_h*out_w to run; 0, 0 = all */
} elo_group_mlp_desc;
int elo_group_mlp_max(const elo_group_mlp_desc *desc, void *stream);

/* attentive cost volume (utils/pointnet_util.py:33-149), mlp1 = [128,64,64], mlp2 = [128,64].
 * elo_cost_volume_1: select-K (window_q, distance as given -- the reference passes 1000) of frame 2
 *   around each frame-1 pixel -> CV_0..2, CV_xyz, sum_CV_0..1 -> masked softmax over K -> stage1_out
 *   (B, H*W, 64).
 * elo_cost_volume_2: random-K (window_p) self-neighbours in frame 1 -> sum_xyz_encoding,
 *   sum_cost_volume_0..1 -> masked softmax over K of the gathered stage-1 features -> out (B, H*W, 64).
 * Stage 2 reads stage 1 at neighbouring pixels, hence two launches. */
typedef struct {
    int batch_size, H, W, C;  /* both frames are (B, H, W, .); C = feature channels, multiple of 4 */
    elo_window window_q;      /* stage 1: kernel_size2, nsample_q */
    elo_window window_p;      /* stage 2: kernel_size1, nsample, distance */
    const float *xyz1, *xyz2; /* (B, H, W, 3) warped frame 1, frame 2 */
    const float *f1, *f2;     /* (B, H, W, C) */
    const float *weights_1, *weights_2;
    float *stage1_out;        /* (B, H*W, 64) written by stage 1, read by stage 2 */
    float *out;               /* (B, H*W, 64) */
    int *dbg_nbr_q, *dbg_nbr_p;
    const int *nbr_q, *nbr_p; /* optional: tables from elo_multi_search (stage 1 / stage 2) */
    long long query_begin, query_end;   /* pixels [begin, end) of the B*H*W to run in THIS call (stage 2 of a row band
                                           needs stage 1 on one more row of its 3-row window either side); 0, 0 = all */
} elo_cost_volume_desc;
int elo_cost_volume_1(const elo_cost_volume_desc *desc, void *stream);
int elo_cost_volume_2(const elo_cost_volume_desc *desc, void *stream);

/* per-point MLP chains on the concatenation of up to three (rows, C_i) tensors
 * (utils/pointnet_util.py:153-175 flow_predictor; :303-311 the second half of up_conv).  A second
 * phase may use the first phase's result as one of its sources without leaving the chip. */
typedef struct {
    int num_sources;
    int channels[3];          /* multiples of 4 */
    int from_previous[3];     /* 1: this source is the previous phase's output */
    int num_layers;
    int cout[3];              /* 64 or 128 */
    const float *src[2][3];   /* [set][source], (rows, channels[i]) */
} elo_row_mlp_phase;
typedef struct {
    long long rows;
    int nsets;
    int num_phases;
    elo_row_mlp_phase phase[2];
    const float *weights[2];
    float *out[2];            /* (rows, cout of the last layer of the last phase) */
    float *out_phase0[2];     /* optional: phase 0's result */
} elo_row_mlp_desc;
int elo_row_mlp(const elo_row_mlp_desc *desc, void *stream);

/* set-conv for the narrow pyramid layers (pwclo_model.py:126-135), one warp per 32/K queries, MLP in
 * registers.  Same descriptor as elo_group_mlp_max with nsets = 1, num_layers = 3, xyz2 == xyz1,
 * feat2[0] may be NULL (all-zero input features, pwclo_model.py:69-70); supported (feat_channels;
 * cout; K): (3; 8,8,16; 32), (16; 16,16,32; 32), (32; 32,32,64; 16).  weights[0] here is the PLAIN
 * packing W1[(3+Cf)][c1], b1, W2[c1][c2], b2, W3[c2][c3], b3 (batch-norm folded).  nsets = 2 runs two
 * sample ranges with their own scan orders (window[1].random_hw) but the same weights / tensors. */
int elo_set_conv_small(const elo_group_mlp_desc *desc, void *stream);

/* PreProcess / pose warp fused with ProjectPC2SphericalRing (model_util.py:181-292, 346-445;
 * pwclo_model.py:213-232).  mode 0: project the points as they are; mode 1: 35 m crop, optional
 * 4x4 augmentation T (B,4,4) (NULL = none) applied to the samples whose T_apply flag is set (NULL =
 * all), then * valid as the reference does (signed zeros survive); mode 2: p' = q p q^-1 + t with
 * per-sample q (B,4) (w,x,y,z) and t (B,3), empty points stay empty.  Per cell the nearest point
 * wins (equal ranges accumulate); out_xyz (B,H,W,3), out_feat (B,H,W,C) are fully written.
 * points: sample b, point n at points + b*batch_stride + n*point_stride (floats) -- lets the kernel
 * read xyz straight out of the reference's (B, 2N, 6) input.
 * cellmin: (B,H,W) uint64 scratch, all ones before the first call; state: 2 x uint32, zero before the first call.
 * Both persist between calls on the same images (an epoch in state[0] tags the minima, so nothing is cleared);
 * calls that may overlap on different streams need their own pair. */
typedef struct {
    int batch_size, num_points, H, W, C, mode;
    const float *points;
    long long point_stride, batch_stride;
    int inner_batch;          /* 0: sample s at s*batch_stride.  >0: s = f*inner_batch + b lives at
                                 b*batch_stride + f*outer_stride (frame-major view of a (B,2N,6) cloud) */
    long long outer_stride;
    const float *feat;        /* (B, num_points, C) or NULL */
    const float *T, *q, *t;
    float pi, az_res, v_res, v_off;   /* fp32 constants of model_util.py:204-210 */
    unsigned long long *cellmin;
    unsigned *state;
    float *out_xyz, *out_feat;
    float *out_points;        /* optional (B, num_points, 3): the transformed points */
    const int *T_apply;       /* optional (B): mode 1 multiplies sample b by T[b] only if T_apply[b] != 0 */
    int *out_cell;            /* optional (B, num_points): cell (row*W+col) if the point is a nearest point of its
                                 cell, else -1 -- lets a differentiable scatter be rebuilt on top (training) */
    int *point_keys;          /* optional scratch (B, num_points, 2) int32: the binning pass leaves every point's (cell,
                                 range bits) here and the scatter pass reads them instead of transforming and binning
                                 the point a second time (atan2, asin, sqrt, divisions) */
} elo_project_desc;
int elo_project(const elo_project_desc *desc, void *stream);

/* softmax_valid (model_util.py:319-343) + conv1d heads + pose composition (pwclo_model.py:194-208,
 * 262-280).  feature / weight (B,N,64); xyz (B,N,3) marks valid points (not exactly zero).
 * Outputs q_out (B,4) (composed, un-normalised as the next level consumes it), t_out (B,3),
 * q_norm_out (B,4) (pwclo_model.py:427-430), pooled_out (B,64) optional.
 * w_big == NULL: only the pooling (softmax_valid) is done and pooled_out is required.
 * partial: (B, num_slices, 192) scratch; counter: (B) uint32, zero before the first call. */
typedef struct {
    int batch_size, num_points, num_slices, has_coarse;
    const float *feature, *weight, *xyz;
    const float *w_big, *b_big, *w_q, *b_q, *w_t, *b_t;
    const float *q_coarse, *t_coarse;
    float *partial;
    unsigned *counter;
    float *q_out, *t_out, *q_norm_out, *pooled_out;
} elo_pose_head_desc;
int elo_pose_head(const elo_pose_head_desc *desc, void *stream);

/* Strided xyz pyramid (pwclo_model.py:88-114, get_selected_idx + gather_nd): level l of 4 keeps pixel
 * (i*stride_h[l], j*stride_w[l]) of xyz_in (samples,H,W,3) for i < out_h[l], j < out_w[l]; strides are
 * cumulative with respect to xyz_in.  out: 4 device pointers (host array), (samples,out_h[l],out_w[l],3). */
int elo_pyramid_xyz(int samples, int H, int W, const int *out_h, const int *out_w, const int *stride_h,
                    const int *stride_w, const float *xyz_in, float *const *out, void *stream);

/* Ground-truth pose as the network's (q, t) parametrisation (model_util.py:386-426): per sample
 * T = T_trans T_gt (aug_frame 2) or T_gt T_trans_inv (aug_frame 1); q from the zyx Euler angles of R,
 * t = T[:3,3].  T_* are (B,4,4) row-major, aug_frame (B) int32 (NULL = all 2); q_gt (B,4), t_gt (B,3). */
int elo_gt_pose(int batch_size, const float *T_gt, const float *T_trans, const float *T_trans_inv,
                const int *aug_frame, float *q_gt, float *t_gt, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ELO_B200_H */
