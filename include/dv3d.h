/*
 * dv3d.h — C ABI of lib3dvnet_b200.so: the B200-native hot path of 3DVNet.
 *
 * The reference (alexrich021/3dvnet) is pure Python and has no FFI of its own: its hot
 * path is a sequence of third-party library calls made from mv3d/subnetworks/*.py,
 * mv3d/lightningmodel.py and mv3d/utils.py.  Each entry point below replaces one such
 * call sequence; the comment above it cites the reference lines it stands in for.  The
 * Python host side that mirrors the reference's module interface (3dvnet_b200/mv3d/…)
 * binds these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - fp32 row-major tensors, shapes in the comments, innermost dimension last;
 *   - `stream` is a cudaStream_t; all work is enqueued on it, no call synchronises
 *     unless documented ("syncs");
 *   - return 0 on success, a negative DV3D_E* code otherwise; dv3d_last_error() returns a
 *     thread-local human-readable message for the last failure;
 *   - workspaces are caller-allocated; *_workspace_bytes() functions size them.
 */
#ifndef DV3D_H
#define DV3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DV3D_OK 0
#define DV3D_EINVAL (-1)   /* bad argument / unsupported shape */
#define DV3D_ECUDA (-2)    /* CUDA runtime error, see dv3d_last_error() */
#define DV3D_ENOSPC (-3)   /* caller-provided capacity too small */

const char* dv3d_last_error(void);
/* ABI version of this header; bump on any signature change. */
int dv3d_abi_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
long long dv3d_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Layout helper: NCHW -> NHWC (channels-last) copy of the quarter-resolution feature maps,
 * so that one bilinear tap is one contiguous C*4-byte line.  src [n,C,H*W] -> dst [n,H*W,C].
 */
int dv3d_nchw_to_nhwc(const float* src, float* dst, int n, int C, int HW, void* stream);

/* ------------------------------------------------------------------------------------
 * Per-image camera table used by the two warp kernels: out [n_imgs, 36] =
 *     Kinv (9) | P = K [R|t] (12) | R (9) | t (3) | 0 (3)
 * with the reference's own fp32 arithmetic: P as torch.bmm forms it for 3x3 @ 3x4 operands
 * (plain loop, every product and sum rounded; mvsnet.py:196-197) and Kinv is what torch.inverse returns for a zero-skew pinhole K
 * ([[1/fx,0,-(cx/fx)],[0,1/fy,-(cy/fy)],[0,0,1]]; any other K: fp64 adjugate, rounded once).
 * The kernels then evaluate X = R^T (Kinv [u z, v z, z] - t) and q = P_src [X;1] in the
 * reference's operation order, so sample positions and back-projected points are
 * bit-identical to the CPU PyTorch path (utils.py:102-106, mvsnet.py:199-206,
 * lightningmodel.py:138-160).
 *   rotmats [n_imgs,3,3], tvecs [n_imgs,3], K [n_imgs,3,3] (world->camera, full-res K)
 * Replaces: torch.inverse + 3 torch.bmm per call site.
 */
int dv3d_camera_tables(const float* rotmats, const float* tvecs, const float* K, int n_imgs, float* out,
                       void* stream);

/* ------------------------------------------------------------------------------------
 * Fused plane-sweep warp + variance  (mvsnet.py:187-216 with utils.py:86-108).
 *   feats_nhwc [n_imgs,Hf,Wf,C]   quarter-resolution features, channels-last, C == 32
 *   cams       [n_imgs,36]        from dv3d_camera_tables
 *   ref_img    [n_ref]            image index of every reference (int32, ascending)
 *   edge_rowptr[n_ref+1], edge_src[E]   CSR of the edges of every reference (int32), the edges
 *                                 of a reference in their original order (it is the summation order)
 *   depth hypotheses z_d = linspace(depth_start, depth_start+(D-1)*depth_interval, D)
 *   plane lattice u_j = linspace(0,W-1,w), v_i = linspace(0,H-1,h); grid normalised by the
 *   FULL image size (W-1,H-1) and un-normalised by (Wf-1,Hf-1) as grid_sample does;
 *   bilinear, zero padding per tap, z = |q_z| + 1e-8; divisor = number of edges.
 *   x_var [n_ref,C,D,h,w]         var = E[x^2] - E[x]^2, sums in edge order like scatter-add
 * No x_vox[E,C,D,h,w] is ever materialised.
 */
int dv3d_planesweep_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                        const int* ref_img, const int* edge_rowptr, const int* edge_src, int n_ref,
                        double depth_start, double depth_interval, int D, int h, int w, int H, int W, float* x_var,
                        void* stream);
/* Arithmetic of dv3d_planesweep_var: 0 = exact (default: the reference's fp32 operation chain replayed instruction
 * for instruction, x_var bit-identical to the CPU PyTorch path), 1 = fast (opt-in, DV3D_WARP=fast: affine-in-depth
 * projection + reciprocal, x_var within ~2e-5 of its scale, 12 % faster - measured, see csrc/planesweep.cu).
 * The point-level kernels (dv3d_points_var) always use the exact chain. Process-wide. */
int dv3d_set_warp_mode(int mode);
int dv3d_get_warp_mode(void);

/* ------------------------------------------------------------------------------------
 * Point-level back-projection + re-projection warp + variance, 1 or 2n+1 hypotheses per
 * pixel (lightningmodel.py:132-174 with n_side = 0, and :187-235 with n_side = 3).
 *   depth    [n_ref,h,w];  cams / ref_img / CSR as above
 *   hypothesis i (i=-n..n) has depth d + (float)(i*offset), offset a double like the python float
 *   pts_out  [n_ref*h*w, n_hyp, 3]   world points
 *   feat_out [n_ref*h*w, rows_per_point >= n_hyp, feat_stride] variance feature written at
 *            channel offset feat_off (lets the caller write straight into the decoder operand)
 */
int dv3d_points_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                    const int* ref_img, const int* edge_rowptr, const int* edge_src, const float* depth, int n_ref,
                    int h, int w, int H, int W, int n_side, double offset, float* pts_out, float* feat_out,
                    int rows_per_point, int feat_stride, int feat_off, void* stream);
/* Backward of the two warp + variance entry points above w.r.t. the source feature maps (the reference trains
 * through F.grid_sample + two scatter means, mvsnet.py:209-216 / lightningmodel.py:165-169; its sampling grid is
 * built under no_grad): grad_feats_nhwc [n_imgs,Hf,Wf,32] must be ZERO on entry and receives
 * sum over (pixel, plane, edge) of g * (2/n_edges) (x_e - mean) * bilinear tap weight, accumulated with vector
 * atomics (reproducible to fp32 rounding, like ATen's CUDA grid_sampler backward).  Everything but the gradient
 * arguments is as in the forward call. */
int dv3d_planesweep_var_backward(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                                 const int* ref_img, const int* edge_rowptr, const int* edge_src, int n_ref,
                                 double depth_start, double depth_interval, int D, int h, int w, int H, int W,
                                 const float* grad_x_var, float* grad_feats_nhwc, void* stream);
int dv3d_points_var_backward(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                             const int* ref_img, const int* edge_rowptr, const int* edge_src, const float* depth,
                             int n_ref, int h, int w, int H, int W, int n_side, double offset, const float* grad_feat,
                             int rows_per_point, int feat_stride, int feat_off, float* grad_feats_nhwc, void* stream);


/* ------------------------------------------------------------------------------------
 * CostRegNet layers, inference mode (mvsnet.py:18-36,133-163).  NCDHW fp32.
 * BatchNorm is folded by the caller into per-channel (scale, shift); y = relu(conv*scale+shift) [+ skip].
 *   conv:   weight [Cout,Cin,3,3,3], stride 1 or 2, padding 1
 *   deconv: weight [Cin,Cout,3,3,3], stride 2, padding 1, output_padding 1 (output dims = 2x input)
 *   skip (optional, may be NULL): added AFTER the ReLU (x = conv4 + conv7(x), mvsnet.py:159-161)
 */
int dv3d_conv3d_bn_relu(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                        const float* scale, const float* shift, int Cout, int stride, const float* skip,
                        float* y, void* stream);
/* The first CostRegNet layer (Cin 32, Cout 8, stride 1, no skip) runs on tcgen05 with 3xTF32 arithmetic
 * (csrc/conv3d_tc.cu); mode 1 (or DV3D_CONV3D=ffma) keeps every layer on the fp32 CUDA-core kernels. */
int dv3d_set_conv3d_mode(int mode);
int dv3d_get_conv3d_mode(void);
/* profiling aid (tools/conv3d_phases.py): 64 int64 clock stamps per CTA (148 CTAs at most) of the next launches, or NULL */
int dv3d_conv3d_set_timing_buffer(void* device_buffer);
int dv3d_deconv3d_bn_relu(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                          const float* scale, const float* shift, int Cout, const float* skip, float* y,
                          void* stream);
/* prob conv (8->1, bias) + softmax(-x) over D + expectation of the plane depths
 * (mvsnet.py:152,162,219-227).  x [n,Cin,D,H,W]; weight [1,Cin,3,3,3]; x_reg_out optional
 * [n,D,H,W]; depth_out [n,H,W].  depth values follow torch.linspace's fp32 formula. */
int dv3d_prob_softargmin(const float* x, int n, int Cin, int D, int H, int W, const float* weight, float bias,
                         float depth_start, float depth_end, float* x_reg_out, float* depth_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Voxelisation (utils.py:38-64 incl. torch_geometric voxel_grid -> torch_cluster
 * grid_cluster; bit-exact index math, anchors in ascending voxel-id order like torch.unique).
 *
 * Two calls, each SYNCS the stream once:
 *   dv3d_voxel_grid   reduces the bounding box / batch count and derives the grid on the host
 *                     (both of the reference's formulas: grid_cluster's trunc(.)+1 cells and
 *                     voxelize's ceil(.) grid_size, utils.py:41);
 *   dv3d_voxelize     marks occupied cells in a bitmap, ranks them with a scan, and emits
 *                     anchors + the point->anchor map; reads the anchor count back.
 */
typedef struct dv3d_voxel_grid_t {
    float bbox_min[3];
    float bbox_max[3];
    float edge_len;
    long long n_cells[3];   /* torch_cluster: trunc((max-min)/e) + 1 */
    long long grid_size[3]; /* utils.py:41:   ceil((max-min)/e)      */
    long long n_batch;      /* batch.max() + 1 */
    long long total_cells;  /* prod(n_cells) * n_batch */
} dv3d_voxel_grid_t;

/* pts [N,3] f32, batch [N] int64 (device); scratch64: >= 64 bytes of device memory */
int dv3d_voxel_grid(const float* pts, const long long* batch, long long N, float edge_len,
                    dv3d_voxel_grid_t* grid_host, void* scratch64, void* stream);
size_t dv3d_voxelize_workspace_bytes(const dv3d_voxel_grid_t* grid_host, long long N);
/*   out: *n_anchors_host; anchor_pts [cap,3] f32; anchor_idx3d [cap,3] int32 (per-batch min
 *        subtracted, utils.py:61-62); anchor_batch [cap] int64; point_anchor [N] int32
 *        (row 0 of anchor_pts_edges; row 1 is arange(N)).
 *   DV3D_ENOSPC if workspace or cap is too small. */
int dv3d_voxelize(const float* pts, const long long* batch, long long N, const dv3d_voxel_grid_t* grid_host,
                  void* workspace, size_t workspace_bytes, long long cap, long long* n_anchors_host,
                  float* anchor_pts, int* anchor_idx3d, long long* anchor_batch, int* point_anchor, void* stream);

/* ------------------------------------------------------------------------------------
 * Dense contractions of the volumetric refinement (every nn.Linear, MinkowskiConvolution and
 * Conv1d of scenemodeling.py / refinement.py) run through one gather-GEMM.  Each entry point
 * below takes the layer's weight twice:
 *   W / weight_*  fp32 [K, N] row-major            -> fp32 CUDA-core kernel (verification path)
 *   W_packed      image made by dv3d_gemm_pack_weights -> tcgen05 tensor-core kernel (TMEM
 *                 accumulators); when non-NULL it wins and W may be NULL.  K % 32 == 0.
 * dv3d_set_gemm_precision: 1 = 3xTF32 split accumulation (fp32-grade, default), 2 = TF32.
 */
size_t dv3d_gemm_pack_bytes(int Ktot, int N);
int dv3d_gemm_pack_weights(const float* W, int Ktot, int N, void* packed, void* stream);
int dv3d_set_gemm_precision(int mode);
int dv3d_get_gemm_precision(void);
/* Persistent variant of the tcgen05 gather-GEMM for single-slice launches (pair-major sparse convolution, Linear,
 * Conv2d rows): -1 = never, 0 = automatic (more 128-row tiles than SMs; default), 1 = whenever the launch qualifies.
 * Results are bit-identical in all modes. Process-wide; for tests and A/B measurements. */
int dv3d_set_gemm_persistent(int mode);
/* pair-major sparse-convolution GEMM: 1 = weight-stationary kernel (default), 0 = the general gather-GEMM kernels
 * (A/B measurements; also DV3D_PAIR_WS=0) */
int dv3d_set_pair_gemm_mode(int weight_stationary);
/* profiling aid (tools/gemm_phases.py): device buffer of (grid.x * grid.y) * 8 int64 clock stamps written by every
 * tcgen05 gather-GEMM launch on the current device; NULL switches it off */
int dv3d_gemm_set_timing_buffer(void* device_buffer);

/* ------------------------------------------------------------------------------------
 * PointNet (scenemodeling.py:116-144).
 *   input rows [pts - anchor_pts[seg] | pts_feat | 0...] (lightningmodel.py:182): out [N,out_ld],
 *   out_ld >= 3 + C (padded so that the GEMM K extent is a multiple of 32)
 */
int dv3d_pointnet_input(const float* pts, const float* pts_feat, int feat_ld, const float* anchor_pts,
                        const int* seg, long long N, int C, int out_ld, float* out, void* stream);
/*   y = [relu](x) W + b on rows x = [x_a | pool[seg]] (scenemodeling.py:130-138).
 *   x_a [N,lda] (first Ca columns used, Ca % 16 == 0); pool [n_seg,Cb] gathered through seg [N]
 *   int32 (NULL with Cb = 0); weight_kn [Ca+Cb, Cout] = nn.Linear.weight transposed;
 *   bias [Cout]; Cout in {64,128}; y [N,Cout] */
int dv3d_linear(const float* x_a, int Ca, int lda, const float* pool, const int* seg, int Cb, long long N,
                const float* weight_kn, const void* W_packed, const float* bias, int Cout, int relu_input, float* y,
                void* stream);
/* segment max with empty segments = 0 (torch_scatter 'max'): out [n_seg,C] */
/* dv3d_linear whose epilogue also max-pools the output rows per segment (scatter max of scenemodeling.py:129 fused
 * into the layer that produces its operand): pool_out[pool_seg[m], :] = max(pool_out[...], y[m, :]), row pitch Cout.
 * pool_out must be pre-filled with 0xFF bytes (cudaMemsetAsync) and every segment must own at least one row; the
 * result is the exact maximum (order independent), identical to dv3d_segment_max on y. */
int dv3d_linear_pool(const float* x_a, int Ca, int lda, const float* pool, const int* seg, int Cb, long long N,
                     const float* weight_kn, const void* W_packed, const float* bias, int Cout, int relu_input, float* y,
                     float* pool_out, const int* pool_seg, void* stream);
int dv3d_segment_max(const float* x, const int* seg, long long N, int C, long long n_seg, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Sparse 3D-UNet building blocks (scenemodeling.py:16-44,78-113,147-237; MinkowskiEngine
 * 0.5 semantics per SURVEY.md A.4).
 *
 * A coordinate level is: coords [n,4] int32 (batch,x,y,z) in ascending (batch,z,y,x)
 * order, its tensor stride, and an open-addressed hash table (key -> row).
 */
/* [batch | idx3d] rows of the finest level (scenemodeling.py:192) */
int dv3d_make_coords(const int* idx3d, const long long* batch, long long n, int* coords, void* stream);
size_t dv3d_hash_bytes(long long n_rows);
/* build the table of a level; *err_flag (device int, caller-zeroed) is set on out-of-range coordinates */
int dv3d_hash_build(const int* coords, long long n, void* table, size_t table_bytes, int* err_flag, void* stream);
/* n_tables (<= 4) tables in two launches (the coordinate maps MinkowskiEngine's manager creates for the strided
 * convolutions of scenemodeling.py:160-162); coords / n / tables / table_bytes are HOST arrays of device pointers */
int dv3d_hash_build_batch(const int* const* coords, const long long* n, void* const* tables, const size_t* table_bytes,
                          int n_tables, int* err_flag, void* stream);
/* coarser level of a stride-2 convolution: unique(floor(c / new_stride) * new_stride) in
 * (batch,z,y,x) order.  dim_* bound the finest-level index range (cells per axis).  SYNCS once. */
size_t dv3d_coarsen_workspace_bytes(int dim_x, int dim_y, int dim_z, int n_batch, int new_stride);
int dv3d_coarsen(const int* coords, long long n, int new_stride, int dim_x, int dim_y, int dim_z, int n_batch,
                 void* workspace, size_t workspace_bytes, long long cap, int* coarse_coords,
                 long long* n_coarse_host, void* stream);
/* the same in two halves, so that several levels can be coarsened from the finest coordinates
 * (floor(floor(c/2)*2/4)*4 = floor(c/4)*4) with ONE sync: enqueue every level, then finish each */
int dv3d_coarsen_enqueue(const int* coords, long long n, int new_stride, int dim_x, int dim_y, int dim_z, int n_batch,
                         void* workspace, size_t workspace_bytes, long long cap, int* coarse_coords, void* stream);
/* n_levels (<= 4) coarser levels (output coordinates of the stride-2 convolutions, scenemodeling.py:160-162,194-204)
 * from the same finest coordinates with four launches and one memset when the workspaces are contiguous;
 * new_strides / workspaces / workspace_bytes / coarse_coords are HOST arrays; finish each with dv3d_coarsen_finish */
int dv3d_coarsen_enqueue_batch(const int* coords, long long n, const int* new_strides, int n_levels, int dim_x, int dim_y,
                               int dim_z, int n_batch, void* const* workspaces, const size_t* workspace_bytes,
                               long long cap, int* const* coarse_coords, void* stream);
int dv3d_coarsen_finish(const void* workspace, int new_stride, int dim_x, int dim_y, int dim_z, int n_batch,
                        long long cap, long long* n_coarse_host, void* stream);
/* the counts of several levels enqueued by dv3d_coarsen_enqueue_batch with ONE read-back (HOST arrays) */
int dv3d_coarsen_finish_batch(const void* const* workspaces, const int* new_strides, int n_levels, int dim_x, int dim_y,
                              int dim_z, int n_batch, long long cap, long long* n_coarse_host, void* stream);
/* kernel map: nbr[o*27+k] = row of coords_out[o] + offset_k*step in the input level, or -1
 * (offset_k x fastest).  k3s1: same level, step = ts.  k3s2: out = coarse, in = fine table,
 * step = ts_fine.  transposed k3s2: out = fine, in = coarse table, step = -ts_fine
 * (fine_j = coarse_i + offset_k*ts_fine, same W[k], no flip). */
int dv3d_kernel_map(const int* coords_out, long long n_out, const void* table_in, size_t table_bytes, int step,
                    int* nbr, void* stream);
/* up to 16 kernel maps (all the maps of a scene) with one launch; HOST arrays of device pointers / sizes */
int dv3d_kernel_map_batch(const int* const* coords_out, const long long* n_out, const void* const* table_in,
                          const size_t* table_bytes, const int* step, int* const* nbr, int n_maps, void* stream);
/* Row ranges of one sparse level for `world` GPUs with equal WORK instead of equal row counts: the occupied 3x3x3
 * neighbours (offset `step`) of every sample_stride-th row are counted through the level's own hash table and the
 * prefix sum is cut into `world` parts.  bounds_dev[0..world] (device ints): rank k owns rows
 * [bounds[k], bounds[k+1]); interior bounds are multiples of 128.  Deterministic: identical on every rank that holds
 * the same level.  scratch: dv3d_balanced_row_bounds_scratch_bytes(n, sample_stride). */
size_t dv3d_balanced_row_bounds_scratch_bytes(long long n, int sample_stride);
int dv3d_balanced_row_bounds(const int* coords, long long n, const void* table, size_t table_bytes, int step, int world,
                             int sample_stride, void* scratch, int* bounds_dev, void* stream);
/* out[o] = sum_k feat[nbr[o,k]] @ W[k]   (W [27,Cin,Cout], ME layout); optional fused per-row
 * GroupNorm (gn_weight/gn_bias [Cout], 16 channels per group, eps 1e-5), optional residual
 * add (before the ReLU) and ReLU — the SparseResidual3d / down / up blocks.  Cin % 16 == 0
 * (% 32 for the tensor-core kernel), Cout in {64,128}.
 * workspace (optional, tensor-core kernel only): lets a small level fill the GPU by splitting
 * the 27 offsets over several CTAs per 128-row tile; partials are added in a fixed order by the
 * last CTA of the tile.  dv3d_sparse_conv_workspace_bytes(max Cout) sizes it for any level; it
 * must be 256-byte aligned and ZERO before its first use - every launch leaves its counters
 * zero again, so one buffer serves all layers of a scene on one stream. */
size_t dv3d_sparse_conv_workspace_bytes(int Cout);
int dv3d_sparse_conv(const float* feat, long long n_in, int Cin, const int* nbr, long long n_out, const float* W,
                     const void* W_packed, int Cout, const float* gn_weight, const float* gn_bias,
                     const float* residual, int relu, void* workspace, size_t workspace_bytes, float* out,
                     void* stream);
/* Pair-major variant for SPARSELY connected levels (csrc/sparse_pairs.cu): the rows of the GEMM
 * are the existing (output row, offset) pairs grouped by offset - what MinkowskiEngine's in/out
 * kernel maps are - so no tensor-core work is spent on absent neighbours:
 *   plan    built once per kernel map (no sync; the maps of a scene in one call): pairs of every offset in ascending output row,
 *           padded to 128-row tiles;
 *   GEMM    P[slot,:] = feat[in_row(slot),:] @ W[offset(tile)]  (tcgen05 kernel, weight block per tile);
 *   reduce  out[m,:] = epilogue(sum over offsets ascending of P[slot(m,k),:]) - fixed order, bit-reproducible.
 * dv3d_pair_plan_counts reads the tile / pair counts of several plans back with ONE sync;
 * dv3d_sparse_conv_prefers_pairs is the library's choice between the two variants (callers that
 * must agree bit for bit - the engine and the composed path - both ask it). */
size_t dv3d_pair_plan_bytes(long long n_out);
/* plans n_maps (<= 16) kernel maps with two launches; arrays are HOST arrays of device pointers / sizes */
int dv3d_pair_plan_build(const int* const* nbrs, const long long* n_outs, void* const* plans, const size_t* plan_bytes,
                         int n_maps, void* stream);
int dv3d_pair_plan_counts(const void* const* plans, int n_plans, long long* n_tiles_host, long long* n_pairs_host,
                          void* stream);
int dv3d_sparse_conv_prefers_pairs(long long n_out, long long n_tiles);
size_t dv3d_sparse_conv_pairs_workspace_bytes(long long n_tiles, int Cout);
/* W_packed: dv3d_gemm_pack_weights image of W [27*Cin, Cout] (required: tensor-core path only) */
int dv3d_sparse_conv_pairs(const float* feat, long long n_in, int Cin, const void* plan, long long n_tiles,
                           long long n_out, const void* W_packed, int Cout, const float* gn_weight,
                           const float* gn_bias, const float* residual, int relu, void* workspace,
                           size_t workspace_bytes, float* out, void* stream);
/* 1x1 "feature adjust" on the concatenation [a | b] (ME.cat + k=1 conv, scenemodeling.py:206)
 * followed by GroupNorm + ReLU: W [Ca+Cb, Cout]. */
int dv3d_concat_linear_gn_relu(const float* a, int Ca, const float* b, int Cb, long long n, const float* W,
                               const void* W_packed, int Cout, const float* gn_weight, const float* gn_bias,
                               float* out, void* stream);
/* voxel positions of a level (scenemodeling.py:211-226): origin[b] = anchor_pts[first voxel of b]
 * - idx3d[first] * res; pts = coord * res + origin[batch]; optional int64 views of idx / batch */
int dv3d_batch_origin(const float* anchor_pts, const int* idx3d, const long long* batch, long long n, float res,
                      float* origin, void* stream);
int dv3d_level_points(const int* coords, long long n, const float* origin, float res, float* pts,
                      long long* idx_out, long long* batch_out, void* stream);

/* ------------------------------------------------------------------------------------
 * PointFlow hypothesis decoder (refinement.py:28-44, lightningmodel.py:238-241).
 * The decoder operand is [n_pts, rows_per_point = 8, ld]: 7 hypotheses + one all-zero row,
 * channels [L0 64 | L1 128 | L2 128 | var 32] (refinement.py:41 prepends each level).
 *
 * Trilinear sparse interpolation of one level at the hypothesis points, written at channel
 * offset out_off (missing voxels contribute 0, no renormalisation).
 *   pts [n_pts*n_hyp,3] world points; origin [n_batch,3] from dv3d_batch_origin (equals the
 *   scatter-min of the level's voxel positions, refinement.py:33); res = level voxel size
 *   (stride * base edge), stride = tensor stride; pts_batch [n_pts] int64.
 */
int dv3d_sparse_interp(const float* pts, const long long* pts_batch, long long n_pts, int n_hyp, int rows_per_point,
                       const float* origin, float res, int stride, const void* table, size_t table_bytes,
                       const float* feat, int C, float* out, int out_ld, int out_off, void* stream);
/* the same interpolation for up to 4 levels with one launch (HOST arrays of per-level values) */
int dv3d_sparse_interp_batch(const float* pts, const long long* pts_batch, long long n_pts, int n_hyp,
                             int rows_per_point, const float* origin, int n_levels, const float* res, const int* stride,
                             const void* const* table, const size_t* table_bytes, const float* const* feat, const int* C,
                             const int* out_off, float* out, int out_ld, void* stream);
/* Conv1d(k=3,pad=1,no bias)+BN(folded scale/shift)+ReLU over the hypothesis axis on the padded
 * layout: x [n_pts*8, ldx] -> y [n_pts*8, ldy] (row 7 of every point written as 0);
 * weight_tkn [3, Cin, Cout] = torch weight [Cout,Cin,3] permuted (2,1,0). */
int dv3d_conv1d_bn_relu(const float* x, long long n_pts, int rows_per_point, int Cin, int ldx,
                        const float* weight_tkn, const void* W_packed, const float* scale, const float* shift,
                        int Cout, float* y, int ldy, void* workspace, size_t workspace_bytes, void* stream);
/* last Conv1d (Cin->1, bias; weight [1,Cin,3] torch layout) + softmax over hypotheses + expected
 * offset sum_i p_i * linspace(-n*offset, n*offset)_i; prob_out optional [n_pts,n_hyp]; offset_out [n_pts] */
int dv3d_decoder_head(const float* x, long long n_pts, int n_hyp, int rows_per_point, int Cin, int ldx,
                      const float* weight, float bias, double offset, float* prob_out, float* offset_out,
                      void* stream);

/* The whole decoder stack of a PointFlow pass as ONE tcgen05 kernel (csrc/decoder_fused.cu): the three
 * Conv1d(k=3)+BN+ReLU layers, the 1-channel Conv1d head, the softmax over the 7 hypotheses and the expected
 * offset (refinement.py:17-25,42-44; lightningmodel.py:238-241).  A 128-row tile is 16 whole points, so the
 * activations never leave the SM between layers.  x [n_pts*8, ldx] as for dv3d_conv1d_bn_relu (the 8th row of a
 * point is treated as zero, whatever it holds); hidden width must be 128.
 *   dv3d_decoder_pack_weights: weight_tkn [3, Cin, 128] -> the tap-stationary shared-memory images
 *   (dv3d_decoder_pack_bytes(Cin) bytes; Cin % 16 == 0).
 *   W_packed / scale / shift: HOST arrays of 3 device pointers (layer 1..3); precision 1 = 3xTF32, 2 = TF32.
 *   prob_out [n_pts,7], offset_out [n_pts] optional; depth_accum [n_pts] optional: depth += offset
 *   (eval-3dvnet.py:99) in the same kernel. */
size_t dv3d_decoder_pack_bytes(int Cin);
/* profiling aid: device buffer of n_tiles * 8 int64 clock stamps written by every dv3d_decoder_fused launch
 * (NULL switches it off; tools/decoder_phases.py) */
int dv3d_decoder_set_timing_buffer(void* device_buffer);
int dv3d_decoder_pack_weights(const float* weight_tkn, int Cin, int Cout, void* packed, void* stream);
int dv3d_decoder_fused(const float* x, long long n_pts, int rows_per_point, int Cin, int ldx,
                       const void* const* W_packed, const float* const* scale, const float* const* shift,
                       int hidden, const float* head_weight, float head_bias, double offset, int precision,
                       float* prob_out, float* offset_out, float* depth_accum, void* stream);

/* ------------------------------------------------------------------------------------
 * Coarse-to-fine depth upsampling: the step right after the hot path.
 * PropagationNet (upsampling.py:14-36): four Conv2d(3x3, pad 1, no bias)+BN+ReLU, softmax over
 * the 9 output channels, weighted sum of the replicate-padded 3x3 depth neighbourhood; in front of
 * it F.interpolate(mode='nearest') of the depth map (eval-3dvnet.py:101-125).
 * Activations are channels-last rows [n*H*W, ld]; the convolutions run on the tcgen05 gather-GEMM.
 */
/* nbr [n*H*W, 9] int32: row of pixel (y+dy, x+dx), t = (dy+1)*3 + (dx+1); -1 = zero padding */
int dv3d_grid_map_2d(int n, int H, int W, int* nbr, void* stream);
/* x [n*H*W, ld] = [features (C channels of feats_nchw [n,C,H,W]) | depth_up | 0 ...], ld % 32 == 0, ld > C;
 * depth_up [n,H,W] = nearest upsampling of depth_lo [n,h,w] (src = min(floor(dst * (h/H)), h-1) in fp32, as ATen) */
int dv3d_propagation_input(const float* feats_nchw, int C, const float* depth_lo, int n, int h, int w, int H, int W,
                           int ld, float* x, float* depth_up, void* stream);
/* y [M, ldy] (first Cout columns) = relu(conv3x3(x)[.,0:Cout] * scale + shift); x [M, ldx] (first Cin channels,
 * Cin % 32 == 0); W_kn [9*Cin, Cout] with row t*Cin + ci = weight[co][ci][t/3][t%3], Cout in {64,128} (zero-pad
 * the layer's output channels; padded scale/shift = 0); W_packed from dv3d_gemm_pack_weights or NULL */
int dv3d_conv2d3x3_bn_relu_rows(const float* x, long long M, int Cin, int ldx, const int* nbr, const float* W_kn,
                                const void* W_packed, const float* scale, const float* shift, int Cout, float* y,
                                int ldy, void* stream);
/* out [n,H,W] = sum_t softmax(logits[m, 0:9])_t * depth[clamp(y+dy), clamp(x+dx)]; logits rows 16-byte aligned */
int dv3d_propagation_output(const float* logits, int ld, const float* depth, int n, int H, int W, float* out,
                            void* stream);

/* ------------------------------------------------------------------------------------
 * Depth-map fusion (pointcloudfusion_custom.py:10-116), the consumer of the path's output
 * (SURVEY.md §8f.4): for the reference images 0..n_ref-1 of a scene, every pixel is back-projected,
 * re-projected into every OTHER image, compared with the nearest-sampled depth there
 * (|z - z_sample| < z_thresh, inside the image, z > 1e-4) and averaged with the consistent samples.
 *   depths [n_imgs,h,w]; poses [n_imgs,4,4] world->camera; K [n_imgs,3,3] at depth-map resolution
 *   pts_avg [n_ref,h*w,3] = (pts + sum of consistent back-projected samples) / (n_valid + 1)
 *   n_valid [n_ref,h*w] int32; valid [n_ref,h*w] uint8 = n_valid >= n_consistent_thresh
 *   workspace: dv3d_depth_fusion_workspace_bytes(n_imgs) (per-image K, K^-1, P, P^-1, inverted in fp64)
 */
size_t dv3d_depth_fusion_workspace_bytes(int n_imgs);
int dv3d_depth_fusion(const float* depths, const float* poses, const float* K, int n_imgs, int n_ref, int h, int w,
                      float z_thresh, int n_consistent_thresh, void* workspace, size_t workspace_bytes, float* pts_avg,
                      int* n_valid, unsigned char* valid, void* stream);

/* ------------------------------------------------------------------------------------
 * Engine: the whole hot path enqueued by ONE call from native code.
 *
 * The reference drives its hot path from Python, one library call per tensor op
 * (eval-3dvnet.py:58-99 -> lightningmodel.py:124-242); at B200 kernel durations (5-50 us) a
 * per-op interpreter round trip would bound the step.  dv3d_hot_path runs
 *   plane-sweep cost volume -> CostRegNet -> soft-argmin                    (mvsnet.py:176-229)
 *   n_outer x { feature-rich point cloud -> voxelise -> PointNet -> sparse 3D-UNet
 *               (lightningmodel.py:132-185), n_inner x PointFlow pass (:187-242), depth += offset
 *               (eval-3dvnet.py:73-99) }
 * with exactly the kernels and the order of the per-op entry points above (results are
 * bit-identical to composing them), temporaries bump-allocated from a caller-provided arena.
 * It SYNCS the stream 4 times per outer iteration (voxel grid, anchor count, two coarse-level
 * counts: the sizes of the sparse levels are data dependent).
 *
 * Parameters are passed as a table of device pointers the caller has prepared once: folded
 * BatchNorm (scale, shift), nn.Linear / Conv1d weights transposed to [K, N], tensor-core
 * images from dv3d_gemm_pack_weights (Wp; NULL selects the fp32 CUDA-core GEMM).
 */
typedef struct dv3d_conv3d_params_t {
    const float* weight;  /* torch layout: conv [Cout,Cin,3,3,3], deconv [Cin,Cout,3,3,3] */
    const float* scale;   /* folded BatchNorm3d */
    const float* shift;
    int Cin, Cout;
    int kind;             /* 0 = conv stride 1, 1 = conv stride 2, 2 = transposed conv stride 2 */
    int reserved;
} dv3d_conv3d_params_t;

typedef struct dv3d_dense_params_t {
    const float* W;       /* fp32 [K, N] row-major (sparse conv: [27*Cin, Cout]; conv1d: [3*Cin, Cout]) */
    const void* Wp;       /* dv3d_gemm_pack_weights image or NULL */
    const float* a;       /* Linear: unused; Conv1d: folded BN scale; sparse conv / 1x1: GroupNorm weight */
    const float* b;       /* Linear: bias;   Conv1d: folded BN shift; sparse conv / 1x1: GroupNorm bias */
    int K, N;
} dv3d_dense_params_t;

#define DV3D_MAX_LEVELS 3
#define DV3D_MAX_RES 4
typedef struct dv3d_net_params_t {
    /* CostRegNet conv0..conv6, conv7..conv9 (mvsnet.py:133-150), then the prob conv */
    dv3d_conv3d_params_t costreg[10];
    const float* prob_weight;  /* [1,8,3,3,3] */
    float prob_bias;
    /* PointNet fc_pos (input rows zero-padded to pointnet_in_pad), fc1, fc2, fc3, fc4, fc_out */
    int pointnet_in_pad;
    dv3d_dense_params_t pointnet[6];
    /* SparseUNet (scenemodeling.py:147-237): res blocks are [level][block][conv1|conv2] */
    int n_levels;
    int n_res[DV3D_MAX_LEVELS];
    dv3d_dense_params_t res_down[DV3D_MAX_LEVELS][DV3D_MAX_RES][2];
    dv3d_dense_params_t down[DV3D_MAX_LEVELS - 1];      /* k3 s2, level i -> i+1 */
    dv3d_dense_params_t up[DV3D_MAX_LEVELS - 1];        /* transposed k3 s2; up[i] produces level n_levels-2-i */
    dv3d_dense_params_t feat_adj[DV3D_MAX_LEVELS - 1];  /* 1x1 on [up | skip] */
    dv3d_dense_params_t res_up[DV3D_MAX_LEVELS - 1][DV3D_MAX_RES][2];  /* res_up[i] on level n_levels-2-i */
    /* HypothesisDecoder: three Conv1d+BN+ReLU, then the 1-channel head (refinement.py:17-26) */
    dv3d_dense_params_t dec[3];
    const float* dec_head_weight;  /* [1,Cin,3] torch layout */
    float dec_head_bias;
    /* dv3d_decoder_pack_weights images of the three layers: all set -> the engine runs dv3d_decoder_fused */
    const void* dec_fused[3];
} dv3d_net_params_t;

/* Optional stage timing of dv3d_hot_path: CUDA events recorded on the launching stream around
 * the stages below.  dv3d_engine_profile(1) clears the records and enables recording (0 disables);
 * after synchronising the stream, dv3d_engine_profile_read fills ids[] / ms[] with one record per
 * stage execution in launch order and returns their number (or a negative error). */
#define DV3D_STAGE_PLANESWEEP 0   /* planesweep_var_kernel alone */
#define DV3D_STAGE_COSTREG 1      /* the ten CostRegNet layers */
#define DV3D_STAGE_SOFTARGMIN 2
#define DV3D_STAGE_POINTCLOUD 3   /* points_var_kernel, 1 hypothesis */
#define DV3D_STAGE_VOXELIZE 4
#define DV3D_STAGE_POINTNET 5
#define DV3D_STAGE_LEVELS 6       /* coordinate levels + hash tables */
#define DV3D_STAGE_UNET 7         /* kernel maps + sparse convolutions */
#define DV3D_STAGE_FLOW_WARP 8    /* points_var_kernel, 7 hypotheses */
#define DV3D_STAGE_FLOW_INTERP 9
#define DV3D_STAGE_DEC_GEMM0 10   /* dv3d_decoder_fused (whole decoder + depth update); per-layer path: first Conv1d */
#define DV3D_STAGE_DEC_REST 11    /* per-layer path only: two more Conv1d GEMMs + head + depth update */
#define DV3D_STAGE_EXCHANGE 12    /* dv3d_hot_path_sharded only: point rows into the peers' heaps + the barrier */
#define DV3D_STAGE_BARRIER 13     /* sharded only: every cross-GPU barrier (time inside the pointnet / unet stages) */
int dv3d_engine_profile(int enable);
int dv3d_engine_profile_read(int* ids, float* ms, int cap);

/* arena size that dv3d_hot_path needs for these shapes (the voxel bitmaps get a fixed 64 MiB
 * share; DV3D_ENOSPC from dv3d_hot_path means the scene's bounding box needs more) */
size_t dv3d_hot_path_workspace_bytes(const dv3d_net_params_t* net, int n_imgs, int n_ref, int D, int h, int w);
/*   feats_nhwc [n_imgs,Hf,Wf,32]; rotmats/tvecs/K as dv3d_camera_tables; ref_img/edge CSR as
 *   dv3d_planesweep_var; depth_batch [n_ref] int64 (scene id of every reference view);
 *   offsets_host [n_outer*n_inner] doubles (HOST memory; eval-3dvnet.py:23), n_side = 3;
 *   workspace: 256-byte aligned device arena; depth_out [n_ref,h,w];
 *   depth_init_out optional [n_ref,h,w] (the soft-argmin depth before refinement) */
int dv3d_hot_path(const dv3d_net_params_t* net, const float* feats_nhwc, int n_imgs, int Hf, int Wf,
                  const float* rotmats, const float* tvecs, const float* K, const int* ref_img,
                  const int* edge_rowptr, const int* edge_src, int n_ref, const long long* depth_batch,
                  double depth_start, double depth_interval, int D, int h, int w, int H, int W, double edge_len,
                  const double* offsets_host, int n_outer, int n_inner, void* workspace, size_t workspace_bytes,
                  float* depth_init_out, float* depth_out, void* stream);

/* ---- one scene spanning GPUs: symmetric regions (csrc/symm.cu; host side 3dvnet_b200/parallel.py) --------------
 * The reference is single-GPU (mv3d/config.py:3-5); this is the SURVEY.md section 8e "shard UNet by voxel-id range"
 * path.  Every rank owns one buffer of the same size (dv3d_symm_alloc) and maps its peers' copies (dv3d_symm_open on
 * the handles the host side exchanged; legacy CUDA IPC, one process per GPU).
 * Once registered, every sparse-convolution / concat-linear launch whose `out` lies inside the local region also
 * stores its rows at the same offset of every peer copy (NVLink stores from the epilogue), so a row-sharded layer
 * needs no separate all-gather.  dv3d_symm_barrier enqueues the cross-GPU barrier between layers: flags are
 * n_ranks 32-bit words inside the region (zero at start), peers in ascending rank order with this rank left out,
 * epoch strictly increasing.  A peer that does not arrive within ~20 s sets *err_flag and traps the kernel (every later
 * CUDA call of the process then fails) instead of hanging. */
int dv3d_symm_alloc(size_t bytes, void** ptr, void* handle64);   /* zeroed cudaMalloc block + its 64-byte IPC handle */
int dv3d_symm_open(const void* handle64, void** ptr);            /* a PEER process's block, mapped for the current device */
int dv3d_symm_close(void* ptr);
int dv3d_symm_free(void* ptr);
int dv3d_symm_register(const void* base, size_t bytes, void* const* peer_bases, int n_peers);
int dv3d_symm_unregister(const void* base);
int dv3d_symm_barrier(void* local_flags, void* const* peer_flags, int n_peers, int rank, int epoch, int* err_flag,
                      void* stream);
/* as dv3d_symm_barrier, but this rank only waits for the ranks whose bit is set in *wait_mask_dev (a device int,
 * read when the barrier runs; NULL = all).  Every peer is signalled either way. */
int dv3d_symm_barrier_masked(void* local_flags, void* const* peer_flags, int n_peers, int rank, int epoch, int* err_flag,
                             const int* wait_mask_dev, void* stream);

/* BASELINE config C4 from one native call per rank: the hot path of a scene whose reference views are sharded over
 * `world` GPUs (one process each).  This rank computes the cost volumes, depths and PointFlow passes of the
 * contiguous range [ref_start, ref_start + n_ref_local) of the scene's n_ref_total sorted reference views
 * (eval-3dvnet.py:60-99 for those views); ref_img / edge_rowptr / edge_src describe that range only, depth_batch_all
 * [n_ref_total] all views.  Per refinement iteration the rank's point rows are copied into every peer's heap and one
 * flag barrier follows (the all-gather utils.voxelize needs, mv3d/utils.py:39-48), every rank voxelises the whole
 * cloud (identical tables everywhere), and the sparse U-Net runs for this rank's rows of every level with its
 * epilogues storing into all heaps and one barrier per layer (scenemodeling.py:191-237).
 *   heap / peer_heaps: dv3d_symm_alloc + dv3d_symm_open blocks of heap_bytes >= dv3d_hot_path_sharded_heap_bytes,
 *   registered with dv3d_symm_register; peers in ascending rank order with this rank left out; the first 256 bytes
 *   are the barrier flags.  *epoch: last barrier epoch used on this heap (0 at first), updated on return; every
 *   rank must make the same sequence of calls.  err_flag: device int, set if a peer misses a barrier (the kernel
 *   traps).  workspace: dv3d_hot_path_workspace_bytes(net, n_imgs, n_ref_total, D, h, w).
 *   n_ref_local may be 0 (more ranks than views): the rank still takes part in the scene model.
 * Results are those of dv3d_hot_path on the whole scene, rows [ref_start, ref_start + n_ref_local). */
size_t dv3d_hot_path_sharded_heap_bytes(const dv3d_net_params_t* net, int n_ref_total, int h, int w);
int dv3d_hot_path_sharded(const dv3d_net_params_t* net, const float* feats_nhwc, int n_imgs, int Hf, int Wf,
                          const float* rotmats, const float* tvecs, const float* K, const int* ref_img,
                          const int* edge_rowptr, const int* edge_src, int n_ref_local, int ref_start, int n_ref_total,
                          const long long* depth_batch_all, double depth_start, double depth_interval, int D, int h,
                          int w, int H, int W, double edge_len, const double* offsets_host, int n_outer, int n_inner,
                          void* workspace, size_t workspace_bytes, int rank, int world, void* heap, size_t heap_bytes,
                          void* const* peer_heaps, int* epoch, int* err_flag, float* depth_init_out, float* depth_out,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DV3D_H */
