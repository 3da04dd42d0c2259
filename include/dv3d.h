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
 * Per-edge composed camera transform.  For edge e = (ref r, src s):
 *     q = P_s [X;1],  X = R_r^T (K_r^-1 [u z, v z, z] - t_r)      (utils.py:102-106, mvsnet.py:196-199)
 *       = z * (M_e [u,v,1]) + b_e,   M_e = K_s R_s R_r^T K_r^-1,  b_e = K_s (t_s - R_s R_r^T t_r)
 * composed in fp64 on the device and rounded once to fp32.
 *   rotmats [n_imgs,3,3], tvecs [n_imgs,3], K [n_imgs,3,3] (world->camera, full-res K)
 *   edge_ref [E], edge_src [E]  int32 image indices
 *   xform_out [E,12] = M (row-major 9) then b (3)
 * Replaces: torch.inverse + 3 torch.bmm of mvsnet.py:188-199 / lightningmodel.py:138-160.
 */
int dv3d_edge_transforms(const float* rotmats, const float* tvecs, const float* K, const int* edge_ref,
                         const int* edge_src, int n_edges, float* xform_out, void* stream);

/* Per-reference back-projection ray basis: X(u,v,d) = d * (B_r [u,v,1]) + C_r with
 * B_r = R_r^T K_r^-1, C_r = -R_r^T t_r (lightningmodel.py:138-144), fp64-composed.
 *   ref_img [n_ref] int32 image index of every reference;  out [n_ref,12] = B (9) then C (3).
 */
int dv3d_ref_backprojection(const float* rotmats, const float* tvecs, const float* K, const int* ref_img,
                            int n_ref, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused plane-sweep warp + variance  (mvsnet.py:187-216 with utils.py:86-108).
 *   feats_nhwc [n_imgs,Hf,Wf,C]   quarter-resolution features, channels-last, C == 32
 *   xform      [E,12]             from dv3d_edge_transforms, edges sorted by reference
 *   edge_rowptr[n_ref+1], edge_src[E]   CSR of the edges of every reference (int32)
 *   depth hypotheses z_d = linspace(depth_start, depth_start+(D-1)*depth_interval, D)
 *   plane lattice u_j = linspace(0,W-1,w), v_i = linspace(0,H-1,h); grid normalised by the
 *   FULL image size (W-1,H-1) and un-normalised by (Wf-1,Hf-1) as grid_sample does;
 *   bilinear, zero padding per tap, z = |q_z| + 1e-8; divisor = number of edges.
 *   x_var [n_ref,C,D,h,w]         var = E[x^2] - E[x]^2
 * No x_vox[E,C,D,h,w] is ever materialised.
 */
int dv3d_planesweep_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* xform,
                        const int* edge_rowptr, const int* edge_src, int n_ref, double depth_start,
                        double depth_interval, int D, int h, int w, int H, int W, float* x_var, void* stream);

/* ------------------------------------------------------------------------------------
 * Point-level back-projection + re-projection warp + variance, 1 or 2n+1 hypotheses per
 * pixel (lightningmodel.py:132-174 with n_side = 0, and :187-235 with n_side = 3).
 *   depth    [n_ref,h,w];  backproj [n_ref,12] from dv3d_ref_backprojection
 *   hypothesis i (i=-n..n) has depth d + i*offset
 *   pts_out  [n_ref*h*w, n_hyp, 3]   world points
 *   feat_out [n_ref*h*w, n_hyp, feat_stride] variance feature written at channel offset
 *            feat_off (lets the caller write straight into the decoder operand)
 */
int dv3d_points_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* xform,
                    const int* edge_rowptr, const int* edge_src, const float* backproj, const float* depth,
                    int n_ref, int h, int w, int H, int W, int n_side, float offset, float* pts_out,
                    float* feat_out, int feat_stride, int feat_off, void* stream);

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
int dv3d_deconv3d_bn_relu(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                          const float* scale, const float* shift, int Cout, const float* skip, float* y,
                          void* stream);
/* prob conv (8->1, bias) + softmax(-x) over D + expectation of the plane depths
 * (mvsnet.py:152,162,219-227).  x [n,Cin,D,H,W]; weight [1,Cin,3,3,3]; x_reg_out optional
 * [n,D,H,W]; depth_out [n,H,W].  depth values follow torch.linspace's fp32 formula. */
int dv3d_prob_softargmin(const float* x, int n, int Cin, int D, int H, int W, const float* weight, float bias,
                         float depth_start, float depth_end, float* x_reg_out, float* depth_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Voxelisation (utils.py:38-64 incl. torch_cluster grid_cluster; bit-exact index math).
 * SYNCS once (reads the voxel count back).
 *   pts [N,3] f32, batch [N] int64
 *   out: n_anchors_host; anchor_pts [cap,3] f32; anchor_idx3d [cap,3] int32;
 *        anchor_batch [cap] int64; point_anchor [N] int64 (row 0 of anchor_pts_edges)
 *   workspace: dv3d_voxelize_workspace_bytes(N); returns DV3D_ENOSPC if the bounding box
 *   needs more cells than the workspace bitmap holds or anchors exceed `cap`.
 */
size_t dv3d_voxelize_workspace_bytes(long long n_points);
int dv3d_voxelize(const float* pts, const long long* batch, long long N, float edge_len, void* workspace,
                  size_t workspace_bytes, long long cap, long long* n_anchors_host, float* anchor_pts,
                  int* anchor_idx3d, long long* anchor_batch, long long* point_anchor, void* stream);

/* ------------------------------------------------------------------------------------
 * PointNet (scenemodeling.py:116-144).
 *   y = W x + b with optional ReLU on the INPUT (fc(relu(x))), x rows optionally the
 *   concatenation [x_a | pool[seg]] (scenemodeling.py:130-138).
 *   x_a [N,Ca]; pool [n_seg,Cb] gathered through seg [N] (may be NULL with Cb = 0);
 *   weight [Cout, Ca+Cb] (torch nn.Linear layout), bias [Cout]; y [N,Cout]
 */
int dv3d_linear(const float* x_a, int Ca, const float* pool, const long long* seg, int Cb, long long N,
                const float* weight, const float* bias, int Cout, int relu_input, float* y, void* stream);
/* segment max with empty segments = 0 (torch_scatter 'max'): out [n_seg,C] */
int dv3d_segment_max(const float* x, const long long* seg, long long N, int C, long long n_seg, float* out,
                     void* stream);
/* PointNet input rows [pts - anchor_pts[seg] | pts_feat] (lightningmodel.py:182): out [N,3+C] */
int dv3d_pointnet_input(const float* pts, const float* pts_feat, const float* anchor_pts, const long long* seg,
                        long long N, int C, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Sparse 3D-UNet building blocks (scenemodeling.py:16-44,78-113,147-237; MinkowskiEngine
 * 0.5 semantics per SURVEY.md A.4).
 *
 * A coordinate level is: coords [n,4] int32 (batch,x,y,z) in ascending (batch,z,y,x)
 * order, its tensor stride, and an open-addressed hash table (key -> row).
 */
size_t dv3d_hash_bytes(long long n_rows);
/* build the table of a level from its coordinates */
int dv3d_hash_build(const int* coords, long long n, void* table, size_t table_bytes, void* stream);
/* coarser level of a stride-2 convolution: unique(floor(c/(2 ts)) * 2 ts), sorted.  SYNCS.
 * workspace as for voxelize.  coarse_coords capacity = n rows. */
int dv3d_coarsen(const int* coords, long long n, int new_stride, void* workspace, size_t workspace_bytes,
                 int* coarse_coords, long long* n_coarse_host, void* stream);
/* kernel map: for every output row o and offset k (x fastest, 27 offsets, scaled by `step`)
 * nbr[o*27+k] = row of coords_out[o] + offset_k*step in the input level, or -1. */
int dv3d_kernel_map(const int* coords_out, long long n_out, const void* table_in, size_t table_bytes, int step,
                    int* nbr, void* stream);
/* out[o] = sum_k feat[nbr[o,k]] @ W[k]   (W [27,Cin,Cout]); optional fused per-row
 * GroupNorm (gn_weight/gn_bias [Cout], group size Cout/n_groups, eps 1e-5), optional
 * residual add (before the ReLU) and ReLU — the SparseResidual3d / down / up blocks. */
int dv3d_sparse_conv(const float* feat, long long n_in, int Cin, const int* nbr, long long n_out, const float* W,
                     int Cout, const float* gn_weight, const float* gn_bias, int n_groups, const float* residual,
                     int relu, float* out, void* stream);
/* transposed map: nbrT[j*27+k] = coarse row i with fine_j = coarse_i + offset_k*ts_fine, or -1 */
int dv3d_kernel_map_transpose(const int* coords_fine, long long n_fine, const void* table_coarse,
                              size_t table_bytes, int ts_fine, int* nbr, void* stream);
/* 1x1 "feature adjust" on the concatenation [a | b] (ME.cat + k=1 conv, scenemodeling.py:206)
 * followed by GroupNorm + ReLU: W [Ca+Cb, Cout]. */
int dv3d_concat_linear_gn_relu(const float* a, int Ca, const float* b, int Cb, long long n, const float* W,
                               int Cout, const float* gn_weight, const float* gn_bias, int n_groups, float* out,
                               void* stream);

/* ------------------------------------------------------------------------------------
 * PointFlow hypothesis decoder (refinement.py:28-44, lightningmodel.py:238-241).
 * Trilinear sparse interpolation of one level at the query points, written into the
 * decoder operand at a channel offset (missing voxels contribute 0, no renormalisation).
 *   pts [Nq,3] world points; origin [n_batch,3] = position of index (0,0,0) per batch
 *   (scatter-min of the level's voxel positions, refinement.py:33); res = level voxel size,
 *   stride = level tensor stride; pts_batch [Nq] int64 (already unrolled per hypothesis).
 */
int dv3d_sparse_interp(const float* pts, const long long* pts_batch, long long Nq, const float* origin, float res,
                       int stride, const void* table, size_t table_bytes, const float* feat, int C,
                       float* out, int out_stride, int out_off, void* stream);
/* Conv1d(k=3,pad=1,no bias)+BN(folded scale/shift)+ReLU over the hypothesis axis:
 * x [Np,n_hyp,Cin] -> y [Np,n_hyp,Cout]; weight [Cout,Cin,3] (torch layout). */
int dv3d_conv1d_bn_relu(const float* x, long long Np, int n_hyp, int Cin, const float* weight, const float* scale,
                        const float* shift, int Cout, float* y, void* stream);
/* last Conv1d (Cin->1, bias) + softmax over hypotheses + expected offset
 * sum_i p_i * linspace(-n*offset, n*offset)_i; prob_out optional [Np,n_hyp]; offset_out [Np] */
int dv3d_decoder_head(const float* x, long long Np, int n_hyp, int Cin, const float* weight, float bias,
                      float offset, float* prob_out, float* offset_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DV3D_H */
