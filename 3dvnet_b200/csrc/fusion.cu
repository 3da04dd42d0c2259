// Depth-map fusion: the consumer of the path's output.
// pointcloudfusion_custom.py:10-116: every pixel of a reference depth map is back-projected,
// re-projected into every other image of the scene, compared with the depth sampled there
// (nearest neighbour, zero padding), counted as consistent when |z - z_sample| < z_thresh and
// the projection is inside the image and in front of the camera; the consistent samples are
// back-projected again and averaged with the point itself.  The reference materialises
// [n_src, 3, h*w] tensors per reference image and loops over images in Python ("very slow",
// README.md:94); here one thread owns a reference pixel and walks the other images with the
// cameras in shared memory - no intermediate tensor, one launch for the whole scene.
#include <math.h>

#include "common.cuh"

namespace dv3d {

constexpr int FC_STRIDE = 48;  // per image: K (9) | Kinv (9) | P rows 0-2 (12) | Pinv rows 0-2 (12) | pad
constexpr int FC_K = 0, FC_KINV = 9, FC_P = 18, FC_PINV = 30;

// inverse of a 3x3 (adjugate) and of a 4x4 (Gauss-Jordan with partial pivoting), both in fp64
__device__ void inv3_d(const double* m, double* o) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double id = 1.0 / (a * A + b * B + c * C);
    o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
    o[3] = B * id; o[4] = (a * i - c * g) * id;  o[5] = -(a * f - c * d) * id;
    o[6] = C * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}
__device__ void inv4_d(const double* m, double* o) {
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            a[r][c] = m[4 * r + c];
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        for (int c = 0; c < 8; ++c) {
            const double t = a[col][c];
            a[col][c] = a[piv][c];
            a[piv][c] = t;
        }
        const double ip = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] *= ip;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = a[r][col];
            for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) o[4 * r + c] = a[r][4 + c];
}

__global__ void fusion_cameras_kernel(const float* __restrict__ poses, const float* __restrict__ K, int n, float* __restrict__ out) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double k[9], ki[9], p[16], pi[16];
    for (int a = 0; a < 9; ++a) k[a] = (double)K[9 * i + a];
    for (int a = 0; a < 16; ++a) p[a] = (double)poses[16 * i + a];
    inv3_d(k, ki);
    inv4_d(p, pi);
    float* o = out + (size_t)i * FC_STRIDE;
    for (int a = 0; a < 9; ++a) o[FC_K + a] = (float)k[a], o[FC_KINV + a] = (float)ki[a];
    for (int a = 0; a < 12; ++a) o[FC_P + a] = (float)p[a], o[FC_PINV + a] = (float)pi[a];
    for (int a = 42; a < FC_STRIDE; ++a) o[a] = 0.f;
}

__device__ __forceinline__ void mat3(const float* m, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = fmaf(m[2], z, fmaf(m[1], y, m[0] * x));
    oy = fmaf(m[5], z, fmaf(m[4], y, m[3] * x));
    oz = fmaf(m[8], z, fmaf(m[7], y, m[6] * x));
}

constexpr int FUSE_CHUNK = 64;  // cameras staged in shared memory per pass

__global__ void __launch_bounds__(256)
depth_fusion_kernel(const float* __restrict__ depths, const float* __restrict__ cams, int n_imgs, int h, int w,
                    float z_thresh, int n_consistent, float* __restrict__ pts_avg, int* __restrict__ n_valid_out,
                    unsigned char* __restrict__ valid_out) {
    pdl_wait();
    __shared__ float s_cam[FUSE_CHUNK][FC_STRIDE];
    const int ref = blockIdx.y;
    const int n_pts = h * w;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < n_pts;
    const float wm1 = (float)(w - 1), hm1 = (float)(h - 1);
    // world point of the reference pixel (pointcloudfusion_custom.py:26-32)
    float X = 0.f, Y = 0.f, Z = 0.f;
    if (live) {
        const float* c = cams + (size_t)ref * FC_STRIDE;
        const float d = __ldg(depths + (size_t)ref * n_pts + p);
        float cx, cy, cz;
        mat3(c + FC_KINV, (float)(p % w) * d, (float)(p / w) * d, d, cx, cy, cz);
        const float* pi = c + FC_PINV;
        X = fmaf(pi[2], cz, fmaf(pi[1], cy, pi[0] * cx)) + pi[3];
        Y = fmaf(pi[6], cz, fmaf(pi[5], cy, pi[4] * cx)) + pi[7];
        Z = fmaf(pi[10], cz, fmaf(pi[9], cy, pi[8] * cx)) + pi[11];
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
    int nv = 0;
    for (int j0 = 0; j0 < n_imgs; j0 += FUSE_CHUNK) {
        const int nj = min(FUSE_CHUNK, n_imgs - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < nj * FC_STRIDE; i += blockDim.x) s_cam[i / FC_STRIDE][i % FC_STRIDE] = __ldg(cams + (size_t)j0 * FC_STRIDE + i);
        __syncthreads();
        if (!live) continue;
        for (int jj = 0; jj < nj; ++jj) {
            const int j = j0 + jj;
            if (j == ref) continue;  // the sources of a reference are all the other images (:104)
            const float* c = s_cam[jj];
            const float* P = c + FC_P;
            // re-projection (:46-50)
            const float px = fmaf(P[2], Z, fmaf(P[1], Y, P[0] * X)) + P[3];
            const float py = fmaf(P[6], Z, fmaf(P[5], Y, P[4] * X)) + P[7];
            const float pz = fmaf(P[10], Z, fmaf(P[9], Y, P[8] * X)) + P[11];
            float qx, qy, z;
            mat3(c + FC_K, px, py, pz, qx, qy, z);
            qx = qx / z;
            qy = qy / z;
            const bool ok_geom = z > 1e-4f && qx >= 0.f && qx <= wm1 && qy >= 0.f && qy <= hm1;  // (:52-54)
            // grid_sample(mode='nearest', align_corners=True, zeros) on the normalised grid (:56-60)
            const float gx = (qx / wm1) * 2.f - 1.f, gy = (qy / hm1) * 2.f - 1.f;
            const float ix = ((gx + 1.f) * 0.5f) * wm1, iy = ((gy + 1.f) * 0.5f) * hm1;
            const float rx = nearbyintf(ix), ry = nearbyintf(iy);
            float zs = 0.f;
            if (rx >= 0.f && rx <= wm1 && ry >= 0.f && ry <= hm1)
                zs = __ldg(depths + (size_t)j * n_pts + (size_t)((int)ry) * w + (int)rx);
            const bool ok = ok_geom && fabsf(z - zs) < z_thresh;  // (:63-66)
            nv += ok;
            // back-projection of the sampled depth (:70-73)
            float sx, sy, sz;
            mat3(c + FC_KINV, qx * zs, qy * zs, zs, sx, sy, sz);
            sx -= P[3]; sy -= P[7]; sz -= P[11];
            const float wx = fmaf(P[8], sz, fmaf(P[4], sy, P[0] * sx));   // R^T
            const float wy = fmaf(P[9], sz, fmaf(P[5], sy, P[1] * sx));
            const float wz = fmaf(P[10], sz, fmaf(P[6], sy, P[2] * sx));
            if (ok && !(isnan(wx) || isnan(wy) || isnan(wz))) {  // (:84-88)
                ax += wx; ay += wy; az += wz;
            }
        }
    }
    if (!live) return;
    const float inv = 1.f / (float)(nv + 1);  // (:89)
    float* o = pts_avg + ((size_t)ref * n_pts + p) * 3;
    o[0] = (X + ax) * inv;
    o[1] = (Y + ay) * inv;
    o[2] = (Z + az) * inv;
    n_valid_out[(size_t)ref * n_pts + p] = nv;
    valid_out[(size_t)ref * n_pts + p] = nv >= n_consistent;
}

}  // namespace dv3d

using namespace dv3d;

extern "C" size_t dv3d_depth_fusion_workspace_bytes(int n_imgs) { return n_imgs <= 0 ? 0 : (size_t)n_imgs * FC_STRIDE * sizeof(float); }

extern "C" int dv3d_depth_fusion(const float* depths, const float* poses, const float* K, int n_imgs, int n_ref, int h, int w,
                                 float z_thresh, int n_consistent_thresh, void* workspace, size_t workspace_bytes,
                                 float* pts_avg, int* n_valid, unsigned char* valid, void* stream) {
    DV3D_REQUIRE(depths && poses && K && workspace && pts_avg && n_valid && valid, "depth_fusion: null pointer");
    DV3D_REQUIRE(n_imgs > 0 && n_ref >= 0 && n_ref <= n_imgs && n_ref <= 65535 && h > 1 && w > 1, "depth_fusion: bad shape");
    DV3D_REQUIRE(workspace_bytes >= dv3d_depth_fusion_workspace_bytes(n_imgs), "depth_fusion: workspace too small");
    if (n_ref == 0) return DV3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* cams = (float*)workspace;
    DV3D_LAUNCH((fusion_cameras_kernel), cdiv(n_imgs, 64), 64, 0, st, poses, K, n_imgs, cams);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((depth_fusion_kernel), dim3(cdiv((long long)h * w, 256), n_ref), 256, 0, st, depths, cams, n_imgs, h, w, z_thresh,
                n_consistent_thresh, pts_avg, n_valid, valid);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
