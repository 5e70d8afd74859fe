// backproject.cu -- depth map -> the point sets the fusion schedule searches (sm_100a).
//
// The reference back-projects the depth image on the CPU (dpt_2_pcld, datasets/ycb/ycb_dataset.py:
// 165-176), keeps the organised cloud [H,W,3], samples N pixels from it (`choose`, :218-237) and
// slices stride-2/4/8 sub-grids out of it (:253-267).  Only those four point sets are ever searched
// (sr = 1 never is), so this kernel produces them directly from the depth map: the 3.7 MB organised
// cloud per frame is neither shipped to the GPU nor materialised on it; the host sends the depth
// map (1.2 MB) and `choose` (48 KB).
//
// Arithmetic is the reference's, which numpy evaluates in float64 (integer pixel grid minus a
// float64 intrinsic): x = ((col - cx) * d) / fx, y = ((row - cy) * d) / fy, z = d, each multiplied by
// the validity mask (d > 1e-8 ? 1 : 0 -- holes become signed zeros), then rounded once to float32
// where the reference casts (`cld.astype(np.float32)`, NN/knn.pyx:95-96).
#include "common.cuh"

namespace ffb6d {

__device__ __forceinline__ void backproject_px(const float *__restrict__ depth, int W, int row, int col,
                                               double fx, double fy, double cx, double cy, float *o)
{
    const float d = __ldg(depth + (size_t)row * W + col);
    const double msk = (d > 1e-8f) ? 1.0 : 0.0;
    const double dd = (double)d;
    o[0] = __double2float_rn(__dmul_rn(__ddiv_rn(__dmul_rn((double)col - cx, dd), fx), msk));
    o[1] = __double2float_rn(__dmul_rn(__ddiv_rn(__dmul_rn((double)row - cy, dd), fy), msk));
    o[2] = __double2float_rn(__dmul_rn(dd, msk));
}

// flat work list per frame: [0,N) sampled points, then the stride-2, -4, -8 sub-grids
__global__ void __launch_bounds__(256)
backproject_kernel(const float *__restrict__ depth, int H, int W, const double *__restrict__ intr,
                   int intr_per_frame, const int *__restrict__ choose, int N, float *__restrict__ cld,
                   float *__restrict__ p2, float *__restrict__ p4, float *__restrict__ p8)
{
    const int b = blockIdx.y;
    const int n2 = (H / 2) * (W / 2), n4 = (H / 4) * (W / 4), n8 = (H / 8) * (W / 8);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N + n2 + n4 + n8) return;
    const double *k = intr + (intr_per_frame ? (size_t)b * 4 : 0);
    const double fx = k[0], fy = k[1], cx = k[2], cy = k[3];
    const float *dp = depth + (size_t)b * H * W;
    int row, col;
    float *o;
    if (t < N) {
        const int px = __ldg(choose + (size_t)b * N + t);
        row = px / W;
        col = px % W;
        o = cld + ((size_t)b * N + t) * 3;
    } else {
        int u = t - N, s, nw;
        float *base;
        if (u < n2) { s = 2; nw = W / 2; base = p2 + (size_t)b * n2 * 3; }
        else if ((u -= n2) < n4) { s = 4; nw = W / 4; base = p4 + (size_t)b * n4 * 3; }
        else { u -= n4; s = 8; nw = W / 8; base = p8 + (size_t)b * n8 * 3; }
        row = (u / nw) * s;
        col = (u % nw) * s;
        o = base + (size_t)u * 3;
    }
    float v[3];
    backproject_px(dp, W, row, col, fx, fy, cx, cy, v);
    o[0] = v[0];
    o[1] = v[1];
    o[2] = v[2];
}

}  // namespace ffb6d

using namespace ffb6d;

extern "C" int ffb6d_backproject(const float *depth, int64_t B, int64_t H, int64_t W,
                                 const double *intrinsics, int intrinsics_per_frame,
                                 const int *choose, int64_t N, float *cld, float *pyr2, float *pyr4,
                                 float *pyr8, ffb6d_stream_t stream)
{
    FFB6D_CHECK_ARG(B >= 0 && H >= 8 && W >= 8 && N >= 0 && H * W < (1ll << 31) && B < 65536,
                    "backproject: bad size (B=%lld H=%lld W=%lld N=%lld)", (long long)B, (long long)H,
                    (long long)W, (long long)N);
    if (B == 0) return FFB6D_OK;
    FFB6D_CHECK_ARG(depth && intrinsics && cld && pyr2 && pyr4 && pyr8 && (N == 0 || choose),
                    "backproject: null pointer");
    const int64_t work = N + (H / 2) * (W / 2) + (H / 4) * (W / 4) + (H / 8) * (W / 8);
    dim3 grid((unsigned)ceil_div(work, 256), (unsigned)B);
    backproject_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth, (int)H, (int)W, intrinsics,
                                                              intrinsics_per_frame, choose, (int)N, cld,
                                                              pyr2, pyr4, pyr8);
    FFB6D_LAUNCH_OK("backproject_kernel");
    return FFB6D_OK;
}
