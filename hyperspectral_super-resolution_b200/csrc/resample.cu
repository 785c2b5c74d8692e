// resample.cu — grid resampling, the ALIGNED INTEGER-RATIO case of SURVEY section 8f row 4: bringing the Sentinel-2
// 10 m stack onto the EMIT 60 m grid with "average" resampling (the notebook's downsample_s2_to_grid,
// Pairs_EMIT_S2_demo-2.ipynb cell 73, called at s2_emit/poly_regression.py:110-116 with src_scale = 1/255).
// nc_to_envi snaps the EMIT grid to the Sentinel-2 origin with an integer pixel ratio (emit_proj.py:794-797), so
// every 60 m pixel covers exactly factor x factor 10 m pixels and GDAL's average is their plain mean (nodata
// pixels excluded; no valid pixel -> the destination keeps its initial 0).  Parity with GDAL itself is UNPINNED
// (no GDAL / rasterio in the build image); general warps (other CRS, rotation, cubic, bilinear) are out of scope.
// One thread per output pixel; a warp covers 32 consecutive outputs, i.e. a contiguous span of every source row.
#include "hsr_common.cuh"

namespace hsr {

namespace {

template <typename T>
__global__ void __launch_bounds__(256) block_average_kernel(const T* __restrict__ src, long long src_plane_stride,
                                                            long long Ws, int factor, long long Hd, long long Wd,
                                                            int has_nodata, double nodata, float scale, int has_scale,
                                                            float* __restrict__ dst, long long dst_plane_stride) {
    const long long c = blockIdx.y;
    const T* sp = src + c * src_plane_stride;
    float* dp = dst + c * dst_plane_stride;
    const long long n = Hd * Wd;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
        const long long y = o / Wd, x = o - y * Wd;
        const T* blk = sp + (y * factor) * Ws + x * factor;
        double sum = 0.0;
        int cnt = 0;
        for (int r = 0; r < factor; ++r) {
            const T* row = blk + (long long)r * Ws;
            for (int q = 0; q < factor; ++q) {
                const double v = (double)row[q];
                const bool use = !(has_nodata && v == nodata) && v == v;
                sum += use ? v : 0.0;
                cnt += use ? 1 : 0;
            }
        }
        float out = cnt ? (float)(sum / (double)cnt) : 0.f;
        if (has_scale) out = __fmul_rn(out, scale);   // `out *= float(src_scale)` on the float32 result
        dp[o] = out;
    }
}

}  // namespace

int block_average_impl(const void* src, int src_dtype, int C, long long Hs, long long Ws, long long src_plane_stride,
                       int factor, int has_nodata, double nodata, int has_scale, float scale, float* dst,
                       long long dst_plane_stride, cudaStream_t stream) {
    HSR_REQUIRE(src && dst, HSR_EINVAL, "null src / dst pointer");
    HSR_REQUIRE(C >= 1 && C <= 65535 && Hs >= 0 && Ws >= 0 && factor >= 1, HSR_EINVAL, "bad shape or factor");
    HSR_REQUIRE(src_dtype >= 0 && src_dtype <= 2, HSR_EINVAL, "src_dtype must be 0 (u8), 1 (u16) or 2 (f32)");
    const long long Hd = Hs / factor, Wd = Ws / factor;
    HSR_REQUIRE(src_plane_stride >= Hs * Ws && dst_plane_stride >= Hd * Wd, HSR_EINVAL, "plane stride too small");
    if (Hd == 0 || Wd == 0) return HSR_OK;
    long long blocks = (Hd * Wd + 255) / 256;
    long long cap = (long long)device_sm_count() * 8 / C;
    if (cap < 1) cap = 1;
    dim3 grid((unsigned int)(blocks < cap ? blocks : cap), (unsigned int)C);
    if (src_dtype == 0)
        block_average_kernel<uint8_t><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(src), src_plane_stride, Ws, factor,
                                                               Hd, Wd, has_nodata, nodata, scale, has_scale, dst,
                                                               dst_plane_stride);
    else if (src_dtype == 1)
        block_average_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(src), src_plane_stride, Ws,
                                                                factor, Hd, Wd, has_nodata, nodata, scale, has_scale, dst,
                                                                dst_plane_stride);
    else
        block_average_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), src_plane_stride, Ws, factor, Hd,
                                                             Wd, has_nodata, nodata, scale, has_scale, dst, dst_plane_stride);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

}  // namespace hsr

namespace hsr {

namespace {

// Bilinear resampling onto a `factor`-times FINER aligned grid (the notebook's reproject_stack_to_grid with
// Resampling.bilinear, cell 73; s2_emit/poly_regression.py:150-156): the destination pixel centre
// ((x + 0.5) / factor - 0.5 in source pixel coordinates) is interpolated from its 2 x 2 source neighbours; neighbours
// outside the source, equal to nodata or NaN are skipped and the weights renormalised (GDAL's 4-sample bilinear
// kernel), none usable -> 0.  fp64 weights, fp32 result.
__global__ void __launch_bounds__(256) bilinear_up_kernel(const float* __restrict__ src, long long src_plane_stride,
                                                          long long Hs, long long Ws, int factor, int has_nodata,
                                                          float nodata, float* __restrict__ dst,
                                                          long long dst_plane_stride) {
    const long long c = blockIdx.y;
    const float* sp = src + c * src_plane_stride;
    float* dp = dst + c * dst_plane_stride;
    const long long Hd = Hs * factor, Wd = Ws * factor, n = Hd * Wd;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
        const long long y = o / Wd, x = o - y * Wd;
        const double sy = ((double)y + 0.5) / factor - 0.5, sx = ((double)x + 0.5) / factor - 0.5;
        const double fy = floor(sy), fx = floor(sx);
        const long long y0 = (long long)fy, x0 = (long long)fx;
        const double wy1 = sy - fy, wx1 = sx - fx;
        double acc = 0.0, wsum = 0.0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const long long yy = y0 + dy;
            const double wy = dy ? wy1 : 1.0 - wy1;
            if (yy < 0 || yy >= Hs) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const long long xx = x0 + dx;
                const double w = wy * (dx ? wx1 : 1.0 - wx1);
                if (xx < 0 || xx >= Ws || w == 0.0) continue;
                const float v = __ldg(sp + yy * Ws + xx);
                if ((has_nodata && v == nodata) || v != v) continue;
                acc += w * (double)v;
                wsum += w;
            }
        }
        dp[o] = wsum > 0.0 ? (float)(acc / wsum) : 0.f;
    }
}

}  // namespace

int bilinear_upsample_impl(const float* src, int C, long long Hs, long long Ws, long long src_plane_stride, int factor,
                           int has_nodata, float nodata, float* dst, long long dst_plane_stride, cudaStream_t stream) {
    HSR_REQUIRE(src && dst, HSR_EINVAL, "null src / dst pointer");
    HSR_REQUIRE(C >= 1 && C <= 65535 && Hs >= 0 && Ws >= 0 && factor >= 1, HSR_EINVAL, "bad shape or factor");
    HSR_REQUIRE(src_plane_stride >= Hs * Ws && dst_plane_stride >= Hs * Ws * factor * factor, HSR_EINVAL,
                "plane stride too small");
    if (Hs == 0 || Ws == 0) return HSR_OK;
    long long blocks = (Hs * Ws * factor * factor + 255) / 256;
    long long cap = (long long)device_sm_count() * 8 / C;
    if (cap < 1) cap = 1;
    dim3 grid((unsigned int)(blocks < cap ? blocks : cap), (unsigned int)C);
    bilinear_up_kernel<<<grid, 256, 0, stream>>>(src, src_plane_stride, Hs, Ws, factor, has_nodata, nodata, dst,
                                                 dst_plane_stride);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

}  // namespace hsr
