// tiles.cu — tile validity and uint16 quantisation: the arithmetic of
// tiles_helpers/utils.py on band-sequential (bands, H, W) float32 tiles, the layout rasterio hands the reference.
//   is_black_mask      :201-220   pixel is black if ALL bands ~ nodata, or ALL ~ masked_val (-0.01), or ALL |v| < 1e-6
//   save_tile_pair     :362-373   valid = isfinite(v) & (v != nodata); u16 = clip(int32(rint(v * scale)), 0, nodata_u16 - 1),
//                                 nodata_u16 where not valid
// Both are single-pass, HBM-bound elementwise kernels; comparisons follow numpy's float32 semantics (np.isclose
// on a float32 array with Python-float targets compares |x - f32(y)| <= f32(atol + rtol * |y|) in float32).
#include "hsr_common.cuh"

namespace hsr {

namespace {

struct BlackParams {
    const float* arr;  // [G][B][n]: group g, band b at arr + g * g_stride + b * b_stride
    long long g_stride, b_stride, n;
    int B, has_nodata;
    float nodata, nodata_tol, masked, masked_tol, zero_tol;
    uint8_t* out;                   // [G][n]
    unsigned long long* count;      // nullable [G]: black pixels per group (accumulated)
};

// grid = (blocks, G); a thread owns 4 consecutive pixels (16-byte loads when aligned) and walks the bands.
__global__ void __launch_bounds__(256) black_mask_kernel(const BlackParams P, int vec) {
    const long long g = blockIdx.y;
    const float* ag = P.arr + g * P.g_stride;
    uint8_t* og = P.out + g * P.n;
    const long long n4 = (P.n + 3) >> 2;
    unsigned int cnt = 0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        const long long i = q << 2;
        const bool full = i + 4 <= P.n;
        unsigned int nod = P.has_nodata ? 0xfu : 0u, msk = 0xfu, zer = 0xfu;
        for (int b = 0; b < P.B && (nod | msk | zer); ++b) {
            const float* p = ag + (long long)b * P.b_stride + i;
            float v[4];
            if (full && vec) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (i + j < P.n) ? __ldg(p + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned int bit = 1u << j;
                if (!close32(v[j], P.nodata, P.nodata_tol)) nod &= ~bit;
                if (!close32(v[j], P.masked, P.masked_tol)) msk &= ~bit;
                if (!(fabsf(v[j]) < P.zero_tol)) zer &= ~bit;
            }
        }
        const unsigned int m = nod | msk | zer;
        if (full && vec) {
            *reinterpret_cast<uchar4*>(og + i) = make_uchar4(m & 1u, (m >> 1) & 1u, (m >> 2) & 1u, (m >> 3) & 1u);
            cnt += __popc(m);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i + j < P.n) {
                    og[i + j] = (m >> j) & 1u;
                    cnt += (m >> j) & 1u;
                }
        }
    }
    if (P.count) {
        cnt = (unsigned int)warp_sum((int)cnt);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(P.count + g, (unsigned long long)cnt);
    }
}

__device__ __forceinline__ unsigned short quant1(float v, int has_nodata, float nodata, float scale, int hi, unsigned short nd) {
    return quant_u16(v, has_nodata, nodata, scale, (float)hi, nd);
}

__global__ void __launch_bounds__(256) quantize_u16_kernel(const float* __restrict__ x, long long n, int has_nodata,
                                                           float nodata, float scale, int hi, unsigned short nd,
                                                           unsigned short* __restrict__ out, int vec) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long n8 = vec ? (n >> 3) : 0;
    for (long long i = tid; i < n8; i += nthreads) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
        uint4 o;
        o.x = quant1(a.x, has_nodata, nodata, scale, hi, nd) | ((unsigned int)quant1(a.y, has_nodata, nodata, scale, hi, nd) << 16);
        o.y = quant1(a.z, has_nodata, nodata, scale, hi, nd) | ((unsigned int)quant1(a.w, has_nodata, nodata, scale, hi, nd) << 16);
        o.z = quant1(b.x, has_nodata, nodata, scale, hi, nd) | ((unsigned int)quant1(b.y, has_nodata, nodata, scale, hi, nd) << 16);
        o.w = quant1(b.z, has_nodata, nodata, scale, hi, nd) | ((unsigned int)quant1(b.w, has_nodata, nodata, scale, hi, nd) << 16);
        __stcs(reinterpret_cast<uint4*>(out) + i, o);
    }
    for (long long i = (n8 << 3) + tid; i < n; i += nthreads) out[i] = quant1(__ldg(x + i), has_nodata, nodata, scale, hi, nd);
}

}  // namespace

int black_mask_impl(const float* arr, long long g_stride, long long b_stride, long long n, int B, int G, int has_nodata,
                    float nodata, float nodata_tol, float masked, float masked_tol, float zero_tol, uint8_t* out,
                    unsigned long long* count, cudaStream_t stream) {
    HSR_REQUIRE(arr && out, HSR_EINVAL, "null arr / out pointer");
    HSR_REQUIRE(n >= 0 && B >= 1 && G >= 1 && G <= 65535, HSR_ERANGE, "bad n = %lld, B = %d or G = %d", n, B, G);
    HSR_REQUIRE((reinterpret_cast<uintptr_t>(arr) & 3) == 0, HSR_EALIGN, "arr not 4-byte aligned");
    if (n == 0) return HSR_OK;
    BlackParams P{};
    P.arr = arr, P.g_stride = g_stride, P.b_stride = b_stride, P.n = n, P.B = B, P.has_nodata = has_nodata;
    P.nodata = nodata, P.nodata_tol = nodata_tol, P.masked = masked, P.masked_tol = masked_tol, P.zero_tol = zero_tol;
    P.out = out, P.count = count;
    const long long s_or = (B > 1 ? b_stride : 0) | (G > 1 ? g_stride : 0);
    const int vec = ((reinterpret_cast<uintptr_t>(arr) | (uintptr_t)(s_or * 4)) & 15) == 0 &&
                    ((reinterpret_cast<uintptr_t>(out) | (uintptr_t)(G > 1 ? n : 0)) & 3) == 0;
    long long per = ((n + 3) / 4 + 255) / 256;
    long long cap = (long long)device_sm_count() * 8 / G;
    if (cap < 1) cap = 1;
    dim3 grid((unsigned int)(per < cap ? per : cap), (unsigned int)G);
    black_mask_kernel<<<grid, 256, 0, stream>>>(P, vec);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int quantize_u16_impl(const float* x, long long n, int has_nodata, float nodata, float scale, int nodata_u16,
                      uint16_t* out, cudaStream_t stream) {
    HSR_REQUIRE(x && out, HSR_EINVAL, "null x / out pointer");
    HSR_REQUIRE(n >= 0, HSR_EINVAL, "negative n");
    HSR_REQUIRE(nodata_u16 >= 1 && nodata_u16 <= 65535, HSR_ERANGE, "nodata_u16 = %d outside [1, 65535]", nodata_u16);
    HSR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 1) == 0, HSR_EALIGN,
                "x / out misaligned");
    if (n == 0) return HSR_OK;
    const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    long long blocks = (n / 8 + 255) / 256 + 1;
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    quantize_u16_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(x, n, has_nodata, nodata, scale, nodata_u16 - 1,
                                                                 (unsigned short)nodata_u16, out, vec);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

}  // namespace hsr

namespace hsr {

namespace {

// out[ty * ntx + tx] = sum of the mask bytes of the non-overlapping tile (ty, tx): one block per tile.
__global__ void __launch_bounds__(256) tile_sums_kernel(const uint8_t* __restrict__ mask, long long W, int tile_h, int tile_w,
                                                        int ntx, unsigned int* __restrict__ out) {
    const int t = blockIdx.x;
    const int ty = t / ntx, tx = t - ty * ntx;
    const uint8_t* base = mask + (long long)ty * tile_h * W + (long long)tx * tile_w;
    unsigned int c = 0;
    for (int e = threadIdx.x; e < tile_h * tile_w; e += 256) {
        const int y = e / tile_w, x = e - y * tile_w;
        c += base[(long long)y * W + x] ? 1u : 0u;
    }
    __shared__ unsigned int red[8];
    c = (unsigned int)warp_sum((int)c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        out[t] = s;
    }
}

}  // namespace

int tile_sums_impl(const uint8_t* mask, long long H, long long W, int tile_h, int tile_w, int nty, int ntx,
                   unsigned int* out, cudaStream_t stream) {
    HSR_REQUIRE(mask && out, HSR_EINVAL, "null mask / out pointer");
    HSR_REQUIRE(tile_h >= 1 && tile_w >= 1 && nty >= 0 && ntx >= 0, HSR_EINVAL, "bad tile geometry");
    HSR_REQUIRE((long long)nty * tile_h <= H && (long long)ntx * tile_w <= W, HSR_EINVAL,
                "%d x %d tiles of %d x %d do not fit a %lld x %lld raster", nty, ntx, tile_h, tile_w, H, W);
    if (nty == 0 || ntx == 0) return HSR_OK;
    tile_sums_kernel<<<(unsigned int)(nty * ntx), 256, 0, stream>>>(mask, W, tile_h, tile_w, ntx, out);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

}  // namespace hsr
