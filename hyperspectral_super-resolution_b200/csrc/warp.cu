// warp.cu — general grid warp: the band-interleaved WGS-84 ortho cube resampled onto the
// Sentinel-2 UTM 60 m grid with GDAL's cubic kernel — the job nc_to_envi hands to a `gdalwarp -r cubic -srcnodata
// -9999 -dstnodata -9999 -t_srs <S2 CRS> -te ... -ts ...` subprocess (EMIT_data/emit_proj.py:876-940) — and the
// same-CRS affine case of the notebook's reproject_stack_to_grid.  Parity with GDAL / PROJ is UNPINNED (neither is
// installable in the build image); the algorithm restated is documented in oracle/warp.py:
//   * destination pixel centre -> projected coordinates (destination geotransform) -> lon / lat by the inverse
//     transverse Mercator (Krueger series to n^6, Karney 2011; fp64, one destination pixel per lane) -> source pixel
//     coordinates (inverse source geotransform); exact per pixel (gdalwarp -et 0);
//   * separable cubic-convolution (a = -0.5) or bilinear weights, radius widened to ceil(r / scale) with the argument
//     scaled where the destination is coarser than the source; taps outside the source or equal to nodata are
//     skipped PER BAND and the sum is divided by the accumulated weight; centre outside the source or accumulated
//     weight < 1e-6 -> dst_nodata.  NaN is an ordinary value.
// Default kernel: warp_lane_kernel (lanes across destination pixels, the tap box of a 16 x 8 tile staged transposed in
// shared memory, per-pixel weights in registers for the whole spectrum).  The band-per-lane kernels it replaced remain
// as fall-backs: warp_pipe_kernel (filter radii beyond 4) and warp_tile_kernel (no coordinate workspace).  All of them
// classify what they stage: no nodata -> plain FMAs under a uniform weight sum; all nodata -> skipped; band-specific
// nodata -> exact per-element weights.
#include <math.h>
#include <stdlib.h>

#include "hsr_common.cuh"

namespace hsr {

namespace {

constexpr int WARPS = 8;        // warps per CTA
constexpr int MAX_TAPS = 16;    // taps per axis: radius <= 8, i.e. scale >= 0.25 for cubic
constexpr int GB = 128;         // bands per CTA (32 lanes x 4)
constexpr int TILE_PX = 64;     // destination pixels per tile (<= 8 x 8)
constexpr int BOX_CAP = 192;    // source pixels staged per tile and band group: 192 x 512 B = 96 KB, two CTAs per SM

struct WarpParams {
    const float* src;
    long long Hs, Ws, src_pix_stride;
    int bands;
    double dgt[6];              // destination geotransform
    double sx0, sy0, inv[4];    // source: px = inv0 * (X - sx0) + inv1 * (Y - sy0); py = inv2 * (X - sx0) + inv3 * (Y - sy0)
    int utm;                    // 1: destination is transverse Mercator (UTM), source is lon / lat in degrees
    double lon0_deg, false_northing, k0A_inv, e, e2m, beta[6];
    long long Hd, Wd;
    float* dst;
    long long dst_pix_stride;
    int has_nodata;
    float nodata, dst_nodata;
    int kind;                   // 1 bilinear, 2 cubic
    int rx, ry;                 // radius in taps per axis
    double fx, fy;              // min(scale, 1) per axis
    double* coords;             // hsr_warp_coords_f64 only
    int tile_w, tile_h;         // destination tile of one CTA (tile_w * tile_h <= TILE_PX)
};

__device__ __forceinline__ double taup_of(double tau, double e) {
    const double s1 = sqrt(1.0 + tau * tau);
    const double sigma = sinh(e * atanh(e * tau / s1));
    return tau * sqrt(1.0 + sigma * sigma) - sigma * s1;
}

// centre of destination pixel (col, row) -> source pixel coordinates (pixel (i, j) has its centre at (i + .5, j + .5))
__device__ void dst_to_src(const WarpParams& P, double c, double r, double& px, double& py) {
    double X = P.dgt[0] + c * P.dgt[1] + r * P.dgt[2];
    double Y = P.dgt[3] + c * P.dgt[4] + r * P.dgt[5];
    if (P.utm) {
        const double xi = (Y - P.false_northing) * P.k0A_inv, eta = (X - 500000.0) * P.k0A_inv;
        double xip = xi, etap = eta;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double s, co;
            sincos(2.0 * (j + 1) * xi, &s, &co);
            const double a = 2.0 * (j + 1) * eta;
            xip -= P.beta[j] * s * cosh(a);
            etap -= P.beta[j] * co * sinh(a);
        }
        const double sh = sinh(etap), cx = cos(xip);
        const double tp = sin(xip) / sqrt(sh * sh + cx * cx);
        const double lam = atan2(sh, cx);
        double tau = tp;
        for (int it = 0; it < 4; ++it) {     // Newton on tau'(tau) = tp; quadratic, 2-3 steps reach 1 ulp
            const double ti = taup_of(tau, P.e);
            tau += (tp - ti) / sqrt(1.0 + ti * ti) * (1.0 + P.e2m * tau * tau) / (P.e2m * sqrt(1.0 + tau * tau));
        }
        X = lam * (180.0 / 3.14159265358979323846) + P.lon0_deg;
        Y = atan(tau) * (180.0 / 3.14159265358979323846);
    }
    const double dx = X - P.sx0, dy = Y - P.sy0;
    px = P.inv[0] * dx + P.inv[1] * dy;
    py = P.inv[2] * dx + P.inv[3] * dy;
}

__device__ __forceinline__ double tap_weight(int kind, double x) {
    const double ax = fabs(x);
    double w;
    if (kind == 2) {    // GWKCubic, a = -0.5
        if (ax <= 1.0) w = (1.5 * ax - 2.5) * ax * ax + 1.0;
        else if (ax <= 2.0) w = ((-0.5 * ax + 2.5) * ax - 4.0) * ax + 2.0;
        else w = 0.0;
    } else {
        w = ax <= 1.0 ? 1.0 - ax : 0.0;
    }
    return w;
}

__global__ void __launch_bounds__(32 * WARPS) warp_coords_kernel(const WarpParams P) {
    const long long n = P.Hd * P.Wd;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
        const long long r = o / P.Wd, c = o - r * P.Wd;
        double px, py;
        dst_to_src(P, (double)c + 0.5, (double)r + 0.5, px, py);
        P.coords[2 * o] = px;
        P.coords[2 * o + 1] = py;
    }
}

// One CTA = one TW x TH tile of the destination and one group of GB = 128 bands (lane l of a warp owns bands
// 4l .. 4l+3 of the group).  (1) 64 threads transform the tile's pixel centres (fp64) and the CTA takes the bounding
// box of all their taps in the source; (2) the 8 warps copy the box's spectra (512 B per pixel and group) into shared
// memory ONCE — the only global reads of the tile; 16-byte loads when the records allow, scalar otherwise; (3) each warp
// resamples 8 destination pixels from shared memory (conflict-free 16-byte loads, fp32 FMAs).  A box larger than the
// staging buffer (extreme down-scaling or rotation) falls back to reading the taps from global memory, same arithmetic.
template <bool SRC_VEC>
__device__ __forceinline__ float4 load_group_raw(const float* __restrict__ rec, int b, int bands, bool act) {
    // bands b .. b+3 of one spectrum (words beyond `bands`: record padding with SRC_VEC, 0 otherwise)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!act) return v;
    if (SRC_VEC) {
        v = __ldg(reinterpret_cast<const float4*>(rec + b));
    } else {
        v.x = __ldg(rec + b);
        if (b + 1 < bands) v.y = __ldg(rec + b + 1);
        if (b + 2 < bands) v.z = __ldg(rec + b + 2);
        if (b + 3 < bands) v.w = __ldg(rec + b + 3);
    }
    return v;
}

// words beyond `bands` take the value of band b, so that they never change what the lane sees of the nodata pattern
__device__ __forceinline__ float4 pad_fix(float4 v, int b, int bands) {
    if (b + 1 >= bands) v.y = v.x;
    if (b + 2 >= bands) v.z = v.x;
    if (b + 3 >= bands) v.w = v.x;
    return v;
}

template <bool SRC_VEC>
__device__ __forceinline__ float4 load_group(const float* __restrict__ rec, int b, int bands, bool act) {
    return pad_fix(load_group_raw<SRC_VEC>(rec, b, bands, act), b, bands);
}

template <bool SRC_VEC, bool DST_VEC>
__global__ void __launch_bounds__(32 * WARPS, 2) warp_tile_kernel(const WarpParams P) {
    extern __shared__ __align__(16) float4 box[];          // [BOX_CAP][32]
    __shared__ double s_xy[TILE_PX][2];
    __shared__ int s_box[4];                                // min ix, max ix, min iy, max iy over the tile's pixels
    __shared__ unsigned char s_cls[BOX_CAP];                // per staged source pixel: 0 clean, 1 fill, 2 mixed
    __shared__ double s_w1[WARPS][32];
    __shared__ float s_w[WARPS][MAX_TAPS * MAX_TAPS];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    double* w1 = s_w1[wib];
    float* tw = s_w[wib];
    const int TW = P.tile_w, TH = P.tile_h, npix_tile = TW * TH;
    const long long tiles_x = (P.Wd + TW - 1) / TW, tiles_y = (P.Hd + TH - 1) / TH, ntiles = tiles_x * tiles_y;
    const int b = (int)blockIdx.y * GB + 4 * lane;          // my first band
    const bool act = b < P.bands;
    const int ntx = 2 * P.rx, nty = 2 * P.ry;
    const float nd = P.nodata;
    const bool has_nd = P.has_nodata != 0;
    const unsigned int FULL = 0xffffffffu;
    const float dnd = P.dst_nodata;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long ty = tile / tiles_x, tx = tile - ty * tiles_x;
        // ---- 1. transform the pixel centres, bounding box of the taps
        if (tid < 4) s_box[tid] = (tid & 1) ? -2147483647 : 2147483647;
        __syncthreads();
        if (tid < npix_tile) {
            const long long r = ty * TH + tid / TW, c = tx * TW + tid % TW;
            double px = -1.0, py = -1.0;
            if (r < P.Hd && c < P.Wd) {
                if (P.coords) {     // transformed beforehand by warp_coords_kernel (all SMs busy instead of one warp per CTA)
                    const double2 xy = __ldg(reinterpret_cast<const double2*>(P.coords) + (r * P.Wd + c));
                    px = xy.x;
                    py = xy.y;
                } else {
                    dst_to_src(P, (double)c + 0.5, (double)r + 0.5, px, py);
                }
            }
            const bool inside = px >= 0.0 && px < (double)P.Ws && py >= 0.0 && py < (double)P.Hs;
            s_xy[tid][0] = inside ? px : -1.0;
            s_xy[tid][1] = inside ? py : -1.0;
            if (inside) {
                const int ix = (int)floor(px - 0.5), iy = (int)floor(py - 0.5);
                atomicMin(&s_box[0], ix);
                atomicMax(&s_box[1], ix);
                atomicMin(&s_box[2], iy);
                atomicMax(&s_box[3], iy);
            }
        }
        __syncthreads();
        long long bx0 = (long long)s_box[0] + 1 - P.rx, bx1 = (long long)s_box[1] + P.rx;
        long long by0 = (long long)s_box[2] + 1 - P.ry, by1 = (long long)s_box[3] + P.ry;
        const bool any_inside = s_box[1] >= s_box[0];
        bx0 = bx0 < 0 ? 0 : bx0;
        by0 = by0 < 0 ? 0 : by0;
        bx1 = bx1 >= P.Ws ? P.Ws - 1 : bx1;
        by1 = by1 >= P.Hs ? P.Hs - 1 : by1;
        const int bw = any_inside ? (int)(bx1 - bx0 + 1) : 0, bh = any_inside ? (int)(by1 - by0 + 1) : 0;
        const bool staged = any_inside && (long long)bw * bh <= BOX_CAP;
        // ---- 2. stage the box: one warp per source pixel, 512 B of its spectrum; classify it once for all its uses:
        //         0 = no band of the group is nodata, 1 = every band is (a fill pixel), 2 = some are
        if (staged) {
            constexpr int CB = 8;      // loads in flight per warp
            const int nbox = bw * bh;
            for (int p0 = wib; p0 < nbox; p0 += WARPS * CB) {
                float4 v[CB];
#pragma unroll
                for (int u = 0; u < CB; ++u) {
                    const int p = p0 + u * WARPS;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p < nbox) {
                        const long long yy = by0 + p / bw, xx = bx0 + p % bw;
                        v[u] = load_group_raw<SRC_VEC>(P.src + (yy * P.Ws + xx) * P.src_pix_stride, b, P.bands, act);
                    }
                }
#pragma unroll
                for (int u = 0; u < CB; ++u) {
                    const int p = p0 + u * WARPS;
                    if (p >= nbox) break;
                    const float4 x = pad_fix(v[u], b, P.bands);
                    box[p * 32 + lane] = x;
                    int cls = 0;
                    if (has_nd) {
                        const bool any = act && (x.x == nd || x.y == nd || x.z == nd || x.w == nd);
                        const bool all = !act || (x.x == nd && x.y == nd && x.z == nd && x.w == nd);
                        cls = !__any_sync(FULL, any) ? 0 : (__all_sync(FULL, all) ? 1 : 2);
                    }
                    if (lane == 0) s_cls[p] = (unsigned char)cls;
                }
            }
        }
        __syncthreads();
        // (barrier) + does any staged pixel hold a nodata?  Tiles inside the swath do not: they take the lean loop below
        const bool box_clean = __syncthreads_or(staged && tid < bw * bh && s_cls[tid] != 0) == 0;
        // ---- 3. resample: warp w takes pixels w, w + 8, ... of the tile
        for (int q = wib; q < npix_tile; q += WARPS) {
            const long long r = ty * TH + q / TW, c = tx * TW + q % TW;
            if (r >= P.Hd || c >= P.Wd) continue;
            const double px = s_xy[q][0], py = s_xy[q][1];
            float* outp = P.dst + (r * P.Wd + c) * P.dst_pix_stride + b;
            float4 o = make_float4(dnd, dnd, dnd, dnd);
            if (px >= 0.0) {
                // separable weights: lane t < 16 computes column tap t, lane 16 + t row tap t (fp64), then the products
                const double fxp = floor(px - 0.5), fyp = floor(py - 0.5);
                const long long ix = (long long)fxp, iy = (long long)fyp;
                {
                    const bool isx = lane < 16;
                    const int t = (isx ? lane : lane - 16) + 1 - (isx ? P.rx : P.ry);
                    const double dd = isx ? px - 0.5 - fxp : py - 0.5 - fyp;
                    __syncwarp();                       // everybody is done with the previous pixel's tables
                    w1[lane] = tap_weight(P.kind, ((double)t - dd) * (isx ? P.fx : P.fy));
                    __syncwarp();
                    for (int t2 = lane; t2 < ntx * nty; t2 += 32) {
                        const int j = t2 / ntx, k = t2 - j * ntx;
                        tw[t2] = (float)(w1[16 + j] * w1[k]);
                    }
                    __syncwarp();
                }
                // taps inside the source: rows [jlo, jhi), columns [klo, khi) of the window whose first tap is (x0t, y0t)
                const long long x0t = ix + 1 - P.rx, y0t = iy + 1 - P.ry;
                int jlo = y0t < 0 ? (int)-y0t : 0, klo = x0t < 0 ? (int)-x0t : 0;
                int jhi = y0t + nty > P.Hs ? (int)(P.Hs - y0t) : nty, khi = x0t + ntx > P.Ws ? (int)(P.Ws - x0t) : ntx;
                // a tap of weight zero contributes nothing (not even its NaN): zero rows / columns at the ends of the
                // window (the widened filter's support is rarely full) are trimmed, the others are skipped tap by tap
                const unsigned int nzw = __ballot_sync(FULL, w1[lane] != 0.0);      // bits 0..15 columns, 16..31 rows
                {
                    const unsigned int cm = (nzw & 0xffffu) & (khi >= 32 ? 0xffffffffu : ((1u << khi) - 1u)) & ~((1u << klo) - 1u);
                    const unsigned int rm = (nzw >> 16) & ((1u << jhi) - 1u) & ~((1u << jlo) - 1u);
                    if (cm == 0u || rm == 0u) {
                        jlo = jhi = 0;
                    } else {
                        klo = __ffs((int)cm) - 1;
                        khi = 32 - __clz((int)cm);
                        jlo = __ffs((int)rm) - 1;
                        jhi = 32 - __clz((int)rm);
                    }
                }
                const unsigned int span = ((1u << khi) - 1u) & ~((1u << klo) - 1u);
                const bool dense_cols = ((nzw & 0xffffu) & span) == span && (((nzw >> 16) >> jlo) & ((1u << (jhi - jlo)) - 1u)) == ((1u << (jhi - jlo)) - 1u);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, wu = 0.f;
                if (staged && box_clean && dense_cols) {
                    // lean loop: no nodata anywhere in the box, no zero weight inside the trimmed window
                    const float4* bp = box + ((int)(y0t - by0) * bw + (int)(x0t - bx0)) * 32 + lane;
                    for (int j = jlo; j < jhi; ++j) {
                        const float4* rowb = bp + j * bw * 32;
                        const float* twj = tw + j * ntx;
#pragma unroll 2
                        for (int k = klo; k < khi; ++k) {
                            const float w = twj[k];
                            const float4 v = rowb[k * 32];
                            a0 = fmaf(w, v.x, a0);
                            a1 = fmaf(w, v.y, a1);
                            a2 = fmaf(w, v.z, a2);
                            a3 = fmaf(w, v.w, a3);
                            wu += w;
                        }
                    }
                } else if (staged) {
                    const int base = (int)(y0t - by0) * bw + (int)(x0t - bx0);
                    for (int j = jlo; j < jhi; ++j) {
                        const int rowi = base + j * bw;
                        const float* twj = tw + j * ntx;
#pragma unroll 2
                        for (int k = klo; k < khi; ++k) {
                            const int pi = rowi + k;
                            const int cls = s_cls[pi];                  // three independent shared-memory loads, then branch
                            const float w = twj[k];
                            const float4 v = box[pi * 32 + lane];
                            if (cls == 1 || w == 0.f) continue;         // a fill pixel (skipped by every band) or a zero weight
                            if (cls == 0) {                             // plain FMAs under a warp-uniform weight sum
                                a0 = fmaf(w, v.x, a0);
                                a1 = fmaf(w, v.y, a1);
                                a2 = fmaf(w, v.z, a2);
                                a3 = fmaf(w, v.w, a3);
                                wu += w;
                            } else {                                    // band-specific nodata: exact per element
                                const float w0 = v.x == nd ? 0.f : w, w1q = v.y == nd ? 0.f : w;
                                const float w2 = v.z == nd ? 0.f : w, w3 = v.w == nd ? 0.f : w;
                                a0 = fmaf(w0, v.x, a0);
                                a1 = fmaf(w1q, v.y, a1);
                                a2 = fmaf(w2, v.z, a2);
                                a3 = fmaf(w3, v.w, a3);
                                m0 += w0;
                                m1 += w1q;
                                m2 += w2;
                                m3 += w3;
                            }
                        }
                    }
                } else {    // box too large for the staging buffer: the same taps straight from global memory
                    for (int j = jlo; j < jhi; ++j) {
                        const float* rowp = P.src + ((y0t + j) * P.Ws + x0t) * P.src_pix_stride;
                        for (int k = klo; k < khi; ++k) {
                            const float w = tw[j * ntx + k];
                            if (w == 0.f) continue;
                            const float4 v = load_group<SRC_VEC>(rowp + k * P.src_pix_stride, b, P.bands, act);
                            const bool h = has_nd && act;
                            const float w0 = (h && v.x == nd) ? 0.f : w, w1q = (h && v.y == nd) ? 0.f : w;
                            const float w2 = (h && v.z == nd) ? 0.f : w, w3 = (h && v.w == nd) ? 0.f : w;
                            a0 = fmaf(w0, v.x, a0);
                            a1 = fmaf(w1q, v.y, a1);
                            a2 = fmaf(w2, v.z, a2);
                            a3 = fmaf(w3, v.w, a3);
                            m0 += w0;
                            m1 += w1q;
                            m2 += w2;
                            m3 += w3;
                        }
                    }
                }
                const float s0 = wu + m0, s1 = wu + m1, s2 = wu + m2, s3 = wu + m3;
                o.x = s0 >= 1e-6f ? __fdiv_rn(a0, s0) : dnd;
                o.y = s1 >= 1e-6f ? __fdiv_rn(a1, s1) : dnd;
                o.z = s2 >= 1e-6f ? __fdiv_rn(a2, s2) : dnd;
                o.w = s3 >= 1e-6f ? __fdiv_rn(a3, s3) : dnd;
            }
            if (act) {
                if (DST_VEC) {
                    __stcs(reinterpret_cast<float4*>(outp), o);
                } else {
                    __stcs(outp, o.x);
                    if (b + 1 < P.bands) __stcs(outp + 1, o.y);
                    if (b + 2 < P.bands) __stcs(outp + 2, o.z);
                    if (b + 3 < P.bands) __stcs(outp + 3, o.w);
                }
            }
        }
        __syncthreads();        // the next tile overwrites s_xy / s_box / box
    }
}

// ---------------------------------------------------------------------------- pipelined variant
// The fast path (coordinates transformed beforehand, <= 64 taps): ONE persistent CTA per SM with 16 warps and TWO staging
// buffers.  Work is a sequence of stages (tile, band group); while the warps resample stage s from buffer s & 1, the
// cp.async copies of stage s + 1 land in the other buffer.  What depends on the tile only — bounding box, separable
// weights, trimmed windows of its 32 pixels — is prepared once per tile (kept in shared memory for its band groups).
constexpr int PWARPS = 16;
constexpr int PTILE = 32;       // destination pixels per tile (8 x 4 at most)
constexpr int PTAPS = 64;       // taps per pixel

struct TileState {
    int box[4];                 // scratch: min ix, max ix, min iy, max iy
    int bx0, by0, bw, bh, staged;
    double xy[PTILE][2];
    float tw[PTILE][PTAPS];
    float wsum[PTILE];          // sum of the weights of the window [jlo, jhi) x [0, ntx)
    int meta[PTILE][8];         // inside, jlo, jhi, klo, khi, x0t, y0t, flags: 1 dense (no zero weight inside the trimmed
                                // window), 2 whole rows usable (all ntx columns inside the source)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one row of N taps: acc += w[k] * v[k] for the lane's four bands; weights by 8-byte loads (rows of N = 4, 6, 8 floats
// start on 8-byte boundaries), values by conflict-free 16-byte loads
template <int N>
__device__ __forceinline__ void row_fma(const float4* __restrict__ rowb, const float* __restrict__ twj, float& a0, float& a1,
                                        float& a2, float& a3) {
    float w[N];
#pragma unroll
    for (int k = 0; k < N; k += 2) {
        const float2 t = *reinterpret_cast<const float2*>(twj + k);
        w[k] = t.x;
        w[k + 1] = t.y;
    }
    float4 v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = rowb[k * 32];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        a0 = fmaf(w[k], v[k].x, a0);
        a1 = fmaf(w[k], v[k].y, a1);
        a2 = fmaf(w[k], v[k].z, a2);
        a3 = fmaf(w[k], v[k].w, a3);
    }
}

template <bool SRC_VEC, bool DST_VEC>
__global__ void __launch_bounds__(32 * PWARPS, 1) warp_pipe_kernel(const WarpParams P) {
    extern __shared__ __align__(16) float4 dyn[];            // two boxes of [BOX_CAP][32] float4
    __shared__ TileState s_tile[2];
    __shared__ unsigned char s_cls[2][BOX_CAP];
    __shared__ double s_w1[PWARPS][32];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int TW = P.tile_w, TH = P.tile_h, npix_tile = TW * TH;
    const long long tiles_x = (P.Wd + TW - 1) / TW, tiles_y = (P.Hd + TH - 1) / TH, ntiles = tiles_x * tiles_y;
    const int G = (P.bands + GB - 1) / GB;
    const int ntx = 2 * P.rx, nty = 2 * P.ry;
    const float nd = P.nodata, dnd = P.dst_nodata;
    const bool has_nd = P.has_nodata != 0;
    const unsigned int FULL = 0xffffffffu;
    const int tw_shift = 31 - __clz(TW);
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long nstages = my_tiles * G;

    // ---- per tile: coordinates -> box -> weights and windows of its pixels
    auto prepare_tile = [&](long long it) {
        TileState& T = s_tile[it & 1];
        const long long tile = blockIdx.x + it * gridDim.x;
        const long long ty = tile / tiles_x, tx = tile - ty * tiles_x;
        if (tid < 4) T.box[tid] = (tid & 1) ? -2147483647 : 2147483647;
        __syncthreads();
        if (tid < npix_tile) {
            const long long r = ty * TH + tid / TW, c = tx * TW + tid % TW;
            double px = -1.0, py = -1.0;
            if (r < P.Hd && c < P.Wd) {
                const double2 xy = __ldg(reinterpret_cast<const double2*>(P.coords) + (r * P.Wd + c));
                px = xy.x;
                py = xy.y;
            }
            const bool inside = px >= 0.0 && px < (double)P.Ws && py >= 0.0 && py < (double)P.Hs;
            T.xy[tid][0] = inside ? px : -1.0;
            T.xy[tid][1] = inside ? py : -1.0;
            if (inside) {
                const int ix = (int)floor(px - 0.5), iy = (int)floor(py - 0.5);
                atomicMin(&T.box[0], ix);
                atomicMax(&T.box[1], ix);
                atomicMin(&T.box[2], iy);
                atomicMax(&T.box[3], iy);
            }
        }
        __syncthreads();
        long long bx0 = (long long)T.box[0] + 1 - P.rx, bx1 = (long long)T.box[1] + P.rx;
        long long by0 = (long long)T.box[2] + 1 - P.ry, by1 = (long long)T.box[3] + P.ry;
        const bool any_inside = T.box[1] >= T.box[0];
        bx0 = bx0 < 0 ? 0 : bx0;
        by0 = by0 < 0 ? 0 : by0;
        bx1 = bx1 >= P.Ws ? P.Ws - 1 : bx1;
        by1 = by1 >= P.Hs ? P.Hs - 1 : by1;
        const int bw = any_inside ? (int)(bx1 - bx0 + 1) : 0, bh = any_inside ? (int)(by1 - by0 + 1) : 0;
        if (tid == 0) {
            T.bx0 = (int)bx0;
            T.by0 = (int)by0;
            T.bw = bw;
            T.bh = bh;
            T.staged = any_inside && (long long)bw * bh <= BOX_CAP;
        }
        double* w1 = s_w1[wib];
        for (int q = wib; q < npix_tile; q += PWARPS) {
            const double px = T.xy[q][0], py = T.xy[q][1];
            int* M = T.meta[q];
            if (!(px >= 0.0)) {
                if (lane == 0) M[0] = 0;
                continue;
            }
            const double fxp = floor(px - 0.5), fyp = floor(py - 0.5);
            const long long ix = (long long)fxp, iy = (long long)fyp;
            const bool isx = lane < 16;
            const int t = (isx ? lane : lane - 16) + 1 - (isx ? P.rx : P.ry);
            const double dd = isx ? px - 0.5 - fxp : py - 0.5 - fyp;
            __syncwarp();
            w1[lane] = tap_weight(P.kind, ((double)t - dd) * (isx ? P.fx : P.fy));
            __syncwarp();
            for (int t2 = lane; t2 < ntx * nty; t2 += 32) {
                const int j = t2 / ntx, k = t2 - j * ntx;
                T.tw[q][t2] = (float)(w1[16 + j] * w1[k]);
            }
            const long long x0t = ix + 1 - P.rx, y0t = iy + 1 - P.ry;
            int jlo = y0t < 0 ? (int)-y0t : 0, klo = x0t < 0 ? (int)-x0t : 0;
            int jhi = y0t + nty > P.Hs ? (int)(P.Hs - y0t) : nty, khi = x0t + ntx > P.Ws ? (int)(P.Ws - x0t) : ntx;
            const unsigned int nzw = __ballot_sync(FULL, w1[lane] != 0.0);
            const unsigned int cm = (nzw & 0xffffu) & ((1u << khi) - 1u) & ~((1u << klo) - 1u);
            const unsigned int rm = (nzw >> 16) & ((1u << jhi) - 1u) & ~((1u << jlo) - 1u);
            if (cm == 0u || rm == 0u) {
                jlo = jhi = klo = khi = 0;
            } else {
                klo = __ffs((int)cm) - 1;
                khi = 32 - __clz((int)cm);
                jlo = __ffs((int)rm) - 1;
                jhi = 32 - __clz((int)rm);
            }
            const unsigned int cspan = ((1u << khi) - 1u) & ~((1u << klo) - 1u), rspan = ((1u << jhi) - 1u) & ~((1u << jlo) - 1u);
            // weight sum of the rows [jlo, jhi) over ALL ntx columns (the lean loop below runs whole rows: columns
            // trimmed for a zero weight add exactly 0); fixed order: lanes stride the taps, then a shuffle tree
            __syncwarp();
            float ws = 0.f;
            for (int t2 = jlo * ntx + lane; t2 < jhi * ntx; t2 += 32) ws += T.tw[q][t2];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(FULL, ws, o);
            const bool whole = x0t >= 0 && x0t + ntx <= P.Ws;
            if (lane == 0) {
                M[0] = 1;
                M[1] = jlo;
                M[2] = jhi;
                M[3] = klo;
                M[4] = khi;
                M[5] = (int)x0t;
                M[6] = (int)y0t;
                M[7] = (((cm & cspan) == cspan && (rm & rspan) == rspan) ? 1 : 0) | (whole && rm != 0u ? 2 : 0);
                T.wsum[q] = ws;
            }
        }
        __syncthreads();
    };

    // ---- copy of one stage's box into its buffer (asynchronous; one commit group per stage, possibly empty)
    auto issue_copy = [&](long long s) {
        const TileState& T = s_tile[(s / G) & 1];
        const int b = (int)(s % G) * GB + 4 * lane;
        float4* box = dyn + (size_t)(s & 1) * BOX_CAP * 32;
        if (T.staged && b < P.bands) {
            const int bw = T.bw, nbox = bw * T.bh;
            // running source pointer: no multiplication, division or shared-memory re-read inside the loop (the
            // cp.async statements are memory clobbers, so everything the loop needs lives in registers)
            const long long stride = P.src_pix_stride;
            const float* rec = P.src + ((long long)T.by0 * P.Ws + T.bx0) * stride + b + wib * stride;
            const long long step = PWARPS * stride, wrap = P.Ws * stride - bw * stride;
            const int nb = P.bands;
            float4* dstp = box + wib * 32 + lane;
            int col = wib;                              // pixel p = wib, wib + 16, ... walks the box row by row
            for (int p = wib; p < nbox; p += PWARPS, col += PWARPS, rec += step, dstp += PWARPS * 32) {
                while (col >= bw) {
                    col -= bw;
                    rec += wrap;
                }
                if (SRC_VEC) {
                    cp_async16(dstp, rec);
                } else {
                    float* d = reinterpret_cast<float*>(dstp);
                    cp_async4(d, rec);
                    if (b + 1 < nb) cp_async4(d + 1, rec + 1);
                    if (b + 2 < nb) cp_async4(d + 2, rec + 2);
                    if (b + 3 < nb) cp_async4(d + 3, rec + 3);
                }
            }
        }
        cp_async_commit();
    };

    if (nstages > 0) {
        prepare_tile(0);
        issue_copy(0);
    }
    for (long long s = 0; s < nstages; ++s) {
        const long long it = s / G;
        const int g = (int)(s - it * G);
        if (s + 1 < nstages) {
            if (g == G - 1) prepare_tile(it + 1);      // the other TileState slot: nobody reads it any more
            issue_copy(s + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const TileState& T = s_tile[it & 1];
        float4* box = dyn + (size_t)(s & 1) * BOX_CAP * 32;
        unsigned char* cls = s_cls[s & 1];
        const int b = g * GB + 4 * lane;
        const bool act = b < P.bands;
        const bool staged = T.staged != 0;
        const int bw = T.bw, nbox = staged ? bw * T.bh : 0;
        // ---- classify the staged pixels (0 clean, 1 fill, 2 band-specific nodata); pad words take band b's value
        bool dirty = false;
        if (act && b + 3 >= P.bands) {                  // the lane holding the last, partial vector of the spectrum
            for (int p = wib; p < nbox; p += PWARPS) box[p * 32 + lane] = pad_fix(box[p * 32 + lane], b, P.bands);
        }
        bool allfill = nbox > 0;
        {
            // class 0: no nodata and every value finite (the lean loop may then multiply any of them by a zero weight);
            // 1: every band is nodata (a fill pixel); 2: anything else
            const float4* bp = box + wib * 32 + lane;
            for (int p = wib; p < nbox; p += PWARPS, bp += PWARPS * 32) {
                const float4 x = *bp;
                const float z = fmaf(x.x, 0.f, fmaf(x.y, 0.f, fmaf(x.z, 0.f, x.w * 0.f)));      // NaN iff a value is not finite
                const bool any = act && ((has_nd && (x.x == nd || x.y == nd || x.z == nd || x.w == nd)) || !(z == 0.f));
                int c = 0;
                if (__any_sync(FULL, any)) {
                    const bool all = !act || (x.x == nd && x.y == nd && x.z == nd && x.w == nd);
                    c = (has_nd && __all_sync(FULL, all)) ? 1 : 2;
                    dirty = true;
                }
                allfill = allfill && c == 1;
                if (lane == 0) cls[p] = (unsigned char)c;
            }
        }
        const bool box_clean = __syncthreads_or(dirty) == 0;
        // every staged pixel is fill (the tile lies outside the swath): nothing to resample (uniform: second reduction)
        const bool box_fill = !box_clean && staged && __syncthreads_and(allfill) != 0;
        // ---- resample
        const long long tile = blockIdx.x + it * gridDim.x;
        const long long ty = tile / tiles_x, tx = tile - ty * tiles_x;
        for (int q = wib; q < npix_tile; q += PWARPS) {
            const long long r = ty * TH + (q >> tw_shift), c = tx * TW + (q & (TW - 1));      // TW is a power of two
            if (r >= P.Hd || c >= P.Wd) continue;
            const int* M = T.meta[q];
            float* outp = P.dst + (r * P.Wd + c) * P.dst_pix_stride + b;
            float4 o = make_float4(dnd, dnd, dnd, dnd);
            if (M[0]) {
                const int jlo = M[1], jhi = M[2], klo = M[3], khi = M[4];
                const float* tw = T.tw[q];
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, wu = 0.f;
                if (box_fill) {
                    // nothing valid under any tap: the weight sum stays 0 -> nodata
                } else if (staged && box_clean && (M[7] & 2) && (ntx == 4 || ntx == 6 || ntx == 8)) {
                    // lean loop: clean finite box, whole rows inside the source -> fully unrolled rows, weights by vector
                    // loads, the weight sum precomputed per pixel (zero-weight columns contribute exactly 0)
                    const float4* rowb = box + ((M[6] + jlo - T.by0) * bw + (M[5] - T.bx0)) * 32 + lane;
                    const float* twj = tw + jlo * ntx;
                    const int rstep = bw * 32;
                    if (ntx == 6) {
                        for (int j = jlo; j < jhi; ++j, rowb += rstep, twj += 6) row_fma<6>(rowb, twj, a0, a1, a2, a3);
                    } else if (ntx == 4) {
                        for (int j = jlo; j < jhi; ++j, rowb += rstep, twj += 4) row_fma<4>(rowb, twj, a0, a1, a2, a3);
                    } else {
                        for (int j = jlo; j < jhi; ++j, rowb += rstep, twj += 8) row_fma<8>(rowb, twj, a0, a1, a2, a3);
                    }
                    wu = T.wsum[q];
                } else if (staged && box_clean && (M[7] & 1)) {
                    const float4* bp = box + ((M[6] - T.by0) * bw + (M[5] - T.bx0)) * 32 + lane;
                    for (int j = jlo; j < jhi; ++j) {
                        const float4* rowb = bp + j * bw * 32;
                        const float* twj = tw + j * ntx;
#pragma unroll 2
                        for (int k = klo; k < khi; ++k) {
                            const float w = twj[k];
                            const float4 v = rowb[k * 32];
                            a0 = fmaf(w, v.x, a0);
                            a1 = fmaf(w, v.y, a1);
                            a2 = fmaf(w, v.z, a2);
                            a3 = fmaf(w, v.w, a3);
                            wu += w;
                        }
                    }
                } else if (staged) {
                    const int base = (M[6] - T.by0) * bw + (M[5] - T.bx0);
                    for (int j = jlo; j < jhi; ++j) {
                        const int rowi = base + j * bw;
                        const float* twj = tw + j * ntx;
#pragma unroll 2
                        for (int k = klo; k < khi; ++k) {
                            const int pi = rowi + k;
                            const int cl = cls[pi];
                            const float w = twj[k];
                            const float4 v = box[pi * 32 + lane];
                            if (cl == 1 || w == 0.f) continue;
                            if (cl == 0) {
                                a0 = fmaf(w, v.x, a0);
                                a1 = fmaf(w, v.y, a1);
                                a2 = fmaf(w, v.z, a2);
                                a3 = fmaf(w, v.w, a3);
                                wu += w;
                            } else {
                                const float w0 = v.x == nd ? 0.f : w, w1q = v.y == nd ? 0.f : w;
                                const float w2 = v.z == nd ? 0.f : w, w3 = v.w == nd ? 0.f : w;
                                a0 = fmaf(w0, v.x, a0);
                                a1 = fmaf(w1q, v.y, a1);
                                a2 = fmaf(w2, v.z, a2);
                                a3 = fmaf(w3, v.w, a3);
                                m0 += w0;
                                m1 += w1q;
                                m2 += w2;
                                m3 += w3;
                            }
                        }
                    }
                } else {    // box too large for the staging buffer: the same taps straight from global memory
                    for (int j = jlo; j < jhi; ++j) {
                        const float* rowp = P.src + (((long long)M[6] + j) * P.Ws + M[5]) * P.src_pix_stride;
                        for (int k = klo; k < khi; ++k) {
                            const float w = tw[j * ntx + k];
                            if (w == 0.f) continue;
                            const float4 v = load_group<SRC_VEC>(rowp + k * P.src_pix_stride, b, P.bands, act);
                            const bool h = has_nd && act;
                            const float w0 = (h && v.x == nd) ? 0.f : w, w1q = (h && v.y == nd) ? 0.f : w;
                            const float w2 = (h && v.z == nd) ? 0.f : w, w3 = (h && v.w == nd) ? 0.f : w;
                            a0 = fmaf(w0, v.x, a0);
                            a1 = fmaf(w1q, v.y, a1);
                            a2 = fmaf(w2, v.z, a2);
                            a3 = fmaf(w3, v.w, a3);
                            m0 += w0;
                            m1 += w1q;
                            m2 += w2;
                            m3 += w3;
                        }
                    }
                }
                const float s0 = wu + m0, s1 = wu + m1, s2 = wu + m2, s3 = wu + m3;
                o.x = s0 >= 1e-6f ? __fdiv_rn(a0, s0) : dnd;
                o.y = s1 >= 1e-6f ? __fdiv_rn(a1, s1) : dnd;
                o.z = s2 >= 1e-6f ? __fdiv_rn(a2, s2) : dnd;
                o.w = s3 >= 1e-6f ? __fdiv_rn(a3, s3) : dnd;
            }
            if (act) {
                if (DST_VEC) {
                    __stcs(reinterpret_cast<float4*>(outp), o);
                } else {
                    __stcs(outp, o.x);
                    if (b + 1 < P.bands) __stcs(outp + 1, o.y);
                    if (b + 2 < P.bands) __stcs(outp + 2, o.z);
                    if (b + 3 < P.bands) __stcs(outp + 3, o.w);
                }
            }
        }
        __syncthreads();        // buffer s & 1 and (after the tile's last group) its TileState are free again
    }
}

// ---------------------------------------------------------------------------- lane-per-pixel variant
// Lanes across DESTINATION PIXELS instead of bands.  A CTA owns a 32 x LROWS tile of the destination and walks the
// spectrum in groups of 32 bands (8 float4 "quads"); a thread owns one pixel of the tile for the whole tile: its window
// origin, its separable weights (registers, fully unrolled NT x NT window) and its weight sum are computed ONCE and
// serve every band of the pixel — in the band-per-lane kernels that per-pixel work is redone by a whole warp for every
// pixel and band group, and it dominated them.  Per group the box of the tile's windows is staged in shared memory
// TRANSPOSED ([quad][pixel], odd pitch): the coalesced global reads (lanes across the quads of two source pixels) become
// conflict-free shared-memory writes, and the taps of 32 neighbouring destination pixels — 32 nearly consecutive source
// pixels of one quad — are conflict-free 16-byte reads.  The box covers the windows UNCLIPPED (pixels outside the source
// are staged as zeros and carry zero weight), so tap addresses are base + j * bw + k with k an immediate.  Groups whose
// box holds no nodata and no non-finite value take a lean loop (5 instructions per tap and quad); the others select
// weights and values per element.  Three CTAs per SM (single buffer each): staging of one overlaps resampling of another.
constexpr int LROWS = 8;            // destination tile: LCOLS x LROWS pixels, one per thread and quad half
constexpr int LCOLS = 16;
constexpr int LQ = 8;               // quads (float4) per band group: 32 bands
constexpr int LBOX = 480;           // staged source pixels per tile (pitch LBOX + 1, odd): 8 * 481 * 16 B = 61.6 KB, 3 CTAs per SM
constexpr int LPITCH = LBOX + 1;

template <int NT, bool DST_VEC>
__global__ void __launch_bounds__(256, 3) warp_lane_kernel(const WarpParams P, const int src_vec) {
    extern __shared__ __align__(16) float4 lbox[];           // [LQ][LPITCH]
    __shared__ int s_mm[4];
    __shared__ float s_wy[NT][256];                          // row weights per thread (the row loop stays rolled)
    __shared__ unsigned int s_pix[LBOX];                     // source pixel index of every box pixel, 0xffffffff outside
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    // a quarter-warp (the unit of a 16-byte shared-memory access) = 8 ROWS of one destination column: their taps sit
    // ~one box row apart, and the box pitch is odd, so the 8 addresses fall into 8 different 16-byte bank groups
    // (neighbouring columns would be 1.2 source pixels apart at the granule's x-scale: two of eight always collide)
    const int prow = lane & 7;                               // my pixel: row prow, column pcol of the tile
    const int pcol = (wib & 3) * 4 + (lane >> 3);
    const int qhalf = wib >> 2;                              // 8 warps = 4 column groups x 2 halves of the group's quads
    constexpr int R = NT / 2;
    const long long tiles_x = (P.Wd + LCOLS - 1) / LCOLS, tiles_y = (P.Hd + LROWS - 1) / LROWS, ntiles = tiles_x * tiles_y;
    const int nvec = (P.bands + 3) >> 2;
    const int ngroups = (nvec + LQ - 1) / LQ;
    const float nd = P.nodata, dnd = P.dst_nodata;
    const bool has_nd = P.has_nodata != 0;
    const long long stride = P.src_pix_stride;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const long long r = ty * LROWS + prow, c = tx * LCOLS + pcol;
        // ---- my pixel: source coordinates, window origin, weights (once per tile)
        double px = -1.0, py = -1.0;
        if (r < P.Hd && c < P.Wd) {
            const double2 xy = __ldg(reinterpret_cast<const double2*>(P.coords) + (r * P.Wd + c));
            px = xy.x;
            py = xy.y;
        }
        const bool inside = px >= 0.0 && px < (double)P.Ws && py >= 0.0 && py < (double)P.Hs;
        const double fxp = floor(px - 0.5), fyp = floor(py - 0.5);
        const int wx0 = inside ? (int)fxp + 1 - R : 0, wy0 = inside ? (int)fyp + 1 - R : 0;     // first tap of my window
        float wx[NT], wy[NT];
        float wsum_all = 0.f;
        {
            const double ddx = px - 0.5 - fxp, ddy = py - 0.5 - fyp;
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                // tap k of the NT-wide register window is tap k + P.rx - R of the filter (NT >= 2 * radius)
                const int tk = k + 1 - R;
                const bool inx = inside && tk >= 1 - P.rx && tk <= P.rx && wx0 + k >= 0 && wx0 + k < P.Ws;
                const bool iny = inside && tk >= 1 - P.ry && tk <= P.ry && wy0 + k >= 0 && wy0 + k < P.Hs;
                wx[k] = inx ? (float)tap_weight(P.kind, ((double)tk - ddx) * P.fx) : 0.f;
                wy[k] = iny ? (float)tap_weight(P.kind, ((double)tk - ddy) * P.fy) : 0.f;
                sx += wx[k];
                sy += wy[k];
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
                for (int k = 0; k < NT; ++k) wsum_all += wy[j] * wx[k];
                s_wy[j][tid] = wy[j];       // read back by this thread only
            }
            (void)sx;
            (void)sy;
        }
        // ---- bounding box of the tile's windows (unclipped)
        if (tid < 4) s_mm[tid] = (tid & 1) ? -2147483647 : 2147483647;
        __syncthreads();
        if (qhalf == 0) {
            const unsigned int FULL = 0xffffffffu;
            const int mnx = __reduce_min_sync(FULL, inside ? wx0 : 2147483647), mxx = __reduce_max_sync(FULL, inside ? wx0 : -2147483647);
            const int mny = __reduce_min_sync(FULL, inside ? wy0 : 2147483647), mxy = __reduce_max_sync(FULL, inside ? wy0 : -2147483647);
            if (lane == 0) {
                atomicMin(&s_mm[0], mnx);
                atomicMax(&s_mm[1], mxx);
                atomicMin(&s_mm[2], mny);
                atomicMax(&s_mm[3], mxy);
            }
        }
        __syncthreads();
        const bool any_inside = s_mm[1] >= s_mm[0];
        const int bx0 = s_mm[0], by0 = s_mm[2];
        const int bw = any_inside ? ((s_mm[1] - s_mm[0] + NT) | 1) : 0, bh = any_inside ? s_mm[3] - s_mm[2] + NT : 0;   // odd pitch
        const int nbox = bw * bh;
        const bool staged = any_inside && nbox <= LBOX;
        const int mybase = (wy0 - by0) * bw + (wx0 - bx0);          // my window's first tap inside the box
        // where every box pixel lives in the source (the same for all band groups of the tile): computed once
        if (staged) {
            for (int p = tid; p < nbox; p += 256) {
                const int by = p / bw, bx = p - by * bw;
                const long long yy = (long long)by0 + by, xx = (long long)bx0 + bx;
                s_pix[p] = (yy >= 0 && yy < P.Hs && xx >= 0 && xx < P.Ws) ? (unsigned int)(yy * P.Ws + xx) : 0xffffffffu;
            }
        }
        __syncthreads();
        float* outp = P.dst + (r * P.Wd + c) * P.dst_pix_stride;

        for (int g = 0; g < ngroups; ++g) {
            const int q0 = g * LQ;
            // ---- stage the box: item i = pixel * LQ + quad; out-of-source pixels and quads beyond the spectrum are zeros
            bool dirty = false, notfill = false;
            if (staged) {
                // thread -> quad (tid & 7) of pixels (tid >> 3), + 32, ...: a warp reads 4 pixels x 128 contiguous bytes and
                // writes 8 runs of 4 consecutive slots (pitch odd: conflict-free); no division, one running box position
                const int q = tid & (LQ - 1);
                const int b = (q0 + q) * 4;
                const bool qok = b < P.bands;
                float4* dstq = lbox + q * LPITCH;
                constexpr int SU = 4;
                for (int p0 = tid >> 3; p0 < nbox; p0 += 32 * SU) {
                    float4 v[SU];
                    bool live[SU];
#pragma unroll
                    for (int u = 0; u < SU; ++u) {
                        const int p = p0 + 32 * u;
                        const unsigned int pix = p < nbox ? s_pix[p] : 0xffffffffu;
                        live[u] = qok && pix != 0xffffffffu;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live[u]) {
                            const float* rec = P.src + (long long)pix * stride + b;
                            if (src_vec) {
                                v[u] = __ldg(reinterpret_cast<const float4*>(rec));
                            } else {
                                v[u].x = __ldg(rec);
                                if (b + 1 < P.bands) v[u].y = __ldg(rec + 1);
                                if (b + 2 < P.bands) v[u].z = __ldg(rec + 2);
                                if (b + 3 < P.bands) v[u].w = __ldg(rec + 3);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < SU; ++u) {
                        const int p = p0 + 32 * u;
                        if (p >= nbox) break;
                        float4 x = v[u];
                        if (live[u]) {
                            x = pad_fix(x, b, P.bands);
                            const float z = fmaf(x.x, 0.f, fmaf(x.y, 0.f, fmaf(x.z, 0.f, x.w * 0.f)));
                            dirty = dirty || !(z == 0.f) || (has_nd && (x.x == nd || x.y == nd || x.z == nd || x.w == nd));
                            notfill = notfill || !has_nd || !(x.x == nd && x.y == nd && x.z == nd && x.w == nd);
                        }
                        dstq[p] = x;
                    }
                }
            }
            const bool clean = __syncthreads_or(dirty) == 0;
            // a group whose box holds nothing but nodata (the tile lies outside the swath): every weight sum is 0
            const bool allfill = staged && !clean && __syncthreads_or(notfill) == 0;
            // ---- resample my pixel for my half of the group's quads
            auto store_quad = [&](int b, const float4& o) {
                if (!(r < P.Hd && c < P.Wd) || b >= P.bands) return;
                if (DST_VEC) {
                    __stcs(reinterpret_cast<float4*>(outp + b), o);
                } else {
                    __stcs(outp + b, o.x);
                    if (b + 1 < P.bands) __stcs(outp + b + 1, o.y);
                    if (b + 2 < P.bands) __stcs(outp + b + 2, o.z);
                    if (b + 3 < P.bands) __stcs(outp + b + 3, o.w);
                }
            };
            const float4 fillq = make_float4(dnd, dnd, dnd, dnd);
            if (!inside || allfill) {
                for (int qq = 0; qq < LQ / 2; ++qq) store_quad((q0 + qhalf * (LQ / 2) + qq) * 4, fillq);
            } else if (staged && clean) {
                // lean loop, two quads at a time: one weight product serves eight FMAs, two independent accumulator sets
                const float inv_ok = wsum_all >= 1e-6f ? 1.f : 0.f;
                for (int qq = 0; qq < LQ / 2; qq += 2) {
                    const int q = qhalf * (LQ / 2) + qq;
                    const float4* wp = lbox + q * LPITCH + mybase;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll 1
                    for (int j = 0; j < NT; ++j) {
                        const float4* rp = wp + j * bw;
                        const float wyj = s_wy[j][tid];
#pragma unroll
                        for (int k = 0; k < NT; ++k) {
                            const float w = wyj * wx[k];
                            const float4 v = rp[k], u = rp[k + LPITCH];
                            a0 = fmaf(w, v.x, a0);
                            a1 = fmaf(w, v.y, a1);
                            a2 = fmaf(w, v.z, a2);
                            a3 = fmaf(w, v.w, a3);
                            c0 = fmaf(w, u.x, c0);
                            c1 = fmaf(w, u.y, c1);
                            c2 = fmaf(w, u.z, c2);
                            c3 = fmaf(w, u.w, c3);
                        }
                    }
                    float4 o = fillq, o2 = fillq;
                    if (inv_ok != 0.f) {
                        o = make_float4(__fdiv_rn(a0, wsum_all), __fdiv_rn(a1, wsum_all), __fdiv_rn(a2, wsum_all), __fdiv_rn(a3, wsum_all));
                        o2 = make_float4(__fdiv_rn(c0, wsum_all), __fdiv_rn(c1, wsum_all), __fdiv_rn(c2, wsum_all), __fdiv_rn(c3, wsum_all));
                    }
                    store_quad((q0 + q) * 4, o);
                    store_quad((q0 + q + 1) * 4, o2);
                }
            } else {
                for (int qq = 0; qq < LQ / 2; ++qq) {
                    const int q = qhalf * (LQ / 2) + qq;
                    const int b = (q0 + q) * 4;
                    if (b >= P.bands) break;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll 1
                    for (int j = 0; j < NT; ++j) {
                        const float wyj = s_wy[j][tid];
#pragma unroll
                        for (int k = 0; k < NT; ++k) {
                            const float w = wyj * wx[k];
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (staged) {
                                v = lbox[q * LPITCH + mybase + j * bw + k];
                            } else if (w != 0.f) {      // box too large for the buffer: the tap from global memory
                                const float* rec = P.src + (((long long)wy0 + j) * P.Ws + (wx0 + k)) * stride + b;
                                if (src_vec) {
                                    v = __ldg(reinterpret_cast<const float4*>(rec));
                                } else {
                                    v.x = __ldg(rec);
                                    if (b + 1 < P.bands) v.y = __ldg(rec + 1);
                                    if (b + 2 < P.bands) v.z = __ldg(rec + 2);
                                    if (b + 3 < P.bands) v.w = __ldg(rec + 3);
                                }
                                v = pad_fix(v, b, P.bands);
                            }
                            // a zero weight contributes nothing (not even a NaN); nodata is skipped per element
                            const bool z = w == 0.f;
                            const bool s0 = z || (has_nd && v.x == nd), s1 = z || (has_nd && v.y == nd);
                            const bool s2 = z || (has_nd && v.z == nd), s3 = z || (has_nd && v.w == nd);
                            a0 = fmaf(s0 ? 0.f : w, s0 ? 0.f : v.x, a0);
                            a1 = fmaf(s1 ? 0.f : w, s1 ? 0.f : v.y, a1);
                            a2 = fmaf(s2 ? 0.f : w, s2 ? 0.f : v.z, a2);
                            a3 = fmaf(s3 ? 0.f : w, s3 ? 0.f : v.w, a3);
                            m0 += s0 ? 0.f : w;
                            m1 += s1 ? 0.f : w;
                            m2 += s2 ? 0.f : w;
                            m3 += s3 ? 0.f : w;
                        }
                    }
                    float4 o;
                    o.x = m0 >= 1e-6f ? __fdiv_rn(a0, m0) : dnd;
                    o.y = m1 >= 1e-6f ? __fdiv_rn(a1, m1) : dnd;
                    o.z = m2 >= 1e-6f ? __fdiv_rn(a2, m2) : dnd;
                    o.w = m3 >= 1e-6f ? __fdiv_rn(a3, m3) : dnd;
                    store_quad(b, o);
                }
            }
            __syncthreads();            // the next group overwrites the box
        }
    }
}

// ---------------------------------------------------------------------------- host side
int fill_params(WarpParams& P, const hsr_warp_geo_t* geo, long long Hs, long long Ws, long long Hd, long long Wd,
                int kernel) {
    HSR_REQUIRE(geo, HSR_EINVAL, "null geo pointer");
    HSR_REQUIRE(kernel == 1 || kernel == 2, HSR_EINVAL, "kernel must be 1 (bilinear) or 2 (cubic), got %d", kernel);
    HSR_REQUIRE(geo->utm_zone >= 0 && geo->utm_zone <= 60, HSR_ERANGE, "utm_zone = %d outside [0, 60]", geo->utm_zone);
    const double* s = geo->src_gt;
    const double det = s[1] * s[5] - s[2] * s[4];
    HSR_REQUIRE(det != 0.0 && det == det, HSR_EINVAL, "source geotransform is singular");
    for (int i = 0; i < 6; ++i) P.dgt[i] = geo->dst_gt[i];
    P.sx0 = s[0];
    P.sy0 = s[3];
    P.inv[0] = s[5] / det;
    P.inv[1] = -s[2] / det;
    P.inv[2] = -s[4] / det;
    P.inv[3] = s[1] / det;
    P.utm = geo->utm_zone > 0 ? 1 : 0;
    if (P.utm) {    // WGS-84, k0 = 0.9996; Krueger series to n^6 (Karney 2011, eq. 36)
        const double a = 6378137.0, f = 1.0 / 298.257223563;
        const double n = f / (2.0 - f), n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
        const double A = a / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0);
        P.k0A_inv = 1.0 / (0.9996 * A);
        P.e = sqrt(f * (2.0 - f));
        P.e2m = 1.0 - f * (2.0 - f);
        P.lon0_deg = -183.0 + 6.0 * geo->utm_zone;
        P.false_northing = geo->south ? 10000000.0 : 0.0;
        P.beta[0] = n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800;
        P.beta[1] = n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720;
        P.beta[2] = 17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720;
        P.beta[3] = 4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600;
        P.beta[4] = 4583 * n5 / 161280 - 108847 * n6 / 3991680;
        P.beta[5] = 20648693 * n6 / 638668800;
    }
    P.Hs = Hs, P.Ws = Ws, P.Hd = Hd, P.Wd = Wd;
    P.kind = kernel;
    const int r0 = kernel == 2 ? 2 : 1;
    const double xs = geo->xscale > 0.0 ? geo->xscale : 1.0, ys = geo->yscale > 0.0 ? geo->yscale : 1.0;
    P.fx = xs < 1.0 ? xs : 1.0;
    P.fy = ys < 1.0 ? ys : 1.0;
    const double rxd = xs < 1.0 ? ceil(r0 / xs) : (double)r0, ryd = ys < 1.0 ? ceil(r0 / ys) : (double)r0;
    HSR_REQUIRE(rxd <= MAX_TAPS / 2 && ryd <= MAX_TAPS / 2, HSR_ERANGE,
                "scale (%g, %g) needs a filter radius beyond %d taps", xs, ys, MAX_TAPS / 2);
    P.rx = (int)rxd;
    P.ry = (int)ryd;
    return HSR_OK;
}

}  // namespace

size_t warp_workspace(long long Hd, long long Wd) { return Hd > 0 && Wd > 0 ? (size_t)Hd * (size_t)Wd * 16 : 0; }

int warp_impl(const float* src, long long Hs, long long Ws, int bands, long long src_pix_stride, const hsr_warp_geo_t* geo,
              int kernel, int has_nodata, float nodata, float dst_nodata, long long Hd, long long Wd, float* dst,
              long long dst_pix_stride, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    HSR_REQUIRE(src && dst, HSR_EINVAL, "null src / dst pointer");
    HSR_REQUIRE(Hs > 0 && Ws > 0 && bands > 0 && Hd >= 0 && Wd >= 0, HSR_EINVAL, "bad shape");
    HSR_REQUIRE(Hs < 2147483000LL && Ws < 2147483000LL && Hs * Ws < (1LL << 40) && Hd * Wd < (1LL << 40), HSR_ERANGE,
                "grid too large (source sides are limited to 2^31 pixels, grids to 2^40)");
    HSR_REQUIRE(src_pix_stride >= bands && dst_pix_stride >= bands, HSR_EINVAL, "pixel stride < bands");
    HSR_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0, HSR_EALIGN,
                "src / dst not 4-byte aligned");
    WarpParams P{};
    int rc = fill_params(P, geo, Hs, Ws, Hd, Wd, kernel);
    if (rc != HSR_OK) return rc;
    if (Hd == 0 || Wd == 0) return HSR_OK;
    P.src = src;
    P.src_pix_stride = src_pix_stride;
    P.bands = bands;
    P.dst = dst;
    P.dst_pix_stride = dst_pix_stride;
    P.has_nodata = has_nodata ? 1 : 0;
    P.nodata = nodata;
    P.dst_nodata = dst_nodata;
    // destination tile: the largest of 8x4, 4x4, 4x2, 2x2, 1x1 whose tap footprint fits the staging buffer
    // (estimated from the scales plus two pixels of slack for rotation; the kernel checks the real box per tile)
    const double xs = geo->xscale > 0.0 ? geo->xscale : 1.0, ys = geo->yscale > 0.0 ? geo->yscale : 1.0;
    const int cand[5][2] = {{8, 4}, {4, 4}, {4, 2}, {2, 2}, {1, 1}};
    P.tile_w = P.tile_h = 1;
    for (int i = 0; i < 5; ++i) {
        const double fw = cand[i][0] / xs + 2 * P.rx + 2, fh = cand[i][1] / ys + 2 * P.ry + 2;
        if (fw * fh <= BOX_CAP) {
            P.tile_w = cand[i][0];
            P.tile_h = cand[i][1];
            break;
        }
    }
    const long long ntiles = ((Hd + P.tile_h - 1) / P.tile_h) * ((Wd + P.tile_w - 1) / P.tile_w);
    if (workspace) {    // source coordinates of every destination pixel, once, instead of per tile and band group
        HSR_REQUIRE(workspace_bytes >= warp_workspace(Hd, Wd), HSR_EINVAL, "workspace too small (%zu < %zu bytes)",
                    workspace_bytes, warp_workspace(Hd, Wd));
        HSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, HSR_EALIGN, "workspace not 16-byte aligned");
        P.coords = static_cast<double*>(workspace);
        long long cb = (Hd * Wd + 32 * WARPS - 1) / (32 * WARPS);
        const long long ccap = (long long)device_sm_count() * 8;
        warp_coords_kernel<<<(unsigned int)(cb < ccap ? cb : ccap), 32 * WARPS, 0, stream>>>(P);
        HSR_CUDA(cudaGetLastError());
    }
    const int groups = (bands + GB - 1) / GB;
    const bool src_vec = (src_pix_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const bool dst_vec = (dst_pix_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    const bool fast = P.coords && 4 * P.rx * P.ry <= PTAPS && P.tile_w * P.tile_h <= PTILE && Ws < 2147483647LL &&
                      Hs < 2147483647LL;
    if (P.coords && P.rx <= 4 && P.ry <= 4 && Hs * Ws < 4294967295LL && Hs < 2147483000LL && Ws < 2147483000LL &&
        exp_int("HSR_WARP_NO_LANE", 0, 0, 1) == 0) {
        // lane-per-pixel kernel: register window of NT x NT taps, NT = 2 * max radius rounded up to 2, 4, 6, 8
        const int rmax = P.rx > P.ry ? P.rx : P.ry;
        const long long ltiles = ((Wd + LCOLS - 1) / LCOLS) * ((Hd + LROWS - 1) / LROWS);
        long long blocks = (long long)device_sm_count() * 3;
        if (blocks > ltiles) blocks = ltiles;
        const size_t smem = (size_t)LQ * LPITCH * sizeof(float4);
        const int sv = src_vec ? 1 : 0;
#define HSR_LAUNCH_LANE(NTV)                                                                                              \
    do {                                                                                                                  \
        if (dst_vec) {                                                                                                    \
            { static int set__[HSR_MAX_DEVICES]; HSR_CUDA(ensure_dynamic_smem(warp_lane_kernel<NTV, true>, (int)smem, set__)); } \
            warp_lane_kernel<NTV, true><<<(unsigned int)blocks, 256, smem, stream>>>(P, sv);                              \
        } else {                                                                                                          \
            { static int set__[HSR_MAX_DEVICES]; HSR_CUDA(ensure_dynamic_smem(warp_lane_kernel<NTV, false>, (int)smem, set__)); } \
            warp_lane_kernel<NTV, false><<<(unsigned int)blocks, 256, smem, stream>>>(P, sv);                             \
        }                                                                                                                 \
    } while (0)
        if (rmax <= 1) HSR_LAUNCH_LANE(2);
        else if (rmax == 2) HSR_LAUNCH_LANE(4);
        else if (rmax == 3) HSR_LAUNCH_LANE(6);
        else HSR_LAUNCH_LANE(8);
#undef HSR_LAUNCH_LANE
        HSR_CUDA(cudaGetLastError());
        return HSR_OK;
    }
    if (fast) {   // radii beyond 4 (NT > 8): the band-per-lane pipeline
        // pipelined: one persistent CTA per SM, two staging buffers
        const size_t smem = (size_t)2 * BOX_CAP * 32 * sizeof(float4);
        long long blocks = device_sm_count();
        if (blocks > ntiles) blocks = ntiles;
#define HSR_LAUNCH_PIPE(SV, DV)                                                                                      \
    do {                                                                                                             \
        { static int set__[HSR_MAX_DEVICES]; HSR_CUDA(ensure_dynamic_smem(warp_pipe_kernel<SV, DV>, (int)smem, set__)); } \
        warp_pipe_kernel<SV, DV><<<(unsigned int)blocks, 32 * PWARPS, smem, stream>>>(P);                            \
    } while (0)
        if (src_vec && dst_vec) HSR_LAUNCH_PIPE(true, true);
        else if (src_vec) HSR_LAUNCH_PIPE(true, false);
        else if (dst_vec) HSR_LAUNCH_PIPE(false, true);
        else HSR_LAUNCH_PIPE(false, false);
#undef HSR_LAUNCH_PIPE
        HSR_CUDA(cudaGetLastError());
        return HSR_OK;
    }
    long long blocks = ntiles;
    const long long cap = (long long)device_sm_count() * 2 * 4;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)BOX_CAP * 32 * sizeof(float4);
    const dim3 grid((unsigned int)blocks, (unsigned int)groups);
#define HSR_LAUNCH_WARP(SV, DV)                                                                                      \
    do {                                                                                                             \
        { static int set__[HSR_MAX_DEVICES]; HSR_CUDA(ensure_dynamic_smem(warp_tile_kernel<SV, DV>, (int)smem, set__)); } \
        warp_tile_kernel<SV, DV><<<grid, 32 * WARPS, smem, stream>>>(P);                                             \
    } while (0)
    if (src_vec && dst_vec) HSR_LAUNCH_WARP(true, true);
    else if (src_vec) HSR_LAUNCH_WARP(true, false);
    else if (dst_vec) HSR_LAUNCH_WARP(false, true);
    else HSR_LAUNCH_WARP(false, false);
#undef HSR_LAUNCH_WARP
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int warp_coords_impl(const hsr_warp_geo_t* geo, long long Hd, long long Wd, double* coords, cudaStream_t stream) {
    HSR_REQUIRE(coords, HSR_EINVAL, "null coords pointer");
    HSR_REQUIRE(Hd >= 0 && Wd >= 0, HSR_EINVAL, "bad shape");
    WarpParams P{};
    int rc = fill_params(P, geo, 1, 1, Hd, Wd, 2);
    if (rc != HSR_OK) return rc;
    if (Hd == 0 || Wd == 0) return HSR_OK;
    P.coords = coords;
    long long blocks = (Hd * Wd + 32 * WARPS - 1) / (32 * WARPS);
    const long long cap = (long long)device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    warp_coords_kernel<<<(unsigned int)blocks, 32 * WARPS, 0, stream>>>(P);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

}  // namespace hsr
