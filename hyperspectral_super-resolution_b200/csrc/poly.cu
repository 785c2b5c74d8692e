// poly.cu — kernel 3 of the hot path: per-series polynomial regression
//   moments (block-then-grid fp64 reduction of the normal equations X^T X / X^T y)
//   -> warp-level solve of the column-scaled system -> fused Horner apply + mask + clip.
// Arithmetic follows np.polyfit as called at s2_emit/poly_regression.py:58-60 and
// apply_poly_rgb at s2_emit/poly_regression.py:65-84.
#include "hsr_common.cuh"

namespace hsr {

namespace {

constexpr int MAXDEG = HSR_MAX_POLY_DEG;
constexpr int MOM_THREADS = 256;
constexpr int MOM_MAX_BLOCKS = 1184;  // 148 SMs x 8 resident CTAs

__host__ __device__ inline int n_moments(int deg) { return 3 * deg + 2; }

int moments_blocks(long long n, int K) {
    long long per_series = (n + (long long)MOM_THREADS * 16 - 1) / ((long long)MOM_THREADS * 16);
    if (per_series < 1) per_series = 1;
    long long cap = MOM_MAX_BLOCKS / (K > 0 ? K : 1);
    if (cap < 1) cap = 1;
    if (per_series > cap) per_series = cap;
    return (int)per_series;
}

// One sample into the 3*DEG+2 sums: S_j = sum x^j (j = 0..2*DEG), T_j = sum x^j y (j = 0..DEG).
// Only the powers up to DEG are formed (DEG-1 multiplies); every higher S_j is a single fused
// multiply-add of two of them (x^a * x^b, a + b = j, rounded once), so a degree-2 sample costs 8
// fp64 operations.  x is fp32, so x and x^2 are exact in fp64.
template <int DEG>
__device__ __forceinline__ void accumulate(double (&acc)[3 * DEG + 2], float xf, float yf, bool use) {
    if (!use) return;
    const double x = (double)xf, y = (double)yf;
    double pw[DEG + 1];
    pw[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= DEG; ++j) pw[j] = pw[j - 1] * x;
    acc[0] += 1.0;
    acc[1] += x;
#pragma unroll
    for (int j = 2; j <= 2 * DEG; ++j) acc[j] = fma(pw[j / 2], pw[j - j / 2], acc[j]);
    acc[2 * DEG + 1] += y;
#pragma unroll
    for (int j = 1; j <= DEG; ++j) acc[2 * DEG + 1 + j] = fma(pw[j], y, acc[2 * DEG + 1 + j]);
}

// grid = (blocks_per_series, K).  Each thread keeps 3*DEG+2 fp64 accumulators, the block
// reduces them by warp shuffles + shared memory, and writes ONE partial row; the finalize
// kernel then sums the rows of a series in a fixed order (bit-reproducible).
template <int DEG>
__global__ void __launch_bounds__(MOM_THREADS) poly_moments_kernel(
    const float* __restrict__ x, long long x_k_stride, long long x_n_stride, const float* __restrict__ y,
    long long y_k_stride, long long y_n_stride, const uint8_t* __restrict__ mask, long long mask_k_div, long long mask_k_mod, long long n,
    double* __restrict__ partial) {
    constexpr int M = 3 * DEG + 2;
    const int k = blockIdx.y;
    const float* xk = x + (long long)k * x_k_stride;
    const float* yk = y + (long long)k * y_k_stride;
    const uint8_t* mk = mask ? mask + (((long long)k / mask_k_div) % mask_k_mod) * n : nullptr;

    double acc[M];
#pragma unroll
    for (int j = 0; j < M; ++j) acc[j] = 0.0;

    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;

    const bool vec = x_n_stride == 1 && y_n_stride == 1 &&
                     ((reinterpret_cast<uintptr_t>(xk) | reinterpret_cast<uintptr_t>(yk)) & 15) == 0 &&
                     (mk == nullptr || (reinterpret_cast<uintptr_t>(mk) & 3) == 0);
    long long done = 0;
    if (vec) {
        const long long n4 = n >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(xk);
        const float4* y4 = reinterpret_cast<const float4*>(yk);
        const uchar4* m4 = reinterpret_cast<const uchar4*>(mk);
        for (long long i = tid; i < n4; i += nthreads) {
            const float4 xv = __ldg(x4 + i);
            const float4 yv = __ldg(y4 + i);
            uchar4 mv = make_uchar4(1, 1, 1, 1);
            if (mk) mv = __ldg(m4 + i);
            accumulate<DEG>(acc, xv.x, yv.x, mv.x && finite_f32(xv.x) && finite_f32(yv.x));
            accumulate<DEG>(acc, xv.y, yv.y, mv.y && finite_f32(xv.y) && finite_f32(yv.y));
            accumulate<DEG>(acc, xv.z, yv.z, mv.z && finite_f32(xv.z) && finite_f32(yv.z));
            accumulate<DEG>(acc, xv.w, yv.w, mv.w && finite_f32(xv.w) && finite_f32(yv.w));
        }
        done = n4 << 2;
    }
    for (long long i = done + tid; i < n; i += nthreads) {
        const float xv = __ldg(xk + i * x_n_stride);
        const float yv = __ldg(yk + i * y_n_stride);
        const bool use = (mk == nullptr || mk[i]) && finite_f32(xv) && finite_f32(yv);
        accumulate<DEG>(acc, xv, yv, use);
    }

    __shared__ double red[MOM_THREADS / 32][M];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double s = warp_sum(acc[j]);
        if (lane == 0) red[warp][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < M) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < MOM_THREADS / 32; ++w) s += red[w][threadIdx.x];
        partial[((long long)k * gridDim.x + blockIdx.x) * M + threadIdx.x] = s;
    }
}

__global__ void poly_moments_finalize_kernel(const double* __restrict__ partial, int nblocks, int M,
                                             double* __restrict__ moments) {
    const int k = blockIdx.x;
    const int j = threadIdx.x;
    if (j >= M) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[((long long)k * nblocks + b) * M + j];
    moments[(long long)k * M + j] = s;
}

// One warp solves one series.  Normal equations G c = r with G_ij = S_{i+j}, r_i = T_i, scaled by
// s_j = sqrt(S_{2j}) (the column norms np.polyfit divides its Vandermonde by), solved by
// Gaussian elimination with partial pivoting; lane j owns column j of the augmented matrix.
// A: [MAXDEG+1][MAXDEG+2] and scale: [MAXDEG+1] doubles of shared memory private to the warp.
// out[0..deg]: highest power first (np.polyfit order); every lane < deg+1 writes one entry.
__device__ __forceinline__ void solve_series(const double* __restrict__ mom, int deg, long long min_count,
                                             double (*A)[MAXDEG + 2], double* scale, double* out, int lane) {
    const int N = deg + 1;
    const double count = mom[0];
    if (!(count >= (double)min_count)) {  // identity: poly_regression.py:38-41
        if (lane < N) out[lane] = (lane == N - 2) ? 1.0 : 0.0;
        return;
    }
    if (lane < N) scale[lane] = sqrt(mom[2 * lane]);
    __syncwarp();
    if (lane <= N) {
        for (int i = 0; i < N; ++i) {
            double v;
            if (lane < N)
                v = mom[i + lane] / (scale[i] * scale[lane]);
            else
                v = mom[2 * deg + 1 + i] / scale[i];
            A[i][lane] = v;
        }
    }
    __syncwarp();
    for (int c = 0; c < N; ++c) {
        // pivot search (every lane computes the same answer)
        int piv = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < N; ++r) {
            const double v = fabs(A[r][c]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        __syncwarp();
        if (lane <= N && piv != c) {
            const double t = A[c][lane];
            A[c][lane] = A[piv][lane];
            A[piv][lane] = t;
        }
        __syncwarp();
        const double inv = 1.0 / A[c][c];
        __syncwarp();
        if (lane <= N) A[c][lane] *= inv;
        __syncwarp();
        double f[MAXDEG + 1];  // column c, read by every lane before lane c rewrites it
#pragma unroll
        for (int r = 0; r <= MAXDEG; ++r) f[r] = (r < N) ? A[r][c] : 0.0;
        __syncwarp();
        if (lane <= N) {
            const double pc = A[c][lane];
#pragma unroll
            for (int r = 0; r <= MAXDEG; ++r)
                if (r < N && r != c) A[r][lane] = fma(-f[r], pc, A[r][lane]);
        }
        __syncwarp();
    }
    // un-scale; np.polyfit order is highest power first
    if (lane < N) out[deg - lane] = A[lane][N] / scale[lane];
}

__global__ void __launch_bounds__(128) poly_solve_kernel(const double* __restrict__ moments, int K, int deg,
                                                         long long min_count, double* __restrict__ coeffs) {
    __shared__ double A[4][MAXDEG + 1][MAXDEG + 2];
    __shared__ double scale[4][MAXDEG + 1];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + w;
    if (k >= K) return;
    solve_series(moments + (long long)k * (3 * deg + 2), deg, min_count, A[w], scale[w],
                 coeffs + (long long)k * (deg + 1), lane);
}

template <int DEG>
__device__ __forceinline__ float horner_clip(float xf, const double (&c)[DEG + 1], bool use, float lo, float hi) {
    float r = xf;
    if (use) {
        const double x = (double)xf;
        double acc = 0.0;  // np.polyval starts from zeros_like(x): 0*Inf = NaN is kept
#pragma unroll
        for (int j = 0; j <= DEG; ++j) acc = fma(acc, x, c[j]);
        r = (float)acc;
    }
    if (lo <= hi) r = r < lo ? lo : (r > hi ? hi : r);  // NaN compares false twice and survives, as np.clip
    return r;
}

// np.polyval evaluates Horner as y = y*x + c (mul then add, two roundings); an fma differs
// by < 1 ulp of fp64, far below the fp32 result's resolution.
template <int DEG>
__global__ void __launch_bounds__(256) poly_apply_kernel(const float* __restrict__ x, long long x_k_stride,
                                                         long long x_n_stride, const double* __restrict__ coeffs,
                                                         const uint8_t* __restrict__ mask, long long mask_k_div,
                                                         long long mask_k_mod, long long n, float lo, float hi,
                                                         float* __restrict__ out,
                                                         long long out_k_stride, long long out_n_stride) {
    const int k = blockIdx.y;
    const float* xk = x + (long long)k * x_k_stride;
    float* ok = out + (long long)k * out_k_stride;
    const uint8_t* mk = mask ? mask + (((long long)k / mask_k_div) % mask_k_mod) * n : nullptr;
    double c[DEG + 1];
#pragma unroll
    for (int j = 0; j <= DEG; ++j) c[j] = coeffs[(long long)k * (DEG + 1) + j];

    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const bool vec = x_n_stride == 1 && out_n_stride == 1 &&
                     ((reinterpret_cast<uintptr_t>(xk) | reinterpret_cast<uintptr_t>(ok)) & 15) == 0 &&
                     (mk == nullptr || (reinterpret_cast<uintptr_t>(mk) & 3) == 0);
    long long done = 0;
    if (vec) {
        const long long n4 = n >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(xk);
        float4* o4 = reinterpret_cast<float4*>(ok);
        const uchar4* m4 = reinterpret_cast<const uchar4*>(mk);
        for (long long i = tid; i < n4; i += nthreads) {
            const float4 xv = __ldg(x4 + i);
            uchar4 mv = make_uchar4(1, 1, 1, 1);
            if (mk) mv = __ldg(m4 + i);
            float4 r;
            r.x = horner_clip<DEG>(xv.x, c, mv.x != 0, lo, hi);
            r.y = horner_clip<DEG>(xv.y, c, mv.y != 0, lo, hi);
            r.z = horner_clip<DEG>(xv.z, c, mv.z != 0, lo, hi);
            r.w = horner_clip<DEG>(xv.w, c, mv.w != 0, lo, hi);
            o4[i] = r;
        }
        done = n4 << 2;
    }
    for (long long i = done + tid; i < n; i += nthreads) {
        const float xv = __ldg(xk + i * x_n_stride);
        ok[i * out_n_stride] = horner_clip<DEG>(xv, c, mk == nullptr || mk[i] != 0, lo, hi);
    }
}

// mask[i] = valid[i] && all_k finite(x[k,i]) && x[gate_k,i] > gate_gt   (poly_regression.py:106)
__global__ void __launch_bounds__(256) fit_mask_kernel(const float* __restrict__ x, long long x_k_stride, long long n,
                                                       int K, const uint8_t* __restrict__ valid, int gate_k,
                                                       float gate_gt, uint8_t* __restrict__ mask) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        bool m = valid ? valid[i] != 0 : true;
        for (int k = 0; k < K; ++k) m = m && finite_f32(__ldg(x + (long long)k * x_k_stride + i));
        if (gate_k >= 0) m = m && (__ldg(x + (long long)gate_k * x_k_stride + i) > gate_gt);
        mask[i] = m ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Fused fit mask + moments: ONE pass over the K pseudo-S2 planes and the K reference planes.
//   mask[g, i] = valid[g, i] & all_k isfinite(x[k, g, i]) & (x[gate_k, g, i] > gate_gt)  (poly_regression.py:106)
//   moments[k, g, :] over the samples with mask & isfinite(y[k, g, i])                   (:35-36, :58-60)
// One warp per band k (block = K warps), grid = (blocks per group, G groups).  Per iteration a block
// takes 256 pixels; every lane holds 8 of them for its band (two 16-byte loads of x, two of y), the
// bands exchange their 8 finite bits through shared memory (double-buffered, one __syncthreads per
// iteration) and accumulate 3*DEG+2 fp64 sums privately.  Partial rows are reduced in a fixed order
// by poly_moments_finalize_kernel (bit-reproducible).
struct FitParams {
    const float* x;
    long long xks, xgs;
    const float* y;
    long long yks, ygs;
    const uint8_t* valid;
    uint8_t* mask;
    long long n;
    int K, G, gate_k;
    float gate_gt;
    int vecx, vecy, vecm;  // 16-byte loads of x / y planes, 4-byte accesses of valid / mask allowed (alignment)
    double* partial;
};

constexpr int FIT_PX = 256;  // pixels per block iteration: 32 lanes x 2 x 4

template <int DEG>
__global__ void __launch_bounds__(32 * HSR_MAX_SRF_BANDS) fit_moments_kernel(const FitParams P) {
    constexpr int M = 3 * DEG + 2;
    __shared__ unsigned int fin[2][HSR_MAX_SRF_BANDS][32];
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y;
    const long long n = P.n;
    const float* xk = P.x + (long long)k * P.xks + (long long)g * P.xgs;
    const float* yk = P.y + (long long)k * P.yks + (long long)g * P.ygs;
    const uint8_t* vg = P.valid ? P.valid + (long long)g * n : nullptr;
    uint8_t* mg = P.mask ? P.mask + (long long)g * n : nullptr;
    const float qnan = __int_as_float(0x7fc00000);

    double acc[M];
#pragma unroll
    for (int j = 0; j < M; ++j) acc[j] = 0.0;

    // loads of one iteration (8 pixels of this lane's band + their valid bits for the band-0 warp)
    auto fetch = [&](long long base, float (&xv)[8], float (&yv)[8], unsigned int& vb) {
        const long long i0 = base + lane * 4, i1 = i0 + 128;  // my two groups of 4 pixels
        vb = 0xffu;
        if (base + FIT_PX <= n) {
            if (P.vecx) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(xk + i0));
                const float4 b = __ldg(reinterpret_cast<const float4*>(xk + i1));
                xv[0] = a.x, xv[1] = a.y, xv[2] = a.z, xv[3] = a.w, xv[4] = b.x, xv[5] = b.y, xv[6] = b.z, xv[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) xv[j] = __ldg(xk + (j < 4 ? i0 : i1) + (j & 3));
            }
            if (P.vecy) {
                const float4 c = __ldcs(reinterpret_cast<const float4*>(yk + i0));
                const float4 d = __ldcs(reinterpret_cast<const float4*>(yk + i1));
                yv[0] = c.x, yv[1] = c.y, yv[2] = c.z, yv[3] = c.w, yv[4] = d.x, yv[5] = d.y, yv[6] = d.z, yv[7] = d.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) yv[j] = __ldg(yk + (j < 4 ? i0 : i1) + (j & 3));
            }
            if (k == 0 && vg) {
                if (P.vecm) {
                    const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(vg + i0));
                    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(vg + i1));
                    vb = (u.x ? 1u : 0u) | (u.y ? 2u : 0u) | (u.z ? 4u : 0u) | (u.w ? 8u : 0u) | (v.x ? 16u : 0u) |
                         (v.y ? 32u : 0u) | (v.z ? 64u : 0u) | (v.w ? 128u : 0u);
                } else {
                    vb = 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) vb |= vg[(j < 4 ? i0 : i1) + (j & 3)] ? (1u << j) : 0u;
                }
            }
        } else {
            vb = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long i = (j < 4 ? i0 : i1) + (j & 3);
                const bool in = i < n;
                xv[j] = in ? __ldg(xk + i) : qnan;  // out of range: not finite, never used
                yv[j] = in ? __ldg(yk + i) : qnan;
                if (in && (k != 0 || vg == nullptr || vg[i] != 0)) vb |= 1u << j;
            }
        }
    };

    // one iteration: exchange the finite bits of the 8 pixels, write the mask, accumulate
    auto process = [&](long long base, const float (&xv)[8], const float (&yv)[8], unsigned int vb, int buf) {
        const long long i0 = base + lane * 4, i1 = i0 + 128;
        const bool full = base + FIT_PX <= n;
        unsigned int bits = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            bool okj = finite_f32(xv[j]);
            if (k == P.gate_k) okj = okj && (xv[j] > P.gate_gt);
            bits |= okj ? (1u << j) : 0u;
        }
        if (k == 0) bits &= vb;
        fin[buf][k][lane] = bits;
        __syncthreads();
        unsigned int m = 0xffu;
        for (int kk = 0; kk < P.K; ++kk) m &= fin[buf][kk][lane];
        if (k == 0 && mg) {
            if (P.vecm && full) {
                *reinterpret_cast<uchar4*>(mg + i0) = make_uchar4(m & 1u, (m >> 1) & 1u, (m >> 2) & 1u, (m >> 3) & 1u);
                *reinterpret_cast<uchar4*>(mg + i1) =
                    make_uchar4((m >> 4) & 1u, (m >> 5) & 1u, (m >> 6) & 1u, (m >> 7) & 1u);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const long long i = (j < 4 ? i0 : i1) + (j & 3);
                    if (i < n) mg[i] = (m >> j) & 1u;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) accumulate<DEG>(acc, xv[j], yv[j], ((m >> j) & 1u) && finite_f32(yv[j]));
    };

    // Two register buffers, refilled right after they are consumed: one or two iterations' loads are always in
    // flight while the block does the other one's math.  The loop bounds are uniform across the block (the
    // __syncthreads inside process() is reached by every warp).
    const long long step = (long long)gridDim.x * FIT_PX;
    long long b0 = (long long)blockIdx.x * FIT_PX, b1 = b0 + step;
    float xa[8], ya[8], xb[8], yb[8];
    unsigned int va = 0u, vb2 = 0u;
    if (b0 < n) fetch(b0, xa, ya, va);
    if (b1 < n) fetch(b1, xb, yb, vb2);
    while (b0 < n) {
        process(b0, xa, ya, va, 0);
        if (b0 + 2 * step < n) fetch(b0 + 2 * step, xa, ya, va);
        if (b1 >= n) break;
        process(b1, xb, yb, vb2, 1);
        if (b1 + 2 * step < n) fetch(b1 + 2 * step, xb, yb, vb2);
        b0 += 2 * step;
        b1 += 2 * step;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double s = warp_sum(acc[j]);
        if (lane == 0) P.partial[(((long long)k * P.G + g) * gridDim.x + blockIdx.x) * M + j] = s;
    }
}

// Grid of the fit: every block resident at once (no tail wave); `per_sm` is the occupancy of the kernel.
int fit_blocks(long long n, int G, int per_sm) {
    long long per = (n + FIT_PX - 1) / FIT_PX;
    if (per < 1) per = 1;
    long long cap = (long long)device_sm_count() * (per_sm > 0 ? per_sm : 1) / (G > 0 ? G : 1);
    if (cap < 1) cap = 1;
    return (int)(per < cap ? per : cap);
}

template <int DEG>
int fit_occupancy(int K) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fit_moments_kernel<DEG>, 32 * K, 0) != cudaSuccess || nb < 1)
        nb = 1;
    return nb;
}

// ------------------------------------------------------------------------------------------------
// Fused solve + apply: grid = (blocks per series, S series).  Warp 0 of every block solves its
// series' scaled normal equations (a 3 x 3 system for degree 2: cheaper than a launch), the block
// then maps its share of the samples (Horner in fp64, mask, clip: poly_regression.py:65-84).
struct ApplyParams {
    const float* x;
    long long xks, xgs;  // series s = k * G + g  ->  x + k * xks + g * xgs
    float* out;
    long long oks, ogs;
    const uint8_t* mask;  // nullable, [G, n]
    const double* moments;  // [S, 3*deg+2]
    double* coeffs;         // [S, deg+1] (written by block x == 0 of every series)
    long long n, min_count;
    int G, vec;
    float lo, hi;
};

template <int DEG>
__global__ void __launch_bounds__(256) solve_apply_kernel(const ApplyParams P) {
    __shared__ double A[MAXDEG + 1][MAXDEG + 2];
    __shared__ double scale[MAXDEG + 1];
    __shared__ double cs[MAXDEG + 1];
    const int s = blockIdx.y;
    const int k = s / P.G, g = s - k * P.G;
    if (threadIdx.x < 32) {
        solve_series(P.moments + (long long)s * (3 * DEG + 2), DEG, P.min_count, A, scale, cs, threadIdx.x);
        __syncwarp();
        if (blockIdx.x == 0 && threadIdx.x <= DEG) P.coeffs[(long long)s * (DEG + 1) + threadIdx.x] = cs[threadIdx.x];
    }
    __syncthreads();
    double c[DEG + 1];
#pragma unroll
    for (int j = 0; j <= DEG; ++j) c[j] = cs[j];

    const float* xs = P.x + (long long)k * P.xks + (long long)g * P.xgs;
    float* os = P.out + (long long)k * P.oks + (long long)g * P.ogs;
    const uint8_t* mg = P.mask ? P.mask + (long long)g * P.n : nullptr;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    long long done = 0;
    if (P.vec) {
        const long long n4 = P.n >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(xs);
        float4* o4 = reinterpret_cast<float4*>(os);
        const uchar4* m4 = reinterpret_cast<const uchar4*>(mg);
        long long i = tid;
        for (; i + nthreads < n4; i += 2 * nthreads) {  // two 16-byte loads in flight per thread
            const float4 xa = __ldg(x4 + i), xb = __ldg(x4 + i + nthreads);
            uchar4 ma = make_uchar4(1, 1, 1, 1), mb = ma;
            if (mg) {
                ma = __ldg(m4 + i);
                mb = __ldg(m4 + i + nthreads);
            }
            float4 ra, rb;
            ra.x = horner_clip<DEG>(xa.x, c, ma.x != 0, P.lo, P.hi);
            ra.y = horner_clip<DEG>(xa.y, c, ma.y != 0, P.lo, P.hi);
            ra.z = horner_clip<DEG>(xa.z, c, ma.z != 0, P.lo, P.hi);
            ra.w = horner_clip<DEG>(xa.w, c, ma.w != 0, P.lo, P.hi);
            rb.x = horner_clip<DEG>(xb.x, c, mb.x != 0, P.lo, P.hi);
            rb.y = horner_clip<DEG>(xb.y, c, mb.y != 0, P.lo, P.hi);
            rb.z = horner_clip<DEG>(xb.z, c, mb.z != 0, P.lo, P.hi);
            rb.w = horner_clip<DEG>(xb.w, c, mb.w != 0, P.lo, P.hi);
            __stcs(o4 + i, ra);
            __stcs(o4 + i + nthreads, rb);
        }
        for (; i < n4; i += nthreads) {
            const float4 xa = __ldg(x4 + i);
            uchar4 ma = make_uchar4(1, 1, 1, 1);
            if (mg) ma = __ldg(m4 + i);
            float4 ra;
            ra.x = horner_clip<DEG>(xa.x, c, ma.x != 0, P.lo, P.hi);
            ra.y = horner_clip<DEG>(xa.y, c, ma.y != 0, P.lo, P.hi);
            ra.z = horner_clip<DEG>(xa.z, c, ma.z != 0, P.lo, P.hi);
            ra.w = horner_clip<DEG>(xa.w, c, ma.w != 0, P.lo, P.hi);
            __stcs(o4 + i, ra);
        }
        done = n4 << 2;
    }
    for (long long i = done + tid; i < P.n; i += nthreads)
        os[i] = horner_clip<DEG>(__ldg(xs + i), c, mg == nullptr || mg[i] != 0, P.lo, P.hi);
}

template <int DEG>
void launch_moments(const float* x, long long xks, long long xns, const float* y, long long yks, long long yns,
                    const uint8_t* mask, long long mdiv, long long mmod, long long n, int K, int nblk,
                    double* partial, cudaStream_t stream) {
    dim3 grid((unsigned int)nblk, (unsigned int)K);
    poly_moments_kernel<DEG><<<grid, MOM_THREADS, 0, stream>>>(x, xks, xns, y, yks, yns, mask, mdiv, mmod, n,
                                                               partial);
}

template <int DEG>
void launch_apply(const float* x, long long xks, long long xns, const double* coeffs, const uint8_t* mask,
                  long long mdiv, long long mmod, long long n, int K, float lo, float hi, float* out, long long oks,
                  long long ons, int nblk, cudaStream_t stream) {
    dim3 grid((unsigned int)nblk, (unsigned int)K);
    poly_apply_kernel<DEG><<<grid, 256, 0, stream>>>(x, xks, xns, coeffs, mask, mdiv, mmod, n, lo, hi, out, oks,
                                                     ons);
}

}  // namespace

size_t poly_moments_workspace(long long n, int K, int deg) {
    if (n < 0 || K < 1 || deg < 1 || deg > MAXDEG) return 0;
    return (size_t)moments_blocks(n, K) * (size_t)K * (size_t)n_moments(deg) * sizeof(double);
}

#define HSR_DEG_SWITCH(deg, CALL)  \
    switch (deg) {                 \
        case 1: CALL(1); break;    \
        case 2: CALL(2); break;    \
        case 3: CALL(3); break;    \
        case 4: CALL(4); break;    \
        case 5: CALL(5); break;    \
        case 6: CALL(6); break;    \
        case 7: CALL(7); break;    \
        default: CALL(8); break;   \
    }

int poly_moments_impl(const float* x, long long xks, long long xns, const float* y, long long yks, long long yns,
                      const uint8_t* mask, long long mask_k_div, long long mask_k_mod, long long n, int K, int deg,
                      double* partial, double* moments, cudaStream_t stream) {
    HSR_REQUIRE(x && y && partial && moments, HSR_EINVAL, "null x / y / partial / moments pointer");
    HSR_REQUIRE(n >= 0 && K >= 1 && K <= 65535, HSR_EINVAL, "bad n = %lld or K = %d", n, K);
    HSR_REQUIRE(deg >= 1 && deg <= MAXDEG, HSR_ERANGE, "deg = %d outside [1, %d]", deg, MAXDEG);
    HSR_REQUIRE(mask == nullptr || (mask_k_div >= 1 && mask_k_mod >= 1), HSR_EINVAL,
                "mask_k_div / mask_k_mod must be >= 1");
    HSR_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 3) == 0, HSR_EALIGN,
                "x / y not 4-byte aligned");
    const int nblk = moments_blocks(n, K);
#define CALL(D) \
    launch_moments<D>(x, xks, xns, y, yks, yns, mask, mask_k_div, mask_k_mod, n, K, nblk, partial, stream)
    HSR_DEG_SWITCH(deg, CALL)
#undef CALL
    HSR_CUDA(cudaGetLastError());
    poly_moments_finalize_kernel<<<(unsigned int)K, 32, 0, stream>>>(partial, nblk, n_moments(deg), moments);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int poly_solve_impl(const double* moments, int K, int deg, long long min_count, double* coeffs, cudaStream_t stream) {
    HSR_REQUIRE(moments && coeffs, HSR_EINVAL, "null moments / coeffs pointer");
    HSR_REQUIRE(K >= 1, HSR_EINVAL, "K = %d", K);
    HSR_REQUIRE(deg >= 1 && deg <= MAXDEG, HSR_ERANGE, "deg = %d outside [1, %d]", deg, MAXDEG);
    poly_solve_kernel<<<(unsigned int)((K + 3) / 4), 128, 0, stream>>>(moments, K, deg, min_count, coeffs);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int poly_apply_impl(const float* x, long long xks, long long xns, const double* coeffs, const uint8_t* mask,
                    long long mask_k_div, long long mask_k_mod, long long n, int K, int deg, float lo, float hi,
                    float* out, long long oks, long long ons, cudaStream_t stream) {
    HSR_REQUIRE(x && coeffs && out, HSR_EINVAL, "null x / coeffs / out pointer");
    HSR_REQUIRE(n >= 0 && K >= 1 && K <= 65535, HSR_EINVAL, "bad n = %lld or K = %d", n, K);
    HSR_REQUIRE(deg >= 1 && deg <= MAXDEG, HSR_ERANGE, "deg = %d outside [1, %d]", deg, MAXDEG);
    HSR_REQUIRE(mask == nullptr || (mask_k_div >= 1 && mask_k_mod >= 1), HSR_EINVAL,
                "mask_k_div / mask_k_mod must be >= 1");
    if (n == 0) return HSR_OK;
    long long nblk = (n + 256 * 8 - 1) / (256 * 8);
    long long cap = MOM_MAX_BLOCKS / K;
    if (cap < 1) cap = 1;
    if (nblk > cap) nblk = cap;
#define CALL(D) \
    launch_apply<D>(x, xks, xns, coeffs, mask, mask_k_div, mask_k_mod, n, K, lo, hi, out, oks, ons, (int)nblk, \
                    stream)
    HSR_DEG_SWITCH(deg, CALL)
#undef CALL
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int fit_mask_impl(const float* x, long long xks, long long n, int K, const uint8_t* valid, int gate_k, float gate_gt,
                  uint8_t* mask, cudaStream_t stream) {
    HSR_REQUIRE(x && mask, HSR_EINVAL, "null x / mask pointer");
    HSR_REQUIRE(n >= 0 && K >= 1, HSR_EINVAL, "bad n = %lld or K = %d", n, K);
    HSR_REQUIRE(gate_k < K, HSR_EINVAL, "gate_k = %d >= K = %d", gate_k, K);
    if (n == 0) return HSR_OK;
    long long nblk = (n + 255) / 256;
    if (nblk > MOM_MAX_BLOCKS) nblk = MOM_MAX_BLOCKS;
    fit_mask_kernel<<<(unsigned int)nblk, 256, 0, stream>>>(x, xks, n, K, valid, gate_k, gate_gt, mask);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int fit_occupancy_of(int deg, int K) {
    int nb = 1;
#define CALL(D) nb = fit_occupancy<D>(K)
    HSR_DEG_SWITCH(deg, CALL)
#undef CALL
    return nb;
}

int fit_moments_impl(const float* x, long long xks, long long xgs, const float* y, long long yks, long long ygs,
                     const uint8_t* valid, long long n, int K, int G, int deg, int gate_k, float gate_gt,
                     uint8_t* mask, double* partial, double* moments, cudaStream_t stream) {
    HSR_REQUIRE(x && y && partial && moments, HSR_EINVAL, "null x / y / partial / moments pointer");
    HSR_REQUIRE(n >= 0 && K >= 1 && K <= HSR_MAX_SRF_BANDS, HSR_ERANGE, "bad n = %lld or K = %d (K <= %d)", n, K,
                HSR_MAX_SRF_BANDS);
    HSR_REQUIRE(G >= 1 && G <= 65535, HSR_ERANGE, "G = %d outside [1, 65535]", G);
    HSR_REQUIRE(deg >= 1 && deg <= MAXDEG, HSR_ERANGE, "deg = %d outside [1, %d]", deg, MAXDEG);
    HSR_REQUIRE(gate_k < K, HSR_EINVAL, "gate_k = %d >= K = %d", gate_k, K);
    HSR_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 3) == 0, HSR_EALIGN,
                "x / y not 4-byte aligned");
    FitParams P{};
    P.x = x, P.xks = xks, P.xgs = xgs, P.y = y, P.yks = yks, P.ygs = ygs;
    P.valid = valid, P.mask = mask, P.n = n, P.K = K, P.G = G, P.gate_k = gate_k, P.gate_gt = gate_gt;
    P.partial = partial;
    // strides only matter along dimensions that are actually stepped
    const long long xs_or = (K > 1 ? xks : 0) | (G > 1 ? xgs : 0), ys_or = (K > 1 ? yks : 0) | (G > 1 ? ygs : 0);
    P.vecx = ((reinterpret_cast<uintptr_t>(x) | (uintptr_t)(xs_or * 4)) & 15) == 0 ? 1 : 0;
    P.vecy = ((reinterpret_cast<uintptr_t>(y) | (uintptr_t)(ys_or * 4)) & 15) == 0 ? 1 : 0;
    P.vecm = ((reinterpret_cast<uintptr_t>(valid) | reinterpret_cast<uintptr_t>(mask) | (uintptr_t)(G > 1 ? n : 0)) & 3) == 0
                 ? 1 : 0;
    const int nblk = fit_blocks(n, G, fit_occupancy_of(deg, K));
    dim3 grid((unsigned int)nblk, (unsigned int)G);
#define CALL(D) fit_moments_kernel<D><<<grid, 32 * K, 0, stream>>>(P)
    HSR_DEG_SWITCH(deg, CALL)
#undef CALL
    HSR_CUDA(cudaGetLastError());
    poly_moments_finalize_kernel<<<(unsigned int)(K * G), 32, 0, stream>>>(partial, nblk, n_moments(deg), moments);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

size_t fit_moments_workspace(long long n, int K, int G, int deg) {
    if (n < 0 || K < 1 || K > HSR_MAX_SRF_BANDS || G < 1 || deg < 1 || deg > MAXDEG) return 0;
    return (size_t)fit_blocks(n, G, fit_occupancy_of(deg, K)) * (size_t)K * (size_t)G * (size_t)n_moments(deg) *
           sizeof(double);
}

int poly_solve_apply_impl(const float* x, long long xks, long long xgs, const double* moments, const uint8_t* mask,
                          long long n, int K, int G, int deg, long long min_count, float lo, float hi, double* coeffs,
                          float* out, long long oks, long long ogs, cudaStream_t stream) {
    HSR_REQUIRE(x && moments && coeffs && out, HSR_EINVAL, "null x / moments / coeffs / out pointer");
    HSR_REQUIRE(n >= 0 && K >= 1 && G >= 1 && (long long)K * G <= 65535, HSR_ERANGE,
                "bad n = %lld, K = %d or G = %d (K * G <= 65535)", n, K, G);
    HSR_REQUIRE(deg >= 1 && deg <= MAXDEG, HSR_ERANGE, "deg = %d outside [1, %d]", deg, MAXDEG);
    ApplyParams P{};
    P.x = x, P.xks = xks, P.xgs = xgs, P.out = out, P.oks = oks, P.ogs = ogs, P.mask = mask, P.moments = moments;
    P.coeffs = coeffs, P.n = n, P.min_count = min_count, P.G = G, P.lo = lo, P.hi = hi;
    const long long s_or = (K > 1 ? (xks | oks) : 0) | (G > 1 ? (xgs | ogs) : 0);  // strides that are stepped
    const uintptr_t a16 = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | (uintptr_t)(s_or * 4);
    const uintptr_t a4 = reinterpret_cast<uintptr_t>(mask) | (uintptr_t)(G > 1 ? n : 0);
    P.vec = ((a16 & 15) == 0 && (a4 & 3) == 0) ? 1 : 0;
    const long long S = (long long)K * G;
    long long nblk = (n + 256 * 8 - 1) / (256 * 8);
    long long cap = (long long)device_sm_count() * 8 / S;
    if (cap < 1) cap = 1;
    if (nblk > cap) nblk = cap;
    if (nblk < 1) nblk = 1;
    dim3 grid((unsigned int)nblk, (unsigned int)S);
#define CALL(D) solve_apply_kernel<D><<<grid, 256, 0, stream>>>(P)
    HSR_DEG_SWITCH(deg, CALL)
#undef CALL
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

}  // namespace hsr
