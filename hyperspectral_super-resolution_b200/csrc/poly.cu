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

template <int DEG>
__device__ __forceinline__ void accumulate(double (&acc)[3 * DEG + 2], float xf, float yf, bool use) {
    if (!use) return;
    const double x = (double)xf, y = (double)yf;
    double pw = 1.0;
#pragma unroll
    for (int j = 0; j <= 2 * DEG; ++j) {
        acc[j] += pw;
        if (j <= DEG) acc[2 * DEG + 1 + j] = fma(pw, y, acc[2 * DEG + 1 + j]);
        pw *= x;
    }
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

// One warp per series.  Normal equations G c = r with G_ij = S_{i+j}, r_i = T_i, scaled by
// s_j = sqrt(S_{2j}) (the column norms np.polyfit divides its Vandermonde by), solved by
// Gaussian elimination with partial pivoting; lane j owns column j of the augmented matrix.
__global__ void __launch_bounds__(128) poly_solve_kernel(const double* __restrict__ moments, int K, int deg,
                                                         long long min_count, double* __restrict__ coeffs) {
    __shared__ double A[4][MAXDEG + 1][MAXDEG + 2];
    __shared__ double scale[4][MAXDEG + 1];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + w;
    if (k >= K) return;
    const int N = deg + 1, M = 3 * deg + 2;
    const double* mom = moments + (long long)k * M;
    double* out = coeffs + (long long)k * N;

    const double count = mom[0];
    if (!(count >= (double)min_count)) {  // identity: poly_regression.py:38-41
        if (lane < N) out[lane] = (lane == N - 2) ? 1.0 : 0.0;
        return;
    }
    if (lane < N) scale[w][lane] = sqrt(mom[2 * lane]);
    __syncwarp();
    if (lane <= N) {
        for (int i = 0; i < N; ++i) {
            double v;
            if (lane < N)
                v = mom[i + lane] / (scale[w][i] * scale[w][lane]);
            else
                v = mom[2 * deg + 1 + i] / scale[w][i];
            A[w][i][lane] = v;
        }
    }
    __syncwarp();
    for (int c = 0; c < N; ++c) {
        // pivot search (every lane computes the same answer)
        int piv = c;
        double best = fabs(A[w][c][c]);
        for (int r = c + 1; r < N; ++r) {
            const double v = fabs(A[w][r][c]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        __syncwarp();
        if (lane <= N && piv != c) {
            const double t = A[w][c][lane];
            A[w][c][lane] = A[w][piv][lane];
            A[w][piv][lane] = t;
        }
        __syncwarp();
        const double inv = 1.0 / A[w][c][c];
        __syncwarp();
        if (lane <= N) A[w][c][lane] *= inv;
        __syncwarp();
        double f[MAXDEG + 1];  // column c, read by every lane before lane c rewrites it
#pragma unroll
        for (int r = 0; r <= MAXDEG; ++r) f[r] = (r < N) ? A[w][r][c] : 0.0;
        __syncwarp();
        if (lane <= N) {
            const double pc = A[w][c][lane];
#pragma unroll
            for (int r = 0; r <= MAXDEG; ++r)
                if (r < N && r != c) A[w][r][lane] = fma(-f[r], pc, A[w][r][lane]);
        }
        __syncwarp();
    }
    // un-scale; np.polyfit order is highest power first
    if (lane < N) out[deg - lane] = A[w][lane][N] / scale[w][lane];
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

}  // namespace hsr
