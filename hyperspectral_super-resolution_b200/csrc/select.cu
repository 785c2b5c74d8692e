// select.cu — shared percentile stretch: exact masked percentiles of fp32 planes by a
// three-pass MSB-first radix select, and the stretch itself.
//
// Reference: s2_emit/color.py:25-34 (apply_shared_percentile_stretch), called between the SRF synthesis and
// the polynomial fit at s2_emit/poly_regression.py:126-127:
//     vals = img[..., c][mask];  lo, hi = np.percentile(vals, [pmin, pmax])
//     out[..., c] = np.clip((img[..., c] - lo) / (hi - lo + 1e-12), 0, 1)          (float64 -> float32)
// np.percentile (method "linear") on n masked samples: virtual index v = (n - 1) * q/100, neighbours
// a = sorted[floor(v)], b = sorted[floor(v) + 1] (both the last element when v >= n - 1), gamma = v - floor(v),
// result = a + (b - a) * gamma, or b - (b - a) * (1 - gamma) when gamma >= 0.5, with b - a rounded in fp32 and
// the rest in fp64; NaN anywhere among the samples makes every percentile NaN.  All of that is reproduced
// bit for bit: the order statistics are found exactly (no sampling, no sort), the interpolation follows
// numpy's operation order with explicitly un-fused fp64 arithmetic.
//
// Radix select: keys are the usual order-preserving u32 image of fp32 (-0 folded onto +0, NaNs counted apart).
// Each series has T = 2 * Q targets (the two neighbours of every percentile).  Pass p histograms the next
// 11 / 11 / 10 key bits of the samples that match a target's prefix (shared-memory histograms, warp-aggregated
// in the first pass where natural images put whole warps into one bin; merged into a small global histogram
// by atomics), a one-block scan then extends every target's prefix and rebases its rank.  Three streaming
// reads of the planes in total; the second and third mostly hit L2 for RGB-sized inputs.
#include "hsr_common.cuh"

namespace hsr {

namespace {

constexpr int SEL_QMAX = HSR_MAX_PERCENTILES;
constexpr int SEL_TMAX = 2 * SEL_QMAX;
constexpr int SEL_BINS = 2048;
constexpr int SEL_THREADS = 256;

struct SelSeries {
    unsigned int prefix[SEL_TMAX];  // key bits fixed so far for target t (right-aligned)
    long long rank[SEL_TMAX];       // rank of target t among the samples that share its prefix
    int slot[SEL_TMAX];             // histogram slot target t reads; targets with equal prefixes share one
    unsigned int slot_prefix[SEL_TMAX];
    int nslot;
    int dead;                // no samples, or a NaN among them: every percentile is NaN
    long long count, nan;    // masked samples (NaNs included) and NaNs among them (atomically accumulated)
    double gamma[SEL_QMAX];
};

struct SelParams {
    const float* x;
    long long xks, xgs;
    const uint8_t* mask;  // nullable [G, n]
    long long n;
    int G, vec;
    SelSeries* st;         // [S]
    unsigned int* hist;    // [S][SEL_TMAX][SEL_BINS]
};

__device__ __forceinline__ bool key_of(float v, unsigned int& key) {
    unsigned int u = __float_as_uint(v);
    if ((u & 0x7fffffffu) > 0x7f800000u) return false;  // NaN
    if ((u << 1) == 0u) u = 0u;                         // -0 -> +0 (equal under numpy's ordering)
    key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return true;
}

__device__ __forceinline__ float value_of(unsigned int key) {
    const unsigned int u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    return __uint_as_float(u);
}

// PASS 0: bits [31:21] of every sample; PASS 1: bits [20:10] under an 11-bit prefix; PASS 2: bits [9:0]
// under a 22-bit prefix.
template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) select_hist_kernel(const SelParams P) {
    constexpr int SHIFT = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
    constexpr int PSHIFT = PASS == 1 ? 21 : 10;  // key >> PSHIFT is the prefix (passes 1, 2)
    constexpr unsigned int DMASK = PASS == 2 ? 1023u : 2047u;
    constexpr int NS = PASS == 0 ? 1 : SEL_TMAX;
    __shared__ unsigned int h[NS][SEL_BINS];
    const long long s = blockIdx.y;
    const long long k = s / P.G, g = s - k * P.G;
    const float* xs = P.x + k * P.xks + g * P.xgs;
    const uint8_t* mg = P.mask ? P.mask + g * P.n : nullptr;
    SelSeries* st = P.st + s;
    const int lane = threadIdx.x & 31;

    int nslot = 1;
    unsigned int sp[SEL_TMAX];  // slot prefixes; unused slots hold a value no prefix can take
    if (PASS > 0) {
        if (st->dead) return;
        nslot = st->nslot;
#pragma unroll
        for (int t = 0; t < SEL_TMAX; ++t) sp[t] = t < nslot ? st->slot_prefix[t] : 0xffffffffu;
    }
    for (int i = threadIdx.x; i < nslot * SEL_BINS; i += SEL_THREADS) (&h[0][0])[i] = 0u;
    __syncthreads();

    unsigned int cnt = 0, nan = 0;  // per thread: at most n / threads samples, far below 2^32
    auto take = [&](float v, bool use) {
        unsigned int key = 0u;
        const bool finite_key = use && key_of(v, key);
        if (PASS == 0) {
            cnt += use ? 1u : 0u;
            nan += (use && !finite_key) ? 1u : 0u;
            // neighbouring pixels of natural images mostly share their top 11 key bits: when the whole warp
            // agrees, one lane adds 32; otherwise plain shared-memory atomics (conflicts replay, still cheap)
            const unsigned int bin = finite_key ? (key >> SHIFT) : 0xffffffffu;
            const unsigned int b0 = __shfl_sync(0xffffffffu, bin, 0);
            if (__all_sync(0xffffffffu, bin == b0)) {
                if (lane == 0 && b0 != 0xffffffffu) atomicAdd(&h[0][b0], 32u);
            } else if (finite_key) {
                atomicAdd(&h[0][bin], 1u);
            }
        } else {
            const unsigned int pre = finite_key ? (key >> PSHIFT) : 0xfffffffeu;
            bool hit = false;
#pragma unroll
            for (int t = 0; t < SEL_TMAX; ++t) hit |= (pre == sp[t]);
            if (hit) {  // rare: a few per cent of the samples after the first pass, a handful after the second
#pragma unroll
                for (int t = 0; t < SEL_TMAX; ++t)
                    if (pre == sp[t]) atomicAdd(&h[t][(key >> SHIFT) & DMASK], 1u);
            }
        }
    };

    const long long n = P.n;
    // warp-uniform trip counts (the first pass uses warp votes); 32-bit indices: n < 2^33 is checked by the host
    const unsigned int tid = blockIdx.x * SEL_THREADS + threadIdx.x;
    const unsigned int nthreads = gridDim.x * SEL_THREADS;
    const unsigned int n4 = P.vec ? (unsigned int)(n >> 2) : 0u;
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    const uchar4* m4 = reinterpret_cast<const uchar4*>(mg);
    constexpr int UNR = 4;  // 16-byte loads in flight per thread (the passes are latency-bound otherwise)
    for (unsigned int base = tid - lane; base < n4; base += UNR * nthreads) {
        float4 v[UNR];
        uchar4 m[UNR];
        bool in[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const unsigned int i = base + u * nthreads + lane;
            in[u] = (base + u * nthreads < n4) && i < n4;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            m[u] = make_uchar4(1, 1, 1, 1);
            if (in[u]) {
                v[u] = __ldg(x4 + i);
                if (mg) m[u] = __ldg(m4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (base + u * nthreads < n4) {  // warp-uniform
                take(v[u].x, in[u] && m[u].x);
                take(v[u].y, in[u] && m[u].y);
                take(v[u].z, in[u] && m[u].z);
                take(v[u].w, in[u] && m[u].w);
            }
        }
    }
    for (long long base = ((long long)n4 << 2) + tid - lane; base < n; base += nthreads) {
        const long long i = base + lane;
        const bool in = i < n;
        const float v = in ? __ldg(xs + i) : 0.f;
        take(v, in && (mg == nullptr || mg[i] != 0));
    }
    __syncthreads();

    unsigned int* gh = P.hist + s * (long long)(SEL_TMAX * SEL_BINS);
    for (int i = threadIdx.x; i < nslot * SEL_BINS; i += SEL_THREADS) {
        const unsigned int c = (&h[0][0])[i];
        if (c) atomicAdd(gh + i, c);
    }
    if (PASS == 0) {
        unsigned long long c64 = cnt, n64 = nan;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c64 += __shfl_xor_sync(0xffffffffu, c64, o);
            n64 += __shfl_xor_sync(0xffffffffu, n64, o);
        }
        if (lane == 0 && c64) atomicAdd(reinterpret_cast<unsigned long long*>(&st->count), c64);
        if (lane == 0 && n64) atomicAdd(reinterpret_cast<unsigned long long*>(&st->nan), n64);
    }
}

// One block per series: locate, for every target, the bin its rank falls into; extend the prefix, rebase the
// rank, regroup the targets into histogram slots, and clear the histogram for the next pass.  After the
// last pass the prefixes are complete keys: interpolate and write the percentiles.
template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) select_scan_kernel(SelSeries* __restrict__ sts,
                                                                  unsigned int* __restrict__ hist,
                                                                  const double* __restrict__ q, int Q,
                                                                  double* __restrict__ out) {
    constexpr int BITS = PASS == 2 ? 10 : 11;
    constexpr int PER = SEL_BINS / SEL_THREADS;  // bins per thread
    constexpr int WORDS = sizeof(SelSeries) / 4;
    __shared__ SelSeries st;  // the series' state lives in shared memory while this block works on it
    __shared__ unsigned long long wsum[SEL_THREADS / 32];
    __shared__ unsigned int found_bin[SEL_TMAX];
    __shared__ long long found_rank[SEL_TMAX];
    const long long s = blockIdx.x;
    unsigned int* gh = hist + s * (long long)(SEL_TMAX * SEL_BINS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = 2 * Q;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = tid; i < WORDS; i += SEL_THREADS)
        reinterpret_cast<unsigned int*>(&st)[i] = reinterpret_cast<const unsigned int*>(sts + s)[i];
    __syncthreads();

    if (PASS == 0) {
        if (tid == 0) {
            const long long n = st.count;
            st.dead = (n == 0 || st.nan != 0) ? 1 : 0;
            for (int qi = 0; qi < Q; ++qi) {
                // np.percentile, method "linear": virtual index (n - 1) * q, neighbours floor / floor + 1,
                // both the last element when the index is >= n - 1 (numpy then forms gamma from index -1)
                const double vi = __dmul_rn((double)(n - 1), q[qi]);
                double prev = floor(vi);
                long long r0 = (long long)prev, r1 = r0 + 1;
                if (vi >= (double)(n - 1)) {
                    r0 = r1 = n - 1;
                    prev = -1.0;
                }
                if (vi < 0.0) {
                    r0 = r1 = 0;
                    prev = 0.0;
                }
                st.rank[2 * qi] = r0;
                st.rank[2 * qi + 1] = r1;
                st.gamma[qi] = vi - prev;
            }
            for (int t = 0; t < SEL_TMAX; ++t) {
                st.prefix[t] = 0u;
                st.slot[t] = 0;  // one shared histogram in the first pass
            }
            st.nslot = 1;
        }
        __syncthreads();
    }
    const int used = st.nslot * SEL_BINS;  // histogram words the pass before this scan has filled
    if (st.dead) {
        if (PASS == 2 && tid < Q) out[s * Q + tid] = qnan;
        if (PASS == 0 && tid == 0) sts[s].dead = 1;
        for (int i = tid; i < used; i += SEL_THREADS) gh[i] = 0u;
        return;
    }

    for (int t = 0; t < T; ++t) {
        const unsigned int* hrow = gh + (long long)st.slot[t] * SEL_BINS;
        const long long rank = st.rank[t];
        unsigned int c[PER];
        unsigned long long local = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            c[j] = hrow[tid * PER + j];
            local += c[j];
        }
        unsigned long long incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long before = 0;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        unsigned long long excl = before + incl - local;
        if ((unsigned long long)rank >= excl && (unsigned long long)rank < excl + local) {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if ((unsigned long long)rank >= excl && (unsigned long long)rank < excl + c[j]) {
                    found_bin[t] = (unsigned int)(tid * PER + j);
                    found_rank[t] = rank - (long long)excl;
                }
                excl += c[j];
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        int nslot = 0;
        for (int t = 0; t < T; ++t) {
            const unsigned int pre = (st.prefix[t] << BITS) | found_bin[t];
            st.prefix[t] = pre;
            st.rank[t] = found_rank[t];
            int sl = -1;
            for (int u = 0; u < nslot; ++u)
                if (st.slot_prefix[u] == pre) sl = u;
            if (sl < 0) {
                sl = nslot++;
                st.slot_prefix[sl] = pre;
            }
            st.slot[t] = sl;
        }
        st.nslot = nslot;
        if (PASS == 2) {
            for (int qi = 0; qi < Q; ++qi) {
                const float a = value_of(st.prefix[2 * qi]), b = value_of(st.prefix[2 * qi + 1]);
                const double t = st.gamma[qi];
                const float diff = __fsub_rn(b, a);  // numpy subtracts the two float32 neighbours first
                double r = __dadd_rn((double)a, __dmul_rn((double)diff, t));
                if (t >= 0.5) r = __dsub_rn((double)b, __dmul_rn((double)diff, __dsub_rn(1.0, t)));
                out[s * Q + qi] = r;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < WORDS; i += SEL_THREADS)
        reinterpret_cast<unsigned int*>(sts + s)[i] = reinterpret_cast<const unsigned int*>(&st)[i];
    for (int i = tid; i < used; i += SEL_THREADS) gh[i] = 0u;
}

__global__ void select_init_kernel(SelSeries* sts, long long S) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s < S) {
        sts[s].count = 0;
        sts[s].nan = 0;
        sts[s].dead = 0;
        sts[s].nslot = 1;
    }
}

// out = (f32) clip((f64(x) - lo) / (hi - lo + 1e-12), 0, 1), series s = k * G + g   (color.py:33)
struct StretchParams {
    const float* x;
    long long xks, xgs;
    float* out;
    long long oks, ogs;
    const double* lohi;  // [S][2]
    long long n;
    int G, vec;
};

__global__ void __launch_bounds__(256) stretch_kernel(const StretchParams P) {
    const long long s = blockIdx.y;
    const long long k = s / P.G, g = s - k * P.G;
    const float* xs = P.x + k * P.xks + g * P.xgs;
    float* os = P.out + k * P.oks + g * P.ogs;
    const double lo = P.lohi[2 * s], den = P.lohi[2 * s + 1] - lo + 1e-12;
    auto f = [&](float v) {
        double r = ((double)v - lo) / den;
        r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);  // NaN survives, as np.clip
        return (float)r;
    };
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long n4 = P.vec ? (P.n >> 2) : 0;
    for (long long i = tid; i < n4; i += nthreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xs) + i);
        __stcs(reinterpret_cast<float4*>(os) + i, make_float4(f(v.x), f(v.y), f(v.z), f(v.w)));
    }
    for (long long i = (n4 << 2) + tid; i < P.n; i += nthreads) os[i] = f(__ldg(xs + i));
}

// out = clip((f64(x) - lo) / (hi - lo + 1e-12), 0, 1) as float64, NaN outside the mask: robust_norm / robust_norm_rgb
// (color.py:6-23), whose results are float64 arrays
struct Stretch64Params {
    const float* x;
    long long xks, xgs;
    double* out;
    long long oks, ogs;
    const double* lohi;
    const uint8_t* mask;    // nullable: [G][n]
    long long n;
    int G;
};

__global__ void __launch_bounds__(256) stretch64_kernel(const Stretch64Params P) {
    const long long s = blockIdx.y;
    const long long k = s / P.G, g = s - k * P.G;
    const float* xs = P.x + k * P.xks + g * P.xgs;
    double* os = P.out + k * P.oks + g * P.ogs;
    const uint8_t* ms = P.mask ? P.mask + g * P.n : nullptr;
    const double lo = P.lohi[2 * s], den = P.lohi[2 * s + 1] - lo + 1e-12;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P.n; i += (long long)gridDim.x * blockDim.x) {
        double r = ((double)__ldg(xs + i) - lo) / den;
        r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);  // NaN survives, as np.clip
        if (ms && !ms[i]) r = __longlong_as_double(0x7ff8000000000000LL);
        os[i] = r;
    }
}

// out[i] = !isnan(x[i]) (& base[i]): the sample set of np.nanpercentile as a mask for the select kernels
__global__ void __launch_bounds__(256) notnan_mask_kernel(const float* __restrict__ x, const uint8_t* __restrict__ base,
                                                          long long n, uint8_t* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(x + i);
        out[i] = (v == v && (base == nullptr || base[i])) ? 1 : 0;
    }
}

// ---- histogram matching (color.py:36-61): flags of the last element of every run of equal values in a sorted array
__global__ void __launch_bounds__(256) run_end_kernel(const float* __restrict__ v, long long n, uint8_t* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (i == n - 1 || !(v[i] == v[i + 1])) ? 1 : 0;
}

// number of elements <= x in the ascending array v[0, n)
__device__ __forceinline__ long long count_le(const float* __restrict__ v, long long n, float x) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (v[mid] <= x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// One thread per pixel.  Masked pixels: q = (#source samples <= x) / ns  (the source CDF at x, color.py:44-47), then
// np.interp(q, r_quant, r_values) (:49) with r_quant[j] = (end[j] + 1) / nr the reference CDF at its j-th distinct value
// r_values[j] = ref_sorted[end[j]] — numpy's arithmetic: one division per quantile, slope * (x - xp[j]) + fp[j] without
// contraction — cast to float32; every pixel is then clipped to [0, 1] (:60).
__global__ void __launch_bounds__(256) hist_match_kernel(const float* __restrict__ src, const uint8_t* __restrict__ mask,
                                                         long long n, const float* __restrict__ src_sorted, long long ns,
                                                         const float* __restrict__ ref_sorted, long long nr,
                                                         const int* __restrict__ end, long long nu, float* __restrict__ out) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        float x = src[p];
        if (mask[p]) {
            const double q = (double)count_le(src_sorted, ns, x) / (double)ns;
            const double x_first = (double)(end[0] + 1) / (double)nr, x_last = (double)(end[nu - 1] + 1) / (double)nr;
            double r;
            if (q > x_last) {
                r = (double)ref_sorted[end[nu - 1]];
            } else if (q < x_first) {
                r = (double)ref_sorted[end[0]];
            } else {
                long long lo = 0, hi = nu;                       // largest j with r_quant[j] <= q
                while (hi - lo > 1) {
                    const long long mid = (lo + hi) >> 1;
                    if ((double)(end[mid] + 1) / (double)nr <= q) lo = mid;
                    else hi = mid;
                }
                const double xj = (double)(end[lo] + 1) / (double)nr, fj = (double)ref_sorted[end[lo]];
                if (lo == nu - 1 || xj == q) {
                    r = fj;
                } else {
                    const double xk = (double)(end[lo + 1] + 1) / (double)nr, fk = (double)ref_sorted[end[lo + 1]];
                    const double slope = __ddiv_rn(__dsub_rn(fk, fj), __dsub_rn(xk, xj));
                    r = __dadd_rn(__dmul_rn(slope, __dsub_rn(q, xj)), fj);
                }
            }
            x = (float)r;
        }
        out[p] = x < 0.f ? 0.f : (x > 1.f ? 1.f : x);             // NaN survives, as np.clip
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

int run_ends_impl(const float* sorted, long long n, uint8_t* flags, cudaStream_t stream) {
    HSR_REQUIRE(sorted && flags, HSR_EINVAL, "null sorted / flags pointer");
    HSR_REQUIRE(n >= 0, HSR_EINVAL, "negative n");
    if (n == 0) return HSR_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8;
    run_end_kernel<<<(unsigned int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(sorted, n, flags);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int hist_match_impl(const float* src, const uint8_t* mask, long long n, const float* src_sorted, long long ns,
                    const float* ref_sorted, long long nr, const int* ref_run_ends, long long nu, float* out,
                    cudaStream_t stream) {
    HSR_REQUIRE(src && mask && src_sorted && ref_sorted && ref_run_ends && out, HSR_EINVAL, "null pointer");
    HSR_REQUIRE(n >= 0 && ns >= 1 && nr >= 1 && nu >= 1 && nu <= nr && nr < 2147483647LL, HSR_ERANGE,
                "need n >= 0, ns >= 1, 1 <= nu <= nr < 2^31 (got %lld, %lld, %lld, %lld)", n, ns, nu, nr);
    if (n == 0) return HSR_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8;
    hist_match_kernel<<<(unsigned int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(src, mask, n, src_sorted, ns, ref_sorted,
                                                                                     nr, ref_run_ends, nu, out);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int stretch64_impl(const float* x, long long xks, long long xgs, const double* lohi, const uint8_t* mask, long long n,
                   int K, int G, double* out, long long oks, long long ogs, cudaStream_t stream) {
    HSR_REQUIRE(x && lohi && out, HSR_EINVAL, "null x / lohi / out pointer");
    HSR_REQUIRE(n >= 0 && K >= 1 && G >= 1 && (long long)K * G <= 65535, HSR_ERANGE,
                "bad n = %lld, K = %d or G = %d (K * G <= 65535)", n, K, G);
    HSR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, HSR_EALIGN,
                "x not 4-byte / out not 8-byte aligned");
    if (n == 0) return HSR_OK;
    Stretch64Params P{};
    P.x = x, P.xks = xks, P.xgs = xgs, P.out = out, P.oks = oks, P.ogs = ogs, P.lohi = lohi, P.mask = mask, P.n = n, P.G = G;
    const long long S = (long long)K * G;
    long long per = (n + 256 * 4 - 1) / (256 * 4);
    long long cap = (long long)device_sm_count() * 6 / S;
    if (cap < 1) cap = 1;
    dim3 grid((unsigned int)(per < cap ? per : cap), (unsigned int)S);
    stretch64_kernel<<<grid, 256, 0, stream>>>(P);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int notnan_mask_impl(const float* x, const uint8_t* base, long long n, uint8_t* out, cudaStream_t stream) {
    HSR_REQUIRE(x && out, HSR_EINVAL, "null x / out pointer");
    HSR_REQUIRE(n >= 0, HSR_EINVAL, "negative n");
    if (n == 0) return HSR_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8;
    notnan_mask_kernel<<<(unsigned int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(x, base, n, out);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

size_t percentiles_workspace(int K, int G) {
    if (K < 1 || G < 1) return 0;
    const size_t S = (size_t)K * (size_t)G;
    return align_up(S * sizeof(SelSeries), 256) + S * SEL_TMAX * SEL_BINS * sizeof(unsigned int);
}

int masked_percentiles_impl(const float* x, long long xks, long long xgs, const uint8_t* mask, long long n, int K, int G,
                            const double* q, int Q, void* workspace, double* out, cudaStream_t stream) {
    HSR_REQUIRE(x && q && workspace && out, HSR_EINVAL, "null x / q / workspace / out pointer");
    HSR_REQUIRE(n >= 0 && n < (1LL << 33) && K >= 1 && G >= 1 && (long long)K * G <= 65535, HSR_ERANGE,
                "bad n = %lld (< 2^33), K = %d or G = %d (K * G <= 65535)", n, K, G);
    HSR_REQUIRE(Q >= 1 && Q <= SEL_QMAX, HSR_ERANGE, "Q = %d outside [1, %d]", Q, SEL_QMAX);
    HSR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0, HSR_EALIGN, "x not 4-byte aligned");
    HSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, HSR_EALIGN, "workspace not 256-byte aligned");
    const long long S = (long long)K * G;
    SelParams P{};
    P.x = x, P.xks = xks, P.xgs = xgs, P.mask = mask, P.n = n, P.G = G;
    P.st = reinterpret_cast<SelSeries*>(workspace);
    P.hist = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(workspace) +
                                             align_up((size_t)S * sizeof(SelSeries), 256));
    const long long s_or = (K > 1 ? xks : 0) | (G > 1 ? xgs : 0);
    const uintptr_t a16 = reinterpret_cast<uintptr_t>(x) | (uintptr_t)(s_or * 4);
    const uintptr_t a4 = reinterpret_cast<uintptr_t>(mask) | (uintptr_t)(G > 1 ? n : 0);
    P.vec = ((a16 & 15) == 0 && (a4 & 3) == 0) ? 1 : 0;

    HSR_CUDA(cudaMemsetAsync(P.hist, 0, (size_t)S * SEL_TMAX * SEL_BINS * sizeof(unsigned int), stream));
    select_init_kernel<<<(unsigned int)((S + 255) / 256), 256, 0, stream>>>(P.st, S);
    long long per = (n + (long long)SEL_THREADS * 16 - 1) / ((long long)SEL_THREADS * 16);
    if (per < 1) per = 1;
    long long cap = (long long)device_sm_count() * 4 / S;
    if (cap < 1) cap = 1;
    dim3 grid((unsigned int)(per < cap ? per : cap), (unsigned int)S);
    select_hist_kernel<0><<<grid, SEL_THREADS, 0, stream>>>(P);
    select_scan_kernel<0><<<(unsigned int)S, SEL_THREADS, 0, stream>>>(P.st, P.hist, q, Q, out);
    select_hist_kernel<1><<<grid, SEL_THREADS, 0, stream>>>(P);
    select_scan_kernel<1><<<(unsigned int)S, SEL_THREADS, 0, stream>>>(P.st, P.hist, q, Q, out);
    select_hist_kernel<2><<<grid, SEL_THREADS, 0, stream>>>(P);
    select_scan_kernel<2><<<(unsigned int)S, SEL_THREADS, 0, stream>>>(P.st, P.hist, q, Q, out);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int stretch_impl(const float* x, long long xks, long long xgs, const double* lohi, long long n, int K, int G,
                 float* out, long long oks, long long ogs, cudaStream_t stream) {
    HSR_REQUIRE(x && lohi && out, HSR_EINVAL, "null x / lohi / out pointer");
    HSR_REQUIRE(n >= 0 && K >= 1 && G >= 1 && (long long)K * G <= 65535, HSR_ERANGE,
                "bad n = %lld, K = %d or G = %d (K * G <= 65535)", n, K, G);
    HSR_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 3) == 0, HSR_EALIGN,
                "x / out not 4-byte aligned");
    if (n == 0) return HSR_OK;
    StretchParams P{};
    P.x = x, P.xks = xks, P.xgs = xgs, P.out = out, P.oks = oks, P.ogs = ogs, P.lohi = lohi, P.n = n, P.G = G;
    const long long s_or = (K > 1 ? (xks | oks) : 0) | (G > 1 ? (xgs | ogs) : 0);
    P.vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | (uintptr_t)(s_or * 4)) & 15) == 0;
    const long long S = (long long)K * G;
    long long per = (n + 256 * 8 - 1) / (256 * 8);
    long long cap = (long long)device_sm_count() * 6 / S;
    if (cap < 1) cap = 1;
    dim3 grid((unsigned int)(per < cap ? per : cap), (unsigned int)S);
    stretch_kernel<<<grid, 256, 0, stream>>>(P);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

}  // namespace hsr
