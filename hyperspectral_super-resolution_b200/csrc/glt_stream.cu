// glt_stream.cu — kernels 1 and 2 of the hot path: GLT-indexed ortho gather of the
// band-interleaved cube (bit-exact copy/fill) and the fused gather + SRF contraction.
//
// One persistent CTA per SM.  Warp 0 is the PRODUCER: lane i owns pixel i of a 32-pixel
// ortho tile, reads its GLT entry from a shared-memory ring that is itself filled by 1-D
// bulk copies (TMA, SASS UBLKCP) eight tiles ahead, evaluates the validity rule of
// EMIT_data/emit_proj.py:691-703 and issues ONE bulk copy of the 16-byte-aligned window that
// covers the pixel's spectrum (1140 B for 285 bands -> 1152 B window) into a stage slot.
// Warps 1..nstage are CONSUMERS: warp s owns stage s of the ring, waits on its mbarrier, and either re-aligns the windows into 16-byte vector stores of the ortho cube
// (emit_proj.py:981-982) or runs the SRF contraction of s2_emit/synth.py:41-43 with one lane
// per pixel, or both.  No data-path instruction touches the raw cube outside the TMA unit.
#include "hsr_common.cuh"

namespace hsr {

namespace {

constexpr int TILE = HSR_TILE_PX;  // ortho pixels per tile == lanes of the producer warp
constexpr int MAX_STAGES = 8;
// One consumer warp per stage: a stage's uses are then consumed in order by a single warp,
// which is what makes the mbarrier phase-parity test unambiguous.
constexpr int NTHREADS = 32 * (1 + MAX_STAGES);
constexpr int GLT_DEPTH = 8;  // GLT tiles staged ahead of the gather
constexpr int MAXK = HSR_MAX_SRF_BANDS;

constexpr int MODE_COPY = 1;
constexpr int MODE_SRF = 2;

constexpr int META_FILL = -1;  // pixel inside the grid, GLT invalid -> fill value
constexpr int META_OOB = -2;   // lane beyond the end of the grid (last tile only)

struct StreamParams {
    const float* raw;
    long long raw_h, raw_w, raw_pix_stride;
    int bands;
    int transpose;
    int identity;  // 1: no GLT, source pixel == output pixel (un-fused SRF)
    const int32_t* glt_x;
    const int32_t* glt_y;
    long long out_w, glt_row_stride, npix, ntiles;
    int glt_tma;  // GLT planes contiguous and 16-byte aligned -> staged with bulk copies
    float fill;
    float* ortho;
    long long out_pix_stride;
    uint8_t* valid;
    unsigned long long* diag;
    const float* W;
    const float* fill_out;
    int K;
    float* bands_out;
    long long plane_stride;
    int slot_f4;    // float4 per pixel slot (odd, so lane-per-pixel LDS.128 is conflict-free)
    int stage_f4;   // float4 per stage
    int nstage;
    int wt_pitch;   // floats per row of the transposed weight table in smem
    unsigned long long raw_lo, raw_hi;  // byte range of the raw cube that may be read
};

struct SmemHeader {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t glt_full[GLT_DEPTH];
    int meta[MAX_STAGES][TILE];
    int glt[GLT_DEPTH][2][TILE];
    int run_b0[MAXK];
    int run_len[MAXK];
    float fill_out[MAXK];
};
static_assert(offsetof(SmemHeader, glt) % 16 == 0, "GLT ring must be 16-byte aligned for bulk copies");

__host__ __device__ inline int header_bytes() { return (int)((sizeof(SmemHeader) + 127) / 128 * 128); }

// float4 offset of pixel slot l inside a stage: odd pitch + one float4 of skew per 8 lanes
__device__ __forceinline__ int slot_off_f4(int l, int slot_f4) { return l * slot_f4 + (l >> 3); }

template <int D>
__device__ __forceinline__ float4 realign(const float4 a, const float4 b) {
    if (D == 0) return a;
    if (D == 1) return make_float4(a.y, a.z, a.w, b.x);
    if (D == 2) return make_float4(a.z, a.w, b.x, b.y);
    return make_float4(a.w, b.x, b.y, b.z);
}

// Copy nfull aligned float4 of one pixel: source window words start D words into w4[0].
template <int D>
__device__ __forceinline__ void copy_pixel_f4(const float4* __restrict__ w4, float4* __restrict__ dst4, int nfull,
                                               int lane) {
    for (int base = 0; base < nfull; base += 96) {
        const int j0 = base + lane, j1 = j0 + 32, j2 = j0 + 64;
        const bool p0 = j0 < nfull, p1 = j1 < nfull, p2 = j2 < nfull;
        float4 a0, b0, a1, b1, a2, b2;
        a0 = b0 = a1 = b1 = a2 = b2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p0) { a0 = w4[j0]; if (D) b0 = w4[j0 + 1]; }
        if (p1) { a1 = w4[j1]; if (D) b1 = w4[j1 + 1]; }
        if (p2) { a2 = w4[j2]; if (D) b2 = w4[j2 + 1]; }
        if (p0) __stcs(dst4 + j0, realign<D>(a0, b0));
        if (p1) __stcs(dst4 + j1, realign<D>(a1, b1));
        if (p2) __stcs(dst4 + j2, realign<D>(a2, b2));
    }
}

// Consumer, ortho materialisation: one warp walks the 32 pixels of the tile; per pixel the
// lanes stream 16-byte stores (head/tail words of the unaligned 1140-byte record are scalar).
__device__ __forceinline__ void copy_tile(const StreamParams& P, const float4* __restrict__ st4, int m, long long tile,
                                          int lane) {
    const int B = P.bands;
    const float fillv = P.fill;
    const float4 fill4 = make_float4(fillv, fillv, fillv, fillv);
    for (int k = 0; k < TILE; ++k) {
        const int mk = __shfl_sync(0xffffffffu, m, k);
        if (mk == META_OOB) break;
        const long long p = tile * TILE + k;
        float* dst = P.ortho + p * P.out_pix_stride;
        const int dp = (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3);
        int h = (4 - dp) & 3;
        if (h > B) h = B;
        const int nfull = (B - h) >> 2;
        const int t = (B - h) & 3;
        float4* dst4 = reinterpret_cast<float4*>(dst + h);
        if (mk >= 0) {
            const float* wf = reinterpret_cast<const float*>(st4 + slot_off_f4(k, P.slot_f4));
            const int s = mk + h;
            const float4* w4 = reinterpret_cast<const float4*>(wf) + (s >> 2);
            switch (s & 3) {
                case 0: copy_pixel_f4<0>(w4, dst4, nfull, lane); break;
                case 1: copy_pixel_f4<1>(w4, dst4, nfull, lane); break;
                case 2: copy_pixel_f4<2>(w4, dst4, nfull, lane); break;
                default: copy_pixel_f4<3>(w4, dst4, nfull, lane); break;
            }
            if (lane < h) __stcs(dst + lane, wf[mk + lane]);
            const int tl = lane - 8;
            if (tl >= 0 && tl < t) __stcs(dst + h + 4 * nfull + tl, wf[s + 4 * nfull + tl]);
        } else {
            for (int j = lane; j < nfull; j += 32) __stcs(dst4 + j, fill4);
            if (lane < h) __stcs(dst + lane, fillv);
            const int tl = lane - 8;
            if (tl >= 0 && tl < t) __stcs(dst + h + 4 * nfull + tl, fillv);
        }
    }
}

// Consumer, SRF contraction: lane l owns pixel l of the tile.
//   1. non-finite scan of all `bands` samples (0*x accumulates NaN iff any sample is NaN/Inf):
//      in synth.py:41 a non-finite sample under a ZERO weight still poisons the integral;
//   2. per S2 band k, an fp32 FMA chain over the contiguous non-zero run of W[:,k]; pixels
//      flagged by the scan take a dense chain over all bands instead (exact IEEE propagation).
__device__ __forceinline__ void srf_tile(const StreamParams& P, const SmemHeader* hd, const float* __restrict__ wt,
                                         const float4* __restrict__ st4, int m, long long tile, int lane) {
    const bool ok = m >= 0;
    if (__ballot_sync(0xffffffffu, ok) == 0u) {  // whole tile is fill
        const long long p = tile * TILE + lane;
        if (m != META_OOB)
            for (int k = 0; k < P.K; ++k) P.bands_out[(long long)k * P.plane_stride + p] = hd->fill_out[k];
        return;
    }
    const int B = P.bands;
    const int sp = ok ? m : 0;
    const float4* w4 = st4 + slot_off_f4(lane, P.slot_f4);
    const float* wf = reinterpret_cast<const float*>(w4);

    // ---- 1. non-finite scan
    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
    {
        const int n0 = (B + 3) >> 2;  // float4 count of the window when sp == 0
        const int end = sp + B;       // first word index past the spectrum
        // float4 0 and the last two candidates are masked word by word
        auto masked = [&](int i) {
            const float4 v = w4[i];
            const int w = 4 * i;
            if (w + 0 >= sp && w + 0 < end) z0 = fmaf(v.x, 0.f, z0);
            if (w + 1 >= sp && w + 1 < end) z1 = fmaf(v.y, 0.f, z1);
            if (w + 2 >= sp && w + 2 < end) z2 = fmaf(v.z, 0.f, z2);
            if (w + 3 >= sp && w + 3 < end) z3 = fmaf(v.w, 0.f, z3);
        };
        if (ok) {
            masked(0);
            if (n0 >= 2) masked(n0 - 1);
            masked(n0);  // slot holds at least n0 + 1 float4
        }
        if (ok) {
#pragma unroll 4
            for (int i = 1; i < n0 - 1; ++i) {
                const float4 v = w4[i];
                z0 = fmaf(v.x, 0.f, z0);
                z1 = fmaf(v.y, 0.f, z1);
                z2 = fmaf(v.z, 0.f, z2);
                z3 = fmaf(v.w, 0.f, z3);
            }
        }
    }
    const float z = (z0 + z1) + (z2 + z3);
    const bool bad = !(z == 0.f);

    // ---- 2. per-band FMA over the non-zero run of the folded weights
    const long long p = tile * TILE + lane;
    const float* xs = wf + sp;
    for (int k = 0; k < P.K; ++k) {
        const int b0 = hd->run_b0[k];
        const int len = hd->run_len[k];
        const float* wk = wt + k * P.wt_pitch + b0;
        const float* xk = xs + b0;
        float acc0 = 0.f, acc1 = 0.f;
        int e = 0;
        for (; e + 1 < len; e += 2) {
            acc0 = fmaf(xk[e], wk[e], acc0);
            acc1 = fmaf(xk[e + 1], wk[e + 1], acc1);
        }
        if (e < len) acc0 = fmaf(xk[e], wk[e], acc0);
        float r = acc0 + acc1;
        if (bad) {
            // rare: redo this band densely over ALL samples so that IEEE propagation matches
            // synth.py:41 exactly (NaN anywhere or Inf under a zero weight -> NaN; Inf under a
            // non-zero weight -> +-Inf).
            const float* wd = wt + k * P.wt_pitch;
            r = 0.f;
            for (int b = 0; b < B; ++b) r = fmaf(xs[b], wd[b], r);
        }
        if (!ok) r = hd->fill_out[k];
        if (m != META_OOB) P.bands_out[(long long)k * P.plane_stride + p] = r;
    }
}

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) glt_stream_kernel(const StreamParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SmemHeader* hd = reinterpret_cast<SmemHeader*>(smem_raw);
    float* wt = reinterpret_cast<float*>(smem_raw + header_bytes());
    const int wt_bytes = (MODE & MODE_SRF) ? ((P.K * P.wt_pitch * 4 + 127) / 128 * 128) : 0;
    float4* stages = reinterpret_cast<float4*>(smem_raw + header_bytes() + wt_bytes);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const long long bid = blockIdx.x;
    const long long grid = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(&hd->full[s], 1);
            mbar_init(&hd->empty[s], 1);
        }
        for (int g = 0; g < GLT_DEPTH; ++g) mbar_init(&hd->glt_full[g], 1);
        fence_mbar_init();
    }
    if (MODE & MODE_SRF) {
        // transposed weight table Wt[k][b] = W[b][k], zero padded
        const int total = P.K * P.wt_pitch;
        for (int i = tid; i < total; i += (int)blockDim.x) {
            const int k = i / P.wt_pitch, b = i - k * P.wt_pitch;
            wt[i] = (b < P.bands) ? P.W[(long long)b * P.K + k] : 0.f;
        }
        if (tid < P.K) hd->fill_out[tid] = P.fill_out ? P.fill_out[tid] : 0.f;
    }
    __syncthreads();
    if (MODE & MODE_SRF) {
        // contiguous non-zero run [b0, b0+len) of each folded response
        if (tid < P.K) {
            int first = -1, last = -1;
            for (int b = 0; b < P.bands; ++b) {
                if (!(wt[tid * P.wt_pitch + b] == 0.f)) {
                    if (first < 0) first = b;
                    last = b;
                }
            }
            hd->run_b0[tid] = first < 0 ? 0 : first;
            hd->run_len[tid] = first < 0 ? 0 : last - first + 1;
        }
        __syncthreads();
    }

    const int nstage = P.nstage;

    if (warp == 0) {
        // =================================================================== PRODUCER
        const int B = P.bands;
        const bool contiguous = P.glt_row_stride == P.out_w;
        const bool glt_ring = P.glt_tma && !P.identity;
        auto tile_is_full = [&](long long tile) { return tile * TILE + TILE <= P.npix; };
        auto issue_glt = [&](int g, long long tile) {  // lane 0 only
            mbar_arrive_expect_tx(&hd->glt_full[g], 2 * TILE * 4);
            bulk_g2s(&hd->glt[g][0][0], P.glt_x + tile * TILE, TILE * 4, &hd->glt_full[g]);
            bulk_g2s(&hd->glt[g][1][0], P.glt_y + tile * TILE, TILE * 4, &hd->glt_full[g]);
        };
        if (glt_ring && lane == 0) {
            for (int g = 0; g < GLT_DEPTH; ++g) {
                const long long tile = bid + g * grid;
                if (tile < P.ntiles && tile_is_full(tile)) issue_glt(g, tile);
            }
        }
        unsigned int cnt_nz = 0, cnt_ib = 0;
        for (long long it = 0;; ++it) {
            const long long tile = bid + it * grid;
            if (tile >= P.ntiles) break;
            const int stage = (int)(it % nstage);
            const unsigned int use = (unsigned int)(it / nstage);
            const long long p = tile * TILE + lane;
            const bool inb = p < P.npix;

            int gx = 0, gy = 0;
            if (!P.identity) {
                if (glt_ring && tile_is_full(tile)) {
                    const int g = (int)(it % GLT_DEPTH);
                    mbar_wait(&hd->glt_full[g], (unsigned int)(it / GLT_DEPTH) & 1u);
                    gx = hd->glt[g][0][lane];
                    gy = hd->glt[g][1][lane];
                    __syncwarp();
                    if (lane == 0) {
                        const long long nt = tile + GLT_DEPTH * grid;
                        if (nt < P.ntiles && tile_is_full(nt)) issue_glt(g, nt);
                    }
                } else if (inb) {
                    const long long gi = contiguous ? p : (p / P.out_w) * P.glt_row_stride + (p % P.out_w);
                    gx = __ldg(P.glt_x + gi);
                    gy = __ldg(P.glt_y + gi);
                }
            }
            // validity rule: emit_proj.py:691 (both != 0), :694 (1-based -> 0-based), :698-703 (in bounds)
            const bool nz = (gx != 0) && (gy != 0);
            long long x0 = (long long)gx - 1, y0 = (long long)gy - 1;
            bool ib = inb && nz && x0 >= 0 && x0 < P.raw_w && y0 >= 0 && y0 < P.raw_h;
            long long q = P.transpose ? x0 * P.raw_h + y0 : y0 * P.raw_w + x0;
            if (P.identity) {
                ib = inb;
                q = p;
            }
            if (!ib) q = 0;
            const float* src = P.raw + q * P.raw_pix_stride;
            const unsigned long long a = reinterpret_cast<unsigned long long>(src);
            const unsigned long long lo = a & ~15ull;
            const int sp = (int)((a & 15ull) >> 2);
            const unsigned int bytes = (unsigned int)((sp + B + 3) >> 2) * 16u;
            const bool slow = ib && (lo < P.raw_lo || lo + bytes > P.raw_hi);
            const bool fast = ib && !slow;

            if (inb && !P.identity) {
                if (P.valid) P.valid[p] = ib ? 1 : 0;
                cnt_nz += nz ? 1u : 0u;
                cnt_ib += ib ? 1u : 0u;
            }

            mbar_wait(&hd->empty[stage], (use & 1u) ^ 1u);

            hd->meta[stage][lane] = inb ? (ib ? sp : META_FILL) : META_OOB;
            float* slot = reinterpret_cast<float*>(stages + (long long)stage * P.stage_f4 + slot_off_f4(lane, P.slot_f4));
            if (slow) {  // window would cross the ends of the allocation: plain loads
                for (int b = 0; b < B; ++b) slot[sp + b] = __ldg(src + b);
                fence_proxy_async_smem();
            }
            const int tx = warp_sum(fast ? (int)bytes : 0);
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&hd->full[stage], (unsigned int)tx);
            if (fast) bulk_g2s(slot, reinterpret_cast<const void*>(lo), bytes, &hd->full[stage]);
        }
        if (P.diag && !P.identity) {
            const int snz = warp_sum((int)cnt_nz), sib = warp_sum((int)cnt_ib);
            if (lane == 0) {
                atomicAdd(P.diag + 0, (unsigned long long)snz);
                atomicAdd(P.diag + 1, (unsigned long long)sib);
                atomicAdd(P.diag + 2, (unsigned long long)(snz - sib));
            }
        }
    } else {
        // =================================================================== CONSUMERS
        const int stage = warp - 1;  // < nstage by construction of the launch
        unsigned int use = 0;
        for (long long it = stage;; it += nstage, ++use) {
            const long long tile = bid + it * grid;
            if (tile >= P.ntiles) break;
            mbar_wait(&hd->full[stage], use & 1u);
            const int m = hd->meta[stage][lane];
            const float4* st4 = stages + (long long)stage * P.stage_f4;
            if (MODE & MODE_COPY) copy_tile(P, st4, m, tile, lane);
            if (MODE & MODE_SRF) srf_tile(P, hd, wt, st4, m, tile, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&hd->empty[stage]);
        }
    }
}

// Fallback for short records (LOC / OBS planes, bands < 32): one thread per output element.
__global__ void __launch_bounds__(256) glt_small_kernel(const StreamParams P) {
    const long long total = P.npix * P.bands;
    unsigned int cnt_nz = 0, cnt_ib = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / P.bands;
        const int b = (int)(i - p * P.bands);
        const long long gi = (p / P.out_w) * P.glt_row_stride + (p % P.out_w);
        const int gx = __ldg(P.glt_x + gi), gy = __ldg(P.glt_y + gi);
        const bool nz = (gx != 0) && (gy != 0);
        const long long x0 = (long long)gx - 1, y0 = (long long)gy - 1;
        const bool ib = nz && x0 >= 0 && x0 < P.raw_w && y0 >= 0 && y0 < P.raw_h;
        float v = P.fill;
        if (ib) {
            const long long q = P.transpose ? x0 * P.raw_h + y0 : y0 * P.raw_w + x0;
            v = __ldg(P.raw + q * P.raw_pix_stride + b);
        }
        P.ortho[p * P.out_pix_stride + b] = v;
        if (b == 0) {
            if (P.valid) P.valid[p] = ib ? 1 : 0;
            cnt_nz += nz ? 1u : 0u;
            cnt_ib += ib ? 1u : 0u;
        }
    }
    if (P.diag) {
        const int snz = warp_sum((int)cnt_nz), sib = warp_sum((int)cnt_ib);
        if ((threadIdx.x & 31) == 0 && (snz | sib)) {
            atomicAdd(P.diag + 0, (unsigned long long)snz);
            atomicAdd(P.diag + 1, (unsigned long long)sib);
            atomicAdd(P.diag + 2, (unsigned long long)(snz - sib));
        }
    }
}

// ---------------------------------------------------------------------------- host side
int plan_smem(StreamParams& P, int mode, size_t* smem_bytes) {
    const int n0 = (P.bands + 3) / 4;
    P.slot_f4 = (n0 + 1) | 1;                  // >= n0 + 1 float4 and odd
    P.stage_f4 = (TILE * P.slot_f4 + 4 + 7) / 8 * 8;  // 128-byte multiple
    P.wt_pitch = (P.bands + 3) / 4 * 4;
    const size_t wt_bytes = (mode & MODE_SRF) ? ((size_t)P.K * P.wt_pitch * 4 + 127) / 128 * 128 : 0;
    const size_t fixed = (size_t)header_bytes() + wt_bytes;
    const size_t stage_bytes = (size_t)P.stage_f4 * 16;
    const size_t cap = (size_t)device_max_smem_optin();
    if (cap <= fixed + 1024) return HSR_ENOSMEM;
    long long ns = (long long)((cap - fixed - 1024) / stage_bytes);  // 1 KB headroom for base alignment
    if (ns > MAX_STAGES) ns = MAX_STAGES;
    if (ns < 2) return HSR_ENOSMEM;
    P.nstage = (int)ns;
    *smem_bytes = fixed + (size_t)ns * stage_bytes;
    return HSR_OK;
}

template <int MODE>
int launch_stream(StreamParams& P, cudaStream_t stream) {
    size_t smem = 0;
    int rc = plan_smem(P, MODE, &smem);
    if (rc != HSR_OK) {
        set_error("spectrum of %d bands does not fit the shared-memory staging ring", P.bands);
        return rc;
    }
    HSR_CUDA(cudaFuncSetAttribute(glt_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = device_sm_count();
    if (grid > P.ntiles) grid = P.ntiles;
    glt_stream_kernel<MODE><<<(unsigned int)grid, 32 * (1 + P.nstage), smem, stream>>>(P);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int check_common(const float* raw, long long raw_h, long long raw_w, int bands, long long raw_pix_stride,
                 const int32_t* glt_x, const int32_t* glt_y, long long out_h, long long out_w,
                 long long glt_row_stride) {
    HSR_REQUIRE(raw && glt_x && glt_y, HSR_EINVAL, "null raw / GLT pointer");
    HSR_REQUIRE(raw_h > 0 && raw_w > 0 && bands > 0, HSR_EINVAL, "raw shape must be positive (got %lld x %lld x %d)",
                raw_h, raw_w, bands);
    HSR_REQUIRE(out_h >= 0 && out_w >= 0, HSR_EINVAL, "negative ortho shape");
    HSR_REQUIRE(raw_pix_stride >= bands, HSR_EINVAL, "raw_pix_stride %lld < bands %d", raw_pix_stride, bands);
    HSR_REQUIRE(glt_row_stride >= out_w, HSR_EINVAL, "glt_row_stride %lld < out_w %lld", glt_row_stride, out_w);
    HSR_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 3) == 0, HSR_EALIGN, "raw is not 4-byte aligned");
    HSR_REQUIRE(((reinterpret_cast<uintptr_t>(glt_x) | reinterpret_cast<uintptr_t>(glt_y)) & 3) == 0, HSR_EALIGN,
                "GLT planes are not 4-byte aligned");
    return HSR_OK;
}

void fill_common(StreamParams& P, const float* raw, long long raw_h, long long raw_w, int bands,
                 long long raw_pix_stride, int transpose, const int32_t* glt_x, const int32_t* glt_y, long long out_h,
                 long long out_w, long long glt_row_stride, float fill) {
    P.raw = raw;
    P.raw_h = raw_h;
    P.raw_w = raw_w;
    P.raw_pix_stride = raw_pix_stride;
    P.bands = bands;
    P.transpose = transpose ? 1 : 0;
    P.identity = 0;
    P.glt_x = glt_x;
    P.glt_y = glt_y;
    P.out_w = out_w;
    P.glt_row_stride = glt_row_stride;
    P.npix = out_h * out_w;
    P.ntiles = (P.npix + TILE - 1) / TILE;
    P.glt_tma = (glt_row_stride == out_w || out_h <= 1) &&
                ((reinterpret_cast<uintptr_t>(glt_x) | reinterpret_cast<uintptr_t>(glt_y)) & 15) == 0;
    P.fill = fill;
    P.raw_lo = reinterpret_cast<unsigned long long>(raw);
    P.raw_hi = P.raw_lo + ((unsigned long long)(raw_h * raw_w - 1) * raw_pix_stride + bands) * 4ull;
}

}  // namespace

int glt_ortho_impl(const float* raw, long long raw_h, long long raw_w, int bands, long long raw_pix_stride,
                   int transpose, const int32_t* glt_x, const int32_t* glt_y, long long out_h, long long out_w,
                   long long glt_row_stride, float fill, float* out, long long out_pix_stride, uint8_t* valid,
                   unsigned long long* diag, cudaStream_t stream) {
    if (out_h == 0 || out_w == 0) return HSR_OK;  // empty grid: nothing to read or write
    int rc = check_common(raw, raw_h, raw_w, bands, raw_pix_stride, glt_x, glt_y, out_h, out_w, glt_row_stride);
    if (rc != HSR_OK) return rc;
    HSR_REQUIRE(out, HSR_EINVAL, "null output pointer");
    HSR_REQUIRE(out_pix_stride >= bands, HSR_EINVAL, "out_pix_stride %lld < bands %d", out_pix_stride, bands);
    HSR_REQUIRE((reinterpret_cast<uintptr_t>(out) & 3) == 0, HSR_EALIGN, "out is not 4-byte aligned");
    if (out_h == 0 || out_w == 0) return HSR_OK;
    StreamParams P{};
    fill_common(P, raw, raw_h, raw_w, bands, raw_pix_stride, transpose, glt_x, glt_y, out_h, out_w, glt_row_stride,
                fill);
    P.ortho = out;
    P.out_pix_stride = out_pix_stride;
    P.valid = valid;
    P.diag = diag;
    size_t smem = 0;
    if (bands >= 32 && plan_smem(P, MODE_COPY, &smem) == HSR_OK) return launch_stream<MODE_COPY>(P, stream);
    // short records: planes of the LOC / OBS cubes (emit_proj.py:1123-1131, :1217-1224)
    const long long total = P.npix * bands;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    glt_small_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(P);
    HSR_CUDA(cudaGetLastError());
    return HSR_OK;
}

int glt_srf_impl(const float* raw, long long raw_h, long long raw_w, int bands, long long raw_pix_stride,
                 int transpose, const int32_t* glt_x, const int32_t* glt_y, long long out_h, long long out_w,
                 long long glt_row_stride, float fill, const float* W, const float* fill_out, int K,
                 float* bands_out, long long bands_plane_stride, float* ortho_out, long long out_pix_stride,
                 uint8_t* valid, unsigned long long* diag, cudaStream_t stream) {
    if (out_h == 0 || out_w == 0) return HSR_OK;
    int rc = check_common(raw, raw_h, raw_w, bands, raw_pix_stride, glt_x, glt_y, out_h, out_w, glt_row_stride);
    if (rc != HSR_OK) return rc;
    HSR_REQUIRE(W && fill_out && bands_out, HSR_EINVAL, "null W / fill_out / bands_out pointer");
    HSR_REQUIRE(K >= 1 && K <= MAXK, HSR_ERANGE, "K = %d outside [1, %d]", K, MAXK);
    HSR_REQUIRE(bands_plane_stride >= out_h * out_w, HSR_EINVAL, "bands_plane_stride %lld < out_h*out_w",
                bands_plane_stride);
    HSR_REQUIRE(((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(bands_out)) & 3) == 0, HSR_EALIGN,
                "W / bands_out not 4-byte aligned");
    if (ortho_out) {
        HSR_REQUIRE(out_pix_stride >= bands, HSR_EINVAL, "out_pix_stride %lld < bands %d", out_pix_stride, bands);
        HSR_REQUIRE((reinterpret_cast<uintptr_t>(ortho_out) & 3) == 0, HSR_EALIGN, "ortho_out not 4-byte aligned");
    }
    if (out_h == 0 || out_w == 0) return HSR_OK;
    StreamParams P{};
    fill_common(P, raw, raw_h, raw_w, bands, raw_pix_stride, transpose, glt_x, glt_y, out_h, out_w, glt_row_stride,
                fill);
    P.W = W;
    P.fill_out = fill_out;
    P.K = K;
    P.bands_out = bands_out;
    P.plane_stride = bands_plane_stride;
    P.ortho = ortho_out;
    P.out_pix_stride = out_pix_stride;
    P.valid = valid;
    P.diag = diag;
    if (ortho_out) return launch_stream<MODE_COPY | MODE_SRF>(P, stream);
    return launch_stream<MODE_SRF>(P, stream);
}

int srf_impl(const float* cube, long long n_pix, int bands, long long pix_stride, const float* W, int K,
             float* bands_out, long long bands_plane_stride, cudaStream_t stream) {
    HSR_REQUIRE(cube && W && bands_out, HSR_EINVAL, "null cube / W / bands_out pointer");
    HSR_REQUIRE(n_pix >= 0 && bands > 0, HSR_EINVAL, "bad cube shape (%lld x %d)", n_pix, bands);
    HSR_REQUIRE(pix_stride >= bands, HSR_EINVAL, "pix_stride %lld < bands %d", pix_stride, bands);
    HSR_REQUIRE(K >= 1 && K <= MAXK, HSR_ERANGE, "K = %d outside [1, %d]", K, MAXK);
    HSR_REQUIRE(bands_plane_stride >= n_pix, HSR_EINVAL, "bands_plane_stride %lld < n_pix", bands_plane_stride);
    HSR_REQUIRE(((reinterpret_cast<uintptr_t>(cube) | reinterpret_cast<uintptr_t>(W) |
                  reinterpret_cast<uintptr_t>(bands_out)) & 3) == 0,
                HSR_EALIGN, "cube / W / bands_out not 4-byte aligned");
    if (n_pix == 0) return HSR_OK;
    StreamParams P{};
    P.raw = cube;
    P.raw_h = 1;
    P.raw_w = n_pix;
    P.raw_pix_stride = pix_stride;
    P.bands = bands;
    P.identity = 1;
    P.out_w = n_pix;
    P.glt_row_stride = n_pix;
    P.npix = n_pix;
    P.ntiles = (n_pix + TILE - 1) / TILE;
    P.raw_lo = reinterpret_cast<unsigned long long>(cube);
    P.raw_hi = P.raw_lo + ((unsigned long long)(n_pix - 1) * pix_stride + bands) * 4ull;
    P.W = W;
    P.K = K;
    P.bands_out = bands_out;
    P.plane_stride = bands_plane_stride;
    return launch_stream<MODE_SRF>(P, stream);
}

}  // namespace hsr
