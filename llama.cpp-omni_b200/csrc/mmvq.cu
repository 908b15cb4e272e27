// mmvq.cu — decode matvec  y[m] = W[m,k] . x[k]  for quantised W (q4_0, q8_0, q4_K, q5_K, q6_K), up to 8 columns.
//
// Replaces mul_mat_vec_q (ggml-cuda/mmvq.cu:139-227 + vecdotq.cuh).  B200-first design:
//   * HBM-bound: every weight byte is read exactly once with 128-bit streaming loads (ld.global.nc.L1::no_allocate),
//     8 / 4 / 2 / 1 lanes per block so that a warp instruction covers 512 contiguous bytes of the row.
//   * arithmetic = the CPU oracle's: activations pre-quantised to q8_K / q8_0 (quant_act.cu), integer sub-block dot
//     products with dp4a, one float multiply per (block, lane) — ggml-cpu/quants.c:115-149, 305-333, 550-758.
//   * persistent grid sized to the SM count; several weight matrices sharing one activation (q/k/v, gate/up) run as
//     "jobs" of ONE launch; optional fused residual add and swiglu epilogues remove 3 launches per layer.
#include "common.cuh"
#include "stream_decode.cuh"

namespace b200 {

constexpr int MAX_JOBS = 4;
constexpr int MMVQ_WARPS = 8;

struct MatvecArgs {
    const uint8_t * payload[MAX_JOBS];   // payload plane (native: block base)
    const uint8_t * dplane[MAX_JOBS];    // f16 d plane (native: block base + d_off)
    int64_t row_stride_p[MAX_JOBS];      // bytes between rows in the payload plane
    int64_t row_stride_d[MAX_JOBS];      // bytes between rows in the d plane
    int32_t blk_stride_p[MAX_JOBS];      // bytes between blocks (payload)
    int32_t blk_stride_d[MAX_JOBS];      // bytes between blocks (d)
    float * y[MAX_JOBS];
    const float * residual[MAX_JOBS];
    int64_t row_begin[MAX_JOBS + 1];     // prefix sum of m over jobs
    int64_t y_col_stride[MAX_JOBS];      // elements between columns of y
    int njobs;
    const uint8_t * act; int64_t act_bytes, act_d_off, act_bsum_off;
    int64_t k;
    int swiglu;                          // 1: jobs 0/1 are gate/up of equal m, y[0] = silu(g)*u
};

__device__ __forceinline__ int dp4a(int a, int b, int c) { return __dp4a(a, b, c); }
__device__ __forceinline__ uint4 lda16(const uint8_t * p) { return *(const uint4 *) p; }     // activation record: L1-resident

// ---- per-type partial dot products: one (block, lane-part) against NCOLS activation records ------------------------
template <int T, bool AL, int NC> struct BlockDot;

// q4_K: 8 lanes per block; lane part c owns qs[16c, 16c+16): 16 low nibbles of sub-block 2j and 16 high nibbles of 2j+1
template <bool AL, int NC> struct BlockDot<B200_Q4_K, AL, NC> {
    static constexpr int LPB = 8;
    __device__ static __forceinline__ void run(const uint8_t * pb, const uint8_t * /*db*/, int c, int64_t b, const MatvecArgs & A, float * acc) {
        const uint4 hdr = ld16_w<AL>(pb);
        const uint4 qs  = ld16_w<AL>(pb + 16 + 16 * c);
        const int j = c >> 1, half = c & 1;
        const float dw = h2f(hdr.x & 0xffff), dmin = h2f(hdr.x >> 16);
        const uint32_t sc03 = hdr.y & 0x3f3f3f3fu, mn03 = hdr.z & 0x3f3f3f3fu;
        const uint32_t sc47 = (hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4);
        const uint32_t mn47 = ((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4);
        const uint32_t scw = j < 2 ? sc03 : sc47, mnw = j < 2 ? mn03 : mn47;
        const int sh = (j & 1) * 16;
        const int sc_lo = (scw >> sh) & 0xff, sc_hi = (scw >> (sh + 8)) & 0xff;
        const int mn_lo = (mnw >> sh) & 0xff, mn_hi = (mnw >> (sh + 8)) & 0xff;
        const uint32_t w[4] = { qs.x, qs.y, qs.z, qs.w };
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            const uint8_t * rec = A.act + n * A.act_bytes;
            const uint4 a_lo = lda16(rec + b * 256 + 64 * j + 16 * half);
            const uint4 a_hi = lda16(rec + b * 256 + 64 * j + 32 + 16 * half);
            const int al[4] = { (int) a_lo.x, (int) a_lo.y, (int) a_lo.z, (int) a_lo.w };
            const int ah[4] = { (int) a_hi.x, (int) a_hi.y, (int) a_hi.z, (int) a_hi.w };
            int dlo = 0, dhi = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { dlo = dp4a((int) (w[i] & 0x0f0f0f0fu), al[i], dlo); dhi = dp4a((int) ((w[i] >> 4) & 0x0f0f0f0fu), ah[i], dhi); }
            const int16_t * bs = (const int16_t *) (rec + A.act_bsum_off) + b * 16 + 4 * j + half;
            const int isum = sc_lo * dlo + sc_hi * dhi;
            const int msum = mn_lo * (int) bs[0] + mn_hi * (int) bs[2];
            const float da = ((const float *) (rec + A.act_d_off))[b];
            acc[n] += (dw * da) * (float) isum - (dmin * da) * (float) msum;
        }
    }
};

// q5_K: as q4_K plus one high bit per weight from qh[32] (bit 2j for the low-nibble sub-block, 2j+1 for the high one)
template <bool AL, int NC> struct BlockDot<B200_Q5_K, AL, NC> {
    static constexpr int LPB = 8;
    __device__ static __forceinline__ void run(const uint8_t * pb, const uint8_t *, int c, int64_t b, const MatvecArgs & A, float * acc) {
        const uint4 hdr = ld16_w<AL>(pb);
        const int j = c >> 1, half = c & 1;
        const uint4 qh  = ld16_w<AL>(pb + 16 + 16 * half);
        const uint4 qs  = ld16_w<AL>(pb + 48 + 16 * c);
        const float dw = h2f(hdr.x & 0xffff), dmin = h2f(hdr.x >> 16);
        const uint32_t sc03 = hdr.y & 0x3f3f3f3fu, mn03 = hdr.z & 0x3f3f3f3fu;
        const uint32_t sc47 = (hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4);
        const uint32_t mn47 = ((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4);
        const uint32_t scw = j < 2 ? sc03 : sc47, mnw = j < 2 ? mn03 : mn47;
        const int sh = (j & 1) * 16;
        const int sc_lo = (scw >> sh) & 0xff, sc_hi = (scw >> (sh + 8)) & 0xff;
        const int mn_lo = (mnw >> sh) & 0xff, mn_hi = (mnw >> (sh + 8)) & 0xff;
        const uint32_t w[4] = { qs.x, qs.y, qs.z, qs.w }, h[4] = { qh.x, qh.y, qh.z, qh.w };
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t hb = h[i] >> (2 * j);
            lo[i] = (w[i] & 0x0f0f0f0fu) | ((hb & 0x01010101u) << 4);
            hi[i] = ((w[i] >> 4) & 0x0f0f0f0fu) | (((hb >> 1) & 0x01010101u) << 4);
        }
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            const uint8_t * rec = A.act + n * A.act_bytes;
            const uint4 a_lo = lda16(rec + b * 256 + 64 * j + 16 * half);
            const uint4 a_hi = lda16(rec + b * 256 + 64 * j + 32 + 16 * half);
            const int al[4] = { (int) a_lo.x, (int) a_lo.y, (int) a_lo.z, (int) a_lo.w };
            const int ah[4] = { (int) a_hi.x, (int) a_hi.y, (int) a_hi.z, (int) a_hi.w };
            int dlo = 0, dhi = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { dlo = dp4a((int) lo[i], al[i], dlo); dhi = dp4a((int) hi[i], ah[i], dhi); }
            const int16_t * bs = (const int16_t *) (rec + A.act_bsum_off) + b * 16 + 4 * j + half;
            const int isum = sc_lo * dlo + sc_hi * dhi;
            const int msum = mn_lo * (int) bs[0] + mn_hi * (int) bs[2];
            const float da = ((const float *) (rec + A.act_d_off))[b];
            acc[n] += (dw * da) * (float) isum - (dmin * da) * (float) msum;
        }
    }
};

// q6_K: 4 lanes per block; part p = (half h, l0 = 16*(p&1)) owns 16 positions l of half h -> 64 weights (4 quads of 16)
template <bool AL, int NC> struct BlockDot<B200_Q6_K, AL, NC> {
    static constexpr int LPB = 4;
    __device__ static __forceinline__ void run(const uint8_t * pb, const uint8_t * db, int p, int64_t b, const MatvecArgs & A, float * acc) {
        const int h = p >> 1, s = p & 1;
        const uint4 qa = ld16_w<AL>(pb + 64 * h + 16 * s);            // ql[64h + l0 ..]       quads 0 (lo nibble) and 2 (hi nibble)
        const uint4 qb = ld16_w<AL>(pb + 64 * h + 32 + 16 * s);       // ql[64h + 32 + l0 ..]  quads 1 and 3
        const uint4 qh = ld16_w<AL>(pb + 128 + 32 * h + 16 * s);
        const uint4 scv = ld16_w<AL>(pb + 192);
        const float dw = h2f(__ldg((const uint16_t *) db));
        const uint32_t scw0 = h ? scv.z : scv.x, scw1 = h ? scv.w : scv.y;    // scales[8h .. 8h+7]
        int sc[4];
        sc[0] = (int) (int8_t) (scw0 >> (8 * s));  sc[1] = (int) (int8_t) (scw0 >> (8 * s + 16));
        sc[2] = (int) (int8_t) (scw1 >> (8 * s));  sc[3] = (int) (int8_t) (scw1 >> (8 * s + 16));
        const uint32_t a[4] = { qa.x, qa.y, qa.z, qa.w }, bq[4] = { qb.x, qb.y, qb.z, qb.w }, hh[4] = { qh.x, qh.y, qh.z, qh.w };
        uint32_t w[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            w[0][i] = (a[i]  & 0x0f0f0f0fu)       | ((hh[i] << 4) & 0x30303030u);
            w[1][i] = (bq[i] & 0x0f0f0f0fu)       | ((hh[i] << 2) & 0x30303030u);
            w[2][i] = ((a[i]  >> 4) & 0x0f0f0f0fu) | ( hh[i]       & 0x30303030u);
            w[3][i] = ((bq[i] >> 4) & 0x0f0f0f0fu) | ((hh[i] >> 2) & 0x30303030u);
        }
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            const uint8_t * rec = A.act + n * A.act_bytes;
            const int16_t * bs = (const int16_t *) (rec + A.act_bsum_off) + b * 16 + 8 * h + s;
            int isum = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 av = lda16(rec + b * 256 + 128 * h + 32 * q + 16 * s);
                int d = dp4a((int) w[q][0], (int) av.x, 0);
                d = dp4a((int) w[q][1], (int) av.y, d); d = dp4a((int) w[q][2], (int) av.z, d); d = dp4a((int) w[q][3], (int) av.w, d);
                isum += sc[q] * (d - 32 * (int) bs[2 * q]);
            }
            const float da = ((const float *) (rec + A.act_d_off))[b];
            acc[n] += (dw * da) * (float) isum;
        }
    }
};

// q8_0: 2 lanes per block (16 int8 each); the pair's integer sums are added before the float multiply (oracle order)
template <bool AL, int NC> struct BlockDot<B200_Q8_0, AL, NC> {
    static constexpr int LPB = 2;
    __device__ static __forceinline__ void run(const uint8_t * pb, const uint8_t * db, int p, int64_t b, const MatvecArgs & A, float * acc) {
        const uint4 qs = ld16_w<AL>(pb + 16 * p);
        const float dw = h2f(__ldg((const uint16_t *) db));
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            const uint8_t * rec = A.act + n * A.act_bytes;
            const uint4 av = lda16(rec + b * 32 + 16 * p);
            int d = dp4a((int) qs.x, (int) av.x, 0);
            d = dp4a((int) qs.y, (int) av.y, d); d = dp4a((int) qs.z, (int) av.z, d); d = dp4a((int) qs.w, (int) av.w, d);
            d += __shfl_xor_sync(3u << (threadIdx.x & 30), d, 1);   // pair-local mask: other lanes may have left the loop
            const float da = __half2float(((const __half *) (rec + A.act_d_off))[b]);
            if (p == 0) acc[n] += (float) d * (dw * da);
        }
    }
};

// q4_0: 1 lane per block: 16 bytes = 32 nibbles; codes are offset by 8 -> subtract 8 * sum(q8) via the 32-wide bsum
template <bool AL, int NC> struct BlockDot<B200_Q4_0, AL, NC> {
    static constexpr int LPB = 1;
    __device__ static __forceinline__ void run(const uint8_t * pb, const uint8_t * db, int, int64_t b, const MatvecArgs & A, float * acc) {
        const uint4 qs = ld16_w<AL>(pb);
        const float dw = h2f(__ldg((const uint16_t *) db));
        const uint32_t w[4] = { qs.x, qs.y, qs.z, qs.w };
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            const uint8_t * rec = A.act + n * A.act_bytes;
            const uint4 a0 = lda16(rec + b * 32), a1 = lda16(rec + b * 32 + 16);
            const int al[4] = { (int) a0.x, (int) a0.y, (int) a0.z, (int) a0.w }, ah[4] = { (int) a1.x, (int) a1.y, (int) a1.z, (int) a1.w };
            int d = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { d = dp4a((int) (w[i] & 0x0f0f0f0fu), al[i], d); d = dp4a((int) ((w[i] >> 4) & 0x0f0f0f0fu), ah[i], d); }
            d -= 8 * (int) ((const int16_t *) (rec + A.act_bsum_off))[b];
            const float da = __half2float(((const __half *) (rec + A.act_d_off))[b]);
            acc[n] += ((float) d * dw) * da;
        }
    }
};

// ---- the kernel: persistent warps, one row per warp-iteration ----------------------------------------------------------
template <int T, bool AL, int NC>
__global__ void __launch_bounds__(MMVQ_WARPS * 32) k_mmvq(const __grid_constant__ MatvecArgs A) {
    using BD = BlockDot<T, AL, NC>;
    constexpr int LPB = BD::LPB, BPI = 32 / LPB;                      // lanes per block, blocks per warp-iteration
    const int lane = threadIdx.x & 31, part = lane % LPB, bsub = lane / LPB;
    const int64_t nblk = A.k / qtraits<T>::qk;
    const int64_t total_rows = A.swiglu ? A.row_begin[1] : A.row_begin[A.njobs];
    const int64_t warp0 = (int64_t) blockIdx.x * MMVQ_WARPS + (threadIdx.x >> 5), nwarps = (int64_t) gridDim.x * MMVQ_WARPS;
    for (int64_t g = warp0; g < total_rows; g += nwarps) {
        int job = 0;
#pragma unroll
        for (int jn = 1; jn < MAX_JOBS; ++jn) if (jn < A.njobs && !A.swiglu && g >= A.row_begin[jn]) job = jn;
        const int64_t r = g - A.row_begin[job];
        float acc[NC], acc2[NC];
#pragma unroll
        for (int n = 0; n < NC; ++n) { acc[n] = 0.0f; acc2[n] = 0.0f; }
        {
            const uint8_t * prow = A.payload[job] + r * A.row_stride_p[job];
            const uint8_t * drow = A.dplane[job]  + r * A.row_stride_d[job];
            const int bp = A.blk_stride_p[job], bd = A.blk_stride_d[job];
#pragma unroll 4
            for (int64_t b = bsub; b < nblk; b += BPI) BD::run(prow + b * bp, drow + b * bd, part, b, A, acc);
        }
        if (A.swiglu) {
            const uint8_t * prow = A.payload[1] + r * A.row_stride_p[1];
            const uint8_t * drow = A.dplane[1]  + r * A.row_stride_d[1];
            const int bp = A.blk_stride_p[1], bd = A.blk_stride_d[1];
#pragma unroll 4
            for (int64_t b = bsub; b < nblk; b += BPI) BD::run(prow + b * bp, drow + b * bd, part, b, A, acc2);
        }
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            float v = warp_sum(acc[n]);
            if (A.swiglu) {
                const float u = warp_sum(acc2[n]);
                v = (v / (1.0f + expf(-v))) * u;
            }
            if (lane == 0) {
                const int64_t o = n * A.y_col_stride[job] + r;
                if (A.residual[job]) v += A.residual[job][o];
                A.y[job][o] = v;
            }
        }
    }
}

template <int T, bool AL, int NC>
static int launch_mmvq(const MatvecArgs & A, cudaStream_t st) {
    const int64_t rows = A.swiglu ? A.row_begin[1] : A.row_begin[A.njobs];
    int64_t ctas = (rows + MMVQ_WARPS - 1) / MMVQ_WARPS;
    const int64_t cap = (int64_t) sm_count() * 8;                     // 8 CTAs x 8 warps = 64 resident warps per SM
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    k_mmvq<T, AL, NC><<<(unsigned) ctas, MMVQ_WARPS * 32, 0, st>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <int T, bool AL>
static int dispatch_nc(const MatvecArgs & A, int nc, cudaStream_t st) {
    switch (nc) {
        case 1: return launch_mmvq<T, AL, 1>(A, st);  case 2: return launch_mmvq<T, AL, 2>(A, st);
        case 3: return launch_mmvq<T, AL, 3>(A, st);  case 4: return launch_mmvq<T, AL, 4>(A, st);
        case 5: return launch_mmvq<T, AL, 5>(A, st);  case 6: return launch_mmvq<T, AL, 6>(A, st);
        case 7: return launch_mmvq<T, AL, 7>(A, st);  case 8: return launch_mmvq<T, AL, 8>(A, st);
    }
    return B200_ERR_UNSUPPORTED;
}

template <int T>
static int dispatch_al(const MatvecArgs & A, bool aligned, int nc, cudaStream_t st) {
    return aligned ? dispatch_nc<T, true>(A, nc, st) : dispatch_nc<T, false>(A, nc, st);
}

// fill plane pointers / strides for one job; returns whether all 16-byte loads are aligned
static bool setup_job(MatvecArgs & A, int i, const void * w, int type, int layout, int64_t m, int64_t row_stride, int64_t k) {
    const int qk = blck_size(type), bytes = type_size(type), pay = payload_size(type);
    const int64_t nblk = k / qk;
    const uint8_t * base = (const uint8_t *) w;
    if (layout == B200_LAYOUT_PLANAR && pay != bytes) {
        A.payload[i] = base;                       A.blk_stride_p[i] = pay;  A.row_stride_p[i] = nblk * pay;
        A.dplane[i]  = base + m * nblk * pay;      A.blk_stride_d[i] = 2;    A.row_stride_d[i] = nblk * 2;
    } else {
        const int p_off = (type == B200_Q4_0 || type == B200_Q8_0) ? 2 : 0, d_off = type == B200_Q6_K ? 208 : 0;
        A.payload[i] = base + p_off;               A.blk_stride_p[i] = bytes; A.row_stride_p[i] = row_stride;
        A.dplane[i]  = base + d_off;               A.blk_stride_d[i] = bytes; A.row_stride_d[i] = row_stride;
    }
    return ((uintptr_t) A.payload[i] % 16 == 0) && (A.blk_stride_p[i] % 16 == 0) && (A.row_stride_p[i] % 16 == 0);
}

int matvec_launch(MatvecArgs & A, int type, bool aligned, int ncols, cudaStream_t st) {
    switch (type) {
        case B200_Q4_0: return dispatch_al<B200_Q4_0>(A, aligned, ncols, st);
        case B200_Q8_0: return dispatch_al<B200_Q8_0>(A, aligned, ncols, st);
        case B200_Q4_K: return dispatch_al<B200_Q4_K>(A, aligned, ncols, st);
        case B200_Q5_K: return dispatch_al<B200_Q5_K>(A, aligned, ncols, st);
        case B200_Q6_K: return dispatch_al<B200_Q6_K>(A, aligned, ncols, st);
    }
    return B200_ERR_UNSUPPORTED;
}

static bool try_stream(const b200_matvec_job * jobs, int njobs, const void * act, int64_t k, bool swiglu, float * y_swiglu, cudaStream_t st, int & rc);

// used by mul_mat.cu: single job, ncols <= 8, explicit y column stride
int matvec_q_cols(const void * w, int type, int layout, int64_t m, int64_t row_stride, const void * act, int64_t k, int ncols,
                  float * y, int64_t y_col_stride, cudaStream_t st) {
    if (ncols == 1) {
        b200_matvec_job job = { w, type, layout, m, row_stride, y, nullptr };
        int rc = 0;
        if (try_stream(&job, 1, act, k, false, nullptr, st, rc)) return rc;
    }
    MatvecArgs A = {};
    const ActLayout L = act_layout(type, k);
    A.njobs = 1; A.act = (const uint8_t *) act; A.act_bytes = L.bytes; A.act_d_off = L.d_off; A.act_bsum_off = L.bsum_off; A.k = k;
    const bool al = setup_job(A, 0, w, type, layout, m, row_stride, k);
    A.y[0] = y; A.residual[0] = nullptr; A.row_begin[0] = 0; A.row_begin[1] = m; A.y_col_stride[0] = y_col_stride;
    return matvec_launch(A, type, al, ncols, st);
}

bool sd_fill_mat(SdMat & M, const void * w, int type, int layout, int64_t m, int64_t k, int64_t row_stride_bytes, float * y, const float * residual);
bool sd_phase_ok(const SdPhase & P);
int  sd_launch(const SdPhase * phases_dev, int n_phases, const SdPhase * single, unsigned * gbar, const SdRuntime & rt, cudaStream_t st);

// batch-1 jobs whose rows can be bulk-copied go to the persistent streaming kernel (stream_decode.cu); anything else (native
// q4_0/q8_0/q6_K blocks, strided rows, k beyond the shared-memory record) stays on k_mmvq above.
static bool try_stream(const b200_matvec_job * jobs, int njobs, const void * act, int64_t k, bool swiglu, float * y_swiglu, cudaStream_t st, int & rc) {
    if (njobs > 3) return false;
    SdPhase P = {};
    P.ksplit = 1; P.next_kind = -1; P.next_mv = -1;
    P.kind = SD_MATVEC; P.n_mat = njobs; P.epilogue = swiglu ? SD_EPI_SWIGLU : SD_EPI_STORE; P.prologue = SD_PRO_ACT;
    P.k = (int32_t) k; P.act_group = is_kquant(jobs[0].type) ? 256 : 32; P.act = (const uint8_t *) act;
    if ((uintptr_t) act % 16) return false;
    for (int i = 0; i < njobs; ++i) {
        if ((is_kquant(jobs[i].type) ? 256 : 32) != P.act_group) return false;
        if (!sd_fill_mat(P.mat[i], jobs[i].w, jobs[i].type, jobs[i].layout, jobs[i].m, k, jobs[i].row_stride_bytes,
                         swiglu ? y_swiglu : jobs[i].y, swiglu ? nullptr : jobs[i].residual)) return false;
    }
    if (!sd_phase_ok(P)) return false;
    SdRuntime rt = {};
    rc = sd_launch(nullptr, 1, &P, nullptr, rt, st);
    return true;
}

} // namespace b200

using namespace b200;

extern "C" int b200_matvec_q(const b200_matvec_job * jobs, int njobs, const void * act, int64_t k, void * stream) {
    if (njobs < 1 || njobs > MAX_JOBS || !jobs || !act) return B200_ERR_ARG;
    const int type = jobs[0].type;
    if (!is_quant(type) || k % blck_size(type) != 0) return B200_ERR_UNSUPPORTED;
    { int rc = 0; if (try_stream(jobs, njobs, act, k, false, nullptr, (cudaStream_t) stream, rc)) return rc; }
    MatvecArgs A = {};
    const ActLayout L = act_layout(type, k);
    A.njobs = njobs; A.act = (const uint8_t *) act; A.act_bytes = L.bytes; A.act_d_off = L.d_off; A.act_bsum_off = L.bsum_off; A.k = k;
    bool al = true;
    A.row_begin[0] = 0;
    for (int i = 0; i < njobs; ++i) {
        if (jobs[i].type != type) return B200_ERR_UNSUPPORTED;       // one launch = one weight type (template instance)
        al = setup_job(A, i, jobs[i].w, type, jobs[i].layout, jobs[i].m, jobs[i].row_stride_bytes, k) && al;
        A.y[i] = jobs[i].y; A.residual[i] = jobs[i].residual; A.row_begin[i + 1] = A.row_begin[i] + jobs[i].m; A.y_col_stride[i] = jobs[i].m;
    }
    return matvec_launch(A, type, al, 1, (cudaStream_t) stream);
}

extern "C" int b200_matvec_q_swiglu(const b200_matvec_job * gate, const b200_matvec_job * up, float * y, const void * act, int64_t k,
                                    void * stream) {
    if (!gate || !up || !y || !act) return B200_ERR_ARG;
    const int type = gate->type;
    if (!is_quant(type) || up->type != type || up->m != gate->m || k % blck_size(type) != 0) return B200_ERR_UNSUPPORTED;
    { const b200_matvec_job two[2] = { *gate, *up }; int rc = 0; if (try_stream(two, 2, act, k, true, y, (cudaStream_t) stream, rc)) return rc; }
    MatvecArgs A = {};
    const ActLayout L = act_layout(type, k);
    A.njobs = 2; A.swiglu = 1; A.act = (const uint8_t *) act; A.act_bytes = L.bytes; A.act_d_off = L.d_off; A.act_bsum_off = L.bsum_off; A.k = k;
    bool al = setup_job(A, 0, gate->w, type, gate->layout, gate->m, gate->row_stride_bytes, k);
    al = setup_job(A, 1, up->w, type, up->layout, up->m, up->row_stride_bytes, k) && al;
    A.y[0] = y; A.residual[0] = nullptr; A.y_col_stride[0] = gate->m;
    A.row_begin[0] = 0; A.row_begin[1] = gate->m; A.row_begin[2] = 2 * gate->m;
    return matvec_launch(A, type, al, 1, (cudaStream_t) stream);
}
