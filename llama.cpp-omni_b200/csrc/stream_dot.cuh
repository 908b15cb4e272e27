// stream_dot.cuh — warp-level dot products of weight rows resident in shared memory against the q8 activation record (also in shared
// memory, int8 plane swizzled: act_qs_off<true>).  Included by stream_decode.cu INSIDE namespace b200.
//
// Lane mapping for the K-quants (q4_K, q5_K, q6_K): 8 lanes per 256-weight super-block, 4 super-blocks per warp pass.  The 8 lanes of
// a block form one shared-memory quarter-warp phase and read 8 different 16-byte bank groups of the block -> conflict free:
//   q4_K / q5_K  lane part c: qs[16c .. 16c+16): low nibbles = sub-block 2(c>>1), high nibbles = sub-block 2(c>>1)+1, positions 16(c&1)..
//   q6_K         lane part c = 4h + 2q + s: ql[64h + 32q + 16s ..+16): low nibbles = 16 weights of quad q, high nibbles = quad q+2,
//                with their 2 high bits from qh[32h + 16s ..+16) (bit pairs 2q and 2q+4) and int8 scales sc[8h + 2q + s], sc[.. + 4].
// Every lane therefore needs, per block, two 16-byte activation slices, their two 16-element sums and the block's d: a KFrag.  For
// k <= 4096 (16 blocks = 4 passes) the 4 KFrags of a lane are loop-invariant across rows and live in registers.
struct ActS { uint32_t qs, d, bsum; };     // shared-memory addresses of the record's planes
struct KFrag { uint4 lo, hi; int bs_lo, bs_hi; float d; };

template <int T> __device__ __forceinline__ KFrag kfrag_load(const ActS & A, int b, int c) {
    int off_lo, off_hi, bi_lo, bi_hi;
    if (T == B200_Q6_K) { const int h = c >> 2, q = (c >> 1) & 1, s = c & 1;
                          off_lo = b * 256 + 128 * h + 32 * q + 16 * s; off_hi = off_lo + 64; bi_lo = b * 16 + 8 * h + 2 * q + s; bi_hi = bi_lo + 4; }
    else                { const int j = c >> 1, half = c & 1;
                          off_lo = b * 256 + 64 * j + 16 * half;        off_hi = off_lo + 32; bi_lo = b * 16 + 4 * j + half;      bi_hi = bi_lo + 2; }
    KFrag f;
    f.lo = lds16(A.qs + (uint32_t) act_qs_off<true>(off_lo)); f.hi = lds16(A.qs + (uint32_t) act_qs_off<true>(off_hi));
    f.bs_lo = lds_s16(A.bsum + bi_lo * 2); f.bs_hi = lds_s16(A.bsum + bi_hi * 2);
    f.d = lds_f32(A.d + b * 4);
    return f;
}

__device__ __forceinline__ void kfrag_fill(int type, const ActS & A, int nblk, KFrag (&fr)[4]) {
    const int lane = threadIdx.x & 31, c = lane & 7;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int b = (lane >> 3) + 4 * t;
        if (b < nblk) fr[t] = type == B200_Q6_K ? kfrag_load<B200_Q6_K>(A, b, c) : kfrag_load<B200_Q4_K>(A, b, c);
    }
}

__device__ __forceinline__ int dot16(const uint32_t (&w)[4], const uint4 & a) {
    int d = __dp4a((int) w[0], (int) a.x, 0);
    d = __dp4a((int) w[1], (int) a.y, d); d = __dp4a((int) w[2], (int) a.z, d); d = __dp4a((int) w[3], (int) a.w, d);
    return d;
}

// one (block, lane part) against one fragment; pb = shared address of the block payload, dp = shared address of its f16 d (q6_K)
template <int T> __device__ __forceinline__ float kblock_dot(uint32_t pb, uint32_t dp, int c, const KFrag & f);

__device__ __forceinline__ void k4_scales(const uint4 & hdr, int j, int & sc_lo, int & sc_hi, int & mn_lo, int & mn_hi) {
    // get_scale_min_k4 (ggml-quants.c:703-711) for sub-blocks 2j and 2j+1, from the 12 packed bytes in hdr.y/z/w
    const uint32_t sc03 = hdr.y & 0x3f3f3f3fu, mn03 = hdr.z & 0x3f3f3f3fu;
    const uint32_t sc47 = (hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4);
    const uint32_t mn47 = ((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4);
    const uint32_t scw = j < 2 ? sc03 : sc47, mnw = j < 2 ? mn03 : mn47;
    const int sh = (j & 1) * 16;
    sc_lo = (scw >> sh) & 0xff; sc_hi = (scw >> (sh + 8)) & 0xff; mn_lo = (mnw >> sh) & 0xff; mn_hi = (mnw >> (sh + 8)) & 0xff;
}

template <> __device__ __forceinline__ float kblock_dot<B200_Q4_K>(uint32_t pb, uint32_t, int c, const KFrag & f) {
    const uint4 hdr = lds16(pb), qs = lds16(pb + 16 + 16 * c);
    int sc_lo, sc_hi, mn_lo, mn_hi; k4_scales(hdr, c >> 1, sc_lo, sc_hi, mn_lo, mn_hi);
    const uint32_t wl[4] = { qs.x & 0x0f0f0f0fu, qs.y & 0x0f0f0f0fu, qs.z & 0x0f0f0f0fu, qs.w & 0x0f0f0f0fu };
    const uint32_t wh[4] = { (qs.x >> 4) & 0x0f0f0f0fu, (qs.y >> 4) & 0x0f0f0f0fu, (qs.z >> 4) & 0x0f0f0f0fu, (qs.w >> 4) & 0x0f0f0f0fu };
    const int isum = sc_lo * dot16(wl, f.lo) + sc_hi * dot16(wh, f.hi);
    const int msum = mn_lo * f.bs_lo + mn_hi * f.bs_hi;
    const float dw = h2f(hdr.x & 0xffff), dmin = h2f(hdr.x >> 16);
    return (dw * f.d) * (float) isum - (dmin * f.d) * (float) msum;
}

template <> __device__ __forceinline__ float kblock_dot<B200_Q5_K>(uint32_t pb, uint32_t, int c, const KFrag & f) {
    const int j = c >> 1;
    const uint4 hdr = lds16(pb), qh = lds16(pb + 16 + 16 * (c & 1)), qs = lds16(pb + 48 + 16 * c);
    int sc_lo, sc_hi, mn_lo, mn_hi; k4_scales(hdr, j, sc_lo, sc_hi, mn_lo, mn_hi);
    const uint32_t q[4] = { qs.x, qs.y, qs.z, qs.w }, h[4] = { qh.x, qh.y, qh.z, qh.w };
    uint32_t wl[4], wh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t hb = h[i] >> (2 * j);
        wl[i] = (q[i] & 0x0f0f0f0fu) | ((hb & 0x01010101u) << 4);
        wh[i] = ((q[i] >> 4) & 0x0f0f0f0fu) | (((hb >> 1) & 0x01010101u) << 4);
    }
    const int isum = sc_lo * dot16(wl, f.lo) + sc_hi * dot16(wh, f.hi);
    const int msum = mn_lo * f.bs_lo + mn_hi * f.bs_hi;
    const float dw = h2f(hdr.x & 0xffff), dmin = h2f(hdr.x >> 16);
    return (dw * f.d) * (float) isum - (dmin * f.d) * (float) msum;
}

template <> __device__ __forceinline__ float kblock_dot<B200_Q6_K>(uint32_t pb, uint32_t dp, int c, const KFrag & f) {
    const int h = c >> 2, q = (c >> 1) & 1, s = c & 1;
    const uint4 ql = lds16(pb + 16 * c), qh = lds16(pb + 128 + 32 * h + 16 * s);
    const uint32_t scw = lds4(pb + 192 + 8 * h + 4 * (q >> 1));            // sc[8h .. 8h+3] : holds sc[8h + 2q + s] for q in {0, 1}
    const uint32_t scw2 = lds4(pb + 196 + 8 * h);                          // sc[8h+4 .. 8h+7]: holds sc[8h + 2q + s + 4]
    const int sc_lo = (int) (int8_t) (scw >> (8 * (2 * q + s))), sc_hi = (int) (int8_t) (scw2 >> (8 * (2 * q + s)));
    const uint32_t l[4] = { ql.x, ql.y, ql.z, ql.w }, hh[4] = { qh.x, qh.y, qh.z, qh.w };
    uint32_t wl[4], wh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t hb = hh[i] >> (2 * q);
        wl[i] = (l[i] & 0x0f0f0f0fu) | ((hb & 0x03030303u) << 4);
        wh[i] = ((l[i] >> 4) & 0x0f0f0f0fu) | (((hb >> 4) & 0x03030303u) << 4);
    }
    const int isum = sc_lo * (dot16(wl, f.lo) - 32 * f.bs_lo) + sc_hi * (dot16(wh, f.hi) - 32 * f.bs_hi);
    const float dw = h2f(lds_u16(dp));
    return (dw * f.d) * (float) isum;
}

// ---- register path (K-slice <= 4096 = 16 super-blocks): TWO lanes per super-block, the whole row in ONE warp pass -------------------
// Lane (b = lane >> 1, h = lane & 1) owns half h of block b: 128 consecutive weights, i.e. 128 consecutive int8 activations (8 x 16 B), their
// eight 16-element sums and the block's d — the same fragment for q4_K and q6_K, loaded once per phase.  Per lane and block this costs
// ~30 instructions per 16 weight bytes instead of ~55 with 8 lanes per block (scale decoding and the float tail are amortised over 4x
// more bytes); the 8 lanes of a shared-memory phase read 8 distinct 16-byte bank groups (q4_K: segments 9b+1+4h+i, q6_K: 13b+4h+i).
struct HFrag { uint4 a[8]; uint4 bs; float d; };

__device__ __forceinline__ void hfrag_fill(const ActS & A, int nblk, HFrag & f) {
    const int lane = threadIdx.x & 31, b = lane >> 1, h = lane & 1;
    if (b < nblk) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f.a[i] = lds16(A.qs + (uint32_t) act_qs_off<true>(b * 256 + 128 * h + 16 * i));
        f.bs = lds16(A.bsum + (b * 16 + 8 * h) * 2);
        f.d = lds_f32(A.d + b * 4);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f.a[i] = make_uint4(0, 0, 0, 0);
        f.bs = make_uint4(0, 0, 0, 0); f.d = 0.0f;
    }
}

__device__ __forceinline__ int dp2a_lo(int a, int b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dot16u(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, const uint4 & a, int acc) {
    acc = __dp4a((int) w0, (int) a.x, acc); acc = __dp4a((int) w1, (int) a.y, acc); acc = __dp4a((int) w2, (int) a.z, acc); return __dp4a((int) w3, (int) a.w, acc);
}

// unsigned-weight x signed-activation dp4a: lets a nibble stay in the HIGH half of its byte (value x16) and a 2-bit field stay where it
// is in qh (value x4^q) — the shifts the unpacking would need move to ONE exact shift of the accumulated sum.  The logic pipe (LOP3/SHF),
// not HBM, is what bounds this kernel (ncu: math_pipe_throttle), so every removed shift counts.
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) { int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dot16m(const uint4 & w, uint32_t mask, const uint4 & a, int acc) {     // sum over 16 bytes of (w & mask) * a
    acc = dp4a_us(w.x & mask, a.x, acc); acc = dp4a_us(w.y & mask, a.y, acc); acc = dp4a_us(w.z & mask, a.z, acc); return dp4a_us(w.w & mask, a.w, acc);
}

template <int T> __device__ __forceinline__ float hblock_dot(uint32_t pb, uint32_t dp, int h, const HFrag & f);

template <> __device__ __forceinline__ float hblock_dot<B200_Q4_K>(uint32_t pb, uint32_t, int h, const HFrag & f) {
    const uint4 hdr = lds16(pb);
    uint4 q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = lds16(pb + 16 + 64 * h + 16 * i);
    // scales / mins of sub-blocks 4h .. 4h+3 as packed bytes (get_scale_min_k4, ggml-quants.c:703-711)
    const uint32_t scw = h ? ((hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4)) : (hdr.y & 0x3f3f3f3fu);
    const uint32_t mnw = h ? (((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4)) : (hdr.z & 0x3f3f3f3fu);
    int isum = 0;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {                      // pair jj: low nibbles = sub-block 2jj (acts a[4jj], a[4jj+1]), high = 2jj+1 (a[4jj+2], a[4jj+3])
        int dlo  = dot16m(q[2 * jj], 0x0f0f0f0fu, f.a[4 * jj], 0);
        dlo      = dot16m(q[2 * jj + 1], 0x0f0f0f0fu, f.a[4 * jj + 1], dlo);
        int dh16 = dot16m(q[2 * jj], 0xf0f0f0f0u, f.a[4 * jj + 2], 0);           // 16 x the high-nibble dot product
        dh16     = dot16m(q[2 * jj + 1], 0xf0f0f0f0u, f.a[4 * jj + 3], dh16);
        isum += (int) ((scw >> (16 * jj)) & 0xff) * dlo + (int) ((scw >> (16 * jj + 8)) & 0xff) * (dh16 >> 4);
    }
    // mins: sub-block s' of this half sums the 16-element groups 2s', 2s'+1 -> dp2a of the packed s16 pair with the min byte duplicated
    int msum = dp2a_lo((int) f.bs.x, (int) __byte_perm(mnw, 0, 0x4400), 0);
    msum = dp2a_lo((int) f.bs.y, (int) __byte_perm(mnw, 0, 0x4411), msum);
    msum = dp2a_lo((int) f.bs.z, (int) __byte_perm(mnw, 0, 0x4422), msum);
    msum = dp2a_lo((int) f.bs.w, (int) __byte_perm(mnw, 0, 0x4433), msum);
    const float dw = h2f(hdr.x & 0xffff), dmin = h2f(hdr.x >> 16);
    return (dw * f.d) * (float) isum - (dmin * f.d) * (float) msum;
}

template <> __device__ __forceinline__ float hblock_dot<B200_Q6_K>(uint32_t pb, uint32_t dp, int h, const HFrag & f) {
    uint4 l[4], hq[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = lds16(pb + 64 * h + 16 * i);          // ql[64h ..+64): bytes 0..31 -> quads 0 (lo) / 2 (hi), 32..63 -> quads 1 / 3
#pragma unroll
    for (int i = 0; i < 2; ++i) hq[i] = lds16(pb + 128 + 32 * h + 16 * i);   // qh[32h ..+32): 2 bits per quad for position l
    const uint32_t sc0 = lds4(pb + 192 + 8 * h), sc1 = lds4(pb + 196 + 8 * h);   // int8 scales sc[8h + 2q + s]
    const float dw = h2f(lds_u16(dp));
    const int16_t * bs = (const int16_t *) &f.bs;
    int isum = 0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {                         // s: positions l = 16s .. 16s+15 of each quad; activations of quad q: f.a[2q + s]
        // w = nibble + 16 * (2-bit field): by linearity dot(w, a) = dot(nibble, a) + 16 * dot(field, a); fields stay in place (x 4^q)
        const int n0 = dot16m(l[s],     0x0f0f0f0fu, f.a[s],     0);          // quad 0: low nibbles of ql[.. 0..31]
        const int n1 = dot16m(l[2 + s], 0x0f0f0f0fu, f.a[2 + s], 0);          // quad 1: low nibbles of ql[.. 32..63]
        const int n2 = dot16m(l[s],     0xf0f0f0f0u, f.a[4 + s], 0);          // quad 2: high nibbles, x16
        const int n3 = dot16m(l[2 + s], 0xf0f0f0f0u, f.a[6 + s], 0);          // quad 3: high nibbles, x16
        const int h0 = dot16m(hq[s], 0x03030303u, f.a[s],     0);             // x1
        const int h1 = dot16m(hq[s], 0x0c0c0c0cu, f.a[2 + s], 0);             // x4
        const int h2 = dot16m(hq[s], 0x30303030u, f.a[4 + s], 0);             // x16
        const int h3 = dot16m(hq[s], 0xc0c0c0c0u, f.a[6 + s], 0);             // x64
        const int d0 = n0 + 16 * h0, d1 = n1 + 4 * h1, d2 = (n2 >> 4) + h2, d3 = (n3 >> 4) + (h3 >> 2);      // all exact
        // quad q, half s -> 16-element group 2q + s of this block half; scale byte index 2q + s (sc0: q = 0, 1; sc1: q = 2, 3)
        isum += (int) (int8_t) (sc0 >> (8 * s))      * (d0 - 32 * (int) bs[s]);
        isum += (int) (int8_t) (sc0 >> (8 * s + 16)) * (d1 - 32 * (int) bs[2 + s]);
        isum += (int) (int8_t) (sc1 >> (8 * s))      * (d2 - 32 * (int) bs[4 + s]);
        isum += (int) (int8_t) (sc1 >> (8 * s + 16)) * (d3 - 32 * (int) bs[6 + s]);
    }
    return (dw * f.d) * (float) isum;
}

// one or two rows (cnt) of a unit, whole row in one pass; two rows interleave for ILP
template <int T>
__device__ __forceinline__ float2 krow_regs(uint32_t prow, uint32_t pstride, uint32_t drow, uint32_t dstride, int nblk, int cnt, const HFrag & f) {
    constexpr int PB = T == B200_Q4_K ? 144 : 208;
    const int lane = threadIdx.x & 31, b = lane >> 1, h = lane & 1;
    float a0 = 0.0f, a1 = 0.0f;
    if (b < nblk) {
        a0 = hblock_dot<T>(prow + b * PB, drow + b * 2, h, f);
        if (cnt == 2) a1 = hblock_dot<T>(prow + pstride + b * PB, drow + dstride + b * 2, h, f);
    }
    a0 = warp_sum(a0);
    if (cnt == 2) a1 = warp_sum(a1);
    return make_float2(a0, a1);
}

// NR rows (same matrix, consecutive in the stage), fragments re-read from shared memory once per block and shared by the rows
template <int T, int NR>
__device__ __forceinline__ float2 krow_lds(uint32_t prow, uint32_t pstride, uint32_t drow, uint32_t dstride, const ActS & A, int nblk) {
    constexpr int PB = T == B200_Q4_K ? 144 : T == B200_Q5_K ? 176 : 208;
    const int lane = threadIdx.x & 31, c = lane & 7;
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 2
    for (int b = lane >> 3; b < nblk; b += 4) {
        const KFrag f = kfrag_load<T>(A, b, c);
        a0 += kblock_dot<T>(prow + b * PB, drow + b * 2, c, f);
        if (NR == 2) a1 += kblock_dot<T>(prow + pstride + b * PB, drow + dstride + b * 2, c, f);
    }
    float2 r; r.x = warp_sum(a0); r.y = NR == 2 ? warp_sum(a1) : 0.0f;
    return r;
}

// legacy 32-element blocks (planar payload + f16 d plane), one lane per block
template <int T> __device__ __forceinline__ float legacy_row_dot(uint32_t prow, uint32_t drow, const ActS & A, int k) {
    const int lane = threadIdx.x & 31;
    const int nblk = k >> 5;
    float acc = 0.0f;
    for (int b = lane; b < nblk; b += 32) {
        const uint4 a0 = lds16(A.qs + (uint32_t) act_qs_off<true>(b * 32)), a1 = lds16(A.qs + (uint32_t) act_qs_off<true>(b * 32 + 16));
        const float dw = h2f(lds_u16(drow + b * 2)), da = h2f(lds_u16(A.d + b * 2));
        if (T == B200_Q8_0) {
            const uint4 q0 = lds16(prow + b * 32), q1 = lds16(prow + b * 32 + 16);
            const uint32_t w0[4] = { q0.x, q0.y, q0.z, q0.w }, w1[4] = { q1.x, q1.y, q1.z, q1.w };
            const int d = dot16(w0, a0) + dot16(w1, a1);
            acc += (float) d * (dw * da);
        } else {
            const uint4 qs = lds16(prow + b * 16);
            const uint32_t wl[4] = { qs.x & 0x0f0f0f0fu, qs.y & 0x0f0f0f0fu, qs.z & 0x0f0f0f0fu, qs.w & 0x0f0f0f0fu };
            const uint32_t wh[4] = { (qs.x >> 4) & 0x0f0f0f0fu, (qs.y >> 4) & 0x0f0f0f0fu, (qs.z >> 4) & 0x0f0f0f0fu, (qs.w >> 4) & 0x0f0f0f0fu };
            const int d = dot16(wl, a0) + dot16(wh, a1) - 8 * lds_s16(A.bsum + b * 2);
            acc += ((float) d * dw) * da;
        }
    }
    return warp_sum(acc);
}

// `cnt` (1 or 2) consecutive rows of one matrix from shared-memory fragments.  The two-row q4_K / q6_K variants (ffn_down, k = 12288)
// are inlined; the rest goes through one out-of-line function to keep the kernel's code small.
__device__ __noinline__ float2 unit_dots_cold(int type, int cnt, uint32_t prow, uint32_t pstride, uint32_t drow, uint32_t dstride,
                                              uint32_t a_qs, uint32_t a_d, uint32_t a_bsum, int k) {
    const int nblk = k >> 8;
    ActS A; A.qs = a_qs; A.d = a_d; A.bsum = a_bsum;
    switch (type) {
        case B200_Q4_K: return krow_lds<B200_Q4_K, 1>(prow, pstride, drow, dstride, A, nblk);
        case B200_Q6_K: return krow_lds<B200_Q6_K, 1>(prow, pstride, drow, dstride, A, nblk);
        case B200_Q5_K: return cnt == 2 ? krow_lds<B200_Q5_K, 2>(prow, pstride, drow, dstride, A, nblk) : krow_lds<B200_Q5_K, 1>(prow, pstride, drow, dstride, A, nblk);
        case B200_Q8_0: return make_float2(legacy_row_dot<B200_Q8_0>(prow, drow, A, k), 0.0f);
        default:        return make_float2(legacy_row_dot<B200_Q4_0>(prow, drow, A, k), 0.0f);
    }
}

__device__ __forceinline__ float2 unit_dots_lds(int type, int cnt, uint32_t prow, uint32_t pstride, uint32_t drow, uint32_t dstride,
                                                const ActS & A, int k) {
    if (cnt == 2 && type == B200_Q4_K) return krow_lds<B200_Q4_K, 2>(prow, pstride, drow, dstride, A, k >> 8);
    if (cnt == 2 && type == B200_Q6_K) return krow_lds<B200_Q6_K, 2>(prow, pstride, drow, dstride, A, k >> 8);
    return unit_dots_cold(type, cnt, prow, pstride, drow, dstride, A.qs, A.d, A.bsum, k);
}
