// stream_attn.cuh — attention phase of the persistent decode kernel.  Included by stream_decode.cu INSIDE namespace b200.
//
// One CTA per (kv head, KV chunk): q-norm + RoPE of the 4 (GQA) query heads, k-norm + RoPE + cache write of the new K/V row (by the CTA
// whose chunk holds the slot), split-KV online-softmax attention over the F16 cache — arithmetic of flash_attn.cu / fused_decode.cu (and so of
// the CPU oracle ops.cpp:7912-8148 up to f32-vs-f16 V accumulation and the order of the online-softmax merges) — then the last CTA of each kv
// head to arrive merges the chunk partials.  Replaces rms_norm_f32 x2, rope_neox x2, k_set_rows x2, flash_attn_ext_vec and
// flash_attn_combine_results of the reference (norm.cu:107-185, rope.cu:83-123, set-rows.cu:264, fattn-vec.cuh:19, fattn-common.cuh).
//
// Latency design: the phase is a chain of dependent round trips, not bandwidth, so (1) the chunk's K/V rows are pulled into L2 during the
// PREVIOUS (qkv matvec) phase — old cache rows do not depend on this token (sa_prefetch_kv); (2) every warp owns 16-position tiles and issues
// the K, V and mask loads of its first tile before the q-norm/RoPE work; (3) the tile arithmetic runs on the tensor cores (legacy
// mma.sync.m16n8k16: 4 query heads x 16 positions x 128 dims is far too small for a tcgen05 tile, and at 12 warps per SM the SIMT version was
// bound by its ~1500 instructions per lane): S = Q.K^T with the 4 GQA heads as the M rows (12 of 16 padded) and P.V with P split into f16
// hi + lo parts, so the result is the f32-P x f16-V product of the scalar path to ~1e-7; every warp keeps its own online-softmax state, the
// 12 warps are merged once at the end.  Fragments are loaded straight from global memory with 16-byte loads: a dot product is invariant under
// a permutation of its reduction index applied to both operands, so a lane's 8 consecutive halves serve as two k-steps of its fragment.
constexpr int SA_G    = 4;              // query heads per kv head handled together (GQA ratio must be a multiple; Qwen3: 32/8)
constexpr int SA_TILE = 16;             // KV positions per warp tile (two m16n8k16 column tiles)

struct SaSmem {                         // carved from the phase scratch (SD_ATTN_BYTES)
    float red[SD_WARPS][SA_G * 128];    // per-warp un-normalised outputs, merged by the CTA
    float2 red_ml[SD_WARPS][SA_G];
    __half q16[SA_G][128];              // normalised + rotated queries, rounded to f16 exactly as the oracle does before the dot products
    __half knew[128], vnew[128];
    int is_last;
};
static_assert(sizeof(SaSmem) <= SD_ATTN_BYTES, "attention scratch");

// which (kv head, chunk) this CTA owns; false = idle CTA
__device__ __forceinline__ bool sa_geometry(const SdAttn & A, int n_kv, int & kvh, int & split, int & splits, int & c0, int & c1) {
    splits = (int) gridDim.x / A.n_head_kv;
    int chunk = (n_kv + splits - 1) / splits; chunk = (chunk + 31) / 32 * 32;
    splits = (n_kv + chunk - 1) / chunk;
    if ((int) blockIdx.x >= A.n_head_kv * splits) return false;
    kvh = blockIdx.x / splits; split = blockIdx.x % splits;
    c0 = split * chunk; c1 = min(c0 + chunk, n_kv);
    return true;
}

// issued by the producer warp when it stages the phase (two phases early): L2 prefetch of this CTA's chunk of the layer's K and V cache
__device__ __forceinline__ void sa_prefetch_kv(const SdAttn & A, const SdRuntime & rt, int tid, int nthreads) {
    int kvh, split, splits, c0, c1;
    if (!sa_geometry(A, rt.n_kv, kvh, split, splits, c0, c1)) return;
    const int rows = c1 - c0;
    for (int r = tid; r < rows * 4; r += nthreads) {           // 256 B of K and of V per position = 4 lines
        const int pos = c0 + (r >> 2), which = (r >> 1) & 1, half = r & 1;
        const uint8_t * a = (which ? A.v_cache + (int64_t) pos * A.v_row_bytes : A.k_cache + (int64_t) pos * A.k_row_bytes) + kvh * 256 + half * 128;
        asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
    }
}

// (cos, sin) * mscale of the token's position for the 64 rotation pairs: computed ONCE per launch (every layer rotates by the same
// angles), with the oracle's theta chain (theta *= theta_scale per pair, ops.cpp ggml_rope_cache_init) and the accurate sincosf
__device__ __forceinline__ void sa_rope_table(float2 * tab, const SdRuntime & rt, int lane) {
    const float posf = (float) rt.pos[0];
    for (int p = lane; p < 64; p += 32) {
        float theta = posf;
        for (int j = 0; j < p; ++j) theta = __fmul_rn(theta, rt.theta_scale);
        float th = __fmul_rn(rt.freq_scale, theta), ms = rt.attn_factor;
        if (rt.ext_factor != 0.0f) {
            const float yv = ((float) p - rt.corr0) / fmaxf(0.001f, rt.corr1 - rt.corr0);
            const float ramp = (1.0f - fminf(1.0f, fmaxf(0.0f, yv))) * rt.ext_factor;
            th = __fadd_rn(__fmul_rn(th, 1.0f - ramp), __fmul_rn(theta, ramp));
            ms *= 1.0f + 0.1f * logf(1.0f / rt.freq_scale);
        }
        float sn, cs; sincosf(th, &sn, &cs);
        tab[p] = make_float2(cs * ms, sn * ms);
    }
}

// warp-level: RMS-norm (optional; w = the lane's 4 norm weights) + rotary embedding of one 128-wide head held as 4 contiguous elements per lane
__device__ __forceinline__ void sa_norm_rope(float (&v)[4], const float * w, float eps, int rope_mode, const float2 * tab) {
    const int lane = threadIdx.x & 31;
    if (w) {
        float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
        ss = warp_sum(ss);
        const float scale = 1.0f / sqrtf(ss / 128.0f + eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __fmul_rn(__fmul_rn(v[i], scale), w[i]);
    }
    float out[4];
    if (rope_mode & 2) {                // neox: pairs (p, p + 64) live in lanes (l, l + 16)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 cs = tab[(lane & 15) * 4 + i];
            const float other = __shfl_xor_sync(0xffffffffu, v[i], 16);
            out[i] = lane < 16 ? v[i] * cs.x - other * cs.y : other * cs.y + v[i] * cs.x;
        }
    } else {                            // norm: pairs (2p, 2p + 1) inside a lane
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
            const float2 cs = tab[(lane * 4 + i) / 2];
            out[i] = v[i] * cs.x - v[i + 1] * cs.y; out[i + 1] = v[i] * cs.y + v[i + 1] * cs.x;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = out[i];
}

// D(16x8, f32) += A(16x16, f16, row) . B(16x8, f16, col): the legacy warp-level tensor-core op (HMMA.16816.F32)
__device__ __forceinline__ void mma_f16_16816(float & d0, float & d1, float & d2, float & d3, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t u4_at(const uint4 & v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// four tagged outputs of one head (16-byte aligned pair address)
__device__ __forceinline__ void sa_store_out4(float * out_pairs, int idx, float4 v, uint32_t tag) {
    st_tagged2(out_pairs + 2 * idx, v.x, v.y, tag); st_tagged2(out_pairs + 2 * idx + 4, v.z, v.w, tag);
}

// Out of line ON PURPOSE: the phase needs ~150 registers of its own (K/V fragments in flight + 32 accumulators); inlined into the phase loop it
// made the matvec consumer loop spill.  Shared-memory pointers are re-derived from the extern symbol (see the map in stream_decode.cu).
__device__ __noinline__ void sd_attention(int staged_slot, int n_kv, const int64_t * kv_idx, const __half * mask, int rope_mode, uint32_t tag_base, uint32_t my_tag,
                                          unsigned long long * pf) {
    extern __shared__ __align__(128) uint8_t smem[];
    const SdAttn & A = ((const StagedPhase *) (smem + SM_PHASES))[staged_slot].P.attn;
    SaSmem & sm = *(SaSmem *) (smem + SM_ATTN);
    const float2 * rope_tab = (const float2 *) (smem + SM_ROPE);
    constexpr int D = 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int kvh, split, splits, c0, c1;
    if (!sa_geometry(A, n_kv, kvh, split, splits, c0, c1)) return;           // idle CTA
    const int ratio = A.n_head / A.n_head_kv, head0 = kvh * ratio;           // ratio == SA_G (checked on the host)
    const int64_t slot = kv_idx[0];
    const bool owner = slot >= c0 && slot < c1;
    const __half * mrow = mask;

    // ---- the first tile's K, V and mask loads go out before anything else ---------------------------------------------------------------
    // lane (g, t) = (lane >> 2, lane & 3).  K: B fragments of the two 8-position column tiles: row c0' + 8 nt + g, halves [32 blk + 8 t, +8).
    // V: the 4 positions {2t, 2t+1, 8+2t, 9+2t} this lane's P fragment multiplies, dims [16 g, 16 g + 16).
    const int g = lane >> 2, t = lane & 3;
    uint4 kf[2][4], vf[4][2]; float mv[2][2]; bool live[2][2];
    auto load_tile = [&](int tile) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int pos = min(tile + 8 * nt + g, c1 - 1);                // out-of-chunk rows are clamped (and masked below)
            const char * kr = (const char *) A.k_cache + (int64_t) pos * A.k_row_bytes + (int64_t) kvh * D * 2 + t * 16;
#pragma unroll
            for (int blk = 0; blk < 4; ++blk) kf[nt][blk] = ldg_stream16(kr + blk * 64);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int pp = tile + 8 * nt + 2 * t + e;
                const float m = pp < c1 ? __half2float(mrow[pp]) : -INFINITY;
                mv[nt][e] = m; live[nt][e] = m != -INFINITY;
                const char * vr = (const char *) A.v_cache + (int64_t) min(pp, c1 - 1) * A.v_row_bytes + (int64_t) kvh * D * 2 + g * 32;
                vf[2 * nt + e][0] = ldg_stream16(vr); vf[2 * nt + e][1] = ldg_stream16(vr + 16);
            }
        }
    };
    const int tile0 = c0 + warp * SA_TILE;
    if (tile0 < c1) load_tile(tile0);

    // ---- q heads (warps 0..3), new K row (warp 4), new V row (warp 5): their loads (and the norm weights') join the same round trip ------
    {
        const bool isq = warp < SA_G, isk = warp == SA_G && owner, isv = warp == SA_G + 1 && owner;
        const float * srcp = isq ? A.q + 2 * (head0 + warp) * D : isk ? A.k_new + 2 * kvh * D : A.v_new + 2 * kvh * D;     // pairs
        const float * nw = isq ? A.q_norm_w : isk ? A.k_norm_w : nullptr;
        float4 r = make_float4(0, 0, 0, 0), wn = make_float4(1, 1, 1, 1);
        if (nw) wn = __ldg((const float4 *) (nw + lane * 4));
        if (isq || isk || isv) {                                            // tagged pairs from the qkv phase: poll until this warp's 128 values have landed
            const uint32_t want = tag_base + (uint32_t) A.in_tag;
            uint32_t spins = 0; long long t0 = 0;
            for (;;) {
                const Pairs4 pr = ld_pairs4(srcp + 2 * (lane * 4));
                const bool ok = pair_tag(pr.p[0]) == want && pair_tag(pr.p[1]) == want && pair_tag(pr.p[2]) == want && pair_tag(pr.p[3]) == want;
                r = make_float4(pair_val(pr.p[0]), pair_val(pr.p[1]), pair_val(pr.p[2]), pair_val(pr.p[3]));
                if (__all_sync(0xffffffffu, ok)) break;
                sd_spin_guard(spins, t0);
            }
        }
        float v[4] = { r.x, r.y, r.z, r.w };
        const float w4[4] = { wn.x, wn.y, wn.z, wn.w };
        if (isq || isk) sa_norm_rope(v, nw ? w4 : nullptr, A.eps, rope_mode, rope_tab);
        if (isq) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sm.q16[warp][lane * 4 + i] = __float2half_rn(v[i]);                    // the oracle rounds Q to f16
        } else if (isk || isv) {
            __half * dst = (__half *) ((isk ? A.k_cache + slot * A.k_row_bytes : A.v_cache + slot * A.v_row_bytes)) + kvh * D + lane * 4;
            __half * snew = isk ? sm.knew : sm.vnew;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const __half hv = __float2half_rn(v[i]); dst[i] = hv; snew[lane * 4 + i] = hv; }
        }
    }
    cons_sync();
    if (pf) pf[4] = globaltimer();

    // A fragments of Q (rows = the 4 heads; rows 4..15 of the tile are zero) are re-read from shared memory per use: 16 registers matter more here
    const __half * qrow = &sm.q16[g & (SA_G - 1)][8 * t];

    float o[16][2], m_run = -INFINITY, l_run = 0.0f, zz0 = 0.0f, zz1 = 0.0f;    // o[j] = out[head g][dims 32 t + j, 32 t + 16 + j]
#pragma unroll
    for (int j = 0; j < 16; ++j) { o[j][0] = 0.0f; o[j][1] = 0.0f; }

    for (int tile = tile0; tile < c1; ) {
        // this token's row is not in the cache yet for the loads above: patch it from shared memory
        if (owner) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                if (tile + 8 * nt + g == slot) {
#pragma unroll
                    for (int blk = 0; blk < 4; ++blk) kf[nt][blk] = *(const uint4 *) (sm.knew + 32 * blk + 8 * t);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) if (tile + 8 * nt + 2 * t + e == slot) {
                    vf[2 * nt + e][0] = *(const uint4 *) (sm.vnew + 16 * g); vf[2 * nt + e][1] = *(const uint4 *) (sm.vnew + 16 * g + 8);
                }
            }
        }
        // ---- S[head][pos] = Q . K^T (f16 x f16 -> f32) ----------------------------------------------------------------------------------
        float sc[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            float c0_ = 0.0f, c1_ = 0.0f;
#pragma unroll
            for (int blk = 0; blk < 4; ++blk) {
                uint4 qa = *(const uint4 *) (qrow + 32 * blk);
                if (g >= SA_G) qa = make_uint4(0, 0, 0, 0);
                mma_f16_16816(c0_, c1_, zz0, zz1, qa.x, 0u, qa.y, 0u, kf[nt][blk].x, kf[nt][blk].y);
                mma_f16_16816(c0_, c1_, zz0, zz1, qa.z, 0u, qa.w, 0u, kf[nt][blk].z, kf[nt][blk].w);
            }
            sc[nt][0] = live[nt][0] ? c0_ * A.scale + mv[nt][0] : -INFINITY;
            sc[nt][1] = live[nt][1] ? c1_ * A.scale + mv[nt][1] : -INFINITY;
        }
        if (pf && tile == tile0) pf[5] = globaltimer();
        // ---- online softmax of row g over the tile's 16 positions (4 per lane, 4 lanes per row) ------------------------------------------
        float mx = fmaxf(fmaxf(sc[0][0], sc[0][1]), fmaxf(sc[1][0], sc[1][1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run, mx);
        const float corr = m_new == -INFINITY ? 1.0f : __expf(m_run - m_new);
        float pr[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) pr[nt][e] = sc[nt][e] == -INFINITY ? 0.0f : __expf(sc[nt][e] - m_new);
        }
        l_run = l_run * corr + (pr[0][0] + pr[0][1]) + (pr[1][0] + pr[1][1]);          // lane-local; summed over the row's 4 lanes at the end
        m_run = m_new;
        if (tile != tile0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { o[j][0] *= corr; o[j][1] *= corr; }
        }
        // P = hi + lo (two f16 parts: ~22 bits) as the A fragments of P.V; masked positions contribute exactly nothing, so their V rows are
        // zeroed (they may hold anything: uninitialised cache cells, NaN included)
        uint32_t ahi[2], alo[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const __half2 hi = __floats2half2_rn(pr[nt][0], pr[nt][1]);
            const float2 hf = __half22float2(hi);
            const __half2 lo = __floats2half2_rn(pr[nt][0] - hf.x, pr[nt][1] - hf.y);
            ahi[nt] = *(const uint32_t *) &hi; alo[nt] = *(const uint32_t *) &lo;
#pragma unroll
            for (int e = 0; e < 2; ++e) if (!live[nt][e]) { vf[2 * nt + e][0] = make_uint4(0, 0, 0, 0); vf[2 * nt + e][1] = make_uint4(0, 0, 0, 0); }
        }
        // ---- out[head][dim] += P . V: column tile j holds dims { 16 c + j } -> B fragment = halves j of this lane's 4 V rows ----------------
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
            const uint32_t b0 = __byte_perm(u4_at(vf[0][j >> 3], (j >> 1) & 3), u4_at(vf[1][j >> 3], (j >> 1) & 3), sel);
            const uint32_t b1 = __byte_perm(u4_at(vf[2][j >> 3], (j >> 1) & 3), u4_at(vf[3][j >> 3], (j >> 1) & 3), sel);
            mma_f16_16816(o[j][0], o[j][1], zz0, zz1, ahi[0], 0u, ahi[1], 0u, b0, b1);
            mma_f16_16816(o[j][0], o[j][1], zz0, zz1, alo[0], 0u, alo[1], 0u, b0, b1);
        }
        tile += SD_WARPS * SA_TILE;
        if (tile < c1) load_tile(tile);
    }
    if (pf) pf[6] = globaltimer();
    // ---- hand this warp's state to the CTA merge: rows g < 4 are real heads -----------------------------------------------------------------
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1); l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    if (g < SA_G) {
        float * dstp = &sm.red[warp][g * D + 32 * t];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            *(float4 *) (dstp + j)      = make_float4(o[j][0], o[j + 1][0], o[j + 2][0], o[j + 3][0]);
            *(float4 *) (dstp + 16 + j) = make_float4(o[j][1], o[j + 1][1], o[j + 2][1], o[j + 3][1]);
        }
        if (t == 0) sm.red_ml[warp][g] = make_float2(m_run, l_run);
    }
    cons_sync();
    if (tid < SA_G * D / 4) {                                               // 128 threads x 4 consecutive outputs of one head
        const int o = tid * 4, g = o / D, d = o % D, head = head0 + g;
        float M = -INFINITY;
#pragma unroll
        for (int r = 0; r < SD_WARPS; ++r) M = fmaxf(M, sm.red_ml[r][g].x);
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f); float L = 0.0f;
#pragma unroll
        for (int r = 0; r < SD_WARPS; ++r) {
            const float2 ml = sm.red_ml[r][g];
            const float w = ml.x == -INFINITY ? 0.0f : __expf(ml.x - M);
            const float4 a = *(const float4 *) &sm.red[r][o];
            v.x = fmaf(a.x, w, v.x); v.y = fmaf(a.y, w, v.y); v.z = fmaf(a.z, w, v.z); v.w = fmaf(a.w, w, v.w); L = fmaf(ml.y, w, L);
        }
        if (splits == 1) {
            sa_store_out4(A.out, head * D + d, L == 0.0f ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : make_float4(v.x / L, v.y / L, v.z / L, v.w / L), tag_base + my_tag);
        } else {
            const int64_t ps = (int64_t) head * splits + split;
            *(float4 *) &A.part_acc[ps * D + d] = v;
            if (d == 0) A.part_ml[ps] = make_float2(M, L);
        }
    }
    if (pf) pf[7] = globaltimer();
    if (splits == 1) return;
    // ---- the last chunk of this kv head to finish merges the partials ---------------------------------------------------------------------
    cons_sync();
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(A.tickets + kvh, 1u);
        sm.is_last = t == (unsigned) splits - 1;
        if (sm.is_last) { A.tickets[kvh] = 0; __threadfence(); }
    }
    cons_sync();
    if (!sm.is_last) return;
    // merge: 128 threads x 4 consecutive output elements; a thread loads its float4 column of partials AND the head's (m, l) pairs in one
    // round trip (18 + 18 requests instead of 36 scalar ones per element: the phase's tail is bound by outstanding requests), then turns them
    // into w_s = exp(m_s - M) / sum_s l_s w_s itself (cheaper than a second round trip through shared memory)
    if (tid < SA_G * D / 4) {
        const int o = tid * 4, g = o / D, d = o % D, head = head0 + g;
        const float * pa = A.part_acc + (int64_t) head * splits * D + d;
        const float2 * pm = A.part_ml + (int64_t) head * splits;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f); float L = 0.0f, M = -INFINITY;
        constexpr int NP = 18;                                              // chunks per pass = 148 SMs / 8 kv heads: the whole merge is ONE round trip
        for (int s0 = 0; s0 < splits; s0 += NP) {
            float4 t[NP]; float2 ml[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const bool in = s0 + i < splits;
                t[i] = in ? __ldcg((const float4 *) (pa + (int64_t) (s0 + i) * D)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                ml[i] = in ? __ldcg(pm + s0 + i) : make_float2(-INFINITY, 0.0f);
            }
            float Mn = M;
#pragma unroll
            for (int i = 0; i < NP; ++i) Mn = fmaxf(Mn, ml[i].x);
            const float c = Mn == -INFINITY ? 1.0f : __expf(M - Mn);          // M == -inf: v = L = 0 anyway
            v.x *= c; v.y *= c; v.z *= c; v.w *= c; L *= c; M = Mn;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const float w = ml[i].x == -INFINITY ? 0.0f : __expf(ml[i].x - M);
                v.x = fmaf(t[i].x, w, v.x); v.y = fmaf(t[i].y, w, v.y); v.z = fmaf(t[i].z, w, v.z); v.w = fmaf(t[i].w, w, v.w); L = fmaf(ml[i].y, w, L);
            }
        }
        sa_store_out4(A.out, head * D + d, L == 0.0f ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : make_float4(v.x / L, v.y / L, v.z / L, v.w / L), tag_base + my_tag);
    }
}
