// stream_attn.cuh — attention phase of the persistent decode kernel.  Included by stream_decode.cu INSIDE namespace b200.
//
// One CTA per (kv head, KV chunk): q-norm + RoPE of the 4 (GQA) query heads, k-norm + RoPE + cache write of the new K/V row (by the CTA
// whose chunk holds the slot), split-KV online-softmax attention over the F16 cache — arithmetic of flash_attn.cu / fused_decode.cu (and so of
// the CPU oracle ops.cpp:7912-8148 up to f32-vs-f16 V accumulation and the order of the online-softmax merges) — then the last CTA of each kv
// head to arrive merges the chunk partials.  Replaces rms_norm_f32 x2, rope_neox x2, k_set_rows x2, flash_attn_ext_vec and
// flash_attn_combine_results of the reference (norm.cu:107-185, rope.cu:83-123, set-rows.cu:264, fattn-vec.cuh:19, fattn-common.cuh).
//
// Latency design: the phase is a chain of dependent round trips, not bandwidth, so (1) the chunk's K/V rows are pulled into L2 during the
// PREVIOUS (qkv matvec) phase — old cache rows do not depend on this token (sa_prefetch_kv); (2) the K, V and mask loads of a 144-position
// tile are all issued together, before the q-norm/RoPE work, into registers; (3) each 16-lane row group keeps its own online-softmax state
// (m, l, acc), so the tile loop has no CTA barrier; the 24 groups are merged once at the end.
constexpr int SA_G    = 4;              // query heads per kv head handled together (GQA ratio must be a multiple; Qwen3: 32/8)
constexpr int SA_U    = 6;              // KV rows per row group and tile
constexpr int SA_NRG  = SD_WARPS * 2;   // row groups (16 lanes each: 8 head dims per lane)
constexpr int SA_TILE = SA_NRG * SA_U;  // KV positions per tile

struct SaSmem {                         // carved from the phase scratch (SD_ATTN_BYTES)
    float q[SA_G][128];
    float red[SD_WARPS][SA_G * 128];
    float2 red_ml[SD_WARPS][SA_G];
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

__device__ void sd_attention(const SdPhase & P, const SdRuntime & rt, uint8_t * scratch, const float2 * rope_tab, unsigned long long * pf) {
    const SdAttn & A = P.attn;
    SaSmem & sm = *(SaSmem *) scratch;
    constexpr int D = 128, LPR = 16, U = SA_U;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp * 2 + lane / LPR, hl = lane % LPR;
    int kvh, split, splits, c0, c1;
    if (!sa_geometry(A, rt.n_kv, kvh, split, splits, c0, c1)) return;        // idle CTA: straight to the grid barrier
    const int ratio = A.n_head / A.n_head_kv, head0 = kvh * ratio;           // ratio == SA_G (checked on the host)
    const int64_t slot = rt.kv_idx[0];
    const bool owner = slot >= c0 && slot < c1;
    const char * kb = (const char *) A.k_cache + (int64_t) kvh * D * 2 + hl * 16;
    const char * vb = (const char *) A.v_cache + (int64_t) kvh * D * 2 + hl * 16;
    const __half * mrow = rt.mask;

    // ---- the first tile's K, V and mask loads go out before anything else ---------------------------------------------------------------
    uint4 kk[U], vv[U]; float mv[U];
    auto load_tile = [&](int t0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pos = t0 + rg + u * SA_NRG;
            const bool in = pos < c1;
            mv[u] = in ? __half2float(mrow[pos]) : -INFINITY;
            kk[u] = in ? ldg_stream16(kb + (int64_t) pos * A.k_row_bytes) : make_uint4(0, 0, 0, 0);
            vv[u] = in ? ldg_stream16(vb + (int64_t) pos * A.v_row_bytes) : make_uint4(0, 0, 0, 0);
        }
    };
    load_tile(c0);

    // ---- q heads (warps 0..3), new K row (warp 4), new V row (warp 5): their loads (and the norm weights') join the same round trip ------
    {
        const bool isq = warp < SA_G, isk = warp == SA_G && owner, isv = warp == SA_G + 1 && owner;
        const float * srcp = isq ? A.q + (head0 + warp) * D : isk ? A.k_new + kvh * D : A.v_new + kvh * D;
        const float * nw = isq ? A.q_norm_w : isk ? A.k_norm_w : nullptr;
        float4 r = make_float4(0, 0, 0, 0), wn = make_float4(1, 1, 1, 1);
        if (isq || isk || isv) r = __ldcg((const float4 *) (srcp + lane * 4));
        if (nw) wn = __ldg((const float4 *) (nw + lane * 4));
        float v[4] = { r.x, r.y, r.z, r.w };
        const float w4[4] = { wn.x, wn.y, wn.z, wn.w };
        if (isq || isk) sa_norm_rope(v, nw ? w4 : nullptr, A.eps, rt.rope_mode, rope_tab);
        if (isq) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sm.q[warp][lane * 4 + i] = __half2float(__float2half_rn(v[i]));       // the oracle rounds Q to f16
        } else if (isk || isv) {
            __half * dst = (__half *) ((isk ? A.k_cache + slot * A.k_row_bytes : A.v_cache + slot * A.v_row_bytes)) + kvh * D + lane * 4;
            __half * snew = isk ? sm.knew : sm.vnew;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const __half hv = __float2half_rn(v[i]); dst[i] = hv; snew[lane * 4 + i] = hv; }
        }
    }
    cons_sync();
    if (pf) pf[4] = globaltimer();

    float acc[SA_G][8], m[SA_G], l[SA_G];
#pragma unroll
    for (int g = 0; g < SA_G; ++g) {
        m[g] = -INFINITY; l[g] = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[g][i] = 0.0f;
    }

    for (int t0 = c0; ; ) {
        // ---- s = scale * K.q + mask for the group's U rows -------------------------------------------------------------------------------
        float s[U][SA_G];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pos = t0 + rg + u * SA_NRG;
            if (pos == slot) { kk[u] = *(const uint4 *) (sm.knew + hl * 8); vv[u] = *(const uint4 *) (sm.vnew + hl * 8); }   // this token's row: not in the cache yet for the loads above
            const __half2 * h2 = (const __half2 *) &kk[u];
            float kf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h2[i]); kf[2 * i] = f.x; kf[2 * i + 1] = f.y; }
#pragma unroll
            for (int g = 0; g < SA_G; ++g) {
                const float4 q0 = *(const float4 *) &sm.q[g][hl * 8], q1 = *(const float4 *) &sm.q[g][hl * 8 + 4];   // re-read per row: keeps 32 registers free
                float d = kf[0] * q0.x;
                d = fmaf(kf[1], q0.y, d); d = fmaf(kf[2], q0.z, d); d = fmaf(kf[3], q0.w, d);
                d = fmaf(kf[4], q1.x, d); d = fmaf(kf[5], q1.y, d); d = fmaf(kf[6], q1.z, d); d = fmaf(kf[7], q1.w, d);
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                s[u][g] = mv[u] != -INFINITY ? d * A.scale + mv[u] : -INFINITY;
            }
        }
        // ---- the group's online softmax + acc = acc * corr + P.V -------------------------------------------------------------------------
        float pr[U][SA_G];
#pragma unroll
        for (int g = 0; g < SA_G; ++g) {
            float mx = s[0][g];
#pragma unroll
            for (int u = 1; u < U; ++u) mx = fmaxf(mx, s[u][g]);
            const float m_new = fmaxf(m[g], mx);
            const float corr = m_new == -INFINITY ? 1.0f : __expf(m[g] - m_new);
            float sum = 0.0f;
#pragma unroll
            for (int u = 0; u < U; ++u) { pr[u][g] = s[u][g] == -INFINITY ? 0.0f : __expf(s[u][g] - m_new); sum += pr[u][g]; }
            m[g] = m_new; l[g] = l[g] * corr + sum;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[g][i] *= corr;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool any = pr[u][0] != 0.0f || pr[u][1] != 0.0f || pr[u][2] != 0.0f || pr[u][3] != 0.0f;
            if (!any) vv[u] = make_uint4(0, 0, 0, 0);                          // masked rows may hold anything (uninitialised cache)
            const __half2 * h2 = (const __half2 *) &vv[u];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h2[i]);
#pragma unroll
                for (int g = 0; g < SA_G; ++g) { acc[g][2 * i] = fmaf(pr[u][g], f.x, acc[g][2 * i]); acc[g][2 * i + 1] = fmaf(pr[u][g], f.y, acc[g][2 * i + 1]); }
            }
        }
        t0 += SA_TILE;
        if (t0 >= c1) break;
        load_tile(t0);
    }
    if (pf) pf[5] = globaltimer();
    // ---- merge the 24 row groups: the two groups of a warp by shuffle, the 12 warps through shared memory (16-byte stores and loads) ---------
#pragma unroll
    for (int g = 0; g < SA_G; ++g) {
        const float mo = __shfl_xor_sync(0xffffffffu, m[g], 16), lo = __shfl_xor_sync(0xffffffffu, l[g], 16);
        const float M = fmaxf(m[g], mo);
        const float ws = m[g] == -INFINITY ? 0.0f : __expf(m[g] - M), wo = mo == -INFINITY ? 0.0f : __expf(mo - M);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[g][i] * ws + __shfl_xor_sync(0xffffffffu, acc[g][i], 16) * wo;
        if (lane < 16) {
            *(float4 *) &sm.red[warp][g * D + hl * 8]     = make_float4(v[0], v[1], v[2], v[3]);
            *(float4 *) &sm.red[warp][g * D + hl * 8 + 4] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (lane == 0) sm.red_ml[warp][g] = make_float2(M, l[g] * ws + lo * wo);
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
            const float inv = L == 0.0f ? 0.0f : 1.0f / L;
            *(float4 *) &A.out[head * D + d] = L == 0.0f ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : make_float4(v.x / L, v.y / L, v.z / L, v.w / L);
            (void) inv;
        } else {
            const int64_t ps = (int64_t) head * splits + split;
            *(float4 *) &A.part_acc[ps * D + d] = v;
            if (d == 0) A.part_ml[ps] = make_float2(M, L);
        }
    }
    if (pf) pf[6] = globaltimer();
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
    if (pf) pf[7] = globaltimer();
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
        *(float4 *) &A.out[head * D + d] = L == 0.0f ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : make_float4(v.x / L, v.y / L, v.z / L, v.w / L);
    }
}
