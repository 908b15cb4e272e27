// stream_attn.cuh — attention phase of the persistent decode kernel.  Included by stream_decode.cu INSIDE namespace b200.
//
// One CTA per (kv head, KV chunk): q-norm + RoPE of the 4 (GQA) query heads, k-norm + RoPE + cache write of the new K/V row (by the CTA
// whose chunk holds the slot), split-KV online-softmax attention over the F16 cache — arithmetic identical to flash_attn.cu /
// fused_decode.cu (and so to the CPU oracle ops.cpp:7912-8148 up to f32-vs-f16 V accumulation) — then the last CTA of each kv head to
// arrive merges the chunk partials.  Replaces rms_norm_f32 x2, rope_neox x2, k_set_rows x2, flash_attn_ext_vec and
// flash_attn_combine_results of the reference (norm.cu:107-185, rope.cu:83-123, set-rows.cu:264, fattn-vec.cuh:19, fattn-common.cuh).
constexpr int SA_TILE = 384;            // KV positions per softmax tile = 24 row groups x 16
constexpr int SA_G    = 4;              // query heads per kv head handled together (GQA ratio must be a multiple; Qwen3: 32/8)

struct SaSmem {                         // carved from the phase scratch (SD_ATTN_BYTES)
    float S[SA_TILE][SA_G];
    float q[SA_G][128];
    float red[SD_WARPS][SA_G * 128 + 4];
    __half knew[128], vnew[128];
    float corr[SA_G], m[SA_G], l[SA_G];
    int is_last;
};
static_assert(sizeof(SaSmem) <= SD_ATTN_BYTES, "attention scratch");

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

// warp-level: RMS-norm (optional) + weight + rotary embedding of one 128-wide head held as 4 contiguous elements per lane
__device__ __forceinline__ void sa_norm_rope(float (&v)[4], const float * w, float eps, int rope_mode, const float2 * tab) {
    const int lane = threadIdx.x & 31;
    if (w) {
        float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
        ss = warp_sum(ss);
        const float scale = 1.0f / sqrtf(ss / 128.0f + eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __fmul_rn(__fmul_rn(v[i], scale), w[lane * 4 + i]);
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
    constexpr int D = 128, LPR = 16, NRG = SD_WARPS * 2, U = 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp * 2 + lane / LPR, hl = lane % LPR;
    const int n_kv = rt.n_kv;
    int splits = (int) gridDim.x / A.n_head_kv;
    int chunk = (n_kv + splits - 1) / splits; chunk = (chunk + 31) / 32 * 32;
    splits = (n_kv + chunk - 1) / chunk;
    if ((int) blockIdx.x >= A.n_head_kv * splits) return;                   // idle CTA: straight to the grid barrier
    const int kvh = blockIdx.x / splits, split = blockIdx.x % splits;
    const int ratio = A.n_head / A.n_head_kv, head0 = kvh * ratio;           // ratio == SA_G (checked on the host)
    const int c0 = split * chunk, c1 = min(c0 + chunk, n_kv);
    const int64_t slot = rt.kv_idx[0];
    const bool owner = slot >= c0 && slot < c1;

    // ---- q heads (warps 0..3), new K row (warp 4), new V row (warp 5) ---------------------------------------------------------------
    if (warp < SA_G) {
        const float4 r = __ldcg((const float4 *) (A.q + (head0 + warp) * D + lane * 4));
        float v[4] = { r.x, r.y, r.z, r.w };
        sa_norm_rope(v, A.q_norm_w, A.eps, rt.rope_mode, rope_tab);
#pragma unroll
        for (int i = 0; i < 4; ++i) sm.q[warp][lane * 4 + i] = __half2float(__float2half_rn(v[i]));       // the oracle rounds Q to f16
        if (lane == 0) { sm.m[warp] = -INFINITY; sm.l[warp] = 0.0f; }
    } else if (warp == SA_G && owner) {
        const float4 r = __ldcg((const float4 *) (A.k_new + kvh * D + lane * 4));
        float v[4] = { r.x, r.y, r.z, r.w };
        sa_norm_rope(v, A.k_norm_w, A.eps, rt.rope_mode, rope_tab);
        __half * dst = (__half *) (A.k_cache + slot * A.k_row_bytes) + kvh * D + lane * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const __half hv = __float2half_rn(v[i]); dst[i] = hv; sm.knew[lane * 4 + i] = hv; }
    } else if (warp == SA_G + 1 && owner) {
        const float4 r = __ldcg((const float4 *) (A.v_new + kvh * D + lane * 4));
        const float v[4] = { r.x, r.y, r.z, r.w };
        __half * dst = (__half *) (A.v_cache + slot * A.v_row_bytes) + kvh * D + lane * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const __half hv = __float2half_rn(v[i]); dst[i] = hv; sm.vnew[lane * 4 + i] = hv; }
    }
    __syncthreads();
    if (pf) pf[4] = globaltimer();

    float qreg[SA_G][8], acc[SA_G][8];
#pragma unroll
    for (int g = 0; g < SA_G; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) { qreg[g][i] = sm.q[g][hl * 8 + i]; acc[g][i] = 0.0f; }
    const char * kb = (const char *) A.k_cache + (int64_t) kvh * D * 2 + hl * 16;
    const char * vb = (const char *) A.v_cache + (int64_t) kvh * D * 2 + hl * 16;
    const __half * mrow = rt.mask;

    for (int t0 = c0; t0 < c1; t0 += SA_TILE) {
        // ---- S = scale * K.q + mask ---------------------------------------------------------------------------------------------------
        for (int j0 = rg; j0 < SA_TILE; j0 += NRG * U) {
            uint4 kk[U]; float mv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int pos = t0 + j0 + u * NRG;
                mv[u] = pos < c1 ? __half2float(mrow[pos]) : -INFINITY;
                kk[u] = make_uint4(0, 0, 0, 0);
                if (mv[u] != -INFINITY) kk[u] = pos == slot ? *(const uint4 *) (sm.knew + hl * 8) : ldg_stream16(kb + (int64_t) pos * A.k_row_bytes);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const __half2 * h2 = (const __half2 *) &kk[u];
                float kf[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h2[i]); kf[2 * i] = f.x; kf[2 * i + 1] = f.y; }
#pragma unroll
                for (int g = 0; g < SA_G; ++g) {
                    float d = 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) d = fmaf(kf[i], qreg[g][i], d);
#pragma unroll
                    for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    if (hl == 0) sm.S[j0 + u * NRG][g] = mv[u] != -INFINITY ? d * A.scale + mv[u] : -INFINITY;
                }
            }
        }
        __syncthreads();
        // ---- online softmax bookkeeping: warp g owns head g -----------------------------------------------------------------------------
        if (warp < SA_G) {
            const int g = warp;
            float mx = -INFINITY;
            for (int j = lane; j < SA_TILE; j += 32) mx = fmaxf(mx, sm.S[j][g]);
            mx = warp_max(mx);
            const float m_old = sm.m[g], m_new = fmaxf(m_old, mx);
            float sum = 0.0f;
            for (int j = lane; j < SA_TILE; j += 32) {
                const float s = sm.S[j][g];
                const float p = s == -INFINITY ? 0.0f : expf(s - m_new);
                sm.S[j][g] = p; sum += p;
            }
            sum = warp_sum(sum);
            if (lane == 0) {
                const float corr = m_old == -INFINITY ? 1.0f : expf(m_old - m_new);
                sm.corr[g] = corr; sm.m[g] = m_new; sm.l[g] = sm.l[g] * corr + sum;
            }
        }
        __syncthreads();
        // ---- acc = acc * corr + P.V -------------------------------------------------------------------------------------------------------
#pragma unroll
        for (int g = 0; g < SA_G; ++g) { const float c = sm.corr[g];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[g][i] *= c; }
        for (int j0 = rg; j0 < SA_TILE; j0 += NRG * U) {
            uint4 vv[U]; float4 pv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * NRG, pos = t0 + j;
                pv[u] = *(const float4 *) &sm.S[j][0];
                const bool any = pv[u].x != 0.0f || pv[u].y != 0.0f || pv[u].z != 0.0f || pv[u].w != 0.0f;
                vv[u] = make_uint4(0, 0, 0, 0);
                if (any) vv[u] = pos == slot ? *(const uint4 *) (sm.vnew + hl * 8) : ldg_stream16(vb + (int64_t) pos * A.v_row_bytes);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const __half2 * h2 = (const __half2 *) &vv[u];
                const float pg[SA_G] = { pv[u].x, pv[u].y, pv[u].z, pv[u].w };
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
#pragma unroll
                    for (int g = 0; g < SA_G; ++g) { acc[g][2 * i] = fmaf(pg[g], f.x, acc[g][2 * i]); acc[g][2 * i + 1] = fmaf(pg[g], f.y, acc[g][2 * i + 1]); }
                }
            }
        }
        __syncthreads();
    }
    if (pf) pf[5] = globaltimer();
    // ---- reduce the 24 row-group accumulators: the two groups of a warp by shuffle, the 12 warps through shared memory --------------------
#pragma unroll
    for (int g = 0; g < SA_G; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = acc[g][i] + __shfl_xor_sync(0xffffffffu, acc[g][i], 16);
            if (lane < 16) sm.red[warp][g * D + hl * 8 + i] = v;
        }
    __syncthreads();
    for (int o = tid; o < SA_G * D; o += SD_THREADS) {
        float v = 0.0f;
#pragma unroll
        for (int r = 0; r < SD_WARPS; ++r) v += sm.red[r][o];
        const int g = o / D, d = o % D, head = head0 + g;
        if (splits == 1) {
            const float l = sm.l[g];
            A.out[head * D + d] = l == 0.0f ? 0.0f : v / l;
        } else {
            const int64_t ps = (int64_t) head * splits + split;
            A.part_acc[ps * D + d] = v;
            if (d == 0) A.part_ml[ps] = make_float2(sm.m[g], sm.l[g]);
        }
    }
    if (pf) pf[6] = globaltimer();
    if (splits == 1) return;
    // ---- the last chunk of this kv head to finish merges the partials ---------------------------------------------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(A.tickets + kvh, 1u);
        sm.is_last = t == (unsigned) splits - 1;
        if (sm.is_last) { A.tickets[kvh] = 0; __threadfence(); }
    }
    __syncthreads();
    if (pf) pf[7] = globaltimer();
    if (!sm.is_last) return;
    // merge weights: warp g (one head) turns the chunks' (m, l) into w_s = exp(m_s - M) / sum_s l_s w_s, lanes = chunks
    float * wgt = &sm.S[0][0];                                              // [SA_G][splits], the score tile is free now
    if (warp < SA_G) {
        const int64_t ps0 = (int64_t) (head0 + warp) * splits;
        float M = -INFINITY;
        for (int s = lane; s < splits; s += 32) M = fmaxf(M, __ldcg(&A.part_ml[ps0 + s]).x);
        M = warp_max(M);
        float L = 0.0f;
        for (int s = lane; s < splits; s += 32) {
            const float2 ml = __ldcg(&A.part_ml[ps0 + s]);
            const float w = ml.x == -INFINITY ? 0.0f : expf(ml.x - M);
            wgt[warp * splits + s] = w; L += ml.y * w;
        }
        L = warp_sum(L);
        __syncwarp();
        const float inv = L == 0.0f ? 0.0f : 1.0f / L;
        for (int s = lane; s < splits; s += 32) wgt[warp * splits + s] *= inv;
    }
    __syncthreads();
    for (int o = tid; o < SA_G * D; o += SD_THREADS) {
        const int g = o / D, d = o % D, head = head0 + g;
        const float * pa = A.part_acc + (int64_t) head * splits * D + d;
        float v = 0.0f;
        int s = 0;
        for (; s + 6 <= splits; s += 6) {                                   // six independent L2 loads in flight per thread
            float t[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) t[i] = __ldcg(pa + (int64_t) (s + i) * D);
#pragma unroll
            for (int i = 0; i < 6; ++i) v = fmaf(t[i], wgt[g * splits + s + i], v);
        }
        for (; s < splits; ++s) v = fmaf(__ldcg(pa + (int64_t) s * D), wgt[g * splits + s], v);
        A.out[head * D + d] = v;
    }
}
