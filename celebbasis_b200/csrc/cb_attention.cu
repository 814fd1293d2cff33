// cb_attention_fwd — fused scaled-dot-product attention forward for sm_100a (flash style).
//
// Replaces the materialised `sim = einsum(q,k)*scale; attn = sim.softmax(-1); out = einsum(attn, v)` of
// ldm/modules/attention.py:178-191 (and the CLIP layers' masked variant, encoders/modules.py:24-31): the
// (heads x N x N) score tensor -- 512 MiB fp32 per 4096-token block in the reference -- never reaches HBM unless
// the caller asks for the probabilities (training keeps them for the backward pass).
//
// One CTA = 128 query rows of one (image, head):
//   warp 0     : TMA producer -- Q once, then K_j / V_j blocks of 128 keys through a 2-stage mbarrier ring
//   warp 1     : single-thread tcgen05.mma issuer: S_j = Q K_j^T into one of two TMEM score buffers,
//                O += P_j V_j (V read MN-major straight from its [keys][d] layout) into the TMEM output tile
//   warps 2..5 : one thread per query row: tcgen05.ld of the score row, online softmax (running max / sum in
//                registers, exp2 with the scale folded in), rescale of the TMEM output tile when the max moves,
//                P_j written as the next MMA's K-major SWIZZLE_128B A operand in shared memory;
//                finally O / l -> HBM and logsumexp.
// Head dims 40 / 64 / 80 / 128 (any multiple of 8 up to 128): the tensor maps declare the head's d columns as the
// K extent, so TMA zero-fills the rest of each 64-column box and the MMAs run on 16-column multiples.
#include <string.h>

#include <type_traits>

#include "cb_common.cuh"

namespace cb {

constexpr int kAttnThreads = 192;
constexpr int kBQ = 128;    // query rows per CTA
constexpr int kBKV = 64;    // keys per block (64 scores per softmax thread keeps 2-4 CTAs resident per SM)

struct AttnParams {
    int nq, nk, heads, images;
    int d, dpad16;          // head dim, head dim rounded up to 16
    int dboxes;             // 64-column boxes per row of Q/K/V (1 or 2)
    int causal;
    int two_pass;           // pass A: row max / sum only; pass B: normalised probabilities (needed when P is stored)
    float scale_log2e;      // scale * log2(e)
    void* O;
    int o_dtype;
    long long ldo;          // elements between consecutive query rows of O
    float* lse;             // [images][heads][nq] natural-log sum-exp (scaled scores), or NULL
    void* P;                // optional probabilities [images*heads][nq][ldp] (same 16-bit dtype as the operands)
    long long ldp;
    int p_is_bf16;
    unsigned idesc_s, idesc_o;
    unsigned long long* dbg;      // profiling aid (tools/attn_timeline.py): per-CTA accumulated cycles of the softmax phases
};

template <int DBOX>  // number of 64-column boxes of the head dim (1: d <= 64, 2: d <= 128)
struct AttnCfg {
    static constexpr int kQBytes = DBOX * kBQ * 128;
    static constexpr int kKBytes = DBOX * kBKV * 128;
    static constexpr int kVBytes = DBOX * kBKV * 128;
    static constexpr int kPBytes = kBQ * 128;              // 64 keys = one 128-byte chunk per query row
    // K/V ring depth: a stage is re-requested only when the P.V of its previous block has retired, so with 2 stages the
    // TMA round trip (~1.5 us under load) was exposed once per block; 4 stages keep ~3 blocks of K/V in flight.
    static constexpr int kStages = 3;
    static constexpr int kSmemBytes = kQBytes + kStages * (kKBytes + kVBytes) + 2 * kPBytes + 1024 + 256;   // two P buffers
    static constexpr int kTmemCols = DBOX == 1 ? 128 : 256;   // S [0,64)  O [64, 64 + DBOX*64)
    static constexpr int kMinBlocks = 2;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int DBOX>
__global__ void __launch_bounds__(kAttnThreads, AttnCfg<DBOX>::kMinBlocks)
cb_attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
    using Cfg = AttnCfg<DBOX>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;
    const uint32_t sKV = sQ + Cfg::kQBytes;                                 // stage s: K at sKV + s*(K+V), V after K
    const uint32_t sP = sKV + Cfg::kStages * (Cfg::kKBytes + Cfg::kVBytes);
    const uint32_t bars = sP + 2 * Cfg::kPBytes;
    // barriers: q_full, kv_full[2], kv_empty[2], s_full[2], s_empty[2], p_full, pv_done ; then the TMEM slot
    const uint32_t bar_q = bars;
    constexpr int kSt = Cfg::kStages;
    auto bar_kv_full = [&](int s) { return bars + 8u * (1 + s); };
    auto bar_kv_empty = [&](int s) { return bars + 8u * (1 + kSt + s); };
    const uint32_t bar_s_full = bars + 8u * (1 + 2 * kSt);
    const uint32_t bar_s_empty = bars + 8u * (2 + 2 * kSt);
    // P is double-buffered (block j uses buffer j&1), each buffer with its own full / done barrier: the softmax of block
    // j only waits for P.V of block j-2 (buffer reuse); it waits for block j-1 only when it must rescale the output tile.
    auto bar_p_full = [&](int b) { return bars + 8u * (3 + 2 * kSt + b); };
    auto bar_pv_done = [&](int b) { return bars + 8u * (5 + 2 * kSt + b); };
    const uint32_t tmem_slot = bars + 8u * (7 + 2 * kSt);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBQ;
    const int head = blockIdx.y, img = blockIdx.z;
    int nblk = (p.nk + kBKV - 1) / kBKV;
    if (p.causal) nblk = min(nblk, (min(q0 + kBQ, p.nq) + kBKV - 1) / kBKV);
    const int nstat = p.two_pass ? nblk : 0;      // leading statistics-only virtual blocks
    const int nv = nstat + nblk;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(bar_q, 1);
        for (int s = 0; s < kSt; ++s) {
            mbar_init(bar_kv_full(s), 1);
            mbar_init(bar_kv_empty(s), 1);
        }
        mbar_init(bar_s_full, 1);
        mbar_init(bar_s_empty, 128);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_p_full(b), 128);
            mbar_init(bar_pv_done(b), 1);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
    const uint32_t tS0 = tmem, tO = tmem + 64;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            // ============================ TMA producer ============================
            mbar_arrive_expect_tx(bar_q, Cfg::kQBytes);
#pragma unroll
            for (int b = 0; b < DBOX; ++b) tma_load_4d(sQ + b * (kBQ * 128), &tmQ, bar_q, b * 64, q0, head, img);
            for (int vj = 0; vj < nv; ++vj) {
                const int s = vj % kSt;
                const uint32_t ph = (vj / kSt) & 1;
                const bool full = vj >= nstat;
                const int j = full ? vj - nstat : vj;
                mbar_wait(bar_kv_empty(s), ph ^ 1u);
                mbar_arrive_expect_tx(bar_kv_full(s), Cfg::kKBytes + (full ? Cfg::kVBytes : 0));
                const uint32_t dK = sKV + s * (Cfg::kKBytes + Cfg::kVBytes), dV = dK + Cfg::kKBytes;
#pragma unroll
                for (int b = 0; b < DBOX; ++b) {
                    tma_load_4d(dK + b * (kBKV * 128), &tmK, bar_kv_full(s), b * 64, j * kBKV, head, img);
                    if (full) tma_load_4d(dV + b * (kBKV * 128), &tmV, bar_kv_full(s), b * 64, j * kBKV, head, img);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ============================ MMA issuer ============================
            const int ks_qk = p.dpad16 / 16;
            auto issue_S = [&](int vj) {
                const int s = vj % kSt;
                mbar_wait(bar_kv_full(s), (vj / kSt) & 1);
                mbar_wait(bar_s_empty, (vj & 1) ^ 1u);      // softmax has copied the previous scores to registers
                tc_fence_after();
                const uint32_t aK = sKV + s * (Cfg::kKBytes + Cfg::kVBytes);
                for (int k = 0; k < ks_qk; ++k) {
                    const uint32_t offq = (k >> 2) * (kBQ * 128) + (k & 3) * 32;    // 64-col box, then 16-col step
                    const uint32_t offk = (k >> 2) * (kBKV * 128) + (k & 3) * 32;
                    umma_f16(tS0, umma_smem_desc_sw128(sQ + offq, 16, 1024), umma_smem_desc_sw128(aK + offk, 16, 1024),
                             p.idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(bar_s_full);
                if (vj < nstat) umma_commit(bar_kv_empty(s));   // statistics pass: K stage is free once S is done
            };
            mbar_wait(bar_q, 0);
            if (nv > 0) issue_S(0);
            for (int vj = 0; vj < nv; ++vj) {
                if (vj + 1 < nv) issue_S(vj + 1);          // next block's scores overlap this block's softmax
                if (vj < nstat) continue;
                const int fj = vj - nstat;
                const int s = vj % kSt;
                mbar_wait(bar_p_full(fj & 1), (fj >> 1) & 1);
                tc_fence_after();
                const uint32_t aV = sKV + s * (Cfg::kKBytes + Cfg::kVBytes) + Cfg::kKBytes;
                for (int k = 0; k < kBKV / 16; ++k) {
                    // A = P [128 rows][64 keys] K-major: 16-key step k
                    const uint64_t ad = umma_smem_desc_sw128(sP + (fj & 1) * Cfg::kPBytes + k * 32, 16, 1024);
                    // B = V [128 keys][d] read MN-major: 16 keys = 2 groups of 8 rows (SBO 1024), 64-col chunks at LBO
                    const uint64_t bd = umma_smem_desc_sw128(aV + k * 2048, kBKV * 128, 1024);
                    umma_f16(tO, ad, bd, p.idesc_o, (fj > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_kv_empty(s));   // K_j / V_j stage reusable
                umma_commit(bar_pv_done(fj & 1));       // this P buffer reusable, O stable up to block fj
            }
        }
    } else {
        // ============================ softmax / epilogue: one thread per query row ============================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int qrow = q0 + r;
        const bool row_ok = qrow < p.nq;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        float m = -INFINITY, l = 0.f;
        const int et = threadIdx.x - 64;
        const int ocols = DBOX * 64;
        float inv_l = 1.f;
        unsigned tph[6] = {0, 0, 0, 0, 0, 0};
        const bool prof = p.dbg != nullptr && et == 0;
        for (int vj = 0; vj < nv; ++vj) {
            const bool full = vj >= nstat;
            const int j = full ? vj - nstat : vj;
            unsigned tc0 = 0, tc1;
            if (prof) tc0 = (unsigned)clock64();
            mbar_wait(bar_s_full, vj & 1);
            tc_fence_after();
            if (prof) { tc1 = (unsigned)clock64(); tph[0] += tc1 - tc0; tc0 = tc1; }
            uint32_t sc[2][32];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld_32x32(tS0 + lane_off + c * 32, sc[c]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_s_empty);                       // the score tile may be overwritten by S_{vj+1}
            if (prof) { tc1 = (unsigned)clock64(); tph[1] += tc1 - tc0; tc0 = tc1; }
            const int kbase = j * kBKV;
            int kvalid = min(kBKV, p.nk - kbase);
            if (p.causal) kvalid = min(kvalid, qrow - kbase + 1);
            // row max on the raw scores (scale > 0 commutes with max); masked blocks overwrite the invalid tail with
            // -inf first, so the exp loop below is the same for both: e = ex2(fma(s, scale*log2e, -m))
            float mx = -INFINITY;
            if (kvalid >= kBKV) {
                // four independent chains: a single running max is a 64-deep dependent FMNMX chain per block
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sc[c][i]));
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = (c * 32 + i < kvalid) ? __uint_as_float(sc[c][i]) : -INFINITY;
                        sc[c][i] = __float_as_uint(v);
                        mx = fmaxf(mx, v);
                    }
            }
            mx *= p.scale_log2e;
            if (prof) { tc1 = (unsigned)clock64(); tph[2] += tc1 - tc0; tc0 = tc1; }
            const bool online = !(p.two_pass && full);      // running max / sum still being built
            float m_use = m;
            // wait until P.V of block jj has retired (barrier of its P buffer, phase jj>>1)
            auto wait_pv = [&](int jj) { mbar_wait(bar_pv_done(jj & 1), (jj >> 1) & 1); tc_fence_after(); };
            if (online) {
                // Lazy rescale: the reference maximum only moves when the block maximum exceeds it by more than 2^8
                // (probabilities stay <= 256, exact in fp16/fp32), so the TMEM output tile is rescaled -- and the previous
                // P.V waited for -- a handful of times per row instead of nearly every block.
                const float m_new = fmaxf(m, mx);
                float corr = 1.f;
                if (m == -INFINITY) {
                    m = m_new;                               // first block: nothing accumulated yet (l = 0, O untouched)
                } else if (m_new > m + 8.f) {
                    corr = fast_exp2(m - m_new);
                    m = m_new;
                }
                l *= corr;
                if (full) {
                    const bool need = __any_sync(0xffffffffu, j > 0 && corr != 1.f);   // tcgen05.ld/st are warp-collective
                    if (need) {
                        wait_pv(j - 1);                     // the previous P.V must be complete before O is rescaled
                        for (int c = 0; c * 32 < ocols; ++c) {
                            uint32_t o[32];
                            tmem_ld_32x32(tO + lane_off + c * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                            tmem_st_32x32(tO + lane_off + c * 32, o);
                        }
                        tmem_st_wait();
                        tc_fence_before();
                    }
                }
            }
            m_use = (m == -INFINITY) ? 0.f : m;             // (fully masked so far: the exponent argument stays -inf)
            if (full && j > 1) wait_pv(j - 2);              // P buffer j&1 was last read by the P.V of block j-2
            if (prof) { tc1 = (unsigned)clock64(); tph[3] += tc1 - tc0; tc0 = tc1; }
            if (!full) {                                    // statistics pass: accumulate the row sum only
                float rs = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i) rs += fast_exp2(fmaf(__uint_as_float(sc[c][i]), p.scale_log2e, -m_use));
                l += rs;
                if (vj == nstat - 1) inv_l = l > 0.f ? 1.f / l : 0.f;
                continue;
            }
            // P = exp2(s - m) (normalised by 1/l in two-pass mode): swizzled K-major A tile in smem (+ optional HBM copy)
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};      // independent partial row sums (no 64-deep FADD chain)
            uint16_t* prow = (p.P && row_ok)
                                 ? reinterpret_cast<uint16_t*>(p.P) +
                                       ((static_cast<long long>(img) * p.heads + head) * p.nq + qrow) * p.ldp + kbase
                                 : nullptr;
            const float pscale = p.two_pass ? inv_l : 1.f;
            auto write_p = [&](auto scaled) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        e[i] = fast_exp2(fmaf(__uint_as_float(sc[c][g * 8 + i]), p.scale_log2e, -m_use));
                        if (decltype(scaled)::value) e[i] *= pscale;
                        rs4[i & 3] += e[i];
                    }
                    uint4 pk;
                    if (p.p_is_bf16) {
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(e[2 * i], e[2 * i + 1]);
                    } else {
                        __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(e[2 * i], e[2 * i + 1]);
                    }
                    const int key8 = c * 4 + g;                               // 8-key group 0..7
                    const uint32_t dst = sP + (j & 1) * Cfg::kPBytes + r * 128 + ((key8 ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z),
                                 "r"(pk.w)
                                 : "memory");
                    if (prow && kbase + key8 * 8 < p.ldp) *reinterpret_cast<uint4*>(prow + key8 * 8) = pk;
                }
            }
            };
            if (p.two_pass) write_p(std::true_type{}); else write_p(std::false_type{});
            if (online) l += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
            if (prof) { tc1 = (unsigned)clock64(); tph[4] += tc1 - tc0; tc0 = tc1; }
            fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            mbar_arrive(bar_p_full(j & 1));
            if (prof) { tc1 = (unsigned)clock64(); tph[5] += tc1 - tc0; }
        }
        if (prof) {
            unsigned long long* d = p.dbg + 8ull * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
#pragma unroll
            for (int i = 0; i < 6; ++i) d[i] = tph[i];
            d[6] = nv;
        }
        // ---- epilogue: O / l, logsumexp ----
        if (nblk > 1) mbar_wait(bar_pv_done((nblk - 2) & 1), ((nblk - 2) >> 1) & 1);
        if (nblk > 0) mbar_wait(bar_pv_done((nblk - 1) & 1), ((nblk - 1) >> 1) & 1);
        tc_fence_after();
        const float inv = p.two_pass ? 1.f : (l > 0.f ? 1.f / l : 0.f);
        if (row_ok && p.lse && et >= 0)
            p.lse[(static_cast<long long>(img) * p.heads + head) * p.nq + qrow] = (m + log2f(l)) * 0.6931471805599453f;
        for (int c = 0; c * 32 < ocols; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + lane_off + c * 32, o);
            tmem_ld_wait();
            if (row_ok) {
                const long long orow = (static_cast<long long>(img) * p.nq + qrow) * p.ldo + static_cast<long long>(head) * p.d;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int col = c * 32 + g * 8;
                    if (col < p.d) {
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[g * 8 + i]) * inv;
                        uint4 pk;
                        if (p.o_dtype == CB_BF16) {
                            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                            for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                        } else {
                            __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                        }
                        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.O) + orow + col) = pk;
                    }
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem);
}

int make_tmap(CUtensorMap* out, int dtype, int rank, const void* ptr, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr);

template <int DBOX>
static int launch_attn(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnParams& p, dim3 grid,
                       cudaStream_t st) {
    using Cfg = AttnCfg<DBOX>;
    static bool done = false;
    auto kern = cb_attention_fwd_kernel<DBOX>;
    if (!done) {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        done = true;
    }
CB_LAUNCH((kern), grid, kAttnThreads, Cfg::kSmemBytes, st, q, k, v, p);
    CB_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}


// =================================================================================================================
// Backward.  Two launches of one kernel template, both without atomics and without the (heads x N x N) tensors:
//   MODE 0 (dQ)    : CTA = 128 query rows.  X1 = Q, X2 = dO stationary; per 64-key block Y1 = K_j, Y2 = V_j.
//   MODE 1 (dK,dV) : CTA = 128 key rows.    X1 = K, X2 = V  stationary; per 64-query block Y1 = Q_j, Y2 = dO_j.
// Every block: T1 = X1 Y1^T (scores, or their transpose) and T2 = X2 Y2^T (dP, or its transpose) into TMEM;
// one thread per stationary row rebuilds P = exp2(T1*scale*log2e - lse*log2e) and dS = P o (T2 - delta) * scale with the
// forward's log-sum-exp and delta = rowsum(dO o O) (indexed by row in MODE 0, by column in MODE 1) and writes them as
// K-major SWIZZLE_128B A tiles; then  MODE 0: acc += dS Y1 (= dS K);  MODE 1: accV += P^T Y2 (= P^T dO),
// accK += dS^T Y1 (= dS^T Q) with the streamed tiles read MN-major, exactly like V in the forward kernel.
// MODE 0 also computes delta (it owns dO and reads O) and stores it for the MODE 1 launch that follows it in stream order.
struct AttnBwdParams {
    int n_stat, n_stream, heads, images;     // rows of the stationary / streamed operands (nq,nk in MODE 0; nk,nq in MODE 1)
    int d, dpad16, dboxes, causal;
    float scale, scale_log2e;
    const float* lse;        // [images][heads][nq]
    float* delta;            // [images][heads][nq]   (written by MODE 0, read by MODE 1)
    const void* O;           // MODE 0 only: forward output, for delta
    const void* dO;
    long long ldo, lddo;
    void* out1;              // MODE 0: dQ ; MODE 1: dV
    void* out2;              // MODE 1: dK
    long long ld1, ld2;
    int is_bf16;
    int nq;                  // query count (lse / delta row pitch)
    void* dS;                // MODE 0, optional: dS = P o (dP - delta) * scale exported as [images*heads][nq][ldds] (16-bit)
    long long ldds;
    unsigned idesc_t, idesc_acc;
};

template <int DBOX, int MODE>
struct AttnBwdCfg {
    static constexpr int kXBytes = DBOX * kBQ * 128;        // one stationary tile
    static constexpr int kYBytes = DBOX * kBKV * 128;       // one streamed tile
    static constexpr int kABytes = kBQ * 128;               // one A tile (128 rows x 64 block items)
    static constexpr int kNumA = MODE == 0 ? 1 : 2;
    static constexpr int kNumAcc = MODE == 0 ? 1 : 2;
    static constexpr int kSmemBytes = 2 * kXBytes + 2 * 2 * kYBytes + kNumA * kABytes + 1024 + 256 + 1024;
    static constexpr int kTmemNeed = 128 + kNumAcc * DBOX * 64;
    static constexpr int kTmemCols = kTmemNeed <= 256 ? 256 : 512;
    static constexpr int kMinBlocks = (kTmemCols <= 256 && kSmemBytes <= 110 * 1024) ? 2 : 1;
};

__device__ __forceinline__ uint4 store_a16(uint32_t tile, int r, int key8, const float (&e)[8], bool bf16) {
    uint4 pk;
    if (bf16) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(e[2 * i], e[2 * i + 1]);
    } else {
        __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(e[2 * i], e[2 * i + 1]);
    }
    const uint32_t dst = tile + r * 128 + ((key8 ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
    return pk;
}

template <int DBOX, int MODE>
__global__ void __launch_bounds__(kAttnThreads, AttnBwdCfg<DBOX, MODE>::kMinBlocks)
cb_attention_bwd_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                        const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
                        const __grid_constant__ AttnBwdParams p) {
    using Cfg = AttnBwdCfg<DBOX, MODE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sX1 = base, sX2 = base + Cfg::kXBytes;
    const uint32_t sY = sX2 + Cfg::kXBytes;                       // stage s: Y1 at sY + s*2*kYBytes, Y2 right after
    const uint32_t sA = sY + 4 * Cfg::kYBytes;                    // A tile 0 (dS in MODE 0, P^T in MODE 1), tile 1 (dS^T)
    const uint32_t bars = sA + Cfg::kNumA * Cfg::kABytes;
    const uint32_t bar_x = bars;
    auto bar_y_full = [&](int s) { return bars + 8u * (1 + s); };
    auto bar_y_empty = [&](int s) { return bars + 8u * (3 + s); };
    const uint32_t bar_t_full = bars + 8u * 5;
    const uint32_t bar_t_empty = bars + 8u * 6;
    const uint32_t bar_a_full = bars + 8u * 7;
    const uint32_t bar_a_done = bars + 8u * 8;
    const uint32_t tmem_slot = bars + 8u * 9;
    float* s_stat = reinterpret_cast<float*>(smem_raw + (bars + 256 - smem_u32(smem_raw)));    // [2 buffers][2][64] (MODE 1)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * kBQ;
    const int head = blockIdx.y, img = blockIdx.z;
    // streamed block range (causal: keys <= query)
    int jbeg = 0, jend = (p.n_stream + kBKV - 1) / kBKV;
    if (p.causal) {
        if (MODE == 0) jend = min(jend, (min(x0 + kBQ, p.n_stat) + kBKV - 1) / kBKV);     // keys up to the last query row
        else jbeg = x0 / kBKV;                                                            // queries from the first key row
    }
    const int nit = max(0, jend - jbeg);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX1); tma_prefetch_desc(&tmX2); tma_prefetch_desc(&tmY1); tma_prefetch_desc(&tmY2);
        mbar_init(bar_x, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_y_full(s), 1); mbar_init(bar_y_empty(s), 1); }
        mbar_init(bar_t_full, 1);
        mbar_init(bar_t_empty, 128);
        mbar_init(bar_a_full, 128);
        mbar_init(bar_a_done, 1);
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
    const uint32_t tT1 = tmem, tT2 = tmem + 64, tAcc0 = tmem + 128, tAcc1 = tmem + 128 + DBOX * 64;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_x, 2 * Cfg::kXBytes);
#pragma unroll
            for (int b = 0; b < DBOX; ++b) {
                tma_load_4d(sX1 + b * (kBQ * 128), &tmX1, bar_x, b * 64, x0, head, img);
                tma_load_4d(sX2 + b * (kBQ * 128), &tmX2, bar_x, b * 64, x0, head, img);
            }
            for (int it = 0; it < nit; ++it) {
                const int s = it & 1;
                mbar_wait(bar_y_empty(s), ((it >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(bar_y_full(s), 2 * Cfg::kYBytes);
                const uint32_t d1 = sY + s * 2 * Cfg::kYBytes, d2 = d1 + Cfg::kYBytes;
#pragma unroll
                for (int b = 0; b < DBOX; ++b) {
                    tma_load_4d(d1 + b * (kBKV * 128), &tmY1, bar_y_full(s), b * 64, (jbeg + it) * kBKV, head, img);
                    tma_load_4d(d2 + b * (kBKV * 128), &tmY2, bar_y_full(s), b * 64, (jbeg + it) * kBKV, head, img);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const int ks = p.dpad16 / 16;
            auto issue_T = [&](int it) {
                const int s = it & 1;
                mbar_wait(bar_y_full(s), (it >> 1) & 1);
                mbar_wait(bar_t_empty, (it & 1) ^ 1u);
                tc_fence_after();
                const uint32_t y1 = sY + s * 2 * Cfg::kYBytes, y2 = y1 + Cfg::kYBytes;
                for (int k = 0; k < ks; ++k) {
                    const uint32_t offx = (k >> 2) * (kBQ * 128) + (k & 3) * 32;
                    const uint32_t offy = (k >> 2) * (kBKV * 128) + (k & 3) * 32;
                    umma_f16(tT1, umma_smem_desc_sw128(sX1 + offx, 16, 1024), umma_smem_desc_sw128(y1 + offy, 16, 1024),
                             p.idesc_t, k > 0 ? 1u : 0u);
                }
                for (int k = 0; k < ks; ++k) {
                    const uint32_t offx = (k >> 2) * (kBQ * 128) + (k & 3) * 32;
                    const uint32_t offy = (k >> 2) * (kBKV * 128) + (k & 3) * 32;
                    umma_f16(tT2, umma_smem_desc_sw128(sX2 + offx, 16, 1024), umma_smem_desc_sw128(y2 + offy, 16, 1024),
                             p.idesc_t, k > 0 ? 1u : 0u);
                }
                umma_commit(bar_t_full);
            };
            mbar_wait(bar_x, 0);
            if (nit > 0) issue_T(0);
            for (int it = 0; it < nit; ++it) {
                if (it + 1 < nit) issue_T(it + 1);
                const int s = it & 1;
                mbar_wait(bar_a_full, it & 1);
                tc_fence_after();
                const uint32_t y1 = sY + s * 2 * Cfg::kYBytes, y2 = y1 + Cfg::kYBytes;
                for (int k = 0; k < kBKV / 16; ++k) {
                    const uint64_t a0 = umma_smem_desc_sw128(sA + k * 32, 16, 1024);
                    if (MODE == 0) {
                        umma_f16(tAcc0, a0, umma_smem_desc_sw128(y1 + k * 2048, kBKV * 128, 1024), p.idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
                    } else {
                        const uint64_t a1 = umma_smem_desc_sw128(sA + Cfg::kABytes + k * 32, 16, 1024);
                        umma_f16(tAcc0, a0, umma_smem_desc_sw128(y2 + k * 2048, kBKV * 128, 1024), p.idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
                        umma_f16(tAcc1, a1, umma_smem_desc_sw128(y1 + k * 2048, kBKV * 128, 1024), p.idesc_acc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(bar_y_empty(s));
                umma_commit(bar_a_done);
            }
        }
    } else {
        // ====================== one thread per stationary row ======================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int xrow = x0 + r;
        const bool row_ok = xrow < p.n_stat;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const int et = threadIdx.x - 64;
        const long long stat_base = (static_cast<long long>(img) * p.heads + head) * p.nq;
        const float kLog2e = 1.4426950408889634f;
        float lse_r = 0.f, delta_r = 0.f;          // MODE 0: this query row's statistics
        float nlse = 0.f, ndelta = 0.f;            // MODE 1: prefetched column statistics of the next block (threads et < 64)
        if (MODE == 0) {
            if (row_ok) {
                lse_r = p.lse[stat_base + xrow] * kLog2e;
                const uint16_t* orow = reinterpret_cast<const uint16_t*>(p.O) + (static_cast<long long>(img) * p.n_stat + xrow) * p.ldo + static_cast<long long>(head) * p.d;
                const uint16_t* grow = reinterpret_cast<const uint16_t*>(p.dO) + (static_cast<long long>(img) * p.n_stat + xrow) * p.lddo + static_cast<long long>(head) * p.d;
                float acc = 0.f;
                for (int c = 0; c < p.d; c += 8) {
                    const uint4 a = *reinterpret_cast<const uint4*>(orow + c), b = *reinterpret_cast<const uint4*>(grow + c);
                    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float2 fa, fb;
                        if (p.is_bf16) {
                            fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
                            fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[i]));
                        } else {
                            fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[i]));
                            fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[i]));
                        }
                        acc += fa.x * fb.x + fa.y * fb.y;
                    }
                }
                delta_r = acc * p.scale;            // used as fma(T2, scale, -delta*scale)
                p.delta[stat_base + xrow] = acc;
            }
        } else if (et < 64 && nit > 0) {
            const int qc = jbeg * kBKV + et;
            if (qc < p.n_stream) { nlse = p.lse[stat_base + qc] * kLog2e; ndelta = p.delta[stat_base + qc] * p.scale; }
        }
        for (int it = 0; it < nit; ++it) {
            const int j = jbeg + it;
            const int cbase = j * kBKV;
            float2* st_ld = reinterpret_cast<float2*>(s_stat) + (it & 1) * 64;     // (lse*log2e, delta*scale) per streamed query
            if (MODE == 1) {
                if (et < 64) st_ld[et] = make_float2(nlse, ndelta);
                asm volatile("bar.sync 3, 128;" ::: "memory");
                if (et < 64 && it + 1 < nit) {
                    const int qc = (j + 1) * kBKV + et;
                    nlse = 0.f; ndelta = 0.f;
                    if (qc < p.n_stream) { nlse = p.lse[stat_base + qc] * kLog2e; ndelta = p.delta[stat_base + qc] * p.scale; }
                }
            }
            // valid streamed items of this block for this row
            int cvalid = min(kBKV, p.n_stream - cbase);      // columns past the end of the streamed operand
            int cfirst = 0;
            if (p.causal) {
                if (MODE == 0) cvalid = min(cvalid, xrow - cbase + 1);      // keys <= this query
                else cfirst = max(0, xrow - cbase);                         // queries >= this key
            }
            if (!row_ok && MODE == 0) cvalid = 0;
            mbar_wait(bar_t_full, it & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t t1[32], t2[32];
                tmem_ld_32x32(tT1 + lane_off + half * 32, t1);
                tmem_ld_32x32(tT2 + lane_off + half * 32, t2);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    mbar_arrive(bar_t_empty);           // T1/T2 may be overwritten by the next block's MMAs
                } else if (it > 0) {
                    mbar_wait(bar_a_done, (it - 1) & 1);    // the previous block's accumulation MMAs have read the A tiles
                    tc_fence_after();
                }
                // P = ex2(fma(T1, scale*log2e, -lse*log2e)); dS = P * fma(T2, scale, -delta*scale).  Unmasked blocks (all
                // but the ragged tail / the causal diagonal) skip the per-element predicates.
                const bool masked = cvalid < kBKV || cfirst > 0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float pe[8], ds[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = half * 32 + g * 8 + i;
                        float l2, dl;
                        if (MODE == 0) { l2 = lse_r; dl = delta_r; }
                        else { const float2 sd = st_ld[c]; l2 = sd.x; dl = sd.y; }
                        float pv = fast_exp2(fmaf(__uint_as_float(t1[g * 8 + i]), p.scale_log2e, -l2));
                        if (masked) pv = (c < cvalid && c >= cfirst) ? pv : 0.f;
                        pe[i] = pv;
                        ds[i] = pv * fmaf(__uint_as_float(t2[g * 8 + i]), p.scale, -dl);
                    }
                    const int key8 = half * 4 + g;
                    if (MODE == 0) {
                        const uint4 pk = store_a16(sA, r, key8, ds, p.is_bf16 != 0);
                        if (p.dS && row_ok && cbase + key8 * 8 < p.ldds)
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dS) + (stat_base + xrow) * p.ldds + cbase + key8 * 8) = pk;
                    } else {
                        store_a16(sA, r, key8, pe, p.is_bf16 != 0);
                        store_a16(sA + Cfg::kABytes, r, key8, ds, p.is_bf16 != 0);
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_a_full);
        }
        if (MODE == 0 && p.dS && row_ok) {
            // causal: key blocks past this tile's last query were never visited -- their dS is zero
            const int nblk_all = (p.n_stream + kBKV - 1) / kBKV;
            uint16_t* dsrow = reinterpret_cast<uint16_t*>(p.dS) + (stat_base + xrow) * p.ldds;
            for (int j = jend; j < nblk_all; ++j)
                for (int key8 = 0; key8 < 8; ++key8)
                    if (j * kBKV + key8 * 8 < p.ldds) *reinterpret_cast<uint4*>(dsrow + j * kBKV + key8 * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
        // ---- epilogue: accumulators -> HBM ----
        if (nit > 0) mbar_wait(bar_a_done, (nit - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int a = 0; a < Cfg::kNumAcc; ++a) {
            const uint32_t tacc = a == 0 ? tAcc0 : tAcc1;
            uint16_t* outp = reinterpret_cast<uint16_t*>(a == 0 ? p.out1 : p.out2);
            if (outp == nullptr) continue;
            const long long ldout = a == 0 ? p.ld1 : p.ld2;
            for (int c = 0; c * 32 < DBOX * 64; ++c) {
                if (c * 32 >= p.d) break;
                uint32_t o[32];
                if (nit > 0) {
                    tmem_ld_32x32(tacc + lane_off + c * 32, o);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = 0u;
                }
                if (row_ok) {
                    const long long orow = (static_cast<long long>(img) * p.n_stat + xrow) * ldout + static_cast<long long>(head) * p.d;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int col = c * 32 + g * 8;
                        if (col < p.d) {
                            uint4 pk;
                            if (p.is_bf16) {
                                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                                for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(__uint_as_float(o[g * 8 + 2 * i]), __uint_as_float(o[g * 8 + 2 * i + 1]));
                            } else {
                                __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                                for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * i]), __uint_as_float(o[g * 8 + 2 * i + 1]));
                            }
                            *reinterpret_cast<uint4*>(outp + orow + col) = pk;
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem);
}

template <int DBOX, int MODE>
static int launch_attn_bwd(const CUtensorMap& x1, const CUtensorMap& x2, const CUtensorMap& y1, const CUtensorMap& y2,
                           const AttnBwdParams& p, dim3 grid, cudaStream_t st) {
    using Cfg = AttnBwdCfg<DBOX, MODE>;
    static bool done = false;
    auto kern = cb_attention_bwd_kernel<DBOX, MODE>;
    if (!done) {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        done = true;
    }
    CB_LAUNCH((kern), grid, kAttnThreads, Cfg::kSmemBytes, st, x1, x2, y1, y2, p);
    CB_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

static int attn_tmap(CUtensorMap* out, int dtype, const void* ptr, long long ld, int rows, int d, int heads, int images, int box_rows) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)rows, (uint64_t)heads, (uint64_t)images};
    uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)d * 2, (uint64_t)rows * ld * 2};
    uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
    uint32_t estr[4] = {1, 1, 1, 1};
    return make_tmap(out, dtype, 4, ptr, dims, str, box, estr);
}

}  // namespace cb

using namespace cb;

static unsigned long long* g_attn_dbg = nullptr;
// profiling hook (not part of the public header): device buffer of 8 x uint64 per CTA, or NULL to disable
extern "C" void cb_debug_attention_timeline(void* buf) { g_attn_dbg = reinterpret_cast<unsigned long long*>(buf); }

extern "C" int cb_attention_fwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                                void* O, long long ldo, float* lse, void* P, long long ldp, int dtype, int images,
                                int heads, int nq, int nk, int d, float scale, int causal, void* stream) {
    CB_REQUIRE(dtype == CB_F16 || dtype == CB_BF16, CB_ERR_ARG, "attention_fwd: dtype must be f16/bf16");
    CB_REQUIRE(Q && K && V && O && images > 0 && heads > 0 && nq > 0 && nk > 0 && scale > 0.f, CB_ERR_ARG, "attention_fwd: bad args");
    CB_REQUIRE(d >= 8 && d <= 128 && d % 8 == 0, CB_ERR_ARG, "attention_fwd: head dim %d unsupported (8..128, multiple of 8)", d);
    CB_REQUIRE((ldo * 2) % 16 == 0 && ((long long)d * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(O) & 15u) == 0,
               CB_ERR_ALIGN, "attention_fwd: O alignment");
    const int es = 2;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    p.nq = nq; p.nk = nk; p.heads = heads; p.images = images;
    p.d = d; p.dpad16 = (d + 15) / 16 * 16; p.dboxes = d <= 64 ? 1 : 2;
    p.causal = causal;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.O = O; p.o_dtype = dtype; p.ldo = ldo; p.lse = lse;
    p.P = P; p.ldp = ldp; p.p_is_bf16 = dtype == CB_BF16;
    p.two_pass = P != nullptr ? 1 : 0;
    p.dbg = g_attn_dbg;
    if (P) CB_REQUIRE(ldp >= nk && ldp % 8 == 0 && (reinterpret_cast<uintptr_t>(P) & 15u) == 0, CB_ERR_ALIGN, "attention_fwd: P row pitch must be a multiple of 8 elements >= nk");
    // S = Q K^T : M=128, N=128 keys, both K-major.  O = P V : M=128, N=dpad16, A K-major, B (V) MN-major.
    p.idesc_s = umma_idesc_f16(128, kBKV, dtype == CB_BF16, false, false);
    p.idesc_o = umma_idesc_f16(128, p.dboxes * 64, dtype == CB_BF16, false, true);   // N = whole 64-col chunks (zero-filled past d)
    CUtensorMap tq, tk, tv;
    uint32_t estr[4] = {1, 1, 1, 1};
    uint32_t box[4] = {64, kBQ, 1, 1};
    uint32_t boxkv[4] = {64, kBKV, 1, 1};
    {
        uint64_t dims[4] = {(uint64_t)d, (uint64_t)nq, (uint64_t)heads, (uint64_t)images};
        uint64_t str[3] = {(uint64_t)ldq * es, (uint64_t)d * es, (uint64_t)nq * ldq * es};
        int rc = make_tmap(&tq, dtype, 4, Q, dims, str, box, estr);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)d, (uint64_t)nk, (uint64_t)heads, (uint64_t)images};
        uint64_t str[3] = {(uint64_t)ldk * es, (uint64_t)d * es, (uint64_t)nk * ldk * es};
        int rc = make_tmap(&tk, dtype, 4, K, dims, str, boxkv, estr);
        if (rc) return rc;
        uint64_t strv[3] = {(uint64_t)ldv * es, (uint64_t)d * es, (uint64_t)nk * ldv * es};
        rc = make_tmap(&tv, dtype, 4, V, dims, strv, boxkv, estr);
        if (rc) return rc;
    }
    dim3 grid((unsigned)ceil_div(nq, kBQ), (unsigned)heads, (unsigned)images);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p.dboxes == 1) return launch_attn<1>(tq, tk, tv, p, grid, st);
    return launch_attn<2>(tq, tk, tv, p, grid, st);
}

static int attention_bwd_impl(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                              const void* O, long long ldo, const void* dO, long long lddo, const float* lse, float* delta,
                              void* dQ, long long lddq, void* dK, long long lddk, void* dV, long long lddv, void* dS,
                              long long ldds, bool run_dq, bool run_dkdv, int dtype, int images, int heads, int nq, int nk,
                              int d, float scale, int causal, void* stream) {
    CB_REQUIRE(dtype == CB_F16 || dtype == CB_BF16, CB_ERR_ARG, "attention_bwd: dtype must be f16/bf16");
    CB_REQUIRE(Q && K && V && O && dO && lse && delta && images > 0 && heads > 0 && nq > 0 && nk > 0 && scale > 0.f,
               CB_ERR_ARG, "attention_bwd: bad args");
    CB_REQUIRE(!run_dkdv || (dK && dV), CB_ERR_ARG, "attention_bwd: dK/dV required");
    CB_REQUIRE(d >= 8 && d <= 128 && d % 8 == 0, CB_ERR_ARG, "attention_bwd: head dim %d unsupported (8..128, multiple of 8)", d);
    const void* ptrs[8] = {Q, K, V, O, dO, dQ, dK, dV};
    const long long lds[8] = {ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv};
    for (int i = 0; i < 8; ++i)
        CB_REQUIRE(ptrs[i] == nullptr || ((lds[i] * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(ptrs[i]) & 15u) == 0), CB_ERR_ALIGN,
                   "attention_bwd: operand %d must be 16-byte aligned with a 16-byte multiple row pitch", i);
    if (dS) CB_REQUIRE(ldds >= nk && ldds % 8 == 0 && (reinterpret_cast<uintptr_t>(dS) & 15u) == 0, CB_ERR_ALIGN, "attention_bwd: dS row pitch must be a multiple of 8 elements >= nk");
    AttnBwdParams p;
    memset(&p, 0, sizeof(p));
    p.heads = heads; p.images = images; p.d = d; p.dpad16 = (d + 15) / 16 * 16; p.dboxes = d <= 64 ? 1 : 2;
    p.causal = causal; p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
    p.lse = lse; p.delta = delta; p.O = O; p.dO = dO; p.ldo = ldo; p.lddo = lddo;
    p.is_bf16 = dtype == CB_BF16; p.nq = nq;
    p.idesc_t = umma_idesc_f16(128, kBKV, dtype == CB_BF16, false, false);
    p.idesc_acc = umma_idesc_f16(128, p.dboxes * 64, dtype == CB_BF16, false, true);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap x1, x2, y1, y2;
    int rc = 0;
    if (run_dq) {
        // ---- MODE 0: dQ (+ delta, + optional dS export) ----
        if ((rc = attn_tmap(&x1, dtype, Q, ldq, nq, d, heads, images, kBQ))) return rc;
        if ((rc = attn_tmap(&x2, dtype, dO, lddo, nq, d, heads, images, kBQ))) return rc;
        if ((rc = attn_tmap(&y1, dtype, K, ldk, nk, d, heads, images, kBKV))) return rc;
        if ((rc = attn_tmap(&y2, dtype, V, ldv, nk, d, heads, images, kBKV))) return rc;
        p.n_stat = nq; p.n_stream = nk; p.out1 = dQ; p.ld1 = lddq; p.out2 = nullptr; p.ld2 = 0; p.dS = dS; p.ldds = ldds;
        dim3 grid((unsigned)ceil_div(nq, kBQ), (unsigned)heads, (unsigned)images);
        rc = p.dboxes == 1 ? launch_attn_bwd<1, 0>(x1, x2, y1, y2, p, grid, st) : launch_attn_bwd<2, 0>(x1, x2, y1, y2, p, grid, st);
        if (rc) return rc;
    }
    if (run_dkdv) {
        // ---- MODE 1: dK, dV ----
        if ((rc = attn_tmap(&x1, dtype, K, ldk, nk, d, heads, images, kBQ))) return rc;
        if ((rc = attn_tmap(&x2, dtype, V, ldv, nk, d, heads, images, kBQ))) return rc;
        if ((rc = attn_tmap(&y1, dtype, Q, ldq, nq, d, heads, images, kBKV))) return rc;
        if ((rc = attn_tmap(&y2, dtype, dO, lddo, nq, d, heads, images, kBKV))) return rc;
        p.n_stat = nk; p.n_stream = nq; p.out1 = dV; p.ld1 = lddv; p.out2 = dK; p.ld2 = lddk; p.dS = nullptr; p.ldds = 0;
        dim3 grid((unsigned)ceil_div(nk, kBQ), (unsigned)heads, (unsigned)images);
        rc = p.dboxes == 1 ? launch_attn_bwd<1, 1>(x1, x2, y1, y2, p, grid, st) : launch_attn_bwd<2, 1>(x1, x2, y1, y2, p, grid, st);
    }
    return rc;
}

extern "C" int cb_attention_bwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                                const void* O, long long ldo, const void* dO, long long lddo, const float* lse, float* delta,
                                void* dQ, long long lddq, void* dK, long long lddk, void* dV, long long lddv, int dtype,
                                int images, int heads, int nq, int nk, int d, float scale, int causal, void* stream) {
    CB_REQUIRE(dQ != nullptr, CB_ERR_ARG, "attention_bwd: dQ required");
    return attention_bwd_impl(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, dK, lddk, dV, lddv, nullptr, 0,
                              true, true, dtype, images, heads, nq, nk, d, scale, causal, stream);
}

extern "C" int cb_attention_bwd_dq(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                                   const void* O, long long ldo, const void* dO, long long lddo, const float* lse,
                                   float* delta, void* dQ, long long lddq, void* dS, long long ldds, int dtype, int images,
                                   int heads, int nq, int nk, int d, float scale, int causal, void* stream) {
    CB_REQUIRE(dQ != nullptr || dS != nullptr, CB_ERR_ARG, "attention_bwd_dq: nothing to compute");
    return attention_bwd_impl(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, delta, dQ, lddq, nullptr, 0, nullptr, 0, dS, ldds,
                              true, false, dtype, images, heads, nq, nk, d, scale, causal, stream);
}
