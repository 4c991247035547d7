// Federated logistic GLM on a BLOCK-SCALED FP8 design matrix (tcgen05.mma kind::mxf8f6f4.block_scale).
//
// Storage: X as e4m3 bytes + one UE8M0 scale per 32 x 32 block (32 rows x 32 features):
//     x[r, f] = e4m3(Xq[r, f]) * 2^(S[r/32, f/32] - 127)
// The kernel receives the scales pre-packed per 128-row tile in the word order of the TMEM
// scale-factor columns (16 words per tile), so the epilogue warps only move them.
// i.e. 1.03 bytes per element instead of 2 for bf16 -> the HBM-bound evaluation gets ~2x faster.
//
// Both GEMMs consume the quantised tile directly; the dequantisation is done by the tensor core.
// TMEM scale-factor layout (cta_group::1, M = 128, 32-element blocks): the scale of operand row m
// for K-block j sits in lane m % 32 (replicated in all four 32-lane subpartitions), 32-bit column
// base + 4*(j/4) + m/32, byte j % 4 — the byte is selected by the instruction descriptor's sf_id.
// With 32-row x 32-feature blocks every lane holds the same word, so the epilogue warps can
// write the scale words with plain tcgen05.st:
//   MMA #1  eta[128 x 16] += Xq_tile[:, feature block fb] . Theta^T     rows m -> row group m/32:
//           column 4*(fb/4) + q holds the scales of (row group q, feature blocks 4*(fb/4)..+3),
//           sf_id = fb % 4.
//   MMA #2  G[128 feat x 16] += Xq_tile^T[:, row group q] . R           rows m -> feature block
//           4h + m/32: column 4h + qq holds the scales of (row groups 0..3, feature block 4h+qq),
//           sf_id = q.
//   Scale-factor-B is the constant 2^0 for MMA #1 and for the logistic family (|r| <= 1).
//   Poisson / Gaussian residuals are unbounded: there (template DYN) every 32-row group of R gets
//   its own UE8M0 scale — warp max -> exponent, exchanged through shared memory, written to the
//   scale-factor-B columns of the R buffer by the epilogue warps, byte selected by b_sf_id = row
//   group — so the radix-16 expansion is relative to the block maximum, which is what the
//   block-scaled MMA is for.
// Theta and the residuals are themselves e4m3: theta is a 5-term, r a 4-term radix-16 expansion
// (term k carries weight 16^-k), the terms sit in separate N columns and are recombined in the
// epilogue in fp32, so the low precision of the operands does not leak into the result.
//
// Pipeline/roles as in glm_tc.cu; scale words are written to TMEM kEG tiles ahead by the
// epilogue warps (tcgen05.st), 2*kEG-deep ring.  Workload: BASELINE.json "hierarchical GLM, 8
// partial-pooling groups (one per GPU), fp8 block-scaled design matrix" (groups = intercepts).
#include <cuda.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "fed_comm.cuh"
#include "models.h"
#include "tc_common.cuh"
#include "chunks.h"

namespace fp8 {
using namespace tc;

constexpr int kTile = 128;            // rows per tile
constexpr int kPanelF = 128;          // features per 128-byte swizzle span (1 byte / element)
constexpr int kPanelB = kTile * 128;  // 16 KB
constexpr int kEG = 3;                // epilogue groups; tile t belongs to group t % kEG (own eta / R buffers)
constexpr int kThreadsF = 32 * (3 + 4 * kEG);  // warps: 0 TMA, 1 MMA#1, 2-5 group 0, 6 MMA#2, 7-10 group 1, 11-14 group 2
constexpr int kMaxChunkF = 30;        // tiles per chunk at most (fp32 accumulation in TMEM), a multiple of kEG
constexpr int kMinChunkF = 6;
constexpr int kRingF = 16;            // published chunks the consumers may lag behind
constexpr int kLLRowsF = 4 * kEG;     // epilogue warps: per-warp slots of the warp-level sums
constexpr int kN = 16;                // MMA N for both GEMMs
constexpr int kThetaTerms = 5;
constexpr int kResidTerms = 4;
constexpr int kSfRing = 2 * kEG;      // scale words are written kEG tiles ahead

__device__ __forceinline__ void umma_fp8_block_scaled(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                      uint32_t accumulate, uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// Block-scaled instruction descriptor: e4m3 x e4m3 -> f32, UE8M0 scales, K = 32.
__host__ __device__ constexpr uint32_t make_idesc_bs(int M, int N, int a_mn_major, int b_mn_major) {
    return ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | (1u << 23) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&w)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(w[0]),
                 "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_x1(uint32_t taddr, uint32_t w) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(w) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint8_t to_e4m3(float v) {
    return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float from_e4m3(uint8_t b) {
    const __half_raw h = __nv_cvt_fp8_to_halfraw(b, __NV_E4M3);
    return __half2float(*reinterpret_cast<const __half*>(&h));
}
// radix-16 expansion: v ~= sum_k terms[k] * 16^-k, each term an e4m3 of magnitude <= 448
template <int T>
__device__ __forceinline__ void expand16(float v, uint8_t (&terms)[T]) {
    float rem = v;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        terms[k] = to_e4m3(rem);
        rem = (rem - from_e4m3(terms[k])) * 16.f;
    }
}

struct SmemLayoutF {
    uint32_t stages, stage_bytes, off_theta_b, theta_b_bytes, off_r, r_bytes, off_theta_f, off_ring, off_bars, off_tmem, total;
};
// doubles per CTA row of the partial array: (hi, lo) pairs of the n_vals outputs, then the per-warp slots
// [kLLRowsF][KF][1 + G] of the warp-level sums (log-likelihood, intercept gradients) — see csrc/glm_tc.cu
__host__ __device__ constexpr size_t partial_row_doubles(int n_vals, int kf, int n_out, int n_groups) {
    return 2 * ((size_t)n_vals + (size_t)kLLRowsF * n_out * kf * (1 + n_groups));
}
__host__ __device__ inline SmemLayoutF smem_layout(int P, int n_theta, int n_groups) {
    SmemLayoutF L;
    const uint32_t panels = P / kPanelF;
    L.stage_bytes = panels * kPanelB;
    L.theta_b_bytes = panels * kN * 128;
    L.r_bytes = kTile * kN;  // 1 byte per element
    const uint32_t fixed = L.theta_b_bytes + kEG * L.r_bytes + ((n_theta * 4 + 15) & ~15) + kRingF * 24 + 512 + 1024;
    uint32_t stages = (227u * 1024u - fixed) / L.stage_bytes;
    if (stages > 6) stages = 6;
    L.stages = stages;
    uint32_t o = stages * L.stage_bytes;
    L.off_theta_b = o; o += L.theta_b_bytes;
    L.off_r = o; o += kEG * L.r_bytes;
    L.off_theta_f = o; o += (n_theta * 4 + 15) & ~15;
    L.off_ring = o; o += kRingF * 24;   // published chunks + one mbarrier per ring slot
    L.off_bars = o; o += 320;
    L.off_tmem = o; o += 192;  // tmem slot, theta norms, residual-exponent exchange (DYN)
    L.total = o + 1024;
    return L;
}

// KF = chains per launch: 1 or 3 (3 x 5 theta terms and 3 x 4 residual terms fit the 16 MMA columns);
// DYN = per-row-group residual scales (families with unbounded residuals)
template <int KF, bool DYN>
__global__ void __launch_bounds__(kThreadsF, 1)
fed_glm_fp8_kernel(FedComm comm, const GlmSegment* __restrict__ segs_g, GlmParams prm, const CUtensorMap* __restrict__ tmaps,
                   const GlmChunk* __restrict__ chunks, int n_chunks, unsigned int* __restrict__ work_counter) {
    extern __shared__ unsigned char smem_dyn[];
    // 1 KB alignment (128B-swizzled TMA tiles) by offsetting INSIDE the shared array: the pointer keeps its
    // shared address space, so the compiler emits LDS / STS instead of generic LD / ST for everything below
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);

    const int P = prm.n_features;
    const int G = prm.n_groups;
    const int NH = P / 128;       // 128-feature panels == UMMA M blocks of MMA #2
    const int NFB = P / 32;       // 32-feature scale blocks == K steps of MMA #1
    const SmemLayoutF L = smem_layout(P, comm.n_theta, G);
    const int S = (int)L.stages;

    unsigned char* theta_b = smem + L.off_theta_b;
    unsigned char* r_buf = smem + L.off_r;
    float* theta_f = reinterpret_cast<float*>(smem + L.off_theta_f);
    int4* ring = reinterpret_cast<int4*>(smem + L.off_ring);          // published chunks: (segment or -1, first tile, tiles, -)
    uint64_t* bar_ring = reinterpret_cast<uint64_t*>(smem + L.off_ring + kRingF * 16);   // slot j % kRingF: chunk j published
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.off_tmem);
    float* theta_norm = reinterpret_cast<float*>(smem + L.off_tmem + 16);  // c: theta = c * sum_k t_k 16^-k
    uint8_t* r_expo = smem + L.off_tmem + 64;  // [group][tile parity][chain (3)][row group (4)] UE8M0 bytes
    uint64_t* bar_full = bars;            // [6]
    uint64_t* bar_empty = bars + 6;       // [6]
    uint64_t* bar_eta_full = bars + 12;             // [kEG]
    uint64_t* bar_eta_empty = bars + 12 + kEG;      // [kEG]
    uint64_t* bar_r_full = bars + 12 + 2 * kEG;     // [kEG]
    uint64_t* bar_r_empty = bars + 12 + 3 * kEG;    // [kEG]
    uint64_t* bar_g_full = bars + 12 + 4 * kEG;     // [2]
    uint64_t* bar_g_empty = bars + 14 + 4 * kEG;    // [2]
    uint64_t* bar_sf_full = bars + 16 + 4 * kEG;    // [kSfRing]   (16 + 6 kEG <= 40 barriers = 320 bytes)
    static_assert(16 + 4 * kEG + kSfRing <= 40, "barrier block too small");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- theta-independent setup: BEFORE the dependency wait inside fed::prologue, so that it overlaps with the
    // tail of the previous evaluation under programmatic dependent launch (see csrc/glm_tc.cu)
    constexpr uint32_t kTmemCols = 256;
    for (int i = threadIdx.x; i < (int)(kEG * L.r_bytes / 16); i += blockDim.x)
        reinterpret_cast<uint4*>(r_buf)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        *pipeline_fault() = 0;
        for (int i = 0; i < kRingF; ++i) mbar_init(&bar_ring[i], 1);
        for (int i = 0; i < 6; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < kEG; ++i) {
            mbar_init(&bar_eta_full[i], 1);
            mbar_init(&bar_eta_empty[i], 128);
            mbar_init(&bar_r_full[i], 128);
            mbar_init(&bar_r_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_g_full[i], 1);
            mbar_init(&bar_g_empty[i], 128);
        }
        for (int i = 0; i < kSfRing; ++i) mbar_init(&bar_sf_full[i], 128);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0)
        for (int i = 0; i < prm.n_segments; ++i) tma_prefetch_desc(&tmaps[i]);
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    // Early loads (see csrc/glm_tc.cu): the TMA warp claims the first chunk and fills the stage ring before
    // theta has arrived; theta has its own shared memory here, so every stage can be used.
    unsigned int claim0 = 0;
    int preloaded = 0;
    if (warp == 0) {
        __syncwarp();   // lane 0's mbarrier inits
        fed::pdl_wait();
        if (lane == 0) claim0 = atomicAdd(work_counter, 1u);
        claim0 = __shfl_sync(0xffffffffu, claim0, 0);
        if (claim0 < (unsigned int)n_chunks) {
            const GlmChunk ch = chunks[claim0];
            preloaded = ch.n_tiles < S ? ch.n_tiles : S;
            if (!prm.early_loads) preloaded = 0;
            if (elect_one()) {
                for (int t = 0; t < preloaded; ++t) {
                    mbar_expect_tx(&bar_full[t], L.stage_bytes);
                    for (int pnl = 0; pnl < NH; ++pnl)
                        tma_load_2d(smem + (size_t)t * L.stage_bytes + pnl * kPanelB, &tmaps[ch.seg], pnl * kPanelF,
                                    (ch.first_tile + t) * kTile, &bar_full[t]);
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();

    fed::Prologue pro = fed::prologue(comm, theta_f);   // contains __syncthreads()
    const bool active = !pro.stop && !pro.timed_out;
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr float kResidNorm = 1.f / 256.f;  // r = kResidNorm * sum_k t_k 16^-k, |r| <= 1

    const int nch = prm.n_chains < KF ? prm.n_chains : KF;
    const int PG = G + P;  // parameters per chain
    const int NV1 = 1 + PG;   // outputs per chain: [LL, gi[G], g[P]]
    const int NS1 = 1 + G;    // warp-level values per chain: LL and the G intercept gradients
    // this CTA's running sums as (hi, lo) pairs (see csrc/glm_tc.cu: dynamic chunks + double-double sums)
    const int NOUT = prm.n_out;   // output blocks (1 = everything summed; else one per node, GlmSegment::out_group)
    const size_t row_doubles = partial_row_doubles(comm.n_vals, KF, NOUT, G);
    double* out = comm.cta_partials + (size_t)blockIdx.x * row_doubles;
    double* ll_slots = out + 2 * (size_t)comm.n_vals;   // [kLLRowsF][NOUT][KF][1 + G] pairs

    if (active) {
        for (size_t i = threadIdx.x; i < row_doubles / 2; i += blockDim.x) reinterpret_cast<double2*>(out)[i] = make_double2(0.0, 0.0);
        // ---- theta normalisation per chain: power of two c with max|beta| / c in [128, 256)
        if (warp < KF) {
            float m = 0.f;
            if (warp < nch)
                for (int f = lane; f < P; f += 32) m = fmaxf(m, fabsf(theta_f[warp * PG + G + f]));
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) {
                int e = 0;
                if (m > 0.f) {
                    frexpf(m, &e);      // m = frac * 2^e, frac in [0.5, 1)
                    e -= 8;             // m / 2^(e-8) in [128, 256)
                }
                theta_norm[warp] = ldexpf(1.f, e);
            }
        }
        __syncthreads();
        float c_theta[KF];
#pragma unroll
        for (int k = 0; k < KF; ++k) c_theta[k] = theta_norm[k];
        // ---- Theta^T as the K-major, 128B-swizzled e4m3 B operand: row n = 5 * chain + term
        for (int idx = threadIdx.x; idx < NH * kN * 8; idx += blockDim.x) {
            const int j = idx & 7;                 // 16-byte chunk = 16 features
            const int n = (idx >> 3) % kN;
            const int pnl = idx / (8 * kN);
            const int chain = n / kThetaTerms, term = n % kThetaTerms;
            uint32_t packed[4] = {0, 0, 0, 0};
            if (chain < nch) {
                const float inv_c = 1.f / theta_norm[chain];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    uint8_t terms[kThetaTerms];
                    expand16<kThetaTerms>(theta_f[chain * PG + G + pnl * kPanelF + j * 16 + e] * inv_c, terms);
                    packed[e >> 2] |= (uint32_t)terms[term] << ((e & 3) * 8);
                }
            }
            *reinterpret_cast<uint4*>(theta_b + pnl * (kN * 128) + n * 128 + ((j ^ (n & 7)) * 16)) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        fence_proxy_async();
        __syncthreads();
        const uint32_t tmem_eta = tmem_base;                         // kEG x 16            (<= 64)
        const uint32_t tmem_g = tmem_base + 64;                      // 2 buffers x NH x 16 (<= 64)
        const uint32_t tmem_sfa1 = tmem_base + 128;                  // ring x 8 columns
        const uint32_t tmem_sfa2 = tmem_sfa1 + kSfRing * 8;          // ring x 8 columns
        const uint32_t tmem_sfb = tmem_sfa2 + kSfRing * 8;           // 4 columns of 2^0
        const uint32_t tmem_sfb_r = tmem_sfb + 4;                    // DYN: kEG R buffers x 4 columns (col 0 used)
        static_assert(128 + kSfRing * 16 + 4 + kEG * 4 <= (int)kTmemCols, "TMEM budget");

        // Consumers: the j-th chunk of this CTA, or x < 0 when the producer found the work counter exhausted.
        // (an mbarrier per ring slot: the producer arrives after writing the entry — release —, the consumers wait
        // on the slot's phase — acquire; slot j % kRingF is reused every kRingF chunks, consumers lag ~3 at most)
        auto next_chunk = [&](int j) -> int4 {
            mbar_wait(&bar_ring[j & (kRingF - 1)], (uint32_t)((j / kRingF) & 1));
            if (*pipeline_fault()) return make_int4(-1, 0, 0, 0);   // a stalled pipeline ends every role loop
            return ring[j & (kRingF - 1)];
        };
        // non-blocking variant for look-ahead: x == -2 when chunk j has not been published yet
        auto peek_chunk = [&](int j) -> int4 {
            if (!mbar_try_wait(&bar_ring[j & (kRingF - 1)], (uint32_t)((j / kRingF) & 1))) return make_int4(-2, 0, 0, 0);
            return ring[j & (kRingF - 1)];
        };

        if (warp == 0) {
            // ================= TMA producer + chunk scheduler (see csrc/glm_tc.cu) ==============
            Ring stage;
            unsigned int claim = claim0, ahead = 0;   // the first chunk was claimed before theta arrived
            for (int j = 0;; ++j) {
                const bool have = claim < (unsigned int)n_chunks;
                GlmChunk ch{};
                if (have) ch = chunks[claim];
                if (lane == 0) {
                    ring[j & (kRingF - 1)] = have ? make_int4(ch.seg, ch.first_tile, ch.n_tiles, 0) : make_int4(-1, 0, 0, 0);
                    mbar_arrive(&bar_ring[j & (kRingF - 1)]);
                    if (have) ahead = atomicAdd(work_counter, 1u);
                }
                __syncwarp();
                if (!have) break;
                for (int t = 0; t < ch.n_tiles; ++t) {
                    const int st = stage.idx;
                    if (j == 0 && t < preloaded) {   // already in flight (early loads)
                        stage.advance(S);
                        continue;
                    }
                    mbar_wait(&bar_empty[st], stage.phase ^ 1);
                    const int row0 = (ch.first_tile + t) * kTile;
                    unsigned char* dst = smem + (size_t)st * L.stage_bytes;
                    if (elect_one()) {
                        mbar_expect_tx(&bar_full[st], L.stage_bytes);
                        for (int pnl = 0; pnl < NH; ++pnl)
                            tma_load_2d(dst + pnl * kPanelB, &tmaps[ch.seg], pnl * kPanelF, row0, &bar_full[st]);
                    }
                    __syncwarp();
                    stage.advance(S);
                }
                claim = __shfl_sync(0xffffffffu, ahead, 0);
            }
        } else if (warp == 1) {
            {
                constexpr uint32_t idesc1 = make_idesc_bs(128, kN, 0, 0);
                // descriptors: constant fields once, the 14-bit (address >> 4) field added per MMA
                const uint64_t desc_k = make_desc(0, 16, 1024, 2);
                const uint32_t theta_b_a4 = smem_u32(theta_b) >> 4;
                const uint32_t x_base_a4 = smem_u32(smem) >> 4;
                const uint32_t stage_a4 = L.stage_bytes >> 4;
                Ring stage, buf, sf;   // TMA stages (S), eta buffers (kEG), scale-word slots (kSfRing)
                for (int j = 0;; ++j) {
                  const int4 ch = next_chunk(j);
                  if (ch.x < 0) break;
                  for (int t = 0; t < ch.z; ++t) {
                    mbar_wait(&bar_sf_full[sf.idx], sf.phase);
                    mbar_wait(&bar_eta_empty[buf.idx], buf.phase ^ 1);
                    mbar_wait(&bar_full[stage.idx], stage.phase);
                    tc_fence_after();
                    const uint32_t x_a4 = x_base_a4 + (uint32_t)stage.idx * stage_a4;
                    const uint32_t d_eta = tmem_eta + buf.idx * kN;
                    const uint32_t sfa = tmem_sfa1 + sf.idx * 8;
                    if (elect_one()) {
                        for (int pnl = 0; pnl < NH; ++pnl) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {   // 4 K-steps of 32 features per 128-feature panel
                                const uint64_t adesc = desc_k | (uint64_t)(x_a4 + ((pnl * kPanelB + ks * 32) >> 4));
                                const uint64_t bdesc = desc_k | (uint64_t)(theta_b_a4 + ((pnl * (kN * 128) + ks * 32) >> 4));
                                umma_fp8_block_scaled(d_eta, adesc, bdesc, idesc1 | ((uint32_t)ks << 29), (pnl | ks) ? 1u : 0u,
                                                      sfa + pnl * 4, tmem_sfb);
                            }
                        }
                        umma_commit(&bar_eta_full[buf.idx]);
                    }
                    __syncwarp();
                    stage.advance(S);
                    buf.advance(kEG);
                    sf.advance(kSfRing);
                  }
                }
            }
        } else if (warp == 6) {
            {
                constexpr uint32_t idesc2 = make_idesc_bs(128, kN, 1, 1);
                const uint64_t desc_x = make_desc(0, kPanelB, 1024, 2);
                const uint64_t desc_r = make_desc(0, 128, 128, 0);
                const uint32_t r_a4 = smem_u32(r_buf) >> 4;
                const uint32_t x_base_a4 = smem_u32(smem) >> 4;
                const uint32_t stage_a4 = L.stage_bytes >> 4;
                Ring stage, buf, sf, gbuf;   // gbuf: TMEM gradient accumulator of the current chunk (2, alternating)
                for (int j = 0;; ++j) {
                  const int4 ch = next_chunk(j);
                  if (ch.x < 0) break;
                  const int gb = gbuf.idx;
                  mbar_wait(&bar_g_empty[gb], gbuf.phase ^ 1);   // the epilogue has drained this accumulator
                  for (int t = 0; t < ch.z; ++t) {
                    const bool first = t == 0;
                    const bool last = t == ch.z - 1;
                    mbar_wait(&bar_r_full[buf.idx], buf.phase);
                    tc_fence_after();
                    const uint32_t x_a4 = x_base_a4 + (uint32_t)stage.idx * stage_a4;
                    const uint32_t rb_a4 = r_a4 + (uint32_t)buf.idx * (L.r_bytes >> 4);
                    const uint32_t sfa = tmem_sfa2 + sf.idx * 8;
                    const uint32_t sfbr = DYN ? tmem_sfb_r + buf.idx * 4 : tmem_sfb;
                    if (elect_one()) {
                    for (int h = 0; h < NH; ++h) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {   // K step = one 32-row group
                            const uint64_t adesc = desc_x | (uint64_t)(x_a4 + ((h * kPanelB + q * 4 * 1024) >> 4));
                            const uint64_t bdesc = desc_r | (uint64_t)(rb_a4 + ((q * 4 * 128) >> 4));
                            // DYN: scale-factor-B byte q of the R buffer's column = exponent of row group q
                            umma_fp8_block_scaled(tmem_g + (gb * NH + h) * kN, adesc, bdesc,
                                                  idesc2 | ((uint32_t)q << 29) | (DYN ? ((uint32_t)q << 4) : 0u),
                                                  (first && q == 0) ? 0u : 1u, sfa + h * 4, sfbr);
                        }
                    }
                    umma_commit(&bar_empty[stage.idx]);
                    umma_commit(&bar_r_empty[buf.idx]);
                    if (last) umma_commit(&bar_g_full[gb]);
                    }
                    __syncwarp();
                    stage.advance(S);
                    buf.advance(kEG);
                    sf.advance(kSfRing);
                  }
                  gbuf.advance(2);
                }
            }
        } else {
            // ================= epilogue warps: group g = warps 2..5 / 7..10 / 11..14 owns tiles t % kEG == g
            // The per-tile epilogue is a long serial chain (barrier wake-up, TMEM load, link maths,
            // smem store + proxy fence, TMEM scale stores).  kEG groups rotate over the tiles so that kEG
            // chains overlap; group g owns eta/R buffer g, so every barrier still sees 128 arrivals.
            // Measured (8 x 10M x 256): 1 group 269, 2 groups 279 -> 295 with host-packed scale words,
            // 3 groups 297 evals/s — beyond two groups the group count is no longer the bound
            // (prefetching y one iteration ahead did not move it either: 293).
            const int eg = warp >= 7 ? (warp - 7) / 4 + 1 : 0;
            const int q4 = warp & 3;
            const int row = q4 * 32 + lane;
            const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
            // scale-factor-B: 2^0 everywhere
#pragma unroll
            for (int cidx = 0; cidx < 4; ++cidx) tmem_st_x1(tmem_sfb + lane_addr + cidx, 0x7F7F7F7Fu);

            // Scale words of a tile: loaded from global memory early (load_scales, results are not
            // consumed until store_scales) so that the L2/HBM latency never sits on the per-tile chain.
            auto load_scales = [&](int seg, int tile_in_seg, uint4 (&pk)[4]) {
                // 16 words per tile, packed on the host in TMEM order (models/glm.py: pack_tile_scales):
                //   words 0-7  (MMA #1): word 4g + q  = scales of (row group q, feature blocks 4g .. 4g+3)
                //   words 8-15 (MMA #2): word 4h + qq = scales of (row groups 0..3, feature block 4h + qq)
                const long long seg_tiles = (segs_g[seg].n_rows + kTile - 1) / kTile;
                if (tile_in_seg < seg_tiles) {
                    const uint4* sp = reinterpret_cast<const uint4*>(segs_g[seg].scales) + (size_t)tile_in_seg * 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i) pk[i] = __ldg(sp + i);
                } else {   // an empty tile that pads a chunk: 2^0 everywhere (its rows are all zero)
#pragma unroll
                    for (int i = 0; i < 4; ++i) pk[i] = make_uint4(0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0x7F7F7F7Fu);
                }
            };
            // writes the scale words of the CTA's t_it-th tile into ring slot t_it % kSfRing
            auto store_scales = [&](int t_it, const uint4 (&pk)[4]) {
                const uint32_t sfa1[8] = {pk[0].x, pk[0].y, pk[0].z, pk[0].w, pk[1].x, pk[1].y, pk[1].z, pk[1].w};
                const uint32_t sfa2[8] = {pk[2].x, pk[2].y, pk[2].z, pk[2].w, pk[3].x, pk[3].y, pk[3].z, pk[3].w};
                const int slot = t_it % kSfRing;
                tmem_st_x8(tmem_sfa1 + lane_addr + slot * 8, sfa1);
                tmem_st_x8(tmem_sfa2 + lane_addr + slot * 8, sfa2);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&bar_sf_full[slot]);
            };

            const int ew = eg * 4 + q4;          // epilogue warp ordinal: row of the per-warp slots
            const int b = eg;                    // eta / R buffer of this group
            int it = 0;                          // tiles this CTA has been handed before the current chunk (multiple of kEG)
            bool stored_ahead = false;           // the scales of this group's first tile of the chunk are already in TMEM
            Ring gbuf;
            for (int j = 0;; ++j) {
                const int4 ch = next_chunk(j);
                if (ch.x < 0) break;
                const float* __restrict__ seg_y = segs_g[ch.x].y;
                const long long seg_rows = segs_g[ch.x].n_rows;
                const int seg_group = segs_g[ch.x].group;
                const int og = segs_g[ch.x].out_group;   // output block of this chunk's segment
                // this warp's (hi, lo) slots: LL at +0, intercept gradient g at +(1 + g); fetched a whole chunk early
                double* slot0 = ll_slots + 2 * ((((size_t)ew * NOUT + og) * KF) * NS1);
                double2 pre_l[KF], pre_g[KF];
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < KF; ++k)
                        if (k < nch) {
                            pre_l[k] = *reinterpret_cast<const double2*>(slot0 + 2 * (size_t)k * NS1);
                            pre_g[k] = *reinterpret_cast<const double2*>(slot0 + 2 * (size_t)k * NS1 + 2 * (1 + seg_group));
                        }
                }
                if (!stored_ahead) {
                    uint4 first[4];
                    load_scales(ch.x, ch.y + eg, first);
                    store_scales(it + eg, first);
                }
                stored_ahead = false;
                float ll_acc[KF], gi_cur[KF];
#pragma unroll
                for (int k = 0; k < KF; ++k) ll_acc[k] = gi_cur[k] = 0.f;
                for (int t = eg; t < ch.z; t += kEG) {
                    const int itl = it + t;
                    const long long grow = ((long long)ch.y + t) * kTile + row;
                    const bool valid = grow < seg_rows;
                    const float y = valid ? __ldg(seg_y + grow) : 0.f;
                    const uint32_t bph = (uint32_t)((itl / kEG) & 1);
                    // scale words of this group's NEXT tile (consumed at the end of this iteration): the next tile
                    // of the chunk, or the group's first tile of the next chunk if that chunk is already known
                    uint4 next_scales[4];
                    int next_it = -1;
                    if (t + kEG < ch.z) {
                        load_scales(ch.x, ch.y + t + kEG, next_scales);
                        next_it = itl + kEG;
                    } else {
                        const int4 nx = peek_chunk(j + 1);
                        if (nx.x >= 0) {
                            load_scales(nx.x, nx.y + eg, next_scales);
                            next_it = it + ch.z + eg;
                            stored_ahead = true;
                        }
                    }

                    mbar_wait(&bar_eta_full[b], bph);
                    tc_fence_after();
                    float v[16];
                    tmem_ld_x16(tmem_eta + lane_addr + b * kN, v);
                    tc_fence_before();
                    mbar_arrive(&bar_eta_empty[b]);
                    uint32_t rwords[4] = {0, 0, 0, 0};  // 16 residual bytes of this row: 4 terms per chain
                    uint8_t* expo_slot = r_expo + (eg * 2 + (int)bph) * 12;
#pragma unroll
                    for (int k = 0; k < KF; ++k) {
                        const float* vk = v + kThetaTerms * k;
                        const float eta = c_theta[k] * (vk[0] + vk[1] * (1.f / 16) + vk[2] * (1.f / 256) + vk[3] * (1.f / 4096) +
                                                        vk[4] * (1.f / 65536)) + theta_f[k * PG + seg_group];
                        float ll = 0.f, r = 0.f;
                        if (valid && k < nch) link_loglik(DYN ? prm.family : 0, y, eta, ll, r);
                        ll_acc[k] += ll;
                        gi_cur[k] += r;
                        if constexpr (DYN) {
                            // this 32-row group's scale: 2^e with max|r| / 2^e in [0.5, 1)
                            float m = fabsf(r);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                            int e = 0;
                            if (m > 0.f) frexpf(m, &e);
                            e = max(-126, min(126, e));
                            r *= __uint_as_float((uint32_t)(127 - e) << 23);
                            if (lane == 0) expo_slot[k * 4 + q4] = (uint8_t)(127 + e);
                        }
                        uint8_t rt[kResidTerms];
                        expand16<kResidTerms>(r * 256.f, rt);
                        rwords[k] = (uint32_t)rt[0] | ((uint32_t)rt[1] << 8) | ((uint32_t)rt[2] << 16) | ((uint32_t)rt[3] << 24);
                    }
                    mbar_wait(&bar_r_empty[b], bph ^ 1);
                    *reinterpret_cast<uint4*>(r_buf + b * L.r_bytes + (row >> 3) * 128 + (row & 7) * 16) =
                        make_uint4(rwords[0], rwords[1], rwords[2], rwords[3]);
                    fence_proxy_async();
                    if constexpr (DYN) {
                        // all four row-group exponents of this tile -> one word per chain; B row n = 4 * chain + term
                        // reads it from lane n of every subpartition
                        if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                        else if (eg == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
                        else asm volatile("bar.sync 3, 128;" ::: "memory");
                        const int chn = min(lane >> 2, KF - 1);
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(expo_slot + chn * 4);
                        tmem_st_x1(tmem_sfb_r + lane_addr + b * 4, w);
                        tmem_wait_st();
                        tc_fence_before();
                    }
                    mbar_arrive(&bar_r_full[b]);

                    // scales for this group's next tile: its ring slot was last used by the tile kSfRing before it,
                    // whose MMAs are complete (we just passed r_empty of the previous own tile and eta_full of this one)
                    if (next_it >= 0) store_scales(next_it, next_scales);
                }
                // ---- end of the chunk for this group: per-thread fp32 sums -> fixed butterfly over the warp
                // (double) -> lane 0 adds the warp's value to its own (hi, lo) slot; all chunk-determined.  The slot
                // values were fetched at the start of the chunk (`pre`), so no L2 round trip sits here.
                double lsum[KF], gsum[KF];
#pragma unroll
                for (int k = 0; k < KF; ++k) {
                    lsum[k] = (double)ll_acc[k];
                    gsum[k] = (double)gi_cur[k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        lsum[k] += __shfl_xor_sync(0xffffffffu, lsum[k], o);
                        gsum[k] += __shfl_xor_sync(0xffffffffu, gsum[k], o);
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < KF; ++k)
                        if (k < nch) {
                            double* slot = slot0 + 2 * (size_t)k * NS1;
                            fed::dd_add(pre_l[k].x, pre_l[k].y, lsum[k], 0.0);
                            fed::dd_add(pre_g[k].x, pre_g[k].y, gsum[k], 0.0);
                            *reinterpret_cast<double2*>(slot) = pre_l[k];
                            *reinterpret_cast<double2*>(slot + 2 * (1 + seg_group)) = pre_g[k];
                        }
                }
                // gradient: chunks hold a multiple of kEG tiles, so the last group always owns a chunk's last tile
                if (eg == kEG - 1) {
                    const int gb = gbuf.idx;
                    mbar_wait(&bar_g_full[gb], gbuf.phase);
                    tc_fence_after();
                    for (int h = 0; h < NH; ++h) {
                        float gv[16];
                        tmem_ld_x16(tmem_g + lane_addr + (gb * NH + h) * kN, gv);
                        if (h == NH - 1) {
                            tc_fence_before();
                            mbar_arrive(&bar_g_empty[gb]);
                        }
#pragma unroll
                        for (int k = 0; k < KF; ++k)
                            if (k < nch) {
                                double* slot = out + 2 * (((size_t)og * nch + k) * NV1 + 1 + G + h * 128 + row);
                                double2 cur = *reinterpret_cast<double2*>(slot);
                                fed::dd_add(cur.x, cur.y,
                                            (double)kResidNorm * ((double)gv[4 * k] + (double)gv[4 * k + 1] * (1.0 / 16) +
                                                                  (double)gv[4 * k + 2] * (1.0 / 256) + (double)gv[4 * k + 3] * (1.0 / 4096)),
                                            0.0);
                                *reinterpret_cast<double2*>(slot) = cur;
                            }
                    }
                }
                gbuf.advance(2);
                it += ch.z;
            }
        }

        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        fed::pdl_trigger();   // the next evaluation's CTA may take this SM as soon as we exit
        // layout per chain: [LL, gi[G], g[P]] as (hi, lo) pairs; g[] was accumulated in place, LL and gi[] are the
        // per-warp slots summed in warp order
        for (int i = threadIdx.x; i < NOUT * nch * NS1; i += blockDim.x) {
            const int jv = i % NS1, k = (i / NS1) % nch, o = i / (NS1 * nch);
            double hi = 0.0, lo = 0.0;
            for (int w = 0; w < kLLRowsF; ++w) {
                const double* slot = ll_slots + 2 * ((((size_t)w * NOUT + o) * KF + k) * NS1 + jv);
                fed::dd_add(hi, lo, slot[0], slot[1]);
            }
            out[2 * (((size_t)o * nch + k) * NV1 + jv)] = hi;
            out[2 * (((size_t)o * nch + k) * NV1 + jv) + 1] = lo;
        }
    } else if (warp == 0) {
        // nothing will be computed (stop / idle): the early loads still have to land before the CTA may exit
        for (int t = 0; t < preloaded; ++t) mbar_wait(&bar_full[t], 0u);
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
    const bool fin = fed::epilogue_t<true>(comm, pro, (active && *pipeline_fault()) ? B200FED_ERR_PIPELINE : 0ull, row_doubles,
                                           comm.group_partials);
    if (fin && threadIdx.x == 0) *work_counter = 0u;   // every CTA has stopped claiming: ready for the next launch
}

}  // namespace fp8

namespace {
typedef CUresult (*EncodeTiledFn8)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn8 get_encode8() {
    static EncodeTiledFn8 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn8>(p);
    }
    return fn;
}
}  // namespace

extern "C" int b200_glm_fp8_prepare(const GlmSegment* segs_host, int n_segments, const GlmParams* prm, int sm_count,
                                    void** tmaps_dev, void** chunks_dev, int* n_chunks) {
    if (prm->n_features != 128 && prm->n_features != 256) return -21;   // NFB <= 8 scale columns per tile
    for (int s = 0; s < n_segments; ++s)
        if (segs_host[s].n_rows + fp8::kTile * fp8::kEG >= (1ll << 31)) return -22;   // row coordinates are 32-bit
    if (prm->n_chains < 1 || prm->n_chains > 3 || prm->family < 0 || prm->family > 2) return -23;
    if (prm->ld % 16 != 0) return -24;
    EncodeTiledFn8 encode = get_encode8();
    if (!encode) return -25;
    CUtensorMap* host = new CUtensorMap[n_segments];
    for (int s = 0; s < n_segments; ++s) {
        if (((uintptr_t)segs_host[s].X & 15) != 0 || segs_host[s].scales == nullptr) { delete[] host; return -26; }
        cuuint64_t dims[2] = {(cuuint64_t)prm->n_features, (cuuint64_t)segs_host[s].n_rows};
        cuuint64_t strides[1] = {(cuuint64_t)prm->ld};
        cuuint32_t box[2] = {(cuuint32_t)fp8::kPanelF, (cuuint32_t)fp8::kTile};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&host[s], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(segs_host[s].X), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete[] host; return -27; }
    }
    if (*tmaps_dev) cudaFree(*tmaps_dev);
    cudaError_t e = cudaMalloc(tmaps_dev, sizeof(CUtensorMap) * n_segments);
    if (e == cudaSuccess) e = cudaMemcpy(*tmaps_dev, host, sizeof(CUtensorMap) * n_segments, cudaMemcpyHostToDevice);
    delete[] host;
    if (e != cudaSuccess) return (int)e;
    const std::vector<GlmChunk> chunks = build_chunks(segs_host, n_segments, sm_count > 0 ? sm_count : 148, fp8::kTile, fp8::kEG,
                                                        fp8::kMaxChunkF, fp8::kMinChunkF);
    if (*chunks_dev) cudaFree(*chunks_dev);
    e = cudaMalloc(chunks_dev, sizeof(GlmChunk) * (chunks.size() + 1));
    if (e == cudaSuccess) e = cudaMemcpy(*chunks_dev, chunks.data(), sizeof(GlmChunk) * chunks.size(), cudaMemcpyHostToDevice);
    *n_chunks = (int)chunks.size();
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" size_t b200_glm_fp8_partial_row_doubles(int n_vals, int n_chains, int n_out, int n_groups) {
    return fp8::partial_row_doubles(n_vals, n_chains == 1 ? 1 : 3, n_out > 0 ? n_out : 1, n_groups);
}

extern "C" int b200_launch_glm_fp8(const FedComm* comm, const GlmSegment* segs_dev, const GlmParams* prm, const void* tmaps,
                                   const void* chunks_dev, int n_chunks, unsigned int* work_counter, int grid,
                                   cudaStream_t stream) {
    const fp8::SmemLayoutF L = fp8::smem_layout(prm->n_features, comm->n_theta, prm->n_groups);
    if (L.stages < 2) return -2;
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(tmaps);
#define B200FED_FP8_LAUNCH(KF, DYN)                                                                                       \
    do {                                                                                                                 \
        cudaFuncSetAttribute(fp8::fed_glm_fp8_kernel<KF, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total); \
        cudaLaunchConfig_t cfg{};                                                                                        \
        cfg.gridDim = dim3(grid);                                                                                        \
        cfg.blockDim = dim3(fp8::kThreadsF);                                                                             \
        cfg.dynamicSmemBytes = L.total;                                                                                  \
        cfg.stream = stream;                                                                                             \
        cudaLaunchAttribute attr[1];                                                                                     \
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                 \
        attr[0].val.programmaticStreamSerializationAllowed = 1;                                                          \
        cfg.attrs = attr;                                                                                                \
        cfg.numAttrs = tc::use_pdl() ? 1 : 0;                                                                            \
        cudaLaunchKernelEx(&cfg, fp8::fed_glm_fp8_kernel<KF, DYN>, *comm, segs_dev, *prm, maps,                          \
                           reinterpret_cast<const GlmChunk*>(chunks_dev), n_chunks, work_counter);                       \
    } while (0)
    const bool dyn = prm->family != 0;  // unbounded residuals: per-row-group scales for the R operand
    if (prm->n_chains == 1) {
        if (dyn) B200FED_FP8_LAUNCH(1, true); else B200FED_FP8_LAUNCH(1, false);
    } else {
        if (dyn) B200FED_FP8_LAUNCH(3, true); else B200FED_FP8_LAUNCH(3, false);
    }
#undef B200FED_FP8_LAUNCH
    return (int)cudaGetLastError();
}
