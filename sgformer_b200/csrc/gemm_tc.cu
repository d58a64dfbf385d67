// Dense contractions of the SGFormer encoder on tcgen05 tensor cores (sm_100a).
//
//  gemm_nt : out[rows, n_out] = epilogue( sum_seg A_seg[rows, k] . B_seg[n_out, k]^T )      "row-streaming"
//            rows = nodes (huge), n_out, k <= a few hundred.  Replaces nn.Linear forward / input-gradient,
//            q~.(K^T V) and the attention backward products (reference medium/ours.py:22,28,76-85; large/ours.py:38-40,
//            123-128,141,199,275).  Both operands K-major bf16, TMA SWIZZLE_128B tiles, BM=128 x BN<=256(+16) x BK=64,
//            4-stage smem ring, fp32 accumulators in TMEM (double-buffered when BN<=256), warp-specialised:
//            warp0 = TMA producer, warp1 = MMA issuer + TMEM owner, warps 2-5 = epilogue (one TMEM lane quadrant each).
//  gemm_tn : out[m, n] = alpha * sum_rows A[rows, m]^T B[rows, n]                               "node-contracting"
//            Replaces K^T V (medium/ours.py:21), its backward q^T gnum and every weight gradient dW = dY^T X.
//            Both operands are MN-major for the MMA (features contiguous), loaded as [64 feat x 64 node] boxes;
//            the node range is split across CTAs, per-CTA partials go to a workspace and are reduced deterministically.
//
// Descriptor formats follow cute/arch/mma_sm100_desc.hpp and cute/atom/mma_traits_sm100.hpp (canonical SW128 layouts).
#include "common.cuh"
#include "launch_count.h"
#include "../../include/sgformer_b200.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace sgf {

// ------------------------------------------------------------------------------------------------
// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    });
    return fn;
}

// 2-D row-major [rows, cols] (pitch ld elements) of bf16 (dtype 1) or fp32 (dtype 0); box = [box_rows x 128 bytes],
// 128-byte swizzle, OOB loads -> 0, OOB stores clipped
static int make_tmap_2d(CUtensorMap* tm, const void* base, int dtype, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return SGF_ERR_DRIVER;
    const int es = dtype == 1 ? 2 : 4;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * es) % 16 != 0 || rows <= 0 || cols <= 0 || box_rows <= 0 || box_rows > 256)
        return SGF_ERR_ARG;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(tm, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SGF_OK : SGF_ERR_DRIVER;
}
static int make_tmap_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    return make_tmap_2d(tm, base, 1, rows, cols, ld, box_rows);
}

// ================================================================================================
// gemm_nt
// ================================================================================================
namespace nt {
constexpr int BM = 128, BK = 64, STAGES = 3;
constexpr int EPI_WARPS = 8;                      // two warps per TMEM lane quadrant
constexpr int THREADS = 32 * (2 + EPI_WARPS);     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES_MAX = (256 + 16) * BK * 2;  // 34 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int STG_BYTES = BM * 128;               // output staging: 128 rows x 128 bytes (SW128) = one 4 KB tile per epilogue warp ...
constexpr int WSTG_BYTES = 32 * 128;              // ... of 32 rows (its TMEM lane quadrant), stored by the warp's own TMA store
constexpr int BAR_BYTES = 512;
constexpr int STAT_COLS = 1024;                   // fused column statistics cover n_out <= 1024
constexpr int STAT_BYTES = 2 * STAT_COLS * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * STG_BYTES + BAR_BYTES + STAT_BYTES + 1024;
// Resident-B mode (p.b_res): the whole B operand of one n-block (all k-blocks) stays in shared memory across the row tiles
// of a CTA, so that only A streams through the ring.  Without it every 128-row tile re-fetches K x n_block of weights from
// L2 and the kernel runs into the L2->SM throughput cap (~12 TB/s) before the HBM roofline (ncu, profiles/r1b_*).
constexpr int RES_BYTES = (256 + 16) * 256 * 2;   // 136 KB: K = 256 x (256 + 16-column tail)
constexpr int RES_MAX_KB = 16;                    // one full/empty mbarrier pair per resident k-block
constexpr int RES_STAGES = 3;                     // A-only ring
constexpr int RES_STG = 2;                        // one staging tile per epilogue half
constexpr int SMEM_BYTES_RES = RES_BYTES + RES_STAGES * A_BYTES + RES_STG * STG_BYTES + BAR_BYTES + STAT_BYTES + 1024;
static_assert(SMEM_BYTES <= 232448 && SMEM_BYTES_RES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

struct Seg {
    int a_idx, a_koff, b_idx, b_koff, k_blocks;
};
struct Params {
    int64_t rows;
    int n_out, bn_main, has_tail, n_blocks;
    int64_t num_tiles;
    int b_res;          // 1: resident-B schedule (see RES_BYTES)
    int64_t chunk;      // b_res: row tiles per CTA between two reloads of B (only matters when n_blocks > 1)
    int n_seg, total_kb;
    Seg seg[SGF_MAX_SEG];
    int epi;
    void* out; int64_t ldo; int out_dtype;
    const float* bias;
    const void* aux; int64_t ld_aux; int aux_dtype;
    const float* row_scale;
    float alpha, beta;
    const float* alpha_dev; const float* beta_dev;
    int relu, accumulate;
    float nf; const float* nf_dev; float* den_out;
    const float* r1_row; const float* r1_col;
    int tma_store;   // 1: epilogue stages 128-byte rows in smem and stores them with TMA; 0: direct global stores
    float* col_sum; float* col_sumsq;   // optional fused column statistics of the STORED output (tma_store path only)
};
struct Tmaps {
    CUtensorMap a[SGF_MAX_SRC];
    CUtensorMap b[SGF_MAX_SRC];
    CUtensorMap tail;
    CUtensorMap out;
};

__device__ __forceinline__ float load1(const void* base, int dtype, int64_t off) {
    return dtype == 1 ? __bfloat162float(static_cast<const __nv_bfloat16*>(base)[off]) : static_cast<const float*>(base)[off];
}
__device__ __forceinline__ void store1(void* base, int dtype, int64_t off, float v) {
    if (dtype == 1) static_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(v);
    else static_cast<float*>(base)[off] = v;
}
// 32 consecutive elements at element offset `off` (vec: 16-byte aligned and all 32 valid; else the first nv, rest 0)
__device__ __forceinline__ void load32(const void* base, int dtype, int64_t off, bool vec, int nv, float* f) {
    if (vec) {
        if (dtype == 1) {
            const uint4* q = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + off);
#pragma unroll
            for (int i = 0; i < 4; ++i) Vec16<__nv_bfloat16>::unpack(q[i], f + 8 * i);
        } else {
            const uint4* q = reinterpret_cast<const uint4*>(static_cast<const float*>(base) + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) Vec16<float>::unpack(q[i], f + 4 * i);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = j < nv ? load1(base, dtype, off + j) : 0.f;
    }
}
__device__ __forceinline__ void store32(void* base, int dtype, int64_t off, bool vec, int nv, const float* f) {
    if (vec) {
        if (dtype == 1) {
            uint4* q = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + off);
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = Vec16<__nv_bfloat16>::pack(f + 8 * i);
        } else {
            uint4* q = reinterpret_cast<uint4*>(static_cast<float*>(base) + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = Vec16<float>::pack(f + 4 * i);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < nv) store1(base, dtype, off + j, f[j]);
    }
}
// The three warp roles walk the CTA's tiles in the same order.
//  streaming:  tile = blockIdx.x + i*gridDim.x over (m_blk, n_blk) pairs, n_blk fastest.
//  resident-B: the CTA owns row tiles m = blockIdx.x + i*gridDim.x; they are visited in chunks of p.chunk, and inside a chunk
//              n_blk is the OUTER loop: B(n_blk) is loaded once per (chunk, n_blk) "group", the chunk's A tiles are re-read
//              from L2 for the following n-blocks (a chunk is sized to stay L2-resident).
struct TileIter {
    const Params& p;
    bool started = false;
    int64_t tile = 0;                 // streaming
    int64_t cnt = 0, c0 = 0, i = 0;   // resident
    int nb = 0;
    int64_t group = 0;
    int m_blk = 0, n_blk = 0;
    bool first = false, last = false;   // first / last tile of its group (resident)
    __device__ explicit TileIter(const Params& pp) : p(pp) {
        if (p.b_res) {
            const int64_t m_tiles = p.num_tiles / p.n_blocks;
            cnt = m_tiles > blockIdx.x ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        }
    }
    __device__ bool next() {
        if (!p.b_res) {
            tile = started ? tile + gridDim.x : blockIdx.x;
            started = true;
            if (tile >= p.num_tiles) return false;
            m_blk = (int)(tile / p.n_blocks);
            n_blk = (int)(tile % p.n_blocks);
            return true;
        }
        if (!started) {
            started = true;
            if (cnt == 0) return false;
        } else {
            ++i;
            const int64_t cend = c0 + p.chunk < cnt ? c0 + p.chunk : cnt;
            if (i >= cend) {
                ++group;
                if (++nb >= p.n_blocks) {
                    nb = 0;
                    c0 += p.chunk;
                    if (c0 >= cnt) return false;
                }
                i = c0;
            }
        }
        const int64_t cend = c0 + p.chunk < cnt ? c0 + p.chunk : cnt;
        m_blk = (int)(blockIdx.x + i * gridDim.x);
        n_blk = nb;
        first = i == c0;
        last = i == cend - 1;
        return true;
    }
};

// Epilogue feature bits.  A specialised instantiation gemm_nt_kernel<F> compiles exactly the features in F (present and
// unconditional: bf16 output through the TMA-store path, every 32-column piece complete, bf16 16-byte-aligned addends); the
// F_GENERIC instantiation reads every feature from Params at run time and handles ragged widths, fp32 and unaligned tensors.
enum : int { F_BIAS = 1, F_AUX = 2, F_RELU = 4, F_ROWSCALE = 8, F_ACCUM = 16, F_R1 = 32, F_ATTN = 64, F_GENERIC = 128 };

template <int F>
__global__ void __launch_bounds__(THREADS, 1) gemm_nt_kernel(const __grid_constant__ Tmaps tm, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    // streaming: [ring: STAGES x (A | B)] [staging 4 tiles];  resident: [B: RES_BYTES] [ring: RES_STAGES x A] [staging 2 tiles]
    const bool res = p.b_res != 0;
    uint8_t* bres = smem;
    uint8_t* ring = res ? smem + RES_BYTES : smem;
    const int ring_stages = res ? RES_STAGES : STAGES;
    const int ring_stride = res ? A_BYTES : STAGE_BYTES;
    const int stg_per_warp = res ? 1 : 2;                 // warp-private [32 rows x 128 B] staging tiles
    uint8_t* staging = ring + ring_stages * ring_stride;
    uint64_t* full = reinterpret_cast<uint64_t*>(staging + EPI_WARPS * stg_per_warp * WSTG_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* bres_full = tmem_empty + 2;
    uint64_t* bres_empty = bres_full + RES_MAX_KB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_empty + RES_MAX_KB);
    float* stat_sm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + BAR_BYTES);   // [2][STAT_COLS]: sum, sumsq
    static_assert((2 * STAGES + 4 + 2 * RES_MAX_KB) * 8 + 4 <= BAR_BYTES, "barrier area");
    static_assert(RES_STAGES <= STAGES, "ring barriers");
    const bool want_stats = p.col_sum != nullptr || p.col_sumsq != nullptr;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int bn_total = p.bn_main + (p.has_tail ? 16 : 0);
    const int acc_stages = bn_total <= 256 ? 2 : 1;
    if (want_stats)
        for (int i = threadIdx.x; i < 2 * STAT_COLS; i += THREADS) stat_sm[i] = 0.f;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.n_seg; ++i) {
            tma_prefetch_desc(&tm.a[p.seg[i].a_idx]);
            tma_prefetch_desc(&tm.b[p.seg[i].b_idx]);
        }
        if (p.has_tail) tma_prefetch_desc(&tm.tail);
        if (p.tma_store) tma_prefetch_desc(&tm.out);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI_WARPS); }
        for (int s = 0; s < RES_MAX_KB; ++s) { mbar_init(&bres_full[s], 1); mbar_init(&bres_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const uint32_t b_tx = p.bn_main * BK * 2 + (p.has_tail ? 16 * BK * 2 : 0);
            int stage = 0; uint32_t phase = 0;
            TileIter ti(p);
            while (ti.next()) {
                int kbg = 0;
                for (int s = 0; s < p.n_seg; ++s) {
                    const Seg sg = p.seg[s];
                    for (int kb = 0; kb < sg.k_blocks; ++kb, ++kbg) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = ring + stage * ring_stride;
                        if (!res) {
                            uint8_t* sb = sa + A_BYTES;
                            mbar_arrive_expect_tx(&full[stage], A_BYTES + b_tx);
                            tma_load_2d(sa, &tm.a[sg.a_idx], &full[stage], sg.a_koff + kb * BK, ti.m_blk * BM);
                            tma_load_2d(sb, &tm.b[sg.b_idx], &full[stage], sg.b_koff + kb * BK, ti.n_blk * p.bn_main);
                            if (p.has_tail) tma_load_2d(sb + p.bn_main * BK * 2, &tm.tail, &full[stage], sg.b_koff + kb * BK, 0);
                        } else {
                            // A first (its ring slot frees early), then - on the first tile of a group - the k-block of B,
                            // as soon as the previous group's last tile has consumed the old one
                            mbar_arrive_expect_tx(&full[stage], A_BYTES);
                            tma_load_2d(sa, &tm.a[sg.a_idx], &full[stage], sg.a_koff + kb * BK, ti.m_blk * BM);
                            if (ti.first) {
                                mbar_wait(&bres_empty[kbg], (uint32_t)(ti.group & 1) ^ 1);
                                uint8_t* sb = bres + (size_t)kbg * b_tx;
                                mbar_arrive_expect_tx(&bres_full[kbg], b_tx);
                                tma_load_2d(sb, &tm.b[sg.b_idx], &bres_full[kbg], sg.b_koff + kb * BK, ti.n_blk * p.bn_main);
                                if (p.has_tail) tma_load_2d(sb + p.bn_main * BK * 2, &tm.tail, &bres_full[kbg], sg.b_koff + kb * BK, 0);
                            }
                        }
                        if (++stage == ring_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc_main = make_idesc_bf16(BM, p.bn_main, 0, 0);
            const uint32_t idesc_tail = make_idesc_bf16(BM, 16, 0, 0);
            const uint32_t b_tx = p.bn_main * BK * 2 + (p.has_tail ? 16 * BK * 2 : 0);
            int stage = 0; uint32_t phase = 0;
            int64_t it = 0;
            TileIter ti(p);
            for (; ti.next(); ++it) {
                const int acc = (int)(it % acc_stages);
                const uint32_t acc_phase = (uint32_t)((it / acc_stages) & 1);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_main = tmem_base + acc * 256;
                for (int kbg = 0; kbg < p.total_kb; ++kbg) {
                    if (res && ti.first) mbar_wait(&bres_full[kbg], (uint32_t)(ti.group & 1));
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(ring + stage * ring_stride);
                    const uint32_t sb = res ? smem_u32(bres) + (uint32_t)kbg * b_tx : sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(sa + k * 32, 0, 1024);
                        const uint64_t db = make_smem_desc_sw128(sb + k * 32, 0, 1024);
                        const uint32_t accum = (kbg > 0 || k > 0) ? 1u : 0u;
                        umma_bf16(d_main, da, db, idesc_main, accum);
                        if (p.has_tail) {
                            const uint64_t dt = make_smem_desc_sw128(sb + p.bn_main * BK * 2 + k * 32, 0, 1024);
                            umma_bf16(d_main + p.bn_main, da, dt, idesc_tail, accum);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (res && ti.last) umma_commit(&bres_empty[kbg]);
                    if (kbg == p.total_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == ring_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (warp-private smem staging -> TMA store | direct global stores) =====
        // Each warp owns the 32 rows of its TMEM lane quadrant and every other 128-byte column group (two warps per
        // quadrant), stages a [32 x 128 B] tile and stores it with its own TMA store: no cross-warp synchronisation.
        constexpr bool GEN = (F & F_GENERIC) != 0;
        const bool f_bias = GEN ? p.bias != nullptr : (F & F_BIAS) != 0;
        const bool f_aux = GEN ? p.aux != nullptr : (F & F_AUX) != 0;
        const bool f_relu = GEN ? p.relu != 0 : (F & F_RELU) != 0;
        const bool f_rs = GEN ? p.row_scale != nullptr : (F & F_ROWSCALE) != 0;
        const bool f_acc = GEN ? p.accumulate != 0 : (F & F_ACCUM) != 0;
        const bool f_r1 = GEN ? p.r1_row != nullptr : (F & F_R1) != 0;
        const bool f_attn = GEN ? (p.epi == SGF_EPI_ATTN_APPLY || p.epi == SGF_EPI_ATTN_GRAM) : (F & F_ATTN) != 0;
        const bool f_tma = GEN ? p.tma_store != 0 : true;
        const bool f_stats = GEN ? want_stats : false;
        const int out_dtype = GEN ? p.out_dtype : 1;
        const int ew = warp - 2;     // 0..7
        const int q = warp & 3;      // TMEM lane quadrant this warp may access
        const int half = ew >> 2;    // the two warps of a quadrant alternate over column groups
        float alpha = p.alpha, beta = p.beta;
        if (p.alpha_dev) alpha *= *p.alpha_dev;
        if (p.beta_dev) beta *= *p.beta_dev;
        const int out_es = out_dtype == 1 ? 2 : 4, aux_es = p.aux_dtype == 1 ? 2 : 4;
        const bool vec_ok = !GEN || (((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && ((p.ldo * out_es) % 16 == 0));
        const bool aux_vec_ok = !GEN || !p.aux || (((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) && ((p.ld_aux * aux_es) % 16 == 0));
        const bool vecf_ok = !GEN || ((!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) &&
                                      (!p.r1_col || (reinterpret_cast<uintptr_t>(p.r1_col) & 15) == 0));
        const int GW = out_dtype == 1 ? 64 : 32;              // columns per 128-byte staging row
        uint8_t* wstg = staging + ew * stg_per_warp * WSTG_BYTES;
        const int r_local = q * 32 + lane;
        // The addend of the epilogue (aux, else the old output when accumulating) is fetched one 32-column piece AHEAD as
        // packed 16-byte chunks: a thread owns a whole row, so an un-prefetched load costs a DRAM latency per piece.
        const void* pre_src = f_aux ? p.aux : (f_acc ? p.out : nullptr);
        const int64_t pre_ld = f_aux ? p.ld_aux : p.ldo;
        const bool pre_on = GEN ? (pre_src && (f_aux ? p.aux_dtype == 1 && aux_vec_ok : out_dtype == 1 && vec_ok)) : (f_aux || f_acc);
        uint4 pre[4];
        bool pre_ok = false;     // pre[] holds the addend of the NEXT piece to be processed
        auto prefetch = [&](int64_t row, bool row_ok, int col0) {
            pre_ok = pre_on && row_ok && (!GEN || col0 + 32 <= p.n_out);
            if (pre_ok) {
                const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(pre_src) + row * pre_ld + col0);
#pragma unroll
                // L1-allocating loads: a thread reads its row 64 bytes at a time, the second half of each 128-byte line
                // must hit L1 (ld.global.nc.L1::no_allocate here made the addend GEMMs 1.5x slower than r1b)
                for (int c = 0; c < 4; ++c) pre[c] = f_aux ? __ldg(src + c) : src[c];
            }
        };
        const int n_pieces = (p.bn_main + 31) / 32;
        const int ppg = GW / 32;                                   // pieces per staging group (bf16: 2, fp32: 1)
        const int n_groups = (n_pieces + ppg - 1) / ppg;
        uint32_t gcount = 0;
        int64_t it = 0;
        TileIter ti(p);
        for (; ti.next(); ++it) {
            const int m_blk = ti.m_blk, n_blk = ti.n_blk;
            const int acc = (int)(it % acc_stages);
            const uint32_t acc_phase = (uint32_t)((it / acc_stages) & 1);
            const int64_t row = (int64_t)m_blk * BM + r_local;
            const bool row_ok = row < p.rows;
            const int col_base = n_blk * p.bn_main;
            if (half < n_groups) prefetch(row, row_ok, col_base + half * ppg * 32);   // overlaps the wait for the accumulator
            const float rs = (f_rs && row_ok) ? p.row_scale[row] : 1.f;
            const float r1r = (f_r1 && row_ok) ? p.r1_row[row] : 0.f;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
            float inv_den = 1.f;
            const float nfv = (f_attn && p.nf_dev) ? *p.nf_dev : p.nf;
            if (f_attn) {
                float t[16];
                __syncwarp();
                tmem_ld16(taddr + p.bn_main, t);
                tmem_ld_wait();
                const float den = t[0] + nfv;
                inv_den = 1.f / den;
                if (row_ok && p.den_out && half == 0 && n_blk == 0) p.den_out[row] = den;
            }
            // -------- 32-column pieces; piece index pc covers tile columns [32*pc, 32*pc+32) --------
            for (int g = half; g < n_groups; g += 2) {
                const uint32_t buf = stg_per_warp == 2 ? (gcount & 1) : 0;
                if (f_tma) {
                    if (lane == 0) {                                   // the store that last used this buffer has drained
                        if (stg_per_warp == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
                    }
                    __syncwarp();
                }
                for (int pp = 0; pp < ppg; ++pp) {
                    const int pc = g * ppg + pp;
                    if (pc >= n_pieces) break;
                    // take this piece's prefetched addend (kept packed), then put the next piece's in flight
                    uint4 cur[4];
                    const bool have_ad = pre_ok;
                    if (have_ad) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) cur[c] = pre[c];
                    }
                    {
                        int next_pc = -1;
                        if (pp + 1 < ppg && pc + 1 < n_pieces) next_pc = pc + 1;
                        else if (g + 2 < n_groups) next_pc = (g + 2) * ppg;
                        if (next_pc >= 0) prefetch(row, row_ok, col_base + next_pc * 32); else pre_ok = false;
                    }
                    float v[32];
                    __syncwarp();  // tcgen05.ld is .sync.aligned
                    tmem_ld32(taddr + pc * 32, v);
                    tmem_ld_wait();
                    const int col0 = col_base + pc * 32;
                    const int ncol = GEN ? (p.n_out - col0 < 32 ? p.n_out - col0 : 32) : 32;   // valid columns (may be <= 0)
                    const bool full32 = ncol == 32;
                    if (row_ok && ncol > 0) {
                        if (f_attn) {
                            // out = (acc + nf*aux + bias) / (tail + nf): ATTN_APPLY carries aux (= v), ATTN_GRAM the bias bt
                            if (f_aux) {
                                if (have_ad) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c) {
                                        float a8[8];
                                        Vec16<__nv_bfloat16>::unpack(cur[c], a8);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) v[8 * c + j] += nfv * a8[j];
                                    }
                                } else if constexpr (GEN) {
                                    float ax[32];
                                    load32(p.aux, p.aux_dtype, row * p.ld_aux + col0, full32 && aux_vec_ok, ncol, ax);
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] += nfv * ax[j];
                                }
                            }
                            if (f_bias) {
                                if (!GEN || (full32 && vecf_ok && (col0 & 3) == 0)) {
                                    const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) { float4 b4 = __ldg(bp + j); v[4*j] += b4.x; v[4*j+1] += b4.y; v[4*j+2] += b4.z; v[4*j+3] += b4.w; }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) if (j < ncol) v[j] += p.bias[col0 + j];
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= inv_den;
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= alpha;
                            if (f_aux) {
                                if (have_ad) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c) {
                                        float a8[8];
                                        Vec16<__nv_bfloat16>::unpack(cur[c], a8);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) v[8 * c + j] += beta * a8[j];
                                    }
                                } else if constexpr (GEN) {
                                    float ax[32];
                                    load32(p.aux, p.aux_dtype, row * p.ld_aux + col0, full32 && aux_vec_ok, ncol, ax);
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] += beta * ax[j];
                                }
                            }
                            if (f_bias) {
                                if (!GEN || (full32 && vecf_ok && (col0 & 3) == 0)) {
                                    const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) { float4 b4 = __ldg(bp + j); v[4*j] += b4.x; v[4*j+1] += b4.y; v[4*j+2] += b4.z; v[4*j+3] += b4.w; }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) if (j < ncol) v[j] += p.bias[col0 + j];
                                }
                            }
                            if (f_r1) {
                                if (!GEN || (full32 && vecf_ok && (col0 & 3) == 0)) {
                                    const float4* cp = reinterpret_cast<const float4*>(p.r1_col + col0);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) { float4 c4 = __ldg(cp + j); v[4*j] += r1r * c4.x; v[4*j+1] += r1r * c4.y; v[4*j+2] += r1r * c4.z; v[4*j+3] += r1r * c4.w; }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) if (j < ncol) v[j] += r1r * p.r1_col[col0 + j];
                                }
                            }
                            if (f_relu)
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                            if (f_rs)
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] *= rs;
                        }
                        if (f_acc) {
                            if (have_ad && !f_aux) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    float o8[8];
                                    Vec16<__nv_bfloat16>::unpack(cur[c], o8);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[8 * c + j] += o8[j];
                                }
                            } else if (GEN || f_aux) {
                                float old[32];
                                load32(p.out, out_dtype, row * p.ldo + col0, full32 && vec_ok, ncol, old);
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] += old[j];
                            }
                        }
                        if constexpr (GEN) {
                            if (!f_tma) store32(p.out, out_dtype, row * p.ldo + col0, full32 && vec_ok, ncol, v);
                        }
                    }
                    if (f_tma) {
                        // 128-byte staging row, 16-byte chunks XOR-swizzled with (row & 7) (== TMA SWIZZLE_128B); OOB rows/cols
                        // are clipped by the TMA store, so garbage there is harmless
                        uint8_t* rowp = wstg + buf * WSTG_BYTES + lane * 128;
                        if (out_dtype == 1) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int chunk = (pp * 4 + c) ^ (lane & 7);
                                *reinterpret_cast<uint4*>(rowp + chunk * 16) = Vec16<__nv_bfloat16>::pack(v + 8 * c);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const int chunk = c ^ (lane & 7);
                                *reinterpret_cast<uint4*>(rowp + chunk * 16) = Vec16<float>::pack(v + 4 * c);
                            }
                        }
                    }
                }
                if (f_tma) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tm.out, wstg + buf * WSTG_BYTES, col_base + g * GW, m_blk * BM + q * 32);
                        bulk_commit();
                    }
                    if (f_stats) {
                        // column sums of the staged (already rounded) 32-row tile: lane l owns columns l (and l + 32 for bf16);
                        // the buffer is rewritten only after this warp's next bulk_wait_read + __syncwarp
                        const uint8_t* tile = wstg + buf * WSTG_BYTES;
                        const int64_t valid_rows = p.rows - ((int64_t)m_blk * BM + q * 32);
                        for (int cg = lane; cg < GW; cg += 32) {
                            const int colg = col_base + g * GW + cg;
                            if (colg >= p.n_out || colg >= col_base + p.bn_main) continue;
                            float s1 = 0.f, s2 = 0.f;
                            for (int r = 0; r < 32 && r < valid_rows; ++r) {
                                float x;
                                if (out_dtype == 1) {
                                    const int chunk = (cg >> 3) ^ (r & 7);
                                    x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + r * 128 + chunk * 16 + (cg & 7) * 2));
                                } else {
                                    const int chunk = (cg >> 2) ^ (r & 7);
                                    x = *reinterpret_cast<const float*>(tile + r * 128 + chunk * 16 + (cg & 3) * 4);
                                }
                                s1 += x;
                                s2 += x * x;
                            }
                            atomicAdd(&stat_sm[colg], s1);
                            atomicAdd(&stat_sm[STAT_COLS + colg], s2);
                        }
                    }
                    ++gcount;
                }
            }
            __syncwarp();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (f_tma && lane == 0) bulk_wait<0>();
        if (f_stats) {
            named_bar_sync(3, 32 * EPI_WARPS);
            for (int i = threadIdx.x - 64; i < p.n_out; i += 32 * EPI_WARPS) {
                if (p.col_sum) atomicAdd(&p.col_sum[i], stat_sm[i]);
                if (p.col_sumsq) atomicAdd(&p.col_sumsq[i], stat_sm[STAT_COLS + i]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}
}  // namespace nt

// ================================================================================================
// gemm_tn
// ================================================================================================
namespace tn {
constexpr int BKN = 64, STAGES = 3, THREADS = 192;
constexpr int CHUNK_BYTES = 64 * BKN * 2;   // one [64 feat x 64 node] box = 8 KB
constexpr int A_BYTES = 4 * CHUNK_BYTES;    // up to 256 features
constexpr int B_BYTES = 4 * CHUNK_BYTES;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 64 KB
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;

struct Params {
    int64_t rows;
    int m, n, m_blocks, n_chunks, un;  // un = UMMA N (n rounded up to 16)
    int64_t kb_total;
    float* ws;  // [grid][un][m_blocks*128]
    int n_pairs;                                  // partial products accumulated per node block (1, or the 6 of bf16x3)
    int a_off[SGF_TN_MAX_PAIRS], b_off[SGF_TN_MAX_PAIRS];   // column offset (elements) of the plane each product reads
};
struct Tmaps {
    CUtensorMap a, b;
};

__global__ void __launch_bounds__(THREADS, 1) gemm_tn_kernel(const __grid_constant__ Tmaps tm, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // contiguous slice of 64-row node blocks for this CTA
    const int64_t per = p.kb_total / gridDim.x, rem = p.kb_total % gridDim.x;
    const int64_t kb0 = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
    const int64_t kb1 = kb0 + per + (blockIdx.x < rem ? 1 : 0);
    const int a_chunks = p.m_blocks * 2;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm.a);
        tma_prefetch_desc(&tm.b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t stage_tx = (a_chunks + p.n_chunks) * CHUNK_BYTES;
            int stage = 0; uint32_t phase = 0;
            // all partial products of a node block before the next block: the planes re-read by later products hit L2
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                for (int pr = 0; pr < p.n_pairs; ++pr) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_arrive_expect_tx(&full[stage], stage_tx);
                    for (int c = 0; c < a_chunks; ++c)
                        tma_load_2d(sa + c * CHUNK_BYTES, &tm.a, &full[stage], p.a_off[pr] + c * 64, (int32_t)(kb * BKN));
                    for (int c = 0; c < p.n_chunks; ++c)
                        tma_load_2d(sb + c * CHUNK_BYTES, &tm.b, &full[stage], p.b_off[pr] + c * 64, (int32_t)(kb * BKN));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.un, 1, 1);  // both operands MN-major
            int stage = 0; uint32_t phase = 0;
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                for (int pr = 0; pr < p.n_pairs; ++pr) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BKN / 16; ++k) {
                        // MN-major SW128: LBO = stride between 64-element feature chunks, SBO = stride between 8-row node groups
                        const uint64_t db = make_smem_desc_sw128(sb + k * 2048, CHUNK_BYTES, 1024);
                        const uint32_t accum = (kb > kb0 || pr > 0 || k > 0) ? 1u : 0u;
                        for (int mb = 0; mb < p.m_blocks; ++mb) {
                            const uint64_t da = make_smem_desc_sw128(sa + mb * 2 * CHUNK_BYTES + k * 2048, CHUNK_BYTES, 1024);
                            umma_bf16(tmem_base + mb * 256, da, db, idesc, accum);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (kb == kb1 - 1 && pr == p.n_pairs - 1) umma_commit(tmem_full);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const int q = warp & 3;
        const int mp = p.m_blocks * 128;
        float* ws = p.ws + (int64_t)blockIdx.x * p.un * mp;
        if (kb1 > kb0) {
            mbar_wait(tmem_full, 0);
            tcgen05_fence_after();
        }
        for (int mb = 0; mb < p.m_blocks; ++mb) {
            const int mrow = mb * 128 + q * 32 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mb * 256;
            for (int c = 0; c < p.un / 16; ++c) {
                float v[16];
                if (kb1 > kb0) {
                    __syncwarp();
                    tmem_ld16(taddr + c * 16, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) ws[(int64_t)(c * 16 + j) * mp + mrow] = v[j];
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// out[i,j] = alpha * sum_cta ws[cta][j][i] (+ beta*out[i,j]);  i < m, j < n
__global__ void tn_reduce_kernel(const float* __restrict__ ws, int nparts, int mp, int un, int m, int n, float alpha,
                                 const float* __restrict__ alpha_dev, float beta, float* __restrict__ out, int64_t ldo, int transpose_out) {
    const int64_t total = (int64_t)n * mp;
    const float a = alpha * (alpha_dev ? *alpha_dev : 1.f);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t / mp), i = (int)(t % mp);
        if (i >= m) continue;
        float s = 0.f;
        for (int c = 0; c < nparts; ++c) s += ws[((int64_t)c * un + j) * mp + i];
        float* o = transpose_out ? &out[(int64_t)j * ldo + i] : &out[(int64_t)i * ldo + j];
        *o = a * s + (beta != 0.f ? beta * *o : 0.f);
    }
}
}  // namespace tn
// ================================================================================================
// gram : G[h,h] = X^T X and s[h] = X^T 1 of one bf16 operand X [rows, h <= 256] — pass 1 of the Gram-form attention
// ================================================================================================
// The node-contracting product of gemm_tn specialised for A == B (reference contractions it replaces: k^T v, k^T 1, ||q||^2,
// ||k||^2 of medium/ours.py:16-31 — all functions of X^T X, see csrc/attn_gram.cu):
//  * every [64 feat x 64 node] box is loaded ONCE and feeds both operands of the MMA (A and B descriptors point at the same
//    shared-memory tile), so the kernel moves rows*h*2 bytes for 2*rows*h^2 flops (AI = h flop/B: 256 at h = 256, at the
//    bf16 ridge) instead of twice that;
//  * G is symmetric: only the blocks (0,0), (0,1) and (1,1) of the 2 x 2 block matrix are accumulated
//    (M=128 x N=256 + M=128 x N=128 per k-step: 3/4 of the flops, 384 TMEM columns), the reduction kernel mirrors (0,1);
//  * X^T 1 rides along as an N = 16 MMA against a constant all-ones tile (16 more TMEM columns per row block);
//  * bf16x3 operands (fp32 mode): the six plane pairs of kernels._PAIRS3 accumulate into the same TMEM tiles.
// The node range is split across the CTAs; per-CTA partials -> workspace -> fixed-order reduction (deterministic).
namespace gramk {
constexpr int BKN = 64, THREADS = 192;
constexpr int CHUNK_BYTES = 64 * BKN * 2;        // one [64 feat x 64 node] box = 8 KB
constexpr int ONES_BYTES = CHUNK_BYTES;          // constant tile of bf16 1.0
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BUDGET = 200 * 1024;          // stages * planes * chunks * 8 KB
constexpr int MAX_STAGES = 8;

struct Params {
    int64_t rows, kb_total;
    int h, chunks, m_blocks, un, n_planes, stages;
    int plane_off[3];                             // column offset (elements) of each plane
    int n_pairs, pa[6], pb[6];
    int ncols;                                    // fp32 columns per CTA partial: un + (un - 128 if m_blocks == 2) + m_blocks
    float* ws;                                    // [grid][ncols][128]
};

__global__ void __launch_bounds__(THREADS, 1) gram_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int plane_bytes = p.chunks * CHUNK_BYTES;
    const int stage_bytes = p.n_planes * plane_bytes;
    uint8_t* ones = smem + p.stages * stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(ones + ONES_BYTES);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tmem_full = empty + MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    static_assert((2 * MAX_STAGES + 1) * 8 + 4 <= BAR_BYTES, "barrier area");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t per = p.kb_total / gridDim.x, rem = p.kb_total % gridDim.x;
    const int64_t kb0 = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem);
    const int64_t kb1 = kb0 + per + (blockIdx.x < rem ? 1 : 0);

    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;   // bf16 1.0 x2
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();        // the generic-proxy writes of the ones tile must be visible to the tensor core (async proxy)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: [0, un) block row 0 | [256, 256 + un - 128) block (1,1) | 384 + 16*mb: X^T 1
    constexpr uint32_t COL_D1 = 256, COL_S = 384;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * stage_bytes;
                mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
                for (int pl = 0; pl < p.n_planes; ++pl)
                    for (int c = 0; c < p.chunks; ++c)
                        tma_load_2d(sa + pl * plane_bytes + c * CHUNK_BYTES, &tm, &full[stage], p.plane_off[pl] + c * 64, (int32_t)(kb * BKN));
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc0 = make_idesc_bf16(128, p.un, 1, 1);                                   // both operands MN-major
            const uint32_t idesc1 = make_idesc_bf16(128, p.m_blocks == 2 ? p.un - 128 : 16, 1, 1);
            const uint32_t idescs = make_idesc_bf16(128, 16, 1, 1);
            const uint32_t s_ones = smem_u32(ones);
            int stage = 0; uint32_t phase = 0;
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                for (int pr = 0; pr < p.n_pairs; ++pr) {
                    const uint32_t a0 = sa + p.pa[pr] * plane_bytes, b0 = sa + p.pb[pr] * plane_bytes;
#pragma unroll
                    for (int k = 0; k < BKN / 16; ++k) {
                        // MN-major SW128: LBO = stride between 64-element feature chunks, SBO = stride between 8-row node groups
                        const uint32_t accum = (kb > kb0 || pr > 0 || k > 0) ? 1u : 0u;
                        const uint64_t da0 = make_smem_desc_sw128(a0 + k * 2048, CHUNK_BYTES, 1024);
                        const uint64_t db0 = make_smem_desc_sw128(b0 + k * 2048, CHUNK_BYTES, 1024);
                        umma_bf16(tmem_base, da0, db0, idesc0, accum);
                        if (p.m_blocks == 2) {
                            const uint64_t da1 = make_smem_desc_sw128(a0 + 2 * CHUNK_BYTES + k * 2048, CHUNK_BYTES, 1024);
                            const uint64_t db1 = make_smem_desc_sw128(b0 + 2 * CHUNK_BYTES + k * 2048, CHUNK_BYTES, 1024);
                            umma_bf16(tmem_base + COL_D1, da1, db1, idesc1, accum);
                        }
                    }
                }
                for (int pl = 0; pl < p.n_planes; ++pl) {
                    const uint32_t a0 = sa + pl * plane_bytes;
#pragma unroll
                    for (int k = 0; k < BKN / 16; ++k) {
                        const uint32_t accum = (kb > kb0 || pl > 0 || k > 0) ? 1u : 0u;
                        const uint64_t dones = make_smem_desc_sw128(s_ones + k * 2048, CHUNK_BYTES, 1024);
                        for (int mb = 0; mb < p.m_blocks; ++mb) {
                            const uint64_t da = make_smem_desc_sw128(a0 + mb * 2 * CHUNK_BYTES + k * 2048, CHUNK_BYTES, 1024);
                            umma_bf16(tmem_base + COL_S + 16 * mb, da, dones, idescs, accum);
                        }
                    }
                }
                umma_commit(&empty[stage]);
                if (kb == kb1 - 1) umma_commit(tmem_full);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        const int l = q * 32 + lane;
        float* ws = p.ws + (int64_t)blockIdx.x * p.ncols * 128;
        const bool any = kb1 > kb0;
        if (any) {
            mbar_wait(tmem_full, 0);
            tcgen05_fence_after();
        }
        const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        auto dump = [&](uint32_t tcol, int ncol16, int ws_col0) {
            for (int c = 0; c < ncol16; ++c) {
                float v[16];
                if (any) {
                    __syncwarp();
                    tmem_ld16(tlane + tcol + c * 16, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) ws[(int64_t)(ws_col0 + c * 16 + j) * 128 + l] = v[j];
            }
        };
        dump(0, p.un / 16, 0);
        int col = p.un;
        if (p.m_blocks == 2) { dump(COL_D1, (p.un - 128) / 16, col); col += p.un - 128; }
        for (int mb = 0; mb < p.m_blocks; ++mb) {
            float v[16];
            if (any) {
                __syncwarp();
                tmem_ld16(tlane + COL_S + 16 * mb, v);
                tmem_ld_wait();
            } else {
                v[0] = 0.f;
            }
            ws[(int64_t)(col + mb) * 128 + l] = v[0];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// G (both triangles) and s from the per-CTA partials, summed in CTA order (deterministic)
__global__ void gram_reduce_kernel(const float* __restrict__ ws, int nparts, int ncols, int un, int m_blocks, int h, float* __restrict__ G,
                                   int64_t ldg, float* __restrict__ s) {
    const int total = ncols * 128;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int c = t / 128, l = t % 128;
        float acc = 0.f;
        for (int part = 0; part < nparts; ++part) acc += ws[((int64_t)part * ncols + c) * 128 + l];
        if (c < un) {                                       // block row 0: G[l][c], mirrored into block (1,0)
            if (l < h && c < h) {
                G[(int64_t)l * ldg + c] = acc;
                if (m_blocks == 2 && c >= 128) G[(int64_t)c * ldg + l] = acc;
            }
        } else if (m_blocks == 2 && c < un + (un - 128)) {   // block (1,1)
            const int i = 128 + l, j = 128 + (c - un);
            if (i < h && j < h) G[(int64_t)i * ldg + j] = acc;
        } else {                                            // X^T 1
            const int mb = c - (ncols - m_blocks);
            const int i = mb * 128 + l;
            if (i < h) s[i] = acc;
        }
    }
}
}  // namespace gramk
}  // namespace sgf

using namespace sgf;

extern "C" int sgf_gemm_nt(const sgf_gemm_nt_args* a, void* stream) {
    if (!a || a->rows < 0 || a->n_out <= 0 || a->n_seg <= 0 || a->n_seg > SGF_MAX_SEG || a->n_a <= 0 || a->n_a > SGF_MAX_SRC ||
        a->n_b <= 0 || a->n_b > SGF_MAX_SRC || !a->out)
        return SGF_ERR_ARG;
    if (a->out_dtype != 0 && a->out_dtype != 1) return SGF_ERR_ARG;
    if (a->aux && a->aux_dtype != 0 && a->aux_dtype != 1) return SGF_ERR_ARG;
    if (a->rows == 0) return SGF_OK;
    const bool has_tail = a->b_tail != nullptr;
    if (has_tail && (a->n_b != 1 || a->n_out > 256)) return SGF_ERR_ARG;
    if (a->epi == SGF_EPI_ATTN_APPLY && (!has_tail || !a->aux)) return SGF_ERR_ARG;
    if (a->epi == SGF_EPI_ATTN_GRAM && (!has_tail || !a->bias || !a->nf_dev || a->aux)) return SGF_ERR_ARG;
    if (a->epi != SGF_EPI_AFFINE && a->epi != SGF_EPI_ATTN_APPLY && a->epi != SGF_EPI_ATTN_GRAM) return SGF_ERR_ARG;
    if ((a->r1_row == nullptr) != (a->r1_col == nullptr)) return SGF_ERR_ARG;

    nt::Params p;
    memset(&p, 0, sizeof(p));
    nt::Tmaps tm;
    memset(&tm, 0, sizeof(tm));
    p.rows = a->rows;
    p.n_out = a->n_out;
    const int n16 = (a->n_out + 15) / 16 * 16;
    p.n_blocks = (n16 + 255) / 256;
    // a 16-column tail needs n-block + 16 <= 256 TMEM columns per accumulator stage to keep the two stages (MMA of tile i+1
    // overlapping the epilogue of tile i): split a 256-wide output into two 128(+16) blocks (the tail is recomputed, cheap)
    // ... unless the whole B (all k-blocks of the 256 + 16 columns) can stay resident in shared memory: then ONE n-block with a
    // single 272-column accumulator stage reads A once and B never again (the split form re-reads the A tile and re-streams 73 KB of
    // B per 128 x 128 output tile: 5.2 GB through L2 for the 1.25 GB attention apply at the products shape, 0.76 ms, L2-bound).
    bool tail_one_block = false;
    if (has_tail && n16 + 16 > 256) {
        static const bool res_ok = [] { const char* e = getenv("SGF_ATTN_RESIDENT"); return !(e && e[0] == '0'); }();
        int kb = 0;
        for (int s = 0; s < a->n_seg; ++s) kb += (a->seg_klen[s] + nt::BK - 1) / nt::BK;
        tail_one_block = res_ok && a->schedule != SGF_GEMM_STREAM_B && n16 <= 256 && kb <= nt::RES_MAX_KB &&
                         (int64_t)(n16 + 16) * nt::BK * 2 * kb <= nt::RES_BYTES;
        if (!tail_one_block) p.n_blocks = 2;
    }
    // equal-width n-blocks: a multiple of 16 (UMMA N), and of 64 when there are several - the epilogue stores 128-byte
    // groups (64 bf16 / 32 fp32 columns), which must not reach into the next n-block's columns
    p.bn_main = p.n_blocks == 1 ? n16 : ((n16 / p.n_blocks + 63) / 64) * 64;
    if (p.bn_main > 256) return SGF_ERR_UNSUPPORTED;
    p.n_blocks = (n16 + p.bn_main - 1) / p.bn_main;
    p.has_tail = has_tail ? 1 : 0;
    const int64_t m_blocks = (a->rows + nt::BM - 1) / nt::BM;
    p.num_tiles = m_blocks * p.n_blocks;
    p.n_seg = a->n_seg;
    p.total_kb = 0;
    for (int s = 0; s < a->n_seg; ++s) {
        const int ai = a->seg_a[s], bi = a->seg_b[s];
        if (ai < 0 || ai >= a->n_a || bi < 0 || bi >= a->n_b || a->seg_klen[s] <= 0) return SGF_ERR_ARG;
        const int klen = a->seg_klen[s];
        // A partial last k-block (klen % 64 != 0) reads A/B columns past koff+klen: the caller guarantees that those A
        // columns are zero (zero padding or the end of the tensor, where TMA zero-fills) and the B columns finite.
        if (a->seg_akoff[s] + klen > a->a_cols[ai] || a->seg_bkoff[s] + klen > a->b_cols[bi]) return SGF_ERR_ARG;
        p.seg[s].a_idx = ai; p.seg[s].a_koff = a->seg_akoff[s];
        p.seg[s].b_idx = bi; p.seg[s].b_koff = a->seg_bkoff[s];
        p.seg[s].k_blocks = (klen + nt::BK - 1) / nt::BK;
        p.total_kb += p.seg[s].k_blocks;
    }
    // explicit resident-B request: narrow the n-blocks until one block of B fits (A is then re-read from L2 per n-block)
    if (a->schedule == SGF_GEMM_RESIDENT_B && !has_tail) {
        while ((int64_t)p.bn_main * nt::BK * 2 * p.total_kb > nt::RES_BYTES && p.bn_main > 64) {
            p.bn_main -= 64;
            p.n_blocks = (n16 + p.bn_main - 1) / p.bn_main;
        }
        p.num_tiles = m_blocks * p.n_blocks;
    }
    int rc;
    for (int i = 0; i < a->n_a; ++i)
        if ((rc = make_tmap_bf16(&tm.a[i], a->a[i], a->rows, a->a_cols[i], a->lda[i], nt::BM))) return rc;
    for (int i = 0; i < a->n_b; ++i)
        if ((rc = make_tmap_bf16(&tm.b[i], a->b[i], a->n_out, a->b_cols[i], a->ldb[i], p.bn_main))) return rc;
    if (has_tail && (rc = make_tmap_bf16(&tm.tail, a->b_tail, 16, a->b_cols[0], a->ldb_tail, 16))) return rc;

    p.epi = a->epi;
    p.out = a->out; p.ldo = a->ldo; p.out_dtype = a->out_dtype;
    p.bias = a->bias;
    p.aux = a->aux; p.ld_aux = a->ld_aux; p.aux_dtype = a->aux_dtype;
    p.row_scale = a->row_scale;
    p.alpha = a->alpha; p.beta = a->beta; p.alpha_dev = a->alpha_dev; p.beta_dev = a->beta_dev;
    p.relu = a->relu; p.accumulate = a->accumulate;
    p.nf = a->nf; p.nf_dev = a->nf_dev; p.den_out = a->den_out;
    p.r1_row = a->r1_row; p.r1_col = a->r1_col;
    {
        const int es = a->out_dtype == 1 ? 2 : 4;
        p.tma_store = ((reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (a->ldo * es) % 16 == 0) ? 1 : 0;
        if (p.tma_store && (rc = make_tmap_2d(&tm.out, a->out, a->out_dtype, a->rows, a->n_out, a->ldo, 32))) return rc;
        p.col_sum = a->col_sum; p.col_sumsq = a->col_sumsq;
        if ((p.col_sum || p.col_sumsq) && (!p.tma_store || a->n_out > nt::STAT_COLS)) return SGF_ERR_UNSUPPORTED;
    }

    // resident-B schedule whenever one n-block of B (all k-blocks) fits: see nt::RES_BYTES
    {
        const int64_t b_tx = (int64_t)(p.bn_main + (has_tail ? 16 : 0)) * nt::BK * 2;
        const bool fits = p.total_kb <= nt::RES_MAX_KB && b_tx * p.total_kb <= nt::RES_BYTES;
        if (a->schedule == SGF_GEMM_RESIDENT_B && !fits) return SGF_ERR_UNSUPPORTED;
        // AUTO: resident when the whole B is one n-block (loaded once per CTA): same speed as streaming at K = 256 and 1.7x
        // faster for short K (K = 100 -> 256: 0.32 vs 0.53 ms at 2.4 M rows), where a tile's few k-blocks cannot cover the TMA
        // latency of re-fetching B.  With several n-blocks the per-chunk reload of B stalls the MMA (QKV 1.41 vs 1.15 ms).
        p.b_res = (fits && (a->schedule == SGF_GEMM_RESIDENT_B || (a->schedule == SGF_GEMM_AUTO && p.n_blocks == 1))) ? 1 : 0;
        // several n-blocks: B is re-loaded per (chunk, n-block); a chunk of 8 row tiles per CTA keeps the A tiles that are
        // re-read for the following n-blocks inside L2 (148 CTAs x 8 x 128 rows x K x 2 B = 75 MB at K = 256)
        p.chunk = p.n_blocks == 1 ? (int64_t)1 << 40 : 8;
    }
    // resident-B: one CTA per SM over ROW tiles (every CTA visits all n-blocks of its rows)
    const int64_t work = p.b_res ? m_blocks : p.num_tiles;
    const int64_t grid = work < num_sms() ? work : num_sms();
    const int smem_bytes = p.b_res ? nt::SMEM_BYTES_RES : nt::SMEM_BYTES;

    // epilogue specialisation: the exact feature set of this call if it has a compiled instantiation, else the generic kernel
    int feat = (a->bias ? nt::F_BIAS : 0) | (a->aux ? nt::F_AUX : 0) | (a->relu ? nt::F_RELU : 0) |
               (a->row_scale ? nt::F_ROWSCALE : 0) | (a->accumulate ? nt::F_ACCUM : 0) | (a->r1_row ? nt::F_R1 : 0) |
               (a->epi != SGF_EPI_AFFINE ? nt::F_ATTN : 0);
    const bool al16 = (!a->bias || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0) &&
                      (!a->r1_col || (reinterpret_cast<uintptr_t>(a->r1_col) & 15) == 0) &&
                      (!a->aux || (a->aux_dtype == 1 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0 && (a->ld_aux * 2) % 16 == 0));
    const bool fast_ok = p.tma_store && a->out_dtype == 1 && a->n_out % 32 == 0 && al16 && !p.col_sum && !p.col_sumsq &&
                         p.n_blocks * p.bn_main == a->n_out;     // every 32-column piece of every n-block is complete
    static const bool no_special = [] { const char* e = getenv("SGF_GEMM_NT_GENERIC"); return e && e[0] == '1'; }();
    if (!fast_ok || no_special) feat = nt::F_GENERIC;
    const int smem_max = nt::SMEM_BYTES > nt::SMEM_BYTES_RES ? nt::SMEM_BYTES : nt::SMEM_BYTES_RES;
#define SGF_NT_CASE(FEAT)                                                                                                   \
    case (FEAT): {                                                                                                         \
        static bool attr_set = false;                                                                                      \
        if (!attr_set) {                                                                                                   \
            SGF_CUDA_TRY(cudaFuncSetAttribute(nt::gemm_nt_kernel<(FEAT)>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                              smem_max));                                                                  \
            attr_set = true;                                                                                               \
        }                                                                                                                  \
        nt::gemm_nt_kernel<(FEAT)><<<(unsigned)grid, nt::THREADS, smem_bytes, (cudaStream_t)stream>>>(tm, p);              \
        launched = true;                                                                                                   \
    } break;
    for (int attempt = 0; attempt < 2; ++attempt) {
        bool launched = false;
        switch (feat) {
            SGF_NT_CASE(0)
            SGF_NT_CASE(nt::F_BIAS)
            SGF_NT_CASE(nt::F_BIAS | nt::F_RELU)
            SGF_NT_CASE(nt::F_ROWSCALE)
            SGF_NT_CASE(nt::F_ACCUM)
            SGF_NT_CASE(nt::F_AUX)
            SGF_NT_CASE(nt::F_AUX | nt::F_ACCUM)
            SGF_NT_CASE(nt::F_AUX | nt::F_BIAS)
            SGF_NT_CASE(nt::F_AUX | nt::F_R1)
            SGF_NT_CASE(nt::F_AUX | nt::F_ATTN)
            SGF_NT_CASE(nt::F_BIAS | nt::F_ATTN)
            SGF_NT_CASE(nt::F_BIAS | nt::F_R1)
            SGF_NT_CASE(nt::F_BIAS | nt::F_R1 | nt::F_ACCUM)
            SGF_NT_CASE(nt::F_GENERIC)
            default: break;
        }
        if (launched) break;
        feat = nt::F_GENERIC;     // no instantiation for this feature set
    }
#undef SGF_NT_CASE
    SGF_LAUNCH_CHECK(); count_launch();
    return SGF_OK;
}

static inline int tn_grid(int64_t kb_total) { return (int)(kb_total < num_sms() ? kb_total : num_sms()); }

extern "C" int sgf_gemm_tn_ws_bytes(int32_t m, int32_t n, int64_t rows, size_t* bytes) {
    if (!bytes || m <= 0 || m > 256 || n <= 0 || n > 256 || rows < 0) return SGF_ERR_ARG;
    const int m_blocks = (m + 127) / 128, un = (n + 15) / 16 * 16;
    const int64_t kb_total = (rows + tn::BKN - 1) / tn::BKN;
    int grid = tn_grid(kb_total < 1 ? 1 : kb_total);
    *bytes = (size_t)grid * un * m_blocks * 128 * sizeof(float);
    return SGF_OK;
}

extern "C" int sgf_gemm_tn(const sgf_gemm_tn_args* a, void* stream) {
    if (!a || a->m <= 0 || a->m > 256 || a->n <= 0 || a->n > 256 || a->rows < 0 || !a->out || !a->ws) return SGF_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    tn::Params p;
    memset(&p, 0, sizeof(p));
    p.rows = a->rows; p.m = a->m; p.n = a->n;
    p.m_blocks = (a->m + 127) / 128;
    p.un = (a->n + 15) / 16 * 16;
    p.n_chunks = (p.un + 63) / 64;
    p.kb_total = (a->rows + tn::BKN - 1) / tn::BKN;
    p.ws = (float*)a->ws;
    if (a->n_pairs < 0 || a->n_pairs > SGF_TN_MAX_PAIRS) return SGF_ERR_ARG;
    p.n_pairs = a->n_pairs > 0 ? a->n_pairs : 1;
    int64_t a_span = a->m, b_span = a->n;      // columns the tensor maps must cover
    for (int i = 0; i < p.n_pairs; ++i) {
        p.a_off[i] = a->n_pairs > 0 ? a->a_off[i] : 0;
        p.b_off[i] = a->n_pairs > 0 ? a->b_off[i] : 0;
        if (p.a_off[i] < 0 || p.b_off[i] < 0 || p.a_off[i] % 64 || p.b_off[i] % 64) return SGF_ERR_ARG;
        if (p.a_off[i] + a->m > a_span) a_span = p.a_off[i] + a->m;
        if (p.b_off[i] + a->n > b_span) b_span = p.b_off[i] + a->n;
    }
    const int mp = p.m_blocks * 128;
    int grid = 0;
    if (a->rows > 0) {
        size_t need = 0;
        sgf_gemm_tn_ws_bytes(a->m, a->n, a->rows, &need);
        if (a->ws_bytes < need) return SGF_ERR_ARG;
        tn::Tmaps tm;
        memset(&tm, 0, sizeof(tm));
        int rc;
        if ((rc = make_tmap_bf16(&tm.a, a->a, a->rows, a_span, a->lda, tn::BKN))) return rc;
        if ((rc = make_tmap_bf16(&tm.b, a->b, a->rows, b_span, a->ldb, tn::BKN))) return rc;
        static bool attr_set = false;
        if (!attr_set) {
            SGF_CUDA_TRY(cudaFuncSetAttribute(tn::gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tn::SMEM_BYTES));
            attr_set = true;
        }
        grid = tn_grid(p.kb_total);
        tn::gemm_tn_kernel<<<grid, tn::THREADS, tn::SMEM_BYTES, st>>>(tm, p);
        SGF_LAUNCH_CHECK(); count_launch();
    }
    int64_t total = (int64_t)a->n * mp;
    int rgrid = (int)((total + 255) / 256);
    tn::tn_reduce_kernel<<<rgrid, 256, 0, st>>>(p.ws, grid, mp, p.un, a->m, a->n, a->alpha, a->alpha_dev, a->beta, a->out, a->ldo,
                                                a->transpose_out);
    SGF_LAUNCH_CHECK(); count_launch();
    return SGF_OK;
}


static int gram_geometry(int h, int planes, int64_t rows, gramk::Params& p) {
    if (h <= 0 || h > 256 || (planes != 1 && planes != 3) || rows < 0) return SGF_ERR_ARG;
    memset(&p, 0, sizeof(p));
    p.rows = rows; p.h = h;
    p.chunks = (h + 63) / 64;
    p.m_blocks = h > 128 ? 2 : 1;
    p.un = (h + 15) / 16 * 16;
    p.n_planes = planes;
    p.kb_total = (rows + gramk::BKN - 1) / gramk::BKN;
    p.stages = gramk::SMEM_BUDGET / (planes * p.chunks * gramk::CHUNK_BYTES);
    if (p.stages > gramk::MAX_STAGES) p.stages = gramk::MAX_STAGES;
    if (p.stages < 2) return SGF_ERR_UNSUPPORTED;
    p.ncols = p.un + (p.m_blocks == 2 ? p.un - 128 : 0) + p.m_blocks;
    return SGF_OK;
}
static inline int gram_grid(int64_t kb_total) { return (int)(kb_total < num_sms() ? (kb_total < 1 ? 1 : kb_total) : num_sms()); }

extern "C" int sgf_gram_ws_bytes(int32_t h, int32_t planes, int64_t rows, size_t* bytes) {
    gramk::Params p;
    int rc = gram_geometry(h, planes, rows, p);
    if (rc || !bytes) return rc ? rc : SGF_ERR_ARG;
    *bytes = (size_t)gram_grid(p.kb_total) * p.ncols * 128 * sizeof(float);
    return SGF_OK;
}

extern "C" int sgf_gram(const void* x, int64_t ldx, int64_t rows, int32_t h, int32_t planes, int64_t plane_ld, float* G, int64_t ldg,
                        float* s, void* ws, size_t ws_bytes, void* stream) {
    gramk::Params p;
    int rc = gram_geometry(h, planes, rows, p);
    if (rc) return rc;
    if (!x || !G || !s || !ws || ldg < h || (planes == 3 && (plane_ld < h || plane_ld % 64 != 0))) return SGF_ERR_ARG;
    size_t need = 0;
    sgf_gram_ws_bytes(h, planes, rows, &need);
    if (ws_bytes < need) return SGF_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    p.ws = (float*)ws;
    for (int i = 0; i < planes; ++i) p.plane_off[i] = (int)(i * plane_ld);
    if (planes == 1) {
        p.n_pairs = 1;
    } else {      // bf16x3: the six partial products, smallest terms first (kernels._PAIRS3)
        static const int PA[6] = {0, 2, 1, 0, 1, 0}, PB[6] = {2, 0, 1, 1, 0, 0};
        p.n_pairs = 6;
        for (int i = 0; i < 6; ++i) { p.pa[i] = PA[i]; p.pb[i] = PB[i]; }
    }
    const int grid = gram_grid(p.kb_total);
    if (rows > 0) {
        CUtensorMap tm;
        memset(&tm, 0, sizeof(tm));
        const int64_t span = planes == 1 ? h : 2 * plane_ld + h;
        if ((rc = make_tmap_bf16(&tm, x, rows, span, ldx, gramk::BKN))) return rc;
        const int smem_bytes = p.stages * planes * p.chunks * gramk::CHUNK_BYTES + gramk::ONES_BYTES + gramk::BAR_BYTES + 1024;
        static int attr_set = 0;
        if (attr_set < smem_bytes) {
            SGF_CUDA_TRY(cudaFuncSetAttribute(gramk::gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set = 227 * 1024;
        }
        gramk::gram_kernel<<<grid, gramk::THREADS, smem_bytes, st>>>(tm, p);
        SGF_LAUNCH_CHECK(); count_launch();
    } else {
        SGF_CUDA_TRY(cudaMemsetAsync(ws, 0, need, st));
    }
    gramk::gram_reduce_kernel<<<(p.ncols * 128 + 255) / 256, 256, 0, st>>>(p.ws, grid, p.ncols, p.un, p.m_blocks, h, G, ldg, s);
    SGF_LAUNCH_CHECK(); count_launch();
    return SGF_OK;
}
