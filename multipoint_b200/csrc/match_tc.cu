// Rows 6-7 of the hot path on the 5th-generation tensor cores: brute-force nearest-neighbour
// search (cv2.BFMatcher / NNMatcher, multipoint/utils/matching.py:7,31,50-58) as a GEMM whose
// similarity matrix never leaves the SM.
//
//   S = A . B^T  with fp32 descriptors split into bf16 planes a = a_hi + a_mid (+ 2^-18 |a|):
//   S ~= A_hi.B_hi^T + A_hi.B_mid^T + A_mid.B_hi^T      (three tcgen05.mma passes, fp32 accumulate)
//   |S~ - S| <= 4e-5 |a||b|  (match_internal.cuh); rows whose top-3 margins are inside that bound are
//   re-ranked exactly in fp64 (match.cu: pair / row / 8-row recheck kernels), so the argmax that leaves this
//   file plus the recheck equals the fp64 argmax with ties to the lowest index.
//
// One CTA (10 warps) owns 128 rows of A for one image pair:
//   - warp 0: TMA producer.  A_hi / A_mid for the whole K = D (<= 256) arrive as the first D/64 items of a
//     six-stage ring of 32 KB blocks (128 rows x 64 k, hi + mid, 128B swizzle), B_hi / B_mid follow tile by tile
//   - warps 2-9: copy A from the ring into tensor memory (tcgen05.st, one row per thread), then act as the
//     epilogue: tcgen05.ld of the finished accumulator (two warps per TMEM lane quarter, alternate 32-column
//     chunks), packed-key arg-top-3 per row while the next tile's MMAs run
//   - warp 1: one elect.sync thread issues tcgen05.mma (M=128, N=128, K=16, kind::f16, A from TMEM, B from
//     shared memory) into one of two 128-column TMEM accumulators; tcgen05.commit releases the ring stage /
//     publishes the tile
// Both directions come out of the one GEMM (matching.py:53-60: argmin over axis 1 and axis 0 of the same
// distance matrix): a thread owns one row of the accumulator, so the row side is an in-thread running arg-top-3;
// for the column side the 32 rows a warp holds of one column are reduced with two warp-collective reductions
// (redux.sync -> CREDUX): the packed maximum, and the gap to the runner-up.  Each (32-row chunk, column) leaves
// 8 bytes; match_colmerge_kernel folds the chunks of a column into the same Top2 record the row side writes.
// The only global traffic is the bf16 operands (B re-read once per 128-row block, from L2), 20 B of result
// per row and 8 B per (32-row chunk, column).
#include <cuda.h>
#include <stdlib.h>

#include "match_internal.cuh"
#include "mp_tma.cuh"

namespace mp {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;  // BK bf16 = one 128 B swizzle row
constexpr int TC_EPI_WARPS = 8;                      // two per TMEM lane quarter, alternating 32-column chunks
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr uint32_t TC_TILE_BYTES = TC_BM * TC_BK * 2;  // 16 KB: one (128 x 64) bf16 block

// ---------------------------------------------------------------- PTX wrappers
// one lane of a converged warp (cute::elect_one_sync): ptxas then knows a single thread is active and moves the
// tcgen05 / TMA operands to uniform registers without the per-instruction waterfall it emits under `lane == 0`
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// A operand from tensor memory (cute SM100_MMA_F16BF16_TS)
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// tcgen05.ld is asynchronous: the registers are valid only after this wait
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// warp-collective minimum (SASS: CREDUX.MIN into a uniform register); volatile keeps the batches in source order
__device__ __forceinline__ uint32_t warp_min_u32(uint32_t x) {
    uint32_t r;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(x));
    return r;
}
// d = a * b + c on the integer multiply-add pipe (keeps the key arithmetic off the ALU pipe that the min/max chain saturates)
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t x) {
    uint32_t r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(x));
    return r;
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (=1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024)      [46,48) version = 1 (sm_100)
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// UMMA instruction descriptor, kind::f16 (cute::UMMA::InstrDescriptor):
//   [4,6) D format 1 = f32   [7,10) A format 1 = bf16   [10,13) B format 1 = bf16
//   [15] A major 0 = K   [16] B major 0 = K   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The A operand lives in tensor memory: the CTA's 128 rows of A_hi / A_mid (D/2 32-bit columns each, two
// bf16 per column, row r in TMEM lane r) are written once with tcgen05.st and every MMA reads A from
// TMEM (tcgen05.mma ... [d], [a], b_desc).  Shared memory then only holds the B ring (six 32 KB
// stages) and serves half the operand bytes per MMA: with both operands in shared memory an M=128, N=128
// MMA needs 8 KB per 64 cycles -- all of the 128 B/cycle an SM has -- on top of the TMA fills (that variant
// was measured in round 1 and removed).  TMEM: accumulators in columns [0, ACC*128), A in the last 64*KB columns.
template <int KB>
struct TcSmem {
    static constexpr int B_STAGES = 6;
    static constexpr int ACC_STAGES = (512 - 64 * KB) / TC_BN;
    static constexpr uint32_t A_COL0 = 512 - 64 * KB;
    static constexpr uint32_t B_STAGE_BYTES = 2u * TC_TILE_BYTES;
    static constexpr uint32_t B_OFF = 0;
    static constexpr uint32_t BAR_OFF = B_OFF + B_STAGES * B_STAGE_BYTES;  // multiple of 1024
    static constexpr int NUM_BARS = 1 + 2 * B_STAGES + 2 * ACC_STAGES;
    static constexpr uint32_t BIAS_OFF = (BAR_OFF + NUM_BARS * 8 + 16 + 15) & ~15u;  // after the barriers and the tmem pointer
    // per epilogue warp: the 2 x 32 column terms of a tile, double-buffered over tiles (512 B), and the (32 x 8 B)
    // column records of a chunk, double-buffered over chunks (512 B)
    static constexpr uint32_t BIAS_BYTES = TC_EPI_WARPS * 1024;
    static constexpr uint32_t TOTAL = BIAS_OFF + BIAS_BYTES + 1024;        // + alignment slack
};

template <int KB, bool COLS, bool BIAS>
__global__ void __launch_bounds__(TC_THREADS, 1)
match_top2_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_mid,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_mid,
                     const int32_t *__restrict__ na, int NA, const int32_t *__restrict__ nb, int NB,
                     const float *__restrict__ norms_a, const float *__restrict__ norms_b,
                     const unsigned *__restrict__ max_a, const unsigned *__restrict__ max_b, Top2 *__restrict__ top,
                     uint2 *__restrict__ colpart, int RC, int NBP) {
    using L = TcSmem<KB>;
    constexpr int S = L::B_STAGES;
    constexpr int ACC = L::ACC_STAGES;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.y, m0 = blockIdx.x * TC_BM;
    const int n_a = na ? min(na[p], NA) : NA, n_b = nb ? min(nb[p], NB) : NB;

    if (m0 >= n_a || n_b <= 0) {  // nothing to search: empty results (uniform per CTA)
        for (int r = threadIdx.x; r < TC_BM; r += TC_THREADS)
            if (m0 + r < NA) {
                top[(size_t)p * NA + m0 + r] = top2_empty();
            }
        return;
    }

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle needs 1024 B alignment
    const uint32_t smem_b = base + L::B_OFF, bars = base + L::BAR_OFF;
    const uint32_t bar_a_full = bars;
    auto bar_b_full = [&](int s) { return bars + 8u * (1 + s); };
    auto bar_b_empty = [&](int s) { return bars + 8u * (1 + S + s); };
    auto bar_acc_full = [&](int t) { return bars + 8u * (1 + 2 * S + t); };
    auto bar_acc_empty = [&](int t) { return bars + 8u * (1 + 2 * S + ACC + t); };
    const uint32_t tmem_ptr_addr = bars + 8u * L::NUM_BARS;
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

    const int n_tiles = (n_b + TC_BN - 1) / TC_BN;

    if (warp == 0 && lane == 0) {
        mbar_init(bar_a_full, TC_EPI_WARPS);
        for (int s = 0; s < S; ++s) { mbar_init(bar_b_full(s), 1); mbar_init(bar_b_empty(s), 1); }
        for (int t = 0; t < ACC; ++t) { mbar_init(bar_acc_full(t), 1); mbar_init(bar_acc_empty(t), TC_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 1) {  // one warp allocates all 512 TMEM columns (1 CTA per SM: smem-limited)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int it = 0;
            // ring items 0..KB-1 are this CTA's own rows of A (hi + mid block per k-block, the same 32 KB stage
            // layout as B): the epilogue warps move them from shared to tensor memory.  Pulling A in with
            // per-thread global loads took 6 us per CTA (43 us per launch) in front of the first MMA.
            for (int kb = 0; kb < KB; ++kb, ++it) {
                mbar_expect_tx(bar_b_full(kb), L::B_STAGE_BYTES);
                const uint32_t dst = smem_b + kb * L::B_STAGE_BYTES;
                tma_load_3d(dst, &map_a_hi, bar_b_full(kb), kb * TC_BK, m0, p);
                tma_load_3d(dst + TC_TILE_BYTES, &map_a_mid, bar_b_full(kb), kb * TC_BK, m0, p);
            }
            for (int nt = 0; nt < n_tiles; ++nt) {
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % S;
                    mbar_wait(bar_b_empty(s), ((it / S) & 1) ^ 1);
                    mbar_expect_tx(bar_b_full(s), L::B_STAGE_BYTES);
                    const uint32_t dst = smem_b + s * L::B_STAGE_BYTES;
                    tma_load_3d(dst, &map_b_hi, bar_b_full(s), kb * TC_BK, nt * TC_BN, p);
                    tma_load_3d(dst + TC_TILE_BYTES, &map_b_mid, bar_b_full(s), kb * TC_BK, nt * TC_BN, p);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
            mbar_wait(bar_a_full, 0);
            tc_fence_after();
            int it = 0;
            // A has left the ring stages it arrived in: hand them back to the producer
            for (int kb = 0; kb < KB; ++kb, ++it) mbar_arrive(bar_b_empty(kb));
            for (int nt = 0; nt < n_tiles; ++nt) {
                const int t = nt % ACC;
                mbar_wait(bar_acc_empty(t), ((nt / ACC) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(t * TC_BN);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % S;
                    mbar_wait(bar_b_full(s), (it / S) & 1);
                    tc_fence_after();
                    const uint32_t b_hi = smem_b + s * L::B_STAGE_BYTES, b_mid = b_hi + TC_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint32_t ko = k * 32;  // 16 bf16 = 32 B inside the 128 B swizzle row
                        const uint64_t db_hi = umma_desc_sw128(b_hi + ko), db_mid = umma_desc_sw128(b_mid + ko);
                        // 16 bf16 of a row = 8 TMEM columns; plane offset 32*KB columns
                        const uint32_t ta_hi = tmem_base + L::A_COL0 + (uint32_t)(kb * 32 + k * 8);
                        const uint32_t ta_mid = ta_hi + 32u * KB;
                        tc_mma_f16_ts(d_tmem, ta_mid, db_hi, idesc, (kb | k) != 0);  // small terms first
                        tc_mma_f16_ts(d_tmem, ta_hi, db_mid, idesc, 1);
                        tc_mma_f16_ts(d_tmem, ta_hi, db_hi, idesc, 1);
                    }
                    tc_commit(bar_b_empty(s));  // smem stage free once these MMAs have read it
                }
                tc_commit(bar_acc_full(t));  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: running arg-top-3 per row, nearest row per column =====================
        // Every accumulator value becomes one 32-bit integer key (match_internal.cuh: KeyScale) whose low 5 bits are the
        // column's position inside its 32-column chunk (row side) or the row's inside its 32-row chunk (column side):
        // one integer max then carries value and index together; equal keys prefer the lower index.
        // Two warps share each TMEM lane quarter and take alternate chunks, so every scheduler has
        // two epilogue warps to hide the tcgen05.ld latency (with one, ncu showed 3.4 stall cycles
        // per issued instruction and the epilogue, not the MMAs, set the tile time).
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        {
            // stage this CTA's rows of A into tensor memory: warps with half 0 write the hi plane,
            // half 1 the mid plane; each thread owns one row (= one TMEM lane)
            const uint32_t col = L::A_COL0 + (uint32_t)(half * 32 * KB);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(bar_b_full(kb), 0);
                // this thread's row of the (128 x 64) bf16 block: eight 16 B chunks, chunk c at (c ^ (row & 7)) (128 B swizzle)
                const uint32_t src = smem_b + kb * L::B_STAGE_BYTES + (uint32_t)half * TC_TILE_BYTES + (uint32_t)row * 128u;
                uint32_t r[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t addr = src + (uint32_t)((q ^ (row & 7)) << 4);
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r[4 * q + 0]), "=r"(r[4 * q + 1]), "=r"(r[4 * q + 2]), "=r"(r[4 * q + 3]) : "r"(addr));
                }
                tc_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + col + (uint32_t)(kb * 32), r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a_full);
        }
        const float ma = __uint_as_float(max_a[p]), mb = __uint_as_float(max_b[p]);
        // Packed keys (match_internal.cuh: KeyScale).  Row side: f = a.b [- |b_j|^2/2] + C with C = 3 * 2^E and
        // 2^E > |a.b - ...|, so every f lies in the one binade [2^(E+1), 2^(E+2)): its 23 mantissa bits order the keys
        // and bits(f) * 32 + (31 - j) -- one integer multiply-add -- is the key with the column inside the chunk in the
        // low 5 bits.  Columns beyond n_b get the term 0: their operand rows are zero, f = 0, key < 32 = "nothing".
        const KeyScale kr = key_scale(1.002f * ma * mb + (BIAS ? 0.5f * mb * mb : 0.f));
        const float *bias = norms_b + (size_t)p * NB;
        // Column side (COLS): f = a.b [- |a_i|^2/2] + CC, the row's own term being a per-thread constant; the key is
        // complemented (smaller = closer) so that the winner and the gap to the runner-up both come out of
        // warp-collective reductions.  Rows beyond n_a carry the key ~0 and never win.
        const KeyScale kc = key_scale(1.002f * ma * mb + (BIAS ? 0.5f * ma * ma : 0.f));
        const bool row_ok = m0 + row < n_a;
        const float add_row = (COLS && BIAS && row_ok) ? fmaf(-0.5f, __ldg(norms_a + (size_t)p * NA + m0 + row), kc.C) : kc.C;
        const uint32_t opaque0 = (uint32_t)(n_tiles >> 30);           // 0, but not to the compiler: keeps the multipliers in registers
        const uint32_t mul_r = 32u + opaque0;
        const uint32_t mul_c = row_ok ? 0u - 32u : opaque0;            // ~(32 f + code) = -32 f + ~code
        const uint32_t code_c = row_ok ? ~(uint32_t)(31 - lane) : ~0u;
        uint2 *cdst = COLS ? colpart + ((size_t)p * RC + (m0 >> 5) + quarter) * NBP + lane : nullptr;
        uint8_t *wsm = smem_raw + (base + L::BIAS_OFF - smem_u32(smem_raw)) + (warp - 2) * 1024;
        float *sbias = reinterpret_cast<float *>(wsm);            // [2 tiles][64]
        uint4 *srec = reinterpret_cast<uint4 *>(wsm + 512);       // [2 chunks][16] = (m, g) of two columns each

        uint32_t best = 0, second = 0, third = 0;  // packed keys; < 32 = nothing yet
        int best_chunk = -1, second_chunk = -1;    // global chunk index (32 columns each)
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);

        // one 32-column chunk of this thread's accumulator row
        auto process = [&](const uint32_t (&v)[32], int col0, const float *sb, uint4 *rec) {
            const uint32_t old_best = best, old_second = second;
#pragma unroll
            for (int j8 = 0; j8 < 32; j8 += 8) {
                uint32_t nx[8];
#pragma unroll
                for (int j4 = j8; j4 < j8 + 8; j4 += 4) {
                    const float4 q = *reinterpret_cast<const float4 *>(sb + j4);   // broadcast LDS.128
                    const float add4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = j4 + jj;
                        const float val = __uint_as_float(v[j]);
                        const float fr = val + add4[jj];
                        const uint32_t x = mad_u32(__float_as_uint(fr), mul_r, (uint32_t)(31 - j));
                        // running sorted triple: the chain through (best, second, third) is one min/max deep per element
                        third = max(third, min(x, second));
                        second = max(second, min(x, best));
                        best = max(best, x);
                        if (COLS) {
                            const float fc = BIAS ? val + add_row : fr;   // NN: same constant, one add serves both sides
                            nx[j - j8] = mad_u32(__float_as_uint(fc), mul_c, code_c);
                        }
                    }
                }
                if (COLS) {
                    // eight columns per batch: the eight winners first, then the eight gaps, so that the collectives'
                    // latencies overlap instead of forming one dependent chain per column.  m - nx is 0 for the winner
                    // and 2^32 - gap for everybody else: its maximum is the runner-up.
                    uint32_t m[8], g[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) m[k] = warp_min_u32(nx[k]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[k] = warp_max_u32(m[k] - nx[k]);
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < 8; k += 2) rec[(j8 + k) >> 1] = make_uint4(m[k], g[k], m[k + 1], g[k + 1]);
                    }
                }
            }
            if (COLS) {
                __syncwarp();
                cdst[col0] = reinterpret_cast<const uint2 *>(rec)[lane];   // 256 B per warp, coalesced; columns >= n_b are padding
            }
            // which chunk the new best / second came from (keys of different chunks can be equal only in their low bits'
            // meaning, never in value order: a changed key is a key of this chunk)
            const int chunk = col0 >> 5;
            if (second != old_second) second_chunk = (second == old_best && best != old_best) ? best_chunk : chunk;
            if (best != old_best) best_chunk = chunk;
        };
        // the raw |b_j|^2 of the two columns this lane stages for tile nt; the -|b_j|^2/2 + C arithmetic happens when the
        // value is stored a tile later (next to the load it would wait for the whole global-memory latency)
        auto fetch_terms = [&](int nt, float &a0, float &a1) {
            const int ca = nt * TC_BN + half * 32 + lane, cb_ = ca + 64;
            a0 = 0.f; a1 = 0.f;
            if (BIAS) {
                if (ca < n_b) a0 = __ldg(bias + ca);
                if (cb_ < n_b) a1 = __ldg(bias + cb_);
            }
        };
        auto store_terms = [&](int nt, float a0, float a1, float *dst) {
            const int ca = nt * TC_BN + half * 32 + lane, cb_ = ca + 64;
            dst[lane] = ca < n_b ? (BIAS ? fmaf(-0.5f, a0, kr.C) : kr.C) : 0.f;
            dst[32 + lane] = cb_ < n_b ? (BIAS ? fmaf(-0.5f, a1, kr.C) : kr.C) : 0.f;
        };
        // Software pipeline over the 2 * n_tiles chunks of this warp: the tcgen05.ld of chunk s+1 is in flight while chunk s
        // is processed (tcgen05.wait::ld waits for every outstanding load, so it is issued right after the wait for chunk s).
        // The 2 x 32 column terms of a tile, one per lane, are fetched a whole tile ahead and become warp-visible through
        // shared memory (double-buffered over tiles).
        uint32_t va[32], vb[32];
        float a0n, a1n;
        fetch_terms(0, a0n, a1n);
        store_terms(0, a0n, a1n, sbias);
        if (n_tiles > 1) fetch_terms(1, a0n, a1n);
        mbar_wait(bar_acc_full(0), 0);
        tc_fence_after();
        tc_ld32(tlane + (uint32_t)(half * 32), va);
        for (int nt = 0; nt < n_tiles; ++nt) {
            const int t = nt % ACC;
            const int n0 = nt * TC_BN;
            const uint32_t tacc = tlane + (uint32_t)(t * TC_BN);
            // ---- chunk 0 of the tile is in va (loading); start chunk 1 behind it
            tc_ld_wait();
            tc_ld32(tacc + (uint32_t)((half + 2) * 32), vb);
            __syncwarp();   // the tile's column terms are visible; the previous chunk's records have been read
            process(va, n0 + half * 32, sbias + (nt & 1) * 64, srec);
            // ---- chunk 1: wait, hand the accumulator stage back, start the next tile's chunk 0
            tc_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(t));
            if (nt + 1 < n_tiles) {
                store_terms(nt + 1, a0n, a1n, sbias + ((nt + 1) & 1) * 64);      // last read two tiles ago
                if (nt + 2 < n_tiles) fetch_terms(nt + 2, a0n, a1n);
                const int t1 = (nt + 1) % ACC;
                mbar_wait(bar_acc_full(t1), ((nt + 1) / ACC) & 1);
                tc_fence_after();
                tc_ld32(tlane + (uint32_t)(t1 * TC_BN + half * 32), va);
            }
            process(vb, n0 + (half + 2) * 32, sbias + (nt & 1) * 64 + 32, srec + 16);
        }
        // merge the two column subsets of each row (operand smem is free: every MMA has completed)
        uint32_t *mrg = reinterpret_cast<uint32_t *>(smem_raw + (base - smem_u32(smem_raw)));
        if (half == 1) {
            mrg[row * 4 + 0] = best; mrg[row * 4 + 1] = second;
            mrg[row * 4 + 2] = (uint32_t)best_chunk; mrg[row * 4 + 3] = (uint32_t)second_chunk;
            mrg[512 + row] = third;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
        if (half == 0) {
            const uint32_t ob = mrg[row * 4 + 0], os = mrg[row * 4 + 1];
            const int obc = (int)mrg[row * 4 + 2], osc = (int)mrg[row * 4 + 3];
            // (best, second) of the union; on equal packed keys either index is fine (flagged anyway)
            uint32_t nb_, ns_; int nbc, nsc;
            if (ob > best) { nb_ = ob; nbc = obc; if (best >= os) { ns_ = best; nsc = best_chunk; } else { ns_ = os; nsc = osc; } }
            else { nb_ = best; nbc = best_chunk; if (ob >= second) { ns_ = ob; nsc = obc; } else { ns_ = second; nsc = second_chunk; } }
            const uint32_t ot = mrg[512 + row];
            third = max(max(third, ot), max(min(second, ob), min(best, os)));  // before best/second are overwritten
            best = nb_; best_chunk = nbc; second = ns_; second_chunk = nsc;
        }
        if (half == 0 && m0 + row < NA) {
            Top2 out = top2_empty();
            if (m0 + row < n_a) {
                if (best >= 32u) {
                    out.best_idx = best_chunk * 32 + 31 - (int)(best & 31u);
                    out.best = key_value(kr, best);
                }
                if (second >= 32u) {
                    out.second_idx = second_chunk * 32 + 31 - (int)(second & 31u);
                    out.second = key_value(kr, second);
                }
                if (third >= 32u) out.third = key_value(kr, third);
            }
            top[(size_t)p * NA + m0 + row] = out;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------- host side
// (P, N, D) bf16 row-major -> 3-D map, box = 64 k x 128 rows x 1 pair, 128-byte swizzle
static int make_operand_map(CUtensorMap *map, const __nv_bfloat16 *ptr, int P, int N, int D) {
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) {
        set_error("match_top2_tensor: cuTensorMapEncodeTiled not available from the driver");
        return MP_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)P};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)N * D * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("match_top2_tensor: cuTensorMapEncodeTiled failed with CUresult %d (N=%d D=%d P=%d)", (int)r, N, D, P);
        return MP_ERR_CUDA;
    }
    return MP_OK;
}

// ---------------------------------------------------------------- column side: fold the 32-row chunks
// One thread per column j of pair p: the chunk records (~maximum, gap g) give the chunk's two best packed keys
// (low 5 bits = 31 - row inside the chunk).  The chunk's third key is unknown but not larger than
// its second, so the second is entered twice: a near-tie whose best and second share a chunk is then classified
// for the full exact rescan (conservative), every other case is exact.
__global__ void __launch_bounds__(256)
match_colmerge_kernel(const uint2 *__restrict__ colpart, int RC, int NBP, const int32_t *__restrict__ na, int NA,
                      const int32_t *__restrict__ nb, int NB, int use_bias, const unsigned *__restrict__ max_a,
                      const unsigned *__restrict__ max_b, Top2 *__restrict__ top_cols) {
    const int p = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= NB) return;
    const int n_a = na ? min(na[p], NA) : NA, n_b = nb ? min(nb[p], NB) : NB;
    Top2 out = top2_empty();
    if (j < n_b && n_a > 0) {
        const float ma = __uint_as_float(max_a[p]), mb = __uint_as_float(max_b[p]);
        const KeyScale kc = key_scale(1.002f * ma * mb + (use_bias ? 0.5f * ma * ma : 0.f));
        const uint2 *src = colpart + (size_t)p * RC * NBP + j;
        const int chunks = (n_a + 31) >> 5;
        uint32_t b = 0, s = 0, t = 0;
        int bc = -1, sc = -1;
        auto insert = [&](uint32_t x, int c, bool indexed) {
            if (x > b) { t = s; s = b; sc = bc; b = x; bc = c; }
            else if (x > s) { t = s; s = x; sc = indexed ? c : -2; }
            else if (x > t) t = x;
        };
        // eight chunk records in flight per thread: the loop is a chain of dependent L2 round trips otherwise
        for (int c0 = 0; c0 < chunks; c0 += 8) {
            uint2 cur[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                cur[u] = c0 + u < chunks ? __ldg(src + (size_t)(c0 + u) * NBP) : make_uint2(~0u, 0u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t k1 = ~cur[u].x;   // the kernel reduces complemented keys
                if (k1 < 32u) continue;
                // cur.y = 2^32 - (runner-up - winner) of the complemented keys, 0 = no runner-up
                const uint32_t k2 = cur[u].y == 0u ? 0u : ~(cur[u].x - cur[u].y);
                insert(k1, c0 + u, true);
                if (k2 >= 32u) { insert(k2, c0 + u, true); insert(k2, c0 + u, false); }
            }
        }
        // equal packed keys cannot occur inside a chunk (distinct row codes); across chunks the strict '>' keeps the
        // earlier chunk = the lower row index, and such ties are inside the recheck margin anyway
        if (bc >= 0) { out.best_idx = bc * 32 + 31 - (int)(b & 31u); out.best = key_value(kc, b); }
        if (sc >= 0) { out.second_idx = sc * 32 + 31 - (int)(s & 31u); out.second = key_value(kc, s); }
        if (t >= 32u) out.third = key_value(kc, t);
    }
    top_cols[(size_t)p * NB + j] = out;
}

template <int KB, bool COLS, bool BIAS>
static int launch_tc(const CUtensorMap &ah, const CUtensorMap &am, const CUtensorMap &bh, const CUtensorMap &bm,
                     const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_mid, const int32_t *na, int NA,
                     const int32_t *nb, int NB, int P, const float *norms_a, const float *norms_b,
                     const unsigned *max_a, const unsigned *max_b, Top2 *top, uint2 *colpart, int RC, int NBP, cudaStream_t s) {
    auto k = match_top2_tc_kernel<KB, COLS, BIAS>;
    MP_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcSmem<KB>::TOTAL));
    dim3 grid((NA + TC_BM - 1) / TC_BM, P);
    k<<<grid, TC_THREADS, TcSmem<KB>::TOTAL, s>>>(ah, am, bh, bm, na, NA, nb, NB, norms_a, norms_b,
                                                        max_a, max_b, top, colpart, RC, NBP);
    MP_LAUNCH_OK_S("match_top2_tc_kernel", s);
    return MP_OK;
}

size_t match_colpart_bytes(int P, int NA, int NB) {
    const size_t rc = 4 * (size_t)((NA + TC_BM - 1) / TC_BM), nbp = (size_t)((NB + TC_BN - 1) / TC_BN) * TC_BN;
    return sizeof(uint2) * (size_t)P * rc * nbp;
}

int match_top2_tensor(const __nv_bfloat16 *a_hi, const __nv_bfloat16 *a_mid, const int32_t *na, int NA,
                      const __nv_bfloat16 *b_hi, const __nv_bfloat16 *b_mid, const int32_t *nb, int NB, int P, int D,
                      const float *norms_a, const float *norms_b, int use_bias, const unsigned *max_a, const unsigned *max_b,
                      Top2 *top, Top2 *top_cols, void *colpart, cudaStream_t stream) {
    if (D % 64 != 0 || D > 256 || D <= 0) {
        set_error("match_top2_tensor: D=%d must be a multiple of 64 and <= 256", D);
        return MP_ERR_UNSUPPORTED;
    }
    CUtensorMap ah, am, bh, bm;
    int rc;
    if ((rc = make_operand_map(&ah, a_hi, P, NA, D)) != MP_OK) return rc;
    if ((rc = make_operand_map(&am, a_mid, P, NA, D)) != MP_OK) return rc;
    if ((rc = make_operand_map(&bh, b_hi, P, NB, D)) != MP_OK) return rc;
    if ((rc = make_operand_map(&bm, b_mid, P, NB, D)) != MP_OK) return rc;
    const int RC = 4 * ((NA + TC_BM - 1) / TC_BM), NBP = (NB + TC_BN - 1) / TC_BN * TC_BN;
    uint2 *cp = (uint2 *)colpart;
    const bool cols = top_cols != nullptr;
    if (cols && cp == nullptr) {
        set_error("match_top2_tensor: the column side needs its chunk buffer");
        return MP_ERR_WORKSPACE;
    }
#define MP_TC_ARGS ah, am, bh, bm, a_hi, a_mid, na, NA, nb, NB, P, norms_a, norms_b, max_a, max_b, top, cp, RC, NBP, stream
#define MP_TC_LAUNCH(KB)                                                                             \
    rc = cols ? (use_bias ? launch_tc<KB, true, true>(MP_TC_ARGS) : launch_tc<KB, true, false>(MP_TC_ARGS))     \
              : (use_bias ? launch_tc<KB, false, true>(MP_TC_ARGS) : launch_tc<KB, false, false>(MP_TC_ARGS));  \
    break
    switch (D / 64) {
        case 1: MP_TC_LAUNCH(1);
        case 2: MP_TC_LAUNCH(2);
        case 3: MP_TC_LAUNCH(3);
        default: MP_TC_LAUNCH(4);
    }
#undef MP_TC_ARGS
#undef MP_TC_LAUNCH
    if (rc != MP_OK || !cols) return rc;
    dim3 grid((NB + 255) / 256, P);
    match_colmerge_kernel<<<grid, 256, 0, stream>>>(cp, RC, NBP, na, NA, nb, NB, use_bias, max_a, max_b, top_cols);
    MP_LAUNCH_OK_S("match_colmerge_kernel", stream);
    return MP_OK;
}

}  // namespace mp
