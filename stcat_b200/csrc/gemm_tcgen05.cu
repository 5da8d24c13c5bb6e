// bf16 tensor-core GEMM for sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma (fp32 accumulators in
// TMEM) -> tcgen05.ld epilogue (bias / ReLU / dtype) -> swizzled shared staging -> TMA store or TMA reduce-add.
//
//   C[M,N] (+)= A(M,K) . B(N,K)^T      A, B bf16; C fp32 or bf16; fp32 accumulation
//
// One kernel serves the three GEMMs of a Linear layer (stcat_linear_{fwd,bwd_data,bwd_weight}) by letting
// either operand be "K-major" (contraction index contiguous in memory) or "MN-major" (row index contiguous):
//   fwd        y  = x . w^T        A = x  [M,K]  K-major     B = w  [N,K]  K-major
//   bwd_data   dx = dy . w         A = dy [M,N'] K-major     B(k,n') = w[n',k]    MN-major
//   bwd_weight dw = dy^T . x       A(n,m) = dy[m,n] MN-major B(k,m) = x[m,k]      MN-major
// so no transposed copies of activations are ever materialised.
//
// Structure (persistent, warp-specialised, 1 CTA / SM; 640 threads for the 128 x 256 tile, 256 for the 128 x 64 tile):
//   warp 0      TMA producer  : 4-stage ring of {A 128x64, B 256x64} bf16 tiles (48 KB / stage), mbarrier full/empty
//                               (weight-resident variant: the [256 x K] weight tile stays in shared memory, the ring holds A only)
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 128x256x16, 4 per stage;
//                               tcgen05.commit releases the stage / publishes the accumulator
//   warp 2      TMEM allocator: 512 columns = 2 accumulator stages of 128 lanes x 256 fp32 columns
//   warps 4..19 epilogue      : 128 x 256 tile ("warp epilogue"): 16 independent warps, (TMEM lane quadrant) x (64-column
//                               group); each tcgen05.ld's 32 lanes x 32 columns at a time, adds the bias slice staged in shared
//                               memory, ReLU / mask / dropout, converts, st.shared into ITS OWN 64B-swizzled [32 x 64 B] staging
//                               tile and issues its own TMA store / reduce-add -- no barrier between warps inside a tile.
//   warps 4..7  epilogue      : 128 x 64 tile: one group of 4 warps with a shared 128B-swizzled [128 x 128 B] staging tile
//                               (also the former epilogue of the wide tile, kept behind STCAT_GEMM_WEPI=0);
//                               cluster split-K variant: partial tiles parked in shared memory, reduced through DSMEM
// M/N/K tails need no code: TMA zero-fills out-of-bounds loads and clips out-of-bounds stores.
// Split-K (needed by bwd_weight, whose contraction runs over all M = T*S tokens while the output is one
// or a few tiles) uses the TMA reduce-add epilogue on a pre-zeroed fp32 output.
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace stcat {

namespace tc {

constexpr int BM = 128, BK = 64;  // BK * 2 B = 128 B = one swizzle row
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int ACC_STAGES = 2;

// Two tile shapes.  BN = 256 is the throughput shape (128 x 256 accumulator, two epilogue groups).  BN = 64 is the
// latency shape for GEMMs with so few 128 x 256 tiles that most SMs would idle (the decoder's [t, 256] query-side
// layers): 4x as many CTAs, each with a short K loop, and -- unlike split-K -- a deterministic summation order.
// EPIMODE (BN = 256 only):
//   0  two epilogue groups of 4 warps, one [128 x 128 B] staging tile per group;
//   1  two staging tiles per group, so the conversion of chunk c+1 runs while the TMA store of chunk c still reads its tile
//      (with one tile the group waits for every store to drain), paid for with one operand stage (3 instead of 4);
//   3  "warp epilogue": 16 epilogue warps (4 per SM sub-partition), each an independent pipeline over its own 32 rows x 64
//      columns of the accumulator with its own [32 x 64 B] staging tile (64B swizzle) and its own TMA stores: no barrier
//      between warps inside a tile, latencies (TMEM load, store drain) hidden by the other three warps of the sub-partition,
//      32 KB of staging instead of 64.
//
// BRES ("B resident", BN = 256, warp epilogue, single-term jobs with K <= 256): at K = 256 a 128 x 256 tile needs 16 MMAs
// (2.1 k clocks) but 192 KB of operands, 128 KB of which is the weight tile that every tile of the same column block re-reads
// (FFN linear1: 164 MB of operand loads + 56 MB of stores per launch; TMA load latency under that load is ~5 k clocks, so a
// 3-stage ring cannot keep the MMAs fed).  With BRES the CTA keeps the whole [256 x K] weight tile in shared memory, takes a
// CONTIGUOUS range of the work list (column-block-major, so the tile changes at most a few times per CTA) and streams only
// the A tiles (64 KB per output tile) through a 4-stage ring: operand traffic drops 3x.
template <int BN, int EPIMODE = 0, bool BRES = false> struct Cfg {
    static constexpr bool WEPI = EPIMODE == 3;
    static_assert(!BRES || (BN == 256 && WEPI), "BRES: 128 KB weight tile + 64 KB A ring + 32 KB warp-epilogue staging");
    static_assert(!WEPI || BN == 256, "warp epilogue: 4 column groups of 64");
    static constexpr int STAGES = BRES ? 4 : (BN >= 256 ? (EPIMODE == 1 ? 3 : 4) : 8);   // ~192 KB of operands in flight
    static constexpr int EPI_BUFS = EPIMODE == 1 ? 2 : 1;
    static constexpr int EPI_BYTES = BM * 128;                      // group epilogue: one staging tile, 128 rows x 128 B
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int BRES_KB = 4;                               // k-blocks of the resident weight tile (K <= 256)
    static constexpr int BRES_BYTES = BRES ? BRES_KB * B_BYTES : 0;
    static constexpr int STAGE_BYTES = BRES ? A_BYTES : A_BYTES + B_BYTES;
    static constexpr int GROUPS = WEPI ? 4 : (BN >= 256 ? 2 : 1);   // epilogue column groups (4 warps each)
    static constexpr int GC = BN / GROUPS;                          // accumulator columns per group
    static constexpr int EPI_WARPS = 4 * GROUPS;
    static constexpr int THREADS = 128 + 32 * EPI_WARPS;            // 4 control warps + epilogue warps
    static constexpr int WEPI_TILE = 32 * 64;                       // warp epilogue: [32 rows x 64 B] per warp
    static constexpr int EPI_SMEM = WEPI ? EPI_WARPS * WEPI_TILE : GROUPS * EPI_BUFS * EPI_BYTES;
    static constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512 / 128 (power of two >= 32)
    static constexpr int BIAS_BYTES = (WEPI ? 2 : 1) * BN * 4;      // the tile's bias slice (warp epilogue: double-buffered)
    // the dynamic shared memory window starts 1024-aligned when the kernel has no static shared memory (checked at run
    // time); the warp-epilogue configurations have no room for alignment slack
    static constexpr int SMEM_BYTES = BRES_BYTES + STAGES * STAGE_BYTES + EPI_SMEM + BIAS_BYTES + (WEPI ? 0 : 1024) /*align slack*/ + 256 /*barriers*/;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
};

constexpr int MAX_TERMS = 3;   // y = sum_t A_t . B_t^T: the reference adds up to three Linear outputs (query_decoder.py:329-339)
constexpr int MAX_JOBS = 12;   // independent GEMMs served by one launch (grouped launch)

// One GEMM of a launch: C[M,N] (+)= sum_t A_t(M,K_t) . B_t(N,K_t)^T (+ sum_t bias_t) with its own tensor maps.
struct alignas(64) Job {
    CUtensorMap tmA[MAX_TERMS];
    CUtensorMap tmB[MAX_TERMS];
    CUtensorMap tmC;
    const float* bias[MAX_TERMS];  // [N] or null; all are added (by k-split 0)
    int kb[MAX_TERMS];             // k-blocks of each term
    int nterms;
    int M, N;
    int tiles_m, tiles_n, splits, kb_per_split;  // splits > 1 only with nterms == 1
    int work0;                     // first work item (tile x split) of this job in the launch
    int relu;
    int reduce_add;                // epilogue uses TMA reduce-add instead of store
    const __nv_bfloat16* mask;     // optional [M, N] bf16: C is zeroed where mask <= 0 (ReLU backward), else null
    int64_t ld_mask;
    float* colsum;                 // optional [N]: accumulated (atomicAdd) with the column sums of the stored C
    void* c_ptr;                   // CLK (cluster split-K): the output is written with plain stores by the row owners
    int64_t ldc;
    int accumulate;
    float alpha;                   // scale applied before the ReLU-backward mask (1 = none)
    DropArgs drop;                 // dropout on the output after bias / ReLU (thresh == 0: none); index row * N + col
};
template <int NJ> struct GroupParams {
    Job jobs[NJ];
    int njobs, total;
    long long* trace;  // diagnostics (TRACE instantiation only): SM clock at the phase boundaries of CTA 0's first 8 tiles
};

// CLK ("cluster split-K", BN = 64, one single-term job): few output tiles with a long contraction (the decoders' FFN at
// [t <= 128, 2048]: 4 tiles x 32 k-blocks, 12 us on 4 SMs).  A thread-block cluster of S = 2 / 4 / 8 CTAs takes one tile; CTA r
// runs k-blocks [r kb/S, (r+1) kb/S) into its TMEM accumulator and parks the fp32 partial tile in its own shared memory; after
// a cluster barrier CTA r sums rows [128 r / S, 128 (r+1) / S) over the S partials through distributed shared memory in rank
// order (a fixed summation order: bit-reproducible, unlike a reduce-add split-K), applies bias / ReLU / accumulate and writes
// them with plain stores.
// EPIX ("epilogue extras", single-job bf16-output instantiations only): the output is scaled by Job::alpha and / or passed
// through the train-mode dropout mask of Job::drop -- the FFN's inner dropout in the linear1 epilogue, its backward (with the
// ReLU mask) in the linear2 data-gradient epilogue.  A template parameter, so that the default instantiations carry none of it.
template <bool A_MN, bool B_MN, bool OUT_BF16, int BN, int NJ, bool TRACE = false, int EPIMODE = 0, bool BRES = false, bool CLK = false,
          bool EPIX = false>
__global__ void __launch_bounds__((Cfg<BN, EPIMODE, BRES>::THREADS), 1)
gemm_tc_kernel(const __grid_constant__ GroupParams<NJ> gp) {
    static_assert(!CLK || (BN == 64 && NJ == 1 && EPIMODE == 0 && !BRES && !A_MN), "CLK: skinny tile, one job");
    static_assert(!EPIX || (NJ == 1 && OUT_BF16 && !A_MN && !CLK && !TRACE), "EPIX: one job, bf16 output");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t bres_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B-swizzle atoms need 1024 B alignment
    using C = Cfg<BN, EPIMODE, BRES>;
    constexpr bool EPI2 = EPIMODE == 1;  // two staging tiles per group
    constexpr bool WEPI = C::WEPI;
    if (WEPI && (bres_base != smem_u32(smem_raw))) asm volatile("trap;");  // no alignment slack in these configurations
    constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, GROUPS = C::GROUPS, GC = C::GC, TMEM_COLS = C::TMEM_COLS;
    constexpr int EPI_BYTES = C::EPI_BYTES, B_BYTES = C::B_BYTES;
    const uint32_t base = bres_base + C::BRES_BYTES;  // operand ring
    const uint32_t epi_base = base + STAGES * STAGE_BYTES;
    const uint32_t bias_base = epi_base + C::EPI_SMEM;
    const uint32_t bar_base = bias_base + C::BIAS_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto accf_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto acce_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
    const uint32_t bfull_bar = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);  // BRES: the resident weight tile has landed
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = gp.total;
    // work items of this CTA: w_lo, w_lo + w_step, ... < w_hi.  Strided over the grid by default; BRES: a contiguous range of the
    // (job, column block, row block)-ordered list, so that successive items share the weight tile.
    uint32_t crank = 0, csize = 1;  // CLK: rank in / size of the cluster that shares this CTA's tile
    if (CLK) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    }
    const int w_lo = CLK ? (int)(blockIdx.x / csize) : BRES ? (int)(((long long)total * blockIdx.x) / gridDim.x) : (int)blockIdx.x;
    const int w_hi = CLK ? w_lo + 1 : BRES ? (int)(((long long)total * (blockIdx.x + 1)) / gridDim.x) : total;
    const int w_step = (BRES || CLK) ? 1 : (int)gridDim.x;
    // work item -> (job, split, m_blk, n_blk); jobs are few, a linear scan of the prefix table is enough
    auto find_job = [&](int w) {
        int j = 0;
        if (NJ > 1) while (j + 1 < gp.njobs && w >= gp.jobs[j + 1].work0) ++j;
        return j;
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&gp.jobs[0].tmA[0])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&gp.jobs[0].tmB[0])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&gp.jobs[0].tmC)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(accf_bar(s), 1); mbar_init(acce_bar(s), C::EPI_WARPS); }
        mbar_init(bfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // everything above (barriers, TMEM, descriptor prefetch) ran under the previous kernel's tail; its data is needed now
    pdl_launch_dependents();
    pdl_wait();
    // diagnostics: event ev of this CTA's ti-th tile -> trace[ti * 8 + ev] (CTA 0 only, one lane per role)
    auto T = [&](int ti, int ev) {
        if (TRACE && gp.trace != nullptr && blockIdx.x == 0 && ti < 8) gp.trace[ti * 8 + ev] = clock64();
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int b_key = -1;                  // BRES: (job, column block) of the weight tile in shared memory
            int tiles_done = 0;              // BRES: tiles issued so far (selects the accumulator stage / phase of the last one)
            for (int w = w_lo; w < w_hi; w += w_step) {
                const int ti = TRACE ? (w - w_lo) / w_step : 0;
                bool first = true;
                const int ji = find_job(w);
                const Job& J = gp.jobs[ji];
                const int lw = w - J.work0;
                const int split = CLK ? (int)crank : lw % J.splits;
                const int t = lw / J.splits;
                const int m_blk = t % J.tiles_m, n_blk = t / J.tiles_m;
                if (BRES) {
                    const int key = ji * 65536 + n_blk;
                    if (key != b_key) {
                        if (tiles_done > 0) {
                            // the MMAs of the previous tile (the last readers of the old weight tile) must have completed:
                            // their commit is what the epilogue waits for, too
                            const int pt = tiles_done - 1;
                            mbar_wait(accf_bar(pt % ACC_STAGES), (uint32_t)(pt / ACC_STAGES) & 1u);
                        }
                        const int kbs = J.kb[0];
                        mbar_expect_tx(bfull_bar, (uint32_t)(kbs * B_BYTES));
                        for (int kb = 0; kb < kbs; ++kb) {
                            const uint32_t sb = bres_base + kb * B_BYTES;
                            if (!B_MN) {
                                tma_load_2d(sb, &J.tmB[0], bfull_bar, kb * BK, n_blk * BN);
                            } else {
#pragma unroll
                                for (int j = 0; j < BN / 64; ++j)
                                    tma_load_2d(sb + j * (BK * 128), &J.tmB[0], bfull_bar, n_blk * BN + j * 64, kb * BK);
                            }
                        }
                        b_key = key;
                    }
                    ++tiles_done;
                }
                for (int term = 0; term < J.nterms; ++term) {
                    const int kb0 = split * J.kb_per_split;  // splits == 1 for multi-term jobs: kb0 = 0
                    const int kb1 = min(J.kb[term], kb0 + J.kb_per_split);
                    const CUtensorMap* tmA = &J.tmA[term];
                    const CUtensorMap* tmB = &J.tmB[term];
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        if (TRACE && first) { T(ti, 0); first = false; }  // first stage of the tile free: loads start
                        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                        mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                        if (!A_MN) {
                            tma_load_2d(sa, tmA, full_bar(stage), kb * BK, m_blk * BM);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                tma_load_2d(sa + j * (BK * 128), tmA, full_bar(stage), m_blk * BM + j * 64, kb * BK);
                        }
                        if (!BRES) {
                            if (!B_MN) {
                                tma_load_2d(sb, tmB, full_bar(stage), kb * BK, n_blk * BN);
                            } else {
#pragma unroll
                                for (int j = 0; j < BN / 64; ++j)
                                    tma_load_2d(sb + j * (BK * 128), tmB, full_bar(stage), n_blk * BN + j * 64, kb * BK);
                            }
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                T(ti, 1);  // all loads of the tile issued
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, majors, N >> 3, M >> 4
            const uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            int b_key = -1;
            uint32_t bphase = 0;
            for (int w = w_lo; w < w_hi; w += w_step) {
                const int ti = TRACE ? (w - w_lo) / w_step : 0;
                bool first = true;
                const int ji = find_job(w);
                const Job& J = gp.jobs[ji];
                const int split = CLK ? (int)crank : (w - J.work0) % J.splits;
                if (BRES) {
                    const int key = ji * 65536 + ((w - J.work0) / J.splits) / J.tiles_m;
                    if (key != b_key) {  // a new weight tile: wait for it once
                        mbar_wait(bfull_bar, bphase);
                        bphase ^= 1;
                        b_key = key;
                    }
                }
                mbar_wait(acce_bar(as), aphase ^ 1);  // epilogue has drained this accumulator stage
                T(ti, 2);  // accumulator stage free
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                uint32_t acc = 0;  // the first MMA of the tile overwrites the accumulator
                for (int term = 0; term < J.nterms; ++term) {
                    const int kb0 = split * J.kb_per_split;
                    const int kb1 = min(J.kb[term], kb0 + J.kb_per_split);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        if (TRACE && first) { T(ti, 3); first = false; }  // first k-block landed
                        tc_fence_after();
                        const uint32_t sa = base + stage * STAGE_BYTES, sb = BRES ? bres_base + kb * B_BYTES : sa + A_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; 8-row groups 1024 B apart (SBO).
                            // MN-major: 16 k-rows = 2 swizzle atoms of 8 rows x 128 B (SBO = 1024 B); successive
                            //           64-element MN blocks are BK*128 B apart (LBO).
                            const uint64_t ad = A_MN ? make_desc(sa + k * 2048, BK * 128, 1024) : make_desc(sa + k * 32, 16, 1024);
                            const uint64_t bd = B_MN ? make_desc(sb + k * 2048, BK * 128, 1024) : make_desc(sb + k * 32, 16, 1024);
                            umma_bf16(d_tmem, ad, bd, idesc, acc);
                            acc = 1u;
                        }
                        umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit(accf_bar(as));  // accumulator complete -> epilogue
                T(ti, 4);  // all MMAs of the tile issued
                if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
            }
        }
    } else if (WEPI && warp >= 4) {
        // ================= warp epilogue: 16 independent warps, (TMEM lane quadrant q) x (64-column group cg) =================
        const int ew = warp - 4;
        const int q = ew & 3;                       // the hardware ties a warp to TMEM lanes [32 (warp % 4), +32)
        const int cg = ew >> 2;
        const int gt = (int)threadIdx.x - 128 - cg * 128;  // thread index inside the column group (4 warps)
        const uint32_t sbuf = epi_base + ew * C::WEPI_TILE;
        constexpr int CHUNK_COLS = OUT_BF16 ? 32 : 16;    // 64 B of output per row per chunk
        constexpr int NCHUNK = GC / CHUNK_COLS;
        // 16-byte unit j of row r in the 64B-swizzled staging tile (CU_TENSOR_MAP_SWIZZLE_64B: address bits [4,6) ^= bits [7,9))
        auto swz = [](int j, int r) { return j ^ ((r >> 1) & 3); };
        int as = 0;
        uint32_t aphase = 0;
        // Bias slice of the column group (GC floats), staged in shared memory by the group's 4 warps.  It only depends on (job,
        // column block): restaged when that changes (BRES: a few times per CTA), from a register loaded one tile ahead so that
        // the global-load latency is off the tile's critical path.  Double-buffered by restage count: one barrier of the group
        // per restage (a warp can run at most one restage ahead of the slowest warp of its group, which then still reads the
        // other buffer).
        auto bias_key = [&](int w) {
            const int ji = find_job(w);
            const Job& J = gp.jobs[ji];
            const int lw = w - J.work0;
            return (ji << 20) | (((lw / J.splits) / J.tiles_m) << 1) | ((lw % J.splits) == 0 ? 1 : 0);
        };
        auto bias_load = [&](int w) {
            const Job& J = gp.jobs[find_job(w)];
            const int lw = w - J.work0;
            const int col = ((lw / J.splits) / J.tiles_m) * BN + cg * GC + gt;
            float bv = 0.f;
            if ((lw % J.splits) == 0 && gt < GC && col < J.N)
                for (int term = 0; term < J.nterms; ++term)
                    if (J.bias[term] != nullptr) bv += __ldg(J.bias[term] + col);
            return bv;
        };
        int staged_key = -1, next_key = w_lo < w_hi ? bias_key(w_lo) : -1;
        float bv_next = w_lo < w_hi ? bias_load(w_lo) : 0.f;
        uint32_t nstaged = 0;
        uint32_t sbias = bias_base + cg * GC * 4;
        for (int w = w_lo; w < w_hi; w += w_step) {
            const int ti = TRACE ? (w - w_lo) / w_step : 0;
            const bool tr = TRACE && ew == 0 && lane == 0;
            const Job& p = gp.jobs[find_job(w)];
            const CUtensorMap& tmC = p.tmC;
            const int lw = w - p.work0;
            const int t = lw / p.splits;
            const int m_blk = t % p.tiles_m, n_blk = t / p.tiles_m;
            const int col0 = n_blk * BN + cg * GC;        // first output column of this warp
            const int row0 = m_blk * BM + q * 32;         // first output row of this warp
            const int n_valid = p.N - col0;               // valid columns of this warp's share (<= 0: none)
            const int job_relu = p.relu, job_reduce = p.reduce_add, job_N = p.N, job_M = p.M;
            const __nv_bfloat16* const job_mask = p.mask;
            const int64_t job_ld_mask = p.ld_mask;
            float* const job_colsum = p.colsum;
            const float job_alpha = EPIX ? p.alpha : 1.f;
            const DropArgs job_drop = (EPIX && p.drop.thresh) ? drop_resolve(p.drop) : DropArgs();
            if (next_key != staged_key) {
                sbias = bias_base + (nstaged & 1u) * (BN * 4) + cg * GC * 4;
                if (gt < GC) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + gt * 4), "f"(bv_next) : "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + cg) : "memory");
                staged_key = next_key;
                ++nstaged;
            }
            if (w + w_step < w_hi) {
                next_key = bias_key(w + w_step);
                if (next_key != staged_key) bv_next = bias_load(w + w_step);
            }
            if (tr) T(ti, 5);  // epilogue ready for the tile
            mbar_wait(accf_bar(as), aphase);
            if (tr) T(ti, 6);  // accumulator complete
            tc_fence_after();
            const uint32_t t_row = tmem_base + as * BN + cg * GC + ((uint32_t)(q * 32) << 16);
            const bool live = row0 < job_M && n_valid > 0;  // warp-uniform: anything of this warp's share inside the output?
            if (live) {
#pragma unroll 1
                for (int c = 0; c < NCHUNK; ++c) {
                    if (c * CHUNK_COLS >= n_valid) break;
                    uint32_t r[CHUNK_COLS];
                    if constexpr (CHUNK_COLS == 32) tmem_ld32(t_row + c * CHUNK_COLS, r);
                    else tmem_ld16(t_row + c * CHUNK_COLS, r);
                    uint4 mk[CHUNK_COLS / 8];  // this row's slice of the ReLU mask (bf16), in flight with the TMEM load
                    if (job_mask) {
                        const int gr = row0 + lane;
                        const uint4* mp = reinterpret_cast<const uint4*>(job_mask + (int64_t)gr * job_ld_mask + col0 + c * CHUNK_COLS);
#pragma unroll
                        for (int j = 0; j < CHUNK_COLS / 8; ++j) mk[j] = gr < job_M ? __ldg(mp + j) : make_uint4(0, 0, 0, 0);
                    }
                    if (lane == 0) tma_wait_read<0>();  // this warp's previous store has read the staging tile
                    __syncwarp();
                    tmem_ld_wait();
                    const uint32_t sb = sbias + c * CHUNK_COLS * 4;
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                        float b0, b1, b2, b3;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(sb + j * 16));
                        float x0 = __uint_as_float(r[4 * j + 0]) + b0, x1 = __uint_as_float(r[4 * j + 1]) + b1;
                        float x2 = __uint_as_float(r[4 * j + 2]) + b2, x3 = __uint_as_float(r[4 * j + 3]) + b3;
                        if (job_relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
                        if (EPIX && job_alpha != 1.f) { x0 *= job_alpha; x1 *= job_alpha; x2 *= job_alpha; x3 *= job_alpha; }
                        if (EPIX && job_drop.thresh) {  // dropout behind the ReLU (the FFN's inner dropout): element row * N + col
                            const uint64_t e0 = (uint64_t)(row0 + lane) * (uint64_t)job_N + (uint64_t)(col0 + c * CHUNK_COLS + 4 * j);
                            x0 = drop_apply(job_drop, e0, x0); x1 = drop_apply(job_drop, e0 + 1, x1);
                            x2 = drop_apply(job_drop, e0 + 2, x2); x3 = drop_apply(job_drop, e0 + 3, x3);
                        }
                        r[4 * j + 0] = __float_as_uint(x0); r[4 * j + 1] = __float_as_uint(x1);
                        r[4 * j + 2] = __float_as_uint(x2); r[4 * j + 3] = __float_as_uint(x3);
                    }
                    if (job_mask) {  // y > 0 for a bf16 y  <=>  its bits, read as int16, are > 0
#pragma unroll
                        for (int j = 0; j < CHUNK_COLS / 8; ++j) {
                            const uint32_t w4[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if ((int16_t)(w4[e] & 0xffffu) <= 0) r[8 * j + 2 * e] = 0u;
                                if ((int16_t)(w4[e] >> 16) <= 0) r[8 * j + 2 * e + 1] = 0u;
                            }
                        }
                    }
                    const uint32_t srow = sbuf + lane * 64;
                    if (OUT_BF16) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // 16 B = 8 bf16 per store
                            uint32_t wv[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1]));
                                wv[e] = *reinterpret_cast<uint32_t*>(&bb);
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + swz(j, lane) * 16), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)  // 16 B = 4 fp32 per store
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + swz(j, lane) * 16), "r"(r[4 * j + 0]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    const int c0 = col0 + c * CHUNK_COLS;
                    if (lane == 0) {
                        if (job_reduce) tma_reduce_add_2d(&tmC, sbuf, c0, row0);
                        else tma_store_2d(&tmC, sbuf, c0, row0);
                        tma_commit();
                    }
                    if (job_colsum) {
                        // column sums of this warp's staged (rounded) 32-row tile; rows past M hold zeros (TMA zero-fills A, the
                        // fused column sum is only used without a bias).  The next chunk overwrites the tile after the __syncwarp
                        // that follows lane 0's wait, which every lane reaches after these reads.
                        if (lane < CHUNK_COLS) {
                            float sum = 0.f;
#pragma unroll 8
                            for (int rr = 0; rr < 32; ++rr) {
                                if (OUT_BF16) {
                                    uint16_t hv;
                                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(sbuf + rr * 64 + (swz(lane >> 3, rr) << 4) + (lane & 7) * 2));
                                    sum += __uint_as_float((uint32_t)hv << 16);
                                } else {
                                    float fv;
                                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(fv) : "r"(sbuf + rr * 64 + (swz(lane >> 2, rr) << 4) + (lane & 3) * 4));
                                    sum += fv;
                                }
                            }
                            if (c0 + lane < job_N) atomicAdd(job_colsum + c0 + lane, sum);
                        }
                    }
                }
            }
            // all TMEM reads of this warp are complete (tmem_ld_wait above): hand the accumulator stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acce_bar(as));
            if (tr) T(ti, 7);  // last store of the tile issued
            if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
        }
        if (lane == 0) tma_wait_all();
    } else if (CLK && warp >= 4 && warp < 8) {
        // ================= cluster split-K: park the partial tile, reduce the owned rows after the cluster barrier =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t pbase = base;  // [128 rows x 64 fp32], 16-byte units XOR-swizzled by (row & 15); the operand ring is free
        mbar_wait(accf_bar(0), 0);    // all MMAs of this CTA's k-range are complete: nothing reads the ring any more
        tc_fence_after();
        {
            uint32_t r[64];
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
            tmem_ld32(t_row, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
            tmem_ld32(t_row + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pbase + row * 256 + ((j ^ (row & 15)) << 4)),
                             "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
        }
        tc_fence_before();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        {
            const Job& p = gp.jobs[0];
            const int t = w_lo - p.work0;
            const int m_blk = t % p.tiles_m, n_blk = t / p.tiles_m;
            const int rows_per = BM / (int)csize;
            const int gt = (int)threadIdx.x - 128;
            for (int idx = gt; idx < rows_per * 16; idx += 128) {
                const int lr = (int)crank * rows_per + (idx >> 4), ch = idx & 15;
                const int grow = m_blk * BM + lr, gcol = n_blk * BN + ch * 4;
                if (grow >= p.M || gcol >= p.N) continue;
                const uint32_t local = pbase + lr * 256 + ((ch ^ (lr & 15)) << 4);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (uint32_t sr = 0; sr < csize; ++sr) {
                    uint32_t remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(sr));
                    float x0, x1, x2, x3;
                    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(remote) : "memory");
                    a0 += x0; a1 += x1; a2 += x2; a3 += x3;
                }
                float acc[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (gcol + e >= p.N) break;
                    float x = acc[e];
                    if (p.bias[0] != nullptr) x += __ldg(p.bias[0] + gcol + e);
                    if (p.relu) x = fmaxf(x, 0.f);
                    if (OUT_BF16) {
                        __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c_ptr) + (int64_t)grow * p.ldc + gcol + e;
                        if (p.accumulate) x += __bfloat162float(*c);
                        *c = __float2bfloat16_rn(x);
                    } else {
                        float* c = reinterpret_cast<float*>(p.c_ptr) + (int64_t)grow * p.ldc + gcol + e;
                        if (p.accumulate) x += *c;
                        *c = x;
                    }
                }
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");  // partial tiles stay alive until every reader is done
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else if (!CLK && !WEPI && warp >= 4 && warp < 4 + 4 * GROUPS) {
        // ================= epilogue: GROUPS column groups x 4 TMEM lane quadrants =================
        // Group g (warps 4+4g .. 7+4g) drains columns [GC g, GC g + GC) of the accumulator; warp (q = warp & 3)
        // of a group reads TMEM lanes [32 q, 32 q + 32) (the hardware ties a warp to lane quadrant warp % 4).
        const int ew = warp - 4;
        const int q = ew & 3;
        const int g = ew >> 2;
        const int row = q * 32 + lane;             // row of the 128-row tile owned by this thread
        const int gt = threadIdx.x - 128 - g * 128;  // thread index inside the group
        const uint32_t bar_id = 1 + g;
        const uint32_t sbuf_group = epi_base + g * C::EPI_BUFS * EPI_BYTES;  // the group's 128 x 128 B staging tile(s)
        uint32_t nchunk_done = 0;  // running chunk count of this group (EPI2: selects the staging tile)
        const uint32_t sbias = bias_base + g * GC * 4;  // this group's GC bias values (fp32)
        constexpr int ROWB = 128;                             // bytes of output per row per chunk (= the staging tile's row)
        constexpr int CHUNK_COLS = OUT_BF16 ? ROWB / 2 : ROWB / 4;
        // 16-byte unit j of row `row` inside the swizzled staging tile (CU_TENSOR_MAP_SWIZZLE_128B: address bits [4,7) ^= [7,10))
        auto swz = [](int j, int row) { return j ^ (row & 7); };
        constexpr int NCHUNK = GC / CHUNK_COLS;               // chunks per group
        int as = 0;
        uint32_t aphase = 0;
        for (int w = w_lo; w < w_hi; w += w_step) {
            const int ti = TRACE ? (w - w_lo) / w_step : 0;
            const bool tr = TRACE && ew == 0 && lane == 0;
            const Job& p = gp.jobs[find_job(w)];
            const CUtensorMap& tmC = p.tmC;
            const int lw = w - p.work0;
            const int split = lw % p.splits;
            const int t = lw / p.splits;
            const int m_blk = t % p.tiles_m, n_blk = t / p.tiles_m;
            const int n_valid = min(BN, p.N - n_blk * BN) - g * GC;  // valid columns of this group's share
            // job fields used per chunk, read once per tile (they live in the constant bank)
            const int job_relu = p.relu, job_reduce = p.reduce_add, job_N = p.N, job_M = p.M;
            const __nv_bfloat16* const job_mask = p.mask;
            const int64_t job_ld_mask = p.ld_mask;
            float* const job_colsum = p.colsum;
            const float job_alpha = EPIX ? p.alpha : 1.f;
            const DropArgs job_drop = (EPIX && p.drop.thresh) ? drop_resolve(p.drop) : DropArgs();
            {   // stage the bias slice (zero where there is none / out of range): no per-element predicates below.
                // Safe to overwrite: every thread of the group passed the previous tile's last bar.sync, which
                // follows all of that tile's bias reads.
                const int col = n_blk * BN + g * GC + gt;
                float bv = 0.f;
                if (split == 0 && col < p.N && gt < GC)
                    for (int term = 0; term < p.nterms; ++term)
                        if (p.bias[term] != nullptr) bv += __ldg(p.bias[term] + col);
                if (gt < GC) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + gt * 4), "f"(bv) : "memory");
            }
            if (tr) T(ti, 5);  // epilogue ready for the tile
            mbar_wait(accf_bar(as), aphase);
            if (tr) T(ti, 6);  // accumulator complete
            tc_fence_after();
            const uint32_t t_row = tmem_base + as * BN + g * GC + ((uint32_t)(q * 32) << 16);
            bool released = false;
            auto release_acc = [&]() {  // all TMEM reads of this accumulator stage are complete
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acce_bar(as));
                released = true;
            };
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                if (c * CHUNK_COLS >= n_valid) break;  // uniform across the group
                uint32_t r[CHUNK_COLS];
#pragma unroll
                for (int h = 0; h < CHUNK_COLS / 32; ++h)
                    tmem_ld32(t_row + c * CHUNK_COLS + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[h * 32]));
                uint4 mk[CHUNK_COLS / 8];  // this row's slice of the ReLU mask (bf16), in flight with the TMEM loads
                if (job_mask) {
                    const int gr = m_blk * BM + row;
                    const uint4* mp = reinterpret_cast<const uint4*>(job_mask + (int64_t)gr * job_ld_mask + n_blk * BN + g * GC + c * CHUNK_COLS);
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 8; ++j) mk[j] = gr < job_M ? __ldg(mp + j) : make_uint4(0, 0, 0, 0);
                }
                // the staging tile must have been read by the TMA store that used it last (EPI2: two tiles alternate, so
                // the most recent store may still be in flight)
                const uint32_t sbuf_base = sbuf_group + (EPI2 ? (nchunk_done & 1u) * EPI_BYTES : 0u);
                if (gt == 0) { if (EPI2) tma_wait_read<1>(); else tma_wait_read<0>(); }
                if (c == NCHUNK - 1 || (c + 1) * CHUNK_COLS >= n_valid) {
                    // last TMEM read of this accumulator stage: hand it back to the MMA warp as early as possible
                    tmem_ld_wait();
                    release_acc();
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                tmem_ld_wait();
                const uint32_t srow = sbuf_base + row * ROWB;
                const uint32_t sb = sbias + c * CHUNK_COLS * 4;
#pragma unroll
                for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                    float b0, b1, b2, b3;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(sb + j * 16));
                    float x0 = __uint_as_float(r[4 * j + 0]) + b0, x1 = __uint_as_float(r[4 * j + 1]) + b1;
                    float x2 = __uint_as_float(r[4 * j + 2]) + b2, x3 = __uint_as_float(r[4 * j + 3]) + b3;
                    if (job_relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
                    if (EPIX && job_alpha != 1.f) { x0 *= job_alpha; x1 *= job_alpha; x2 *= job_alpha; x3 *= job_alpha; }
                    if (EPIX && job_drop.thresh) {  // dropout behind the ReLU (the FFN's inner dropout): element row * N + col
                        const uint64_t e0 = (uint64_t)(m_blk * BM + row) * (uint64_t)job_N + (uint64_t)(n_blk * BN + g * GC + c * CHUNK_COLS + 4 * j);
                        x0 = drop_apply(job_drop, e0, x0); x1 = drop_apply(job_drop, e0 + 1, x1);
                        x2 = drop_apply(job_drop, e0 + 2, x2); x3 = drop_apply(job_drop, e0 + 3, x3);
                    }
                    r[4 * j + 0] = __float_as_uint(x0); r[4 * j + 1] = __float_as_uint(x1);
                    r[4 * j + 2] = __float_as_uint(x2); r[4 * j + 3] = __float_as_uint(x3);
                }
                if (job_mask) {  // y > 0 for a bf16 y  <=>  its bits, read as int16, are > 0
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 8; ++j) {
                        const uint32_t w4[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if ((int16_t)(w4[e] & 0xffffu) <= 0) r[8 * j + 2 * e] = 0u;
                            if ((int16_t)(w4[e] >> 16) <= 0) r[8 * j + 2 * e + 1] = 0u;
                        }
                    }
                }
                if (OUT_BF16) {
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 8; ++j) {  // 16 B = 8 bf16 per store
                        uint32_t wv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1]));
                            wv[e] = *reinterpret_cast<uint32_t*>(&bb);
                        }
                        const int chunk = swz(j, row);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + chunk * 16), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 4; ++j) {  // 16 B = 4 fp32 per store
                        const int chunk = swz(j, row);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + chunk * 16), "r"(r[4 * j + 0]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                const int c0 = n_blk * BN + g * GC + c * CHUNK_COLS;
                if (gt == 0) {
                    if (job_reduce) tma_reduce_add_2d(&tmC, sbuf_base, c0, m_blk * BM);
                    else tma_store_2d(&tmC, sbuf_base, c0, m_blk * BM);
                    tma_commit();
                }
                if (job_colsum) {
                    // column sums of the staged (rounded) tile: 128 threads = CHUNK_COLS columns x (128 / CHUNK_COLS)
                    // row slabs; rows past M hold zeros (TMA zero-fills A).  The next chunk overwrites the staging
                    // tile only after the group's next bar.sync, which this thread reaches after its reads.
                    constexpr int SLABS = 128 / CHUNK_COLS, ROWS = BM / SLABS;
                    const int cc = gt % CHUNK_COLS, r0 = (gt / CHUNK_COLS) * ROWS;
                    float sum = 0.f;
#pragma unroll 8
                    for (int rr = r0; rr < r0 + ROWS; ++rr) {
                        if (OUT_BF16) {
                            uint16_t hv;
                            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(sbuf_base + rr * ROWB + (swz(cc >> 3, rr) << 4) + (cc & 7) * 2));
                            sum += __uint_as_float((uint32_t)hv << 16);
                        } else {
                            float fv;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(fv) : "r"(sbuf_base + rr * ROWB + (swz(cc >> 2, rr) << 4) + (cc & 3) * 4));
                            sum += fv;
                        }
                    }
                    if (c0 + cc < job_N) atomicAdd(job_colsum + c0 + cc, sum);
                }
                if (EPI2) ++nchunk_done;
            }
            if (!released) release_acc();  // this group's half lies entirely outside N
            if (tr) T(ti, 7);  // last store of the tile issued
            if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
        }
        if (gt == 0) tma_wait_all();
    }
    if (CLK && (warp < 4 || warp >= 8)) {  // the control warps take part in the two cluster barriers of the epilogue warps
        __syncwarp();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_map(CUtensorMap* tm, const void* ptr, bool bf16, int64_t rows, int64_t cols, int64_t ld, int box_cols,
             int box_rows, bool swizzle64) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_err(STCAT_EINVAL, "tensor map: cuTensorMapEncodeTiled not available from the driver");
    const int es = bf16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(STCAT_EINVAL, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows, (long long)cols, (long long)ld);
    return 0;
}

int make_map_3d(CUtensorMap* tm, const void* ptr, int64_t batch, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                int box_rows, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_err(STCAT_EINVAL, "tensor map: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(STCAT_EINVAL, "cuTensorMapEncodeTiled(3d) failed (%d) batch=%lld rows=%lld cols=%lld ld=%lld", (int)r, (long long)batch, (long long)rows, (long long)cols, (long long)ld);
    return 0;
}

}  // namespace tc

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Shapes / alignments the tensor-core kernel takes; everything else (tiny heads such as N = 1, 2, 4, odd leading
// dimensions) stays on the exact SIMT kernel.
int gemm_tc_supported(int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, const void* A, const void* B,
                      const void* C, int a_mn_major, int b_mn_major) {
    (void)a_mn_major; (void)b_mn_major;
    if (getenv("STCAT_DISABLE_TC")) return 0;
    if (M < 1 || N < 8 || K < 8) return 0;
    if ((int64_t)M * N * K < (int64_t)(1 << 18)) return 0;  // launch-latency regime: SIMT kernel is as fast
    if (!aligned16(A) || !aligned16(B) || !aligned16(C)) return 0;
    if ((lda % 8) || (ldb % 8) || (ldc % 8)) return 0;  // 16 B row pitch for bf16 (fp32 C: 32 B, stricter than needed)
    return 1;
}

// one GEMM of a grouped launch as the C-ABI layer describes it (capi.cu)
struct TcTerm { const void* A; int64_t lda; const void* B; int64_t ldb; const float* bias; int K; };
struct TcJob {
    TcTerm term[tc::MAX_TERMS];
    int nterms;
    void* C; int64_t ldc; int M, N;
    int relu, accumulate;
    GemmEpilogue epi;
};

// Upper bound on the CTAs of the following tcgen05 GEMM launches (0 = all SMs).  The decoder's memory-side key / value
// projections run on a side stream next to the latency-bound query chains: a persistent GEMM on every SM leaves no SM for
// the chains' small kernels (its CTAs hold all registers / shared memory of their SM), so those launches are capped.
static int g_gemm_sm_limit = 0;
void gemm_tc_set_sm_limit(int n) { g_gemm_sm_limit = n > 0 ? n : 0; }

static long long* g_gemm_trace = nullptr;
void gemm_tc_set_trace(long long* buf) { g_gemm_trace = buf; }

// one instantiation: shared-memory attribute once, programmatic dependent launch
template <bool AMN, bool BMN, bool OBF, int BN, int NJ, bool TRACE, int EPIMODE, bool BRES, bool EPIX = false>
static int launch_inst(const tc::GroupParams<NJ>& gp, int grid, cudaStream_t st) {
    using namespace tc;
    constexpr int SMEM = Cfg<BN, EPIMODE, BRES>::SMEM_BYTES;
    auto kern = gemm_tc_kernel<AMN, BMN, OBF, BN, NJ, TRACE, EPIMODE, BRES, false, EPIX>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return set_err((int)e, "gemm_tc: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    cudaError_t le = launch_pdl(kern, dim3(grid), dim3(Cfg<BN, EPIMODE, BRES>::THREADS), SMEM, st, gp);
    if (le != cudaSuccess) return set_err((int)le, "gemm_tc_kernel launch: %s", cudaGetErrorString(le));
    return check_launch("gemm_tc_kernel");
}

// cluster split-K launch (BN = 64, one job): `clk` CTAs per output tile
template <bool BMN, bool OBF>
static int launch_clk(const tc::GroupParams<1>& gp, int tiles, int clk, cudaStream_t st) {
    using namespace tc;
    constexpr int SMEM = Cfg<64, 0, false>::SMEM_BYTES;
    auto kern = gemm_tc_kernel<false, BMN, OBF, 64, 1, false, 0, false, true>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return set_err((int)e, "gemm_tc<clk>: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles * clk);
    cfg.blockDim = dim3(Cfg<64, 0, false>::THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = clk;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, gp);
    if (le != cudaSuccess) return set_err((int)le, "gemm_tc_kernel<clk> launch: %s", cudaGetErrorString(le));
    return check_launch("gemm_tc_kernel<clk>");
}

// epilogue / operand-residency variant of a launch (see Cfg)
enum TcMode { TC_PLAIN = 0, TC_EPI2 = 1, TC_BRES = 2, TC_WEPI = 3 };

template <bool AMN, bool BMN, bool OBF, int BN, int NJ>
static int launch_tc(const tc::GroupParams<NJ>& gp, int grid, cudaStream_t st, int mode) {
    using namespace tc;
    bool extras = false;
    for (int j = 0; j < gp.njobs; ++j) extras = extras || gp.jobs[j].alpha != 1.f || gp.jobs[j].drop.thresh != 0;
    if (extras) {
        if constexpr (NJ == 1 && OBF && !AMN) {
            if constexpr (BN == 256) {
                if (mode == TC_BRES) return launch_inst<AMN, BMN, OBF, BN, NJ, false, 3, true, true>(gp, grid, st);
                return launch_inst<AMN, BMN, OBF, BN, NJ, false, 3, false, true>(gp, grid, st);
            } else {
                return launch_inst<AMN, BMN, OBF, BN, NJ, false, 0, false, true>(gp, grid, st);
            }
        } else {
            return set_err(STCAT_ESHAPE, "gemm_tc: the scaled / dropout epilogue is instantiated for single launches with a bf16 output only");
        }
    }
    if constexpr (BN == 256) {
        if constexpr (!AMN && OBF) {
            if (mode == TC_BRES) {
                if constexpr (!BMN && NJ == 1) {  // diagnostics: traced instantiation of the plain forward GEMM
                    if (g_gemm_trace != nullptr) {
                        GroupParams<NJ> gt = gp;
                        gt.trace = g_gemm_trace;
                        return launch_inst<AMN, BMN, OBF, BN, NJ, true, 3, true>(gt, grid, st);
                    }
                }
                return launch_inst<AMN, BMN, OBF, BN, NJ, false, 3, true>(gp, grid, st);
            }
        }
        if constexpr (!AMN && !BMN && NJ == 1) {  // diagnostics: traced instantiations of the plain forward GEMM
            if (g_gemm_trace != nullptr) {
                GroupParams<NJ> gt = gp;
                gt.trace = g_gemm_trace;
                if (mode == TC_WEPI) return launch_inst<AMN, BMN, OBF, BN, NJ, true, 3, false>(gt, grid, st);
                if (mode == TC_EPI2) return launch_inst<AMN, BMN, OBF, BN, NJ, true, 1, false>(gt, grid, st);
                return launch_inst<AMN, BMN, OBF, BN, NJ, true, 0, false>(gt, grid, st);
            }
        }
        if (mode == TC_WEPI) return launch_inst<AMN, BMN, OBF, BN, NJ, false, 3, false>(gp, grid, st);
        if (mode == TC_EPI2) return launch_inst<AMN, BMN, OBF, BN, NJ, false, 1, false>(gp, grid, st);
    }
    return launch_inst<AMN, BMN, OBF, BN, NJ, false, 0, false>(gp, grid, st);
}

// Fills one device-side Job.  `work0` is the running prefix of work items; returns <0 on error via set_err code.
template <int BN>
static int fill_job(tc::Job& J, const TcJob& in, int a_mn_major, int b_mn_major, bool out_bf16, int& work, cudaStream_t st,
                    bool warp_epi) {
    using namespace tc;
    int rc;
    const int M = in.M, N = in.N;
    memset(&J, 0, sizeof(J));
    int kb_max = 0;
    for (int t = 0; t < in.nterms; ++t) {
        const TcTerm& T = in.term[t];
        // A: K-major -> tensor [M rows, K cols], box {BK, BM};  MN-major -> tensor [K rows, M cols], box {64, BK}
        rc = a_mn_major ? make_map(&J.tmA[t], T.A, true, T.K, M, T.lda, 64, BK) : make_map(&J.tmA[t], T.A, true, M, T.K, T.lda, BK, BM);
        if (rc) return rc;
        rc = b_mn_major ? make_map(&J.tmB[t], T.B, true, T.K, N, T.ldb, 64, BK) : make_map(&J.tmB[t], T.B, true, N, T.K, T.ldb, BK, BN);
        if (rc) return rc;
        J.bias[t] = T.bias;
        J.kb[t] = (T.K + BK - 1) / BK;
        kb_max = J.kb[t] > kb_max ? J.kb[t] : kb_max;
    }
    // output staging tile: [BM rows x 128 B] (128B swizzle) per epilogue group or, for the warp epilogue, [32 x 64 B] (64B swizzle)
    rc = make_map(&J.tmC, in.C, out_bf16, M, N, in.ldc, (out_bf16 ? 64 : 32) / (warp_epi ? 2 : 1), warp_epi ? 32 : BM, warp_epi);
    if (rc) return rc;
    J.nterms = in.nterms;
    J.M = M; J.N = N;
    J.tiles_m = (M + BM - 1) / BM;
    J.tiles_n = (N + BN - 1) / BN;
    const int tiles = J.tiles_m * J.tiles_n;
    const int sms = num_sms();
    int splits = 1;
    // Split-K only for weight gradients (A MN-major: the contraction runs over all tokens while the output is a
    // few tiles); they are accumulated into fp32 gradient buffers anyway.  Forward / data-gradient GEMMs keep a
    // fixed summation order (bit-reproducible results); their few-tile cases use the BN = 64 shape instead.
    if (in.nterms == 1 && a_mn_major && !in.relu && !out_bf16 && tiles * 2 <= sms && kb_max >= 8) {
        splits = sms / tiles;
        const int max_by_k = kb_max / 4;  // at least 4 k-blocks (256 contraction elements) per split
        if (splits > max_by_k) splits = max_by_k;
        if (splits < 1) splits = 1;
    }
    static const bool no_splitk = getenv("STCAT_NO_SPLITK") != nullptr;  // diagnosis: deterministic summation order
    if (no_splitk) splits = 1;
    J.kb_per_split = (kb_max + splits - 1) / splits;
    J.splits = (kb_max + J.kb_per_split - 1) / J.kb_per_split;
    J.relu = in.relu;
    J.mask = (const __nv_bfloat16*)in.epi.relu_mask;
    J.ld_mask = in.epi.ld_mask;
    J.colsum = in.epi.colsum;
    J.alpha = in.epi.alpha;
    J.drop = in.epi.drop;
    if (J.drop.thresh && (in.ldc != N || in.accumulate || J.splits > 1))
        return set_err(STCAT_ESHAPE, "gemm_tc: fused dropout epilogue needs a contiguous output, no accumulate, no split-K");
    if ((J.mask || J.colsum) && (a_mn_major || in.accumulate || N % 64 != 0))
        return set_err(STCAT_ESHAPE, "gemm_tc: fused ReLU-mask / column-sum epilogue needs N %% 64 == 0, no accumulate, K-major A");
    if (in.relu && in.accumulate) return set_err(STCAT_ESHAPE, "gemm_tc: relu with accumulate is not supported");
    J.reduce_add = (in.accumulate || J.splits > 1) ? 1 : 0;
    if (J.splits > 1 && !in.accumulate) {
        cudaError_t e = cudaMemset2DAsync(in.C, (size_t)in.ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
        if (e != cudaSuccess) return set_err((int)e, "gemm_tc memset: %s", cudaGetErrorString(e));
    }
    J.c_ptr = in.C;
    J.ldc = in.ldc;
    J.accumulate = in.accumulate;
    J.work0 = work;
    work += tiles * J.splits;
    return 0;
}

template <int BN, int NJ>
static int gemm_tc_launch_jobs(const TcJob* jobs, int njobs, int a_mn_major, int b_mn_major, int out_dtype, cudaStream_t st) {
    using namespace tc;
    const bool out_bf16 = out_dtype == STCAT_BF16;
    static thread_local GroupParams<NJ> gp;  // ~1 KB per job: kept off the stack, rebuilt per call
    int kmax = 0;  // longest K loop of the launch
    bool single_short = true;  // every job: one term with K <= 256 (its [256 x K] weight tile fits the resident region)
    long tiles256 = 0;
    for (int j = 0; j < njobs; ++j) {
        int k = 0;
        for (int t = 0; t < jobs[j].nterms; ++t) k += jobs[j].term[t].K;
        kmax = k > kmax ? k : kmax;
        if (jobs[j].nterms != 1 || jobs[j].term[0].K > Cfg<256, 3, true>::BRES_KB * BK) single_short = false;
        tiles256 += (long)((jobs[j].M + BM - 1) / BM) * ((jobs[j].N + 255) / 256);
    }
    const int sms = (g_gemm_sm_limit > 0 && g_gemm_sm_limit < num_sms()) ? g_gemm_sm_limit : num_sms();
    // Variant (BN = 256): weight-resident (BRES) when every job is a single K <= 256 product with a bf16 output and each CTA
    // gets at least two tiles (else nothing is re-used); otherwise two staging tiles (EPI2) for short K loops -- measured
    // (profiles/r2_a_validate_staged.log): K = 256: FFN linear1 23.2 -> 19.9 us, slower by 3 % where the K loop is long
    // (K = 2048) and the fourth operand stage matters more.  STCAT_GEMM_EPI2 / STCAT_GEMM_BRES = 0 / 1 force them off / on.
    static const char* epi2_env = getenv("STCAT_GEMM_EPI2");
    static const char* bres_env = getenv("STCAT_GEMM_BRES");
    int mode = TC_PLAIN;
    if (BN == 256) {
        const bool bres_ok = single_short && out_bf16 && !a_mn_major;
        const bool bres = bres_ok && (bres_env ? atoi(bres_env) != 0 : tiles256 >= 2L * sms);
        const bool epi2 = epi2_env ? atoi(epi2_env) != 0 : kmax <= 512;
        static const char* wepi_env = getenv("STCAT_GEMM_WEPI");
        const bool wepi = wepi_env ? atoi(wepi_env) != 0 : true;
        mode = bres ? TC_BRES : (wepi ? TC_WEPI : (epi2 ? TC_EPI2 : TC_PLAIN));
        for (int j = 0; j < njobs; ++j)  // the scaled / dropout epilogue exists in the warp epilogue only
            if ((jobs[j].epi.alpha != 1.f || jobs[j].epi.drop.thresh) && mode != TC_BRES) mode = TC_WEPI;
    }
    int work = 0;
    for (int j = 0; j < njobs; ++j) {
        int rc = fill_job<BN>(gp.jobs[j], jobs[j], a_mn_major, b_mn_major, out_bf16, work, st, mode == TC_BRES || mode == TC_WEPI);
        if (rc) return rc;
    }
    gp.njobs = njobs;
    gp.total = work;
    gp.trace = nullptr;
    if constexpr (BN == 64 && NJ == 1) {
        // cluster split-K: few tiles, long contraction (see the kernel's CLK note); STCAT_GEMM_CLK=0 turns it off
        static const bool clk_on = !(getenv("STCAT_GEMM_CLK") && atoi(getenv("STCAT_GEMM_CLK")) == 0);
        const TcJob& j0 = jobs[0];
        const int kb = (j0.term[0].K + BK - 1) / BK;
        if (clk_on && !a_mn_major && j0.nterms == 1 && !j0.epi.relu_mask && !j0.epi.colsum && j0.epi.alpha == 1.f && !j0.epi.drop.thresh &&
            kb >= 16 && gp.jobs[0].splits == 1) {
            int clk = 0;
            for (int sft = 3; sft >= 1 && !clk; --sft)
                if (kb % (1 << sft) == 0 && work * (1 << sft) <= sms) clk = 1 << sft;
            if (clk) {
                gp.jobs[0].kb_per_split = kb / clk;
                if (b_mn_major) return out_bf16 ? launch_clk<true, true>(gp, work, clk, st) : launch_clk<true, false>(gp, work, clk, st);
                return out_bf16 ? launch_clk<false, true>(gp, work, clk, st) : launch_clk<false, false>(gp, work, clk, st);
            }
        }
    }
    const int grid = work < sms ? work : sms;
    if (!a_mn_major && !b_mn_major)
        return out_bf16 ? launch_tc<false, false, true, BN, NJ>(gp, grid, st, mode) : launch_tc<false, false, false, BN, NJ>(gp, grid, st, mode);
    if (!a_mn_major && b_mn_major)
        return out_bf16 ? launch_tc<false, true, true, BN, NJ>(gp, grid, st, mode) : launch_tc<false, true, false, BN, NJ>(gp, grid, st, mode);
    if (a_mn_major && b_mn_major)
        return out_bf16 ? launch_tc<true, true, true, BN, NJ>(gp, grid, st, mode) : launch_tc<true, true, false, BN, NJ>(gp, grid, st, mode);
    return set_err(STCAT_ESHAPE, "gemm_tc: A MN-major with B K-major is not instantiated");
}

// tile shape: the latency shape when the 128 x 256 tiling would leave most SMs idle and the K loops are short
static bool pick_skinny(const TcJob* jobs, int njobs) {
    static const int force_bn = getenv("STCAT_TC_BN") ? atoi(getenv("STCAT_TC_BN")) : 0;
    if (force_bn) return force_bn == 64;
    int tiles256 = 0, kb = 0;
    for (int j = 0; j < njobs; ++j) {
        tiles256 += ((jobs[j].M + tc::BM - 1) / tc::BM) * ((jobs[j].N + 255) / 256);
        int k = 0;
        for (int t = 0; t < jobs[j].nterms; ++t) k += (jobs[j].term[t].K + tc::BK - 1) / tc::BK;
        kb = k > kb ? k : kb;
    }
    return tiles256 * 4 <= num_sms() && kb <= 64;
}

int gemm_tc(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* C,
            int64_t ldc, int out_dtype, const float* bias, int M, int N, int K, int relu, int accumulate,
            cudaStream_t st, const GemmEpilogue* epi) {
    TcJob job;
    job.term[0] = TcTerm{A, lda, B, ldb, bias, K};
    job.nterms = 1;
    job.C = C; job.ldc = ldc; job.M = M; job.N = N;
    job.relu = relu; job.accumulate = accumulate;
    if (epi) job.epi = *epi;
    if (pick_skinny(&job, 1)) return gemm_tc_launch_jobs<64, 1>(&job, 1, a_mn_major, b_mn_major, out_dtype, st);
    return gemm_tc_launch_jobs<256, 1>(&job, 1, a_mn_major, b_mn_major, out_dtype, st);
}

// Grouped launch: up to MAX_JOBS independent multi-term GEMMs sharing operand majors and output dtype.
int gemm_tc_group(const TcJob* jobs, int njobs, int a_mn_major, int b_mn_major, int out_dtype, cudaStream_t st) {
    if (njobs < 1 || njobs > tc::MAX_JOBS) return set_err(STCAT_EINVAL, "gemm_tc_group: njobs=%d (1..%d)", njobs, tc::MAX_JOBS);
    if (pick_skinny(jobs, njobs)) return gemm_tc_launch_jobs<64, tc::MAX_JOBS>(jobs, njobs, a_mn_major, b_mn_major, out_dtype, st);
    return gemm_tc_launch_jobs<256, tc::MAX_JOBS>(jobs, njobs, a_mn_major, b_mn_major, out_dtype, st);
}

}  // namespace stcat
