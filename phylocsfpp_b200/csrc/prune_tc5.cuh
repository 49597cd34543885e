// prune_tc5.cuh — k_prune_tc5: Felsenstein pruning on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators and the A operand in TMEM, P tiles and leaf tables streamed by bulk TMA, cherry messages gathered from
// L2-resident tables).  FP32-class arithmetic with per-window log-scaling; the FP64 DMMA kernel (k_prune) stays the parity anchor.
//
// Formulation (ensure_alpha, fixed_lik.hpp:125-164).  For 128 codon windows at a time every edge above a NON-CHERRY inner
// node c is one GEMM  D[w][a] = sum_b A[w][b] * P_c[a][b]  on the tensor core (M = 128 windows, K = 64 child states):
// A = alpha_c split into TF32 hi + lo (per-window power-of-two normalised, the exponent kept as an integer);
// B = [hi(P_c) | lo(P_c)] side by side (N = 128): per k-step one N = 128 MMA with hi(A) gives hi*hi | hi*lo and one
// N = 64 MMA with lo(A) adds lo*hi, in FP32 accumulators; msg = D[:, 0:64] + D[:, 64:128] (~2^-21 per product).  A TS-form MMA
// reads its 4 KB A slab from TMEM at 64 B/clk, i.e. ~64 cycles whatever its N <= 128 (tools/tc5_probe.cu), so the
// side-by-side layout is what makes two instructions per k-step enough.
// LEAF edges are not GEMMs: the message of a leaf with codon x is column x of P_l (all ones for a gap/N codon, the row
// sums of fixed_lik.hpp:111-118).  The epilogue threads gather it from a 17 KB per-leaf table that TMA streams into shared
// memory in program order — and they do so WHILE their GEMM runs, so leaves cost no tensor time and no latency.
// CHERRY edges are not GEMMs either (round 2): the message of the edge above a node whose two children are leaves depends only on
// the two codons, P_c . (P_l[:, x] * P_r[:, y]), 65 x 65 rows of 64 floats tabulated once per model in FP64 (k_build_cherry) and
// kept in global memory (1.08 MB per cherry and ECM: L2 resident).  Every epilogue warp copies the 32 rows of its own windows
// with 16-byte cp.async (sixteen lanes per 256-byte row: full sectors) into a private, XOR-swizzled 8 KB staging area one
// cherry ahead of the program, and each thread then reads its own row without bank conflicts.  That removes 17 of 56 GEMMs,
// 34 of 58 leaf gathers and half of the stack pushes for 58mammals (28 of 98, 56 of 100 for 100vertebrates).
//
// One persistent CTA per SM works on a PAIR of 128-window tiles (chains X and Y) that share both TMA rings: while the
// tensor core runs chain Y's GEMM of step g, chain X's epilogue turns D_g into A_{g+1} (and vice versa).
//   warps 0-3   epilogue of chain X: thread = window = TMEM lane; the 64-state partial lives in registers
//   warps 4-7   epilogue of chain Y        TMEM columns 256c..256c+255: two 128-column regions, A_g in one, D_g in the
//   warp  8     MMA issue (one elected lane)    other; A_{g+1} overwrites D_g in place
//   warp  9     TMA producer, inner-edge tiles (32 KB each, 2- or 3-stage full/empty mbarrier ring)
//   warp 10     TMA producer, leaf tables (17 KB each, up to 6 stages: both chains read every table, so a stage lives
//               until the trailing chain has used it)
// The producer warpgroup gives its registers to the epilogue warpgroups (setmaxnreg 40 / 232).
// Waiting sibling partials (stack depth = Strahler number over the non-cherry inner nodes - 1) spill to an L2-resident scratch
// in global memory, every thread only ever touching its own column: the push is issued after the next A has been handed to
// the tensor core, the pop is prefetched while the GEMM runs, so neither is on the critical path.
#pragma once

#include "kernels.cuh"
#include "tc5.cuh"

namespace pcsf {

constexpr int T5_MAX_NSTAGE = 3;         // inner-edge tile ring (stages chosen at model creation: what fits)
constexpr int T5_TILE_BYTES = 32768;
constexpr int T5_MAX_NLSTAGE = 6;        // leaf table ring
constexpr int T5_LEAF_BYTES = T5_LEAF_FLOATS * 4;
constexpr int T5_THREADS = 384;
constexpr int T5_STACK_ENTRY_FLOATS = 64 * 128 + 128;   // 128 windows x 64 states + 128 exponents

struct PruneTc5Args {
    WinSpace ws;
    const uint32_t *uniq;
    const uint32_t *n_unique;
    const uint32_t *steps;
    int n_steps, max_stack;
    uint32_t start;              // src1 | src2 << 8: the chain start the program begins with
    int n_leaf_tabs, n_cherry;
    int nstage, nlstage;         // ring depths
    const float *pstream[2];     // [n_steps][8192]
    const float *leaftab[2];     // [n_leaf_tabs][T5_LEAF_FLOATS], program order
    const float *cherrytab[2];   // [n_cherry][T5_CHERRY_ROWS][64], program order
    const uint16_t *cherry_leaves;   // [n_cherry] left leaf | right leaf << 8
    const double *pi[2];
    double *logz[2];
    float *scratch;              // [grid][2][max_stack][T5_STACK_ENTRY_FLOATS]
};

constexpr int T5_CHERRY_STAGE_BYTES = 32 * 256;     // per epilogue warp: the rows of its 32 windows

__host__ __device__ inline size_t prune_tc5_smem_bytes(int nl, int n_steps, int n_cherry, int nstage, int nlstage) {
    size_t b = (size_t)nstage * T5_TILE_BYTES + (size_t)nlstage * T5_LEAF_BYTES;
    b += (size_t)8 * T5_CHERRY_STAGE_BYTES;                 // cherry row staging
    b += (size_t)2 * nl * 128;                              // leaf codon ids of both tiles
    b += (size_t)(((n_steps + 1) * 4 + 15) / 16) * 16;      // steps
    b += (size_t)(((n_cherry + 1) * 2 + 15) / 16) * 16;     // cherry leaves
    b += 2 * 64 * 8;                                        // pi
    b += 32 * 8;                                            // mbarriers + TMEM base (2*3 + 2*6 + 4 + 2 barriers)
    return b;
}
// Deepest rings that fit into one SM's shared memory.
inline void prune_tc5_pick_stages(int nl, int n_steps, int n_cherry, int *nstage, int *nlstage) {
    *nstage = T5_MAX_NSTAGE; *nlstage = T5_MAX_NLSTAGE;
    while (prune_tc5_smem_bytes(nl, n_steps, n_cherry, *nstage, *nlstage) > 227 * 1024 && *nlstage > 4) --*nlstage;
    while (prune_tc5_smem_bytes(nl, n_steps, n_cherry, *nstage, *nlstage) > 227 * 1024 && *nstage > 2) --*nstage;
    while (prune_tc5_smem_bytes(nl, n_steps, n_cherry, *nstage, *nlstage) > 227 * 1024 && *nlstage > 3) --*nlstage;
}

// Cherry tables: T[k][x * 65 + y][a] = sum_b P_c[a][b] * u_x[b] * v_y[b] with u_x = P_l[:, x], v_y = P_r[:, y] (all ones for
// index 64, the gap/N codon — the leaf rule of this path), FP64 accumulation, FP32 storage.  grid (65, n_cherry), 64 threads.
__global__ void __launch_bounds__(64) k_build_cherry(const double *__restrict__ cherry_P, const double *__restrict__ leafPT,
                                                     const uint16_t *__restrict__ cherry_leaves, float *__restrict__ tab) {
    __shared__ double u[64];
    __shared__ double v[65][64];
    const int k = blockIdx.y, x = blockIdx.x, a = threadIdx.x;
    const int l = cherry_leaves[k] & 0xff, r = cherry_leaves[k] >> 8;
    u[a] = x < 64 ? leafPT[((size_t)l * 65 + x) * 64 + a] : 1.0;
    for (int y = 0; y < 65; ++y) v[y][a] = y < 64 ? leafPT[((size_t)r * 65 + y) * 64 + a] : 1.0;
    __syncthreads();
    double pa[64];
#pragma unroll
    for (int b = 0; b < 64; ++b) pa[b] = cherry_P[(size_t)k * 4096 + a * 64 + b] * u[b];
    float *out = tab + ((size_t)k * T5_CHERRY_ROWS + (size_t)x * 65) * 64 + a;
    for (int y = 0; y < 65; ++y) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < 64; ++b) s += pa[b] * v[y][b];
        out[(size_t)y * 64] = (float)s;
    }
}

__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(T5_THREADS, 1) k_prune_tc5(const PruneTc5Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sp_ = smem;
    const uint32_t T5_NSTAGE = a.nstage, T5_NLSTAGE = a.nlstage;
    unsigned char *stage_buf = sp_; sp_ += (size_t)T5_NSTAGE * T5_TILE_BYTES;
    unsigned char *leaf_buf = sp_; sp_ += (size_t)T5_NLSTAGE * T5_LEAF_BYTES;
    unsigned char *cherry_buf = sp_; sp_ += (size_t)8 * T5_CHERRY_STAGE_BYTES;
    uint8_t *ids = sp_; sp_ += (size_t)2 * a.ws.nl * 128;
    uint32_t *steps = reinterpret_cast<uint32_t *>(sp_); sp_ += (size_t)(((a.n_steps + 1) * 4 + 15) / 16) * 16;
    uint16_t *cherry_leaves = reinterpret_cast<uint16_t *>(sp_); sp_ += (size_t)(((a.n_cherry + 1) * 2 + 15) / 16) * 16;
    double *s_pi = reinterpret_cast<double *>(sp_); sp_ += 2 * 64 * 8;
    uint64_t *full = reinterpret_cast<uint64_t *>(sp_);
    uint64_t *empty = full + T5_MAX_NSTAGE;
    uint64_t *lfull = empty + T5_MAX_NSTAGE;
    uint64_t *lempty = lfull + T5_MAX_NLSTAGE;
    uint64_t *a_ready = lempty + T5_MAX_NLSTAGE;   // [2 chains][2 halves] epilogue -> MMA: states 0..31 / 32..63 of the step's A are in TMEM
    uint64_t *d_ready = a_ready + 4;           // [2] MMA -> epilogue: D of the step is complete
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(d_ready + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef PCSF_TC5_TRACE
    __shared__ long long trace[3][128][4];
#define T5_TRACE(who, s, i) do { if (blockIdx.x == 0 && first_seq && (s) < 128) trace[who][s][i] = clock64(); } while (0)
#else
#define T5_TRACE(who, s, i) do { } while (0)
#endif
    if (tid == 0) {
        for (int s = 0; s < T5_MAX_NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < T5_MAX_NLSTAGE; ++s) { mbar_init(lfull + s, 1); mbar_init(lempty + s, 8); }
        for (int c = 0; c < 2; ++c) { for (int k = 0; k < 2; ++k) mbar_init(a_ready + 2 * c + k, 128); mbar_init(d_ready + c, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) { tc5::tmem_alloc(tmem_base_slot, 512); tc5::tmem_relinquish(); }
    for (int i = tid; i < a.n_steps; i += blockDim.x) steps[i] = a.steps[i];
    for (int i = tid; i < a.n_cherry; i += blockDim.x) cherry_leaves[i] = a.cherry_leaves[i];
    for (int i = tid; i < 128; i += blockDim.x) s_pi[i] = a.pi[i >> 6][i & 63];
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_base_slot;

    const uint32_t n_unique = *a.n_unique;
    const uint32_t npairs = (n_unique + 255) / 256;

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 9) {
            // ---- TMA producer: inner-edge tiles
            if (lane == 0) {
                uint32_t use = 0;
                for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
                    for (int m = 0; m < 2; ++m)
                        for (int s = 0; s < a.n_steps; ++s, ++use) {
                            const uint32_t st = use % T5_NSTAGE;
                            mbar_wait(empty + st, ((use / T5_NSTAGE) & 1) ^ 1);
                            mbar_arrive_expect_tx(full + st, T5_TILE_BYTES);
                            tma_bulk_g2s(stage_buf + (size_t)st * T5_TILE_BYTES, a.pstream[m] + (size_t)s * 8192, T5_TILE_BYTES, full + st);
                        }
            }
        } else if (warp == 10) {
            // ---- TMA producer: leaf gather tables in program order
            if (lane == 0) {
                uint32_t use = 0;
                for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
                    for (int m = 0; m < 2; ++m)
                        for (int k = 0; k < a.n_leaf_tabs; ++k, ++use) {
                            const uint32_t st = use % T5_NLSTAGE;
                            mbar_wait(lempty + st, ((use / T5_NLSTAGE) & 1) ^ 1);
                            mbar_arrive_expect_tx(lfull + st, T5_LEAF_BYTES);
                            tma_bulk_g2s(leaf_buf + (size_t)st * T5_LEAF_BYTES, a.leaftab[m] + (size_t)k * T5_LEAF_FLOATS, T5_LEAF_BYTES, lfull + st);
                        }
            }
        } else if (warp == 8) {
            // ---- MMA issue: warp-uniform control flow, one elected lane issues
            const uint32_t idesc = tc5::idesc_tf32(128, 128), idesc_hi = tc5::idesc_tf32(128, 64);
            uint32_t use = 0;
            for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
                for (int m = 0; m < 2; ++m)
                    for (int s = 0; s < a.n_steps; ++s, ++use) {
                        const uint32_t st = use % T5_NSTAGE;
#ifdef PCSF_TC5_TRACE
                        const bool first_seq = pair == blockIdx.x + 4 * gridDim.x && m == 0 && lane == 0;
#endif
                        mbar_wait(full + st, (use / T5_NSTAGE) & 1);
                        const uint32_t sb = tc5::smem_addr(stage_buf + (size_t)st * T5_TILE_BYTES);
                        for (int c = 0; c < 2; ++c) {
                            // k-steps 0..3 only need states 0..31 of A: they are issued while the epilogue threads still split and store
                            // the other half (measured: 101.2 -> 95.8 ms per 8 Mi columns; four quarters: 98.6 ms, the extra
                            // tcgen05.wait::st round trips cost more than the earlier start gains)
                            const uint32_t ta = tmem + c * 256 + (use & 1) * 128, td = tmem + c * 256 + ((use & 1) ^ 1) * 128;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                mbar_wait(a_ready + 2 * c + k, use & 1);
                                if (k == 0) T5_TRACE(2, s, 2 * c);
                                tc5::fence_after_sync();
                                if (tc5::elect_one()) {
#pragma unroll
                                    for (int j = 4 * k; j < 4 * k + 4; ++j) {
                                        const uint64_t bd = tc5::smem_desc(sb + j * 4096, 128, 256);
                                        tc5::mma_tf32_ts(td, ta + 8 * j, bd, idesc, j > 0);
                                        tc5::mma_tf32_ts(td, ta + 64 + 8 * j, bd, idesc_hi, 1);      // lo(A) x hi(B) only
                                    }
                                    if (k == 1) {
                                        tc5::commit(d_ready + c);
                                        if (c == 1) tc5::commit(empty + st);
                                    }
                                }
                                __syncwarp();
                            }
                            __syncwarp();
                            T5_TRACE(2, s, 2 * c + 1);
                        }
                    }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ---- epilogue warps: chain c, thread = window t = TMEM lane t
        const int c = warp >> 2, t = tid & 127;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c * 256;
        uint8_t *myids = ids + (size_t)c * a.ws.nl * 128 + t;                      // [leaf * 128]
        float *stk = a.scratch + ((size_t)blockIdx.x * 2 + c) * (size_t)(a.max_stack > 0 ? a.max_stack : 1) * T5_STACK_ENTRY_FLOATS;
        uint32_t use = 0, luse = 0;

        // L (= or *=) message of the next leaf in program order, gathered from its shared-memory table
        auto gather = [&](float (&L)[64], int leaf, bool mul) {
            const uint32_t st = luse % T5_NLSTAGE;
            mbar_wait(lfull + st, (luse / T5_NLSTAGE) & 1);
            const uint32_t x = myids[leaf * 128];
            if (x != 64u) {
                const float4 *row = reinterpret_cast<const float4 *>(leaf_buf + (size_t)st * T5_LEAF_BYTES) + x * (T5_LEAF_ROW / 4);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 v = row[j];
                    if (mul) { L[4 * j] *= v.x; L[4 * j + 1] *= v.y; L[4 * j + 2] *= v.z; L[4 * j + 3] *= v.w; }
                    else { L[4 * j] = v.x; L[4 * j + 1] = v.y; L[4 * j + 2] = v.z; L[4 * j + 3] = v.w; }
                }
            } else if (!mul) {
#pragma unroll
                for (int i = 0; i < 64; ++i) L[i] = 1.0f;
            }
            // the table is overwritten by TMA (async proxy) once all 8 warps have arrived: order the generic-proxy loads
            // above before it — they may still sit in the LSU queue behind bank-conflicted wavefronts
            tc5::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(lempty + st);
            ++luse;
        };

        // Cherry rows.  The warp owns a 32-row x 256-byte staging area; chunk ch (16 bytes) of row rr sits at chunk position
        // (ch & 8) | ((ch ^ rr) & 7): the sixteen lanes that copy one row write two full 128-byte lines, and the eight threads of a
        // quarter warp that later read chunk j of their OWN rows (LDS.128) hit eight different bank groups.
        unsigned char *cstage = cherry_buf + (size_t)warp * T5_CHERRY_STAGE_BYTES;
        const uint32_t cstage_s = tc5::smem_addr(cstage);
        int ck = 0;                        // cherry (program order) whose rows are in flight / staged
        auto prefetch_cherry = [&](int m, int k) {
            const uint32_t cl = cherry_leaves[k];
            const uint32_t myrow = (uint32_t)myids[(cl & 0xffu) * 128] * 65u + (uint32_t)myids[(cl >> 8) * 128];
            const float *tab = a.cherrytab[m] + (size_t)k * ((size_t)T5_CHERRY_ROWS * 64);
            const int half = lane >> 4, ch = lane & 15;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int rr = 2 * i + half;                      // the window (row of the staging area) this lane copies a chunk of
                const uint32_t row = __shfl_sync(0xffffffffu, myrow, rr);
                cp_async_16(cstage_s + rr * 256 + (((ch & 8) | ((ch ^ rr) & 7)) << 4), tab + (size_t)row * 64 + ch * 4);
            }
            cp_async_commit();
        };
        // L (= or *=) the staged cherry message of this thread's window; then starts the copy of the next cherry of the program
        // (same ECM, or the other ECM's first one; the next pair's first one is started once its leaf ids are known)
        auto take_cherry = [&](float (&L)[64], bool mul, int m) {
            cp_async_wait_all();
            __syncwarp();
            const float4 *row = reinterpret_cast<const float4 *>(cstage + lane * 256);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 v = row[(j & 8) | ((j ^ lane) & 7)];
                if (mul) { L[4 * j] *= v.x; L[4 * j + 1] *= v.y; L[4 * j + 2] *= v.z; L[4 * j + 3] *= v.w; }
                else { L[4 * j] = v.x; L[4 * j + 1] = v.y; L[4 * j + 2] = v.z; L[4 * j + 3] = v.w; }
            }
            __syncwarp();
            // the copy of the program's next cherry starts right away: it needs a whole step of lead time (issuing it only after
            // this step's hand-over was measured 40 % slower at 8 Mi columns: the rows then come from DRAM as often as from L2)
            if (++ck < a.n_cherry) prefetch_cherry(m, ck);
            else { ck = 0; if (m == 0) prefetch_cherry(1, 0); }
        };
        auto fetch = [&](float (&L)[64], uint32_t src, bool mul, int m) {
            if (src & T5_SRC_CHERRY) take_cherry(L, mul, m);
            else gather(L, (int)src, mul);
        };

#ifdef PCSF_TC5_TRACE
        long long t_ids = 0, t_seq = 0, t_pro = 0, t_begin = clock64();
#endif
        for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
            // leaf codon ids of this thread's window
            const uint32_t u = pair * 256 + c * 128 + t;
#ifdef PCSF_TC5_TRACE
            long long tt0 = clock64();
#endif
            {
                const uint32_t lw = a.uniq[u < n_unique ? u : n_unique - 1];
                int64_t o; uint32_t strand;
                if (a.ws.mode == 0) { o = a.ws.c0 + (lw >> 1); strand = lw & 1; }
                else { o = a.ws.win_off[lw]; strand = 0; }
                const uint8_t *p = a.ws.codes + o;
                // 32 species at a time: all loads first (the byte stores below may alias anything as far as the compiler
                // knows, which would otherwise serialise one DRAM round trip per species)
                for (int s0 = 0; s0 < a.ws.nl; s0 += 32) {
                    uint32_t v[32][3];
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const uint8_t *q = p + (int64_t)(s0 + k < a.ws.nl ? s0 + k : s0) * a.ws.ld;
                        v[k][0] = __ldg(q); v[k][1] = __ldg(q + 1); v[k][2] = __ldg(q + 2);
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (s0 + k < a.ws.nl)
                            myids[(s0 + k) * 128] = (uint8_t)(strand ? codon_minus(v[k][0], v[k][1], v[k][2]) : codon_plus(v[k][0], v[k][1], v[k][2]));
                }
            }
            if (a.n_cherry > 0) prefetch_cherry(0, 0);
#ifdef PCSF_TC5_TRACE
            t_ids += clock64() - tt0;
#endif
            for (int m = 0; m < 2; ++m) {
                float R[64];
                int E = 0, sp = 0;
#ifdef PCSF_TC5_TRACE
                long long tt1 = clock64();
#endif
                fetch(R, a.start & 0xffu, false, m);
                fetch(R, (a.start >> 8) & 0xffu, true, m);
#ifdef PCSF_TC5_TRACE
                long long tt2 = clock64();
                t_pro += tt2 - tt1;
#endif
                // A = split(alpha): per-window normalisation by an exact power of two, then TF32 hi + lo (hi = the 19 bits
                // the tensor core reads, lo = the exact remainder); hands the step to the MMA warp
                float amax_prev = 0.f;          // largest entry of the A that was handed over last (in [1, 2), or 0)
                auto split_and_arrive = [&](const float (&V)[64], uint32_t areg) {
                    float mx = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; ++i) mx = fmaxf(mx, V[i]);
                    const int e = mx > 0.f ? (int)((__float_as_uint(mx) >> 23) & 0xff) - 127 : 0;
                    const float sc = __uint_as_float((uint32_t)(127 - e) << 23);
                    const float2 sc2 = make_float2(sc, sc);
                    E += e;
                    amax_prev = mx * sc;
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const float2 r = __fmul2_rn(make_float2(V[16 * h + i], V[16 * h + i + 1]), sc2);
                            hi[i] = __float_as_uint(r.x) & 0xffffe000u;
                            hi[i + 1] = __float_as_uint(r.y) & 0xffffe000u;
                            const float2 l = __fadd2_rn(r, make_float2(-__uint_as_float(hi[i]), -__uint_as_float(hi[i + 1])));
                            lo[i] = __float_as_uint(l.x);
                            lo[i + 1] = __float_as_uint(l.y);
                        }
                        if (h == 2) {
                            // states 0..31 are on their way: hand them to the MMA warp before storing the rest
                            tc5::wait_st();
                            tc5::fence_before_sync();
                            mbar_arrive(a_ready + 2 * c);
                        }
                        tc5::st16(areg + 16 * h, hi);
                        tc5::st16(areg + 64 + 16 * h, lo);
                    }
                    tc5::wait_st();
                    tc5::fence_before_sync();
                    mbar_arrive(a_ready + 2 * c + 1);
                };
                if (a.n_steps > 0) split_and_arrive(R, lane_base + (use & 1) * 128);

                for (int s = 0; s < a.n_steps; ++s, ++use) {
                    const uint32_t step = steps[s];
                    const uint32_t post = (step >> 16) & 3u;
#ifdef PCSF_TC5_TRACE
                    const bool first_seq = pair == blockIdx.x + 4 * gridDim.x && m == 0 && (t == 0);
#endif
                    T5_TRACE(c, s, 0);
                    // ---- while the GEMM runs: what the program multiplies in before the next GEMM — leaf messages from
                    // their shared-memory tables, cherry messages from the staged table rows, or the waiting sibling partial
                    // from the stack
                    float L[64];
                    int Epop = 0;
                    if (post == T5_MUL) {
                        fetch(L, step & 0xffu, false, m);
                    } else if (post == T5_PUSH_START) {
                        fetch(L, step & 0xffu, false, m);
                        fetch(L, (step >> 8) & 0xffu, true, m);
                    } else if (post == T5_POP_MUL) {
                        --sp;
                        const float4 *e4 = reinterpret_cast<const float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float4 v = __ldcg(e4 + j * 128 + t);
                            L[4 * j] = v.x; L[4 * j + 1] = v.y; L[4 * j + 2] = v.z; L[4 * j + 3] = v.w;
                        }
                        Epop = __ldcg(reinterpret_cast<const int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t);
                    }
                    T5_TRACE(c, s, 1);
                    mbar_wait(d_ready + c, use & 1);
                    T5_TRACE(c, s, 2);
                    tc5::fence_after_sync();
                    const uint32_t dreg = lane_base + ((use & 1) ^ 1) * 128;     // D_s; A_{s+1} overwrites it in place
                    if (post == T5_PUSH_START || s + 1 == a.n_steps) {
                        // the message itself is needed (pushed onto the stack / dotted with pi): load all of it
                        {
                            uint32_t x0[32], y0[32], x1[32], y1[32];
                            tc5::ld32(dreg, x0);
                            tc5::ld32(dreg + 64, y0);
                            tc5::ld32(dreg + 32, x1);
                            tc5::ld32(dreg + 96, y1);
                            tc5::wait_ld();
                            // msg = D[0:64] + D[64:128]
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                const float2 v0 = __fadd2_rn(make_float2(__uint_as_float(x0[i]), __uint_as_float(x0[i + 1])),
                                                             make_float2(__uint_as_float(y0[i]), __uint_as_float(y0[i + 1])));
                                const float2 v1 = __fadd2_rn(make_float2(__uint_as_float(x1[i]), __uint_as_float(x1[i + 1])),
                                                             make_float2(__uint_as_float(y1[i]), __uint_as_float(y1[i + 1])));
                                R[i] = v0.x; R[i + 1] = v0.y; R[32 + i] = v1.x; R[32 + i + 1] = v1.y;
                            }
                        }
                        if (post == T5_PUSH_START) {
                            // the next GEMM's input is the new chain's start (already in L): hand it over first, then push the message
                            const int Epush = E;
                            E = 0;
                            split_and_arrive(L, dreg);
                                        float4 *e4 = reinterpret_cast<float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                            for (int j = 0; j < 16; ++j) __stcg(e4 + j * 128 + t, make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]));
                            __stcg(reinterpret_cast<int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t, Epush);
                            ++sp;
                        } else if (post == T5_MUL || post == T5_POP_MUL) {
#pragma unroll
                            for (int i = 0; i < 64; i += 2) {
                                const float2 v = __fmul2_rn(make_float2(R[i], R[i + 1]), make_float2(L[i], L[i + 1]));
                                R[i] = v.x; R[i + 1] = v.y;
                            }
                            E += Epop;
                        }
                    } else {
                        // Streamed hand-over.  Any power of two is an exact scale, so the normalisation does not have to wait
                        // for this message's own maximum: every entry is <= the largest entry of the A it was computed from
                        // (rows of P sum to <= 1, leaf and sibling factors are <= the scale they carry), hence the exponent of
                        // that previous maximum keeps the new A below 2, and how far below is corrected one step later.  Without
                        // a separate max pass the scale is folded into the combine (streaming the TMEM loads in two halves was measured slower:
                        // a second tcgen05.wait::ld round trip costs more than the overlap gains, 114 against 97 ms).
                        const bool mul = post == T5_MUL || post == T5_POP_MUL;
                        const int e = amax_prev > 0.f ? (int)((__float_as_uint(amax_prev) >> 23) & 0xff) - 127 : 0;
                        const float sc = __uint_as_float((uint32_t)(127 - e) << 23);
                        const float2 sc2 = make_float2(sc, sc);
                        E += e + (mul ? Epop : 0);
                        float amax = 0.f;
                        {
                            uint32_t x0[32], y0[32], x1[32], y1[32];
                            tc5::ld32(dreg, x0);
                            tc5::ld32(dreg + 64, y0);
                            tc5::ld32(dreg + 32, x1);
                            tc5::ld32(dreg + 96, y1);
                            tc5::wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                float2 v0 = __fadd2_rn(make_float2(__uint_as_float(x0[i]), __uint_as_float(x0[i + 1])),
                                                       make_float2(__uint_as_float(y0[i]), __uint_as_float(y0[i + 1])));
                                float2 v1 = __fadd2_rn(make_float2(__uint_as_float(x1[i]), __uint_as_float(x1[i + 1])),
                                                       make_float2(__uint_as_float(y1[i]), __uint_as_float(y1[i + 1])));
                                if (mul) {
                                    v0 = __fmul2_rn(v0, make_float2(L[i], L[i + 1]));
                                    v1 = __fmul2_rn(v1, make_float2(L[32 + i], L[32 + i + 1]));
                                }
                                v0 = __fmul2_rn(v0, sc2);
                                v1 = __fmul2_rn(v1, sc2);
                                amax = fmaxf(amax, fmaxf(fmaxf(v0.x, v0.y), fmaxf(v1.x, v1.y)));
                                R[i] = v0.x; R[i + 1] = v0.y; R[32 + i] = v1.x; R[32 + i + 1] = v1.y;
                            }
                        }
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                hi[i] = __float_as_uint(R[16 * h + i]) & 0xffffe000u;
                                lo[i] = __float_as_uint(R[16 * h + i] - __uint_as_float(hi[i]));
                            }
                            if (h == 2) {
                                tc5::wait_st();
                                tc5::fence_before_sync();
                                mbar_arrive(a_ready + 2 * c);
                            }
                            tc5::st16(dreg + 16 * h, hi);
                            tc5::st16(dreg + 64 + 16 * h, lo);
                        }
                        tc5::wait_st();
                        tc5::fence_before_sync();
                        mbar_arrive(a_ready + 2 * c + 1);
                        amax_prev = amax;
                            }
                    T5_TRACE(c, s, 3);
                }
                // z = pi . alpha_root (fixed_lik.hpp:159-163), log z with the exponents taken out so far
                {
                    const double *pi = s_pi + m * 64;
                    double z = 0.0;
#pragma unroll
                    for (int i = 0; i < 64; ++i) z += pi[i] * (double)R[i];
                    if (u < n_unique) a.logz[m][u] = log(z) + (double)E * 0.6931471805599453;
                }
#ifdef PCSF_TC5_TRACE
                t_seq += clock64() - tt2;
#endif
            }
        }
#ifdef PCSF_TC5_TRACE
        if ((blockIdx.x == 0 || blockIdx.x == 77) && t == 0)
            printf("T5 cta %d chain %d: total %lld cycles, ids %lld, prologue gathers %lld, steps+END %lld, pairs %u\n", blockIdx.x, c,
                   clock64() - t_begin, t_ids, t_pro, t_seq, (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x);
#endif
    }
#ifdef PCSF_TC5_TRACE
    __syncwarp();
    if (blockIdx.x == 0 && tid == 0) {
        const long long t0 = trace[0][0][0];
        for (int s = 0; s < a.n_steps && s < 128; ++s)
            printf("T5 s=%3d post=%u | X: A-ready %7lld gather %5lld wait-D %5lld combine %5lld | Y: A-ready %7lld gather %5lld wait-D %5lld combine %5lld | MMA: aX %7lld issX %4lld aY %7lld issY %4lld\n",
                   s, (steps[s] >> 16) & 3u, trace[0][s][0] - t0, trace[0][s][1] - trace[0][s][0], trace[0][s][2] - trace[0][s][1],
                   trace[0][s][3] - trace[0][s][2], trace[1][s][0] - t0, trace[1][s][1] - trace[1][s][0], trace[1][s][2] - trace[1][s][1],
                   trace[1][s][3] - trace[1][s][2], trace[2][s][0] - t0, trace[2][s][1] - trace[2][s][0], trace[2][s][2] - t0,
                   trace[2][s][3] - trace[2][s][2]);
    }
#endif
    // teardown: every tcgen05.ld has completed and every MMA was waited for by its epilogue
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 8) tc5::tmem_dealloc(tmem, 512);
}

}  // namespace pcsf
