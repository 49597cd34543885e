// prune_tc5.cuh — k_prune_tc5: Felsenstein pruning on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators and the A operand in TMEM, P tiles and leaf tables streamed by bulk TMA).  FP32-class arithmetic with
// per-window log-scaling; the FP64 DMMA kernel (k_prune) stays the parity anchor.
//
// Formulation (ensure_alpha, fixed_lik.hpp:125-164).  For 128 codon windows at a time every INNER edge c -> parent is
// one GEMM  D[w][a] = sum_b A[w][b] * P_c[a][b]  on the tensor core (M = 128 windows, K = 64 child states):
// A = alpha_c split into TF32 hi + lo (per-window power-of-two normalised, the exponent kept as an integer);
// B = [hi(P_c) | lo(P_c)] side by side (N = 128): per k-step one N = 128 MMA with hi(A) gives hi*hi | hi*lo and one
// N = 64 MMA with lo(A) adds lo*hi, in FP32 accumulators; msg = D[:, 0:64] + D[:, 64:128] (~2^-21 per product).  One MMA
// instruction has a ~64-cycle floor whatever its N <= 128 (tools/tc5_probe.cu), so the side-by-side layout is what makes
// two instructions per k-step enough; shared-memory bandwidth (B reads + leaf gathers) is the co-bottleneck.
// LEAF edges are not GEMMs: the message of a leaf with codon x is column x of P_l (all ones for a gap/N codon, the row
// sums of fixed_lik.hpp:111-118).  The epilogue threads gather it from a 17 KB per-leaf table that TMA streams into shared
// memory in program order — and they do so WHILE their GEMM runs, so leaves cost no tensor time and no latency.
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
// Waiting sibling partials (stack depth = Strahler number - 1) spill to an L2-resident scratch in global memory, every
// thread only ever touching its own column: the push is issued after the next A has been handed to the tensor core, the
// pop is prefetched while the GEMM runs, so neither is on the critical path.
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
    int first0, first1;          // the cherry the program starts with
    int nstage, nlstage;         // ring depths
    const float *pstream[2];     // [n_steps][8192]
    const float *leaftab[2];     // [nl][T5_LEAF_FLOATS], program order
    const double *pi[2];
    double *logz[2];
    float *scratch;              // [grid][2][max_stack][T5_STACK_ENTRY_FLOATS]
};

__host__ __device__ inline size_t prune_tc5_smem_bytes(int nl, int n_steps, int nstage, int nlstage) {
    size_t b = (size_t)nstage * T5_TILE_BYTES + (size_t)nlstage * T5_LEAF_BYTES;
    b += (size_t)2 * nl * 128;                              // leaf codon ids of both tiles
    b += (size_t)(((n_steps + 1) * 4 + 15) / 16) * 16;      // steps
    b += 2 * 64 * 8;                                        // pi
    b += 32 * 8;                                            // mbarriers + TMEM base (2*3 + 2*6 + 4 + 2 barriers)
    return b;
}
// Deepest rings that fit into one SM's shared memory.
inline void prune_tc5_pick_stages(int nl, int n_steps, int *nstage, int *nlstage) {
    *nstage = T5_MAX_NSTAGE; *nlstage = T5_MAX_NLSTAGE;
    while (prune_tc5_smem_bytes(nl, n_steps, *nstage, *nlstage) > 227 * 1024 && *nlstage > 3) --*nlstage;
    while (prune_tc5_smem_bytes(nl, n_steps, *nstage, *nlstage) > 227 * 1024 && *nstage > 2) --*nstage;
}

__global__ void __launch_bounds__(T5_THREADS, 1) k_prune_tc5(const PruneTc5Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sp_ = smem;
    const uint32_t T5_NSTAGE = a.nstage, T5_NLSTAGE = a.nlstage;
    unsigned char *stage_buf = sp_; sp_ += (size_t)T5_NSTAGE * T5_TILE_BYTES;
    unsigned char *leaf_buf = sp_; sp_ += (size_t)T5_NLSTAGE * T5_LEAF_BYTES;
    uint8_t *ids = sp_; sp_ += (size_t)2 * a.ws.nl * 128;
    uint32_t *steps = reinterpret_cast<uint32_t *>(sp_); sp_ += (size_t)(((a.n_steps + 1) * 4 + 15) / 16) * 16;
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
                        for (int k = 0; k < a.ws.nl; ++k, ++use) {
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
#ifdef PCSF_TC5_TRACE
            t_ids += clock64() - tt0;
#endif
            for (int m = 0; m < 2; ++m) {
                float R[64];
                int E = 0, sp = 0;
#ifdef PCSF_TC5_TRACE
                long long tt1 = clock64();
#endif
                gather(R, a.first0, false);
                gather(R, a.first1, true);
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
                    // their shared-memory tables, or the waiting sibling partial from the stack
                    float L[64];
                    int Epop = 0;
                    if (post == T5_MUL_LEAF) {
                        gather(L, (int)(step & 0xffu), false);
                    } else if (post == T5_PUSH_CHERRY) {
                        gather(L, (int)(step & 0xffu), false);
                        gather(L, (int)((step >> 8) & 0xffu), true);
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
                    if (post == T5_PUSH_CHERRY || s + 1 == a.n_steps) {
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
                        if (post == T5_PUSH_CHERRY) {
                            // the next GEMM's input is the cherry (already in L): hand it over first, then push the message
                            const int Epush = E;
                            E = 0;
                            split_and_arrive(L, dreg);
                            float4 *e4 = reinterpret_cast<float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                            for (int j = 0; j < 16; ++j) __stcg(e4 + j * 128 + t, make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]));
                            __stcg(reinterpret_cast<int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t, Epush);
                            ++sp;
                        } else if (post == T5_MUL_LEAF || post == T5_POP_MUL) {
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
                        const bool mul = post == T5_MUL_LEAF || post == T5_POP_MUL;
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
