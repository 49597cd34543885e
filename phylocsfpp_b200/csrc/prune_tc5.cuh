// prune_tc5.cuh — k_prune_tc5: Felsenstein pruning on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators and the A operand in TMEM, P tiles and leaf tables streamed by bulk TMA).  FP32-class arithmetic with
// per-window log-scaling; the FP64 DMMA kernel (k_prune) stays the parity anchor.
//
// Formulation (ensure_alpha, fixed_lik.hpp:125-164).  For 128 codon windows at a time every INNER edge c -> parent is
// one GEMM  D[w][a] = sum_b A[w][b] * P_c[a][b]  on the tensor core (M = 128 windows, K = 64 child states):
// A = alpha_c split into TF32 hi + lo (per-window power-of-two normalised, the exponent kept as an integer);
// B = [hi(P_c) | lo(P_c)] side by side (N = 128), so 2 MMAs per k-step give all four products hi*hi, hi*lo, lo*hi,
// lo*lo in FP32 accumulators and msg = D[:, 0:64] + D[:, 64:128] (~2^-21 per product).  N = 128 is also the smallest N
// that runs at the full MMA rate: one MMA instruction has a ~64-cycle floor (tools/tc5_probe.cu).
// LEAF edges are not GEMMs: the message of a leaf with codon x is column x of P_l (all ones for a gap/N codon, the row
// sums of fixed_lik.hpp:111-118).  The epilogue threads gather it from a 17 KB per-leaf table that TMA streams into shared
// memory in program order — and they do so WHILE their GEMM runs, so leaves cost no tensor time and no latency.
//
// One persistent CTA per SM works on a PAIR of 128-window tiles (chains X and Y) that share both TMA rings: while the
// tensor core runs chain Y's GEMM of step g, chain X's epilogue turns D_g into A_{g+1} (and vice versa).
//   warps 0-3   epilogue of chain X: thread = window = TMEM lane; the 64-state partial lives in registers
//   warps 4-7   epilogue of chain Y        TMEM columns 256c..256c+255: two 128-column regions, A_g in one, D_g in the
//   warp  8     MMA issue (one elected lane)    other; A_{g+1} overwrites D_g in place
//   warp  9     TMA producer, inner-edge tiles (32 KB each, 3-stage full/empty mbarrier ring)
//   warp 10     TMA producer, leaf tables (17 KB each, 4-stage ring)
// The producer warpgroup gives its registers to the epilogue warpgroups (setmaxnreg 40 / 232).
// Waiting sibling partials (stack depth = Strahler number - 1) spill to an L2-resident scratch in global memory; every
// thread only ever touches its own column of it.
#pragma once

#include "kernels.cuh"
#include "tc5.cuh"

namespace pcsf {

constexpr int T5_NSTAGE = 3;             // inner-edge tile ring
constexpr int T5_TILE_BYTES = 32768;
constexpr int T5_NLSTAGE = 4;            // leaf table ring
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
    const float *pstream[2];     // [n_steps][8192]
    const float *leaftab[2];     // [nl][T5_LEAF_FLOATS], program order
    const double *pi[2];
    double *logz[2];
    float *scratch;              // [grid][2][max_stack][T5_STACK_ENTRY_FLOATS]
};

__host__ __device__ inline size_t prune_tc5_smem_bytes(int nl, int n_steps) {
    size_t b = (size_t)T5_NSTAGE * T5_TILE_BYTES + (size_t)T5_NLSTAGE * T5_LEAF_BYTES;
    b += (size_t)2 * nl * 128;                              // leaf codon ids of both tiles
    b += (size_t)(((n_steps + 1) * 4 + 15) / 16) * 16;      // steps
    b += 2 * 64 * 8;                                        // pi
    b += 32 * 8;                                            // mbarriers + TMEM base
    return b;
}

__global__ void __launch_bounds__(T5_THREADS, 1) k_prune_tc5(const PruneTc5Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sp_ = smem;
    unsigned char *stage_buf = sp_; sp_ += (size_t)T5_NSTAGE * T5_TILE_BYTES;
    unsigned char *leaf_buf = sp_; sp_ += (size_t)T5_NLSTAGE * T5_LEAF_BYTES;
    uint8_t *ids = sp_; sp_ += (size_t)2 * a.ws.nl * 128;
    uint32_t *steps = reinterpret_cast<uint32_t *>(sp_); sp_ += (size_t)(((a.n_steps + 1) * 4 + 15) / 16) * 16;
    double *s_pi = reinterpret_cast<double *>(sp_); sp_ += 2 * 64 * 8;
    uint64_t *full = reinterpret_cast<uint64_t *>(sp_);
    uint64_t *empty = full + T5_NSTAGE;
    uint64_t *lfull = empty + T5_NSTAGE;
    uint64_t *lempty = lfull + T5_NLSTAGE;
    uint64_t *a_ready = lempty + T5_NLSTAGE;   // [2] epilogue -> MMA: A of the step is in TMEM
    uint64_t *d_ready = a_ready + 2;           // [2] MMA -> epilogue: D of the step is complete
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(d_ready + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef PCSF_TC5_TRACE
    __shared__ long long trace[3][128][4];
#define T5_TRACE(who, s, i) do { if (blockIdx.x == 0 && first_seq && (s) < 128) trace[who][s][i] = clock64(); } while (0)
#else
#define T5_TRACE(who, s, i) do { } while (0)
#endif
    if (tid == 0) {
        for (int s = 0; s < T5_NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < T5_NLSTAGE; ++s) { mbar_init(lfull + s, 1); mbar_init(lempty + s, 8); }
        for (int c = 0; c < 2; ++c) { mbar_init(a_ready + c, 128); mbar_init(d_ready + c, 1); }
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
            const uint32_t idesc = tc5::idesc_tf32(128, 128);
            uint32_t use = 0;
            for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
                for (int m = 0; m < 2; ++m)
                    for (int s = 0; s < a.n_steps; ++s, ++use) {
                        const uint32_t st = use % T5_NSTAGE;
#ifdef PCSF_TC5_TRACE
                        const bool first_seq = pair == blockIdx.x && m == 0 && lane == 0;
#endif
                        mbar_wait(full + st, (use / T5_NSTAGE) & 1);
                        const uint32_t sb = tc5::smem_addr(stage_buf + (size_t)st * T5_TILE_BYTES);
                        for (int c = 0; c < 2; ++c) {
                            mbar_wait(a_ready + c, use & 1);
                            T5_TRACE(2, s, 2 * c);
                            tc5::fence_after_sync();
                            if (tc5::elect_one()) {
                                const uint32_t ta = tmem + c * 256 + (use & 1) * 128, td = tmem + c * 256 + ((use & 1) ^ 1) * 128;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint64_t bd = tc5::smem_desc(sb + j * 4096, 128, 256);
                                    tc5::mma_tf32_ts(td, ta + 8 * j, bd, idesc, j > 0);
                                    tc5::mma_tf32_ts(td, ta + 64 + 8 * j, bd, idesc, 1);
                                }
                                tc5::commit(d_ready + c);
                                if (c == 1) tc5::commit(empty + st);
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

        for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
            // leaf codon ids of this thread's window
            const uint32_t u = pair * 256 + c * 128 + t;
            {
                const uint32_t lw = a.uniq[u < n_unique ? u : n_unique - 1];
                int64_t o; uint32_t strand;
                if (a.ws.mode == 0) { o = a.ws.c0 + (lw >> 1); strand = lw & 1; }
                else { o = a.ws.win_off[lw]; strand = 0; }
                const uint8_t *p = a.ws.codes + o;
#pragma unroll 4
                for (int s = 0; s < a.ws.nl; ++s, p += a.ws.ld) {
                    const uint32_t x0 = p[0], x1 = p[1], x2 = p[2];
                    myids[s * 128] = (uint8_t)(strand ? codon_minus(x0, x1, x2) : codon_plus(x0, x1, x2));
                }
            }
            for (int m = 0; m < 2; ++m) {
                float R[64];
                int E = 0, sp = 0;
                gather(R, a.first0, false);
                gather(R, a.first1, true);

                for (int s = 0; s < a.n_steps; ++s, ++use) {
                    const uint32_t step = steps[s];
                    const uint32_t post = (step >> 16) & 3u;
#ifdef PCSF_TC5_TRACE
                    const bool first_seq = pair == blockIdx.x && m == 0 && (t == 0);
#endif
                    // ---- A_s = split(alpha): per-window normalisation by an exact power of two, then TF32 hi/lo
                    {
                        const uint32_t areg = lane_base + (use & 1) * 128;
                        float mx = 0.f;
#pragma unroll
                        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, R[i]);
                        const int e = mx > 0.f ? (int)((__float_as_uint(mx) >> 23) & 0xff) - 127 : 0;
                        const float sc = __uint_as_float((uint32_t)(127 - e) << 23);
                        E += e;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t hi[32], lo[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float r = R[32 * h + i] * sc;
                                hi[i] = (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
                                lo[i] = __float_as_uint(r - __uint_as_float(hi[i]));
                            }
                            tc5::st32(areg + 32 * h, hi);
                            tc5::st32(areg + 64 + 32 * h, lo);
                        }
                        tc5::wait_st();
                        tc5::fence_before_sync();
                        mbar_arrive(a_ready + c);
                    }
                    T5_TRACE(c, s, 0);
                    // ---- while the GEMM runs: the leaf messages the program multiplies in before the next GEMM
                    float L[64];
                    if (post == T5_MUL_LEAF) {
                        gather(L, (int)(step & 0xffu), false);
                    } else if (post == T5_PUSH_CHERRY) {
                        gather(L, (int)(step & 0xffu), false);
                        gather(L, (int)((step >> 8) & 0xffu), true);
                    }
                    T5_TRACE(c, s, 1);
                    mbar_wait(d_ready + c, use & 1);
                    T5_TRACE(c, s, 2);
                    tc5::fence_after_sync();
                    {
                        const uint32_t dreg = lane_base + ((use & 1) ^ 1) * 128;
                        uint32_t x0[32], y0[32], x1[32], y1[32];
                        tc5::ld32(dreg, x0);
                        tc5::ld32(dreg + 64, y0);
                        tc5::ld32(dreg + 32, x1);
                        tc5::ld32(dreg + 96, y1);
                        tc5::wait_ld();
                        // msg = D[0:64] + D[64:128]
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            R[i] = __uint_as_float(x0[i]) + __uint_as_float(y0[i]);
                            R[32 + i] = __uint_as_float(x1[i]) + __uint_as_float(y1[i]);
                        }
                    }
                    if (post == T5_MUL_LEAF) {
#pragma unroll
                        for (int i = 0; i < 64; ++i) R[i] *= L[i];
                    } else if (post == T5_PUSH_CHERRY) {
                        float4 *e4 = reinterpret_cast<float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                        for (int j = 0; j < 16; ++j) __stcg(e4 + j * 128 + t, make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]));
                        __stcg(reinterpret_cast<int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t, E);
                        ++sp;
#pragma unroll
                        for (int i = 0; i < 64; ++i) R[i] = L[i];
                        E = 0;
                    } else if (post == T5_POP_MUL) {
                        --sp;
                        const float4 *e4 = reinterpret_cast<const float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float4 v = __ldcg(e4 + j * 128 + t);
                            R[4 * j] *= v.x; R[4 * j + 1] *= v.y; R[4 * j + 2] *= v.z; R[4 * j + 3] *= v.w;
                        }
                        E += __ldcg(reinterpret_cast<const int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t);
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
            }
        }
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
