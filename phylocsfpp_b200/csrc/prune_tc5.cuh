// prune_tc5.cuh — k_prune_tc5: Felsenstein pruning on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulators and the A operand in TMEM, P tiles streamed by bulk TMA).  FP32-class arithmetic with per-window
// log-scaling; the FP64 DMMA kernel (k_prune) stays the parity anchor.
//
// Formulation (ensure_alpha, fixed_lik.hpp:125-164).  For 128 codon windows at a time every edge c -> parent of the
// tree is ONE GEMM  D[w][a] = sum_b A[w][b] * P_c[a][b]  on the tensor core (M = 128 windows, K = 64 child states):
//   * inner edge: A = alpha_c, split into TF32 hi + lo (per-window power-of-two normalised, exponent kept as an
//     integer); B = [hi(P_c) | lo(P_c)] side by side (N = 128), so 2 MMAs per k-step give all four products
//     hi*hi, hi*lo, lo*hi, lo*lo in FP32 accumulators and msg = D[:, 0:64] + D[:, 64:128] (~2^-21 per product;
//     N = 128 is also the smallest N that runs at the full MMA rate: an MMA instruction has a ~64-cycle floor,
//     tools/tc5_probe.cu);
//   * leaf edge: A = one-hot row of the leaf's codon (all ones for a gap/N codon, id 64: the row sums of
//     fixed_lik.hpp:111-118), exact in TF32 -> 1 MMA per k-step; the gather of P_l[:, x] runs on the tensor pipe
//     instead of 32 KB of scattered L1/L2 reads per leaf and tile.
// The Hadamard products, push/pop of waiting sibling partials, normalisation and the hi/lo split run in the epilogue
// warps: thread = window = TMEM lane, the 64-state partial lives in registers between steps.
//
// One persistent CTA per SM works on a PAIR of 128-window tiles (chains X and Y) that share the TMA ring: while the
// tensor core runs chain Y's GEMM of step s, chain X's epilogue turns D_s into A_{s+1} (and vice versa).
//   warps 0-3   epilogue of chain X (TMEM lanes 32*(warp%4)..)      TMEM columns   0..255: two 128-column regions,
//   warps 4-7   epilogue of chain Y                                  TMEM columns 256..511  A_s in one, D_s in the other;
//   warp  8     MMA issue (one elected lane), TMEM alloc/dealloc      A_{s+1} overwrites D_s in place
//   warp  9     TMA producer: one 32 KB tile per step through a 3-stage full/empty mbarrier ring
// Waiting sibling partials (stack depth = Strahler number - 1) spill to an L2-resident scratch in global memory;
// every thread only ever touches its own column of it.
#pragma once

#include "kernels.cuh"
#include "tc5.cuh"

namespace pcsf {

constexpr int T5_NSTAGE = 3;
constexpr int T5_TILE_BYTES = 32768;
constexpr int T5_THREADS = 320;
constexpr int T5_STACK_ENTRY_FLOATS = 64 * 128 + 128;   // 128 windows x 64 states + 128 exponents

struct PruneTc5Args {
    WinSpace ws;
    const uint32_t *uniq;
    const uint32_t *n_unique;
    const uint32_t *steps;
    int n_steps, max_stack;
    const float *pstream[2];     // [n_steps][8192]
    const double *pi[2];
    double *logz[2];
    float *scratch;              // [grid][2][max_stack][T5_STACK_ENTRY_FLOATS]
};

__host__ __device__ inline size_t prune_tc5_smem_bytes(int nl, int n_steps) {
    size_t b = (size_t)T5_NSTAGE * T5_TILE_BYTES;
    b += (size_t)2 * nl * 128;                       // leaf codon ids of both tiles
    b += (size_t)((n_steps * 4 + 15) / 16) * 16;     // steps
    b += 2 * 64 * 8;                                 // pi
    b += 16 * 8;                                     // mbarriers + TMEM base
    return b;
}

__global__ void __launch_bounds__(T5_THREADS, 1) k_prune_tc5(const PruneTc5Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sp_ = smem;
    unsigned char *stage_buf = sp_; sp_ += (size_t)T5_NSTAGE * T5_TILE_BYTES;
    uint8_t *ids = sp_; sp_ += (size_t)2 * a.ws.nl * 128;
    uint32_t *steps = reinterpret_cast<uint32_t *>(sp_); sp_ += (size_t)((a.n_steps * 4 + 15) / 16) * 16;
    double *s_pi = reinterpret_cast<double *>(sp_); sp_ += 2 * 64 * 8;
    uint64_t *full = reinterpret_cast<uint64_t *>(sp_);
    uint64_t *empty = full + T5_NSTAGE;
    uint64_t *a_ready = empty + T5_NSTAGE;     // [2] epilogue -> MMA: A of the next step is in TMEM
    uint64_t *d_ready = a_ready + 2;           // [2] MMA -> epilogue: D of the step is complete
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(d_ready + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < T5_NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
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

    if (warp == 9) {
        // ---- TMA producer
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
        return;
    }

    if (warp == 8) {
        // ---- MMA issue: warp-uniform control flow, one elected lane issues
        const uint32_t idesc = tc5::idesc_tf32(128, 128);
        uint32_t use = 0;
        for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
            for (int m = 0; m < 2; ++m)
                for (int s = 0; s < a.n_steps; ++s, ++use) {
                    const uint32_t st = use % T5_NSTAGE;
                    const bool inner = ((steps[s] >> 16) & 3u) == T5_INNER;
                    mbar_wait(full + st, (use / T5_NSTAGE) & 1);
                    const uint32_t sb = tc5::smem_addr(stage_buf + (size_t)st * T5_TILE_BYTES);
                    for (int c = 0; c < 2; ++c) {
                        mbar_wait(a_ready + c, use & 1);
                        tc5::fence_after_sync();
                        if (tc5::elect_one()) {
                            const uint32_t ta = tmem + c * 256 + (use & 1) * 128, td = tmem + c * 256 + ((use & 1) ^ 1) * 128;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const uint64_t bd = tc5::smem_desc(sb + j * 4096, 128, 256);
                                tc5::mma_tf32_ts(td, ta + 8 * j, bd, idesc, j > 0);
                                if (inner) tc5::mma_tf32_ts(td, ta + 64 + 8 * j, bd, idesc, 1);
                            }
                            tc5::commit(d_ready + c);
                            if (c == 1) tc5::commit(empty + st);
                        }
                        __syncwarp();
                    }
                }
    } else {
        // ---- epilogue warps: chain c, thread = window t = TMEM lane t
        const int c = warp >> 2, t = tid & 127;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c * 256;
        uint8_t *myids = ids + (size_t)c * a.ws.nl * 128 + t;                      // [leaf * 128]
        float *stk = a.scratch + ((size_t)blockIdx.x * 2 + c) * (size_t)(a.max_stack > 0 ? a.max_stack : 1) * T5_STACK_ENTRY_FLOATS;
        uint32_t use = 0;

        // one-hot row of a leaf codon (all ones for id 64) into the A region
        auto write_onehot = [&](uint32_t region, int leaf) {
            const uint32_t x = myids[leaf * 128];
            uint32_t v[32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (x == 64u || x == (uint32_t)(32 * h + i)) ? 0x3f800000u : 0u;
                tc5::st32(region + 32 * h, v);
            }
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
#pragma unroll
                for (int i = 0; i < 64; ++i) R[i] = 0.f;
                int E = 0, sp = 0;
                // A of step 0 (always a leaf step)
                write_onehot(lane_base + (use & 1) * 128, (int)(steps[0] & 0xffffu));
                tc5::wait_st();
                tc5::fence_before_sync();
                mbar_arrive(a_ready + c);

                for (int s = 0; s < a.n_steps; ++s, ++use) {
                    const uint32_t step = steps[s];
                    const uint32_t kind = (step >> 16) & 3u;
                    mbar_wait(d_ready + c, use & 1);
                    tc5::fence_after_sync();
                    const uint32_t dreg = lane_base + ((use & 1) ^ 1) * 128;
                    // msg = D[0:64] + D[64:128]; a gap/N leaf codon (id 64) contributes the row sums of P_l, which are 1
                    // to FP32 precision (fixed_lik.hpp:111-118; PhyloModel_make normalises the rows, instance.hpp:625-639)
                    const bool miss = kind != T5_INNER && myids[(step & 0xffffu) * 128] == 64;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t x[32], y[32];
                        tc5::ld32(dreg + 32 * h, x);
                        tc5::ld32(dreg + 64 + 32 * h, y);
                        tc5::wait_ld();
                        if (kind == T5_LEAF_MUL) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) R[32 * h + i] *= miss ? 1.0f : __uint_as_float(x[i]) + __uint_as_float(y[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) R[32 * h + i] = miss ? 1.0f : __uint_as_float(x[i]) + __uint_as_float(y[i]);
                        }
                    }
                    if (kind == T5_LEAF_SET) E = 0;
                    if (step & T5_PUSH) {
                        float4 *e4 = reinterpret_cast<float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                        for (int j = 0; j < 16; ++j) __stcg(e4 + j * 128 + t, make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]));
                        __stcg(reinterpret_cast<int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t, E);
                        ++sp;
                    }
                    if (step & T5_POP_MUL) {
                        --sp;
                        const float4 *e4 = reinterpret_cast<const float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float4 v = __ldcg(e4 + j * 128 + t);
                            R[4 * j] *= v.x; R[4 * j + 1] *= v.y; R[4 * j + 2] *= v.z; R[4 * j + 3] *= v.w;
                        }
                        E += __ldcg(reinterpret_cast<const int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t);
                    }
                    if (step & T5_END) {
                        // z = pi . alpha_root (fixed_lik.hpp:159-163), log z with the exponents taken out so far
                        const double *pi = s_pi + m * 64;
                        double z = 0.0;
#pragma unroll
                        for (int i = 0; i < 64; ++i) z += pi[i] * (double)R[i];
                        if (u < n_unique) a.logz[m][u] = log(z) + (double)E * 0.6931471805599453;
                    }
                    if (s + 1 < a.n_steps) {
                        const uint32_t nstep = steps[s + 1];
                        const uint32_t areg = lane_base + ((use + 1) & 1) * 128;      // == dreg: D_s is dead now
                        if (((nstep >> 16) & 3u) == T5_INNER) {
                            // per-window normalisation by an exact power of two, then TF32 hi/lo split
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
                        } else {
                            write_onehot(areg, (int)(nstep & 0xffffu));
                        }
                        tc5::wait_st();
                        tc5::fence_before_sync();
                        mbar_arrive(a_ready + c);
                    }
                }
            }
        }
    }
    // teardown: every tcgen05.ld has completed and every MMA was waited for by its epilogue
    tc5::fence_before_sync();
    named_bar_sync(1, 9 * 32);
    if (warp == 8) tc5::tmem_dealloc(tmem, 512);
}

}  // namespace pcsf
