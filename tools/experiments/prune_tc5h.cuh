// prune_tc5h.cuh — k_prune_tc5h: the tcgen05/TMEM pruning kernel of prune_tc5.cuh with the epilogue spread over SIXTEEN warps.
//
// Same formulation, same data (P tiles, leaf tables, step program, scratch) and same MMA / TMA warps as k_prune_tc5.  What
// changes is who turns D_g into A_{g+1}: per-step traces of k_prune_tc5 (profiles/tc5_step_trace_r1.txt) show a chain's period
// = leaf gathers (1.0-2.6 k cycles) + combine (1.3 k cycles) against 2 x 0.8 k cycles of GEMM — the tensor core waits for the
// 128 threads of a chain, each of which owns a whole 64-state partial.  Here TWO threads own a window: warps w and w + 8
// address the same TMEM lane quarter (w % 4), one holds states 0..31, the other 32..63.  Each thread gathers half a leaf row,
// loads / combines / splits / stores half a partial; the only cross-thread quantity, the per-window maximum for the power-of-two
// normalisation, is exchanged through a double-buffered shared-memory word and one 64-thread named barrier per step.
//   warps  0-15  epilogue: quarter q = w & 3, chain c = (w >> 2) & 1, half h = w >> 3
//   warp  16     MMA issue        warp 17  TMA: inner-edge tiles        warp 18  TMA: leaf tables
// 640 threads (warp 19 only completes the producer warpgroup): the epilogue warps run with 104 registers (setmaxnreg), the
// producer warpgroup with 40.
#pragma once

#include "prune_tc5.cuh"

namespace pcsf {

constexpr int T5H_THREADS = 640;          // 20 warps: setmaxnreg is a warpgroup-wide instruction, so the producer warpgroup must be complete
constexpr int T5H_XCH_BYTES = 2 * 2 * 2 * 128 * 4 + 2 * 2 * 128 * 8;          // xmax[parity][c][h][t] floats + zx[c][h][t] doubles

__host__ __device__ inline size_t prune_tc5h_smem_bytes(int nl, int n_steps, int nstage, int nlstage) {
    return prune_tc5_smem_bytes(nl, n_steps, nstage, nlstage) + T5H_XCH_BYTES;
}
inline void prune_tc5h_pick_stages(int nl, int n_steps, int *nstage, int *nlstage) {
    *nstage = T5_MAX_NSTAGE; *nlstage = T5_MAX_NLSTAGE;
    while (prune_tc5h_smem_bytes(nl, n_steps, *nstage, *nlstage) > 227 * 1024 && *nlstage > 3) --*nlstage;
    while (prune_tc5h_smem_bytes(nl, n_steps, *nstage, *nlstage) > 227 * 1024 && *nstage > 2) --*nstage;
}

__global__ void __launch_bounds__(T5H_THREADS, 1) k_prune_tc5h(const PruneTc5Args a) {
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
    uint64_t *a_ready = lempty + T5_MAX_NLSTAGE;
    uint64_t *d_ready = a_ready + 2;
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(d_ready + 2);
    sp_ += 32 * 8;
    double *zx = reinterpret_cast<double *>(sp_); sp_ += 2 * 2 * 128 * 8;          // [c][h][t]
    float *xmax = reinterpret_cast<float *>(sp_);                                  // [parity][c][h][t]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < T5_MAX_NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < T5_MAX_NLSTAGE; ++s) { mbar_init(lfull + s, 1); mbar_init(lempty + s, 16); }
        for (int c = 0; c < 2; ++c) { mbar_init(a_ready + c, 256); mbar_init(d_ready + c, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) { tc5::tmem_alloc(tmem_base_slot, 512); tc5::tmem_relinquish(); }
    for (int i = tid; i < a.n_steps; i += blockDim.x) steps[i] = a.steps[i];
    for (int i = tid; i < 128; i += blockDim.x) s_pi[i] = a.pi[i >> 6][i & 63];
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_base_slot;

    const uint32_t n_unique = *a.n_unique;
    const uint32_t npairs = (n_unique + 255) / 256;

    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 17) {
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
        } else if (warp == 18) {
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
        } else if (warp == 16) {
            // ---- MMA issue (as in k_prune_tc5)
            const uint32_t idesc = tc5::idesc_tf32(128, 128), idesc_hi = tc5::idesc_tf32(128, 64);
            uint32_t use = 0;
            for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
                for (int m = 0; m < 2; ++m)
                    for (int s = 0; s < a.n_steps; ++s, ++use) {
                        const uint32_t st = use % T5_NSTAGE;
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
                                    tc5::mma_tf32_ts(td, ta + 64 + 8 * j, bd, idesc_hi, 1);
                                }
                                tc5::commit(d_ready + c);
                                if (c == 1) tc5::commit(empty + st);
                            }
                            __syncwarp();
                        }
                    }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");          // 512 x 104 + 128 x 40 <= 640 x 96: the pool is the CTA's own allocation
        const int q = warp & 3, c = (warp >> 2) & 1, h = warp >> 3, t = q * 32 + lane;
        const int pair_bar = 1 + c * 4 + q;                                   // named barrier of the two warps that share windows
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + c * 256;
        uint8_t *myids = ids + (size_t)c * a.ws.nl * 128 + t;                      // [leaf * 128]
        float *stk = a.scratch + ((size_t)blockIdx.x * 2 + c) * (size_t)(a.max_stack > 0 ? a.max_stack : 1) * T5_STACK_ENTRY_FLOATS;
        float *my_xmax = xmax + (c * 2 + h) * 128 + t, *other_xmax = xmax + (c * 2 + (h ^ 1)) * 128 + t;
        uint32_t use = 0, luse = 0, nsplit = 0;

        // L (= or *=) this thread's half of the message of the next leaf in program order
        auto gather = [&](float (&L)[32], int leaf, bool mul) {
            const uint32_t st = luse % T5_NLSTAGE;
            mbar_wait(lfull + st, (luse / T5_NLSTAGE) & 1);
            const uint32_t x = myids[leaf * 128];
            if (x != 64u) {
                const float4 *row = reinterpret_cast<const float4 *>(leaf_buf + (size_t)st * T5_LEAF_BYTES) + x * (T5_LEAF_ROW / 4) + 8 * h;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 v = row[j];
                    if (mul) { L[4 * j] *= v.x; L[4 * j + 1] *= v.y; L[4 * j + 2] *= v.z; L[4 * j + 3] *= v.w; }
                    else { L[4 * j] = v.x; L[4 * j + 1] = v.y; L[4 * j + 2] = v.z; L[4 * j + 3] = v.w; }
                }
            } else if (!mul) {
#pragma unroll
                for (int i = 0; i < 32; ++i) L[i] = 1.0f;
            }
            tc5::fence_proxy_async_smem();          // generic-proxy reads before the async-proxy overwrite (see k_prune_tc5)
            __syncwarp();
            if (lane == 0) mbar_arrive(lempty + st);
            ++luse;
        };

        for (uint32_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
            const uint32_t u = pair * 256 + c * 128 + t;
            // both threads of a window are done with the previous pair's ids
            named_bar_sync(pair_bar, 64);
            {
                const uint32_t lw = a.uniq[u < n_unique ? u : n_unique - 1];
                int64_t o; uint32_t strand;
                if (a.ws.mode == 0) { o = a.ws.c0 + (lw >> 1); strand = lw & 1; }
                else { o = a.ws.win_off[lw]; strand = 0; }
                const uint8_t *p = a.ws.codes + o;
                // the two threads of a window split the species; 16 at a time, all loads before the byte stores
                const int half_n = (a.ws.nl + 1) >> 1, sbeg = h * half_n, send = h ? a.ws.nl : half_n;
                for (int s0 = sbeg; s0 < send; s0 += 16) {
                    uint32_t v[16][3];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const uint8_t *qq = p + (int64_t)(s0 + k < send ? s0 + k : s0) * a.ws.ld;
                        v[k][0] = __ldg(qq); v[k][1] = __ldg(qq + 1); v[k][2] = __ldg(qq + 2);
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (s0 + k < send)
                            myids[(s0 + k) * 128] = (uint8_t)(strand ? codon_minus(v[k][0], v[k][1], v[k][2]) : codon_plus(v[k][0], v[k][1], v[k][2]));
                }
            }
            named_bar_sync(pair_bar, 64);
            for (int m = 0; m < 2; ++m) {
                float R[32];
                int E = 0, sp = 0;
                gather(R, a.first0, false);
                gather(R, a.first1, true);
                // A = split(alpha): per-window power-of-two normalisation (maximum over both halves), TF32 hi + lo
                auto split_and_arrive = [&](const float (&V)[32], uint32_t areg) {
                    float mx = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, V[i]);
                    const uint32_t par = (nsplit & 1) * 512;
                    my_xmax[par] = mx;
                    named_bar_sync(pair_bar, 64);
                    mx = fmaxf(mx, other_xmax[par]);
                    ++nsplit;
                    const int e = mx > 0.f ? (int)((__float_as_uint(mx) >> 23) & 0xff) - 127 : 0;
                    const float sc = __uint_as_float((uint32_t)(127 - e) << 23);
                    const float2 sc2 = make_float2(sc, sc);
                    E += e;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; i += 2) {
                            const float2 r = __fmul2_rn(make_float2(V[8 * k + i], V[8 * k + i + 1]), sc2);
                            hi[i] = __float_as_uint(r.x) & 0xffffe000u;
                            hi[i + 1] = __float_as_uint(r.y) & 0xffffe000u;
                            const float2 l = __fadd2_rn(r, make_float2(-__uint_as_float(hi[i]), -__uint_as_float(hi[i + 1])));
                            lo[i] = __float_as_uint(l.x);
                            lo[i + 1] = __float_as_uint(l.y);
                        }
                        tc5::st8(areg + 32 * h + 8 * k, hi);
                        tc5::st8(areg + 64 + 32 * h + 8 * k, lo);
                    }
                    tc5::wait_st();
                    tc5::fence_before_sync();
                    mbar_arrive(a_ready + c);
                };
                if (a.n_steps > 0) split_and_arrive(R, lane_base + (use & 1) * 128);

                for (int s = 0; s < a.n_steps; ++s, ++use) {
                    const uint32_t step = steps[s];
                    const uint32_t post = (step >> 16) & 3u;
                    // ---- while the GEMM runs: make sure the leaf tables of this step have landed and find this window's rows
                    // (the multiply reads them straight from shared memory once D is there: holding a gathered copy in registers
                    // across the TMEM load does not fit into 104 registers); the popped sibling partial is prefetched from L2
                    const float4 *rowA = nullptr, *rowB = nullptr;
                    uint32_t stA = 0, stB = 0;
                    float L[32];
                    int Epop = 0;
                    if (post == T5_MUL_LEAF || post == T5_PUSH_CHERRY) {
                        stA = luse % T5_NLSTAGE;
                        mbar_wait(lfull + stA, (luse / T5_NLSTAGE) & 1);
                        ++luse;
                        const uint32_t x = myids[(step & 0xffu) * 128];
                        if (x != 64u) rowA = reinterpret_cast<const float4 *>(leaf_buf + (size_t)stA * T5_LEAF_BYTES) + x * (T5_LEAF_ROW / 4) + 8 * h;
                        if (post == T5_PUSH_CHERRY) {
                            stB = luse % T5_NLSTAGE;
                            mbar_wait(lfull + stB, (luse / T5_NLSTAGE) & 1);
                            ++luse;
                            const uint32_t y = myids[((step >> 8) & 0xffu) * 128];
                            if (y != 64u) rowB = reinterpret_cast<const float4 *>(leaf_buf + (size_t)stB * T5_LEAF_BYTES) + y * (T5_LEAF_ROW / 4) + 8 * h;
                        }
                    } else if (post == T5_POP_MUL) {
                        --sp;
                        const float4 *e4 = reinterpret_cast<const float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 v = __ldcg(e4 + (8 * h + j) * 128 + t);
                            L[4 * j] = v.x; L[4 * j + 1] = v.y; L[4 * j + 2] = v.z; L[4 * j + 3] = v.w;
                        }
                        Epop = __ldcg(reinterpret_cast<const int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t);
                    }
                    mbar_wait(d_ready + c, use & 1);
                    tc5::fence_after_sync();
                    const uint32_t dreg = lane_base + ((use & 1) ^ 1) * 128;     // D_s; A_{s+1} overwrites it in place
                    // msg = D[0:64] + D[64:128], this thread's 32 states
                    {
                        uint32_t x0[32];
                        tc5::ld16(dreg + 32 * h, x0);
                        tc5::ld16(dreg + 32 * h + 16, x0 + 16);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            uint32_t y0[16];
                            tc5::ld16(dreg + 64 + 32 * h + 16 * k, y0);
                            tc5::wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                const float2 v0 = __fadd2_rn(make_float2(__uint_as_float(x0[16 * k + i]), __uint_as_float(x0[16 * k + i + 1])),
                                                             make_float2(__uint_as_float(y0[i]), __uint_as_float(y0[i + 1])));
                                R[16 * k + i] = v0.x; R[16 * k + i + 1] = v0.y;
                            }
                        }
                    }
                    if (post == T5_PUSH_CHERRY) {
                        // the message waits on the stack (fire-and-forget stores); the next GEMM's input is the cherry
                        float4 *e4 = reinterpret_cast<float4 *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS);
#pragma unroll
                        for (int j = 0; j < 8; ++j) __stcg(e4 + (8 * h + j) * 128 + t, make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]));
                        if (h == 0) __stcg(reinterpret_cast<int *>(stk + (size_t)sp * T5_STACK_ENTRY_FLOATS + 8192) + t, E);
                        ++sp;
                        E = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 v = rowA ? rowA[j] : make_float4(1.f, 1.f, 1.f, 1.f);
                            if (rowB) { const float4 w = rowB[j]; v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w; }
                            R[4 * j] = v.x; R[4 * j + 1] = v.y; R[4 * j + 2] = v.z; R[4 * j + 3] = v.w;
                        }
                    } else if (post == T5_MUL_LEAF) {
                        if (rowA) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 v = rowA[j];
                                R[4 * j] *= v.x; R[4 * j + 1] *= v.y; R[4 * j + 2] *= v.z; R[4 * j + 3] *= v.w;
                            }
                        }
                    } else if (post == T5_POP_MUL) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float2 v = __fmul2_rn(make_float2(R[i], R[i + 1]), make_float2(L[i], L[i + 1]));
                            R[i] = v.x; R[i + 1] = v.y;
                        }
                        E += Epop;
                    }
                    if (post == T5_MUL_LEAF || post == T5_PUSH_CHERRY) {
                        // release the leaf tables: generic-proxy reads before the async-proxy overwrite (see k_prune_tc5)
                        tc5::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(lempty + stA); if (post == T5_PUSH_CHERRY) mbar_arrive(lempty + stB); }
                    }
                    if (post == T5_PUSH_CHERRY || s + 1 < a.n_steps) split_and_arrive(R, dreg);
                }
                // z = pi . alpha_root (fixed_lik.hpp:159-163): the upper half hands its partial sum to the lower half
                {
                    const double *pi = s_pi + m * 64 + 32 * h;
                    double z = 0.0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) z += pi[i] * (double)R[i];
                    if (h == 1) zx[(c * 2 + 1) * 128 + t] = z;
                    named_bar_sync(pair_bar, 64);
                    if (h == 0) {
                        z += zx[(c * 2 + 1) * 128 + t];
                        if (u < n_unique) a.logz[m][u] = log(z) + (double)E * 0.6931471805599453;
                    }
                    named_bar_sync(pair_bar, 64);          // zx is free again (trees without inner edges have no other barrier in between)
                }
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 16) tc5::tmem_dealloc(tmem, 512);
}

}  // namespace pcsf
