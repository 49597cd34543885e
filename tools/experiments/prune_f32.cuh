// prune_f32.cuh — FP32-class tensor-core variant of k_prune (north star: "tries FP64 DMMA against FP32 with
// per-column log-scaling").
//
// Same program, same register-chained GEMM structure and the same TMA-streamed tile ring as k_prune, but
//   * partials are FP32, 16 codon windows per warp (mma.sync.m16n8k8 TF32, SASS HMMA.1688.F32.TF32);
//   * every product is done in split TF32: x = hi + lo (cvt.rna.tf32), D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi
//     with FP32 accumulation — ~2^-21 relative error per product instead of TF32's 2^-11 (plain TF32 misses the
//     1e-3 deciban contract);
//   * per-window log-scaling: after every GEMM the 64-vector of a window is renormalised by an exact power of
//     two (max -> [1,2)) and the exponent is accumulated as an integer, so FP32's range never underflows
//     (the reference multiplies raw FP64 likelihoods, fixed_lik.hpp:155; 58 leaves reach ~1e-100);
//   * log z = log(sum_a pi[a] R[a]) + E ln 2 is finished in FP64.
// The K permutation that makes the accumulator fragment of one GEMM the A fragment of the next:
//   k-index q <-> child state 8*ks + 2q,  k-index q+4 <-> child state 8*ks + 2q + 1
//   => a0 = c0, a1 = c2, a2 = c1, a3 = c3 of n-tile ks.
#pragma once

#include "kernels.cuh"

namespace pcsf {

constexpr int PF_MAX_NWARP = 8;    // more warps do not help this kernel and would cap registers below its need
constexpr int PF_THREADS = (PF_MAX_NWARP + 1) * 32;   // launch bound; the launch uses (nwarp + 1) * 32
constexpr int PF_NSTAGE = 3;
constexpr int PF_TILE_BYTES = 2 * NS * NS * 4;   // 32 KB: hi and lo of one 64x64 P
constexpr int PF_STACK_ENTRY = 4096 + 256;       // 16 windows x 64 floats + per-thread exponent pair

struct PruneF32Args {
    WinSpace ws;
    const uint32_t *uniq;
    const uint32_t *n_unique;
    const int32_t *program;
    int n_ops, n_gemm, max_stack;
    int nwarp;                   // compute warps per CTA (16 windows each)
    const float *pstream[2];     // [n_gemm][2048] float4 {hi0,hi1,lo0,lo1}
    const float *leafPT[2];      // [nl][65][64]
    const double *pi[2];
    double *logz[2];
    uint32_t stagger_ns;
};

__host__ __device__ inline size_t prune_f32_smem_bytes(int nl, int n_ops, int max_stack, int nwarp) {
    size_t b = (size_t)PF_NSTAGE * PF_TILE_BYTES;
    b += (size_t)nwarp * (max_stack > 0 ? max_stack : 1) * PF_STACK_ENTRY;
    b += (size_t)((nl * nwarp * 16 + 15) / 16) * 16;
    b += (size_t)nwarp * 16 * 8;
    b += (size_t)((n_ops * 4 + 15) / 16) * 16;
    b += 2 * 64 * 8;
    b += 2 * PF_NSTAGE * 8;
    return b;
}

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void hmma_tf32(float *c, const uint32_t *a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(PF_THREADS, 1) k_prune_f32(const PruneF32Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sp_ = smem;
    const int PF_NWARP = a.nwarp, PF_TILE_W = a.nwarp * 16;
    float *stage_buf = reinterpret_cast<float *>(sp_); sp_ += (size_t)PF_NSTAGE * PF_TILE_BYTES;
    unsigned char *stack = sp_; sp_ += (size_t)PF_NWARP * (a.max_stack > 0 ? a.max_stack : 1) * PF_STACK_ENTRY;
    uint8_t *ids = sp_; sp_ += (size_t)((a.ws.nl * PF_TILE_W + 15) / 16) * 16;
    int64_t *s_woff = reinterpret_cast<int64_t *>(sp_); sp_ += (size_t)PF_TILE_W * 8;
    int32_t *prog = reinterpret_cast<int32_t *>(sp_); sp_ += (size_t)((a.n_ops * 4 + 15) / 16) * 16;
    double *s_pi = reinterpret_cast<double *>(sp_); sp_ += 2 * 64 * 8;
    uint64_t *full = reinterpret_cast<uint64_t *>(sp_);
    uint64_t *empty = full + PF_NSTAGE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < PF_NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, PF_NWARP); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < a.n_ops; i += blockDim.x) prog[i] = a.program[i];
    for (int i = tid; i < 128; i += blockDim.x) s_pi[i] = a.pi[i >> 6][i & 63];
    __syncthreads();

    const uint32_t n_unique = *a.n_unique;
    const uint32_t ntiles = (n_unique + PF_TILE_W - 1) / PF_TILE_W;

    if (warp == PF_NWARP) {
        if (lane == 0) {
            uint32_t use = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int m = 0; m < 2; ++m)
                    for (int g = 0; g < a.n_gemm; ++g, ++use) {
                        const uint32_t st = use % PF_NSTAGE;
                        mbar_wait(empty + st, ((use / PF_NSTAGE) & 1) ^ 1);
                        mbar_arrive_expect_tx(full + st, PF_TILE_BYTES);
                        tma_bulk_g2s(stage_buf + (size_t)st * 2 * NS * NS, a.pstream[m] + (size_t)g * 2 * NS * NS,
                                     PF_TILE_BYTES, full + st);
                    }
        }
        return;
    }

    const int g = lane >> 2, q = lane & 3;
    const int win_lo = warp * 16 + g, win_hi = win_lo + 8;     // the two windows (rows g, g+8) of this thread
    unsigned char *mystack = stack + (size_t)warp * (a.max_stack > 0 ? a.max_stack : 1) * PF_STACK_ENTRY;
    uint32_t use = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        named_bar_sync(1, PF_NWARP * 32);
        if (tid < PF_TILE_W) {
            uint32_t u = tile * PF_TILE_W + tid;
            if (u >= n_unique) u = n_unique - 1;
            const uint32_t lw = a.uniq[u];
            int64_t o; uint32_t strand;
            if (a.ws.mode == 0) { o = a.ws.c0 + (lw >> 1); strand = lw & 1; }
            else { o = a.ws.win_off[lw]; strand = 0; }
            s_woff[tid] = (o << 1) | strand;
        }
        named_bar_sync(1, PF_NWARP * 32);
#pragma unroll 4
        for (int i = tid; i < a.ws.nl * PF_TILE_W; i += PF_NWARP * 32) {
            const int s = i / PF_TILE_W, wi = i - s * PF_TILE_W;
            const int64_t ow = s_woff[wi];
            const uint8_t *p = a.ws.codes + (int64_t)s * a.ws.ld + (ow >> 1);
            const uint32_t x0 = p[0], x1 = p[1], x2 = p[2];
            ids[i] = (uint8_t)((ow & 1) ? codon_minus(x0, x1, x2) : codon_plus(x0, x1, x2));
        }
        named_bar_sync(1, PF_NWARP * 32);
        if (warp >= PF_NWARP / 2 && a.stagger_ns) __nanosleep(a.stagger_ns);   // phase offset, see k_prune

        for (int m = 0; m < 2; ++m) {
            const float *leafPT = a.leafPT[m];
            // v[nt] = {row g: states 8nt+2q, +1 ; row g+8: states 8nt+2q, +1}
            auto load_leaf = [&](int leaf, float4(&v)[8]) {
                const int x0 = ids[leaf * PF_TILE_W + win_lo], x1 = ids[leaf * PF_TILE_W + win_hi];
                const float2 *r0 = reinterpret_cast<const float2 *>(leafPT + ((size_t)leaf * 65 + x0) * NS);
                const float2 *r1 = reinterpret_cast<const float2 *>(leafPT + ((size_t)leaf * 65 + x1) * NS);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float2 a0 = __ldg(r0 + nt * 4 + q), a1 = __ldg(r1 + nt * 4 + q);
                    v[nt] = make_float4(a0.x, a0.y, a1.x, a1.y);
                }
            };
            // R[nt][c]: c0 = (row g, state 8nt+2q), c1 = (g, 8nt+2q+1), c2 = (g+8, 8nt+2q), c3 = (g+8, 8nt+2q+1)
            float R[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) R[nt][c] = 0.f;
            int E0 = 0, E1 = 0;    // power-of-two exponents taken out of rows g and g+8
            int sp = 0;
            for (int pc = 0; pc < a.n_ops; ++pc) {
                const int32_t op = prog[pc];
                const int code = op >> 16, arg = op & 0xffff;
                if (code == OP_GATHER_SET || code == OP_GATHER_MUL) {
                    float4 v[8];
                    load_leaf(arg, v);
                    if (code == OP_GATHER_SET) {
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) { R[nt][0] = v[nt].x; R[nt][1] = v[nt].y; R[nt][2] = v[nt].z; R[nt][3] = v[nt].w; }
                        E0 = 0; E1 = 0;
                    } else {
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) { R[nt][0] *= v[nt].x; R[nt][1] *= v[nt].y; R[nt][2] *= v[nt].z; R[nt][3] *= v[nt].w; }
                    }
                } else if (code == OP_GEMM) {
                    const uint32_t st = use % PF_NSTAGE;
                    mbar_wait(full + st, (use / PF_NSTAGE) & 1);
                    const float4 *bt = reinterpret_cast<const float4 *>(stage_buf + (size_t)st * 2 * NS * NS) + lane;
                    float acc[8][4];
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[nt][c] = 0.f;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const float x[4] = {R[ks][0], R[ks][2], R[ks][1], R[ks][3]};
                        uint32_t ah[4], al[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // cvt.rna.tf32 done with integer ALU ops (round half away on the 13 dropped bits); the
                            // residual is exact in FP32 and the tensor core ignores its low 13 bits
                            ah[i] = (__float_as_uint(x[i]) + 0x1000u) & 0xffffe000u;
                            al[i] = __float_as_uint(x[i] - __uint_as_float(ah[i]));
                        }
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) {
                            const float4 b = bt[(ks * 8 + nt) * 32];
                            const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y);
                            hmma_tf32(acc[nt], al, bh0, bh1);
                            hmma_tf32(acc[nt], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                            hmma_tf32(acc[nt], ah, bh0, bh1);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + st);
                    ++use;
                    // per-window renormalisation by an exact power of two
                    float m0 = 0.f, m1 = 0.f;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        m0 = fmaxf(m0, fmaxf(acc[nt][0], acc[nt][1]));
                        m1 = fmaxf(m1, fmaxf(acc[nt][2], acc[nt][3]));
                    }
                    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
                    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
                    const int e0 = m0 > 0.f ? (int)((__float_as_uint(m0) >> 23) & 0xff) - 127 : 0;
                    const int e1 = m1 > 0.f ? (int)((__float_as_uint(m1) >> 23) & 0xff) - 127 : 0;
                    const float s0 = __uint_as_float((uint32_t)(127 - e0) << 23), s1 = __uint_as_float((uint32_t)(127 - e1) << 23);
                    E0 += e0; E1 += e1;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        R[nt][0] = acc[nt][0] * s0; R[nt][1] = acc[nt][1] * s0;
                        R[nt][2] = acc[nt][2] * s1; R[nt][3] = acc[nt][3] * s1;
                    }
                } else if (code == OP_PUSH) {
                    float4 *st4 = reinterpret_cast<float4 *>(mystack + (size_t)sp * PF_STACK_ENTRY);
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) st4[nt * 32 + lane] = make_float4(R[nt][0], R[nt][1], R[nt][2], R[nt][3]);
                    reinterpret_cast<int2 *>(mystack + (size_t)sp * PF_STACK_ENTRY + 4096)[lane] = make_int2(E0, E1);
                    ++sp;
                } else if (code == OP_POP_MUL) {
                    --sp;
                    const float4 *st4 = reinterpret_cast<const float4 *>(mystack + (size_t)sp * PF_STACK_ENTRY);
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const float4 v = st4[nt * 32 + lane];
                        R[nt][0] *= v.x; R[nt][1] *= v.y; R[nt][2] *= v.z; R[nt][3] *= v.w;
                    }
                    const int2 e = reinterpret_cast<const int2 *>(mystack + (size_t)sp * PF_STACK_ENTRY + 4096)[lane];
                    E0 += e.x; E1 += e.y;
                } else {  // OP_END
                    const double *pi = s_pi + m * 64;
                    double z0 = 0.0, z1 = 0.0;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        z0 += pi[8 * nt + 2 * q] * (double)R[nt][0] + pi[8 * nt + 2 * q + 1] * (double)R[nt][1];
                        z1 += pi[8 * nt + 2 * q] * (double)R[nt][2] + pi[8 * nt + 2 * q + 1] * (double)R[nt][3];
                    }
                    z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
                    z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
                    const uint32_t u0 = tile * PF_TILE_W + win_lo, u1 = tile * PF_TILE_W + win_hi;
                    if (q == 0) {
                        if (u0 < n_unique) a.logz[m][u0] = log(z0) + (double)E0 * 0.6931471805599453;
                        if (u1 < n_unique) a.logz[m][u1] = log(z1) + (double)E1 * 0.6931471805599453;
                    }
                }
            }
        }
    }
}

}  // namespace pcsf
