// tc5_probe.cu — tcgen05 / TMEM micro-measurements on sm_100a that size k_prune_tc5:
//   1. correctness of the kind::tf32 TS MMA with the K-major no-swizzle B layout used by the kernel
//   2. MMA issue rate (M=128, N=64/128/256, K=8, A in TMEM)
//   3. tcgen05.ld / tcgen05.st throughput (32x32b), 4 / 8 / 16 warps
//   4. ld+st running next to a saturating MMA stream
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I phylocsfpp_b200/csrc -o tools/tc5_probe tools/tc5_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc5.cuh"

using namespace pcsf::tc5;

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W1:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D1;\n"
        "bra W1;\n"
        "D1:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// B tile layout (floats): chunk j (K = 8j..8j+7) is 2 KB contiguous; inside: (n/8)*64 + (k%8/4)*32 + (n%8)*4 + k%4
__host__ __device__ inline int b_index(int n, int k) { return (k / 8) * 512 + (n / 8) * 64 + ((k % 8) / 4) * 32 + (n % 8) * 4 + (k % 4); }

// ---- 1. correctness: D[128][64] = A[128][64] * B[64][64]^T-ish (B given as Bm[n][k]) ----------------------
__global__ void __launch_bounds__(128) k_check(const float *A, const float *Bt /* tile layout */, float *D) {
    __shared__ __align__(128) float sB[64 * 64];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tmem_alloc(&tbase, 256); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 4096; i += 128) sB[i] = Bt[i];
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t t0 = tbase;
    const uint32_t lane_addr = t0 + ((uint32_t)(warp * 32) << 16);
    // A -> TMEM columns [64, 128): thread = window/lane
    uint32_t r[32];
    for (int h = 0; h < 2; ++h) {
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(A[tid * 64 + h * 32 + c]);
        st32(lane_addr + 64 + h * 32, r);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, 64);
        for (int j = 0; j < 8; ++j)
            mma_tf32_ts(t0, t0 + 64 + 8 * j, smem_desc(smem_addr(sB) + j * 2048, 128, 256), idesc, j > 0);
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int h = 0; h < 2; ++h) {
        ld32(lane_addr + h * 32, r);
        wait_ld();
        for (int c = 0; c < 32; ++c) D[tid * 64 + h * 32 + c] = __uint_as_float(r[c]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t0, 256);
}

// ---- 2/4. MMA rate, optionally with ld/st traffic from the other warps -------------------------------------
// warp 0 lane 0 issues `reps` groups of 24 MMAs (TS, N columns) and commits after each group; warps 1.. do
// `ldst` rounds of (ld 64 cols, st 128 cols) on their lane quarter in parallel.
__global__ void __launch_bounds__(544) k_rate(int N, int reps, int ss_mode, int ldst_rounds, long long *out) {
    extern __shared__ __align__(1024) unsigned char dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<float *>(dyn)[i] = 0.001f * (i % 97);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t t0 = tbase;
    long long c0 = clock64();
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(128, N);
            const uint32_t sb = smem_addr(dyn);
            for (int rep = 0; rep < reps; ++rep) {
                for (int j = 0; j < 8; ++j)
                    for (int s = 0; s < 3; ++s) {
                        if (ss_mode)
                            mma_tf32_ss(t0, smem_desc(sb + 32768 + j * 4096, 128, 256), smem_desc(sb + (j & 3) * 2048 * (N / 64), 128, 256), idesc, 1);
                        else
                            mma_tf32_ts(t0, t0 + 256 + 8 * j + 64 * (s & 1), smem_desc(sb + (j & 3) * 2048 * (N / 64), 128, 256), idesc, 1);
                    }
                commit(&bar);
                mbar_wait(&bar, rep & 1);
            }
        }
        __syncwarp();
    } else if (ldst_rounds > 0) {
        const uint32_t lane_addr = t0 + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t colbase = 384 + ((warp - 1) / 4 & 1) * 64;   // scratch columns away from D/A
        uint32_t r[32], q[32];
        for (int i = 0; i < 32; ++i) { r[i] = i; q[i] = 2 * i; }
        for (int it = 0; it < ldst_rounds; ++it) {
            ld32(lane_addr + colbase, r);
            ld32(lane_addr + colbase + 32, q);
            wait_ld();
            for (int i = 0; i < 32; ++i) { r[i] += 1; q[i] ^= r[i]; }
            st32(lane_addr + colbase, r);
            st32(lane_addr + colbase + 32, q);
            st32(lane_addr + colbase, q);
            st32(lane_addr + colbase + 32, r);
            wait_st();
        }
        if (r[3] == 0x12345678u) out[100] = q[5];
    }
    long long c1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = c1 - c0;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t0, 512);
}

// ---- 3. ld / st throughput --------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_ldst(int mode /*0 ld, 1 st*/, int iters, long long *out) {
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t t0 = tbase;
    const uint32_t lane_addr = t0 + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t r[32], q[32];
    for (int i = 0; i < 32; ++i) { r[i] = i + tid; q[i] = i * tid; }
    st32(lane_addr, r); st32(lane_addr + 32, q); wait_st();
    __syncthreads();
    long long c0 = clock64();
    if (mode == 0) {
        uint32_t acc = 0;
        for (int it = 0; it < iters; ++it) {
            ld32(lane_addr, r);
            ld32(lane_addr + 32, q);
            wait_ld();
            acc += r[0] ^ q[31] ^ r[17];
        }
        if (acc == 0x12345678u) out[200] = acc;
    } else {
        for (int it = 0; it < iters; ++it) {
            r[0] += it;
            st32(lane_addr, r);
            st32(lane_addr + 32, q);
            wait_st();
        }
    }
    long long c1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = c1 - c0;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t0, 512);
}


// ---- 5. MMA rate with `nacc` independent accumulators round-robin, one commit at the end ------------------------
__global__ void __launch_bounds__(32) k_rate_acc(int N, int nacc, int total, long long *out) {
    extern __shared__ __align__(1024) unsigned char dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x;
    tmem_alloc(&tbase, 512); tmem_relinquish();
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<float *>(dyn)[i] = 0.001f * (i % 97);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t t0 = tbase;
    long long c0 = clock64();
    {
        const uint32_t idesc = idesc_tf32(128, N);
        const uint32_t sb = smem_addr(dyn);
        int acc = 0;
        for (int i = 0; i < total; ++i) {
            if (elect_one()) mma_tf32_ts(t0 + acc * N, t0 + 448 + 8 * (i & 7), smem_desc(sb + (i & 7) * 2048, 128, 256), idesc, 1);
            if (++acc == nacc) acc = 0;
        }
        if (elect_one()) commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
    }
    long long c1 = clock64();
    if (tid == 0) out[blockIdx.x] = c1 - c0;
    __syncthreads();
    tmem_dealloc(t0, 512);
}

// ---- 6. tcgen05.ld 16x256b.x16 (64 regs: 16 lanes x 128 columns) ---------------------------------------------
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(512) k_ld_shape(int iters, long long *out) {
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t t0 = tbase;
    const uint32_t lane_addr = t0 + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t r[32], q[32];
    long long c0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        ld_16x256b_x8(lane_addr, r);                        // lanes 0-15 of the quarter, 64 columns
        ld_16x256b_x8(lane_addr + (16u << 16), q);          // lanes 16-31
        wait_ld();
        acc += r[0] ^ q[31] ^ r[17];
    }
    if (acc == 0x12345678u) out[200] = acc;
    long long c1 = clock64();
    if (lane == 0) out[blockIdx.x * 32 + warp] = c1 - c0;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t0, 512);
}

// ---- 7. shared-memory leaf gather: 128 threads, each reads the 256-byte row of a random codon (row stride 272 B) -----
__global__ void __launch_bounds__(128) k_lds_gather(int iters, int stride_bytes, int spread, long long *out, float *sink) {
    extern __shared__ __align__(1024) unsigned char dyn[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 65 * 80; i += 128) reinterpret_cast<float *>(dyn)[i] = 1.0f + i * 1e-6f;
    __syncthreads();
    uint32_t x = (tid * 2654435761u) >> 7;
    float acc[64];
    for (int i = 0; i < 64; ++i) acc[i] = 1.f;
    long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        const int row = spread ? (x >> 10) % 65 : 64;
        const float4 *p = reinterpret_cast<const float4 *>(dyn + row * stride_bytes);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 v = p[j];
            acc[4 * j] *= v.x; acc[4 * j + 1] *= v.y; acc[4 * j + 2] *= v.z; acc[4 * j + 3] *= v.w;
        }
    }
    long long c1 = clock64();
    float s = 0;
    for (int i = 0; i < 64; ++i) s += acc[i];
    sink[tid] = s;
    if (tid == 0) out[0] = c1 - c0;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
    // 1. correctness
    std::vector<float> A(128 * 64), Bm(64 * 64), Bt(4096), D(128 * 64);
    srand(1);
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; };
    for (auto &v : A) v = tf((float)rand() / RAND_MAX);
    for (auto &v : Bm) v = tf((float)rand() / RAND_MAX - 0.3f);
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) Bt[b_index(n, k)] = Bm[n * 64 + k];
    float *dA, *dB, *dD; long long *dout;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bt.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMalloc(&dout, 8192 * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
    k_check<<<1, 128>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)A[m * 64 + k] * Bm[n * 64 + k];
            maxerr = fmax(maxerr, fabs(ref - D[m * 64 + n]));
        }
    printf("{\"check_max_abs_err\": %.3e}\n", maxerr);

    std::vector<long long> h(8192);
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    // 2. MMA rates
    for (int ss = 0; ss < 2; ++ss)
        for (int N : {64, 128, 256}) {
            for (int grid : {1, 148}) {
                const int reps = 200;
                k_rate<<<grid, 32, 128 * 1024>>>(N, reps, ss, 0, dout);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h.data(), dout, 8192 * 8, cudaMemcpyDeviceToHost));
                printf("{\"mma\": \"%s\", \"N\": %d, \"grid\": %d, \"cycles_per_mma\": %.2f}\n", ss ? "SS" : "TS", N, grid,
                       (double)h[0] / (reps * 24.0));
            }
        }
    // 3. ld / st
    for (int mode = 0; mode < 2; ++mode)
        for (int nw : {4, 8, 16}) {
            const int iters = 2000;
            k_ldst<<<1, nw * 32>>>(mode, iters, dout);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), dout, 8192 * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            printf("{\"tmem\": \"%s\", \"warps\": %d, \"bytes_per_cycle_per_sm\": %.1f, \"cycles_per_64col_warp_op\": %.1f}\n",
                   mode ? "st" : "ld", nw, (double)nw * 32 * 64 * 4 * iters / mx, (double)mx / iters);
        }
    // 4. MMA + ld/st
    for (int nw : {4, 8}) {
        const int reps = 200;
        k_rate<<<148, 32 * (1 + nw), 128 * 1024>>>(64, reps, 0, reps, dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), dout, 8192 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int w = 1; w <= nw; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("{\"mixed_epi_warps\": %d, \"mma_cycles_per_mma\": %.2f, \"epi_cycles_per_round(ld64+st128)\": %.1f}\n", nw,
               (double)h[0] / (reps * 24.0), (double)mx / reps);
    }
    // 5. independent accumulators
    CK(cudaFuncSetAttribute(k_rate_acc, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int N : {64, 128})
        for (int nacc : {1, 2, 3, 4}) {
            if (nacc * N > 448) continue;
            const int total = 4800;
            k_rate_acc<<<1, 32, 64 * 1024>>>(N, nacc, total, dout);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), dout, 8192 * 8, cudaMemcpyDeviceToHost));
            printf("{\"mma_nacc\": %d, \"N\": %d, \"cycles_per_mma\": %.2f}\n", nacc, N, (double)h[0] / total);
        }
    // 6. ld 16x256b
    for (int nw : {4, 8}) {
        const int iters = 2000;
        k_ld_shape<<<1, nw * 32>>>(iters, dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), dout, 8192 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("{\"tmem\": \"ld16x256b\", \"warps\": %d, \"bytes_per_cycle_per_sm\": %.1f}\n", nw, (double)nw * 32 * 64 * 4 * iters / mx);
    }
    // 7. smem gather
    {
        float *sink; CK(cudaMalloc(&sink, 4096));
        CK(cudaFuncSetAttribute(k_lds_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        for (int stride : {256, 272})
            for (int spread : {0, 1}) {
                const int iters = 2000;
                k_lds_gather<<<1, 128, 64 * 1024>>>(iters, stride, spread, dout, sink);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h.data(), dout, 8, cudaMemcpyDeviceToHost));
                printf("{\"lds_gather_stride\": %d, \"random_rows\": %d, \"cycles_per_leaf_tile(128 rows)\": %.1f}\n", stride, spread, (double)h[0] / iters);
            }
    }
    return 0;
}
