// row_fetch_probe.cu — how to bring per-window 256-byte table rows (cherry / leaf messages of k_prune_tc5) from L2 into a
// warp's shared-memory staging area: cost of the three candidate mechanisms on sm_100a, per warp and source (32 rows).
//   L  16-byte cp.async (LDGSTS): 16 per lane, sixteen lanes per row, XOR-swizzled staging           (LSU pipe)
//   B  one 256-byte 1-D bulk TMA copy per lane (cp.async.bulk -> UBLKCP, serialised per lane), padded rows (async proxy)
//   G  TMA gather4 through a tensor map (box 32 x 1 floats, SWIZZLE_128B): 16 instructions per 32 rows     (async proxy)
// Every variant is verified against the table, then timed: cycles to ISSUE and cycles until the rows have LANDED, with
// 1, 4 and 8 warps of the CTA fetching at the same time (148 CTAs, one per SM, random rows of a 40 MB table).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/row_fetch_probe tools/row_fetch_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WL:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WD;\n"
        "bra WL;\n"
        "WD:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int ROWS = 160000;      // 41 MB of 256-byte rows
constexpr int NSRC = 64;          // sources fetched per warp in the timed loop

struct Res { long long issue, landed; int bad; };

// mode 0 = L, 1 = B, 2 = G
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_probe(const float *tab, const __grid_constant__ CUtensorMap tmap, const uint32_t *rows, int nwarps,
                                                  Res *res, int verify) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= nwarps) return;
    unsigned char *stage = smem + warp * 9216;          // 8704 needed for the padded variant; 1024-aligned (9216 = 9 * 1024)
    const uint32_t stage_s = smem_u32(stage);
    uint64_t *bar = bars + warp;
    long long t_issue = 0, t_land = 0;
    int bad = 0;
    uint32_t phase = 0;
    for (int src = 0; src < NSRC; ++src) {
        const uint32_t myrow = rows[((blockIdx.x * 8 + warp) * NSRC + src) * 32 + lane];
        __syncwarp();
        const long long t0 = clock64();
        if (MODE == 0) {
            const int half = lane >> 4, ch = lane & 15;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int rr = 2 * i + half;
                const uint32_t row = __shfl_sync(0xffffffffu, myrow, rr);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage_s + rr * 256 + (((ch & 8) | ((ch ^ rr) & 7)) << 4)),
                             "l"(tab + (size_t)row * 64 + ch * 4) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else if (MODE == 1) {
            if (lane == 0) mbar_arrive_expect_tx(bar, 32 * 256);
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stage_s + lane * 272),
                         "l"(tab + (size_t)myrow * 64), "r"(256), "r"(smem_u32(bar)) : "memory");
        } else {
            if (lane == 0) mbar_arrive_expect_tx(bar, 32 * 256);
            __syncwarp();
            // lanes 0..15: lane = 2 * group + half; group g gathers rows 4g..4g+3, half h the floats 32h..32h+31
            const int g = (lane & 15) >> 1, h = lane & 1;
            const uint32_t r0 = __shfl_sync(0xffffffffu, myrow, 4 * g), r1 = __shfl_sync(0xffffffffu, myrow, 4 * g + 1),
                           r2 = __shfl_sync(0xffffffffu, myrow, 4 * g + 2), r3 = __shfl_sync(0xffffffffu, myrow, 4 * g + 3);
            if (lane < 16)
                asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                             ::"r"(stage_s + h * 4096 + g * 512), "l"(&tmap), "r"(32 * h), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
        }
        const long long t1 = clock64();
        if (MODE == 0) { asm volatile("cp.async.wait_all;" ::: "memory"); __syncwarp(); }
        else { mbar_wait(bar, phase); phase ^= 1; }
        const long long t2 = clock64();
        t_issue += t1 - t0; t_land += t2 - t0;
        // read this thread's row the way the kernel would (16 x LDS.128) and check it
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t off;
            if (MODE == 0) off = lane * 256 + (((j & 8) | ((j ^ lane) & 7)) << 4);
            else if (MODE == 1) off = lane * 272 + j * 16;
            else off = (j >> 3) * 4096 + lane * 128 + (((j & 7) ^ (lane & 7)) << 4);
            const float4 v = *reinterpret_cast<const float4 *>(stage + off);
            if (verify) {
                const float e0 = (float)(myrow * 64 + 4 * j);
                bad += (v.x != e0) + (v.y != e0 + 1.f) + (v.z != e0 + 2.f) + (v.w != e0 + 3.f);
            }
            s += v.x + v.y + v.z + v.w;
        }
        if (s == -1.f) bad += 1000;
        __syncwarp();
    }
    if (lane == 0) res[blockIdx.x * 8 + warp] = Res{t_issue / NSRC, t_land / NSRC, bad};
    int b = bad;
    for (int d = 16; d; d >>= 1) b += __shfl_xor_sync(0xffffffffu, b, d);
    if (lane == 0) res[blockIdx.x * 8 + warp].bad = b;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    float *tab;
    CK(cudaMalloc(&tab, (size_t)ROWS * 256));
    {
        std::vector<float> h((size_t)ROWS * 64);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 16777216);          // exactly representable
        CK(cudaMemcpy(tab, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    }
    const int nblk = 148;
    std::vector<uint32_t> hr((size_t)nblk * 8 * NSRC * 32);
    uint64_t x = 88172645463325252ull;
    for (auto &r : hr) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; r = (uint32_t)(x % ROWS); }     // ROWS * 64 < 2^24: the float encoding is exact
    uint32_t *rows;
    CK(cudaMalloc(&rows, hr.size() * 4));
    CK(cudaMemcpy(rows, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
    Res *res;
    CK(cudaMalloc(&res, sizeof(Res) * nblk * 8));

    CUtensorMap tmap{};
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
        const cuuint64_t gdim[2] = {64, (cuuint64_t)ROWS}, gstride[1] = {256};
        const cuuint32_t box[2] = {32, 1}, estr[2] = {1, 1};
        const CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    }
    const size_t sh = 8 * 9216;
    CK(cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    CK(cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    CK(cudaFuncSetAttribute(k_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    const char *names[3] = {"L  cp.async 16 B x 16 per lane", "B  bulk 256 B per lane", "G  TMA gather4 (16 per warp)"};
    std::vector<Res> hres(nblk * 8);
    for (int mode = 0; mode < 3; ++mode)
        for (int nw : {1, 4, 8})
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaMemset(res, 0, sizeof(Res) * nblk * 8));
                if (mode == 0) k_probe<0><<<nblk, 256, sh>>>(tab, tmap, rows, nw, res, rep == 0);
                if (mode == 1) k_probe<1><<<nblk, 256, sh>>>(tab, tmap, rows, nw, res, rep == 0);
                if (mode == 2) k_probe<2><<<nblk, 256, sh>>>(tab, tmap, rows, nw, res, rep == 0);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s, %d warps: %s\n", names[mode], nw, cudaGetErrorString(e)); return 1; }
                CK(cudaMemcpy(hres.data(), res, sizeof(Res) * nblk * 8, cudaMemcpyDeviceToHost));
                long long is = 0, la = 0, bad = 0;
                int n = 0;
                for (int b = 0; b < nblk; ++b)
                    for (int w = 0; w < nw; ++w) { is += hres[b * 8 + w].issue; la += hres[b * 8 + w].landed; bad += hres[b * 8 + w].bad; ++n; }
                if (rep == 0) printf("%-34s %d warps: verify %s (%lld wrong floats)\n", names[mode], nw, bad ? "FAILED" : "ok", bad);
                else printf("%-34s %d warps: issue %5lld cycles, landed after %5lld cycles per warp and source (32 rows)\n", names[mode], nw, is / n, la / n);
            }
    return 0;
}
