// tools/mma_peak_i8.cu -- measures the tcgen05 kind::i8 (u8 x u8 -> s32) ceiling the integer GEMM roofline is quoted against
// on THIS box: every SM issues back-to-back tcgen05.mma (operands resident in shared memory, accumulators in TMEM),
// no global traffic.  Two shapes: cta_group::1 M=128 N=256 K=32 and cta_group::2 M=256 N=256 K=32.
// The run is long enough (seconds) to sit at the power-limited sustained clock, like the GEMM itself; a short burst
// figure is printed too.  Byte-plane ceilings: i8 = this number, i16 / 3, i32 / 10, i64 / 36.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iwekua_b200/csrc -o build/mma_peak_i8 tools/mma_peak_i8.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace wk::tc;

constexpr int BM = 128, BN = 256, BK = 128;  // one 128-byte swizzle row = 128 k-bytes = 4 MMAs of K = 32
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK;  // per CTA (the 2-CTA shape uses half of B per CTA)
constexpr int BATCH = 32;                                     // k-blocks (4 MMAs each) between two commits

template <int CTAS>
__global__ void __launch_bounds__(128, 1) k_i8(const float *seed, uint32_t batches, float *sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_holder;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = CTAS == 1 ? 0u : cluster_ctarank();
    // operands: random bytes (so the multiplier array toggles like real data)
    float *tiles = reinterpret_cast<float *>(smem);
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES / CTAS) / 4; i += blockDim.x) tiles[i] = seed[(i + 977 * blockIdx.x) & 65535];
    fence_proxy_async();
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        if (CTAS == 1) tmem_alloc(&tmem_holder, 512);
        else tmem_alloc_2cta(&tmem_holder, 512);
    }
    tc_fence_before();
    if (CTAS == 1) __syncthreads();
    else cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_holder;
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = umma_idesc_u8(BM * CTAS, BN);
        const uint64_t base = umma_desc_base(16, 1024, UMMA_SW128);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + A_BYTES;
        for (uint32_t b = 0; b < batches; b++) {
            if (b >= 2) mbar_wait(&bar[b & 1], ((b >> 1) - 1) & 1);  // batch b-2 retired: at most two batches in flight
            if (elect_one()) {
                const uint32_t d = tmem + (b & 1) * BN;
                for (int kb = 0; kb < BATCH; kb++) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint64_t da = umma_desc(base, a0 + k * 32), db = umma_desc(base, b0 + k * 32);
                        if (CTAS == 1) mma_i8_ss(d, da, db, idesc, (kb | k) != 0);
                        else mma_i8_ss_2cta(d, da, db, idesc, (kb | k) != 0);
                    }
                }
                if (CTAS == 1) mma_commit(&bar[b & 1]);
                else mma_commit_2cta_multicast(&bar[b & 1], 1);
            }
            __syncwarp();
        }
        for (uint32_t b = batches >= 2 ? batches - 2 : 0; b < batches; b++) mbar_wait(&bar[b & 1], (b >> 1) & 1);
        tc_fence_after();
    }
    tc_fence_before();
    if (CTAS == 1) __syncthreads();
    else cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        if (sink && threadIdx.x == 0 && blockIdx.x == 0xffffffffu) sink[0] = tiles[0];
        if (CTAS == 1) tmem_dealloc(tmem, 512);
        else tmem_dealloc_2cta(tmem, 512);
    }
}

template <int CTAS> static double run(int sms, const float *seed, uint32_t batches, int reps) {
    const int smem = A_BYTES + B_BYTES / CTAS + 1024;
    cudaFuncSetAttribute(k_i8<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms / CTAS * CTAS);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    float *sink = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, k_i8<CTAS>, seed, batches, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) cudaLaunchKernelEx(&cfg, k_i8<CTAS>, seed, batches, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * BM * CTAS * BN * 32 * 4 * BATCH * (double)batches * (sms / CTAS) * reps;
    return flop / (ms * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    std::vector<float> h(65536);
    srand(7);
    for (auto &v : h) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;
    float *seed;
    cudaMalloc(&seed, h.size() * 4);
    cudaMemcpy(seed, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    // burst: ~10 ms launches after idle; sustained: ~4 s back to back
    printf("tcgen05.mma kind::i8   cta_group::1 128x256x32 burst     %8.1f Top/s\n", run<1>(sms, seed, 256, 3));
    printf("tcgen05.mma kind::i8   cta_group::2 256x256x32 burst     %8.1f Top/s\n", run<2>(sms, seed, 256, 3));
    printf("tcgen05.mma kind::i8   cta_group::1 128x256x32 sustained %8.1f Top/s\n", run<1>(sms, seed, 4096, 60));
    printf("tcgen05.mma kind::i8   cta_group::2 256x256x32 sustained %8.1f Top/s\n", run<2>(sms, seed, 4096, 60));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
