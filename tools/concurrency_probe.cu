// Dev tool: can a high-priority kernel start while a low-priority persistent kernel holds most of the SM slots?
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/concurrency_probe.bin tools/concurrency_probe.cu
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 4) spin(long long cycles, int *sink, volatile int *flag) {
    __shared__ int cta_flag;
    if (threadIdx.x == 0) cta_flag = 0;
    __syncthreads();
    const long long t0 = clock64();
    int acc = *(volatile int *)&cta_flag;
    // ~48 live values so that the kernel really needs its 64 registers (occupancy 4 CTAs of 256 threads per SM)
    float v[48];
#pragma unroll
    for (int k = 0; k < 48; k++) v[k] = (float)(threadIdx.x + k) * 1.0001f;
    while (clock64() - t0 < cycles) {
        acc += *flag;
#pragma unroll
        for (int k = 0; k < 48; k++) v[k] = __fmaf_rn(v[k], 1.0001f, (float)acc);
    }
    float sum = 0;
#pragma unroll
    for (int k = 0; k < 48; k++) sum += v[k];
    if (sum == 123456789.0f) *sink = acc;
}
__global__ void __launch_bounds__(256) tiny(int *out) { if (threadIdx.x == 0 && blockIdx.x == 0) *out = 1; }
__global__ void __launch_bounds__(256) tiny2(int *out) { if (threadIdx.x == 0 && blockIdx.x == 0) *out = 2; }
__global__ void __launch_bounds__(256) tiny_smem(int *out) {
    __shared__ unsigned long long s[24];
    if (threadIdx.x < 24) s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = (int)s[3];
}
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    int lo, hi;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStream_t a, b;
    cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, lo);
    cudaStreamCreateWithPriority(&b, cudaStreamNonBlocking, hi);
    int *d = nullptr, *pin = nullptr;
    const bool async_alloc = argc > 1;
    if (async_alloc) { cudaMallocAsync(&d, 256, a); cudaMemsetAsync(d, 0, 256, a); cudaStreamSynchronize(a); }
    else { cudaMalloc(&d, 256); cudaMemset(d, 0, 256); }
    cudaMallocHost(&pin, 64);
    *pin = 0;
    printf("device buffer from %s\n", async_alloc ? "cudaMallocAsync(stream a)" : "cudaMalloc");
    const long long cyc = 300LL * 1900000; // ~300 ms
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spin, 256, 0);
    const int cap = getenv("PROBE_CAP") ? atoi(getenv("PROBE_CAP")) : per_sm;
    const int grid = 148 * cap - (getenv("PROBE_FULL") ? 0 : 2);
    printf("spin: occupancy %d CTAs/SM, grid %d\n", per_sm, grid);
    // warm: load both kernels
    if (getenv("PROBE_CARVEOUT")) {
        const int pct = atoi(getenv("PROBE_CARVEOUT"));
        printf("carveout %d %% on every kernel: %d %d\n", pct, (int)cudaFuncSetAttribute(spin, cudaFuncAttributePreferredSharedMemoryCarveout, pct),
               (int)cudaFuncSetAttribute(tiny_smem, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    tiny<<<1, 32, 0, b>>>(d + 4); tiny2<<<1, 32, 0, b>>>(d + 5); tiny_smem<<<1, 256, 0, b>>>(d + 6); cudaDeviceSynchronize();
    const char *names[] = {"tiny (no smem) x148", "tiny_smem (192 B) x1184", "tiny (no smem) <<<1,1>>>", "tiny_smem then tiny", "tiny, carveout=spin's"};
    for (int what = 0; what < 5; what++) {
        cudaDeviceSynchronize();
        if (what == 4) {
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, spin);
            printf("spin preferredShmemCarveout=%d; setting 8 %% on spin, tiny, tiny_smem\n", fa.preferredShmemCarveout);
            cudaFuncSetAttribute(spin, cudaFuncAttributePreferredSharedMemoryCarveout, 8);
            cudaFuncSetAttribute(tiny, cudaFuncAttributePreferredSharedMemoryCarveout, 8);
            cudaFuncSetAttribute(tiny_smem, cudaFuncAttributePreferredSharedMemoryCarveout, 8);
        }
        const double t0 = now();
        spin<<<grid, 256, 0, a>>>(cyc, d, d + 16);
        while (now() - t0 < 50) {}
        const double t1 = now();
        switch (what) {
        case 0: tiny<<<148, 256, 0, b>>>(d + 4); break;
        case 1: tiny_smem<<<1184, 256, 0, b>>>(d + 6); break;
        case 2: tiny<<<1, 1, 0, b>>>(d + 4); break;
        case 3: tiny_smem<<<1184, 256, 0, b>>>(d + 6); tiny<<<1, 1, 0, b>>>(d + 4); break;
        case 4: tiny<<<148, 256, 0, b>>>(d + 4); break;
        }
        cudaStreamSynchronize(b);
        const double t2 = now();
        cudaStreamSynchronize(a);
        const double t3 = now();
        printf("%-28s on hi-prio stream: back after %8.3f ms   (spin kernel total %7.1f ms)  err=%d\n", names[what], t2 - t1, t3 - t0,
               (int)cudaGetLastError());
    }
    return 0;
}
