// Micro-benchmark: what does it cost to retire a CTA and start the next one in its slot?
// 4096 CTAs of 256 threads with 73 KiB of dynamic shared memory (3 per SM), each spinning a fixed time.
// ideal = ceil(4096 / (3 * SMs)) * spin; the difference is launch / retire overhead per CTA generation.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 3) spin_kernel(long long cycles, int* sink) {
    extern __shared__ unsigned char smem[];
    const long long t0 = clock64();
    if (threadIdx.x == 0) smem[0] = 1;
    while (clock64() - t0 < cycles) {}
    if (smem[0] == 2) *sink = 1;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int smem = 73088;
    cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int* sink; cudaMalloc(&sink, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (double us : {1.0, 3.0, 8.0}) {
        const long long cycles = (long long)(us * 1e-6 * khz * 1e3);
        for (int grid : {3 * sms, 4096}) {
            for (int w = 0; w < 3; ++w) spin_kernel<<<grid, 256, smem>>>(cycles, sink);
            cudaEventRecord(a);
            for (int r = 0; r < 20; ++r) spin_kernel<<<grid, 256, smem>>>(cycles, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double waves = (grid + 3 * sms - 1) / (3 * sms);
            printf("spin %.0f us, grid %5d: %.1f us per launch, ideal %.1f us (%.0f waves) -> %.2f us overhead per wave\n", us, grid,
                   ms / 20 * 1e3, waves * us, waves, (ms / 20 * 1e3 - waves * us) / waves);
        }
    }
    return 0;
}
