// pcie_probe.cu -- host<->device bandwidth of the box by mechanism: copy engines (cudaMemcpyAsync) versus
// SM-driven access to mapped pinned memory (kernels that read / write host memory in place), alone and
// with both directions busy.  Decides how AMV_MEM_HOST moves the bulk planes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pcie_probe pcie_probe.cu && ./pcie_probe [MiB]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void k_copy(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char **argv) {
    const size_t mib = argc > 1 ? atoi(argv[1]) : 512, bytes = mib << 20, n16 = bytes / 16;
    const int reps = 6;
    void *h_in, *h_out, *d_in, *d_out;
    CK(cudaMallocHost(&h_in, bytes)); CK(cudaMallocHost(&h_out, bytes));
    CK(cudaMalloc(&d_in, bytes)); CK(cudaMalloc(&d_out, bytes));
    CK(cudaMemset(d_out, 1, bytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // mode bits: 1 = H2D by DMA, 2 = D2H by DMA, 4 = H2D by kernel (reads host), 8 = D2H by kernel (writes host)
    const int modes[] = { 1, 2, 3, 4, 8, 12, 1 | 8, 4 | 2 };
    const char *names[] = { "h2d_dma", "d2h_dma", "both_dma", "h2d_kernel", "d2h_kernel", "both_kernel", "h2d_dma+d2h_kernel", "h2d_kernel+d2h_dma" };
    for (int grid = 64; grid <= 592; grid *= 3) {
        printf("{\"buffer_MiB\": %zu, \"kernel_grid\": %d", mib, grid);
        for (int m = 0; m < 8; m++) {
            for (int warm = 0; warm < 2; warm++) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, 0));
                CK(cudaStreamWaitEvent(s1, e0, 0)); CK(cudaStreamWaitEvent(s2, e0, 0));
                for (int r = 0; r < reps; r++) {
                    if (modes[m] & 1) CK(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s1));
                    if (modes[m] & 4) k_copy<<<grid, 256, 0, s1>>>((const uint4 *)h_in, (uint4 *)d_in, n16);
                    if (modes[m] & 2) CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s2));
                    if (modes[m] & 8) k_copy<<<grid, 256, 0, s2>>>((const uint4 *)d_out, (uint4 *)h_out, n16);
                }
                cudaEvent_t f1, f2;
                CK(cudaEventCreate(&f1)); CK(cudaEventCreate(&f2));
                CK(cudaEventRecord(f1, s1)); CK(cudaEventRecord(f2, s2));
                CK(cudaStreamWaitEvent(0, f1, 0)); CK(cudaStreamWaitEvent(0, f2, 0));
                CK(cudaEventRecord(e1, 0));
                CK(cudaDeviceSynchronize());
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (warm) printf(", \"%s_GBs_per_direction\": %.1f", names[m], (double)bytes * reps / (ms / 1e3) / 1e9);
                cudaEventDestroy(f1); cudaEventDestroy(f2);
            }
        }
        printf("}\n");
    }
    return 0;
}
