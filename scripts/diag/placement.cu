// Diagnostic: where do the two warps of a 64-thread CTA land (SM, hardware warp slot) when 14 CTAs share an SM?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o placement placement.cu ; run: ./placement
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(64, 14) k(int* out, long long spin)
{
    extern __shared__ unsigned char sm[];
    unsigned wid, sid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sid));
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2] = sid; out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2 + 1] = wid; }
    long long t0 = clock64();
    while (clock64() - t0 < spin) sm[threadIdx.x] += 1;
}
int main()
{
    const int nb = 4096, smem = 15568;
    int* d; cudaMalloc(&d, nb * 4 * sizeof(int));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<nb, 64, smem>>>(d, 2000000);
    cudaDeviceSynchronize();
    static int h[4096 * 4];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // first wave on SM of block 0 and 1
    for (int target = 0; target < 2; ++target) {
        int sm0 = h[target * 4];
        printf("SM %d:", sm0);
        for (int b = 0; b < nb; ++b) if (h[b * 4] == sm0) printf(" b%d(%d,%d)", b, h[b * 4 + 1], h[b * 4 + 3]);
        printf("\n");
    }
    // histogram of (slot of warp0 % 4, slot of warp1 % 4)
    int hist[4][4] = {};
    for (int b = 0; b < nb; ++b) hist[h[b * 4 + 1] & 3][h[b * 4 + 3] & 3]++;
    for (int a = 0; a < 4; ++a) printf("w0 on %d: w1 on 0..3: %d %d %d %d\n", a, hist[a][0], hist[a][1], hist[a][2], hist[a][3]);
    // with swap bit (slot0 >> 2) & 1: scheduler of the acceptance warp
    int acc[4] = {}, acc0[4] = {};
    for (int b = 0; b < nb; ++b) { int s = (h[b * 4 + 1] >> 2) & 1; acc[(s ? h[b * 4 + 3] : h[b * 4 + 1]) & 3]++; acc0[h[b * 4 + 1] & 3]++; }
    printf("acceptance warps per scheduler: fixed %d %d %d %d, alternating %d %d %d %d\n", acc0[0], acc0[1], acc0[2], acc0[3], acc[0], acc[1], acc[2], acc[3]);
    return 0;
}
