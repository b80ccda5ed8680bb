// dependent-chain latencies on one warp / one block (cycles per operation), for the design of rti_solo.cuh
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat tools/micro/lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define REP 256
__global__ void k(double* out, long long* cyc, int nthreads_active)
{
    __shared__ double sm[1024];
    const int t = threadIdx.x;
    double a = out[0] + t * 1e-9, b = out[1], c = out[2];
    long long t0, t1;
    sm[t] = a; __syncthreads();
    // 0: DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; i++) a = fma(a, b, c);
    t1 = clock64(); if (t == 0) cyc[0] = t1 - t0;
    // 1: DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; i++) a = a + c;
    t1 = clock64(); if (t == 0) cyc[1] = t1 - t0;
    // 2: DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; i++) a = a * b;
    t1 = clock64(); if (t == 0) cyc[2] = t1 - t0;
    // 3: 64-bit shuffle chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; i++) a = __shfl_xor_sync(0xffffffffu, a, 1);
    t1 = clock64(); if (t == 0) cyc[3] = t1 - t0;
    // 4: LDS dependent chain (pointer chase through shared memory)
    {
        __shared__ int idx[64];
        if (t < 64) idx[t] = (t + 1) & 63;
        __syncthreads();
        int p = t & 63;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < REP; i++) p = idx[p];
        t1 = clock64(); if (t == 0) cyc[4] = t1 - t0;
        a += p;
    }
    // 5: STS -> __syncthreads -> LDS round trip (value passed to the neighbour thread)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; i++) { sm[t] = a; __syncthreads(); a = sm[(t + 1) % blockDim.x] + 1.0; __syncthreads(); }
    t1 = clock64(); if (t == 0) cyc[5] = t1 - t0;
    // 6: STS -> __syncwarp -> LDS round trip
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; i++) { sm[t] = a; __syncwarp(); a = sm[(t & ~31) | ((t + 1) & 31)] + 1.0; __syncwarp(); }
    t1 = clock64(); if (t == 0) cyc[6] = t1 - t0;
    // 7: reciprocal chain, 8: division chain, 9: rsqrt chain, 10: 1/sqrt chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; i++) a = __drcp_rn(a) + c;
    t1 = clock64(); if (t == 0) cyc[7] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; i++) a = b / a + c;
    t1 = clock64(); if (t == 0) cyc[8] = t1 - t0;
    a = fabs(a) + 1.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; i++) a = rsqrt(a) + 1.0;
    t1 = clock64(); if (t == 0) cyc[9] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; i++) a = 1.0 / sqrt(a) + 1.0;
    t1 = clock64(); if (t == 0) cyc[10] = t1 - t0;
    // 11: two independent DFMA chains (issue rate)
    {
        double a2 = a + 1.0;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < REP; i++) { a = fma(a, b, c); a2 = fma(a2, b, c); }
        t1 = clock64(); if (t == 0) cyc[11] = t1 - t0;
        a += a2;
    }
    // 12: clock64 back to back
    t0 = clock64(); t1 = clock64(); if (t == 0) cyc[12] = (t1 - t0) * REP;
    out[3 + t] = a;
}
int main()
{
    double* d; long long* c; cudaMalloc(&d, 4096 * 8); cudaMalloc(&c, 16 * 8);
    double h[3] = {1.0000001, 0.9999999, 1e-9}; cudaMemcpy(d, h, 24, cudaMemcpyHostToDevice);
    const char* names[13] = {"DFMA", "DADD", "DMUL", "SHFL64", "LDS chase", "STS+syncthreads+LDS+syncthreads", "STS+syncwarp+LDS+syncwarp", "drcp+DADD", "div+DADD", "rsqrt+DADD", "1/sqrt+DADD", "2 DFMA chains (per pair)", "clock64 pair"};
    for (int nt : {32, 96, 224, 256}) {
        for (int rep = 0; rep < 2; rep++) { k<<<1, nt>>>(d, c, nt); cudaDeviceSynchronize(); }
        long long hc[16]; cudaMemcpy(hc, c, 13 * 8, cudaMemcpyDeviceToHost);
        printf("threads %d:", nt);
        for (int i = 0; i < 13; i++) printf("  %s %.1f", names[i], (double)hc[i] / REP);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
