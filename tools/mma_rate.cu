// Issue-rate probe for the legacy warp-level tensor path on sm_100a: mma.sync m16n8k16 f16 vs m16n8k32 s8 (and u4 k64),
// 16 warps per SM, 8 independent accumulator chains per warp.  Prints MMAs per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND>
__global__ void __launch_bounds__(512, 1) probe(int iters, int* sink, long long* cycles) {
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
    float cf[8][4] = {};
    int ci[8][4] = {};
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(cf[c][0]), "+f"(cf[c][1]), "+f"(cf[c][2]), "+f"(cf[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(ci[c][0]), "+r"(ci[c][1]), "+r"(ci[c][2]), "+r"(ci[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k64.row.col.s32.u4.s4.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(ci[c][0]), "+r"(ci[c][1]), "+r"(ci[c][2]), "+r"(ci[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    long long t1 = clock64();
    int s = 0;
    for (int c = 0; c < 8; c++) for (int k = 0; k < 4; k++) s += ci[c][k] + (int)cf[c][k];
    if (s == 123456789) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[KIND] = t1 - t0;
}
int main() {
    int* sink; long long* cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 64);
    const int iters = 4096;
    probe<0><<<148, 512>>>(iters, sink, cyc);
    probe<1><<<148, 512>>>(iters, sink, cyc);
    probe<2><<<148, 512>>>(iters, sink, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3];
    cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    const char* names[3] = {"f16 m16n8k16", "u8.s8 m16n8k32", "u4.s4 m16n8k64"};
    for (int k = 0; k < 3; k++) printf("%-16s %lld cycles, %.3f MMA/clk/SM (16 warps x 8 chains)\n", names[k], h[k], 16.0 * 8 * iters / (double)h[k]);
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
