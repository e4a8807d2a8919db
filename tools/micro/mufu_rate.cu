// MUFU throughput probe (B200): warp-instructions per clock per SM for tanh.approx.f32, tanh.approx.f16x2, tanh.approx.bf16x2,
// ex2.approx.f32, rcp.approx.f32 and plain FFMA, with 1 / 2 / 4 / 8 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate tools/micro/mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define CH 8
template <int OP> __device__ __forceinline__ uint32_t op(uint32_t x) {
    uint32_t y;
    if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
    else if (OP == 1) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    else if (OP == 2) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    else if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
    else if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
    else asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
    return y;
}
template <int OP> __global__ void k(uint32_t* out, long long* cyc) {
    uint32_t v[CH];
    for (int i = 0; i < CH; ++i) v[i] = 0x3c003c00u + threadIdx.x * 8 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = op<OP>(v[i]);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < CH; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, uint32_t* out, long long* cyc) {
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<OP><<<148, warps * 32>>>(out, cyc);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        double winst = (double)ITERS * CH * warps;
        printf("%-18s warps/SM %2d: %7.2f cycles per warp-instruction per SMSP -> %6.2f lanes/clk/SM\n", name, warps,
               avg / (winst / 4), winst * 32 / avg);
    }
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("tanh.approx.f32", out, cyc);
    run<1>("tanh.approx.f16x2", out, cyc);
    run<2>("tanh.approx.bf16x2", out, cyc);
    run<3>("ex2.approx.f32", out, cyc);
    run<4>("rcp.approx.f32", out, cyc);
    run<5>("fma.rn.f32", out, cyc);
    return 0;
}
