// fp64_rate.cu -- measured fp64 issue rate / latency of one SM (B200, sm_100a): DFMA, DMUL, DADD, DSETP, DMMA m8n8k4.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k(double *out, long long *cyc, int iters, double seed) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) a[i] = fma(a[i], b, c);
            if (OP == 1) a[i] = a[i] * b;
            if (OP == 2) a[i] = a[i] + c;
            if (OP == 3) a[i] = (a[i] > b) ? a[i] - 1.0 : a[i] + 1.0;  // DSETP + select + DADD
            if (OP == 4) {
                double d0 = a[i], d1 = a[i];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(d0), "+d"(d1) : "d"(b), "d"(c));
                a[i] = d0 + d1 * 0.0;
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int ILP>
void run(const char *name, int threads) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    k<OP, ILP><<<1, threads>>>(out, cyc, iters, 1.0);
    k<OP, ILP><<<1, threads>>>(out, cyc, iters, 1.0);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / iters / ILP;
    printf("%-6s ILP=%d warps=%2d : %.2f cycles per op per warp-slot ; SM rate = %.2f lanes/clk\n", name, ILP, threads / 32, per,
           threads / per);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>("DFMA", 32);   // latency (1 warp, dependent chain)
    run<0, 8>("DFMA", 32);   // 1 warp, 8 independent chains
    run<0, 4>("DFMA", 512);  // 16 warps
    run<0, 4>("DFMA", 1024); // 32 warps: SM throughput
    run<1, 4>("DMUL", 1024);
    run<2, 4>("DADD", 1024);
    run<3, 4>("DSETP+", 1024);
    run<4, 1>("DMMA", 32);
    run<4, 2>("DMMA", 1024);
    return 0;
}
