// Micro-benchmark of the FP64 pipe on sm_100a: throughput and latency of the instructions the
// DTW fill is made of (DADD, DSETP, FSEL pairs).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// Run on a B200; prints cycles per warp-instruction per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

#define N_ITER 4096

template <int OP>
__global__ void k(double *out, double c, double d, long long *cyc) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = c * (threadIdx.x + i);
    int cnt = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_ITER; ++it) {
        if (OP == 0) {          // 8 independent DADD
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(d));
        } else if (OP == 1) {   // 8 independent DSETP (+ predicated integer add)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(cnt) : "d"(a[i]), "d"(d));
        } else if (OP == 2) {   // dependent DADD chain (latency)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.f64 %0, %0, %1;" : "+d"(a[0]) : "d"(d));
        } else if (OP == 3) {   // DP critical path: add -> setp -> select, dependent
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("{.reg .pred p; .reg .f64 s; add.f64 s, %0, %1; setp.lt.f64 p, %2, s; selp.f64 %0, %2, s, p;}"
                             : "+d"(a[0]) : "d"(d), "d"(c));
        } else if (OP == 4) {   // the DP's per-state mix, 8 independent states: 5 DADD + 1 DSETP + select + code
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("{.reg .pred p; .reg .f64 e, s, q;\n"
                             "add.f64 e, %2, %3;\n abs.f64 e, e;\n add.f64 s, %0, e;\n add.f64 q, %4, e;\n"
                             "add.f64 %4, %4, e;\n add.f64 %3, %3, e;\n setp.lt.f64 p, q, s;\n selp.f64 %0, q, s, p;\n"
                             "@p or.b32 %1, %1, 16;}"
                             : "+d"(a[i]), "+r"(cnt) : "d"(d), "d"(c), "d"(a[(i + 1) & 7]));
        } else if (OP == 5) {   // 8 independent FSEL pairs via integer select (ALU pipe)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("{.reg .pred p; setp.ne.s32 p, %1, 0; selp.f64 %0, %0, %2, p;}" : "+d"(a[i]) : "r"(cnt + it), "d"(d));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + cnt;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, int ops_per_iter, int warps_per_smsp) {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 8);
    const int threads = 128 * warps_per_smsp;
    k<OP><<<148, threads>>>(out, 1e-9, 0.5, cyc);
    k<OP><<<148, threads>>>(out, 1e-9, 0.5, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)N_ITER * ops_per_iter * warps_per_smsp);
    printf("%-28s warps/SMSP=%d  cycles=%lld  cycles per warp-instr per SMSP=%.3f  (per iter per warp %.1f)\n", name,
           warps_per_smsp, h, per, (double)h / N_ITER);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {1, 2, 4, 8}) {
        run<0>("DADD x8 independent", 8, w);
        run<1>("DSETP x8 independent", 8, w);
        run<2>("DADD dependent chain", 8, w);
        run<3>("DADD->DSETP->SEL dependent", 8, w);
        run<4>("DP state mix x8 (5DADD+DSETP)", 8, w);
        run<5>("SEL.f64 x8", 8, w);
    }
    return 0;
}
