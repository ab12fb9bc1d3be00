// Which companion instructions slow a DADD-dominated stream on sm_100a?  Per "state": 5 DADD + 1 DSETP
// (FP64 pipe, 12 cycles) plus a variable set of ALU / FMA-pipe / LSU instructions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o <name> <name>.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITER 2048
template <int V>
__global__ void k(double *out, double c, double d, long long *cyc, int one) {
    __shared__ double sm[1024];
    double D[8], P[8], Q[8];
    unsigned codes = 0, junk = one;
    for (int i = 0; i < 8; ++i) { D[i] = c * (threadIdx.x + i); P[i] = D[i] + 1; Q[i] = D[i] + 2; }
    sm[threadIdx.x] = c;
    const int lane8 = (threadIdx.x & 31) + 32 * (threadIdx.x >> 5 & 7);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_ITER; ++it) {
        double x = d + it;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double ae = fabs(x - Q[i]);                       // DADD
            double stay = D[i] + ae;                          // DADD
            double cand = P[(i + 7) & 7] + ae;                // DADD
            P[i] = P[i] + ae;                                 // DADD
            Q[i] = Q[i] + 1e-30;                              // DADD (stands for the second pipeline add)
            bool p = cand < stay;                             // DSETP
            D[i] = p ? cand : stay;                           // 2 FSEL
            if (p) codes |= 1u << i;                          // predicated LOP3
            if (V & 1) { junk = junk * 3 + i; }               // FMA pipe (IMAD)
            if (V & 2) { junk = (junk ^ (junk >> 3)) | i; }   // 2 ALU ops
            if (V & 4) { if ((i & 1) == 0) { sm[lane8] = D[i]; } else { x += sm[(lane8 + i) & 255] * 1e-300; } }  // LSU
            if (V & 8) { asm volatile("mov.b32 %0, %0;" : "+r"(junk)); }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += D[i] + P[i] + Q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + codes + junk;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int V>
void run(const char *name, int w) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
    k<V><<<148, 128 * w>>>(out, 1e-9, 0.5, cyc, 1);
    k<V><<<148, 128 * w>>>(out, 1e-9, 0.5, cyc, 1);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP=%d  cycles per state per SMSP = %.2f\n", name, w, (double)h / (N_ITER * 8.0 * w));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {3, 4}) {
        run<0>("base (5 DADD+DSETP+2FSEL+LOP3)", w);
        run<1>("+1 IMAD", w);
        run<2>("+2 ALU", w);
        run<4>("+0.5 STS +0.5 LDS", w);
        run<3>("+1 IMAD +2 ALU", w);
        run<7>("+1 IMAD +2 ALU + LSU", w);
    }
}
