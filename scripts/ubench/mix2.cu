// Which part of the per-state instruction set costs FP64 issue rate on sm_100a?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o <name> <name>.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITER 2048
template <int V>
__global__ void k(double *out, double c, double d, long long *cyc, int one) {
    double D[8], P[8], Q[8], R[8];
    unsigned codes = 0;
    for (int i = 0; i < 8; ++i) { D[i] = c * (threadIdx.x + i); P[i] = D[i] + 1; Q[i] = D[i] + 2; R[i] = D[i] + 3; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_ITER; ++it) {
        double x = d + it;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double ae = fabs(x - Q[i]);                       // DADD
            double stay = D[i] + ae;                          // DADD
            double cand = P[(i + 7) & 7] + ae;                // DADD
            P[i] = R[i] + ae;                                 // DADD
            R[i] = stay;
            Q[i] = Q[i] + 1e-30;                              // DADD
            if (V == 0) { D[i] = stay + cand; }               // 6th DADD, no compare
            if (V >= 1) {
                bool p = cand < stay;                         // DSETP
                if (V == 1) { if (p) codes += 1; D[i] = stay; }
                if (V == 2) { D[i] = p ? cand : stay; }       // + 2 FSEL
                if (V == 3) { D[i] = p ? cand : stay; if (p) codes |= 1u << i; }   // + LOP3
                if (V == 4) {                                 // select through predicated moves (FMA pipe?)
                    double b = stay;
                    asm("{.reg .pred q; setp.lt.f64 q, %1, %0; @q mov.f64 %0, %1;}" : "+d"(b) : "d"(cand));
                    D[i] = b;
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += D[i] + P[i] + Q[i] + R[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + codes;
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
    printf("%-44s warps/SMSP=%d  cycles per state per SMSP = %.2f\n", name, w, (double)h / (N_ITER * 8.0 * w));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {3, 4, 6}) {
        run<0>("6 DADD", w);
        run<1>("5 DADD + DSETP (+pred IADD)", w);
        run<2>("5 DADD + DSETP + 2 FSEL", w);
        run<3>("5 DADD + DSETP + 2 FSEL + LOP3", w);
        run<4>("5 DADD + DSETP + pred mov", w);
    }
}
