// fp64_probe.cu -- measures the FP64 pipe peak (dependent-free DFMA streams) and the cost of the
// IEEE division / sqrt sequences that dominate the force kernels.  MEASURED_PEAKS.json has no
// FP64 figure; this provides the denominator for the FP64-bound sweeps.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double seed, int iters) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = seed + threadIdx.x*1e-3 + j;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) a[j] = fma(a[j], b, c);
            if (MODE == 1) a[j] = a[j]/b + c;          // one IEEE division + add
            if (MODE == 2) a[j] = sqrt(a[j]) + b;      // one IEEE sqrt + add
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out; CK(cudaMalloc(&out, (size_t)sms*8*256*sizeof(double)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000, grid = sms*8;
    const char* names[3] = {"DFMA", "DDIV+DADD", "DSQRT+DADD"};
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int r = 0; r < 4; r++) {
            CK(cudaEventRecord(e0));
            if (mode == 0) k<0><<<grid, 256>>>(out, 1.0, iters);
            if (mode == 1) k<1><<<grid, 256>>>(out, 1.0, iters);
            if (mode == 2) k<2><<<grid, 256>>>(out, 2.0, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        const double ops = (double)grid*256*8*iters;
        printf("{\"op\": \"%s\", \"ops_per_s\": %.4e, \"per_sm_per_clk_at_1965MHz\": %.2f%s}\n", names[mode], ops/(best*1e-3),
               ops/(best*1e-3)/sms/1.965e9, mode == 0 ? ", \"tflops\": " : "");
        if (mode == 0) printf("{\"fp64_fma_tflops\": %.2f, \"sms\": %d}\n", 2*ops/(best*1e-3)/1e12, sms);
    }
    return 0;
}
