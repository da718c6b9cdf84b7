// hostpack_probe.cu -- is a host-thread gather/scatter pipeline faster than DMA of whole records?
//
// The host-facing path receives REBOUND's AoS array (128-byte reb_particle records).  Round-1's
// pipeline DMAs whole records both ways (2 x 128 B/particle over PCIe).  This probe measures the
// alternative: T host threads gather only the fields a sweep reads (x..az = 72 B, or x..vz = 48 B
// when the host adds the delta itself) into pinned SoA staging, the DMA engines move those, a
// kernel runs on the SoA chunk, only ax..az (24 B) come back and the threads scatter them.
// Each thread owns whole chunks (own stream, two staging buffers, software-pipelined), so there is
// no cross-thread barrier.   Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// stand-in sweep: reads nin arrays, read-modify-writes (or writes) 3 acceleration arrays
__global__ void sweep(const double* __restrict__ in, double* __restrict__ acc, long n, long stride, int nin, int rmw) {
    for (long i = (long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x*blockDim.x) {
        double s = 0;
        for (int k = 0; k < nin; k++) s += in[k*stride + i];
        for (int k = 0; k < 3; k++) acc[k*stride + i] = (rmw ? acc[k*stride + i] : 0.0) + s*(k + 1);
    }
}

struct Slot {
    double* up_h; double* dn_h; double* up_d; double* dn_d; cudaStream_t s;
};

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 10000000;
    const unsigned hw = std::thread::hardware_concurrency();
    printf("hardware_concurrency %u, n %ld\n", hw, n);
    const size_t bytes = (size_t)n*128;
    double* aos = nullptr;
    if (posix_memalign((void**)&aos, 4096, bytes)) return 1;
    for (long i = 0; i < n*16; i++) aos[i] = (double)(i & 1023);      // first touch by one thread, like a REBOUND realloc
    CK(cudaFree(0));

    for (int mode = 0; mode < 2; mode++) {          // 0: gather x..az (72 B) up, RMW on device; 1: gather x..vz (48 B) up, host adds the delta
        const int nup = mode == 0 ? 9 : 6;
        for (long chunk : {65536L, 262144L}) {
            for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
                if ((unsigned)T > hw && T > 8) continue;
                std::vector<Slot> slots((size_t)T*2);
                for (auto& sl : slots) {
                    CK(cudaMallocHost((void**)&sl.up_h, (size_t)chunk*nup*8));
                    CK(cudaMallocHost((void**)&sl.dn_h, (size_t)chunk*3*8));
                    CK(cudaMalloc((void**)&sl.up_d, (size_t)chunk*9*8));
                    sl.dn_d = sl.up_d + 6*chunk;
                    CK(cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking));
                }
                const long nchunks = (n + chunk - 1)/chunk;
                double best = 1e30;
                for (int rep = 0; rep < 4; rep++) {
                    std::atomic<long> next{0};
                    CK(cudaDeviceSynchronize());
                    const double t0 = now();
                    std::vector<std::thread> th;
                    for (int t = 0; t < T; t++) th.emplace_back([&, t] {
                        CK(cudaSetDevice(0));
                        auto gather = [&](long c, Slot& sl) {
                            const long c0 = c*chunk, m = std::min(chunk, n - c0);
                            const double* src = aos + c0*16;
                            double* d = sl.up_h;
                            for (long i = 0; i < m; i++) {
                                const double* r = src + i*16;
                                for (int k = 0; k < nup; k++) d[k*chunk + i] = r[k];
                            }
                            CK(cudaMemcpyAsync(sl.up_d, sl.up_h, (size_t)chunk*nup*8, cudaMemcpyHostToDevice, sl.s));
                            sweep<<<296, 256, 0, sl.s>>>(sl.up_d, sl.dn_d, m, chunk, 6, mode == 0);
                            CK(cudaMemcpyAsync(sl.dn_h, sl.dn_d, (size_t)chunk*3*8, cudaMemcpyDeviceToHost, sl.s));
                        };
                        auto scatter = [&](long c, Slot& sl) {
                            CK(cudaStreamSynchronize(sl.s));
                            const long c0 = c*chunk, m = std::min(chunk, n - c0);
                            double* dst = aos + c0*16;
                            const double* a = sl.dn_h;
                            if (mode == 0) for (long i = 0; i < m; i++) { double* r = dst + i*16; r[6] = a[i]; r[7] = a[chunk + i]; r[8] = a[2*chunk + i]; }
                            else for (long i = 0; i < m; i++) { double* r = dst + i*16; r[6] += a[i]; r[7] += a[chunk + i]; r[8] += a[2*chunk + i]; }
                        };
                        long pending = -1; int ps = 0, cur = 0;
                        for (;;) {
                            const long c = next.fetch_add(1);
                            if (c >= nchunks) break;
                            gather(c, slots[(size_t)t*2 + cur]);
                            if (pending >= 0) scatter(pending, slots[(size_t)t*2 + ps]);
                            pending = c; ps = cur; cur ^= 1;
                        }
                        if (pending >= 0) scatter(pending, slots[(size_t)t*2 + ps]);
                    });
                    for (auto& x : th) x.join();
                    const double dt = now() - t0;
                    if (dt < best) best = dt;
                }
                printf("mode %s chunk %7ld threads %2d: %8.2f ms  (%.1f GB/s PCIe useful, %.2e particles/s)\n", mode == 0 ? "72up/24down" : "48up/24down+hostadd",
                       chunk, T, best*1e3, (double)n*(nup + 3)*8/best/1e9, n/best);
                for (auto& sl : slots) { cudaFreeHost(sl.up_h); cudaFreeHost(sl.dn_h); cudaFree(sl.up_d); cudaStreamDestroy(sl.s); }
            }
        }
    }
    // for comparison: whole records both ways through pinned memory, 3-stream pipeline (round-1 path)
    {
        CK(cudaHostRegister(aos, bytes, cudaHostRegisterDefault));
        const size_t chunkb = size_t(64) << 20;
        char* dev; CK(cudaMalloc((void**)&dev, 3*chunkb));
        cudaStream_t ps[3]; for (auto& q : ps) CK(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
        double best = 1e30;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaDeviceSynchronize());
            const double t0 = now();
            const size_t nchunks = (bytes + chunkb - 1)/chunkb;
            for (size_t c = 0; c < nchunks; c++) {
                const size_t off = c*chunkb, len = std::min(chunkb, bytes - off);
                CK(cudaMemcpyAsync(dev + (c % 3)*chunkb, (char*)aos + off, len, cudaMemcpyHostToDevice, ps[c % 3]));
                CK(cudaMemcpyAsync((char*)aos + off, dev + (c % 3)*chunkb, len, cudaMemcpyDeviceToHost, ps[c % 3]));
            }
            CK(cudaDeviceSynchronize());
            best = std::min(best, now() - t0);
        }
        printf("whole records both ways, 3 streams x 64 MB: %8.2f ms\n", best*1e3);
    }
    return 0;
}
