// pcie_probe.cu -- measures the host<->device strategies available to the host-facing path
// (AoS reb_particle records of 128 B, of which x..az = 72 B / m,r = 16 B are used and
// ax..az = 24 B are produced).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void zc_read(const double* __restrict__ aos, double* __restrict__ out, long n) {
    // each thread reads the first 96 B (12 doubles) of its 128 B record from mapped host memory
    long stride = (long)gridDim.x*blockDim.x;
    for (long i = (long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2* rec = reinterpret_cast<const double2*>(aos + i*16);
        double2 a = rec[0], b = rec[1], c = rec[2], d = rec[3], e = rec[4], f = rec[5];
        out[i] = a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y + e.x + e.y + f.x + f.y;
    }
}
// warp-cooperative: a warp reads 32 records = 4 KB contiguous with 16-byte loads (coalesced 512 B per instruction)
__global__ void zc_read_coalesced(const double2* __restrict__ aos, double* __restrict__ out, long n) {
    long warp = ((long)blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    long nwarps = ((long)gridDim.x*blockDim.x) >> 5;
    for (long w = warp; w*32 < n; w += nwarps) {
        const double2* base = aos + w*32*8;       // 8 double2 per record
        double acc = 0;
        #pragma unroll
        for (int k = 0; k < 8; k++) { double2 v = base[k*32 + lane]; acc += v.x + v.y; }
        out[w*32 + lane] = acc;
    }
}
__global__ void zc_write(double* __restrict__ aos, const double* __restrict__ in, long n) {
    long stride = (long)gridDim.x*blockDim.x;
    for (long i = (long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = in[i];
        double* rec = aos + i*16;
        rec[6] = v; rec[7] = v + 1; rec[8] = v + 2;
    }
}

// sector-exact variants: read bytes [0,96) of each record (x..last_collision; sector 3 = the pointers is never
// touched), write bytes [32,96) (vy vz ax ay | az m r lc: the two sectors holding ax..az) as whole 32-byte sectors
__global__ void zc_read96(const double2* __restrict__ aos, double* __restrict__ out, long n) {
    long warp = ((long)blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    long nwarps = ((long)gridDim.x*blockDim.x) >> 5;
    for (long w = warp; w*32 < n; w += nwarps) {
        const double2* base = aos + w*32*8;
        double acc = 0;
        #pragma unroll
        for (int it = 0; it < 6; it++) { int j = it*32 + (int)lane; double2 v = base[(j/6)*8 + (j%6)]; acc += v.x + v.y; }
        out[w*32 + lane] = acc;
    }
}
__global__ void zc_write64(double2* __restrict__ aos, const double* __restrict__ in, long n) {
    long warp = ((long)blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    long nwarps = ((long)gridDim.x*blockDim.x) >> 5;
    for (long w = warp; w*32 < n; w += nwarps) {
        double2* base = aos + w*32*8;
        double v = in[w*32 + lane];
        #pragma unroll
        for (int it = 0; it < 4; it++) { int j = it*32 + (int)lane; base[(j/4)*8 + 2 + (j%4)] = make_double2(v, v + it); }
    }
}
__global__ void zc_write128(double2* __restrict__ aos, const double* __restrict__ in, long n) {
    long warp = ((long)blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    long nwarps = ((long)gridDim.x*blockDim.x) >> 5;
    for (long w = warp; w*32 < n; w += nwarps) {
        double2* base = aos + w*32*8;
        double v = in[w*32 + lane];
        #pragma unroll
        for (int k = 0; k < 8; k++) base[k*32 + lane] = make_double2(v, v + k);
    }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 10000000;
    size_t bytes = (size_t)n*128;
    void* host = nullptr;
    if (posix_memalign(&host, 4096, bytes)) return 1;
    memset(host, 1, bytes);
    double t0 = now();
    CK(cudaHostRegister(host, bytes, cudaHostRegisterDefault));
    printf("cudaHostRegister %.1f MB: %.1f ms\n", bytes/1e6, (now() - t0)*1e3);
    void* dev; CK(cudaMalloc(&dev, bytes + (1 << 20)));
    double* dsmall; CK(cudaMalloc(&dsmall, n*8*9)); CK(cudaMemset(dsmall, 0, n*8*9));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto timeit = [&](const char* name, auto fn, double useful_bytes) {
        fn(); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            CK(cudaDeviceSynchronize());
            double w0 = now();
            fn();
            CK(cudaDeviceSynchronize());
            float ms = (float)((now() - w0)*1e3);
            if (ms < best) best = ms;
        }
        printf("%-52s %8.2f ms  %7.1f GB/s (useful bytes)\n", name, best, useful_bytes/best/1e6);
    };
    timeit("H2D contiguous 128 B/p", [&] { CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s1)); }, bytes);
    timeit("D2H contiguous 128 B/p", [&] { CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s1)); }, bytes);
    timeit("H2D + D2H concurrent, 128 B/p each", [&] {
        CK(cudaMemcpyAsync(dev, host, bytes/2, cudaMemcpyHostToDevice, s1));
        CK(cudaMemcpyAsync((char*)host + bytes/2, (char*)dev + bytes/2, bytes/2, cudaMemcpyDeviceToHost, s2)); }, bytes);
    timeit("H2D 2D width 96 of pitch 128", [&] { CK(cudaMemcpy2DAsync(dev, 96, host, 128, 96, n, cudaMemcpyHostToDevice, s1)); }, (double)n*96);
    timeit("H2D 2D width 72 of pitch 128", [&] { CK(cudaMemcpy2DAsync(dev, 72, host, 128, 72, n, cudaMemcpyHostToDevice, s1)); }, (double)n*72);
    timeit("D2H 2D width 24 into pitch 128", [&] { CK(cudaMemcpy2DAsync((char*)host + 48, 128, dev, 24, 24, n, cudaMemcpyDeviceToHost, s1)); }, (double)n*24);
    timeit("D2H 2D width 32 into pitch 128", [&] { CK(cudaMemcpy2DAsync((char*)host + 32, 128, dev, 32, 32, n, cudaMemcpyDeviceToHost, s1)); }, (double)n*32);
    // chunked duplex pipeline like the dispatcher's: H2D(c) -> [kernel] -> D2H(c) per stream, 3 streams
    cudaStream_t ps[4]; for (auto& q : ps) CK(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    for (size_t chunk : {size_t(16) << 20, size_t(64) << 20, size_t(256) << 20}) {
        for (int nb : {2, 3, 4}) {
            char name[128]; snprintf(name, sizeof name, "pipeline chunk %zu MB, %d streams: H2D(c); D2H(c)", chunk >> 20, nb);
            timeit(name, [&] {
                size_t nchunks = (bytes + chunk - 1)/chunk;
                for (size_t c = 0; c < nchunks; c++) {
                    size_t off = c*chunk, len = (off + chunk <= bytes) ? chunk : bytes - off;
                    cudaStream_t q = ps[c % nb];
                    char* dbuf = (char*)dev + (c % nb)*chunk;
                    CK(cudaMemcpyAsync(dbuf, (char*)host + off, len, cudaMemcpyHostToDevice, q));
                    CK(cudaMemcpyAsync((char*)host + off, dbuf, len, cudaMemcpyDeviceToHost, q));
                }
            }, 2.0*bytes);
        }
    }
    // same pipeline from a buffer that starts 16 bytes into a page (what glibc hands REBOUND for a large realloc)
    for (int variant = 0; variant < 4; variant++) {
        const size_t off0 = variant == 3 ? 2048 : 16;
        const size_t doff = variant == 1 ? 16 : 0;                         // device address congruent to the host address mod 4096
        char name[160]; snprintf(name, sizeof name, "pipeline 64 MB x3, host = page + %zu, device = aligned + %zu%s", off0, doff, variant == 2 ? ", head/body split at the page" : "");
        const size_t chunk = size_t(64) << 20;
        char* hb = (char*)host + off0;
        const size_t nb = bytes - 8192;
        timeit(name, [&] {
            size_t nchunks = (nb + chunk - 1)/chunk;
            for (size_t c = 0; c < nchunks; c++) {
                size_t off = c*chunk, len = (off + chunk <= nb) ? chunk : nb - off;
                cudaStream_t q = ps[c % 3];
                char* dbuf = (char*)dev + (c % 3)*(chunk + 4096) + doff;
                if (variant == 2) {      // copy [hb+off, next page) separately, then the page-aligned remainder
                    const size_t head = 4096 - off0;
                    CK(cudaMemcpyAsync(dbuf, hb + off, head, cudaMemcpyHostToDevice, q));
                    CK(cudaMemcpyAsync(dbuf + head, hb + off + head, len - head, cudaMemcpyHostToDevice, q));
                    CK(cudaMemcpyAsync(hb + off, dbuf, head, cudaMemcpyDeviceToHost, q));
                    CK(cudaMemcpyAsync(hb + off + head, dbuf + head, len - head, cudaMemcpyDeviceToHost, q));
                    continue;
                }
                CK(cudaMemcpyAsync(dbuf, hb + off, len, cudaMemcpyHostToDevice, q));
                CK(cudaMemcpyAsync(hb + off, dbuf, len, cudaMemcpyDeviceToHost, q));
            }
        }, 2.0*nb);
    }
    // two independent queues: all H2D on one stream, all D2H on another, D2H(c) gated by an event after H2D(c)
    {
        size_t chunk = size_t(64) << 20;
        static cudaEvent_t ev[64]; for (auto& evx : ev) CK(cudaEventCreateWithFlags(&evx, cudaEventDisableTiming));
        timeit("two queues, 64 MB chunks, event-gated D2H", [&] {
            size_t nchunks = (bytes + chunk - 1)/chunk;
            for (size_t c = 0; c < nchunks; c++) {
                size_t off = c*chunk, len = (off + chunk <= bytes) ? chunk : bytes - off;
                CK(cudaMemcpyAsync((char*)dev + off, (char*)host + off, len, cudaMemcpyHostToDevice, s1));
                CK(cudaEventRecord(ev[c % 64], s1));
                CK(cudaStreamWaitEvent(s2, ev[c % 64], 0));
                CK(cudaMemcpyAsync((char*)host + off, (char*)dev + off, len, cudaMemcpyDeviceToHost, s2));
            }
        }, 2.0*bytes);
    }
    // zero-copy
    void* hdev = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&hdev, host, 0);
    if (e == cudaSuccess) {
        timeit("zero-copy read 96 B/p, thread per record", [&] { zc_read<<<148*8, 256, 0, s1>>>((const double*)hdev, dsmall, n); }, (double)n*96);
        timeit("zero-copy read 128 B/p, warp-coalesced", [&] { zc_read_coalesced<<<148*8, 256, 0, s1>>>((const double2*)hdev, dsmall, n); }, (double)n*128);
        timeit("zero-copy write 24 B/p into records", [&] { zc_write<<<148*8, 256, 0, s1>>>((double*)hdev, dsmall, n); }, (double)n*24);
        timeit("zero-copy read 128 (s1) + zero-copy write 24 (s2)", [&] {
            zc_read_coalesced<<<148*4, 256, 0, s1>>>((const double2*)hdev, dsmall, n/2);
            zc_write<<<148*4, 256, 0, s2>>>((double*)hdev + (n/2)*16, dsmall + n/2, n/2); }, (double)(n/2)*152);
        timeit("H2D contiguous (s1) + zero-copy write 24 (s2)", [&] {
            CK(cudaMemcpyAsync(dev, host, bytes/2, cudaMemcpyHostToDevice, s1));
            zc_write<<<148*4, 256, 0, s2>>>((double*)hdev + (n/2)*16, dsmall + n/2, n/2); }, (double)(n/2)*152);
        for (int g : {1, 2, 4, 8, 16}) {
            char name[128];
            snprintf(name, sizeof name, "zero-copy read 96 B/p sector-exact, grid 148x%d", g);
            timeit(name, [&] { zc_read96<<<148*g, 256, 0, s1>>>((const double2*)hdev, dsmall, n); }, (double)n*96);
            snprintf(name, sizeof name, "zero-copy write 64 B/p sector-exact, grid 148x%d", g);
            timeit(name, [&] { zc_write64<<<148*g, 256, 0, s1>>>((double2*)hdev, dsmall, n); }, (double)n*64);
        }
        timeit("zero-copy write 128 B/p coalesced", [&] { zc_write128<<<148*8, 256, 0, s1>>>((double2*)hdev, dsmall, n); }, (double)n*128);
        timeit("zero-copy read 96 (s1) + write 64 (s2), all records each", [&] {
            zc_read96<<<148*4, 256, 0, s1>>>((const double2*)hdev, dsmall, n);
            zc_write64<<<148*4, 256, 0, s2>>>((double2*)hdev, dsmall + n, n); }, (double)n*160);
        timeit("H2D contiguous 128 (s1) + zero-copy write 64 (s2), all records each", [&] {
            CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s1));
            zc_write64<<<148*4, 256, 0, s2>>>((double2*)hdev, dsmall + n, n); }, (double)n*192);
        timeit("zero-copy read 96 (s1) + D2H contiguous 128 (s2), all records each", [&] {
            zc_read96<<<148*4, 256, 0, s1>>>((const double2*)hdev, dsmall, n);
            CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s2)); }, (double)n*224);
    } else printf("no device pointer for registered memory: %s\n", cudaGetErrorString(e));
    // pageable
    CK(cudaHostUnregister(host));
    timeit("H2D contiguous from PAGEABLE memory", [&] { CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s1)); }, bytes);
    return 0;
}
