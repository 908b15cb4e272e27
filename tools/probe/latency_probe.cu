// latency_probe.cu — measures, on the box it runs on: dependent-load latency (L2 hit / HBM), cross-SM flag ping-pong, grid-barrier cost,
// and the loaded L2 latency while other SMs stream HBM.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency_probe latency_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void chase(const uint32_t * p, int n, uint32_t * out, long long * cyc) {
    uint32_t i = 0;
    for (int k = 0; k < 64; ++k) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(i) : "l"(p + i));   // warm
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(i) : "l"(p + i));
    long long t1 = clock64();
    *out = i; *cyc = (t1 - t0) / n;
}
// CTA 0 and CTA 1 (different SMs) ping-pong a flag with release/acquire
__global__ void pingpong(unsigned * flag, int n, long long * cyc) {
    if (threadIdx.x) return;
    const unsigned me = blockIdx.x;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        const unsigned want = 2 * k + me;                       // CTA0 waits for even.. simple: flag counts up; CTA me writes when flag % 2 == me
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v != want);
        asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flag), "r"(want + 1) : "memory");
    }
    long long t1 = clock64();
    if (me == 0) *cyc = (t1 - t0) / n;                          // one full round trip (two hops)
}
// grid barrier (monotonic counter) repeated n times, all CTAs, with `work` cycles of streaming in between optional
__global__ void gridbar(unsigned * bar, int n, long long * cyc) {
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
            unsigned v; const unsigned target = (unsigned) (k + 1) * gridDim.x;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int) (v - target) < 0);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = (t1 - t0) / n;
}
// loaded latency: CTA 0 chases L2-resident pointers while all other CTAs stream `big` from HBM
__global__ void loaded(const uint32_t * p, int n, const uint4 * big, size_t nbig, uint32_t * out, long long * cyc, unsigned * stop) {
    if (blockIdx.x == 0) {
        if (threadIdx.x) return;
        uint32_t i = 0;
        for (int k = 0; k < 2000; ++k) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(i) : "l"(p + i));
        long long t0 = clock64();
        for (int k = 0; k < n; ++k) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(i) : "l"(p + i));
        long long t1 = clock64();
        *out = i; *cyc = (t1 - t0) / n;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(stop), "r"(1u) : "memory");
        return;
    }
    uint4 acc = make_uint4(0, 0, 0, 0);
    size_t idx = (size_t) (blockIdx.x - 1) * blockDim.x + threadIdx.x, stride = (size_t) (gridDim.x - 1) * blockDim.x;
    for (int rep = 0; rep < 64; ++rep) {
        for (size_t i = idx; i < nbig; i += stride * 4) {
            uint4 a = __ldcs(big + i), b = i + stride < nbig ? __ldcs(big + i + stride) : a, c = i + 2 * stride < nbig ? __ldcs(big + i + 2 * stride) : a, d = i + 3 * stride < nbig ? __ldcs(big + i + 3 * stride) : a;
            acc.x ^= a.x ^ b.x ^ c.x ^ d.x;
        }
        unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(stop) : "memory");
        if (v) break;
    }
    if (acc.x == 0x12345678) out[1] = acc.x;
}

int main() {
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock attr: %d kHz\n", clk);
    long long * cyc; uint32_t * out; CK(cudaMallocManaged(&cyc, 64)); CK(cudaMalloc(&out, 64));
    for (size_t bytes : { (size_t) 1 << 20, (size_t) 32 << 20, (size_t) 1 << 30 }) {
        size_t n = bytes / 4; uint32_t * h = (uint32_t *) malloc(bytes), * d;
        // stride pattern that defeats line reuse: jump by 4099 lines (coprime) modulo n
        const size_t step = 32 * 4099 % n ? 32 * 4099 : 32 * 4097;
        for (size_t i = 0; i < n; ++i) h[i] = (uint32_t) ((i + step) % n);
        CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
        chase<<<1, 1>>>(d, 20000, out, cyc); CK(cudaDeviceSynchronize());
        printf("dependent ld.cg latency, %4zu MB footprint: %lld cycles\n", bytes >> 20, *cyc);
        cudaFree(d); free(h);
    }
    unsigned * flag; CK(cudaMalloc(&flag, 256)); CK(cudaMemset(flag, 0, 256));
    pingpong<<<2, 32>>>(flag, 2000, cyc); CK(cudaDeviceSynchronize());
    printf("cross-SM flag ping-pong round trip (2 hops, release/acquire): %lld cycles\n", *cyc);
    for (int threads : { 32, 384 }) {
        CK(cudaMemset(flag, 0, 256));
        void * args[] = { &flag, nullptr, &cyc }; int n = 2000; args[1] = &n;
        CK(cudaLaunchCooperativeKernel((void *) gridbar, dim3(148), dim3(threads), args, 0, 0)); CK(cudaDeviceSynchronize());
        printf("grid barrier (148 CTAs x %d threads, red.release + ld.acquire poll): %lld cycles\n", threads, *cyc);
    }
    {   // loaded latency
        size_t bytes = 1 << 20, n = bytes / 4; uint32_t * h = (uint32_t *) malloc(bytes), * d; const size_t step = 32 * 4099;
        for (size_t i = 0; i < n; ++i) h[i] = (uint32_t) ((i + step) % n);
        CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
        uint4 * big; size_t nbig = ((size_t) 4 << 30) / 16; CK(cudaMalloc(&big, nbig * 16)); CK(cudaMemset(big, 1, nbig * 16));
        unsigned * stop; CK(cudaMalloc(&stop, 4)); CK(cudaMemset(stop, 0, 4));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        loaded<<<148, 512>>>(d, 20000, big, nbig, out, cyc, stop); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("dependent ld.cg latency (1 MB, L2 hits) while 147 SMs stream HBM: %lld cycles  (kernel %.2f ms)\n", *cyc, ms);
    }
    return 0;
}
