// io_ceiling.cu -- what can the step kernel's ACCESS PATTERN reach with no arithmetic at all?
//
// Same streams as k_step<N=4, EPT=2>: per env read 8 B state + 4 B action, write 8 B next state + 8 B reward + 8 B prob +
// 1 B done + 1 B collision (38 B), 128-bit accesses for the 8-byte fields, persistent grid-stride CTAs of 256 threads,
// 4 per SM, a ring of buffers larger than L2.  The result calibrates the roofline fraction of the real kernel:
// MEASURED_PEAKS.json's hbm_gbs is a 50/50 read/write copy, this stream is 68 % writes spread over five arrays.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o io_ceiling tools/io_ceiling.cu && ./io_ceiling
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

typedef unsigned long long u64;

__global__ void __launch_bounds__(256, 4)
k_io(const u64 *__restrict__ states, const int *__restrict__ actions, unsigned n_items, u64 *__restrict__ ns,
     double *__restrict__ reward, double *__restrict__ prob, unsigned char *__restrict__ done,
     unsigned char *__restrict__ coll) {
    // programmatic dependent launch, exactly as k_step does it
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += gridDim.x * blockDim.x) {
        const ulonglong2 s = reinterpret_cast<const ulonglong2 *>(states)[it];
        const int2 a = reinterpret_cast<const int2 *>(actions)[it];
        reinterpret_cast<ulonglong2 *>(ns)[it] = make_ulonglong2(s.x + a.x, s.y + a.y);
        reinterpret_cast<double2 *>(reward)[it] = make_double2((double)a.x, (double)a.y);
        reinterpret_cast<double2 *>(prob)[it] = make_double2(1.0, 0.5);
        reinterpret_cast<unsigned short *>(done)[it] = (unsigned short)(a.x & 0x101);
        reinterpret_cast<unsigned short *>(coll)[it] = (unsigned short)(a.y & 0x101);
    }
}

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e = (x);                                                         \
        if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } \
    } while (0)

int main() {
    const double peak = 6436.1;
    for (long B : {1L << 20, 1L << 23}) {
        const int ring = B == (1L << 20) ? 32 : 4, K = 128;
        std::vector<u64 *> st(ring), ns(ring);
        std::vector<int *> ac(ring);
        std::vector<double *> rw(ring), pb(ring);
        std::vector<unsigned char *> dn(ring), cl(ring);
        for (int j = 0; j < ring; ++j) {
            CK(cudaMalloc(&st[j], B * 8)); CK(cudaMalloc(&ns[j], B * 8)); CK(cudaMalloc(&ac[j], B * 4));
            CK(cudaMalloc(&rw[j], B * 8)); CK(cudaMalloc(&pb[j], B * 8)); CK(cudaMalloc(&dn[j], B)); CK(cudaMalloc(&cl[j], B));
            CK(cudaMemset(st[j], 1, B * 8)); CK(cudaMemset(ac[j], 1, B * 4));
        }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaStream_t stream;
        CK(cudaStreamCreate(&stream));
        for (int mode = 0; mode < 2; ++mode) {  // 0: plain stream launches, 1: one CUDA graph of K PDL launches
            for (int grid : {592, 1184, 148 * 16, (int)(B / 2 / 256)}) {
                auto launch_all = [&]() -> cudaError_t {
                    for (int i = 0; i < K; ++i) {
                        const int j = i % ring;
                        unsigned n_items = (unsigned)(B / 2);
                        void *args[] = {&st[j], &ac[j], &n_items, &ns[j], &rw[j], &pb[j], &dn[j], &cl[j]};
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(grid);
                        cfg.blockDim = dim3(256);
                        cfg.stream = stream;
                        cudaLaunchAttribute attr[1];
                        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        attr[0].val.programmaticStreamSerializationAllowed = mode;
                        cfg.attrs = attr;
                        cfg.numAttrs = 1;
                        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)k_io, args);
                        if (e != cudaSuccess) return e;
                    }
                    return cudaSuccess;
                };
                cudaGraphExec_t exec = nullptr;
                if (mode == 1) {
                    cudaGraph_t graph;
                    CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                    CK(launch_all());
                    CK(cudaStreamEndCapture(stream, &graph));
                    CK(cudaGraphInstantiate(&exec, graph, 0));
                }
                float best = 1e9f;
                for (int rep = 0; rep < 6; ++rep) {
                    cudaEventRecord(e0, stream);
                    if (mode == 1) CK(cudaGraphLaunch(exec, stream));
                    else CK(launch_all());
                    cudaEventRecord(e1, stream);
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                const double us = best * 1e3 / K, gbs = B * 38.0 / (us * 1e-6) / 1e9;
                printf("io_ceiling %s B=%ld grid=%d  %.2f us/launch  %.1f GB/s  %.1f%% of %.1f\n",
                       mode ? "graph+PDL" : "stream   ", B, grid, us, gbs, 100 * gbs / peak, peak);
            }
        }
        for (int j = 0; j < ring; ++j) {
            cudaFree(st[j]); cudaFree(ns[j]); cudaFree(ac[j]); cudaFree(rw[j]); cudaFree(pb[j]); cudaFree(dn[j]); cudaFree(cl[j]);
        }
    }
    return 0;
}
