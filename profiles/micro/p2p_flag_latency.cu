// Round-trip latency of a flag written into a peer GPU's memory and polled there (two GPUs, one process, peer access):
// the floor under the decomposed swarm's exchange. Build: nvcc -arch=sm_100a -o p2p_flag_latency p2p_flag_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// `mine` lives on this GPU (polled), `theirs` on the peer (written). The starter writes first.
__global__ void pingpong(volatile int *mine, volatile int *theirs, int rounds, int starter, unsigned long long *out, int fence) {
    unsigned long long t0 = now_ns();
    for (int r = 1; r <= rounds; r++) {
        if (starter) {
            if (fence) __threadfence_system();
            *theirs = r;
            while (*mine < r) {}
        } else {
            while (*mine < r) {}
            if (fence) __threadfence_system();
            *theirs = r;
        }
    }
    out[0] = now_ns() - t0;
}
int main() {
    int n = 0; cudaGetDeviceCount(&n);
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    int *f[2]; unsigned long long *o[2]; cudaStream_t st[2];
    for (int d = 0; d < 2; d++) {
        cudaSetDevice(d); cudaDeviceEnablePeerAccess(1 - d, 0);
        cudaMalloc(&f[d], 256); cudaMemset(f[d], 0, 256); cudaMalloc(&o[d], 8); cudaStreamCreate(&st[d]);
    }
    for (int fence = 0; fence < 2; fence++) {
        const int rounds = 2000;
        for (int d = 0; d < 2; d++) { cudaSetDevice(d); cudaMemset(f[d], 0, 256); cudaDeviceSynchronize(); }
        for (int d = 0; d < 2; d++) { cudaSetDevice(d); pingpong<<<1, 1, 0, st[d]>>>(f[d], f[1 - d], rounds, d == 0, o[d], fence); }
        unsigned long long ns = 0;
        for (int d = 0; d < 2; d++) { cudaSetDevice(d); cudaDeviceSynchronize(); }
        cudaSetDevice(0); cudaMemcpy(&ns, o[0], 8, cudaMemcpyDeviceToHost);
        printf("{\"fence_system\": %d, \"round_trip_us\": %.3f, \"one_way_us\": %.3f}\n", fence, ns / 1e3 / rounds, ns / 2e3 / rounds);
    }
    return 0;
}
