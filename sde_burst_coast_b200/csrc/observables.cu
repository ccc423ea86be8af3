// observables.cu - per-replicate swarm observables (subset of Out_swarm,
// /root/reference/src/swarmdyn.cpp:237-344, and basic_particle_averages :486-510):
// out[rep][8] = N_alive, pol_order, avg_x0, avg_x1, avg_v0, avg_v1, avg_speed, NND
// NND is the true mean nearest-neighbour distance (the reference's loop skips neighbour i-1, :275).
// One block per replicate; sums in FP64 so the reduction order does not show at FP32 tolerance.
#include "sbc_kernels.cuh"

namespace {
constexpr int OT = 256;

__global__ void __launch_bounds__(OT)
observables_kernel(const __grid_constant__ DevParams P, DevState S, double *out) {
    __shared__ double sh[8][OT / 32];
    const int rep = blockIdx.x, N = S.N;
    const float4 *pv = S.pv + (size_t)rep * N;
    const float2 *hv = S.hv + (size_t)rep * N;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // n, ux, uy, x, y, vx, vy, speed ; nnd separately
    double nnd = 0;
    for (int i = threadIdx.x; i < N; i += OT) {
        const float4 p = pv[i];
        if (p.x != p.x) continue;
        float c, s;
        sincosf(hv[i].x, &s, &c);
        acc[0] += 1; acc[1] += c; acc[2] += s; acc[3] += p.x; acc[4] += p.y;
        acc[5] += p.z; acc[6] += p.w; acc[7] += sqrtf(p.z * p.z + p.w * p.w);
        float best = -1.f;
        for (int j = 0; j < N; j++) {
            if (j == i) continue;
            const float4 q = pv[j];
            if (q.x != q.x) continue;
            float dx = q.x - p.x, dy = q.y - p.y;
            if (P.BC == 0) { dx = sbc_delta_pbc(p.x, q.x, P); dy = sbc_delta_pbc(p.y, q.y, P); }
            const float d2 = dx * dx + dy * dy;
            if (best < 0.f || d2 < best) best = d2;
        }
        if (best >= 0.f) nnd += sqrtf(best);
    }
    double v[9];
    for (int k = 0; k < 8; k++) v[k] = acc[k];
    v[8] = nnd;
    double tot[9];
    for (int k = 0; k < 9; k++) {
        double x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = x;
        __syncthreads();
        double t = 0;
        for (int w = 0; w < OT / 32; w++) t += sh[0][w];
        tot[k] = t;
    }
    if (threadIdx.x == 0) {
        double *o = out + (size_t)rep * 8;
        const double n = tot[0];
        o[0] = n;
        if (n == 0) {
            for (int k = 1; k < 8; k++) o[k] = 0;
        } else {
            const double ux = tot[1] / n, uy = tot[2] / n;
            o[1] = sqrt(ux * ux + uy * uy);
            o[2] = tot[3] / n; o[3] = tot[4] / n; o[4] = tot[5] / n; o[5] = tot[6] / n;
            o[6] = tot[7] / n;
            o[7] = tot[8] / n;
        }
    }
}
}  // namespace

void sbc_launch_observables(const DevParams &P, const DevState &S, double *out_dev, cudaStream_t st) {
    observables_kernel<<<S.R, OT, 0, st>>>(P, S, out_dev);
}
