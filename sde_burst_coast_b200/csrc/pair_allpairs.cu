// pair_allpairs.cu - three-zone social interaction over ALL ordered pairs (metric neighbour rule),
// the GPU counterpart of InteractionGlobal -> IntCalcPrey -> SFM_4Zone
// (/root/reference/src/agents_interact.cpp:144-157,172-250; social_forces.cpp:25-52).
//
// Two kernels, one result:
//   pair_dense_kernel : every alive row is active (reference state at s = 1). One thread
//                       per row, j-tiles of 256 agents staged in shared memory and broadcast to the
//                       warp (one LDS.128 feeds 32 pairs); FP32 fast classification with a rounding
//                       band, the tile is re-done with FP64 zone decisions when any pair of the
//                       thread fell inside the band.
//   pair_rows_kernel  : burst-gated rows (a fraction burst_rate*dt of the agents per step, F2).
//                       Rows are compacted first; one warp per active row, lanes stride over j with
//                       coalesced 16-byte loads, counters by warp reduction.
// Both write the same accumulator layout: acc[g][8] = rep.xy alg.xy att.xy (flee.xy untouched),
// cnt[g][4] = c_rep c_alg c_att (c_flee untouched).
#include "sbc_kernels.cuh"

namespace {

constexpr int DENSE_THREADS = 256;
constexpr int DENSE_TILE = 256;

struct Acc9 {
    float rx, ry, ax, ay, tx, ty;
    int cr, ca, ct;
};

template <bool PERIODIC>
__device__ __forceinline__ void pair_fast(Acc9 &A, bool &amb, float xi, float yi, const float4 pj,
                                          const DevParams &P) {
    float dx = pj.x - xi, dy = pj.y - yi;
    if (PERIODIC) {
        dx = sbc_delta_pbc(xi, pj.x, P);
        dy = sbc_delta_pbc(yi, pj.y, P);
    }
    const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
    const float rinv = rsqrtf(fmaxf(d2, 1e-30f));  // d = 0 -> u = 0 (agents_interact.cpp:199)
    amb = amb || (d2 >= P.band_lo[0] && d2 <= P.band_hi[0]) ||
          (d2 >= P.band_lo[1] && d2 <= P.band_hi[1]) || (d2 >= P.band_lo[2] && d2 <= P.band_hi[2]);
    if (d2 <= P.rep2) {
        A.rx = __fmaf_rn(-dx, rinv, A.rx);
        A.ry = __fmaf_rn(-dy, rinv, A.ry);
        A.cr++;
    } else if (d2 <= P.alg2) {
        A.ax += pj.z;
        A.ay += pj.w;
        A.ca++;
    } else if (d2 <= P.att2) {
        A.tx = __fmaf_rn(dx, rinv, A.tx);
        A.ty = __fmaf_rn(dy, rinv, A.ty);
        A.ct++;
    }
}

template <bool PERIODIC>
__device__ __noinline__ void tile_exact(Acc9 &A, float xi, float yi, const float4 *tile, int n,
                                        const DevParams &P) {
    // slow path: same arithmetic for the force terms, zone decided in FP64 inside the band
    for (int j = 0; j < n; j++) {
        const float4 pj = tile[j];
        float dx = pj.x - xi, dy = pj.y - yi;
        if (PERIODIC) {
            dx = sbc_delta_pbc(xi, pj.x, P);
            dy = sbc_delta_pbc(yi, pj.y, P);
        }
        const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
        const float rinv = rsqrtf(fmaxf(d2, 1e-30f));
        bool amb;
        int z = sbc_zone_fast(d2, P, amb);
        if (amb) z = sbc_zone_exact(xi, yi, pj.x, pj.y, P);
        if (z == 0) {
            A.rx = __fmaf_rn(-dx, rinv, A.rx);
            A.ry = __fmaf_rn(-dy, rinv, A.ry);
            A.cr++;
        } else if (z == 1) {
            A.ax += pj.z;
            A.ay += pj.w;
            A.ca++;
        } else if (z == 2) {
            A.tx = __fmaf_rn(dx, rinv, A.tx);
            A.ty = __fmaf_rn(dy, rinv, A.ty);
            A.ct++;
        }
    }
}

template <bool PERIODIC>
__global__ void __launch_bounds__(DENSE_THREADS)
pair_dense_kernel(const __grid_constant__ DevParams P, DevState S) {
    __shared__ float4 tile[DENSE_TILE];
    __shared__ unsigned long long s_amb;
    const int N = S.N;
    const int rep = blockIdx.y;
    const float4 *pv = S.pv + (size_t)rep * N;
    const int i = blockIdx.x * DENSE_THREADS + threadIdx.x;
    const bool row_ok = i < N;
    float xi = 0.f, yi = 0.f;
    bool alive = false;
    if (row_ok) {
        const float4 p = pv[i];
        xi = p.x;
        yi = p.y;
        alive = (p.x == p.x);  // NaN marks a dead agent
    }
    if (threadIdx.x == 0) s_amb = 0ull;
    Acc9 A = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0, 0, 0};
    unsigned long long n_amb_tiles = 0;
    for (int j0 = 0; j0 < N; j0 += DENSE_TILE) {
        const int n = min(DENSE_TILE, N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += DENSE_THREADS) tile[t] = pv[j0 + t];
        __syncthreads();
        if (alive) {
            const Acc9 snap = A;
            bool amb = false;
            if (n == DENSE_TILE) {
#pragma unroll 8
                for (int j = 0; j < DENSE_TILE; j++) pair_fast<PERIODIC>(A, amb, xi, yi, tile[j], P);
            } else {
                for (int j = 0; j < n; j++) pair_fast<PERIODIC>(A, amb, xi, yi, tile[j], P);
            }
            if (amb) {
                A = snap;
                tile_exact<PERIODIC>(A, xi, yi, tile, n, P);
                n_amb_tiles++;
            }
        }
    }
    if (row_ok) {
        const size_t g = (size_t)rep * N + i;
        float *acc = S.acc + g * 8;
        int *cnt = S.cnt + g * 4;
        if (alive) A.cr -= 1;  // the self pair (d = 0, zero force) was counted as repulsion
        acc[0] = alive ? A.rx : 0.f; acc[1] = alive ? A.ry : 0.f;
        acc[2] = alive ? A.ax : 0.f; acc[3] = alive ? A.ay : 0.f;
        acc[4] = alive ? A.tx : 0.f; acc[5] = alive ? A.ty : 0.f;
        cnt[0] = alive ? A.cr : 0; cnt[1] = alive ? A.ca : 0; cnt[2] = alive ? A.ct : 0;
    }
    if (n_amb_tiles) atomicAdd(&s_amb, n_amb_tiles);
    __syncthreads();
    if (threadIdx.x == 0 && s_amb) atomicAdd(S.stats + 4, s_amb);
}

// ---- burst-gated rows ------------------------------------------------------------------------
__global__ void compact_active_kernel(DevState S, uint32_t burst_steps, int *list, int *count) {
    const size_t total = (size_t)S.N * S.R;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total;
         g += (size_t)gridDim.x * blockDim.x) {
        const bool active = S.alive[g] && S.timers[g].x == burst_steps;
        if (!active) continue;
        const unsigned m = __activemask();
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(count, __popc(m));
        base = __shfl_sync(m, base, leader);
        list[base + __popc(m & ((1u << lane) - 1))] = (int)g;
    }
}

__global__ void __launch_bounds__(256)
pair_rows_kernel(const __grid_constant__ DevParams P, DevState S, const int *__restrict__ list,
                 const int *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_rows = *count;
    const int N = S.N;
    unsigned long long n_amb = 0;
    for (int k = warp; k < n_rows; k += n_warps) {
        const int g = list[k];
        const int rep = g / N, i = g - rep * N;
        const float4 *pv = S.pv + (size_t)rep * N;
        const float4 pi = pv[i];
        RowAcc A;
        row_acc_clear(A);
        for (int j = lane; j < N; j += 32) {
            if (j == i) continue;
            const float4 pj = pv[j];
            float dx = pj.x - pi.x, dy = pj.y - pi.y;
            if (P.BC == 0) {
                dx = sbc_delta_pbc(pi.x, pj.x, P);
                dy = sbc_delta_pbc(pi.y, pj.y, P);
            }
            const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
            bool amb;
            int z = sbc_zone_fast(d2, P, amb);
            if (amb) {
                z = sbc_zone_exact(pi.x, pi.y, pj.x, pj.y, P);
                n_amb++;
            }
            const float rinv = rsqrtf(fmaxf(d2, 1e-30f));
            if (z == 0) {
                A.rep_x = __fmaf_rn(-dx, rinv, A.rep_x);
                A.rep_y = __fmaf_rn(-dy, rinv, A.rep_y);
                A.c_rep++;
            } else if (z == 1) {
                A.alg_x += pj.z;
                A.alg_y += pj.w;
                A.c_alg++;
            } else if (z == 2) {
                A.att_x = __fmaf_rn(dx, rinv, A.att_x);
                A.att_y = __fmaf_rn(dy, rinv, A.att_y);
                A.c_att++;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A.rep_x += __shfl_xor_sync(0xffffffffu, A.rep_x, o);
            A.rep_y += __shfl_xor_sync(0xffffffffu, A.rep_y, o);
            A.alg_x += __shfl_xor_sync(0xffffffffu, A.alg_x, o);
            A.alg_y += __shfl_xor_sync(0xffffffffu, A.alg_y, o);
            A.att_x += __shfl_xor_sync(0xffffffffu, A.att_x, o);
            A.att_y += __shfl_xor_sync(0xffffffffu, A.att_y, o);
        }
        const int cr = __reduce_add_sync(0xffffffffu, A.c_rep);
        const int ca = __reduce_add_sync(0xffffffffu, A.c_alg);
        const int ct = __reduce_add_sync(0xffffffffu, A.c_att);
        if (lane == 0) {
            float *acc = S.acc + (size_t)g * 8;
            int *cnt = S.cnt + (size_t)g * 4;
            acc[0] = A.rep_x; acc[1] = A.rep_y; acc[2] = A.alg_x; acc[3] = A.alg_y;
            acc[4] = A.att_x; acc[5] = A.att_y;
            cnt[0] = cr; cnt[1] = ca; cnt[2] = ct;
        }
    }
    n_amb = __reduce_add_sync(0xffffffffu, (unsigned)n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, n_amb);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(S.stats + 3, (unsigned long long)n_rows * (unsigned long long)(N - 1));
}

// test hook: neighbour ids of every row in ascending j (same zone functions as the kernels above)
__global__ void nn_fill_kernel(const __grid_constant__ DevParams P, DevState S, int all_rows,
                               const long long *__restrict__ nn_ptr, int *__restrict__ nn_idx) {
    const size_t total = (size_t)S.N * S.R;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int N = S.N;
    const int rep = (int)(g / N), i = (int)(g - (size_t)rep * N);
    if (!S.alive[g]) return;
    if (!all_rows && S.timers[g].x != P.burst_steps) return;
    const float4 *pv = S.pv + (size_t)rep * N;
    const float4 pi = pv[i];
    long long w = nn_ptr[g];
    for (int j = 0; j < N; j++) {
        if (j == i) continue;
        const float4 pj = pv[j];
        float dx = pj.x - pi.x, dy = pj.y - pi.y;
        if (P.BC == 0) {
            dx = sbc_delta_pbc(pi.x, pj.x, P);
            dy = sbc_delta_pbc(pi.y, pj.y, P);
        }
        const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
        bool amb;
        int z = sbc_zone_fast(d2, P, amb);
        if (amb) z = sbc_zone_exact(pi.x, pi.y, pj.x, pj.y, P);
        if (z < 3) nn_idx[w++] = j;
    }
}

__global__ void clear_acc_kernel(DevState S) {
    const size_t total = (size_t)S.N * S.R;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total;
         g += (size_t)gridDim.x * blockDim.x) {
        float4 *a = reinterpret_cast<float4 *>(S.acc + g * 8);
        a[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        a[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<int4 *>(S.cnt + g * 4) = make_int4(0, 0, 0, 0);
    }
}

}  // namespace

void sbc_launch_pair_allpairs(const DevParams &P, const DevState &S, int all_rows, cudaStream_t st) {
    const size_t total = (size_t)S.N * S.R;
    const int clear_blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
    if (all_rows) {
        dim3 grid((S.N + DENSE_THREADS - 1) / DENSE_THREADS, S.R);
        if (P.BC == 0) pair_dense_kernel<true><<<grid, DENSE_THREADS, 0, st>>>(P, S);
        else pair_dense_kernel<false><<<grid, DENSE_THREADS, 0, st>>>(P, S);
        return;
    }
    int *list = S.active_list;
    int *count = S.active_count;
    cudaMemsetAsync(count, 0, sizeof(int), st);
    clear_acc_kernel<<<clear_blocks, 256, 0, st>>>(S);
    compact_active_kernel<<<clear_blocks, 256, 0, st>>>(S, P.burst_steps, list, count);
    pair_rows_kernel<<<148 * 4, 256, 0, st>>>(P, S, list, count);
}

void sbc_launch_nn_fill(const DevParams &P, const DevState &S, int all_rows, const long long *nn_ptr,
                        int *nn_idx, cudaStream_t st) {
    const size_t total = (size_t)S.N * S.R;
    nn_fill_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(P, S, all_rows, nn_ptr, nn_idx);
}
