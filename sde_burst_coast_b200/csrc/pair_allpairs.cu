// pair_allpairs.cu - three-zone social interaction over ALL ordered pairs (metric neighbour rule),
// the GPU counterpart of InteractionGlobal -> IntCalcPrey -> SFM_4Zone
// (/root/reference/src/agents_interact.cpp:144-157,172-250; social_forces.cpp:25-52).
//
// This file holds the burst-gated pass: only rows with bin_step == burst_steps are active (a
// fraction burst_rate*dt of the agents per step). Rows are compacted first; one warp per
// active row, lanes stride over j with coalesced 16-byte loads, counters by warp reduction. The
// dense pass (every row active) lives in pair_dense.cu. Both write the same accumulator layout:
// acc[g][8] = rep.xy alg.xy att.xy (flee.xy untouched), cnt[g][4] = c_rep c_alg c_att.
#include "sbc_kernels.cuh"

namespace {

constexpr int ROWS_THREADS = 128;  // one block per active row

__global__ void __launch_bounds__(ROWS_THREADS, 4)
pair_rows_kernel(const __grid_constant__ DevParams P, DevState S, StepArgs SA) {
    const uint32_t step = sbc_step_of(SA);
    const int par = (int)(step % 3u);
    const int2 *__restrict__ list = S.active_list[par];
    __shared__ float shf[ROWS_THREADS / 32][6];
    __shared__ int shi[ROWS_THREADS / 32][3];
    __shared__ float s_flee[2];
    __shared__ int s_cflee;
    const int lane = threadIdx.x & 31;
    const int n_rows = S.active_count[par];
    const int N = S.N;
    unsigned n_amb = 0;
    sbc_trace(S, step, SBC_TR_ROWS, true);
    for (int k = blockIdx.x; k < n_rows; k += gridDim.x) {
        const int g = list[k].x;
        const int rep = g / N, i = g - rep * N;
        const float4 *pv = S.pv + (size_t)rep * N;
        const float4 pi = pv[i];
        RowState rs = {};
        if (SA.finish && threadIdx.x == 0) rs = sbc_row_prefetch(S, (size_t)g);
        RowAcc A;
        row_acc_clear(A);
        for (int j = threadIdx.x; j < N; j += ROWS_THREADS) {
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
            if (P.voronoi && z < 3) {
                DelaunayEdge e;
                sbc_delaunay_begin(e, pi.x, pi.y, sbc_vor_pos(pi.x, pj.x, P), sbc_vor_pos(pi.y, pj.y, P));
                for (int kk = 0; kk < N; kk++) {
                    const float4 pk = pv[kk];
                    if (kk != i && kk != j && pk.x == pk.x) sbc_delaunay_add(e, sbc_vor_pos(pi.x, pk.x, P), sbc_vor_pos(pi.y, pk.y, P));
                }
                if (SA.pred_active)   // the predators are vertices of F2FP's triangulation (agents_interact.cpp:103-110)
                    for (int q = 0; q < S.Npred; q++) {
                        const float4 pq = S.pred_pv[(size_t)rep * S.Npred + q];
                        sbc_delaunay_add(e, sbc_vor_pos(pi.x, pq.x, P), sbc_vor_pos(pi.y, pq.y, P));
                    }
                if (!sbc_delaunay_end(e)) z = 3;
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
        sbc_block_reduce_row<ROWS_THREADS / 32>(A, shf, shi);
        if (SA.pred_active) {   // IntCalcPred of this row by the last warp (warp 0's lane 0 holds the zone sums)
            if (threadIdx.x >= ROWS_THREADS - 32) {
                RowAcc F;
                row_acc_clear(F);
                sbc_detect_row_warp(P, S, (size_t)g, pi, step, lane, F);
                if (lane == 0) { s_flee[0] = F.flee_x; s_flee[1] = F.flee_y; s_cflee = F.c_flee; }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            float *acc = S.acc + (size_t)g * 8;
            int *cnt = S.cnt + (size_t)g * 4;
            acc[0] = A.rep_x; acc[1] = A.rep_y; acc[2] = A.alg_x; acc[3] = A.alg_y;
            acc[4] = A.att_x; acc[5] = A.att_y;
            cnt[0] = A.c_rep; cnt[1] = A.c_alg; cnt[2] = A.c_att;
            if (SA.finish) {
                if (SA.pred_active) { A.flee_x = s_flee[0]; A.flee_y = s_flee[1]; A.c_flee = s_cflee; }
                sbc_finish_row(P, S, (size_t)g, A, rs, step, SA);
            }
        }
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, (unsigned long long)n_amb);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(S.stats + 3, (unsigned long long)n_rows * (unsigned long long)(N - 1));
    __syncthreads();
    sbc_trace(S, step, SBC_TR_ROWS, false);
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
    if (!all_rows && sbc_tm_bin(S.tm[g], P) != P.burst_steps) return;
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
        if (P.voronoi && z < 3) {
            DelaunayEdge e;
            sbc_delaunay_begin(e, pi.x, pi.y, sbc_vor_pos(pi.x, pj.x, P), sbc_vor_pos(pi.y, pj.y, P));
            for (int k = 0; k < N; k++) {
                const float4 pk = pv[k];
                if (k != i && k != j && pk.x == pk.x) sbc_delaunay_add(e, sbc_vor_pos(pi.x, pk.x, P), sbc_vor_pos(pi.y, pk.y, P));
            }
            if (!sbc_delaunay_end(e)) z = 3;
        }
        if (z < 3) nn_idx[w++] = j;
    }
}

}  // namespace

int sbc_pair_rows_blocks(const DevState &S) {
    // grid: enough blocks for the synchronised first burst (every row active), few enough to be cheap
    // to launch when only a handful of rows start a burst
    // resident at once (see sbc_pair_cells_blocks): the bulk update's blocks follow this grid's in the dispatch order
    const size_t total = (size_t)S.N * S.R;
    return (int)min((size_t)148 * 4, max((size_t)16, total / 64));
}

void sbc_launch_pair_rows(const DevParams &P, const DevState &S, const StepArgs &A, cudaStream_t st) {
    sbc_launch_prio(pair_rows_kernel, sbc_pair_rows_blocks(S), ROWS_THREADS, 0, st, P, S, A);
}

void sbc_launch_nn_fill(const DevParams &P, const DevState &S, int all_rows, const long long *nn_ptr,
                        int *nn_idx, cudaStream_t st) {
    const size_t total = (size_t)S.N * S.R;
    nn_fill_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(P, S, all_rows, nn_ptr, nn_idx);
}
