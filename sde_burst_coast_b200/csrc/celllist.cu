// celllist.cu - the burst-gated three-zone pass for LARGE swarms (BASELINE config C5): candidates
// come from a uniform grid of cells at least att_range wide instead of all N agents. Same result
// as InteractionGlobal -> IntCalcPrey -> SFM_4Zone (/root/reference/src/agents_interact.cpp:144-157,
// 172-250): a pair farther apart than att_range contributes nothing, and every agent within
// att_range of row i lies in the 3 x 3 block of cells around i's cell.
//
// Build (stable, hence deterministic): cell key per agent -> CUB radix sort of (key, agent id) ->
// first/last sorted position per cell. Cells are att_range + skin wide, so one build serves several
// steps (Verlet skin): a pass looks around the row's BUILD-TIME cell and reads the candidates' CURRENT
// positions, which is exact as long as nobody moved more than skin/2 since the build - the host side
// bounds that by (steps since build) x (largest possible speed) x dt and rebuilds in time.
// Pass: rows that start a burst are compacted (pair_allpairs.cu), one block per active row, threads
// stride over the agents of the nine cells, zone decisions exactly as on the all-pairs path (FP32
// band + FP64 re-decision), fixed-order block reduction.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "sbc_kernels.cuh"

namespace {

struct CellGeom {
    float ox, oy;        // origin of the grid
    float inv_w, inv_h;  // 1 / cell size
    int ncx, ncy;
    int periodic;
};

__device__ __forceinline__ int cell_coord(float x, float o, float inv, int n) {
    int c = (int)floorf((x - o) * inv);
    return min(max(c, 0), n - 1);
}

// geometry for non-periodic domains from the bounding box of the alive agents (device-side, no
// host round trip); the grid is capped at max_cells by widening the cells
__global__ void cell_geom_kernel(const int *__restrict__ bbox, float min_w, int max_per_axis, CellGeom *geom) {
    auto unord = [](int i) { return __int_as_float((i >= 0) ? i : (i ^ 0x7fffffff)); };
    const float lo_x = unord(bbox[0]), lo_y = unord(bbox[1]), hi_x = unord(bbox[2]), hi_y = unord(bbox[3]);
    const float ex = fmaxf(hi_x - lo_x, 0.f), ey = fmaxf(hi_y - lo_y, 0.f);
    const float w = fmaxf(min_w, fmaxf(ex, ey) / (float)(max_per_axis - 1));
    CellGeom g;
    g.ox = lo_x; g.oy = lo_y;
    g.inv_w = 1.f / w; g.inv_h = 1.f / w;
    g.ncx = min(max_per_axis, (int)floorf(ex / w) + 1);
    g.ncy = min(max_per_axis, (int)floorf(ey / w) + 1);
    g.periodic = 0;
    *geom = g;
}

__global__ void bbox_init_kernel2(int *bbox) {
    bbox[0] = bbox[1] = 0x7fffffff;
    bbox[2] = bbox[3] = (int)0x80000000;
}
__global__ void bbox_kernel2(const float4 *__restrict__ pv, int n, int *bbox) {
    auto ord = [](float f) { const int i = __float_as_int(f); return (i >= 0) ? i : (i ^ 0x7fffffff); };
    float lo_x = INFINITY, lo_y = INFINITY, hi_x = -INFINITY, hi_y = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pv[i];
        if (p.x != p.x) continue;
        lo_x = fminf(lo_x, p.x); hi_x = fmaxf(hi_x, p.x);
        lo_y = fminf(lo_y, p.y); hi_y = fmaxf(hi_y, p.y);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
        hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    if ((threadIdx.x & 31) == 0 && lo_x <= hi_x) {
        atomicMin(bbox + 0, ord(lo_x)); atomicMin(bbox + 1, ord(lo_y));
        atomicMax(bbox + 2, ord(hi_x)); atomicMax(bbox + 3, ord(hi_y));
    }
}

__global__ void cell_key_kernel(const float4 *__restrict__ pv, int n, const CellGeom *__restrict__ geom,
                                uint32_t dead_key, uint32_t *__restrict__ key, int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CellGeom g = *geom;
    const float4 p = pv[i];
    uint32_t k = dead_key;  // dead agents sort behind every cell
    if (p.x == p.x) k = (uint32_t)(cell_coord(p.y, g.oy, g.inv_h, g.ncy) * g.ncx + cell_coord(p.x, g.ox, g.inv_w, g.ncx));
    key[i] = k;
    idx[i] = i;
}

__global__ void cell_bounds_kernel(const uint32_t *__restrict__ key, const int *__restrict__ idx, int n,
                                   uint32_t dead_key, int2 *__restrict__ cell_range) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t k = key[r];
    if (k == dead_key) return;
    if (r == 0 || key[r - 1] != k) cell_range[k].x = r;
    if (r == n - 1 || key[r + 1] != k) cell_range[k].y = r + 1;
}

// ---- counting-sort build for sparse grids (a few agents per cell, e.g. config C5) ---------------------------------
// key + count (atomic per agent), exclusive scan of the counts, scatter (the counters run back down to zero), then one
// thread per cell puts its handful of agents into ascending index order - the same order the stable radix sort gives,
// so results stay reproducible - and writes the cell's range. Four light passes instead of a three-pass radix sort.
__global__ void cell_count_kernel(const float4 *__restrict__ pv, int n, const CellGeom *__restrict__ geom, uint32_t dead_key,
                                  uint32_t *__restrict__ key, int *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CellGeom g = *geom;
    const float4 p = pv[i];
    uint32_t k = dead_key;
    if (p.x == p.x) {
        k = (uint32_t)(cell_coord(p.y, g.oy, g.inv_h, g.ncy) * g.ncx + cell_coord(p.x, g.ox, g.inv_w, g.ncx));
        atomicAdd(cnt + k, 1);
    }
    key[i] = k;
}
__global__ void cell_scatter_kernel(const uint32_t *__restrict__ key, int n, uint32_t dead_key, const int *__restrict__ start,
                                    int *__restrict__ cnt, int *__restrict__ sidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = key[i];
    if (k == dead_key) return;
    sidx[start[k] + atomicSub(cnt + k, 1) - 1] = i;
}
// cell bounds + the agents of every cell in ascending id order (the scatter's order depends on the atomics' timing; sums
// must not). A cell of up to eight agents is fetched at once, sorted in registers by a 19-comparator network and
// written back: two memory round trips per cell instead of a dependent load/store chain per element (37 -> 9 us per
// build at C5 size).
__global__ void cell_finish_kernel(int ncells, const int *__restrict__ start, int *__restrict__ sidx, int2 *__restrict__ cell_range) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int r0 = start[c], r1 = start[c + 1];
    cell_range[c] = make_int2(r0, r1);
    const int n = r1 - r0;
    if (n < 2) return;
    if (n <= 8) {
        int v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (k < n) ? sidx[r0 + k] : 0x7fffffff;
#define SBC_CSWAP(a, b) { const int lo = min(v[a], v[b]), hi = max(v[a], v[b]); v[a] = lo; v[b] = hi; }
        SBC_CSWAP(0, 1) SBC_CSWAP(2, 3) SBC_CSWAP(4, 5) SBC_CSWAP(6, 7)
        SBC_CSWAP(0, 2) SBC_CSWAP(1, 3) SBC_CSWAP(4, 6) SBC_CSWAP(5, 7)
        SBC_CSWAP(1, 2) SBC_CSWAP(5, 6) SBC_CSWAP(0, 4) SBC_CSWAP(3, 7)
        SBC_CSWAP(1, 5) SBC_CSWAP(2, 6)
        SBC_CSWAP(1, 4) SBC_CSWAP(3, 6)
        SBC_CSWAP(2, 4) SBC_CSWAP(3, 5)
        SBC_CSWAP(3, 4)
#undef SBC_CSWAP
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < n) sidx[r0 + k] = v[k];
        return;
    }
    for (int a = r0 + 1; a < r1; a++) {   // fuller cells: insertion sort in place
        const int v = sidx[a];
        int b = a - 1;
        while (b >= r0 && sidx[b] > v) { sidx[b + 1] = sidx[b]; b--; }
        sidx[b + 1] = v;
    }
}

constexpr int CELLS_THREADS = 128;        // warp-per-row kernel: four rows per block
constexpr int CELLS_BLOCK_THREADS = 512;  // block-per-row kernel: a dense neighbourhood holds thousands of candidates

// One active row against the agents of its 3 x 3 cell neighbourhood, swept by a GROUP of TPR threads
// (a warp when the neighbourhood holds a few dozen agents, a whole block when it holds thousands).
// The nine cell ranges are fetched at once by nine threads and laid end to end, so that the group walks ONE
// candidate range: the dependent chain per row is key -> cell bounds -> sorted index -> position (four
// memory round trips) instead of three per cell.
template <bool PERIODIC, int TPR>
__device__ __forceinline__ void cells_row_sweep(const DevParams &P, const DevState &S, const CellGeom &g,
                                                const int2 *__restrict__ cell_range,
                                                const uint32_t *__restrict__ build_key, const int *__restrict__ sidx, int i,
                                                int t /* thread in group */, int *pref /* [10] shared, per group */,
                                                int *first /* [9] shared, per group */, RowAcc &A, unsigned &n_amb,
                                                unsigned long long &n_cand, const StepArgs &SA, uint32_t step) {
    const float4 pi = S.pv[i];
    // the row's cell at build time (agents alive now were alive then: nobody is born between builds)
    const int ck = (int)build_key[i];
    const int cy = ck / g.ncx, cx = ck - cy * g.ncx;
    if (t < 9) {
        const int tc = t;
        int yy = cy + tc / 3 - 1, xx = cx + tc % 3 - 1;
        bool ok = true;
        if (PERIODIC) {
            yy = (yy + g.ncy) % g.ncy;
            xx = (xx + g.ncx) % g.ncx;
        } else {
            ok = yy >= 0 && yy < g.ncy && xx >= 0 && xx < g.ncx;
        }
        int r0 = 0, r1 = 0;
        if (ok) {
            const int2 cr = cell_range[yy * g.ncx + xx];
            r0 = cr.x;
            r1 = cr.y;
        }
        first[tc] = r0;
        pref[tc + 1] = r1 - r0;
    }
    if (TPR <= 32) __syncwarp(); else __syncthreads();
    if (t == 0) {
        int run = 0;
        pref[0] = 0;
        for (int c = 0; c < 9; c++) { run += pref[c + 1]; pref[c + 1] = run; }
        n_cand += (unsigned long long)run;
    }
    if (TPR <= 32) __syncwarp(); else __syncthreads();
    const int total = pref[9];
    row_acc_clear(A);
    if (SA.ghost_ready && (cx <= SA.edge_lo || cx >= SA.edge_hi)) sbc_wait_ghosts(SA, step);   // a row next to a slab edge
#pragma unroll 2
    for (int q = t; q < total; q += TPR) {
        int c = 0;
#pragma unroll
        for (int k = 1; k < 9; k++) c += (q >= pref[k]) ? 1 : 0;
        const int j = sidx[first[c] + (q - pref[c])];
        if (j == i) continue;
        const float4 pj = __ldcg(S.pv + j);   // current position (NaN if the agent died since the build); past L1: a ghost's
                                              // slot may have been written by the exchange while this kernel runs
        float dx = pj.x - pi.x, dy = pj.y - pi.y;
        if (PERIODIC) {
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
}
__device__ __forceinline__ void cells_row_store(const DevState &S, int i, const RowAcc &A) {
    float4 *acc = reinterpret_cast<float4 *>(S.acc + (size_t)i * 8);
    acc[0] = make_float4(A.rep_x, A.rep_y, A.alg_x, A.alg_y);
    S.acc[(size_t)i * 8 + 4] = A.att_x;
    S.acc[(size_t)i * 8 + 5] = A.att_y;
    *reinterpret_cast<int4 *>(S.cnt + (size_t)i * 4) = make_int4(A.c_rep, A.c_alg, A.c_att, 0);
}

// dense neighbourhoods: one block per active row
template <bool PERIODIC>
__global__ void __launch_bounds__(CELLS_BLOCK_THREADS, 2)
pair_cells_kernel(const __grid_constant__ DevParams P, DevState S, const CellGeom *__restrict__ geom,
                  const int2 *__restrict__ cell_range,
                  const uint32_t *__restrict__ build_key, const int *__restrict__ sidx, StepArgs SA) {
    const uint32_t step = sbc_step_of(SA);
    const int par = (int)(step % 3u);
    const int2 *__restrict__ list = S.active_list[par];
    __shared__ float shf[CELLS_BLOCK_THREADS / 32][6];
    __shared__ int shi[CELLS_BLOCK_THREADS / 32][3];
    __shared__ int pref[10], first[9];
    __shared__ float s_flee[2];
    __shared__ int s_cflee;
    const int lane = threadIdx.x & 31;
    const int n_rows = S.active_count[par];
    const CellGeom g = *geom;
    unsigned n_amb = 0;
    unsigned long long n_cand = 0;
    sbc_trace(S, step, SBC_TR_ROWS, true);
    for (int k = blockIdx.x; k < n_rows; k += gridDim.x) {
        const int i = list[k].x;  // single swarm on this path: slot index == global index
        RowState rs = {};
        if (SA.finish && threadIdx.x == 0) rs = sbc_row_prefetch(S, (size_t)i);
        RowAcc A;
        cells_row_sweep<PERIODIC, CELLS_BLOCK_THREADS>(P, S, g, cell_range, build_key, sidx, i, threadIdx.x, pref, first, A,
                                                 n_amb, n_cand, SA, step);
        sbc_block_reduce_row<CELLS_BLOCK_THREADS / 32>(A, shf, shi);
        if (SA.pred_active) {   // IntCalcPred of this row by the last warp (thread 0 holds the zone sums)
            if (threadIdx.x >= CELLS_BLOCK_THREADS - 32) {
                RowAcc F;
                row_acc_clear(F);
                sbc_detect_row_warp(P, S, (size_t)i, S.pv[i], step, lane, F);
                if (lane == 0) { s_flee[0] = F.flee_x; s_flee[1] = F.flee_y; s_cflee = F.c_flee; }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            cells_row_store(S, i, A);
            if (SA.finish) {
                if (SA.pred_active) { A.flee_x = s_flee[0]; A.flee_y = s_flee[1]; A.c_flee = s_cflee; }
                sbc_finish_row<PERIODIC ? 0 : SBC_BC_RUNTIME>(P, S, (size_t)i, A, rs, step, SA);
            }
        }
        __syncthreads();   // pref / first are rewritten by the next row
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, (unsigned long long)n_amb);
    if (threadIdx.x == 0 && n_cand) atomicAdd(S.stats + 3, n_cand);  // candidates examined on this path
    __syncthreads();
    sbc_trace(S, step, SBC_TR_ROWS, false);
}

// one candidate of a row: zone decision as on the all-pairs path (FP32 band + FP64 re-decision), accumulation
template <bool PERIODIC, bool ALG = true>
__device__ __forceinline__ void cells_pair(const DevParams &P, const float4 pi, const float4 pj, RowAcc &A, unsigned &n_amb) {
    float dx = pj.x - pi.x, dy = pj.y - pi.y;
    if (PERIODIC) {
        dx = sbc_delta_pbc(pi.x, pj.x, P);
        dy = sbc_delta_pbc(pi.y, pj.y, P);
    }
    const float d2 = __fmaf_rn(dy, dy, __fmul_rn(dx, dx));   // NaN for a dead or skipped candidate: no zone
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
    } else if (ALG && z == 1) {   // (!ALG: alg_range == rep_range, the zone is empty and pj carries no velocity)
        A.alg_x += pj.z;
        A.alg_y += pj.w;
        A.c_alg++;
    } else if (z == 2) {
        A.att_x = __fmaf_rn(dx, rinv, A.att_x);
        A.att_y = __fmaf_rn(dy, rinv, A.att_y);
        A.c_att++;
    }
}

// sparse neighbourhoods (a few agents per cell): EIGHT lanes per active row, four rows per warp. Lane t walks cell t of
// the row's 3 x 3 block (lane 0 also the ninth), four candidates at a time: the sorted indices of a batch are loaded
// together, then their positions together, so the dependent chain of a row is list -> key -> cell bounds -> indices ->
// positions, five memory round trips whatever the cell holds. Few warps carry a step's rows (about 0.5 % of the
// agents), which leaves the register file to the bulk update that runs beside this kernel.
// The row's own update, NOT inlined: the interior and the edge blocks of the lanes kernel must run the very same
// instructions - an inlined copy per call site may contract floating-point expressions differently, and a row's
// result must not depend on which list it was on. (P, S, A are __grid_constant__ kernel parameters: no copies.)
template <int BCT>
// The row's own state is asked for twice: a prefetch into L1 when the row is picked up (no register is held through the
// sweep - ten registers kept alive there were spilled, and a spill store waits for its load right where it is issued),
// the loads themselves here, after the sweep.
static __device__ __noinline__ void cells_finish_row(const DevParams &P, const DevState &S, size_t g, const RowAcc &acc,
                                                     uint32_t step, const StepArgs &A) {
    const RowState r = sbc_row_prefetch(S, g);
    sbc_finish_row<BCT>(P, S, g, acc, r, step, A);
}
__device__ __forceinline__ void cells_row_touch(const DevState &S, size_t g) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.tm + g));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.hv + g));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.force + g));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.alive + g));
}

// profiling hook (sbc_debug_trace, undecomposed swarms only - the slots are the exchange's otherwise): the LATEST moment
// any warp of the step's rows kernel passed a phase of its dependent chain. Entry 1 of slot 2: row positions and cell
// bounds have arrived; slot 3: candidate indices and positions have arrived and the sweep is done; slot 6: reductions
// done; slot 7: the rows are updated.
__device__ __forceinline__ void cells_phase_mark(const DevState &S, const StepArgs &SA, uint32_t step, int slot, int lane) {
    if (!S.trace || SA.halo_idx || lane != 0) return;
    atomicMax(S.trace + ((size_t)(step % SBC_TRACE_STEPS) * SBC_TR_KERNELS + slot) * 2 + 1, sbc_now_ns());
}

// One pass of a group of eight lanes over the rows of `list`: lane t walks cell t of the row's 3 x 3 block (lane 0 also
// the ninth). The sorted indices of up to eight candidates per cell are loaded together, then their positions together,
// so the dependent chain of a row is list -> (key) -> cell bounds -> indices -> positions whatever the cell holds.
// EDGE: the rows next to a slab edge of a decomposed swarm - everything up to the indices goes on while the ghosts
// are on their way; the group holds right before it reads positions (past L1: the exchange writes them meanwhile).
template <bool PERIODIC, bool EDGE, bool ALG>
__device__ __forceinline__ void cells_lanes_rows(const DevParams &P, const DevState &S, const CellGeom &g,
                                                 const int2 *__restrict__ cell_range, const uint32_t *__restrict__ build_key,
                                                 const int *__restrict__ sidx, const StepArgs &SA, uint32_t step,
                                                 const int2 *__restrict__ list, int n_rows, int bid, int nblk, unsigned &n_amb,
                                                 unsigned &n_cand) {
    // candidates a lane takes per pass: one pass covers 8 lanes x NB of them. A second pass is two more memory round trips
// for the whole warp, and the kernel ends with its slowest warp: at rho = 0.01 a row meets 52 +- 7 candidates, 80 cover
// all but one row in ten thousand (64 left one warp in six with a second pass). The order of a lane's sum does not depend
// on NB (candidate q always goes to lane q mod 8).
#ifndef SBC_ROWS_NB
#define SBC_ROWS_NB 10
#endif
    constexpr int TPR = 8, WPB = CELLS_THREADS / 32, GPW = 32 / TPR, GPB = WPB * GPW, NB = SBC_ROWS_NB;
    __shared__ int s_first[GPB][9], s_pref[GPB][10];   // per row group: start of each cell's range, running candidate count
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / TPR, t = lane % TPR;
    const unsigned gmask = ((1u << TPR) - 1u) << (grp * TPR);
    const float qnan = __int_as_float(0x7fc00000);
    // warp-uniform loop: the groups of a warp take GPW consecutive rows per pass
    for (int k0 = (bid * WPB + warp) * GPW; k0 < n_rows; k0 += nblk * GPB) {
        const int k = k0 + grp;
        const bool has = k < n_rows;
        const int2 e = has ? list[k] : make_int2(0, 0);
        const int i = e.x;
        if (SA.finish && t == 0 && has) cells_row_touch(S, (size_t)i);
        const float4 pi = has ? S.pv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        // the row's cell at build time (agents alive now were alive then: nobody is born between builds); the key
        // travelled with the list entry unless the cell order was rebuilt since
        const int ck = !has ? 0 : ((SA.fresh_cells || e.y < 0) ? (int)build_key[i] : e.y);
        const int cy = ck / g.ncx, cx = ck - cy * g.ncx;
        // The nine cell ranges of the row are laid end to end (lane t fetches the bounds of cell t, lane 0 also the ninth)
        // and the candidates are dealt round-robin to the eight lanes: every lane sweeps total / 8 of them, whatever the
        // single cells hold. (A lane per cell made every row of the warp wait for its fullest cell - twice the work.)
        int *first = s_first[warp * GPW + grp], *pref = s_pref[warp * GPW + grp];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int tc = t + u * TPR;
            if (tc < 9) {
                int c0 = 0, c1 = 0;
                if (has) {
                    int yy = cy + tc / 3 - 1, xx = cx + tc % 3 - 1;
                    bool ok = true;
                    if (PERIODIC) {
                        yy += (yy < 0) ? g.ncy : 0; yy -= (yy >= g.ncy) ? g.ncy : 0;
                        xx += (xx < 0) ? g.ncx : 0; xx -= (xx >= g.ncx) ? g.ncx : 0;
                    } else {
                        ok = yy >= 0 && yy < g.ncy && xx >= 0 && xx < g.ncx;
                    }
                    if (ok) {
                        const int2 cr = cell_range[yy * g.ncx + xx];
                        c0 = cr.x;
                        c1 = cr.y;
                    }
                }
                first[tc] = c0;
                pref[tc + 1] = c1 - c0;
            }
        }
        __syncwarp();
        if (t == 0) {
            int run = 0;
            pref[0] = 0;
#pragma unroll
            for (int c = 0; c < 9; c++) { run += pref[c + 1]; pref[c + 1] = run; }
        }
        __syncwarp();
        const int total = pref[9];
        if (!EDGE) cells_phase_mark(S, SA, step, 2, lane);
        if (t == 0) n_cand += (unsigned)total;
        RowAcc A;
        row_acc_clear(A);
        for (int qb = t; qb < total; qb += TPR * NB) {
            int jj[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const int q = qb + b * TPR;
                int c = 0;
#pragma unroll
                for (int kk = 1; kk < 9; kk++) c += (q >= pref[kk]) ? 1 : 0;
                jj[b] = (q < total) ? sidx[first[c] + (q - pref[c])] : -1;
            }
            if (EDGE) sbc_wait_ghosts(SA, step);
            if (!ALG) {
                // no alignment zone (alg_range == rep_range: natPred, the default tank): only the candidates' POSITIONS
                // are read - eight bytes each, all eight of a lane in flight at once (one round trip per pass; with the
                // velocities the registers allow four at a time, i.e. two)
                float2 qj[NB];
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    const float2 *src = reinterpret_cast<const float2 *>(S.pv + (jj[b] >= 0 ? jj[b] : 0));
                    qj[b] = !(jj[b] >= 0 && jj[b] != i) ? make_float2(qnan, qnan) : EDGE ? __ldcg(src) : *src;
                }
#pragma unroll
                for (int b = 0; b < NB; b++) cells_pair<PERIODIC, false>(P, pi, make_float4(qj[b].x, qj[b].y, 0.f, 0.f), A, n_amb);
            } else {
#pragma unroll
                for (int h = 0; h < NB; h += 4) {
                    float4 pj[4];
#pragma unroll
                    for (int b = 0; b < 4; b++)
                        if (h + b < NB)
                            pj[b] = !(jj[h + b] >= 0 && jj[h + b] != i) ? make_float4(qnan, qnan, 0.f, 0.f)
                                    : EDGE ? __ldcg(S.pv + jj[h + b]) : S.pv[jj[h + b]];
#pragma unroll
                    for (int b = 0; b < 4; b++)
                        if (h + b < NB) cells_pair<PERIODIC, true>(P, pi, pj[b], A, n_amb);
                }
            }
        }
        if (!EDGE) cells_phase_mark(S, SA, step, 3, lane);
        A.c_rep = __reduce_add_sync(gmask, A.c_rep);
        A.c_alg = __reduce_add_sync(gmask, A.c_alg);
        A.c_att = __reduce_add_sync(gmask, A.c_att);
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) {
            A.rep_x += __shfl_xor_sync(0xffffffffu, A.rep_x, o); A.rep_y += __shfl_xor_sync(0xffffffffu, A.rep_y, o);
            A.alg_x += __shfl_xor_sync(0xffffffffu, A.alg_x, o); A.alg_y += __shfl_xor_sync(0xffffffffu, A.alg_y, o);
            A.att_x += __shfl_xor_sync(0xffffffffu, A.att_x, o); A.att_y += __shfl_xor_sync(0xffffffffu, A.att_y, o);
        }
        if (!EDGE) cells_phase_mark(S, SA, step, 6, lane);
        if (t == 0 && has) {
            cells_row_store(S, i, A);
            if (SA.finish) cells_finish_row<PERIODIC ? 0 : SBC_BC_RUNTIME>(P, S, (size_t)i, A, step, SA);
        }
        __syncwarp();
        if (!EDGE) cells_phase_mark(S, SA, step, 7, lane);
    }
}

// The rows next to a slab edge while the exchange is in flight (GHOSTS): a whole WARP per row, three lanes per cell of
// the 3 x 3 block, so that a lane holds at most four candidates of a cell of up to twelve agents. Everything up to the
// candidates' indices goes on while the ghosts are on their way; after the wait ONE round of position loads (past L1:
// the exchange wrote them meanwhile) is all that is left before the row's update. These rows are few (two cell columns
// per slab edge); their sums are accumulated in another - equally fixed - order than the eight-lane sweep's.
template <bool PERIODIC>
__device__ __forceinline__ void cells_warp_edge_rows(const DevParams &P, const DevState &S, const CellGeom &g,
                                                     const int2 *__restrict__ cell_range, const uint32_t *__restrict__ build_key,
                                                     const int *__restrict__ sidx, const StepArgs &SA, uint32_t step,
                                                     const int2 *__restrict__ list, int n_rows, int bid, int nblk, unsigned &n_amb,
                                                     unsigned &n_cand) {
    constexpr int WPB = CELLS_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tc = lane / 3, sub = lane % 3;   // cell of the 3 x 3 block (lanes 27..31 idle), place among its three lanes
    const float qnan = __int_as_float(0x7fc00000);
    for (int k = bid * WPB + warp; k < n_rows; k += nblk * WPB) {
        const int2 e = list[k];
        const int i = e.x;
        if (SA.finish && lane == 0) cells_row_touch(S, (size_t)i);
        const float4 pi = S.pv[i];
        const int ck = (SA.fresh_cells || e.y < 0) ? (int)build_key[i] : e.y;
        const int cy = ck / g.ncx, cx = ck - cy * g.ncx;
        int r0 = 0, r1 = 0;
        if (tc < 9) {
            int yy = cy + tc / 3 - 1, xx = cx + tc % 3 - 1;
            bool ok = true;
            if (PERIODIC) {
                yy += (yy < 0) ? g.ncy : 0; yy -= (yy >= g.ncy) ? g.ncy : 0;
                xx += (xx < 0) ? g.ncx : 0; xx -= (xx >= g.ncx) ? g.ncx : 0;
            } else {
                ok = yy >= 0 && yy < g.ncy && xx >= 0 && xx < g.ncx;
            }
            if (ok) {
                const int2 cr = cell_range[yy * g.ncx + xx];
                r0 = cr.x;
                r1 = cr.y;
            }
        }
        if (sub == 0) n_cand += (unsigned)(r1 - r0);
        RowAcc A;
        row_acc_clear(A);
        for (int r = r0 + sub; r < r1; r += 12) {
            int jj[4];
            float4 pj[4];
#pragma unroll
            for (int b = 0; b < 4; b++) jj[b] = (r + 3 * b < r1) ? sidx[r + 3 * b] : -1;
            sbc_wait_ghosts(SA, step);
            if (S.trace) {   // when the first / last waiting lane of the step got through (kernel id 7)
                unsigned long long *w = S.trace + ((size_t)(step % SBC_TRACE_STEPS) * SBC_TR_KERNELS + 7) * 2;
                const unsigned long long now = sbc_now_ns();
                atomicMin(w, now);
                atomicMax(w + 1, now);
            }
#pragma unroll
            for (int b = 0; b < 4; b++)
                pj[b] = (jj[b] >= 0 && jj[b] != i) ? __ldcg(S.pv + jj[b]) : make_float4(qnan, qnan, 0.f, 0.f);
#pragma unroll
            for (int b = 0; b < 4; b++) cells_pair<PERIODIC>(P, pi, pj[b], A, n_amb);
        }
        A.c_rep = __reduce_add_sync(0xffffffffu, A.c_rep);
        A.c_alg = __reduce_add_sync(0xffffffffu, A.c_alg);
        A.c_att = __reduce_add_sync(0xffffffffu, A.c_att);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A.rep_x += __shfl_xor_sync(0xffffffffu, A.rep_x, o); A.rep_y += __shfl_xor_sync(0xffffffffu, A.rep_y, o);
            A.alg_x += __shfl_xor_sync(0xffffffffu, A.alg_x, o); A.alg_y += __shfl_xor_sync(0xffffffffu, A.alg_y, o);
            A.att_x += __shfl_xor_sync(0xffffffffu, A.att_x, o); A.att_y += __shfl_xor_sync(0xffffffffu, A.att_y, o);
        }
        if (lane == 0) {
            cells_row_store(S, i, A);
            if (SA.finish) cells_finish_row<PERIODIC ? 0 : SBC_BC_RUNTIME>(P, S, (size_t)i, A, step, SA);
        }
        __syncwarp();
    }
}

// sparse neighbourhoods (a few agents per cell): EIGHT lanes per active row, four rows per warp. Few warps carry a step's
// rows (about 0.5 % of the agents), which leaves the register file to the bulk update that runs beside this kernel.
// Decomposed swarm: the last SA.edge_blocks blocks of the grid take the rows next to a slab edge (listed apart by
// whoever re-armed them); with GHOSTS they wait for the exchange that runs beside this kernel. The other blocks never
// see a ghost and never wait.
template <bool PERIODIC, bool GHOSTS, bool ALG>
// Register budget of the rows kernel: 80 (six resident blocks at most). Only the ~2 blocks per SM that carry rows stay -
// a third of the register file; the others exit at once. At 64 registers the sweep spilled (A/B on one GPU, C5: rows
// kernel done after 20.5 us at 64 registers, 19.3 at 80; with ten candidates per lane and pass 18.2 / 16.9).
#ifndef SBC_ROWS_MINB
#define SBC_ROWS_MINB 6
#endif
__global__ void __launch_bounds__(CELLS_THREADS, SBC_ROWS_MINB)
pair_cells_lanes_kernel(const __grid_constant__ DevParams P, const __grid_constant__ DevState S,
                        const CellGeom *__restrict__ geom, const int2 *__restrict__ cell_range,
                        const uint32_t *__restrict__ build_key, const int *__restrict__ sidx,
                        const __grid_constant__ StepArgs SA) {
    const uint32_t step = sbc_step_of(SA);
    const int par = (int)(step % 3u);
    const int n_main = (int)gridDim.x - SA.edge_blocks;
    const bool edge_part = (int)blockIdx.x >= n_main;
    const CellGeom g = *geom;
    const int lane = threadIdx.x & 31;
    unsigned n_amb = 0, n_cand = 0;
    if (!edge_part) {
        sbc_trace(S, step, SBC_TR_ROWS, true);
        cells_lanes_rows<PERIODIC, false, ALG>(P, S, g, cell_range, build_key, sidx, SA, step, S.active_list[par], S.active_count[par],
                                          (int)blockIdx.x, n_main, n_amb, n_cand);
    } else if (GHOSTS) {
        if (S.trace && threadIdx.x == 0)
            atomicMin(S.trace + ((size_t)(step % SBC_TRACE_STEPS) * SBC_TR_KERNELS + SBC_TR_EDGE) * 2, sbc_now_ns());
        cells_warp_edge_rows<PERIODIC>(P, S, g, cell_range, build_key, sidx, SA, step, S.edge_list[par], S.active_count[4 + par],
                                       (int)blockIdx.x - n_main, SA.edge_blocks, n_amb, n_cand);
    } else {
        cells_lanes_rows<PERIODIC, false, ALG>(P, S, g, cell_range, build_key, sidx, SA, step, S.edge_list[par], S.active_count[4 + par],
                                          (int)blockIdx.x - n_main, SA.edge_blocks, n_amb, n_cand);
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, (unsigned long long)n_amb);
    n_cand = __reduce_add_sync(0xffffffffu, n_cand);
    if (lane == 0 && n_cand) atomicAdd(S.stats + 3, (unsigned long long)n_cand);
    __syncthreads();
    if (!edge_part) sbc_trace(S, step, SBC_TR_ROWS, false);
    else if (S.trace && threadIdx.x == 0)   // the edge blocks' last exit, apart (kernel id SBC_TR_EDGE, entry 1)
        atomicMax(S.trace + ((size_t)(step % SBC_TRACE_STEPS) * SBC_TR_KERNELS + SBC_TR_EDGE) * 2 + 1, sbc_now_ns());
}

// the same neighbourhoods in the predator phase: one WARP per active row, which also evaluates IntCalcPred for it
template <bool PERIODIC>
__global__ void __launch_bounds__(CELLS_THREADS, 4)
pair_cells_warp_kernel(const __grid_constant__ DevParams P, DevState S, const CellGeom *__restrict__ geom,
                       const int2 *__restrict__ cell_range,
                       const uint32_t *__restrict__ build_key, const int *__restrict__ sidx, StepArgs SA) {
    const uint32_t step = sbc_step_of(SA);
    const int par = (int)(step % 3u);
    const int2 *__restrict__ list = S.active_list[par];
    constexpr int WPB = CELLS_THREADS / 32;
    __shared__ int pref[WPB][10], first[WPB][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_rows = S.active_count[par];
    const CellGeom g = *geom;
    unsigned n_amb = 0;
    unsigned long long n_cand = 0;
    sbc_trace(S, step, SBC_TR_ROWS, true);
    for (int k = blockIdx.x * WPB + warp; k < n_rows; k += gridDim.x * WPB) {
        const int i = list[k].x;
        RowState rs = {};
        if (SA.finish && lane == 0) rs = sbc_row_prefetch(S, (size_t)i);
        RowAcc A;
        cells_row_sweep<PERIODIC, 32>(P, S, g, cell_range, build_key, sidx, i, lane, pref[warp], first[warp], A, n_amb,
                                      n_cand, SA, step);
        A.c_rep = __reduce_add_sync(0xffffffffu, A.c_rep);
        A.c_alg = __reduce_add_sync(0xffffffffu, A.c_alg);
        A.c_att = __reduce_add_sync(0xffffffffu, A.c_att);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A.rep_x += __shfl_xor_sync(0xffffffffu, A.rep_x, o); A.rep_y += __shfl_xor_sync(0xffffffffu, A.rep_y, o);
            A.alg_x += __shfl_xor_sync(0xffffffffu, A.alg_x, o); A.alg_y += __shfl_xor_sync(0xffffffffu, A.alg_y, o);
            A.att_x += __shfl_xor_sync(0xffffffffu, A.att_x, o); A.att_y += __shfl_xor_sync(0xffffffffu, A.att_y, o);
        }
        if (lane == 0) cells_row_store(S, i, A);
        __syncwarp();
        if (SA.pred_active) sbc_detect_row_warp(P, S, (size_t)i, S.pv[i], step, lane, A);
        if (SA.finish && lane == 0) sbc_finish_row<PERIODIC ? 0 : SBC_BC_RUNTIME>(P, S, (size_t)i, A, rs, step, SA);
        __syncwarp();
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, (unsigned long long)n_amb);
    if (lane == 0 && n_cand) atomicAdd(S.stats + 3, n_cand);
    __syncthreads();
    sbc_trace(S, step, SBC_TR_ROWS, false);
}

}  // namespace

// smallest grid the 3 x 3 sweep is valid on
bool sbc_cells_supported(const DevParams &P, int N, int R) {
    if (R != 1) return false;
    if (P.BC == 0) return (int)floor((double)P.L / ((double)P.att * 1.00001)) >= 3;
    return true;
}

// widest skin (up to `want`) that still leaves three cells per axis in a periodic box
float sbc_cells_skin(const DevParams &P, float want) {
    if (P.BC != 0) return want;
    const double room = (double)P.L / 3.0 / 1.00001 - (double)P.att;
    return (float)fmax(0.0, fmin((double)want, room));
}

cudaError_t sbc_cell_scratch_alloc(CellScratch &C, const DevParams &P, int N) {
    cudaError_t e;
    C.max_per_axis = 2048;
    CellGeom hg;
    if (P.BC == 0) {
        const int n = (int)floor((double)P.L / (((double)P.att + (double)C.skin) * 1.00001));
        hg.ox = hg.oy = 0.f;
        hg.inv_w = hg.inv_h = (float)((double)n / (double)P.L);
        hg.ncx = hg.ncy = n;
        hg.periodic = 1;
        C.ncells = (size_t)n * n;
        C.ncx = n;
    } else {
        C.ncells = (size_t)C.max_per_axis * C.max_per_axis;
    }
    if ((e = cudaMalloc((void **)&C.key_in, (size_t)N * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.key_out, (size_t)N * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.idx_in, (size_t)N * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.idx_out, (size_t)N * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.cell_range, C.ncells * sizeof(int2))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.bbox, 4 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&C.geom, sizeof(CellGeom))) != cudaSuccess) return e;
    if (P.BC == 0) cudaMemcpy(C.geom, &hg, sizeof(hg), cudaMemcpyHostToDevice);
    C.key_bits = 1;
    while (((size_t)1 << C.key_bits) <= C.ncells) C.key_bits++;
    C.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, C.cub_bytes, C.key_in, C.key_out, C.idx_in, C.idx_out, N, 0, C.key_bits);
    // sparse grids (periodic box, at most 32 agents per cell on average) are built by a counting sort
    C.counting = P.BC == 0 && (double)N / (double)C.ncells <= 32.0;
    if (C.counting) {
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, C.cnt, C.start, (int)C.ncells + 1);
        if (scan_bytes > C.cub_bytes) C.cub_bytes = scan_bytes;
        if ((e = cudaMalloc((void **)&C.cnt, (C.ncells + 1) * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&C.start, (C.ncells + 1) * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMemset(C.cnt, 0, (C.ncells + 1) * sizeof(int))) != cudaSuccess) return e;
    }
    if ((e = cudaMalloc(&C.cub_tmp, C.cub_bytes ? C.cub_bytes : 16)) != cudaSuccess) return e;
    return cudaSuccess;
}

void sbc_cell_scratch_free(CellScratch &C) {
    cudaFree(C.key_in); cudaFree(C.key_out); cudaFree(C.idx_in); cudaFree(C.idx_out);
    cudaFree(C.cell_range); cudaFree(C.bbox); cudaFree(C.geom); cudaFree(C.cub_tmp); cudaFree(C.cnt); cudaFree(C.start);
    C = CellScratch();
}

void sbc_launch_cell_build(const DevParams &P, const DevState &S, CellScratch &C, cudaStream_t st) {
    const int N = S.N;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    CellGeom *geom = static_cast<CellGeom *>(C.geom);
    if (P.BC != 0) {
        bbox_init_kernel2<<<1, 1, 0, st>>>(C.bbox);
        bbox_kernel2<<<148, 256, 0, st>>>(S.pv, N, C.bbox);
        cell_geom_kernel<<<1, 1, 0, st>>>(C.bbox, (P.att + C.skin) * 1.00001f, C.max_per_axis, geom);
    }
    const uint32_t dead_key = (uint32_t)C.ncells;
    if (C.counting) {
        cell_count_kernel<<<blocks, 256, 0, st>>>(S.pv, N, geom, dead_key, C.key_in, C.cnt);
        size_t bytes = C.cub_bytes;
        cub::DeviceScan::ExclusiveSum(C.cub_tmp, bytes, C.cnt, C.start, (int)C.ncells + 1, st);
        cell_scatter_kernel<<<blocks, 256, 0, st>>>(C.key_in, N, dead_key, C.start, C.cnt, C.idx_out);
        cell_finish_kernel<<<(unsigned)((C.ncells + 255) / 256), 256, 0, st>>>((int)C.ncells, C.start, C.idx_out, C.cell_range);
        return;
    }
    cell_key_kernel<<<blocks, 256, 0, st>>>(S.pv, N, geom, dead_key, C.key_in, C.idx_in);
    cub::DeviceRadixSort::SortPairs(C.cub_tmp, C.cub_bytes, C.key_in, C.key_out, C.idx_in, C.idx_out, N, 0,
                                    C.key_bits, st);
    cudaMemsetAsync(C.cell_range, 0, C.ncells * sizeof(int2), st);
    cell_bounds_kernel<<<blocks, 256, 0, st>>>(C.key_out, C.idx_out, N, dead_key, C.cell_range);
}

// mean number of candidates a row meets in its nine cells decides between a warp and a block per row
static bool cells_warp_rows(const DevParams &P, const DevState &S, const CellScratch &C) {
    const double per_row = 9.0 * (double)S.N / (double)(C.ncells > 0 ? C.ncells : 1);
    return P.BC == 0 && per_row < 512.0;
}
// The rows kernel is the first of the step's two kernels; the bulk update's blocks are only dispatched once every
// block of this grid has been, so the grid must be resident at once: 4 blocks of four warps per SM in the
// lanes-per-row kernel (four rows per warp: 9472 rows in flight cover the ~0.5 % of a million agents that start a
// burst in a step; warps without a row exit at once), 2 blocks of 512 threads per SM in the block-per-row kernel.
bool sbc_pair_cells_lanes(const DevParams &P, const DevState &S, const CellScratch &C, int pred_active) {
    return P.BC == 0 && cells_warp_rows(P, S, C) && !pred_active;
}
int sbc_pair_cells_blocks(const DevParams &P, const DevState &S, const CellScratch &C) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int per_sm = cells_warp_rows(P, S, C) ? 4 : 2;
    if (const char *e = getenv("SBC_ROWS_BLOCKS_PER_SM")) { const int v = atoi(e); if (v > 0 && v <= 8) per_sm = v; }   // tuning runs
    const size_t cap = (size_t)sms * per_sm;
    return (int)min(cap, max((size_t)16, (size_t)S.N / 64));
}

// the rows are S.active_list[step % 3] (kept by the steps themselves, or compacted by sbc_launch_compact_rows)
void sbc_launch_pair_cells(const DevParams &P, const DevState &S, CellScratch &C, const StepArgs &A, cudaStream_t st) {
    const CellGeom *geom = static_cast<const CellGeom *>(C.geom);
    const int blocks = sbc_pair_cells_blocks(P, S, C) + A.edge_blocks;   // (edge_blocks > 0 only with the lanes kernel)
#define SBC_CELLS_ARGS P, S, geom, C.cell_range, C.key_in, C.idx_out, A
    if (P.BC == 0) {
        if (cells_warp_rows(P, S, C)) {
            if (A.pred_active) sbc_launch_prio(pair_cells_warp_kernel<true>, blocks, CELLS_THREADS, 0, st, SBC_CELLS_ARGS);
            else if (A.ghost_ready && P.has_alg_zone) sbc_launch_prio(pair_cells_lanes_kernel<true, true, true>, blocks, CELLS_THREADS, 0, st, SBC_CELLS_ARGS);
            else if (A.ghost_ready) sbc_launch_prio(pair_cells_lanes_kernel<true, true, false>, blocks, CELLS_THREADS, 0, st, SBC_CELLS_ARGS);
            else if (P.has_alg_zone) sbc_launch_prio(pair_cells_lanes_kernel<true, false, true>, blocks, CELLS_THREADS, 0, st, SBC_CELLS_ARGS);
            else sbc_launch_prio(pair_cells_lanes_kernel<true, false, false>, blocks, CELLS_THREADS, 0, st, SBC_CELLS_ARGS);
        } else {
            sbc_launch_prio(pair_cells_kernel<true>, blocks, CELLS_BLOCK_THREADS, 0, st, SBC_CELLS_ARGS);
        }
    } else {
        sbc_launch_prio(pair_cells_kernel<false>, blocks, CELLS_BLOCK_THREADS, 0, st, SBC_CELLS_ARGS);
    }
#undef SBC_CELLS_ARGS
}
// (see sbc_pred_preload, predator.cu: nothing may load lazily once all-gathers wait inside kernels)
void sbc_cells_preload(void) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, pair_cells_warp_kernel<true>);
    cudaFuncGetAttributes(&fa, pair_cells_lanes_kernel<true, true, true>);
    cudaFuncGetAttributes(&fa, pair_cells_lanes_kernel<true, true, false>);
    cudaFuncGetAttributes(&fa, pair_cells_lanes_kernel<true, false, true>);
    cudaFuncGetAttributes(&fa, pair_cells_lanes_kernel<true, false, false>);
    cudaFuncGetAttributes(&fa, pair_cells_kernel<true>);
}
