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
                                   uint32_t dead_key, int *__restrict__ cell_start, int *__restrict__ cell_end) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t k = key[r];
    if (k == dead_key) return;
    if (r == 0 || key[r - 1] != k) cell_start[k] = r;
    if (r == n - 1 || key[r + 1] != k) cell_end[k] = r + 1;
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
                                                const int *__restrict__ cell_start, const int *__restrict__ cell_end,
                                                const uint32_t *__restrict__ build_key, const int *__restrict__ sidx, int i,
                                                int t /* thread in group */, int *pref /* [10] shared, per group */,
                                                int *first /* [9] shared, per group */, RowAcc &A, unsigned &n_amb,
                                                unsigned long long &n_cand) {
    const float4 pi = S.pv[i];
    // the row's cell at build time (agents alive now were alive then: nobody is born between builds)
    const int ck = (int)build_key[i];
    const int cy = ck / g.ncx, cx = ck - cy * g.ncx;
    if (t < 9) {
        int yy = cy + t / 3 - 1, xx = cx + t % 3 - 1;
        bool ok = true;
        if (PERIODIC) {
            yy = (yy + g.ncy) % g.ncy;
            xx = (xx + g.ncx) % g.ncx;
        } else {
            ok = yy >= 0 && yy < g.ncy && xx >= 0 && xx < g.ncx;
        }
        int r0 = 0, r1 = 0;
        if (ok) {
            const int c = yy * g.ncx + xx;
            r0 = cell_start[c];
            r1 = cell_end[c];
        }
        first[t] = r0;
        pref[t + 1] = r1 - r0;
    }
    if (TPR == 32) __syncwarp(); else __syncthreads();
    if (t == 0) {
        int run = 0;
        pref[0] = 0;
        for (int c = 0; c < 9; c++) { run += pref[c + 1]; pref[c + 1] = run; }
        n_cand += (unsigned long long)run;
    }
    if (TPR == 32) __syncwarp(); else __syncthreads();
    const int total = pref[9];
    row_acc_clear(A);
    for (int q = t; q < total; q += TPR) {
        int c = 0;
#pragma unroll
        for (int k = 1; k < 9; k++) c += (q >= pref[k]) ? 1 : 0;
        const int j = sidx[first[c] + (q - pref[c])];
        if (j == i) continue;
        const float4 pj = S.pv[j];   // current position (NaN if the agent died since the build)
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
__global__ void __launch_bounds__(CELLS_BLOCK_THREADS)
pair_cells_kernel(const __grid_constant__ DevParams P, DevState S, const CellGeom *__restrict__ geom,
                  const int *__restrict__ cell_start, const int *__restrict__ cell_end,
                  const uint32_t *__restrict__ build_key, const int *__restrict__ sidx, StepArgs SA) {
    const uint32_t step = SA.step_dev ? *SA.step_dev : SA.step;
    const int par = (int)(step & 1u);
    const int *__restrict__ list = S.active_list[par];
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
    for (int k = blockIdx.x; k < n_rows; k += gridDim.x) {
        const int i = list[k];  // single swarm on this path: slot index == global index
        RowAcc A;
        cells_row_sweep<PERIODIC, CELLS_BLOCK_THREADS>(P, S, g, cell_start, cell_end, build_key, sidx, i, threadIdx.x, pref, first, A,
                                                 n_amb, n_cand);
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
                sbc_finish_row(P, S, (size_t)i, A, step, SA);
            }
        }
        __syncthreads();   // pref / first are rewritten by the next row
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, (unsigned long long)n_amb);
    if (threadIdx.x == 0 && n_cand) atomicAdd(S.stats + 3, n_cand);  // candidates examined on this path
    sbc_step_close(S, SA, step);
}

// sparse neighbourhoods (a few dozen candidates): one WARP per active row, reduction by shuffles only
template <bool PERIODIC>
__global__ void __launch_bounds__(CELLS_THREADS)
pair_cells_warp_kernel(const __grid_constant__ DevParams P, DevState S, const CellGeom *__restrict__ geom,
                       const int *__restrict__ cell_start, const int *__restrict__ cell_end,
                       const uint32_t *__restrict__ build_key, const int *__restrict__ sidx, StepArgs SA) {
    const uint32_t step = SA.step_dev ? *SA.step_dev : SA.step;
    const int par = (int)(step & 1u);
    const int *__restrict__ list = S.active_list[par];
    constexpr int WPB = CELLS_THREADS / 32;
    __shared__ int pref[WPB][10], first[WPB][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_rows = S.active_count[par];
    const CellGeom g = *geom;
    unsigned n_amb = 0;
    unsigned long long n_cand = 0;
    for (int k = blockIdx.x * WPB + warp; k < n_rows; k += gridDim.x * WPB) {
        const int i = list[k];
        RowAcc A;
        cells_row_sweep<PERIODIC, 32>(P, S, g, cell_start, cell_end, build_key, sidx, i, lane, pref[warp], first[warp], A, n_amb,
                                      n_cand);
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
        if (SA.finish && lane == 0) sbc_finish_row(P, S, (size_t)i, A, step, SA);
        __syncwarp();
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(S.stats + 4, (unsigned long long)n_amb);
    if (lane == 0 && n_cand) atomicAdd(S.stats + 3, n_cand);
    sbc_step_close(S, SA, step);
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
    } else {
        C.ncells = (size_t)C.max_per_axis * C.max_per_axis;
    }
    if ((e = cudaMalloc((void **)&C.key_in, (size_t)N * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.key_out, (size_t)N * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.idx_in, (size_t)N * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.idx_out, (size_t)N * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.cell_start, C.ncells * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.cell_end, C.ncells * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&C.bbox, 4 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&C.geom, sizeof(CellGeom))) != cudaSuccess) return e;
    if (P.BC == 0) cudaMemcpy(C.geom, &hg, sizeof(hg), cudaMemcpyHostToDevice);
    C.key_bits = 1;
    while (((size_t)1 << C.key_bits) <= C.ncells) C.key_bits++;
    C.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, C.cub_bytes, C.key_in, C.key_out, C.idx_in, C.idx_out, N, 0, C.key_bits);
    if ((e = cudaMalloc(&C.cub_tmp, C.cub_bytes ? C.cub_bytes : 16)) != cudaSuccess) return e;
    return cudaSuccess;
}

void sbc_cell_scratch_free(CellScratch &C) {
    cudaFree(C.key_in); cudaFree(C.key_out); cudaFree(C.idx_in); cudaFree(C.idx_out);
    cudaFree(C.cell_start); cudaFree(C.cell_end); cudaFree(C.bbox); cudaFree(C.geom); cudaFree(C.cub_tmp);
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
    cell_key_kernel<<<blocks, 256, 0, st>>>(S.pv, N, geom, dead_key, C.key_in, C.idx_in);
    cub::DeviceRadixSort::SortPairs(C.cub_tmp, C.cub_bytes, C.key_in, C.key_out, C.idx_in, C.idx_out, N, 0,
                                    C.key_bits, st);
    cudaMemsetAsync(C.cell_start, 0, C.ncells * sizeof(int), st);
    cudaMemsetAsync(C.cell_end, 0, C.ncells * sizeof(int), st);
    cell_bounds_kernel<<<blocks, 256, 0, st>>>(C.key_out, C.idx_out, N, dead_key, C.cell_start, C.cell_end);
}

// mean number of candidates a row meets in its nine cells decides between a warp and a block per row
static bool cells_warp_rows(const DevParams &P, const DevState &S, const CellScratch &C) {
    const double per_row = 9.0 * (double)S.N / (double)(C.ncells > 0 ? C.ncells : 1);
    return P.BC == 0 && per_row < 512.0;
}
int sbc_pair_cells_blocks(const DevParams &P, const DevState &S, const CellScratch &C) {
    return (int)min((size_t)148 * 16, max((size_t)16, (size_t)S.N / 64));
}

// the rows are S.active_list[step & 1] (kept by the steps themselves, or compacted by sbc_launch_compact_rows)
void sbc_launch_pair_cells(const DevParams &P, const DevState &S, CellScratch &C, const StepArgs &A, cudaStream_t st) {
    const CellGeom *geom = static_cast<const CellGeom *>(C.geom);
    const int blocks = sbc_pair_cells_blocks(P, S, C);
#define SBC_CELLS_ARGS P, S, geom, C.cell_start, C.cell_end, C.key_in, C.idx_out, A
    if (P.BC == 0) {
        if (cells_warp_rows(P, S, C)) pair_cells_warp_kernel<true><<<blocks, CELLS_THREADS, 0, st>>>(SBC_CELLS_ARGS);
        else pair_cells_kernel<true><<<blocks, CELLS_BLOCK_THREADS, 0, st>>>(SBC_CELLS_ARGS);
    } else {
        pair_cells_kernel<false><<<blocks, CELLS_BLOCK_THREADS, 0, st>>>(SBC_CELLS_ARGS);
    }
#undef SBC_CELLS_ARGS
}
