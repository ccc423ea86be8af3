// pair_dense.cu - the dense all-pairs three-zone pass: every alive row active (the reference's
// state at s = 1), all N(N-1) directed pairs evaluated - the GPU counterpart of
// InteractionGlobal -> IntCalcPrey -> SFM_4Zone (/root/reference/src/agents_interact.cpp:144-157,
// 172-250; social_forces.cpp:25-52) and the kernel the FP32 roofline is quoted on.
//
// Design:
//   1. agents are put in Hilbert-curve order of their position (keys: curve_key_kernel, sort: CUB radix
//      sort, gather: gather_sorted_kernel), so that the 64 rows of a warp are neighbours in space;
//   2. grid = (row blocks of 256 rows) x (j splits) x (replicates); a block stages j-tiles of 256
//      agents in shared memory, every thread owns two rows (ILP 2) and sweeps the tile with one
//      broadcast LDS per j;
//   3. per pair the distance is ALWAYS evaluated; the three-zone arithmetic (precise periodic fold,
//      rounding-band check with FP64 re-decision, rsqrt, zone sums, counters) runs only behind a
//      branch `min(d2_row0, d2_row1) <= far2`, which is warp-uniformly false for most j once rows
//      are spatially sorted;
//   4. periodic boxes: rows and tile are expressed relative to the block's first row (minimum
//      image, computed once per j per block), so the inner loop needs no fold; blocks whose rows
//      are not compact enough for that to be safe (unsorted input, tiny boxes) take a generic
//      loop that folds every pair;
//   5. partial sums of the j splits are combined in a fixed order by reduce_partials_kernel
//      (deterministic results), which also scatters rows back to agent-id order and removes the
//      self pair from c_rep.
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "sbc_kernels.cuh"

namespace {

constexpr int DT = 128;          // threads per block
constexpr int ROWS = 2 * DT;     // rows per block (two per thread)
constexpr int TILE = 256;        // j per shared-memory tile

struct Acc9 {
    float rx, ry, ax, ay, tx, ty;
    int cr, ca, ct;
};

// index of cell (x, y) of a 65536 x 65536 grid along the Hilbert curve. Unlike the Morton order the
// Hilbert curve has no jumps: any run of consecutive agents is a compact blob, which is what the
// warp-uniform early-out and the block-relative periodic coordinates rely on.
__device__ __forceinline__ uint32_t hilbert_key(uint32_t x, uint32_t y) {
    uint32_t d = 0;
#pragma unroll
    for (uint32_t s = 32768u; s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = 65535u - x; y = 65535u - y; }
            const uint32_t t = x; x = y; y = t;
        }
    }
    return d;
}

// ordered-int image of a float for atomicMin/atomicMax
__device__ __forceinline__ int float_ordered(float f) {
    const int i = __float_as_int(f);
    return (i >= 0) ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float ordered_float(int i) {
    return __int_as_float((i >= 0) ? i : (i ^ 0x7fffffff));
}

__global__ void bbox_init_kernel(int *bbox) {
    bbox[0] = bbox[1] = 0x7fffffff;             // min x, min y
    bbox[2] = bbox[3] = (int)0x80000000;        // max x, max y
}
__global__ void bbox_kernel(const float4 *__restrict__ pv, int n, int *bbox) {
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
        atomicMin(bbox + 0, float_ordered(lo_x));
        atomicMin(bbox + 1, float_ordered(lo_y));
        atomicMax(bbox + 2, float_ordered(hi_x));
        atomicMax(bbox + 3, float_ordered(hi_y));
    }
}

// 32-bit Hilbert key of the position quantised to 16 bits per axis; dead agents sort last
__global__ void curve_key_kernel(const float4 *__restrict__ pv, int n, const int *__restrict__ bbox,
                                  float fixed_lo, float fixed_hi, int use_bbox,
                                  uint32_t *__restrict__ key, int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo_x = fixed_lo, lo_y = fixed_lo, hi_x = fixed_hi, hi_y = fixed_hi;
    if (use_bbox) {
        lo_x = ordered_float(bbox[0]); lo_y = ordered_float(bbox[1]);
        hi_x = ordered_float(bbox[2]); hi_y = ordered_float(bbox[3]);
    }
    const float4 p = pv[i];
    uint32_t k = 0xffffffffu;
    if (p.x == p.x) {
        const float sx = 65535.f / fmaxf(hi_x - lo_x, 1e-20f), sy = 65535.f / fmaxf(hi_y - lo_y, 1e-20f);
        const uint32_t qx = (uint32_t)fminf(fmaxf((p.x - lo_x) * sx, 0.f), 65535.f);
        const uint32_t qy = (uint32_t)fminf(fmaxf((p.y - lo_y) * sy, 0.f), 65535.f);
        k = hilbert_key(qx, qy);
        if (k == 0xffffffffu) k = 0xfffffffeu;
    }
    key[i] = k;
    idx[i] = i;
}

__global__ void iota_kernel(int *idx, int n_per_rep, size_t total) {
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g < total) idx[g] = (int)(g % n_per_rep);
}

// ensembles: every swarm is put in Hilbert order of its own agents. One sort over all of them with 64-bit keys
// (replicate in the high word, curve index in the low one); the box of a swarm is the domain (periodic box, tank) or,
// in an open domain, the bounding box of the whole ensemble (coarser cells, same locality).
__global__ void curve_key_rep_kernel(const float4 *__restrict__ pv, int n_per_rep, size_t total, const int *__restrict__ bbox,
                                     float fixed_lo, float fixed_hi, int use_bbox, unsigned long long *__restrict__ key,
                                     int *__restrict__ idx) {
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    float lo_x = fixed_lo, lo_y = fixed_lo, hi_x = fixed_hi, hi_y = fixed_hi;
    if (use_bbox) {
        lo_x = ordered_float(bbox[0]); lo_y = ordered_float(bbox[1]);
        hi_x = ordered_float(bbox[2]); hi_y = ordered_float(bbox[3]);
    }
    const float4 p = pv[g];
    uint32_t k = 0xffffffffu;   // dead agents sort behind the alive ones of their swarm
    if (p.x == p.x) {
        const float sx = 65535.f / fmaxf(hi_x - lo_x, 1e-20f), sy = 65535.f / fmaxf(hi_y - lo_y, 1e-20f);
        const uint32_t qx = (uint32_t)fminf(fmaxf((p.x - lo_x) * sx, 0.f), 65535.f);
        const uint32_t qy = (uint32_t)fminf(fmaxf((p.y - lo_y) * sy, 0.f), 65535.f);
        k = hilbert_key(qx, qy);
        if (k == 0xffffffffu) k = 0xfffffffeu;
    }
    const size_t rep = g / n_per_rep;
    key[g] = ((unsigned long long)rep << 32) | k;
    idx[g] = (int)(g - rep * n_per_rep);
}

__global__ void gather_sorted_kernel(const float4 *__restrict__ pv, const int *__restrict__ idx,
                                     int n_per_rep, size_t total, float4 *__restrict__ spv) {
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    const size_t rep = g / n_per_rep;
    spv[g] = pv[rep * n_per_rep + idx[g]];
}

// x_j - x_i folded to the minimum image as a two-float (hi + lo): hi is what sbc_delta_pbc returns,
// lo the exact rounding residual of its last subtraction (TwoSum) plus the low word of the box
// length. The dense kernel expresses rows and tile relative to a block reference this way, so that
// (hi_j - hi_i) + (lo_j - lo_i) is the pair displacement to ulp accuracy without a per-pair fold.
__device__ __forceinline__ float sbc_delta_pbc2(float xi, float xj, const DevParams &P, float &lo) {
    const float n = sbc_fold_count(xj - xi, P);
    const float a = __fmaf_rn(-fmaxf(n, 0.f), P.L, xj);
    const float b = __fmaf_rn(fminf(n, 0.f), P.L, xi);
    const float hi = __fsub_rn(a, b);
    const float bv = __fsub_rn(hi, a);            // TwoSum(a, -b)
    const float av = __fsub_rn(hi, bv);
    const float err = __fadd_rn(__fsub_rn(a, av), __fsub_rn(-b, bv));
    lo = __fmaf_rn(-n, P.L_lo, err);
    return hi;
}

// the three-zone arithmetic for one directed pair (row i <- j) given its displacement; branch-free
// except for the rare rounding-band case (then the zone is re-decided in FP64 from the stored
// coordinates). xi, yi: the row's stored position; pjp: the tile entry of j as stored.
template <bool HAS_ALG>
__device__ __forceinline__ void near_pair(Acc9 &A, unsigned &n_amb, float dx, float dy, float d2,
                                          float xi, float yi, const float4 *pjp, const DevParams &P) {
    // two compares per threshold: below the band / not above the band. Equal answers = the zone
    // decision is safe in FP32; different answers = the pair sits in the rounding band.
    const bool r_lo = d2 < P.band_lo[0], r_hi = d2 <= P.band_hi[0];
    const bool t_lo = d2 < P.band_lo[2], t_hi = d2 <= P.band_hi[2];   // NaN (dead) -> all false
    bool amb = (r_lo != r_hi) || (t_lo != t_hi);
    bool in_rep = r_hi;
    bool in_alg = false;
    if (HAS_ALG) {
        const bool a_lo = d2 < P.band_lo[1], a_hi = d2 <= P.band_hi[1];
        amb = amb || (a_lo != a_hi);
        in_alg = !in_rep && a_hi;
    }
    bool in_att = !in_rep && !in_alg && t_hi;
    if (amb) {
        const float4 pj = *pjp;
        const int z = sbc_zone_exact(xi, yi, pj.x, pj.y, P);
        in_rep = (z == 0);
        in_alg = (z == 1);
        in_att = (z == 2);
        n_amb++;
    }
    const float rinv = rsqrtf(fmaxf(d2, 1e-30f));  // d = 0 -> dx = dy = 0 -> u = 0 (agents_interact.cpp:199)
    const float wr = in_rep ? -rinv : 0.f;
    const float wt = in_att ? rinv : 0.f;
    A.rx = __fmaf_rn(dx, wr, A.rx);
    A.ry = __fmaf_rn(dy, wr, A.ry);
    A.tx = __fmaf_rn(dx, wt, A.tx);
    A.ty = __fmaf_rn(dy, wt, A.ty);
    A.cr += in_rep ? 1 : 0;
    A.ct += in_att ? 1 : 0;
    if (HAS_ALG) {
        const float4 pj = *pjp;
        A.ax += in_alg ? pj.z : 0.f;
        A.ay += in_alg ? pj.w : 0.f;
        A.ca += in_alg ? 1 : 0;
    }
}

template <bool PERIODIC, bool HAS_ALG>
__global__ void __launch_bounds__(DT, 6)
pair_dense_kernel(const __grid_constant__ DevParams P, int N, int R, const float4 *__restrict__ spv_all,
                  int nsplit, float *__restrict__ part_f, int *__restrict__ part_i,
                  unsigned long long *__restrict__ stats) {
    __shared__ float4 tA[TILE];                 // far-test tile: block-relative {x', y', w = -(x'^2+y'^2)(1-8u), 0}
    __shared__ __align__(16) float2 tL[TILE];   // periodic: lo parts of the block-relative tile positions
    __shared__ float4 tB[TILE];                 // tile agents as stored
    __shared__ int s_ref;
    const int rep = blockIdx.z, sp = blockIdx.y, rb = blockIdx.x;
    const float4 *spv = spv_all + (size_t)rep * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = rb * ROWS + warp * 64 + lane, row1 = row0 + 32;
    const float qnan = __int_as_float(0x7fc00000);
    const float4 nan4 = make_float4(qnan, qnan, 0.f, 0.f);
    const float4 p0 = (row0 < N) ? spv[row0] : nan4;
    const float4 p1 = (row1 < N) ? spv[row1] : nan4;
    // block reference point = the block's first alive row (dead agents sort last, but small or
    // replicated swarms are processed in agent-id order)
    if (threadIdx.x == 0) s_ref = 0x7fffffff;
    __syncthreads();
    if (p0.x == p0.x) atomicMin(&s_ref, row0);
    if (p1.x == p1.x) atomicMin(&s_ref, row1);
    __syncthreads();
    const bool any_alive = (s_ref != 0x7fffffff);
    const float4 c = any_alive ? spv[s_ref] : nan4;
    Acc9 A0 = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0, 0, 0}, A1 = A0;
    unsigned n_amb = 0;
    // block-relative row coordinates: two-floats (hi, lo) of the minimum image in a periodic box,
    // plain differences to the reference point otherwise
    float x0 = p0.x - c.x, y0 = p0.y - c.y, x1 = p1.x - c.x, y1 = p1.y - c.y;
    float l0x = 0.f, l0y = 0.f, l1x = 0.f, l1y = 0.f;
    int generic = 0;
    if (PERIODIC) {
        x0 = sbc_delta_pbc2(c.x, p0.x, P, l0x); y0 = sbc_delta_pbc2(c.y, p0.y, P, l0y);
        x1 = sbc_delta_pbc2(c.x, p1.x, P, l1x); y1 = sbc_delta_pbc2(c.y, p1.y, P, l1y);
        // block-relative coordinates are only safe while |x'_i| < L/2 - att (see header)
        const float m = fmaxf(fmaxf(fabsf(x0), fabsf(y0)), fmaxf(fabsf(x1), fabsf(y1)));
        generic = __syncthreads_or(m > P.compact_lim);   // NaN rows compare false
    }
    // Far test in expanded form: |p_j - p_i|^2 <= far2  <=>  (|p_i|^2 - far2) - 2 p_i.p_j <= -|p_j|^2.
    // Two FFMAs per pair; the FP32 rounding of the three terms (<= 4u(2|p_i|^2 + |p_j|^2 + far2)) is
    // covered by shifting both sides: the test may pass a pair that is not near, never the reverse.
    const float u8 = 8.f * 5.9604645e-8f;
    const float n0 = __fmaf_rn(y0, y0, x0 * x0), n1 = __fmaf_rn(y1, y1, x1 * x1);
    const float c0 = (n0 - P.far2) - u8 * (n0 + P.far2), c1 = (n1 - P.far2) - u8 * (n1 + P.far2);
    const float m2x0 = -2.f * x0, m2y0 = -2.f * y0, m2x1 = -2.f * x1, m2y1 = -2.f * y1;
    const float far2 = P.far2;
    (void)far2;
    if (any_alive) {
        // tiles are dealt to the splits round-robin: a split's tiles are spread over the whole
        // curve, so every block sees about the same share of near pairs
        for (int j0 = sp * TILE; j0 < N; j0 += nsplit * TILE) {
            __syncthreads();
            for (int t = threadIdx.x; t < TILE; t += DT) {
                const int j = j0 + t;
                const float4 pj = (j < N) ? spv[j] : nan4;
                tB[t] = pj;
                float hx = pj.x - c.x, hy = pj.y - c.y;
                if (PERIODIC) {
                    float lx, ly;
                    hx = sbc_delta_pbc2(c.x, pj.x, P, lx);
                    hy = sbc_delta_pbc2(c.y, pj.y, P, ly);
                    tL[t] = make_float2(lx, ly);
                }
                tA[t] = make_float4(hx, hy, -__fmaf_rn(hy, hy, hx * hx) * (1.f - 8.f * 5.9604645e-8f), 0.f);
            }
            __syncthreads();
            if (!generic) {
#pragma unroll 2
                for (int jj = 0; jj < TILE; jj += 2) {
                    const float4 qa = tA[jj], qb = tA[jj + 1];
                    const float ea0 = __fmaf_rn(m2x0, qa.x, __fmaf_rn(m2y0, qa.y, c0));
                    const float ea1 = __fmaf_rn(m2x1, qa.x, __fmaf_rn(m2y1, qa.y, c1));
                    const float eb0 = __fmaf_rn(m2x0, qb.x, __fmaf_rn(m2y0, qb.y, c0));
                    const float eb1 = __fmaf_rn(m2x1, qb.x, __fmaf_rn(m2y1, qb.y, c1));
                    const float ta = fminf(ea0, ea1) - qa.z, tb = fminf(eb0, eb1) - qb.z;
                    if (fminf(ta, tb) <= 0.f) {  // NaN (dead or padding) never passes
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            if ((u ? tb : ta) <= 0.f) {
                                const float4 q = u ? qb : qa;
                                const float4 *pjp = &tB[jj + u];
                                float dx0, dy0, dx1, dy1;
                                if (PERIODIC) {
                                    const float2 l = tL[jj + u];
                                    dx0 = (q.x - x0) + (l.x - l0x); dy0 = (q.y - y0) + (l.y - l0y);
                                    dx1 = (q.x - x1) + (l.x - l1x); dy1 = (q.y - y1) + (l.y - l1y);
                                } else {
                                    const float4 pj = *pjp;   // stored coordinates: one rounding per difference
                                    dx0 = pj.x - p0.x; dy0 = pj.y - p0.y;
                                    dx1 = pj.x - p1.x; dy1 = pj.y - p1.y;
                                }
                                near_pair<HAS_ALG>(A0, n_amb, dx0, dy0, __fmaf_rn(dy0, dy0, __fmul_rn(dx0, dx0)), p0.x, p0.y, pjp, P);
                                near_pair<HAS_ALG>(A1, n_amb, dx1, dy1, __fmaf_rn(dy1, dy1, __fmul_rn(dx1, dx1)), p1.x, p1.y, pjp, P);
                            }
                        }
                    }
                }
            } else {
                // rows of this block are too spread out for block-relative coordinates (unsorted
                // input or a box barely larger than 2 att_range): fold every pair, rough first
                const int n = min(TILE, N - j0);
                for (int jj = 0; jj < n; jj++) {
                    const float4 *pjp = &tB[jj];
                    const float4 pj = *pjp;
                    const float rx0 = sbc_delta_pbc_rough(p0.x, pj.x, P), ry0 = sbc_delta_pbc_rough(p0.y, pj.y, P);
                    const float rx1 = sbc_delta_pbc_rough(p1.x, pj.x, P), ry1 = sbc_delta_pbc_rough(p1.y, pj.y, P);
                    const float d0 = __fmaf_rn(ry0, ry0, rx0 * rx0), d1 = __fmaf_rn(ry1, ry1, rx1 * rx1);
                    if (fminf(d0, d1) <= far2) {
                        const float dx0 = sbc_delta_pbc(p0.x, pj.x, P), dy0 = sbc_delta_pbc(p0.y, pj.y, P);
                        const float dx1 = sbc_delta_pbc(p1.x, pj.x, P), dy1 = sbc_delta_pbc(p1.y, pj.y, P);
                        near_pair<HAS_ALG>(A0, n_amb, dx0, dy0, __fmaf_rn(dy0, dy0, __fmul_rn(dx0, dx0)), p0.x, p0.y, pjp, P);
                        near_pair<HAS_ALG>(A1, n_amb, dx1, dy1, __fmaf_rn(dy1, dy1, __fmul_rn(dx1, dx1)), p1.x, p1.y, pjp, P);
                    }
                }
            }
        }
    }
    // partial sums: layout [k][rep][split][row] so that a warp's stores are contiguous
    const size_t plane = (size_t)R * nsplit * N;
    const size_t base = ((size_t)rep * nsplit + sp) * N;
    if (row0 < N) {
        float *f = part_f + base + row0;
        f[0] = A0.rx; f[plane] = A0.ry; f[2 * plane] = A0.ax; f[3 * plane] = A0.ay;
        f[4 * plane] = A0.tx; f[5 * plane] = A0.ty;
        int *q = part_i + base + row0;
        q[0] = A0.cr; q[plane] = A0.ca; q[2 * plane] = A0.ct;
    }
    if (row1 < N) {
        float *f = part_f + base + row1;
        f[0] = A1.rx; f[plane] = A1.ry; f[2 * plane] = A1.ax; f[3 * plane] = A1.ay;
        f[4 * plane] = A1.tx; f[5 * plane] = A1.ty;
        int *q = part_i + base + row1;
        q[0] = A1.cr; q[plane] = A1.ca; q[2 * plane] = A1.ct;
    }
    n_amb = __reduce_add_sync(0xffffffffu, n_amb);
    if (lane == 0 && n_amb) atomicAdd(stats + 4, (unsigned long long)n_amb);
}

// sum the j-split partials in split order, drop the self pair, scatter to agent-id order.
// One thread per (plane, row): planes 0..5 are the zone sums, 6..8 the zone counts.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(DevState S, const float4 *__restrict__ spv, const int *__restrict__ idx,
                       int nsplit, const float *__restrict__ part_f, const int *__restrict__ part_i) {
    const int N = S.N, R = S.R;
    const size_t total = (size_t)N * R;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= total * 9) return;
    const int k = (int)(t / total);
    const size_t rr = t - (size_t)k * total;
    const int rep = (int)(rr / N), row = (int)(rr - (size_t)rep * N);
    const size_t plane = (size_t)R * nsplit * N;
    const bool alive = (spv[rr].x == spv[rr].x);
    const size_t g = (size_t)rep * N + idx[rr];
    const size_t o = (size_t)rep * nsplit * N + row;
    if (k < 6) {
        const float *src = part_f + (size_t)k * plane + o;
        float f = 0.f;
        for (int sp = 0; sp < nsplit; sp++) f += src[(size_t)sp * N];
        S.acc[g * 8 + k] = alive ? f : 0.f;
    } else {
        const int *src = part_i + (size_t)(k - 6) * plane + o;
        int c = 0;
        for (int sp = 0; sp < nsplit; sp++) c += src[(size_t)sp * N];
        if (k == 6) c -= 1;  // the self pair (d = 0, zero force) was counted as repulsion
        S.cnt[g * 4 + (k - 6)] = alive ? c : 0;
    }
}

}  // namespace

// ---- host side --------------------------------------------------------------------------------
static void dense_plan(int N, int R, int &nsplit, int &jchunk) {
    const int row_blocks = (N + ROWS - 1) / ROWS;
    int waves = 8;
    if (const char *e = getenv("SBC_DENSE_WAVES")) waves = atoi(e) > 0 ? atoi(e) : 8;
    const int target = 148 * 6 * waves;  // four waves at 6 resident blocks per SM: block run times differ
                                     // by the share of near pairs, the block scheduler evens them out
    int want = (target + row_blocks * R - 1) / (row_blocks * R);
    const int max_split = (N + TILE - 1) / TILE;
    if (want < 1) want = 1;
    if (want > max_split) want = max_split;
    jchunk = (N + want - 1) / want;
    jchunk = ((jchunk + TILE - 1) / TILE) * TILE;
    nsplit = (N + jchunk - 1) / jchunk;
}

cudaError_t sbc_dense_scratch_alloc(DenseScratch &D, int N, int R) {
    const size_t total = (size_t)N * R;
    int nsplit, jchunk;
    dense_plan(N, R, nsplit, jchunk);
    D.nsplit = nsplit;
    D.jchunk = jchunk;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&D.key_in, total * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.key_out, total * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.idx_in, total * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.idx_out, total * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.spv, total * sizeof(float4))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.part_f, total * nsplit * 6 * sizeof(float))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.part_i, total * nsplit * 3 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&D.bbox, 4 * sizeof(int))) != cudaSuccess) return e;
    D.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, D.cub_bytes, D.key_in, D.key_out, D.idx_in, D.idx_out, N);
    if (R > 1) {   // ensembles: one sort of (replicate, curve index) keys over all swarms
        if ((e = cudaMalloc((void **)&D.key64_in, total * sizeof(unsigned long long))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&D.key64_out, total * sizeof(unsigned long long))) != cudaSuccess) return e;
        size_t b64 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b64, D.key64_in, D.key64_out, D.idx_in, D.idx_out, (int)total);
        if (b64 > D.cub_bytes) D.cub_bytes = b64;
    }
    if ((e = cudaMalloc(&D.cub_tmp, D.cub_bytes ? D.cub_bytes : 16)) != cudaSuccess) return e;
    D.sorted_valid = false;
    return cudaSuccess;
}

void sbc_dense_scratch_free(DenseScratch &D) {
    cudaFree(D.key_in); cudaFree(D.key_out); cudaFree(D.idx_in); cudaFree(D.idx_out);
    cudaFree(D.spv); cudaFree(D.part_f); cudaFree(D.part_i); cudaFree(D.bbox); cudaFree(D.cub_tmp);
    cudaFree(D.key64_in); cudaFree(D.key64_out);
    D = DenseScratch();
}

// (re)build the Morton order of the current positions; R > 1 keeps agent-id order
void sbc_launch_dense_sort(const DevParams &P, const DevState &S, DenseScratch &D, cudaStream_t st) {
    const int N = S.N;
    const size_t total = (size_t)N * S.R;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (S.R == 1 && N > ROWS) {
        float lo = 0.f, hi = P.L;
        int use_bbox = 0;
        if (P.BC == 5 || P.BC == 6) { lo = -P.L; hi = P.L; }
        if (P.BC == -1) {
            use_bbox = 1;
            bbox_init_kernel<<<1, 1, 0, st>>>(D.bbox);
            bbox_kernel<<<148, 256, 0, st>>>(S.pv, N, D.bbox);
        }
        curve_key_kernel<<<blocks, 256, 0, st>>>(S.pv, N, D.bbox, lo, hi, use_bbox, D.key_in, D.idx_in);
        cub::DeviceRadixSort::SortPairs(D.cub_tmp, D.cub_bytes, D.key_in, D.key_out, D.idx_in, D.idx_out,
                                        N, 0, 32, st);
    } else if (S.R > 1 && N > 64 && D.key64_in) {
        float lo = 0.f, hi = P.L;
        int use_bbox = 0;
        if (P.BC == 5 || P.BC == 6) { lo = -P.L; hi = P.L; }
        if (P.BC == -1) {
            use_bbox = 1;
            bbox_init_kernel<<<1, 1, 0, st>>>(D.bbox);
            bbox_kernel<<<148, 256, 0, st>>>(S.pv, (int)total, D.bbox);
        }
        curve_key_rep_kernel<<<blocks, 256, 0, st>>>(S.pv, N, total, D.bbox, lo, hi, use_bbox, D.key64_in, D.idx_in);
        int rep_bits = 1;
        while ((1ll << rep_bits) < (long long)S.R) rep_bits++;
        cub::DeviceRadixSort::SortPairs(D.cub_tmp, D.cub_bytes, D.key64_in, D.key64_out, D.idx_in, D.idx_out, (int)total, 0,
                                        32 + rep_bits, st);
    } else {
        iota_kernel<<<blocks, 256, 0, st>>>(D.idx_out, N, total);
    }
    D.sorted_valid = true;
}

// the pass itself: gather in the stored order, all pairs, combine
void sbc_launch_pair_dense(const DevParams &P, const DevState &S, DenseScratch &D, cudaStream_t st) {
    const int N = S.N;
    const size_t total = (size_t)N * S.R;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    gather_sorted_kernel<<<blocks, 256, 0, st>>>(S.pv, D.idx_out, N, total, D.spv);
    dim3 grid((N + ROWS - 1) / ROWS, D.nsplit, S.R);
#define SBC_DENSE_LAUNCH(PER, ALG) pair_dense_kernel<PER, ALG><<<grid, DT, 0, st>>>( \
        P, N, S.R, D.spv, D.nsplit, D.part_f, D.part_i, S.stats)
    if (P.BC == 0) { if (P.has_alg_zone) SBC_DENSE_LAUNCH(true, true); else SBC_DENSE_LAUNCH(true, false); }
    else { if (P.has_alg_zone) SBC_DENSE_LAUNCH(false, true); else SBC_DENSE_LAUNCH(false, false); }
#undef SBC_DENSE_LAUNCH
    reduce_partials_kernel<<<(unsigned)((total * 9 + 255) / 256), 256, 0, st>>>(S, D.spv, D.idx_out, D.nsplit, D.part_f, D.part_i);
}
