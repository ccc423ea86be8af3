// observables.cu - the reference's per-output-step swarm observables on the device:
//   Out_swarm            /root/reference/src/swarmdyn.cpp:237-344   (16 columns of /swarm)
//   Out_swarm_fishNet    /root/reference/src/swarmdyn.cpp:358-410   (4 columns of /swarm_fishNet)
//   basic_particle_averages :486-510, get_elongation / AreaConvexHull agents_operation.cpp:130-177
// for every replicate of a handle, from the FP32 state as it stands, in FP64 where the reference's value depends on
// a selection (nearest neighbour, most distant pair, extreme projections), so that those are the SAME agents.
//
// Reference quirks reproduced (SURVEY f-1): agents are the alive ones in id order (split_dead); the nearest-neighbour
// loop skips the agent just before i (:275 runs j < i-1) and starts from N*alg_range (:273); IID sums j > i only and is
// divided by N(N-1) (:296); ND is 0/0 = NaN whenever an agent has an empty neighbour list (:270) - on the device always
// (the lists of the last step are not kept); max_IID_vec starts as (1,1) (:247). Corrected values travel in `extra`.
//
// Passes (all per-block partials are combined in a fixed order: results do not depend on scheduling):
//   select   ordered compaction of the alive slots (CUB DeviceSelect), replicate offsets by binary search
//   sums     per chunk of 1024 agents: sums of x, v, u, |v|, v^2; support points in 16 directions (hull filter)
//   pairs    one thread per agent i over all j of its replicate (tiles in shared memory): FP32 distance for every
//            ordered pair; a pair that comes within 2e-6 (relative) of the thread's running minimum / maximum is
//            re-evaluated in FP64 in the reference's operation order (mathtools.h:54-83) - minima, the maximum and the
//            pair that attains it (first in (i, j) order, as the reference's strict `>` keeps it) are exact
//   second   per agent: angular-momentum term about the centre, projections on avg_u / max_IID_vec and their normals,
//            hull filter: agents not strictly inside the polygon of the 16 support points are handed to the host
//   host     monotone-chain hull of the few surviving points (libsbc's host side, double precision)
#include "sbc_kernels.cuh"
#include "../../include/sbc.h"
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <algorithm>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

namespace {
constexpr int OT = 256;          // threads per block, every kernel here
constexpr int CHUNK = 1024;      // agents per block of the O(N) passes
constexpr int NDIR = 16;         // support directions of the hull filter
constexpr int NSUM = 8;          // x y vx vy ux uy |v| v^2
constexpr int NSEC = 9;          // L term, min/max of 4 projections

struct MaxRec { double d2; int i, j, pad; };
struct HullPt { int rep; float x, y; };
// what the finalize kernels leave per replicate
struct RepRes {
    double n, avg_x[2], avg_v[2], avg_u[2], avg_s, avg_v2;
    double nnd_q, nnd_true, iid_sum, max_d2, max_vec[2];
    double dir_u[2], dir_m[2];
    float2 poly[NDIR];
    int npoly, max_i, max_j, pad;
};

struct AliveAt {
    const float4 *pv;
    __device__ __forceinline__ bool operator()(const int &g) const { const float x = pv[g].x; return x == x; }
};

__device__ __forceinline__ double block_sum(double v, double *sh) {   // fixed order: xor tree in the warp, warps by index
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < OT / 32; w++) t += sh[w];
    return t;
}
__device__ __forceinline__ double block_min(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = sh[0];
    for (int w = 1; w < OT / 32; w++) t = fmin(t, sh[w]);
    return t;
}
__device__ __forceinline__ bool max_better(double d2, int i, int j, double D2, int I, int J) {
    return d2 > D2 || (d2 == D2 && (i < I || (i == I && j < J)));
}

// CalcDistVec + the squared vec_length in the reference's order (mathtools.h:54-83), arguments promoted exactly
__device__ __forceinline__ void dist_vec_dd(float xi, float yi, float xj, float yj, const DevParams &P, double r[2]) {
    r[0] = __dsub_rn((double)xj, (double)xi);
    r[1] = __dsub_rn((double)yj, (double)yi);
    if (P.BC == 0) { r[0] = sbc_min_image_dd(r[0], P.dL); r[1] = sbc_min_image_dd(r[1], P.dL); }
}
__device__ __noinline__ double exact_d2(float xi, float yi, float xj, float yj, const DevParams &P) {
    double r[2];
    dist_vec_dd(xi, yi, xj, yj, P, r);
    return __dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1]));
}

__global__ void obs_offsets_kernel(const int *__restrict__ cidx, const int *__restrict__ n_sel, int N, int R, int *off) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > R) return;
    const long long want = (long long)r * N;
    int lo = 0, hi = *n_sel;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (cidx[m] < want) lo = m + 1; else hi = m; }
    off[r] = lo;
}

__global__ void __launch_bounds__(OT)
obs_sums_kernel(DevState S, const int *__restrict__ cidx, const int *__restrict__ off, int chunks,
                double *part, unsigned long long *ext) {
    __shared__ double sh[OT / 32];
    __shared__ unsigned long long shx[NDIR];
    const int rep = blockIdx.y, c = blockIdx.x;
    const int o0 = off[rep], n = off[rep + 1] - o0;
    if (c * CHUNK >= n && c > 0) return;
    if (threadIdx.x < NDIR) shx[threadIdx.x] = 0ull;
    __syncthreads();
    double a[NSUM] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int end = min(n, (c + 1) * CHUNK);
    for (int k = c * CHUNK + threadIdx.x; k < end; k += OT) {
        const int g = cidx[o0 + k];
        const float4 p = S.pv[g];
        double su, cu;
        sincos((double)S.hv[g].x, &su, &cu);                   // particle::u = (cos phi, sin phi)
        const double vx = p.z, vy = p.w, v2 = __dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy));
        a[0] += p.x; a[1] += p.y; a[2] += vx; a[3] += vy; a[4] += cu; a[5] += su; a[6] += sqrt(v2); a[7] += v2;
#pragma unroll
        for (int d = 0; d < NDIR; d++) {                       // support point of direction d (FP32 is enough for a filter)
            float sd, cd;
            sincospif(2.f * d / NDIR, &sd, &cd);
            const float pr = p.x * cd + p.y * sd;
            unsigned u = __float_as_uint(pr);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);    // order-preserving map to unsigned
            atomicMax(&shx[d], ((unsigned long long)u << 32) | (unsigned)k);
        }
    }
    for (int q = 0; q < NSUM; q++) {
        const double t = block_sum(a[q], sh);
        if (threadIdx.x == 0) part[((size_t)rep * chunks + c) * NSUM + q] = t;
    }
    __syncthreads();
    if (threadIdx.x < NDIR && n > 0) atomicMax(&ext[(size_t)rep * NDIR + threadIdx.x], shx[threadIdx.x]);
}

__global__ void obs_final1_kernel(DevState S, const int *__restrict__ cidx, const int *__restrict__ off, int chunks,
                                  const double *__restrict__ part, const unsigned long long *__restrict__ ext, RepRes *res) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= S.R) return;
    RepRes &q = res[rep];
    const int o0 = off[rep], n = off[rep + 1] - o0;
    double t[NSUM] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int used = max(1, (n + CHUNK - 1) / CHUNK);
    for (int c = 0; c < used; c++)
        for (int k = 0; k < NSUM; k++) t[k] += part[((size_t)rep * chunks + c) * NSUM + k];
    q.n = n;
    q.avg_x[0] = t[0] / n; q.avg_x[1] = t[1] / n; q.avg_v[0] = t[2] / n; q.avg_v[1] = t[3] / n;
    q.avg_u[0] = t[4] / n; q.avg_u[1] = t[5] / n; q.avg_s = t[6] / n; q.avg_v2 = t[7] / n;
    // polygon of the distinct support points, counter-clockwise because the directions are
    int np = 0, last = -1, first = -1;
    for (int d = 0; d < NDIR && n > 0; d++) {
        const int k = (int)(ext[(size_t)rep * NDIR + d] & 0xffffffffull);
        if (k == last || (d == NDIR - 1 && k == first)) continue;
        const float4 p = S.pv[cidx[o0 + k]];
        bool dup = false;                                       // the same agent can serve two non-adjacent directions only
        for (int m = 0; m < np; m++) dup = dup || (q.poly[m].x == p.x && q.poly[m].y == p.y);   // in degenerate sets
        if (dup) continue;
        q.poly[np++] = make_float2(p.x, p.y);
        last = k;
        if (first < 0) first = k;
    }
    q.npoly = np;
}

// ---- the O(N^2) pass -----------------------------------------------------------------------------------------------
struct RowTrack {
    double minq, maxd;      // exact squared distances
    float min_thr, max_thr; // FP32 squared distances beyond these cannot change minq / maxd
    int mj;
    float iid;              // sum of FP32 distances of the current tile
};
// MODE 0: every j of the tile is below k-1 (minimum only), 1: every j is above k (minimum, sum, maximum), 2: mixed
template <int MODE>
__device__ __forceinline__ void sweep_tile(const float2 *__restrict__ tile, int j0, int cnt, int k, float xi, float yi,
                                           const DevParams &P, RowTrack &T) {
#pragma unroll 4
    for (int t = 0; t < cnt; t++) {
        const float2 q = tile[t];
        float dx, dy;
        if (P.BC == 0) { dx = sbc_delta_pbc(xi, q.x, P); dy = sbc_delta_pbc(yi, q.y, P); }
        else { dx = q.x - xi; dy = q.y - yi; }
        const float d2 = fmaf(dy, dy, dx * dx);
        const int j = j0 + t;
        const bool upper = MODE == 1 || (MODE == 2 && j > k);
        const bool lower = MODE == 0 || (MODE == 2 && j < k - 1);
        if (upper) T.iid += sqrtf(d2);
        if ((upper || lower) && d2 <= T.min_thr) {
            const double e = exact_d2(xi, yi, q.x, q.y, P);
            if (e < T.minq) { T.minq = e; T.min_thr = __double2float_ru(e * (1.0 + 2e-6)); }
        }
        if (upper && d2 >= T.max_thr) {
            const double e = exact_d2(xi, yi, q.x, q.y, P);
            if (e > T.maxd) { T.maxd = e; T.mj = j; T.max_thr = __double2float_rd(e * (1.0 - 2e-6)); }
        }
    }
}

__global__ void __launch_bounds__(OT)
obs_pairs_kernel(const __grid_constant__ DevParams P, DevState S, const int *__restrict__ cidx, const int *__restrict__ off,
                 int row_blocks, double hd0, double *nnd_q, double *nnd_t, double *iid_row, MaxRec *part_max) {
    __shared__ float2 tile[OT];
    __shared__ MaxRec shm[OT / 32];
    const int rep = blockIdx.y;
    const int o0 = off[rep], n = off[rep + 1] - o0;
    const int k0 = blockIdx.x * OT;
    if (k0 >= n) {
        if (threadIdx.x == 0) part_max[(size_t)rep * row_blocks + blockIdx.x] = MaxRec{-1.0, 0, 0, 0};
        return;
    }
    const int k = k0 + threadIdx.x;
    const bool row = k < n;
    float xi = 0.f, yi = 0.f;
    if (row) { const float4 p = S.pv[cidx[o0 + k]]; xi = p.x; yi = p.y; }
    RowTrack T;
    T.minq = INFINITY; T.maxd = -1.0; T.min_thr = INFINITY; T.max_thr = -1.f; T.mj = 0; T.iid = 0.f;
    double iid = 0.0;
    const int kmax = min(n, k0 + OT) - 1;
    for (int j0 = 0; j0 < n; j0 += OT) {
        const int cnt = min(OT, n - j0);
        __syncthreads();
        if (threadIdx.x < cnt) { const float4 p = S.pv[cidx[o0 + j0 + threadIdx.x]]; tile[threadIdx.x] = make_float2(p.x, p.y); }
        __syncthreads();
        if (!row) continue;
        if (j0 + cnt - 1 < k0 - 1) sweep_tile<0>(tile, j0, cnt, k, xi, yi, P, T);
        else if (j0 > kmax) sweep_tile<1>(tile, j0, cnt, k, xi, yi, P, T);
        else sweep_tile<2>(tile, j0, cnt, k, xi, yi, P, T);
        iid += (double)T.iid;
        T.iid = 0.f;
    }
    if (row) {
        const size_t g = (size_t)o0 + k;
        nnd_q[g] = fmin(hd0, sqrt(T.minq));                                 // :273-294, hd starts at N * alg_range
        double e = T.minq;
        if (k > 0) { const float4 p = S.pv[cidx[o0 + k - 1]]; e = fmin(e, exact_d2(xi, yi, p.x, p.y, P)); }
        nnd_t[g] = sqrt(e);                                                // the true nearest neighbour
        iid_row[g] = iid;
    }
    // the block's most distant pair, first in (i, j) order among equals
    MaxRec m{row ? T.maxd : -1.0, k, T.mj, 0};
    for (int o = 16; o > 0; o >>= 1) {
        MaxRec b;
        b.d2 = __shfl_xor_sync(0xffffffffu, m.d2, o); b.i = __shfl_xor_sync(0xffffffffu, m.i, o); b.j = __shfl_xor_sync(0xffffffffu, m.j, o);
        if (max_better(b.d2, b.i, b.j, m.d2, m.i, m.j)) { m.d2 = b.d2; m.i = b.i; m.j = b.j; }
    }
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < OT / 32; w++)
            if (max_better(shm[w].d2, shm[w].i, shm[w].j, m.d2, m.i, m.j)) m = shm[w];
        part_max[(size_t)rep * row_blocks + blockIdx.x] = m;
    }
}

__global__ void __launch_bounds__(OT)
obs_final2_kernel(const __grid_constant__ DevParams P, DevState S, const int *__restrict__ cidx, const int *__restrict__ off,
                  int row_blocks, int with_pairs, const double *__restrict__ nnd_q, const double *__restrict__ nnd_t,
                  const double *__restrict__ iid_row, const MaxRec *__restrict__ part_max, RepRes *res) {
    __shared__ double sh[OT / 32];
    const int rep = blockIdx.x;
    const int o0 = off[rep], n = off[rep + 1] - o0;
    RepRes &q = res[rep];
    double a = 0, b = 0, c = 0;
    if (with_pairs)
        for (int k = threadIdx.x; k < n; k += OT) { a += nnd_q[o0 + k]; b += nnd_t[o0 + k]; c += iid_row[o0 + k]; }
    a = block_sum(a, sh); b = block_sum(b, sh); c = block_sum(c, sh);
    if (threadIdx.x != 0) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    q.nnd_q = with_pairs ? a : nan; q.nnd_true = with_pairs ? b : nan; q.iid_sum = with_pairs ? c : nan;
    MaxRec m{-1.0, 0, 0, 0};
    if (with_pairs)
        for (int w = 0; w < (n + OT - 1) / OT; w++) {
            const MaxRec x = part_max[(size_t)rep * row_blocks + w];
            if (max_better(x.d2, x.i, x.j, m.d2, m.i, m.j)) m = x;
        }
    q.max_vec[0] = q.max_vec[1] = 1.0;                                        // :247
    q.max_d2 = 0.0; q.max_i = q.max_j = -1;
    if (m.d2 > 0.0) {                                                         // :285 `dist > max_IID` from 0
        const float4 pi = S.pv[cidx[o0 + m.i]], pj = S.pv[cidx[o0 + m.j]];
        dist_vec_dd(pi.x, pi.y, pj.x, pj.y, P, q.max_vec);
        q.max_d2 = m.d2; q.max_i = m.i; q.max_j = m.j;
    }
    if (!with_pairs) q.max_vec[0] = q.max_vec[1] = nan;
    // get_elongation's unit vectors (agents_operation.cpp:155-160)
    const double lu = sqrt(__dadd_rn(__dmul_rn(q.avg_u[0], q.avg_u[0]), __dmul_rn(q.avg_u[1], q.avg_u[1])));
    q.dir_u[0] = q.avg_u[0] / lu; q.dir_u[1] = q.avg_u[1] / lu;
    const double lm = sqrt(__dadd_rn(__dmul_rn(q.max_vec[0], q.max_vec[0]), __dmul_rn(q.max_vec[1], q.max_vec[1])));
    q.dir_m[0] = q.max_vec[0] / lm; q.dir_m[1] = q.max_vec[1] / lm;
}

__global__ void __launch_bounds__(OT)
obs_second_kernel(DevState S, const int *__restrict__ cidx, const int *__restrict__ off, int chunks,
                  const RepRes *__restrict__ res, double *part, HullPt *hull, int *hull_count, int hull_cap) {
    __shared__ double sh[OT / 32];
    const int rep = blockIdx.y, c = blockIdx.x;
    const int o0 = off[rep], n = off[rep + 1] - o0;
    if (c * CHUNK >= n && c > 0) return;
    const RepRes &q = res[rep];
    const double big = 1.7976931348623157e308;
    double Ls = 0, mn[4] = {big, big, big, big}, mx[4] = {-big, -big, -big, -big};
    const double dir[4][2] = {{q.dir_u[0], q.dir_u[1]}, {q.dir_u[1], -q.dir_u[0]}, {q.dir_m[0], q.dir_m[1]}, {q.dir_m[1], -q.dir_m[0]}};
    const int end = min(n, (c + 1) * CHUNK);
    for (int k = c * CHUNK + threadIdx.x; k < end; k += OT) {
        const float4 p = S.pv[cidx[o0 + k]];
        const double x = p.x, y = p.y;
        const double hx = x - q.avg_x[0], hy = y - q.avg_x[1];
        Ls += __dsub_rn(__dmul_rn(hx, (double)p.w), __dmul_rn(hy, (double)p.z)) / sqrt(__dadd_rn(__dmul_rn(hx, hx), __dmul_rn(hy, hy)));   // :301-305
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const double pr = __dadd_rn(__dmul_rn(x, dir[d][0]), __dmul_rn(y, dir[d][1]));
            mn[d] = fmin(mn[d], pr); mx[d] = fmax(mx[d], pr);
        }
        // hull filter: strictly inside the support polygon (by a margin far above the rounding of the test) = no vertex
        bool inside = q.npoly >= 3;
        for (int e = 0; e < q.npoly && inside; e++) {
            const float2 A = q.poly[e], B = q.poly[e + 1 == q.npoly ? 0 : e + 1];
            const double ex = (double)B.x - A.x, ey = (double)B.y - A.y;
            const double cr = ex * (y - A.y) - ey * (x - A.x);
            inside = cr > 1e-9 * (fabs(ex) + fabs(ey)) * (fabs(x - A.x) + fabs(y - A.y) + 1e-30);
        }
        if (!inside) {
            const int slot = atomicAdd(hull_count, 1);
            if (slot < hull_cap) hull[slot] = HullPt{rep, p.x, p.y};
        }
    }
    double *o = part + ((size_t)rep * chunks + c) * NSEC;
    const double t = block_sum(Ls, sh);
    if (threadIdx.x == 0) o[0] = t;
    for (int d = 0; d < 4; d++) {
        const double a = block_min(mn[d], sh), b = -block_min(-mx[d], sh);
        if (threadIdx.x == 0) { o[1 + 2 * d] = a; o[2 + 2 * d] = b; }
    }
}

__global__ void obs_final3_kernel(int R, double hdN, int with_pairs, const int *__restrict__ off, int chunks,
                                  const double *__restrict__ part, const RepRes *__restrict__ res, double *out16, double *extra) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= R) return;
    const RepRes &q = res[rep];
    const int n = off[rep + 1] - off[rep];
    const double big = 1.7976931348623157e308, nan = __longlong_as_double(0x7ff8000000000000ll);
    double Ls = 0, mn[4] = {big, big, big, big}, mx[4] = {-big, -big, -big, -big};
    const int used = max(1, (n + CHUNK - 1) / CHUNK);
    for (int c = 0; c < used; c++) {
        const double *o = part + ((size_t)rep * chunks + c) * NSEC;
        Ls += o[0];
        for (int d = 0; d < 4; d++) { mn[d] = fmin(mn[d], o[1 + 2 * d]); mx[d] = fmax(mx[d], o[2 + 2 * d]); }
    }
    const double N = q.n;
    const double avgvel = sqrt(__dadd_rn(__dmul_rn(q.avg_v[0], q.avg_v[0]), __dmul_rn(q.avg_v[1], q.avg_v[1])));
    const double pol = sqrt(__dadd_rn(__dmul_rn(q.avg_u[0], q.avg_u[0]), __dmul_rn(q.avg_u[1], q.avg_u[1])));
    const double lm = sqrt(__dadd_rn(__dmul_rn(q.max_vec[0], q.max_vec[0]), __dmul_rn(q.max_vec[1], q.max_vec[1])));
    double a1 = (q.avg_v[0] * q.max_vec[0] + q.avg_v[1] * q.max_vec[1]) / (avgvel * lm);     // :311-316
    const double a2 = acos(-a1);
    a1 = acos(a1);
    double *o = out16 + (size_t)rep * 16;
    o[0] = N; o[1] = pol; o[2] = fabs(Ls) / N; o[3] = q.avg_x[0]; o[4] = q.avg_x[1]; o[5] = q.avg_v[0]; o[6] = q.avg_v[1];
    o[7] = q.avg_s; o[8] = q.avg_v2;
    o[9] = fabs((mx[0] - mn[0]) / (mx[1] - mn[1]));
    o[10] = with_pairs ? fabs((mx[2] - mn[2]) / (mx[3] - mn[3])) : nan;
    o[11] = with_pairs ? fmin(a1, a2) : nan;
    o[12] = 0.0;                                   // hull area: filled in by the host from the filtered points
    o[13] = q.nnd_q / N;
    o[14] = q.iid_sum / ((N - 1.0) * N);
    o[15] = nan;                                   // ND, see the header
    if (extra) {
        double *e = extra + (size_t)rep * 4;
        e[0] = q.nnd_true / N;                     // mean distance to the true nearest neighbour
        e[1] = q.iid_sum / (0.5 * (N - 1.0) * N);  // mean distance over unordered pairs
        e[2] = sqrt(q.max_d2);                     // largest inter-individual distance
        e[3] = 0.0;                                // hull vertices (host)
    }
    (void)hdN;
}

// ---- fish net (swarmdyn.cpp:358-410) -------------------------------------------------------------------------------
__global__ void __launch_bounds__(OT)
obs_fishnet_kernel(const __grid_constant__ DevParams P, DevState S, const int *__restrict__ cidx, const int *__restrict__ off,
                   int chunks, double *part) {
    __shared__ double sh[OT / 32];
    const int rep = blockIdx.y, c = blockIdx.x;
    const int o0 = off[rep], n = off[rep + 1] - o0;
    if (c * CHUNK >= n && c > 0) return;
    const int np = S.Npred;
    const float4 *pp = S.pred_pv + (size_t)rep * np;
    double mvx = 0, mvy = 0, vnet[2] = {0, 0}, net_len = 0;
    if (np > 1) {
        sincos((double)S.pred_phi[(size_t)rep * np], &mvy, &mvx);      // predator::u of node 0
        dist_vec_dd(pp[0].x, pp[0].y, pp[1].x, pp[1].y, P, vnet);
        net_len = (np - 1) * sqrt(__dadd_rn(__dmul_rn(vnet[0], vnet[0]), __dmul_rn(vnet[1], vnet[1])));
    }
    double dsum = 0, nfront = 0, nclose = 0;
    const int end = min(n, (c + 1) * CHUNK);
    for (int k = c * CHUNK + threadIdx.x; k < end; k += OT) {
        const float4 p = S.pv[cidx[o0 + k]];
        double r[2];
        if (np > 1) {
            dist_vec_dd(pp[0].x, pp[0].y, p.x, p.y, P, r);
            const double front = __dadd_rn(__dmul_rn(mvx, r[0]), __dmul_rn(mvy, r[1]));
            const double side = __dadd_rn(__dmul_rn(vnet[0], r[0]), __dmul_rn(vnet[1], r[1]));
            if (side > 0 && side <= net_len) { nfront += 1; if (front > 0 && front < P.d_kill_range) nclose += 1; }
        }
        double dmin = P.dL * 2;
        for (int j = 0; j < np; j++) {
            dist_vec_dd(pp[j].x, pp[j].y, p.x, p.y, P, r);
            dmin = fmin(dmin, sqrt(__dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1]))));
        }
        dsum += dmin;
    }
    double *o = part + ((size_t)rep * chunks + c) * 3;
    const double a = block_sum(dsum, sh), b = block_sum(nfront, sh), d = block_sum(nclose, sh);
    if (threadIdx.x == 0) { o[0] = a; o[1] = b; o[2] = d; }
}
__global__ void obs_fishnet_final_kernel(DevState S, const int *__restrict__ off, int chunks, const double *__restrict__ part, double *out4) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= S.R) return;
    const int n = off[rep + 1] - off[rep];
    double a = 0, b = 0, d = 0;
    const int used = max(1, (n + CHUNK - 1) / CHUNK);
    for (int c = 0; c < used; c++) { const double *o = part + ((size_t)rep * chunks + c) * 3; a += o[0]; b += o[1]; d += o[2]; }
    double *o = out4 + (size_t)rep * 4;
    o[0] = d; o[1] = a / n; o[2] = b; o[3] = S.n_dead[rep];
}

// the 8-column subset sbc_observables has always returned (NND = the true nearest neighbour)
__global__ void obs_subset_kernel(int R, const double *__restrict__ out16, const double *__restrict__ extra, double *out8) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= R) return;
    const double *o = out16 + (size_t)rep * 16;
    double *q = out8 + (size_t)rep * 8;
    const bool none = o[0] == 0;
    q[0] = o[0];
    q[1] = none ? 0 : o[1]; q[2] = none ? 0 : o[3]; q[3] = none ? 0 : o[4]; q[4] = none ? 0 : o[5]; q[5] = none ? 0 : o[6];
    q[6] = none ? 0 : o[7];
    q[7] = none ? 0 : (o[0] < 2 ? 0 : extra[(size_t)rep * 4]);
}

double hull_area(std::vector<std::pair<double, double>> &pts, int *n_vertices) {   // monotone chain; AreaConvexHull agents_operation.cpp:130-150
    std::sort(pts.begin(), pts.end());
    pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
    const int n = (int)pts.size();
    *n_vertices = n;
    if (n < 3) return 0.0;
    std::vector<std::pair<double, double>> h(2 * n);
    auto cross = [](const std::pair<double, double> &o, const std::pair<double, double> &a, const std::pair<double, double> &b) {
        return (a.first - o.first) * (b.second - o.second) - (a.second - o.second) * (b.first - o.first);
    };
    int k = 0;
    for (int i = 0; i < n; i++) { while (k >= 2 && cross(h[k - 2], h[k - 1], pts[i]) <= 0) k--; h[k++] = pts[i]; }
    for (int i = n - 2, t = k + 1; i >= 0; i--) { while (k >= t && cross(h[k - 2], h[k - 1], pts[i]) <= 0) k--; h[k++] = pts[i]; }
    *n_vertices = k - 1;
    double area = 0;
    for (int i = 0; i + 1 < k; i++) area += h[i].first * h[i + 1].second - h[i + 1].first * h[i].second;
    return 0.5 * area;
}
}  // namespace

struct ObsScratch {
    size_t total = 0; int R = 0, chunks = 0, row_blocks = 0, hull_cap = 0;
    int *cidx = nullptr, *n_sel = nullptr, *off = nullptr, *hull_count = nullptr;
    void *cub_tmp = nullptr; size_t cub_bytes = 0;
    double *part = nullptr, *nnd_q = nullptr, *nnd_t = nullptr, *iid_row = nullptr, *out16 = nullptr, *extra = nullptr, *out8 = nullptr;
    unsigned long long *ext = nullptr;
    MaxRec *part_max = nullptr;
    RepRes *res = nullptr;
    HullPt *hull = nullptr;
};

void sbc_obs_free(ObsScratch *o) {
    if (!o) return;
    cudaFree(o->cidx); cudaFree(o->n_sel); cudaFree(o->off); cudaFree(o->hull_count); cudaFree(o->cub_tmp); cudaFree(o->part);
    cudaFree(o->nnd_q); cudaFree(o->nnd_t); cudaFree(o->iid_row); cudaFree(o->out16); cudaFree(o->extra); cudaFree(o->out8);
    cudaFree(o->ext); cudaFree(o->part_max); cudaFree(o->res); cudaFree(o->hull);
    delete o;
}

#define OBS_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return SBC_ERR_CUDA; } } while (0)

static int obs_prepare(const DevState &S, ObsScratch **scr, std::string &err) {
    if (*scr) return SBC_OK;
    ObsScratch *o = new ObsScratch;
    *scr = o;
    o->total = (size_t)S.N * S.R; o->R = S.R;
    o->chunks = (S.N + CHUNK - 1) / CHUNK;
    o->row_blocks = (S.N + OT - 1) / OT;
    o->hull_cap = (int)std::min<size_t>(o->total, (size_t)1 << 24);
    AliveAt pred{S.pv};
    cub::CountingInputIterator<int> it(0);
    OBS_CUDA(cub::DeviceSelect::If(nullptr, o->cub_bytes, it, o->cidx, o->n_sel, (int)o->total, pred));
    OBS_CUDA(cudaMalloc(&o->cub_tmp, o->cub_bytes ? o->cub_bytes : 16));
    OBS_CUDA(cudaMalloc(&o->cidx, o->total * sizeof(int)));
    OBS_CUDA(cudaMalloc(&o->n_sel, sizeof(int)));
    OBS_CUDA(cudaMalloc(&o->off, ((size_t)S.R + 1) * sizeof(int)));
    OBS_CUDA(cudaMalloc(&o->hull_count, sizeof(int)));
    OBS_CUDA(cudaMalloc(&o->part, (size_t)S.R * o->chunks * NSEC * sizeof(double)));
    OBS_CUDA(cudaMalloc(&o->nnd_q, o->total * sizeof(double)));
    OBS_CUDA(cudaMalloc(&o->nnd_t, o->total * sizeof(double)));
    OBS_CUDA(cudaMalloc(&o->iid_row, o->total * sizeof(double)));
    OBS_CUDA(cudaMalloc(&o->out16, (size_t)S.R * 16 * sizeof(double)));
    OBS_CUDA(cudaMalloc(&o->extra, (size_t)S.R * 4 * sizeof(double)));
    OBS_CUDA(cudaMalloc(&o->out8, (size_t)S.R * 8 * sizeof(double)));
    OBS_CUDA(cudaMalloc(&o->ext, (size_t)S.R * NDIR * sizeof(unsigned long long)));
    OBS_CUDA(cudaMalloc(&o->part_max, (size_t)S.R * o->row_blocks * sizeof(MaxRec)));
    OBS_CUDA(cudaMalloc(&o->res, (size_t)S.R * sizeof(RepRes)));
    OBS_CUDA(cudaMalloc(&o->hull, (size_t)o->hull_cap * sizeof(HullPt)));
    return SBC_OK;
}

static int obs_select(const DevState &S, ObsScratch *o, cudaStream_t st, std::string &err) {
    AliveAt pred{S.pv};
    cub::CountingInputIterator<int> it(0);
    size_t bytes = o->cub_bytes;
    OBS_CUDA(cub::DeviceSelect::If(o->cub_tmp, bytes, it, o->cidx, o->n_sel, (int)o->total, pred, st));
    obs_offsets_kernel<<<(S.R + 1 + 255) / 256, 256, 0, st>>>(o->cidx, o->n_sel, S.N, S.R, o->off);
    return SBC_OK;
}

// out16_host [R][16], extra_host [R][4] or nullptr, out8_host [R][8] or nullptr (the subset of sbc_observables)
int sbc_obs_out_swarm(const DevParams &P, const DevState &S, const sbc_params &hp, ObsScratch **scr, int flags,
                      double *out16_host, double *extra_host, double *out8_host, cudaStream_t st, std::string &err) {
    if (int rc = obs_prepare(S, scr, err)) return rc;
    ObsScratch *o = *scr;
    const int R = S.R, with_pairs = (flags & SBC_OBS_NO_PAIRS) ? 0 : 1;
    if (int rc = obs_select(S, o, st, err)) return rc;
    OBS_CUDA(cudaMemsetAsync(o->ext, 0, (size_t)R * NDIR * sizeof(unsigned long long), st));
    OBS_CUDA(cudaMemsetAsync(o->hull_count, 0, sizeof(int), st));
    const dim3 gchunks(o->chunks, R);
    obs_sums_kernel<<<gchunks, OT, 0, st>>>(S, o->cidx, o->off, o->chunks, o->part, o->ext);
    obs_final1_kernel<<<(R + 127) / 128, 128, 0, st>>>(S, o->cidx, o->off, o->chunks, o->part, o->ext, o->res);
    const double hd0 = (double)S.N * hp.alg_range;   // SP.N keeps the initial agent count
    if (with_pairs)
        obs_pairs_kernel<<<dim3(o->row_blocks, R), OT, 0, st>>>(P, S, o->cidx, o->off, o->row_blocks, hd0, o->nnd_q, o->nnd_t,
                                                                 o->iid_row, o->part_max);
    obs_final2_kernel<<<R, OT, 0, st>>>(P, S, o->cidx, o->off, o->row_blocks, with_pairs, o->nnd_q, o->nnd_t, o->iid_row,
                                        o->part_max, o->res);
    obs_second_kernel<<<gchunks, OT, 0, st>>>(S, o->cidx, o->off, o->chunks, o->res, o->part, o->hull, o->hull_count, o->hull_cap);
    obs_final3_kernel<<<(R + 127) / 128, 128, 0, st>>>(R, hd0, with_pairs, o->off, o->chunks, o->part, o->res, o->out16, o->extra);
    if (out8_host) obs_subset_kernel<<<(R + 127) / 128, 128, 0, st>>>(R, o->out16, o->extra, o->out8);
    OBS_CUDA(cudaGetLastError());
    int nh = 0;
    OBS_CUDA(cudaMemcpyAsync(&nh, o->hull_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    OBS_CUDA(cudaStreamSynchronize(st));
    if (out8_host) OBS_CUDA(cudaMemcpy(out8_host, o->out8, (size_t)R * 8 * sizeof(double), cudaMemcpyDeviceToHost));
    if (!out16_host) return SBC_OK;
    if (nh > o->hull_cap) { err = "sbc_out_swarm: hull candidate list overflow"; return SBC_ERR_STATE; }
    std::vector<HullPt> hp_host((size_t)nh);
    if (nh) OBS_CUDA(cudaMemcpy(hp_host.data(), o->hull, (size_t)nh * sizeof(HullPt), cudaMemcpyDeviceToHost));
    OBS_CUDA(cudaMemcpy(out16_host, o->out16, (size_t)R * 16 * sizeof(double), cudaMemcpyDeviceToHost));
    if (extra_host) OBS_CUDA(cudaMemcpy(extra_host, o->extra, (size_t)R * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<int> count((size_t)R + 1, 0);
    for (const HullPt &p : hp_host) count[p.rep + 1]++;
    for (int r = 0; r < R; r++) count[r + 1] += count[r];
    std::vector<std::pair<double, double>> sorted((size_t)nh);
    std::vector<int> fill(count.begin(), count.end() - 1);
    for (const HullPt &p : hp_host) sorted[fill[p.rep]++] = {(double)p.x, (double)p.y};
    for (int r = 0; r < R; r++) {
        std::vector<std::pair<double, double>> pts(sorted.begin() + count[r], sorted.begin() + count[r + 1]);
        int nv = 0;
        out16_host[(size_t)r * 16 + 12] = hull_area(pts, &nv);
        if (extra_host) extra_host[(size_t)r * 4 + 3] = nv;
    }
    return SBC_OK;
}

int sbc_obs_fishnet(const DevParams &P, const DevState &S, ObsScratch **scr, double *out4_host, cudaStream_t st, std::string &err) {
    if (int rc = obs_prepare(S, scr, err)) return rc;
    ObsScratch *o = *scr;
    if (int rc = obs_select(S, o, st, err)) return rc;
    obs_fishnet_kernel<<<dim3(o->chunks, S.R), OT, 0, st>>>(P, S, o->cidx, o->off, o->chunks, o->part);
    obs_fishnet_final_kernel<<<(S.R + 127) / 128, 128, 0, st>>>(S, o->off, o->chunks, o->part, o->out16);
    OBS_CUDA(cudaGetLastError());
    OBS_CUDA(cudaStreamSynchronize(st));
    OBS_CUDA(cudaMemcpy(out4_host, o->out16, (size_t)S.R * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return SBC_OK;
}
