// state_io.cu - conversion between the C ABI's host layout (arrays of doubles, the reference's
// `particle` fields, /root/reference/src/agents.h:29-58) and the FP32 structure-of-arrays device
// state, done ON THE DEVICE: the host side of sbc_set_state / sbc_get_state / sbc_get_particles
// only issues copies into / out of a staging area, no per-agent host loop.
#include "sbc_kernels.cuh"

namespace {

// one real number of the caller's staged array k (double, or float for the _f32 entry points)
__device__ __forceinline__ float stage_real(const StateStage &st, int k, size_t g) {
    return st.f32 ? reinterpret_cast<const float *>(st.d[k])[g] : (float)st.d[k][g];
}

__global__ void __launch_bounds__(256)
pack_state_kernel(const __grid_constant__ DevParams P, DevState S, StateStage st, unsigned mask, int have_prev) {
    const size_t total = (size_t)S.N * S.R;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 hv = make_float2(0.f, 1.f);  // InitSystem: phi = 0, vproj = 1 (settings.cpp:168-193)
    float2 cs = make_float2(1.f, 0.f);
    float2 fo = make_float2(0.f, 0.f);
    uint32_t bin = 0u, stb = 0u;
    uint8_t al = 1;
    if (have_prev) {
        al = S.alive[g];
        pv = al ? S.pv[g] : S.pv_dead[g];
        hv = S.hv[g];
        cs = S.cs[g];
        fo = S.force[g];
        const uint32_t t = S.tm[g];
        bin = sbc_tm_bin(t, P);
        stb = sbc_tm_stb(t, P);
    }
    if (mask & SBC_ST_X) pv.x = stage_real(st, 0, g);
    if (mask & SBC_ST_Y) pv.y = stage_real(st, 1, g);
    if (mask & SBC_ST_PHI) {
        hv.x = stage_real(st, 4, g);
        sincosf(hv.x, &cs.y, &cs.x);
    }
    if (mask & SBC_ST_VPROJ) hv.y = stage_real(st, 5, g);
    if (mask & SBC_ST_VX) pv.z = stage_real(st, 2, g);
    else if (mask & SBC_ST_PHI) pv.z = hv.y * cs.x;  // v = vproj u(phi), settings.cpp:374-377
    if (mask & SBC_ST_VY) pv.w = stage_real(st, 3, g);
    else if (mask & SBC_ST_PHI) pv.w = hv.y * cs.y;
    if (mask & SBC_ST_FX) fo.x = stage_real(st, 6, g);
    if (mask & SBC_ST_FY) fo.y = stage_real(st, 7, g);
    if (mask & SBC_ST_BIN) bin = min(st.u[0][g], P.tm_mask);
    if (mask & SBC_ST_STB) stb = st.u[1][g];
    if (mask & SBC_ST_ALIVE) al = st.al[g] ? 1 : 0;
    S.pv_dead[g] = pv;
    if (!al) {
        const float qnan = __int_as_float(0x7fc00000);
        pv.x = qnan;
        pv.y = qnan;
        atomicAdd(S.n_dead + g / S.N, 1);
    }
    S.pv[g] = pv;
    S.pv_nxt[g] = pv;
    S.hv[g] = hv;
    S.cs[g] = cs;
    S.force[g] = fo;
    S.tm[g] = sbc_tm_pack(bin, stb, P);
    S.alive[g] = al;
}

__global__ void __launch_bounds__(256)
unpack_state_kernel(const __grid_constant__ DevParams P, DevState S, StateStage st, unsigned mask) {
    const size_t total = (size_t)S.N * S.R;
    const size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (g >= total) return;
    const uint8_t al = S.alive[g];
    const float4 pv = al ? S.pv[g] : S.pv_dead[g];
    const float2 hv = S.hv[g];
    const float2 fo = S.force[g];
    const uint32_t tm = S.tm[g];
    if (mask & SBC_ST_X) st.d[0][g] = pv.x;
    if (mask & SBC_ST_Y) st.d[1][g] = pv.y;
    if (mask & SBC_ST_VX) st.d[2][g] = pv.z;
    if (mask & SBC_ST_VY) st.d[3][g] = pv.w;
    if (mask & SBC_ST_PHI) st.d[4][g] = hv.x;
    if (mask & SBC_ST_VPROJ) st.d[5][g] = hv.y;
    if (mask & SBC_ST_FX) st.d[6][g] = fo.x;
    if (mask & SBC_ST_FY) st.d[7][g] = fo.y;
    if (mask & SBC_ST_BIN) st.u[0][g] = sbc_tm_bin(tm, P);
    if (mask & SBC_ST_STB) st.u[1][g] = sbc_tm_stb(tm, P);
    if (mask & SBC_ST_ALIVE) st.al[g] = al;
}

// rows of particle::out() (agents.cpp:4-17): x0 x1 v0 v1 fitness(=0) force0 force1; dead -> zeros.
// One thread per output double so that the stores are contiguous.
__global__ void __launch_bounds__(256)
particles_kernel(DevState S, double *__restrict__ out, uint8_t *__restrict__ alive_out, int f32) {
    const size_t total = (size_t)S.N * S.R;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= total * 7) return;
    const size_t g = t / 7;
    const int k = (int)(t - g * 7);
    const uint8_t al = S.alive[g];
    double v = 0.0;
    if (al) {
        if (k < 4) {
            const float4 pv = S.pv[g];
            v = (k == 0) ? pv.x : (k == 1) ? pv.y : (k == 2) ? pv.z : pv.w;
        } else if (k > 4) {
            const float2 fo = S.force[g];
            v = (k == 5) ? fo.x : fo.y;
        }
    }
    if (f32) reinterpret_cast<float *>(out)[t] = (float)v;
    else out[t] = v;
    if (k == 0 && alive_out) alive_out[g] = al;
}

}  // namespace


namespace {
// ---- initial conditions on the device (ResetSystem, /root/reference/src/settings.cpp:196-387) -------------------------
// Every uniform the reference pops from its serial GSL stream is a word of a Philox block here:
//   counter (agent = loop index i, step = 0xFFFFFFFF, slot = 16, replicate):  .x -> x draw, .y -> y draw, .z -> heading
//   counter (agent = 0xFFFFFFFE, same step / slot):                           .x -> the swarm-wide heading of IC 1, 2, 3
//   counter (agent = 0xFFFFFFFE, slot = 17):                                  four round keys of the ring shuffle
// The shuffle of IC 3-5 (std::shuffle / random_shuffle with a clock seed in the reference) is a keyed bijection of
// [0, N): a four-round Feistel network on the next even power of two, cycle-walked back into range. Arithmetic is FP64 in
// the reference's operation order (no contraction), the state is rounded to FP32 by the pack kernel as for any upload.
#define SBC_IC_STEP 0xFFFFFFFFu
#define SBC_IC_SWARM_ID 0xFFFFFFFEu

__device__ __forceinline__ uint32_t ic_mix(uint32_t v) {
    v *= 0x9E3779B1u; v ^= v >> 15; v *= 0x85EBCA77u; v ^= v >> 13;
    return v;
}
__device__ uint32_t ic_permute(uint32_t i, uint32_t n, const uint4 key) {
    uint32_t bits = 2;
    while ((1ull << bits) < n) bits += 2;
    const uint32_t half = bits / 2, mask = (1u << half) - 1u;
    const uint32_t k[4] = {key.x, key.y, key.z, key.w};
    uint32_t x = i;
    do {
        uint32_t l = x >> half, r = x & mask;
        for (int q = 0; q < 4; q++) { const uint32_t t = l ^ (ic_mix(r ^ k[q]) & mask); l = r; r = t; }
        x = (l << half) | r;
    } while (x >= n);
    return x;
}

// Boundary<particle> (agents_dynamics.cpp:276-471) on a fresh agent, FP64 in the reference's operation order. As there,
// a.phi follows the velocity only at the circular walls (:447-449, :467-469); the box walls flip or replace v and leave phi.
__device__ void ic_boundary(double &x, double &y, double &vx, double &vy, double &phi, double L, int BC) {
    const double dphi = 0.1;
    if (BC == 0) {
        x = fmod(__dadd_rn(x, L), L);
        y = fmod(__dadd_rn(y, L), L);
    } else if (BC == 1) {                        // :300-356 inelastic box
        double t = 0;
        bool hit = true;
        if (x > L) { x = __dmul_rn(0.9999, L); t = (vy > 0.) ? __dmul_rn(0.5 + dphi, SBC_PI_D) : __dmul_rn(-(0.5 + dphi), SBC_PI_D); }
        else if (x < 0) { x = 0.0001; t = (vy > 0.) ? __dmul_rn(0.5 - dphi, SBC_PI_D) : __dmul_rn(-(0.5 - dphi), SBC_PI_D); }
        else hit = false;
        if (hit) { vx = cos(t); vy = sin(t); }
        hit = true;
        if (y > L) { y = __dmul_rn(0.9999, L); t = (vx > 0.) ? -dphi : __dadd_rn(SBC_PI_D, dphi); }
        else if (y < 0) { y = 0.0001; t = (vx > 0.) ? dphi : __dsub_rn(SBC_PI_D, dphi); }
        else hit = false;
        if (hit) { vx = cos(t); vy = sin(t); }
    } else if (BC == 2 || BC == 3) {             // :358-406 elastic walls (BC 3: x periodic)
        if (BC == 3) { double tx = fmod(x, L); if (tx < 0.0) tx = __dadd_rn(tx, L); x = tx; }
        else if (x > L) { x = __dsub_rn(x, __dmul_rn(2., __dsub_rn(x, L))); vx = -vx; }
        else if (x < 0) { x = __dsub_rn(x, __dmul_rn(2., x)); vx = -vx; }
        if (y > L) { y = __dsub_rn(y, __dmul_rn(2., __dsub_rn(y, L))); vy = -vy; }
        else if (y < 0) { y = __dsub_rn(y, __dmul_rn(2., y)); vy = -vy; }
    } else if (BC == 4) {                        // :407-430 x periodic, y inelastic
        double tx = fmod(x, L);
        if (tx < 0.0) tx = __dadd_rn(tx, L);
        x = tx;
        if (y > L) { y = __dsub_rn(y, __dadd_rn(__dsub_rn(y, L), 0.0001)); vy = 0.0; }
        else if (y < 0) { y = __dsub_rn(y, __dsub_rn(y, 0.0001)); vy = 0.0; }
    } else if (BC == 5 || BC == 6) {             // :432-470 circular tank
        const double r = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
        double diff = __dsub_rn(r, L);
        if (diff > 0) {
            const double inv = 1 / r, nx = __dmul_rn(x, inv), ny = __dmul_rn(y, inv), f = __dmul_rn(-2, diff);
            x = __dadd_rn(x, __dmul_rn(nx, f));
            y = __dadd_rn(y, __dmul_rn(ny, f));
            diff = __dadd_rn(__dmul_rn(nx, vx), __dmul_rn(ny, vy));
            const double g = (BC == 5) ? __dmul_rn(-2, diff) : -diff;
            vx = __dadd_rn(vx, __dmul_rn(nx, g));
            vy = __dadd_rn(vy, __dmul_rn(ny, g));
            const double len = sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
            const double corr = 1 / len;                   // vec_set_mag, mathtools.h:138-144
            phi = atan2(__dmul_rn(vy, corr), __dmul_rn(vx, corr));
        }
    }
}

__global__ void __launch_bounds__(256)
init_cond_kernel(const __grid_constant__ DevParams P, int N, int R, int ic, double rep_range, StateStage st) {
    const size_t total = (size_t)N * R;
    const size_t gi = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (gi >= total) return;
    const uint32_t rep = (uint32_t)(gi / N), i = (uint32_t)(gi % N);     // i = the reference's loop index
    const double u32 = 1.0 / 4294967296.0, two_pi = __dmul_rn(2, SBC_PI_D), L = P.dL;
    const double range = __dmul_rn(3, rep_range);
    const uint4 d = sbc_draw(P, rep, i, SBC_IC_STEP, 16u);
    const uint4 sw = sbc_draw(P, rep, SBC_IC_SWARM_ID, SBC_IC_STEP, 16u);
    const double ux = d.x * u32, uy = d.y * u32, up = d.z * u32, theta_all = __dmul_rn(two_pi, sw.x * u32);
    double x, y, phi;
    uint32_t agent = i;
    if (ic == 1) {                                   // :212-221
        x = __dmul_rn(L, ux); y = __dmul_rn(L, uy); phi = theta_all;
    } else if (ic == 0 || ic == 2) {                 // :237-257
        const double side = fmin(L, __dmul_rn(sqrt((double)N), range));
        x = __dmul_rn(side, ux); y = __dmul_rn(side, uy);
        phi = (ic == 0) ? __dmul_rn(two_pi, up) : theta_all;
    } else if (ic == 3 || ic == 4 || ic == 5) {      // :258-355, agents on perturbed concentric rings
        int circle = (ic == 3) ? 1 : 3, ncir = (ic == 3) ? 3 : (int)__dmul_rn(two_pi, 3.0), ncap = ncir, base = 0;
        while ((int)i >= ncap) { circle++; ncir = (int)__dmul_rn(two_pi, (double)circle); base = ncap; ncap += ncir; }
        const double ang = __dmul_rn(two_pi, (double)((int)i - base)) / (double)ncir;
        const double sig = (ic == 3) ? 0.8 : 1.0, rc = __dmul_rn(range, (double)circle), sr = __dmul_rn(sig, range);
        x = __dadd_rn(__dmul_rn(rc, cos(ang)), __dmul_rn(sr, ux));
        y = __dadd_rn(__dmul_rn(rc, sin(ang)), __dmul_rn(sr, uy));
        phi = (ic == 3) ? theta_all : (ic == 4) ? fmod(__dadd_rn(ang, SBC_PI_D / 2.), two_pi) : __dmul_rn(two_pi, up);
        agent = ic_permute(i, (uint32_t)N, sbc_draw(P, rep, SBC_IC_SWARM_ID, SBC_IC_STEP, 17u));
    } else {                                         // :357-366
        x = __dmul_rn(L, ux); y = __dmul_rn(L, uy); phi = __dmul_rn(two_pi, up);
    }
    double vx = cos(phi), vy = sin(phi);           // :368-378 unit speed
    if (P.BC != -1) ic_boundary(x, y, vx, vy, phi, L, P.BC);
    const size_t g = (size_t)rep * N + agent;
    st.d[0][g] = x; st.d[1][g] = y; st.d[2][g] = vx; st.d[3][g] = vy; st.d[4][g] = phi; st.d[5][g] = 1.0;
}

}  // namespace

void sbc_launch_init_cond(const DevParams &P, const DevState &S, int ic, double rep_range, const StateStage &st, cudaStream_t stream) {
    const size_t total = (size_t)S.N * S.R;
    init_cond_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, S.N, S.R, ic, rep_range, st);
}

void sbc_launch_pack_state(const DevParams &P, const DevState &S, const StateStage &st, unsigned mask, int have_prev,
                           cudaStream_t stream) {
    const size_t total = (size_t)S.N * S.R;
    cudaMemsetAsync(S.n_dead, 0, (size_t)S.R * sizeof(int), stream);
    pack_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, S, st, mask, have_prev);
}
void sbc_launch_unpack_state(const DevParams &P, const DevState &S, const StateStage &st, unsigned mask, cudaStream_t stream) {
    const size_t total = (size_t)S.N * S.R;
    unpack_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, S, st, mask);
}
void sbc_launch_particles(const DevState &S, double *out_dev, uint8_t *alive_dev, int f32, cudaStream_t stream) {
    const size_t n = (size_t)S.N * S.R * 7;
    particles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(S, out_dev, alive_dev, f32);
}
