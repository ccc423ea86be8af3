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
