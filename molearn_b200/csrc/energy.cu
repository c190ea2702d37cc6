// B200 (sm_100a) kernels and C ABI for the molearn physics hot path: bonded bond/angle/torsion
// energy + analytic gradient, tiled O(N^2) soft non-bonded energy + force, RMSD reducer.
//
// Replaces the ATen op chains of TorchProteinEnergy._roll_bond_angle_torsion_loss
// (src/molearn/loss_functions/torch_protein_energy.py:276-305, ~500 launches) and _cdist_nb /
// _cdist_nb_full (:224-239, ~15 dense [B,N,N] passes) by one launch each.  FP32 on the FMA
// pipes; no tensor cores (not a dense contraction); O(N) memory.
//
// Bonded kernel (bandwidth / issue bound)
//   CTA = (atom tile, chunk of conformations).  The tile's term records live in REGISTERS for
//   the whole chunk, so topology is read once per CTA, not once per conformation.  Per
//   conformation: coordinates of the tile (+halo) are staged in shared memory as float4, each
//   thread evaluates its terms and writes every role's gradient to a private shared-memory slot;
//   slots are laid out so that each atom's incoming contributions are contiguous, and one thread
//   per atom sums its range and writes the gradient once - no atomics.  Terms that straddle a
//   tile boundary are evaluated by both tiles (energy counted by the owner of role 0).
//
// Non-bonded kernel (FP32 FMA bound)
//   Work item = (conformation, 128x128 atom tile pair J>=I), one warp each.  A lane owns NI=4
//   i-atoms in registers; j-atoms are broadcast from shared memory 4 at a time; Newton's third
//   law is used, the j-forces being reduced across the warp by a halving exchange and added to
//   the gradient with RED.  Tile pairs that contain excluded (1-2/1-3/1-4) pairs or the diagonal
//   carry a 128x128 bit mask applied in registers; all other tiles run a fast path (d > 1.9 A
//   everywhere: w(d,k) = d) guarded by a min-distance test per micro-tile.
#include "../../include/molearn_b200.h"
#include "topology.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------
constexpr int BT = 256;          // bonded: threads per CTA
constexpr int RB = 2, RA = 2, RT = 2;  // bonded: max bonds / angles / torsions per thread
constexpr int NF = 4;            // torsion Fourier orders kept (ff14SB periods are 1..4)
constexpr int NB_T = 128;        // non-bonded tile edge (atoms)
constexpr int NB_NI = 4;         // i-atoms per lane
constexpr int NB_WARPS = 4;      // warps per CTA, one work item each
#ifndef NB_MINB
#define NB_MINB 3
#endif
constexpr float LJ_K = 1.9f, C_K = 0.4f;  // warp thresholds, torch_protein_energy.py:233-234
constexpr float CLAMP = 0.999999f;        // acos clamp, :289

struct BondRec { uint32_t ij, slots; float r0, k; };                       // 16 B
struct AngleRec { uint32_t ij, k, s_ij, s_k; float th0, kf, pad0, pad1; }; // 32 B
struct TorsRec { uint32_t ij, kl, s_ij, s_kl; float c0, cg[NF], sg[NF], pad[3]; };  // 64 B
static_assert(sizeof(BondRec) == 16 && sizeof(AngleRec) == 32 && sizeof(TorsRec) == 64, "record layout");
constexpr uint32_t OWNER = 0x8000u;  // bit 15 of the first local index: this tile counts the energy

struct BTile {
    int a0, n_own, n_halo, halo_off;
    int bond_off, n_bond, angle_off, n_angle, tors_off, n_tors;
    int csr_off, n_slots;
};

struct mlp_energy {
    int device = 0;
    uint32_t nb_mode = 0;
    int N = 0, Npad = 0;
    // bonded
    int n_tiles = 0, max_slots = 0, max_atoms = 0;
    BTile* d_tiles = nullptr;
    int* d_halo = nullptr;
    BondRec* d_bonds = nullptr;
    AngleRec* d_angles = nullptr;
    TorsRec* d_tors = nullptr;
    int* d_csr = nullptr;
    size_t bonded_smem = 0;
    // non-bonded
    int n_nbt = 0;                 // tiles per edge
    int n_pairs = 0, n_masked = 0; // tile pairs (masked ones first)
    int2* d_pairs = nullptr;       // (I,J)
    uint32_t* d_masks = nullptr;   // [n_masked][128][4]
    float* d_q = nullptr;          // [Npad]
    float2* d_par = nullptr;       // [Npad] (R/2, sqrt(12 eps))
    float* d_ones = nullptr;       // [4] default weights
    int last_launches = 0;
    int sm_count = 148;
    // optional per-kernel timing (mlp_energy_profile): events on the launching stream
    bool profile = false;
    cudaEvent_t ev[8] = {};
    std::vector<std::array<cudaEvent_t, 2>> pending[4];  // per kernel: (start, stop) pairs not read yet
    double prof_ms[4] = {0, 0, 0, 0};
    long prof_n[4] = {0, 0, 0, 0};
    std::vector<void*> allocs;
};

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t err__ = (call);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            mlp_set_error(std::string(#call) + ": " + cudaGetErrorString(err__));        \
            return MLP_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(const float4& a) { return make_float3(a.x, a.y, a.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float4 f4(float3 a) { return make_float4(a.x, a.y, a.z, 0.f); }
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// bonded kernel
// ------------------------------------------------------------------------------------------
// E = K (|ri-rj| - r0)^2                      (torch_protein_energy.py:282-284)
__device__ __forceinline__ float bond_term(const float4* sc, float4* slots, const BondRec& r, float w, bool grad, float& e) {
    const uint32_t i = r.ij & 0x7fffu, j = r.ij >> 16;
    float3 v = f3(sc[i]) - f3(sc[j]);
    float d2 = dot(v, v);
    float d = sqrtf(d2);
    float dl = d - r.r0;
    float en = r.k * dl * dl;
    if (r.ij & OWNER) e += en;
    if (grad) {
        float c = d > 0.f ? w * 2.f * r.k * dl / d : 0.f;  // autograd: d|v|/dv = 0 at v = 0
        float3 g = c * v;
        slots[r.slots & 0xffffu] = f4(g);
        slots[r.slots >> 16] = f4(-1.f * g);
    }
    return en;
}

// theta = acos(clamp(v1.v2/(|v1||v2|))), E = K (theta-theta0)^2      (:286-293)
__device__ __forceinline__ void angle_term(const float4* sc, float4* slots, const AngleRec& r, float w, bool grad, float& e) {
    const uint32_t i = r.ij & 0x7fffu, j = r.ij >> 16, k = r.k;
    float3 rj = f3(sc[j]);
    float3 v1 = f3(sc[i]) - rj, v2 = f3(sc[k]) - rj;
    float n1 = dot(v1, v1), n2 = dot(v2, v2);
    float inv12 = 1.0f / (sqrtf(n1) * sqrtf(n2));
    float c = dot(v1, v2) * inv12;
    float cc = fminf(fmaxf(c, -CLAMP), CLAMP);
    float th = acosf(cc);
    if (c != c) th = c;  // NaN propagates (fminf/fmaxf would swallow it)
    float dt = th - r.th0;
    if (r.ij & OWNER) e += r.kf * dt * dt;
    if (grad) {
        // dtheta/dv1 = v1 x (v1 x v2) / (|v1|^2 |v1 x v2|), dtheta/dv2 = -v2 x (v1 x v2) / (|v2|^2 |v1 x v2|):
        // the cross-product form of -1/sqrt(1-c^2) * dc/dv, which does not cancel near c = +-1.
        // torch.clamp passes no gradient outside the clamp, so the term is dropped there.
        float3 n = cross(v1, v2);
        float nn = dot(n, n);
        float coef = (c > -CLAMP && c < CLAMP) ? w * 2.f * r.kf * dt * rsqrtf(nn) : 0.f;
        if (c != c) coef = c;
        float3 gi = (coef / n1) * cross(v1, n);
        float3 gk = (-coef / n2) * cross(v2, n);
        slots[r.s_ij & 0xffffu] = f4(gi);
        slots[r.s_k] = f4(gk);
        slots[r.s_ij >> 16] = f4(-1.f * (gi + gk));
    }
}

// phi = atan2(b2.(c32 x c12), |b2| c12.c32), E = sum_p (V/f)(1+cos(n phi - gamma))   (:295-304)
// evaluated as a Fourier series in (cos phi, sin phi) - no atan2 / cos calls.
__device__ __forceinline__ void torsion_term(const float4* sc, float4* slots, const TorsRec& r, float w, bool grad, float& e) {
    const uint32_t i = r.ij & 0x7fffu, j = r.ij >> 16, k = r.kl & 0xffffu, l = r.kl >> 16;
    float3 rj = f3(sc[j]), rk = f3(sc[k]);
    float3 F = f3(sc[i]) - rj, G = rj - rk, H = f3(sc[l]) - rk;
    float3 A = cross(F, G), B = cross(H, G);
    float G2 = dot(G, G);
    float Gn = sqrtf(G2);
    float X = dot(A, B), Y = -Gn * dot(B, F);
    float rho2 = X * X + Y * Y;
    float irho = rsqrtf(rho2);
    // atan2(0,0) = 0 in the reference's forward
    float c1 = rho2 > 0.f ? X * irho : 1.f, s1 = rho2 > 0.f ? Y * irho : 0.f;
    float cn = c1, sn = s1;
    float en = r.c0, de = 0.f;  // de = dE/dphi
#pragma unroll
    for (int n = 1; n <= NF; ++n) {
        en = fmaf(r.cg[n - 1], cn, fmaf(r.sg[n - 1], sn, en));
        de = fmaf((float)n, fmaf(r.sg[n - 1], cn, -r.cg[n - 1] * sn), de);
        float c2 = cn * c1 - sn * s1;
        sn = fmaf(sn, c1, cn * s1);
        cn = c2;
    }
    if (r.ij & OWNER) e += en;
    if (grad) {
        de *= w;
        float A2 = dot(A, A), B2 = dot(B, B);
        float iA2 = 1.f / A2, iB2 = 1.f / B2, iG = 1.f / Gn;
        float3 gi = (-de * Gn * iA2) * A;
        float3 gl = (de * Gn * iB2) * B;
        float fa = de * dot(F, G) * iA2 * iG, hb = de * dot(H, G) * iB2 * iG;
        float3 sv = fa * A - hb * B;
        slots[r.s_ij & 0xffffu] = f4(gi);
        slots[r.s_kl >> 16] = f4(gl);
        slots[r.s_ij >> 16] = f4(sv - gi);
        slots[r.s_kl & 0xffffu] = f4(-1.f * (gl + sv));
    }
}

template <bool GRAD>
__global__ void __launch_bounds__(BT, 2)
bonded_kernel(const BTile* __restrict__ tiles, const int* __restrict__ halo, const BondRec* __restrict__ bonds,
              const AngleRec* __restrict__ angles, const TorsRec* __restrict__ tors, const int* __restrict__ csr,
              const float* __restrict__ x, long sxb, long sxc, long sxn,
              float* __restrict__ grad, long sgb, long sgc, long sgn,
              const float* __restrict__ weights, double* __restrict__ acc,
              int B, int conf_per_cta, uint32_t term_mask, int accumulate, int max_atoms) {
    extern __shared__ float4 smem[];
    float4* sc = smem;                 // [max_atoms] coordinates of own + halo atoms
    float4* slots = smem + max_atoms;  // [n_slots + 1] gradient slots (last = dump)
    __shared__ float s_e[3];

    const BTile t = tiles[blockIdx.x];
    const int tid = threadIdx.x;
    const int n_at = t.n_own + t.n_halo;

    // term records -> registers, once per CTA
    BondRec rb[RB];
    AngleRec ra[RA];
    TorsRec rt[RT];
#pragma unroll
    for (int r = 0; r < RB; ++r)
        if (tid + r * BT < t.n_bond) rb[r] = bonds[t.bond_off + tid + r * BT];
#pragma unroll
    for (int r = 0; r < RA; ++r)
        if (tid + r * BT < t.n_angle) ra[r] = angles[t.angle_off + tid + r * BT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
        if (tid + r * BT < t.n_tors) rt[r] = tors[t.tors_off + tid + r * BT];
    int s_lo = 0, s_hi = 0;
    if (GRAD && tid < t.n_own) { s_lo = csr[t.csr_off + tid]; s_hi = csr[t.csr_off + tid + 1]; }

    // element e of the tile's coordinate block -> (local atom, component, global atom)
    const int n_el = 3 * n_at;
    const bool aos = (sxc == 1);
    auto elem = [&](int e, int& la, int& c) {
        if (e < 3 * t.n_own) {
            if (aos) { la = e / 3; c = e - 3 * la; } else { c = e / t.n_own; la = e - c * t.n_own; }
        } else {
            int h = e - 3 * t.n_own;
            la = t.n_own + h / 3; c = h - 3 * (h / 3);
        }
    };
    auto gatom = [&](int la) { return la < t.n_own ? t.a0 + la : halo[t.halo_off + la - t.n_own]; };

    const int b0 = blockIdx.y * conf_per_cta;
    const int b1 = min(B, b0 + conf_per_cta);
    constexpr int PF = 4;  // prefetch registers: 3*(256+halo)/256 <= 4 elements per thread
    float pf[PF];
    auto prefetch = [&](int b) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            int e = tid + q * BT;
            if (e < n_el) {
                int la, c; elem(e, la, c);
                pf[q] = x[(long)b * sxb + c * sxc + (long)gatom(la) * sxn];
            }
        }
    };

    if (b0 < b1) prefetch(b0);
    for (int b = b0; b < b1; ++b) {
        float wb = 1.f, wa = 1.f, wt = 1.f;
        if (weights) { wb = weights[4 * b]; wa = weights[4 * b + 1]; wt = weights[4 * b + 2]; }
        const bool do_b = term_mask & MLP_TERM_BOND, do_a = term_mask & MLP_TERM_ANGLE, do_t = term_mask & MLP_TERM_TORSION;
        if (GRAD && accumulate && !acc && (!do_b || wb == 0.f) && (!do_a || wa == 0.f) && (!do_t || wt == 0.f)) {
            // backward correction with nothing to add for this conformation (CTA-uniform)
            if (b + 1 < b1) prefetch(b + 1);
            continue;
        }
        // stage coordinates (previous iteration's term phase is over: all threads passed the sync below)
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            int e = tid + q * BT;
            if (e < n_el) { int la, c; elem(e, la, c); reinterpret_cast<float*>(&sc[la])[c] = pf[q]; }
        }
        for (int e = tid + PF * BT; e < n_el; e += BT) {
            int la, c; elem(e, la, c);
            reinterpret_cast<float*>(&sc[la])[c] = x[(long)b * sxb + c * sxc + (long)gatom(la) * sxn];
        }
        if (tid < 3) s_e[tid] = 0.f;
        __syncthreads();
        if (b + 1 < b1) prefetch(b + 1);

        float eb = 0.f, ea = 0.f, et = 0.f;
        if (do_b) {
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (tid + r * BT < t.n_bond) bond_term(sc, slots, rb[r], wb, GRAD, eb);
        }
        if (do_a) {
#pragma unroll
            for (int r = 0; r < RA; ++r)
                if (tid + r * BT < t.n_angle) angle_term(sc, slots, ra[r], wa, GRAD, ea);
        }
        if (do_t) {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (tid + r * BT < t.n_tors) torsion_term(sc, slots, rt[r], wt, GRAD, et);
        }
        if (GRAD && term_mask != MLP_TERM_BONDED) {
            // a masked-out term type leaves its slots unwritten: zero them
            // (rare path: backward corrections with a subset of terms)
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (!do_b && tid + r * BT < t.n_bond) { slots[rb[r].slots & 0xffffu] = make_float4(0, 0, 0, 0); slots[rb[r].slots >> 16] = make_float4(0, 0, 0, 0); }
#pragma unroll
            for (int r = 0; r < RA; ++r)
                if (!do_a && tid + r * BT < t.n_angle) {
                    slots[ra[r].s_ij & 0xffffu] = make_float4(0, 0, 0, 0); slots[ra[r].s_ij >> 16] = make_float4(0, 0, 0, 0);
                    slots[ra[r].s_k] = make_float4(0, 0, 0, 0);
                }
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (!do_t && tid + r * BT < t.n_tors) {
                    slots[rt[r].s_ij & 0xffffu] = make_float4(0, 0, 0, 0); slots[rt[r].s_ij >> 16] = make_float4(0, 0, 0, 0);
                    slots[rt[r].s_kl & 0xffffu] = make_float4(0, 0, 0, 0); slots[rt[r].s_kl >> 16] = make_float4(0, 0, 0, 0);
                }
        }
        eb = warp_sum(eb); ea = warp_sum(ea); et = warp_sum(et);
        if ((tid & 31) == 0) { atomicAdd(&s_e[0], eb); atomicAdd(&s_e[1], ea); atomicAdd(&s_e[2], et); }
        __syncthreads();
        if (tid < 3 && acc) atomicAdd(&acc[4 * (long)b + tid], (double)s_e[tid]);
        if (GRAD && tid < t.n_own) {
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (int s = s_lo; s < s_hi; ++s) { float4 v = slots[s]; gx += v.x; gy += v.y; gz += v.z; }
            float* g = grad + (long)b * sgb + (long)(t.a0 + tid) * sgn;
            if (accumulate) { g[0] += gx; g[sgc] += gy; g[2 * sgc] += gz; }
            else { g[0] = gx; g[sgc] = gy; g[2 * sgc] = gz; }
        }
        // next iteration's coordinate staging only touches sc (term phase done); its slot
        // writes happen after the next __syncthreads, i.e. after every gather above.
    }
}

// ------------------------------------------------------------------------------------------
// non-bonded kernels
// ------------------------------------------------------------------------------------------
// gather x (any strides) into [B][Npad] float4 (x,y,z,q); padding atoms are parked far away with
// zero charge / epsilon so that they contribute exactly 0 without any test in the pair loop.
__global__ void nb_pack_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, const float* __restrict__ q,
                               float4* __restrict__ xq, int N, int Npad, int B) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)B * Npad) return;
    int b = (int)(idx / Npad), n = (int)(idx - (long)b * Npad);
    float4 v;
    if (n < N) {
        const float* p = x + (long)b * sxb + (long)n * sxn;
        v = make_float4(p[0], p[sxc], p[2 * sxc], q[n]);
    } else {
        v = make_float4(1.0e6f + 100.f * (float)(n - N), -1.0e6f, 1.0e6f, 0.f);
    }
    xq[idx] = v;
}

__device__ __forceinline__ void warp_fn(float d, float k, float& w, float& dw) {
    // w(d,k) = elu(d-k)+k  (torch_protein_energy.py:238-239), dw = dw/dd
    if (d > k) { w = d; dw = 1.f; }
    else { float ex = expf(d - k); w = ex - 1.f + k; dw = ex; }
}

// Squared distance below which a pair leaves the fast path (w(d,1.9) = d needs d > 1.9 A; the
// margin covers the 2-ulp rsqrt).  Pairs closer than this are evaluated by the fast path at the
// clamped distance sqrt(NB_THR2) - a small, finite value - and then corrected by nb_pair_fix().
constexpr float NB_THR2 = LJ_K * LJ_K * 1.0001f;
constexpr float NB_MASKED_R2 = 1.0e4f;  // stand-in r^2 of statically excluded pairs (their parameters are zeroed)

// Fast-path pair term at squared distance r2c >= NB_THR2.  Returns s = -(dE/dd)/d; adds 12*E_lj to
// Sl and E_coulomb to Sc.
template <bool FULL>
__device__ __forceinline__ float nb_pair_fast(float r2c, float Rij, float e12, float qq, float& Sl, float& Sc) {
    float inv = rsqrt_fast(r2c);
    float t = Rij * inv;
    float t2 = t * t;
    float t6 = t2 * t2 * t2;
    float u6 = e12 * t6;
    float ec = qq * inv;
    float u;
    if (FULL) { u = fmaf(u6, t6 - 1.f, ec); Sl = fmaf(u6, t6 - 2.f, Sl); }
    else { u = fmaf(u6, t6, ec); Sl = fmaf(u6, t6, Sl); }
    Sc = fmaf(qq, inv, Sc);
    return u * (inv * inv);
}

// Exact pair term (any distance), minus what the fast path added for this pair at the clamped
// distance.  Rare path (clashing structures): kept out of line to protect the hot loop's registers
// and instruction cache.  Returns (ds, dSl, dSc).
template <bool FULL>
__device__ __noinline__ float3 nb_pair_fix(float r2, float Rij, float e12, float qq) {
    float Sl_f = 0.f, Sc_f = 0.f;
    float s_f = nb_pair_fast<FULL>(NB_THR2, Rij, e12, qq, Sl_f, Sc_f);
    // the fast path multiplied s by the true (dx,dy,dz); so must the correction
    float d = sqrtf(r2);
    float dinv = r2 > 0.f ? 1.f / d : 0.f;  // coincident atoms: zero force (cdist's backward does the same)
    float w1, dw1, w2, dw2;
    warp_fn(d, LJ_K, w1, dw1);
    warp_fn(d, C_K, w2, dw2);
    float iw1 = 1.f / w1, iw2 = 1.f / w2;
    float t = Rij * iw1;
    float t2 = t * t;
    float t6 = t2 * t2 * t2;
    float u6 = e12 * t6;
    float ec = qq * iw2;
    float ulj, sl;
    if (FULL) { ulj = u6 * (t6 - 1.f); sl = u6 * (t6 - 2.f); }
    else { ulj = u6 * t6; sl = ulj; }
    float s = (ulj * iw1 * dw1 + ec * iw2 * dw2) * dinv;
    return make_float3(s - s_f, sl - Sl_f, ec - Sc_f);
}

// One micro-tile: NB_NI i-atoms (registers) x 4 j-atoms, fast path with clamped distance.
// m16: bit (4a+b) = pair allowed (MASKED tiles).  Returns the smallest true r^2 of the allowed pairs.
template <bool MASKED, bool FULL>
__device__ __forceinline__ float nb_micro(const float4 (&xi)[NB_NI], const float2 (&pi)[NB_NI], const float4 (&xj)[4],
                                          const float2 (&pj)[4], uint32_t m16,
                                          float3 (&Fi)[NB_NI], float3 (&Fj)[4], float& Sl, float& Sc) {
    float rmin = 1e30f;
#pragma unroll
    for (int a = 0; a < NB_NI; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float dx = xi[a].x - xj[b].x, dy = xi[a].y - xj[b].y, dz = xi[a].z - xj[b].z;
            float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            float e12 = pi[a].y * pj[b].y;
            float qq = xi[a].w * xj[b].w;
            float Rij = pi[a].x + pj[b].x;
            if (MASKED) {
                bool on = (m16 >> (4 * a + b)) & 1u;
                r2 = on ? r2 : NB_MASKED_R2;
                e12 = on ? e12 : 0.f;
                qq = on ? qq : 0.f;
            }
            rmin = fminf(rmin, r2);
            float s = nb_pair_fast<FULL>(fmaxf(r2, NB_THR2), Rij, e12, qq, Sl, Sc);
            Fi[a].x = fmaf(s, dx, Fi[a].x); Fi[a].y = fmaf(s, dy, Fi[a].y); Fi[a].z = fmaf(s, dz, Fi[a].z);
            Fj[b].x = fmaf(s, dx, Fj[b].x); Fj[b].y = fmaf(s, dy, Fj[b].y); Fj[b].z = fmaf(s, dz, Fj[b].z);
        }
    }
    return rmin;
}

// Correction pass over a micro-tile that holds at least one pair closer than sqrt(NB_THR2).
template <bool MASKED, bool FULL>
__device__ __forceinline__ void nb_micro_fix(const float4 (&xi)[NB_NI], const float2 (&pi)[NB_NI], const float4 (&xj)[4],
                                             const float2 (&pj)[4], uint32_t m16,
                                             float3 (&Fi)[NB_NI], float3 (&Fj)[4], float& Sl, float& Sc) {
    // Launder the operands: without this the compiler shares dx/r2/... with the hot micro-tile
    // and keeps 60+ values alive (spilled to local memory) just in case this path runs.
    float4 yi[NB_NI], yj[4];
#pragma unroll
    for (int a = 0; a < NB_NI; ++a) {
        yi[a] = xi[a];
        asm volatile("" : "+f"(yi[a].x), "+f"(yi[a].y), "+f"(yi[a].z), "+f"(yi[a].w));
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        yj[b] = xj[b];
        asm volatile("" : "+f"(yj[b].x), "+f"(yj[b].y), "+f"(yj[b].z), "+f"(yj[b].w));
    }
#pragma unroll
    for (int a = 0; a < NB_NI; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float dx = yi[a].x - yj[b].x, dy = yi[a].y - yj[b].y, dz = yi[a].z - yj[b].z;
            float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            bool on = !MASKED || ((m16 >> (4 * a + b)) & 1u);
            // NaN coordinates: !(r2 >= thr) is true -> exact path -> NaN propagates
            if (on && !(r2 >= NB_THR2)) {
                float3 c = nb_pair_fix<FULL>(r2, pi[a].x + pj[b].x, pi[a].y * pj[b].y, yi[a].w * yj[b].w);
                Fi[a].x = fmaf(c.x, dx, Fi[a].x); Fi[a].y = fmaf(c.x, dy, Fi[a].y); Fi[a].z = fmaf(c.x, dz, Fi[a].z);
                Fj[b].x = fmaf(c.x, dx, Fj[b].x); Fj[b].y = fmaf(c.x, dy, Fj[b].y); Fj[b].z = fmaf(c.x, dz, Fj[b].z);
                Sl += c.y; Sc += c.z;
            }
        }
    }
}

// One warp = one (conformation, tile pair).  Lane L owns i-atoms i0+L+32a (a<4) for the whole
// tile; the 32 groups of 4 j-atoms rotate through the lanes (lane L works on group (L+r)%32 at
// step r) together with their force accumulators, so Newton's third law costs 12 shuffles per
// 16 pairs and no reduction.
template <bool GRAD, bool FULL>
__global__ void __launch_bounds__(NB_WARPS * 32, NB_MINB)
nb_kernel(const float4* __restrict__ xq, const float2* __restrict__ par, const int2* __restrict__ pairs,
          const uint32_t* __restrict__ masks, int n_pairs, int n_masked, int Npad, int N, int B,
          float* __restrict__ grad, long sgb, long sgc, long sgn, const float* __restrict__ weights,
          double* __restrict__ acc) {
    // j tile, group-minor: atom 4g+b lives at [b*32+g] so that lanes (different g) never conflict
    __shared__ float4 s_xj[NB_WARPS][NB_T];
    __shared__ float2 s_pj[NB_WARPS][NB_T];
    __shared__ uint32_t s_m[NB_WARPS][NB_T * 4];   // uint16 [g][lane] pair masks, as words
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long item = (long)blockIdx.x * NB_WARPS + warp;
    if (item >= (long)B * n_pairs) return;
    const int b = (int)(item / n_pairs), p = (int)(item - (long)b * n_pairs);
    const float wnb = weights ? weights[4 * b + 3] : 1.f;
    if (weights && wnb == 0.f && acc == nullptr) return;  // backward correction with nothing to add
    const int2 IJ = pairs[p];
    const int i0 = IJ.x * NB_T, j0 = IJ.y * NB_T;
    const float4* xb = xq + (long)b * Npad;
    const bool masked = p < n_masked;

    float4 xi[NB_NI];
    float2 pi[NB_NI];
    float3 Fi[NB_NI];
#pragma unroll
    for (int a = 0; a < NB_NI; ++a) {
        xi[a] = xb[i0 + lane + 32 * a];
        pi[a] = par[i0 + lane + 32 * a];
        Fi[a] = make_float3(0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < NB_T / 32; ++k) {
        int jj = lane + 32 * k;
        int slot = (jj & 3) * 32 + (jj >> 2);
        s_xj[warp][slot] = xb[j0 + jj];
        s_pj[warp][slot] = par[j0 + jj];
    }
    if (masked) {
        const uint4* src = reinterpret_cast<const uint4*>(masks + (long)p * (NB_T * 4));
        uint4* dst = reinterpret_cast<uint4*>(s_m[warp]);
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[lane + 32 * k] = src[lane + 32 * k];
    }
    __syncwarp();

    float Sl = 0.f, Sc = 0.f;
    float3 Fj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) Fj[q] = make_float3(0.f, 0.f, 0.f);
    const int src_lane = (lane + 1) & 31;
    const uint16_t* sm16 = reinterpret_cast<const uint16_t*>(s_m[warp]);

#pragma unroll 1
    for (int r = 0; r < 32; ++r) {
        const int g = (lane + r) & 31;
        float4 xj[4];
        float2 pj[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { xj[q] = s_xj[warp][q * 32 + g]; pj[q] = s_pj[warp][q * 32 + g]; }
        if (masked) {
            const uint32_t m16 = sm16[g * 32 + lane];
            float rmin = nb_micro<true, FULL>(xi, pi, xj, pj, m16, Fi, Fj, Sl, Sc);
            if (!(rmin >= NB_THR2)) nb_micro_fix<true, FULL>(xi, pi, xj, pj, m16, Fi, Fj, Sl, Sc);
        } else {
            float rmin = nb_micro<false, FULL>(xi, pi, xj, pj, 0xffffu, Fi, Fj, Sl, Sc);
            if (!(rmin >= NB_THR2)) nb_micro_fix<false, FULL>(xi, pi, xj, pj, 0xffffu, Fi, Fj, Sl, Sc);
        }
        if (GRAD) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                Fj[q].x = __shfl_sync(0xffffffffu, Fj[q].x, src_lane);
                Fj[q].y = __shfl_sync(0xffffffffu, Fj[q].y, src_lane);
                Fj[q].z = __shfl_sync(0xffffffffu, Fj[q].z, src_lane);
            }
        }
    }
    if (GRAD) {
        // after 32 rotations lane L holds the finished forces of group L (atoms j0+4L+q);
        // grad_i = -sum s*(ri-rj), grad_j = +sum s*(ri-rj)   (s = -(dE/dd)/d)
        float* gb = grad + (long)b * sgb;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int ja = j0 + 4 * lane + q;
            if (ja < N) {
                float* g = gb + (long)ja * sgn;
                atomicAdd(g, wnb * Fj[q].x); atomicAdd(g + sgc, wnb * Fj[q].y); atomicAdd(g + 2 * sgc, wnb * Fj[q].z);
            }
        }
#pragma unroll
        for (int a = 0; a < NB_NI; ++a) {
            int ia = i0 + lane + 32 * a;
            if (ia < N) {
                float* g = gb + (long)ia * sgn;
                atomicAdd(g, -wnb * Fi[a].x); atomicAdd(g + sgc, -wnb * Fi[a].y); atomicAdd(g + 2 * sgc, -wnb * Fi[a].z);
            }
        }
    }
    if (acc) {
        // E = sum_lj12/12 + sum_coulomb
        double e = (double)warp_sum(Sl) * (1.0 / 12.0) + (double)warp_sum(Sc);
        if (lane == 0) atomicAdd(&acc[4 * (long)b + 3], e);
    }
}

__global__ void zero_grad_kernel(float* __restrict__ grad, long sgb, long sgc, long sgn, int N, int B) {
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)B * N * 3) return;
    int c = (int)(idx % 3);
    long r = idx / 3;
    int n = (int)(r % N), b = (int)(r / N);
    grad[(long)b * sgb + c * sgc + (long)n * sgn] = 0.f;
}

__global__ void finalize_kernel(const double* __restrict__ acc, float* __restrict__ energies, int n) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) energies[idx] = (float)acc[idx];
}

// RMSD of x*scale+shift to K targets: out[b,k] = sqrt(sum_{n,c} d^2 / N)
__global__ void __launch_bounds__(256)
rmsd_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, float scale, float shift,
            const float* __restrict__ targets, int N, int K, float* __restrict__ out) {
    const int b = blockIdx.x, k = blockIdx.y;
    const float* t = targets + (long)k * N * 3;
    float s = 0.f;
    for (int e = threadIdx.x; e < 3 * N; e += blockDim.x) {
        int n = e / 3, c = e - 3 * n;
        float v = fmaf(x[(long)b * sxb + c * sxc + (long)n * sxn], scale, shift) - t[e];
        s = fmaf(v, v, s);
    }
    __shared__ float red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[(long)b * K + k] = sqrtf(v / (float)N);
    }
}

// ------------------------------------------------------------------------------------------
// host: packing
// ------------------------------------------------------------------------------------------
namespace {

template <class T> int upload(mlp_energy* E, T** dst, const std::vector<T>& v) {
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CK(cudaMalloc(reinterpret_cast<void**>(dst), bytes));
    E->allocs.push_back(*dst);
    if (!v.empty()) CK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MLP_OK;
}

struct Term { int type; int idx; };  // type 0 bond, 1 angle, 2 torsion

int pack_bonded(mlp_energy* E, const mlp_topology& T, int tile_atoms) {
    const int N = T.n_atoms;
    const int nb = (int)T.bonds.size() / 2, na = (int)T.angles.size() / 3, nt = (int)T.torsions.size() / 4, P = T.P;
    // torsion Fourier coefficients from the fp32-rounded parameters (the reference moves its
    // tables to the device with .float(), torch_protein_energy.py:123)
    std::vector<char> t_live(nt, 0);
    std::vector<std::array<float, 1 + 2 * NF>> t_coef(nt);
    for (int k = 0; k < nt; ++k) {
        double c0 = 0, cg[NF] = {0}, sg[NF] = {0};
        bool live = false;
        for (int s = 0; s < P; ++s) {
            const double* p = &T.torsion_para[(size_t)k * 4 * P];
            double f = (double)(float)p[s], v = (double)(float)p[P + s], ph = (double)(float)p[2 * P + s], pe = (double)(float)p[3 * P + s];
            double amp = v / f;
            if (amp == 0.0) continue;
            int n = (int)std::lround(pe);
            if (std::fabs(pe - n) > 1e-6 || n < 0 || n > NF) {
                mlp_set_error("torsion period " + std::to_string(pe) + " is not an integer in [0," + std::to_string(NF) + "]: unsupported by the sm_100a kernel");
                return MLP_E_ARG;
            }
            live = true;
            c0 += amp;
            if (n == 0) { c0 += amp * std::cos(ph); continue; }  // cos(0*phi - gamma)
            cg[n - 1] += amp * std::cos(ph);
            sg[n - 1] += amp * std::sin(ph);
        }
        t_live[k] = live;
        t_coef[k][0] = (float)c0;
        for (int n = 0; n < NF; ++n) { t_coef[k][1 + n] = (float)cg[n]; t_coef[k][1 + NF + n] = (float)sg[n]; }
    }

    std::vector<BTile> tiles;
    std::vector<int> halo_all, csr_all;
    std::vector<BondRec> bonds_all;
    std::vector<AngleRec> angles_all;
    std::vector<TorsRec> tors_all;
    // per-atom incident terms
    std::vector<std::vector<Term>> inc(N);
    auto touch = [&](int type, int idx, const int32_t* a, int n) {
        for (int r = 0; r < n; ++r) {
            bool dup = false;
            for (int q = 0; q < r; ++q) dup = dup || a[q] == a[r];
            if (!dup) inc[a[r]].push_back({type, idx});
        }
    };
    for (int k = 0; k < nb; ++k) touch(0, k, &T.bonds[2 * k], 2);
    for (int k = 0; k < na; ++k) touch(1, k, &T.angles[3 * k], 3);
    for (int k = 0; k < nt; ++k) if (t_live[k]) touch(2, k, &T.torsions[4 * k], 4);

    int max_slots = 0, max_atoms = 0;
    for (int a0 = 0; a0 < N; a0 += tile_atoms) {
        const int n_own = std::min(tile_atoms, N - a0);
        BTile bt{};
        bt.a0 = a0; bt.n_own = n_own;
        // unique terms of the tile, in (type, index) order
        std::vector<int> tb, ta, tt;
        for (int a = a0; a < a0 + n_own; ++a)
            for (const Term& tm : inc[a]) (tm.type == 0 ? tb : tm.type == 1 ? ta : tt).push_back(tm.idx);
        auto uniq = [](std::vector<int>& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
        uniq(tb); uniq(ta); uniq(tt);
        if ((int)tb.size() > RB * BT || (int)ta.size() > RA * BT || (int)tt.size() > RT * BT) return 1;  // tile too dense
        // halo
        std::vector<int> halo;
        auto need = [&](int a) { if (a < a0 || a >= a0 + n_own) halo.push_back(a); };
        for (int k : tb) for (int r = 0; r < 2; ++r) need(T.bonds[2 * k + r]);
        for (int k : ta) for (int r = 0; r < 3; ++r) need(T.angles[3 * k + r]);
        for (int k : tt) for (int r = 0; r < 4; ++r) need(T.torsions[4 * k + r]);
        uniq(halo);
        if (n_own + (int)halo.size() >= 0x7fff) return 1;
        auto local = [&](int a) -> uint32_t {
            if (a >= a0 && a < a0 + n_own) return (uint32_t)(a - a0);
            return (uint32_t)(n_own + (std::lower_bound(halo.begin(), halo.end(), a) - halo.begin()));
        };
        // slots: contiguous per own atom, in the order (bonds, angles, torsions) x role
        std::vector<std::vector<std::array<int, 3>>> per_atom(n_own);  // (type, pos in tile list, role)
        for (size_t q = 0; q < tb.size(); ++q) for (int r = 0; r < 2; ++r) { int a = T.bonds[2 * tb[q] + r]; if (a >= a0 && a < a0 + n_own) per_atom[a - a0].push_back({0, (int)q, r}); }
        for (size_t q = 0; q < ta.size(); ++q) for (int r = 0; r < 3; ++r) { int a = T.angles[3 * ta[q] + r]; if (a >= a0 && a < a0 + n_own) per_atom[a - a0].push_back({1, (int)q, r}); }
        for (size_t q = 0; q < tt.size(); ++q) for (int r = 0; r < 4; ++r) { int a = T.torsions[4 * tt[q] + r]; if (a >= a0 && a < a0 + n_own) per_atom[a - a0].push_back({2, (int)q, r}); }
        std::vector<std::array<uint32_t, 2>> sb(tb.size());
        std::vector<std::array<uint32_t, 3>> sa(ta.size());
        std::vector<std::array<uint32_t, 4>> st(tt.size());
        int n_slots = 0;
        for (auto& v : per_atom) n_slots += (int)v.size();
        if (n_slots >= 0xffff) return 1;
        const uint32_t dump = (uint32_t)n_slots;
        for (auto& s : sb) s.fill(dump);
        for (auto& s : sa) s.fill(dump);
        for (auto& s : st) s.fill(dump);
        bt.csr_off = (int)csr_all.size();
        int slot = 0;
        for (int la = 0; la < n_own; ++la) {
            csr_all.push_back(slot);
            for (auto& e : per_atom[la]) {
                if (e[0] == 0) sb[e[1]][e[2]] = slot; else if (e[0] == 1) sa[e[1]][e[2]] = slot; else st[e[1]][e[2]] = slot;
                ++slot;
            }
        }
        csr_all.push_back(slot);
        bt.n_slots = n_slots;
        auto owner = [&](int first_atom) { return (first_atom >= a0 && first_atom < a0 + n_own) ? OWNER : 0u; };
        bt.bond_off = (int)bonds_all.size(); bt.n_bond = (int)tb.size();
        for (size_t q = 0; q < tb.size(); ++q) {
            const int k = tb[q];
            BondRec r{};
            r.ij = local(T.bonds[2 * k]) | owner(T.bonds[2 * k]) | (local(T.bonds[2 * k + 1]) << 16);
            r.slots = sb[q][0] | (sb[q][1] << 16);
            r.r0 = (float)T.bond_para[2 * k]; r.k = (float)T.bond_para[2 * k + 1];
            bonds_all.push_back(r);
        }
        bt.angle_off = (int)angles_all.size(); bt.n_angle = (int)ta.size();
        for (size_t q = 0; q < ta.size(); ++q) {
            const int k = ta[q];
            AngleRec r{};
            r.ij = local(T.angles[3 * k]) | owner(T.angles[3 * k]) | (local(T.angles[3 * k + 1]) << 16);
            r.k = local(T.angles[3 * k + 2]);
            r.s_ij = sa[q][0] | (sa[q][1] << 16); r.s_k = sa[q][2];
            r.th0 = (float)T.angle_para[2 * k]; r.kf = (float)T.angle_para[2 * k + 1];
            angles_all.push_back(r);
        }
        bt.tors_off = (int)tors_all.size(); bt.n_tors = (int)tt.size();
        for (size_t q = 0; q < tt.size(); ++q) {
            const int k = tt[q];
            TorsRec r{};
            r.ij = local(T.torsions[4 * k]) | owner(T.torsions[4 * k]) | (local(T.torsions[4 * k + 1]) << 16);
            r.kl = local(T.torsions[4 * k + 2]) | (local(T.torsions[4 * k + 3]) << 16);
            r.s_ij = st[q][0] | (st[q][1] << 16); r.s_kl = st[q][2] | (st[q][3] << 16);
            r.c0 = t_coef[k][0];
            for (int n = 0; n < NF; ++n) { r.cg[n] = t_coef[k][1 + n]; r.sg[n] = t_coef[k][1 + NF + n]; }
            tors_all.push_back(r);
        }
        bt.halo_off = (int)halo_all.size(); bt.n_halo = (int)halo.size();
        halo_all.insert(halo_all.end(), halo.begin(), halo.end());
        max_slots = std::max(max_slots, n_slots + 1);
        max_atoms = std::max(max_atoms, n_own + (int)halo.size());
        tiles.push_back(bt);
    }
    E->n_tiles = (int)tiles.size(); E->max_slots = max_slots; E->max_atoms = max_atoms;
    E->bonded_smem = (size_t)(max_slots + max_atoms) * sizeof(float4);
    int rc;
    if ((rc = upload(E, &E->d_tiles, tiles))) return rc;
    if ((rc = upload(E, &E->d_halo, halo_all))) return rc;
    if ((rc = upload(E, &E->d_bonds, bonds_all))) return rc;
    if ((rc = upload(E, &E->d_angles, angles_all))) return rc;
    if ((rc = upload(E, &E->d_tors, tors_all))) return rc;
    if ((rc = upload(E, &E->d_csr, csr_all))) return rc;
    return MLP_OK;
}

int pack_nonbonded(mlp_energy* E, const mlp_topology& T) {
    const int N = T.n_atoms;
    const int nt = (N + NB_T - 1) / NB_T;
    E->n_nbt = nt; E->Npad = nt * NB_T;
    std::vector<float> q(E->Npad, 0.f);
    std::vector<float2> par(E->Npad, make_float2(1.f, 0.f));
    for (int i = 0; i < N; ++i) {
        q[i] = (float)T.charges[i];
        // R_ij = |R_i+R_j|/2, eps_ij = sqrt(eps_i eps_j) (utils.py:814-815); 12 folded into eps
        par[i] = make_float2(0.5f * T.vdw_R[i], (float)std::sqrt(12.0 * (double)T.vdw_e[i]));
    }
    // exclusions per tile pair
    std::map<std::pair<int, int>, std::vector<std::pair<int, int>>> ex;
    for (size_t k = 0; k < T.excl.size() / 2; ++k) {
        int i = T.excl[2 * k], j = T.excl[2 * k + 1];
        ex[{i / NB_T, j / NB_T}].push_back({i, j});
    }
    for (int t = 0; t < nt; ++t) ex[{t, t}];  // diagonal tiles are always masked (j > i)
    std::vector<int2> pairs;
    std::vector<uint32_t> masks;
    for (auto& kv : ex) {
        const int I = kv.first.first, J = kv.first.second;
        pairs.push_back(make_int2(I, J));
        std::vector<uint16_t> m16((size_t)32 * 32, 0xffffu);
        auto clear = [&](int li, int lj) { m16[(size_t)(lj >> 2) * 32 + (li & 31)] &= (uint16_t)~(1u << (4 * (li >> 5) + (lj & 3))); };
        if (I == J)
            for (int li = 0; li < NB_T; ++li)
                for (int lj = 0; lj <= li; ++lj) clear(li, lj);
        for (auto& pr : kv.second) clear(pr.first - I * NB_T, pr.second - J * NB_T);
        std::vector<uint32_t> m((size_t)NB_T * 4);
        std::memcpy(m.data(), m16.data(), m.size() * sizeof(uint32_t));
        masks.insert(masks.end(), m.begin(), m.end());
    }
    E->n_masked = (int)pairs.size();
    for (int I = 0; I < nt; ++I)
        for (int J = I; J < nt; ++J)
            if (!ex.count({I, J})) pairs.push_back(make_int2(I, J));
    E->n_pairs = (int)pairs.size();
    int rc;
    if ((rc = upload(E, &E->d_pairs, pairs))) return rc;
    if ((rc = upload(E, &E->d_masks, masks))) return rc;
    if ((rc = upload(E, &E->d_q, q))) return rc;
    if ((rc = upload(E, &E->d_par, par))) return rc;
    std::vector<float> ones(4, 1.f);
    if ((rc = upload(E, &E->d_ones, ones))) return rc;
    return MLP_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int mlp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mlp_energy_create(const mlp_topology* t, int device, uint32_t nb_mode, mlp_energy** out) {
    if (!t || !out || nb_mode > MLP_NB_FULL) { mlp_set_error("mlp_energy_create: bad argument"); return MLP_E_ARG; }
    *out = nullptr;
    if (mlp_device_count() <= device || device < 0) {
        mlp_set_error("mlp_energy_create: CUDA device " + std::to_string(device) + " not available (there is no CPU fallback)");
        return MLP_E_NOGPU;
    }
    int prev = 0;
    CK(cudaGetDevice(&prev));
    CK(cudaSetDevice(device));
    auto* E = new mlp_energy();
    E->device = device; E->nb_mode = nb_mode; E->N = t->n_atoms;
    cudaDeviceGetAttribute(&E->sm_count, cudaDevAttrMultiProcessorCount, device);
    int rc = MLP_OK;
    int tile = BT;
    if (const char* env = std::getenv("MLP_BONDED_TILE")) tile = std::max(32, std::min(BT, std::atoi(env)));
    while (true) {
        rc = pack_bonded(E, *t, tile);
        if (rc == 1 && tile > 32) { tile /= 2; for (void* p : E->allocs) cudaFree(p); E->allocs.clear(); continue; }
        break;
    }
    if (rc == 1) { mlp_set_error("mlp_energy_create: topology too dense for the bonded tile"); rc = MLP_E_ARG; }
    if (rc == MLP_OK) rc = pack_nonbonded(E, *t);
    if (rc == MLP_OK) {
        cudaError_t e1 = cudaFuncSetAttribute(bonded_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E->bonded_smem);
        cudaError_t e2 = cudaFuncSetAttribute(bonded_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E->bonded_smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { mlp_set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); rc = MLP_E_CUDA; }
    }
    cudaSetDevice(prev);
    if (rc != MLP_OK) { mlp_energy_destroy(E); return rc; }
    *out = E;
    return MLP_OK;
}

void mlp_energy_destroy(mlp_energy* e) {
    if (!e) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    for (void* p : e->allocs) cudaFree(p);
    cudaSetDevice(prev);
    delete e;
}

int64_t mlp_energy_workspace_bytes(const mlp_energy* e, int64_t B) {
    if (!e || B < 0) return MLP_E_ARG;
    return 256 + B * 4 * (int64_t)sizeof(double) + 256 + B * (int64_t)e->Npad * (int64_t)sizeof(float4);
}

int mlp_energy_last_launches(const mlp_energy* e) { return e ? e->last_launches : 0; }

int mlp_energy_profile(mlp_energy* e, int enable) {
    if (!e) { mlp_set_error("mlp_energy_profile: null handle"); return MLP_E_ARG; }
    e->profile = enable != 0;
    return MLP_OK;
}

int mlp_energy_profile_read(mlp_energy* e, double ms[4], int64_t counts[4]) {
    if (!e || !ms || !counts) { mlp_set_error("mlp_energy_profile_read: null argument"); return MLP_E_ARG; }
    for (int k = 0; k < 4; ++k) {
        for (auto& p : e->pending[k]) {
            CK(cudaEventSynchronize(p[1]));
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, p[0], p[1]));
            e->prof_ms[k] += t; e->prof_n[k] += 1;
            cudaEventDestroy(p[0]); cudaEventDestroy(p[1]);
        }
        e->pending[k].clear();
        ms[k] = e->prof_ms[k]; counts[k] = e->prof_n[k];
        e->prof_ms[k] = 0; e->prof_n[k] = 0;
    }
    return MLP_OK;
}

int mlp_energy_eval(mlp_energy* e, const float* x, int64_t B, int64_t sxb, int64_t sxc, int64_t sxn,
                    float* energies, float* grad, int64_t sgb, int64_t sgc, int64_t sgn,
                    const float* weights, uint32_t term_mask, uint32_t flags,
                    void* workspace, int64_t workspace_bytes, void* stream_) {
    if (!e || B < 0 || (term_mask & ~MLP_TERM_ALL)) { mlp_set_error("mlp_energy_eval: bad argument"); return MLP_E_ARG; }
    if (B == 0 || e->N == 0) { e->last_launches = 0; return MLP_OK; }  // empty batch: nothing to do (pointers may be null)
    if (!x || !workspace) { mlp_set_error("mlp_energy_eval: null x / workspace"); return MLP_E_ARG; }
    if (workspace_bytes < mlp_energy_workspace_bytes(e, B)) { mlp_set_error("mlp_energy_eval: workspace too small"); return MLP_E_WORKSPACE; }
    if (reinterpret_cast<uintptr_t>(workspace) & 15) { mlp_set_error("mlp_energy_eval: workspace must be 16-byte aligned"); return MLP_E_ARG; }
    e->last_launches = 0;
    if (B == 0 || e->N == 0) return MLP_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int prev = 0;
    CK(cudaGetDevice(&prev));
    if (prev != e->device) CK(cudaSetDevice(e->device));
    char* ws = static_cast<char*>(workspace);
    double* acc = reinterpret_cast<double*>(ws);
    size_t off = ((size_t)B * 4 * sizeof(double) + 255) & ~(size_t)255;
    float4* xq = reinterpret_cast<float4*>(ws + off);
    const int N = e->N;
    const bool accumulate = flags & MLP_EVAL_ACCUMULATE_GRAD;
    int launches = 0;
    if (energies) { CK(cudaMemsetAsync(acc, 0, (size_t)B * 4 * sizeof(double), stream)); ++launches; }
    double* acc_arg = energies ? acc : nullptr;

    auto prof_begin = [&](int k) {
        if (!e->profile) return;
        std::array<cudaEvent_t, 2> p;
        cudaEventCreate(&p[0]); cudaEventCreate(&p[1]);
        cudaEventRecord(p[0], stream);
        e->pending[k].push_back(p);
    };
    auto prof_end = [&](int k) { if (e->profile) cudaEventRecord(e->pending[k].back()[1], stream); };

    if (term_mask & MLP_TERM_BONDED) {
        prof_begin(1);
        int cpc = (int)std::max<int64_t>(1, std::min<int64_t>(16, (B * e->n_tiles) / (int64_t)(e->sm_count * 8)));
        dim3 grid(e->n_tiles, (unsigned)((B + cpc - 1) / cpc));
        if (grad)
            bonded_kernel<true><<<grid, BT, e->bonded_smem, stream>>>(e->d_tiles, e->d_halo, e->d_bonds, e->d_angles, e->d_tors, e->d_csr,
                x, sxb, sxc, sxn, grad, sgb, sgc, sgn, weights, acc_arg, (int)B, cpc, term_mask & MLP_TERM_BONDED, accumulate, e->max_atoms);
        else
            bonded_kernel<false><<<grid, BT, e->bonded_smem, stream>>>(e->d_tiles, e->d_halo, e->d_bonds, e->d_angles, e->d_tors, e->d_csr,
                x, sxb, sxc, sxn, nullptr, 0, 0, 0, weights, acc_arg, (int)B, cpc, term_mask & MLP_TERM_BONDED, 0, e->max_atoms);
        prof_end(1);
        ++launches;
    } else if (grad && !accumulate) {
        long n = (long)B * N * 3;
        zero_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(grad, sgb, sgc, sgn, N, (int)B);
        ++launches;
    }
    if (term_mask & MLP_TERM_NB) {
        long n = (long)B * e->Npad;
        prof_begin(0);
        nb_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, sxb, sxc, sxn, e->d_q, xq, N, e->Npad, (int)B);
        prof_end(0);
        ++launches;
        prof_begin(2);
        long items = (long)B * e->n_pairs;
        unsigned blocks = (unsigned)((items + NB_WARPS - 1) / NB_WARPS);
        const bool full = e->nb_mode == MLP_NB_FULL;
#define NB_LAUNCH(G, F) nb_kernel<G, F><<<blocks, NB_WARPS * 32, 0, stream>>>(xq, e->d_par, e->d_pairs, e->d_masks, e->n_pairs, e->n_masked, e->Npad, N, (int)B, grad, sgb, sgc, sgn, weights, acc_arg)
        if (grad) { if (full) NB_LAUNCH(true, true); else NB_LAUNCH(true, false); }
        else { if (full) NB_LAUNCH(false, true); else NB_LAUNCH(false, false); }
#undef NB_LAUNCH
        prof_end(2);
        ++launches;
    }
    if (energies) {
        int n = (int)(B * 4);
        finalize_kernel<<<(n + 255) / 256, 256, 0, stream>>>(acc, energies, n);
        ++launches;
    }
    e->last_launches = launches;
    cudaError_t err = cudaGetLastError();
    if (prev != e->device) cudaSetDevice(prev);
    if (err != cudaSuccess) { mlp_set_error(std::string("kernel launch: ") + cudaGetErrorString(err)); return MLP_E_CUDA; }
    return MLP_OK;
}

int mlp_rmsd_eval(const float* x, int64_t B, int64_t N, int64_t sxb, int64_t sxc, int64_t sxn, float scale, float shift,
                  const float* targets, int64_t K, float* out, void* stream_) {
    if (!x || !targets || !out || B < 0 || N <= 0 || K <= 0 || K > 65535) { mlp_set_error("mlp_rmsd_eval: bad argument"); return MLP_E_ARG; }
    if (B == 0) return MLP_OK;
    rmsd_kernel<<<dim3((unsigned)B, (unsigned)K), 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, sxb, sxc, sxn, scale, shift, targets, (int)N, (int)K, out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { mlp_set_error(std::string("rmsd kernel launch: ") + cudaGetErrorString(err)); return MLP_E_CUDA; }
    return MLP_OK;
}

}  // extern "C"
