// B200 (sm_100a) kernels and C ABI for the molearn physics hot path: bonded bond/angle/torsion
// energy + analytic gradient, tiled O(N^2) soft non-bonded energy + force, RMSD reducer.
//
// Replaces the ATen op chains of TorchProteinEnergy._roll_bond_angle_torsion_loss
// (src/molearn/loss_functions/torch_protein_energy.py:276-305, ~500 launches) and _cdist_nb /
// _cdist_nb_full (:224-239, ~15 dense [B,N,N] passes) by one launch each.  FP32 on the FMA
// pipes; no tensor cores (not a dense contraction); O(N) memory; NO ATOMICS anywhere - every
// output is produced by a fixed-order sum, so results are bit-reproducible run to run.
//
// Bonded kernel
//   CTA = (atom tile, chunk of conformations).  The host matches every angle to a torsion that contains it and
//   every bond to an item that contains it, so a WORK ITEM is up to one torsion + one angle + one bond on the same
//   four atoms: four coordinate reads, shared difference vectors / cross product, one summed gradient per atom.
//   The tile's item records live in REGISTERS for the whole chunk (topology is read once per CTA, not once per
//   conformation).  Per conformation: coordinates of the tile (+halo) are staged in shared memory as float4, each
//   thread evaluates its items (branch-free) and writes every atom's gradient to a private shared-memory slot;
//   slot (atom a, k-th contribution) lives at [k*stride + a]; one thread per atom sums its column and writes the
//   gradient once.  Items that straddle a tile boundary are evaluated by both tiles (energy counted by the tile
//   that owns the term's first atom).
//
// Non-bonded kernel
//   Work item = (conformation, 128x128 atom tile pair J>=I), one one-warp CTA each, tile pairs ordered
//   longest first (the diagonal tiles last, cut into two step ranges).  A lane owns NI=4 i-atoms in registers for
//   the whole tile; the 32 groups of 4 j-atoms rotate through the lanes together with their force accumulators
//   (Newton's third law for 12 shuffles per 16 pairs; diagonal tiles need only 17 of the 32 rotation steps).  Tile
//   pairs that contain excluded (1-2/1-3/1-4) pairs or the diagonal carry a 128x128 bit mask, applied in registers
//   in the rotation steps that touch an excluded pair.  The hot loop is the fast path w(d,k) = d at a clamped
//   distance, written in packed FP32 (FFMA2/FMUL2 over two j-atoms per instruction: the scalar version was
//   issue-bound, the packed one is FMA-pipe-bound); the rare pairs closer than 1.9 A are corrected exactly.  Items
//   whose first step is dominated by close pairs (collapsed structures) run their other steps in a second loop that
//   evaluates every pair exactly (packed as well) and skips the fast path.  Each item writes its partial forces to a
//   scratch buffer; a second kernel sums, per atom, the partials of the tile pairs in its row and column in a fixed
//   order, adds them to the gradient and reduces the conformation's energy partials.  nb_kernel (scalar) is kept
//   for comparison (MLP_NB_PACKED=0).
#include "../../include/molearn_b200.h"
#include "topology.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

// ------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------
#ifndef BONDED_THREADS
#define BONDED_THREADS 128
#endif
constexpr int BT = BONDED_THREADS;  // bonded: threads per CTA (a multiple of 32, >= BTA)
constexpr int BTA = 128;         // bonded: most atoms per tile (one gather thread per atom)
static_assert(BT % 32 == 0 && BT >= BTA, "one gather thread per tile atom");
constexpr int RI = 2;            // bonded: max work items per thread
constexpr int SLOT_STRIDE = BTA + 1;  // odd stride (in float4s): an atom's consecutive slots fall into different banks
constexpr int PF = 4;            // bonded: coordinate elements staged per thread (3*(own+halo) <= PF*BT)
constexpr int NF = 4;            // torsion Fourier orders kept (ff14SB periods are 1..4)
constexpr int NB_T = 128;        // non-bonded tile edge (atoms)
constexpr int NB_NI = 4;         // i-atoms per lane
#ifndef NB_WARPS_PER_CTA
#define NB_WARPS_PER_CTA 1
#endif
constexpr int NB_WARPS = NB_WARPS_PER_CTA;  // warps per CTA, one work item each (1: a finished item frees its slot at once)
#ifndef BONDED_CPC_MAX
#define BONDED_CPC_MAX 32   // conformations per bonded CTA (the item records are read once per CTA)
#endif
#ifndef BONDED_DB
#define BONDED_DB 0   // 1: bonded kernel double-buffered over conformations, one barrier per conformation instead of two (twice the shared memory; measured SLOWER: 0.270 vs 0.258 ms at B = 4096)
#endif
#ifndef BONDED_MINB
#define BONDED_MINB (640 / BONDED_THREADS)   // bonded CTAs per SM the register budget is sized for (20 warps x 96 registers)
#endif
#ifndef NB2_MINB
#define NB2_MINB 3   // packed kernel: resident warps per SM sub-partition the register budget is sized for
#endif
#ifndef NB_BLK_MEM_MB
#define NB_BLK_MEM_MB 64   // most partial forces of one conformation in one pair launch; larger systems run band by band through their tile-pair rows
#endif
#ifndef NB2_RCP_LUT
#define NB2_RCP_LUT 0   // table form: 1/r^2 = (1/r)^2 on the FMA pipe (one MUFU per pair; with four resident warps the XU pipe was 50 % busy and MUFU issue the top stall: -1.3 %)
#endif
#ifndef NB2_MINB_LUT
#define NB2_MINB_LUT 4   // the table form needs fewer registers: four resident warps per sub-partition (128 registers) - equal at MurD x 64, +2 % at 10 k atoms
#endif
#ifndef NB2_RCP
#define NB2_RCP 1    // 1/r^2 of the force scalar from MUFU.RCP instead of inv*inv: one FMA-pipe operation per pair moves to the XU pipe (+1.1 %)
#endif
#ifndef NB_PACKED
#define NB_PACKED 1  // 1: nb_kernel2 (packed FFMA2 pair loop), 0: scalar nb_kernel
#endif
#ifndef NB_MINB
#define NB_MINB 4    // resident warps per SM sub-partition the register budget is sized for (4 -> 128 registers)
#endif
constexpr float LJ_K = 1.9f, C_K = 0.4f;  // warp thresholds, torch_protein_energy.py:233-234
constexpr float CLAMP = 0.999999f;        // acos clamp, :289

// one bonded work item: torsion (i,j,k,l) + angle (i,j,k) + bond (i,j)|(j,k), any subset (flags)
struct ItemRec {
    uint32_t ij, kl;          // local atom indices, 16 bits each
    uint32_t s_ij, s_kl;      // gradient slot ids of the four atoms, 16 bits each
    float c0, cg[NF];         // torsion Fourier series (cosine coefficients)
    float th0, ka, r0, kb;    // angle (theta0, K), bond (r0, K)
    uint32_t flags, pad[2];
};
static_assert(sizeof(ItemRec) == 64, "record layout");
constexpr uint32_t IT_T = 1u, IT_A = 2u, IT_B = 4u, IT_B_JK = 8u;       // which terms the item carries; bond on (j,k)
constexpr uint32_t IT_OWN_T = 16u, IT_OWN_A = 32u, IT_OWN_B = 64u;      // this tile counts the term's energy

struct BTile {
    int a0, n_own, n_halo, halo_off;
    int item_off, n_item;            // padded to a multiple of 32
    int cnt_off, pad;
};

struct NbBlockItem;
constexpr int NB_BLK_LEVELS = 3;
constexpr int NB_BLK_ROWS[NB_BLK_LEVELS] = {2, 4, 8};

struct mlp_energy {
    int device = 0;
    uint32_t nb_mode = 0;
    int N = 0, Npad = 0;
    // bonded
    int n_tiles = 0, tile_atoms = 0, max_cnt = 0, max_atoms = 0;
    long n_items = 0;              // bonded work items over all tiles (boundary items are evaluated by each tile they touch)
    bool torsion_sines = false;    // some torsion phase is not 0 / pi: the kernel keeps the sine coefficients
    BTile* d_tiles = nullptr;
    int* d_halo = nullptr;
    ItemRec* d_items = nullptr;
    float4* d_sg = nullptr;        // [items] torsion sine coefficients (only read when torsion_sines)
    int* d_cnt = nullptr;
    size_t bonded_smem = 0;
    // non-bonded
    int n_nbt = 0;                 // tiles per edge
    int n_pairs = 0, n_masked = 0; // tile pairs (masked ones first)
    int4* d_pairs = nullptr;       // (I, J, mask index or -1, rotation steps), longest items first
    uint32_t* d_masks = nullptr;   // [n_masked][32 groups][32 lanes] uint16 pair masks
    uint32_t* d_stepbits = nullptr; // [n_masked] bit r: rotation step r touches an excluded pair
    float2* d_parw = nullptr;      // [Npad] (a, b) product-form LJ parameters of the packed repulsive kernel
    float* d_q = nullptr;          // [Npad]
    float2* d_par = nullptr;       // [Npad] (R/2, sqrt(12 eps))
    // type-pair table of the repulsive mode (systems with few Lennard-Jones classes): w_ij^2 looked up, not computed
    bool lut_on = false;           // MLP_NB_LUT=0 switches it off
    int lut_classes = 0, lut_n = 0; // classes incl. the zero class of the padding atoms; table entries (classes^3)
    uint32_t* d_lrow = nullptr;    // [Npad] byte offset of the atom's table row
    uint2* d_lpc = nullptr;        // [Npad/4] byte offsets, within a row, of the group's two j-atom pairs
    float2* d_lut = nullptr;       // [c][c][c] (w^2(ci, cj0), w^2(ci, cj1))
    // row bands of the tile-pair list: when one conformation's partial forces exceed NB_BLK_MEM_MB the pair term runs
    // band by band (rows [I_k, I_k+1) of the tile-pair triangle: pair launch, then a reduction that adds to the gradient)
    struct PairBand { int begin = 0, count = 0; int* d_red_off = nullptr; int* d_red_list = nullptr; };
    std::vector<PairBand> bands;
    int4* d_pairs_b = nullptr;     // the pair list ordered band by band (longest first inside a band)
    int band_max = 0;              // most pairs in a band
    size_t band_bytes = 0;         // partial forces of one conformation per band (NB_BLK_MEM_MB; MLP_NB_BAND_KB)
    int band_mode = -1;            // MLP_NB_BANDS: -1 auto, 0 never (block items instead)
    int* d_red_off = nullptr;      // [n_nbt+1] per tile: range in d_red_list
    int* d_red_list = nullptr;     // entries pair*2 + side (0: the tile is the pair's I, 1: its J)
    // block-item schedules of the pair term (nb_block_kernel): used when there is much more work than warp slots;
    // level k walks blocks of NB_BLK_ROWS[k] rows x NB_BLK_C columns (bigger blocks need more work per slot)
    struct BlockSchedule {
        int rows = 0, items = 0, blocks = 0;   // items and partial-force blocks per conformation
        NbBlockItem* d_items = nullptr;
        int* d_red_off = nullptr;
        int* d_red_list = nullptr;
    } blk[NB_BLK_LEVELS];
    int blk_acc_smem = 0;                // dynamic shared memory of a block-item CTA (j-force accumulators)
    int blk_mode = -1;                   // -1: by scratch per conformation, 0: never, R: always the level with R rows (MLP_NB_BLOCKS)
    int ws_cap_mb = 0;                   // partial-force scratch of one pair launch (MLP_NB_WS_MB; batches are evaluated in chunks)
    int2* d_pair_meta = nullptr;         // [nt(nt+1)/2] (mask index | -1, step bits) of tile pair (I, J >= I)
    float* d_tpar = nullptr;             // [nt][3][128] charge, R/2, sqrt(12 eps)
    float* d_tparw = nullptr;            // [nt][3][128] charge, a, b (product form of the repulsive mode)
    float* d_tparl = nullptr;            // [nt][320] charge, table row offsets, pair offsets (table form of the repulsive mode)
    int last_launches = 0;
    int sm_count = 148;
    // side stream: the (latency-bound) bonded kernel runs next to the non-bonded pair kernel
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap = true;
    bool nb_first = false;          // launch order of the two concurrent kernels (MLP_NB_FIRST)
    bool packed = NB_PACKED != 0;   // non-bonded pair loop in packed FP32 (FFMA2)
    // optional per-kernel timing (mlp_energy_profile): events on the launching stream
    bool profile = false;
    std::vector<std::array<cudaEvent_t, 2>> pending[4];  // per kernel: (start, stop) pairs not read yet
    std::vector<char> pending_first[4];                   // 0: a further chunk of the same call (its time is added, the call is counted once)
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
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 1/sqrt(x) to ~1 ulp: MUFU + one Newton step (the bare MUFU is 2 ulp; (d - r0) and cos(n phi) magnify it)
__device__ __forceinline__ float rsqrt_nr(float x) {
    float y = rsqrt_fast(x);
    float h = 0.5f * x * y;
    return y * fmaf(-h, y, 1.5f);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// bonded kernel
// ------------------------------------------------------------------------------------------
// One work item = up to one torsion (i,j,k,l), one angle (i,j,k) and one bond ((i,j) or (j,k)) on the SAME
// four atoms (pack_bonded matches every angle to a torsion that contains it and every bond to an item that
// contains it; torsions are reversed where needed - phi(i,j,k,l) = phi(l,k,j,i)).  The three terms share the
// four coordinate reads, F = ri-rj, G = rj-rk, A = F x G and the squared lengths, and their gradients are summed
// in registers: 4 shared-memory reads and 4 slot writes per item instead of 9 + 9 for three separate terms.
// (The separate-term kernel was bound by the shared-memory pipe: 78 % LSU wavefront utilisation.)
//   bond   E = K (|v| - r0)^2                                      (torch_protein_energy.py:282-284)
//   angle  theta = acos(clamp(v1.v2/(|v1||v2|))), E = K (theta-theta0)^2                  (:286-293)
//   torsion phi = atan2(b2.(c32 x c12), |b2| c12.c32), E = sum_p (V/f)(1+cos(n phi - gamma)) (:295-304),
//          evaluated as a Fourier series in (cos phi, sin phi) - no atan2 / cos calls.
// SG: the series has sine coefficients (phases other than 0 / pi).  ff14SB has none, so the usual
// instantiation drops them.
// Branch-free form: every item evaluates all three terms in ONE basic block, so the scheduler can interleave the
// three dependency chains (the kernel is latency-bound).  Terms an item does not carry have zero force constants /
// Fourier coefficients; the few reciprocals that could turn 0 * inf into NaN on an item's dummy atoms are forced
// to zero instead (hT / hA selects), and energies are added under the OWN flags, which imply presence.
template <bool GRAD, bool SG>
__device__ __forceinline__ void item_term(const float4* sc, float4* slots, const ItemRec& r, const float4& sg, uint32_t f,
                                          float wb, float wa, float wt, float& eb, float& ea, float& et) {
    const uint32_t i = r.ij & 0xffffu, j = r.ij >> 16, k = r.kl & 0xffffu, l = r.kl >> 16;
    const bool hT = f & IT_T, hA = f & IT_A;
    const float3 rj = f3(sc[j]), rk = f3(sc[k]);
    const float3 F = f3(sc[i]) - rj, G = rj - rk, H = f3(sc[l]) - rk;
    const float F2 = dot(F, F), G2 = dot(G, G), FG = dot(F, G);
    const float3 A = cross(F, G), B = cross(H, G);
    const float A2 = dot(A, A);
    float3 gi, gj, gk, gl;
    // ---- torsion
    {
        float iG = G2 > 0.f ? rsqrt_nr(G2) : (G2 == 0.f ? 0.f : G2);   // coincident j,k: |G| = 0, phi = atan2(0,0) = 0 like the reference
        float Gn = G2 * iG;
        float X = dot(A, B), Y = -Gn * dot(B, F);
        float rho2 = fmaf(X, X, Y * Y);
        float irho = rsqrt_nr(rho2);
        // atan2(0,0) = 0 in the reference's forward; NaN/inf fall through and propagate
        float c1 = rho2 == 0.f ? 1.f : X * irho, s1 = rho2 == 0.f ? 0.f : Y * irho;
        float cn = c1, sn = s1;
        float en = r.c0, de = 0.f;  // de = dE/dphi
        const float sgv[NF] = {sg.x, sg.y, sg.z, sg.w};
#pragma unroll
        for (int n = 1; n <= NF; ++n) {
            if (SG) {
                en = fmaf(r.cg[n - 1], cn, fmaf(sgv[n - 1], sn, en));
                if (GRAD) de = fmaf((float)n, fmaf(sgv[n - 1], cn, -r.cg[n - 1] * sn), de);
            } else {
                en = fmaf(r.cg[n - 1], cn, en);
                if (GRAD) de = fmaf(-(float)n * r.cg[n - 1], sn, de);
            }
            if (n < NF) {
                float c2 = fmaf(cn, c1, -sn * s1);
                sn = fmaf(sn, c1, cn * s1);
                cn = c2;
            }
        }
        et += (f & IT_OWN_T) ? en : 0.f;
        if (GRAD) {
            de = hT ? de * wt : 0.f;
            float iA2 = hT ? rcp_fast(A2) : 0.f, iB2 = hT ? rcp_fast(dot(B, B)) : 0.f;
            float3 ti = (-de * Gn * iA2) * A;
            float3 tl = (de * Gn * iB2) * B;
            float fa = de * FG * iA2 * iG, hb = de * dot(H, G) * iB2 * iG;
            float3 sv = fa * A - hb * B;
            gi = ti; gl = tl; gj = sv - ti; gk = -1.f * (tl + sv);
        }
    }
    // ---- angle (i,j,k): v1 = ri - rj = F, v2 = rk - rj = -G, n = v1 x v2 = -A
    {
        float c = -FG * rsqrt_nr(F2 * G2);
        float cc = fminf(fmaxf(c, -CLAMP), CLAMP);
        float th = acosf(cc);
        if (c != c) th = c;  // NaN propagates (fminf/fmaxf would swallow it)
        float dt = th - r.th0;
        float kdt = r.ka * dt;
        ea += (f & IT_OWN_A) ? kdt * dt : 0.f;
        if (GRAD) {
            // dtheta/dv1 = v1 x (v1 x v2) / (|v1|^2 |v1 x v2|), dtheta/dv2 = -v2 x (v1 x v2) / (|v2|^2 |v1 x v2|):
            // the cross-product form of -1/sqrt(1-c^2) * dc/dv, which does not cancel near c = +-1.
            // torch.clamp passes no gradient outside the clamp, so the term is dropped there.
            float coef = (c > -CLAMP && c < CLAMP) ? 2.f * wa * kdt * rsqrt_fast(A2) : 0.f;
            if (c != c) coef = c;
            coef = hA ? coef : 0.f;
            float qi = hA ? -coef * rcp_fast(F2) : 0.f, qk = hA ? -coef * rcp_fast(G2) : 0.f;
            float3 ai = qi * cross(F, A);
            float3 ak = qk * cross(G, A);
            gi = gi + ai; gk = gk + ak; gj = gj - (ai + ak);
        }
    }
    // ---- bond (i,j) or (j,k)
    {
        const bool jk = f & IT_B_JK;
        const float3 v = jk ? G : F;
        const float d2 = jk ? G2 : F2;
        float inv = d2 > 0.f ? rsqrt_nr(d2) : 0.f;   // autograd: d|v|/dv = 0 at v = 0
        float d = d2 * inv;
        d = d2 != d2 ? d2 : d;
        float dl = d - r.r0;
        float kdl = r.kb * dl;
        eb += (f & IT_OWN_B) ? kdl * dl : 0.f;
        if (GRAD) {
            float sc2 = (f & IT_B) ? 2.f * wb * kdl * inv : 0.f;
            float3 g = sc2 * v;
            float mi = jk ? 0.f : 1.f, mk = jk ? -1.f : 0.f;     // (i,j): +g on i, -g on j; (j,k): +g on j, -g on k
            gi = gi + mi * g; gk = gk + mk * g; gj = gj - (mi + mk) * g;
        }
    }
    if (GRAD) {
        slots[r.s_ij & 0xffffu] = f4(gi);
        slots[r.s_ij >> 16] = f4(gj);
        slots[r.s_kl & 0xffffu] = f4(gk);
        slots[r.s_kl >> 16] = f4(gl);
    }
}

template <bool GRAD, bool SG>
__global__ void __launch_bounds__(BT, BONDED_MINB)
bonded_kernel(const BTile* __restrict__ tiles, const int* __restrict__ halo, const ItemRec* __restrict__ items,
              const float4* __restrict__ sgs, const int* __restrict__ cnts,
              const float* __restrict__ x, long sxb, int sxc, int sxn,
              float* __restrict__ grad, long sgb, int sgc, int sgn,
              const float* __restrict__ weights, float* __restrict__ e_part /* [B][n_tiles][3] or null */,
              int B, int conf_per_cta, uint32_t term_mask, int accumulate, int max_atoms, int max_cnt) {
    constexpr int STRIDE = SLOT_STRIDE;  // slot (atom a, k-th contribution) at k*STRIDE + a
    extern __shared__ float4 smem[];
    // Double-buffered over conformations (BONDED_DB): conformation b's items write slots[b & 1] while the slower
    // threads still gather conformation b-1 from the other half, and the coordinates of b+1 are staged into the other
    // coordinate buffer - ONE barrier per conformation instead of two.
    constexpr int NBUF = BONDED_DB ? 2 : 1;
    const int slot_span = max_cnt * STRIDE + 1;
    float4* sc0 = smem;                          // [NBUF][max_atoms] coordinates of own + halo atoms
    float4* slots0 = smem + NBUF * max_atoms;    // [NBUF][max_cnt*STRIDE + 1] gradient slots; last = dump
    __shared__ float s_e[NBUF][3][BT];

    const BTile t = tiles[blockIdx.x];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // item records -> registers, once per CTA (the list is padded to whole warps with inert items).
    // 32-item unit u goes to warp u % NW, register slot u / NW.
    ItemRec ri[RI];
    float4 rs[RI];
    bool ai[RI];                   // warp-uniform activity flags
    constexpr int NW = BT / 32;
    // terms switched off by term_mask lose their flag bits (the item still writes its four slots)
    const uint32_t fmask = ((term_mask & MLP_TERM_TORSION) ? (IT_T | IT_OWN_T) : 0u) | ((term_mask & MLP_TERM_ANGLE) ? (IT_A | IT_OWN_A) : 0u) |
                           ((term_mask & MLP_TERM_BOND) ? (IT_B | IT_B_JK | IT_OWN_B) : 0u);
#pragma unroll
    for (int r = 0; r < RI; ++r) {
        int i0 = (r * NW + warp) * 32;
        ai[r] = i0 < t.n_item;
        rs[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ai[r]) {
            ri[r] = items[t.item_off + i0 + lane];
            ri[r].flags &= fmask;
            if (SG) rs[r] = sgs[t.item_off + i0 + lane];
        }
    }
    const int my_cnt = (GRAD && tid < t.n_own) ? cnts[t.cnt_off + tid] : 0;
    const int wcnt = (__reduce_max_sync(0xffffffffu, my_cnt) + 3) & ~3;   // <= max_cnt (a multiple of 4)
    if (GRAD) {
        for (int s = tid; s < NBUF * slot_span; s += BT) slots0[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        // ordered before the first slot write by the first __syncthreads of the conformation loop
    }

    // coordinate staging plan: element e = tid + q*BT of the tile's 3*(own+halo) floats
    // element offsets inside one conformation are 32-bit (the host checks the extent): a 64-bit multiply-add
    // per staged element and per stored gradient component was 10 % of the kernel's instructions
    int goff[PF];
    int sidx[PF];
    {
        const int n_el = 3 * (t.n_own + t.n_halo);
        const bool aos = (sxc == 1);
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            int e = tid + q * BT;
            sidx[q] = -1;
            goff[q] = 0;
            if (e < n_el) {
                int la, c;
                if (e < 3 * t.n_own) {
                    if (aos) { la = e / 3; c = e - 3 * la; } else { c = e / t.n_own; la = e - c * t.n_own; }
                } else {
                    int h = e - 3 * t.n_own;
                    la = t.n_own + h / 3; c = h - 3 * (h / 3);
                }
                int ga = la < t.n_own ? t.a0 + la : halo[t.halo_off + la - t.n_own];
                goff[q] = c * sxc + ga * sxn;
                sidx[q] = 4 * la + c;
            }
        }
    }

    const int b0 = blockIdx.y * conf_per_cta;
    const int b1 = min(B, b0 + conf_per_cta);
    float pf[PF];
    auto prefetch = [&](const float* xb) {
#pragma unroll
        for (int q = 0; q < PF; ++q) if (sidx[q] >= 0) pf[q] = xb[goff[q]];
    };
    const float* xb = x + (long)b0 * sxb;                 // conformation b (running pointers: no per-element 64-bit multiply)
    float* gb = GRAD ? grad + (long)b0 * sgb : nullptr;
    const int gofs = (t.a0 + tid) * sgn;
    const bool do_b = term_mask & MLP_TERM_BOND, do_a = term_mask & MLP_TERM_ANGLE, do_t = term_mask & MLP_TERM_TORSION;

    auto stage = [&](float4* sc) {        // prefetched registers -> coordinate buffer
        float* scf = reinterpret_cast<float*>(sc);
#pragma unroll
        for (int q = 0; q < PF; ++q) if (sidx[q] >= 0) scf[sidx[q]] = pf[q];
    };
    auto energy_and_gather = [&](int b, const float4* slots, const float (*se)[BT], float* gb) {
        if (e_part && warp < 3) {
            // warp w sums term w over the CTA in a fixed order (reproducible); the other warps go on
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < BT / 32; ++k) s += se[warp][lane + 32 * k];
            s = warp_sum(s);
            if (lane == 0) e_part[((long)b * gridDim.x + blockIdx.x) * 3 + warp] = s;
        }
        if (GRAD) {
            // per-atom column sum; the bound is warp-uniform and a multiple of 4: slots past an atom's own
            // count are never written by any term and hold the zeros stored at kernel start
            float gx = 0.f, gy = 0.f, gz = 0.f;
            const float4* col = slots + tid;
            for (int k = 0; k < wcnt; k += 4, col += 4 * STRIDE) {      // fixed stride: immediate-offset loads
                float4 v0 = col[0], v1 = col[STRIDE], v2 = col[2 * STRIDE], v3 = col[3 * STRIDE];
                gx += (v0.x + v1.x) + (v2.x + v3.x); gy += (v0.y + v1.y) + (v2.y + v3.y); gz += (v0.z + v1.z) + (v2.z + v3.z);
            }
            if (tid < t.n_own) {
                float* g = gb + gofs;
                if (accumulate) { g[0] += gx; g[sgc] += gy; g[2 * sgc] += gz; }
                else { g[0] = gx; g[sgc] = gy; g[2 * sgc] = gz; }
            }
        }
    };
    // every requested term of a conformation has zero upstream weight (CTA-uniform): nothing to add, or an exact zero to
    // store - its coordinates never enter any arithmetic, so a non-finite conformation cannot leak NaN * 0
    auto all_zero = [&](float wb, float wa, float wt) {
        return GRAD && weights && !e_part && (!do_b || wb == 0.f) && (!do_a || wa == 0.f) && (!do_t || wt == 0.f);
    };
#if BONDED_DB
    if (b0 < b1) {
        prefetch(xb);
        stage(sc0);
        if (b0 + 1 < b1) prefetch(xb + sxb);
    }
    __syncthreads();
    for (int b = b0; b < b1; ++b, xb += sxb, gb += sgb) {
        const int cur = (b - b0) & 1;
        float4* sc = sc0 + cur * max_atoms;
        float4* slots = slots0 + cur * slot_span;
        float wb = 1.f, wa = 1.f, wt = 1.f;
        if (weights) { wb = weights[4 * b]; wa = weights[4 * b + 1]; wt = weights[4 * b + 2]; }
        const bool skip = all_zero(wb, wa, wt);
        // a term with zero upstream weight is dropped like a term the item does not carry (its degenerate-geometry
        // NaN must not reach the gradient as NaN * 0); its energy is still counted
        const uint32_t wdrop = weights ? ((wt == 0.f ? IT_T : 0u) | (wa == 0.f ? IT_A : 0u) | (wb == 0.f ? IT_B : 0u)) : 0u;
        if (!skip) {
            float eb = 0.f, ea = 0.f, et = 0.f;
#pragma unroll
            for (int r = 0; r < RI; ++r)
                if (ai[r]) item_term<GRAD, SG>(sc, slots, ri[r], rs[r], ri[r].flags & ~wdrop, wb, wa, wt, eb, ea, et);
            if (e_part) { s_e[cur][0][tid] = eb; s_e[cur][1][tid] = ea; s_e[cur][2][tid] = et; }
        }
        // next conformation: registers -> the other coordinate buffer (its last readers, the items of b-1, are behind the
        // previous barrier), then the loads of the one after it
        if (b + 1 < b1) {
            stage(sc0 + (cur ^ 1) * max_atoms);
            if (b + 2 < b1) prefetch(xb + 2 * sxb);
        }
        __syncthreads();
        // slots[cur] / s_e[cur] are complete; they are written again two conformations later, behind the next barrier
        if (!skip) energy_and_gather(b, slots, s_e[cur], gb);
        else if (!accumulate && tid < t.n_own) { float* g = gb + gofs; g[0] = 0.f; g[sgc] = 0.f; g[2 * sgc] = 0.f; }
    }
#else
    float4* sc = sc0;
    float4* slots = slots0;
    if (b0 < b1) prefetch(xb);
    for (int b = b0; b < b1; ++b, xb += sxb, gb += sgb) {
        float wb = 1.f, wa = 1.f, wt = 1.f;
        if (weights) { wb = weights[4 * b]; wa = weights[4 * b + 1]; wt = weights[4 * b + 2]; }
        if (all_zero(wb, wa, wt)) {
            if (!accumulate && tid < t.n_own) { float* g = gb + gofs; g[0] = 0.f; g[sgc] = 0.f; g[2 * sgc] = 0.f; }
            if (b + 1 < b1) prefetch(xb + sxb);
            continue;
        }
        const uint32_t wdrop = weights ? ((wt == 0.f ? IT_T : 0u) | (wa == 0.f ? IT_A : 0u) | (wb == 0.f ? IT_B : 0u)) : 0u;
        // stage coordinates (the previous iteration's term phase is over: everybody passed its second barrier)
        stage(sc);
        __syncthreads();
        if (b + 1 < b1) prefetch(xb + sxb);
        float eb = 0.f, ea = 0.f, et = 0.f;
#pragma unroll
        for (int r = 0; r < RI; ++r)
            if (ai[r]) item_term<GRAD, SG>(sc, slots, ri[r], rs[r], ri[r].flags & ~wdrop, wb, wa, wt, eb, ea, et);
        if (e_part) { s_e[0][0][tid] = eb; s_e[0][1][tid] = ea; s_e[0][2][tid] = et; }
        __syncthreads();
        energy_and_gather(b, slots, s_e[0], gb);
        // the next iteration's staging only touches sc (term phase done); its slot writes happen after the
        // next barrier, i.e. after every gather above.
    }
#endif
}

// ------------------------------------------------------------------------------------------
// non-bonded kernels
// ------------------------------------------------------------------------------------------
// Atom n of a conformation as (x, y, z, q), read from the caller's tensor with its own strides (the batch
// dimension is already applied to xb).  Atoms past N pad the last 128-atom tile: they are parked far away
// with zero charge / epsilon so that they contribute exactly 0 without any test in the pair loop.
__device__ __forceinline__ float4 nb_load_atom(const float* __restrict__ xb, long sxc, long sxn,
                                               const float* __restrict__ q, int n, int N) {
    if (n < N) {
        const float* p = xb + (long)n * sxn;
        return make_float4(p[0], p[sxc], p[2 * sxc], q[n]);
    }
    return make_float4(1.0e6f + 100.f * (float)(n - N), -1.0e6f, 1.0e6f, 0.f);
}

// Squared distance below which a pair leaves the fast path (w(d,1.9) = d needs d > 1.9 A; the
// margin covers the 2-ulp rsqrt).  Pairs closer than this are evaluated by the fast path at the
// clamped distance sqrt(NB_THR2) - a small, finite value - and then corrected by nb_pair_fix().
constexpr float NB_THR2 = LJ_K * LJ_K * 1.0001f;
constexpr float NB_MASKED_R2 = 1.0e4f;  // stand-in r^2 of statically excluded pairs (their parameters are zeroed)

// Fast-path pair term at squared distance r2c >= NB_THR2.  Returns s = -(dE/dd)/d; adds 12*E_lj to
// Sl and E_coulomb to Sc.
// W2 (type-pair table form of the repulsive mode, nb_micro2<.., LUT>): Rij holds w_ij^2 and e12 = 1.
template <bool FULL, bool W2 = false>
__device__ __forceinline__ float nb_pair_fast(float r2c, float Rij, float e12, float qq, float& Sl, float& Sc) {
    float inv = rsqrt_fast(r2c);
    float t = Rij * inv;
    float t2 = W2 ? Rij * rcp_fast(r2c) : t * t;
    float t6 = t2 * t2 * t2;
    float u6 = e12 * t6;
    float ec = qq * inv;
    float u;
    if (FULL) { u = fmaf(u6, t6 - 1.f, ec); Sl = fmaf(u6, t6 - 2.f, Sl); }
    else { u = fmaf(u6, t6, ec); Sl = fmaf(u6, t6, Sl); }
    Sc = fmaf(qq, inv, Sc);
    return W2 ? u * rcp_fast(r2c) : u * (inv * inv);
}

// Exact pair term (any distance), minus what the fast path added for this pair at the clamped
// distance.  Rare path (clashing structures): kept out of line to protect the hot loop's registers
// and instruction cache.  Returns (ds, dSl, dSc).
#ifndef NB_DENSE_MODE
#define NB_DENSE_MODE 1   // 1: items whose first step is dominated by close pairs evaluate their other steps exactly, without the fast path
#endif
#ifndef NB_FIX_INLINE
#define NB_FIX_INLINE 1   // 1: the exact pair correction is inlined into the micro-tile fix (1.2x in the clash regime, hot loop unchanged)
#endif
#if NB_FIX_INLINE
#define NB_FIX_ATTR __forceinline__
#else
#define NB_FIX_ATTR __noinline__
#endif
template <bool FULL, bool W2 = false>
__device__ NB_FIX_ATTR float3 nb_pair_fix(float r2, float Rij, float e12, float qq) {
    float Sl_f = 0.f, Sc_f = 0.f;
    float s_f = nb_pair_fast<FULL, W2>(NB_THR2, Rij, e12, qq, Sl_f, Sc_f);
    // the fast path multiplied s by the true (dx,dy,dz); so must the correction.
    // Lean arithmetic (a collapsed decode sends most pairs through here): rsqrt + Newton for d, ex2.approx for the
    // two elu branches (|d - k| < 2: 2 ulp), rcp.approx for 1/w (1 ulp, x12 in the LJ term = 7e-7).
    float dinv = r2 > 0.f ? rsqrt_nr(r2) : 0.f;   // coincident atoms: zero force (cdist's backward does the same); NaN stays NaN
    float d = r2 * dinv;
    float ex1 = __expf(d - LJ_K), ex2 = __expf(d - C_K);
    const bool far1 = d > LJ_K, far2 = d > C_K;
    float w1 = far1 ? d : ex1 - 1.f + LJ_K, dw1 = far1 ? 1.f : ex1;     // w(d,k) = elu(d-k)+k  (torch_protein_energy.py:238-239)
    float w2 = far2 ? d : ex2 - 1.f + C_K, dw2 = far2 ? 1.f : ex2;
    float iw1 = rcp_fast(w1), iw2 = rcp_fast(w2);
    float t = Rij * iw1;
    float t2 = W2 ? Rij * (iw1 * iw1) : t * t;
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
// m16: bit (4a+b) = pair allowed (MASKED tiles).  Returns true if any allowed pair is closer than
// sqrt(NB_THR2) or not comparable (NaN coordinates) - the caller then runs the exact correction.
#ifndef NB_RMIN3
#define NB_RMIN3 0   // 1: one three-input FMNMX3 per two pairs; 0: two FMNMX.NAN (the three-input form issues at a lower rate: -1 % kernel time)
#endif
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
#if NB_RMIN3
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // sm_100 three-input min (one FMNMX3); NaN wins
#else
    asm("{ .reg .f32 t; min.NaN.f32 t, %2, %3; min.NaN.f32 %0, %1, t; }" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // NaN wins
#endif
    return r;
}

template <bool MASKED, bool FULL>
__device__ __forceinline__ bool nb_micro(const float4 (&xi)[NB_NI], const float2 (&pi)[NB_NI], const float4 (&xj)[4],
                                          const float2 (&pj)[4], uint32_t m16,
                                          float3 (&Fi)[NB_NI], float3 (&Fj)[4], float& Sl, float& Sc) {
    // smallest r^2 over the allowed pairs, two pairs per instruction; a NaN distance makes it NaN, which fails
    // the comparison below and sends the micro-tile to the exact path (where NaN propagates)
    float rmin = 3.0e38f;
#pragma unroll
    for (int a = 0; a < NB_NI; ++a) {
        float r2s[4];
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
            r2s[b] = r2;
            float s = nb_pair_fast<FULL>(fmaxf(r2, NB_THR2), Rij, e12, qq, Sl, Sc);
            Fi[a].x = fmaf(s, dx, Fi[a].x); Fi[a].y = fmaf(s, dy, Fi[a].y); Fi[a].z = fmaf(s, dz, Fi[a].z);
            Fj[b].x = fmaf(s, dx, Fj[b].x); Fj[b].y = fmaf(s, dy, Fj[b].y); Fj[b].z = fmaf(s, dz, Fj[b].z);
        }
        rmin = fmin3(rmin, r2s[0], r2s[1]);
        rmin = fmin3(rmin, r2s[2], r2s[3]);
    }
    return !(rmin >= NB_THR2);
}

// Correction pass over a micro-tile that holds at least one pair closer than sqrt(NB_THR2).
// WFORM: the per-atom parameters are (a, b) of nb_micro2's product form, R_ij (12 eps_ij)^(1/12) = a_i b_j + b_i a_j.
// PFORM 2: the table form - w2m[a][b] holds w_ij^2 of the pair (pi / pj are not read).
template <bool MASKED, bool FULL, int PFORM = 0>
__device__ __forceinline__ void nb_micro_fix(const float4 (&xi)[NB_NI], const float2 (&pi)[NB_NI], const float4 (&xj)[4],
                                             const float2 (&pj)[4], uint32_t m16,
                                             float3 (&Fi)[NB_NI], float3 (&Fj)[4], float& Sl, float& Sc,
                                             const float (*w2m)[4] = nullptr) {
    constexpr bool WFORM = PFORM == 1;
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
                const float Rij = PFORM == 2 ? w2m[a][b] : WFORM ? fmaf(pi[a].y, pj[b].x, pi[a].x * pj[b].y) : pi[a].x + pj[b].x;
                const float e12 = PFORM ? 1.f : pi[a].y * pj[b].y;
                float3 c = nb_pair_fix<FULL, PFORM == 2>(r2, Rij, e12, yi[a].w * yj[b].w);
                Fi[a].x = fmaf(c.x, dx, Fi[a].x); Fi[a].y = fmaf(c.x, dy, Fi[a].y); Fi[a].z = fmaf(c.x, dz, Fi[a].z);
                Fj[b].x = fmaf(c.x, dx, Fj[b].x); Fj[b].y = fmaf(c.x, dy, Fj[b].y); Fj[b].z = fmaf(c.x, dz, Fj[b].z);
                Sl += c.y; Sc += c.z;
            }
        }
    }
}

// One warp = one (conformation, tile pair).  Lane L owns i-atoms i0+4L+a (a<4) for the whole
// tile; the 32 groups of 4 j-atoms rotate through the lanes (lane L works on group (L+r)%32 at
// step r) together with their force accumulators, so Newton's third law costs 12 shuffles per
// 16 pairs and no reduction.  Output: partial forces of the item's 128 i-atoms and 128 j-atoms
// ([6][128] floats: -Fi xyz, +Fj xyz) and its energy - no atomics.
template <bool GRAD, bool FULL>
__global__ void __launch_bounds__(NB_WARPS * 32, (4 * NB_MINB) / NB_WARPS)
nb_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, const float* __restrict__ q, int N,
          const float2* __restrict__ par, const int4* __restrict__ pairs,
          const uint32_t* __restrict__ masks, int n_pairs, int B,
          float* __restrict__ fpart, const float* __restrict__ weights, double* __restrict__ e_part) {
    // j tile, group-minor: atom 4g+b lives at [b*32+g] so that lanes (different g) never conflict
    __shared__ float4 s_xj[NB_WARPS][NB_T];
    __shared__ float2 s_pj[NB_WARPS][NB_T];
    __shared__ uint32_t s_m[NB_WARPS][NB_T * 4];   // uint16 [g][lane] pair masks, as words
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work order: tile pair major, conformation minor; the pair list is sorted longest first (band tiles,
    // clean tiles, then the half-length diagonal tiles), so the grid ends with the short items
    const long work = (long)blockIdx.x * NB_WARPS + warp;
    if (work >= (long)B * n_pairs) return;
    const int p = (int)(work / B), b = (int)(work - (long)p * B);
    if (weights && !e_part && weights[4 * b + 3] == 0.f) return;  // backward correction with nothing to add (nb_reduce skips it too)
    const int4 IJ = pairs[p];                                       // (I, J, mask index or -1, rotation steps)
    const long item = (long)b * n_pairs + p;                        // slot in the partial buffers
    const int i0 = IJ.x * NB_T, j0 = IJ.y * NB_T;
    const float* xb = x + (long)b * sxb;
    const bool masked = IJ.z >= 0;

    float4 xi[NB_NI];
    float2 pi[NB_NI];
    float3 Fi[NB_NI];
#pragma unroll
    for (int a = 0; a < NB_NI; ++a) {
        xi[a] = nb_load_atom(xb, sxc, sxn, q, i0 + 4 * lane + a, N);
        pi[a] = par[i0 + 4 * lane + a];
        Fi[a] = make_float3(0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < NB_T / 32; ++k) {
        int jj = lane + 32 * k;
        int slot = (jj & 3) * 32 + (jj >> 2);
        s_xj[warp][slot] = nb_load_atom(xb, sxc, sxn, q, j0 + jj, N);
        s_pj[warp][slot] = par[j0 + jj];
    }
    if (masked) {
        const uint4* src = reinterpret_cast<const uint4*>(masks + (long)IJ.z * (NB_T * 4));
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

    // Diagonal tile: block (L,g) mirrors block (g,L), so 17 of the 32 rotation steps visit every unordered
    // block pair once (step 0 = the in-block triangle, step 16 = lanes < 16 only; the masks encode both).
    const int r_begin = IJ.w & 0xff, n_steps = IJ.w >> 8;    // rotation steps [r_begin, n_steps) of this tile pair
#pragma unroll 1
    for (int r = r_begin; r < n_steps; ++r) {
        const int g = (lane + r) & 31;
        float4 xj[4];
        float2 pj[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { xj[q] = s_xj[warp][q * 32 + g]; pj[q] = s_pj[warp][q * 32 + g]; }
        if (masked) {
            const uint32_t m16 = sm16[g * 32 + lane];
            if (nb_micro<true, FULL>(xi, pi, xj, pj, m16, Fi, Fj, Sl, Sc))
                nb_micro_fix<true, FULL>(xi, pi, xj, pj, m16, Fi, Fj, Sl, Sc);
        } else {
            if (nb_micro<false, FULL>(xi, pi, xj, pj, 0xffffu, Fi, Fj, Sl, Sc))
                nb_micro_fix<false, FULL>(xi, pi, xj, pj, 0xffffu, Fi, Fj, Sl, Sc);
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
        // after n_steps rotations lane L holds the forces of group (L + n_steps) % 32 (atoms j0+4g+q);
        // grad_i = -sum s*(ri-rj), grad_j = +sum s*(ri-rj)   (s = -(dE/dd)/d)
        float* out = fpart + item * (6 * NB_T);
        const int gl = (lane + n_steps) & 31;
        reinterpret_cast<float4*>(out + 0 * NB_T)[lane] = make_float4(-Fi[0].x, -Fi[1].x, -Fi[2].x, -Fi[3].x);
        reinterpret_cast<float4*>(out + 1 * NB_T)[lane] = make_float4(-Fi[0].y, -Fi[1].y, -Fi[2].y, -Fi[3].y);
        reinterpret_cast<float4*>(out + 2 * NB_T)[lane] = make_float4(-Fi[0].z, -Fi[1].z, -Fi[2].z, -Fi[3].z);
        reinterpret_cast<float4*>(out + 3 * NB_T)[gl] = make_float4(Fj[0].x, Fj[1].x, Fj[2].x, Fj[3].x);
        reinterpret_cast<float4*>(out + 4 * NB_T)[gl] = make_float4(Fj[0].y, Fj[1].y, Fj[2].y, Fj[3].y);
        reinterpret_cast<float4*>(out + 5 * NB_T)[gl] = make_float4(Fj[0].z, Fj[1].z, Fj[2].z, Fj[3].z);
    }
    if (e_part) {
        // E = sum_lj12/12 + sum_coulomb
        double e = (double)warp_sum(Sl) * (1.0 / 12.0) + (double)warp_sum(Sc);
        if (lane == 0) e_part[item] = e;
    }
}

// ---- packed-FP32 variant -------------------------------------------------------------------------------
// Blackwell's FMA pipe executes fma/mul/add.f32x2 (SASS FFMA2/FMUL2/FADD2): two FP32 lanes per instruction at
// the same flop rate.  The scalar kernel above is issue-bound (79 % issue, 63 % FMA pipe), so here every FMA-pipe
// operation of the pair term is packed over two j-atoms: j data sit in shared memory component-wise (x[128],
// y[128], ...), a lane reads float4s = two (j, j+1) pairs per component; the lane's i-atoms enter as 32-bit
// broadcast operands (the compiler keeps one register per value); i-force accumulators are (even-j, odd-j) partial
// sums folded at the end of the item.  Same arithmetic per pair as nb_pair_fast.
// What the instructions cost on sm_100a (tools/micro/issue_bench.cu, operand_bench.cu): a packed
// operation holds the pipe for two cycles - three when it reads three distinct register pairs -, and every MUFU / FMNMX /
// IADD3 / SHFL between two of them costs roughly one more (LDS: none).  The loop is bound by 2 * packed + others.
// NB_WFORM (repulsive mode only): the Lennard-Jones pair parameters are folded into one length,
//   12 eps_ij (R_ij/d)^12 = (w_ij/d)^12,  w_ij = (12 eps_ij)^(1/12) R_ij = a_i b_j + b_i a_j,
//   b = (12 eps)^(1/24), a = b R/2   (eps_ij = sqrt(eps_i eps_j), R_ij = R_i/2 + R_j/2; utils.py:814-815),
// which is 2 FMA-pipe operations per pair instead of 4 (R_ij, eps_ij, R_ij/d, eps_ij t^6): 26 instead of 27 per pair.
#ifndef NB_DIAG_SPLIT
#define NB_DIAG_SPLIT 2
#endif
#ifndef NB_REUSE_ORDER
#define NB_REUSE_ORDER 1   // source order / operand choice that lets consecutive FFMA2s share operands (operand reuse cache)
#endif
#ifndef NB_PIPE_J
#define NB_PIPE_J 0    // nb_kernel2: the j-group of the next rotation step is loaded before the shuffles of the current one
#endif
#ifndef NB_STEP_UNROLL
#define NB_STEP_UNROLL 1   // rotation steps per trip of the pair loop
#endif
#ifndef NB_KO
#define NB_KO 0      // knock-out experiments (wrong results; tools/nb_variants.sh): 1 no clamp, 2 no r_min, 4 no rotation, 8 no energy sums, 16 1/r^2 = inv^2, 32 no MUFU, 64 table value not loaded
#endif
#ifndef NB_WFORM
#define NB_WFORM 1
#endif
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 dup(float a) { return make_float2(a, a); }

// The lane's i-atom values enter every packed operation as (v, v).  NB_ISCALAR = 0 keeps them duplicated in
// register pairs (one FFMA2 per use); NB_ISCALAR = 1 keeps one copy and spends two scalar FFMA/FMUL/FADD per use:
// the same FMA-pipe cycles, twice the issue slots, 28 registers fewer (a fourth resident warp per scheduler).
// NB_FISCALAR likewise folds the (even j, odd j) halves of the i-force accumulators on the fly (12 registers).
#ifndef NB_ISCALAR
#define NB_ISCALAR 0
#endif
#ifndef NB_FISCALAR
#define NB_FISCALAR 0
#endif
#if NB_ISCALAR
typedef float ival;
__device__ __forceinline__ ival imake(float v) { return v; }
__device__ __forceinline__ float ilo(ival v) { return v; }
__device__ __forceinline__ float2 isub(float2 xj, ival xi) { return make_float2(xi - xj.x, xi - xj.y); }
__device__ __forceinline__ float2 imul(ival a, float2 b) { return make_float2(a * b.x, a * b.y); }
__device__ __forceinline__ float2 ifma(ival a, float2 b, float2 c) { return make_float2(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y)); }
__device__ __forceinline__ float2 iadd(ival a, float2 b) { return make_float2(a + b.x, a + b.y); }
#else
typedef float2 ival;
__device__ __forceinline__ ival imake(float v) { return make_float2(v, v); }
__device__ __forceinline__ float ilo(ival v) { return v.x; }
__device__ __forceinline__ float2 isub(float2 xj, ival xi) { return __ffma2_rn(xj, make_float2(-1.f, -1.f), xi); }
__device__ __forceinline__ float2 imul(ival a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 ifma(ival a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 iadd(ival a, float2 b) { return __ffma2_rn(a, make_float2(1.f, 1.f), b); }   // add.f32x2 is lowered to two FADDs; FFMA2 by one stays packed
#endif
#if NB_FISCALAR
typedef float fval;
__device__ __forceinline__ fval fzero() { return 0.f; }
__device__ __forceinline__ void facc(fval& F, float2 s, float2 d) { F = fmaf(s.x, d.x, fmaf(s.y, d.y, F)); }
__device__ __forceinline__ float fsum(fval F) { return F; }
__device__ __forceinline__ void faddlo(fval& F, float v) { F += v; }
#else
typedef float2 fval;
__device__ __forceinline__ fval fzero() { return make_float2(0.f, 0.f); }
__device__ __forceinline__ void facc(fval& F, float2 s, float2 d) { F = __ffma2_rn(s, d, F); }
__device__ __forceinline__ float fsum(fval F) { return F.x + F.y; }
__device__ __forceinline__ void faddlo(fval& F, float v) { F.x += v; }
#endif

// LUT (repulsive mode, few Lennard-Jones classes): w_ij^2 of the two packed pairs comes from a shared-memory table
// indexed by (class of the i-atom, classes of the j-atom pair) - one IADD + one LDS.64 instead of three packed
// FMA-pipe operations (w_ij: 2, t = w/d: 1): (w/d)^2 = w_ij^2 * rcp(r^2), 21 FMA-pipe operations per pair instead of 24.
// lrow[a]: byte offset of the table row of i-atom a; pc: byte offsets of the lane's two j-atom pairs within a row.
struct NbLut { const char* base; uint32_t lrow[NB_NI]; uint2 pc; };
__device__ __forceinline__ float2 nb_lut_w2(const NbLut& L, int a, int p) {
#if NB_KO & 64
    return make_float2(__uint_as_float(L.lrow[a]), __uint_as_float(p ? L.pc.y : L.pc.x));
#endif
    return *reinterpret_cast<const float2*>(L.base + (L.lrow[a] + (p ? L.pc.y : L.pc.x)));
}

template <bool GRAD, bool MASKED, bool FULL, bool LUT = false>
__device__ __forceinline__ bool nb_micro2(const ival (&xi)[NB_NI][4], const ival (&hRi)[NB_NI], const ival (&sei)[NB_NI],
                                           const float4 (&cj)[6], uint32_t m16,
                                           fval (&Fi)[NB_NI][3], float2 (&Fj)[2][3], float2& Sl, float2& Sc, const NbLut& L) {
    static_assert(!(LUT && FULL), "the table form is the repulsive mode's");
    constexpr bool WF = NB_WFORM && !FULL;
    float rmin = 3.0e38f;
    const float2 neg1 = dup(-1.f), one = dup(1.f);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const float2 xj = p ? f2(cj[0].z, cj[0].w) : f2(cj[0].x, cj[0].y);
        const float2 yj = p ? f2(cj[1].z, cj[1].w) : f2(cj[1].x, cj[1].y);
        const float2 zj = p ? f2(cj[2].z, cj[2].w) : f2(cj[2].x, cj[2].y);
        const float2 qj = p ? f2(cj[3].z, cj[3].w) : f2(cj[3].x, cj[3].y);
        const float2 hj = p ? f2(cj[4].z, cj[4].w) : f2(cj[4].x, cj[4].y);
        const float2 sj = p ? f2(cj[5].z, cj[5].w) : f2(cj[5].x, cj[5].y);
#pragma unroll
        for (int a = 0; a < NB_NI; ++a) {
            float2 dx = isub(xj, xi[a][0]), dy = isub(yj, xi[a][1]), dz = isub(zj, xi[a][2]);
            float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            // WF: (hRi, sei) = (a_i, b_i), (hj, sj) = (a_j, b_j); e12 holds w_ij.  LUT: Rij holds w_ij^2 (e12 unused)
            float2 e12 = LUT ? dup(1.f) : WF ? ifma(sei[a], hj, imul(hRi[a], sj)) : imul(sei[a], sj);
            float2 qq = imul(xi[a][3], qj);
            float2 Rij = LUT ? nb_lut_w2(L, a, p) : WF ? e12 : iadd(hRi[a], hj);
            if (MASKED) {
                const bool on0 = (m16 >> (4 * a + 2 * p)) & 1u, on1 = (m16 >> (4 * a + 2 * p + 1)) & 1u;
                r2.x = on0 ? r2.x : NB_MASKED_R2; r2.y = on1 ? r2.y : NB_MASKED_R2;
                if (WF) { Rij.x = on0 ? Rij.x : 0.f; Rij.y = on1 ? Rij.y : 0.f; }
                else { e12.x = on0 ? e12.x : 0.f; e12.y = on1 ? e12.y : 0.f; }
                qq.x = on0 ? qq.x : 0.f; qq.y = on1 ? qq.y : 0.f;
            }
#if NB_KO & 2
            rmin = r2.x;
#else
            rmin = fmin3(rmin, r2.x, r2.y);
#endif
#if NB_KO & 1
            const float2 r2c = r2;
#else
            const float2 r2c = f2(fmaxf(r2.x, NB_THR2), fmaxf(r2.y, NB_THR2));
#endif
#if NB_KO & 32
            float2 inv = __fmul2_rn(r2c, dup(0.01f));
#else
            float2 inv = f2(rsqrt_fast(r2c.x), rsqrt_fast(r2c.y));
#endif
            float2 rinv2 = dup(0.f);
            float2 t2;
#if (NB_KO & 48) || !NB2_RCP_LUT
            if (LUT) { rinv2 = __fmul2_rn(inv, inv); t2 = __fmul2_rn(Rij, rinv2); }
#else
            if (LUT) { rinv2 = f2(rcp_fast(r2c.x), rcp_fast(r2c.y)); t2 = __fmul2_rn(Rij, rinv2); }
#endif
            else { const float2 t = __fmul2_rn(Rij, inv); t2 = __fmul2_rn(t, t); }
            float2 t6 = __fmul2_rn(__fmul2_rn(t2, t2), t2);
            float2 u6 = WF ? t6 : __fmul2_rn(e12, t6);
            float2 ec = __fmul2_rn(qq, inv);
            float2 u;
            if (FULL) {
                u = __ffma2_rn(u6, __ffma2_rn(t6, one, neg1), ec);
                Sl = __ffma2_rn(u6, __ffma2_rn(t6, one, dup(-2.f)), Sl);
            } else {
                u = __ffma2_rn(u6, t6, ec);
#if !(NB_KO & 8)
                Sl = __ffma2_rn(u6, t6, Sl);
#endif
            }
#if !(NB_KO & 8)
#if NB_REUSE_ORDER
            Sc = __ffma2_rn(ec, one, Sc);      // two register pairs read instead of three (qq, inv, Sc)
#else
            Sc = __ffma2_rn(qq, inv, Sc);
#endif
#endif
            if (GRAD) {
#if NB2_RCP
                float2 s = __fmul2_rn(u, LUT ? rinv2 : f2(rcp_fast(r2c.x), rcp_fast(r2c.y)));   // 1/r^2 on the (idle) MUFU pipe
#else
                float2 s = __fmul2_rn(u, LUT ? rinv2 : __fmul2_rn(inv, inv));
#endif
#if NB_KO & 256
                // knock-out: no accumulator operand (two register pairs read per operation instead of three)
                Fi[a][0] = __fmul2_rn(s, dx); Fj[p][0] = __fmul2_rn(s, dx);
                Fi[a][1] = __fmul2_rn(s, dy); Fj[p][1] = __fmul2_rn(s, dy);
                Fi[a][2] = __fmul2_rn(s, dz); Fj[p][2] = __fmul2_rn(s, dz);
#elif NB_KO & 512
                // knock-out: the j-side accumulation only (half of the force operations)
                Fj[p][0] = __ffma2_rn(s, dx, Fj[p][0]); Fj[p][1] = __ffma2_rn(s, dy, Fj[p][1]); Fj[p][2] = __ffma2_rn(s, dz, Fj[p][2]);
#elif NB_REUSE_ORDER
                // the i- and the j-side accumulation of a component share two of their three operands
                facc(Fi[a][0], s, dx); Fj[p][0] = __ffma2_rn(s, dx, Fj[p][0]);
                facc(Fi[a][1], s, dy); Fj[p][1] = __ffma2_rn(s, dy, Fj[p][1]);
                facc(Fi[a][2], s, dz); Fj[p][2] = __ffma2_rn(s, dz, Fj[p][2]);
#else
                facc(Fi[a][0], s, dx); facc(Fi[a][1], s, dy); facc(Fi[a][2], s, dz);
                Fj[p][0] = __ffma2_rn(s, dx, Fj[p][0]); Fj[p][1] = __ffma2_rn(s, dy, Fj[p][1]); Fj[p][2] = __ffma2_rn(s, dz, Fj[p][2]);
#endif
            }
        }
    }
    return !(rmin >= NB_THR2);
}

// All 16 pairs of a micro-tile evaluated EXACTLY (any distance), branch-free and packed like nb_micro2, without the fast
// path: the second loop of nb_kernel2 uses it for items dominated by close pairs (collapsed structures - an untrained
// decoder's output - where "fast path + correction" would do the work twice).  Per pair: 38 FMA-pipe lane-operations and
// 4 MUFU (rsqrt, 2 x ex2, one shared rcp).  w(d,k) = elu(d-k)+k (torch_protein_energy.py:238-239).
__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool GRAD, bool MASKED, bool FULL, bool LUT = false>
__device__ __forceinline__ bool nb_micro2_exact(const ival (&xi)[NB_NI][4], const ival (&hRi)[NB_NI], const ival (&sei)[NB_NI],
                                                 const float4 (&cj)[6], uint32_t m16,
                                                 fval (&Fi)[NB_NI][3], float2 (&Fj)[2][3], float2& Sl, float2& Sc, const NbLut& L) {
    constexpr bool WF = NB_WFORM && !FULL;
    constexpr float LOG2E = 1.4426950408889634f;
    float rmin = 3.0e38f;
    const float2 neg1 = dup(-1.f), one = dup(1.f);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const float2 xj = p ? f2(cj[0].z, cj[0].w) : f2(cj[0].x, cj[0].y);
        const float2 yj = p ? f2(cj[1].z, cj[1].w) : f2(cj[1].x, cj[1].y);
        const float2 zj = p ? f2(cj[2].z, cj[2].w) : f2(cj[2].x, cj[2].y);
        const float2 qj = p ? f2(cj[3].z, cj[3].w) : f2(cj[3].x, cj[3].y);
        const float2 hj = p ? f2(cj[4].z, cj[4].w) : f2(cj[4].x, cj[4].y);
        const float2 sj = p ? f2(cj[5].z, cj[5].w) : f2(cj[5].x, cj[5].y);
#pragma unroll
        for (int a = 0; a < NB_NI; ++a) {
            float2 dx = isub(xj, xi[a][0]), dy = isub(yj, xi[a][1]), dz = isub(zj, xi[a][2]);
            float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            float2 e12 = LUT ? dup(1.f) : WF ? ifma(sei[a], hj, imul(hRi[a], sj)) : imul(sei[a], sj);
            float2 qq = imul(xi[a][3], qj);
            float2 Rij = LUT ? nb_lut_w2(L, a, p) : WF ? e12 : iadd(hRi[a], hj);
            if (MASKED) {
                const bool on0 = (m16 >> (4 * a + 2 * p)) & 1u, on1 = (m16 >> (4 * a + 2 * p + 1)) & 1u;
                r2.x = on0 ? r2.x : NB_MASKED_R2; r2.y = on1 ? r2.y : NB_MASKED_R2;
                if (WF) { Rij.x = on0 ? Rij.x : 0.f; Rij.y = on1 ? Rij.y : 0.f; }
                else { e12.x = on0 ? e12.x : 0.f; e12.y = on1 ? e12.y : 0.f; }
                qq.x = on0 ? qq.x : 0.f; qq.y = on1 ? qq.y : 0.f;
            }
            rmin = fmin3(rmin, r2.x, r2.y);
            // d = r2 * rsqrt(r2), one Newton step; coincident atoms: 1/d := 0 (zero force, like cdist's backward); NaN stays NaN
            const float2 y0 = f2(rsqrt_fast(r2.x), rsqrt_fast(r2.y));
            const float2 hn = __fmul2_rn(__fmul2_rn(r2, dup(-0.5f)), y0);
            float2 dinv = __fmul2_rn(y0, __ffma2_rn(hn, y0, dup(1.5f)));
            dinv.x = r2.x > 0.f ? dinv.x : 0.f; dinv.y = r2.y > 0.f ? dinv.y : 0.f;
            const float2 d = __fmul2_rn(r2, dinv);
            const float2 a1 = __ffma2_rn(d, dup(LOG2E), dup(-LJ_K * LOG2E)), a2 = __ffma2_rn(d, dup(LOG2E), dup(-C_K * LOG2E));
            const float2 ex1 = f2(ex2_fast(a1.x), ex2_fast(a1.y)), ex2 = f2(ex2_fast(a2.x), ex2_fast(a2.y));
            const float2 n1 = __ffma2_rn(ex1, one, dup(LJ_K - 1.f)), n2 = __ffma2_rn(ex2, one, dup(C_K - 1.f));
            const bool f1x = d.x > LJ_K, f1y = d.y > LJ_K, f2x = d.x > C_K, f2y = d.y > C_K;
            const float2 w1 = f2(f1x ? d.x : n1.x, f1y ? d.y : n1.y), w2 = f2(f2x ? d.x : n2.x, f2y ? d.y : n2.y);
            const float2 ww = __fmul2_rn(w1, w2);
            const float2 iw = f2(rcp_fast(ww.x), rcp_fast(ww.y));
            const float2 iw1 = __fmul2_rn(iw, w2), iw2 = __fmul2_rn(iw, w1);
            const float2 t = __fmul2_rn(Rij, iw1);
            const float2 t2 = LUT ? __fmul2_rn(Rij, __fmul2_rn(iw1, iw1)) : __fmul2_rn(t, t);
            const float2 t6 = __fmul2_rn(__fmul2_rn(t2, t2), t2);
            const float2 u6 = WF ? t6 : __fmul2_rn(e12, t6);
            const float2 ec = __fmul2_rn(qq, iw2);
            float2 ulj;
            if (FULL) {
                ulj = __fmul2_rn(u6, __ffma2_rn(t6, one, neg1));
                Sl = __ffma2_rn(u6, __ffma2_rn(t6, one, dup(-2.f)), Sl);
            } else {
                ulj = __fmul2_rn(u6, t6);
                Sl = __ffma2_rn(u6, t6, Sl);
            }
            Sc = __ffma2_rn(qq, iw2, Sc);
            if (GRAD) {
                const float2 dw1 = f2(f1x ? 1.f : ex1.x, f1y ? 1.f : ex1.y), dw2 = f2(f2x ? 1.f : ex2.x, f2y ? 1.f : ex2.y);
                const float2 gl = __fmul2_rn(__fmul2_rn(ulj, iw1), dw1), gc = __fmul2_rn(__fmul2_rn(ec, iw2), dw2);
                const float2 sf = __fmul2_rn(__ffma2_rn(gl, one, gc), dinv);
                facc(Fi[a][0], sf, dx); facc(Fi[a][1], sf, dy); facc(Fi[a][2], sf, dz);
                Fj[p][0] = __ffma2_rn(sf, dx, Fj[p][0]); Fj[p][1] = __ffma2_rn(sf, dy, Fj[p][1]); Fj[p][2] = __ffma2_rn(sf, dz, Fj[p][2]);
            }
        }
    }
    return !(rmin >= NB_THR2);
}

template <bool GRAD, bool FULL, bool LUT>
__global__ void __launch_bounds__(32, 4 * (LUT ? NB2_MINB_LUT : NB2_MINB))
nb_kernel2(const float* __restrict__ x, long sxb, long sxc, long sxn, const float* __restrict__ q, int N,
           const float2* __restrict__ par, const int4* __restrict__ pairs,
           const uint32_t* __restrict__ masks, const uint32_t* __restrict__ stepbits, int n_pairs, int B,
           float* __restrict__ fpart, const float* __restrict__ weights, double* __restrict__ e_part,
           const uint32_t* __restrict__ lrow /* [Npad] */, const uint2* __restrict__ lpc /* [Npad/4] */,
           const float2* __restrict__ lut, int lut_n, int e_stride /* energy partials per conformation */,
           int e_off /* of this launch's first pair (row bands) */) {
    constexpr int NC = LUT ? 4 : 6;                // LUT: (x, y, z, q) only; the pair codes replace the two LJ parameters
    __shared__ float4 s_j[NC][32];                 // component c of the 128 j-atoms: s_j[c][g] = atoms 4g..4g+3
    __shared__ uint2 s_pc[LUT ? 32 : 1];           // LUT: table offsets of the two j-atom pairs of group g
    __shared__ uint32_t s_m[NB_T * 4];             // uint16 [g][lane] pair masks, as words
    extern __shared__ __align__(16) float2 s_lut[];   // LUT: [class i][class j0][class j1] -> (w^2(i,j0), w^2(i,j1))
    const int lane = threadIdx.x;
    // grid (B, y, z): conformation in x, tile pair p = z * gridDim.y + y - the launch order (x fastest) is the pair-major
    // order of the item list (longest first), without an integer division in every warp
    const int b = blockIdx.x, p = blockIdx.z * gridDim.y + blockIdx.y;
    if (p >= n_pairs) return;
    if (weights && !e_part && weights[4 * b + 3] == 0.f) return;
    const int4 IJ = pairs[p];
    const long item = (long)b * n_pairs + p;
    const int i0 = IJ.x * NB_T, j0 = IJ.y * NB_T;
    const float* xb = x + (long)b * sxb;
    const bool masked = IJ.z >= 0;

    ival xi[NB_NI][4], hRi[NB_NI], sei[NB_NI];
    fval Fi[NB_NI][3];
    NbLut L;
    L.base = reinterpret_cast<const char*>(s_lut);
    L.pc = make_uint2(0u, 0u);
    if (LUT) {
        const uint4 r4 = reinterpret_cast<const uint4*>(lrow + i0)[lane];
        L.lrow[0] = r4.x; L.lrow[1] = r4.y; L.lrow[2] = r4.z; L.lrow[3] = r4.w;
        for (int k = lane; k < lut_n; k += 32) s_lut[k] = lut[k];
        s_pc[lane] = lpc[(j0 >> 2) + lane];
    }
    // Element offsets inside one conformation are 32-bit (mlp_energy_eval checks the extent) and are stepped by
    // additions: the prologue of an item runs next to two warps that keep the FMA pipe busy, where every IMAD of a
    // 64-bit "n * stride" waited for the pipe (12 % of the kernel's warp-time was spent on 45 such instructions).
    const int sn = (int)sxn, sc = (int)sxc;
    const float* pi = xb + (i0 + 4 * lane) * sn;
    const float4 qi4 = reinterpret_cast<const float4*>(q + i0)[lane];
    const float qiv[4] = {qi4.x, qi4.y, qi4.z, qi4.w};
#pragma unroll
    for (int a = 0; a < NB_NI; ++a, pi += sn) {
        const int n = i0 + 4 * lane + a;
        float4 v = make_float4(1.0e6f + 100.f * (float)(n - N), -1.0e6f, 1.0e6f, 0.f);   // padding atoms: parked far away (nb_load_atom)
        if (n < N) { const float* pc = pi + sc; v = make_float4(pi[0], pc[0], pc[sc], qiv[a]); }
        xi[a][0] = imake(v.x); xi[a][1] = imake(v.y); xi[a][2] = imake(v.z); xi[a][3] = imake(v.w);
        if (!LUT) {
            const float2 pr = par[i0 + 4 * lane + a];
            hRi[a] = imake(pr.x); sei[a] = imake(pr.y);
            L.lrow[a] = 0u;
        } else hRi[a] = sei[a] = imake(0.f);
        Fi[a][0] = Fi[a][1] = Fi[a][2] = fzero();
    }
    {
        float* sf = reinterpret_cast<float*>(s_j);
        const float* pj = xb + (j0 + lane) * sn;
        const int sn32 = sn << 5;
#pragma unroll
        for (int k = 0; k < NB_T / 32; ++k, pj += sn32) {
            const int jj = lane + 32 * k, n = j0 + jj;
            float4 v = make_float4(1.0e6f + 100.f * (float)(n - N), -1.0e6f, 1.0e6f, 0.f);
            if (n < N) { const float* pc = pj + sc; v = make_float4(pj[0], pc[0], pc[sc], q[n]); }
            sf[0 * NB_T + jj] = v.x; sf[1 * NB_T + jj] = v.y; sf[2 * NB_T + jj] = v.z; sf[3 * NB_T + jj] = v.w;
            if (!LUT) {
                const float2 pr = par[j0 + jj];
                sf[4 * NB_T + jj] = pr.x; sf[5 * NB_T + jj] = pr.y;
            }
        }
    }
    if (masked) {
        const uint4* src = reinterpret_cast<const uint4*>(masks + (long)IJ.z * (NB_T * 4));
        uint4* dst = reinterpret_cast<uint4*>(s_m);
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[lane + 32 * k] = src[lane + 32 * k];
    }
    __syncwarp();

    float2 Sl = dup(0.f), Sc = dup(0.f);
    float2 Fj[2][3];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) Fj[pp][0] = Fj[pp][1] = Fj[pp][2] = dup(0.f);
    const int src_lane = (lane + 1) & 31;
    const uint16_t* sm16 = reinterpret_cast<const uint16_t*>(s_m);
    const int r_begin = IJ.w & 0xff, n_steps = IJ.w >> 8;    // rotation steps [r_begin, n_steps) of this tile pair
    // bit r: rotation step r of this tile pair has an excluded pair in some lane (a band tile has 3 such steps of 32,
    // a diagonal tile 4-5 of 17 for backbone systems); the other steps run the mask-free micro-tile
    const uint32_t sbits = masked ? stepbits[IJ.z] : 0u;
    // exact correction of a micro-tile's close pairs (rare) through the scalar path, on un-packed views of the lane's data
    auto fix_tile = [&](const float4 (&cj)[6], uint32_t m16) {
        float4 ui[NB_NI], uj[4];
        float2 upi[NB_NI], upj[4];
        float3 dFi[NB_NI], dFj[4];
        float w2m[NB_NI][4];
        float dSl = 0.f, dSc = 0.f;
#pragma unroll
        for (int a = 0; a < NB_NI; ++a) {
            ui[a] = make_float4(ilo(xi[a][0]), ilo(xi[a][1]), ilo(xi[a][2]), ilo(xi[a][3]));
            upi[a] = make_float2(ilo(hRi[a]), ilo(sei[a]));
            dFi[a] = make_float3(0.f, 0.f, 0.f);
            if (LUT) {
                const float2 w0 = nb_lut_w2(L, a, 0), w1 = nb_lut_w2(L, a, 1);
                w2m[a][0] = w0.x; w2m[a][1] = w0.y; w2m[a][2] = w1.x; w2m[a][3] = w1.y;
            }
        }
        const float* cf = reinterpret_cast<const float*>(cj);
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            uj[bb] = make_float4(cf[0 * 4 + bb], cf[1 * 4 + bb], cf[2 * 4 + bb], cf[3 * 4 + bb]);
            upj[bb] = LUT ? make_float2(0.f, 0.f) : make_float2(cf[4 * 4 + bb], cf[5 * 4 + bb]);
            dFj[bb] = make_float3(0.f, 0.f, 0.f);
        }
        nb_micro_fix<true, FULL, LUT ? 2 : (NB_WFORM && !FULL)>(ui, upi, uj, upj, m16, dFi, dFj, dSl, dSc, w2m);
#pragma unroll
        for (int a = 0; a < NB_NI; ++a) { faddlo(Fi[a][0], dFi[a].x); faddlo(Fi[a][1], dFi[a].y); faddlo(Fi[a][2], dFi[a].z); }
        Fj[0][0].x += dFj[0].x; Fj[0][1].x += dFj[0].y; Fj[0][2].x += dFj[0].z;
        Fj[0][0].y += dFj[1].x; Fj[0][1].y += dFj[1].y; Fj[0][2].y += dFj[1].z;
        Fj[1][0].x += dFj[2].x; Fj[1][1].x += dFj[2].y; Fj[1][2].x += dFj[2].z;
        Fj[1][0].y += dFj[3].x; Fj[1][1].y += dFj[3].y; Fj[1][2].y += dFj[3].z;
        Sl.x += dSl; Sc.x += dSc;
    };
    auto rotate = [&]() {
        if (GRAD && !(NB_KO & 4)) {
#pragma unroll
            for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    Fj[pp][c].x = __shfl_sync(0xffffffffu, Fj[pp][c].x, src_lane);
                    Fj[pp][c].y = __shfl_sync(0xffffffffu, Fj[pp][c].y, src_lane);
                }
        }
    };
    int r = r_begin;
    constexpr int kStepUnroll = NB_STEP_UNROLL;
    float4 cj[6];
    cj[4] = cj[5] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_j = [&](int rr) {       // the lane's j-group of rotation step rr
        const int g = (lane + rr) & 31;
#pragma unroll
        for (int c = 0; c < NC; ++c) cj[c] = s_j[c][g];
        if (LUT) L.pc = s_pc[g];
    };
#if NB_PIPE_J
    load_j(r);
#endif
#pragma unroll kStepUnroll
    for (; r < n_steps; ++r) {
        const int g = (lane + r) & 31;
#if !NB_PIPE_J
        load_j(r);
#endif
        const bool mstep = (sbits >> r) & 1u;
        const uint32_t m16 = mstep ? sm16[g * 32 + lane] : 0xffffu;
        const bool close = mstep ? nb_micro2<GRAD, true, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L)
                                  : nb_micro2<GRAD, false, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L);
        if (close) fix_tile(cj, m16);
#if NB_PIPE_J
        load_j(r + 1);        // next step's shared-memory loads are issued before the shuffles of the rotation (their latencies overlap); the group past the last step is loaded and not used
#endif
        rotate();
        // collapsed structures (an untrained decoder's output): when at least half of the lanes met a close pair in the
        // item's first step, the remaining steps are evaluated exactly, without the fast path (second loop)
        if (NB_DENSE_MODE && r == r_begin && __popc(__ballot_sync(0xffffffffu, close)) >= 16) { ++r; break; }
    }
#pragma unroll 1
    for (; r < n_steps; ++r) {
        const int g = (lane + r) & 31;
        load_j(r);
        const bool mstep = (sbits >> r) & 1u;
        const uint32_t m16 = mstep ? sm16[g * 32 + lane] : 0xffffu;
        if (mstep) nb_micro2_exact<GRAD, true, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L);
        else nb_micro2_exact<GRAD, false, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L);
        rotate();
    }
    if (GRAD) {
        float* out = fpart + item * (6 * NB_T);
        const int gl = (lane + n_steps) & 31;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            reinterpret_cast<float4*>(out + c * NB_T)[lane] =
                make_float4(-fsum(Fi[0][c]), -fsum(Fi[1][c]), -fsum(Fi[2][c]), -fsum(Fi[3][c]));
            reinterpret_cast<float4*>(out + (3 + c) * NB_T)[gl] = make_float4(Fj[0][c].x, Fj[0][c].y, Fj[1][c].x, Fj[1][c].y);
        }
    }
    if (e_part) {
        double e = (double)warp_sum(Sl.x + Sl.y) * (1.0 / 12.0) + (double)warp_sum(Sc.x + Sc.y);
        if (lane == 0) e_part[(long)b * e_stride + e_off + p] = e;
    }
}

// ---- block-item variant: large batches x systems ------------------------------------------------------------
// nb_kernel2 gives every 128 x 128 tile pair its own warp and its own pair of partial-force blocks: the finest
// granularity (what a MurD-sized batch of 64 needs to fill 148 SMs evenly), at the price of 3 KB of scratch per tile
// pair and conformation and of one prologue per 32 rotation steps.  When there is much more work than warp slots the
// items below are used instead: one warp walks a BLOCK of tile pairs, rows I0..I0+R-1 times columns J0..J0+C-1,
// row by row;
//   * the row's 128 i-atoms and their force accumulators stay in registers across the row's C tiles and are written
//     once per row;
//   * the j-atom forces of a column are accumulated over the block's R rows in shared memory (fixed order, one warp:
//     still reproducible) and written once per column - R + C partial blocks instead of 2 R C;
//   * the next tile's j-atoms (coordinates with the caller's strides, packed per-tile parameters, exclusion masks) are
//     fetched with cp.async into the other half of a double buffer while the current tile is evaluated.
// Items are either rectangles strictly above the diagonal band (R <= NB_BLK_R, C <= NB_BLK_C) or single rows
// (R = 1: the diagonal tile and the tiles up to the next rectangle).  Scratch per conformation drops from
// 3 KB x tile pairs to about 1.5 KB x tile pairs x (1/R + 1/C).
#ifndef NB_LUT_MAXC
#define NB_LUT_MAXC 7   // most Lennard-Jones classes the table form is used for (8^3 entries of 8 bytes = 4 KB per warp)
#endif
#ifndef NB2B_MINB
#define NB2B_MINB NB2_MINB   // block-item kernel: resident warps per SM sub-partition the register budget is sized for
#endif
#define NB_BLK_R 8    // largest block height (NB_BLK_ROWS)
#ifndef NB_BLK_C
#define NB_BLK_C 4
#endif
static_assert(NB_BLK_R * NB_BLK_C <= 32, "the lanes of the warp preload the meta data of the item's tiles");
struct NbBlockItem { int I0, R, J0, C, iblk, jblk, pad0, pad1; };   // iblk / jblk: first partial block of the rows / columns
static_assert(sizeof(NbBlockItem) == 32, "record layout");

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <bool GRAD, bool FULL, bool LUT>
__global__ void __launch_bounds__(32, 4 * (LUT ? NB2_MINB_LUT : NB2B_MINB))
nb_block_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, int N, int nt,
                const float* __restrict__ tpar /* [nt][3][128]: charge, then the two per-atom LJ parameters; LUT: [nt][320] charge,
                                                   table row offsets of the 128 atoms, pair offsets of the 32 groups */,
                const NbBlockItem* __restrict__ items, const int2* __restrict__ pair_meta /* [nt(nt+1)/2]: mask index | -1, step bits */,
                const uint32_t* __restrict__ masks, int n_items, int n_blocks, int B,
                float* __restrict__ fpart, const float* __restrict__ weights, double* __restrict__ e_part,
                const float2* __restrict__ lut, int lut_n, int acc_floats) {
    constexpr int TPW = LUT ? NB_T + NB_T + NB_T / 2 : 3 * NB_T;   // words of per-tile parameters
    constexpr int NC = LUT ? 4 : 6;
    __shared__ __align__(16) float s_j[2][6 * NB_T];          // double buffer: component c of the 128 j-atoms at [c*128 + jj]
    __shared__ __align__(16) uint32_t s_m[2][NB_T * 4];       // uint16 [g][lane] pair masks of a masked tile, as words
    extern __shared__ __align__(16) float s_acc[];            // [C][3][128] j-force accumulators of a rectangle (R > 1)
    __shared__ int2 s_meta[32];                                // (mask index | -1, step bits) of the item's tiles
    const int lane = threadIdx.x;
    const int b = blockIdx.x, q = blockIdx.z * gridDim.y + blockIdx.y;   // grid (B, y, z): item-major launch order, no integer division
    if (q >= n_items) return;
    if (weights && !e_part && weights[4 * b + 3] == 0.f) return;
    const int sn = (int)sxn, sc = (int)sxc, sn32 = sn << 5;   // 32-bit element offsets, stepped by additions (see nb_kernel2)
    const NbBlockItem it = items[q];
    const float* xb = x + (long)b * sxb;
    const int n_tile = it.R * it.C;
    // tile t of the item = (row t / C, column t % C)
    {
        int2 m = make_int2(-1, 0);
        if (lane < n_tile) {
            const int I = it.I0 + lane / it.C, J = it.J0 + lane % it.C;
            m = pair_meta[I * nt - (I * (I - 1)) / 2 + (J - I)];
        }
        s_meta[lane] = m;
        __syncwarp();
    }
    auto issue = [&](int t, int buf) {
        const int J = it.J0 + t % it.C;
        const int j0 = J * NB_T;
        float* dst = s_j[buf];
        const float* p = xb + (j0 + lane) * sn;
#pragma unroll
        for (int k = 0; k < NB_T / 32; ++k, p += sn32) {
            const int jj = lane + 32 * k, n = j0 + jj;
            if (n < N) {
                const float* pc = p + sc;
                cp_async4(dst + jj, p); cp_async4(dst + NB_T + jj, pc); cp_async4(dst + 2 * NB_T + jj, pc + sc);
            } else {   // padding of the last tile: parked far away (zero charge / LJ parameters in tpar)
                dst[jj] = 1.0e6f + 100.f * (float)(n - N); dst[NB_T + jj] = -1.0e6f; dst[2 * NB_T + jj] = 1.0e6f;
            }
        }
        const float* src = tpar + (long)J * TPW;
        if (LUT) {      // charges of the 128 atoms, then the pair offsets of the 32 groups (skipping the row offsets)
            cp_async16(dst + 3 * NB_T + 4 * lane, src + 4 * lane);
            if (lane < 16) cp_async16(dst + 4 * NB_T + 4 * lane, src + 2 * NB_T + 4 * lane);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) cp_async16(dst + 3 * NB_T + 4 * (lane + 32 * k), src + 4 * (lane + 32 * k));
        }
        const int mi = s_meta[t].x;
        if (mi >= 0) {
            const uint32_t* ms = masks + (long)mi * (NB_T * 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) cp_async16(s_m[buf] + 4 * (lane + 32 * k), ms + 4 * (lane + 32 * k));
        }
        cp_async_commit();
    };
    issue(0, 0);
    if (n_tile > 1) issue(1, 1); else cp_async_commit();

    float2 Sl = dup(0.f), Sc = dup(0.f);
    NbLut L{};
    if (LUT) {       // the type-pair table sits behind the j-force accumulators
        float2* s_lut = reinterpret_cast<float2*>(s_acc + acc_floats);
        for (int k = lane; k < lut_n; k += 32) s_lut[k] = lut[k];
        L.base = reinterpret_cast<const char*>(s_lut);
        __syncwarp();
    }
    ival xi[NB_NI][4], hRi[NB_NI], sei[NB_NI];
    fval Fi[NB_NI][3];
    const int src_lane = (lane + 1) & 31;
    int t = 0;
    for (int r = 0; r < it.R; ++r) {
        const int I = it.I0 + r, i0 = I * NB_T;
        {
            const float4* tp = reinterpret_cast<const float4*>(tpar + (long)I * TPW);
            const float4 q4 = tp[lane], h4 = tp[32 + lane], s4 = LUT ? h4 : tp[64 + lane];
            const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, hv[4] = {h4.x, h4.y, h4.z, h4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
            if (LUT) { L.lrow[0] = __float_as_uint(h4.x); L.lrow[1] = __float_as_uint(h4.y); L.lrow[2] = __float_as_uint(h4.z); L.lrow[3] = __float_as_uint(h4.w); }
            const float* p = xb + (i0 + 4 * lane) * sn;
#pragma unroll
            for (int a = 0; a < NB_NI; ++a, p += sn) {
                const int n = i0 + 4 * lane + a;
                float vx = 1.0e6f + 100.f * (float)(n - N), vy = -1.0e6f, vz = 1.0e6f;
                if (n < N) { const float* pc = p + sc; vx = p[0]; vy = pc[0]; vz = pc[sc]; }
                xi[a][0] = imake(vx); xi[a][1] = imake(vy); xi[a][2] = imake(vz); xi[a][3] = imake(qv[a]);
                hRi[a] = imake(LUT ? 0.f : hv[a]); sei[a] = imake(LUT ? 0.f : sv[a]);
                Fi[a][0] = Fi[a][1] = Fi[a][2] = fzero();
            }
        }
        for (int c = 0; c < it.C; ++c, ++t) {
            const int buf = t & 1;
            if (t + 1 < n_tile) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncwarp();
            const int J = it.J0 + c;
            const int2 meta = s_meta[t];
            const uint32_t sbits = meta.x >= 0 ? (uint32_t)meta.y : 0u;   // bit rs: rotation step rs touches an excluded pair
            const int n_steps = (J == I) ? 17 : 32;            // diagonal tile: block (L,g) mirrors (g,L)
            const float4(*sj)[32] = reinterpret_cast<const float4(*)[32]>(s_j[buf]);
            const uint16_t* sm16 = reinterpret_cast<const uint16_t*>(s_m[buf]);
            const uint2* spc = reinterpret_cast<const uint2*>(s_j[buf] + 4 * NB_T);
            float2 Fj[2][3];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) Fj[pp][0] = Fj[pp][1] = Fj[pp][2] = dup(0.f);
            auto fix_tile = [&](const float4 (&cj)[6], uint32_t m16) {
                float4 ui[NB_NI], uj[4];
                float2 upi[NB_NI], upj[4];
                float3 dFi[NB_NI], dFj[4];
                float w2m[NB_NI][4];
                float dSl = 0.f, dSc = 0.f;
#pragma unroll
                for (int a = 0; a < NB_NI; ++a) {
                    ui[a] = make_float4(ilo(xi[a][0]), ilo(xi[a][1]), ilo(xi[a][2]), ilo(xi[a][3]));
                    upi[a] = make_float2(ilo(hRi[a]), ilo(sei[a]));
                    dFi[a] = make_float3(0.f, 0.f, 0.f);
                    if (LUT) {
                        const float2 w0 = nb_lut_w2(L, a, 0), w1 = nb_lut_w2(L, a, 1);
                        w2m[a][0] = w0.x; w2m[a][1] = w0.y; w2m[a][2] = w1.x; w2m[a][3] = w1.y;
                    }
                }
                const float* cf = reinterpret_cast<const float*>(cj);
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    uj[bb] = make_float4(cf[0 * 4 + bb], cf[1 * 4 + bb], cf[2 * 4 + bb], cf[3 * 4 + bb]);
                    upj[bb] = LUT ? make_float2(0.f, 0.f) : make_float2(cf[4 * 4 + bb], cf[5 * 4 + bb]);
                    dFj[bb] = make_float3(0.f, 0.f, 0.f);
                }
                nb_micro_fix<true, FULL, LUT ? 2 : (NB_WFORM && !FULL)>(ui, upi, uj, upj, m16, dFi, dFj, dSl, dSc, w2m);
#pragma unroll
                for (int a = 0; a < NB_NI; ++a) { faddlo(Fi[a][0], dFi[a].x); faddlo(Fi[a][1], dFi[a].y); faddlo(Fi[a][2], dFi[a].z); }
                Fj[0][0].x += dFj[0].x; Fj[0][1].x += dFj[0].y; Fj[0][2].x += dFj[0].z;
                Fj[0][0].y += dFj[1].x; Fj[0][1].y += dFj[1].y; Fj[0][2].y += dFj[1].z;
                Fj[1][0].x += dFj[2].x; Fj[1][1].x += dFj[2].y; Fj[1][2].x += dFj[2].z;
                Fj[1][0].y += dFj[3].x; Fj[1][1].y += dFj[3].y; Fj[1][2].y += dFj[3].z;
                Sl.x += dSl; Sc.x += dSc;
            };
            auto rotate = [&]() {
                if (GRAD) {
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            Fj[pp][cc].x = __shfl_sync(0xffffffffu, Fj[pp][cc].x, src_lane);
                            Fj[pp][cc].y = __shfl_sync(0xffffffffu, Fj[pp][cc].y, src_lane);
                        }
                }
            };
            int rs = 0;
#pragma unroll 1
            for (; rs < n_steps; ++rs) {
                const int g = (lane + rs) & 31;
                float4 cj[6];
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) cj[cc] = sj[cc][g];
                if (LUT) { L.pc = spc[g]; cj[4] = cj[5] = make_float4(0.f, 0.f, 0.f, 0.f); }
                const bool mstep = (sbits >> rs) & 1u;
                const uint32_t m16 = mstep ? sm16[g * 32 + lane] : 0xffffu;
                const bool close = mstep ? nb_micro2<GRAD, true, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L)
                                         : nb_micro2<GRAD, false, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L);
                if (close) fix_tile(cj, m16);
                rotate();
                if (NB_DENSE_MODE && rs == 0 && __popc(__ballot_sync(0xffffffffu, close)) >= 16) { ++rs; break; }
            }
#pragma unroll 1
            for (; rs < n_steps; ++rs) {       // collapsed structures: every pair exactly, no fast path
                const int g = (lane + rs) & 31;
                float4 cj[6];
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) cj[cc] = sj[cc][g];
                if (LUT) { L.pc = spc[g]; cj[4] = cj[5] = make_float4(0.f, 0.f, 0.f, 0.f); }
                const bool mstep = (sbits >> rs) & 1u;
                const uint32_t m16 = mstep ? sm16[g * 32 + lane] : 0xffffu;
                if (mstep) nb_micro2_exact<GRAD, true, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L);
                else nb_micro2_exact<GRAD, false, FULL, LUT>(xi, hRi, sei, cj, m16, Fi, Fj, Sl, Sc, L);
                rotate();
            }
            if (GRAD) {
                const int gl = (lane + n_steps) & 31;          // after n_steps rotations the lane holds the forces of this group
                if (it.R == 1) {
                    float* out = fpart + ((long)b * n_blocks + it.jblk + c) * (3 * NB_T);
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)
                        reinterpret_cast<float4*>(out + cc * NB_T)[gl] = make_float4(Fj[0][cc].x, Fj[0][cc].y, Fj[1][cc].x, Fj[1][cc].y);
                } else {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        float4* a4 = reinterpret_cast<float4*>(s_acc + (c * 3 + cc) * NB_T) + gl;
                        float4 v = make_float4(Fj[0][cc].x, Fj[0][cc].y, Fj[1][cc].x, Fj[1][cc].y);
                        if (r > 0) { const float4 o = *a4; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        *a4 = v;
                    }
                }
            }
            __syncwarp();                                       // every lane is done with this half of the double buffer
            if (t + 2 < n_tile) issue(t + 2, buf); else cp_async_commit();
        }
        if (GRAD) {
            float* out = fpart + ((long)b * n_blocks + it.iblk + r) * (3 * NB_T);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                reinterpret_cast<float4*>(out + cc * NB_T)[lane] =
                    make_float4(-fsum(Fi[0][cc]), -fsum(Fi[1][cc]), -fsum(Fi[2][cc]), -fsum(Fi[3][cc]));
        }
    }
    if (GRAD && it.R > 1) {
        __syncwarp();
        for (int c = 0; c < it.C; ++c) {
            float* out = fpart + ((long)b * n_blocks + it.jblk + c) * (3 * NB_T);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                reinterpret_cast<float4*>(out + cc * NB_T)[lane] = reinterpret_cast<const float4*>(s_acc + (c * 3 + cc) * NB_T)[lane];
        }
    }
    if (e_part) {
        double e = (double)warp_sum(Sl.x + Sl.y) * (1.0 / 12.0) + (double)warp_sum(Sc.x + Sc.y);
        if (lane == 0) e_part[(long)b * n_items + q] = e;
    }
}

// grad[b, n] (+)= w_nb[b] * (sum of the partial forces of every tile pair in n's tile row / column),
// summed in list order: reproducible.
#ifndef RED_PARTS_N
#define RED_PARTS_N 4
#endif
constexpr int RED_PARTS = RED_PARTS_N;     // warps of a reduction CTA: each sums one contiguous share of the tile's block list
__global__ void __launch_bounds__(32 * RED_PARTS)
nb_reduce_kernel(const float* __restrict__ fpart, const int* __restrict__ red_off, const int* __restrict__ red_list,
                 int n_pairs /* energy partials per conformation */, int n_blocks /* partial-force blocks per conformation */,
                 int N, float* __restrict__ grad, long sgb, long sgc, long sgn,
                 const float* __restrict__ weights, int accumulate, int skip_zero_weight,
                 const float* __restrict__ e_bonded, int n_tiles, const double* __restrict__ e_nb,
                 float* __restrict__ energies /* null: no energy output */, uint32_t term_mask) {
    if (energies && blockIdx.x == gridDim.x - 1) {
        // the conformation's energies: fp64 sums of the kernels' partials in a fixed order (this used to be a fifth
        // launch next to the reduction; the last CTA of each conformation's row does it now)
        __shared__ double s_en[RED_PARTS];
        const int b = blockIdx.y, tid = threadIdx.x;
        double s = 0.0;
        if (term_mask & MLP_TERM_NB)
            for (int p = tid; p < n_pairs; p += 32 * RED_PARTS) s += e_nb[(long)b * n_pairs + p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) s_en[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            double v = s_en[0];
#pragma unroll
            for (int q = 1; q < RED_PARTS; ++q) v += s_en[q];
            energies[4 * b + 3] = (float)v;
        }
        if (tid < 3) {
            double v = 0.0;
            if (term_mask & (1u << tid))
                for (int t = 0; t < n_tiles; ++t) v += (double)e_bonded[((long)b * n_tiles + t) * 3 + tid];
            energies[4 * b + tid] = (float)v;
        }
    }
    // CTA = one 128-atom tile of one conformation; a lane owns four consecutive atoms (16-byte loads of the partial
    // blocks), warp `part` sums its contiguous share of the tile's block list, then part 0 adds the RED_PARTS sums in
    // order: a fixed summation tree, reproducible.  30 MB out of L2 for the headline batch in 9 us + launch; the same
    // time as one atom per lane and two warps per 32 atoms, with a quarter of the load instructions.
    __shared__ float4 s_part[RED_PARTS][3][32];
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int T = blockIdx.x;
    const int n = T * NB_T + 4 * lane;
    const float w = weights ? weights[4 * b + 3] : 1.f;
    if (skip_zero_weight && w == 0.f) {          // nb_kernel wrote no partials for this conformation
        if (part == 0 && !accumulate) {
            float* g = grad + (long)b * sgb + (long)n * sgn;
            for (int a = 0; a < 4 && n + a < N; ++a, g += sgn) { g[0] = 0.f; g[sgc] = 0.f; g[2 * sgc] = 0.f; }
        }
        return;
    }
    float4 sx = make_float4(0.f, 0.f, 0.f, 0.f), sy = sx, sz = sx;
    const float4* base = reinterpret_cast<const float4*>(fpart + (long)b * n_blocks * (3 * NB_T)) + lane;
    const int k0 = red_off[T], k1 = red_off[T + 1];
    if (k0 == k1 && accumulate) return;          // a row band that does not touch this tile (CTA-uniform)
    const int per = (k1 - k0 + RED_PARTS - 1) / RED_PARTS;
    const int ka = min(k1, k0 + part * per), kb = min(k1, ka + per);
    auto add4 = [](float4& s, const float4& v) { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; };
    int k = ka;
    for (; k + 4 <= kb; k += 4) {
        float4 vx[4], vy[4], vz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4* p = base + (long)red_list[k + u] * (3 * NB_T / 4);       // entry = block index (nb_kernel2: pair*2 + side)
            vx[u] = p[0]; vy[u] = p[NB_T / 4]; vz[u] = p[2 * NB_T / 4];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { add4(sx, vx[u]); add4(sy, vy[u]); add4(sz, vz[u]); }
    }
    for (; k < kb; ++k) {
        const float4* p = base + (long)red_list[k] * (3 * NB_T / 4);
        add4(sx, p[0]); add4(sy, p[NB_T / 4]); add4(sz, p[2 * NB_T / 4]);
    }
    s_part[part][0][lane] = sx; s_part[part][1][lane] = sy; s_part[part][2][lane] = sz;
    __syncthreads();
    if (part == 0) {
#pragma unroll
        for (int q = 1; q < RED_PARTS; ++q) { add4(sx, s_part[q][0][lane]); add4(sy, s_part[q][1][lane]); add4(sz, s_part[q][2][lane]); }
        const float fx[4] = {sx.x, sx.y, sx.z, sx.w}, fy[4] = {sy.x, sy.y, sy.z, sy.w}, fz[4] = {sz.x, sz.y, sz.z, sz.w};
        float* g = grad + (long)b * sgb + (long)n * sgn;
#pragma unroll
        for (int a = 0; a < 4; ++a, g += sgn) {
            if (n + a < N) {
                if (accumulate) { g[0] = fmaf(w, fx[a], g[0]); g[sgc] = fmaf(w, fy[a], g[sgc]); g[2 * sgc] = fmaf(w, fz[a], g[2 * sgc]); }
                else { g[0] = w * fx[a]; g[sgc] = w * fy[a]; g[2 * sgc] = w * fz[a]; }
            }
        }
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

// energies[b] = (sum over bonded tiles, sum over NB tile pairs), fixed order, fp64 accumulation
__global__ void __launch_bounds__(128)
finalize_kernel(const float* __restrict__ e_bonded, int n_tiles, const double* __restrict__ e_nb, int n_pairs,
                float* __restrict__ energies, uint32_t term_mask) {
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ double red[128];
    double s = 0.0;
    if (term_mask & MLP_TERM_NB)
        for (int p = tid; p < n_pairs; p += 128) s += e_nb[(long)b * n_pairs + p];
    red[tid] = s;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) energies[4 * b + 3] = (float)red[0];
    if (tid < 3) {
        double v = 0.0;
        if (term_mask & (1u << tid))
            for (int t = 0; t < n_tiles; ++t) v += (double)e_bonded[((long)b * n_tiles + t) * 3 + tid];
        energies[4 * b + tid] = (float)v;
    }
}

// RMSD of x*scale+shift to K targets: out[b,k] = sqrt(sum_{n,c} d^2 / N).  One block scores a structure against up
// to RMSD_KC targets, so the structure is read once per RMSD_KC targets (the targets stay in L2).
// paired: one target per structure, target index = b (get_error: structure b against its own reconstruction).
constexpr int RMSD_KC = 8;
__global__ void __launch_bounds__(256)
rmsd_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, float scale, float shift,
            const float* __restrict__ targets, int N, int K, int paired, float* __restrict__ out) {
    const int b = blockIdx.x;
    const int k0 = paired ? b : blockIdx.y * RMSD_KC;
    const int nk = paired ? 1 : min(RMSD_KC, K - k0);
    const float* t = targets + (long)k0 * N * 3;
    const long tstride = (long)N * 3;
    float s[RMSD_KC];
#pragma unroll
    for (int k = 0; k < RMSD_KC; ++k) s[k] = 0.f;
    for (int e = threadIdx.x; e < 3 * N; e += blockDim.x) {
        int n = e / 3, c = e - 3 * n;
        const float v = fmaf(x[(long)b * sxb + c * sxc + (long)n * sxn], scale, shift);
#pragma unroll
        for (int k = 0; k < RMSD_KC; ++k)
            if (k < nk) { const float d = v - t[k * tstride + e]; s[k] = fmaf(d, d, s[k]); }
    }
    __shared__ float red[RMSD_KC][8];
#pragma unroll
    for (int k = 0; k < RMSD_KC; ++k) {
        s[k] = warp_sum(s[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 32 * RMSD_KC && (threadIdx.x >> 5) < nk) {
        const int k = threadIdx.x >> 5, l = threadIdx.x & 31;
        float v = l < (int)(blockDim.x >> 5) ? red[k][l] : 0.f;
        v = warp_sum(v);
        if (l == 0) out[paired ? (long)b : (long)b * K + k0 + k] = sqrtf(v / (float)N);
    }
}

// Largest eigenvalue of a symmetric 4x4 matrix by cyclic Jacobi rotations (fp64, one thread).
__device__ double max_eig_sym4(double a[4][4]) {
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += a[p][q] * a[p][q];
        if (off < 1e-28 * (a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2] + a[3][3] * a[3][3]) || off == 0.0) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {  // A <- A J
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {  // A <- J^T A
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk;
                }
            }
    }
    return fmax(fmax(a[0][0], a[1][1]), fmax(a[2][2], a[3][3]));
}

// Optimally superposed (Kabsch) RMSD of x*scale+shift onto K targets, without forming the rotation:
// rmsd^2 = (G_x + G_t - 2 lambda_max(K_Horn)) / N with the 4x4 quaternion matrix of the 3x3
// cross-covariance (Horn 1987).  Replaces the biobox Molecule.rmsd loop of analyser.py:293-297, 665-671.
// All sums in fp64: the result is a small difference of large numbers.
__global__ void __launch_bounds__(256)
rmsd_aligned_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, float scale, float shift,
                    const float* __restrict__ targets, int N, int K, int paired, float* __restrict__ out) {
    const int b = blockIdx.x, k = paired ? b : blockIdx.y;
    const float* t = targets + (long)k * N * 3;
    float* const o = out + (paired ? (long)b : (long)b * K + k);
    double acc[17];  // sum x (3), sum t (3), sum |x|^2, sum |t|^2, sum x_i t_j (9)
#pragma unroll
    for (int i = 0; i < 17; ++i) acc[i] = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float* px = x + (long)b * sxb + (long)n * sxn;
        double xv[3], tv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { xv[c] = (double)fmaf(px[c * sxc], scale, shift); tv[c] = (double)t[3 * n + c]; }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc[c] += xv[c]; acc[3 + c] += tv[c];
            acc[6] += xv[c] * xv[c]; acc[7] += tv[c] * tv[c];
#pragma unroll
            for (int d = 0; d < 3; ++d) acc[8 + 3 * c + d] += xv[c] * tv[d];
        }
    }
    __shared__ double red[8][17];
#pragma unroll
    for (int i = 0; i < 17; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < 17; ++i) red[threadIdx.x >> 5][i] = acc[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[17];
        for (int i = 0; i < 17; ++i) { s[i] = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s[i] += red[w][i]; }
        const double n = (double)N;
        double cx[3] = {s[0] / n, s[1] / n, s[2] / n}, ct[3] = {s[3] / n, s[4] / n, s[5] / n};
        double gx = s[6] - n * (cx[0] * cx[0] + cx[1] * cx[1] + cx[2] * cx[2]);
        double gt = s[7] - n * (ct[0] * ct[0] + ct[1] * ct[1] + ct[2] * ct[2]);
        double S[3][3];
        for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) S[c][d] = s[8 + 3 * c + d] - n * cx[c] * ct[d];
        double Km[4][4] = {
            {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
            {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
            {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
            {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
        double lam = max_eig_sym4(Km);
        double msd = (gx + gt - 2.0 * lam) / n;
        *o = msd != msd ? (float)msd : (float)sqrt(fmax(msd, 0.0));
    }
}

// Per-structure geometry statistics for the scan reducers scan_bondlength / scan_ca_chirality
// (analyser.py:340-473, 831-861, 980-1014): mean and population std of the lengths of up to 8 groups of atom
// pairs, and the number of quadruples (n, ca, c, cb) whose triple product ((n-ca) x (c-ca)) . (cb-ca) is
// negative.  out[b] = (mean_0, std_0, ..., mean_{G-1}, std_{G-1}, inversions).  fp64 sums.
constexpr int GEOM_MAX_GROUPS = 8;
__global__ void __launch_bounds__(256)
geom_stats_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, float scale,
                  const int2* __restrict__ pairs, const int* __restrict__ pair_group, int n_pairs, int n_groups,
                  const int4* __restrict__ quads, int n_quads, float* __restrict__ out) {
    const int b = blockIdx.x;
    const float* xb = x + (long)b * sxb;
    auto at = [&](int n) { const float* p = xb + (long)n * sxn; return make_float3(p[0] * scale, p[sxc] * scale, p[2 * sxc] * scale); };
    double acc[2 * GEOM_MAX_GROUPS + 1];
#pragma unroll
    for (int i = 0; i < 2 * GEOM_MAX_GROUPS + 1; ++i) acc[i] = 0.0;
    for (int k = threadIdx.x; k < n_pairs; k += blockDim.x) {
        const int2 pr = pairs[k];
        const int gidx = pair_group[k];
        float3 d = at(pr.x) - at(pr.y);
        double len = sqrt((double)d.x * d.x + (double)d.y * d.y + (double)d.z * d.z);
#pragma unroll
        for (int q = 0; q < GEOM_MAX_GROUPS; ++q)
            if (q == gidx) { acc[2 * q] += len; acc[2 * q + 1] += len * len; }
    }
    for (int k = threadIdx.x; k < n_quads; k += blockDim.x) {
        const int4 qd = quads[k];
        float3 ca = at(qd.y);
        float3 v = cross(at(qd.x) - ca, at(qd.z) - ca);
        if (dot(v, at(qd.w) - ca) < 0.f) acc[2 * GEOM_MAX_GROUPS] += 1.0;
    }
    __shared__ double red[8][2 * GEOM_MAX_GROUPS + 1];
#pragma unroll
    for (int i = 0; i < 2 * GEOM_MAX_GROUPS + 1; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < 2 * GEOM_MAX_GROUPS + 1; ++i) red[threadIdx.x >> 5][i] = acc[i];
    __syncthreads();
    if (threadIdx.x <= 2 * n_groups && threadIdx.x <= 2 * GEOM_MAX_GROUPS) {
        // thread t < 2G: statistic t of group t/2; thread 2G: inversions
        const int t = threadIdx.x;
        auto total = [&](int i) { double s = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][i]; return s; };
        float* o = out + (long)b * (2 * n_groups + 1);
        if (t == 2 * n_groups) o[t] = (float)total(2 * GEOM_MAX_GROUPS);
        else {
            // per-group pair counts are recomputed here (cheap) to avoid another accumulator
            const int gq = t >> 1;
            int cnt = 0;
            for (int k = 0; k < n_pairs; ++k) cnt += pair_group[k] == gq;
            double m = cnt ? total(2 * gq) / cnt : 0.0;
            if (t & 1) { double v = cnt ? total(2 * gq + 1) / cnt - m * m : 0.0; o[t] = (float)sqrt(v > 0.0 ? v : 0.0); }
            else o[t] = (float)m;
        }
    }
}

// Internal coordinates of every structure: pair distances, bond angles and dihedral angles over index lists -
// the device form of MolearnAnalysis.get_bondlengths / get_angles / get_dihedrals (analyser.py:418-566 with the
// formulas of :791-861): only [B, n] scalars leave the GPU instead of the decoded [B, N, 3] structures.
// angle(p0,p1,p2) = acos(b0^ . b1^), b0 = p0-p1, b1 = p2-p1; dihedral(p0,p1,p2,p3) = atan2((b1^ x v) . w, v . w) with
// v, w the components of p0-p1 and p3-p2 orthogonal to b1 = p2-p1 (the standard formula; analyser.py:826 multiplies two
// sums where it means the sum of a product).  fp64 arithmetic on the fp32 coordinates: a few thousand values per structure.
__global__ void __launch_bounds__(128)
internal_coords_kernel(const float* __restrict__ x, long sxb, long sxc, long sxn, float scale,
                       const int2* __restrict__ pairs, int n_pairs, const int* __restrict__ triples, int n_triples,
                       const int4* __restrict__ quads, int n_quads, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_all = n_pairs + n_triples + n_quads;
    if (e >= n_all) return;
    const float* xb = x + (long)b * sxb;
    auto at = [&](int n, double (&p)[3]) { const float* q = xb + (long)n * sxn; p[0] = q[0]; p[1] = q[sxc]; p[2] = q[2 * sxc]; };
    auto dotd = [](const double (&a)[3], const double (&c)[3]) { return a[0] * c[0] + a[1] * c[1] + a[2] * c[2]; };
    double v;
    if (e < n_pairs) {
        const int2 pr = pairs[e];
        double p0[3], p1[3];
        at(pr.x, p0); at(pr.y, p1);
        double d[3] = {p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]};
        v = sqrt(dotd(d, d)) * (double)scale;
    } else if (e < n_pairs + n_triples) {
        const int* t = triples + 3 * (e - n_pairs);
        double p0[3], p1[3], p2[3];
        at(t[0], p0); at(t[1], p1); at(t[2], p2);
        double b0[3] = {p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]}, b1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        double c = dotd(b0, b1) / (sqrt(dotd(b0, b0)) * sqrt(dotd(b1, b1)));
        v = acos(fmin(1.0, fmax(-1.0, c)));
        if (c != c) v = c;
    } else {
        const int4 qd = quads[e - n_pairs - n_triples];
        double p0[3], p1[3], p2[3], p3[3];
        at(qd.x, p0); at(qd.y, p1); at(qd.z, p2); at(qd.w, p3);
        double b0[3] = {p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2]}, b1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        double b2[3] = {p3[0] - p2[0], p3[1] - p2[1], p3[2] - p2[2]};
        const double inv = 1.0 / sqrt(dotd(b1, b1));
        b1[0] *= inv; b1[1] *= inv; b1[2] *= inv;
        const double s0 = dotd(b0, b1), s2 = dotd(b2, b1);
        double vv[3] = {b0[0] - s0 * b1[0], b0[1] - s0 * b1[1], b0[2] - s0 * b1[2]};
        double ww[3] = {b2[0] - s2 * b1[0], b2[1] - s2 * b1[1], b2[2] - s2 * b1[2]};
        double cr[3] = {b1[1] * vv[2] - b1[2] * vv[1], b1[2] * vv[0] - b1[0] * vv[2], b1[0] * vv[1] - b1[1] * vv[0]};
        v = atan2(dotd(cr, ww), dotd(vv, ww));
    }
    out[(long)b * n_all + e] = (float)v;
}

// FP32 FMA-pipe probe (bench.py measures the pipe's live peak with it, next to the nominal 148 x 128 x 2 x f):
// 8 independent chains of FFMA (packed = 0) or FFMA2 (packed = 1) per thread, 16 lane-FMAs per chain step.
constexpr int PROBE_BLOCKS = 148 * 8, PROBE_THREADS = 256;
template <int PACKED>
__global__ void __launch_bounds__(PROBE_THREADS) fp32_probe_kernel(float* __restrict__ out, int iters) {
    float2 a[8], b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) a[i] = __ffma2_rn(a[i], b, c);
            else { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * PROBE_THREADS + threadIdx.x] = s;
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

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// A bonded work item: atoms (i,j,k,l) and the ids of the torsion / angle / bond it carries (-1: none).
struct Item { int a[4]; int tors, angle, bond; bool bond_jk; };
struct TileItems { int a0, n_own; std::vector<Item> items; std::vector<int> halo; std::vector<int> cnt; };
struct BondedPlan {
    std::vector<char> t_live;                                  // torsion has a non-zero series
    std::vector<std::array<float, 1 + 2 * NF>> t_coef;         // c0, cosine[NF], sine[NF]
    bool torsion_sines = false;
    std::vector<TileItems> tiles;
    int max_cnt = 1;                                           // max gradient contributions per atom (multiple of 4)
};

// Host-only part of the bonded packing: Fourier coefficients, per-tile work items (angle -> torsion and
// bond -> item matchings), halo lists, per-atom contribution counts.
// returns 1 if the tile size does not fit the kernel's static limits (the caller retries smaller)
int plan_bonded(const mlp_topology& T, int tile_atoms, BondedPlan& plan) {
    const int N = T.n_atoms;
    const int nb = (int)T.bonds.size() / 2, na = (int)T.angles.size() / 3, nt = (int)T.torsions.size() / 4, P = T.P;
    // torsion Fourier coefficients from the fp32-rounded parameters (the reference moves its
    // tables to the device with .float(), torch_protein_energy.py:123)
    std::vector<char>& t_live = plan.t_live;
    std::vector<std::array<float, 1 + 2 * NF>>& t_coef = plan.t_coef;
    t_live.assign(nt, 0);
    t_coef.assign(nt, {});
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
        for (int n = 0; n < NF; ++n)   // sin(fp32(pi)) = -8.7e-8: below 1e-6 of the cosine coefficient it is rounding noise
            if (std::fabs(sg[n]) > 1e-6 * std::max(std::fabs(cg[n]), 1e-30)) plan.torsion_sines = true;
        t_coef[k][0] = (float)c0;
        for (int n = 0; n < NF; ++n) { t_coef[k][1 + n] = (float)cg[n]; t_coef[k][1 + NF + n] = (float)sg[n]; }
    }

    if (const char* env = std::getenv("MLP_FORCE_TORSION_SINES")) if (std::atoi(env)) plan.torsion_sines = true;  // test hook

    // terms incident to each atom
    std::vector<std::vector<int>> inc_b(N), inc_a(N), inc_t(N);
    for (int k = 0; k < nb; ++k) for (int r = 0; r < 2; ++r) inc_b[T.bonds[2 * k + r]].push_back(k);
    for (int k = 0; k < na; ++k) for (int r = 0; r < 3; ++r) inc_a[T.angles[3 * k + r]].push_back(k);
    for (int k = 0; k < nt; ++k) if (t_live[k]) for (int r = 0; r < 4; ++r) inc_t[T.torsions[4 * k + r]].push_back(k);

    std::vector<TileItems>& TI = plan.tiles;
    TI.clear();
    auto uniq = [](std::vector<int>& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
    auto key2 = [](int a, int b) { return std::make_pair(std::min(a, b), std::max(a, b)); };
    int max_cnt = 1;
    for (int a0 = 0; a0 < N; a0 += tile_atoms) {
        TileItems ti;
        ti.a0 = a0; ti.n_own = std::min(tile_atoms, N - a0);
        std::vector<int> tb, ta, tt;   // the tile's terms: every term that touches an owned atom
        for (int a = a0; a < a0 + ti.n_own; ++a) {
            tb.insert(tb.end(), inc_b[a].begin(), inc_b[a].end());
            ta.insert(ta.end(), inc_a[a].begin(), inc_a[a].end());
            tt.insert(tt.end(), inc_t[a].begin(), inc_t[a].end());
        }
        uniq(tb); uniq(ta); uniq(tt);

        // ---- angles -> torsions that contain them as (b1,b2,b3) or (b2,b3,b4): maximum bipartite matching (Kuhn)
        std::map<std::array<int, 3>, int> angle_at;   // (min end, centre, max end) -> index into ta
        for (size_t q = 0; q < ta.size(); ++q) {
            const int32_t* a = &T.angles[3 * ta[q]];
            angle_at[{std::min(a[0], a[2]), a[1], std::max(a[0], a[2])}] = (int)q;
        }
        std::vector<std::vector<int>> cand_a(ta.size());          // angle -> candidate torsions (indices into tt)
        for (size_t q = 0; q < tt.size(); ++q) {
            const int32_t* b = &T.torsions[4 * tt[q]];
            for (int side = 0; side < 2; ++side) {
                const int e0 = b[side], c = b[side + 1], e1 = b[side + 2];
                auto it = angle_at.find({std::min(e0, e1), c, std::max(e0, e1)});
                if (it != angle_at.end()) cand_a[it->second].push_back((int)q);
            }
        }
        std::vector<int> tors_angle(tt.size(), -1), angle_tors(ta.size(), -1);
        {
            std::vector<char> seen;
            std::function<bool(int)> grow = [&](int qa) -> bool {
                for (int qt : cand_a[qa]) {
                    if (seen[qt]) continue;
                    seen[qt] = 1;
                    if (tors_angle[qt] < 0 || grow(tors_angle[qt])) { tors_angle[qt] = qa; angle_tors[qa] = qt; return true; }
                }
                return false;
            };
            for (size_t qa = 0; qa < ta.size(); ++qa) { seen.assign(tt.size(), 0); grow((int)qa); }
        }
        // ---- items: torsions (oriented so that the carried angle is (i,j,k)), then the unmatched angles
        auto& items = ti.items;
        for (size_t q = 0; q < tt.size(); ++q) {
            const int32_t* b = &T.torsions[4 * tt[q]];
            Item it{{b[0], b[1], b[2], b[3]}, tt[q], -1, -1, false};
            if (tors_angle[q] >= 0) {
                it.angle = ta[tors_angle[q]];
                if (T.angles[3 * it.angle + 1] != b[1]) { it.a[0] = b[3]; it.a[1] = b[2]; it.a[2] = b[1]; it.a[3] = b[0]; }  // centre is b3: reverse
            }
            items.push_back(it);
        }
        for (size_t q = 0; q < ta.size(); ++q)
            if (angle_tors[q] < 0) {
                const int32_t* a = &T.angles[3 * ta[q]];
                items.push_back(Item{{a[0], a[1], a[2], a[0]}, -1, ta[q], -1, false});   // l: dummy (= i)
            }
        // ---- bonds -> items that contain them as (i,j) or (j,k) ((k,l) too for a torsion without angle: it can be reversed)
        std::map<std::pair<int, int>, int> bond_at;
        for (size_t q = 0; q < tb.size(); ++q) bond_at[key2(T.bonds[2 * tb[q]], T.bonds[2 * tb[q] + 1])] = (int)q;
        std::vector<std::vector<std::pair<int, int>>> cand_b(tb.size());   // bond -> (item, position 0: ij, 1: jk, 2: kl)
        for (size_t q = 0; q < items.size(); ++q) {
            const Item& it = items[q];
            const int npos = it.tors >= 0 ? (it.angle >= 0 ? 2 : 3) : 2;
            for (int pos = 0; pos < npos; ++pos) {
                auto f = bond_at.find(key2(it.a[pos], it.a[pos + 1]));
                if (f != bond_at.end()) cand_b[f->second].push_back({(int)q, pos});
            }
        }
        std::vector<int> item_bond(items.size(), -1), item_pos(items.size(), 0), bond_item(tb.size(), -1);
        {
            std::vector<char> seen;
            std::function<bool(int)> grow = [&](int qb) -> bool {
                for (auto& c : cand_b[qb]) {
                    if (seen[c.first]) continue;
                    seen[c.first] = 1;
                    if (item_bond[c.first] < 0 || grow(item_bond[c.first])) { item_bond[c.first] = qb; item_pos[c.first] = c.second; bond_item[qb] = c.first; return true; }
                }
                return false;
            };
            for (size_t qb = 0; qb < tb.size(); ++qb) { seen.assign(items.size(), 0); grow((int)qb); }
        }
        for (size_t q = 0; q < items.size(); ++q)
            if (item_bond[q] >= 0) {
                Item& it = items[q];
                it.bond = tb[item_bond[q]];
                if (item_pos[q] == 2) { std::swap(it.a[0], it.a[3]); std::swap(it.a[1], it.a[2]); it.bond_jk = false; }   // (k,l) -> reversed (i,j)
                else it.bond_jk = item_pos[q] == 1;
            }
        for (size_t q = 0; q < tb.size(); ++q)
            if (bond_item[q] < 0) {
                const int32_t* b = &T.bonds[2 * tb[q]];
                items.push_back(Item{{b[0], b[1], b[0], b[0]}, -1, -1, tb[q], false});   // k, l: dummies
            }
        // warps should be uniform in what their items carry
        auto kind = [](const Item& it) { return (it.tors >= 0 ? 0 : 4) + (it.angle >= 0 ? 0 : 2) + (it.bond >= 0 ? 0 : 1); };
        std::stable_sort(items.begin(), items.end(), [&](const Item& x, const Item& y) { return kind(x) < kind(y); });
        if (((items.size() + 31) & ~(size_t)31) > (size_t)RI * BT) return 1;

        // halo atoms and per-atom contribution counts (one slot per item that touches the atom with a real role)
        for (const Item& it : items) for (int r = 0; r < 4; ++r) if (it.a[r] < a0 || it.a[r] >= a0 + ti.n_own) ti.halo.push_back(it.a[r]);
        uniq(ti.halo);
        const int n_at = ti.n_own + (int)ti.halo.size();
        if (3 * n_at > PF * BT || n_at >= 0xffff) return 1;
        ti.cnt.assign(ti.n_own, 0);
        for (const Item& it : items) {
            const int roles = it.tors >= 0 ? 4 : it.angle >= 0 ? 3 : 2;
            for (int r = 0; r < roles; ++r) if (it.a[r] >= a0 && it.a[r] < a0 + ti.n_own) max_cnt = std::max(max_cnt, ++ti.cnt[it.a[r] - a0]);
        }
        TI.push_back(std::move(ti));
    }
    max_cnt = (max_cnt + 3) & ~3;   // the gather reads whole groups of 4 slots
    if ((size_t)max_cnt * SLOT_STRIDE + 1 >= 0xffff) return 1;
    plan.max_cnt = max_cnt;
    return MLP_OK;
}

int pack_bonded(mlp_energy* E, const mlp_topology& T, int tile_atoms) {
    BondedPlan plan;
    if (int rc = plan_bonded(T, tile_atoms, plan)) return rc;
    E->torsion_sines = plan.torsion_sines;
    const auto& t_coef = plan.t_coef;
    std::vector<TileItems>& TI = plan.tiles;
    const int max_cnt = plan.max_cnt;
    const int stride = SLOT_STRIDE;  // the kernel's compile-time slot stride
    const uint32_t dump = (uint32_t)(max_cnt * stride);

    std::vector<BTile> tiles;
    std::vector<int> halo_all, cnt_all;
    std::vector<ItemRec> items_all;
    std::vector<float4> sg_all;
    int max_atoms = 1;
    size_t n_items_real = 0;
    for (TileItems& ti : TI) {
        const int a0 = ti.a0, n_own = ti.n_own;
        const int n_at = n_own + (int)ti.halo.size();
        BTile bt{};
        bt.a0 = a0; bt.n_own = n_own;
        auto own = [&](int a) { return a >= a0 && a < a0 + n_own; };
        auto local = [&](int a) -> uint32_t {
            if (own(a)) return (uint32_t)(a - a0);
            return (uint32_t)(n_own + (std::lower_bound(ti.halo.begin(), ti.halo.end(), a) - ti.halo.begin()));
        };
        // slot of (own atom la, its k-th contribution) = k*stride + la; roles on halo atoms and dummy roles go to the dump slot
        std::vector<int> next(n_own, 0);
        bt.item_off = (int)items_all.size();
        bt.cnt_off = (int)cnt_all.size();
        for (const Item& it : ti.items) {
            ItemRec r{};
            const int roles = it.tors >= 0 ? 4 : it.angle >= 0 ? 3 : 2;
            uint32_t la[4], sl[4];
            for (int q = 0; q < 4; ++q) {
                la[q] = local(it.a[q]);
                sl[q] = (q < roles && own(it.a[q])) ? (uint32_t)(next[it.a[q] - a0]++ * stride + (it.a[q] - a0)) : dump;
            }
            r.ij = la[0] | (la[1] << 16); r.kl = la[2] | (la[3] << 16);
            r.s_ij = sl[0] | (sl[1] << 16); r.s_kl = sl[2] | (sl[3] << 16);
            float4 sg = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it.tors >= 0) {
                r.flags |= IT_T | (own(T.torsions[4 * it.tors]) ? IT_OWN_T : 0u);
                r.c0 = t_coef[it.tors][0];
                for (int n = 0; n < NF; ++n) r.cg[n] = t_coef[it.tors][1 + n];
                sg = make_float4(t_coef[it.tors][1 + NF], t_coef[it.tors][2 + NF], t_coef[it.tors][3 + NF], t_coef[it.tors][4 + NF]);
            }
            if (it.angle >= 0) {
                r.flags |= IT_A | (own(T.angles[3 * it.angle]) ? IT_OWN_A : 0u);
                r.th0 = (float)T.angle_para[2 * it.angle]; r.ka = (float)T.angle_para[2 * it.angle + 1];
            }
            if (it.bond >= 0) {
                r.flags |= IT_B | (it.bond_jk ? IT_B_JK : 0u) | (own(T.bonds[2 * it.bond]) ? IT_OWN_B : 0u);
                r.r0 = (float)T.bond_para[2 * it.bond]; r.kb = (float)T.bond_para[2 * it.bond + 1];
            }
            items_all.push_back(r);
            sg_all.push_back(sg);
            ++n_items_real;
        }
        while ((items_all.size() - bt.item_off) % 32) {      // inert items: no term, four stores to the dump slot
            ItemRec r{};
            r.s_ij = dump | (dump << 16); r.s_kl = dump | (dump << 16);
            items_all.push_back(r);
            sg_all.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
        }
        bt.n_item = (int)items_all.size() - bt.item_off;
        for (int la = 0; la < n_own; ++la) {
            if (next[la] != ti.cnt[la]) { mlp_set_error("pack_bonded: internal error, slot count mismatch"); return MLP_E_ARG; }
            cnt_all.push_back(next[la]);
        }
        // self-check of the packed records (compute-sanitizer is not available on the target pool): every local
        // atom index addresses a staged atom, every slot id addresses an allocated slot or the dump slot
        for (size_t q = bt.item_off; q < items_all.size(); ++q) {
            const ItemRec& r = items_all[q];
            const uint32_t at[4] = {r.ij & 0xffffu, r.ij >> 16, r.kl & 0xffffu, r.kl >> 16};
            const uint32_t sl[4] = {r.s_ij & 0xffffu, r.s_ij >> 16, r.s_kl & 0xffffu, r.s_kl >> 16};
            for (int c = 0; c < 4; ++c)
                if (at[c] >= (uint32_t)n_at || sl[c] > dump) { mlp_set_error("pack_bonded: internal error, record index out of range"); return MLP_E_ARG; }
        }
        bt.halo_off = (int)halo_all.size(); bt.n_halo = (int)ti.halo.size();
        halo_all.insert(halo_all.end(), ti.halo.begin(), ti.halo.end());
        max_atoms = std::max(max_atoms, n_at);
        tiles.push_back(bt);
    }
    const size_t smem = (BONDED_DB ? 2 : 1) * ((size_t)max_atoms + (size_t)max_cnt * stride + 1) * sizeof(float4);   // both halves of the double buffer
    if (smem > 200 * 1024) { mlp_set_error("pack_bonded: too many gradient contributions per atom for the shared-memory slot table"); return MLP_E_ARG; }
    E->n_tiles = (int)tiles.size(); E->tile_atoms = tile_atoms; E->max_cnt = max_cnt;
    E->max_atoms = max_atoms; E->bonded_smem = smem; E->n_items = (long)n_items_real;
    int rc;
    if ((rc = upload(E, &E->d_tiles, tiles))) return rc;
    if ((rc = upload(E, &E->d_halo, halo_all))) return rc;
    if ((rc = upload(E, &E->d_items, items_all))) return rc;
    if (!E->torsion_sines) sg_all.assign(1, make_float4(0.f, 0.f, 0.f, 0.f));
    if ((rc = upload(E, &E->d_sg, sg_all))) return rc;
    if ((rc = upload(E, &E->d_cnt, cnt_all))) return rc;
    return MLP_OK;
}

// Lennard-Jones classes of a topology: atoms with the same (R, eps) in order of first appearance, and the squared pair
// length w^2(a, b) = ((12 eps_ab)^(1/12) R_ab)^2 with R_ab = (R_a + R_b)/2, eps_ab = sqrt(eps_a eps_b) (utils.py:814-815),
// evaluated in fp64 and rounded once.  (w/d)^12 / 12 is the reference's vdw_A / d^12.
struct LjClasses {
    std::vector<int> of;          // [N] class of each atom
    std::vector<double> R, eps;   // per class
    int n() const { return (int)R.size(); }
    float w2(int a, int b) const {
        const double w = std::pow(12.0 * std::sqrt(eps[a] * eps[b]), 1.0 / 12.0) * 0.5 * (R[a] + R[b]);
        return (float)(w * w);
    }
};
LjClasses lj_classes(const mlp_topology& T) {
    LjClasses L;
    std::map<std::pair<float, float>, int> cls;
    L.of.resize(T.n_atoms);
    for (int i = 0; i < T.n_atoms; ++i) {
        const auto key = std::make_pair(T.vdw_R[i], T.vdw_e[i]);
        auto it = cls.find(key);
        if (it == cls.end()) {
            it = cls.emplace(key, (int)cls.size()).first;
            L.R.push_back(key.first); L.eps.push_back(key.second);
        }
        L.of[i] = it->second;
    }
    return L;
}

int pack_nonbonded(mlp_energy* E, const mlp_topology& T) {
    const int N = T.n_atoms;
    const int nt = (N + NB_T - 1) / NB_T;
    E->n_nbt = nt; E->Npad = nt * NB_T;
    std::vector<float> q(E->Npad, 0.f);
    std::vector<float2> par(E->Npad, make_float2(1.f, 0.f)), parw(E->Npad, make_float2(0.f, 0.f));
    for (int i = 0; i < N; ++i) {
        q[i] = (float)T.charges[i];
        // R_ij = |R_i+R_j|/2, eps_ij = sqrt(eps_i eps_j) (utils.py:814-815); 12 folded into eps
        par[i] = make_float2(0.5f * T.vdw_R[i], (float)std::sqrt(12.0 * (double)T.vdw_e[i]));
        const double bw = std::pow(12.0 * (double)T.vdw_e[i], 1.0 / 24.0);   // 0 for eps = 0
        parw[i] = make_float2((float)(bw * 0.5 * (double)T.vdw_R[i]), (float)bw);
    }
    std::vector<uint32_t> lut_lrow;
    std::vector<uint2> lut_lpc;
    // Lennard-Jones classes: atoms with the same (R, eps).  Up to NB_LUT_MAXC classes (plus a zero class for the
    // padding atoms) the repulsive kernel reads w_ij^2 = ((12 eps_ij)^(1/12) R_ij)^2 from a table in shared memory.
    {
        const LjClasses LJ = lj_classes(T);
        const int C = LJ.n() + 1;
        std::vector<int> c_of(E->Npad, C - 1);
        for (int i = 0; i < N; ++i) c_of[i] = LJ.of[i];
        E->lut_classes = C;
        E->lut_on = false;
        if (C <= NB_LUT_MAXC + 1) {
            auto w2 = [&](int a, int b) { return (a == C - 1 || b == C - 1) ? 0.f : LJ.w2(a, b); };
            std::vector<float2> lut((size_t)C * C * C);
            for (int a = 0; a < C; ++a)
                for (int b0 = 0; b0 < C; ++b0)
                    for (int b1 = 0; b1 < C; ++b1) lut[((size_t)a * C + b0) * C + b1] = make_float2(w2(a, b0), w2(a, b1));
            std::vector<uint32_t> lrow(E->Npad);
            std::vector<uint2> lpc(E->Npad / 4);
            for (int i = 0; i < E->Npad; ++i) lrow[i] = (uint32_t)c_of[i] * C * C * (uint32_t)sizeof(float2);
            for (int g = 0; g < E->Npad / 4; ++g)
                lpc[g] = make_uint2((uint32_t)(c_of[4 * g] * C + c_of[4 * g + 1]) * (uint32_t)sizeof(float2),
                                    (uint32_t)(c_of[4 * g + 2] * C + c_of[4 * g + 3]) * (uint32_t)sizeof(float2));
            E->lut_n = (int)lut.size();
            int rc;
            if ((rc = upload(E, &E->d_lut, lut))) return rc;
            if ((rc = upload(E, &E->d_lrow, lrow))) return rc;
            if ((rc = upload(E, &E->d_lpc, lpc))) return rc;
            lut_lrow = lrow; lut_lpc = lpc;
            E->lut_on = true;
        }
    }
    // exclusions per tile pair
    std::map<std::pair<int, int>, std::vector<std::pair<int, int>>> ex;
    for (size_t k = 0; k < T.excl.size() / 2; ++k) {
        int i = T.excl[2 * k], j = T.excl[2 * k + 1];
        ex[{i / NB_T, j / NB_T}].push_back({i, j});
    }
    for (int t = 0; t < nt; ++t) ex[{t, t}];  // diagonal tiles are always masked (j > i)
    std::vector<int4> band, clean, diag;
    std::vector<uint32_t> masks, stepbits;
    std::map<std::pair<int, int>, int> mask_of;   // tile pair -> mask index
    for (auto& kv : ex) {
        const int I = kv.first.first, J = kv.first.second;
        // uint16 [g][lane]: bit 4a+b <-> pair (i = 4*lane+a, j = 4g+b) allowed
        std::vector<uint16_t> m16((size_t)32 * 32, 0xffffu);
        auto clear = [&](int li, int lj) { m16[(size_t)(lj >> 2) * 32 + (li >> 2)] &= (uint16_t)~(1u << (4 * (li & 3) + (lj & 3))); };
        if (I == J) {
            // keep each unordered pair in exactly one of its two blocks: the one reached in rotation steps 0..16
            for (int li = 0; li < NB_T; ++li)
                for (int lj = 0; lj < NB_T; ++lj) {
                    const int gi = li >> 2, gj = lj >> 2, dist = (gj - gi) & 31;
                    const bool keep = (dist >= 1 && dist <= 15) || (dist == 16 && gi < 16) || (dist == 0 && li < lj);
                    if (!keep) clear(li, lj);
                }
            for (auto& pr : kv.second) {  // an excluded pair may sit in either orientation
                clear(pr.first - I * NB_T, pr.second - J * NB_T);
                clear(pr.second - J * NB_T, pr.first - I * NB_T);
            }
        } else {
            for (auto& pr : kv.second) clear(pr.first - I * NB_T, pr.second - J * NB_T);
        }
        const int mi = (int)(masks.size() / (NB_T * 4));
        mask_of[{I, J}] = mi;
        std::vector<uint32_t> m((size_t)NB_T * 4);
        std::memcpy(m.data(), m16.data(), m.size() * sizeof(uint32_t));
        masks.insert(masks.end(), m.begin(), m.end());
        uint32_t sb = 0;
        for (int g = 0; g < 32; ++g)
            for (int ln = 0; ln < 32; ++ln)
                if (m16[(size_t)g * 32 + ln] != 0xffffu) sb |= 1u << ((g - ln) & 31);   // lane ln meets group g at step (g - ln) mod 32
        stepbits.push_back(sb);
        if (I == J) {
            // diagonal tiles run rotation steps 0..16; they are the last (shortest) items of the launch and are cut
            // into NB_DIAG_SPLIT step ranges, each with its own partial block, to shorten the tail of the last wave
            for (int k = 0; k < NB_DIAG_SPLIT; ++k) {
                const int r0 = 17 * k / NB_DIAG_SPLIT, r1 = 17 * (k + 1) / NB_DIAG_SPLIT;
                diag.push_back(make_int4(I, J, mi, r0 | (r1 << 8)));
            }
        } else band.push_back(make_int4(I, J, mi, 32 << 8));
    }
    E->n_masked = (int)(band.size() + diag.size());
    for (int I = 0; I < nt; ++I)
        for (int J = I; J < nt; ++J)
            if (!ex.count({I, J})) clean.push_back(make_int4(I, J, -1, 32 << 8));
    std::vector<int4> pairs(band);
    pairs.insert(pairs.end(), clean.begin(), clean.end());
    pairs.insert(pairs.end(), diag.begin(), diag.end());
    E->n_pairs = (int)pairs.size();
    // per tile: the partial-force blocks to sum (pair index order, i-side before j-side)
    std::vector<std::vector<int>> lists(nt);
    for (int p = 0; p < E->n_pairs; ++p) {
        lists[pairs[p].x].push_back(2 * p);
        lists[pairs[p].y].push_back(2 * p + 1);
    }
    std::vector<int> red_off(nt + 1, 0), red_list;
    for (int t = 0; t < nt; ++t) {
        red_list.insert(red_list.end(), lists[t].begin(), lists[t].end());
        red_off[t + 1] = (int)red_list.size();
    }
    // ---- row bands (large systems): at most NB_BLK_MEM_MB of partial forces per conformation and launch ----
    E->bands.clear(); E->band_max = 0;
    E->band_bytes = (size_t)NB_BLK_MEM_MB << 20;
    if (const char* env = std::getenv("MLP_NB_BAND_KB")) E->band_bytes = (size_t)std::max(3, std::atoi(env)) << 10;   // test hook: bands on small systems
    if ((size_t)E->n_pairs * 6 * NB_T * sizeof(float) > E->band_bytes) {
        const size_t cap = E->band_bytes / (6 * NB_T * sizeof(float));
        std::vector<int4> pairs_b;
        int I0 = 0;
        while (I0 < nt) {
            size_t cnt = 0;
            int I1 = I0;
            while (I1 < nt && (I1 == I0 || cnt + (size_t)(nt - I1 - 1 + NB_DIAG_SPLIT) <= cap)) { cnt += (size_t)(nt - I1 - 1 + NB_DIAG_SPLIT); ++I1; }
            mlp_energy::PairBand bd;
            bd.begin = (int)pairs_b.size();
            for (const std::vector<int4>* src : {&band, &clean, &diag})
                for (const int4& q4 : *src) if (q4.x >= I0 && q4.x < I1) pairs_b.push_back(q4);
            bd.count = (int)pairs_b.size() - bd.begin;
            std::vector<std::vector<int>> bl(nt);
            for (int p = 0; p < bd.count; ++p) {
                bl[pairs_b[bd.begin + p].x].push_back(2 * p);
                bl[pairs_b[bd.begin + p].y].push_back(2 * p + 1);
            }
            std::vector<int> boff(nt + 1, 0), blist;
            for (int t = 0; t < nt; ++t) { blist.insert(blist.end(), bl[t].begin(), bl[t].end()); boff[t + 1] = (int)blist.size(); }
            if (blist.empty()) blist.push_back(0);
            int rcb;
            if ((rcb = upload(E, &bd.d_red_off, boff))) return rcb;
            if ((rcb = upload(E, &bd.d_red_list, blist))) return rcb;
            E->band_max = std::max(E->band_max, bd.count);
            E->bands.push_back(bd);
            I0 = I1;
        }
        int rcb;
        if ((rcb = upload(E, &E->d_pairs_b, pairs_b))) return rcb;
    }
    // ---- block-item schedule (nb_block_kernel) ----
    std::vector<float> tpar((size_t)nt * 3 * NB_T), tparw((size_t)nt * 3 * NB_T);
    for (int n = 0; n < E->Npad; ++n) {
        const size_t o = (size_t)(n / NB_T) * 3 * NB_T + n % NB_T;
        tpar[o] = q[n]; tpar[o + NB_T] = par[n].x; tpar[o + 2 * NB_T] = par[n].y;
        tparw[o] = q[n]; tparw[o + NB_T] = parw[n].x; tparw[o + 2 * NB_T] = parw[n].y;
    }
    if (E->lut_on) {     // per tile: 128 charges, 128 table row offsets, 32 x 2 pair offsets
        constexpr int TPW = NB_T + NB_T + NB_T / 2;
        std::vector<float> tparl((size_t)nt * TPW);
        for (int t = 0; t < nt; ++t) {
            float* o = tparl.data() + (size_t)t * TPW;
            for (int k = 0; k < NB_T; ++k) { o[k] = q[t * NB_T + k]; std::memcpy(o + NB_T + k, &lut_lrow[t * NB_T + k], 4); }
            std::memcpy(o + 2 * NB_T, &lut_lpc[(size_t)t * (NB_T / 4)], (NB_T / 4) * sizeof(uint2));
        }
        int rcl;
        if ((rcl = upload(E, &E->d_tparl, tparl))) return rcl;
    }
    std::vector<int2> pair_meta((size_t)nt * (nt + 1) / 2, make_int2(-1, 0));
    auto pid = [&](int I, int J) { return (size_t)I * nt - (size_t)I * (I - 1) / 2 + (J - I); };
    for (auto& kv : mask_of) pair_meta[pid(kv.first.first, kv.first.second)] = make_int2(kv.second, (int)stepbits[kv.second]);
    E->blk_acc_smem = NB_BLK_C * 3 * NB_T * (int)sizeof(float);
    int rc;
    for (int lv = 0; lv < NB_BLK_LEVELS; ++lv) {
        const int RB = NB_BLK_ROWS[lv];
        std::vector<NbBlockItem> bitems;
        for (int I0 = 0; I0 < nt; I0 += RB) {
            const int R = std::min(RB, nt - I0);
            for (int I = I0; I < I0 + R; ++I)                          // the diagonal band of the block row: single rows
                bitems.push_back(NbBlockItem{I, 1, I, I0 + R - I, 0, 0, 0, 0});
            for (int J0 = I0 + R; J0 < nt; J0 += NB_BLK_C)            // rectangles strictly above it
                bitems.push_back(NbBlockItem{I0, R, J0, std::min(NB_BLK_C, nt - J0), 0, 0, 0, 0});
        }
        auto cost = [&](const NbBlockItem& b) {
            long c = 0;
            for (int r = 0; r < b.R; ++r) for (int k = 0; k < b.C; ++k) c += (b.I0 + r == b.J0 + k) ? 17 : 32;
            return c;
        };
        std::stable_sort(bitems.begin(), bitems.end(), [&](const NbBlockItem& a, const NbBlockItem& b) { return cost(a) > cost(b); });
        std::vector<std::vector<int>> blists(nt);
        int nblk = 0;
        for (NbBlockItem& b : bitems) {
            b.iblk = nblk; nblk += b.R;
            b.jblk = nblk; nblk += b.C;
            for (int r = 0; r < b.R; ++r) blists[b.I0 + r].push_back(b.iblk + r);
            for (int k = 0; k < b.C; ++k) blists[b.J0 + k].push_back(b.jblk + k);
        }
        std::vector<int> bred_off(nt + 1, 0), bred_list;
        for (int t = 0; t < nt; ++t) {
            bred_list.insert(bred_list.end(), blists[t].begin(), blists[t].end());
            bred_off[t + 1] = (int)bred_list.size();
        }
        auto& S = E->blk[lv];
        S.rows = RB; S.items = (int)bitems.size(); S.blocks = nblk;
        if ((rc = upload(E, &S.d_items, bitems))) return rc;
        if ((rc = upload(E, &S.d_red_off, bred_off))) return rc;
        if ((rc = upload(E, &S.d_red_list, bred_list))) return rc;
    }
    if ((rc = upload(E, &E->d_pair_meta, pair_meta))) return rc;
    if ((rc = upload(E, &E->d_tpar, tpar))) return rc;
    if ((rc = upload(E, &E->d_tparw, tparw))) return rc;
    if ((rc = upload(E, &E->d_pairs, pairs))) return rc;
    if ((rc = upload(E, &E->d_masks, masks))) return rc;
    if ((rc = upload(E, &E->d_stepbits, stepbits))) return rc;
    if ((rc = upload(E, &E->d_parw, parw))) return rc;
    if ((rc = upload(E, &E->d_q, q))) return rc;
    if ((rc = upload(E, &E->d_par, par))) return rc;
    if ((rc = upload(E, &E->d_red_off, red_off))) return rc;
    if ((rc = upload(E, &E->d_red_list, red_list))) return rc;
    return MLP_OK;
}

struct Workspace { size_t e_bonded, e_nb, fpart, total; };

// Schedule of the pair term.  One warp per tile pair (nb_kernel2) is the faster kernel at every size measured (MurD x 64,
// 10 715 atoms x 32 / 128, 51 432 atoms x 8: 10-12 % ahead of the block items since the table form and the lean item
// prologue), but it needs 3 KB of partial forces per tile pair and conformation.  Two bounds keep that scratch O(N):
//   * a batch is evaluated in chunks of conformations whose partial forces fit NB_WS_CAP_MB (MLP_NB_WS_MB; each chunk is
//     one pair launch + one reduction, and a chunk of one 10 k-atom conformation is already 1.5 waves of warps);
//   * when ONE conformation needs more than NB_BLK_MEM_MB (systems beyond ~26 k atoms) the tile-pair list is walked in
//     row bands of at most that size (pair launch + reduction per band, the later bands adding to the gradient), or,
//     with MLP_NB_BANDS=0, as block items: rows in registers, column forces accumulated in shared memory, about a
//     fifth of the scratch, 2 % slower.
// A pure function of the handle and the batch size: mlp_energy_workspace_bytes and mlp_energy_eval agree on it.
#ifndef NB_BLK_MIN_ITEMS
#define NB_BLK_MIN_ITEMS 8
#endif
#ifndef NB_BLK_MEM_MB
#define NB_BLK_MEM_MB 64
#endif
#ifndef NB_WS_CAP_MB
#define NB_WS_CAP_MB 384
#endif
int block_level(const mlp_energy* e, int64_t B) {   // -1: one warp per tile pair (nb_kernel2)
    if (!e->packed || e->blk_mode == 0) return -1;
    if (e->blk_mode > 0) {
        for (int lv = 0; lv < NB_BLK_LEVELS; ++lv) if (NB_BLK_ROWS[lv] == e->blk_mode) return lv;
        return NB_BLK_LEVELS - 1;
    }
    if ((size_t)e->n_pairs * 6 * NB_T * sizeof(float) <= e->band_bytes) return -1;
    if (e->band_mode != 0 && !e->bands.empty()) return -1;       // the tile-pair kernel, band by band
    // memory-bound regime: the biggest blocks that still leave every resident warp slot NB_BLK_MIN_ITEMS items, at least 4 rows
    const bool tab = e->lut_on && e->nb_mode != MLP_NB_FULL;
    const double per_slot = (double)e->n_nbt * (e->n_nbt + 1) / 2 * (double)B / ((double)e->sm_count * 4 * (tab ? NB2_MINB_LUT : NB2_MINB));
    int best = 1;
    for (int lv = 1; lv < NB_BLK_LEVELS; ++lv)
        if (per_slot >= (double)NB_BLK_MIN_ITEMS * NB_BLK_ROWS[lv] * NB_BLK_C) best = lv;
    return best;
}
// partial-force blocks (3 x 128 floats) and energy partials per conformation under the schedule in use
bool banded(const mlp_energy* e, int64_t B) { return e->packed && e->band_mode != 0 && !e->bands.empty() && block_level(e, B) < 0; }
int nb_blocks(const mlp_energy* e, int64_t B) {
    const int lv = block_level(e, B);
    return lv >= 0 ? e->blk[lv].blocks : banded(e, B) ? 2 * e->band_max : 2 * e->n_pairs;
}
int nb_eparts(const mlp_energy* e, int64_t B) { const int lv = block_level(e, B); return lv >= 0 ? e->blk[lv].items : e->n_pairs; }

// conformations per pair launch: the partial forces of a chunk fit the scratch cap (at least one conformation)
int64_t nb_chunk(const mlp_energy* e, int64_t B) {
    const size_t per = (size_t)nb_blocks(e, B) * 3 * NB_T * sizeof(float);
    const int64_t c = (int64_t)(((size_t)e->ws_cap_mb << 20) / std::max<size_t>(per, 1));
    return std::max<int64_t>(1, std::min<int64_t>(B, c));
}

Workspace layout(const mlp_energy* e, int64_t B, uint32_t term_mask, bool grad) {
    Workspace w{};
    size_t off = 0;
    w.e_bonded = off; off += align256((size_t)B * std::max(e->n_tiles, 1) * 3 * sizeof(float));
    w.e_nb = off;
    if (term_mask & MLP_TERM_NB) off += align256((size_t)B * nb_eparts(e, B) * sizeof(double));
    w.fpart = off;
    if ((term_mask & MLP_TERM_NB) && grad) off += align256((size_t)nb_chunk(e, B) * nb_blocks(e, B) * 3 * NB_T * sizeof(float));
    w.total = off + 256;
    return w;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int mlp_bonded_plan(const mlp_topology* t, int tile_atoms, int64_t* n_items, int32_t* items, int64_t capacity, int64_t stats[4]) {
    if (!t || !n_items || tile_atoms < 1 || tile_atoms > BTA) { mlp_set_error("mlp_bonded_plan: bad argument"); return MLP_E_ARG; }
    BondedPlan plan;
    int rc = plan_bonded(*t, tile_atoms, plan);
    if (rc == 1) { mlp_set_error("mlp_bonded_plan: topology too dense for this tile size"); return MLP_E_ARG; }
    if (rc) return rc;
    int64_t n = 0;
    for (size_t ti = 0; ti < plan.tiles.size(); ++ti)
        for (const Item& it : plan.tiles[ti].items) {
            if (items && n < capacity) {
                int32_t* row = items + 9 * n;
                row[0] = (int32_t)ti;
                for (int q = 0; q < 4; ++q) row[1 + q] = it.a[q];
                row[5] = it.tors; row[6] = it.angle; row[7] = it.bond; row[8] = it.bond_jk ? 1 : 0;
            }
            ++n;
        }
    *n_items = n;
    if (stats) {
        stats[0] = (int64_t)plan.tiles.size(); stats[1] = plan.max_cnt; stats[2] = plan.torsion_sines ? 1 : 0;
        stats[3] = 0;
        for (char c : plan.t_live) stats[3] += c ? 1 : 0;
    }
    return MLP_OK;
}

int mlp_nb_pair_table(const mlp_topology* t, int32_t* atom_class, float* w2, int64_t capacity, int64_t* n_classes) {
    if (!t || !n_classes) { mlp_set_error("mlp_nb_pair_table: null argument"); return MLP_E_ARG; }
    const LjClasses L = lj_classes(*t);
    const int C = L.n();
    *n_classes = C;
    if (atom_class) for (int i = 0; i < t->n_atoms; ++i) atom_class[i] = L.of[i];
    if (w2) {
        if (capacity < (int64_t)C * C) { mlp_set_error("mlp_nb_pair_table: w2 holds fewer than n_classes^2 floats"); return MLP_E_ARG; }
        for (int a = 0; a < C; ++a) for (int b = 0; b < C; ++b) w2[(size_t)a * C + b] = L.w2(a, b);
    }
    return MLP_OK;
}

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
    int tile = BTA;
    if (const char* env = std::getenv("MLP_BONDED_TILE")) tile = std::max(32, std::min(BTA, std::atoi(env) & ~7));
    while (true) {
        rc = pack_bonded(E, *t, tile);
        if (rc == 1 && tile > 32) { tile -= 32; for (void* p : E->allocs) cudaFree(p); E->allocs.clear(); continue; }
        break;
    }
    if (rc == 1) { mlp_set_error("mlp_energy_create: topology too dense for the bonded tile"); rc = MLP_E_ARG; }
    if (rc == MLP_OK) rc = pack_nonbonded(E, *t);
    if (rc == MLP_OK) {
        if (const char* env = std::getenv("MLP_NO_OVERLAP")) E->overlap = std::atoi(env) == 0;
        if (const char* env = std::getenv("MLP_NB_FIRST")) E->nb_first = std::atoi(env) != 0;
        E->ws_cap_mb = NB_WS_CAP_MB;
        if (const char* env = std::getenv("MLP_NB_WS_MB")) E->ws_cap_mb = std::max(1, std::atoi(env));
        if (const char* env = std::getenv("MLP_NB_BANDS")) E->band_mode = std::atoi(env);
        if (const char* env = std::getenv("MLP_NB_BLOCKS")) E->blk_mode = std::max(0, std::atoi(env));   // 0: never, 2/4/8: always that block height
        if (const char* env = std::getenv("MLP_NB_PACKED")) E->packed = std::atoi(env) != 0;
        if (const char* env = std::getenv("MLP_NB_LUT")) E->lut_on = E->lut_on && std::atoi(env) != 0;
        if (cudaStreamCreateWithFlags(&E->side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&E->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&E->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            mlp_set_error("mlp_energy_create: could not create the side stream"); rc = MLP_E_CUDA;
        }
    }
    if (rc == MLP_OK) {
        // The opt-in dynamic shared-memory limit is a property of the kernel FUNCTION on a device, shared by every
        // handle: it is only ever raised (a later, smaller topology must not lower it under an earlier handle).
        static std::mutex mtx;
        static std::map<int, size_t> raised;
        std::lock_guard<std::mutex> lock(mtx);
        size_t& have = raised[device];
        if (E->bonded_smem > have) {
            const int want = (int)E->bonded_smem;
            cudaError_t err = cudaFuncSetAttribute(bonded_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            if (err == cudaSuccess) err = cudaFuncSetAttribute(bonded_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            if (err == cudaSuccess) err = cudaFuncSetAttribute(bonded_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            if (err == cudaSuccess) err = cudaFuncSetAttribute(bonded_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
            if (err != cudaSuccess) { mlp_set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err)); rc = MLP_E_CUDA; }
            else have = E->bonded_smem;
        }
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
    for (int k = 0; k < 4; ++k)
        for (auto& p : e->pending[k]) { cudaEventDestroy(p[0]); cudaEventDestroy(p[1]); }
    for (void* p : e->allocs) cudaFree(p);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->side) cudaStreamDestroy(e->side);
    cudaSetDevice(prev);
    delete e;
}

int mlp_energy_info(const mlp_energy* e, int64_t info[8]) {
    if (!e || !info) { mlp_set_error("mlp_energy_info: null argument"); return MLP_E_ARG; }
    info[0] = e->n_tiles; info[1] = e->tile_atoms; info[2] = e->max_cnt; info[3] = (int64_t)e->bonded_smem;
    info[4] = e->n_nbt; info[5] = e->n_pairs; info[6] = e->n_masked; info[7] = e->Npad;
    return MLP_OK;
}

int mlp_energy_nb_table_classes(const mlp_energy* e) { return (e && e->lut_on && e->nb_mode != MLP_NB_FULL) ? e->lut_classes : 0; }

int64_t mlp_energy_workspace_bytes(const mlp_energy* e, int64_t B, uint32_t term_mask, int want_grad) {
    if (!e || B < 0) return MLP_E_ARG;
    return (int64_t)layout(e, B, term_mask, want_grad != 0).total;
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
            e->prof_ms[k] += t;
            cudaEventDestroy(p[0]); cudaEventDestroy(p[1]);
        }
        for (char f : e->pending_first[k]) e->prof_n[k] += f ? 1 : 0;
        e->pending[k].clear(); e->pending_first[k].clear();
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
    if (B > 65535) { mlp_set_error("mlp_energy_eval: at most 65535 conformations per call (split the batch)"); return MLP_E_ARG; }
    if (!x || !workspace) { mlp_set_error("mlp_energy_eval: null x / workspace"); return MLP_E_ARG; }
    const Workspace W = layout(e, B, term_mask, grad != nullptr);
    if (workspace_bytes < (int64_t)W.total) { mlp_set_error("mlp_energy_eval: workspace too small"); return MLP_E_WORKSPACE; }
    if (reinterpret_cast<uintptr_t>(workspace) & 15) { mlp_set_error("mlp_energy_eval: workspace must be 16-byte aligned"); return MLP_E_ARG; }
    {   // every argument is checked before any device state (current device, events, side stream) is touched
        auto fits = [&](int64_t sc, int64_t sn) { return sc >= 0 && sn >= 0 && (int64_t)(e->N - 1) * sn + 2 * sc < ((int64_t)1 << 31); };
        if (!fits(sxc, sxn) || (grad && !fits(sgc, sgn))) {
            mlp_set_error("mlp_energy_eval: one conformation is addressed with 32-bit element offsets "
                          "(strides must be non-negative and a conformation must span less than 2^31 elements)");
            return MLP_E_ARG;
        }
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    struct DeviceGuard {     // restores the caller's current device on every return path
        int prev = -1, want;
        explicit DeviceGuard(int w) : want(w) {}
        cudaError_t enter() {
            cudaError_t r = cudaGetDevice(&prev);
            if (r == cudaSuccess && prev != want) r = cudaSetDevice(want);
            return r;
        }
        ~DeviceGuard() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
    } guard(e->device);
    CK(guard.enter());
    cudaGetLastError();   // a stale (non-sticky) error left in this host thread by an earlier, unrelated runtime call is not ours to report
    cudaError_t lerr = cudaSuccess;
#define LAUNCH_CHECK(name)                                                                                    \
    if ((lerr = cudaGetLastError()) != cudaSuccess) {                                                        \
        mlp_set_error(std::string("launch of " name ": ") + cudaGetErrorString(lerr));                      \
        return MLP_E_CUDA;                                                                                    \
    }
    char* ws = static_cast<char*>(workspace);
    ws += (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
    float* e_bonded = reinterpret_cast<float*>(ws + W.e_bonded);
    double* e_nb = reinterpret_cast<double*>(ws + W.e_nb);
    float* fpart = reinterpret_cast<float*>(ws + W.fpart);
    const int N = e->N;
    const bool accumulate = flags & MLP_EVAL_ACCUMULATE_GRAD;
    const bool bonded = term_mask & MLP_TERM_BONDED, nonbonded = term_mask & MLP_TERM_NB;
    int launches = 0;

    auto prof_begin = [&](int k, bool first = true) {
        if (!e->profile) return;
        std::array<cudaEvent_t, 2> p;
        cudaEventCreate(&p[0]); cudaEventCreate(&p[1]);
        cudaEventRecord(p[0], stream);
        e->pending[k].push_back(p);
        e->pending_first[k].push_back(first ? 1 : 0);
    };
    auto prof_end = [&](int k) { if (e->profile) cudaEventRecord(e->pending[k].back()[1], stream); };

    // With both parts requested the bonded kernel goes to the handle's side stream (fork after everything
    // already queued on `stream`, join before the first kernel that touches the gradient again): it is
    // latency-bound at small batches and fills the pair kernel's tail.  Per-kernel timing keeps one stream.
    const bool fork = bonded && nonbonded && e->overlap && !e->profile;
    bool energies_done = false;
    cudaStream_t bstream = stream;
    if (fork) {
        CK(cudaEventRecord(e->ev_fork, stream));
        CK(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
        bstream = e->side;
    }
    // nb_first: the pair kernel is launched BEFORE the bonded kernel, so that the (short, latency-bound) bonded CTAs
    // are handed out when the pair kernel's grid is exhausted - in its last, partially filled wave - instead of taking
    // slots at the start.
    const bool nb_first = fork && e->nb_first;
    auto launch_bonded = [&]() {
        prof_begin(1);
        int cpc = (int)std::max<int64_t>(1, std::min<int64_t>(BONDED_CPC_MAX, (B * e->n_tiles) / (int64_t)(e->sm_count * 8)));
        dim3 grid(e->n_tiles, (unsigned)((B + cpc - 1) / cpc));
#define BONDED_LAUNCH(G, S) bonded_kernel<G, S><<<grid, BT, e->bonded_smem, bstream>>>(e->d_tiles, e->d_halo, e->d_items, e->d_sg, e->d_cnt, \
            x, sxb, (int)sxc, (int)sxn, grad, sgb, (int)sgc, (int)sgn, weights, energies ? e_bonded : nullptr, (int)B, cpc, \
            term_mask & MLP_TERM_BONDED, grad ? (int)accumulate : 0, e->max_atoms, e->max_cnt)
        if (grad) { if (e->torsion_sines) BONDED_LAUNCH(true, true); else BONDED_LAUNCH(true, false); }
        else { if (e->torsion_sines) BONDED_LAUNCH(false, true); else BONDED_LAUNCH(false, false); }
#undef BONDED_LAUNCH
        prof_end(1);
        ++launches;
    };
    auto bonded_checked = [&]() -> int { launch_bonded(); LAUNCH_CHECK("bonded_kernel"); return MLP_OK; };
    if (bonded && !nb_first) {
        if (int rc = bonded_checked()) return rc;
        if (fork) CK(cudaEventRecord(e->ev_join, e->side));
    } else if (!bonded && grad && !accumulate && !nonbonded) {
        long n = (long)B * N * 3;
        zero_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(grad, sgb, sgc, sgn, N, (int)B);
        LAUNCH_CHECK("zero_grad_kernel");
        ++launches;
    }
    if (nonbonded) {
        const bool full = e->nb_mode == MLP_NB_FULL;
        const int lv = block_level(e, B);
        const bool blocks_on = lv >= 0;
        const mlp_energy::BlockSchedule& S = e->blk[blocks_on ? lv : 0];
        const bool tab = !full && e->lut_on;
        const int n_ep = nb_eparts(e, B), n_blk = nb_blocks(e, B);
        // with a gradient the batch goes in chunks of conformations whose partial forces fit the scratch (usually one chunk),
        // and a large system band by band through its tile-pair rows
        const int64_t chunk = grad ? nb_chunk(e, B) : B;
        const bool bands_on = grad && !blocks_on && banded(e, B);
        const int n_bands = bands_on ? (int)e->bands.size() : 1;
        for (int64_t b0 = 0; b0 < B; b0 += chunk) {
            const int Bc = (int)std::min<int64_t>(chunk, B - b0);
            const float* xc = x + b0 * sxb;
            const float* wc = weights ? weights + 4 * b0 : nullptr;
            double* ep = energies ? e_nb + b0 * n_ep : nullptr;
            for (int kb = 0; kb < n_bands; ++kb) {
                const int4* plist = bands_on ? e->d_pairs_b + e->bands[kb].begin : e->d_pairs;
                const int np = bands_on ? e->bands[kb].count : e->n_pairs;
                const int e_off = bands_on ? e->bands[kb].begin : 0;
                const bool first = b0 == 0 && kb == 0, last_band = kb == n_bands - 1;
                prof_begin(2, first);
                const unsigned blocks = (unsigned)(((long)Bc * np + NB_WARPS - 1) / NB_WARPS);
#define NB_LAUNCH(G, F) nb_kernel<G, F><<<blocks, NB_WARPS * 32, 0, stream>>>(xc, sxb, sxc, sxn, e->d_q, N, e->d_par, plist, e->d_masks, np, Bc, fpart, wc, ep)
#define NB_LAUNCH2(G, F, T) nb_kernel2<G, F, T><<<grid2, 32, T ? e->lut_n * sizeof(float2) : 0, stream>>>(xc, sxb, sxc, sxn, e->d_q, N, (NB_WFORM && !F) ? e->d_parw : e->d_par, plist, e->d_masks, e->d_stepbits, np, Bc, fpart, wc, ep, e->d_lrow, e->d_lpc, e->d_lut, e->lut_n, n_ep, e_off)
                const unsigned gy = (unsigned)std::min(np, 65535);
                const dim3 grid2((unsigned)Bc, gy, (unsigned)((np + gy - 1) / gy));
#define NB_LAUNCHB(G, F, T) nb_block_kernel<G, F, T><<<dim3((unsigned)Bc, (unsigned)std::min(S.items, 65535), (unsigned)((S.items + 65534) / 65535)), 32, e->blk_acc_smem + (T ? e->lut_n * sizeof(float2) : 0), stream>>>(xc, sxb, sxc, sxn, N, e->n_nbt, \
                    T ? e->d_tparl : (NB_WFORM && !F) ? e->d_tparw : e->d_tpar, S.d_items, e->d_pair_meta, e->d_masks, S.items, S.blocks, Bc, fpart, wc, ep, \
                    e->d_lut, e->lut_n, e->blk_acc_smem / (int)sizeof(float))
                if (blocks_on) {
                    if (grad) { if (full) NB_LAUNCHB(true, true, false); else if (tab) NB_LAUNCHB(true, false, true); else NB_LAUNCHB(true, false, false); }
                    else { if (full) NB_LAUNCHB(false, true, false); else if (tab) NB_LAUNCHB(false, false, true); else NB_LAUNCHB(false, false, false); }
                } else if (e->packed) {
                    if (grad) { if (full) NB_LAUNCH2(true, true, false); else if (tab) NB_LAUNCH2(true, false, true); else NB_LAUNCH2(true, false, false); }
                    else { if (full) NB_LAUNCH2(false, true, false); else if (tab) NB_LAUNCH2(false, false, true); else NB_LAUNCH2(false, false, false); }
                } else {
                    if (grad) { if (full) NB_LAUNCH(true, true); else NB_LAUNCH(true, false); }
                    else { if (full) NB_LAUNCH(false, true); else NB_LAUNCH(false, false); }
                }
#undef NB_LAUNCHB
#undef NB_LAUNCH2
#undef NB_LAUNCH
                LAUNCH_CHECK("the non-bonded pair kernel");
                prof_end(2);
                ++launches;
                if (first) {
                    if (nb_first) {
                        if (int rc = bonded_checked()) return rc;
                        CK(cudaEventRecord(e->ev_join, e->side));
                    }
                    if (fork) CK(cudaStreamWaitEvent(stream, e->ev_join, 0));   // the reduction adds to the bonded gradient (and sums its energy partials)
                }
                if (grad) {
                    // a band's reduction adds to what the bands before it stored; the last one also sums the conformation's
                    // energy partials (all pair launches of the chunk are behind it in stream order)
                    prof_begin(3, first);
                    nb_reduce_kernel<<<dim3(e->n_nbt, (unsigned)Bc), 32 * RED_PARTS, 0, stream>>>(fpart,
                        bands_on ? e->bands[kb].d_red_off : blocks_on ? S.d_red_off : e->d_red_off,
                        bands_on ? e->bands[kb].d_red_list : blocks_on ? S.d_red_list : e->d_red_list, n_ep, bands_on ? 2 * np : n_blk, N,
                        grad + b0 * sgb, sgb, sgc, sgn, wc, (accumulate || bonded || kb > 0) ? 1 : 0, (weights && !energies) ? 1 : 0,
                        e_bonded + b0 * e->n_tiles * 3, e->n_tiles, e_nb + b0 * n_ep, (energies && last_band) ? energies + 4 * b0 : nullptr, term_mask);
                    LAUNCH_CHECK("nb_reduce_kernel");
                    prof_end(3);
                    ++launches;
                    energies_done = energies != nullptr;
                }
            }
        }
    }
    if (energies && !energies_done) {
        finalize_kernel<<<(unsigned)B, 128, 0, stream>>>(e_bonded, e->n_tiles, e_nb, nb_eparts(e, B), energies, term_mask);
        LAUNCH_CHECK("finalize_kernel");
        ++launches;
    }
#undef LAUNCH_CHECK
    e->last_launches = launches;
    return launches;       // >= 0: success, the number of kernels queued (also mlp_energy_last_launches)
}

int mlp_rmsd_eval(const float* x, int64_t B, int64_t N, int64_t sxb, int64_t sxc, int64_t sxn, float scale, float shift,
                  const float* targets, int64_t K, float* out, uint32_t flags, void* stream_) {
    if (B == 0) return MLP_OK;
    const bool paired = flags & MLP_RMSD_PAIRED;
    if (!x || !targets || !out || B < 0 || N <= 0 || K <= 0 || (flags & ~(MLP_RMSD_ALIGN | MLP_RMSD_PAIRED)) ||
        (paired && K != B) || (!paired && K > 65535 * (int64_t)RMSD_KC)) {
        mlp_set_error("mlp_rmsd_eval: bad argument (paired mode needs K == B)");
        return MLP_E_ARG;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (flags & MLP_RMSD_ALIGN) {
        if (!paired && K > 65535) { mlp_set_error("mlp_rmsd_eval: at most 65535 targets with MLP_RMSD_ALIGN"); return MLP_E_ARG; }
        rmsd_aligned_kernel<<<dim3((unsigned)B, paired ? 1u : (unsigned)K), 256, 0, stream>>>(x, sxb, sxc, sxn, scale, shift, targets, (int)N, (int)K, (int)paired, out);
    } else {
        rmsd_kernel<<<dim3((unsigned)B, paired ? 1u : (unsigned)((K + RMSD_KC - 1) / RMSD_KC)), 256, 0, stream>>>(x, sxb, sxc, sxn, scale, shift, targets, (int)N, (int)K, (int)paired, out);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { mlp_set_error(std::string("rmsd kernel launch: ") + cudaGetErrorString(err)); return MLP_E_CUDA; }
    return MLP_OK;
}

int mlp_internal_coords_eval(const float* x, int64_t B, int64_t sxb, int64_t sxc, int64_t sxn, float scale,
                             const int32_t* pairs, int64_t n_pairs, const int32_t* triples, int64_t n_triples,
                             const int32_t* quads, int64_t n_quads, float* out, void* stream_) {
    const int64_t n_all = n_pairs + n_triples + n_quads;
    if (B == 0 || n_all == 0) return MLP_OK;
    if (!x || !out || B < 0 || B > 65535 || n_pairs < 0 || n_triples < 0 || n_quads < 0 || n_all > (1LL << 30) ||
        (n_pairs > 0 && !pairs) || (n_triples > 0 && !triples) || (n_quads > 0 && !quads)) {
        mlp_set_error("mlp_internal_coords_eval: bad argument (at most 65535 structures per call)");
        return MLP_E_ARG;
    }
    internal_coords_kernel<<<dim3((unsigned)((n_all + 127) / 128), (unsigned)B), 128, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, sxb, sxc, sxn, scale, reinterpret_cast<const int2*>(pairs), (int)n_pairs, triples, (int)n_triples,
        reinterpret_cast<const int4*>(quads), (int)n_quads, out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { mlp_set_error(std::string("internal-coordinates kernel launch: ") + cudaGetErrorString(err)); return MLP_E_CUDA; }
    return MLP_OK;
}

int64_t mlp_probe_fp32(float* scratch, int64_t scratch_floats, int packed, int iters, void* stream_) {
    const int64_t need = (int64_t)PROBE_BLOCKS * PROBE_THREADS;
    if (!scratch) return need;                                  // size query
    if (scratch_floats < need || iters <= 0) { mlp_set_error("mlp_probe_fp32: scratch too small / bad iteration count"); return MLP_E_ARG; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (packed) fp32_probe_kernel<1><<<PROBE_BLOCKS, PROBE_THREADS, 0, stream>>>(scratch, iters);
    else fp32_probe_kernel<0><<<PROBE_BLOCKS, PROBE_THREADS, 0, stream>>>(scratch, iters);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { mlp_set_error(std::string("probe kernel launch: ") + cudaGetErrorString(err)); return MLP_E_CUDA; }
    return 2LL * 16 * (int64_t)iters * need;                    // flops of the launch
}

int mlp_geometry_eval(const float* x, int64_t B, int64_t sxb, int64_t sxc, int64_t sxn, float scale,
                      const int32_t* pairs, const int32_t* pair_group, int64_t n_pairs, int64_t n_groups,
                      const int32_t* quads, int64_t n_quads, float* out, void* stream_) {
    if (B == 0) return MLP_OK;
    if (!x || !out || B < 0 || n_pairs < 0 || n_quads < 0 || n_groups < 0 || n_groups > GEOM_MAX_GROUPS ||
        (n_pairs > 0 && (!pairs || !pair_group)) || (n_quads > 0 && !quads)) {
        mlp_set_error("mlp_geometry_eval: bad argument");
        return MLP_E_ARG;
    }
    geom_stats_kernel<<<(unsigned)B, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, sxb, sxc, sxn, scale, reinterpret_cast<const int2*>(pairs), pair_group, (int)n_pairs, (int)n_groups,
        reinterpret_cast<const int4*>(quads), (int)n_quads, out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { mlp_set_error(std::string("geometry kernel launch: ") + cudaGetErrorString(err)); return MLP_E_CUDA; }
    return MLP_OK;
}

}  // extern "C"
