/* TEST INFRASTRUCTURE ONLY - fp64 CPU restatement of the reference energy (SURVEY.md 8a rows
 * E-1..E-4, NB-1, NB-2) evaluated on the explicit term lists.  Never on the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg load this library.
 *
 * Pinned against fp64 runs of the unmodified reference (tests/golden/energy_*.npz, produced by
 * oracle/make_golden.py) in tests/test_oracle_energy.py.
 *
 * Reference being restated (src/molearn/loss_functions/torch_protein_energy.py):
 *   bonds     :282-284   E = sum K (|r_i - r_j| - r0)^2
 *   angles    :286-293   cos = v1.v2/(|v1||v2|), theta = acos(clamp(cos, +-0.999999)),
 *                        E = sum K (theta - theta0)^2
 *   torsions  :295-304   b1=r_i-r_j, b2=r_j-r_k, b3=r_l-r_k, c32=b3 x b2, c12=b1 x b2,
 *                        phi = atan2(b2.(c32 x c12), |b2| (c12.c32)),
 *                        E = sum_p (V_p/f_p)(1 + cos(n_p phi - gamma_p))
 *   nonbonded :224-239   w(d,k) = elu(d-k)+k;  E = sum_{i<j, not excluded} A_ij/w(d,1.9)^12
 *                        [- B_ij/w(d,1.9)^6] + q_i q_j / w(d,0.4)
 *   with A_ij = eps_ij R_ij^12, B_ij = 2 eps_ij R_ij^6, R_ij = |R_i+R_j|/2,
 *   eps_ij = sqrt(eps_i eps_j) (torch_protein_energy_utils.py:813-830, energy.py:125-127) and
 *   the 1-2/1-3/1-4 pairs removed (pair semantics, SURVEY NB-3).
 *
 * Gradients come from forward-mode automatic differentiation of exactly those formulas (dual
 * numbers below), not from closed-form force expressions, so they are an independent check of
 * the CUDA kernels' analytic gradients and reproduce autograd's conventions (zero gradient
 * where the clamp is active).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ND 12
typedef struct { double v; double d[ND]; } dual;

static dual d_const(double v) { dual r; r.v = v; memset(r.d, 0, sizeof r.d); return r; }
static dual d_var(double v, int k) { dual r = d_const(v); r.d[k] = 1.0; return r; }
static dual d_add(dual a, dual b) { dual r; r.v = a.v + b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
static dual d_sub(dual a, dual b) { dual r; r.v = a.v - b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
static dual d_mul(dual a, dual b) { dual r; r.v = a.v * b.v; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
static dual d_div(dual a, dual b) { dual r; r.v = a.v / b.v; for (int i = 0; i < ND; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
static dual d_scale(dual a, double s) { dual r; r.v = a.v * s; for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * s; return r; }
static dual d_sqrt(dual a) { dual r; r.v = sqrt(a.v); for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] / (2.0 * r.v); return r; }
static dual d_cos(dual a) { dual r; r.v = cos(a.v); double s = -sin(a.v); for (int i = 0; i < ND; ++i) r.d[i] = s * a.d[i]; return r; }
static dual d_atan2(dual y, dual x) {
    /* atan2(+-0, +-0): with exactly coincident atoms both arguments are signed zeros whose signs depend on the
     * summation order of the dot products (torch's reductions and this loop can disagree: 0 vs +-pi, seen on a
     * collapsed 51k-atom decode).  The convention pinned here - and implemented by the kernels - is phi = 0, the
     * value torch returns for atan2(0, 0). */
    dual r; r.v = (y.v == 0.0 && x.v == 0.0) ? 0.0 : atan2(y.v, x.v);
    double q = x.v * x.v + y.v * y.v;
    for (int i = 0; i < ND; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) / q;
    return r;
}
/* acos(clamp(c, lo, hi)): torch.clamp's backward passes the gradient only strictly inside */
static dual d_acos_clamped(dual c, double lo, double hi) {
    dual r;
    if (c.v < lo || c.v > hi) { r = d_const(acos(c.v < lo ? lo : hi)); if (c.v != c.v) r.v = c.v; return r; }
    r.v = acos(c.v);
    double s = -1.0 / sqrt(1.0 - c.v * c.v);
    for (int i = 0; i < ND; ++i) r.d[i] = s * c.d[i];
    return r;
}
typedef struct { dual x, y, z; } dvec;
static dvec v_sub(dvec a, dvec b) { dvec r = { d_sub(a.x, b.x), d_sub(a.y, b.y), d_sub(a.z, b.z) }; return r; }
static dual v_dot(dvec a, dvec b) { return d_add(d_add(d_mul(a.x, b.x), d_mul(a.y, b.y)), d_mul(a.z, b.z)); }
static dvec v_cross(dvec a, dvec b) {
    dvec r = { d_sub(d_mul(a.y, b.z), d_mul(a.z, b.y)), d_sub(d_mul(a.z, b.x), d_mul(a.x, b.z)),
               d_sub(d_mul(a.x, b.y), d_mul(a.y, b.x)) };
    return r;
}
static dual v_norm(dvec a) { return d_sqrt(v_dot(a, a)); }
static dvec load_atom(const double* x, int atom, int slot) {
    dvec r = { d_var(x[3 * atom], 3 * slot), d_var(x[3 * atom + 1], 3 * slot + 1), d_var(x[3 * atom + 2], 3 * slot + 2) };
    return r;
}
static void scatter(double* g, const int32_t* atoms, int n, const dual* e) {
    if (!g) return;
    for (int s = 0; s < n; ++s)
        for (int c = 0; c < 3; ++c) g[3 * atoms[s] + c] += e->d[3 * s + c];
}

/* x [N,3] fp64.  e[3] = (bond, angle, torsion).  g = NULL or [3][N][3] per-term gradients
 * (zeroed here). */
void oracle_bonded(int32_t N, const double* x,
                   int32_t nb, const int32_t* bonds, const double* bond_para,
                   int32_t na, const int32_t* angles, const double* angle_para,
                   int32_t nt, const int32_t* torsions, const double* torsion_para, int32_t P,
                   double* e, double* g) {
    double* gb = g, *ga = g ? g + (size_t)3 * N : 0, *gt = g ? g + (size_t)6 * N : 0;
    if (g) memset(g, 0, sizeof(double) * 9 * (size_t)N);
    e[0] = e[1] = e[2] = 0.0;
    for (int32_t k = 0; k < nb; ++k) {
        const int32_t* a = bonds + 2 * k;
        dvec v = v_sub(load_atom(x, a[0], 0), load_atom(x, a[1], 1));
        dual dl = d_sub(v_norm(v), d_const(bond_para[2 * k]));
        dual en = d_scale(d_mul(dl, dl), bond_para[2 * k + 1]);
        if (dl.v == 0.0 && v_dot(v, v).v == 0.0) memset(en.d, 0, sizeof en.d); /* autograd: norm'(0)=0 */
        e[0] += en.v;
        scatter(gb, a, 2, &en);
    }
    for (int32_t k = 0; k < na; ++k) {
        const int32_t* a = angles + 3 * k;
        dvec rj = load_atom(x, a[1], 1);
        dvec v1 = v_sub(load_atom(x, a[0], 0), rj), v2 = v_sub(load_atom(x, a[2], 2), rj);
        dual c = d_div(v_dot(v1, v2), d_mul(v_norm(v1), v_norm(v2)));
        dual th = d_acos_clamped(c, -0.999999, 0.999999);
        dual dt = d_sub(th, d_const(angle_para[2 * k]));
        dual en = d_scale(d_mul(dt, dt), angle_para[2 * k + 1]);
        e[1] += en.v;
        scatter(ga, a, 3, &en);
    }
    for (int32_t k = 0; k < nt; ++k) {
        const int32_t* a = torsions + 4 * k;
        dvec ri = load_atom(x, a[0], 0), rj = load_atom(x, a[1], 1), rk = load_atom(x, a[2], 2), rl = load_atom(x, a[3], 3);
        dvec b1 = v_sub(ri, rj), b2 = v_sub(rj, rk), b3 = v_sub(rl, rk);
        dvec c32 = v_cross(b3, b2), c12 = v_cross(b1, b2);
        dual phi = d_atan2(v_dot(b2, v_cross(c32, c12)), d_mul(v_norm(b2), v_dot(c12, c32)));
        const double* p = torsion_para + (size_t)k * 4 * P;
        dual en = d_const(0.0);
        for (int32_t s = 0; s < P; ++s) {
            double amp = p[P + s] / p[s];
            dual arg = d_sub(d_scale(phi, p[3 * P + s]), d_const(p[2 * P + s]));
            en = d_add(en, d_scale(d_add(d_const(1.0), d_cos(arg)), amp));
        }
        e[2] += en.v;
        scatter(gt, a, 4, &en);
    }
}

static double warp(double d, double k, double* dw) {
    if (d > k) { *dw = 1.0; return d; }
    double ex = exp(d - k);
    *dw = ex;
    return ex - 1.0 + k;
}

/* Non-bonded energy and gradient of one conformation.
 * x [N,3] fp64; R, eps, q [N] (the reference's fp32 per-atom values, promoted);
 * excl [ne,2] unique pairs i<j, sorted lexicographically; full != 0 adds the -B/w^6 term.
 * g = NULL or [N,3] (zeroed here).  Returns E. */
double oracle_nonbonded(int32_t N, const double* x, const double* R, const double* eps, const double* q,
                        int32_t ne, const int32_t* excl, int32_t full, double* g) {
    double E = 0.0;
    if (g) memset(g, 0, sizeof(double) * 3 * (size_t)N);
    /* first exclusion of every row (the list is sorted by i, then j): rows can then be evaluated independently */
    int32_t* row = (int32_t*)malloc(sizeof(int32_t) * ((size_t)N + 1));
    {
        int32_t ep = 0;
        for (int32_t i = 0; i <= N; ++i) {
            while (ep < ne && excl[2 * ep] < i) ++ep;
            row[i] = ep;
        }
    }
    /* rows are spread over the host threads (OpenMP when compiled with it); every thread keeps its own energy and
     * gradient accumulators, added in thread order at the end (fp64: the order is immaterial at the 1e-13 level) */
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        double El = 0.0;
        double* gl = g ? (double*)calloc(3 * (size_t)N, sizeof(double)) : NULL;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8) nowait
#endif
        for (int32_t i = 0; i < N; ++i) {
            int32_t ep = row[i];
            for (int32_t j = i + 1; j < N; ++j) {
                while (ep < row[i + 1] && excl[2 * ep + 1] < j) ++ep;
                if (ep < row[i + 1] && excl[2 * ep + 1] == j) continue;
                double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
                double d = sqrt(dx * dx + dy * dy + dz * dz);
                double Rij = 0.5 * fabs(R[i] + R[j]), eij = sqrt(eps[i] * eps[j]);
                double R6 = Rij * Rij * Rij; R6 *= R6;
                double A = eij * R6 * R6, Bc = 2.0 * eij * R6, qq = q[i] * q[j];
                double dw1, dw2;
                double w1 = warp(d, 1.9, &dw1), w2 = warp(d, 0.4, &dw2);
                double i6 = 1.0 / (w1 * w1 * w1 * w1 * w1 * w1);
                double elj = A * i6 * i6 - (full ? Bc * i6 : 0.0);
                El += elj + qq / w2;
                if (gl && d > 0.0) {
                    double dE = (-12.0 * A * i6 * i6 + (full ? 6.0 * Bc * i6 : 0.0)) / w1 * dw1 - qq / (w2 * w2) * dw2;
                    double s = dE / d;
                    gl[3 * i] += s * dx; gl[3 * i + 1] += s * dy; gl[3 * i + 2] += s * dz;
                    gl[3 * j] -= s * dx; gl[3 * j + 1] -= s * dy; gl[3 * j + 2] -= s * dz;
                }
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            E += El;
            if (gl) for (size_t k = 0; k < 3 * (size_t)N; ++k) g[k] += gl[k];
        }
        free(gl);
    }
    free(row);
    return E;
}

/* Unaligned RMSD sqrt(mean_atoms sum_xyz (a-b)^2) (SURVEY S-3/S-4 intended semantics). */
double oracle_rmsd(int32_t N, const double* a, const double* b) {
    double s = 0.0;
    for (int32_t i = 0; i < 3 * N; ++i) { double d = a[i] - b[i]; s += d * d; }
    return sqrt(s / N);
}
