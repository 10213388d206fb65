// cpu_frame.cpp -- CPU ORACLE / CPU BASELINE of the FERS analysis hot path for plain 3D frames, C++ + OpenMP.
// TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library built from this file (oracle/build.py -> oracle/_build/libfers_cpu.so).
//
// What it restates (the reference holds no solver source -- fers_core/solver.py is empty, the arithmetic is the closed
// wheel fers_calculations==0.2.43, pyproject.toml:26 -- so this follows the published contract, file:line of
// /root/reference in every function): the configs[3] subset of oracle/fers_oracle.py (Normal and Truss members without
// modifier records, nodal loads, distributed and point member loads, combinations, linear static), with the solve done
// the way the reference describes its own first-order path, "a batch first-order solve that shares one factorization
// across many load cases" (fers_core/types/pydantic_models.py:1216-1219): a multifrontal Cholesky over the front tree of
// the product's host-side analysis phase (ordering + symbolic: plain integer arrays, exported by chol.CholPlan.export()),
// dense front kernels from BLAS-3 / LAPACK (dpotrf, dtrsm, dsyrk, dgemm of the OpenBLAS that ships inside the scipy
// wheel, bound at run time with dlopen), fronts of a tree level in parallel with OpenMP where the level is wide and
// threaded BLAS where it is narrow.  SURVEY.md 8(d) names exactly this as the CPU baseline ("an OpenMP C++ restatement").
//
// Pinning: tests/test_cpu_frame.py checks every function against oracle/fers_oracle.py (itself pinned on the reference's
// own test-suite and known-answer numbers) on seeded frames: K, load vectors, displacements, member forces, reactions.
#include <dlfcn.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace {

std::string g_error;
int fail(const std::string& m) { g_error = m; return 1; }

// ---- BLAS / LAPACK of scipy's OpenBLAS (LP64, symbols prefixed scipy_) -------------------------------------------------
enum { RowMajor = 101, ColMajor = 102, NoTrans = 111, Trans = 112, Upper = 121, Lower = 122, NonUnit = 131, Left = 141, Right = 142 };
typedef void (*dgemm_t)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*dsyrk_t)(int, int, int, int, int, double, const double*, int, double, double*, int);
typedef void (*dtrsm_t)(int, int, int, int, int, int, int, double, const double*, int, double*, int);
typedef void (*dpotrf_t)(const char*, const int*, double*, const int*, int*);
typedef void (*setthr_t)(int);
dgemm_t p_dgemm = nullptr;
dsyrk_t p_dsyrk = nullptr;
dtrsm_t p_dtrsm = nullptr;
dpotrf_t p_dpotrf = nullptr;
setthr_t p_set_threads = nullptr;
int g_threads = 0;

void blas_threads(int n) { if (p_set_threads) p_set_threads(n); }

// ---- element geometry (fers_core/members/member.py:309-354) -------------------------------------------------------------
// rows of R = local x, y, z in global coordinates; ref = +Y, or +Z when |lx . Y| > 1 - 1e-6; lz = lx x ref, ly = lz x lx; roll
struct Frame { double R[3][3]; double L; };

inline void cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
inline void member_frame(const double* xyz, int n1, int n2, double roll, Frame& f) {
    double d[3] = {xyz[3 * n2] - xyz[3 * n1], xyz[3 * n2 + 1] - xyz[3 * n1 + 1], xyz[3 * n2 + 2] - xyz[3 * n1 + 2]};
    f.L = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double lx[3] = {d[0] / f.L, d[1] / f.L, d[2] / f.L};
    double ref[3] = {0.0, 1.0, 0.0};
    if (fabs(lx[1]) > 1.0 - 1e-6) { ref[1] = 0.0; ref[2] = 1.0; }
    double lz[3], ly[3];
    cross(lx, ref, lz);
    double n = sqrt(lz[0] * lz[0] + lz[1] * lz[1] + lz[2] * lz[2]);
    for (int i = 0; i < 3; ++i) lz[i] /= n;
    cross(lz, lx, ly);
    n = sqrt(ly[0] * ly[0] + ly[1] * ly[1] + ly[2] * ly[2]);
    for (int i = 0; i < 3; ++i) ly[i] /= n;
    if (roll != 0.0) {
        const double c = cos(roll), s = sin(roll);
        for (int i = 0; i < 3; ++i) { const double y = ly[i], z = lz[i]; ly[i] = c * y + s * z; lz[i] = -s * y + c * z; }
    }
    for (int i = 0; i < 3; ++i) { f.R[0][i] = lx[i]; f.R[1][i] = ly[i]; f.R[2][i] = lz[i]; }
}

constexpr int NPROP = 8;          // EA EIy EIz GJ GAsy GAsz Wpl EIw  (wire.P_*)
constexpr int MT_TRUSS = 1;

// shear parameters (phi_y couples v / theta_z bending, EIz with G*a_sz)
inline void shear_phis(const double* P, double L, double& phi_y, double& phi_z) {
    phi_y = P[5] > 0.0 ? 12.0 * P[2] / (P[5] * L * L) : 0.0;
    phi_z = P[4] > 0.0 ? 12.0 * P[1] / (P[4] * L * L) : 0.0;
}

// 12x12 local stiffness, DOF order [u v w rx ry rz] x 2 (pydantic_models.py:1052-1114; Timoshenko when shear areas given)
inline void local_stiffness(const double* P, double L, int mtype, double k[12][12]) {
    memset(k, 0, sizeof(double) * 144);
    const bool truss = mtype == MT_TRUSS;
    const double a = P[0] / L;
    k[0][0] = k[6][6] = a; k[0][6] = k[6][0] = -a;
    if (truss) return;
    const double t = P[3] / L + 12.0 * P[7] / (L * L * L);
    k[3][3] = k[9][9] = t; k[3][9] = k[9][3] = -t;
    double phi_y, phi_z;
    shear_phis(P, L, phi_y, phi_z);
    {
        const double b = P[2] / (1.0 + phi_y);
        const double k12 = 12 * b / (L * L * L), k6 = 6 * b / (L * L), k4 = (4 + phi_y) * b / L, k2 = (2 - phi_y) * b / L;
        const int v1 = 1, r1 = 5, v2 = 7, r2 = 11;
        const int I[10] = {v1, v1, v1, v1, r1, r1, r1, v2, v2, r2}, J[10] = {v1, r1, v2, r2, r1, v2, r2, v2, r2, r2};
        const double V[10] = {k12, k6, -k12, k6, k4, -k6, k2, k12, -k6, k4};
        for (int q = 0; q < 10; ++q) { k[I[q]][J[q]] = V[q]; k[J[q]][I[q]] = V[q]; }
    }
    {
        const double b = P[1] / (1.0 + phi_z);
        const double k12 = 12 * b / (L * L * L), k6 = 6 * b / (L * L), k4 = (4 + phi_z) * b / L, k2 = (2 - phi_z) * b / L;
        const int w1 = 2, q1 = 4, w2 = 8, q2 = 10;
        const int I[10] = {w1, w1, w1, w1, q1, q1, q1, w2, w2, q2}, J[10] = {w1, q1, w2, q2, q1, w2, q2, w2, q2, q2};
        const double V[10] = {k12, -k6, -k12, -k6, k4, k6, k2, k12, k6, k4};
        for (int q = 0; q < 10; ++q) { k[I[q]][J[q]] = V[q]; k[J[q]][I[q]] = V[q]; }
    }
}

inline int find_slot(const int* rowptr, const int* colidx, int r, int c) {
    const int* b = colidx + rowptr[r];
    const int* e = colidx + rowptr[r + 1];
    const int* p = std::lower_bound(b, e, c);
    return (p != e && *p == c) ? (int)(p - colidx) : -1;
}

// Hermite shapes with shear (oracle/fers_oracle.py hermite / hermite_rot; consistent equivalent loads,
// pydantic_models.py:1438-1445, 1562-1591; exactness demanded by tests/beam_equation_checks/test_beam_equations.py:167-323)
inline void hermite(double xi, double phi, double L, double n[4]) {
    const double d = 1.0 + phi;
    n[0] = (1 - 3 * xi * xi + 2 * xi * xi * xi + phi * (1 - xi)) / d;
    n[1] = L * (xi - 2 * xi * xi + xi * xi * xi + phi * (xi - xi * xi) / 2) / d;
    n[2] = (3 * xi * xi - 2 * xi * xi * xi + phi * xi) / d;
    n[3] = L * (-xi * xi + xi * xi * xi - phi * (xi - xi * xi) / 2) / d;
}
inline void hermite_rot(double xi, double phi, double L, double g[4]) {
    const double d = 1.0 + phi;
    g[0] = 6 * (xi * xi - xi) / (L * d);
    g[1] = (1 - 4 * xi + 3 * xi * xi + phi * (1 - xi)) / d;
    g[2] = -g[0];
    g[3] = (3 * xi * xi - 2 * xi + phi * xi) / d;
}
// integrals of a linear load (q1 at a, q2 at b) against the shapes and against (1 - xi), xi; 3-point Gauss is exact
inline void shape_integrals(double a, double b, double q1, double q2, double phi, double L, double out[6]) {
    static const double GX[3] = {-0.7745966692414834, 0.0, 0.7745966692414834}, GW[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    for (int i = 0; i < 6; ++i) out[i] = 0.0;
    const double half = 0.5 * (b - a), mid = 0.5 * (a + b);
    if (half <= 0.0) return;
    for (int g = 0; g < 3; ++g) {
        const double xi = mid + half * GX[g];
        const double q = q1 + (q2 - q1) * (xi - a) / (b - a);
        const double w = GW[g] * half * L * q;
        double n[4];
        hermite(xi, phi, L, n);
        out[0] += w * n[0]; out[1] += w * n[1]; out[2] += w * n[2]; out[3] += w * n[3];
        out[4] += w * (1.0 - xi); out[5] += w * xi;
    }
}

// ---- front tree (the arrays of chol.CholPlan.export(); layouts of fers_core_b200/csrc/chol.cuh) ------------------------
struct FrontRec {
    int n1, n2, nodes_off, child_begin, child_end, parent, pad0, pad1;
    long long panel_off, schur_off, dinv_off, rhs_off;
};
struct KScatterRec { int front, k, bi, bj; };

}  // namespace

extern "C" {

struct FersCpuPlan {
    int n_nodes, n_fronts, n_levels, reserved;
    const FrontRec* fronts;
    const int* nodes;
    const int* child;
    const long long* inv_off;
    const int* inv;
    const int* level_ptr;          // [n_levels + 1]
    const int* level_fronts;
    const KScatterRec* kscatter;
    const long long* kscatter_ptr; // [n_levels + 1]
    const int* pos;                // node -> elimination position or -1
    long long factor_doubles, schur_doubles0, schur_doubles1, rhs_rows;
};

const char* fers_cpu_last_error(void) { return g_error.c_str(); }

int fers_cpu_init(const char* openblas_path) {
    if (p_dgemm) return 0;
    void* h = dlopen(openblas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(std::string("dlopen failed: ") + dlerror());
    auto sym = [&](const char* name) -> void* {
        for (const char* pre : {"scipy_", ""}) {
            void* p = dlsym(h, (std::string(pre) + name).c_str());
            if (p) return p;
        }
        return nullptr;
    };
    p_dgemm = (dgemm_t)sym("cblas_dgemm");
    p_dsyrk = (dsyrk_t)sym("cblas_dsyrk");
    p_dtrsm = (dtrsm_t)sym("cblas_dtrsm");
    p_dpotrf = (dpotrf_t)sym("dpotrf_");
    p_set_threads = (setthr_t)sym("openblas_set_num_threads");
    if (!p_dgemm || !p_dsyrk || !p_dtrsm || !p_dpotrf) return fail("OpenBLAS symbols (cblas_dgemm, cblas_dsyrk, cblas_dtrsm, dpotrf_) not found");
    g_threads = omp_get_max_threads();
    return 0;
}

int fers_cpu_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }
void fers_cpu_set_threads(int n) {
    g_threads = n < 1 ? 1 : n;
    omp_set_num_threads(g_threads);
    blas_threads(g_threads);
}

// ---- element stiffness + assembly into the block-CSR values (T^T k T) -------------------------
int fers_cpu_assemble(int n_nodes, int n_members, const double* xyz, const int* n1, const int* n2, const int* m_prop,
                      const int* m_type, const double* roll, const double* prop, const int* rowptr, const int* colidx,
                      double* vals) {
    const long long nnzb = rowptr[n_nodes];
    memset(vals, 0, sizeof(double) * 36 * (size_t)nnzb);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int m = 0; m < n_members; ++m) {
        Frame f;
        member_frame(xyz, n1[m], n2[m], roll[m], f);
        double k[12][12];
        local_stiffness(prop + (size_t)NPROP * m_prop[m], f.L, m_type[m], k);
        const int nodes[2] = {n1[m], n2[m]};
        for (int p = 0; p < 2; ++p)
            for (int q = 0; q < 2; ++q) {
                const int slot = find_slot(rowptr, colidx, nodes[p], nodes[q]);
                if (slot < 0) { bad |= 1; continue; }
                double* blk = vals + (size_t)36 * slot;
                // 6x6 block = four 3x3 blocks R^T k_sub R
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        double tmp[3][3], out[3][3];
                        for (int i = 0; i < 3; ++i)
                            for (int j = 0; j < 3; ++j) {
                                double s = 0.0;
                                for (int l = 0; l < 3; ++l) s += k[6 * p + 3 * a + i][6 * q + 3 * b + l] * f.R[l][j];
                                tmp[i][j] = s;
                            }
                        for (int i = 0; i < 3; ++i)
                            for (int j = 0; j < 3; ++j) {
                                double s = 0.0;
                                for (int l = 0; l < 3; ++l) s += f.R[l][i] * tmp[l][j];
                                out[i][j] = s;
                            }
                        for (int i = 0; i < 3; ++i)
                            for (int j = 0; j < 3; ++j) {
                                double* dst = blk + (3 * a + i) * 6 + 3 * b + j;
#pragma omp atomic
                                *dst += out[i][j];
                            }
                    }
            }
    }
    return bad ? fail("assemble: member block missing from the pattern") : 0;
}

// ---- one load case: nodal loads, distributed and point member loads -----------------------------------
// dl_data rows [q1(3) q2(3) a b] in global axes; pl_data rows [F(3) M(3) xi is_local]
int fers_cpu_load_case(int n_nodes, int n_members, const double* xyz, const int* n1, const int* n2, const int* m_prop,
                       const int* m_type, const double* roll, const double* prop, int n_nod, const int* nod_node,
                       const double* nod_vec, int n_dl, const int* dl_member, const double* dl_data, int n_pl,
                       const int* pl_member, const double* pl_data, double* F, double* peq) {
    memset(F, 0, sizeof(double) * 6 * (size_t)n_nodes);
    memset(peq, 0, sizeof(double) * 12 * (size_t)n_members);
    for (int i = 0; i < n_nod; ++i)
        for (int d = 0; d < 6; ++d) F[6 * (size_t)nod_node[i] + d] += nod_vec[6 * (size_t)i + d];
    auto rot = [](const Frame& f, const double* g, double* l) {
        for (int i = 0; i < 3; ++i) l[i] = f.R[i][0] * g[0] + f.R[i][1] * g[1] + f.R[i][2] * g[2];
    };
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_dl + n_pl; ++i) {
        const bool dist = i < n_dl;
        const int m = dist ? dl_member[i] : pl_member[i - n_dl];
        Frame f;
        member_frame(xyz, n1[m], n2[m], roll[m], f);
        const double* P = prop + (size_t)NPROP * m_prop[m];
        double phi_y, phi_z;
        shear_phis(P, f.L, phi_y, phi_z);
        const bool truss = m_type[m] == MT_TRUSS;
        if (truss) phi_y = phi_z = 0.0;
        double p[12] = {0};
        if (dist) {
            const double* row = dl_data + 8 * (size_t)i;
            double q1[3], q2[3], s[6];
            rot(f, row, q1); rot(f, row + 3, q2);
            const double a = row[6], b = row[7];
            shape_integrals(a, b, q1[0], q2[0], 0.0, f.L, s);
            p[0] += s[4]; p[6] += s[5];
            double sy[6], sz[6];
            shape_integrals(a, b, q1[1], q2[1], phi_y, f.L, sy);
            shape_integrals(a, b, q1[2], q2[2], phi_z, f.L, sz);
            if (truss) {
                p[1] += sy[4]; p[7] += sy[5]; p[2] += sz[4]; p[8] += sz[5];
            } else {
                p[1] += sy[0]; p[5] += sy[1]; p[7] += sy[2]; p[11] += sy[3];
                p[2] += sz[0]; p[4] -= sz[1]; p[8] += sz[2]; p[10] -= sz[3];
            }
        } else {
            const double* row = pl_data + 8 * (size_t)(i - n_dl);
            double Fl[3] = {row[0], row[1], row[2]}, Ml[3] = {row[3], row[4], row[5]};
            const double xi = row[6];
            if (row[7] == 0.0) { double t[3]; rot(f, row, t); memcpy(Fl, t, sizeof t); rot(f, row + 3, t); memcpy(Ml, t, sizeof t); }
            p[0] += (1 - xi) * Fl[0]; p[6] += xi * Fl[0];
            p[3] += (1 - xi) * Ml[0]; p[9] += xi * Ml[0];
            if (truss) {
                p[1] += (1 - xi) * Fl[1]; p[7] += xi * Fl[1]; p[2] += (1 - xi) * Fl[2]; p[8] += xi * Fl[2];
            } else {
                double n[4], g[4];
                hermite(xi, phi_y, f.L, n); hermite_rot(xi, phi_y, f.L, g);
                const int dy[4] = {1, 5, 7, 11};
                for (int q = 0; q < 4; ++q) p[dy[q]] += n[q] * Fl[1] + g[q] * Ml[2];
                hermite(xi, phi_z, f.L, n); hermite_rot(xi, phi_z, f.L, g);
                const int dz[4] = {2, 4, 8, 10};
                const double sg[4] = {1.0, -1.0, 1.0, -1.0};
                for (int q = 0; q < 4; ++q) p[dz[q]] += sg[q] * n[q] * Fl[2] - sg[q] * g[q] * Ml[1];
            }
        }
        // peq (local) and F += T^T p
        const int nodes[2] = {n1[m], n2[m]};
        for (int d = 0; d < 12; ++d) {
#pragma omp atomic
            peq[12 * (size_t)m + d] += p[d];
        }
        for (int e = 0; e < 2; ++e)
            for (int h = 0; h < 2; ++h)
                for (int j = 0; j < 3; ++j) {
                    double s = 0.0;
                    for (int l = 0; l < 3; ++l) s += f.R[l][j] * p[6 * e + 3 * h + l];
#pragma omp atomic
                    F[6 * (size_t)nodes[e] + 3 * h + j] += s;
                }
    }
    return 0;
}

// ---- combinations: out[row][c] = sum_l fac[c][l] per_case[l][row]   ------------------------------------
int fers_cpu_combine(long long rows, int nlc, int nr, const double* per_case, const double* fac, double* out) {
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < rows; ++r) {
        double* o = out + (size_t)r * nr;
        for (int c = 0; c < nr; ++c) o[c] = 0.0;
        for (int l = 0; l < nlc; ++l) {
            const double v = per_case[(size_t)l * rows + r];
            if (v == 0.0) continue;
            for (int c = 0; c < nr; ++c) o[c] += fac[(size_t)c * nlc + l] * v;
        }
    }
    return 0;
}

// ---- y = K x for multi-vectors [Ndof][nr] (reactions R = K u - F, residual check) --------------------------------------------
int fers_cpu_spmm(int n_nodes, int nr, const int* rowptr, const int* colidx, const double* vals, const double* x, double* y) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n_nodes; ++i) {
        double* yi = y + (size_t)6 * i * nr;
        for (int q = 0; q < 6 * nr; ++q) yi[q] = 0.0;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const double* b = vals + (size_t)36 * e;
            const double* xj = x + (size_t)6 * colidx[e] * nr;
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < 6; ++c) {
                    const double v = b[6 * a + c];
                    if (v == 0.0) continue;
                    const double* xr = xj + (size_t)c * nr;
                    double* yr = yi + (size_t)a * nr;
                    for (int r = 0; r < nr; ++r) yr[r] += v * xr[r];
                }
        }
    }
    return 0;
}

// ---- multifrontal Cholesky: numeric factorisation ---------------------------------------------------
// factor: [factor_doubles] panels (column-major [6(n1+n2)] x [6 n1] at panel_off) followed by the equilibration scales
// sc[6 n_nodes]; schur0 / schur1: the two Schur arenas (levels alternate).  info: 0 or the 1-based elimination-order DOF of the
// first non-positive pivot.
int fers_cpu_mf_factor(const FersCpuPlan* P, const double* vals, const int* diag_slot, const uint8_t* fixed,
                       const double* kspring, double* factor, double* schur0, double* schur1, int* info) {
    if (!p_dgemm) return fail("fers_cpu_init was not called");
    const int threads = fers_cpu_threads();
    double* sc = factor + P->factor_doubles;
    const int nd = 6 * P->n_nodes;
#pragma omp parallel for schedule(static)
    for (int d = 0; d < nd; ++d) {
        const int n = d / 6, a = d - 6 * n;
        double k = vals[(size_t)36 * diag_slot[n] + 7 * a];
        if (kspring) k += kspring[d];
        sc[d] = (fixed[d] || !(k > 0.0)) ? 1.0 : 1.0 / sqrt(k);
    }
    double* arena[2] = {schur0, schur1};
    int first_bad = 0;
    for (int l = 0; l < P->n_levels; ++l) {
        const int f0 = P->level_ptr[l], f1 = P->level_ptr[l + 1];
        const int nf = f1 - f0;
        if (nf == 0) continue;
        double* cur = arena[l & 1];
        const double* prev = arena[(l & 1) ^ 1];
        const bool wide = nf >= 2 * threads;          // fronts in parallel, sequential BLAS; else threaded BLAS front by front
        // (1) extend-add of the children's Schur complements
        auto init_front = [&](int fi, bool inner_parallel) {
            const FrontRec& f = P->fronts[fi];
            const int nt = f.n1 + f.n2, m = 6 * nt, k1 = 6 * f.n1, m2 = 6 * f.n2;
            double* panel = factor + f.panel_off;
            double* S = cur + f.schur_off;
#pragma omp parallel for schedule(static) if (inner_parallel)
            for (int c = 0; c < m; ++c) {
                if (c < k1) memset(panel + (size_t)c * m + c, 0, sizeof(double) * (m - c));
                else memset(S + (size_t)(c - k1) * m2 + (c - k1), 0, sizeof(double) * (m - c));
            }
            for (int ce = f.child_begin; ce < f.child_end; ++ce) {
                const int* mp = P->inv + P->inv_off[ce];
                const FrontRec& fc = P->fronts[P->child[ce]];
                const size_t ldc = 6 * (size_t)fc.n2;
                const double* Sc = prev + fc.schur_off;
                // parent-local nodes present in the child's boundary, in order (both lists follow the elimination order)
                std::vector<int> here;
                for (int b = 0; b < nt; ++b) if (mp[b] >= 0) here.push_back(b);
                const int nh = (int)here.size();
#pragma omp parallel for schedule(dynamic, 4) if (inner_parallel)
                for (int jj = 0; jj < nh; ++jj) {
                    const int bj = here[jj], cj = mp[bj];
                    for (int b = 0; b < 6; ++b) {
                        const int c = 6 * bj + b;
                        double* dst = c < k1 ? panel + (size_t)c * m : S + (size_t)(c - k1) * m2 - k1;
                        const double* src = Sc + (size_t)(6 * cj + b) * ldc;
                        for (int ii = jj; ii < nh; ++ii) {
                            const int bi = here[ii], ci = mp[bi];
                            const int a0 = (ii == jj) ? b : 0;       // lower triangle of the diagonal block only
                            for (int a = a0; a < 6; ++a) dst[6 * bi + a] += src[6 * ci + a];
                        }
                    }
                }
            }
        };
        if (wide) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int q = f0; q < f1; ++q) init_front(P->level_fronts[q], false);
        } else {
            for (int q = f0; q < f1; ++q) init_front(P->level_fronts[q], true);
        }
        // (2) the stiffness blocks of the own nodes (equilibrated; constrained DOFs become identity rows)
        const long long ks0 = P->kscatter_ptr[l], ks1 = P->kscatter_ptr[l + 1];
#pragma omp parallel for schedule(static)
        for (long long w = ks0; w < ks1; ++w) {
            const KScatterRec e = P->kscatter[w];
            const FrontRec& f = P->fronts[e.front];
            const int m = 6 * (f.n1 + f.n2);
            const int gi = P->nodes[f.nodes_off + e.bi], gj = P->nodes[f.nodes_off + e.bj];
            double* panel = factor + f.panel_off;
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) {
                    if (e.bi == e.bj && a < b) continue;
                    const int dr = 6 * gi + a, dc = 6 * gj + b;
                    double v;
                    if (fixed[dr] || fixed[dc]) v = (dr == dc) ? 1.0 : 0.0;
                    else {
                        v = vals[(size_t)36 * e.k + 6 * a + b];
                        if (dr == dc && kspring) v += kspring[dr];
                        v *= sc[dr] * sc[dc];
                    }
                    panel[(size_t)(6 * e.bj + b) * m + 6 * e.bi + a] += v;
                }
        }
        // (3) dense partial factorisation of every front: L11 = chol(F11), L21 = F21 L11^-T, F22 -= L21 L21^T
        auto factor_front = [&](int fi) {
            const FrontRec& f = P->fronts[fi];
            const int m = 6 * (f.n1 + f.n2), k1 = 6 * f.n1, m2 = 6 * f.n2;
            double* panel = factor + f.panel_off;
            int inf = 0;
            p_dpotrf("L", &k1, panel, &m, &inf);
            if (inf > 0) {
#pragma omp critical
                { const int g = 6 * f.pad0 + inf; if (first_bad == 0 || g < first_bad) first_bad = g; }
                return;
            }
            if (m2 == 0) return;
            p_dtrsm(ColMajor, Right, Lower, Trans, NonUnit, m2, k1, 1.0, panel, m, panel + k1, m);
            if (f.parent >= 0)
                p_dsyrk(ColMajor, Lower, NoTrans, m2, k1, -1.0, panel + k1, m, 1.0, cur + f.schur_off, m2);
        };
        if (wide) {
            blas_threads(1);
#pragma omp parallel for schedule(dynamic, 1)
            for (int q = f0; q < f1; ++q) factor_front(P->level_fronts[q]);
            blas_threads(threads);
        } else {
            for (int q = f0; q < f1; ++q) factor_front(P->level_fronts[q]);
        }
        if (first_bad) break;
    }
    *info = first_bad;
    return 0;
}

// ---- multifrontal substitution for nrhs right-hand sides, b and x [6 n_nodes][nrhs] (right-hand side index fastest) ----------
// work: (rhs_rows + 6 n_nodes) * nrhs doubles
int fers_cpu_mf_solve(const FersCpuPlan* P, const double* factor, const uint8_t* fixed, int nrhs, const double* b, double* x,
                      double* work) {
    if (!p_dgemm) return fail("fers_cpu_init was not called");
    const int threads = fers_cpu_threads();
    const double* sc = factor + P->factor_doubles;
    double* W = work;
    double* Y = work + (size_t)P->rhs_rows * nrhs;
    const int nr = nrhs;
    for (int l = 0; l < P->n_levels; ++l) {                 // forward: children -> parents
        const int f0 = P->level_ptr[l], f1 = P->level_ptr[l + 1];
        const bool wide = (f1 - f0) >= 2 * threads;
        auto fwd = [&](int fi) {
            const FrontRec& f = P->fronts[fi];
            const int nt = f.n1 + f.n2, m = 6 * nt, k1 = 6 * f.n1, m2 = 6 * f.n2;
            double* Wf = W + (size_t)f.rhs_off * nr;
            for (int q = 0; q < nt; ++q)
                for (int a = 0; a < 6; ++a) {
                    double* w = Wf + (size_t)(6 * q + a) * nr;
                    if (q < f.n1) {
                        const int g = 6 * P->nodes[f.nodes_off + q] + a;
                        if (fixed[g]) for (int r = 0; r < nr; ++r) w[r] = 0.0;
                        else for (int r = 0; r < nr; ++r) w[r] = b[(size_t)g * nr + r] * sc[g];
                    } else for (int r = 0; r < nr; ++r) w[r] = 0.0;
                }
            for (int ce = f.child_begin; ce < f.child_end; ++ce) {
                const int* mp = P->inv + P->inv_off[ce];
                const FrontRec& fc = P->fronts[P->child[ce]];
                for (int q = 0; q < nt; ++q) {
                    const int ci = mp[q];
                    if (ci < 0) continue;
                    const double* src = W + (size_t)(fc.rhs_off + 6 * (fc.n1 + ci)) * nr;
                    double* dst = Wf + (size_t)6 * q * nr;
                    for (int t = 0; t < 6 * nr; ++t) dst[t] += src[t];
                }
            }
            const double* panel = factor + f.panel_off;
            // own rows: (nr x k1, ld nr) <- (nr x k1) L11^-T, i.e. y = L11^-1 w
            p_dtrsm(ColMajor, Right, Lower, Trans, NonUnit, nr, k1, 1.0, panel, m, Wf, nr);
            if (m2) p_dgemm(ColMajor, NoTrans, Trans, nr, m2, k1, -1.0, Wf, nr, panel + k1, m, 1.0, Wf + (size_t)k1 * nr, nr);
            memcpy(Y + (size_t)6 * f.pad0 * nr, Wf, sizeof(double) * (size_t)k1 * nr);
        };
        if (wide) {
            blas_threads(1);
#pragma omp parallel for schedule(dynamic, 1)
            for (int q = f0; q < f1; ++q) fwd(P->level_fronts[q]);
            blas_threads(threads);
        } else
            for (int q = f0; q < f1; ++q) fwd(P->level_fronts[q]);
    }
    for (int l = P->n_levels - 1; l >= 0; --l) {            // backward: parents -> children
        const int f0 = P->level_ptr[l], f1 = P->level_ptr[l + 1];
        const bool wide = (f1 - f0) >= 2 * threads;
        auto bwd = [&](int fi) {
            const FrontRec& f = P->fronts[fi];
            const int m = 6 * (f.n1 + f.n2), k1 = 6 * f.n1, m2 = 6 * f.n2;
            double* Wf = W + (size_t)f.rhs_off * nr;
            double* Yf = Y + (size_t)6 * f.pad0 * nr;
            const double* panel = factor + f.panel_off;
            for (int q = 0; q < f.n2; ++q) {
                const int node = P->nodes[f.nodes_off + f.n1 + q];
                memcpy(Wf + (size_t)(k1 + 6 * q) * nr, Y + (size_t)6 * P->pos[node] * nr, sizeof(double) * 6 * nr);
            }
            if (m2) p_dgemm(ColMajor, NoTrans, NoTrans, nr, k1, m2, -1.0, Wf + (size_t)k1 * nr, nr, panel + k1, m, 1.0, Yf, nr);
            p_dtrsm(ColMajor, Right, Lower, NoTrans, NonUnit, nr, k1, 1.0, panel, m, Yf, nr);
        };
        if (wide) {
            blas_threads(1);
#pragma omp parallel for schedule(dynamic, 1)
            for (int q = f0; q < f1; ++q) bwd(P->level_fronts[q]);
            blas_threads(threads);
        } else
            for (int q = f0; q < f1; ++q) bwd(P->level_fronts[q]);
    }
    const int nd = 6 * P->n_nodes;
#pragma omp parallel for schedule(static)
    for (int d = 0; d < nd; ++d) {
        const int n = d / 6, a = d - 6 * n;
        const int pn = P->pos[n];
        double* xd = x + (size_t)d * nr;
        if (fixed[d] || pn < 0) for (int r = 0; r < nr; ++r) xd[r] = 0.0;
        else {
            const double* y = Y + (size_t)(6 * pn + a) * nr;
            for (int r = 0; r < nr; ++r) xd[r] = sc[d] * y[r];
        }
    }
    return 0;
}

// ---- member recovery: local end displacements, local / global end forces (stiffness sign: k u - p_eq) ---------
// U [Ndof][nr], peq [Nm*12][nr]; outputs [Nm][12][nr]
int fers_cpu_recover_members(int n_members, int nr, const double* xyz, const int* n1, const int* n2, const int* m_prop,
                             const int* m_type, const double* roll, const double* prop, const double* U, const double* peq,
                             double* f_local, double* f_global, double* u_local) {
#pragma omp parallel
    {
        std::vector<double> ul(12 * (size_t)nr), fl(12 * (size_t)nr);
#pragma omp for schedule(static)
        for (int m = 0; m < n_members; ++m) {
            Frame f;
            member_frame(xyz, n1[m], n2[m], roll[m], f);
            double k[12][12];
            local_stiffness(prop + (size_t)NPROP * m_prop[m], f.L, m_type[m], k);
            const int nodes[2] = {n1[m], n2[m]};
            for (int e = 0; e < 2; ++e)
                for (int h = 0; h < 2; ++h)
                    for (int i = 0; i < 3; ++i) {
                        double* o = ul.data() + (size_t)(6 * e + 3 * h + i) * nr;
                        const double* u0 = U + (size_t)(6 * nodes[e] + 3 * h) * nr;
                        for (int r = 0; r < nr; ++r) o[r] = f.R[i][0] * u0[r] + f.R[i][1] * u0[nr + r] + f.R[i][2] * u0[2 * nr + r];
                    }
            const double* pe = peq + (size_t)12 * m * nr;
            for (int i = 0; i < 12; ++i) {
                double* o = fl.data() + (size_t)i * nr;
                for (int r = 0; r < nr; ++r) o[r] = -pe[(size_t)i * nr + r];
                for (int j = 0; j < 12; ++j) {
                    const double v = k[i][j];
                    if (v == 0.0) continue;
                    const double* uj = ul.data() + (size_t)j * nr;
                    for (int r = 0; r < nr; ++r) o[r] += v * uj[r];
                }
            }
            double* ol = f_local + (size_t)12 * m * nr;
            double* og = f_global + (size_t)12 * m * nr;
            double* ou = u_local + (size_t)12 * m * nr;
            memcpy(ol, fl.data(), sizeof(double) * 12 * nr);
            memcpy(ou, ul.data(), sizeof(double) * 12 * nr);
            for (int blk = 0; blk < 4; ++blk)
                for (int j = 0; j < 3; ++j) {
                    double* o = og + (size_t)(3 * blk + j) * nr;
                    const double* s = fl.data() + (size_t)(3 * blk) * nr;
                    for (int r = 0; r < nr; ++r) o[r] = f.R[0][j] * s[r] + f.R[1][j] * s[nr + r] + f.R[2][j] * s[2 * nr + r];
                }
        }
    }
    return 0;
}

}  // extern "C"
