// fers_elem.cuh -- element-level arithmetic of the FERS hot path (fp64), shared by every kernel.
//
// All functions are FERS_HD (= __host__ __device__ under nvcc, plain inline under g++) so the
// same code is exercised on the CPU by tests/test_elem_host.py against the oracle before it
// ever runs on a GPU.  Nothing in the product path calls the host instantiation.
//
// Reference semantics restated here (the reference ships no solver source):
//   local axes ............ fers_core/members/member.py:309-354
//   section properties .... fers_core/types/pydantic_models.py:1052-1114 (shear pairing: SURVEY.md A.5)
//   member types .......... pydantic_models.py:592-598
//   P-Delta K_g ........... pydantic_models.py:1377-1380
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FERS_HD __host__ __device__ __forceinline__
#else
#define FERS_HD inline
#endif

namespace fers {

enum MemberType { MT_NORMAL = 0, MT_TRUSS = 1, MT_TENSION = 2, MT_COMPRESSION = 3, MT_RIGID = 4, MT_CABLE = 5 };
enum { P_EA = 0, P_EIY = 1, P_EIZ = 2, P_GJ = 3, P_GASY = 4, P_GASZ = 5, P_EIW = 7, NPROP = 8 };

// Local frame of a member: R rows = local x, y, z expressed in global coordinates.
struct Frame {
    double r[3][3];
    double L;
};

FERS_HD Frame member_frame(const double a[3], const double b[3], double roll) {
    Frame f;
    double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    double L = sqrt(dx * dx + dy * dy + dz * dz);
    f.L = L;
    double inv = 1.0 / L;
    double lx0 = dx * inv, lx1 = dy * inv, lx2 = dz * inv;
    // reference vector: +Y, or +Z for (near-)vertical members
    double rf0 = 0.0, rf1 = 1.0, rf2 = 0.0;
    if (fabs(lx1) > 1.0 - 1e-6) { rf1 = 0.0; rf2 = 1.0; }
    // lz = lx x ref
    double z0 = lx1 * rf2 - lx2 * rf1, z1 = lx2 * rf0 - lx0 * rf2, z2 = lx0 * rf1 - lx1 * rf0;
    double zn = 1.0 / sqrt(z0 * z0 + z1 * z1 + z2 * z2);
    z0 *= zn; z1 *= zn; z2 *= zn;
    // ly = lz x lx
    double y0 = z1 * lx2 - z2 * lx1, y1 = z2 * lx0 - z0 * lx2, y2 = z0 * lx1 - z1 * lx0;
    double yn = 1.0 / sqrt(y0 * y0 + y1 * y1 + y2 * y2);
    y0 *= yn; y1 *= yn; y2 *= yn;
    if (roll != 0.0) {
        double c = cos(roll), s = sin(roll);
        double ny0 = c * y0 + s * z0, ny1 = c * y1 + s * z1, ny2 = c * y2 + s * z2;
        double nz0 = -s * y0 + c * z0, nz1 = -s * y1 + c * z1, nz2 = -s * y2 + c * z2;
        y0 = ny0; y1 = ny1; y2 = ny2; z0 = nz0; z1 = nz1; z2 = nz2;
    }
    f.r[0][0] = lx0; f.r[0][1] = lx1; f.r[0][2] = lx2;
    f.r[1][0] = y0;  f.r[1][1] = y1;  f.r[1][2] = y2;
    f.r[2][0] = z0;  f.r[2][1] = z1;  f.r[2][2] = z2;
    return f;
}

// Closed-form coefficients of the 12x12 local stiffness (Euler-Bernoulli / Timoshenko).
struct BeamCoef {
    double a;                    // EA/L
    double t;                    // GJ/L
    double z12, z6, z4, z2;      // bending about local z (v, theta_z): 12EIz/L^3(1+phi_y) ...
    double y12, y6, y4, y2;      // bending about local y (w, theta_y)
    double phi_y, phi_z;
};

FERS_HD BeamCoef beam_coef(const double* P, double L, int mtype, double scale) {
    BeamCoef c;
    double EA = P[P_EA], EIy = P[P_EIY], EIz = P[P_EIZ], GJ = P[P_GJ], GAsy = P[P_GASY], GAsz = P[P_GASZ];
    double L2 = L * L, L3 = L2 * L;
    c.phi_y = (GAsz > 0.0) ? 12.0 * EIz / (GAsz * L2) : 0.0;
    c.phi_z = (GAsy > 0.0) ? 12.0 * EIy / (GAsy * L2) : 0.0;
    bool axial_only = (mtype == MT_TRUSS);
    c.a = scale * (EA / L);
    // torsion: St-Venant GJ/L plus the theta_x part 12 EIw / L^3 of the Hermite warping matrix (warping.cu; 0 without)
    c.t = axial_only ? 0.0 : scale * (GJ / L + 12.0 * P[P_EIW] / L3);
    double bz = axial_only ? 0.0 : EIz / (1.0 + c.phi_y);
    double by = axial_only ? 0.0 : EIy / (1.0 + c.phi_z);
    c.z12 = scale * (12.0 * bz / L3); c.z6 = scale * (6.0 * bz / L2);
    c.z4 = scale * ((4.0 + c.phi_y) * bz / L); c.z2 = scale * ((2.0 - c.phi_y) * bz / L);
    c.y12 = scale * (12.0 * by / L3); c.y6 = scale * (6.0 * by / L2);
    c.y4 = scale * ((4.0 + c.phi_z) * by / L); c.y2 = scale * ((2.0 - c.phi_z) * by / L);
    return c;
}

// Dense 12x12 local stiffness, row-major, DOF order [u v w rx ry rz] x {start, end}.
FERS_HD void beam_local_k(const BeamCoef& c, double* k /*144*/) {
    for (int i = 0; i < 144; ++i) k[i] = 0.0;
#define KS(i, j, v) { k[(i) * 12 + (j)] = (v); k[(j) * 12 + (i)] = (v); }
    KS(0, 0, c.a) KS(0, 6, -c.a) KS(6, 6, c.a)
    KS(3, 3, c.t) KS(3, 9, -c.t) KS(9, 9, c.t)
    KS(1, 1, c.z12) KS(1, 5, c.z6) KS(1, 7, -c.z12) KS(1, 11, c.z6)
    KS(5, 5, c.z4) KS(5, 7, -c.z6) KS(5, 11, c.z2)
    KS(7, 7, c.z12) KS(7, 11, -c.z6) KS(11, 11, c.z4)
    KS(2, 2, c.y12) KS(2, 4, -c.y6) KS(2, 8, -c.y12) KS(2, 10, -c.y6)
    KS(4, 4, c.y4) KS(4, 8, c.y6) KS(4, 10, c.y2)
    KS(8, 8, c.y12) KS(8, 10, c.y6) KS(10, 10, c.y4)
#undef KS
}

// Local geometric stiffness added into k (tension-positive axial force N).
// consistent != 0: Przemieniecki string+bending terms; else N/L on transverse translations.
FERS_HD void beam_add_kg(double* k /*144*/, double N, double L, int consistent, int mtype) {
#define KA(i, j, v) { k[(i) * 12 + (j)] += (v); if ((i) != (j)) k[(j) * 12 + (i)] += (v); }
    if (!consistent || mtype == MT_TRUSS) {
        double c = N / L;
        KA(1, 1, c) KA(7, 7, c) KA(1, 7, -c)
        KA(2, 2, c) KA(8, 8, c) KA(2, 8, -c)
    } else {
        double c = N / (30.0 * L), L2 = L * L;
        KA(1, 1, 36.0 * c) KA(1, 5, 3.0 * L * c) KA(1, 7, -36.0 * c) KA(1, 11, 3.0 * L * c)
        KA(5, 5, 4.0 * L2 * c) KA(5, 7, -3.0 * L * c) KA(5, 11, -L2 * c)
        KA(7, 7, 36.0 * c) KA(7, 11, -3.0 * L * c) KA(11, 11, 4.0 * L2 * c)
        KA(2, 2, 36.0 * c) KA(2, 4, -3.0 * L * c) KA(2, 8, -36.0 * c) KA(2, 10, -3.0 * L * c)
        KA(4, 4, 4.0 * L2 * c) KA(4, 8, 3.0 * L * c) KA(4, 10, -L2 * c)
        KA(8, 8, 36.0 * c) KA(8, 10, 3.0 * L * c) KA(10, 10, 4.0 * L2 * c)
    }
#undef KA
}

// Static condensation of end-release springs (pydantic_models.py:2062-2087): for each local
// end DOF d with spring ks >= 0 (ks < 0: rigid connection):  k -= k[:,d] k[d,:] / (k[d,d]+ks),
// p -= k[:,d] p[d] / (k[d,d]+ks)  for each of the np load rows p (stride 12).
FERS_HD void beam_condense(double* k /*144*/, const double* rel /*12*/, double* p, int np) {
    for (int d = 0; d < 12; ++d) {
        double ks = rel[d];
        if (ks < 0.0) continue;
        double den = k[d * 12 + d] + ks;
        if (!(den > 0.0)) continue;
        double col[12];
        for (int i = 0; i < 12; ++i) col[i] = k[i * 12 + d];
        double inv = 1.0 / den;
        for (int q = 0; q < np; ++q) {
            double s = p[q * 12 + d] * inv;
            for (int i = 0; i < 12; ++i) p[q * 12 + i] -= s * col[i];
        }
        for (int i = 0; i < 12; ++i) {
            double s = col[i] * inv;
            for (int j = 0; j < 12; ++j) k[i * 12 + j] -= s * col[j];
        }
    }
}

// local = R * global (each 3-vector of a 6-vector pair) and its transpose
FERS_HD void rot_to_local(const Frame& f, const double* g, double* l) {
    for (int i = 0; i < 3; ++i) l[i] = f.r[i][0] * g[0] + f.r[i][1] * g[1] + f.r[i][2] * g[2];
}
FERS_HD void rot_to_global(const Frame& f, const double* l, double* g) {
    for (int i = 0; i < 3; ++i) g[i] = f.r[0][i] * l[0] + f.r[1][i] * l[1] + f.r[2][i] * l[2];
}

// Shear-centre eccentricity (pydantic_models.py:1106-1111 `y_s`, `z_s`; toggle `include_shear_center_coupling`
// 1358-1361, "T^T.K.T transformation"; `mirror` negates both, 1521-1524).  The element's transverse
// displacements and its twist are those of the shear-centre axis, the nodes sit on the centroid axis:
//   v_sc = v - z_s * theta_x ,  w_sc = w + y_s * theta_x     (rigid section, both ends)
// so  k <- S^T k S,  p <- S^T p  with  S = I + E,  E[v][rx] = -z_s, E[w][rx] = y_s per end.
FERS_HD void beam_shear_centre(double* k /*144*/, double ys, double zs, double* p, int np) {
    if (ys == 0.0 && zs == 0.0) return;
    for (int e = 0; e < 2; ++e) {
        const int v = 6 * e + 1, w = 6 * e + 2, t = 6 * e + 3;
        for (int i = 0; i < 12; ++i) k[i * 12 + t] += ys * k[i * 12 + w] - zs * k[i * 12 + v];   // k S
    }
    for (int e = 0; e < 2; ++e) {
        const int v = 6 * e + 1, w = 6 * e + 2, t = 6 * e + 3;
        for (int j = 0; j < 12; ++j) k[t * 12 + j] += ys * k[w * 12 + j] - zs * k[v * 12 + j];   // S^T (k S)
        for (int q = 0; q < np; ++q) p[q * 12 + t] += ys * p[q * 12 + w] - zs * p[q * 12 + v];
    }
}

// Per-member modifier record (REL_STRIDE doubles, one row of the `rel` table): [0..11] end-release springs,
// [12] y_s, [13] z_s, [14..16] start offset, [17..19] end offset (global X Y Z; MemberEndOffset,
// pydantic_models.py:525-531, Member.start_offset / end_offset 1541 / 1518), [20] E*I_yz (Section.i_yz, 1078-1082).
enum { REL_YS = 12, REL_ZS = 13, REL_O1 = 14, REL_O2 = 17, REL_EIYZ = 20, REL_STRIDE = 21 };

// End offsets: the elastic axis runs from (start node + start offset) to (end node + end offset).
FERS_HD void rec_offset_points(const double* rec, double pa[3], double pb[3]) {
    for (int c = 0; c < 3; ++c) { pa[c] += rec[REL_O1 + c]; pb[c] += rec[REL_O2 + c]; }
}
FERS_HD bool rec_has_offsets(const double* rec) {
    bool any = false;
    for (int c = 0; c < 6; ++c) any |= (rec[REL_O1 + c] != 0.0);
    return any;
}
// Rigid arms between the nodes and the ends of the elastic axis:  u_end = u_node + theta_node x o  (o = the end's
// offset in LOCAL axes), i.e. u_axis = E u_node.  arm_fwd: u <- E u;  arm_bwd: f <- E^T f  (m_node = m_axis + o x f).
FERS_HD void arm_local(const double* rec, const Frame& f, double o[6]) {
    rot_to_local(f, rec + REL_O1, o);
    rot_to_local(f, rec + REL_O2, o + 3);
}
FERS_HD void arm_fwd(const double o[6], double* u /*12*/) {
    for (int e = 0; e < 2; ++e) {
        const double* a = o + 3 * e;
        double* t = u + 6 * e;
        const double tx = t[3], ty = t[4], tz = t[5];
        t[0] += ty * a[2] - tz * a[1];
        t[1] += tz * a[0] - tx * a[2];
        t[2] += tx * a[1] - ty * a[0];
    }
}
FERS_HD void arm_bwd(const double o[6], double* g /*12*/) {
    for (int e = 0; e < 2; ++e) {
        const double* a = o + 3 * e;
        double* t = g + 6 * e;
        const double fx = t[0], fy = t[1], fz = t[2];
        t[3] += a[1] * fz - a[2] * fy;
        t[4] += a[2] * fx - a[0] * fz;
        t[5] += a[0] * fy - a[1] * fx;
    }
}
// k <- E^T k E,  p <- E^T p  (dense local 12x12, row-major)
FERS_HD void beam_arms(double* k /*144*/, const double* rec, const Frame& f, double* p, int np) {
    if (!rec_has_offsets(rec)) return;
    double o[6];
    arm_local(rec, f, o);
    for (int i = 0; i < 12; ++i) {            // rows of k E = E^T applied to each row (k symmetric in structure only)
        double* row = k + i * 12;
        arm_bwd(o, row);
    }
    for (int j = 0; j < 12; ++j) {            // columns of E^T (k E)
        double col[12];
        for (int i = 0; i < 12; ++i) col[i] = k[i * 12 + j];
        arm_bwd(o, col);
        for (int i = 0; i < 12; ++i) k[i * 12 + j] = col[i];
    }
    for (int q = 0; q < np; ++q) arm_bwd(o, p + q * 12);
}

// Product of inertia: E I_yz int v'' w'' dx couples the two bending planes (Euler-Bernoulli).  v over (v1, rz1, v2, rz2)
// with the Hermite cubics, w over (w1, ry1, w2, ry2) with ry = -w'; equivalent loads do not depend on the stiffness.
FERS_HD void beam_product_inertia(double* k /*144*/, double eiyz, double L) {
    if (eiyz == 0.0) return;
    const double c = eiyz / (L * L * L);
    const double B[4][4] = {{12.0, 6.0 * L, -12.0, 6.0 * L}, {6.0 * L, 4.0 * L * L, -6.0 * L, 2.0 * L * L},
                            {-12.0, -6.0 * L, 12.0, -6.0 * L}, {6.0 * L, 2.0 * L * L, -6.0 * L, 4.0 * L * L}};
    const int vs[4] = {1, 5, 7, 11}, ws[4] = {2, 4, 8, 10};
    const double sw[4] = {1.0, -1.0, 1.0, -1.0};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const double x = B[i][j] * c * sw[j];
            k[vs[i] * 12 + ws[j]] += x;
            k[ws[j] * 12 + vs[i]] += x;
        }
}

// The whole record applied to the dense local stiffness in this order: product of inertia, shear-centre eccentricity,
// release condensation (the releases sit at the ends of the elastic axis), rigid arms to the nodes.
FERS_HD void beam_modify(double* k /*144*/, const double* rec /*REL_STRIDE*/, const Frame& f, double* p, int np) {
    beam_product_inertia(k, rec[REL_EIYZ], f.L);
    beam_shear_centre(k, rec[REL_YS], rec[REL_ZS], p, np);
    beam_condense(k, rec, p, np);
    beam_arms(k, rec, f, p, np);
}

// G = T^T k T for one 6x6 node block (bi, bj in {0,1}) of the 12x12 local matrix.
// out[36] row-major.  T = blockdiag(R,R,R,R).
FERS_HD void beam_rotate_block(const double* k /*144*/, const Frame& f, int bi, int bj, double* out /*36*/) {
    for (int si = 0; si < 2; ++si)
        for (int sj = 0; sj < 2; ++sj) {
            const int r0 = 6 * bi + 3 * si, c0 = 6 * bj + 3 * sj;
            double m[3][3];   // m = K_sub * R
            for (int a = 0; a < 3; ++a)
                for (int c = 0; c < 3; ++c)
                    m[a][c] = k[(r0 + a) * 12 + c0] * f.r[0][c] + k[(r0 + a) * 12 + c0 + 1] * f.r[1][c] +
                              k[(r0 + a) * 12 + c0 + 2] * f.r[2][c];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    out[(3 * si + r) * 6 + 3 * sj + c] = f.r[0][r] * m[0][c] + f.r[1][r] * m[1][c] + f.r[2][r] * m[2][c];
        }
}

// Closed-form global 6x6 block for an un-released member (no dense 12x12 needed).
// Each 3x3 sub-block is  cxx x x^T + cyy y y^T + czz z z^T + cyz y z^T + czy z y^T.
FERS_HD void beam_block_closed(const BeamCoef& c, const Frame& f, int bi, int bj, double* out /*36*/) {
    const double* x = f.r[0];
    const double* y = f.r[1];
    const double* z = f.r[2];
    const double sgn = (bi == bj) ? 1.0 : -1.0;          // translational coupling sign
    // translation-translation
    {
        double cxx = sgn * c.a, cyy = sgn * c.z12, czz = sgn * c.y12;
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) out[r * 6 + q] = cxx * x[r] * x[q] + cyy * y[r] * y[q] + czz * z[r] * z[q];
    }
    // rotation-rotation
    {
        double cxx = sgn * c.t;
        double cyy = (bi == bj) ? c.y4 : c.y2;
        double czz = (bi == bj) ? c.z4 : c.z2;
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q)
                out[(3 + r) * 6 + 3 + q] = cxx * x[r] * x[q] + cyy * y[r] * y[q] + czz * z[r] * z[q];
    }
    // translation(bi)-rotation(bj): local entries k[v_i][rz_j] and k[w_i][ry_j]
    {
        // k[v1][rz1]=z6 k[v1][rz2]=z6 k[v2][rz1]=-z6 k[v2][rz2]=-z6 ; k[w1][ry*]=-y6 k[w2][ry*]=+y6
        double cyz = (bi == 0) ? c.z6 : -c.z6;   // y (v) row, z (rz) col
        double czy = (bi == 0) ? -c.y6 : c.y6;   // z (w) row, y (ry) col
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) out[r * 6 + 3 + q] = cyz * y[r] * z[q] + czy * z[r] * y[q];
    }
    // rotation(bi)-translation(bj): transpose of the (bj, bi) translation-rotation block
    {
        double cyz = (bj == 0) ? c.z6 : -c.z6;
        double czy = (bj == 0) ? -c.y6 : c.y6;
        for (int r = 0; r < 3; ++r)
            for (int q = 0; q < 3; ++q) out[(3 + r) * 6 + q] = cyz * z[r] * y[q] + czy * y[r] * z[q];
    }
}

// f = k u for the closed-form (un-released) local stiffness.
FERS_HD void beam_k_times_u(const BeamCoef& c, const double* u /*12*/, double* f /*12*/) {
    double du = u[0] - u[6], dt = u[3] - u[9], dv = u[1] - u[7], dw = u[2] - u[8];
    f[0] = c.a * du; f[6] = -f[0];
    f[3] = c.t * dt; f[9] = -f[3];
    f[1] = c.z12 * dv + c.z6 * (u[5] + u[11]); f[7] = -f[1];
    f[5] = c.z6 * dv + c.z4 * u[5] + c.z2 * u[11];
    f[11] = c.z6 * dv + c.z2 * u[5] + c.z4 * u[11];
    f[2] = c.y12 * dw - c.y6 * (u[4] + u[10]); f[8] = -f[2];
    f[4] = -c.y6 * dw + c.y4 * u[4] + c.y2 * u[10];
    f[10] = -c.y6 * dw + c.y2 * u[4] + c.y4 * u[10];
}

// f += k_g(N) u, closed form of beam_add_kg.
FERS_HD void beam_kg_times_u(double N, double L, int consistent, int mtype, const double* u /*12*/, double* f /*12*/) {
    double dv = u[1] - u[7], dw = u[2] - u[8];
    if (!consistent || mtype == MT_TRUSS) {
        double c = N / L;
        f[1] += c * dv; f[7] -= c * dv;
        f[2] += c * dw; f[8] -= c * dw;
    } else {
        double c = N / (30.0 * L), L2 = L * L;
        double fv = c * (36.0 * dv + 3.0 * L * (u[5] + u[11]));
        f[1] += fv; f[7] -= fv;
        f[5] += c * (3.0 * L * dv + 4.0 * L2 * u[5] - L2 * u[11]);
        f[11] += c * (3.0 * L * dv - L2 * u[5] + 4.0 * L2 * u[11]);
        double fw = c * (36.0 * dw - 3.0 * L * (u[4] + u[10]));
        f[2] += fw; f[8] -= fw;
        f[4] += c * (-3.0 * L * dw + 4.0 * L2 * u[4] - L2 * u[10]);
        f[10] += c * (-3.0 * L * dw - L2 * u[4] + 4.0 * L2 * u[10]);
    }
}


// ---- Timoshenko-exact shape functions (phi = 0: Hermite cubics) ----------------------
FERS_HD void shape_n(double xi, double phi, double L, double n[4]) {
    double d = 1.0 / (1.0 + phi), x2 = xi * xi, x3 = x2 * xi;
    n[0] = (1.0 - 3.0 * x2 + 2.0 * x3 + phi * (1.0 - xi)) * d;
    n[1] = L * (xi - 2.0 * x2 + x3 + 0.5 * phi * (xi - x2)) * d;
    n[2] = (3.0 * x2 - 2.0 * x3 + phi * xi) * d;
    n[3] = L * (-x2 + x3 - 0.5 * phi * (xi - x2)) * d;
}
FERS_HD void shape_g(double xi, double phi, double L, double g[4]) {
    double d = 1.0 / (1.0 + phi), x2 = xi * xi;
    g[0] = 6.0 * (x2 - xi) * d / L;
    g[1] = (1.0 - 4.0 * xi + 3.0 * x2 + phi * (1.0 - xi)) * d;
    g[2] = -g[0];
    g[3] = (3.0 * x2 - 2.0 * xi + phi * xi) * d;
}

// Equivalent nodal loads (local, 12) of a linearly varying load (global vectors q1 at fraction a,
// q2 at fraction b) -- pydantic_models.py:1438-1445.  Accumulates into p.
FERS_HD void beam_dist_load(const Frame& f, const BeamCoef& c, int mtype, const double* row /*q1[3] q2[3] a b*/,
                            double* p /*12*/) {
    const double gx[3] = {-0.7745966692414834, 0.0, 0.7745966692414834};
    const double gw[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    double a = row[6], b = row[7];
    double half = 0.5 * (b - a), mid = 0.5 * (a + b);
    if (!(half > 0.0)) return;
    double q1[3], q2[3];
    for (int i = 0; i < 3; ++i) {
        q1[i] = f.r[i][0] * row[0] + f.r[i][1] * row[1] + f.r[i][2] * row[2];
        q2[i] = f.r[i][0] * row[3] + f.r[i][1] * row[4] + f.r[i][2] * row[5];
    }
    bool truss = (mtype == MT_TRUSS);
    for (int g = 0; g < 3; ++g) {
        double xi = mid + half * gx[g];
        double s = (xi - a) / (b - a);
        double w = gw[g] * half * f.L;
        double qx = (q1[0] + (q2[0] - q1[0]) * s) * w;
        double qy = (q1[1] + (q2[1] - q1[1]) * s) * w;
        double qz = (q1[2] + (q2[2] - q1[2]) * s) * w;
        p[0] += (1.0 - xi) * qx; p[6] += xi * qx;
        if (truss) {
            p[1] += (1.0 - xi) * qy; p[7] += xi * qy;
            p[2] += (1.0 - xi) * qz; p[8] += xi * qz;
        } else {
            double n[4];
            shape_n(xi, c.phi_y, f.L, n);
            p[1] += n[0] * qy; p[5] += n[1] * qy; p[7] += n[2] * qy; p[11] += n[3] * qy;
            shape_n(xi, c.phi_z, f.L, n);
            p[2] += n[0] * qz; p[4] -= n[1] * qz; p[8] += n[2] * qz; p[10] -= n[3] * qz;
        }
    }
}

// Equivalent nodal loads of a member point load / moment (pydantic_models.py:1562-1591).
// row = F[3] M[3] position is_local.
FERS_HD void beam_point_load(const Frame& f, const BeamCoef& c, int mtype, const double* row, double* p /*12*/) {
    double F[3], M[3];
    if (row[7] != 0.0) {
        for (int i = 0; i < 3; ++i) { F[i] = row[i]; M[i] = row[3 + i]; }
    } else {
        for (int i = 0; i < 3; ++i) {
            F[i] = f.r[i][0] * row[0] + f.r[i][1] * row[1] + f.r[i][2] * row[2];
            M[i] = f.r[i][0] * row[3] + f.r[i][1] * row[4] + f.r[i][2] * row[5];
        }
    }
    double xi = row[6];
    p[0] += (1.0 - xi) * F[0]; p[6] += xi * F[0];
    p[3] += (1.0 - xi) * M[0]; p[9] += xi * M[0];
    if (mtype == MT_TRUSS) {
        p[1] += (1.0 - xi) * F[1]; p[7] += xi * F[1];
        p[2] += (1.0 - xi) * F[2]; p[8] += xi * F[2];
        return;
    }
    double n[4], g[4];
    shape_n(xi, c.phi_y, f.L, n); shape_g(xi, c.phi_y, f.L, g);
    p[1] += n[0] * F[1] + g[0] * M[2]; p[5] += n[1] * F[1] + g[1] * M[2];
    p[7] += n[2] * F[1] + g[2] * M[2]; p[11] += n[3] * F[1] + g[3] * M[2];
    shape_n(xi, c.phi_z, f.L, n); shape_g(xi, c.phi_z, f.L, g);
    p[2] += n[0] * F[2] - g[0] * M[1]; p[4] += -n[1] * F[2] + g[1] * M[1];
    p[8] += n[2] * F[2] - g[2] * M[1]; p[10] += -n[3] * F[2] + g[3] * M[1];
}

// Resultant of a distributed load on [0, x] about the section at fraction x (local frame):
// adds force into s[0..2] and moment into s[3..5], scaled by `fac`.
FERS_HD void span_dist_resultant(const Frame& f, const double* row, double x, double fac, double* s /*6*/) {
    const double gx[3] = {-0.7745966692414834, 0.0, 0.7745966692414834};
    const double gw[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    double a = row[6], b = row[7];
    double hi = b < x ? b : x;
    if (!(hi > a)) return;
    double q1[3], q2[3];
    for (int i = 0; i < 3; ++i) {
        q1[i] = f.r[i][0] * row[0] + f.r[i][1] * row[1] + f.r[i][2] * row[2];
        q2[i] = f.r[i][0] * row[3] + f.r[i][1] * row[4] + f.r[i][2] * row[5];
    }
    double half = 0.5 * (hi - a), mid = 0.5 * (hi + a);
    for (int g = 0; g < 3; ++g) {
        double xi = mid + half * gx[g];
        double t = (xi - a) / (b - a);
        double w = fac * gw[g] * half * f.L;
        double d = (x - xi) * f.L;
        double qx = q1[0] + (q2[0] - q1[0]) * t, qy = q1[1] + (q2[1] - q1[1]) * t, qz = q1[2] + (q2[2] - q1[2]) * t;
        s[0] += w * qx; s[1] += w * qy; s[2] += w * qz;
        s[4] += w * d * qz; s[5] -= w * d * qy;
    }
}

FERS_HD void span_point_resultant(const Frame& f, const double* row, double x, double fac, double* s /*6*/) {
    double pos = row[6];
    if (pos > x) return;
    double F[3], M[3];
    if (row[7] != 0.0) {
        for (int i = 0; i < 3; ++i) { F[i] = row[i]; M[i] = row[3 + i]; }
    } else {
        for (int i = 0; i < 3; ++i) {
            F[i] = f.r[i][0] * row[0] + f.r[i][1] * row[1] + f.r[i][2] * row[2];
            M[i] = f.r[i][0] * row[3] + f.r[i][1] * row[4] + f.r[i][2] * row[5];
        }
    }
    double d = (x - pos) * f.L;
    s[0] += fac * F[0]; s[1] += fac * F[1]; s[2] += fac * F[2];
    s[3] += fac * M[0]; s[4] += fac * (M[1] + d * F[2]); s[5] += fac * (M[2] - d * F[1]);
}

// ---- TRI3 shell: CST membrane + DSG3 Mindlin bending --------
// Per-plate record (NPDATA doubles): thickness, E, nu, then the stiffness modifiers of PlateElement.stiffness_modifiers
// (membrane, bending, shear; pydantic_models.py:801-807).  flags: bit0 membrane, bit1 bending, bit2 plane strain
// (PlateElement.plane_state, 1712-1715: in-plane constitutive law of the membrane part).
// Orthotropic plate material (OrthotropicPlateMaterial, pydantic_models.py:663-681; material axes = the plate's local x, y):
// PD_ORTHO != 0 -> E_x, E_y, G_xy, nu_xy (nu_yx from E_x nu_yx = E_y nu_xy) for membrane and bending, G_xz / G_yz for the
// transverse shear of the Mindlin part.
enum { PD_T = 0, PD_E = 1, PD_NU = 2, PD_MODM = 3, PD_MODB = 4, PD_MODS = 5, PD_THETA = 6, PD_ORTHO = 7,
       PD_EX = 8, PD_EY = 9, PD_GXY = 10, PD_NUXY = 11, PD_GXZ = 12, PD_GYZ = 13, PD_OFFSET = 14, NPDATA = 16 };

// Plate offset (PlateElement.offset, pydantic_models.py:1711-1714: "distance from the reference surface along the local
// plate normal"): the mid-surface sits at local z = o from the nodes, every node carries it on a rigid arm,
//   u_mid = u + o ry,  v_mid = v - o rx   (local axes; w and the rotations are shared),   u_mid = E u_node.
// plate_offset_fwd: local nodal displacements -> mid-surface displacements (18 values in place); the stiffness is E^T k E.
FERS_HD void plate_offset_fwd(double o, double* u /*18*/) {
    if (o == 0.0) return;
    for (int n = 0; n < 3; ++n) { u[6 * n] += o * u[6 * n + 4]; u[6 * n + 1] -= o * u[6 * n + 3]; }
}
FERS_HD void plate_offset_k(double o, double* k /*324*/) {
    if (o == 0.0) return;
    for (int i = 0; i < 18; ++i)              // k E: columns of ry take o x the u column, columns of rx take -o x the v column
        for (int n = 0; n < 3; ++n) {
            k[i * 18 + 6 * n + 4] += o * k[i * 18 + 6 * n];
            k[i * 18 + 6 * n + 3] -= o * k[i * 18 + 6 * n + 1];
        }
    for (int j = 0; j < 18; ++j)              // E^T (k E): the same on the rows
        for (int n = 0; n < 3; ++n) {
            k[(6 * n + 4) * 18 + j] += o * k[(6 * n) * 18 + j];
            k[(6 * n + 3) * 18 + j] -= o * k[(6 * n + 1) * 18 + j];
        }
}

// in-plane constitutive matrix: plane stress, or plane strain when asked for (membrane part only)
FERS_HD void plate_D(double E, double nu, bool plane_strain, double D[3][3]) {
    double d11, d12, d33;
    if (plane_strain) {
        const double f = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
        d11 = f * (1.0 - nu); d12 = f * nu; d33 = f * (1.0 - 2.0 * nu) / 2.0;
    } else {
        const double f = E / (1.0 - nu * nu);
        d11 = f; d12 = f * nu; d33 = f * (1.0 - nu) / 2.0;
    }
    D[0][0] = d11; D[0][1] = d12; D[0][2] = 0.0;
    D[1][0] = d12; D[1][1] = d11; D[1][2] = 0.0;
    D[2][0] = 0.0; D[2][1] = 0.0; D[2][2] = d33;
}
// membrane (Dm) and bending (Db) in-plane constitutive matrices and the two transverse shear moduli of one plate record
FERS_HD void plate_constitutive(const double* data, int flags, double Dm[3][3], double Db[3][3], double Gs[2]) {
    if (data[PD_ORTHO] != 0.0) {
        const double ex = data[PD_EX], ey = data[PD_EY], nxy = data[PD_NUXY];
        const double nyx = nxy * ey / ex, den = 1.0 - nxy * nyx;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Dm[i][j] = 0.0;
        Dm[0][0] = ex / den; Dm[1][1] = ey / den; Dm[0][1] = Dm[1][0] = nxy * ey / den; Dm[2][2] = data[PD_GXY];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Db[i][j] = Dm[i][j];
        Gs[0] = data[PD_GXZ]; Gs[1] = data[PD_GYZ];
    } else {
        plate_D(data[PD_E], data[PD_NU], (flags & 4) != 0, Dm);      // membrane: plane stress or plane strain
        plate_D(data[PD_E], data[PD_NU], false, Db);                 // bending: plate theory (plane stress)
        Gs[0] = Gs[1] = data[PD_E] / (2.0 * (1.0 + data[PD_NU]));
    }
}
struct PlateGeom {
    double r[3][3];     // rows = local x, y, z
    double xl[3][2];    // local corner coordinates (corner 0 at origin)
    double A;           // signed area
};

// theta: in-plane angle of local x from the default frame (PlateElement.local_x_direction, pydantic_models.py:1703,
// resolved on the host by wire.plate_local_x_angle); 0 keeps the default frame bit for bit.
FERS_HD PlateGeom plate_geom(const double p0[3], const double p1[3], const double p2[3], double theta = 0.0) {
    PlateGeom g;
    double e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    double e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    double nn = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    double ez[3] = {n[0] * nn, n[1] * nn, n[2] * nn};
    double gx[3] = {1.0, 0.0, 0.0};
    if (fabs(ez[0]) > 1.0 - 1e-6) { gx[0] = 0.0; gx[1] = 1.0; }
    double dot = gx[0] * ez[0] + gx[1] * ez[1] + gx[2] * ez[2];
    double ex[3] = {gx[0] - dot * ez[0], gx[1] - dot * ez[1], gx[2] - dot * ez[2]};
    double en = 1.0 / sqrt(ex[0] * ex[0] + ex[1] * ex[1] + ex[2] * ex[2]);
    ex[0] *= en; ex[1] *= en; ex[2] *= en;
    double ey[3] = {ez[1] * ex[2] - ez[2] * ex[1], ez[2] * ex[0] - ez[0] * ex[2], ez[0] * ex[1] - ez[1] * ex[0]};
    if (theta != 0.0) {
        const double ct = cos(theta), st = sin(theta);
        for (int c = 0; c < 3; ++c) {
            const double x0 = ex[c], y0 = ey[c];
            ex[c] = ct * x0 + st * y0;
            ey[c] = -st * x0 + ct * y0;
        }
    }
    for (int c = 0; c < 3; ++c) { g.r[0][c] = ex[c]; g.r[1][c] = ey[c]; g.r[2][c] = ez[c]; }
    g.xl[0][0] = 0.0; g.xl[0][1] = 0.0;
    g.xl[1][0] = e1[0] * ex[0] + e1[1] * ex[1] + e1[2] * ex[2];
    g.xl[1][1] = e1[0] * ey[0] + e1[1] * ey[1] + e1[2] * ey[2];
    g.xl[2][0] = e2[0] * ex[0] + e2[1] * ex[1] + e2[2] * ex[2];
    g.xl[2][1] = e2[0] * ey[0] + e2[1] * ey[1] + e2[2] * ey[2];
    double a_ = g.xl[1][0], b_ = g.xl[1][1], c_ = g.xl[2][1], d_ = g.xl[2][0];
    g.A = 0.5 * (a_ * c_ - b_ * d_);
    return g;
}

// Strain-displacement rows in LOCAL shell DOFs (18 = 3 nodes x [u v w rx ry rz]).
// Bm[3][18] membrane, Bb[3][18] curvature, Bs[2][18] transverse shear (DSG3, bx=+ry, by=-rx).
FERS_HD void plate_B(const PlateGeom& g, double Bm[3][18], double Bb[3][18], double Bs[2][18]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 18; ++j) { Bm[i][j] = 0.0; Bb[i][j] = 0.0; if (i < 2) Bs[i][j] = 0.0; }
    double x1 = g.xl[0][0], y1 = g.xl[0][1], x2 = g.xl[1][0], y2 = g.xl[1][1], x3 = g.xl[2][0], y3 = g.xl[2][1];
    double A = g.A, i2A = 1.0 / (2.0 * A);
    double bb[3] = {y2 - y3, y3 - y1, y1 - y2};
    double cc[3] = {x3 - x2, x1 - x3, x2 - x1};
    for (int n = 0; n < 3; ++n) {
        Bm[0][6 * n + 0] = bb[n] * i2A;
        Bm[1][6 * n + 1] = cc[n] * i2A;
        Bm[2][6 * n + 0] = cc[n] * i2A;
        Bm[2][6 * n + 1] = bb[n] * i2A;
    }
    double a_ = x2 - x1, b_ = y2 - y1, c_ = y3 - y1, d_ = x3 - x1;
    // columns in (w, bx, by) per node
    double bbm[3][9] = {{0, b_ - c_, 0, 0, c_, 0, 0, -b_, 0},
                        {0, 0, d_ - a_, 0, 0, -d_, 0, 0, a_},
                        {0, d_ - a_, b_ - c_, 0, -d_, c_, 0, a_, -b_}};
    double bsm[2][9] = {{b_ - c_, A, 0, c_, a_ * c_ / 2, b_ * c_ / 2, -b_, -b_ * d_ / 2, -b_ * c_ / 2},
                        {d_ - a_, 0, A, -d_, -a_ * d_ / 2, -b_ * d_ / 2, a_, a_ * d_ / 2, a_ * c_ / 2}};
    for (int n = 0; n < 3; ++n) {
        for (int i = 0; i < 3; ++i) {
            Bb[i][6 * n + 2] = bbm[i][3 * n + 0] * i2A;      // w
            Bb[i][6 * n + 4] = bbm[i][3 * n + 1] * i2A;      // bx = +ry
            Bb[i][6 * n + 3] = -bbm[i][3 * n + 2] * i2A;     // by = -rx
        }
        for (int i = 0; i < 2; ++i) {
            Bs[i][6 * n + 2] = bsm[i][3 * n + 0] * i2A;
            Bs[i][6 * n + 4] = bsm[i][3 * n + 1] * i2A;
            Bs[i][6 * n + 3] = -bsm[i][3 * n + 2] * i2A;
        }
    }
}

// ---- thin plates: discrete Kirchhoff triangle (PlateTheory.Kirchhoff, pydantic_models.py:816-819, field 1723) -------------
// Batoz, Bathe, Ho (1980): the rotations of the normal bx = +ry, by = -rx are quadratic over the element, tied to w by the
// Kirchhoff condition at the corners and, along each side, at its mid-point (w cubic along the side); no transverse shear
// energy.  Curvature rows kappa = [bx,x  by,y  bx,y + by,x] at the area coordinates (xi, eta) = (L2, L3), in the local shell
// DOFs (columns w, rx, ry of every node; the others stay zero): the explicit derivatives of the paper's Hx, Hy.
FERS_HD void plate_dkt_B(const PlateGeom& g, double xi, double eta, double Bb[3][18]) {
    const double x1 = g.xl[0][0], y1 = g.xl[0][1], x2 = g.xl[1][0], y2 = g.xl[1][1], x3 = g.xl[2][0], y3 = g.xl[2][1];
    const double x23 = x2 - x3, x31 = x3 - x1, x12 = x1 - x2, y23 = y2 - y3, y31 = y3 - y1, y12 = y1 - y2;
    const double i4 = 1.0 / (x23 * x23 + y23 * y23), i5 = 1.0 / (x31 * x31 + y31 * y31), i6 = 1.0 / (x12 * x12 + y12 * y12);
    const double P4 = -6.0 * x23 * i4, P5 = -6.0 * x31 * i5, P6 = -6.0 * x12 * i6;
    const double t4 = -6.0 * y23 * i4, t5 = -6.0 * y31 * i5, t6 = -6.0 * y12 * i6;
    const double q4 = 3.0 * x23 * y23 * i4, q5 = 3.0 * x31 * y31 * i5, q6 = 3.0 * x12 * y12 * i6;
    const double r4 = 3.0 * y23 * y23 * i4, r5 = 3.0 * y31 * y31 * i5, r6 = 3.0 * y12 * y12 * i6;
    const double a = 1.0 - 2.0 * xi, b = 1.0 - 2.0 * eta, x = xi, e = eta;
    const double hxx[9] = {P6 * a + (P5 - P6) * e, q6 * a - (q5 + q6) * e, -4.0 + 6.0 * (x + e) + r6 * a - e * (r5 + r6),
                           -P6 * a + e * (P4 + P6), q6 * a - e * (q6 - q4), -2.0 + 6.0 * x + r6 * a + e * (r4 - r6),
                           -e * (P5 + P4), e * (q4 - q5), -e * (r5 - r4)};
    const double hyx[9] = {t6 * a + e * (t5 - t6), 1.0 + r6 * a - e * (r5 + r6), -q6 * a + e * (q5 + q6),
                           -t6 * a + e * (t4 + t6), -1.0 + r6 * a + e * (r4 - r6), -q6 * a - e * (q4 - q6),
                           -e * (t4 + t5), e * (r4 - r5), -e * (q4 - q5)};
    const double hxe[9] = {-P5 * b - x * (P6 - P5), q5 * b - x * (q5 + q6), -4.0 + 6.0 * (x + e) + r5 * b - x * (r5 + r6),
                           x * (P4 + P6), x * (q4 - q6), -x * (r6 - r4),
                           P5 * b - x * (P4 + P5), q5 * b + x * (q4 - q5), -2.0 + 6.0 * e + r5 * b + x * (r4 - r5)};
    const double hye[9] = {-t5 * b - x * (t6 - t5), 1.0 + r5 * b - x * (r5 + r6), -q5 * b + x * (q5 + q6),
                           x * (t4 + t6), x * (r4 - r6), -x * (q4 - q6),
                           t5 * b - x * (t4 + t5), -1.0 + r5 * b + x * (r4 - r5), -q5 * b - x * (q4 - q5)};
    const double i2A = 1.0 / (x31 * y12 - x12 * y31);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 18; ++j) Bb[i][j] = 0.0;
    for (int n = 0; n < 3; ++n)
        for (int c = 0; c < 3; ++c) {
            const int j = 3 * n + c, col = 6 * n + 2 + c;       // (w, theta_x, theta_y) = local DOFs (w, rx, ry)
            Bb[0][col] = (y31 * hxx[j] + y12 * hxe[j]) * i2A;
            Bb[1][col] = (-x31 * hyx[j] - x12 * hye[j]) * i2A;
            Bb[2][col] = (-x31 * hxx[j] - x12 * hxe[j] + y31 * hyx[j] + y12 * hye[j]) * i2A;
        }
}

// Strain operators the results are formed from: Mindlin -> plate_B.  Kirchhoff -> curvature of the DKT element at the
// centroid and NO shear rows: a thin plate has no transverse shear strain, its shear forces are derivatives of the moments,
// and differentiating the curvature inside one DKT element is not a usable estimate (measured on a cantilever strip: 30 %
// off at every mesh size; the nodal-force form has exact resultants but oscillates by several times its mean) -- the
// result carries zeros and the analysis says so (W_KIRCHHOFF_NO_SHEAR_FORCES).
FERS_HD bool plate_is_kirchhoff(int flags) { return (flags & 8) != 0 && (flags & 2) != 0; }
FERS_HD void plate_B_theory(const PlateGeom& g, int flags, double Bm[3][18], double Bb[3][18], double Bs[2][18]) {
    plate_B(g, Bm, Bb, Bs);
    if (!plate_is_kirchhoff(flags)) return;
    plate_dkt_B(g, 1.0 / 3.0, 1.0 / 3.0, Bb);
    for (int j = 0; j < 18; ++j) Bs[0][j] = Bs[1][j] = 0.0;
}

// 18x18 local stiffness (row-major).  data = plate record (NPDATA doubles); flags bit0 membrane bit1 bending bit2 plane strain
// bit3 Kirchhoff (DKT bending, three mid-side integration points, no shear part).
FERS_HD void plate_local_k(const PlateGeom& g, const double* data, int flags, double* k /*324*/) {
    double Bm[3][18], Bb[3][18], Bs[2][18];
    plate_B(g, Bm, Bb, Bs);
    double t = data[PD_T], E = data[PD_E];
    double A = fabs(g.A);
    double Dm[3][3], D[3][3], Gs[2];
    plate_constitutive(data, flags, Dm, D, Gs);
    double cm = (flags & 1) ? data[PD_MODM] * t * A : 0.0;
    double cb = (flags & 2) ? data[PD_MODB] * A * t * t * t / 12.0 : 0.0;
    double cs = (flags & 2) ? data[PD_MODS] * A * (5.0 / 6.0) * t : 0.0;
    const bool dkt = plate_is_kirchhoff(flags);
    if (dkt) cs = 0.0;
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) {
            double sm = 0.0, sb = 0.0;
            for (int a = 0; a < 3; ++a) {
                double dmj = Dm[a][0] * Bm[0][j] + Dm[a][1] * Bm[1][j] + Dm[a][2] * Bm[2][j];
                double dbj = D[a][0] * Bb[0][j] + D[a][1] * Bb[1][j] + D[a][2] * Bb[2][j];
                sm += Bm[a][i] * dmj;
                sb += Bb[a][i] * dbj;
            }
            double ss = Gs[0] * Bs[0][i] * Bs[0][j] + Gs[1] * Bs[1][i] * Bs[1][j];
            k[i * 18 + j] = cm * sm + (dkt ? 0.0 : cb * sb) + cs * ss;
        }
    if (dkt) {
        const double gp[3][2] = {{0.5, 0.5}, {0.0, 0.5}, {0.5, 0.0}};
        for (int p = 0; p < 3; ++p) {
            plate_dkt_B(g, gp[p][0], gp[p][1], Bb);
            for (int j = 0; j < 18; ++j) {
                double db[3];
                for (int a = 0; a < 3; ++a) db[a] = D[a][0] * Bb[0][j] + D[a][1] * Bb[1][j] + D[a][2] * Bb[2][j];
                for (int i = 0; i < 18; ++i)
                    k[i * 18 + j] += (cb / 3.0) * (Bb[0][i] * db[0] + Bb[1][i] * db[1] + Bb[2][i] * db[2]);
            }
        }
    }
    double kd = 1e-6 * E * t * A;     // drilling stabilisation
    for (int n = 0; n < 3; ++n) k[(6 * n + 5) * 18 + 6 * n + 5] += kd;
    plate_offset_k(data[PD_OFFSET], k);
}

// Global 6x6 block (node bi, node bj) of T^T k T for a plate.
FERS_HD void plate_rotate_block(const double* k /*324*/, const PlateGeom& g, int bi, int bj, double* out /*36*/) {
    for (int si = 0; si < 2; ++si)
        for (int sj = 0; sj < 2; ++sj) {
            const int r0 = 6 * bi + 3 * si, c0 = 6 * bj + 3 * sj;
            double m[3][3];
            for (int a = 0; a < 3; ++a)
                for (int c = 0; c < 3; ++c)
                    m[a][c] = k[(r0 + a) * 18 + c0] * g.r[0][c] + k[(r0 + a) * 18 + c0 + 1] * g.r[1][c] +
                              k[(r0 + a) * 18 + c0 + 2] * g.r[2][c];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    out[(3 * si + r) * 6 + 3 * sj + c] = g.r[0][r] * m[0][c] + g.r[1][r] * m[1][c] + g.r[2][r] * m[2][c];
        }
}

}  // namespace fers
