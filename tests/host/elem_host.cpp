// Host instantiation of fers_elem.cuh for CPU-side unit tests (tests/test_elem_host.py).
// TEST HARNESS ONLY: the product never links this file.
#include "../../fers_core_b200/csrc/fers_elem.cuh"
using namespace fers;

extern "C" {
// global 12x12 via the closed-form block path (mode 0) or dense local k + rotation (mode 1)
void h_beam_ke(const double* a, const double* b, double roll, const double* P, int mtype, double scale,
               const double* rel /*REL_STRIDE or null*/, int mode, double* out /*144*/) {
    double pa[3] = {a[0], a[1], a[2]}, pb[3] = {b[0], b[1], b[2]};
    if (rel) rec_offset_points(rel, pa, pb);
    Frame f = member_frame(pa, pb, roll);
    BeamCoef c = beam_coef(P, f.L, mtype, scale);
    double blk[36];
    double k[144];
    if (mode == 1) {
        beam_local_k(c, k);
        if (rel) beam_modify(k, rel, f, nullptr, 0);
    }
    for (int bi = 0; bi < 2; ++bi)
        for (int bj = 0; bj < 2; ++bj) {
            if (mode == 0) beam_block_closed(c, f, bi, bj, blk);
            else beam_rotate_block(k, f, bi, bj, blk);
            for (int r = 0; r < 6; ++r)
                for (int q = 0; q < 6; ++q) out[(6 * bi + r) * 12 + 6 * bj + q] = blk[r * 6 + q];
        }
}
void h_beam_local(const double* a, const double* b, double roll, const double* P, int mtype, double scale,
                  double N, int consistent, double* k /*144*/, double* R /*9*/, double* L) {
    Frame f = member_frame(a, b, roll);
    BeamCoef c = beam_coef(P, f.L, mtype, scale);
    beam_local_k(c, k);
    if (N != 0.0) beam_add_kg(k, N, f.L, consistent, mtype);
    for (int i = 0; i < 9; ++i) R[i] = f.r[i / 3][i % 3];
    *L = f.L;
}
void h_beam_loads(const double* a, const double* b, double roll, const double* P, int mtype,
                  const double* dl_rows, int ndl, const double* pl_rows, int npl, const double* rel, double* p /*12*/) {
    double pa[3] = {a[0], a[1], a[2]}, pb[3] = {b[0], b[1], b[2]};
    if (rel) rec_offset_points(rel, pa, pb);
    Frame f = member_frame(pa, pb, roll);
    BeamCoef c = beam_coef(P, f.L, mtype, 1.0);
    for (int i = 0; i < 12; ++i) p[i] = 0.0;
    for (int i = 0; i < ndl; ++i) beam_dist_load(f, c, mtype, dl_rows + 8 * i, p);
    for (int i = 0; i < npl; ++i) beam_point_load(f, c, mtype, pl_rows + 8 * i, p);
    if (rel) {
        double k[144];
        beam_local_k(c, k);
        beam_modify(k, rel, f, p, 1);
    }
}
void h_span(const double* a, const double* b, double roll, const double* dl_rows, int ndl, const double* pl_rows,
            int npl, double x, double* s /*6*/) {
    Frame f = member_frame(a, b, roll);
    for (int i = 0; i < 6; ++i) s[i] = 0.0;
    for (int i = 0; i < ndl; ++i) span_dist_resultant(f, dl_rows + 8 * i, x, 1.0, s);
    for (int i = 0; i < npl; ++i) span_point_resultant(f, pl_rows + 8 * i, x, 1.0, s);
}
void h_plate_ke(const double* p0, const double* p1, const double* p2, const double* data, int flags, double* out /*324*/) {
    PlateGeom g = plate_geom(p0, p1, p2, data[PD_THETA]);
    double k[324], blk[36];
    plate_local_k(g, data, flags, k);
    for (int bi = 0; bi < 3; ++bi)
        for (int bj = 0; bj < 3; ++bj) {
            plate_rotate_block(k, g, bi, bj, blk);
            for (int r = 0; r < 6; ++r)
                for (int q = 0; q < 6; ++q) out[(6 * bi + r) * 18 + 6 * bj + q] = blk[r * 6 + q];
        }
}
// strain operators of the plate results (rows: 3 membrane, 3 curvature, 2 shear), local DOFs
void h_plate_B(const double* p0, const double* p1, const double* p2, const double* data, int flags, double* out /*8*18*/) {
    PlateGeom g = plate_geom(p0, p1, p2, data[PD_THETA]);
    double Bm[3][18], Bb[3][18], Bs[2][18];
    plate_B_theory(g, flags, Bm, Bb, Bs);
    for (int j = 0; j < 18; ++j) {
        for (int i = 0; i < 3; ++i) { out[i * 18 + j] = Bm[i][j]; out[(3 + i) * 18 + j] = Bb[i][j]; }
        out[6 * 18 + j] = Bs[0][j]; out[7 * 18 + j] = Bs[1][j];
    }
}
}
