// unity.cu -- unity checks: evaluation of a compiled "Generic" check
// (fers_core/types/pydantic_models.py: GenericSpec 2136-2151, UnityCheckDefinition 2202-2220, quantity
// sources 829-969) for every (entity, right-hand-side column) and the envelope over the columns.
//
// The host (fers_core_b200/unity.py) compiles the check into
//   - a table of variables (kind + three integers each): per-entity constants (section / material /
//     geometry / literal), member-force aggregations read from the recovery arrays, node displacement /
//     deflected position, chord-relative member deflection, or a sub-expression,
//   - postfix programs (one per expression variable, one for the demand, one for the capacity).
// One warp owns one entity; lanes stride over the columns, evaluate utilisation = demand / capacity and
// the warp keeps the column with the largest utilisation (ties: lowest column) -- the "enveloped over
// the applicable load combinations" of the schema.  Everything is read in the [..][nrhs] layout the
// recovery kernels write, so a warp reads contiguous runs of right-hand-side values.
#include <math.h>

#include "common.cuh"

namespace fers {

enum UnityOp { U_PUSHC = 0, U_PUSHV, U_ADD, U_SUB, U_MUL, U_DIV, U_POW, U_NEG, U_ABS, U_SQRT, U_MIN, U_MAX };
enum UnityVar { V_CONST = 0, V_FORCE, V_EXPR, V_NODE_DISP, V_NODE_POS, V_DEFLECTION, V_FORCE_REF, V_ENTITY_DISP, V_PLATE };
enum UnityAgg { A_MAXABS = 0, A_MAX, A_MIN, A_START, A_END, A_STATION };

constexpr int UNITY_MAX_VARS = 96;       // the EN 1993-1-1 program (fers_core_b200/ec3.py) binds about 70 names
constexpr int UNITY_STACK = 16;

struct UnityArgs {
    int n_ent, nrhs, n_vars, n_ct;
    const uint8_t* col_mask;     // [nrhs] 1 = column takes part in the envelope
    const int* ent_ptr;          // [n_ent+1] into ent_members
    const int* ent_members;      // member indices, chain order
    const int* ent_node_ptr;     // [n_ent+1] into ent_nodes
    const int* ent_nodes;        // node indices along the entity (start .. end), chain order
    const int* var_desc;         // [n_vars][4]: kind, a, b, c
    const int* code;             // postfix programs: pairs (op, arg)
    const double* consts;
    const double* ctab;          // [n_ent][n_ct]
    int demand_off, demand_len, cap_off, cap_len;
    const double* xyz;           // [n_nodes][3]
    const double* U;             // [n_dof][nrhs]
    const double* f_local;       // [n_members][12][nrhs]
    const double* sec;           // [n_members][18][nrhs]
    const double* lmax;          // [n_members][6][nrhs]
    const double* lmin;
    const double* frames;        // [n_ent][9] local axes of the entity's first member (deflection components)
    const double* p_res;         // [n_plates][8][nrhs] plate resultants nx ny nxy mx my mxy qx qy (plate entities), or NULL
    const double* p_data;        // [n_plates][NPDATA] plate records (thickness first)
    const int* ent_ref;          // [n_ent][4]: reference member, its start node, its end node, reference node (-1: none)
    double* out;                 // [n_ent][4]: utilisation, demand, capacity, column (as double; -1 none)
    double* vars_out;            // [n_ent][n_vars] variable values at the governing column, or NULL
};

__device__ double unity_run(const int* __restrict__ code, int off, int len, const double* __restrict__ consts,
                            const double* vals) {
    double st[UNITY_STACK];
    int sp = 0;
    for (int i = 0; i < len; ++i) {
        const int op = code[2 * (off + i)], arg = code[2 * (off + i) + 1];
        switch (op) {
            case U_PUSHC: st[sp++] = consts[arg]; break;
            case U_PUSHV: st[sp++] = vals[arg]; break;
            case U_ADD: --sp; st[sp - 1] = st[sp - 1] + st[sp]; break;
            case U_SUB: --sp; st[sp - 1] = st[sp - 1] - st[sp]; break;
            case U_MUL: --sp; st[sp - 1] = st[sp - 1] * st[sp]; break;
            case U_DIV: --sp; st[sp - 1] = st[sp - 1] / st[sp]; break;
            case U_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
            case U_NEG: st[sp - 1] = -st[sp - 1]; break;
            case U_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
            case U_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
            case U_MIN: --sp; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
            case U_MAX: --sp; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
            default: break;
        }
    }
    return sp > 0 ? st[sp - 1] : 0.0;
}

// all variables of the check for one (entity, column)
__device__ void unity_eval_vars(const UnityArgs& A, int ent, int col, double* vals) {
    const int nr = A.nrhs;
    const int mb = A.ent_ptr[ent], me = A.ent_ptr[ent + 1];
    const int nb = A.ent_node_ptr[ent], ne = A.ent_node_ptr[ent + 1];
    const int* ref = A.ent_ref + (size_t)ent * 4;
    for (int v = 0; v < A.n_vars; ++v) {
        const int kind = A.var_desc[4 * v], a = A.var_desc[4 * v + 1], b = A.var_desc[4 * v + 2], c = A.var_desc[4 * v + 3];
        double x = 0.0;
        if (kind == V_CONST) {
            x = A.ctab[(size_t)ent * A.n_ct + a];
        } else if (kind == V_FORCE || kind == V_FORCE_REF) {   // a = component 0..5, b = aggregation, c = station
            // the entity's own member chain, or the single member it refers to (`of: ReferenceMember`)
            const bool own = kind == V_FORCE;
            const int first = own ? A.ent_members[mb] : ref[0], last = own ? A.ent_members[me - 1] : ref[0];
            if (first < 0) {
                x = NAN;                                        // no reference member
            } else if (b == A_START) {
                x = A.f_local[((size_t)first * 12 + a) * nr + col];
            } else if (b == A_END) {
                x = A.f_local[((size_t)last * 12 + 6 + a) * nr + col];
            } else if (b == A_STATION) {                        // single members: c = 0 start, 1..3 quarter points, 4 end
                const size_t m = first;
                if (c == 0) x = A.f_local[(m * 12 + a) * nr + col];
                else if (c == 4) x = -A.f_local[(m * 12 + 6 + a) * nr + col];
                else x = A.sec[(m * 18 + (size_t)(c - 1) * 6 + a) * nr + col];
            } else {
                double hi = -INFINITY, lo = INFINITY;
                const int q0 = own ? mb : 0, q1 = own ? me : 1;
                for (int q = q0; q < q1; ++q) {
                    const size_t m = own ? A.ent_members[q] : first;
                    hi = fmax(hi, A.lmax[(m * 6 + a) * nr + col]);
                    lo = fmin(lo, A.lmin[(m * 6 + a) * nr + col]);
                }
                x = (b == A_MAX) ? hi : (b == A_MIN) ? lo : fmax(fabs(hi), fabs(lo));
            }
        } else if (kind == V_EXPR) {                            // a = program offset, b = length
            x = unity_run(A.code, a, b, A.consts, vals);
        } else if (kind == V_NODE_DISP || kind == V_NODE_POS) {
            // a = 0 start, 1 end of the entity; 2 / 3 start / end of the reference member; 4 reference node; b = component
            const int node_i = a == 0 ? A.ent_nodes[nb] : a == 1 ? A.ent_nodes[ne - 1] : ref[a - 1];
            if (node_i < 0) {
                x = NAN;
            } else {
                const size_t node = node_i;
                if (b < 6) {
                    x = A.U[(node * 6 + b) * nr + col];
                    if (kind == V_NODE_POS) x += A.xyz[node * 3 + b];
                } else {                                        // translation magnitude
                    const double ux = A.U[(node * 6) * nr + col], uy = A.U[(node * 6 + 1) * nr + col], uz = A.U[(node * 6 + 2) * nr + col];
                    x = sqrt(ux * ux + uy * uy + uz * uz);
                }
            }
        } else if (kind == V_ENTITY_DISP) {                     // a = component (6 = magnitude): largest |.| over the entity's nodes
            for (int q = nb; q < ne; ++q) {
                const size_t n = A.ent_nodes[q];
                double val;
                if (a < 6) {
                    val = fabs(A.U[(n * 6 + a) * nr + col]);
                } else {
                    const double ux = A.U[(n * 6) * nr + col], uy = A.U[(n * 6 + 1) * nr + col], uz = A.U[(n * 6 + 2) * nr + col];
                    val = sqrt(ux * ux + uy * uy + uz * uz);
                }
                x = fmax(x, val);
            }
        } else if (kind == V_PLATE) {                           // a = 0 von Mises, 1 largest principal, 2 smallest principal
            // plate entities: ent_members lists the TRI3 rows of the plate; surface stresses sigma = n / t +- 6 m / t^2
            // at the centroid of every row, both surfaces, enveloped (max for 0 and 1, min for 2)
            double best = (a == 2) ? INFINITY : -INFINITY;
            for (int q = mb; q < me && A.p_res != nullptr; ++q) {
                const size_t e = A.ent_members[q];
                const double t = A.p_data[e * NPDATA + PD_T];
                const double* r = A.p_res + (e * 8) * (size_t)nr + col;
                for (int side = -1; side <= 1; side += 2) {
                    const double sx = r[0] / t + side * 6.0 * r[3 * (size_t)nr] / (t * t);
                    const double sy = r[(size_t)nr] / t + side * 6.0 * r[4 * (size_t)nr] / (t * t);
                    const double txy = r[2 * (size_t)nr] / t + side * 6.0 * r[5 * (size_t)nr] / (t * t);
                    const double mean = 0.5 * (sx + sy), rad = sqrt(0.25 * (sx - sy) * (sx - sy) + txy * txy);
                    const double val = a == 0 ? sqrt(sx * sx - sx * sy + sy * sy + 3.0 * txy * txy) : a == 1 ? mean + rad : mean - rad;
                    best = (a == 2) ? fmin(best, val) : fmax(best, val);
                }
            }
            x = isfinite(best) ? best : NAN;
        } else if (kind == V_DEFLECTION) {                      // a = 0 local y, 1 local z, 2 magnitude; b = aggregation
            // displacement of the entity's nodes relative to the chord between its end nodes
            const size_t n0 = A.ent_nodes[nb], n1 = A.ent_nodes[ne - 1];
            double p0[3], ch[3], u0[3], du[3], l2 = 0.0;
            for (int q = 0; q < 3; ++q) {
                p0[q] = A.xyz[n0 * 3 + q];
                ch[q] = A.xyz[n1 * 3 + q] - p0[q];
                u0[q] = A.U[(n0 * 6 + q) * nr + col];
                du[q] = A.U[(n1 * 6 + q) * nr + col] - u0[q];
                l2 += ch[q] * ch[q];
            }
            const double* fr = A.frames + (size_t)ent * 9;
            double hi = 0.0, lo = 0.0, mag = 0.0, first = 0.0, last = 0.0;
            for (int q = nb; q < ne; ++q) {
                const size_t n = A.ent_nodes[q];
                double s = 0.0, r[3];
                for (int t = 0; t < 3; ++t) s += (A.xyz[n * 3 + t] - p0[t]) * ch[t];
                s = l2 > 0.0 ? s / l2 : 0.0;
                for (int t = 0; t < 3; ++t) r[t] = A.U[(n * 6 + t) * nr + col] - (u0[t] + s * du[t]);
                double val;
                if (a == 2) val = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                else val = r[0] * fr[3 * (a + 1)] + r[1] * fr[3 * (a + 1) + 1] + r[2] * fr[3 * (a + 1) + 2];
                hi = fmax(hi, val); lo = fmin(lo, val); mag = fmax(mag, fabs(val));
                if (q == nb) first = val;
                last = val;
            }
            x = (b == A_MAX) ? hi : (b == A_MIN) ? lo : (b == A_START) ? first : (b == A_END) ? last : mag;
        }
        vals[v] = x;
    }
}

__global__ void __launch_bounds__(256) k_unity_generic(UnityArgs A) {
    const int lane = threadIdx.x & 31;
    const int ent = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ent >= A.n_ent) return;
    const int nr = A.nrhs;
    double best_u = -INFINITY, best_d = 0.0, best_c = 0.0;
    int best_col = -1;
    double vals[UNITY_MAX_VARS];
    for (int col = lane; col < nr; col += 32) {
        if (!A.col_mask[col]) continue;
        unity_eval_vars(A, ent, col, vals);
        const double d = unity_run(A.code, A.demand_off, A.demand_len, A.consts, vals);
        const double cpy = unity_run(A.code, A.cap_off, A.cap_len, A.consts, vals);
        const double u = d / cpy;
        // columns are visited in increasing order: a later column wins only with a strictly larger utilisation;
        // a NaN utilisation (0/0) is kept only until a number turns up
        const bool take = best_col < 0 || (isnan(best_u) && !isnan(u)) || (!isnan(u) && u > best_u);
        if (take) { best_u = u; best_d = d; best_c = cpy; best_col = col; }
    }
    // warp envelope: largest utilisation, lowest column on ties (a NaN never wins over a number)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ou = __shfl_xor_sync(FULL, best_u, off), od = __shfl_xor_sync(FULL, best_d, off),
                     oc = __shfl_xor_sync(FULL, best_c, off);
        const int ocol = __shfl_xor_sync(FULL, best_col, off);
        const bool both_nan = isnan(ou) && isnan(best_u);
        const bool take = ocol >= 0 && (best_col < 0 || (isnan(best_u) && !isnan(ou)) || (both_nan && ocol < best_col) ||
                                        (!isnan(ou) && !isnan(best_u) && (ou > best_u || (ou == best_u && ocol < best_col))));
        if (take) { best_u = ou; best_d = od; best_c = oc; best_col = ocol; }
    }
    if (lane == 0) {
        double* o = A.out + (size_t)ent * 4;
        o[0] = best_col >= 0 ? best_u : 0.0; o[1] = best_d; o[2] = best_c; o[3] = (double)best_col;
    }
    // hand-calculation trace: the variables at the governing column (re-evaluated by the lane that owns it)
    if (A.vars_out != nullptr && best_col >= 0 && lane == (best_col & 31)) {
        unity_eval_vars(A, ent, best_col, vals);
        for (int v = 0; v < A.n_vars; ++v) A.vars_out[(size_t)ent * A.n_vars + v] = vals[v];
    }
}

}  // namespace fers

using namespace fers;

extern "C" int fers_b200_unity_generic(int n_ent, int nrhs, int n_vars, int n_ct, const uint8_t* col_mask, const int* ent_ptr,
                                       const int* ent_members, const int* ent_node_ptr, const int* ent_nodes,
                                       const int* var_desc, const int* code, const double* consts, const double* ctab,
                                       int demand_off, int demand_len, int cap_off, int cap_len, const double* xyz,
                                       const double* u, const double* f_local, const double* sec, const double* lmax,
                                       const double* lmin, const double* frames, const int* ent_ref, const double* p_res,
                                       const double* p_data, double* out, double* vars_out, void* stream) {
    if (n_ent <= 0) return 0;
    FERS_REQUIRE(n_vars >= 0 && n_vars <= UNITY_MAX_VARS, "unity_generic: at most 96 variables per check");
    FERS_REQUIRE(col_mask && ent_ptr && ent_members && ent_node_ptr && ent_nodes && code && consts && ent_ref && out,
                 "unity_generic: bad arguments");
    FERS_REQUIRE(nrhs >= 1, "unity_generic: no columns");
    UnityArgs A;
    A.n_ent = n_ent; A.nrhs = nrhs; A.n_vars = n_vars; A.n_ct = n_ct; A.col_mask = col_mask; A.ent_ptr = ent_ptr;
    A.ent_members = ent_members; A.ent_node_ptr = ent_node_ptr; A.ent_nodes = ent_nodes; A.var_desc = var_desc;
    A.code = code; A.consts = consts; A.ctab = ctab; A.demand_off = demand_off; A.demand_len = demand_len;
    A.cap_off = cap_off; A.cap_len = cap_len; A.xyz = xyz; A.U = u; A.f_local = f_local; A.sec = sec; A.lmax = lmax;
    A.lmin = lmin; A.frames = frames; A.ent_ref = ent_ref; A.p_res = p_res; A.p_data = p_data; A.out = out;
    A.vars_out = vars_out;
    const long long threads = (long long)n_ent * 32;
    k_unity_generic<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(A);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}
