// loads.cu -- load-vector assembly.
//
// NodalLoad / NodalMoment: F = magnitude * direction (pydantic_models.py:1684-1697);
// DistributedLoad: linear magnitude -> end_magnitude over [start_frac, end_frac], global direction
// (pydantic_models.py:1438-1445); MemberPointLoad / MemberPointMoment in global or local axes
// (pydantic_models.py:1562-1591); PlatePressure lumped q*A/3 per corner;
// LoadCombination = factored sum of cases (pydantic_models.py:491-505).
#include "common.cuh"

namespace fers {

struct LoadArgs {
    int n_nodes, n_members;
    const double* xyz;
    const int* n1;
    const int* n2;
    const int* m_prop;
    const int* m_type;
    const double* m_roll;
    const double* m_scale;
    const int* m_rel;
    const double* prop;
    const double* rel;
    int n_nod;
    const int* nod_node;
    const double* nod_vec;
    int n_dl;
    const int* dl_member;
    const double* dl_data;
    int n_pl;
    const int* pl_member;
    const double* pl_data;
    int n_pp;
    const int* pp_plate;
    const double* pp_vec;
    const int* p_nodes;
    double* f_case;
    double* peq_case;
};

__device__ __forceinline__ Frame load_frame(const LoadArgs& A, int m) {
    const int a = A.n1[m], b = A.n2[m];
    double pa[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
    double pb[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
    if (A.m_rel && A.m_rel[m] >= 0) rec_offset_points(A.rel + REL_STRIDE * (size_t)A.m_rel[m], pa, pb);
    return member_frame(pa, pb, A.m_roll[m]);
}

// pass 1: every load record -> nodal vector or local equivalent member loads
__global__ void k_load_records(LoadArgs A) {
    const int total = A.n_nod + A.n_dl + A.n_pl + A.n_pp;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        if (k < A.n_nod) {
            const int n = A.nod_node[k];
            for (int c = 0; c < 6; ++c) {
                double v = A.nod_vec[6 * (size_t)k + c];
                if (v != 0.0) atomicAdd(A.f_case + 6 * (size_t)n + c, v);
            }
        } else if (k < A.n_nod + A.n_dl + A.n_pl) {
            const bool dist = k < A.n_nod + A.n_dl;
            const int q = dist ? k - A.n_nod : k - A.n_nod - A.n_dl;
            const int m = dist ? A.dl_member[q] : A.pl_member[q];
            const double* row = (dist ? A.dl_data : A.pl_data) + 8 * (size_t)q;
            Frame f = load_frame(A, m);
            const int mtype = A.m_type[m];
            BeamCoef c = beam_coef(A.prop + (size_t)A.m_prop[m] * NPROP, f.L, mtype, 1.0);
            double p[12];
            for (int i = 0; i < 12; ++i) p[i] = 0.0;
            if (dist) beam_dist_load(f, c, mtype, row, p);
            else beam_point_load(f, c, mtype, row, p);
            for (int i = 0; i < 12; ++i)
                if (p[i] != 0.0) atomicAdd(A.peq_case + 12 * (size_t)m + i, p[i]);
        } else {
            const int q = k - A.n_nod - A.n_dl - A.n_pl;
            const int e = A.pp_plate[q];
            const int a = A.p_nodes[3 * e], b = A.p_nodes[3 * e + 1], c = A.p_nodes[3 * e + 2];
            double p0[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
            double p1[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
            double p2[3] = {A.xyz[3 * c], A.xyz[3 * c + 1], A.xyz[3 * c + 2]};
            PlateGeom g = plate_geom(p0, p1, p2);
            const double w = fabs(g.A) / 3.0;
            const int nd[3] = {a, b, c};
            for (int i = 0; i < 3; ++i)
                for (int d = 0; d < 3; ++d) atomicAdd(A.f_case + 6 * (size_t)nd[i] + d, w * A.pp_vec[3 * (size_t)q + d]);
        }
    }
}

// pass 2: per member, condense equivalent loads for end releases and add T^T p to the nodes
__global__ void k_load_members(LoadArgs A) {
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < A.n_members; m += gridDim.x * blockDim.x) {
        double p[12];
        bool any = false;
        for (int i = 0; i < 12; ++i) { p[i] = A.peq_case[12 * (size_t)m + i]; any |= (p[i] != 0.0); }
        if (!any) continue;
        Frame f = load_frame(A, m);
        const int ri = A.m_rel ? A.m_rel[m] : -1;
        if (ri >= 0) {
            const double scale = A.m_scale ? A.m_scale[m] : 1.0;
            BeamCoef c = beam_coef(A.prop + (size_t)A.m_prop[m] * NPROP, f.L, A.m_type[m], scale);
            double k[144];
            beam_local_k(c, k);
            beam_modify(k, A.rel + REL_STRIDE * (size_t)ri, f, p, 1);
            for (int i = 0; i < 12; ++i) A.peq_case[12 * (size_t)m + i] = p[i];
        }
        const int nd[2] = {A.n1[m], A.n2[m]};
        for (int b = 0; b < 4; ++b) {
            const int node = nd[b >> 1];
            for (int c = 0; c < 3; ++c) {
                double v = f.r[0][c] * p[3 * b] + f.r[1][c] * p[3 * b + 1] + f.r[2][c] * p[3 * b + 2];
                if (v != 0.0) atomicAdd(A.f_case + 6 * (size_t)node + 3 * (b & 1) + c, v);
            }
        }
    }
}

// out[row][rhs] = sum_l fac[rhs][l] * in[l][row].  A group of nrhs/2 lanes owns one row (two columns per lane,
// one 16-byte store: a warp writes 512 contiguous bytes), the row's case values are loaded once per group
// as broadcasts, four rows in flight per group.
__global__ void __launch_bounds__(256)
k_combine(long long n_rows, int n_cases, int nrhs, const double* __restrict__ in, const double* __restrict__ fac,
          double* __restrict__ out) {
    extern __shared__ double s_fac[];          // [n_cases][nrhs] (transposed: conflict-free column pairs)
    for (int i = threadIdx.x; i < nrhs * n_cases; i += blockDim.x) {
        const int r = i / n_cases, l = i - r * n_cases;
        s_fac[l * nrhs + r] = fac[i];
    }
    __syncthreads();
    const int cp = nrhs / 2;                   // lanes per row
    const int lane_in_row = threadIdx.x % cp;
    const long long groups = (long long)gridDim.x * (blockDim.x / cp);
    const long long g0 = (long long)blockIdx.x * (blockDim.x / cp) + threadIdx.x / cp;
    for (long long row0 = g0 * 4; row0 < n_rows; row0 += groups * 4) {
        double2 acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = make_double2(0.0, 0.0);
        for (int l = 0; l < n_cases; ++l) {
            const double2 f = *reinterpret_cast<const double2*>(s_fac + l * nrhs + 2 * lane_in_row);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long row = row0 + u;
                const double v = row < n_rows ? __ldg(in + (size_t)l * n_rows + row) : 0.0;
                acc[u].x = fma(f.x, v, acc[u].x); acc[u].y = fma(f.y, v, acc[u].y);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = row0 + u;
            if (row < n_rows) reinterpret_cast<double2*>(out + (size_t)row * nrhs)[lane_in_row] = acc[u];
        }
    }
}

// general shape (odd nrhs or nrhs/2 not dividing 256): one thread per output value
__global__ void k_combine_any(int n_rows, int n_cases, int nrhs, const double* __restrict__ in,
                              const double* __restrict__ fac, double* __restrict__ out) {
    extern __shared__ double s_fac[];
    for (int i = threadIdx.x; i < nrhs * n_cases; i += blockDim.x) s_fac[i] = fac[i];
    __syncthreads();
    const size_t total = (size_t)n_rows * nrhs;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t row = t / nrhs;
        const int r = (int)(t - row * nrhs);
        double acc = 0.0;
        for (int l = 0; l < n_cases; ++l) acc += s_fac[r * n_cases + l] * __ldg(in + (size_t)l * n_rows + row);
        out[t] = acc;
    }
}

}  // namespace fers

using namespace fers;

extern "C" {

int fers_b200_load_case(int n_nodes, int n_members, const double* xyz, const int* m_n1, const int* m_n2,
                        const int* m_prop, const int* m_type, const double* m_roll, const double* m_scale,
                        const int* m_rel, const double* prop, const double* rel, int n_nod, const int* nod_node,
                        const double* nod_vec, int n_dl, const int* dl_member, const double* dl_data, int n_pl,
                        const int* pl_member, const double* pl_data, int n_pp, const int* pp_plate,
                        const double* pp_vec, const int* p_nodes, double* f_case, double* peq_case, void* stream) {
    LoadArgs A;
    A.n_nodes = n_nodes; A.n_members = n_members; A.xyz = xyz; A.n1 = m_n1; A.n2 = m_n2; A.m_prop = m_prop;
    A.m_type = m_type; A.m_roll = m_roll; A.m_scale = m_scale; A.m_rel = m_rel; A.prop = prop; A.rel = rel;
    A.n_nod = n_nod; A.nod_node = nod_node; A.nod_vec = nod_vec; A.n_dl = n_dl; A.dl_member = dl_member;
    A.dl_data = dl_data; A.n_pl = n_pl; A.pl_member = pl_member; A.pl_data = pl_data; A.n_pp = n_pp;
    A.pp_plate = pp_plate; A.pp_vec = pp_vec; A.p_nodes = p_nodes; A.f_case = f_case; A.peq_case = peq_case;
    cudaStream_t st = as_stream(stream);
    const int T = 128;
    long long total = (long long)n_nod + n_dl + n_pl + n_pp;
    if (total > 0) {
        int grid = (int)((total + T - 1) / T);
        k_load_records<<<grid, T, 0, st>>>(A);
        FERS_LAUNCH_CHECK();
        count_launches(1);
    }
    if (n_members > 0 && (n_dl + n_pl) > 0) {
        int grid = (n_members + T - 1) / T;
        k_load_members<<<grid, T, 0, st>>>(A);
        FERS_LAUNCH_CHECK();
        count_launches(1);
    }
    return 0;
}

int fers_b200_combine(int n_rows, int n_cases, int nrhs, const double* in, const double* factors, double* out,
                      void* stream) {
    if (n_rows <= 0 || nrhs <= 0) return 0;
    FERS_REQUIRE((size_t)nrhs * n_cases * 8 <= 48 * 1024, "combine: factor table exceeds shared memory");
    const int T = 256;
    const size_t smem = (size_t)nrhs * n_cases * 8;
    const int cp = nrhs / 2;
    if (nrhs % 2 == 0 && cp >= 1 && T % cp == 0 && (((uintptr_t)out) & 15) == 0) {
        const long long rows_per_cta = (long long)(T / cp) * 4;
        long long grid = (n_rows + rows_per_cta - 1) / rows_per_cta;
        const long long cap = (long long)sm_count() * 16;
        if (grid > cap) grid = cap;
        k_combine<<<(unsigned)grid, T, smem, as_stream(stream)>>>(n_rows, n_cases, nrhs, in, factors, out);
    } else {
        size_t total = (size_t)n_rows * nrhs;
        int grid = (int)((total + T - 1) / T);
        int cap = sm_count() * 16;
        if (grid > cap) grid = cap;
        k_combine_any<<<grid, T, smem, as_stream(stream)>>>(n_rows, n_cases, nrhs, in, factors, out);
    }
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
