// pattern.cu -- symbolic block-CSR pattern, element slot maps, boundary-condition masks and
// free-DOF numbering on the device.
//
// Replaces the reference engine's DOF numbering / sparsity set-up (SURVEY.md 8a row A7;
// supports: fers_core/types/pydantic_models.py:2104-2115, 1956-1959).  The outputs are the
// "bit-exact" part of the parity contract: rowptr/colidx/free_index/fixed must equal the
// oracle's (oracle/fers_oracle.py: bsr_pattern, bc_masks) integer for integer.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fers {

// key = row << 32 | col for: every diagonal, both orientations of every member, the six
// off-diagonal pairs of every TRI3 plate.
__global__ void k_pattern_keys(int n_nodes, int n_members, const int* __restrict__ n1, const int* __restrict__ n2,
                               int n_plates, const int* __restrict__ pn, unsigned long long* __restrict__ keys) {
    const int total = n_nodes + 2 * n_members + 6 * n_plates;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        unsigned long long r, c;
        if (k < n_nodes) {
            r = c = (unsigned)k;
        } else if (k < n_nodes + 2 * n_members) {
            int q = k - n_nodes, m = q >> 1;
            unsigned a = (unsigned)n1[m], b = (unsigned)n2[m];
            r = (q & 1) ? b : a;
            c = (q & 1) ? a : b;
        } else {
            int q = k - n_nodes - 2 * n_members, e = q / 6, t = q % 6;
            // ordered pairs (i,j), i != j : t -> (0,1)(0,2)(1,0)(1,2)(2,0)(2,1)
            int i = t >> 1, j = (t & 1);
            if (j >= i) j += 1;
            r = (unsigned)pn[3 * e + i];
            c = (unsigned)pn[3 * e + j];
        }
        keys[k] = (r << 32) | c;
    }
}

__global__ void k_pattern_rows(int n_nodes, const unsigned long long* __restrict__ ukeys, const int* __restrict__ nnzb_p,
                               int* __restrict__ rowptr, int* __restrict__ colidx) {
    const int nnzb = *nnzb_p;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnzb; k += gridDim.x * blockDim.x) {
        unsigned long long key = ukeys[k];
        int row = (int)(key >> 32);
        colidx[k] = (int)(key & 0xffffffffull);
        if (k == 0 || (int)(ukeys[k - 1] >> 32) != row) rowptr[row] = k;   // every row owns a diagonal
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) rowptr[n_nodes] = nnzb;
}

__device__ __forceinline__ int find_slot(const int* __restrict__ rowptr, const int* __restrict__ colidx, int r, int c) {
    int lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (colidx[mid] < c) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_pattern_slots(int n_nodes, int n_members, const int* __restrict__ n1, const int* __restrict__ n2,
                                int n_plates, const int* __restrict__ pn, const int* __restrict__ rowptr,
                                const int* __restrict__ colidx, int* __restrict__ m_slot, int* __restrict__ p_slot,
                                int* __restrict__ diag_slot) {
    const int total = n_nodes + n_members + n_plates;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        if (k < n_nodes) {
            diag_slot[k] = find_slot(rowptr, colidx, k, k);
        } else if (k < n_nodes + n_members) {
            int m = k - n_nodes, a = n1[m], b = n2[m];
            int4 s;
            s.x = find_slot(rowptr, colidx, a, a);
            s.y = find_slot(rowptr, colidx, a, b);
            s.z = find_slot(rowptr, colidx, b, a);
            s.w = find_slot(rowptr, colidx, b, b);
            reinterpret_cast<int4*>(m_slot)[m] = s;
        } else {
            int e = k - n_nodes - n_members;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    p_slot[9 * e + 3 * i + j] = find_slot(rowptr, colidx, pn[3 * e + i], pn[3 * e + j]);
        }
    }
}

// fixed mask from support codes + dead-DOF detection (zero diagonal stiffness).
__global__ void k_bc_mask(int n_nodes, const uint8_t* __restrict__ code, const double* __restrict__ bc_k,
                          const double* __restrict__ vals, const int* __restrict__ diag_slot,
                          uint8_t* __restrict__ fixed, double* __restrict__ kspring, int* __restrict__ flag_free,
                          int* __restrict__ n_dead) {
    const int n_dof = 6 * n_nodes;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n_dof; d += gridDim.x * blockDim.x) {
        int node = d / 6, c = d % 6;
        uint8_t cd = code[d];
        double ks = (cd == 2) ? bc_k[d] : 0.0;
        bool fx = (cd == 1);
        if (!fx) {
            double diag = vals[(size_t)diag_slot[node] * 36 + c * 7] + ks;
            if (diag == 0.0) { fx = true; atomicAdd(n_dead, 1); }
        }
        fixed[d] = fx ? 1 : 0;
        kspring[d] = ks;
        flag_free[d] = fx ? 0 : 1;
    }
}

__global__ void k_bc_number(int n_dof, const uint8_t* __restrict__ fixed, const int* __restrict__ scan,
                            int* __restrict__ free_index) {
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n_dof; d += gridDim.x * blockDim.x)
        free_index[d] = fixed[d] ? -1 : scan[d];
}

}  // namespace fers

using namespace fers;

extern "C" {

size_t fers_b200_pattern_workspace(int n_nodes, int n_members, int n_plates) {
    size_t nk = (size_t)n_nodes + 2 * (size_t)n_members + 6 * (size_t)n_plates;
    size_t sort_tmp = 0, uniq_tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)nk);
    cub::DeviceSelect::Unique(nullptr, uniq_tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                              (int*)nullptr, (int)nk);
    size_t tmp = sort_tmp > uniq_tmp ? sort_tmp : uniq_tmp;
    return 2 * align_up(nk * 8) + align_up(tmp) + 256;
}

int fers_b200_pattern_build(int n_nodes, int n_members, const int* m_n1, const int* m_n2, int n_plates,
                            const int* p_nodes, int* rowptr, int* colidx, int cap, int* m_slot, int* p_slot,
                            int* diag_slot, int* nnzb_out, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = as_stream(stream);
    FERS_REQUIRE(n_nodes > 0, "pattern_build: model has no nodes");
    size_t nk = (size_t)n_nodes + 2 * (size_t)n_members + 6 * (size_t)n_plates;
    FERS_REQUIRE(nk < (size_t)INT32_MAX, "pattern_build: too many node pairs for 32-bit indexing");
    FERS_REQUIRE(workspace_bytes >= fers_b200_pattern_workspace(n_nodes, n_members, n_plates),
                 "pattern_build: workspace too small");
    char* w = (char*)workspace;
    unsigned long long* keys = (unsigned long long*)w;
    w += align_up(nk * 8);
    unsigned long long* ukeys = (unsigned long long*)w;
    w += align_up(nk * 8);
    int* d_n = (int*)w;
    w += 256;
    size_t tmp_bytes = workspace_bytes - (size_t)(w - (char*)workspace);
    const int T = 256;
    int grid = (int)((nk + T - 1) / T);
    k_pattern_keys<<<grid, T, 0, st>>>(n_nodes, n_members, m_n1, m_n2, n_plates, p_nodes, keys);
    FERS_LAUNCH_CHECK();
    // number of significant key bits: 32 + bits(n_nodes)
    int bits = 1;
    while ((1ll << bits) < (long long)n_nodes) ++bits;
    size_t t1 = tmp_bytes;
    FERS_CK(cub::DeviceRadixSort::SortKeys(w, t1, keys, ukeys, (int)nk, 0, 32 + bits, st));
    size_t t2 = tmp_bytes;
    FERS_CK(cub::DeviceSelect::Unique(w, t2, ukeys, keys, d_n, (int)nk, st));
    int nnzb = 0;
    FERS_CK(cudaMemcpyAsync(&nnzb, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    FERS_CK(cudaStreamSynchronize(st));
    FERS_REQUIRE(nnzb <= cap, "pattern_build: colidx capacity too small");
    grid = (nnzb + T - 1) / T;
    k_pattern_rows<<<grid, T, 0, st>>>(n_nodes, keys, d_n, rowptr, colidx);
    FERS_LAUNCH_CHECK();
    int total = n_nodes + n_members + n_plates;
    grid = (total + T - 1) / T;
    k_pattern_slots<<<grid, T, 0, st>>>(n_nodes, n_members, m_n1, m_n2, n_plates, p_nodes, rowptr, colidx, m_slot,
                                        p_slot, diag_slot);
    FERS_LAUNCH_CHECK();
    count_launches(3 + 6);   // keys, rows, slots + cub sort/unique passes
    if (nnzb_out) *nnzb_out = nnzb;
    return 0;
}

size_t fers_b200_bc_workspace(int n_dof) {
    size_t scan_tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, n_dof);
    return 2 * align_up((size_t)n_dof * 4) + align_up(scan_tmp) + 512;
}

int fers_b200_bc_build(int n_nodes, const uint8_t* bc_code, const double* bc_k, const double* vals,
                       const int* diag_slot, uint8_t* fixed, int* free_index, double* kspring, int* n_free_out,
                       int* n_dead_out, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = as_stream(stream);
    const int n_dof = 6 * n_nodes;
    FERS_REQUIRE(workspace_bytes >= fers_b200_bc_workspace(n_dof), "bc_build: workspace too small");
    char* w = (char*)workspace;
    int* flag = (int*)w;
    w += align_up((size_t)n_dof * 4);
    int* scan = (int*)w;
    w += align_up((size_t)n_dof * 4);
    int* d_cnt = (int*)w;   // [0] dead count
    w += 512;
    size_t tmp_bytes = workspace_bytes - (size_t)(w - (char*)workspace);
    FERS_CK(cudaMemsetAsync(d_cnt, 0, 512, st));
    const int T = 256;
    int grid = (n_dof + T - 1) / T;
    k_bc_mask<<<grid, T, 0, st>>>(n_nodes, bc_code, bc_k, vals, diag_slot, fixed, kspring, flag, d_cnt);
    FERS_LAUNCH_CHECK();
    FERS_CK(cub::DeviceScan::ExclusiveSum(w, tmp_bytes, flag, scan, n_dof, st));
    k_bc_number<<<grid, T, 0, st>>>(n_dof, fixed, scan, free_index);
    FERS_LAUNCH_CHECK();
    count_launches(2 + 2);
    if (n_free_out == nullptr && n_dead_out == nullptr) return 0;      // re-assembly inside an iteration: no host sync
    int last_scan = 0, last_flag = 0, dead = 0;
    FERS_CK(cudaMemcpyAsync(&last_scan, scan + n_dof - 1, 4, cudaMemcpyDeviceToHost, st));
    FERS_CK(cudaMemcpyAsync(&last_flag, flag + n_dof - 1, 4, cudaMemcpyDeviceToHost, st));
    FERS_CK(cudaMemcpyAsync(&dead, d_cnt, 4, cudaMemcpyDeviceToHost, st));
    FERS_CK(cudaStreamSynchronize(st));
    if (n_free_out) *n_free_out = last_scan + last_flag;
    if (n_dead_out) *n_dead_out = dead;
    return 0;
}

}  // extern "C"
