// chol.cuh -- data structures of the sparse multifrontal Cholesky (chol_plan.cu: ordering + symbolic
// analysis on the host; chol_numeric.cu: numeric factorisation and multi-RHS substitution on the GPU).
//
// The reference engine documents its first-order path as "a batch first-order solve that shares one
// factorization across many load cases" (fers_core/types/pydantic_models.py:1216-1219); this is that
// factorisation.  Unknowns are eliminated node by node (6x6 blocks), ordered by nested dissection on the
// node coordinates; every tree node ("front") owns a set of nodes and a boundary (the later-eliminated
// nodes its Schur complement touches).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace fers {

constexpr int CHOL_NB = 64;            // panel width (DOFs) of the right-looking factorisation / substitution

// One front as the device kernels see it (all sizes in NODES; DOF counts are 6x).
struct FrontDev {
    int n1;               // own nodes (eliminated here)
    int n2;               // boundary nodes (later-eliminated nodes coupled to the own nodes)
    int nodes_off;        // into nodes[]: n1 own nodes then n2 boundary nodes, elimination order
    int child_begin;      // into child[] (front ids) and inv_off[]
    int child_end;
    int parent;           // front id or -1
    int pad0;             // elimination position of the first own node (row 6*pad0 of the permuted solution)
    int pad1;
    long long panel_off;  // doubles into the factor store: panel [6(n1+n2)] x [6 n1], column-major
    long long schur_off;  // doubles into the Schur arena of this front's level parity: [6 n2] x [6 n2]
    long long dinv_off;   // doubles into the factor store: ceil(6 n1 / NB) inverted diagonal tiles NB x NB
    long long rhs_off;    // rows into the front-local right-hand-side store ([6(n1+n2)] rows of nrhs)
};

// One K block assembled into a front panel: vals block k -> panel position (block row bi, block col bj)
struct KScatter {
    int front;
    int k;       // block index in the block-CSR values
    int bi, bj;  // local node indices inside the front (bi >= bj, bj < n1)
};

struct CholPlanHost {
    int n_nodes = 0;
    int n_active = 0;               // nodes taking part in the elimination (n_nodes minus fully restrained ones)
    int n_fronts = 0;
    int n_levels = 0;
    std::vector<int> perm;          // elimination position -> node
    std::vector<int> pos;           // node -> elimination position (-1: fully restrained node, left out)
    std::vector<FrontDev> fronts;   // in elimination (post-)order
    std::vector<int> front_first;   // elimination position of a front's first own node
    std::vector<int> nodes;         // concatenated front node lists
    std::vector<int> child;         // concatenated child lists
    std::vector<long long> inv_off; // per child entry: offset into inv[] of the parent-sized inverse map
    std::vector<int> inv;           // parent local node -> child boundary index or -1
    std::vector<int> level_ptr;     // levels are processed 0 .. n_levels-1 (deepest tree depth first)
    std::vector<int> level_fronts;
    std::vector<KScatter> kscatter; // grouped by level
    std::vector<long long> kscatter_ptr;
    std::vector<int> level_max_n1, level_max_n;   // per level: largest own / total node count
    long long factor_doubles = 0;   // panels + inverted diagonal tiles
    long long schur_doubles[2] = {0, 0};
    long long rhs_rows = 0;         // total front-local rhs rows
    double flops = 0;               // factorisation flops (potrf + trsm + syrk)
    long long nnz_l = 0;            // scalar entries of L (lower)
    int max_front_nodes = 0;
    // device index buffer layout (bytes), filled by layout()
    size_t off_fronts = 0, off_nodes = 0, off_child = 0, off_inv_off = 0, off_inv = 0, off_level_fronts = 0,
           off_kscatter = 0, off_pos = 0, index_bytes = 0;
};

// fp64-equivalent Schur update on the int8 tensor cores (ozaki_syrk.cu): C (m2 x m2, lower) -= A A^T, A m2 x k1
size_t oz_work_bytes(int m2, int k1, int S);
// ncols: only the first ncols columns of C (0: all); persistent: one CTA per SM walking the tiles, else one CTA per tile
int oz_syrk(const double* A, long long lda, double* C, long long ldc, int m2, int k1, int S, void* work, cudaStream_t st,
            int ncols = 0, bool persistent = true);

}  // namespace fers
