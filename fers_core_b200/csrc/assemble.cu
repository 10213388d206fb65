// assemble.cu -- element stiffness formation + global assembly into the block-CSR value array.
//
// SURVEY.md 8a rows A3-A7.  Beam: 12x12 Euler-Bernoulli/Timoshenko local stiffness, optional
// end-release condensation (pydantic_models.py:2062-2087) and geometric stiffness
// (pydantic_models.py:1377-1384), T^T k T with the local axes of fers_core/members/member.py:309-354.
// Plate: CST + DSG3 TRI3.
//
// Kernel shape (beams): one thread forms one member's stiffness, one warp owns a batch of 32
// members.  Each of the four 6x6 node blocks is staged through a per-warp shared-memory tile
// (padded stride 33: conflict-free writes, 2-wavefront fp64 reads) so that the scatter-add into
// HBM is issued as 32-lane runs of CONSECUTIVE addresses (one 288-byte block = 9 sectors) instead
// of 32 scattered 8-byte atomics.  Members of the batch that hit the same block (members sharing
// their start node sit next to each other because members are ordered by start node) are summed
// in shared memory first and only the group leader issues the atomic ("warp-aggregated").
// The section/material table is staged into shared memory with one TMA bulk copy
// (cp.async.bulk + mbarrier) while the threads fetch their coordinates.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fers {

constexpr int ASM_THREADS = 128;
constexpr int ASM_WARPS = ASM_THREADS / 32;
constexpr int STAGE_STRIDE = 33;
constexpr int STAGE_DOUBLES = 36 * STAGE_STRIDE;
constexpr int MAX_PROP_SMEM_ROWS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}

// 256-bit global store (sm_100+): one full 32-byte sector per lane
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

struct BeamArgs {
    int n_members;
    const double* xyz;
    const int* n1;
    const int* n2;
    const int* m_prop;
    const int* m_type;
    const double* m_roll;
    const double* m_scale;
    const int* m_rel;
    const double* prop;
    int n_prop;
    const double* rel;
    const double* m_axial;
    int kg_mode;
    int kg_suppress;
    const int* m_slot;
    double* vals;
    int aggregate;
    int use_tma;
};

// zero the rows/cols of a global 6x6 node block of K_g that belong to suppressed translations
__device__ __forceinline__ void suppress_block(double* blk, int mask) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (mask & (1 << a)) {
#pragma unroll
            for (int q = 0; q < 6; ++q) { blk[a * 6 + q] = 0.0; blk[q * 6 + a] = 0.0; }
        }
}

// GENERAL = false: closed-form blocks only (no releases, no K_g) -- the cfg4 hot path.
// GENERAL = true : dense local matrices in thread-local memory (releases and/or P-Delta).
template <bool GENERAL>
__global__ void __launch_bounds__(ASM_THREADS) k_assemble_beams(BeamArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_stage = reinterpret_cast<double*>(smem_raw);
    double* s_prop = s_stage + ASM_WARPS * STAGE_DOUBLES;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_prop + (size_t)(A.use_tma ? A.n_prop : 0) * NPROP);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t prop_bytes = (uint32_t)A.n_prop * NPROP * sizeof(double);
    if (A.use_tma) {
        if (tid == 0) mbar_init(s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(s_bar, prop_bytes);
            tma_bulk_g2s(s_prop, A.prop, prop_bytes, s_bar);
        }
    }

    const int m = blockIdx.x * ASM_THREADS + tid;
    const bool valid = m < A.n_members;
    Frame f;
    BeamCoef c;
    int4 slots = make_int4(-1, -1, -1, -1);
    int mtype = 0, pi = 0, ri = -1;
    double scale = 1.0, axial = 0.0;
    bool general = false;
    if (valid) {
        const int a = A.n1[m], b = A.n2[m];
        double pa[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
        double pb[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
        if (GENERAL && A.m_rel) ri = A.m_rel[m];
        if (GENERAL && ri >= 0) rec_offset_points(A.rel + (size_t)ri * REL_STRIDE, pa, pb);
        f = member_frame(pa, pb, A.m_roll[m]);
        mtype = A.m_type[m];
        pi = A.m_prop[m];
        if (A.m_scale) scale = A.m_scale[m];
        if (A.m_axial && A.kg_mode) axial = A.m_axial[m];
        slots = reinterpret_cast<const int4*>(A.m_slot)[m];
        general = GENERAL && ((ri >= 0) || (axial != 0.0));
    }
    if (A.use_tma) mbar_wait(s_bar, 0);
    if (valid) {
        const double* P = A.use_tma ? (s_prop + (size_t)pi * NPROP) : (A.prop + (size_t)pi * NPROP);
        double Pr[NPROP];
#pragma unroll
        for (int i = 0; i < NPROP; ++i) Pr[i] = P[i];
        c = beam_coef(Pr, f.L, mtype, scale);
    }
    // general path: dense local matrices in thread-local memory (rare: releases / P-Delta)
    double kE[GENERAL ? 144 : 1], kG[GENERAL ? 144 : 1];
    if (GENERAL && general) {
        beam_local_k(c, kE);
        bool have_g = (axial != 0.0);
        if (have_g) {
            for (int i = 0; i < 144; ++i) kG[i] = 0.0;
            beam_add_kg(kG, axial, f.L, A.kg_mode == 1, mtype);
        }
        // releases condense the elastic part only (K_g is added un-condensed, as the oracle does); the rigid arms
        // of the end offsets carry both
        if (ri >= 0) {
            beam_modify(kE, A.rel + (size_t)ri * REL_STRIDE, f, nullptr, 0);
            if (have_g) beam_arms(kG, A.rel + (size_t)ri * REL_STRIDE, f, nullptr, 0);
        }
    }

    double* stage = s_stage + warp * STAGE_DOUBLES;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int bi = b >> 1, bj = b & 1;
        double blk[36];
        if (valid) {
            if (!GENERAL || !general) {
                beam_block_closed(c, f, bi, bj, blk);
            } else {
                beam_rotate_block(kE, f, bi, bj, blk);
                if (axial != 0.0) {
                    double g[36];
                    beam_rotate_block(kG, f, bi, bj, g);
                    if (A.kg_suppress) suppress_block(g, A.kg_suppress);
#pragma unroll
                    for (int e = 0; e < 36; ++e) blk[e] += g[e];
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 36; ++e) blk[e] = 0.0;
        }
#pragma unroll
        for (int e = 0; e < 36; ++e) stage[e * STAGE_STRIDE + lane] = blk[e];
        const int slot = valid ? (b == 0 ? slots.x : b == 1 ? slots.y : b == 2 ? slots.z : slots.w) : -1;
        unsigned peers = A.aggregate ? __match_any_sync(FULL, slot) : (1u << lane);
        const int leader = __ffs(peers) - 1;
        __syncwarp();
        // cooperative scatter: w enumerates (member-in-batch, entry) with the entry fastest
        for (int w = lane; w < 32 * 36; w += 32) {
            const int mm = w / 36, e = w - mm * 36;
            const int s = __shfl_sync(FULL, slot, mm);
            const unsigned pr = __shfl_sync(FULL, peers, mm);
            const int ld = __shfl_sync(FULL, leader, mm);
            if (s >= 0 && ld == mm) {
                double v = 0.0;
                unsigned rest = pr;
                while (rest) {
                    const int q = __ffs(rest) - 1;
                    rest &= rest - 1;
                    v += stage[e * STAGE_STRIDE + q];
                }
                atomicAdd(A.vals + (size_t)s * 36 + e, v);
            }
        }
        __syncwarp();
    }
}

// ---- gather formulation (default): no atomics, no zero pass, bit-reproducible ---------------------
// Work unit = "half-member" (member, end): the thread forms the member's frame and coefficients and
// the two 6x6 blocks that live in the block row of its own node: the off-diagonal block (own, other),
// which has exactly one contributor and is stored straight from registers, and the member's
// contribution to the diagonal block (own, own).  Half-members are sorted by block slot (= by node),
// a CTA owns GA_NODES consecutive nodes, the diagonal contributions meet in shared memory and are
// summed in list order by fixed (node, entry) owners, so every value of K is written exactly once.
// Duplicate members between the same node pair (flagged at map-build time) are summed by the first
// of them.  Element arithmetic is the same as in k_assemble_beams (each member is formed twice, two
// blocks each time, instead of once with four blocks).
constexpr int GA_THREADS = 128;
constexpr int GA_NODES = 20;      // 20 nodes x up to 6 members = 120 half-members: one chunk of GA_THREADS on a 3D grid frame
constexpr int GA_STRIDE = 37;
constexpr int GA_PAIRS = (GA_NODES * 36 + GA_THREADS - 1) / GA_THREADS;

struct GatherArgs {
    BeamArgs B;
    int n_nodes;
    const int* hm_ptr;     // [n_nodes + 1]
    const int* hm_list;    // [2 * n_members]  member << 2 | dup << 1 | end
    const int* diag_slot;  // [n_nodes]
    double* cache;         // [n_members][MC_DOUBLES] frames + coefficients written by k_member_cache, or NULL
};

// Per-member invariants of one assembly pass: local frame (9), length, closed-form coefficients (12).
// Formed once per member by k_member_cache and read by both of its half-members in k_assemble_gather
// (64 MB at configs[3]: stays in the 126 MB L2 between the two kernels).
constexpr int MC_DOUBLES = 24;

__global__ void __launch_bounds__(256) k_member_cache(BeamArgs A, double* __restrict__ cache) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_prop = reinterpret_cast<double*>(smem_raw);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_prop + (size_t)(A.use_tma ? A.n_prop : 0) * NPROP);
    if (A.use_tma) {          // section / material table: one TMA bulk copy per CTA, overlapped with the coordinate loads
        if (threadIdx.x == 0) mbar_init(s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t prop_bytes = (uint32_t)A.n_prop * NPROP * sizeof(double);
            mbar_expect_tx(s_bar, prop_bytes);
            tma_bulk_g2s(s_prop, A.prop, prop_bytes, s_bar);
        }
    }
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = m < A.n_members;
    Frame f;
    int mtype = 0, pi = 0;
    double scale = 1.0;
    if (valid) {
        const int a = A.n1[m], b = A.n2[m];
        double pa[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
        double pb[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
        if (A.m_rel && A.m_rel[m] >= 0) rec_offset_points(A.rel + (size_t)A.m_rel[m] * REL_STRIDE, pa, pb);
        f = member_frame(pa, pb, A.m_roll[m]);
        mtype = A.m_type[m];
        pi = A.m_prop[m];
        if (A.m_scale) scale = A.m_scale[m];
    }
    if (A.use_tma) mbar_wait(s_bar, 0);
    if (!valid) return;
    const double* P = (A.use_tma ? s_prop : A.prop) + (size_t)pi * NPROP;
    double Pr[NPROP];
#pragma unroll
    for (int i = 0; i < NPROP; ++i) Pr[i] = P[i];
    const BeamCoef c = beam_coef(Pr, f.L, mtype, scale);
    double* o = cache + (size_t)m * MC_DOUBLES;
    st_global_v4(o, f.r[0][0], f.r[0][1], f.r[0][2], f.r[1][0]);
    st_global_v4(o + 4, f.r[1][1], f.r[1][2], f.r[2][0], f.r[2][1]);
    st_global_v4(o + 8, f.r[2][2], f.L, c.a, c.t);
    st_global_v4(o + 12, c.z12, c.z6, c.z4, c.z2);
    st_global_v4(o + 16, c.y12, c.y6, c.y4, c.y2);
    st_global_v4(o + 20, c.phi_y, c.phi_z, 0.0, 0.0);
}

__device__ __forceinline__ void load_member_cache(const double* __restrict__ cache, int m, Frame& f, BeamCoef& c) {
    const double2* p = reinterpret_cast<const double2*>(cache + (size_t)m * MC_DOUBLES);
    double2 v[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) v[i] = __ldg(p + i);
    f.r[0][0] = v[0].x; f.r[0][1] = v[0].y; f.r[0][2] = v[1].x; f.r[1][0] = v[1].y; f.r[1][1] = v[2].x; f.r[1][2] = v[2].y;
    f.r[2][0] = v[3].x; f.r[2][1] = v[3].y; f.r[2][2] = v[4].x; f.L = v[4].y;
    c.a = v[5].x; c.t = v[5].y; c.z12 = v[6].x; c.z6 = v[6].y; c.z4 = v[7].x; c.z2 = v[7].y;
    c.y12 = v[8].x; c.y6 = v[8].y; c.y4 = v[9].x; c.y2 = v[9].y; c.phi_y = v[10].x; c.phi_z = v[10].y;
}

// one 6x6 block (bi, bj) of member m, self-contained (duplicate-member path)
template <bool GENERAL>
__device__ __noinline__ void member_block(const BeamArgs& A, const double* prop_tab, int m, int bi, int bj, double* blk) {
    const int a = A.n1[m], b = A.n2[m];
    double pa[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
    double pb[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
    const int ri = A.m_rel ? A.m_rel[m] : -1;
    if (ri >= 0) rec_offset_points(A.rel + (size_t)ri * REL_STRIDE, pa, pb);
    const Frame f = member_frame(pa, pb, A.m_roll[m]);
    const int mtype = A.m_type[m];
    const double scale = A.m_scale ? A.m_scale[m] : 1.0;
    const double axial = (A.m_axial && A.kg_mode) ? A.m_axial[m] : 0.0;
    double Pr[NPROP];
    const double* P = prop_tab + (size_t)A.m_prop[m] * NPROP;
#pragma unroll
    for (int i = 0; i < NPROP; ++i) Pr[i] = P[i];
    const BeamCoef c = beam_coef(Pr, f.L, mtype, scale);
    if (GENERAL && ((ri >= 0) || (axial != 0.0))) {
        double kE[144];
        beam_local_k(c, kE);
        if (ri >= 0) beam_modify(kE, A.rel + (size_t)ri * REL_STRIDE, f, nullptr, 0);
        beam_rotate_block(kE, f, bi, bj, blk);
        if (axial != 0.0) {
            for (int i = 0; i < 144; ++i) kE[i] = 0.0;
            beam_add_kg(kE, axial, f.L, A.kg_mode == 1, mtype);
            if (ri >= 0) beam_arms(kE, A.rel + (size_t)ri * REL_STRIDE, f, nullptr, 0);
            double g[36];
            beam_rotate_block(kE, f, bi, bj, g);
            if (A.kg_suppress) suppress_block(g, A.kg_suppress);
#pragma unroll
            for (int e = 0; e < 36; ++e) blk[e] += g[e];
        }
    } else {
        beam_block_closed(c, f, bi, bj, blk);
    }
}

template <bool GENERAL, bool CACHED>
__global__ void __launch_bounds__(GA_THREADS, GENERAL ? 1 : 5) k_assemble_gather(GatherArgs G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BeamArgs& A = G.B;
    double* s_stage = reinterpret_cast<double*>(smem_raw);
    double* s_prop = s_stage + GA_THREADS * GA_STRIDE;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_prop + (size_t)(A.use_tma ? A.n_prop : 0) * NPROP);
    __shared__ int s_hp[GA_NODES + 1];

    const int tid = threadIdx.x;
    if (A.use_tma) {          // (host: only when the member cache is off -- the cache kernel reads the table)
        if (tid == 0) mbar_init(s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t prop_bytes = (uint32_t)A.n_prop * NPROP * sizeof(double);
            mbar_expect_tx(s_bar, prop_bytes);
            tma_bulk_g2s(s_prop, A.prop, prop_bytes, s_bar);
        }
    }
    const int node_a = blockIdx.x * GA_NODES;
    const int nn = min(GA_NODES, G.n_nodes - node_a);
    if (tid <= nn) s_hp[tid] = G.hm_ptr[node_a + tid];
    __syncthreads();
    const int h_begin = s_hp[0], h_end = s_hp[nn];
    const double* prop_tab = A.use_tma ? s_prop : A.prop;
    bool prop_ready = !A.use_tma;

    double acc[GA_PAIRS];
#pragma unroll
    for (int q = 0; q < GA_PAIRS; ++q) acc[q] = 0.0;

    for (int chunk0 = h_begin; chunk0 < h_end; chunk0 += GA_THREADS) {
        const int h = chunk0 + tid;
        const bool valid = h < h_end;
        if (valid) {
            const int code = G.hm_list[h];
            const int m = code >> 2, end = code & 1;
            const bool dup = (code & 2) != 0;
            const int ri = (GENERAL && A.m_rel) ? A.m_rel[m] : -1;
            const double axial = (GENERAL && A.m_axial && A.kg_mode) ? A.m_axial[m] : 0.0;
            const int4 slots = reinterpret_cast<const int4*>(A.m_slot)[m];
            const int off_slot = end ? slots.z : slots.y, dg_slot = end ? slots.w : slots.x;
            Frame f;
            BeamCoef c;
            int mtype = 0;
            if (CACHED) {
                load_member_cache(G.cache, m, f, c);
                if (GENERAL) mtype = A.m_type[m];
            } else {
                const int a = A.n1[m], b = A.n2[m];
                double pa[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
                double pb[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
                if (GENERAL && ri >= 0) rec_offset_points(A.rel + (size_t)ri * REL_STRIDE, pa, pb);
                f = member_frame(pa, pb, A.m_roll[m]);
                mtype = A.m_type[m];
                const int pi = A.m_prop[m];
                const double scale = A.m_scale ? A.m_scale[m] : 1.0;
                if (!prop_ready) { mbar_wait(s_bar, 0); prop_ready = true; }
                double Pr[NPROP];
#pragma unroll
                for (int i = 0; i < NPROP; ++i) Pr[i] = prop_tab[(size_t)pi * NPROP + i];
                c = beam_coef(Pr, f.L, mtype, scale);
            }
            const bool general = GENERAL && ((ri >= 0) || (axial != 0.0));
            double kE[GENERAL ? 144 : 1], kG[GENERAL ? 144 : 1];
            if (GENERAL && general) {
                beam_local_k(c, kE);
                if (axial != 0.0) {
                    for (int i = 0; i < 144; ++i) kG[i] = 0.0;
                    beam_add_kg(kG, axial, f.L, A.kg_mode == 1, mtype);
                }
                if (ri >= 0) {
                    beam_modify(kE, A.rel + (size_t)ri * REL_STRIDE, f, nullptr, 0);
                    if (axial != 0.0) beam_arms(kG, A.rel + (size_t)ri * REL_STRIDE, f, nullptr, 0);
                }
            }
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {          // 0: diagonal contribution, 1: off-diagonal block
                const int bi = end, bj = pass ? 1 - end : end;
                double blk[36];
                if (!GENERAL || !general) {
                    beam_block_closed(c, f, bi, bj, blk);
                } else {
                    beam_rotate_block(kE, f, bi, bj, blk);
                    if (axial != 0.0) {
                        double g[36];
                        beam_rotate_block(kG, f, bi, bj, g);
                        if (A.kg_suppress) suppress_block(g, A.kg_suppress);
#pragma unroll
                        for (int e = 0; e < 36; ++e) blk[e] += g[e];
                    }
                }
                if (pass == 0) {
#pragma unroll
                    for (int e = 0; e < 36; ++e) s_stage[tid * GA_STRIDE + e] = blk[e];
                } else if (!dup && off_slot != dg_slot) {       // duplicated node pairs: k_gather_dups
                    // 288-byte block = nine full 32-byte sectors: 256-bit stores, no partial-sector writes
                    double* dst = A.vals + (size_t)off_slot * 36;
#pragma unroll
                    for (int e = 0; e < 9; ++e) st_global_v4(dst + 4 * e, blk[4 * e], blk[4 * e + 1], blk[4 * e + 2], blk[4 * e + 3]);
                }
            }
        }
        __syncthreads();
        // fixed owners sum the diagonal contributions of this chunk in list order
#pragma unroll
        for (int q = 0; q < GA_PAIRS; ++q) {
            const int p = tid + q * GA_THREADS;
            const int nl = p / 36, e = p - nl * 36;
            if (nl < nn) {
                const int lo = max(s_hp[nl], chunk0), hi = min(s_hp[nl + 1], chunk0 + GA_THREADS);
                for (int hh = lo; hh < hi; ++hh) acc[q] += s_stage[(hh - chunk0) * GA_STRIDE + e];
            }
        }
        __syncthreads();
    }
    if (!prop_ready) mbar_wait(s_bar, 0);     // never leave the CTA with a bulk copy in flight
#pragma unroll
    for (int q = 0; q < GA_PAIRS; ++q) {
        const int p = tid + q * GA_THREADS;
        const int nl = p / 36, e = p - nl * 36;
        if (nl < nn) A.vals[(size_t)G.diag_slot[node_a + nl] * 36 + e] = acc[q];
    }
}

// ---- gather assembly of plain members (no modifier record, no K_g): element-lane kernel -----------------------------------
// Every 6x6 block of an un-released member is +- one of four 3x3 matrices (x, y, z = rows of the member frame):
//   TT = a x x^T + z12 y y^T + y12 z z^T     R4 = t x x^T + y4 y y^T + z4 z z^T     R2 = -t x x^T + y2 y y^T + z2 z z^T
//   TR = z6 y z^T - y6 z y^T                 (TT, R4, R2 symmetric: 6 + 6 + 6 + 9 = 27 values per member)
//   K_ss = [TT TR; TR^T R4]   K_se = [-TT TR; -TR^T R2]   K_es = [-TT -TR; TR^T R2]   K_ee = [TT -TR; -TR^T R4]
// Phase A: one thread per half-member forms the 27 values in shared memory: frame and coefficients on one reciprocal
// square root of the length (the rule of member_frame / beam_coef, equal to them to a few ulp; the fp64 divisions of the
// shared functions were two thirds of the instructions of this kernel).  Phase B: groups of 18 lanes, one lane per pair of
// neighbouring entries of a 6x6 block, stream the off-diagonal blocks out (consecutive half-members of the list own
// consecutive slots: 16-byte stores coalesce into whole 288-byte blocks; signs are applied to the high word) and sum the
// diagonal contributions of the CTA's nodes in list order (the order of k_assemble_gather: no atomics, bit-reproducible).
constexpr int GP_GROUPS = 8;                      // groups of 18 lanes: a lane owns two neighbouring entries of a 6x6 block
constexpr int GP_LANES = 18;
constexpr int GP_THREADS = GP_LANES * GP_GROUPS;  // 144
constexpr int GP_NODES = 24;                      // 24 nodes x 6 members = 144 half-members: one chunk on a 3D grid frame
constexpr int GP_VALS = 27;
constexpr int GP_QS = GP_NODES / GP_GROUPS;       // nodes per group
static_assert(GP_NODES % GP_GROUPS == 0 && GP_NODES + 1 <= GP_THREADS, "plain gather kernel geometry");

__device__ __forceinline__ int gp_sym6(int r, int q) {
    const int lo = min(r, q), hi = max(r, q);
    return lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
}
// where entry (R, C) of a 6x6 block sits in the 27 values (diagonal kinds K_ss / K_ee and off-diagonal kinds K_se / K_es)
// and its sign per kind as a mask on the high word: bit 0 of `sg` = K_se, bit 1 = K_es, bit 2 = K_ee (K_ss carries none)
__device__ __forceinline__ void gp_entry(int R, int C, int& idx_dg, int& idx_off, int& sg) {
    const int sr = R >= 3, sc = C >= 3, r = R - 3 * sr, q = C - 3 * sc;
    if (!sr && !sc) { idx_dg = idx_off = gp_sym6(r, q); sg = 1 | 2; }
    else if (sr && sc) { idx_dg = 6 + gp_sym6(r, q); idx_off = 12 + gp_sym6(r, q); sg = 0; }
    else if (!sr) { idx_dg = idx_off = 18 + 3 * r + q; sg = 2 | 4; }       // translation row, rotation column: sign of the row end
    else { idx_dg = idx_off = 18 + 3 * q + r; sg = 1 | 4; }                // rotation row, translation column: sign of the column end
}
__device__ __forceinline__ double gp_flip(double v, unsigned mask) {
    return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}
// rolled members are rare: the double-precision sin / cos (whose slow path owns a stack frame) stays out of line
__device__ __noinline__ double2 gp_cos_sin(double roll) { return make_double2(cos(roll), sin(roll)); }
__device__ __forceinline__ void st_global_v2(double* p, double a, double b) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

template <int MINB>
__global__ void __launch_bounds__(GP_THREADS, MINB) k_assemble_gather_plain(GatherArgs G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BeamArgs& A = G.B;
    double* s_val = reinterpret_cast<double*>(smem_raw);                 // [GP_THREADS][GP_VALS]
    double* s_prop = s_val + GP_THREADS * GP_VALS;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_prop + (size_t)(A.use_tma ? A.n_prop : 0) * NPROP);
    __shared__ int s_hp[GP_NODES + 1];
    __shared__ int s_meta[GP_THREADS];         // slot of the off-diagonal block << 2 | not written here << 1 | end

    const int tid = threadIdx.x;
    if (A.use_tma) {
        if (tid == 0) mbar_init(s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t prop_bytes = (uint32_t)A.n_prop * NPROP * sizeof(double);
            mbar_expect_tx(s_bar, prop_bytes);
            tma_bulk_g2s(s_prop, A.prop, prop_bytes, s_bar);
        }
    }
    const int node_a = blockIdx.x * GP_NODES;
    const int nn = min(GP_NODES, G.n_nodes - node_a);
    if (tid <= nn) s_hp[tid] = G.hm_ptr[node_a + tid];
    __syncthreads();
    const int h_begin = s_hp[0], h_end = s_hp[nn];
    const double* prop_tab = A.use_tma ? s_prop : A.prop;
    bool prop_ready = !A.use_tma;

    const int g = tid / GP_LANES, pr = tid - GP_LANES * g;
    // diagonal sums: written straight out when the CTA's half-members fit one chunk (always, on frames with up to six members
    // per node); carried from chunk to chunk in shared memory otherwise -- no accumulator lives across the register-heavy phase A
    __shared__ double s_acc[GP_NODES * 36];
    const bool one_chunk = h_end - h_begin <= GP_THREADS;
    if (!one_chunk)
        for (int i = tid; i < GP_NODES * 36; i += GP_THREADS) s_acc[i] = 0.0;

    for (int chunk0 = h_begin; chunk0 < h_end; chunk0 += GP_THREADS) {
        const int h = chunk0 + tid;
        if (h < h_end) {
            const int code = G.hm_list[h];
            const int m = code >> 2, end = code & 1;
            const int4 slots = reinterpret_cast<const int4*>(A.m_slot)[m];
            const int off_slot = end ? slots.z : slots.y, dg_slot = end ? slots.w : slots.x;
            const int a = A.n1[m], b = A.n2[m];
            const double roll = A.m_roll[m];
            const int mtype = A.m_type[m], pi = A.m_prop[m];
            const double scale = A.m_scale ? A.m_scale[m] : 1.0;
            // local frame (the rule of member_frame) and closed-form coefficients (those of beam_coef) on one reciprocal
            // of the length: the divisions of the shared functions cost more than everything else in this kernel
            const double dx = A.xyz[3 * b] - A.xyz[3 * a], dy = A.xyz[3 * b + 1] - A.xyz[3 * a + 1], dz = A.xyz[3 * b + 2] - A.xyz[3 * a + 2];
            const double iL = rsqrt(dx * dx + dy * dy + dz * dz);
            double x[3] = {dx * iL, dy * iL, dz * iL}, y[3], z[3];
            {
                double rf1 = 1.0, rf2 = 0.0;
                if (fabs(x[1]) > 1.0 - 1e-6) { rf1 = 0.0; rf2 = 1.0; }
                z[0] = x[1] * rf2 - x[2] * rf1; z[1] = -x[0] * rf2; z[2] = x[0] * rf1;
                const double zn = rsqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
                z[0] *= zn; z[1] *= zn; z[2] *= zn;
                y[0] = z[1] * x[2] - z[2] * x[1]; y[1] = z[2] * x[0] - z[0] * x[2]; y[2] = z[0] * x[1] - z[1] * x[0];
                const double yn = rsqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
                y[0] *= yn; y[1] *= yn; y[2] *= yn;
                if (roll != 0.0) {
                    const double2 cs = gp_cos_sin(roll);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double ny = cs.x * y[i] + cs.y * z[i], nz = -cs.y * y[i] + cs.x * z[i];
                        y[i] = ny; z[i] = nz;
                    }
                }
            }
            if (!prop_ready) { mbar_wait(s_bar, 0); prop_ready = true; }
            const double* P = prop_tab + (size_t)pi * NPROP;
            const double iL2 = iL * iL, iL3 = iL2 * iL;
            const bool axial_only = (mtype == MT_TRUSS);
            double bz = 0.0, by = 0.0, phi_y = 0.0, phi_z = 0.0;
            if (!axial_only) {
                bz = P[P_EIZ]; by = P[P_EIY];
                if (P[P_GASZ] > 0.0) { phi_y = 12.0 * bz * iL2 / P[P_GASZ]; bz = bz / (1.0 + phi_y); }
                if (P[P_GASY] > 0.0) { phi_z = 12.0 * by * iL2 / P[P_GASY]; by = by / (1.0 + phi_z); }
            }
            const double ca = scale * (P[P_EA] * iL);
            const double ct = axial_only ? 0.0 : scale * (P[P_GJ] * iL + 12.0 * P[P_EIW] * iL3);
            const double z12 = scale * (12.0 * bz * iL3), z6 = scale * (6.0 * bz * iL2);
            const double z4 = scale * ((4.0 + phi_y) * bz * iL), z2 = scale * ((2.0 - phi_y) * bz * iL);
            const double y12 = scale * (12.0 * by * iL3), y6 = scale * (6.0 * by * iL2);
            const double y4 = scale * ((4.0 + phi_z) * by * iL), y2 = scale * ((2.0 - phi_z) * by * iL);
            double* o = s_val + tid * GP_VALS;
            int k = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = i; j < 3; ++j, ++k) {
                    const double xx = x[i] * x[j], yy = y[i] * y[j], zz = z[i] * z[j];
                    o[k] = ca * xx + z12 * yy + y12 * zz;
                    o[6 + k] = ct * xx + y4 * yy + z4 * zz;
                    o[12 + k] = -ct * xx + y2 * yy + z2 * zz;
                }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) o[18 + 3 * i + j] = z6 * y[i] * z[j] - y6 * z[i] * y[j];
            // duplicated node pairs are left to k_gather_dups
            s_meta[tid] = (off_slot << 2) | ((((code & 2) != 0) || off_slot == dg_slot) ? 2 : 0) | end;
        }
        __syncthreads();
        // the two neighbouring entries (R, C0), (R, C0 + 1) of a 6x6 block this lane owns
        const int R = pr / 3, C0 = 2 * (pr - 3 * R);
        int idg0, iof0, sg0, idg1, iof1, sg1;
        gp_entry(R, C0, idg0, iof0, sg0);
        gp_entry(R, C0 + 1, idg1, iof1, sg1);
        const int cnt = min(GP_THREADS, h_end - chunk0);
        double* const out = A.vals + 2 * pr;
#pragma unroll 4
        for (int i = g; i < cnt; i += GP_GROUPS) {
            const int meta = s_meta[i];
            if (!(meta & 2)) {
                const int bit = 1 << (meta & 1);                     // K_se for a start end, K_es for an end end
                const double* v = s_val + i * GP_VALS;
                st_global_v2(out + (size_t)(meta >> 2) * 36, gp_flip(v[iof0], (sg0 & bit) ? 0x80000000u : 0u),
                             gp_flip(v[iof1], (sg1 & bit) ? 0x80000000u : 0u));
            }
        }
#pragma unroll
        for (int k = 0; k < GP_QS; ++k) {
            const int nl = g + GP_GROUPS * k;
            if (nl < nn) {
                const int lo = max(s_hp[nl], chunk0) - chunk0, hi = min(s_hp[nl + 1], chunk0 + GP_THREADS) - chunk0;
                double a0 = one_chunk ? 0.0 : s_acc[nl * 36 + 2 * pr], a1 = one_chunk ? 0.0 : s_acc[nl * 36 + 2 * pr + 1];
                for (int i = lo; i < hi; ++i) {
                    const unsigned ee = (s_meta[i] & 1) ? 4u : 0u;   // K_ee for an end end, K_ss (no sign) for a start end
                    const double* v = s_val + i * GP_VALS;
                    a0 += gp_flip(v[idg0], (sg0 & ee) ? 0x80000000u : 0u);
                    a1 += gp_flip(v[idg1], (sg1 & ee) ? 0x80000000u : 0u);
                }
                if (one_chunk || chunk0 + GP_THREADS >= h_end)        // the node's list is complete
                    st_global_v2(A.vals + (size_t)G.diag_slot[node_a + nl] * 36 + 2 * pr, a0, a1);
                else { s_acc[nl * 36 + 2 * pr] = a0; s_acc[nl * 36 + 2 * pr + 1] = a1; }
            }
        }
        __syncthreads();
    }
    if (!prop_ready) mbar_wait(s_bar, 0);     // never leave the CTA with a bulk copy in flight
    if (h_begin >= h_end) {                   // nodes without members: zero diagonal blocks
#pragma unroll
        for (int k = 0; k < GP_QS; ++k) {
            const int nl = g + GP_GROUPS * k;
            if (nl < nn) st_global_v2(A.vals + (size_t)G.diag_slot[node_a + nl] * 36 + 2 * pr, 0.0, 0.0);
        }
    }
}

static int g_gather_kernel = 0;      // 0: element-lane kernel for plain models (default); 1: always k_assemble_gather

// Off-diagonal blocks shared by several members (flagged dup): the first half-member of each run of
// equal slots forms every contribution of the run and stores the sum.  Launched only when the map has
// duplicates.
template <bool GENERAL>
__global__ void __launch_bounds__(128) k_gather_dups(GatherArgs G) {
    const BeamArgs& A = G.B;
    const int n_half = 2 * A.n_members;
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_half) return;
    const int code = G.hm_list[h];
    if (!(code & 2)) return;
    const int m = code >> 2, end = code & 1;
    const int4 sl = reinterpret_cast<const int4*>(A.m_slot)[m];
    const int off_slot = end ? sl.z : sl.y, dg_slot = end ? sl.w : sl.x;
    if (off_slot == dg_slot) return;
    auto slot_of = [&](int hh) {
        const int c2 = G.hm_list[hh];
        const int4 s2 = reinterpret_cast<const int4*>(A.m_slot)[c2 >> 2];
        return (c2 & 1) ? s2.z : s2.y;
    };
    if (h > 0 && slot_of(h - 1) == off_slot) return;          // not the first of its run
    double sum[36], blk[36];
    member_block<GENERAL>(A, A.prop, m, end, 1 - end, sum);
    for (int h2 = h + 1; h2 < n_half && slot_of(h2) == off_slot; ++h2) {
        const int c2 = G.hm_list[h2];
        member_block<GENERAL>(A, A.prop, c2 >> 2, c2 & 1, 1 - (c2 & 1), blk);
        for (int e = 0; e < 36; ++e) sum[e] += blk[e];
    }
    for (int e = 0; e < 36; ++e) A.vals[(size_t)off_slot * 36 + e] = sum[e];
}

// ---- gather map: half-members sorted by the block slot of their off-diagonal block ---------------
__global__ void k_gather_keys(int n_members, const int* __restrict__ n1, const int* __restrict__ n2,
                              const int* __restrict__ m_slot, unsigned long long* __restrict__ keys, int* __restrict__ count) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < 2 * n_members; k += gridDim.x * blockDim.x) {
        const int m = k >> 1, end = k & 1;
        const int4 s = reinterpret_cast<const int4*>(m_slot)[m];
        const unsigned long long slot = (unsigned)(end ? s.z : s.y);
        keys[k] = (slot << 32) | (unsigned)k;
        atomicAdd(count + (end ? n2[m] : n1[m]), 1);
    }
}

__global__ void k_gather_list(int n_half, const unsigned long long* __restrict__ sorted, int* __restrict__ hm_list,
                              int* __restrict__ n_dup) {
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < n_half; h += gridDim.x * blockDim.x) {
        const unsigned long long key = sorted[h];
        const unsigned slot = (unsigned)(key >> 32), k = (unsigned)(key & 0xffffffffull);
        const bool dup = (h > 0 && (unsigned)(sorted[h - 1] >> 32) == slot) || (h + 1 < n_half && (unsigned)(sorted[h + 1] >> 32) == slot);
        hm_list[h] = (int)(((k >> 1) << 2) | (dup ? 2u : 0u) | (k & 1u));
        if (dup) atomicAdd(n_dup, 1);
    }
}

// ---- plates: one thread per TRI3, nine 6x6 blocks, same staged scatter -------------------------
__global__ void __launch_bounds__(ASM_THREADS) k_assemble_plates(int n_plates, const double* __restrict__ xyz,
                                                                 const int* __restrict__ pn,
                                                                 const double* __restrict__ pdata,
                                                                 const int* __restrict__ pflags,
                                                                 const int* __restrict__ p_slot,
                                                                 double* __restrict__ vals) {
    __shared__ double s_stage[ASM_WARPS * STAGE_DOUBLES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = blockIdx.x * ASM_THREADS + tid;
    const bool valid = e < n_plates;
    double k[324];
    PlateGeom g;
    if (valid) {
        const int a = pn[3 * e], b = pn[3 * e + 1], c = pn[3 * e + 2];
        double p0[3] = {xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2]};
        double p1[3] = {xyz[3 * b], xyz[3 * b + 1], xyz[3 * b + 2]};
        double p2[3] = {xyz[3 * c], xyz[3 * c + 1], xyz[3 * c + 2]};
        g = plate_geom(p0, p1, p2, pdata[NPDATA * (size_t)e + PD_THETA]);
        plate_local_k(g, pdata + NPDATA * (size_t)e, pflags[e], k);
    }
    double* stage = s_stage + warp * STAGE_DOUBLES;
#pragma unroll 1
    for (int b = 0; b < 9; ++b) {
        double blk[36];
        if (valid) {
            plate_rotate_block(k, g, b / 3, b % 3, blk);
        } else {
#pragma unroll
            for (int q = 0; q < 36; ++q) blk[q] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < 36; ++q) stage[q * STAGE_STRIDE + lane] = blk[q];
        const int slot = valid ? p_slot[9 * e + b] : -1;
        __syncwarp();
        for (int w = lane; w < 32 * 36; w += 32) {
            const int mm = w / 36, q = w - mm * 36;
            const int s = __shfl_sync(FULL, slot, mm);
            if (s >= 0) atomicAdd(vals + (size_t)s * 36 + q, stage[q * STAGE_STRIDE + mm]);
        }
        __syncwarp();
    }
}

}  // namespace fers

using namespace fers;

extern "C" {

int fers_b200_assemble_beams(int n_members, const double* xyz, const int* m_n1, const int* m_n2, const int* m_prop,
                             const int* m_type, const double* m_roll, const double* m_scale, const int* m_rel,
                             const double* prop, int n_prop, const double* rel, int n_rel, const double* m_axial,
                             int kg_mode, int kg_suppress, const int* m_slot, double* vals, int variant, void* stream) {
    (void)n_rel;
    if (n_members <= 0) return 0;
    FERS_REQUIRE(n_prop > 0, "assemble_beams: empty property table");
    BeamArgs A;
    A.n_members = n_members; A.xyz = xyz; A.n1 = m_n1; A.n2 = m_n2; A.m_prop = m_prop; A.m_type = m_type;
    A.m_roll = m_roll; A.m_scale = m_scale; A.m_rel = m_rel; A.prop = prop; A.n_prop = n_prop; A.rel = rel;
    A.m_axial = m_axial; A.kg_mode = kg_mode; A.kg_suppress = kg_suppress; A.m_slot = m_slot; A.vals = vals;
    A.aggregate = (variant & 1) ? 0 : 1;
    A.use_tma = ((variant & 2) == 0 && n_prop <= MAX_PROP_SMEM_ROWS && (((uintptr_t)prop) & 15) == 0) ? 1 : 0;
    size_t smem = (size_t)ASM_WARPS * STAGE_DOUBLES * sizeof(double) +
                  (A.use_tma ? (size_t)n_prop * NPROP * sizeof(double) : 0) + 16;
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        const int max_smem = (int)((size_t)ASM_WARPS * STAGE_DOUBLES * 8 + MAX_PROP_SMEM_ROWS * NPROP * 8 + 16);
        FERS_CK(cudaFuncSetAttribute(k_assemble_beams<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FERS_CK(cudaFuncSetAttribute(k_assemble_beams<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set.done();
    }
    int grid = (n_members + ASM_THREADS - 1) / ASM_THREADS;
    const bool general = (m_rel != nullptr) || (m_axial != nullptr && kg_mode != 0);
    if (general) k_assemble_beams<true><<<grid, ASM_THREADS, smem, as_stream(stream)>>>(A);
    else k_assemble_beams<false><<<grid, ASM_THREADS, smem, as_stream(stream)>>>(A);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

size_t fers_b200_gather_map_workspace(int n_nodes, int n_members) {
    size_t sort_tmp = 0, scan_tmp = 0;
    const int nh = 2 * (n_members > 0 ? n_members : 1);
    cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, nh);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int*)nullptr, (int*)nullptr, n_nodes + 1);
    const size_t tmp = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
    return 2 * align_up((size_t)nh * 8) + align_up(((size_t)n_nodes + 1) * 4) + align_up(tmp) + 512;
}

int fers_b200_gather_map_build(int n_nodes, int n_members, int nnzb, const int* m_n1, const int* m_n2, const int* m_slot,
                               int* hm_ptr, int* hm_list, int* n_dup_out, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = as_stream(stream);
    FERS_REQUIRE(n_nodes > 0 && hm_ptr, "gather_map_build: bad arguments");
    FERS_REQUIRE(n_members < (1 << 29), "gather_map_build: too many members for the packed half-member code");
    FERS_REQUIRE(workspace_bytes >= fers_b200_gather_map_workspace(n_nodes, n_members), "gather_map_build: workspace too small");
    if (n_members <= 0) {
        FERS_CK(cudaMemsetAsync(hm_ptr, 0, ((size_t)n_nodes + 1) * 4, st));
        if (n_dup_out) *n_dup_out = 0;
        return 0;
    }
    const int nh = 2 * n_members;
    char* w = (char*)workspace;
    unsigned long long* keys = (unsigned long long*)w;   w += align_up((size_t)nh * 8);
    unsigned long long* sorted = (unsigned long long*)w; w += align_up((size_t)nh * 8);
    int* count = (int*)w;                                w += align_up(((size_t)n_nodes + 1) * 4);
    int* d_dup = (int*)w;                                w += 256;
    size_t tmp_bytes = workspace_bytes - (size_t)(w - (char*)workspace);
    FERS_CK(cudaMemsetAsync(count, 0, ((size_t)n_nodes + 1) * 4, st));
    FERS_CK(cudaMemsetAsync(d_dup, 0, 4, st));
    const int T = 256, grid = (nh + T - 1) / T;
    k_gather_keys<<<grid, T, 0, st>>>(n_members, m_n1, m_n2, m_slot, keys, count);
    FERS_LAUNCH_CHECK();
    int bits = 1;
    while ((1ll << bits) < (long long)(nnzb > 1 ? nnzb : 2)) ++bits;
    size_t t1 = tmp_bytes;
    FERS_CK(cub::DeviceRadixSort::SortKeys(w, t1, keys, sorted, nh, 0, 32 + bits, st));
    size_t t2 = tmp_bytes;
    FERS_CK(cub::DeviceScan::ExclusiveSum(w, t2, count, hm_ptr, n_nodes + 1, st));
    k_gather_list<<<grid, T, 0, st>>>(nh, sorted, hm_list, d_dup);
    FERS_LAUNCH_CHECK();
    count_launches(2 + 5);
    int n_dup = 0;
    FERS_CK(cudaMemcpyAsync(&n_dup, d_dup, 4, cudaMemcpyDeviceToHost, st));
    FERS_CK(cudaStreamSynchronize(st));
    if (n_dup_out) *n_dup_out = n_dup;
    return 0;
}

int fers_b200_assemble_beams_gather(int n_nodes, int n_members, const double* xyz, const int* m_n1, const int* m_n2,
                                    const int* m_prop, const int* m_type, const double* m_roll, const double* m_scale,
                                    const int* m_rel, const double* prop, int n_prop, const double* rel, int n_rel,
                                    const double* m_axial, int kg_mode, int kg_suppress, const int* m_slot,
                                    const int* hm_ptr, const int* hm_list, int n_dup, const int* diag_slot,
                                    double* member_cache, double* vals, void* stream) {
    (void)n_rel;
    FERS_REQUIRE(n_nodes > 0 && hm_ptr && diag_slot && vals, "assemble_beams_gather: bad arguments");
    FERS_REQUIRE(n_members <= 0 || (n_prop > 0 && hm_list && m_slot), "assemble_beams_gather: empty property table or missing maps");
    GatherArgs G;
    BeamArgs& A = G.B;
    A.n_members = n_members; A.xyz = xyz; A.n1 = m_n1; A.n2 = m_n2; A.m_prop = m_prop; A.m_type = m_type;
    A.m_roll = m_roll; A.m_scale = m_scale; A.m_rel = m_rel; A.prop = prop; A.n_prop = n_prop; A.rel = rel;
    A.m_axial = m_axial; A.kg_mode = kg_mode; A.kg_suppress = kg_suppress; A.m_slot = m_slot; A.vals = vals;
    A.aggregate = 0;
    A.use_tma = (n_members > 0 && n_prop <= MAX_PROP_SMEM_ROWS && (((uintptr_t)prop) & 15) == 0) ? 1 : 0;
    G.n_nodes = n_nodes; G.hm_ptr = hm_ptr; G.hm_list = hm_list; G.diag_slot = diag_slot; G.cache = member_cache;
    const bool cached = member_cache != nullptr && n_members > 0;
    const size_t prop_smem = A.use_tma ? (size_t)n_prop * NPROP * sizeof(double) : 0;
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        const int max_smem = (int)((size_t)GA_THREADS * GA_STRIDE * 8 + MAX_PROP_SMEM_ROWS * NPROP * 8 + 16);
        FERS_CK(cudaFuncSetAttribute(k_assemble_gather<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FERS_CK(cudaFuncSetAttribute(k_assemble_gather<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FERS_CK(cudaFuncSetAttribute(k_assemble_gather<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FERS_CK(cudaFuncSetAttribute(k_assemble_gather<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FERS_CK(cudaFuncSetAttribute(k_assemble_gather_plain<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)GP_THREADS * GP_VALS * 8 + MAX_PROP_SMEM_ROWS * NPROP * 8 + 16)));
        FERS_CK(cudaFuncSetAttribute(k_assemble_gather_plain<6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)GP_THREADS * GP_VALS * 8 + MAX_PROP_SMEM_ROWS * NPROP * 8 + 16)));
        attr_set.done();
    }
    if (cached) {
        FERS_REQUIRE((((uintptr_t)member_cache) & 31) == 0, "assemble_beams_gather: member_cache must be 32-byte aligned");
        k_member_cache<<<(n_members + 255) / 256, 256, prop_smem + 16, as_stream(stream)>>>(A, member_cache);
        FERS_LAUNCH_CHECK();
        count_launches(1);
        A.use_tma = 0;       // the gather kernel reads frames and coefficients from the cache
    }
    const size_t smem = (size_t)GA_THREADS * GA_STRIDE * sizeof(double) + (A.use_tma ? prop_smem : 0) + 16;
    const int grid = (n_nodes + GA_NODES - 1) / GA_NODES;
    const bool general = (m_rel != nullptr) || (m_axial != nullptr && kg_mode != 0);
    if (!general && !cached && g_gather_kernel == 0 && n_members > 0) {
        const size_t smem_p = (size_t)GP_THREADS * GP_VALS * sizeof(double) + prop_smem + 16;
        // 5 CTAs per SM at 72 registers beat 6 at 64 (73 against 79 us at configs[3]: the spills of the tighter budget go
        // through local memory, which misses L1 under this kernel's shared-memory carve-out); FERS_GATHER_MINB=6 for the other
        static const int minb = getenv("FERS_GATHER_MINB") ? atoi(getenv("FERS_GATHER_MINB")) : 5;
        if (minb == 5) k_assemble_gather_plain<5><<<(n_nodes + GP_NODES - 1) / GP_NODES, GP_THREADS, smem_p, as_stream(stream)>>>(G);
        else k_assemble_gather_plain<6><<<(n_nodes + GP_NODES - 1) / GP_NODES, GP_THREADS, smem_p, as_stream(stream)>>>(G);
    } else if (cached) {
        if (general) k_assemble_gather<true, true><<<grid, GA_THREADS, smem, as_stream(stream)>>>(G);
        else k_assemble_gather<false, true><<<grid, GA_THREADS, smem, as_stream(stream)>>>(G);
    } else {
        if (general) k_assemble_gather<true, false><<<grid, GA_THREADS, smem, as_stream(stream)>>>(G);
        else k_assemble_gather<false, false><<<grid, GA_THREADS, smem, as_stream(stream)>>>(G);
    }
    FERS_LAUNCH_CHECK();
    count_launches(1);
    if (n_dup > 0 && n_members > 0) {
        const int g2 = (2 * n_members + 127) / 128;
        if (general) k_gather_dups<true><<<g2, 128, 0, as_stream(stream)>>>(G);
        else k_gather_dups<false><<<g2, 128, 0, as_stream(stream)>>>(G);
        FERS_LAUNCH_CHECK();
        count_launches(1);
    }
    return 0;
}

int fers_b200_assemble_set_gather_kernel(int kind) {
    FERS_REQUIRE(kind == 0 || kind == 1, "assemble_set_gather_kernel: 0 = element-lane kernel for plain models, 1 = per-half-member kernel");
    g_gather_kernel = kind;
    return 0;
}

int fers_b200_assemble_plates(int n_plates, const double* xyz, const int* p_nodes, const double* p_data,
                              const int* p_flags, const int* p_slot, double* vals, void* stream) {
    if (n_plates <= 0) return 0;
    int grid = (n_plates + ASM_THREADS - 1) / ASM_THREADS;
    k_assemble_plates<<<grid, ASM_THREADS, 0, as_stream(stream)>>>(n_plates, xyz, p_nodes, p_data, p_flags, p_slot, vals);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
