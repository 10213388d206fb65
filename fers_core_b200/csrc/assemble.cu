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
        f = member_frame(pa, pb, A.m_roll[m]);
        mtype = A.m_type[m];
        pi = A.m_prop[m];
        if (A.m_scale) scale = A.m_scale[m];
        if (A.m_rel) ri = A.m_rel[m];
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
        // releases condense the elastic part only (K_g is added un-condensed, as the oracle does)
        if (ri >= 0) beam_condense(kE, A.rel + (size_t)ri * 12, nullptr, 0);
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
        g = plate_geom(p0, p1, p2);
        plate_local_k(g, pdata + 4 * (size_t)e, pflags[e], k);
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
    static bool attr_set = false;
    if (!attr_set) {
        const int max_smem = (int)((size_t)ASM_WARPS * STAGE_DOUBLES * 8 + MAX_PROP_SMEM_ROWS * NPROP * 8 + 16);
        FERS_CK(cudaFuncSetAttribute(k_assemble_beams<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        FERS_CK(cudaFuncSetAttribute(k_assemble_beams<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set = true;
    }
    int grid = (n_members + ASM_THREADS - 1) / ASM_THREADS;
    const bool general = (m_rel != nullptr) || (m_axial != nullptr && kg_mode != 0);
    if (general) k_assemble_beams<true><<<grid, ASM_THREADS, smem, as_stream(stream)>>>(A);
    else k_assemble_beams<false><<<grid, ASM_THREADS, smem, as_stream(stream)>>>(A);
    FERS_LAUNCH_CHECK();
    count_launches(1);
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
