/* fers_b200.h -- C ABI of libfers_b200.so, the sm_100a implementation of the FERS analysis hot path.
 *
 * What this boundary replaces.  The reference front end (Jeroen124/FERS_core) hands the whole
 * analysis to one native call, `fers_calculations.calculate_from_json(str) -> str`
 * (fers_core/fers/fers.py:300; file variant fers.py:260; direct use tests/nafems/nafems_harness.py:38).
 * The Python module `fers_calculations/` in this repo keeps exactly that signature; below it the
 * work is split into the stages the reference engine documents (element stiffness -> assembly ->
 * boundary conditions -> "batch first-order solve that shares one factorization across many load
 * cases" -> recovery; fers_core/types/pydantic_models.py:1200-1219, 1334-1398, 1594-1635) and
 * each stage is one entry point here.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; fers_b200_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - pointers are DEVICE pointers unless the parameter is marked [host];
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); calls are asynchronous on
 *     that stream unless they return a value to the host (documented per function);
 *   - no allocation happens inside the library except cub scratch carved from caller workspace;
 *   - multi-vector layout is [n_dof][nrhs] (right-hand side index fastest), n_dof = 6 * n_nodes,
 *     DOF order per node = X Y Z RX RY RZ (fers_core/supports/nodalsupport.py:146-159);
 *   - the stiffness matrix is block-CSR with 6x6 node blocks, row-major inside a block.
 */
#ifndef FERS_B200_H
#define FERS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FERS_B200_ABI_VERSION 1

/* ---- library -------------------------------------------------------------------------------- */
const char* fers_b200_version(void);
const char* fers_b200_last_error(void);
/* [host] outputs.  Fails when `device` is not a CUDA device of compute capability 10.x. */
int fers_b200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
/* Make `device` current for this library's CUDA runtime (one process per GPU: call once with
 * LOCAL_RANK before any other entry point). */
int fers_b200_set_device(int device);
/* Number of kernels this library has launched in this process (monotonic). */
long long fers_b200_launch_count(void);

/* ---- [host] bulk JSON writer of the result document (fers_core/fers/fers.py:300-311: the boundary returns text) ----
 * '{' "key_r": <tmpl with every byte 0x01 (number) / 0x02 (integer) replaced by values[r][slot]> , ... '}' for n_records
 * equally-shaped records.
 * keys: n_records NUL-terminated strings back to back; numbers in shortest round-trip form, integers as "5.0", non-finite
 * values as NaN / Infinity / -Infinity (Python's json spelling).  out_cap >= fers_b200_json_records_bound(...). */
size_t fers_b200_json_records_bound(size_t template_len, int n_slots, long long n_records, size_t keys_bytes);
int fers_b200_json_records(const char* tmpl, size_t template_len, int n_slots, long long n_records, const char* keys,
                           const double* values, char* out, size_t out_cap, size_t* written_out);

/* Live fp64 roofline denominators (bench.py): register-resident throughput of the fp64 tensor pipe
 * (DMMA.8x8x4) and of the vector pipe (DFMA) in TFLOP/s, each the best of bursts lasting at least
 * min_ms.  scratch: device buffer of >= 8 * 1024 * SM-count bytes.  Synchronises the stream. */
int fers_b200_fp64_peak(double min_ms, double* dmma_tflops, double* dfma_tflops, double* scratch, size_t scratch_bytes,
                        void* stream);

/* ---- symbolic pattern (replaces the engine's DOF numbering / sparsity set-up; contract
 *      SURVEY.md 8a row A7, pydantic_models.py:2104-2115) ------------------------------------- */
size_t fers_b200_pattern_workspace(int n_nodes, int n_members, int n_plates);
/* Sorted unique node pairs (every diagonal present) -> rowptr[n_nodes+1], colidx[<=cap];
 * per-element block slots: m_slot[n_members*4] = (s,s)(s,e)(e,s)(e,e), p_slot[n_plates*9],
 * diag_slot[n_nodes].  *nnzb_out [host] receives the block count (synchronises the stream). */
int fers_b200_pattern_build(int n_nodes, int n_members, const int* m_n1, const int* m_n2,
                            int n_plates, const int* p_nodes,
                            int* rowptr, int* colidx, int cap,
                            int* m_slot, int* p_slot, int* diag_slot, int* nnzb_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- boundary conditions (pydantic_models.py:1956-1959, 1232-1237) -------------------------- */
/* bc_code[n_dof] (0 free, 1 fixed, 2 spring) -> fixed[n_dof] (uint8 0/1) and
 * free_index[n_dof] (running number of free DOFs, -1 when constrained).  DOFs whose diagonal
 * stiffness (vals + spring) is exactly zero are restrained too ("dead" DOFs, e.g. rotations of
 * truss-only nodes).  *n_free_out, *n_dead_out [host] (synchronises the stream); pass both as NULL to skip
 * the read-back and stay asynchronous (re-assemblies inside an iteration). */
size_t fers_b200_bc_workspace(int n_dof);
int fers_b200_bc_build(int n_nodes, const uint8_t* bc_code, const double* bc_k, const double* vals,
                       const int* diag_slot, uint8_t* fixed, int* free_index, double* kspring,
                       int* n_free_out, int* n_dead_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- element stiffness + assembly --------------------------------- */
/* Beam members: 12x12 local stiffness (Euler-Bernoulli / Timoshenko), optional end-release
 * condensation and geometric stiffness, T^T k T, scatter-add into vals[nnzb*36].
 *   m_scale  : per-member stiffness multiplier or NULL (active-set switch of tension-only members)
 *   m_rel    : per-member row of rel[n_rel*20] or -1; NULL = none.  One record = 12 end-release spring
 *              stiffnesses in local DOF order (< 0: rigid), the shear-centre offset y_s, z_s
 *              (pydantic_models.py:1106-1111, 1358-1361: k <- S^T k S before the condensation) and the
 *              start / end offsets in global X Y Z (MemberEndOffset, 525-531: the elastic axis runs between
 *              the offset points, tied to the nodes by rigid arms, k <- E^T k E after the condensation)
 *   m_axial  : per-member axial force (tension +) for K_g or NULL; kg_mode 0 none, 1 consistent,
 *              2 simplified; kg_suppress = bit mask of global axes removed from K_g (bit0 X..)
 *   variant  : 0 = warp-aggregated atomic scatter-add, 1 = plain atomic scatter-add
 * vals must be zeroed by the caller (fers_b200_zero).                                           */
int fers_b200_assemble_beams(int n_members, const double* xyz, const int* m_n1, const int* m_n2,
                             const int* m_prop, const int* m_type, const double* m_roll,
                             const double* m_scale, const int* m_rel, const double* prop, int n_prop,
                             const double* rel, int n_rel, const double* m_axial, int kg_mode, int kg_suppress,
                             const int* m_slot, double* vals, int variant, void* stream);
/* Gather formulation of the same assembly (default of the engine): no atomics and no zero pass, every
 * value of K is written exactly once and the result is bit-reproducible.  The map lists the
 * "half-members" (member, end) sorted by the slot of their off-diagonal block, i.e. grouped by the
 * block row they write: hm_ptr[n_nodes+1] (CSR over nodes), hm_list[2*n_members] = member<<2 | dup<<1 | end
 * (dup: another member joins the same node pair; *n_dup_out [host] = how many are flagged).  Built once per
 * pattern from m_slot; synchronises the stream. */
size_t fers_b200_gather_map_workspace(int n_nodes, int n_members);
int fers_b200_gather_map_build(int n_nodes, int n_members, int nnzb, const int* m_n1, const int* m_n2, const int* m_slot,
                               int* hm_ptr, int* hm_list, int* n_dup_out, void* workspace, size_t workspace_bytes, void* stream);
/* Writes every diagonal block and every off-diagonal block that has a member (blocks coupled only by
 * plates are not touched: zero vals first when the model has plates, then add the plates).
 * member_cache: scratch of n_members * 24 doubles (32-byte aligned) or NULL.  With it the local frame and
 * the stiffness coefficients of a member are formed once (k_member_cache) and read by both of its ends;
 * without it every half-member forms them itself. */
int fers_b200_assemble_beams_gather(int n_nodes, int n_members, const double* xyz, const int* m_n1, const int* m_n2,
                                    const int* m_prop, const int* m_type, const double* m_roll, const double* m_scale,
                                    const int* m_rel, const double* prop, int n_prop, const double* rel, int n_rel,
                                    const double* m_axial, int kg_mode, int kg_suppress, const int* m_slot,
                                    const int* hm_ptr, const int* hm_list, int n_dup, const int* diag_slot,
                                    double* member_cache, double* vals, void* stream);
/* Which kernel the gather assembly runs for plain models (no modifier records, no K_g, no member cache): 0 (default) = the
 * element-lane kernel k_assemble_gather_plain (27 values per half-member in shared memory, one lane per entry of a 6x6
 * block), 1 = the per-half-member kernel k_assemble_gather that also serves the general case.  Same K either way. */
int fers_b200_assemble_set_gather_kernel(int kind);
/* Restrained-warping torsion (7th nodal DOF of the reference's beam: pydantic_models.py:1366-1369, 1384-1387;
 * formulation SURVEY.md A.11, pinned by NAFEMS LE5).  The warping DOF psi of a node is DOF 0 of an auxiliary node;
 * prop[.][7] = E*Iw (its theta-theta part 12 EIw/L^3 is already in the member's torsion coefficient).  Adds the
 * psi couplings of n_w warping members to vals: w_member[n_w] member indices, w_slot[n_w][5][4] = the m_slot rows
 * of the node pairs (s,s'), (s,e'), (e,s'), (e,e'), (s',e') (s', e' = auxiliary nodes of start / end). */
int fers_b200_assemble_warping(int n_w, const int* w_member, const double* xyz, const int* m_n1, const int* m_n2,
                               const int* m_prop, const double* m_scale, const double* prop, const int* w_slot, double* vals,
                               void* stream);
/* Elastic catenary cables (MemberType.Cable, pydantic_models.py:592-598; pretension 1525-1528, unstretched_length
 * 1542-1545): one Newton step of the cable state, formed on the host (fers_core_b200/cable.py), into the device matrix
 * and right-hand side.  c_member[n_cables] member rows; g[n_cables][3][3] symmetric tangent between the translations of
 * the end nodes: K(a,a) += G, K(a,b) -= G, K(b,a) -= G, K(b,b) += G; corr[n_cables][6] added to column `col` of
 * f[Ndof][nrhs] at the translations of (start, end). */
int fers_b200_assemble_cables(int n_cables, const int* c_member, const double* g, const int* m_slot, double* vals,
                              void* stream);
int fers_b200_cable_rhs(int n_cables, const int* c_member, const int* m_n1, const int* m_n2, const double* corr, double* f,
                        int nrhs, int col, void* stream);
/* TRI3 shell elements: CST membrane + DSG3 Mindlin bending (pydantic_models.py:1700-1726).
 * p_data[n_plates][8] = thickness, E, nu, stiffness modifiers (membrane, bending, shear), in-plane angle (rad) of local x from the default frame (PlateElement.local_x_direction, pydantic_models.py:1703), 1 spare;
 * p_flags bit0 membrane, bit1 bending, bit2 plane strain (membrane part). */
int fers_b200_assemble_plates(int n_plates, const double* xyz, const int* p_nodes, const double* p_data,
                              const int* p_flags, const int* p_slot, double* vals, void* stream);
int fers_b200_zero(void* ptr, size_t bytes, void* stream);

/* ---- load vectors (pydantic_models.py:2048-2059, 491-505) ---------- */
/* One load case: nodal loads, member span loads (-> consistent equivalent nodal loads) and plate
 * pressures accumulate into f_case[n_dof] and peq_case[n_members*12] (local frame, after release
 * condensation).  Both must be zeroed by the caller. */
int fers_b200_load_case(int n_nodes, int n_members, const double* xyz, const int* m_n1, const int* m_n2,
                        const int* m_prop, const int* m_type, const double* m_roll, const double* m_scale,
                        const int* m_rel, const double* prop, const double* rel,
                        int n_nod, const int* nod_node, const double* nod_vec,
                        int n_dl, const int* dl_member, const double* dl_data,
                        int n_pl, const int* pl_member, const double* pl_data,
                        int n_pp, const int* pp_plate, const double* pp_vec, const int* p_nodes,
                        double* f_case, double* peq_case, void* stream);
/* out[n_rows][nrhs] = sum_l factors[rhs][l] * in[l][n_rows]   (factors [nrhs][n_cases], row-major) */
int fers_b200_combine(int n_rows, int n_cases, int nrhs, const double* in, const double* factors,
                      double* out, void* stream);

/* ---- solve ------------------------------------------------------- */
/* y = K x on the free DOFs (rows of constrained DOFs give 0), springs included; nrhs in
 * {1,8,16,32,64}.  fixed == NULL: plain product with the unconstrained matrix (used for
 * reactions R = K u - F). */
/* SpMM kernel variant: 0 = register/LDG kernel (all nrhs), 1 = TMA-staged shared-memory pipeline,
 * 2 = TMA-staged with dedicated producer warps (1 and 2: nrhs = 64 only; other nrhs fall back to 0).  Default from env FERS_SPMM_VARIANT, else 1. */
int fers_b200_set_spmm_variant(int variant);
int fers_b200_spmm(int n_nodes, int nrhs, const int* rowptr, const int* colidx, const double* vals,
                   const uint8_t* fixed, const double* kspring, const double* x, double* y, void* stream);
/* Block-Jacobi preconditioner: inverse of each node's 6x6 diagonal block restricted to its free
 * DOFs -> minv[n_nodes*36]. */
int fers_b200_block_jacobi(int n_nodes, const double* vals, const int* diag_slot, const uint8_t* fixed,
                           const double* kspring, double* minv, void* stream);
size_t fers_b200_pcg_workspace(int n_nodes, int nrhs);
/* [host] Device-side timing of the last fers_b200_pcg_solve, from CUDA events recorded on its
 * stream around a sample of the iterations: out[0..2] = mean ms of the SpMM, update and direction
 * kernels, out[3] = sampled iterations, out[4] = total iterations, out[5] = mean ms of one whole
 * iteration (all five launches). */
int fers_b200_pcg_profile(double* out6);
/* Batched preconditioned CG, one independent recurrence per right-hand side column.
 * b, x: [n_dof][nrhs]; x holds the initial guess when warm_start != 0, else it is zeroed.
 * Stops when every column has ||r|| <= tol * ||b|| or after max_iter iterations.
 * [host] outputs: *iters_out, relres_out[nrhs].  Returns 0 also when not converged (caller
 * inspects relres_out).  Synchronises the stream. */
int fers_b200_pcg_solve(int n_nodes, int nrhs, const int* rowptr, const int* colidx, const double* vals,
                        const uint8_t* fixed, const double* kspring, const double* minv,
                        const double* b, double* x, int warm_start, double tol, int max_iter, int check_every,
                        int* iters_out, double* relres_out, void* workspace, size_t workspace_bytes, void* stream);
/* Dense direct path for small models and batches of small models: expand block-CSR to dense
 * (constrained DOFs -> identity rows), Cholesky, multi-RHS substitution.
 * a: [batch][n][n] column-major lower factor on exit; bx: [batch][n][nrhs] in: b, out: x.
 * info_out [host][batch]: 0 ok, k>0 = leading minor k not positive definite. Synchronises. */
int fers_b200_dense_expand(int n_nodes, const int* rowptr, const int* colidx, const double* vals,
                           const uint8_t* fixed, const double* kspring, double* a, void* stream);
/* Batch of `batch` disconnected frames with nodes_per_frame nodes each (nodes numbered frame by
 * frame): frame f -> a[f][n][n], n = 6*nodes_per_frame.  flag_dev: one device int of scratch.
 * Fails when a block couples two frames.  Synchronises the stream. */
int fers_b200_dense_expand_batched(int nodes_per_frame, int batch, const int* rowptr, const int* colidx,
                                   const double* vals, const uint8_t* fixed, const double* kspring, double* a,
                                   int* flag_dev, void* stream);
int fers_b200_dense_cholesky(int n, int batch, double* a, int* info_dev, int* info_out, void* stream);
int fers_b200_dense_solve(int n, int nrhs, int batch, const double* a, double* bx, void* stream);
/* Batch of frames inside a P-Delta fixed point (BASELINE.json configs[4]; reference: the per-model iteration of
 * pydantic_models.py:1370-1398 run for every frame of the batch).  All three calls are asynchronous (no host
 * synchronisation) and skip the frames whose active[f] is 0 (NULL: every frame):
 *  frame_batch_factor_solve: zero a, expand the block-CSR matrix of every active frame (fers_b200_dense_expand_batched
 *    without its check-and-synchronise: *flag_dev is OR-ed when a block couples two frames, the caller zeroes and reads
 *    it), factorise (one CTA per frame, n = 6 nodes_per_frame <= 512; info_dev[f] = failing minor or 0) and substitute
 *    bx[f][n] in place;
 *  frame_batch_step: r = |un - u| / |un| per active frame; it takes un, counts the iteration, records r and freezes
 *    when r < tol; flags[0] |= a frame is still active, flags[1] |= a frame had a non-positive pivot (bad[f] = 1);
 *  member_axial_masked: axial[m] = f_local[m][6] (one right-hand side) for the members of active frames. */
int fers_b200_frame_batch_factor_solve(int nodes_per_frame, int batch, const int* rowptr, const int* colidx, const double* vals,
                                       const uint8_t* fixed, const double* kspring, double* a, const uint8_t* active,
                                       int* info_dev, int* flag_dev, double* bx, void* stream);
int fers_b200_frame_batch_step(int batch, int n, const double* un, double* u, uint8_t* active, long long* iters, double* resid,
                               double tol, const int* info_dev, uint8_t* bad, int* flags, void* stream);
int fers_b200_member_axial_masked(int n_members, int members_per_frame, const double* f_local, const uint8_t* active,
                                  double* axial, void* stream);

/* ---- sparse direct solve: multifrontal Cholesky (the reference engine's
 *      "batch first-order solve that shares one factorization across many load cases",
 *      pydantic_models.py:1216-1219) -------------------------------------------------------------
 * Analysis phase, host only (no CUDA call): nested dissection on the node coordinates, front tree,
 * symbolic factorisation.  rowptr/colidx [host]: the block-CSR node pattern (structurally symmetric,
 * diagonal present, columns sorted); xyz [host][n_nodes*3].  leaf_nodes: largest undivided node set.
 * node_restrained [host][n_nodes] or NULL: non-zero = all six DOFs of the node are restrained by supports
 * for the lifetime of the plan; such nodes are left out of the elimination (their rows of the solution
 * are 0), which shrinks every front they would sit in.  The `fixed` mask given to the numeric calls must
 * restrain all their DOFs.
 * *plan_out receives an opaque host handle (free with fers_b200_chol_plan_destroy). */
int fers_b200_chol_plan_create(int n_nodes, const int* rowptr, const int* colidx, const double* xyz, int leaf_nodes,
                               const uint8_t* node_restrained, void** plan_out);
void fers_b200_chol_plan_destroy(void* plan);
/* [host] the same block-CSR pattern fers_b200_pattern_build leaves on the device (rowptr[n_nodes+1], colidx[cap >=
 * n_nodes + 2 n_pairs + 6 n_plates], *nnzb_out), built from the host copies of the connectivity: lets the plan of the
 * factorisation start on a host thread while the device is still being set up. */
int fers_b200_host_pattern(int n_nodes, int n_pairs, const int* p1, const int* p2, int n_plates, const int* p_nodes,
                           int* rowptr, int* colidx, int cap, int* nnzb_out);
/* [host] out[12]: fronts, levels, largest front (nodes), factor doubles, Schur arena doubles (even
 * levels, odd levels), front-local rhs rows, index bytes, factorisation Mflop, nnz(L), K blocks
 * scattered, root separator nodes. */
int fers_b200_chol_plan_info(void* plan, long long* out, int n_out);
/* [host] perm_out[n_nodes]: elimination position -> node (-1 past the number of eliminated nodes). */
int fers_b200_chol_plan_perm(void* plan, int* perm_out);
/* [host] out[level*4 + {0,1,2,3}] = fronts, max own nodes, max front nodes, Mflop of the level. */
int fers_b200_chol_plan_levels(void* plan, long long* out, int n_levels_cap);
/* [host] copy of the device index buffer and its section offsets (tests, diagnostics). */
int fers_b200_chol_plan_export(void* plan, void* index_out, long long* offsets_out, int* level_ptr_out,
                               long long* kscatter_ptr_out);
/* Device buffers the caller provides: index (index bytes of plan_info), factor (factor_bytes),
 * work (work_bytes; shared by factorisation and substitution). */
size_t fers_b200_chol_factor_bytes(void* plan);
size_t fers_b200_chol_work_bytes(void* plan, int nrhs);
/* Copy the index arrays to index_dev (synchronises the stream). */
int fers_b200_chol_upload(void* plan, void* index_dev, void* stream);
/* Numeric factorisation K_ff = L L^T of the block-CSR values (constrained DOFs -> identity rows,
 * support springs added, symmetric diagonal equilibration).  info_dev: one device int; info_out
 * [host] or NULL: 0 ok, d > 0 = the d-th unknown in elimination order has a non-positive pivot
 * (mechanism / buckled).  Synchronises the stream when info_out != NULL. */
int fers_b200_chol_factor(void* plan, const void* index_dev, const double* vals, const int* diag_slot,
                          const uint8_t* fixed, const double* kspring, double* factor, void* work, size_t work_bytes,
                          int* info_dev, int* info_out, void* stream);
/* x = K_ff^-1 b (accumulate == 0) or x += K_ff^-1 b (accumulate != 0, iterative refinement);
 * b, x: [n_dof][nrhs], nrhs even, 2..64; rows of constrained DOFs give 0. */
int fers_b200_chol_solve(void* plan, const void* index_dev, const double* factor, const uint8_t* fixed, int nrhs,
                         const double* b, double* x, int accumulate, void* work, size_t work_bytes, void* stream);
/* Iterative refinement helper: r = f - kx on the free DOFs (0 on constrained rows); norms2
 * [2][nrhs] (device) receives ||r_c||^2 and ||f_c||^2 per column (diagnostics; atomics). */
int fers_b200_residual(long long n_dof, int nrhs, const double* f, const double* kx, const uint8_t* fixed, double* r,
                       double* norms2, void* stream);
/* Per-kernel-class device timing of the next factorisations / substitutions (CUDA events on their
 * stream; makes both calls synchronise).  out[10] [host]: ms of front init, K scatter, diagonal
 * tiles, trsm, panel updates, Schur updates, whole factorisation, whole substitution; launches of
 * the factorisation and of the substitution.  on = 1: every kernel class; on = 2: only the Schur-update
 * spans and the two totals (a few events per tree level: cheap enough for a timed region); 0 = off. */
int fers_b200_chol_set_profile(int on);
/* Tile shape of the factorisation GEMMs: bit 0 = panel updates, bit 1 = Schur update; bit set =
 * 128x128 tile with one CTA per SM, clear (default) = 128x64 tile with two CTAs per SM; bit 2 set = no
 * look-ahead (the update of all later columns is then not overlapped with the next block's panel chain). */
/* Tree-parallel use of the sparse direct solve on two GPUs (strong scaling of ONE model; SURVEY.md 8e asks for sharding
 * by combinations only -- this goes further).  chol_plan_partition (before chol_upload): rank r keeps the fronts of the
 * subtree under child r of the root front (rank 0 also the root); all offsets stay those of the whole plan.
 * out[12] = root front, child 0, child 1, root level, [first, one past last) elimination position of subtree 0, of
 * subtree 1, first position and node count of the root, fronts kept, K blocks kept.
 * chol_factor_levels: levels [l_begin, l_end) of the numeric factorisation, asynchronous (level 0 starts a factorisation).
 * chol_solve_phase: phases (bit 0 forward, bit 1 backward, bit 2 un-permute into x) over levels [l_begin, l_end).
 * Between the calls the caller moves, rank to rank: the Schur complement of child 1 (arena of its level parity at its
 * schur_off), its front-local right-hand side rows, the root rows and the subtree rows of the permuted solution
 * (fers_core_b200.chol.TreeParallelSolver does, with torch.distributed send / recv over NVLink). */
int fers_b200_chol_plan_partition(void* plan, int rank, int world, long long* out, int n_out);
int fers_b200_chol_factor_levels(void* plan, const void* index_dev, const double* vals, const int* diag_slot,
                                 const uint8_t* fixed, const double* kspring, double* factor, void* work, size_t work_bytes,
                                 int* info_dev, int l_begin, int l_end, void* stream);
int fers_b200_chol_solve_phase(void* plan, const void* index_dev, const double* factor, const uint8_t* fixed, int nrhs,
                               const double* b, double* x, int accumulate, void* work, size_t work_bytes, int phases,
                               int l_begin, int l_end, void* stream);
int fers_b200_chol_set_variant(int variant);
/* Blocking of the panel updates: `steps` 64-column steps per outer block (default 4 = 256 columns) and `group`
 * outer blocks per group (default 1; measured: 4 x 1 is the fastest of 2 x 1, 4 x 1, 8 x 1, 4 x 2, 4 x 4, 4 x 8 at
 * configs[3]).  A finished block updates the rest of its group, a finished group all later columns. */
int fers_b200_chol_set_outer(int steps, int group);
/* n_out >= 16 additionally returns sums since the last fers_b200_chol_set_profile call: out[10..12] = ms
 * of the Schur updates, of whole factorisations, of whole substitutions; out[13..15] = factorisations,
 * substitutions, Schur-update launches. */
int fers_b200_chol_profile(double* out, int n_out);

/* ---- fp64-equivalent Schur update on the int8 tensor cores (tcgen05.mma kind::i8, Ozaki splitting) -------------
 * C (m2 x m2, lower triangle, column-major, leading dimension ldc) -= A A^T with A m2 x k1 (column-major, lda), both
 * fp64 on the device.  A is cut into `slices` (4..8) signed digit planes per row exponent (6 bits, then 8 bits each); the
 * digit-plane products are exact in int32 (k1 < 16 384) and are summed back in fp64: 7 slices carry 54 bits, 6 carry 46.
 * work: >= fers_b200_oz_syrk_work_bytes(m2, k1, slices) device bytes (digit planes + row scales). */
/* The sparse factorisation sends the Schur updates of tree levels whose fronts all have >= min_k own columns through this
 * kernel (slices = 0 switches it off: fp64 tensor pipe everywhere; default 7 slices, min_k 1024; FERS_CHOL_OZAKI,
 * FERS_CHOL_OZAKI_MIN_K).  The solve checks the residual of every column afterwards and refines while it exceeds 1e-12. */
int fers_b200_chol_set_ozaki(int slices, int min_k);
size_t fers_b200_oz_syrk_work_bytes(int m2, int k1, int slices);
int fers_b200_oz_syrk(const double* A, long long lda, double* C, long long ldc, int m2, int k1, int slices, void* work,
                      size_t work_bytes, void* stream);

/* ---- recovery (pydantic_models.py:1594-1635) --------------------------- */
/* Member end forces f = k_local (T u_e) - p_eq for every (member, rhs):
 *   f_local, f_global, u_local : [n_members][12][nrhs]
 * peq: [n_members*12][nrhs] or NULL.  m_axial_rhs: [n_members][nrhs] axial forces for K_g or NULL. */
int fers_b200_recover_members(int n_members, int nrhs, const double* xyz, const int* m_n1, const int* m_n2,
                              const int* m_prop, const int* m_type, const double* m_roll, const double* m_scale,
                              const int* m_rel, const double* prop, const double* rel,
                              const double* m_axial_rhs, int kg_mode, int kg_suppress,
                              const double* u, const double* peq,
                              double* f_local, double* f_global, double* u_local, void* stream);
/* Warping members: adds the psi coupling to the end torques of f_local / f_global (f_global may be NULL) and
 * writes the bimoments bw[n_w][2][nrhs] (nodal force at the warping DOF of the start / end node, the `bw` of
 * NodeForces, pydantic_models.py:642-649).  w_aux[n_w][2]: auxiliary nodes of the member ends.  Call after
 * fers_b200_recover_members. */
int fers_b200_recover_warping(int n_w, int nrhs, const int* w_member, const int* w_aux, const double* xyz, const int* m_n1,
                              const int* m_n2, const int* m_prop, const double* m_scale, const double* prop, const double* u,
                              double* f_local, double* f_global, double* bw, void* stream);
/* Interior section forces at x_frac 0.25/0.5/0.75 and per-component extrema:
 *   sec [n_members][18][nrhs]; lmax,lmin,gmax,gmin [n_members][6][nrhs].
 * Span loads are given per load case as CSR over members (sp_ptr[n_cases*(n_members+1)] into
 * rows sp_rows[*][9] = kind(0 dist,1 point) + 8 data); factors [nrhs][n_cases].
 * m_rel / rel: the member modifier records of fers_b200_assemble_beams (end offsets: stations are fractions of the
 * elastic axis, transport starts from the forces at its start) or NULL. */
int fers_b200_recover_sections(int n_members, int nrhs, int n_cases, const double* xyz, const int* m_n1,
                               const int* m_n2, const double* m_roll, const int* sp_ptr, const double* sp_rows,
                               const double* factors, const double* f_local, const double* f_global,
                               const int* m_rel, const double* rel,
                               double* sec, double* lmax, double* lmin, double* gmax, double* gmin, void* stream);
/* r[n_dof][nrhs] = supported(node) ? (K u - f) : 0, K WITHOUT support springs. */
int fers_b200_reactions(int n_nodes, int nrhs, const int* rowptr, const int* colidx, const double* vals,
                        const uint8_t* supported, const double* u, const double* f, double* r, void* stream);
/* Plate resultants at the centroid (nx ny nxy mx my mxy qx qy), corner forces and centroid
 * displacements:  p_res [n_plates][8][nrhs], p_nodal [n_plates][18][nrhs], p_cd [n_plates][12][nrhs]. */
int fers_b200_recover_plates(int n_plates, int nrhs, const double* xyz, const int* p_nodes, const double* p_data,
                             const int* p_flags, const double* u, double* p_res, double* p_nodal, double* p_cd,
                             void* stream);
/* ---- unity checks (pydantic_models.py: GenericSpec 2136-2151, UnityCheckDefinition
 *      2202-2220, EntityUnityResult 1472-1489) --------------------------------------------------------
 * Evaluates one compiled Generic check for every (entity, column) and keeps, per entity, the column with
 * the largest utilisation = demand / capacity among the columns with col_mask != 0:
 *   out[n_ent][4] = utilisation, demand, capacity, column (-1: no column applies).
 * Entities are member chains: ent_ptr/ent_members (CSR, chain order), ent_node_ptr/ent_nodes (their nodes,
 * start to end), frames[n_ent][9] = local axes of the first member.  Variables var_desc[n_vars][4] =
 * (kind, a, b, c): 0 per-entity constant ctab[e][a]; 1 member force component a (N Vy Vz Mx My Mz) with
 * aggregation b (0 MaxAbs 1 Max 2 Min 3 Start 4 End 5 station c of 0, 1/4, 1/2, 3/4, 1); 2 postfix
 * program code[a .. a+b); 3 / 4 displacement / deflected position of the start (a = 0) or end node,
 * component b (6 = translation magnitude); 5 chord-relative deflection, component a (local y, local z,
 * magnitude), aggregation b; 6 member force of the entity's REFERENCE member (ent_ref[e][0]); 7 largest |displacement
 * component a| over the entity's nodes; 8 plate stress at the centroid surfaces (a = 0 von Mises, 1 / 2 largest /
 * smallest principal) of plate entities (ent_members = TRI3 rows, p_res / p_data as in fers_b200_recover_plates /
 * fers_b200_assemble_plates; NULL without plates).  Node selector a of kinds 3 / 4: 0 start, 1 end, 2 / 3 start / end of the
 * reference member, 4 reference node (ent_ref[n_ent][4] = reference member, its start node, its end node, reference
 * node; -1 = none -> NaN).  vars_out[n_ent][n_vars] (or NULL) receives the variable values at the governing
 * column (the calculation trace).  code = (op, arg) pairs: 0 push consts[arg], 1 push variable arg, 2 + 3 -
 * 4 * 5 / 6 pow 7 neg 8 abs 9 sqrt 10 min 11 max.  Result arrays in the layout of the recovery calls. */
int fers_b200_unity_generic(int n_ent, int nrhs, int n_vars, int n_ct, const uint8_t* col_mask, const int* ent_ptr,
                            const int* ent_members, const int* ent_node_ptr, const int* ent_nodes,
                            const int* var_desc, const int* code, const double* consts, const double* ctab,
                            int demand_off, int demand_len, int cap_off, int cap_len, const double* xyz,
                            const double* u, const double* f_local, const double* sec, const double* lmax,
                            const double* lmin, const double* frames, const int* ent_ref, const double* p_res,
                            const double* p_data, double* out, double* vars_out, void* stream);
/* axial[n_members][nrhs] = f_local[m][6][rhs] (tension +), used between P-Delta iterations. */
int fers_b200_member_axial(int n_members, int nrhs, const double* f_local, double* axial, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FERS_B200_H */
