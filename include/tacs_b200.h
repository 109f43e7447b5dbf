/*
 * tacs_b200.h -- C ABI of libtacs_b200.so, the B200-native drop-in for the TACS
 * finite-element assembly + Krylov-operator hot path.
 *
 * The reference has no FFI on this path: callers bind to C++ virtual classes in
 * libtacs.so (C++ drivers directly, Python through Cython, tacs/TACS.pyx).  Each entry
 * point below is the flat-C form of the reference member function it replaces (cited as
 * file:line under /root/reference); INTEGRATION.md shows the C++ shim and the Cython
 * stub a maintainer would add on the reference side.  All pointers are host pointers
 * unless a name says `device`; sizes are element counts; arrays are copied, never kept.
 *
 * Conventions (the reference's own, src/TACSAssembler.cpp:513-581, src/TACSObject.h:106-135):
 *   - objects are opaque, reference-counted handles; every *_create returns a handle
 *     holding one reference (NULL on failure); drop it with tacsb200_release
 *   - int-returning calls return 0 on success, non-zero on failure after printing a
 *     message to stderr; nothing throws across this boundary
 *   - there is no CPU execution path: every compute call fails loudly without a GPU
 *   - one process drives one GPU; calls are synchronous with respect to the host unless
 *     documented otherwise (the work is enqueued on the library's stream and the call
 *     returns after the result it hands back is complete)
 */
#ifndef TACS_B200_H
#define TACS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void *tacsb200_handle;

int tacsb200_abi_version(void);
/* Select the GPU (device < 0: cudaSetDevice(LOCAL_RANK or 0)). Called implicitly by the first
   object that needs the device. */
int tacsb200_init(int device);
/* One process per GPU: rank/size of this process and the 128-byte ncclUniqueId created on rank 0
   (tacsb200_comm_unique_id) and distributed by the caller (torch.distributed, MPI, a file...).
   Replaces the MPI_Comm every reference object carries (src/TACSObject.h:39). */
int tacsb200_comm_unique_id(unsigned char id[128]);
int tacsb200_comm_init(int rank, int size, const unsigned char id[128]);
int tacsb200_comm_rank(void);
int tacsb200_comm_size(void);
int tacsb200_synchronize(void);
/* number of tacs_b200 kernels launched since the last call with reset != 0 */
long tacsb200_kernel_launches(int reset);
void tacsb200_release(tacsb200_handle h);

/* ---- constitutive: src/constitutive ------------------------------------------------------ */
/* TACSMaterialProperties(rho, specific_heat, E, nu, ys, alpha, kappa)  TACSMaterialProperties.h:44 */
tacsb200_handle tacsb200_material_properties_create(double rho, double specific_heat, double E, double nu,
                                                    double ys, double alpha, double kappa);
/* orthotropic constructor, TACSMaterialProperties.h:47 */
tacsb200_handle tacsb200_material_properties_create_ortho(double rho, double specific_heat, double E1,
                                                          double E2, double E3, double nu12, double nu13,
                                                          double nu23, double G12, double G13, double G23);
/* TACSOrthotropicPly(plyThickness, properties)  TACSMaterialProperties.h:174 */
tacsb200_handle tacsb200_orthotropic_ply_create(double ply_thickness, tacsb200_handle props);
/* TACSIsoShellConstitutive(props, t, tNum=-1, tlb, tub, tOffset, kcorr)  TACSIsoShellConstitutive.h:33 */
tacsb200_handle tacsb200_iso_shell_constitutive_create(tacsb200_handle props, double t, double tOffset,
                                                       double kcorr);
/* TACSCompositeShellConstitutive(num_plies, ply_props, thickness, angles, kcorr, tOffset)
   TACSCompositeShellConstitutive.h:27 */
tacsb200_handle tacsb200_composite_shell_constitutive_create(int num_plies, tacsb200_handle *plies,
                                                             const double *ply_thickness,
                                                             const double *ply_angles, double kcorr,
                                                             double tOffset);
/* TACSSolidConstitutive(properties, t)  TACSSolidConstitutive.h:34 */
tacsb200_handle tacsb200_solid_constitutive_create(tacsb200_handle props, double t);
/* TACSShellConstitutive::setDrillingRegularization  TACSShellConstitutive.cpp:66 */
/* Constitutive objects given by their constant values: the 22-entry shell tangent stiffness [A B D As drill] and
   mass moments, or the 21-entry solid tangent stiffness and density -- what TACSShellConstitutive::
   evalTangentStiffness / evalMassMoments (src/constitutive/TACSShellConstitutive.h:100-115) and
   TACSSolidConstitutive::evalTangentStiffness / evalDensity (TACSSolidConstitutive.cpp:104-178) return for the classes
   of this path. The reference-side shim (shim/) reads them off the caller's own objects. */
tacsb200_handle tacsb200_shell_constitutive_create_raw(const double *C22, const double *moments3);
tacsb200_handle tacsb200_solid_constitutive_create_raw(const double *C21, double density);
void tacsb200_shell_set_drilling_regularization(double k);
/* TACSConstitutive::evalTangentStiffness  TACSConstitutive.h:340 (22 values shell, 21 solid) */
int tacsb200_constitutive_eval_tangent_stiffness(tacsb200_handle con, double *C);
/* TACSShellConstitutive::evalMassMoments  TACSShellConstitutive.h */
int tacsb200_shell_constitutive_eval_mass_moments(tacsb200_handle con, double *moments);

/* ---- transforms and elements: src/elements ---------------------------------------------------- */
/* TACSShellNaturalTransform / TACSShellRefAxisTransform  shell/TACSShellElementTransform.h:21,95 */
tacsb200_handle tacsb200_shell_natural_transform_create(void);
tacsb200_handle tacsb200_shell_ref_axis_transform_create(const double axis[3]);
/* TACSQuad4Shell / TACSQuad9Shell (transform, constitutive)  shell/TACSShellElementDefs.h:19-25 */
tacsb200_handle tacsb200_quad4_shell_create(tacsb200_handle transform, tacsb200_handle con);
tacsb200_handle tacsb200_quad9_shell_create(tacsb200_handle transform, tacsb200_handle con);
/* TACSLinearHexaBasis / TACSQuadraticHexaBasis  basis/TACSHexaBasis.h */
tacsb200_handle tacsb200_linear_hexa_basis_create(void);
tacsb200_handle tacsb200_quadratic_hexa_basis_create(void);
/* TACSLinearElasticity3D(con, TACS_LINEAR_STRAIN)  TACSLinearElasticity.h:178 */
tacsb200_handle tacsb200_linear_elasticity3d_create(tacsb200_handle con);
/* TACSElement3D(model, basis)  TACSElement3D.h:26 */
tacsb200_handle tacsb200_element3d_create(tacsb200_handle model, tacsb200_handle basis);
/* TACSElement::getNumNodes / getVarsPerNode  TACSElement.h:98-105 */
int tacsb200_element_num_nodes(tacsb200_handle e);
int tacsb200_element_vars_per_node(tacsb200_handle e);
/* TACSElement::addJacobian (TACSElement.h:450) / addResidual (:422) for a batch of `count` elements
   sharing this descriptor, evaluated by the device element kernel. Element-major arrays:
   Xpts[count][3*nn], vars/dvars/ddvars[count][nv] (dvars, ddvars may be NULL),
   res[count][nv], mat[count][nv*nv] row-major; res and mat are overwritten. */
int tacsb200_element_add_jacobian(tacsb200_handle e, int count, double alpha, double beta, double gamma,
                                  const double *Xpts, const double *vars, const double *dvars,
                                  const double *ddvars, double *res, double *mat);
int tacsb200_element_add_residual(tacsb200_handle e, int count, const double *Xpts, const double *vars,
                                  const double *dvars, const double *ddvars, double *res);

/* ---- TACSCreator: src/TACSCreator.h:46-120 --------------------------------------------------- */
/* Every rank passes the same global mesh (the reference reads it on the root rank only and
   scatters it, TACSCreator.cpp:548-747; here the scatter is a local selection). */
tacsb200_handle tacsb200_creator_create(int vars_per_node);
int tacsb200_creator_set_global_connectivity(tacsb200_handle c, int num_nodes, int num_elements,
                                             const int *ptr, const int *conn, const int *elem_id_nums);
int tacsb200_creator_set_boundary_conditions(tacsb200_handle c, int num_bcs, const int *bc_nodes,
                                             const int *bc_ptr, const int *bc_vars, const double *bc_vals);
int tacsb200_creator_set_nodes(tacsb200_handle c, const double *Xpts);
int tacsb200_creator_set_elements(tacsb200_handle c, int num_elems, tacsb200_handle *elems);
/* partitionMesh(split_size, part): part == NULL runs METIS exactly as TACSCreator.cpp:1104-1125 */
/* Adopt the caller's node numbering as final (no first-touch renumbering, one rank): the mesh comes from an existing
   TACSAssembler (getElementConnectivity, src/TACSAssembler.h:90) whose vectors and matrices the caller keeps using. */
int tacsb200_creator_set_keep_numbering(tacsb200_handle creator, int keep_numbering);
int tacsb200_creator_partition_mesh(tacsb200_handle c, int split_size, const int *part);
int tacsb200_creator_get_node_nums(tacsb200_handle c, int *new_nodes);
int tacsb200_creator_get_element_partition(tacsb200_handle c, int *partition);
tacsb200_handle tacsb200_creator_create_tacs(tacsb200_handle c);

/* Host-only planning (no GPU is touched): the integer pipeline of rank `rank` out of `size` for the
   creator's mesh -- local element order, node maps, Aloc/Bext sparsity, gather plans and the neighbour
   exchange lists (src/TACSCreator.cpp:436-909, src/TACSAssembler.cpp:1013-1098, 3384-3440,
   src/bpmat/TACSMatDistribute.cpp:65-415, src/bpmat/TACSBVecDistribute.cpp:299-467). Used by the
   CPU test-suite (including 2-rank gloo runs) and by tools that inspect a decomposition. */
tacsb200_handle tacsb200_creator_create_plan(tacsb200_handle c, int rank, int size);
/* copies the named integer array (out may be NULL to query the length); returns the length or -1.
   names: elem_global elem_ptr elem_conn_global elem_conn_local ext_nodes owner_range scalars
          Aloc_rowp Aloc_cols Bext_rowp Bext_cols ext_col_nodes a_ptr a_src b_ptr b_src r_ptr r_src
          {state,cols,rows,blocks}_{send_peers,send_ptr,send_idx,recv_peers,recv_ptr}
   scalars = [nelems, nowned, nlocal, ext_before, ext_after, np, local_blocks, recv_blocks,
              local_node_slots, recv_node_slots] */
int tacsb200_plan_get_array(tacsb200_handle plan, const char *name, int *out);

/* ---- TACSAssembler: src/TACSAssembler.h:61-523 ------------------------------------------------ */
int tacsb200_assembler_get_vars_per_node(tacsb200_handle a);
int tacsb200_assembler_get_num_nodes(tacsb200_handle a);
int tacsb200_assembler_get_num_owned_nodes(tacsb200_handle a);
int tacsb200_assembler_get_num_elements(tacsb200_handle a);
int tacsb200_assembler_get_owner_range(tacsb200_handle a, int *lo, int *hi);
/* getElementConnectivity (:90): local elements, global node numbers; returns conn length */
int tacsb200_assembler_get_element_connectivity(tacsb200_handle a, int *ptr, int *conn);
/* getGlobalNodeNum (:328) for every local node */
int tacsb200_assembler_get_local_to_global(tacsb200_handle a, int *global);
tacsb200_handle tacsb200_assembler_create_vec(tacsb200_handle a);      /* createVec :179 */
tacsb200_handle tacsb200_assembler_create_node_vec(tacsb200_handle a); /* createNodeVec :163 */
tacsb200_handle tacsb200_assembler_create_mat(tacsb200_handle a);      /* createMat :206 */
int tacsb200_assembler_get_nodes(tacsb200_handle a, tacsb200_handle X);
int tacsb200_assembler_set_nodes(tacsb200_handle a, tacsb200_handle X);
int tacsb200_assembler_set_variables(tacsb200_handle a, tacsb200_handle q, tacsb200_handle qdot,
                                     tacsb200_handle qddot); /* :197 */
int tacsb200_assembler_zero_variables(tacsb200_handle a);
int tacsb200_assembler_apply_bcs_vec(tacsb200_handle a, tacsb200_handle v); /* applyBCs(TACSVec*) :182 */
int tacsb200_assembler_apply_bcs_mat(tacsb200_handle a, tacsb200_handle m); /* applyBCs(TACSMat*) :183 */
int tacsb200_assembler_set_bcs(tacsb200_handle a, tacsb200_handle v);       /* setBCs :187 */
int tacsb200_assembler_set_num_threads(tacsb200_handle a, int t); /* accepted and ignored (:317) */
/* assembleRes(residual, lambda = 1, applyBCs = true)  :221 */
int tacsb200_assembler_assemble_res(tacsb200_handle a, tacsb200_handle res);
/* assembleJacobian(alpha, beta, gamma, residual, A, TACS_MAT_NORMAL, lambda = 1, applyBCs = true) :224 */
int tacsb200_assembler_assemble_jacobian(tacsb200_handle a, double alpha, double beta, double gamma,
                                         tacsb200_handle res, tacsb200_handle mat);

/* Enqueue-only variant (SURVEY 8b "_async"): returns once the kernels are on the compute stream. A following
   tacsb200_vec_get_array of the residual copies it to the host behind the residual kernels only, while the block
   gather of the matrix is still running; pair with tacsb200_synchronize before the matrix is read on the host. */
int tacsb200_assembler_assemble_jacobian_async(tacsb200_handle a, double alpha, double beta, double gamma,
                                               tacsb200_handle res, tacsb200_handle mat);
/* setVariables(q) + assembleJacobian(alpha, beta, gamma, res, A) with the state vector taken from and the residual
   returned to host memory (the owned entries, as getArray / setArray). What a host-resident caller does every Newton /
   time step, in one call so that the transfers overlap the kernels: the state is uploaded in pieces on a copy stream
   and each chunk of elements starts as soon as the piece with its last node has arrived; the residual goes back while
   the matrix is still being gathered. Pinned buffers make the copies asynchronous. Returns when res_host is complete;
   tacsb200_synchronize before the matrix is used from the host. */
int tacsb200_assembler_assemble_jacobian_host(tacsb200_handle a, double alpha, double beta, double gamma,
                                              const double *q_host, double *res_host, tacsb200_handle mat);
/* assembleMatType(matType, A, TACS_MAT_NORMAL, lambda = 1, applyBCs) :227, src/TACSAssembler.cpp:4418-4504.
   matType follows ElementMatrixType (src/elements/TACSElementTypes.h:113-119): 1 = TACS_STIFFNESS_MATRIX,
   2 = TACS_MASS_MATRIX; any other type returns non-zero (not evaluated on the device). */
int tacsb200_assembler_assemble_mat_type(tacsb200_handle a, int mat_type, tacsb200_handle mat, int apply_bcs);
/* addJacobianVecProduct(scale, alpha, beta, gamma, x, y, TACS_MAT_NORMAL, lambda = 1, applyBCs)
   src/TACSAssembler.cpp:5416-5496: y <- y + scale * (alpha K + gamma M) x, matrix free */
int tacsb200_assembler_add_jacobian_vec_product(tacsb200_handle a, double scale, double alpha, double beta,
                                                double gamma, tacsb200_handle x, tacsb200_handle y, int apply_bcs);

/* ---- TACSBVec: src/bpmat/TACSBVec.h:67-163 (TACSVec interface KSM.h:91-115) ---------------------- */
int tacsb200_vec_get_size(tacsb200_handle v);                /* length of getArray(): bs * owned nodes */
int tacsb200_vec_get_array(tacsb200_handle v, double *out);  /* device -> host copy */
int tacsb200_vec_set_array(tacsb200_handle v, const double *in);
double *tacsb200_vec_device_ptr(tacsb200_handle v);          /* owned entries in HBM */
double tacsb200_vec_norm(tacsb200_handle v);
double tacsb200_vec_dot(tacsb200_handle x, tacsb200_handle y);
int tacsb200_vec_mdot(tacsb200_handle x, int n, tacsb200_handle *ys, double *out);
int tacsb200_vec_axpy(tacsb200_handle y, double alpha, tacsb200_handle x);
int tacsb200_vec_axpby(tacsb200_handle y, double alpha, double beta, tacsb200_handle x);
int tacsb200_vec_scale(tacsb200_handle y, double alpha);
int tacsb200_vec_copy_values(tacsb200_handle y, tacsb200_handle x);
int tacsb200_vec_zero_entries(tacsb200_handle y);

/* ---- TACSParallelMat / BCSRMat: src/bpmat/TACSParallelMat.h:52-136, BCSRMat.h:34-206 ---------- */
/* which: 0 = Aloc, 1 = Bext (getBCSRMat :103, BCSRMat::getArrays BCSRMat.h:82) */
int tacsb200_mat_get_sizes(tacsb200_handle m, int which, int *bsize, int *nrows, int *ncols, int *nnzb);
int tacsb200_mat_get_pattern(tacsb200_handle m, int which, int *rowp, int *cols);
int tacsb200_mat_get_values(tacsb200_handle m, int which, double *out);
double *tacsb200_mat_device_values(tacsb200_handle m, int which);
int tacsb200_mat_get_ext_col_nodes(tacsb200_handle m, int *nodes); /* getExtColMap :107 */
/* TACSParallelMat::copyValues / scale / axpy (src/bpmat/TACSParallelMat.cpp:198-246), same non-zero pattern */
int tacsb200_mat_copy_values(tacsb200_handle mat, tacsb200_handle other);
int tacsb200_mat_scale(tacsb200_handle mat, double alpha);
int tacsb200_mat_axpy(tacsb200_handle mat, double alpha, tacsb200_handle other);
/* TACSParallelMat::multTranspose (src/bpmat/TACSParallelMat.cpp:267-290): y = A^T x (one rank) */
int tacsb200_mat_mult_transpose(tacsb200_handle mat, tacsb200_handle x, tacsb200_handle y);
int tacsb200_mat_zero_entries(tacsb200_handle m);
int tacsb200_mat_mult(tacsb200_handle m, tacsb200_handle x, tacsb200_handle y); /* mult :248 */
/* asynchronous variant for benchmarking: enqueue only, pair with tacsb200_synchronize */
int tacsb200_mat_mult_async(tacsb200_handle m, tacsb200_handle x, tacsb200_handle y);
tacsb200_handle tacsb200_mat_create_vec(tacsb200_handle m); /* TACSMat::createVec KSM.h */

/* ---- TACSAuxElements: src/TACSAuxElements.h:52-100, TACSAssembler::setAuxElements src/TACSAssembler.h:158 ------
   State-independent loads bound to an element number of the creator's global numbering and added, scaled by the load
   factor, to that element's residual in assembleRes / assembleJacobian (TACSAssembler.cpp:4207-4223):
   TACSShellTraction(t, useConstTrac) src/elements/shell/TACSShellTraction.h:17-33 (t[3], or t[3 nn] node by node),
   TACSShellPressure(p) src/elements/shell/TACSShellPressure.h:17-29 (p[1] or p[nn]). order = 2 Quad4, 3 Quad9. */
tacsb200_handle tacsb200_aux_elements_create(void);
int tacsb200_aux_elements_add_shell_traction(tacsb200_handle aux, int elem_num, int order, const double *t,
                                             int use_const_trac);
int tacsb200_aux_elements_add_shell_pressure(tacsb200_handle aux, int elem_num, int order, const double *p,
                                             int use_const_pressure);
int tacsb200_assembler_set_aux_elements(tacsb200_handle assembler, tacsb200_handle aux);

/* ---- TACSSchurMat: src/bpmat/TACSSchurMat.h:58-125, TACSSchurMat.cpp:453-552 (addValues), 883-938 (mult) ------
   The four blocks [B E; F C] in the reference's local ordering (interior unknowns b, interface unknowns c), as a view
   of an assembled matrix: b_nodes[nb] / c_nodes[nc] give the owned node of every local index (TACSSchurMat::
   getLocalMap / getSchurMap ->getIndices), the rowp / cols arrays are the patterns of the live reference object
   (getBCSRMat -> getArrays). update() gathers the values of the source matrix (after assembleJacobian) into the four
   blocks on the device; get_values(which = 0 B, 1 E, 2 F, 3 C) copies one block's values to the host so that
   TACSSchurPc::factor runs on them; mult is TACSSchurMat::mult on the device. One rank. */
tacsb200_handle tacsb200_schur_mat_create(tacsb200_handle mat, int nb, const int *b_nodes, int nc, const int *c_nodes,
                                          const int *Browp, const int *Bcols, const int *Erowp, const int *Ecols,
                                          const int *Frowp, const int *Fcols, const int *Crowp, const int *Ccols);
int tacsb200_schur_mat_update(tacsb200_handle schur);
int tacsb200_schur_mat_get_values(tacsb200_handle schur, int which, double *vals);
int tacsb200_schur_mat_mult(tacsb200_handle schur, tacsb200_handle x, tacsb200_handle y);

/* ---- TACSChebyshevSmoother: src/bpmat/TACSParallelMat.h:180-217, TACSParallelMat.cpp:871-1113 ---- */
/* TACSChebyshevSmoother(mat, degree, lower_factor = 1/30, upper_factor = 1.1, iters = 1) */
tacsb200_handle tacsb200_chebyshev_create(tacsb200_handle mat, int degree, double lower_factor,
                                          double upper_factor, int iters);
int tacsb200_chebyshev_factor(tacsb200_handle pc);                                    /* factor() :930 */
int tacsb200_chebyshev_apply_factor(tacsb200_handle pc, tacsb200_handle x, tacsb200_handle y); /* :981 */
double tacsb200_chebyshev_get_spectral_radius(tacsb200_handle pc);  /* Gershgorin bound used by factor() :1024 */

/* ---- GMRES: src/bpmat/KSM.h:392-440, KSM.cpp:547-956 -------------------------------------------- */
tacsb200_handle tacsb200_gmres_create(tacsb200_handle mat, int m, int nrestart);
/* GMRES(mat, pc, m, nrestart, isFlexible) KSM.cpp:547: right preconditioning by a Chebyshev smoother */
tacsb200_handle tacsb200_gmres_create_pc(tacsb200_handle mat, tacsb200_handle pc, int m, int nrestart,
                                         int is_flexible);
int tacsb200_gmres_set_tolerances(tacsb200_handle k, double rtol, double atol);
int tacsb200_gmres_solve(tacsb200_handle k, tacsb200_handle b, tacsb200_handle x, int zero_guess);
/* GMRES::setOrthoType KSM.cpp:747-756 (classical != 0: one mdot sweep + one update sweep per iteration; default
   modified Gram-Schmidt, the reference's default), setMonitor with a KSMPrintStdout(descript, rank, freq)
   KSM.cpp:239-283, setTimeMonitor KSM.cpp:765 */
int tacsb200_gmres_set_ortho_type(tacsb200_handle k, int classical);
int tacsb200_gmres_set_monitor(tacsb200_handle k, const char *descript, int freq);
int tacsb200_gmres_set_time_monitor(tacsb200_handle k);
int tacsb200_gmres_get_iter_count(tacsb200_handle k);
double tacsb200_gmres_get_residual_norm(tacsb200_handle k);

/* ---- device-timed helpers for bench.py (CUDA events on the library's stream) ------------------------ */
/* Run `reps` back-to-back assembleJacobian / assembleRes / mult calls and return the elapsed device
   time in milliseconds (cudaEvent pair on the launching stream); < 0 on failure. */
double tacsb200_time_assemble_jacobian(tacsb200_handle a, double alpha, double beta, double gamma,
                                       tacsb200_handle res, tacsb200_handle mat, int reps);
double tacsb200_time_assemble_res(tacsb200_handle a, tacsb200_handle res, int reps);
double tacsb200_time_mat_mult(tacsb200_handle m, tacsb200_handle x, tacsb200_handle y, int reps);


/* Plan statistics behind bench.py's algorithmic byte counts, valid after the first createMat: out[8] =
   {local staging slots, received staging slots, blocks written directly by the element kernels, upper node-pair
   blocks staged by the element kernels, gathered blocks, gather sources, blocks of [Aloc | Bext], local node pairs}. */
int tacsb200_assembler_get_plan_stats(tacsb200_handle assembler, long *out);

/* Per-kernel device timing: when enabled every launch is bracketed by CUDA events on the launching
   stream. collect() synchronises, sums the log into ms[8] / count[8] indexed by
   {0 element, 1 residual gather, 2 block gather, 3 boundary conditions, 4 SpMV, 5 vector, 6 dot, 7 halo}
   and clears it. */
int tacsb200_profile_enable(int on);
int tacsb200_profile_collect(double *ms, long *count);
/* The same log by kernel as launched: lines "kernel name|launches|ms" of the last collect() (library-owned
   string, valid until the next collect). */
const char *tacsb200_profile_named(void);
/* Roofline denominators measured on this device: FP64 FMA throughput (TFLOP/s, 2 flop per FMA) of a
   register-resident DFMA stream, and device-to-device copy bandwidth (GB/s, read + write bytes). */
double tacsb200_measure_fp64_tflops(void);
double tacsb200_measure_copy_gbs(void);

#ifdef __cplusplus
}
#endif
#endif /* TACS_B200_H */
