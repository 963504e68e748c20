/*
 * femb200 -- B200-native (sm_100a) finite-element assembly, C ABI.
 *
 * Drop-in boundary for FEMpy's assembly hot path.  The reference has no FFI layer; the seam it
 * offers is three private methods of FEMpyProblem plus the batched Element methods and the
 * AssemblyUtils helpers.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; femb_last_error() gives the text
 *   - float64 C-contiguous arrays; int64 connectivity / DOF / CSR indices at the boundary
 *   - host-pointer calls are synchronous; *_dev calls are asynchronous on the plan's stream
 *   - the caller owns every buffer passed in; the plan owns its device memory until destroy
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 */
#ifndef FEMB200_H
#define FEMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct femb_plan femb_plan;

/* constitutive model ids (FEMpy/Constitutive/IsoPlaneStress.py, IsoPlaneStrain.py) */
#define FEMB_PLANE_STRESS 0
#define FEMB_PLANE_STRAIN 1
#define FEMB_ISO_3D 2 /* FEMpy/Constitutive/Iso3D.py: 3-D plans (numDim = numStates = 3), no design variable */
#define FEMB_ISO_1D 3 /* FEMpy/Constitutive/Iso1D.py: 1-D plans, the per-element scaling array is the cross-sectional area */

/* element functions (ConstitutiveModel.getFunction names: "Mass", "Von-Mises-Stress") */
#define FEMB_FUNC_MASS 0
#define FEMB_FUNC_VON_MISES 1 /* 2-D and 3-D models (StressModels/VonMises.py:26-89) */
#define FEMB_FUNC_STRESS 2    /* Iso1D "Stress" (Iso1D.py:217-225) */
#define FEMB_FUNC_STRAIN 3    /* Iso1D "Strain" (Iso1D.py:210-215) */

/* element reductions (FEMpy/Elements/Element.py:221-259) */
#define FEMB_REDUCE_NONE 0 /* (E, P) point values */
#define FEMB_REDUCE_SUM 1
#define FEMB_REDUCE_MEAN 2
#define FEMB_REDUCE_INTEGRATE 3
#define FEMB_REDUCE_MIN 4
#define FEMB_REDUCE_MAX 5
#define FEMB_REDUCE_KSMAX 6 /* KS aggregate, rho = 100 (FEMpy/Utils/KSAgg.py:25-58): max + log(mean(exp(rho (v - max)))) / rho.
                               The reference reaches this for "ksmax" AND "ksmin" (Element.py:261-265) */

/* what femb_assemble produces */
#define FEMB_WANT_K 1
#define FEMB_WANT_R 2

/* material + kinematics of one assembly call.
 * E, nu  -> D matrix exactly as FEMpy/Constitutive/StressModels/IsoStress.py:25-78
 * model  -> FEMB_PLANE_STRESS | FEMB_PLANE_STRAIN
 * nonlinear != 0 -> Green strain (ContinuumStrains.py:112-126,158-172), else small strain
 * rho    -> density returned by the "Mass" function (IsoPlaneStress.py:211-216) */
typedef struct femb_material {
    double E;
    double nu;
    double rho;
    int32_t model;
    int32_t nonlinear;
} femb_material;

const char* femb_last_error(void);
/* page-locked host memory for result buffers (the Python layer keeps a pool of them: fempy_b200/_pinned.py) */
int femb_host_alloc(size_t bytes, void** out);
int femb_host_free(void* ptr);
int femb_device_count(void);
const char* femb_version(void);

/* ---- plan: mesh topology + element tables -> CSR pattern + scatter map (built on the GPU, once)
 * replaces FEMpyModel.__init__'s per-type element registry and getDOFfromNodeInds
 * (FEMpy/Model.py:123-142,387-414) and the COO->CSR structure scipy derives on every call
 * (FEMpy/Problem.py:613-615). */
int femb_plan_create(int64_t numNodes, int32_t numDim, int32_t numStates, int32_t device, femb_plan** out);

/* one element type ("block"); element ids run across blocks in the order they are added
 * (FEMpy/Model.py:137-142).  N: (P, n) shape functions, dNdxi: (P, numDim, n) parametric gradients,
 * w: (P) weights at the element's own Gauss points (Element.computeShapeFunctions /
 * computeShapeFunctionGradients / getIntegrationPointWeights).  conn: (numElems, n) node ids.
 * Supported: 2-D plans -- 3/6/10-node triangles and 4/9/16-node quads with the families' default rules (1/3/3 and 4/9/16
 * points) or the other rules `intOrder=` can select (triangles 1/3/4 points, quads n x n with n = 1..4); 1-D / 3-D plans --
 * any Lagrange family of up to 27 nodes.  Which kernel a 2-D block gets is decided from its TABLES, checked here on the
 * host: the row-owner / fan kernels evaluate a 3- or 4-node element renumbered so that the row's node comes first, which is
 * exact only for tables invariant under cyclic renumbering of the nodes (counter-clockwise order + a rotation-symmetric
 * rule, as QuadElement2D / TriElement2D order 1); the closed-form fan kernel additionally needs the standard bilinear /
 * linear tables.  Blocks whose tables fail a check (e.g. a 4-node block in tensor order BL, BR, TL, TR) run on the scatter
 * kernels -- same result, FP64 atomics (femb_plan_get_info "rows_path" / "fan_path" / "closed_form_q4" tell). */
int femb_plan_add_block(femb_plan* plan, int32_t nodesPerElem, int32_t numQuad, const double* N, const double* dNdxi,
                        const double* w, int64_t numElems, const int64_t* conn);
/* the same with the connectivity already in device memory (device-side mesh ingest): the element registry,
 * DOF map (FEMpy/Model.py:123-142,387-414), pattern and kernel records are then built without the mesh touching the host */
int femb_plan_add_block_dev(femb_plan* plan, int32_t nodesPerElem, int32_t numQuad, const double* N, const double* dNdxi,
                        const double* w, int64_t numElems, const int64_t* d_conn);

/* builds the node->element adjacency, the sorted structural CSR pattern and the corner records of the row-owner
 * kernel (the element scatter map of the fallback kernels is built on their first use) */
int femb_plan_finalize(femb_plan* plan);
void femb_plan_destroy(femb_plan* plan);

/* options:
 *   "assembly_path": 2 (default) row-owner kernel -- one thread per node row, one launch per element block, no
 *       atomics, bit-reproducible; 3/4/6/9/10-node elements, node degree <= 48, else the scatter kernels run;
 *       0: slot-map scatter kernels with FP64 atomics (every family);
 *   "fan_rows": 1 (default) single-block plans of 3- / 4-node elements with the reference's order-1 tables run the
 *       write-once fan kernel for the small-strain model (corners of a node chained into fans, every row block finished
 *       in registers and stored once); 0: the accumulating row-owner kernel.  "fan_compact" (before finalize): 0 keeps the
 *       wide fan records even when node / element ids fit the compact forms (4 nodes: 24-bit differences to the owner;
 *       3 nodes: 8-byte chained records) (tests); "fan_prefetch": L2 look-ahead distance of the fan kernel in 32-node
 *       warps (-1 default: a sixth of the resident wave, 0 off; a hint only, results do not depend on it);
 *       "fan_carveout": shared-memory carve-out of the fan kernel in percent (-1 default: 85 % of what the register-limited
 *       warps would need; performance only);
 *   "points_only" (before the first block): blocks of any node / point count, for femb_point_quantities only;
 *   "balance_rows" (before finalize): row schedule of the row-owner kernel -- 0 threads take nodes in node order,
 *       1 (default) nodes of every 256-node window are ordered by their number of adjacent elements when that saves
 *       more than 8 % of the warps' element iterations (9-node quads, mixed meshes), 2 always; results do not depend on it */
int femb_plan_set_option(femb_plan* plan, const char* name, int64_t value);
/* "assembly_path", "rows_path" (1: the row-owner kernel will run), "max_degree", "rows_balanced", "row_iters_node_order",
 * "row_iters_balanced" (sum over warps of the longest element walk in either schedule), "closed_form_q4" (1: every
 * 4-node block carries the standard bilinear tables and runs the closed-form kernel), "fan_path" (1: the write-once fan
 * kernel will run for the small-strain model), "fan_compact"; -1 for an unknown name */
int64_t femb_plan_get_info(const femb_plan* plan, const char* name);

/* external != 0: run the plan's kernels on the caller's CUDA stream (e.g.
 * torch.cuda.current_stream().cuda_stream; NULL is the legacy default stream); external == 0: back to the
 * plan's own stream */
int femb_plan_set_stream(femb_plan* plan, void* cudaStream, int32_t external);
int femb_plan_sync(femb_plan* plan);

int64_t femb_plan_num_dof(const femb_plan* plan);
int64_t femb_plan_num_elems(const femb_plan* plan);
int64_t femb_plan_nnz(const femb_plan* plan);
double femb_plan_build_ms(const femb_plan* plan); /* device time spent in finalize */
int64_t femb_plan_last_launches(const femb_plan* plan); /* kernels launched by the last compute call */

/* CSR structure of the stiffness matrix: indptr (numDOF+1), indices (nnz), int64, rows sorted --
 * what scipy's csr_array(coo_array(...)) yields in FEMpyProblem._assembleMatrix (Problem.py:613-615) */
int femb_plan_pattern(femb_plan* plan, int64_t* indptr, int64_t* indices);
int femb_plan_pattern_dev(femb_plan* plan, int64_t* d_indptr, int64_t* d_indices);

/* ---- device-side model.  Node coordinates and per-element thickness are constants of a Newton or
 * optimisation loop (the reference changes them through FEMpyModel.setCoordinates / setDesignVariables, Model.py:240-276
 * and gathers them again per element on every assembly, Model.py:288-329).  After this call femb_assemble,
 * femb_function and femb_assemble_solve accept NULL for coords / thickness and use the resident copies: only states
 * (and loads) cross PCIe per call.  Passing NULL here, or an explicit pointer to one of those entries later, ends the
 * residency of that array.  The arrays are copied before the call returns. */
int femb_plan_set_geometry(femb_plan* plan, const double* coords /* N*2 or NULL */, const double* thickness /* E or NULL */);

/* ---- K + R assembly: FEMpyProblem._assembleMatrix + _assembleResidual (FEMpy/Problem.py:565-661)
 * coords (numNodes, numDim), states (numNodes, numStates), thickness (numElems) ["Thickness" DV],
 * bcDof/bcVal (nBC) = AssemblyUtils.convertBCDictToLists output (duplicates allowed, reference
 * semantics: rows zeroed, +1.0 on the diagonal per occurrence, last value wins in the residual;
 * AssemblyUtils.py:26-94), applied when applyBCs != 0;  loads (numDOF) or NULL =
 * convertLoadsDictToVector output.  Outputs: Kdata (nnz) in femb_plan_pattern order, R (numDOF).
 * `want` selects FEMB_WANT_K / FEMB_WANT_R / both; unused outputs may be NULL. */
int femb_assemble(femb_plan* plan, const double* coords, const double* states, const double* thickness,
                  const femb_material* mat, const int64_t* bcDof, const double* bcVal, int64_t nBC, int32_t applyBCs,
                  const double* loads, int32_t want, double* Kdata, double* R);

/* device-resident variant: BCs are staged once with femb_plan_set_bcs; all pointers are device pointers */
int femb_plan_set_bcs(femb_plan* plan, const int64_t* bcDof, const double* bcVal, int64_t nBC);
int femb_assemble_dev(femb_plan* plan, const double* d_coords, const double* d_states, const double* d_thickness,
                      const femb_material* mat, int32_t applyBCs, const double* d_loads, int32_t want, double* d_Kdata,
                      double* d_R);

/* CUDA-event timing of the dominant assembly kernel(s) inside femb_assemble_dev, on the plan's stream
 * (measurement only: bench.py's roofline leg).  read() synchronises, returns the summed kernel time and
 * the number of timed calls since the last read, and resets the counters. */
int femb_plan_profile(femb_plan* plan, int32_t enable);
int femb_plan_profile_read(femb_plan* plan, double* sumMs, int64_t* count);

/* measurement builds only (-DROWS_TRACE=1): warp lifetime statistics of the row-owner kernel.  enable != 0 /
 * out == NULL switches the counters on (0 off); out != NULL copies up to `capacity` int64 values back. */
int femb_plan_row_stats(femb_plan* plan, int32_t enable, int64_t* out, int64_t capacity);

/* ---- element functions: FEMpyProblem.computeFunction / Element.computeFunction
 * (FEMpy/Problem.py:282-355, FEMpy/Elements/Element.py:163-273).  out: (numElems) or (sum_b E_b*P_b)
 * for FEMB_REDUCE_NONE, blocks concatenated in plan order. */
int femb_function(femb_plan* plan, const double* coords, const double* states, const femb_material* mat,
                  int32_t func, int32_t reduction, double* out);
int femb_function_dev(femb_plan* plan, const double* d_coords, const double* d_states, const femb_material* mat,
                      int32_t func, int32_t reduction, double* d_out);

/* ---- point quantities: Element._computeFunctionEvaluationQuantities (FEMpy/Elements/Element.py:275-368) with its helpers
 * _interpolationProduct (:918-939), _computeNPrimeCoordProduct (:942-978), _computeUPrimeProduct (:981-1018),
 * _computeDUPrimeDqProduct (:1021-1048) and det2 / inv2 (FEMpy/LinAlg/LinAlg.py:70-158), evaluated at the points the
 * plan's blocks were added with (their N / dNdxi tables: the Gauss points of any rule, or any other parametric points).
 * Outputs, (numPoints, ...) with the blocks concatenated in plan order, any may be NULL: coord (.,2), state (.,2),
 * stateGrad (.,2,2) [state][dim], jac (.,2,2), jacDet (.), stateGradSens (.,2,nodesPerElem) per block. */
int femb_point_quantities(femb_plan* plan, const double* coords, const double* states, double* coord, double* state,
                          double* stateGrad, double* jac, double* jacDet, double* stateGradSens);

/* ---- device-resident linear solve: the hand-off after assembly.
 * Replaces, on request, the host factorisation of FEMpyProblem.solveJacLinear (FEMpy/Problem.py:171-198).  K: the
 * plan's CSR values as femb_assemble[_dev] left them with applyBCs = 1 and the plan's current BC set (constrained rows =
 * zeros + occurrence count on the diagonal, AssemblyUtils.py:26-62); b, x: numDOF.  Solves K x = b by Jacobi-
 * preconditioned conjugate gradients on the free DOFs (x_c = b_c / count substituted; symmetric positive definite for
 * the elasticity models) until ||r|| <= relTol ||r_0|| or maxIter; iters / relRes may be NULL.  Three launches per
 * iteration on the plan's stream, host check every 25 iterations.  Not a bit-for-bit replacement of a direct solve:
 * the answer agrees with it to about relTol x condition number. */
int femb_solve_pcg(femb_plan* plan, const double* K, const double* b, double* x, double relTol, int32_t maxIter,
                   int32_t* iters, double* relRes);
int femb_solve_pcg_dev(femb_plan* plan, const double* d_K, const double* d_b, double* d_x, double relTol, int32_t maxIter,
                       int32_t* iters, double* relRes);
/* One step of FEMpyProblem.solve (Problem.py:137-169) without the CSR values crossing PCIe: H2D of the nodal inputs,
 * one fused K + R assembly with the BCs applied, update = K^-1 (-R) by femb_solve_pcg_dev on the device values, D2H of
 * update (numDOF) and, when R != NULL, of the residual.  The caller adds update to its states. */
int femb_assemble_solve(femb_plan* plan, const double* coords, const double* states, const double* thickness,
                        const femb_material* mat, const int64_t* bcDof, const double* bcVal, int64_t nBC,
                        const double* loads /* may be NULL */, double relTol, int32_t maxIter, double* update,
                        double* R /* may be NULL */, int32_t* iters, double* relRes);

/* ---- per-element blocks: Element.computeResidualJacobians / computeResiduals
 * (FEMpy/Elements/Element.py:370-486).  Ke: (E_b, n*s, n*s) per block, Re: (E_b, n, s) per block,
 * blocks concatenated.  Either output may be NULL. */
int femb_element_blocks(femb_plan* plan, const double* coords, const double* states, const double* thickness,
                        const femb_material* mat, double* Ke, double* Re);

/* ---- COO expansion of element matrices: AssemblyUtils.localMatricesToCOOArrays
 * (FEMpy/Utils/AssemblyUtils.py:168-214): element-major, row-major triplets with exactly-zero values
 * dropped.  rows/cols/vals have capacity numElems*nd*nd; *numOut receives the kept count. */
int femb_local_matrices_to_coo(const double* localMats, const int64_t* localDOF, int64_t numElems, int32_t nd,
                               int32_t device, int64_t* rows, int64_t* cols, double* vals, int64_t* numOut);

/* ---- residual scatter: AssemblyUtils.scatterLocalResiduals (FEMpy/Utils/AssemblyUtils.py:217-238),
 * globalResidual (numNodes*numStates) updated in place */
int femb_scatter_local_residuals(const double* localRes, const int64_t* conn, int64_t numElems, int32_t nodesPerElem,
                                 int32_t numStates, int64_t numNodes, int32_t device, double* globalResidual);

#ifdef __cplusplus
}
#endif
#endif /* FEMB200_H */
