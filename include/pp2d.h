/*
 * pp2d.h -- C ABI of the B200-native hot path of path_planning_2d.
 *
 * This is the drop-in boundary.  The reference has no
 * FFI: its host classes share C++-linkage symbols, global device pointers and
 * raw kernel launches with the .cu files.  Each entry point below names the
 * reference call sites it replaces (paths relative to
 * /root/reference/path_planning_2d/); INTEGRATION.md shows the mechanical
 * edit of the reference classes that binds them.
 *
 * Conventions kept from the reference:
 *   - the caller owns every host buffer; inputs are copied, never retained;
 *   - every call is synchronous on return (the reference synchronises after
 *     every launch: src/mdp/path_planning_2d.cu:231,236);
 *   - row-major cells, idx = y*width + x; map value 1 = occupied, 0 = free
 *     (src/mdp/path_planning_2d.cu:191-205);
 *   - 9 actions, action u moves by (u%3-1, u/3-1)
 *     (src/mdp/path_planning_2d_cuda.cu:83-88).
 * Errors: the reference prints and exit()s inside checkCudaErrors
 * (include/path_planning_2d/helper_cuda/helper_cuda.h:984-999).  A library
 * must not exit: every function returns PP2D_OK (0) or a negative code and
 * pp2d_last_error() returns the message; the reference-side shim keeps the
 * print-and-exit behaviour.
 *
 * There is no CPU fallback: every entry point needs a CUDA device of compute
 * capability 10.x and fails with PP2D_ERR_CUDA otherwise.
 */
#ifndef PP2D_H_
#define PP2D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP2D_OK 0
#define PP2D_ERR_INVALID (-1) /* bad argument (null, zero size, goal outside) */
#define PP2D_ERR_GOAL_OCCUPIED (-2) /* src/mdp/path_planning_2d.cu:84-88 */
#define PP2D_ERR_CUDA (-3)    /* CUDA runtime error, see pp2d_last_error() */
#define PP2D_ERR_STATE (-4)   /* call order violated */

/* Message of the last failure on the calling thread ("" if none). */
const char* pp2d_last_error(void);
/* ABI version of this library (bumped on any signature change). */
int pp2d_abi_version(void);
/* Number of CUDA kernels this library has launched in this process so far
 * (all handles).  bench.py reports the delta over its timed region. */
uint64_t pp2d_kernel_launches(void);

/* ------------------------------------------------------------------------
 * MDP value iteration
 * ------------------------------------------------------------------------ */
typedef struct pp2d_mdp pp2d_mdp;

/*
 * Replaces allocateDeviceMemory + map upload + cudaGenerateModelData
 * (src/mdp/path_planning_2d_cuda.cu:40-64,174-213;
 *  src/mdp/path_planning_2d.cu:84-106).
 * map: height*width bytes, 1 = occupied.  (goal_x, goal_y) must be a free
 * cell, otherwise PP2D_ERR_GOAL_OCCUPIED (the reference's initialize() fails
 * the same way).  J is initialised to 0 and the action grid to 0 exactly as
 * the reference's cudaMemset does.  The per-cell 360-byte tables of the
 * reference are never materialised: the model is a pure function of the 3x3
 * occupancy and is stored as a 2-byte code per cell.
 */
int pp2d_mdp_create(uint32_t height, uint32_t width, const uint8_t* map,
                    uint32_t goal_x, uint32_t goal_y, float gamma,
                    pp2d_mdp** out);

/*
 * Row-sharded variant for one-process-per-GPU runs.
 * The handle owns rows [row_begin, row_end) of a height x width grid; `map`
 * is still the whole grid (only rows row_begin-3 .. row_end+2 are read).
 * Ghost rows of J are exchanged by the caller between pp2d_mdp_sweeps calls
 * through the pointers of pp2d_mdp_halo (NCCL send/recv in
 * path_planning_2d_b200/distributed.py).
 */
int pp2d_mdp_create_shard(uint32_t height, uint32_t width, const uint8_t* map,
                          uint32_t goal_x, uint32_t goal_y, float gamma,
                          uint32_t row_begin, uint32_t row_end,
                          pp2d_mdp** out);

/*
 * Start over on the same handle with a new map (same height/width/rows) and
 * goal: J = 0, action = 0, sweep count 0, codes rebuilt.  Same checks as
 * pp2d_mdp_create; lets a long-running planner re-solve without paying the
 * device allocations again (the reference allocates once per process,
 * src/mdp/path_planning_2d.cu:91).
 */
int pp2d_mdp_reset(pp2d_mdp* h, const uint8_t* map, uint32_t goal_x,
                   uint32_t goal_y);

/* Replaces freeDeviceMemory (src/mdp/path_planning_2d_cuda.cu:66-74). */
void pp2d_mdp_destroy(pp2d_mdp* h);

/* Run the library's kernels on `stream` (a cudaStream_t; NULL = default
 * stream) from now on.  Calls still synchronise that stream before
 * returning unless pp2d_mdp_set_async(h, 1) was called. */
int pp2d_mdp_set_stream(pp2d_mdp* h, void* stream);
/* async != 0: pp2d_mdp_sweeps / pp2d_mdp_residual_begin only enqueue work
 * (used by the multi-GPU driver to overlap with NCCL on the same stream). */
int pp2d_mdp_set_async(pp2d_mdp* h, int async);

/*
 * n Jacobi Bellman backups of every owned cell; replaces n launches of
 * cudaOneStepValueIteration (src/mdp/path_planning_2d_cuda.cu:215-264,
 * call sites src/mdp/path_planning_2d.cu:226-237).  After the call J and the
 * greedy action grid are exactly those the reference holds after the same n
 * launches (the action grid is produced by the last backup only, which is
 * all that is observable).  For a shard, n must not exceed the halo depth
 * the caller refreshed (pp2d_mdp_halo: depth 2), i.e. n <= 2 between
 * exchanges; for an unsharded handle any n.
 */
int pp2d_mdp_sweeps(pp2d_mdp* h, uint32_t n);
/* Same, but with want_action == 0 the greedy action grid is left untouched
 * (all n backups are value-only; the multi-GPU driver uses this between
 * ghost-row exchanges and asks for the action only on the last sweep). */
int pp2d_mdp_sweeps_ex(pp2d_mdp* h, uint32_t n, int want_action);

/*
 * Inf-norm of the change of J since the previous call (or since creation),
 * max_s |J_now(s) - J_then(s)| over the owned cells, as the reference
 * computes it on the host every 100 sweeps
 * (src/mdp/path_planning_2d.cu:243-251).
 */
int pp2d_mdp_residual(pp2d_mdp* h, float* inf_norm);

/*
 * The reference's valueIteration() loop (src/mdp/path_planning_2d.cu:207-269):
 * batches of 100 sweeps until the residual of a batch is
 * <= 5.0/(1.0-gamma)*1e-3 (double arithmetic on the float gamma).
 * sweeps_out receives the total number of sweeps; residuals (optional,
 * capacity max_residuals) one inf-norm per batch.  Unsharded handles only.
 */
int pp2d_mdp_solve(pp2d_mdp* h, uint32_t* sweeps_out, double* residuals,
                   uint32_t max_residuals);

/*
 * Replaces the two result cudaMemcpy calls
 * (src/mdp/path_planning_2d.cu:118-126).  cost: rows*width floats,
 * action: rows*width bytes (rows = owned rows); either may be NULL.
 */
int pp2d_mdp_download(pp2d_mdp* h, float* cost, uint8_t* action);

/*
 * MdpPathPlanning2d::beliefCallback (src/mdp/path_planning_2d.cu:168-189):
 * the action at the first strict maximum of `belief` (height*width floats,
 * host memory).  Unsharded handles only.
 */
int pp2d_mdp_plan(pp2d_mdp* h, const float* belief, uint8_t* action);
/* Batched form: n_beliefs beliefs laid out [n][height*width]. */
int pp2d_mdp_plan_batch(pp2d_mdp* h, const float* beliefs, uint32_t n_beliefs,
                        uint8_t* actions);

/*
 * Row W of SURVEY.md section 8a: greedy rollout of the action grid from
 * (start_x, start_y): follow (u%3-1, u/3-1) until action 4 (stay), the map
 * border, or max_len cells.  Writes cell indices y*width+x (start included)
 * and their count.  Unsharded handles only.
 */
int pp2d_mdp_waypoints(pp2d_mdp* h, uint32_t start_x, uint32_t start_y,
                       uint32_t* cells, uint32_t max_len, uint32_t* n_out);

/* Number of sweeps applied so far. */
uint32_t pp2d_mdp_sweep_count(const pp2d_mdp* h);

/*
 * Ghost-row exchange for shards.  depth = 2 rows.  After any
 * pp2d_mdp_sweeps call the caller copies `bytes` bytes
 *   from this rank's send_top  to the upper neighbour's recv_bottom,
 *   from this rank's send_bottom to the lower neighbour's recv_top
 * (device pointers into the CURRENT J buffer; they change after every
 * pp2d_mdp_sweeps call, so query them again).  Pointers for a missing
 * neighbour (first/last shard) are still valid memory and may be ignored.
 */
typedef struct pp2d_halo {
  void* send_top;
  void* send_bottom;
  void* recv_top;
  void* recv_bottom;
  size_t bytes;
} pp2d_halo;
int pp2d_mdp_halo(pp2d_mdp* h, pp2d_halo* out);

/* Device-side residual for shards: enqueue the reduction, then read it. */
int pp2d_mdp_residual_device(pp2d_mdp* h, void** dev_float_out);

#ifdef __cplusplus
}
#endif
#endif /* PP2D_H_ */
