/*
 * oracle/ref_mdp_driver.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * ROS/OpenCV-free driver around the UNMODIFIED reference MDP kernels.  The
 * reference translation unit is compiled where it lies
 * (/root/reference/path_planning_2d/src/mdp/path_planning_2d_cuda.cu, pulled
 * in by the #include below through -I; no reference source is copied into
 * this repository) with the reference's own flag --use_fast_math, for
 * sm_100a.  Output: oracle/_ref/libpp2d_ref_mdp.so (git-ignored, travels to
 * the GPU box).  It needs a GPU: it is used by `-m gpu` tests to pin
 * oracle/mdp_oracle.c, to generate tests/golden/, and by bench.py to time
 * "the reference's own CUDA kernels on the same B200".
 *
 * What is restated here is only host glue, mirroring
 *   src/mdp/path_planning_2d.cu:90-126  (initialize: alloc, upload, model,
 *                                        download)
 *   src/mdp/path_planning_2d.cu:207-269 (valueIteration: 50x ping-pong pairs,
 *                                        D2H, inf-norm, stopping rule)
 * with OpenCV's absdiff/minMaxIdx replaced by a float loop and imshow dropped.
 */
#include <path_planning_2d_cuda.cu>   /* the reference TU, via -I.../src/mdp */

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {
dim3 ref_grid(uint32_t h, uint32_t w) {
  /* mdp_host:97-100 */
  return dim3(static_cast<int>(std::ceil(static_cast<float>(w) / 8.0)),
              static_cast<int>(std::ceil(static_cast<float>(h) / 8.0)));
}
void ref_setup(uint32_t h, uint32_t w, const uint8_t* map, uint32_t gx,
               uint32_t gy) {
  allocateDeviceMemory(h, w);                                  /* mdp_host:91 */
  checkCudaErrors(cudaMemcpy(dev_map, map, sizeof(uint8_t) * h * w,
                             cudaMemcpyHostToDevice));         /* mdp_host:94 */
  cudaGenerateModelData<<<ref_grid(h, w), dim3(8, 8)>>>(
      h, w, gx, gy, dev_map, dev_trans_prob, dev_stage_cost);  /* :103 */
  checkCudaErrors(cudaDeviceSynchronize());
}
}  // namespace

extern "C" {

/* Dump the model tables the reference builds (mdp_cuda:174-213). */
int ref_mdp_tables(uint32_t h, uint32_t w, const uint8_t* map, uint32_t gx,
                   uint32_t gy, float* trans_prob, float* stage_cost) {
  ref_setup(h, w, map, gx, gy);
  checkCudaErrors(cudaMemcpy(trans_prob, dev_trans_prob,
                             sizeof(float) * h * w * 81,
                             cudaMemcpyDeviceToHost));
  checkCudaErrors(cudaMemcpy(stage_cost, dev_stage_cost,
                             sizeof(float) * h * w * 9,
                             cudaMemcpyDeviceToHost));
  freeDeviceMemory();
  return 0;
}

/* Full solve with the reference's loop.  max_batches <= 0: until converged.
 * Returns the number of sweeps. */
int ref_mdp_solve(uint32_t h, uint32_t w, const uint8_t* map, uint32_t gx,
                  uint32_t gy, float gamma, float* J_out, uint8_t* action_out,
                  double* residuals, int max_batches) {
  ref_setup(h, w, map, gx, gy);
  const size_t n = static_cast<size_t>(h) * w;
  float* prev = static_cast<float*>(calloc(n, sizeof(float)));
  float* curr = static_cast<float*>(calloc(n, sizeof(float)));
  dim3 grid = ref_grid(h, w), block(8, 8);
  int total = 0, batch = 0;
  double inf_norm = 0.0;
  double max_optimal_cost = 5.0 / (1.0 - gamma);               /* :221 */
  do {
    for (int i = 0; i < 50; ++i) {                             /* :226-237 */
      cudaOneStepValueIteration<<<grid, block>>>(
          h, w, gamma, dev_trans_prob, dev_stage_cost, dev_optimal_cost1,
          dev_optimal_cost2, dev_optimal_action);
      checkCudaErrors(cudaDeviceSynchronize());
      cudaOneStepValueIteration<<<grid, block>>>(
          h, w, gamma, dev_trans_prob, dev_stage_cost, dev_optimal_cost2,
          dev_optimal_cost1, dev_optimal_action);
      checkCudaErrors(cudaDeviceSynchronize());
    }
    total += 100;
    checkCudaErrors(cudaMemcpy(curr, dev_optimal_cost1, sizeof(float) * n,
                               cudaMemcpyDeviceToHost));       /* :243 */
    float m = 0.0f;                                            /* :247-249 */
    for (size_t i = 0; i < n; ++i) {
      float d = std::fabs(prev[i] - curr[i]);
      if (d > m) m = d;
    }
    inf_norm = m;
    memcpy(prev, curr, sizeof(float) * n);
    if (residuals && (max_batches <= 0 || batch < max_batches))
      residuals[batch] = inf_norm;
    ++batch;
    if (max_batches > 0 && batch >= max_batches) break;
  } while (inf_norm > max_optimal_cost * 1e-3);                /* :263 */
  checkCudaErrors(cudaMemcpy(J_out, dev_optimal_cost1, sizeof(float) * n,
                             cudaMemcpyDeviceToHost));         /* :123 */
  checkCudaErrors(cudaMemcpy(action_out, dev_optimal_action,
                             sizeof(uint8_t) * n, cudaMemcpyDeviceToHost));
  free(prev);
  free(curr);
  freeDeviceMemory();
  return total;
}

/* policyIteration (path_planning_2d.cu:271-357, dead code in the reference:
 * its call is commented out at :115-116) with the reference kernels
 * cudaOneStepPolicyEvaluation / cudaPolicyImprovment; OpenCV's absdiff /
 * minMaxIdx / compare replaced by float loops.  Returns the number of
 * evaluation sweeps. */
int ref_mdp_policy_iteration(uint32_t h, uint32_t w, const uint8_t* map, uint32_t gx,
                             uint32_t gy, float gamma, float* J_out, uint8_t* action_out,
                             double* residuals, uint32_t* changed, int max_rounds) {
  ref_setup(h, w, map, gx, gy);
  const size_t n = static_cast<size_t>(h) * w;
  float* prev = static_cast<float*>(calloc(n, sizeof(float)));
  float* curr = static_cast<float*>(calloc(n, sizeof(float)));
  uint8_t* aprev = static_cast<uint8_t*>(calloc(n, 1));
  uint8_t* acurr = static_cast<uint8_t*>(calloc(n, 1));
  dim3 grid = ref_grid(h, w), block(8, 8);
  int total = 0, round = 0;
  double inf_norm = 0.0;
  double max_optimal_cost = 5.0 / (1.0 - gamma);               /* :290 */
  do {
    for (int i = 0; i < 25; ++i) {                             /* :295-306 */
      cudaOneStepPolicyEvaluation<<<grid, block>>>(
          h, w, gamma, dev_trans_prob, dev_stage_cost, dev_optimal_cost1,
          dev_optimal_cost2, dev_optimal_action);
      checkCudaErrors(cudaDeviceSynchronize());
      cudaOneStepPolicyEvaluation<<<grid, block>>>(
          h, w, gamma, dev_trans_prob, dev_stage_cost, dev_optimal_cost2,
          dev_optimal_cost1, dev_optimal_action);
      checkCudaErrors(cudaDeviceSynchronize());
    }
    total += 50;
    checkCudaErrors(cudaMemcpy(curr, dev_optimal_cost1, sizeof(float) * n,
                               cudaMemcpyDeviceToHost));       /* :312 */
    float m = 0.0f;
    for (size_t i = 0; i < n; ++i) {
      float d = std::fabs(prev[i] - curr[i]);
      if (d > m) m = d;
    }
    inf_norm = m;
    memcpy(prev, curr, sizeof(float) * n);
    cudaPolicyImprovment<<<grid, block>>>(h, w, gamma, dev_trans_prob, dev_stage_cost,
                                          dev_optimal_cost1, dev_optimal_action);  /* :333 */
    checkCudaErrors(cudaDeviceSynchronize());
    checkCudaErrors(cudaMemcpy(acurr, dev_optimal_action, n, cudaMemcpyDeviceToHost));
    uint32_t diff = 0;
    for (size_t i = 0; i < n; ++i) diff += aprev[i] != acurr[i];
    memcpy(aprev, acurr, n);
    if (residuals && (max_rounds <= 0 || round < max_rounds)) residuals[round] = inf_norm;
    if (changed && (max_rounds <= 0 || round < max_rounds)) changed[round] = diff;
    ++round;
    if (max_rounds > 0 && round >= max_rounds) break;
  } while (inf_norm > max_optimal_cost * 1e-3);                /* :353 */
  checkCudaErrors(cudaMemcpy(J_out, dev_optimal_cost1, sizeof(float) * n, cudaMemcpyDeviceToHost));
  memcpy(action_out, acurr, n);
  free(prev); free(curr); free(aprev); free(acurr);
  freeDeviceMemory();
  return total;
}

/* Time n_sweeps (even) of the reference sweep kernel, launched and
 * synchronised exactly as mdp_host:226-237 does, with CUDA events around the
 * loop.  Model build and upload are outside the timed region.  Returns 0 and
 * the elapsed milliseconds. */
int ref_mdp_time_sweeps(uint32_t h, uint32_t w, const uint8_t* map,
                        uint32_t gx, uint32_t gy, float gamma, int n_sweeps,
                        int warmup_sweeps, float* ms_out) {
  ref_setup(h, w, map, gx, gy);
  dim3 grid = ref_grid(h, w), block(8, 8);
  cudaEvent_t e0, e1;
  checkCudaErrors(cudaEventCreate(&e0));
  checkCudaErrors(cudaEventCreate(&e1));
  for (int phase = 0; phase < 2; ++phase) {
    int n = phase == 0 ? warmup_sweeps : n_sweeps;
    if (phase == 1) checkCudaErrors(cudaEventRecord(e0));
    for (int i = 0; i < n / 2; ++i) {
      cudaOneStepValueIteration<<<grid, block>>>(
          h, w, gamma, dev_trans_prob, dev_stage_cost, dev_optimal_cost1,
          dev_optimal_cost2, dev_optimal_action);
      checkCudaErrors(cudaDeviceSynchronize());
      cudaOneStepValueIteration<<<grid, block>>>(
          h, w, gamma, dev_trans_prob, dev_stage_cost, dev_optimal_cost2,
          dev_optimal_cost1, dev_optimal_action);
      checkCudaErrors(cudaDeviceSynchronize());
    }
  }
  checkCudaErrors(cudaEventRecord(e1));
  checkCudaErrors(cudaEventSynchronize(e1));
  checkCudaErrors(cudaEventElapsedTime(ms_out, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  freeDeviceMemory();
  return 0;
}

}  // extern "C"
