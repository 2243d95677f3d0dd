/*
 * oracle/ref_pomdp_driver.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Driver around the reference's POMDP device code:
 *   - src/pomdp/model_generation_cuda.cu is compiled UNMODIFIED (included
 *     below through -I; ROS / std_srvs / boost::shared_ptr headers are
 *     replaced by the declarations in oracle/stubs/);
 *   - the kernels cudaBayesBeliefUpdate
 *     (src/pomdp/point_based_value_iteration_cuda.cu:88-133),
 *     cudaFIBValueIteration (src/pomdp/fast_informed_bound_cuda.cu:97-204) and
 *     cudaInitializeCurandStates / cudaForwardSampling
 *     (src/pomdp/search_tree_cuda.cu:84-147) live in files whose HOST code
 *     needs Boost.MultiArray, which is absent; oracle/Makefile cuts exactly
 *     those line ranges out of the reference files into the generated,
 *     git-ignored oracle/_ref/ref_pomdp_kernels.inc at build time (no
 *     reference source is stored in this repository).
 * Flags: the reference's --use_fast_math, arch sm_100a.
 */
#include <curand_kernel.h>
#include <model_generation_cuda.cu>      /* reference TU, via -I.../src/pomdp */
#include "ref_pomdp_kernels.inc"         /* generated into oracle/_ref/ */

#include <cmath>
#include <cstring>

extern "C" {

/* generateModelData (model_generation_cuda.cu:349-375): tables to host. */
int ref_pomdp_model(uint32_t h, uint32_t w, const uint8_t* map, int32_t gx,
                    int32_t gy, float* trans_prob, float* meas_prob,
                    float* stage_reward) {
  allocateDeviceMemoryOfModel(h, w);
  int32_t goal[2] = {gx, gy};
  generateModelData(h, w, map, goal);
  memcpy(trans_prob, host_trans_prob, sizeof(float) * h * w * 81);
  memcpy(meas_prob, host_meas_prob, sizeof(float) * h * w * 16);
  memcpy(stage_reward, host_stage_reward, sizeof(float) * h * w * 9);
  freeDeviceMemoryOfModel();
  return 0;
}

/* n_updates Bayes updates (u[i], z[i]) applied in sequence to one belief,
 * un-normalised outputs [n_updates][hw], each starting from belief_in
 * (search_tree_cuda.cu:204-223 launch configuration). */
int ref_pomdp_bayes(uint32_t h, uint32_t w, const uint8_t* map, int32_t gx,
                    int32_t gy, const float* belief_in, uint32_t n_updates,
                    const uint8_t* u, const uint8_t* z, float* out) {
  allocateDeviceMemoryOfModel(h, w);
  int32_t goal[2] = {gx, gy};
  generateModelData(h, w, map, goal);
  float *din, *dout;
  size_t n = (size_t)h * w;
  checkCudaErrors(cudaMalloc(&din, n * sizeof(float)));
  checkCudaErrors(cudaMalloc(&dout, n * sizeof(float)));
  checkCudaErrors(cudaMemcpy(din, belief_in, n * sizeof(float), cudaMemcpyHostToDevice));
  dim3 grid((w + 7) / 8, (h + 7) / 8), block(8, 8);
  for (uint32_t i = 0; i < n_updates; ++i) {
    cudaBayesBeliefUpdate<<<grid, block>>>(h, w, dev_trans_prob, dev_meas_prob,
                                           din, u[i], z[i], dout);
    checkCudaErrors(cudaDeviceSynchronize());
    checkCudaErrors(cudaMemcpy(out + (size_t)i * n, dout, n * sizeof(float),
                               cudaMemcpyDeviceToHost));
  }
  cudaFree(din); cudaFree(dout);
  freeDeviceMemoryOfModel();
  return 0;
}

/* fastInformedBound's loop (fast_informed_bound_cuda.cu:206-276) around the
 * reference kernel; alphas [hw][9].  Returns the number of sweeps. */
int ref_pomdp_fib(uint32_t h, uint32_t w, const uint8_t* map, int32_t gx,
                  int32_t gy, float gamma, float* alphas, int max_sweeps) {
  allocateDeviceMemoryOfModel(h, w);
  int32_t goal[2] = {gx, gy};
  generateModelData(h, w, map, goal);
  size_t n = (size_t)h * w * 9;
  float *a1, *a2;
  checkCudaErrors(cudaMalloc(&a1, n * sizeof(float)));
  checkCudaErrors(cudaMalloc(&a2, n * sizeof(float)));
  checkCudaErrors(cudaMemset(a1, 0, n * sizeof(float)));
  checkCudaErrors(cudaMemset(a2, 0, n * sizeof(float)));
  float* prev = (float*)calloc(n, sizeof(float));
  float* curr = (float*)calloc(n, sizeof(float));
  dim3 grid((int)ceil((float)w / 8.0f), (int)ceil((float)h / 8.0f)), block(8, 8);
  int total = 0;
  float inf_norm;
  do {
    for (int i = 0; i < 5; ++i) {
      cudaFIBValueIteration<<<grid, block>>>(h, w, gamma, dev_trans_prob,
                                             dev_meas_prob, dev_stage_reward, a1, a2);
      checkCudaErrors(cudaDeviceSynchronize());
      cudaFIBValueIteration<<<grid, block>>>(h, w, gamma, dev_trans_prob,
                                             dev_meas_prob, dev_stage_reward, a2, a1);
      checkCudaErrors(cudaDeviceSynchronize());
    }
    total += 10;
    checkCudaErrors(cudaMemcpy(curr, a1, n * sizeof(float), cudaMemcpyDeviceToHost));
    inf_norm = 0.0f;
    for (size_t i = 0; i < n; ++i) {
      float d = fabs(prev[i] - curr[i]);
      if (d > inf_norm) inf_norm = d;
    }
    memcpy(prev, curr, n * sizeof(float));
  } while (inf_norm > 0.01f && (max_sweeps <= 0 || total < max_sweeps));
  memcpy(alphas, curr, n * sizeof(float));
  free(prev); free(curr);
  cudaFree(a1); cudaFree(a2);
  freeDeviceMemoryOfModel();
  return total;
}

/* The 2*n uniforms the reference's sampling kernel consumes: XORWOW,
 * curand_init(1234, idx, 0), then curand_uniform twice
 * (search_tree_cuda.cu:84-92, 117, 134). */
__global__ void ref_uniforms_kernel(uint32_t n, float* out) {
  uint32_t idx = blockDim.x * blockIdx.x + threadIdx.x;
  if (idx >= n) return;
  curandState st;
  curand_init(1234, idx, 0, &st);
  out[2 * idx] = curand_uniform(&st);
  out[2 * idx + 1] = curand_uniform(&st);
}
int ref_pomdp_uniforms(uint32_t n, float* out) {
  float* d;
  checkCudaErrors(cudaMalloc(&d, 2 * n * sizeof(float)));
  ref_uniforms_kernel<<<(n + 31) / 32, 32>>>(n, d);
  checkCudaErrors(cudaDeviceSynchronize());
  checkCudaErrors(cudaMemcpy(out, d, 2 * n * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

/* forwardSampling's device half with the reference kernels: given the
 * state samples, returns the observations (search_tree_cuda.cu:314-358). */
int ref_pomdp_forward_sampling(uint32_t h, uint32_t w, const uint8_t* map,
                               int32_t gx, int32_t gy, uint32_t n,
                               const uint32_t* samples, uint8_t action,
                               uint8_t* observations) {
  allocateDeviceMemoryOfModel(h, w);
  int32_t goal[2] = {gx, gy};
  generateModelData(h, w, map, goal);
  curandState* st; uint32_t* ds; uint8_t* dobs;
  checkCudaErrors(cudaMalloc(&st, sizeof(curandState) * n));
  checkCudaErrors(cudaMalloc(&ds, sizeof(uint32_t) * n));
  checkCudaErrors(cudaMalloc(&dobs, n));
  dim3 grid((n + 31) / 32), block(32);
  cudaInitializeCurandStates<<<grid, block>>>(n, st);
  checkCudaErrors(cudaDeviceSynchronize());
  checkCudaErrors(cudaMemcpy(ds, samples, sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
  cudaForwardSampling<<<grid, block>>>(n, h, w, st, dev_trans_prob, dev_meas_prob,
                                       ds, action, dobs);
  checkCudaErrors(cudaDeviceSynchronize());
  checkCudaErrors(cudaMemcpy(observations, dobs, n, cudaMemcpyDeviceToHost));
  cudaFree(st); cudaFree(ds); cudaFree(dobs);
  freeDeviceMemoryOfModel();
  return 0;
}

}  // extern "C"
