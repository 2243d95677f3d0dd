/*
 * oracle/mdp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's MDP value-iteration path, used only as
 * the parity checker (tests/, __graft_entry__.smoke()) and as bench.py's
 * cpu_baseline / --impl reference arm.  Nothing under path_planning_2d_b200/
 * may call into this file.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/path_planning_2d/):
 *   mdp_cuda = src/mdp/path_planning_2d_cuda.cu
 *   mdp_host = src/mdp/path_planning_2d.cu
 *
 * Arithmetic contract (what "the reference" means numerically): the
 * reference kernels as compiled by nvcc 12.9 for sm_100a with the
 * reference's own flag --use_fast_math (CMakeLists.txt:36-38).  In the SASS
 * of cudaOneStepValueIteration the inner statement
 *     cost += gamma*tp[i]*local_cost_to_go[i];
 * is t = FMUL(gamma, tp[i]); cost = FFMA(t, J[i], cost)   (81 + 81 per
 * thread, no FADD), so this file uses a float multiply followed by fmaf().
 * Build with -ffp-contract=off so the C compiler adds no contraction of its
 * own.  --use_fast_math also sets FTZ; no subnormal can arise on this path
 * (costs are 0 or >= 0.095) so FTZ is not modelled.
 *
 * Parity pin: the reference ships no golden vectors.
 * This restatement is pinned against the UNMODIFIED reference kernels run on
 * a B200 (oracle/_ref, built by oracle/Makefile from the sources in
 * /root/reference); the outputs of that run are committed under
 * tests/golden/ and tests/test_oracle_cpu.py checks this file against
 * them bit for bit.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Thread count of the OpenMP loops below (bench.py: the reference arm sets it
 * explicitly because torchrun exports OMP_NUM_THREADS=1).  n <= 0 leaves the
 * setting alone.  Returns the number of threads a parallel region really gets. */
int oracle_set_threads(int n) {
  int got = 1;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#pragma omp parallel
  {
#pragma omp single
    got = omp_get_num_threads();
  }
#else
  (void)n;
#endif
  return got;
}

#if defined(__x86_64__) && defined(__GNUC__) && !defined(PP2D_ORACLE_NO_CLONES)
#define ORACLE_CLONES __attribute__((target_clones("fma", "default")))
#else
#define ORACLE_CLONES
#endif

/* mdp_cuda:76-150  cudaTransitionProbability (MDP flavour: the trapped-cell
 * override is applied BEFORE the copy into the naive table). */
static void transition_probability(uint8_t u, const uint8_t* map,
                                   float* tp, float* tp_naive) {
  switch (u) {
    case 0: tp[0] = 0.7f; tp[1] = 0.1f; tp[3] = 0.1f; tp[4] = 0.1f; break;
    case 1: tp[0] = 0.1f; tp[1] = 0.7f; tp[2] = 0.1f; tp[4] = 0.1f; break;
    case 2: tp[1] = 0.1f; tp[2] = 0.7f; tp[4] = 0.1f; tp[5] = 0.1f; break;
    case 3: tp[0] = 0.1f; tp[3] = 0.7f; tp[4] = 0.1f; tp[6] = 0.1f; break;
    case 4: tp[4] = 1.0f; break;
    case 5: tp[2] = 0.1f; tp[4] = 0.1f; tp[5] = 0.7f; tp[8] = 0.1f; break;
    case 6: tp[3] = 0.1f; tp[4] = 0.1f; tp[6] = 0.7f; tp[7] = 0.1f; break;
    case 7: tp[4] = 0.1f; tp[6] = 0.1f; tp[7] = 0.7f; tp[8] = 0.1f; break;
    case 8: tp[4] = 0.1f; tp[5] = 0.1f; tp[7] = 0.1f; tp[8] = 0.7f; break;
  }
  if (map[4] == 1) {                       /* mdp_cuda:131-134 */
    for (int i = 0; i < 9; ++i) tp[i] = 0.0f;
    tp[4] = 1.0f;
  }
  memcpy(tp_naive, tp, sizeof(float) * 9); /* mdp_cuda:137 */
  for (int i = 0; i < 9; ++i) {            /* mdp_cuda:142-147 */
    if (map[i] == 1 && i != 4) {
      tp[4] += tp[i];
      tp[i] = 0.0f;
    }
  }
}

/* mdp_cuda:152-172  cudaStageCost.  map_cost is 1 or 2, so the product with
 * the naive probability is exact and fmaf() == FADD of the exact product. */
static void stage_cost_fn(uint32_t x, uint32_t y, uint32_t gx, uint32_t gy,
                          const uint8_t* map, const float* tp_naive,
                          float* stage_cost) {
  float map_cost[9];
  for (int i = 0; i < 9; ++i) map_cost[i] = (map[i] == 1) ? 2.0f : 1.0f;
  for (int u = 0; u < 9; ++u)
    for (int i = 0; i < 9; ++i)
      stage_cost[u] = fmaf(map_cost[i], tp_naive[9 * u + i], stage_cost[u]);
  stage_cost[4] = (x != gx || y != gy) ? 2.0f : 0.0f;
}

/* mdp_cuda:174-213  cudaGenerateModelData: one call per cell. */
void oracle_mdp_generate_model(uint32_t height, uint32_t width,
                               uint32_t gx, uint32_t gy, const uint8_t* map,
                               float* trans_prob, float* stage_cost) {
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)height; ++y) {
    for (int64_t x = 0; x < (int64_t)width; ++x) {
      int64_t idx = y * width + x;
      uint8_t local_map[9];
      int i = 0;
      for (int oy = -1; oy < 2; ++oy)
        for (int ox = -1; ox < 2; ++ox, ++i) {
          int64_t nx = x + ox, ny = y + oy;
          if (nx < 0 || nx >= (int64_t)width || ny < 0 || ny >= (int64_t)height)
            local_map[i] = 1;               /* out of map = occupied */
          else
            local_map[i] = map[ny * width + nx];
        }
      float tp[81] = {0.0f}, tpn[81] = {0.0f};
      for (uint8_t u = 0; u < 9; ++u)
        transition_probability(u, local_map, tp + u * 9, tpn + u * 9);
      memcpy(trans_prob + idx * 81, tp, sizeof(tp));
      float sc[9] = {0.0f};
      stage_cost_fn((uint32_t)x, (uint32_t)y, gx, gy, local_map, tpn, sc);
      memcpy(stage_cost + idx * 9, sc, sizeof(sc));
    }
  }
}

/* mdp_cuda:215-264  cudaOneStepValueIteration: one Jacobi backup of every
 * cell; curr and action are written, prev is read. */
ORACLE_CLONES
void oracle_mdp_sweep(uint32_t height, uint32_t width, float gamma,
                      const float* trans_prob, const float* stage_cost,
                      const float* prev, float* curr, uint8_t* action) {
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)height; ++y) {
    for (int64_t x = 0; x < (int64_t)width; ++x) {
      int64_t idx = y * width + x;
      const float* tp_all = trans_prob + idx * 81;
      const float* sc = stage_cost + idx * 9;
      float ctg[9] = {0.0f};
      int i = 0;
      for (int oy = -1; oy < 2; ++oy)
        for (int ox = -1; ox < 2; ++ox, ++i) {
          int64_t nx = x + ox, ny = y + oy;
          if (nx >= 0 && nx < (int64_t)width && ny >= 0 && ny < (int64_t)height)
            ctg[i] = prev[ny * width + nx];
        }
      float opt_cost = FLT_MAX;
      uint8_t opt_action = 0;
      for (uint8_t u = 0; u < 9; ++u) {
        const float* tp = tp_all + u * 9;
        float cost = sc[u];
        for (int k = 0; k < 9; ++k) {
          float t = gamma * tp[k];          /* FMUL.FTZ */
          cost = fmaf(t, ctg[k], cost);     /* FFMA.FTZ */
        }
        if (cost < opt_cost) { opt_cost = cost; opt_action = u; }
      }
      curr[idx] = opt_cost;
      action[idx] = opt_action;
    }
  }
}

/* mdp_host:243-251: inf-norm of the change of J between two check points
 * (OpenCV absdiff + minMaxIdx on CV_32F, widened to double). */
double oracle_mdp_inf_norm(uint64_t n, const float* a, const float* b) {
  float m = 0.0f;
  for (uint64_t i = 0; i < n; ++i) {
    float d = fabsf(a[i] - b[i]);
    if (d > m) m = d;
  }
  return (double)m;
}

/* mdp_host:207-269 valueIteration + mdp_host:90-126 of initialize():
 * J1 = J2 = 0, action = 0; batches of 100 sweeps (50 ping-pong pairs); stop
 * when the inf-norm over the batch is <= 5.0/(1.0-gamma)*1e-3 evaluated in
 * double on the float gamma.  Returns the number of sweeps; J/action are the
 * downloaded dev_optimal_cost1 / dev_optimal_action.  residuals (optional)
 * receives one inf-norm per batch, up to max_batches entries.
 * max_batches <= 0 means "until converged". */
int oracle_mdp_value_iteration(uint32_t height, uint32_t width,
                               uint32_t gx, uint32_t gy, float gamma,
                               const uint8_t* map, float* J_out,
                               uint8_t* action_out, double* residuals,
                               int max_batches) {
  uint64_t n = (uint64_t)height * width;
  float* tp = (float*)malloc(sizeof(float) * n * 81);
  float* sc = (float*)malloc(sizeof(float) * n * 9);
  float* j1 = (float*)calloc(n, sizeof(float));
  float* j2 = (float*)calloc(n, sizeof(float));
  float* jprev = (float*)calloc(n, sizeof(float));
  if (!tp || !sc || !j1 || !j2 || !jprev) {
    free(tp); free(sc); free(j1); free(j2); free(jprev);
    return -1;
  }
  memset(action_out, 0, n);
  oracle_mdp_generate_model(height, width, gx, gy, map, tp, sc);

  int total = 0, batch = 0;
  double inf_norm = 0.0;
  double max_optimal_cost = 5.0 / (1.0 - gamma);   /* mdp_host:221 */
  do {
    for (int i = 0; i < 50; ++i) {
      oracle_mdp_sweep(height, width, gamma, tp, sc, j1, j2, action_out);
      oracle_mdp_sweep(height, width, gamma, tp, sc, j2, j1, action_out);
    }
    total += 100;
    inf_norm = oracle_mdp_inf_norm(n, jprev, j1);
    memcpy(jprev, j1, sizeof(float) * n);
    if (residuals && (max_batches <= 0 || batch < max_batches))
      residuals[batch] = inf_norm;
    ++batch;
    if (max_batches > 0 && batch >= max_batches) break;
  } while (inf_norm > max_optimal_cost * 1e-3);    /* mdp_host:263 */

  memcpy(J_out, j1, sizeof(float) * n);
  free(tp); free(sc); free(j1); free(j2); free(jprev);
  return total;
}

/* Flush-to-zero of the reference's device arithmetic (--use_fast_math:
 * FMUL.FTZ / FFMA.FTZ flush denormal inputs and results to a signed zero).
 * It matters in policy iteration only: J of the goal decays geometrically to
 * 0 once the policy says "stay" there; value iteration never produces a
 * denormal (J is 0 or >= 1). */
static inline float ftz(float x) {
  return (fabsf(x) < FLT_MIN) ? copysignf(0.0f, x) : x;
}

/* mdp_cuda:266-306 cudaOneStepPolicyEvaluation: one backup under the FIXED
 * policy in `action`.  nvcc 12.9 / sm_100a: t = FMUL.FTZ(J_k, gamma), then
 * cost = FFMA.FTZ(P_k, t, cost) for k = 0..8 -- note (gamma*J)*P here, against
 * (gamma*P)*J in the value-iteration kernel. */
ORACLE_CLONES
void oracle_mdp_policy_evaluation(uint32_t height, uint32_t width, float gamma,
                                  const float* trans_prob, const float* stage_cost,
                                  const float* prev, float* curr, const uint8_t* action) {
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)height; ++y) {
    for (int64_t x = 0; x < (int64_t)width; ++x) {
      int64_t idx = y * width + x;
      const uint8_t u = action[idx];
      const float* tp = trans_prob + idx * 81 + u * 9;
      float cost = stage_cost[idx * 9 + u];
      int i = 0;
      for (int oy = -1; oy < 2; ++oy)
        for (int ox = -1; ox < 2; ++ox, ++i) {
          int64_t nx = x + ox, ny = y + oy;
          float ctg = 0.0f;
          if (nx >= 0 && nx < (int64_t)width && ny >= 0 && ny < (int64_t)height)
            ctg = prev[ny * width + nx];
          float t = ftz(gamma * ftz(ctg));              /* FMUL.FTZ */
          cost = ftz(fmaf(t, ftz(tp[i]), ftz(cost)));   /* FFMA.FTZ */
        }
      curr[idx] = cost;
    }
  }
}

/* mdp_cuda:308-355 cudaPolicyImprovment: greedy action for the given J (first
 * strict minimum); same product order as the evaluation kernel. */
ORACLE_CLONES
void oracle_mdp_policy_improvement(uint32_t height, uint32_t width, float gamma,
                                   const float* trans_prob, const float* stage_cost,
                                   const float* J, uint8_t* action) {
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)height; ++y) {
    for (int64_t x = 0; x < (int64_t)width; ++x) {
      int64_t idx = y * width + x;
      float t[9];
      int i = 0;
      for (int oy = -1; oy < 2; ++oy)
        for (int ox = -1; ox < 2; ++ox, ++i) {
          int64_t nx = x + ox, ny = y + oy;
          float ctg = 0.0f;
          if (nx >= 0 && nx < (int64_t)width && ny >= 0 && ny < (int64_t)height)
            ctg = J[ny * width + nx];
          t[i] = ftz(gamma * ftz(ctg));
        }
      float opt_cost = FLT_MAX;
      uint8_t opt_action = 0;
      for (uint8_t u = 0; u < 9; ++u) {
        const float* tp = trans_prob + idx * 81 + u * 9;
        float cost = stage_cost[idx * 9 + u];
        for (int k = 0; k < 9; ++k) cost = ftz(fmaf(t[k], ftz(tp[k]), ftz(cost)));
        if (cost < opt_cost) { opt_cost = cost; opt_action = u; }
      }
      action[idx] = opt_action;
    }
  }
}

/* mdp_host:271-357 policyIteration (dead code in the reference: its call is
 * commented out, mdp_host:115-116): J1 = J2 = 0, action = 0; rounds of 50
 * evaluation sweeps, the inf-norm of the change of J over the round, one
 * improvement; stop when the inf-norm is <= 5.0/(1.0-gamma)*1e-3.  Returns the
 * number of evaluation sweeps; residuals / changed (optional) get one entry
 * per round (changed = number of actions the improvement altered). */
int oracle_mdp_policy_iteration(uint32_t height, uint32_t width, uint32_t gx, uint32_t gy,
                                float gamma, const uint8_t* map, float* J_out,
                                uint8_t* action_out, double* residuals, uint32_t* changed,
                                int max_rounds) {
  uint64_t n = (uint64_t)height * width;
  float* tp = (float*)malloc(sizeof(float) * n * 81);
  float* sc = (float*)malloc(sizeof(float) * n * 9);
  float* j1 = (float*)calloc(n, sizeof(float));
  float* j2 = (float*)calloc(n, sizeof(float));
  float* jprev = (float*)calloc(n, sizeof(float));
  uint8_t* aprev = (uint8_t*)calloc(n, 1);
  if (!tp || !sc || !j1 || !j2 || !jprev || !aprev) {
    free(tp); free(sc); free(j1); free(j2); free(jprev); free(aprev);
    return -1;
  }
  memset(action_out, 0, n);
  oracle_mdp_generate_model(height, width, gx, gy, map, tp, sc);
  int total = 0, round = 0;
  double inf_norm = 0.0;
  double max_optimal_cost = 5.0 / (1.0 - gamma);
  do {
    for (int i = 0; i < 25; ++i) {
      oracle_mdp_policy_evaluation(height, width, gamma, tp, sc, j1, j2, action_out);
      oracle_mdp_policy_evaluation(height, width, gamma, tp, sc, j2, j1, action_out);
    }
    total += 50;
    inf_norm = oracle_mdp_inf_norm(n, jprev, j1);
    memcpy(jprev, j1, sizeof(float) * n);
    oracle_mdp_policy_improvement(height, width, gamma, tp, sc, j1, action_out);
    uint32_t diff = 0;
    for (uint64_t i = 0; i < n; ++i) diff += aprev[i] != action_out[i];
    memcpy(aprev, action_out, n);
    if (residuals && (max_rounds <= 0 || round < max_rounds)) residuals[round] = inf_norm;
    if (changed && (max_rounds <= 0 || round < max_rounds)) changed[round] = diff;
    ++round;
    if (max_rounds > 0 && round >= max_rounds) break;
  } while (inf_norm > max_optimal_cost * 1e-3);
  memcpy(J_out, j1, sizeof(float) * n);
  free(tp); free(sc); free(j1); free(j2); free(jprev); free(aprev);
  return total;
}

/* mdp_host:168-189 beliefCallback: action at the first strict maximum of the
 * belief (initial mode 0.0f at index 0). */
uint8_t oracle_mdp_plan(uint64_t n, const float* belief,
                        const uint8_t* optimal_action) {
  float mode = 0.0f;
  uint64_t mode_idx = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (belief[i] > mode) { mode_idx = i; mode = belief[i]; }
  return optimal_action[mode_idx];
}

/* Row W of SURVEY.md section 8a (defined by this build; the reference never
 * returns a path): greedy rollout of optimal_action from a start cell using
 * the nominal move of action u, (u%3-1, u/3-1) (mdp_cuda:83-88).  Stops at
 * action 4 (stay), on leaving the map, or after max_len cells.  Returns the
 * number of cell indices y*width+x written, the start cell included. */
uint32_t oracle_mdp_waypoints(uint32_t height, uint32_t width,
                              const uint8_t* optimal_action,
                              uint32_t sx, uint32_t sy,
                              uint32_t* out, uint32_t max_len) {
  uint32_t n = 0;
  int64_t x = sx, y = sy;
  while (n < max_len) {
    if (x < 0 || x >= (int64_t)width || y < 0 || y >= (int64_t)height) break;
    out[n++] = (uint32_t)(y * width + x);
    uint8_t u = optimal_action[y * width + x];
    if (u == 4) break;
    x += (int)(u % 3) - 1;
    y += (int)(u / 3) - 1;
  }
  return n;
}
