/*
 * oracle/pomdp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's POMDP / QV-Tree path (SURVEY.md rows
 * B1-B10), used only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Nothing under path_planning_2d_b200/
 * may call into this file.
 *
 * Reference files (relative to /root/reference/path_planning_2d/):
 *   model_gen = src/pomdp/model_generation_cuda.cu
 *   pbvi      = src/pomdp/point_based_value_iteration_cuda.cu
 *   fib       = src/pomdp/fast_informed_bound_cuda.cu
 *   tree      = src/pomdp/search_tree_cuda.cu
 *   tree_h    = include/path_planning_2d/search_tree.h
 *   pomdp     = src/pomdp/path_planning_2d.cu
 *
 * Arithmetic contract:
 *   - device code of the reference is compiled with --use_fast_math
 *     (CMakeLists.txt:36-38): FMUL/FFMA contraction and flush-to-zero.
 *     Beliefs do reach the subnormal range, so the device-side functions
 *     here (bayes update, forward sampling) run with the SSE FTZ+DAZ bits
 *     set and use fmaf() where nvcc contracts (p += a*b).
 *   - host code of the reference (normalisation, inner products, tree
 *     bookkeeping) is plain x86-64 g++: sequential float mul then add, no
 *     FMA.  Build this file with -ffp-contract=off.
 *
 * Parity pin: every function here is checked against the reference's OWN code
 * run on a B200.  Kernels (B1, B2, FIB sweep, sampling): oracle/_ref/
 * libpp2d_ref_pomdp.so -> tests/golden/pomdp_*.npz.  Host-side tree logic
 * (B3-B10: SearchTree, VNode, QNode, evaluateFibCpu, evaluatePbviCpu): the
 * four reference translation units compiled unmodified against stand-in
 * headers for ROS and Boost.MultiArray (oracle/_ref/libpp2d_ref_pomdp_full.so,
 * oracle/stubs/) -> tests/golden/tree_*.npz, compared node by node and bit for
 * bit by tests/test_tree_pin_cpu.py.  See DESIGN.md section 2.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <xmmintrin.h>
#define FTZ_ON()  unsigned int csr_saved_ = _mm_getcsr(); _mm_setcsr(csr_saved_ | 0x8040)
#define FTZ_OFF() _mm_setcsr(csr_saved_)
#else
#define FTZ_ON()
#define FTZ_OFF()
#endif

#if defined(__x86_64__) && defined(__GNUC__) && !defined(PP2D_ORACLE_NO_CLONES)
#define ORACLE_CLONES __attribute__((target_clones("fma", "default")))
#else
#define ORACLE_CLONES
#endif

/* ------------------------------------------------------------------ B1 -- */
/* model_gen:161-233 cudaTransitionProbability (POMDP flavour: the naive copy
 * is taken BEFORE the trapped-cell override). */
static void transition_probability(uint8_t u, const uint8_t* map, float* tp,
                                   float* tp_naive) {
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
  memcpy(tp_naive, tp, sizeof(float) * 9);            /* model_gen:213 */
  for (int i = 0; i < 9; ++i) {                       /* model_gen:218-223 */
    if (map[i] == 1 && i != 4) {
      tp[4] += tp[i];
      tp[i] = 0.0f;
    }
  }
  if (map[4] == 1) {                                  /* model_gen:229-232 */
    for (int i = 0; i < 9; ++i) tp[i] = 0.0f;
    tp[4] = 1.0f;
  }
}

/* model_gen:235-263 cudaMeasurementLikelihood.  0.98 / 0.02 are double
 * literals converted to float per use; the four factors are multiplied left
 * to right in float. */
static void measurement_likelihood(const uint8_t* map, float* meas_prob) {
  uint8_t m[4] = {map[1], map[3], map[5], map[7]};
  for (uint8_t i = 0; i < 16; ++i) {
    float l0 = ((i >> 0) & 1) == m[0] ? 0.98 : 0.02;
    float l1 = ((i >> 1) & 1) == m[1] ? 0.98 : 0.02;
    float l2 = ((i >> 2) & 1) == m[2] ? 0.98 : 0.02;
    float l3 = ((i >> 3) & 1) == m[3] ? 0.98 : 0.02;
    meas_prob[i] = l0 * l1 * l2 * l3;
  }
}

/* model_gen:265-296 cudaStageReward (-1 free, -2 occupied; stay = -2, 0 at
 * the goal).  The rewards are -1 / -2, the product is exact. */
static void stage_reward_fn(uint32_t x, uint32_t y, int32_t gx, int32_t gy,
                            const uint8_t* map, const float* tp_naive,
                            float* stage_reward) {
  float map_reward[9];
  for (int i = 0; i < 9; ++i) map_reward[i] = (map[i] == 1) ? -2.0f : -1.0f;
  for (int u = 0; u < 9; ++u)
    for (int i = 0; i < 9; ++i)
      stage_reward[u] = fmaf(map_reward[i], tp_naive[9 * u + i], stage_reward[u]);
  stage_reward[4] = ((int32_t)x != gx || (int32_t)y != gy) ? -2.0f : 0.0f;
}

/* model_gen:298-347 cudaGenerateModelData. */
void oracle_pomdp_generate_model(uint32_t height, uint32_t width, int32_t gx,
                                 int32_t gy, const uint8_t* map,
                                 float* trans_prob, float* meas_prob,
                                 float* stage_reward) {
  for (int64_t y = 0; y < (int64_t)height; ++y) {
    for (int64_t x = 0; x < (int64_t)width; ++x) {
      int64_t idx = y * width + x;
      uint8_t local_map[9];
      int i = 0;
      for (int oy = -1; oy < 2; ++oy)
        for (int ox = -1; ox < 2; ++ox, ++i) {
          int64_t nx = x + ox, ny = y + oy;
          if (nx < 0 || nx >= (int64_t)width || ny < 0 || ny >= (int64_t)height)
            local_map[i] = 1;
          else
            local_map[i] = map[ny * width + nx];
        }
      float tp[81] = {0.0f}, tpn[81] = {0.0f};
      for (uint8_t u = 0; u < 9; ++u)
        transition_probability(u, local_map, tp + u * 9, tpn + u * 9);
      memcpy(trans_prob + idx * 81, tp, sizeof(tp));
      float mp[16] = {0.0f};
      measurement_likelihood(local_map, mp);
      memcpy(meas_prob + idx * 16, mp, sizeof(mp));
      float sr[9] = {0.0f};
      stage_reward_fn((uint32_t)x, (uint32_t)y, gx, gy, local_map, tpn, sr);
      memcpy(stage_reward + idx * 9, sr, sizeof(sr));
    }
  }
}

/* ------------------------------------------------------------------ B2 -- */
/* pbvi:88-133 cudaBayesBeliefUpdate: gather-form predict + update,
 * un-normalised.  Arithmetic as compiled for sm_100a (see the comment at the
 * accumulation below). */
ORACLE_CLONES
void oracle_pomdp_bayes_update(uint32_t height, uint32_t width,
                               const float* trans_prob, const float* meas_prob,
                               const float* belief_in, uint8_t u, uint8_t z,
                               float* belief_out) {
  FTZ_ON();
  for (int64_t y = 0; y < (int64_t)height; ++y) {
    for (int64_t x = 0; x < (int64_t)width; ++x) {
      int64_t idx = y * width + x;
      float ltp[9] = {0.0f}, lb[9] = {0.0f};
      int s = 0;
      for (int oy = -1; oy < 2; ++oy)
        for (int ox = -1; ox < 2; ++ox, ++s) {
          int64_t sx = x + ox, sy = y + oy;
          if (sx < 0 || sx >= (int64_t)width || sy < 0 || sy >= (int64_t)height)
            continue;
          int64_t sidx = sy * width + sx;
          ltp[s] = trans_prob[81 * sidx + 9 * u + (8 - s)];
          lb[s] = belief_in[sidx];
        }
      /* SASS of the reference kernel (nvcc 12.9, sm_100a, --use_fast_math):
       * slots 0..7 are an FFMA.FTZ chain, the last slot is a predicated
       * FMUL.FTZ whose ROUNDED product is added with FADD.FTZ, then the
       * likelihood is applied with FMUL.FTZ. */
      float p = 0.0f;
      for (int k = 0; k < 8; ++k) p = fmaf(ltp[k], lb[k], p);
      {
        volatile float last = ltp[8] * lb[8];
        p = p + last;
      }
      p *= meas_prob[16 * idx + z];
      belief_out[idx] = p;
    }
  }
  FTZ_OFF();
}

/* ------------------------------------------------------------------ B3 -- */
/* tree:226-229 / tree:609-612: sequential float accumulate from 0.0f, then a
 * float division of every element.  Returns the sum. */
float oracle_pomdp_normalize(uint64_t n, float* b) {
  float sum = 0.0f;
  for (uint64_t i = 0; i < n; ++i) sum = sum + b[i];
  for (uint64_t i = 0; i < n; ++i) b[i] /= sum;
  return sum;
}

/* ------------------------------------------------------------- B4 / B5 -- */
/* fib:278-297 evaluateFibCpu: 9 stride-9 inner products, first maximum. */
void oracle_pomdp_evaluate_fib(uint64_t hw, const float* belief,
                               const float* fib_alphas,
                               const uint8_t* fib_actions, float* value,
                               uint8_t* action) {
  float v[9];
  for (int a = 0; a < 9; ++a) {
    float acc = 0.0f;
    for (uint64_t s = 0; s < hw; ++s) acc = acc + belief[s] * fib_alphas[s * 9 + a];
    v[a] = acc;
  }
  int best = 0;
  for (int a = 1; a < 9; ++a)
    if (v[best] < v[a]) best = a;          /* std::max_element: first max */
  *value = v[best];
  *action = fib_actions ? fib_actions[best] : (uint8_t)best;
}

/* pbvi:678-699 evaluatePbviCpu: n inner products over [n][hw], first max. */
void oracle_pomdp_evaluate_pbvi(uint64_t hw, const float* belief,
                                const float* pbvi_alphas, uint32_t n,
                                const uint8_t* pbvi_actions, float* value,
                                uint8_t* action) {
  float best_v = 0.0f;
  uint32_t best = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const float* al = pbvi_alphas + (uint64_t)i * hw;
    float acc = 0.0f;
    for (uint64_t s = 0; s < hw; ++s) acc = acc + belief[s] * al[s];
    if (i == 0 || best_v < acc) { best_v = acc; best = i; }
  }
  *value = best_v;
  *action = pbvi_actions ? pbvi_actions[best] : 0;
}

/* -------------------------------------------------- FIB solver ("next") -- */
/* fib:97-204 cudaFIBValueIteration for one sweep (device code: FMUL/FFMA
 * contraction of  acc += a*b, FTZ). */
ORACLE_CLONES
static void fib_sweep(uint32_t height, uint32_t width, float gamma,
                      const float* trans_prob, const float* meas_prob,
                      const float* stage_reward, const float* prev,
                      float* curr) {
  FTZ_ON();
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)height; ++y) {
    for (int64_t x = 0; x < (int64_t)width; ++x) {
      int64_t idx = y * width + x;
      float lmp[144] = {0.0f}, lpa[81] = {0.0f};
      int i = 0;
      for (int oy = -1; oy < 2; ++oy)
        for (int ox = -1; ox < 2; ++ox, ++i) {
          int64_t nx = x + ox, ny = y + oy;
          if (nx < 0 || nx >= (int64_t)width || ny < 0 || ny >= (int64_t)height)
            continue;
          int64_t nidx = ny * width + nx;
          memcpy(lmp + 16 * i, meas_prob + 16 * nidx, sizeof(float) * 16);
          memcpy(lpa + 9 * i, prev + 9 * nidx, sizeof(float) * 9);
        }
      float out[9];
      for (int a = 0; a < 9; ++a) {
        const float* ltp = trans_prob + 81 * idx + 9 * a;
        float reward = stage_reward[idx * 9 + a];
        float rtg = 0.0f;
        for (int o = 0; o < 16; ++o) {
          float ltm[9];
          for (int sp = 0; sp < 9; ++sp) ltm[sp] = ltp[sp] * lmp[sp * 16 + o];
          float rtg_o = -FLT_MAX;
          for (int ap = 0; ap < 9; ++ap) {
            float acc = 0.0f;
            for (int sp = 0; sp < 9; ++sp) acc = fmaf(ltm[sp], lpa[sp * 9 + ap], acc);
            if (rtg_o < acc) rtg_o = acc;
          }
          rtg += rtg_o;
        }
        out[a] = fmaf(gamma, rtg, reward);     /* reward + gamma*rtg -> FFMA */
      }
      memcpy(curr + 9 * idx, out, sizeof(out));
    }
  }
  FTZ_OFF();
}

/* fib:206-276 fastInformedBound: batches of 10 sweeps until the inf-norm of
 * the change is <= 0.01.  alphas: [hw][9], zero initialised by the callee.
 * Returns the number of sweeps. */
int oracle_pomdp_fib_solve(uint32_t height, uint32_t width, float gamma,
                           const float* trans_prob, const float* meas_prob,
                           const float* stage_reward, float* alphas,
                           int max_sweeps) {
  uint64_t n = (uint64_t)height * width * 9;
  float* a1 = (float*)calloc(n, sizeof(float));
  float* a2 = (float*)calloc(n, sizeof(float));
  float* prev = (float*)calloc(n, sizeof(float));
  int total = 0;
  float inf_norm;
  do {
    for (int i = 0; i < 5; ++i) {
      fib_sweep(height, width, gamma, trans_prob, meas_prob, stage_reward, a1, a2);
      fib_sweep(height, width, gamma, trans_prob, meas_prob, stage_reward, a2, a1);
    }
    total += 10;
    inf_norm = 0.0f;
    for (uint64_t i = 0; i < n; ++i) {
      float d = fabsf(prev[i] - a1[i]);
      if (d > inf_norm) inf_norm = d;
    }
    memcpy(prev, a1, n * sizeof(float));
  } while (inf_norm > 0.01f && (max_sweeps <= 0 || total < max_sweeps));
  memcpy(alphas, a1, n * sizeof(float));
  free(a1); free(a2); free(prev);
  return total;
}

/* ------------------------------------------------------- glibc rand() -- */
/* The planner never calls srand(): glibc's TYPE_3 additive-feedback
 * generator with seed 1 (tree:332).  Restated so that every query of a batch
 * can own a private stream "as if started in a fresh process". */
typedef struct {
  int32_t r[34];
  int k;
} glibc_rand_t;

void oracle_glibc_srand(glibc_rand_t* g, uint32_t seed) {
  int32_t r[344];
  if (seed == 0) seed = 1;
  r[0] = (int32_t)seed;
  for (int i = 1; i < 31; ++i) {
    int64_t v = (16807LL * r[i - 1]) % 2147483647LL;
    if (v < 0) v += 2147483647LL;
    r[i] = (int32_t)v;
  }
  for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
  for (int i = 34; i < 344; ++i)
    r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
  for (int i = 0; i < 34; ++i) g->r[i] = r[310 + i];
  g->k = 0;
}

uint32_t oracle_glibc_rand(glibc_rand_t* g) {
  /* ring of the last 34 values: new = r[-31] + r[-3] */
  int k = g->k;
  uint32_t v = (uint32_t)g->r[(k + 34 - 31) % 34] + (uint32_t)g->r[(k + 34 - 3) % 34];
  g->r[k] = (int32_t)v;
  g->k = (k + 1) % 34;
  return v >> 1;
}

/* ------------------------------------------------------------------ B7 -- */
/* tree:311-366 QNode::forwardSampling + tree:94-147 cudaForwardSampling.
 * uniforms: 2*sample_num values of curand_uniform for XORWOW states
 * curand_init(1234, idx, 0) -- (u_next_state, u_observation) per sample idx;
 * the reference re-initialises the states on every call, so they are the same
 * 100 numbers for every Q node (tests/golden/curand_xorwow_1234.npy).
 * A draw that falls beyond the last prefix sum indexes one past the belief in
 * the reference (undefined behaviour); here it is clamped to the last cell.
 * Returns the number of clamped samples. */
int oracle_pomdp_forward_sampling(uint32_t height, uint32_t width,
                                  const float* trans_prob,
                                  const float* meas_prob, const float* belief,
                                  uint8_t action, uint32_t sample_num,
                                  const float* uniforms, glibc_rand_t* rng,
                                  uint8_t* observations) {
  uint64_t hw = (uint64_t)height * width;
  float* dist = (float*)malloc(hw * sizeof(float));
  float acc = 0.0f;
  for (uint64_t i = 0; i < hw; ++i) { acc = acc + belief[i]; dist[i] = acc; }
  int clamped = 0;
  for (uint32_t i = 0; i < sample_num; ++i) {
    float sample_rand = (float)oracle_glibc_rand(rng) / ((float)2147483647 + 1.0f);
    uint64_t sample1 = hw;
    for (uint64_t s = 0; s < hw; ++s)
      if (dist[s] >= sample_rand) { sample1 = s; break; }
    if (sample1 >= hw) { sample1 = hw - 1; ++clamped; }
    /* device part, FTZ adds */
    FTZ_ON();
    float td[9];
    memcpy(td, trans_prob + sample1 * 81 + (uint64_t)action * 9, sizeof(td));
    for (int k = 1; k < 9; ++k) td[k] += td[k - 1];
    float r2 = uniforms[2 * i];
    uint32_t sample2 = 0;
    for (int k = 0; k < 9; ++k)
      if (r2 <= td[k]) { sample2 = (uint32_t)k; break; }
    int64_t next = (int64_t)sample1 + ((int64_t)(sample2 / 3) - 1) * width +
                   ((int64_t)(sample2 % 3) - 1);
    if (next < 0) next = 0;                 /* cannot happen for valid models */
    if ((uint64_t)next >= hw) next = hw - 1;
    float md[16];
    memcpy(md, meas_prob + 16 * next, sizeof(md));
    for (int k = 1; k < 16; ++k) md[k] += md[k - 1];
    float r3 = uniforms[2 * i + 1];
    uint8_t obs = 0;
    for (int k = 0; k < 16; ++k)
      if (r3 <= md[k]) { obs = (uint8_t)k; break; }
    FTZ_OFF();
    observations[i] = obs;
  }
  free(dist);
  return clamped;
}

/* ------------------------------------------------------ B6, B8, B9, B10 -- */
typedef struct VNode VNode;
typedef struct QNode QNode;

typedef struct {
  uint32_t height, width;
  float gamma;
  const float *trans_prob, *meas_prob, *stage_reward;
  const float* fib_alphas;      /* [hw][9] */
  const uint8_t* fib_actions;   /* [9] */
  const float* pbvi_alphas;     /* [n_pbvi][hw] */
  const uint8_t* pbvi_actions;  /* [n_pbvi] */
  uint32_t n_pbvi;
  const float* uniforms;        /* 100 cuRAND uniforms */
  glibc_rand_t rng;
  /* statistics */
  uint64_t n_vnodes, n_qnodes, n_bayes, n_clamped;
} pomdp_ctx;

struct QNode {                   /* tree_h:30-80 */
  float* belief;
  uint8_t action;
  VNode* parent;
  VNode** children;
  int n_children;
  float upper_bound, lower_bound, heuristic, reward;
  VNode* vnode_to_expand;
  uint32_t depth;
};

struct VNode {                   /* tree_h:82-128 */
  float* belief;
  uint8_t observation;
  QNode* parent;
  QNode** children;
  int n_children;
  float upper_bound, lower_bound, heuristic;
  VNode* vnode_to_expand;
  float weight;
  uint32_t depth;
};

static void qnode_update(pomdp_ctx* c, QNode* q);

/* tree:368-388 VNode::VNode */
static VNode* vnode_new(pomdp_ctx* c, const float* b, uint8_t z, float w, QNode* p) {
  uint64_t hw = (uint64_t)c->height * c->width;
  VNode* v = (VNode*)calloc(1, sizeof(VNode));
  v->belief = (float*)malloc(hw * sizeof(float));
  memcpy(v->belief, b, hw * sizeof(float));
  v->observation = z;
  v->weight = w;
  v->parent = p;
  v->depth = 0;
  uint8_t dummy;
  oracle_pomdp_evaluate_fib(hw, v->belief, c->fib_alphas, c->fib_actions,
                            &v->upper_bound, &dummy);
  oracle_pomdp_evaluate_pbvi(hw, v->belief, c->pbvi_alphas, c->n_pbvi,
                             c->pbvi_actions, &v->lower_bound, &dummy);
  v->heuristic = v->upper_bound - v->lower_bound;
  v->vnode_to_expand = v;
  c->n_vnodes++;
  return v;
}

/* tree:161-242 QNode::QNode */
static QNode* qnode_new(pomdp_ctx* c, const float* b, uint8_t a, VNode* p) {
  uint64_t hw = (uint64_t)c->height * c->width;
  QNode* q = (QNode*)calloc(1, sizeof(QNode));
  q->belief = (float*)malloc(hw * sizeof(float));
  memcpy(q->belief, b, hw * sizeof(float));
  q->action = a;
  q->parent = p;
  q->upper_bound = FLT_MAX;
  q->lower_bound = -FLT_MAX;
  q->heuristic = FLT_MIN;
  q->depth = 1;
  c->n_qnodes++;
  /* tree:168-173 reward = <b, R(:,a)>, sequential float, stride 9 */
  float r = 0.0f;
  for (uint64_t s = 0; s < hw; ++s) r = r + q->belief[s] * c->stage_reward[s * 9 + a];
  q->reward = r;
  /* tree:176-195 sample 50 observations, unique + frequency (std::set order) */
  enum { SAMPLES = 50 };
  uint8_t obs[SAMPLES];
  c->n_clamped += oracle_pomdp_forward_sampling(
      c->height, c->width, c->trans_prob, c->meas_prob, q->belief, a, SAMPLES,
      c->uniforms, &c->rng, obs);
  int count[16] = {0};
  for (int i = 0; i < SAMPLES; ++i) count[obs[i] & 15]++;
  int n_unique = 0;
  for (int z = 0; z < 16; ++z) n_unique += count[z] > 0;
  q->children = (VNode**)calloc(n_unique, sizeof(VNode*));
  q->n_children = n_unique;
  float* out = (float*)malloc(hw * sizeof(float));
  int ci = 0;
  for (int z = 0; z < 16; ++z) {               /* tree:213-232 */
    if (!count[z]) continue;
    float weight = (float)count[z] / (float)SAMPLES;
    oracle_pomdp_bayes_update(c->height, c->width, c->trans_prob, c->meas_prob,
                              q->belief, a, (uint8_t)z, out);
    c->n_bayes++;
    oracle_pomdp_normalize(hw, out);
    q->children[ci++] = vnode_new(c, out, (uint8_t)z, weight, q);
  }
  free(out);
  qnode_update(c, q);
  return q;
}

/* tree:251-286 QNode::update */
static void qnode_update(pomdp_ctx* c, QNode* q) {
  float up = 0.0f, lo = 0.0f;
  for (int i = 0; i < q->n_children; ++i) {
    up += q->children[i]->upper_bound * q->children[i]->weight;
    lo += q->children[i]->lower_bound * q->children[i]->weight;
  }
  q->upper_bound = q->reward + c->gamma * up;
  q->lower_bound = q->reward + c->gamma * lo;
  q->heuristic = 0.0f;
  for (int i = 0; i < q->n_children; ++i) {
    VNode* v = q->children[i];
    float h = c->gamma * v->weight * v->heuristic;
    if (h > q->heuristic) {
      q->heuristic = h;
      q->vnode_to_expand = v->vnode_to_expand;
    }
  }
  uint32_t child_depth = 0;
  for (int i = 0; i < q->n_children; ++i)
    if (q->children[i]->depth > child_depth) {
      child_depth = q->children[i]->depth;
      q->depth = child_depth + 1;
    }
}

/* tree:397-435 VNode::update */
static void vnode_update(VNode* v) {
  int umax = 0, lmax = 0;
  for (int i = 1; i < v->n_children; ++i) {
    if (v->children[umax]->upper_bound < v->children[i]->upper_bound) umax = i;
    if (v->children[lmax]->lower_bound < v->children[i]->lower_bound) lmax = i;
  }
  v->upper_bound = v->children[umax]->upper_bound;
  v->lower_bound = v->children[lmax]->lower_bound;
  v->heuristic = -FLT_MAX;
  for (int i = 0; i < v->n_children; ++i) {
    QNode* q = v->children[i];
    if (q->upper_bound <= v->lower_bound) continue;
    if (q->heuristic > v->heuristic) {
      v->heuristic = q->heuristic;
      v->vnode_to_expand = q->vnode_to_expand;
    }
  }
  uint32_t child_depth = 0;
  for (int i = 0; i < v->n_children; ++i)
    if (v->children[i]->depth > child_depth) {
      child_depth = v->children[i]->depth;
      v->depth = child_depth + 1;
    }
}

/* tree:437-450 VNode::expand */
static void vnode_expand(pomdp_ctx* c, VNode* v) {
  v->children = (QNode**)calloc(9, sizeof(QNode*));
  v->n_children = 9;
  for (uint8_t a = 0; a < 9; ++a) v->children[a] = qnode_new(c, v->belief, a, v);
  vnode_update(v);
}

static void free_vnode(VNode* v);
static void free_qnode(QNode* q) {
  if (!q) return;
  for (int i = 0; i < q->n_children; ++i) free_vnode(q->children[i]);
  free(q->children);
  free(q->belief);
  free(q);
}
static void free_vnode(VNode* v) {
  if (!v) return;
  for (int i = 0; i < v->n_children; ++i) free_qnode(v->children[i]);
  free(v->children);
  free(v->belief);
  free(v);
}

typedef struct {
  pomdp_ctx ctx;
  VNode* root;
} pomdp_tree;

/* Context + tree:479-482 SearchTree::SearchTree.  All table pointers are
 * borrowed and must outlive the tree. */
pomdp_tree* oracle_pomdp_tree_create(
    uint32_t height, uint32_t width, float gamma, const float* trans_prob,
    const float* meas_prob, const float* stage_reward, const float* fib_alphas,
    const uint8_t* fib_actions, const float* pbvi_alphas,
    const uint8_t* pbvi_actions, uint32_t n_pbvi, const float* uniforms,
    uint32_t rand_seed, const float* belief) {
  pomdp_tree* t = (pomdp_tree*)calloc(1, sizeof(pomdp_tree));
  pomdp_ctx* c = &t->ctx;
  c->height = height; c->width = width; c->gamma = gamma;
  c->trans_prob = trans_prob; c->meas_prob = meas_prob;
  c->stage_reward = stage_reward; c->fib_alphas = fib_alphas;
  c->fib_actions = fib_actions; c->pbvi_alphas = pbvi_alphas;
  c->pbvi_actions = pbvi_actions; c->n_pbvi = n_pbvi; c->uniforms = uniforms;
  oracle_glibc_srand(&c->rng, rand_seed);
  t->root = vnode_new(c, belief, 0, 0.0f, NULL);
  return t;
}

void oracle_pomdp_tree_destroy(pomdp_tree* t) {
  if (!t) return;
  free_vnode(t->root);
  free(t);
}

uint32_t oracle_pomdp_tree_depth(const pomdp_tree* t) { return t->root->depth; }

/* tree:490-508 SearchTree::expand.  Returns 0, or -1 where the reference
 * would dereference a null vnode_to_expand. */
int oracle_pomdp_tree_expand(pomdp_tree* t) {
  VNode* v = t->root->vnode_to_expand;
  if (!v) return -1;
  if (v->n_children) {            /* re-expansion leaks in the reference */
    for (int i = 0; i < v->n_children; ++i) free_qnode(v->children[i]);
    free(v->children);
    v->children = NULL;
    v->n_children = 0;
  }
  vnode_expand(&t->ctx, v);
  while (v->parent != NULL) {
    QNode* pq = v->parent;
    qnode_update(&t->ctx, pq);
    VNode* pv = pq->parent;
    vnode_update(pv);
    v = pv;
  }
  return 0;
}

/* tree:510-524 SearchTree::getOptimalAction */
void oracle_pomdp_tree_best_action(const pomdp_tree* t, uint8_t* a, float* r) {
  *a = 0;
  *r = -FLT_MAX;
  for (int i = 0; i < t->root->n_children; ++i) {
    const QNode* q = t->root->children[i];
    if (q->upper_bound > *r) { *r = q->upper_bound; *a = q->action; }
  }
}

/* tree:548-626 SearchTree::update(a, z): re-root. */
int oracle_pomdp_tree_update(pomdp_tree* t, uint8_t a, uint8_t z) {
  pomdp_ctx* c = &t->ctx;
  VNode* root = t->root;
  if (root->n_children == 0) return -1;       /* reference: null deref */
  QNode* root_q = NULL;
  for (int i = 0; i < root->n_children; ++i) {
    if (root->children[i]->action == a) root_q = root->children[i];
    else free_qnode(root->children[i]);
  }
  if (!root_q) return -1;
  VNode* root_v = NULL;
  for (int i = 0; i < root_q->n_children; ++i) {
    if (root_q->children[i]->observation == z) root_v = root_q->children[i];
    else free_vnode(root_q->children[i]);
  }
  uint64_t hw = (uint64_t)c->height * c->width;
  if (root_v == NULL) {
    float* cur = (float*)malloc(hw * sizeof(float));
    oracle_pomdp_bayes_update(c->height, c->width, c->trans_prob, c->meas_prob,
                              root->belief, a, z, cur);
    c->n_bayes++;
    oracle_pomdp_normalize(hw, cur);
    root_v = vnode_new(c, cur, 0, 0.0f, NULL);
    free(cur);
  } else {
    root_v->parent = NULL;
  }
  free(root_q->children); free(root_q->belief); free(root_q);
  free(root->children); free(root->belief); free(root);
  t->root = root_v;
  return 0;
}

/* pomdp:199-241 beliefCallback for a fresh tree: root + up to max_iter
 * expansions while depth < max_depth (uint8_t counter), then the action with
 * the largest upper bound.  stats (optional): [n_vnodes, n_qnodes, n_bayes,
 * n_clamped, depth]. */
int oracle_pomdp_plan(pomdp_tree* t, uint32_t max_depth, uint32_t max_iter,
                      uint8_t* action, float* value, uint64_t* stats) {
  uint8_t counter = 0;
  int rc = 0;
  while (oracle_pomdp_tree_depth(t) < max_depth && counter++ < max_iter) {
    rc = oracle_pomdp_tree_expand(t);
    if (rc) break;
  }
  oracle_pomdp_tree_best_action(t, action, value);
  if (stats) {
    stats[0] = t->ctx.n_vnodes; stats[1] = t->ctx.n_qnodes;
    stats[2] = t->ctx.n_bayes; stats[3] = t->ctx.n_clamped;
    stats[4] = oracle_pomdp_tree_depth(t);
  }
  return rc;
}

/* Root statistics for parity diagnostics. */
void oracle_pomdp_tree_root_bounds(const pomdp_tree* t, float* upper, float* lower) {
  *upper = t->root->upper_bound;
  *lower = t->root->lower_bound;
}

/* Per-root-action (Q node) upper/lower bounds; returns the child count. */
int oracle_pomdp_tree_root_q(const pomdp_tree* t, float* upper, float* lower,
                             float* reward) {
  for (int i = 0; i < t->root->n_children; ++i) {
    upper[i] = t->root->children[i]->upper_bound;
    lower[i] = t->root->children[i]->lower_bound;
    reward[i] = t->root->children[i]->reward;
  }
  return t->root->n_children;
}

/* Pre-order dump of the whole tree (the information SearchTree::print shows,
 * tree:288-309, 452-473, 628-633): 9 floats per node = kind (0 = V, 1 = Q),
 * observation | action, weight | reward, upper, lower, heuristic, depth,
 * #children, pre-order id of vnode_to_expand (-1 = NULL or not in the tree).
 * Same format as oracle/ref_pomdp_full_driver.cu:ref_full_tree_dump and
 * pp2d_tree_dump.  Returns the node count. */
typedef struct { const void** key; int n, cap; } idmap_t;
static void idmap_add(idmap_t* m, const void* k) {
  if (m->n == m->cap) {
    m->cap = m->cap ? 2 * m->cap : 256;
    m->key = (const void**)realloc((void*)m->key, (size_t)m->cap * sizeof(void*));
  }
  m->key[m->n++] = k;
}
static float idmap_find(const idmap_t* m, const void* k) {
  if (!k) return -1.0f;
  for (int i = 0; i < m->n; ++i) if (m->key[i] == k) return (float)i;
  return -1.0f;
}
static void number_v(const VNode* v, idmap_t* m);
static void number_q(const QNode* q, idmap_t* m) {
  idmap_add(m, q);
  for (int i = 0; i < q->n_children; ++i) number_v(q->children[i], m);
}
static void number_v(const VNode* v, idmap_t* m) {
  idmap_add(m, v);
  for (int i = 0; i < v->n_children; ++i) number_q(v->children[i], m);
}
int64_t oracle_pomdp_tree_dump(const pomdp_tree* t, float* out, uint64_t cap_nodes) {
  idmap_t m = {NULL, 0, 0};
  number_v(t->root, &m);
  /* pre-order: entry i of the map is node i; V and Q alternate by level, so
   * the kind is recovered by walking again in the same order */
  int64_t n = m.n;
  if (out) {
    /* iterative re-walk with an explicit stack of (node, kind) */
    const void** stack = (const void**)malloc((size_t)(n + 1) * sizeof(void*));
    uint8_t* kind = (uint8_t*)malloc((size_t)n + 1);
    int sp = 0;
    uint64_t row = 0;
    stack[sp] = t->root; kind[sp] = 0; ++sp;
    while (sp > 0 && row < cap_nodes) {
      --sp;
      float* o = out + 9 * row++;
      if (kind[sp] == 0) {
        const VNode* v = (const VNode*)stack[sp];
        o[0] = 0.0f; o[1] = (float)v->observation; o[2] = v->weight;
        o[3] = v->upper_bound; o[4] = v->lower_bound; o[5] = v->heuristic;
        o[6] = (float)v->depth; o[7] = (float)v->n_children;
        o[8] = idmap_find(&m, v->vnode_to_expand);
        for (int i = v->n_children - 1; i >= 0; --i) { stack[sp] = v->children[i]; kind[sp] = 1; ++sp; }
      } else {
        const QNode* q = (const QNode*)stack[sp];
        o[0] = 1.0f; o[1] = (float)q->action; o[2] = q->reward;
        o[3] = q->upper_bound; o[4] = q->lower_bound; o[5] = q->heuristic;
        o[6] = (float)q->depth; o[7] = (float)q->n_children;
        o[8] = idmap_find(&m, q->vnode_to_expand);
        for (int i = q->n_children - 1; i >= 0; --i) { stack[sp] = q->children[i]; kind[sp] = 0; ++sp; }
      }
    }
    free(stack); free(kind);
  }
  free((void*)m.key);
  return n;
}

/* ---- helpers for fixtures (not in the reference) ------------------------ */
/* Value of the blind policy "always action a": V <- R(:,a) + gamma * P_a V,
 * `sweeps` Jacobi iterations from V = R(:,a)/(1-gamma) lower estimate 0.
 * Any such vector is a valid lower-bound alpha vector; tests and bench use
 * them as stand-ins for the PBVI alpha set (the PBVI solver is a "next" row). */
void oracle_pomdp_blind_policy(uint32_t height, uint32_t width, float gamma,
                               const float* trans_prob,
                               const float* stage_reward, uint8_t a,
                               int sweeps, float* out) {
  uint64_t hw = (uint64_t)height * width;
  float* v0 = (float*)calloc(hw, sizeof(float));
  float* v1 = (float*)calloc(hw, sizeof(float));
  for (int it = 0; it < sweeps; ++it) {
    for (int64_t y = 0; y < (int64_t)height; ++y)
      for (int64_t x = 0; x < (int64_t)width; ++x) {
        int64_t idx = y * width + x;
        float acc = 0.0f;
        for (int k = 0; k < 9; ++k) {
          int64_t nx = x + k % 3 - 1, ny = y + k / 3 - 1;
          if (nx < 0 || nx >= (int64_t)width || ny < 0 || ny >= (int64_t)height) continue;
          acc += trans_prob[idx * 81 + a * 9 + k] * v0[ny * width + nx];
        }
        v1[idx] = stage_reward[idx * 9 + a] + gamma * acc;
      }
    float* t = v0; v0 = v1; v1 = t;
  }
  memcpy(out, v0, hw * sizeof(float));
  free(v0); free(v1);
}

/* libc rand() replica self-check helper: n values from seed. */
void oracle_glibc_rand_fill(uint32_t seed, uint32_t n, uint32_t* out) {
  glibc_rand_t g;
  oracle_glibc_srand(&g, seed);
  for (uint32_t i = 0; i < n; ++i) out[i] = oracle_glibc_rand(&g);
}
