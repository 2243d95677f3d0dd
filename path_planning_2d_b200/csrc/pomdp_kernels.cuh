// pomdp_kernels.cuh -- sm_100a kernels of the QV-Tree half (SURVEY.md rows
// B1-B7).  Reference files, relative to /root/reference/path_planning_2d/:
//   model_gen = src/pomdp/model_generation_cuda.cu
//   pbvi      = src/pomdp/point_based_value_iteration_cuda.cu
//   fib       = src/pomdp/fast_informed_bound_cuda.cu
//   tree      = src/pomdp/search_tree_cuda.cu
//
// Layout: every belief of a batch is a COLUMN of one matrix
//   bel[s * cap + slot]      s = cell (y*W+x), slot = belief id, cap % 32 == 0
// so that "one thread per belief, sequential over the cells" kernels are
// coalesced.  That shape is what makes bit-exact parity possible: the
// reference normalises, prefix-sums and takes inner products with sequential
// single-accumulator float loops on the host (tree:226-229, 326-328,
// fib:289-292, pbvi:689-694).  Here each of those loops is run by one thread
// in the same order with the same roundings (separate multiply and add, IEEE
// division, subnormals kept), and the parallelism comes from the thousands of
// beliefs a batch of queries holds.  The device-side arithmetic of the
// reference (--use_fast_math: FFMA contraction, flush-to-zero) is reproduced
// with explicit .ftz PTX.
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include "async_copy.cuh"
#include "fp_exact.cuh"
#include "pomdp_host.h"

namespace pp2d {

// ---------------------------------------------------------------- B1 -------
// model_gen:161-347: per cell trans_prob[9][9] (blocked mass moved to "stay"
// in ascending slot order, trapped override AFTER the naive copy),
// meas_prob[16] and stage_reward[9] (-1 free / -2 occupied over the naive
// probabilities; stay = -2, 0 at the goal).  Same table layout as the
// reference so that save_data / the FIB and PBVI solvers can consume them.
__global__ void pomdp_model_kernel(int H, int W, int gx, int gy,
                                   const uint8_t* __restrict__ map,
                                   float* __restrict__ trans_prob,
                                   float* __restrict__ meas_prob,
                                   float* __restrict__ stage_reward) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t idx = (size_t)y * W + x;
  uint8_t m[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int nx = x + i % 3 - 1, ny = y + i / 3 - 1;
    m[i] = (nx < 0 || nx >= W || ny < 0 || ny >= H) ? 1
           : (map[(size_t)ny * W + nx] == 1 ? 1 : 0);
  }
  // naive probabilities of action u: 0.7 at slot u, 0.1 at the two side
  // slots and at the centre (model_gen:175-211).
  const int side[9][2] = {{1, 3}, {0, 2}, {1, 5}, {0, 6}, {4, 4},
                          {2, 8}, {3, 7}, {6, 8}, {5, 7}};
  for (int u = 0; u < 9; ++u) {
    float tp[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) tp[i] = 0.0f;
    if (u == 4) {
      tp[4] = 1.0f;
    } else {
      tp[u] = 0.7f; tp[side[u][0]] = 0.1f; tp[side[u][1]] = 0.1f; tp[4] = 0.1f;
    }
    // stage reward over the naive probabilities (model_gen:287-290): the
    // reward is -1 / -2, so the product is exact and the chain is the
    // reference's FFMA chain.
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) r = fma_ftz(m[i] ? -2.0f : -1.0f, tp[i], r);
    if (u == 4) r = (x != gx || y != gy) ? -2.0f : 0.0f;
    stage_reward[idx * 9 + u] = r;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (m[i] && i != 4) { tp[4] = add_ftz(tp[4], tp[i]); tp[i] = 0.0f; }
    }
    if (m[4]) {
#pragma unroll
      for (int i = 0; i < 9; ++i) tp[i] = 0.0f;
      tp[4] = 1.0f;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) trans_prob[idx * 81 + u * 9 + i] = tp[i];
  }
  // model_gen:235-263: 0.98 / 0.02 are double literals rounded to float.
  const float hit = (float)0.98, miss = (float)0.02;
  const uint8_t mm[4] = {m[1], m[3], m[5], m[7]};
  for (int z = 0; z < 16; ++z) {
    const float l0 = ((z >> 0) & 1) == mm[0] ? hit : miss;
    const float l1 = ((z >> 1) & 1) == mm[1] ? hit : miss;
    const float l2 = ((z >> 2) & 1) == mm[2] ? hit : miss;
    const float l3 = ((z >> 3) & 1) == mm[3] ? hit : miss;
    meas_prob[idx * 16 + z] = mul_ftz(mul_ftz(mul_ftz(l0, l1), l2), l3);
  }
}

// Cells no probability mass can ENTER: dead[s'] = 1 when P(s, u, s') == 0 for
// every action u and every neighbour s != s' inside the map (the only inflow
// of such a cell is from itself).  For the generated model these are the
// occupied cells (blocked mass is shifted to "stay", model_gen:213-233) plus
// free cells walled in on all sides; the mask is derived from the TABLES, so
// it also holds for tables loaded from a checkpoint.  A belief that is +0 on
// the dead cells keeps exact +0 there through every Bayes update
// (sum of +0 products, times the likelihood, divided by the sum), which is
// what lets the sequential inner products below skip those cells bit-exactly:
// acc + (+-0) == acc for an accumulator that starts at +0 (it can never become
// -0 under round-to-nearest).
__global__ void pomdp_dead_cells_kernel(int H, int W, const float* __restrict__ trans_prob,
                                        uint8_t* __restrict__ dead) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= H * W) return;
  const int x = cell % W, y = cell / W;
  bool inflow = false;
  for (int s = 0; s < 9; ++s) {
    if (s == 4) continue;
    const int sx = x + s % 3 - 1, sy = y + s / 3 - 1;
    if (sx < 0 || sx >= W || sy < 0 || sy >= H) continue;
    const size_t sidx = (size_t)sy * W + sx;
    for (int u = 0; u < 9; ++u)
      inflow = inflow || (trans_prob[81 * sidx + 9 * u + (8 - s)] != 0.0f);
  }
  dead[cell] = inflow ? 0 : 1;
}

// ---------------------------------------------------------------- B2 -------
// pbvi:88-133 cudaBayesBeliefUpdate, batched: child c is made from belief
// column src[c] with action act[c] and observation obs[c] and written,
// un-normalised, to column dst[c].  threadIdx.x runs over children so that
// siblings (same source column) read the same addresses.
// (BayesItem is declared in pomdp_host.h)

__global__ void __launch_bounds__(256)
pomdp_bayes_kernel(int H, int W, int cap, const float* __restrict__ trans_prob,
                   const float* __restrict__ meas_prob,
                   const BayesItem* __restrict__ items, int n_items,
                   const float* bel_in, float* bel_out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int cell = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (c >= n_items || cell >= H * W) return;
  const BayesItem it = items[c];
  const int x = cell % W, y = cell / W;
  // Arithmetic of the reference kernel as nvcc 12.9 compiles it for sm_100a
  // with --use_fast_math: FFMA.FTZ chain over slots 0..7, the last slot as a
  // rounded FMUL.FTZ product added with FADD.FTZ, then FMUL.FTZ by L.
  float p = 0.0f;
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const int sx = x + s % 3 - 1, sy = y + s / 3 - 1;
    if (sx < 0 || sx >= W || sy < 0 || sy >= H) continue;
    const size_t sidx = (size_t)sy * W + sx;
    const float tp = __ldg(trans_prob + 81 * sidx + 9 * it.act + (8 - s));
    const float b = bel_in[sidx * cap + it.src];
    if (s < 8) p = fma_ftz(tp, b, p);
    else p = add_ftz(p, mul_ftz(tp, b));
  }
  p = mul_ftz(p, __ldg(meas_prob + 16 * (size_t)cell + it.obs));
  bel_out[(size_t)cell * cap + it.dst] = p;
}

// The children of one Q node (same parent belief, same action) differ only in
// the observation: the predicted belief sum_s P(s,u,s') b(s) is computed once
// per (Q node, cell).  group g: children items[first[g] .. first[g+1]) (all
// with the src / act of the group's first item); threadIdx.x runs over groups,
// so the 9 Q nodes of one expanded node read the same belief addresses.
// Bayes update, column sum and division of a round without ever storing the
// un-normalised children: the prediction of a Q node is written
// once to pred[cell * ngp + g]; the sequential sum of a child and its
// normalised belief are both formed from pred * L on the fly.  Per element the
// operations are those of pomdp_bayes_kernel + pomdp_colsum_kernel +
// pomdp_scale_kernel (FMUL.FTZ by the likelihood, sequential rounded adds,
// IEEE division), so the bits are the same; the traffic drops from
// 4 x |children| to 2 x |Q nodes| + 1 x |children| belief-sized passes.
__global__ void __launch_bounds__(256)
pomdp_predict_kernel(int H, int W, int K, const int* __restrict__ kidx, int cap, int ngp,
                     const float* __restrict__ trans_prob,
                     const BayesItem* __restrict__ items, const int* __restrict__ first,
                     int n_groups, const float* __restrict__ bel, float* __restrict__ pred) {
  // Only the K target cells listed in kidx are predicted: on the others (cells
  // no mass can enter, belief +0) the prediction is +0 and nobody reads it --
  // pomdp_child_sum_kernel walks the same list and pomdp_child_write_kernel
  // writes those cells without it.  kidx = identity: every cell.
  const int g = blockIdx.x * 32 + (threadIdx.x & 31);
  const int kc = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (g >= n_groups || kc >= K) return;
  const int cell = __ldg(kidx + kc);
  const BayesItem it = items[first[g]];
  const int x = cell % W, y = cell / W;
  float p = 0.0f;
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const int sx = x + s % 3 - 1, sy = y + s / 3 - 1;
    if (sx < 0 || sx >= W || sy < 0 || sy >= H) continue;
    const size_t sidx = (size_t)sy * W + sx;
    const float tp = __ldg(trans_prob + 81 * sidx + 9 * it.act + (8 - s));
    const float b = bel[sidx * cap + it.src];
    if (s < 8) p = fma_ftz(tp, b, p);
    else p = add_ftz(p, mul_ftz(tp, b));
  }
  pred[(size_t)cell * ngp + g] = p;
}

// sums[k] = accumulate over cells of pred * L(., z_k), in cell order.  Only
// the K cells listed in kidx (ascending) are visited: the prediction is +0 on
// the others and sum + (+0 * L) == sum (see pomdp_dead_cells_kernel); with
// kidx = identity this is the plain loop over all HW cells.
// PP2D_CHILD_SUM_ORDERED = 0 rebuilds the plain-load version of round 1.
#ifndef PP2D_CHILD_SUM_ORDERED
#define PP2D_CHILD_SUM_ORDERED 1
#endif
#ifndef PP2D_CS_BATCH
#define PP2D_CS_BATCH 32
#endif
#if PP2D_CHILD_SUM_ORDERED
// Register double buffer with ordered (volatile) loads: the operands of the
// NEXT kCsBatch cells are requested before the current ones are added, so
// 2 * kCsBatch loads per thread are in flight during every batch.  (Plain
// loads get interleaved with the adds by ptxas -- 40 registers, ~8 loads in
// flight, long-scoreboard stall 22 per issue, 289 us per launch; this version
// 132 us.  A cp.async ring through shared memory was no faster than the plain
// version: 297 us.  Batch 8 / 16 / 32 cells: 217 / 132 / 86 us.)
constexpr int kCsBatch = PP2D_CS_BATCH;
__device__ __forceinline__ float ldg_f32_ordered(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__global__ void __launch_bounds__(128)
pomdp_child_sum_kernel(int K, const int* __restrict__ kidx, int ngp,
                       const float* __restrict__ meas_prob,
                       const BayesItem* __restrict__ items, const int* __restrict__ kgroup,
                       int n, const float* __restrict__ pred, float* __restrict__ sums) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float* pc = pred + kgroup[k];
  const float* L = meas_prob + items[k].obs;
  float v[2][kCsBatch], l[2][kCsBatch];
  auto request = [&](int buf, int c0) {
#pragma unroll
    for (int j = 0; j < kCsBatch; ++j) {
      const int s = __ldg(kidx + min(c0 + j, K - 1));        // clamped cells are not added
      v[buf][j] = ldg_f32_ordered(pc + (size_t)s * ngp);
      l[buf][j] = ldg_f32_ordered(L + (size_t)s * 16);
    }
  };
  float sum = 0.0f;
  request(0, 0);
  for (int c0 = 0; c0 < K; c0 += 2 * kCsBatch) {
    request(1, c0 + kCsBatch);
#pragma unroll
    for (int j = 0; j < kCsBatch; ++j)
      if (c0 + j < K) sum = __fadd_rn(sum, mul_ftz(v[0][j], l[0][j]));
    request(0, c0 + 2 * kCsBatch);
#pragma unroll
    for (int j = 0; j < kCsBatch; ++j)
      if (c0 + kCsBatch + j < K) sum = __fadd_rn(sum, mul_ftz(v[1][j], l[1][j]));
  }
  sums[k] = sum;
}
#else
__global__ void __launch_bounds__(128)
pomdp_child_sum_kernel(int K, const int* __restrict__ kidx, int ngp,
                       const float* __restrict__ meas_prob,
                       const BayesItem* __restrict__ items, const int* __restrict__ kgroup,
                       int n, const float* __restrict__ pred, float* __restrict__ sums) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float* pc = pred + kgroup[k];
  const float* L = meas_prob + items[k].obs;
  float sum = 0.0f;
  int c = 0;
  for (; c + 32 <= K; c += 32) {
    float v[32], l[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int s = __ldg(kidx + c + j);
      v[j] = pc[(size_t)s * ngp];
      l[j] = __ldg(L + (size_t)s * 16);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) sum = __fadd_rn(sum, mul_ftz(v[j], l[j]));
  }
  for (; c < K; ++c) {
    const int s = __ldg(kidx + c);
    sum = __fadd_rn(sum, mul_ftz(pc[(size_t)s * ngp], __ldg(L + (size_t)s * 16)));
  }
  sums[k] = sum;
}

#endif
// child belief = (pred * L) / sum, written once into its pool column.  dead
// (may be NULL = no cell is skipped): cells whose prediction is +0 by
// construction and was not computed: (+0 * L) / sum.
__global__ void __launch_bounds__(256)
pomdp_child_write_kernel(int HW, int cap, int ngp, const float* __restrict__ meas_prob,
                         const uint8_t* __restrict__ dead,
                         const BayesItem* __restrict__ items, const int* __restrict__ kgroup,
                         int n, const float* __restrict__ pred, const float* __restrict__ sums,
                         float* __restrict__ bel) {
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int cell = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (k >= n || cell >= HW) return;
  const BayesItem it = items[k];
  const float p = (dead != nullptr && dead[cell]) ? 0.0f : pred[(size_t)cell * ngp + kgroup[k]];
  const float v = mul_ftz(p, __ldg(meas_prob + 16 * (size_t)cell + it.obs));
  bel[(size_t)cell * cap + it.dst] = __fdiv_rn(v, sums[k]);
}

// ---------------------------------------------------------------- B3 -------
// tree:226-229: sum = accumulate(b, 0.0f) sequentially, then b /= sum (IEEE
// division).  One thread per belief column.
// The sum is a serial chain of float adds by construction; the loads are
// issued 32 at a time so that memory latency overlaps.
__global__ void __launch_bounds__(128)
pomdp_colsum_kernel(int HW, int cap, const int* __restrict__ slots, int n,
                    const float* __restrict__ bel, float* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* col = bel + slots[i];
  float sum = 0.0f;
  int s = 0;
  for (; s + 32 <= HW; s += 32) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = col[(size_t)(s + j) * cap];
#pragma unroll
    for (int j = 0; j < 32; ++j) sum = __fadd_rn(sum, v[j]);
  }
  for (; s < HW; ++s) sum = __fadd_rn(sum, col[(size_t)s * cap]);
  sums[i] = sum;
}

// b /= sum for every cell of every listed column (tree:228-229), IEEE division.
__global__ void __launch_bounds__(256)
pomdp_scale_kernel(int HW, int cap, const int* __restrict__ slots, int n,
                   const float* __restrict__ sums, float* __restrict__ bel) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int s = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (i >= n || s >= HW) return;
  const size_t q = (size_t)s * cap + slots[i];
  bel[q] = __fdiv_rn(bel[q], sums[i]);
}

// ---------------------------------------------------------------- B7 -------
// tree:326-328: distribution = partial_sum(belief) (sequential float adds).
// One WARP per expanded V node: the lanes fetch 32 consecutive cells (their
// addresses are a pool row apart, so all 32 loads are in flight together),
// then every lane replays the 32 adds in order from registers (shuffles) and
// keeps the partial sum of its own cell.  prefix[i * HW + s].
__global__ void __launch_bounds__(128)
pomdp_prefix_kernel(int HW, int cap, const int* __restrict__ slots, int n,
                    const float* __restrict__ bel, float* __restrict__ prefix) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* col = bel + slots[i];
  float* out = prefix + (size_t)i * HW;
  float acc = 0.0f;
  float nxt = lane < HW ? col[(size_t)lane * cap] : 0.0f;
  for (int s0 = 0; s0 < HW; s0 += 32) {
    const float v = nxt;
    const int sn = s0 + 32 + lane;
    nxt = sn < HW ? col[(size_t)sn * cap] : 0.0f;      // next chunk while this one is summed
    float mine = 0.0f;
    const int cnt = min(32, HW - s0);
    for (int j = 0; j < cnt; ++j) {
      acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, v, j));
      if (j == lane) mine = acc;
    }
    if (lane < cnt) out[s0 + lane] = mine;
  }
}

// The 2*n uniforms cudaForwardSampling consumes (tree:84-92, 117, 134):
// XORWOW, curand_init(1234, idx, 0), two curand_uniform draws.  The reference
// re-creates the states for every Q node, so these numbers never change.
__global__ void pomdp_uniforms_kernel(int n, float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  curandState st;
  curand_init(1234, idx, 0, &st);
  out[2 * idx] = curand_uniform(&st);
  out[2 * idx + 1] = curand_uniform(&st);
}

// tree:331-337 (state sample by inverse CDF: first prefix >= draw; the prefix
// is non-decreasing, so a binary search finds the same element) followed by
// tree:94-147 cudaForwardSampling.  One thread per (expanded node i, action
// a, sample k); draws[(i*9+a)*S + k] are the host rand() values already
// divided by RAND_MAX+1.  A draw beyond the last prefix sum is clamped to the
// last cell (the reference indexes one past the belief there).
__global__ void __launch_bounds__(128)
pomdp_sample_kernel(int H, int W, int n, int S,
                    const float* __restrict__ trans_prob,
                    const float* __restrict__ meas_prob,
                    const float* __restrict__ prefix,
                    const float* __restrict__ draws,
                    const float* __restrict__ uniforms,
                    uint8_t* __restrict__ observations) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 9 * S) return;
  const int k = t % S, a = (t / S) % 9, i = t / (9 * S);
  const int HW = H * W;
  const float r = draws[t];
  int lo = 0, hi = HW;                       // first s with prefix[s] >= r
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (prefix[(size_t)i * HW + mid] >= r) hi = mid; else lo = mid + 1;
  }
  int s1 = lo < HW ? lo : HW - 1;
  float cum = 0.0f;
  const float r2 = uniforms[2 * k];
  int s2 = 0;
  bool found = false;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const float tp = trans_prob[(size_t)s1 * 81 + a * 9 + j];
    cum = j == 0 ? tp : add_ftz(tp, cum);
    if (!found && r2 <= cum) { s2 = j; found = true; }
  }
  long long nxt = (long long)s1 + (long long)(s2 / 3 - 1) * W + (s2 % 3 - 1);
  if (nxt < 0) nxt = 0;
  if (nxt >= HW) nxt = HW - 1;
  const float r3 = uniforms[2 * k + 1];
  int z = 0;
  found = false;
  cum = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float mp = meas_prob[(size_t)nxt * 16 + j];
    cum = j == 0 ? mp : add_ftz(mp, cum);
    if (!found && r3 <= cum) { z = j; found = true; }
  }
  observations[t] = (uint8_t)z;
}

// ------------------------------------------------------------ B4 / B5 ------
// Values of many beliefs against many alpha vectors:
//   out[i][j] = sum_s bel[s][slots[i]] * alpha[s][j]     (j < ncol)
// evaluated exactly like std::inner_product(b, b+HW, alpha, 0.0f) on the
// reference's host (fib:289-292, pbvi:689-694, tree:172): one accumulator per
// (i, j), cells in ascending order, float multiply rounded, then float add
// rounded (no FMA).  alpha is [HW][ld]: the bound matrix has the columns
//   0..8 FIB, 9..9+N-1 PBVI
// (509 -> 512 = four column tiles for the reference's 500 vectors); ld is a
// multiple of the column tile and the padding is zero.
// CTA tile 128 beliefs x 128 columns, 256 threads, 8x8 accumulators each
// (4 LDS.128 per 128 math instructions: the 4x4 version was bound by the
// shared-memory pipe), K chunks of 16 cells double-buffered with cp.async.
constexpr int kEvM = 128, kEvN = 128, kEvK = 16, kEvPad = 4;

// The inner dimension runs over the K cells listed in kidx (ascending cell
// order); alpha holds the K matching rows.  K = HW with kidx = identity is the
// dense product; with the live cells only (pomdp_dead_cells_kernel) every
// belief of the launch must be +0 on the cells left out, and the result is
// bit-identical: the skipped terms are acc + (+-0).
__global__ void __launch_bounds__(256, 2)
pomdp_values_kernel(int HW, const int* __restrict__ kidx, int cap, int ld, int ncol,
                    const int* __restrict__ slots, int n,
                    const float* __restrict__ bel,
                    const float* __restrict__ alpha, float* __restrict__ out) {
  __shared__ __align__(16) float sb[2][kEvK][kEvM + kEvPad];
  __shared__ __align__(16) float sa[2][kEvK][kEvN + kEvPad];
  __shared__ int sslot[kEvM];
  const int m0 = blockIdx.x * kEvM, n0 = blockIdx.y * kEvN;
  const int tid = threadIdx.x;
  if (tid < kEvM) sslot[tid] = (m0 + tid < n) ? slots[m0 + tid] : slots[0];
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  // Tile loaders (HW = the number K of inner-dimension rows of this launch).
  // Rows past it are clamped to the last row: they are never accumulated (the
  // k loop stops at HW), only kept in bounds.  Columns past ncol read the zero
  // padding of alpha (ld is a multiple of 128).
  auto load_tiles = [&](int buf, int k0) {
#pragma unroll
    for (int e = 0; e < (kEvK * kEvM) / 256; ++e) {
      const int idx = tid + e * 256;
      const int kk = idx / kEvM, mm = idx % kEvM;
      const int s = __ldg(kidx + min(k0 + kk, HW - 1));
      cp_async<4>((uint32_t)__cvta_generic_to_shared(&sb[buf][kk][mm]),
                  bel + (size_t)s * cap + sslot[mm]);
    }
#pragma unroll
    for (int e = 0; e < (kEvK * kEvN / 4) / 256; ++e) {
      const int idx = tid + e * 256;
      const int kk = idx / (kEvN / 4), nn = (idx % (kEvN / 4)) * 4;
      const int s = min(k0 + kk, HW - 1);
      cp_async<16>((uint32_t)__cvta_generic_to_shared(&sa[buf][kk][nn]),
                   alpha + (size_t)s * ld + n0 + nn);
    }
    cp_async_commit();
  };

  const int nchunks = (HW + kEvK - 1) / kEvK;
  load_tiles(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      load_tiles(buf ^ 1, (c + 1) * kEvK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int kend = min(kEvK, HW - c * kEvK);
#pragma unroll 4
    for (int kk = 0; kk < kend; ++kk) {
      const float4 b0 = *reinterpret_cast<const float4*>(&sb[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sb[buf][kk][64 + tx * 4]);
      const float4 a0 = *reinterpret_cast<const float4*>(&sa[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sa[buf][kk][64 + ty * 4]);
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(bv[i], av[j]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
    if (m >= n) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
      if (col < ncol) out[(size_t)m * ncol + col] = acc[i][j];
    }
  }
}

// tree:168-173: reward[i][a] = inner_product(b_i, R(:,a), 0.0f) for the nodes
// being expanded: one WARP per (belief, action).  The lanes fetch and multiply
// 32 consecutive cells, then every lane replays the 32 rounded adds in order
// (the same sequential multiply-then-add chain as pomdp_values_kernel).
// stage_reward is the reference table [HW][9].
__global__ void __launch_bounds__(128)
pomdp_rewards_kernel(int K, const int* __restrict__ kidx, int cap,
                     const int* __restrict__ slots, int n,
                     const float* __restrict__ bel, const float* __restrict__ stage_reward,
                     float* __restrict__ out) {
  // cells kidx[0..K) only (ascending): the belief is +0 on the others and
  // acc + (+0 * r) == acc (see pomdp_dead_cells_kernel)
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= n * 9) return;
  const int a = w % 9, i = w / 9;
  const float* col = bel + slots[i];
  const float* r = stage_reward + a;
  float acc = 0.0f;
  int s = lane < K ? __ldg(kidx + lane) : 0;
  float nb = lane < K ? col[(size_t)s * cap] : 0.0f;
  float nr = lane < K ? __ldg(r + (size_t)s * 9) : 0.0f;
  for (int c0 = 0; c0 < K; c0 += 32) {
    const float p = __fmul_rn(nb, nr);
    const int cn = c0 + 32 + lane;
    s = cn < K ? __ldg(kidx + cn) : 0;
    nb = cn < K ? col[(size_t)s * cap] : 0.0f;
    nr = cn < K ? __ldg(r + (size_t)s * 9) : 0.0f;
    const int cnt = min(32, K - c0);
    for (int j = 0; j < cnt; ++j) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, p, j));
  }
  if (lane == 0) out[(size_t)i * 9 + a] = acc;
}

// Bounds of every evaluated belief from its row of values: first maximum over
// the FIB columns (upper, fib:294-296) and over the PBVI columns (lower,
// pbvi:696-698), as std::max_element does.  Packed per belief:
//   out[i*4 + 0] upper, [1] lower, [2] = fib index | pbvi index << 8 (as int
//   bits).
__global__ void __launch_bounds__(128)
pomdp_bounds_kernel(int n, int ncol, int n_pbvi, const float* __restrict__ vals,
                    float* __restrict__ out) {
  // one warp per belief: lane-local first maximum over a strided subset of the
  // columns, then (value, index) reduction where ties keep the smaller index
  // -- the element std::max_element returns.
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* v = vals + (size_t)i * ncol;
  float bu = 0.0f, bl = 0.0f;
  int iu = -1, il = -1;
  if (lane < 9) { bu = v[lane]; iu = lane; }
  for (int j = lane; j < n_pbvi; j += 32) {
    const float x = v[9 + j];
    if (il < 0 || bl < x) { bl = x; il = j; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ou = __shfl_xor_sync(0xffffffffu, bu, off);
    const int oiu = __shfl_xor_sync(0xffffffffu, iu, off);
    if (oiu >= 0 && (iu < 0 || bu < ou || (bu == ou && oiu < iu))) { bu = ou; iu = oiu; }
    const float ol = __shfl_xor_sync(0xffffffffu, bl, off);
    const int oil = __shfl_xor_sync(0xffffffffu, il, off);
    if (oil >= 0 && (il < 0 || bl < ol || (bl == ol && oil < il))) { bl = ol; il = oil; }
  }
  if (lane == 0) {
    float* o = out + (size_t)i * 4;
    o[0] = bu;
    o[1] = n_pbvi > 0 ? bl : 0.0f;
    o[2] = __int_as_float(iu | ((il < 0 ? 0 : il) << 8));
    o[3] = 0.0f;
  }
}

// ------------------------------------------------ FIB solver ("next" #1) ----
// fib:97-204 cudaFIBValueIteration: one sweep of the Fast Informed Bound
// backup on the 9 alpha vectors, alpha[s][a] layout.  One thread per
// (cell, action); same operation order as the reference kernel: per
// observation o, trans*meas (FMUL.FTZ), per next action an FFMA.FTZ chain over
// the 9 next states from 0, running max, FADD.FTZ into reward-to-go, and
// finally fma(gamma, reward_to_go, reward).
__global__ void __launch_bounds__(128)
pomdp_fib_kernel(int H, int W, float gamma, const float* __restrict__ trans_prob,
                 const float* __restrict__ meas_prob,
                 const float* __restrict__ stage_reward,
                 const float* __restrict__ prev, float* __restrict__ curr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W * 9) return;
  const int a = t % 9, cell = t / 9;
  const int x = cell % W, y = cell / W;
  float tp[9];
  int nidx[9];
#pragma unroll
  for (int sp = 0; sp < 9; ++sp) {
    const int nx = x + sp % 3 - 1, ny = y + sp / 3 - 1;
    nidx[sp] = (nx < 0 || nx >= W || ny < 0 || ny >= H) ? -1 : ny * W + nx;
    tp[sp] = __ldg(trans_prob + 81 * (size_t)cell + 9 * a + sp);
  }
  float rtg = 0.0f;
  for (int o = 0; o < 16; ++o) {
    float ltm[9];
#pragma unroll
    for (int sp = 0; sp < 9; ++sp)
      ltm[sp] = mul_ftz(tp[sp], nidx[sp] >= 0 ? __ldg(meas_prob + 16 * (size_t)nidx[sp] + o)
                                              : 0.0f);
    float best = -3.402823466e+38f;
    for (int ap = 0; ap < 9; ++ap) {
      float acc = 0.0f;
#pragma unroll
      for (int sp = 0; sp < 9; ++sp)
        acc = fma_ftz(ltm[sp], nidx[sp] >= 0 ? __ldg(prev + 9 * (size_t)nidx[sp] + ap) : 0.0f,
                      acc);
      if (best < acc) best = acc;
    }
    rtg = add_ftz(rtg, best);
  }
  curr[(size_t)cell * 9 + a] = fma_ftz(gamma, rtg, __ldg(stage_reward + (size_t)cell * 9 + a));
}

// max |a - b| over n floats (fib:245-251), atomicMax on the float bits.
__global__ void __launch_bounds__(256)
pomdp_maxdiff_kernel(const float* __restrict__ a, float* __restrict__ b_and_copy,
                     size_t n, unsigned int* result) {
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float v = a[i];
    m = fmaxf(m, fabsf(b_and_copy[i] - v));
    b_and_copy[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(result, __float_as_uint(m));
}

// Gather host-provided beliefs ([n][HW] row major) into belief columns.
__global__ void pomdp_scatter_kernel(int HW, int cap, const int* __restrict__ slots,
                                     int n, const float* __restrict__ rows,
                                     float* __restrict__ bel) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (s >= HW || i >= n) return;
  bel[(size_t)s * cap + slots[i]] = rows[(size_t)i * HW + s];
}

// Belief columns back to [n][HW] rows.
__global__ void pomdp_gather_kernel(int HW, int cap, const int* __restrict__ slots,
                                    int n, const float* __restrict__ bel,
                                    float* __restrict__ rows) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (s >= HW || i >= n) return;
  rows[(size_t)i * HW + s] = bel[(size_t)s * cap + slots[i]];
}

}  // namespace pp2d
