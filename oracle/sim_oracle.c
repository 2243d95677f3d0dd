/*
 * oracle/sim_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the histogram Bayes filter of the reference's
 * dummy_simulator, citing
 * /root/reference/dummy_simulator/src/dummy_simulator.cpp:
 *   440-522  DummySimulator::transitionProbability
 *   671-718  DummySimulator::updateBelief(const uint8_t& u)      prediction
 *   720-773  DummySimulator::updateBelief(const vector<uint8_t>&) correction
 * Host arithmetic of the reference: x86-64 g++ without FMA, so every product
 * and every sum is rounded separately; built here with -ffp-contract=off.
 *
 * Parity pin: oracle/_ref/libpp2d_ref_sim.so is those three methods cut out of
 * the reference source by line range (oracle/Makefile: ref_sim) and compiled
 * unmodified inside a stand-in class that only declares the members they use;
 * it needs no GPU, so tests/golden/sim_*.npz were generated in the build
 * container (tests/golden/make_golden.py sim) and tests/test_sim_cpu.py checks
 * this file against them bit for bit.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* sim:440-522.  Unlike the planners' tables there is no "trapped" override:
 * an occupied centre cell is treated like any other. */
static void transition_probability(int32_t w, int32_t h, const uint8_t* grid,
                                   int32_t x, int32_t y, uint8_t u, float* tp) {
  for (int i = 0; i < 9; ++i) tp[i] = 0.0f;
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
    default: break;
  }
  int i = 0;
  for (int oy = -1; oy < 2; ++oy)
    for (int ox = -1; ox < 2; ++ox, ++i) {
      const int32_t px = x + ox, py = y + oy;
      if (px < 0 || px >= w || py < 0 || py >= h) {       /* sim:507-512 */
        tp[4] += tp[i];
        tp[i] = 0.0f;
        continue;
      }
      if (grid[py * w + px] > 0 && i != 4) {              /* sim:513-517 */
        tp[4] += tp[i];
        tp[i] = 0.0f;
      }
    }
}

/* sim:671-718: scatter-form prediction, sequential sum, division. */
void oracle_sim_update_action(int32_t h, int32_t w, const uint8_t* grid,
                              float* belief, uint8_t u) {
  const int n = h * w;
  float* nb = (float*)calloc((size_t)n, sizeof(float));
  static const int8_t ox[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};
  static const int8_t oy[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1};
  for (int y = 0, idx = 0; y < h; ++y)
    for (int x = 0; x < w; ++x, ++idx) {
      if (belief[idx] == 0.0f) continue;
      float tp[9];
      transition_probability(w, h, grid, x, y, u, tp);
      for (int i = 0; i < 9; ++i) {
        const int px = x + ox[i], py = y + oy[i];
        if (px < 0 || px >= w || py < 0 || py >= h) continue;
        nb[py * w + px] += belief[idx] * tp[i];
      }
    }
  float sum = 0.0f;
  for (int i = 0; i < n; ++i) sum += nb[i];
  for (int i = 0; i < n; ++i) nb[i] /= sum;
  memcpy(belief, nb, sizeof(float) * (size_t)n);
  free(nb);
}

/* sim:720-773: likelihood of the four cell measurements (up, left, right,
 * down; out of map reads as occupied) times the prior, normalised. */
void oracle_sim_update_measurement(int32_t h, int32_t w, const uint8_t* grid,
                                   float* belief, const uint8_t* meas) {
  const int n = h * w;
  float* nb = (float*)calloc((size_t)n, sizeof(float));
  float sum = 0.0f;
  static const int8_t ox[4] = {0, -1, 1, 0};
  static const int8_t oy[4] = {-1, 0, 0, 1};
  for (int y = 0, idx = 0; y < h; ++y)
    for (int x = 0; x < w; ++x, ++idx) {
      if (belief[idx] == 0.0f) { nb[idx] = 0.0f; continue; }
      float l = 1.0f;
      for (int i = 0; i < 4; ++i) {
        const int32_t mx = x + ox[i], my = y + oy[i];
        uint8_t m;
        if (mx < 0 || mx >= w || my < 0 || my >= h) m = 1;
        else m = grid[my * w + mx];
        l *= (m == meas[i]) ? 0.98f : 0.02f;
      }
      nb[idx] = l * belief[idx];
      sum += nb[idx];
    }
  for (int i = 0; i < n; ++i) nb[i] /= sum;
  memcpy(belief, nb, sizeof(float) * (size_t)n);
  free(nb);
}
