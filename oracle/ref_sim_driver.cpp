/*
 * oracle/ref_sim_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The Bayes filter of the reference's dummy_simulator, compiled from the
 * reference's own source lines.  dummy_simulator.cpp as a whole needs ROS and
 * OpenCV (absent from this image); the three methods of the filter
 *   DummySimulator::transitionProbability   dummy_simulator.cpp:440-522
 *   DummySimulator::updateBelief(u)         dummy_simulator.cpp:671-718
 *   DummySimulator::updateBelief(meas)      dummy_simulator.cpp:720-773
 * use only four members of the class, so oracle/Makefile (target ref_sim) cuts
 * exactly those line ranges out of /root/reference into
 * oracle/_ref/ref_sim_methods.inc (never committed) and this file compiles
 * them unmodified inside a stand-in class that declares those members and
 * nothing else.  No arithmetic is restated here.  CPU only.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace dummy_simulator {
class DummySimulator {                 // dummy_simulator.h:69-82, 102-110
 public:
  void transitionProbability(const int32_t& x, const int32_t& y, const uint8_t& u,
                             std::vector<float>& trans_prob_naive,
                             std::vector<float>& trans_prob);
  void updateBelief(const uint8_t& u);
  void updateBelief(const std::vector<uint8_t>& meas);
  int32_t map_width, map_height;
  uint8_t* grid_map;
  float* belief;
};
#include "ref_sim_methods.inc"
}  // namespace dummy_simulator

extern "C" {
int ref_sim_update_action(int32_t h, int32_t w, const uint8_t* grid, float* belief, uint8_t u) {
  dummy_simulator::DummySimulator s;
  s.map_width = w; s.map_height = h;
  s.grid_map = const_cast<uint8_t*>(grid);
  s.belief = belief;
  s.updateBelief(u);
  return 0;
}
int ref_sim_update_measurement(int32_t h, int32_t w, const uint8_t* grid, float* belief,
                               const uint8_t* meas) {
  dummy_simulator::DummySimulator s;
  s.map_width = w; s.map_height = h;
  s.grid_map = const_cast<uint8_t*>(grid);
  s.belief = belief;
  s.updateBelief(std::vector<uint8_t>(meas, meas + 4));
  return 0;
}
}
