// stand-in for <dummy_simulator/Belief.h> (TEST INFRASTRUCTURE)
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace dummy_simulator {
struct Belief {
  uint8_t action;
  std::vector<uint8_t> measurement;
  std::vector<float> belief;
};
typedef std::shared_ptr<const Belief> BeliefConstPtr;
}  // namespace dummy_simulator
