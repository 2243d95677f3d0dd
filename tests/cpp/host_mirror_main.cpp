// Test driver for the C++ host mirror (include/pp2d/planners.hpp).
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <string>

#include "pp2d/planners.hpp"

using namespace path_planning_2d;

static uint64_t fnv(const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[1];
  if (mode == "maps") {
    uint32_t w, h;
    std::vector<uint8_t> grid;
    if (!pp2d::load_occupancy(argv[2], w, h, grid)) { std::printf("FAIL\n"); return 1; }
    size_t occ = 0;
    for (uint8_t v : grid) occ += v;
    std::printf("%u %u %zu %016" PRIx64 "\n", w, h, occ, fnv(grid.data(), grid.size()));
    return 0;
  }
  if (mode == "mdp" && argc >= 8) {
    Params p{{"map_path", argv[2]}, {"goal_x", argv[3]}, {"goal_y", argv[4]},
             {"discount_factor", argv[5]}, {"map_resolution", "0.2"}};
    // optional: "gpus=N" (devices 0..N-1) or "devices=0,0,1" (one shard per entry)
    for (int i = 8; i < argc; ++i) {
      const std::string a = argv[i];
      if (a.rfind("gpus=", 0) == 0) p["num_gpus"] = a.substr(5);
      if (a.rfind("devices=", 0) == 0) p["gpu_devices"] = a.substr(8);
    }
    MdpPathPlanning2d planner(p);
    if (!planner.initialize()) { std::printf("INIT_FAILED\n"); return 0; }
    auto path = planner.waypoints((uint32_t)atoi(argv[6]), (uint32_t)atoi(argv[7]));
    Belief b;
    b.belief.assign(planner.optimal_action.size(), 0.0f);
    b.belief[(size_t)atoi(argv[7]) * planner.width() + atoi(argv[6])] = 1.0f;
    std::printf("RESULT sweeps=%d cost=%016" PRIx64 " action=%016" PRIx64
                " path_len=%zu path=%016" PRIx64 " cb=%u\n",
                planner.total_iterations,
                fnv(planner.optimal_cost.data(), planner.optimal_cost.size() * 4),
                fnv(planner.optimal_action.data(), planner.optimal_action.size()),
                path.size(), fnv(path.data(), path.size() * 4), planner.beliefCallback(b));
    return 0;
  }
  if (mode == "pomdp" && argc >= 10) {
    Params p{{"map_path", argv[2]}, {"goal_x", argv[3]}, {"goal_y", argv[4]},
             {"discount_factor", argv[5]}, {"map_resolution", "0.2"},
             {"read_data_from_file", "true"}, {"data_dir", argv[6]},
             {"belief_set_size", argv[8]}, {"max_search_tree_depth", "50"},
             {"max_online_iteration", argv[9]},
             {"data_format", argc >= 11 ? argv[10] : "text"}};
    PomdpPathPlanning2d planner(p);
    if (!planner.initialize()) { std::printf("INIT_FAILED\n"); return 0; }
    {
      // what the planner now plans with: tables as uploaded, alphas as loaded
      const size_t n = (size_t)planner.height() * planner.width();
      std::vector<float> tp(n * 81), mp(n * 16), sr(n * 9);
      PP2D_CHECK(pp2d_pomdp_model_tables(planner.handle(), tp.data(), mp.data(), sr.data()));
      std::printf("TABLES tp=%016" PRIx64 " mp=%016" PRIx64 " sr=%016" PRIx64 " fib=%016" PRIx64
                  " pbvi=%016" PRIx64 " acts=%016" PRIx64 "\n",
                  fnv(tp.data(), tp.size() * 4), fnv(mp.data(), mp.size() * 4),
                  fnv(sr.data(), sr.size() * 4),
                  fnv(planner.fibAlphas().data(), planner.fibAlphas().size() * 4),
                  fnv(planner.pbviAlphas().data(), planner.pbviAlphas().size() * 4),
                  fnv(planner.pbviActions().data(), planner.pbviActions().size()));
    }
    Belief b;
    std::vector<uint8_t> raw;
    if (!pp2d::read_file(argv[7], raw)) return 1;
    b.belief.resize(raw.size() / 4);
    memcpy(b.belief.data(), raw.data(), b.belief.size() * 4);
    uint8_t a0 = planner.beliefCallback(b);
    uint32_t r0;
    memcpy(&r0, &planner.last_reward, 4);
    // second callback: the robot executed a0 and measured 0,0,0,0
    b.action = a0;
    uint8_t a1 = planner.beliefCallback(b);
    uint32_t r1;
    memcpy(&r1, &planner.last_reward, 4);
    std::printf("RESULT a0=%u r0=%08x a1=%u r1=%08x\n", a0, r0, a1, r1);
    return 0;
  }
  if (mode == "pomdp_solve" && argc >= 9) {
    // read_data_from_file = false: FIB + PBVI solved on the GPU at start-up
    // (src/pomdp/path_planning_2d.cu:109-125), then save_data.
    Params p{{"map_path", argv[2]}, {"goal_x", argv[3]}, {"goal_y", argv[4]},
             {"discount_factor", argv[5]}, {"map_resolution", "0.2"},
             {"read_data_from_file", "false"}, {"belief_set_size", argv[6]},
             {"max_search_tree_depth", "50"}, {"max_online_iteration", argv[7]},
             {"data_format", argc >= 10 ? argv[9] : "text"}};
    PomdpPathPlanning2d planner(p);
    if (!planner.initialize()) { std::printf("INIT_FAILED\n"); return 0; }
    if (!planner.saveDataCallback(argv[8])) { std::printf("SAVE_FAILED\n"); return 0; }
    Belief b;
    b.belief = planner.initial_belief;
    const uint8_t a0 = planner.beliefCallback(b);
    std::printf("RESULT fib_sweeps=%u fib=%016" PRIx64 " pbvi=%016" PRIx64 " acts=%016" PRIx64
                " a0=%u\n", planner.fib_sweeps,
                fnv(planner.fibAlphas().data(), planner.fibAlphas().size() * 4),
                fnv(planner.pbviAlphas().data(), planner.pbviAlphas().size() * 4),
                fnv(planner.pbviActions().data(), planner.pbviActions().size()), a0);
    return 0;
  }
  return 2;
}
