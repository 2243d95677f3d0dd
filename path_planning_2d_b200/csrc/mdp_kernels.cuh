// mdp_kernels.cuh -- sm_100a kernels of the MDP value-iteration hot path.
//
// Reference behaviour being reproduced (paths relative to
// /root/reference/path_planning_2d/):
//   src/mdp/path_planning_2d_cuda.cu:76-213  model tables (P, g) per cell
//   src/mdp/path_planning_2d_cuda.cu:215-264 one Jacobi Bellman backup
//
// Design (see DESIGN.md for the derivation):
//   * The reference's 360 B/cell tables are never materialised.  Everything a
//     backup needs is a function of the 3x3 occupancy, stored as one 16-bit
//     code per cell ("ring code": the 8 neighbours in ring order
//     s0 s1 s2 s5 s8 s7 s6 s3 s0 s1, so that the three neighbour slots every
//     action can move into are 3 consecutive bits).
//   * J of occupied cells, of the goal and of out-of-map padding is stored as
//     0.  A blocked neighbour then contributes fma(coef, 0, cost) = cost,
//     exactly what the reference's zero probability does, with no select.
//     Occupied cells follow the closed form J_n = fma(gamma, J_{n-1}, 2) and
//     are filled in on download.
//   * Per action the only occupancy-dependent numbers are the stage cost g_u
//     and the centre coefficient gamma*P_u[4]; both come from an 8 KB
//     shared-memory table indexed by 4 ring bits (two actions per 16-byte
//     row, replicated 8x so that a quarter-warp LDS.128 never bank-conflicts).
//   * Each lane owns CW consecutive columns and marches down its rows; the
//     3-row window of J lives in registers, horizontal neighbours come from
//     warp shuffles, so there is no shared-memory tile and every global load
//     is a coalesced row segment.
//   * T = 2 fuses two sweeps: J^1 of row y is produced in registers, shuffled
//     to the neighbours and consumed for J^2 of row y-1 while the LUT rows of
//     row y-1 are still in registers (register-level temporal blocking).  A
//     warp recomputes HL lanes of halo on each side.
//   * min over the 9 actions uses FMNMX3; the arg-min (POLICY) variant is
//     only run for the last sweep of a pp2d_mdp_sweeps call.
//
// Bit-exactness: each action cost is the reference's chain
//   cost = g_u; for k ascending: cost = fma(gamma*P_u[k], J[n_k], cost)
// with gamma*0.7f, gamma*0.1f and gamma*P_u[4] rounded once (FMUL) before the
// FFMA, as in the SASS of the reference kernel.  Terms
// whose probability is 0 are skipped: fma(0, J, c) == c for finite J.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "async_copy.cuh"
#include "fp_exact.cuh"

namespace pp2d {

constexpr int kPadRows = 2;   // ghost rows above and below the owned rows
constexpr int kSlackRows = 8; // extra rows at the bottom: prefetch may overrun
constexpr int kPrefetch = 2;  // CW == 1 path: rows held in registers ahead of use
#ifndef PP2D_KRING
#define PP2D_KRING 3
#endif
constexpr int kRing = PP2D_KRING;  // CW >= 2 path: cp.async ring slots per lane (1, 2, 3 or 6)
// PP2D_TMA = 1 (experiment, A/B-timed in profiles/): the row ring of the fused
// CW = 2 kernel is filled by bulk async copies (cp.async.bulk, the TMA engine)
// issued by one lane per warp and completed on a per-slot mbarrier, instead of
// per-lane cp.async.  A bulk copy needs 16-byte aligned addresses: with 10 pad
// columns a warp's 64-float row segment (strip k starts at column 60 k - 2)
// is aligned; the CW = 4 kernels (16-byte per-lane vectors) then are not, so
// such a build only runs CW <= 2.
#ifndef PP2D_TMA
#define PP2D_TMA 0
#endif
constexpr int kPadLeft = PP2D_TMA ? 10 : 8;   // zero columns left of x = 0
// Table layouts (build-time choice, both give the same bits):
//   PP2D_LUT6 = 0: 4 tables (one per action pair) x 16 rows (4 ring bits) x 8
//     lane replicas of one float4 = 8 KB; a cell needs 4 row addresses.
//   PP2D_LUT6 = 1: two ring-adjacent action pairs share 6 ring bits, so ONE row
//     address serves both: 2 super-pairs x 2 halves x 64 rows x 8 replicas =
//     32 KB, 2 row addresses per cell (each SHF + LOP3) instead of 4.
#ifndef PP2D_LUT6
#define PP2D_LUT6 0
#endif
// PP2D_UWARP = 1: the warp index is read through a shuffle from lane 0, which
// tells the compiler it is warp-uniform (unit, row counts and loop exits then
// are, too: no divergence guard in front of the shuffles of the marching loop).
#ifndef PP2D_UWARP
#define PP2D_UWARP 1
#endif
#if PP2D_LUT6
constexpr int kLutFloat4 = 2 * 2 * 64 * 8;
constexpr int kLutAlign = 8192;         // row bits 7..12 + replica bits 4..6 are OR-ed in
#else
constexpr int kLutFloat4 = 4 * 16 * 8;  // 4 action pairs x 16 rows x 8 copies
constexpr int kLutAlign = 2048;
#endif
// Tuning knobs (tools/sweep_variants.py builds alternates): how the minimum
// over the 9 actions is taken, warps per CTA and resident CTAs per SM of the
// fused kernel.
#ifndef PP2D_MINV
#define PP2D_MINV 0
#endif
// PP2D_ARGMIN_EQ = 1: the arg-min variant takes the minimum with FMNMX3 and
// finds its first index by equality (see backup()); 0: compare-and-keep chain.
// (236 against 245 instructions per marching step of the fused arg-min kernel,
// but the P2P strip-major instantiation then spills two registers: off.)
#ifndef PP2D_ARGMIN_EQ
#define PP2D_ARGMIN_EQ 0
#endif
#ifndef PP2D_WARPS
#define PP2D_WARPS 8
#endif
#ifndef PP2D_MINCTAS
#define PP2D_MINCTAS 2
#endif
constexpr int kWarpsPerCta = PP2D_WARPS;

// Code bits.
constexpr uint32_t kCodeRingMask = 0x3FFu;   // bits 0..9
constexpr uint32_t kCodeOccBit = 1u << 13;   // occupied or padding
constexpr uint32_t kCodeLiveBit = 1u << 14;  // free, in map, not the goal

struct SweepParams {
  const float* jin;      // padded plane, element (y, x) at [(y+2)*pitch + x+8]
  float* jout;
  const uint16_t* code;  // same geometry as J
  uint8_t* action;       // dense [H][W] (POLICY only)
  const float4* lut;     // kLutFloat4 / 8 rows (the kernel writes the 8 lane replicas)
  int W, H, pitch;
  int n_strips, rows_per_unit, n_units;
  int y_begin, y_end;    // rows to produce (may extend 1 row into the ghosts)
  int prefetch_rows;     // rows ahead touched with prefetch.global.L1 (>= kPrefetch)
  float gamma, ga, gb;   // gamma*1.0f, gamma*0.7f, gamma*0.1f
  // --- peer-to-peer ghost rows (P2P kernels only, see P2pParams) ---
  float* peer_up_out;    // up neighbour's jout plane at ITS ghost row H_up, col 0
  float* peer_down_out;  // down neighbour's jout plane at ITS ghost row -2, col 0
  unsigned int* flags;   // this shard's flag block (kFlag*)
  unsigned int* up_flag_remote;    // &up.flags[kFlagFromDown]
  unsigned int* down_flag_remote;  // &down.flags[kFlagFromUp]
  unsigned int iter;               // 1-based count of fused P2P launches
  unsigned int expect_top, expect_bot;  // cumulative boundary-unit counts
  unsigned int p2p_debug;               // bit0 no waits, bit1 no peer stores, bit2 no signals
  int edge_rows;                        // rows of a boundary unit processed first
  int lin_len;           // LIN kernels: units are runs of lin_len rows of the strip-major sequence
  unsigned int spin_limit;  // P2P: polls of a neighbour flag before giving up (kFlagError)
  // P2P LIN kernels: unit u covers rows [unit_lo[u], unit_lo[u+1]) of the
  // strip-major sequence.  The host balances COST, not rows: a unit that
  // touches the first / last two rows of a strip also does the hand-shake
  // (system fences wait for NVLink write acknowledgements) and gets fewer rows,
  // so that it does not finish after everybody else.  NULL: runs of lin_len.
  const int* unit_lo;
  // P2P row-block units: the same idea with two numbers: the first and last
  // row block of a strip get rows_edge < rows_per_unit rows, the blocks between
  // them rows_inner; rows_edge == 0: all blocks have rows_per_unit rows.
  int rows_edge, rows_inner;
  // P2P: index of the warp that publishes this launch's flags to the
  // neighbours (see publish_flags), -1: the last boundary unit to finish its
  // edge rows does it itself.
  int publisher_unit;
};

// Flag block of a shard (uint32 each, cudaMalloc'ed, IPC-shared).
constexpr int kFlagFromUp = 0;     // written by the up neighbour: its launch #
constexpr int kFlagFromDown = 1;   // written by the down neighbour
constexpr int kFlagCountTop = 2;   // local: finished top-boundary units
constexpr int kFlagCountBot = 3;   // local: finished bottom-boundary units
constexpr int kFlagError = 4;      // local: a wait timed out
constexpr int kFlagWords = 8;

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// One Bellman backup of one cell.  j0..j8: J of the 3x3 neighbourhood in the
// reference's slot order (row-major, slot 4 = the cell), blocked slots = 0.
// t[0]={g1,c1,g2,c2} t[1]={g5,c5,g8,c8} t[2]={g7,c7,g6,c6} t[3]={g3,c3,g0,c0}
// (g = stage cost, c = gamma*P[4] of that action), g4 = stage cost of "stay".
template <bool POLICY>
__device__ __forceinline__ float backup(float j0, float j1, float j2, float j3,
                                        float j4, float j5, float j6, float j7,
                                        float j8, const float4 (&t)[4],
                                        float g4, float gam, float ga,
                                        float gb, uint32_t& act) {
  // Slot order of the non-zero probabilities of each action
  // (path_planning_2d_cuda.cu:89-125), centre coefficient from the table.
  float c0 = fmaf(t[3].w, j4, fmaf(gb, j3, fmaf(gb, j1, fmaf(ga, j0, t[3].z))));
  float c1 = fmaf(t[0].y, j4, fmaf(gb, j2, fmaf(ga, j1, fmaf(gb, j0, t[0].x))));
  float c2 = fmaf(gb, j5, fmaf(t[0].w, j4, fmaf(ga, j2, fmaf(gb, j1, t[0].z))));
  float c3 = fmaf(gb, j6, fmaf(t[3].y, j4, fmaf(ga, j3, fmaf(gb, j0, t[3].x))));
  float c4 = fmaf(gam, j4, g4);
  float c5 = fmaf(gb, j8, fmaf(ga, j5, fmaf(t[1].y, j4, fmaf(gb, j2, t[1].x))));
  float c6 = fmaf(gb, j7, fmaf(ga, j6, fmaf(t[2].w, j4, fmaf(gb, j3, t[2].z))));
  float c7 = fmaf(gb, j8, fmaf(ga, j7, fmaf(gb, j6, fmaf(t[2].y, j4, t[2].x))));
  float c8 = fmaf(ga, j8, fmaf(gb, j7, fmaf(gb, j5, fmaf(t[1].w, j4, t[1].z))));
  if (POLICY) {
#if PP2D_ARGMIN_EQ
    // path_planning_2d_cuda.cu:244-258 keeps the first strict minimum, u
    // ascending = the lowest u whose cost equals the minimum.  The minimum is
    // taken as in the value-only kernel (costs are finite and >= +0: no NaN or
    // -0 for FMNMX to choose differently), then one compare + select per
    // action in DESCENDING order, so that the lowest index is written last: 4
    // FMNMX3 + 8 FSETP + 8 SEL with one live predicate, instead of a chain of
    // 8 x (FSETP, FSEL, SEL) that keeps 8 predicates alive.
    const float best = min3(min3(c0, c1, c2), min3(c3, c4, c5), min3(c6, c7, c8));
    uint32_t a = 8;
    a = (c7 == best) ? 7u : a;
    a = (c6 == best) ? 6u : a;
    a = (c5 == best) ? 5u : a;
    a = (c4 == best) ? 4u : a;
    a = (c3 == best) ? 3u : a;
    a = (c2 == best) ? 2u : a;
    a = (c1 == best) ? 1u : a;
    a = (c0 == best) ? 0u : a;
    act = a;
    return best;
#else
    // path_planning_2d_cuda.cu:244-258: first strict minimum, u ascending.
    float best = c0;
    uint32_t a = 0;
    if (c1 < best) { best = c1; a = 1; }
    if (c2 < best) { best = c2; a = 2; }
    if (c3 < best) { best = c3; a = 3; }
    if (c4 < best) { best = c4; a = 4; }
    if (c5 < best) { best = c5; a = 5; }
    if (c6 < best) { best = c6; a = 6; }
    if (c7 < best) { best = c7; a = 7; }
    if (c8 < best) { best = c8; a = 8; }
    act = a;
    return best;
#endif
  } else {
#if PP2D_MINV == 1
    return fminf(fminf(fminf(c0, c1), fminf(c2, c3)),
                 fminf(fminf(fminf(c4, c5), fminf(c6, c7)), c8));
#elif PP2D_MINV == 2
    // Costs are finite and >= +0, so their order as floats is their order as
    // unsigned integers: the 3-input integer minimum (VIMNMX3) gives the same
    // bits as FMNMX3.
    const uint32_t m0 = __vimin3_u32(__float_as_uint(c0), __float_as_uint(c1), __float_as_uint(c2));
    const uint32_t m1 = __vimin3_u32(__float_as_uint(c3), __float_as_uint(c4), __float_as_uint(c5));
    const uint32_t m2 = __vimin3_u32(__float_as_uint(c6), __float_as_uint(c7), __float_as_uint(c8));
    return __uint_as_float(__vimin3_u32(m0, m1, m2));
#else
    return min3(min3(c0, c1, c2), min3(c3, c4, c5), min3(c6, c7, c8));
#endif
  }
}

// Row loads are volatile asm on purpose: a plain __ldg() is sunk by the
// compiler to just before its first use, which collapses the kPrefetch-deep
// software pipeline (seen in the ncu source view: long-scoreboard stalls on
// the first use of every prefetched register).
// NC: the non-coherent (read-only) path.  It is only legal for data nobody
// writes while the kernel runs, so the P2P instantiations -- whose ghost rows
// of jin are written by the neighbour GPU during the launch, ordered by the
// flag hand-shake -- use ordinary weak loads, which the acquire of the flag
// wait does order.
template <int CW, bool NC = true>
__device__ __forceinline__ void load_own(const float* __restrict__ p,
                                         float (&o)[CW]) {
  if constexpr (!NC) {
    if constexpr (CW == 1) {
      asm volatile("ld.global.f32 %0, [%1];" : "=f"(o[0]) : "l"(p) : "memory");
    } else if constexpr (CW == 2) {
      asm volatile("ld.global.v2.f32 {%0, %1}, [%2];"
                   : "=f"(o[0]), "=f"(o[1]) : "l"(p) : "memory");
    } else {
      asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p) : "memory");
    }
  } else if constexpr (CW == 1) {
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(o[0]) : "l"(p));
  } else if constexpr (CW == 2) {
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];"
                 : "=f"(o[0]), "=f"(o[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p));
  }
}

template <int CW>
__device__ __forceinline__ void store_own(float* p, const float (&o)[CW]) {
  if constexpr (CW == 1) {
    *p = o[0];
  } else if constexpr (CW == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]);
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Codes of the lane's CW cells, packed two per 32-bit word.
template <int CW>
__device__ __forceinline__ void load_codes(const uint16_t* __restrict__ p,
                                           uint32_t (&c)[(CW + 1) / 2]) {
  if constexpr (CW == 1) {
    uint16_t v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
    c[0] = v;
  } else if constexpr (CW == 2) {
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(c[0]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];"
                 : "=r"(c[0]), "=r"(c[1]) : "l"(p));
  }
}

// own[CW] -> row[CW+2] with the left/right neighbours taken from the
// adjacent lanes (lane 0 / lane 31 get their own value: those lanes are halo).
template <int CW>
__device__ __forceinline__ void fill_row(const float (&own)[CW],
                                         float (&row)[CW + 2]) {
  row[0] = __shfl_up_sync(0xffffffffu, own[CW - 1], 1);
  row[CW + 1] = __shfl_down_sync(0xffffffffu, own[0], 1);
#pragma unroll
  for (int j = 0; j < CW; ++j) row[1 + j] = own[j];
}

// Fetch the 4 table rows of one cell.  `word` holds two 16-bit codes; HI
// selects the upper one.  lane_base = shared address of the table (8 KB
// aligned) | (lane & 7) * 16, so the row offset can be OR-ed in: one shift
// and one LOP3 per row.
template <int IMM>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "r"(addr), "n"(IMM));
  return v;
}

template <bool HI, int P>
__device__ __forceinline__ float4 lut_row(uint32_t lane_base, uint32_t word) {
  const uint32_t sh = HI ? (word >> (9 + 2 * P)) : (word << (7 - 2 * P));
  return lds128<P * 2048>((sh & 0x780u) | lane_base);
}

template <bool HI>
__device__ __forceinline__ void lut_fetch(uint32_t lane_base, uint32_t word,
                                          float4 (&t)[4], float& g4) {
#if PP2D_LUT6
  // super-pair 0: ring bits 0..5 (pairs 0 and 1), super-pair 1: ring bits 4..9
  // (pairs 2 and 3); row offset = 6 bits << 7, the two halves 8 KB apart
  const uint32_t s0 = HI ? (word >> 9) : (word << 7);
  const uint32_t a0 = (s0 & 0x1F80u) | lane_base;
  t[0] = lds128<0>(a0);
  t[1] = lds128<8192>(a0);
  const uint32_t s1 = HI ? (word >> 13) : (word << 3);
  const uint32_t a1 = (s1 & 0x1F80u) | lane_base;
  t[2] = lds128<16384>(a1);
  t[3] = lds128<24576>(a1);
#else
  t[0] = lut_row<HI, 0>(lane_base, word);
  t[1] = lut_row<HI, 1>(lane_base, word);
  t[2] = lut_row<HI, 2>(lane_base, word);
  t[3] = lut_row<HI, 3>(lane_base, word);
#endif
  // live bit (14 of the code) -> 2.0f (0x40000000); goal, occupied cells and
  // padding -> 0.0f, which makes "stay" cost exactly 0 there (their J is 0).
  g4 = __uint_as_float((HI ? word : (word << 16)) & 0x40000000u);
}

// ---- cp.async staging (CW >= 2) --------------------------------------------
// Every lane copies its own CW floats of J and CW codes of a row into its
// private slot of a per-warp shared-memory ring, kRing rows ahead, and reads
// them back itself: cp.async.wait_group is per thread, so no warp or block
// synchronisation is needed.  Unlike register loads (which ptxas sinks to
// their first use, sharing a scoreboard with younger loads) the copies are
// issued where they are written and waited for with a counting barrier
// (LDGDEPBAR / DEPBAR.LE in SASS).
template <int CW, int OFF>
__device__ __forceinline__ void lds_row(uint32_t saddr, float (&o)[CW]) {
  if constexpr (CW == 2) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];"
                 : "=f"(o[0]), "=f"(o[1]) : "r"(saddr), "n"(OFF) : "memory");
  } else {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
                 : "r"(saddr), "n"(OFF) : "memory");
  }
}
template <int CW, int OFF>
__device__ __forceinline__ void lds_codes(uint32_t saddr,
                                          uint32_t (&c)[(CW + 1) / 2]) {
  if constexpr (CW == 2) {
    asm volatile("ld.shared.u32 %0, [%1+%2];"
                 : "=r"(c[0]) : "r"(saddr), "n"(OFF) : "memory");
  } else {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];"
                 : "=r"(c[0]), "=r"(c[1]) : "r"(saddr), "n"(OFF) : "memory");
  }
}

// ---- bulk-copy (TMA) staging helpers (PP2D_TMA builds) -----------------------
// (lane 0 only, predicated: no divergent branch)
__device__ __forceinline__ void mbar_init_lane0(int lane, uint32_t bar, uint32_t count) {
  asm volatile(
      "{\n"
      ".reg .pred P0;\n"
      "setp.eq.s32 P0, %0, 0;\n"
      "@P0 mbarrier.init.shared::cta.b64 [%1], %2;\n"
      "}\n" :: "r"(lane), "r"(bar), "r"(count) : "memory");
}
// Spin until the phase with the given parity of the barrier has completed.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" :: "r"(bar), "r"(parity) : "memory");
}
// Lane 0 (predicated, no branch: the warp stays converged for the compiler):
// expect JB + CB bytes on the barrier and start the two bulk copies.
template <int JB, int CB>
__device__ __forceinline__ void tma_fill(int lane, uint32_t bar, uint32_t dst_j, const void* src_j,
                                         uint32_t dst_c, const void* src_c) {
  asm volatile(
      "{\n"
      ".reg .pred P0;\n"
      "setp.eq.s32 P0, %0, 0;\n"
      "@P0 mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %6;\n"
      "@P0 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %7, [%1];\n"
      "@P0 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%4], [%5], %8, [%1];\n"
      "}\n"
      :: "r"(lane), "r"(bar), "r"(dst_j), "l"(src_j), "r"(dst_c), "l"(src_c), "n"(JB + CB),
         "n"(JB), "n"(CB)
      : "memory");
}

template <int T, int CW>
struct StripGeom {
  static constexpr int HL = (T + CW - 1) / CW;          // halo lanes per side
  static constexpr int S = (32 - 2 * HL) * CW;          // valid columns/strip
  static constexpr int XOFF = -HL * CW;                 // x of lane 0, cell 0
  // cp.async ring: per warp and slot, 32 lanes x (CW floats + CW codes).
  // (TMA builds: 16 more bytes of codes per slot -- the code segment is copied
  // from the 16-byte boundary below it -- and kRing mbarriers per warp.)
  static constexpr int kSlotBytes = 32 * CW * 6 + (PP2D_TMA ? 16 : 0);
  static constexpr int kCodeOff = 32 * CW * 4;
  static constexpr int kBarOff = kRing * kSlotBytes;
  static constexpr int kRingBytesPerWarp =
      (CW >= 2) ? kRing * kSlotBytes + (PP2D_TMA ? 32 : 0) : 0;
};

// LIN: units are runs of the strip-major row sequence (see run()); a separate
// instantiation because the loop around the marching code costs registers the
// row-block scheme cannot spare (126 of 128).
template <int T, int CW, bool POLICY, bool P2P = false, bool LIN = false>
struct Sweeper {
  using G = StripGeom<T, CW>;
  // bulk-copy ring: the fused row-block kernel only (one march per warp)
  static constexpr bool kTma = PP2D_TMA && T == 2 && CW == 2 && !P2P && !LIN;
  uint32_t ring_base = 0;    // kTma: shared address of this warp's ring (slot 0)
  int lane_id = 0;
  float A[3][CW + 2];        // J^0 rows y-1, y, y+1 (rotating)
  float B[3][CW + 2];        // J^1 rows y-2, y-1, y (T == 2)
  float4 L[2][CW][4];        // LUT rows of row y (cur) and y-1 (prev)
  float G4[2][CW];
  float nxt[kPrefetch][CW];              // prefetched raw J^0 rows
  uint32_t cnx[kPrefetch][(CW + 1) / 2]; // prefetched code rows
  uint32_t cprev[(CW + 1) / 2];          // codes of row y-1 (T == 2 && POLICY)

  const SweepParams& p;
  uint32_t lane_base;        // shared address of this lane's table replica
  const float* pin;          // next J^0 row to prefetch (row y+1+kPrefetch)
  const uint16_t* pcode;     // next code row to prefetch (row y+kPrefetch)
  float* pout;               // row written at step y (y for T=1, y-1 for T=2)
  uint8_t* pact;             // action row y (POLICY)
  float *pup, *pdn;          // P2P: this lane's column in the neighbours' ghost rows
  bool peer_top, peer_bot;   // P2P: this segment feeds the upper / lower neighbour (warp-uniform)
  int yout;                  // P2P: row index of pout
  size_t l1_ahead;           // CW == 1: rows ahead for prefetch.global.L1
  uint32_t ring_j, ring_c;   // CW >= 2: shared addresses of this lane's ring slots
  int x0;                    // map x of the lane's first cell
  bool valid;

  __device__ __forceinline__ Sweeper(const SweepParams& p_, uint32_t lb,
                                     uint32_t ring_warp)
      : p(p_), lane_base(lb), ring_j(ring_warp), ring_c(ring_warp) {}

  // One marching step on row y.  I: rotation index (compile time).  J2: also
  // produce the second sweep of row y-1 (steady state of T == 2).
  template <int I, bool J2, bool PEER = false>
  __device__ __forceinline__ void step() {
    constexpr int a0 = I % 3, a1 = (I + 1) % 3, a2 = (I + 2) % 3;
    constexpr int lc = I % 2, lp = (I + 1) % 2;
    uint32_t cc[(CW + 1) / 2];
    if constexpr (kTma) {
      constexpr int slot = I % kRing;
      // use number (step / kRing) of this slot: its parity is a compile-time
      // constant because the unrolled period (6) is a multiple of 2 * kRing
      static_assert(6 % (2 * kRing) == 0 || kRing == 3, "ring depth vs unroll");
      mbar_wait(ring_base + G::kBarOff + 8 * slot, (I / kRing) & 1);
      float own[CW];
      lds_row<CW, slot * G::kSlotBytes>(ring_j, own);
      lds_codes<CW, slot * G::kSlotBytes + G::kCodeOff>(ring_c, cc);
      fill_row<CW>(own, A[a2]);
    } else if constexpr (CW >= 2) {
      constexpr int slot = I % kRing;
      // The oldest of the kRing groups in flight (row y+1, codes of row y)
      // has landed in this lane's slot.
      cp_async_wait<kRing - 1>();
      float own[CW];
      lds_row<CW, slot * G::kSlotBytes>(ring_j, own);
      lds_codes<CW, slot * G::kSlotBytes + G::kCodeOff>(ring_c, cc);
      fill_row<CW>(own, A[a2]);
      // Refill the slot with row y+1+kRing / codes of row y+kRing (the planes
      // have kSlackRows spare rows, so the last steps may run past the end).
      // (Issued here, right after the slot was read: refilling at the END of the
      // step saves the three predicated-off LDS ptxas puts in front of the
      // LDGSTS pair -- 193 instead of 195 instructions per step -- but runs
      // slower, 64.8 vs 63.4 us per launch: the copy has one step less to land.)
      cp_async<CW * 4>(ring_j + slot * G::kSlotBytes, pin);
      cp_async<CW * 2>(ring_c + slot * G::kSlotBytes + G::kCodeOff, pcode);
      cp_async_commit();
    } else {
      constexpr int q = I % kPrefetch;
      fill_row<CW>(nxt[q], A[a2]);
      cc[0] = cnx[q][0];
      load_own<CW, !P2P>(pin, nxt[q]);
      load_codes<CW>(pcode, cnx[q]);
      asm volatile("prefetch.global.L1 [%0];" :: "l"(pin + l1_ahead));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(pcode + l1_ahead));
    }
    if constexpr (!kTma) {
      pin += p.pitch;
      pcode += p.pitch;
    }
    // Table rows of row y.
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      if (j & 1) lut_fetch<true>(lane_base, cc[j >> 1], L[lc][j], G4[lc][j]);
      else lut_fetch<false>(lane_base, cc[j >> 1], L[lc][j], G4[lc][j]);
    }
    float v1[CW];
    uint32_t act[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      v1[j] = backup<(POLICY && T == 1)>(A[a0][j], A[a0][j + 1], A[a0][j + 2],
                             A[a1][j], A[a1][j + 1], A[a1][j + 2],
                             A[a2][j], A[a2][j + 1], A[a2][j + 2],
                             L[lc][j], G4[lc][j], p.gamma, p.ga, p.gb, act[j]);
    }
    if constexpr (T == 1) {
      if (valid) {
        store_own<CW>(pout, v1);
        if constexpr (POLICY) {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            if (x0 + j < p.W) {
              const uint32_t cj = (j & 1) ? (cc[j >> 1] >> 16) : cc[j >> 1];
              // Occupied cells tie on every action in the reference -> 0.
              pact[j] = (cj & kCodeOccBit) ? 0 : (uint8_t)act[j];
            }
          }
        }
      }
      pout += p.pitch;
      if constexpr (POLICY) pact += p.W;
    } else {
      fill_row<CW>(v1, B[a2]);
      if constexpr (J2) {
        float v2[CW];
        uint32_t act2[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          // the arg-min of a fused launch belongs to its SECOND sweep
          v2[j] = backup<POLICY>(B[a0][j], B[a0][j + 1], B[a0][j + 2],
                                 B[a1][j], B[a1][j + 1], B[a1][j + 2],
                                 B[a2][j], B[a2][j + 1], B[a2][j + 2],
                                 L[lp][j], G4[lp][j], p.gamma, p.ga, p.gb,
                                 act2[j]);
        }
        if (valid) store_own<CW>(pout, v2);
        if constexpr (POLICY) {
          // (single predicated stores: no divergent region inside the marching loop)
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const uint32_t occ_bit = (j & 1) ? (kCodeOccBit << 16) : kCodeOccBit;
            // Occupied cells tie on every action in the reference -> 0.
            const uint32_t a = (cprev[j >> 1] & occ_bit) ? 0u : act2[j];
            if (valid && x0 + j < p.W) pact[j] = (uint8_t)a;
          }
        }
        if constexpr (PEER) {
          // The first / last two owned rows are also the neighbours' ghost
          // rows: written straight into their HBM over NVLink.  peer_top /
          // peer_bot and yout are warp-uniform, `valid` only predicates the
          // store (no divergent region in front of the next shuffle).
          if (peer_top && yout < kPadRows) {
            if (valid) store_own<CW>(pup + (size_t)yout * p.pitch, v2);
          }
          if (peer_bot && yout >= p.H - kPadRows) {
            if (valid) store_own<CW>(pdn + (size_t)(yout - (p.H - kPadRows)) * p.pitch, v2);
          }
        }
        if constexpr (POLICY) pact += p.W;
        if constexpr (PEER) ++yout;
        pout += p.pitch;
      }
      if constexpr (POLICY) {
#pragma unroll
        for (int j = 0; j < (CW + 1) / 2; ++j) cprev[j] = cc[j];
      }
    }
    if constexpr (kTma) {
      // Refill the slot read at the top of this step (its values have been
      // consumed by the arithmetic above) with row y+1+kRing / codes of row
      // y+kRing; pcode points at the 16-byte boundary below the code segment.
      constexpr int slot = I % kRing;
      tma_fill<32 * CW * 4, 32 * CW * 2 + 16>(
          lane_id, ring_base + G::kBarOff + 8 * slot, ring_base + slot * G::kSlotBytes, pin,
          ring_base + slot * G::kSlotBytes + G::kCodeOff, pcode);
      pin += p.pitch;
      pcode += p.pitch;
    }
  }

  // March over rows [y0, y1) of this lane's columns.  PEER: rows 0,1 / H-2,H-1
  // are also stored into the neighbours' ghost rows.
  template <bool PEER>
  __device__ __forceinline__ void march(int y0, int y1, size_t col) {
    const size_t pitch = (size_t)p.pitch;
    // First row stepped on: y0-1 for T=2 (J^1 of the row above), y0 for T=1.
    const int ys = (T == 2) ? y0 - 1 : y0;
    int steps = y1 - ys + (T == 2 ? 1 : 0);               // rows ys .. ye
    if constexpr (PEER) yout = y0;
    const float* jin = p.jin + col + (size_t)(ys - 1 + kPadRows) * pitch;
    {
      float r[CW];
      load_own<CW, !P2P>(jin, r);
      fill_row<CW>(r, A[0]);
      load_own<CW, !P2P>(jin + pitch, r);
      fill_row<CW>(r, A[1]);
    }
    pin = jin + 2 * pitch;                                  // row ys+1
    pcode = p.code + col + (size_t)(ys + kPadRows) * pitch; // row ys
    const int total_steps = steps;
    if constexpr (kTma) {
      // lane 0's pointers are the warp's row segment; codes are fetched from
      // the 16-byte boundary below it and read back at that offset
      const uint32_t coff =
          __shfl_sync(0xffffffffu, (uint32_t)(reinterpret_cast<uintptr_t>(pcode) & 15u), 0);
      pcode = reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(pcode) - coff);
      ring_c += coff;
#pragma unroll
      for (int d = 0; d < kRing; ++d) mbar_init_lane0(lane_id, ring_base + G::kBarOff + 8 * d, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int d = 0; d < kRing; ++d) {
        tma_fill<32 * CW * 4, 32 * CW * 2 + 16>(
            lane_id, ring_base + G::kBarOff + 8 * d, ring_base + d * G::kSlotBytes, pin,
            ring_base + d * G::kSlotBytes + G::kCodeOff, pcode);
        pin += pitch;
        pcode += pitch;
      }
    } else if constexpr (CW >= 2) {
#pragma unroll
      for (int d = 0; d < kRing; ++d) {
        cp_async<CW * 4>(ring_j + d * G::kSlotBytes, pin);
        cp_async<CW * 2>(ring_c + d * G::kSlotBytes + G::kCodeOff, pcode);
        cp_async_commit();
        pin += pitch;
        pcode += pitch;
      }
    } else {
#pragma unroll
      for (int d = 0; d < kPrefetch; ++d) {
        load_own<CW, !P2P>(pin, nxt[d]);
        load_codes<CW>(pcode, cnx[d]);
        pin += pitch;
        pcode += pitch;
      }
    }
    pout = p.jout + col + (size_t)(y0 + kPadRows) * pitch;
    if constexpr (POLICY) pact = p.action + (size_t)y0 * p.W + x0;
    if constexpr (T == 2) {
      // Two priming steps (rows y0-1 and y0) produce J^1 only.
      step<0, false, PEER>();
      step<1, false, PEER>();
      steps -= 2;
      while (true) {
        step<2, true, PEER>(); if (--steps == 0) break;
        step<3, true, PEER>(); if (--steps == 0) break;
        step<4, true, PEER>(); if (--steps == 0) break;
        step<5, true, PEER>(); if (--steps == 0) break;
        step<0, true, PEER>(); if (--steps == 0) break;
        step<1, true, PEER>(); if (--steps == 0) break;
      }
    } else {
      while (true) {
        step<0, false>(); if (--steps == 0) break;
        step<1, false>(); if (--steps == 0) break;
        step<2, false>(); if (--steps == 0) break;
        step<3, false>(); if (--steps == 0) break;
        step<4, false>(); if (--steps == 0) break;
        step<5, false>(); if (--steps == 0) break;
      }
    }
    // Do not leave the slot ring with copies in flight.
    if constexpr (kTma) {
      // slot d was filled once by the priming and once by every step j with
      // j % kRing == d; wait for its last fill (use number fills - 1)
#pragma unroll
      for (int d = 0; d < kRing; ++d) {
        const int fills = 1 + (total_steps > d ? (total_steps - d + kRing - 1) / kRing : 0);
        mbar_wait(ring_base + G::kBarOff + 8 * d, (uint32_t)(fills - 1) & 1u);
      }
    } else if constexpr (CW >= 2) cp_async_wait<0>();
  }

  // A unit is a run of `lin_len` rows of the strip-major sequence (strip 0 rows
  // y_begin..y_end-1, strip 1 rows ..., ...): every warp of the one-wave launch
  // gets the same number of rows whatever W and H are (with whole row blocks
  // per strip, 274 strips x 8 blocks fill only 92.6 % of the 2368 warp slots at
  // W = 16384).  A unit that crosses a strip boundary marches twice.
  __device__ __forceinline__ void run(int unit, int lane) {
    if constexpr (LIN) {
      l1_ahead = (size_t)(p.prefetch_rows - kPrefetch) * (size_t)p.pitch;
      if constexpr (CW >= 2) {
        ring_j += lane * (CW * 4);
        ring_c += lane * (CW * 2);
      }
      const int R = p.y_end - p.y_begin;
      int lo, hi;
      if (P2P && p.unit_lo != nullptr) {
        // (through lane 0: the compiler must keep seeing warp-uniform bounds)
        lo = __shfl_sync(0xffffffffu, __ldg(p.unit_lo + unit), 0);
        hi = __shfl_sync(0xffffffffu, __ldg(p.unit_lo + unit + 1), 0);
      } else {
        lo = unit * p.lin_len;
        hi = min(lo + p.lin_len, p.n_strips * R);
      }
      while (lo < hi) {
        const int k = lo / R;
        const int a = lo - k * R;
        const int b = min(R, a + (hi - lo));
        run_segment(k, p.y_begin + a, p.y_begin + b, lane);
        lo += b - a;
      }
    } else {
      const int k = unit % p.n_strips;
      const int rb = unit / p.n_strips;
      int y0, y1;                                             // rows [y0, y1)
      if (P2P && p.rows_edge > 0) {
        y0 = rb == 0 ? p.y_begin : p.y_begin + p.rows_edge + (rb - 1) * p.rows_inner;
        y1 = min(p.y_begin + p.rows_edge + rb * p.rows_inner, p.y_end);
      } else {
        y0 = p.y_begin + rb * p.rows_per_unit;
        y1 = min(y0 + p.rows_per_unit, p.y_end);
      }
      run_segment(k, y0, y1, lane);
    }
  }

  // Rows [y0, y1) of strip k.
  __device__ __forceinline__ void run_segment(const int k, const int y0, const int y1,
                                              const int lane) {
    x0 = k * G::S + G::XOFF + lane * CW;
    valid = (lane >= G::HL) && (lane < 32 - G::HL) && (x0 < p.W);
    const size_t col = (size_t)(x0 + kPadLeft);
    if constexpr (!LIN) {
      l1_ahead = (size_t)(p.prefetch_rows - kPrefetch) * (size_t)p.pitch;
      if constexpr (CW >= 2) {
        ring_base = ring_j;
        lane_id = lane;
        ring_j += lane * (CW * 4);
        ring_c += lane * (CW * 2);
      }
    }
    int r0 = y0, r1 = y1;            // rows marched without peer stores
    if constexpr (P2P) {
      // Units that touch the first / last two owned rows read ghost rows and
      // write the neighbours' ghost rows.  They process the `edge_rows` rows
      // next to the boundary FIRST and publish their flag right after, so the
      // hand-shake is off the critical path of the launch.
      const bool top = p.peer_up_out != nullptr && y0 < kPadRows;
      const bool bot = p.peer_down_out != nullptr && y1 > p.H - kPadRows;
      if (top || bot) {
        pup = p.peer_up_out + col;       // only dereferenced under peer_top / peer_bot
        pdn = p.peer_down_out + col;
        peer_top = top && !(p.p2p_debug & 2u);
        peer_bot = bot && !(p.p2p_debug & 2u);
        int e0 = y0, e1 = y1;
        const int e = p.edge_rows;
        if (top != bot && y1 - y0 > e) {
          if (top) { e1 = y0 + e; r0 = e1; }
          else     { e0 = y1 - e; r1 = e0; }
        } else {
          r1 = r0;                     // the whole unit is the edge segment
        }
        if (!(p.p2p_debug & 1u)) {
          // Wait until the neighbour's previous launch has (a) written my
          // ghost rows of this buffer parity and (b) finished reading its own
          // ghost rows that I am about to overwrite; both are implied by its
          // flag, published after ALL its edge segments facing me are done.
          // Every lane polls (one broadcast transaction per poll) and the
          // value is taken from lane 0 through a shuffle: the loop exit is
          // then warp-uniform FOR THE COMPILER, which keeps the marching code
          // below free of divergence guards (a lane-0-only spin loop cost the
          // whole kernel its uniformity: 63 -> 82 us per launch).
          const unsigned int want = p.iter - 1;
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            if (!(side == 0 ? top : bot)) continue;
            const unsigned int* f = p.flags + (side == 0 ? kFlagFromUp : kFlagFromDown);
            // A neighbour that never shows up (crashed rank, mismatched launch
            // counts) must not hang the GPU: after spin_limit polls the error
            // word is set.  The rows this launch produces from stale ghost rows
            // are wrong from here on; mdp_residual_kernel turns the error word
            // into an infinite residual and every host entry point that returns
            // results checks it, so the failure cannot pass silently.
            unsigned int spins = 0;
            while (true) {
              const unsigned int v = __shfl_sync(0xffffffffu, ld_acquire_sys(f), 0);
              if ((int)(v - want) >= 0) break;
              if (++spins > p.spin_limit) {
                if (lane == 0) atomicExch(p.flags + kFlagError, 1u);
                break;
              }
              __nanosleep(spins < 64u ? 32u : 256u);
            }
          }
        }
        march<true>(e0, e1, col);
        // Publish: the last edge segment on each side to finish writes this
        // launch's number into the neighbour's flag.  (All lanes fence; lane 0
        // counts; the count comes back through a shuffle so that the decision
        // is warp-uniform.)
        if (!(p.p2p_debug & 4u)) {
          // Release of this unit's peer stores into the local count.  A system
          // fence waits for the flush of EVERY egress port of the GPU, PCIe
          // included: while a download is on its way to the host it takes tens
          // of microseconds, and a marching warp that executes it becomes the
          // tail of the launch (measured: +17 us per launch at 2 GPUs).  So the
          // marching warps only order their stores at GPU scope (barrier, then
          // fence, then lane 0's atomic: the release pattern of the PTX memory
          // model, cumulative over the whole warp's stores); the one system
          // fence per side and launch is executed by whoever publishes the
          // flag -- a warp without rows (publish_flags) when there is one.
          const bool gpu_scope = p.publisher_unit >= 0 || (p.p2p_debug & 8u);
          if (gpu_scope) {
            __syncwarp();
            __threadfence();
          } else {
            __threadfence_system();        // every lane: its peer stores before the count
            __syncwarp();                  // ... and all of them before lane 0 counts
          }
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            if (!(side == 0 ? top : bot)) continue;
            unsigned int c = 0;
            if (lane == 0)
              c = atomicAdd(p.flags + (side == 0 ? kFlagCountTop : kFlagCountBot), 1u) + 1u;
            c = __shfl_sync(0xffffffffu, c, 0);
            if (p.publisher_unit < 0 && c == (side == 0 ? p.expect_top : p.expect_bot)) {
              __threadfence_system();      // (acquires the other units' counts, too)
              if (lane == 0)
                st_release_sys(side == 0 ? p.up_flag_remote : p.down_flag_remote, p.iter);
            }
          }
        }
      }
    }
    if (r1 > r0) march<false>(r0, r1, col);
  }
};

// The publisher warp of a P2P launch (SweepParams::publisher_unit): waits until
// every boundary unit of a side has counted itself in (their peer stores happen
// before their count, the fence after the last poll acquires it), executes the system fence
// and writes this launch's number into the neighbour's flag.  The counts are
// cumulative over launches.  Lane 0 serves the upper, lane 1 the lower side.
__device__ __forceinline__ void publish_flags(const SweepParams& p, int lane) {
  if (p.p2p_debug & 4u) return;
  if (lane < 2) {
    unsigned int* remote = lane == 0 ? p.up_flag_remote : p.down_flag_remote;
    const bool feeds = (lane == 0 ? p.peer_up_out : p.peer_down_out) != nullptr;
    if (feeds && remote != nullptr) {
      const unsigned int* cnt = p.flags + (lane == 0 ? kFlagCountTop : kFlagCountBot);
      const unsigned int expect = lane == 0 ? p.expect_top : p.expect_bot;
      unsigned int spins = 0;
      bool ok = true;
      // (relaxed polls: an acquire load invalidates the L1 of the SM on every
      // poll, which the marching warps of this SM pay for -- 68.0 against 65.8
      // us per launch; the fence below is the acquire)
      while ((int)(ld_relaxed_gpu(cnt) - expect) < 0) {
        if (++spins > p.spin_limit) {       // a unit never finished: do not publish
          atomicExch(p.flags + kFlagError, 1u);
          ok = false;
          break;
        }
        __nanosleep(200u);
      }
      if (ok) {
        __threadfence_system();             // acquire of the counts + release of the flag
        st_release_sys(remote, p.iter);
      }
    }
  }
}

// Dynamic shared memory of a launch (PP2D_LUT6 builds; the 8 KB table keeps
// its static buffer so that the default build's code is untouched).
template <int T, int CW>
constexpr size_t sweep_smem_bytes() {
#if PP2D_LUT6
  return (size_t)kLutFloat4 * 16 + kLutAlign + (size_t)kWarpsPerCta * StripGeom<T, CW>::kRingBytesPerWarp;
#else
  return 0;
#endif
}

// One warp per (column strip, row block) unit; kWarpsPerCta warps per CTA.
template <int T, int CW, bool POLICY, bool P2P = false, bool LIN = false>
#ifdef PP2D_MAXNREG
__global__ void __maxnreg__(PP2D_MAXNREG)
#else
__global__ void __launch_bounds__(kWarpsPerCta * 32, (T == 2 && CW == 2) ? PP2D_MINCTAS : 1)
#endif
mdp_sweep_kernel(const SweepParams p) {
  // The table must start on a 2 KB boundary of the shared window so that the
  // row offset (bits 7..10) and the lane replica (bits 4..6) can be OR-ed
  // into the base.  The window itself starts at 1 KB (system reserved), so
  // the alignment is done at run time on an over-allocated buffer.
  using G = StripGeom<T, CW>;
  constexpr int kWarps = kWarpsPerCta;
  // Programmatic dependent launch: the next launch of the stream may start its
  // prologue (the table copy below, which reads nothing a sweep writes) while
  // this one is still running; it blocks in griddepcontrol.wait until this grid
  // has completed and its writes are visible.  Both are no-ops for a launch
  // without the attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#if PP2D_LUT6
  extern __shared__ __align__(16) unsigned char lut_raw[];   // sweep_smem_bytes<T, CW>()
#else
  __shared__ __align__(16) unsigned char
      lut_raw[kLutFloat4 * 16 + kLutAlign + kWarps * G::kRingBytesPerWarp];
#endif
  const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(lut_raw);
  const uint32_t lut_addr = (raw_addr + (uint32_t)(kLutAlign - 1)) & ~(uint32_t)(kLutAlign - 1);
  float4* lut_s = reinterpret_cast<float4*>(lut_raw + (lut_addr - raw_addr));
  // p.lut holds ONE copy of every row (kLutFloat4 / 8 rows); the 8 lane
  // replicas are written here: one global load and 8 shared stores per row.
  for (int i = threadIdx.x; i < kLutFloat4 / 8; i += blockDim.x) {
    const float4 row = p.lut[i];
#pragma unroll
    for (int c = 0; c < 8; ++c) lut_s[i * 8 + c] = row;
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
#if PP2D_UWARP
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
#else
  const int warp = (int)(threadIdx.x >> 5);
#endif
  const int unit = blockIdx.x * (blockDim.x >> 5) + warp;
  if constexpr (P2P) {
    if (unit == p.publisher_unit) {
      publish_flags(p, lane);
      return;
    }
  }
  if (unit >= p.n_units) return;
  // Produced by a volatile asm after the barrier: the table loads (plain asm,
  // free to be scheduled) depend on it and so cannot move above the barrier.
  uint32_t lane_base;
  asm volatile("or.b32 %0, %1, %2;"
               : "=r"(lane_base) : "r"(lut_addr), "r"((lane & 7) << 4) : "memory");
  const uint32_t ring_warp =
      lut_addr + kLutFloat4 * 16 + warp * G::kRingBytesPerWarp;
  Sweeper<T, CW, POLICY, P2P, LIN> s(p, lane_base, ring_warp);
  s.run(unit, lane);
}

// ---------------------------------------------------------------------------
// Code plane from the occupancy grid (replaces cudaGenerateModelData's crop of
// the 3x3 neighbourhood, path_planning_2d_cuda.cu:185-196: out of map =
// occupied).  occ holds global rows [occ_row0, occ_row0 + occ_rows).
struct CodeParams {
  const uint8_t* occ;
  uint16_t* code;
  int W, Htot, pitch;
  int rows_phys;       // owned rows + 2*kPadRows
  int row_begin;       // global row of local y = 0
  int occ_row0, occ_rows;
  int gx, gy;
};

// Occupancy of the 6 columns x0-1 .. x0+4 of global row gy as a bit mask (bit i
// = column x0-1+i; out of map = occupied, a cell is occupied iff its byte is
// exactly 1).  FAST: x0 is 4-aligned, inside the map with 3 columns to spare
// and W % 4 == 0, so the 4 own bytes are one aligned 32-bit load.
template <bool FAST>
__device__ __forceinline__ uint32_t occ_mask6(const CodeParams& p, int x0, int gy) {
  const int r = gy - p.occ_row0;
  if (gy < 0 || gy >= p.Htot || r < 0 || r >= p.occ_rows) return 0x3Fu;
  const uint8_t* __restrict__ q = p.occ + (size_t)r * p.W;
  uint32_t m = 0;
  if (FAST) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(q + x0));
    m |= ((w & 0xFFu) == 1u) << 1;
    m |= (((w >> 8) & 0xFFu) == 1u) << 2;
    m |= (((w >> 16) & 0xFFu) == 1u) << 3;
    m |= ((w >> 24) == 1u) << 4;
    m |= (x0 > 0 ? (uint32_t)(__ldg(q + x0 - 1) == 1) : 1u);
    m |= (x0 + 4 < p.W ? (uint32_t)(__ldg(q + x0 + 4) == 1) : 1u) << 5;
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int gx = x0 - 1 + i;
      m |= ((gx < 0 || gx >= p.W) ? 1u : (uint32_t)(__ldg(q + gx) == 1)) << i;
    }
  }
  return m;
}

// One thread builds the codes of 4 consecutive cells and marches kCodeRows rows
// down with the three 6-column occupancy masks of rows y-1, y, y+1 in
// registers: 3 loads (1 word + 2 bytes) and one 8-byte store per 4 cells
// instead of 9 byte loads and a 2-byte store per cell.
constexpr int kCodeRows = 16;
constexpr int kCodeThreads = 128;

template <bool FAST>
__device__ __forceinline__ void code_march(const CodeParams& p, int c0, int r0) {
  const int x0 = c0 - kPadLeft;
  const int r1 = min(r0 + kCodeRows, p.rows_phys);
  int gy = p.row_begin + r0 - kPadRows;
  uint32_t m0 = occ_mask6<FAST>(p, x0, gy - 1), m1 = occ_mask6<FAST>(p, x0, gy);
  for (int r = r0; r < r1; ++r, ++gy) {
    const uint32_t m2 = occ_mask6<FAST>(p, x0, gy + 1);
    uint32_t code[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j;
      uint32_t cd = kCodeOccBit;   // padding
      if (x >= 0 && x < p.W && gy >= 0 && gy < p.Htot) {
        // slots: 0 1 2 / 3 4 5 / 6 7 8 ; ring: s0 s1 s2 s5 s8 s7 s6 s3 s0 s1
        const uint32_t t = m0 >> j, c = m1 >> j, b = m2 >> j;
        const uint32_t s0 = t & 1u, s1 = (t >> 1) & 1u, s2 = (t >> 2) & 1u, s3 = c & 1u,
                       s4 = (c >> 1) & 1u, s5 = (c >> 2) & 1u, s6 = b & 1u,
                       s7 = (b >> 1) & 1u, s8 = (b >> 2) & 1u;
        cd = s0 | (s1 << 1) | (s2 << 2) | (s5 << 3) | (s8 << 4) | (s7 << 5) |
             (s6 << 6) | (s3 << 7) | (s0 << 8) | (s1 << 9);
        if (s4) cd |= kCodeOccBit;
        else if (!(x == p.gx && gy == p.gy)) cd |= kCodeLiveBit;
      }
      code[j] = cd;
    }
    // pitch is a multiple of 32 and c0 of 4: the 4 codes are 8-byte aligned
    *reinterpret_cast<uint2*>(p.code + (size_t)r * p.pitch + c0) =
        make_uint2(code[0] | (code[1] << 16), code[2] | (code[3] << 16));
    m0 = m1;
    m1 = m2;
  }
}

// grid: (ceil(pitch / 4 / kCodeThreads), ceil(rows_phys / kCodeRows))
__global__ void __launch_bounds__(kCodeThreads) mdp_code_kernel(const CodeParams p) {
  const int c0 = 4 * (blockIdx.x * kCodeThreads + threadIdx.x);
  const int r0 = blockIdx.y * kCodeRows;
  if (c0 >= p.pitch || r0 >= p.rows_phys) return;
  const int x0 = c0 - kPadLeft;
  const bool fast = (p.W & 3) == 0 && (x0 & 3) == 0 && x0 >= 0 && x0 + 3 < p.W &&
                    (reinterpret_cast<uintptr_t>(p.occ) & 3u) == 0;
  if (fast) code_march<true>(p, c0, r0);
  else code_march<false>(p, c0, r0);
}

// ---------------------------------------------------------------------------
// max |J - Jchk| over the owned rows, then Jchk = J
// (path_planning_2d.cu:243-251).  result: float bits, atomicMax on uint is
// order preserving for non-negative floats.
// CHK_ZERO: first check after a reset -- Jchk is J_0 = 0 by definition and is
// not read (the reset then does not have to zero it).
template <bool CHK_ZERO>
__global__ void __launch_bounds__(256)
mdp_residual_kernel(const float4* __restrict__ j, float4* __restrict__ chk,
                    size_t n4, uint32_t floor_bits, uint32_t* result,
                    const unsigned int* p2p_error) {
  float m = __uint_as_float(floor_bits);
  // a ghost-row hand-shake timed out: J is not to be trusted, and the
  // stopping rule must never be satisfied by it
  if (p2p_error != nullptr && *p2p_error != 0u) m = __uint_as_float(0x7f800000u);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = j[i];
    const float4 b = CHK_ZERO ? make_float4(0.f, 0.f, 0.f, 0.f) : chk[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x - b.x), fabsf(a.y - b.y)),
                       fmaxf(fabsf(a.z - b.z), fabsf(a.w - b.w))));
    chk[i] = a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, wm[i]);
    atomicMax(result, __float_as_uint(m));
  }
}

// ---------------------------------------------------------------------------
// Policy iteration (path_planning_2d_cuda.cu:266-355, called from the
// reference's policyIteration(), path_planning_2d.cu:271-357 -- dead code
// there).  Not a hot path: one thread per cell,
// the model of the cell rebuilt from its code word.
//   evaluation:  J'(s) = g(s,u) + sum_k P(s,u,k) * (gamma * J(n_k)),  u = pi(s)
//   improvement: pi(s) = first arg-min over u of the same expression
// Arithmetic of the reference kernels (nvcc 12.9, sm_100a): t_k = FMUL(J_k,
// gamma) once per neighbour, cost = FFMA(P_k, t_k, cost) for k = 0..8 --
// (gamma*J)*P, NOT the (gamma*P)*J of the value-iteration kernel, all with
// flush-to-zero (J of the goal decays into the denormal range).  Zero
// probabilities contribute fma(0, t, c) = c and are skipped.  Occupied cells
// (stored as 0, see the header of this file) follow J_n = (gamma*J_{n-1}) + 2
// and keep action 0; unlike in value iteration the goal is an ordinary cell
// here (its J is only 0 once the policy says "stay").
struct PolicyParams {
  const float* jin;
  float* jout;             // evaluation only
  const uint16_t* code;
  uint8_t* action;
  int W, H, pitch;
  float gamma;
};

// P(s,u,.) and g(s,u) of one action from the ring code of the cell
// (path_planning_2d_cuda.cu:76-172).  blocked: bit i = slot i occupied or out
// of map (bit 4 unused: the cell itself is free here).
__device__ __forceinline__ void policy_model(uint32_t blocked, bool is_goal, int u,
                                             float (&P)[9], float& g) {
  // slots with non-zero naive probability, ascending; side[u] as in the
  // reference's switch
  const int s0[9] = {0, 0, 1, 0, 4, 2, 3, 4, 4};
  const int s1[9] = {1, 1, 2, 3, 4, 4, 4, 6, 5};
  const int s2[9] = {3, 2, 4, 4, 4, 5, 6, 7, 7};
  const int s3[9] = {4, 4, 5, 6, 4, 8, 7, 8, 8};
#pragma unroll
  for (int i = 0; i < 9; ++i) P[i] = 0.0f;
  if (u == 4) {
    P[4] = 1.0f;
    g = is_goal ? 0.0f : 2.0f;                       // cuda.cu:171
    return;
  }
  const int sl[4] = {s0[u], s1[u], s2[u], s3[u]};
  float naive[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) naive[j] = (sl[j] == u) ? 0.7f : 0.1f;
  // stage cost over the NAIVE probabilities, ascending slot (cuda.cu:166-169)
  g = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    g = fmaf(((blocked >> sl[j]) & 1u) && sl[j] != 4 ? 2.0f : 1.0f, naive[j], g);
  // blocked mass moves to "stay", ascending slot (cuda.cu:142-147)
  float stay = 0.1f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (sl[j] == 4) continue;
    if ((blocked >> sl[j]) & 1u) stay += naive[j];
    else P[sl[j]] = naive[j];
  }
  P[4] = stay;
}

// ring bit r of the code -> neighbour slot: s0 s1 s2 s5 s8 s7 s6 s3
__device__ __forceinline__ uint32_t blocked_slots(uint32_t code) {
  return ((code >> 0) & 1u) << 0 | ((code >> 1) & 1u) << 1 | ((code >> 2) & 1u) << 2 |
         ((code >> 3) & 1u) << 5 | ((code >> 4) & 1u) << 8 | ((code >> 5) & 1u) << 7 |
         ((code >> 6) & 1u) << 6 | ((code >> 7) & 1u) << 3;
}

template <bool IMPROVE>
__global__ void __launch_bounds__(256)
mdp_policy_kernel(const PolicyParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.W || y >= p.H) return;
  const size_t q = (size_t)(y + kPadRows) * p.pitch + x + kPadLeft;
  const uint32_t code = p.code[q];
  const size_t dense = (size_t)y * p.W + x;
  if (code & kCodeOccBit) {                // trapped: closed form on download
    if (IMPROVE) p.action[dense] = 0;
    else p.jout[q] = 0.0f;
    return;
  }
  const bool is_goal = !(code & kCodeLiveBit);
  const uint32_t blocked = blocked_slots(code);
  float t[9];
#pragma unroll
  for (int i = 0; i < 9; ++i)
    t[i] = mul_ftz(p.jin[q + (size_t)((i / 3 - 1) * p.pitch) + (i % 3 - 1)], p.gamma);
  if (IMPROVE) {
    float opt = 3.402823466e+38f;
    int oa = 0;
    for (int u = 0; u < 9; ++u) {
      float P[9], g;
      policy_model(blocked, is_goal, u, P, g);
      float cost = g;
#pragma unroll
      for (int i = 0; i < 9; ++i)
        if (P[i] != 0.0f) cost = fma_ftz(P[i], t[i], cost);
      if (cost < opt) { opt = cost; oa = u; }
    }
    p.action[dense] = (uint8_t)oa;
  } else {
    float P[9], g;
    policy_model(blocked, is_goal, p.action[dense], P, g);
    float cost = g;
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (P[i] != 0.0f) cost = fma_ftz(P[i], t[i], cost);
    p.jout[q] = cost;
  }
}

// Dense J for download: occupied cells get the closed-form trapped cost.
// VEC: 4 cells per thread (W % 4 == 0 and 4-aligned pad columns: 16-byte loads
// and stores); otherwise one cell per thread.
template <bool VEC>
__global__ void mdp_export_kernel(const float* __restrict__ j,
                                  const uint16_t* __restrict__ code,
                                  float* __restrict__ out, int W, int H,
                                  int pitch, float occupied_cost) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 4 : 1);
  const int y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t q = (size_t)(y + kPadRows) * pitch + x + kPadLeft;
  if (VEC) {
    float4 v = *reinterpret_cast<const float4*>(j + q);
    const uint2 c = *reinterpret_cast<const uint2*>(code + q);
    if (c.x & kCodeOccBit) v.x = occupied_cost;
    if (c.x & (kCodeOccBit << 16)) v.y = occupied_cost;
    if (c.y & kCodeOccBit) v.z = occupied_cost;
    if (c.y & (kCodeOccBit << 16)) v.w = occupied_cost;
    *reinterpret_cast<float4*>(out + (size_t)y * W + x) = v;
  } else {
    out[(size_t)y * W + x] = (code[q] & kCodeOccBit) ? occupied_cost : j[q];
  }
}

// MdpPathPlanning2d::beliefCallback (path_planning_2d.cu:168-189): index of
// the first strict maximum of the belief starting from (0.0f, index 0), then
// the action stored there.  One CTA per belief.
__global__ void __launch_bounds__(256)
mdp_plan_kernel(const float* __restrict__ beliefs, size_t n,
                const uint8_t* __restrict__ action, uint8_t* __restrict__ out,
                unsigned long long* __restrict__ out_index) {
  const float* b = beliefs + (size_t)blockIdx.x * n;
  float bm = 0.0f;
  unsigned long long bi = 0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    float v = b[i];
    if (v > bm) { bm = v; bi = i; }   // ascending i per thread: first max kept
  }
  // (value, index) reduction: larger value wins, ties -> smaller index.
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bm, o);
    unsigned long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bm || (ov == bm && oi < bi)) { bm = ov; bi = oi; }
  }
  __shared__ float sv[8];
  __shared__ unsigned long long si[8];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bm; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i)
      if (sv[i] > bm || (sv[i] == bm && si[i] < bi)) { bm = sv[i]; bi = si[i]; }
    // all beliefs <= 0: the reference keeps index 0.
    const unsigned long long mode = bm > 0.0f ? bi : 0;
    // (a multi-GPU handle keeps the action grid spread over its devices: the
    // cell index goes back and the host looks the action up)
    if (action != nullptr) out[blockIdx.x] = action[mode];
    else out_index[blockIdx.x] = mode;
  }
}

}  // namespace pp2d
