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

namespace pp2d {

constexpr int kPadRows = 2;   // ghost rows above and below the owned rows
constexpr int kPadLeft = 8;   // zero columns left of x = 0 (32 B)
constexpr int kLutFloat4 = 4 * 16 * 8;  // 4 action pairs x 16 rows x 8 copies

// Code bits.
constexpr uint32_t kCodeRingMask = 0x3FFu;   // bits 0..9
constexpr uint32_t kCodeOccBit = 1u << 13;   // occupied or padding
constexpr uint32_t kCodeLiveBit = 1u << 14;  // free, in map, not the goal

struct SweepParams {
  const float* jin;      // padded plane, element (y, x) at [(y+2)*pitch + x+8]
  float* jout;
  const uint16_t* code;  // same geometry as J
  uint8_t* action;       // dense [H][W] (POLICY only)
  const float4* lut;     // kLutFloat4 entries, already lane-replicated
  int W, H, pitch;
  int n_strips, rows_per_unit, n_units;
  float gamma, ga, gb;   // gamma*1.0f, gamma*0.7f, gamma*0.1f
};

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
  } else {
    return min3(min3(c0, c1, c2), min3(c3, c4, c5), min3(c6, c7, c8));
  }
}

template <int CW>
__device__ __forceinline__ void load_own(const float* __restrict__ p,
                                         float (&o)[CW]) {
  if constexpr (CW == 1) {
    o[0] = __ldg(p);
  } else if constexpr (CW == 2) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    o[0] = v.x; o[1] = v.y;
  } else {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
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
    c[0] = __ldg(p);
  } else if constexpr (CW == 2) {
    c[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
  } else {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    c[0] = v.x; c[1] = v.y;
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

// Fetch the 4 table rows of one cell.  lut_lane already points at this
// lane's replica ((lane & 7) * 16 bytes into the table).
__device__ __forceinline__ void lut_fetch(const char* lut_lane, uint32_t code,
                                          float4 (&t)[4], float& g4) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint32_t off = (code & (0xFu << (2 * p))) << (7 - 2 * p);
    t[p] = *reinterpret_cast<const float4*>(lut_lane + p * 2048 + off);
  }
  // live bit (14) -> 2.0f (0x40000000); goal, occupied, padding -> 0.0f.
  g4 = __uint_as_float((code << 16) & 0x40000000u);
}

template <int T, int CW>
struct StripGeom {
  static constexpr int HL = (T + CW - 1) / CW;          // halo lanes per side
  static constexpr int S = (32 - 2 * HL) * CW;          // valid columns/strip
  static constexpr int XOFF = -HL * CW;                 // x of lane 0, cell 0
};

template <int T, int CW, bool POLICY>
struct Sweeper {
  using G = StripGeom<T, CW>;
  float A[3][CW + 2];        // J^0 rows y-1, y, y+1 (rotating)
  float B[3][CW + 2];        // J^1 rows y-2, y-1, y (T == 2)
  float4 L[2][CW][4];        // LUT rows of row y (cur) and y-1 (prev)
  float G4[2][CW];
  float nxt[CW];             // prefetched raw J^0 row
  uint32_t cnx[(CW + 1) / 2];  // prefetched codes of the next row

  const SweepParams& p;
  const char* lut_lane;
  const float* jin;          // lane's column, row 0
  float* jout;
  const uint16_t* code;
  int x0;                    // map x of the lane's first cell
  int y0, y1;                // owned rows of this unit [y0, y1)
  bool valid;

  __device__ __forceinline__ Sweeper(const SweepParams& p_, const char* lut_)
      : p(p_), lut_lane(lut_) {}

  template <int I>
  __device__ __forceinline__ void step(int y) {
    constexpr int a0 = I % 3, a1 = (I + 1) % 3, a2 = (I + 2) % 3;
    constexpr int lc = I % 2, lp = (I + 1) % 2;
    const int pitch = p.pitch;
    // Row y+1 arrived (prefetched one step ago): add the horizontal halo.
    fill_row<CW>(nxt, A[a2]);
    uint32_t cc[(CW + 1) / 2];
#pragma unroll
    for (int j = 0; j < (CW + 1) / 2; ++j) cc[j] = cnx[j];
    // Prefetch row y+2 and the codes of row y+1.
    const int ylast = (T == 2) ? y1 : y1 - 1;   // last y this unit steps on
    if (y < ylast) {
      load_own<CW>(jin + (size_t)(y + 2 + kPadRows) * pitch, nxt);
      load_codes<CW>(code + (size_t)(y + 1 + kPadRows) * pitch, cnx);
    }
    // Table rows of row y.
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      uint32_t cj = (j & 1) ? (cc[j >> 1] >> 16) : cc[j >> 1];
      lut_fetch(lut_lane, cj, L[lc][j], G4[lc][j]);
    }
    float v1[CW];
    uint32_t act[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      v1[j] = backup<POLICY>(A[a0][j], A[a0][j + 1], A[a0][j + 2],
                             A[a1][j], A[a1][j + 1], A[a1][j + 2],
                             A[a2][j], A[a2][j + 1], A[a2][j + 2],
                             L[lc][j], G4[lc][j], p.gamma, p.ga, p.gb, act[j]);
    }
    if constexpr (T == 1) {
      if (valid) {
        store_own<CW>(jout + (size_t)(y + kPadRows) * pitch, v1);
        if constexpr (POLICY) {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            if (x0 + j < p.W) {
              uint32_t cj = (j & 1) ? (cc[j >> 1] >> 16) : cc[j >> 1];
              // Occupied cells tie on every action in the reference -> 0.
              p.action[(size_t)y * p.W + x0 + j] =
                  (cj & kCodeOccBit) ? 0 : (uint8_t)act[j];
            }
          }
        }
      }
    } else {
      fill_row<CW>(v1, B[a2]);
      if (y > y0) {
        float v2[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          v2[j] = backup<false>(B[a0][j], B[a0][j + 1], B[a0][j + 2],
                                B[a1][j], B[a1][j + 1], B[a1][j + 2],
                                B[a2][j], B[a2][j + 1], B[a2][j + 2],
                                L[lp][j], G4[lp][j], p.gamma, p.ga, p.gb,
                                act[j]);
        }
        if (valid) store_own<CW>(jout + (size_t)(y - 1 + kPadRows) * pitch, v2);
      }
    }
  }

  __device__ __forceinline__ void run(int unit, int lane) {
    const int k = unit % p.n_strips;
    const int rb = unit / p.n_strips;
    y0 = rb * p.rows_per_unit;
    y1 = min(y0 + p.rows_per_unit, p.H);
    x0 = k * G::S + G::XOFF + lane * CW;
    valid = (lane >= G::HL) && (lane < 32 - G::HL) && (x0 < p.W);
    const size_t col = (size_t)(x0 + kPadLeft);
    jin = p.jin + col;
    jout = p.jout + col;
    code = p.code + col;
    const int pitch = p.pitch;
    // First row stepped on: y0-1 for T=2 (J^1 of the row above), y0 for T=1.
    const int ys = (T == 2) ? y0 - 1 : y0;
    const int ye = (T == 2) ? y1 : y1 - 1;
    {
      float r[CW];
      load_own<CW>(jin + (size_t)(ys - 1 + kPadRows) * pitch, r);
      fill_row<CW>(r, A[0]);
      load_own<CW>(jin + (size_t)(ys + kPadRows) * pitch, r);
      fill_row<CW>(r, A[1]);
      load_own<CW>(jin + (size_t)(ys + 1 + kPadRows) * pitch, nxt);
      load_codes<CW>(code + (size_t)(ys + kPadRows) * pitch, cnx);
    }
    int y = ys;
    while (true) {
      step<0>(y); if (++y > ye) break;
      step<1>(y); if (++y > ye) break;
      step<2>(y); if (++y > ye) break;
      step<3>(y); if (++y > ye) break;
      step<4>(y); if (++y > ye) break;
      step<5>(y); if (++y > ye) break;
    }
  }
};

// One warp per (column strip, row block) unit; 8 warps per CTA.
template <int T, int CW, bool POLICY>
__global__ void __launch_bounds__(256)
mdp_sweep_kernel(const SweepParams p) {
  __shared__ float4 lut_s[kLutFloat4];
  for (int i = threadIdx.x; i < kLutFloat4; i += blockDim.x) lut_s[i] = p.lut[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= p.n_units) return;
  Sweeper<T, CW, POLICY> s(p, reinterpret_cast<const char*>(lut_s) +
                                  (lane & 7) * 16);
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

__device__ __forceinline__ uint32_t occ_at(const CodeParams& p, int gx, int gy) {
  if (gx < 0 || gx >= p.W || gy < 0 || gy >= p.Htot) return 1u;
  int r = gy - p.occ_row0;
  if (r < 0 || r >= p.occ_rows) return 1u;   // never needed for owned rows
  return p.occ[(size_t)r * p.W + gx] == 1 ? 1u : 0u;
}

__global__ void mdp_code_kernel(const CodeParams p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= p.pitch || r >= p.rows_phys) return;
  const int x = c - kPadLeft;
  const int gy = p.row_begin + r - kPadRows;
  uint32_t code = kCodeOccBit;   // padding
  if (x >= 0 && x < p.W && gy >= 0 && gy < p.Htot) {
    // slots: 0 1 2 / 3 4 5 / 6 7 8 ; ring: s0 s1 s2 s5 s8 s7 s6 s3 s0 s1
    const uint32_t s0 = occ_at(p, x - 1, gy - 1), s1 = occ_at(p, x, gy - 1),
                   s2 = occ_at(p, x + 1, gy - 1), s3 = occ_at(p, x - 1, gy),
                   s4 = occ_at(p, x, gy), s5 = occ_at(p, x + 1, gy),
                   s6 = occ_at(p, x - 1, gy + 1), s7 = occ_at(p, x, gy + 1),
                   s8 = occ_at(p, x + 1, gy + 1);
    code = s0 | (s1 << 1) | (s2 << 2) | (s5 << 3) | (s8 << 4) | (s7 << 5) |
           (s6 << 6) | (s3 << 7) | (s0 << 8) | (s1 << 9);
    if (s4) code |= kCodeOccBit;
    else if (!(x == p.gx && gy == p.gy)) code |= kCodeLiveBit;
  }
  p.code[(size_t)r * p.pitch + c] = (uint16_t)code;
}

// ---------------------------------------------------------------------------
// max |J - Jchk| over the owned rows, then Jchk = J
// (path_planning_2d.cu:243-251).  result: float bits, atomicMax on uint is
// order preserving for non-negative floats.
__global__ void __launch_bounds__(256)
mdp_residual_kernel(const float4* __restrict__ j, float4* __restrict__ chk,
                    size_t n4, uint32_t floor_bits, uint32_t* result) {
  float m = __uint_as_float(floor_bits);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 a = j[i], b = chk[i];
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

// Dense J for download: occupied cells get the closed-form trapped cost.
__global__ void mdp_export_kernel(const float* __restrict__ j,
                                  const uint16_t* __restrict__ code,
                                  float* __restrict__ out, int W, int H,
                                  int pitch, float occupied_cost) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t q = (size_t)(y + kPadRows) * pitch + x + kPadLeft;
  out[(size_t)y * W + x] = (code[q] & kCodeOccBit) ? occupied_cost : j[q];
}

// MdpPathPlanning2d::beliefCallback (path_planning_2d.cu:168-189): index of
// the first strict maximum of the belief starting from (0.0f, index 0), then
// the action stored there.  One CTA per belief.
__global__ void __launch_bounds__(256)
mdp_plan_kernel(const float* __restrict__ beliefs, size_t n,
                const uint8_t* __restrict__ action, uint8_t* __restrict__ out) {
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
    out[blockIdx.x] = action[bm > 0.0f ? bi : 0];
  }
}

}  // namespace pp2d
