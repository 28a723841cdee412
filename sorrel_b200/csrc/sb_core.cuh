// sb_core.cuh -- per-lane building blocks of the step / observation kernels.
//
// Everything here is a pure function of its arguments (plus the shared-memory image of one world
// instance), marked SB_HD so that tests/hostcheck can compile the very same source with g++ and
// unit-test it against the oracle without a GPU.  The kernels in sb_step.cu only add the tiling,
// the synchronisation and the global-memory traffic around these functions.
#pragma once
#include <stdint.h>

#include "../../include/sorrel_b200.h"

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD static inline
#endif

// ------------------------------------------------------------------------------------------------
// Device-side view of SbWorldDesc: everything the kernels need, precomputed once per call on the
// host (sb_abi.cu: make_dev) and passed by value as a __grid_constant__ kernel parameter.
// ------------------------------------------------------------------------------------------------
struct SbDev {
  int32_t N, A, H, W, HW;
  int32_t plane_stride;  // bytes per stored plane (H*W rounded up to 16)
  int32_t n_planes;      // dynamic layers stored per instance
  int32_t env_stride;    // n_planes * plane_stride
  int32_t n_chunks;      // env_stride / 16
  int32_t agent_plane;   // plane index of the agent layer
  int32_t L;             // all layers (static ones live in tmpl)
  int32_t plane_of_layer[SB_MAX_LAYERS];  // -1 for static layers
  int32_t K, C, v, S, SS, obs_size;
  uint32_t inv_S;  // ceil(65536 / S): ij / S == (ij * inv_S) >> 16 for ij < 512
  uint32_t inv_W;  // ceil(2^32 / W) when cell / W == umulhi(cell, inv_W) for every cell < H*W (H*W*W < 2^32), else 0
  int32_t fill_kind, agent_kind, default_kind, n_actions, n_choices;
  int32_t fast_obs;           // appearance rows are zero / one-hot and static layers invisible
  int32_t fill_channel;       // fast path: channel of the fill kind (or 255 = none)
  int32_t single_spawn_kind;  // >= 0 when exactly one kind carries SB_RULE_SPAWN
  int32_t simd_sweep;         // byte-SIMD level-1 filter usable (single spawn kind, 1 <= n2 <= 128)
  uint32_t hmask4;            // ~(n2 - 1) replicated in 4 bytes
  uint32_t passable_mask, spawn_mask;  // bit k: kind k passable / has SB_RULE_SPAWN
  uint32_t n2;                         // level-1 bound of the spawn rule (sb_level1_bound): pass iff r8 < n2
  uint64_t thr64m1;                    // level-2 bound (sb_level2_bound): spawn iff the 64 fresh bits <= this
  uint64_t choices_packed;             // spawn_choice_kinds, one byte each
  uint64_t channel8;                   // channel[0..7] packed (used when K <= 8)
  uint64_t dy_packed, dx_packed;       // action deltas, one signed byte each
  uint64_t env_id_base;
  double p;
  float value[SB_MAX_KINDS];
  uint8_t channel[SB_MAX_KINDS];  // fast path: one-hot channel of the kind, 255 = all-zero row
  int32_t sparse_ok;              // spawn_cdf is valid (SB_RNG_SPARSE available)
  int32_t obs_transform;          // SB_OBS_*
  uint64_t cdf[SB_SPARSE_KMAX];   // Binomial(env_stride, p) CDF in 0.64 fixed point
  const double* app;              // device [K][C]
  const uint8_t* tmpl;            // device [L][plane_stride]
  // ---- column rules / facings / beams / positional code (ABI v2; examples/cleanup) ----
  int32_t oh_size;    // C*S*S: the crop part of an observation row; obs_size = oh_size + obs_extra
  int32_t obs_extra;  // floats appended from extra_tbl[cell]
  const float* extra_tbl;
  int32_t ext_rules;  // some kind carries a rule other than SB_RULE_SPAWN
  int32_t ext_act;    // facings / beams / all-layer rewards / reward_count != 1
  uint32_t draw_mask; // kinds whose rule consumes a uniform (SPAWN, SPAWN_TO)
  uint32_t det_mask;  // kinds with a deterministic rule (BECOME_IF, BECOME)
  uint32_t else_mask; // SPAWN_TO kinds whose failed draw still changes the kind (rule_else != SB_NO_KIND)
  uint8_t rule[SB_MAX_KINDS], rule_to[SB_MAX_KINDS], rule_else[SB_MAX_KINDS];
  int32_t toroidal;  // engine option: crops and moves wrap around the map edges (general observation path)
  int32_t agent_kind_alt, touch_swap, reward_own, obs_flag;  // agents with a second kind (tag)
  int32_t* agent_flag;           // device [N][A]
  int32_t use_value64;           // rewards are value64 (Python floats that float32 cannot hold)
  double value64[SB_MAX_KINDS];  // total_reward accumulates these; the ring stores (float)value64
  int32_t rule_plane_off[SB_MAX_KINDS];  // BECOME_IF: byte offset of the tested plane
  uint32_t rule_mask[SB_MAX_KINDS];
  uint64_t thr64m1k[SB_MAX_KINDS];  // level-2 bound per drawing kind (SPAWN / SPAWN_TO)
  uint16_t n2k[SB_MAX_KINDS];       // level-1 bound per drawing kind
  double prob[SB_MAX_KINDS];     // replay compares the recorded f64 uniform with this
  int8_t act_facing[SB_MAX_ACTIONS];
  uint8_t act_beam[SB_MAX_ACTIONS];
  int32_t beam_plane, n_beam;
  int8_t beam_dy[4][SB_MAX_BEAM_CELLS], beam_dx[4][SB_MAX_BEAM_CELLS];
  uint32_t beam_block_mask;
  int32_t reward_all, reward_count;
  int32_t par_radius;  // agents farther apart than this (Chebyshev) cannot interact within one turn: max(v + m, 2m)
  // row-wise window gather (narrow groups): channel8 with slot 7 := the fill kind's channel, so that an out-of-bounds
  // window cell becomes a lookup of "kind 7" by a byte mask instead of a branch.  Valid when K <= 7 (slot 7 unused).
  uint64_t channel8_oob;
  int32_t rowwise_ok;
  int32_t S5;  // S - S % 5: window columns covered by whole 5-cell groups of the row-wise gather
  // column gather of several staged planes (sb_step.cu: scatter_cols_multi): channel[8..15] packed like channel8, and
  // per group width 2^w the row segmentation -- cg_nseg[w] segments of cg_len[w] cells per window row
  // byte-SIMD classification of the generic sweep (sb_step.cu): per-kind class bytes for kinds 0..7 / 8..15
  uint64_t cls_and[2], cls_or[2];
  int32_t cls_ok;  // K <= 15 (slot 15 = the 0xFF padding byte: never visited)
  uint64_t channel8_hi;
  int32_t colgather_ok;  // fast_obs, 2 or 3 stored planes, K <= 16, not toroidal
  int8_t cg_nseg[6], cg_len[6];
  int8_t rw_nseg[6], rw_len[6];  // the same for the segmented row-wise gather of one stored plane (five cells per step)
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
// ------------------------------------------------------------------------------------------------
struct SbU4 {
  uint32_t x, y, z, w;
};

#if !defined(__CUDA_ARCH__)
static inline uint32_t sb_host_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
#define __umulhi sb_host_umulhi
#endif
SB_HD void sb_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;  // mul.wide.u32 -> one IMAD.WIDE.U32
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
}

SB_HD SbU4 sb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    sb_mulhilo(M0, c0, hi0, lo0);
    sb_mulhilo(M1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  SbU4 o;
  o.x = c0;
  o.y = c1;
  o.z = c2;
  o.w = c3;
  return o;
}

// The engine's keyed stream: Philox4x32-10 with a FIXED key (the first 64 fractional bits of pi, as in the Random123
// examples) and the 64-bit SEED folded into counter words 1 and 3:
//     draw(env, tick, index, domain) = Philox4x32-10((env, tick ^ seed_lo, index, domain ^ seed_hi), K)
// Every (seed, env, tick, index, domain) still maps to its own counter -- Philox is a bijection of the 128-bit counter
// for a fixed key, so distinct counters give distinct, statistically independent outputs -- but the ten round keys are
// now COMPILE-TIME constants: immediates of the XORs instead of twenty `LDC` loads (or twenty registers) per pair of
// calls in the sweep loop, which sits at the 64-register cap.  `sk` = {seed_lo, seed_hi}; the kernels pass a pointer
// into their __grid_constant__ parameters, so both XORs are loop-invariant uniform operations.
#define SB_PHILOX_K0 0xA4093822u
#define SB_PHILOX_K1 0x299F31D0u
SB_HD SbU4 sb_philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* __restrict__ sk) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  c1 ^= sk[0];
  c3 ^= sk[1];
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t k0 = SB_PHILOX_K0 + (uint32_t)r * 0x9E3779B9u, k1 = SB_PHILOX_K1 + (uint32_t)r * 0xBB67AE85u;
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  SbU4 o;
  o.x = c0;
  o.y = c1;
  o.z = c2;
  o.w = c3;
  return o;
}

SB_HD uint64_t sb_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// uniform integer in [0, n) from 64 random bits (multiply-shift; bias <= n / 2^64)
SB_HD uint32_t sb_bounded(uint32_t hi, uint32_t lo, uint32_t n) {
  return (uint32_t)sb_mulhi64(((uint64_t)hi << 32) | lo, (uint64_t)n);
}

SB_HD uint32_t sb_byte_of(const SbU4& r, int b) {
  uint32_t w = (b & 8) ? ((b & 4) ? r.w : r.z) : ((b & 4) ? r.y : r.x);
  return (w >> (8 * (b & 3))) & 0xFFu;
}

// ------------------------------------------------------------------------------------------------
// Counter-based Bernoulli draw (DESIGN.md "Counter-based RNG"): cell `cell` (= plane * plane_stride + y*W + x)
// of instance `env` at tick `tick` "hits" with probability EXACTLY P / 2^53, P = ceil(p * 2^53) -- the
// probability of the reference's `np.random.random() < p` -- evaluated in two levels so that almost every cell
// costs one random BYTE:
//   level 1: r8 = byte (cell & 15) of Philox(env, tick, cell >> 4, SWEEP); the cell passes iff r8 < n2, n2 = the
//            smallest power of two >= ((P - 1) >> 45) + 1 (the byte-SIMD filter of 16 cells per Philox call);
//   level 2: only for passing cells, 64 FRESH bits v from Philox(env, tick, cell, CELL): hit iff v < T,
//            T = P * 2^19 / n2 = P << (19 - log2 n2)   (<= 2^64 because P <= n2 * 2^45).
// P(hit) = (n2 / 256) * (T / 2^64) = P / 2^53.  The other 64 bits of the level-2 call choose the kind.
// (Round 1 refined the SAME 53-bit uniform lazily, which made the resolver recompute the level-1 call; two
// independent levels need one Philox call per candidate and no state carried from the filter.)
// ------------------------------------------------------------------------------------------------
SB_HD uint32_t sb_level1_bound(uint64_t P) {  // n2; 0 = never passes (p == 0)
  if (P == 0) return 0u;
  const uint32_t hi = (uint32_t)((P - 1) >> 45);  // <= 255
  uint32_t n2 = 1;
  while (n2 < hi + 1u) n2 <<= 1;
  return n2;
}

SB_HD uint64_t sb_level2_bound(uint64_t P, uint32_t n2) {  // T - 1 (T = 2^64 wraps to 0, minus one = all ones)
  int lg = 0;
  while ((1u << lg) < n2) ++lg;
  return (P << (19 - lg)) - 1ull;
}

SB_HD bool sb_level2_hit(uint64_t thr64m1, const SbU4& w) { return (((uint64_t)w.x << 32) | w.y) <= thr64m1; }

// Level 2 of the spawn rule for a cell that passed level 1.  Returns the kind to spawn, or -1.
SB_HD int sb_resolve_candidate(const SbDev& d, uint32_t env, uint32_t tick, uint32_t cell,
                               const uint32_t* __restrict__ rk) {
  const SbU4 w = sb_philox_rk(env, tick, cell, SB_DOM_CELL, rk);
  if (!sb_level2_hit(d.thr64m1, w)) return -1;
  const uint32_t ci = sb_bounded(w.z, w.w, (uint32_t)d.n_choices);
  return (int)((d.choices_packed >> (8 * ci)) & 0xFFu);
}

// SB_RULE_SPAWN_TO: the same two levels, no choice draw.  `r8` is the cell's level-1 byte; true = hit.
SB_HD bool sb_bernoulli_hit(uint32_t n2, uint64_t thr64m1, uint32_t r8, uint32_t env, uint32_t tick, uint32_t cell,
                            const uint32_t* __restrict__ rk) {
  if (r8 >= n2) return false;
  return sb_level2_hit(thr64m1, sb_philox_rk(env, tick, cell, SB_DOM_CELL, rk));
}

// Deterministic column rules (SB_RULE_BECOME_IF / SB_RULE_BECOME) for the cell at byte offset `cell` = plane
// base + `off` of the instance image `genv`.  Returns the new kind, or -1 when the cell keeps its kind.
// `ktab` (optional): the per-kind tables in SHARED memory, see sb_fill_kind_tables -- lanes look up DIFFERENT kinds, and
// a constant-bank load with a divergent index is replayed once per distinct address, a shared load is not.
SB_HD int sb_column_rule(const SbDev& d, const uint8_t* genv, uint32_t kind, uint32_t rule, int off,
                         const uint32_t* ktab = nullptr) {
  const uint32_t to = ktab ? (ktab[kind] >> 16) & 0xFFu : (uint32_t)d.rule_to[kind];
  if (rule == SB_RULE_BECOME) return (int)to;
  if (rule == SB_RULE_BECOME_IF) {
    const uint32_t other = genv[(ktab ? (int32_t)ktab[64 + kind] : d.rule_plane_off[kind]) + off];
    const uint32_t mask = ktab ? ktab[32 + kind] : d.rule_mask[kind];
    if (other < (uint32_t)SB_MAX_KINDS && ((mask >> other) & 1u)) return (int)to;
  }
  return -1;
}

// Per-kind tables for shared memory, 3 x 32 words: [kind] = channel | rule << 8 | rule_to << 16 | rule_else << 24,
// [32 + kind] = rule_mask, [64 + kind] = rule_plane_off.  32 consecutive words = 32 different banks: conflict-free.
#define SB_KTAB_WORDS 96
SB_HD void sb_fill_kind_tables(const SbDev& d, uint32_t* ktab, int k) {
  ktab[k] = (uint32_t)d.channel[k] | ((uint32_t)d.rule[k] << 8) | ((uint32_t)d.rule_to[k] << 16) |
            ((uint32_t)d.rule_else[k] << 24);
  ktab[32 + k] = d.rule_mask[k];
  ktab[64 + k] = (uint32_t)d.rule_plane_off[k];
}

// Sparse sampling of the spawn field (SB_RNG_SPARSE).  K = number of table entries <= u.
SB_HD int sb_sparse_count(const SbDev& d, uint64_t u) {
  int k = 0;
  while (k < SB_SPARSE_KMAX && u >= d.cdf[k]) ++k;
  return k;
}

// Byte-SIMD first-level FILTER for 4 cells (one 32-bit word of kinds / random bytes): returns 0x80 in
// every byte whose kind == single_spawn_kind and whose random byte r8 < n2 (`hmask4` = ~(n2-1) replicated): exactly
// the cells that pass level 1 of the Bernoulli draw above.  One zero-byte detection on (kinds ^ spawn4) | (rnd & hmask4).
SB_HD uint32_t sb_simd_cand_word(uint32_t kinds, uint32_t rnd, uint32_t spawn4, uint32_t hmask4) {
  uint32_t x = (kinds ^ spawn4) | (rnd & hmask4);
  uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | x | 0x7F7F7F7Fu);  // 0x80 where the byte of x is zero (exact, no cross-byte carries)
}

// Candidate mask of a 16-cell chunk: bit (8*b + w) set <=> byte b of word w is a candidate.
SB_HD uint32_t sb_simd_cand_chunk(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, uint32_t r0, uint32_t r1,
                                  uint32_t r2, uint32_t r3, uint32_t spawn4, uint32_t hmask4) {
  return (sb_simd_cand_word(k0, r0, spawn4, hmask4) >> 7) | (sb_simd_cand_word(k1, r1, spawn4, hmask4) >> 6) |
         (sb_simd_cand_word(k2, r2, spawn4, hmask4) >> 5) | (sb_simd_cand_word(k3, r3, spawn4, hmask4) >> 4);
}

// ------------------------------------------------------------------------------------------------
// Agent move: MovingAgent.act agents/agent.py:215-225 + Gridworld.move gridworld.py:95-122.
// `genv` is the (shared-memory) image of the instance.  Returns the reward; updates *yx (packed
// y<<16|x) and, when `writer`, the two grid cells.  `err` collects SB_ERR_* bits.
// ------------------------------------------------------------------------------------------------
// Cell index the action of an agent at packed position `yx` targets on the agent plane, or -1 (bad action /
// off the grid).  Lets a caller start the load of the target kind early (planes in global memory).
SB_HD int sb_target_cell(const SbDev& d, uint32_t yx, int action) {
  if ((unsigned)action >= (unsigned)d.n_actions) return -1;
  const int ny = (int)(yx >> 16) + (int)(int8_t)((d.dy_packed >> (8 * action)) & 0xFF);
  const int nx = (int)(yx & 0xFFFFu) + (int)(int8_t)((d.dx_packed >> (8 * action)) & 0xFF);
  if ((unsigned)ny >= (unsigned)d.H || (unsigned)nx >= (unsigned)d.W) return -1;  // (no prefetch across the seam)
  return ny * d.W + nx;
}

// `pre_tk` >= 0: kind of the target cell already loaded by the caller (same cell, same state).
// `reward64` != nullptr (worlds with use_value64 / all-layer rewards): receives the reward as the double the
// reference adds to world.total_reward; the returned float is its float32 rounding (Buffer.rewards).
// `own_kind` >= 0: the byte the moving agent is drawn with (worlds where an agent has two kinds); else agent_kind.
SB_HD float sb_agent_act(const SbDev& d, uint8_t* genv, uint32_t* yx, int action, bool writer,
                         uint32_t& err, bool all_layers = false, int pre_tk = -1, double* reward64 = nullptr,
                         int own_kind = -1) {
  int y = (int)(*yx >> 16), x = (int)(*yx & 0xFFFFu);
  int dy = 0, dx = 0;
  if ((unsigned)action < (unsigned)d.n_actions) {
    dy = (int)(int8_t)((d.dy_packed >> (8 * action)) & 0xFF);
    dx = (int)(int8_t)((d.dx_packed >> (8 * action)) & 0xFF);
  } else {
    err |= SB_ERR_BAD_ACTION;
  }
  int ny = y + dy, nx = x + dx;
  if ((unsigned)ny >= (unsigned)d.H || (unsigned)nx >= (unsigned)d.W) {
    if (!d.toroidal) {  // numpy index wrap / IndexError in the reference: reported, not emulated
      err |= SB_ERR_OFFGRID;
      return 0.0f;
    }
    ny = ny < 0 ? ny + d.H : (ny >= d.H ? ny - d.H : ny);  // explicit toroidal mode: moves cross the seam
    nx = nx < 0 ? nx + d.W : (nx >= d.W ? nx - d.W : nx);
    pre_tk = -1;
  }
  uint8_t* plane = genv + d.agent_plane * d.plane_stride;
  uint32_t tk = pre_tk >= 0 ? (uint32_t)pre_tk : plane[ny * d.W + nx];
  if (tk >= (uint32_t)d.K) {
    err |= SB_ERR_BAD_KIND;
    return 0.0f;
  }
  float reward = d.value[tk];  // reward = value of the entity at the target cell (agent.py:219-220)
  if (reward64) *reward64 = d.value64[tk];
  if (all_layers) {  // sum over every layer of the target column (observe_all_layers, cleanup/agents.py:160-170)
    // float32-exact values (cleanup: 0, 1, -1) are summed in float; the double path is taken only when asked for
    double r64 = 0.0;
    reward = 0.0f;
    for (int z = 0; z < d.L; ++z) {
      const int pl = d.plane_of_layer[z];
      const uint32_t k = pl >= 0 ? genv[pl * d.plane_stride + ny * d.W + nx] : d.tmpl[z * d.plane_stride + ny * d.W + nx];
      if (k >= (uint32_t)d.K) err |= SB_ERR_BAD_KIND;
      else if (reward64) r64 += d.value64[k];
      else reward += d.value[k];
    }
    if (reward64) {
      reward = (float)r64;
      *reward64 = r64;
    }
  } else if (reward64) {
    reward = (float)*reward64;  // Buffer.rewards is float32 (buffers.py:33,58)
  }
  if ((d.passable_mask >> tk) & 1u) {
    if (writer) {
      plane[ny * d.W + nx] = (uint8_t)(own_kind >= 0 ? own_kind : d.agent_kind);  // target discarded, agent placed
      plane[y * d.W + x] = (uint8_t)d.default_kind;  // vacated cell <- default entity
    }
    *yx = ((uint32_t)ny << 16) | (uint32_t)nx;
  }
  return reward;
}

// TagAgent.act (examples/tag/agents.py:84-105) after the move: an agent drawn with agent_kind_alt ("It") hands the role
// to the first neighbour drawn with agent_kind ("NotIt") in the order of Location.adjacent (location.py:144): up,
// right, down, left.  The state lives in the grid byte of the agents.  Returns the agent's kind afterwards.
SB_HD uint32_t sb_touch_swap(const SbDev& d, uint8_t* genv, uint32_t yx) {
  uint8_t* plane = genv + d.agent_plane * d.plane_stride;
  const int y = (int)(yx >> 16), x = (int)(yx & 0xFFFFu);
  const int cur = y * d.W + x;
  if (plane[cur] == (uint8_t)d.agent_kind_alt) {
    const int oy[4] = {-1, 0, 1, 0}, ox[4] = {0, 1, 0, -1};
    for (int k = 0; k < 4; ++k) {
      const int qy = y + oy[k], qx = x + ox[k];
      if ((unsigned)qy >= (unsigned)d.H || (unsigned)qx >= (unsigned)d.W) continue;
      if (plane[qy * d.W + qx] == (uint8_t)d.agent_kind) {
        plane[cur] = (uint8_t)d.agent_kind;
        plane[qy * d.W + qx] = (uint8_t)d.agent_kind_alt;
        break;
      }
    }
  }
  return plane[cur];
}

// ------------------------------------------------------------------------------------------------
// Observation, general path: closed form of visual_field (observation/visual_field.py:9-101):
// element (c, i, j) of the crop centred on (y, x).  float64 sum over layers in z order, like
// np.sum(axis=-1) on a short axis; the caller rounds to float32 (Buffer.states, buffers.py:31,56).
// ------------------------------------------------------------------------------------------------
SB_HD double sb_obs_general(const SbDev& d, const uint8_t* genv, const uint8_t* tmpl, const double* app,
                            int y, int x, int c, int i, int j) {
  int yy = y - d.v + i, xx = x - d.v + j;
  double s = 0.0;
  const bool outside = (unsigned)yy >= (unsigned)d.H || (unsigned)xx >= (unsigned)d.W;
  if (outside && d.toroidal) {  // |offset| <= v <= (H-1)/2, so one fold is enough
    yy = yy < 0 ? yy + d.H : (yy >= d.H ? yy - d.H : yy);
    xx = xx < 0 ? xx + d.W : (xx >= d.W ? xx - d.W : xx);
  }
  if (outside && !d.toroidal) {
    s = app[d.fill_kind * d.C + c];
  } else {
    int cell = yy * d.W + xx;
    for (int z = 0; z < d.L; ++z) {
      int pl = d.plane_of_layer[z];
      uint32_t kind = pl >= 0 ? genv[pl * d.plane_stride + cell] : tmpl[z * d.plane_stride + cell];
      s += app[kind * d.C + c];
    }
  }
  if (d.obs_transform == SB_OBS_RGB) s = (s < 0.0 ? 0.0 : (s > 255.0 ? 255.0 : s)) / 255.0;  // np.clip(..)/255
  return s;
}

// One-hot channel of a kind (255 = all-zero appearance row / padding byte).
SB_HD uint32_t sb_kind_channel(const SbDev& d, uint32_t kind) {
  if (d.K <= 8) {  // 8-entry table in a register pair, one byte-permute
#if defined(__CUDA_ARCH__)
    return __byte_perm((uint32_t)d.channel8, (uint32_t)(d.channel8 >> 32), kind & 7u) & 0xFFu;
#else
    return (uint32_t)(d.channel8 >> (8 * (kind & 7u))) & 0xFFu;
#endif
  }
  return kind < (uint32_t)SB_MAX_KINDS ? d.channel[kind] : 255u;
}

// Fast path, phase A: channel code of window cell (i, j) on plane `pl` (255 = contributes nothing).
SB_HD uint32_t sb_window_code(const SbDev& d, const uint8_t* genv, int y, int x, int i, int j, int pl) {
  int yy = y - d.v + i, xx = x - d.v + j;
  if ((unsigned)yy >= (unsigned)d.H || (unsigned)xx >= (unsigned)d.W) {
    if (!d.toroidal) return pl == 0 ? (uint32_t)d.fill_channel : 255u;  // out of bounds: the fill kind, once
    // toroidal worlds: the window shows the opposite side (|offset| <= v <= (H-1)/2, so one fold is enough)
    yy = yy < 0 ? yy + d.H : (yy >= d.H ? yy - d.H : yy);
    xx = xx < 0 ? xx + d.W : (xx >= d.W ? xx - d.W : xx);
  }
  return sb_kind_channel(d, genv[pl * d.plane_stride + yy * d.W + xx]);
}
