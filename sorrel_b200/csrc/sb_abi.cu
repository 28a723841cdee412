// sb_abi.cu -- extern "C" entry points of libsorrel_b200.so (see include/sorrel_b200.h):
// argument validation, derivation of the device-side world view, and the small kernels that do
// not live in sb_step.cu (reset, replay-ring add / sample / current_state, reward statistics).
#include <atomic>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "sb_internal.h"

namespace sb {

static thread_local char g_cuda_err[256] = "";
static std::atomic<uint64_t> g_launches{0};

int note_cuda_error(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s", cudaGetErrorString(e));
  return SB_ECUDA;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ceil(p * 2^53): `u < p` for u = k / 2^53 (numpy random_sample) <=> k < threshold
static uint64_t spawn_threshold(double p) {
  if (!(p > 0.0)) return 0;
  if (p >= 1.0) return 1ull << 53;
  return (uint64_t)ceil(ldexp(p, 53));
}

int make_dev(const SbWorldDesc* w, SbDev* out) {
  if (!w || !out) return SB_EINVAL;
  if (w->abi_version != SB_ABI_VERSION) return SB_EINVAL;
  if (w->num_envs <= 0 || w->num_agents <= 0 || w->height <= 0 || w->width <= 0) return SB_EINVAL;
  if (w->num_layers <= 0 || w->num_layers > SB_MAX_LAYERS) return SB_EUNSUPPORTED;
  if (w->num_dyn_layers <= 0 || w->num_dyn_layers > w->num_layers) return SB_EINVAL;
  if (w->num_agents > SB_MAX_AGENTS) return SB_EUNSUPPORTED;
  if (w->num_kinds <= 0 || w->num_kinds > SB_MAX_KINDS) return SB_EUNSUPPORTED;
  if (w->num_channels <= 0 || w->num_channels > SB_MAX_CHANNELS) return SB_EUNSUPPORTED;
  if (w->height > 4096 || w->width > 4096) return SB_EUNSUPPORTED;
  if (w->vision < 0) return SB_EINVAL;
  // visual_field.py:78-87 returns a wrong-shaped crop beyond this
  if (w->vision > (w->height - 1) / 2 || w->vision > (w->width - 1) / 2) return SB_EVISION;
  if (w->vision > 10) return SB_EUNSUPPORTED;  // window index is 9 bits in the kernel LUT
  if (w->num_actions <= 0 || w->num_actions > SB_MAX_ACTIONS) return SB_EUNSUPPORTED;
  if (w->num_spawn_choices < 0 || w->num_spawn_choices > SB_MAX_CHOICES) return SB_EUNSUPPORTED;
  auto kind_ok = [&](int k) { return k >= 0 && k < w->num_kinds; };
  if (!kind_ok(w->fill_kind) || !kind_ok(w->agent_kind) || !kind_ok(w->default_kind)) return SB_EINVAL;
  if (!(w->spawn_prob >= 0.0) || w->spawn_prob > 1.0) return SB_EINVAL;
  if (!w->d_appearance || !w->h_appearance || !w->d_template) return SB_EINVAL;
  if (w->env_id_base + (uint64_t)w->num_envs > 0xFFFFFFFFull) return SB_EUNSUPPORTED;

  SbDev d;
  memset(&d, 0, sizeof(d));
  d.N = w->num_envs;
  d.A = w->num_agents;
  d.H = w->height;
  d.W = w->width;
  d.HW = d.H * d.W;
  d.plane_stride = sb_plane_stride(d.H, d.W);
  d.n_planes = w->num_dyn_layers;
  d.env_stride = d.n_planes * d.plane_stride;
  if (d.env_stride >= (1 << 20)) return SB_EUNSUPPORTED;  // cell id is 20 bits in the spawn queue
  d.n_chunks = d.env_stride / 16;
  d.L = w->num_layers;
  for (int z = 0; z < SB_MAX_LAYERS; ++z) d.plane_of_layer[z] = -1;
  d.agent_plane = -1;
  int prev = -1;
  for (int pl = 0; pl < d.n_planes; ++pl) {
    int z = w->dyn_layers[pl];
    if (z <= prev || z >= d.L) return SB_EINVAL;
    prev = z;
    d.plane_of_layer[z] = pl;
    if (z == w->agent_layer) d.agent_plane = pl;
  }
  if (d.agent_plane < 0) return SB_EINVAL;
  d.K = w->num_kinds;
  d.C = w->num_channels;
  d.v = w->vision;
  d.S = 2 * d.v + 1;
  d.SS = d.S * d.S;
  d.oh_size = d.C * d.SS;
  if (w->obs_extra < 0 || w->obs_extra > 256 || (w->obs_extra > 0 && !w->d_obs_extra)) return SB_EINVAL;
  d.obs_extra = w->obs_extra;
  d.extra_tbl = w->d_obs_extra;
  d.obs_flag = w->obs_flag ? 1 : 0;
  d.obs_size = d.oh_size + d.obs_extra + d.obs_flag;
  d.inv_S = (65536u + (uint32_t)d.S - 1u) / (uint32_t)d.S;
  d.inv_W = ((uint64_t)d.HW * (uint64_t)d.W < (1ull << 32) && d.W > 1)
                ? (uint32_t)(((1ull << 32) + (uint64_t)d.W - 1) / (uint64_t)d.W) : 0u;
  d.fill_kind = w->fill_kind;
  d.agent_kind = w->agent_kind;
  d.default_kind = w->default_kind;
  d.n_actions = w->num_actions;
  d.n_choices = w->num_spawn_choices;
  d.env_id_base = w->env_id_base;
  d.p = w->spawn_prob;
  const uint64_t thr = spawn_threshold(d.p);
  d.n2 = sb_level1_bound(thr);
  d.thr64m1 = sb_level2_bound(thr, d.n2);
  int n_spawn_kinds = 0;
  d.single_spawn_kind = -1;
  for (int k = 0; k < d.K; ++k) {
    d.value[k] = w->kind_value[k];
    if (w->kind_passable[k]) d.passable_mask |= 1u << k;
    const int rule = w->kind_rule[k];
    d.rule[k] = (uint8_t)rule;
    d.rule_to[k] = w->kind_rule_to[k];
    d.rule_else[k] = w->kind_rule_else[k];
    d.value64[k] = w->use_value64 ? w->kind_value64[k] : (double)w->kind_value[k];
    if (rule == SB_RULE_SPAWN) {
      d.spawn_mask |= 1u << k;
      d.draw_mask |= 1u << k;
      d.single_spawn_kind = k;
      d.prob[k] = w->spawn_prob;
      d.n2k[k] = (uint16_t)d.n2;
      d.thr64m1k[k] = d.thr64m1;
      ++n_spawn_kinds;
    } else if (rule == SB_RULE_SPAWN_TO) {
      if (!(w->kind_rule_prob[k] >= 0.0) || w->kind_rule_prob[k] > 1.0 || !kind_ok(w->kind_rule_to[k])) return SB_EINVAL;
      if (w->kind_rule_else[k] != SB_NO_KIND) {
        if (!kind_ok(w->kind_rule_else[k])) return SB_EINVAL;
        d.else_mask |= 1u << k;
      }
      d.draw_mask |= 1u << k;
      d.prob[k] = w->kind_rule_prob[k];
      d.n2k[k] = (uint16_t)sb_level1_bound(spawn_threshold(d.prob[k]));
      d.thr64m1k[k] = sb_level2_bound(spawn_threshold(d.prob[k]), d.n2k[k]);
      d.ext_rules = 1;
    } else if (rule == SB_RULE_BECOME_IF) {
      const int z = w->kind_rule_layer[k];
      if (z < 0 || z >= d.L || d.plane_of_layer[z] < 0 || !kind_ok(w->kind_rule_to[k])) return SB_EINVAL;
      d.rule_plane_off[k] = d.plane_of_layer[z] * d.plane_stride;
      d.rule_mask[k] = w->kind_rule_mask[k];
      d.det_mask |= 1u << k;
      d.ext_rules = 1;
    } else if (rule == SB_RULE_BECOME) {
      if (!kind_ok(w->kind_rule_to[k])) return SB_EINVAL;
      d.det_mask |= 1u << k;
      d.ext_rules = 1;
    } else if (rule != SB_RULE_NONE) {
      return SB_EUNSUPPORTED;
    }
  }
  if (n_spawn_kinds != 1 || d.ext_rules) d.single_spawn_kind = -1;
  // facings / beams / all-layer rewards (MovingAgent.movement, CleanupAgent.spawn_beam / act)
  d.reward_all = w->reward_all_layers ? 1 : 0;
  d.reward_count = w->reward_count > 0 ? w->reward_count : 1;
  d.beam_plane = -1;
  for (int i = 0; i < SB_MAX_ACTIONS; ++i) {
    d.act_facing[i] = -1;
    d.act_beam[i] = SB_NO_KIND;
  }
  if (w->uses_facing) {
    bool any_beam = false;
    for (int i = 0; i < w->num_actions; ++i) {
      if (w->action_facing[i] > 3) return SB_EINVAL;
      d.act_facing[i] = w->action_facing[i] < 0 ? -1 : w->action_facing[i];
      d.act_beam[i] = w->action_beam_kind[i];
      if (d.act_beam[i] != SB_NO_KIND) {
        if (!kind_ok(d.act_beam[i])) return SB_EINVAL;
        any_beam = true;
      }
    }
    if (any_beam) {
      if (w->beam_layer < 0 || w->beam_layer >= d.L || d.plane_of_layer[w->beam_layer] < 0) return SB_EINVAL;
      if (w->num_beam_cells < 0 || w->num_beam_cells > SB_MAX_BEAM_CELLS) return SB_EUNSUPPORTED;
      d.beam_plane = d.plane_of_layer[w->beam_layer];
      d.n_beam = w->num_beam_cells;
      memcpy(d.beam_dy, w->beam_dy, sizeof(d.beam_dy));
      memcpy(d.beam_dx, w->beam_dx, sizeof(d.beam_dx));
      d.beam_block_mask = w->beam_block_mask;
    }
  }
  d.use_value64 = w->use_value64 ? 1 : 0;
  d.agent_kind_alt = -1;
  if (w->agent_kind_alt >= 0) {
    if (!kind_ok(w->agent_kind_alt) || w->agent_kind_alt == w->agent_kind || !w->d_agent_flag) return SB_EINVAL;
    d.agent_kind_alt = w->agent_kind_alt;
    d.agent_flag = w->d_agent_flag;
  } else if (w->touch_swap || w->obs_flag) {
    return SB_EINVAL;
  }
  d.touch_swap = w->touch_swap ? 1 : 0;
  d.reward_own = w->reward_own ? 1 : 0;
  d.ext_act = (w->uses_facing || d.reward_all || d.reward_count != 1 || d.use_value64 || d.agent_kind_alt >= 0 ||
               d.reward_own) ? 1 : 0;
  if (n_spawn_kinds > 0 && d.n_choices == 0 && d.p > 0.0) return SB_EINVAL;
  for (int i = 0; i < d.n_choices; ++i) {
    if (!kind_ok(w->spawn_choice_kinds[i])) return SB_EINVAL;
    d.choices_packed |= (uint64_t)w->spawn_choice_kinds[i] << (8 * i);
  }
  int max_delta = 0;
  for (int i = 0; i < d.n_actions; ++i) {
    d.dy_packed |= (uint64_t)(uint8_t)w->action_dy[i] << (8 * i);
    d.dx_packed |= (uint64_t)(uint8_t)w->action_dx[i] << (8 * i);
    max_delta = max(max_delta, max(abs((int)w->action_dy[i]), abs((int)w->action_dx[i])));
  }
  // two agents farther apart than this cannot read or write a cell the other's move touches this turn: a move
  // changes cells within max_delta of the agent, a window reaches v, two targets are at most 2*max_delta apart
  d.par_radius = max(d.v + max_delta, 2 * max_delta);
  const bool ext = d.ext_rules || d.ext_act || d.obs_extra > 0 || d.obs_flag;  // only the generic / replay variants implement these
  d.simd_sweep = (d.single_spawn_kind >= 0 && d.n2 >= 1u && d.n2 <= 128u && !ext) ? 1 : 0;
  d.hmask4 = ((~(d.n2 - 1u)) & 0xFFu) * 0x01010101u;
  // one-hot fast path: every appearance row is all-zero or exactly one 1.0
  if (w->obs_transform != SB_OBS_SUM && w->obs_transform != SB_OBS_RGB) return SB_EINVAL;
  d.obs_transform = w->obs_transform;
  d.toroidal = w->toroidal ? 1 : 0;
  // (toroidal worlds keep the fast path: window indices are folded where the cell code is looked up)
  d.fast_obs = (w->static_layers_invisible || d.n_planes == d.L) && d.obs_transform == SB_OBS_SUM ? 1 : 0;
  for (int k = 0; k < d.K; ++k) {
    int ch = 255, ones = 0;
    for (int c = 0; c < d.C; ++c) {
      double v = w->h_appearance[k * d.C + c];
      if (v == 1.0) {
        ch = c;
        ++ones;
      } else if (v != 0.0) {
        d.fast_obs = 0;
      }
    }
    if (ones > 1) d.fast_obs = 0;
    d.channel[k] = (uint8_t)ch;
  }
  for (int k = d.K; k < SB_MAX_KINDS; ++k) d.channel[k] = 255;
  for (int k = 0; k < 8; ++k) d.channel8 |= (uint64_t)d.channel[k] << (8 * k);
  d.fill_channel = d.channel[d.fill_kind];
  d.S5 = d.S - d.S % 5;
  d.rowwise_ok = d.K <= 7 && !d.toroidal ? 1 : 0;  // (a toroidal window row is not one run of bytes)
  d.channel8_oob = (d.channel8 & 0x00FFFFFFFFFFFFFFull) | ((uint64_t)(uint8_t)d.fill_channel << 56);
  // class bytes of the generic sweep: A = ~(n2 - 1) of a drawing kind (level 1 passes iff (r8 & A) == 0); O bit 0 =
  // the kind does not draw (or never passes: p == 0), bit 7 = visit always (deterministic rule; a draw whose failure
  // still changes the kind)
  d.cls_ok = d.K <= 15 ? 1 : 0;
  for (int k = 0; k < 16; ++k) {
    uint32_t a = 0xFFu, o = 0x01u;
    if (k < d.K) {
      if (((d.draw_mask >> k) & 1u) && d.n2k[k] >= 1) a = (~((uint32_t)d.n2k[k] - 1u)) & 0xFFu, o = 0u;
      if (((d.det_mask >> k) & 1u) || ((d.else_mask >> k) & 1u)) o |= 0x80u;
    }
    d.cls_and[k >> 3] |= (uint64_t)a << (8 * (k & 7));
    d.cls_or[k >> 3] |= (uint64_t)o << (8 * (k & 7));
  }
  for (int k = 8; k < 16; ++k) d.channel8_hi |= (uint64_t)d.channel[k] << (8 * (k - 8));
  d.colgather_ok = (d.fast_obs && d.n_planes >= 2 && d.n_planes <= 3 && d.K <= 16 && !d.toroidal) ? 1 : 0;
  for (int lg = 0; lg < 6; ++lg) {
    // segments per window row for groups of 2^lg lanes: a lane walks its segment four cells at a time, so a window
    // costs ceil(S * nseg / lanes) trips x ceil(len / 4) chunk steps; take the cheapest split (at most 8 cells per
    // segment; ties: fewer segments).  11x11 windows: 3 segments (4, 4, 3 cells) on 16 lanes = 3 steps, not 4.
    const int lanes = 1 << lg;
    int best = 0, best_cost = 1 << 30;
    for (int nseg = (d.S + 7) / 8; nseg <= d.S; ++nseg) {
      const int len = (d.S + nseg - 1) / nseg;
      if ((nseg - 1) * len >= d.S) continue;  // the last segment would be empty
      const int cost = ((d.S * nseg + lanes - 1) / lanes) * ((len + 3) / 4);
      if (cost < best_cost) best = nseg, best_cost = cost;
    }
    d.cg_nseg[lg] = (int8_t)best;
    d.cg_len[lg] = (int8_t)((d.S + best - 1) / best);
    // one stored plane (scatter_rows_onehot_seg): five cells per step, any segment length
    best = 1, best_cost = 1 << 30;
    for (int nseg = 1; nseg <= d.S; ++nseg) {
      const int len = (d.S + nseg - 1) / nseg;
      if ((nseg - 1) * len >= d.S) continue;
      const int cost = ((d.S * nseg + lanes - 1) / lanes) * ((len + 4) / 5);
      if (cost < best_cost) best = nseg, best_cost = cost;
    }
    d.rw_nseg[lg] = (int8_t)best;
    d.rw_len[lg] = (int8_t)((d.S + best - 1) / best);
  }
  d.sparse_ok = (w->spawn_cdf_len == SB_SPARSE_KMAX && d.env_stride < 65536 && !ext) ? 1 : 0;
  if (d.sparse_ok) memcpy(d.cdf, w->spawn_cdf, sizeof(d.cdf));
  d.app = w->d_appearance;
  d.tmpl = w->d_template;
  *out = d;
  return SB_OK;
}

// ------------------------------------------------------------------------------------------------
// reset: Environment.reset + populate_environment (environment.py:72-79, treasurehunt/env.py:114-147)
// ------------------------------------------------------------------------------------------------
constexpr int RESET_LPE = 4;
struct ResetGroups {
  int32_t n;
  int32_t kind[SB_MAX_ITEM_GROUPS], plane[SB_MAX_ITEM_GROUPS], count[SB_MAX_ITEM_GROUPS], n_cells[SB_MAX_ITEM_GROUPS];
  const int32_t* cells[SB_MAX_ITEM_GROUPS];
  const int32_t* placement[SB_MAX_ITEM_GROUPS];
};

__global__ void __launch_bounds__(SB_THREADS)
    reset_kernel(const __grid_constant__ SbDev d, uint8_t* __restrict__ grid, int32_t* __restrict__ agent_pos,
                 double* __restrict__ total_reward, const int32_t* __restrict__ spawnable, int n_spawnable,
                 const int32_t* __restrict__ placement, const __grid_constant__ ResetGroups groups, uint64_t seed,
                 uint32_t tick) {
  const int tid = threadIdx.x, lane = tid & (RESET_LPE - 1);
  const long long env = (long long)blockIdx.x * (SB_THREADS / RESET_LPE) + tid / RESET_LPE;
  const bool active = env < d.N;
  if (active) {
    uint4* dst = reinterpret_cast<uint4*>(grid + (size_t)env * d.env_stride);
    const int cpp = d.plane_stride / 16;
    for (int c = lane; c < d.n_chunks; c += RESET_LPE) {
      int pl = c / cpp, off = c - pl * cpp;
      int z = 0;
      for (int zz = 0; zz < d.L; ++zz)
        if (d.plane_of_layer[zz] == pl) z = zz;
      dst[c] = __ldg(reinterpret_cast<const uint4*>(d.tmpl + (size_t)z * d.plane_stride) + off);
    }
  }
  __syncwarp();  // orders the template copy before lane 0's placement writes
  if (active && lane == 0) {
    uint8_t* plane = grid + (size_t)env * d.env_stride + (size_t)d.agent_plane * d.plane_stride;
    const uint32_t envid = (uint32_t)(d.env_id_base + (uint64_t)env);
    const uint32_t sk[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    // item groups first, in order (cleanup/env.py:126-132: apples before agents)
    for (int g = 0; g < groups.n; ++g) {
      uint8_t* gp = grid + (size_t)env * d.env_stride + (size_t)groups.plane[g] * d.plane_stride;
      uint32_t jg = (uint32_t)(g + 1) << 24;
      for (int i = 0; i < groups.count[g]; ++i) {
        int cell;
        if (groups.placement[g]) {
          cell = groups.cells[g][groups.placement[g][(size_t)env * groups.count[g] + i]];
        } else {
          for (;;) {
            SbU4 w = sb_philox_rk(envid, tick, jg++, SB_DOM_RESET, sk);
            cell = groups.cells[g][sb_bounded(w.x, w.y, (uint32_t)groups.n_cells[g])];
            if (gp[cell] != (uint8_t)groups.kind[g]) break;
          }
        }
        gp[cell] = (uint8_t)groups.kind[g];
      }
    }
    uint32_t j = 0;
    for (int a = 0; a < d.A; ++a) {
      int cell;
      if (placement) {
        cell = spawnable[placement[(size_t)env * d.A + a]];
      } else {
        // sequential rejection: uniform over the spawnable cells not taken yet
        for (;;) {
          SbU4 w = sb_philox_rk(envid, tick, j++, SB_DOM_RESET, sk);
          cell = spawnable[sb_bounded(w.x, w.y, (uint32_t)n_spawnable)];
          if (plane[cell] != (uint8_t)d.agent_kind && (d.agent_kind_alt < 0 || plane[cell] != (uint8_t)d.agent_kind_alt))
            break;
        }
      }
      // world.add(loc, agent) overwrites, gridworld.py:67-76; a flagged agent is drawn with its other kind
      const bool flagged = d.agent_kind_alt >= 0 && d.agent_flag[(size_t)env * d.A + a] != 0;
      plane[cell] = (uint8_t)(flagged ? d.agent_kind_alt : d.agent_kind);
      agent_pos[(size_t)env * d.A + a] = cell;
    }
    total_reward[env] = 0.0;  // create_world, gridworld.py:65
  }
}

// ------------------------------------------------------------------------------------------------
// replay ring
// ------------------------------------------------------------------------------------------------
__global__ void buffer_add_kernel(SbBufferDesc b, int slot, const float* __restrict__ obs,
                                  const int64_t* __restrict__ actions, const float* __restrict__ rewards,
                                  const float* __restrict__ dones, const int64_t* __restrict__ positions) {
  // one warp per row: states[row][slot][:] = obs[row][:]  (Buffer.add, buffers.py:56-62)
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= b.rows) return;
  const float* src = obs + (size_t)row * b.obs_size;
  float* dst = b.d_states + (size_t)row * b.states_row_stride + (size_t)slot * b.states_slot_stride;
  for (int e = lane; e < b.obs_size; e += 32) dst[e] = src[e];
  if (lane == 0) {
    const size_t at = (size_t)slot * b.rows + row;
    b.d_actions[at] = actions[row];
    b.d_rewards[at] = rewards[row];
    b.d_dones[at] = dones[row];
    if (b.d_positions && positions) {
      b.d_positions[2 * at] = positions[2 * row];
      b.d_positions[2 * at + 1] = positions[2 * row + 1];
    }
  }
}

// dst_a[e] = src[e] for e < hi_a and dst_b[e - lo_b] = src[e] for e >= lo_b, e in [0, total): one pass over the source
// run feeding two (overlapping) destinations.  SB_COPY_DEPTH loads per lane are issued before their stores, so a
// warp keeps several KB in flight (the plain load-store loop left the gather latency-bound at ~3.6 TB/s).
#ifndef SB_COPY_DEPTH
#define SB_COPY_DEPTH 4  // measured at config 5: 4 -> 0.89 ms, 8 -> 0.92 ms, 16 -> 1.37 ms
#endif
template <typename V>
__device__ __forceinline__ void copy_run2(const V* __restrict__ src, V* __restrict__ dst_a, V* __restrict__ dst_b,
                                          int hi_a, int lo_b, int total, int lane) {
  for (int e0 = lane; e0 < total; e0 += 32 * SB_COPY_DEPTH) {
    V v[SB_COPY_DEPTH];
#pragma unroll
    for (int u = 0; u < SB_COPY_DEPTH; ++u)
      if (e0 + 32 * u < total) v[u] = __ldg(src + e0 + 32 * u);
#pragma unroll
    for (int u = 0; u < SB_COPY_DEPTH; ++u) {
      const int e = e0 + 32 * u;
      if (e < total) {
        if (e < hi_a) dst_a[e] = v[u];
        if (e >= lo_b) dst_b[e - lo_b] = v[u];
      }
    }
  }
}

__global__ void buffer_sample_kernel(SbBufferDesc b, const int64_t* __restrict__ rows_idx,
                                     const int64_t* __restrict__ starts, int batch, float* __restrict__ states,
                                     int64_t* __restrict__ actions, float* __restrict__ rewards,
                                     float* __restrict__ next_states, float* __restrict__ dones,
                                     double* __restrict__ valid, int64_t* __restrict__ action_seq,
                                     float* __restrict__ next_action_seq) {
  // one warp per sample (Buffer.sample, buffers.py:112-124): the F stacked frames of a row are
  // contiguous in the ring, so states / next_states are two overlapping linear copies.
  const int sample = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (sample >= batch) return;
  const int64_t row = rows_idx[sample], s0 = starts[sample];
  const int F = b.n_frames;
  // numpy raises IndexError for these (buffers.py:112-119); report instead of reading a neighbouring row
  if (row < 0 || row >= b.rows || s0 < 0 || s0 + F >= b.capacity) {
    if (lane == 0 && b.d_error) atomicOr(b.d_error, SB_ERR_BAD_INDEX);
    return;
  }
  const size_t n = (size_t)F * b.obs_size;
  const float* src = b.d_states + (size_t)row * b.states_row_stride + (size_t)s0 * b.states_slot_stride;
  float* ds = states + (size_t)sample * n;
  float* dn = next_states + (size_t)sample * n;
  // 8-byte accesses when everything is 8-byte aligned (obs_size and the strides even), else scalar
  const bool wide = (b.obs_size & 1) == 0 && (b.states_slot_stride & 1) == 0 &&
                    ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(ds) |
                      reinterpret_cast<uintptr_t>(dn)) & 7u) == 0;
  if (b.states_slot_stride == b.obs_size) {
    // frames of a row are contiguous (agent-major ring): states / next_states are two overlapping linear copies of
    // (F + 1) * obs floats
    if (wide)
      copy_run2(reinterpret_cast<const float2*>(src), reinterpret_cast<float2*>(ds), reinterpret_cast<float2*>(dn),
                (int)(n / 2), b.obs_size / 2, (int)(n / 2) + b.obs_size / 2, lane);
    else
      copy_run2(src, ds, dn, (int)n, b.obs_size, (int)n + b.obs_size, lane);
  } else {
    // any other layout (slot-major ring): frame f of the transition is its own run of obs floats; frame f goes to
    // states[f] (f < F) and next_states[f - 1] (f > 0)
    for (int f = 0; f <= F; ++f) {
      const float* sf = src + (size_t)f * b.states_slot_stride;
      float* to_s = ds + (size_t)f * b.obs_size;
      float* to_n = dn + (size_t)f * b.obs_size - b.obs_size;
      // copy_run2 writes to_s[e] for e < hi_s and to_n'[e - o] for e >= o: one frame, both or one destination
      const int len = wide ? b.obs_size / 2 : b.obs_size;
      const int hi_s = f < F ? len : 0, lo_n = f > 0 ? 0 : len;
      if (wide)
        copy_run2(reinterpret_cast<const float2*>(sf), reinterpret_cast<float2*>(to_s),
                  reinterpret_cast<float2*>(to_n), hi_s, lo_n, len, lane);
      else
        copy_run2(sf, to_s, to_n, hi_s, lo_n, len, lane);
    }
  }
  if (action_seq && lane < F) {  // sample_transformer: actions[indices], actions[indices + 1] as f32
    action_seq[(size_t)sample * F + lane] = b.d_actions[(size_t)(s0 + lane) * b.rows + row];
    next_action_seq[(size_t)sample * F + lane] = (float)b.d_actions[(size_t)(s0 + lane + 1) * b.rows + row];
  }
  if (lane == 0) {
    const size_t last = (size_t)(s0 + F - 1) * b.rows + row;  // indices[:, -1]
    if (actions) actions[sample] = b.d_actions[last];
    if (rewards) rewards[sample] = b.d_rewards[last];
    dones[sample] = b.d_dones[last];
    bool any = false;  // valid = 1 - any(dones[indices[:, :-1]]), buffers.py:120-122
    for (int f = 0; f < F - 1; ++f) any = any || (b.d_dones[(size_t)(s0 + f) * b.rows + row] != 0.0f);
    valid[sample] = any ? 0.0 : 1.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Buffer.sample gather, STAGED form: the F + 1 frames of a transition are pulled into shared memory with 8-byte
// asynchronous copies (no registers held while the loads are in flight: a warp keeps its whole 3.6 KB run
// outstanding instead of 1 KB, and ~50 warps per SM are resident), then written out twice -- states = frames
// [0, F), next_states = frames [1, F] -- from shared memory.  The register-staged loop above was latency-bound
// (ncu: 0.21 eligible warps per cycle, DRAM 61 % busy).  Needs 8-byte aligned runs (even obs_size / strides).
// ------------------------------------------------------------------------------------------------
#ifndef SB_SAMPLE_ASYNC
#define SB_SAMPLE_ASYNC 1
#endif
constexpr int SAMPLE_WARPS = 4;  // samples per block (one warp each)
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__global__ void __launch_bounds__(SAMPLE_WARPS * 32)
    buffer_sample_staged_kernel(SbBufferDesc b, const int64_t* __restrict__ rows_idx,
                                const int64_t* __restrict__ starts, int batch, float* __restrict__ states,
                                int64_t* __restrict__ actions, float* __restrict__ rewards,
                                float* __restrict__ next_states, float* __restrict__ dones,
                                double* __restrict__ valid, int64_t* __restrict__ action_seq,
                                float* __restrict__ next_action_seq) {
  extern __shared__ __align__(16) float2 sample_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sample = blockIdx.x * SAMPLE_WARPS + warp;
  if (sample >= batch) return;  // (warp-level synchronisation only below)
  const int64_t row = rows_idx[sample], s0 = starts[sample];
  const int F = b.n_frames;
  if (row < 0 || row >= b.rows || s0 < 0 || s0 + F >= b.capacity) {  // IndexError in numpy (buffers.py:112-119)
    if (lane == 0 && b.d_error) atomicOr(b.d_error, SB_ERR_BAD_INDEX);
    return;
  }
  const int obs2 = b.obs_size / 2, n2 = F * obs2;  // float2 units
  float2* buf = sample_smem + (size_t)warp * (size_t)(n2 + obs2);
  const float* src = b.d_states + (size_t)row * b.states_row_stride + (size_t)s0 * b.states_slot_stride;
  if (b.states_slot_stride == b.obs_size) {  // agent-major ring: one run of (F + 1) * obs floats
    const float2* s2 = reinterpret_cast<const float2*>(src);
    for (int e = lane; e < n2 + obs2; e += 32) cp_async8(buf + e, s2 + e);
  } else {  // slot-major ring: F + 1 runs of obs floats
    for (int f = 0; f <= F; ++f) {
      const float2* s2 = reinterpret_cast<const float2*>(src + (size_t)f * b.states_slot_stride);
      for (int e = lane; e < obs2; e += 32) cp_async8(buf + f * obs2 + e, s2 + e);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // the scalar part of the sample while the copies are in flight
  if (action_seq && lane < F) {  // sample_transformer: actions[indices], actions[indices + 1] as f32
    action_seq[(size_t)sample * F + lane] = b.d_actions[(size_t)(s0 + lane) * b.rows + row];
    next_action_seq[(size_t)sample * F + lane] = (float)b.d_actions[(size_t)(s0 + lane + 1) * b.rows + row];
  }
  if (lane == 0) {
    const size_t last = (size_t)(s0 + F - 1) * b.rows + row;  // indices[:, -1]
    if (actions) actions[sample] = b.d_actions[last];
    if (rewards) rewards[sample] = b.d_rewards[last];
    dones[sample] = b.d_dones[last];
    bool any = false;  // valid = 1 - any(dones[indices[:, :-1]]), buffers.py:120-122
    for (int f = 0; f < F - 1; ++f) any = any || (b.d_dones[(size_t)(s0 + f) * b.rows + row] != 0.0f);
    valid[sample] = any ? 0.0 : 1.0;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
  float2* ds = reinterpret_cast<float2*>(states + (size_t)sample * (size_t)F * b.obs_size);
  float2* dn = reinterpret_cast<float2*>(next_states + (size_t)sample * (size_t)F * b.obs_size);
  for (int e = lane; e < n2; e += 32) {
    const float2 a = buf[e], c = buf[e + obs2];
    __stcs(ds + e, a);  // written once, read by the learner later: streaming stores
    __stcs(dn + e, c);
  }
}

static bool sample_staged_ok(const SbBufferDesc& b, const float* states, const float* next_states, int* smem_out) {
#if SB_SAMPLE_ASYNC
  if ((b.obs_size & 1) || (b.states_slot_stride & 1) || (b.states_row_stride & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(b.d_states) | reinterpret_cast<uintptr_t>(states) |
       reinterpret_cast<uintptr_t>(next_states)) & 7u)
    return false;
  const long long bytes = (long long)SAMPLE_WARPS * (b.n_frames + 1) * b.obs_size * 4;
  if (bytes > 64 * 1024) return false;  // long frame stacks: too few resident warps, keep the register-staged loop
  *smem_out = (int)bytes;
  return true;
#else
  return false;
#endif
}

static int launch_sample(const SbBufferDesc& nb, const int64_t* d_rows, const int64_t* d_starts, int batch,
                         float* d_states, int64_t* d_actions, float* d_rewards, float* d_next_states, float* d_dones,
                         double* d_valid, int64_t* d_action_seq, float* d_next_action_seq, cudaStream_t stream) {
  int smem = 0;
  if (sample_staged_ok(nb, d_states, d_next_states, &smem)) {
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(buffer_sample_staged_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return note_cuda_error(e);
    }
    buffer_sample_staged_kernel<<<(unsigned)((batch + SAMPLE_WARPS - 1) / SAMPLE_WARPS), SAMPLE_WARPS * 32, smem,
                                  stream>>>(nb, d_rows, d_starts, batch, d_states, d_actions, d_rewards,
                                            d_next_states, d_dones, d_valid, d_action_seq, d_next_action_seq);
  } else {
    const long long threads = (long long)batch * 32;
    buffer_sample_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        nb, d_rows, d_starts, batch, d_states, d_actions, d_rewards, d_next_states, d_dones, d_valid, d_action_seq,
        d_next_action_seq);
  }
  count_launch();
  return check_launch();
}

__global__ void buffer_current_state_kernel(SbBufferDesc b, int idx, const uint32_t* __restrict__ clock,
                                            float* __restrict__ out) {
  // rows [idx-(F-1), idx) with ring wrap (Buffer.current_state, buffers.py:143-154)
  if (clock) idx = (int)clock[1];  // write head kept on the device (graph-replayed turns)
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= b.rows) return;
  const int F1 = b.n_frames - 1;
  for (int f = 0; f < F1; ++f) {
    int slot = idx - F1 + f;
    if (slot < 0) slot += b.capacity;
    const float* src = b.d_states + (size_t)row * b.states_row_stride + (size_t)slot * b.states_slot_stride;
    float* dst = out + ((size_t)row * F1 + f) * b.obs_size;
    for (int e = lane; e < b.obs_size; e += 32) dst[e] = src[e];
  }
}

// ------------------------------------------------------------------------------------------------
// reward statistics: deterministic two-level reduction (fixed block order)
// ------------------------------------------------------------------------------------------------
constexpr int STATS_BLOCKS = 296, STATS_THREADS = 256;
struct StatsScratch {
  double part[STATS_BLOCKS][4];
  unsigned int ticket;
};

__device__ inline void stats_merge(double& s, double& ss, double& mn, double& mx, double s2, double ss2,
                                   double mn2, double mx2) {
  s += s2;
  ss += ss2;
  mn = fmin(mn, mn2);
  mx = fmax(mx, mx2);
}

__global__ void __launch_bounds__(STATS_THREADS)
    reward_stats_kernel(const double* __restrict__ x, long long n, double* __restrict__ out, StatsScratch* sc) {
  __shared__ double sh[STATS_THREADS / 32][4];
  __shared__ bool last;
  double s = 0, ss = 0, mn = INFINITY, mx = -INFINITY;
  for (long long i = (long long)blockIdx.x * STATS_THREADS + threadIdx.x; i < n;
       i += (long long)gridDim.x * STATS_THREADS) {
    double v = x[i];
    s += v;
    ss += v * v;
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  auto warp_reduce = [&]() {
    for (int o = 16; o > 0; o >>= 1)
      stats_merge(s, ss, mn, mx, __shfl_down_sync(~0u, s, o), __shfl_down_sync(~0u, ss, o),
                  __shfl_down_sync(~0u, mn, o), __shfl_down_sync(~0u, mx, o));
  };
  warp_reduce();
  if ((threadIdx.x & 31) == 0) {
    double* r = sh[threadIdx.x >> 5];
    r[0] = s, r[1] = ss, r[2] = mn, r[3] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < STATS_THREADS / 32; ++w) stats_merge(s, ss, mn, mx, sh[w][0], sh[w][1], sh[w][2], sh[w][3]);
    double* r = sc->part[blockIdx.x];
    r[0] = s, r[1] = ss, r[2] = mn, r[3] = mx;
    __threadfence();
    last = atomicAdd(&sc->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    s = 0, ss = 0, mn = INFINITY, mx = -INFINITY;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      volatile double* r = sc->part[b];
      stats_merge(s, ss, mn, mx, r[0], r[1], r[2], r[3]);
    }
    out[0] = s, out[1] = ss, out[2] = mn, out[3] = mx, out[4] = (double)n;
    sc->ticket = 0;
  }
}

__global__ void stats_combine_kernel(const double* __restrict__ parts, int ranks, double* __restrict__ out) {
  double s = 0, ss = 0, mn = INFINITY, mx = -INFINITY, n = 0;
  for (int r = 0; r < ranks; ++r) {  // fixed rank order: every rank computes the identical result
    const double* p = parts + 5 * r;
    stats_merge(s, ss, mn, mx, p[0], p[1], p[2], p[3]);
    n += p[4];
  }
  out[0] = s, out[1] = ss, out[2] = mn, out[3] = mx, out[4] = n;
}

__global__ void clock_advance_kernel(uint32_t* clock, uint32_t add_tick, uint32_t add_slot, uint32_t capacity) {
  clock[0] += add_tick;
  clock[1] = (clock[1] + add_slot) % capacity;
}

// validates the ring description and returns it with the default (agent-major) strides filled in
static int check_buf(const SbBufferDesc* in, SbBufferDesc* b) {
  if (!in) return SB_EINVAL;
  *b = *in;
  if (b->capacity <= 0 || b->obs_size <= 0 || b->n_frames <= 0 || b->rows <= 0) return SB_EINVAL;
  if (!b->d_states || !b->d_actions || !b->d_rewards || !b->d_dones) return SB_EINVAL;
  if (b->states_row_stride == 0 && b->states_slot_stride == 0) {
    b->states_row_stride = (int64_t)b->capacity * b->obs_size;
    b->states_slot_stride = b->obs_size;
  }
  if (b->states_row_stride < b->obs_size || b->states_slot_stride < b->obs_size) return SB_EINVAL;
  return SB_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_abi_version(void) { return SB_ABI_VERSION; }

const char* sb_strerror(int status) {
  switch (status) {
    case SB_OK: return "ok";
    case SB_EINVAL: return "invalid argument";
    case SB_EUNSUPPORTED: return "shape not supported by the sm_100a kernels";
    case SB_EALIGN: return "pointer is not 16-byte aligned";
    case SB_ECUDA: return "CUDA error (see sb_last_cuda_error)";
    case SB_ENODEVICE: return "no CUDA device with compute capability 10.x";
    case SB_EVISION: return "vision radius exceeds (H-1)/2 or (W-1)/2";
    default: return "unknown status";
  }
}

const char* sb_last_cuda_error(void) { return g_cuda_err; }
uint64_t sb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int32_t sb_plane_stride(int32_t height, int32_t width) { return (height * width + 15) / 16 * 16; }

size_t sb_grid_bytes(const SbWorldDesc* w) {
  if (!w) return 0;
  return (size_t)w->num_envs * (size_t)w->num_dyn_layers * (size_t)sb_plane_stride(w->height, w->width);
}

int sb_validate(const SbWorldDesc* w) {
  SbDev d;
  return make_dev(w, &d);
}

int sb_reset(const SbWorldDesc* w, uint8_t* d_grid, int32_t* d_agent_pos, double* d_total_reward,
             const int32_t* d_spawnable, int32_t n_spawnable, const int32_t* d_placement, uint64_t seed,
             uint32_t tick, void* stream) {
  return sb_reset_ex(w, d_grid, d_agent_pos, d_total_reward, d_spawnable, n_spawnable, d_placement, nullptr, 0,
                     seed, tick, stream);
}

int sb_reset_ex(const SbWorldDesc* w, uint8_t* d_grid, int32_t* d_agent_pos, double* d_total_reward,
                const int32_t* d_spawnable, int32_t n_spawnable, const int32_t* d_placement,
                const SbItemGroup* item_groups, int32_t n_groups, uint64_t seed, uint32_t tick, void* stream) {
  SbDev d;
  int rc = make_dev(w, &d);
  if (rc) return rc;
  if (!d_grid || !d_agent_pos || !d_total_reward || !d_spawnable) return SB_EINVAL;
  if (n_spawnable < d.A) return SB_EINVAL;  // np.random.choice(replace=False) raises ValueError
  if (!aligned16(d_grid) || !aligned16(d.tmpl)) return SB_EALIGN;
  if (n_groups < 0 || n_groups > SB_MAX_ITEM_GROUPS || (n_groups > 0 && !item_groups)) return SB_EINVAL;
  ResetGroups groups;
  memset(&groups, 0, sizeof(groups));
  groups.n = n_groups;
  for (int g = 0; g < n_groups; ++g) {
    const SbItemGroup& ig = item_groups[g];
    if (ig.kind < 0 || ig.kind >= d.K || ig.layer < 0 || ig.layer >= d.L || d.plane_of_layer[ig.layer] < 0)
      return SB_EINVAL;
    if (ig.count < 0 || ig.n_cells < ig.count || !ig.d_cells) return SB_EINVAL;  // choice(replace=False)
    groups.kind[g] = ig.kind;
    groups.plane[g] = d.plane_of_layer[ig.layer];
    groups.count[g] = ig.count;
    groups.n_cells[g] = ig.n_cells;
    groups.cells[g] = ig.d_cells;
    groups.placement[g] = ig.d_placement;
  }
  const int epb = SB_THREADS / RESET_LPE;
  const unsigned blocks = (unsigned)(((long long)d.N + epb - 1) / epb);
  reset_kernel<<<blocks, SB_THREADS, 0, (cudaStream_t)stream>>>(d, d_grid, d_agent_pos, d_total_reward,
                                                                d_spawnable, n_spawnable, d_placement, groups,
                                                                seed, tick);
  count_launch();
  return check_launch();
}

// validation shared by sb_step and sb_step_plan (`plan_only`: the state pointers are not required)
static int prepare_step(const SbWorldDesc* w, const SbStepArgs* a, const SbBufferDesc* buf, bool plan_only, SbDev* dout,
                        StepParams* pout) {
  SbDev& d = *dout;
  int rc = make_dev(w, &d);
  if (rc) return rc;
  if (!a) return SB_EINVAL;
  if (a->num_turns < 1) return SB_EINVAL;
  StepParams& p = *pout;
  memset(&p, 0, sizeof(p));
  p.a = *a;
  if (a->mode == SB_MODE_SPLIT) {
    if (a->num_turns != 1) return SB_EINVAL;
    if (a->act_agent >= d.A || a->obs_agent >= d.A) return SB_EINVAL;
    if (a->act_agent < 0) p.a.act_agent = -1;
    if (a->obs_agent < 0 && a->obs_agent != SB_OBS_ALL) p.a.obs_agent = -1;
    if (!plan_only && a->act_agent >= 0 && a->policy == SB_POLICY_TAPE && !a->d_actions) return SB_EINVAL;
  } else if (a->mode == SB_MODE_FUSED) {
    if (!plan_only && a->policy == SB_POLICY_TAPE && !a->d_actions) return SB_EINVAL;
  } else {
    return SB_EINVAL;
  }
  const bool sweeps = a->mode == SB_MODE_FUSED || a->do_sweep;
  if (!plan_only && sweeps && a->rng == SB_RNG_TAPE && (!a->d_spawn_u || !a->d_spawn_c)) return SB_EINVAL;
  if (a->rng != SB_RNG_TAPE && a->rng != SB_RNG_PHILOX && a->rng != SB_RNG_SPARSE) return SB_EINVAL;
  if (a->rng == SB_RNG_SPARSE && !d.sparse_ok) return SB_EINVAL;
  if (!plan_only && w->uses_facing && !a->d_agent_facing) return SB_EINVAL;
  if (a->policy != SB_POLICY_TAPE && a->policy != SB_POLICY_UNIFORM) return SB_EINVAL;
  if (buf) {
    rc = check_buf(buf, &p.buf);
    if (rc) return rc;
    if (buf->obs_size != d.obs_size || buf->rows != (int64_t)d.N * d.A) return SB_EINVAL;
    if (!a->d_clock && (a->slot0 < 0 || a->slot0 >= buf->capacity)) return SB_EINVAL;
    p.has_buf = 1;
  }
  p.write_back = 1;
  return SB_OK;
}

int sb_step(const SbWorldDesc* w, const SbStepArgs* a, uint8_t* d_grid, int32_t* d_agent_pos,
            double* d_total_reward, const SbBufferDesc* buf, void* stream) {
  SbDev d;
  StepParams p;
  if (!d_grid || !d_agent_pos || !d_total_reward) return SB_EINVAL;
  if (!aligned16(d_grid)) return SB_EALIGN;
  const int rc = prepare_step(w, a, buf, false, &d, &p);
  if (rc) return rc;
  p.grid = d_grid;
  p.agent_pos = d_agent_pos;
  p.total_reward = d_total_reward;
  return launch_step(d, p, (cudaStream_t)stream);
}

int sb_step_plan(const SbWorldDesc* w, const SbStepArgs* a, SbStepPlan* out) {
  SbDev d;
  StepParams p;
  if (!out) return SB_EINVAL;
  memset(out, 0, sizeof(*out));
  const int rc = prepare_step(w, a, nullptr, true, &d, &p);
  if (rc) return rc;
  p.has_buf = 1;  // the plan is for the usual launch: a replay ring attached, no second observation output
  return launch_step(d, p, nullptr, out);
}

int sb_clock_advance(uint32_t* d_clock, uint32_t add_tick, uint32_t add_slot, uint32_t capacity, void* stream) {
  if (!d_clock || capacity == 0) return SB_EINVAL;
  clock_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_clock, add_tick, add_slot, capacity);
  count_launch();
  return check_launch();
}

int sb_observe(const SbWorldDesc* w, const uint8_t* d_grid, const int32_t* d_agent_pos, float* d_obs_out,
               void* stream) {
  SbDev d;
  int rc = make_dev(w, &d);
  if (rc) return rc;
  if (!d_grid || !d_agent_pos || !d_obs_out) return SB_EINVAL;
  if (!aligned16(d_grid)) return SB_EALIGN;
  return launch_observe(d, d_grid, d_agent_pos, d_obs_out, (cudaStream_t)stream);
}

int sb_buffer_add(const SbBufferDesc* buf, int32_t slot, const float* d_obs, const int64_t* d_actions,
                  const float* d_rewards, const float* d_dones, const int64_t* d_positions, void* stream) {
  SbBufferDesc nb;
  int rc = check_buf(buf, &nb);
  if (rc) return rc;
  if (slot < 0 || slot >= buf->capacity || !d_obs || !d_actions || !d_rewards || !d_dones) return SB_EINVAL;
  const long long threads = buf->rows * 32;
  buffer_add_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      nb, slot, d_obs, d_actions, d_rewards, d_dones, d_positions);
  count_launch();
  return check_launch();
}

int sb_buffer_sample(const SbBufferDesc* buf, const int64_t* d_rows, const int64_t* d_starts, int32_t batch,
                     float* d_states, int64_t* d_actions, float* d_rewards, float* d_next_states,
                     float* d_dones, double* d_valid, void* stream) {
  SbBufferDesc nb;
  int rc = check_buf(buf, &nb);
  if (rc) return rc;
  if (batch <= 0 || !d_rows || !d_starts || !d_states || !d_actions || !d_rewards || !d_next_states ||
      !d_dones || !d_valid)
    return SB_EINVAL;
  if (buf->capacity < buf->n_frames + 1) return SB_EINVAL;  // a transition spans n_frames + 1 consecutive slots
  return launch_sample(nb, d_rows, d_starts, batch, d_states, d_actions, d_rewards, d_next_states, d_dones, d_valid,
                       nullptr, nullptr, (cudaStream_t)stream);
}

int sb_buffer_sample_transformer(const SbBufferDesc* buf, const int64_t* d_rows, const int64_t* d_starts,
                                 int32_t batch, float* d_states, int64_t* d_action_seq, float* d_next_action_seq,
                                 float* d_next_states, float* d_dones, double* d_valid, void* stream) {
  SbBufferDesc nb;
  int rc = check_buf(buf, &nb);
  if (rc) return rc;
  if (batch <= 0 || !d_rows || !d_starts || !d_states || !d_action_seq || !d_next_action_seq || !d_next_states ||
      !d_dones || !d_valid)
    return SB_EINVAL;
  if (buf->n_frames > 32) return SB_EUNSUPPORTED;
  if (buf->capacity < buf->n_frames + 1) return SB_EINVAL;
  return launch_sample(nb, d_rows, d_starts, batch, d_states, nullptr, nullptr, d_next_states, d_dones, d_valid,
                       d_action_seq, d_next_action_seq, (cudaStream_t)stream);
}

int sb_buffer_current_state(const SbBufferDesc* buf, int32_t idx, float* d_out, void* stream) {
  SbBufferDesc nb;
  int rc = check_buf(buf, &nb);
  if (rc) return rc;
  if (idx < 0 || idx >= buf->capacity || !d_out) return SB_EINVAL;
  if (buf->n_frames < 2) return SB_OK;  // zero rows, like states[idx:idx]
  const long long threads = buf->rows * 32;
  buffer_current_state_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nb, idx, nullptr,
                                                                                                   d_out);
  count_launch();
  return check_launch();
}

int sb_buffer_current_state_dev(const SbBufferDesc* buf, const uint32_t* d_clock, float* d_out, void* stream) {
  SbBufferDesc nb;
  int rc = check_buf(buf, &nb);
  if (rc) return rc;
  if (!d_clock || !d_out) return SB_EINVAL;
  if (buf->n_frames < 2) return SB_OK;
  const long long threads = buf->rows * 32;
  buffer_current_state_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nb, 0, d_clock,
                                                                                                   d_out);
  count_launch();
  return check_launch();
}

size_t sb_stats_scratch_bytes(void) { return sizeof(StatsScratch); }

int sb_reward_stats(const double* d_total_reward, int64_t n, double* d_out, void* d_scratch, void* stream) {
  if (!d_total_reward || n <= 0 || !d_out || !d_scratch) return SB_EINVAL;
  long long want = (n + STATS_THREADS - 1) / STATS_THREADS;
  unsigned blocks = (unsigned)(want < STATS_BLOCKS ? want : STATS_BLOCKS);
  reward_stats_kernel<<<blocks, STATS_THREADS, 0, (cudaStream_t)stream>>>(d_total_reward, n, d_out,
                                                                         (StatsScratch*)d_scratch);
  count_launch();
  return check_launch();
}

int sb_stats_combine(const double* d_parts, int32_t num_ranks, double* d_out, void* stream) {
  if (!d_parts || !d_out || num_ranks <= 0) return SB_EINVAL;
  stats_combine_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_parts, num_ranks, d_out);
  count_launch();
  return check_launch();
}

}  // extern "C"
