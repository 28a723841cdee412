// hostcheck.cpp -- compiles the kernels' per-lane device functions (sorrel_b200/csrc/sb_core.cuh) for the
// HOST so that the CPU test-suite can unit-test them against the oracle without a GPU.
// TEST-ONLY: nothing in the product loads this library; it is not a CPU fallback (no kernel, no tiling,
// no entry point of include/sorrel_b200.h is implemented here).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../sorrel_b200/csrc/sb_core.cuh"

extern "C" {

void hc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out, uint32_t* out_rk) {
  SbU4 r = sb_philox(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
  // the engine's keyed stream: fixed key, `key` folded into counter words 1 and 3
  SbU4 q = sb_philox_rk(ctr[0], ctr[1], ctr[2], ctr[3], key);
  out_rk[0] = q.x, out_rk[1] = q.y, out_rk[2] = q.z, out_rk[3] = q.w;
}

static void set_prob(SbDev* d, double p, int n_choices) {
  uint64_t thr = !(p > 0.0) ? 0 : (p >= 1.0 ? (1ull << 53) : (uint64_t)ceil(ldexp(p, 53)));
  d->p = p;
  d->n2 = sb_level1_bound(thr);
  d->thr64m1 = sb_level2_bound(thr, d->n2);
  d->n_choices = n_choices;
  d->choices_packed = 0x0706050403020100ull;
  d->hmask4 = ((~(d->n2 - 1u)) & 0xFFu) * 0x01010101u;
}

// first-level filter of one chunk: returns the 16-bit mask (bit = byte index in the chunk) of candidates
uint32_t hc_filter_chunk(const uint8_t* kinds16, const uint8_t* rnd16, int spawn_kind, double p) {
  SbDev d;
  memset(&d, 0, sizeof d);
  set_prob(&d, p, 3);
  uint32_t k[4], r[4];
  memcpy(k, kinds16, 16);
  memcpy(r, rnd16, 16);
  uint32_t m = sb_simd_cand_chunk(k[0], k[1], k[2], k[3], r[0], r[1], r[2], r[3], (uint32_t)spawn_kind * 0x01010101u,
                                  d.hmask4);
  uint32_t out = 0;
  for (int pos = 0; pos < 32; ++pos)
    if (m & (1u << pos)) out |= 1u << (4 * (pos & 7) + (pos >> 3));
  return out;
}

uint32_t hc_filter_bound(double p) {  // n2: the filter passes bytes < n2
  SbDev d;
  memset(&d, 0, sizeof d);
  set_prob(&d, p, 3);
  return ((~d.hmask4) & 0xFFu) + 1u;
}

int hc_resolve_candidate(double p, int n_choices, uint32_t env, uint32_t tick, uint32_t cell, uint64_t seed) {
  SbDev d;
  memset(&d, 0, sizeof d);
  set_prob(&d, p, n_choices);
  const uint32_t rk[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  // level 1 the way the sweep evaluates it (byte (cell & 15) of the chunk's call), then the resolver = level 2
  const SbU4 r = sb_philox_rk(env, tick, cell >> 4, SB_DOM_SWEEP, rk);
  if (sb_byte_of(r, (int)(cell & 15u)) >= d.n2) return -1;
  return sb_resolve_candidate(d, env, tick, cell, rk);
}

uint32_t hc_action(uint32_t env, uint32_t tick, uint32_t agent, uint32_t n_actions, uint64_t seed) {
  const uint32_t sk[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  SbU4 w = sb_philox_rk(env, tick, agent, SB_DOM_ACTION, sk);
  return sb_bounded(w.x, w.y, n_actions);
}

// one agent move on a single-plane world; returns reward, updates grid / yx / err
float hc_agent_act(uint8_t* grid, int H, int W, int K, const float* value, uint32_t passable_mask, int agent_kind,
                   int default_kind, uint32_t* yx, int action, uint32_t* err) {
  SbDev d;
  memset(&d, 0, sizeof d);
  d.H = H, d.W = W, d.K = K, d.plane_stride = (H * W + 15) / 16 * 16, d.agent_plane = 0;
  d.n_actions = 4;
  const int8_t dy[4] = {-1, 1, 0, 0}, dx[4] = {0, 0, -1, 1};
  for (int i = 0; i < 4; ++i) {
    d.dy_packed |= (uint64_t)(uint8_t)dy[i] << (8 * i);
    d.dx_packed |= (uint64_t)(uint8_t)dx[i] << (8 * i);
  }
  for (int k = 0; k < K; ++k) d.value[k] = value[k];
  d.passable_mask = passable_mask, d.agent_kind = agent_kind, d.default_kind = default_kind;
  uint32_t e = 0;
  float r = sb_agent_act(d, grid, yx, action, true, e);
  *err = e;
  return r;
}

// general-path observation of one element
double hc_obs_general(const uint8_t* planes, const uint8_t* tmpl, const double* app, int H, int W, int L,
                      const int* plane_of_layer, int K, int C, int v, int fill_kind, int y, int x, int c, int i, int j) {
  SbDev d;
  memset(&d, 0, sizeof d);
  d.H = H, d.W = W, d.L = L, d.K = K, d.C = C, d.v = v, d.fill_kind = fill_kind;
  d.plane_stride = (H * W + 15) / 16 * 16;
  for (int z = 0; z < L; ++z) d.plane_of_layer[z] = plane_of_layer[z];
  return sb_obs_general(d, planes, tmpl, app, y, x, c, i, j);
}

// SB_RULE_SPAWN_TO under the counter-based stream: 1 = the cell spawns (first-level byte recomputed here the
// way the generic sweep does: byte (cell & 15) of Philox(env, tick, cell >> 4, SWEEP))
int hc_bernoulli_hit(double p, uint32_t env, uint32_t tick, uint32_t cell, uint64_t seed) {
  const uint64_t thr = !(p > 0.0) ? 0 : (p >= 1.0 ? (1ull << 53) : (uint64_t)ceil(ldexp(p, 53)));
  const uint32_t rk[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const SbU4 r = sb_philox_rk(env, tick, cell >> 4, SB_DOM_SWEEP, rk);
  const uint32_t n2 = sb_level1_bound(thr);
  return sb_bernoulli_hit(n2, sb_level2_bound(thr, n2), sb_byte_of(r, (int)(cell & 15u)), env, tick, cell, rk) ? 1 : 0;
}

// deterministic column rules on a 2-plane image: planes[2][stride]; the rule of `kind` tests plane `other_plane`
int hc_column_rule(const uint8_t* planes, int stride, int off, int rule, int to, int other_plane, uint32_t mask) {
  SbDev d;
  memset(&d, 0, sizeof d);
  const uint32_t kind = planes[off];
  d.rule_to[kind] = (uint8_t)to;
  d.rule_plane_off[kind] = other_plane * stride;
  d.rule_mask[kind] = mask;
  return sb_column_rule(d, planes, kind, (uint32_t)rule, off);
}

// all-layer reward (observe_all_layers) through sb_agent_act on an L-plane image
float hc_agent_act_all_layers(uint8_t* planes, int H, int W, int L, int K, const float* value, uint32_t passable,
                              int agent_plane, int agent_kind, int default_kind, uint32_t* yx, int action,
                              uint32_t* err) {
  SbDev d;
  memset(&d, 0, sizeof d);
  d.H = H, d.W = W, d.K = K, d.L = L, d.n_planes = L;
  d.plane_stride = (H * W + 15) / 16 * 16;
  for (int z = 0; z < L; ++z) d.plane_of_layer[z] = z;
  d.agent_plane = agent_plane;
  d.agent_kind = agent_kind, d.default_kind = default_kind, d.n_actions = 6;
  d.passable_mask = passable;
  const int8_t dy[6] = {-1, 1, 0, 0, 0, 0}, dx[6] = {0, 0, -1, 1, 0, 0};
  for (int i = 0; i < 6; ++i) {
    d.dy_packed |= (uint64_t)(uint8_t)dy[i] << (8 * i);
    d.dx_packed |= (uint64_t)(uint8_t)dx[i] << (8 * i);
  }
  for (int k = 0; k < K; ++k) d.value[k] = value[k], d.value64[k] = (double)value[k];
  uint32_t e = 0;
  float r = sb_agent_act(d, planes, yx, action, true, e, true);
  *err = e;
  return r;
}

// role hand-over of examples/tag on a 1-plane image; returns the acting agent's kind afterwards
int hc_touch_swap(uint8_t* plane, int H, int W, int y, int x, int agent_kind, int agent_kind_alt) {
  SbDev d;
  memset(&d, 0, sizeof d);
  d.H = H, d.W = W, d.plane_stride = (H * W + 15) / 16 * 16, d.agent_plane = 0;
  d.agent_kind = agent_kind, d.agent_kind_alt = agent_kind_alt;
  return (int)sb_touch_swap(d, plane, ((uint32_t)y << 16) | (uint32_t)x);
}
}
