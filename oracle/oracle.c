/*
 * oracle.c -- plain C restatement of the reference step / observation / replay-buffer path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs as the CHECKER and the CPU baseline.  The product
 * (sorrel_b200/) never links, loads or calls it.
 *
 * Follows (file:line relative to the reference root, social-ai-uoft/sorrel v1.4.0):
 *   Environment.reset / take_turn        sorrel/environment.py:72-93
 *   populate_environment                 sorrel/examples/treasurehunt/env.py:114-147
 *   EmptyEntity.transition               sorrel/examples/treasurehunt/entities.py:69-85
 *   Agent.transition / MovingAgent.act   sorrel/agents/agent.py:155-173, 187-225
 *   Gridworld.observe / move             sorrel/worlds/gridworld.py:95-133
 *   visual_field (closed form)           sorrel/observation/visual_field.py:9-101
 *   Buffer.add                           sorrel/buffers.py:46-63
 *   cleanup rule set: River/Pollution/AppleTree/Apple.transition  sorrel/examples/cleanup/entities.py:43-105,
 *                     Beam.transition, CleanupAgent.spawn_beam / act, CleanupObservation.observe
 *                                                                  sorrel/examples/cleanup/agents.py:21-216,
 *                     populate_environment                         sorrel/examples/cleanup/env.py:85-140
 *
 * Randomness (three interchangeable sources, same step code):
 *   OR_RNG_NUMPY  -- own MT19937 + the numpy *legacy* RandomState draw algorithms (numpy is a
 *                    third-party dependency of the reference, pinned ^2.5.0 in pyproject.toml:33,
 *                    2.3.5 installed; its legacy stream is frozen by NEP 19): init_genrand seeding,
 *                    random_sample = (a>>5, b>>6) 53-bit doubles, randint/choice = masked rejection on
 *                    32-bit outputs, permutation = Fisher-Yates from the top with masked rejection.
 *                    Validated against numpy itself in tests/test_coracle.py.
 *   OR_RNG_PHILOX -- the engine's counter-based stream (Philox4x32-10).
 *   OR_RNG_SPARSE -- the same stream with the spawn field sampled sparsely (Binomial count by inverse CDF on a
 *                    caller-supplied exact table, then distinct uniform cells); placement and actions as PHILOX.
 *   tapes         -- OR_RNG_NUMPY can record the dense replay tapes the CUDA kernels consume.
 *
 * Parity status: pinned against tests/golden/ (generated from the unmodified reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define OR_MAX_KINDS 32
#define OR_MAX_LAYERS 8
#define OR_RNG_NUMPY 0
#define OR_RNG_PHILOX 1
#define OR_RNG_SPARSE 2 /* Philox stream, sparse sampling of the spawn field (DESIGN.md "Counter-based RNG") */
#define OR_SPARSE_KMAX 64
#define OR_RULE_SPAWN 1
#define OR_RULE_SPAWN_TO 2  /* Bernoulli(rule_prob[k]) -> rule_to[k], one uniform, no choice draw */
#define OR_RULE_BECOME_IF 3 /* same column on rule_layer[k] holds a kind of rule_mask[k] -> rule_to[k] */
#define OR_RULE_BECOME 4    /* -> rule_to[k] (beam ageing: turn_counter folded into the kind id) */
#define OR_NO_KIND 255
#define OR_MAX_BEAM 16
#define OR_MAX_GROUPS 4

#define DOM_SWEEP 0x53570000u
#define DOM_CELL 0x43450000u
#define DOM_ACTION 0x41430000u
#define DOM_RESET 0x52530000u
#define DOM_SPARSE 0x53500000u

/* ---------------------------------------------------------------------------------------------
 * MT19937 + numpy legacy draws
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t key[624];
  int pos;
} mt_t;

static void mt_seed(mt_t* s, uint32_t seed) { /* numpy mt19937_seed == init_genrand */
  for (int i = 0; i < 624; ++i) {
    s->key[i] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
  }
  s->pos = 624;
}

static void mt_gen(mt_t* s) {
  const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
  uint32_t* k = s->key;
  int i;
  for (i = 0; i < 624 - 397; ++i) {
    uint32_t y = (k[i] & UP) | (k[i + 1] & LO);
    k[i] = k[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  }
  for (; i < 623; ++i) {
    uint32_t y = (k[i] & UP) | (k[i + 1] & LO);
    k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  }
  uint32_t y = (k[623] & UP) | (k[0] & LO);
  k[623] = k[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  s->pos = 0;
}

static inline uint32_t mt_next32(mt_t* s) {
  if (s->pos == 624) mt_gen(s);
  uint32_t y = s->key[s->pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* RandomState.random_sample(): mt19937_next_double */
static inline double np_random_sample(mt_t* s) {
  uint32_t a = mt_next32(s) >> 5, b = mt_next32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

static inline uint32_t mask_ge(uint32_t m) {
  m |= m >> 1;
  m |= m >> 2;
  m |= m >> 4;
  m |= m >> 8;
  m |= m >> 16;
  return m;
}

/* RandomState.randint(0, n) / choice(n): masked rejection on 32-bit draws (n-1 < 2^32-1) */
static inline uint32_t np_randint(mt_t* s, uint32_t n) {
  uint32_t rng = n - 1;
  if (rng == 0) return 0;
  uint32_t mask = mask_ge(rng), v;
  while ((v = (mt_next32(s) & mask)) > rng) {
  }
  return v;
}

/* RandomState.permutation(n)[:k]: arange + legacy shuffle (i = n-1..1, j = random_interval(i)) */
static void np_permutation_head(mt_t* s, int n, int k, int32_t* out, int32_t* scratch) {
  for (int i = 0; i < n; ++i) scratch[i] = i;
  for (int i = n - 1; i >= 1; --i) {
    uint32_t mask = mask_ge((uint32_t)i), j;
    while ((j = (mt_next32(s) & mask)) > (uint32_t)i) {
    }
    int32_t t = scratch[i];
    scratch[i] = scratch[j];
    scratch[j] = t;
  }
  for (int i = 0; i < k; ++i) out[i] = scratch[i];
}

/* ---------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11)
 * ------------------------------------------------------------------------------------------- */
static inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

/* The engine's keyed stream (sb_core.cuh sb_philox_rk): FIXED key, the 64-bit seed folded into counter words 1, 3. */
#define PHILOX_K0 0xA4093822u
#define PHILOX_K1 0x299F31D0u
static inline void draw(uint32_t env, uint32_t tick, uint32_t index, uint32_t domain, uint32_t s0, uint32_t s1,
                        uint32_t out[4]) {
  philox(env, tick ^ s0, index, domain ^ s1, PHILOX_K0, PHILOX_K1, out);
}

static inline uint32_t bounded64(uint32_t hi, uint32_t lo, uint32_t n) {
  return (uint32_t)(((unsigned __int128)(((uint64_t)hi << 32) | lo) * n) >> 64);
}

/* ---------------------------------------------------------------------------------------------
 * world description / state
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t H, W, L, A, K, C, v;
  int32_t n_dyn, dyn_layers[OR_MAX_LAYERS];
  int32_t agent_layer, fill_kind, agent_kind, default_kind;
  int32_t n_actions, dy[8], dx[8];
  int32_t n_choices, choices[8];
  int32_t n_spawnable;
  int32_t obs_transform; /* 0 none, 1 RGB: clip(sum, 0, 255) / 255 (observation_spec.py:483) */
  double p;
  double value[OR_MAX_KINDS]; /* Entity.value as the Python float it is (iowa: 1.1, -1.4, ...) */
  uint8_t passable[OR_MAX_KINDS], rule[OR_MAX_KINDS];
  const double* appearance; /* [K][C] */
  const uint8_t* tmpl;      /* [H][W][L] (reference index order y, x, z) */
  const int32_t* spawnable; /* [n_spawnable] y*W+x on the agent layer, reference order */
  /* ---- column rules, facings + beams, all-layer rewards, positional code, item groups (examples/cleanup) ---- */
  double rule_prob[OR_MAX_KINDS];
  uint8_t rule_to[OR_MAX_KINDS];
  uint8_t rule_else[OR_MAX_KINDS]; /* SPAWN_TO: kind when the draw fails, OR_NO_KIND = unchanged */
  int8_t rule_layer[OR_MAX_KINDS];
  uint32_t rule_mask[OR_MAX_KINDS];
  int32_t uses_facing, initial_facing;
  int8_t action_facing[8];
  uint8_t action_beam_kind[8];
  int32_t beam_layer, n_beam;
  int8_t beam_dy[4][OR_MAX_BEAM], beam_dx[4][OR_MAX_BEAM];
  uint32_t beam_block_mask;
  int32_t reward_all_layers, reward_count;
  int32_t obs_extra;
  const double* extra; /* [H*W][obs_extra] float64 (rounded to f32 when stored, like Buffer.states) */
  int32_t n_groups, group_kind[OR_MAX_GROUPS], group_layer[OR_MAX_GROUPS], group_count[OR_MAX_GROUPS],
      group_n_cells[OR_MAX_GROUPS];
  const int32_t* group_cells[OR_MAX_GROUPS]; /* y*W+x, reference order */
  /* agents with a binary state shown as a second kind (examples/tag): TagAgent.it, tag/agents.py:40-52,76-106 */
  int32_t agent_kind_alt, touch_swap, reward_own, obs_flag;
  int32_t initial_flags[64];
} or_world_t;

typedef struct {
  or_world_t w;
  int64_t N;
  int rng_mode, threads, capacity, n_frames, obs_size, plane_stride, record;
  uint64_t seed, env_id_base;
  uint32_t tick; /* global turn counter */
  int idx, size;
  /* state */
  uint8_t* grid; /* [N][H][W][L] */
  int32_t* pos;  /* [N][A] cell index */
  int32_t* facing; /* [N][A] MovingAgent.direction: survives reset */
  int32_t* flags;  /* [N][A] TagAgent.it: survives reset */
  double* total; /* [N] */
  mt_t* mt;      /* [N] numpy streams (seed = seed + n) */
  /* ring, engine layout */
  float* states;      /* [N*A][cap][obs] */
  int64_t* actions;   /* [cap][N*A] */
  float* rewards;     /* [cap][N*A] */
  float* dones;       /* [cap][N*A] */
  int64_t* positions; /* [cap][N*A][2] */
  /* optional tapes of the last or_step call (kernel layout: padded planes) */
  double* tape_u;   /* [T][N][n_dyn*plane_stride] */
  uint8_t* tape_c;  /* same */
  int32_t* tape_a;  /* [T][N][A] */
  int32_t* tape_pl; /* [N][A] placement of the last reset */
  int tape_T;
  uint64_t cdf[OR_SPARSE_KMAX]; /* OR_RNG_SPARSE: floor(2^64 P(X <= k)), X ~ Binomial(n_dyn*plane_stride, P/2^53) */
  uint64_t* env_ids;            /* optional [N]: global id of every instance (else env_id_base + n) */
} or_env_t;

static inline uint32_t env_id_of(const or_env_t* e, int64_t n) {
  return (uint32_t)(e->env_ids ? e->env_ids[n] : e->env_id_base + (uint64_t)n);
}

static int zd_of_layer(const or_world_t* w, int z) {
  for (int i = 0; i < w->n_dyn; ++i)
    if (w->dyn_layers[i] == z) return i;
  return -1;
}

or_env_t* or_create(const or_world_t* w, int64_t N, int capacity, int n_frames, int rng_mode, uint64_t seed,
                    uint64_t env_id_base, int threads, int record) {
  or_env_t* e = (or_env_t*)calloc(1, sizeof(or_env_t));
  e->w = *w;
  e->N = N;
  e->rng_mode = rng_mode;
  e->seed = seed;
  e->env_id_base = env_id_base;
  e->threads = threads > 0 ? threads : 1;
  e->capacity = capacity;
  e->n_frames = n_frames;
  e->record = record;
  int S = 2 * w->v + 1;
  e->obs_size = w->C * S * S + w->obs_extra + (w->obs_flag ? 1 : 0);
  e->plane_stride = (w->H * w->W + 15) / 16 * 16;
  size_t cells = (size_t)w->H * w->W * w->L, rows = (size_t)N * w->A;
  e->grid = (uint8_t*)malloc(N * cells);
  e->pos = (int32_t*)calloc(rows, sizeof(int32_t));
  e->facing = (int32_t*)calloc(rows, sizeof(int32_t));
  for (size_t r = 0; r < rows; ++r) e->facing[r] = w->initial_facing;
  e->flags = (int32_t*)calloc(rows, sizeof(int32_t));
  for (size_t r = 0; r < rows; ++r) e->flags[r] = w->initial_flags[r % (size_t)w->A];
  e->total = (double*)calloc(N, sizeof(double));
  if (rng_mode == OR_RNG_NUMPY) {
    e->mt = (mt_t*)malloc(N * sizeof(mt_t));
    for (int64_t n = 0; n < N; ++n) mt_seed(&e->mt[n], (uint32_t)(seed + (uint64_t)n));
  }
  if (capacity > 0) {
    e->states = (float*)calloc(rows * capacity * e->obs_size, sizeof(float));
    e->actions = (int64_t*)calloc(rows * capacity, sizeof(int64_t));
    e->rewards = (float*)calloc(rows * capacity, sizeof(float));
    e->dones = (float*)calloc(rows * capacity, sizeof(float));
    e->positions = (int64_t*)calloc(rows * capacity * 2, sizeof(int64_t));
  }
  e->tape_pl = (int32_t*)calloc(rows, sizeof(int32_t));
  return e;
}

void or_destroy(or_env_t* e) {
  if (!e) return;
  free(e->facing);
  free(e->flags);
  free(e->grid), free(e->pos), free(e->total), free(e->mt), free(e->states), free(e->actions);
  free(e->rewards), free(e->dones), free(e->positions), free(e->tape_u), free(e->tape_c), free(e->tape_a);
  free(e->tape_pl);
  free(e->env_ids);
  free(e);
}

/* Environment.reset (environment.py:72-79) + populate_environment (treasurehunt/env.py:114-147) */
static void reset_range(or_env_t* e, int64_t n0, int64_t n1) {
  const or_world_t* w = &e->w;
  const size_t cells = (size_t)w->H * w->W * w->L;
  int32_t* scratch = (int32_t*)malloc(sizeof(int32_t) * (w->n_spawnable > 0 ? w->n_spawnable : 1));
  int32_t picks[64];
  for (int64_t n = n0; n < n1; ++n) {
    uint8_t* g = e->grid + n * cells;
    memcpy(g, w->tmpl, cells);
    e->total[n] = 0.0; /* create_world, gridworld.py:65 */
    /* item groups first, in order (cleanup/env.py:126-132: apples before agents) */
    for (int gi = 0; gi < w->n_groups; ++gi) {
      int32_t gp[256];
      const int cnt = w->group_count[gi], nc = w->group_n_cells[gi];
      if (e->rng_mode == OR_RNG_NUMPY) {
        int32_t* sc = (int32_t*)malloc(sizeof(int32_t) * (nc > 0 ? nc : 1));
        np_permutation_head(&e->mt[n], nc, cnt, gp, sc);
        free(sc);
      } else {
        uint32_t env = env_id_of(e, n), j = (uint32_t)(gi + 1) << 24, r[4];
        for (int a = 0; a < cnt; ++a) {
          for (;;) {
            draw(env, e->tick + 1, j++, DOM_RESET, (uint32_t)e->seed, (uint32_t)(e->seed >> 32), r);
            int32_t cand = (int32_t)bounded64(r[0], r[1], (uint32_t)nc);
            int dup = 0;
            for (int b = 0; b < a; ++b) dup |= gp[b] == cand;
            if (!dup) {
              gp[a] = cand;
              break;
            }
          }
        }
      }
      for (int a = 0; a < cnt; ++a)
        g[(size_t)w->group_cells[gi][gp[a]] * w->L + w->group_layer[gi]] = (uint8_t)w->group_kind[gi];
    }
    if (e->rng_mode == OR_RNG_NUMPY) {
      /* np.random.choice(len(valid), size=A, replace=False) == permutation(len)[:A], env.py:141 */
      np_permutation_head(&e->mt[n], w->n_spawnable, w->A, picks, scratch);
    } else {
      uint32_t env = env_id_of(e, n), j = 0, r[4];
      for (int a = 0; a < w->A; ++a) {
        for (;;) {
          draw(env, e->tick + 1, j++, DOM_RESET, (uint32_t)e->seed, (uint32_t)(e->seed >> 32), r);
          int32_t cand = (int32_t)bounded64(r[0], r[1], (uint32_t)w->n_spawnable);
          int dup = 0;
          for (int b = 0; b < a; ++b) dup |= picks[b] == cand;
          if (!dup) {
            picks[a] = cand;
            break;
          }
        }
      }
    }
    for (int a = 0; a < w->A; ++a) {
      int cell = w->spawnable[picks[a]];
      /* world.add, gridworld.py:67-76; a flagged agent is drawn with its other kind */
      g[(size_t)cell * w->L + w->agent_layer] =
          (uint8_t)((w->agent_kind_alt >= 0 && e->flags[n * w->A + a]) ? w->agent_kind_alt : w->agent_kind);
      e->pos[n * w->A + a] = cell;
      e->tape_pl[n * w->A + a] = picks[a];
    }
  }
  free(scratch);
}

/* closed form of visual_field (visual_field.py:9-101): float64 sum over layers in z order,
 * stored as float32 (Buffer.states, buffers.py:31,56) */
static void observe(const or_world_t* w, const uint8_t* g, int y, int x, float* out) {
  const int S = 2 * w->v + 1, C = w->C;
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) {
      int yy = y - w->v + i, xx = x - w->v + j;
      if (yy >= 0 && yy < w->H && xx >= 0 && xx < w->W) {
        const uint8_t* cell = g + ((size_t)yy * w->W + xx) * w->L;
        for (int c = 0; c < C; ++c) {
          double s = 0.0;
          for (int z = 0; z < w->L; ++z) s += w->appearance[cell[z] * C + c];
          if (w->obs_transform == 1) s = (s < 0.0 ? 0.0 : (s > 255.0 ? 255.0 : s)) / 255.0;
          out[(c * S + i) * S + j] = (float)s;
        }
      } else {
        for (int c = 0; c < C; ++c) {
          double s = w->appearance[w->fill_kind * C + c];
          if (w->obs_transform == 1) s = (s < 0.0 ? 0.0 : (s > 255.0 ? 255.0 : s)) / 255.0;
          out[(c * S + i) * S + j] = (float)s;
        }
      }
    }
  /* CleanupObservation.observe (cleanup/agents.py:52-62): positional code of the agent's cell appended */
  for (int k = 0; k < w->obs_extra; ++k)
    out[C * S * S + k] = (float)w->extra[(size_t)(y * w->W + x) * w->obs_extra + k];
}

static void step_env(or_env_t* e, int64_t n, int T, int t_base) {
  const or_world_t* w = &e->w;
  const size_t cells = (size_t)w->H * w->W * w->L, rows = (size_t)e->N * w->A;
  uint8_t* g = e->grid + n * cells;
  mt_t* mt = e->mt ? &e->mt[n] : 0;
  const uint32_t env = env_id_of(e, n);
  const uint32_t k0 = (uint32_t)e->seed, k1 = (uint32_t)(e->seed >> 32);
  const size_t tape_stride = (size_t)w->n_dyn * e->plane_stride;
  float obs_local[32 * 21 * 21 + 256];
  for (int t = 0; t < T; ++t) {
    const uint32_t tick = e->tick + 1 + (uint32_t)t;
    double* tu = e->record ? e->tape_u + ((size_t)(t_base + t) * e->N + n) * tape_stride : 0;
    uint8_t* tc = e->record ? e->tape_c + ((size_t)(t_base + t) * e->N + n) * tape_stride : 0;
    if (e->rng_mode == OR_RNG_SPARSE) {
      /* The iid Bernoulli(p) field of entities.py:72-75 sampled sparsely: K ~ Binomial(cells, P/2^53) by inverse
         CDF, then K distinct uniform cells of the padded plane layout by sequential rejection over the draw stream
         j = 1, 2, ...; a hit spawns iff the cell (still) carries the spawn rule. */
      uint32_t r[4];
      const uint32_t n_cells = (uint32_t)tape_stride;
      draw(env, tick, 0u, DOM_SPARSE, k0, k1, r);
      const uint64_t u = ((uint64_t)r[0] << 32) | r[1];
      int K = 0;
      while (K < OR_SPARSE_KMAX && u >= e->cdf[K]) ++K;
      uint32_t hits[OR_SPARSE_KMAX], j = 1;
      for (int got = 0; got < K;) {
        draw(env, tick, j++, DOM_SPARSE, k0, k1, r);
        const uint32_t pos = bounded64(r[0], r[1], n_cells);
        int dup = 0;
        for (int h = 0; h < got; ++h) dup |= hits[h] == pos;
        if (dup) continue;
        hits[got++] = pos;
        const int zd = (int)(pos / (uint32_t)e->plane_stride), off = (int)(pos % (uint32_t)e->plane_stride);
        if (off >= w->H * w->W) continue; /* padding byte */
        uint8_t* cellp = g + (size_t)off * w->L + w->dyn_layers[zd];
        if (w->rule[*cellp] == OR_RULE_SPAWN)
          *cellp = (uint8_t)w->choices[bounded64(r[2], r[3], (uint32_t)w->n_choices)];
      }
    } else
    /* entity sweep in C order (y, x, z) over the live map, environment.py:88-91 */
    for (int y = 0; y < w->H; ++y)
      for (int x = 0; x < w->W; ++x)
        for (int z = 0; z < w->L; ++z) {
          uint8_t* cellp = g + ((size_t)y * w->W + x) * w->L + z;
          const int kind = *cellp, rule = w->rule[kind];
          if (rule == 0 || kind == w->agent_kind) continue; /* `not isinstance(x, Agent)`, environment.py:90 */
          if (rule == OR_RULE_BECOME) { /* Beam.transition, cleanup/agents.py:199-204 */
            *cellp = w->rule_to[kind];
            continue;
          }
          if (rule == OR_RULE_BECOME_IF) { /* Pollution / Apple, cleanup/entities.py:64-71,100-105 */
            const int other = g[((size_t)y * w->W + x) * w->L + w->rule_layer[kind]];
            if ((w->rule_mask[kind] >> other) & 1u) *cellp = w->rule_to[kind];
            continue;
          }
          int zd = zd_of_layer(w, z);
          uint32_t pcell = (uint32_t)(zd * e->plane_stride + y * w->W + x);
          const double prob = rule == OR_RULE_SPAWN ? w->p : w->rule_prob[kind];
          if (e->rng_mode == OR_RNG_NUMPY) {
            double u = np_random_sample(mt); /* entities.py:73; cleanup/entities.py:52,86 */
            if (tu) tu[pcell] = u;
            if (u < prob) {
              if (rule == OR_RULE_SPAWN) {
                uint32_t ci = np_randint(mt, (uint32_t)w->n_choices); /* np.random.choice, entities.py:75 */
                if (tc) tc[pcell] = (uint8_t)ci;
                *cellp = (uint8_t)w->choices[ci];
              } else {
                *cellp = w->rule_to[kind];
              }
            } else if (rule == OR_RULE_SPAWN_TO && w->rule_else[kind] != OR_NO_KIND) {
              *cellp = w->rule_else[kind]; /* iowa Deck.draw: the no-loss value state */
            }
          } else {
            uint32_t r[4];
            const uint64_t thr = !(prob > 0.0) ? 0 : (prob >= 1.0 ? (1ull << 53) : (uint64_t)ceil(ldexp(prob, 53)));
            /* exact Bernoulli(thr / 2^53) in two levels (sb_core.cuh "Counter-based Bernoulli draw") */
            uint32_t n2 = 0;
            int lg = 0;
            if (thr) {
              const uint32_t hi = (uint32_t)((thr - 1) >> 45);
              for (n2 = 1; n2 < hi + 1u; n2 <<= 1) ++lg;
            }
            draw(env, tick, pcell >> 4, DOM_SWEEP, k0, k1, r);
            uint32_t b = pcell & 15, r8 = (r[b >> 2] >> (8 * (b & 3))) & 0xFF;
            int hit = r8 < n2; /* level 1 */
            if (hit) {         /* level 2: 64 fresh bits below thr << (19 - log2 n2) */
              draw(env, tick, pcell, DOM_CELL, k0, k1, r);
              hit = ((((uint64_t)r[0] << 32) | r[1]) <= (thr << (19 - lg)) - 1ull);
            }
            if (!hit) {
              if (rule == OR_RULE_SPAWN_TO && w->rule_else[kind] != OR_NO_KIND) *cellp = w->rule_else[kind];
              continue;
            }
            if (rule == OR_RULE_SPAWN) *cellp = (uint8_t)w->choices[bounded64(r[2], r[3], (uint32_t)w->n_choices)];
            else *cellp = w->rule_to[kind];
          }
        }
    /* agents in list order, agents/agent.py:155-173 */
    for (int a = 0; a < w->A; ++a) {
      const size_t row = (size_t)n * w->A + a;
      int cell = e->pos[row], y = cell / w->W, x = cell % w->W;
      int slot = e->capacity > 0 ? (e->idx + t) % e->capacity : 0;
      float* obs = e->capacity > 0 ? e->states + (row * e->capacity + slot) * e->obs_size : obs_local;
      observe(w, g, y, x, obs); /* pov(): state BEFORE the move */
      if (w->obs_flag) obs[e->obs_size - 1] = (float)e->flags[row]; /* TagAgent.pov, tag/agents.py:55-65 */
      int action;
      if (e->rng_mode == OR_RNG_NUMPY) {
        action = (int)np_randint(mt, (uint32_t)w->n_actions); /* RandomModel, base_model.py:111 */
      } else {
        uint32_t r[4];
        draw(env, tick, (uint32_t)a, DOM_ACTION, k0, k1, r);
        action = (int)bounded64(r[0], r[1], (uint32_t)w->n_actions);
      }
      if (e->record) e->tape_a[((size_t)(t_base + t) * e->N + n) * w->A + a] = action;
      if (w->uses_facing) {
        if (w->action_facing[action] >= 0) e->facing[row] = w->action_facing[action]; /* agents/agent.py:199-211 */
        const int bk = w->action_beam_kind[action], f = e->facing[row] & 3;
        if (bk != OR_NO_KIND) /* CleanupAgent.spawn_beam, cleanup/agents.py:93-145 */
          for (int b = 0; b < w->n_beam; ++b) {
            int by = y + w->beam_dy[f][b], bx = x + w->beam_dx[f][b];
            if (by < 0 || by >= w->H || bx < 0 || bx >= w->W) continue;
            uint8_t* bp = g + ((size_t)by * w->W + bx) * w->L + w->beam_layer;
            if (!((w->beam_block_mask >> *bp) & 1u)) *bp = (uint8_t)bk;
          }
      }
      int ny = y + w->dy[action], nx = x + w->dx[action];
      double reward = 0.0; /* the Python float the reference adds to world.total_reward; the ring stores float32 */
      if (ny >= 0 && ny < w->H && nx >= 0 && nx < w->W) {
        uint8_t* tp = g + ((size_t)ny * w->W + nx) * w->L + w->agent_layer;
        if (w->reward_all_layers) { /* observe_all_layers, cleanup/agents.py:160-170 */
          for (int z = 0; z < w->L; ++z) reward += w->value[g[((size_t)ny * w->W + nx) * w->L + z]];
        } else {
          reward = w->value[*tp]; /* reward = target.value, agent.py:219-220 */
        }
        if (w->passable[*tp]) { /* Gridworld.move, gridworld.py:95-122: the agent object itself moves */
          *tp = g[((size_t)y * w->W + x) * w->L + w->agent_layer];
          g[((size_t)y * w->W + x) * w->L + w->agent_layer] = (uint8_t)w->default_kind;
          e->pos[row] = ny * w->W + nx;
          y = ny, x = nx;
        }
      }
      if (w->touch_swap && e->flags[row]) { /* TagAgent.act, tag/agents.py:84-99; Location.adjacent order */
        static const int oy[4] = {-1, 0, 1, 0}, ox[4] = {0, 1, 0, -1};
        for (int k = 0; k < 4; ++k) {
          int qy = y + oy[k], qx = x + ox[k];
          if (qy < 0 || qy >= w->H || qx < 0 || qx >= w->W) continue;
          int b = -1;
          for (int i = 0; i < w->A; ++i)
            if (e->pos[n * w->A + i] == qy * w->W + qx) b = i;
          if (b >= 0 && !e->flags[n * w->A + b]) {
            e->flags[row] = 0;
            e->flags[n * w->A + b] = 1;
            g[((size_t)y * w->W + x) * w->L + w->agent_layer] = (uint8_t)w->agent_kind;
            g[((size_t)qy * w->W + qx) * w->L + w->agent_layer] = (uint8_t)w->agent_kind_alt;
            break;
          }
        }
      }
      if (w->reward_own) /* tag/agents.py:101-105: by the agent's own state after the exchange */
        reward = w->value[g[((size_t)y * w->W + x) * w->L + w->agent_layer]];
      /* agent.py:172 (cleanup adds the reward a second time in act(): cleanup/agents.py:172) */
      e->total[n] += reward * (double)(w->reward_count > 0 ? w->reward_count : 1);
      if (e->capacity > 0) {          /* Buffer.add, buffers.py:46-63; position AFTER the move */
        size_t at = (size_t)slot * rows + row;
        e->actions[at] = action;
        e->rewards[at] = (float)reward;
        e->dones[at] = 0.0f; /* world.is_done is False during the episode */
        e->positions[2 * at] = y;
        e->positions[2 * at + 1] = x;
      }
    }
  }
}

/* contiguous ranges of instances on plain pthreads (instances are independent) */
typedef struct {
  or_env_t* e;
  int64_t n0, n1;
  int T, do_reset;
} job_t;

static void* job_main(void* arg) {
  job_t* j = (job_t*)arg;
  if (j->do_reset) reset_range(j->e, j->n0, j->n1);
  else
    for (int64_t n = j->n0; n < j->n1; ++n) step_env(j->e, n, j->T, 0);
  return 0;
}

static void run_parallel(or_env_t* e, int T, int do_reset) {
  int nt = e->threads;
  if (nt > e->N) nt = (int)e->N;
  if (nt <= 1) {
    job_t j = {e, 0, e->N, T, do_reset};
    job_main(&j);
    return;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * nt);
  for (int i = 0; i < nt; ++i) {
    jobs[i].e = e;
    jobs[i].n0 = e->N * i / nt;
    jobs[i].n1 = e->N * (i + 1) / nt;
    jobs[i].T = T;
    jobs[i].do_reset = do_reset;
    pthread_create(&th[i], 0, job_main, &jobs[i]);
  }
  for (int i = 0; i < nt; ++i) pthread_join(th[i], 0);
  free(th), free(jobs);
}

void or_reset(or_env_t* e) { run_parallel(e, 0, 1); }

/* T x Environment.take_turn for every instance (instance-major: one world stays in cache) */
void or_step(or_env_t* e, int T) {
  const or_world_t* w = &e->w;
  if (e->record) {
    const size_t tape_stride = (size_t)w->n_dyn * e->plane_stride;
    free(e->tape_u), free(e->tape_c), free(e->tape_a);
    size_t cnt = (size_t)T * e->N * tape_stride;
    e->tape_u = (double*)malloc(cnt * sizeof(double));
    e->tape_c = (uint8_t*)malloc(cnt);
    e->tape_a = (int32_t*)malloc((size_t)T * e->N * w->A * sizeof(int32_t));
    for (size_t i = 0; i < cnt; ++i) e->tape_u[i] = NAN;
    memset(e->tape_c, 255, cnt);
    e->tape_T = T;
  }
  run_parallel(e, T, 0);
  e->tick += (uint32_t)T;
  if (e->capacity > 0) {
    e->idx = (e->idx + T) % e->capacity;
    e->size = e->size + T < e->capacity ? e->size + T : e->capacity;
  }
}

void or_add_empty(or_env_t* e) { /* buffers.py:65-69 */
  e->idx = (e->idx + e->n_frames - 1) % e->capacity;
  e->size = e->size + 1 < e->capacity ? e->size + 1 : e->capacity;
}

/* accessors for the ctypes wrapper */
uint8_t* or_grid(or_env_t* e) { return e->grid; }
int32_t* or_pos(or_env_t* e) { return e->pos; }
int32_t* or_facing(or_env_t* e) { return e->facing; }
int32_t* or_flags(or_env_t* e) { return e->flags; }
double* or_total(or_env_t* e) { return e->total; }
float* or_states(or_env_t* e) { return e->states; }
int64_t* or_actions(or_env_t* e) { return e->actions; }
float* or_rewards(or_env_t* e) { return e->rewards; }
float* or_dones(or_env_t* e) { return e->dones; }
int64_t* or_positions(or_env_t* e) { return e->positions; }
double* or_tape_u(or_env_t* e) { return e->tape_u; }
uint8_t* or_tape_c(or_env_t* e) { return e->tape_c; }
int32_t* or_tape_a(or_env_t* e) { return e->tape_a; }
int32_t* or_tape_placement(or_env_t* e) { return e->tape_pl; }
int or_idx(or_env_t* e) { return e->idx; }
int or_size(or_env_t* e) { return e->size; }
void or_set_tick(or_env_t* e, uint32_t tick) { e->tick = tick; }
void or_set_sparse_cdf(or_env_t* e, const uint64_t* cdf) { memcpy(e->cdf, cdf, sizeof(e->cdf)); }
/* global instance ids (Philox counter word 0) of a SAMPLE of a larger batch: lets a test replay instances
   {i0, i1, ...} of a 2^20-instance launch without stepping the rest */
void or_set_env_ids(or_env_t* e, const uint64_t* ids) {
  free(e->env_ids);
  e->env_ids = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)e->N);
  memcpy(e->env_ids, ids, sizeof(uint64_t) * (size_t)e->N);
}
int or_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* numpy-legacy stream probes (validated against numpy in tests/test_coracle.py) */
void or_np_probe(uint32_t seed, int n_uniform, double* u_out, int n_randint, uint32_t bound, uint32_t* r_out,
                 int perm_n, int perm_k, int32_t* p_out) {
  mt_t s;
  mt_seed(&s, seed);
  for (int i = 0; i < n_uniform; ++i) u_out[i] = np_random_sample(&s);
  for (int i = 0; i < n_randint; ++i) r_out[i] = np_randint(&s, bound);
  if (perm_n > 0) {
    int32_t* scratch = (int32_t*)malloc(sizeof(int32_t) * perm_n);
    np_permutation_head(&s, perm_n, perm_k, p_out, scratch);
    free(scratch);
  }
}

void or_philox_probe(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
