/*
 * sorrel_b200.h -- C ABI of the B200-native batched gridworld engine.
 *
 * One shared library (libsorrel_b200.so, sm_100a only) implements the environment-step /
 * one-hot observation / replay-buffer hot path of social-ai-uoft/sorrel v1.4.0 for N
 * independent world instances per call.  The reference has no FFI: its boundary is a set of
 * Python methods.  Each entry point below names the reference method(s) it replaces
 * (file:line relative to the reference root); INTEGRATION.md shows the ctypes binding.
 *
 * ABI history: v1 treasurehunt (SB_RULE_SPAWN, one-hot / RGB crops, ring, stats); v2 column rules, facings and beams,
 * all-layer rewards, positional code, item groups (examples/cleanup); v3 value states and float64 rewards
 * (examples/iowa); v4 agents with a second kind (examples/tag); v5 launch tuning in SbStepArgs (the library reads no
 * environment variables), sb_step_plan, device error word for the ring gathers.  Fields are only ever appended.
 *
 * Conventions
 *  - plain C types, pointers and sizes only; no C++/torch types.
 *  - every pointer named `d_*` or living in a struct documented "device" is a DEVICE pointer on
 *    the current CUDA device, owned by the caller (PyTorch); the library allocates nothing
 *    persistent, keeps no global state and never frees caller memory.
 *  - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*).
 *  - return value: 0 (SB_OK) or a negative SbStatus; never throws, never exits.
 *  - kernel-side invariant violations (an agent stepping off the grid, a replay tape that does
 *    not line up with the cells that drew) set bits in the caller-provided device error word.
 */
#ifndef SORREL_B200_H
#define SORREL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 5

#define SB_MAX_KINDS 32    /* entity classes per world (kind ids are uint8 < SB_MAX_KINDS) */
#define SB_MAX_CHANNELS 32 /* observation channels C = len(entity_list) */
#define SB_MAX_LAYERS 8
#define SB_MAX_CHOICES 8 /* kinds a spawn rule can choose from */
#define SB_MAX_ACTIONS 8
#define SB_MAX_AGENTS 64 /* agents per world instance */
#define SB_PAD_KIND 0xFF /* value of the padding bytes after each H*W plane */
#define SB_NO_KIND 0xFF  /* "no kind" in kind-valued tables (action_beam_kind) */
#define SB_MAX_BEAM_CELLS 16 /* cells one beam covers (cleanup: 3 * beam_radius = 9) */
#define SB_MAX_ITEM_GROUPS 4 /* item groups placed by sb_reset_ex before the agents */

typedef enum SbStatus {
  SB_OK = 0,
  SB_EINVAL = -1,       /* NULL / negative / inconsistent argument            -> ValueError  */
  SB_EUNSUPPORTED = -2, /* shape outside what the kernels implement            -> ValueError  */
  SB_EALIGN = -3,       /* pointer not 16-byte aligned                         -> ValueError  */
  SB_ECUDA = -4,        /* CUDA runtime / launch error (sb_last_cuda_error)    -> RuntimeError */
  SB_ENODEVICE = -5,    /* no sm_100 device visible                            -> RuntimeError */
  SB_EVISION = -6       /* vision radius > (H-1)/2 or (W-1)/2: the reference silently returns a
                           wrong-shaped crop there (visual_field.py:78-87)     -> ValueError  */
} SbStatus;

/* bits of the device error word */
#define SB_ERR_OFFGRID 1u   /* agent moved outside [0,H)x[0,W): numpy index wrap / IndexError in
                               the reference (gridworld.py:95-122 with a negative index) */
#define SB_ERR_TAPE_ALIGN 2u /* replay: a cell drew a uniform where the tape holds NaN */
#define SB_ERR_BAD_ACTION 4u /* action outside [0, num_actions) */
#define SB_ERR_BAD_KIND 8u   /* grid byte >= num_kinds */
#define SB_ERR_BAD_INDEX 16u /* Buffer.sample gather: row outside [0, rows) or start outside [0, capacity - n_frames):
                                numpy raises IndexError there (buffers.py:112-119) */

/* observation transforms */
#define SB_OBS_SUM 0 /* OneHotObservationSpec / raw table sums (observation_spec.py:175-205) */
#define SB_OBS_RGB 1 /* RGBObservationSpec: clip(sum, 0, 255) / 255 in f64 (observation_spec.py:386-483) */

/* transition rules (per kind) */
#define SB_RULE_NONE 0
#define SB_RULE_SPAWN 1 /* Bernoulli(spawn_prob) -> uniform choice among spawn_choice_kinds;
                           EmptyEntity.transition, examples/treasurehunt/entities.py:69-85 */
/* "column rules": a transition may read the other layers of its own (y, x) column, which is all the
   reference's shipped rule sets do.  The sweep visits the planes in ascending layer order, so a rule
   sees lower layers after, higher layers before their own transition of the turn -- exactly the
   C-order np.ndenumerate sweep of Environment.take_turn (environment.py:88-91). */
#define SB_RULE_SPAWN_TO 2  /* Bernoulli(kind_rule_prob[k]) -> kind_rule_to[k], else kind_rule_else[k] (SB_NO_KIND =
                               unchanged), one uniform, no choice draw: River -> Pollution, AppleTree -> Apple
                               (examples/cleanup/entities.py:43-54,74-90); Deck.transition / draw with the mutable
                               value folded into the kind id (examples/iowa/entities.py:33-69) */
#define SB_RULE_BECOME_IF 3 /* the cell of the same column on layer kind_rule_layer[k] holds a kind of
                               kind_rule_mask[k] -> kind_rule_to[k]: Pollution under a CleanBeam -> River,
                               Apple under an agent -> AppleTree (cleanup/entities.py:57-71,93-105) */
#define SB_RULE_BECOME 4    /* unconditionally -> kind_rule_to[k]: Beam.transition (cleanup/agents.py:199-204)
                               with the per-entity turn_counter folded into the kind id */

/* sources of randomness / actions for sb_step */
#define SB_RNG_TAPE 0   /* replay recorded numpy draws: bit-exact to the reference */
#define SB_RNG_PHILOX 1 /* counter-based Philox4x32-10 over (seed, global env id, tick, index, domain): one random
                           byte per cell per turn (dense; mirrors the reference's draw-per-cell loop) */
#define SB_RNG_SPARSE 2 /* same generator, sparse sampling of the same Bernoulli field: the number of
                           spawning cells K ~ Binomial(cells, p) is drawn first (inverse CDF on
                           SbWorldDesc.spawn_cdf), then K distinct uniform cells.  Identical in
                           distribution to per-cell draws; needs spawn_cdf_len > 0 */
#define SB_POLICY_TAPE 0    /* actions supplied by the caller */
#define SB_POLICY_UNIFORM 1 /* in-kernel uniform policy == RandomModel, models/base_model.py:107 */

/* The keyed stream: draw(env, tick, index, domain) = Philox4x32-10((env, tick ^ seed_lo, index, domain ^ seed_hi), K)
   with a FIXED key K (sb_core.cuh: SB_PHILOX_K0 / K1): word0 = global env id, word1 = tick, word2 = index, word3 =
   "domain"; the 64-bit seed is folded into words 1 and 3, so the round keys are compile-time constants. */
#define SB_DOM_SWEEP 0x53570000u  /* index = cell >> 4 : 16 cells per call, one byte each */
#define SB_DOM_CELL 0x43450000u   /* index = cell      : 64 level-2 bits + 64 choice bits (cells that passed level 1) */
#define SB_DOM_ACTION 0x41430000u /* index = agent     : 64 action bits */
#define SB_DOM_RESET 0x52530000u  /* index = try       : 64 placement bits */
#define SB_DOM_SPARSE 0x53500000u /* index 0: 64 bits -> K;  index j >= 1: 64 position bits + 64 choice bits */
#define SB_SPARSE_KMAX 64         /* entries of the Binomial CDF table */

/*
 * World family description (host memory, copied by value at every call).
 *
 * Replaces the per-object state of Gridworld / Entity (sorrel/worlds/gridworld.py:36-65,
 * sorrel/entities/entity.py:9-68): a *kind* is an entity class; its `value`, `passable`,
 * transition rule and appearance row live in tables indexed by a uint8 kind id.
 *
 * Grid memory: uint8 d_grid[num_envs][num_dyn_layers][plane_stride], plane_stride =
 * sb_plane_stride(H, W) = H*W rounded up to 16, padding bytes = SB_PAD_KIND.  Only layers that
 * can change are stored per instance ("dynamic"); static layers are read from d_template.
 */
typedef struct SbWorldDesc {
  int32_t abi_version; /* SB_ABI_VERSION */
  int32_t num_envs;    /* N: world instances owned by this rank */
  int32_t num_agents;  /* A per instance (Environment.agents, environment.py:31-54) */
  int32_t height, width, num_layers; /* Gridworld(height,width,layers,...) gridworld.py:36 */
  int32_t num_dyn_layers;
  int32_t dyn_layers[SB_MAX_LAYERS]; /* layer index of each dynamic plane, ascending */
  int32_t agent_layer;               /* layer the agents live on (must be dynamic) */
  int32_t num_kinds;
  int32_t num_channels;              /* C */
  int32_t vision;                    /* v; S = 2v+1 (ObservationSpec.vision_radius) */
  int32_t fill_kind;                 /* ObservationSpec.fill_entity_kind ("Wall") */
  int32_t agent_kind;
  int32_t default_kind;              /* Gridworld.default_entity: what a vacated cell becomes */
  int32_t num_actions;               /* ActionSpec.n_actions, action/action_spec.py:4-47 */
  int8_t action_dy[SB_MAX_ACTIONS];  /* MovingAgent.movement agents/agent.py:187-213 */
  int8_t action_dx[SB_MAX_ACTIONS];
  int32_t num_spawn_choices;
  uint8_t spawn_choice_kinds[SB_MAX_CHOICES]; /* order of the choice array, entities.py:77-81 */
  double spawn_prob;                 /* world.spawn_prob, treasurehunt/world.py:29 */
  uint64_t env_id_base;              /* global id of instance 0 (rank offset; shard invariance) */
  float kind_value[SB_MAX_KINDS];    /* Entity.value */
  uint8_t kind_passable[SB_MAX_KINDS];
  uint8_t kind_rule[SB_MAX_KINDS];
  const double* d_appearance; /* device [num_kinds][C]: ObservationSpec.entity_map rows */
  const double* h_appearance; /* HOST copy of the same table (lets the library pick the one-hot
                                 fast path without a device read) */
  const uint8_t* d_template;  /* device [num_layers][plane_stride]: world after
                                 populate_environment, before agents are placed */
  int32_t static_layers_invisible; /* 1 when every kind on a static layer has an all-zero
                                      appearance row (treasurehunt Sand, entities.py:23-31) */
  int32_t spawn_cdf_len;           /* 0, or SB_SPARSE_KMAX when spawn_cdf is filled (SB_RNG_SPARSE) */
  /* spawn_cdf[k] = floor(2^64 * P(X <= k)), X ~ Binomial(num_dyn_layers * plane_stride, P/2^53),
     P = ceil(spawn_prob * 2^53), saturated at 2^64-1; computed exactly (rational arithmetic) by the
     caller, see sorrel_b200/engine.py: binomial_cdf_u64.  K = #{k : u64 >= spawn_cdf[k]}. */
  uint64_t spawn_cdf[SB_SPARSE_KMAX];
  int32_t obs_transform; /* SB_OBS_*: applied to the per-cell sum over layers before the f32 store */
  int32_t reserved1;
  /* ---- ABI v2: column rules, facings + beams, all-layer rewards, positional code (examples/cleanup) ---- */
  double kind_rule_prob[SB_MAX_KINDS];   /* SB_RULE_SPAWN_TO */
  uint8_t kind_rule_to[SB_MAX_KINDS];    /* SB_RULE_SPAWN_TO / BECOME_IF / BECOME */
  int8_t kind_rule_layer[SB_MAX_KINDS];  /* SB_RULE_BECOME_IF: a dynamic layer */
  uint32_t kind_rule_mask[SB_MAX_KINDS]; /* SB_RULE_BECOME_IF: bit j set <=> kind j satisfies the test */
  int8_t action_facing[SB_MAX_ACTIONS];  /* facing (0 N, 1 E, 2 S, 3 W) the agent takes with the action, -1 =
                                            unchanged (MovingAgent.movement, agents/agent.py:199-211) */
  uint8_t action_beam_kind[SB_MAX_ACTIONS]; /* kind the action's beam places, SB_NO_KIND = no beam
                                               (CleanupAgent.spawn_beam, cleanup/agents.py:93-145) */
  int32_t uses_facing;    /* 1: action_facing / action_beam_kind are meaningful and SbStepArgs.d_agent_facing
                             must be set; 0: plain MovingAgent (treasurehunt) */
  int32_t beam_layer;     /* dynamic layer the beams live on */
  int32_t num_beam_cells; /* <= SB_MAX_BEAM_CELLS */
  int8_t beam_dy[4][SB_MAX_BEAM_CELLS]; /* beam cells relative to the agent, per facing */
  int8_t beam_dx[4][SB_MAX_BEAM_CELLS];
  uint32_t beam_block_mask;  /* kinds on the beam layer a beam does not replace (Wall) */
  int32_t reward_all_layers; /* reward = sum of Entity.value over all layers of the target column
                                (world.observe_all_layers, cleanup/agents.py:160-170) instead of the agent layer's */
  int32_t reward_count;      /* times the reward is added to world.total_reward per agent turn: 1, or 2 for
                                cleanup (cleanup/agents.py:172 and agents/agent.py:172); 0 means 1 */
  int32_t obs_extra;         /* floats appended to every observation after the C*S*S crop */
  const float* d_obs_extra;  /* device [H*W][obs_extra], row of the agent's cell: positional code
                                (CleanupObservation.observe cleanup/agents.py:52-62, observation/embedding.py:8-44),
                                computed by the caller (numpy, f64 -> f32 like Buffer.states) */
  /* ---- ABI v3: value states (examples/iowa) ---- */
  uint8_t kind_rule_else[SB_MAX_KINDS]; /* SB_RULE_SPAWN_TO: kind when the draw fails; SB_NO_KIND = unchanged */
  int32_t use_value64;                  /* 1: rewards are kind_value64 (Python floats such as 1.1 or -1.4): the ring
                                           stores (float)value, world.total_reward accumulates the double */
  int32_t reserved2;
  double kind_value64[SB_MAX_KINDS];
  /* ---- ABI v4: agents with a binary state shown as a second kind (examples/tag: "It" / "NotIt") ---- */
  int32_t agent_kind_alt; /* -1, or the kind an agent whose flag is set is drawn with (TagAgent.it setter,
                             examples/tag/agents.py:44-52); agent_kind is the kind of the others */
  int32_t touch_swap;     /* 1: after its move a flagged agent hands the flag to the first unflagged agent among the
                             neighbours (y-1,x), (y,x+1), (y+1,x), (y,x-1) (TagAgent.act tag/agents.py:84-99,
                             Location.adjacent location.py:144) */
  int32_t reward_own;     /* 1: reward = kind_value[own kind after the exchange] instead of the target's value
                             (tag/agents.py:101-105) */
  int32_t obs_flag;       /* 1: the flag is appended to every observation as one more float (TagAgent.pov :55-65) */
  int32_t toroidal;       /* engine option, NOT a reference behaviour (the reference pads crops with the fill kind and
                             has no toroidal worlds): 1 = observation windows and moves wrap around
                             the map edges.  One-hot worlds keep the fast observation path (bf16 staging row, 128-bit
                             stores): window indices are folded where the cell is looked up */
  int32_t reserved3;
  int32_t* d_agent_flag;  /* device int32 [N][A]: the flags; read by sb_reset / sb_reset_ex (they survive
                             Environment.reset like the agent attribute does), written back by sb_step; required iff
                             agent_kind_alt >= 0 */
} SbWorldDesc;

/* One group of items placed on distinct cells at reset, before the agents (cleanup/env.py:126-132:
 * `initial_apples` Apples on the orchard cells). */
typedef struct SbItemGroup {
  int32_t kind, layer, count, n_cells;
  const int32_t* d_cells;     /* device int32 [n_cells] plane cell index (y*W+x), reference order */
  const int32_t* d_placement; /* device int32 [N][count] indices into d_cells (replayed np.random.choice),
                                 or NULL to draw in-kernel: Philox(seed, env id, tick, ((g+1)<<24) + try, SB_DOM_RESET) */
} SbItemGroup;

/*
 * Batched replay ring: one reference `Buffer` (sorrel/buffers.py:11-44) per agent row,
 * rows = num_envs * num_agents, row r = env * A + agent.  `idx`/`size` are host scalars shared
 * by all rows (every agent adds exactly one experience per turn).
 *   states    float  [rows][capacity][obs_size]  (row-major per agent: frame stacks contiguous; or any other
 *                                                 (row, slot) stride pair, see states_row_stride below)
 *   actions   int64  [capacity][rows]
 *   rewards   float  [capacity][rows]
 *   dones     float  [capacity][rows]
 *   positions int64  [capacity][rows][2]         (extra_data["positions"], treasurehunt/env.py:100)
 */
typedef struct SbBufferDesc {
  int32_t capacity;
  int32_t obs_size; /* C*S*S + obs_extra (+ 1 with obs_flag) floats */
  int32_t n_frames;
  int32_t reserved;
  int64_t rows;
  float* d_states;
  int64_t* d_actions;
  float* d_rewards;
  float* d_dones;
  int64_t* d_positions; /* may be NULL */
  /* ---- ABI v5 ---- */
  uint32_t* d_error;    /* optional device error word: the gathers OR SB_ERR_BAD_INDEX into it instead of reading
                           outside the ring */
  /* Layout of `states`, in floats: element (row, slot, e) lives at row * states_row_stride + slot * states_slot_stride
     + e.  Both 0 = agent-major, the default documented above ([rows][capacity][obs]: row stride capacity * obs_size,
     slot stride obs_size; the F stacked frames of a transition are contiguous).  Slot-major ([capacity][rows][obs]: row
     stride obs_size, slot stride rows * obs_size) makes the rows ONE TURN writes contiguous in HBM -- whole DRAM pages
     instead of one 600-byte piece per 19 KB -- which is what multi-agent worlds want (profiles/README.md). */
  int64_t states_row_stride;
  int64_t states_slot_stride;
} SbBufferDesc;

/* sb_step modes */
#define SB_MODE_FUSED 0 /* per turn: sweep; for each agent in list order: observe, act, record */
#define SB_MODE_SPLIT 1 /* one launch = [sweep] -> [act agent act_agent] -> [observe agent obs_agent];
                           lets a live policy run between observe(k) and act(k) */

typedef struct SbStepArgs {
  int32_t mode;      /* SB_MODE_* */
  int32_t num_turns; /* T >= 1 turns per launch (FUSED); must be 1 for SPLIT */
  uint32_t tick0;    /* global turn counter of the first turn (Philox counter word 1) */
  int32_t slot0;     /* ring slot (Buffer.idx) the first turn writes */
  int32_t rng;       /* SB_RNG_* */
  int32_t policy;    /* SB_POLICY_* */
  uint64_t seed;     /* Philox key */
  int32_t do_sweep;  /* SPLIT: run the entity sweep in this launch */
  int32_t act_agent; /* SPLIT: agent index to act, or -1 */
  int32_t obs_agent; /* SPLIT: agent index to observe, or -1 */
  int32_t reserved;
  const int32_t* d_actions;  /* [T][N][A]   (policy TAPE / SPLIT act) */
  const double* d_spawn_u;   /* [T][N][Ldyn*H*W] f64, NaN where the reference drew nothing */
  const uint8_t* d_spawn_c;  /* [T][N][Ldyn*H*W] index into spawn_choice_kinds, 255 = none */
  float* d_obs_out;          /* optional [T][N][A][obs_size] (when no ring is attached) */
  float* d_reward_out;       /* optional [T][N][A] */
  int32_t* d_action_out;     /* optional [T][N][A] actions actually taken */
  uint8_t* d_drew_out;       /* optional [T][N][Ldyn*H*W] 1 where a cell consumed a uniform */
  uint32_t* d_error;         /* device error word (OR of SB_ERR_*) */
  const uint32_t* d_clock;   /* optional device uint32[2] = {tick0, slot0}: when set, the kernel reads the
                                two counters from device memory instead of the fields above, so that an
                                identical launch can be replayed from a CUDA graph (sb_clock_advance) */
  int32_t* d_agent_facing;   /* [N][A] MovingAgent.direction, read and written; required iff desc.uses_facing */
  /* ---- ABI v5: launch tuning, 0 = the library chooses (tests pin the variant they check; sb_step_plan reports it).
     These replace the SB_LPE / SB_STAGED / SB_NO_PAR environment variables v4 read at every launch. ---- */
  int32_t tune_lpe;    /* lanes that own one instance: 0 auto, or 2 / 4 / 8 / 16 / 32, or 128 = one block per instance
                          (16..64 agents, one-hot rows, byte-SIMD sweep; SB_EUNSUPPORTED where it does not apply) */
  int32_t tune_staged; /* instance planes staged in shared memory for the launch: 0 auto, 1 staged, -1 used in place
                          in global memory (one warp per instance: tune_lpe must be 0 or 32) */
  int32_t tune_par;    /* agent-parallel path of warp-per-instance worlds: 0 auto, -1 off */
  int32_t actions_u8;  /* 1: d_actions points to uint8 [T][N][A] instead of int32 (an ActionSpec has at most
                          SB_MAX_ACTIONS = 8 actions): a quarter of the host-to-device bytes of a host policy */
} SbStepArgs;

/* What sb_step would launch for (desc, args): filled by sb_step_plan without touching the device. */
typedef struct SbStepPlan {
  int32_t lpe;          /* lanes per instance */
  int32_t kernel_rng;   /* 0 replay, 1 byte-SIMD Philox sweep, 2 generic Philox sweep, 3 sparse sampler */
  int32_t fast_onehot;  /* 1: one stored plane + one-hot fast path (bf16 staging row, 128-bit stores) */
  int32_t staged;       /* planes staged in shared memory */
  int32_t par;          /* agent-parallel path enabled */
  int32_t threads, smem_bytes;
  int64_t blocks;
} SbStepPlan;

/* -- library / geometry ------------------------------------------------------------------ */
int sb_abi_version(void);
const char* sb_strerror(int status);
/* text of the last CUDA error seen by the calling thread (for SB_ECUDA) */
const char* sb_last_cuda_error(void);
/* number of CUDA kernels this library has launched in the calling process (bench "gpu_launches") */
uint64_t sb_launch_count(void);
int32_t sb_plane_stride(int32_t height, int32_t width);
/* bytes of d_grid for a descriptor */
size_t sb_grid_bytes(const SbWorldDesc* desc);
/* validate a descriptor without touching the device (shape limits, vision rule) */
int sb_validate(const SbWorldDesc* desc);

/* -- Environment.reset + populate_environment -------------------------------------------- *
 * environment.py:72-79 -> examples/treasurehunt/env.py:114-147: copy the template into every
 * instance, place the A agents on distinct cells of the spawnable list, zero total_reward.
 *   d_spawnable  int32 [n_spawnable]  plane cell index (y*W+x) on the agent layer, reference order
 *   d_placement  int32 [N][A] indices into the spawnable list (replayed np.random.choice), or
 *                NULL to draw them in-kernel from Philox (seed, env id, tick, SB_DOM_RESET)
 */
int sb_reset(const SbWorldDesc* desc, uint8_t* d_grid, int32_t* d_agent_pos, double* d_total_reward,
             const int32_t* d_spawnable, int32_t n_spawnable, const int32_t* d_placement,
             uint64_t seed, uint32_t tick, void* stream);

/* Same with item groups placed first (examples/cleanup/env.py:85-140): groups[g] host array, n_groups <=
 * SB_MAX_ITEM_GROUPS.  sb_reset(...) == sb_reset_ex(..., NULL, 0, ...). */
int sb_reset_ex(const SbWorldDesc* desc, uint8_t* d_grid, int32_t* d_agent_pos, double* d_total_reward,
                const int32_t* d_spawnable, int32_t n_spawnable, const int32_t* d_placement,
                const SbItemGroup* groups, int32_t n_groups, uint64_t seed, uint32_t tick, void* stream);

/* -- Environment.take_turn ---------------------------------------------------------------- *
 * environment.py:81-93 (entity sweep -> EmptyEntity.transition treasurehunt/entities.py:69-85;
 * then Agent.transition agents/agent.py:155-173 = pov/observe -> action -> act/move -> reward ->
 * Buffer.add) for all N instances, `num_turns` turns per launch.
 *   d_agent_pos    int32 [N][A] plane cell index of every agent
 *   d_total_reward f64   [N]    world.total_reward
 *   buf            ring the observations/experiences are written into, or NULL
 */
int sb_step(const SbWorldDesc* desc, const SbStepArgs* args, uint8_t* d_grid, int32_t* d_agent_pos,
            double* d_total_reward, const SbBufferDesc* buf, void* stream);

/* The launch sb_step(desc, args, ..., buf != NULL, ...) would make (variant, tiling, shared memory, grid); no device
 * work.  lpe 128 = the block-per-instance kernel of large many-agent worlds. */
int sb_step_plan(const SbWorldDesc* desc, const SbStepArgs* args, SbStepPlan* out);

/* Advance a device-resident {tick, slot} pair (see SbStepArgs.d_clock): tick += add_tick,
 * slot = (slot + add_slot) % capacity.  Graph-capturable companion of sb_step. */
int sb_clock_advance(uint32_t* d_clock, uint32_t add_tick, uint32_t add_slot, uint32_t capacity, void* stream);

/* -- ObservationSpec.observe -------------------------------------------------------------- *
 * observation/observation_spec.py:175-205 -> visual_field.py:9-101 for every (instance, agent):
 * d_obs_out float [N][A][C*S*S + obs_extra].  Pure function of the grid; does not modify state.
 */
int sb_observe(const SbWorldDesc* desc, const uint8_t* d_grid, const int32_t* d_agent_pos,
               float* d_obs_out, void* stream);

/* -- Buffer.add (buffers.py:46-63) for all rows at ring slot `slot` ------------------------ */
int sb_buffer_add(const SbBufferDesc* buf, int32_t slot, const float* d_obs, const int64_t* d_actions,
                  const float* d_rewards, const float* d_dones, const int64_t* d_positions,
                  void* stream);

/* -- Buffer.sample gather (buffers.py:112-124) --------------------------------------------- *
 * sample b reads frames [start_b, start_b + n_frames] of row row_b.  Index draws
 * (buffers.py:109-111) are the caller's.  Outputs: states/next_states float [B][F*obs],
 * actions int64 [B], rewards/dones float [B], valid double [B].
 */
int sb_buffer_sample(const SbBufferDesc* buf, const int64_t* d_rows, const int64_t* d_starts,
                     int32_t batch, float* d_states, int64_t* d_actions, float* d_rewards,
                     float* d_next_states, float* d_dones, double* d_valid, void* stream);

/* -- TransformerBuffer.sample_transformer gather (buffers.py:319-355) ----------------------------- *
 * like sb_buffer_sample, but the actions keep the time axis: d_action_seq int64 [B][F] =
 * actions[start .. start+F), d_next_action_seq float [B][F] = actions[start+1 .. start+F+1] cast to f32.
 */
int sb_buffer_sample_transformer(const SbBufferDesc* buf, const int64_t* d_rows, const int64_t* d_starts,
                                 int32_t batch, float* d_states, int64_t* d_action_seq,
                                 float* d_next_action_seq, float* d_next_states, float* d_dones,
                                 double* d_valid, void* stream);

/* -- Buffer.current_state (buffers.py:143-154): rows [idx-(F-1), idx) with ring wrap --------- *
 * d_out float [rows][F-1][obs_size]
 */
int sb_buffer_current_state(const SbBufferDesc* buf, int32_t idx, float* d_out, void* stream);

/* Same with the write head read from device memory: idx = d_clock[1] (see SbStepArgs.d_clock), so that a captured
 * CUDA graph of a whole split-phase turn can stack frames for its policy without a host-side index. */
int sb_buffer_current_state_dev(const SbBufferDesc* buf, const uint32_t* d_clock, float* d_out, void* stream);

/* -- per-episode reward statistics (feeds Logger.record_turn, utils/logging.py:48-72) ------- *
 * d_out double[5] = {sum, sum of squares, min, max, count} over d_total_reward[0..n); the only
 * quantity that is all-reduced across ranks (NCCL, by the caller).  d_scratch: >= sb_stats_scratch_bytes().
 */
size_t sb_stats_scratch_bytes(void);
int sb_reward_stats(const double* d_total_reward, int64_t n, double* d_out, void* d_scratch,
                    void* stream);

/* Combine the per-rank partials of sb_reward_stats after the all-gather: d_parts double [num_ranks][5] -> d_out
 * double[5] (sums added in rank order, min / max folded, counts added).  One tiny launch, stream-ordered. */
int sb_stats_combine(const double* d_parts, int32_t num_ranks, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SORREL_B200_H */
