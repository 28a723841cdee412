// sb_internal.h -- declarations shared by the translation units of libsorrel_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/sorrel_b200.h"
#include "sb_core.cuh"

#define SB_THREADS 128
#define SB_QCAP 96       // per-warp spawn-candidate queue entries
#ifndef SB_PAR
#define SB_PAR 4  // agents of one instance observed / moved at once by a warp (when independent)
#endif
#ifndef SB_ROWWISE_MAX_LPE
#define SB_ROWWISE_MAX_LPE 8  // groups up to this width gather observation windows row by row (0: never)
#endif
#ifndef SB_COLGATHER
#define SB_COLGATHER 1  // several staged planes: gather observation windows by row segments (scatter_cols_multi)
#endif
#define SB_OBS_ALL (-2)  // SbStepArgs.obs_agent: observe every agent (sb_observe)

// kernel RNG variants (template parameter)
#define SB_K_TAPE 0
#define SB_K_PHILOX_SIMD 1
#define SB_K_PHILOX_GENERIC 2
#define SB_K_SPARSE 3

namespace sb {

struct StepParams {
  SbStepArgs a;
  SbBufferDesc buf;
  int32_t has_buf;
  int32_t write_back;
  uint8_t* grid;
  int32_t* agent_pos;
  double* total_reward;
  // shared-memory layout (struct-of-arrays over the block's instances), computed on the host in
  // sb_step.cu: launch_one.  Per-instance strides are padded to 64 (mod 128) bytes so the two groups
  // served by one quarter-warp of a 128-bit shared access hit disjoint banks.
  int32_t q_off, planes_off, pos_off, stage_off, hits_off;  // block-level byte offsets
  int32_t staged;                                 // planes staged in shared memory (else used in place)
  int32_t plane_env, stage_env;                   // per-instance strides of planes / staging rows
  int32_t stage_row, par;                         // one staging row; agent-parallel path enabled
  int32_t agent_words;                            // 32-bit words of per-instance agent state (yx, facing, reward, actions)
  int32_t act_turns;                              // turns whose actions are prepared at once (action window rows)
  uint32_t inv_A;                                 // ceil(65536 / A)
  uint32_t rk[2];                                 // {seed_lo, seed_hi}: the stream words of sb_philox_rk
  int32_t prefetch_blocks;                        // resident blocks of the launch (L2 prefetch distance), 0 = none
};

int note_cuda_error(cudaError_t e);  // records the error text for sb_last_cuda_error, returns SB_ECUDA
// after a launch / runtime call: SB_OK, or SB_ECUDA with the (single) cudaGetLastError() result recorded
inline int check_launch() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SB_OK : note_cuda_error(e);
}
void count_launch();

int make_dev(const SbWorldDesc* desc, SbDev* out);  // validates + derives the device view
// plan != nullptr: fill it and do not launch
int launch_step(const SbDev& d, const StepParams& p, cudaStream_t s, SbStepPlan* plan = nullptr);
int launch_observe(const SbDev& d, const uint8_t* grid, const int32_t* agent_pos, float* out, cudaStream_t s);

}  // namespace sb
