// sb_internal.h -- declarations shared by the translation units of libsorrel_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/sorrel_b200.h"
#include "sb_core.cuh"

#define SB_THREADS 128
#define SB_QCAP 96       // per-warp spawn-candidate queue entries
#define SB_OBS_ALL (-2)  // SbStepArgs.obs_agent: observe every agent (sb_observe)

// kernel RNG variants (template parameter)
#define SB_K_TAPE 0
#define SB_K_PHILOX_SIMD 1
#define SB_K_PHILOX_GENERIC 2

namespace sb {

struct StepParams {
  SbStepArgs a;
  SbBufferDesc buf;
  int32_t has_buf;
  int32_t write_back;
  uint8_t* grid;
  int32_t* agent_pos;
  double* total_reward;
  // shared-memory layout, computed once on the host (sb_step.cu: launch_one)
  int32_t lut_bytes, q_bytes, env_smem, stage_off;
};

int note_cuda_error();  // records cudaGetLastError() text for sb_last_cuda_error, returns SB_ECUDA
void count_launch();

int make_dev(const SbWorldDesc* desc, SbDev* out);  // validates + derives the device view
int launch_step(const SbDev& d, const StepParams& p, cudaStream_t s);
int pick_lpe(const SbDev& d);

}  // namespace sb
