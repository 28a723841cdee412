// sb_step.cu -- the environment-step kernel (entity sweep + sequential agents + observation +
// replay-ring write), sm_100a.
//
// Replaces Environment.take_turn (sorrel/environment.py:81-93) for N independent instances:
//   sweep   : EmptyEntity.transition           examples/treasurehunt/entities.py:69-85
//   observe : OneHotObservationSpec.observe    observation/observation_spec.py:175-205
//             -> visual_field                  observation/visual_field.py:9-101
//   act     : MovingAgent.act / Gridworld.move agents/agent.py:215-225, worlds/gridworld.py:95-122
//   record  : Agent.add_memory -> Buffer.add   agents/agent.py:113-136, buffers.py:46-63
// and, in the replay / generic variants only (`EXT`), the rule extensions of the other shipped examples:
//   column rules, facings + beams, all-layer rewards, positional code   examples/cleanup/{entities,agents}.py
//   value states, float64 rewards                                       examples/iowa/entities.py:33-73
//   agents with a second kind, role hand-over                           examples/tag/agents.py:40-106
//
// Tiling: a *group* of LPE lanes (4..32, a power of two) owns one world instance whose dynamic
// planes are staged in shared memory for the whole launch (all `num_turns` turns), so the grid is
// read and written once per launch while every turn streams one observation row per agent straight
// into its replay-ring slot with 128-bit stores.  Groups never talk to each other; the only
// intra-warp cooperation is the per-warp spawn-candidate queue of the Philox sweep.  All synchronisation is
// __syncwarp(): every phase that lets ONE lane mutate the instance (a move, a beam, a hand-over) is fenced from the
// phases in which all lanes read it.
//
// Worlds with several stored planes (cleanup, iowa) gather their windows by row segments with byte-permute channel
// lookups and ordered stores of plane counts (`scatter_cols_multi`); their sweep classifies cells four at a time.
// A split-phase launch that only moves one agent takes `act_only_kernel` (one thread per instance, planes in place).
// The standalone observation (`sb_observe`) has three forms: bulk-async stores of whole tiles of small rows
// (`observe_bulk_kernel`), a software-pipelined row-wise kernel for 7..16-wide windows, and `observe_kernel`.
//
// Large maps with many agents (16..64 agents, planes of more than 4 KB) take `step_kernel_block` further down: one
// BLOCK per instance, the plane staged in shared memory, up to sixteen mutually independent agents observed and moved
// at once (dependency levels keep the list order of environment.py:92).  The older form of that idea -- planes used
// in place in global memory (STAGED = false, one warp per instance), four mutually distant agents at once by 8 lanes
// each -- remains for shapes the block kernel does not cover.
#include <cuda_runtime.h>

#include "sb_core.cuh"
#include "sb_internal.h"

namespace sb {

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

__host__ __device__ inline int pad64mod128(int b) {
  b = round_up(b, 16);
  while ((b & 127) != 64) b += 16;
  return b;
}

// Shared-memory layout of a block (see StepParams).  The staging row of an instance holds ONE
// observation (+4 elements of alignment slack) and is all-zero between observations; it is bf16: entries
// are 0 / 1 (one stored plane) or the number of stored planes showing that channel in that cell (several planes
// are summed: small integers, exact in bf16), and the float is the bf16 shifted left by 16.
struct SmemLayout {
  int q_off, planes_off, pos_off, stage_off, hits_off, plane_env, stage_env, stage_row, total;
  int agent_words, act_turns;
};

// Turns whose actions are drawn (or loaded from the tape) at once, by different lanes of the group: with fewer agents
// than lanes, lane l prepares the action of (turn t + l / A, agent l % A).  One agent on four lanes: one Philox call
// per lane every four turns instead of one call by lane 0 (three lanes idle) every turn.
__host__ __device__ inline int act_turns_ahead(int lpe, int A) { return A < lpe ? lpe / A : 1; }

// planes above this size stay in global memory (a 128x128 plane is 16 KB: staging it would leave room for
// ~12 instances, i.e. 12 warps, per SM)
__host__ inline bool stage_planes(const SbDev& d, int tune_staged) {
  if (tune_staged) return tune_staged > 0;  // SbStepArgs.tune_staged
  return d.env_stride <= 4096;
}

__host__ inline SmemLayout smem_layout(const SbDev& d, int threads, int lpe, bool np1, bool staged, bool sparse) {
  SmemLayout L;
  const int epb = threads / lpe;
  L.q_off = round_up(2 * d.SS, 16) + 4 * SB_KTAB_WORDS;  // window LUT, per-kind tables, then the queues
  L.planes_off = L.q_off + round_up((threads / 32) * (SB_QCAP + 1) * 4, 16);
  L.plane_env = staged ? pad64mod128(d.env_stride) : 0;
  L.pos_off = L.planes_off + epb * L.plane_env;
  // per instance: packed yx [A], actions [act_turns x A], facing [A], reward of the turn [A]
  L.act_turns = act_turns_ahead(lpe, d.A);
  L.agent_words = 3 * d.A + L.act_turns * d.A;
  L.stage_off = L.pos_off + round_up(epb * 4 * L.agent_words, 16);
  L.stage_row = pad64mod128((d.obs_size + 4) * 2);  // bf16 on every fast path (entries are small integer counts)
  // warp-per-instance worlds with several agents observe up to SB_PAR independent agents at once
  L.stage_env = (np1 && lpe == 32 && d.A >= SB_PAR) ? SB_PAR * L.stage_row : L.stage_row;
  L.hits_off = L.stage_off + epb * L.stage_env;  // SB_RNG_SPARSE: accepted cells of the turn, u16[KMAX]
  L.total = L.hits_off + (sparse ? epb * SB_SPARSE_KMAX * 2 : 0);  // only the sparse sampler keeps a hit list
  return L;
}

#ifndef SB_ASYNC_STAGE
#define SB_ASYNC_STAGE 1
#endif
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* gmem) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem));
}

__device__ __forceinline__ void st_f4(float* p, float a, float b, float c, float e) {
#ifdef SB_EXP_NOSTORE  // timing experiment only (wrong results): the row is produced but never leaves the SM
  if (a == 123.0f)
#endif
  // streaming 128-bit store: observation rows are written once and not re-read by this kernel
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(e)
               : "memory");
}

// Row-wise form of the window scatter for NARROW groups (LANES <= 8) on a staged single plane.  A lane takes whole
// window ROWS (i = lane, lane + LANES, ...) instead of single cells: the bytes of a row are fetched with two aligned
// 32-bit shared loads per 5 cells and funnel-shifted into place, out-of-bounds cells are turned into "kind 7" -- whose
// slot of channel8_oob holds the fill kind's channel -- by a byte mask, and four kinds are looked up by ONE byte
// permute.  Per cell that leaves: extract the code, test it, one multiply-add for the address, one 16-bit store
// (the cell-wise loop spends ~35 instructions per cell on index arithmetic, bounds tests and a byte load).
// Requires K <= 7 (d.rowwise_ok).  `genv` must be shared memory: loads may touch up to 12 bytes before / after the
// instance image (masked out), which always lie inside the block's allocation (queue before, agent area after).
// PARTIAL = false: S is a multiple of 5 (the common 5x5 window): every group is whole.  PARTIAL = true: the last
// group of a row holds S % 5 cells; cells past the row end get code 255.  Two instantiations rather than one loop with
// a count: the extra mask arithmetic in the whole-group path cost 3-9 % of the step (profiles/README.md).
template <int LANES, bool PARTIAL>
__device__ __forceinline__ void scatter_rows_onehot_impl(const SbDev& d, const uint8_t* genv, uint16_t* srow16, int y,
                                                         int x, int lane) {
  const int x0 = x - d.v;                                 // map column of window column 0 (may be negative)
  const int jlo = max(0, -x0), jhi = min(d.S, d.W - x0);  // window columns [jlo, jhi) lie inside the map
  const uint32_t tlo = (uint32_t)d.channel8_oob, thi = (uint32_t)(d.channel8_oob >> 32);
  for (int i = lane; i < d.S; i += LANES) {
    const int yy = y - d.v + i;
    const bool row_in = (unsigned)yy < (unsigned)d.H;
    const int off = yy * d.W + x0;  // plane offset of window cell (i, 0); out-of-bounds bytes are masked below
    uint16_t* const srow = srow16 + i * d.S;
    for (int j0 = 0; j0 < (PARTIAL ? d.S : d.S5); j0 += 5) {
      // byte mask of the in-bounds cells among j0 .. j0+3, and whether cell j0+4 is in bounds
      const int lc = min(max(jlo - j0, 0), 4), hc = min(max(jhi - j0, 0), 4);
      uint32_t m4 = __funnelshift_lc(~0u, 0u, 8 * hc) & ~__funnelshift_lc(~0u, 0u, 8 * lc);
      bool in5 = jlo <= j0 + 4 && j0 + 4 < jhi;
      uint32_t w0 = 0u, w1 = 0u;
      if (row_in) {
        const int al = (off + j0) & ~3;
        w0 = *reinterpret_cast<const uint32_t*>(genv + al);
        w1 = *reinterpret_cast<const uint32_t*>(genv + al + 4);
      } else {
        m4 = 0u;
        in5 = false;
      }
      const uint32_t sh = ((uint32_t)(off + j0) & 3u) * 8u;
      uint32_t k4 = __funnelshift_r(w0, w1, sh);                   // kinds of cells j0 .. j0+3
      const uint32_t k5 = in5 ? ((w1 >> sh) & 7u) : 7u;             // kind of cell j0+4
      k4 = (k4 & m4) | (0x07070707u & ~m4);                        // out of bounds -> slot 7 (the fill channel)
      const uint32_t t = k4 & 0x07070707u;
      const uint32_t sel = __byte_perm(t | (t >> 4), 0u, 0x4420);  // nibbles k0, k1, k2, k3
      uint32_t c4 = __byte_perm(tlo, thi, sel);                    // four channel codes
      uint32_t c5 = __byte_perm(tlo, thi, k5) & 0xFFu;
      if (PARTIAL) {
        const int cnt = d.S - j0;
        c4 |= ~__funnelshift_lc(~0u, 0u, 8 * min(cnt, 4));
        if (cnt < 5) c5 = 255u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t code = (c4 >> (8 * u)) & 0xFFu;
        if (code != 255u) srow[code * d.SS + j0 + u] = 0x3F80;  // bf16(1.0)
      }
      if (c5 != 255u) srow[c5 * d.SS + j0 + 4] = 0x3F80;
    }
  }
}

// SEGMENTED form of the same gather: a lane takes a SEGMENT of a window row (d.rw_len cells; d.rw_nseg segments per row,
// from the cost model in make_dev: trips x five-cell steps), so that a window whose rows outnumber the lanes by one --
// 9x9 windows on the 8-lane agent slots of step_kernel_block: rows 0..7, then row 8 by one lane -- is not walked
// twice at full cost (segments of 5 and 4 cells: 18 units = 3 trips of ONE step each instead of 2 trips of two).
#ifndef SB_ROWSEG
#define SB_ROWSEG 1
#endif
template <int LANES>
__device__ __forceinline__ void scatter_rows_onehot_seg(const SbDev& d, const uint8_t* genv, uint16_t* srow16, int y,
                                                        int x, int lane) {
  constexpr int LG = LANES == 32 ? 5 : LANES == 16 ? 4 : LANES == 8 ? 3 : LANES == 4 ? 2 : 1;
  const int S = d.S, g = d.rw_len[LG], units = S * d.rw_nseg[LG];
  const int x0 = x - d.v;
  const int jlo = max(0, -x0), jhi = min(S, d.W - x0);
  const uint32_t tlo = (uint32_t)d.channel8_oob, thi = (uint32_t)(d.channel8_oob >> 32);
  for (int un = lane; un < units; un += LANES) {
    const int sg = (int)(((uint32_t)un * d.inv_S) >> 16), i = un - sg * S;  // segment-major: (segment, window row)
    const int yy = y - d.v + i;
    const bool row_in = (unsigned)yy < (unsigned)d.H;
    const int off = yy * d.W + x0;
    uint16_t* const srow = srow16 + i * S;
    const int j_end = min(S, (sg + 1) * g);
    for (int j0 = sg * g; j0 < j_end; j0 += 5) {
      const int lc = min(max(jlo - j0, 0), 4), hc = min(max(jhi - j0, 0), 4);
      uint32_t m4 = __funnelshift_lc(~0u, 0u, 8 * hc) & ~__funnelshift_lc(~0u, 0u, 8 * lc);
      bool in5 = jlo <= j0 + 4 && j0 + 4 < jhi;
      uint32_t w0 = 0u, w1 = 0u;
      if (row_in) {
        const int al = (off + j0) & ~3;
        w0 = *reinterpret_cast<const uint32_t*>(genv + al);
        w1 = *reinterpret_cast<const uint32_t*>(genv + al + 4);
      } else {
        m4 = 0u;
        in5 = false;
      }
      const uint32_t sh = ((uint32_t)(off + j0) & 3u) * 8u;
      uint32_t k4 = __funnelshift_r(w0, w1, sh);
      const uint32_t k5 = in5 ? ((w1 >> sh) & 7u) : 7u;
      k4 = (k4 & m4) | (0x07070707u & ~m4);
      const uint32_t t = k4 & 0x07070707u;
      const uint32_t sel = __byte_perm(t | (t >> 4), 0u, 0x4420);
      uint32_t c4 = __byte_perm(tlo, thi, sel);
      uint32_t c5 = __byte_perm(tlo, thi, k5) & 0xFFu;
      const int cnt = j_end - j0;  // cells of the segment from j0 on
      c4 |= ~__funnelshift_lc(~0u, 0u, 8 * min(cnt, 4));
      if (cnt < 5) c5 = 255u;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t code = (c4 >> (8 * e)) & 0xFFu;
        if (code != 255u) srow[code * d.SS + j0 + e] = 0x3F80;
      }
      if (c5 != 255u) srow[c5 * d.SS + j0 + 4] = 0x3F80;
    }
  }
}

template <int LANES>
__device__ __forceinline__ void scatter_rows_onehot(const SbDev& d, const uint8_t* genv, uint16_t* srow16, int y, int x,
                                                    int lane) {
  if (d.S5 == d.S) scatter_rows_onehot_impl<LANES, false>(d, genv, srow16, y, x, lane);
  else scatter_rows_onehot_impl<LANES, true>(d, genv, srow16, y, x, lane);
}

// 0x80 in every byte of x that is zero (exact: no carries across bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
  const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | x | 0x7F7F7F7Fu);
}

// Window gather for SEVERAL stored planes (2 or 3) staged in shared memory, any group width: the multi-plane
// counterpart of scatter_rows_onehot.  A lane takes a SEGMENT of a window row (d.cg_len cells; d.cg_nseg segments per
// row, chosen per group width so that S x nseg units fill the lanes: 11x11 windows on 32 lanes = 22 units of 6 / 5
// cells) and walks it four cells at a time: per plane two aligned 32-bit shared loads and a funnel shift fetch the four
// kinds, two byte permutes (kinds 0..7 / 8..15) turn them into channel codes, out-of-bounds cells become the fill
// channel on the first plane and "nothing" on the others by a byte mask.  The codes of a cell's column are then in
// registers together, so each (channel, cell) entry is STORED with its final count -- 1, or 2 / 3 where several
// planes show the same channel (cleanup: walls on every layer) -- instead of a shared-memory read-modify-write per
// plane.  Per cell-plane that is ~8 instructions against ~35 for the cell-wise loop (a byte load, a dependent table
// load, bounds tests and index arithmetic per plane).  Requires K <= 16, not toroidal (d.colgather_ok); `genv` must
// be shared memory (loads may touch a few bytes before / after the instance image, masked out, inside the block's
// allocation).
template <int LANES>
__device__ __forceinline__ void scatter_cols_multi(const SbDev& d, const uint8_t* genv, uint16_t* srow16, int y, int x,
                                                   int lane) {
  constexpr int LG = LANES == 32 ? 5 : LANES == 16 ? 4 : LANES == 8 ? 3 : LANES == 4 ? 2 : 1;
  const int S = d.S, g = d.cg_len[LG], units = S * d.cg_nseg[LG];
  const int x0 = x - d.v;                               // map column of window column 0 (may be negative)
  const int jlo = max(0, -x0), jhi = min(S, d.W - x0);  // window columns [jlo, jhi) lie inside the map
  const uint32_t clo0 = (uint32_t)d.channel8, clo1 = (uint32_t)(d.channel8 >> 32);
  const uint32_t chi0 = (uint32_t)d.channel8_hi, chi1 = (uint32_t)(d.channel8_hi >> 32);
  const uint32_t fill4 = (uint32_t)(d.fill_channel & 0xFF) * 0x01010101u;
  for (int u = lane; u < units; u += LANES) {
    const int sg = (int)(((uint32_t)u * d.inv_S) >> 16), i = u - sg * S;  // segment-major: (segment, window row)
    const int yy = y - d.v + i;
    const bool row_in = (unsigned)yy < (unsigned)d.H;
    const int off = yy * d.W + x0;  // plane offset of window cell (i, 0); out-of-bounds bytes are masked below
    uint16_t* const srow = srow16 + i * S;
    const int j_end = min(S, (sg + 1) * g);
    for (int j0 = sg * g; j0 < j_end; j0 += 4) {
      const int cnt = min(4, j_end - j0);
      const int lc = min(max(jlo - j0, 0), 4), hc = min(max(jhi - j0, 0), 4);
      const uint32_t m4 = row_in ? (__funnelshift_lc(~0u, 0u, 8 * hc) & ~__funnelshift_lc(~0u, 0u, 8 * lc)) : 0u;
      const uint32_t past = ~__funnelshift_lc(~0u, 0u, 8 * cnt);  // bytes of cells beyond the segment: show nothing
      uint32_t c[3];
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        c[pl] = 0xFFFFFFFFu;
        if (pl < d.n_planes) {
          uint32_t k4 = 0u;
          if (row_in) {
            const int o = pl * d.plane_stride + off + j0;
            const int al = o & ~3;
            const uint32_t w0 = *reinterpret_cast<const uint32_t*>(genv + al);
            const uint32_t w1 = *reinterpret_cast<const uint32_t*>(genv + al + 4);
            k4 = __funnelshift_r(w0, w1, ((uint32_t)o & 3u) * 8u);  // kinds of cells j0 .. j0+3
          }
          const uint32_t t = k4 & 0x07070707u;
          const uint32_t sel = __byte_perm(t | (t >> 4), 0u, 0x4420);  // nibbles k0, k1, k2, k3 (low three bits)
          uint32_t c4 = __byte_perm(clo0, clo1, sel);
          if (d.K > 8) {  // kinds 8..15: the second table, chosen per byte by bit 3 of the kind
            const uint32_t h4 = __byte_perm(chi0, chi1, sel);
            const uint32_t m8 = ((k4 >> 3) & 0x01010101u) * 0xFFu;
            c4 = (c4 & ~m8) | (h4 & m8);
          }
          c[pl] = (c4 & m4) | ((pl == 0 ? fill4 : 0xFFFFFFFFu) & ~m4) | past;  // out of bounds: the fill kind, once
        }
      }
      // Entry (channel, cell) must end up holding the NUMBER of planes that show that channel there.  The planes are
      // stored in order by the same lane, so plane q writes 1 + (number of earlier planes with the same code) and the
      // last store to an address wins: no read-modify-write, no branch on "shown twice".  The equalities are found
      // four cells at a time (zero bytes of the XOR of two code words); bf16(2) = bf16(1) + 0x80, bf16(3) = + 0xC0.
      uint16_t* const cell0 = srow + j0;
      if (c[0] != 0xFFFFFFFFu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t a = (c[0] >> (8 * e)) & 0xFFu;
          if (a != 255u) cell0[a * d.SS + e] = 0x3F80;
        }
      }
      if (c[1] != 0xFFFFFFFFu) {
        const uint32_t add1 = zero_bytes(c[0] ^ c[1]);  // 0x80 where plane 1 repeats the channel of plane 0
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t b = (c[1] >> (8 * e)) & 0xFFu;
          if (b != 255u) cell0[b * d.SS + e] = (uint16_t)(0x3F80u + ((add1 >> (8 * e)) & 0x80u));
        }
      }
      if (c[2] != 0xFFFFFFFFu) {
        const uint32_t z02 = zero_bytes(c[0] ^ c[2]), z12 = zero_bytes(c[1] ^ c[2]);
        const uint32_t add2 = (z02 | z12) | ((z02 & z12) >> 1);  // 0x80: one earlier plane repeats it, 0xC0: both
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t cc = (c[2] >> (8 * e)) & 0xFFu;
          if (cc != 255u) cell0[cc * d.SS + e] = (uint16_t)(0x3F80u + ((add2 >> (8 * e)) & 0xC0u));
        }
      }
    }
  }
}

// One observation row through the one-hot fast path, by the LANES lanes that own the instance (all
// of them must call; `dst == nullptr` only takes part in the synchronisation; the caller issues the
// trailing __syncwarp, which lets it overlap the agent's move with the stream-out).  The row is all zeros
// except one entry per (window cell, stored plane): the lanes scatter those into the zeroed staging
// row, then stream the row out with wide shared loads / 128-bit global stores, zeroing it again on the
// way.  `off` shifts the row inside the staging buffer so that the first 16-byte aligned global
// address lines up with an aligned shared address (ring rows are only 8-byte aligned: 600 B, 1944 B).
// `genv` may point to shared or global memory.
#ifndef SB_STREAM_UNROLL
#define SB_STREAM_UNROLL 4  // vectors per trip of the stream-out loop of long rows
#endif
#ifndef SB_STREAM_LONG
#define SB_STREAM_LONG 64  // rows of at least this many 16-byte vectors (9x9 windows: 121; cleanup: 275) take it
#endif
template <int LANES, bool NP1, int NB = 1, int ROWWISE = 0>
__device__ __forceinline__ void emit_onehot_row(const SbDev& d, const uint8_t* genv, const uint32_t* ktab,
                                                uint8_t* stage_raw, int y, int x, float* dst, int lane) {
  uint16_t* stage16 = reinterpret_cast<uint16_t*>(stage_raw);  // bf16: 0 / 1.0 (NP1) or a small plane count
  const int head = dst ? (4 - (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3)) & 3 : 0;
  const int off = (4 - head) & 3;
  if (dst) {
    // window cell ij -> (i, j) = (ij / S, ij % S) by multiply-shift (exact for ij < 512, S <= 21); the
    // loop is unrolled so the dependent shared-memory reads of several cells are in flight together
    const uint32_t inv_s = d.inv_S;
    uint16_t* const srow16 = stage16 + off;  // hoisted: the compiler recomputed `off` under the store's predicate
#ifndef SB_SCATTER_UNROLL
#define SB_SCATTER_UNROLL 2  // measured: 7.60 -> 7.55 ms at the default shape
#endif
    constexpr int kScatterUnroll = (NB > 1) ? 1 : SB_SCATTER_UNROLL;
    // ROWWISE (planes staged in shared memory): one stored plane -> row-wise gather of narrow groups; several
    // stored planes -> column gather (scatter_cols_multi), any group width
    const bool rowwise = ROWWISE != 0 && (NP1 ? d.rowwise_ok : d.colgather_ok);
    if (rowwise) {
      if constexpr (NP1 && ROWWISE == 2) scatter_rows_onehot_seg<LANES>(d, genv, srow16, y, x, lane);
      else if constexpr (NP1) scatter_rows_onehot<LANES>(d, genv, srow16, y, x, lane);
      else scatter_cols_multi<LANES>(d, genv, srow16, y, x, lane);
    }
#pragma unroll kScatterUnroll
    for (int ij = rowwise ? d.SS : lane; ij < d.SS; ij += LANES) {
      const int i = (int)(((uint32_t)ij * inv_s) >> 16), j = ij - i * d.S;
      if (NP1 && NB > 1) {
        // planes in GLOBAL memory (standalone observation kernel, unstaged large maps): issue the NB window
        // loads of this lane back to back, then the dependent scatters (the compiler cannot hoist a generic
        // load over a shared store by itself)
        uint32_t kinds[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int q = ij + u * LANES;
          const int qi = (int)(((uint32_t)q * inv_s) >> 16), qj = q - qi * d.S;
          int yy = y - d.v + qi, xx = x - d.v + qj;
          if (d.toroidal) {  // fold once: |offset| <= v <= (H-1)/2
            yy = yy < 0 ? yy + d.H : (yy >= d.H ? yy - d.H : yy);
            xx = xx < 0 ? xx + d.W : (xx >= d.W ? xx - d.W : xx);
          }
          kinds[u] = 0x100u;  // beyond the window
          if (q < d.SS)
            kinds[u] = ((unsigned)yy >= (unsigned)d.H || (unsigned)xx >= (unsigned)d.W) ? 0x200u  // out of bounds
                                                                                      : genv[yy * d.W + xx];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const uint32_t code = kinds[u] < 0x100u ? sb_kind_channel(d, kinds[u])
                                                  : (kinds[u] == 0x200u ? (uint32_t)d.fill_channel : 255u);
          if (code != 255u) stage16[off + code * d.SS + ij + u * LANES] = 0x3F80;
        }
        ij += (NB - 1) * LANES;
      } else if (NP1) {  // one stored plane: every window cell contributes at most one 1.0
        const uint32_t code = sb_window_code(d, genv, y, x, i, j, 0);
        if (code != 255u) srow16[code * d.SS + ij] = 0x3F80;  // bf16(1.0)
      } else {
        // several stored planes are SUMMED (visual_field.py:51).  Slot (channel, ij) of the zeroed row is only
        // ever touched by this lane in this loop iteration, so the first plane stores and only a later plane
        // that lands on the same channel (stacked walls) needs a read-modify-write.
        int yy = y - d.v + i, xx = x - d.v + j;
        if (d.toroidal) {
          yy = yy < 0 ? yy + d.H : (yy >= d.H ? yy - d.H : yy);
          xx = xx < 0 ? xx + d.W : (xx >= d.W ? xx - d.W : xx);
        }
        uint16_t* srow = srow16 + ij;
        if ((unsigned)yy >= (unsigned)d.H || (unsigned)xx >= (unsigned)d.W) {
          if (d.fill_channel != 255) srow[d.fill_channel * d.SS] = 0x3F80;  // out of bounds: the fill kind, once
        } else {
          const uint8_t* pc = genv + yy * d.W + xx;
          uint32_t first = 255u;
          for (int pl = 0; pl < d.n_planes; ++pl, pc += d.plane_stride) {
            // several planes mean more than 8 kinds in practice: channel from the shared-memory kind table
            const uint32_t code = ktab ? (ktab[*pc & 31u] & 0xFFu) : sb_kind_channel(d, *pc);
            if (code == 255u) continue;
            if (first == 255u) {
              srow[code * d.SS] = 0x3F80;  // bf16(1.0)
              first = code;
            } else {  // bf16 count + 1 (exact: counts are at most the number of planes)
              const float f = __uint_as_float((uint32_t)srow[code * d.SS] << 16) + 1.0f;
              srow[code * d.SS] = (uint16_t)(__float_as_uint(f) >> 16);
            }
          }
        }
      }
    }
  }
  __syncwarp();
  if (dst) {
    const int nvec = d.oh_size > head ? (d.oh_size - head) >> 2 : 0;
    const int etail = head + 4 * nvec + lane;
    {
      if (lane < head && lane < d.oh_size) {
        dst[lane] = __uint_as_float((uint32_t)stage16[off + lane] << 16);
        stage16[off + lane] = 0;
      }
      // 4 bf16 per 8-byte shared load -> one 128-bit streaming store
      if constexpr (LANES > 4) {
      uint2* s2 = reinterpret_cast<uint2*>(stage16 + off + head);
      int qv = lane;
#if SB_STREAM_UNROLL > 1
      // long rows (cleanup: 275 vectors): several vectors per trip, loads first (measured: cleanup 9.64 -> 9.55 ms;
      // the short rows of one-plane worlds gain nothing or lose, so they keep the simple loop)
      // (compile-time choice: the extra loop in the one-plane step_kernel variants cost the sparse sampler 8 % at
      // config 3 -- registers / instruction cache -- even when it was skipped at run time)
      constexpr bool kLongRows = !NP1 || ROWWISE == 2;
      for (; kLongRows && nvec >= SB_STREAM_LONG && qv + (SB_STREAM_UNROLL - 1) * LANES < nvec;
           qv += SB_STREAM_UNROLL * LANES) {
        uint2 w[SB_STREAM_UNROLL];
#pragma unroll
        for (int u = 0; u < SB_STREAM_UNROLL; ++u) w[u] = s2[qv + u * LANES];
#pragma unroll
        for (int u = 0; u < SB_STREAM_UNROLL; ++u) {
          s2[qv + u * LANES] = make_uint2(0u, 0u);
          st_f4(dst + head + 4 * (qv + u * LANES), __uint_as_float(w[u].x << 16),
                __uint_as_float(w[u].x & 0xFFFF0000u), __uint_as_float(w[u].y << 16),
                __uint_as_float(w[u].y & 0xFFFF0000u));
        }
      }
#endif
      for (; qv < nvec; qv += LANES) {
        const uint2 w = s2[qv];
        s2[qv] = make_uint2(0u, 0u);
        st_f4(dst + head + 4 * qv, __uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u),
              __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xFFFF0000u));
      }
      } else {
      // narrow groups own ~10 vectors per lane: running pointers and two vectors per trip (the address
      // arithmetic and the branch were half of this loop's instructions; measured 7.85 -> 7.71 ms at LANES = 4,
      // but slower for 8- and 32-lane groups, which keep the simple loop)
      uint2* sp = reinterpret_cast<uint2*>(stage16 + off + head) + lane;
      float* gp = dst + head + 4 * lane;
      int left = nvec - lane;  // vectors this lane still owns: left, left - LANES, ...
      for (; left > LANES; left -= 2 * LANES, sp += 2 * LANES, gp += 8 * LANES) {
        const uint2 w0 = sp[0], w1 = sp[LANES];
        sp[0] = make_uint2(0u, 0u);
        sp[LANES] = make_uint2(0u, 0u);
        st_f4(gp, __uint_as_float(w0.x << 16), __uint_as_float(w0.x & 0xFFFF0000u), __uint_as_float(w0.y << 16),
              __uint_as_float(w0.y & 0xFFFF0000u));
        st_f4(gp + 4 * LANES, __uint_as_float(w1.x << 16), __uint_as_float(w1.x & 0xFFFF0000u),
              __uint_as_float(w1.y << 16), __uint_as_float(w1.y & 0xFFFF0000u));
      }
      if (left > 0) {
        const uint2 w = sp[0];
        sp[0] = make_uint2(0u, 0u);
        st_f4(gp, __uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
              __uint_as_float(w.y & 0xFFFF0000u));
      }
      }
      if (etail < d.oh_size) {
        dst[etail] = __uint_as_float((uint32_t)stage16[off + etail] << 16);
        stage16[off + etail] = 0;
      }
    }
  }
  // NOTE: the caller must __syncwarp() before the staging row is scattered into again.
}

#ifndef SB_ONLY_LPE  // the step-kernel instantiations of one LPE are compiled in their own translation unit
// ----------------------------------------------------------------------------------------------
// Standalone observation kernel (sb_observe): a pure function of grid + positions, so every
// (instance, agent) pair is independent: one warp per pair, window cells read straight from global
// memory (the instance's planes are shared by its agents through L1/L2), rows streamed out with
// 128-bit stores.  Config 4 (128x128, 64 agents, 9x9 window): 1944 B written per pair.
// ----------------------------------------------------------------------------------------------
constexpr int OBS_THREADS = 256;
#ifndef SB_OBS_ROWWISE
#define SB_OBS_ROWWISE 1
#endif
#ifndef SB_OBS_PIPELINED
#define SB_OBS_PIPELINED 1
#endif
#ifndef SB_OBS_SEG
#define SB_OBS_SEG 1
#endif
#ifndef SB_OBS_MINB
#define SB_OBS_MINB 6  // resident blocks per SM the register allocation is sized for (40 registers)
#endif

// LANES lanes (16 or 32) serve one (instance, agent) pair; every group walks a CONTIGUOUS range of
// pairs so that (env, agent) advance incrementally (no divisions in the loop), the agents of one
// instance are served back to back (its planes stay in L1) and consecutive rows land in the same DRAM
// pages.
// ROWS: gather the window row by row (scatter_rows_onehot) from the planes in global memory.  The row gather may touch
// a few bytes before / after an instance's planes, so the first and the last instance of the batch (whose neighbours
// are outside the allocation) take the cell-wise path; the choice is made per warp (both halves agree).
template <bool NP1, int LANES, int ROWS = 0>  // ROWS: 0 cell-wise, 1 one lane per window row, 2 row segments
__global__ void __launch_bounds__(OBS_THREADS, SB_OBS_MINB)
    observe_kernel(const __grid_constant__ SbDev d, const uint8_t* __restrict__ grid,
                   const int32_t* __restrict__ agent_pos, float* __restrict__ out, int stage_bytes, int fast,
                   unsigned pairs_per_group) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int GROUPS = OBS_THREADS / LANES;
  uint16_t* lut_ij = reinterpret_cast<uint16_t*>(smem_raw);
  const int lut_bytes = round_up(2 * d.SS, 16);
  uint32_t* ktab = reinterpret_cast<uint32_t*>(smem_raw + lut_bytes);
  if (threadIdx.x < 32) sb_fill_kind_tables(d, ktab, threadIdx.x);
  const int grp = threadIdx.x / LANES, lane = threadIdx.x & (LANES - 1);
  uint8_t* stage_raw = smem_raw + lut_bytes + 4 * SB_KTAB_WORDS + grp * stage_bytes;
  for (int ij = threadIdx.x; ij < d.SS; ij += OBS_THREADS) {
    int i = ij / d.S;
    lut_ij[ij] = (uint16_t)((i << 8) | (ij - i * d.S));
  }
  for (int e = lane; e < stage_bytes / 4; e += LANES) reinterpret_cast<uint32_t*>(stage_raw)[e] = 0u;
  __syncthreads();
  const unsigned long long pairs = (unsigned long long)d.N * d.A;
  const unsigned long long first = ((unsigned long long)blockIdx.x * GROUPS + grp) * pairs_per_group;
  unsigned long long env = first / (unsigned)d.A;
  int a = (int)(first - env * (unsigned)d.A);
  for (unsigned k = 0; k < pairs_per_group; ++k) {  // uniform trip count: both halves of a warp stay in step
    const unsigned long long pair = first + k;
    const bool valid = pair < pairs;
    int y = 0, x = 0;
    if (valid) {
      const int pos = agent_pos[pair];
      y = d.inv_W ? (int)__umulhi((uint32_t)pos, d.inv_W) : pos / d.W;  // exact multiply-shift division (make_dev)
      x = pos - y * d.W;
    }
    const uint8_t* genv = grid + (size_t)(valid ? env : 0) * d.env_stride;
    float* dst = valid ? out + (size_t)pair * d.obs_size : nullptr;
    if (fast) {
      bool rows = false;
      if constexpr (ROWS != 0)
        rows = __all_sync(0xFFFFFFFFu, !valid || (env > 0 && env + 1 < (unsigned long long)d.N));
      if (ROWS != 0 && rows) {
        emit_onehot_row<LANES, NP1, 1, ROWS>(d, genv, ktab, stage_raw, y, x, dst, lane);
      } else {
        // window loads batched: 2 per lane cover a 5x5 window with 16 lanes, 4 per lane a 9x9 / 11x11 one with 32
        emit_onehot_row<LANES, NP1, (LANES == 16 ? 2 : 4)>(d, genv, ktab, stage_raw, y, x, dst, lane);
      }
      __syncwarp();
    } else if (valid) {
      for (int e = lane; e < d.oh_size; e += LANES) {
        const int c = e / d.SS;
        const uint32_t li = lut_ij[e - c * d.SS];
        dst[e] = (float)sb_obs_general(d, genv, d.tmpl, d.app, y, x, c, (int)(li >> 8), (int)(li & 255u));
      }
    }
    if (d.obs_extra && valid) {  // positional code of the agent's cell (CleanupObservation.observe)
      const float* src = d.extra_tbl + (size_t)(y * d.W + x) * d.obs_extra;
      for (int e = lane; e < d.obs_extra; e += LANES) dst[d.oh_size + e] = __ldg(src + e);
    }
    if (d.obs_flag && valid && lane == 0)  // TagAgent.pov: [self.it]
      dst[d.oh_size + d.obs_extra] = genv[d.agent_plane * d.plane_stride + y * d.W + x] == (uint8_t)d.agent_kind_alt
                                         ? 1.0f : 0.0f;
    if (++a == d.A) {
      a = 0;
      ++env;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Standalone observation, row-wise gather, SOFTWARE-PIPELINED over the pairs a group walks (one-hot worlds with one
// stored plane, 7..16-wide windows, no appended floats).  The plain kernel above was latency-bound: 47 % of its stall
// samples sat on the two dependent global loads of a pair (its position, then the window words that position
// addresses).  Here a lane (= one window row) issues the window loads of pair k+1 and the position load of pair k+2
// BEFORE it scatters and streams out pair k, so both round trips are hidden behind a whole pair's work.
// Loads are guarded against the ends of the grid allocation (a row gather may touch up to 12 bytes outside an
// instance's planes), so no instance needs a special path.
// ----------------------------------------------------------------------------------------------
constexpr int OBS_ROW_LANES = 16, OBS_ROW_MAXG = 4;  // 16 lanes per pair; up to 4 groups of 5 cells per window row

#ifndef SB_OBSROW_MINB
#define SB_OBSROW_MINB 4  // the pipeline holds two pairs' window words in registers; latency is hidden by it, not by warps
#endif
template <bool PARTIAL>
__global__ void __launch_bounds__(OBS_THREADS, SB_OBSROW_MINB)
    observe_rows_kernel(const __grid_constant__ SbDev d, const uint8_t* __restrict__ grid,
                        const int32_t* __restrict__ agent_pos, float* __restrict__ out, int stage_bytes,
                        unsigned pairs_per_group, long long grid_bytes) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int LANES = OBS_ROW_LANES, GROUPS = OBS_THREADS / LANES;
  const int grp = threadIdx.x / LANES, lane = threadIdx.x & (LANES - 1);
  uint16_t* stage16 = reinterpret_cast<uint16_t*>(smem_raw + grp * stage_bytes);
  for (int e = lane; e < stage_bytes / 4; e += LANES) reinterpret_cast<uint32_t*>(stage16)[e] = 0u;
  __syncwarp();
  const unsigned long long pairs = (unsigned long long)d.N * d.A;
  const unsigned long long first = ((unsigned long long)blockIdx.x * GROUPS + grp) * pairs_per_group;
  const int ng = (d.S + 4) / 5;
  const uint32_t tlo = (uint32_t)d.channel8_oob, thi = (uint32_t)(d.channel8_oob >> 32);

  auto load_pos = [&](unsigned long long pair) -> int { return pair < pairs ? __ldg(agent_pos + pair) : -1; };
  // window words of row `lane` of the pair at `pair` (instance env) whose agent stands on cell `pos`
  auto issue = [&](unsigned long long env, int pos, uint32_t (&w0)[OBS_ROW_MAXG], uint32_t (&w1)[OBS_ROW_MAXG]) {
    const int y = pos < 0 ? 0 : (d.inv_W ? (int)__umulhi((uint32_t)pos, d.inv_W) : pos / d.W);
    const int x = pos < 0 ? 0 : pos - y * d.W;
    const int yy = y - d.v + lane;
    const bool row_in = pos >= 0 && lane < d.S && (unsigned)yy < (unsigned)d.H;
    const long long base = (long long)env * d.env_stride + (long long)yy * d.W + (x - d.v);
#pragma unroll
    for (int g = 0; g < OBS_ROW_MAXG; ++g) {
      w0[g] = 0u, w1[g] = 0u;
      if (g < ng && row_in) {
        const long long al = (base + 5 * g) & ~3ll;
        if (al >= 0) w0[g] = __ldg(reinterpret_cast<const uint32_t*>(grid + al));
        if (al + 8 <= grid_bytes) w1[g] = __ldg(reinterpret_cast<const uint32_t*>(grid + al + 4));
      }
    }
  };

  unsigned long long env_c = first / (unsigned)d.A;
  int a_c = (int)(first - env_c * (unsigned)d.A);
  unsigned long long env_n = env_c;
  int a_n = a_c;
  auto advance = [&](unsigned long long& env, int& a) {
    if (++a == d.A) {
      a = 0;
      ++env;
    }
  };
  advance(env_n, a_n);
  int pos_c = load_pos(first), pos_n = load_pos(first + 1);
  uint32_t cw0[OBS_ROW_MAXG], cw1[OBS_ROW_MAXG], nw0[OBS_ROW_MAXG], nw1[OBS_ROW_MAXG];
  issue(env_c, pos_c, cw0, cw1);

  for (unsigned k = 0; k < pairs_per_group; ++k) {  // uniform trip count: both halves of a warp stay in step
    const unsigned long long pair = first + k;
    issue(env_n, pos_n, nw0, nw1);                 // pair k+1: in flight during the work below
    const int pos_nn = load_pos(pair + 2);         // pair k+2
    float* dst = pos_c >= 0 ? out + (size_t)pair * d.obs_size : nullptr;
    const int head = dst ? (4 - (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3)) & 3 : 0;
    const int off = (4 - head) & 3;
    if (dst && lane < d.S) {
      // ---- scatter row `lane` of the window from the words loaded one iteration ago (scatter_rows_onehot_impl) ----
      const int y = d.inv_W ? (int)__umulhi((uint32_t)pos_c, d.inv_W) : pos_c / d.W, x = pos_c - y * d.W;
      const int x0 = x - d.v, yy = y - d.v + lane;
      const int jlo = max(0, -x0), jhi = min(d.S, d.W - x0);
      const bool row_in = (unsigned)yy < (unsigned)d.H;
      const int boff = yy * d.W + x0;
      uint16_t* const srow = stage16 + off + lane * d.S;
#pragma unroll
      for (int g = 0; g < OBS_ROW_MAXG; ++g) {
        if (g < ng) {
          const int j0 = 5 * g;
          const int lc = min(max(jlo - j0, 0), 4), hc = min(max(jhi - j0, 0), 4);
          uint32_t m4 = __funnelshift_lc(~0u, 0u, 8 * hc) & ~__funnelshift_lc(~0u, 0u, 8 * lc);
          bool in5 = jlo <= j0 + 4 && j0 + 4 < jhi;
          if (!row_in) m4 = 0u, in5 = false;
          const uint32_t sh = ((uint32_t)(boff + j0) & 3u) * 8u;
          uint32_t k4 = __funnelshift_r(cw0[g], cw1[g], sh);
          const uint32_t k5 = in5 ? ((cw1[g] >> sh) & 7u) : 7u;
          k4 = (k4 & m4) | (0x07070707u & ~m4);
          const uint32_t t = k4 & 0x07070707u;
          const uint32_t sel = __byte_perm(t | (t >> 4), 0u, 0x4420);
          uint32_t c4 = __byte_perm(tlo, thi, sel);
          uint32_t c5 = __byte_perm(tlo, thi, k5) & 0xFFu;
          if (PARTIAL) {
            const int cnt = d.S - j0;
            c4 |= ~__funnelshift_lc(~0u, 0u, 8 * min(cnt, 4));
            if (cnt < 5) c5 = 255u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t code = (c4 >> (8 * u)) & 0xFFu;
            if (code != 255u) srow[code * d.SS + j0 + u] = 0x3F80;
          }
          if (c5 != 255u) srow[c5 * d.SS + j0 + 4] = 0x3F80;
        }
      }
    }
    __syncwarp();
    if (dst) {  // ---- stream the row out (bf16 -> f32, 128-bit streaming stores), zeroing the staging row ----
      const int nvec = d.oh_size > head ? (d.oh_size - head) >> 2 : 0;
      const int etail = head + 4 * nvec + lane;
      if (lane < head && lane < d.oh_size) {
        dst[lane] = __uint_as_float((uint32_t)stage16[off + lane] << 16);
        stage16[off + lane] = 0;
      }
      uint2* s2 = reinterpret_cast<uint2*>(stage16 + off + head);
      for (int qv = lane; qv < nvec; qv += LANES) {
        const uint2 w = s2[qv];
        s2[qv] = make_uint2(0u, 0u);
        st_f4(dst + head + 4 * qv, __uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u),
              __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xFFFF0000u));
      }
      if (etail < d.oh_size) {
        dst[etail] = __uint_as_float((uint32_t)stage16[off + etail] << 16);
        stage16[off + etail] = 0;
      }
    }
    __syncwarp();
    // rotate the pipeline
#pragma unroll
    for (int g = 0; g < OBS_ROW_MAXG; ++g) cw0[g] = nw0[g], cw1[g] = nw1[g];
    pos_c = pos_n;
    pos_n = pos_nn;
    env_c = env_n, a_c = a_n;
    advance(env_n, a_n);
  }
}

// ----------------------------------------------------------------------------------------------
// Standalone observation with BULK-ASYNC stores (one-hot worlds with one stored plane, windows up to 16 wide, no
// appended floats).  The row-wise kernel above moves every row through LDS -> convert -> STG.128 and zeroes its bf16
// staging row on the way: ~8 shared-memory and 8 global-store instructions per lane per row, and ncu shows it bound
// by exactly that (mio_throttle, 18 % issue-slot use, 58 % of the HBM peak).  Here a group of 16 lanes scatters the
// ones of M consecutive rows (M = 1, 2 or 4: the smallest count whose bytes are a multiple of 16; 9x9 windows: two rows
// = 3888 B) as f32 into one of two shared-memory tiles and ONE lane hands the finished tile to the bulk-copy engine
// (`cp.async.bulk.global.shared::cta`, tracked by bulk groups): the rows leave the SM as one contiguous 16-byte-aligned
// burst with no per-element instruction, while the group already gathers the next tile into the other buffer.  A tile
// is zeroed (128-bit shared stores) only when it comes up for reuse, after `cp.async.bulk.wait_group.read` says the
// engine has finished reading it.  The window loads stay software-pipelined (pair k+1 / position k+2 in flight).
// ----------------------------------------------------------------------------------------------
#ifndef SB_OBS_BULK
#define SB_OBS_BULK 1
#endif
constexpr int OBS_BULK_THREADS = 128;
__device__ __forceinline__ void bulk_store_tile(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
template <bool PARTIAL>
__global__ void __launch_bounds__(OBS_BULK_THREADS)
    observe_bulk_kernel(const __grid_constant__ SbDev d, const uint8_t* __restrict__ grid,
                        const int32_t* __restrict__ agent_pos, float* __restrict__ out, int rows_per_tile,
                        unsigned pairs_per_stream, long long grid_bytes) {
  // every 16-lane group walks its own STREAM of consecutive tiles.  (A form in which the two halves of a warp fill the
  // even / odd rows of one tile -- half the shared memory per warp -- was measured: 9x9 windows 0.142 -> 0.137 ms, 5x5
  // windows 0.182 -> 0.202 ms; not kept.)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int LANES = OBS_ROW_LANES;
  const int lane = threadIdx.x & (LANES - 1);
  const int M = rows_per_tile, tile_floats = M * d.obs_size;  // tile_floats * 4 is a multiple of 16
  const int stream_in_block = (int)(threadIdx.x / LANES);
  constexpr int streams_per_block = OBS_BULK_THREADS / LANES;
  const int zl = lane;
  constexpr int zn = LANES;  // lanes that zero / copy a tile together
  const bool elected = zl == 0;
  float* const tiles = reinterpret_cast<float*>(smem_raw) + (size_t)stream_in_block * 2 * tile_floats;
  for (int e = zl; e < 2 * tile_floats / 4; e += zn) reinterpret_cast<float4*>(tiles)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  const unsigned long long pairs = (unsigned long long)d.N * d.A;
  const unsigned long long stream_first =
      ((unsigned long long)blockIdx.x * streams_per_block + stream_in_block) * pairs_per_stream;
  const unsigned long long first = stream_first;
  const int ng = (d.S + 4) / 5;
  const uint32_t tlo = (uint32_t)d.channel8_oob, thi = (uint32_t)(d.channel8_oob >> 32);

  auto load_pos = [&](unsigned long long pair) -> int { return pair < pairs ? __ldg(agent_pos + pair) : -1; };
  auto issue = [&](unsigned long long env, int pos, uint32_t (&w0)[OBS_ROW_MAXG], uint32_t (&w1)[OBS_ROW_MAXG]) {
    const int y = pos < 0 ? 0 : (d.inv_W ? (int)__umulhi((uint32_t)pos, d.inv_W) : pos / d.W);
    const int x = pos < 0 ? 0 : pos - y * d.W;
    const int yy = y - d.v + lane;
    const bool row_in = pos >= 0 && lane < d.S && (unsigned)yy < (unsigned)d.H;
    const long long base = (long long)env * d.env_stride + (long long)yy * d.W + (x - d.v);
#pragma unroll
    for (int g = 0; g < OBS_ROW_MAXG; ++g) {
      w0[g] = 0u, w1[g] = 0u;
      if (g < ng && row_in) {
        const long long al = (base + 5 * g) & ~3ll;
        if (al >= 0) w0[g] = __ldg(reinterpret_cast<const uint32_t*>(grid + al));
        if (al + 8 <= grid_bytes) w1[g] = __ldg(reinterpret_cast<const uint32_t*>(grid + al + 4));
      }
    }
  };
  auto advance = [&](unsigned long long& env, int& a) {
    if (++a == d.A) {
      a = 0;
      ++env;
    }
  };

  unsigned long long env_c = first / (unsigned)d.A;
  int a_c = (int)(first - env_c * (unsigned)d.A);
  unsigned long long env_n = env_c;
  int a_n = a_c;
  advance(env_n, a_n);
  int pos_c = load_pos(first), pos_n = load_pos(first + 1);
  uint32_t cw0[OBS_ROW_MAXG], cw1[OBS_ROW_MAXG], nw0[OBS_ROW_MAXG], nw1[OBS_ROW_MAXG];
  issue(env_c, pos_c, cw0, cw1);

  const int steps_per_tile = M;            // rows of a tile
  const unsigned steps = pairs_per_stream;  // pairs this group walks
  int rs = 0;          // step within the current tile
  unsigned tile = 0;   // tiles this stream has started
  for (unsigned k = 0; k < steps; ++k) {  // uniform trip count: both halves of a warp stay in step
    const unsigned long long pair = first + k;
    float* const buf = tiles + (tile & 1u) * tile_floats;
    if (rs == 0 && tile >= 2) {
      // this buffer held tile - 2: the engine must have finished READING it before it is zeroed and refilled
      if (elected) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      for (int e = zl; e < tile_floats / 4; e += zn) reinterpret_cast<float4*>(buf)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
    }
    issue(env_n, pos_n, nw0, nw1);  // the next pair: in flight during the work below
    const int pos_nn = load_pos(pair + 2);  // ... and the one after
    if (pos_c >= 0 && lane < d.S) {
      // ---- scatter row `lane` of the window (scatter_rows_onehot_impl on the words loaded one iteration ago) ----
      const int y = d.inv_W ? (int)__umulhi((uint32_t)pos_c, d.inv_W) : pos_c / d.W, x = pos_c - y * d.W;
      const int x0 = x - d.v, yy = y - d.v + lane;
      const int jlo = max(0, -x0), jhi = min(d.S, d.W - x0);
      const bool row_in = (unsigned)yy < (unsigned)d.H;
      const int boff = yy * d.W + x0;
      float* const srow = buf + rs * d.obs_size + lane * d.S;
#pragma unroll
      for (int g = 0; g < OBS_ROW_MAXG; ++g) {
        if (g < ng) {
          const int j0 = 5 * g;
          const int lc = min(max(jlo - j0, 0), 4), hc = min(max(jhi - j0, 0), 4);
          uint32_t m4 = __funnelshift_lc(~0u, 0u, 8 * hc) & ~__funnelshift_lc(~0u, 0u, 8 * lc);
          bool in5 = jlo <= j0 + 4 && j0 + 4 < jhi;
          if (!row_in) m4 = 0u, in5 = false;
          const uint32_t sh = ((uint32_t)(boff + j0) & 3u) * 8u;
          uint32_t k4 = __funnelshift_r(cw0[g], cw1[g], sh);
          const uint32_t k5 = in5 ? ((cw1[g] >> sh) & 7u) : 7u;
          k4 = (k4 & m4) | (0x07070707u & ~m4);
          const uint32_t t = k4 & 0x07070707u;
          const uint32_t sel = __byte_perm(t | (t >> 4), 0u, 0x4420);
          uint32_t c4 = __byte_perm(tlo, thi, sel);
          uint32_t c5 = __byte_perm(tlo, thi, k5) & 0xFFu;
          if (PARTIAL) {
            const int cnt = d.S - j0;
            c4 |= ~__funnelshift_lc(~0u, 0u, 8 * min(cnt, 4));
            if (cnt < 5) c5 = 255u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t code = (c4 >> (8 * u)) & 0xFFu;
            if (code != 255u) srow[code * d.SS + j0 + u] = 1.0f;
          }
          if (c5 != 255u) srow[c5 * d.SS + j0 + 4] = 1.0f;
        }
      }
    }
    if (++rs == steps_per_tile) {  // the tile is complete
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the scattered ones, seen by the copy engine
      __syncwarp();
      const unsigned long long t0 = stream_first + (unsigned long long)tile * (unsigned)M;  // first pair of the tile
      if (t0 + (unsigned)M <= pairs) {
        if (elected) bulk_store_tile(out + (size_t)t0 * d.obs_size, buf, (uint32_t)tile_floats * 4u);
      } else if (t0 < pairs) {  // the last tile of the batch is short: plain stores of its valid rows
        const int n_valid = (int)(pairs - t0) * d.obs_size;
        for (int e = zl; e < n_valid; e += zn) out[(size_t)t0 * d.obs_size + e] = buf[e];
      }
      if (elected) asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // (an empty group keeps the count)
      rs = 0;
      ++tile;
    }
    // rotate the pipeline
#pragma unroll
    for (int g = 0; g < OBS_ROW_MAXG; ++g) cw0[g] = nw0[g], cw1[g] = nw1[g];
    pos_c = pos_n;
    pos_n = pos_nn;
    env_c = env_n, a_c = a_n;
    advance(env_n, a_n);
  }
  // shared memory must outlive the reads of the copy engine
  if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

int launch_observe(const SbDev& d, const uint8_t* grid, const int32_t* agent_pos, float* out, cudaStream_t s) {
  const bool np1 = d.fast_obs && d.n_planes == 1;
  // row-wise window gather (one lane per window row) when the rows fit the 16 lanes of a half warp
  // (measured: 9x9 windows 0.589 -> 0.608 of the HBM peak, 5x5 windows 0.72 -> 0.70: only for the wider ones)
  const bool rows = np1 && d.rowwise_ok && d.S >= 7 && d.S <= 16 && SB_OBS_ROWWISE;
  const int lanes = (d.obs_size <= 256 || rows) ? 16 : 32;  // two pairs per warp / one
  const int groups = OBS_THREADS / lanes;
  const int stage_bytes = round_up((d.obs_size + 4) * 2, 16);
  const int smem = round_up(2 * d.SS, 16) + 4 * SB_KTAB_WORDS + groups * stage_bytes;
  if (smem > 227 * 1024) return SB_EUNSUPPORTED;
  const unsigned long long pairs = (unsigned long long)d.N * d.A;
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return note_cuda_error(e);
    }
    // persistent grid of exactly ONE wave: (SMs) x (blocks that are resident together, from the occupancy
    // calculator: registers cap it at 6, not 8); each group takes a contiguous run of pairs.  A fixed 148 x 8
    // grid ran 1.33 waves, i.e. a second, one-third-full wave (ncu: launch__waves_per_multiprocessor).
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, OBS_THREADS, smem) != cudaSuccess || per_sm < 1)
      per_sm = 4;
    unsigned long long blocks = (pairs + groups - 1) / groups;
    if (blocks > (unsigned long long)sms * per_sm) blocks = (unsigned long long)sms * per_sm;
    const unsigned long long ppg = (pairs + blocks * groups - 1) / (blocks * groups);
    if (ppg > 0xFFFFFFFFull) return SB_EUNSUPPORTED;
    kern<<<(unsigned)blocks, OBS_THREADS, smem, s>>>(d, grid, agent_pos, out, stage_bytes, d.fast_obs, (unsigned)ppg);
    count_launch();
    return check_launch();
  };
  if (SB_OBS_BULK && np1 && d.rowwise_ok && d.S <= 16 && !d.obs_extra && !d.obs_flag &&
      (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {  // bulk-async stores of whole tiles of rows
    const int M = (d.obs_size & 3) == 0 ? 1 : ((d.obs_size & 1) == 0 ? 2 : 4);  // rows per 16-byte-multiple tile
    const int streams = OBS_BULK_THREADS / OBS_ROW_LANES;  // tile streams per block
    const int sm = streams * 2 * M * d.obs_size * 4;
    // f32 tiles cost shared memory: the bulk form wins while ~9 blocks stay resident (5x5 windows, 1200-byte tiles:
    // 0.224 -> 0.182 ms = 0.75 -> 0.92 of the HBM peak at 2^20 x 1; 0.407 -> 0.332 ms at 262144 x 8); with 3888-byte tiles
    // (9x9 windows) three blocks fit and it loses to the bf16 row kernel (0.142 against 0.137 ms)
    if (sm <= 24 * 1024) {
      auto launch_bulk = [&](auto kern) -> int {
        if (sm > 48 * 1024) {
          const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
          if (e != cudaSuccess) return note_cuda_error(e);
        }
        int dev = 0, sms = 148, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, OBS_BULK_THREADS, sm) != cudaSuccess || per_sm < 1)
          per_sm = 2;
        unsigned long long blocks = (pairs + streams - 1) / streams;
        if (blocks > (unsigned long long)sms * per_sm) blocks = (unsigned long long)sms * per_sm;
        unsigned long long pps = (pairs + blocks * streams - 1) / (blocks * streams);
        pps = (pps + M - 1) / M * M;  // whole tiles per stream: tiles never straddle two streams
        if (pps > 0xFFFFFFFFull) return SB_EUNSUPPORTED;
        blocks = (pairs + pps * streams - 1) / (pps * streams);
        kern<<<(unsigned)blocks, OBS_BULK_THREADS, sm, s>>>(d, grid, agent_pos, out, M, (unsigned)pps,
                                                            (long long)d.N * d.env_stride);
        count_launch();
        return check_launch();
      };
      return d.S5 == d.S ? launch_bulk(observe_bulk_kernel<false>) : launch_bulk(observe_bulk_kernel<true>);
    }
  }
  if (rows && !d.obs_extra && !d.obs_flag && SB_OBS_PIPELINED) {  // software-pipelined row gather
    const int sb = round_up((d.obs_size + 4) * 2, 16);
    const int sm = groups * sb;
    auto launch_rows = [&](auto kern) -> int {
      if (sm > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        if (e != cudaSuccess) return note_cuda_error(e);
      }
      int dev = 0, sms = 148, per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, OBS_THREADS, sm) != cudaSuccess || per_sm < 1)
        per_sm = 4;
      unsigned long long blocks = (pairs + groups - 1) / groups;
      if (blocks > (unsigned long long)sms * per_sm) blocks = (unsigned long long)sms * per_sm;
      const unsigned long long ppg = (pairs + blocks * groups - 1) / (blocks * groups);
      if (ppg > 0xFFFFFFFFull) return SB_EUNSUPPORTED;
      kern<<<(unsigned)blocks, OBS_THREADS, sm, s>>>(d, grid, agent_pos, out, sb, (unsigned)ppg,
                                                     (long long)d.N * d.env_stride);
      count_launch();
      return check_launch();
    };
    return d.S5 == d.S ? launch_rows(observe_rows_kernel<false>) : launch_rows(observe_rows_kernel<true>);
  }
  if (rows) return launch(observe_kernel<true, 16, 1>);
  // windows wider than 16 (19x19: the other reading of config 4's "9x9 vision radius"): 32 lanes per pair, each taking
  // row SEGMENTS (scatter_rows_onehot_seg: 19x19 = 76 units of 5 / 4 cells, three single-step trips) instead of 361
  // cell-wise visits
  if (SB_OBS_SEG && np1 && d.rowwise_ok && d.S > 16 && lanes == 32) return launch(observe_kernel<true, 32, 2>);
  if (lanes == 16) return np1 ? launch(observe_kernel<true, 16>) : launch(observe_kernel<false, 16>);
  return np1 ? launch(observe_kernel<true, 32>) : launch(observe_kernel<false, 32>);
}

#endif  // SB_ONLY_LPE

// Resolve the queued spawn candidates of a warp, all 32 lanes at once: entry = group << 20 | cell.
__device__ __forceinline__ void drain_queue(const SbDev& d, const uint32_t* q, uint32_t n, uint8_t* envs, int esz,
                                            long long env0, uint32_t tick, const uint32_t* __restrict__ rk) {
  for (uint32_t i = threadIdx.x & 31u; i < n; i += 32) {
    const uint32_t entry = q[i];
    const uint32_t g = entry >> 20, cell = entry & 0xFFFFFu;
    const uint32_t eid = (uint32_t)(d.env_id_base + (uint64_t)(env0 + g));
    const int kind = sb_resolve_candidate(d, eid, tick, cell, rk);
    if (kind >= 0) envs[(size_t)g * esz + cell] = (uint8_t)kind;
  }
}

// The same, out of line: called from inside the sweep loop only when the queue is about to overflow (dense
// spawning), so that the two Philox bodies of the resolution exist ONCE in the hot instruction stream.
static __device__ __noinline__ void drain_queue_cold(const SbDev* d, const uint32_t* q, uint32_t n, uint8_t* envs, int esz,
                                              long long env0, uint32_t tick, const uint32_t* rk) {
  drain_queue(*d, q, n, envs, esz, env0, tick, rk);
}

// LPE  lanes per instance; RNG  SB_K_* sweep variant; NP1  one stored plane + one-hot fast path;
// STAGED  planes live in shared memory for the whole launch (compile-time, so that every access to the
// instance image is an LDS / STS with 32-bit addressing instead of a generic load); false only with LPE == 32
#ifndef SB_MIN_BLOCKS
#define SB_MIN_BLOCKS 8  // resident blocks per SM the register allocation is sized for (64 registers)
#endif
#ifndef SB_MIN_BLOCKS_EXT
#define SB_MIN_BLOCKS_EXT 6  // replay / generic variants: shared memory caps their residency anyway (cleanup: 6 blocks), so they
                             // get 80 registers instead of 64 (measured: cleanup 8.60 -> 8.13 ms; 5 blocks / 87 registers: 8.59)
#endif
template <int LPE, int RNG, bool NP1, bool STAGED>
__global__ void __launch_bounds__(SB_THREADS, (RNG == SB_K_TAPE || RNG == SB_K_PHILOX_GENERIC) ? SB_MIN_BLOCKS_EXT
                                                                                             : SB_MIN_BLOCKS)
    step_kernel(const __grid_constant__ SbDev d, const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int EPB = SB_THREADS / LPE;
  constexpr int GPW = 32 / LPE;  // groups per warp
  // facings / beams / positional code are compiled only into the replay and generic variants: a world that
  // uses them never takes the byte-SIMD or sparse fast paths (make_dev clears simd_sweep / sparse_ok)
  constexpr bool EXT = RNG == SB_K_TAPE || RNG == SB_K_PHILOX_GENERIC;
  const int tid = threadIdx.x;
  const int lane = tid & (LPE - 1);
  const int grp = tid / LPE;
  const int wlane = tid & 31;
  const int warp = tid >> 5;

  // ---- carve shared memory --------------------------------------------------------------
  uint16_t* lut_ij = reinterpret_cast<uint16_t*>(smem_raw);  // [SS] (i << 8) | j of window cell ij
  uint32_t* ktab = reinterpret_cast<uint32_t*>(smem_raw + p.q_off) - SB_KTAB_WORDS;  // per-kind tables (sb_core.cuh)
  if ((!NP1 || EXT) && tid < 32) sb_fill_kind_tables(d, ktab, tid);  // (the one-plane fast variants never read it)
  uint32_t* qbase = reinterpret_cast<uint32_t*>(smem_raw + p.q_off);
  uint32_t* q = qbase + warp * (SB_QCAP + 1);  // spawn candidates of this warp's instances
  // planes: staged in shared memory for the whole launch, or (large maps, p.staged == 0) used in
  // place in global memory so that shared memory does not cap the number of resident warps
  const long long env0 = (long long)blockIdx.x * EPB;
  const int esz = STAGED ? p.plane_env : d.env_stride;
  uint8_t* envs;
  if constexpr (STAGED) envs = smem_raw + p.planes_off;
  else envs = p.grid + (size_t)env0 * d.env_stride;
  uint8_t* genv = envs + (STAGED ? grp * esz : (size_t)grp * esz);
  uint32_t* spos = reinterpret_cast<uint32_t*>(smem_raw + p.pos_off) + grp * p.agent_words;
  int32_t* sfacing = reinterpret_cast<int32_t*>(spos + d.A);  // MovingAgent.direction (only read when d.ext_act)
  float* srew = reinterpret_cast<float*>(sfacing + d.A);      // reward of every agent's move this turn
  int32_t* sact_win = reinterpret_cast<int32_t*>(srew + d.A);  // actions of the next act_turns turns: [turn][agent]
  uint8_t* stage_raw = smem_raw + p.stage_off + grp * p.stage_env;

  const long long env = env0 + grp;
  const bool active = env < d.N;
  uint4* genv4 = reinterpret_cast<uint4*>(genv);
#if SB_ASYNC_STAGE
  // ---- stage the instance: asynchronous copies (global -> shared without a register round trip) issued before the
  // block's set-up work, waited for after it; and an L2 prefetch of the planes of the block that will follow this one
  // on the same SM slot.  Matters for short launches (T = 1: the staging load was 22 % of the stall samples).
  if (STAGED && active) {
    const uint4* src = reinterpret_cast<const uint4*>(p.grid + (size_t)env * d.env_stride);
    for (int c = lane; c < d.n_chunks; c += LPE) cp_async16(genv4 + c, src + c);
  }
  if (STAGED && p.prefetch_blocks) {
    const size_t block_bytes = (size_t)EPB * d.env_stride, all_bytes = (size_t)d.N * d.env_stride;
    const size_t at = ((size_t)blockIdx.x + (size_t)p.prefetch_blocks) * block_bytes;
    for (size_t o = (size_t)tid * 128; o < block_bytes; o += (size_t)SB_THREADS * 128)
      if (at + o < all_bytes) prefetch_l2(p.grid + at + o);
  }
#endif

  for (int ij = tid; ij < d.SS; ij += SB_THREADS) {
    int i = ij / d.S;
    lut_ij[ij] = (uint16_t)((i << 8) | (ij - i * d.S));
  }
  for (int e = lane; e < p.stage_env / 4; e += LPE) reinterpret_cast<uint32_t*>(stage_raw)[e] = 0u;
  __syncthreads();

  if (env0 + (long long)warp * GPW >= d.N) return;  // whole warp idle (only warp-level syncs below)
  const uint32_t envid = (uint32_t)(d.env_id_base + (uint64_t)env);
  const uint32_t* __restrict__ rk = p.rk;

  if (active) {
#if SB_ASYNC_STAGE
    if (STAGED) cp_async_wait_all();
#else
    if (STAGED) {
      const uint4* src = reinterpret_cast<const uint4*>(p.grid + (size_t)env * d.env_stride);
      for (int c = lane; c < d.n_chunks; c += LPE) genv4[c] = __ldg(src + c);
    }
#endif
    for (int a = lane; a < d.A; a += LPE) {
      int pos = p.agent_pos[(size_t)env * d.A + a];
      int y = pos / d.W;
      spos[a] = ((uint32_t)y << 16) | (uint32_t)(pos - y * d.W);
      if (EXT) sfacing[a] = p.a.d_agent_facing ? p.a.d_agent_facing[(size_t)env * d.A + a] : 2;
    }
  } else {  // lanes past the last instance: never touch a grid (every access below is guarded by `active`)
    if (STAGED)
      for (int c = lane; c < d.n_chunks; c += LPE) genv4[c] = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (int a = lane; a < d.A; a += LPE) spos[a] = 0, sfacing[a] = 0;
  }
  double total = (active && lane == 0 && p.write_back) ? p.total_reward[env] : 0.0;
  uint32_t err = 0;
  __syncwarp();

  const int T = p.a.num_turns;
  const size_t rows = (size_t)d.N * d.A;
  const size_t tape_stride = (size_t)d.env_stride;  // tape entries per instance per turn
  const bool fused = p.a.mode == SB_MODE_FUSED;
  const size_t row0 = (size_t)(active ? env : 0) * d.A;  // first ring row of this instance
  // observation rows go straight into the ring: element (row, slot) at row * row_stride + slot * slot_stride
  float* const ring0 = p.has_buf ? p.buf.d_states + row0 * (size_t)p.buf.states_row_stride : nullptr;
  auto ring_row = [&](int a, int slot_) -> float* {
    return ring0 + (size_t)a * (size_t)p.buf.states_row_stride + (size_t)slot_ * (size_t)p.buf.states_slot_stride;
  };
  // {tick0, slot0} from the launch arguments, or from device memory for graph-replayed launches
  const uint32_t tick0 = p.a.d_clock ? p.a.d_clock[0] : p.a.tick0;
  int slot = p.has_buf ? (p.a.d_clock ? (int)p.a.d_clock[1] : p.a.slot0) : 0;

  int tw = 0;  // row of the action window this turn reads
  for (int t = 0; t < T; ++t, slot = (p.has_buf && slot + 1 < p.buf.capacity) ? slot + 1 : 0) {
    const uint32_t tick = tick0 + (uint32_t)t;

    // ---- actions (parallel over (turn, agent) pairs; hides the tape-load latency) -------------------
    // sact points at this turn's row of the action window; the window is refilled every act_turns turns
    if (tw == p.act_turns) tw = 0;
    int32_t* const sact = sact_win + tw * d.A;
    if ((fused || p.a.act_agent >= 0) && tw == 0) {
      for (int e = lane; e < p.act_turns * d.A; e += LPE) {
        const int dt = p.act_turns > 1 ? (int)(((uint32_t)e * p.inv_A) >> 16) : 0;
        const int a = e - dt * d.A;
        int act = 0;
        if (t + dt < T) {
          if (p.a.policy == SB_POLICY_UNIFORM) {
            SbU4 w = sb_philox_rk(envid, tick + (uint32_t)dt, (uint32_t)a, SB_DOM_ACTION, rk);
            act = (int)sb_bounded(w.x, w.y, (uint32_t)d.n_actions);
          } else if (active && p.a.d_actions) {
            {
              const size_t at = ((size_t)(t + dt) * d.N + env) * d.A + a;
              act = p.a.actions_u8 ? (int)reinterpret_cast<const uint8_t*>(p.a.d_actions)[at] : p.a.d_actions[at];
            }
          }
        }
        sact_win[e] = act;
      }
    }
    ++tw;

    // ---- entity sweep ----------------------------------------------------------------------
#ifdef SB_EXP_NOSWEEP  // timing experiment only (wrong results): what the turn costs without the sweep
    if (false) {
#else
    if (fused || p.a.do_sweep) {
#endif
      if (RNG == SB_K_TAPE) {
        // Replay, all rule types.  Planes in ascending layer order with a warp barrier in between: a column
        // rule then sees lower layers after and higher layers before their own transition of this turn,
        // like the C-order sweep of environment.py:88-91 (rules only read their own column).
        const double* tu = p.a.d_spawn_u + ((size_t)t * d.N + (active ? env : 0)) * tape_stride;
        const uint8_t* tc = p.a.d_spawn_c + ((size_t)t * d.N + (active ? env : 0)) * tape_stride;
        uint8_t* drew = p.a.d_drew_out ? p.a.d_drew_out + ((size_t)t * d.N + (active ? env : 0)) * tape_stride
                                       : nullptr;
        for (int pl = 0; pl < d.n_planes; ++pl) {
          const int base = pl * d.plane_stride;
          for (int off = lane; off < d.plane_stride && active; off += LPE) {
            const int cell = base + off;
            const uint32_t kind = genv[cell];
            const uint32_t rule = kind < (uint32_t)SB_MAX_KINDS ? d.rule[kind] : (uint32_t)SB_RULE_NONE;
            const bool draws = rule == SB_RULE_SPAWN || rule == SB_RULE_SPAWN_TO;
            const double u = tu[cell];
            if (drew) drew[cell] = draws ? 1 : 0;
            if (draws) {
              if (u != u) {
                err |= SB_ERR_TAPE_ALIGN;  // the reference drew here, the tape did not
              } else if (u < d.prob[kind]) {  // np.random.random() < prob (entities.py:72-74, cleanup/entities.py:52,86)
                if (rule == SB_RULE_SPAWN) {
                  const uint32_t ci = tc[cell];
                  if (ci < (uint32_t)d.n_choices) genv[cell] = (uint8_t)((d.choices_packed >> (8 * ci)) & 0xFF);
                  else err |= SB_ERR_TAPE_ALIGN;
                } else {
                  genv[cell] = d.rule_to[kind];
                }
              } else if (rule == SB_RULE_SPAWN_TO && d.rule_else[kind] != SB_NO_KIND) {
                genv[cell] = d.rule_else[kind];  // iowa Deck.draw: the no-loss value state
              }
            } else {
              if (u == u) err |= SB_ERR_TAPE_ALIGN;  // tape holds a draw for a cell that has no drawing rule
              if (rule != SB_RULE_NONE) {
                const int nk = sb_column_rule(d, genv, kind, rule, off);
                if (nk >= 0) genv[cell] = (uint8_t)nk;
              }
            }
          }
          __syncwarp();
        }
      } else if (RNG == SB_K_PHILOX_SIMD) {
        // First level, byte-SIMD: one Philox call covers the 16 cells of a chunk; cells whose random byte is below
        // the filter bound AND that carry the spawn rule are rare (about 2 per instance-turn at p = 0.005), so they
        // are appended to a per-warp queue -- positions from a ballot + popcount, the count lives in a (warp-
        // uniform) register, no shared-memory atomics -- and resolved by all 32 lanes at once after the loop
        // instead of diverging inside it.  Two chunks per trip: two independent Philox chains in flight, and one
        // compaction round usually serves both (x4 / x8 no longer fit the instruction cache, profiles/README.md).
        const uint32_t spawn4 = (uint32_t)d.single_spawn_kind * 0x01010101u;
        uint32_t qcount = 0;
        for (int base = 0; base < d.n_chunks; base += 2 * LPE) {  // uniform trip count: ballots inside
          const int c0 = base + lane, c1 = c0 + LPE;
          uint32_t m0 = 0, m1 = 0;
          if (active && c0 < d.n_chunks) {
            const uint4 k = genv4[c0];
            const SbU4 r = sb_philox_rk(envid, tick, (uint32_t)c0, SB_DOM_SWEEP, rk);
            m0 = sb_simd_cand_chunk(k.x, k.y, k.z, k.w, r.x, r.y, r.z, r.w, spawn4, d.hmask4);
            if (c1 < d.n_chunks) {
              const uint4 k1 = genv4[c1];
              const SbU4 r1 = sb_philox_rk(envid, tick, (uint32_t)c1, SB_DOM_SWEEP, rk);
              m1 = sb_simd_cand_chunk(k1.x, k1.y, k1.z, k1.w, r1.x, r1.y, r1.z, r1.w, spawn4, d.hmask4);
            }
          }
          uint32_t b = __ballot_sync(0xFFFFFFFFu, (m0 | m1) != 0);
          while (b) {  // one candidate per lane per round (a second round is needed in < 40 % of the trips)
            if (qcount + 32u > (uint32_t)SB_QCAP) {  // dense spawning: make room (warp-uniform branch, cold)
              __syncwarp();
              drain_queue_cold(&d, q, qcount, envs, esz, env0, tick, rk);
              __syncwarp();
              qcount = 0;
            }
            if (m0 | m1) {
              const bool first = m0 != 0;
              const uint32_t m = first ? m0 : m1;
              const int pos = __ffs(m) - 1;  // bit 8*b + w: byte b of word w of the chunk
              if (first) m0 &= m0 - 1;
              else m1 &= m1 - 1;
              const uint32_t cell = (uint32_t)((first ? c0 : c1) * 16 + 4 * (pos & 7) + (pos >> 3));
              q[qcount + __popc(b & ((1u << wlane) - 1u))] = ((uint32_t)grp << 20) | cell;
            }
            qcount += __popc(b);
            b = __ballot_sync(0xFFFFFFFFu, (m0 | m1) != 0);
          }
        }
        __syncwarp();
        drain_queue(d, q, qcount, envs, esz, env0, tick, rk);
        __syncwarp();
      } else if (RNG == SB_K_SPARSE) {
        // Sparse sampling of the same Bernoulli field: K ~ Binomial(cells, p) by inverse CDF, then K
        // distinct uniform cells (sequential rejection over the draw stream j = 1, 2, ...).  The lanes of
        // the group draw LPE stream entries at a time; lane 0 accepts them in stream order.
        uint16_t* hits = reinterpret_cast<uint16_t*>(smem_raw + p.hits_off) + grp * SB_SPARSE_KMAX;
        const SbU4 r0 = sb_philox_rk(envid, tick, 0u, SB_DOM_SPARSE, rk);
        const int K = active ? sb_sparse_count(d, ((uint64_t)r0.x << 32) | r0.y) : 0;
        int accepted = 0;
        uint32_t base = 1;
        while (__any_sync(0xFFFFFFFFu, accepted < K)) {
          const SbU4 r = sb_philox_rk(envid, tick, base + (uint32_t)lane, SB_DOM_SPARSE, rk);
          const uint32_t pos = (uint32_t)sb_mulhi64(((uint64_t)r.x << 32) | r.y, (uint64_t)d.env_stride);
          const uint32_t ci = sb_bounded(r.z, r.w, (uint32_t)d.n_choices);
#pragma unroll
          for (int l = 0; l < LPE; ++l) {
            const uint32_t pl = __shfl_sync(0xFFFFFFFFu, pos, l, LPE);
            const uint32_t cl = __shfl_sync(0xFFFFFFFFu, ci, l, LPE);
            if (lane == 0 && accepted < K) {
              bool dup = false;
              for (int h = 0; h < accepted; ++h) dup |= hits[h] == (uint16_t)pl;
              if (!dup) {
                hits[accepted++] = (uint16_t)pl;
                const uint32_t kind = genv[pl];
                if (kind < (uint32_t)SB_MAX_KINDS && ((d.spawn_mask >> kind) & 1u))
                  genv[pl] = (uint8_t)((d.choices_packed >> (8 * cl)) & 0xFFu);
              }
            }
          }
          accepted = __shfl_sync(0xFFFFFFFFu, accepted, 0, LPE);
          base += LPE;
        }
      } else {  // SB_K_PHILOX_GENERIC: several spawning kinds, a large probability, or column rules
        const int cpp = d.plane_stride / 16;
        const uint32_t ruled = d.draw_mask | d.det_mask;  // kinds that carry any rule
        for (int pl = 0; pl < d.n_planes; ++pl) {
          for (int cc = lane; cc < cpp && active; cc += LPE) {
            const int c = pl * cpp + cc;
            const uint4 k4 = genv4[c];
            const uint32_t kw[4] = {k4.x, k4.y, k4.z, k4.w};
            // classify the 16 cells from registers (common case: nothing to do), then visit only the cells where
            // something can happen in a ROLLED loop: the slow path (two Philox bodies) exists once in the
            // instruction stream
            uint32_t m = 0;
            bool have_r = false;
            SbU4 r;
            if (d.cls_ok) {
              // Byte-SIMD, four cells per word: two per-kind class bytes (A = level-1 mask of a drawing kind; O = bit 0
              // "does not draw", bit 7 "always visit": deterministic rules, draws with an else branch) are looked up
              // by byte permutes; a cell is visited iff it must always be, or it draws and its random byte passes
              // level 1 of its kind -- (r8 & A) | (O & 1) == 0.  A River / AppleTree cell (a third of the cleanup map
              // each) used to cost a trip of the loop below every turn; now 1 in 64 / 1 in 256 of them does.
              uint32_t ca[4], co[4], nd = 0xFFFFFFFFu;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const uint32_t t = kw[w] & 0x07070707u;
                const uint32_t sel = __byte_perm(t | (t >> 4), 0u, 0x4420);
                const uint32_t m8 = ((kw[w] >> 3) & 0x01010101u) * 0xFFu;  // kinds 8..15 (and the 0xFF padding)
                ca[w] = (__byte_perm((uint32_t)d.cls_and[0], (uint32_t)(d.cls_and[0] >> 32), sel) & ~m8) |
                        (__byte_perm((uint32_t)d.cls_and[1], (uint32_t)(d.cls_and[1] >> 32), sel) & m8);
                co[w] = (__byte_perm((uint32_t)d.cls_or[0], (uint32_t)(d.cls_or[0] >> 32), sel) & ~m8) |
                        (__byte_perm((uint32_t)d.cls_or[1], (uint32_t)(d.cls_or[1] >> 32), sel) & m8);
                nd &= co[w];
              }
              r.x = r.y = r.z = r.w = 0u;
              if ((~nd & 0x01010101u) != 0u) {  // some cell of the chunk draws
                r = sb_philox_rk(envid, tick, (uint32_t)c, SB_DOM_SWEEP, rk);
                have_r = true;
              }
              const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const uint32_t x = (rw[w] & ca[w]) | (co[w] & 0x01010101u);
                const uint32_t tz = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                const uint32_t vis = (~(tz | x | 0x7F7F7F7Fu)) | (co[w] & 0x80808080u);  // 0x80 per cell to visit
                m |= vis >> (7 - w);  // bit 8*b + w: byte b of word w
              }
            } else {
#pragma unroll
              for (int b = 0; b < 16; ++b) {
                const uint32_t kind = (kw[b >> 2] >> (8 * (b & 3))) & 0xFFu;
                if (kind < (uint32_t)SB_MAX_KINDS && ((ruled >> kind) & 1u)) m |= 1u << b;
              }
            }
#pragma unroll 1
            while (m) {
              const int pos = __ffs(m) - 1;
              m &= m - 1;
              const int b = d.cls_ok ? 4 * (pos & 7) + (pos >> 3) : pos;  // cell of the chunk
              const uint32_t cell = (uint32_t)(c * 16 + b);
              const uint32_t kind = genv[cell];
              const uint32_t kt = ktab[kind];  // channel | rule << 8 | rule_to << 16 | rule_else << 24
              const uint32_t rule = (kt >> 8) & 0xFFu;
              if ((d.draw_mask >> kind) & 1u) {
                if (!have_r) {
                  r = sb_philox_rk(envid, tick, (uint32_t)c, SB_DOM_SWEEP, rk);
                  have_r = true;
                }
                const uint32_t r8 = sb_byte_of(r, b);
                if (rule == SB_RULE_SPAWN) {
                  if (r8 >= d.n2) continue;  // level 1
                  const int nk = sb_resolve_candidate(d, envid, tick, cell, rk);
                  if (nk >= 0) genv[cell] = (uint8_t)nk;
                } else if (sb_bernoulli_hit(d.n2k[kind], d.thr64m1k[kind], r8, envid, tick, cell, rk)) {
                  genv[cell] = (uint8_t)(kt >> 16);
                } else if ((d.else_mask >> kind) & 1u) {  // register-resident mask: most failed draws change nothing
                  genv[cell] = (uint8_t)(kt >> 24);
                }
              } else {
                const int nk = sb_column_rule(d, genv, kind, rule, cc * 16 + b, ktab);
                if (nk >= 0) genv[cell] = (uint8_t)nk;
              }
            }
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();

    // ---- agents, sequentially in list order (environment.py:92) -------------------------------
    auto observe = [&](int a) {
      const uint32_t yx = spos[a];
      const int y = (int)(yx >> 16), x = (int)(yx & 0xFFFFu);
      float* dsts[2];
      dsts[0] = (p.has_buf && active) ? ring_row(a, slot) : nullptr;
      dsts[1] = (p.a.d_obs_out && active)
                    ? p.a.d_obs_out + (((size_t)t * d.N + env) * d.A + a) * d.obs_size
                    : nullptr;
      if (NP1 && LPE == 32 && !STAGED) {
        // large maps used in place in global memory: batch the window loads (one L2 round trip per row, not 3)
        emit_onehot_row<LPE, NP1, (NP1 && LPE == 32) ? 4 : 1>(d, genv, ktab, stage_raw, y, x, dsts[0], lane);
        if (p.a.d_obs_out) {
          __syncwarp();
          emit_onehot_row<LPE, NP1, (NP1 && LPE == 32) ? 4 : 1>(d, genv, ktab, stage_raw, y, x, dsts[1], lane);
        }
      } else if (NP1 || d.fast_obs) {
        // narrow groups on a staged plane gather the window row by row (scatter_rows_onehot)
        // (several staged planes: the column gather of scatter_cols_multi, any group width)
        constexpr int RW = (STAGED && (NP1 ? LPE <= SB_ROWWISE_MAX_LPE : SB_COLGATHER != 0)) ? 1 : 0;
        emit_onehot_row<LPE, NP1, 1, RW>(d, genv, ktab, stage_raw, y, x, dsts[0], lane);
        if (p.a.d_obs_out) {
          __syncwarp();
          emit_onehot_row<LPE, NP1, 1, RW>(d, genv, ktab, stage_raw, y, x, dsts[1], lane);
        }
      } else if (active) {
        for (int e = lane; e < d.oh_size; e += LPE) {
          const int c = e / d.SS;
          const uint32_t li = lut_ij[e - c * d.SS];
          float vf = (float)sb_obs_general(d, genv, d.tmpl, d.app, y, x, c, (int)(li >> 8), (int)(li & 255u));
          if (dsts[0]) dsts[0][e] = vf;
          if (dsts[1]) dsts[1][e] = vf;
        }
      }
      // every lane has read the grid and the agent's position before lane 0 moves the agent (the fast paths get
      // this from the barrier inside emit_onehot_row)
      if (!(NP1 || d.fast_obs)) __syncwarp();
      if (EXT && d.obs_extra && active) {  // positional code of the agent's cell (CleanupObservation.observe)
        const float* src = d.extra_tbl + (size_t)(y * d.W + x) * d.obs_extra;
        for (int e = lane; e < d.obs_extra; e += LPE) {
          const float vf = __ldg(src + e);
          if (dsts[0]) dsts[0][d.oh_size + e] = vf;
          if (dsts[1]) dsts[1][d.oh_size + e] = vf;
        }
      }
      if (EXT && d.obs_flag && active && lane == 0) {  // TagAgent.pov: [self.it] appended (tag/agents.py:55-65)
        const float vf = genv[d.agent_plane * d.plane_stride + y * d.W + x] == (uint8_t)d.agent_kind_alt ? 1.0f : 0.0f;
        if (dsts[0]) dsts[0][d.oh_size + d.obs_extra] = vf;
        if (dsts[1]) dsts[1][d.oh_size + d.obs_extra] = vf;
      }
    };

    // The scalar part of Buffer.add (buffers.py:57-62) and the optional per-turn outputs for agents [a_lo, a_hi):
    // one agent per LANE, after the agents of the turn have moved -- with A >= LPE every lane stores, where lane 0
    // used to issue five stores per agent while the others waited (neighbouring rows are contiguous in the slot-major
    // ring arrays, so the stores of a warp coalesce).
    auto record_turn = [&](int a_lo, int a_hi) {
      __syncwarp();
      for (int a = a_lo + lane; a < a_hi && active; a += LPE) {
        const uint32_t yx = spos[a];
        const int action = sact[a];
        const float reward = srew[a];
        if (p.has_buf) {
          const size_t at = (size_t)slot * rows + row0 + a;
          p.buf.d_actions[at] = (int64_t)action;
          p.buf.d_rewards[at] = reward;
          p.buf.d_dones[at] = 0.0f;  // world.is_done is False for the whole episode
          if (p.buf.d_positions)     // location AFTER the move (agent.py:131-134); 16-byte aligned pair
            *reinterpret_cast<longlong2*>(p.buf.d_positions + 2 * at) =
                make_longlong2((long long)(yx >> 16), (long long)(yx & 0xFFFFu));
        }
        const size_t o = ((size_t)t * d.N + env) * d.A + a;
        if (p.a.d_reward_out) p.a.d_reward_out[o] = reward;
        if (p.a.d_action_out) p.a.d_action_out[o] = action;
      }
      __syncwarp();  // the action window may be refilled right after
    };

    // one agent's move by ONE lane: MovingAgent.act + Gridworld.move
    auto move_and_record = [&](int a, int pre_tk, double* reward64 = nullptr) -> float {
      uint32_t yx = spos[a];
      const int action = sact[a];
      float reward;
      if (EXT && d.agent_kind_alt >= 0) {
        // the agent moves as the kind it is drawn with; then the role hand-over and the own-state reward (tag)
        const int own = genv[d.agent_plane * d.plane_stride + (int)(yx >> 16) * d.W + (int)(yx & 0xFFFFu)];
        reward = sb_agent_act(d, genv, &yx, action, true, err, d.reward_all, pre_tk, reward64, own);
        uint32_t mine = (uint32_t)own;
        if (d.touch_swap) mine = sb_touch_swap(d, genv, yx);
        if (d.reward_own) {
          reward = d.value[mine];
          if (reward64) *reward64 = d.value64[mine];
        }
      } else {
        reward = sb_agent_act(d, genv, &yx, action, true, err, EXT && d.reward_all, pre_tk, reward64);
      }
      spos[a] = yx;
      srew[a] = reward;  // written to the ring (with the action and the new position) by record_turn below
      return reward;
    };

    auto act = [&](int a, int pre_tk) {
      if (EXT && d.ext_act) {
        // MovingAgent.movement sets the facing (agents/agent.py:199-211); beam actions then replace every
        // in-map, non-blocked cell of the beam pattern on the beam layer (CleanupAgent.spawn_beam,
        // cleanup/agents.py:93-145) BEFORE the reward is read (agents.py:158-170).  Lanes share the cells.
        const int action = sact[a];
        const bool ok = (unsigned)action < (unsigned)d.n_actions;
        int facing = sfacing[a] & 3;
        if (ok && d.act_facing[action] >= 0) facing = d.act_facing[action];
        const uint32_t bk = ok ? (uint32_t)d.act_beam[action] : (uint32_t)SB_NO_KIND;
        if (active && bk != (uint32_t)SB_NO_KIND) {
          const uint32_t yx0 = spos[a];
          const int y0 = (int)(yx0 >> 16), x0 = (int)(yx0 & 0xFFFFu);
          uint8_t* bp = genv + d.beam_plane * d.plane_stride;
          for (int b = lane; b < d.n_beam; b += LPE) {
            const int yy = y0 + d.beam_dy[facing][b], xx = x0 + d.beam_dx[facing][b];
            if ((unsigned)yy < (unsigned)d.H && (unsigned)xx < (unsigned)d.W) {  // world.valid_location
              const uint32_t k = bp[yy * d.W + xx];
              if (!(k < (uint32_t)SB_MAX_KINDS && ((d.beam_block_mask >> k) & 1u))) bp[yy * d.W + xx] = (uint8_t)bk;
            }
          }
        }
        __syncwarp();
        if (lane == 0) sfacing[a] = facing;
      }
      if (lane == 0 && active) {
        // world.total_reward += reward (agent.py:172; cleanup adds it twice: cleanup/agents.py:172)
        if (EXT && d.use_value64) {  // values float32 cannot hold (iowa): accumulate the double the reference adds
          double r64 = 0.0;
          move_and_record(a, pre_tk, &r64);
          total += r64 * (double)d.reward_count;
        } else {
          const float reward = move_and_record(a, pre_tk);
          total += (double)(EXT ? reward * (float)d.reward_count : reward);
        }
      }
      __syncwarp();
    };

    if (fused && LPE == 32 && NP1 && p.par) {
      // Warp-per-instance worlds (large maps, many agents): SB_PAR consecutive agents whose positions are
      // pairwise more than par_radius = max(v + m, 2m) apart (Chebyshev; m = longest action delta, 1 for
      // MovingAgent) cannot see or touch each other's moves this turn -- an agent's move changes cells within m of
      // its position, a window reaches v, two targets are 2m apart at most -- so their observations and moves
      // commute with the list order of environment.py:92 and are done AT ONCE, 8 lanes per agent.  Groups that
      // fail the test (or the tail) take the sequential path.  Rewards are added in list order.
      constexpr int SL = 32 / SB_PAR;
      for (int a0 = 0; a0 < d.A; a0 += SB_PAR) {
        bool par = a0 + SB_PAR <= d.A;
        if (par) {
          bool conflict = false;
          if (lane < SB_PAR * (SB_PAR - 1) / 2) {
            int i = 0, j = lane + 1;  // lane -> pair (i, j), i < j, row-major over the upper triangle
            while (j >= SB_PAR) {
              ++i;
              j = j - SB_PAR + i + 1;
            }
            const uint32_t pi = spos[a0 + i], pj = spos[a0 + j];
            const int dy = (int)(pi >> 16) - (int)(pj >> 16), dx = (int)(pi & 0xFFFFu) - (int)(pj & 0xFFFFu);
            conflict = max(abs(dy), abs(dx)) <= d.par_radius;
          }
          par = !__any_sync(0xFFFFFFFFu, conflict);
        }
        if (par) {
          const int sub = lane / SL, sl = lane % SL, a = a0 + sub;
          const uint32_t yx = spos[a];
          const int y = (int)(yx >> 16), x = (int)(yx & 0xFFFFu);
          int pre_tk = -1;
          if (sl == 0 && active) {  // start the load of the cell this agent will step on
            const int tc = sb_target_cell(d, yx, sact[a]);
            if (tc >= 0) pre_tk = genv[d.agent_plane * d.plane_stride + tc];
          }
          uint8_t* srow = stage_raw + sub * p.stage_row;
          float* dst0 = (p.has_buf && active) ? ring_row(a, slot) : nullptr;
          emit_onehot_row<SL, NP1, 4>(d, genv, ktab, srow, y, x, dst0, sl);
          if (p.a.d_obs_out) {
            float* dst1 = active ? p.a.d_obs_out + (((size_t)t * d.N + env) * d.A + a) * d.obs_size : nullptr;
            __syncwarp();
            emit_onehot_row<SL, NP1, 4>(d, genv, ktab, srow, y, x, dst1, sl);
          }
          // every window has been read (the barrier inside emit_onehot_row follows the scatter): move
          float reward = 0.0f;
          if (sl == 0 && active) reward = move_and_record(a, pre_tk);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < SB_PAR; ++k) {
            const float rk_ = __shfl_sync(0xFFFFFFFFu, reward, k * SL);
            if (lane == 0) total += (double)rk_;  // agent.py:172, in list order
          }
        } else {
          for (int a = a0; a < min(a0 + SB_PAR, d.A); ++a) {
            observe(a);
            act(a, -1);
          }
        }
      }
      record_turn(0, d.A);
    } else if (fused) {
      for (int a = 0; a < d.A; ++a) {
        // planes in global memory: start the load of the cell agent a will step on NOW (its position, its
        // action and every earlier agent's move are final), so the L2 round trip overlaps the observation
        int pre_tk = -1;
        if (!STAGED && lane == 0 && active && !(EXT && d.ext_act)) {
          const int tc = sb_target_cell(d, spos[a], sact[a]);
          if (tc >= 0) pre_tk = genv[d.agent_plane * d.plane_stride + tc];
        }
        observe(a);
        act(a, pre_tk);
      }
      record_turn(0, d.A);
    } else {
      if (p.a.act_agent >= 0) {
        act(p.a.act_agent, -1);
        record_turn(p.a.act_agent, p.a.act_agent + 1);
      }
      if (p.a.obs_agent >= 0) {
        observe(p.a.obs_agent);
      } else if (p.a.obs_agent == SB_OBS_ALL) {
        for (int a = 0; a < d.A; ++a) {
          observe(a);
          __syncwarp();
        }
      }
      __syncwarp();
    }
  }

  // ---- write the instance back ---------------------------------------------------------------
  if (active && p.write_back) {
    if (STAGED) {
      uint4* dst = reinterpret_cast<uint4*>(p.grid + (size_t)env * d.env_stride);
      for (int c = lane; c < d.n_chunks; c += LPE) dst[c] = genv4[c];
    }
    for (int a = lane; a < d.A; a += LPE) {
      uint32_t yx = spos[a];
      p.agent_pos[(size_t)env * d.A + a] = (int32_t)((yx >> 16) * d.W + (yx & 0xFFFFu));
    }
    if (EXT && p.a.d_agent_facing)
      for (int a = lane; a < d.A; a += LPE) p.a.d_agent_facing[(size_t)env * d.A + a] = sfacing[a];
    if (EXT && d.agent_kind_alt >= 0)  // the flags live in the grid while stepping; persist them for the next reset
      for (int a = lane; a < d.A; a += LPE) {
        const uint32_t yx = spos[a];
        d.agent_flag[(size_t)env * d.A + a] =
            genv[d.agent_plane * d.plane_stride + (int)(yx >> 16) * d.W + (int)(yx & 0xFFFFu)] == (uint8_t)d.agent_kind_alt;
      }
    if (lane == 0) p.total_reward[env] = total;
  }
  if (active && err && p.a.d_error) atomicOr(p.a.d_error, err);
}

#ifndef SB_ONLY_LPE
// ----------------------------------------------------------------------------------------------
// Block-per-instance variant for LARGE worlds with MANY agents (BASELINE config 4: 128x128, 64 agents).
//
// One warp per instance (above) leaves the GPU underfilled at the sizes such worlds are run at (2048 instances = 14
// warps per SM) and every window cell is an L2 round trip because a 16 KB plane per warp does not fit in shared
// memory.  Here ONE BLOCK of 4 warps owns the instance: the plane is staged in shared memory for the whole launch (16
// KB + 16 staging rows: six blocks per SM), the sweep runs on all 128 lanes, and up to SIXTEEN agents are observed and
// moved at once -- 4 warps x 4 subgroups of 8 lanes -- without giving up the list-order semantics of
// environment.py:92:
//
//   Two agents whose start-of-turn positions are more than par_radius apart (Chebyshev) cannot read or write a cell
//   the other's move touches this turn, so their sub-steps commute.  Agent k therefore only has to wait for the
//   CONFLICTING agents before it in the list: level(k) = 1 + max{level(j) : j < k, |p_j - p_k| <= par_radius}.  Agents
//   of one level are pairwise independent and run concurrently; levels run in order with a block barrier in between.
//   A 64-bit mask per agent (its conflicting predecessors) is built once per turn by 64 lanes; "ready" is then
//   (mask & ~done) == 0, evaluated identically by every warp (no exchange needed).  Typical turn: ~4 levels.
//
// Only the treasurehunt-type fast path is compiled (one stored plane, one-hot rows, byte-SIMD Philox sweep, fused
// mode, ring attached); everything else keeps using the warp-per-instance kernel.
// ----------------------------------------------------------------------------------------------
#ifndef SB_BLK_SPW
#define SB_BLK_SPW 4  // agent slots per warp (of 8 lanes each): 4 = all lanes; 3 = 12 slots, smaller blocks (7 / SM)
#endif
#ifndef SB_BLK_MINB
#define SB_BLK_MINB 6
#endif
#ifndef SB_BLK_ORDER_TABLE
#define SB_BLK_ORDER_TABLE 1
#endif
constexpr int BLK_THREADS = 128, BLK_SPW = SB_BLK_SPW, BLK_SLOTS = 4 * BLK_SPW, BLK_SL = 8;

struct BlockLayout {
  int q_off, plane_off, agent_off, mask_off, order_off, stage_off, stage_row, act_turns, total;
};

__host__ __device__ inline BlockLayout block_layout(const SbDev& d) {
  BlockLayout L;
  L.q_off = 0;
  L.plane_off = round_up((BLK_THREADS / 32) * (SB_QCAP + 1) * 4, 16) + 16;  // 16 bytes of slack before the plane
  L.act_turns = act_turns_ahead(BLK_THREADS, d.A);
  L.agent_off = L.plane_off + round_up(d.env_stride, 16) + 16;                // ... and after it (row-wise gathers)
  L.mask_off = L.agent_off + round_up(4 * (2 * d.A + L.act_turns * d.A), 16);  // yx, reward, action window
  L.order_off = L.mask_off + round_up(8 * d.A, 16);  // per warp: the ready agents of the level in list order, u8[64]
  L.stage_off = L.order_off + (BLK_THREADS / 32) * 64;
  L.stage_row = BLK_SPW == 4 ? pad64mod128((d.obs_size + 4) * 2) : round_up((d.obs_size + 4) * 2, 32);
  L.total = L.stage_off + BLK_SLOTS * L.stage_row;
  return L;
}

__device__ __forceinline__ int nth_set_bit64(uint64_t m, int k) {  // position of the k-th (0-based) set bit
  const uint32_t lo = (uint32_t)m;
  const int c = __popc(lo);
  return k < c ? (int)__fns(lo, 0u, k + 1) : 32 + (int)__fns((uint32_t)(m >> 32), 0u, k - c + 1);
}

__global__ void __launch_bounds__(BLK_THREADS, SB_BLK_MINB)
    step_kernel_block(const __grid_constant__ SbDev d, const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const BlockLayout L = block_layout(d);
  const int tid = threadIdx.x, warp = tid >> 5, wlane = tid & 31;
  const bool has_slot = wlane / BLK_SL < BLK_SPW;  // (with 3 slots per warp the last 8 lanes only synchronise)
  const int slot_id = warp * BLK_SPW + (has_slot ? wlane / BLK_SL : 0), sl = wlane % BLK_SL;
  const long long env = blockIdx.x;
  uint32_t* q = reinterpret_cast<uint32_t*>(smem_raw + L.q_off) + warp * (SB_QCAP + 1);
  uint8_t* plane = smem_raw + L.plane_off;
  uint4* plane4 = reinterpret_cast<uint4*>(plane);
  uint32_t* spos = reinterpret_cast<uint32_t*>(smem_raw + L.agent_off);
  float* srew = reinterpret_cast<float*>(spos + d.A);
  int32_t* sact_win = reinterpret_cast<int32_t*>(srew + d.A);
  unsigned long long* smask = reinterpret_cast<unsigned long long*>(smem_raw + L.mask_off);
  uint8_t* srow = smem_raw + L.stage_off + slot_id * L.stage_row;
  uint8_t* sorder = smem_raw + L.order_off + warp * 64;
  const uint32_t envid = (uint32_t)(d.env_id_base + (uint64_t)env);
  const uint32_t* __restrict__ rk = p.rk;

  // ---- stage the instance ------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.grid + (size_t)env * d.env_stride);
    for (int c = tid; c < d.n_chunks; c += BLK_THREADS) plane4[c] = __ldg(src + c);
    for (int a = tid; a < d.A; a += BLK_THREADS) {
      const int pos = p.agent_pos[(size_t)env * d.A + a];
      const int y = pos / d.W;
      spos[a] = ((uint32_t)y << 16) | (uint32_t)(pos - y * d.W);
    }
    for (int e = tid; e < BLK_SLOTS * L.stage_row / 4; e += BLK_THREADS)
      reinterpret_cast<uint32_t*>(smem_raw + L.stage_off)[e] = 0u;
  }
  double total = tid == 0 ? p.total_reward[env] : 0.0;
  uint32_t err = 0;
  __syncthreads();

  const int T = p.a.num_turns;
  const size_t rows = (size_t)d.N * d.A, row0 = (size_t)env * d.A;
  float* const ring0 = p.buf.d_states + row0 * (size_t)p.buf.states_row_stride;
  const uint32_t tick0 = p.a.d_clock ? p.a.d_clock[0] : p.a.tick0;
  int slot = p.a.d_clock ? (int)p.a.d_clock[1] : p.a.slot0;
  const uint64_t all = d.A >= 64 ? ~0ull : ((1ull << d.A) - 1ull);
  const uint32_t spawn4 = (uint32_t)d.single_spawn_kind * 0x01010101u;
  int tw = 0;

  for (int t = 0; t < T; ++t, slot = slot + 1 < p.buf.capacity ? slot + 1 : 0) {
    const uint32_t tick = tick0 + (uint32_t)t;

    // ---- actions of the next act_turns turns (one (turn, agent) pair per lane) -------------------
    if (tw == L.act_turns) tw = 0;
    int32_t* const sact = sact_win + tw * d.A;
    if (tw == 0) {
      for (int e = tid; e < L.act_turns * d.A; e += BLK_THREADS) {
        const int dt = L.act_turns > 1 ? (int)(((uint32_t)e * p.inv_A) >> 16) : 0;
        const int a = e - dt * d.A;
        int act = 0;
        if (t + dt < T) {
          if (p.a.policy == SB_POLICY_UNIFORM) {
            const SbU4 w = sb_philox_rk(envid, tick + (uint32_t)dt, (uint32_t)a, SB_DOM_ACTION, rk);
            act = (int)sb_bounded(w.x, w.y, (uint32_t)d.n_actions);
          } else {
            {
              const size_t at = ((size_t)(t + dt) * d.N + env) * d.A + a;
              act = p.a.actions_u8 ? (int)reinterpret_cast<const uint8_t*>(p.a.d_actions)[at] : p.a.d_actions[at];
            }
          }
        }
        sact_win[e] = act;
      }
    }
    ++tw;

    // ---- entity sweep: all 128 lanes, per-warp candidate queues (see step_kernel) ----------------
    {
      uint32_t qcount = 0;
      for (int base = 0; base < d.n_chunks; base += 2 * BLK_THREADS) {
        const int c0 = base + tid, c1 = c0 + BLK_THREADS;
        uint32_t m0 = 0, m1 = 0;
        if (c0 < d.n_chunks) {
          const uint4 k = plane4[c0];
          const SbU4 r = sb_philox_rk(envid, tick, (uint32_t)c0, SB_DOM_SWEEP, rk);
          m0 = sb_simd_cand_chunk(k.x, k.y, k.z, k.w, r.x, r.y, r.z, r.w, spawn4, d.hmask4);
          if (c1 < d.n_chunks) {
            const uint4 k1 = plane4[c1];
            const SbU4 r1 = sb_philox_rk(envid, tick, (uint32_t)c1, SB_DOM_SWEEP, rk);
            m1 = sb_simd_cand_chunk(k1.x, k1.y, k1.z, k1.w, r1.x, r1.y, r1.z, r1.w, spawn4, d.hmask4);
          }
        }
        uint32_t b = __ballot_sync(0xFFFFFFFFu, (m0 | m1) != 0);
        while (b) {
          if (qcount + 32u > (uint32_t)SB_QCAP) {
            __syncwarp();
            drain_queue_cold(&d, q, qcount, plane, 0, env, tick, rk);
            __syncwarp();
            qcount = 0;
          }
          if (m0 | m1) {
            const bool first = m0 != 0;
            const uint32_t m = first ? m0 : m1;
            const int pos = __ffs(m) - 1;
            if (first) m0 &= m0 - 1;
            else m1 &= m1 - 1;
            q[qcount + __popc(b & ((1u << wlane) - 1u))] = (uint32_t)((first ? c0 : c1) * 16 + 4 * (pos & 7) + (pos >> 3));
          }
          qcount += __popc(b);
          b = __ballot_sync(0xFFFFFFFFu, (m0 | m1) != 0);
        }
      }
      __syncwarp();
      drain_queue(d, q, qcount, plane, 0, env, tick, rk);  // entries: group 0, instance `env`
    }

    // ---- conflicting predecessors of every agent (start-of-turn positions) ------------------------
    // lane -> (agent a, half h): bits [32h, 32h + 32) of a's mask, i.e. predecessors j in that range -- all four warps
    // work, 32 tests per lane, and the two halves are separate words (no reduction)
    for (int e = tid; e < 2 * d.A; e += BLK_THREADS) {
      const int a = e >> 1, h = e & 1;
      const uint32_t pa = spos[a];
      const int ya = (int)(pa >> 16), xa = (int)(pa & 0xFFFFu);
      uint32_t m = 0;
      const int j_hi = min(a, 32 * h + 32);
      for (int j = 32 * h; j < j_hi; ++j) {
        const uint32_t pj = spos[j];
        const int dy = ya - (int)(pj >> 16), dx = xa - (int)(pj & 0xFFFFu);
        if (max(abs(dy), abs(dx)) <= d.par_radius) m |= 1u << (j & 31);
      }
      reinterpret_cast<uint32_t*>(smask)[e] = m;  // little-endian halves of the 64-bit mask
    }
    __syncthreads();  // sweep writes, masks and the action window are visible to every warp

    // ---- agents, level by level ---------------------------------------------------------------
    uint64_t done = 0;
    while (done != all) {  // block-uniform: every warp derives the same ready set
      const int a_lo = wlane, a_hi = wlane + 32;
      const bool rd_lo = a_lo < d.A && !((done >> a_lo) & 1ull) && (smask[a_lo] & ~done) == 0ull;
      const bool rd_hi = a_hi < d.A && !((done >> a_hi) & 1ull) && (smask[a_hi] & ~done) == 0ull;
      const uint64_t ready = (uint64_t)__ballot_sync(0xFFFFFFFFu, rd_lo) |
                             ((uint64_t)__ballot_sync(0xFFFFFFFFu, rd_hi) << 32);
      const int n = __popcll(ready);
#if SB_BLK_ORDER_TABLE
      // the k-th ready agent: every warp writes its own copy of the rank -> agent table (a lane knows the rank of its
      // two agents from a popcount; finding the k-th set bit of a 64-bit mask is a ~40-instruction routine)
      if (rd_lo) sorder[__popcll(ready & ((1ull << a_lo) - 1ull))] = (uint8_t)a_lo;
      if (rd_hi) sorder[__popcll(ready & ((1ull << a_hi) - 1ull))] = (uint8_t)a_hi;
      __syncwarp();
#endif
      for (int r0 = 0; r0 < n; r0 += BLK_SLOTS) {  // uniform trip count; a slot without an agent only synchronises
        const int k = r0 + slot_id;
#if SB_BLK_ORDER_TABLE
        const int a = (has_slot && k < n) ? (int)sorder[k] : -1;
#else
        const int a = (has_slot && k < n) ? nth_set_bit64(ready, k) : -1;
#endif
        uint32_t yx = a >= 0 ? spos[a] : 0u;
        float* dst = a >= 0 ? ring0 + (size_t)a * (size_t)p.buf.states_row_stride +
                                  (size_t)slot * (size_t)p.buf.states_slot_stride
                            : nullptr;
        emit_onehot_row<BLK_SL, true, 1, SB_ROWSEG ? 2 : 1>(d, plane, nullptr, srow, (int)(yx >> 16), (int)(yx & 0xFFFFu),
                                                            dst, sl);
        // every lane of the subgroup has read the window (barrier inside emit_onehot_row): move
        if (a >= 0 && sl == 0) {
          srew[a] = sb_agent_act(d, plane, &yx, sact[a], true, err);
          spos[a] = yx;
        }
        __syncwarp();
      }
      __syncthreads();  // the level's moves are visible before the next level observes
      done |= ready;
    }

    // ---- Buffer.add scalars (one agent per lane) and world.total_reward in list order -------------
    for (int a = tid; a < d.A; a += BLK_THREADS) {
      const uint32_t yx = spos[a];
      const size_t at = (size_t)slot * rows + row0 + a;
      p.buf.d_actions[at] = (int64_t)sact[a];
      p.buf.d_rewards[at] = srew[a];
      p.buf.d_dones[at] = 0.0f;
      if (p.buf.d_positions)
        *reinterpret_cast<longlong2*>(p.buf.d_positions + 2 * at) =
            make_longlong2((long long)(yx >> 16), (long long)(yx & 0xFFFFu));
      const size_t o = ((size_t)t * d.N + env) * d.A + a;
      if (p.a.d_reward_out) p.a.d_reward_out[o] = srew[a];
      if (p.a.d_action_out) p.a.d_action_out[o] = sact[a];
    }
    if (tid == 0)
      for (int a = 0; a < d.A; ++a) total += (double)srew[a];  // agent.py:172, list order
    __syncthreads();  // the action window / rewards may be overwritten by the next turn
  }

  // ---- write the instance back ---------------------------------------------------------------
  {
    uint4* dst = reinterpret_cast<uint4*>(p.grid + (size_t)env * d.env_stride);
    for (int c = tid; c < d.n_chunks; c += BLK_THREADS) dst[c] = plane4[c];
    for (int a = tid; a < d.A; a += BLK_THREADS) {
      const uint32_t yx = spos[a];
      p.agent_pos[(size_t)env * d.A + a] = (int32_t)((yx >> 16) * d.W + (yx & 0xFFFFu));
    }
    if (tid == 0) p.total_reward[env] = total;
  }
  if (err && p.a.d_error) atomicOr(p.a.d_error, err);
}

// The block-per-instance kernel applies to: fused mode, ring attached and no second observation output, one stored
// plane with the one-hot fast path and the row-wise gather table, the byte-SIMD Philox sweep, at most 64 agents
// (64-bit predecessor masks) but enough of them to occupy the 16 slots, a plane too large for the per-warp tilings
// yet small enough for several blocks per SM.
static bool block_kernel_applies(const SbDev& d, const StepParams& p, int kr) {
  if (p.a.mode != SB_MODE_FUSED || !p.has_buf || p.a.d_obs_out) return false;
  if (kr != SB_K_PHILOX_SIMD || !(d.fast_obs && d.n_planes == 1) || !d.rowwise_ok || d.toroidal) return false;
  if (d.A < 16 || d.A > 64) return false;
  if (p.a.tune_lpe != 0 && p.a.tune_lpe != 128) return false;
  if (p.a.tune_staged < 0 || p.a.tune_par < 0) return false;
  if (p.a.tune_lpe == 0 && d.env_stride <= 4096) return false;  // small maps: the per-warp tilings stage them already
  return block_layout(d).total <= 227 * 1024 / 2;
}

static int launch_block(const SbDev& d, StepParams p, cudaStream_t s, SbStepPlan* plan) {
  const BlockLayout L = block_layout(d);
  p.rk[0] = (uint32_t)p.a.seed, p.rk[1] = (uint32_t)(p.a.seed >> 32);  // the stream words of sb_philox_rk
  p.inv_A = (65536u + (uint32_t)d.A - 1u) / (uint32_t)d.A;
  p.staged = 1, p.par = 1;
  if (plan) {
    plan->lpe = BLK_THREADS, plan->kernel_rng = SB_K_PHILOX_SIMD, plan->fast_onehot = 1, plan->staged = 1, plan->par = 1;
    plan->threads = BLK_THREADS, plan->smem_bytes = L.total, plan->blocks = d.N;
    return SB_OK;
  }
  if (L.total > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(step_kernel_block, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return note_cuda_error(e);
  }
  step_kernel_block<<<(unsigned)d.N, BLK_THREADS, L.total, s>>>(d, p);
  count_launch();
  return check_launch();
}
#endif  // SB_ONLY_LPE

#ifndef SB_ONLY_LPE
// ----------------------------------------------------------------------------------------------
// Split-phase launch that ONLY moves one agent (`act_agent` set, nothing to observe afterwards, no sweep): the last
// launch of every live-policy turn (Agent.act of the last agent slot, agents/agent.py:166-173).  MovingAgent.act
// touches two cells, one position and the ring scalars, so one THREAD per instance works on the planes in place in
// global memory -- ~5 sectors per instance instead of staging the whole instance into shared memory and writing it
// back (896 bytes at 21x21).  Worlds with facings / beams / all-layer rewards / agent states keep the general kernel.
// ----------------------------------------------------------------------------------------------
#ifndef SB_ACT_ONLY
#define SB_ACT_ONLY 1
#endif
constexpr int ACT_THREADS = 256;
__global__ void __launch_bounds__(ACT_THREADS)
    act_only_kernel(const __grid_constant__ SbDev d, const __grid_constant__ StepParams p) {
  const long long env = (long long)blockIdx.x * ACT_THREADS + threadIdx.x;
  if (env >= d.N) return;
  const int a = p.a.act_agent;
  const size_t ia = (size_t)env * d.A + a;
  uint8_t* genv = p.grid + (size_t)env * d.env_stride;
  const uint32_t tick = p.a.d_clock ? p.a.d_clock[0] : p.a.tick0;
  int action = 0;
  if (p.a.policy == SB_POLICY_UNIFORM) {
    const SbU4 w = sb_philox_rk((uint32_t)(d.env_id_base + (uint64_t)env), tick, (uint32_t)a, SB_DOM_ACTION, p.rk);
    action = (int)sb_bounded(w.x, w.y, (uint32_t)d.n_actions);
  } else if (p.a.d_actions) {
    action = p.a.actions_u8 ? (int)reinterpret_cast<const uint8_t*>(p.a.d_actions)[ia] : p.a.d_actions[ia];
  }
  const int pos = p.agent_pos[ia];
  const int y = pos / d.W;
  uint32_t yx = ((uint32_t)y << 16) | (uint32_t)(pos - y * d.W);
  uint32_t err = 0;
  const float reward = sb_agent_act(d, genv, &yx, action, true, err);
  p.agent_pos[ia] = (int32_t)((yx >> 16) * d.W + (yx & 0xFFFFu));
  p.total_reward[env] += (double)reward;  // world.total_reward += reward (agent.py:172)
  if (p.has_buf) {  // the scalar part of Buffer.add (buffers.py:57-62)
    const int slot = p.a.d_clock ? (int)p.a.d_clock[1] : p.a.slot0;
    const size_t at = (size_t)slot * ((size_t)d.N * d.A) + ia;
    p.buf.d_actions[at] = (int64_t)action;
    p.buf.d_rewards[at] = reward;
    p.buf.d_dones[at] = 0.0f;
    if (p.buf.d_positions)
      *reinterpret_cast<longlong2*>(p.buf.d_positions + 2 * at) =
          make_longlong2((long long)(yx >> 16), (long long)(yx & 0xFFFFu));
  }
  if (p.a.d_reward_out) p.a.d_reward_out[ia] = reward;
  if (p.a.d_action_out) p.a.d_action_out[ia] = action;
  if (err && p.a.d_error) atomicOr(p.a.d_error, err);
}

static bool act_only_applies(const SbDev& d, const StepParams& p) {
  if (!SB_ACT_ONLY || p.a.mode != SB_MODE_SPLIT || p.a.do_sweep || p.a.act_agent < 0 || p.a.obs_agent != -1) return false;
  if (d.ext_act || d.agent_kind_alt >= 0 || d.obs_flag) return false;
  return p.a.tune_lpe == 0 && p.a.tune_staged == 0;
}

static int launch_act_only(const SbDev& d, StepParams p, cudaStream_t s, SbStepPlan* plan) {
  p.rk[0] = (uint32_t)p.a.seed, p.rk[1] = (uint32_t)(p.a.seed >> 32);  // the stream words of sb_philox_rk
  const long long blocks = ((long long)d.N + ACT_THREADS - 1) / ACT_THREADS;
  if (plan) {
    plan->lpe = 1, plan->kernel_rng = SB_K_TAPE, plan->fast_onehot = 0, plan->staged = 0, plan->par = 0;
    plan->threads = ACT_THREADS, plan->smem_bytes = 0, plan->blocks = blocks;
    return SB_OK;
  }
  act_only_kernel<<<(unsigned)blocks, ACT_THREADS, 0, s>>>(d, p);
  count_launch();
  return check_launch();
}
#endif  // SB_ONLY_LPE

// ----------------------------------------------------------------------------------------------
// launch
// ----------------------------------------------------------------------------------------------
template <int LPE, int RNG, bool NP1>
static int launch_one(const SbDev& d, StepParams p, cudaStream_t s, SbStepPlan* plan) {
  const bool staged = stage_planes(d, p.a.tune_staged);
  const SmemLayout L = smem_layout(d, SB_THREADS, LPE, NP1, staged, RNG == SB_K_SPARSE);
  p.rk[0] = (uint32_t)p.a.seed, p.rk[1] = (uint32_t)(p.a.seed >> 32);  // the stream words of sb_philox_rk
  p.q_off = L.q_off, p.planes_off = L.planes_off, p.pos_off = L.pos_off, p.stage_off = L.stage_off;
  p.plane_env = L.plane_env, p.stage_env = L.stage_env, p.hits_off = L.hits_off;
  p.stage_row = L.stage_row;
  p.agent_words = L.agent_words;
  p.act_turns = p.a.mode == SB_MODE_FUSED ? L.act_turns : 1;
  p.inv_A = (65536u + (uint32_t)d.A - 1u) / (uint32_t)d.A;  // e / A == (e * inv_A) >> 16 for e < 32, A <= 64
  p.par = (NP1 && LPE == 32 && d.A >= SB_PAR && !d.ext_act && !d.obs_extra && !d.toroidal && p.a.tune_par >= 0) ? 1 : 0;
  p.staged = staged ? 1 : 0;
  const int smem = L.total;
  if (smem > 227 * 1024) return SB_EUNSUPPORTED;
  if (!staged && LPE != 32) return SB_EUNSUPPORTED;  // planes used in place: one warp per instance
  const int epb = SB_THREADS / LPE;
  const long long blocks = ((long long)d.N + epb - 1) / epb;
  if (plan) {
    plan->lpe = LPE, plan->kernel_rng = RNG, plan->fast_onehot = NP1 ? 1 : 0, plan->staged = p.staged, plan->par = p.par;
    plan->threads = SB_THREADS, plan->smem_bytes = smem, plan->blocks = blocks;
    return SB_OK;
  }
  auto go = [&](auto kern) -> int {
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return note_cuda_error(e);
    }
    // L2 prefetch distance of the staging copies: the blocks resident at once (looked up once per kernel / size)
    static thread_local struct { const void* kern; int smem, dev, resident; } cache = {nullptr, 0, -1, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (cache.kern != (const void*)kern || cache.smem != smem || cache.dev != dev) {
      int per_sm = 0, sms = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SB_THREADS, smem) != cudaSuccess ||
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        per_sm = 0, (void)cudaGetLastError();
      cache = {(const void*)kern, smem, dev, per_sm * sms};
    }
    p.prefetch_blocks = (staged && blocks > cache.resident) ? cache.resident : 0;
    kern<<<(unsigned)blocks, SB_THREADS, smem, s>>>(d, p);
    count_launch();
    return check_launch();
  };
  if (staged) return go(step_kernel<LPE, RNG, NP1, true>);
  if constexpr (LPE == 32) return go(step_kernel<LPE, RNG, NP1, false>);
  return SB_EUNSUPPORTED;
}

template <int LPE>
static int launch_rng(const SbDev& d, const StepParams& p, int kr, cudaStream_t s, SbStepPlan* plan) {
  const bool np1 = d.fast_obs && d.n_planes == 1;
  switch (kr) {
    case SB_K_TAPE:  // replay: speed is irrelevant, but it shares the observation / agent code under test
      return np1 ? launch_one<LPE, SB_K_TAPE, true>(d, p, s, plan) : launch_one<LPE, SB_K_TAPE, false>(d, p, s, plan);
    case SB_K_PHILOX_SIMD:
      return np1 ? launch_one<LPE, SB_K_PHILOX_SIMD, true>(d, p, s, plan)
                 : launch_one<LPE, SB_K_PHILOX_SIMD, false>(d, p, s, plan);
    case SB_K_SPARSE:
      return np1 ? launch_one<LPE, SB_K_SPARSE, true>(d, p, s, plan) : launch_one<LPE, SB_K_SPARSE, false>(d, p, s, plan);
    default: return launch_one<LPE, SB_K_PHILOX_GENERIC, false>(d, p, s, plan);
  }
}

#define SB_CAT2(a, b) a##b
#define SB_CAT(a, b) SB_CAT2(a, b)
#ifdef SB_ONLY_LPE
// translation unit of ONE group width: every step_kernel<SB_ONLY_LPE, ...> instantiation (make -j compiles the
// widths and the rest of the library in parallel)
int SB_CAT(launch_rng_, SB_ONLY_LPE)(const SbDev& d, const StepParams& p, int kr, cudaStream_t s, SbStepPlan* plan) {
  return launch_rng<SB_ONLY_LPE>(d, p, kr, s, plan);
}
#else
int launch_rng_2(const SbDev& d, const StepParams& p, int kr, cudaStream_t s, SbStepPlan* plan);
int launch_rng_4(const SbDev& d, const StepParams& p, int kr, cudaStream_t s, SbStepPlan* plan);
int launch_rng_8(const SbDev& d, const StepParams& p, int kr, cudaStream_t s, SbStepPlan* plan);
int launch_rng_16(const SbDev& d, const StepParams& p, int kr, cudaStream_t s, SbStepPlan* plan);
int launch_rng_32(const SbDev& d, const StepParams& p, int kr, cudaStream_t s, SbStepPlan* plan);

// Lanes per instance.  `hint` = SbStepArgs.tune_lpe (0: choose).
static int pick_lpe(const SbDev& d, int hint, bool staged, bool sparse) {
  if (!staged) return 32;  // planes used in place in global memory: one warp per instance
  if (hint) return hint;
  int lpe = 4;
  while (lpe < 32 && d.n_chunks / lpe > 8) lpe *= 2;
  if (d.A >= 4 && lpe < 8) lpe = 8;  // agent phases dominate: halve their latency (measured, profiles/README.md)
  // small batches: widen the group until ~8 warps per SM are in flight (148 SMs), trading lane efficiency for
  // latency.  Config 2 (4096 instances x 100 turns): LPE 4 / 8 / 16 / 32 = 0.457 / 0.368 / 0.362 / 0.453 ms --
  // beyond ~14 warps per SM the extra instruction streams cost more than the latency they hide.
  while (lpe < 32 && (long long)d.N * lpe < 148LL * 8 * 32) lpe *= 2;
  // shared memory must fit, and should not cap the SM below 6 resident blocks (24 warps): big staging rows
  // (cleanup: 4.4 KB of f32 per instance) are better served by wider groups (measured: 16.9 -> 13.7 ms)
  const bool np1 = d.fast_obs && d.n_planes == 1;
  while (lpe < 32 && smem_layout(d, SB_THREADS, lpe, np1, staged, sparse).total > 227 * 1024 / 6) lpe *= 2;
  return lpe;
}

int launch_step(const SbDev& d, const StepParams& p, cudaStream_t s, SbStepPlan* plan) {
  int kr;
  if (p.a.rng == SB_RNG_TAPE) kr = SB_K_TAPE;
  else if (p.a.rng == SB_RNG_SPARSE) kr = SB_K_SPARSE;
  else kr = d.simd_sweep ? SB_K_PHILOX_SIMD : SB_K_PHILOX_GENERIC;
  const int hint = p.a.tune_lpe;
  if (hint != 0 && hint != 2 && hint != 4 && hint != 8 && hint != 16 && hint != 32 && hint != 128) return SB_EINVAL;
  if (p.a.tune_staged < -1 || p.a.tune_staged > 1 || p.a.tune_par < -1 || p.a.tune_par > 0) return SB_EINVAL;
  if (p.a.tune_staged < 0 && hint != 0 && hint != 32) return SB_EINVAL;
  if (act_only_applies(d, p)) return launch_act_only(d, p, s, plan);
  if (block_kernel_applies(d, p, kr)) return launch_block(d, p, s, plan);
  if (hint == 128) return SB_EUNSUPPORTED;  // the block-per-instance kernel was asked for but does not apply
  switch (pick_lpe(d, hint, stage_planes(d, p.a.tune_staged), kr == SB_K_SPARSE)) {
    case 2: return launch_rng_2(d, p, kr, s, plan);
    case 4: return launch_rng_4(d, p, kr, s, plan);
    case 8: return launch_rng_8(d, p, kr, s, plan);
    case 16: return launch_rng_16(d, p, kr, s, plan);
    default: return launch_rng_32(d, p, kr, s, plan);
  }
}

#endif  // SB_ONLY_LPE

}  // namespace sb
