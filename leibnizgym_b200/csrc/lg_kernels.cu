// sm_100a kernels + C ABI of the TriFinger MDP hot path (see include/leibniz_b200.h).
//
// Two launches per env step, both env-major and HBM-bound (no tensor cores: the path is
// elementwise, ~2 FLOP/B):
//   pre_physics_kernel   ordered mask compaction (decoupled look-back) + reset / goal-reset
//                        sampling and scatter + action store + action->torque; tile slabs by TMA bulk copy
//   post_physics_kernel  obs/states fill + scale_transform, six reward terms, termination,
//                        step counters / timeouts / dones, episode statistics; one lane per output column
// Reference paths are relative to /root/reference/leibnizgym/.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>

#include "lg_device.cuh"

namespace lg {

// Programmatic dependent launch (PDL): a kernel launched with programmatic stream serialisation may
// begin (CTA scheduling, parameter fetch, address set-up) before its predecessor has finished;
// griddepcontrol.wait then blocks until the predecessor's grid has completed and its writes are
// visible.  Two ~2 us launches per step make this worth ~1/4 of the step time at 16k envs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// =========================================================================================
// post-physics: one CTA = one tile of E envs, 4 threads per env
// =========================================================================================
constexpr int kPostThreads = 256;

template <int A, bool ASYM>
struct Layout {
  static constexpr int OBS = 32 + A;              // q 9 | qdot 9 | object pose 7 | goal pose 7 | action A
  static constexpr int STATE = OBS + 72;          // + object vel 6 | fingertips 39 | torque 9 | wrench 18
  static constexpr int ROW = ASYM ? STATE : OBS;  // floats per env in the staged tile
  static constexpr int OFF_OBJ = 18, OFF_GOAL = 25, OFF_ACT = 32;
  static constexpr int OFF_OBJVEL = OBS, OFF_TIPS = OBS + 6, OFF_TORQUE = OBS + 45, OFF_FT = OBS + 54;
};

// coefficient slots computed once per CTA (python-float arithmetic of the reward modules)
enum Coef { C_REACH = 0, C_MOVE, C_DIST, C_ROT_SCALE, C_ROT_SCHED, C_ROT_W, C_DELTA_RAMP, C_DELTA_W,
            C_OBJMOVE, C_DT, C_DT_RCP, C_POS_TOL, C_ROT_TOL, C_BONUS, C_KP_W, C_KP_SCALE, C_KP_EPS, C_NOISE_EPOCH,
            C_COUNT };

__host__ __device__ inline double sched_gate(const LgRewardTerm& t, double T) {  // rewards.py:56-60
  if (t.sched_start != t.sched_end) return (t.sched_start <= T && T <= t.sched_end) ? 1.0 : 0.0;
  return 1.0;
}
__host__ __device__ inline double sched_ramp(const LgRewardTerm& t, double T) {  // rewards.py:14-17, :169-172
  if (t.sched_start != t.sched_end) {
    const double v = (T - t.sched_start) / (t.sched_end - t.sched_start);
    return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
  }
  return 1.0;
}
// The reward modules' Python-float arithmetic for env_steps_count = T.  Same code on host and device
// (IEEE double in both places), so the two clock modes produce identical coefficients.
__host__ __device__ inline void compute_coefs(const LgParams& P, double T, float* c) {
  const LgRewardTerm* t = P.terms;
  c[C_REACH] = (float)(t[0].weight * sched_gate(t[0], T));             // rewards.py:235
  c[C_MOVE] = (float)t[1].weight;                                       // rewards.py:263
  c[C_DIST] = (float)((t[2].weight * P.dt) * sched_gate(t[2], T));     // rewards.py:63
  c[C_ROT_SCALE] = (float)t[3].scale;                                   // rewards.py:137
  c[C_ROT_SCHED] = (float)(sched_gate(t[3], T) * P.dt);
  c[C_ROT_W] = (float)t[3].weight;                                      // rewards.py:139
  c[C_DELTA_RAMP] = (float)sched_ramp(t[4], T);                         // rewards.py:182
  c[C_DELTA_W] = (float)t[4].weight;                                    // rewards.py:184
  c[C_OBJMOVE] = (float)t[5].weight;                                    // rewards.py:91
  c[C_DT] = (float)P.dt;
  c[C_DT_RCP] = 1.0f / (float)P.dt;                                     // IEEE division: correctly rounded
  c[C_POS_TOL] = (float)P.position_tolerance;
  c[C_ROT_TOL] = (float)P.orientation_tolerance;
  c[C_BONUS] = (float)P.success_bonus;
  c[C_KP_W] = (float)(t[6].weight * P.dt);
  c[C_KP_SCALE] = (float)t[6].scale;
  c[C_KP_EPS] = (float)t[6].eps;
  // step identifier of the domain-randomisation noise stream: env_steps_count itself (bit pattern, not a value)
  union { uint32_t u; float f; } bits;
  bits.u = (uint32_t)(unsigned long long)T;
  c[C_NOISE_EPOCH] = bits.f;
}
static_assert(C_COUNT <= LG_NUM_COEF, "LgCoef too small");

// Fast division by a per-column constant: q = x*r, q' = fma(fma(-q, span, x), r, q) with r = fp32(1/span)
// correctly rounded (Markstein).  Contract, verified exhaustively by lg_selftest_division over all 2^32
// numerators for every span the env uses:
//   * 2^-100 <= |x| < 2^100 : bit-identical to IEEE x / span;
//   * x = +-0                : 0 (a negative zero comes out as +0);
//   * 0 < |x| < 2^-100       : within 1 ulp (the residual may be subnormal);
//   * |x| >= 2^100, inf      : flagged through `amax`; the caller redoes the column with __fdiv_rn.
__device__ __forceinline__ float div_by_const(float num, float span, float rcp, float& amax) {
  const float q = num * rcp;
  const float r = __fmaf_rn(-q, span, num);
  amax = fmaxf(amax, fabsf(num));
  return __fmaf_rn(r, rcp, q);
}
constexpr float kDivSafeMax = 1.2676506e30f;  // 2^100

// bfloat16 bits of a float, round-to-nearest-even (what torch's .to(torch.bfloat16) does)
__device__ __forceinline__ uint16_t to_bf16(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }

// Cold path: re-emit one lane's output column with IEEE division (only when a numerator left the fast
// division's window).  Re-reads the source so the hot path keeps its registers.
template <int STATE, int OBS, bool ASYM>
__device__ __noinline__ void output_exact(const LgParams& P, const LgBuffers& B, const float* src, int stride, int cnt,
                                          int64_t env0, int dcol) {
  const float centre = P.scale_centre[dcol], span = P.scale_span[dcol], clip = P.clip_obs;
  for (int k = 0; k < cnt; ++k) {
    const int64_t e = env0 + k;
    const float raw = src[(int64_t)k * stride];
    const float v = P.normalize_obs ? __fdiv_rn(2.0f * (raw - centre), span) : raw;
    const float vc = fminf(fmaxf(v, -clip), clip);
    if (ASYM) B.states[e * STATE + dcol] = v;
    if (dcol < OBS) B.obs[e * OBS + dcol] = v;
    if (ASYM && B.states_clipped) B.states_clipped[e * STATE + dcol] = vc;
    if (dcol < OBS && B.obs_clipped) B.obs_clipped[e * OBS + dcol] = vc;
    if (ASYM && B.states_bf16) B.states_bf16[e * STATE + dcol] = to_bf16(B.states_clipped ? vc : v);
    if (dcol < OBS && B.obs_bf16) B.obs_bf16[e * OBS + dcol] = to_bf16(B.obs_clipped ? vc : v);
  }
}

// One lane = one OUTPUT column of the tile ("role"), walking over the envs of its part of the tile.
// The role fixes, per lane and once per launch: the source pointer and row stride, the output column
// with its scale constants, and where (if anywhere) the raw value is staged for the reward math.
// The per-element code is then identical for every lane — no divergence although the lanes of a warp
// read from seven different tensors — and free of index arithmetic:
//     load   v[k] = src[k * stride]                       (all loads of the tile in flight at once)
//     emit   states[k][dcol] = obs[k][dcol] = scale(v[k])  (compile-time row offsets)
// Role order = output column order of the states row, except that the nine fingertip POSITION columns
// come right after the observation columns: the lanes that feed obs, the reward staging and the next
// history entry are then all in the first two warps of a part, and the other warps skip that code.
template <int A, bool ASYM, int E>
struct Roles {
  static constexpr int OBS = 32 + A;
  static constexpr int R_TIPPOS = OBS;                  // 9 roles
  static constexpr int R_TIPREST = R_TIPPOS + 9;        // 30 roles (asymmetric only)
  static constexpr int R_OBJVEL = R_TIPREST + 30;       // 6
  static constexpr int R_FT = R_OBJVEL + 6;             // 18
  static constexpr int R_TQ = R_FT + 18;                // 9
  static constexpr int R_END = ASYM ? R_TQ + 9 : R_TIPREST;
  static constexpr int LANES = R_END <= 64 ? 64 : 128;  // role lanes per tile part
  static constexpr int PARTS = kPostThreads / LANES;    // the tile's envs are split over the parts
  static constexpr int EP = E / PARTS;                  // envs per lane
  static_assert(E % PARTS == 0 && E <= 32, "tile must split evenly; the reward math uses one lane per env");
  static constexpr int FRONT = R_TIPREST;               // roles < FRONT may feed obs / staging / history
  static_assert(R_END <= LANES, "more output columns than role lanes");
};

template <int A, bool ASYM, bool REWARD, bool CLIP, int E, bool EXT>
__global__ void __launch_bounds__(kPostThreads, 4)
post_physics_kernel(const __grid_constant__ LgParams P, const __grid_constant__ LgSimState S,
                    const __grid_constant__ LgBuffers B, const __grid_constant__ LgCoef CF) {
  using L = Layout<A, ASYM>;
  using R = Roles<A, ASYM, E>;          // E envs per CTA
  constexpr int EP = R::EP;
  constexpr int HS = LG_HISTORY_COLS + 1;  // padded row: lane-per-env reads stay conflict free
  // only what the reward terms read is staged in shared memory (raw, unscaled)
  __shared__ float s_obj[E * 7];        // object pose
  __shared__ float s_goal[E * 7];       // goal pose
  __shared__ float s_tips[E * 9];       // fingertip positions
  __shared__ float s_hist[E * HS];      // previous fingertip positions (9) + previous object pose (7)
  __shared__ float s_coef[C_COUNT];
  __shared__ float s_part[12][E];       // sub-task results of the reward warps
  __shared__ float s_stat[LG_NUM_STATS][E + 1];
  __shared__ float s_noise[EXT ? L::OBS : 1][EXT ? E + 1 : 1];   // extension: standard normals of the obs columns

  const int tid = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * E;
  const int nvalid = (int)min((int64_t)E, P.num_envs - e0);

  // ---- role of this lane ---------------------------------------------------------------------------
  int role = tid % R::LANES;
  if (role >= R::R_END) role -= (R::LANES - R::R_END);  // spare lanes duplicate a column (same value, same address)
  const bool front = (tid % R::LANES) / 32 * 32 < R::FRONT;  // warp-uniform: this warp holds obs/stage/history roles
  const int env_first = (tid / R::LANES) * EP;     // first env (within the tile) of this lane's part
  const float* src;             // source of (env_first, column)
  int stride;                   // source row stride in floats
  int dcol;                     // output column (scaled): states[:, dcol], obs[:, dcol] if dcol < OBS; -1: none
  float* stage = nullptr;       // shared destination of env_first's raw value, or null
  int stage_stride = 0;
  int hist_col = -1;            // column of the NEXT history entry this lane provides, or -1
  {
    const int body_stride = P.bodies_per_env * 13, actor_stride = P.actors_per_env * 13;
    auto tip_src = [&](int tip, int c) {
      const int body = tip == 0 ? P.fingertip_body[0] : tip == 1 ? P.fingertip_body[1] : P.fingertip_body[2];
      return S.rigid_body + body * 13 + c;
    };
    if (role < 18) {                                         // dof_state (pos, vel) interleaved  trifinger_env.py:1003-1007
      src = S.dof_state + role; stride = 18;
      dcol = (role & 1) * 9 + (role >> 1);
    } else if (role < 25) {                                  // object pose (root row of actor 4e+2)  :975, :1011
      const int c = role - 18;
      src = S.root_state + P.object_slot * 13 + c; stride = actor_stride;
      dcol = L::OFF_OBJ + c;
      stage = s_obj + c; stage_stride = 7; hist_col = 9 + c;
    } else if (role < 32) {                                  // goal pose buffer                  :1015
      const int c = role - 25;
      src = B.goal_pose + c; stride = 7;
      dcol = L::OFF_GOAL + c;
      stage = s_goal + c; stage_stride = 7;
    } else if (role < R::OBS) {                              // last action                       :1019
      const int c = role - 32;
      src = B.action + c; stride = A;
      dcol = L::OFF_ACT + c;
    } else if (role < R::R_TIPREST) {                        // fingertip positions (bodies 6/11/16)  :974, :1040
      const int j = role - R::R_TIPPOS, tip = j / 3, c = j - tip * 3;
      src = tip_src(tip, c); stride = body_stride;
      dcol = ASYM ? L::OFF_TIPS + tip * 13 + c : -1;
      stage = s_tips + j; stage_stride = 9; hist_col = j;
    } else if (role < R::R_OBJVEL) {                         // fingertip orientation + velocity
      const int j = role - R::R_TIPREST, tip = j / 10, c = 3 + (j - tip * 10);
      src = tip_src(tip, c); stride = body_stride;
      dcol = L::OFF_TIPS + tip * 13 + c;
    } else if (role < R::R_FT) {                             // object velocity                   :1035
      const int c = role - R::R_OBJVEL;
      src = S.root_state + P.object_slot * 13 + 7 + c; stride = actor_stride;
      dcol = L::OFF_OBJVEL + c;
    } else if (role < R::R_TQ) {                             // fingertip wrenches                :1051
      const int c = role - R::R_FT;
      src = S.ft_sensors + c; stride = 18;
      dcol = L::OFF_FT + c;
    } else {                                                 // dof torque                        :1047
      const int c = role - R::R_TQ;
      src = S.dof_force + c; stride = 9;
      dcol = L::OFF_TORQUE + c;
    }
  }
  src += (e0 + env_first) * stride;
  pdl_wait();  // everything above is independent of the previous kernel's results
  const int cnt = max(0, min(EP, nvalid - env_first));   // envs of this lane: env_first .. env_first + cnt - 1
  const bool full = nvalid == E;                           // every CTA but possibly the last

  // ---- phase 1: every global load of the tile in flight -----------------------------------------
  float v[EP];
  if (full) {
#pragma unroll
    for (int k = 0; k < EP; ++k) v[k] = ld_stream1(src + (int64_t)k * stride);
  } else {
#pragma unroll
    for (int k = 0; k < EP; ++k) v[k] = k < cnt ? ld_stream1(src + (int64_t)k * stride) : 0.0f;
  }
  // reward warps: thread (w, env) = (tid / 32, tid % 32), w < 4.  Each fetches one 16-byte piece of the
  // env's previous history entry (64-byte rows), warp 0 also the env's flags and step counter.
  const int rw = tid >> 5, renv = tid & 31;
  const bool rlive = REWARD && rw < 4 && renv < nvalid;
  float4 hprev = make_float4(0.f, 0.f, 0.f, 0.f);
  uint8_t in_goal_reset = 0, in_succ = 0, in_reset = 0;
  int64_t in_steps = 0;
  if (rlive) {
    hprev = ld_stream4(reinterpret_cast<const float4*>(B.history + (e0 + renv) * LG_HISTORY_COLS) + rw);
    if (rw == 0) {
      const int64_t e = e0 + renv;
      in_goal_reset = B.goal_reset[e]; in_succ = B.successes[e]; in_reset = B.reset[e]; in_steps = B.steps_count[e];
    }
  }
  // this lane's scale_transform constants (torch_utils.py:33-36) in half-span form:
  // 2 (x - c) / span == (x - c) / (span / 2), both scalings exact.  normalize_obs = False: x / 1.
  float centre = 0.0f, half_span = 1.0f, rcp_half = 1.0f;
  if (P.normalize_obs && dcol >= 0) {
    centre = __ldg(B.scale_table + dcol);
    half_span = 0.5f * __ldg(B.scale_table + LG_MAX_STATE_DIM + dcol);
    rcp_half = 2.0f * __ldg(B.scale_table + 2 * LG_MAX_STATE_DIM + dcol);
  }

  pdl_launch_dependents();  // the next kernel may start launching; it still waits for this grid to finish

  // ---- reward coefficients: from the launch arguments, or (device clock) from what lg_pre_physics wrote ----
  if (tid < C_COUNT) s_coef[tid] = (REWARD && P.use_device_clock) ? __ldg(B.reward_coef + tid) : CF.v[tid];

  // ---- phase 2: stage what the reward terms read, then ONE barrier -------------------------------
  // Only the warps that hold staged columns wait for (a small part of) their data here; the others
  // reach the barrier right after issuing their loads.  The reward math then runs on warps 0-3 while
  // the bulk of the tile is still arriving and being scaled and stored by warps 4-7.
  if (rlive) {
    float* h = s_hist + renv * HS + rw * 4;
    h[0] = hprev.x; h[1] = hprev.y; h[2] = hprev.z; h[3] = hprev.w;
  }
  // extension (DR observation noise): the tile's standard normals, generated by ALL threads while the loads are in
  // flight — one Philox block + two Box-Muller pairs per (column, group of 4 envs by GLOBAL index), so the stream
  // does not depend on tiling or sharding.  Columns with sigma = 0 are skipped.
  if (EXT && P.dr_activate) {
    const uint64_t base = (uint64_t)(P.env_offset + e0);
    const uint64_t g0 = base >> 2;
    const int ngroups = (int)(((base + E - 1) >> 2) - g0) + 1;
    const uint32_t noise_epoch = __float_as_uint((REWARD && P.use_device_clock) ? __ldg(B.reward_coef + C_NOISE_EPOCH)
                                                                                  : CF.v[C_NOISE_EPOCH]);
    for (int i = tid; i < L::OBS * ngroups; i += kPostThreads) {
      const int col = i / ngroups, g = i - col * ngroups;
      if (__ldg(B.scale_table + 3 * LG_MAX_STATE_DIM + col) != 0.0f) {
        const uint64_t grp = g0 + g;
        const U4 r = philox4x32_10(U4{(uint32_t)grp, (uint32_t)(grp >> 32) ^ kPurposeNoise, (uint32_t)col, noise_epoch},
                                   (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
        float n[4];
        box_muller(r.x, r.y, n[0], n[1]);
        box_muller(r.z, r.w, n[2], n[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t local = (int64_t)(4 * grp + q) - (int64_t)base;
          if (local >= 0 && local < E) s_noise[col][local] = n[q];
        }
      }
    }
  }
  if (front && stage) {
    float* dst = stage + env_first * stage_stride;
    if (full) {
#pragma unroll
      for (int k = 0; k < EP; ++k) dst[k * stage_stride] = v[k];
    } else {
#pragma unroll
      for (int k = 0; k < EP; ++k)
        if (k < cnt) dst[k * stage_stride] = v[k];
    }
  }
  __syncthreads();

  // ---- phase 3 (all warps; the reward warps come back to it after their math) -------------------------
  auto emit_outputs = [&](auto full_c) {
    constexpr bool FULLC = decltype(full_c)::value;  // full tile: no per-element bound checks at all
    // history shift (deque.appendleft, trifinger_env.py:974-975): current -> entry read next step.
    // After the barrier: every read of the previous entry has completed.
    if (front && hist_col >= 0) {
      float* dst = B.history + (e0 + env_first) * LG_HISTORY_COLS + hist_col;
#pragma unroll
      for (int k = 0; k < EP; ++k)
        if (FULLC || k < cnt) dst[k * LG_HISTORY_COLS] = v[k];
    }
    float amax = 0.0f;  // largest numerator seen: beyond 2^100 (never, for physical data) the column is redone exactly
    const float clip = P.clip_obs;
    const int st_off = env_first * L::STATE + dcol, ob_off = env_first * L::OBS + dcol;
    float* st = ASYM ? B.states + e0 * L::STATE + st_off : nullptr;                          // trifinger_env.py:990-994
    float* stc = (CLIP && ASYM) ? B.states_clipped + e0 * L::STATE + st_off : nullptr;       // vec_task.py:147
    const bool to_obs = front && dcol >= 0 && dcol < L::OBS;
    float* ob = B.obs + e0 * L::OBS + ob_off;                                                // trifinger_env.py:983-987
    float* obc = CLIP ? B.obs_clipped + e0 * L::OBS + ob_off : nullptr;                      // vec_task.py:167
    // extension (no reference code; the TODO at trifinger_env.py:979): additive Gaussian noise on the RAW
    // actor observation, before scale_transform; the critic's states stay clean.  sigma = 0 -> skipped.
    const float sigma = (EXT && P.dr_activate && to_obs) ? __ldg(B.scale_table + 3 * LG_MAX_STATE_DIM + dcol) : 0.0f;
    const bool noisy = EXT && __any_sync(0xffffffffu, sigma != 0.0f);
    // optional bf16 copies for the policy / value networks: of the clipped values when the
    // wrapper's clamp is fused in, else of the scaled values
    uint16_t* stb = (EXT && ASYM && B.states_bf16) ? B.states_bf16 + e0 * L::STATE + st_off : nullptr;
    uint16_t* obb = (EXT && B.obs_bf16) ? B.obs_bf16 + e0 * L::OBS + ob_off : nullptr;
    // outputs go out with streaming stores (st.global.cs): nothing on the device reads them again before the
    // learner does, and evict-first keeps them from displacing the simulator rows in L2 (measured: 84.0 -> 80.3 us
    // at 262 144 envs, 7.25 -> 7.15 us at 16 384)
#pragma unroll
    for (int k = 0; k < EP; ++k) {
      const float sv = div_by_const(v[k] - centre, half_span, rcp_half, amax);
      if (FULLC || k < cnt) {
        if (ASYM) {
          __stcs(st + k * L::STATE, sv);
          const float svc = CLIP ? fminf(fmaxf(sv, -clip), clip) : sv;
          if (CLIP) __stcs(stc + k * L::STATE, svc);
          if (EXT && stb) __stcs(stb + k * L::STATE, to_bf16(svc));
        }
        if (to_obs) {
          float ov = sv;
          if (noisy) {
            const float n0 = s_noise[dcol][env_first + k];   // generated before the barrier, see above
            ov = div_by_const((v[k] + sigma * n0) - centre, half_span, rcp_half, amax);
          }
          __stcs(ob + k * L::OBS, ov);
          const float ovc = CLIP ? fminf(fmaxf(ov, -clip), clip) : ov;
          if (CLIP) __stcs(obc + k * L::OBS, ovc);
          if (EXT && obb) __stcs(obb + k * L::OBS, to_bf16(ovc));
        }
      }
    }
    if (dcol >= 0 && !(amax < kDivSafeMax))  // cold: same addresses, same thread: plain overwrite
      output_exact<L::STATE, L::OBS, ASYM>(P, B, src, stride, cnt, e0 + env_first, dcol);
    // moving goal (__update_goal_movement_post, trifinger_env.py:1279-1284): after the rewards, the goal pose
    // buffer takes the pose the simulator integrated for the goal body.  The goal-role lanes own their elements
    // of goal_pose (read above, overwritten here), so no other lane observes the change within this step.
    if (EXT && REWARD && P.goal_rotation && front && role >= 25 && role < 32) {
      const int c = role - 25, actor_stride = P.actors_per_env * 13;
      const float* g_src = S.root_state + ((e0 + env_first) * P.actors_per_env + P.goal_slot) * 13 + c;
      float* g_dst = B.goal_pose + (e0 + env_first) * 7 + c;
#pragma unroll
      for (int k = 0; k < EP; ++k)
        if (FULLC || k < cnt) g_dst[k * 7] = g_src[(int64_t)k * actor_stride];
    }
  };
  // ======== reward warps: lane = env, warp = sub-task (uniform control flow inside a warp) ================
  if (REWARD && rw < 4) {
    const int env = renv;
    const bool live = renv < nvalid;
    const float* obj = s_obj + env * 7;
    const float* goal = s_goal + env * 7;
    const float* tips = s_tips + env * 9;
    const float* hist = s_hist + env * HS;
    if (live) {
      const float gx = goal[0], gy = goal[1], gz = goal[2];
      if (rw == 0) {
        // finger_reach_object_rate (rewards.py:219-235): sum_i (|tip_i - obj| - |tip_i' - obj'|)
        const float ox = obj[0], oy = obj[1], oz = obj[2];
        const float px = hist[9], py = hist[10], pz = hist[11];
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float cur = norm3(tips[3 * i] - ox, tips[3 * i + 1] - oy, tips[3 * i + 2] - oz);
          const float prev = norm3(hist[3 * i] - px, hist[3 * i + 1] - py, hist[3 * i + 2] - pz);
          acc = acc + (cur - prev);
        }
        s_part[0][env] = s_coef[C_REACH] * acc;
      } else if (rw == 1) {
        // finger_move_penalty (rewards.py:261-263): sum_9 ((tip - tip') / dt)^2
        const float dt = s_coef[C_DT], dt_rcp = s_coef[C_DT_RCP];
        float acc = 0.0f, dmax = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float d = div_by_const(tips[k] - hist[k], dt, dt_rcp, dmax);
          acc = acc + d * d;
        }
        if (!(dmax < kDivSafeMax)) {  // cold
          acc = 0.0f;
          for (int k = 0; k < 9; ++k) {
            const float d = __fdiv_rn(tips[k] - hist[k], dt);
            acc = acc + d * d;
          }
        }
        s_part[1][env] = s_coef[C_MOVE] * acc;
        // object_dist (rewards.py:62-63) and object_move (rewards.py:88-91)
        const float d = norm3(obj[0] - gx, obj[1] - gy, obj[2] - gz);
        const float dprev = norm3(hist[9] - gx, hist[10] - gy, hist[11] - gz);
        s_part[2][env] = lgsk(d, 50.0f) * s_coef[C_DIST];
        s_part[3][env] = s_coef[C_OBJMOVE] * (d - dprev);
        s_part[4][env] = d;
      } else {
        const Quat gq{goal[3], goal[4], goal[5], goal[6]};
        if (rw == 2) {
          // object_rot (rewards.py:134-139): w * (gate*dt) / (scale*|theta| + scale)
          const Quat oq{obj[3], obj[4], obj[5], obj[6]};
          const float theta = quat_diff_rad(oq, gq);
          const float den = s_coef[C_ROT_SCALE] * fabsf(theta) + s_coef[C_ROT_SCALE];
          s_part[5][env] = (__frcp_rn(den) * s_coef[C_ROT_SCHED]) * s_coef[C_ROT_W];
          s_part[6][env] = theta;
        } else {
          // previous-orientation angle for object_rot_delta (rewards.py:179)
          const Quat pq{hist[12], hist[13], hist[14], hist[15]};
          s_part[7][env] = fabsf(quat_diff_rad(pq, gq));
        }
      }
      // extension (no reference code): keypoint pose reward
      //   w dt mean_k lgsk(|kp_k - kp_k^goal|; scale, eps), kp_k = p + R(q) c_k over the 8 cube corners;
      // two corners per reward warp, summed by warp 0
      if (EXT && ((P.term_active_mask >> LG_TERM_KEYPOINT) & 1)) {
        const Quat oq{obj[3], obj[4], obj[5], obj[6]};
        const Quat gq{goal[3], goal[4], goal[5], goal[6]};
        const float h = (float)P.cube_half_size;
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int kk = 2 * rw + q;
          const float cx = (kk & 1) ? h : -h, cy = (kk & 2) ? h : -h, cz = (kk & 4) ? h : -h;
          float ax, ay, az, bx, by, bz;
          quat_rotate(oq, cx, cy, cz, ax, ay, az);
          quat_rotate(gq, cx, cy, cz, bx, by, bz);
          const float dd = norm3((obj[0] + ax) - (gx + bx), (obj[1] + ay) - (gy + by), (obj[2] + az) - (gz + bz));
          acc = acc + lgsk(dd, s_coef[C_KP_SCALE], s_coef[C_KP_EPS]);
        }
        s_part[8 + rw][env] = acc;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four reward warps
  }
  if (REWARD && rw == 0) {
    // ---- warp 0: combine, terminate, count (one lane per env) -----------------------------------------
    const int env = renv;
    const bool live = renv < nvalid;
    float st[LG_NUM_STATS];
#pragma unroll
    for (int i = 0; i < LG_NUM_STATS; ++i) st[i] = 0.0f;
    if (live) {
      const int64_t e = e0 + env;
      const float dist = s_part[4][env], theta = s_part[6][env];
      // object_rot_delta (rewards.py:180-184): w * (ramp * (|theta| - |theta'|))
      const float t_delta = s_coef[C_DELTA_W] * (s_coef[C_DELTA_RAMP] * (fabsf(theta) - s_part[7][env]));
      const bool kp_on = EXT && ((P.term_active_mask >> LG_TERM_KEYPOINT) & 1);
      const float terms[7] = {s_part[0][env], s_part[1][env], s_part[2][env], s_part[5][env], t_delta, s_part[3][env],
                              kp_on ? s_coef[C_KP_W] * ((((s_part[8][env] + s_part[9][env]) + s_part[10][env]) + s_part[11][env]) * 0.125f)
                                    : 0.0f};
      float reward = 0.0f;  // trifinger_env.py:511, :551-553 — accumulation in dict order (the extension term last)
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        if ((P.term_active_mask >> k) & 1) { reward = reward + terms[k]; st[LG_STAT_TERM0 + k] = terms[k]; }
        if (B.term_rewards) B.term_rewards[(int64_t)k * P.num_envs + e] = terms[k];
      }
      // __check_termination (trifinger_env.py:1053-1099)
      const bool pos_ok = dist <= s_coef[C_POS_TOL];
      const bool rot_ok = theta <= s_coef[C_ROT_TOL];
      bool done;
      if (P.task_difficulty < 4) done = pos_ok;
      else if (P.task_difficulty == 4) done = pos_ok && rot_ok;
      else done = rot_ok;
      bool goal_reset = in_goal_reset != 0;
      bool succ = in_succ != 0;
      if (P.success_activate) {
        if (done) reward = reward + s_coef[C_BONUS];
        goal_reset = done;
        succ = succ || goal_reset;
        B.goal_reset[e] = goal_reset;
      } else {
        succ = goal_reset && succ;
      }
      B.successes[e] = succ;
      B.reward[e] = reward;
      // step counter, timeout, dones (envs/env_base.py:391-399)
      bool reset = in_reset != 0;
      if (P.fuse_bookkeeping) {
        const int64_t steps = in_steps + 1;
        B.steps_count[e] = steps;
        if (P.episode_length >= 0) reset = reset || (steps >= P.episode_length);
        B.reset[e] = reset;
      }
      const bool dn = reset && goal_reset;
      if (B.dones) B.dones[e] = dn;
      st[LG_STAT_POSITION_GOAL] = pos_ok;
      st[LG_STAT_ORIENTATION_GOAL] = rot_ok;
      st[LG_STAT_SUCCESSES] = succ;
      st[LG_STAT_REWARD] = reward;
      st[LG_STAT_RESETS] = reset;
      st[LG_STAT_DONES] = dn;
    }
    // ---- episode statistics: per-CTA fp64 sums in a fixed order, one RED per slot, no fence ----------
    if (env < E) {  // lanes beyond the tile hold nothing (E < 32)
#pragma unroll
      for (int i = 0; i <= LG_STAT_DONES; ++i) s_stat[i][env] = st[i];
    }
    __syncwarp();
    if (env <= LG_STAT_DONES) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four chains: the fp64 adds are dependent otherwise
#pragma unroll
      for (int k = 0; k < E; k += 4) {
        a0 += (double)s_stat[env][k];
        a1 += (double)s_stat[env][k + 1];
        a2 += (double)s_stat[env][k + 2];
        a3 += (double)s_stat[env][k + 3];
      }
      double acc = (a0 + a1) + (a2 + a3);
      // reward-term, reward and success entries are means over this shard (trifinger_env.py:554, :1098),
      // the rest are counts (:1067, :1076)
      const bool is_mean = env < LG_STAT_POSITION_GOAL || env == LG_STAT_SUCCESSES || env == LG_STAT_REWARD;  // 0..6: terms
      if (is_mean) acc = acc / (double)(P.stats_num_envs > 0 ? P.stats_num_envs : P.num_envs);
      atomicAdd(B.step_stats + env, acc);
    }
  }
  if (full) emit_outputs(std::true_type{});
  else emit_outputs(std::false_type{});
}

// history seeding (trifinger_env.py:619-628): both entries = initial simulator state
__global__ void init_history_kernel(const __grid_constant__ LgParams P, const LgSimState S, const LgBuffers B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.num_envs * LG_HISTORY_COLS) return;
  const int64_t e = i >> 4;
  const int c = (int)(i & 15);
  float v;
  if (c < 9) v = S.rigid_body[(e * P.bodies_per_env + P.fingertip_body[c / 3]) * 13 + (c % 3)];
  else v = S.root_state[((int64_t)P.actors_per_env * e + P.object_slot) * 13 + (c - 9)];
  B.history[i] = v;
}

// =========================================================================================
// pre-physics: ordered compaction by decoupled look-back, fused with the resets
// =========================================================================================
constexpr int kPreThreads = 128;  // one env per thread, one tile per CTA

// status word: [63:48] epoch | [47:46] state | [45:23] count A | [22:0] count B
constexpr uint64_t kStateAggregate = 1, kStateInclusive = 2;
__device__ __forceinline__ uint64_t pack_status(uint32_t epoch, uint64_t state, uint32_t a, uint32_t b) {
  return ((uint64_t)(epoch & 0xffffu) << 48) | (state << 46) | ((uint64_t)(a & 0x7fffffu) << 23) | (uint64_t)(b & 0x7fffffu);
}
__device__ __forceinline__ bool status_valid(uint64_t w, uint32_t epoch) {
  return (uint32_t)(w >> 48) == (epoch & 0xffffu) && ((w >> 46) & 3u) != 0;
}

// Exclusive prefix of (a, b) over all tiles before `tile`.  Block-wide: thread i inspects predecessor
// tile-1-i (128 predecessors per round trip to L2), each warp reduces up to its nearest tile that
// already holds an inclusive prefix, thread 0 chains the warps.  Called by every thread of the CTA.
__device__ __forceinline__ void lookback(uint64_t* status, int tile, uint32_t epoch, uint32_t my_a, uint32_t my_b,
                                         uint32_t& ex_a, uint32_t& ex_b) {
  __shared__ uint32_t s_sum[kPreThreads / 32][2];
  __shared__ int s_found[kPreThreads / 32];
  __shared__ uint32_t s_res[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t acc_a = 0, acc_b = 0;
  int pos = tile - 1;
  bool done = tile == 0;
  while (!done) {
    const int idx = pos - (int)threadIdx.x;
    uint64_t w = 0;
    if (idx >= 0) {
      do { w = ld_volatile_u64(status + idx); } while (!status_valid(w, epoch));
    }
    const bool incl = idx >= 0 && ((w >> 46) & 3u) == kStateInclusive;
    const unsigned incl_mask = __ballot_sync(0xffffffffu, incl);
    const int stop = incl_mask ? __ffs(incl_mask) - 1 : 31;  // nearest predecessor holding an inclusive prefix
    uint32_t a = (idx >= 0 && lane <= stop) ? (uint32_t)((w >> 23) & 0x7fffffu) : 0u;
    uint32_t b = (idx >= 0 && lane <= stop) ? (uint32_t)(w & 0x7fffffu) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { s_sum[warp][0] = a; s_sum[warp][1] = b; s_found[warp] = incl_mask != 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
      bool found = false;
      for (int k = 0; k < kPreThreads / 32 && !found; ++k) { acc_a += s_sum[k][0]; acc_b += s_sum[k][1]; found = s_found[k] != 0; }
      s_res[0] = acc_a; s_res[1] = acc_b; s_res[2] = found || pos - kPreThreads < 0;
    }
    __syncthreads();
    done = s_res[2] != 0;
    pos -= kPreThreads;
  }
  if (threadIdx.x == 0) {
    ex_a = acc_a; ex_b = acc_b;
    if (tile > 0)
      atomicExch(reinterpret_cast<unsigned long long*>(status + tile),
                 (unsigned long long)pack_status(epoch, kStateInclusive, acc_a + my_a, acc_b + my_b));
  }
}

struct TileScan {
  int tile;
  uint32_t epoch;
  uint32_t rank_a, rank_b;   // this thread's rank inside the tile (valid where its flag is set)
  uint32_t total_a, total_b; // tile totals
};

// Resolves the global exclusive prefix; the last tile re-arms ticket and epoch for the next launch.
template <bool TICKET>
__device__ __forceinline__ void tile_scan_finish(LgControl* ctl, uint64_t* status, const TileScan& t, int num_tiles,
                                                 uint32_t& ex_a, uint32_t& ex_b, int32_t* counts_out) {
  __shared__ uint32_t s_ex[2];
  uint32_t a = 0, b = 0;
  lookback(status, t.tile, t.epoch, t.total_a, t.total_b, a, b);
  if (threadIdx.x == 0) {
    s_ex[0] = a; s_ex[1] = b;
    if (t.tile == num_tiles - 1) {
      // every tile has read the epoch and taken its ticket by now (their aggregates are visible)
      if (counts_out) { counts_out[0] = (int32_t)(a + t.total_a); counts_out[1] = (int32_t)(b + t.total_b); }
      if (TICKET) { ctl->scan_ticket = 0; __threadfence(); }
      ctl->scan_epoch = t.epoch + 1;
    }
  }
  __syncthreads();
  ex_a = s_ex[0]; ex_b = s_ex[1];
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP): one thread moves a whole contiguous tile slab -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int A, bool TICKET>
__global__ void __launch_bounds__(kPreThreads)
pre_physics_kernel(const __grid_constant__ LgParams P, const __grid_constant__ LgSimState S,
                   const __grid_constant__ LgBuffers B,
                   const float* __restrict__ action_in, int num_tiles) {
  constexpr int E = kPreThreads, NT = kPreThreads;
  // tile slabs: the action and joint-state rows of the tile's 128 envs are contiguous in memory, so they
  // come in (and the action / torque rows go out) as TMA bulk copies issued by one thread
  __shared__ __align__(128) float s_act[E * A];
  __shared__ __align__(128) float s_dof[E * 18];
  __shared__ __align__(128) float s_tq[E * 9];
  __shared__ __align__(8) uint64_t s_mbar;
  __shared__ int s_tile;
  __shared__ uint32_t s_wa[NT / 32], s_wb[NT / 32];
  const int tid = threadIdx.x;
  // Launched programmatically after lg_post_physics (see pdl_mode): what this kernel reads before pdl_wait() below
  // must not be written by that kernel.  The control block (epoch, ticket) is only written by this kernel's own
  // previous launch, the action and the joint state by the caller / simulator — all complete before the preceding
  // post-physics pass was allowed past its own dependency wait.
  // every thread reads the epoch before the tile publishes anything, so the last tile may advance it
  const uint32_t epoch = ld_volatile_u32(&B.control->scan_epoch);
  int tile = blockIdx.x;
  if (TICKET) {  // grids larger than what is co-resident: tiles by ticket, so predecessors always run
    if (tid == 0) s_tile = (int)atomicAdd(&B.control->scan_ticket, 1u);
    __syncthreads();
    tile = s_tile;
  }
  const int64_t e0 = (int64_t)tile * E;
  const int64_t e = e0 + tid;
  const int nvalid = (int)min((int64_t)E, P.num_envs - e0);
  const bool live = tid < nvalid;
  const bool want_torque = B.applied_torque != nullptr;
  const bool full_tile = nvalid == E;   // bulk copies need 16-byte multiples: ragged last tile goes lane by lane

  // ---- every global load of the tile up front ---------------------------------------------------
  if (full_tile) {
    if (tid == 0) {
      mbar_init(&s_mbar, 1);
      mbar_expect_tx(&s_mbar, (uint32_t)(sizeof(float) * E * (A + (want_torque ? 18 : 0))));
      bulk_load(s_act, action_in + e0 * A, sizeof(float) * E * A, &s_mbar);
      if (want_torque) bulk_load(s_dof, S.dof_state + e0 * 18, sizeof(float) * E * 18, &s_mbar);
    }
  } else if (live) {
#pragma unroll
    for (int c = 0; c < A; ++c) s_act[tid * A + c] = action_in[e * A + c];
    if (want_torque) {
#pragma unroll
      for (int c = 0; c < 18; ++c) s_dof[tid * 18 + c] = S.dof_state[e * 18 + c];
    }
  }
  pdl_wait();   // the flags, counters and statistics below are results of the preceding post-physics pass
  uint8_t flag_r = 0, flag_g = 0;
  if (live) {
    flag_r = B.reset[e]; flag_g = B.goal_reset[e];
    if (B.force_reset) flag_r |= B.force_reset[e];             // `_reset_buf |= mask` folded into the pass
    if (B.force_goal_reset) flag_g |= B.force_goal_reset[e];
  }
  if (tile == 0 && tid < LG_NUM_STATS && B.step_stats) B.step_stats[tid] = 0.0;  // accumulated by lg_post_physics
  if (tile == 0 && tid == NT - 1 && P.use_device_clock) {
    // device clock: advance the frame counter and publish the reward coefficients of the coming
    // post-physics pass (env_steps_count = frames x global env count, envs/env_base.py:286-289);
    // done by an otherwise idle lane while the tile's loads are in flight
    const int64_t frame = B.control->frame_count + P.control_decimation;
    B.control->frame_count = frame;
    if (B.reward_coef) compute_coefs(P, (double)(frame * P.global_num_envs), B.reward_coef);
  }

  // ---- block scan of both masks + aggregate publication (env_base.py:374-379) -------------------
  // Needs only the two flag bytes, so it runs (and the tile's aggregate is visible to its successors)
  // while the action / joint-state slabs are still in flight; the look-back at the end then never waits.
  const bool f_reset = flag_r != 0, f_goal = flag_g != 0;
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned ba = __ballot_sync(0xffffffffu, f_reset), bb = __ballot_sync(0xffffffffu, f_goal);
  if (lane == 0) { s_wa[warp] = __popc(ba); s_wb[warp] = __popc(bb); }
  __syncthreads();   // also orders the mbarrier initialisation before the waits below
  TileScan t;
  t.tile = tile; t.epoch = epoch;
  {
    uint32_t pa = 0, pb = 0, ta = 0, tb = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
      if (w < warp) { pa += s_wa[w]; pb += s_wb[w]; }
      ta += s_wa[w]; tb += s_wb[w];
    }
    const unsigned below = (1u << lane) - 1u;
    t.rank_a = pa + __popc(ba & below);
    t.rank_b = pb + __popc(bb & below);
    t.total_a = ta; t.total_b = tb;
  }
  if (tid == 0) {
    const uint64_t st = tile == 0 ? kStateInclusive : kStateAggregate;
    atomicExch(reinterpret_cast<unsigned long long*>(B.scan_status + tile),
               (unsigned long long)pack_status(epoch, st, t.total_a, t.total_b));
  }
  uint32_t ex_a = 0, ex_b = 0;
  const bool need_rank_first = P.inject_draws != 0;  // injected draws are indexed by compaction rank
  if (need_rank_first) tile_scan_finish<TICKET>(B.control, B.scan_status, t, num_tiles, ex_a, ex_b, B.counts);

  if (full_tile) mbar_wait(&s_mbar, 0);   // the slabs have landed (each thread only touches its own env row below)

  // ---- this env's action row: noise (extension), clamp, reset ---------------------------------------
  float act[A];
  if (live) {
#pragma unroll
    for (int c = 0; c < A; ++c) act[c] = s_act[tid * A + c];
    if (P.dr_activate && P.dr_action_sigma != 0.0f) {  // extension (no reference code): Gaussian action noise
      const uint64_t genv = (uint64_t)(P.env_offset + e);
#pragma unroll
      for (int c4 = 0; c4 < A; c4 += 4) {   // one Philox block -> four normals -> four action columns
        const U4 r = philox4x32_10(U4{(uint32_t)genv, (uint32_t)(genv >> 32) ^ kPurposeNoise, 0x41435400u + c4, epoch},
                                   (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
        float n[4];
        box_muller(r.x, r.y, n[0], n[1]);
        box_muller(r.z, r.w, n[2], n[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c4 + q < A) act[c4 + q] = act[c4 + q] + P.dr_action_sigma * n[q];
      }
    }
    if (P.clip_input_actions) {   // the wrapper's clamp (wrappers/vec_task.py:162)
      const float clip = P.clip_actions;
#pragma unroll
      for (int c = 0; c < A; ++c) act[c] = fminf(fmaxf(act[c], -clip), clip);
    }
    if (f_reset) {                // the row is zeroed after the store (envs/env_base.py:369, trifinger_env.py:387)
#pragma unroll
      for (int c = 0; c < A; ++c) act[c] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < A; ++c) s_act[tid * A + c] = act[c];
  }
  // ---- resets (trifinger_env.py:373-440) -----------------------------------------------------------------
  // Block-uniform branch: a tile without flagged envs neither fetches nor executes sampler code.  The flagged envs
  // are listed in shared memory by their rank in the tile, then each WARP runs two of the eight reset sub-tasks over
  // that list (lane = listed env): uniform control flow, and a serial chain of ~2 Philox blocks per warp instead of
  // ten per resetting thread.
  if ((t.total_a | t.total_b) != 0) {
    __shared__ uint16_t s_reset_list[E], s_goal_list[E];
    if (f_reset) s_reset_list[t.rank_a] = (uint16_t)(tid | (f_goal ? 0x8000 : 0));
    if (f_goal) s_goal_list[t.rank_b] = (uint16_t)tid;
    __syncthreads();
    const int na = (int)t.total_a, nb = (int)t.total_b;
    auto run_sub = [&](int sub, int first, int step) {
#pragma unroll 1
      for (int r = first; r < na; r += step) {
        const int ent = s_reset_list[r], local = ent & 0x7fff;
        const int64_t env = e0 + local;
        const DrawSource dr = make_draws(P, (uint64_t)epoch, env, kPurposeReset, B.inject_reset_u, B.inject_reset_n,
                                         (int64_t)ex_a + r);
        reset_subtask(P, S, B, env, sub, dr, (ent & 0x8000) != 0, s_dof + local * 18);
      }
    };
    // the two long sub-tasks (object pose: 2 Philox blocks, sqrt, 2 sincos; goal: 2-3 blocks, Box-Muller, normalise)
    // take two warps each, the six short ones (joint blocks 0..4, bookkeeping) are spread over the four warps
    run_sub(warp < 2 ? 5 : 6, tid & 63, 64);
    run_sub(warp, lane, 32);                                   // joint blocks 0..3
    if (warp < 2) run_sub(warp == 0 ? 4 : 7, lane, 32);        // joint block 4, bookkeeping
#pragma unroll 1
    for (int r = tid; r < nb; r += NT) {   // goal resets second, as in env_base.py:374-379
      const int64_t env = e0 + s_goal_list[r];
      const DrawSource dr = make_draws(P, (uint64_t)epoch, env, kPurposeGoal, B.inject_goal_u, B.inject_goal_n,
                                       (int64_t)ex_b + r);
      B.goal_reset[env] = 0;  // trifinger_env.py:427
      apply_goal_sample(P, S, B, env, dr);
    }
    __syncthreads();   // the torque and goal-movement code below reads rows other lanes have just rewritten
  }
  // ---- moving goal (__update_goal_movement_pre, trifinger_env.py:1267-1277): every step the goal body's
  // angular velocity is re-imposed from the movement buffer (freshly sampled above for envs that reset)
  if (P.goal_rotation && live) {
    float* row = S.root_state + (P.actors_per_env * e + P.goal_slot) * 13;
    const float* gm = B.goal_movement + e * 6;
    row[10] = gm[3]; row[11] = gm[4]; row[12] = gm[5];
  }
  // ---- action -> torque (trifinger_env.py:442-498), on the post-reset joint state ---------------
  if (want_torque && live) {
    torque_one_env(P, act, s_dof + tid * 18, s_tq + tid * 9);   // resets mirrored their joint rows into s_dof
  }
  // The post-physics pass may start launching now.  Triggering earlier parks its CTAs (which fill the register file)
  // next to this kernel's one warp per scheduler and slows the latency chain above; measured, us/step at 16k envs /
  // 30 % resets: right after the flag loads 11.50 / 16.95, after the slab wait 11.38 / 16.83, here 11.28 / 16.50,
  // no explicit trigger 11.98 / 17.24.
  pdl_launch_dependents();
  if (full_tile) {
    fence_async_proxy();           // generic-proxy writes to the slabs -> visible to the bulk-copy engine
    __syncthreads();
    if (tid == 0) {
      bulk_store(B.action + e0 * A, s_act, sizeof(float) * E * A);
      if (want_torque) bulk_store(B.applied_torque + e0 * 9, s_tq, sizeof(float) * E * 9);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else if (live) {
#pragma unroll
    for (int c = 0; c < A; ++c) B.action[e * A + c] = act[c];
    if (want_torque) {
#pragma unroll
      for (int c = 0; c < 9; ++c) B.applied_torque[e * 9 + c] = s_tq[tid * 9 + c];
    }
  }

  // ---- ordered id lists (env_base.py:374-379; trifinger_env.py:413-416, :435-436) ---------------
  if (!need_rank_first) tile_scan_finish<TICKET>(B.control, B.scan_status, t, num_tiles, ex_a, ex_b, B.counts);
  if (f_reset) {
    const int64_t j = (int64_t)ex_a + t.rank_a;
    const int32_t base = (int32_t)(P.actors_per_env * e);
    B.reset_ids[j] = e;
    if (B.robot_indices) B.robot_indices[j] = base + P.robot_slot;
    if (B.reset_root_indices) {  // unique(cat(robot, object, goal)) == sorted, and sorted == per-env triples
      B.reset_root_indices[3 * j] = base + P.robot_slot;
      B.reset_root_indices[3 * j + 1] = base + P.object_slot;
      B.reset_root_indices[3 * j + 2] = base + P.goal_slot;
    }
  }
  if (f_goal) {
    const int64_t j = (int64_t)ex_b + t.rank_b;
    B.goal_reset_ids[j] = e;
    if (B.goal_root_indices) B.goal_root_indices[j] = (int32_t)(P.actors_per_env * e) + P.goal_slot;
  }
  // shared memory must outlive the bulk stores that read it
  if (full_tile && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// standalone compaction (torch.nonzero(mask).view(-1))
__global__ void __launch_bounds__(kPreThreads)
compact_kernel(const uint8_t* __restrict__ mask, int64_t n, int64_t* __restrict__ ids, int32_t* counts2,
               uint64_t* status, LgControl* ctl, int num_tiles) {
  __shared__ int s_tile;
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = ld_volatile_u32(&ctl->scan_epoch);
    s_tile = (int)atomicAdd(&ctl->scan_ticket, 1u);
  }
  __syncthreads();
  const int64_t e = (int64_t)s_tile * kPreThreads + threadIdx.x;
  const bool f = e < n && mask[e] != 0;
  __shared__ uint32_t s_w[kPreThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, f);
  if (lane == 0) s_w[warp] = __popc(b);
  __syncthreads();
  TileScan t;
  t.tile = s_tile; t.epoch = s_epoch;
  uint32_t p = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kPreThreads / 32; ++w) { if (w < warp) p += s_w[w]; tot += s_w[w]; }
  t.rank_a = p + __popc(b & ((1u << lane) - 1u));
  t.rank_b = 0; t.total_a = tot; t.total_b = 0;
  if (threadIdx.x == 0)
    atomicExch(reinterpret_cast<unsigned long long*>(status + t.tile),
               (unsigned long long)pack_status(t.epoch, t.tile == 0 ? kStateInclusive : kStateAggregate, tot, 0));
  uint32_t ex_a, ex_b;
  tile_scan_finish<true>(ctl, status, t, num_tiles, ex_a, ex_b, counts2);
  if (f) ids[(int64_t)ex_a + t.rank_a] = e;
}

// hooks on explicit id lists
__global__ void reset_ids_kernel(const __grid_constant__ LgParams P, const LgSimState S, const LgBuffers B,
                                 const int64_t* __restrict__ ids, int64_t k, int goal_only) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < k) {
    const int64_t e = ids[j];
    const uint64_t epoch = B.control->rng_epoch | (1ull << 63);  // never collides with the fused path's epochs
    if (goal_only) {
      const DrawSource dr = make_draws(P, epoch, e, kPurposeGoal, B.inject_goal_u, B.inject_goal_n, j);
      B.goal_reset[e] = 0;
      apply_goal_sample(P, S, B, e, dr);
      if (B.goal_root_indices) B.goal_root_indices[j] = (int32_t)(P.actors_per_env * e) + P.goal_slot;
    } else {
      const DrawSource dr = make_draws(P, epoch, e, kPurposeReset, B.inject_reset_u, B.inject_reset_n, j);
      reset_one_env(P, S, B, e, dr);
      const int32_t base = (int32_t)(P.actors_per_env * e);
      if (B.robot_indices) B.robot_indices[j] = base + P.robot_slot;
      if (B.reset_root_indices) {
        B.reset_root_indices[3 * j] = base + P.robot_slot;
        B.reset_root_indices[3 * j + 1] = base + P.object_slot;
        B.reset_root_indices[3 * j + 2] = base + P.goal_slot;
      }
    }
  }
}
__global__ void bump_epoch_kernel(LgControl* ctl) { ctl->rng_epoch += 1; }

__global__ void pre_step_kernel(const __grid_constant__ LgParams P, const LgSimState S, const LgBuffers B) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.num_envs) return;
  float act[LG_MAX_ACTION_DIM];
  for (int c = 0; c < P.action_dim; ++c) act[c] = B.action[e * P.action_dim + c];
  torque_one_env(P, act, S.dof_state + e * 18, B.applied_torque + e * 9);
  if (P.goal_rotation) {  // __update_goal_movement_pre (trifinger_env.py:1267-1277)
    float* row = S.root_state + (P.actors_per_env * e + P.goal_slot) * 13;
    const float* gm = B.goal_movement + e * 6;
    row[10] = gm[3]; row[11] = gm[4]; row[12] = gm[5];
  }
}

// ---- batched primitives -------------------------------------------------------------------------
__global__ void quat_mul_kernel(const float* a, const float* b, float* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 qa = reinterpret_cast<const float4*>(a)[i], qb = reinterpret_cast<const float4*>(b)[i];
  const Quat r = quat_mul(Quat{qa.x, qa.y, qa.z, qa.w}, Quat{qb.x, qb.y, qb.z, qb.w});
  reinterpret_cast<float4*>(out)[i] = make_float4(r.x, r.y, r.z, r.w);
}
__global__ void quat_diff_kernel(const float* a, const float* b, float* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 qa = reinterpret_cast<const float4*>(a)[i], qb = reinterpret_cast<const float4*>(b)[i];
  out[i] = quat_diff_rad(Quat{qa.x, qa.y, qa.z, qa.w}, Quat{qb.x, qb.y, qb.z, qb.w});
}
template <int OP>
__global__ void rowwise_kernel(const float* x, const float* lo, const float* hi, float* out, int64_t total, int dims) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % dims);
  const float l = lo[c], h = hi[c], v = x[i];
  if (OP == 0) out[i] = scale_transform(v, (l + h) * 0.5f, h - l);
  else if (OP == 1) out[i] = unscale_transform(v, l, h);
  else out[i] = saturate(v, l, h);
}
__global__ void lgsk_kernel(const float* x, float scale, float* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = lgsk(x[i], scale);
}
__global__ void keypoints_kernel(const float* pose, float size, float* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 8) return;
  const int64_t e = i >> 3;
  const int k = (int)(i & 7);
  const float* p = pose + e * 7;
  const float h = size * 0.5f;
  const float vx = (k & 1) ? h : -h, vy = (k & 2) ? h : -h, vz = (k & 4) ? h : -h;
  float rx, ry, rz;
  quat_rotate(Quat{p[3], p[4], p[5], p[6]}, vx, vy, vz, rx, ry, rz);
  out[i * 3] = p[0] + rx; out[i * 3 + 1] = p[1] + ry; out[i * 3 + 2] = p[2] + rz;
}

// Stand-in for the simulator's rigid-body integration of the (gravity-free, collision-free) goal body of the
// moving-goal task: semi-implicit Euler over dt on root rows [n, 13] = pos 3 | quat xyzw 4 | lin vel 3 | ang vel 3,
//   p += v dt;   q <- normalize(exp(w dt / 2) (x) q)   (world-frame angular velocity, quaternion exponential).
__global__ void integrate_rows_kernel(float* rows, int64_t row_stride, float dt, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* r = rows + i * row_stride;
  const float wx = r[10], wy = r[11], wz = r[12];
  r[0] = r[0] + r[7] * dt; r[1] = r[1] + r[8] * dt; r[2] = r[2] + r[9] * dt;
  const float w = norm3(wx, wy, wz);
  if (!(w > 0.0f)) return;   // not turning: the orientation stays bit-identical
  const float half = 0.5f * w * dt;
  const float k = w > 1e-12f ? __fdiv_rn(sinf(half), w) : 0.5f * dt;   // sin(half)/w -> dt/2 as w -> 0
  const Quat dq{wx * k, wy * k, wz * k, cosf(half)};
  Quat q = quat_mul(dq, Quat{r[3], r[4], r[5], r[6]});
  const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  const float inv = __frcp_rn(fmaxf(nrm, 1e-12f));
  r[3] = q.x * inv; r[4] = q.y * inv; r[5] = q.z * inv; r[6] = q.w * inv;
}

// exhaustive check of div_by_const's contract for one (span, rcp): all 2^32 numerators.
// out[0] = in-window numerators that are not bit-identical to IEEE division (must be 0);
// out[1] = worst ulp distance for 0 < |x| < 2^-100 (must be <= 1); out[2] = huge/inf numerators not flagged.
__global__ void selftest_division_kernel(float span, float rcp, unsigned long long* out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0, unflagged = 0;
  unsigned int worst = 0;
  for (uint64_t bits = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; bits < (1ull << 32); bits += stride) {
    const float x = __uint_as_float((uint32_t)bits);
    float amax = 0.0f;
    const float a = div_by_const(x, span, rcp, amax), b = __fdiv_rn(x, span);
    const float ax = fabsf(x);
    if (x != x) continue;                                   // nan in, nan out on both paths
    if (!(ax < kDivSafeMax)) { if (amax < kDivSafeMax) ++unflagged; continue; }
    if (ax == 0.0f) { if (a != 0.0f) ++bad; continue; }
    if (ax >= 7.8886091e-31f /* 2^-100 */) { if (__float_as_uint(a) != __float_as_uint(b)) ++bad; continue; }
    const int d = abs((int)(__float_as_uint(a) & 0x7fffffffu) - (int)(__float_as_uint(b) & 0x7fffffffu));
    worst = max(worst, (unsigned int)d);
  }
  if (bad) atomicAdd(out, bad);
  if (worst) atomicMax(out + 1, (unsigned long long)worst);
  if (unflagged) atomicAdd(out + 2, unflagged);
}

}  // namespace lg

// =========================================================================================
// C ABI
// =========================================================================================
namespace {
thread_local std::string g_error;
int fail(int code, const std::string& msg) { g_error = msg; return code; }
int check_launch(const char* what) {
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(LG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(err));
  return LG_OK;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool env_flag(const char* name) { const char* v = std::getenv(name); return v && v[0] && v[0] != '0'; }
int env_int(const char* name, int dflt) { const char* v = std::getenv(name); return v ? std::atoi(v) : dflt; }

int sm_count() {
  static const int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

// bit 0: pre-physics kernel launched programmatically, bit 1: post-physics kernel.  Measured on B200 (us/step):
// 16k envs, no resets: post only 11.5-11.7, both 11.4-11.5, none 11.9; 30 % resets: post only 17.5, both 21.6 —
// a long pre-physics pass should not have the post pass's CTAs parked on the SMs, so the default is post only.
int pdl_mode() { static const int m = env_flag("LG_NO_PDL") ? 0 : env_int("LG_PDL", 2); return m; }

// Launch with programmatic stream serialisation (PDL) unless LG_NO_PDL is set.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Envs per CTA of the post-physics kernel: 32 when the grid needs several waves anyway, otherwise the
// smallest tile that still fits one wave (4 CTAs per SM) so that all SMs carry the same load.
int pick_tile_envs(int64_t n) {
  static const int forced = env_int("LG_TILE_ENVS", 0);
  if (forced == 32 || forced == 28 || forced == 24 || forced == 16) return forced;
  const int64_t capacity = (int64_t)sm_count() * 4;
  if ((n + 31) / 32 > capacity) return 32;
  const int cands[] = {16, 24, 28, 32};
  for (int e : cands)
    if ((n + e - 1) / e <= capacity) return e;
  return 32;
}

int validate(const LgParams* P, const LgSimState* S, const LgBuffers* B, bool need_states) {
  if (!P || !S || !B) return fail(LG_ERR_BAD_ARG, "null LgParams / LgSimState / LgBuffers");
  if (P->num_envs <= 0) return fail(LG_ERR_BAD_ARG, "num_envs must be positive");
  if (P->num_envs >= (1 << 23)) return fail(LG_ERR_BAD_ARG, "num_envs per shard must be < 2^23");
  if (P->action_dim != 9 && P->action_dim != 18) return fail(LG_ERR_BAD_ARG, "action_dim must be 9 or 18");
  if (!S->dof_state || !S->root_state || !S->rigid_body) return fail(LG_ERR_BAD_ARG, "null simulator tensor");
  if (!B->obs || !B->action || !B->reward || !B->reset || !B->goal_reset || !B->successes || !B->steps_count ||
      !B->goal_pose || !B->goal_movement || !B->history || !B->control)
    return fail(LG_ERR_BAD_ARG, "null env buffer");
  if (need_states && P->asymmetric_obs && (!B->states || !S->dof_force || !S->ft_sensors))
    return fail(LG_ERR_BAD_ARG, "asymmetric_obs needs states, dof_force and ft_sensors");
  const void* al[] = {S->dof_state, S->dof_force, S->ft_sensors, B->obs, B->states, B->obs_clipped, B->states_clipped,
                      B->action, B->goal_pose, B->history, B->applied_torque};
  for (const void* p : al)
    if (p && !aligned16(p)) return fail(LG_ERR_BAD_ARG, "tensor base pointers must be 16-byte aligned");
  return LG_OK;
}

template <bool REWARD>
int launch_post(const LgParams* P, const LgSimState* S, const LgBuffers* B, double sched, cudaStream_t st) {
  if (int rc = validate(P, S, B, true)) return rc;
  if (REWARD && !B->step_stats) return fail(LG_ERR_BAD_ARG, "null statistics buffer");
  if (P->normalize_obs && !B->scale_table) return fail(LG_ERR_BAD_ARG, "null scale_table");
  if (REWARD && !P->fuse_bookkeeping) {
    // stand-alone _post_step: nobody zeroed the accumulators (lg_pre_physics does in the fused sequence)
    if (cudaMemsetAsync(B->step_stats, 0, LG_NUM_STATS * sizeof(double), st) != cudaSuccess) return check_launch("memset");
  }
  const bool asym = P->asymmetric_obs != 0;
  const bool clip = B->obs_clipped != nullptr;
  if (clip && asym && !B->states_clipped) return fail(LG_ERR_BAD_ARG, "obs_clipped and states_clipped go together");
  LgCoef cf;
  std::memset(&cf, 0, sizeof(cf));
  if (REWARD && P->use_device_clock) {
    if (!B->reward_coef) return fail(LG_ERR_BAD_ARG, "use_device_clock needs reward_coef");
  } else {
    lg::compute_coefs(*P, sched, cf.v);
  }
  const int E = P->action_dim == 9 ? pick_tile_envs(P->num_envs) : 32;
  const unsigned grid = (unsigned)((P->num_envs + E - 1) / E);
  cudaError_t err;
  // rarely-used paths (DR noise, keypoint term, moving goal, bf16 copies) live in their own instantiation: switched off they cost nothing
  const bool ext = P->dr_activate || P->goal_rotation || ((P->term_active_mask >> LG_TERM_KEYPOINT) & 1) || B->obs_bf16 || B->states_bf16;
#define LG_K(AD, AS, CL, EE) (ext ? launch_pdl(pdl_mode() & 2, lg::post_physics_kernel<AD, AS, REWARD, CL, EE, true>, grid, lg::kPostThreads, st, *P, *S, *B, cf) \
                                  : launch_pdl(pdl_mode() & 2, lg::post_physics_kernel<AD, AS, REWARD, CL, EE, false>, grid, lg::kPostThreads, st, *P, *S, *B, cf))
#define LG_E(AD, AS, CL) (E == 16 ? LG_K(AD, AS, CL, 16) : E == 24 ? LG_K(AD, AS, CL, 24) : E == 28 ? LG_K(AD, AS, CL, 28) : LG_K(AD, AS, CL, 32))
  if (P->action_dim == 9) {
    if (asym) err = clip ? LG_E(9, true, true) : LG_E(9, true, false);
    else err = clip ? LG_E(9, false, true) : LG_E(9, false, false);
  } else {
    if (asym) err = clip ? LG_K(18, true, true, 32) : LG_K(18, true, false, 32);
    else err = clip ? LG_K(18, false, true, 32) : LG_K(18, false, false, 32);
  }
#undef LG_E
#undef LG_K
  if (err != cudaSuccess) return fail(LG_ERR_CUDA, std::string("post_physics_kernel: ") + cudaGetErrorString(err));
  return check_launch("post_physics_kernel");
}
}  // namespace

extern "C" {

int lg_version(void) { return LG_VERSION; }
int lg_set_l2_fetch_granularity(int bytes) {
  if (bytes != 32 && bytes != 64 && bytes != 128) return fail(LG_ERR_BAD_ARG, "granularity must be 32, 64 or 128");
  if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes) != cudaSuccess) return check_launch("cudaDeviceSetLimit");
  return LG_OK;
}
const char* lg_last_error(void) { return g_error.c_str(); }
size_t lg_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(LgParams);
    case 1: return sizeof(LgSimState);
    case 2: return sizeof(LgBuffers);
    case 3: return sizeof(LgControl);
    case 4: return sizeof(LgRewardTerm);
    case 5: return sizeof(LgHostStep);
    default: return 0;
  }
}
int64_t lg_scan_tiles(int64_t n) { return (n + lg::kPreThreads - 1) / lg::kPreThreads; }

int lg_post_physics(const LgParams* P, const LgSimState* S, const LgBuffers* B, double sched_step, void* stream) {
  return launch_post<true>(P, S, B, sched_step, (cudaStream_t)stream);
}
int lg_fill_observations(const LgParams* P, const LgSimState* S, const LgBuffers* B, void* stream) {
  return launch_post<false>(P, S, B, 0.0, (cudaStream_t)stream);
}
int lg_init_history(const LgParams* P, const LgSimState* S, const LgBuffers* B, void* stream) {
  if (int rc = validate(P, S, B, false)) return rc;
  const int64_t total = P->num_envs * LG_HISTORY_COLS;
  lg::init_history_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*P, *S, *B);
  return check_launch("init_history_kernel");
}

int lg_pre_physics(const LgParams* P, const LgSimState* S, const LgBuffers* B, const float* action_in, void* stream) {
  if (int rc = validate(P, S, B, false)) return rc;
  if (!action_in) return fail(LG_ERR_BAD_ARG, "null action_in");
  if (!B->scan_status || !B->reset_ids || !B->goal_reset_ids || !B->counts)
    return fail(LG_ERR_BAD_ARG, "null compaction buffer");
  if (P->inject_draws && !(B->inject_reset_u || B->inject_goal_u || B->inject_reset_n || B->inject_goal_n))
    return fail(LG_ERR_BAD_ARG, "inject_draws set without injected arrays");
  const int tiles = (int)lg_scan_tiles(P->num_envs);
  if (!aligned16(action_in)) return fail(LG_ERR_BAD_ARG, "action_in must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  // all tiles co-resident (<= 8 CTAs on each of 148 SMs): tiles are block indices; beyond that, tickets
  const bool ticket = tiles > 148 * 8;
cudaError_t err;
#define LG_PRE(AD, TK) err = launch_pdl(pdl_mode() & 1, lg::pre_physics_kernel<AD, TK>, (unsigned)tiles, lg::kPreThreads, st, *P, *S, *B, action_in, tiles)
  if (P->action_dim == 9) { if (ticket) LG_PRE(9, true); else LG_PRE(9, false); }
  else { if (ticket) LG_PRE(18, true); else LG_PRE(18, false); }
  if (err != cudaSuccess) return fail(LG_ERR_CUDA, std::string("pre_physics_kernel: ") + cudaGetErrorString(err));
#undef LG_PRE
  return check_launch("pre_physics_kernel");
}

int lg_compact(const uint8_t* mask, int64_t n, int64_t* ids_out, int32_t* count_out, uint64_t* status,
               LgControl* control, void* stream) {
  if (!mask || !ids_out || !count_out || !status || !control) return fail(LG_ERR_BAD_ARG, "null argument");
  if (n < 0 || n >= (1 << 23)) return fail(LG_ERR_BAD_ARG, "n out of range");
  if (n == 0) { cudaMemsetAsync(count_out, 0, 2 * sizeof(int32_t), (cudaStream_t)stream); return check_launch("memset"); }
  const int tiles = (int)lg_scan_tiles(n);
  lg::compact_kernel<<<tiles, lg::kPreThreads, 0, (cudaStream_t)stream>>>(mask, n, ids_out, count_out, status, control, tiles);
  return check_launch("compact_kernel");
}

static int reset_list(const LgParams* P, const LgSimState* S, const LgBuffers* B, const int64_t* ids, int64_t k,
                      int goal_only, void* stream) {
  if (int rc = validate(P, S, B, false)) return rc;
  if (k < 0 || (k > 0 && !ids)) return fail(LG_ERR_BAD_ARG, "bad id list");
  if (k == 0) return LG_OK;
  lg::reset_ids_kernel<<<(unsigned)((k + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*P, *S, *B, ids, k, goal_only);
  if (int rc = check_launch("reset_ids_kernel")) return rc;
  lg::bump_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(B->control);
  return check_launch("bump_epoch_kernel");
}
int lg_reset_envs(const LgParams* P, const LgSimState* S, const LgBuffers* B, const int64_t* ids, int64_t k, void* stream) {
  return reset_list(P, S, B, ids, k, 0, stream);
}
int lg_goal_reset_envs(const LgParams* P, const LgSimState* S, const LgBuffers* B, const int64_t* ids, int64_t k, void* stream) {
  return reset_list(P, S, B, ids, k, 1, stream);
}

int lg_pre_step(const LgParams* P, const LgSimState* S, const LgBuffers* B, void* stream) {
  if (int rc = validate(P, S, B, false)) return rc;
  if (!B->applied_torque) return fail(LG_ERR_BAD_ARG, "null applied_torque");
  lg::pre_step_kernel<<<(unsigned)((P->num_envs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*P, *S, *B);
  return check_launch("pre_step_kernel");
}

#define LG_GRID(n) (unsigned)(((n) + 255) / 256), 256, 0, (cudaStream_t)stream
int lg_quat_mul(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (!a || !b || !out || n < 0) return fail(LG_ERR_BAD_ARG, "bad argument");
  if (!aligned16(a) || !aligned16(b) || !aligned16(out)) return fail(LG_ERR_BAD_ARG, "quaternion arrays must be 16-byte aligned");
  if (n) lg::quat_mul_kernel<<<LG_GRID(n)>>>(a, b, out, n);
  return check_launch("quat_mul_kernel");
}
int lg_quat_diff_rad(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (!a || !b || !out || n < 0) return fail(LG_ERR_BAD_ARG, "bad argument");
  if (!aligned16(a) || !aligned16(b)) return fail(LG_ERR_BAD_ARG, "quaternion arrays must be 16-byte aligned");
  if (n) lg::quat_diff_kernel<<<LG_GRID(n)>>>(a, b, out, n);
  return check_launch("quat_diff_kernel");
}
static int rowwise(int op, const float* x, const float* lo, const float* hi, float* out, int64_t n, int32_t dims, void* stream) {
  if (!x || !lo || !hi || !out || n < 0 || dims <= 0) return fail(LG_ERR_BAD_ARG, "bad argument");
  const int64_t total = n * dims;
  if (total == 0) return LG_OK;
  if (op == 0) lg::rowwise_kernel<0><<<LG_GRID(total)>>>(x, lo, hi, out, total, dims);
  else if (op == 1) lg::rowwise_kernel<1><<<LG_GRID(total)>>>(x, lo, hi, out, total, dims);
  else lg::rowwise_kernel<2><<<LG_GRID(total)>>>(x, lo, hi, out, total, dims);
  return check_launch("rowwise_kernel");
}
int lg_scale_transform(const float* x, const float* lo, const float* hi, float* out, int64_t n, int32_t d, void* s) { return rowwise(0, x, lo, hi, out, n, d, s); }
int lg_unscale_transform(const float* x, const float* lo, const float* hi, float* out, int64_t n, int32_t d, void* s) { return rowwise(1, x, lo, hi, out, n, d, s); }
int lg_saturate(const float* x, const float* lo, const float* hi, float* out, int64_t n, int32_t d, void* s) { return rowwise(2, x, lo, hi, out, n, d, s); }
int lg_lgsk_kernel(const float* x, float scale, float* out, int64_t n, void* stream) {
  if (!x || !out || n < 0) return fail(LG_ERR_BAD_ARG, "bad argument");
  if (n) lg::lgsk_kernel<<<LG_GRID(n)>>>(x, scale, out, n);
  return check_launch("lgsk_kernel");
}
int lg_cube_keypoints(const float* pose, float cube_size, float* out, int64_t n, void* stream) {
  if (!pose || !out || n < 0) return fail(LG_ERR_BAD_ARG, "bad argument");
  if (n) lg::keypoints_kernel<<<LG_GRID(n * 8)>>>(pose, cube_size, out, n);
  return check_launch("keypoints_kernel");
}

int lg_integrate_goal(const LgParams* P, const LgSimState* S, float dt, void* stream) {
  if (!P || !S || !S->root_state || P->num_envs < 0) return fail(LG_ERR_BAD_ARG, "bad argument");
  const int64_t n = P->num_envs;
  if (n) lg::integrate_rows_kernel<<<LG_GRID(n)>>>(S->root_state + (size_t)P->goal_slot * 13, (int64_t)P->actors_per_env * 13, dt, n);
  return check_launch("integrate_rows_kernel");
}

int lg_selftest_division(float span, float rcp, unsigned long long* mismatches_dev, void* stream) {
  if (!mismatches_dev) return fail(LG_ERR_BAD_ARG, "null argument");
  lg::selftest_division_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(span, rcp, mismatches_dev);
  return check_launch("selftest_division_kernel");
}

int lg_upload_sim_state(const LgParams* P, const LgSimState* S, const LgHostStep* H, void* stream) {
  if (!P || !S || !H || !H->dof_state_host || !H->root_state_host || !H->rigid_body_host)
    return fail(LG_ERR_BAD_ARG, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = P->num_envs;
  const cudaMemcpyKind k = cudaMemcpyHostToDevice;
  cudaMemcpyAsync(S->dof_state, H->dof_state_host, sizeof(float) * N * 18, k, st);
  // Of root_state and rigid_body only the rows the path reads: the object actor's row and the three fingertip
  // bodies, each one strided 2-D copy of 52-byte rows.  Measured on a B200 (PCIe 5 x16, scripts/pcie_probe.cu,
  // 16384 envs): object rows 39 us against 65 us for the whole tensor; the three fingertip copies 143 us against
  // 199 us for the run of bodies 6..16 and 311 us for the whole tensor.
  const size_t row = sizeof(float) * 13;
  cudaMemcpy2DAsync(S->root_state + (size_t)P->object_slot * 13, row * P->actors_per_env,
                    H->root_state_host + (size_t)P->object_slot * 13, row * P->actors_per_env, row, (size_t)N, k, st);
  if (P->goal_rotation)  // moving goal: the pose the simulator integrated for the goal body is read back (:1279-1284)
    cudaMemcpy2DAsync(S->root_state + (size_t)P->goal_slot * 13, row * P->actors_per_env,
                      H->root_state_host + (size_t)P->goal_slot * 13, row * P->actors_per_env, sizeof(float) * 7, (size_t)N, k, st);
  const size_t pitch = row * P->bodies_per_env;
  for (int i = 0; i < 3; ++i) {
    const size_t off = (size_t)P->fingertip_body[i] * 13;
    cudaMemcpy2DAsync(const_cast<float*>(S->rigid_body) + off, pitch, H->rigid_body_host + off, pitch, row, (size_t)N, k, st);
  }
  if (P->asymmetric_obs) {
    if (!H->dof_force_host || !H->ft_sensors_host) return fail(LG_ERR_BAD_ARG, "null host buffer (asymmetric)");
    cudaMemcpyAsync(const_cast<float*>(S->dof_force), H->dof_force_host, sizeof(float) * N * 9, k, st);
    cudaMemcpyAsync(const_cast<float*>(S->ft_sensors), H->ft_sensors_host, sizeof(float) * N * 18, k, st);
  }
  return check_launch("lg_upload_sim_state");
}

int lg_step_host_pipelined(const LgParams* P, const LgSimState* S, const LgBuffers* B, const LgHostStep* H,
                           double sched_step, int chunks, void* stream_main, void* stream_up, void* stream_down) {
  if (int rc = validate(P, S, B, true)) return rc;
  if (!H || !H->dof_state_host || !H->root_state_host || !H->rigid_body_host || !H->action_host || !H->action_staging ||
      !H->reward_host || (!H->obs_host && !(P->asymmetric_obs && H->states_host && !P->dr_activate)))
    return fail(LG_ERR_BAD_ARG, "null host buffer");
  if (chunks < 1 || chunks > 16) return fail(LG_ERR_BAD_ARG, "chunks must be in 1..16");
  if (chunks == 1) return lg_step_host(P, S, B, H, sched_step, stream_main);   // nothing to overlap: one stream, no events
  if (P->asymmetric_obs && (!H->dof_force_host || !H->ft_sensors_host || !H->states_host))
    return fail(LG_ERR_BAD_ARG, "null host buffer (asymmetric)");
  cudaStream_t main = (cudaStream_t)stream_main, up = (cudaStream_t)stream_up, down = (cudaStream_t)stream_down;
  // events are created once per host thread and reused (timing disabled: cheapest kind)
  struct Pool { cudaEvent_t pre, done, up[16], post[16]; bool ok = false; };
  static thread_local Pool pool;
  if (!pool.ok) {
    const unsigned f = cudaEventDisableTiming;
    cudaEventCreateWithFlags(&pool.pre, f); cudaEventCreateWithFlags(&pool.done, f);
    for (int i = 0; i < 16; ++i) { cudaEventCreateWithFlags(&pool.up[i], f); cudaEventCreateWithFlags(&pool.post[i], f); }
    pool.ok = true;
  }
  const int64_t N = P->num_envs;
  const int A = P->action_dim, od = 32 + A, sd = P->asymmetric_obs ? od + 72 : 0;
  const float* obs_src = B->obs_clipped ? B->obs_clipped : B->obs;
  const float* st_src = B->states_clipped ? B->states_clipped : B->states;
  // the action arrives first; resets / torque consume it together with the previous state
  cudaMemcpyAsync(H->action_staging, H->action_host, sizeof(float) * N * A, cudaMemcpyHostToDevice, main);
  if (int rc = lg_pre_physics(P, S, B, H->action_staging, stream_main)) return rc;
  cudaEventRecord(pool.pre, main);
  cudaStreamWaitEvent(up, pool.pre, 0);
  int64_t lo = 0;
  for (int c = 0; c < chunks; ++c) {
    int64_t hi = c == chunks - 1 ? N : ((N * (c + 1) / chunks + 16) / 32) * 32;
    if (hi > N) hi = N;
    if (hi <= lo) continue;
    LgParams pc = *P;
    pc.num_envs = hi - lo; pc.env_offset = P->env_offset + lo; pc.stats_num_envs = N; pc.fuse_bookkeeping = 1;
    LgSimState sc = *S;
    sc.dof_state += lo * 18; sc.root_state += lo * 13 * P->actors_per_env; sc.rigid_body += lo * 13 * P->bodies_per_env;
    if (sc.dof_force) sc.dof_force += lo * 9;
    if (sc.ft_sensors) sc.ft_sensors += lo * 18;
    LgHostStep hc = *H;
    hc.dof_state_host += lo * 18; hc.root_state_host += lo * 13 * P->actors_per_env;
    hc.rigid_body_host += lo * 13 * P->bodies_per_env;
    if (hc.dof_force_host) hc.dof_force_host += lo * 9;
    if (hc.ft_sensors_host) hc.ft_sensors_host += lo * 18;
    LgBuffers bc = *B;
    bc.obs += lo * od; if (bc.states) bc.states += lo * sd;
    if (bc.obs_clipped) bc.obs_clipped += lo * od;
    if (bc.states_clipped) bc.states_clipped += lo * sd;
    bc.action += lo * A; bc.reward += lo; bc.reset += lo; bc.goal_reset += lo; bc.successes += lo;
    if (bc.dones) bc.dones += lo;
    bc.steps_count += lo; bc.goal_pose += lo * 7; bc.goal_movement += lo * 6; bc.history += lo * LG_HISTORY_COLS;
    if (bc.applied_torque) bc.applied_torque += lo * 9;
    bc.term_rewards = nullptr;
    if (int rc = lg_upload_sim_state(&pc, &sc, &hc, stream_up)) return rc;
    cudaEventRecord(pool.up[c], up);
    cudaStreamWaitEvent(main, pool.up[c], 0);
    if (int rc = lg_post_physics(&pc, &sc, &bc, sched_step, stream_main)) return rc;
    cudaEventRecord(pool.post[c], main);
    cudaStreamWaitEvent(down, pool.post[c], 0);
    if (H->obs_host) cudaMemcpyAsync(H->obs_host + lo * od, obs_src + lo * od, sizeof(float) * (hi - lo) * od, cudaMemcpyDeviceToHost, down);
    if (sd) cudaMemcpyAsync(H->states_host + lo * sd, st_src + lo * sd, sizeof(float) * (hi - lo) * sd, cudaMemcpyDeviceToHost, down);
    cudaMemcpyAsync(H->reward_host + lo, B->reward + lo, sizeof(float) * (hi - lo), cudaMemcpyDeviceToHost, down);
    if (H->dones_host && B->dones) cudaMemcpyAsync(H->dones_host + lo, B->dones + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, down);
    lo = hi;
  }
  cudaEventRecord(pool.done, down);
  cudaStreamWaitEvent(main, pool.done, 0);
  return check_launch("lg_step_host_pipelined");
}

int lg_step_host(const LgParams* P, const LgSimState* S, const LgBuffers* B, const LgHostStep* H, double sched_step, void* stream) {
  if (int rc = validate(P, S, B, true)) return rc;
  if (!H || !H->dof_state_host || !H->root_state_host || !H->rigid_body_host || !H->action_host || !H->action_staging ||
      !H->reward_host || (!H->obs_host && !(P->asymmetric_obs && H->states_host && !P->dr_activate)))
    return fail(LG_ERR_BAD_ARG, "null host buffer");
  if (P->asymmetric_obs && (!H->dof_force_host || !H->ft_sensors_host || !H->states_host))
    return fail(LG_ERR_BAD_ARG, "null host buffer (asymmetric)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = P->num_envs;
  const int obs_dim = 32 + P->action_dim, state_dim = obs_dim + 72;
  // the action arrives first: the pre-physics pass (resets, torque) consumes it together with last step's state
  cudaMemcpyAsync(H->action_staging, H->action_host, sizeof(float) * N * P->action_dim, cudaMemcpyHostToDevice, st);
  if (int rc = lg_pre_physics(P, S, B, H->action_staging, stream)) return rc;
  // "physics": the simulator's new state lands in the device tensors (only the rows the path reads)
  if (int rc = lg_upload_sim_state(P, S, H, stream)) return rc;
  if (int rc = lg_post_physics(P, S, B, sched_step, stream)) return rc;
  const float* obs_src = B->obs_clipped ? B->obs_clipped : B->obs;
  if (H->obs_host) cudaMemcpyAsync(H->obs_host, obs_src, sizeof(float) * N * obs_dim, cudaMemcpyDeviceToHost, st);
  if (P->asymmetric_obs) {
    const float* st_src = B->states_clipped ? B->states_clipped : B->states;
    cudaMemcpyAsync(H->states_host, st_src, sizeof(float) * N * state_dim, cudaMemcpyDeviceToHost, st);
  }
  cudaMemcpyAsync(H->reward_host, B->reward, sizeof(float) * N, cudaMemcpyDeviceToHost, st);
  if (H->dones_host && B->dones) cudaMemcpyAsync(H->dones_host, B->dones, N, cudaMemcpyDeviceToHost, st);
  return check_launch("lg_step_host");
}

}  // extern "C"
