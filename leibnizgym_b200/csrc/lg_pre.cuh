// pre-physics pass of the TriFinger MDP hot path (sm_100a): ordered mask compaction (decoupled look-back) fused with
// reset / goal-reset sampling and scatter, action store and action -> torque; tile slabs by TMA bulk copy.
// Also the stand-alone compaction and the hook kernels on explicit id lists.
// Included by lg_kernels.cu (single translation unit).  Reference paths are relative to /root/reference/leibnizgym/.
#pragma once

#include "lg_device.cuh"
#include "lg_post.cuh"   // pdl_wait / pdl_launch_dependents, compute_coefs

namespace lg {

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
  LG_TP(1, 0, tid == 0);
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
  LG_TP(1, 1, tid == 0);
  pdl_wait();   // the flags, counters and statistics below are results of the preceding post-physics pass
  LG_TP(1, 2, tid == 0);
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
  LG_TP(1, 3, tid == 0);
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

  LG_TP(1, 4, tid == 0);
  if (full_tile) mbar_wait(&s_mbar, 0);   // the slabs have landed (each thread only touches its own env row below)
  LG_TP(1, 5, tid == 0);

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
  LG_TP(1, 6, tid == 0);
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
  LG_TP(1, 7, tid == 0);
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
  LG_TP(1, 8, tid == 0);
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

  LG_TP(1, 9, tid == 0);
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
  LG_TP(1, 10, tid == 0);
  if (full_tile && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  LG_TP(1, 11, tid == 0);
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

}  // namespace lg
