/*
 * cc4_b200.h — C-ABI of the B200-native batched CybORG CC4 step engine.
 *
 * The reference (cage-challenge/cage-challenge-4, pure Python) has no FFI seam; its boundary for this path is
 * the Python call chain  BlueFlatWrapper.step -> CybORG.parallel_step -> SimulationController.step
 * (CybORG/Agents/Wrappers/BlueFlatWrapper.py:80, CybORG/env.py:95, CybORG/Simulator/SimulationController.py:211).
 * The entry points below are what a ctypes binding placed behind those methods calls (INTEGRATION.md shows it).
 *
 * Conventions: extern "C", plain pointers and sizes, no torch / C++ types.  Every function returns 0 on
 * success and a negative cc4_status on failure; cc4_last_error() returns a thread-local message.  Pointers
 * named *_dev are device pointers on the handle's device (borrowed for the duration of the call; work is
 * enqueued on `stream`, which is a cudaStream_t passed as void*; NULL = the legacy default stream).
 * Pointers named *_host are host pointers.  A handle is bound to one device and is not thread-safe.
 */
#ifndef CC4_B200_H
#define CC4_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cc4_handle cc4_handle;

enum cc4_status { CC4_OK = 0, CC4_ERR_ARG = -1, CC4_ERR_CUDA = -2, CC4_ERR_STATE = -3 };

/* per-episode sticky error bits (cc4_error_flags) */
enum {
    CC4_E_TAPE_SITE = 1,   /* tape replay: the engine's draw does not match the recorded site / range */
    CC4_E_TAPE_END = 2,    /* tape replay: an actor's segment ran out */
    CC4_E_CAPACITY = 4,    /* a fixed-size table overflowed (the episode's results are no longer exact) */
    CC4_E_UNREACHABLE = 8, /* a branch on which the reference itself raises an exception was reached */
    CC4_E_TAPE_LEFT = 16   /* tape replay: an actor's segment was not fully consumed */
};

#define CC4_N_BLUE 5
#define CC4_OBS_LEN 210   /* BlueFlatWrapper.py:128-170: 1 + 3*59 + 32 (agents 0-3 use the first 92 = 1 + 59 + 32) */
#define CC4_MASK_LEN 242  /* BlueFixedActionWrapper.py:233-309: 82 actions (agents 0-3) / 242 (agent 4) */
#define CC4_N_ACTORS 91   /* tape actor ids: blue 0-4, red 5-10, green 11-90 */

/* step flags */
#define CC4_STEP_OBS 1      /* write the 5 observation vectors            (BlueFlatWrapper.observation_change :172) */
#define CC4_STEP_REWARD 2   /* write the team reward                      (BlueRewardMachine.calculate_reward :70) */
#define CC4_STEP_ALL 3
#define CC4_STEP_OBS_COMPACT 4   /* with CC4_STEP_OBS: obs rows are the five UNPADDED vectors back to back, uint8 [n_envs][578] =
                                    92 + 92 + 92 + 92 + 210 (what BlueFlatWrapper returns without pad_spaces, BlueFlatWrapper.py:128-170);
                                    45 % fewer bytes to move to the host */
#define CC4_OBS_COMPACT_LEN 578
#define CC4_STEP_OBS_BITS 8      /* with CC4_STEP_OBS: obs rows are uint32 [n_envs][20]: bit j of words 0-18 = entry j of the compact row
                                    (every entry but the five mission-phase entries 0, 92, 184, 276, 368 is 0 / 1; those five bits are 0),
                                    word 19 = the mission phase.  80 B per episode instead of 578: what cc4_step_host moves over PCIe
                                    before it expands the rows on the host; a device-side policy can consume the rows as they are */
#define CC4_OBS_BITS_WORDS 20

/* One recorded RNG draw of the reference (tape mode; layout shared with oracle/cc4_oracle.h). */
typedef struct {
    uint64_t val;   /* kind 0: integer in [0,n); kind 1: IEEE-754 bits of a double in [0,1) */
    uint16_t n;
    uint8_t site;
    uint8_t kind;
    uint32_t pad;
} cc4_draw;

/* Replaces: CybORG(scenario_generator, seed) + SimulationController.__init__ (env.py:53-77, SimulationController.py:60-151).
 * n_envs independent episodes; episode i uses the RNG stream (seed, first_episode + i), so results do not depend
 * on how episodes are sharded over devices.  steps = EnterpriseScenarioGenerator(steps=...), 500 in the challenge.
 * rng_mode 0: Philox4x32-10 counter RNG (production); 1: replay of recorded reference draws (parity tests). */
int cc4_create(int n_envs, int device, uint64_t seed, uint64_t first_episode, int steps, int rng_mode, cc4_handle **out);
int cc4_destroy(cc4_handle *h);

/* Re-key the handle: the next cc4_reset generates episodes (seed, first_episode + i).  This is how consecutive resets
 * produce fresh scenarios, as consecutive CybORG.reset() calls do by continuing one RNG (env.py:218-236). */
int cc4_set_episode_base(cc4_handle *h, uint64_t seed, uint64_t first_episode);

/* Scripted red agent class of every episode (EnterpriseScenarioGenerator(red_agent_class=...)): FiniteStateRedAgent
 * (CybORG/Agents/SimpleAgents/FiniteStateRedAgent.py, the challenge default) or its DiscoveryFSRed variant
 * (FSMRedVariants.py:72-122: host-state priority list, server bias, discovery-weighted action matrix), or the
 * RandomSelectRedAgent (RandomSelectRedAgent.py:8-139: a uniformly drawn valid command with uniformly drawn parameters every
 * tick; the only scripted agent that uses Withdraw, Withdraw.py:38-91).  Call before cc4_reset; it takes effect from the
 * next tick on. */
enum { CC4_RED_FINITE_STATE = 0, CC4_RED_DISCOVERY = 1, CC4_RED_RANDOM_SELECT = 2 };
int cc4_set_red_agent(cc4_handle *h, int variant);

/* Packed per-episode state sizes in bytes: hot = the record streamed through shared memory every tick (read + written),
 * cold = dynamic state touched in place by rare transitions only, stat = written at reset (read on Restore / export). */
int cc4_state_bytes(cc4_handle *h, size_t *hot_bytes, size_t *cold_bytes, size_t *stat_bytes);

/* Tape mode only: recs_dev[n_recs]; seg_off_dev[n_envs][n_segs + 1] absolute offsets into recs_dev, segment 0 =
 * reset, segment 1 + tick*CC4_N_ACTORS + actor = that actor's draws in that tick (oracle/ref_harness.py). */
int cc4_set_tape(cc4_handle *h, const cc4_draw *recs_dev, const uint32_t *seg_off_dev, uint32_t n_segs);

/* Replaces: CybORG.reset -> SimulationController.reset -> EnterpriseScenarioGenerator.create_scenario + State.__init__
 * (env.py:218, SimulationController.py:153-209, EnterpriseScenarioGenerator.py:123, State.py:66) and the wrappers'
 * reset outputs (BlueFlatWrapper.reset :61, BlueFixedActionWrapper.reset :76).
 * obs_dev: uint8 [n_envs][5][210] or NULL, 16-byte aligned (CC4_ERR_ARG otherwise: rows go out as 128-bit stores);
 * mask_dev: uint8 [n_envs][5][242] or NULL. */
int cc4_reset(cc4_handle *h, uint8_t *obs_dev, uint8_t *mask_dev, void *stream);

/* Same call with HOST buffers (copies the outputs device->host and synchronises the stream). */
int cc4_reset_host(cc4_handle *h, uint8_t *obs_host, uint8_t *mask_host, void *stream);

/* Replaces: BlueFlatWrapper.step -> CybORG.parallel_step -> SimulationController.step (+ send_messages,
 * observation_change x5, BlueRewardMachine) for every episode of the batch.
 * actions_dev int32 [n_envs][5] BlueFixedActionWrapper indices; msgs_dev uint8 [n_envs][5] (bit i = message[i]) or NULL;
 * obs_dev uint8 [n_envs][5][210] (or [n_envs][578] with CC4_STEP_OBS_COMPACT); reward_dev float [n_envs] (team reward = BlueRewardMachine + action_cost, identical
 * for the 5 blue agents, SimulationController.py:1118); done_dev uint8 [n_envs].  Outputs may be NULL.
 * Alignment (checked, CC4_ERR_ARG): obs_dev 16 bytes (4 with CC4_STEP_OBS_BITS), reward_dev and actions_dev 4 bytes. */
int cc4_step(cc4_handle *h, const int32_t *actions_dev, const uint8_t *msgs_dev, uint8_t *obs_dev, float *reward_dev,
             uint8_t *done_dev, int flags, void *stream);

/* Same call with HOST buffers: copies actions/messages host->device, steps, copies obs/reward/done back and
 * synchronises the stream (the e2e path bench.py times; host buffers should be pinned for full speed).  The batch is
 * cut into chunks that are pipelined over two streams; observation rows cross PCIe as bit rows (CC4_STEP_OBS_BITS) and
 * are expanded to the requested uint8 layout by a small pool of host threads (CC4_HOST_THREADS, default: the host's
 * hardware threads / LOCAL_WORLD_SIZE, at most 8) while the next chunks are in flight.  Environment knobs:
 * CC4_HOST_STREAM=0/1 forces ordinary / streaming stores for the expanded rows (default: streaming when the node's ranks
 * together write more than 64 MB per tick); CC4_HOST_EXPAND=0 copies byte rows instead of bit rows.  With
 * CC4_STEP_OBS_BITS in `flags` the bit rows themselves are returned (obs_host = uint32 [n][20]). */
int cc4_step_host(cc4_handle *h, const int32_t *actions_host, const uint8_t *msgs_host, uint8_t *obs_host,
                  float *reward_host, uint8_t *done_host, int flags, void *stream);

/* Host-only: expands bit rows uint32 [n_rows][20] (CC4_STEP_OBS_BITS) to uint8 [n_rows][578] (compact != 0) or
 * [n_rows][5][210]; no device work.  For callers that fetch the bit rows themselves. */
int cc4_expand_obs_bits(const uint32_t *rows_host, uint8_t *obs_host, size_t n_rows, int compact);

/* Host-only measurement: GB/s that `threads` host threads sustain WRITING a `bytes`-sized buffer `iters` times (streaming
 * != 0: non-temporal stores, as cc4_step_host uses when the node's ranks together write more than 64 MB per tick).  This
 * is the bound of the host-buffer path on several GPUs (bench.py: e2e.host_write_gbs). */
int cc4_host_write_bandwidth(size_t bytes, int threads, int streaming, int iters, double *gbs_out);

/* n_ticks fused ticks with the synthetic uniform-random blue policy of the benchmark (indices drawn on device from
 * Philox keyed (policy_seed, episode, tick); BASELINE config "FiniteStateRed + random blue").  reward_sum_dev float
 * [n_envs] is incremented by each tick's reward (may be NULL).  State is kept in shared memory across the ticks. */
int cc4_rollout(cc4_handle *h, int n_ticks, uint64_t policy_seed, float *reward_sum_dev, int flags, void *stream);

/* Per-episode error bits: out_host uint32 [n_envs] (synchronises). */
int cc4_error_flags(cc4_handle *h, uint32_t *out_host);

/* Canonical flattened state of one episode (layout: oracle/canon.py) for parity tests: out_host int32 [cap];
 * *n_out = number of values.  Synchronises the device. */
int cc4_export_canon(cc4_handle *h, int env, int32_t *out_host, int cap, int *n_out);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int cc4_launch_count(cc4_handle *h, uint64_t *out);

/* Profiling builds only (compiled with -DCC4_PHASE_CLOCKS; otherwise CC4_ERR_STATE): accumulated busy / barrier-wait
 * cycles of every phase of the tick per episode, out_host uint64 [n_envs][18] ([2p] busy cycles of phase p; [2p+1] of a
 * block's first episode: wall cycles of phase p for the block, barrier included); reset != 0 clears the counters.
 * The first call arms the counters.  Synchronises the device. */
int cc4_phase_clocks(cc4_handle *h, unsigned long long *out_host, int reset);
/* Profiling builds only: out_host uint64 [64][2] = cycles spent in, and number of, executed actions per action code
 * (oracle/canon.py ACTION_CODE) since cc4_phase_clocks armed the counters. */
int cc4_action_clocks(cc4_handle *h, unsigned long long *out_host);

const char *cc4_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
