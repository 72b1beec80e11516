/*
 * cc4_oracle.h — CPU restatement of the CybORG CC4 multi-agent step (TEST INFRASTRUCTURE).
 *
 * This is the ORACLE: a plain, serial, single-episode re-statement of what the reference Python simulator
 * does on the hot path (SimulationController.step and everything it calls).  It exists to CHECK the CUDA
 * engine and to provide the CPU baseline in bench.py.  Nothing in the product package may include, link
 * or call it (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg).
 *
 * Parity status: PINNED against the live reference — tests/golden/ holds RNG tapes + per-step canonical
 * state / observation / reward trajectories produced by running the unmodified reference
 * (/root/reference, commit 8c3c50c) through oracle/ref_harness.py; tests/test_oracle_golden.py replays
 * them through this file and requires bit-exact equality after every step.
 *
 * All citations are file:line under /root/reference/CybORG/.
 */
#ifndef CC4_ORACLE_H
#define CC4_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- draw sites (must match oracle/ref_harness.py SITE and cc4_b200/csrc/cc4_common.cuh) ---- */
enum {
    ST_R_NUSER = 1, ST_R_NSERVER = 2, ST_R_NADDON = 3, ST_R_ADDON = 4, ST_R_PID = 5, ST_R_BLUE_PARENT = 6,
    ST_R_RED_SUBNET = 7, ST_R_RED_HOST = 8, ST_CREATE_PID = 9, ST_GREEN_ACT = 10, ST_LW_SERVICE = 11,
    ST_LW_RELIAB = 12, ST_LW_FP = 13, ST_LW_PHISH = 14, ST_PHISH_AGENT = 15, ST_AS_DEST = 16, ST_AS_FP = 17,
    ST_FSM_HOST = 18, ST_FSM_ACTION = 19, ST_FSM_PARAM = 20, ST_PORTSCAN = 21, ST_EXP_OTHER = 22,
    ST_EXP_TOP = 23, ST_EXP_SUCC = 24, ST_EXP_DETECT = 25, ST_SSH_HOP = 26, ST_PRIV_SESS = 27, ST_DD_TP = 28,
    ST_DD_FP = 29, ST_RSC_PRIMARY = 30, ST_DECOY_FACTORY = 31,
    /* DiscoveryFSRed only (FiniteStateRedAgent._choose_host :252-293) */
    ST_FSMD_STATE_P = 32, ST_FSMD_STATE_U = 33, ST_FSMD_ROLL = 34, ST_FSMD_HOST_SRV = 35, ST_FSMD_HOST_OTHER = 36,
    ST_FSMD_HOST_SRV2 = 37, ST_FSMD_HOST_ALL = 38,
    /* RandomSelectRedAgent only (RandomSelectRedAgent.get_action :37-92) */
    ST_RS_COMMAND = 39, ST_RS_PARAM = 40
};

/* One recorded draw. kind 0: integer in [0,n) stored in val; kind 1: double in [0,1), IEEE bits in val. */
typedef struct {
    uint64_t val;
    uint16_t n;
    uint8_t site;
    uint8_t kind;
    uint32_t pad;
} cc4o_draw;

#define CC4O_N_ACTORS 91 /* blue 0-4, red 5-10, green 11-90 */
#define CC4O_OBS_LEN 210
#define CC4O_MASK_LEN 242
#define CC4O_EXPORT_MAX 32768

/* error flags (sticky, per episode) */
enum {
    CC4O_E_TAPE_SITE = 1,      /* tape replay: next record's site/n differs from the engine's draw */
    CC4O_E_TAPE_END = 2,       /* tape replay: segment exhausted */
    CC4O_E_CAPACITY = 4,       /* a fixed-size table overflowed */
    CC4O_E_UNREACHABLE = 8,    /* a branch on which the reference itself raises was reached */
    CC4O_E_TAPE_LEFT = 16      /* tape replay: a segment had draws the engine did not consume */
};

typedef struct cc4o_env cc4o_env;

/* rng_mode 0 = Philox4x32-10 keyed (seed, episode_id); 1 = tape replay */
cc4o_env *cc4o_create(uint64_t seed, uint64_t episode_id, int steps, int rng_mode);
/* scripted red agent class: 0 = FiniteStateRedAgent (default), 1 = DiscoveryFSRed (FSMRedVariants.py:72-122),
 * 2 = RandomSelectRedAgent (RandomSelectRedAgent.py:8-139, the agent that uses Withdraw); before cc4o_reset */
void cc4o_set_red_agent(cc4o_env *e, int variant);
void cc4o_destroy(cc4o_env *e);
/* tape: recs[] plus seg_off[n_segs + 1]; segment 0 = reset, 1 + t*CC4O_N_ACTORS + actor = tick t.
 * The arrays are borrowed (must outlive the env). */
void cc4o_set_tape(cc4o_env *e, const cc4o_draw *recs, const uint32_t *seg_off, uint32_t n_segs);
int cc4o_reset(cc4o_env *e);
/* actions[5]: BlueFixedActionWrapper indices; msgs[5]: one byte per blue agent (bit i = message[i]).
 * obs: [5][210] uint8, reward: team reward (BlueRewardMachine + action_cost), done. returns 0 ok, <0 error */
int cc4o_step(cc4o_env *e, const int32_t *actions, const uint8_t *msgs, uint8_t *obs, float *reward, uint8_t *done);
int cc4o_observe(cc4o_env *e, uint8_t *obs);       /* obs vectors of the current state (used after reset) */
int cc4o_action_mask(cc4o_env *e, uint8_t *mask);  /* [5][242] */
/* canonical flattened state (layout: oracle/canon.py); returns number of int32 written */
int cc4o_export(cc4o_env *e, int32_t *out, int cap);
int cc4o_error_flags(cc4o_env *e);
void cc4o_reward_breakdown(cc4o_env *e, int32_t *out);   /* [3 phases][LWF, ASF, RIA][9 subnets] event counts since reset */
const char *cc4o_last_error(cc4o_env *e);
/* bench helper: run `n_eps` whole episodes (reset + steps-1 ticks) with uniform-random blue indices drawn from
 * a Philox stream keyed like the CUDA engine's synthetic blue policy; returns total env ticks executed and the
 * summed reward via *ret_sum */
long cc4o_run_episodes(uint64_t seed, uint64_t first_episode, int n_eps, int steps, double *ret_sum);

#ifdef __cplusplus
}
#endif
#endif
