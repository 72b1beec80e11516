/*
 * cc4_device.cuh — device-side CC4 tick logic on the packed per-episode record (cc4_state.h).
 *
 * One warp owns one episode; the record lives in shared memory for the duration of the kernel.  The functions in
 * this file fall in two groups:
 *   - lane-parallel phases (names start with w_): record load/store, green agents (lane per agent), end-turn Monitor
 *     (lane per host), observation vectors; red policy + FSM digest (lane per red agent);
 *   - the order-dependent transitions (blue actions, red kill chain, phishing, reassignment, RedSessionCheck) that the
 *     reference executes strictly in agent order.  These run WARP-UNIFORM: all 32 lanes of the episode's warp execute
 *     the same control flow on the same values (scalar stores write the same value from every lane), and every scan
 *     over a list (session pool, process list, known ports, suspicious pids, observation entries) is done 32 entries
 *     at a time with ballots / reductions.  Functions marked [uniform] must be called by the whole converged warp with
 *     identical arguments; functions marked [lane] may be called by any subset of lanes.
 * Citations are file:line under /root/reference/CybORG/ (never read at run time).
 */
/* no include guard: cc4_engine.cu includes this file once per namespace (CC4_NS, CC4_WITH_TAPE) */
#ifndef CC4_NS
#define CC4_NS cc4
#endif
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cc4_b200.h"
#include "cc4_state.h"

namespace CC4_NS {

/* inlining knobs (tools/build_variants.py): which helper groups are compiled inline; defaults = the fastest measured set
 * (profiles/r01x_variants.txt: everything inline except the Philox block, the tape reader and rng_int / rng_f64: 0.62 -> 0.54 ms) */
#ifndef CC4_RNG_INLINE
#define CC4_RNG_INLINE __forceinline__     /* one call level less on every draw: 1.7 % (profiles/r01x_variants.txt) */
#endif
#ifndef CC4_INL4
#define CC4_INL4 __forceinline__
#endif
#ifndef CC4_INL5
#define CC4_INL5 __forceinline__
#endif
#ifndef CC4_INL6
#define CC4_INL6 __forceinline__
#endif
#ifndef CC4_INL7
#define CC4_INL7 __forceinline__
#endif
#ifndef CC4_INL8
#define CC4_INL8 __forceinline__
#endif
#ifndef CC4_INL9
#define CC4_INL9 __noinline__
#endif
#ifndef CC4_INL2
#define CC4_INL2 __noinline__
#endif
#ifndef CC4_INL3
#define CC4_INL3 __forceinline__
#endif
#ifndef CC4_POLICY_PHILOX_INLINE
#define CC4_POLICY_PHILOX_INLINE 0
#endif
#ifndef CC4_GREEN_PHILOX_INLINE
#define CC4_GREEN_PHILOX_INLINE 1     /* the green agents keep their own inlined Philox rounds: their service-word load stays in flight across them (-1.4 %) */
#endif
#ifndef CC4_INL_RARE
#define CC4_INL_RARE CC4_INL6      /* transitions that most ticks do not execute: reassignment, re-election, new server sessions, phishing */
#endif

/* ------------------------------------------------------------------------------------------------ */
/* constants                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
enum { S_RZA, S_OZA, S_RZB, S_OZB, S_CON, S_PAZ, S_ADM, S_OFF, S_INT };   /* EnterpriseScenarioGenerator.py:40-51 */
enum { A_BLUE0 = 0, A_RED0 = 5, A_GREEN0 = 11, NAG = CC4_N_ACTORS };

/* draw sites (oracle/ref_harness.py SITE) */
enum {
    ST_R_NUSER = 1, ST_R_NSERVER, ST_R_NADDON, ST_R_ADDON, ST_R_PID, ST_R_BLUE_PARENT, ST_R_RED_SUBNET, ST_R_RED_HOST,
    ST_CREATE_PID, ST_GREEN_ACT, ST_LW_SERVICE, ST_LW_RELIAB, ST_LW_FP, ST_LW_PHISH, ST_PHISH_AGENT, ST_AS_DEST, ST_AS_FP,
    ST_FSM_HOST, ST_FSM_ACTION, ST_FSM_PARAM, ST_PORTSCAN, ST_EXP_OTHER, ST_EXP_TOP, ST_EXP_SUCC, ST_EXP_DETECT, ST_SSH_HOP,
    ST_PRIV_SESS, ST_DD_TP, ST_DD_FP, ST_RSC_PRIMARY, ST_DECOY_FACTORY,
    /* DiscoveryFSRed only (FiniteStateRedAgent._choose_host :252-293) */
    ST_FSMD_STATE_P, ST_FSMD_STATE_U, ST_FSMD_ROLL, ST_FSMD_HOST_SRV, ST_FSMD_HOST_OTHER, ST_FSMD_HOST_SRV2, ST_FSMD_HOST_ALL,
    /* RandomSelectRedAgent only (RandomSelectRedAgent.get_action :37-92) */
    ST_RS_COMMAND, ST_RS_PARAM
};
static_assert(ST_DECOY_FACTORY == 31 && ST_FSMD_HOST_ALL == 38 && ST_RS_PARAM == 40, "draw sites must match oracle/cc4_oracle.h");

/* action codes (oracle/canon.py ACTION_CODE) */
enum { AC_SLEEP = 0, AC_MONITOR, AC_ANALYSE, AC_REMOVE, AC_RESTORE, AC_DECOY, AC_BLOCK, AC_ALLOW, AC_INVALID,
       AC_DRS = 10, AC_AGGR, AC_STEALTH, AC_DD, AC_EXPLOIT, AC_PRIVESC, AC_IMPACT, AC_DEGRADE, AC_WITHDRAW,
       AC_G_ACCESS = 20, AC_G_LOCAL = 21, AC_RSC = 30 };

enum { F_K, F_KD, F_S, F_SD, F_U, F_UD, F_R, F_RD, F_F };                  /* FiniteStateRedAgent.py:401-411 */
enum { SUCC_TRUE = 1, SUCC_UNKNOWN = 2, SUCC_FALSE = 3, SUCC_INPROG = 4 };  /* Shared/Enums.py:5 */

__constant__ int8_t c_sorted_subnets[9] = { S_ADM, S_CON, S_INT, S_OFF, S_OZA, S_OZB, S_PAZ, S_RZA, S_RZB }; /* BlueFlatWrapper.py:195 */
__constant__ int8_t c_fsm_col_code[9] = { AC_DRS, AC_AGGR, AC_STEALTH, AC_DD, AC_EXPLOIT, AC_PRIVESC, AC_IMPACT, AC_DEGRADE, AC_WITHDRAW };
__constant__ int8_t c_fsm_succ[9][9] = {   /* FiniteStateRedAgent.py:441-451 */
    { F_KD, F_S, F_S, -1, -1, -1, -1, -1, -1 },   { F_KD, F_SD, F_SD, -1, -1, -1, -1, -1, -1 },
    { F_SD, -1, -1, F_S, F_U, -1, -1, -1, -1 },   { F_SD, -1, -1, F_SD, F_UD, -1, -1, -1, -1 },
    { F_UD, -1, -1, -1, -1, F_R, -1, -1, F_S },   { F_UD, -1, -1, -1, -1, F_RD, -1, -1, F_SD },
    { F_RD, -1, -1, -1, -1, -1, F_R, F_R, F_S },  { F_RD, -1, -1, -1, -1, -1, F_RD, F_RD, F_SD },
    { F_F, -1, -1, -1, -1, -1, -1, -1, -1 } };
__constant__ int8_t c_fsm_fail[9][9] = {   /* :481-491 */
    { F_K, F_K, F_K, -1, -1, -1, -1, -1, -1 },    { F_KD, F_KD, F_KD, -1, -1, -1, -1, -1, -1 },
    { F_S, -1, -1, F_S, F_S, -1, -1, -1, -1 },    { F_SD, -1, -1, F_SD, F_SD, -1, -1, -1, -1 },
    { F_U, -1, -1, -1, -1, F_U, -1, -1, F_U },    { F_UD, -1, -1, -1, -1, F_UD, -1, -1, F_UD },
    { F_R, -1, -1, -1, -1, -1, F_R, F_R, F_R },   { F_RD, -1, -1, -1, -1, -1, F_RD, F_RD, F_RD },
    { F_F, -1, -1, -1, -1, -1, -1, -1, -1 } };
/* :540-549 probabilities in quarters (-1 = None); all are multiples of 0.25, so cdf / total is exact in double */
__constant__ int8_t c_fsm_prob4[2][8][9] = { {
    { 2, 1, 1, -1, -1, -1, -1, -1, -1 },  { -1, 2, 2, -1, -1, -1, -1, -1, -1 },
    { 1, -1, -1, 1, 2, -1, -1, -1, -1 },  { -1, -1, -1, 1, 3, -1, -1, -1, -1 },
    { 2, -1, -1, -1, -1, 2, -1, -1, 0 },  { -1, -1, -1, -1, -1, 4, -1, -1, 0 },
    { 2, -1, -1, -1, -1, -1, 1, 1, 0 },   { -1, -1, -1, -1, -1, -1, 2, 2, 0 } },
  { /* DiscoveryFSRed.state_transitions_probability FSMRedVariants.py:110-120 */
    { 1, 3, 0, -1, -1, -1, -1, -1, -1 },  { -1, 4, 0, -1, -1, -1, -1, -1, -1 },
    { 1, -1, -1, 0, 3, -1, -1, -1, -1 },  { -1, -1, -1, 0, 4, -1, -1, -1, -1 },
    { 0, -1, -1, -1, -1, 4, -1, -1, 0 },  { -1, -1, -1, -1, -1, 4, -1, -1, 0 },
    { 4, -1, -1, -1, -1, -1, 0, 0, 0 },   { -1, -1, -1, -1, -1, -1, 2, 2, 0 } } };
/* DiscoveryFSRed.set_host_state_priority_list FSMRedVariants.py:86-99 (K KD S SD U UD R RD) */
__constant__ uint8_t c_fsmd_priority[8] = { 20, 20, 20, 20, 10, 10, 0, 0 };
__constant__ int8_t c_red_aspace_cols[9] = { 0, 1, 2, 4, 5, 7, 3, 6, 8 };  /* EnterpriseScenarioGenerator.py:764-768 */
__constant__ int8_t c_reward[3][9][3] = {  /* BlueRewardMachine.py:35-65 [phase][subnet][LWF, ASF, RIA] */
    { { -1, -3, -1 }, { -1, -1, -1 }, { -1, -3, -1 }, { -1, -1, -1 }, { 0, -5, -5 }, { -1, -1, -3 }, { -1, -1, -3 }, { -1, -1, -3 }, { 0, 0, -1 } },
    { { -2, -1, -3 }, { -10, 0, -10 }, { -1, -1, -1 }, { -1, -1, -1 }, { 0, 0, 0 }, { -1, -1, -3 }, { -1, -1, -3 }, { -1, -1, -3 }, { 0, 0, 0 } },
    { { -1, -3, -3 }, { -1, -1, -1 }, { -2, -1, -3 }, { -10, 0, -10 }, { 0, 0, 0 }, { -1, -1, -3 }, { -1, -1, -3 }, { -1, -1, -3 }, { 0, 0, 0 } } };
/* green allowed-subnet masks [phase][subnet] (EnterpriseScenarioGenerator.py:281-306 + own subnet) and
 * comms-policy adjacency (BlueFlatWrapper.py:267-302); generated by tools/gen_tables.py from the same pair lists
 * as oracle/cc4_oracle.c init_tables() */
__constant__ uint16_t c_green_allowed[3][9] = {
    { 0x0F7, 0x003, 0x0FD, 0x00C, 0x0F5, 0x035, 0x055, 0x095, 0x100 },
    { 0x0E1, 0x002, 0x0FC, 0x00C, 0x0F4, 0x035, 0x055, 0x095, 0x100 },
    { 0x0F3, 0x003, 0x0E4, 0x008, 0x0F1, 0x035, 0x055, 0x095, 0x100 } };
__constant__ uint16_t c_comms_adj[3][9] = {
    { 0x1F6, 0x001, 0x1F9, 0x004, 0x1E5, 0x1D5, 0x1B5, 0x175, 0x0F5 },
    { 0x0E0, 0x000, 0x1F8, 0x004, 0x1E4, 0x1D5, 0x1B5, 0x175, 0x0F4 },
    { 0x1F2, 0x001, 0x0E0, 0x000, 0x1E1, 0x1D5, 0x1B5, 0x175, 0x0F1 } };
__constant__ int8_t c_tree_parent[9] = { S_INT, S_RZA, S_INT, S_RZB, S_INT, S_INT, S_PAZ, S_PAZ, -1 };   /* :388-416 */
__constant__ uint16_t c_links[9] = {   /* server_host_0 knowledge links :431-462 */
    (1 << S_OZA) | (1 << S_CON), (1 << S_RZA), (1 << S_OZB) | (1 << S_CON), (1 << S_RZB),
    (1 << S_RZA) | (1 << S_RZB) | (1 << S_PAZ), (1 << S_ADM) | (1 << S_OFF) | (1 << S_CON), (1 << S_PAZ), (1 << S_PAZ), 0 };
__constant__ int8_t c_kind_portbit[12] = { 0, 1, 2, 3, 4, 2, 5, 4, 2, -1, -1, -1 };   /* canon PORT_BIT */
__constant__ uint8_t c_duration[32] = { 1, 1, 2, 3, 5, 2, 1, 1, 1, 1, 1, 1, 3, 2, 4, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 };
#define PB_22 1
#define PB_80 4
#define PB_3390 8
#define PB_25 16
#define PB_443 32

struct Act { int code, host, subnet, session; };
/* actions travel between functions as one packed word (IP_MAKE layout, cc4_state.h) so that they stay in registers */
#define ACT_VALID (1u << 31)   /* blue_decode: the index is a valid (mask True) action */
__device__ __forceinline__ Act unpack_act(uint32_t w)
{
    Act a = { (int)IP_CODE(w), IP_HOST(w) == 255 ? -1 : (int)IP_HOST(w), IP_SUBNET(w) == 15 ? -1 : (int)IP_SUBNET(w), (int)IP_SESSION(w) };
    return a;
}
__device__ __forceinline__ uint32_t pack_act(const Act &a, int ticks) { return IP_MAKE(a.code, a.host < 0 ? 255 : a.host, a.subnet < 0 ? 15 : a.subnet, a.session, ticks); }
#define ACT_SLEEP_WORD IP_MAKE(AC_SLEEP, 255, 15, 0, 0)

/* per-warp scratch (shared memory, not persisted) */
struct Scratch {
    uint16_t ctr_br[12];                         /* draw counter / tape position of this tick: blue 0-4, red 5-10, [11] = reset stream */
    uint8_t ctr_g[CC4_MAXGREEN];                 /* ... of the green agents (a green agent draws at most 7 times per tick) */
    uint16_t oent[CC4_NRED][CC4_OENT_CAP];       /* explicit observation entries: slot | key_ip<<8 | iface<<9 | subnet_if<<10 | syshost<<11 | sess<<12 */
    uint8_t n_oent[8];
    uint8_t have_first[8], succ0[8], has_act0[8];
    uint32_t act0[8];                             /* IP_MAKE(code, host, subnet, session, 0) */
    uint32_t evpid[8]; int n_evpid;               /* (slot<<16 | pid) carried by this tick's process_creation events: at most one per red agent */
    uint32_t subw[11];                            /* this tick's due action per blue / red actor (IP_MAKE, ticks = 0) */
    uint8_t dropped[16];
    int16_t act_si[8];                            /* pool index of the session each red agent's due action names, resolved lane-parallel before the red phase */
    int sess_moved;                               /* a session was deleted since act_si was resolved (indices may have shifted) */
    /* session-pool digests, rebuilt by w_scan_sessions */
    uint32_t sess_hosts[CC4_NRED][4];             /* hosts on which agent r holds a session */
    int n_red[8];                                 /* sessions per agent */
    uint32_t has0, newss, misplaced;              /* bit r: agent r has ident 0 / got a new abstract session this tick (set by sess_add) / has a session on a foreign subnet */
    int reward, cost;                             /* BlueRewardMachine sum / action cost of the tick */
    uint32_t active;                              /* bit r: red agent r holds a session after this tick's reassignment */
    uint32_t blk_id[CC4_NRED + 1];                /* cached Philox block per serial stream: red agents 0..5, [6] = reset */
    uint32_t blk[CC4_NRED + 1][4];
    union {                                       /* two phase-local areas that are never live together */
        struct { uint32_t hostagents[32];         /* byte h = mask of red agents with a session on host h (built on demand for phishing) */
                 uint32_t phish[3], pad;          /* green agents whose local work triggered a phishing email this tick */
                 uint64_t as_cum[8]; } g;         /* per source subnet: byte sn = number of allowed destination servers in subnets <= sn (GreenAccessService) */
        uint32_t obsbits[40];                     /* observation vectors as one bit string (w_write_obs) */
        struct { uint8_t agent[32], ident[32]; } ra;     /* the misplaced sessions different_subnet_agent_reassignment is moving (reassign) */
    };
};

/* per-warp context (shared memory): everything a device function needs besides the record and the scratch */
struct WarpCtx {
    Cc4Stat *st; Cc4Cold *k;
    const uint32_t *seg;      /* tape mode: this episode's segment offsets */
    int tick;                 /* state step_count at the start of the tick; -1 during reset */
    uint32_t k1, ep_lo; int env;
};
/* what is the same for every episode of the block (set by the kernel prologue) */
struct BlockCtx {
    const cc4_draw *tape; uint32_t n_segs;
    int steps, rng_mode;
    uint32_t k0;
    int red_variant;          /* 0 FiniteStateRedAgent, 1 DiscoveryFSRed, 2 RandomSelectRedAgent */
};
__shared__ BlockCtx g_blk;
/* tape mode (parity replay of the reference's recorded draws) or production mode (Philox); CC4_WITH_TAPE=0 builds a
   library without the tape paths (experiment: what the mode checks and the tape code cost the production kernel) */
/* The scripted red agent class is a COMPILE-TIME parameter of the kernels (template <int RV>: 0 FiniteStateRedAgent, 1 DiscoveryFSRed,
   2 RandomSelectRedAgent; the host picks the instantiation from cc4_set_red_agent).  As a run-time switch the two non-default
   agents cost the default one 3 % (profiles/r03g_red_variant_hooks.txt: every out-of-line call site inside the policy and red
   phases changes the register allocation of the whole phase), as a template parameter nothing. */
#ifndef CC4_WITH_TAPE
#define CC4_WITH_TAPE 1
#endif
#define CC4_TAPE_ON() (CC4_WITH_TAPE && g_blk.rng_mode == 1)
/* per-actor draw counters: 16 bits for the blue / red / reset streams, 8 bits for the green ones */
__device__ __forceinline__ uint32_t ctr_get(const struct Scratch *x, int actor);
__device__ __forceinline__ void ctr_set(struct Scratch *x, int actor, uint32_t v);
struct __align__(16) WarpSmem {
    Cc4Dyn dyn;
    Scratch x;
    WarpCtx c;
};
/* A warp's WarpSmem lives in dynamic shared memory; device functions receive its byte offset (`Cx`, one register) and
 * address it through the __shared__ symbol, so every access is a shared-memory instruction and nothing about the
 * context is passed through the thread stack (ncu r01a: the stack-resident context cost 24 % long-scoreboard stalls). */
extern __shared__ __align__(16) unsigned char smem_raw[];
typedef uint32_t Cx;
__device__ __forceinline__ uint32_t ctr_get(const Scratch *x, int actor) { return actor < A_GREEN0 ? x->ctr_br[actor] : actor == NAG ? x->ctr_br[11] : x->ctr_g[actor - A_GREEN0]; }
__device__ __forceinline__ void ctr_set(Scratch *x, int actor, uint32_t v)
{
    if (actor < A_GREEN0) x->ctr_br[actor] = (uint16_t)v; else if (actor == NAG) x->ctr_br[11] = (uint16_t)v; else x->ctr_g[actor - A_GREEN0] = (uint8_t)v;
}
/* host of green agent g (agents are numbered over the user hosts in creation order) */
__device__ __forceinline__ int green_host_of(const Cc4Dyn *s, int g)
{
    const uint64_t cum = s->green_cum;
    const uint32_t kk = (uint32_t)g * 0x01010101u;
    const int sn = (__popc(__vcmpleu4((uint32_t)cum, kk)) + __popc(__vcmpleu4((uint32_t)(cum >> 32), kk))) >> 3;
    return sn * 16 + 6 + (g - (sn ? (int)((cum >> (8 * (sn - 1))) & 0xFFu) : 0));
}
#define WS(cx_) (reinterpret_cast<WarpSmem *>(smem_raw + (cx_)))
#define DYN(cx_) (&WS(cx_)->dyn)
#define SCR(cx_) (&WS(cx_)->x)
#define WC(cx_) (&WS(cx_)->c)

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void fail(Cx c, uint32_t flag) { atomicOr(&DYN(c)->err, flag); }

/* ------------------------------------------------------------------------------------------------ */
/* RNG.  Production: every (actor, tick) owns a stream of 32-bit words; word j is element j&3 of the    */
/* Philox4x32-10 block with key (seed, episode) and counter (j>>2, actor, tick+1, episode).  A draw     */
/* consumes one word: integer = high half of word*n, double = word / 2^32 (oracle/cc4_oracle.c rng_*). */
/* Parity: replay of the reference's recorded draws.                                                  */
/* ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ CC4_INL9 uint64_t tape_next(Cx c, int actor, int site, int kind, uint32_t n)
{
    uint32_t seg = actor == NAG ? 0u : 1u + (uint32_t)WC(c)->tick * NAG + (uint32_t)actor;
    if (seg >= g_blk.n_segs) { fail(c, CC4_E_TAPE_END); return 0; }
    uint32_t i = WC(c)->seg[seg] + ctr_get(SCR(c), actor);
    if (i >= WC(c)->seg[seg + 1]) { fail(c, CC4_E_TAPE_END); return 0; }
    cc4_draw d = g_blk.tape[i];
    if (d.site != site || d.kind != kind || d.n != n) fail(c, CC4_E_TAPE_SITE);
    ctr_set(SCR(c), actor, ctr_get(SCR(c), actor) + 1);
    return d.val;
}
/* block `blk` of the actor's stream for this tick (registers only) */
#ifndef CC4_INL10
#define CC4_INL10 __noinline__      /* ONE copy of the ten Philox rounds instead of five inlined ones: the kernel is bound by instruction
                                       supply (its executed code does not fit the SM's instruction cache), -4.3 % (profiles/r04_variants.txt) */
#endif
__device__ CC4_INL10 uint4 philox_block(Cx c, int actor, uint32_t blk)
{
    const WarpCtx *g = WC(c);
    uint32_t o[4] = { blk, (uint32_t)actor, actor == NAG ? 0u : (uint32_t)g->tick + 1u, g->ep_lo };
    philox4x32_10(o, g_blk.k0, g->k1);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
/* the same, always inline: for the green agents, whose service-word load stays in flight across the rounds (a call would wait for it) */
__device__ __forceinline__ uint4 philox_block_inl(Cx c, int actor, uint32_t blk)
{
    const WarpCtx *g = WC(c);
    uint32_t o[4] = { blk, (uint32_t)actor, actor == NAG ? 0u : (uint32_t)g->tick + 1u, g->ep_lo };
    philox4x32_10(o, g_blk.k0, g->k1);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ uint32_t u4_word(const uint4 &v, uint32_t i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
/* next word of the actor's stream.  Blue / red actors and the reset stream (serial callers) keep their current block in
 * the scratch; green actors (lane-parallel in green_tick, which keeps its block in registers) recompute on the rare
 * serial follow-ups (phishing). */
__device__ CC4_RNG_INLINE uint32_t rng_word(Cx c, int actor)
{
    Scratch *x = SCR(c);
    const uint32_t j = ctr_get(x, actor);
    ctr_set(x, actor, j + 1);
    int slot = actor == NAG ? CC4_NRED : (actor >= A_RED0 && actor < A_GREEN0) ? actor - A_RED0 : -1;   /* blue / green draws are rare: uncached */
    if (slot < 0) { uint4 v = philox_block(c, actor, j >> 2); return u4_word(v, j & 3u); }
    if (x->blk_id[slot] != (j >> 2)) {
        uint4 v = philox_block(c, actor, j >> 2);
        x->blk_id[slot] = j >> 2; x->blk[slot][0] = v.x; x->blk[slot][1] = v.y; x->blk[slot][2] = v.z; x->blk[slot][3] = v.w;
    }
    return x->blk[slot][j & 3u];
}
__device__ __forceinline__ double f64_from_word(uint32_t w) { return (double)w * 2.3283064365386962890625e-10; }     /* w / 2^32, exact */
__device__ CC4_INL2 uint32_t rng_int(Cx c, int actor, int site, uint32_t n)
{
    if (CC4_TAPE_ON()) { uint64_t v = tape_next(c, actor, site, 0, n); return v < n ? (uint32_t)v : 0u; }
    if (n == 1 && site == ST_FSM_PARAM) return 0u;      /* a one-option parameter pick spends no word (oracle: rng_int) */
    return __umulhi(rng_word(c, actor), n);
}
__device__ CC4_INL2 double rng_f64(Cx c, int actor, int site, uint32_t n)
{
    if (CC4_TAPE_ON()) return __longlong_as_double((long long)tape_next(c, actor, site, 1, n));
    return f64_from_word(rng_word(c, actor));
}

/* The red agents' serial draws (exec_red, re-election) in production mode: the stream position and the cached block of
 * agent r are in the scratch, the fast path is a handful of shared-memory instructions inline, the Philox refill (once per
 * four words; the first block comes from red_policy) is out of line.  Same words as rng_int / rng_f64 would deliver. */
__device__ __noinline__ void red_refill(Cx c, int r, uint32_t b)
{
    Scratch *x = SCR(c);
    const uint4 v = philox_block(c, A_RED0 + r, b);
    /* every calling lane holds the same block: one of them files it, the others read it back after the warp barrier */
    const uint32_t am = __activemask();
    if (lane_id() == __ffs(am) - 1) { x->blk_id[r] = b; x->blk[r][0] = v.x; x->blk[r][1] = v.y; x->blk[r][2] = v.z; x->blk[r][3] = v.w; }
    __syncwarp(am);
}
__device__ __forceinline__ uint32_t red_word(Cx c, int r)
{
    Scratch *x = SCR(c);
    const uint32_t j = x->ctr_br[A_RED0 + r];
    x->ctr_br[A_RED0 + r] = (uint16_t)(j + 1);
    if (x->blk_id[r] != (j >> 2)) red_refill(c, r, j >> 2);
    return x->blk[r][j & 3u];
}
#ifndef CC4_RED_RNG_INLINE
#define CC4_RED_RNG_INLINE 1        /* 0: the red agents' serial draws go through the generic out-of-line rng_int / rng_f64 (tuning knob) */
#endif
__device__ __forceinline__ uint32_t rng_int_red(Cx c, int r, int site, uint32_t n)
{
    if (!CC4_RED_RNG_INLINE || CC4_TAPE_ON()) return rng_int(c, A_RED0 + r, site, n);
    return __umulhi(red_word(c, r), n);
}
__device__ __forceinline__ double rng_f64_red(Cx c, int r, int site)
{
    if (!CC4_RED_RNG_INLINE || CC4_TAPE_ON()) return rng_f64(c, A_RED0 + r, site, 0);
    return f64_from_word(red_word(c, r));
}

/* the reset stream (scenario generation, one serial stream per episode): the same fast path, cache slot CC4_NRED */
__device__ __noinline__ void reset_refill(Cx c, uint32_t b)
{
    Scratch *x = SCR(c);
    const uint4 v = philox_block(c, NAG, b);
    x->blk_id[CC4_NRED] = b; x->blk[CC4_NRED][0] = v.x; x->blk[CC4_NRED][1] = v.y; x->blk[CC4_NRED][2] = v.z; x->blk[CC4_NRED][3] = v.w;
}
__device__ __forceinline__ uint32_t rng_int_reset(Cx c, int site, uint32_t n)
{
    if (CC4_TAPE_ON()) return rng_int(c, NAG, site, n);
    Scratch *x = SCR(c);
    const uint32_t j = x->ctr_br[11];
    x->ctr_br[11] = (uint16_t)(j + 1);
    if (x->blk_id[CC4_NRED] != (j >> 2)) reset_refill(c, j >> 2);
    return __umulhi(x->blk[CC4_NRED][j & 3u], n);
}

/* ------------------------------------------------------------------------------------------------ */
/* small helpers                                                                                    */
/* ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int slot_subnet(int slot) { return slot == CC4_SLOT_INTERNET ? S_INT : slot >= CC4_SLOT_ROUTER ? slot - CC4_SLOT_ROUTER : slot >> 4; }
__device__ __forceinline__ bool host_exists(const Cc4Dyn *s, int slot)
{
    if (slot >= CC4_SLOT_ROUTER) return true;
    int sn = slot >> 4, k = slot & 15;
    return k < 6 ? k < s->nserver[sn] : (k - 6) < s->nuser[sn];
}
__device__ __forceinline__ int red_of_subnet(int sn) { return sn == S_CON ? 0 : sn <= S_OZB ? sn + 1 : sn <= S_OFF ? 5 : -1; }
__device__ __forceinline__ bool red_allowed(int r, int sn) { return red_of_subnet(sn) == r; }
__device__ __forceinline__ int blue_of_subnet(int sn) { return sn <= S_OZB ? sn : (sn >= S_PAZ && sn <= S_OFF) ? 4 : -1; }
__device__ __forceinline__ bool bit128(const uint32_t *b, int i) { return (b[i >> 5] >> (i & 31)) & 1u; }
__device__ __forceinline__ void set128(uint32_t *b, int i) { b[i >> 5] |= 1u << (i & 31); }

/* ---- process lists [uniform] (they live in the cold record's arena, cc4_state.h) ---- */
#define FULLMASK 0xFFFFFFFFu
#define NPROC(c, h) ((int)PD_CNT(WC(c)->k->pdir[h]))
#define PROCS(c, h) (WC(c)->k->parena + PD_OFF(WC(c)->k->pdir[h]))
#define P_NONE P_DEADBIT    /* filler for lanes past the end of a list: looks like a tombstone */
/* port class of a process kind (canon PORT_BIT), 15 = does not listen */
__device__ __forceinline__ uint32_t kind_portbit(uint32_t kind) { return (uint32_t)(0xFFF245243210ull >> (4 * kind)) & 15u; }
__device__ CC4_INL4 void proc_append(Cx c, int h, uint32_t w, uint32_t d, uint32_t pmax)
{   /* host.processes.append; d = pdir[h], pmax = pmax[h] (the caller has them in registers) */
    Cc4Cold *k = WC(c)->k;
    uint32_t off = PD_OFF(d), n = PD_CNT(d), capl = PD_CAPL(d);
    if (n >= PD_MAXCNT) { fail(c, CC4_E_CAPACITY); return; }        /* the directory's count field holds 1023 (reachable only with steps >> 500) */
    if (n == (1u << capl)) {            /* block full: move the list to a block of twice the size */
        uint32_t top = k->arena_top, ncap = 2u << capl;
        if (top + ncap > CC4_ARENA) { fail(c, CC4_E_CAPACITY); return; }
        _Pragma("unroll 1") for (uint32_t i = lane_id(); i < n; i += 32) k->parena[top + i] = k->parena[off + i];
        __syncwarp();
        k->arena_top = top + ncap; off = top; capl++;
    }
    k->parena[off + n] = w;
    k->pdir[h] = PD_MAKE(off, n + 1, capl);
    if (P_PID(w) > pmax) k->pmax[h] = (uint16_t)P_PID(w);
}
__device__ CC4_INL4 void proc_append(Cx c, int h, uint32_t w) { proc_append(c, h, w, WC(c)->k->pdir[h], WC(c)->k->pmax[h]); }
__device__ CC4_INL5 void proc_rescan_max(Cx c, int h)
{
    Cc4Cold *k = WC(c)->k;
    const uint32_t d = k->pdir[h];
    const uint32_t *pp = k->parena + PD_OFF(d); const int n = (int)PD_CNT(d);
    uint32_t mx = 0;
    _Pragma("unroll 1") for (int i = lane_id(); i < n; i += 32) { uint32_t w = pp[i]; if (!P_DEAD(w) && P_PID(w) > mx) mx = P_PID(w); }
    mx = __reduce_max_sync(FULLMASK, mx);
    k->pmax[h] = (uint16_t)mx;
}
__device__ CC4_INL5 void proc_delete(Cx c, int h, int idx)
{
    Cc4Cold *k = WC(c)->k;
    const uint32_t d = k->pdir[h];
    uint32_t *pp = k->parena + PD_OFF(d);
    const int n = (int)PD_CNT(d);
    const uint32_t gone = pp[idx];
    _Pragma("unroll 1") for (int base = idx; base + 1 < n; base += 32) {
        const int i = base + lane_id();
        const bool a = i + 1 < n;
        uint32_t v = 0;
        if (a) v = pp[i + 1];
        __syncwarp();
        if (a) pp[i] = v;
        __syncwarp();
    }
    k->pdir[h] = PD_MAKE(PD_OFF(d), n - 1, PD_CAPL(d));
    if (!P_DEAD(gone) && P_PID(gone) == k->pmax[h]) proc_rescan_max(c, h);
}
__device__ CC4_INL5 int proc_find(Cx c, uint32_t pid, uint32_t d)
{   /* Host.get_process Host.py:251 (tombstones are not processes); d = pdir[host] */
    const uint32_t *pp = WC(c)->k->parena + PD_OFF(d); const int n = (int)PD_CNT(d);
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const uint32_t w = i < n ? pp[i] : P_NONE;
        const uint32_t m = __ballot_sync(FULLMASK, !P_DEAD(w) && P_PID(w) == pid);
        if (m) return base + __ffs(m) - 1;
    }
    return -1;
}
__device__ CC4_INL5 int proc_find(Cx c, int h, uint32_t pid) { return proc_find(c, pid, WC(c)->k->pdir[h]); }
__device__ CC4_INL3 int create_pid(Cx c, int actor, uint32_t pmax)
{   /* Host.create_pid Host.py:198-200; pmax = pmax[h] */
    uint32_t pid = pmax + 1 + rng_int(c, actor, ST_CREATE_PID, 9);
    if (pid > 0xFFFFu) { fail(c, CC4_E_CAPACITY); pid = 0xFFFFu; }
    return (int)pid;
}
__device__ CC4_INL3 int create_pid_red(Cx c, int r, uint32_t pmax)
{
    uint32_t pid = pmax + 1 + rng_int_red(c, r, ST_CREATE_PID, 9);
    if (pid > 0xFFFFu) { fail(c, CC4_E_CAPACITY); pid = 0xFFFFu; }
    return (int)pid;
}
__device__ CC4_INL3 int create_pid_on(Cx c, int actor, int h) { return create_pid(c, actor, (uint32_t)WC(c)->k->pmax[h]); }

/* ---- services ---- */
__device__ __forceinline__ int svc_count(uint64_t v)
{   /* entries are filled from 0 and never removed, so the empty ones (code 15) are a suffix: 9 - their number */
    const uint64_t t = v & (v >> 1) & (v >> 2) & (v >> 3) & 0x0102040810204081ull;      /* bit 7i set iff the low nibble of entry i is 15 */
    return 9 - __popcll(t);
}
__device__ __forceinline__ bool svc_active(uint64_t v, uint32_t entry) { return !((entry & 15u) == K_OT && (v & SVC_OT_STOPPED)); }
__device__ __forceinline__ uint64_t svc_set(uint64_t v, int i, uint32_t entry) { return (v & ~(0x7Full << (7 * i))) | ((uint64_t)entry << (7 * i)); }

/* ---- red session pool [uniform] ---- */
/* first pool index whose (word, flags) satisfies pred, -1 if none */
template <class Pred> __device__ __forceinline__ int sess_first(const Cc4Dyn *s, Pred pred)
{
    const int n = s->n_sess;
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const uint32_t m = __ballot_sync(FULLMASK, i < n && pred(s->sess[i], (uint32_t)s->sessf[i]));
        if (m) return base + __ffs(m) - 1;
    }
    return -1;
}
/* position of the k-th (0-based) set bit of a warp-uniform mask with more than k bits set [uniform]: every lane tests its own bit */
__device__ __forceinline__ int kth_bit(uint32_t m, int k)
{
    const int lane = lane_id();
    return __ffs(__ballot_sync(FULLMASK, (m >> lane & 1u) && __popc(m & ((1u << lane) - 1u)) == k)) - 1;
}
/* the same for one lane on its own mask (red policy): five popcount steps instead of __fns (65 instructions with branches) */
__device__ __forceinline__ int nth_bit(uint32_t m, int n)
{
    int pos = 0, c;
    c = __popc(m & 0xFFFFu); if (n >= c) { n -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (n >= c) { n -= c; pos += 8;  m >>= 8; }
    c = __popc(m & 0xFu);    if (n >= c) { n -= c; pos += 4;  m >>= 4; }
    c = __popc(m & 0x3u);    if (n >= c) { n -= c; pos += 2;  m >>= 2; }
    c = (int)(m & 1u);       if (n >= c) pos += 1;
    return pos;
}
/* pool index of the k-th (0-based) entry that satisfies pred, -1 if there are fewer */
template <class Pred> __device__ __forceinline__ int sess_kth(const Cc4Dyn *s, int k, Pred pred)
{
    const int n = s->n_sess;
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const uint32_t m = __ballot_sync(FULLMASK, i < n && pred(s->sess[i], (uint32_t)s->sessf[i]));
        const int cnt = __popc(m);
        if (k < cnt) return base + kth_bit(m, k);
        k -= cnt;
    }
    return -1;
}
__device__ CC4_INL5 int sess_find(const Cc4Dyn *s, int r, int id)
{
    return sess_first(s, [=](uint32_t w, uint32_t f) { return SF_AGENT(f) == (uint32_t)r && S_ID(w) == (uint32_t)id; });
}
__device__ CC4_INL5 void ports_drop(Cx c, int r, int id)
{   /* stable compaction of the known-port entries */
    Cc4Dyn *s = DYN(c);
    uint32_t *ports = WC(c)->k->ports;
    const int n = s->n_ports;
    int w = 0;
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const uint32_t p = i < n ? ports[i] : 0u;
        const bool keep = i < n && !(PT_AGENT(p) == (uint32_t)r && PT_ID(p) == (uint32_t)id);
        const uint32_t m = __ballot_sync(FULLMASK, keep);
        if (keep) ports[w + __popc(m & ((1u << lane_id()) - 1u))] = p;      /* target index <= i: never ahead of an unread entry */
        w += __popc(m);
        __syncwarp();
    }
    s->n_ports = (uint16_t)w;
}
/* remove pool entries [idx] by shifting the tail down (keeps the order) */
__device__ __forceinline__ void sess_shift_down(Cc4Dyn *s, int idx, int n)
{
    _Pragma("unroll 1") for (int base = idx; base + 1 < n; base += 32) {
        const int i = base + lane_id();
        const bool a = i + 1 < n;
        uint32_t v = 0; uint8_t f = 0;
        if (a) { v = s->sess[i + 1]; f = s->sessf[i + 1]; }
        __syncwarp();
        if (a) { s->sess[i] = v; s->sessf[i] = f; }
        __syncwarp();
    }
}
__device__ CC4_INL5 void sess_delete(Cx c, int idx)
{
    Cc4Dyn *s = DYN(c);
    if (SF_ABS(s->sessf[idx])) ports_drop(c, SF_AGENT(s->sessf[idx]), S_ID(s->sess[idx]));
    const int n = s->n_sess;
    sess_shift_down(s, idx, n);
    s->n_sess = (uint16_t)(n - 1);
    SCR(c)->sess_moved = 1;
}
/* State.add_session State.py:305-323 for a red agent; returns the ident */
__device__ CC4_INL4 int sess_add(Cx c, int r, int host, int abstract, int root, int pid, int has_parent = 0)
{
    Cc4Dyn *s = DYN(c);
    const int n = s->n_sess;
    if (n >= CC4_SESS_CAP) { fail(c, CC4_E_CAPACITY); return 0; }
    int mx = -1;
    _Pragma("unroll 1") for (int i = lane_id(); i < n; i += 32) if (SF_AGENT(s->sessf[i]) == (uint32_t)r) mx = max(mx, (int)S_ID(s->sess[i]));
    int id = __reduce_max_sync(FULLMASK, mx) + 1;       /* max ident + 1, 0 when the agent holds no session */
    if (id > 255) { fail(c, CC4_E_CAPACITY); id = 255; }
    if (s->sess_count[r] < 255) s->sess_count[r]++;
    s->sess[n] = S_MAKE(pid, id, host);
    s->sessf[n] = (uint8_t)(r | abstract << 3 | root << 4 | has_parent << 5);
    s->n_sess = (uint16_t)(n + 1);
    if (abstract) SCR(c)->newss |= 1u << r;     /* its ident becomes a known server_session key with the end-turn observation */
    return id;
}
__device__ CC4_INL5 int session_from_pid(const Cc4Dyn *s, int host, uint32_t pid)
{   /* State.get_session_from_pid State.py:420-444; blue / green session processes are never removed */
    return sess_first(s, [=](uint32_t w, uint32_t) { return S_HOST(w) == (uint32_t)host && S_PID(w) == pid; });
}
__device__ __forceinline__ bool blocking(const Cc4Dyn *s, int src_slot, int other_slot)
{   /* RemoteAction.blocking_host Action.py:127-136 */
    return (s->blocks[slot_subnet(other_slot)] >> slot_subnet(src_slot)) & 1;
}

/* ---- observation bookkeeping for one red agent (ObservationSet / ActionSpace.update ActionSpace.py:170-212) ---- */
#ifndef CC4_INL_OBS
#define CC4_INL_OBS CC4_INL4
#endif
__device__ CC4_INL_OBS void obs_entry(Cx c, int r, int key_ip, int slot, int iface, int subnet_if, int syshost, int sess)
{   /* [uniform] */
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    if (syshost) set128(s->hn_known[r], slot);
    if (iface) set128(s->ip_known[r], slot);
    if (subnet_if) s->sn_known[r] |= (uint16_t)(1u << slot_subnet(slot));
    const uint16_t key = (uint16_t)(slot | key_ip << 8), fl = (uint16_t)(iface << 9 | subnet_if << 10 | syshost << 11 | sess << 12);
    const int n = x->n_oent[r];
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const uint32_t m = __ballot_sync(FULLMASK, i < n && (x->oent[r][i] & 0x1FF) == key);
        if (m) { x->oent[r][base + __ffs(m) - 1] |= fl; return; }
    }
    if (n >= CC4_OENT_CAP) { fail(c, CC4_E_CAPACITY); return; }
    x->oent[r][n] = key | fl;
    x->n_oent[r] = (uint8_t)(n + 1);
}
/* the ident becomes a known server_session key (ActionSpace.update ActionSpace.py:170-212: dict insertion order, no
 * duplicates) */
__device__ CC4_INL4 void know_session(Cx c, int r, int id)
{   /* [uniform] */
    Cc4Dyn *s = DYN(c);
    const int n = s->n_ss[r];
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        if (__any_sync(FULLMASK, i < n && s->ss[r][i] == id)) return;
    }
    if (n >= CC4_SS_CAP) { fail(c, CC4_E_CAPACITY); return; }
    s->ss[r][n] = (uint8_t)id;
    s->n_ss[r] = (uint8_t)(n + 1);
}
/* RedSessionCheck reports every session of the agent (RedSessionCheck.py:57-63): the idents of its abstract sessions become
 * known server_session keys in dict (= pool) order.  Only agents that got a new abstract session this tick can learn
 * something (sess_add sets x->newss).  [uniform]: pool scanned 32 entries at a time */
__device__ CC4_INL_RARE void learn_new_sessions(Cx c)
{
    Cc4Dyn *s = DYN(c);
    const uint32_t news = SCR(c)->newss;
    const int n = s->n_sess;
    _Pragma("unroll 1") for (int r = 0; r < CC4_NRED; r++) {
        if (!(news >> r & 1)) continue;
        _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
            const int i = base + lane_id();
            const uint32_t w = i < n ? s->sess[i] : 0u;
            uint32_t m = __ballot_sync(FULLMASK, i < n && SF_AGENT((uint32_t)s->sessf[i]) == (uint32_t)r && SF_ABS((uint32_t)s->sessf[i]));
            while (m) { const int j = __ffs(m) - 1; m &= m - 1; know_session(c, r, (int)S_ID(__shfl_sync(FULLMASK, w, j))); }
        }
    }
}
__device__ __forceinline__ void obs_push(Cx c, int r, int succ, bool has_act, uint32_t aw)
{
    Scratch *x = SCR(c);
    if (x->have_first[r]) return;
    x->have_first[r] = 1; x->succ0[r] = (uint8_t)succ; x->has_act0[r] = has_act;
    x->act0[r] = aw & 0x03FFFFFFu;
}

/* ------------------------------------------------------------------------------------------------ */
/* FiniteStateRedAgent bookkeeping, applied eagerly at the end of the tick (FiniteStateRedAgent.py:91-93) */
/* ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int fsm_state(const Cc4Fsm *f, int slot)
{
    const int w = slot >> 5, b = slot & 31;
    return (int)((f->pl[0][w] >> b & 1u) | (f->pl[1][w] >> b & 1u) << 1 | (f->pl[2][w] >> b & 1u) << 2 | (f->pl[3][w] >> b & 1u) << 3);
}
__device__ __forceinline__ void fsm_set_state(Cc4Fsm *f, int slot, int st)
{
    const int w = slot >> 5; const uint32_t bit = 1u << (slot & 31);
    _Pragma("unroll") for (int p = 0; p < 4; p++) f->pl[p][w] = (st >> p & 1) ? f->pl[p][w] | bit : f->pl[p][w] & ~bit;
}
/* append host `slot` to agent r's list (state K or U); false when the list is full */
__device__ __forceinline__ bool fsm_append(Cx c, int r, int slot, int hn, int st)
{
    Cc4Dyn *s = DYN(c); Cc4Fsm *f = &s->fsm[r];
    const int n = s->n_fsm[r];
    if (n >= CC4_FSM_CAP) { fail(c, CC4_E_CAPACITY); return false; }
    s->fsm_order[r][n] = (uint8_t)slot; s->n_fsm[r] = (uint8_t)(n + 1);
    const int w = slot >> 5; const uint32_t bit = 1u << (slot & 31);
    f->present[w] |= bit;
    if (hn) f->hn[w] |= bit;
    if (st & 4) f->pl[2][w] |= bit;       /* K = 0, U = 4 */
    return true;
}

/* End of tick for one ACTIVE red agent, run by lane r (agents are independent here) [lane]: the RedSessionCheck
 * observation (RedSessionCheck._all_red_sessions_observation RedSessionCheck.py:57-63: one hostname-keyed entry with
 * Sessions, Interface+Subnet and System info per session host) is represented by the host mask x->sess_hosts[r] instead
 * of explicit entries; ActionSpace.update (ActionSpace.py:170-212) and the FSM bookkeeping consume the mask. */
template <int RV> __device__ CC4_INL8 void agent_end_turn(Cx c, int r)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    Cc4Fsm *f = &s->fsm[r];
    const uint32_t *rsc = x->sess_hosts[r];
    uint32_t snk = s->sn_known[r];
    uint32_t seen[4];
    _Pragma("unroll") for (int w = 0; w < 4; w++) {
        uint32_t m = rsc[w];
        seen[w] = m;
        s->hn_known[r][w] |= m; s->ip_known[r][w] |= m;
        if (m & 0xFFFFu) snk |= 1u << (2 * w);
        if (m >> 16) snk |= 1u << (2 * w + 1);
    }
    s->sn_known[r] = (uint16_t)snk;
    if (RV == 2) return;       /* RandomSelectRedAgent keeps no host states: the action-space update above is all */

    /* _host_state_transition :124-167 on the action the merged observation reports */
    const int started = (s->fsm_started >> r) & 1;
    if (x->has_act0[r] && x->succ0[r] != SUCC_INPROG) {
        const int code = IP_CODE(x->act0[r]), ahost = IP_HOST(x->act0[r]), asub = IP_SUBNET(x->act0[r]);
        const bool succ = x->succ0[r] == SUCC_TRUE;
        const int col = code >= AC_DRS && code <= AC_DEGRADE ? code - AC_DRS : -1;      /* columns follow the action codes; Withdraw (8) never transitions */
        if (col == 0) {
            /* DiscoverRemoteSystems moves every known host of the subnet: success adds the 'D' suffix (K->KD, S->SD, U->UD,
               R->RD; F stays), failure changes nothing (:441-451 / :481-491, column 0) */
            if (succ && asub < 8) {
                const int w = asub >> 1; const uint32_t m = f->present[w] & ~f->pl[3][w] & (0xFFFFu << ((asub & 1) * 16));
                f->pl[0][w] |= m;
            }
        } else if (col > 0 && ahost < 128 && bit128(f->present, ahost) && (col <= 4 || bit128(f->hn, ahost))) {
            /* host-targeted: ip-addressed actions (columns 1-4) find the host by address, the others by hostname */
            const int st = fsm_state(f, ahost);
            int nx = succ ? c_fsm_succ[st][col] : c_fsm_fail[st][col];
            if (nx == F_U && !red_allowed(r, ahost >> 4)) nx = F_F;
            if (nx >= 0 && nx != st) {
                fsm_set_state(f, ahost, nx);
                if (nx == F_F) {        /* F is absorbing: mark the list position for the host pick of the policy */
                    const int n = s->n_fsm[r];
                    _Pragma("unroll 1") for (int i = 0; i < n; i++) if (s->fsm_order[r][i] == ahost) { f->posF[i >> 5] |= 1u << (i & 31); break; }
                }
            }
        }
    }
    /* _process_new_observations :190-250, explicit entries first (they precede the RedSessionCheck observation) */
    _Pragma("unroll 1") for (int i = 0; i < x->n_oent[r]; i++) {
        uint16_t e = x->oent[r][i];
        int slot = e & 0xFF, key_ip = (e >> 8) & 1, syshost = (e >> 11) & 1;
        if ((e & 0x1900) == 0x0100 && bit128(f->present, slot)) continue;      /* ip-keyed, no hostname, no session, host already listed (a repeated scan): nothing to learn */
        int iface = ((e >> 9) & 1) || (!key_ip && bit128(rsc, slot));   /* a hostname-keyed entry merges with the RedSessionCheck entry of the same host */
        int hn = !key_ip || syshost;
        uint32_t w = (uint32_t)slot >> 5, bit = 1u << (slot & 31);
        if ((e >> 12) & 1) {
            if (iface) { _Pragma("unroll") for (uint32_t q = 0; q < 4; q++) seen[q] |= w == q ? bit : 0u; }     /* static indices: seen[] stays in registers */
            else fail(c, CC4_E_UNREACHABLE);
        }
        if (!(key_ip || iface || (f->hn[w] & bit))) { fail(c, CC4_E_UNREACHABLE); continue; }   /* would create host_states[None] */
        if (!(f->present[w] & bit)) fsm_append(c, r, slot, hn, started ? F_K : F_U);
        else if (hn) f->hn[w] |= bit;
    }
    /* RedSessionCheck entries in state.sessions[agent] order: hosts the FSM has not seen yet */
    const uint32_t nw0 = rsc[0] & ~f->present[0], nw1 = rsc[1] & ~f->present[1], nw2 = rsc[2] & ~f->present[2], nw3 = rsc[3] & ~f->present[3];
    const int n_new = __popc(nw0) + __popc(nw1) + __popc(nw2) + __popc(nw3);
    if (n_new == 1) {       /* the usual case: one new session host, no order to respect */
        const int h = nw0 ? __ffs(nw0) - 1 : nw1 ? 31 + __ffs(nw1) : nw2 ? 63 + __ffs(nw2) : 95 + __ffs(nw3);
        fsm_append(c, r, h, 1, started ? F_K : F_U);
    } else if (n_new > 1) {
        _Pragma("unroll 1") for (int i = 0; i < s->n_sess; i++) {
            if (SF_AGENT(s->sessf[i]) != (uint32_t)r) continue;
            uint32_t h = S_HOST(s->sess[i]);
            if (bit128(f->present, h)) continue;
            if (!fsm_append(c, r, h, 1, started ? F_K : F_U)) break;
        }
    }
    _Pragma("unroll") for (int w = 0; w < 4; w++) {
        /* ... the FSM learns the hostname of every session host that has an entry (all of them, unless the list is full) */
        f->hn[w] |= rsc[w] & f->present[w];
        /* _session_removal_state_change :169-188: U / UD / R / RD hosts without a session in this observation fall back to KD */
        const uint32_t m = f->pl[2][w] & ~f->pl[3][w] & ~seen[w];
        f->pl[0][w] |= m; f->pl[1][w] &= ~m; f->pl[2][w] &= ~m;
    }
}

/* action class + parameters for the chosen host (FiniteStateRedAgent._choose_host_and_action :296-336 +
 * replace_action_if_invalid SimulationController.py:1068-1112); ri / rf draw an integer / a double */
template <int RV, class RI, class RF> __device__ __forceinline__ uint32_t fsm_choose_action(Cx c, int r, int slot, RI ri, RF rf)
{
    Cc4Dyn *s = DYN(c);
    const Cc4Fsm *f = &s->fsm[r];
    const int st = fsm_state(f, slot), f_hn = bit128(f->hn, slot);
    /* options in action-space order (c_red_aspace_cols); numpy Generator.choice(p): idx = searchsorted(cumsum(p) / sum, u, side='right') */
    int nc = 0, tot = 0;
    const int8_t (*prob4)[9] = c_fsm_prob4[RV == 1 ? 1 : 0];
    _Pragma("unroll 1") for (int i = 0; i < 9; i++) { int p = prob4[st][c_red_aspace_cols[i]]; if (p >= 0) { tot += p; nc++; } }
    if (tot != 4) fail(c, CC4_E_UNREACHABLE);
    const double u = rf(ST_FSM_ACTION, (uint32_t)nc);
    int col = -1, cum = 0, last = -1;
    _Pragma("unroll 1") for (int i = 0; i < 9 && col < 0; i++) {
        int cc = c_red_aspace_cols[i], p = prob4[st][cc];
        if (p < 0) continue;
        last = cc; cum += p;
        if (!((double)cum * 0.25 <= u)) col = cc;      /* cdf / cdf[-1]: every row of the table sums to 4 quarters, so this is exact */
    }
    if (col < 0) col = last;
    Act a = { c_fsm_col_code[col], -1, -1, 0 };
    #define PICK_SESSION() do { if (s->n_ss[r] == 0) fail(c, CC4_E_UNREACHABLE); else a.session = s->ss[r][ri(ST_FSM_PARAM, s->n_ss[r])]; } while (0)
    if (col == 0) {
        const uint32_t known = s->sn_known[r];
        const int pk = (int)ri(ST_FSM_PARAM, (uint32_t)__popc(known));
        a.subnet = (int)__fns(known, 0, pk + 1);        /* pk-th known subnet in creation order */
        PICK_SESSION(); (void)ri(ST_FSM_PARAM, 1);
    } else if (col <= 3) { PICK_SESSION(); (void)ri(ST_FSM_PARAM, 1); a.host = slot; }
    else { if (col >= 5 && !f_hn) fail(c, CC4_E_UNREACHABLE); a.host = slot; PICK_SESSION(); (void)ri(ST_FSM_PARAM, 1); }
    #undef PICK_SESSION
    if (col >= 1 && col <= 4 && !bit128(s->ip_known[r], a.host)) a.code = AC_INVALID;
    if (col >= 5 && !bit128(s->hn_known[r], a.host)) a.code = AC_INVALID;
    return pack_act(a, 0);
}

/* The same choice in production mode, on the words of the agent's first Philox block (`j` = next unused word, advanced):
 * the action draw u = word / 2^32 is compared as an integer (u < cum / 4  <=>  word < cum * 2^30, both sides exact), the
 * table sanity check is left to the tape path, a one-option parameter pick spends no word (as rng_int). */
template <int RV> __device__ __forceinline__ uint32_t fsm_choose_action_words(Cx c, int r, int slot, const uint4 &blk, uint32_t &j)
{
    Cc4Dyn *s = DYN(c);
    const Cc4Fsm *f = &s->fsm[r];
    const int st = fsm_state(f, slot), f_hn = bit128(f->hn, slot);
    const int8_t (*prob4)[9] = c_fsm_prob4[RV == 1 ? 1 : 0];
    const uint32_t wa = u4_word(blk, j++);
    int col = -1, cum = 0, last = -1;
    _Pragma("unroll 1") for (int i = 0; i < 9 && col < 0; i++) {
        const int cc = c_red_aspace_cols[i], p = prob4[st][cc];
        if (p < 0) continue;
        last = cc; cum += p;
        if (cum >= 4 || wa < ((uint32_t)cum << 30)) col = cc;
    }
    if (col < 0) col = last;
    Act a = { c_fsm_col_code[col], -1, -1, 0 };
    auto pick = [&](uint32_t n) -> uint32_t { return n == 1 ? 0u : __umulhi(u4_word(blk, j++), n); };
    const int nss = s->n_ss[r];
    if (col == 0) {
        const uint32_t known = s->sn_known[r];
        a.subnet = (int)__fns(known, 0, (int)pick((uint32_t)__popc(known)) + 1);        /* pk-th known subnet in creation order */
    } else {
        if (col >= 5 && !f_hn) fail(c, CC4_E_UNREACHABLE);
        a.host = slot;
    }
    if (nss == 0) fail(c, CC4_E_UNREACHABLE); else a.session = s->ss[r][pick((uint32_t)nss)];
    if (col >= 1 && col <= 4 && !bit128(s->ip_known[r], a.host)) a.code = AC_INVALID;
    if (col >= 5 && !bit128(s->hn_known[r], a.host)) a.code = AC_INVALID;
    return pack_act(a, 0);
}

/* DiscoveryFSRed (FSMRedVariants.py:72-122), out of line: draws through the generic stream functions (the same words in
 * the same order as the register-resident path of red_policy would take) */
__device__ __noinline__ uint32_t red_policy_discovery(Cx c, int r, uint32_t lo, uint32_t hi)
{
    Cc4Dyn *s = DYN(c);
    const int actor = A_RED0 + r;
    const Cc4Fsm *f = &s->fsm[r];
    const int nf = s->n_fsm[r];
    auto ri = [&](int site, uint32_t n) -> uint32_t { return rng_int(c, actor, site, n); };
    auto rf = [&](int site, uint32_t n) -> double { return rng_f64(c, actor, site, n); };
    int slot;
    {
        /* DiscoveryFSRed: FiniteStateRedAgent._choose_host :252-293 with a host-state priority list and the server bias.
           First walk over the list: the states in order of first appearance; after the state is drawn, a second walk
           collects the position masks of its hosts / of its servers ('server' in hostname: the entry carries a hostname
           and sits in a server slot). */
        uint32_t st_list = 0, seen = 0; int ns = 0;             /* 4 bits per listed state */
        _Pragma("unroll 1") for (int i = 0; i < nf; i++) {
            if ((i < 32 ? lo >> i : hi >> (i - 32)) & 1u) {
                const int st = fsm_state(f, s->fsm_order[r][i]);
                if (!(seen >> st & 1u)) { seen |= 1u << st; st_list |= (uint32_t)st << (4 * ns); ns++; }
            }
        }
        int total = 0;
        _Pragma("unroll 1") for (int i = 0; i < ns; i++) total += c_fsmd_priority[st_list >> (4 * i) & 15u];
        int pick;
        if (total > 0) {
            /* probs = [(p / base) * p_multiplier]; numpy choice(p=): cdf = cumsum(probs); cdf /= cdf[-1]; searchsorted(cdf, u,
               'right'): the same double operations in the same order, no fused multiply-add */
            const double mult = __ddiv_rn(1.0, __ddiv_rn((double)total, 100.0));
            double last = 0.0;
            _Pragma("unroll 1") for (int i = 0; i < ns; i++) last = __dadd_rn(last, __dmul_rn(__ddiv_rn((double)c_fsmd_priority[st_list >> (4 * i) & 15u], 100.0), mult));
            const double u = rf(ST_FSMD_STATE_P, (uint32_t)ns);
            double acc = 0.0;
            _Pragma("unroll 1") for (pick = 0; pick < ns; pick++) {
                acc = __dadd_rn(acc, __dmul_rn(__ddiv_rn((double)c_fsmd_priority[st_list >> (4 * pick) & 15u], 100.0), mult));
                if (!(__ddiv_rn(acc, last) <= u)) break;
            }
            if (pick >= ns) pick = ns - 1;
        } else pick = (int)ri(ST_FSMD_STATE_U, (uint32_t)ns);
        const int cst = (int)(st_list >> (4 * pick) & 15u);
        uint64_t all = 0, srv = 0;      /* second walk: the position masks of the chosen state only (registers, no local arrays) */
        _Pragma("unroll 1") for (int i = 0; i < nf; i++) {
            if ((i < 32 ? lo >> i : hi >> (i - 32)) & 1u) {
                const int sl_ = s->fsm_order[r][i];
                if (fsm_state(f, sl_) != cst) continue;
                all |= 1ull << i;
                if (bit128(f->hn, sl_) && (sl_ & 15) < 6) srv |= 1ull << i;
            }
        }
        const int no = __popcll(all), nsrv = __popcll(srv);
        uint64_t from; int site, n;
        if (no > 1) {
            if (nsrv > 0) {
                const double roll = rf(ST_FSMD_ROLL, 0);
                if (roll <= 0.75) { from = srv; site = ST_FSMD_HOST_SRV; n = nsrv; }
                else if (nsrv != no) { from = all & ~srv; site = ST_FSMD_HOST_OTHER; n = no - nsrv; }
                else { from = srv; site = ST_FSMD_HOST_SRV2; n = nsrv; }
            } else { from = all; site = ST_FSMD_HOST_ALL; n = no; }
        } else { from = all; site = ST_FSM_HOST; n = no; }
        const int kth = (int)ri(site, (uint32_t)n);
        const uint32_t flo = (uint32_t)from, fhi = (uint32_t)(from >> 32);
        const int cl = __popc(flo);
        const int pos = kth < cl ? (int)__fns(flo, 0, kth + 1) : 32 + (int)__fns(fhi, 0, kth - cl + 1);
        slot = s->fsm_order[r][pos];
    }
    return fsm_choose_action<1>(c, r, slot, ri, rf);
}

/* RandomSelectRedAgent.get_action RandomSelectRedAgent.py:37-92 [lane], out of line: a uniformly drawn valid command, then one
 * uniform draw per parameter that has more than one known option.  Options come in action-space dict order: subnets and
 * hosts in creation order (per subnet the user hosts, then the servers: EnterpriseScenarioGenerator.py:342-366), sessions in the
 * order they became known.  The words are those of the generic stream functions (oracle: red_policy_random). */
__device__ __noinline__ uint32_t red_policy_random(Cx c, int r)
{
    Cc4Dyn *s = DYN(c);
    const int actor = A_RED0 + r;
    auto ri = [&](int site, uint32_t n) -> uint32_t { return rng_int(c, actor, site, n); };
    /* a known-host mask by creation rank: inside every 16-bit subnet group the ten user slots (6..15) come before the six servers */
    auto ranked = [](uint32_t w) -> uint32_t {
        const uint32_t lo = w & 0xFFFFu, hi = w >> 16;
        return ((lo >> 6) | (lo & 0x3Fu) << 10) | ((hi >> 6) | (hi & 0x3Fu) << 10) << 16;
    };
    auto nth_host = [&](const uint32_t *known, int k) -> int {        /* k-th known host in creation order -> slot */
        _Pragma("unroll 1") for (int w = 0; w < 4; w++) {
            const uint32_t m = ranked(known[w]);
            const int cnt = __popc(m);
            if (k < cnt) { const int rank = w * 32 + nth_bit(m, k), j = rank & 15; return (rank & ~15) | (j < 10 ? 6 + j : j - 10); }
            k -= cnt;
        }
        return 0;
    };
    const int nss = s->n_ss[r], nsub = __popc((uint32_t)s->sn_known[r]);
    const int nip = __popc(s->ip_known[r][0]) + __popc(s->ip_known[r][1]) + __popc(s->ip_known[r][2]) + __popc(s->ip_known[r][3]);
    const int nhn = __popc(s->hn_known[r][0]) + __popc(s->hn_known[r][1]) + __popc(s->hn_known[r][2]) + __popc(s->hn_known[r][3]);
    /* action_space['action'] order = red_actions (EnterpriseScenarioGenerator.py:764-768), four bits each; valid = every parameter has an option */
    const uint64_t cmds = (uint64_t)AC_DRS | (uint64_t)AC_AGGR << 5 | (uint64_t)AC_STEALTH << 10 | (uint64_t)AC_EXPLOIT << 15 | (uint64_t)AC_PRIVESC << 20 |
                          (uint64_t)AC_DEGRADE << 25 | (uint64_t)AC_DD << 30 | (uint64_t)AC_IMPACT << 35 | (uint64_t)AC_WITHDRAW << 40 | (uint64_t)AC_SLEEP << 45;
    uint64_t list = 0; int n = 0;
    _Pragma("unroll 1") for (int i = 0; i < 10; i++) {
        const int cd = (int)(cmds >> (5 * i) & 31u);
        const bool ok = cd == AC_SLEEP ? true : cd == AC_DRS ? (nss && nsub) : (cd == AC_PRIVESC || cd == AC_DEGRADE || cd == AC_IMPACT) ? (nss && nhn)
                        : cd == AC_WITHDRAW ? (nss && nip && nhn) : (nss && nip);
        if (ok) list |= (uint64_t)cd << (5 * n++);
    }
    Act a = { (int)(list >> (5 * ri(ST_RS_COMMAND, (uint32_t)n)) & 31u), -1, -1, 0 };
    auto pick = [&](int cnt) -> int { return cnt > 1 ? (int)ri(ST_RS_PARAM, (uint32_t)cnt) : 0; };
    auto session = [&]() { a.session = s->ss[r][pick(nss)]; };
    switch (a.code) {       /* parameters in constructor-signature order */
    case AC_DRS: a.subnet = (int)__fns((uint32_t)s->sn_known[r], 0, pick(nsub) + 1); session(); break;
    case AC_AGGR: case AC_STEALTH: case AC_DD: session(); a.host = nth_host(s->ip_known[r], pick(nip)); break;
    case AC_EXPLOIT: a.host = nth_host(s->ip_known[r], pick(nip)); session(); break;
    case AC_PRIVESC: case AC_DEGRADE: case AC_IMPACT: a.host = nth_host(s->hn_known[r], pick(nhn)); session(); break;
    case AC_WITHDRAW: session(); (void)pick(nip); a.host = nth_host(s->hn_known[r], pick(nhn)); break;     /* ip_address only routes the RemoteAction */
    default: break;
    }
    return pack_act(a, 0);
}

/* FiniteStateRedAgent.get_action :58-123 for one red agent [lane].  Returns the packed action (ticks = 0). */
template <int RV> __device__ CC4_INL8 uint32_t red_policy(Cx c, int r)
{
    Cc4Dyn *s = DYN(c);
    int actor = A_RED0 + r;
    if (RV == 2) return red_policy_random(c, r);         /* draws every tick, also while an action is in flight */
    if (IP_CODE(s->inprog[actor]) != IP_NONE) return ACT_SLEEP_WORD;     /* last observation was IN_PROGRESS :100-102 */
    const Cc4Fsm *f = &s->fsm[r];
    const int nf = s->n_fsm[r];
    /* hosts that are not in state F, by list position */
    const uint32_t lo = ~f->posF[0] & (nf >= 32 ? 0xFFFFFFFFu : (1u << nf) - 1u), hi = nf > 32 ? ~f->posF[1] & (nf >= 64 ? 0xFFFFFFFFu : (1u << (nf - 32)) - 1u) : 0u;
    const int nlo = __popc(lo), nk = nlo + __popc(hi);
    if (nk == 0) { fail(c, CC4_E_UNREACHABLE); return ACT_SLEEP_WORD; }
    if (RV == 1) return red_policy_discovery(c, r, lo, hi);
    /* production mode: host pick, action pick and at most two parameter picks fit the four words of the stream's first
       Philox block, which stays in registers; the block is left in the stream cache for the action code that follows */
    const bool tape = CC4_TAPE_ON();
#if CC4_POLICY_PHILOX_INLINE
    const uint4 blk = tape ? make_uint4(0, 0, 0, 0) : philox_block_inl(c, actor, 0);
#else
    const uint4 blk = tape ? make_uint4(0, 0, 0, 0) : philox_block(c, actor, 0);
#endif
    uint32_t j = 0;
    auto ri = [&](int site, uint32_t n) -> uint32_t {
        if (tape) return rng_int(c, actor, site, n);
        if (n == 1 && site == ST_FSM_PARAM) return 0u;
        return __umulhi(u4_word(blk, j++), n);
    };
    auto rf = [&](int site, uint32_t n) -> double { return tape ? rng_f64(c, actor, site, n) : f64_from_word(u4_word(blk, j++)); };
    /* FiniteStateRedAgent: uniform over the known hosts in list order */
    const int pick = (int)ri(ST_FSM_HOST, (uint32_t)nk);
    const int pos = pick < nlo ? nth_bit(lo, pick) : 32 + nth_bit(hi, pick - nlo);
    const uint32_t act = tape ? fsm_choose_action<RV>(c, r, s->fsm_order[r][pos], ri, rf) : fsm_choose_action_words<RV>(c, r, s->fsm_order[r][pos], blk, j);
    if (!tape) {
        Scratch *x = SCR(c);
        ctr_set(x, actor, j);
        x->blk_id[r] = 0; x->blk[r][0] = blk.x; x->blk[r][1] = blk.y; x->blk[r][2] = blk.z; x->blk[r][3] = blk.w;
    }
    return act;
}

/* ------------------------------------------------------------------------------------------------ */
/* blue: index decoding (BlueFixedActionWrapper._populate_action_space :233-309) and actions         */
/* ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int blue_sorted_subnet(int b, int i) { return b < 4 ? b : (i == 0 ? S_ADM : i == 1 ? S_OFF : S_PAZ); }
/* index -> packed action (ticks = 0) | ACT_VALID when the index is a mask-True entry */
__device__ CC4_INL8 uint32_t blue_decode(const Cc4Dyn *s, int b, int idx)
{
    Act a = { AC_SLEEP, -1, -1, 0 };
    int n = b == 4 ? 3 : 1, H = 16 * n, T = 8 * n, m = 1, hostidx = -1, tidx = -1, code = AC_SLEEP;
    if (idx < 0) m = 0;
    else if (idx < H) { code = AC_ANALYSE; hostidx = idx; }
    else if (idx == H) code = AC_MONITOR;
    else if (idx < 2 * H + 1) { code = AC_REMOVE; hostidx = idx - H - 1; }
    else if (idx < 3 * H + 1) { code = AC_RESTORE; hostidx = idx - 2 * H - 1; }
    else if (idx == 3 * H + 1) code = AC_SLEEP;
    else if (idx < 3 * H + 2 + T) { code = AC_ALLOW; tidx = idx - 3 * H - 2; }
    else if (idx < 3 * H + 2 + 2 * T) { code = AC_BLOCK; tidx = idx - 3 * H - 2 - T; }
    else if (idx < 4 * H + 2 + 2 * T) { code = AC_DECOY; hostidx = idx - 3 * H - 2 - 2 * T; }
    else m = 0;
    if (hostidx >= 0) {
        int slot = blue_sorted_subnet(b, hostidx >> 4) * 16 + (hostidx & 15);
        if (!host_exists(s, slot)) { code = AC_SLEEP; m = 0; } else a.host = slot;
    }
    if (tidx >= 0) {
        int dst = blue_sorted_subnet(b, tidx >> 3), j = tidx & 7, src = -1;
        _Pragma("unroll 1") for (int p = 0, k = 0; p < 9; p++) { if (c_sorted_subnets[p] == dst) continue; if (k == j) { src = c_sorted_subnets[p]; break; } k++; }
        a.subnet = dst; a.host = src;
    }
    a.code = code;
    return pack_act(a, 0) | (m ? ACT_VALID : 0u);
}

__device__ CC4_INL4 void ev_procpid(Cx c, int slot, int pid)
{
    DYN(c)->ev[slot] |= EV_PROC;
    /* only hosts some blue agent monitors ever have their event pids read (Monitor.py:66-70) */
    if (pid >= 0 && blue_of_subnet(slot_subnet(slot)) >= 0) {
        if (SCR(c)->n_evpid < 8) SCR(c)->evpid[SCR(c)->n_evpid++] = (uint32_t)slot << 16 | (uint32_t)pid; else fail(c, CC4_E_CAPACITY);
    }
}

/* StopProcess.execute_targeteted_local_action + kill_process StopProcess.py:23-58 [uniform] */
__device__ CC4_INL6 bool stop_process(Cx c, int actor, int h, uint32_t pid, bool stop_all = false)
{
    Cc4Dyn *s = DYN(c);
    int pi = proc_find(c, h, pid);
    if (pi < 0) return false;
    uint32_t w = PROCS(c, h)[pi];
    if (!stop_all && P_ROOT(w)) return false;
    int si = session_from_pid(s, h, pid);
    /* a stopped OT service keeps Service.process; a pid collision with it takes the reference down a path that
       re-labels the killed process as the service's (StopProcess.py:37-42): not modelled */
    { const uint32_t *pp = PROCS(c, h); const int np = NPROC(c, h);
      uint32_t clash = 0;
      _Pragma("unroll 1") for (int i = lane_id(); i < np; i += 32) clash |= (uint32_t)(P_DEAD(pp[i]) && P_PID(pp[i]) == pid);
      if (__any_sync(FULLMASK, clash)) fail(c, CC4_E_UNREACHABLE); }
    proc_delete(c, h, pi);
    if (P_SVC(w)) {
        int np = create_pid_on(c, actor, h);
        proc_append(c, h, P_MAKE(np, P_KIND(w), P_ROOT(w), 1));
        if (si >= 0) fail(c, CC4_E_UNREACHABLE);   /* a session whose process is a service process */
    }
    if (si < 0) return true;
    int r = SF_AGENT(s->sessf[si]);
    sess_delete(c, si);
    if (s->sess_count[r] > 0) s->sess_count[r]--;
    return true;
}

/* RestoreFromBackup.py:9-19 + Host.restore Host.py:373-429 [uniform] */
__device__ CC4_INL6 void do_restore(Cx c, int h)
{
    Cc4Dyn *s = DYN(c); const Cc4Stat *st = WC(c)->st;
    /* the backup (static record, cold) is requested before the pool scan: three independent loads in flight instead of three
       round trips one after the other; the largest pid comes from the same registers instead of a read-back of the arena */
    const int on = st->o_nproc[h];
    const uint32_t mine = lane_id() < CC4_OSLAB ? st->o_procs[h][lane_id()] : 0u;
    const uint64_t svc0 = st->o_svc64[h];
    _Pragma("unroll 1") for (;;) {      /* every session on the host goes, in pool order */
        int i = sess_first(s, [=](uint32_t w, uint32_t) { return S_HOST(w) == (uint32_t)h; });
        if (i < 0) break;
        sess_delete(c, i);
    }
    s->ev[h] = 0;
    Cc4Cold *k = WC(c)->k;
    if (lane_id() < on) k->parena[h * CC4_OSLAB + lane_id()] = mine;   /* back in the home slab (on <= CC4_OSLAB) */
    k->pdir[h] = PD_MAKE(h * CC4_OSLAB, on, 3);
    k->pmax[h] = (uint16_t)__reduce_max_sync(FULLMASK, lane_id() < on && !P_DEAD(mine) ? P_PID(mine) : 0u);      /* proc_rescan_max of the restored list */
    k->svc64[h] = svc0;
}

/* production mode: the next two words of a blue actor's stream from one Philox block (blue streams are not cached in the
   scratch: a blue agent draws at most twice per tick) */
#ifndef CC4_BLUE_PHILOX_INLINE
#define CC4_BLUE_PHILOX_INLINE 1     /* -2.7 %: DeployDecoy sits in the heaviest blue chains */
#endif
#if CC4_BLUE_PHILOX_INLINE     /* inline rounds: the decoy's cold loads stay in flight across them (a call waits for loads in flight) */
__device__ __forceinline__ uint2 blue_words2(Cx c, int actor)
{
    Scratch *x = SCR(c);
    const uint32_t j = ctr_get(x, actor);
    ctr_set(x, actor, j + 2);
    const uint4 v = philox_block_inl(c, actor, j >> 2);
    const uint32_t w0 = u4_word(v, j & 3u);
    const uint32_t w1 = (j & 3u) < 3u ? u4_word(v, (j & 3u) + 1u) : philox_block(c, actor, (j >> 2) + 1u).x;
    return make_uint2(w0, w1);
}
#else
__device__ __noinline__ uint2 blue_words2(Cx c, int actor)
{
    Scratch *x = SCR(c);
    const uint32_t j = ctr_get(x, actor);
    ctr_set(x, actor, j + 2);
    const uint4 v = philox_block(c, actor, j >> 2);
    const uint32_t w0 = u4_word(v, j & 3u);
    const uint32_t w1 = (j & 3u) < 3u ? u4_word(v, (j & 3u) + 1u) : philox_block(c, actor, (j >> 2) + 1u).x;
    return make_uint2(w0, w1);
}
#endif

/* DecoyAction.execute DecoyAction.py:47-112 with DeployDecoy.CANDIDATE_DECOYS DeployDecoy.py:27 [uniform] */
__device__ CC4_INL6 void do_decoy(Cx c, int b, int h)
{
    int actor = A_BLUE0 + b;
    Cc4Cold *k = WC(c)->k;
    /* the host's cold words are requested together; the first 32 list entries stay in a register for both passes */
    const uint32_t d = k->pdir[h], pmax = k->pmax[h];
    uint64_t v = k->svc64[h];
    /* production mode: the two words of the draws below are taken now, while the cold words are on their way */
    uint2 dw = make_uint2(0u, 0u);
    if (CC4_BLUE_PHILOX_INLINE && !CC4_TAPE_ON()) dw = blue_words2(c, actor);
    uint32_t *pp_h = k->parena + PD_OFF(d); const int np_h = (int)PD_CNT(d);
    const uint32_t w0 = lane_id() < np_h ? pp_h[lane_id()] : P_NONE;
    uint32_t used = 0;
    {
        if (!P_DEAD(w0)) { uint32_t pb = kind_portbit(P_KIND(w0)); if (pb != 15u) used |= 1u << pb; }
        _Pragma("unroll 1") for (int i = 32 + lane_id(); i < np_h; i += 32) { uint32_t w = pp_h[i]; if (P_DEAD(w)) continue; uint32_t pb = kind_portbit(P_KIND(w)); if (pb != 15u) used |= 1u << pb; }
        used = __reduce_or_sync(FULLMASK, used);
    }
    /* compatible factories in CANDIDATE_DECOYS order; VsftpdDecoyFactory tests port 21, which nothing uses, and then opens
       port 80 (DecoyVsftpd.py:17-20): always compatible */
    uint32_t fac = 0; int nf = 0;       /* packed list, 4 bits per entry */
    if (!(used & PB_80)) fac |= (uint32_t)K_D_APACHE << (4 * nf++);
    if (!(used & PB_443)) fac |= (uint32_t)K_D_TOMCAT << (4 * nf++);
    if (!(used & PB_25)) fac |= (uint32_t)K_D_HARAKA << (4 * nf++);
    fac |= (uint32_t)K_D_VSFTPD << (4 * nf++);
    uint32_t pick, pid_draw;        /* DecoyAction.py:64 choice(factories), Host.create_pid Host.py:198-200 */
    if (CC4_TAPE_ON()) { pick = rng_int(c, actor, ST_DECOY_FACTORY, (uint32_t)nf); pid_draw = rng_int(c, actor, ST_CREATE_PID, 9); }
    else { const uint2 w = CC4_BLUE_PHILOX_INLINE ? dw : blue_words2(c, actor); pick = __umulhi(w.x, (uint32_t)nf); pid_draw = __umulhi(w.y, 9u); }
    int kind = (int)(fac >> (4 * pick) & 15u);
    int pid = (int)(pmax + 1 + pid_draw);
    if (pid > 0xFFFF) { fail(c, CC4_E_CAPACITY); pid = 0xFFFF; }
    /* host.services[name] = Service(pid): an earlier process of the same decoy stops being the service's process */
    {
        if (lane_id() < np_h && P_KIND(w0) == (uint32_t)kind && P_SVC(w0)) pp_h[lane_id()] = w0 & ~(1u << 21);
        _Pragma("unroll 1") for (int i = 32 + lane_id(); i < np_h; i += 32) { uint32_t w = pp_h[i]; if (P_KIND(w) == (uint32_t)kind && P_SVC(w)) pp_h[i] = w & ~(1u << 21); }
        __syncwarp();
    }
    proc_append(c, h, P_MAKE(pid, kind, 0, 1), d, pmax);
    int n = svc_count(v), at = -1;
    _Pragma("unroll 1") for (int i = 0; i < n; i++) if ((SVC_ENTRY(v, i) & 15u) == (uint32_t)kind) at = i;
    if (at < 0) { if (n >= 9) { fail(c, CC4_E_CAPACITY); return; } at = n; }
    k->svc64[h] = svc_set(v, at, (uint32_t)kind | 5u << 4);
}

/* a blue agent's own Monitor action (Monitor.py:35-74 for one agent) [uniform] */
__device__ CC4_INL6 void monitor_one(Cx c, int b)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    _Pragma("unroll 1") for (int h = lane_id(); h < CC4_NSLOT; h += 32) if (blue_of_subnet(slot_subnet(h)) == b) { uint32_t e = s->ev[h]; s->ev[h] = (uint8_t)((e & (EV_PROC | EV_NET)) << 2); }
    __syncwarp();
    int w = 0;
    const int ne = x->n_evpid;
    _Pragma("unroll 1") for (int i = 0; i < ne; i++) {
        const uint32_t e = x->evpid[i];
        int slot = e >> 16;
        if (blue_of_subnet(slot_subnet(slot)) != b) { x->evpid[w++] = e; continue; }
        if (s->n_sus >= CC4_SUS_CAP) { fail(c, CC4_E_CAPACITY); continue; }
        WC(c)->k->sus[s->n_sus] = SU_MAKE(e & 0xFFFFu, slot, b);
        s->n_sus++;
    }
    x->n_evpid = w;
}

__device__ CC4_INL6 void exec_blue(Cx c, int b, uint32_t aw)
{   /* [uniform] */
    Cc4Dyn *s = DYN(c);
    const Act a = unpack_act(aw);
    switch (a.code) {
    case AC_REMOVE: {   /* Remove.py:42-69: StopProcess for every suspicious pid recorded for the host, in list order */
        const int n = s->n_sus;
        const uint32_t *sus = WC(c)->k->sus;
        uint32_t w_next = lane_id() < n ? sus[lane_id()] : 0u;     /* the list is cold: the next 32 entries are requested before these are looked at */
        _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
            const int i = base + lane_id();
            const uint32_t w = w_next;
            if (base + 32 < n) w_next = i + 32 < n ? sus[i + 32] : 0u;
            uint32_t m = __ballot_sync(FULLMASK, i < n && SU_AGENT(w) == (uint32_t)b && SU_SLOT(w) == (uint32_t)a.host);
            while (m) {
                const int j = __ffs(m) - 1; m &= m - 1;
                stop_process(c, A_BLUE0 + b, a.host, SU_PID(__shfl_sync(FULLMASK, w, j)));
            }
        }
        break;
    }
    case AC_MONITOR: monitor_one(c, b); break;
    case AC_RESTORE: do_restore(c, a.host); break;        /* Restore.py:38-68 */
    case AC_DECOY: do_decoy(c, b, a.host); break;
    case AC_BLOCK: s->blocks[a.subnet] |= (uint16_t)(1u << a.host); break;     /* ControlTraffic.py:62-116 */
    case AC_ALLOW: s->blocks[a.subnet] &= (uint16_t)~(1u << a.host); break;    /* ControlTraffic.py:132-185 */
    default: break;   /* Analyse (Analyse.py:40-71) changes no state */
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* green                                                                                            */
/* ------------------------------------------------------------------------------------------------ */
/* PhishingEmail._create_new_session PhishingEmail.py:42-113 ([uniform], in green-agent order); x->g.hostagents is current */
__device__ CC4_INL_RARE void do_phishing(Cx c, int g)
{
    Cc4Dyn *s = DYN(c);
    const uint8_t *ha = reinterpret_cast<const uint8_t *>(SCR(c)->g.hostagents);
    int actor = A_GREEN0 + g, gh = green_host_of(s, g), gsn = gh >> 4;
    if (ha[gh]) return;
    const uint32_t pd_gh = WC(c)->k->pdir[gh], pmax_gh = WC(c)->k->pmax[gh];      /* cold; requested ahead of the owner search */
    /* (host, agent) pairs in state.hosts order (per subnet: users, then servers), agents ascending on each host.  A red
       agent with a session in the green host's own subnet owns the new session (the LAST such host wins: only the inner
       loop breaks, PhishingEmail.py:73-81); otherwise the owner is drawn uniformly from the pairs of the other subnets.
       Lane L looks at the four hosts of order-rank 4L..4L+3 (rank = subnet*16 + {user k | 10 + server k}). */
    const int lane = lane_id();
    uint32_t mine = 0;      /* byte j = agent mask of the host with rank 4*lane + j */
    _Pragma("unroll") for (int j = 0; j < 4; j++) {
        int rank = 4 * lane + j, k = rank & 15, slot = (rank & ~15) | (k < 10 ? 6 + k : k - 10);
        mine |= (uint32_t)ha[slot] << (8 * j);
    }
    const bool own = (lane >> 2) == gsn;
    int src = -1;
    {
        const uint32_t m = __ballot_sync(FULLMASK, own && mine != 0);
        if (m) {
            const int L = 31 - __clz(m);
            const uint32_t v = __shfl_sync(FULLMASK, mine, L);
            const uint32_t last = v >> 24 ? v >> 24 : (v >> 16 & 0xFF) ? (v >> 16 & 0xFF) : (v >> 8 & 0xFF) ? (v >> 8 & 0xFF) : (v & 0xFF);
            src = __ffs(last) - 1;
        }
    }
    if (src < 0) {
        const int cnt = own ? 0 : __popc(mine);
        int incl = cnt;
        _Pragma("unroll") for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
        const int nc = __shfl_sync(FULLMASK, incl, 31);
        if (nc == 0) return;
        const int pk = (int)rng_int(c, actor, ST_PHISH_AGENT, (uint32_t)nc);
        const uint32_t m = __ballot_sync(FULLMASK, pk < incl);
        const int L = __ffs(m) - 1;
        const uint32_t v = __shfl_sync(FULLMASK, mine, L);
        const int before = __shfl_sync(FULLMASK, incl - cnt, L);
        src = (int)__fns(v, 0, pk - before + 1) & 7;      /* bit position inside the lane's four bytes -> agent */
    }
    int pid = create_pid(c, actor, pmax_gh);
    proc_append(c, gh, P_MAKE(pid, K_REDABS, 0, 0), pd_gh, pmax_gh);
    sess_add(c, src, gh, 1, 0, pid);
    reinterpret_cast<uint8_t *>(SCR(c)->g.hostagents)[gh] |= (uint8_t)(1u << src);
}

/* one green agent's tick (EnterpriseGreenAgent.get_action :42-78, GreenLocalWork.py:60-125, GreenAccessService.py:137-217);
 * event bits are OR-ed atomically, the phishing side effect is deferred (bit 8 of the result) because it is order
 * dependent.  Result: bits 0-7 = reward penalty (LWF / ASF, BlueRewardMachine.py:35-65) as a positive number.
 * The agent's Philox block stays in registers: 4 words cover Sleep / AccessService, LocalWork needs a second block. */
#define GREEN_PHISH 0x100
/* tape mode (parity replay): the same tick with the recorded draws, out of line */
__device__ __noinline__ int green_tick_tape(Cx c, int g, uint32_t *ev_words)
{
    Cc4Dyn *s = DYN(c);
    const int actor = A_GREEN0 + g, gh = green_host_of(s, g), gsn = gh >> 4;
    const bool tape = true;
    const uint64_t v = WC(c)->k->svc64[gh];     /* issued ahead of the Philox rounds: only LocalWork (1 in 3) uses it, but the latency hides */
    /* production mode: the agent's whole tick fits the four words of ONE Philox block (word 0: action choice and, cut from
       its low half, the service pick; the phishing follow-up continues the stream at word j) */
    const uint4 blk = tape ? make_uint4(0, 0, 0, 0) : philox_block(c, actor, 0);
    uint32_t j = 1;
    int out = 0;
    uint32_t choice, frac = 0;
    if (tape) choice = rng_int(c, actor, ST_GREEN_ACT, 3);
    else { choice = __umulhi(blk.x, 3u); frac = blk.x * 3u; }
    if (choice == 1) {
        const int n = svc_count(v);
        const bool ot_stopped = (v & SVC_OT_STOPPED) && n > 1 && (SVC_ENTRY(v, 1) & 15u) == K_OT;     /* the OT service is entry 1 wherever it exists */
        const int na = n - (int)ot_stopped;
        if (na == 0) out = -c_reward[s->phase][gsn][0];
        else {
            int pick = tape ? (int)rng_int(c, actor, ST_LW_SERVICE, (uint32_t)na) : (int)__umulhi(frac, (uint32_t)na);
            if (ot_stopped && pick >= 1) pick++;        /* pick-th ACTIVE service */
            const int rel = (int)(SVC_ENTRY(v, pick) >> 4) * 20;
            j = 2;
            if ((int)(tape ? rng_int(c, actor, ST_LW_RELIAB, 100) : __umulhi(blk.y, 100u)) >= rel) out = -c_reward[s->phase][gsn][0];
            else {
                j = 4;
                if ((tape ? rng_f64(c, actor, ST_LW_FP, 0) : f64_from_word(blk.z)) < 0.01) atomicOr(&ev_words[gh >> 2], EV_PROC << ((gh & 3) * 8));
                if ((tape ? rng_f64(c, actor, ST_LW_PHISH, 0) : f64_from_word(blk.w)) < 0.01) out = GREEN_PHISH;
            }
        }
    } else if (choice == 0) {
        /* destination = k-th server (creation order: subnets ascending, server_host_0..) among the allowed subnets,
           GreenAccessService.py:94-115; the cumulative counts per source subnet were tabulated by w_tick */
        const uint64_t cum = SCR(c)->g.as_cum[gsn];
        const uint32_t k = tape ? rng_int(c, actor, ST_AS_DEST, (uint32_t)(cum >> 56)) : __umulhi(blk.y, (uint32_t)(cum >> 56)), kk = k * 0x01010101u;
        j = 3;
        const int dsn = (__popc(__vcmpleu4((uint32_t)cum, kk)) + __popc(__vcmpleu4((uint32_t)(cum >> 32), kk))) >> 3;
        const int d = dsn * 16 + (int)(k - (dsn ? (uint32_t)(cum >> (8 * (dsn - 1))) & 0xFFu : 0u));
        if ((s->blocks[gsn] >> dsn & 1) || (s->blocks[dsn] >> gsn & 1)) { atomicOr(&ev_words[d >> 2], EV_NET << ((d & 3) * 8)); out = -c_reward[s->phase][gsn][1]; }
        else if ((tape ? rng_f64(c, actor, ST_AS_FP, 0) : f64_from_word(blk.z)) < 0.01) atomicOr(&ev_words[d >> 2], EV_NET << ((d & 3) * 8));
    }
    if (!tape) SCR(c)->ctr_g[g] = (uint8_t)j;
    return out;
}

__device__ CC4_INL8 int green_tick(Cx c, int g, uint32_t *ev_words)
{
    if (CC4_TAPE_ON()) return green_tick_tape(c, g, ev_words);
    Cc4Dyn *s = DYN(c);
    const int actor = A_GREEN0 + g, gh = green_host_of(s, g), gsn = gh >> 4;
    const uint64_t v = WC(c)->k->svc64[gh];     /* issued ahead of the Philox rounds: only LocalWork (1 in 3) uses it, but the latency hides */
    /* the agent's whole tick fits the four words of ONE Philox block (word 0: action choice and, cut from its low half, the
       service pick; the phishing follow-up continues the stream at word j).  word / 2^32 < 0.01 is the integer test
       word < P01: the double product is exact and 0.01 * 2^32 = 42 949 672.96 */
    const uint32_t P01 = 42949673u;
#if CC4_GREEN_PHILOX_INLINE
    const uint4 blk = philox_block_inl(c, actor, 0);
#else
    const uint4 blk = philox_block(c, actor, 0);
#endif
    uint32_t j = 1;
    int out = 0;
    const uint32_t choice = __umulhi(blk.x, 3u);
    if (choice == 1) {
        const uint32_t frac = blk.x * 3u;
        const int n = svc_count(v);
        const bool ot_stopped = (v & SVC_OT_STOPPED) && n > 1 && (SVC_ENTRY(v, 1) & 15u) == K_OT;     /* the OT service is entry 1 wherever it exists */
        const int na = n - (int)ot_stopped;
        if (na == 0) out = -c_reward[s->phase][gsn][0];
        else {
            int pick = (int)__umulhi(frac, (uint32_t)na);
            if (ot_stopped && pick >= 1) pick++;        /* pick-th ACTIVE service */
            const uint32_t rel = (SVC_ENTRY(v, pick) >> 4) * 20u;
            j = 2;
            if (__umulhi(blk.y, 100u) >= rel) out = -c_reward[s->phase][gsn][0];
            else {
                j = 4;
                if (blk.z < P01) atomicOr(&ev_words[gh >> 2], EV_PROC << ((gh & 3) * 8));
                if (blk.w < P01) out = GREEN_PHISH;
            }
        }
    } else if (choice == 0) {
        /* destination = k-th server (creation order: subnets ascending, server_host_0..) among the allowed subnets,
           GreenAccessService.py:94-115; the cumulative counts per source subnet were tabulated by ph_setup */
        const uint64_t cum = SCR(c)->g.as_cum[gsn];
        const uint32_t k = __umulhi(blk.y, (uint32_t)(cum >> 56)), kk = k * 0x01010101u;
        j = 3;
        const int dsn = (__popc(__vcmpleu4((uint32_t)cum, kk)) + __popc(__vcmpleu4((uint32_t)(cum >> 32), kk))) >> 3;
        const int d = dsn * 16 + (int)(k - (dsn ? (uint32_t)(cum >> (8 * (dsn - 1))) & 0xFFu : 0u));
        if ((s->blocks[gsn] >> dsn & 1) || (s->blocks[dsn] >> gsn & 1)) { atomicOr(&ev_words[d >> 2], EV_NET << ((d & 3) * 8)); out = -c_reward[s->phase][gsn][1]; }
        else if (blk.z < P01) atomicOr(&ev_words[d >> 2], EV_NET << ((d & 3) * 8));
    }
    SCR(c)->ctr_g[g] = (uint8_t)j;
    return out;
}

/* ------------------------------------------------------------------------------------------------ */
/* red actions (lane 0)                                                                             */
/* ------------------------------------------------------------------------------------------------ */
/* subnets whose routers lie on the tree path between two subnets (Action.py:101-117, topology
 * EnterpriseScenarioGenerator.py:373-416): 4 bits each from bit 0, count in bits 28-31 (tools/gen_tables.py) */
__constant__ uint32_t c_route[9][9] = {
    { 0x10000000u, 0x20000010u, 0x30000280u, 0x40003280u, 0x30000480u, 0x30000580u, 0x40006580u, 0x40007580u, 0x20000080u },
    { 0x20000001u, 0x10000001u, 0x40002801u, 0x50032801u, 0x40004801u, 0x40005801u, 0x50065801u, 0x50075801u, 0x30000801u },
    { 0x30000082u, 0x40001082u, 0x10000002u, 0x20000032u, 0x30000482u, 0x30000582u, 0x40006582u, 0x40007582u, 0x20000082u },
    { 0x40000823u, 0x50010823u, 0x20000023u, 0x10000003u, 0x40004823u, 0x40005823u, 0x50065823u, 0x50075823u, 0x30000823u },
    { 0x30000084u, 0x40001084u, 0x30000284u, 0x40003284u, 0x10000004u, 0x30000584u, 0x40006584u, 0x40007584u, 0x20000084u },
    { 0x30000085u, 0x40001085u, 0x30000285u, 0x40003285u, 0x30000485u, 0x10000005u, 0x20000065u, 0x20000075u, 0x20000085u },
    { 0x40000856u, 0x50010856u, 0x40002856u, 0x50032856u, 0x40004856u, 0x20000056u, 0x10000006u, 0x30000756u, 0x30000856u },
    { 0x40000857u, 0x50010857u, 0x40002857u, 0x50032857u, 0x40004857u, 0x20000057u, 0x30000657u, 0x10000007u, 0x30000857u },
    { 0x20000008u, 0x30000108u, 0x20000028u, 0x30000328u, 0x20000048u, 0x20000058u, 0x30000658u, 0x30000758u, 0x10000008u } };
/* number of hosts on the path src -> dst (both real hosts) and the i-th of them */
__device__ __forceinline__ int route_len(int src, int dst) { return src == dst ? 1 : 2 + (int)(c_route[src >> 4][dst >> 4] >> 28); }
__device__ __forceinline__ int route_hop(int src, int dst, int nh, int i)
{
    if (i == 0) return src;
    if (i == nh - 1) return dst;
    const int v = (int)(c_route[src >> 4][dst >> 4] >> (4 * (i - 1)) & 15u);
    return v == S_INT ? CC4_SLOT_INTERNET : CC4_SLOT_ROUTER + v;
}

/* ExploitAction._create_new_session ExploitAction.py:212-262 [uniform] */
__device__ CC4_INL3 int exploit_new_session(Cx c, int r, int target, uint32_t pd_target, uint32_t pmax_target)
{
    int pid = create_pid_red(c, r, pmax_target);
    proc_append(c, target, P_MAKE(pid, K_SHELL, 0, 0), pd_target, pmax_target);
    sess_add(c, r, target, 0, 0, pid, 1);      /* parent = the acting session, ExploitAction.py:256 */
    return pid;
}
/* index of the known-port entry of (agent, abstract session, host), -1 if none; *word = the entry [uniform] */
__device__ CC4_INL3 int ports_find(Cx c, int r, int id, int host, uint32_t *word)
{
    const uint32_t *ports = WC(c)->k->ports;
    const int n = DYN(c)->n_ports;
    const uint32_t key = PT_MAKE(0, host, id, r);
    uint32_t v_next = lane_id() < n ? ports[lane_id()] : 0xFFFFFFFFu;      /* cold table: the next line is requested before this one is looked at */
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const uint32_t v = v_next;
        if (base + 32 < n) v_next = i + 32 < n ? ports[i + 32] : 0xFFFFFFFFu;
        const uint32_t m = __ballot_sync(FULLMASK, (v & ~0x3Fu) == key);
        if (m) { *word = __shfl_sync(FULLMASK, v, __ffs(m) - 1); return base + __ffs(m) - 1; }
    }
    return -1;
}

/* tape-mode (parity replay) forms of the two draw loops that production mode runs lane-parallel; out of line so that
 * they stay out of the instruction stream of the red phase [uniform] */
__device__ __noinline__ int dd_tape(Cx c, int actor, int h)
{   /* DiscoverDeception.py:83-97: the second draw only happens when the first did not report */
    const uint32_t *pp_h = PROCS(c, h); const int np_h = NPROC(c, h);
    int rep = 0;
    _Pragma("unroll 1") for (int base = 0; base < np_h; base += 32) {
        const uint32_t mine = base + lane_id() < np_h ? pp_h[base + lane_id()] : P_NONE;
        const int cnt = min(32, np_h - base);
        _Pragma("unroll 1") for (int j = 0; j < cnt; j++) {
            uint32_t w = __shfl_sync(FULLMASK, mine, j);
            if (P_DEAD(w)) continue;
            bool decoy = P_KIND(w) >= K_D_APACHE && P_KIND(w) <= K_D_VSFTPD;
            if (rng_f64(c, actor, ST_DD_TP, 0) <= 0.5 && decoy) rep = 1;
            else if (rng_f64(c, actor, ST_DD_FP, 0) <= 0.1 && !decoy) rep = 1;
        }
    }
    return rep;
}
__device__ __noinline__ void ssh_hops_tape(Cx c, int actor, int src, int target, int nh)
{   /* SSHBruteForce.py:36-45 */
    _Pragma("unroll 1") for (int i = 0; i < nh; i++) if (0.05 < rng_f64(c, actor, ST_SSH_HOP, 0)) DYN(c)->ev[route_hop(src, target, nh, i)] |= EV_NET;
}

/* Withdraw.py:38-91 (RandomSelectRedAgent only) [uniform], out of line so that the FiniteStateRedAgent instruction stream does
 * not carry it: every session of the agent on the host goes; si / S_host = pool index / host of the acting session */
__device__ __noinline__ int exec_withdraw(Cx c, int r, int h, int acting, int si, int S_host)
{
    Cc4Dyn *s = DYN(c);
    const int actor = A_RED0 + r, lane = lane_id();
    if (si < 0) return SUCC_FALSE;
    /* the sessions to stop, as idents in the scratch: those with a parent first, then the parentless ones except ident 0, both in
       dict (= pool) order; the acting session last when it sits on the host (a second time when its ident is not 0) */
    uint8_t *ids = SCR(c)->ra.ident;
    int n = 0;
    const int ns = s->n_sess;
    _Pragma("unroll 1") for (int pass = 0; pass < 2; pass++)
        _Pragma("unroll 1") for (int base = 0; base < ns; base += 32) {
            const int i = base + lane;
            const uint32_t w = i < ns ? s->sess[i] : 0u, f = i < ns ? (uint32_t)s->sessf[i] : 0u;
            const bool on = i < ns && SF_AGENT(f) == (uint32_t)r && S_HOST(w) == (uint32_t)h;
            const uint32_t m = __ballot_sync(FULLMASK, on && (pass == 0 ? SF_PARENT(f) != 0u : (!SF_PARENT(f) && S_ID(w) != 0u)));
            const int at = n + __popc(m & ((1u << lane) - 1u));
            if ((m >> lane & 1u) && at < 31) ids[at] = (uint8_t)S_ID(w);
            n += __popc(m);
        }
    if (n > 30) { fail(c, CC4_E_CAPACITY); n = 30; }
    __syncwarp();
    if (S_host == h) { if (lane == 0) ids[n] = (uint8_t)acting; n++; }
    __syncwarp();
    if (n == 0) return SUCC_FALSE;
    _Pragma("unroll 1") for (int k = 0; k < n; k++) {
        /* StopProcess(stop_all=True): TargetedLocalAction.py:14-21 needs the acting and the target session, StopProcess.py:23-33 the process */
        const int ti = sess_find(s, r, ids[k]);
        if (ti < 0 || sess_find(s, r, acting) < 0) return SUCC_FALSE;
        if (!stop_process(c, actor, h, S_PID(s->sess[ti]), true)) return SUCC_FALSE;
    }
    return SUCC_TRUE;
}

/* [uniform]: the whole warp executes the action; scans are lane-parallel */
template <int RV> __device__ CC4_INL7 int exec_red(Cx c, int r, uint32_t aw)
{
    Cc4Dyn *s = DYN(c);
    const Act a = unpack_act(aw);
    const int actor = A_RED0 + r, lane = lane_id();
    int si = (a.code == AC_SLEEP || a.code == AC_INVALID) ? -1 : SCR(c)->sess_moved ? sess_find(s, r, a.session) : (int)SCR(c)->act_si[r];
    int S_host = si >= 0 ? (int)S_HOST(s->sess[si]) : -1, S_abs = si >= 0 ? (int)SF_ABS(s->sessf[si]) : 0;
    /* The observation entries an action reports are filed by ONE copy of obs_entry after the switch (the action bodies only
       queue them, 16 bits each, in report order): ten inlined copies were 4.6 KB of the red phase's executed code. */
    uint64_t pend = 0; int npend = 0, succ = SUCC_UNKNOWN;
    #define OBS_Q(key_ip, slot, iface, subnet_if, syshost, sess) do { pend |= (uint64_t)(0x8000u | (uint32_t)(slot) | (uint32_t)(key_ip) << 8 | (uint32_t)(iface) << 9 | \
        (uint32_t)(subnet_if) << 10 | (uint32_t)(syshost) << 11 | (uint32_t)(sess) << 12) << (16 * npend); npend++; } while (0)
    #define RET(v) do { succ = (v); goto done; } while (0)
    switch (a.code) {
    case AC_DRS: {     /* DiscoverRemoteSystems.py:39 -> Pingsweep.execute Pingsweep.py:31-64 */
        if (si < 0) RET(SUCC_FALSE);
        int sn = a.subnet;
        if (sn == S_INT) RET(SUCC_UNKNOWN);
        int nu = s->nuser[sn], nh = nu + s->nserver[sn];
        if (red_allowed(r, sn)) {   /* _filter_obs SimulationController.py:1054 drops interfaces of foreign subnets */
            /* one ip-keyed Interface + Subnet entry per host of the subnet, users first (state.hosts order): obs_entry for
               all of them at once, lane k = k-th host */
            Scratch *x = SCR(c);
            const int slot = lane < nu ? sn * 16 + 6 + lane : sn * 16 + (lane - nu);
            const uint16_t key = (uint16_t)(slot | 1 << 8), fl = (uint16_t)(1 << 9 | 1 << 10);
            const uint32_t hosts = (((1u << nu) - 1u) << 6 | ((1u << s->nserver[sn]) - 1u)) << ((sn & 1) * 16);
            s->ip_known[r][sn >> 1] |= hosts;
            s->sn_known[r] |= (uint16_t)(1u << sn);
            const int n = x->n_oent[r];
            int at = -1;
            if (lane < nh) _Pragma("unroll 1") for (int i = 0; i < n; i++) if ((x->oent[r][i] & 0x1FF) == key) { at = i; break; }
            __syncwarp();
            if (lane < nh && at >= 0) x->oent[r][at] |= fl;
            const uint32_t fresh = __ballot_sync(FULLMASK, lane < nh && at < 0);
            const int pos = n + __popc(fresh & ((1u << lane) - 1u));
            if ((fresh >> lane & 1) && pos < CC4_OENT_CAP) x->oent[r][pos] = key | fl;
            int nn = n + __popc(fresh);
            if (nn > CC4_OENT_CAP) { fail(c, CC4_E_CAPACITY); nn = CC4_OENT_CAP; }
            x->n_oent[r] = (uint8_t)nn;
            __syncwarp();
        }
        RET(SUCC_TRUE);
    }
    case AC_AGGR: case AC_STEALTH: {   /* DiscoverNetworkServices.py:44-86 -> Portscan.py:23-65 */
        if (si < 0 || !S_abs) RET(SUCC_FALSE);
        if (blocking(s, S_host, a.host)) RET(SUCC_FALSE);
        double rate = a.code == AC_AGGR ? 0.75 : 0.25;
        double u = rng_f64_red(c, r, ST_PORTSCAN);
        const int h = a.host;
        uint32_t mask = 0, alert = 0;
        const uint32_t *pp_h = PROCS(c, h); const int np_h = NPROC(c, h);
        _Pragma("unroll 1") for (int i = lane; i < np_h; i += 32) {
            uint32_t w = pp_h[i];
            if (P_DEAD(w)) continue;
            uint32_t pb = kind_portbit(P_KIND(w));
            if (pb == 15u) continue;
            mask |= 1u << pb;
            if (u <= rate || P_KIND(w) >= K_D_APACHE) alert = 1;
        }
        mask = __reduce_or_sync(FULLMASK, mask);
        if (__any_sync(FULLMASK, alert)) s->ev[h] |= EV_NET;
        if (mask) {
            uint32_t was;
            int at = ports_find(c, r, a.session, h, &was);
            if (at < 0) { if (s->n_ports >= CC4_PORT_CAP) { fail(c, CC4_E_CAPACITY); at = 0; } else { at = s->n_ports; s->n_ports = (uint16_t)(at + 1); } }
            WC(c)->k->ports[at] = PT_MAKE(mask, h, a.session, r);
            OBS_Q(1, h, 1, 0, 0, 0);
        }
        RET(SUCC_TRUE);
    }
    case AC_DD: {      /* DiscoverDeception.py:44-101: two draws per process, in list order */
        if (si < 0) RET(SUCC_FALSE);
        int rep = 0, h = a.host;
        const uint32_t *pp_h = PROCS(c, h); const int np_h = NPROC(c, h);
        if (!CC4_TAPE_ON()) {
            /* production mode spends two words per live process whether or not the second is looked at (oracle: AC_DD), so
               the q-th live process owns words p0 + 2q (decoy test) and p0 + 2q + 1 (false positive): lane per process */
            Scratch *x = SCR(c);
            uint32_t p0 = ctr_get(x, actor);
            _Pragma("unroll 1") for (int base = 0; base < np_h; base += 32) {
                const uint32_t w = base + lane < np_h ? pp_h[base + lane] : P_NONE;
                const bool live = !P_DEAD(w);
                const uint32_t lm = __ballot_sync(FULLMASK, live);
                bool hit = false;
                if (live) {
                    const bool decoy = P_KIND(w) >= K_D_APACHE && P_KIND(w) <= K_D_VSFTPD;
                    const uint32_t j = p0 + 2u * (uint32_t)__popc(lm & ((1u << lane) - 1u)) + (decoy ? 0u : 1u);
                    const uint4 v = philox_block(c, actor, j >> 2);
                    hit = u4_word(v, j & 3u) <= (decoy ? 2147483648u : 429496729u);      /* word / 2^32 <= 0.5 / 0.1, as integers (0.1 * 2^32 = 429 496 729.6) */
                }
                rep |= (int)__any_sync(FULLMASK, hit);
                p0 += 2u * (uint32_t)__popc(lm);
            }
            ctr_set(x, actor, p0);
        } else rep = dd_tape(c, actor, h);
        if (rep) OBS_Q(0, h, 1, 0, 0, 0);
        RET(SUCC_TRUE);
    }
    case AC_EXPLOIT: { /* ExploitRemoteService.py:149-202 */
        if (si < 0 || !S_abs) RET(SUCC_FALSE);
        int target = a.host, src = S_host;
        const uint32_t pd_target = WC(c)->k->pdir[target], pmax_target = WC(c)->k->pmax[target];    /* requested ahead of the port-table scan (all cold) */
        uint32_t pw;
        if (ports_find(c, r, a.session, target, &pw) < 0) RET(SUCC_FALSE);
        const int ports = (int)PT_MASK(pw);
        if (blocking(s, src, target)) RET(SUCC_FALSE);
        enum { X_HTTP, X_HTTPS, X_SSH, X_SQL, X_HARAKA };
        /* DefaultExploitActionSelector :37-69: the options in class order, packed four bits each; weights HTTP 3, HTTPS 4,
           SSH 0.1, SQL 5, Haraka 6, so the option with the largest weight is the last one listed unless SSH is alone */
        uint32_t opts = 0; int no = 0, top = 0;
        if (ports & PB_80) { top = no; opts |= (uint32_t)X_HTTP << (4 * no++); }
        if (ports & PB_443) { top = no; opts |= (uint32_t)X_HTTPS << (4 * no++); }
        if (ports & PB_22) { opts |= (uint32_t)X_SSH << (4 * no++); }
        if ((ports & PB_3390) && (ports & (PB_80 | PB_443))) { top = no; opts |= (uint32_t)X_SQL << (4 * no++); }
        if (ports & PB_25) { top = no; opts |= (uint32_t)X_HARAKA << (4 * no++); }
        if (no < 1) RET(SUCC_FALSE);
        int sel = (int)(opts & 15u);
        if (no > 1) {       /* the top-weighted option is popped, the choice is uniform among the rest (and one more draw is spent) */
            const uint32_t rest = (opts & ((1u << (4 * top)) - 1u)) | (opts >> (4 * (top + 1))) << (4 * top);
            no--;
            sel = (int)(rest >> (4 * rng_int_red(c, r, ST_EXP_OTHER, (uint32_t)no)) & 15u);
            (void)rng_f64_red(c, r, ST_EXP_TOP);
        }
        /* the first live process (list order) that the chosen exploit can attack: ExploitAction._get_vulnerable_process
           ExploitAction.py:178-200 / SSHBruteForce.py:24-84 */
        int vk = -1;
        {
            const uint32_t *pp_target = WC(c)->k->parena + PD_OFF(pd_target); const int np_target = (int)PD_CNT(pd_target);
            _Pragma("unroll 1") for (int base = 0; base < np_target && vk < 0; base += 32) {
                const uint32_t w = base + lane < np_target ? pp_target[base + lane] : P_NONE;
                const int k = (int)P_KIND(w);
                bool hit = false;
                if (!P_DEAD(w)) {
                    if (sel == X_SSH) hit = k == K_SSHD;
                    else if (sel == X_HTTP) hit = k == K_APACHE2 || k == K_D_APACHE || k == K_D_VSFTPD;
                    else if (sel == X_HTTPS) hit = k == K_D_TOMCAT;
                    else if (sel == X_SQL) hit = k == K_MYSQLD;
                    else hit = k == K_SMTP || k == K_D_HARAKA;
                }
                const uint32_t m = __ballot_sync(FULLMASK, hit);
                if (m) vk = __shfl_sync(FULLMASK, k, __ffs(m) - 1);
            }
        }
        if (sel == X_SSH) {   /* SSHBruteForce.execute SSHBruteForce.py:24-84 */
            const int nh = route_len(src, target);
            if (CC4_TAPE_ON()) ssh_hops_tape(c, actor, src, target, nh);
            else {    /* lane i takes the i-th hop and the i-th next word of the stream (the same words the serial loop would take) */
                Scratch *x = SCR(c);
                const uint32_t p0 = ctr_get(x, actor);
                if (lane < nh) {
                    const uint32_t j = p0 + (uint32_t)lane;
                    const uint4 v = philox_block(c, actor, j >> 2);
                    if (u4_word(v, j & 3u) > 214748364u) s->ev[route_hop(src, target, nh, lane)] |= EV_NET;        /* 0.05 < word / 2^32 (0.05 * 2^32 = 214 748 364.8) */
                }
                __syncwarp();
                ctr_set(x, actor, p0 + (uint32_t)nh);
            }
            if (vk < 0) RET(SUCC_FALSE);
            s->ev[target] |= EV_NET;
        } else {
            /* ExploitAction.sim_exploit ExploitAction.py:48-116 */
            if (vk < 0) RET(SUCC_FALSE);
            const bool decoy = vk >= K_D_APACHE && vk <= K_D_VSFTPD;
            if (decoy) s->ev[target] |= EV_NET;
            OBS_Q(1, target, 1, 0, 0, 0);
            if (!(sel == X_SQL || decoy)) RET(SUCC_FALSE);
            if (0.0 >= rng_f64_red(c, r, ST_EXP_SUCC)) RET(SUCC_FALSE);
            if (decoy) { s->ev[target] |= EV_NET; RET(SUCC_FALSE); }
        }
        /* both ways in end here (one copy of the session / process creation): the brute force is always noticed, the exploit
           with probability 0.95 (its draw comes after the new pid's) */
        const int pid = exploit_new_session(c, r, target, pd_target, pmax_target);
        if (sel == X_SSH || rng_f64_red(c, r, ST_EXP_DETECT) > 0.05) ev_procpid(c, target, pid);
        OBS_Q(1, target, 1, 0, 1, 1);
        OBS_Q(1, src, 1, 0, 0, 0);
        RET(SUCC_TRUE);
    }
    case AC_PRIVESC: { /* PrivilegeEscalate.py:127-179 */
        const int h = a.host;
        const uint32_t pd_h = WC(c)->k->pdir[h];       /* cold; requested ahead of the pool scan */
        /* one pass over the pool: how many sessions the agent holds on the host, and the first root one among them */
        auto on_host = [=](uint32_t w, uint32_t f) { return SF_AGENT(f) == (uint32_t)r && S_HOST(w) == (uint32_t)h; };
        int n = 0, tgt = -1;
        const int ns = s->n_sess;
        _Pragma("unroll 1") for (int base = 0; base < ns; base += 32) {
            const int i = base + lane;
            const uint32_t f = i < ns ? (uint32_t)s->sessf[i] : 0u;
            const bool o = i < ns && on_host(s->sess[i], f);
            const uint32_t om = __ballot_sync(FULLMASK, o), rm = __ballot_sync(FULLMASK, o && SF_ROOT(f));
            n += __popc(om);
            if (tgt < 0 && rm) tgt = base + __ffs(rm) - 1;
        }
        if (n == 0) RET(SUCC_FALSE);
        if (tgt < 0) {
            int pk = (int)rng_int_red(c, r, ST_PRIV_SESS, (uint32_t)n);
            tgt = sess_kth(s, pk, on_host);
            if (si < 0 || !S_abs) RET(SUCC_FALSE);
            s->sessf[tgt] |= 1u << 4;      /* EscalateAction.__upgrade_session EscalateAction.py:54-87 */
            int pi = proc_find(c, S_PID(s->sess[tgt]), pd_h);
            if (pi >= 0) (WC(c)->k->parena + PD_OFF(pd_h))[pi] |= P_ROOTBIT; else fail(c, CC4_E_UNREACHABLE);
        }
        OBS_Q(0, h, 0, 0, 0, 1);
        if (SF_ABS(s->sessf[tgt])) know_session(c, r, S_ID(s->sess[tgt]));
        if (si >= 0 && (h & 15) == 0) {   /* ExploreHost EscalateAction.py:90-106, Host.info = _between_subnet_links :418-468 */
            uint32_t lk = c_links[h >> 4];
            _Pragma("unroll 1") for (int sn = 0; sn < 8; sn++) if (lk >> sn & 1) OBS_Q(0, sn * 16, 1, 0, 0, 0);
        }
        RET(SUCC_TRUE);
    }
    case AC_IMPACT: case AC_DEGRADE: {   /* Impact.py:38-88, DegradeServices.py:38-82 */
        const int h = a.host;
        uint64_t v = WC(c)->k->svc64[h];        /* cold; requested ahead of the pool scan */
        uint32_t any = 0, root = 0;
        _Pragma("unroll 1") for (int i = lane; i < s->n_sess; i += 32) if (SF_AGENT(s->sessf[i]) == (uint32_t)r && S_HOST(s->sess[i]) == (uint32_t)h) { any = 1; root |= SF_ROOT(s->sessf[i]); }
        if (!__any_sync(FULLMASK, any && root)) RET(SUCC_FALSE);
        int n = svc_count(v);
        /* all nine 7-bit entries at once: `lanes` = bit 7i of every entry in use, `ot` = of the entry that is the OT service */
        const uint64_t SVC_ONES = 0x0102040810204081ull;
        const uint64_t lanes = n >= 9 ? SVC_ONES : SVC_ONES & ((1ull << (7 * n)) - 1ull);
        const uint64_t t = v ^ ((uint64_t)K_OT * SVC_ONES);
        const uint64_t ot = ~(t | t >> 1 | t >> 2 | t >> 3) & lanes;
        if (a.code == AC_IMPACT) {
            if (!ot || (v & SVC_OT_STOPPED)) RET(SUCC_FALSE);
            if (si < 0) RET(SUCC_FALSE);     /* StopService is a TargetedLocalAction: TargetedLocalAction.py:14-21 */
            WC(c)->k->svc64[h] = v | SVC_OT_STOPPED;  /* Host.stop_service Host.py:296-301 */
            /* State.remove_process State.py:390-418: the first live OT service process becomes a tombstone */
            uint32_t *pp_h = PROCS(c, h); const int np_h = NPROC(c, h);
            int at = -1;
            _Pragma("unroll 1") for (int base = 0; base < np_h && at < 0; base += 32) {
                const uint32_t w = base + lane < np_h ? pp_h[base + lane] : P_NONE;
                const uint32_t m = __ballot_sync(FULLMASK, !P_DEAD(w) && P_KIND(w) == K_OT && P_SVC(w));
                if (m) at = base + __ffs(m) - 1;
            }
            if (at >= 0) {
                const uint32_t w = pp_h[at];
                int s2 = session_from_pid(s, h, P_PID(w));
                pp_h[at] = w | P_DEADBIT;
                if (P_PID(w) == WC(c)->k->pmax[h]) proc_rescan_max(c, h);
                if (s2 >= 0) sess_delete(c, s2);
            }
        } else {
            /* every active service loses one reliability step (20 %), not below 0 */
            const uint64_t act = (v & SVC_OT_STOPPED) ? lanes & ~ot : lanes;
            if (!act) RET(SUCC_FALSE);
            const uint64_t nz = (v >> 4 | v >> 5 | v >> 6) & act;       /* reliability > 0 */
            v -= nz << 4;
            WC(c)->k->svc64[h] = v;
        }
        OBS_Q(0, h, 0, 0, 0, 0);
        RET(SUCC_TRUE);
    }
    case AC_WITHDRAW: RET(RV == 2 ? exec_withdraw(c, r, a.host, a.session, si, S_host) : (int)SUCC_UNKNOWN);      /* only the RandomSelectRedAgent ever submits it */
    case AC_INVALID: RET(SUCC_FALSE);
    default: RET(SUCC_UNKNOWN);
    }
done:
    _Pragma("unroll 1") for (; pend; pend >>= 16) {
        const uint32_t e = (uint32_t)pend & 0xFFFFu;
        obs_entry(c, r, (int)(e >> 8 & 1u), (int)(e & 0xFFu), (int)(e >> 9 & 1u), (int)(e >> 10 & 1u), (int)(e >> 11 & 1u), (int)(e >> 12 & 1u));
    }
    #undef OBS_Q
    #undef RET
    return succ;
}


/* RedSessionCheck._choose_new_primary_session RedSessionCheck.py:34-55 ([uniform]; agent r holds n sessions, none with ident 0) */
template <int RV> __device__ CC4_INL_RARE void rsc_reelect(Cx c, int r, int n)
{
    Cc4Dyn *s = DYN(c);
    int pk = (int)rng_int_red(c, r, ST_RSC_PRIMARY, (uint32_t)n);
    int at = sess_kth(s, pk, [=](uint32_t, uint32_t f) { return SF_AGENT(f) == (uint32_t)r; });
    if (at < 0) { fail(c, CC4_E_UNREACHABLE); return; }
    const uint32_t w = s->sess[at]; const uint8_t f = s->sessf[at];
    const int old_id = S_ID(w);
    {   /* the session's known ports follow it to ident 0 */
        uint32_t *ports = WC(c)->k->ports;
        const int np = s->n_ports;
        _Pragma("unroll 1") for (int i = lane_id(); i < np; i += 32) { uint32_t p = ports[i]; if (PT_AGENT(p) == (uint32_t)r && PT_ID(p) == (uint32_t)old_id) ports[i] = PT_MAKE(PT_MASK(p), PT_HOST(p), 0, r); }
        __syncwarp();
    }
    const int ns = s->n_sess;
    sess_shift_down(s, at, ns);      /* the re-keyed entry moves to the end of the agent's dict */
    s->sess[ns - 1] = S_MAKE(S_PID(w), 0, S_HOST(w)); s->sessf[ns - 1] = f;
    if (RV == 2) {      /* only Withdraw (RandomSelectRedAgent) ever looks at the parent flag */
        __syncwarp();
        /* every session of the agent now hangs under the new primary, whose name is None (RedSessionCheck.py:50-54): no parents left */
        _Pragma("unroll 1") for (int i = lane_id(); i < ns; i += 32) if (SF_AGENT((uint32_t)s->sessf[i]) == (uint32_t)r) s->sessf[i] &= (uint8_t)~SF_PARENTBIT;
        __syncwarp();
    }
}

/* SimulationController.different_subnet_agent_reassignment :820-903 [uniform] */
__device__ CC4_INL_RARE void reassign(Cx c)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    uint8_t *list = x->ra.agent, *lid = x->ra.ident; int n = 0;   /* (agent, ident) pairs, agents ascending then dict order; in the
                                                                       scratch, not in a local array: no thread stack in the hot kernel */
    _Pragma("unroll 1") for (int r = 0; r < CC4_NRED; r++) {
        if (!(x->misplaced >> r & 1)) continue;
        const int ns = s->n_sess;
        _Pragma("unroll 1") for (int base = 0; base < ns; base += 32) {
            const int i = base + lane_id();
            const uint32_t w = i < ns ? s->sess[i] : 0u;
            uint32_t m = __ballot_sync(FULLMASK, i < ns && SF_AGENT(s->sessf[i]) == (uint32_t)r && !red_allowed(r, S_HOST(w) >> 4));
            while (m) {
                const int j = __ffs(m) - 1; m &= m - 1;
                if (n >= 32) { fail(c, CC4_E_CAPACITY); break; }
                const uint32_t id = S_ID(__shfl_sync(FULLMASK, w, j));
                if (lane_id() == 0) { list[n] = (uint8_t)r; lid[n] = (uint8_t)id; }
                n++;
            }
        }
    }
    __syncwarp();
    _Pragma("unroll 1") for (int k = 0; k < n; k++) {
        int r = list[k], si = sess_find(s, r, lid[k]);
        uint32_t w = s->sess[si]; uint8_t f = s->sessf[si];
        int host = S_HOST(w), nr = red_of_subnet(host >> 4);
        if (nr < 0) { fail(c, CC4_E_UNREACHABLE); continue; }
        sess_delete(c, si);
        int nid = sess_add(c, nr, host, 1, SF_ROOT(f), S_PID(w));
        if (s->sess_count[r] > 0) s->sess_count[r]--;
        _Pragma("unroll 1") for (int i = 0; i < x->n_oent[r]; i++) {
            uint16_t e = x->oent[r][i];
            if (((e >> 8) & 1) && (e & 0xFF) == host) {
                if (!((e >> 12) & 1)) { fail(c, CC4_E_UNREACHABLE); break; }
                obs_push(c, nr, SUCC_UNKNOWN, false, ACT_SLEEP_WORD);
                obs_entry(c, nr, 1, host, (e >> 9) & 1, (e >> 10) & 1, (e >> 11) & 1, 1);
                know_session(c, nr, nid);
                break;
            }
        }
    }
}

}  // namespace CC4_NS
