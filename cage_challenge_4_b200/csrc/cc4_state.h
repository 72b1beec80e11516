/*
 * cc4_state.h — packed per-episode state of the batched CC4 engine (shared by device kernels and the host-side
 * canonical exporter).  One episode = one Cc4Dyn record (read + written every tick, staged through shared memory
 * by the warp that owns the episode) + one Cc4Stat record (written at reset; read on Restore / export only).
 * Records of consecutive episodes are contiguous in HBM, so a warp moves its record with 128-bit coalesced accesses.
 *
 * Field semantics cite the reference (paths under CybORG/).
 */
#pragma once
#include <stdint.h>

#define CC4_NSUB 9
#define CC4_NSLOT 137        /* 128 regular host slots (subnet*16 + {server k | 6 + user k}), routers 128+subnet, internet host 136 */
#define CC4_SLOT_ROUTER 128
#define CC4_SLOT_INTERNET 136
#define CC4_NRED 6
#define CC4_NBLUE 5
#define CC4_MAXGREEN 80

#define CC4_OSLAB 8          /* processes a host can have at reset: <= 5 services + blue + green + red session */
#define CC4_ARENA 8192       /* process entries per episode: 128 home slabs of CC4_OSLAB + growth blocks (see Cc4Cold) */
#define CC4_SESS_CAP 128     /* red sessions per episode (all six agents; 63 at most in a 3 000-episode probe) */
#define CC4_PORT_CAP 384     /* (red abstract session, host) known-port entries */
#define CC4_SUS_CAP 256      /* blue suspicious-pid entries */
#define CC4_FSM_CAP 64       /* hosts known to one red agent's FSM */
#define CC4_SS_CAP 48        /* server_session ids known to one red agent */
#define CC4_OENT_CAP 24      /* observation entries of one red agent in one tick (a Pingsweep reports <= 16 hosts) */

/* process entry (Simulator/Process.py:12): pid | kind<<16 | root<<20 | svc<<21 | dead<<22
 * svc  = this process backs host.services[kind] (Host.py:124);
 * dead = tombstone of a stopped service's removed process: keeps Service.process for a stopped OT service
 *        (State.stop_service State.py:467); skipped by everything that iterates host.processes. */
#define P_PID(w) ((w) & 0xFFFFu)
#define P_KIND(w) (((w) >> 16) & 0xFu)
#define P_ROOT(w) (((w) >> 20) & 1u)
#define P_SVC(w) (((w) >> 21) & 1u)
#define P_DEAD(w) (((w) >> 22) & 1u)
#define P_MAKE(pid, kind, root, svc) ((uint32_t)(pid) | (uint32_t)(kind) << 16 | (uint32_t)(root) << 20 | (uint32_t)(svc) << 21)
#define P_DEADBIT (1u << 22)
#define P_ROOTBIT (1u << 20)

/* process kinds == service codes for 0..8 (oracle/canon.py PK) */
enum { K_SSHD, K_OT, K_APACHE2, K_MYSQLD, K_SMTP, K_D_APACHE, K_D_TOMCAT, K_D_HARAKA, K_D_VSFTPD, K_SESSION, K_SHELL, K_REDABS };

/* svc64: host.services in dict order (Host.py:124, Service.py:4): entry i = bits [7i,7i+7) = code | (reliability/20)<<4,
 * code 15 = empty; bit 63 = the OT service is stopped (the only service the CC4 action set can stop, Impact.py:38). */
#define SVC_EMPTY 0x7FFFFFFFFFFFFFFFull
#define SVC_ENTRY(v, i) ((uint32_t)((v) >> (7 * (i))) & 0x7Fu)
#define SVC_OT_STOPPED (1ull << 63)

/* red session (Shared/Session.py:6,120): pid | ident<<16 | host<<24 ; flags: agent | abstract<<3 | root<<4 | has-parent<<5
 * (Session.parent is not None: set by an exploit, ExploitAction.py:256; reset by a re-election, RedSessionCheck.py:50-54; only
 * Withdraw, i.e. the RandomSelectRedAgent, looks at it) */
#define S_PID(w) ((w) & 0xFFFFu)
#define S_ID(w) (((w) >> 16) & 0xFFu)
#define S_HOST(w) (((w) >> 24) & 0x7Fu)
#define S_MAKE(pid, id, host) ((uint32_t)(pid) | (uint32_t)(id) << 16 | (uint32_t)(host) << 24)
#define SF_AGENT(f) ((f) & 7u)
#define SF_ABS(f) (((f) >> 3) & 1u)
#define SF_ROOT(f) (((f) >> 4) & 1u)
#define SF_PARENT(f) (((f) >> 5) & 1u)
#define SF_PARENTBIT (1u << 5)

/* known ports of a RedAbstractSession (Session.py:125-133): mask | host<<6 | ident<<13 | agent<<21 */
#define PT_MASK(w) ((w) & 0x3Fu)
#define PT_HOST(w) (((w) >> 6) & 0x7Fu)
#define PT_ID(w) (((w) >> 13) & 0xFFu)
#define PT_AGENT(w) (((w) >> 21) & 7u)
#define PT_MAKE(mask, host, id, agent) ((uint32_t)(mask) | (uint32_t)(host) << 6 | (uint32_t)(id) << 13 | (uint32_t)(agent) << 21)

/* FiniteStateRedAgent.host_states (FiniteStateRedAgent.py:36) of one red agent is kept as
 *   order[i]   host slot of the i-th entry in insertion order (the order `choice` over the known hosts indexes into),
 *   present    128-bit mask: the host has an entry,
 *   hn         128-bit mask: the entry carries a hostname,
 *   pl[b]      128-bit mask: bit b of the entry's 4-bit state (K KD S SD U UD R RD F = 0..8: bit 0 is the 'D' suffix,
 *              states U..RD are exactly pl[2] & ~pl[3], F is pl[3]),
 *   posF       64-bit mask over list POSITIONS: the entry is in state F (F is absorbing),
 * so that the per-tick bookkeeping is mask algebra instead of a walk over the list. */
struct Cc4Fsm {
    uint32_t pl[4][4];
    uint32_t present[4], hn[4];
    uint32_t posF[2];
};

/* in-flight action (SimulationController.actions_in_progress :246): code | host<<6 | subnet<<14 | session<<18 | ticks<<26 */
#define IP_NONE 63u
#define IP_CODE(w) ((w) & 63u)
#define IP_HOST(w) (((w) >> 6) & 0xFFu)
#define IP_SUBNET(w) (((w) >> 14) & 0xFu)
#define IP_SESSION(w) (((w) >> 18) & 0xFFu)
#define IP_TICKS(w) (((w) >> 26) & 0xFu)
#define IP_MAKE(code, host, subnet, session, ticks) ((uint32_t)(code) | (uint32_t)((host) & 0xFF) << 6 | (uint32_t)((subnet) & 0xF) << 14 | (uint32_t)((session) & 0xFF) << 18 | (uint32_t)(ticks) << 26)

/* blue suspicious pid (VelociraptorServer.sus_pids Session.py:154-165): pid | slot<<16 | agent<<24 */
#define SU_PID(w) ((w) & 0xFFFFu)
#define SU_SLOT(w) (((w) >> 16) & 0xFFu)
#define SU_AGENT(w) (((w) >> 24) & 7u)
#define SU_MAKE(pid, slot, agent) ((uint32_t)(pid) | (uint32_t)(slot) << 16 | (uint32_t)(agent) << 24)

/* host event bits (Simulator/HostEvents.py:7) */
#define EV_PROC 1u
#define EV_NET 2u
#define EV_OLDPROC 4u
#define EV_OLDNET 8u

/* HOT record: read and written every tick, staged through shared memory by the warp that owns the episode. */
struct __attribute__((aligned(16))) Cc4Dyn {
    /* ---- red sessions: one pool in creation order; an agent's state.sessions[agent] dict order is the pool order */
    uint32_t sess[CC4_SESS_CAP];
    /* ---- red knowledge (Shared/ActionSpace.py:170-212) ---- */
    uint32_t ip_known[CC4_NRED][4];
    uint32_t hn_known[CC4_NRED][4];
    uint32_t inprog[16];                /* actors 0..10 (blue 0-4, red 5-10) */
    struct Cc4Fsm fsm[CC4_NRED];
    uint8_t fsm_order[CC4_NRED][CC4_FSM_CAP];
    uint16_t sn_known[8];
    uint16_t blocks[12];                /* blocks[to] bit from (State.blocks State.py:148); 9 used */
    uint8_t ev[144];                    /* 137 used */
    uint8_t sessf[CC4_SESS_CAP];
    uint8_t ss[CC4_NRED][CC4_SS_CAP];   /* known server_session ids in dict order */
    uint64_t green_cum;                 /* byte sn = number of green agents (= user hosts, EnterpriseScenarioGenerator.py:700-749) in subnets <= sn:
                                           green agent g sits on user host (g - byte[sn-1]) of the subnet sn whose byte first exceeds g */
    uint8_t n_ss[8];
    uint8_t n_fsm[8];
    uint8_t sess_count[8];              /* state.sessions_count[red agent] (stale after Restore, RestoreFromBackup.py:9-19) */
    uint8_t msgs[8];
    uint8_t nuser[8], nserver[8];
    uint16_t n_sess, n_ports, n_sus, step;
    uint8_t phase, done, fsm_started, red_active, ngreen, pad0[3];
    uint32_t err;
    uint32_t pad1[3];
};

/* COLD record: dynamic state that only the rare transitions touch (process lists on exploit / decoy / remove / restore,
 * known ports on scans and exploits, suspicious pids on detection and Remove).  It stays in HBM and is accessed in place
 * by the lane that executes the transition; it is NOT streamed every tick. */
/* process directory entry of one host: arena offset | count<<13 | log2(capacity)<<23 */
#define PD_OFF(w) ((w) & 0x1FFFu)
#define PD_CNT(w) (((w) >> 13) & 0x3FFu)
#define PD_CAPL(w) (((w) >> 23) & 0xFu)
#define PD_MAKE(off, cnt, capl) ((uint32_t)(off) | (uint32_t)(cnt) << 13 | (uint32_t)(capl) << 23)
#define PD_MAXCNT 1023
struct __attribute__((aligned(16))) Cc4Cold {
    /* host.processes in order (Host.py:117).  Host h starts in its home slab parena[h*8 .. h*8+8); a list that outgrows its
     * block moves to a block of twice the size at arena_top (old blocks are not recycled; Restore returns the host to
     * its home slab).  The reference puts no bound on the list: a blue agent can deploy a decoy on the same host every
     * second tick (VsftpdDecoyFactory is always compatible, DecoyVsftpd.py:17-20), ~250 per episode; 8192 entries hold
     * five such hosts (5 x 1008 entries of growth blocks) next to the 1024 home entries. */
    uint32_t parena[CC4_ARENA];
    uint32_t pdir[128];
    /* host.services (svc64 encoding above): read by one green agent per host per tick (GreenLocalWork), written by decoys,
     * Impact / DegradeServices and Restore: sparse, so it is accessed in place instead of being staged (1 KB of shared
     * memory per warp buys more resident warps than the staging saved latency) */
    uint64_t svc64[128];
    uint32_t ports[CC4_PORT_CAP];
    uint32_t sus[CC4_SUS_CAP];
    uint16_t pmax[128];                 /* largest live pid per host (Host.create_pid Host.py:198-200 starts from it) */
    uint32_t arena_top, pad[3];
};

struct __attribute__((aligned(16))) Cc4Stat {
    uint32_t o_procs[128][CC4_OSLAB];   /* Host.create_backup Host.py:316 (original_processes) */
    uint64_t o_svc64[128];              /* original_services */
    uint8_t o_nproc[128];
    uint16_t router_pid[16];            /* pid of the blue session process on each router (0 = none), internet host has none */
    uint8_t blue_parent[8];
    uint8_t red_start[8];
};

#ifdef __cplusplus
static_assert(sizeof(Cc4Dyn) % 16 == 0, "Cc4Dyn must be a multiple of 16 bytes");
static_assert(sizeof(Cc4Stat) % 16 == 0, "Cc4Stat must be a multiple of 16 bytes");
static_assert(sizeof(Cc4Cold) % 16 == 0, "Cc4Cold must be a multiple of 16 bytes");
#endif
