/*
 * cc4_oracle.c — serial CPU restatement of the CybORG CC4 multi-agent step.  TEST INFRASTRUCTURE ONLY
 * (see cc4_oracle.h for the rules and the parity status).  Plain C, one episode per env, data structures
 * chosen to mirror the reference's Python objects one-to-one (ordered lists where the reference iterates
 * dicts/lists), not for speed.  All citations are file:line under /root/reference/CybORG/.
 */
#include "cc4_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* constants                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
#define NSUB 9
#define NSLOT 137
#define SLOT_ROUTER 128
#define SLOT_INTERNET 136
#define NBLUE 5
#define NRED 6
#define MAXGREEN 80
#define NAG CC4O_N_ACTORS
#define A_BLUE0 0
#define A_RED0 5
#define A_GREEN0 11
#define MAXPROC 640    /* the reference puts no bound on host.processes: one decoy every second tick is possible (DecoyVsftpd.py:17-20) */
#define MAXSVC 9
#define MAXSESS 160
#define MAXSUS 256
#define MAXFSM 140
#define MAXOBS 200
#define MAXSS 160

/* subnets in creation order, Simulator/Scenarios/EnterpriseScenarioGenerator.py:40-51 */
enum { S_RZA, S_OZA, S_RZB, S_OZB, S_CON, S_PAZ, S_ADM, S_OFF, S_INT };
/* sorted-by-name order (BlueFlatWrapper.py:195): admin, contractor, internet, office, oza, ozb, paz, rza, rzb */
static const int SORTED_SUBNETS[NSUB] = { S_ADM, S_CON, S_INT, S_OFF, S_OZA, S_OZB, S_PAZ, S_RZA, S_RZB };
static int SORTED_POS[NSUB]; /* creation idx -> sorted position */

/* process kinds (oracle/canon.py PK) */
enum { K_SSHD, K_OT, K_APACHE2, K_MYSQLD, K_SMTP, K_D_APACHE, K_D_TOMCAT, K_D_HARAKA, K_D_VSFTPD, K_SESSION, K_SHELL, K_REDABS };
/* service codes (canon SERVICE_NAMES) equal kinds 0..8 */
/* listening port bit per kind (canon PORT_BIT: 22->0, 1->1, 80->2, 3390->3, 25->4, 443->5) */
static const int KIND_PORTBIT[12] = { 0, 1, 2, 3, 4, 2, 5, 4, 2, -1, -1, -1 };
#define PB_22 1
#define PB_80 4
#define PB_3390 8
#define PB_25 16
#define PB_443 32

/* action codes (canon ACTION_CODE) */
enum { AC_NONE = -1, AC_SLEEP = 0, AC_MONITOR, AC_ANALYSE, AC_REMOVE, AC_RESTORE, AC_DECOY, AC_BLOCK, AC_ALLOW, AC_INVALID,
       AC_DRS = 10, AC_AGGR, AC_STEALTH, AC_DD, AC_EXPLOIT, AC_PRIVESC, AC_IMPACT, AC_DEGRADE, AC_WITHDRAW,
       AC_G_ACCESS = 20, AC_G_LOCAL = 21 };

/* FSM states, Agents/SimpleAgents/FiniteStateRedAgent.py:401-411 */
enum { F_K, F_KD, F_S, F_SD, F_U, F_UD, F_R, F_RD, F_F };
/* FSM action columns :393-399: DRS, Aggr, Stealth, DD, Exploit, PrivEsc, Impact, Degrade, Withdraw */
static const int FSM_COL_CODE[9] = { AC_DRS, AC_AGGR, AC_STEALTH, AC_DD, AC_EXPLOIT, AC_PRIVESC, AC_IMPACT, AC_DEGRADE, AC_WITHDRAW };
static const int8_t FSM_SUCC[9][9] = { /* :441-451, -1 = None */
    { F_KD, F_S, F_S, -1, -1, -1, -1, -1, -1 },   { F_KD, F_SD, F_SD, -1, -1, -1, -1, -1, -1 },
    { F_SD, -1, -1, F_S, F_U, -1, -1, -1, -1 },   { F_SD, -1, -1, F_SD, F_UD, -1, -1, -1, -1 },
    { F_UD, -1, -1, -1, -1, F_R, -1, -1, F_S },   { F_UD, -1, -1, -1, -1, F_RD, -1, -1, F_SD },
    { F_RD, -1, -1, -1, -1, -1, F_R, F_R, F_S },  { F_RD, -1, -1, -1, -1, -1, F_RD, F_RD, F_SD },
    { F_F, -1, -1, -1, -1, -1, -1, -1, -1 } };
static const int8_t FSM_FAIL[9][9] = { /* :481-491 */
    { F_K, F_K, F_K, -1, -1, -1, -1, -1, -1 },    { F_KD, F_KD, F_KD, -1, -1, -1, -1, -1, -1 },
    { F_S, -1, -1, F_S, F_S, -1, -1, -1, -1 },    { F_SD, -1, -1, F_SD, F_SD, -1, -1, -1, -1 },
    { F_U, -1, -1, -1, -1, F_U, -1, -1, F_U },    { F_UD, -1, -1, -1, -1, F_UD, -1, -1, F_UD },
    { F_R, -1, -1, -1, -1, -1, F_R, F_R, F_R },   { F_RD, -1, -1, -1, -1, -1, F_RD, F_RD, F_RD },
    { F_F, -1, -1, -1, -1, -1, -1, -1, -1 } };
static const double FSM_PROB[2][8][9] = { { /* FiniteStateRedAgent.py:540-549, <0 = None */
    { 0.5, 0.25, 0.25, -1, -1, -1, -1, -1, -1 },  { -1, 0.5, 0.5, -1, -1, -1, -1, -1, -1 },
    { 0.25, -1, -1, 0.25, 0.5, -1, -1, -1, -1 },  { -1, -1, -1, 0.25, 0.75, -1, -1, -1, -1 },
    { 0.5, -1, -1, -1, -1, 0.5, -1, -1, 0.0 },    { -1, -1, -1, -1, -1, 1.0, -1, -1, 0.0 },
    { 0.5, -1, -1, -1, -1, -1, 0.25, 0.25, 0.0 }, { -1, -1, -1, -1, -1, -1, 0.5, 0.5, 0.0 } },
  { /* DiscoveryFSRed.state_transitions_probability FSMRedVariants.py:110-120 */
    { 0.25, 0.75, 0.0, -1, -1, -1, -1, -1, -1 },  { -1, 1.0, 0.0, -1, -1, -1, -1, -1, -1 },
    { 0.25, -1, -1, 0.0, 0.75, -1, -1, -1, -1 },  { -1, -1, -1, 0.0, 1.0, -1, -1, -1, -1 },
    { 0.0, -1, -1, -1, -1, 1.0, -1, -1, 0.0 },    { -1, -1, -1, -1, -1, 1.0, -1, -1, 0.0 },
    { 1.0, -1, -1, -1, -1, -1, 0.0, 0.0, 0.0 },   { -1, -1, -1, -1, -1, -1, 0.5, 0.5, 0.0 } } };
/* DiscoveryFSRed.set_host_state_priority_list FSMRedVariants.py:86-99 (K KD S SD U UD R RD) */
static const int FSMD_PRIORITY[8] = { 20, 20, 20, 20, 10, 10, 0, 0 };
/* order of action_space['action'] for red = red_actions list, EnterpriseScenarioGenerator.py:764-768, as FSM columns */
static const int RED_ASPACE_COLS[9] = { 0, 1, 2, 4, 5, 7, 3, 6, 8 };

/* BlueRewardMachine.py:35-65: [phase][subnet creation idx][LWF, ASF, RIA] */
static const int8_t REWARD[3][NSUB][3] = {
    { { -1, -3, -1 }, { -1, -1, -1 }, { -1, -3, -1 }, { -1, -1, -1 }, { 0, -5, -5 }, { -1, -1, -3 }, { -1, -1, -3 }, { -1, -1, -3 }, { 0, 0, -1 } },
    { { -2, -1, -3 }, { -10, 0, -10 }, { -1, -1, -1 }, { -1, -1, -1 }, { 0, 0, 0 }, { -1, -1, -3 }, { -1, -1, -3 }, { -1, -1, -3 }, { 0, 0, 0 } },
    { { -1, -3, -3 }, { -1, -1, -1 }, { -2, -1, -3 }, { -10, 0, -10 }, { 0, 0, 0 }, { -1, -1, -3 }, { -1, -1, -3 }, { -1, -1, -3 }, { 0, 0, 0 } } };

/* green allowed-subnet partner masks per phase, EnterpriseScenarioGenerator.py:281-306 (symmetric pairs) */
static uint16_t GREEN_ALLOWED[3][NSUB];
/* BlueFlatWrapper.py:267-302 comms policy adjacency per phase (creation-index bit masks) */
static uint16_t COMMS_ADJ[3][NSUB];
/* router tree parent (subnet idx; S_INT is the root, represented by root_internet_host_0), :388-416 */
static const int TREE_PARENT[NSUB] = { S_INT, S_RZA, S_INT, S_RZB, S_INT, S_INT, S_PAZ, S_PAZ, -1 };
/* server_host_0 knowledge links :431-462 (subnet idx lists, in dict-literal order; iteration is state.hosts order) */
static uint16_t LINKS[NSUB];
/* blue / red agent subnets :643-649, :769-776 */
static const int BLUE_SUBNETS[NBLUE][3] = { { S_RZA, -1, -1 }, { S_OZA, -1, -1 }, { S_RZB, -1, -1 }, { S_OZB, -1, -1 }, { S_PAZ, S_ADM, S_OFF } };
static const int RED_SUBNETS[NRED][3] = { { S_CON, -1, -1 }, { S_RZA, -1, -1 }, { S_OZA, -1, -1 }, { S_RZB, -1, -1 }, { S_OZB, -1, -1 }, { S_PAZ, S_ADM, S_OFF } };
static int tables_ready = 0;

static void pair(uint16_t *t, int a, int b) { t[a] |= 1u << b; t[b] |= 1u << a; }
static void init_tables(void)
{
    if (tables_ready) return;
    for (int i = 0; i < NSUB; i++) SORTED_POS[SORTED_SUBNETS[i]] = i;
    memset(GREEN_ALLOWED, 0, sizeof GREEN_ALLOWED);
    for (int p = 0; p < 3; p++) {
        uint16_t *t = GREEN_ALLOWED[p];
        const int hq[3] = { S_PAZ, S_ADM, S_OFF };
        for (int i = 0; i < 3; i++) { pair(t, hq[i], S_CON); pair(t, hq[i], S_RZA); pair(t, hq[i], S_RZB); }
        if (p == 0) { pair(t, S_RZA, S_CON); pair(t, S_OZA, S_RZA); pair(t, S_RZB, S_CON); pair(t, S_RZB, S_RZA); pair(t, S_OZB, S_RZB); }
        if (p == 1) { pair(t, S_RZB, S_CON); pair(t, S_OZB, S_RZB); }
        if (p == 2) { pair(t, S_RZA, S_CON); pair(t, S_OZA, S_RZA); }
        for (int s = 0; s < NSUB; s++) t[s] |= 1u << s; /* own subnet always allowed, SimulationController.py:756 */
    }
    memset(COMMS_ADJ, 0, sizeof COMMS_ADJ);
    for (int p = 0; p < 3; p++) {
        uint16_t *t = COMMS_ADJ[p];
        const int core[7] = { S_INT, S_ADM, S_OFF, S_PAZ, S_CON, S_RZA, S_RZB };
        for (int i = 0; i < 7; i++) for (int j = i + 1; j < 7; j++) pair(t, core[i], core[j]);
        pair(t, S_RZA, S_OZA); pair(t, S_RZB, S_OZB);
        if (p == 1) { const int rm[4] = { S_OZA, S_CON, S_RZB, S_INT }; for (int i = 0; i < 4; i++) { t[S_RZA] &= ~(1u << rm[i]); t[rm[i]] &= ~(1u << S_RZA); } }
        if (p == 2) { const int rm[4] = { S_OZB, S_CON, S_RZA, S_INT }; for (int i = 0; i < 4; i++) { t[S_RZB] &= ~(1u << rm[i]); t[rm[i]] &= ~(1u << S_RZB); } }
    }
    memset(LINKS, 0, sizeof LINKS);
    LINKS[S_CON] = 1u << S_RZA | 1u << S_RZB | 1u << S_PAZ;
    LINKS[S_RZA] = 1u << S_OZA | 1u << S_CON;   LINKS[S_OZA] = 1u << S_RZA;
    LINKS[S_RZB] = 1u << S_OZB | 1u << S_CON;   LINKS[S_OZB] = 1u << S_RZB;
    LINKS[S_PAZ] = 1u << S_ADM | 1u << S_OFF | 1u << S_CON;
    LINKS[S_ADM] = 1u << S_PAZ;                 LINKS[S_OFF] = 1u << S_PAZ;
    tables_ready = 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* data structures                                                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int pid; uint8_t kind, root; } proc_t;              /* Simulator/Process.py:12 */
typedef struct { uint8_t code, active, reliab; int pid; } svc_t;      /* Simulator/Service.py:4 */
typedef struct {
    uint8_t exists, subnet;
    int nproc; proc_t procs[MAXPROC];
    int nsvc; svc_t svcs[MAXSVC];
    uint8_t ev_proc, ev_net, old_proc, old_net;                       /* Simulator/HostEvents.py:7 */
    int n_evpid; int evpid[32];                                       /* pids carried by current process_creation events */
    int o_nproc; proc_t o_procs[MAXPROC]; int o_nsvc; svc_t o_svcs[MAXSVC]; /* Host.create_backup Host.py:316 */
    int o_nsess; struct { uint8_t agent; int id; } o_sess[4];
} host_t;
typedef struct {
    int id, host, pid; uint8_t type /*0 Session, 1 RedAbstractSession, 2 VelociraptorServer*/, root;
    uint8_t has_parent;    /* Session.parent is not None (Session.py:53): set by an exploit (ExploitAction.py:256), reset by a re-election
                              (RedSessionCheck.py:50-54: parent = the new primary's name, which is None); only Withdraw looks at it */
    uint8_t has_ports[NSLOT], ports[NSLOT];                           /* Shared/Session.py:125-133 */
} sess_t;
typedef struct { int code, host, subnet, session; } act_t;            /* an Action instance: type + parameters */
typedef struct { uint8_t key_ip; int16_t slot; uint8_t iface, subnet_if, syshost, sess; } obsent_t;
enum { SUCC_TRUE = 1, SUCC_UNKNOWN = 2, SUCC_FALSE = 3, SUCC_INPROG = 4 }; /* Shared/Enums.py:5 */
typedef struct {
    int n; obsent_t e[MAXOBS];
    int have_first;        /* observations list non-empty */
    int succ0; act_t act0; int has_act0;   /* 'success' / 'action' of observations[0] */
} obsset_t;
typedef struct { int16_t slot; uint8_t state, hn; } fsm_t;

struct cc4o_env {
    uint64_t seed, episode; int steps, rng_mode, red_variant;
    const cc4o_draw *tape; const uint32_t *seg_off; uint32_t n_segs;
    uint32_t seg_pos[NAG]; uint32_t reset_pos; uint32_t draw_ctr[NAG + 1]; uint32_t frac[NAG + 1];
    int32_t breakdown[3][3][NSUB];     /* reward events so far: [mission phase][LWF, ASF, RIA][subnet] (cc4o_reward_breakdown) */
    uint32_t blk_id[NAG + 1]; uint32_t blk[NAG + 1][4];   /* cached Philox block per stream */
    int err; char errmsg[256];
    /* scenario */
    int nuser[8], nserver[8];
    int order[NSLOT], norder;          /* state.hosts / state.ip_addresses iteration order */
    int ngreen; int green_host[MAXGREEN];
    int blue_parent[NBLUE];
    int used_pids[1024], n_used;
    /* dynamic */
    host_t hosts[NSLOT];
    sess_t *sess[NAG]; int nsess[NAG]; int sess_count[NAG];
    uint8_t active[NAG];
    uint16_t blocks[NSUB];
    int phase, step_count, done;
    act_t inprog[NAG]; int inprog_ticks[NAG];   /* code AC_NONE when nothing in flight */
    act_t cur_action[NAG];                       /* self.action[agent][0] of the tick */
    /* red knowledge: ActionSpace.py:170-212 */
    uint8_t ip_known[NRED][NSLOT], hn_known[NRED][NSLOT]; uint16_t sn_known[NRED];
    int n_ss[NRED]; int ss[NRED][MAXSS];
    /* FiniteStateRedAgent */
    int n_fsm[NRED]; fsm_t fsm[NRED][MAXFSM]; int fsm_step[NRED];
    obsset_t obs[NRED];
    /* blue sus pids: Session.py:154-165 */
    int n_sus[NBLUE]; struct { int16_t slot; int pid; } sus[NBLUE][MAXSUS];
    uint8_t msgs[NBLUE];
    int green_succ[MAXGREEN];
};

static void fail(cc4o_env *e, int flag, const char *msg)
{
    if (!(e->err & flag) || !e->errmsg[0]) snprintf(e->errmsg, sizeof e->errmsg, "tick %d: %s", e->step_count, msg);
    e->err |= flag;
}

/* ------------------------------------------------------------------------------------------------ */
/* RNG: tape replay or Philox4x32-10                                                                */
/* ------------------------------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* actor == NAG means "reset segment" */
static int tape_next(cc4o_env *e, int actor, int site, int kind, uint32_t n, uint64_t *val)
{
    uint32_t seg, *pos;
    if (actor == NAG) { seg = 0; pos = &e->reset_pos; }
    else { seg = 1 + (uint32_t)e->step_count * NAG + (uint32_t)actor; pos = &e->seg_pos[actor]; }
    if (seg >= e->n_segs) { fail(e, CC4O_E_TAPE_END, "tape segment out of range"); *val = 0; return -1; }
    uint32_t i = e->seg_off[seg] + *pos;
    if (i >= e->seg_off[seg + 1]) {
        char b[128]; snprintf(b, sizeof b, "tape exhausted: actor %d site %d n %u", actor, site, n);
        fail(e, CC4O_E_TAPE_END, b); *val = 0; return -1;
    }
    const cc4o_draw *d = &e->tape[i];
    if (d->site != site || d->kind != kind || d->n != n) {
        char b[160]; snprintf(b, sizeof b, "tape mismatch: actor %d wants site %d kind %d n %u, tape has site %d kind %d n %u (pos %u)",
                              actor, site, kind, n, d->site, d->kind, d->n, *pos);
        fail(e, CC4O_E_TAPE_SITE, b);
    }
    (*pos)++;
    *val = d->val;
    return 0;
}

/* Production RNG of the engine (not the reference's): every (actor, tick) owns a stream of 32-bit words; word j is
 * element j&3 of the Philox4x32-10 block with key (seed, episode) and counter (j>>2, actor, tick+1, episode).
 * A draw consumes one word: an integer in [0,n) is the high half of word*n (the low half is kept per actor, so that a
 * second, dependent integer can be cut from the same word: rng_int_more), a double is word / 2^32. */
static uint32_t rng_word(cc4o_env *e, int actor)
{
    uint32_t j = e->draw_ctr[actor]++;
    if (e->blk_id[actor] != (j >> 2)) {
        uint32_t c[4];
        c[0] = j >> 2;
        c[1] = (uint32_t)actor;
        c[2] = actor == NAG ? 0u : (uint32_t)e->step_count + 1u;
        c[3] = (uint32_t)e->episode;
        philox4x32_10(c, (uint32_t)e->seed, (uint32_t)(e->seed >> 32) ^ (uint32_t)(e->episode >> 32));
        memcpy(e->blk[actor], c, sizeof c);
        e->blk_id[actor] = j >> 2;
    }
    return e->blk[actor][j & 3u];
}

static uint32_t rng_int(cc4o_env *e, int actor, int site, uint32_t n)
{
    if (e->rng_mode == 1) {
        uint64_t v; tape_next(e, actor, site, 0, n, &v);
        return v < n ? (uint32_t)v : 0;
    }
    if (n == 1 && site == ST_FSM_PARAM) return 0;       /* a one-option parameter pick spends no word */
    uint64_t p = (uint64_t)rng_word(e, actor) * n;
    e->frac[actor] = (uint32_t)p;
    return (uint32_t)(p >> 32);
}

/* a further integer that depends on the previous integer draw of the same actor (the reference draws twice; the tape
 * has both draws): production mode cuts it from the unused low half of that word instead of spending a new one */
static uint32_t rng_int_more(cc4o_env *e, int actor, int site, uint32_t n)
{
    if (e->rng_mode == 1) return rng_int(e, actor, site, n);
    uint64_t p = (uint64_t)e->frac[actor] * n;
    e->frac[actor] = (uint32_t)p;
    return (uint32_t)(p >> 32);
}

static double rng_f64(cc4o_env *e, int actor, int site, uint32_t n)
{
    if (e->rng_mode == 1) {
        uint64_t v; tape_next(e, actor, site, 1, n, &v);
        double d; memcpy(&d, &v, 8); return d;
    }
    return (double)rng_word(e, actor) * (1.0 / 4294967296.0);
}

/* ------------------------------------------------------------------------------------------------ */
/* helpers                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
static int slot_subnet(int slot) { return slot == SLOT_INTERNET ? S_INT : slot >= SLOT_ROUTER ? slot - SLOT_ROUTER : slot >> 4; }
static int slot_is_server(int slot) { return slot < 128 && (slot & 15) < 6; }
static int red_of_subnet(int sn)
{
    for (int r = 0; r < NRED; r++) for (int i = 0; i < 3; i++) if (RED_SUBNETS[r][i] == sn) return r;
    return -1;
}
static int red_allowed(int r, int sn) { for (int i = 0; i < 3; i++) if (RED_SUBNETS[r][i] == sn) return 1; return 0; }

/* Host.create_pid Host.py:198-200 */
static int create_pid(cc4o_env *e, int actor, int slot)
{
    host_t *h = &e->hosts[slot];
    int mx = 0;
    for (int i = 0; i < h->nproc; i++) if (h->procs[i].pid > mx) mx = h->procs[i].pid;
    return mx + 1 + (int)rng_int(e, actor, ST_CREATE_PID, 9);
}
static void add_proc(cc4o_env *e, int slot, int pid, int kind, int root)
{
    host_t *h = &e->hosts[slot];
    if (h->nproc >= MAXPROC) { fail(e, CC4O_E_CAPACITY, "process table overflow"); return; }
    h->procs[h->nproc].pid = pid; h->procs[h->nproc].kind = (uint8_t)kind; h->procs[h->nproc].root = (uint8_t)root; h->nproc++;
}
static int find_proc(host_t *h, int pid) { for (int i = 0; i < h->nproc; i++) if (h->procs[i].pid == pid) return i; return -1; }
static void del_proc(host_t *h, int idx) { for (int i = idx; i + 1 < h->nproc; i++) h->procs[i] = h->procs[i + 1]; h->nproc--; }

static int find_sess(cc4o_env *e, int ag, int id) { for (int i = 0; i < e->nsess[ag]; i++) if (e->sess[ag][i].id == id) return i; return -1; }
static void del_sess(cc4o_env *e, int ag, int idx) { for (int i = idx; i + 1 < e->nsess[ag]; i++) e->sess[ag][i] = e->sess[ag][i + 1]; e->nsess[ag]--; }
/* State.add_session State.py:305-323 + Host.add_session Host.py:189-196.  pid < 0 => create one (actor draws). */
static int add_session(cc4o_env *e, int actor, int ag, int host, int type, int root, int pid, int pkind, int ident)
{
    if (e->nsess[ag] >= MAXSESS) { fail(e, CC4O_E_CAPACITY, "session table overflow"); return -1; }
    if (ident < 0) {
        ident = 0;
        if (e->nsess[ag] > 0) { int mx = -1; for (int i = 0; i < e->nsess[ag]; i++) if (e->sess[ag][i].id > mx) mx = e->sess[ag][i].id; ident = mx + 1; }
    }
    e->sess_count[ag]++;
    sess_t *s = &e->sess[ag][e->nsess[ag]++];
    memset(s, 0, sizeof *s);
    s->id = ident; s->host = host; s->type = (uint8_t)type; s->root = (uint8_t)root;
    if (pid < 0) { pid = create_pid(e, actor, host); add_proc(e, host, pid, pkind, 0); }
    s->pid = pid;
    return ident;
}
/* iteration order of state.sessions (scenario.agents order: blue, green, red) */
static int agent_iter(cc4o_env *e, int k)
{
    if (k < NBLUE) return A_BLUE0 + k;
    if (k < NBLUE + e->ngreen) return A_GREEN0 + (k - NBLUE);
    return A_RED0 + (k - NBLUE - e->ngreen);
}
static int n_agents(cc4o_env *e) { return NBLUE + e->ngreen + NRED; }
/* State.get_session_from_pid State.py:420-444 */
static int session_from_pid(cc4o_env *e, int host, int pid, int *ag_out)
{
    for (int k = 0; k < n_agents(e); k++) {
        int ag = agent_iter(e, k);
        for (int i = 0; i < e->nsess[ag]; i++) if (e->sess[ag][i].pid == pid && e->sess[ag][i].host == host) { *ag_out = ag; return i; }
    }
    return -1;
}
static int red_on_host(cc4o_env *e, int r, int host) { for (int i = 0; i < e->nsess[A_RED0 + r]; i++) if (e->sess[A_RED0 + r][i].host == host) return 1; return 0; }
/* RemoteAction.blocking_host Action.py:127-136 (subnet form; the per-host form is never populated in CC4) */
static int blocking(cc4o_env *e, int src_slot, int other_slot) { return (e->blocks[slot_subnet(other_slot)] >> slot_subnet(src_slot)) & 1; }

/* route src -> dst on the router tree (Action.py:101-117); returns hop count, hosts in out[] */
static int route(int src, int dst, int *out)
{
    int n = 0;
    if (src == dst) { out[n++] = src; return n; }
    int a = slot_subnet(src), b = slot_subnet(dst);
    int up[8], nu = 0, dn[8], nd = 0;
    /* ancestors chain as subnets */
    int ca[8], na = 0, cb[8], nb = 0;
    for (int x = a; x >= 0; x = TREE_PARENT[x]) ca[na++] = x;
    for (int x = b; x >= 0; x = TREE_PARENT[x]) cb[nb++] = x;
    int ia = na - 1, ib = nb - 1;
    while (ia > 0 && ib > 0 && ca[ia - 1] == cb[ib - 1]) { ia--; ib--; }
    /* ca[ia] == cb[ib] is the lowest common ancestor */
    for (int i = 0; i <= ia; i++) up[nu++] = ca[i];
    for (int i = ib - 1; i >= 0; i--) dn[nd++] = cb[i];
    /* src itself: if src is a router / the internet host it is the first tree node */
    int src_is_node = (src >= SLOT_ROUTER), dst_is_node = (dst >= SLOT_ROUTER);
    if (!src_is_node) out[n++] = src;
    for (int i = 0; i < nu; i++) out[n++] = up[i] == S_INT ? SLOT_INTERNET : SLOT_ROUTER + up[i];
    for (int i = 0; i < nd; i++) out[n++] = dn[i] == S_INT ? SLOT_INTERNET : SLOT_ROUTER + dn[i];
    if (!dst_is_node) out[n++] = dst;
    return n;
}

/* ---- per-agent observation bookkeeping (what ObservationSet/ActionSpace/FSM consume) ---- */
static void obs_entry(cc4o_env *e, int r, int key_ip, int slot, int iface, int subnet_if, int syshost, int sess)
{
    obsset_t *o = &e->obs[r];
    /* ActionSpace.update ActionSpace.py:170-212 */
    if (syshost) e->hn_known[r][slot] = 1;
    if (iface) e->ip_known[r][slot] = 1;
    if (subnet_if) e->sn_known[r] |= 1u << slot_subnet(slot);
    for (int i = 0; i < o->n; i++)
        if (o->e[i].key_ip == key_ip && o->e[i].slot == slot) {
            o->e[i].iface |= iface; o->e[i].subnet_if |= subnet_if; o->e[i].syshost |= syshost; o->e[i].sess |= sess;
            return;
        }
    if (o->n >= MAXOBS) { fail(e, CC4O_E_CAPACITY, "observation entry overflow"); return; }
    obsent_t *x = &o->e[o->n++];
    x->key_ip = (uint8_t)key_ip; x->slot = (int16_t)slot; x->iface = (uint8_t)iface; x->subnet_if = (uint8_t)subnet_if;
    x->syshost = (uint8_t)syshost; x->sess = (uint8_t)sess;
}
static void know_session(cc4o_env *e, int r, int id)
{
    for (int i = 0; i < e->n_ss[r]; i++) if (e->ss[r][i] == id) return;
    if (e->n_ss[r] >= MAXSS) { fail(e, CC4O_E_CAPACITY, "server_session overflow"); return; }
    e->ss[r][e->n_ss[r]++] = id;
}
/* append an Observation to observation[agent]: records observations[0] success/action */
static void obs_push(cc4o_env *e, int r, int succ, const act_t *act)
{
    obsset_t *o = &e->obs[r];
    if (!o->have_first) {
        o->have_first = 1; o->succ0 = succ; o->has_act0 = act != NULL;
        if (act) o->act0 = *act;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* reset: EnterpriseScenarioGenerator.create_scenario :123-169 + State.__init__ State.py:66-148      */
/* ------------------------------------------------------------------------------------------------ */
static int gen_pid(cc4o_env *e)
{   /* _generate_pid :563-578 (rejection-sampled unique 4-digit pid) */
    for (;;) {
        int pid = 1000 + (int)rng_int(e, NAG, ST_R_PID, 9000);
        int dup = 0;
        for (int i = 0; i < e->n_used; i++) if (e->used_pids[i] == pid) { dup = 1; break; }
        if (e->err & (CC4O_E_TAPE_END)) return pid;
        if (!dup) { if (e->n_used < 1024) e->used_pids[e->n_used++] = pid; return pid; }
    }
}
static void gen_host(cc4o_env *e, int slot)
{   /* _generate_linux_host :470-528, _generate_linux_host_services :530-561, _generate_linux_host_processes :580-629 */
    host_t *h = &e->hosts[slot];
    memset(h, 0, sizeof *h);
    h->exists = 1; h->subnet = (uint8_t)slot_subnet(slot);
    e->order[e->norder++] = slot;
    if (slot >= SLOT_ROUTER) return;
    int sn = slot_subnet(slot);
    h->svcs[h->nsvc].code = K_SSHD; h->svcs[h->nsvc].pid = gen_pid(e); h->nsvc++;
    if (sn == S_OZA || sn == S_OZB) { h->svcs[h->nsvc].code = K_OT; h->svcs[h->nsvc].pid = gen_pid(e); h->nsvc++; }
    int opt[3] = { K_APACHE2, K_MYSQLD, K_SMTP }, optpid[3], nopt = 3;
    for (int i = 0; i < 3; i++) optpid[i] = gen_pid(e);
    int nadd = (int)rng_int(e, NAG, ST_R_NADDON, 4);
    for (int i = 0; i < nadd; i++) {
        int c = (int)rng_int(e, NAG, ST_R_ADDON, (uint32_t)nopt);
        h->svcs[h->nsvc].code = (uint8_t)opt[c]; h->svcs[h->nsvc].pid = optpid[c]; h->nsvc++;
        for (int j = c; j + 1 < nopt; j++) { opt[j] = opt[j + 1]; optpid[j] = optpid[j + 1]; }
        nopt--;
    }
    for (int i = 0; i < h->nsvc; i++) { h->svcs[i].active = 1; h->svcs[i].reliab = 100; add_proc(e, slot, h->svcs[i].pid, h->svcs[i].code, 0); }
}

int cc4o_reset(cc4o_env *e)
{
    init_tables();
    e->err = 0; e->errmsg[0] = 0;
    e->reset_pos = 0; memset(e->seg_pos, 0, sizeof e->seg_pos); memset(e->draw_ctr, 0, sizeof e->draw_ctr); memset(e->blk_id, 0xFF, sizeof e->blk_id);
    memset(e->breakdown, 0, sizeof e->breakdown);
    e->n_used = 0; e->norder = 0; e->ngreen = 0;
    memset(e->hosts, 0, sizeof e->hosts);
    for (int a = 0; a < NAG; a++) { e->nsess[a] = 0; e->sess_count[a] = 0; e->active[a] = 0; e->inprog[a].code = AC_NONE; e->inprog_ticks[a] = 0; e->cur_action[a].code = AC_SLEEP; }
    memset(e->blocks, 0, sizeof e->blocks);
    e->phase = 0; e->step_count = 0; e->done = 0;
    memset(e->ip_known, 0, sizeof e->ip_known); memset(e->hn_known, 0, sizeof e->hn_known); memset(e->sn_known, 0, sizeof e->sn_known);
    memset(e->n_ss, 0, sizeof e->n_ss); memset(e->n_fsm, 0, sizeof e->n_fsm); memset(e->fsm_step, 0, sizeof e->fsm_step);
    memset(e->obs, 0, sizeof e->obs); memset(e->n_sus, 0, sizeof e->n_sus); memset(e->msgs, 0, sizeof e->msgs);

    /* hosts, :312-371: per subnet router, users, servers; internet host last */
    for (int sn = 0; sn < 8; sn++) {
        gen_host(e, SLOT_ROUTER + sn);
        e->nuser[sn] = 3 + (int)rng_int(e, NAG, ST_R_NUSER, 8);
        for (int i = 0; i < e->nuser[sn]; i++) gen_host(e, sn * 16 + 6 + i);
        e->nserver[sn] = 1 + (int)rng_int(e, NAG, ST_R_NSERVER, 6);
        for (int i = 0; i < e->nserver[sn]; i++) gen_host(e, sn * 16 + i);
    }
    gen_host(e, SLOT_INTERNET);
    /* blue agents :631-698: parent host choice over allowed_hosts (router, users, servers per subnet) */
    int allowed[NBLUE][64], nallowed[NBLUE];
    for (int b = 0; b < NBLUE; b++) {
        int n = 0;
        for (int i = 0; i < 3 && BLUE_SUBNETS[b][i] >= 0; i++) {
            int sn = BLUE_SUBNETS[b][i];
            allowed[b][n++] = SLOT_ROUTER + sn;
            for (int k = 0; k < e->nuser[sn]; k++) allowed[b][n++] = sn * 16 + 6 + k;
            for (int k = 0; k < e->nserver[sn]; k++) allowed[b][n++] = sn * 16 + k;
        }
        nallowed[b] = n;
        e->blue_parent[b] = allowed[b][rng_int(e, NAG, ST_R_BLUE_PARENT, (uint32_t)n)];
    }
    /* green agents :700-749: one per user host in creation order */
    for (int sn = 0; sn < 8; sn++) for (int k = 0; k < e->nuser[sn]; k++) e->green_host[e->ngreen++] = sn * 16 + 6 + k;
    /* red agents :751-812 */
    int red_start[NRED];
    for (int r = 0; r < NRED; r++) {
        int ns = 0; while (ns < 3 && RED_SUBNETS[r][ns] >= 0) ns++;
        int sn = RED_SUBNETS[r][rng_int(e, NAG, ST_R_RED_SUBNET, (uint32_t)ns)];
        int c = (int)rng_int(e, NAG, ST_R_RED_HOST, (uint32_t)(e->nuser[sn] + e->nserver[sn]));
        red_start[r] = c < e->nuser[sn] ? sn * 16 + 6 + c : sn * 16 + (c - e->nuser[sn]);
    }
    /* State.__init__ State.py:108-134: sessions per agent in scenario.agents order; parentless first */
    for (int b = 0; b < NBLUE; b++) {
        add_session(e, NAG, A_BLUE0 + b, e->blue_parent[b], 2, 0, -1, K_SESSION, -1);
        for (int i = 0; i < nallowed[b]; i++) if (allowed[b][i] != e->blue_parent[b]) add_session(e, NAG, A_BLUE0 + b, allowed[b][i], 0, 0, -1, K_SESSION, -1);
        e->active[A_BLUE0 + b] = 1;
    }
    for (int g = 0; g < e->ngreen; g++) { add_session(e, NAG, A_GREEN0 + g, e->green_host[g], 0, 0, -1, K_SESSION, -1); e->active[A_GREEN0 + g] = 1; }
    add_session(e, NAG, A_RED0, red_start[0], 1, 0, -1, K_REDABS, -1);
    e->active[A_RED0] = 1;
    /* Host.create_backup Host.py:316-371 */
    for (int s = 0; s < NSLOT; s++) {
        host_t *h = &e->hosts[s];
        if (!h->exists) continue;
        h->o_nproc = h->nproc; memcpy(h->o_procs, h->procs, sizeof h->procs);
        h->o_nsvc = h->nsvc; memcpy(h->o_svcs, h->svcs, sizeof h->svcs);
        h->o_nsess = 0;
        for (int k = 0; k < n_agents(e); k++) {
            int ag = agent_iter(e, k);
            for (int i = 0; i < e->nsess[ag]; i++) if (e->sess[ag][i].host == s) { h->o_sess[h->o_nsess].agent = (uint8_t)ag; h->o_sess[h->o_nsess].id = e->sess[ag][i].id; h->o_nsess++; }
        }
    }
    /* initial knowledge (AgentInterface.set_init_obs AgentInterface.py:110-117 with the OSINT of :784-785)
       and red_agent_0's initial observation (SimulationController.reset_observation :767-773) */
    for (int r = 0; r < NRED; r++) {
        int s = red_start[r];
        e->hn_known[r][s] = 1; e->ip_known[r][s] = 1; e->sn_known[r] |= 1u << slot_subnet(s);
    }
    know_session(e, 0, 0);
    obs_entry(e, 0, 0, red_start[0], 1, 1, 1, 1);
    e->obs[0].have_first = 1; e->obs[0].succ0 = SUCC_UNKNOWN; e->obs[0].has_act0 = 0;
    if (e->rng_mode == 1 && e->reset_pos != e->seg_off[1] - e->seg_off[0]) fail(e, CC4O_E_TAPE_LEFT, "reset segment not fully consumed");
    return e->err ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* blue action index decoding: BlueFixedActionWrapper._populate_action_space :233-309               */
/* ------------------------------------------------------------------------------------------------ */
static int blue_nsub(int b) { return b == 4 ? 3 : 1; }
static int blue_sorted_subnet(int b, int i)
{   /* self._agent_metadata[agent]["subnets"] = sorted names */
    if (b < 4) return BLUE_SUBNETS[b][0];
    static const int s4[3] = { S_ADM, S_OFF, S_PAZ };
    return s4[i];
}
static int blue_has_session(cc4o_env *e, int b, int slot)
{
    for (int i = 0; i < e->nsess[A_BLUE0 + b]; i++) if (e->sess[A_BLUE0 + b][i].host == slot) return 1;
    return 0;
}
/* decode with the state AT RESET (mask and Action objects are built once per episode in wrapper.reset) */
static act_t blue_decode(cc4o_env *e, int b, int idx, int *mask_out)
{
    act_t a = { AC_SLEEP, -1, -1, 0 };
    int n = blue_nsub(b), H = 16 * n, T = 8 * n, m = 1;
    int i = idx;
    int code = AC_SLEEP, hostidx = -1, tidx = -1;
    if (i < 0) { m = 0; }
    else if (i < H) { code = AC_ANALYSE; hostidx = i; }
    else if (i == H) { code = AC_MONITOR; }
    else if (i < 2 * H + 1) { code = AC_REMOVE; hostidx = i - H - 1; }
    else if (i < 3 * H + 1) { code = AC_RESTORE; hostidx = i - 2 * H - 1; }
    else if (i == 3 * H + 1) { code = AC_SLEEP; }
    else if (i < 3 * H + 2 + T) { code = AC_ALLOW; tidx = i - 3 * H - 2; }
    else if (i < 3 * H + 2 + 2 * T) { code = AC_BLOCK; tidx = i - 3 * H - 2 - T; }
    else if (i < 4 * H + 2 + 2 * T) { code = AC_DECOY; hostidx = i - 3 * H - 2 - 2 * T; }
    else { m = 0; }
    if (hostidx >= 0) {
        int sn = blue_sorted_subnet(b, hostidx / 16), k = hostidx % 16;
        int slot = sn * 16 + k;
        /* the mask is computed from the scenario (host exists; blue has a session on every host of its subnets at reset) */
        int exists = k < 6 ? k < e->nserver[sn] : (k - 6) < e->nuser[sn];
        if (!exists) { code = AC_SLEEP; m = 0; } else a.host = slot;
    }
    if (tidx >= 0) {
        int dst = blue_sorted_subnet(b, tidx / 8), j = tidx % 8, src = -1;
        for (int p = 0, c = 0; p < NSUB; p++) { if (SORTED_SUBNETS[p] == dst) continue; if (c == j) { src = SORTED_SUBNETS[p]; break; } c++; }
        a.subnet = dst; a.host = src; /* to_subnet = dst, from_subnet = src (kept in .host) */
    }
    a.code = code;
    if (mask_out) *mask_out = m;
    return a;
}
int cc4o_action_mask(cc4o_env *e, uint8_t *mask)
{
    memset(mask, 0, NBLUE * CC4O_MASK_LEN);
    for (int b = 0; b < NBLUE; b++) {
        int n = b == 4 ? 242 : 82;
        for (int i = 0; i < n; i++) { int m; blue_decode(e, b, i, &m); mask[b * CC4O_MASK_LEN + i] = (uint8_t)m; }
    }
    return 0;
}
static int action_duration(int code)
{
    switch (code) {
    case AC_ANALYSE: return 2;   /* Analyse.py:38 */
    case AC_REMOVE: return 3;    /* Remove.py:40 */
    case AC_RESTORE: return 5;   /* Restore.py:36 */
    case AC_DECOY: return 2;     /* DeployDecoy.py:31 */
    case AC_STEALTH: return 3;   /* DiscoverNetworkServices.py:136 */
    case AC_DD: return 2;        /* DiscoverDeception.py:39 */
    case AC_EXPLOIT: return 4;   /* ExploitRemoteService.py:147 */
    case AC_PRIVESC: return 2;   /* PrivilegeEscalate.py:82 */
    case AC_IMPACT: return 2;    /* Impact.py:36 */
    case AC_DEGRADE: return 2;   /* DegradeServices.py:36 */
    default: return 1;           /* Action.py:15 */
    }
}
static int action_has_session(int code) { return !(code == AC_SLEEP || code == AC_INVALID); }

/* ------------------------------------------------------------------------------------------------ */
/* FiniteStateRedAgent.get_action FiniteStateRedAgent.py:58-122                                     */
/* ------------------------------------------------------------------------------------------------ */
static int fsm_find(cc4o_env *e, int r, int slot) { for (int i = 0; i < e->n_fsm[r]; i++) if (e->fsm[r][i].slot == slot) return i; return -1; }

static void fsm_transition(cc4o_env *e, int r)
{   /* _host_state_transition :124-167 */
    obsset_t *o = &e->obs[r];
    if (!o->has_act0 || o->succ0 == SUCC_INPROG) return;
    int col = -1;
    for (int c = 0; c < 9; c++) if (FSM_COL_CODE[c] == o->act0.code) col = c;
    if (col < 0) return;
    int targets[MAXFSM], nt = 0;
    if (col == 1 || col == 2 || col == 3 || col == 4) { int i = fsm_find(e, r, o->act0.host); if (i >= 0) targets[nt++] = i; }
    else if (col == 5 || col == 6 || col == 7) { for (int i = 0; i < e->n_fsm[r]; i++) if (e->fsm[r][i].hn && e->fsm[r][i].slot == o->act0.host) { targets[nt++] = i; break; } }
    else if (col == 0) { for (int i = 0; i < e->n_fsm[r]; i++) if (e->fsm[r][i].slot >= 0 && slot_subnet(e->fsm[r][i].slot) == o->act0.subnet) targets[nt++] = i; }
    else return; /* Withdraw: never selected (p = 0) */
    for (int t = 0; t < nt; t++) {
        fsm_t *f = &e->fsm[r][targets[t]];
        int next = o->succ0 == SUCC_TRUE ? FSM_SUCC[f->state][col] : FSM_FAIL[f->state][col];
        if (next == F_U) { next = F_F; if (f->slot >= 0 && red_allowed(r, slot_subnet(f->slot))) next = F_U; }
        if (next < 0) next = f->state;
        f->state = (uint8_t)next;
    }
}
static void fsm_process_obs(cc4o_env *e, int r)
{   /* _process_new_observations :190-250 */
    obsset_t *o = &e->obs[r];
    for (int i = 0; i < o->n; i++) {
        obsent_t *x = &o->e[i];
        int hn = 0, slot = -1;
        if (!x->key_ip) hn = 1; else if (x->syshost) hn = 1;
        if (x->key_ip) slot = x->slot; else if (x->iface) slot = x->slot;
        if (slot < 0 && hn) { for (int j = 0; j < e->n_fsm[r]; j++) if (e->fsm[r][j].hn && e->fsm[r][j].slot == x->slot) { slot = x->slot; break; } }
        int idx = fsm_find(e, r, slot);   /* slot == -1 models the None key */
        if (idx < 0) {
            if (e->n_fsm[r] >= MAXFSM) { fail(e, CC4O_E_CAPACITY, "fsm host list overflow"); return; }
            idx = e->n_fsm[r]++;
            e->fsm[r][idx].slot = (int16_t)slot; e->fsm[r][idx].state = e->fsm_step[r] == 0 ? F_U : F_K; e->fsm[r][idx].hn = (uint8_t)hn;
            if (slot < 0) fail(e, CC4O_E_UNREACHABLE, "fsm host_states[None]");
        }
        if (slot >= 0 && hn && !e->fsm[r][idx].hn) e->fsm[r][idx].hn = 1;
    }
}
static void fsm_session_removal(cc4o_env *e, int r)
{   /* _session_removal_state_change :169-188 */
    obsset_t *o = &e->obs[r];
    for (int i = 0; i < e->n_fsm[r]; i++) {
        fsm_t *f = &e->fsm[r][i];
        if (!(f->state == F_U || f->state == F_UD || f->state == F_R || f->state == F_RD)) continue;
        int seen = 0;
        for (int j = 0; j < o->n; j++) if (o->e[j].sess) {
            if (!o->e[j].iface) { fail(e, CC4O_E_UNREACHABLE, "obs entry with Sessions but no Interface"); continue; }
            if (o->e[j].slot == f->slot) seen = 1;
        }
        if (!seen) f->state = F_KD;
    }
}

static act_t red_policy(cc4o_env *e, int r)
{
    int actor = A_RED0 + r;
    act_t a = { AC_SLEEP, -1, -1, 0 };
    obsset_t *o = &e->obs[r];
    fsm_transition(e, r);
    fsm_process_obs(e, r);
    fsm_session_removal(e, r);
    e->fsm_step[r]++;
    if (o->succ0 == SUCC_INPROG) return a;
    /* _choose_host_and_action :296-336 */
    int known[MAXFSM], nk = 0;
    for (int i = 0; i < e->n_fsm[r]; i++) if (e->fsm[r][i].state != F_F) known[nk++] = i;
    if (nk == 0) { fail(e, CC4O_E_UNREACHABLE, "fsm has no known host"); return a; }
    fsm_t *f;
    if (e->red_variant == 0) f = &e->fsm[r][known[rng_int(e, actor, ST_FSM_HOST, (uint32_t)nk)]];
    else {
        /* DiscoveryFSRed: FiniteStateRedAgent._choose_host :252-293 with a host-state priority list and the server bias */
        int states[8], ns = 0, total = 0;
        for (int i = 0; i < nk && ns < 8; i++) {      /* states in order of first appearance among the known hosts */
            int st = e->fsm[r][known[i]].state, seen = 0;
            for (int j = 0; j < ns; j++) seen |= states[j] == st;
            if (!seen) { states[ns++] = st; total += FSMD_PRIORITY[st]; }
        }
        int pick;
        if (total > 0) {
            /* probs = [(p / base) * p_multiplier]; numpy choice(p=): cdf = cumsum(probs); cdf /= cdf[-1]; searchsorted(cdf, u, 'right') */
            volatile double mult = 1.0 / ((double)total / 100.0);
            double cdfv[8]; volatile double acc = 0.0;
            for (int i = 0; i < ns; i++) { volatile double pi = ((double)FSMD_PRIORITY[states[i]] / 100.0) * mult; acc = acc + pi; cdfv[i] = acc; }
            double u = rng_f64(e, actor, ST_FSMD_STATE_P, (uint32_t)ns);
            for (pick = 0; pick < ns; pick++) if (!(cdfv[pick] / cdfv[ns - 1] <= u)) break;
            if (pick >= ns) pick = ns - 1;
        } else pick = (int)rng_int(e, actor, ST_FSMD_STATE_U, (uint32_t)ns);
        int opts[MAXFSM], no = 0, srv[MAXFSM], nsrv = 0, oth[MAXFSM], noth = 0;
        for (int i = 0; i < nk; i++) {
            fsm_t *g = &e->fsm[r][known[i]];
            if (g->state != states[pick]) continue;
            opts[no++] = known[i];
            if (g->hn && g->slot >= 0 && g->slot < 128 && (g->slot & 15) < 6) srv[nsrv++] = known[i]; else oth[noth++] = known[i];   /* 'server' in hostname */
        }
        int idx;
        if (no > 1) {
            if (nsrv > 0) {
                double roll = rng_f64(e, actor, ST_FSMD_ROLL, 0);
                if (roll <= 0.75) idx = srv[rng_int(e, actor, ST_FSMD_HOST_SRV, (uint32_t)nsrv)];
                else if (nsrv != no) idx = oth[rng_int(e, actor, ST_FSMD_HOST_OTHER, (uint32_t)noth)];
                else idx = srv[rng_int(e, actor, ST_FSMD_HOST_SRV2, (uint32_t)nsrv)];
            } else idx = opts[rng_int(e, actor, ST_FSMD_HOST_ALL, (uint32_t)no)];
        } else idx = opts[rng_int(e, actor, ST_FSM_HOST, (uint32_t)no)];
        f = &e->fsm[r][idx];
    }
    const double (*PROB)[9] = FSM_PROB[e->red_variant];
    int cols[9], nc = 0; double p[9];
    for (int i = 0; i < 9; i++) { int c = RED_ASPACE_COLS[i]; if (PROB[f->state][c] >= 0) { cols[nc] = c; p[nc] = PROB[f->state][c]; nc++; } }
    double u = rng_f64(e, actor, ST_FSM_ACTION, (uint32_t)nc);
    double cdf = 0, tot = 0; int pick = 0;
    for (int i = 0; i < nc; i++) tot += p[i];
    /* numpy Generator.choice(p=): cdf = cumsum(p); cdf /= cdf[-1]; idx = searchsorted(cdf, u, side='right') */
    for (pick = 0; pick < nc; pick++) { cdf += p[pick]; if (!(cdf / tot <= u)) break; }
    if (pick >= nc) pick = nc - 1;
    int col = cols[pick];
    a.code = FSM_COL_CODE[col];
    /* parameters in constructor-signature order; every non-hostname/ip parameter is one choice() :317-334 */
    #define PICK_SESSION() do { if (e->n_ss[r] == 0) { fail(e, CC4O_E_UNREACHABLE, "no known session"); } else a.session = e->ss[r][rng_int(e, actor, ST_FSM_PARAM, (uint32_t)e->n_ss[r])]; } while (0)
    #define PICK_AGENT() (void)rng_int(e, actor, ST_FSM_PARAM, 1)
    if (col == 0) {        /* DiscoverRemoteSystems(subnet, session, agent) */
        int opts[NSUB], no = 0;
        for (int s = 0; s < NSUB; s++) if (e->sn_known[r] >> s & 1) opts[no++] = s;
        a.subnet = opts[rng_int(e, actor, ST_FSM_PARAM, (uint32_t)no)];
        PICK_SESSION(); PICK_AGENT();
    } else if (col == 1 || col == 2 || col == 3) { /* (session, agent, ip_address) */
        PICK_SESSION(); PICK_AGENT(); a.host = f->slot;
    } else if (col == 4) { /* ExploitRemoteService(ip_address, session, agent) */
        a.host = f->slot; PICK_SESSION(); PICK_AGENT();
    } else {               /* PrivilegeEscalate / Impact / DegradeServices (hostname, session, agent) */
        if (!f->hn) fail(e, CC4O_E_UNREACHABLE, "hostname parameter unknown to the FSM");
        a.host = f->slot; PICK_SESSION(); PICK_AGENT();
    }
    /* replace_action_if_invalid SimulationController.py:1068-1112 */
    if ((col >= 1 && col <= 4) && !e->ip_known[r][a.host]) a.code = AC_INVALID;
    if (col >= 5 && !e->hn_known[r][a.host]) a.code = AC_INVALID;
    return a;
}

/* ------------------------------------------------------------------------------------------------ */
/* action execution                                                                                 */
/* ------------------------------------------------------------------------------------------------ */
static void ev_procpid(cc4o_env *e, int slot, int pid)
{
    host_t *h = &e->hosts[slot];
    h->ev_proc = 1;
    if (pid >= 0) { if (h->n_evpid < 32) h->evpid[h->n_evpid++] = pid; else fail(e, CC4O_E_CAPACITY, "event pid overflow"); }
}

/* Monitor.execute Monitor.py:35-74 */
static void do_monitor(cc4o_env *e, int b)
{
    for (int i = 0; i < 3 && BLUE_SUBNETS[b][i] >= 0; i++) {
        int sn = BLUE_SUBNETS[b][i];
        for (int k = -1; k < 16; k++) {
            int slot = k < 0 ? SLOT_ROUTER + sn : sn * 16 + k;
            host_t *h = &e->hosts[slot];
            if (!h->exists) continue;
            h->old_net = h->ev_net; h->ev_net = 0;
            for (int j = 0; j < h->n_evpid; j++) {
                if (e->n_sus[b] >= MAXSUS) { fail(e, CC4O_E_CAPACITY, "sus pid overflow"); break; }
                e->sus[b][e->n_sus[b]].slot = (int16_t)slot; e->sus[b][e->n_sus[b]].pid = h->evpid[j]; e->n_sus[b]++;
            }
            h->old_proc = h->ev_proc; h->ev_proc = 0; h->n_evpid = 0;
        }
    }
}

/* StopProcess.execute_targeteted_local_action + kill_process StopProcess.py:23-58 */
static int stop_process(cc4o_env *e, int actor, int slot, int pid, int stop_all)
{
    host_t *h = &e->hosts[slot];
    int pi = find_proc(h, pid);
    if (pi < 0) return 0;
    if (!stop_all && h->procs[pi].root) return 0;
    proc_t p = h->procs[pi];
    int ag = -1, si = session_from_pid(e, slot, pid, &ag);
    del_proc(h, pi);
    int svc = -1;
    for (int i = 0; i < h->nsvc; i++) if (h->svcs[i].pid == pid) { svc = i; break; }
    if (svc >= 0) {
        int np = create_pid(e, actor, slot);
        add_proc(e, slot, np, p.kind, p.root);
        h->svcs[svc].pid = np;
    }
    if (si < 0) return 1;
    sess_t old = e->sess[ag][si];
    del_sess(e, ag, si);
    e->sess_count[ag]--;
    if (svc < 0) return 1;
    /* a session whose process is also a service process: re-created with a fresh pid (StopProcess.py:48-58) */
    add_session(e, actor, ag, slot, old.type, old.root, -1, old.type == 1 ? K_REDABS : K_SESSION, -1);
    return 1;
}

/* RestoreFromBackup.py:9-19 + Host.restore Host.py:373-429 */
static void do_restore(cc4o_env *e, int slot)
{
    host_t *h = &e->hosts[slot];
    sess_t saved[4]; int have[4];
    for (int i = 0; i < h->o_nsess; i++) {
        int ag = h->o_sess[i].agent, si = find_sess(e, ag, h->o_sess[i].id);
        have[i] = si >= 0 && e->sess[ag][si].host == slot;
        if (have[i]) saved[i] = e->sess[ag][si];
    }
    for (int k = 0; k < n_agents(e); k++) {
        int ag = agent_iter(e, k);
        for (int i = 0; i < e->nsess[ag];) if (e->sess[ag][i].host == slot) del_sess(e, ag, i); else i++;
    }
    h->ev_proc = h->ev_net = h->old_proc = h->old_net = 0; h->n_evpid = 0;
    h->nproc = h->o_nproc; memcpy(h->procs, h->o_procs, sizeof h->procs);
    h->nsvc = h->o_nsvc; memcpy(h->svcs, h->o_svcs, sizeof h->svcs);
    for (int i = 0; i < h->o_nsess; i++) {
        if (!have[i]) { fail(e, CC4O_E_UNREACHABLE, "restore: original session missing"); continue; }
        int ag = h->o_sess[i].agent;
        e->sess[ag][e->nsess[ag]++] = saved[i];
    }
}

/* DecoyAction.execute DecoyAction.py:47-112, DeployDecoy.py:8-31 */
static int do_decoy(cc4o_env *e, int b, int slot)
{
    int actor = A_BLUE0 + b;
    if (!blue_has_session(e, b, slot)) return SUCC_FALSE;
    host_t *h = &e->hosts[slot];
    int use80 = 0, use443 = 0, use25 = 0;
    for (int i = 0; i < h->nproc; i++) {
        int pb = KIND_PORTBIT[h->procs[i].kind];
        use80 |= pb == 2; use443 |= pb == 5; use25 |= pb == 4;
    }
    int fac[4], nf = 0;
    if (!use80) fac[nf++] = K_D_APACHE;
    if (!use443) fac[nf++] = K_D_TOMCAT;
    if (!use25) fac[nf++] = K_D_HARAKA;
    fac[nf++] = K_D_VSFTPD;   /* DecoyVsftpd.py:17-20 checks port 21, which nothing ever uses */
    int kind = fac[rng_int(e, actor, ST_DECOY_FACTORY, (uint32_t)nf)];
    int pid = create_pid(e, actor, slot);
    add_proc(e, slot, pid, kind, 0);
    int si = -1;
    for (int i = 0; i < h->nsvc; i++) if (h->svcs[i].code == kind) si = i;
    if (si < 0) { if (h->nsvc >= MAXSVC) { fail(e, CC4O_E_CAPACITY, "service table overflow"); return SUCC_TRUE; } si = h->nsvc++; }
    h->svcs[si].code = (uint8_t)kind; h->svcs[si].pid = pid; h->svcs[si].active = 1; h->svcs[si].reliab = 100;
    return SUCC_TRUE;
}

static int exec_blue(cc4o_env *e, int b, const act_t *a)
{
    int actor = A_BLUE0 + b;
    switch (a->code) {
    case AC_MONITOR: do_monitor(e, b); return SUCC_TRUE;
    case AC_ANALYSE: return blue_has_session(e, b, a->host) ? SUCC_TRUE : SUCC_FALSE;   /* Analyse.py:40-71: no state change */
    case AC_REMOVE: {   /* Remove.py:42-69 */
        if (!blue_has_session(e, b, a->host)) return SUCC_FALSE;
        int n = e->n_sus[b];
        for (int i = 0; i < n; i++) if (e->sus[b][i].slot == a->host) stop_process(e, actor, a->host, e->sus[b][i].pid, 0);
        return SUCC_TRUE;
    }
    case AC_RESTORE:    /* Restore.py:38-68 */
        if (find_sess(e, actor, 0) < 0 || !blue_has_session(e, b, a->host)) return SUCC_FALSE;
        do_restore(e, a->host);
        return SUCC_TRUE;
    case AC_DECOY: return do_decoy(e, b, a->host);
    case AC_BLOCK:      /* ControlTraffic.py:62-116 */
        if (find_sess(e, actor, 0) < 0) return SUCC_FALSE;
        if (e->blocks[a->subnet] >> a->host & 1) return SUCC_FALSE;
        e->blocks[a->subnet] |= 1u << a->host; return SUCC_TRUE;
    case AC_ALLOW:      /* ControlTraffic.py:132-185 */
        if (find_sess(e, actor, 0) < 0) return SUCC_FALSE;
        if (!(e->blocks[a->subnet] >> a->host & 1)) return SUCC_FALSE;
        e->blocks[a->subnet] &= ~(1u << a->host); return SUCC_TRUE;
    case AC_INVALID: return SUCC_FALSE;
    default: return SUCC_UNKNOWN;
    }
}

/* PhishingEmail._create_new_session PhishingEmail.py:42-113 */
static void do_phishing(cc4o_env *e, int g)
{
    int actor = A_GREEN0 + g, gh = e->green_host[g], gsn = slot_subnet(gh);
    for (int r = 0; r < NRED; r++) if (red_on_host(e, r, gh)) return;
    int src = -1, cand[NSLOT * NRED], nc = 0;
    for (int oi = 0; oi < e->norder; oi++) {
        int s = e->order[oi];
        for (int r = 0; r < NRED; r++) {
            if (!red_on_host(e, r, s)) continue;
            if (slot_subnet(s) == gsn) { src = r; break; }
            cand[nc++] = r;
        }
    }
    if (src < 0) {
        if (nc == 0) return;
        src = cand[rng_int(e, actor, ST_PHISH_AGENT, (uint32_t)nc)];
    }
    add_session(e, actor, A_RED0 + src, gh, 1, 0, -1, K_REDABS, -1);
}

static int exec_green(cc4o_env *e, int g, const act_t *a)
{
    int actor = A_GREEN0 + g, gh = e->green_host[g];
    if (a->code == AC_G_LOCAL) {   /* GreenLocalWork.execute GreenLocalWork.py:60-125 */
        host_t *h = &e->hosts[gh];
        int act[MAXSVC], na = 0;
        for (int i = 0; i < h->nsvc; i++) if (h->svcs[i].active) act[na++] = i;
        if (na == 0) return SUCC_FALSE;
        int si = act[rng_int_more(e, actor, ST_LW_SERVICE, (uint32_t)na)];     /* follows the actor's ST_GREEN_ACT draw */
        if ((int)rng_int(e, actor, ST_LW_RELIAB, 100) >= h->svcs[si].reliab) return SUCC_FALSE;
        if (rng_f64(e, actor, ST_LW_FP, 0) < 0.01) ev_procpid(e, gh, -1);
        if (rng_f64(e, actor, ST_LW_PHISH, 0) < 0.01) do_phishing(e, g);
        return SUCC_TRUE;
    }
    if (a->code == AC_G_ACCESS) {  /* GreenAccessService.execute GreenAccessService.py:137-217, random_reachable_ip :94-115 */
        int gsn = slot_subnet(gh);
        uint16_t allowed = GREEN_ALLOWED[e->phase][gsn];
        int dst[NSLOT], nd = 0;
        for (int oi = 0; oi < e->norder; oi++) { int s = e->order[oi]; if (slot_is_server(s) && (allowed >> slot_subnet(s) & 1)) dst[nd++] = s; }
        int d = dst[rng_int(e, actor, ST_AS_DEST, (uint32_t)nd)], dsn = slot_subnet(d);
        if ((e->blocks[gsn] >> dsn & 1) || (e->blocks[dsn] >> gsn & 1)) { e->hosts[d].ev_net = 1; return SUCC_FALSE; }
        if (rng_f64(e, actor, ST_AS_FP, 0) < 0.01) e->hosts[d].ev_net = 1;
        return SUCC_TRUE;
    }
    return SUCC_UNKNOWN;
}

/* ExploitAction._create_new_session ExploitAction.py:212-262 */
static int exploit_new_session(cc4o_env *e, int r, int target, int parent_session)
{
    (void)parent_session;
    int actor = A_RED0 + r;
    int pid = create_pid(e, actor, target);
    add_proc(e, target, pid, K_SHELL, 0);
    int id = add_session(e, actor, actor, target, 0, 0, pid, K_SHELL, -1);
    int at = find_sess(e, actor, id);
    if (at >= 0) e->sess[actor][at].has_parent = 1;     /* parent = the acting session, ExploitAction.py:256 */
    return pid;
}

static int exec_red(cc4o_env *e, int r, const act_t *a)
{
    int actor = A_RED0 + r;
    int si = action_has_session(a->code) ? find_sess(e, actor, a->session) : -1;
    sess_t *S = si >= 0 ? &e->sess[actor][si] : NULL;
    switch (a->code) {
    case AC_DRS: {     /* DiscoverRemoteSystems.py:39 -> Pingsweep.execute Pingsweep.py:31-64 */
        int sn = a->subnet, succ = SUCC_UNKNOWN;
        if (!S) return SUCC_FALSE;                     /* session vanished after filter_actions: Pingsweep.py:36-39 */
        if (sn == S_INT) return succ;                  /* root_internet_host_0 does not respond to ping */
        for (int k = 0; k < e->nuser[sn] + e->nserver[sn]; k++) {
            int slot = k < e->nuser[sn] ? sn * 16 + 6 + k : sn * 16 + (k - e->nuser[sn]);
            succ = SUCC_TRUE;
            /* SimulationController._filter_obs :1054: interfaces of subnets outside allowed_subnets are dropped */
            if (red_allowed(r, sn)) obs_entry(e, r, 1, slot, 1, 1, 0, 0);
        }
        return succ;
    }
    case AC_AGGR: case AC_STEALTH: {   /* DiscoverNetworkServices.py:44-86 -> Portscan.py:23-65 */
        if (!S || S->type != 1) return SUCC_FALSE;
        if (blocking(e, S->host, a->host)) return SUCC_FALSE;
        host_t *h = &e->hosts[a->host];
        double rate = a->code == AC_AGGR ? 0.75 : 0.25;
        double u = rng_f64(e, actor, ST_PORTSCAN, 0);
        int mask = 0, any = 0;
        for (int i = 0; i < h->nproc; i++) {
            int pb = KIND_PORTBIT[h->procs[i].kind];
            if (pb < 0) continue;
            any = 1; mask |= 1 << pb;
            if (u <= rate || h->procs[i].kind >= K_D_APACHE) h->ev_net = 1;
        }
        if (any) { S->has_ports[a->host] = 1; S->ports[a->host] = (uint8_t)mask; obs_entry(e, r, 1, a->host, 1, 0, 0, 0); }
        return SUCC_TRUE;
    }
    case AC_DD: {      /* DiscoverDeception.py:44-101 */
        if (!S) return SUCC_FALSE;
        host_t *h = &e->hosts[a->host];
        int rep = 0;
        for (int i = 0; i < h->nproc; i++) {
            int decoy = h->procs[i].kind >= K_D_APACHE && h->procs[i].kind <= K_D_VSFTPD;
            if (e->rng_mode != 1) {     /* production mode: both words are spent for every process (the engine takes them lane-parallel) */
                double u1 = rng_f64(e, actor, ST_DD_TP, 0), u2 = rng_f64(e, actor, ST_DD_FP, 0);
                if (decoy ? u1 <= 0.5 : u2 <= 0.1) rep = 1;
                continue;
            }
            if (rng_f64(e, actor, ST_DD_TP, 0) <= 0.5 && decoy) rep = 1;
            else if (rng_f64(e, actor, ST_DD_FP, 0) <= 0.1 && !decoy) rep = 1;
        }
        if (rep) obs_entry(e, r, 0, a->host, 1, 0, 0, 0);
        return SUCC_TRUE;
    }
    case AC_EXPLOIT: { /* ExploitRemoteService.py:149-202 */
        if (!S || S->type != 1) return SUCC_FALSE;
        if (!S->has_ports[a->host]) return SUCC_FALSE;
        if (blocking(e, S->host, a->host)) return SUCC_FALSE;
        int ports = S->ports[a->host];
        /* DefaultExploitActionSelector :37-69: options in list order HTTPRFI, HTTPSRFI, SSH, SQLi, Haraka, FTP */
        enum { X_HTTP, X_HTTPS, X_SSH, X_SQL, X_HARAKA };
        int opt[5], no = 0; double w[5];
        if (ports & PB_80) { opt[no] = X_HTTP; w[no++] = 3.0; }
        if (ports & PB_443) { opt[no] = X_HTTPS; w[no++] = 4.0; }
        if (ports & PB_22) { opt[no] = X_SSH; w[no++] = 0.1; }
        if ((ports & PB_3390) && (ports & (PB_80 | PB_443))) { opt[no] = X_SQL; w[no++] = 5.0; }
        if (ports & PB_25) { opt[no] = X_HARAKA; w[no++] = 6.0; }
        if (no < 1) return SUCC_FALSE;
        int sel = opt[0];
        if (no > 1) {
            int top = 0;
            for (int i = 1; i < no; i++) if (w[i] > w[top]) top = i;
            for (int i = top; i + 1 < no; i++) opt[i] = opt[i + 1];
            no--;
            sel = opt[rng_int(e, actor, ST_EXP_OTHER, (uint32_t)no)];
            (void)rng_f64(e, actor, ST_EXP_TOP, 0);   /* random() < 0 is never true */
        }
        int target = a->host, src = S->host;
        host_t *h = &e->hosts[target];
        if (sel == X_SSH) {   /* SSHBruteForce.execute SSHBruteForce.py:24-84 */
            int hops[16], nh = route(src, target, hops);
            for (int i = 0; i < nh; i++) if (0.05 < rng_f64(e, actor, ST_SSH_HOP, 0)) e->hosts[hops[i]].ev_net = 1;
            int vp = -1;
            for (int i = 0; i < h->nproc; i++) if (h->procs[i].kind == K_SSHD) { vp = i; break; }
            if (vp < 0) return SUCC_FALSE;
            obs_entry(e, r, 1, target, 1, 0, 0, 0);
            h->ev_net = 1;   /* _create_brute_force_event :98-105 */
            int pid = exploit_new_session(e, r, target, a->session);
            ev_procpid(e, target, pid);
            obs_entry(e, r, 1, target, 1, 0, 1, 1);
            obs_entry(e, r, 1, src, 1, 0, 0, 0);
            return SUCC_TRUE;
        }
        /* ExploitAction.sim_exploit ExploitAction.py:48-116, _get_vulnerable_process :178-200 */
        int vp = -1;
        for (int i = 0; i < h->nproc && vp < 0; i++) {
            int k = h->procs[i].kind;
            if (sel == X_HTTP && (k == K_APACHE2 || k == K_D_APACHE || k == K_D_VSFTPD)) vp = i;
            if (sel == X_HTTPS && k == K_D_TOMCAT) vp = i;
            if (sel == X_SQL && k == K_MYSQLD) vp = i;
            if (sel == X_HARAKA && (k == K_SMTP || k == K_D_HARAKA)) vp = i;
        }
        if (vp < 0) return SUCC_FALSE;
        int vk = h->procs[vp].kind, decoy = vk >= K_D_APACHE && vk <= K_D_VSFTPD;
        if (decoy) h->ev_net = 1;
        obs_entry(e, r, 1, target, 1, 0, 0, 0);
        int works = sel == X_SQL ? 1 : decoy;   /* HTTPRFI.py:18, HTTPSRFI.py:18: 'rfi' only on decoys; HarakaRCE.py:19: real SMTP is 2.8.9 */
        if (!works) return SUCC_FALSE;
        if (0.0 >= rng_f64(e, actor, ST_EXP_SUCC, 0)) return SUCC_FALSE;
        if (decoy) { h->ev_net = 1; return SUCC_FALSE; }
        int pid = exploit_new_session(e, r, target, a->session);
        if (rng_f64(e, actor, ST_EXP_DETECT, 0) > 0.05) ev_procpid(e, target, pid);
        obs_entry(e, r, 1, target, 1, 0, 1, 1);
        obs_entry(e, r, 1, src, 1, 0, 0, 0);
        return SUCC_TRUE;
    }
    case AC_PRIVESC: { /* PrivilegeEscalate.py:127-179 */
        int on[MAXSESS], n = 0;
        for (int i = 0; i < e->nsess[actor]; i++) if (e->sess[actor][i].host == a->host) on[n++] = i;
        if (n == 0) return SUCC_FALSE;
        int tgt = -1;
        for (int i = 0; i < n; i++) if (e->sess[actor][on[i]].root) { tgt = on[i]; break; }
        if (tgt < 0) {
            tgt = on[rng_int(e, actor, ST_PRIV_SESS, (uint32_t)n)];
            if (!S || S->type != 1) return SUCC_FALSE; /* DefaultEscalateActionSelector :37-48 */
            sess_t *T = &e->sess[actor][tgt];
            T->root = 1;                                /* EscalateAction.__upgrade_session EscalateAction.py:54-87 */
            int pi = find_proc(&e->hosts[a->host], T->pid);
            if (pi >= 0) e->hosts[a->host].procs[pi].root = 1; else fail(e, CC4O_E_UNREACHABLE, "escalate: session process missing");
        }
        sess_t *T = &e->sess[actor][tgt];
        obs_entry(e, r, 0, a->host, 0, 0, 0, 1);
        if (T->type == 1) know_session(e, r, T->id);
        /* ExploreHost EscalateAction.py:90-106 with Host.info = _between_subnet_links :418-468 */
        if (S && slot_is_server(a->host) && (a->host & 15) == 0) {   /* ExploreHost needs self.session too :99-104 */
            uint16_t lk = LINKS[slot_subnet(a->host)];
            for (int oi = 0; oi < e->norder; oi++) { int s = e->order[oi]; if (s < 128 && (s & 15) == 0 && (lk >> slot_subnet(s) & 1)) obs_entry(e, r, 0, s, 1, 0, 0, 0); }
        }
        return SUCC_TRUE;
    }
    case AC_IMPACT: {  /* Impact.py:38-88 */
        int any = 0, rootsess = 0;
        for (int i = 0; i < e->nsess[actor]; i++) if (e->sess[actor][i].host == a->host) { any = 1; if (e->sess[actor][i].root) rootsess = 1; }
        if (!any || !rootsess) return SUCC_FALSE;
        host_t *h = &e->hosts[a->host];
        int sv = -1;
        for (int i = 0; i < h->nsvc; i++) if (h->svcs[i].code == K_OT && h->svcs[i].active) sv = i;
        if (sv < 0) return SUCC_FALSE;
        if (!S) return SUCC_FALSE;                      /* TargetedLocalAction.py:14-21 needs the acting session too */
        h->svcs[sv].active = 0;                         /* Host.stop_service Host.py:296-301 */
        int pi = find_proc(h, h->svcs[sv].pid);          /* State.remove_process State.py:390-418 */
        if (pi >= 0) {
            int pid = h->procs[pi].pid, ag, s2 = session_from_pid(e, a->host, pid, &ag);
            del_proc(h, pi);
            if (s2 >= 0) del_sess(e, ag, s2);
        }
        obs_entry(e, r, 0, a->host, 0, 0, 0, 0);
        return SUCC_TRUE;
    }
    case AC_DEGRADE: { /* DegradeServices.py:38-82 */
        int any = 0, rootsess = 0;
        for (int i = 0; i < e->nsess[actor]; i++) if (e->sess[actor][i].host == a->host) { any = 1; if (e->sess[actor][i].root) rootsess = 1; }
        if (!any || !rootsess) return SUCC_FALSE;
        host_t *h = &e->hosts[a->host];
        int n = 0;
        for (int i = 0; i < h->nsvc; i++) if (h->svcs[i].active) { n++; h->svcs[i].reliab = h->svcs[i].reliab >= 20 ? h->svcs[i].reliab - 20 : 0; }
        if (n == 0) return SUCC_FALSE;
        obs_entry(e, r, 0, a->host, 0, 0, 0, 0);
        return SUCC_TRUE;
    }
    case AC_WITHDRAW: {   /* Withdraw.py:38-91 (RandomSelectRedAgent only): every session of the agent on the host goes, children first */
        if (!S) return SUCC_FALSE;
        int ids[MAXSESS + 1], n = 0;
        for (int i = 0; i < e->nsess[actor]; i++) if (e->sess[actor][i].host == a->host && e->sess[actor][i].has_parent) ids[n++] = e->sess[actor][i].id;
        for (int i = 0; i < e->nsess[actor]; i++) if (e->sess[actor][i].host == a->host && !e->sess[actor][i].has_parent && e->sess[actor][i].id != 0) ids[n++] = e->sess[actor][i].id;
        if (S->host == a->host) ids[n++] = a->session;     /* the acting session last (a second time when its ident is not 0) */
        if (n == 0) return SUCC_FALSE;
        for (int k = 0; k < n; k++) {
            /* StopProcess(stop_all=True): TargetedLocalAction.py:14-21 needs the acting and the target session, StopProcess.py:23-33 the process */
            int ti = find_sess(e, actor, ids[k]);
            if (find_sess(e, actor, a->session) < 0 || ti < 0) return SUCC_FALSE;
            if (!stop_process(e, actor, a->host, e->sess[actor][ti].pid, 1)) return SUCC_FALSE;
        }
        return SUCC_TRUE;
    }
    case AC_INVALID: return SUCC_FALSE;
    default: return SUCC_UNKNOWN;
    }
}

/* RandomSelectRedAgent.get_action RandomSelectRedAgent.py:37-92: a uniformly drawn valid command, then one uniform draw per
   parameter that has more than one known option (options in action-space dict order: subnets and hosts in creation order,
   sessions in the order they became known) */
static int nth_known_host(cc4o_env *e, const uint8_t *known, int k)
{
    for (int oi = 0; oi < e->norder; oi++) { int s = e->order[oi]; if (known[s] && k-- == 0) return s; }
    return -1;
}
static act_t red_policy_random(cc4o_env *e, int r)
{
    int actor = A_RED0 + r;
    act_t a = { AC_SLEEP, -1, -1, 0 };
    int nss = e->n_ss[r], nsub = 0, nip = 0, nhn = 0;
    for (int s = 0; s < NSUB; s++) nsub += e->sn_known[r] >> s & 1;
    for (int s = 0; s < NSLOT; s++) { nip += e->ip_known[r][s]; nhn += e->hn_known[r][s]; }
    /* action_space['action'] order = red_actions, EnterpriseScenarioGenerator.py:764-768; a command is valid when every parameter has an option */
    static const int cmds[10] = { AC_DRS, AC_AGGR, AC_STEALTH, AC_EXPLOIT, AC_PRIVESC, AC_DEGRADE, AC_DD, AC_IMPACT, AC_WITHDRAW, AC_SLEEP };
    int list[10], n = 0;
    for (int i = 0; i < 10; i++) {
        int c = cmds[i], ok;
        if (c == AC_SLEEP) ok = 1;
        else if (c == AC_DRS) ok = nss && nsub;
        else if (c == AC_PRIVESC || c == AC_DEGRADE || c == AC_IMPACT) ok = nss && nhn;
        else if (c == AC_WITHDRAW) ok = nss && nip && nhn;
        else ok = nss && nip;
        if (ok) list[n++] = c;
    }
    a.code = list[rng_int(e, actor, ST_RS_COMMAND, (uint32_t)n)];
    #define RS_PICK(cnt) ((cnt) > 1 ? (int)rng_int(e, actor, ST_RS_PARAM, (uint32_t)(cnt)) : 0)
    #define RS_SESSION() (a.session = e->ss[r][RS_PICK(nss)])
    #define RS_IP() nth_known_host(e, e->ip_known[r], RS_PICK(nip))
    #define RS_HOSTNAME() nth_known_host(e, e->hn_known[r], RS_PICK(nhn))
    switch (a.code) {   /* parameters in constructor-signature order */
    case AC_DRS: { int k = RS_PICK(nsub); for (int s = 0; s < NSUB; s++) if ((e->sn_known[r] >> s & 1) && k-- == 0) { a.subnet = s; break; } RS_SESSION(); break; }
    case AC_AGGR: case AC_STEALTH: case AC_DD: RS_SESSION(); a.host = RS_IP(); break;
    case AC_EXPLOIT: a.host = RS_IP(); RS_SESSION(); break;
    case AC_PRIVESC: case AC_DEGRADE: case AC_IMPACT: a.host = RS_HOSTNAME(); RS_SESSION(); break;
    case AC_WITHDRAW: RS_SESSION(); (void)RS_IP(); a.host = RS_HOSTNAME(); break;      /* ip_address only routes the RemoteAction */
    default: break;
    }
    #undef RS_PICK
    #undef RS_SESSION
    #undef RS_IP
    #undef RS_HOSTNAME
    return a;
}

/* RedSessionCheck.execute RedSessionCheck.py:9-63 */
static void red_session_check(cc4o_env *e, int r)
{
    int actor = A_RED0 + r;
    act_t rsc = { 100, -1, -1, 0 };
    obs_push(e, r, SUCC_TRUE, &rsc);
    if (e->sess_count[actor] == 0) return;
    if (e->nsess[actor] == 0) { fail(e, CC4O_E_UNREACHABLE, "RedSessionCheck with stale count and no session"); return; }
    if (find_sess(e, actor, 0) < 0) {
        int i = (int)rng_int(e, actor, ST_RSC_PRIMARY, (uint32_t)e->nsess[actor]);
        sess_t s = e->sess[actor][i];
        del_sess(e, actor, i);
        s.id = 0; s.has_parent = 0;
        e->sess[actor][e->nsess[actor]++] = s;
        for (int k = 0; k < e->nsess[actor]; k++) e->sess[actor][k].has_parent = 0;     /* parent = new primary's name = None, :50-54 */
    }
    for (int i = 0; i < e->nsess[actor]; i++) {
        sess_t *s = &e->sess[actor][i];
        obs_entry(e, r, 0, s->host, 1, 1, 1, 1);
        if (s->type == 1) know_session(e, r, s->id);
    }
}

/* SimulationController.different_subnet_agent_reassignment :820-903 */
static void reassign(cc4o_env *e)
{
    struct { int orig, id, host; } lst[64]; int n = 0;
    for (int r = 0; r < NRED; r++) {
        int ag = A_RED0 + r;
        for (int i = 0; i < e->nsess[ag]; i++) if (!red_allowed(r, slot_subnet(e->sess[ag][i].host))) {
            if (n >= 64) { fail(e, CC4O_E_CAPACITY, "reassign overflow"); break; }
            lst[n].orig = r; lst[n].id = e->sess[ag][i].id; lst[n].host = e->sess[ag][i].host; n++;
        }
    }
    for (int k = 0; k < n; k++) {
        int r = lst[k].orig, ag = A_RED0 + r, nr = red_of_subnet(slot_subnet(lst[k].host));
        if (nr < 0) { fail(e, CC4O_E_UNREACHABLE, "session on a subnet no red agent owns"); continue; }
        int si = find_sess(e, ag, lst[k].id);
        sess_t old = e->sess[ag][si];
        del_sess(e, ag, si);
        int nid = add_session(e, A_RED0 + nr, A_RED0 + nr, old.host, 1, old.root, old.pid, K_REDABS, -1);
        e->sess_count[ag]--;
        /* the creating agent's observation of this step (if any) is edited and copied to the new owner :868-893 */
        obsset_t *o = &e->obs[r];
        for (int i = 0; i < o->n; i++) if (o->e[i].key_ip && o->e[i].slot == old.host) {
            if (!o->e[i].sess) { fail(e, CC4O_E_UNREACHABLE, "reassign: obs entry without Sessions"); break; }
            obs_push(e, nr, SUCC_UNKNOWN, NULL);
            obs_entry(e, nr, 1, old.host, o->e[i].iface, o->e[i].subnet_if, o->e[i].syshost, 1);
            know_session(e, nr, nid);
            break;
        }
    }
    for (int a = 0; a < NAG; a++) {
        if (a >= A_GREEN0 + e->ngreen) break;
        e->active[a] = e->nsess[a] > 0;   /* :896-903 (every session is active and either parentless or a child) */
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* observation vector: BlueFlatWrapper.observation_change BlueFlatWrapper.py:172-256                */
/* ------------------------------------------------------------------------------------------------ */
int cc4o_observe(cc4o_env *e, uint8_t *obs)
{
    memset(obs, 0, NBLUE * CC4O_OBS_LEN);
    for (int b = 0; b < NBLUE; b++) {
        uint8_t *v = obs + b * CC4O_OBS_LEN;
        int p = 0;
        v[p++] = (uint8_t)e->phase;
        for (int i = 0; i < blue_nsub(b); i++) {
            int sn = blue_sorted_subnet(b, i);
            for (int j = 0; j < NSUB; j++) v[p++] = SORTED_SUBNETS[j] == sn;
            for (int j = 0; j < NSUB; j++) v[p++] = e->blocks[sn] >> SORTED_SUBNETS[j] & 1;
            for (int j = 0; j < NSUB; j++) v[p++] = !(COMMS_ADJ[e->phase][sn] >> SORTED_SUBNETS[j] & 1);
            for (int k = 0; k < 16; k++) { host_t *h = &e->hosts[sn * 16 + k]; v[p++] = h->exists && (h->old_proc || h->ev_proc); }
            for (int k = 0; k < 16; k++) { host_t *h = &e->hosts[sn * 16 + k]; v[p++] = h->exists && (h->old_net || h->ev_net); }
        }
        /* messages from the other four blue agents in sorted sender order, SimulationController.send_messages :781-804 */
        for (int o = 0; o < NBLUE; o++) { if (o == b) continue; for (int k = 0; k < 8; k++) v[p++] = e->msgs[o] >> k & 1; }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* step: SimulationController.step :211-315                                                         */
/* ------------------------------------------------------------------------------------------------ */
int cc4o_step(cc4o_env *e, const int32_t *actions, const uint8_t *msgs, uint8_t *obs, float *reward, uint8_t *done)
{
    int nact = A_GREEN0 + e->ngreen;
    memset(e->seg_pos, 0, sizeof e->seg_pos); memset(e->draw_ctr, 0, sizeof e->draw_ctr); memset(e->blk_id, 0xFF, sizeof e->blk_id);
    /* mission phase, State.check_next_phase_on_update_step State.py:514-544, _generate_mission_phases :854-860 */
    {
        int q = e->steps / 3, rem = e->steps % 3, l0 = q + (rem >= 1), l1 = q + (rem == 2);
        int np = e->step_count < l0 ? 0 : e->step_count < l0 + l1 ? 1 : e->step_count < e->steps ? 2 : -1;
        if (np < 0) { fail(e, CC4O_E_UNREACHABLE, "step beyond last mission phase"); return -1; }
        if (np > e->phase) e->phase = np;
    }
    /* 1. collect this tick's submitted actions in agent order (blue, green, red) :236-248 */
    act_t sub[NAG]; double cost = 0;
    for (int a = 0; a < NAG; a++) sub[a].code = AC_SLEEP;
    for (int b = 0; b < NBLUE; b++) {
        sub[b] = blue_decode(e, b, actions[b], NULL);
        if (sub[b].code == AC_RESTORE) cost += -1;      /* Restore.py:70, charged on submission :310-311 */
    }
    for (int g = 0; g < e->ngreen; g++) {
        int a = A_GREEN0 + g;
        act_t x = { AC_SLEEP, -1, -1, 0 };
        if (e->active[a]) { uint32_t c = rng_int(e, a, ST_GREEN_ACT, 3); x.code = c == 0 ? AC_G_ACCESS : c == 1 ? AC_G_LOCAL : AC_SLEEP; }
        sub[a] = x;
    }
    for (int r = 0; r < NRED; r++) {
        act_t x = { AC_SLEEP, -1, -1, 0 };
        if (e->active[A_RED0 + r]) x = e->red_variant == 2 ? red_policy_random(e, r) : red_policy(e, r);
        sub[A_RED0 + r] = x;
    }
    /* clear old observations :251 */
    for (int r = 0; r < NRED; r++) { e->obs[r].n = 0; e->obs[r].have_first = 0; e->obs[r].has_act0 = 0; e->obs[r].succ0 = SUCC_UNKNOWN; }
    /* 2. duration queue :245-265 */
    act_t ex[NAG]; int ready[NAG];
    for (int k = 0; k < n_agents(e); k++) {
        int a = agent_iter(e, k);
        if (e->inprog[a].code == AC_NONE) { e->inprog[a] = sub[a]; e->inprog_ticks[a] = action_duration(sub[a].code); }
        e->inprog_ticks[a]--;
        if (e->inprog_ticks[a] < 1) { ex[a] = e->inprog[a]; e->inprog[a].code = AC_NONE; e->inprog_ticks[a] = 0; ready[a] = 1; }
        else {
            ready[a] = 0; ex[a].code = AC_SLEEP; ex[a].host = ex[a].subnet = -1; ex[a].session = 0;
            if (a >= A_RED0 && a < A_GREEN0) obs_push(e, a - A_RED0, SUCC_INPROG, NULL);
        }
        e->cur_action[a] = ex[a];
    }
    /* 3+4. execute: priority-1 actions (Block/AllowTrafficZone) first, then agent order; actions whose session
       no longer exists are dropped :398-404, :466-486 */
    uint8_t dropped[NAG];   /* filter_actions runs on the whole list BEFORE anything executes */
    for (int k = 0; k < n_agents(e); k++) {
        int a = agent_iter(e, k);
        int sid = a < A_RED0 ? 0 : a < A_GREEN0 ? ex[a].session : 0;
        dropped[a] = action_has_session(ex[a].code) && find_sess(e, a, sid) < 0;
    }
    for (int pass = 0; pass < 2; pass++) {
        for (int k = 0; k < n_agents(e); k++) {
            int a = agent_iter(e, k);
            int prio1 = ex[a].code == AC_BLOCK || ex[a].code == AC_ALLOW;
            if ((pass == 0) != prio1) continue;
            if (dropped[a]) continue;
            if (a < A_RED0) exec_blue(e, a, &ex[a]);
            else if (a < A_GREEN0) { int s = exec_red(e, a - A_RED0, &ex[a]); obs_push(e, a - A_RED0, s, &ex[a]); }
            else e->green_succ[a - A_GREEN0] = exec_green(e, a - A_GREEN0, &ex[a]);
        }
    }
    (void)ready; (void)nact;
    /* 5. reassignment :278 */
    reassign(e);
    /* 6. end-turn actions for active agents :281-286 */
    for (int b = 0; b < NBLUE; b++) if (e->active[b] && find_sess(e, b, 0) >= 0) do_monitor(e, b);
    for (int r = 0; r < NRED; r++) if (e->active[A_RED0 + r]) red_session_check(e, r);
    /* 7. :297-300 */
    e->step_count++;
    e->done = e->step_count >= e->steps - 1;
    /* 8. BlueRewardMachine.calculate_reward BlueRewardMachine.py:70-121 */
    int rew = 0;
    for (int g = 0; g < e->ngreen; g++) {
        int a = A_GREEN0 + g, c = e->cur_action[a].code;
        if ((c == AC_G_LOCAL || c == AC_G_ACCESS) && e->nsess[a] > 0 && e->green_succ[g] == SUCC_FALSE)
        { rew += REWARD[e->phase][slot_subnet(e->green_host[g])][c == AC_G_LOCAL ? 0 : 1]; e->breakdown[e->phase][c == AC_G_LOCAL ? 0 : 1][slot_subnet(e->green_host[g])]++; }
    }
    for (int r = 0; r < NRED; r++) {
        int a = A_RED0 + r;
        /* `success and isinstance(action, Impact)` :118 — a TernaryEnum member is always truthy (Enums.py:5-24) */
        if (e->cur_action[a].code == AC_IMPACT && e->nsess[a] > 0) { rew += REWARD[e->phase][slot_subnet(e->cur_action[a].host)][2]; e->breakdown[e->phase][2][slot_subnet(e->cur_action[a].host)]++; }
    }
    if (msgs) memcpy(e->msgs, msgs, NBLUE); else memset(e->msgs, 0, NBLUE);
    if (reward) *reward = (float)((double)rew + cost);
    if (done) *done = (uint8_t)e->done;
    if (obs) cc4o_observe(e, obs);
    if (e->rng_mode == 1) {
        for (int a = 0; a < NAG; a++) {
            uint32_t seg = 1 + (uint32_t)(e->step_count - 1) * NAG + (uint32_t)a;
            if (seg < e->n_segs && e->seg_pos[a] != e->seg_off[seg + 1] - e->seg_off[seg]) {
                char b[128]; snprintf(b, sizeof b, "actor %d consumed %u of %u draws", a, e->seg_pos[a], e->seg_off[seg + 1] - e->seg_off[seg]);
                e->step_count--; fail(e, CC4O_E_TAPE_LEFT, b); e->step_count++;
            }
        }
    }
    return e->err ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* canonical export (oracle/canon.py)                                                               */
/* ------------------------------------------------------------------------------------------------ */
int cc4o_export(cc4o_env *e, int32_t *out, int cap)
{
    int n = 0;
    #define PUT(v) do { if (n < cap) out[n] = (int32_t)(v); n++; } while (0)
    PUT(e->phase); PUT(e->step_count);
    for (int s = 0; s < NSUB; s++) PUT(e->blocks[s]);
    int nh = 0; for (int s = 0; s < NSLOT; s++) nh += e->hosts[s].exists;
    PUT(nh);
    for (int s = 0; s < NSLOT; s++) {
        host_t *h = &e->hosts[s];
        if (!h->exists) continue;
        PUT(s); PUT(h->ev_proc | h->ev_net << 1 | h->old_proc << 2 | h->old_net << 3); PUT(h->nproc);
        for (int i = 0; i < h->nproc; i++) { PUT(h->procs[i].pid); PUT(h->procs[i].kind); PUT(h->procs[i].root); }
        PUT(h->nsvc);
        for (int i = 0; i < h->nsvc; i++) { PUT(h->svcs[i].code); PUT(h->svcs[i].pid); PUT(h->svcs[i].active); PUT(h->svcs[i].reliab); }
    }
    for (int r = 0; r < NRED; r++) {
        int a = A_RED0 + r;
        PUT(e->active[a]); PUT(e->sess_count[a]); PUT(e->nsess[a]);
        for (int i = 0; i < e->nsess[a]; i++) { sess_t *s = &e->sess[a][i]; PUT(s->id); PUT(s->host); PUT(s->type == 1); PUT(s->root); PUT(s->pid); }
        int c = 0; for (int s = 0; s < NSLOT; s++) c += e->ip_known[r][s];
        PUT(c); for (int s = 0; s < NSLOT; s++) if (e->ip_known[r][s]) PUT(s);
        c = 0; for (int s = 0; s < NSLOT; s++) c += e->hn_known[r][s];
        PUT(c); for (int s = 0; s < NSLOT; s++) if (e->hn_known[r][s]) PUT(s);
        PUT(e->sn_known[r]);
        PUT(e->n_ss[r]); for (int i = 0; i < e->n_ss[r]; i++) PUT(e->ss[r][i]);
        {   /* FSM view AFTER the agent has digested this tick's observation (what get_action will do first thing next
               tick, FiniteStateRedAgent.py:91-93); inactive agents are not asked (AgentInterface.py:133-136) */
            static fsm_t sv[MAXFSM]; int sn_ = e->n_fsm[r], serr = e->err; char smsg[256];
            memcpy(sv, e->fsm[r], sizeof sv); memcpy(smsg, e->errmsg, sizeof smsg);
            if (e->active[a] && e->red_variant != 2) { fsm_transition(e, r); fsm_process_obs(e, r); fsm_session_removal(e, r); }
            PUT(e->n_fsm[r]); for (int i = 0; i < e->n_fsm[r]; i++) { PUT(e->fsm[r][i].slot); PUT(e->fsm[r][i].state); PUT(e->fsm[r][i].hn); }
            memcpy(e->fsm[r], sv, sizeof sv); e->n_fsm[r] = sn_; e->err = serr; memcpy(e->errmsg, smsg, sizeof smsg);
        }
        /* ports sorted by (ident, slot) */
        int np = 0;
        for (int i = 0; i < e->nsess[a]; i++) if (e->sess[a][i].type == 1) for (int s = 0; s < NSLOT; s++) np += e->sess[a][i].has_ports[s];
        PUT(np);
        int idx[MAXSESS], ni = 0;
        for (int i = 0; i < e->nsess[a]; i++) idx[ni++] = i;
        for (int i = 0; i < ni; i++) for (int j = i + 1; j < ni; j++) if (e->sess[a][idx[j]].id < e->sess[a][idx[i]].id) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
        for (int i = 0; i < ni; i++) { sess_t *s = &e->sess[a][idx[i]]; if (s->type != 1) continue; for (int h = 0; h < NSLOT; h++) if (s->has_ports[h]) { PUT(s->id); PUT(h); PUT(s->ports[h]); } }
        PUT(e->inprog[a].code); PUT(e->inprog[a].code == AC_NONE ? 0 : e->inprog_ticks[a]);
    }
    for (int b = 0; b < NBLUE; b++) {
        PUT(e->n_sus[b]);
        for (int s = 0; s < NSLOT; s++) for (int i = 0; i < e->n_sus[b]; i++) if (e->sus[b][i].slot == s) { PUT(s); PUT(e->sus[b][i].pid); }
        PUT(e->inprog[b].code); PUT(e->inprog[b].code == AC_NONE ? 0 : e->inprog_ticks[b]);
    }
    #undef PUT
    return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* life cycle                                                                                       */
/* ------------------------------------------------------------------------------------------------ */
void cc4o_set_red_agent(cc4o_env *e, int variant) { e->red_variant = variant == 2 ? 2 : variant ? 1 : 0; }

cc4o_env *cc4o_create(uint64_t seed, uint64_t episode_id, int steps, int rng_mode)
{
    cc4o_env *e = (cc4o_env *)calloc(1, sizeof *e);
    if (!e) return NULL;
    e->seed = seed; e->episode = episode_id; e->steps = steps; e->rng_mode = rng_mode;
    for (int a = 0; a < NAG; a++) {
        int cap = (a >= A_GREEN0) ? 8 : MAXSESS;
        e->sess[a] = (sess_t *)calloc((size_t)cap, sizeof(sess_t));
    }
    init_tables();
    return e;
}
void cc4o_destroy(cc4o_env *e)
{
    if (!e) return;
    for (int a = 0; a < NAG; a++) free(e->sess[a]);
    free(e);
}
void cc4o_set_tape(cc4o_env *e, const cc4o_draw *recs, const uint32_t *seg_off, uint32_t n_segs)
{
    e->tape = recs; e->seg_off = seg_off; e->n_segs = n_segs;
}
int cc4o_error_flags(cc4o_env *e) { return e->err; }
const char *cc4o_last_error(cc4o_env *e) { return e->errmsg; }

long cc4o_run_episodes(uint64_t seed, uint64_t first_episode, int n_eps, int steps, double *ret_sum)
{
    long ticks = 0; double total = 0;
    cc4o_env *e = cc4o_create(seed, first_episode, steps, 0);
    for (int k = 0; k < n_eps; k++) {
        e->episode = first_episode + (uint64_t)k;
        cc4o_reset(e);
        for (int t = 0; t < steps - 1; t++) {
            int32_t act[NBLUE]; uint8_t msg[NBLUE] = { 0 }; float r; uint8_t d;
            uint32_t c[4] = { (uint32_t)t, 0xB10Eu, (uint32_t)e->episode, (uint32_t)(e->episode >> 32) };
            philox4x32_10(c, (uint32_t)seed ^ 0x5EEDu, (uint32_t)(seed >> 32));
            uint32_t c2[4] = { (uint32_t)t, 0xB10Fu, (uint32_t)e->episode, (uint32_t)(e->episode >> 32) };
            philox4x32_10(c2, (uint32_t)seed ^ 0x5EEDu, (uint32_t)(seed >> 32));
            for (int b = 0; b < 4; b++) act[b] = (int32_t)(((uint64_t)c[b] * 82) >> 32);
            act[4] = (int32_t)(((uint64_t)c2[0] * 242) >> 32);
            cc4o_step(e, act, msg, NULL, &r, &d);
            total += r; ticks++;
        }
    }
    cc4o_destroy(e);
    if (ret_sum) *ret_sum = total;
    return ticks;
}

/* capacity probe (sizing of the CUDA engine's fixed tables): runs the same episodes as cc4o_run_episodes and
 * records the maxima over all ticks of
 *   hw[0] processes on one host          hw[1] red sessions of all agents      hw[2] largest session ident
 *   hw[3] (abstract session, host) known-port entries                          hw[4] suspicious pids of all blue agents
 *   hw[5] FSM hosts of one red agent     hw[6] server_session ids of one agent hw[7] observation entries of one agent in a tick
 *   hw[8] event pids pending in a tick   hw[9] services on one host            hw[10] episodes with an error flag */
void cc4o_capacity_probe(uint64_t seed, uint64_t first_episode, int n_eps, int steps, int *hw)
{
    for (int i = 0; i < 11; i++) hw[i] = 0;
    cc4o_env *e = cc4o_create(seed, first_episode, steps, 0);
#define HW(i, v) do { int v_ = (v); if (v_ > hw[i]) hw[i] = v_; } while (0)
    for (int k = 0; k < n_eps; k++) {
        e->episode = first_episode + (uint64_t)k;
        cc4o_reset(e);
        for (int t = 0; t < steps - 1; t++) {
            int32_t act[NBLUE]; uint8_t msg[NBLUE] = { 0 }; float r; uint8_t d;
            uint32_t c[4] = { (uint32_t)t, 0xB10Eu, (uint32_t)e->episode, (uint32_t)(e->episode >> 32) };
            philox4x32_10(c, (uint32_t)seed ^ 0x5EEDu, (uint32_t)(seed >> 32));
            uint32_t c2[4] = { (uint32_t)t, 0xB10Fu, (uint32_t)e->episode, (uint32_t)(e->episode >> 32) };
            philox4x32_10(c2, (uint32_t)seed ^ 0x5EEDu, (uint32_t)(seed >> 32));
            for (int b = 0; b < 4; b++) act[b] = (int32_t)(((uint64_t)c[b] * 82) >> 32);
            act[4] = (int32_t)(((uint64_t)c2[0] * 242) >> 32);
            /* pending event pids are consumed by the end-turn Monitor inside the step: count them via the sus growth */
            int sus0 = 0; for (int b = 0; b < NBLUE; b++) sus0 += e->n_sus[b];
            cc4o_step(e, act, msg, NULL, &r, &d);
            int sus1 = 0; for (int b = 0; b < NBLUE; b++) sus1 += e->n_sus[b];
            HW(8, sus1 - sus0); HW(4, sus1);
            int ns = 0, np = 0;
            for (int s = 0; s < 128; s++) { HW(0, e->hosts[s].nproc); HW(9, e->hosts[s].nsvc); }
            for (int rr = 0; rr < NRED; rr++) {
                int a = A_RED0 + rr;
                ns += e->nsess[a];
                for (int i = 0; i < e->nsess[a]; i++) {
                    HW(2, e->sess[a][i].id);
                    if (e->sess[a][i].type == 1) for (int s = 0; s < NSLOT; s++) np += e->sess[a][i].has_ports[s];
                }
                HW(5, e->n_fsm[rr]); HW(6, e->n_ss[rr]); HW(7, e->obs[rr].n);
            }
            HW(1, ns); HW(3, np);
        }
        if (e->err) hw[10]++;
    }
#undef HW
    cc4o_destroy(e);
}

/* diagnostics: how often each BlueRewardMachine event fired since the last reset, out[3 phases][3 kinds: LWF ASF RIA][9 subnets] */
void cc4o_reward_breakdown(cc4o_env *e, int32_t *out) { memcpy(out, e->breakdown, sizeof e->breakdown); }
