/*
 * cc4_kernels.cuh — the kernels of the batched CC4 step engine (see cc4_engine.cu for the overview).  Included by
 * cc4_engine.cu once per NAMESPACE: `cc4` carries the tape paths (parity replay of recorded draws, rng_mode 1) and the reset
 * kernels; `cc4p` is the production build of k_step / k_rollout without a single tape-mode instruction (CC4_WITH_TAPE 0:
 * 11 % less code, 1.7 % faster, profiles/r04_variants.txt).  The host picks the namespace from the handle's rng_mode.
 */
#ifndef CC4_COMMON_DEFINED
#define CC4_COMMON_DEFINED
namespace cc4c {
struct Params {
    Cc4Dyn *dyn; Cc4Stat *stat; Cc4Cold *cold;
    int n_envs, steps, rng_mode, red_variant;
    uint32_t k0, k1hi; uint64_t first_episode;
    const cc4_draw *tape; const uint32_t *seg_off; uint32_t n_segs;
    unsigned long long *phase_clk;   /* [n_envs][8][2] cycles per phase (only in -DCC4_PHASE_CLOCKS builds) */
};
struct StepIO { const int32_t *actions; const uint8_t *msgs; uint8_t *obs; float *reward; uint8_t *done; int flags;
                int reverse; };     /* blocks take the episodes from the far end first (see cc4_step) */
}  // namespace cc4c
#endif

namespace CC4_NS {

#ifndef CC4_WARPS_PER_BLOCK
#define CC4_WARPS_PER_BLOCK 32  /* one block per SM: 1024 threads */
#endif
#ifndef CC4_EPB_MAX
#define CC4_EPB_MAX 64          /* episodes resident in one block's shared memory (64 x 3 600 B = 225 KB of the 227 KB): two full rounds of 32 warps */
#endif
#ifndef CC4_MERGE
#define CC4_MERGE 0
#endif
#ifndef CC4_SPLIT
#define CC4_SPLIT 0       /* bit 0: filter | blue execution, bit 1: phishing replay | red execution as separate barrier phases (k_step) */
#endif
#ifndef CC4_ENDTURN_PER
#define CC4_ENDTURN_PER 2       /* episodes per warp in the FSM digest phase (6 lanes of work per episode; 2 and 4 measure the same) */
#endif
#ifndef CC4_MIN_BLOCKS
#define CC4_MIN_BLOCKS 1        /* blocks of k_step per SM (tuning knob: 2 with CC4_WARPS_PER_BLOCK=16 CC4_EPB_MAX=32) */
#endif
constexpr int WARPS_PER_BLOCK = CC4_WARPS_PER_BLOCK;
constexpr int EPB_MAX = CC4_EPB_MAX;
constexpr int DYN_U4 = (int)(sizeof(Cc4Dyn) / 16);

using cc4c::Params;
using cc4c::StepIO;

/* ---- record movement ----
 * k_step moves the hot records with the bulk-copy engine (cp.async.bulk, 1-D TMA): at block start one thread per episode
 * starts the copy of its 2.7 KB record into shared memory and arms the episode's mbarrier with the byte count; the
 * warp that takes the episode in the first phase waits on that barrier.  All of a block's loads are in flight at once
 * (HBM-bound instead of one round trip per 512 B), and no warp issues load/store instructions for them.  The store at
 * the end of the tick is a bulk copy as well: the issuing lane only waits for its copies before it exits. */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, uint32_t bytes)
{   /* the warp's generic-proxy writes to the record must be visible to the async proxy that reads it */
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

/* plain 128-bit coalesced copies (k_rollout, k_reset: once per many ticks) */
__device__ __forceinline__ void w_load(Cc4Dyn *dst, const Cc4Dyn *src)
{
    const uint4 *g = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll 4
    for (int i = lane_id(); i < DYN_U4; i += 32) d[i] = __ldg(g + i);
    __syncwarp();
}
__device__ __forceinline__ void w_store(Cc4Dyn *dst, const Cc4Dyn *src)
{
    __syncwarp();
    const uint4 *d = reinterpret_cast<const uint4 *>(src);
    uint4 *g = reinterpret_cast<uint4 *>(dst);
#pragma unroll 4
    for (int i = lane_id(); i < DYN_U4; i += 32) g[i] = d[i];
}

/* ------------------------------------------------------------------------------------------------ */
/* observation vectors: BlueFlatWrapper.observation_change BlueFlatWrapper.py:172-256 (lane per element) */
/* ------------------------------------------------------------------------------------------------ */
/* The five vectors are assembled as ONE bit string (element p of agent b = bit mis + stride*b + p) and expanded into
 * bytes in registers, sixteen elements per lane, straight into 128-bit stores.  `mis` = the episode's byte offset inside
 * its 16-byte line (a row of 1 050 or 578 B is only 2-byte aligned): giving the bit string the same lead makes bit j and
 * global byte (line start + j) the same element, so whole lines go out aligned; only the first and the last line are
 * written by halfwords. */
constexpr int OBS_BYTES = CC4_N_BLUE * CC4_OBS_LEN;               /* 1050 */
static_assert(sizeof(((Scratch *)0)->obsbits) * 8 >= 14 + OBS_BYTES, "bit string too short");
__device__ __forceinline__ void bits_or(uint32_t *bits, int off, uint64_t v)
{   /* OR a <= 59-bit field into the bit string at bit offset off */
    int sh = off & 31;
    uint32_t *w = bits + (off >> 5);
    atomicOr(&w[0], (uint32_t)(v << sh));
    if (sh) { atomicOr(&w[1], (uint32_t)(v >> (32 - sh))); if (sh > 5) atomicOr(&w[2], (uint32_t)(v >> (64 - sh))); }
    else atomicOr(&w[1], (uint32_t)(v >> 32));
}
/* bits 0-8: one-hot own subnet, bits 18-26: "should be blocked" of the comms policy (BlueFlatWrapper.py:258-302), both in
 * sorted-subnet order, per mission phase and per (agent, own subnet) block: agents 0-3, then agent 4's ADM / OFF / PAZ;
 * from c_sorted_subnets and c_comms_adj (tools/gen_tables.py prints the same values) */
__constant__ uint32_t c_obs_fixed[3][7] = {
    { 0x2800080, 0x5FC0010, 0x4400100, 0x3FC0020, 0x0C40001, 0x0E00008, 0x1C00040 },
    { 0x6D80080, 0x7FC0010, 0x6400100, 0x3FC0020, 0x0C40001, 0x0E00008, 0x1C00040 },
    { 0x6800080, 0x5FC0010, 0x6D80100, 0x7FC0020, 0x0C40001, 0x0E00008, 0x1C00040 } };
/* compact: the five vectors without padding, 92 + 92 + 92 + 92 + 210 bytes per episode (CC4_STEP_OBS_COMPACT) */
/* as_bits: the compact row as a bit string, uint32 [n][20] (CC4_STEP_OBS_BITS): words 0-18 = entries, word 19 = mission phase */
__device__ void w_write_obs(const Cc4Dyn *s, Scratch *x, uint8_t *obs_base /* [n][5][210] or [n][578] or uint32 [n][20] */, int env, bool compact = false, bool as_bits = false)
{
    const int lane = lane_id();
    uint32_t *bits = x->obsbits;
    if (as_bits) compact = true;
    const int row_bytes = compact ? CC4_OBS_COMPACT_LEN : OBS_BYTES;
    const int vec_stride = compact ? 92 : CC4_OBS_LEN;      /* agents 0-3 have 92 entries, so agent b starts at 92 b either way for b <= 4 */
    const size_t goff = (size_t)env * row_bytes;
    const int mis = as_bits ? 0 : (int)(goff & 15);
    /* per-subnet 16-bit host masks: malicious-process / network-connection alerts (old + current events).  Lane L looks
       at hosts 4L..4L+3 (one word of ev), the four lanes of a quad cover one subnet. */
    const uint32_t evw = reinterpret_cast<const uint32_t *>(s->ev)[lane];
    uint32_t tp = (evw | evw >> 2) & 0x01010101u, tn = (evw >> 1 | evw >> 3) & 0x01010101u;       /* EV_PROC|EV_OLDPROC = bits 0,2; EV_NET|EV_OLDNET = bits 1,3 */
    tp = (tp | tp >> 7 | tp >> 14 | tp >> 21) & 0xFu; tn = (tn | tn >> 7 | tn >> 14 | tn >> 21) & 0xFu;
    uint32_t alerts = (tp | tn << 16) << (4 * (lane & 3));
    alerts |= __shfl_xor_sync(0xFFFFFFFFu, alerts, 1);
    alerts |= __shfl_xor_sync(0xFFFFFFFFu, alerts, 2);      /* low half: process alerts of the quad's subnet, high half: connection alerts */
    const int ob = lane < 4 ? lane : 4, oi = lane < 4 ? 0 : lane - 4, osn = lane < 7 ? blue_sorted_subnet(ob, oi) : 0;
    const uint32_t mine = __shfl_sync(0xFFFFFFFFu, alerts, 4 * osn);
    bits[lane] = 0;
    if (lane < 8) bits[32 + lane] = 0;
    __syncwarp();
    if (lane < 7) {   /* one lane per (agent, subnet) block of 59 bits */
        const int b = ob, i = oi, sn = osn;
        const uint32_t blk = s->blocks[sn];
        uint64_t v = c_obs_fixed[s->phase][lane];       /* one-hot subnet + comms policy of the mission phase */
#pragma unroll
        for (int j = 0; j < 9; j++) v |= (uint64_t)(blk >> c_sorted_subnets[j] & 1) << (9 + j);
        const uint32_t exist = ((1u << s->nuser[sn]) - 1u) << 6 | ((1u << s->nserver[sn]) - 1u);
        const uint32_t pm = mine & exist, nm = mine >> 16 & exist;
        v |= (uint64_t)pm << 27 | (uint64_t)nm << 43;
        bits_or(bits, mis + vec_stride * b + 1 + 59 * i, v);
    }
    if (lane >= 8 && lane < 13) {   /* messages from the other four agents in sorted order, SimulationController.py:781-804 */
        int b = lane - 8;
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) m |= (uint32_t)s->msgs[k < b ? k : k + 1] << (8 * k);
        bits_or(bits, mis + vec_stride * b + 1 + 59 * (b == 4 ? 3 : 1), (uint64_t)m);
    }
    __syncwarp();
    if (as_bits) {      /* the string itself: 19 words + the mission phase, one 80-byte row per episode */
        uint32_t *row = reinterpret_cast<uint32_t *>(obs_base) + (size_t)env * CC4_OBS_BITS_WORDS;
        if (lane < CC4_OBS_BITS_WORDS) row[lane] = lane < CC4_OBS_BITS_WORDS - 1 ? bits[lane] : (uint32_t)s->phase;
        __syncwarp();
        return;
    }
    /* expand and store: global byte (line start + j) = bit j of the string.  A lane turns 16 bits into one 16-byte line in
       registers (4 bits -> 4 bytes by one multiply) and stores it with one 128-bit store; element 0 of every vector is
       the mission phase, not a bit, and is patched in; the partial first and last line go out by halfwords. */
    const uint16_t *half = reinterpret_cast<const uint16_t *>(bits);
    uint8_t *gline = obs_base + (goff - (size_t)mis);
    const int end = mis + row_bytes;
    auto spread4 = [](uint32_t nib) -> uint32_t { return (nib * 0x00204081u) & 0x01010101u; };
    _Pragma("unroll 1") for (int q = lane; q * 16 < end; q += 32) {
        const int lo = q * 16;
        if (lo >= mis && lo + 16 <= end) {
            const uint32_t h = half[q];
            reinterpret_cast<uint4 *>(gline)[q] = make_uint4(spread4(h & 15u), spread4(h >> 4 & 15u), spread4(h >> 8 & 15u), spread4(h >> 12));
        }
    }
    {   /* lanes 0-7: halfwords of the first line, lanes 8-15: of the last line */
        const int k = lane < 8 ? 2 * lane : (end & ~15) + 2 * (lane - 8);
        const bool head = lane < 8 && mis > 0 && k >= mis, tail = lane >= 8 && lane < 16 && (end & 15) && k < end;
        if (head || tail) {
            const uint32_t two = bits[k >> 5] >> (k & 31) & 3u;        /* k is even: both bits are in one word */
            *reinterpret_cast<uint16_t *>(gline + k) = (uint16_t)((two & 1u) | (two >> 1) << 8);
        }
    }
    __syncwarp();       /* orders the line stores before the five byte stores below (same warp, overlapping addresses) */
    if (lane < CC4_N_BLUE) gline[mis + vec_stride * lane] = s->phase;       /* element 0 of every vector is the mission phase, not a bit */
    __syncwarp();
}
__device__ void w_write_mask(const Cc4Dyn *s, uint8_t *mask)
{
    _Pragma("unroll 1") for (int i = lane_id(); i < CC4_N_BLUE * CC4_MASK_LEN; i += 32) {
        int b = i / CC4_MASK_LEN, k = i % CC4_MASK_LEN;
        mask[i] = (uint8_t)(k < (b == 4 ? 242 : 82) ? blue_decode(s, b, k) >> 31 : 0u);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* end-turn Monitor for all five blue agents (Monitor.py:35-74), lane per host                      */
/* ------------------------------------------------------------------------------------------------ */
__device__ void w_monitor_all(Cx c)
{
    Cc4Dyn *s = DYN(c);
    {   /* events -> old events on every host some blue agent monitors, four hosts (one word of ev) per lane: all subnets
           but the contractor network (slots 64-79, router 132) and the internet host (136) */
        uint32_t *evw = reinterpret_cast<uint32_t *>(s->ev);
        const int lane = lane_id();
        if ((lane >> 2) != S_CON) evw[lane] = (evw[lane] & 0x03030303u) << 2;          /* (EV_PROC | EV_NET) -> (EV_OLDPROC | EV_OLDNET) */
        if (lane < 2) { const uint32_t keep = lane == 1 ? 0x000000FFu : 0u, w = evw[32 + lane]; evw[32 + lane] = (w & keep) | (((w & ~keep) & 0x03030303u) << 2); }
    }
    __syncwarp();
    if (lane_id() == 0) {
        Scratch *x = SCR(c);
        _Pragma("unroll 1") for (int i = 0; i < x->n_evpid; i++) {
            int slot = x->evpid[i] >> 16, b = blue_of_subnet(slot_subnet(slot));
            if (s->n_sus >= CC4_SUS_CAP) { fail(c, CC4_E_CAPACITY); break; }
            WC(c)->k->sus[s->n_sus++] = SU_MAKE(x->evpid[i] & 0xFFFFu, slot, b);
        }
        x->n_evpid = 0;
    }
    __syncwarp();
}
/* ------------------------------------------------------------------------------------------------ */
/* session-pool digests (lane per session)                                                          */
/* ------------------------------------------------------------------------------------------------ */
__device__ void w_scan_sessions(Cx c)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    const int lane = lane_id();
    if (lane < 24) (&x->sess_hosts[0][0])[lane] = 0;
    if (lane < 8) x->n_red[lane] = 0;
    if (lane == 8) { x->has0 = 0; x->misplaced = 0; }
    __syncwarp();
    const int n = s->n_sess;
    _Pragma("unroll 1") for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        uint32_t r = 7, h = 0, id = 1;
        if (i < n) { const uint32_t w = s->sess[i]; r = SF_AGENT((uint32_t)s->sessf[i]); h = S_HOST(w); id = S_ID(w); }
        const uint32_t same = __match_any_sync(0xFFFFFFFFu, r);        /* lanes whose session belongs to the same agent */
        if (i < n) {
            atomicOr(&x->sess_hosts[r][h >> 5], 1u << (h & 31));
            if ((int)__ffs(same) - 1 == lane) atomicAdd(&x->n_red[r], __popc(same));    /* one add per agent and chunk */
            if (id == 0) atomicOr(&x->has0, 1u << r);
            if (!red_allowed((int)r, (int)(h >> 4))) atomicOr(&x->misplaced, 1u << r);
        }
    }
    __syncwarp();
}
__device__ void w_build_hostagents(Cx c)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    x->g.hostagents[lane_id()] = 0;
    __syncwarp();
    _Pragma("unroll 1") for (int i = lane_id(); i < s->n_sess; i += 32) {
        uint32_t h = S_HOST(s->sess[i]);
        atomicOr(&x->g.hostagents[h >> 2], (1u << SF_AGENT(s->sessf[i])) << ((h & 3) * 8));
    }
    __syncwarp();
}

/* ------------------------------------------------------------------------------------------------ */
/* one tick: SimulationController.step SimulationController.py:211-315, cut into phases               */
/* ------------------------------------------------------------------------------------------------ */
/* One block per SM keeps up to CC4_EPB_MAX episodes resident in shared memory.  A tick is a sequence of PHASES separated
 * by block barriers; inside a phase the block's warps take episodes from a shared counter, so (a) every warp of the SM
 * executes the same phase of the tick at the same time and they share its instruction-cache lines (ncu: without
 * barriers the kernel is 1.7x slower, stall_no_inst on top), and (b) a warp that drew a cheap episode (no action due)
 * takes another one instead of waiting at the barrier for the warp that drew an expensive one.  Everything that lives
 * across a phase boundary is in the episode's shared-memory record / scratch, never in registers. */
enum { PH_SETUP, PH_POLICY, PH_BLUE, PH_GREEN, PH_RED, PH_REASSIGN, PH_ENDTURN, PH_FINISH, N_PHASES };
/* The action phases are dominated by their most expensive episodes (an exploit or a Restore against nothing due).  The
 * policy phase sorts the block's episodes into four cost classes from the due actions; the action phases hand them out
 * heaviest class first, so the cheap episodes fill the tail instead of leaving most warps idle behind a late expensive one. */
#ifndef CC4_TYPE_ORDER
#define CC4_TYPE_ORDER 0        /* 1 (with CC4_ORDER_CLASSES=8): order the action phases by the kind of the first due action instead of by summed cost */
#endif
#ifndef CC4_ORDER_CLASSES
#define CC4_ORDER_CLASSES 8     /* cost classes of the two action phases: 8 (measured action costs, profiles/r02s_phase_clocks.txt; -0.6 %) or 4 (coarse weights) */
#endif
constexpr int NCLS = CC4_ORDER_CLASSES;
struct PhaseOrder { int cnt[NCLS]; uint8_t list[NCLS][64]; };
__device__ __forceinline__ void order_push(PhaseOrder *o, int cls, int e)      /* one lane per episode calls this */
{
    int k = atomicAdd(&o->cnt[cls], 1); o->list[cls][k] = (uint8_t)e;
}
__device__ __forceinline__ int order_get(const PhaseOrder *o, int i)
{
#pragma unroll
    for (int cls = NCLS - 1; cls > 0; cls--) { int n = o->cnt[cls]; if (i < n) return o->list[cls][i]; i -= n; }
    return o->list[0][i];
}
#ifdef CC4_PHASE_CLOCKS   /* profiling build (tools/phase_clocks.py): per episode [18]: [2p] = busy cycles of phase p; [2p+1] of a block's
                             first episode = wall cycles of phase p for the block, barrier included; [17] = the episode's slowest red
                             phase, [16] of a block's first episode = the block's slowest red phase */
__device__ unsigned long long *g_phase_clk;
__device__ unsigned long long *g_action_clk;       /* [64][2]: cycles and count per action code (red and blue execution) */
#define PC_STRIDE 18
#endif

/* ---- phase 0: per-tick scratch, known-session mirror, mission phase ---- */
__device__ __forceinline__ void ph_setup(Cx c)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    const int lane = lane_id();
    WC(c)->tick = s->step;
    {   /* draw counters: ctr_br (24 B) + ctr_g (80 B) are contiguous */
        static_assert(offsetof(Scratch, ctr_g) == offsetof(Scratch, ctr_br) + 24, "counter arrays must be adjacent");
        uint32_t *z = reinterpret_cast<uint32_t *>(x->ctr_br);
        if (lane < 26) z[lane] = 0;
    }
    if (lane < 8) { x->n_oent[lane] = 0; x->have_first[lane] = 0; x->succ0[lane] = SUCC_UNKNOWN; x->has_act0[lane] = 0; }
    if (lane <= CC4_NRED) x->blk_id[lane] = 0xFFFFFFFFu;
    if (lane >= 8 && lane < 11) x->g.phish[lane - 8] = 0;
    if (lane == 0) {
        x->n_evpid = 0; x->reward = 0; x->newss = 0;
        /* mission phase: State.check_next_phase_on_update_step State.py:514-544, phases EnterpriseScenarioGenerator.py:854-860 */
        int q = g_blk.steps / 3, rem = g_blk.steps % 3, l0 = q + (rem >= 1), l1 = q + (rem == 2);
        int np = WC(c)->tick < l0 ? 0 : WC(c)->tick < l0 + l1 ? 1 : 2;
        if (WC(c)->tick >= g_blk.steps) fail(c, CC4_E_UNREACHABLE);
        if (np > s->phase) s->phase = (uint8_t)np;
    }
    __syncwarp();
    if (lane < 8) {     /* GreenAccessService destination table of this tick's mission phase, one lane per source subnet */
        uint32_t allowed = c_green_allowed[s->phase][lane];
        uint64_t cum = 0; uint32_t tot = 0;
#pragma unroll
        for (int sn = 0; sn < 8; sn++) { if (allowed >> sn & 1) tot += s->nserver[sn]; cum |= (uint64_t)tot << (8 * sn); }
        x->g.as_cum[lane] = cum;
    }
    __syncwarp();
}

/* ---- phase 1: policies (:236-248) and duration queue (:245-265): one lane per blue / red actor.  The scripted red
   agents only read their own knowledge and draw from their own stream, so the six run side by side.  A work item is the
   blue half (index decoding, 5 lanes) or the red half (FiniteStateRed choice, 6 lanes) of FOUR episodes, 8 lanes each
   (`sl` = lane inside the group, `c` differs per group), so a warp never mixes the two code paths; `live` = the group has
   an episode that is still running.  `actions` = the episode's five BlueFixedActionWrapper
   indices, or null for the benchmark's synthetic policy (k_rollout): index ~ U{0..81} (agents 0-3) / U{0..241} (agent 4),
   Philox keyed (policy_seed, episode, tick) exactly as oracle/cc4_oracle.c cc4o_run_episodes.  Also sorts the episode into
   the cost classes of the two action phases (each_episode). ---- */
/*                                          Sl Mo An Rm Rs De Bl Al In  -  DRS Ag St DD Ex Pr Im Dg Wd */
#if CC4_ORDER_CLASSES == 8      /* mean cycles of the executed action / 256 (profiles/r02s_phase_clocks.txt) */
__constant__ uint8_t c_act_weight[32] = { 0, 5, 0, 9, 13, 14, 1, 1, 0, 0, 9, 13, 16, 14, 20, 27, 10, 11, 12 };
__device__ __forceinline__ int blue_class(int w) { return w == 0 ? 0 : min(7, 1 + (w - 1) / 6); }
__device__ __forceinline__ int red_class(int w) { return w == 0 ? 0 : min(7, 1 + (w - 1) / 9); }
#else
__constant__ uint8_t c_act_weight[32] = { 0, 2, 0, 4, 6, 4, 1, 1, 0, 0, 5, 3, 3, 3, 6, 4, 3, 3, 4 };
__device__ __forceinline__ int blue_class(int w) { return w == 0 ? 0 : w <= 3 ? 1 : w <= 7 ? 2 : 3; }
__device__ __forceinline__ int red_class(int w) { return w == 0 ? 0 : w <= 4 ? 1 : w <= 8 ? 2 : 3; }
#endif
template <int RV, bool SYNTH = false> __device__ __forceinline__ void ph_policy(Cx c, int e, bool exists, bool live, int sl, bool red_part, const int32_t *actions, uint32_t pk0, uint32_t pk1,
                                          uint64_t ep, PhaseOrder *ord_blue, PhaseOrder *ord_red)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    int cost = 0, wgt = 0;
    if (live && sl < (red_part ? CC4_NRED : CC4_NBLUE)) {
        const int actor = red_part ? A_RED0 + sl : sl;
        uint32_t aw = ACT_SLEEP_WORD;
        if (!red_part) {
            int idx;
            if (!SYNTH) idx = actions[sl];      /* SYNTH: the benchmark's synthetic policy (k_rollout), compiled out of k_step */
            else {
                uint32_t cw[4] = { (uint32_t)s->step, sl == 4 ? 0xB10Fu : 0xB10Eu, (uint32_t)ep, (uint32_t)(ep >> 32) };
                philox4x32_10(cw, pk0, pk1);
                idx = sl == 4 ? (int)__umulhi(cw[0], 242u) : (int)__umulhi(sl == 0 ? cw[0] : sl == 1 ? cw[1] : sl == 2 ? cw[2] : cw[3], 82u);
            }
            aw = blue_decode(s, sl, idx) & ~ACT_VALID;
            if (IP_CODE(aw) == AC_RESTORE) cost = -1;      /* Restore.py:70, charged on submission :310-311 */
        }
        else if (s->red_active >> sl & 1) aw = red_policy<RV>(c, sl);
        uint32_t ip = s->inprog[actor], due = ACT_SLEEP_WORD;
        if (IP_CODE(ip) == IP_NONE) ip = aw | (uint32_t)c_duration[IP_CODE(aw)] << 26;
        int ticks = (int)IP_TICKS(ip) - 1;
        if (ticks < 1) { due = ip & ~(0xFu << 26); s->inprog[actor] = IP_NONE; }
        else {
            s->inprog[actor] = (ip & ~(0xFu << 26)) | (uint32_t)ticks << 26;
            if (red_part) obs_push(c, sl, SUCC_INPROG, false, ACT_SLEEP_WORD);
        }
        x->subw[actor] = due;
        wgt = c_act_weight[IP_CODE(due) & 31u];
    }
#if CC4_TYPE_ORDER       /* class = the kind of the episode's FIRST due action (what its warp runs first), heaviest kinds first: warps that run side by
                            side then share the action's code in the instruction cache */
    int first_code = 0xFFFF;
    if (live && sl < (red_part ? CC4_NRED : CC4_NBLUE)) { const uint32_t cd = IP_CODE(x->subw[red_part ? A_RED0 + sl : sl]); if (wgt > 0) first_code = sl << 8 | (int)cd; }
#endif
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        cost += __shfl_xor_sync(0xFFFFFFFFu, cost, o); wgt += __shfl_xor_sync(0xFFFFFFFFu, wgt, o);     /* sums over the episode's 8 lanes */
#if CC4_TYPE_ORDER
        first_code = min(first_code, __shfl_xor_sync(0xFFFFFFFFu, first_code, o));
#endif
    }
    if (exists && sl == 0) {
#if CC4_TYPE_ORDER
        const int fc = first_code & 0xFF;
        /*                                   Sl Mo An Rm Rs De Bl Al In  -  DRS Ag St DD Ex Pr Im Dg Wd */
        const uint64_t kind = 0x0ull | 1ull << 4 | 0ull << 8 | 4ull << 12 | 6ull << 16 | 5ull << 20 | 2ull << 24 | 2ull << 28 | 0ull << 32 | 0ull << 36;
        const uint64_t kindr = 1ull | 3ull << 4 | 5ull << 8 | 4ull << 12 | 7ull << 16 | 6ull << 20 | 2ull << 24 | 2ull << 28 | 3ull << 32;      /* from AC_DRS */
        if (!red_part) { if (live) x->cost = cost; order_push(ord_blue, first_code == 0xFFFF ? 0 : (int)(kind >> (4 * fc) & 15u), e); }
        else order_push(ord_red, first_code == 0xFFFF || fc < AC_DRS ? 0 : (int)(kindr >> (4 * (fc - AC_DRS)) & 15u), e);
#else
        if (!red_part) { if (live) x->cost = cost; order_push(ord_blue, blue_class(wgt), e); }
        else order_push(ord_red, red_class(wgt), e);
#endif
    }
    __syncwarp();
    /* The due actions are known: pull the cold-record lines they will touch towards L2 now, so that the dependent reads of
       the (serial) action code of the next phases find them there instead of in DRAM (worth 3 % of the blue phase). */
    if (live) {
        const Cc4Cold *k = WC(c)->k;
        if (sl < (red_part ? CC4_NRED : CC4_NBLUE)) {   /* directory entry, home slab and largest pid of the actor's target host */
            const uint32_t w = x->subw[red_part ? A_RED0 + sl : sl], code = IP_CODE(w), h = IP_HOST(w);
            const bool hosty = (code == AC_REMOVE || code == AC_RESTORE || code == AC_DECOY || (code >= AC_AGGR && code <= AC_WITHDRAW)) && h < 128;
            if (hosty) {
                prefetch_l2(&k->pdir[h]); prefetch_l2(&k->parena[h * CC4_OSLAB]); prefetch_l2(&k->pmax[h]);
                if (code == AC_RESTORE) { const Cc4Stat *st = WC(c)->st; prefetch_l2(&st->o_procs[h][0]); prefetch_l2(&st->o_svc64[h]); prefetch_l2(&st->o_nproc[h]); }
            }
        }
        if (!red_part) {        /* the service table (green LocalWork reads one entry per agent every tick) and the suspicious pids */
            const char *b = reinterpret_cast<const char *>(k->svc64);
            prefetch_l2(b + 128 * sl);
            if (sl == 0) prefetch_l2(b + sizeof(k->svc64) - 1);
            if (sl * 32 < s->n_sus) prefetch_l2(k->sus + sl * 32);
        } else {                /* known ports: 12 lines = 384 entries */
            if (sl * 32 < s->n_ports) prefetch_l2(k->ports + sl * 32);
            if ((sl + 8) * 32 < s->n_ports && sl < 4) prefetch_l2(k->ports + (sl + 8) * 32);
        }
    }
}

/* ---- phase 2: filter_actions (:466-486), then blue execution ---- */
__device__ __forceinline__ void ph_blue_exec(Cx c, uint32_t agents = 0x1Fu, bool with_prio1 = true);
__device__ __forceinline__ void ph_blue(Cx c, bool with_exec = true, uint32_t agents = 0x1Fu)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    const int lane = lane_id();
    /* filter_actions looks at the sessions as they are BEFORE anything executes; blue / green sessions never go away */
    /* (the pool index found here stays valid for the red phase unless a session is deleted in between: sess_moved) */
    {
        /* every pool entry belongs to one red agent and (agent, ident) is unique, so a lane only compares its entry's ident
           with the one its owner's action names and, on a match, files the pool index under that agent */
        const int n = s->n_sess;
        uint32_t want = 0;      /* lanes 0..5: ident | needs << 8 */
        if (lane < CC4_NRED) {
            const uint32_t w = x->subw[A_RED0 + lane];
            want = IP_SESSION(w) | (uint32_t)!(IP_CODE(w) == AC_SLEEP || IP_CODE(w) == AC_INVALID) << 8;
            x->act_si[lane] = -1;
        }
        const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, want, 0) | __shfl_sync(0xFFFFFFFFu, want, 1) << 10 | __shfl_sync(0xFFFFFFFFu, want, 2) << 20;
        const uint32_t p1 = __shfl_sync(0xFFFFFFFFu, want, 3) | __shfl_sync(0xFFFFFFFFu, want, 4) << 10 | __shfl_sync(0xFFFFFFFFu, want, 5) << 20;
        __syncwarp();
        _Pragma("unroll 1") for (int i = lane; i < n; i += 32) {
            const uint32_t r = SF_AGENT((uint32_t)s->sessf[i]);
            const uint32_t wr = (r < 3 ? p0 >> (10 * r) : p1 >> (10 * (r - 3))) & 0x3FFu;
            if ((wr >> 8) && (wr & 0xFFu) == S_ID(s->sess[i])) x->act_si[r] = (int16_t)i;
        }
        __syncwarp();
        if (lane < CC4_NRED) x->dropped[A_RED0 + lane] = (want >> 8) && x->act_si[lane] < 0;
        if (lane == 0) x->sess_moved = 0;
    }
    __syncwarp();
    if (with_exec) ph_blue_exec(c, agents);
}
/* `agents`: the blue agents whose ordinary actions run in this call (the phase can be cut by agent, CC4_BLUE_PARTS); the
   priority-1 actions of ALL agents run in the call that has with_prio1 */
__device__ __forceinline__ void ph_blue_exec(Cx c, uint32_t agents, bool with_prio1)
{
    Scratch *x = SCR(c);
    const int lane = lane_id();
    /* execution: priority-1 ControlTraffic first (ControlTraffic.py:19), then agent order; warp-uniform: every lane runs the
       same action code, list scans inside are lane-parallel */
    {
        const uint32_t code = lane < CC4_NBLUE ? IP_CODE(x->subw[lane]) : (uint32_t)AC_SLEEP;
        const bool prio1 = code == AC_BLOCK || code == AC_ALLOW;
        uint32_t first = with_prio1 ? __ballot_sync(0xFFFFFFFFu, prio1) : 0u, then = __ballot_sync(0xFFFFFFFFu, !prio1 && code != AC_SLEEP && code != AC_ANALYSE) & agents;
        while (first) { const int b = __ffs(first) - 1; first &= first - 1; exec_blue(c, b, x->subw[b]); }
        while (then) {
            const int b = __ffs(then) - 1; then &= then - 1;
#ifdef CC4_PHASE_CLOCKS
            const long long ta0 = clock64();
            exec_blue(c, b, x->subw[b]);
            if (lane_id() == 0 && g_action_clk) { atomicAdd(&g_action_clk[2 * IP_CODE(x->subw[b])], (unsigned long long)(clock64() - ta0)); atomicAdd(&g_action_clk[2 * IP_CODE(x->subw[b]) + 1], 1ull); }
#else
            exec_blue(c, b, x->subw[b]);
#endif
        }
    }
    __syncwarp();
}

/* ---- phase 3: green agents (EnterpriseGreenAgent.py:42-78).  A green agent reads its host's services and the block
   table and ORs event bits: agents are independent of each other, also across episodes.  The block's green agents are laid
   out as ONE flat list (episode-major) and handed out 32 at a time, so every lane of every item has an agent whatever the
   episodes' sizes (24-80 agents) are; penalties and phishing flags go to the owning episode's scratch with shared-memory
   atomics.  The phishing side effect is order dependent: it is replayed in agent order at the start of the red phase. ---- */
/* inclusive prefix of the episodes' green-agent counts, by ONE warp, at least one block barrier before green_items reads it
   (called between the policy and the blue phase: the counts are final after setup) */
__device__ __forceinline__ void green_prefix(int n_local, const uint8_t *ng_s, short *gend)
{
    const int lane = lane_id();
    int ca = lane < n_local ? ng_s[lane] : 0, cb = 32 + lane < n_local ? ng_s[32 + lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int ta = __shfl_up_sync(0xFFFFFFFFu, ca, o), tb = __shfl_up_sync(0xFFFFFFFFu, cb, o); if (lane >= o) { ca += ta; cb += tb; } }
    cb += __shfl_sync(0xFFFFFFFFu, ca, 31);
    if (lane < n_local) gend[lane] = (short)ca;
    if (32 + lane < n_local) gend[32 + lane] = (short)cb;
}
__device__ __forceinline__ void green_items(int ph, int first, int n_local, int *next, const uint8_t *ng_s, short *gend)
{
#ifdef CC4_PHASE_CLOCKS
    const long long tb0 = clock64();
#endif
    const int lane = lane_id();
    const int total = n_local > 0 ? gend[n_local - 1] : 0;        /* the prefix table was written by green_prefix, a block barrier ago */
    const int n_items = (total + 31) >> 5;
    /* items cost the same: warp w takes items w, w + 32, ... (CC4_GREEN_FIXED=0: from the shared counter) */
    _Pragma("unroll 1") for (int it_fixed = (int)(threadIdx.x >> 5);; it_fixed += WARPS_PER_BLOCK) {
        int it = it_fixed;
#if !CC4_GREEN_FIXED
        if (lane == 0) it = atomicAdd(&next[ph], 1);
        it = __shfl_sync(0xFFFFFFFFu, it, 0);
#endif
        if (it >= n_items) break;
#ifdef CC4_PHASE_CLOCKS
        const long long t0 = clock64();
#endif
        const int gidx = it * 32 + lane;
        int e = 0;
        /* the item starts in the first episode whose end is past it * 32 (one lane-parallel look at the prefix table); an
           episode has at least 24 agents, so a lane's own episode is at most two further on */
        const uint32_t past_a = __ballot_sync(0xFFFFFFFFu, lane < n_local && gend[lane] > it * 32);
        const uint32_t past_b = __ballot_sync(0xFFFFFFFFu, 32 + lane < n_local && gend[32 + lane] > it * 32);
        const int e_first = past_a ? __ffs(past_a) - 1 : 31 + __ffs(past_b);
        if (gidx < total) {
            e = e_first;
            while (gend[e] <= gidx) e++;
            const Cx c = (Cx)(e * (int)sizeof(WarpSmem));
            const int g = gidx - (e ? gend[e - 1] : 0);
            Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
            const int gr = green_tick(c, g, reinterpret_cast<uint32_t *>(s->ev));
            if (gr & 0xFF) atomicSub(&x->reward, gr & 0xFF);
            if (gr & GREEN_PHISH) atomicOr(&x->g.phish[g >> 5], 1u << (g & 31));
        }
        __syncwarp();
#ifdef CC4_PHASE_CLOCKS
        e = __shfl_sync(0xFFFFFFFFu, e, 0);
        if (lane == 0 && g_phase_clk) g_phase_clk[(size_t)(first + e) * PC_STRIDE + 2 * ph] += (unsigned long long)(clock64() - t0);
#endif
    }
    __syncthreads();
#ifdef CC4_PHASE_CLOCKS
    if (threadIdx.x == 0 && g_phase_clk && n_local > 0) g_phase_clk[(size_t)first * PC_STRIDE + 2 * ph + 1] += (unsigned long long)(clock64() - tb0);
#endif
}

/* ---- phase 4: red execution in agent order ---- */
__device__ __forceinline__ void ph_phishing(Cx c)
{
    Scratch *x = SCR(c);
    /* phishing e-mails of this tick's green LocalWork (GreenLocalWork.py:118-124), in green-agent order */
    if (x->g.phish[0] | x->g.phish[1] | x->g.phish[2]) {
        w_build_hostagents(c);
        _Pragma("unroll 1") for (int k = 0; k < 3; k++) { uint32_t m = x->g.phish[k]; while (m) { int b = __ffs(m) - 1; m &= m - 1; do_phishing(c, k * 32 + b); } }
        __syncwarp();
    }
}
/* `agents`: the red agents this call handles (the phase can be cut by agent, CC4_RED_PARTS: the same code runs again on
   the next agents after a block barrier, which halves the heaviest work item of the phase) */
template <int RV> __device__ __forceinline__ void ph_red(Cx c, bool with_phishing = true, uint32_t agents = 0x3Fu)
{
    Scratch *x = SCR(c);
    if (with_phishing) ph_phishing(c);
    {
        const int lane = lane_id();
        bool real = false;
        if (lane < CC4_NRED && (agents >> lane & 1u) && !x->dropped[A_RED0 + lane]) {
            const uint32_t aw = x->subw[A_RED0 + lane], code = IP_CODE(aw);
            if (code == AC_SLEEP) obs_push(c, lane, SUCC_UNKNOWN, true, aw);        /* nothing to execute: the lane files the observation */
            else if (code == AC_INVALID) obs_push(c, lane, SUCC_FALSE, true, aw);
            else real = true;
        }
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, real);
        while (todo) {                  /* warp-uniform, in agent order */
            const int r = __ffs(todo) - 1; todo &= todo - 1;
            const uint32_t aw = x->subw[A_RED0 + r];
#ifdef CC4_PHASE_CLOCKS
            const long long ta0 = clock64();
            const int succ = exec_red<RV>(c, r, aw);
            if (lane_id() == 0 && g_action_clk) { atomicAdd(&g_action_clk[2 * IP_CODE(aw)], (unsigned long long)(clock64() - ta0)); atomicAdd(&g_action_clk[2 * IP_CODE(aw) + 1], 1ull); }
#else
            const int succ = exec_red<RV>(c, r, aw);
#endif
            obs_push(c, r, succ, true, aw);
        }
    }
    __syncwarp();
}

/* ---- phase 5: reassignment (:278), then the warp-wide half of the end-turn actions (:281-286): Monitor for blue,
   the primary-session re-election of RedSessionCheck ---- */
template <int RV> __device__ __forceinline__ void ph_reassign(Cx c)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    w_scan_sessions(c);
    if (x->misplaced) {
        reassign(c);
        __syncwarp();
        w_scan_sessions(c);
    }
    const int lane = lane_id();
    const bool holds = lane < CC4_NRED && x->n_red[lane] > 0;
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, holds);                                   /* :896-903: active iff it holds a session */
    /* RedSessionCheck re-elects a primary session when ident 0 is gone (changes the session order: agent by agent) */
    uint32_t elect = __ballot_sync(0xFFFFFFFFu, holds && s->sess_count[lane] != 0 && !(x->has0 >> lane & 1));
    if (lane == 0) { x->active = active; s->red_active = (uint8_t)active; }
    w_monitor_all(c);
    while (elect) { const int r = __ffs(elect) - 1; elect &= elect - 1; rsc_reelect<RV>(c, r, x->n_red[r]); }
    __syncwarp();
    if (x->newss) learn_new_sessions(c);        /* the observation lists the sessions as the re-election left them */
    __syncwarp();
}

/* ---- phase 6: RedSessionCheck observation + FSM digest of every active red agent, one lane per agent; a warp takes
   FOUR episodes (8 lanes each) ---- */
template <int RV> __device__ __forceinline__ void ph_endturn(Cx c, bool live, int sl)
{
    if (live && sl < CC4_NRED && (SCR(c)->active >> sl & 1)) {
        obs_push(c, sl, SUCC_TRUE, true, IP_MAKE(AC_RSC, 255, 15, 0, 0));
        /* the FSM agents digest the tick's observation now instead of at the start of the next tick */
        agent_end_turn<RV>(c, sl);
    }
}

/* ---- phase 7: counters, done (:297-300), reward (BlueRewardMachine.py:70-121), messages ---- */
__device__ __forceinline__ void ph_finish(Cx c, const uint8_t *msgs /* [5] or null */, float *reward_out, bool accumulate, uint8_t *done_out)
{
    Cc4Dyn *s = DYN(c); Scratch *x = SCR(c);
    const int lane = lane_id();
    const uint32_t active = x->active;
    int rew = 0;
    if (lane < CC4_NRED) {
        /* `success and isinstance(action, Impact)` BlueRewardMachine.py:118: a TernaryEnum member is always truthy
           (Enums.py:5-24), so every Impact that comes due costs RIA while the agent holds a session */
        const uint32_t w = x->subw[A_RED0 + lane];
        if (IP_CODE(w) == AC_IMPACT && (active >> lane & 1)) rew = c_reward[s->phase][IP_HOST(w) >> 4][2];
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) rew += __shfl_xor_sync(0xFFFFFFFFu, rew, o);       /* lanes 0-7 */
    if (lane >= 8 && lane < 8 + CC4_NBLUE) s->msgs[lane - 8] = msgs ? msgs[lane - 8] : 0;
    if (CC4_TAPE_ON())
        _Pragma("unroll 1") for (int a = lane; a < NAG; a += 32) {
            uint32_t seg = 1u + (uint32_t)WC(c)->tick * NAG + (uint32_t)a;
            if (seg < g_blk.n_segs && ctr_get(x, a) != WC(c)->seg[seg + 1] - WC(c)->seg[seg]) fail(c, CC4_E_TAPE_LEFT);
        }
    if (lane == 0) {
        s->fsm_started |= (uint8_t)active;
        const int step = s->step + 1;
        s->step = (uint16_t)step;
        s->done = step >= g_blk.steps - 1;
        const float r = (float)(rew + x->reward + x->cost);
        if (reward_out) { if (accumulate) *reward_out += r; else *reward_out = r; }
        if (done_out) *done_out = s->done;
    }
    __syncwarp();
}

/* ---- episode scheduling inside a phase ----
 * (Measured alternative, r01k: per-phase work queues without block barriers — a warp that finishes phase p on one episode
 * goes on to phase p+1 of whatever episode is ready — raise warp utilisation but let two phases be live at once; the
 * instruction-cache misses cost more than the idle time saved: 1.08 ms against 0.97 ms per tick.) */
#ifndef CC4_RED_PARTS
#define CC4_RED_PARTS 1         /* k_step: the red phase as 1, 2 or 3 passes over the episodes (2 or 3: by agent, a block barrier in between) */
#endif
#ifndef CC4_BLUE_PARTS
#define CC4_BLUE_PARTS 1
#endif
#ifndef CC4_GREEN_FIXED
#define CC4_GREEN_FIXED 1
#endif
#ifndef CC4_GROUP_FIXED
#define CC4_GROUP_FIXED 1       /* the lane-per-actor phases take item = warp instead of drawing from the shared counter */
#endif
#ifndef CC4_FIXED_PHASES
#define CC4_FIXED_PHASES 5      /* k_step phases with a fixed episode -> warp map: bit 0 setup, bit 1 reassignment, bit 2 finish (profiles/r02t_variants.txt) */
#endif
#ifndef CC4_STATIC_ASSIGN
#define CC4_STATIC_ASSIGN 0     /* 1: episode e of a block is always handled by warp e % warps (tuning knob; default: shared counter) */
#endif
template <bool FIXED = false, bool SYNC = true, class F> __device__ __forceinline__ void each_episode(int ph, int first, int n_local, int *next, F f, const PhaseOrder *order = nullptr, int ctr = -1)
{
    if (ctr < 0) ctr = ph;      /* which work counter the phase uses (the parts of a cut phase share `ph`) */
#ifdef CC4_PHASE_CLOCKS
    const long long tb0 = clock64();
#endif
    /* FIXED: phases whose episodes cost about the same (setup, reassignment, finish) hand episode e to warp e % 32 */
    _Pragma("unroll 1") for (int e_fixed = (int)(threadIdx.x >> 5);; e_fixed += WARPS_PER_BLOCK) {
        int e = e_fixed;
        if (FIXED || CC4_STATIC_ASSIGN) { if (e >= n_local) break; }
        else {
            if (lane_id() == 0) e = atomicAdd(&next[ctr], 1);
            e = __shfl_sync(0xFFFFFFFFu, e, 0);
            if (e >= n_local) break;
            if (order) e = order_get(order, e);
        }
        const Cx c = (Cx)(e * (int)sizeof(WarpSmem));
#ifdef CC4_PHASE_CLOCKS
        const long long t0 = clock64();
        f(c, e);
        if (lane_id() == 0 && g_phase_clk && ph < N_PHASES) {
            const unsigned long long dt = (unsigned long long)(clock64() - t0);
            g_phase_clk[(size_t)(first + e) * PC_STRIDE + 2 * ph] += dt;
            if (ph == PH_RED) { atomicMax(&g_phase_clk[(size_t)(first + e) * PC_STRIDE + 17], dt); atomicMax(&g_phase_clk[(size_t)first * PC_STRIDE + 16], dt); }     /* slowest red phase of the episode / of the block's first-episode slot */
        }
#else
        f(c, e);
#endif
    }
    if (SYNC) __syncthreads();
#ifdef CC4_PHASE_CLOCKS
    if (threadIdx.x == 0 && g_phase_clk && n_local > 0 && ph < N_PHASES) g_phase_clk[(size_t)first * PC_STRIDE + 2 * ph + 1] += (unsigned long long)(clock64() - tb0);
#endif
}

/* the same for the lane-per-actor phases: a warp item = `per` consecutive episodes, 32 / per lanes each;
 * f(c, e, exists, sl) runs on all lanes, with the lane's own episode */
template <int per, int parts = 1, class F> __device__ __forceinline__ void each_group(int ph, int first, int n_local, int *next, F f)
{
#ifdef CC4_PHASE_CLOCKS
    const long long tb0 = clock64();
#endif
    const int n_per = (n_local + per - 1) / per, n_items = parts * n_per, sub = lane_id() / (32 / per), sl = lane_id() % (32 / per);
    /* the items of these phases cost about the same and there are at most as many as warps (64 episodes): warp w takes
       item w, w + 32, ... without a trip to the shared counter */
    _Pragma("unroll 1") for (int it_fixed = (int)(threadIdx.x >> 5);; it_fixed += WARPS_PER_BLOCK) {
        int it = it_fixed;
#if !CC4_GROUP_FIXED
        if (lane_id() == 0) it = atomicAdd(&next[ph], 1);
        it = __shfl_sync(0xFFFFFFFFu, it, 0);
#endif
        if (it >= n_items) break;
        const int part = parts > 1 ? it / n_per : 0;       /* `parts` kinds of item over the same episodes (f gets part through sl: sl + 32 * part) */
        const int e = (it - part * n_per) * per + sub;
        const bool exists = e < n_local;
        const Cx c = (Cx)((exists ? e : 0) * (int)sizeof(WarpSmem));
#ifdef CC4_PHASE_CLOCKS
        const long long t0 = clock64();
        f(c, e, exists, sl + 32 * part);
        __syncwarp();
        if (exists && sl == 0 && g_phase_clk) g_phase_clk[(size_t)(first + e) * PC_STRIDE + 2 * ph] += (unsigned long long)(clock64() - t0) / per;
#else
        f(c, e, exists, sl + 32 * part);
        __syncwarp();
#endif
    }
    __syncthreads();
#ifdef CC4_PHASE_CLOCKS
    if (threadIdx.x == 0 && g_phase_clk && n_local > 0) g_phase_clk[(size_t)first * PC_STRIDE + 2 * ph + 1] += (unsigned long long)(clock64() - tb0);
#endif
}

/* -DCC4_PHASE_WALL (profiling build, tools/phase_wall.py): thread 0 of every block reads the clock after each block barrier of
 * the tick and adds the phase's wall cycles to g_wall[phase] with one fire-and-forget atomic; [N_PHASES] = blocks counted,
 * [N_PHASES + 1] = whole block.  Unlike CC4_PHASE_CLOCKS (per-item clocks, global read-modify-write per item) this does not
 * perturb what it measures. */
#ifdef CC4_PHASE_WALL
__device__ unsigned long long g_wall[16];
#define PW_BEGIN() long long pw_t = clock64(); const long long pw_t0 = pw_t
#define PW_MARK(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_wall[i], (unsigned long long)(t_ - pw_t)); pw_t = t_; } } while (0)
#define PW_END() do { if (threadIdx.x == 0) { atomicAdd(&g_wall[N_PHASES], 1ull); atomicAdd(&g_wall[N_PHASES + 1], (unsigned long long)(clock64() - pw_t0)); } } while (0)
#else
#define PW_BEGIN() do { } while (0)
#define PW_MARK(i) do { } while (0)
#define PW_END() do { } while (0)
#endif

/* block prologue: the context every episode of the block shares (callers synchronise before the first use) */
__device__ __forceinline__ void init_block_ctx(const Params &p)
{
    if (threadIdx.x == 0) {
        g_blk.tape = p.tape; g_blk.n_segs = p.seg_off ? p.n_segs : 0;
        g_blk.steps = p.steps; g_blk.rng_mode = p.rng_mode; g_blk.k0 = p.k0; g_blk.red_variant = p.red_variant;
    }
}
__device__ __forceinline__ void init_ctx(Cx c, const Params &p, int env, int tick)
{
    if (lane_id() == 0) {
        uint64_t ep = p.first_episode + (uint64_t)env;
        WarpCtx *g = WC(c);
        g->st = p.stat + env; g->k = p.cold + env;
        g->k1 = p.k1hi ^ (uint32_t)(ep >> 32); g->ep_lo = (uint32_t)ep;
        g->seg = p.seg_off ? p.seg_off + (size_t)env * (p.n_segs + 1) : nullptr;
        g->tick = tick; g->env = env;
    }
    __syncwarp();
}

/* ---- dependency-driven scheduling of the per-episode phases (CC4_FLOW) ----
 * A block barrier between two phases makes every warp wait for the slowest item of the first one (ncu r03i: 27 % of the
 * warp-time of k_step is barrier wait).  What the next phase of an episode really needs is that THIS episode has finished
 * the previous phase.  With CC4_FLOW the phases from the blue actions to the finish are STAGES of one item list per
 * segment: items are handed out from one counter in stage-major order (all blue items, then all green items, ...), so a
 * warp moves on to stage k + 1 only when every item of stage k has been TAKEN, and at most two neighbouring code regions
 * are live.  The items of a stage that follows another one without a barrier are not bound to episodes in advance: item i
 * is "the i-th episode to come out of the stage before" (a READY QUEUE per stage, filled in completion order), so the warps
 * that run out of stage-k items work on the episodes that are already through while the stragglers finish.  (r03a measured
 * a first form in which item i of every stage was the i-th episode of the heaviest-first order: the first items of a stage
 * then wait for exactly the episodes that finish last, nothing overlaps, and every mask was slower than the barriers.)
 * The green stage is a flat list of agents over the episodes in the order they left the blue stage; an episode enters the
 * red stage when its last green agent is done.  No deadlock: an item only waits for completions of the stage before, whose
 * items have all been taken by warps that run them to completion.
 * CC4_FLOW is a bit mask of the stage boundaries WITHOUT a block barrier: 1 blue|green, 2 green|red, 4 red|reassign,
 * 8 reassign|end turn, 16 end turn|finish, 32 setup|policies (a flag per episode, see k_step); 0 = the barrier form. */
#ifndef CC4_FLOW
#define CC4_FLOW 24       /* reassign .. finish without block barriers: -1.5 % (profiles/r04_variants.txt) */
#endif
#ifndef CC4_FLOW_SLEEP
#define CC4_FLOW_SLEEP 0        /* ns between two looks at a ready-queue slot; 0 = spin */
#endif
#ifndef CC4_FLOW_FENCE
#define CC4_FLOW_FENCE 1        /* 0: experiment, no memory fences around the ready flags */
#endif
__device__ __forceinline__ void flow_fence() { if (CC4_FLOW_FENCE) __threadfence_block(); }
__device__ __forceinline__ void flow_pause() { if (CC4_FLOW_SLEEP > 0) __nanosleep(CC4_FLOW_SLEEP); }
enum { FS_BLUE, FS_GREEN, FS_RED, FS_REASSIGN, FS_ENDTURN, FS_FINISH, N_FSTAGES };
struct FlowSmem {
    int gdone[64];              /* green agents of the episode that have had their tick */
    int tail[N_FSTAGES];        /* entries pushed to the stage's ready queue */
    uint32_t gstate;            /* green list: episodes pushed << 16 | their green agents */
    short gend[64];             /* green agents up to and including list position p (positions = order of arrival from the blue stage) */
    uint8_t gq[64];             /* green list position -> episode + 1 (0 = not yet written) */
    uint8_t rq[N_FSTAGES][64];  /* ready queue of a stage: episode + 1 in completion order of the stage before (0 = not yet there) */
};
/* [all lanes, same slot] the episode in a ready-queue slot, waiting for it to arrive */
__device__ __forceinline__ int flow_pop(const volatile uint8_t *slot)
{
    int v = *slot;
    if (v == 0) { do { flow_pause(); v = *slot; } while (v == 0); }
    flow_fence();
    return v - 1;
}
/* [one lane, after the item's last __syncwarp] episode e has left the stage before `st` */
__device__ __forceinline__ void flow_push(FlowSmem *fl, int st, int e)
{
    flow_fence();
    const int k = atomicAdd(&fl->tail[st], 1);
    *(volatile uint8_t *)&fl->rq[st][k] = (uint8_t)(e + 1);
}

/* ------------------------------------------------------------------------------------------------ */
/* kernels                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
/* StepIO: see cc4c (blocks take the episodes from the far end first when `reverse` is set, see cc4_step) */

/* Block b owns episodes [env0 + b*epb, min(env0 + (b+1)*epb, env_end)); env0 / env_end: the slice of the batch this
 * launch covers (the chunked host path launches several slices). */
template <int RV> __global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, CC4_MIN_BLOCKS)
k_step(Params p, int env0, int env_end, int epb, StepIO io)
{
    __shared__ int next[N_PHASES + 6];
    __shared__ uint8_t ng_s[64];
    __shared__ short gend[64];
    __shared__ PhaseOrder ord_blue, ord_red;
    __shared__ __align__(8) uint64_t ld_bar[64];
    const int first = env0 + (int)(io.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * epb, n_local = min(epb, env_end - first);
    PW_BEGIN();
    init_block_ctx(p);
    if (threadIdx.x < N_PHASES + 6) next[threadIdx.x] = 0;
    if (threadIdx.x < NCLS) { ord_blue.cnt[threadIdx.x] = 0; ord_red.cnt[threadIdx.x] = 0; }
    if ((int)threadIdx.x < n_local) mbar_init(&ld_bar[threadIdx.x], 1);
#if CC4_FLOW & 32
    __shared__ uint8_t sdone[64];       /* the episode has had its setup (no block barrier between setup and the policies) */
    if (threadIdx.x < 64) sdone[threadIdx.x] = 0;
#endif
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if ((int)threadIdx.x < n_local) {
        mbar_expect_tx(&ld_bar[threadIdx.x], (uint32_t)sizeof(Cc4Dyn));
        bulk_load(DYN((Cx)(threadIdx.x * (int)sizeof(WarpSmem))), p.dyn + first + threadIdx.x, (uint32_t)sizeof(Cc4Dyn), &ld_bar[threadIdx.x]);
    }
    /* CC4_FLOW bit 5: the policies of an episode only wait for THAT episode's setup (a flag per episode): the record loads of
       a block's last episodes overlap the policies of its first ones.  The setup map is fixed (warp e % 32) and a warp
       finishes its setups before it looks at a policy item, so nothing a policy item waits for can be waiting itself. */
    each_episode<CC4_FIXED_PHASES & 1, !(CC4_FLOW & 32)>(PH_SETUP, first, n_local, next, [&](Cx c, int e) {
        init_ctx(c, p, first + e, 0);
        mbar_wait(&ld_bar[e], 0);
        ph_setup(c);
        if (lane_id() == 0) ng_s[e] = DYN(c)->ngreen;
#if CC4_FLOW & 32
        if (lane_id() == 0) { flow_fence(); *(volatile uint8_t *)&sdone[e] = 1; }
#endif
    });
    PW_MARK(PH_SETUP);
    /* CC4_MERGE (tuning knob): which neighbouring per-episode phases run back to back inside one work item instead of
       being separated by a block barrier: bit 0 red + reassign, bit 1 FSM digest + finish, bit 2 red ... finish all in
       one item.  Every merge measures slower (profiles/r01x_variants.txt: 0.477 ms -> 0.501 / 0.487 / 0.576 ms): fewer
       barriers mean better warp utilisation but two code regions live at once. */
    auto finish = [&](Cx c, int e) {
        const int env = first + e;
        ph_finish(c, io.msgs ? io.msgs + (size_t)env * 5 : nullptr, (io.flags & CC4_STEP_REWARD) && io.reward ? io.reward + env : nullptr, false,
                  io.done ? io.done + env : nullptr);
        if ((io.flags & CC4_STEP_OBS) && io.obs) w_write_obs(DYN(c), SCR(c), io.obs, env, (io.flags & CC4_STEP_OBS_COMPACT) != 0, (io.flags & CC4_STEP_OBS_BITS) != 0);
        __syncwarp();
        if (lane_id() == 0) bulk_store(p.dyn + env, DYN(c), (uint32_t)sizeof(Cc4Dyn));
    };
    each_group<4, 2>(PH_POLICY, first, n_local, next, [&](Cx c, int e, bool exists, int sl) {
#if CC4_FLOW & 32
        if (exists) { const volatile uint8_t *d = &sdone[e]; while (*d == 0) flow_pause(); }
        __syncwarp();
        flow_fence();
#endif
        ph_policy<RV>(c, e, exists, exists, sl & 31, sl >= 32, io.actions + (size_t)(first + (exists ? e : 0)) * 5, 0, 0, 0, &ord_blue, &ord_red);
    });
    PW_MARK(PH_POLICY);
    if (threadIdx.x < 32) green_prefix(n_local, ng_s, gend);        /* read after the blue phase's barrier */
#if CC4_FLOW & 31
    {
        __shared__ FlowSmem fl;
        if (threadIdx.x < 64) { fl.gdone[threadIdx.x] = 0; fl.gq[threadIdx.x] = 0; }
        if (threadIdx.x < N_FSTAGES * 64) (&fl.rq[0][0])[threadIdx.x] = 0;
        if (threadIdx.x < N_FSTAGES) fl.tail[threadIdx.x] = 0;
        if (threadIdx.x == 0) fl.gstate = 0;
        __syncthreads();
        const int lane = lane_id();
        const int total_green = n_local > 0 ? gend[n_local - 1] : 0;     /* green_prefix ran a barrier ago */
        /* run stages [lo, hi] of the item list from one counter, then a block barrier */
        auto segment = [&](const int lo, const int hi) {
            const int cnt[N_FSTAGES] = { n_local, (total_green + 31) >> 5, n_local, n_local, (n_local + CC4_ENDTURN_PER - 1) / CC4_ENDTURN_PER, n_local };
            _Pragma("unroll 1") for (;;) {
                int i = 0;
                if (lane == 0) i = atomicAdd(&next[N_PHASES + lo], 1);
                i = __shfl_sync(0xFFFFFFFFu, i, 0);
                int st = lo;
                while (st <= hi && i >= cnt[st]) { i -= cnt[st]; st++; }
                if (st > hi) break;
                const bool head = st == lo;     /* first stage of the segment: every item is ready, in the static order */
                const bool more = st < hi;      /* a stage follows without a barrier: announce the completions */
                if (st == FS_BLUE) {
                    const int e = order_get(&ord_blue, i);
                    ph_blue((Cx)(e * (int)sizeof(WarpSmem)));
                    if (more && lane == 0) {    /* append the episode's agents to the green list */
                        flow_fence();
                        const uint32_t ng = ng_s[e], old = atomicAdd(&fl.gstate, (1u << 16) | ng);
                        const int pos = (int)(old >> 16);
                        fl.gend[pos] = (short)((old & 0xFFFFu) + ng);
                        flow_fence();
                        *(volatile uint8_t *)&fl.gq[pos] = (uint8_t)(e + 1);
                    }
                } else if (st == FS_GREEN) {
                    /* 32 green agents of consecutive list positions (an episode has >= 24 agents: at most three episodes) */
                    const int gidx = i * 32 + lane;
                    const short *ge = head ? gend : fl.gend;
                    int npos = n_local;
                    if (!head) {    /* the list must have grown past this item, and the positions it covers must be written */
                        const uint32_t need = (uint32_t)min(32 * (i + 1), total_green);
                        uint32_t gs = 0;
                        if (lane == 0) { const volatile uint32_t *g = &fl.gstate; while (((gs = *g) & 0xFFFFu) < need) flow_pause(); }
                        npos = (int)(__shfl_sync(0xFFFFFFFFu, gs, 0) >> 16);
                        if (lane < npos) { const volatile uint8_t *q = &fl.gq[lane]; while (*q == 0) { } }
                        if (32 + lane < npos) { const volatile uint8_t *q = &fl.gq[32 + lane]; while (*q == 0) { } }
                        __syncwarp();
                        flow_fence();
                    }
                    const uint32_t past_a = __ballot_sync(0xFFFFFFFFu, lane < npos && ge[lane] > i * 32);
                    const uint32_t past_b = __ballot_sync(0xFFFFFFFFu, 32 + lane < npos && ge[32 + lane] > i * 32);
                    const int p_first = past_a ? __ffs(past_a) - 1 : 31 + __ffs(past_b);
                    int pos = -1;
                    if (gidx < total_green) { pos = p_first; while (ge[pos] <= gidx) pos++; }
                    const int e = pos < 0 ? -1 : head ? pos : (int)fl.gq[pos] - 1;
                    if (e >= 0) {
                        const Cx c = (Cx)(e * (int)sizeof(WarpSmem));
                        const int g = gidx - (pos ? ge[pos - 1] : 0);
                        Cc4Dyn *sd = DYN(c); Scratch *x = SCR(c);
                        const int gr = green_tick(c, g, reinterpret_cast<uint32_t *>(sd->ev));
                        if (gr & 0xFF) atomicSub(&x->reward, gr & 0xFF);
                        if (gr & GREEN_PHISH) atomicOr(&x->g.phish[g >> 5], 1u << (g & 31));
                    }
                    __syncwarp();
                    if (more) {     /* the item that completes an episode's agents hands the episode to the red stage */
                        const uint32_t same = __match_any_sync(0xFFFFFFFFu, e);
                        if (e >= 0 && (int)__ffs(same) - 1 == lane) {
                            flow_fence();
                            const int k = __popc(same);
                            if (atomicAdd(&fl.gdone[e], k) + k == (int)ng_s[e]) flow_push(&fl, FS_RED, e);
                        }
                    }
                } else if (st == FS_RED) {
                    const int e = head ? order_get(&ord_red, i) : flow_pop(&fl.rq[FS_RED][i]);
                    ph_red<RV>((Cx)(e * (int)sizeof(WarpSmem)));
                    if (more && lane == 0) flow_push(&fl, FS_REASSIGN, e);
                } else if (st == FS_REASSIGN) {
                    const int e = head ? i : flow_pop(&fl.rq[FS_REASSIGN][i]);
                    ph_reassign<RV>((Cx)(e * (int)sizeof(WarpSmem)));
                    if (more && lane == 0) flow_push(&fl, FS_ENDTURN, e);
                } else if (st == FS_ENDTURN) {
                    constexpr int per = CC4_ENDTURN_PER;
                    const int sub = lane / (32 / per), sl = lane % (32 / per), k = i * per + sub;
                    const bool exists = k < n_local;
                    int e = exists ? k : 0;
                    if (!head && exists) e = flow_pop(&fl.rq[FS_ENDTURN][k]);
                    __syncwarp();
                    ph_endturn<RV>((Cx)(e * (int)sizeof(WarpSmem)), exists, sl);
                    __syncwarp();
                    if (more && exists && sl == 0) flow_push(&fl, FS_FINISH, e);
                } else {
                    const int e = head ? i : flow_pop(&fl.rq[FS_FINISH][i]);
                    finish((Cx)(e * (int)sizeof(WarpSmem)), e);
                }
            }
            __syncthreads();
        };
        /* the boundaries whose bit is clear keep their block barrier: cut the stage list there */
        /* a stage with a barrier on both sides runs in its barrier form (fixed item -> warp maps where the items cost the same) */
        auto alone = [&](const int b) {
            if (b == FS_BLUE) each_episode(PH_BLUE, first, n_local, next, [&](Cx c, int) { ph_blue(c); }, &ord_blue);
            else if (b == FS_GREEN) green_items(PH_GREEN, first, n_local, next, ng_s, gend);
            else if (b == FS_RED) each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c); }, &ord_red);
            else if (b == FS_REASSIGN) each_episode<(CC4_FIXED_PHASES >> 1) & 1>(PH_REASSIGN, first, n_local, next, [&](Cx c, int) { ph_reassign<RV>(c); });
            else if (b == FS_ENDTURN) each_group<CC4_ENDTURN_PER>(PH_ENDTURN, first, n_local, next, [&](Cx c, int, bool exists, int sl) { ph_endturn<RV>(c, exists, sl); });
            else each_episode<(CC4_FIXED_PHASES >> 2) & 1>(PH_FINISH, first, n_local, next, finish);
        };
        constexpr int M = CC4_FLOW & 31;
        int lo = 0;
#pragma unroll
        for (int b = 0; b < N_FSTAGES; b++) {
            if (b == N_FSTAGES - 1 || !(M >> b & 1)) { if (lo == b) alone(b); else segment(lo, b); PW_MARK(PH_BLUE + b); lo = b + 1; }
        }
    }
#else
#if CC4_SPLIT & 1
    each_episode(N_PHASES + 1, first, n_local, next, [&](Cx c, int) { ph_blue(c, false); });
    each_episode(PH_BLUE, first, n_local, next, [&](Cx c, int) { ph_blue_exec(c); }, &ord_blue);
#else
#if CC4_BLUE_PARTS == 2
    each_episode(PH_BLUE, first, n_local, next, [&](Cx c, int) { ph_blue(c, true, 0x07u); }, &ord_blue);
    each_episode(PH_BLUE, first, n_local, next, [&](Cx c, int) { ph_blue_exec(c, 0x18u, false); }, &ord_blue, N_PHASES + 3);
#else
    each_episode(PH_BLUE, first, n_local, next, [&](Cx c, int) { ph_blue(c); }, &ord_blue);
#endif
#endif
    PW_MARK(PH_BLUE);
    green_items(PH_GREEN, first, n_local, next, ng_s, gend);
    PW_MARK(PH_GREEN);
#if CC4_SPLIT & 2
    each_episode(N_PHASES + 2, first, n_local, next, [&](Cx c, int) { ph_phishing(c); });
#endif
#if CC4_MERGE & 4
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int e) { ph_red<RV>(c); ph_reassign<RV>(c); ph_endturn<RV>(c, true, lane_id()); __syncwarp(); finish(c, e); }, &ord_red);
#else
#if CC4_MERGE & 1
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c); ph_reassign<RV>(c); }, &ord_red);
#else
#if CC4_RED_PARTS == 2
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c, !(CC4_SPLIT & 2), 0x07u); }, &ord_red);
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c, false, 0x38u); }, &ord_red, N_PHASES + 4);
#elif CC4_RED_PARTS == 3
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c, !(CC4_SPLIT & 2), 0x03u); }, &ord_red);
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c, false, 0x0Cu); }, &ord_red, N_PHASES + 4);
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c, false, 0x30u); }, &ord_red, N_PHASES + 5);
#else
    each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { ph_red<RV>(c, !(CC4_SPLIT & 2)); }, &ord_red);
#endif
    PW_MARK(PH_RED);
#if CC4_MERGE & 8      /* reassignment + FSM digest + finish of an episode back to back in one item (two barriers fewer) */
    each_episode<true>(PH_FINISH, first, n_local, next, [&](Cx c, int e) { ph_reassign<RV>(c); ph_endturn<RV>(c, true, lane_id()); __syncwarp(); finish(c, e); });
    PW_MARK(PH_FINISH);
#else
    each_episode<(CC4_FIXED_PHASES >> 1) & 1>(PH_REASSIGN, first, n_local, next, [&](Cx c, int) { ph_reassign<RV>(c); });
    PW_MARK(PH_REASSIGN);
#endif
#endif
#if CC4_MERGE & 8
#elif CC4_MERGE & 2
    each_episode(PH_FINISH, first, n_local, next, [&](Cx c, int e) { ph_endturn<RV>(c, true, lane_id()); __syncwarp(); finish(c, e); });
#else
    each_group<CC4_ENDTURN_PER>(PH_ENDTURN, first, n_local, next, [&](Cx c, int, bool exists, int sl) { ph_endturn<RV>(c, exists, sl); });
    PW_MARK(PH_ENDTURN);
    each_episode<(CC4_FIXED_PHASES >> 2) & 1>(PH_FINISH, first, n_local, next, finish);
    PW_MARK(PH_FINISH);
#endif
#endif
#endif  /* CC4_FLOW */
    if (lane_id() == 0) bulk_store_wait();      /* shared memory must outlive the copies this lane started */
    PW_END();
}

/* n_ticks fused ticks with the benchmark's synthetic blue policy (see ph_policy); the state stays in shared memory */
template <int RV> __global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, CC4_MIN_BLOCKS)
k_rollout(Params p, int epb, int n_ticks, uint32_t pk0, uint32_t pk1, float *reward_sum, int flags)
{
    __shared__ int next[N_PHASES + 1];
    __shared__ uint8_t ng_s[64];
    __shared__ short gend[64];
    __shared__ PhaseOrder ord_blue, ord_red;
    __shared__ int alive;
    const int first = blockIdx.x * epb, n_local = min(epb, p.n_envs - first);
    init_block_ctx(p);
    __syncthreads();
    _Pragma("unroll 1") for (int e = threadIdx.x >> 5; e < n_local; e += WARPS_PER_BLOCK) { Cx c = (Cx)(e * (int)sizeof(WarpSmem)); init_ctx(c, p, first + e, 0); w_load(DYN(c), p.dyn + first + e); }
    _Pragma("unroll 1") for (int t = 0; t < n_ticks; t++) {
        __syncthreads();
        if (threadIdx.x <= N_PHASES) next[threadIdx.x] = 0;
        if (threadIdx.x < NCLS) { ord_blue.cnt[threadIdx.x] = 0; ord_red.cnt[threadIdx.x] = 0; }
        if (threadIdx.x == 0) alive = 0;
        __syncthreads();
        /* an episode that is over keeps its state; the block stops when all of its episodes are over */
        each_episode(PH_SETUP, first, n_local, next, [&](Cx c, int e) {
            if (lane_id() == 0) ng_s[e] = DYN(c)->done ? 0 : DYN(c)->ngreen;      /* an episode that is over has no green work */
            if (!DYN(c)->done) { if (lane_id() == 0) atomicAdd(&alive, 1); ph_setup(c); }
        });
        if (alive == 0) break;
        each_group<4, 2>(PH_POLICY, first, n_local, next, [&](Cx c, int e, bool exists, int sl) {
            ph_policy<RV, true>(c, e, exists, exists && !DYN(c)->done, sl & 31, sl >= 32, nullptr, pk0, pk1, p.first_episode + (uint64_t)(first + e), &ord_blue, &ord_red);
        });
        if (threadIdx.x < 32) green_prefix(n_local, ng_s, gend);
        each_episode(PH_BLUE, first, n_local, next, [&](Cx c, int) { if (!DYN(c)->done) ph_blue(c); }, &ord_blue);
        green_items(PH_GREEN, first, n_local, next, ng_s, gend);
        each_episode(PH_RED, first, n_local, next, [&](Cx c, int) { if (!DYN(c)->done) ph_red<RV>(c); }, &ord_red);
        each_episode(PH_REASSIGN, first, n_local, next, [&](Cx c, int) { if (!DYN(c)->done) ph_reassign<RV>(c); });
        each_group<CC4_ENDTURN_PER>(PH_ENDTURN, first, n_local, next, [&](Cx c, int, bool exists, int sl) { ph_endturn<RV>(c, exists && !DYN(c)->done, sl); });
        /* `done` flips in this phase: the ticks that follow skip the episode, this tick must not look at the flag again */
        each_episode(PH_FINISH, first, n_local, next, [&](Cx c, int e) {
            if (DYN(c)->done) return;
            ph_finish(c, nullptr, (flags & CC4_STEP_REWARD) && reward_sum ? reward_sum + first + e : nullptr, true, nullptr);
        });
    }
    __syncthreads();
    _Pragma("unroll 1") for (int e = threadIdx.x >> 5; e < n_local; e += WARPS_PER_BLOCK) w_store(p.dyn + first + e, DYN((Cx)(e * (int)sizeof(WarpSmem))));
}

#if !CC4_PROD_ONLY     /* the reset kernels exist once, in the full namespace */
/* ---- reset ---- */
struct ResetTmp {
    uint32_t procs[128][8];
    uint8_t np[128];
    uint32_t used[288];          /* bitmap over pids 1000..9999, EnterpriseScenarioGenerator._generate_pid :563-578 */
    uint16_t router_pid[16];
};
struct ResetSmem {
    WarpSmem w;
    ResetTmp t;
};

__device__ int gen_pid(Cx c, ResetTmp *t)
{
    _Pragma("unroll 1") for (int guard = 0; guard < 64; guard++) {
        int v = (int)rng_int_reset(c, ST_R_PID, 9000);
        if (DYN(c)->err & CC4_E_TAPE_END) return 1000 + v;
        if (!(t->used[v >> 5] >> (v & 31) & 1)) { t->used[v >> 5] |= 1u << (v & 31); return 1000 + v; }
    }
    fail(c, CC4_E_CAPACITY);
    return 1000;
}
__device__ void gen_host(Cx c, ResetTmp *t, int slot)
{   /* _generate_linux_host_services :530-561 + _generate_linux_host_processes :580-629 */
    int sn = slot >> 4, n = 0;
    uint64_t v = SVC_EMPTY;
    uint32_t *pr = t->procs[slot];      /* entries go straight to the host's list in shared memory; the option lists are packed in registers */
    auto add = [&](uint32_t kind, uint32_t pid) { v = svc_set(v, n, kind | 5u << 4); pr[n++] = P_MAKE(pid, kind, 0, 1); };
    add(K_SSHD, (uint32_t)gen_pid(c, t));
    if (sn == S_OZA || sn == S_OZB) add(K_OT, (uint32_t)gen_pid(c, t));
    uint32_t okind = K_APACHE2 | K_MYSQLD << 4 | K_SMTP << 8;      /* remaining optional services, four bits each ... */
    uint64_t opid = 0;                                              /* ... and their pids (drawn up front, :588-590), sixteen bits each */
    _Pragma("unroll 1") for (int i = 0; i < 3; i++) opid |= (uint64_t)(uint32_t)gen_pid(c, t) << (16 * i);
    int nopt = 3;
    int nadd = (int)rng_int_reset(c, ST_R_NADDON, 4);
    _Pragma("unroll 1") for (int i = 0; i < nadd; i++) {
        const int k = (int)rng_int_reset(c, ST_R_ADDON, (uint32_t)nopt);
        add(okind >> (4 * k) & 15u, (uint32_t)(opid >> (16 * k)) & 0xFFFFu);
        okind = (okind & ((1u << (4 * k)) - 1u)) | (okind >> (4 * (k + 1))) << (4 * k);         /* list.pop(k) */
        opid = (opid & ((1ull << (16 * k)) - 1ull)) | (opid >> (16 * (k + 1))) << (16 * k);
        nopt--;
    }
    t->np[slot] = (uint8_t)n;
    WC(c)->k->svc64[slot] = v;
}
__device__ int tmp_create_pid(Cx c, ResetTmp *t, int slot)
{
    uint32_t mx = 0;
    _Pragma("unroll 1") for (int i = 0; i < t->np[slot]; i++) if (P_PID(t->procs[slot][i]) > mx) mx = P_PID(t->procs[slot][i]);
    return (int)(mx + 1 + rng_int_reset(c, ST_CREATE_PID, 9));
}
__device__ void tmp_add_session_proc(Cx c, ResetTmp *t, int slot, int kind)
{
    if (slot >= CC4_SLOT_ROUTER) { t->router_pid[slot - CC4_SLOT_ROUTER] = (uint16_t)(1 + rng_int_reset(c, ST_CREATE_PID, 9)); return; }
    int pid = tmp_create_pid(c, t, slot);
    t->procs[slot][t->np[slot]++] = P_MAKE(pid, kind, 0, 0);
}

__global__ void __launch_bounds__(32) k_reset(Params p, uint8_t *obs, uint8_t *mask)
{
    ResetSmem *rs = reinterpret_cast<ResetSmem *>(smem_raw);
    WarpSmem *w = &rs->w;
    int env = blockIdx.x;
    const bool valid = env < p.n_envs;
    uint32_t *z = reinterpret_cast<uint32_t *>(rs);
    _Pragma("unroll 1") for (int i = lane_id(); i < (int)(sizeof(ResetSmem) / 4); i += 32) z[i] = 0;
    __syncwarp();
    Cx c = 0;
    init_block_ctx(p);
    init_ctx(c, p, env, -1);
    if (lane_id() <= CC4_NRED) SCR(c)->blk_id[lane_id()] = 0xFFFFFFFFu;
    __syncwarp();
    _Pragma("unroll 1") for (int h = lane_id(); h < 128; h += 32) WC(c)->k->svc64[h] = SVC_EMPTY;
    __syncwarp();
    /* scenario generation is ONE serial random stream per episode (the reference's draw order): lane 0 */
    if (lane_id() == 0) {
        Cc4Dyn *s = DYN(c); ResetTmp *t = &rs->t; Cc4Stat *st = WC(c)->st;
        _Pragma("unroll 1") for (int a = 0; a < 16; a++) s->inprog[a] = IP_NONE;
        /* hosts, EnterpriseScenarioGenerator._generate_hosts :312-371 */
        _Pragma("unroll 1") for (int sn = 0; sn < 8; sn++) {
            s->nuser[sn] = (uint8_t)(3 + rng_int_reset(c, ST_R_NUSER, 8));
            _Pragma("unroll 1") for (int i = 0; i < s->nuser[sn]; i++) gen_host(c, t, sn * 16 + 6 + i);
            s->nserver[sn] = (uint8_t)(1 + rng_int_reset(c, ST_R_NSERVER, 6));
            _Pragma("unroll 1") for (int i = 0; i < s->nserver[sn]; i++) gen_host(c, t, sn * 16 + i);
        }
        /* blue agents :631-698 */
        _Pragma("unroll 1") for (int b = 0; b < CC4_NBLUE; b++) {
            int n = 0;
            _Pragma("unroll 1") for (int i = 0; i < (b == 4 ? 3 : 1); i++) { int sn = b < 4 ? b : S_PAZ + i; n += 1 + s->nuser[sn] + s->nserver[sn]; }
            int pk = (int)rng_int_reset(c, ST_R_BLUE_PARENT, (uint32_t)n), parent = -1;
            _Pragma("unroll 1") for (int i = 0, k = 0; i < (b == 4 ? 3 : 1) && parent < 0; i++) {
                int sn = b < 4 ? b : S_PAZ + i, nu = s->nuser[sn], m = 1 + nu + s->nserver[sn];
                if (pk < k + m) { int j = pk - k; parent = j == 0 ? CC4_SLOT_ROUTER + sn : j <= nu ? sn * 16 + 6 + (j - 1) : sn * 16 + (j - 1 - nu); }
                k += m;
            }
            st->blue_parent[b] = (uint8_t)parent;
        }
        /* green agents :700-749 */
        int ng = 0;
        { uint64_t cum = 0; _Pragma("unroll 1") for (int sn = 0; sn < 8; sn++) { ng += s->nuser[sn]; cum |= (uint64_t)ng << (8 * sn); } s->green_cum = cum; }
        s->ngreen = (uint8_t)ng;
        /* red agents :751-812 */
        _Pragma("unroll 1") for (int r = 0; r < CC4_NRED; r++) {
            int ns = r == 5 ? 3 : 1;
            int si = (int)rng_int_reset(c, ST_R_RED_SUBNET, (uint32_t)ns);
            int sn = r == 0 ? S_CON : r <= 4 ? r - 1 : S_PAZ + si;
            int k = (int)rng_int_reset(c, ST_R_RED_HOST, (uint32_t)(s->nuser[sn] + s->nserver[sn]));
            st->red_start[r] = (uint8_t)(k < s->nuser[sn] ? sn * 16 + 6 + k : sn * 16 + (k - s->nuser[sn]));
        }
        /* State.__init__ State.py:108-134: session processes (Host.add_session Host.py:189-196) */
        _Pragma("unroll 1") for (int b = 0; b < CC4_NBLUE; b++) {
            tmp_add_session_proc(c, t, st->blue_parent[b], K_SESSION);
            _Pragma("unroll 1") for (int i = 0; i < (b == 4 ? 3 : 1); i++) {
                int sn = b < 4 ? b : S_PAZ + i;
                if (st->blue_parent[b] != CC4_SLOT_ROUTER + sn) tmp_add_session_proc(c, t, CC4_SLOT_ROUTER + sn, K_SESSION);
                _Pragma("unroll 1") for (int k = 0; k < s->nuser[sn]; k++) if (st->blue_parent[b] != sn * 16 + 6 + k) tmp_add_session_proc(c, t, sn * 16 + 6 + k, K_SESSION);
                _Pragma("unroll 1") for (int k = 0; k < s->nserver[sn]; k++) if (st->blue_parent[b] != sn * 16 + k) tmp_add_session_proc(c, t, sn * 16 + k, K_SESSION);
            }
        }
        _Pragma("unroll 1") for (int g = 0; g < ng; g++) tmp_add_session_proc(c, t, green_host_of(s, g), K_SESSION);
        {
            int h = st->red_start[0], pid = tmp_create_pid(c, t, h);
            t->procs[h][t->np[h]++] = P_MAKE(pid, K_REDABS, 0, 0);
            s->sess[0] = S_MAKE(pid, 0, h); s->sessf[0] = (uint8_t)(0 | 1 << 3); s->n_sess = 1; s->sess_count[0] = 1; s->red_active = 1;
        }
        WC(c)->k->arena_top = 128 * CC4_OSLAB;
        _Pragma("unroll 1") for (int i = 0; i < 16; i++) st->router_pid[i] = t->router_pid[i];
        /* initial knowledge: AgentInterface.set_init_obs AgentInterface.py:110-117 with the OSINT of :784-785 */
        _Pragma("unroll 1") for (int r = 0; r < CC4_NRED; r++) { int h = st->red_start[r]; set128(s->ip_known[r], h); set128(s->hn_known[r], h); s->sn_known[r] |= (uint16_t)(1u << (h >> 4)); }
        s->ss[0][0] = 0; s->n_ss[0] = 1;
        /* red_agent_0 digests its initial observation (its start host becomes 'U', FiniteStateRedAgent.py:221-224) */
        { int h0 = st->red_start[0]; s->fsm_order[0][0] = (uint8_t)h0; s->n_fsm[0] = 1; s->fsm_started = 1;
          set128(s->fsm[0].present, h0); set128(s->fsm[0].hn, h0); set128(s->fsm[0].pl[2], h0); }
        if (CC4_TAPE_ON() && g_blk.n_segs > 0 && ctr_get(SCR(c), NAG) != WC(c)->seg[1] - WC(c)->seg[0]) fail(c, CC4_E_TAPE_LEFT);
    }
    __syncwarp();
    {   /* process lists + Host.create_backup Host.py:316-371: the generated lists leave shared memory, a lane per host */
        ResetTmp *t = &rs->t; Cc4Stat *st = WC(c)->st; Cc4Cold *k = WC(c)->k;
        _Pragma("unroll 1") for (int h = lane_id(); h < 128; h += 32) {
            const int np = t->np[h];
            st->o_nproc[h] = (uint8_t)np;
            k->pdir[h] = PD_MAKE(h * CC4_OSLAB, np, 3);
            uint32_t mx = 0;
            _Pragma("unroll 1") for (int i = 0; i < np; i++) { uint32_t pw = t->procs[h][i]; k->parena[h * CC4_OSLAB + i] = pw; st->o_procs[h][i] = pw; if (P_PID(pw) > mx) mx = P_PID(pw); }
            k->pmax[h] = (uint16_t)mx;
            st->o_svc64[h] = k->svc64[h];
        }
    }
    __syncwarp();
    if (valid && obs) w_write_obs(&w->dyn, &w->x, obs, env);
    if (valid && mask) w_write_mask(&w->dyn, mask + (size_t)env * CC4_N_BLUE * CC4_MASK_LEN);
    w_store(p.dyn + env, &w->dyn);
}


/* ---- reset, production mode: one THREAD per episode ----
 * Scenario generation is one serial random stream per episode (the reference's draw order, ~1 000 draws with data-dependent
 * counts), so an episode cannot use more than one lane; k_reset (above, kept for tape replay) therefore runs one warp per
 * episode with 31 idle lanes, 25 episodes in flight per SM.  In production mode nothing of the generation needs the
 * per-episode shared-memory context: the Philox stream lives in registers, the generated lists go straight to their
 * place in the cold / static records, and what the chain itself reads back (the bitmap of used pids, the largest pid and the
 * process count of every host) is 1.5 KB of shared memory per lane.  So a block of 128 threads generates 128 episodes side
 * by side (lanes diverge only in loop trip counts), the hot records are zero-filled by the whole block with coalesced
 * stores first, and the observation rows and masks are written afterwards by k_reset_outputs (warp per episode, coalesced).
 * Same words, same order, same results as k_reset (tests: production-mode parity with the oracle and the Philox-seam
 * fixtures of the live reference compare the state after reset). */
struct ResetLane {
    uint32_t used[288];          /* bitmap over pids 1000..9999 (EnterpriseScenarioGenerator._generate_pid :563-578) */
    uint16_t pmax[128];          /* largest pid on the host so far (Host.create_pid Host.py:198-200) */
    uint8_t np[128];             /* processes on the host so far */
    uint8_t router_pid[16];      /* 1..9: a router's only process is the blue session's (Host.create_pid on an empty list) */
    uint32_t pad;                /* stride = 389 words (odd): lanes hit different banks at equal offsets */
};
static_assert(sizeof(ResetLane) == 389 * 4, "ResetLane stride");
/* 148 lanes x 1 556 B = 225 KB of shared memory: one block per SM, and 65 536 episodes are 443 blocks = three full waves of 148
   (128 threads per block were 512 blocks = 3.5 waves, i.e. four: 1.61 ms against 1.2 ms) */
constexpr int RESET_FAST_THREADS = 148;

__global__ void __launch_bounds__(RESET_FAST_THREADS) k_reset_fast(Params p)
{
    ResetLane *t = reinterpret_cast<ResetLane *>(smem_raw) + threadIdx.x;
    const int first = (int)blockIdx.x * RESET_FAST_THREADS, env = first + (int)threadIdx.x;
    {   /* the block's hot records start from zero: coalesced 128-bit stores */
        uint4 *d4 = reinterpret_cast<uint4 *>(p.dyn + first);
        const int n4 = min(RESET_FAST_THREADS, p.n_envs - first) * DYN_U4;
        _Pragma("unroll 4") for (int i = (int)threadIdx.x; i < n4; i += RESET_FAST_THREADS) d4[i] = make_uint4(0, 0, 0, 0);
        uint32_t *z = reinterpret_cast<uint32_t *>(t);
        _Pragma("unroll 1") for (int i = 0; i < (int)(sizeof(ResetLane) / 4); i++) z[i] = 0;
    }
    __syncthreads();
    if (env >= p.n_envs) return;
    Cc4Dyn *s = p.dyn + env; Cc4Stat *st = p.stat + env; Cc4Cold *k = p.cold + env;
    /* the reset stream of the episode (cc4_device.cuh philox_block with actor NAG, tick -1): registers only */
    const uint64_t ep = p.first_episode + (uint64_t)env;
    const uint32_t k1 = p.k1hi ^ (uint32_t)(ep >> 32), ep_lo = (uint32_t)ep;
    uint32_t j = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    auto ri = [&](uint32_t n) -> uint32_t {
        if ((j & 3u) == 0u) { uint32_t o[4] = { j >> 2, (uint32_t)NAG, 0u, ep_lo }; philox4x32_10(o, p.k0, k1); b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3]; }
        const uint32_t w = (j & 3u) == 0u ? b0 : (j & 3u) == 1u ? b1 : (j & 3u) == 2u ? b2 : b3;
        j++;
        return __umulhi(w, n);
    };
    uint32_t err = 0;
    auto gen_pid = [&]() -> uint32_t {
        _Pragma("unroll 1") for (int guard = 0; guard < 64; guard++) {
            const uint32_t v = ri(9000);
            if (!(t->used[v >> 5] >> (v & 31) & 1u)) { t->used[v >> 5] |= 1u << (v & 31); return 1000u + v; }
        }
        err |= CC4_E_CAPACITY;
        return 1000u;
    };
    uint64_t nu64 = 0, ns64 = 0;        /* users / servers per subnet, one byte each (registers, no local arrays) */
    #define NUSER(sn_) ((int)(nu64 >> (8 * (sn_)) & 0xFFu))
    #define NSERVER(sn_) ((int)(ns64 >> (8 * (sn_)) & 0xFFu))
    auto gen_host = [&](int slot) {     /* _generate_linux_host_services :530-561 + _generate_linux_host_processes :580-629 */
        const int sn = slot >> 4;
        int n = 0;
        uint64_t v = SVC_EMPTY;
        uint32_t mx = 0;
        auto add = [&](uint32_t kind, uint32_t pid) {
            v = svc_set(v, n, kind | 5u << 4);
            const uint32_t pw = P_MAKE(pid, kind, 0, 1);
            k->parena[slot * CC4_OSLAB + n] = pw; st->o_procs[slot][n] = pw;
            if (pid > mx) mx = pid;
            n++;
        };
        add(K_SSHD, gen_pid());
        if (sn == S_OZA || sn == S_OZB) add(K_OT, gen_pid());
        uint32_t okind = K_APACHE2 | K_MYSQLD << 4 | K_SMTP << 8;
        uint64_t opid = 0;
        _Pragma("unroll 1") for (int i = 0; i < 3; i++) opid |= (uint64_t)gen_pid() << (16 * i);
        int nopt = 3;
        const int nadd = (int)ri(4);
        _Pragma("unroll 1") for (int i = 0; i < nadd; i++) {
            const int q = (int)ri((uint32_t)nopt);
            add(okind >> (4 * q) & 15u, (uint32_t)(opid >> (16 * q)) & 0xFFFFu);
            okind = (okind & ((1u << (4 * q)) - 1u)) | (okind >> (4 * (q + 1))) << (4 * q);         /* list.pop(q) */
            opid = (opid & ((1ull << (16 * q)) - 1ull)) | (opid >> (16 * (q + 1))) << (16 * q);
            nopt--;
        }
        t->np[slot] = (uint8_t)n; t->pmax[slot] = (uint16_t)mx;
        k->svc64[slot] = v; st->o_svc64[slot] = v;
    };
    /* hosts, EnterpriseScenarioGenerator._generate_hosts :312-371 */
    _Pragma("unroll 1") for (int sn = 0; sn < 8; sn++) {
        nu64 |= (uint64_t)(3 + ri(8)) << (8 * sn);
        _Pragma("unroll 1") for (int i = 0; i < NUSER(sn); i++) gen_host(sn * 16 + 6 + i);
        ns64 |= (uint64_t)(1 + ri(6)) << (8 * sn);
        _Pragma("unroll 1") for (int i = 0; i < NSERVER(sn); i++) gen_host(sn * 16 + i);
        _Pragma("unroll 1") for (int q = 0; q < 16; q++) {       /* the slots without a host */
            const bool exists = q < 6 ? q < NSERVER(sn) : (q - 6) < NUSER(sn);
            if (!exists) { k->svc64[sn * 16 + q] = SVC_EMPTY; st->o_svc64[sn * 16 + q] = SVC_EMPTY; }
        }
    }
    /* blue agents :631-698 */
    uint64_t bp64 = 0, rs64 = 0;        /* blue parent hosts / red start hosts, one byte each */
    #define BLUE_PARENT(b_) ((int)(bp64 >> (8 * (b_)) & 0xFFu))
    #define RED_START(r_) ((int)(rs64 >> (8 * (r_)) & 0xFFu))
    _Pragma("unroll 1") for (int b = 0; b < CC4_NBLUE; b++) {
        int n = 0;
        _Pragma("unroll 1") for (int i = 0; i < (b == 4 ? 3 : 1); i++) { const int sn = b < 4 ? b : S_PAZ + i; n += 1 + NUSER(sn) + NSERVER(sn); }
        const int pk = (int)ri((uint32_t)n);
        int parent = -1;
        _Pragma("unroll 1") for (int i = 0, q = 0; i < (b == 4 ? 3 : 1) && parent < 0; i++) {
            const int sn = b < 4 ? b : S_PAZ + i, nu = NUSER(sn), m = 1 + nu + NSERVER(sn);
            if (pk < q + m) { const int jj = pk - q; parent = jj == 0 ? CC4_SLOT_ROUTER + sn : jj <= nu ? sn * 16 + 6 + (jj - 1) : sn * 16 + (jj - 1 - nu); }
            q += m;
        }
        bp64 |= (uint64_t)(uint8_t)parent << (8 * b);
    }
    /* green agents :700-749 */
    int ng = 0;
    uint64_t cum = 0;
    _Pragma("unroll 1") for (int sn = 0; sn < 8; sn++) { ng += NUSER(sn); cum |= (uint64_t)ng << (8 * sn); }
    /* red agents :751-812 */
    _Pragma("unroll 1") for (int r = 0; r < CC4_NRED; r++) {
        const int si = (int)ri(r == 5 ? 3u : 1u);
        const int sn = r == 0 ? S_CON : r <= 4 ? r - 1 : S_PAZ + si;
        const int q = (int)ri((uint32_t)(NUSER(sn) + NSERVER(sn)));
        rs64 |= (uint64_t)(uint8_t)(q < NUSER(sn) ? sn * 16 + 6 + q : sn * 16 + (q - NUSER(sn))) << (8 * r);
    }
    /* State.__init__ State.py:108-134: session processes (Host.add_session Host.py:189-196) */
    auto add_session_proc = [&](int slot, uint32_t kind) -> uint32_t {
        if (slot >= CC4_SLOT_ROUTER) { const uint32_t pid = 1u + ri(9); t->router_pid[slot - CC4_SLOT_ROUTER] = (uint8_t)pid; return pid; }
        const uint32_t pid = (uint32_t)t->pmax[slot] + 1u + ri(9);
        const int n = t->np[slot];
        const uint32_t pw = P_MAKE(pid, kind, 0, 0);
        k->parena[slot * CC4_OSLAB + n] = pw; st->o_procs[slot][n] = pw;
        t->np[slot] = (uint8_t)(n + 1); t->pmax[slot] = (uint16_t)pid;
        return pid;
    };
    _Pragma("unroll 1") for (int b = 0; b < CC4_NBLUE; b++) {
        add_session_proc(BLUE_PARENT(b), K_SESSION);
        _Pragma("unroll 1") for (int i = 0; i < (b == 4 ? 3 : 1); i++) {
            const int sn = b < 4 ? b : S_PAZ + i;
            if (BLUE_PARENT(b) != CC4_SLOT_ROUTER + sn) add_session_proc(CC4_SLOT_ROUTER + sn, K_SESSION);
            _Pragma("unroll 1") for (int q = 0; q < NUSER(sn); q++) if (BLUE_PARENT(b) != sn * 16 + 6 + q) add_session_proc(sn * 16 + 6 + q, K_SESSION);
            _Pragma("unroll 1") for (int q = 0; q < NSERVER(sn); q++) if (BLUE_PARENT(b) != sn * 16 + q) add_session_proc(sn * 16 + q, K_SESSION);
        }
    }
    _Pragma("unroll 1") for (int sn = 0; sn < 8; sn++)         /* one green agent per user host, creation order */
        _Pragma("unroll 1") for (int q = 0; q < NUSER(sn); q++) add_session_proc(sn * 16 + 6 + q, K_SESSION);
    const int h0 = RED_START(0);
    const uint32_t pid0 = add_session_proc(h0, K_REDABS);
    /* ---- the records ---- */
    _Pragma("unroll 1") for (int h = 0; h < 128; h++) {
        const int n = t->np[h];
        k->pdir[h] = PD_MAKE(h * CC4_OSLAB, n, 3); k->pmax[h] = t->pmax[h]; st->o_nproc[h] = (uint8_t)n;
    }
    k->arena_top = 128 * CC4_OSLAB;
    _Pragma("unroll 1") for (int i = 0; i < 16; i++) st->router_pid[i] = t->router_pid[i];
    _Pragma("unroll 1") for (int b = 0; b < CC4_NBLUE; b++) st->blue_parent[b] = (uint8_t)BLUE_PARENT(b);
    _Pragma("unroll 1") for (int r = 0; r < CC4_NRED; r++) st->red_start[r] = (uint8_t)RED_START(r);
    _Pragma("unroll 1") for (int a = 0; a < 16; a++) s->inprog[a] = IP_NONE;
    _Pragma("unroll 1") for (int sn = 0; sn < 8; sn++) { s->nuser[sn] = (uint8_t)NUSER(sn); s->nserver[sn] = (uint8_t)NSERVER(sn); }
    s->green_cum = cum; s->ngreen = (uint8_t)ng;
    s->sess[0] = S_MAKE(pid0, 0, h0); s->sessf[0] = (uint8_t)(0 | 1 << 3); s->n_sess = 1; s->sess_count[0] = 1; s->red_active = 1;
    /* initial knowledge: AgentInterface.set_init_obs AgentInterface.py:110-117 with the OSINT of :784-785 */
    _Pragma("unroll 1") for (int r = 0; r < CC4_NRED; r++) { const int h = RED_START(r); set128(s->ip_known[r], h); set128(s->hn_known[r], h); s->sn_known[r] = (uint16_t)(1u << (h >> 4)); }
    s->ss[0][0] = 0; s->n_ss[0] = 1;
    /* red_agent_0 digests its initial observation (its start host becomes 'U', FiniteStateRedAgent.py:221-224) */
    s->fsm_order[0][0] = (uint8_t)h0; s->n_fsm[0] = 1; s->fsm_started = 1;
    set128(s->fsm[0].present, h0); set128(s->fsm[0].hn, h0); set128(s->fsm[0].pl[2], h0);
    s->err = err;
    #undef NUSER
    #undef NSERVER
    #undef BLUE_PARENT
    #undef RED_START
}

/* observation rows and action masks of freshly reset episodes: warp per episode, the record comes back through shared memory */
__global__ void __launch_bounds__(32) k_reset_outputs(Params p, uint8_t *obs, uint8_t *mask)
{
    WarpSmem *w = reinterpret_cast<WarpSmem *>(smem_raw);
    const int env = blockIdx.x, lane = lane_id();
    {   /* of a freshly reset record the writers only look at the host counts; events, blocks, messages and the phase are zero */
        Cc4Dyn *d = &w->dyn; const Cc4Dyn *g = p.dyn + env;
        for (int i = lane; i < (int)(sizeof(d->ev) / 4); i += 32) reinterpret_cast<uint32_t *>(d->ev)[i] = 0;
        if (lane < 12) d->blocks[lane] = 0;
        if (lane < 8) { d->msgs[lane] = 0; d->nuser[lane] = g->nuser[lane]; d->nserver[lane] = g->nserver[lane]; }
        if (lane == 0) d->phase = 0;
        __syncwarp();
    }
    if (obs) w_write_obs(&w->dyn, &w->x, obs, env);
    if (mask) w_write_mask(&w->dyn, mask + (size_t)env * CC4_N_BLUE * CC4_MASK_LEN);
}

#endif  /* !CC4_PROD_ONLY */

}  // namespace CC4_NS
