/*
 * cc4_engine.cu — kernels and C-ABI (include/cc4_b200.h) of the B200-native batched CC4 step engine.
 *
 * Kernels (sm_100a):
 *   k_reset_fast + k_reset_outputs   EnterpriseScenarioGenerator.create_scenario + State.__init__ + controller reset in production
 *             mode: one THREAD per episode (the reset stream is one serial chain), then the observation rows and masks
 *   k_reset   the same with one warp per episode and the per-episode shared-memory context the tape reader needs (tape mode)
 *   k_step<RV>    one SimulationController.step tick for every episode (+ observation vectors, reward, done): one block per
 *             SM keeps up to 64 episodes' packed records in shared memory and walks them through the phases of the tick;
 *             RV = the scripted red agent class (FiniteStateRedAgent, DiscoveryFSRed, RandomSelectRedAgent), a template parameter
 *   k_rollout<RV> n fused ticks with the benchmark's synthetic random blue policy (state stays in shared memory)
 * Citations are file:line under /root/reference/CybORG/.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <sched.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

/* the device code is compiled into two namespaces: cc4 (with the tape paths; also the reset kernels) and cc4p (production
   k_step / k_rollout without them).  Profiling builds and -DCC4_WITH_TAPE=0/1 given on the command line keep one namespace. */
#if defined(CC4_WITH_TAPE) || defined(CC4_PHASE_CLOCKS) || defined(CC4_PHASE_WALL)
#define CC4_SINGLE_NS 1
#else
#define CC4_SINGLE_NS 0
#endif
#define CC4_NS cc4
#define CC4_PROD_ONLY 0
#if !CC4_SINGLE_NS
#define CC4_WITH_TAPE 1
#endif
#include "cc4_device.cuh"
#include "cc4_kernels.cuh"
#if !CC4_SINGLE_NS
#undef CC4_NS
#undef CC4_PROD_ONLY
#undef CC4_WITH_TAPE
#define CC4_NS cc4p
#define CC4_PROD_ONLY 1
#define CC4_WITH_TAPE 0
#include "cc4_device.cuh"
#include "cc4_kernels.cuh"
#undef CC4_WITH_TAPE
#define CC4_WITH_TAPE 1
#else
namespace cc4p = cc4;
#endif
#undef CC4_NS
#undef CC4_PROD_ONLY


/* ================================================================================================ */
/* host side: C-ABI                                                                                 */
/* ================================================================================================ */
using namespace cc4;

static thread_local char g_err[512];
static int set_err(int code, const char *fmt, const char *detail)
{
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return set_err(CC4_ERR_CUDA, #expr ": %s", cudaGetErrorString(e_)); } while (0)

/* ---- host side of cc4_step_host: observation rows arrive as bit rows and are expanded to bytes by a few threads ---- */
static uint64_t g_spread[256];      /* byte -> its eight bits as eight bytes (bit i -> byte i) */
static std::once_flag g_spread_once;
static void spread_init() { for (int v = 0; v < 256; v++) { uint64_t w = 0; for (int i = 0; i < 8; i++) w |= (uint64_t)(v >> i & 1) << (8 * i); g_spread[v] = w; } }
/* bit rows uint32 [cnt][20] -> uint8 [cnt][578] (compact) or [cnt][5][210] (padded): the layouts w_write_obs produces.
   Portable form: a 256-entry table turns one byte of bits into eight bytes. */
static void expand_rows_generic(const uint32_t *rows, uint8_t *dst, size_t cnt, bool compact)
{
    for (size_t e = 0; e < cnt; e++) {
        const uint32_t *row = rows + e * CC4_OBS_BITS_WORDS;
        const uint8_t *b = reinterpret_cast<const uint8_t *>(row);
        const uint8_t ph = (uint8_t)row[CC4_OBS_BITS_WORDS - 1];
        if (compact) {
            uint8_t *o = dst + e * CC4_OBS_COMPACT_LEN;
            for (int k = 0; k < 72; k++) memcpy(o + 8 * k, &g_spread[b[k]], 8);
            o[576] = b[72] & 1u; o[577] = (b[72] >> 1) & 1u;
            o[0] = o[92] = o[184] = o[276] = o[368] = ph;
        } else {
            uint8_t *o = dst + e * (size_t)(CC4_N_BLUE * CC4_OBS_LEN);
            for (int a = 0; a < CC4_N_BLUE; a++) {
                const int len = a == 4 ? 210 : 92, bit0 = 92 * a;
                uint8_t *oa = o + a * CC4_OBS_LEN;
                int j = 0;
                for (; j + 8 <= len; j += 8) {      /* eight entries from an unaligned 16-bit window (rows are 80 bytes: no read past the end) */
                    const int pos = bit0 + j;
                    uint16_t w16; memcpy(&w16, b + (pos >> 3), 2);
                    memcpy(oa + j, &g_spread[(w16 >> (pos & 7)) & 0xFFu], 8);
                }
                for (; j < len; j++) { const int pos = bit0 + j; oa[j] = (b[pos >> 3] >> (pos & 7)) & 1u; }
                oa[0] = ph;
                memset(oa + len, 0, CC4_OBS_LEN - len);
            }
        }
    }
}
#if defined(__x86_64__)
/* AVX2 form: 32 bits -> 32 bytes in four instructions.  `stream`: rows are built in a cache-resident block and leave it
   with streaming stores (see stream_rows for when). */
__attribute__((target("avx2"))) static inline __m256i spread32(uint32_t w)
{
    const __m256i sel = _mm256_setr_epi64x(0x0000000000000000ll, 0x0101010101010101ll, 0x0202020202020202ll, 0x0303030303030303ll);
    const __m256i bit = _mm256_set1_epi64x((long long)0x8040201008040201ull);
    const __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi32((int)w), sel);       /* byte j of the result = byte j / 8 of w */
    return _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_and_si256(v, bit), bit), _mm256_set1_epi8(1));
}
__attribute__((target("avx2"))) static void expand_rows_avx2(const uint32_t *rows, uint8_t *dst, size_t cnt, bool compact, bool stream)
{
    const size_t row_bytes = compact ? CC4_OBS_COMPACT_LEN : CC4_N_BLUE * CC4_OBS_LEN;
    constexpr size_t BLOCK = 16;
    alignas(64) uint8_t tmp[BLOCK * CC4_N_BLUE * CC4_OBS_LEN + 64];
    alignas(32) uint8_t crow[19 * 32];
    for (size_t e0 = 0; e0 < cnt; e0 += BLOCK) {
        const size_t nb = cnt - e0 < BLOCK ? cnt - e0 : BLOCK;
        uint8_t *out = stream ? tmp : dst + e0 * row_bytes;
        for (size_t e = 0; e < nb; e++) {
            const uint32_t *r = rows + (e0 + e) * CC4_OBS_BITS_WORDS;
            uint8_t *o = compact ? out + e * CC4_OBS_COMPACT_LEN : crow;
            for (int k = 0; k < 18; k++) _mm256_storeu_si256((__m256i *)(o + 32 * k), spread32(r[k]));
            o[576] = r[18] & 1u; o[577] = (r[18] >> 1) & 1u;
            o[0] = o[92] = o[184] = o[276] = o[368] = (uint8_t)r[CC4_OBS_BITS_WORDS - 1];
            if (!compact) {
                uint8_t *p = out + e * row_bytes;
                for (int a = 0; a < CC4_N_BLUE; a++) {
                    const int len = a == 4 ? 210 : 92;
                    memcpy(p + a * CC4_OBS_LEN, crow + 92 * a, len);
                    memset(p + a * CC4_OBS_LEN + len, 0, CC4_OBS_LEN - len);
                }
            }
        }
        if (stream) {
            uint8_t *d = dst + e0 * row_bytes; const uint8_t *s = tmp; size_t n = nb * row_bytes;
            while (n && ((uintptr_t)d & 31u)) { *d++ = *s++; n--; }
            for (; n >= 32; n -= 32, d += 32, s += 32) _mm256_stream_si256((__m256i *)d, _mm256_loadu_si256((const __m256i *)s));
            memcpy(d, s, n);
        }
    }
    if (stream) _mm_sfence();       /* the streamed rows are visible before the job is reported done */
}
#endif
static void expand_rows(const uint32_t *rows, uint8_t *dst, size_t cnt, bool compact, bool stream)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("CC4_HOST_NO_AVX2");
    if (avx2) { expand_rows_avx2(rows, dst, cnt, compact, stream); return; }
#endif
    (void)stream;
    expand_rows_generic(rows, dst, cnt, compact);
}
/* Streaming stores pay when the rows of one tick do not fit the host's caches: with all ranks of the node together above
   64 MB (two or more GPUs at 65 536 episodes each; measured on four: 1.33 G agent-steps/s against 1.20 G), not for one rank
   (38 MB: 568 M against 612 M).  CC4_HOST_STREAM=0 / 1 overrides. */
static bool stream_rows(size_t bytes_per_tick)
{
    if (const char *e = getenv("CC4_HOST_STREAM")) return e[0] == '1';
    size_t lw = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) lw = atoi(e) > 0 ? (size_t)atoi(e) : 1;
    return lw * bytes_per_tick > ((size_t)64 << 20);
}
struct ExpandJob { const uint32_t *rows; uint8_t *dst; size_t cnt; bool compact, stream; };
/* The pool behind cc4_step_host.  Jobs go through a ring with monotonically increasing publish / claim / done counters;
 * idle workers spin for a few hundred microseconds before they go to sleep, because within a tick the chunks arrive
 * every ~60 us and a futex wake-up per chunk (and one more for the caller) would sit on the critical path of the call. */
struct HostPool {
    static constexpr unsigned RING = 512;
    std::vector<std::thread> th;
    ExpandJob ring[RING];
    std::atomic<unsigned> published{0}, next{0}, done{0};
    std::atomic<int> sleepers{0};
    std::atomic<bool> stop{false};
    std::mutex m;
    std::condition_variable cv;
    static void relax()
    {
#if defined(__x86_64__)
        _mm_pause();
#endif
    }
    bool try_run()
    {
        unsigned i = next.load();
        while (i != published.load()) {
            if (next.compare_exchange_weak(i, i + 1)) {
                const ExpandJob j = ring[i % RING];
                expand_rows(j.rows, j.dst, j.cnt, j.compact, j.stream);
                done.fetch_add(1);
                return true;
            }
        }
        return false;
    }
    void run()
    {
        for (;;) {
            int idle = 0;
            while (!try_run()) {
                if (stop.load()) return;
                if (++idle < 20000) { relax(); continue; }
                std::unique_lock<std::mutex> lk(m);         /* nothing for a while: sleep until the next submit */
                sleepers.fetch_add(1);
                cv.wait(lk, [&] { return stop.load() || next.load() != published.load(); });
                sleepers.fetch_sub(1);
                idle = 0;
            }
        }
    }
    /* CC4_HOST_PIN=1 pins every rank's workers to the rank's own share of the CPUs this process may run on (rank r of w takes
       the r-th of w equal slices of the affinity mask).  Off by default: the scheduler does better on its own (see below). */
    void start(int n)
    {
        std::vector<int> cpus;
#if defined(__linux__)
        int lw = 1, lr = 0;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) lw = atoi(e) > 0 ? atoi(e) : 1;
        if (const char *e = getenv("LOCAL_RANK")) lr = atoi(e) >= 0 ? atoi(e) : 0;
        bool pin = false;       /* measured on 8 ranks x 4 workers on 32 hardware threads: 1.24 G agent-steps/s pinned, 1.35 G free (profiles/r03f) */
        if (const char *e = getenv("CC4_HOST_PIN")) pin = e[0] == '1';
        cpu_set_t set;
        if (pin && sched_getaffinity(0, sizeof set, &set) == 0) {
            std::vector<int> all;
            for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &set)) all.push_back(c);
            const size_t per = all.size() / (size_t)lw;
            if (per >= 1 && lr < lw) cpus.assign(all.begin() + (size_t)lr * per, all.begin() + (size_t)(lr + 1) * per);
        }
#endif
        for (int i = 0; i < n; i++) {
            th.emplace_back([this] { run(); });
#if defined(__linux__)
            if (!cpus.empty()) {
                cpu_set_t one; CPU_ZERO(&one); CPU_SET(cpus[(size_t)i % cpus.size()], &one);
                pthread_setaffinity_np(th.back().native_handle(), sizeof one, &one);
            }
#endif
        }
    }
    void submit(const uint32_t *rows, uint8_t *dst, size_t cnt, bool compact, size_t row_bytes, bool stream)
    {   /* one job per thread; the caller waits for every job of a tick before the next tick (at most 16 chunks x 16 jobs in flight) */
        const size_t parts = th.size(), per = (cnt + parts - 1) / parts;
        unsigned p = published.load();
        for (size_t o = 0; o < cnt; o += per) ring[p++ % RING] = { rows + o * CC4_OBS_BITS_WORDS, dst + o * row_bytes, cnt - o < per ? cnt - o : per, compact, stream };
        published.store(p);
        if (sleepers.load() > 0) { std::lock_guard<std::mutex> lk(m); cv.notify_all(); }
    }
    void wait() { while (done.load() != published.load()) if (!try_run()) relax(); }      /* the caller lends a hand with what is left */
    ~HostPool()
    {
        stop.store(true);
        { std::lock_guard<std::mutex> lk(m); cv.notify_all(); }
        for (auto &t : th) t.join();
    }
};
#define CC4_MAX_CHUNKS 16
constexpr size_t HOST_BITS_ROW = CC4_OBS_BITS_WORDS * sizeof(uint32_t);      /* 80 */
constexpr size_t HOST_ROW = HOST_BITS_ROW + sizeof(float) + 1;               /* + reward + done = 85 bytes per episode through PCIe */

struct cc4_handle {
    Params p;
    int device, n_sm;
    uint64_t launches;
    int32_t *stage_actions; uint8_t *stage_msgs, *stage_obs, *stage_mask, *stage_done; float *stage_reward;   /* device staging for cc4_step_host */
    cudaStream_t pipe[2]; cudaEvent_t ev_in, ev_out[2];   /* chunk pipeline of cc4_step_host */
    cudaEvent_t ev_chunk[CC4_MAX_CHUNKS];                 /* ... chunk k's outputs are on the host */
    uint32_t *host_bits;                                  /* pinned + mapped: the bit rows (+ rewards, done flags) of the batch */
    void *host_bits_dev;                                  /* ... as the device sees it */
    int zero_copy;                                        /* CC4_HOST_ZEROCOPY=1: the kernel writes those outputs straight to host memory (default: staging + one copy per chunk) */
    HostPool *pool;
    int sweep, alternate;                                 /* direction of the last cc4_step over the batch; CC4_SWEEP_ALTERNATE=0 keeps one direction */
    int host_expand;                                      /* 0: byte rows over PCIe (CC4_HOST_EXPAND=0) */
    int host_chunk;                                       /* episodes per chunk of cc4_step_host, 0 = one wave of blocks (CC4_HOST_CHUNK_EPISODES) */
    int reset_serial;                                     /* CC4_RESET_SERIAL=1: k_reset (warp per episode) also in production mode */
};

extern "C" const char *cc4_last_error(void) { return g_err; }

extern "C" int cc4_create(int n_envs, int device, uint64_t seed, uint64_t first_episode, int steps, int rng_mode, cc4_handle **out)
{
    if (!out || n_envs <= 0 || steps < 3 || steps > 60000 || (rng_mode != 0 && rng_mode != 1)) return set_err(CC4_ERR_ARG, "cc4_create: bad argument%s", "");
    CUDA_TRY(cudaSetDevice(device));
    cc4_handle *h = new (std::nothrow) cc4_handle();
    if (!h) return set_err(CC4_ERR_STATE, "cc4_create: out of host memory%s", "");
    memset(h, 0, sizeof *h);
    h->device = device;
    h->p.n_envs = n_envs; h->p.steps = steps; h->p.rng_mode = rng_mode;
    CUDA_TRY(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device));
    h->p.k0 = (uint32_t)seed; h->p.k1hi = (uint32_t)(seed >> 32); h->p.first_episode = first_episode;
    { const char *e = getenv("CC4_HOST_EXPAND"); h->host_expand = !(e && e[0] == '0'); }
    { const char *e = getenv("CC4_SWEEP_ALTERNATE"); h->alternate = !(e && e[0] == '0'); }
    { const char *e = getenv("CC4_HOST_CHUNK_EPISODES"); h->host_chunk = e ? atoi(e) : 0; }
    { const char *e = getenv("CC4_HOST_ZEROCOPY"); h->zero_copy = e && e[0] == '1'; }      /* off: measured 643 M against 663 M agent-steps/s with staging + one copy */
    CUDA_TRY(cudaMalloc(&h->p.dyn, sizeof(Cc4Dyn) * (size_t)n_envs));
    CUDA_TRY(cudaMalloc(&h->p.stat, sizeof(Cc4Stat) * (size_t)n_envs));
    CUDA_TRY(cudaMalloc(&h->p.cold, sizeof(Cc4Cold) * (size_t)n_envs));
    CUDA_TRY(cudaMemset(h->p.dyn, 0, sizeof(Cc4Dyn) * (size_t)n_envs));
    const int tick_smem = (int)(sizeof(WarpSmem) * EPB_MAX);
    CUDA_TRY(cudaFuncSetAttribute(k_step<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_step<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_rollout<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_rollout<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_rollout<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
#if !CC4_SINGLE_NS
    CUDA_TRY(cudaFuncSetAttribute(cc4p::k_step<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(cc4p::k_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(cc4p::k_step<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(cc4p::k_rollout<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(cc4p::k_rollout<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
    CUDA_TRY(cudaFuncSetAttribute(cc4p::k_rollout<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tick_smem));
#endif
    CUDA_TRY(cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ResetSmem)));
    CUDA_TRY(cudaFuncSetAttribute(k_reset_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ResetLane) * RESET_FAST_THREADS)));
    CUDA_TRY(cudaFuncSetAttribute(k_reset_outputs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpSmem)));
    { const char *e = getenv("CC4_RESET_SERIAL"); h->reset_serial = e && e[0] == '1'; }      /* diagnostics: the warp-per-episode reset in production mode too */
    *out = h;
    return CC4_OK;
}

extern "C" int cc4_destroy(cc4_handle *h)
{
    if (!h) return CC4_OK;
    cudaSetDevice(h->device);
    cudaFree(h->p.dyn); cudaFree(h->p.stat); cudaFree(h->p.cold);
    for (int i = 0; i < 2; i++) { if (h->pipe[i]) cudaStreamDestroy(h->pipe[i]); if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]); }
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    cudaFree(h->stage_actions); cudaFree(h->stage_msgs); cudaFree(h->stage_obs); cudaFree(h->stage_mask); cudaFree(h->stage_done); cudaFree(h->stage_reward);
    for (int i = 0; i < CC4_MAX_CHUNKS; i++) if (h->ev_chunk[i]) cudaEventDestroy(h->ev_chunk[i]);
    if (h->host_bits) cudaFreeHost(h->host_bits);
    delete h->pool;
    delete h;
    return CC4_OK;
}

extern "C" int cc4_set_episode_base(cc4_handle *h, uint64_t seed, uint64_t first_episode)
{
    if (!h) return set_err(CC4_ERR_ARG, "cc4_set_episode_base: null handle%s", "");
    h->p.k0 = (uint32_t)seed; h->p.k1hi = (uint32_t)(seed >> 32); h->p.first_episode = first_episode;
    return CC4_OK;
}

extern "C" int cc4_set_red_agent(cc4_handle *h, int variant)
{
    if (!h || (variant != CC4_RED_FINITE_STATE && variant != CC4_RED_DISCOVERY && variant != CC4_RED_RANDOM_SELECT)) return set_err(CC4_ERR_ARG, "cc4_set_red_agent: bad argument%s", "");
    h->p.red_variant = variant;
    return CC4_OK;
}

extern "C" int cc4_state_bytes(cc4_handle *h, size_t *hot_bytes, size_t *cold_bytes, size_t *stat_bytes)
{
    (void)h;
    if (hot_bytes) *hot_bytes = sizeof(Cc4Dyn);
    if (cold_bytes) *cold_bytes = sizeof(Cc4Cold);
    if (stat_bytes) *stat_bytes = sizeof(Cc4Stat);
    return CC4_OK;
}

extern "C" int cc4_set_tape(cc4_handle *h, const cc4_draw *recs_dev, const uint32_t *seg_off_dev, uint32_t n_segs)
{
    if (!h) return set_err(CC4_ERR_ARG, "cc4_set_tape: null handle%s", "");
    h->p.tape = recs_dev; h->p.seg_off = seg_off_dev; h->p.n_segs = n_segs;
    return CC4_OK;
}

extern "C" int cc4_reset(cc4_handle *h, uint8_t *obs_dev, uint8_t *mask_dev, void *stream)
{
    if (!h) return set_err(CC4_ERR_ARG, "cc4_reset: null handle%s", "");
    if (h->p.rng_mode == 1 && !h->p.tape) return set_err(CC4_ERR_STATE, "cc4_reset: tape mode without cc4_set_tape%s", "");
    if ((uintptr_t)obs_dev & 15u) return set_err(CC4_ERR_ARG, "cc4_reset: obs_dev must be 16-byte aligned%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->p.rng_mode == 1 || h->reset_serial) {
        k_reset<<<h->p.n_envs, 32, sizeof(ResetSmem), (cudaStream_t)stream>>>(h->p, obs_dev, mask_dev);
    } else {        /* production mode: thread per episode, then the observation rows / masks (see k_reset_fast) */
        k_reset_fast<<<(h->p.n_envs + RESET_FAST_THREADS - 1) / RESET_FAST_THREADS, RESET_FAST_THREADS, sizeof(ResetLane) * RESET_FAST_THREADS, (cudaStream_t)stream>>>(h->p);
        if (obs_dev || mask_dev) { k_reset_outputs<<<h->p.n_envs, 32, sizeof(WarpSmem), (cudaStream_t)stream>>>(h->p, obs_dev, mask_dev); h->launches++; }
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return CC4_OK;
}

static int ensure_staging(cc4_handle *h)
{
    if (h->stage_actions) return CC4_OK;
    size_t n = (size_t)h->p.n_envs;
    CUDA_TRY(cudaMalloc(&h->stage_actions, n * 5 * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&h->stage_msgs, n * 5));
    CUDA_TRY(cudaMalloc(&h->stage_obs, n * CC4_N_BLUE * CC4_OBS_LEN));
    CUDA_TRY(cudaMalloc(&h->stage_mask, n * CC4_N_BLUE * CC4_MASK_LEN));
    CUDA_TRY(cudaMalloc(&h->stage_done, n));
    CUDA_TRY(cudaMalloc(&h->stage_reward, n * sizeof(float)));
    return CC4_OK;
}

extern "C" int cc4_reset_host(cc4_handle *h, uint8_t *obs_host, uint8_t *mask_host, void *stream)
{
    if (!h) return set_err(CC4_ERR_ARG, "cc4_reset_host: null handle%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = ensure_staging(h);
    if (rc) return rc;
    size_t n = (size_t)h->p.n_envs;
    cudaStream_t st = (cudaStream_t)stream;
    rc = cc4_reset(h, obs_host ? h->stage_obs : nullptr, mask_host ? h->stage_mask : nullptr, stream);
    if (rc) return rc;
    if (obs_host) CUDA_TRY(cudaMemcpyAsync(obs_host, h->stage_obs, n * CC4_N_BLUE * CC4_OBS_LEN, cudaMemcpyDeviceToHost, st));
    if (mask_host) CUDA_TRY(cudaMemcpyAsync(mask_host, h->stage_mask, n * CC4_N_BLUE * CC4_MASK_LEN, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CC4_OK;
}

/* Episodes per block for a launch over n episodes: the fewest whole waves of one block per SM that hold them, then the
 * episodes spread evenly over those waves (65 536 episodes on 148 SMs: 9 waves x 148 blocks x 50 episodes, the last wave
 * 86 % full, instead of 52 per block and a 10th wave at 52 %). */
static int episodes_per_block(const cc4_handle *h, size_t n)
{
    size_t per_wave = (size_t)h->n_sm * EPB_MAX * CC4_MIN_BLOCKS, waves = (n + per_wave - 1) / per_wave;
    size_t epb = (n + waves * h->n_sm * CC4_MIN_BLOCKS - 1) / (waves * h->n_sm * CC4_MIN_BLOCKS);
    return (int)(epb < 1 ? 1 : epb > EPB_MAX ? EPB_MAX : epb);
}

/* k_step is instantiated once per scripted red agent class (see cc4_device.cuh: template <int RV>) */
static void launch_step(cc4_handle *h, int blocks, int epb, cudaStream_t st, int env0, int env_end, const StepIO &io)
{
    const size_t smem = sizeof(WarpSmem) * (size_t)epb;
    if (h->p.rng_mode == 1) {       /* tape replay: the instantiations that carry the tape paths */
        switch (h->p.red_variant) {
        case CC4_RED_DISCOVERY: cc4::k_step<1><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(h->p, env0, env_end, epb, io); break;
        case CC4_RED_RANDOM_SELECT: cc4::k_step<2><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(h->p, env0, env_end, epb, io); break;
        default: cc4::k_step<0><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(h->p, env0, env_end, epb, io); break;
        }
        return;
    }
    switch (h->p.red_variant) {
    case CC4_RED_DISCOVERY: cc4p::k_step<1><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(h->p, env0, env_end, epb, io); break;
    case CC4_RED_RANDOM_SELECT: cc4p::k_step<2><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(h->p, env0, env_end, epb, io); break;
    default: cc4p::k_step<0><<<blocks, WARPS_PER_BLOCK * 32, smem, st>>>(h->p, env0, env_end, epb, io); break;
    }
}

extern "C" int cc4_step(cc4_handle *h, const int32_t *actions_dev, const uint8_t *msgs_dev, uint8_t *obs_dev, float *reward_dev,
                        uint8_t *done_dev, int flags, void *stream)
{
    if (!h || !actions_dev) return set_err(CC4_ERR_ARG, "cc4_step: null handle or actions%s", "");
    /* the observation writer stores whole 16-byte lines relative to obs_dev (bit rows: words); reward / actions are 4-byte items */
    if (((uintptr_t)obs_dev & ((flags & CC4_STEP_OBS_BITS) ? 3u : 15u)) || ((uintptr_t)reward_dev & 3u) || ((uintptr_t)actions_dev & 3u))
        return set_err(CC4_ERR_ARG, "cc4_step: misaligned buffer (obs_dev: 16 bytes, 4 for bit rows; reward_dev, actions_dev: 4 bytes)%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    const int epb = episodes_per_block(h, (size_t)h->p.n_envs), blocks = (h->p.n_envs + epb - 1) / epb;
    /* Successive ticks walk the batch in opposite directions: the per-tick footprint (hot records + the cold lines the
       actions touch) is larger than the 126 MB L2, so a tick that starts where the last one ended finds the most recently
       written records still there, while a fixed direction finds none of them (cyclic sweep over an LRU-like cache). */
    h->sweep ^= h->alternate;
    StepIO io = { actions_dev, msgs_dev, obs_dev, reward_dev, done_dev, flags, h->sweep };
    launch_step(h, blocks, epb, (cudaStream_t)stream, 0, h->p.n_envs, io);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return CC4_OK;
}

/* Host-buffer tick.  The batch is cut into chunks that flow through two internal streams: the device->host copy of
 * chunk i (observations dominate: 1 050 B per episode) overlaps the kernel of chunk i+1, so the call costs about
 * max(kernel, copy) instead of their sum.  The call is synchronous with respect to `stream`. */
#ifndef CC4_HOST_CHUNKS
#define CC4_HOST_CHUNKS 8
#endif
extern "C" int cc4_step_host(cc4_handle *h, const int32_t *actions_host, const uint8_t *msgs_host, uint8_t *obs_host, float *reward_host,
                             uint8_t *done_host, int flags, void *stream)
{
    if (!h || !actions_host) return set_err(CC4_ERR_ARG, "cc4_step_host: null handle or actions%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    { int rc0 = ensure_staging(h); if (rc0) return rc0; }
    const size_t n = (size_t)h->p.n_envs;
    cudaStream_t user = (cudaStream_t)stream;
    if (!h->pipe[0]) {
        for (int i = 0; i < 2; i++) { CUDA_TRY(cudaStreamCreateWithFlags(&h->pipe[i], cudaStreamNonBlocking)); CUDA_TRY(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming)); }
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    }
    /* order after whatever the caller queued on its stream (an earlier reset / device-side step) */
    CUDA_TRY(cudaEventRecord(h->ev_in, user));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamWaitEvent(h->pipe[i], h->ev_in, 0));
    const bool want_obs = obs_host && (flags & CC4_STEP_OBS), want_rew = reward_host && (flags & CC4_STEP_REWARD);
    const bool raw_bits = (flags & CC4_STEP_OBS_BITS) != 0, compact = (flags & CC4_STEP_OBS_COMPACT) != 0;
    const size_t obs_row = raw_bits ? CC4_OBS_BITS_WORDS * sizeof(uint32_t) : compact ? CC4_OBS_COMPACT_LEN : CC4_N_BLUE * CC4_OBS_LEN;
    size_t chunk = (n + CC4_HOST_CHUNKS - 1) / CC4_HOST_CHUNKS;
    {   /* whole waves of full blocks per chunk */
        size_t wave = (size_t)h->n_sm * EPB_MAX;
        chunk = (chunk + wave - 1) / wave * wave;
    }
    if (h->host_chunk > 0) chunk = ((size_t)h->host_chunk + EPB_MAX - 1) / EPB_MAX * EPB_MAX;       /* CC4_HOST_CHUNK_EPISODES (tuning) */
    if (want_obs && !raw_bits && h->host_expand) {
        /* byte rows wanted: the kernel writes bit rows (80 B per episode instead of 578 / 1 050), they cross PCIe into pinned
           staging, and the pool expands chunk k on the host while the chunks after it are still on the device */
        if (!h->host_bits) {        /* first call: pool, pinned staging, per-chunk events */
            std::call_once(g_spread_once, spread_init);
            int nt = 0;
            if (const char *e = getenv("CC4_HOST_THREADS")) nt = atoi(e);
            if (nt <= 0) {
                int hw = (int)std::thread::hardware_concurrency(), lw = 1;
                if (const char *e = getenv("LOCAL_WORLD_SIZE")) lw = atoi(e) > 0 ? atoi(e) : 1;
                nt = hw / lw; if (nt > 8) nt = 8;       /* this rank's share of the hardware threads, at most 8: one GPU on 16 hardware threads
                                                           measures 664 M agent-steps/s with 8 workers and 628 M with 16 (they spin while a tick is
                                                           in flight); four ranks on 32 threads: 1.00 G with 8 workers each, 0.80 G with 4 */
            }
            if (nt < 1) nt = 1;
            if (!h->pool) {
                h->pool = new (std::nothrow) HostPool();
                if (!h->pool) return set_err(CC4_ERR_STATE, "cc4_step_host: out of host memory%s", "");
                h->pool->start(nt);
            }
            for (int i = 0; i < CC4_MAX_CHUNKS; i++) if (!h->ev_chunk[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming));
            CUDA_TRY(cudaHostAlloc(&h->host_bits, n * HOST_ROW, cudaHostAllocMapped));
            CUDA_TRY(cudaHostGetDevicePointer(&h->host_bits_dev, h->host_bits, 0));
        }
        if ((n + chunk - 1) / chunk > CC4_MAX_CHUNKS) return set_err(CC4_ERR_STATE, "cc4_step_host: too many chunks%s", "");
        timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
        const double t_start = ts0.tv_sec * 1e6 + ts0.tv_nsec * 1e-3;
        /* staging (device and pinned host alike): chunk [e0, e0 + cnt) owns bytes [e0 * 85, (e0 + cnt) * 85) = its bit rows, then its
           rewards, then its done flags, so that ONE copy brings a chunk's outputs to the host (e0 and the chunk size are
           multiples of 64 episodes: every part stays 16-byte aligned) */
        uint8_t *stage = h->stage_obs;          /* the byte-row staging area is larger than n * 85 */
        uint8_t *hstage = reinterpret_cast<uint8_t *>(h->host_bits);
        const int kflags = (flags & ~CC4_STEP_OBS_COMPACT) | CC4_STEP_OBS_BITS;
        {   /* chunk order, knob CC4_SWEEP_HOST: 0 forward always (default), 1 alternate like cc4_step, 2 reversed always.
               Alternating is SLOWER here (0.72 ms per tick against 0.55; 0 and 2 measure the same): the chunk that was expanded
               last is then the first whose bit rows are copied to the host again, while the host threads still hold those
               lines of the pinned staging area in their caches, and the first three chunks arrive 170 us late. */
            static const int mode = getenv("CC4_SWEEP_HOST") ? atoi(getenv("CC4_SWEEP_HOST")) : 0;
            h->sweep = mode == 1 ? (h->sweep ^ 1) : mode == 2 ? 1 : 0;
        }
        const int nch = (int)((n + chunk - 1) / chunk);
        int k = 0;
        for (int ci = 0; ci < nch; ci++, k ^= 1) {
            const size_t e0 = (size_t)(h->sweep ? nch - 1 - ci : ci) * chunk;
            size_t cnt = n - e0 < chunk ? n - e0 : chunk;
            cudaStream_t st = h->pipe[k];
            CUDA_TRY(cudaMemcpyAsync(h->stage_actions + e0 * 5, actions_host + e0 * 5, cnt * 5 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            if (msgs_host) CUDA_TRY(cudaMemcpyAsync(h->stage_msgs + e0 * 5, msgs_host + e0 * 5, cnt * 5, cudaMemcpyHostToDevice, st));
            const int epb = episodes_per_block(h, cnt), blocks = (int)((cnt + epb - 1) / epb);
            /* the kernel indexes rows / reward / done by global episode number: hand it bases that put episode e0 at the chunk's parts */
            /* zero copy (knob, off): the 85 bytes per episode are written over PCIe by the kernel itself as it goes, so a chunk is on
               the host when its kernel ends instead of one copy later; the kernel loses more than the copy costs */
            uint8_t *cb = (h->zero_copy ? static_cast<uint8_t *>(h->host_bits_dev) : stage) + e0 * HOST_ROW;
            StepIO io = { h->stage_actions, msgs_host ? h->stage_msgs : nullptr, cb - e0 * HOST_BITS_ROW,
                          reinterpret_cast<float *>(cb + cnt * HOST_BITS_ROW) - e0, cb + cnt * (HOST_BITS_ROW + sizeof(float)) - e0, kflags | CC4_STEP_REWARD, h->sweep };
            launch_step(h, blocks, epb, st, (int)e0, (int)(e0 + cnt), io);
            h->launches++;
            CUDA_TRY(cudaGetLastError());
            if (!h->zero_copy) CUDA_TRY(cudaMemcpyAsync(hstage + e0 * HOST_ROW, cb, cnt * HOST_ROW, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaEventRecord(h->ev_chunk[ci], st));
        }
        int c = 0;
        cudaError_t err = cudaSuccess;
        const bool stream = stream_rows(n * obs_row);
        static const bool trace = getenv("CC4_HOST_TRACE") != nullptr;      /* diagnostics: host timeline of one call in microseconds */
        auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
        const double t_enq = trace ? now_us() : 0.0;
        double t_ev[CC4_MAX_CHUNKS];
        for (; c < nch; c++) {
            const size_t e0 = (size_t)(h->sweep ? nch - 1 - c : c) * chunk;
            size_t cnt = n - e0 < chunk ? n - e0 : chunk;
            err = cudaEventSynchronize(h->ev_chunk[c]);
            if (err != cudaSuccess) break;
            if (trace) t_ev[c] = now_us();
            const uint8_t *hb = hstage + e0 * HOST_ROW;
            h->pool->submit(reinterpret_cast<const uint32_t *>(hb), obs_host + e0 * obs_row, cnt, compact, obs_row, stream);
            if (want_rew) memcpy(reward_host + e0, hb + cnt * HOST_BITS_ROW, cnt * sizeof(float));
            if (done_host) memcpy(done_host + e0, hb + cnt * (HOST_BITS_ROW + sizeof(float)), cnt);
        }
        h->pool->wait();
        if (trace) {
            const double t_done = now_us();
            fprintf(stderr, "cc4_step_host: enqueue %.0f us;", t_enq - t_start);
            for (int i = 0; i < c; i++) fprintf(stderr, " chunk%d +%.0f", i, t_ev[i] - t_start);
            fprintf(stderr, "; expanded +%.0f us\n", t_done - t_start);
        }
        if (err != cudaSuccess) return set_err(CC4_ERR_CUDA, "cc4_step_host: %s", cudaGetErrorString(err));
        /* every chunk's last operation has been waited for on the host: both internal streams are idle, and whatever the
           caller queues on its stream from now on starts after this tick (no event hand-back needed) */
        return CC4_OK;
    }
    int k = 0;
    for (size_t e0 = 0; e0 < n; e0 += chunk, k ^= 1) {
        size_t cnt = n - e0 < chunk ? n - e0 : chunk;
        cudaStream_t st = h->pipe[k];
        CUDA_TRY(cudaMemcpyAsync(h->stage_actions + e0 * 5, actions_host + e0 * 5, cnt * 5 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if (msgs_host) CUDA_TRY(cudaMemcpyAsync(h->stage_msgs + e0 * 5, msgs_host + e0 * 5, cnt * 5, cudaMemcpyHostToDevice, st));
        const int epb = episodes_per_block(h, cnt), blocks = (int)((cnt + epb - 1) / epb);
        StepIO io = { h->stage_actions, msgs_host ? h->stage_msgs : nullptr, want_obs ? h->stage_obs : nullptr,
                      want_rew ? h->stage_reward : nullptr, done_host ? h->stage_done : nullptr, flags };
        launch_step(h, blocks, epb, st, (int)e0, (int)(e0 + cnt), io);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        if (want_obs) CUDA_TRY(cudaMemcpyAsync(obs_host + e0 * obs_row, h->stage_obs + e0 * obs_row, cnt * obs_row, cudaMemcpyDeviceToHost, st));
        if (want_rew) CUDA_TRY(cudaMemcpyAsync(reward_host + e0, h->stage_reward + e0, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (done_host) CUDA_TRY(cudaMemcpyAsync(done_host + e0, h->stage_done + e0, cnt, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < 2; i++) { CUDA_TRY(cudaEventRecord(h->ev_out[i], h->pipe[i])); CUDA_TRY(cudaStreamWaitEvent(user, h->ev_out[i], 0)); }
    CUDA_TRY(cudaStreamSynchronize(user));
    return CC4_OK;
}

/* host-only helper: bit rows -> byte rows, single-threaded (what the pool of cc4_step_host runs per slice) */
extern "C" int cc4_expand_obs_bits(const uint32_t *rows_host, uint8_t *obs_host, size_t n_rows, int compact)
{
    if (!rows_host || !obs_host) return set_err(CC4_ERR_ARG, "cc4_expand_obs_bits: null%s", "");
    std::call_once(g_spread_once, spread_init);
    expand_rows(rows_host, obs_host, n_rows, compact != 0, stream_rows(0));
    return CC4_OK;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_fill_avx2(uint8_t *q, size_t n, int v)
{
    const __m256i x = _mm256_set1_epi8((char)v);
    for (size_t i = 0; i < n; i += 32) _mm256_stream_si256((__m256i *)(q + i), x);
    _mm_sfence();
}
#endif
/* host-only: what can `threads` host threads WRITE per second?  The bound of the host-buffer path on a node whose ranks
   together produce hundreds of MB of uint8 observation rows per tick (bench.py reports e2e against it). */
extern "C" int cc4_host_write_bandwidth(size_t bytes, int threads, int streaming, int iters, double *gbs_out)
{
    if (!gbs_out || bytes < 4096 || threads < 1 || iters < 1) return set_err(CC4_ERR_ARG, "cc4_host_write_bandwidth: bad argument%s", "");
    bytes &= ~(size_t)4095;
    uint8_t *buf = static_cast<uint8_t *>(aligned_alloc(4096, bytes));
    if (!buf) return set_err(CC4_ERR_STATE, "cc4_host_write_bandwidth: out of host memory%s", "");
    const size_t per = (bytes / (size_t)threads) & ~(size_t)63;
    auto pass = [&](int it) {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back([=] {
            uint8_t *d = buf + (size_t)t * per;
#if defined(__x86_64__)
            if (streaming && __builtin_cpu_supports("avx2")) { stream_fill_avx2(d, per, it); return; }
#endif
            memset(d, it & 0xFF, per);
        });
        for (auto &x : th) x.join();
    };
    pass(0);        /* page faults */
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 1; it <= iters; it++) pass(it);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *gbs_out = (double)per * threads * iters / dt / 1e9;
    free(buf);
    return CC4_OK;
}

extern "C" int cc4_rollout(cc4_handle *h, int n_ticks, uint64_t policy_seed, float *reward_sum_dev, int flags, void *stream)
{
    if (!h || n_ticks < 0) return set_err(CC4_ERR_ARG, "cc4_rollout: bad argument%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    const int epb = episodes_per_block(h, (size_t)h->p.n_envs), blocks = (h->p.n_envs + epb - 1) / epb;
    const uint32_t pk0 = (uint32_t)policy_seed ^ 0x5EEDu, pk1 = (uint32_t)(policy_seed >> 32);
    const size_t smem = sizeof(WarpSmem) * (size_t)epb;
    if (h->p.rng_mode == 1) {
        switch (h->p.red_variant) {
        case CC4_RED_DISCOVERY: cc4::k_rollout<1><<<blocks, WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(h->p, epb, n_ticks, pk0, pk1, reward_sum_dev, flags); break;
        case CC4_RED_RANDOM_SELECT: cc4::k_rollout<2><<<blocks, WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(h->p, epb, n_ticks, pk0, pk1, reward_sum_dev, flags); break;
        default: cc4::k_rollout<0><<<blocks, WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(h->p, epb, n_ticks, pk0, pk1, reward_sum_dev, flags); break;
        }
    } else
    switch (h->p.red_variant) {
    case CC4_RED_DISCOVERY: cc4p::k_rollout<1><<<blocks, WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(h->p, epb, n_ticks, pk0, pk1, reward_sum_dev, flags); break;
    case CC4_RED_RANDOM_SELECT: cc4p::k_rollout<2><<<blocks, WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(h->p, epb, n_ticks, pk0, pk1, reward_sum_dev, flags); break;
    default: cc4p::k_rollout<0><<<blocks, WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(h->p, epb, n_ticks, pk0, pk1, reward_sum_dev, flags); break;
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return CC4_OK;
}

/* profiling builds only (-DCC4_PHASE_CLOCKS): out_host uint64 [n_envs][18], see each_episode */
extern "C" int cc4_phase_clocks(cc4_handle *h, unsigned long long *out_host, int reset)
{
#ifdef CC4_PHASE_WALL       /* out_host uint64 [16]: wall cycles per phase summed over blocks and launches, [8] = blocks counted, [9] = whole blocks */
    if (!h || !out_host) return set_err(CC4_ERR_ARG, "cc4_phase_clocks: null%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out_host, g_wall, sizeof(unsigned long long) * 16));
    if (reset) { unsigned long long z[16] = { 0 }; CUDA_TRY(cudaMemcpyToSymbol(g_wall, z, sizeof z)); }
    return CC4_OK;
#elif defined(CC4_PHASE_CLOCKS)
    if (!h || !out_host) return set_err(CC4_ERR_ARG, "cc4_phase_clocks: null%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    size_t bytes = (size_t)h->p.n_envs * PC_STRIDE * sizeof(unsigned long long);
    if (!h->p.phase_clk) {
        CUDA_TRY(cudaMalloc(&h->p.phase_clk, bytes));
        CUDA_TRY(cudaMemset(h->p.phase_clk, 0, bytes));
        CUDA_TRY(cudaMemcpyToSymbol(g_phase_clk, &h->p.phase_clk, sizeof(void *)));
        unsigned long long *act = nullptr;
        CUDA_TRY(cudaMalloc(&act, 128 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(act, 0, 128 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemcpyToSymbol(g_action_clk, &act, sizeof(void *)));
    }
    CUDA_TRY(cudaMemcpy(out_host, h->p.phase_clk, bytes, cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(h->p.phase_clk, 0, bytes));
    return CC4_OK;
#else
    (void)h; (void)out_host; (void)reset;
    return set_err(CC4_ERR_STATE, "cc4_phase_clocks: this build has no phase clocks (-DCC4_PHASE_CLOCKS)%s", "");
#endif
}

/* profiling builds only: out_host uint64 [64][2] = cycles, count per action code since the counters were armed */
extern "C" int cc4_action_clocks(cc4_handle *h, unsigned long long *out_host)
{
#ifdef CC4_PHASE_CLOCKS
    if (!h || !out_host) return set_err(CC4_ERR_ARG, "cc4_action_clocks: null%s", "");
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long *act = nullptr;
    CUDA_TRY(cudaMemcpyFromSymbol(&act, g_action_clk, sizeof(void *)));
    if (!act) return set_err(CC4_ERR_STATE, "cc4_action_clocks: call cc4_phase_clocks first%s", "");
    CUDA_TRY(cudaMemcpy(out_host, act, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return CC4_OK;
#else
    (void)h; (void)out_host;
    return set_err(CC4_ERR_STATE, "cc4_action_clocks: this build has no phase clocks (-DCC4_PHASE_CLOCKS)%s", "");
#endif
}

extern "C" int cc4_launch_count(cc4_handle *h, uint64_t *out)
{
    if (!h || !out) return set_err(CC4_ERR_ARG, "cc4_launch_count: null%s", "");
    *out = h->launches;
    return CC4_OK;
}

extern "C" int cc4_error_flags(cc4_handle *h, uint32_t *out_host)
{
    if (!h || !out_host) return set_err(CC4_ERR_ARG, "cc4_error_flags: null%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy2D(out_host, sizeof(uint32_t), reinterpret_cast<const char *>(h->p.dyn) + offsetof(Cc4Dyn, err), sizeof(Cc4Dyn),
                          sizeof(uint32_t), (size_t)h->p.n_envs, cudaMemcpyDeviceToHost));
    return CC4_OK;
}

/* canonical export (layout: oracle/canon.py); decoding the packed record on the host */
extern "C" int cc4_export_canon(cc4_handle *h, int env, int32_t *out, int cap, int *n_out)
{
    if (!h || !out || !n_out || env < 0 || env >= h->p.n_envs) return set_err(CC4_ERR_ARG, "cc4_export_canon: bad argument%s", "");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    static thread_local Cc4Dyn d; static thread_local Cc4Stat st; static thread_local Cc4Cold cold;
    CUDA_TRY(cudaMemcpy(&d, h->p.dyn + env, sizeof d, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&st, h->p.stat + env, sizeof st, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&cold, h->p.cold + env, sizeof cold, cudaMemcpyDeviceToHost));
    int n = 0;
#define PUT(v) do { if (n < cap) out[n] = (int32_t)(v); n++; } while (0)
    PUT(d.phase); PUT(d.step);
    for (int s = 0; s < 9; s++) PUT(d.blocks[s]);
    int nh = 9;
    for (int sn = 0; sn < 8; sn++) nh += d.nuser[sn] + d.nserver[sn];
    PUT(nh);
    for (int slot = 0; slot < CC4_NSLOT; slot++) {
        if (slot < 128) {
            int sn = slot >> 4, k = slot & 15;
            if (!(k < 6 ? k < d.nserver[sn] : (k - 6) < d.nuser[sn])) continue;
            PUT(slot); PUT(d.ev[slot]);
            const uint32_t *pp = cold.parena + PD_OFF(cold.pdir[slot]);
            const int cnt = (int)PD_CNT(cold.pdir[slot]);
            int np = 0;
            for (int i = 0; i < cnt; i++) np += !P_DEAD(pp[i]);
            PUT(np);
            for (int i = 0; i < cnt; i++) { uint32_t w = pp[i]; if (P_DEAD(w)) continue; PUT(P_PID(w)); PUT(P_KIND(w)); PUT(P_ROOT(w)); }
            uint64_t v = cold.svc64[slot];
            int ns = 0; while (ns < 9 && (SVC_ENTRY(v, ns) & 15u) != 15u) ns++;
            PUT(ns);
            for (int i = 0; i < ns; i++) {
                uint32_t e = SVC_ENTRY(v, i), code = e & 15u, pid = 0;
                for (int j = 0; j < cnt; j++) if (P_KIND(pp[j]) == code && P_SVC(pp[j])) pid = P_PID(pp[j]);
                PUT(code); PUT(pid); PUT(!(code == K_OT && (v & SVC_OT_STOPPED))); PUT((e >> 4) * 20);
            }
        } else {
            PUT(slot); PUT(d.ev[slot]);
            int pid = slot < CC4_SLOT_INTERNET ? st.router_pid[slot - CC4_SLOT_ROUTER] : 0;
            if (pid) { PUT(1); PUT(pid); PUT(K_SESSION); PUT(0); } else PUT(0);
            PUT(0);
        }
    }
    for (int r = 0; r < CC4_NRED; r++) {
        int ns = 0;
        for (int i = 0; i < d.n_sess; i++) ns += SF_AGENT(d.sessf[i]) == (uint32_t)r;
        PUT(d.red_active >> r & 1); PUT(d.sess_count[r]); PUT(ns);
        for (int i = 0; i < d.n_sess; i++) if (SF_AGENT(d.sessf[i]) == (uint32_t)r) { PUT(S_ID(d.sess[i])); PUT(S_HOST(d.sess[i])); PUT(SF_ABS(d.sessf[i])); PUT(SF_ROOT(d.sessf[i])); PUT(S_PID(d.sess[i])); }
        int cnt = 0; for (int s = 0; s < 128; s++) cnt += (d.ip_known[r][s >> 5] >> (s & 31)) & 1;
        PUT(cnt); for (int s = 0; s < 128; s++) if ((d.ip_known[r][s >> 5] >> (s & 31)) & 1) PUT(s);
        cnt = 0; for (int s = 0; s < 128; s++) cnt += (d.hn_known[r][s >> 5] >> (s & 31)) & 1;
        PUT(cnt); for (int s = 0; s < 128; s++) if ((d.hn_known[r][s >> 5] >> (s & 31)) & 1) PUT(s);
        PUT(d.sn_known[r]);
        PUT(d.n_ss[r]); for (int i = 0; i < d.n_ss[r]; i++) PUT(d.ss[r][i]);
        const int nfsm = h->p.red_variant == CC4_RED_RANDOM_SELECT ? 0 : d.n_fsm[r];      /* RandomSelectRedAgent keeps no host states */
        PUT(nfsm);
        for (int i = 0; i < nfsm; i++) {
            int slot = d.fsm_order[r][i], w = slot >> 5, b = slot & 31, stt = 0;
            for (int pl = 0; pl < 4; pl++) stt |= (int)(d.fsm[r].pl[pl][w] >> b & 1u) << pl;
            PUT(slot); PUT(stt); PUT(d.fsm[r].hn[w] >> b & 1u);
        }
        /* ports sorted by (ident, slot) */
        int np = 0;
        for (int i = 0; i < d.n_ports; i++) np += PT_AGENT(cold.ports[i]) == (uint32_t)r;
        PUT(np);
        for (int id = 0; id < 256; id++) for (int s = 0; s < 128; s++) for (int i = 0; i < d.n_ports; i++) {
            uint32_t w = cold.ports[i];
            if (PT_AGENT(w) == (uint32_t)r && PT_ID(w) == (uint32_t)id && PT_HOST(w) == (uint32_t)s) { PUT(id); PUT(s); PUT(PT_MASK(w)); }
        }
        uint32_t ip = d.inprog[A_RED0 + r];
        if (IP_CODE(ip) == IP_NONE) { PUT(-1); PUT(0); } else { PUT(IP_CODE(ip)); PUT(IP_TICKS(ip)); }
    }
    for (int b = 0; b < CC4_NBLUE; b++) {
        int cnt = 0;
        for (int i = 0; i < d.n_sus; i++) cnt += SU_AGENT(cold.sus[i]) == (uint32_t)b;
        PUT(cnt);
        for (int s = 0; s < CC4_NSLOT; s++) for (int i = 0; i < d.n_sus; i++) if (SU_AGENT(cold.sus[i]) == (uint32_t)b && SU_SLOT(cold.sus[i]) == (uint32_t)s) { PUT(s); PUT(SU_PID(cold.sus[i])); }
        uint32_t ip = d.inprog[b];
        if (IP_CODE(ip) == IP_NONE) { PUT(-1); PUT(0); } else { PUT(IP_CODE(ip)); PUT(IP_TICKS(ip)); }
    }
#undef PUT
    *n_out = n;
    return CC4_OK;
}
