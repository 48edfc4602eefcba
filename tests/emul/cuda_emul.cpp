// tests/emul/cuda_emul.cpp — TEST INFRASTRUCTURE ONLY (see cuda_emul.hpp).
#include "cuda_emul.hpp"

namespace emul {
dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
uint64_t g_launches = 0;
int g_sync_or_acc[2] = {0, 0};
int g_sync_or_phase = 0;
uint32_t g_dyn_smem[64 * 1024];

static ucontext_t g_sched;
static std::vector<ucontext_t> g_ctx;
static std::vector<char*> g_stacks;
static std::vector<int> g_state;  // 0 ready, 1 at block barrier, 2 done, 3 waiting on warp op
static std::vector<uint32_t> g_wval, g_wres, g_warg;
static std::vector<int> g_wop;
static const std::function<void()>* g_body = nullptr;
static bool g_in_fiber = false;
static int g_cur = 0;
static const size_t STACK = 128 * 1024;

// EMUL_ORDER=reverse|random permutes the order in which the threads of a block (and the blocks of a grid) run
// between barriers: results must not depend on it (a cheap stand-in for the concurrency of a real GPU)
static int order_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("EMUL_ORDER");
        mode = !e ? 0 : (!strncmp(e, "reverse", 7) ? 1 : (!strncmp(e, "random", 6) ? 2 : 0));
        if (e && strstr(e, "_blocks")) mode |= 16;   // permute only the blocks
        if (e && strstr(e, "_threads")) mode |= 32;  // permute only the threads
    }
    return mode;
}
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return (uint32_t)(g_rng >> 32); }
static void make_order(std::vector<int>& ord, int n, bool blocks) {
    ord.resize((size_t)n);
    for (int i = 0; i < n; i++) ord[(size_t)i] = i;
    const int m = order_mode();
    if ((blocks && (m & 32)) || (!blocks && (m & 16))) return;
    if ((m & 3) == 1) std::reverse(ord.begin(), ord.end());
    else if ((m & 3) == 2) for (int i = n - 1; i > 0; i--) std::swap(ord[(size_t)i], ord[(size_t)(rnd() % (uint32_t)(i + 1))]);
}

static void fiber_entry() {
    (*g_body)();
    g_state[g_cur] = 2;
    swapcontext(&g_ctx[g_cur], &g_sched);
}

void sync_threads() {
    if (!g_in_fiber) { fprintf(stderr, "emul: __syncthreads in a BPE_LAUNCH_NS kernel\n"); abort(); }
    g_state[g_cur] = 1;
    swapcontext(&g_ctx[g_cur], &g_sched);
}

// a spinning thread lets the other threads of its block run (the fibers are cooperative; on a GPU they run anyway)
void yield_thread() {
    if (!g_in_fiber) return;
    swapcontext(&g_ctx[g_cur], &g_sched);  // state stays "ready": it is resumed in the next sweep
}

uint32_t warp_exchange(uint32_t v, int op, uint32_t arg) {
    if (!g_in_fiber) { fprintf(stderr, "emul: warp intrinsic in a BPE_LAUNCH_NS kernel\n"); abort(); }
    int me = g_cur;
    g_wval[me] = v; g_wop[me] = op; g_warg[me] = arg;
    g_state[me] = 3;
    swapcontext(&g_ctx[me], &g_sched);
    return g_wres[me];
}

static bool resolve_warps(int n) {
    bool any = false;
    for (int w0 = 0; w0 < n; w0 += 32) {
        int w1 = w0 + 32 < n ? w0 + 32 : n;
        bool all = true, some = false;
        for (int t = w0; t < w1; t++) {
            if (g_state[t] == 2) continue;
            if (g_state[t] == 3) some = true; else all = false;
        }
        if (!some || !all) continue;
        uint32_t ballot = 0;
        for (int t = w0; t < w1; t++) if (g_state[t] == 3 && g_wval[t]) ballot |= 1u << (t - w0);
        for (int t = w0; t < w1; t++) {
            if (g_state[t] != 3) continue;
            int lane = t - w0, src = lane;
            switch (g_wop[t]) {
                case EMUL_BALLOT: g_wres[t] = ballot; break;
                case EMUL_ANY: g_wres[t] = ballot ? 1u : 0u; break;
                case EMUL_SHFL: src = (int)(g_warg[t] & 31); break;
                case EMUL_SHFL_UP: src = lane - (int)g_warg[t]; if (src < 0) src = lane; break;
                case EMUL_SHFL_DOWN: src = lane + (int)g_warg[t]; if (src > 31) src = lane; break;
                case EMUL_SHFL_XOR: src = lane ^ (int)g_warg[t]; break;
            }
            if (g_wop[t] != EMUL_BALLOT && g_wop[t] != EMUL_ANY) {
                int st = w0 + src;
                g_wres[t] = (st < w1 && g_state[st] == 3) ? g_wval[st] : g_wval[t];
            }
        }
        for (int t = w0; t < w1; t++) if (g_state[t] == 3) g_state[t] = 0;
        any = true;
    }
    return any;
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body, bool uses_sync) {
    g_launches++;
    g_gridDim = grid;
    g_blockDim = block;
    const int n = (int)block.x;
    std::vector<int> border, torder;
    make_order(border, (int)grid.x, true);
    if (!uses_sync) {
        g_in_fiber = false;
        for (unsigned bi = 0; bi < grid.x; bi++) {
            g_blockIdx = dim3((unsigned)border[bi]);
            make_order(torder, n, false);
            for (int ti = 0; ti < n; ti++) {
                g_threadIdx = dim3((unsigned)torder[(size_t)ti]);
                body();
            }
        }
        return;
    }
    if ((int)g_stacks.size() < n) {
        size_t old = g_stacks.size();
        g_stacks.resize((size_t)n);
        for (size_t i = old; i < (size_t)n; i++) g_stacks[i] = (char*)malloc(STACK);
    }
    g_ctx.resize((size_t)n);
    g_state.assign((size_t)n, 0);
    g_wval.assign((size_t)n, 0); g_wres.assign((size_t)n, 0); g_warg.assign((size_t)n, 0); g_wop.assign((size_t)n, 0);
    g_body = &body;
    g_in_fiber = true;
    for (unsigned bi = 0; bi < grid.x; bi++) {
        const unsigned b = (unsigned)border[bi];
        g_blockIdx = dim3(b);
        for (int t = 0; t < n; t++) {
            getcontext(&g_ctx[t]);
            g_ctx[t].uc_stack.ss_sp = g_stacks[t];
            g_ctx[t].uc_stack.ss_size = STACK;
            g_ctx[t].uc_link = &g_sched;
            makecontext(&g_ctx[t], fiber_entry, 0);
            g_state[t] = 0;
        }
        g_sync_or_acc[0] = g_sync_or_acc[1] = 0;
        g_sync_or_phase = 0;
        int done = 0;
        while (done < n) {
            bool progressed = false;
            make_order(torder, n, false);
            for (int ti = 0; ti < n; ti++) {
                const int t = torder[(size_t)ti];
                if (g_state[t] != 0) continue;
                g_cur = t;
                g_threadIdx = dim3((unsigned)t);
                swapcontext(&g_sched, &g_ctx[t]);
                progressed = true;
            }
            done = 0;
            int at_bar = 0;
            for (int t = 0; t < n; t++) { if (g_state[t] == 2) done++; else if (g_state[t] == 1) at_bar++; }
            bool resolved = resolve_warps(n);
            if (at_bar > 0 && at_bar + done == n) {
                for (int t = 0; t < n; t++) if (g_state[t] == 1) g_state[t] = 0;
                resolved = true;
            }
            if (!progressed && !resolved && done < n) { fprintf(stderr, "emul: deadlock in block %u\n", b); abort(); }
        }
    }
    g_in_fiber = false;
}
}  // namespace emul
