// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler behind cuda_emu.h.
#include "cuda_emu.h"

#include <ucontext.h>

#include <vector>

uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

namespace dcmt_emu {
namespace {
constexpr size_t kStack = 96 * 1024;
constexpr size_t kSmem = 256 * 1024;
alignas(128) unsigned char g_smem[kSmem];
ucontext_t g_main;
std::vector<ucontext_t> g_ctx;
std::vector<unsigned char> g_stacks;
std::vector<char> g_done;
const std::function<void()>* g_body = nullptr;
int g_cur = 0, g_nthreads = 0, g_alive = 0;
long g_launches = 0;
// block barrier state
int g_bar_count = 0, g_bar_gen = 0, g_bar_acc = 0, g_bar_result = 0;
// warp state
struct Warp { int count = 0, gen = 0, lanes = 32; unsigned slot[32]; unsigned ballot_acc = 0, ballot_res = 0, sum_acc = 0, sum_res = 0; };
std::vector<Warp> g_warps;

void yield() { swapcontext(&g_ctx[g_cur], &g_main); }

void trampoline() {
    (*g_body)();
    g_done[g_cur] = 1;
    --g_alive;
    swapcontext(&g_ctx[g_cur], &g_main);
}
int alive_threads() { return g_alive; }
}  // namespace

void* dyn_smem() { return g_smem; }
long launches() { return g_launches; }

void block_barrier() { (void)block_count(0); }

int block_count(int pred) {
    // CUDA semantics: a barrier among the threads that have not exited.
    const int my_gen = g_bar_gen;
    g_bar_acc += pred ? 1 : 0;
    ++g_bar_count;
    while (g_bar_gen == my_gen) {
        if (g_bar_count >= alive_threads()) {  // last arriver (or others exited meanwhile)
            g_bar_result = g_bar_acc;
            g_bar_acc = 0;
            g_bar_count = 0;
            ++g_bar_gen;
            break;
        }
        yield();
    }
    return g_bar_result;
}

static void warp_sync(Warp& w) {
    const int my_gen = w.gen;
    if (++w.count >= w.lanes) {
        w.count = 0;
        ++w.gen;
        return;
    }
    while (w.gen == my_gen) yield();
}

void warp_barrier() { warp_sync(g_warps[g_cur / 32]); }

unsigned warp_exchange(unsigned v, int width, int mode, int arg) {
    Warp& w = g_warps[g_cur / 32];
    const int lane = g_cur % 32;
    if (width <= 0 || width > 32) width = 32;
    const int seg = lane / width * width;  // shuffles stay inside segments of `width` lanes
    w.slot[lane] = v;
    warp_sync(w);
    int src = lane;
    switch (mode) {
        case 0: src = seg + (arg & (width - 1)); break;
        case 1: src = lane + arg; break;
        case 2: src = lane - arg; break;
        case 3: src = lane ^ arg; break;
    }
    if (src < seg || src >= seg + width) src = lane;
    const unsigned r = (src >= 0 && src < w.lanes) ? w.slot[src] : v;
    warp_sync(w);
    return r;
}

unsigned warp_ballot(int pred) {
    Warp& w = g_warps[g_cur / 32];
    const int lane = g_cur % 32;
    const int my_gen = w.gen;
    if (pred) w.ballot_acc |= 1u << lane;
    if (++w.count >= w.lanes) {
        w.ballot_res = w.ballot_acc;
        w.ballot_acc = 0;
        w.count = 0;
        ++w.gen;
    } else {
        while (w.gen == my_gen) yield();
    }
    const unsigned r = w.ballot_res;
    warp_sync(w);  // nobody overwrites ballot_res before everyone has read it
    return r;
}

unsigned warp_reduce_add(unsigned v) {
    Warp& w = g_warps[g_cur / 32];
    const int my_gen = w.gen;
    w.sum_acc += v;
    if (++w.count >= w.lanes) {
        w.sum_res = w.sum_acc;
        w.sum_acc = 0;
        w.count = 0;
        ++w.gen;
    } else {
        while (w.gen == my_gen) yield();
    }
    const unsigned r = w.sum_res;
    warp_sync(w);  // nobody overwrites sum_res before everyone has read it
    return r;
}

int warp_gather(unsigned v, unsigned* out) {  // every lane's value, in lane order; returns the number of lanes of the warp
    Warp& w = g_warps[g_cur / 32];
    w.slot[g_cur % 32] = v;
    warp_sync(w);
    const int n = w.lanes;
    for (int i = 0; i < n; ++i) out[i] = w.slot[i];
    warp_sync(w);  // nobody overwrites a slot before everyone has read it
    return n;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    ++g_launches;
    if (smem_bytes > kSmem) std::abort();
    const int nthreads = (int)(block.x * block.y * block.z);
    if ((int)g_ctx.size() < nthreads) {
        g_ctx.resize(nthreads);
        g_stacks.resize((size_t)nthreads * kStack);
    }
    g_done.assign(nthreads, 0);
    g_nthreads = nthreads;
    g_body = &body;
    gridDim = grid;
    blockDim = block;
    // DCMT_EMU_ORDER=reverse runs the blocks of a grid and the threads of a block in reverse order: results that
    // depend on block order (two tiles writing one cell) or on thread order (a missing barrier) then change.
    static const bool reverse = [] { const char* e = std::getenv("DCMT_EMU_ORDER"); return e && e[0] == 'r'; }();
    const unsigned nblocks = grid.x * grid.y * grid.z;
    for (unsigned bi = 0; bi < nblocks; ++bi) {
            {
                const unsigned b = reverse ? nblocks - 1 - bi : bi;
                const unsigned bx = b % grid.x, by = (b / grid.x) % grid.y, bz = b / (grid.x * grid.y);
                blockIdx = uint3{bx, by, bz};
                std::memset(g_smem, 0xCD, smem_bytes);  // poison: uninitialised shared memory reads show up
                g_bar_count = g_bar_gen = g_bar_acc = g_bar_result = 0;
                g_warps.assign((nthreads + 31) / 32, Warp());
                for (size_t w = 0; w < g_warps.size(); ++w)
                    g_warps[w].lanes = std::min(32, nthreads - (int)w * 32);
                for (int t = 0; t < nthreads; ++t) {
                    g_done[t] = 0;
                    getcontext(&g_ctx[t]);
                    g_ctx[t].uc_stack.ss_sp = g_stacks.data() + (size_t)t * kStack;
                    g_ctx[t].uc_stack.ss_size = kStack;
                    g_ctx[t].uc_link = &g_main;
                    makecontext(&g_ctx[t], trampoline, 0);
                }
                int alive = nthreads;
                g_alive = nthreads;
                while (alive) {
                    alive = 0;
                    for (int tt = 0; tt < nthreads; ++tt) {
                        const int t = reverse ? nthreads - 1 - tt : tt;
                        if (g_done[t]) continue;
                        g_cur = t;
                        threadIdx = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y,
                                          (unsigned)t / (block.x * block.y)};
                        swapcontext(&g_main, &g_ctx[t]);
                        alive += !g_done[t];
                    }
                }
            }
    }
    g_body = nullptr;
}
}  // namespace dcmt_emu
