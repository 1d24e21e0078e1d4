// TEST TOOL ONLY (see cuda_emu.h).  Fiber scheduler behind the CUDA-thread emulator.
#include "cuda_emu.h"
#include <stdexcept>

emu_uint3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
char* dyn_smem = nullptr;
long long launch_counter = 0;

namespace {
enum State { RUNNABLE, WAIT_BLOCK, WAIT_WARP, WAIT_NAMED, DONE };
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    State st = RUNNABLE;
    emu_uint3 tid;
    int lin = 0;
    int bar_id = 0;
};
constexpr size_t kStack = 512 * 1024;
std::vector<Fiber> fibers;
Fiber* cur = nullptr;
void* sched_sp = nullptr;
const std::function<void()>* body_fn = nullptr;
int named_cnt[16];     // arrivals at the named barriers (bar.sync / bar.arrive id, count)
double xbuf[64][32];  // per-warp exchange
int xpred[64][32];

extern "C" void emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

void yield_to_sched() { emu_switch(&cur->sp, sched_sp); }

void fiber_entry() {
    (*body_fn)();
    cur->st = DONE;
    yield_to_sched();
    std::abort();
}

void init_fiber(Fiber& f) {
    if (!f.stack) f.stack = (char*)std::malloc(kStack);
    // stack top, 16-byte aligned; layout consumed by emu_switch: 6 regs then return address
    uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address of fiber_entry (keeps alignment: entry sees rsp%16==8)
    *--sp = (void*)&fiber_entry;     // 'ret' target
    for (int i = 0; i < 6; i++) *--sp = nullptr;
    f.sp = sp;
    f.st = RUNNABLE;
}
}  // namespace

int lane_id() { return cur->lin & 31; }

static unsigned tmem_cells[128][512];
unsigned* tmem_cell(int lane, int col) {
    if (lane < 0 || lane >= 128 || col < 0 || col + 8 > 512) { std::fprintf(stderr, "emu: TMEM access out of range (lane %d column %d)\n", lane, col); std::abort(); }
    // tcgen05 32x32b: a warp reaches only the lane quarter 32 * (warp id % 4)
    if ((lane >> 5) != ((cur->lin >> 5) & 3)) { std::fprintf(stderr, "emu: TMEM lane %d is not in the quarter of warp %d\n", lane, cur->lin >> 5); std::abort(); }
    return &tmem_cells[lane][col];
}

void sync_block() { cur->st = WAIT_BLOCK; yield_to_sched(); }
void sync_warp() { cur->st = WAIT_WARP; yield_to_sched(); }

// named barriers: the barrier completes when `count` threads have arrived (bar_sync blocks, bar_arrive does not)
static void named_release(int id) {
    named_cnt[id] = 0;
    for (auto& f : fibers) if (f.st == WAIT_NAMED && f.bar_id == id) f.st = RUNNABLE;
}
void bar_arrive(int id, int count) {
    if (++named_cnt[id] >= count) named_release(id);
}
void bar_sync(int id, int count) {
    if (++named_cnt[id] >= count) { named_release(id); return; }
    cur->st = WAIT_NAMED; cur->bar_id = id;
    yield_to_sched();
}

double shfl_exchange(double v, int src_lane) {
    int w = cur->lin >> 5, l = cur->lin & 31;
    xbuf[w][l] = v;
    sync_warp();
    double r = xbuf[w][src_lane & 31];
    sync_warp();
    return r;
}
// mma.sync.aligned.m8n8k4.row.col.f64: lane (g = lane >> 2, t = lane & 3) supplies A[g][t], B[t][g] and owns C[g][2t], C[g][2t+1];
// the products are accumulated in the order k = 0..3 with fused multiply-adds.
double xbuf2[64][32];
void dmma884(double& c0, double& c1, double a, double b) {
    int w = cur->lin >> 5, l = cur->lin & 31;
    xbuf[w][l] = a; xbuf2[w][l] = b;
    sync_warp();
    const int g = l >> 2, t = l & 3;
    for (int k = 0; k < 4; k++) {
        const double ak = xbuf[w][g * 4 + k];
        c0 = std::fma(ak, xbuf2[w][(2 * t) * 4 + k], c0);
        c1 = std::fma(ak, xbuf2[w][(2 * t + 1) * 4 + k], c1);
    }
    sync_warp();
}
unsigned ballot(int pred) {
    int w = cur->lin >> 5, l = cur->lin & 31;
    xpred[w][l] = pred ? 1 : 0;
    sync_warp();
    unsigned m = 0;
    int nthreads = (int)fibers.size();
    for (int i = 0; i < 32; i++) {
        int lin = (w << 5) + i;
        if (lin < nthreads && fibers[lin].st != DONE && xpred[w][i]) m |= (1u << i);
    }
    sync_warp();
    return m;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    launch_counter++;
    int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads > 2048 || nthreads <= 0) throw std::runtime_error("emu: bad block size");
    if ((int)fibers.size() < nthreads) fibers.resize(nthreads);
    std::vector<char> smem(smem_bytes + 64);
    dyn_smem = (char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    body_fn = &body;
    gridDim = {grid.x, grid.y, grid.z};
    blockDim = {block.x, block.y, block.z};
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        std::memset(dyn_smem, 0xCD, smem_bytes);  // poison: uninitialised shared memory reads show up as garbage
        std::memset(tmem_cells, 0xCD, sizeof(tmem_cells));
        blockIdx = {bx, by, bz};
        for (int& c : named_cnt) c = 0;
        int lin = 0;
        for (unsigned tz = 0; tz < block.z; tz++)
        for (unsigned ty = 0; ty < block.y; ty++)
        for (unsigned tx = 0; tx < block.x; tx++, lin++) {
            fibers[lin].tid = {tx, ty, tz};
            fibers[lin].lin = lin;
            init_fiber(fibers[lin]);
        }
        int ndone = 0;
        while (ndone < nthreads) {
            bool progressed = false;
            for (int i = 0; i < nthreads; i++) {
                Fiber& f = fibers[i];
                if (f.st != RUNNABLE) continue;
                cur = &f;
                threadIdx = f.tid;
                emu_switch(&sched_sp, f.sp);
                progressed = true;
                if (f.st == DONE) ndone++;
            }
            // release warps whose live lanes all wait at a warp rendezvous
            int nwarps = (nthreads + 31) / 32;
            for (int w = 0; w < nwarps; w++) {
                int lo = w * 32, hi = lo + 32 < nthreads ? lo + 32 : nthreads;
                bool all = true, any = false;
                for (int i = lo; i < hi; i++) {
                    if (fibers[i].st == WAIT_WARP) any = true;
                    else if (fibers[i].st != DONE) all = false;
                }
                if (any && all) { for (int i = lo; i < hi; i++) if (fibers[i].st == WAIT_WARP) fibers[i].st = RUNNABLE; progressed = true; }
            }
            // release the block barrier when every live thread waits at it
            bool allb = true, anyb = false;
            for (int i = 0; i < nthreads; i++) {
                if (fibers[i].st == WAIT_BLOCK) anyb = true;
                else if (fibers[i].st != DONE) allb = false;
            }
            if (anyb && allb) { for (int i = 0; i < nthreads; i++) if (fibers[i].st == WAIT_BLOCK) fibers[i].st = RUNNABLE; progressed = true; }
            if (!progressed) {
                std::fprintf(stderr, "emu: DEADLOCK in block (%u,%u,%u): barrier not reached by all threads\n", bx, by, bz);
                for (int i = 0; i < nthreads; i++)
                    if (fibers[i].st != DONE && (i % 32 == 0)) std::fprintf(stderr, "  thread %d state %d named-barrier %d (arrivals %d)\n", i, (int)fibers[i].st, fibers[i].bar_id, named_cnt[fibers[i].bar_id & 15]);
                std::abort();
            }
        }
    }
    dyn_smem = nullptr;
}
}  // namespace emu
