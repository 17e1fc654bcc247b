// tools/simt/simt.cpp -- fiber scheduler and runtime stubs of the SIMT emulator (see include/simt.h; test infrastructure only).
//
// One OS thread runs one CTA at a time; every CUDA thread of the CTA is a fiber (own stack, hand-written x86-64 context
// switch).  Scheduling is cooperative and two-level: a lane that waits on a WARP collective hands the core to the next
// live lane of its warp; a lane that spins on a shared flag, or whose whole warp sits in a CTA barrier, hands it to the
// next live warp.  Exited lanes count as arrived in every collective and barrier.
#include "simt.h"

#include <sys/mman.h>
#include <stdio.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "the SIMT emulator's context switch is written for x86-64"
#endif

extern "C" void simt_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
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
.size simt_switch,.-simt_switch
)");

namespace {

const size_t kStack = 512 * 1024;

struct Cta;
struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  SimtThreadView view;
  int lane = 0, warp = 0;
  bool done = true;
  uint32_t coll_seq = 0;
  Cta* cta = nullptr;
};
struct WarpS {
  uint64_t vals[2][32];
  uint32_t seq[32];
  uint32_t live = 0;
  int cur = 0;
  int idle = 0;       // consecutive fruitless wake-ups of lanes blocked at CTA level (barriers); reset on any progress
};
struct NamedBar { int arrived = 0; unsigned gen = 0; };
struct Cta {
  std::vector<Fiber> fibers;
  std::vector<WarpS> warps;
  int nthreads = 0, nwarps = 0, live_threads = 0, curwarp = 0;
  int bar_arrived = 0; unsigned bar_gen = 0;
  NamedBar named[16];
  unsigned char* smem = nullptr;
  size_t smem_cap = 0;
  const std::function<void()>* body = nullptr;
  Fiber main;      // the worker thread's own context
};

thread_local Fiber* t_self = nullptr;
thread_local Cta* t_cta = nullptr;

void switch_to(Fiber* next) {
  Fiber* prev = t_self;
  if (next == prev) return;
  t_self = next;
  simt_switch(&prev->sp, next->sp);
}

// next live lane of warp w after `from` (cyclic); -1 if none
int next_lane(const WarpS& w, int from) {
  if (!w.live) return -1;
  for (int k = 1; k <= 32; ++k) { const int l = (from + k) & 31; if ((w.live >> l) & 1u) return l; }
  return -1;
}
int next_warp(Cta* c, int from) {
  for (int k = 1; k <= c->nwarps; ++k) { const int w = (from + k) % c->nwarps; if (c->warps[w].live) return w; }
  return -1;
}
void yield_warp() {
  Fiber* f = t_self; Cta* c = f->cta; WarpS& w = c->warps[f->warp];
  const int l = next_lane(w, f->lane);
  if (l < 0 || l == f->lane) return;
  w.cur = l;
  switch_to(&c->fibers[f->warp * 32 + l]);
}
void yield_cta() {
  Fiber* f = t_self; Cta* c = f->cta;
  c->warps[f->warp].cur = f->lane;
  const int w = next_warp(c, f->warp);
  if (w < 0) return;
  if (w == f->warp) { yield_warp(); return; }
  c->curwarp = w;
  switch_to(&c->fibers[w * 32 + c->warps[w].cur]);
}

void fiber_exit() __attribute__((noreturn));
void fiber_exit() {
  Fiber* f = t_self; Cta* c = f->cta; WarpS& w = c->warps[f->warp];
  f->done = true;
  w.live &= ~(1u << f->lane);
  --c->live_threads;
  if (c->bar_arrived > 0 && c->bar_arrived >= c->live_threads) { c->bar_arrived = 0; ++c->bar_gen; }
  Fiber* next = nullptr;
  const int l = next_lane(w, f->lane);
  if (l >= 0) { w.cur = l; next = &c->fibers[f->warp * 32 + l]; }
  else {
    const int nw = next_warp(c, f->warp);
    if (nw >= 0) { c->curwarp = nw; next = &c->fibers[nw * 32 + c->warps[nw].cur]; }
    else next = &c->main;
  }
  t_self = next;
  void* dummy;
  simt_switch(&dummy, next->sp);
  __builtin_unreachable();
}

void fiber_entry() {
  (*t_self->cta->body)();
  fiber_exit();
}

void prepare(Fiber* f) {
  if (!f->stack) {
    void* m = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { perror("simt: mmap stack"); abort(); }
    f->stack = (char*)m;
  }
  // initial frame: six callee-saved registers, then the return address (fiber_entry); rsp % 16 == 8 at its entry
  uintptr_t top = ((uintptr_t)f->stack + kStack) & ~(uintptr_t)15;
  void** sp = (void**)(top - 8);
  *--sp = (void*)fiber_entry;
  for (int i = 0; i < 6; ++i) *--sp = nullptr;
  f->sp = sp;
  f->done = false;
  f->coll_seq = 0;
}

void run_cta(Cta* c, dim3 grid, dim3 block, unsigned bx, unsigned by, unsigned bz, size_t smem, const std::function<void()>& body) {
  const int nt = (int)(block.x * block.y * block.z);
  c->nthreads = nt; c->nwarps = (nt + 31) / 32; c->live_threads = nt; c->curwarp = 0;
  c->bar_arrived = 0; c->bar_gen = 0;
  for (auto& nb : c->named) { nb.arrived = 0; nb.gen = 0; }
  if ((int)c->fibers.size() < c->nwarps * 32) c->fibers.resize(c->nwarps * 32);
  if ((int)c->warps.size() < c->nwarps) c->warps.resize(c->nwarps);
  if (c->smem_cap < smem + 64) {
    free(c->smem);
    c->smem_cap = smem + 64;
    c->smem = (unsigned char*)aligned_alloc(128, (c->smem_cap + 127) & ~(size_t)127);
  }
  memset(c->smem, 0xA5, c->smem_cap);
  c->body = &body;
  for (int w = 0; w < c->nwarps; ++w) {
    WarpS& W = c->warps[w];
    W.live = 0; W.cur = 0; W.idle = 0;
    memset(W.seq, 0, sizeof W.seq);
  }
  for (int t = 0; t < nt; ++t) {
    Fiber* f = &c->fibers[t];
    f->cta = c; f->lane = t & 31; f->warp = t >> 5;
    f->view.tid.x = t % block.x; f->view.tid.y = (t / block.x) % block.y; f->view.tid.z = t / (block.x * block.y);
    f->view.bid.x = bx; f->view.bid.y = by; f->view.bid.z = bz;
    f->view.bdim = block; f->view.gdim = grid;
    prepare(f);
    c->warps[f->warp].live |= 1u << f->lane;
  }
  c->main.cta = c;
  t_self = &c->main;
  t_cta = c;
  switch_to(&c->fibers[0]);
  t_self = nullptr;
  if (c->live_threads != 0) { fprintf(stderr, "simt: CTA ended with %d live threads (deadlock?)\n", c->live_threads); abort(); }
}

thread_local Cta* t_pool = nullptr;
std::mutex g_reg_mu;
std::vector<Cta*> g_reg;          // every worker's CTA slot, for the watchdog dump

void watchdog_dump() {
  std::lock_guard<std::mutex> lk(g_reg_mu);
  fprintf(stderr, "simt watchdog: %zu worker slots\n", g_reg.size());
  for (Cta* c : g_reg) {
    if (!c->live_threads) continue;
    const SimtThreadView& v = c->fibers[0].view;
    fprintf(stderr, " CTA (%u,%u,%u) of (%u,%u,%u): %d/%d threads live, barrier arrived %d gen %u, current warp %d\n", v.bid.x, v.bid.y, v.bid.z,
            v.gdim.x, v.gdim.y, v.gdim.z, c->live_threads, c->nthreads, c->bar_arrived, c->bar_gen, c->curwarp);
    for (int w = 0; w < c->nwarps; ++w) {
      const WarpS& W = c->warps[w];
      if (!W.live) continue;
      uint32_t lo = ~0u, hi = 0;
      for (int l = 0; l < 32; ++l) if ((W.live >> l) & 1u) { lo = std::min(lo, W.seq[l]); hi = std::max(hi, W.seq[l]); }
      fprintf(stderr, "  warp %2d live %08x cur %2d idle %2d collective seq %u..%u\n", w, W.live, W.cur, W.idle, lo, hi);
    }
    for (int i = 0; i < 16; ++i) if (c->named[i].arrived) fprintf(stderr, "  named barrier %d: %d arrived\n", i, c->named[i].arrived);
  }
}

std::atomic<size_t> g_alloc{0};
int g_threads = 0;

}  // namespace

extern "C" {

const SimtThreadView* simt_view() { return &t_self->view; }
unsigned char* simt_dyn_smem() { return t_self->cta->smem; }
int simt_lane() { return t_self->lane; }

const uint64_t* simt_warp_exchange(unsigned mask, uint64_t v, unsigned* took_part) {
  Fiber* f = t_self; WarpS& w = f->cta->warps[f->warp];
  const uint32_t k = ++f->coll_seq;
  w.vals[k & 1][f->lane] = v;
  w.seq[f->lane] = k;
  for (;;) {
    const unsigned need = mask & w.live;
    bool ok = true;
    for (unsigned m = need; m; m &= m - 1) {
      const int l = __builtin_ctz(m);
      if ((int32_t)(w.seq[l] - k) < 0) { ok = false; break; }
    }
    if (ok) {
      // lanes that deposited a value for this collective (a lane may have exited since: its value stays valid)
      unsigned tp = 0;
      for (unsigned m = mask; m; m &= m - 1) { const int l = __builtin_ctz(m); if ((int32_t)(w.seq[l] - k) >= 0) tp |= 1u << l; }
      *took_part = tp;
      w.idle = 0;
      return w.vals[k & 1];
    }
    yield_warp();
  }
}

static void bar_wait(Fiber* f, WarpS& w, const unsigned* gen, unsigned my_gen) {
  // give every lane of the warp a turn first (they may be released from an earlier barrier or still on their way here);
  // after a full fruitless rotation hand the core to the next warp
  while (*(volatile const unsigned*)gen == my_gen) {
    if (++w.idle >= __builtin_popcount(w.live)) { w.idle = 0; yield_cta(); }
    else yield_warp();
  }
  w.idle = 0;
  (void)f;
}

void simt_syncthreads() {
  Fiber* f = t_self; Cta* c = f->cta; WarpS& w = c->warps[f->warp];
  const unsigned g = c->bar_gen;
  if (++c->bar_arrived >= c->live_threads) { c->bar_arrived = 0; ++c->bar_gen; return; }
  bar_wait(f, w, &c->bar_gen, g);
}

void simt_named_barrier(int id, int count) {
  Fiber* f = t_self; Cta* c = f->cta; WarpS& w = c->warps[f->warp];
  NamedBar& b = c->named[id & 15];
  const unsigned g = b.gen;
  if (++b.arrived >= count) { b.arrived = 0; ++b.gen; return; }
  bar_wait(f, w, &b.gen, g);
}

// bar.arrive: counts the thread in, never waits (the producer side of a producer / consumer barrier)
void simt_named_arrive(int id, int count) {
  Fiber* f = t_self; Cta* c = f->cta;
  NamedBar& b = c->named[id & 15];
  if (++b.arrived >= count) { b.arrived = 0; ++b.gen; }
}

void simt_spin() {
  // one fruitless look at the flag: let the warp's other lanes run (one of them may be the flag's writer in another
  // role, or may still have to publish something), then the other warps
  Fiber* f = t_self; WarpS& w = f->cta->warps[f->warp];
  if (++w.idle >= __builtin_popcount(w.live)) { w.idle = 0; yield_cta(); }
  else yield_warp();
}

void simt_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const size_t total = (size_t)grid.x * grid.y * grid.z;
  if (!total) return;
  if (!g_threads) {
    const char* e = getenv("SIMT_THREADS");
    g_threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    if (g_threads < 1) g_threads = 1;
  }
  static bool dog = false;
  if (!dog) {
    dog = true;
    if (const char* e = getenv("SIMT_WATCHDOG")) {
      const int sec = atoi(e);
      std::thread([sec]() { std::this_thread::sleep_for(std::chrono::seconds(sec)); watchdog_dump(); fflush(stderr); _exit(3); }).detach();
    }
  }
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    if (!t_pool) { t_pool = new Cta(); std::lock_guard<std::mutex> lk(g_reg_mu); g_reg.push_back(t_pool); }
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= total) break;
      const unsigned bx = (unsigned)(i % grid.x), by = (unsigned)((i / grid.x) % grid.y), bz = (unsigned)(i / ((size_t)grid.x * grid.y));
      run_cta(t_pool, grid, block, bx, by, bz, smem, body);
    }
  };
  const int nth = (int)std::min<size_t>((size_t)g_threads, total);
  if (nth <= 1) { worker(); return; }
  // persistent kernels spin on nothing across CTAs, so CTAs may run in any order and any number at a time
  std::vector<std::thread> th;
  for (int t = 0; t < nth; ++t) th.emplace_back(worker);
  for (auto& t : th) t.join();
}

// ------------------------------------------------------------------ runtime stubs
struct SimtStream { int dummy; };
struct SimtEvent { std::chrono::steady_clock::time_point t; };

static size_t total_mem() {
  const char* e = getenv("SIMT_MEM_GB");
  return (size_t)(e ? atoi(e) : 180) << 30;
}
struct AllocHdr { size_t n; size_t pad[31]; };   // 256 bytes in front of every allocation

cudaError_t cudaMalloc(void** p, size_t n) {
  void* m = mmap(nullptr, n + sizeof(AllocHdr), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (m == MAP_FAILED) { *p = nullptr; return cudaErrorMemoryAllocation; }
  ((AllocHdr*)m)->n = n;
  unsigned char* d = (unsigned char*)m + sizeof(AllocHdr);
  const char* pz = getenv("SIMT_POISON_MB");
  const size_t cap = (size_t)(pz ? atoi(pz) : 1024) << 20;
  memset(d, 0xA5, n < cap ? n : cap);      // real cudaMalloc memory is not zero: catch reads of unwritten tables
  g_alloc += n;
  *p = d;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  AllocHdr* h = (AllocHdr*)((unsigned char*)p - sizeof(AllocHdr));
  g_alloc -= h->n;
  munmap(h, h->n + sizeof(AllocHdr));
  return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot) { *tot = total_mem(); const size_t a = g_alloc.load(); *fr = a < *tot ? *tot - a : 0; return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) { *st = new SimtStream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t st) { delete st; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }      // streams are synchronous here
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new SimtEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof *p);
  snprintf(p->name, sizeof p->name, "SIMT emulator (B200 shape)");
  p->totalGlobalMem = total_mem(); p->sharedMemPerBlockOptin = 232448; p->sharedMemPerMultiprocessor = 233472;
  const char* e = getenv("SIMT_SMS");
  p->multiProcessorCount = e ? atoi(e) : 148; p->major = 10; p->minor = 0; p->l2CacheSize = 126 << 20; p->clockRate = 1965000;
  return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "simt emulator error"; }

}  // extern "C"

long long clock64() { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
