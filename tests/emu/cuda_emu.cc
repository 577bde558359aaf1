// TEST INFRASTRUCTURE ONLY -- fiber scheduler of the SIMT emulator (see cuda_emu.h).
#include "cuda_emu.h"

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
namespace emu {
thread_local Block *g_blk = nullptr;
static const size_t STACK = 256 * 1024;

void yield() {
  Block *b = g_blk;
  int me = b->cur;
  swapcontext(&b->ctx[me], &b->main);
}

static void trampoline() {
  Block *b = g_blk;
  int me = b->cur;
  b->body();
  b->done[me] = 1;
  b->progress++;
  swapcontext(&b->ctx[me], &b->main);
}

void run_block(Block &b, dim3 bid, dim3 bdim, dim3 gdim) {
  g_blk = &b;
  blockIdx.x = bid.x; blockIdx.y = bid.y; blockIdx.z = bid.z;
  blockDim = bdim; gridDim = gdim;
  int n = b.nthreads;
  b.bar_count = 0; b.bar_gen = 0;
  for (auto &w : b.warps) { w.count = 0; w.gen = 0; }
  for (int t = 0; t < n; ++t) {
    b.done[t] = 0;
    getcontext(&b.ctx[t]);
    b.ctx[t].uc_stack.ss_sp = b.stacks[t];
    b.ctx[t].uc_stack.ss_size = STACK;
    b.ctx[t].uc_link = &b.main;
    makecontext(&b.ctx[t], (void (*)())trampoline, 0);
  }
  int alive = n;
  while (alive) {
    long before = b.progress;
    alive = 0;
    for (int t = 0; t < n; ++t) {
      if (b.done[t]) continue;
      b.cur = t;
      threadIdx.x = t % bdim.x; threadIdx.y = (t / bdim.x) % bdim.y; threadIdx.z = t / (bdim.x * bdim.y);
      swapcontext(&b.main, &b.ctx[t]);
      if (!b.done[t]) alive++;
    }
    if (alive && b.progress == before) {
      fprintf(stderr, "cuda_emu: DEADLOCK in block (%u,%u): %d threads stuck at a barrier "
              "(divergent __syncthreads / warp collective?)\n", bid.x, bid.y, alive);
      abort();
    }
  }
}

void launch(std::function<void()> body, dim3 grid, dim3 block, size_t smem) {
  size_t nblk = (size_t)grid.x * grid.y * grid.z;
  int nthr = block.x * block.y * block.z;
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  unsigned nw = (unsigned)std::min<size_t>(hw, nblk);
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    Block b;
    b.nthreads = nthr;
    b.ctx.resize(nthr); b.done.resize(nthr); b.stacks.resize(nthr);
    b.warps.resize((nthr + 31) / 32);
    for (int t = 0; t < nthr; ++t) b.stacks[t] = (char *)malloc(STACK);
    b.smem = (unsigned char *)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
    b.body = body;
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nblk) break;
      dim3 bid(i % grid.x, (i / grid.x) % grid.y, i / ((size_t)grid.x * grid.y));
      memset(b.smem, 0xCD, smem);   // poison: uninitialised shared memory shows up
      run_block(b, bid, block, grid);
    }
    for (int t = 0; t < nthr; ++t) free(b.stacks[t]);
    free(b.smem);
  };
  std::vector<std::thread> th;
  for (unsigned w = 0; w < nw; ++w) th.emplace_back(worker);
  for (auto &t : th) t.join();
}
}  // namespace emu
