// fdr_launch.cuh -- launcher of k_fdr<Cfg, MODE>, shared by the translation units that instantiate it
// (fdr_stage.cu: MODE 0 / 1, fdr_stage_mono.cu: MODE 2 / 3 -- the kernel has 80 instantiations, compiled side by side).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "fdr_stage.cuh"

namespace b2 {

// variant number (FdrInfo::variant) -> ratio family
#define B2_FDR_VARIANTS(X) \
  X(1, FdrCfg160_147) \
  X(2, FdrCfg147_160) \
  X(3, FdrCfg2_1) \
  X(4, FdrCfg1_2) \
  X(5, FdrCfg2_1w) \
  X(6, FdrCfg3_1) \
  X(7, FdrCfg1_3) \
  X(8, FdrCfg441_160) \
  X(9, FdrCfg3_2) \
  X(10, FdrCfg2_3) \
  X(11, FdrCfg6_1) \
  X(12, FdrCfg1_6) \
  X(13, FdrCfg4_1) \
  X(14, FdrCfg1_4) \
  X(15, FdrCfg441_320) \
  X(16, FdrCfg320_441) \
  X(17, FdrCfg147_80) \
  X(18, FdrCfg80_147) \
  X(19, FdrCfg320_147) \
  X(20, FdrCfg147_320) \
  X(21, FdrCfg160_441)

template <class Cfg, int MODE>
inline const char *launch_cfg_t(const FdrInfo &fi, FdrKParams P, long long n_items, cudaStream_t st, Workspace *ws) {
  auto kern = k_fdr<Cfg, MODE>;
  int dev = 0;
  cudaGetDevice(&dev);
  // per instantiation (one set of statics per Cfg and FAST) and per device, initialised once under a mutex:
  // concurrent callers (the reference's thread-pool tests) launch from many threads
  struct DevCfg { bool set = false; int per_sm = 1, sm_count = 148; };
  static std::mutex mu;
  static std::vector<DevCfg> cfgs;
  // B200SOXR_FDR_PAD_KB: extra dynamic shared memory per CTA, to study throughput against CTAs per SM
  static const size_t pad = [] { const char *v = std::getenv("B200SOXR_FDR_PAD_KB"); return v ? size_t(std::atol(v)) << 10 : size_t(0); }();
  const size_t smem_bytes = fi.smem_bytes + pad;
  DevCfg dc;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cfgs.size() <= size_t(dev)) cfgs.resize(size_t(dev) + 1);
    DevCfg &c = cfgs[size_t(dev)];
    if (!c.set) {
      B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, kern, Cfg::THREADS, smem_bytes) != cudaSuccess || c.per_sm < 1) {
        cudaGetLastError();
        c.per_sm = 1;
      }
      // Every resident CTA holds TM_COLS of the SM's 512 tensor-memory columns for its whole life.  The occupancy
      // calculator answers 1 CTA per SM for any kernel that allocates tensor memory (measured: grid 148 instead of
      // 740, profiles/r2q); the hardware co-schedules as many as registers, shared memory and columns allow, which
      // is what __launch_bounds__(THREADS, MIN_CTAS) and the buffer sizes were chosen for.
      if (Cfg::TM_OK && B2_FDR_TMEM != 0) {
        cudaFuncAttributes fa;
        size_t static_smem = 1056;
        if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) static_smem = fa.sharedSizeBytes;
        const int by_smem = int((228u << 10) / (smem_bytes + static_smem + 1024));   // 228 kB per SM, 1 kB reserved per CTA
        c.per_sm = std::max(c.per_sm, std::min(std::min(Cfg::MIN_CTAS, by_smem), 512 / Cfg::TM_COLS));
      }
      if (cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c.sm_count < 1)
        c.sm_count = 148;
      c.set = true;
    }
    dc = c;
  }
  const long long grid = std::min<long long>(n_items, (long long)dc.sm_count * dc.per_sm);
  // dynamic chunks when there is more than one item per CTA (about six chunks per CTA, at most 8 items each: the
  // TMA staging of the next block works inside a chunk); the ticket counter lives in the caller's workspace
  static const bool static_split = [] { const char *v = std::getenv("B200SOXR_FDR_STATIC_SPLIT"); return v && v[0] == '1'; }();
  P.counter = nullptr;
  P.ticket_base = P.n_chunks = P.n_big = 0;
  P.chunk = P.chunk_tail = 1;
  if (ws && !static_split && n_items > grid) {
    if (!ws->fdr_tickets.p) {
      if (const char *e = ws->fdr_tickets.reserve(sizeof(unsigned))) return e;
      B2_CUDA(cudaMemsetAsync(ws->fdr_tickets.p, 0, sizeof(unsigned), st));
      ws->fdr_ticket_base = 0;
    }
    auto env_long = [](const char *name, long dflt) { const char *v = std::getenv(name); return v ? std::atol(v) : dflt; };
    static const long chunk_env = env_long("B200SOXR_FDR_CHUNK", 0), tail_pct = env_long("B200SOXR_FDR_TAIL_PCT", 25),
                      tail_chunk = env_long("B200SOXR_FDR_TAIL_CHUNK", 2);
    P.chunk = chunk_env > 0 ? unsigned(chunk_env)
                            : unsigned(std::max<long long>(1, std::min<long long>(8, n_items / (grid * 4))));
    P.chunk_tail = unsigned(std::max<long>(1, std::min<long>(tail_chunk, long(P.chunk))));
    const long long tail_items = std::min<long long>(n_items, n_items * std::max(0L, std::min(100L, tail_pct)) / 100);
    P.n_big = unsigned((n_items - tail_items) / P.chunk);
    const long long rest = n_items - (long long)P.n_big * P.chunk;
    P.n_chunks = P.n_big + unsigned((rest + P.chunk_tail - 1) / P.chunk_tail);
    P.counter = ws->fdr_tickets.as<unsigned>();
    P.ticket_base = ws->fdr_ticket_base;
    ws->fdr_ticket_base += P.n_chunks + unsigned(grid);   // every CTA draws one ticket past the end
  }
  kern<<<unsigned(grid), Cfg::THREADS, smem_bytes, st>>>(P);
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  return nullptr;
}

// P.fast == MODE_B runs k_fdr<Cfg, MODE_B>, anything else k_fdr<Cfg, MODE_A>
template <int MODE_A, int MODE_B>
inline const char *fdr_launch_modes(const FdrInfo &fi, const FdrKParams &P, long long n_items, cudaStream_t st, Workspace *ws) {
  switch (fi.variant) {
#define B2_X(v, Cfg) \
    case v:          \
      return P.fast == MODE_B ? launch_cfg_t<Cfg, MODE_B>(fi, P, n_items, st, ws) : launch_cfg_t<Cfg, MODE_A>(fi, P, n_items, st, ws);
    B2_FDR_VARIANTS(B2_X)
#undef B2_X
  }
  return "internal: unknown fdr variant";
}

// fdr_stage_mono.cu: the block-paired mono instantiations (P.fast 2 / 3)
const char *fdr_launch_mono(const FdrInfo &fi, const FdrKParams &P, long long n_items, cudaStream_t st, Workspace *ws);

}  // namespace b2
