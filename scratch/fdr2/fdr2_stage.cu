// fdr2_stage.cu -- host side of the two-pass form of the 160:147 block (fdr2_stage.cuh).
#include "fdr2_stage.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace b2 {

// twiddle bases w_N^(g 2^i): forward [6][64] for N = 2560, then inverse [6][48] for N = 2352 (applied conjugated)
std::vector<float> fdr2_tables() {
  std::vector<float> t;
  auto root = [&](long long k, long long n) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)(k % n) / (long double)n;
    t.push_back(float(std::cos(a)));
    t.push_back(float(std::sin(a)));
  };
  for (int i = 0; i < 6; ++i)
    for (int g = 0; g < Fdr2::RB; ++g) root((long long)g << i, Fdr2::NIN);
  for (int i = 0; i < 6; ++i)
    for (int q = 0; q < Fdr2::RD; ++q) root((long long)q << i, Fdr2::NOUT);
  return t;
}

bool fdr2_enabled() {
  static const bool v = [] { const char *e = std::getenv("B200SOXR_FDR2"); return e && e[0] == '1'; }();
  return v;
}

const char *fdr2_launch(const FdrInfo &fi, FdrKParams P, long long n_items, cudaStream_t st, Workspace *ws) {
  if (!ws) return "internal: k_fdr2 needs a workspace";
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  const size_t smem = size_t(2) * Fdr2::SLOT_ELEMS * sizeof(float2);
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(k_fdr2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(k_fdr2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  });
  if (attr_err != cudaSuccess) return cuda_err(attr_err);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  if (!ws->fdr_tickets.p) {
    if (const char *e = ws->fdr_tickets.reserve(sizeof(unsigned))) return e;
    B2_CUDA(cudaMemsetAsync(ws->fdr_tickets.p, 0, sizeof(unsigned), st));
    ws->fdr_ticket_base = 0;
  }
  P.f2tab = fi.f2tab.p;
  P.n_chunks = unsigned((n_items + 1) / 2);            // a ticket = two consecutive (pair, block) items
  const long long grid = std::min<long long>(P.n_chunks, (long long)sms * 2);
  P.counter = ws->fdr_tickets.as<unsigned>();
  P.ticket_base = ws->fdr_ticket_base;
  ws->fdr_ticket_base += P.n_chunks + unsigned(grid);   // every CTA draws one ticket past the end
  k_fdr2<<<unsigned(grid), 128, smem, st>>>(P);
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  return nullptr;
}

}  // namespace b2
