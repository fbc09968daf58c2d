#include "profile.cuh"

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace amc {
namespace {
struct Rec {
  const char* name;
  cudaEvent_t a, b;
  double flops, bytes;
};
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
}  // namespace

bool profile_enabled() { return g_on; }

ProfScope::ProfScope(const char* name, cudaStream_t st_, double flops, double bytes) : slot(-1), st(st_) {
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  Rec r;
  r.name = name;
  r.flops = flops;
  r.bytes = bytes;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  slot = (int)g_recs.size();
  g_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[slot].b, st);
}
}  // namespace amc

using namespace amc;

extern "C" {
int amc_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  return 0;
}
// Synchronises the device, writes "name count total_ms flops bytes\n" per class into buf, clears the records.
int amc_profile_dump(char* buf, size_t cap) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("amc_profile_dump: %s", cudaGetErrorString(e));
    return (int)e;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  struct Agg { int n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    if (!agg.count(r.name)) order.push_back(r.name);
    Agg& a = agg[r.name];
    a.n++;
    a.ms += ms;
    a.flops += r.flops;
    a.bytes += r.bytes;
  }
  g_recs.clear();
  size_t off = 0;
  if (cap) buf[0] = 0;
  for (auto& k : order) {
    const Agg& a = agg[k];
    int w = snprintf(buf + off, off < cap ? cap - off : 0, "%s %d %.6f %.6e %.6e\n", k.c_str(), a.n, a.ms, a.flops,
                     a.bytes);
    if (w < 0 || off + (size_t)w >= cap) break;
    off += (size_t)w;
  }
  return 0;
}
}
