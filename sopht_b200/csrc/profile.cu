// Per-kernel CUDA-event timers (off by default). When enabled, every instrumented launch site records an
// event pair on the stream it launches on; sopht_profile_report() synchronises once and returns the totals
// per kernel label. bench.py uses this for the roofline of the dominant kernel (live, same stream, same
// timed region) - it is not a profiler replacement and costs two event records per launch while on.
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace sopht {

bool g_prof_on = false;

namespace {
struct Rec {
  const char* label;
  cudaEvent_t e0, e1;
};
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
std::string g_report;

cudaEvent_t take_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

ProfScope::ProfScope(const char* label, cudaStream_t st) : st_(st), idx_(-1) {
  if (!g_prof_on) return;
  Rec r{label, take_event(), take_event()};
  cudaEventRecord(r.e0, st);
  idx_ = (int)g_recs.size();
  g_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx_ >= 0) cudaEventRecord(g_recs[idx_].e1, st_);
}

}  // namespace sopht

using namespace sopht;

extern "C" {

/* host-layer ranges (collectives issued by the Python layer between the kernels): label must stay valid */
int sopht_profile_range_begin(const char* label, void* stream) {
  if (!g_prof_on) return -1;
  static std::vector<std::string>* names = new std::vector<std::string>();
  const char* stable = nullptr;
  for (auto& n : *names)
    if (n == label) stable = n.c_str();
  if (!stable) {
    names->reserve(256);
    names->push_back(label);
    stable = names->back().c_str();
  }
  Rec r{stable, take_event(), take_event()};
  cudaEventRecord(r.e0, as_stream(stream));
  g_recs.push_back(r);
  return (int)g_recs.size() - 1;
}

int sopht_profile_range_end(int index, void* stream) {
  if (index >= 0 && index < (int)g_recs.size()) cudaEventRecord(g_recs[index].e1, as_stream(stream));
  return SOPHT_OK;
}

int sopht_profile_enable(int on) {
  g_prof_on = on != 0;
  return SOPHT_OK;
}

const char* sopht_profile_report(void) {
  std::map<std::string, std::pair<int64_t, double>> agg;
  std::vector<std::string> order;
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto it = agg.find(r.label);
    if (it == agg.end()) {
      order.push_back(r.label);
      it = agg.emplace(r.label, std::make_pair<int64_t, double>(0, 0.0)).first;
    }
    it->second.first++;
    it->second.second += ms;
    g_pool.push_back(r.e0);
    g_pool.push_back(r.e1);
  }
  g_recs.clear();
  g_report = "{";
  bool first = true;
  for (auto& k : order) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", k.c_str(),
             (long long)agg[k].first, agg[k].second);
    g_report += buf;
    first = false;
  }
  g_report += "}";
  return g_report.c_str();
}

}  // extern "C"
