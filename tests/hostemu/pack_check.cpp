// Host-emulation unit check: ref::pack_all_pairs must write exactly what ref::pack_all writes, for every table entry
// type (forward layout, transposed layout, block-entry combination, stem) and both storage types.
// Build: g++ -O1 -std=c++17 -fopenmp -DLBC_HOST_EMU -I learningbycheating_b200/csrc tests/hostemu/pack_check.cpp
#include <cstdio>
#include <vector>

#include "lbc_ref_ops.h"

namespace lbc {
long long g_launches = 0;   // normally defined by lbc_fast.cu
bool g_trace_on = false;
void trace_note(const char*) {}
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;
}  // namespace lbc
using namespace lbc;

template <class T>
static int run() {
  struct Conv { int Co, Ci, K; };
  const Conv convs[] = {{64, 3, 7}, {64, 64, 3}, {128, 64, 3}, {128, 64, 1}, {256, 128, 3}, {640, 256, 3}, {20, 64, 1}};
  std::vector<float> P;
  std::vector<ref::PackEntry> table;
  std::vector<int64_t> sizes;
  std::vector<int64_t> offs;
  for (const Conv& c : convs) {
    offs.push_back((int64_t)P.size());
    int64_t n = (int64_t)c.Co * c.Ci * c.K * c.K;
    for (int64_t i = 0; i < n; ++i) P.push_back((float)((i * 2654435761u) % 100003) / 977.f - 51.f);
  }
  auto add = [&](int type, int64_t src, int64_t src2, int Co, int Ci, int K, int aux, int64_t n) {
    ref::PackEntry e;
    e.src_off = src;
    e.src2_off = src2;
    e.dst = nullptr;
    e.type = type;
    e.Co = Co;
    e.Ci = Ci;
    e.K = K;
    e.aux = aux;
    e.n = n;
    table.push_back(e);
    sizes.push_back(n);
  };
  for (size_t i = 0; i < sizeof(convs) / sizeof(convs[0]); ++i) {
    const Conv& c = convs[i];
    add(0, offs[i], 0, c.Co, c.Ci, c.K, 0, (int64_t)c.Co * c.Ci * c.K * c.K);
    add(1, offs[i], 0, c.Co, c.Ci, c.K, 0, (int64_t)c.Co * c.Ci * c.K * c.K);
  }
  add(2, offs[2], offs[3], 128, 64, 3, 0, (int64_t)64 * 2 * 128);   // conv1 centre tap | 1x1 downsample
  add(3, offs[0], 0, 64, 3, 7, 192, (int64_t)64 * 192);             // stem rows padded to Kp = 192
  int bad = 0;
  std::vector<std::vector<T>> a(table.size()), b(table.size());
  std::vector<ref::PackEntry> ta = table, tb = table;
  for (size_t i = 0; i < table.size(); ++i) {
    a[i].assign((size_t)sizes[i], T());
    b[i].assign((size_t)sizes[i], T());
    memset(a[i].data(), 0xAB, sizeof(T) * a[i].size());
    memset(b[i].data(), 0xCD, sizeof(T) * b[i].size());
    ta[i].dst = a[i].data();
    tb[i].dst = b[i].data();
  }
  ref::pack_all<T>(nullptr, P.data(), ta.data(), (int)ta.size());
  ref::pack_all_pairs<T>(nullptr, P.data(), tb.data(), (int)tb.size());
  for (size_t i = 0; i < table.size(); ++i)
    if (memcmp(a[i].data(), b[i].data(), sizeof(T) * a[i].size()) != 0) {
      printf("entry %zu (type %d, Co %d Ci %d K %d) differs\n", i, table[i].type, table[i].Co, table[i].Ci, table[i].K);
      ++bad;
    }
  return bad;
}

int main() {
  int bad = run<float>() + run<bf16>();
  printf(bad ? "FAILED\n" : "pack_all_pairs == pack_all on %d entries x 2 dtypes\n", 16);
  return bad ? 1 : 0;
}
