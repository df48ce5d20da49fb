// postfilter_asan.cc — TEST HARNESS: processResults (postfilter.cu, host-only) compiled as plain C++ with ASAN + UBSAN and
// driven with random UTF-16 inputs (empty strings, lone surrogates, CJK punctuation, long runs, duplicates) and random
// options. Checks only what must hold for ANY input: no sanitizer report, count <= min(n, max_results), indices in range
// and unique. Run by tests/test_postfilter.py::test_process_results_under_sanitizers.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
static thread_local char g_err[1024];
int rag_set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); return code; }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }
static std::u16string rand_text() {
  static const char16_t pool[] = u" \t\n.。,，;；:：!！?？、【】（）\"'第1234567890章节页条款abcdefgXYZ检索增强生成记忆\xD83D\xDE00\xD800\xFEFF\x3000";
  const uint32_t kind = rnd() % 8;
  uint32_t len = kind == 0 ? 0 : kind == 1 ? rnd() % 4 : kind == 2 ? 150 + rnd() % 200 : kind == 3 ? 2000 + rnd() % 3000 : 10 + rnd() % 80;
  std::u16string s;
  for (uint32_t i = 0; i < len; i++) s.push_back(pool[rnd() % (sizeof(pool) / sizeof(pool[0]) - 1)]);
  return s;
}
int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 300;
  int failures = 0;
  for (int it = 0; it < iters; it++) {
    const uint32_t n = rnd() % 40;
    std::vector<std::u16string> texts(n);
    for (uint32_t i = 0; i < n; i++) texts[i] = (i > 0 && rnd() % 4 == 0) ? texts[rnd() % i] : rand_text();  // exact duplicates too
    std::vector<rag_text> t(n ? n : 1);
    std::vector<double> scores(n ? n : 1);
    std::vector<uint8_t> sources(n ? n : 1);
    for (uint32_t i = 0; i < n; i++) {
      t[i] = rag_text{reinterpret_cast<const uint16_t*>(texts[i].data()), (uint32_t)texts[i].size()};
      scores[i] = (rnd() % 1000) / 1000.0;
      sources[i] = (uint8_t)(rnd() % 4);
    }
    const std::u16string q = rand_text();
    rag_process_opts o = {(rnd() % 101) / 100.0, rnd() % 40, 1 + rnd() % 15, rnd() % 2, rnd() % 2};
    const uint32_t cap = n ? n : 1;
    std::vector<uint32_t> idx(cap), mask(cap), ns(cap);
    std::vector<double> fs(cap);
    std::vector<uint8_t> dd(cap);
    rag_processed_out out = {cap, idx.data(), fs.data(), dd.data(), mask.data(), ns.data(), 0};
    const int rc = rag_process_results(t.data(), scores.data(), sources.data(), n, rag_text{reinterpret_cast<const uint16_t*>(q.data()), (uint32_t)q.size()},
                                       (it % 5 == 0) ? nullptr : &o, &out);
    if (rc != RAG_OK) { printf("iteration %d: rc=%d %s\n", it, rc, g_err); failures++; continue; }
    const uint32_t max_results = (it % 5 == 0) ? 10u : o.max_results;
    bool ok = out.count <= n && out.count <= max_results;
    std::vector<char> seen(cap, 0);
    for (uint32_t i = 0; ok && i < out.count; i++) {
      ok = idx[i] < n && !seen[idx[i]];
      if (ok) seen[idx[i]] = 1;
    }
    if (!ok) { printf("iteration %d: invariant broken (n=%u count=%u)\n", it, n, out.count); failures++; }
  }
  printf("done: %d iterations, %d failures\n", iters, failures);
  return failures ? 1 : 0;
}
