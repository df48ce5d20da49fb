// postfilter.cu — §8f N3: processResults, the host-side step right after the hot path in
// ContextEngine.buildContext (src/lib/context/engine.ts:289 → src/lib/context/rag/dedup-filter.ts).
//
// String work over <= ~30 fused results, so it stays on the host; it is here so the
// native side can hand the engine the same list the reference would. Restated line by line on UTF-16
// code units, because every JS string operation involved (length, substring, split(''), includes,
// regex classes without the u flag) counts code units:
//   processResults        dedup-filter.ts:193-247   keyword-presence filter → noise filter → dedup → rerank
//   filterNoise           :106-127                  5 whole-string patterns + punctuation density > 0.3
//   deduplicateResults    :42-91                    char-set Jaccard >= 0.85 on the first 200 units, keep-first,
//                                                   merge sources / max score, slice(maxResults)
//   rerankByRelevance     :132-155                  0.7*fusionScore + 0.3*keyword coverage, stable sort desc
// Host code only.
#include "common.cuh"

#include <algorithm>
#include <set>
#include <string>
#include <vector>

namespace {

typedef std::u16string str;

// JS \s (WhiteSpace + LineTerminator) — also what String.prototype.trim removes
bool is_space(char16_t c) {
  return c == 0x09 || c == 0x0A || c == 0x0B || c == 0x0C || c == 0x0D || c == 0x20 || c == 0xA0 || c == 0x1680 ||
         (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000 || c == 0xFEFF;
}
bool in_set(char16_t c, const char16_t* set) {
  for (; *set; set++)
    if (*set == c) return true;
  return false;
}
const char16_t kQuerySplit[] = u"\uFF0C\u3002\uFF01\uFF1F\u3001";                            // ，。！？、  (:216)
const char16_t kPunct[] = u"\uFF0C\u3002\uFF01\uFF1F\u3001\uFF1B\uFF1A\"\"''\u3010\u3011\uFF08\uFF09";  // (:119, :162)
const char16_t kPurePunct[] = u".\u3002,\uFF0C;\uFF1B:\uFF1A!\uFF01?\uFF1F";                 // (:98)
const char16_t kSection[] = u"\u7AE0\u8282\u9875\u6761\u6B3E";                                // 章节页条款 (:100)

str trim(const str& s) {
  size_t a = 0, b = s.size();
  while (a < b && is_space(s[a])) a++;
  while (b > a && is_space(s[b - 1])) b--;
  return s.substr(a, b - a);
}
bool all_of_set(const str& s, bool (*pred)(char16_t)) {
  if (s.empty()) return false;  // the patterns use +
  for (char16_t c : s)
    if (!pred(c)) return false;
  return true;
}
bool is_digit(char16_t c) { return c >= u'0' && c <= u'9'; }

// NOISE_PATTERNS (:96-102) on the trimmed content
bool is_noise(const str& c) {
  if (all_of_set(c, is_space)) return true;                                     // /^[\s\n]+$/
  if (all_of_set(c, [](char16_t x) { return in_set(x, kPurePunct); })) return true;  // /^[.。,，;；:：!！?？]+$/
  if (all_of_set(c, is_digit)) return true;                                     // /^\d+$/
  {                                                                             // /^第?\d+[章节页条款]$/
    size_t i = 0;
    if (i < c.size() && c[i] == 0x7B2C) i++;
    size_t d = i;
    while (d < c.size() && is_digit(c[d])) d++;
    if (d > i && d + 1 == c.size() && in_set(c[d], kSection)) return true;
  }
  if (c == u"\u76EE\u5F55" || c == u"\u7D22\u5F15" || c == u"\u53C2\u8003\u6587\u732E") return true;  // 目录|索引|参考文献
  return false;
}

// extractKeywords (:160-165)
std::set<str> extract_keywords(const str& text) {
  std::set<str> out;
  str word;
  auto flush = [&]() {
    if (word.size() >= 2) out.insert(word);
    word.clear();
  };
  for (char16_t c : text) {
    if (in_set(c, kPunct) || is_space(c)) flush();  // replaced by ' ' then split(' ')
    else word.push_back(c);
  }
  flush();
  return out;
}

// calculateKeywordCoverage (:170-188)
double coverage(const std::set<str>& query, const std::set<str>& content) {
  if (query.empty()) return 0.0;
  int covered = 0;
  for (const str& kw : query)
    for (const str& w : content)
      if (w.find(kw) != str::npos || kw.find(w) != str::npos) { covered++; break; }
  return (double)covered / (double)query.size();
}

// calculateSimilarity (:26-37) on the first 200 code units of each text; NaN (both empty) compares false
bool similar(const str& a, const str& b, double threshold) {
  std::set<char16_t> s1(a.begin(), a.begin() + std::min<size_t>(200, a.size()));
  std::set<char16_t> s2(b.begin(), b.begin() + std::min<size_t>(200, b.size()));
  size_t inter = 0;
  for (char16_t c : s1) inter += s2.count(c);
  const size_t uni = s1.size() + s2.size() - inter;
  if (uni == 0) return false;
  return (double)inter / (double)uni >= threshold;
}

struct fused {
  uint32_t index;
  double fusion_score;
  bool deduplicated;
  uint32_t source_mask, n_sources;
};

}  // namespace

extern "C" int rag_process_results(const rag_text* contents, const double* scores, const uint8_t* sources, uint32_t n,
                                   rag_text query, const rag_process_opts* opts, rag_processed_out* out) {
  if ((n && (!contents || !scores)) || !out || (out->capacity && (!out->index || !out->fusion_score)))
    return rag_set_error(RAG_ERR_INVALID, "rag_process_results: null argument");
  rag_process_opts o = {0.85, 20, 10, 1, 1};  // DEFAULT_CONFIG (:16-20) + processResults defaults (:204-207)
  if (opts) o = *opts;
  const str q(reinterpret_cast<const char16_t*>(query.units), query.units ? query.len : 0);
  std::vector<str> text(n);
  for (uint32_t i = 0; i < n; i++)
    text[i].assign(reinterpret_cast<const char16_t*>(contents[i].units), contents[i].units ? contents[i].len : 0);

  // 0. keyword-presence filter (:216-231): query.split(/[\s，。！？、]+/).filter(w => w.length >= 2)
  std::vector<str> kws;
  {
    str w;
    for (char16_t c : q) {
      if (is_space(c) || in_set(c, kQuerySplit)) { if (w.size() >= 2) kws.push_back(w); w.clear(); }
      else w.push_back(c);
    }
    if (w.size() >= 2) kws.push_back(w);
  }
  std::vector<uint32_t> keep;
  for (uint32_t i = 0; i < n; i++) {
    bool hit = kws.empty();
    for (const str& kw : kws)
      if (text[i].find(kw) != str::npos) { hit = true; break; }
    if (hit) keep.push_back(i);
  }
  // 1. noise filter (:106-127)
  if (o.enable_noise_filter) {
    std::vector<uint32_t> k2;
    for (uint32_t i : keep) {
      const str c = trim(text[i]);
      if (is_noise(c)) continue;
      size_t punct = 0;
      for (char16_t x : c) punct += in_set(x, kPunct) ? 1 : 0;
      if (!c.empty() && (double)punct / (double)c.size() > 0.3) continue;
      k2.push_back(i);
    }
    keep.swap(k2);
  }
  // 2. dedup (:42-91)
  std::vector<fused> dd;
  for (uint32_t i : keep) {
    if (text[i].size() < o.min_content_length) continue;
    bool dup = false;
    fused* target = nullptr;
    for (fused& e : dd) {
      if (similar(text[i], text[e.index], o.similarity_threshold)) {
        dup = true;
        if (scores[i] > e.fusion_score) target = &e;
        break;
      }
    }
    const uint32_t bit = 1u << (sources ? (sources[i] & 31) : 0);
    if (!dup) dd.push_back(fused{i, scores[i], false, bit, 1});
    else if (target) {
      target->source_mask |= bit;
      target->n_sources++;
      target->fusion_score = std::max(target->fusion_score, scores[i]);
      target->deduplicated = true;
    }
  }
  if (dd.size() > o.max_results) dd.resize(o.max_results);
  // 3. rerank (:132-155)
  if (o.enable_rerank) {
    const std::set<str> qk = extract_keywords(q);
    for (fused& e : dd) {
      const double cov = coverage(qk, extract_keywords(text[e.index]));
      e.fusion_score = e.fusion_score * 0.7 + cov * 0.3;
    }
    std::stable_sort(dd.begin(), dd.end(), [](const fused& a, const fused& b) { return a.fusion_score > b.fusion_score; });
  }
  out->count = (uint32_t)dd.size();
  for (uint32_t i = 0; i < dd.size() && i < out->capacity; i++) {
    out->index[i] = dd[i].index;
    out->fusion_score[i] = dd[i].fusion_score;
    if (out->deduplicated) out->deduplicated[i] = dd[i].deduplicated ? 1 : 0;
    if (out->source_mask) out->source_mask[i] = dd[i].source_mask;
    if (out->n_sources) out->n_sources[i] = dd[i].n_sources;
  }
  if (dd.size() > out->capacity) return rag_set_error(RAG_ERR_INVALID, "rag_process_results: output capacity %u < %zu results", out->capacity, dd.size());
  return RAG_OK;
}
