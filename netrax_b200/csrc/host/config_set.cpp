/*
 * config_set.cpp — reticulation-configuration algebra on (care, second) bit masks.
 * Behaviour (including the ORDER in which simplifyReticulationChoices merges and removes entries, which fixes
 * the canonical form later compared with ==) follows src/graph/ReticulationConfigSet.cpp of the reference.
 */
#include <algorithm>
#include <cmath>

#include "netrax_likelihood_api.hpp"

namespace netrax {

bool ReticulationConfigSet::operator==(const ReticulationConfigSet &o) const {  // ReticulationConfigSet.hpp:21-42
  if (max_reticulations != o.max_reticulations || configs.size() != o.configs.size()) return false;
  for (const ReticulationConfig &a : configs)
    if (std::find(o.configs.begin(), o.configs.end(), a) == o.configs.end()) return false;
  return true;
}

static inline bool compatible(const ReticulationConfig &l, const ReticulationConfig &r) {  // .cpp:9-25
  return ((l.second ^ r.second) & l.care & r.care) == 0;
}

static double configLogProb(const ReticulationConfig &c, const std::vector<double> &first, const std::vector<double> &second) {  // .cpp:71-88
  double lp = 0;
  for (uint32_t m = c.care; m; m &= m - 1) {  // ascending reticulation index, like the reference's loop
    const int i = __builtin_ctz(m);
    lp += ((c.second >> i) & 1) ? second[i] : first[i];
  }
  return lp;
}

double computeReticulationConfigLogProb(const ReticulationConfigSet &c, const std::vector<double> &first, const std::vector<double> &second) {  // .cpp:98-113
  if (c.configs.size() == 1) return configLogProb(c.configs[0], first, second);
  double prob = 0.0;  // probabilities of reticulation choices never leave double range (mpreal only in the reference)
  for (const ReticulationConfig &x : c.configs) prob += std::exp(configLogProb(x, first, second));
  return std::log(prob);
}

double computeReticulationConfigProb(const ReticulationConfigSet &c, const std::vector<double> &first, const std::vector<double> &second) {  // .cpp:115-130
  if (c.configs.size() == 1) return std::exp(configLogProb(c.configs[0], first, second));
  double prob = 0.0;
  for (const ReticulationConfig &x : c.configs) prob += std::exp(configLogProb(x, first, second));
  return prob;
}

bool reticulationConfigsCompatible(const ReticulationConfigSet &l, const ReticulationConfigSet &r) {  // .cpp:132-142
  for (const ReticulationConfig &a : l.configs)
    for (const ReticulationConfig &b : r.configs)
      if (compatible(a, b)) return true;
  return false;
}

void simplifyReticulationChoices(ReticulationConfigSet &res) {  // .cpp:151-215
  std::vector<ReticulationConfig> &v = res.configs;
  auto drop = [&](size_t k) { std::swap(v[k], v.back()); v.pop_back(); };
  for (;;) {
    bool shrunk = false;
    for (size_t i = 0; i < v.size() && !shrunk; ++i)  // duplicates first
      for (size_t j = i + 1; j < v.size(); ++j)
        if (v[i] == v[j]) { drop(j); shrunk = true; break; }
    if (shrunk) continue;
    for (size_t i = 0; i < v.size() && !shrunk; ++i)
      for (size_t j = 0; j < res.max_reticulations && !shrunk; ++j) {
        if (!((v[i].care >> j) & 1)) continue;
        ReticulationConfig query = v[i];
        query.second ^= (1u << j);  // the same configuration with reticulation j flipped
        for (size_t k = 0; k < v.size(); ++k) {
          if (k == i) continue;
          if (v[k] == query) {  // 0x and 1x -> -x
            v[i].care &= ~(1u << j);
            v[i].second &= ~(1u << j);
            drop(k);
            shrunk = true;
            break;
          }
          if (v[k] == v[i]) { drop(k); shrunk = true; break; }
        }
      }
    if (!shrunk) break;
  }
}

ReticulationConfigSet combineReticulationChoices(const ReticulationConfigSet &l, const ReticulationConfigSet &r) {  // .cpp:217-231
  ReticulationConfigSet res(l.max_reticulations);
  for (const ReticulationConfig &a : l.configs)
    for (const ReticulationConfig &b : r.configs)
      if (compatible(a, b)) res.configs.push_back(ReticulationConfig{a.care | b.care, (a.second & a.care) | (b.second & b.care)});
  simplifyReticulationChoices(res);
  return res;
}

std::string toString(const ReticulationConfigSet &c, size_t nret) {
  std::vector<std::string> rows;
  for (const ReticulationConfig &x : c.configs) {
    std::string s;
    for (size_t i = 0; i < nret; ++i) s += !((x.care >> i) & 1) ? '-' : (((x.second >> i) & 1) ? '1' : '0');
    rows.push_back(s);
  }
  std::sort(rows.begin(), rows.end());
  std::string out;
  for (size_t i = 0; i < rows.size(); ++i) { if (i) out += '|'; out += rows[i]; }
  return out;
}

}  // namespace netrax
