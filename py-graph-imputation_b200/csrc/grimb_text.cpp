// grimb_text.cpp -- host-side text pipeline of libgrimb200.so (see include/grimb200.h, "Text
// pipeline").  Multi-threaded C++: line split, GL-string tokenising, race -> prior matrices,
// and the six output texts with Python's str(float) layout.  Behaviour mirrors, line for line,
// grim/imputation/impute.py:_process/_encode_gl/_prior_matrix/_format_subject of this package,
// which in turn follow the reference (nmdp-bioinformatics/py-graph-imputation,
// grim/imputation/impute.py:24-118,246-272,1844-1975,1985-2155).  No CUDA calls in this file
// except through grimb_impute_host.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/grimb200.h"

extern "C" void* grimb_pinned_alloc(size_t bytes);  // grimb200.cu: cudaMallocHost or nullptr
extern "C" void grimb_pinned_free(void* p);
extern "C" void grimb_set_error(const char* msg);   // grimb200.cu: the message grimb_last_error() returns

namespace {

using sv = std::string_view;

int tfail(int code, const std::string& m) {
  grimb_set_error(m.c_str());
  return code;
}

enum { H_OK = 0, H_PROBLEM = 1, H_FAULT = 2 };

struct Unknown {
  uint8_t locus;
  uint16_t id;
  std::string name;
};

struct Line {
  sv raw, sid, race1, race2;
  bool has_race = false;
  bool no_fields = false;  // rejected by the field-count check (IndexError in the reference's parser)
  uint8_t hclass = H_OK;
  uint16_t mask = 0;
  uint16_t counts[GRIMB_MAX_LOCI * 2];
  uint32_t ids_off = 0, ids_cnt = 0;   // into the owning thread's id vector
  uint32_t unk_off = 0, unk_cnt = 0;   // into the owning thread's Unknown vector
  uint32_t thread = 0;
  uint32_t prior = 0;
};

struct SvHash {
  size_t operator()(sv s) const { return std::hash<sv>()(s); }
};

inline bool py_space(unsigned char c) { return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x20); }

void split(sv s, char d, std::vector<sv>& out) {
  out.clear();
  size_t b = 0;
  for (;;) {
    size_t e = s.find(d, b);
    if (e == sv::npos) {
      out.push_back(s.substr(b));
      return;
    }
    out.push_back(s.substr(b, e - b));
    b = e + 1;
  }
}

// Python's repr(float) / str(float): shortest round-trip digits, fixed notation iff
// -4 < decpt <= 16 (Python/pystrtod.c format_float_short, mode 'r')
void py_float(double x, std::string& out) {
  if (x == 0.0) {
    out += std::signbit(x) ? "-0.0" : "0.0";
    return;
  }
  if (std::isnan(x)) { out += "nan"; return; }
  if (std::isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
  char buf[48];
  auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);
  sv s(buf, (size_t)(r.ptr - buf));
  if (s[0] == '-') {
    out += '-';
    s.remove_prefix(1);
  }
  size_t e = s.find('e');
  int exp10 = 0;
  std::from_chars(s.data() + e + 1 + (s[e + 1] == '+' ? 1 : 0), s.data() + s.size(), exp10);
  char digits[24];
  int nd = 0;
  for (size_t i = 0; i < e; ++i)
    if (s[i] != '.') digits[nd++] = s[i];
  const int decpt = exp10 + 1;
  if (decpt <= -4 || decpt > 16) {
    out += digits[0];
    if (nd > 1) {
      out += '.';
      out.append(digits + 1, (size_t)(nd - 1));
    }
    out += 'e';
    int ex = decpt - 1;
    out += ex < 0 ? '-' : '+';
    if (ex < 0) ex = -ex;
    char eb[8];
    int n = 0;
    do {
      eb[n++] = (char)('0' + ex % 10);
      ex /= 10;
    } while (ex);
    if (n < 2) eb[n++] = '0';
    while (n) out += eb[--n];
  } else if (decpt <= 0) {
    out += "0.";
    out.append((size_t)(-decpt), '0');
    out.append(digits, (size_t)nd);
  } else if (decpt >= nd) {
    out.append(digits, (size_t)nd);
    out.append((size_t)(decpt - nd), '0');
    out += ".0";
  } else {
    out.append(digits, (size_t)decpt);
    out += '.';
    out.append(digits + decpt, (size_t)(nd - decpt));
  }
}

void put_uint(uint64_t v, std::string& out) {
  char b[24];
  auto r = std::to_chars(b, b + sizeof(b), v);
  out.append(b, (size_t)(r.ptr - b));
}

struct PriorKey {
  bool has;
  std::string r1, r2;
  bool operator==(const PriorKey& o) const { return has == o.has && r1 == o.r1 && r2 == o.r2; }
};
struct PriorKeyHash {
  size_t operator()(const PriorKey& k) const {
    return std::hash<std::string>()(k.r1) * 1000003u ^ std::hash<std::string>()(k.r2) ^ (k.has ? 0x9e3779b9u : 0u);
  }
};

}  // namespace

struct GrimbText {
  int L = 0, P = 0, n_threads = 1;
  std::vector<std::string> loci, pops;
  std::vector<std::vector<std::string>> alleles;              // [L][n]
  std::vector<std::unordered_map<sv, uint16_t, SvHash>> ids;  // [L] name -> id
  std::unordered_map<sv, int, SvHash> locus_index;
  std::unordered_map<sv, int, SvHash> pop_index;
  std::vector<double> count_by_prob;
  double alpha = 0, eta = 0, beta = 0, gamma = 0, delta = 0;
  bool mr = true;
  int key_bits[GRIMB_MAX_LOCI];
  int shift[GRIMB_MAX_LOCI];
  // priors (memoised across calls)
  std::unordered_map<PriorKey, uint32_t, PriorKeyHash> prior_index;
  std::vector<double> priors;
  // state of the current call
  std::string text;   // private copy of the input (string_views point into it)
  int64_t first_index = 0;
  std::vector<Line> lines;
  std::vector<std::string> fmt_parts;   // per-thread pieces of the six outputs (capacity reused)
  std::vector<std::vector<uint16_t>> t_ids;
  std::vector<std::vector<Unknown>> t_unk;
  std::vector<std::string> t_clean;  // unused placeholder (cleaned GL strings are per-call locals)
  std::vector<uint16_t> b_typed, b_counts, b_alleles;
  std::vector<uint32_t> b_off, b_prior;
  std::string out[6];
  // pinned staging for grimb_impute_text (grow-only) and the adaptive row-capacity estimate
  struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    void* reserve(size_t bytes) {
      if (bytes <= cap && p) return p;
      release();
      size_t want = bytes + bytes / 4 + 4096;
      p = grimb_pinned_alloc(want);
      pinned = p != nullptr;
      if (!p) p = malloc(want);
      cap = p ? want : 0;
      return p;
    }
    void release() {
      if (p) {
        if (pinned) grimb_pinned_free(p);
        else free(p);
      }
      p = nullptr;
      cap = 0;
    }
    ~HostBuf() { release(); }
  };
  HostBuf hb_compact, hb_words, hb_general, hb_hap, hb_pop;
  double per_subject[4] = {0.5, 0.05, 0.3, 0.3};   // words, general records, hap rows, pop rows

  template <class F>
  void parallel(size_t n, F f) const {
    int nt = n_threads;
    if ((size_t)nt > n) nt = (int)(n ? n : 1);
    if (nt <= 1) {
      f(0, (size_t)0, n);
      return;
    }
    std::vector<std::thread> th;
    size_t per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
      size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
      th.emplace_back([=]() { f(t, lo, hi); });
    }
    for (auto& x : th) x.join();
  }

  // calc_priority_matrix (reference impute.py:1844-1924), same operation order
  void prior_matrix(const std::vector<int>& r1, const std::vector<int>& r2, std::vector<double>& M) const {
    const int n = P;
    M.assign((size_t)n * n, 0.0);
    std::vector<double> T((size_t)n * n), U((size_t)n * n);
    for (int a : r1)
      for (int b : r2) {
        if (a < 0 && b < 0) continue;
        std::fill(T.begin(), T.end(), 0.0);
        if (a < 0 || b < 0) {
          const int r = a < 0 ? b : a;
          for (int i = 0; i < n; ++i) T[(size_t)r * n + i] = T[(size_t)r * n + i] + gamma * 2;
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) U[(size_t)i * n + j] = T[(size_t)i * n + j] + T[(size_t)j * n + i];
          T.swap(U);
          T[(size_t)r * n + r] -= gamma * 2;
        } else {
          for (int i = 0; i < n; ++i) {
            T[(size_t)a * n + i] = T[(size_t)a * n + i] + gamma;
            T[(size_t)i * n + b] = T[(size_t)i * n + b] + gamma;
          }
          T[(size_t)a * n + b] -= gamma;
          T[(size_t)a * n + b] = T[(size_t)a * n + b] + alpha;
          if (a != b) {
            for (int i = 0; i < n; ++i)
              for (int j = 0; j < n; ++j) U[(size_t)i * n + j] = T[(size_t)i * n + j] + T[(size_t)j * n + i];
            T.swap(U);
            T[(size_t)a * n + a] -= gamma;
            T[(size_t)b * n + b] -= gamma;
          }
          T[(size_t)a * n + a] += delta;
          if (a != b) T[(size_t)b * n + b] += delta;
        }
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            const double e = (i == j) ? 1.0 : 0.0;
            const double v = (eta * 1.0 + T[(size_t)i * n + j]) + beta * e;
            M[(size_t)i * n + j] += v;
          }
      }
    double total = 0.0;
    bool first = true;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double& m = M[(size_t)i * n + j];
        m = m * count_by_prob[i] * count_by_prob[j];
        if (first) {
          total = m;
          first = false;
        } else {
          total += m;
        }
      }
    for (double& m : M) m = m / total;
  }

  // consecutive lines mostly repeat the race fields: remember the last key
  bool last_has_race = false, last_valid = false;
  sv last_r1, last_r2;
  uint32_t last_prior = 0;

  uint32_t prior_for(const Line& ln) {
    if (last_valid && ln.has_race == last_has_race && ln.race1 == last_r1 && ln.race2 == last_r2) return last_prior;
    const uint32_t idx = prior_lookup(ln);
    last_valid = true;
    last_has_race = ln.has_race;
    last_r1 = ln.race1;      // views into this call's text buffer: reset by grimb_text_tokenise
    last_r2 = ln.race2;
    last_prior = idx;
    return idx;
  }

  uint32_t prior_lookup(const Line& ln) {
    PriorKey k{ln.has_race, std::string(ln.race1), std::string(ln.race2)};
    auto it = prior_index.find(k);
    if (it != prior_index.end()) return it->second;
    std::vector<double> M((size_t)P * P, mr ? 1.0 : 0.0);
    if (!mr)
      for (int i = 0; i < P; ++i) M[(size_t)i * P + i] = 1.0;
    if (ln.has_race && (!ln.race1.empty() || !ln.race2.empty())) {
      bool known = false;
      std::vector<sv> parts;
      std::vector<int> r1, r2;
      split(ln.race1, ';', parts);
      for (sv p : parts) {
        auto f = pop_index.find(p);
        r1.push_back(f == pop_index.end() ? -1 : f->second);
        known = known || f != pop_index.end();
      }
      split(ln.race2, ';', parts);
      for (sv p : parts) {
        auto f = pop_index.find(p);
        r2.push_back(f == pop_index.end() ? -1 : f->second);
        known = known || f != pop_index.end();
      }
      if (known) prior_matrix(r1, r2, M);
    }
    uint32_t idx = (uint32_t)(priors.size() / ((size_t)P * P));
    priors.insert(priors.end(), M.begin(), M.end());
    prior_index.emplace(std::move(k), idx);
    return idx;
  }

  // one input line -> Line (reference impute.py:2022-2036 + clean_up_gl + gl2haps)
  void parse_line(sv raw_line, Line& ln, int thread, bool planb, std::string& clean, std::vector<sv>& f1,
                  std::vector<sv>& f2, std::vector<sv>& t1, std::vector<sv>& t2, std::vector<sv>& names,
                  std::vector<uint16_t>& scr) {
    size_t e = raw_line.size();
    while (e > 0 && py_space((unsigned char)raw_line[e - 1])) --e;
    sv raw = raw_line.substr(0, e);
    ln.raw = raw;
    ln.thread = (uint32_t)thread;
    split(raw, raw.find(',') != sv::npos ? ',' : '%', f1);
    ln.sid = f1[0];
    if (f1.size() < 2 || f1.size() == 3) {
      ln.hclass = H_FAULT;  // IndexError in the reference's line parser
      ln.no_fields = true;
      return;
    }
    sv gl = f1[1];
    if (f1.size() > 2) {
      ln.has_race = true;
      ln.race1 = f1[2];
      ln.race2 = f1[3];
    }
    if (gl.empty()) {
      ln.hclass = H_PROBLEM;
      return;
    }
    // clean_up_gl: drop every 'g' and 'L', then loci that start or end with 'U'
    clean.clear();
    for (char c : gl)
      if (c != 'g' && c != 'L') clean += c;
    split(sv(clean), '^', f2);
    size_t w = 0;
    for (sv p : f2)
      if (!(!p.empty() && (p.front() == 'U' || p.back() == 'U'))) f2[w++] = p;
    f2.resize(w);
    if (f2.empty() || (f2.size() == 1 && (f2[0].empty() || f2[0] == " "))) {
      ln.hclass = H_PROBLEM;
      return;
    }
    t1.clear();
    t2.clear();
    for (sv chunk : f2) {
      if (chunk.empty()) {
        ln.hclass = H_FAULT;  // IndexError on chunk[0]
        return;
      }
      if (chunk[0] == '+') chunk.remove_prefix(1);
      size_t p = chunk.find('+');
      if (p == sv::npos) {
        if (chunk.empty()) continue;
        ln.hclass = H_PROBLEM;  // a locus without '+': gl2haps returns []
        return;
      }
      sv rest = chunk.substr(p + 1);
      size_t q = rest.find('+');
      t1.push_back(chunk.substr(0, p));
      t2.push_back(q == sv::npos ? rest : rest.substr(0, q));
    }
    if (t1.empty()) {
      ln.hclass = H_FAULT;  // 2 ** (0 - 1) phases
      return;
    }
    std::sort(t1.begin(), t1.end());
    std::sort(t2.begin(), t2.end());
    std::vector<uint16_t>& idv = t_ids[thread];
    std::vector<Unknown>& unk = t_unk[thread];
    const uint32_t ids_off = (uint32_t)idv.size(), unk_off = (uint32_t)unk.size();
    // the id lists of this line, built in a reused scratch vector: [off, off + cnt) per locus and side
    scr.clear();
    uint32_t per_off[GRIMB_MAX_LOCI][2] = {{0}}, per_cnt[GRIMB_MAX_LOCI][2] = {{0}};
    bool used[GRIMB_MAX_LOCI] = {false};
    auto foreign = [&]() {
      idv.resize(ids_off);
      unk.resize(unk_off);
      ln.hclass = planb ? H_FAULT : H_OK;  // mask stays 0: nothing is imputed
      ln.mask = 0;
    };
    for (size_t k = 0; k < t1.size(); ++k) {
      sv first = t1[k].substr(0, t1[k].find('/'));
      sv prefix = first.substr(0, first.find('*'));
      auto li = locus_index.find(prefix);
      if (li == locus_index.end() || used[li->second]) return foreign();
      const int l = li->second;
      used[l] = true;
      const int n_tab = (int)alleles[l].size();
      const int cap = (1 << key_bits[l]) - 1;
      std::vector<std::pair<sv, uint16_t>> local;
      for (int x = 0; x < 2; ++x) {
        split(x ? t2[k] : t1[k], '/', names);
        const uint32_t lst_off = (uint32_t)scr.size();
        per_off[l][x] = lst_off;
        for (sv name : names) {
          if (name.substr(0, name.find('*')) != sv(loci[l])) return foreign();
          uint16_t id = 0;
          auto it = ids[l].find(name);
          if (it != ids[l].end()) id = it->second;
          else {
            for (auto& pr : local)
              if (pr.first == name) id = pr.second;
            if (!id) {
              int nid = n_tab + 1 + (int)local.size();
              if (nid > cap) {
                idv.resize(ids_off);
                unk.resize(unk_off);
                ln.hclass = H_FAULT;
                return;
              }
              id = (uint16_t)nid;
              local.emplace_back(name, id);
              unk.push_back(Unknown{(uint8_t)l, id, std::string(name)});
            }
          }
          if (std::find(scr.begin() + lst_off, scr.end(), id) == scr.end()) scr.push_back(id);
        }
        per_cnt[l][x] = (uint32_t)scr.size() - lst_off;
      }
    }
    memset(ln.counts, 0, sizeof(ln.counts));
    for (int l = 0; l < L; ++l)
      if (used[l]) {
        ln.mask |= (uint16_t)(1u << l);
        for (int x = 0; x < 2; ++x) {
          ln.counts[l * 2 + x] = (uint16_t)per_cnt[l][x];
          idv.insert(idv.end(), scr.begin() + per_off[l][x], scr.begin() + per_off[l][x] + per_cnt[l][x]);
        }
      }
    ln.ids_off = ids_off;
    ln.ids_cnt = (uint32_t)idv.size() - ids_off;
    ln.unk_off = unk_off;
    ln.unk_cnt = (uint32_t)unk.size() - unk_off;
  }

  sv allele_name(int l, uint32_t id, const Line& ln) const {
    if (id >= 1 && id <= alleles[l].size()) return alleles[l][id - 1];
    const std::vector<Unknown>& u = t_unk[ln.thread];
    for (uint32_t i = 0; i < ln.unk_cnt; ++i)
      if (u[ln.unk_off + i].locus == l && u[ln.unk_off + i].id == id) return u[ln.unk_off + i].name;
    return "?";
  }

  // allele id of locus l inside a packed key (GRIMB_KEY_WORDS little-endian 64-bit words)
#if GRIMB_KEY_WORDS == 1
  typedef uint64_t keyref;
  uint32_t field(keyref key, int l) const { return (uint32_t)((key >> shift[l]) & ((1ull << key_bits[l]) - 1ull)); }
#else
  typedef const uint64_t* keyref;
  uint32_t field(keyref key, int l) const {
    const unsigned __int128 k = (unsigned __int128)key[0] | ((unsigned __int128)key[1] << 64);
    return (uint32_t)((uint64_t)(k >> shift[l]) & ((1ull << key_bits[l]) - 1ull));
  }
#endif

  void put_hap(keyref key, const Line& ln, std::string& o) const {
    bool first = true;
    for (int l = 0; l < L; ++l) {
      uint32_t id = field(key, l);
      if (!id) continue;
      if (!first) o += '~';
      first = false;
      o += allele_name(l, id, ln);
    }
  }

  sv pop_name(uint16_t p) const { return p == 0xFFFF ? sv("all_pops") : sv(pops[p]); }

  static void put_row_tail(double prob, uint32_t k, std::string& s) {
    s += ',';
    py_float(prob, s);
    s += ',';
    put_uint(k, s);
    s += '\n';
  }

  void put_pop_row(sv sid, sv x, sv y, double prob, uint32_t k, bool sorted, std::string& s) const {
    if (sorted && y < x) std::swap(x, y);
    s += sid;
    s += ',';
    s += x;
    s += ',';
    s += y;
    put_row_tail(prob, k, s);
  }

  // Rows of a subject the warp-per-subject kernels finished (GRIMB_KIND_SIMPLE / GRIMB_KIND_TYPED, ABI v4):
  // every locus typed with one allele per side, so the single UMUG genotype is the subject's own allele
  // pairs and a PMUG row's haplotypes follow from its phase id (bit m: locus m takes its side-2 allele in
  // the first haplotype).  Same row order and text as the general layout below.
  void format_compact(const Line& ln, const GrimbCompact& c, const uint64_t* words, const GrimbConfig* cfg,
                      std::string* o) const {
    const uint16_t* ids = t_ids[ln.thread].data() + ln.ids_off;   // [L][2]
    const uint32_t kind = c.kind_flags & 3u, n_pmug = c.kind_flags >> 4;
    const bool has = (c.kind_flags & GRIMB_KIND_HAS_RESULTS) != 0;
    const uint64_t* w = words + c.off;
    uint32_t n_pops = 0;
    uint32_t phase[4] = {0, 0, 0, 0};
    const uint64_t* pmug_prob = nullptr;   // nullptr: the single PMUG row carries `total`
    const uint64_t* pop_prob = nullptr;
    const uint64_t* pop_code = nullptr;
    if (kind == GRIMB_KIND_SIMPLE) {
      for (uint32_t k = 0; k < n_pmug; ++k) phase[k] = (c.phases >> (4 * k)) & 15u;
      if (c.kind_flags & GRIMB_KIND_WORDS) pmug_prob = w;
      n_pops = (has && cfg->n_pop_results >= 1) ? 1u : 0u;
    } else {
      const uint64_t hdr = w[0];
      n_pops = (uint32_t)(hdr & 0xFFFFu);
      for (uint32_t k = 0; k < n_pmug; ++k) phase[k] = (uint32_t)(hdr >> (16 + 12 * k)) & 0xFFFu;
      pmug_prob = w + 1;
      pop_prob = w + 1 + n_pmug;
      pop_code = pop_prob + n_pops;
    }
    auto dbl = [](uint64_t u) {
      double d;
      memcpy(&d, &u, 8);
      return d;
    };
    auto pop_pair = [&](uint32_t k, sv& x, sv& y, double& p) {
      if (!pop_code) {
        x = y = sv(pops[0]);
        p = c.total;
        return;
      }
      const uint32_t code = (uint32_t)(pop_code[k >> 2] >> (16 * (k & 3u))) & 0xFFFFu;
      x = sv(pops[code >> 8]);
      y = sv(pops[code & 0xFFu]);
      p = dbl(pop_prob[k]);
    };
    if (cfg->output_pmug) {
      for (uint32_t k = 0; k < n_pmug; ++k) {
        std::string& s = o[GRIMB_OUT_PMUG];
        s += ln.sid;
        s += ',';
        for (int side = 0; side < 2; ++side) {
          if (side) s += '+';
          for (int l = 0; l < L; ++l) {
            if (l) s += '~';
            s += allele_name(l, ids[2 * l + (int)(((phase[k] >> l) & 1u) ^ (uint32_t)side)], ln);
          }
        }
        put_row_tail(pmug_prob ? dbl(pmug_prob[k]) : c.total, k, s);
      }
      for (uint32_t k = 0; k < n_pops; ++k) {
        sv x, y;
        double p;
        pop_pair(k, x, y, p);
        put_pop_row(ln.sid, x, y, p, k, false, o[GRIMB_OUT_PMUG_POPS]);
      }
    }
    if (cfg->output_umug && has) {
      if (cfg->n_results >= 1) {
        std::string& s = o[GRIMB_OUT_UMUG];
        s += ln.sid;
        s += ',';
        for (int l = 0; l < L; ++l) {
          sv x = allele_name(l, ids[2 * l], ln), y = allele_name(l, ids[2 * l + 1], ln);
          if (y < x) std::swap(x, y);
          if (l) s += '^';
          s += x;
          s += '+';
          s += y;
        }
        put_row_tail(c.total, 0, s);
      }
      for (uint32_t k = 0; k < n_pops; ++k) {
        sv x, y;
        double p;
        pop_pair(k, x, y, p);
        put_pop_row(ln.sid, x, y, p, k, true, o[GRIMB_OUT_UMUG_POPS]);
      }
    }
  }

  void format_subject(const Line& ln, const GrimbSubjectResult& r, const GrimbHapRow* hr, const GrimbPopRow* pr,
                      const GrimbConfig* cfg, std::string* o) const {
    if (cfg->output_pmug) {
      for (uint32_t k = 0; k < r.n_pmug; ++k) {
        const GrimbHapRow& row = hr[r.hap_off + r.n_umug + k];
        std::string& s = o[GRIMB_OUT_PMUG];
        s += ln.sid;
        s += ',';
        put_hap(row.a, ln, s);
        s += '+';
        put_hap(row.b, ln, s);
        put_row_tail(row.prob, k, s);
      }
      for (uint32_t k = 0; k < r.n_pmug_pops; ++k) {
        const GrimbPopRow& row = pr[r.pop_off + r.n_umug_pops + k];
        put_pop_row(ln.sid, pop_name(row.pop_a), pop_name(row.pop_b), row.prob, k, false, o[GRIMB_OUT_PMUG_POPS]);
      }
    }
    if (cfg->output_umug) {
      for (uint32_t k = 0; k < r.n_umug; ++k) {
        const GrimbHapRow& row = hr[r.hap_off + k];
        std::string& s = o[GRIMB_OUT_UMUG];
        s += ln.sid;
        s += ',';
        bool first = true;
        for (int l = 0; l < L; ++l) {
          uint32_t a = field(row.a, l), b = field(row.b, l);
          if (!a) continue;
          sv x = allele_name(l, a, ln), y = allele_name(l, b, ln);
          if (y < x) std::swap(x, y);
          if (!first) s += '^';
          first = false;
          s += x;
          s += '+';
          s += y;
        }
        put_row_tail(row.prob, k, s);
      }
      const bool planc_empty = r.plan_umug == GRIMB_PLAN_C && r.tot_umug == 0;
      for (uint32_t k = 0; k < r.n_umug_pops; ++k) {
        const GrimbPopRow& row = pr[r.pop_off + k];
        std::string& s = o[GRIMB_OUT_UMUG_POPS];
        sv x = pop_name(row.pop_a), y = pop_name(row.pop_b);
        if (y < x) std::swap(x, y);
        s += ln.sid;
        s += ',';
        s += x;
        s += ',';
        s += y;
        s += ',';
        if (planc_empty) s += '0';  // sum() of an empty dict is the int 0 (reference impute.py:1376)
        else py_float(row.prob, s);
        s += ',';
        put_uint(k, s);
        s += '\n';
      }
    }
  }
};

extern "C" int grimb_text_create(const GrimbTextDesc* d, GrimbText** out) {
  if (!d || !out || d->n_loci < 1 || d->n_loci > GRIMB_MAX_LOCI || d->n_pops < 1) return tfail(GRIMB_E_ARG, "bad text descriptor");
  GrimbText* t = new GrimbText();
  t->L = d->n_loci;
  t->P = d->n_pops;
  t->n_threads = d->n_threads > 0 ? d->n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  size_t pos = 0;
  int sh = 0;
  t->alleles.resize(t->L);
  t->ids.resize(t->L);
  for (int l = 0; l < t->L; ++l) {
    t->loci.emplace_back(d->locus_names[l]);
    t->key_bits[l] = d->key_bits[l];
    t->shift[l] = sh;
    sh += d->key_bits[l];
    t->alleles[l].reserve((size_t)d->allele_counts[l]);
    for (int i = 0; i < d->allele_counts[l]; ++i) t->alleles[l].emplace_back(d->allele_names[pos++]);
  }
  for (int l = 0; l < t->L; ++l) {
    t->locus_index.emplace(sv(t->loci[l]), l);
    for (size_t i = 0; i < t->alleles[l].size(); ++i) t->ids[l].emplace(sv(t->alleles[l][i]), (uint16_t)(i + 1));
  }
  for (int p = 0; p < t->P; ++p) t->pops.emplace_back(d->pop_names[p]);
  for (int p = 0; p < t->P; ++p) t->pop_index.emplace(sv(t->pops[p]), p);  // first occurrence wins, like list.index
  t->count_by_prob.assign(d->count_by_prob, d->count_by_prob + t->P);
  t->alpha = d->alpha;
  t->eta = d->eta;
  t->beta = d->beta;
  t->gamma = d->gamma;
  t->delta = d->delta;
  t->mr = d->unk_priors_mr != 0;
  *out = t;
  return GRIMB_OK;
}

extern "C" int grimb_text_free(GrimbText* t) {
  delete t;
  return GRIMB_OK;
}

extern "C" int grimb_text_tokenise(GrimbText* t, const GrimbConfig* cfg, const char* text, int64_t len,
                                   int64_t first_line_index, GrimbBatch* b) {
  if (!t || !cfg || !text || !b || len < 0) return tfail(GRIMB_E_ARG, "null argument");
  t->text.assign(text, (size_t)len);
  t->first_index = first_line_index;
  // line boundaries (Python file iteration: split on '\n', a final line without '\n' counts)
  std::vector<std::pair<size_t, size_t>> bounds;
  {
    const char* p = t->text.data();
    size_t n = t->text.size(), b0 = 0;
    while (b0 < n) {
      const void* q = memchr(p + b0, '\n', n - b0);
      size_t e = q ? (size_t)((const char*)q - p) + 1 : n;
      bounds.emplace_back(b0, e);
      b0 = e;
    }
  }
  const size_t S = bounds.size();
  t->lines.assign(S, Line());
  t->t_ids.assign((size_t)t->n_threads, {});
  t->t_unk.assign((size_t)t->n_threads, {});
  t->last_valid = false;   // the remembered race fields point into the previous call's text
  const bool planb = cfg->planb != 0;
  // the cleaned GL strings must outlive the call for unknown-allele names: names are copied
  t->parallel(S, [&](int th, size_t lo, size_t hi) {
    std::string clean;
    std::vector<sv> f1, f2, t1, t2, names;
    std::vector<uint16_t> scr;
    for (size_t i = lo; i < hi; ++i)
      t->parse_line(sv(t->text.data() + bounds[i].first, bounds[i].second - bounds[i].first), t->lines[i], th, planb,
                    clean, f1, f2, t1, t2, names, scr);
  });
  // sequential: prior indices (memoised) + flat batch arrays
  const int L = t->L;
  t->b_typed.assign(S, 0);
  t->b_counts.assign(S * (size_t)L * 2, 0);
  t->b_off.assign(S + 1, 0);
  t->b_prior.assign(S, 0);
  size_t total = 0;
  for (size_t i = 0; i < S; ++i) {
    Line& ln = t->lines[i];
    if (!ln.no_fields) ln.prior = t->prior_for(ln);  // the reference computes the prior before looking at the GL
    t->b_off[i] = (uint32_t)total;
    if (ln.hclass == H_OK && ln.mask) total += ln.ids_cnt;
  }
  t->b_off[S] = (uint32_t)total;
  if (t->priors.empty()) {
    Line dummy;
    t->prior_for(dummy);
  }
  t->b_alleles.assign(total ? total : 1, 0);
  t->parallel(S, [&](int, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      const Line& ln = t->lines[i];
      t->b_prior[i] = ln.prior;
      if (ln.hclass != H_OK || !ln.mask) continue;
      t->b_typed[i] = ln.mask;
      for (int q = 0; q < L * 2; ++q) t->b_counts[i * (size_t)L * 2 + q] = ln.counts[q];
      const std::vector<uint16_t>& idv = t->t_ids[ln.thread];
      std::copy(idv.begin() + ln.ids_off, idv.begin() + ln.ids_off + ln.ids_cnt, t->b_alleles.begin() + t->b_off[i]);
    }
  });
  // ABI v4: the counts travel only when some subject lists several alleles on a side (every typed side
  // lists at least one), the prior indices only when more than one prior matrix is in use
  bool all_single = true, one_prior = true;
  for (size_t i = 0; i < S && (all_single || one_prior); ++i) {
    const Line& ln = t->lines[i];
    if (t->b_prior[i] != 0) one_prior = false;
    if (ln.hclass == H_OK && ln.mask && ln.ids_cnt != 2u * (uint32_t)__builtin_popcount(ln.mask)) all_single = false;
  }
  b->n_subjects = (int64_t)S;
  b->typed_mask = t->b_typed.data();
  b->counts = all_single ? nullptr : t->b_counts.data();
  b->allele_off = t->b_off.data();
  b->alleles = t->b_alleles.data();
  b->n_alleles_total = (int64_t)total;
  b->prior_index = one_prior ? nullptr : t->b_prior.data();
  b->priors = t->priors.data();
  b->n_priors = (int32_t)(t->priors.size() / ((size_t)t->P * t->P));
  b->phase_mask = nullptr;   // default phase enumeration; masks are served by the numpy host front end
  return GRIMB_OK;
}

extern "C" int grimb_text_format(GrimbText* t, const GrimbConfig* cfg, const GrimbResults* res, GrimbTextOut* out) {
  if (!t || !cfg || !res || !out || !res->compact) return tfail(GRIMB_E_ARG, "null argument");
  const size_t S = t->lines.size();
  const int nt = t->n_threads;
  // per-thread output pieces live in the GrimbText so their capacity is reused from call to call
  std::vector<std::string>& parts = t->fmt_parts;
  parts.resize((size_t)nt * 6);
  for (auto& ps : parts) ps.clear();
  std::vector<int64_t> plans((size_t)nt * 4, 0);
  static const GrimbSubjectResult kNoRecord = {};   // a skipped subject: nothing was computed
  t->parallel(S, [&](int th, size_t lo, size_t hi) {
    std::string* o = &parts[(size_t)th * 6];
    for (size_t i = lo; i < hi; ++i) {
      const Line& ln = t->lines[i];
      const uint64_t idx = (uint64_t)t->first_index + i;
      if (ln.hclass == H_PROBLEM) {
        put_uint(idx, o[GRIMB_OUT_PROBLEM]);
        o[GRIMB_OUT_PROBLEM] += ',';
        o[GRIMB_OUT_PROBLEM] += ln.sid;
        o[GRIMB_OUT_PROBLEM] += '\n';
        continue;
      }
      if (ln.hclass == H_FAULT) {
        o[GRIMB_OUT_PROBLEM] += ln.raw;
        o[GRIMB_OUT_PROBLEM] += '\n';
        continue;
      }
      const GrimbCompact& c = res->compact[i];
      if (c.status == GRIMB_ST_FAULT) {
        o[GRIMB_OUT_PROBLEM] += ln.raw;
        o[GRIMB_OUT_PROBLEM] += '\n';
        continue;
      }
      if (c.status == GRIMB_ST_NO_PHASES) {
        // nothing opens: the reference returns its defaults and the PMUG writer then raises
        if (cfg->output_pmug) {
          o[GRIMB_OUT_PROBLEM] += ln.raw;
          o[GRIMB_OUT_PROBLEM] += '\n';
        }
        continue;
      }
      const uint32_t kind = c.kind_flags & 3u;
      if (kind != GRIMB_KIND_GENERAL) {
        plans[(size_t)th * 4 + GRIMB_PLAN_A] += 1;
        if (cfg->output_pmug && !(c.kind_flags & GRIMB_KIND_HAS_RESULTS)) {
          put_uint(idx, o[GRIMB_OUT_MISS]);
          o[GRIMB_OUT_MISS] += ',';
          o[GRIMB_OUT_MISS] += ln.sid;
          o[GRIMB_OUT_MISS] += '\n';
        }
        t->format_compact(ln, c, res->words, cfg, o);
        continue;
      }
      const GrimbSubjectResult& r = c.off == 0xFFFFFFFFu ? kNoRecord : res->general[c.off];
      plans[(size_t)th * 4 + ((cfg->output_umug ? r.plan_umug : r.plan_pmug) & 3)] += 1;
      const bool pm_empty = cfg->output_pmug ? r.tot_pmug == 0 : false;
      if (pm_empty && r.tot_umug == 0) {
        put_uint(idx, o[GRIMB_OUT_MISS]);
        o[GRIMB_OUT_MISS] += ',';
        o[GRIMB_OUT_MISS] += ln.sid;
        o[GRIMB_OUT_MISS] += '\n';
      }
      t->format_subject(ln, r, res->hap_rows, res->pop_rows, cfg, o);
    }
  });
  out->pair_evals = res->totals ? res->totals[4] : 0;
  for (int k = 0; k < 4; ++k) out->plan_count[k] = 0;
  // concatenate the pieces of every output in thread order, the copies themselves in parallel
  std::vector<size_t> offs((size_t)nt * 6, 0);
  for (int k = 0; k < 6; ++k) {
    size_t n = 0;
    for (int th = 0; th < nt; ++th) {
      offs[(size_t)th * 6 + k] = n;
      n += parts[(size_t)th * 6 + k].size();
    }
    if (t->out[k].size() < n) t->out[k].resize(n);   // grow-only buffer; size[k] carries the valid length
    out->data[k] = t->out[k].data();
    out->size[k] = (int64_t)n;
  }
  t->parallel((size_t)nt, [&](int, size_t lo, size_t hi) {
    for (size_t th = lo; th < hi; ++th)
      for (int k = 0; k < 6; ++k) {
        const std::string& ps = parts[th * 6 + k];
        if (!ps.empty()) memcpy(&t->out[k][offs[th * 6 + k]], ps.data(), ps.size());
      }
  });
  for (int th = 0; th < nt; ++th)
    for (int k = 0; k < 4; ++k) out->plan_count[k] += plans[(size_t)th * 4 + k];
  out->n_lines = (int64_t)S;
  return GRIMB_OK;
}

extern "C" int grimb_impute_text(GrimbText* t, GrimbEngine* const* engines, int32_t n_engines, const GrimbConfig* cfg,
                                 const char* text, int64_t len, int64_t first_line_index, GrimbTextOut* out) {
  if (!t || !engines || n_engines < 1 || !cfg || !out) return tfail(GRIMB_E_ARG, "null argument");
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  GrimbBatch b;
  int rc = grimb_text_tokenise(t, cfg, text, len, first_line_index, &b);
  if (rc) return rc;
  auto t1 = clk::now();
  const size_t S = (size_t)b.n_subjects;
  const int L = t->L;
  // One ABI call per workspace tier.  Tier 0 takes the whole batch and its results are formatted straight
  // from the pinned staging buffers; only when some subject overflowed its workspace (rare) are the tiers'
  // results merged into one set of arrays first.
  std::vector<GrimbCompact> m_compact;
  std::vector<uint64_t> m_words;
  std::vector<GrimbSubjectResult> m_general;
  std::vector<GrimbHapRow> m_hap;
  std::vector<GrimbPopRow> m_pop;
  std::vector<uint32_t> todo;
  int64_t retries = 0, evals = 0;
  int64_t tot[6] = {0, 0, 0, 0, 0, 0};
  GrimbResults fin;
  memset(&fin, 0, sizeof(fin));
  for (int tier = 0; tier < n_engines && (tier == 0 || !todo.empty()); ++tier) {
    // gather the sub-batch (tier 0: the whole batch as is)
    GrimbBatch sb = b;
    std::vector<uint16_t> g_typed, g_counts, g_all;
    std::vector<uint32_t> g_off, g_prior;
    const size_t n = tier == 0 ? S : todo.size();
    if (tier > 0) {
      g_typed.resize(n);
      g_counts.resize(n * (size_t)L * 2);
      g_off.resize(n + 1);
      g_prior.resize(n);
      size_t tot_al = 0;
      for (size_t k = 0; k < n; ++k) {
        const uint32_t s = todo[k];
        g_typed[k] = b.typed_mask[s];
        g_prior[k] = b.prior_index ? b.prior_index[s] : 0u;
        if (b.counts) memcpy(&g_counts[k * (size_t)L * 2], b.counts + (size_t)s * L * 2, (size_t)L * 4);
        g_off[k] = (uint32_t)tot_al;
        tot_al += b.allele_off[s + 1] - b.allele_off[s];
      }
      g_off[n] = (uint32_t)tot_al;
      g_all.resize(tot_al ? tot_al : 1);
      for (size_t k = 0; k < n; ++k) {
        const uint32_t s = todo[k];
        std::copy(b.alleles + b.allele_off[s], b.alleles + b.allele_off[s + 1], g_all.begin() + g_off[k]);
      }
      sb.n_subjects = (int64_t)n;
      sb.typed_mask = g_typed.data();
      sb.counts = b.counts ? g_counts.data() : nullptr;
      sb.allele_off = g_off.data();
      sb.alleles = g_all.data();
      sb.n_alleles_total = (int64_t)tot_al;
      sb.prior_index = g_prior.data();
    }
    // pinned, grow-only staging; capacities follow what the previous call needed per subject (+25 %)
    GrimbCompact* rc_ = (GrimbCompact*)t->hb_compact.reserve((n + 1) * sizeof(GrimbCompact));
    int64_t cap[4];
    for (int k = 0; k < 4; ++k) cap[k] = std::max<int64_t>(1024, (int64_t)((double)n * t->per_subject[k] * 1.25) + 64);
    GrimbResults r;
    for (;;) {
      r.compact = rc_;
      r.words = (uint64_t*)t->hb_words.reserve((size_t)cap[0] * 8);
      r.word_capacity = cap[0];
      r.general = (GrimbSubjectResult*)t->hb_general.reserve((size_t)cap[1] * sizeof(GrimbSubjectResult));
      r.general_capacity = cap[1];
      r.hap_rows = (GrimbHapRow*)t->hb_hap.reserve((size_t)cap[2] * sizeof(GrimbHapRow));
      r.hap_capacity = cap[2];
      r.pop_rows = (GrimbPopRow*)t->hb_pop.reserve((size_t)cap[3] * sizeof(GrimbPopRow));
      r.pop_capacity = cap[3];
      r.totals = tot;
      if (!rc_ || !r.words || !r.general || !r.hap_rows || !r.pop_rows) return tfail(GRIMB_E_NOMEM, "host staging allocation failed");
      rc = grimb_impute_host(engines[tier], cfg, &sb, &r);
      if (rc == GRIMB_E_CAPACITY) {
        for (int k = 0; k < 4; ++k) cap[k] = std::max(cap[k], tot[k]);
        continue;
      }
      if (rc) return rc;
      break;
    }
    if (tier == 0 && n > 0)
      for (int k = 0; k < 4; ++k) t->per_subject[k] = std::max(k == 1 ? 0.01 : 0.1, (double)tot[k] / (double)n);
    evals += tot[4];
    // subjects whose workspace overflowed (counted in parallel: the records are only 16 bytes each)
    std::vector<std::vector<uint32_t>> again_t((size_t)t->n_threads);
    t->parallel(n, [&](int th, size_t lo, size_t hi) {
      for (size_t k = lo; k < hi; ++k)
        if (rc_[k].status == GRIMB_ST_WORKSPACE) again_t[th].push_back(tier == 0 ? (uint32_t)k : todo[k]);
    });
    std::vector<uint32_t> again;
    for (auto& v : again_t) again.insert(again.end(), v.begin(), v.end());
    if (tier == 0 && again.empty()) {
      fin = r;   // the usual case: format from the staging buffers
      break;
    }
    // merge this tier into the final arrays, re-basing the offsets
    if (tier == 0) m_compact.assign(rc_, rc_ + n);
    const uint32_t wbase = (uint32_t)m_words.size(), gbase = (uint32_t)m_general.size();
    const uint64_t hbase = m_hap.size(), pbase = m_pop.size();
    m_words.insert(m_words.end(), r.words, r.words + tot[0]);
    for (int64_t k = 0; k < tot[1]; ++k) {
      GrimbSubjectResult o = r.general[k];
      o.hap_off += hbase;
      o.pop_off += pbase;
      m_general.push_back(o);
    }
    m_hap.insert(m_hap.end(), r.hap_rows, r.hap_rows + tot[2]);
    m_pop.insert(m_pop.end(), r.pop_rows, r.pop_rows + tot[3]);
    for (size_t k = 0; k < n; ++k) {
      GrimbCompact c = rc_[k];
      if (c.status == GRIMB_ST_WORKSPACE) continue;
      if ((c.kind_flags & 3u) == GRIMB_KIND_GENERAL) {
        if (c.off != 0xFFFFFFFFu) c.off += gbase;
      } else {
        c.off += wbase;
      }
      m_compact[tier == 0 ? k : todo[k]] = c;
    }
    retries += (int64_t)again.size();
    todo.swap(again);
    fin.compact = m_compact.data();
    fin.words = m_words.data();
    fin.general = m_general.data();
    fin.hap_rows = m_hap.data();
    fin.pop_rows = m_pop.data();
  }
  if (!todo.empty()) return tfail(GRIMB_E_NOMEM, "a subject exceeds the largest workspace tier");
  auto t2 = clk::now();
  int64_t ftot[6] = {0, 0, 0, 0, evals, 0};
  fin.totals = ftot;
  rc = grimb_text_format(t, cfg, &fin, out);
  auto t3 = clk::now();
  out->workspace_retries = retries;
  out->seconds_tokenise = std::chrono::duration<double>(t1 - t0).count();
  out->seconds_gpu = std::chrono::duration<double>(t2 - t1).count();
  out->seconds_format = std::chrono::duration<double>(t3 - t2).count();
  return rc;
}
