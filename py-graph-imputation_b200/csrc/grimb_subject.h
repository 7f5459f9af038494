// grimb_subject.h -- the per-subject imputation algorithm, executed by one thread group (CTA).
//
// What it reproduces (reference = nmdp-bioinformatics/py-graph-imputation, file
// grim/imputation/impute.py unless another file is named):
//   phases            gen_phases :274-303
//   opening           open_phases :914-989 + cutils.pyx:6-31 (mixed-radix decode, never
//                     materialised) and cutils.pyx:35-51 (filter of a label's haplotypes)
//   Plan A probe      comp_hap_prob :353 -> Graph.adjs_query networkx_graph.py:253-278
//   flatten + cap     convert_list_to_one_dim :424-442 (streaming stable top-K)
//   pair evaluation   calc_haps_pairs :444-548 / calc_haps_pairs_haplotype :550-658
//   epsilon schedule  call_comp_phase_prob :1658-1724
//   Plan B            comp_phase_prob_plan_b :1392-1570 and helpers :994-1258
//   Plan C            comp_phase_prob_plan_c :1313-1389, comp_hap_prob_plan_c :1264-1311
//   reductions        reduce_phase_to_valid_allels :864-879, ..._commons_alleles :881-912
//   top-N             write_best_prob :24-58, write_best_prob_genotype :61-76
//
// Ordering contract: the reference's results are Python dicts in first-encounter order of a
// sequential (phase, h, k) traversal, summed with += in that order and ranked with stable
// descending sorts.  Here every accepted pair carries its encounter rank; duplicates are
// resolved with atomicMin on the rank, sums run sequentially in rank order inside each group,
// and all sorts use (value, rank) as a total order, so results are bit-identical for any thread
// interleaving.  FP64 expressions keep the reference's operation order; the translation unit
// must be compiled without FMA contraction (-fmad=false / -ffp-contract=off).
#pragma once
#include "../../include/grimb200.h"
#include "grimb_tables.h"

namespace grimb {

constexpr int MAXT = 512;   // max threads per group
constexpr int MAXL = GRIMB_MAX_LOCI;
constexpr int MAXPH = 256;  // 2^(MAXL-1)
constexpr int NVAR = 4;
constexpr int VAR_ORIG = 0, VAR_VALID = 1, VAR_C10 = 2, VAR_C1 = 3;
constexpr uint32_t NEVER_ROW = 10;  // impute.py:1410

struct SelItem {
  uint64_t wkey;  // ~orderable(weight): ascending = weight descending
  uint64_t ord;   // encounter order inside the side's list
  double f;
  hkey hap;
  uint32_t pop;
  uint32_t pad;
};
struct TopItem {
  double f;
  hkey hap;
  uint32_t pop;
  uint32_t pad;
};
struct Entry {
  hkey h1, h2;
  double prob;
  uint16_t p1, p2;
  uint32_t pad;
};
struct SlotDesc {
  uint64_t ncand;
  const uint32_t* filt;  // filter mode: candidate node ids
  uint8_t var;           // which list variant the slot uses
  uint8_t mode;          // 0 Cartesian, 1 filter
  uint8_t valid;         // phase opened (both sides non-empty)
  uint8_t pad;
  uint32_t first_row;    // Plan B: first matrix row that returned anything (NEVER_ROW)
  uint32_t cached_row;   // row whose top list is currently stored (0xFFFFFFFF none)
};

struct Shared {
  uint32_t scratch[80];
  uint32_t sel_n;
  uint32_t fault;
  uint32_t cnt[8];
  uint32_t chunk_node[MAXT];
  uint32_t chunk_off[MAXT];
  const uint16_t* lptr[NVAR][MAXL][2];
  uint16_t lcnt[NVAR][MAXL][2];
  uint8_t lexist[NVAR][MAXL][2];
  uint8_t lhave[NVAR];
  uint16_t ph[MAXPH];
  uint16_t kbreak[512];
  uint32_t work;
  uint32_t nonempty;  // the side evaluation in progress returned a non-empty dict
  double dmax;
};

struct OutArrays {  // device pointers of GrimbResults + global counters
  GrimbResults r;
  unsigned long long* hap_counter;
  unsigned long long* pop_counter;
  unsigned long long* word_counter;      // 8-byte words of SIMPLE / TYPED subjects
  unsigned long long* general_counter;   // GrimbSubjectResult records appended to r.general
  unsigned long long* evals_counter;     // pair evaluations of the whole batch
  unsigned long long* probe_counters;    // [3] probes issued, probes answered, frequency vectors read
                                         // (general and typed kernels; the single-population fast path
                                         // issues 2 probes per kept phase and is not instrumented)
};

// batch accessors: `counts` and `prior_index` may be NULL (ABI v4: all ones / all zero)
GD uint32_t batch_count(const GrimbBatch& B, uint64_t s, int L, int l, int x) {
  return (B.counts && !B.packed_keys) ? (uint32_t)B.counts[s * (uint64_t)L * 2 + (uint64_t)l * 2 + x] : 1u;
}
GD uint32_t batch_prior(const GrimbBatch& B, uint64_t s) { return B.prior_index ? B.prior_index[s] : 0u; }
// packed form (include/grimb200.h): every subject typed at every locus, one allele per side, or skipped
GD uint32_t batch_typed(const GrimbBatch& B, uint64_t s, uint32_t full) {
  if (B.packed_keys) return (B.packed_flags[s] & 0x8000u) ? 0u : full;
  return B.typed_mask[s];
}

GD GrimbCompact make_compact(uint32_t status, uint32_t kind_flags, uint32_t phases, uint32_t off, double total) {
  GrimbCompact c;
  c.status = (uint8_t)status;
  c.kind_flags = (uint8_t)kind_flags;
  c.phases = (uint16_t)phases;
  c.off = off;
  c.total = total;
  return c;
}

// Publishes the 48-byte record of a GENERAL subject: appended to r.general, the subject's compact
// record points at it.  One thread calls this.
GD void publish_general(const OutArrays& O, uint64_t s, const GrimbSubjectResult& o) {
  GrimbResults& R = const_cast<GrimbResults&>(O.r);
  const unsigned long long gi = atom_add64(O.general_counter, 1ull);
  if ((int64_t)gi < R.general_capacity) R.general[gi] = o;
  const bool has = o.tot_umug != 0 || o.tot_pmug != 0;
  R.compact[s] = make_compact(o.status, GRIMB_KIND_GENERAL | (has ? GRIMB_KIND_HAS_RESULTS : 0), 0, (uint32_t)gi, 0.0);
  if (o.pair_evals) atom_add64(O.evals_counter, (unsigned long long)o.pair_evals);
}

GD GrimbHapRow make_hap_row(hkey a, hkey b, double prob) {
  GrimbHapRow o;
#if GRIMB_KW == 1
  o.a = a;
  o.b = b;
#else
  o.a[0] = (uint64_t)a;
  o.a[1] = (uint64_t)(a >> 64);
  o.b[0] = (uint64_t)b;
  o.b[1] = (uint64_t)(b >> 64);
#endif
  o.prob = prob;
  return o;
}

GD uint64_t order_key_desc(double w) {
  w = w + 0.0;
  uint64_t u = dbits(w);
  u = (u >> 63) ? ~u : (u | (1ull << 63));
  return ~u;
}

struct Ctx {
  Grp g;
  Shared* sh;
  TablesView T;
  const GrimbConfig* cfg;
  const double* ones;  // [P*P] all-ones prior
  // arena (bump allocator; every thread mirrors the same offsets)
  char* ar_base;
  uint64_t ar_cap, ar_used;
  bool ws_fail;
  // subject
  int n;             // typed loci
  int loc[MAXL];     // locus index of typed position t
  uint32_t typed;    // label mask of the typed loci
  uint32_t full;     // label mask of all loci
  const double* M;   // current prior matrix [P][P]
  int nph;           // kept phases (sh->ph)
  SlotDesc* slots;   // [2*nph]
  TopItem* top;      // [2*nph][K]
  uint32_t* top_n;   // [2*nph]
  uint32_t* slot_ne; // [2*nph] last evaluation of the slot returned a non-empty dict
  int K;
  // selector
  SelItem* sel;
  SelItem* sel2;
  uint32_t* sel_idx;
  uint32_t capsel;
  bool thr_on;
  uint64_t thr;
  // entries of the current evaluation
  Entry* ent;
  uint32_t ent_cap, ent_n;
  uint64_t pair_evals;
  mutable uint32_t c_probes, c_hits;   // per-thread: hash probes issued / answered with a node (metric numerators)
  uint64_t c_vecs;             // group-wide (thread 0): frequency vectors the probes' expansions read
  bool plan_c_single;  // P_eff = 1 (Plan C: vectors summed over populations)

  template <class X>
  GD X* alloc(uint64_t count) {
    uint64_t bytes = (count * sizeof(X) + 15ull) & ~15ull;
    if (ar_used + bytes > ar_cap) {
      ws_fail = true;
      return reinterpret_cast<X*>(ar_base);
    }
    X* p = reinterpret_cast<X*>(ar_base + ar_used);
    ar_used += bytes;
    return p;
  }

  GD int side_of(int slot, int t) const { return ((sh->ph[slot >> 1] >> t) & 1) ^ (slot & 1); }

  // ---------------------------------------------------------------- candidates
  GD void decode(const SlotDesc& sd, int slot, uint64_t c, uint16_t* ids) const {
    if (sd.mode) {
      hkey k = T.node_key[sd.filt[c]];
      for (int t = 0; t < n; ++t) ids[t] = (uint16_t)key_field(T, k, loc[t]);
    } else {
      for (int t = n - 1; t >= 0; --t) {
        int x = side_of(slot, t);
        uint32_t cn = sh->lcnt[sd.var][t][x];
        uint32_t d = (uint32_t)(c % cn);
        c /= cn;
        ids[t] = sh->lptr[sd.var][t][x][d];
      }
    }
  }

  // pack the ids at the typed positions selected by posmask; false if an id is not a table allele
  GD bool pack(const uint16_t* ids, uint32_t posmask, hkey& key, uint32_t& label) const {
    key = 0;
    label = 0;
    bool known = true;
    for (int t = 0; t < n; ++t)
      if (posmask >> t & 1u) {
        int l = loc[t];
        if (ids[t] > T.n_alleles[l] || ids[t] == 0) known = false;
        key |= (hkey)ids[t] << T.shift[l];
        label |= 1u << l;
      }
    return known;
  }

  // ---------------------------------------------------------------- streaming stable top-K
  GD void sel_begin() {
    g.sync();
    if (g.tid == 0) {
      sh->sel_n = 0;
      sh->nonempty = 0;
    }
    thr_on = false;
    thr = 0;
    g.sync();
  }

  GDN void sel_compact() {
    uint32_t cnt = sh->sel_n;
    for (uint32_t i = g.tid; i < cnt; i += g.n) sel_idx[i] = i;
    g.sync();
    SelItem* s = sel;
    uint32_t* ix = sel_idx;
    group_sort(
        g, cnt,
        [=](uint32_t a, uint32_t b) {
          const SelItem& x = s[ix[a]];
          const SelItem& y = s[ix[b]];
          return x.wkey < y.wkey || (x.wkey == y.wkey && x.ord < y.ord);
        },
        [=](uint32_t a, uint32_t b) {
          uint32_t t = ix[a];
          ix[a] = ix[b];
          ix[b] = t;
        });
    uint32_t keep = cnt < (uint32_t)K ? cnt : (uint32_t)K;
    for (uint32_t i = g.tid; i < keep; i += g.n) sel2[i] = sel[sel_idx[i]];
    g.sync();
    SelItem* tmp = sel;
    sel = sel2;
    sel2 = tmp;
    if (g.tid == 0) sh->sel_n = keep;
    if (cnt >= (uint32_t)K) {
      thr_on = true;
      thr = sel[K - 1].wkey;
    }
    g.sync();
  }

  // call before every batch of at most g.n * per_thread pushes (uniform)
  GD void sel_reserve(uint32_t per_thread = 1) {
    g.sync();
    if (sh->sel_n + (uint32_t)g.n * per_thread > capsel) sel_compact();
  }

  GD void sel_push(double w, uint64_t ord, double f, hkey hap, uint32_t pop) {
    uint64_t wk = order_key_desc(w);
    if (thr_on && !(wk < thr)) return;
    uint32_t p = atom_add(&sh->sel_n, 1u);
    SelItem it;
    it.wkey = wk;
    it.ord = ord;
    it.f = f;
    it.hap = hap;
    it.pop = pop;
    it.pad = 0;
    sel[p] = it;
  }

  GD void sel_finish(int slot) {
    g.sync();
    sel_compact();
    uint32_t cnt = sh->sel_n;
    TopItem* dst = top + (uint64_t)slot * K;
    for (uint32_t i = g.tid; i < cnt; i += g.n) {
      TopItem t;
      t.f = sel[i].f;
      t.hap = sel[i].hap;
      t.pop = sel[i].pop;
      t.pad = 0;
      dst[i] = t;
    }
    if (g.tid == 0) {
      top_n[slot] = cnt;
      slot_ne[slot] = sh->nonempty;
    }
    g.sync();
  }

  // Expand a chunk of (node, degree) pairs held in sh->chunk_* into selector items:
  // node itself when `self`, else its CSR neighbours `adj[start[node]+t]`; hap key = node key |
  // extra[owner]; frequency scaled by `scale` when scaled.
  template <class Extra>
  GD void expand_chunk(uint32_t total, uint64_t base, bool self, const uint32_t* start, const uint32_t* adj,
                       Extra extra, bool scaled, double scale) {
    const int P = T.P;
    uint64_t items = (uint64_t)total * (uint64_t)P;
    if (g.tid == 0) c_vecs += total;
    for (uint64_t q0 = 0; q0 < items; q0 += g.n) {
      sel_reserve();
      uint64_t q = q0 + g.tid;
      if (q < items) {
        uint32_t hit = (uint32_t)(q / P);
        uint32_t j = (uint32_t)(q % P);
        int lo = 0, hi = g.n - 1;  // largest i with chunk_off[i] <= hit
        while (lo < hi) {
          int mid = (lo + hi + 1) >> 1;
          if (sh->chunk_off[mid] <= hit) lo = mid; else hi = mid - 1;
        }
        uint32_t node = sh->chunk_node[lo];
        uint32_t t = hit - sh->chunk_off[lo];
        uint32_t fn = self ? node : adj[start[node] + t];
        double f = T.freq[(uint64_t)fn * P + j];
        if (scaled) f = f * scale;
        if (f > 0) {
          double w = f * M[j * P + j];
          sel_push(w, ((base + (uint64_t)lo) << 32) | ((uint64_t)t * P + j), f, T.node_key[fn] | extra(lo), j);
        }
      }
    }
    g.sync();
  }

  // Plan A probe of one side (adjs_query): full-label node -> itself, partial -> top links.
  GDN void slot_plan_a(int slot) {
    const SlotDesc sd = slots[slot];
    sel_begin();
    const bool isfull = (typed == full);
    for (uint64_t base = 0; base < sd.ncand; base += g.n) {
      uint64_t c = base + g.tid;
      uint32_t node = GRIMB_NONE, deg = 0;
      if (c < sd.ncand) {
        if (sd.mode) {
          node = sd.filt[c];
        } else {
          uint16_t ids[MAXL];
          decode(sd, slot, c, ids);
          hkey key;
          uint32_t label;
          if (pack(ids, (1u << n) - 1u, key, label)) {
            node = ht_lookup(T, label, key);
            ++c_probes;
            if (node != GRIMB_NONE) ++c_hits;
          }
        }
        if (node != GRIMB_NONE) {
          deg = isfull ? 1u : T.tl_cnt[node];
          if (deg == GRIMB_ADJ_FAULT) {
            sh->fault = 1;
            deg = 0;
          }
        }
      }
      uint32_t total;
      uint32_t off = g.scan_excl(deg, total);
      sh->chunk_node[g.tid] = node;
      sh->chunk_off[g.tid] = off;
      if (total && g.tid == 0) sh->nonempty = 1;
      g.sync();
      expand_chunk(total, base, isfull, T.tl_start, T.tl_adj, [](int) { return (hkey)0; }, false, 1.0);
    }
    sel_finish(slot);
  }

  // ---------------------------------------------------------------- pair evaluation
  // accept test of calc_haps_pairs (impute.py:458-491) for one (h, k) pair; fills e when accepted
  GD bool pair_accept(const TopItem& a, const TopItem& b, double eps, int P, Entry& e) const {
    const double x = eps / a.f;
    const double m = M[a.pop * P + b.pop];
    if (!(m > 0)) return false;
    const double mf = m * b.f;
    const bool same = a.hap == b.hap;
    if (!((!same && mf >= x) || (same && mf >= x * 2))) return false;
    e.h1 = a.hap;
    e.h2 = b.hap;
    e.p1 = (uint16_t)a.pop;
    e.p2 = (uint16_t)b.pop;
    e.pad = 0;
    double pr = a.f * b.f * m;
    if (!same) pr = pr * 2;
    e.prob = pr;
    return true;
  }

  // Appends the accepted pairs of every opened phase, in (phase, h, k) order, to ent[].
  // All pairs of all phases form one index space; every thread owns a contiguous range of it, so
  // thread order == encounter order and one scan over the per-thread counts places the entries
  // (pass 1 counts, pass 2 writes): a handful of barriers per evaluation instead of one scan per
  // chunk of pairs.
  GDN void gen_entries(double eps) {
    const int P = plan_c_single ? 1 : T.P;
    ent_n = 0;
    const uint64_t mark = ar_used;
    uint32_t* ph_id = alloc<uint32_t>(nph);
    uint64_t* ph_off = alloc<uint64_t>(nph + 1);
    uint32_t* row_off = alloc<uint32_t>(nph + 1);
    if (ws_fail) return;
    g.sync();
    if (g.tid == 0) {
      uint32_t np = 0, roff = 0;
      uint64_t off = 0;
      for (int p = 0; p < nph; ++p) {
        if (!slots[2 * p].valid) continue;
        const uint32_t n1 = top_n[2 * p], n2 = top_n[2 * p + 1];
        if (n1 == 0 || n2 == 0) continue;
        ph_id[np] = (uint32_t)p;
        ph_off[np] = off;
        row_off[np] = roff;
        off += (uint64_t)n1 * n2;
        roff += n1;
        ++np;
      }
      ph_off[np] = off;
      row_off[np] = roff;
      sh->cnt[0] = np;
    }
    g.sync();
    const uint32_t np = sh->cnt[0];
    if (np == 0) {
      ar_used = mark;
      return;
    }
    const uint64_t total_pairs = ph_off[np];
    const uint32_t total_rows = row_off[np];
    uint16_t* kb = alloc<uint16_t>(total_rows);
    if (ws_fail) return;
    // the `break` of the k loop: first k with f2 < eps / f1 (impute.py:464,545-546)
    uint64_t evals = 0;
    for (uint32_t r = g.tid; r < total_rows; r += g.n) {
      uint32_t lo = 0, hi = np - 1;
      while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (row_off[mid] <= r) lo = mid; else hi = mid - 1;
      }
      const uint32_t p = ph_id[lo], h = r - row_off[lo];
      const uint32_t n2 = top_n[2 * p + 1];
      const TopItem* T2 = top + (uint64_t)(2 * p + 1) * K;
      const double x = eps / top[(uint64_t)(2 * p) * K + h].f;
      uint32_t k = 0;
      while (k < n2 && T2[k].f >= x) ++k;
      kb[r] = (uint16_t)k;
      evals += (k < n2) ? k + 1 : n2;
    }
    pair_evals += evals;  // per-thread partial; reduced at the end of the subject
    g.sync();
    const uint64_t per = (total_pairs + g.n - 1) / g.n;
    uint64_t q0 = per * g.tid, q1 = q0 + per;
    if (q0 > total_pairs) q0 = total_pairs;
    if (q1 > total_pairs) q1 = total_pairs;
    uint32_t my_cnt = 0, my_off = 0;
    for (int pass = 0; pass < 2; ++pass) {
      if (q0 < q1) {
        uint32_t lo = 0, hi = np - 1;
        while (lo < hi) {
          uint32_t mid = (lo + hi + 1) >> 1;
          if (ph_off[mid] <= q0) lo = mid; else hi = mid - 1;
        }
        uint32_t j = lo, p = ph_id[j];
        uint32_t n1 = top_n[2 * p], n2 = top_n[2 * p + 1];
        const TopItem* T1 = top + (uint64_t)(2 * p) * K;
        const TopItem* T2 = top + (uint64_t)(2 * p + 1) * K;
        uint64_t rem = q0 - ph_off[j];
        uint32_t h = (uint32_t)(rem / n2), k = (uint32_t)(rem % n2);
        uint32_t w = my_off;
        uint64_t q = q0;
        while (q < q1) {
          const uint32_t kbreak = kb[row_off[j] + h];
          if (k < kbreak) {
            Entry e;
            if (pair_accept(T1[h], T2[k], eps, P, e)) {
              if (pass == 0) ++my_cnt;
              else {
                if (w < ent_cap) ent[w] = e;
                ++w;
              }
            }
            ++q;
            ++k;
          } else {
            const uint64_t skip = (uint64_t)(n2 - k);  // the rest of this h row is past the break
            q += skip;
            k = n2;
          }
          if (k >= n2) {
            k = 0;
            if (++h >= n1) {
              h = 0;
              if (++j < np) {
                p = ph_id[j];
                n1 = top_n[2 * p];
                n2 = top_n[2 * p + 1];
                T1 = top + (uint64_t)(2 * p) * K;
                T2 = top + (uint64_t)(2 * p + 1) * K;
              } else {
                break;
              }
            }
          }
        }
      }
      if (pass == 0) {
        uint32_t total;
        my_off = g.scan_excl(my_cnt, total);
        ent_n = total;
        if (total == 0) break;
      }
    }
    g.sync();
    if (ent_n > ent_cap) ws_fail = true;
    ar_used = mark;
  }

  // geno_seen (impute.py:508-513): keep the first entry of every unordered {(hap,pop),(hap,pop)}.
  // Compacts ent[] in place (order preserved); returns MaxProb over the kept entries.
  GDN double dedup_entries() {
    if (ent_n == 0 || ws_fail) return 0.0;
    uint64_t mark = ar_used;
    uint32_t tsz = 2;
    while (tsz < 2 * ent_n) tsz <<= 1;
    uint32_t* tab = alloc<uint32_t>(tsz);
    uint32_t* where = alloc<uint32_t>(ent_n);
    Entry* tmp = alloc<Entry>(ent_n);
    if (ws_fail) return 0.0;
    for (uint32_t i = g.tid; i < tsz; i += g.n) tab[i] = GRIMB_NONE;
    g.sync();
    const Entry* E = ent;
    auto canon = [=](uint32_t i, hkey& ha, uint32_t& pa, hkey& hb, uint32_t& pb) {
      const Entry& e = E[i];
      bool sw = e.h1 > e.h2 || (e.h1 == e.h2 && e.p1 > e.p2);
      ha = sw ? e.h2 : e.h1;
      pa = sw ? e.p2 : e.p1;
      hb = sw ? e.h1 : e.h2;
      pb = sw ? e.p1 : e.p2;
    };
    for (uint32_t i = g.tid; i < ent_n; i += g.n) {
      hkey ha, hb;
      uint32_t pa, pb;
      canon(i, ha, pa, hb, pb);
      uint32_t h = (uint32_t)mix64(fold_key(ha) ^ mix64(fold_key(hb) + 0x9e3779b97f4a7c15ULL) ^ ((uint64_t)pa << 17) ^ ((uint64_t)pb << 41)) & (tsz - 1);
      for (;;) {
        uint32_t cur = tab[h];
        if (cur == GRIMB_NONE) {
          cur = atom_cas(&tab[h], GRIMB_NONE, i);
          if (cur == GRIMB_NONE) break;
        }
        hkey xa, xb;
        uint32_t qa, qb;
        canon(cur, xa, qa, xb, qb);
        if (xa == ha && xb == hb && qa == pa && qb == pb) {
          atom_min(&tab[h], i);
          break;
        }
        h = (h + 1) & (tsz - 1);
      }
      where[i] = h;
    }
    g.sync();
    double mx = 0.0;
    uint32_t kept = 0;
    for (uint32_t i0 = 0; i0 < ent_n; i0 += g.n) {
      uint32_t i = i0 + g.tid;
      bool keep = i < ent_n && tab[where[i]] == i;
      uint32_t total;
      uint32_t pos = g.scan_excl(keep ? 1u : 0u, total);
      if (keep) {
        tmp[kept + pos] = ent[i];
        if (ent[i].prob > mx) mx = ent[i].prob;
      }
      kept += total;
    }
    g.sync();
    for (uint32_t i = g.tid; i < kept; i += g.n) ent[i] = tmp[i];
    ent_n = kept;
    mx = g.maxd(mx);
    g.sync();
    ar_used = mark;
    return mx;
  }

  // ---------------------------------------------------------------- aggregation + top-N
  // kind 0: UMUG genotype (per-locus unordered pairs)   impute.py:497-504,529-533
  // kind 1: PMUG haplotype pair, first-seen orientation  impute.py:24-38
  // kind 2: population pair                               impute.py:535-543 / :24-38
  GD void group_key(int kind, const Entry& e, hkey& a, hkey& b) const {
    if (kind == 0) {
      hkey lo = 0, hi = 0;
      for (int l = 0; l < T.L; ++l) {
        uint64_t x = key_field(T, e.h1, l), y = key_field(T, e.h2, l);
        uint64_t mn = x < y ? x : y, mxv = x < y ? y : x;
        lo |= (hkey)mn << T.shift[l];
        hi |= (hkey)mxv << T.shift[l];
      }
      a = lo;
      b = hi;
    } else if (kind == 1) {
      a = e.h1 < e.h2 ? e.h1 : e.h2;
      b = e.h1 < e.h2 ? e.h2 : e.h1;
    } else {
      a = e.p1 < e.p2 ? e.p1 : e.p2;
      b = e.p1 < e.p2 ? e.p2 : e.p1;
    }
  }

  // Groups ent[0..ent_n) by `kind`, sums probabilities per group in encounter order, ranks the
  // groups by (sum desc, first encounter asc) and writes the best `limit` rows.  Returns the
  // number of groups; *n_rows = rows written.  Rows: kind 0 -> (lo,hi); kind 1 -> (h1,h2) of
  // the first member; kind 2 -> pops of the first member (orientation as encountered).
  //
  // No sort of the entries: a hash table (atomicMin on the entry index) names each group by its
  // first member; groups are numbered in first-encounter order by a scan over the heads; then
  // every thread walks ALL entries in encounter order and adds those whose group it owns
  // (group id mod #threads), so each group's += chain runs in the reference's order without
  // ordering anything.  Top-N is a repeated block arg-max when N is small, a bitonic sort of the
  // group list otherwise.
  GDN uint32_t aggregate(int kind, uint32_t limit, GrimbHapRow* hap_rows, GrimbPopRow* pop_rows, uint32_t* n_rows) {
    if (g.tid == 0) *n_rows = 0;
    if (ent_n == 0 || ws_fail) return 0;
    uint64_t mark = ar_used;
    uint32_t tsz = 2;
    while (tsz < 2 * ent_n) tsz <<= 1;
    uint32_t* tab = alloc<uint32_t>(tsz);        // slot -> first member (entry index)
    uint32_t* slot_gid = alloc<uint32_t>(tsz);   // slot -> group id
    uint32_t* where = alloc<uint32_t>(ent_n);    // entry -> slot, later entry -> group id
    uint32_t* ghead = alloc<uint32_t>(ent_n);    // group -> first member
    double* gsum = alloc<double>(ent_n);
    uint32_t* gord = alloc<uint32_t>(ent_n);
    uint8_t* own = alloc<uint8_t>((uint64_t)ent_n + 16);   // entry -> owning thread of its group (0xFF: group head)
    if (ws_fail) return 0;
    for (uint32_t i = g.tid; i < tsz; i += g.n) tab[i] = GRIMB_NONE;
    g.sync();
    const Entry* E = ent;
    for (uint32_t i = g.tid; i < ent_n; i += g.n) {
      hkey a, b;
      group_key(kind, E[i], a, b);
      uint32_t h = (uint32_t)mix64(fold_key(a) ^ mix64(fold_key(b) + 0x9e3779b97f4a7c15ULL)) & (tsz - 1);
      for (;;) {
        uint32_t cur = tab[h];
        if (cur == GRIMB_NONE) {
          cur = atom_cas(&tab[h], GRIMB_NONE, i);
          if (cur == GRIMB_NONE) break;
        }
        hkey xa, xb;
        group_key(kind, E[cur], xa, xb);
        if (xa == a && xb == b) {
          atom_min(&tab[h], i);
          break;
        }
        h = (h + 1) & (tsz - 1);
      }
      where[i] = h;
    }
    g.sync();
    // heads in encounter order -> group ids
    uint32_t ng = 0;
    for (uint32_t i0 = 0; i0 < ent_n; i0 += g.n) {
      uint32_t i = i0 + g.tid;
      bool head = i < ent_n && tab[where[i]] == i;
      uint32_t total;
      uint32_t pos = g.scan_excl(head ? 1u : 0u, total);
      if (head) {
        ghead[ng + pos] = i;
        slot_gid[where[i]] = ng + pos;
        gsum[ng + pos] = E[i].prob;
        gord[ng + pos] = ng + pos;
      }
      ng += total;
    }
    g.sync();
    for (uint32_t i = g.tid; i < ent_n; i += g.n) where[i] = slot_gid[where[i]];
    g.sync();
    if (ng < ent_n) {
      // += in encounter order: thread t owns the groups with (id & tmask) == t and walks ALL entries in
      // order, adding the ones it owns.  The walk reads one owner byte per entry, 16 entries per load and
      // 4 per SIMD compare, so the scan costs ~0.6 instructions per entry and thread instead of 3.
      uint32_t tpow = 1;
      while (tpow * 2 <= (uint32_t)g.n && tpow * 2 <= 128u) tpow <<= 1;   // owner ids stay below the 0xFF marker
      const uint32_t tmask = tpow - 1;
      for (uint32_t i = g.tid; i < ent_n + 16u; i += g.n)
        own[i] = (i < ent_n && ghead[where[i]] != i) ? (uint8_t)(where[i] & tmask) : (uint8_t)0xFF;
      g.sync();
      if ((uint32_t)g.tid < tpow) {
        const uint32_t me = (uint32_t)g.tid;
#if GRIMB_DEVICE
        const uint32_t me4 = me * 0x01010101u;
        for (uint32_t i0 = 0; i0 < ent_n; i0 += 16) {
          const uint4 v = *reinterpret_cast<const uint4*>(own + i0);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t m = __vcmpeq4(w[k], me4);   // 0xFF in every byte that names this thread
            while (m) {
              const int b = (__ffs(m) - 1) >> 3;   // lowest byte = lowest entry index first
              const uint32_t i = i0 + 4u * k + (uint32_t)b;
              const uint32_t gi = where[i];
              gsum[gi] = gsum[gi] + E[i].prob;
              m &= ~(0xFFu << (8 * b));
            }
          }
        }
#else
        for (uint32_t i = 0; i < ent_n; ++i)
          if (own[i] == (uint8_t)me) {
            const uint32_t gi = where[i];
            gsum[gi] = gsum[gi] + E[i].prob;
          }
#endif
      }
      g.sync();
    }
    uint32_t rows = ng < limit ? ng : limit;
    if (cfg->encounter_order) {
      // group ids ARE the first-encounter order (the insertion order of the reference's dicts)
      for (uint32_t r = g.tid; r < rows; r += g.n) slot_gid[r] = r;
      g.sync();
    } else if (rows <= 64 && ng > 1024) {
      // repeated arg-max over (sum desc, group id asc); taken groups are marked in gord
      for (uint32_t r = 0; r < rows; ++r) {
        double best = -1.0;
        uint32_t bi = GRIMB_NONE;
        for (uint32_t gi = g.tid; gi < ng; gi += g.n)
          if (gord[gi] != GRIMB_NONE) {
            const double v = gsum[gi];
            if (v > best) {   // strided scan visits ids in ascending order: first maximum wins
              best = v;
              bi = gi;
            }
          }
        const double top = g.maxd(best);
        uint32_t cand = (bi != GRIMB_NONE && best == top) ? bi : GRIMB_NONE;
        // smallest group id among the threads holding the maximum
        g.sync();
        if (g.tid == 0) sh->cnt[0] = GRIMB_NONE;
        g.sync();
        if (cand != GRIMB_NONE) atom_min(&sh->cnt[0], cand);
        g.sync();
        const uint32_t win = sh->cnt[0];
        g.sync();
        if (g.tid == 0) {
          gord[win] = GRIMB_NONE;
          slot_gid[r] = win;     // slot_gid is free now: reuse as the ranked list
        }
        g.sync();                // the next sweep must see the winner marked as taken
      }
    } else {
      const double* gs = gsum;
      uint32_t* go = gord;
      if (ng > rows || ng > 1)
        group_sort(
            g, ng,
            [=](uint32_t a, uint32_t b) {
              double x = gs[go[a]], y = gs[go[b]];
              return x > y || (x == y && go[a] < go[b]);
            },
            [=](uint32_t a, uint32_t b) {
              uint32_t t = go[a];
              go[a] = go[b];
              go[b] = t;
            });
      for (uint32_t r = g.tid; r < rows; r += g.n) slot_gid[r] = gord[r];
      g.sync();
    }
    for (uint32_t r = g.tid; r < rows; r += g.n) {
      uint32_t gi = slot_gid[r];
      const Entry& e = E[ghead[gi]];
      if (kind == 2) {
        GrimbPopRow o;
        o.pop_a = e.p1;
        o.pop_b = e.p2;
        o.pad = 0;
        o.prob = gsum[gi];
        pop_rows[r] = o;
      } else {
        hkey ka = e.h1, kb = e.h2;
        if (kind == 0) group_key(0, e, ka, kb);
        hap_rows[r] = make_hap_row(ka, kb, gsum[gi]);
      }
    }
    if (g.tid == 0) *n_rows = rows;
    g.sync();
    ar_used = mark;
    return ng;
  }
};

}  // namespace grimb
