// grimb_plan.h -- per-subject control flow on top of grimb_subject.h: opening, reductions,
// epsilon schedule, Plan B / Plan C side evaluation, emission.  See grimb_subject.h for the
// reference line map ("impute.py" = grim/imputation/impute.py of the reference).
#pragma once
#include "grimb_subject.h"

namespace grimb {


constexpr uint32_t SEL_IPT = 8;  // items per thread between two selector barriers in the streaming loops
constexpr uint64_t PRE_MIN_CANDIDATES = 2048;   // cooperative slot pass: smallest Cartesian product worth a CTA

struct BlockList {
  const uint32_t* ids;  // nullptr: implicit node-id range
  uint32_t first;
  uint32_t n;
};

// Plan A side lists computed ahead of the subject's own CTA by the cooperative slot kernel
// (k_impute_slots in grimb200.cu): for the heaviest subjects of a batch every (phase, side) "slot" is
// opened, probed and reduced to its top-K list by a CTA of its own, so that up to 2^L CTAs work on one
// subject at once; the subject's CTA then starts from the finished lists.
struct PreView {
  TopItem* top;            // [max_subjects][slots_per_subject][K]
  uint32_t* n;             // [max_subjects][slots_per_subject] list length
  uint32_t* ne;            // ... the probe returned anything (sh->nonempty)
  uint32_t* ready;         // ... 1: list valid (slot opened in its original variant, no workspace overflow)
  uint32_t max_subjects;
  uint32_t slots_per_subject;   // 2^L
  uint32_t K;
};

struct Subject : Ctx {
  const PreView* pre;   // nullptr: none
  uint32_t pre_j;       // this subject's index in the pre-computed lists (0xFFFFFFFF: none)
  // allele-membership bitmasks of the subject's lists, per list variant (arena; built on first use by the
  // over-threshold filter): bit id of the mask of (typed position t, side x) = table allele id is listed
  uint32_t* lmask[NVAR];
  uint32_t lmask_off[MAXL];   // first word of position t's two masks
  uint32_t lmask_w[MAXL];     // words per mask of position t
  hkey* chunk_extra;  // [g.n] arena: key bits re-inserted by the missing-data path
  GrimbHapRow* st_hap[2];
  GrimbPopRow* st_pop[2];
  GrimbPopRow* st_pair;   // hap_pop_pair mode: populations of the staged PMUG rows
  uint32_t* st_cnt;       // [4] arena: rows staged (umug, pmug, umug pops, pmug pops)
  const double* Msubj;

  // Python's sum() over a list of floats: plain adds before CPython 3.12, Neumaier-compensated
  // from 3.12 on (bltinmodule.c); the first element is taken as is (0 + x).
  GD double py_sum(const double* v, int cnt) const {
    double f = 0.0 + v[0];
    if (!cfg->compensated_sum) {
      for (int i = 1; i < cnt; ++i) f = f + v[i];
      return f;
    }
    double c = 0.0;
    for (int i = 1; i < cnt; ++i) {
      double x = v[i];
      double t = f + x;
      double af = f < 0 ? -f : f, ax = x < 0 ? -x : x;
      if (af >= ax) c += (f - t) + x;
      else c += (x - t) + f;
      f = t;
    }
    if (c != 0.0 && c - c == 0.0) f += c;
    return f;
  }

  // ------------------------------------------------------------------ lists / variants
  GD uint64_t slot_options(int slot, int var) const {
    uint64_t o = 1;
    for (int t = 0; t < n; ++t) {
      uint64_t c = sh->lcnt[var][t][side_of(slot, t)];
      o = (o > (1ull << 40)) ? o : o * c;
    }
    return o;
  }

  GD bool allele_node(int l, uint32_t id, uint32_t& node) const {
    if (id == 0 || id > T.n_alleles[l]) return false;
    node = ht_lookup(T, 1u << l, (hkey)id << T.shift[l]);
    ++c_probes;
    if (node != GRIMB_NONE) ++c_hits;
    return node != GRIMB_NONE;
  }

  // lexist[var][t][x]: does any allele of the list exist as a single-locus node (impute.py:1218-1241)
  GD void compute_exist(int var) {
    g.sync();
    for (int q = g.tid; q < n * 2; q += g.n) {
      int t = q >> 1, x = q & 1;
      const uint16_t* lst = sh->lptr[var][t][x];
      bool any = false;
      uint32_t node;
      for (uint32_t i = 0; i < sh->lcnt[var][t][x] && !any; ++i) any = allele_node(loc[t], lst[i], node);
      sh->lexist[var][t][x] = any ? 1 : 0;
    }
    g.sync();
  }

  // reduce_phase_to_valid_allels :864-879 (VALID), reduce_phase_to_commons_alleles :881-912
  // (C10: ten best, C1: the best) -- a function of the original list and the prior diagonal.
  GDN void make_variant(int var) {
    lmask[var] = nullptr;   // the lists of this variant are about to change
    uint32_t total = 0;
    for (int t = 0; t < n; ++t) total += sh->lcnt[VAR_ORIG][t][0] + sh->lcnt[VAR_ORIG][t][1];
    uint16_t* buf = alloc<uint16_t>(total);
    double* score = alloc<double>(total);
    if (ws_fail) return;
    const int P = T.P;
    g.sync();
    for (int q = g.tid; q < n * 2; q += g.n) {
      int t = q >> 1, x = q & 1;
      uint32_t off = 0;
      for (int tt = 0; tt < t; ++tt) off += sh->lcnt[VAR_ORIG][tt][0] + sh->lcnt[VAR_ORIG][tt][1];
      if (x) off += sh->lcnt[VAR_ORIG][t][0];
      const uint16_t* src = sh->lptr[VAR_ORIG][t][x];
      uint32_t cn = sh->lcnt[VAR_ORIG][t][x];
      uint16_t* dst = buf + off;
      double* sc = score + off;
      uint32_t m = 0;
      for (uint32_t i = 0; i < cn; ++i) {
        uint32_t node;
        if (allele_node(loc[t], src[i], node)) {
          double s = 0;
          for (int p = 0; p < P; ++p) s += T.freq[(uint64_t)node * P + p] * M[p * P + p];
          dst[m] = src[i];
          sc[m] = s;
          ++m;
        }
      }
      if (m == 0) {  // nothing exists: the reference leaves the string unchanged
        sh->lptr[var][t][x] = src;
        sh->lcnt[var][t][x] = (uint16_t)cn;
      } else {
        uint32_t keep = m;
        if (var == VAR_C10 || var == VAR_C1) {
          keep = (var == VAR_C1) ? 1u : (m < 10u ? m : 10u);
          // stable descending order: repeatedly take the first maximum of the remainder
          for (uint32_t r = 0; r < keep; ++r) {
            uint32_t best = r;
            for (uint32_t i = r + 1; i < m; ++i)
              if (sc[i] > sc[best]) best = i;
            uint16_t a = dst[best];
            double s = sc[best];
            for (uint32_t i = best; i > r; --i) {
              dst[i] = dst[i - 1];
              sc[i] = sc[i - 1];
            }
            dst[r] = a;
            sc[r] = s;
          }
        }
        sh->lptr[var][t][x] = dst;
        sh->lcnt[var][t][x] = (uint16_t)keep;
      }
    }
    g.sync();
    compute_exist(var);
  }

  // ------------------------------------------------------------------ opening (open_phases)
  // every allele of the node is in the slot's lists (cutils.create_hap_list, cutils.pyx:35-51): one bit test
  // per locus against the membership masks of the variant (ensure_masks)
  GD bool node_in_lists(uint32_t node, int slot, int var) const {
    const hkey k = T.node_key[node];
    const uint32_t* mk = lmask[var];
    for (int t = 0; t < n; ++t) {
      const uint32_t id = (uint32_t)key_field(T, k, loc[t]);
      const uint32_t* w = mk + lmask_off[t] + (uint32_t)side_of(slot, t) * lmask_w[t];
      if (!((w[id >> 5] >> (id & 31u)) & 1u)) return false;
    }
    return true;
  }

  // builds the membership masks of a list variant once per subject (uniform over the group)
  GDN void ensure_masks(int var) {
    if (lmask[var] != nullptr) return;
    uint32_t words = 0;
    for (int t = 0; t < n; ++t) {
      lmask_w[t] = (T.n_alleles[loc[t]] + 32u) >> 5;
      lmask_off[t] = words;
      words += 2u * lmask_w[t];
    }
    uint32_t* mk = alloc<uint32_t>(words ? words : 1);
    if (ws_fail) return;
    for (uint32_t i = g.tid; i < words; i += g.n) mk[i] = 0;
    g.sync();
    for (int t = 0; t < n; ++t)
      for (int x = 0; x < 2; ++x) {
        const uint16_t* lst = sh->lptr[var][t][x];
        const uint32_t cn = sh->lcnt[var][t][x];
        uint32_t* w = mk + lmask_off[t] + (uint32_t)x * lmask_w[t];
        for (uint32_t i = g.tid; i < cn; i += g.n) {
          const uint32_t id = lst[i];
          if (id >= 1 && id <= T.n_alleles[loc[t]]) atom_or(&w[id >> 5], 1u << (id & 31u));
        }
      }
    g.sync();
    lmask[var] = mk;
  }

  // Returns the number of opened phases.
  // opens one (phase, side): Cartesian mode below the options threshold, else the filter of the typed
  // label's nodes (open_phases, impute.py:914-989)
  GDN void open_slot(int slot) {
    const uint64_t thr_opt = (uint64_t)cfg->options_threshold;
    SlotDesc sd = slots[slot];
    uint64_t opt = slot_options(slot, sd.var);
    if (opt < thr_opt) {
      g.sync();
      if (g.tid == 0) {
        sd.mode = 0;
        sd.ncand = opt;
        sd.filt = nullptr;
        slots[slot] = sd;
      }
    } else {
      ensure_masks(sd.var);
      if (ws_fail) return;
      const uint32_t first = T.label_first[typed], cnt = T.label_count[typed];
      uint32_t found = 0;
      for (uint32_t b = 0; b < cnt; b += g.n) {
        uint32_t i = b + g.tid;
        found += g.sum((i < cnt && node_in_lists(first + i, slot, sd.var)) ? 1u : 0u);
      }
      uint32_t* lst = alloc<uint32_t>(found ? found : 1);
      if (ws_fail) return;
      uint32_t w = 0;
      for (uint32_t b = 0; b < cnt; b += g.n) {
        uint32_t i = b + g.tid;
        bool in = i < cnt && node_in_lists(first + i, slot, sd.var);
        uint32_t total;
        uint32_t pos = g.scan_excl(in ? 1u : 0u, total);
        if (in) lst[w + pos] = first + i;
        w += total;
      }
      g.sync();
      if (g.tid == 0) {
        sd.mode = 1;
        sd.ncand = found;
        sd.filt = lst;
        slots[slot] = sd;
      }
    }
  }

  GDN int open_all() {
    g.sync();
    for (int slot = 0; slot < 2 * nph; ++slot) {
      open_slot(slot);
      if (ws_fail) return 0;
    }
    g.sync();
    int nvalid = 0;
    for (int p = 0; p < nph; ++p) {
      bool v = slots[2 * p].ncand > 0 && slots[2 * p + 1].ncand > 0;
      nvalid += v ? 1 : 0;
    }
    g.sync();
    if (g.tid == 0)
      for (int p = 0; p < nph; ++p) {
        bool v = slots[2 * p].ncand > 0 && slots[2 * p + 1].ncand > 0;
        slots[2 * p].valid = v;
        slots[2 * p + 1].valid = v;
      }
    g.sync();
    return nvalid;
  }

  // set the list variant of every slot whose current option count reaches the threshold
  // (or of every slot when `all`), as the reduce_* functions do phase by phase
  GD void set_variant(int var, bool all) {
    g.sync();
    if (g.tid == 0)
      for (int slot = 0; slot < 2 * nph; ++slot)
        if (all || slot_options(slot, slots[slot].var) >= (uint64_t)cfg->options_threshold) slots[slot].var = (uint8_t)var;
    g.sync();
  }

  // ------------------------------------------------------------------ Plan B side evaluation
  // One block of a matrix row -> ordered list of nodes (impute.py:1015-1039, 1207-1216,
  // networkx_graph.py:280-321).  Returns false if the side result must be empty.
  GDN bool build_block(int slot, uint32_t bm, bool first_block, BlockList& out) {
    const SlotDesc sd = slots[slot];
    const uint32_t tp = bm & typed, up = bm & ~typed;
    out.ids = nullptr;
    out.first = 0;
    out.n = 0;
    if (tp == 0) {  // no typed locus in the block: strings == []
      if (first_block) return false;
      out.first = T.label_first[bm];
      out.n = T.label_count[bm];
      return true;  // all nodes of the label (impute.py:1100-1105)
    }
    uint32_t posmask = 0;
    for (int t = 0; t < n; ++t)
      if (tp >> loc[t] & 1u) posmask |= 1u << t;
    const int nup = popc16(up);
    int uploc = 0;
    for (int l = 0; l < T.L; ++l)
      if (up >> l & 1u) uploc = l;
    const hkey km = key_mask_of(T, tp);
    // distinct block strings in first-occurrence order
    uint64_t ncomb;
    uint32_t* dd_first = nullptr;  // filter mode: candidate index of each distinct block string
    if (sd.mode == 0) {
      ncomb = 1;
      for (int t = 0; t < n; ++t)
        if (posmask >> t & 1u) ncomb *= sh->lcnt[sd.var][t][side_of(slot, t)];
    } else {
      // candidates are table nodes: dedup their projections (hash, min candidate index)
      uint32_t nc = (uint32_t)sd.ncand;
      uint32_t tsz = 2;
      while (tsz < 2 * nc) tsz <<= 1;
      uint32_t* tab = alloc<uint32_t>(tsz);
      uint32_t* where = alloc<uint32_t>(nc);
      dd_first = alloc<uint32_t>(nc);
      if (ws_fail) return false;
      for (uint32_t i = g.tid; i < tsz; i += g.n) tab[i] = GRIMB_NONE;
      g.sync();
      for (uint32_t c = g.tid; c < nc; c += g.n) {
        hkey k = T.node_key[sd.filt[c]] & km;
        uint32_t h = (uint32_t)mix64(fold_key(k)) & (tsz - 1);
        for (;;) {
          uint32_t cur = tab[h];
          if (cur == GRIMB_NONE) {
            cur = atom_cas(&tab[h], GRIMB_NONE, c);
            if (cur == GRIMB_NONE) break;
          }
          if ((T.node_key[sd.filt[cur]] & km) == k) {
            atom_min(&tab[h], c);
            break;
          }
          h = (h + 1) & (tsz - 1);
        }
        where[c] = h;
      }
      g.sync();
      uint32_t w = 0;
      for (uint32_t b = 0; b < nc; b += g.n) {
        uint32_t c = b + g.tid;
        bool head = c < nc && tab[where[c]] == c;
        uint32_t total;
        uint32_t pos = g.scan_excl(head ? 1u : 0u, total);
        if (head) dd_first[w + pos] = c;
        w += total;
      }
      g.sync();
      ncomb = w;
    }
    // two passes: count, then fill
    uint32_t* lst = nullptr;
    uint32_t total_out = 0;
    for (int pass = 0; pass < 2; ++pass) {
      uint32_t w = 0;
      for (uint64_t base = 0; base < ncomb; base += g.n) {
        uint64_t q = base + g.tid;
        uint32_t node = GRIMB_NONE, deg = 0;
        if (q < ncomb) {
          hkey key = 0;
          bool known = true;
          if (sd.mode == 0) {
            uint64_t c = q;
            for (int t = n - 1; t >= 0; --t)
              if (posmask >> t & 1u) {
                int x = side_of(slot, t);
                uint32_t cn = sh->lcnt[sd.var][t][x];
                uint16_t id = sh->lptr[sd.var][t][x][c % cn];
                c /= cn;
                if (id == 0 || id > T.n_alleles[loc[t]]) known = false;
                key |= (hkey)id << T.shift[loc[t]];
              }
          } else {
            key = T.node_key[sd.filt[dd_first[q]]] & km;
          }
          if (known) {
            node = ht_lookup(T, tp, key);
            ++c_probes;
            if (node != GRIMB_NONE) ++c_hits;
          }
          if (node != GRIMB_NONE) {
            if (nup == 0) deg = 1;
            else if (nup == 1) {
              deg = T.cn_cnt[(uint64_t)node * T.L + uploc];
              if (deg == GRIMB_ADJ_FAULT) {
                sh->fault = 1;
                deg = 0;
              }
            }
          }
        }
        uint32_t total;
        uint32_t off = g.scan_excl(deg, total);
        if (pass == 1 && deg) {
          if (nup == 0) lst[w + off] = node;
          else {
            uint32_t st = T.cn_start[(uint64_t)node * T.L + uploc];
            for (uint32_t i = 0; i < deg; ++i) lst[w + off + i] = T.cn_adj[st + i];
          }
        }
        w += total;
      }
      if (pass == 0) {
        total_out = w;
        if (total_out == 0) return false;  // D == {} with typed loci in the block: side empty
        lst = alloc<uint32_t>(total_out);
        if (ws_fail) return false;
      }
    }
    g.sync();
    out.ids = lst;
    out.n = total_out;
    return true;
  }

  GD uint32_t block_node(const BlockList& b, uint64_t i) const { return b.ids ? b.ids[i] : b.first + (uint32_t)i; }

  // save_space_mode (impute.py:1048-1059): an operand with more than 10 entries keeps the 10
  // with the largest sum over populations (ascending stable sort, delete from the front = the
  // ten largest by (sum, position)); original order preserved.  pe = vector length.
  GD uint32_t prune10(hkey* keys, double* vecs, uint32_t cnt, int pe) {
    if (cnt <= 10) return cnt;
    g.sync();
    if (g.tid == 0) {
      uint32_t best[10];
      double bsum[10];
      uint32_t nb = 0;
      for (uint32_t i = 0; i < cnt; ++i) {
        double s = py_sum(vecs + (uint64_t)i * pe, pe);
        // insert into the descending list; later position wins ties
        uint32_t pos = nb;
        while (pos > 0 && bsum[pos - 1] <= s) --pos;
        if (pos < 10) {
          uint32_t last = nb < 10 ? nb : 9;
          for (uint32_t q = last; q > pos; --q) {
            best[q] = best[q - 1];
            bsum[q] = bsum[q - 1];
          }
          best[pos] = i;
          bsum[pos] = s;
          if (nb < 10) ++nb;
        }
      }
      // keep in original order
      for (uint32_t a = 1; a < 10; ++a) {
        uint32_t v = best[a];
        int b = (int)a - 1;
        while (b >= 0 && best[b] > v) {
          best[b + 1] = best[b];
          --b;
        }
        best[b + 1] = v;
      }
      for (uint32_t w = 0; w < 10; ++w) {
        uint32_t i = best[w];
        keys[w] = keys[i];
        for (int p = 0; p < pe; ++p) vecs[(uint64_t)w * pe + p] = vecs[(uint64_t)i * pe + p];
      }
    }
    g.sync();
    return 10;
  }

  // find_option_freq (impute.py:1072-1115) for one side and one matrix row, streamed into the
  // top-K selector.
  GDN void slot_plan_b_row(int slot, int row) {
    const int P = T.P;
    const uint64_t mark = ar_used;
    const int nb = cfg->row_blocks[row];
    BlockList bl[GRIMB_MAX_BLOCKS];
    bool ok = true;
    for (int b = 0; b < nb && ok; ++b) {
      ok = build_block(slot, cfg->block_mask[row][b], b == 0, bl[b]);
      if (ws_fail) return;
      if (ok && bl[b].n == 0) ok = false;  // an all-untyped label without nodes: product empty
    }
    sel_begin();
    if (ok && !cfg->save_space) {
      uint64_t ntot = 1;
      for (int b = 0; b < nb; ++b) ntot = (ntot > (1ull << 44)) ? ntot : ntot * bl[b].n;
      if (ntot > (1ull << 44)) {
        ws_fail = true;
        return;
      }
      const uint64_t items = ntot * P;
      for (uint64_t q0 = 0; q0 < items; q0 += (uint64_t)g.n * SEL_IPT) {
        sel_reserve(SEL_IPT);   // one barrier per SEL_IPT items per thread
        for (uint32_t u = 0; u < SEL_IPT; ++u) {
        uint64_t q = q0 + (uint64_t)u * g.n + g.tid;
        if (q < items) {
          uint64_t c = q / P;
          uint32_t j = (uint32_t)(q % P);
          uint32_t nd[GRIMB_MAX_BLOCKS];
          for (int b = nb - 1; b >= 0; --b) {
            nd[b] = block_node(bl[b], c % bl[b].n);
            c /= bl[b].n;
          }
          double v = T.freq[(uint64_t)nd[0] * P + j];
          hkey hap = T.node_key[nd[0]];
          for (int b = 1; b < nb; ++b) {
            v = v * T.freq[(uint64_t)nd[b] * P + j] * 0.0001;
            hap |= T.node_key[nd[b]];
          }
          if (v > 0) {
            sh->nonempty = 1;
            sel_push(v * M[j * P + j], q, v, hap, j);
          }
        }
        }
      }
    } else if (ok) {
      // save_space_mode: materialise, pruning both operands to 10 entries before every product
      uint32_t cnt = bl[0].n;
      uint32_t cap0 = cnt > 100 ? cnt : 100;
      hkey* keys = alloc<hkey>(cap0);
      double* vecs = alloc<double>((uint64_t)cap0 * P);
      hkey* keys2 = alloc<hkey>(100);
      double* vecs2 = alloc<double>(100ull * P);
      if (ws_fail) return;
      for (uint32_t i = g.tid; i < cnt; i += g.n) {
        uint32_t nd = block_node(bl[0], i);
        keys[i] = T.node_key[nd];
        for (int p = 0; p < P; ++p) vecs[(uint64_t)i * P + p] = T.freq[(uint64_t)nd * P + p];
      }
      g.sync();
      for (int b = 1; b < nb && cnt > 0; ++b) {
        uint32_t nn = bl[b].n;
        hkey* nk = alloc<hkey>(nn);
        double* nv = alloc<double>((uint64_t)nn * P);
        if (ws_fail) return;
        for (uint32_t i = g.tid; i < nn; i += g.n) {
          uint32_t nd = block_node(bl[b], i);
          nk[i] = T.node_key[nd];
          for (int p = 0; p < P; ++p) nv[(uint64_t)i * P + p] = T.freq[(uint64_t)nd * P + p];
        }
        g.sync();
        cnt = prune10(keys, vecs, cnt, P);
        nn = prune10(nk, nv, nn, P);
        g.sync();
        if (g.tid == 0) {
          uint32_t w = 0;
          for (uint32_t a = 0; a < cnt; ++a)
            for (uint32_t c2 = 0; c2 < nn; ++c2) {
              bool pos = false;
              for (int p = 0; p < P; ++p) {
                double v = vecs[(uint64_t)a * P + p] * nv[(uint64_t)c2 * P + p] * 0.0001;
                vecs2[(uint64_t)w * P + p] = v;
                pos = pos || v > 0;
              }
              if (pos) {
                keys2[w] = keys[a] | nk[c2];
                ++w;
              }
            }
          for (uint32_t i = 0; i < w; ++i) {
            keys[i] = keys2[i];
            for (int p = 0; p < P; ++p) vecs[(uint64_t)i * P + p] = vecs2[(uint64_t)i * P + p];
          }
          sh->cnt[0] = w;
        }
        g.sync();
        cnt = sh->cnt[0];
        g.sync();
      }
      if (cnt && g.tid == 0) sh->nonempty = 1;
      const uint64_t items = (uint64_t)cnt * P;
      for (uint64_t q0 = 0; q0 < items; q0 += g.n) {
        sel_reserve();
        uint64_t q = q0 + g.tid;
        if (q < items) {
          double v = vecs[q];
          uint32_t j = (uint32_t)(q % P);
          if (v > 0) sel_push(v * M[j * P + j], q, v, keys[q / P], j);
        }
      }
    }
    sel_finish(slot);
    ar_used = mark;
  }

  // Expand sh->chunk_* through the connector CSR (or the node itself) into selector items.
  GD void expand_chunk_cn(uint32_t total, uint64_t base, bool self, int uploc, const hkey* extra, double scale) {
    const int P = T.P;
    uint64_t items = (uint64_t)total * (uint64_t)P;
    for (uint64_t q0 = 0; q0 < items; q0 += g.n) {
      sel_reserve();
      uint64_t q = q0 + g.tid;
      if (q < items) {
        uint32_t hit = (uint32_t)(q / P);
        uint32_t j = (uint32_t)(q % P);
        int lo = 0, hi = g.n - 1;
        while (lo < hi) {
          int mid = (lo + hi + 1) >> 1;
          if (sh->chunk_off[mid] <= hit) lo = mid; else hi = mid - 1;
        }
        uint32_t node = sh->chunk_node[lo];
        uint32_t t = hit - sh->chunk_off[lo];
        uint32_t fn = self ? node : T.cn_adj[T.cn_start[(uint64_t)node * T.L + uploc] + t];
        double f = T.freq[(uint64_t)fn * P + j] * scale;
        if (f > 0)
          sel_push(f * M[j * P + j], ((base + (uint64_t)lo) << 32) | ((uint64_t)t * P + j), f, T.node_key[fn] | extra[lo], j);
      }
    }
    g.sync();
  }

  // find_option_freq_missing_data (impute.py:1142-1172): alleles of the loci in `nid` are not in
  // the table; look the rest up, re-insert them, scale by factor_missing_data^|nid|.
  GDN void slot_missing_data(int slot, uint32_t nid) {
    const SlotDesc sd = slots[slot];
    sel_begin();
    const uint32_t have = typed & ~nid, want = full & ~nid, up = want & ~have;
    const int nup = popc16(up);
    int uploc = 0;
    for (int l = 0; l < T.L; ++l)
      if (up >> l & 1u) uploc = l;
    uint32_t keepmask = 0;
    for (int t = 0; t < n; ++t)
      if (have >> loc[t] & 1u) keepmask |= 1u << t;
    const double scale = cfg->factor_missing_pow[popc16(nid)];
    if (have != 0) {
      for (uint64_t base = 0; base < sd.ncand; base += g.n) {
        uint64_t c = base + g.tid;
        uint32_t node = GRIMB_NONE, deg = 0;
        hkey ex = 0;
        if (c < sd.ncand) {
          uint16_t ids[MAXL];
          decode(sd, slot, c, ids);
          hkey key;
          uint32_t label;
          if (pack(ids, keepmask, key, label)) {
            node = ht_lookup(T, label, key);
            ++c_probes;
            if (node != GRIMB_NONE) ++c_hits;
          }
          for (int t = 0; t < n; ++t)
            if (!(keepmask >> t & 1u)) ex |= (hkey)ids[t] << T.shift[loc[t]];
          if (node != GRIMB_NONE) {
            if (nup == 0) deg = 1;
            else if (nup == 1) {
              deg = T.cn_cnt[(uint64_t)node * T.L + uploc];
              if (deg == GRIMB_ADJ_FAULT) {
                sh->fault = 1;
                deg = 0;
              }
            }
          }
        }
        uint32_t total;
        uint32_t off = g.scan_excl(deg, total);
        sh->chunk_node[g.tid] = node;
        sh->chunk_off[g.tid] = off;
        chunk_extra[g.tid] = ex;
        if (total && g.tid == 0) {
          sh->nonempty = 1;
          if (nid == 0) sh->fault = 1;  // `not_in_data[0]` on an empty list (impute.py:1163)
        }
        g.sync();
        expand_chunk_cn(total, base, nup == 0, uploc, chunk_extra, scale);
      }
    }
    sel_finish(slot);
  }

  GD void slot_plan_b(int slot, int row) {
    if (cfg->row_is_plan_a[row]) slot_plan_a(slot);
    else slot_plan_b_row(slot, row);
  }

  // loci (bitmask) none of whose alleles on this side exist in the table (impute.py:1224-1258)
  GD uint32_t not_in_data(int side, int only_phase) const {
    uint32_t nid = 0;
    for (int t = 0; t < n; ++t) {
      bool any = false;
      for (int p = 0; p < nph && !any; ++p) {
        if (!slots[2 * p].valid) continue;
        if (only_phase >= 0 && p != only_phase) continue;
        const SlotDesc& sd = slots[2 * p + side];
        any = sd.mode ? true : sh->lexist[sd.var][t][side_of(2 * p + side, t)] != 0;
      }
      if (!any) nid |= 1u << loc[t];
    }
    return nid;
  }

  // comp_phase_prob_plan_b (impute.py:1392-1570), evaluated at epsilon = 0
  GDN void plan_b() {
    g.sync();
    const uint32_t nid1 = not_in_data(0, -1), nid2 = not_in_data(1, -1);
    g.sync();
    if (g.tid == 0)
      for (int s = 0; s < 2 * nph; ++s) {
        slots[s].first_row = NEVER_ROW;
        slots[s].cached_row = 0xFFFFFFFFu;
        top_n[s] = 0;
        slot_ne[s] = 0;
      }
    g.sync();
    const uint32_t MISSROW = 0xFFFFFFF0u;
    ent_n = 0;
    for (int row = 0; row < cfg->n_rows; ++row) {
      if (cfg->row_blocks[row] == 0) break;
      for (int p = 0; p < nph; ++p) {
        if (!slots[2 * p].valid) continue;
        for (int s = 0; s < 2; ++s) {
          const int slot = 2 * p + s;
          const uint32_t nid = s ? nid2 : nid1;
          g.sync();
          if (nid == 0) {
            const uint32_t fr = slots[slot].first_row;
            const uint32_t idx = (uint32_t)row < fr ? (uint32_t)row : fr;
            if (slots[slot].cached_row != idx) {
              slot_plan_b(slot, (int)idx);
              if (ws_fail) return;
              g.sync();
              if (g.tid == 0) {
                slots[slot].cached_row = idx;
                if (slot_ne[slot]) slots[slot].first_row = idx;
              }
              g.sync();
            }
          } else {
            // side 2 is only evaluated when side 1 returned something (impute.py:1447)
            const bool need = s == 0 || slot_ne[2 * p] != 0;
            if (need && slots[slot].cached_row != MISSROW) {
              slot_missing_data(slot, nid);
              if (ws_fail) return;
              g.sync();
              if (g.tid == 0) slots[slot].cached_row = MISSROW;
              g.sync();
            }
          }
        }
      }
      gen_entries(0.0);
      if (ws_fail) return;
      if (ent_n > 0) {
        dedup_entries();
        return;
      }
    }
    // second stage (impute.py:1490-1558); its six passes are identical, so it runs once
    bool any = false;
    for (int p = 0; p < nph; ++p) {
      if (!slots[2 * p].valid) continue;
      g.sync();
      const uint32_t i1 = slots[2 * p].first_row, i2 = slots[2 * p + 1].first_row;
      bool run = false;
      if (i1 == NEVER_ROW && i2 != NEVER_ROW) {
        slot_missing_data(2 * p, not_in_data(0, p));
        if (ws_fail) return;
        if (slots[2 * p + 1].cached_row != i2) slot_plan_b(2 * p + 1, (int)i2);
        run = true;
      } else if (i2 == NEVER_ROW && i1 != NEVER_ROW) {
        if (slots[2 * p].cached_row != i1) slot_plan_b(2 * p, (int)i1);
        if (ws_fail) return;
        slot_missing_data(2 * p + 1, not_in_data(1, p));
        run = true;
      }
      if (ws_fail) return;
      g.sync();
      if (!run && g.tid == 0) {
        // both sides found (the reference re-evaluates stale lists: no new pairs) or neither
        top_n[2 * p] = 0;
        top_n[2 * p + 1] = 0;
      }
      g.sync();
      any = any || run;
    }
    if (any) {
      gen_entries(0.0);
      if (ws_fail) return;
      if (ent_n > 0) dedup_entries();
    }
  }

  // ------------------------------------------------------------------ Plan C
  // comp_hap_prob_plan_c (impute.py:1264-1311): per-locus frequencies summed over populations,
  // multiplied across loci, unknown alleles re-inserted with factor_missing_data^k, untyped
  // loci filled with every node of the untyped label.
  GD double sr_sum(uint32_t node) const { return py_sum(T.freq + (uint64_t)node * T.P, T.P); }

  GDN void slot_plan_c(int slot) {
    const SlotDesc sd = slots[slot];
    const uint32_t ul = full & ~typed;
    const uint32_t ufirst = ul ? T.label_first[ul] : 0, ucnt = ul ? T.label_count[ul] : 0;
    const uint64_t mark = ar_used;
    // pass 1: value + key per candidate
    double* cval = alloc<double>(sd.ncand);
    hkey* ckey = alloc<hkey>(sd.ncand);
    if (ws_fail) return;
    for (uint64_t c = g.tid; c < sd.ncand; c += g.n) {
      uint16_t ids[MAXL];
      decode(sd, slot, c, ids);
      double v = 0;
      bool have = false, dead = false;
      int miss = 0;
      hkey key = 0;
      for (int t = 0; t < n && !dead; ++t) {
        key |= (hkey)ids[t] << T.shift[loc[t]];
        uint32_t node;
        if (!allele_node(loc[t], ids[t], node)) {
          ++miss;
        } else {
          double s = sr_sum(node);
          if (!have) {
            v = s;
            have = true;
          } else {
            v = v * s * 0.0001;
            if (!(v > 0)) dead = true;  // open_option_ drops it; the candidate yields nothing
          }
        }
      }
      if (!have || dead) v = -1.0;
      else if (miss > 0) v = v * cfg->factor_missing_pow[miss];
      cval[c] = v;
      ckey[c] = key;
    }
    sel_begin();
    if (cfg->save_space) {
      // open_option_ prunes both operands to 10 entries (impute.py:1048-1059); the candidate
      // dict and the untyped-label dict are materialised (vectors of length 1)
      uint32_t nc = 0;
      for (uint64_t b = 0; b < sd.ncand; b += g.n) {
        uint64_t c = b + g.tid;
        bool in = c < sd.ncand && cval[c] >= 0;
        double v = in ? cval[c] : 0.0;
        hkey k = in ? ckey[c] : (hkey)0;
        uint32_t total;
        uint32_t pos = g.scan_excl(in ? 1u : 0u, total);  // barriers inside: all reads are done
        if (in) {
          cval[nc + pos] = v;
          ckey[nc + pos] = k;
        }
        nc += total;
        g.sync();
      }
      g.sync();
      if (ul && ucnt && nc) {
        hkey* uk = alloc<hkey>(ucnt);
        double* uv = alloc<double>(ucnt);
        if (ws_fail) return;
        for (uint32_t u = g.tid; u < ucnt; u += g.n) {
          uk[u] = T.node_key[ufirst + u];
          uv[u] = sr_sum(ufirst + u);
        }
        g.sync();
        uint32_t a = prune10(ckey, cval, nc, 1);
        uint32_t b2 = prune10(uk, uv, ucnt, 1);
        const uint32_t items = a * b2;
        for (uint32_t q0 = 0; q0 < items; q0 += g.n) {
          sel_reserve();
          uint32_t q = q0 + g.tid;
          if (q < items) {
            double f = cval[q / b2] * uv[q % b2] * 0.0001;
            if (f > 0) {
              sh->nonempty = 1;
              sel_push(f * M[0], q, f, ckey[q / b2] | uk[q % b2], 0);
            }
          }
        }
      } else if (!ul) {
        for (uint32_t q0 = 0; q0 < nc; q0 += g.n) {
          sel_reserve();
          uint32_t q = q0 + g.tid;
          if (q < nc && cval[q] > 0) {
            sh->nonempty = 1;
            sel_push(cval[q] * M[0], q, cval[q], ckey[q], 0);
          }
        }
      }
    } else if (ul && ucnt) {
      const uint64_t items = sd.ncand * (uint64_t)ucnt;
      for (uint64_t q0 = 0; q0 < items; q0 += (uint64_t)g.n * SEL_IPT) {
        sel_reserve(SEL_IPT);
        for (uint32_t u2 = 0; u2 < SEL_IPT; ++u2) {
        uint64_t q = q0 + (uint64_t)u2 * g.n + g.tid;
        if (q < items) {
          uint64_t c = q / ucnt;
          uint32_t u = (uint32_t)(q % ucnt);
          double v = cval[c];
          if (v >= 0) {
            double f = v * sr_sum(ufirst + u) * 0.0001;
            if (f > 0) {
              sh->nonempty = 1;
              sel_push(f * M[0], q, f, ckey[c] | T.node_key[ufirst + u], 0);
            }
          }
        }
        }
      }
    } else if (!ul) {
      for (uint64_t q0 = 0; q0 < sd.ncand; q0 += g.n) {
        sel_reserve();
        uint64_t q = q0 + g.tid;
        if (q < sd.ncand) {
          double v = cval[q];
          if (v > 0) {
            sh->nonempty = 1;
            sel_push(v * M[0], q, v, ckey[q], 0);
          }
        }
      }
    }
    sel_finish(slot);
    ar_used = mark;
  }

  // comp_phase_prob_plan_c (impute.py:1313-1389): epsilon 0, prior all ones, one pseudo-population
  GDN void plan_c() {
    g.sync();
    if (g.tid == 0)
      for (int s = 0; s < 2 * nph; ++s) {
        top_n[s] = 0;
        slot_ne[s] = 0;
      }
    g.sync();
    for (int p = 0; p < nph; ++p) {
      if (!slots[2 * p].valid) continue;
      slot_plan_c(2 * p);
      if (ws_fail) return;
      g.sync();
      if (top_n[2 * p]) slot_plan_c(2 * p + 1);
      if (ws_fail) return;
    }
    plan_c_single = true;
    gen_entries(0.0);
    if (!ws_fail && ent_n > 0) dedup_entries();
    plan_c_single = false;
  }

  // ------------------------------------------------------------------ per-subject set-up
  // lists in GL-string order, phases (gen_phases, impute.py:274-303) and the arena allocations every
  // later step relies on; `typed` must be set.  Uniform over the group.
  GDN void setup(const GrimbBatch& B, uint64_t s) {
    const int L = T.L, P = T.P;
    // lists (original GL string order); a packed batch carries the two alleles of every locus inside the
    // subject's two keys: they are unpacked into the arena
    uint16_t* unpacked = nullptr;
#if GRIMB_KW == 1
    if (B.packed_keys) {
      unpacked = alloc<uint16_t>(2 * GRIMB_MAX_LOCI);
      if (g.tid == 0 && !ws_fail) {
        const uint64_t k0 = B.packed_keys[2 * s], k1 = B.packed_keys[2 * s + 1];
        for (int l = 0; l < L; ++l) {
          unpacked[2 * l] = (uint16_t)key_field(T, k0, l);
          unpacked[2 * l + 1] = (uint16_t)key_field(T, k1, l);
        }
      }
    }
#endif
    if (g.tid == 0) {
      sh->fault = 0;
      const uint16_t* cur = unpacked ? unpacked : B.alleles + B.allele_off[s];
      int t = 0;
      for (int l = 0; l < L; ++l)
        if (typed >> l & 1u) {
          for (int x = 0; x < 2; ++x) {
            const uint32_t c = batch_count(B, s, L, l, x);
            sh->lptr[VAR_ORIG][t][x] = cur;
            sh->lcnt[VAR_ORIG][t][x] = (uint16_t)c;
            cur += c;
          }
          ++t;
        }
      for (int v = 0; v < NVAR; ++v) sh->lhave[v] = v == VAR_ORIG;
    }
    for (int v = 0; v < NVAR; ++v) lmask[v] = nullptr;
    n = 0;
    for (int l = 0; l < L; ++l)
      if (typed >> l & 1u) loc[n++] = l;
    g.sync();
    // phases (gen_phases impute.py:274-303): keep i unless both orientations were seen
    if (g.tid == 0) {
      uint32_t het = 0;
      for (int t = 0; t < n; ++t) {
        bool same = sh->lcnt[0][t][0] == sh->lcnt[0][t][1];
        for (uint32_t i = 0; same && i < sh->lcnt[0][t][0]; ++i) same = sh->lptr[0][t][0][i] == sh->lptr[0][t][1][i];
        if (!same) het |= 1u << t;
      }
      const uint32_t low = het & ((1u << (n - 1)) - 1u);
      const bool last_het = het >> (n - 1) & 1u;
      // per-subject phase mask (impute.py:277-290): only the positions in pm may switch sides.  The
      // first index producing a flip set c is c itself; its mirror image (c ^ low, both
      // orientations equal when the last locus is homozygous) is only ever produced if low fits pm.
      const uint32_t pm = (B.phase_mask && !B.packed_keys) ? (uint32_t)B.phase_mask[s] : 0xFFFFu;
      const uint32_t eff = low & pm;
      const bool mirror = !last_het && (low & ~pm) == 0;
      int k = 0;
      for (uint32_t i = 0; i < (1u << (n - 1)); ++i) {
        if (i & ~eff) continue;
        if (mirror && i > (low ^ i)) continue;
        sh->ph[k++] = (uint16_t)i;
      }
      sh->cnt[0] = k;
    }
    g.sync();
    nph = sh->cnt[0];
    g.sync();
    const int ns = 2 * nph;
    slots = alloc<SlotDesc>(ns);
    top = alloc<TopItem>((uint64_t)ns * K);
    top_n = alloc<uint32_t>(ns);
    slot_ne = alloc<uint32_t>(ns);
    capsel = (uint32_t)(2 * K + (int)SEL_IPT * g.n + 64);
    if (capsel < 2048) capsel = 2048;
    sel = alloc<SelItem>(capsel);
    sel2 = alloc<SelItem>(capsel);
    sel_idx = alloc<uint32_t>(capsel);
    chunk_extra = alloc<hkey>(g.n);
    st_hap[0] = alloc<GrimbHapRow>(cfg->n_results);
    st_hap[1] = alloc<GrimbHapRow>(cfg->n_results);
    st_pop[0] = alloc<GrimbPopRow>(cfg->n_pop_results > 0 ? cfg->n_pop_results : 1);
    st_pop[1] = alloc<GrimbPopRow>(cfg->n_pop_results > 0 ? cfg->n_pop_results : 1);
    st_pair = alloc<GrimbPopRow>(cfg->hap_pop_pair ? (cfg->n_results > 0 ? cfg->n_results : 1) : 1);
    st_cnt = alloc<uint32_t>(4);
    Msubj = B.priors + (uint64_t)batch_prior(B, s) * P * P;
    M = Msubj;
    if (!ws_fail) {
      if (g.tid == 0) {
        for (int q = 0; q < ns; ++q) {
          SlotDesc sd;
          sd.ncand = 0;
          sd.filt = nullptr;
          sd.var = VAR_ORIG;
          sd.mode = 0;
          sd.valid = 0;
          sd.pad = 0;
          sd.first_row = NEVER_ROW;
          sd.cached_row = 0xFFFFFFFFu;
          slots[q] = sd;
        }
        for (int q = 0; q < 4; ++q) st_cnt[q] = 0;
      }
    }
  }

  // one slot for the cooperative slot kernel: opens the slot and its partner (a phase only counts when both
  // of its sides open, impute.py:987-988), then the Plan A probe + top-K of the slot.  false: not usable.
  GDN bool prepass_slot(int slot) {
    g.sync();   // the slot descriptors were initialised by one thread (setup)
    // worth a CTA of its own only when the slot is a large Cartesian product: an over-threshold slot is
    // a short filtered list, a small product is done faster than it is handed over
    const uint64_t opt = slot_options(slot, VAR_ORIG);
    if (opt >= (uint64_t)cfg->options_threshold || opt < PRE_MIN_CANDIDATES) return false;
    open_slot(slot);
    if (!ws_fail) open_slot(slot ^ 1);
    g.sync();
    if (ws_fail || slots[slot].ncand == 0 || slots[slot ^ 1].ncand == 0) return false;
    if (g.tid == 0) {
      top_n[slot] = 0;
      slot_ne[slot] = 0;
    }
    slot_plan_a(slot);
    g.sync();
    return !ws_fail && sh->fault == 0;
  }

  // the slot's Plan A list as computed by the cooperative slot kernel, if there is one (uniform over the group)
  GD bool load_pre(int slot) {
    if (pre == nullptr || pre_j == 0xFFFFFFFFu || slots[slot].var != VAR_ORIG) return false;
    const uint64_t idx = (uint64_t)pre_j * pre->slots_per_subject + (uint32_t)slot;
    if ((uint32_t)slot >= pre->slots_per_subject || !pre->ready[idx]) return false;
    g.sync();
    const uint32_t cnt = pre->n[idx];
    const TopItem* src = pre->top + idx * pre->K;
    TopItem* dst = top + (uint64_t)slot * K;
    for (uint32_t i = g.tid; i < cnt; i += g.n) dst[i] = src[i];
    if (g.tid == 0) {
      top_n[slot] = cnt;
      slot_ne[slot] = pre->ne[idx];
    }
    g.sync();
    return true;
  }

  // ------------------------------------------------------------------ epsilon schedule
  // call_comp_phase_prob (impute.py:1658-1724).  Leaves the final (deduplicated) accepted
  // pairs in ent[0..ent_n) and returns the plan that produced them.
  GDN int evaluate() {
    g.sync();
    if (g.tid == 0)
      for (int s = 0; s < 2 * nph; ++s) {
        top_n[s] = 0;
        slot_ne[s] = 0;
      }
    g.sync();
    for (int p = 0; p < nph; ++p) {
      if (!slots[2 * p].valid) continue;
      if (!load_pre(2 * p)) slot_plan_a(2 * p);
      if (ws_fail) return GRIMB_PLAN_A;
      g.sync();
      if (slot_ne[2 * p] && !load_pre(2 * p + 1)) slot_plan_a(2 * p + 1);  // side 2 only if side 1 returned anything
      if (ws_fail) return GRIMB_PLAN_A;
    }
    double eps = cfg->epsilon;
    bool last = false;
    ent_n = 0;
    while (eps > 0) {
      eps /= 10;
      if (eps < 1.0e-9) eps = 0.0;
      gen_entries(eps);
      if (ws_fail) return GRIMB_PLAN_A;
      if (ent_n > 0) {
        double mx = dedup_entries();
        if (ws_fail) return GRIMB_PLAN_A;
        if (eps > 0) {
          eps = mx / 100000;
          last = true;
        }
        break;
      }
    }
    if (last) {
      gen_entries(eps);
      if (ws_fail) return GRIMB_PLAN_A;
      dedup_entries();
    }
    int plan = GRIMB_PLAN_A;
    for (int level = 0; level < 2; ++level) {
      if (level == 1) M = ones;
      if (cfg->planb && ent_n == 0) {
        plan = GRIMB_PLAN_B;
        plan_b();
        if (ws_fail) return plan;
      }
    }
    return plan;
  }

  // ------------------------------------------------------------------ emission
  // write_best_hap_race_pairs (impute.py:79-99): the `limit` most probable single entries, ranked by
  // (probability desc, encounter asc), nothing merged; the populations of row k go to pair_rows[k].
  GDN uint32_t top_entries(uint32_t limit, GrimbHapRow* hap_rows, GrimbPopRow* pair_rows, uint32_t* n_rows) {
    if (g.tid == 0) *n_rows = 0;
    if (ent_n == 0 || ws_fail) return 0;
    const uint64_t mark = ar_used;
    uint32_t* ord = alloc<uint32_t>(ent_n);
    if (ws_fail) return 0;
    for (uint32_t i = g.tid; i < ent_n; i += g.n) ord[i] = i;
    g.sync();
    const Entry* E = ent;
    if (!cfg->encounter_order)   // (encounter_order: the entries as the traversal met them, impute.py:1940-1983)
    group_sort(
        g, ent_n,
        [=](uint32_t a, uint32_t b) {
          const double x = E[ord[a]].prob, y = E[ord[b]].prob;
          return x > y || (x == y && ord[a] < ord[b]);
        },
        [=](uint32_t a, uint32_t b) {
          uint32_t t = ord[a];
          ord[a] = ord[b];
          ord[b] = t;
        });
    const uint32_t rows = ent_n < limit ? ent_n : limit;
    for (uint32_t r = g.tid; r < rows; r += g.n) {
      const Entry& e = E[ord[r]];
      hap_rows[r] = make_hap_row(e.h1, e.h2, e.prob);
      GrimbPopRow o;
      o.pop_a = e.p1;
      o.pop_b = e.p2;
      o.pad = 0;
      o.prob = e.prob;
      pair_rows[r] = o;
    }
    if (g.tid == 0) *n_rows = rows;
    g.sync();
    ar_used = mark;
    return ent_n;
  }

  GDN void emit(int which, bool planc, uint32_t* tot_out) {
    // which 0: UMUG (+ sorted pops), 1: PMUG (+ first-seen pops)
    const bool pairs = which == 1 && cfg->hap_pop_pair;
    uint32_t ng;
    if (pairs) {
      ng = top_entries((uint32_t)cfg->n_results, st_hap[1], st_pair, &st_cnt[1]);
      if (planc) {   // Plan C reports populations as "all_pops" (impute.py:1375-1382)
        for (uint32_t r = g.tid; r < st_cnt[1]; r += g.n) {
          st_pair[r].pop_a = 0xFFFF;
          st_pair[r].pop_b = 0xFFFF;
        }
        g.sync();
      }
    }
    else ng = aggregate(which == 0 ? 0 : 1, (uint32_t)cfg->n_results, st_hap[which], nullptr, &st_cnt[which]);
    if (g.tid == 0) *tot_out = which == 0 ? ng : ent_n;
    // write_best_prob(subject_id, pops, probs, 1, ...) in the hap_pop_pair mode (impute.py:2088)
    aggregate(2, pairs ? 1u : (uint32_t)cfg->n_pop_results, nullptr, st_pop[which], &st_cnt[2 + which]);
    g.sync();
    if (planc && g.tid == 0) {
      // populations are reported as "all_pops" (impute.py:1375-1382); UMUG keeps the row even
      // when nothing was found (sum of an empty dict)
      uint32_t c = st_cnt[2 + which];
      if (c) {
        st_pop[which][0].pop_a = 0xFFFF;
        st_pop[which][0].pop_b = 0xFFFF;
      } else if (which == 0) {
        GrimbPopRow o;
        o.pop_a = 0xFFFF;
        o.pop_b = 0xFFFF;
        o.pad = 0;
        o.prob = 0.0;
        st_pop[0][0] = o;
        st_cnt[2] = 1;
      }
    }
    g.sync();
  }
};

// Runs one subject.  All threads of the group call it with identical arguments.
GD void run_subject(Subject& S, const GrimbBatch& B, const OutArrays& O, uint64_t s) {
  const Grp& g = S.g;
  Shared* sh = S.sh;
  const GrimbConfig* cfg = S.cfg;
  const int L = S.T.L, P = S.T.P;
  GrimbResults& R = const_cast<GrimbResults&>(O.r);
  g.sync();
  S.ar_used = 0;
  S.ws_fail = false;
  S.pair_evals = 0;
  S.c_probes = S.c_hits = 0;
  S.c_vecs = 0;
  S.plan_c_single = false;
  S.ent_n = 0;
  S.full = (1u << L) - 1u;
  S.typed = batch_typed(B, s, S.full);
  S.K = cfg->max_haps_in_phase;
  uint8_t status = GRIMB_ST_OK, plan_u = GRIMB_PLAN_NONE, plan_p = GRIMB_PLAN_NONE;
  uint32_t tot_u = 0, tot_p = 0;
  bool have_rows = false;

  if (S.typed == 0) {
    status = GRIMB_ST_SKIPPED;
  } else {
    S.setup(B, s);
    if (!S.ws_fail) {
      g.sync();
      S.compute_exist(VAR_ORIG);
      // the rest of the arena holds the accepted pairs of the evaluation in progress
      int nvalid = S.open_all();
      if (!S.ws_fail && nvalid == 0) {
        S.make_variant(VAR_VALID);
        if (!S.ws_fail) {
          S.set_variant(VAR_VALID, false);
          nvalid = S.open_all();
        }
      }
      if (!S.ws_fail && nvalid == 0) {
        S.make_variant(VAR_C10);
        if (!S.ws_fail) {
          S.set_variant(VAR_C10, false);
          nvalid = S.open_all();
        }
      }
      if (!S.ws_fail && nvalid == 0) status = GRIMB_ST_NO_PHASES;
      if (!S.ws_fail && nvalid > 0) {
        // entries region: what is left, minus room for dedup/aggregation scratch (~3x)
        uint64_t left = S.ar_cap - S.ar_used;
        uint64_t cap = left / (sizeof(Entry) * 2 + 64);
        if (cap > 0x7FFFFFFFull) cap = 0x7FFFFFFFull;
        S.ent_cap = (uint32_t)cap;
        S.ent = S.alloc<Entry>(S.ent_cap);
        bool reduced = false, faulted = false;
        uint64_t shared_evals = 0;
        if (cfg->output_umug) {
          const uint64_t e0 = S.pair_evals;
          plan_u = (uint8_t)S.evaluate();
          shared_evals = S.pair_evals - e0;
          if (!S.ws_fail && cfg->planb && S.ent_n == 0) {
            plan_u = GRIMB_PLAN_C;
            S.make_variant(VAR_C1);  // prior is all ones here
            if (!S.ws_fail) {
              S.set_variant(VAR_C1, true);
              reduced = true;
              if (S.open_all() == 0) faulted = true;  // check_full_haplo on [] raises
              else S.plan_c();
            }
          }
          if (!S.ws_fail && !faulted) S.emit(0, plan_u == GRIMB_PLAN_C, &tot_u);
          S.M = S.Msubj;
        }
        if (cfg->output_pmug && !S.ws_fail && !faulted) {
          if (!cfg->output_umug || reduced) plan_p = (uint8_t)S.evaluate();
          else {
            plan_p = plan_u;
            S.pair_evals += shared_evals;  // the reference evaluates again for the haplotype output
          }
          if (!S.ws_fail && cfg->planb && S.ent_n == 0 && !cfg->em) {
            plan_p = GRIMB_PLAN_C;
            if (!reduced) {
              S.make_variant(VAR_C1);
              if (!S.ws_fail) S.set_variant(VAR_C1, true);
            }
            if (!S.ws_fail) {
              if (S.open_all() == 0) faulted = true;
              else S.plan_c();
            }
          }
          if (!S.ws_fail && !faulted) S.emit(1, plan_p == GRIMB_PLAN_C, &tot_p);
        }
        g.sync();
        if (faulted || sh->fault) status = GRIMB_ST_FAULT;
        else have_rows = true;
      }
    }
    if (S.ws_fail) status = GRIMB_ST_WORKSPACE;
  }
  // ---- publish
  g.sync();
  uint64_t evals = g.sum64(S.pair_evals);
  const uint64_t n_probes = g.sum64((uint64_t)S.c_probes), n_hits = g.sum64((uint64_t)S.c_hits);
  if (g.tid == 0) {
    if (n_probes) atom_add64(O.probe_counters + 0, (unsigned long long)n_probes);
    if (n_hits) atom_add64(O.probe_counters + 1, (unsigned long long)n_hits);
    if (S.c_vecs) atom_add64(O.probe_counters + 2, (unsigned long long)S.c_vecs);
    sh->cnt[4] = tot_u;
    sh->cnt[5] = tot_p;
  }
  g.sync();
  uint32_t nu = 0, np = 0, nup = 0, npp = 0;
  if (have_rows && status == GRIMB_ST_OK) {
    nu = S.st_cnt[0];
    np = S.st_cnt[1];
    nup = S.st_cnt[2];
    npp = S.st_cnt[3];
  }
  const uint32_t npair = cfg->hap_pop_pair ? np : 0u;   // companion population rows of the PMUG rows
  if (g.tid == 0) {
    unsigned long long hb = 0, pb = 0;
    if (nu + np) hb = atom_add64(O.hap_counter, (unsigned long long)(nu + np));
    if (nup + npp + npair) pb = atom_add64(O.pop_counter, (unsigned long long)(nup + npp + npair));
    GrimbSubjectResult o;
    o.status = status;
    o.plan_umug = plan_u;
    o.plan_pmug = plan_p;
    o.reserved = 0;
    o.n_umug = nu;
    o.n_pmug = np;
    o.n_umug_pops = nup;
    o.n_pmug_pops = npp;
    o.tot_umug = sh->cnt[4];
    o.tot_pmug = sh->cnt[5];
    o.pair_evals = evals > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)evals;
    o.hap_off = hb;
    o.pop_off = pb;
    publish_general(O, s, o);
    sh->cnt[6] = (uint32_t)(hb & 0xFFFFFFFFu);
    sh->cnt[7] = (uint32_t)(hb >> 32);
    sh->cnt[0] = (uint32_t)(pb & 0xFFFFFFFFu);
    sh->cnt[1] = (uint32_t)(pb >> 32);
  }
  g.sync();
  const uint64_t hb = (uint64_t)sh->cnt[6] | ((uint64_t)sh->cnt[7] << 32);
  const uint64_t pb = (uint64_t)sh->cnt[0] | ((uint64_t)sh->cnt[1] << 32);
  if ((int64_t)(hb + nu + np) <= R.hap_capacity) {
    for (uint32_t i = g.tid; i < nu; i += g.n) R.hap_rows[hb + i] = S.st_hap[0][i];
    for (uint32_t i = g.tid; i < np; i += g.n) R.hap_rows[hb + nu + i] = S.st_hap[1][i];
  }
  if ((int64_t)(pb + nup + npp + npair) <= R.pop_capacity) {
    for (uint32_t i = g.tid; i < nup; i += g.n) R.pop_rows[pb + i] = S.st_pop[0][i];
    for (uint32_t i = g.tid; i < npp; i += g.n) R.pop_rows[pb + nup + i] = S.st_pop[1][i];
    for (uint32_t i = g.tid; i < npair; i += g.n) R.pop_rows[pb + nup + npp + i] = S.st_pair[i];
  }
  g.sync();
}

// One work item of the cooperative slot pass: slot `slot` of the subject at position j of the heavy list.
// All threads of the group call it with identical arguments.
GD void run_slot_item(Subject& S, const GrimbBatch& B, const OutArrays& O, uint64_t s, uint32_t j, int slot, const PreView& pv) {
  const Grp& g = S.g;
  const GrimbConfig* cfg = S.cfg;
  const uint64_t idx = (uint64_t)j * pv.slots_per_subject + (uint32_t)slot;
  g.sync();
  S.ar_used = 0;
  S.ws_fail = false;
  S.pair_evals = 0;
  S.c_probes = S.c_hits = 0;
  S.c_vecs = 0;
  S.plan_c_single = false;
  S.ent_n = 0;
  S.full = (1u << S.T.L) - 1u;
  S.typed = batch_typed(B, s, S.full);
  S.K = cfg->max_haps_in_phase;
  bool ok = S.typed != 0;
  if (ok) {
    S.setup(B, s);
    ok = !S.ws_fail && slot < 2 * S.nph;
  }
  if (ok) ok = S.prepass_slot(slot);
  if (ok) {
    const uint32_t cnt = S.top_n[slot];
    const TopItem* src = S.top + (uint64_t)slot * S.K;
    TopItem* dst = pv.top + idx * pv.K;
    for (uint32_t i = g.tid; i < cnt; i += g.n) dst[i] = src[i];
    if (g.tid == 0) {
      pv.n[idx] = cnt;
      pv.ne[idx] = S.slot_ne[slot];
    }
    // the probes of this slot are counted here; the subject's own CTA loads the list instead of probing
    const uint64_t np = g.sum64((uint64_t)S.c_probes), nh = g.sum64((uint64_t)S.c_hits);
    if (g.tid == 0) {
      if (np) atom_add64(O.probe_counters + 0, (unsigned long long)np);
      if (nh) atom_add64(O.probe_counters + 1, (unsigned long long)nh);
      if (S.c_vecs) atom_add64(O.probe_counters + 2, (unsigned long long)S.c_vecs);
    }
  }
  g.sync();
  if (g.tid == 0) {
#if GRIMB_DEVICE
    __threadfence();
#endif
    pv.ready[idx] = ok ? 1u : 0u;
  }
}

}  // namespace grimb
