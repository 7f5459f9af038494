// grimb_emu.cpp -- TEST-ONLY single-thread emulation build of the per-subject algorithm.
//
// Compiles py-graph-imputation_b200/csrc/grimb_plan.h with -DGRIMB_EMU (group size 1, plain
// memory operations instead of atomics) so the integer/FP64 logic of the CUDA kernels can be
// checked against the oracle on a machine without a GPU.  It is built by tests/emu/build.sh
// into tests/emu/libgrimb_emu.so, loaded only by tests/, and is not part of libgrimb200.so:
// the product has no CPU path.
#include <stdlib.h>
#include <string.h>

#include "../../py-graph-imputation_b200/csrc/grimb_plan.h"

using namespace grimb;

extern "C" {

typedef struct {
  int32_t L, P;
  uint32_t n_nodes, n_full;
  uint8_t shift[9];
  uint8_t width[9];
  uint32_t n_alleles[9];
  const uint32_t* label_first;
  const uint32_t* label_count;
  const uint64_t* ht_off;
  const uint32_t* ht_mask;
  const void* slots;
  const uint64_t* node_key;
  const double* freq;
  const uint32_t* tl_start;
  const uint32_t* tl_cnt;
  const uint32_t* tl_adj;
  const uint32_t* cn_start;
  const uint32_t* cn_cnt;
  const uint32_t* cn_adj;
} GrimbEmuTables;

int grimb_emu_impute(const GrimbEmuTables* t, const GrimbConfig* cfg_in, const GrimbBatch* batch,
                     GrimbResults* res, uint64_t arena_bytes) {
  GrimbConfig cfg_dev = *cfg_in;                // as upload_cfg (grimb200.cu) hands it to the kernels
  if (cfg_dev.plan_a_only) cfg_dev.planb = 0;
  const GrimbConfig* cfg = &cfg_dev;
  static Shared sh;
  Subject S;
  memset(&S, 0, sizeof(S));
  S.g.tid = 0;
  S.g.n = 1;
  S.g.scratch = sh.scratch;
  S.sh = &sh;
  S.T.L = t->L;
  S.T.P = t->P;
  S.T.n_nodes = t->n_nodes;
  S.T.n_full = t->n_full;
  memcpy(S.T.shift, t->shift, 9);
  memcpy(S.T.width, t->width, 9);
  memcpy(S.T.n_alleles, t->n_alleles, sizeof(t->n_alleles));
  S.T.label_first = t->label_first;
  S.T.label_count = t->label_count;
  S.T.ht_off = t->ht_off;
  S.T.ht_mask = t->ht_mask;
  S.T.slots = (const HSlot*)t->slots;
  S.T.node_key = t->node_key;
  S.T.freq = t->freq;
  S.T.tl_start = t->tl_start;
  S.T.tl_cnt = t->tl_cnt;
  S.T.tl_adj = t->tl_adj;
  S.T.cn_start = t->cn_start;
  S.T.cn_cnt = t->cn_cnt;
  S.T.cn_adj = t->cn_adj;
  S.cfg = cfg;
  const int P = t->P;
  double* ones = (double*)malloc(sizeof(double) * P * P);
  for (int i = 0; i < P * P; ++i) ones[i] = 1.0;
  S.ones = ones;
  S.ar_base = (char*)malloc(arena_bytes);
  S.ar_cap = arena_bytes;
  OutArrays O;
  O.r = *res;
  unsigned long long hc = 0, pc = 0, wc = 0, gc = 0, ec = 0, prc[3] = {0, 0, 0};
  O.probe_counters = prc;
  O.hap_counter = &hc;
  O.pop_counter = &pc;
  O.word_counter = &wc;
  O.general_counter = &gc;
  O.evals_counter = &ec;
  // GRIMB_EMU_GROUP=1: the cooperative slot pass (run_slot_item for every slot of every subject) first, then
  // every subject starts from the pre-computed Plan A lists -- the CPU twin of k_impute's mode 1 / mode 0
  PreView pv;
  memset(&pv, 0, sizeof(pv));
  const char* grp = getenv("GRIMB_EMU_GROUP");
  if (grp && grp[0] == '1' && !batch->phase_mask && batch->n_subjects > 0) {
    const uint32_t spp = 1u << t->L, K = (uint32_t)cfg->max_haps_in_phase;
    const uint64_t nS = (uint64_t)batch->n_subjects;
    pv.top = (TopItem*)malloc(nS * spp * K * sizeof(TopItem));
    pv.n = (uint32_t*)calloc(nS * spp, 4);
    pv.ne = (uint32_t*)calloc(nS * spp, 4);
    pv.ready = (uint32_t*)calloc(nS * spp, 4);
    pv.max_subjects = (uint32_t)nS;
    pv.slots_per_subject = spp;
    pv.K = K;
    for (uint64_t s = 0; s < nS; ++s)
      for (uint32_t q = 0; q < spp; ++q) run_slot_item(S, *batch, O, s, (uint32_t)s, (int)q, pv);
    S.pre = &pv;
  }
  for (int64_t s = 0; s < batch->n_subjects; ++s) {
    S.pre_j = pv.top ? (uint32_t)s : 0xFFFFFFFFu;
    run_subject(S, *batch, O, (uint64_t)s);
  }
  free(pv.top);
  free(pv.n);
  free(pv.ne);
  free(pv.ready);
  res->totals[0] = (int64_t)wc;
  res->totals[1] = (int64_t)gc;
  res->totals[2] = (int64_t)hc;
  res->totals[3] = (int64_t)pc;
  res->totals[4] = (int64_t)ec;
  res->totals[5] = 0;
  res->totals[6] = (int64_t)prc[0];
  res->totals[7] = (int64_t)prc[1];
  res->totals[8] = (int64_t)prc[2];
  free(S.ar_base);
  free(ones);
  return (int64_t)gc > res->general_capacity || (int64_t)hc > res->hap_capacity || (int64_t)pc > res->pop_capacity
             ? GRIMB_E_CAPACITY : 0;
}
}
