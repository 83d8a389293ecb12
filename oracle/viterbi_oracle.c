/* viterbi_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker).
 *
 * A plain-C, single-threaded CPU restatement of the reference's Viterbi hot path
 * over the machine-state x DNA-position lattice.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (dnastore_b200/) never links, imports or falls back to it.
 *
 * Parity pin: checked by tests/test_oracle.py against (i) the reference's own 11
 * Viterbi known-answer tests (reference Makefile:147,148,154,169,170,171,177,178,
 * 184,185,186) and (ii) log-likelihoods, decoded strings, traceback paths and full
 * DP matrices produced by the UNMODIFIED reference (oracle/_ref/refdriver, built
 * by oracle/Makefile) and committed under tests/golden/.
 *
 * What each function follows (all line numbers are /root/reference/src/...):
 *   oracle_toposort  -> Machine::decoderToposort        trans.cpp:604-634
 *   oracle_fill      -> ViterbiMatrix::ViterbiMatrix    viterbi.cpp:62-176
 *   oracle_traceback -> ViterbiMatrix::traceback        viterbi.cpp:195-304
 * The matrix layout, cell(k+2)*(pos*nStates+state)+mut with mut 0=S,1=D,2+i=T(i+1),
 * is the reference's (viterbi.h:52-76).  All arithmetic is IEEE fp64 with the
 * reference's association order; compile with -ffp-contract=off, no fast-math.
 *
 * The one liberty: outgoing lists are rebuilt from the destination-indexed tables,
 * so within one source the push order of the phase-2 worklist may differ from the
 * reference's transition order.  That changes the schedule only, not the result:
 * phase 2 computes a least fixed point of a monotone system, and
 * the full-matrix golden comparison confirms it bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dnab_tables.h"

#define NEG_INF (-INFINITY)

static inline double dmax(double a, double b) { return (a < b) ? b : a; } /* std::max: keeps a on ties */

typedef struct {
  uint32_t *out_emit_off, *out_emit_dest; /* per-source outgoing emit edges */
  double* out_emit_score;
  uint32_t *out_null_off, *out_null_dest;
  double* out_null_score;
  uint32_t* order; /* toposort */
} oracle_graph;

/* Machine::decoderToposort (trans.cpp:604-634): Kahn, FIFO queue, over the kept
 * transitions with no DNA output.  Returns 0, or -1 on a cycle (domain_error). */
static int oracle_toposort(const dnab_tables* t, const oracle_graph* g, uint32_t* order) {
  const uint32_t n = t->n_states;
  int* nParents = (int*)calloc(n, sizeof(int));
  uint32_t* queue = (uint32_t*)malloc(n * sizeof(uint32_t));
  long edges = 0;
  uint32_t head = 0, tail = 0, nOut = 0;
  for (uint32_t s = 0; s < n; ++s) {
    nParents[s] = (int)(t->null_off[s + 1] - t->null_off[s]);
    edges += nParents[s];
  }
  for (uint32_t s = 0; s < n; ++s)
    if (nParents[s] == 0) queue[tail++] = s;
  while (head < tail) {
    const uint32_t v = queue[head++];
    order[nOut++] = v;
    for (uint32_t e = g->out_null_off[v]; e < g->out_null_off[v + 1]; ++e) {
      const uint32_t m = g->out_null_dest[e];
      --edges;
      if (--nParents[m] == 0) queue[tail++] = m;
    }
  }
  free(nParents);
  free(queue);
  return edges > 0 ? -1 : 0;
}

static void build_out(uint32_t n, const uint32_t* in_off, const uint32_t* in_src, const double* in_score,
                      uint32_t n_edges, uint32_t** out_off, uint32_t** out_dest, double** out_score) {
  uint32_t* off = (uint32_t*)calloc(n + 2, sizeof(uint32_t));
  uint32_t* dest = (uint32_t*)malloc((n_edges ? n_edges : 1) * sizeof(uint32_t));
  double* score = (double*)malloc((n_edges ? n_edges : 1) * sizeof(double));
  for (uint32_t e = 0; e < n_edges; ++e) off[in_src[e] + 2]++;
  for (uint32_t s = 0; s < n; ++s) off[s + 2] += off[s + 1];
  for (uint32_t d = 0; d < n; ++d)
    for (uint32_t e = in_off[d]; e < in_off[d + 1]; ++e) {
      const uint32_t p = off[in_src[e] + 1]++;
      dest[p] = d;
      score[p] = in_score[e];
    }
  *out_off = off;
  *out_dest = dest;
  *out_score = score;
}

static void graph_init(const dnab_tables* t, oracle_graph* g) {
  build_out(t->n_states, t->emit_off, t->emit_src, t->emit_score, t->n_emit, &g->out_emit_off, &g->out_emit_dest,
            &g->out_emit_score);
  build_out(t->n_states, t->null_off, t->null_src, t->null_score, t->n_null, &g->out_null_off, &g->out_null_dest,
            &g->out_null_score);
  g->order = (uint32_t*)malloc(t->n_states * sizeof(uint32_t));
}

static void graph_free(oracle_graph* g) {
  free(g->out_emit_off);
  free(g->out_emit_dest);
  free(g->out_emit_score);
  free(g->out_null_off);
  free(g->out_null_dest);
  free(g->out_null_score);
  free(g->order);
}

#define CELL(state, pos, m) cell[(size_t)(k + 2) * ((size_t)(pos) * n + (state)) + (m)]
#define SC(state, pos) CELL(state, pos, 0)
#define DC(state, pos) CELL(state, pos, 1)
#define TC(state, pos, i) CELL(state, pos, 2 + (i))
#define CTX(state, i) t->ctx[(size_t)(state) * k + (i)]

/* ViterbiMatrix constructor, viterbi.cpp:62-176.  seq[] holds tokens 0..3. */
static void oracle_fill(const dnab_tables* t, const oracle_graph* g, const uint8_t* seq, int L, double* cell) {
  const uint32_t n = t->n_states, k = t->k;
  const size_t nCells = (size_t)(k + 2) * n * ((size_t)L + 1);
  for (size_t i = 0; i < nCells; ++i) cell[i] = NEG_INF; /* :66 */

  if (t->local) { /* :75-79 */
    for (uint32_t s = 0; s < n; ++s) SC(s, 0) = 0;
  } else
    SC(0, 0) = 0;

  uint32_t* pushStates = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
  uint8_t* onStack = (uint8_t*)malloc(n);

  for (int pos = 0; pos <= L; ++pos) { /* :86 */
    /* phase 1, :88-108 */
    for (uint32_t oi = 0; oi < n; ++oi) {
      const uint32_t state = g->order[oi];
      const int mdl = t->mdl[state];
      if (pos > 0)
        for (uint32_t e = t->emit_off[state]; e < t->emit_off[state + 1]; ++e)
          SC(state, pos) = dmax(SC(state, pos), SC(t->emit_src[e], pos - 1) + t->emit_score[e] + t->noGap +
                                                    t->sub[t->emit_base[e] * 4 + seq[pos - 1]]); /* :94-95 */
      for (uint32_t e = t->null_off[state]; e < t->null_off[state + 1]; ++e)
        SC(state, pos) = dmax(SC(state, pos), SC(t->null_src[e], pos) + t->null_score[e]); /* :98-99 */
      if (mdl > 0 && pos > 0) {
        SC(state, pos) = dmax(SC(state, pos), TC(state, pos - 1, 0) + t->sub[CTX(state, 0) * 4 + seq[pos - 1]]); /* :102-103 */
        for (int dupIdx = 0; dupIdx < mdl - 1; ++dupIdx)
          TC(state, pos, dupIdx) =
              TC(state, pos - 1, dupIdx + 1) + t->sub[CTX(state, dupIdx + 1) * 4 + seq[pos - 1]]; /* :105-106 */
      }
    }

    /* phase 2, :110-159: LIFO worklist seeded with every state in toposort order */
    uint32_t top = n;
    memcpy(pushStates, g->order, n * sizeof(uint32_t));
    memset(onStack, 1, n);
    while (top > 0) {
      const uint32_t state = pushStates[--top];
      onStack[state] = 0;
      const double dsrc = DC(state, pos);
      const double ssrc = dmax(SC(state, pos), dsrc + t->delEnd); /* :119-120 */
      SC(state, pos) = ssrc;
      for (uint32_t e = g->out_emit_off[state]; e < g->out_emit_off[state + 1]; ++e) {
        const double dsc = dmax(dsrc + t->delExtend, ssrc + t->delOpen) + g->out_emit_score[e]; /* :124-125 */
        const uint32_t dest = g->out_emit_dest[e];
        if (dsc > DC(dest, pos)) {
          DC(dest, pos) = dsc;
          if (!onStack[dest]) {
            pushStates[top++] = dest;
            onStack[dest] = 1;
          }
        }
      }
      for (uint32_t e = g->out_null_off[state]; e < g->out_null_off[state + 1]; ++e) {
        int push = 0;
        const uint32_t dest = g->out_null_dest[e];
        const double dsc = dsrc + g->out_null_score[e]; /* :140 */
        if (dsc > DC(dest, pos)) {
          DC(dest, pos) = dsc;
          push = 1;
        }
        const double ssc = ssrc + g->out_null_score[e]; /* :147 */
        if (ssc > SC(dest, pos)) {
          SC(dest, pos) = ssc;
          push = 1;
        }
        if (push && !onStack[dest]) {
          pushStates[top++] = dest;
          onStack[dest] = 1;
        }
      }
    }

    /* phase 3, :161-168 */
    if (pos > 0)
      for (uint32_t state = 0; state < n; ++state) {
        const int mdl = t->mdl[state];
        for (int dupIdx = 0; dupIdx < mdl; ++dupIdx)
          TC(state, pos, dupIdx) = dmax(TC(state, pos, dupIdx), SC(state, pos) + t->tanDup + t->len[dupIdx]); /* :166-167 */
      }
  }

  if (t->local) /* :171-173 */
    for (uint32_t state = 0; state < n; ++state) SC(n - 1, L) = dmax(SC(n - 1, L), SC(state, L));

  free(pushStates);
  free(onStack);
}

typedef struct {
  double best;
  uint32_t bestState;
  int bestPos, bestMut;
  uint8_t bestIn;
  int found;
} tb_best;

static inline void update_best(tb_best* b, const double* cell, uint32_t n, uint32_t k, uint32_t srcState, int srcPos,
                               int srcMut, double transScore, uint8_t in) {
  const double score = CELL(srcState, srcPos, srcMut) + transScore; /* :218 */
  if (score > b->best) {                                            /* :219 */
    b->best = score;
    b->bestState = srcState;
    b->bestPos = srcPos;
    b->bestMut = srcMut;
    b->bestIn = in;
    b->found = 1;
  }
}

/* ViterbiMatrix::traceback, viterbi.cpp:195-304.
 * Returns 0 ok, 1 no valid decoding (loglike == -inf, :198-201), -2 output buffer too
 * small, -3 traceback failure (the reference's Assert at :232-233). */
static int oracle_traceback(const dnab_tables* t, const uint8_t* seq, int L, const double* cell, char* decoded,
                            int decoded_cap, int* decoded_len, int32_t* path, int path_cap, int* path_len) {
  const uint32_t n = t->n_states, k = t->k;
  int nDec = 0, nPath = 0;
  *decoded_len = 0;
  if (path_len) *path_len = 0;
  if (!(SC(n - 1, L) > NEG_INF)) return 1;

  uint32_t state = n - 1;
  int pos = L, mut = 0;
  tb_best b;
  /* :239-245 */
  b.best = NEG_INF;
  b.found = 0;
  b.bestIn = 0;
  if (t->local)
    for (uint32_t s = 0; s < n; ++s) update_best(&b, cell, n, k, s, L, 0, 0, 0);
  else
    update_best(&b, cell, n, k, n - 1, L, 0, 0, 0);
  if (!b.found) return -3;
  state = b.bestState;
  pos = b.bestPos;
  mut = b.bestMut;

  char* rev = (char*)malloc((size_t)decoded_cap + 1);
  while (pos >= 0 && state > 0) { /* :247 */
    const int mdl = t->mdl[state];
    if (path && nPath < path_cap) {
      path[3 * nPath] = (int32_t)state;
      path[3 * nPath + 1] = pos;
      path[3 * nPath + 2] = mut;
    }
    ++nPath;
    b.best = NEG_INF;
    b.found = 0;
    b.bestIn = 0;
    if (mut == 0) { /* :251-264 */
      if (pos > 0)
        for (uint32_t e = t->emit_off[state]; e < t->emit_off[state + 1]; ++e)
          update_best(&b, cell, n, k, t->emit_src[e], pos - 1, 0,
                      t->emit_score[e] + t->noGap + t->sub[t->emit_base[e] * 4 + seq[pos - 1]], t->emit_in[e]);
      for (uint32_t e = t->null_off[state]; e < t->null_off[state + 1]; ++e)
        update_best(&b, cell, n, k, t->null_src[e], pos, 0, t->null_score[e], t->null_in[e]);
      update_best(&b, cell, n, k, state, pos, 1, t->delEnd, 0);
      if (mdl > 0 && pos > 0) update_best(&b, cell, n, k, state, pos - 1, 2, t->sub[CTX(state, 0) * 4 + seq[pos - 1]], 0);
      if (pos == 0 && t->local) update_best(&b, cell, n, k, 0, 0, 0, 0, 0);
    } else if (mut == 1) { /* :269-276 */
      for (uint32_t e = t->emit_off[state]; e < t->emit_off[state + 1]; ++e) {
        update_best(&b, cell, n, k, t->emit_src[e], pos, 1, t->emit_score[e] + t->delExtend, t->emit_in[e]);
        update_best(&b, cell, n, k, t->emit_src[e], pos, 0, t->emit_score[e] + t->delOpen, t->emit_in[e]);
      }
      for (uint32_t e = t->null_off[state]; e < t->null_off[state + 1]; ++e)
        update_best(&b, cell, n, k, t->null_src[e], pos, 1, t->null_score[e], t->null_in[e]);
    } else { /* :281-286 */
      const int dupIdx = mut - 2;
      if (dupIdx < mdl - 1)
        update_best(&b, cell, n, k, state, pos - 1, 2 + dupIdx + 1, t->sub[CTX(state, dupIdx + 1) * 4 + seq[pos - 1]], 0);
      update_best(&b, cell, n, k, state, pos, 0, t->tanDup + t->len[dupIdx], 0);
    }
    /* checkBest, :230-237 */
    {
      const double expected = CELL(state, pos, mut);
      const double denom = fabs(expected) < 1e-6 ? 1 : expected;
      if (!b.found || !(fabs((b.best - expected) / denom) < 1e-6)) {
        free(rev);
        return -3;
      }
    }
    state = b.bestState;
    pos = b.bestPos;
    mut = b.bestMut;
    if (b.bestIn) { /* :299-300 */
      if (nDec >= decoded_cap) {
        free(rev);
        return -2;
      }
      rev[nDec++] = (char)b.bestIn;
    }
  }
  for (int i = 0; i < nDec; ++i) decoded[i] = rev[nDec - 1 - i];
  free(rev);
  *decoded_len = nDec;
  if (path_len) *path_len = nPath;
  return 0;
}

/* ---- exported entry points (ctypes) ------------------------------------ */

/* Decode one read.  cells may be NULL (scratch is allocated) or a caller buffer of
 * (L+1)*n_states*(k+2) doubles that receives the full matrix in the reference's
 * layout.  path (may be NULL) receives (state,pos,mut) triples in traceback order,
 * starting at the cell the traceback starts from; *path_len is the number of
 * triples the traceback visited (may exceed path_cap).
 * Returns 0 ok, 1 "No valid Viterbi decoding found" (decoded_len = 0), -1 null
 * cycle (the reference throws std::domain_error), -2/-3 see oracle_traceback. */
int dnab_oracle_viterbi(const dnab_tables* t, const uint8_t* seq, int L, double* loglike, char* decoded,
                        int decoded_cap, int* decoded_len, int32_t* path, int path_cap, int* path_len, double* cells) {
  oracle_graph g;
  graph_init(t, &g);
  if (oracle_toposort(t, &g, g.order) != 0) {
    graph_free(&g);
    return -1;
  }
  const uint32_t n = t->n_states, k = t->k;
  double* cell = cells ? cells : (double*)malloc((size_t)(k + 2) * n * ((size_t)L + 1) * sizeof(double));
  oracle_fill(t, &g, seq, L, cell);
  *loglike = SC(n - 1, L);
  const int rc = oracle_traceback(t, seq, L, cell, decoded, decoded_cap, decoded_len, path, path_cap, path_len);
  if (!cells) free(cell);
  graph_free(&g);
  return rc;
}

/* Batch form used by bench.py's CPU baseline: reads are concatenated tokens with
 * offsets[n_reads+1]; decoded strings are written at decoded + i*decoded_stride.
 * Single-threaded, like the reference's decodeFastSeqs loop (viterbi.cpp:312-318),
 * but (unlike the reference) the graph is built once per batch, not once per read. */
int dnab_oracle_viterbi_batch(const dnab_tables* t, const uint8_t* seqs, const int64_t* offsets, int n_reads,
                              double* loglike, char* decoded, int decoded_stride, int* decoded_len, int* status) {
  oracle_graph g;
  graph_init(t, &g);
  if (oracle_toposort(t, &g, g.order) != 0) {
    graph_free(&g);
    return -1;
  }
  const uint32_t n = t->n_states, k = t->k;
  for (int r = 0; r < n_reads; ++r) {
    const int L = (int)(offsets[r + 1] - offsets[r]);
    double* cell = (double*)malloc((size_t)(k + 2) * n * ((size_t)L + 1) * sizeof(double));
    oracle_fill(t, &g, seqs + offsets[r], L, cell);
    loglike[r] = SC(n - 1, L);
    status[r] = oracle_traceback(t, seqs + offsets[r], L, cell, decoded + (size_t)r * decoded_stride, decoded_stride,
                                 &decoded_len[r], NULL, 0, NULL);
    free(cell);
  }
  graph_free(&g);
  return 0;
}
