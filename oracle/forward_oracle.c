/* forward_oracle.c -- TEST INFRASTRUCTURE ONLY (the checker for SURVEY.md 8a-12).
 *
 * PARITY UNPINNED.  The reference has NO forward/backward over the machine-state x DNA-position
 * lattice (its only forward-backward is the pair-HMM of src/fwdback.cpp, covered by
 * pairhmm_oracle.c).  This file is therefore a SPECIFICATION BY ANALOGY, not a restatement: the
 * sum-product analogue of ViterbiMatrix's fill (reference src/viterbi.cpp:62-176) with every `max`
 * replaced by the reference's TABLE-BASED log_sum_exp (src/logsumexp.h:19-74, logsumexp.cpp:5-15):
 *
 *   column pos, phase 1 (viterbi.cpp:88-108)   S0(d) = lse_{emit-in}( ((F_S(src,pos-1)+score)+noGap)+sub[base][x] )
 *                                              (+) T(d,pos-1,0)+sub[ctx0][x];  T(d,pos,i) = T(d,pos-1,i+1)+sub[ctx(i+1)][x]
 *   closure (viterbi.cpp:97-99,110-159)        the least solution of
 *        D(d) = lse_{emit-in}( lse(D(s)+delExtend, S(s)+delOpen) + score ) (+) lse_{null-in}( D(s)+score )
 *        S(d) = S0(d) (+) lse_{null-in}( S(s)+score ) (+) D(d)+delEnd
 *     computed by SYNCHRONOUS (Jacobi) sweeps from S = S0, D = -inf: every sweep evaluates all states
 *     from the previous sweep's values, operands accumulated in the tables' list order, and the
 *     iteration stops after the first sweep that changes no cell (bitwise).  The table method drops
 *     terms more than 10 nats below the running sum, so the sweeps settle a few hops after the
 *     closure's depth.  The schedule is part of the specification: the GPU kernel performs the same
 *     sweeps and must agree BIT FOR BIT.
 *   phase 3 (viterbi.cpp:161-168)              T(d,pos,i) = lse( T(d,pos,i), (S(d,pos)+tanDup)+len[i] )
 *   init / result (viterbi.cpp:75-79,171-173; viterbi.h:102)  global: F_S(0,0)=0, loglike = F_S(end,L);
 *                                              local: F_S(s,0)=0 for all s, loglike = lse over s of F_S(s,L)
 *
 * What pins it instead of a reference run (tests/test_forward.py): an independent numpy
 * probability-space computation with a direct linear solve per column (agreement to the accuracy
 * of the table log_sum_exp, which ignores terms below e^-10), forward >= Viterbi for every read,
 * and the machine-independent identities of the error model.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dnab_tables.h"

#define LSE_MAX 10
#define LSE_PRECISION .0001
#define LSE_ENTRIES (((int)(LSE_MAX / LSE_PRECISION)) + 1)
#define NEG_INF (-INFINITY)

static double* lse_lookup = 0;

static void lse_table_init(void) { /* logsumexp.cpp:5-15 */
  if (lse_lookup) return;
  lse_lookup = (double*)malloc(sizeof(double) * LSE_ENTRIES);
  for (int n = 0; n < LSE_ENTRIES; ++n) {
    const double x = n * LSE_PRECISION;
    lse_lookup[n] = log(1. + exp(-x));
  }
}

static double lse_unary(double x) { /* logsumexp.h:52-74 */
  if (x >= LSE_MAX || isnan(x) || isinf(x)) return 0;
  if (x < 0) return -x;
  const int n = (int)(x / LSE_PRECISION);
  const double dx = x - (n * LSE_PRECISION);
  const double f0 = lse_lookup[n], f1 = lse_lookup[n + 1];
  const double df = f1 - f0;
  return f0 + df * (dx / LSE_PRECISION);
}

static double lse(double a, double b) { /* logsumexp.h:34-50 */
  double mx, diff;
  if (a == b) {
    mx = a;
    diff = 0;
  } else if (a < b) {
    mx = b;
    diff = b - a;
  } else {
    mx = a;
    diff = a - b;
  }
  return mx + lse_unary(diff);
}

static int same_bits(double a, double b) { return memcmp(&a, &b, sizeof(double)) == 0; }

/* Returns 0, or 1 if some column's closure did not settle within max_sweeps.
 * cells (optional): [(L+1)][n_states][k+2] in the reference's ViterbiMatrix layout (viterbi.h:52-76). */
int dnab_oracle_forward(const dnab_tables* t, const uint8_t* seq, int L, int max_sweeps, double* loglike,
                        long* total_sweeps, double* cells) {
  const uint32_t n = t->n_states, k = t->k;
  lse_table_init();
  double* Sprev = (double*)malloc(sizeof(double) * n);
  double* S0 = (double*)malloc(sizeof(double) * n);
  double* S[2] = {(double*)malloc(sizeof(double) * n), (double*)malloc(sizeof(double) * n)};
  double* D[2] = {(double*)malloc(sizeof(double) * n), (double*)malloc(sizeof(double) * n)};
  double* T[2] = {(double*)malloc(sizeof(double) * (size_t)n * (k ? k : 1)), (double*)malloc(sizeof(double) * (size_t)n * (k ? k : 1))};
  int rc = 0;
  long sweeps = 0;
  for (size_t i = 0; i < (size_t)n * k; ++i) T[0][i] = T[1][i] = NEG_INF;
  int cur = 0, tc = 0; /* S[cur], D[cur]: latest sweep; T[tc]: this column */
  for (int pos = 0; pos <= L; ++pos) {
    const int tp = tc;
    tc ^= 1;
    /* phase 1 */
    for (uint32_t d = 0; d < n; ++d) {
      double acc = NEG_INF;
      for (uint32_t i = 0; i < k; ++i) T[tc][(size_t)i * n + d] = NEG_INF;
      if (pos == 0)
        acc = (t->local || d == 0) ? 0. : NEG_INF;
      else {
        const int x = seq[pos - 1];
        for (uint32_t e = t->emit_off[d]; e < t->emit_off[d + 1]; ++e)
          acc = lse(acc, ((Sprev[t->emit_src[e]] + t->emit_score[e]) + t->noGap) + t->sub[t->emit_base[e] * 4 + x]);
        const uint32_t mdl = t->mdl[d];
        if (mdl > 0) {
          acc = lse(acc, T[tp][d] + t->sub[t->ctx[(size_t)d * k + 0] * 4 + x]);
          for (uint32_t i = 0; i + 1 < mdl; ++i)
            T[tc][(size_t)i * n + d] = T[tp][(size_t)(i + 1) * n + d] + t->sub[t->ctx[(size_t)d * k + i + 1] * 4 + x];
        }
      }
      S0[d] = acc;
    }
    /* closure: Jacobi sweeps until a sweep changes nothing */
    cur = 0;
    for (uint32_t d = 0; d < n; ++d) {
      S[0][d] = S0[d];
      D[0][d] = NEG_INF;
    }
    for (int sweep = 0;; ++sweep) {
      if (sweep >= max_sweeps) {
        rc = 1;
        break;
      }
      const double *So = S[cur], *Do = D[cur];
      double *Sn = S[cur ^ 1], *Dn = D[cur ^ 1];
      int changed = 0;
      for (uint32_t d = 0; d < n; ++d) {
        double nd = NEG_INF, ns = S0[d];
        for (uint32_t e = t->emit_off[d]; e < t->emit_off[d + 1]; ++e) {
          const uint32_t s = t->emit_src[e];
          nd = lse(nd, lse(Do[s] + t->delExtend, So[s] + t->delOpen) + t->emit_score[e]);
        }
        for (uint32_t e = t->null_off[d]; e < t->null_off[d + 1]; ++e) {
          const uint32_t s = t->null_src[e];
          nd = lse(nd, Do[s] + t->null_score[e]);
          ns = lse(ns, So[s] + t->null_score[e]);
        }
        ns = lse(ns, nd + t->delEnd);
        Dn[d] = nd;
        Sn[d] = ns;
        if (!same_bits(nd, Do[d]) || !same_bits(ns, So[d])) changed = 1;
      }
      cur ^= 1;
      ++sweeps;
      if (!changed) break;
    }
    /* phase 3 */
    if (pos > 0)
      for (uint32_t d = 0; d < n; ++d)
        for (uint32_t i = 0; i < t->mdl[d]; ++i)
          T[tc][(size_t)i * n + d] = lse(T[tc][(size_t)i * n + d], (S[cur][d] + t->tanDup) + t->len[i]);
    if (cells)
      for (uint32_t d = 0; d < n; ++d) {
        double* c = cells + ((size_t)pos * n + d) * (k + 2);
        c[0] = S[cur][d];
        c[1] = D[cur][d];
        for (uint32_t i = 0; i < k; ++i) c[2 + i] = T[tc][(size_t)i * n + d];
      }
    memcpy(Sprev, S[cur], sizeof(double) * n);
  }
  if (t->local) {
    double acc = NEG_INF;
    for (uint32_t d = 0; d < n; ++d) acc = lse(acc, Sprev[d]);
    *loglike = acc;
  } else
    *loglike = Sprev[n - 1];
  if (total_sweeps) *total_sweeps = sweeps;
  free(Sprev);
  free(S0);
  free(S[0]);
  free(S[1]);
  free(D[0]);
  free(D[1]);
  free(T[0]);
  free(T[1]);
  return rc;
}

/* Backward pass and posterior expected counts of the error-model events (the machine-lattice analogue of
 * FwdBackMatrix::counts, reference src/fwdback.cpp:154-188 / fwdback.h:92-113; PARITY UNPINNED like the rest
 * of this file).  Every forward move  from-cell --w--> to-cell  listed in the header has the backward
 * recursion  B(from) = lse over its moves of (w + B(to))  and the posterior usage exp(F(from)+w+B(to)-ll):
 *   pos-1 -> pos   S(s) --(score+noGap)+sub[base][x]--> S(d)      counts nNoGap, nSub[base][x]
 *                  T(d,0) --sub[ctx0][x]--> S(d);  T(d,i+1) --sub[ctx(i+1)][x]--> T(d,i)     nSub[ctx][x]
 *   within pos     D(s) --delExtend+score--> D(d)   nDelExtend      S(s) --delOpen+score--> D(d)   nDelOpen
 *                  D(s) --score--> D(d), S(s) --score--> S(d) (null transitions)
 *                  D(d) --delEnd--> S(d)            nDelEnd         S(d) --tanDup+len[i]--> T(d,i)  nTanDup, nLen[i]
 * The within-column system is again solved by synchronous sweeps from the base values until a sweep changes
 * no cell.  ll_back = B_S(start,0) (global) must agree with the forward value up to the table's accuracy.
 * counts: [nDelOpen, nTanDup, nNoGap, nDelExtend, nDelEnd, nLen[k], nSub[16]] (the order of pairhmm_oracle.c).
 * F: the forward cells of dnab_oracle_forward (ViterbiMatrix layout). */
/* classes/post (optional): posterior, per read base, of the class of the move that emitted it -- classes[0] = '-' (a
 * transition without input symbol), classes[1..] the input symbols, the last class '+' = tandem duplication;
 * post[p * n_classes + c], p = 0..L-1.  The same posterior weights as the counts, binned by position and symbol. */
int dnab_oracle_backward_posterior(const dnab_tables* t, const uint8_t* seq, int L, int max_sweeps, const double* F,
                                   double ll, double* ll_back, double* counts, long* total_sweeps, const char* classes,
                                   int n_classes, double* post) {
  const uint32_t n = t->n_states, k = t->k;
  int class_of[256];
  for (int c = 0; c < 256; ++c) class_of[c] = 0;
  if (post) {
    for (int c = 1; c + 1 < n_classes; ++c) class_of[(unsigned char)classes[c]] = c;
    for (size_t i = 0; i < (size_t)L * (size_t)n_classes; ++i) post[i] = 0.;
  }
  lse_table_init();
  /* outgoing lists, built from the destination-indexed tables (order: destination ascending, list order) */
  uint32_t* eoff = (uint32_t*)calloc(n + 2, sizeof(uint32_t));
  uint32_t* noff = (uint32_t*)calloc(n + 2, sizeof(uint32_t));
  for (uint32_t e = 0; e < t->n_emit; ++e) eoff[t->emit_src[e] + 2]++;
  for (uint32_t e = 0; e < t->n_null; ++e) noff[t->null_src[e] + 2]++;
  for (uint32_t s = 0; s < n; ++s) {
    eoff[s + 2] += eoff[s + 1];
    noff[s + 2] += noff[s + 1];
  }
  uint32_t* eedge = (uint32_t*)malloc(sizeof(uint32_t) * (t->n_emit ? t->n_emit : 1)); /* edge id */
  uint32_t* edst = (uint32_t*)malloc(sizeof(uint32_t) * (t->n_emit ? t->n_emit : 1));
  uint32_t* nedge = (uint32_t*)malloc(sizeof(uint32_t) * (t->n_null ? t->n_null : 1));
  uint32_t* ndst = (uint32_t*)malloc(sizeof(uint32_t) * (t->n_null ? t->n_null : 1));
  for (uint32_t d = 0; d < n; ++d) {
    for (uint32_t e = t->emit_off[d]; e < t->emit_off[d + 1]; ++e) {
      const uint32_t p = eoff[t->emit_src[e] + 1]++;
      eedge[p] = e;
      edst[p] = d;
    }
    for (uint32_t e = t->null_off[d]; e < t->null_off[d + 1]; ++e) {
      const uint32_t p = noff[t->null_src[e] + 1]++;
      nedge[p] = e;
      ndst[p] = d;
    }
  }
  const int W = (int)k + 2;
  double* Bn = (double*)malloc(sizeof(double) * (size_t)n * W); /* column pos+1 */
  double* Bc = (double*)malloc(sizeof(double) * (size_t)n * W); /* column pos   */
  double* base = (double*)malloc(sizeof(double) * n);
  double* S[2] = {(double*)malloc(sizeof(double) * n), (double*)malloc(sizeof(double) * n)};
  double* D[2] = {(double*)malloc(sizeof(double) * n), (double*)malloc(sizeof(double) * n)};
  const int nc = 5 + (int)k + 16;
  for (int i = 0; i < nc; ++i) counts[i] = 0;
  long sweeps = 0;
  int rc = 0;
#define FC(pos, s, m) F[((size_t)(pos) * n + (s)) * W + (m)]
#define POST(f, w, b) (((f) == NEG_INF || (b) == NEG_INF || (w) == NEG_INF) ? 0. : exp((f) + (w) + (b)-ll))
  for (int pos = L; pos >= 0; --pos) {
    const int xn = pos < L ? seq[pos] : 0; /* base consumed by the moves into column pos+1 */
    /* T cells of this column and the base value of S */
    for (uint32_t s = 0; s < n; ++s) {
      const uint32_t mdl = t->mdl[s];
      double* bc = Bc + (size_t)s * W;
      for (uint32_t i = 0; i < k; ++i) bc[2 + i] = NEG_INF;
      double b = NEG_INF;
      if (pos == L)
        b = (t->local || s == n - 1) ? 0. : NEG_INF;
      else {
        const double* bn = Bn + (size_t)s * W;
        if (mdl > 0) {
          bc[2] = t->sub[t->ctx[(size_t)s * k] * 4 + xn] + bn[0];
          for (uint32_t i = 0; i + 1 < mdl; ++i) bc[2 + i + 1] = t->sub[t->ctx[(size_t)s * k + i + 1] * 4 + xn] + bn[2 + i];
        }
        for (uint32_t p = eoff[s]; p < eoff[s + 1]; ++p) {
          const uint32_t e = eedge[p];
          b = lse(b, ((t->emit_score[e] + t->noGap) + t->sub[t->emit_base[e] * 4 + xn]) + Bn[(size_t)edst[p] * W]);
        }
      }
      if (pos > 0)
        for (uint32_t i = 0; i < mdl; ++i) b = lse(b, (t->tanDup + t->len[i]) + bc[2 + i]);
      base[s] = b;
    }
    /* closure by synchronous sweeps */
    int cur = 0;
    for (uint32_t s = 0; s < n; ++s) {
      S[0][s] = base[s];
      D[0][s] = NEG_INF;
    }
    for (int sweep = 0;; ++sweep) {
      if (sweep >= max_sweeps) {
        rc = 1;
        break;
      }
      const double *So = S[cur], *Do = D[cur];
      double *Sn = S[cur ^ 1], *Dn = D[cur ^ 1];
      int changed = 0;
      for (uint32_t s = 0; s < n; ++s) {
        double ns = base[s], nd = NEG_INF;
        for (uint32_t p = eoff[s]; p < eoff[s + 1]; ++p) {
          const double sc = t->emit_score[eedge[p]], bd = Do[edst[p]];
          ns = lse(ns, (t->delOpen + sc) + bd);
          nd = lse(nd, (t->delExtend + sc) + bd);
        }
        for (uint32_t p = noff[s]; p < noff[s + 1]; ++p) {
          const double sc = t->null_score[nedge[p]];
          ns = lse(ns, sc + So[ndst[p]]);
          nd = lse(nd, sc + Do[ndst[p]]);
        }
        nd = lse(nd, t->delEnd + ns);
        Sn[s] = ns;
        Dn[s] = nd;
        if (!same_bits(ns, So[s]) || !same_bits(nd, Do[s])) changed = 1;
      }
      cur ^= 1;
      ++sweeps;
      if (!changed) break;
    }
    for (uint32_t s = 0; s < n; ++s) {
      Bc[(size_t)s * W] = S[cur][s];
      Bc[(size_t)s * W + 1] = D[cur][s];
    }
    /* posterior usage of the moves leaving column pos */
    for (uint32_t s = 0; s < n; ++s) {
      const double fS = FC(pos, s, 0), fD = FC(pos, s, 1);
      const uint32_t mdl = t->mdl[s];
      const double* bc = Bc + (size_t)s * W;
      for (uint32_t p = eoff[s]; p < eoff[s + 1]; ++p) {
        const uint32_t e = eedge[p], d = edst[p];
        const double sc = t->emit_score[e];
        counts[0] += POST(fS, t->delOpen + sc, Bc[(size_t)d * W + 1]);   /* nDelOpen */
        counts[3] += POST(fD, t->delExtend + sc, Bc[(size_t)d * W + 1]); /* nDelExtend */
        if (pos < L) {
          const int b = t->emit_base[e];
          const double u = POST(fS, (sc + t->noGap) + t->sub[b * 4 + xn], Bn[(size_t)d * W]);
          counts[2] += u;                   /* nNoGap */
          counts[5 + k + b * 4 + xn] += u; /* nSub */
          if (post) post[(size_t)pos * n_classes + class_of[t->emit_in[e]]] += u;
        }
      }
      counts[4] += POST(fD, t->delEnd, bc[0]); /* nDelEnd */
      if (pos > 0)
        for (uint32_t i = 0; i < mdl; ++i) {
          const double u = POST(fS, t->tanDup + t->len[i], bc[2 + i]);
          counts[1] += u;     /* nTanDup */
          counts[5 + i] += u; /* nLen[i] */
        }
      if (pos < L && mdl > 0) {
        const double* bn = Bn + (size_t)s * W;
        const int c0 = t->ctx[(size_t)s * k];
        double dupU = POST(FC(pos, s, 2), t->sub[c0 * 4 + xn], bn[0]);
        counts[5 + k + c0 * 4 + xn] += dupU;
        for (uint32_t i = 0; i + 1 < mdl; ++i) {
          const int ci = t->ctx[(size_t)s * k + i + 1];
          const double u = POST(FC(pos, s, 2 + i + 1), t->sub[ci * 4 + xn], bn[2 + i]);
          counts[5 + k + ci * 4 + xn] += u;
          dupU += u;
        }
        if (post) post[(size_t)pos * n_classes + (n_classes - 1)] += dupU;
      }
    }
    double* tmp = Bn;
    Bn = Bc;
    Bc = tmp;
  }
  /* Bn now holds column 0 */
  if (t->local) {
    double acc = NEG_INF;
    for (uint32_t s = 0; s < n; ++s) acc = lse(acc, Bn[(size_t)s * W]);
    *ll_back = acc;
  } else
    *ll_back = Bn[0];
  if (total_sweeps) *total_sweeps = sweeps;
  free(eoff); free(noff); free(eedge); free(edst); free(nedge); free(ndst);
  free(Bn); free(Bc); free(base); free(S[0]); free(S[1]); free(D[0]); free(D[1]);
  return rc;
}

int dnab_oracle_backward_counts(const dnab_tables* t, const uint8_t* seq, int L, int max_sweeps, const double* F,
                                double ll, double* ll_back, double* counts, long* total_sweeps) {
  return dnab_oracle_backward_posterior(t, seq, L, max_sweeps, F, ll, ll_back, counts, total_sweeps, NULL, 0, NULL);
}
