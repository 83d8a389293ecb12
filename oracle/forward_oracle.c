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
