/* pairhmm_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker for SURVEY.md 8a-10 / 8a-11).
 *
 * Plain-C restatement of the reference's pair-HMM forward / backward / expected counts
 * over the (original DNA x observed DNA) lattice of a 2-row alignment, with the
 * reference's TABLE-BASED log-sum-exp.  Only tests/ (and bench legs that time the CPU
 * reference) may load this; the product never links or falls back to it.
 *
 * Parity pin: tests/test_pairhmm.py checks it against the UNMODIFIED reference
 * (oracle/_ref/refdriver fb|fit, i.e. FwdBackMatrix / baumWelchParams themselves) --
 * forward and backward log-likelihoods and every expected count bit for bit -- and
 * against the reference's own goldens data/dup*.counts*.json, tiny/test.params.json
 * (reference Makefile:156-163).
 *
 * What each function follows (line numbers in /root/reference/src/):
 *   lse_table_init / lse_unary / lse   logsumexp.h:19-74, logsumexp.cpp:5-15,44-46
 *   scores_init                         mutator.cpp:56-75
 *   in_range                            alignpath.h:48-53 (a[] = cumulativeMatches[row1PosToCol[ip]],
 *                                       b[] = cumulativeMatches[row2PosToCol[op]], built by the caller)
 *   forward                             fwdback.cpp:43-78
 *   backward                            fwdback.cpp:80-116
 *   counts                              fwdback.cpp:154-188 with fwdback.h:92-113
 * Cells outside the envelope read as -inf exactly like the reference's dummyCell
 * (fwdback.h:49-53); storage here is dense.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define LSE_MAX 10
#define LSE_PRECISION .0001
#define LSE_ENTRIES (((int)(LSE_MAX / LSE_PRECISION)) + 1)
#define NEG_INF (-INFINITY)

static double* lse_lookup = 0;

static void lse_table_init(void) {
  if (lse_lookup) return;
  lse_lookup = (double*)malloc(sizeof(double) * LSE_ENTRIES);
  for (int n = 0; n < LSE_ENTRIES; ++n) {
    const double x = n * LSE_PRECISION;
    lse_lookup[n] = log(1. + exp(-x));
  }
}

/* the table itself, for the product to be checked against (dnab_lse_table) */
const double* dnab_oracle_lse_table(int* entries) {
  lse_table_init();
  *entries = LSE_ENTRIES;
  return lse_lookup;
}

static double lse_unary(double x) {
  if (x >= LSE_MAX || isnan(x) || isinf(x)) return 0;
  if (x < 0) return -x;
  const int n = (int)(x / LSE_PRECISION);
  const double dx = x - (n * LSE_PRECISION);
  const double f0 = lse_lookup[n], f1 = lse_lookup[n + 1];
  const double df = f1 - f0;
  return f0 + df * (dx / LSE_PRECISION);
}

double dnab_oracle_lse(double a, double b) {
  double mx, diff;
  lse_table_init();
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
#define lse dnab_oracle_lse

typedef struct {
  double delOpen, tanDup, noGap, delExtend, delEnd, sub[16], len[32];
} scores_t;

static void scores_init(scores_t* s, const double* p, const double* pLen, int k) {
  const double pDelOpen = p[0], pDelExtend = p[1], pTanDup = p[2], pTransition = p[3], pTransversion = p[4];
  s->delOpen = log(pDelOpen);
  s->tanDup = log(pTanDup);
  s->noGap = log(1. - pDelOpen - pTanDup);
  s->delExtend = log(pDelExtend);
  s->delEnd = log(1. - pDelExtend);
  const double nullScore = log(1. / 4.);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      s->sub[i * 4 + j] = (i == j ? log(1. - pTransition - pTransversion)
                                  : ((i != j && (i & 1) == (j & 1)) ? log(pTransition) : log(pTransversion / 2))) -
                          nullScore;
  for (int l = 0; l < k; ++l) s->len[l] = log(pLen[l]);
}

/* Forward, backward and counts of ONE alignment.
 * p = {pDelOpen, pDelExtend, pTanDup, pTransition, pTransversion}; k = maxDupLen = |pLen|;
 * in/out = tokens 0..3; a[0..inLen], b[0..outLen] = envelope coordinates; maxDist = 0 (strict) or k.
 * counts = {nDelOpen, nTanDup, nNoGap, nDelExtend, nDelEnd, nLen[k], nSub[16]}.
 * fcells/bcells (optional) receive the dense matrices [(inLen+1)][(outLen+1)][2+k]. */
int dnab_oracle_pairhmm_fb(const double* p, const double* pLen, int k, const uint8_t* in, int inLen, const uint8_t* out,
                           int outLen, const int32_t* a, const int32_t* b, int maxDist, double* fwdLL, double* backLL,
                           double* counts, double* fcells, double* bcells) {
  lse_table_init();
  scores_t sc;
  scores_init(&sc, p, pLen, k);
  const int W = 2 + k;
  const size_t nCells = (size_t)(inLen + 1) * (outLen + 1) * W;
  double* F = fcells ? fcells : (double*)malloc(nCells * sizeof(double));
  double* B = bcells ? bcells : (double*)malloc(nCells * sizeof(double));
  for (size_t i = 0; i < nCells; ++i) F[i] = B[i] = NEG_INF;
#define IDX(ip, op) (((size_t)(ip) * (outLen + 1) + (op)) * W)
#define INR(ip, op) (abs(a[ip] - b[op]) <= maxDist)
#define MDL(ip) ((k) < (ip) ? (k) : (ip))
#define SUB(ip, op) sc.sub[in[(ip)-1] * 4 + out[(op)-1]]
#define TSUB(ip, op, d) sc.sub[in[(ip)-1 - (d)] * 4 + out[(op)-1]]

  /* forward, fwdback.cpp:46-76 */
  F[IDX(0, 0)] = 0;
  for (int ip = 0; ip <= inLen; ++ip)
    for (int op = 0; op <= outLen; ++op)
      if (INR(ip, op)) {
        double* cell = F + IDX(ip, op);
        if (ip > 0 && op > 0) {
          if (INR(ip - 1, op - 1)) cell[0] = F[IDX(ip - 1, op - 1)] + sc.noGap + SUB(ip, op);
          if (INR(ip, op - 1)) {
            const double* ins = F + IDX(ip, op - 1);
            for (int d = 0; d < MDL(ip) - 1; ++d) cell[2 + d] = ins[2 + d + 1] + TSUB(ip, op, d + 1);
            cell[0] = lse(cell[0], ins[2] + TSUB(ip, op, 0));
          }
        }
        if (ip > 0 && INR(ip - 1, op)) {
          const double* del = F + IDX(ip - 1, op);
          cell[1] = lse(del[0] + sc.delOpen, del[1] + sc.delExtend);
        }
        cell[0] = lse(cell[0], cell[1] + sc.delEnd);
        for (int d = 0; d < MDL(ip); ++d) cell[2 + d] = lse(cell[2 + d], cell[0] + sc.tanDup + sc.len[d]);
      }
  const double ll = F[IDX(inLen, outLen)];
  *fwdLL = ll;

  /* backward, fwdback.cpp:84-114 */
  B[IDX(inLen, outLen)] = 0;
  for (int ip = inLen; ip >= 0; --ip)
    for (int op = outLen; op >= 0; --op)
      if (INR(ip, op)) {
        double* cell = B + IDX(ip, op);
        if (op < outLen) {
          if (ip < inLen && INR(ip + 1, op + 1)) cell[0] = sc.noGap + SUB(ip + 1, op + 1) + B[IDX(ip + 1, op + 1)];
          if (ip > 0 && INR(ip, op + 1)) {
            const double* ins = B + IDX(ip, op + 1);
            for (int d = 1; d < MDL(ip); ++d) cell[2 + d] = TSUB(ip, op + 1, d) + ins[2 + d - 1];
            cell[2] = TSUB(ip, op + 1, 0) + ins[0];
          }
        }
        if (ip < inLen && INR(ip + 1, op)) {
          const double* del = B + IDX(ip + 1, op);
          cell[0] = lse(cell[0], sc.delOpen + del[1]);
          cell[1] = sc.delExtend + del[1];
        }
        for (int d = 0; d < MDL(ip); ++d) cell[0] = lse(cell[0], cell[2 + d] + sc.tanDup + sc.len[d]);
        cell[1] = lse(cell[1], cell[0] + sc.delEnd);
      }
  *backLL = B[IDX(0, 0)];

  /* counts, fwdback.cpp:154-188; out-of-lattice neighbours read as -inf (dummyCell) */
  double *nDelOpen = counts, *nTanDup = counts + 1, *nNoGap = counts + 2, *nDelExtend = counts + 3, *nDelEnd = counts + 4;
  double *nLen = counts + 5, *nSub = counts + 5 + k;
  for (int i = 0; i < 5 + k + 16; ++i) counts[i] = 0;
  for (int ip = 0; ip <= inLen; ++ip)
    for (int op = 0; op <= outLen; ++op)
      if (INR(ip, op)) {
        const double* bc = B + IDX(ip, op);
        if (ip > 0 && op > 0) {
          const double c = exp(F[IDX(ip - 1, op - 1)] + sc.noGap + SUB(ip, op) + bc[0] - ll);
          *nNoGap += c;
          nSub[in[ip - 1] * 4 + out[op - 1]] += c;
          const double* fi = F + IDX(ip, op - 1);
          for (int d = 0; d < MDL(ip) - 1; ++d) {
            const double ci = exp(fi[2 + d + 1] + TSUB(ip, op, d + 1) + bc[2 + d] - ll);
            nSub[in[ip - 1 - (d + 1)] * 4 + out[op - 1]] += ci;
          }
          const double c0 = exp(fi[2] + TSUB(ip, op, 0) + bc[0] - ll);
          nSub[in[ip - 1] * 4 + out[op - 1]] += c0;
        }
        if (ip > 0) {
          const double* fd = F + IDX(ip - 1, op);
          *nDelOpen += exp(fd[0] + sc.delOpen + bc[1] - ll);
          *nDelExtend += exp(fd[1] + sc.delExtend + bc[1] - ll);
        }
        const double* fc = F + IDX(ip, op);
        *nDelEnd += exp(fc[1] + sc.delEnd + bc[0] - ll);
        for (int d = 0; d < MDL(ip); ++d) {
          const double c = exp(fc[0] + sc.tanDup + sc.len[d] + bc[2 + d] - ll);
          *nTanDup += c;
          nLen[d] += c;
        }
      }
  if (!fcells) free(F);
  if (!bcells) free(B);
  return 0;
}
