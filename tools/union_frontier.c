/* union_frontier.c -- ANALYSIS TOOL (not product, not oracle): how much does the closure's work grow when R
 * reads share one frontier?
 *
 * The read-batched fill kernel (csrc/viterbi_fill_batch.cu) makes the reads the SIMD lanes: a state is
 * relaxed for all R reads of a group whenever ANY of them raised it.  This tool replays the closure of
 * src/viterbi.cpp:110-159 level by level (breadth-first: level 0 pushes every state, level n+1 pushes the
 * states raised during level n) for every read of a group on the CPU and counts, per column,
 *   - the per-read visits   (sum over levels of the read's own frontier)
 *   - the union visits      (sum over levels of |union over the group's reads of the frontiers|)
 *   - the levels            (per read, and of the group = the maximum)
 * The emission step uses the fill's arithmetic (src/viterbi.cpp:92-106) so the S columns are the real ones.
 *
 * Build: gcc -O2 -shared -fPIC -ffp-contract=off -o tools/_build/libunion_frontier.so tools/union_frontier.c
 * Driven by tools/union_frontier.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dnab_tables.h"

static inline double dmax(double a, double b) { return (a < b) ? b : a; }

typedef struct {
  uint32_t *eoff, *edst, *noff, *ndst;
  double *esc, *nsc;
} outg;

static void build_out(uint32_t n, const uint32_t* in_off, const uint32_t* in_src, const double* in_score, uint32_t ne,
                      uint32_t** off_, uint32_t** dst_, double** sc_) {
  uint32_t* off = (uint32_t*)calloc(n + 2, sizeof(uint32_t));
  uint32_t* dst = (uint32_t*)malloc((ne ? ne : 1) * sizeof(uint32_t));
  double* sc = (double*)malloc((ne ? ne : 1) * sizeof(double));
  for (uint32_t e = 0; e < ne; ++e) off[in_src[e] + 2]++;
  for (uint32_t s = 0; s < n; ++s) off[s + 2] += off[s + 1];
  for (uint32_t d = 0; d < n; ++d)
    for (uint32_t e = in_off[d]; e < in_off[d + 1]; ++e) {
      const uint32_t p = off[in_src[e] + 1]++;
      dst[p] = d;
      sc[p] = in_score[e];
    }
  *off_ = off;
  *dst_ = dst;
  *sc_ = sc;
}

/* seqs: concatenated tokens, offsets[R+1].  out[0] = sum over columns of per-read visits (all reads),
 * out[1] = sum over columns of union visits, out[2] = sum over columns of group levels,
 * out[3] = sum over columns and reads of per-read levels, out[4] = columns (of the longest read),
 * out[5] = sum over columns of states (n * columns), out[6] = union visits at level >= 1 only. */
int union_frontier(const dnab_tables* t, const uint8_t* seqs, const int64_t* offsets, int R, double* out) {
  const uint32_t n = t->n_states, k = t->k;
  outg g;
  build_out(n, t->emit_off, t->emit_src, t->emit_score, t->n_emit, &g.eoff, &g.edst, &g.esc);
  build_out(n, t->null_off, t->null_src, t->null_score, t->n_null, &g.noff, &g.ndst, &g.nsc);
  int Lmax = 0;
  for (int r = 0; r < R; ++r) Lmax = (int)(offsets[r + 1] - offsets[r]) > Lmax ? (int)(offsets[r + 1] - offsets[r]) : Lmax;
  /* per read: S(prev), S, D, T[k] (prev and cur) */
  const size_t col = (size_t)n;
  double* Sp = (double*)malloc(sizeof(double) * col * R);
  double* S = (double*)malloc(sizeof(double) * col * R);
  double* D = (double*)malloc(sizeof(double) * col * R);
  double* Tp = (double*)malloc(sizeof(double) * col * R * (k ? k : 1));
  double* T = (double*)malloc(sizeof(double) * col * R * (k ? k : 1));
  uint8_t* cur = (uint8_t*)malloc((size_t)n * R);  /* frontier flags per read */
  uint8_t* nxt = (uint8_t*)malloc((size_t)n * R);
  memset(out, 0, 8 * sizeof(double));
  for (size_t i = 0; i < col * R; ++i) Sp[i] = -INFINITY;
  for (size_t i = 0; i < col * R * (k ? k : 1); ++i) Tp[i] = -INFINITY;
  for (int pos = 0; pos <= Lmax; ++pos) {
    int active = 0;
    for (int r = 0; r < R; ++r) {
      const int L = (int)(offsets[r + 1] - offsets[r]);
      double *s = S + col * r, *sp = Sp + col * r, *d = D + col * r;
      double *tt = T + col * r * k, *tp = Tp + col * r * k;
      if (pos > L) {
        memset(cur + (size_t)n * r, 0, n);
        continue;
      }
      ++active;
      const uint8_t x = pos > 0 ? seqs[offsets[r] + pos - 1] : 0;
      for (uint32_t st = 0; st < n; ++st) {
        double v = -INFINITY;
        if (pos == 0) v = (t->local || st == 0) ? 0.0 : -INFINITY;
        if (pos > 0) {
          for (uint32_t e = t->emit_off[st]; e < t->emit_off[st + 1]; ++e)
            v = dmax(v, sp[t->emit_src[e]] + t->emit_score[e] + t->noGap + t->sub[t->emit_base[e] * 4 + x]);
          const int mdl = t->mdl[st];
          if (mdl > 0) {
            v = dmax(v, tp[(size_t)st * k] + t->sub[t->ctx[(size_t)st * k] * 4 + x]);
            for (int i = 0; i < (int)k; ++i) tt[(size_t)st * k + i] = -INFINITY;
            for (int i = 0; i < mdl - 1; ++i)
              tt[(size_t)st * k + i] = tp[(size_t)st * k + i + 1] + t->sub[t->ctx[(size_t)st * k + i + 1] * 4 + x];
          }
        }
        s[st] = v;
        d[st] = -INFINITY;
      }
      memset(cur + (size_t)n * r, 1, n);
    }
    if (!active) break;
    out[4] += 1;
    out[5] += n;
    int glevels = 0;
    for (int level = 0;; ++level) {
      /* union of the frontiers */
      long uni = 0, any = 0;
      for (uint32_t st = 0; st < n; ++st) {
        int u = 0;
        for (int r = 0; r < R; ++r) u |= cur[(size_t)n * r + st];
        uni += u;
      }
      if (!uni) break;
      out[1] += uni;
      if (level >= 1) out[6] += uni;
      ++glevels;
      for (int r = 0; r < R; ++r) {
        uint8_t *c = cur + (size_t)n * r, *nx = nxt + (size_t)n * r;
        double *s = S + col * r, *d = D + col * r;
        long mine = 0;
        memset(nx, 0, n);
        for (uint32_t st = 0; st < n; ++st) {
          if (!c[st]) continue;
          ++mine;
          const double dsrc = d[st];
          const double ssrc = dmax(s[st], dsrc + t->delEnd);
          s[st] = ssrc;
          for (uint32_t e = g.eoff[st]; e < g.eoff[st + 1]; ++e) {
            const double dsc = dmax(dsrc + t->delExtend, ssrc + t->delOpen) + g.esc[e];
            if (dsc > d[g.edst[e]]) {
              d[g.edst[e]] = dsc;
              nx[g.edst[e]] = 1;
            }
          }
          for (uint32_t e = g.noff[st]; e < g.noff[st + 1]; ++e) {
            const double dsc = dsrc + g.nsc[e], ssc = ssrc + g.nsc[e];
            if (dsc > d[g.ndst[e]]) {
              d[g.ndst[e]] = dsc;
              nx[g.ndst[e]] = 1;
            }
            if (ssc > s[g.ndst[e]]) {
              s[g.ndst[e]] = ssc;
              nx[g.ndst[e]] = 1;
            }
          }
        }
        out[0] += mine;
        if (mine) out[3] += 1;
        any |= mine;
      }
      uint8_t* tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
    out[2] += glevels;
    /* duplication opens, then rotate */
    for (int r = 0; r < R; ++r) {
      const int L = (int)(offsets[r + 1] - offsets[r]);
      if (pos > L) continue;
      double *s = S + col * r, *tt = T + col * r * k;
      if (pos > 0)
        for (uint32_t st = 0; st < n; ++st)
          for (int i = 0; i < t->mdl[st]; ++i)
            tt[(size_t)st * k + i] = dmax(tt[(size_t)st * k + i], s[st] + t->tanDup + t->len[i]);
      else
        for (size_t i = 0; i < col * k; ++i) tt[i] = -INFINITY;
      memcpy(Sp + col * r, s, sizeof(double) * col);
      memcpy(Tp + col * r * k, tt, sizeof(double) * col * k);
    }
  }
  free(Sp); free(S); free(D); free(Tp); free(T); free(cur); free(nxt);
  free(g.eoff); free(g.edst); free(g.esc); free(g.noff); free(g.ndst); free(g.nsc);
  return 0;
}
