/* dnab_tables.h -- the flat, destination-indexed transition/score tables that
 * cross the C ABI between the host (which keeps dnastore's Machine/JSON format)
 * and the CUDA decoder.  Plain C, plain pointers and sizes.
 *
 * One dnab_tables value is the flattened form of what the reference rebuilds PER
 * READ inside ViterbiMatrix:
 *   MachineScores  (src/viterbi.cpp:23-60, types src/viterbi.h:18-40)
 *   MutatorScores  (src/mutator.cpp:56-75, src/mutator.h:33-41)
 *   maxDupLen k    (src/viterbi.cpp:63)  and  MutatorParams::local (src/mutator.h:12)
 * The host computes every double with the same libm calls and summation order as
 * the reference (src/viterbi.cpp:6-14,41; src/mutator.cpp:56-75), so the device
 * only ever needs fp64 +, max and >.
 *
 * Edge lists are CSR by DESTINATION state and keep the reference's list order
 * (source index ascending, then transition index within the source,
 * src/viterbi.cpp:30-58): that order is the traceback tie-break order
 * (src/viterbi.cpp:219,254-276).
 */
#ifndef DNAB_TABLES_H
#define DNAB_TABLES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dnab_tables {
  uint32_t n_states;  /* Machine::nStates(); start = 0, end = n_states-1 (viterbi.h:83,102) */
  uint32_t k;         /* min(machine.maxLeftContext(), params.maxDupLen()) (viterbi.cpp:63) */
  uint32_t local;     /* MutatorParams::local: 0 = --error-global */
  uint32_t n_emit;    /* transitions WITH a DNA output kept by the decoder (viterbi.cpp:38,52-55) */
  uint32_t n_null;    /* kept transitions WITHOUT a DNA output (viterbi.cpp:49-51) */

  /* incomingEmit, CSR by destination, reference list order */
  const uint32_t* emit_off;   /* [n_states+1] */
  const uint32_t* emit_src;   /* [n_emit] IncomingTransScore::src */
  const double*   emit_score; /* [n_emit] log(symProb[in]) or 0 for null input (viterbi.cpp:41) */
  const uint8_t*  emit_base;  /* [n_emit] charToBase(t.out), 0..3 = ACGT */
  const uint8_t*  emit_in;    /* [n_emit] input symbol character, 0 = none */

  /* incomingNull, same layout */
  const uint32_t* null_off;   /* [n_states+1] */
  const uint32_t* null_src;   /* [n_null] */
  const double*   null_score; /* [n_null] */
  const uint8_t*  null_in;    /* [n_null] */

  /* tandem-duplication context: ctx[s*k+i] = tanDupBase(ss,i) = leftContext[size-1-i]
   * over the non-'*' left-context characters (viterbi.h:104-105, viterbi.cpp:33-36);
   * mdl[s] = min(k, number of non-'*' characters). Entries i >= mdl[s] are 0. */
  const uint8_t* ctx;         /* [n_states*k] */
  const uint8_t* mdl;         /* [n_states] */

  /* MutatorScores (mutator.cpp:56-75) */
  double noGap, delOpen, delExtend, delEnd, tanDup;
  double sub[16];             /* sub[base*4+observed], log-odds against 1/4 */
  const double* len;          /* [k] log pLen[i] (only the first k entries are used) */
} dnab_tables;

#ifdef __cplusplus
}
#endif
#endif /* DNAB_TABLES_H */
