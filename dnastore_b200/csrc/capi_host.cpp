// Host half of the C ABI (include/dnastore_b200.h): machines, error model, table
// compilation, read packing and the decodeFastSeqs drop-in.  Exceptions from the
// C++ host classes are turned into error codes here; none crosses the boundary.
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dnastore_b200.h"
#include "capi_error.h"
#include "capi_types.h"
#include "host/exact_decoder.h"
#include "host/fasta.h"
#include "host/machine.h"
#include "host/pairhmm.h"
#include "host/tables.h"

namespace dnab {
static thread_local std::string g_lastError;
void setLastError(const std::string& msg) { g_lastError = msg; }
}  // namespace dnab

using namespace dnab;

struct dnab_machine {
  Machine m;
};
struct dnab_compiled {
  CompiledTables c;
};
struct dnab_exact_decoder {
  ExactDecoder d;
  explicit dnab_exact_decoder(const Machine& m) : d(m) {}
};
struct dnab_pair_db {
  std::vector<PairAlignment> aligns;
};

template <class F>
static auto guarded(F&& f, decltype(f()) onError) -> decltype(f()) {
  try {
    return f();
  } catch (const std::exception& e) {
    setLastError(e.what());
  } catch (...) {
    setLastError("unknown error");
  }
  return onError;
}

extern "C" {

const char* dnab_last_error(void) { return g_lastError.c_str(); }
const char* dnab_version(void) { return "dnastore_b200 0.1 (sm_100a)"; }
void dnab_free(void* p) { std::free(p); }

dnab_machine* dnab_machine_load(const char* json_path) {
  return guarded([&]() { return new dnab_machine{Machine::fromFile(json_path)}; }, (dnab_machine*)nullptr);
}
dnab_machine* dnab_machine_from_json(const char* json_text) {
  return guarded([&]() { return new dnab_machine{Machine::fromJSONText(json_text)}; }, (dnab_machine*)nullptr);
}
dnab_machine* dnab_machine_compose(const dnab_machine* first, const dnab_machine* second) {
  if (!first || !second) {
    setLastError("dnab_machine_compose: null machine");
    return nullptr;
  }
  return guarded([&]() { return new dnab_machine{Machine::compose(first->m, second->m)}; }, (dnab_machine*)nullptr);
}
char* dnab_machine_to_json(const dnab_machine* m) {
  if (!m) return nullptr;
  return guarded(
      [&]() {
        const std::string s = m->m.toJSON();
        char* out = (char*)std::malloc(s.size() + 1);
        if (!out) throw std::bad_alloc();
        std::memcpy(out, s.c_str(), s.size() + 1);
        return out;
      },
      (char*)nullptr);
}
uint32_t dnab_machine_n_states(const dnab_machine* m) { return m ? (uint32_t)m->m.nStates() : 0; }
uint32_t dnab_machine_max_left_context(const dnab_machine* m) { return m ? (uint32_t)m->m.maxLeftContext() : 0; }
int dnab_machine_input_alphabet(const dnab_machine* m, int flags, char* out, size_t cap) {
  if (!m || !out || !cap) return DNAB_EINVAL;
  const std::string a = m->m.inputAlphabet(flags);
  if (a.size() + 1 > cap) {
    setLastError("dnab_machine_input_alphabet: buffer too small");
    return DNAB_EINVAL;
  }
  std::memcpy(out, a.c_str(), a.size() + 1);
  return DNAB_OK;
}
void dnab_machine_free(dnab_machine* m) { delete m; }

void dnab_error_flags_default(dnab_error_flags* f) {
  if (!f) return;
  f->length = 12;
  f->global = 0;
  f->sub_prob = .01;
  f->iv_ratio = 10;
  f->dup_prob = .001;
  f->del_open = .001;
  f->del_ext = .01;
}

static dnab_compiled* compileWith(const dnab_machine* m, const MutatorParams& params) {
  auto* c = new dnab_compiled();
  try {
    compileTables(m->m, params, c->c);
  } catch (...) {
    delete c;
    throw;
  }
  return c;
}

dnab_compiled* dnab_compile(const dnab_machine* m, const dnab_error_flags* f) {
  if (!m || !f) {
    setLastError("dnab_compile: null argument");
    return nullptr;
  }
  return guarded(
      [&]() {
        return compileWith(m, MutatorParams::fromFlags(f->length, f->sub_prob, f->iv_ratio, f->dup_prob, f->del_open,
                                                       f->del_ext, f->global != 0));
      },
      (dnab_compiled*)nullptr);
}
dnab_compiled* dnab_compile_with_error_file(const dnab_machine* m, const char* error_json_path) {
  if (!m || !error_json_path) {
    setLastError("dnab_compile_with_error_file: null argument");
    return nullptr;
  }
  return guarded([&]() { return compileWith(m, MutatorParams::fromFile(error_json_path)); }, (dnab_compiled*)nullptr);
}
const dnab_tables* dnab_compiled_tables(const dnab_compiled* c) { return c ? &c->c.t : nullptr; }
void dnab_compiled_free(dnab_compiled* c) { delete c; }

size_t dnab_packed_size(const int32_t* read_len, int64_t n_reads) { return packedSize(read_len, n_reads); }

int dnab_pack_reads(const char* bases, const int64_t* base_off, int64_t n_reads, uint8_t* packed, int64_t* read_byte_off,
                    int32_t* read_len) {
  char bad = 0;
  const int64_t r = packReads(bases, base_off, n_reads, packed, read_byte_off, read_len, &bad);
  if (r >= 0) {
    setLastError(std::string("Unknown symbol ") + bad + " in sequence " + std::to_string(r) + " (alphabet is ACGT)");
    return DNAB_EFORMAT;
  }
  return DNAB_OK;
}

int64_t dnab_decoded_count(const dnab_decoded_set* s) { return s ? (int64_t)s->seqs.size() : 0; }
const char* dnab_decoded_name(const dnab_decoded_set* s, int64_t i) { return s->names[i].c_str(); }
const char* dnab_decoded_seq(const dnab_decoded_set* s, int64_t i) { return s->seqs[i].c_str(); }
double dnab_decoded_loglike(const dnab_decoded_set* s, int64_t i) { return s->loglike[i]; }
int32_t dnab_decoded_status(const dnab_decoded_set* s, int64_t i) { return s->status[i]; }
void dnab_decoded_free(dnab_decoded_set* s) { delete s; }

// ---- pair-HMM forward/backward ------------------------------------------------------------------
static MutatorParams toParams(const dnab_mutator_params* p) {
  MutatorParams m;
  m.pDelOpen = p->p_del_open;
  m.pDelExtend = p->p_del_extend;
  m.pTanDup = p->p_tan_dup;
  m.pTransition = p->p_transition;
  m.pTransversion = p->p_transversion;
  m.pLen.assign(p->p_len, p->p_len + p->max_dup_len);
  m.local = p->local != 0;
  return m;
}
static void fromParams(const MutatorParams& m, dnab_mutator_params* p) {
  std::memset(p, 0, sizeof *p);
  p->p_del_open = m.pDelOpen;
  p->p_del_extend = m.pDelExtend;
  p->p_tan_dup = m.pTanDup;
  p->p_transition = m.pTransition;
  p->p_transversion = m.pTransversion;
  p->max_dup_len = (int32_t)m.maxDupLen();
  for (size_t i = 0; i < m.pLen.size() && i < DNAB_MAX_DUP; ++i) p->p_len[i] = m.pLen[i];
  p->local = m.local ? 1 : 0;
}
static void fromCounts(const MutatorCounts& m, dnab_mutator_counts* c) {
  std::memset(c, 0, sizeof *c);
  c->n_del_open = m.nDelOpen;
  c->n_tan_dup = m.nTanDup;
  c->n_no_gap = m.nNoGap;
  c->n_del_extend = m.nDelExtend;
  c->n_del_end = m.nDelEnd;
  c->max_dup_len = (int32_t)m.nLen.size();
  for (size_t i = 0; i < m.nLen.size() && i < DNAB_MAX_DUP; ++i) c->n_len[i] = m.nLen[i];
  for (int i = 0; i < 16; ++i) c->n_sub[i] = m.nSub[i];
}
static MutatorCounts toCounts(const dnab_mutator_counts* c) {
  MutatorCounts m((size_t)c->max_dup_len);
  m.nDelOpen = c->n_del_open;
  m.nTanDup = c->n_tan_dup;
  m.nNoGap = c->n_no_gap;
  m.nDelExtend = c->n_del_extend;
  m.nDelEnd = c->n_del_end;
  for (int i = 0; i < c->max_dup_len; ++i) m.nLen[i] = c->n_len[i];
  for (int i = 0; i < 16; ++i) m.nSub[i] = c->n_sub[i];
  return m;
}
static char* dupString(const std::string& s) {
  char* out = (char*)std::malloc(s.size() + 1);
  if (out) std::memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

void dnab_mutator_params_from_flags(const dnab_error_flags* f, dnab_mutator_params* out) {
  if (!f || !out) return;
  fromParams(MutatorParams::fromFlags(f->length, f->sub_prob, f->iv_ratio, f->dup_prob, f->del_open, f->del_ext,
                                      f->global != 0), out);
}
char* dnab_mutator_params_json(const dnab_mutator_params* p) { return p ? dupString(toParams(p).asJSON()) : nullptr; }
char* dnab_mutator_counts_json(const dnab_mutator_counts* c) { return c ? dupString(toCounts(c).asJSON()) : nullptr; }
const double* dnab_lse_table(int32_t* n_entries) {
  const std::vector<double>& t = logSumExpLookupTable();
  if (n_entries) *n_entries = (int32_t)t.size();
  return t.data();
}

dnab_pair_db* dnab_pair_db_load(const char* stockholm_path) {
  if (!stockholm_path) return nullptr;
  return guarded(
      [&]() {
        auto* db = new dnab_pair_db();
        try {
          for (const auto& s : readStockholmDatabase(stockholm_path)) db->aligns.push_back(makePairAlignment(s));
        } catch (...) {
          delete db;
          throw;
        }
        return db;
      },
      (dnab_pair_db*)nullptr);
}
int64_t dnab_pair_db_count(const dnab_pair_db* db) { return db ? (int64_t)db->aligns.size() : 0; }
int32_t dnab_pair_db_in_len(const dnab_pair_db* db, int64_t i) { return (int32_t)db->aligns[i].in.size(); }
int32_t dnab_pair_db_out_len(const dnab_pair_db* db, int64_t i) { return (int32_t)db->aligns[i].out.size(); }
const uint8_t* dnab_pair_db_in(const dnab_pair_db* db, int64_t i) { return db->aligns[i].in.data(); }
const uint8_t* dnab_pair_db_out(const dnab_pair_db* db, int64_t i) { return db->aligns[i].out.data(); }
const int32_t* dnab_pair_db_env_a(const dnab_pair_db* db, int64_t i) { return db->aligns[i].a.data(); }
const int32_t* dnab_pair_db_env_b(const dnab_pair_db* db, int64_t i) { return db->aligns[i].b.data(); }
void dnab_pair_db_free(dnab_pair_db* db) { delete db; }

int dnab_pairhmm_set_chunk_cells(int64_t cells) {
  dnab::setPairHmmChunkCells(cells);
  return DNAB_OK;
}

int dnab_pairhmm_fb_batch(int device, const dnab_mutator_params* p, int strict, int64_t n_align, const uint8_t* in_tok,
                          const int64_t* in_off, const uint8_t* out_tok, const int64_t* out_off, const int32_t* env_a,
                          const int32_t* env_b, double* fwd_ll, double* back_ll, dnab_mutator_counts* counts,
                          double* kernel_ms) {
  if (!p || n_align < 0 || p->max_dup_len < 0 || p->max_dup_len > DNAB_MAX_DUP) {
    setLastError("dnab_pairhmm_fb_batch: bad argument");
    return DNAB_EINVAL;
  }
  return guarded(
      [&]() {
        std::vector<PairAlignment> aligns((size_t)n_align);
        for (int64_t i = 0; i < n_align; ++i) {
          PairAlignment& a = aligns[(size_t)i];
          a.in.assign(in_tok + in_off[i], in_tok + in_off[i + 1]);
          a.out.assign(out_tok + out_off[i], out_tok + out_off[i + 1]);
          a.a.assign(env_a + in_off[i] + i, env_a + in_off[i + 1] + i + 1);
          a.b.assign(env_b + out_off[i] + i, env_b + out_off[i + 1] + i + 1);
        }
        std::vector<double> f, b;
        std::vector<MutatorCounts> c;
        if (!pairHmmFwdBackBatch(device, toParams(p), strict != 0, aligns, f, b, c, kernel_ms)) return (int)DNAB_ECUDA;
        for (int64_t i = 0; i < n_align; ++i) {
          if (fwd_ll) fwd_ll[i] = f[(size_t)i];
          if (back_ll) back_ll[i] = b[(size_t)i];
          if (counts) fromCounts(c[(size_t)i], &counts[i]);
        }
        return (int)DNAB_OK;
      },
      (int)DNAB_EINVAL);
}

int dnab_expected_counts(int device, const dnab_mutator_params* p, const dnab_pair_db* db, int strict,
                         dnab_mutator_counts* total, double* loglike) {
  if (!p || !db || !total || !loglike) return DNAB_EINVAL;
  return guarded(
      [&]() {
        MutatorCounts t;
        double ll = 0;
        if (!expectedCounts(device, toParams(p), db->aligns, strict != 0, t, ll)) return (int)DNAB_ECUDA;
        fromCounts(t, total);
        *loglike = ll;
        return (int)DNAB_OK;
      },
      (int)DNAB_EINVAL);
}

int dnab_baum_welch(int device, const dnab_mutator_params* init, const dnab_pair_db* db, int strict,
                    dnab_mutator_params* fitted, int32_t* iterations) {
  if (!init || !db || !fitted) return DNAB_EINVAL;
  return guarded(
      [&]() {
        const MutatorParams start = toParams(init);
        MutatorCounts prior(start.maxDupLen());
        prior.initLaplace();
        MutatorParams fit;
        int iters = 0;
        if (!baumWelchParams(device, start, prior, db->aligns, strict != 0, fit, &iters)) return (int)DNAB_ECUDA;
        fromParams(fit, fitted);
        if (iterations) *iterations = iters;
        return (int)DNAB_OK;
      },
      (int)DNAB_EINVAL);
}

/* ---- exact decoding (host/exact_decoder.h) ---- */
static std::string joinLines(const std::vector<std::string>& lines) {
  std::string all;
  for (const auto& l : lines) all += l + "\n";
  return all;
}

dnab_exact_decoder* dnab_exact_decoder_create(const dnab_machine* m) {
  if (!m) {
    setLastError("dnab_exact_decoder_create: null machine");
    return nullptr;
  }
  return guarded([&]() { return new dnab_exact_decoder(m->m); }, (dnab_exact_decoder*)nullptr);
}
int dnab_exact_decoder_feed(dnab_exact_decoder* d, const char* bases, size_t n) {
  if (!d || (!bases && n)) return DNAB_EINVAL;
  return guarded(
      [&]() {
        for (size_t i = 0; i < n; ++i) d->d.decodeSymbol(bases[i]);
        return (int)DNAB_OK;
      },
      (int)DNAB_EINVAL);
}
int dnab_exact_decoder_close(dnab_exact_decoder* d) {
  if (!d) return DNAB_EINVAL;
  return guarded(
      [&]() {
        d->d.close();
        return (int)DNAB_OK;
      },
      (int)DNAB_EINVAL);
}
char* dnab_exact_decoder_take_symbols(dnab_exact_decoder* d) {
  if (!d) return nullptr;
  return guarded([&]() { return dupString(d->d.takeSymbols()); }, (char*)nullptr);
}
char* dnab_exact_decoder_warnings(const dnab_exact_decoder* d) {
  if (!d) return nullptr;
  return guarded([&]() { return dupString(joinLines(d->d.warnings())); }, (char*)nullptr);
}
int64_t dnab_exact_decoder_hypotheses(const dnab_exact_decoder* d) { return d ? (int64_t)d->d.hypotheses() : 0; }
void dnab_exact_decoder_destroy(dnab_exact_decoder* d) { delete d; }

int64_t dnab_pack_decoded_symbols(const char* symbols, size_t n, uint8_t* bytes, size_t cap, char* leftover_bits,
                                  char** warnings) {
  if (!symbols && n) return DNAB_EINVAL;
  return guarded(
      [&]() -> int64_t {
        const PackedBits p = packDecodedSymbols(std::string(symbols ? symbols : "", n));
        if (p.bytes.size() > cap) {
          setLastError("dnab_pack_decoded_symbols: output buffer too small");
          return DNAB_EINVAL;
        }
        if (bytes && !p.bytes.empty()) std::memcpy(bytes, p.bytes.data(), p.bytes.size());
        if (leftover_bits) std::memcpy(leftover_bits, p.leftoverBits.c_str(), p.leftoverBits.size() + 1);
        if (warnings) *warnings = dupString(joinLines(p.warnings));
        return (int64_t)p.bytes.size();
      },
      (int64_t)DNAB_EINVAL);
}

int dnab_exact_decode_fasta(const dnab_machine* m, const char* fasta_path, uint8_t** bytes, size_t* n_bytes,
                            char** warnings) {
  if (!m || !fasta_path || !bytes || !n_bytes) return DNAB_EINVAL;
  *bytes = nullptr;
  *n_bytes = 0;
  if (warnings) *warnings = nullptr;
  try {
    std::vector<FastSeq> seqs;
    try {
      seqs = readFastSeqs(fasta_path);
    } catch (const std::exception& e) {
      setLastError(e.what());
      return DNAB_EIO;
    }
    ExactDecoder dec(m->m);  // ONE decoder for every record: its state carries over (t/dnastore.cpp:187-190)
    for (const auto& fs : seqs) dec.decodeString(fs.seq);
    dec.close();
    const PackedBits p = packDecodedSymbols(dec.symbols());
    *bytes = (uint8_t*)std::malloc(p.bytes.size() + 1);
    if (!*bytes) throw std::bad_alloc();
    std::memcpy(*bytes, p.bytes.data(), p.bytes.size());
    *n_bytes = p.bytes.size();
    if (warnings) *warnings = dupString(joinLines(dec.warnings()) + joinLines(p.warnings));
    return DNAB_OK;
  } catch (const std::exception& e) {
    setLastError(e.what());
    std::free(*bytes);
    *bytes = nullptr;
    return DNAB_EINVAL;
  }
}

}  // extern "C"
