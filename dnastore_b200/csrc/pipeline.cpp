// Ingest/egress pipeline and multi-GPU sharding of the Viterbi decode.
//
// The reference's seam is one loop over independent reads (reference src/viterbi.cpp:312-318, fed by readFastSeqs,
// src/fastseq.cpp:123-148).  Here that loop becomes
//   producer thread   parses the FASTA/FASTQ(.gz) file in bounded chunks and packs the bases 2 bits each into
//                     page-locked host memory (non-ACGT characters are an error, as in the reference)
//   bounded queue     two chunks per device in flight: parsing chunk i+1 overlaps copying and decoding chunk i
//   one worker thread per device (its own dnab_decoder): H2D, fill + traceback, D2H; only the reads whose decoded
//                     string overflowed the chunk's slot are decoded again with a larger one
//   ordered gather    results are stored by chunk index, so the output keeps the input order whatever device took
//                     which chunk; chunks are handed out dynamically (work balance follows sum(L+1) by itself)
// No collective: reads are independent.  dnab_viterbi_batch_multi is the same worker pool over caller-provided
// packed arrays (the `e2e` path of bench.py --gpus N).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dnastore_b200.h"
#include "capi_error.h"
#include "capi_types.h"
#include "host/fasta.h"

using namespace dnab;

struct dnab_multi_decoder {
  std::vector<int> devices;
  std::vector<dnab_decoder*> dec;
  bool owns = true;
  int64_t chunkReads = 0;  // 0 = automatic
  int64_t slotBytes = 0;   // decoded-string slot per read of the first attempt, 0 = 2*maxLen+256
  dnab_pipeline_stats stats{};
};

namespace {

double nowSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// page-locked host buffer (plain malloc when no CUDA context can pin it: slower copies, same results)
struct HostBuf {
  void* p = nullptr;
  size_t n = 0;
  bool pinned = false;
  ~HostBuf() { release(); }
  void release() {
    if (!p) return;
    if (pinned)
      cudaFreeHost(p);
    else
      std::free(p);
    p = nullptr;
    n = 0;
  }
  void ensure(size_t bytes) {
    if (bytes <= n) return;
    release();
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess)
      pinned = true;
    else {
      cudaGetLastError();
      p = std::malloc(bytes);
      pinned = false;
      if (!p) throw std::bad_alloc();
    }
    n = bytes;
  }
};

struct Chunk {
  int64_t index = 0, n = 0;
  std::vector<std::string> names;
  HostBuf packed;
  std::vector<int64_t> byteOff;
  std::vector<int32_t> len;
  int32_t maxLen = 0;
  // results
  std::vector<double> loglike;
  std::vector<int32_t> status;
  std::vector<std::string> decoded;
};

// One chunk through one decoder: decoded strings land in a per-chunk slot of 2*maxLen+256 bytes per read; the
// reads that overflow it (status DNAB_READ_OVERFLOW) -- and only those -- are decoded again with 4x the slot.
void decodeChunk(dnab_decoder* d, int64_t n, const uint8_t* packed, const int64_t* byteOff, const int32_t* len, int32_t maxLen,
                 HostBuf& decBuf, double* loglike, int32_t* status, std::vector<std::string>& decoded, int64_t* reruns,
                 int64_t slotBytes) {
  int64_t stride = slotBytes > 0 ? slotBytes : 2 * (int64_t)maxLen + 256;
  std::vector<int32_t> decLen((size_t)n);
  decBuf.ensure((size_t)n * (size_t)stride);
  int rc = dnab_viterbi_batch(d, n, packed, byteOff, len, loglike, (char*)decBuf.p, (int32_t)stride, decLen.data(), status, nullptr,
                              0, nullptr);
  if (rc != DNAB_OK) throw std::runtime_error(dnab_last_error());
  decoded.assign((size_t)n, std::string());
  std::vector<int64_t> redo;
  for (int64_t r = 0; r < n; ++r) {
    if (status[r] == DNAB_READ_OVERFLOW)
      redo.push_back(r);
    else
      decoded[(size_t)r].assign((const char*)decBuf.p + (size_t)r * (size_t)stride, (size_t)decLen[(size_t)r]);
  }
  for (int attempt = 0; attempt < 6 && !redo.empty(); ++attempt) {
    stride *= 4;
    if (stride > ((int64_t)1 << 30)) throw std::runtime_error("decoded string exceeds 1 GiB");
    const int64_t m = (int64_t)redo.size();
    if (reruns) *reruns += m;
    std::vector<int64_t> off2((size_t)m);
    std::vector<int32_t> len2((size_t)m), decLen2((size_t)m), st2((size_t)m);
    std::vector<double> ll2((size_t)m);
    for (int64_t i = 0; i < m; ++i) {
      off2[(size_t)i] = byteOff[redo[(size_t)i]];
      len2[(size_t)i] = len[redo[(size_t)i]];
    }
    decBuf.ensure((size_t)m * (size_t)stride);
    rc = dnab_viterbi_batch(d, m, packed, off2.data(), len2.data(), ll2.data(), (char*)decBuf.p, (int32_t)stride, decLen2.data(),
                            st2.data(), nullptr, 0, nullptr);
    if (rc != DNAB_OK) throw std::runtime_error(dnab_last_error());
    std::vector<int64_t> still;
    for (int64_t i = 0; i < m; ++i) {
      const int64_t r = redo[(size_t)i];
      status[r] = st2[(size_t)i];
      loglike[r] = ll2[(size_t)i];
      if (st2[(size_t)i] == DNAB_READ_OVERFLOW)
        still.push_back(r);
      else
        decoded[(size_t)r].assign((const char*)decBuf.p + (size_t)i * (size_t)stride, (size_t)decLen2[(size_t)i]);
    }
    redo.swap(still);
  }
}

int64_t autoChunk(const dnab_multi_decoder* m, int64_t nReads) {
  if (m->chunkReads > 0) return m->chunkReads;
  const int64_t nDev = (int64_t)m->dec.size();
  int64_t c = (nReads + 4 * nDev - 1) / (4 * nDev);  // about four chunks per device: dynamic balance, few launches
  c = std::max<int64_t>(32, (c + 31) / 32 * 32);
  return std::min<int64_t>(c, 1 << 16);
}

}  // namespace

extern "C" {

dnab_multi_decoder* dnab_multi_decoder_create(const dnab_tables* t, const int* devices, int n_devices) {
  if (!t || !devices || n_devices < 1) {
    setLastError("dnab_multi_decoder_create: bad argument");
    return nullptr;
  }
  std::unique_ptr<dnab_multi_decoder> m(new dnab_multi_decoder());
  for (int i = 0; i < n_devices; ++i) {
    dnab_decoder* d = dnab_decoder_create(t, devices[i]);
    if (!d) {
      for (dnab_decoder* x : m->dec) dnab_decoder_destroy(x);
      return nullptr;  // dnab_last_error() says why (no CPU fallback)
    }
    m->devices.push_back(devices[i]);
    m->dec.push_back(d);
  }
  return m.release();
}

void dnab_multi_decoder_destroy(dnab_multi_decoder* m) {
  if (!m) return;
  if (m->owns)
    for (dnab_decoder* d : m->dec) dnab_decoder_destroy(d);
  delete m;
}

int dnab_multi_decoder_count(const dnab_multi_decoder* m) { return m ? (int)m->dec.size() : 0; }
dnab_decoder* dnab_multi_decoder_at(dnab_multi_decoder* m, int i) {
  return (m && i >= 0 && i < (int)m->dec.size()) ? m->dec[(size_t)i] : nullptr;
}
int dnab_multi_decoder_set_option(dnab_multi_decoder* m, const char* key, int64_t value) {
  if (!m || !key || value < 0) {
    setLastError("dnab_multi_decoder_set_option: bad argument");
    return DNAB_EINVAL;
  }
  const std::string k(key);
  if (k == "chunk_reads")
    m->chunkReads = value;
  else if (k == "decoded_slot_bytes")
    m->slotBytes = value;
  else {  // everything else is a decoder option, applied to every device
    for (dnab_decoder* d : m->dec) {
      const int rc = dnab_decoder_set_option(d, key, value);
      if (rc != DNAB_OK) return rc;
    }
  }
  return DNAB_OK;
}
int dnab_multi_decoder_last_stats(const dnab_multi_decoder* m, dnab_pipeline_stats* s) {
  if (!m || !s) return DNAB_EINVAL;
  *s = m->stats;
  return DNAB_OK;
}

int dnab_viterbi_batch_multi(dnab_multi_decoder* m, int64_t n_reads, const uint8_t* packed, const int64_t* read_byte_off,
                             const int32_t* read_len, double* loglike, char* decoded, int32_t decoded_stride,
                             int32_t* decoded_len, int32_t* status) {
  if (!m || n_reads < 0 || decoded_stride <= 0) {
    setLastError("dnab_viterbi_batch_multi: bad argument");
    return DNAB_EINVAL;
  }
  if (n_reads == 0) return DNAB_OK;
  const double t0 = nowSeconds();
  const int64_t chunk = autoChunk(m, n_reads), nChunks = (n_reads + chunk - 1) / chunk;
  std::atomic<int64_t> next(0);
  std::mutex errMutex;
  std::string err;
  std::atomic<bool> failed(false);
  std::vector<double> busy(m->dec.size(), 0.);
  auto worker = [&](size_t w) {
    std::vector<int64_t> off;
    for (;;) {
      const int64_t c = next.fetch_add(1);
      if (c >= nChunks || failed.load()) return;
      const int64_t a = c * chunk, b = std::min(n_reads, a + chunk), n = b - a;
      const int64_t base = read_byte_off[a];  // shard-relative offsets: a worker copies only its own bases to its device
      off.resize((size_t)n);
      int64_t lo = base;
      for (int64_t r = 0; r < n; ++r) lo = std::min(lo, read_byte_off[a + r]);
      for (int64_t r = 0; r < n; ++r) off[(size_t)r] = read_byte_off[a + r] - lo;
      const double t = nowSeconds();
      const int rc = dnab_viterbi_batch(m->dec[w], n, packed + lo, off.data(), read_len + a, loglike + a,
                                        decoded + (size_t)a * (size_t)decoded_stride, decoded_stride, decoded_len + a, status + a,
                                        nullptr, 0, nullptr);
      busy[w] += nowSeconds() - t;
      if (rc != DNAB_OK) {
        std::lock_guard<std::mutex> g(errMutex);
        if (!failed.exchange(true)) err = dnab_last_error();
        return;
      }
    }
  };
  std::vector<std::thread> threads;
  for (size_t w = 1; w < m->dec.size(); ++w) threads.emplace_back(worker, w);
  worker(0);
  for (auto& t : threads) t.join();
  m->stats = dnab_pipeline_stats{};
  m->stats.reads = n_reads;
  m->stats.chunks = nChunks;
  m->stats.wall_seconds = nowSeconds() - t0;
  for (double b : busy) m->stats.decode_busy_seconds += b;
  if (failed.load()) {
    setLastError(err);
    return DNAB_ECUDA;
  }
  return DNAB_OK;
}

dnab_decoded_set* dnab_decode_fasta_multi(dnab_multi_decoder* m, const char* fasta_path) {
  if (!m || !fasta_path || m->dec.empty()) {
    setLastError("dnab_decode_fasta_multi: bad argument");
    return nullptr;
  }
  const double t0 = nowSeconds();
  const size_t nDev = m->dec.size();
  const int64_t chunkReads = m->chunkReads > 0 ? m->chunkReads : 16384;
  const size_t chunkBases = (size_t)64 << 20;  // a chunk never holds more than 64 M bases, however long the reads are
  std::mutex mu;
  std::condition_variable cvFull, cvEmpty;
  std::deque<std::unique_ptr<Chunk>> queue;
  bool producerDone = false;
  std::atomic<bool> failed(false);
  std::string err;
  std::vector<std::unique_ptr<Chunk>> done;
  double parseSeconds = 0;
  std::vector<double> busy(nDev, 0.);
  std::vector<int64_t> reruns(nDev, 0);
  auto fail = [&](const std::string& msg) {
    std::lock_guard<std::mutex> g(mu);
    if (!failed.exchange(true)) err = msg;
    cvFull.notify_all();
    cvEmpty.notify_all();
  };

  std::thread producer([&]() {
    try {
      FastSeqStream in(fasta_path);
      std::vector<FastSeq> recs;
      int64_t index = 0, readsSoFar = 0;
      for (;;) {
        const double t = nowSeconds();
        recs.clear();
        if (!in.next((size_t)chunkReads, chunkBases, recs)) break;
        std::unique_ptr<Chunk> c(new Chunk());
        c->index = index++;
        c->n = (int64_t)recs.size();
        c->len.resize(recs.size());
        c->byteOff.resize(recs.size());
        std::string bases;
        std::vector<int64_t> baseOff(recs.size() + 1, 0);
        for (size_t r = 0; r < recs.size(); ++r) {
          c->names.push_back(recs[r].name);
          bases += recs[r].seq;
          baseOff[r + 1] = (int64_t)bases.size();
          if (recs[r].seq.size() > (size_t)0x7FFFFF00) throw std::runtime_error("read " + recs[r].name + " is too long");
          c->len[r] = (int32_t)recs[r].seq.size();
          c->maxLen = std::max(c->maxLen, c->len[r]);
        }
        c->packed.ensure(packedSize(c->len.data(), c->n));
        char bad = 0;
        const int64_t badRead = packReads(bases.data(), baseOff.data(), c->n, (uint8_t*)c->packed.p, c->byteOff.data(), c->len.data(), &bad);
        if (badRead >= 0)
          throw std::runtime_error(std::string("Unknown symbol ") + bad + " in sequence " + recs[(size_t)badRead].name +
                                   " (alphabet is ACGT)");
        readsSoFar += c->n;
        parseSeconds += nowSeconds() - t;
        std::unique_lock<std::mutex> lk(mu);
        cvFull.wait(lk, [&]() { return queue.size() < 2 * nDev || failed.load(); });
        if (failed.load()) return;
        queue.push_back(std::move(c));
        cvEmpty.notify_one();
      }
    } catch (const std::exception& e) {
      fail(e.what());
    }
    std::lock_guard<std::mutex> g(mu);
    producerDone = true;
    cvEmpty.notify_all();
  });

  auto worker = [&](size_t w) {
    HostBuf decBuf;
    try {
      for (;;) {
        std::unique_ptr<Chunk> c;
        {
          std::unique_lock<std::mutex> lk(mu);
          cvEmpty.wait(lk, [&]() { return !queue.empty() || producerDone || failed.load(); });
          if (failed.load()) return;
          if (queue.empty()) return;  // producer finished
          c = std::move(queue.front());
          queue.pop_front();
          cvFull.notify_one();
        }
        const double t = nowSeconds();
        c->loglike.resize((size_t)c->n);
        c->status.resize((size_t)c->n);
        decodeChunk(m->dec[w], c->n, (const uint8_t*)c->packed.p, c->byteOff.data(), c->len.data(), c->maxLen, decBuf,
                    c->loglike.data(), c->status.data(), c->decoded, &reruns[w], m->slotBytes);
        c->packed.release();
        busy[w] += nowSeconds() - t;
        std::lock_guard<std::mutex> g(mu);
        done.push_back(std::move(c));
      }
    } catch (const std::exception& e) {
      fail(e.what());
    }
  };
  std::vector<std::thread> workers;
  for (size_t w = 0; w < nDev; ++w) workers.emplace_back(worker, w);
  producer.join();
  for (auto& t : workers) t.join();
  if (failed.load()) {
    setLastError(err);
    return nullptr;
  }
  std::sort(done.begin(), done.end(), [](const std::unique_ptr<Chunk>& a, const std::unique_ptr<Chunk>& b) { return a->index < b->index; });
  std::unique_ptr<dnab_decoded_set> out(new dnab_decoded_set());
  for (auto& c : done)
    for (int64_t r = 0; r < c->n; ++r) {
      out->names.push_back(std::move(c->names[(size_t)r]));
      out->seqs.push_back(std::move(c->decoded[(size_t)r]));
      out->loglike.push_back(c->loglike[(size_t)r]);
      out->status.push_back(c->status[(size_t)r]);
    }
  m->stats = dnab_pipeline_stats{};
  m->stats.reads = (int64_t)out->seqs.size();
  m->stats.chunks = (int64_t)done.size();
  m->stats.parse_seconds = parseSeconds;
  m->stats.wall_seconds = nowSeconds() - t0;
  for (double b : busy) m->stats.decode_busy_seconds += b;
  for (int64_t r : reruns) m->stats.overflow_reruns += r;
  return out.release();
}

dnab_decoded_set* dnab_decode_fasta(dnab_decoder* d, const char* fasta_path) {
  if (!d || !fasta_path) {
    setLastError("dnab_decode_fasta: null argument");
    return nullptr;
  }
  dnab_multi_decoder one;
  one.owns = false;
  one.dec.push_back(d);
  one.devices.push_back(-1);
  return dnab_decode_fasta_multi(&one, fasta_path);
}

}  // extern "C"
