#include "fasta.h"

#include <zlib.h>

#include <cctype>
#include <cstring>
#include <ostream>
#include <stdexcept>

namespace dnab {

namespace {

// Line reader over zlib's gzFile, which also reads uncompressed files transparently.
class GzLines {
 public:
  explicit GzLines(const std::string& filename) : fp(gzopen(filename.c_str(), "r")) {
    if (!fp) throw std::runtime_error("Couldn't open " + filename);
    gzbuffer(fp, 1 << 20);
  }
  ~GzLines() { gzclose(fp); }
  bool next(std::string& line) {
    line.clear();
    char buf[1 << 16];
    bool got = false;
    while (gzgets(fp, buf, sizeof buf)) {
      got = true;
      const size_t n = std::strlen(buf);
      line.append(buf, n);
      if (n && buf[n - 1] == '\n') break;
    }
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    return got;
  }

 private:
  gzFile fp;
};

void splitHeader(const std::string& line, FastSeq& fs) {
  size_t p = 1;
  while (p < line.size() && !std::isspace((unsigned char)line[p])) ++p;
  fs.name = line.substr(1, p - 1);
  while (p < line.size() && std::isspace((unsigned char)line[p])) ++p;
  fs.comment = line.substr(p);
}

void appendNonSpace(std::string& dst, const std::string& line) {
  for (char c : line)
    if (!std::isspace((unsigned char)c)) dst.push_back(c);
}

}  // namespace

struct FastSeqStream::Impl {
  GzLines in;
  std::string line;
  bool have;
  explicit Impl(const std::string& filename) : in(filename) { have = in.next(line); }
};

FastSeqStream::FastSeqStream(const std::string& filename) : impl(new Impl(filename)) {}
FastSeqStream::~FastSeqStream() { delete impl; }

bool FastSeqStream::next(size_t maxReads, size_t maxBases, std::vector<FastSeq>& out) {
  GzLines& in = impl->in;
  std::string& line = impl->line;
  bool& have = impl->have;
  size_t nReads = 0, nBases = 0;
  while (have && nReads < maxReads && nBases < maxBases) {
    if (line.empty() || (line[0] != '>' && line[0] != '@')) {  // junk before a header
      have = in.next(line);
      continue;
    }
    FastSeq fs;
    splitHeader(line, fs);
    bool sawPlus = false;
    while ((have = in.next(line))) {
      if (!line.empty() && (line[0] == '>' || line[0] == '@')) break;
      if (!line.empty() && line[0] == '+') {
        sawPlus = true;
        break;
      }
      appendNonSpace(fs.seq, line);
    }
    if (sawPlus) {  // FASTQ quality block: as many characters as bases
      while (fs.qual.size() < fs.seq.size() && (have = in.next(line))) appendNonSpace(fs.qual, line);
      have = in.next(line);
    }
    ++nReads;
    nBases += fs.seq.size();
    out.push_back(std::move(fs));
  }
  return nReads > 0;
}

std::vector<FastSeq> readFastSeqs(const std::string& filename) {
  FastSeqStream in(filename);
  std::vector<FastSeq> seqs;
  while (in.next(1u << 20, (size_t)1 << 30, seqs)) {
  }
  return seqs;
}

void writeFastaSeqs(std::ostream& out, const std::vector<FastSeq>& seqs, size_t lineWidth) {
  for (const auto& fs : seqs) {
    out << '>' << fs.name;
    if (!fs.comment.empty()) out << ' ' << fs.comment;
    out << '\n';
    for (size_t i = 0; i < fs.seq.size(); i += lineWidth) out << fs.seq.substr(i, lineWidth) << '\n';
  }
}

int baseToken(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

size_t packedSize(const int32_t* readLen, int64_t nReads) {
  size_t total = 0;
  for (int64_t r = 0; r < nReads; ++r) total += (((size_t)readLen[r] + 3) / 4 + 15) & ~(size_t)15;
  return total ? total : 16;
}

int64_t packReads(const char* bases, const int64_t* baseOff, int64_t nReads, uint8_t* packed, int64_t* byteOff,
                  int32_t* readLen, char* badChar) {
  int64_t at = 0;
  for (int64_t r = 0; r < nReads; ++r) {
    const int64_t len = baseOff[r + 1] - baseOff[r];
    const int64_t nBytes = ((len + 3) / 4 + 15) & ~(int64_t)15;
    byteOff[r] = at;
    readLen[r] = (int32_t)len;
    std::memset(packed + at, 0, (size_t)nBytes);
    for (int64_t i = 0; i < len; ++i) {
      const int tok = baseToken(bases[baseOff[r] + i]);
      if (tok < 0) {
        if (badChar) *badChar = bases[baseOff[r] + i];
        return r;
      }
      packed[at + i / 4] |= (uint8_t)(tok << (2 * (i % 4)));
    }
    at += nBytes;
  }
  return -1;
}

}  // namespace dnab
