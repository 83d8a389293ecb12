// FASTA/FASTQ ingest and 2-bit packing for the decoder boundary.
// Behaviour follows the reference's readFastSeqs / FastSeq::tokens / writeFastaSeqs
// (reference src/fastseq.cpp:9-56,82-105,123-148): plain or gzip input, multi-line
// records concatenated, name = header up to the first whitespace, bases
// case-insensitive over "ACGT", any other character is an error.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

namespace dnab {

struct FastSeq {
  std::string name, comment, seq, qual;
};

std::vector<FastSeq> readFastSeqs(const std::string& filename);

// The same parser as readFastSeqs, one bounded chunk at a time (the ingest side of the decode pipeline): next()
// appends records until it holds maxReads of them or maxBases bases and returns false once the file is exhausted
// and nothing was appended.
class FastSeqStream {
 public:
  explicit FastSeqStream(const std::string& filename);
  ~FastSeqStream();
  FastSeqStream(const FastSeqStream&) = delete;
  FastSeqStream& operator=(const FastSeqStream&) = delete;
  bool next(size_t maxReads, size_t maxBases, std::vector<FastSeq>& out);

 private:
  struct Impl;
  Impl* impl;
};
void writeFastaSeqs(std::ostream& out, const std::vector<FastSeq>& seqs, size_t lineWidth = 50);

// A,C,G,T -> 0..3 (case-insensitive); -1 otherwise.
int baseToken(char c);

// Packed layout shared with the device: read r starts at byte byteOff[r] (16-byte
// aligned), base i sits in bits 2*(i%4) of byte i/4.
size_t packedSize(const int32_t* readLen, int64_t nReads);
// Returns -1 on success, else the index of the first read holding a non-ACGT character
// (badChar receives it).
int64_t packReads(const char* bases, const int64_t* baseOff, int64_t nReads, uint8_t* packed, int64_t* byteOff,
                  int32_t* readLen, char* badChar);

}  // namespace dnab
