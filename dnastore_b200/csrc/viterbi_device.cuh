// Device-side table layout shared by the kernels (viterbi_kernels.cu) and the
// host code that builds it (decoder.cu).
#pragma once
#include <cstdint>

namespace dnab {

constexpr int kMaxSyms = 32;     // distinct input symbols incl. the "no input" symbol 0
constexpr int kMaxK = 6;         // duplication depth supported by the packed state block (-l <= 13)
constexpr int kMaxCluster = 16;
constexpr uint32_t kNoPred = 255;  // predecessor record: "no candidate"
constexpr uint32_t kBlockWords = 8;  // state blocks are padded to 32 bytes (one L2 sector)

// Packed edge word: the OTHER endpoint = (rank, local index) in the cluster partition.
//   bits  0..19  local index of that state inside its CTA's slice
//   bits 20..23  cluster rank of the CTA owning it
//   bits 24..28  input-symbol id (0 = no input symbol, score 0)   [incoming edges]
//   bits 29..30  emitted base                                     [incoming emit edges]
__host__ __device__ inline uint32_t edgeLocal(uint32_t w) { return w & 0xFFFFFu; }
__host__ __device__ inline uint32_t edgeRank(uint32_t w) { return (w >> 20) & 0xFu; }
__host__ __device__ inline uint32_t edgeSym(uint32_t w) { return (w >> 24) & 0x1Fu; }
__host__ __device__ inline uint32_t edgeBase(uint32_t w) { return (w >> 29) & 0x3u; }
__host__ __device__ inline uint32_t packEdge(uint32_t local, uint32_t rank, uint32_t sym, uint32_t base) {
  return local | (rank << 20) | (sym << 24) | (base << 29);
}

// State block (32-byte aligned run of 32-bit words in `blocks`, found through blockOff):
//   word 0   nEmit | nNull<<8 | nOut<<16 | mdl<<24
//   word 1   ctx: 2 bits per duplication index i = tanDupBase(ss,i)
//   then     nEmit incoming emit edges, nNull incoming null edges -- each group in the
//            reference's list order (source index, transition index) --
//   then     nOut outgoing edges (destinations to wake when this state's cells grow)
// One sector fetch brings the header and the first six edge words.
__host__ __device__ inline uint32_t hdrNEmit(uint32_t w0) { return w0 & 0xFFu; }
__host__ __device__ inline uint32_t hdrNNull(uint32_t w0) { return (w0 >> 8) & 0xFFu; }
__host__ __device__ inline uint32_t hdrNOut(uint32_t w0) { return (w0 >> 16) & 0xFFu; }
__host__ __device__ inline uint32_t hdrMdl(uint32_t w0) { return (w0 >> 24) & 0xFu; }
__host__ __device__ inline uint32_t hdrCtx(uint32_t w1, uint32_t i) { return (w1 >> (2 * i)) & 0x3u; }

struct DevTables {
  uint32_t nStates;   // real states
  uint32_t M;         // states per CTA slice
  uint32_t C;         // cluster size; padded state count Np = C*M
  uint32_t k;         // duplication depth
  uint32_t local;     // local-alignment mode
  uint32_t nSyms;
  uint32_t startG;    // padded-space index of reference state 0
  uint32_t endG;      // padded-space index of the reference's last state
  uint32_t tInSmem;   // T columns in shared memory (else in global scratch)
  const uint32_t* blocks;     // state blocks
  const uint32_t* blockOff;   // [Np] word offset of each state's block (multiple of kBlockWords)
  const uint32_t* origId;     // [Np] reference state index, 0xFFFFFFFF for padding
  const uint8_t* symChar;     // [nSyms] input-symbol character of each id
  double symScore[kMaxSyms];  // log(symProb) per id (0 for id 0)
  double sub[16];
  double len[kMaxK > 0 ? kMaxK : 1];
  double noGap, delOpen, delExtend, delEnd, tanDup;
};

}  // namespace dnab
