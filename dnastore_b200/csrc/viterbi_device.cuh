// Device-side table layout shared by the kernels (viterbi_kernels.cu) and the
// host code that builds it (decoder.cu).
#pragma once
#include <cstdint>

namespace dnab {

constexpr int kMaxSyms = 32;     // distinct input symbols incl. the "no input" symbol 0
constexpr int kMaxK = 6;         // duplication depth supported by the packed state record (-l <= 13)
constexpr int kMaxCluster = 16;
constexpr uint32_t kNoPred = 255;  // predecessor record: "no candidate"

// Packed incoming edge: source = (rank, local index) in the cluster partition.
//   bits  0..19  local index of the source state inside its CTA's slice
//   bits 20..23  cluster rank of the CTA owning the source
//   bits 24..28  input-symbol id (0 = no input symbol, score 0)
//   bits 29..30  emitted base (emit edges only)
__host__ __device__ inline uint32_t edgeLocal(uint32_t w) { return w & 0xFFFFFu; }
__host__ __device__ inline uint32_t edgeRank(uint32_t w) { return (w >> 20) & 0xFu; }
__host__ __device__ inline uint32_t edgeSym(uint32_t w) { return (w >> 24) & 0x1Fu; }
__host__ __device__ inline uint32_t edgeBase(uint32_t w) { return (w >> 29) & 0x3u; }
__host__ __device__ inline uint32_t packEdge(uint32_t local, uint32_t rank, uint32_t sym, uint32_t base) {
  return local | (rank << 20) | (sym << 24) | (base << 29);
}

// Per-state record (uint2):
//   .x  offset of the state's incoming edges in inEdges: [emit edges][null edges],
//       each group in the reference's list order (source index, transition index)
//   .y  bits 0..7 nEmit, 8..15 nNull, 16..19 mdl, 20..31 ctx (2 bits per
//       duplication index i: tanDupBase(ss,i))
__host__ __device__ inline uint32_t recNEmit(uint32_t y) { return y & 0xFFu; }
__host__ __device__ inline uint32_t recNNull(uint32_t y) { return (y >> 8) & 0xFFu; }
__host__ __device__ inline uint32_t recMdl(uint32_t y) { return (y >> 16) & 0xFu; }
__host__ __device__ inline uint32_t recCtx(uint32_t y, uint32_t i) { return (y >> (20 + 2 * i)) & 0x3u; }

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
  const uint2* stateRec;      // [Np]
  const uint32_t* inEdges;    // packed incoming edges
  const uint32_t* outOff;     // [Np+1] outgoing (emit+null) adjacency, for dirty marking
  const uint32_t* outEdges;   // local | rank<<20
  const uint32_t* origId;     // [Np] reference state index, 0xFFFFFFFF for padding
  const uint8_t* symChar;     // [nSyms] input-symbol character of each id
  double symScore[kMaxSyms];  // log(symProb) per id (0 for id 0)
  double sub[16];
  double len[kMaxK > 0 ? kMaxK : 1];
  double noGap, delOpen, delExtend, delEnd, tanDup;
};

}  // namespace dnab
