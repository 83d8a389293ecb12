// Launch interface of the CUDA kernels (viterbi_kernels.cu) used by decoder.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "viterbi_device.cuh"

namespace dnab {

// mirror include/dnastore_b200.h DNAB_READ_* without pulling the C header into device code
constexpr int32_t DNAB_READ_OK_ = 0;
constexpr int32_t DNAB_READ_NO_DECODING_ = 1;
constexpr int32_t DNAB_READ_OVERFLOW_ = 2;
constexpr int32_t DNAB_READ_TRACEBACK_FAILED_ = 3;

struct FillArgs {
  int64_t nReads;
  int32_t maxLen;            // pred stride: every read owns (maxLen+1) columns of records
  const uint8_t* packed;     // 2-bit reads
  const int64_t* byteOff;    // [nReads]
  const int32_t* readLen;    // [nReads]
  uint8_t* pred;             // [nReads][maxLen+1][k+2][Np] predecessor records, 1 byte per DP cell
  double* tScratch;          // [nClusters*C][k][M] T columns when they do not fit in shared memory
  double* loglike;           // [nReads] (global mode; local mode: written by the traceback kernel)
  uint32_t* startState;      // [nReads] traceback start (padded index), global mode
  double* partVal;           // [nReads][C] local mode: per-CTA best final S ...
  uint32_t* partOrig;        //             ... its reference state index (tie-break) ...
  uint32_t* partG;           //             ... and padded index
  double* cells;             // optional debug dump of read 0: [(L+1)][nStates][k+2]
};

struct TracebackArgs {
  int64_t nReads;
  int32_t maxLen;
  const int32_t* readLen;
  const uint8_t* pred;
  double* loglike;
  const uint32_t* startState;
  const double* partVal;
  const uint32_t* partOrig;
  const uint32_t* partG;
  char* decoded;             // [nReads][decodedStride]
  int32_t decodedStride;
  int32_t* decodedLen;
  int32_t* status;
  int32_t* path;             // optional [nReads][pathStride][3]
  int32_t pathStride;
  int32_t* pathLen;
};

uint32_t fillSmemBytes(uint32_t M, uint32_t k, uint32_t tInSmem, uint32_t maxLen);
cudaError_t queryMaxClusters(const DevTables& tb, uint32_t threads, uint32_t smemBytes, int* nClusters);
cudaError_t launchFill(const DevTables& tb, const FillArgs& args, uint32_t nClusters, uint32_t threads,
                       uint32_t smemBytes, cudaStream_t stream);
cudaError_t launchTraceback(const DevTables& tb, const TracebackArgs& args, cudaStream_t stream);

}  // namespace dnab
