// Micro-probe (not part of the product): what does a release fence cost on this GPU, alone and under load, and what is the
// latency of one cross-CTA hop (publish a 512-byte row, order it, notify; the peer polls, reads the row, answers)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fence_probe fence_probe.cu && ./fence_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fenceRelease() { asm volatile("fence.release.gpu;" ::: "memory"); }
__device__ __forceinline__ void fenceAcqRel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void stPub2(double2* p, double a, double b) {
  asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double2 ldPub2(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void redAdd32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void redRelAdd32(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldVol(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ldAcq(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// mode 0: store row + fence.release;  1: fence only (nothing outstanding);  2: row + red + fence;  3: row + fence while the
// other warps of the CTA stream stores without fencing
__global__ void fenceCost(double2* rows, uint32_t* ctr, unsigned long long* out, int activeWarps, int mode, int iters) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2* mine = rows + ((size_t)blockIdx.x * 32 + warp) * 64 * 32 + lane;
  unsigned long long sum = 0;
  if ((int)warp < activeWarps) {
    for (int i = 0; i < iters; ++i) {
      if (mode != 1) stPub2(mine + (i & 63) * 32, (double)i, 1.0);
      if (mode == 2 && lane == 0) redAdd32(ctr + blockIdx.x, 1u);
      const unsigned long long t0 = clock64();
      fenceRelease();
      sum += clock64() - t0;
      __syncwarp();
    }
    if (lane == 0) atomicAdd(out, sum);
  } else if (mode == 3) {
    for (int i = 0; i < iters * 4; ++i) stPub2(mine + (i & 63) * 32, (double)i, 1.0);
  }
}

// Ping-pong between CTA 2p and CTA 2p+1 (one warp each); variant 0: row, fence.release, red;  1: row, red.release (no separate
// fence);  2: row, red with NO ordering (incorrect -- lower bound);  3: as 0 but the peer polls with ld.acquire
__global__ void pingPong(double2* rows, uint32_t* flags, unsigned long long* out, int variant, int iters, int otherWarpsSpin) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t me = blockIdx.x, peer = me ^ 1u;
  if (warp != 0) {
    if (otherWarpsSpin) {  // background: idle warps polling shared memory-free global word like the kernel's idle loop
      volatile uint32_t* f = flags + 8192 + me;
      while (*f == 0) __nanosleep(100);
    }
    return;
  }
  double2* myRow = rows + (size_t)me * 32 + lane;
  const double2* peerRow = rows + (size_t)peer * 32 + lane;
  uint32_t* myFlag = flags + me * 32;      // 128 bytes apart
  uint32_t* peerFlag = flags + peer * 32;
  uint32_t seen = 0;
  double acc = 0;
  const unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if ((me & 1u) == 0 || i > 0 || true) {
      if ((me & 1u) == 1u || i > 0) {  // wait for the peer's notification (CTA 2p starts)
        uint32_t c;
        do { c = variant == 3 ? ldAcq(myFlag) : ldVol(myFlag); } while (c == seen);
        seen = c;
        const double2 v = ldPub2(peerRow);
        acc += v.x;
      }
      stPub2(myRow, acc + 1.0, (double)i);
      if (variant == 0 || variant == 3) fenceRelease();
      __syncwarp();
      if (lane == 0) {
        if (variant == 1) redRelAdd32(peerFlag, 1u);
        else redAdd32(peerFlag, 1u);
      }
    }
  }
  // drain: the odd CTA sent `iters` notes, the even one too; wait for the last one so that nobody exits early
  if ((me & 1u) == 0) {
    uint32_t c;
    do { c = ldVol(myFlag); } while (c == seen);
  }
  const unsigned long long t1 = clock64();
  if (lane == 0 && me == 0) out[0] = t1 - t0;
  if (lane == 0) flags[8192 + me] = 1;
  if (acc == -1.0) out[1] = 1;
}

int main() {
  double2* rows;
  uint32_t* ctr;
  unsigned long long* out;
  cudaMalloc(&rows, (size_t)148 * 32 * 64 * 32 * sizeof(double2));
  cudaMalloc(&ctr, 1 << 20);
  cudaMalloc(&out, 64);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int aw : {1, 8, 32}) {
      if (mode == 3 && aw == 32) continue;
      cudaMemset(out, 0, 64);
      fenceCost<<<148, 1024>>>(rows, ctr, out, aw, mode, iters);
      unsigned long long h = 0;
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("fenceCost mode %d (0 row+fence, 1 fence only, 2 row+red+fence, 3 row+fence, other warps streaming) active warps/CTA %2d: %.0f cycles per fence\n",
             mode, aw, (double)h / (148.0 * aw * iters));
    }
  for (int pairs : {1, 74})
    for (int spin : {0, 1})
      for (int variant = 0; variant < 4; ++variant) {
        cudaMemset(out, 0, 64);
        cudaMemset(ctr, 0, 1 << 20);
        pingPong<<<pairs * 2, spin ? 1024 : 32>>>(rows, ctr, out, variant, iters, spin);
        unsigned long long h = 0;
        cudaError_t e = cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("pingPong pairs %2d idle-warps %d variant %d (0 fence+red, 1 red.release, 2 unordered, 3 fence+red / ld.acquire poll): %.0f cycles per hop  %s\n",
               pairs, spin, variant, (double)h / (2.0 * iters), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
