// Peer exchange over NVLink for the partitioned (multi-GPU) solve.
//
// What couples the ranks of a partitioned network is additive and tiny (<= 3 n_top doubles: the top
// chunk's partial pivots / link conductances / right-hand side, and the shared multiplier rows of a
// residual with two norm partials).  Instead of returning to the host for an NCCL all-reduce, the
// kernel that produces the partial sums stores them straight into EVERY rank's exchange buffer
// (cudaIpc-mapped peer memory; the stores travel over NVLink / NVSwitch), raises a flag there, waits
// for the flags of all ranks in its own buffer and adds the contributions up in rank order -- the
// same order on every rank, so all ranks continue with bit-identical values (replicated top chunk,
// identical norms, identical refinement decisions) and the whole solve keeps the single-GPU launch
// sequence with no host round trip.
//
// Buffer of one rank (cudaMalloc, exported with cudaIpcGetMemHandle):
//   [0, 256)   unsigned flags[kPeerChannels][kMaxPeers]   flags[ch][src] = last epoch src has delivered
//   [256, ..)  double data[kPeerChannels][2][nranks][slot]  parity = epoch & 1
// Two parities suffice: a rank can only deliver epoch e+2 after it has consumed everybody's e+1,
// which everybody sends only after consuming epoch e.
#pragma once

#include <cstdint>

namespace nxfx {

constexpr int kMaxPeers = 16;
constexpr int kPeerChannels = 2;  // 0: top chunk of the tree solve, 1: residual
constexpr size_t kPeerHeaderBytes = 256;

struct PeerDev {
  unsigned long long base[kMaxPeers];  // exchange buffer of every rank as seen from this device
  int rank, nranks;                    // nranks <= 1: single GPU, no exchange
  int slot;                            // doubles per (channel, parity, source)
  unsigned int epoch;                  // of this use of the channel (starts at 1)
  int* err;                            // mapped host word: set to 1 when a peer did not arrive in time
};

__device__ __forceinline__ double* peer_data(const PeerDev& c, int dst, int ch, int src) {
  char* b = reinterpret_cast<char*>(c.base[dst]) + kPeerHeaderBytes;
  return reinterpret_cast<double*>(b) + ((size_t)((ch * 2 + (int)(c.epoch & 1u)) * c.nranks + src)) * (size_t)c.slot;
}
__device__ __forceinline__ unsigned int* peer_flag(const PeerDev& c, int dst, int ch, int src) {
  return reinterpret_cast<unsigned int*>(c.base[dst]) + ch * kMaxPeers + src;
}

// All threads of ONE block: deliver n doubles get(i) to every rank (slot of this rank), then wait until
// every rank has delivered its epoch.  Afterwards peer_data(c, c.rank, ch, src)[i] (read with __ldcg:
// the lines were written by remote stores) holds the contribution of rank src.
template <typename Get>
__device__ __forceinline__ void peer_allgather(const PeerDev& c, int ch, int n, Get get) {
  for (int dst = 0; dst < c.nranks; ++dst) {
    double* out = peer_data(c, dst, ch, c.rank);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = get(i);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < c.nranks) {
    unsigned int* f = peer_flag(c, (int)threadIdx.x, ch, c.rank);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(c.epoch) : "memory");
    const unsigned int* mine = peer_flag(c, c.rank, ch, (int)threadIdx.x);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - c.epoch) >= 0) break;
      __nanosleep(40);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) {  // 4 s: a peer is gone -- flag the error and let the kernel finish
        *reinterpret_cast<volatile int*>(c.err) = 1;
        break;
      }
    }
  }
  __syncthreads();
}

}  // namespace nxfx
