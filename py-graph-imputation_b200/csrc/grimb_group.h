// grimb_group.h -- the thread-group abstraction the per-subject algorithm is written against.
//
// On the GPU a "group" is one CTA: tid = threadIdx.x, n = blockDim.x, barriers are
// __syncthreads(), scans/reductions use warp shuffles + shared memory.  When compiled with
// -DGRIMB_EMU by a host compiler (tests only, never shipped in libgrimb200.so) the group is a
// single thread, which lets the integer/FP64 logic of grimb_subject.h be exercised on a
// machine without a GPU.  The emulation build is a debugging aid for the test-suite; the
// product library contains only the CUDA path.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(GRIMB_EMU)
#define GRIMB_DEVICE 1
#define GD __device__ __forceinline__
#define GDN __device__ __noinline__
#else
#define GRIMB_DEVICE 0
#define GD inline
#define GDN inline
#endif

// Packed haplotype key: one bit-field per locus.  GRIMB_KW = 1: 64-bit keys (<= 63 bits in use;
// libgrimb200.so); GRIMB_KW = 2: 128-bit keys for wide 9-locus tables (libgrimb200w.so, same ABI
// with GRIMB_KEY_WORDS = 2).
#ifndef GRIMB_KW
#define GRIMB_KW 1
#endif

namespace grimb {

#if GRIMB_KW == 1
typedef uint64_t hkey;
#else
typedef unsigned __int128 hkey;
#endif

struct Grp {
  int tid;             // index of this thread in the group
  int n;               // threads in the group
  uint32_t* scratch;   // shared memory: >= 80 uint32_t

  GD void sync() const {
#if GRIMB_DEVICE
    __syncthreads();
#endif
  }

  // exclusive scan of one uint32 per thread; every thread of the group must call
  GD uint32_t scan_excl(uint32_t v, uint32_t& total) const {
#if GRIMB_DEVICE
    const int lane = tid & 31, w = tid >> 5, nw = (n + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31 || tid == n - 1) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
      uint32_t t = (lane < nw) ? scratch[lane] : 0u;
      uint32_t ti = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, ti, d);
        if (lane >= d) ti += o;
      }
      if (lane < nw) scratch[32 + lane] = ti - t;
      if (lane == 31) scratch[64] = ti;
    }
    __syncthreads();
    uint32_t res = inc - v + scratch[32 + w];
    total = scratch[64];
    __syncthreads();
    return res;
#else
    total = v;
    return 0;
#endif
  }

  GD uint32_t sum(uint32_t v) const {
    uint32_t t;
    scan_excl(v, t);
    return t;
  }

  GD bool any(bool p) const { return sum(p ? 1u : 0u) != 0; }

  GD double maxd(double v) const {
#if GRIMB_DEVICE
    const int lane = tid & 31, w = tid >> 5, nw = (n + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      double o = __shfl_xor_sync(0xffffffffu, v, d);
      v = o > v ? o : v;
    }
    double* ds = reinterpret_cast<double*>(scratch);
    if (lane == 0) ds[w] = v;
    __syncthreads();
    double r = ds[0];
    for (int i = 1; i < nw; ++i) r = ds[i] > r ? ds[i] : r;
    __syncthreads();
    return r;
#else
    return v;
#endif
  }

  GD uint64_t sum64(uint64_t v) const {
#if GRIMB_DEVICE
    const int lane = tid & 31, w = tid >> 5, nw = (n + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    uint64_t* ds = reinterpret_cast<uint64_t*>(scratch);
    if (lane == 0) ds[w] = v;
    __syncthreads();
    uint64_t r = 0;
    for (int i = 0; i < nw; ++i) r += ds[i];
    __syncthreads();
    return r;
#else
    return v;
#endif
  }
};

// ---- atomics (plain operations in the single-thread emulation) ----
GD uint32_t atom_add(uint32_t* p, uint32_t v) {
#if GRIMB_DEVICE
  return atomicAdd(p, v);
#else
  uint32_t o = *p; *p = o + v; return o;
#endif
}
GD unsigned long long atom_add64(unsigned long long* p, unsigned long long v) {
#if GRIMB_DEVICE
  return atomicAdd(p, v);
#else
  unsigned long long o = *p; *p = o + v; return o;
#endif
}
GD uint32_t atom_min(uint32_t* p, uint32_t v) {
#if GRIMB_DEVICE
  return atomicMin(p, v);
#else
  uint32_t o = *p; if (v < o) *p = v; return o;
#endif
}
GD uint32_t atom_cas(uint32_t* p, uint32_t expect, uint32_t v) {
#if GRIMB_DEVICE
  return atomicCAS(p, expect, v);
#else
  uint32_t o = *p; if (o == expect) *p = v; return o;
#endif
}
GD uint32_t atom_or(uint32_t* p, uint32_t v) {
#if GRIMB_DEVICE
  return atomicOr(p, v);
#else
  uint32_t o = *p; *p = o | v; return o;
#endif
}

GD uint64_t dbits(double x) {
#if GRIMB_DEVICE
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}

GD uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

// Hash of a packed key for the table regions: fold to 32 bits, then the murmur3 finaliser.  All
// 32-bit arithmetic (a 64-bit multiply is four instructions on the GPU); on the 1M-haplotype table
// its probe lengths equal those of a full 64-bit mixer (1.03 sectors per hit, 1.10 per miss at
// load 0.25).
GD uint32_t hash_key(hkey k) {
#if GRIMB_KW == 1
  uint32_t h = (uint32_t)k ^ ((uint32_t)(k >> 32) * 0x9E3779B1u);
#else
  uint32_t h = (uint32_t)k ^ ((uint32_t)(k >> 32) * 0x9E3779B1u) ^ ((uint32_t)(k >> 64) * 0x85EBCA77u) ^
               ((uint32_t)(k >> 96) * 0xC2B2AE3Du);
#endif
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

// 64-bit digest of a key for the per-subject scratch hash tables (identity for 64-bit keys)
GD uint64_t fold_key(hkey k) {
#if GRIMB_KW == 1
  return k;
#else
  return (uint64_t)k ^ mix64((uint64_t)(k >> 64));
#endif
}

GD int popc16(uint32_t m) {
#if GRIMB_DEVICE
  return __popc(m);
#else
  return __builtin_popcount(m);
#endif
}

// Bitonic sorting network that always merges ascending, so virtual +inf padding beyond n can be
// skipped.  `less(i, j)` must be a strict total order over positions; `swp(i, j)` exchanges.
template <class Less, class Swap>
GD void group_sort(const Grp& g, uint32_t n, Less less, Swap swp) {
  if (n < 2) return;
  uint32_t np = 1;
  while (np < n) np <<= 1;
  for (uint32_t k = 2; k <= np; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = g.tid; i < np; i += g.n) {
        uint32_t p = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (p > i && p < n) {
          if (less(p, i)) swp(i, p);
        }
      }
      g.sync();
    }
  }
}

}  // namespace grimb
