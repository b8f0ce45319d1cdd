// Primitives of the in-kernel exchanges over NVLink peer memory (symmetric allocations mapped into every process):
// system-scope acquire/release flags with bounded waits, self-validating 16-byte scalar lines, last-block detection.
// Shared by the fused multi-GPU PCG (peer_pcg.cu) and the multigrid exchanges (mg.cu).
#pragma once
#include "common.cuh"

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}


// Poll a flag in this rank's own block until a peer has stored a sequence number >= want.
static __device__ __noinline__ void wait_flag(const uint64_t* flag, const uint64_t want, uint64_t* err, const int nowait, const uint64_t timeout_ns) {
  if (ld_acquire_sys(flag) >= want || nowait) return;
  if (*reinterpret_cast<volatile uint64_t*>(err)) return;  // an earlier wait already failed: do not stall again
  const uint64_t t0 = global_timer_ns();
  while (ld_acquire_sys(flag) < want) {
    if (global_timer_ns() - t0 > timeout_ns) {
      atomicExch(reinterpret_cast<unsigned long long*>(err), 1ull);
      return;
    }
  }
}

// 16-byte line {lo, seq, hi, seq}: whichever way the fabric splits the store into 8-byte pieces, a reader that sees
// both sequence words equal to the one it expects has both halves of the value.
__device__ __forceinline__ void line_store(uint64_t* line, const double v, const uint32_t seq) {
  const uint64_t b = (uint64_t)__double_as_longlong(v);
  const uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(lo), "r"(seq), "r"(hi), "r"(seq) : "memory");
}
__device__ __forceinline__ bool line_try(const uint64_t* line, const uint32_t seq, double* v) {
  uint32_t lo, f0, hi, f1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(line) : "memory");
  *v = __longlong_as_double((long long)(((uint64_t)hi << 32) | lo));
  return f0 == seq && f1 == seq;
}
static __device__ __noinline__ double line_wait(const uint64_t* line, const uint32_t seq, uint64_t* err, const int nowait, const uint64_t timeout_ns) {
  double v;
  if (line_try(line, seq, &v) || nowait) return v;
  if (*reinterpret_cast<volatile uint64_t*>(err)) return 0.0;
  const uint64_t t0 = global_timer_ns();
  while (!line_try(line, seq, &v)) {
    if (global_timer_ns() - t0 > timeout_ns) {
      atomicExch(reinterpret_cast<unsigned long long*>(err), 1ull);
      return 0.0;
    }
  }
  return v;
}

// True in exactly one block among the `n_blocks` that call it: the one that arrives last.  What thread 0 of a block
// wrote or accumulated (atomics) before the call is visible to that block afterwards; other threads' plain stores are
// only ordered by the kernel boundary (or by their own system fence, for the interface rows).
__device__ __forceinline__ bool arrive_last(uint64_t* ticket_word, const unsigned n_blocks, int* sh) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(ticket_word), 1u);
    *sh = (t == n_blocks - 1) ? 1 : 0;
    __threadfence();
  }
  __syncthreads();
  return *sh != 0;
}

